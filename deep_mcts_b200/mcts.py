"""Reference-compatible ``MCTS`` / ``MCTSAgent`` on top of the CUDA tree pool.

Keeps the constructor signature, methods and return types of reference deep_mcts/mcts.py
(MCTS 57-245, MCTSAgent 248-300, play_random_mcts 303-316); the search itself runs in the
kernels of csrc/tree.cu.  A plain ``MCTS`` is a pool of one tree (the batch-width-1 shim);
``BatchedMCTS`` drives many trees in lock-step with the same semantics per tree.

Evaluators and rollout policies
  * ``state_evaluator=None`` / ``rollout_policy=None`` keep the reference meaning.
  * The factories below return callables the engine recognises and runs entirely on the
    device: ``uniform_state_evaluator`` (train.py:240-248), ``hashed_state_evaluator`` (test
    evaluator), ``random_rollout_policy`` (``random.choice(legal_actions(s))``),
    ``min_action_policy`` / ``max_action_policy``; a ``GameNet`` evaluator
    (``net.evaluate`` / ``cached_state_evaluator(net)``) runs on the device network.
  * Any other Python callable is honoured through the leaf-exchange path: the device
    selects a leaf, the callable is evaluated on the host, the device expands and backs up.
"""
import random
import time
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._engine import BoardManager, state_to_bits
from .game import Action, GameManager, Player, State
from .pool import TreePool, default_node_capacity
from .tournament import Agent

StateEvaluator = Callable[[State], Tuple[float, Sequence[float]]]
RolloutPolicy = Callable[[State], Action]


# ------------------------------------------------------------------ recognised callables
def uniform_state_evaluator(manager: GameManager) -> StateEvaluator:
    """(0.5, uniform over legal actions) -- the self-play evaluator before the first weight
    transfer (reference train.py:240-248).  Runs on the device when given to MCTS."""

    def evaluator(state):
        legal = set(manager.legal_actions(state))
        return 0.5, [1 / len(legal) if a in legal else 0.0 for a in range(manager.num_actions)]

    evaluator._dm_eval_kind = _lib.DM_EVAL_UNIFORM
    evaluator._dm_eval_salt = 0
    return evaluator


def hashed_state_evaluator(manager: GameManager, salt: int = 0) -> StateEvaluator:
    """Deterministic synthetic evaluator built into the engine (parity tests)."""

    def evaluator(state):
        raise NotImplementedError("hashed_state_evaluator only runs on the device")

    evaluator._dm_eval_kind = _lib.DM_EVAL_HASHED
    evaluator._dm_eval_salt = int(salt)
    return evaluator


def random_rollout_policy(manager: GameManager) -> RolloutPolicy:
    policy = lambda state: random.choice(manager.legal_actions(state))  # noqa: E731
    policy._dm_rollout_kind = _lib.DM_ROLLOUT_RANDOM
    return policy


def greedy_policy_rollout(game_net) -> RolloutPolicy:
    """``lambda state: np.argmax(cached_model(state)[1])`` -- the policy-network rollouts of
    Experiment 3 (reference evaluate_complex_rollouts.py:57-59).  Given to MCTS it runs as batched
    device playouts: every ply is one network batch over all pending leaves plus an argmax."""
    evaluator = game_net.evaluate

    def policy(state):
        return int(np.argmax(evaluator(state)[1]))

    policy._dm_rollout_kind = "greedy_net"
    policy._dm_net = game_net
    return policy


def min_action_policy(manager: GameManager) -> RolloutPolicy:
    policy = lambda state: min(manager.legal_actions(state))  # noqa: E731
    policy._dm_rollout_kind = _lib.DM_ROLLOUT_MIN_ACTION
    return policy


def max_action_policy(manager: GameManager) -> RolloutPolicy:
    policy = lambda state: max(manager.legal_actions(state))  # noqa: E731
    policy._dm_rollout_kind = _lib.DM_ROLLOUT_MAX_ACTION
    return policy


class Node:
    """Read-only view of a tree node (reference mcts.py:27-54 field names)."""

    __slots__ = ["children", "value_sum", "visits", "probability"]

    def __init__(self, visits: int = 0, probability: float = -1.0) -> None:
        self.visits = visits
        self.value_sum = 0.0
        self.probability = probability
        self.children: Dict[Action, "Node"] = {}


class BatchedMCTS:
    """``n_trees`` independent searches with MCTS.step semantics per tree."""

    def __init__(self, game_manager: BoardManager, n_trees: int, num_simulations: int,
                 rollout_policy: Optional[RolloutPolicy], state_evaluator: Optional[StateEvaluator],
                 sample_move_cutoff: int = 0, dirichlet_alpha: float = 0.0, dirichlet_factor: float = 0.25,
                 rollout_share: float = 1.0, max_nodes_per_tree: Optional[int] = None, seed: int = 0,
                 leaves_per_round: int = 1) -> None:
        if rollout_policy is None and state_evaluator is None:
            raise ValueError("Both rollout_policy and state_evaluator cannot be None")
        self.leaves_per_round = int(leaves_per_round)
        self.game_manager = game_manager
        self.num_simulations = num_simulations
        self.rollout_policy = rollout_policy
        self.state_evaluator = state_evaluator
        self.sample_move_cutoff = sample_move_cutoff
        self.dirichlet_alpha = dirichlet_alpha
        self.dirichlet_factor = dirichlet_factor
        self.rollout_share = rollout_share
        sims_for_cap = num_simulations if num_simulations != float("inf") else 4096
        cap = max_nodes_per_tree or default_node_capacity(game_manager, int(sims_for_cap))
        self.pool = TreePool(game_manager, n_trees, cap, seed=seed)
        self._classify()

    # which device path serves the two callables
    def _classify(self) -> None:
        ev, ro = self.state_evaluator, self.rollout_policy
        self._host_eval = None
        self._host_rollout = None
        self._net = None
        if ev is None:
            self._eval_kind, self._salt = _lib.DM_EVAL_NONE, 0
        elif hasattr(ev, "_dm_eval_kind"):
            self._eval_kind, self._salt = ev._dm_eval_kind, getattr(ev, "_dm_eval_salt", 0)
        else:
            self._eval_kind, self._salt = _lib.DM_EVAL_EXTERNAL, 0
            net = getattr(ev, "_dm_net", None) or getattr(getattr(ev, "__self__", None), "_dm_net_self", None)
            if net is not None and getattr(net, "on_device", False):
                self._net = net
            else:
                self._host_eval = ev
        self._greedy_net = None
        if ro is None:
            self._rollout_kind = _lib.DM_ROLLOUT_NONE
        elif getattr(ro, "_dm_rollout_kind", None) == "greedy_net":
            # device policy-network playouts: leaves go through the exchange path, values are mixed on the device
            self._rollout_kind = _lib.DM_ROLLOUT_NONE
            self._greedy_net = ro._dm_net
            if self._eval_kind not in (_lib.DM_EVAL_NONE, _lib.DM_EVAL_EXTERNAL) or self._host_eval is not None:
                raise ValueError("policy-network rollouts combine with state_evaluator=None or a GameNet evaluator")
            self._eval_kind = _lib.DM_EVAL_EXTERNAL
        elif hasattr(ro, "_dm_rollout_kind"):
            self._rollout_kind = ro._dm_rollout_kind
        else:
            # arbitrary Python policy: play it out on the host inside the leaf exchange
            self._rollout_kind = _lib.DM_ROLLOUT_NONE
            self._host_rollout = ro
            if self._eval_kind != _lib.DM_EVAL_EXTERNAL:
                self._host_eval = self.state_evaluator or None
                self._eval_kind = _lib.DM_EVAL_EXTERNAL
        if (self._net is not None and getattr(ev, "_dm_net", None) is not None and self._greedy_net is None
                and self._host_rollout is None and self.leaves_per_round == 1):
            # cached_state_evaluator(net) (reference train.py:408-414): memoise on the device.  (The policy-network
            # rollout path and host rollouts drive the leaf exchange themselves and evaluate every leaf.)
            work = self.pool.n_trees * (int(self.num_simulations) if self.num_simulations != float("inf") else 4096)
            self.pool.enable_cache(min(22, max(14, (8 * work).bit_length())), torch.float32)
            self._cache_weights_version = None

    def _configure(self, sims: int, alpha: Optional[float] = None) -> None:
        vl = self.leaves_per_round if self._eval_kind == _lib.DM_EVAL_EXTERNAL and self._greedy_net is None else 1
        if vl > 1 and self._rollout_kind != _lib.DM_ROLLOUT_NONE:
            # the virtual-loss kernels back up the evaluator's value only (no evaluate_leaf mixing, mcts.py:187-196)
            raise ValueError("leaves_per_round > 1 (virtual-loss mode) does not support rollout policies")
        self.pool.configure(sims, self._eval_kind, self._rollout_kind, self.sample_move_cutoff,
                            self.dirichlet_alpha if alpha is None else alpha, self.dirichlet_factor,
                            self.rollout_share, self._salt, vl)

    # ---- host leaf exchange -----------------------------------------------------------
    def _host_evaluate(self, player, first, second):
        m = self.game_manager
        values, priors = [], []
        net_values = net_probs = None
        if self._host_eval is None and self._net is not None:
            # GameNet evaluator next to a Python rollout policy: the network still runs as one device batch
            net_values, net_probs = self._net.evaluate_batch(player, first, second)
        for i, (p, f, s) in enumerate(zip(player, first, second)):
            state = m._make_state(p, f, s)
            if net_values is not None:
                value, probs = float(net_values[i]), [float(x) for x in net_probs[i]]
            elif self._host_eval is not None:
                value, probs = self._host_eval(state)
                probs = [float(x) for x in probs]
                assert len(probs) == m.num_actions
            else:  # state_evaluator=None with a host rollout policy (mcts.py:202-212)
                legal = set(m.legal_actions(state))
                value, probs = 0.5, [1 / len(legal) if a in legal else 0.0 for a in range(m.num_actions)]
            value = float(value)
            if self._host_rollout is not None:
                # evaluate_leaf's mixing rule (mcts.py:187-196) with the rollout played on the host
                if self.state_evaluator is None:
                    value = self._play_out(state)
                elif not (self.rollout_share < 1.0 and random.random() >= self.rollout_share):
                    value = (self._play_out(state) + value) / 2
            values.append(value)
            priors.append(probs)
        return np.asarray(values, dtype=np.float64), np.asarray(priors, dtype=np.float64)

    def _play_out(self, state) -> float:
        m = self.game_manager
        while not m.is_final_state(state):
            state = m.generate_child_state(state, self._host_rollout(state))
        return m.evaluate_final_state(state).value

    # ---- one MCTS.step for every tree -------------------------------------------------
    def run_simulations(self, sims: int, alpha: Optional[float] = None, tree_ids=None) -> None:
        self._configure(sims, alpha)
        self.pool.begin_step(tree_ids)
        if self._greedy_net is not None:
            self._run_greedy_rollout_search(sims)
        elif self._eval_kind != _lib.DM_EVAL_EXTERNAL:
            self.pool.search_local()
        elif self._net is not None and self._host_rollout is None:
            if self.pool.cache_entries:
                version = self._net.device_net().weights_version
                if version != self._cache_weights_version:      # the network was re-trained: forget its old outputs
                    self.pool.cache_clear()
                    self._cache_weights_version = version
            self._net.run_search(self.pool, sims)
        else:
            self.pool.step_external(self._host_evaluate, max_rounds=sims + 2)

    def _run_greedy_rollout_search(self, sims: int) -> None:
        """select -> (network evaluation | uniform priors) -> greedy policy playouts -> expand/backup.
        evaluate_leaf's mixing rule with rollout_share = 1 (mcts.py:187-196): the playout value alone
        without a state evaluator, the mean of playout and network value with one."""
        pool = self.pool
        rnet = self._greedy_net.device_net()
        enet = self._net.device_net() if self._net is not None else None
        T = pool.n_trees
        if rnet.max_batch < T or (enet is not None and enet.max_batch < T):
            raise ValueError("GameNet.max_batch must be >= the number of trees")
        if not hasattr(self, "_rollout_out"):
            self._rollout_out = torch.zeros(T, dtype=torch.float32, device=pool.device)
            self._uniform_value = torch.zeros(T, dtype=torch.float32, device=pool.device)
            self._uniform_probs = torch.zeros((T, _lib.DM_MAX_ACTIONS), dtype=torch.float32, device=pool.device)
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(int(sims) + 1):
            pool.select()
            n = int(pool.leaf_count.item())
            if n == 0:
                break
            if enet is not None:
                value, probs = enet.evaluate_device(pool.leaf_first, pool.leaf_second, pool.leaf_player, T,
                                                    pool.leaf_count)
                probs = probs[:T].clone()   # the playouts below reuse the network's output buffers
                value = value[:T].clone()
            else:
                with torch.cuda.device(pool.device):
                    _lib.call("dm_eval_uniform", pool.manager.game_id, pool.manager.board, T,
                              pool.leaf_count.data_ptr(), pool.leaf_first.data_ptr(), pool.leaf_second.data_ptr(),
                              pool.leaf_player.data_ptr(), self._uniform_value.data_ptr(),
                              self._uniform_probs.data_ptr(), stream)
                value, probs = self._uniform_value, self._uniform_probs
            rnet.rollout_greedy_device(pool.leaf_first, pool.leaf_second, pool.leaf_player, n, self._rollout_out)
            mixed = self._rollout_out if enet is None else (self._rollout_out + value) / 2
            pool.expand_backup(probs, mixed.contiguous())

    def step(self, tree_ids=None) -> Tuple[np.ndarray, np.ndarray]:
        """(visits[n_trees, A], stats[n_trees, 2]) after ``num_simulations`` simulations per tree
        (only on ``tree_ids`` when given; the other trees are left untouched)."""
        self.run_simulations(int(self.num_simulations), tree_ids=tree_ids)
        visits, stats = self.pool.root_visits()
        self.pool.counters()  # surfaces capacity errors
        return visits, stats

    def advance(self, actions) -> None:
        self.pool.advance(actions)

    def reset(self, tree_ids=None) -> None:
        self.pool.reset(tree_ids)


class MCTS(BatchedMCTS):
    """Drop-in for the reference's ``MCTS`` (one tree)."""

    def __init__(self, game_manager: BoardManager, num_simulations: int, rollout_policy: Optional[RolloutPolicy],
                 state_evaluator: Optional[StateEvaluator], sample_move_cutoff: int = 0,
                 dirichlet_alpha: float = 0.0, dirichlet_factor: float = 0.25, rollout_share: float = 1.0,
                 time_per_move: float = float("inf"), max_nodes_per_tree: Optional[int] = None,
                 seed: int = 0, leaves_per_round: int = 1) -> None:
        super().__init__(game_manager, 1, num_simulations, rollout_policy, state_evaluator, sample_move_cutoff,
                         dirichlet_alpha, dirichlet_factor, rollout_share, max_nodes_per_tree, seed, leaves_per_round)
        self.time_per_move = time_per_move
        self.state = game_manager.initial_game_state()

    # ---- reference attributes -----------------------------------------------------------
    @property
    def root(self) -> Node:
        order, count, root_visits = self.pool.root_children()
        visits, _ = self.pool.root_visits()
        prior, value_sum = self.pool.root_child_stats()
        node = Node(int(root_visits[0]))
        for j, a in enumerate(order[0, : int(count[0])]):
            child = Node(int(visits[0, int(a)]), float(prior[0, j]))
            child.value_sum = float(value_sum[0, j])
            node.children[int(a)] = child
        return node

    def step(self) -> Tuple[List[float], Tuple[int, int]]:
        """One move's search (reference mcts.py:113-148)."""
        with_exp = without_exp = 0
        if self.time_per_move < float("inf"):
            start = time.perf_counter()
            budget = self.num_simulations
            first_chunk = True
            while time.perf_counter() - start < self.time_per_move and with_exp + without_exp < budget:
                chunk = int(min(64, budget - (with_exp + without_exp)))
                self.run_simulations(chunk, alpha=None if first_chunk else 0.0)
                first_chunk = False
                _, stats = self.pool.root_visits()
                with_exp += int(stats[0, 0])
                without_exp += int(stats[0, 1])
        else:
            self.run_simulations(int(self.num_simulations))
            _, stats = self.pool.root_visits()
            with_exp, without_exp = int(stats[0, 0]), int(stats[0, 1])
        visits, _ = self.pool.root_visits()
        self.pool.counters()
        counts = [int(v) for v in visits[0]]
        total = sum(counts)
        return [c / total for c in counts], (with_exp, without_exp)

    def advance(self, action: Action) -> None:
        """``root = root.children[action]; state = child state`` (mcts.py:105-108)."""
        self.pool.advance([int(action)])
        self.state = self.game_manager.generate_child_state(self.state, int(action))

    def observe(self, state) -> None:
        """Re-root onto an observed state (MCTSAgent.play, mcts.py:267-283)."""
        p, f, s = state_to_bits(state)
        self.pool.observe([p], [f], [s])
        self.state = state

    def self_play(self) -> Iterable[Tuple[State, State, Action, Sequence[float]]]:
        i = 0
        while not self.game_manager.is_final_state(self.state):
            probabilities, _ = self.step()
            if i < self.sample_move_cutoff:
                action = int(np.random.choice(len(probabilities), p=probabilities))
            else:
                action = int(np.argmax(probabilities))
            old_state = self.state
            self.advance(action)
            i += 1
            yield old_state, self.state, action, probabilities

    def reset(self) -> None:
        self.pool.reset()
        self.state = self.game_manager.initial_game_state()


class MCTSAgent(Agent):
    """Tree-reusing player (reference mcts.py:248-300)."""

    def __init__(self, mcts: MCTS, epsilon: float = 0.0,
                 reset_fn: Optional[Callable[["MCTSAgent"], None]] = None) -> None:
        self.mcts = mcts
        self.epsilon = epsilon
        self.current_game_simulation_stats: List[Tuple[int, int]] = []
        self.simulation_stats = [self.current_game_simulation_stats]
        self.reset_fn = reset_fn

    def play(self, state) -> Action:
        self.mcts.observe(state)
        probabilities, simulations = self.mcts.step()
        self.current_game_simulation_stats.append(simulations)
        if self.epsilon > 0 and random.random() < self.epsilon:
            action = int(np.random.choice(len(probabilities), p=probabilities))
        else:
            action = int(np.argmax(probabilities))
        self.mcts.advance(action)
        return action

    def reset(self) -> None:
        self.mcts.reset()
        self.current_game_simulation_stats = []
        self.simulation_stats.append(self.current_game_simulation_stats)
        if self.reset_fn is not None:
            self.reset_fn(self)


def play_random_mcts(manager: BoardManager, num_simulations: int) -> None:
    mcts = MCTS(manager, num_simulations, random_rollout_policy(manager), None)
    next_state = mcts.state
    for state, next_state, action, _ in mcts.self_play():
        print(state, "\n", manager.action_str(action), "\n", sep="\n")
    print(next_state)
