"""Flat-array restatement of the reference's MCTS (mcts.py:27-300).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Nodes live in parallel Python lists (visits, value_sum, prior, action,
first_child, n_children) with the children of a node stored contiguously in the
order the reference's dict would iterate them (oracle/setorder.py), so that
first-extremum tie-breaking (mcts.py:167-176) is reproduced by taking the lowest
child slot among equal scores.  All arithmetic is Python float (IEEE double),
evaluated in the reference's operation order:
``u = ((1.25 * P) * sqrt(N_parent)) / (1 + n)`` (mcts.py:40-43).
"""
import random
from math import sqrt
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .games import Game, State

Evaluator = Callable[[Game, State], Tuple[float, Sequence[float]]]
RolloutPolicy = Callable[[Game, State], int]

C_PUCT = 1.25  # mcts.py:42


class OracleMCTS:
    def __init__(
        self,
        game: Game,
        num_simulations: int,
        rollout_policy: Optional[RolloutPolicy],
        state_evaluator: Optional[Evaluator],
        sample_move_cutoff: int = 0,
        dirichlet_alpha: float = 0.0,
        dirichlet_factor: float = 0.25,
        rollout_share: float = 1.0,
    ) -> None:
        if rollout_policy is None and state_evaluator is None:
            # mcts.py:82-83
            raise ValueError("Both rollout_policy and state_evaluator cannot be None")
        self.game = game
        self.num_simulations = num_simulations
        self.rollout_policy = rollout_policy
        self.state_evaluator = state_evaluator
        self.sample_move_cutoff = sample_move_cutoff
        self.dirichlet_alpha = dirichlet_alpha
        self.dirichlet_factor = dirichlet_factor
        self.rollout_share = rollout_share
        self.reset()

    # mcts.py:243-245
    def reset(self) -> None:
        """Fresh tree on the initial state (MCTS.reset, mcts.py:243-245)."""
        self.visits: List[int] = [0]
        self.value_sum: List[float] = [0.0]
        self.prior: List[float] = [-1.0]
        self.action: List[int] = [-1]
        self.first_child: List[int] = [0]
        self.n_children: List[int] = [0]
        self.root = 0
        self.state: State = self.game.initial_state()
        self.leaf_evals = 0
        self.depth_sum = 0
        self.children_scanned = 0

    def set_root_state(self, state: State) -> None:
        """Fresh root on an arbitrary state (MCTSAgent.play's ``Node()`` branch, mcts.py:274-283)."""
        self.visits, self.value_sum, self.prior = [0], [0.0], [-1.0]
        self.action, self.first_child, self.n_children = [-1], [0], [0]
        self.root = 0
        self.state = state

    # ----------------------------------------------------------------- expand
    def _expand(self, node: int, state: State) -> float:
        # mcts.py:198-217
        order = self.game.children_order(state)
        if self.state_evaluator is None:
            value = 0.5
            p = 1 / len(order)
            probs = {a: p for a in order}
        else:
            value, probabilities = self.state_evaluator(self.game, state)
            self.leaf_evals += 1
            assert len(probabilities) == self.game.num_actions
            probs = {a: float(probabilities[a]) for a in order}
        self.first_child[node] = len(self.visits)
        self.n_children[node] = len(order)
        for a in order:
            self.visits.append(0)
            self.value_sum.append(0.0)
            self.prior.append(probs[a])
            self.action.append(a)
            self.first_child.append(0)
            self.n_children.append(0)
        return float(value)

    # ------------------------------------------------------------------ search
    def _tree_search(self) -> Tuple[List[int], State]:
        # mcts.py:162-179 with Node.u (40-43) and Node.value (45-54)
        node = self.root
        state = self.state
        path = [node]
        while self.n_children[node]:
            first, n = self.first_child[node], self.n_children[node]
            sq = sqrt(self.visits[node])
            maximise = state[0] == 0  # SECOND is the max player (game.py:31-33)
            unvisited = 0.0 if maximise else 1.0  # parent.player.loss().value (game.py:25-29)
            best, best_score = -1, 0.0
            for c in range(first, first + n):
                v = self.visits[c]
                value = unvisited if v == 0 else self.value_sum[c] / v
                u = C_PUCT * self.prior[c] * sq / (1 + v)
                score = value + u if maximise else value - u
                if best < 0 or (score > best_score if maximise else score < best_score):
                    best, best_score = c, score
            self.children_scanned += n
            node = best
            path.append(node)
            state = self.game.child(state, self.action[node])
        self.depth_sum += len(path) - 1
        return path, state

    def _rollout(self, state: State) -> float:
        # mcts.py:219-226
        while not self.game.is_final(state):
            state = self.game.child(state, self.rollout_policy(self.game, state))
        return self.game.outcome(state)

    def _evaluate_leaf(self, leaf: int, state: State) -> float:
        # mcts.py:181-196
        if self.game.is_final(state):
            return self.game.outcome(state)
        value = self._expand(leaf, state)
        if self.state_evaluator is None:
            return self._rollout(state)
        if self.rollout_policy is None or (
            self.rollout_share < 1.0 and random.random() >= self.rollout_share
        ):
            return value
        return (self._rollout(state) + value) / 2

    def simulation(self) -> State:
        # mcts.py:150-160, backpropagate 228-231
        path, state = self._tree_search()
        evaluation = self._evaluate_leaf(path[-1], state)
        for node in path:
            self.visits[node] += 1
            self.value_sum[node] += evaluation
        return state

    def _add_dirichlet_noise(self, noise: Optional[Sequence[float]] = None) -> None:
        # mcts.py:233-241
        if self.dirichlet_alpha == 0:
            return
        first, n = self.first_child[self.root], self.n_children[self.root]
        if noise is None:
            noise = np.random.dirichlet([self.dirichlet_alpha] * n)
        for i in range(n):
            c = first + i
            self.prior[c] = self.prior[c] * (1 - self.dirichlet_factor) + noise[i] * self.dirichlet_factor

    def step(self, noise: Optional[Sequence[float]] = None) -> Tuple[List[float], Tuple[int, int]]:
        # mcts.py:113-148 (fixed simulation budget branch, 133-138)
        if not self.n_children[self.root]:
            self._expand(self.root, self.state)
            self.visits[self.root] += 1
        self._add_dirichlet_noise(noise)
        with_exp = without_exp = 0
        for _ in range(self.num_simulations):
            leaf_state = self.simulation()
            if self.game.is_final(leaf_state):
                without_exp += 1
            else:
                with_exp += 1
        first, n = self.first_child[self.root], self.n_children[self.root]
        visit_sum = sum(self.visits[first : first + n])
        dist = [0.0] * self.game.num_actions
        for c in range(first, first + n):
            dist[self.action[c]] = self.visits[c] / visit_sum
        return dist, (with_exp, without_exp)

    def root_visit_counts(self) -> List[int]:
        """Visits of the root's children by action, the numerators of MCTS.step's result (mcts.py:139-147)."""
        first, n = self.first_child[self.root], self.n_children[self.root]
        counts = [0] * self.game.num_actions
        for c in range(first, first + n):
            counts[self.action[c]] = self.visits[c]
        return counts

    def root_children_actions(self) -> List[int]:
        """``list(root.children)``: the children in creation order = set iteration order (mcts.py:200-215)."""
        first, n = self.first_child[self.root], self.n_children[self.root]
        return [self.action[c] for c in range(first, first + n)]

    def advance(self, action: int) -> None:
        """Re-root onto the child for ``action`` keeping its subtree (mcts.py:105-108, 290-291)."""
        first, n = self.first_child[self.root], self.n_children[self.root]
        for c in range(first, first + n):
            if self.action[c] == action:
                self.root = c
                self.state = self.game.child(self.state, action)
                return
        raise KeyError(action)

    def observe(self, state: State) -> None:
        """MCTSAgent.play's re-rooting onto an observed state (mcts.py:267-283)."""
        first, n = self.first_child[self.root], self.n_children[self.root]
        for c in range(first, first + n):
            if self.game.child(self.state, self.action[c]) == state:
                self.root = c
                self.state = state
                return
        self.set_root_state(state)

    def self_play(self):
        # mcts.py:96-111
        i = 0
        while not self.game.is_final(self.state):
            dist, _ = self.step()
            if i < self.sample_move_cutoff:
                action = int(np.random.choice(len(dist), p=dist))
            else:
                action = int(np.argmax(dist))
            old = self.state
            self.advance(action)
            i += 1
            yield old, self.state, action, dist
