"""Training loop (Experiment 1) on the batched engine.

Interface of reference deep_mcts/train.py: ``TrainingConfiguration`` (38-68), ``train`` (71-108),
``evaluate`` (351-405), ``cached_state_evaluator`` (408-414), with the same on-disk artefacts
(``anet-<iteration>.tar``, ``parameters.json``, ``evaluations.csv``).

What changes underneath: the 25 self-play worker processes, the ``mp.Queue`` and the CUDA-IPC
shared network (train.py:128-134, 215-334) are replaced by ``BatchedSelfPlay`` -- thousands of
games searched on the device, examples labelled on the device and appended to a device replay
ring -- and by one process per GPU whose gradients are averaged with a single flat NCCL
all-reduce (``distributed.allreduce_gradients``).  Like the reference, games are played with a
uniform evaluator until the first weight transfer (train.py:240-257) and the self-play network
is refreshed every ``transfer_interval`` iterations (169-171).

Because device self-play is orders of magnitude faster than the reference's workers, the ratio
of training iterations to self-play is explicit here (``train_steps_per_block``) instead of
being whatever two free-running processes happened to produce.  Like the reference's two
processes, the two halves run concurrently: the SGD steps of a block are issued on their own
stream while the next self-play block is searched (``overlap_training``), sampling only replay
positions that were complete when the block started.
"""
import json
import random
import time
from dataclasses import asdict, dataclass
from typing import Any, Dict, Iterable, Optional, Tuple

import torch

from . import _lib, distributed as dd
from .gamenet import GameNet, cached_state_evaluator  # noqa: F401  (re-exported like the reference)
from .mcts import MCTS, MCTSAgent, RolloutPolicy, random_rollout_policy
from .selfplay import BatchedSelfPlay
from .tournament import AgentComparison, compare_agents

EVALUATION_COLUMNS = [
    "training_iterations", "training_games",
    "random_first_wins", "random_second_wins", "random_first_draws", "random_second_draws",
    "random_first_losses", "random_second_losses",
    "previous_first_wins", "previous_second_wins", "previous_first_draws", "previous_second_draws",
    "previous_first_losses", "previous_second_losses",
]


@dataclass(frozen=True)
class TrainingConfiguration:
    num_games: int
    num_simulations: int
    save_interval: int
    evaluation_interval: int
    save_dir: str
    sample_move_cutoff: int
    dirichlet_alpha: float
    dirichlet_factor: float = 0.25
    rollout_policy: Optional[RolloutPolicy] = None
    epsilon: float = 0.05
    nprocs: int = 25
    batch_size: int = 1024
    replay_buffer_max_size: int = 100_000
    train_device: torch.device = torch.device("cuda:1")
    self_play_device: torch.device = torch.device("cuda:0")
    evaluation_games: int = 20
    transfer_interval: int = 1000
    # ---- engine knobs (not in the reference) ----
    n_trees: int = 4096              # concurrent self-play games per GPU
    train_steps_per_block: int = 8   # SGD steps after every block of num_simulations search rounds
    uniform_until_first_transfer: bool = True
    precision: Optional[str] = None
    overlap_training: bool = True    # SGD steps run on their own stream while the next self-play block is searched
    # replay the SGD step (forward, backward, all-reduce, update) as one CUDA graph.  Off by default: measured +2 % at
    # 8 GPUs (batch 128 per GPU), nothing at 1 GPU -- the step is cuDNN time, not launch latency
    graph_train_step: bool = False

    def to_json_dict(self) -> Dict[str, Any]:
        d = asdict(self)
        d["has_rollout_policy"] = d.pop("rollout_policy") is not None
        for key in ("train_device", "self_play_device"):
            dev = d[key]
            d[key] = f"{dev.type}{'' if dev.index is None else ':' + str(dev.index)}"
        return d


def bitboards_to_planes(first: torch.Tensor, second: torch.Tensor, player: torch.Tensor, g: int,
                        transpose_second: bool) -> torch.Tensor:
    """Device twin of ``states_to_tensor``: [B, 3, g, g] fp32 planes from bitboard tensors."""
    shifts = torch.arange(g * g, device=first.device, dtype=torch.int64)
    f = ((first.unsqueeze(1) >> shifts) & 1).reshape(-1, g, g)
    s = ((second.unsqueeze(1) >> shifts) & 1).reshape(-1, g, g)
    is_first = (player == 1).view(-1, 1, 1)
    own = torch.where(is_first, f, s)
    opp = torch.where(is_first, s, f)
    if transpose_second:
        second_to_move = (player == 0)
        own = torch.where(second_to_move.view(-1, 1, 1), own.transpose(1, 2), own)
        opp = torch.where(second_to_move.view(-1, 1, 1), opp.transpose(1, 2), opp)
    plane = player.float().view(-1, 1, 1).expand(-1, g, g)
    return torch.stack([own.float(), opp.float(), plane], dim=1)


class SelfPlayTrainer:
    """One rank's share of the loop: device self-play + SGD with gradient all-reduce."""

    def __init__(self, game_net: GameNet, config: TrainingConfiguration) -> None:
        self.net = game_net
        self.config = config
        self.rank, self.world = dd.world()
        manager = game_net.manager
        self.device = manager.device
        game_net.to(self.device)
        dd.broadcast_parameters(game_net.net, 0)
        game_net.grad_sync = dd.allreduce_gradients
        # gradients as views into one flat buffer: the data-parallel step is one in-place all-reduce
        game_net.grad_bucket = dd.GradientBucket(game_net.net.parameters())
        game_net.precision = config.precision
        self.transpose = manager.game_id in (_lib.DM_HEX, _lib.DM_HEX_SWAP)
        positions = config.replay_buffer_max_size * (manager.board * manager.board)
        self.engine = BatchedSelfPlay(game_net, config.n_trees, config.num_simulations, config.sample_move_cutoff,
                                      config.dirichlet_alpha, config.dirichlet_factor,
                                      replay_capacity=min(positions, 1 << 22), seed=17 + self.rank)
        self.uniform_phase = config.uniform_until_first_transfer
        self.iterations = 0
        self.allreduce_ms = 0.0
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(1234 + self.rank)
        self.engine.use_graph = True
        self.sp_stream = torch.cuda.Stream(device=self.device) if config.overlap_training else None
        self._train_graph = None         # CUDA graph of one whole SGD step (see _sgd)
        self._train_stream = None
        self._eager_steps = 0

    # ---- self-play -------------------------------------------------------------------
    def _uniform_round(self) -> None:
        """One search round with the uniform evaluator (train.py:240-248) instead of the network."""
        pool = self.engine.pool
        dn = self.engine.device_nets[0]
        pool.selfplay_select()
        with torch.cuda.device(self.device):
            _lib.call("dm_eval_uniform", pool.manager.game_id, pool.manager.board, pool.n_trees,
                      pool.leaf_count.data_ptr(), pool.leaf_first.data_ptr(), pool.leaf_second.data_ptr(),
                      pool.leaf_player.data_ptr(), dn._value.data_ptr(), dn._probs.data_ptr(),
                      torch.cuda.current_stream().cuda_stream)
        pool.expand_backup(dn._probs, dn._value)

    def self_play_block(self) -> None:
        rounds = self.config.num_simulations
        if self.uniform_phase:
            for _ in range(rounds):
                self._uniform_round()
        else:
            self.engine.run_rounds(rounds)

    def have_examples(self) -> bool:
        """Collective: True once EVERY rank has replay positions to train on (ranks must take the same number
        of all-reduced SGD steps, so the decision cannot be rank-local)."""
        mine = 1.0 if self.engine.positions_written() > 0 else 0.0
        sums, _ = dd.reduce_counters({"have": mine}, {}, self.device)
        return sums["have"] >= self.world

    def block(self, train_steps: int):
        """One block of the loop: ``num_simulations`` search rounds for every game and ``train_steps`` SGD steps.
        With ``overlap_training`` the two run concurrently on two streams.  The steps sample through a snapshot of
        the replay ring taken BEFORE the block's self-play is launched -- positions complete at that moment, minus the
        ring slots the block can overwrite (at most one finished game = DM_MAX_DEPTH positions per tree) -- so the SGD
        stream never reads a slot the search is writing and never synchronises with it; the weight transfer happens at
        the block boundary, when both streams are idle.
        Returns the list of (policy_loss, value_loss, accuracy) of the steps taken."""
        if self.sp_stream is None or self.uniform_phase:
            self.self_play_block()
            return [self.train_step() for _ in range(train_steps)] if self.have_examples() else []
        main = torch.cuda.current_stream()
        have = self.have_examples()
        snapshot = self.snapshot() if have else None
        # a ring too small to keep a protected range (tiny test configurations): gather the batches before self-play starts
        batches = [self.sample() for _ in range(train_steps)] if have and not snapshot else None
        self.sp_stream.wait_stream(main)
        with torch.cuda.stream(self.sp_stream):
            self.self_play_block()
        losses = []
        for i in range(train_steps if have else 0):
            batch = batches[i] if batches is not None else \
                self.engine.sample_batch(self.config.batch_size, self.generator, snapshot=snapshot)
            losses.append(self.train_step(batch, defer_transfer=True))
        main.wait_stream(self.sp_stream)
        if self._transfer_due:
            self._transfer_weights()
        return losses

    # ---- training --------------------------------------------------------------------
    def snapshot(self):
        """Replay ranges the block's SGD steps may read while its self-play runs (see ``block``)."""
        return self.engine.replay_snapshot(max(1, self.config.replay_buffer_max_size // self.world),
                                           protect=self.config.n_trees * _lib.DM_MAX_DEPTH)

    def sample(self):
        """One training batch from the last ``replay_buffer_max_size`` games (train.py:149-151, 337-348);
        every rank holds its share of that window."""
        return self.engine.sample_batch(self.config.batch_size, self.generator,
                                        max_games=max(1, self.config.replay_buffer_max_size // self.world))

    def _sgd(self, planes: torch.Tensor, pi: torch.Tensor, z: torch.Tensor):
        """``GameNet.train`` on one batch.  The step is ~150 small kernels (batch 1024 / world on a 1.6 M-parameter
        net) and launch-bound when issued eagerly, so after a few eager steps (optimizer state allocated, cuDNN
        autotuned) the whole step -- forward, backward, the in-place gradient all-reduce, the SGD update -- is
        captured once into a CUDA graph over static input buffers and replayed from then on."""
        if not self.config.graph_train_step or not planes.is_cuda:
            return self.net.train(planes, pi, z)
        if self._train_graph is None:
            current = torch.cuda.current_stream()
            if self._train_stream is None:
                self._train_stream = torch.cuda.Stream(device=self.device)
            if self._eager_steps < 3 or planes.shape[0] != self.config.batch_size:
                # warm-up steps run on the stream the capture will use: autograd's gradient accumulators belong to
                # the stream of the first forward pass, and a capture that has to synchronise with another stream
                # for them is invalidated
                self._eager_steps += 1
                self._train_stream.wait_stream(current)
                with torch.cuda.stream(self._train_stream):
                    losses = self.net.train(planes, pi, z)
                current.wait_stream(self._train_stream)
                return losses
            self._static = (planes.clone(), pi.clone(), z.clone())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=self._train_stream):
                self._static_losses = self.net.train(*self._static)
            self._train_graph = graph
        for dst, src in zip(self._static, (planes, pi, z)):
            dst.copy_(src)
        self._train_graph.replay()
        self.net._device_net_stale = True
        return tuple(t.clone() for t in self._static_losses)

    def train_step(self, batch=None, defer_transfer: bool = False):
        first, second, player, pi, z = batch if batch is not None else self.sample()
        planes = bitboards_to_planes(first, second, player, self.net.manager.board, self.transpose)
        if pi.shape[1:] != tuple(self.net.net.policy_head.out_shape):
            pi = pi.reshape(pi.shape[0], *self.net.net.policy_head.out_shape)
        losses = self._sgd(planes, pi, z)
        self.iterations += 1
        if self.iterations % self.config.transfer_interval == 0:
            if defer_transfer:
                self._transfer_due = True    # the self-play stream is still using the current weights
            else:
                self._transfer_weights()
        return losses

    _transfer_due = False

    def _transfer_weights(self) -> None:
        """Weight transfer to the self-play kernels (train.py:169-171).  Parameters are identical on every rank
        (same start, same averaged gradients); the BatchNorm running statistics are not -- each rank normalises
        its own shard of the batch -- so they are averaged here, which keeps the replicas' self-play networks
        identical."""
        dd.allreduce_buffers(self.net.net)
        self.engine.sync_weights()
        self.uniform_phase = False
        self._transfer_due = False

    def games_finished(self) -> int:
        """Games finished on ALL ranks (collective)."""
        sums, _ = dd.reduce_counters({"games": self.engine.games_written()}, {}, self.device)
        return int(sums["games"])


def _train(game_net: GameNet, config: TrainingConfiguration) -> Iterable[Tuple[int, int, AgentComparison, AgentComparison]]:
    trainer = SelfPlayTrainer(game_net, config)
    rank, world = trainer.rank, trainer.world
    if rank == 0:
        game_net.save_full(f"{config.save_dir}/anet-0.tar")
        with open(f"{config.save_dir}/parameters.json", "w") as f:
            net_parameters = game_net.parameters()
            net_parameters.pop("state_dict")
            net_parameters["optimizer_cls"] = net_parameters["optimizer_cls"].__name__
            json.dump({"net_parameters": net_parameters, "training_configuration": config.to_json_dict()}, f, indent=2)
    previous = game_net.copy().to(trainer.device)
    # the reference plays num_games per worker process; the stop condition is collective (the summed game count),
    # so every rank runs the same number of blocks and of all-reduced SGD steps
    target_games = max(1, config.num_games * config.nprocs)
    policy_loss = value_loss = acc = 0.0
    games = 0
    while games < target_games:
        losses = trainer.block(config.train_steps_per_block)
        games = trainer.games_finished()
        for pl, vl, ac in losses:
            policy_loss, value_loss, acc = (policy_loss + pl.detach().item(), value_loss + vl.detach().item(),
                                             acc + ac.detach().item())
            it = trainer.iterations
            if rank == 0 and config.save_interval and it % config.save_interval == 0:
                game_net.save_full(f"{config.save_dir}/anet-{it}.tar")
            if config.evaluation_interval and it % config.evaluation_interval == 0:
                if rank == 0:
                    random_eval, previous_eval = evaluate(game_net, previous, game_net.manager, config)
                    print(f"{time.strftime('%H:%M:%S')} iterations: {it} games: {games} "
                          f"previous: {previous_eval} random MCTS: {random_eval} "
                          f"policy_loss: {policy_loss / config.evaluation_interval:.2f} "
                          f"value_loss: {value_loss / config.evaluation_interval:.2f} "
                          f"accuracy: {acc / config.evaluation_interval:.2f}")
                    yield it - 1, games, random_eval, previous_eval
                dd.barrier()    # the other ranks wait here, not inside the next gradient all-reduce
                policy_loss = value_loss = acc = 0.0
                previous.load_state_dict(game_net.net.state_dict())


def train(game_net: GameNet, config: TrainingConfiguration) -> None:
    """Run the loop and write ``evaluations.csv`` with the reference's columns (train.py:72-108)."""
    import pandas as pd
    rows = [[it, games, *[*r[0], *r[1], *r[2]], *[*p[0], *p[1], *p[2]]] for it, games, r, p in _train(game_net, config)]
    if dd.world()[0] == 0:
        pd.DataFrame(rows, columns=EVALUATION_COLUMNS).to_csv(f"{config.save_dir}/evaluations.csv")


def evaluate(game_net: GameNet, previous_game_net: GameNet, game_manager, config: TrainingConfiguration
             ) -> Tuple[AgentComparison, AgentComparison]:
    """``evaluation_games`` games against a random-rollout MCTS with twice the simulations and against
    the previous evaluation's network (reference train.py:351-405)."""
    evaluator = cached_state_evaluator(game_net)
    previous_evaluator = cached_state_evaluator(previous_game_net)
    sims = config.num_simulations
    random_eval = compare_agents(
        (MCTSAgent(MCTS(game_manager, sims, config.rollout_policy, evaluator)),
         MCTSAgent(MCTS(game_manager, sims * 2, random_rollout_policy(game_manager), None))),
        config.evaluation_games, game_manager)
    previous_eval = compare_agents(
        (MCTSAgent(MCTS(game_manager, sims, config.rollout_policy, evaluator), epsilon=config.epsilon),
         MCTSAgent(MCTS(game_manager, sims, config.rollout_policy, previous_evaluator), epsilon=config.epsilon)),
        config.evaluation_games, game_manager)
    return random_eval, previous_eval
