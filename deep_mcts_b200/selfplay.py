"""BatchedSelfPlay: continuous self-play of thousands of games on one GPU.

Device-resident replacement for the reference's 25 self-play worker processes
(reference deep_mcts/train.py:215-283 driving mcts.py:96-111): every tree runs
``MCTS.self_play`` independently -- ``num_simulations`` simulations per move, move sampled in
proportion to visits for the first ``sample_move_cutoff`` plies then argmax, Dirichlet noise on
every new root, subtree reuse, outcome labelling of the recorded positions, immediate restart
-- and every search round evaluates one leaf per tree in a single network batch.  There is no
host synchronisation inside ``run_rounds``.
"""
from typing import Optional, Tuple

import torch

from . import _lib
from .device_net import DeviceNet
from .pool import TreePool, default_node_capacity


class BatchedSelfPlay:
    def __init__(self, net, n_trees: int, num_simulations: int, sample_move_cutoff: int = 0,
                 dirichlet_alpha: float = 0.0, dirichlet_factor: float = 0.25, replay_capacity: int = 1 << 20,
                 seed: int = 0, max_nodes_per_tree: Optional[int] = None) -> None:
        """``net`` is a GameNet (its DeviceNet is used) or a DeviceNet."""
        self.net = net
        self.device_net: DeviceNet = net if isinstance(net, DeviceNet) else None
        manager = net.manager
        if self.device_net is None:
            net.max_batch = max(net.max_batch, n_trees)
        elif self.device_net.max_batch < n_trees:
            raise ValueError("DeviceNet.max_batch must be >= n_trees")
        cap = max_nodes_per_tree or default_node_capacity(manager, num_simulations)
        self.pool = TreePool(manager, n_trees, cap, seed=seed)
        self.pool.configure(num_simulations, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, sample_move_cutoff,
                            dirichlet_alpha, dirichlet_factor)
        self.pool.enable_selfplay(replay_capacity)
        self.num_simulations = num_simulations
        self.n_trees = n_trees

    def _dn(self) -> DeviceNet:
        return self.device_net if self.device_net is not None else self.net.device_net()

    def run_rounds(self, rounds: int) -> None:
        """``rounds`` x (select one leaf per tree -> network batch -> expand + backup)."""
        dn = self._dn()
        for _ in range(int(rounds)):
            dn.search_round(self.pool, selfplay=True)

    def counters(self) -> dict:
        return self.pool.counters()

    def positions_written(self) -> int:
        return int(self.pool.replay_written.item())

    def sample_batch(self, batch_size: int, generator: Optional[torch.Generator] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """Uniform sample over the replay ring: (first, second, player, pi[B, A], z[B, 1]) on the device
        (reference train.py:337-348 samples uniformly over positions)."""
        written = self.positions_written()
        n = min(written, self.pool.replay_capacity)
        if n == 0:
            raise RuntimeError("replay ring is empty")
        idx = torch.randint(0, n, (batch_size,), device=self.pool.device, generator=generator)
        A = self.pool.num_actions
        return (self.pool.replay_first[idx], self.pool.replay_second[idx], self.pool.replay_player[idx],
                self.pool.replay_pi[idx, :A], self.pool.replay_z[idx].unsqueeze(1))
