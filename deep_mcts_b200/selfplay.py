"""BatchedSelfPlay: continuous self-play of thousands of games on one GPU.

Device-resident replacement for the reference's 25 self-play worker processes
(reference deep_mcts/train.py:215-283 driving mcts.py:96-111): every tree runs
``MCTS.self_play`` independently -- ``num_simulations`` simulations per move, move sampled in
proportion to visits for the first ``sample_move_cutoff`` plies then argmax, Dirichlet noise on
every new root, subtree reuse, outcome labelling of the recorded positions, immediate restart
-- and every search round evaluates one leaf per tree in a single network batch.  There is no
host synchronisation inside ``run_rounds``.
"""
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from .device_net import DeviceNet
from .pool import TreePool, default_node_capacity


class BatchedSelfPlay:
    def __init__(self, net, n_trees: int, num_simulations: int, sample_move_cutoff: int = 0,
                 dirichlet_alpha: float = 0.0, dirichlet_factor: float = 0.25, replay_capacity: int = 1 << 20,
                 seed: int = 0, max_nodes_per_tree: Optional[int] = None, n_pools: int = 1,
                 pool_sizes: Optional[Sequence[int]] = None) -> None:
        """``net`` is a GameNet (its weights are used) or a DeviceNet.

        ``n_pools`` > 1 splits the games into independent pools, each with its own network
        buffers and CUDA stream.  The pools' rounds are dependency chains of their own, so the
        latency-bound tree kernels of one pool run under the tensor-bound network kernels of
        another (no collective, no data shared between pools).  ``pool_sizes`` gives the games per
        pool explicitly (default: an even split); sizes that are multiples of 4 x the SM count
        keep every trunk launch at whole waves of work items."""
        if pool_sizes is None:
            if n_trees % n_pools:
                raise ValueError("n_trees must be divisible by n_pools")
            pool_sizes = [n_trees // n_pools] * n_pools
        pool_sizes = [int(x) for x in pool_sizes]
        if len(pool_sizes) != n_pools or sum(pool_sizes) != n_trees or min(pool_sizes) <= 0:
            raise ValueError("pool_sizes must be n_pools positive sizes that add up to n_trees")
        self.net = net
        manager = net.manager
        self.pool_sizes = pool_sizes
        cap = max_nodes_per_tree or default_node_capacity(manager, num_simulations)
        self.num_simulations = num_simulations
        self.n_trees = n_trees
        self.n_pools = n_pools
        self.pools = []
        self.device_nets = []
        for i in range(n_pools):
            pool = TreePool(manager, pool_sizes[i], cap, seed=seed * 1000 + i)
            pool.configure(num_simulations, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, sample_move_cutoff,
                           dirichlet_alpha, dirichlet_factor)
            pool.enable_selfplay(max(1, replay_capacity * pool_sizes[i] // n_trees))
            self.pools.append(pool)
        if isinstance(net, DeviceNet):
            if n_pools != 1:
                raise ValueError("pass a GameNet when n_pools > 1 (each pool needs its own network buffers)")
            if net.max_batch < pool_sizes[0]:
                raise ValueError("DeviceNet.max_batch must be >= trees per pool")
            self.device_nets = [net]
        else:
            self._build_device_nets()
        self.pool = self.pools[0]
        self.tree_timers = None     # set to a list to collect CUDA-event pairs around the tree kernels
        self.use_graph = False      # set True to replay one captured round instead of re-launching it
        self._graph = None
        self._phase_offset = False  # pools[1:] hold a pending selection (pipelined graph mode)
        self._graph_launches = 0
        self.graph_replays = 0
        self.streams = [torch.cuda.Stream(device=self.pool.device) for _ in range(n_pools)] if n_pools > 1 else [None]

    def _build_device_nets(self) -> None:
        net = self.net
        sd = net.net.state_dict()
        self.device_nets = []
        for size in self.pool_sizes:
            dn = DeviceNet(net.manager, net.channels, net.num_residual, net.value_head_hidden_units, size,
                           net.precision)
            dn.load_state_dict(sd)
            self.device_nets.append(dn)

    def sync_weights(self) -> None:
        """Re-pack the GameNet's current weights into every pool's kernels (weight transfer)."""
        if not isinstance(self.net, DeviceNet):
            sd = self.net.net.state_dict()
            for dn in self.device_nets:
                dn.load_state_dict(sd)

    def run_rounds(self, rounds: int) -> None:
        """``rounds`` search rounds for every pool.  A single pool runs fused rounds (expand + backup of the
        previous batch and the next selection in one tree kernel, then the network batch), so the
        last batch stays pending between calls; several pools run select -> network -> expand/backup."""
        if self.n_pools == 1:
            dn, pool = self.device_nets[0], self.pools[0]
            if self.use_graph:
                if self._graph is None:
                    self._capture_round()
                for _ in range(int(rounds)):
                    self._graph.replay()
                self.graph_replays += int(rounds)
                return
            for _ in range(int(rounds)):
                dn.search_round(pool, selfplay=True, timers=self.tree_timers, fused=True)
            return
        if self.use_graph:
            if self._graph is None:
                self._capture_pipelined_round()
            if not self._phase_offset:
                for pool in self.pools[1:]:
                    pool.selfplay_select()          # prime the phase offset: a selection is pending
                self._phase_offset = True
            for _ in range(int(rounds)):
                self._graph.replay()
            self.graph_replays += int(rounds)
            return
        if self._phase_offset:
            for dn, pool in zip(self.device_nets[1:], self.pools[1:]):
                dn.evaluate_pending(pool)           # finish the round the graph left half done
            self._phase_offset = False
        current = torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(current)
        for _ in range(int(rounds)):
            for dn, pool, st in zip(self.device_nets, self.pools, self.streams):
                with torch.cuda.stream(st):
                    dn.search_round(pool, selfplay=True)
        for st in self.streams:
            current.wait_stream(st)

    def _capture_round(self) -> None:
        """Capture one fused round (tree kernel: expand/backup of the previous batch and the next selection into
        dense leaf slots, then the network kernels) into a CUDA graph; the round has no host-visible state, so
        replaying it is identical to launching it."""
        dn, pool = self.device_nets[0], self.pools[0]
        dn.search_round(pool, selfplay=True, fused=True)        # warm every kernel outside the capture
        torch.cuda.synchronize()
        lib = _lib.load()
        before = lib.dm_launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            dn.search_round(pool, selfplay=True, fused=True)
        self._graph_launches = int(lib.dm_launch_count() - before)
        self._graph = graph

    def _capture_pipelined_round(self) -> None:
        """One graph with a branch per pool.  Pool 0's branch is select -> network -> expand/backup;
        the other pools run a phase apart, network -> expand/backup -> select (their selection for
        the next replay), so a pool's tree kernels are in flight while another pool's trunk
        occupies the tensor cores.  Every pool still does exactly one round per replay."""
        for dn, pool in zip(self.device_nets, self.pools):
            dn.search_round(pool, selfplay=True)    # warm every kernel outside the capture
        torch.cuda.synchronize()
        lib = _lib.load()
        before = lib.dm_launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            current = torch.cuda.current_stream()
            for st in self.streams:
                st.wait_stream(current)
            for i, (dn, pool, st) in enumerate(zip(self.device_nets, self.pools, self.streams)):
                with torch.cuda.stream(st):
                    if i == 0:
                        dn.search_round(pool, selfplay=True)
                    else:
                        dn.evaluate_pending(pool)
                        pool.selfplay_select()
            for st in self.streams:
                current.wait_stream(st)
        self._graph_launches = int(lib.dm_launch_count() - before)
        self._graph = graph

    def graph_launches(self) -> int:
        """Kernel launches executed through graph replays so far."""
        return self.graph_replays * self._graph_launches

    def counters(self) -> dict:
        total = None
        for pool in self.pools:
            c = pool.counters()
            if total is None:
                total = dict(c)
            else:
                for k, v in c.items():
                    total[k] = max(total[k], v) if k == "nodes_high_water" else total[k] + v
        return total

    def positions_written(self) -> int:
        return sum(int(pool.replay_written.item()) for pool in self.pools)

    def sample_batch(self, batch_size: int, generator: Optional[torch.Generator] = None,
                     written: Optional[Sequence[int]] = None, skip_oldest: int = 0
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """Uniform sample over the replay rings: (first, second, player, pi[B, A], z[B, 1]) on the device
        (reference train.py:337-348 samples uniformly over positions).

        ``written`` (positions written per pool, e.g. from ``positions_written_per_pool()`` at a moment
        when self-play was idle) restricts the sample to positions that were complete at that moment;
        ``skip_oldest`` additionally leaves out that many of each ring's oldest positions -- the ones
        self-play may be overwriting while the sample is taken on another stream."""
        if written is None:
            written = self.positions_written_per_pool()
        dev = self.pool.device
        A = self.pool.num_actions
        parts = []
        for pool, w in zip(self.pools, written):
            cap = pool.replay_capacity
            n = min(int(w), cap)
            if n <= 0:
                continue
            if w > cap and skip_oldest > 0:
                # ring is full: slots [head, head + skip) are the next to be overwritten
                head = int(w) % cap
                keep = torch.arange(n, device=dev)
                keep = keep[((keep - head) % cap) >= min(skip_oldest, cap - 1)]
            else:
                keep = torch.arange(n, device=dev)
            parts.append((pool, keep))
        total = sum(int(k.numel()) for _, k in parts)
        if total == 0:
            raise RuntimeError("replay ring is empty")
        idx = torch.randint(0, total, (batch_size,), device=dev, generator=generator)
        if len(parts) == 1:
            pool, keep = parts[0]
            sel = keep[idx]
            return (pool.replay_first[sel], pool.replay_second[sel], pool.replay_player[sel], pool.replay_pi[sel, :A],
                    pool.replay_z[sel].unsqueeze(1))
        first = torch.cat([p.replay_first[k] for p, k in parts])
        second = torch.cat([p.replay_second[k] for p, k in parts])
        player = torch.cat([p.replay_player[k] for p, k in parts])
        pi = torch.cat([p.replay_pi[k, :A] for p, k in parts])
        z = torch.cat([p.replay_z[k] for p, k in parts])
        return first[idx], second[idx], player[idx], pi[idx], z[idx].unsqueeze(1)

    def positions_written_per_pool(self):
        return [int(pool.replay_written.item()) for pool in self.pools]
