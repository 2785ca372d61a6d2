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
                 pool_sizes: Optional[Sequence[int]] = None, eval_cache_log2: Optional[int] = 22) -> None:
        """``net`` is a GameNet (its weights are used) or a DeviceNet.

        ``eval_cache_log2``: size (log2 entries) of each pool's device evaluation cache -- the reference's
        self-play workers evaluate through ``cached_state_evaluator`` (train.py:258-266, 408-414); ``None``
        evaluates every leaf.  Search results are identical either way.

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
            if eval_cache_log2 is not None:
                pool.enable_cache(eval_cache_log2, torch.float32)
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
        self._warmed = False        # every kernel of the round has run once (needed only before the first capture)
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
        """Re-pack the GameNet's current weights into every pool's kernels (weight transfer, reference
        train.py:169-171).  The captured round graph is dropped -- the trunk kernels take the folded
        biases and the value head's 1x1 weights as by-value arguments, which a capture freezes -- and the
        evaluation caches are cleared (they memoise the old network)."""
        if self._phase_offset:
            for dn, pool in zip(self.device_nets[1:], self.pools[1:]):
                dn.evaluate_pending(pool)           # finish the half-done round with the weights it was started with
            self._phase_offset = False
        if not isinstance(self.net, DeviceNet):
            sd = self.net.net.state_dict()
            for dn in self.device_nets:
                dn.load_state_dict(sd)
        self.invalidate_weights()

    def invalidate_weights(self) -> None:
        """Call after the device network's weights changed behind this engine's back (a shared DeviceNet)."""
        self._graph = None
        for pool in self.pools:
            pool.cache_clear()

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

    @staticmethod
    def _capture(fn) -> "torch.cuda.CUDAGraph":
        """Capture ``fn`` on a side stream with the low-level API: no device synchronisation, no gc / cache flush
        (``torch.cuda.graph`` does all three), so re-capturing after a weight transfer costs a fraction of a
        millisecond and does not stall the training stream."""
        graph = torch.cuda.CUDAGraph()
        current = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=current.device)
        side.wait_stream(current)
        with torch.cuda.stream(side):
            graph.capture_begin(capture_error_mode="thread_local")
            try:
                fn()
            finally:
                graph.capture_end()
        current.wait_stream(side)
        return graph

    def _capture_round(self) -> None:
        """Capture one fused round (tree kernel: expand/backup of the previous batch and the next selection into
        dense leaf slots, then the network kernels) into a CUDA graph; the round has no host-visible state, so
        replaying it is identical to launching it."""
        dn, pool = self.device_nets[0], self.pools[0]
        if not self._warmed:
            dn.search_round(pool, selfplay=True, fused=True)    # first capture: load every kernel outside the capture
            torch.cuda.synchronize()
            self._warmed = True
        lib = _lib.load()
        before = lib.dm_launch_count()
        self._graph = self._capture(lambda: dn.search_round(pool, selfplay=True, fused=True))
        self._graph_launches = int(lib.dm_launch_count() - before)

    def _capture_pipelined_round(self) -> None:
        """One graph with a branch per pool.  Pool 0's branch is select -> network -> expand/backup;
        the other pools run a phase apart, network -> expand/backup -> select (their selection for
        the next replay), so a pool's tree kernels are in flight while another pool's trunk
        occupies the tensor cores.  Every pool still does exactly one round per replay."""
        if not self._warmed:
            for dn, pool in zip(self.device_nets, self.pools):
                dn.search_round(pool, selfplay=True)    # first capture: load every kernel outside the capture
            torch.cuda.synchronize()
            self._warmed = True
        lib = _lib.load()
        before = lib.dm_launch_count()

        def body():
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

        self._graph = self._capture(body)
        self._graph_launches = int(lib.dm_launch_count() - before)

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
        return sum(pool.positions_written() for pool in self.pools)

    def replay_snapshot(self, max_games: Optional[int] = None, protect: int = 0):
        """[(pool, lo, hi)]: per pool the range of (monotonic) replay positions a trainer may read -- complete now,
        inside the window of the last ``max_games`` games, and not among the ``protect`` ring slots that come next in
        line to be overwritten.  ``protect`` = the most positions self-play can append while the snapshot is in use
        (n_trees x DM_MAX_DEPTH per block covers every case: a tree finishes at most one game per block) makes reads
        through the snapshot race-free while ``run_rounds`` is in flight on another stream.  Synchronises once."""
        ranges = []
        for pool, size in zip(self.pools, self.pool_sizes):
            hi, _ = pool.replay_progress()
            lo = max(0, hi - pool.replay_capacity + min(int(protect), pool.replay_capacity))
            if max_games is not None:
                lo = max(lo, pool.replay_window_start(max(1, int(max_games) * size // self.n_trees)))
            if hi > lo:
                ranges.append((pool, lo, hi))
        return ranges

    def sample_batch(self, batch_size: int, generator: Optional[torch.Generator] = None,
                     max_games: Optional[int] = None, snapshot=None
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """Uniform sample over the replay rings: (first, second, player, pi[B, A], z[B, 1]) on the device.
        The reference picks a game with probability proportional to its length, then one of its positions
        (train.py:337-348): uniform over positions.  ``max_games`` restricts the sample to the positions of
        the last ``max_games`` finished games (the reference trims its buffer to ``replay_buffer_max_size``
        GAMES, train.py:149-151); with several pools every pool contributes its share of that window.

        The sample reads the rings on the current stream: take it while self-play is idle (or ordered before it on
        the same stream), or pass a ``snapshot`` from ``replay_snapshot(max_games, protect)`` taken before self-play
        was launched -- then no host synchronisation happens here and the reads stay clear of the slots being
        written."""
        dev = self.pool.device
        A = self.pool.num_actions
        ranges = snapshot if snapshot is not None else self.replay_snapshot(max_games)
        total = sum(hi - lo for _, lo, hi in ranges)
        if total == 0:
            raise RuntimeError("replay ring is empty")
        idx = torch.randint(0, total, (batch_size,), device=dev, generator=generator)
        if len(ranges) == 1:
            pool, lo, _ = ranges[0]
            sel = (idx + lo) % pool.replay_capacity
            return (pool.replay_first[sel], pool.replay_second[sel], pool.replay_player[sel], pool.replay_pi[sel, :A],
                    pool.replay_z[sel].unsqueeze(1))
        first = torch.empty(batch_size, dtype=torch.int64, device=dev)
        second = torch.empty(batch_size, dtype=torch.int64, device=dev)
        player = torch.empty(batch_size, dtype=torch.uint8, device=dev)
        pi = torch.empty((batch_size, A), dtype=torch.float32, device=dev)
        z = torch.empty(batch_size, dtype=torch.float32, device=dev)
        offset = 0
        for pool, lo, hi in ranges:
            mine = (idx >= offset) & (idx < offset + (hi - lo))
            sel = (idx[mine] - offset + lo) % pool.replay_capacity
            first[mine], second[mine], player[mine] = pool.replay_first[sel], pool.replay_second[sel], pool.replay_player[sel]
            pi[mine], z[mine] = pool.replay_pi[sel, :A], pool.replay_z[sel]
            offset += hi - lo
        return first, second, player, pi, z.unsqueeze(1)

    def games_written(self) -> int:
        return sum(pool.replay_progress()[1] for pool in self.pools)

    def positions_written_per_pool(self):
        return [pool.positions_written() for pool in self.pools]
