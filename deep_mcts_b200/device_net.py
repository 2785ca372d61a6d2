"""DeviceNet: the CUDA inference engine for ConvolutionalNet (wraps dm_net_* of the C ABI).

Packs a reference-format state_dict (BatchNorm folded, weights re-laid for the kernels) and
evaluates batches of bitboard states entirely on the device:
``states -> planes -> conv trunk -> heads -> tanh/sign/softmax/mask`` (reference
deep_mcts/gamenet.py:61-85 over convolutionalnet.py:53-68).  precision "fp32" is the
CUDA-core path (1e-5 parity), "bf16" the tcgen05 tensor-core path (1e-2 parity, channels=128).
"""
from typing import Dict, Optional, Tuple

import weakref

import numpy as np
import torch

from . import _lib
from ._engine import BoardManager, current_stream

PRECISIONS = {"fp32": 0, "bf16": 1}


class DeviceNet:
    on_device = True

    def __init__(self, manager: BoardManager, channels: int, num_residual: int, hidden: int, max_batch: int = 4096,
                 precision: Optional[str] = None) -> None:
        self.manager = manager
        self.device = manager.device
        self.channels, self.num_residual, self.hidden = int(channels), int(num_residual), int(hidden)
        self.max_batch = int(max_batch)
        if precision is None:
            precision = "bf16" if (self.channels == 128 and self.num_residual >= 1) else "fp32"
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        cfg = _lib.NetConfig(manager.game_id, manager.board, self.channels, self.num_residual, self.hidden,
                             self.device.index or 0, self.max_batch, 0)
        handle = _lib.c_void_p()
        _lib.call("dm_net_create", cfg, handle)
        self._h = handle
        self._loaded = False
        self.weights_version = 0     # bumped by load_state_dict: evaluation caches of older weights are stale
        self.graph_rounds = True     # run_search replays a captured graph of the fused round for large pools
        self.graph_min_trees = 256
        self._graphs = {}
        A = _lib.DM_MAX_ACTIONS
        # outputs of the search path: a fused round leaves its last batch pending here until the next round
        self._value = torch.zeros(self.max_batch, dtype=torch.float32, device=self.device)
        self._probs = torch.zeros((self.max_batch, A), dtype=torch.float32, device=self.device)
        # outputs of the host-facing calls (evaluate / forward), kept apart so they never clobber a pending batch
        self._host_value = self._host_probs = None

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.call("dm_net_destroy", self._h)
            self._h = _lib.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -------------------------------------------------------------------------- weights
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """Upload a reference-format state_dict and (re)pack it for the kernels."""
        for key, tensor in state_dict.items():
            if key.endswith("num_batches_tracked"):
                continue
            host = np.ascontiguousarray(tensor.detach().to("cpu", torch.float32).numpy())
            _lib.call("dm_net_set_tensor", self._h, key.encode(), host.ctypes.data, host.size)
        with torch.cuda.device(self.device):
            _lib.call("dm_net_finalize", self._h, current_stream())
        self._loaded = True
        self.weights_version += 1

    # ------------------------------------------------------------------------ inference
    def _check(self, n: int) -> None:
        if not self._loaded:
            raise RuntimeError("DeviceNet has no weights: call load_state_dict first")
        if n > self.max_batch:
            raise ValueError(f"batch {n} exceeds max_batch {self.max_batch}")

    def evaluate_device(self, first: torch.Tensor, second: torch.Tensor, player: torch.Tensor, n: int,
                        count_dev: Optional[torch.Tensor] = None, precision: Optional[str] = None
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
        """GameNet.evaluate for ``n`` device-resident states; returns device tensors
        (value[n] absolute in [0,1], probs[n, 65]) that stay valid until the next call."""
        self._check(n)
        prec = PRECISIONS[precision or self.precision]
        with torch.cuda.device(self.device):
            _lib.call("dm_net_evaluate", self._h, prec, int(n), count_dev.data_ptr() if count_dev is not None else 0,
                      first.data_ptr(), second.data_ptr(), player.data_ptr(), self._value.data_ptr(),
                      self._probs.data_ptr(), current_stream())
        return self._value, self._probs

    def evaluate_indexed(self, pool) -> Tuple[torch.Tensor, torch.Tensor]:
        """Evaluate a dense leaf batch through its compacted index (``pool.leaf_index`` / ``pool.leaf_count``,
        written by ``pool.round(selfplay=True)``): only the slots that need the network are run, every result
        lands in its own slot of the output buffers."""
        self._check(pool.n_trees)
        with torch.cuda.device(self.device):
            _lib.call("dm_net_evaluate_indexed", self._h, PRECISIONS[self.precision], int(pool.n_trees),
                      pool.leaf_count.data_ptr(), pool.leaf_index.data_ptr(), pool.leaf_first.data_ptr(),
                      pool.leaf_second.data_ptr(), pool.leaf_player.data_ptr(), self._value.data_ptr(),
                      self._probs.data_ptr(), current_stream())
        return self._value, self._probs

    def _host_outputs(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._host_value is None:
            self._host_value = torch.zeros_like(self._value)
            self._host_probs = torch.zeros_like(self._probs)
        return self._host_value, self._host_probs

    def _upload(self, player, first, second):
        f = torch.from_numpy(np.asarray(first, dtype=np.uint64).view(np.int64)).to(self.device)
        s = torch.from_numpy(np.asarray(second, dtype=np.uint64).view(np.int64)).to(self.device)
        p = torch.from_numpy(np.asarray(player, dtype=np.uint8)).to(self.device)
        return f, s, p

    def evaluate(self, player, first, second, precision: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
        """Host arrays in, host arrays out: (value[n], probs[n, num_actions])."""
        f, s, p = self._upload(player, first, second)
        n = f.numel()
        self._check(n)
        value, probs = self._host_outputs()
        with torch.cuda.device(self.device):
            _lib.call("dm_net_evaluate", self._h, PRECISIONS[precision or self.precision], n, 0, f.data_ptr(),
                      s.data_ptr(), p.data_ptr(), value.data_ptr(), probs.data_ptr(), current_stream())
        return value[:n].cpu().numpy(), probs[:n, : self.manager.num_actions].cpu().numpy()

    def forward(self, player, first, second, precision: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
        """GameNet.forward: raw value (pre-tanh) and logits (Hex un-transpose applied)."""
        f, s, p = self._upload(player, first, second)
        n = f.numel()
        self._check(n)
        prec = PRECISIONS[precision or self.precision]
        value, probs = self._host_outputs()
        with torch.cuda.device(self.device):
            _lib.call("dm_net_forward", self._h, prec, n, f.data_ptr(), s.data_ptr(), p.data_ptr(),
                      value.data_ptr(), probs.data_ptr(), current_stream())
        return value[:n].cpu().numpy(), probs[:n, : self.manager.num_actions].cpu().numpy()

    def rollout_greedy(self, player, first, second, precision: Optional[str] = None) -> np.ndarray:
        """Outcome of playing argmax of the masked policy from every state
        (reference deep_mcts/evaluate_complex_rollouts.py:57-59)."""
        f, s, p = self._upload(player, first, second)
        n = f.numel()
        self._check(n)
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("dm_net_rollout_greedy", self._h, PRECISIONS[precision or self.precision], n, f.data_ptr(),
                      s.data_ptr(), p.data_ptr(), out.data_ptr(), current_stream())
        return out.cpu().numpy()

    def rollout_greedy_device(self, first: torch.Tensor, second: torch.Tensor, player: torch.Tensor, n: int,
                              out: torch.Tensor, precision: Optional[str] = None) -> torch.Tensor:
        """Greedy policy playouts for ``n`` device-resident states into ``out`` (device, float32)."""
        self._check(n)
        with torch.cuda.device(self.device):
            _lib.call("dm_net_rollout_greedy", self._h, PRECISIONS[precision or self.precision], int(n),
                      first.data_ptr(), second.data_ptr(), player.data_ptr(), out.data_ptr(), current_stream())
        return out

    # ------------------------------------------------------------------------- profiling
    def profile(self, enable: bool) -> None:
        _lib.call("dm_net_profile", self._h, 1 if enable else 0)

    def profile_read(self) -> dict:
        """{class: (milliseconds, launches)} for 'encode', 'trunk_conv', 'policy_fc' (all heads on the fp32 path), 'heads_finish' since profile(True)."""
        ms = (_lib.c_double * 4)()
        launches = (_lib.c_int64 * 4)()
        _lib.call("dm_net_profile_read", self._h, ms, launches)
        return {name: (float(ms[i]), int(launches[i])) for i, name in enumerate(("encode", "trunk_conv", "policy_fc", "heads_finish"))}

    # --------------------------------------------------------------------------- search
    def search_round(self, pool, selfplay: bool = False, timers: Optional[list] = None, fused: bool = False) -> None:
        """One search round for every tree of the pool (no host sync).

        ``fused = False``: select -> evaluate -> expand/backup (three steps, the tree is complete at
        the end).  ``fused = True``: one tree kernel (expand/backup of the leaves evaluated in the
        previous round, then the next selection) -> evaluate; the last batch stays pending and is
        consumed by the next fused round (or by ``evaluate_pending`` / an unfused round, which
        simply selects again).
        ``timers`` (a list) receives CUDA-event pairs around the tree kernels:
        ("select", start, stop) and ("expand_backup", start, stop), or ("tree_round", start, stop)."""
        if fused:
            if timers is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                ev[0].record()
            pool.round(self._probs, self._value, selfplay)
            if timers is not None:
                ev[1].record()
                timers.append(("tree_round", ev[0], ev[1]))
            if selfplay:
                self.evaluate_indexed(pool)     # dense batch: slot t belongs to tree t, live slots are listed
            else:
                self.evaluate_device(pool.leaf_first, pool.leaf_second, pool.leaf_player, pool.leaf_capacity,
                                     pool.leaf_count)
            return
        if timers is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        if selfplay:
            pool.selfplay_select()
        else:
            pool.select()
        if timers is not None:
            ev[1].record()
        value, probs = self.evaluate_device(pool.leaf_first, pool.leaf_second, pool.leaf_player, pool.leaf_capacity,
                                            pool.leaf_count)
        if timers is not None:
            ev[2].record()
        pool.expand_backup(probs, value)
        if timers is not None:
            ev[3].record()
            timers.append(("select", ev[0], ev[1]))
            timers.append(("expand_backup", ev[2], ev[3]))

    def _round_graph(self, pool):
        """The fused search round of ``pool`` as a CUDA graph (memset + tree kernel + network kernels: a launch-bound
        loop of ~8 small launches per round), captured on first use and again after every weight load -- the
        trunk kernels take folded biases by value, which a capture freezes."""
        key = id(pool)
        hit = self._graphs.get(key)
        if hit is not None and hit[0] == self.weights_version and hit[2]() is pool:
            return hit[1]
        self.search_round(pool, fused=True)     # warm every kernel outside the capture (a real round: harmless,
        torch.cuda.synchronize(self.device)     # run_search's caller counts simulations, not rounds)
        graph = torch.cuda.CUDAGraph()
        current = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(current)
        with torch.cuda.stream(side):           # low-level capture: no gc / allocator flush per capture
            graph.capture_begin(capture_error_mode="thread_local")
            try:
                self.search_round(pool, fused=True)
            finally:
                graph.capture_end()
        current.wait_stream(side)
        self._graphs = {k: v for k, v in self._graphs.items() if v[2]() is not None}   # drop graphs of dead pools
        self._graphs[key] = (self.weights_version, graph, weakref.ref(pool))
        return graph

    def evaluate_pending(self, pool) -> None:
        """Network evaluation + expand/backup of the leaves a previous select left pending."""
        if pool.dense_batch:
            value, probs = self.evaluate_indexed(pool)
        else:
            value, probs = self.evaluate_device(pool.leaf_first, pool.leaf_second, pool.leaf_player,
                                                pool.leaf_capacity, pool.leaf_count)
        pool.expand_backup(probs, value)

    def run_search(self, pool, sims: int, round_graph=None) -> None:
        """All armed simulations of one MCTS.step.  Exact mode: sims + 1 selections (root expansion),
        no host synchronisation; ``round_graph`` is an optional captured CUDA graph of
        ``search_round(pool, fused=True)`` to replay instead of launching the round.
        Virtual-loss mode: rounds until no tree hands out a leaf."""
        if pool.leaves_per_round == 1:
            # select + evaluate, then `sims` fused rounds (expand/backup -> select -> evaluate), then the
            # last batch's expand/backup
            pool.select()
            self.evaluate_device(pool.leaf_first, pool.leaf_second, pool.leaf_player, pool.leaf_capacity, pool.leaf_count)
            if round_graph is None and self.graph_rounds and pool.n_trees >= self.graph_min_trees and int(sims) >= 8:
                round_graph = self._round_graph(pool)
            for _ in range(int(sims)):
                if round_graph is not None:
                    round_graph.replay()
                else:
                    self.search_round(pool, fused=True)
            pool.expand_backup(self._probs, self._value)
            return
        for _ in range(int(sims) + 2):
            self.search_round(pool)
            if int(pool.leaf_count.item()) == 0:
                break
