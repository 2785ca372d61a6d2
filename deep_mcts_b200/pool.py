"""TreePool: thousands of independent MCTS trees resident in HBM, searched by CUDA kernels.

This is the batched engine underneath the reference-compatible ``MCTS`` / ``MCTSAgent``
classes (deep_mcts_b200/mcts.py).  One pool holds ``n_trees`` structure-of-arrays trees for
one game on one GPU; every method is a thin wrapper over a ``dm_pool_*`` entry point of the
C ABI (include/deepmcts_b200.h), which replaces reference deep_mcts/mcts.py:57-300.
"""
from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._engine import BoardManager, current_stream

_TYPESTR = {torch.int64: "<i8", torch.uint8: "|u1", torch.int32: "<i4", torch.float32: "<f4",
            torch.float64: "<f8"}


class _DevicePtr:
    """Exposes a library-owned device buffer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], dtype: torch.dtype) -> None:
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": _TYPESTR[dtype], "data": (int(ptr), False), "version": 2,
        }


def wrap_device_buffer(ptr: int, shape, dtype: torch.dtype, device) -> torch.Tensor:
    return torch.as_tensor(_DevicePtr(ptr, shape, dtype), device=device)


def default_node_capacity(manager: BoardManager, num_simulations: int) -> int:
    """Arena size per tree for a whole game without compaction: every simulation expands at most
    one node (at most num_actions children), once per ply, plus the root expansions."""
    plies = manager.board * manager.board + 8
    per_move = (int(num_simulations) + 1) * manager.num_actions
    return int(min(max(4096, plies * per_move // 2), 1 << 22))


class TreePool:
    def __init__(self, manager: BoardManager, n_trees: int, max_nodes_per_tree: int = 1 << 16, seed: int = 0,
                 c_puct: float = 1.25, compaction: bool = True) -> None:
        """``compaction``: slide a tree's live subtree to the front of its arena at a re-root once the
        arena may not hold the next move's search (or is half full, whichever comes first), so long games / long
        searches never outgrow ``max_nodes_per_tree``
        as long as one move's search fits in half of it."""
        self.manager = manager
        self.n_trees = int(n_trees)
        self.device = manager.device
        self.num_actions = manager.num_actions
        cfg = _lib.PoolConfig(manager.game_id, manager.board, self.n_trees, int(max_nodes_per_tree),
                              self.device.index or 0, 1 if compaction else 0, float(c_puct), int(seed))
        handle = _lib.c_void_p()
        _lib.call("dm_pool_create", cfg, handle)
        self._h = handle
        self._search = None
        self._noise = None
        ptrs = [_lib.c_void_p() for _ in range(5)]
        _lib.call("dm_pool_leaf_buffers", self._h, *ptrs)
        T = self.n_trees
        cap = T * _lib.DM_MAX_LEAVES   # leaf-batch capacity (virtual-loss mode hands out several leaves per tree)
        self.leaves_per_round = 1
        self.leaf_first = wrap_device_buffer(ptrs[0].value, (cap,), torch.int64, self.device)
        self.leaf_second = wrap_device_buffer(ptrs[1].value, (cap,), torch.int64, self.device)
        self.leaf_player = wrap_device_buffer(ptrs[2].value, (cap,), torch.uint8, self.device)
        self.leaf_tree = wrap_device_buffer(ptrs[3].value, (cap,), torch.int32, self.device)
        self.leaf_count = wrap_device_buffer(ptrs[4].value, (1,), torch.int32, self.device)
        idx = _lib.c_void_p()
        _lib.call("dm_pool_leaf_index", self._h, idx)
        # dense batches (round(selfplay=True)): the slots whose leaf needs the evaluator, compacted; count in leaf_count
        self.leaf_index = wrap_device_buffer(idx.value, (T,), torch.int32, self.device)
        self.cache_entries = 0
        self.dense_batch = False    # the pending leaf batch is dense (slot = tree) and must be read through leaf_index
        # read-out staging: ONE device buffer per call and a pinned host mirror, so a read-out is one kernel, one
        # asynchronous copy and one stream synchronisation (visits | stats; first | second | outcome | player | final)
        A = _lib.DM_MAX_ACTIONS
        self._visit_words = T * A + T * 2
        self._visit_buf = torch.zeros(self._visit_words, dtype=torch.int32, device=self.device)
        self._visits = self._visit_buf[: T * A].view(T, A)
        self._stats = self._visit_buf[T * A:].view(T, 2)
        self._state_buf = torch.zeros(T * 22, dtype=torch.uint8, device=self.device)
        self._actions_dev = torch.zeros(T, dtype=torch.int32, device=self.device)
        self._h_visit_buf = self._h_state_buf = self._h_actions = None
        self._actions_copied = None

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.call("dm_pool_destroy", self._h)
            self._h = _lib.c_void_p()

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ configuration
    def configure(self, num_simulations: int, eval_kind: int = _lib.DM_EVAL_EXTERNAL,
                  rollout_kind: int = _lib.DM_ROLLOUT_NONE, sample_move_cutoff: int = 0,
                  dirichlet_alpha: float = 0.0, dirichlet_factor: float = 0.25, rollout_share: float = 1.0,
                  eval_salt: int = 0, leaves_per_round: int = 1) -> None:
        """``leaves_per_round`` = 1 is the exact mode (bit-identical visit counts); 2..8 hands out that
        many leaves per tree per round with virtual loss (throughput mode for small pools)."""
        sc = _lib.SearchConfig(int(num_simulations), int(eval_kind), int(rollout_kind), int(sample_move_cutoff),
                               float(dirichlet_alpha), float(dirichlet_factor), float(rollout_share),
                               int(eval_salt), int(leaves_per_round), 0)
        _lib.call("dm_pool_configure", self._h, sc)
        self._search = sc
        self.leaves_per_round = max(1, int(leaves_per_round))

    def enable_cache(self, log2_entries: int = 22, dtype: torch.dtype = torch.float32) -> None:
        """Device-side ``cached_state_evaluator`` (reference train.py:408-414): memoise the external evaluator's
        outputs per state in a 2**log2_entries-entry table and evaluate equal leaves of one round once.
        Result-transparent for a fixed evaluator; ``dtype`` is the evaluator's output type."""
        assert dtype in (torch.float32, torch.float64)
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_cache_enable", self._h, int(log2_entries), 1 if dtype == torch.float64 else 0)
        self.cache_entries = 1 << int(log2_entries)

    def cache_clear(self) -> None:
        """Forget every cached evaluation (after a weight transfer)."""
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_cache_clear", self._h, self._stream())

    @property
    def leaf_capacity(self) -> int:
        """Largest number of leaves one select can hand out with the current configuration."""
        return self.n_trees * self.leaves_per_round

    def _ids(self, tree_ids):
        if tree_ids is None:
            return None, self.n_trees, 0
        ids = torch.as_tensor(np.asarray(tree_ids, dtype=np.int32)).to(self.device)
        return ids, ids.numel(), ids.data_ptr()

    def _stream(self):
        return current_stream()

    # ----------------------------------------------------------------------- tree admin
    def reset(self, tree_ids=None) -> None:
        ids, n, ptr = self._ids(tree_ids)
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_reset", self._h, ptr, n, self._stream())

    def _state_args(self, player, first, second):
        f = torch.from_numpy(np.asarray(first, dtype=np.uint64).view(np.int64)).to(self.device)
        s = torch.from_numpy(np.asarray(second, dtype=np.uint64).view(np.int64)).to(self.device)
        p = torch.from_numpy(np.asarray(player, dtype=np.uint8)).to(self.device)
        return f, s, p

    def set_root_states(self, player, first, second, tree_ids=None) -> None:
        ids, n, ptr = self._ids(tree_ids)
        f, s, p = self._state_args(player, first, second)
        assert f.numel() == n
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_set_root_state", self._h, ptr, n, f.data_ptr(), s.data_ptr(), p.data_ptr(),
                      self._stream())

    def observe(self, player, first, second, tree_ids=None) -> None:
        ids, n, ptr = self._ids(tree_ids)
        f, s, p = self._state_args(player, first, second)
        assert f.numel() == n
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_observe", self._h, ptr, n, f.data_ptr(), s.data_ptr(), p.data_ptr(), self._stream())

    def set_noise(self, noise: Optional[np.ndarray]) -> None:
        """Host-supplied Dirichlet noise [n_trees, <=65] in child order (parity tests); None = device RNG."""
        if noise is None:
            self._noise = None
            _lib.call("dm_pool_set_noise", self._h, 0)
            return
        buf = np.zeros((self.n_trees, _lib.DM_MAX_ACTIONS), dtype=np.float64)
        noise = np.asarray(noise, dtype=np.float64)
        buf[:, : noise.shape[1]] = noise
        self._noise = torch.from_numpy(buf).to(self.device)
        _lib.call("dm_pool_set_noise", self._h, self._noise.data_ptr())

    # --------------------------------------------------------------------------- search
    def begin_step(self, tree_ids=None) -> None:
        """Arm one MCTS.step on every tree, or only on ``tree_ids`` (the others sit the step out)."""
        with torch.cuda.device(self.device):
            if tree_ids is None:
                _lib.call("dm_pool_begin_step", self._h, self._stream())
            else:
                ids, n, ptr = self._ids(tree_ids)
                _lib.call("dm_pool_begin_step_ids", self._h, ptr, n, self._stream())

    def search_local(self) -> None:
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_search_local", self._h, self._stream())

    def select(self) -> None:
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_select", self._h, self._stream())
        self.dense_batch = False

    def expand_backup(self, priors: torch.Tensor, values: torch.Tensor) -> None:
        assert priors.is_cuda and values.is_cuda and priors.dtype == values.dtype
        assert priors.is_contiguous() and values.is_contiguous()
        is_f64 = 1 if priors.dtype == torch.float64 else 0
        assert priors.dtype in (torch.float32, torch.float64)
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_expand_backup", self._h, priors.data_ptr(), int(priors.shape[1]),
                      values.data_ptr(), is_f64, self._stream())

    def round(self, priors: torch.Tensor, values: torch.Tensor, selfplay: bool = False) -> None:
        """Expand + back up every pending leaf from the evaluator output of the previous batch, then
        select the next leaf of every tree, in one launch (``dm_pool_round``)."""
        assert priors.is_cuda and values.is_cuda and priors.dtype == values.dtype
        assert priors.is_contiguous() and values.is_contiguous()
        assert priors.dtype in (torch.float32, torch.float64)
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_round", self._h, priors.data_ptr(), int(priors.shape[1]), values.data_ptr(),
                      1 if priors.dtype == torch.float64 else 0, 1 if selfplay else 0, self._stream())
        self.dense_batch = bool(selfplay)

    def step_external(self, evaluate: Callable[[np.ndarray, np.ndarray, np.ndarray], Tuple[np.ndarray, np.ndarray]],
                      max_rounds: Optional[int] = None) -> None:
        """Run the armed simulations with a HOST evaluator: ``evaluate(player, first, second)``
        returns (values[L] float64, priors[L, A] float64) for the L pending leaves.  This is the
        injection point for arbitrary Python ``state_evaluator`` callables and for the
        deterministic evaluators of the parity tests."""
        rounds = 0
        limit = max_rounds if max_rounds is not None else (self._search.num_simulations + 2)
        while rounds < limit:
            self.select()
            n = int(self.leaf_count.item())
            if n == 0:
                break
            p = self.leaf_player[:n].cpu().numpy()
            f = self.leaf_first[:n].cpu().numpy().view(np.uint64)
            s = self.leaf_second[:n].cpu().numpy().view(np.uint64)
            values, priors = evaluate(p, f, s)
            pri = np.zeros((self.leaf_capacity, _lib.DM_MAX_ACTIONS), dtype=np.float64)
            pri[:n, : self.num_actions] = np.asarray(priors, dtype=np.float64)[:, : self.num_actions]
            val = np.zeros(self.leaf_capacity, dtype=np.float64)
            val[:n] = np.asarray(values, dtype=np.float64)
            self.expand_backup(torch.from_numpy(pri).to(self.device), torch.from_numpy(val).to(self.device))
            rounds += 1

    # -------------------------------------------------------------------------- read-out
    def root_visits(self) -> Tuple[np.ndarray, np.ndarray]:
        """(visits[n_trees, num_actions], stats[n_trees, 2]) of the last step (mcts.py:139-148)."""
        T, A = self.n_trees, _lib.DM_MAX_ACTIONS
        if self._h_visit_buf is None:
            self._h_visit_buf = torch.empty(self._visit_words, dtype=torch.int32, pin_memory=True)
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_root_visits", self._h, self._visits.data_ptr(), self._stats.data_ptr(),
                      self._stream())
            self._h_visit_buf.copy_(self._visit_buf, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        host = self._h_visit_buf.numpy()
        # copies: the pinned mirror is overwritten by the next read-out
        return host[: T * A].reshape(T, A)[:, : self.num_actions].copy(), host[T * A:].reshape(T, 2).copy()

    def root_children(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        T = self.n_trees
        order = torch.empty((T, _lib.DM_MAX_ACTIONS), dtype=torch.uint8, device=self.device)
        count = torch.empty(T, dtype=torch.int32, device=self.device)
        rv = torch.empty(T, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_root_children", self._h, order.data_ptr(), count.data_ptr(), rv.data_ptr(),
                      self._stream())
        return order.cpu().numpy(), count.cpu().numpy(), rv.cpu().numpy()

    def root_child_stats(self) -> Tuple[np.ndarray, np.ndarray]:
        """(prior[n_trees, 65], value_sum[n_trees, 65]) of the root's children in child order: ``Node.probability``
        (after the Dirichlet mixing of mcts.py:233-241) and ``Node.value_sum``."""
        T = self.n_trees
        prior = torch.empty((T, _lib.DM_MAX_ACTIONS), dtype=torch.float64, device=self.device)
        vsum = torch.empty((T, _lib.DM_MAX_ACTIONS), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_root_child_stats", self._h, prior.data_ptr(), vsum.data_ptr(), self._stream())
        return prior.cpu().numpy(), vsum.cpu().numpy()

    def advance(self, actions) -> None:
        a = np.asarray(actions, dtype=np.int32).reshape(-1)
        assert a.size == self.n_trees
        if self._h_actions is None:
            self._h_actions = torch.empty(self.n_trees, dtype=torch.int32, pin_memory=True)
            self._actions_copied = torch.cuda.Event()
        else:
            self._actions_copied.synchronize()      # the previous upload has left the pinned staging buffer
        self._h_actions.numpy()[:] = a
        with torch.cuda.device(self.device):
            self._actions_dev.copy_(self._h_actions, non_blocking=True)
            self._actions_copied.record()
            _lib.call("dm_pool_advance", self._h, self._actions_dev.data_ptr(), self._stream())

    def states(self):
        """(player, first, second, is_final, outcome) numpy arrays of the current roots."""
        T = self.n_trees
        if self._h_state_buf is None:
            self._h_state_buf = torch.empty(T * 22, dtype=torch.uint8, pin_memory=True)
        base = self._state_buf.data_ptr()       # first | second (8 B each) | outcome (4 B) | player | final (1 B each)
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_states", self._h, base, base + 8 * T, base + 20 * T, base + 21 * T, base + 16 * T,
                      self._stream())
            self._h_state_buf.copy_(self._state_buf, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        host = self._h_state_buf.numpy()
        return (host[20 * T: 21 * T].copy(), host[: 8 * T].view(np.uint64).copy(),
                host[8 * T: 16 * T].view(np.uint64).copy(), host[21 * T: 22 * T].astype(bool),
                host[16 * T: 20 * T].view(np.float32).copy())

    COUNTER_NAMES = ("simulations", "leaf_evals", "terminal_sims", "nodes_high_water", "games", "depth_sum",
                     "children_scanned", "children_created", "capacity_errors", "moves", "rollout_plies",
                     "invalid_actions", "compactions", "cache_hits", "cache_aliases", "cache_inserts")

    def counters(self) -> dict:
        out = (_lib.c_uint64 * 16)()
        with torch.cuda.device(self.device):
            _lib.call("dm_pool_counters", self._h, out)
        return {name: int(out[i]) for i, name in enumerate(self.COUNTER_NAMES)}

    # ------------------------------------------------------------------------- self-play
    def enable_selfplay(self, replay_capacity: int) -> None:
        with torch.cuda.device(self.device):
            _lib.call("dm_selfplay_enable", self._h, int(replay_capacity))
        ptrs = [_lib.c_void_p() for _ in range(6)]
        _lib.call("dm_selfplay_replay", self._h, *ptrs)
        R = int(replay_capacity)
        self.replay_capacity = R
        self.replay_first = wrap_device_buffer(ptrs[0].value, (R,), torch.int64, self.device)
        self.replay_second = wrap_device_buffer(ptrs[1].value, (R,), torch.int64, self.device)
        self.replay_player = wrap_device_buffer(ptrs[2].value, (R,), torch.uint8, self.device)
        self.replay_pi = wrap_device_buffer(ptrs[3].value, (R, _lib.DM_MAX_ACTIONS), torch.float32, self.device)
        self.replay_z = wrap_device_buffer(ptrs[4].value, (R,), torch.float32, self.device)
        # counter word: positions written (low DM_REPLAY_POS_BITS bits) | games written (bits above)
        self.replay_counter = wrap_device_buffer(ptrs[5].value, (1,), torch.int64, self.device)
        ge = _lib.c_void_p()
        _lib.call("dm_selfplay_games", self._h, ge)
        self.replay_game_end = wrap_device_buffer(ge.value, (_lib.DM_REPLAY_GAME_RING,), torch.int64, self.device)

    def replay_progress(self) -> Tuple[int, int]:
        """(positions, games) appended to the replay ring so far (synchronises)."""
        word = int(self.replay_counter.item())
        return word & ((1 << _lib.DM_REPLAY_POS_BITS) - 1), word >> _lib.DM_REPLAY_POS_BITS

    def positions_written(self) -> int:
        return self.replay_progress()[0]

    def replay_window_start(self, max_games: int) -> int:
        """Position (monotonic index) where the last ``max_games`` finished games begin -- the reference
        trims its replay buffer to that many GAMES (train.py:149-151).  0 while fewer have been played."""
        positions, games = self.replay_progress()
        if max_games <= 0 or games <= max_games:
            return 0
        if max_games >= _lib.DM_REPLAY_GAME_RING:
            return 0
        return int(self.replay_game_end[(games - max_games - 1) % _lib.DM_REPLAY_GAME_RING].item())

    def selfplay_select(self) -> None:
        with torch.cuda.device(self.device):
            _lib.call("dm_selfplay_select", self._h, self._stream())
        self.dense_batch = False
