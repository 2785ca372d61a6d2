"""ctypes binding of libdeepmcts_b200.so (C ABI declared in include/deepmcts_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module
raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C deep_mcts_b200/csrc``.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_uint64, c_void_p
from pathlib import Path

DM_MAX_ACTIONS = 65
DM_MAX_DEPTH = 128
DM_MAX_LEAVES = 8
DM_REPLAY_POS_BITS = 38
DM_REPLAY_GAME_RING = 1 << 18

DM_TICTACTOE, DM_HEX, DM_HEX_SWAP, DM_OTHELLO = 0, 1, 2, 3
DM_EVAL_NONE, DM_EVAL_UNIFORM, DM_EVAL_HASHED, DM_EVAL_EXTERNAL = 0, 1, 2, 3
DM_ROLLOUT_NONE, DM_ROLLOUT_RANDOM, DM_ROLLOUT_MIN_ACTION, DM_ROLLOUT_MAX_ACTION = 0, 1, 2, 3

DM_ERR_INVALID, DM_ERR_CUDA, DM_ERR_CAPACITY, DM_ERR_STATE = -1, -2, -3, -4

# DM_LIB_PATH selects another build of the same library (A/B experiments); there is still no fallback.
LIB_PATH = Path(os.environ.get("DM_LIB_PATH") or (Path(__file__).resolve().parent / "libdeepmcts_b200.so"))


class PoolConfig(ctypes.Structure):
    _fields_ = [
        ("game", c_int32), ("grid", c_int32), ("n_trees", c_int32), ("max_nodes_per_tree", c_int32),
        ("device", c_int32), ("compaction", c_int32), ("c_puct", c_double), ("seed", c_uint64),
    ]


class SearchConfig(ctypes.Structure):
    _fields_ = [
        ("num_simulations", c_int32), ("eval_kind", c_int32), ("rollout_kind", c_int32),
        ("sample_move_cutoff", c_int32), ("dirichlet_alpha", c_double), ("dirichlet_factor", c_double),
        ("rollout_share", c_double), ("eval_salt", c_uint64), ("leaves_per_round", c_int32), ("reserved1", c_int32),
    ]


class NetConfig(ctypes.Structure):
    _fields_ = [
        ("game", c_int32), ("grid", c_int32), ("channels", c_int32), ("num_residual", c_int32),
        ("hidden", c_int32), ("device", c_int32), ("max_batch", c_int32), ("reserved0", c_int32),
    ]


class DeepMctsError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libdeepmcts_b200 error {code}: {message}")
        self.code = code


_P = c_void_p  # device pointers and streams travel as plain integers

_SIGNATURES = {
    "dm_abi_version": (c_int, []),
    "dm_last_error": (c_char_p, []),
    "dm_launch_count": (c_uint64, []),
    "dm_game_num_actions": (c_int, [c_int, c_int]),
    "dm_game_initial": (c_int, [c_int, c_int, c_int, _P, _P, _P, _P]),
    "dm_game_legal": (c_int, [c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "dm_game_child": (c_int, [c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dm_game_status": (c_int, [c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "dm_game_rollout": (c_int, [c_int, c_int, c_int, _P, _P, _P, c_int, c_uint64, _P, _P, _P]),
    "dm_eval_uniform": (c_int, [c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "dm_eval_hashed": (c_int, [c_int, c_int, c_int, _P, _P, _P, _P, _P, c_uint64, _P, _P, _P]),
    "dm_pool_create": (c_int, [POINTER(PoolConfig), POINTER(c_void_p)]),
    "dm_pool_destroy": (c_int, [_P]),
    "dm_pool_configure": (c_int, [_P, POINTER(SearchConfig)]),
    "dm_pool_reset": (c_int, [_P, _P, c_int, _P]),
    "dm_pool_set_root_state": (c_int, [_P, _P, c_int, _P, _P, _P, _P]),
    "dm_pool_observe": (c_int, [_P, _P, c_int, _P, _P, _P, _P]),
    "dm_pool_begin_step": (c_int, [_P, _P]),
    "dm_pool_begin_step_ids": (c_int, [_P, _P, c_int, _P]),
    "dm_pool_set_noise": (c_int, [_P, _P]),
    "dm_pool_search_local": (c_int, [_P, _P]),
    "dm_pool_select": (c_int, [_P, _P]),
    "dm_pool_leaf_buffers": (c_int, [_P, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                     POINTER(c_void_p), POINTER(c_void_p)]),
    "dm_pool_leaf_index": (c_int, [_P, POINTER(c_void_p)]),
    "dm_pool_debug_timing": (c_int, [_P, POINTER(c_void_p)]),
    "dm_pool_cache_enable": (c_int, [_P, c_int, c_int]),
    "dm_pool_cache_clear": (c_int, [_P, _P]),
    "dm_pool_expand_backup": (c_int, [_P, _P, c_int, _P, c_int, _P]),
    "dm_pool_round": (c_int, [_P, _P, c_int, _P, c_int, c_int, _P]),
    "dm_pool_root_visits": (c_int, [_P, _P, _P, _P]),
    "dm_pool_root_children": (c_int, [_P, _P, _P, _P, _P]),
    "dm_pool_root_child_stats": (c_int, [_P, _P, _P, _P]),
    "dm_pool_advance": (c_int, [_P, _P, _P]),
    "dm_pool_states": (c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "dm_pool_counters": (c_int, [_P, POINTER(c_uint64)]),
    "dm_selfplay_enable": (c_int, [_P, c_int]),
    "dm_selfplay_select": (c_int, [_P, _P]),
    "dm_selfplay_replay": (c_int, [_P, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                   POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p)]),
    "dm_selfplay_games": (c_int, [_P, POINTER(c_void_p)]),
    "dm_net_create": (c_int, [POINTER(NetConfig), POINTER(c_void_p)]),
    "dm_net_destroy": (c_int, [_P]),
    "dm_net_set_tensor": (c_int, [_P, c_char_p, _P, c_int64]),
    "dm_net_finalize": (c_int, [_P, _P]),
    "dm_net_evaluate": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "dm_net_evaluate_indexed": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dm_net_forward": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "dm_net_rollout_greedy": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P]),
    "dm_net_profile": (c_int, [_P, c_int]),
    "dm_net_profile_read": (c_int, [_P, POINTER(c_double), POINTER(c_int64)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.dm_abi_version() != 1:
        raise ImportError(f"ABI mismatch: library reports version {lib.dm_abi_version()}")
    _lib = lib
    return lib


def check(code):
    """Turn a negative dm_status into an exception carrying dm_last_error()."""
    if code is not None and code < 0:
        message = load().dm_last_error().decode("utf-8", "replace")
        if code == DM_ERR_INVALID:
            raise ValueError(message)
        raise DeepMctsError(code, message)
    return code


def call(name, *args):
    return check(getattr(load(), name)(*args))
