"""Host-side adaptor between reference-style state objects and the CUDA game engines.

A ``BoardManager`` keeps the reference's GameManager method names and return types
(reference deep_mcts/game.py:76-113) but every rule evaluation -- legal moves and their
order, child states, terminal test, outcome -- runs in the kernels of csrc/games.cu through
the C ABI (dm_game_*).  Single-state calls are batches of one; the ``*_batch`` methods take
numpy arrays of bitboards and are what the search engine and the tests use.
"""
from dataclasses import dataclass
from functools import lru_cache
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .game import Action, CellState, GameManager, Outcome, Player, State

_CELL = {-1: CellState.EMPTY, 0: CellState.SECOND_PLAYER, 1: CellState.FIRST_PLAYER}
_OUTCOME = {0.0: Outcome.FIRST_PLAYER_WIN, 1.0: Outcome.SECOND_PLAYER_WIN, 0.5: Outcome.DRAW}


@dataclass(unsafe_hash=True)
class BoardState(State):
    """player + grid of CellState rows, as in the reference's Othello/Hex/TicTacToe states."""

    __slots__ = ["grid"]
    grid: Tuple[Tuple[CellState, ...], ...]


def state_to_bits(state) -> Tuple[int, int, int]:
    """(player, first_bitboard, second_bitboard); bit y*g+x."""
    g = len(state.grid)
    first = second = 0
    for y, row in enumerate(state.grid):
        for x, cell in enumerate(row):
            c = int(cell)
            if c == 1:
                first |= 1 << (y * g + x)
            elif c == 0:
                second |= 1 << (y * g + x)
    return int(state.player), first, second


def bits_to_grid(first: int, second: int, g: int):
    return tuple(
        tuple(
            _CELL[1 if (first >> (y * g + x)) & 1 else (0 if (second >> (y * g + x)) & 1 else -1)]
            for x in range(g)
        )
        for y in range(g)
    )


def _dev(array: np.ndarray, device) -> torch.Tensor:
    if array.dtype == np.uint64:
        array = array.view(np.int64)
    return torch.from_numpy(np.ascontiguousarray(array)).to(device)


def current_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class BoardManager(GameManager):
    """GameManager whose rules run on the GPU.  ``device`` defaults to the current CUDA device."""

    game_id: int
    state_cls = BoardState

    def __init__(self, game_id: int, board: int, grid_size: int, num_actions: int, device=None) -> None:
        super().__init__(grid_size, num_actions)
        self.game_id = game_id
        self.board = board  # true board side (tic-tac-toe passes grid_size=9 like the reference)
        self._device = device

    # ------------------------------------------------------------------ device plumbing
    @property
    def device(self) -> torch.device:
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("deep_mcts_b200 needs a CUDA device: the game engines are CUDA kernels")
            self._device = torch.device("cuda", torch.cuda.current_device())
        return torch.device(self._device)

    def _upload(self, player, first, second):
        dev = self.device
        return (_dev(np.asarray(first, dtype=np.uint64), dev), _dev(np.asarray(second, dtype=np.uint64), dev),
                _dev(np.asarray(player, dtype=np.uint8), dev))

    # ----------------------------------------------------------------------- batch API
    def initial_batch(self, n: int):
        dev = self.device
        first = torch.empty(n, dtype=torch.int64, device=dev)
        second = torch.empty(n, dtype=torch.int64, device=dev)
        player = torch.empty(n, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.call("dm_game_initial", self.game_id, self.board, n, first.data_ptr(), second.data_ptr(),
                      player.data_ptr(), current_stream())
        return (player.cpu().numpy(), first.cpu().numpy().view(np.uint64), second.cpu().numpy().view(np.uint64))

    def legal_batch(self, player, first, second):
        """(list_order[n, 65], children_order[n, 65], count[n]); unused entries are 255."""
        f, s, p = self._upload(player, first, second)
        n = f.numel()
        dev = self.device
        lst = torch.empty((n, _lib.DM_MAX_ACTIONS), dtype=torch.uint8, device=dev)
        order = torch.empty((n, _lib.DM_MAX_ACTIONS), dtype=torch.uint8, device=dev)
        count = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("dm_game_legal", self.game_id, self.board, n, f.data_ptr(), s.data_ptr(), p.data_ptr(),
                      lst.data_ptr(), order.data_ptr(), count.data_ptr(), current_stream())
        return lst.cpu().numpy(), order.cpu().numpy(), count.cpu().numpy()

    def child_batch(self, player, first, second, actions):
        f, s, p = self._upload(player, first, second)
        n = f.numel()
        dev = self.device
        a = _dev(np.asarray(actions, dtype=np.int32), dev)
        of, os_, op = torch.empty_like(f), torch.empty_like(s), torch.empty_like(p)
        ok = torch.empty(n, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.call("dm_game_child", self.game_id, self.board, n, f.data_ptr(), s.data_ptr(), p.data_ptr(),
                      a.data_ptr(), of.data_ptr(), os_.data_ptr(), op.data_ptr(), ok.data_ptr(), current_stream())
        return (op.cpu().numpy(), of.cpu().numpy().view(np.uint64), os_.cpu().numpy().view(np.uint64),
                ok.cpu().numpy().astype(bool))

    def status_batch(self, player, first, second):
        f, s, p = self._upload(player, first, second)
        n = f.numel()
        dev = self.device
        fin = torch.empty(n, dtype=torch.uint8, device=dev)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("dm_game_status", self.game_id, self.board, n, f.data_ptr(), s.data_ptr(), p.data_ptr(),
                      fin.data_ptr(), out.data_ptr(), current_stream())
        return fin.cpu().numpy().astype(bool), out.cpu().numpy()

    def rollout_batch(self, player, first, second, kind=_lib.DM_ROLLOUT_RANDOM, seed=0):
        """Outcome values and ply counts of one playout per state (MCTS.rollout)."""
        f, s, p = self._upload(player, first, second)
        n = f.numel()
        dev = self.device
        out = torch.empty(n, dtype=torch.float32, device=dev)
        plies = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("dm_game_rollout", self.game_id, self.board, n, f.data_ptr(), s.data_ptr(), p.data_ptr(),
                      int(kind), int(seed), out.data_ptr(), plies.data_ptr(), current_stream())
        return out.cpu().numpy(), plies.cpu().numpy()

    # ------------------------------------------------------------- single-state helpers
    def _make_state(self, player: int, first: int, second: int):
        return self.state_cls(Player(int(player)), bits_to_grid(int(first), int(second), self.board))

    def bits(self, state) -> Tuple[int, int, int]:
        return state_to_bits(state)

    def initial_game_state(self):
        p, f, s = self.initial_batch(1)
        return self._make_state(p[0], f[0], s[0])

    @lru_cache(maxsize=2 ** 16)
    def _legal(self, state) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
        p, f, s = state_to_bits(state)
        lst, order, count = self.legal_batch([p], [f], [s])
        n = int(count[0])
        return tuple(int(a) for a in lst[0, :n]), tuple(int(a) for a in order[0, :n])

    def legal_actions(self, state) -> List[Action]:
        return list(self._legal(state)[0])

    def children_order(self, state) -> List[Action]:
        """Order in which the search creates this state's children (= ``list(set(legal_actions))``
        in the reference)."""
        return list(self._legal(state)[1])

    @lru_cache(maxsize=2 ** 16)
    def generate_child_state(self, state, action: Action):
        p, f, s = state_to_bits(state)
        cp, cf, cs, ok = self.child_batch([p], [f], [s], [int(action)])
        if not ok[0]:
            # the reference asserts `action in self.legal_actions(state)`
            raise AssertionError(f"illegal action {action}")
        return self._make_state(cp[0], cf[0], cs[0])

    @lru_cache(maxsize=2 ** 16)
    def _status(self, state) -> Tuple[bool, float]:
        p, f, s = state_to_bits(state)
        fin, out = self.status_batch([p], [f], [s])
        return bool(fin[0]), float(out[0])

    def is_final_state(self, state) -> bool:
        return self._status(state)[0]

    def evaluate_final_state(self, state) -> Outcome:
        return _OUTCOME[self._status(state)[1]]

    def _board_rows(self, action_probabilities: Sequence[float]) -> List[List[float]]:
        g = self.board
        return [[float(action_probabilities[y * g + x]) for x in range(g)] for y in range(g)]
