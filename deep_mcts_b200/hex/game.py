"""Hex on the CUDA bitboard engine (interface of reference deep_mcts/hex/game.py:13-157).

FIRST connects row 0 to row g-1, SECOND connects column 0 to column g-1 (SECOND is tested
first); win detection is a bitmask flood fill over the six hex neighbours in csrc/games.cuh.
"""
import string
from dataclasses import dataclass
from typing import Optional, Sequence

from .. import _lib
from .._engine import BoardManager, BoardState
from ..game import Action  # noqa: F401  (re-exported like the reference module)


@dataclass(unsafe_hash=True)
class HexState(BoardState):
    __slots__ = []

    def __str__(self) -> str:
        g = len(self.grid)
        letters = " ".join(string.ascii_uppercase[:g])
        width = 2 * g + 3
        glyph = {-1: ".", 0: "0", 1: "1"}
        lines = [letters.center(width - 1)]
        for i, row in enumerate(self.grid):
            indent = " " * (i if i < 9 else i - 1)
            lines.append(f"{indent}{i + 1} {' '.join(glyph[int(c)] for c in row)} {i + 1}".center(width))
        lines.append(f"{' ' * (g + 2)}{letters}".center(width))
        return "\n".join(lines)


class HexManager(BoardManager):
    state_cls = HexState

    def __init__(self, grid_size: int, num_actions: Optional[int] = None, device=None,
                 _game_id: int = _lib.DM_HEX) -> None:
        super().__init__(_game_id, grid_size, grid_size, grid_size ** 2 if num_actions is None else num_actions,
                         device)

    def probabilities_grid(self, action_probabilities: Sequence[float]) -> str:
        g = self.board
        width = 4 * g + 3
        lines = []
        for i, row in enumerate(self._board_rows(action_probabilities)):
            indent = " " * (i if i < 9 else i - 1)
            lines.append(f"{indent}{i + 1} {' '.join(f'{p:.2f}' for p in row)} {i + 1}".center(width))
        return "\n".join(lines)
