"""Tic-tac-toe on the CUDA engine (interface of reference deep_mcts/tictactoe/game.py:10-109).

Like the reference, the manager reports ``grid_size = 9`` (tictactoe/game.py:25-26); the real
board side is 3.  ``copy()`` therefore takes no argument here.
"""
from dataclasses import dataclass
from typing import Sequence

from .. import _lib
from .._engine import BoardManager, BoardState
from ..game import Action  # noqa: F401


@dataclass(unsafe_hash=True)
class TicTacToeState(BoardState):
    __slots__ = []

    def __str__(self) -> str:
        glyph = {-1: "#", 1: "X", 0: "O"}
        return "\n".join("".join(glyph[int(c)] for c in row) for row in self.grid)


class TicTacToeManager(BoardManager):
    state_cls = TicTacToeState

    def __init__(self, device=None) -> None:
        super().__init__(_lib.DM_TICTACTOE, 3, 9, 9, device)

    def copy(self) -> "TicTacToeManager":
        return TicTacToeManager(self._device)

    def probabilities_grid(self, action_probabilities: Sequence[float]) -> str:
        return "\n".join(" ".join(f"{p:.2f}" for p in row) for row in self._board_rows(action_probabilities))
