"""Othello on the CUDA bitboard engine (interface of reference deep_mcts/othello/game.py:13-158).

Move generation (8-direction shift fill), flips, pass handling, terminal test and outcome run
in csrc/games.cuh; action ``grid_size**2`` is "pass", legal only when the mover has no move.
"""
import string
from dataclasses import dataclass
from typing import Sequence

from .. import _lib
from .._engine import BoardManager, BoardState
from ..game import Action


@dataclass(unsafe_hash=True)
class OthelloState(BoardState):
    __slots__ = []

    def __str__(self) -> str:
        g = len(self.grid)
        letters = " ".join(string.ascii_uppercase[:g])
        width = 2 * g + 3
        glyph = {-1: ".", 0: "0", 1: "1"}
        lines = [letters.center(width)]
        for i, row in enumerate(self.grid, 1):
            lines.append(f"{i} {' '.join(glyph[int(c)] for c in row)} {i}".center(width))
        lines.append(letters.center(width))
        return "\n".join(lines)


class OthelloManager(BoardManager):
    state_cls = OthelloState
    pass_move: Action

    def __init__(self, grid_size: int, device=None) -> None:
        super().__init__(_lib.DM_OTHELLO, grid_size, grid_size, grid_size ** 2 + 1, device)
        self.pass_move = grid_size ** 2

    def probabilities_grid(self, action_probabilities: Sequence[float]) -> str:
        rows = [" ".join(f"{p:.2f}" for p in row) for row in self._board_rows(action_probabilities)]
        rows.append(f"pass: {action_probabilities[self.pass_move]}")
        return "\n".join(rows)

    def action_str(self, action: Action) -> str:
        return "pass" if action == self.pass_move else super().action_str(action)
