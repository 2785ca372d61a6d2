"""Hex with the swap rule (interface of reference deep_mcts/hex_with_swap/game.py:9-44).

Action ``grid_size**2`` is "swap": legal only for SECOND while exactly one stone is on the
board; it recolours that stone in place (no transposition) and hands the move back to FIRST.
"""
from typing import Sequence

from .. import _lib
from ..game import Action
from ..hex.game import HexManager, HexState  # noqa: F401


class HexWithSwapManager(HexManager):
    swap_move: Action

    def __init__(self, grid_size: int, device=None) -> None:
        super().__init__(grid_size, num_actions=grid_size ** 2 + 1, device=device, _game_id=_lib.DM_HEX_SWAP)
        self.swap_move = grid_size ** 2

    def probabilities_grid(self, action_probabilities: Sequence[float]) -> str:
        return f"{super().probabilities_grid(action_probabilities[:-1])}\nswap: {action_probabilities[-1]}"

    def action_str(self, action: Action) -> str:
        return "swap" if action == self.swap_move else super().action_str(action)
