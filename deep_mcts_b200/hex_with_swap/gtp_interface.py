"""Hex-with-swap GTP engine (interface of reference deep_mcts/hex_with_swap/gtp_interface.py)."""
import sys
from typing import List

from ..game import Action
from ..gtp_interface import GTPInterface, format_cell, net_from_argv, parse_cell
from .convolutionalnet import ConvolutionalHexWithSwapNet


class HexWithSwapGTPInterface(GTPInterface):
    def __init__(self, board_size: int = 5) -> None:
        super().__init__(board_size)
        self.commands["hexgui-analyze_commands"] = self.analyze_commands

    def analyze_commands(self, args: List[str]) -> None:
        return None

    @staticmethod
    def parse_move(move: str, board_size: int) -> Action:
        if move.lower() == "swap":
            return board_size ** 2
        return parse_cell(move, board_size)

    @staticmethod
    def format_move(move: Action, board_size: int) -> str:
        return "swap" if move == board_size ** 2 else format_cell(move, board_size)

    @staticmethod
    def get_game_net(board_size: int) -> ConvolutionalHexWithSwapNet:
        return net_from_argv(ConvolutionalHexWithSwapNet, board_size)


if __name__ == "__main__":
    HexWithSwapGTPInterface(board_size=int(sys.argv[1]) if len(sys.argv) == 3 else 5).start()
