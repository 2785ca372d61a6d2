"""Othello GTP engine (interface of reference deep_mcts/othello/gtp_interface.py)."""
import sys

from ..game import Action
from ..gtp_interface import GTPInterface, format_cell, net_from_argv, parse_cell
from .convolutionalnet import ConvolutionalOthelloNet


class OthelloGTPInterface(GTPInterface):
    def __init__(self, board_size: int = 6) -> None:
        super().__init__(board_size)

    @staticmethod
    def parse_move(move: str, board_size: int) -> Action:
        if move.lower() == "pass":
            return board_size ** 2
        return parse_cell(move, board_size)

    @staticmethod
    def format_move(move: Action, board_size: int) -> str:
        return "pass" if move == board_size ** 2 else format_cell(move, board_size)

    @staticmethod
    def get_game_net(board_size: int) -> ConvolutionalOthelloNet:
        return net_from_argv(ConvolutionalOthelloNet, board_size)


if __name__ == "__main__":
    OthelloGTPInterface(board_size=int(sys.argv[1]) if len(sys.argv) == 3 else 6).start()
