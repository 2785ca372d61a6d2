"""Othello GTP engine; the extra move is "pass" (interface of reference deep_mcts/othello/gtp_interface.py)."""
from ..gtp_interface import board_gtp_interface, run_engine
from .convolutionalnet import ConvolutionalOthelloNet

OthelloGTPInterface = board_gtp_interface("OthelloGTPInterface", ConvolutionalOthelloNet, default_size=6,
                                          special_move="pass")

if __name__ == "__main__":
    run_engine(OthelloGTPInterface)
