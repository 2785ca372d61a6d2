"""Hex GTP engine, usable from HexGui (interface of reference deep_mcts/hex/gtp_interface.py)."""
from ..gtp_interface import board_gtp_interface, run_engine
from .convolutionalnet import ConvolutionalHexNet

HexGTPInterface = board_gtp_interface("HexGTPInterface", ConvolutionalHexNet, default_size=5, hexgui=True)

if __name__ == "__main__":
    run_engine(HexGTPInterface)
