"""Hex-with-swap GTP engine; the extra move is "swap"
(interface of reference deep_mcts/hex_with_swap/gtp_interface.py)."""
from ..gtp_interface import board_gtp_interface, run_engine
from .convolutionalnet import ConvolutionalHexWithSwapNet

HexWithSwapGTPInterface = board_gtp_interface("HexWithSwapGTPInterface", ConvolutionalHexWithSwapNet, default_size=5,
                                              special_move="swap", hexgui=True)

if __name__ == "__main__":
    run_engine(HexWithSwapGTPInterface)
