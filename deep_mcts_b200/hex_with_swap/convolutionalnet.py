"""Hex-with-swap policy/value net: flat policy over the cells plus "swap", transposed board for SECOND
(interface of reference deep_mcts/hex_with_swap/convolutionalnet.py:13-137)."""
from ..boardnet import BoardGameNet
from .game import HexWithSwapManager


class ConvolutionalHexWithSwapNet(BoardGameNet):
    manager_cls = HexWithSwapManager
    extra_actions = 1
    transpose_second = True
