"""Hex policy/value net: [g, g] policy; SECOND-to-move positions are evaluated on the transposed board
(interface of reference deep_mcts/hex/convolutionalnet.py:13-134)."""
from ..boardnet import BoardGameNet
from .game import HexManager


class ConvolutionalHexNet(BoardGameNet):
    manager_cls = HexManager
    policy_as_grid = True
    transpose_second = True
