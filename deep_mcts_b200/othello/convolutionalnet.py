"""Othello policy/value net: flat policy over the g*g cells plus "pass"
(interface of reference deep_mcts/othello/convolutionalnet.py:13-113)."""
from ..boardnet import BoardGameNet
from .game import OthelloManager


class ConvolutionalOthelloNet(BoardGameNet):
    manager_cls = OthelloManager
    extra_actions = 1
