"""Tic-tac-toe policy/value net: fixed 3x3 board, [3, 3] policy, 3 channels unless told otherwise
(interface of reference deep_mcts/tictactoe/convolutionalnet.py:13-107; unlike the other nets its
constructor has no grid-size argument)."""
from typing import Any, Optional

from ..boardnet import BoardGameNet
from ..game import GameManager
from .game import TicTacToeManager


class ConvolutionalTicTacToeNet(BoardGameNet):
    manager_cls = TicTacToeManager
    policy_as_grid = True
    fixed_grid = 3
    default_channels = 3

    def __init__(self, manager: Optional[GameManager] = None, *args: Any, **kwargs: Any) -> None:
        super().__init__(None, manager, *args, **kwargs)
