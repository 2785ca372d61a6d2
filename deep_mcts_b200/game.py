"""Game vocabulary of the engine: outcomes, cells, players, states and the rules interface.

The names, members and numeric values are the reference's (deep_mcts/game.py:9-113), because
callers and checkpoints depend on them:

* tree values are absolute: ``Outcome`` is 0.0 for a win of the player who moved first, 1.0 for a
  win of the other, 0.5 for a draw -- nothing in the search ever negates a value;
* ``Player.FIRST`` (numeric 1) moves first and therefore *minimises*, ``Player.SECOND`` (0) maximises;
* a board cell holds -1 (empty), 0 (SECOND's stone) or 1 (FIRST's stone), so a stone's cell value is
  its owner's numeric value;
* an ``Action`` is a plain int: ``x + y * grid_size``, with ``grid_size ** 2`` for pass / swap.

On the device the same facts live in csrc/games.cuh (two 64-bit boards + a player bit).
"""
import abc
import dataclasses
import enum
from typing import Generic, List, Sequence, TypeVar

Action = int


class Outcome(enum.Enum):
    FIRST_PLAYER_WIN = 0.0
    DRAW = 0.5
    SECOND_PLAYER_WIN = 1.0


class CellState(enum.IntEnum):
    EMPTY = -1
    SECOND_PLAYER = 0
    FIRST_PLAYER = 1

    def opposite(self) -> "CellState":
        """The other colour; an empty cell stays empty."""
        return self if self is CellState.EMPTY else CellState(1 - self.value)


class Player(enum.IntEnum):
    SECOND = 0
    FIRST = 1

    def opposite(self) -> "Player":
        return Player(1 - self.value)

    def win(self) -> Outcome:
        """Outcome in which this player has won."""
        return _WIN_OF[self]

    def loss(self) -> Outcome:
        """Outcome in which this player has lost -- also the value the search gives an unvisited
        child of a node where this player is to move."""
        return _WIN_OF[self.opposite()]

    @staticmethod
    def max_player() -> "Player":
        return Player.SECOND

    @staticmethod
    def min_player() -> "Player":
        return Player.FIRST


_WIN_OF = {Player.FIRST: Outcome.FIRST_PLAYER_WIN, Player.SECOND: Outcome.SECOND_PLAYER_WIN}


@dataclasses.dataclass(unsafe_hash=True)
class State:
    """What every game state has in common: the player to move.  States are values: hashable,
    compared by content, never mutated (``dataclasses.replace`` makes variants)."""

    __slots__ = ["player"]
    player: Player


_S = TypeVar("_S", bound=State)
_M = TypeVar("_M", bound="GameManager")


class GameManager(abc.ABC, Generic[_S]):
    """The rules of one game on one board size.  ``num_actions`` counts every action id the game
    can ever use (board cells plus pass / swap where the game has one)."""

    grid_size: int
    num_actions: int

    def __init__(self, grid_size: int, num_actions: int) -> None:
        self.grid_size, self.num_actions = grid_size, num_actions

    # -- the rules ---------------------------------------------------------------------------
    @abc.abstractmethod
    def initial_game_state(self) -> _S:
        """The position a game starts from."""

    @abc.abstractmethod
    def legal_actions(self, state: _S) -> List[Action]:
        """Actions the player to move may take, in the reference's order."""

    @abc.abstractmethod
    def generate_child_state(self, state: _S, action: Action) -> _S:
        """The position after ``action``."""

    @abc.abstractmethod
    def is_final_state(self, state: _S) -> bool:
        """Whether the game is over."""

    @abc.abstractmethod
    def evaluate_final_state(self, state: _S) -> Outcome:
        """Who won a finished game."""

    # -- presentation ------------------------------------------------------------------------
    @abc.abstractmethod
    def probabilities_grid(self, action_probabilities: Sequence[float]) -> str:
        """A distribution over actions laid out like the board, for logs."""

    def action_str(self, action: Action) -> str:
        column, row = action % self.grid_size, action // self.grid_size
        return str((column, row))

    def copy(self: _M) -> _M:
        """A fresh manager of the same game and size (managers hold no game state)."""
        return type(self)(self.grid_size)
