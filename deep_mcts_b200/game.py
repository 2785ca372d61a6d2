"""Game vocabulary of the engine: players, cells, outcomes, states and the manager interface.

Mirrors the reference's public names and values (reference deep_mcts/game.py:9-113) so code
written against it keeps working: ``Player.FIRST == 1`` moves first and minimises,
``Player.SECOND == 0`` maximises; tree values are absolute (``Outcome`` values 0 / 0.5 / 1
for a FIRST win / draw / SECOND win); cells hold -1 / 0 / 1.
"""
from abc import ABC, abstractmethod
from dataclasses import dataclass
from enum import Enum, IntEnum
from typing import Generic, List, Sequence, TypeVar


class Player(IntEnum):
    FIRST = 1
    SECOND = 0

    def opposite(self) -> "Player":
        return Player(1 - int(self))

    def win(self) -> "Outcome":
        return Outcome.FIRST_PLAYER_WIN if self is Player.FIRST else Outcome.SECOND_PLAYER_WIN

    def loss(self) -> "Outcome":
        return Outcome.SECOND_PLAYER_WIN if self is Player.FIRST else Outcome.FIRST_PLAYER_WIN

    @staticmethod
    def max_player() -> "Player":
        return Player.SECOND

    @staticmethod
    def min_player() -> "Player":
        return Player.FIRST


class CellState(IntEnum):
    FIRST_PLAYER = 1
    SECOND_PLAYER = 0
    EMPTY = -1

    def opposite(self) -> "CellState":
        if self is CellState.EMPTY:
            return self
        return CellState(1 - int(self))


class Outcome(Enum):
    FIRST_PLAYER_WIN = 0.0
    SECOND_PLAYER_WIN = 1.0
    DRAW = 0.5


@dataclass(unsafe_hash=True)
class State:
    """Base of every game state: who moves next.  Hashable, never mutated."""

    __slots__ = ["player"]
    player: Player


_S = TypeVar("_S", bound=State)
_T = TypeVar("_T", bound="GameManager")
Action = int


class GameManager(ABC, Generic[_S]):
    """Rules of one game (reference deep_mcts/game.py:76-113)."""

    grid_size: int
    num_actions: int

    def __init__(self, grid_size: int, num_actions: int) -> None:
        self.grid_size = grid_size
        self.num_actions = num_actions

    @abstractmethod
    def initial_game_state(self) -> _S: ...

    @abstractmethod
    def generate_child_state(self, state: _S, action: Action) -> _S: ...

    @abstractmethod
    def legal_actions(self, state: _S) -> List[Action]: ...

    @abstractmethod
    def is_final_state(self, state: _S) -> bool: ...

    @abstractmethod
    def evaluate_final_state(self, state: _S) -> Outcome: ...

    @abstractmethod
    def probabilities_grid(self, action_probabilities: Sequence[float]) -> str: ...

    def action_str(self, action: Action) -> str:
        return str((action % self.grid_size, action // self.grid_size))

    def copy(self: _T) -> _T:
        return type(self)(self.grid_size)
