"""Agents and match play (interface of reference deep_mcts/tournament.py:16-108).

Host-side control flow only: the agents' decisions come from the CUDA search
(``MCTSAgent``) or from the game engines' legal-move lists (``RandomAgent``).
"""
import itertools
import random
from abc import ABC, abstractmethod
from typing import List, Sequence, Tuple

from .game import Action, GameManager, Outcome

AgentComparison = Tuple[Tuple[float, float], Tuple[float, float], Tuple[float, float]]


class Agent(ABC):
    @abstractmethod
    def play(self, state) -> Action: ...

    @abstractmethod
    def reset(self) -> None: ...


class RandomAgent(Agent):
    def __init__(self, manager: GameManager) -> None:
        self.manager = manager

    def play(self, state) -> Action:
        return random.choice(self.manager.legal_actions(state))

    def reset(self) -> None:
        pass


def play(players: Tuple[Agent, Agent], game_manager: GameManager) -> Outcome:
    """One game; players[0] moves first.  Every agent is reset afterwards."""
    state = game_manager.initial_game_state()
    turn = 0
    while not game_manager.is_final_state(state):
        state = game_manager.generate_child_state(state, players[turn].play(state))
        turn ^= 1
    for agent in players:
        agent.reset()
    return game_manager.evaluate_final_state(state)


def compare_agents(players: Tuple[Agent, Agent], num_games: int, game_manager: GameManager) -> AgentComparison:
    """Alternate who moves first; report players[0]'s (win, draw, loss) rates as
    ((as first, as second), ...) over num_games // 2 games per side."""
    tally = {"first": [0, 0, 0], "second": [0, 0, 0]}  # win, draw, loss of players[0]
    for k in range(num_games):
        if k % 2 == 0:
            result = play(players, game_manager)
            mine, theirs, side = Outcome.FIRST_PLAYER_WIN, Outcome.SECOND_PLAYER_WIN, "first"
        else:
            result = play((players[1], players[0]), game_manager)
            mine, theirs, side = Outcome.SECOND_PLAYER_WIN, Outcome.FIRST_PLAYER_WIN, "second"
        tally[side][0 if result == mine else (2 if result == theirs else 1)] += 1
    half = num_games // 2
    return tuple((tally["first"][i] / half, tally["second"][i] / half) for i in range(3))


def tournament(agents: Sequence[Agent], num_games: int, game_manager: GameManager) -> List[List[AgentComparison]]:
    n = len(agents)
    zero = ((0.0, 0.0), (0.0, 0.0), (0.0, 0.0))
    results = [[zero for _ in range(n)] for _ in range(n)]
    for (i, a), (j, b) in itertools.combinations_with_replacement(enumerate(agents), 2):
        result = compare_agents((a, b), num_games, game_manager)
        results[i][j] = result
        results[j][i] = (result[2], result[1], result[0])
    return results
