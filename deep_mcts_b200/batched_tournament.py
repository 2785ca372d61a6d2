"""Batched agent matches: ``compare_agents`` for many games at once.

Same semantics as reference deep_mcts/tournament.py:55-108 with ``MCTSAgent`` players
(mcts.py:248-300) -- sides alternate between games, each agent keeps and re-roots its own tree
(``observe``), searches only on its own turn, plays argmax of the visit distribution or samples
from it with probability epsilon -- but every game of the match advances in lock-step on the
device: one ``BatchedMCTS`` per agent with one tree per game.
"""
import itertools
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from ._engine import BoardManager
from .mcts import BatchedMCTS
from .tournament import AgentComparison


@dataclass
class AgentSpec:
    """Constructor arguments of ``MCTSAgent(MCTS(manager, ...), epsilon)``."""
    num_simulations: int
    rollout_policy: Optional[Callable] = None
    state_evaluator: Optional[Callable] = None
    epsilon: float = 0.0
    rollout_share: float = 1.0
    leaves_per_round: int = 1


def play_batched(specs: Tuple[AgentSpec, AgentSpec], first_mover: Sequence[int], manager: BoardManager,
                 seed: int = 0):
    """Play len(first_mover) games; game k is opened by agent first_mover[k].  Returns the outcome
    values (0 FIRST win, 1 SECOND win, 0.5 draw) and per-agent (with, without) simulation stats."""
    n = len(first_mover)
    first_mover = np.asarray(first_mover, dtype=np.int64)
    agents = [BatchedMCTS(manager, n, sp.num_simulations, sp.rollout_policy, sp.state_evaluator,
                          rollout_share=sp.rollout_share, seed=seed + 101 * i, leaves_per_round=sp.leaves_per_round)
              for i, sp in enumerate(specs)]
    rs = np.random.RandomState(seed)
    p, f, s = manager.initial_batch(n)
    final, outcome = manager.status_batch(p, f, s)
    ply = 0
    stats = [[], []]
    while not final.all():
        # FIRST moves on even plies; the agent on turn in game k is first_mover[k] on even plies
        on_turn = np.where(ply % 2 == 0, first_mover, 1 - first_mover)
        for a in (0, 1):
            ids = np.where((on_turn == a) & ~final)[0].astype(np.int32)
            if ids.size == 0:
                continue
            agent = agents[a]
            agent.pool.observe(p[ids], f[ids], s[ids], tree_ids=ids)       # MCTSAgent.play re-rooting
            visits, st = agent.step(tree_ids=ids)
            stats[a].append(st[ids])
            v = visits[ids].astype(np.float64)
            actions = v.argmax(axis=1)
            eps = specs[a].epsilon
            if eps > 0:
                for j in range(ids.size):
                    if rs.random_sample() < eps:
                        actions[j] = rs.choice(v.shape[1], p=v[j] / v[j].sum())
            full = np.full(n, -1, dtype=np.int32)
            full[ids] = actions
            agent.pool.advance(full)
            cp, cf, cs, ok = manager.child_batch(p[ids], f[ids], s[ids], actions.astype(np.int32))
            assert ok.all()
            p[ids], f[ids], s[ids] = cp, cf, cs
        final, outcome = manager.status_batch(p, f, s)
        ply += 1
    for agent in agents:
        agent.pool.counters()   # surfaces capacity errors
        agent.pool.close()
    return outcome, stats


def compare_agents_batched(specs: Tuple[AgentSpec, AgentSpec], num_games: int, manager: BoardManager,
                           seed: int = 0) -> AgentComparison:
    """((win, draw, loss) rates of specs[0], each as (when moving first, when moving second)) over
    num_games // 2 games per side, like ``compare_agents``."""
    first_mover = np.arange(num_games) % 2          # even games: agent 0 opens (tournament.py:66-76)
    outcome, _ = play_batched(specs, first_mover, manager, seed)
    half = num_games // 2
    as_first = outcome[first_mover == 0]            # agent 0 is FIRST: wins at outcome 0
    as_second = outcome[first_mover == 1]           # agent 0 is SECOND: wins at outcome 1
    wins = ((as_first == 0.0).sum() / half, (as_second == 1.0).sum() / half)
    draws = ((as_first == 0.5).sum() / half, (as_second == 0.5).sum() / half)
    losses = ((as_first == 1.0).sum() / half, (as_second == 0.0).sum() / half)
    return (tuple(float(x) for x in wins), tuple(float(x) for x in draws), tuple(float(x) for x in losses))


def tournament_batched(specs: Sequence[AgentSpec], num_games: int, manager: BoardManager,
                       seed: int = 0) -> List[List[AgentComparison]]:
    """Round robin incl. self-matches, result[i][j] = comparison of agent i against agent j and the
    mirrored entry with wins and losses swapped (reference tournament.py:84-94)."""
    n = len(specs)
    zero = ((0.0, 0.0), (0.0, 0.0), (0.0, 0.0))
    results = [[zero for _ in range(n)] for _ in range(n)]
    for k, ((i, a), (j, b)) in enumerate(itertools.combinations_with_replacement(enumerate(specs), 2)):
        result = compare_agents_batched((a, b), num_games, manager, seed + 7919 * k)
        results[i][j] = result
        results[j][i] = (result[2], result[1], result[0])
    return results
