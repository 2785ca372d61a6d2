"""Batched tournaments: lock-step matches must reproduce the sequential
compare_agents / MCTSAgent results for deterministic agents and show sane strength ordering."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_games_gpu import manager_for  # noqa: E402


def test_batched_match_equals_sequential_compare_agents():
    from deep_mcts_b200 import mcts as dm
    from deep_mcts_b200.batched_tournament import AgentSpec, compare_agents_batched
    from deep_mcts_b200.tournament import compare_agents
    manager = manager_for("othello", 6)
    spec_a = AgentSpec(30, None, dm.uniform_state_evaluator(manager))
    spec_b = AgentSpec(20, None, dm.hashed_state_evaluator(manager, 0))
    batched = compare_agents_batched((spec_a, spec_b), 4, manager)
    seq = compare_agents(
        (dm.MCTSAgent(dm.MCTS(manager, 30, None, dm.uniform_state_evaluator(manager))),
         dm.MCTSAgent(dm.MCTS(manager, 20, None, dm.hashed_state_evaluator(manager, 0)))), 4, manager)
    assert batched == seq


def test_stronger_agent_wins_more():
    from deep_mcts_b200 import mcts as dm
    from deep_mcts_b200.batched_tournament import AgentSpec, compare_agents_batched
    manager = manager_for("hex", 5)
    strong = AgentSpec(300, dm.random_rollout_policy(manager), None)
    weak = AgentSpec(4, dm.random_rollout_policy(manager), None)
    (w1, w2), (d1, d2), (l1, l2) = compare_agents_batched((strong, weak), 64, manager, seed=1)
    assert d1 == d2 == 0.0                      # no draws in hex
    assert (w1 + w2) / 2 > 0.8, ((w1, w2), (l1, l2))
    assert abs(w1 + l1 - 1) < 1e-9 and abs(w2 + l2 - 1) < 1e-9


def test_begin_step_ids_leaves_other_trees_untouched():
    from deep_mcts_b200 import mcts as dm
    manager = manager_for("tictactoe", 3)
    b = dm.BatchedMCTS(manager, 8, 20, None, dm.uniform_state_evaluator(manager))
    visits, stats = b.step(tree_ids=[1, 5])
    assert stats[[1, 5]].sum() == 40 and stats[[0, 2, 3, 4, 6, 7]].sum() == 0
    assert visits[[0, 2, 3, 4, 6, 7]].sum() == 0 and visits[1].sum() == 20


def test_network_policy_agents_play_legal_games():
    """GreedyGameNetAgent / SamplingGameNetAgent (declared but unimplemented in the reference, gamenet.py:196-210) play
    whole games of legal moves through the sequential tournament driver."""
    from deep_mcts_b200.gamenet import GreedyGameNetAgent, SamplingGameNetAgent
    from deep_mcts_b200.tictactoe.convolutionalnet import ConvolutionalTicTacToeNet
    from deep_mcts_b200.tournament import RandomAgent, compare_agents
    net = ConvolutionalTicTacToeNet()
    manager = net.manager
    for agent in (GreedyGameNetAgent(net), GreedyGameNetAgent(net, epsilon=0.5), SamplingGameNetAgent(net)):
        wins, draws, losses = compare_agents((agent, RandomAgent(manager)), 6, manager)
        assert abs(sum(wins) + sum(draws) + sum(losses) - 2.0) < 1e-9
