"""Config 4 result parity (BASELINE.json configs[3]; reference evaluate_complex_rollouts.py:53-90): the complex agent
(MCTS with greedy policy-network rollouts) against the simple agent (uniform random rollouts), with and without the
network as state evaluator, rollout_share = 1, fixed simulation budget, sides alternating.  The same pairing played
game by game on the oracle (tests/golden/make_config4_golden.py, committed outcomes) and as one batched tournament on
the device must give the complex agent the same win rates as first and as second mover, within binomial bounds."""
import json

import numpy as np
import pytest

from oracle.games import make_game
from oracle.synth import synthetic_state_dict

pytestmark = pytest.mark.gpu

CASES = ["small7", "full6"]


def run_device_pairing(cfg, with_evaluator, games, seed, precision="fp32"):
    from deep_mcts_b200 import mcts as dm
    from deep_mcts_b200.batched_tournament import AgentSpec, play_batched
    from deep_mcts_b200.gamenet import cached_state_evaluator
    from deep_mcts_b200.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet
    game = make_game("hex_with_swap", cfg["grid"])
    sd = synthetic_state_dict(game, cfg["channels"], cfg["num_residual"], cfg["hidden"], seed=cfg["seed"],
                              conv_gain=1.0, fc_scale=1.0)
    net = ConvolutionalHexWithSwapNet(cfg["grid"], num_residual=cfg["num_residual"], channels=cfg["channels"],
                                      value_head_hidden_units=cfg["hidden"])
    net.load_state_dict(sd)
    net.precision = precision
    net.max_batch = games
    manager = net.manager
    evaluator = cached_state_evaluator(net) if with_evaluator else None
    complex_spec = AgentSpec(cfg["sims"], dm.greedy_policy_rollout(net), evaluator, rollout_share=1.0)
    simple_spec = AgentSpec(cfg["sims"], dm.random_rollout_policy(manager), evaluator, rollout_share=1.0)
    first_mover = np.arange(games) % 2
    outcome, stats = play_batched((complex_spec, simple_spec), first_mover, manager, seed=seed)
    as_first, as_second = outcome[first_mover == 0], outcome[first_mover == 1]
    assert set(np.unique(outcome).tolist()) <= {0.0, 1.0}          # no draws in Hex
    return {"complex_wins_as_first": int((as_first == 0.0).sum()), "games_as_first": int(as_first.size),
            "complex_wins_as_second": int((as_second == 1.0).sum()), "games_as_second": int(as_second.size)}


def within_binomial_bounds(k1, n1, k2, n2, sigmas):
    p1, p2 = k1 / n1, k2 / n2
    pooled = (k1 + k2) / (n1 + n2)
    sigma = np.sqrt(max(pooled * (1 - pooled), 1.0 / (n1 + n2)) * (1 / n1 + 1 / n2))
    return abs(p1 - p2) <= sigmas * sigma, p1, p2, sigma


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("with_evaluator", [False, True])
def test_complex_vs_simple_rollout_agents_match_the_oracle(golden_dir, case, with_evaluator):
    path = golden_dir / f"config4_{case}.json"
    if not path.exists():
        pytest.skip(f"{path.name} has not been generated")
    data = json.load(open(path))
    name = "with_state_evaluator" if with_evaluator else "without_state_evaluator"
    if name not in data["pairings"]:
        pytest.skip(f"{path.name} has no {name} pairing yet")
    want = data["pairings"][name]
    cfg = data["config"]
    games = 480 if case == "small7" else 240
    # small7 runs the fp32 kernels (same arithmetic as the oracle's network); full6 runs the bf16 tensor-core path
    # (the fp32 CUDA-core path needs two minutes per pairing at C=128; tests/config4_parity_record.py records both)
    got = run_device_pairing(cfg, with_evaluator, games, seed=3, precision="fp32" if case == "small7" else "bf16")
    for side in ("first", "second"):
        ok, p_dev, p_cpu, sigma = within_binomial_bounds(got[f"complex_wins_as_{side}"], got[f"games_as_{side}"],
                                                         want[f"complex_wins_as_{side}"], want[f"games_as_{side}"], 4.0)
        assert ok, (case, name, side, p_dev, p_cpu, sigma)
