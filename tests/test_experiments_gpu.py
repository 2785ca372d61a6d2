"""The three experiment drivers (reference evaluate_training.py, evaluate_simple_rollouts.py,
evaluate_complex_rollouts.py) on a two-checkpoint run with tiny budgets: artefact locations, schemas
and value ranges."""
import json

import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def run_dir(tmp_path_factory):
    from deep_mcts_b200.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet
    root = tmp_path_factory.mktemp("experiment")
    run = root / "saves" / "run0"
    run.mkdir(parents=True)
    for iteration, seed in ((0, 0), (10, 1)):
        torch.manual_seed(seed)
        ConvolutionalHexWithSwapNet(4).save_full(str(run / f"anet-{iteration}.tar"))
    return root


def rates_ok(comparison):
    (w1, w2), (d1, d2), (l1, l2) = comparison
    return abs(w1 + d1 + l1 - 1.0) < 1e-9 and abs(w2 + d2 + l2 - 1.0) < 1e-9


def test_evaluate_training_writes_reference_csv(run_dir):
    from deep_mcts_b200.evaluate_training import COLUMNS, evaluate_training
    from deep_mcts_b200.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet
    from deep_mcts_b200.hex_with_swap.game import HexWithSwapManager
    evaluate_training(run_dir / "saves", ConvolutionalHexWithSwapNet, HexWithSwapManager(4), torch.device("cuda:0"),
                      num_games=8, num_simulations=12)
    df = pd.read_csv(run_dir / "training" / "run0.csv", index_col=0)
    assert list(df.columns) == COLUMNS and list(df["training_iterations"]) == [0, 10]
    first, second = df.iloc[0], df.iloc[1]
    assert all(first[c] == 0.0 for c in COLUMNS if c.startswith("previous_"))     # nothing to compare the first one with
    for row in (first, second):
        assert abs(row["random_first_wins"] + row["random_first_draws"] + row["random_first_losses"] - 1.0) < 1e-9
    assert abs(second["previous_second_wins"] + second["previous_second_draws"] + second["previous_second_losses"] - 1.0) < 1e-9


def test_evaluate_simple_rollouts_round_robin(run_dir):
    from deep_mcts_b200.evaluate_simple_rollouts import evaluate_simple_rollouts
    from deep_mcts_b200.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet
    from deep_mcts_b200.hex_with_swap.game import HexWithSwapManager
    evaluate_simple_rollouts(run_dir / "saves", ConvolutionalHexWithSwapNet, HexWithSwapManager(4),
                             torch.device("cuda:0"), num_games=4, num_simulations=10)
    results = json.load(open(run_dir / "simple_rollouts" / "run0.json"))
    assert len(results) == 6 and all(len(r) == 6 for r in results)
    for i in range(6):
        for j in range(6):
            assert rates_ok(results[i][j])
            assert results[j][i] == [results[i][j][2], results[i][j][1], results[i][j][0]] or i == j


def test_evaluate_complex_rollouts_time_budget(run_dir):
    from deep_mcts_b200.evaluate_complex_rollouts import evaluate_complex_rollouts
    from deep_mcts_b200.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet
    from deep_mcts_b200.hex_with_swap.game import HexWithSwapManager
    evaluate_complex_rollouts(run_dir / "saves", ConvolutionalHexWithSwapNet, HexWithSwapManager(4),
                              torch.device("cuda:0"), num_games=2, time_per_move=0.05)
    for name in ("with_state_evaluator", "without_state_evaluator"):
        out = json.load(open(run_dir / "complex_rollouts" / name / "run0.json"))
        assert set(out) == {"results", "complex_simulations", "simple_simulations"} and rates_ok(out["results"])
        # one list of (with, without) pairs per game, plus the list opened by the last reset
        assert len(out["complex_simulations"]) == 3 and len(out["simple_simulations"]) == 3
        assert sum(a + b for a, b in out["simple_simulations"][0]) > 0
