"""GPU test of the training loop on the batched engine: a tiny Othello 4x4 run
must play games with the uniform evaluator, train, transfer weights, write the reference's
artefacts and evaluate."""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_bitboards_to_planes_matches_states_to_tensor():
    from deep_mcts_b200.hex.convolutionalnet import ConvolutionalHexNet
    from deep_mcts_b200.train import bitboards_to_planes
    from oracle.games import make_game
    from oracle.synth import random_positions
    game = make_game("hex", 5)
    net = ConvolutionalHexNet(5, num_residual=1, channels=8, value_head_hidden_units=8)
    states = random_positions(game, 40, seed=1)
    objs = [net.manager._make_state(*s) for s in states]
    want = net.states_to_tensor(objs)
    dev = torch.device("cuda")
    f = torch.from_numpy(np.array([s[1] for s in states], np.uint64).view(np.int64)).to(dev)
    s2 = torch.from_numpy(np.array([s[2] for s in states], np.uint64).view(np.int64)).to(dev)
    p = torch.from_numpy(np.array([s[0] for s in states], np.uint8)).to(dev)
    got = bitboards_to_planes(f, s2, p, 5, True).cpu()
    assert torch.equal(got, want)


def test_tiny_training_run(tmp_path):
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from deep_mcts_b200.train import EVALUATION_COLUMNS, TrainingConfiguration, train
    torch.manual_seed(0)
    net = ConvolutionalOthelloNet(4, num_residual=1, channels=16, value_head_hidden_units=16)
    before = {k: v.clone() for k, v in net.net.state_dict().items()}
    cfg = TrainingConfiguration(
        num_games=8, num_simulations=10, save_interval=4, evaluation_interval=8, save_dir=str(tmp_path),
        sample_move_cutoff=4, dirichlet_alpha=1.0, nprocs=2, batch_size=32, replay_buffer_max_size=100,
        evaluation_games=2, transfer_interval=4, n_trees=64, train_steps_per_block=2)
    train(net, cfg)
    assert (tmp_path / "anet-0.tar").exists() and (tmp_path / "anet-4.tar").exists()
    params = json.load(open(tmp_path / "parameters.json"))
    assert params["training_configuration"]["num_simulations"] == 10
    assert params["net_parameters"]["optimizer_cls"] == "SGD"
    header = open(tmp_path / "evaluations.csv").readline().strip().split(",")
    assert header[1:] == EVALUATION_COLUMNS
    assert any(not torch.equal(before[k].to(v.device), v) for k, v in net.net.state_dict().items())
    loaded = ConvolutionalOthelloNet.from_path_full(str(tmp_path / "anet-4.tar"), net.manager)
    assert loaded.channels == 16
