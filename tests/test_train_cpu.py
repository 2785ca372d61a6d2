"""CPU pins of the training-side rows against fixtures written by the UNMODIFIED reference
(tests/golden/make_train_golden.py): a checkpoint saved by the reference's own ``GameNet.save_full`` must load
through ``from_path_full`` (reference gamenet.py:130-150, othello/convolutionalnet.py:88-113), and
``GameNet.train`` must reproduce the reference's losses and updated weights step for step (gamenet.py:87-110)."""
import numpy as np
import pytest
import torch

CASES = [("othello", 6), ("hex", 5), ("hex_with_swap", 5), ("tictactoe", 3)]


def net_class(name):
    from deep_mcts_b200.hex.convolutionalnet import ConvolutionalHexNet
    from deep_mcts_b200.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from deep_mcts_b200.tictactoe.convolutionalnet import ConvolutionalTicTacToeNet
    return {"hex": ConvolutionalHexNet, "hex_with_swap": ConvolutionalHexWithSwapNet,
            "othello": ConvolutionalOthelloNet, "tictactoe": ConvolutionalTicTacToeNet}[name]


def load_reference_checkpoint(golden_dir, name):
    return net_class(name).from_path_full(str(golden_dir / f"anet-0-{name}.tar"))


@pytest.mark.parametrize("name,g", CASES)
def test_reference_written_checkpoint_loads(golden_dir, name, g):
    z = np.load(golden_dir / f"train_{name}.npz")
    channels, num_residual, hidden, grid = (int(v) for v in z["arch"])
    net = load_reference_checkpoint(golden_dir, name)
    assert (net.channels, net.num_residual, net.value_head_hidden_units) == (channels, num_residual, hidden)
    assert net.manager.board == grid and net.optimizer_cls is torch.optim.SGD
    assert dict(net.optimizer_kwargs) == {"lr": 0.01, "momentum": 0.9, "weight_decay": 0.001}
    raw = torch.load(str(golden_dir / f"anet-0-{name}.tar"), map_location="cpu", weights_only=False)
    assert set(raw["state_dict"]) == set(net.net.state_dict())
    for key, t in raw["state_dict"].items():
        assert torch.equal(t, net.net.state_dict()[key]), key
    # the planes our states_to_tensor builds for the recorded positions are the reference's
    m = net.manager
    states = [m._make_state(int(p), int(f), int(s)) for p, f, s in zip(z["player"], z["first"], z["second"])]
    assert np.array_equal(net.states_to_tensor(states).numpy(), z["planes"])


@pytest.mark.parametrize("name,g", CASES)
def test_train_steps_match_the_reference(golden_dir, name, g):
    """Two consecutive SGD steps (momentum, weight decay, BatchNorm batch statistics and running-stat updates) from
    the checkpoint's weights on the recorded batch: losses, accuracy and every tensor of the state_dict."""
    z = np.load(golden_dir / f"train_{name}.npz")
    net = load_reference_checkpoint(golden_dir, name)
    planes, pi, target = torch.from_numpy(z["planes"]), torch.from_numpy(z["pi"]), torch.from_numpy(z["z"])
    for step in range(2):
        pl, vl, acc = net.train(planes, pi, target)
        got = np.array([pl.item(), vl.item(), acc.item()])
        np.testing.assert_allclose(got, z["losses"][step], atol=1e-6, rtol=0)
        sd = net.net.state_dict()
        for key, t in sd.items():
            want = z[f"after{step}/{key}"]
            if key.endswith("num_batches_tracked"):
                assert int(t) == int(want)
            else:
                np.testing.assert_allclose(t.detach().numpy(), want, atol=1e-6, rtol=0, err_msg=f"{key} step {step}")
