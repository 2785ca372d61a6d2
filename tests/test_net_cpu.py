"""CPU checks of the host-side network mirror: the trainable torch module has the reference's
state_dict layout and reproduces the reference's recorded forward outputs (golden fixtures)."""
import numpy as np
import pytest
import torch

from oracle.games import make_game
from oracle.synth import synthetic_state_dict

NET_CFGS = [("tictactoe", 3, 3), ("hex", 7, 128), ("hex_with_swap", 7, 128), ("hex_with_swap", 5, 32),
            ("othello", 8, 128), ("othello", 6, 128)]


def build_net(name, g, channels, num_residual, hidden):
    from deep_mcts_b200.hex.convolutionalnet import ConvolutionalHexNet
    from deep_mcts_b200.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from deep_mcts_b200.tictactoe.convolutionalnet import ConvolutionalTicTacToeNet
    kw = dict(num_residual=num_residual, channels=channels, value_head_hidden_units=hidden)
    if name == "tictactoe":
        return ConvolutionalTicTacToeNet(**kw)
    cls = {"hex": ConvolutionalHexNet, "hex_with_swap": ConvolutionalHexWithSwapNet,
           "othello": ConvolutionalOthelloNet}[name]
    return cls(g, **kw)


@pytest.mark.parametrize("name,g,c", NET_CFGS)
def test_torch_module_matches_reference_fixture(golden_dir, name, g, c):
    z = np.load(golden_dir / f"net_{name}_{g}_c{c}.npz")
    channels, num_residual, hidden, seed = (int(v) for v in z["arch"])
    game = make_game(name, g)
    sd = synthetic_state_dict(game, channels, num_residual, hidden, seed)
    net = build_net(name, g, channels, num_residual, hidden)
    assert set(net.net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd)
    net.net.eval()
    m = net.manager
    states = [m._make_state(int(p), int(f), int(s)) for p, f, s in zip(z["player"], z["first"], z["second"])]
    planes = net.states_to_tensor(states)
    assert np.array_equal(planes.numpy(), z["planes"])
    with torch.no_grad():
        v, logits = net.forward(planes)
    np.testing.assert_allclose(v.numpy(), z["raw_value"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(logits.reshape(len(states), -1).numpy(), z["raw_logits"], atol=2e-5, rtol=0)


def test_train_step_and_copy_and_checkpoint(tmp_path):
    net = build_net("othello", 4, 8, 1, 16)
    m = net.manager
    n = 6
    states = torch.rand(n, 3, 4, 4)
    pi = torch.softmax(torch.rand(n, 17), dim=1)
    z = torch.rand(n, 1) * 2 - 1
    before = {k: v.clone() for k, v in net.net.state_dict().items()}
    pl, vl, acc = net.train(states, pi, z)
    assert pl.item() > 0 and vl.item() >= 0 and 0 <= acc.item() <= 1
    assert any(not torch.equal(before[k], v) for k, v in net.net.state_dict().items())
    twin = net.copy()
    for k, v in net.net.state_dict().items():
        assert torch.equal(v, twin.net.state_dict()[k])
    path = str(tmp_path / "anet-0.tar")
    net.save_full(path)
    loaded = type(net).from_path_full(path, m)
    assert loaded.channels == 8 and loaded.num_residual == 1 and loaded.value_head_hidden_units == 16
    for k, v in net.net.state_dict().items():
        assert torch.equal(v, loaded.net.state_dict()[k])
    assert set(net.parameters()) >= {"optimizer_cls", "state_dict", "grid_size", "channels"}
