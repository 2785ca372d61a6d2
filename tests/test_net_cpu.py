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


@pytest.mark.parametrize("name,g,policy_shape", [("othello", 4, (17,)), ("hex", 4, (4, 4)), ("hex_with_swap", 4, (17,)),
                                                  ("tictactoe", 3, (3, 3))])
def test_every_game_net_copies_and_round_trips_through_a_full_checkpoint(tmp_path, name, g, policy_shape):
    """copy() / save_full / from_path_full / parameters() of all four per-game nets (one shared body,
    deep_mcts_b200/boardnet.py), with the reference's constructor conventions: positional grid size
    for the board games, none for tic-tac-toe."""
    net = build_net(name, g, 8, 1, 16)
    with torch.no_grad():
        value, logits = net.forward(torch.rand(5, 3, g, g))
    assert value.shape == (5, 1) and tuple(logits.shape[1:]) == policy_shape
    twin = net.copy()
    assert type(twin) is type(net) and twin.manager is net.manager
    path = str(tmp_path / "anet-0.tar")
    net.save_full(path)
    loaded = type(net).from_path_full(path, net.manager)
    for other in (twin, loaded):
        assert (other.channels, other.num_residual, other.value_head_hidden_units, other.grid_size) == (8, 1, 16, g)
        for k, v in net.net.state_dict().items():
            assert torch.equal(v, other.net.state_dict()[k])
    keys = set(net.parameters())
    assert keys >= {"optimizer_cls", "optimizer_args", "optimizer_kwargs", "state_dict", "num_residual", "channels",
                    "value_head_hidden_units"}
    assert ("grid_size" in keys) == (name != "tictactoe")     # the reference's tic-tac-toe net saves no grid size


def test_hex_policy_is_transposed_back_for_second_to_move():
    """Rows whose player plane says SECOND are fed a transposed board; forward() must hand their policy back
    in board coordinates (reference hex/convolutionalnet.py:54-62, hex_with_swap/convolutionalnet.py:54-65)."""
    from deep_mcts_b200.game import Player
    for name, flat in (("hex", False), ("hex_with_swap", True)):
        net = build_net(name, 4, 8, 1, 16)
        planes = torch.rand(6, 3, 4, 4)
        planes[:, 2] = torch.tensor([1., 0., 1., 0., 0., 1.]).view(6, 1, 1)
        with torch.no_grad():
            _, logits = net.forward(planes)
            _, raw = net.net(planes)
        second = planes[:, 2, 0, 0] == int(Player.SECOND)
        if flat:
            assert torch.equal(logits[:, 16], raw[:, 16])                        # the swap logit is not moved
            board, raw_board = logits[:, :16].reshape(6, 4, 4), raw[:, :16].reshape(6, 4, 4)
        else:
            board, raw_board = logits, raw
        assert torch.equal(board[~second], raw_board[~second])
        assert torch.equal(board[second], raw_board[second].transpose(1, 2))
