"""Checkpoint and training-step fixtures from the UNMODIFIED reference.

Run in the build container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_train_golden.py

Outputs (committed), one pair per game, small architectures so the files stay in the kilobytes:
  anet-0-<game>.tar   written by the reference's own ``GameNet.save_full`` (gamenet.py:130-150 and the per-game
                      ``parameters()``, e.g. othello/convolutionalnet.py:107-114): the pickled dict with
                      optimizer class / args / kwargs, architecture keys and the state_dict
  train_<game>.npz    ``GameNet.evaluate`` outputs of that checkpoint on reachable positions (gamenet.py:61-85),
                      and TWO consecutive reference ``GameNet.train`` steps (gamenet.py:87-110; SGD lr .01,
                      momentum .9, weight decay .001, BatchNorm in train mode) from the checkpoint's weights on a
                      recorded batch: the three returned scalars per step and every state_dict tensor afterwards
"""
import random
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

from deep_mcts.hex.convolutionalnet import ConvolutionalHexNet  # noqa: E402
from deep_mcts.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet  # noqa: E402
from deep_mcts.othello.convolutionalnet import ConvolutionalOthelloNet  # noqa: E402
from deep_mcts.tictactoe.convolutionalnet import ConvolutionalTicTacToeNet  # noqa: E402

from make_golden import to_bits  # noqa: E402  (reference state -> (player, first, second) bitboards)

CASES = [
    # name, grid, constructor, architecture
    ("othello", 6, lambda kw: ConvolutionalOthelloNet(6, **kw), dict(num_residual=2, channels=16, value_head_hidden_units=32)),
    ("hex", 5, lambda kw: ConvolutionalHexNet(5, **kw), dict(num_residual=1, channels=16, value_head_hidden_units=32)),
    ("hex_with_swap", 5, lambda kw: ConvolutionalHexWithSwapNet(5, **kw), dict(num_residual=1, channels=8, value_head_hidden_units=16)),
    ("tictactoe", 3, lambda kw: ConvolutionalTicTacToeNet(**kw), dict(num_residual=3, channels=3, value_head_hidden_units=128)),
]
BATCH = 24


def reachable_states(manager, n, seed):
    rng = random.Random(seed)
    out = []
    while len(out) < n:
        s = manager.initial_game_state()
        for _ in range(rng.randrange(0, 3 * manager.grid_size)):
            if manager.is_final_state(s):
                break
            s = manager.generate_child_state(s, rng.choice(manager.legal_actions(s)))
        if not manager.is_final_state(s):
            out.append(s)
    return out


def main():
    for k, (name, g, make, arch) in enumerate(CASES):
        torch.manual_seed(100 + k)
        net = make(arch)
        # give the BatchNorm layers non-trivial running statistics and affine parameters
        with torch.no_grad():
            for key, t in net.net.state_dict().items():
                if key.endswith("running_mean"):
                    t.copy_(torch.randn_like(t) * 0.1)
                elif key.endswith("running_var"):
                    t.copy_(torch.rand_like(t) * 0.5 + 0.75)
                elif ".bn" in key and key.endswith("weight"):
                    t.copy_(torch.rand_like(t) * 0.5 + 0.75)
                elif ".bn" in key and key.endswith("bias"):
                    t.copy_(torch.randn_like(t) * 0.1)
        path = HERE / f"anet-0-{name}.tar"
        net.save_full(str(path))
        manager = net.manager
        states = reachable_states(manager, BATCH, seed=7 + k)
        bits = np.array([to_bits(s) for s in states], dtype=np.uint64)
        # evaluate (eval mode) on the checkpoint's weights
        net.net.eval()
        values, probs = [], []
        for s in states:
            v, p = net.evaluate(s)
            values.append(v)
            probs.append(p.numpy())
        # two training steps
        planes = net.states_to_tensor(states)
        gen = torch.Generator().manual_seed(900 + k)
        A = manager.num_actions
        pi = torch.softmax(torch.randn(BATCH, A, generator=gen) * 2, dim=1)
        z = (torch.randint(0, 3, (BATCH, 1), generator=gen).float() - 1.0)
        pi_target = net.distributions_to_tensor(pi.tolist())
        if name in ("hex", "tictactoe"):     # these nets' policy output is [g, g] (hex/train.py, tictactoe/train.py)
            pi_target = pi_target.reshape(BATCH, g, g)
        steps = []
        after = {}
        for step in range(2):
            pl, vl, acc = net.train(planes, pi_target, z)
            steps.append([pl.item(), vl.item(), acc.item()])
            after[step] = {key: t.detach().clone().numpy() for key, t in net.net.state_dict().items()}
        out = {"arch": np.array([arch["channels"], arch["num_residual"], arch["value_head_hidden_units"], g]),
               "player": bits[:, 0].astype(np.uint8), "first": bits[:, 1], "second": bits[:, 2],
               "eval_value": np.array(values, np.float64), "eval_probs": np.array(probs, np.float32),
               "planes": planes.numpy(), "pi": pi_target.numpy(), "z": z.numpy(),
               "losses": np.array(steps, np.float64)}
        for step in range(2):
            for key, t in after[step].items():
                out[f"after{step}/{key}"] = t
        np.savez_compressed(HERE / f"train_{name}.npz", **out)
        print(name, "checkpoint", path.stat().st_size, "bytes; losses", steps)


if __name__ == "__main__":
    main()
