"""Reproducible synthetic weights and positions shared by the golden generator,
the tests and bench.py.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Weights come from numpy's legacy ``RandomState`` (stream frozen by numpy's
compatibility policy), not from torch's RNG, so the same state_dict can be
rebuilt on any box without committing megabytes of parameters.  BatchNorm
running statistics and affine terms are randomised too, so that BN folding is
exercised (a freshly constructed reference net has identity BN).
"""
from typing import Dict, List

import numpy as np
import torch

from .games import Game, State


def synthetic_state_dict(game: Game, channels: int = 128, num_residual: int = 3,
                         hidden: int = 128, seed: int = 0, conv_gain: float = 1.6,
                         fc_scale: float = 2.0) -> Dict[str, torch.Tensor]:
    """A state_dict with the reference's key names and shapes (SURVEY.md §8 a-N).

    ``conv_gain`` / ``fc_scale`` set the weight standard deviation as gain / sqrt(fan_in).  The
    defaults (1.6, 2.0) give deliberately sharp policies (used by the golden fixtures and the
    fp32 tests); (0.577, 0.577) has the spread of a freshly constructed reference net
    (PyTorch default init, kaiming_uniform(a=sqrt(5)) => std = 1 / sqrt(3 fan_in))."""
    rs = np.random.RandomState(seed)
    g, A, C = game.g, game.num_actions, channels
    sd: Dict[str, torch.Tensor] = {}

    def conv(prefix, cout, cin, k):
        fan_in = cin * k * k
        sd[prefix + ".weight"] = torch.from_numpy(
            (rs.standard_normal((cout, cin, k, k)) * (conv_gain / np.sqrt(fan_in))).astype(np.float32))
        sd[prefix + ".bias"] = torch.from_numpy((rs.standard_normal(cout) * 0.1).astype(np.float32))

    def bn(prefix, c):
        sd[prefix + ".weight"] = torch.from_numpy((1.0 + 0.2 * rs.standard_normal(c)).astype(np.float32))
        sd[prefix + ".bias"] = torch.from_numpy((0.1 * rs.standard_normal(c)).astype(np.float32))
        sd[prefix + ".running_mean"] = torch.from_numpy((0.1 * rs.standard_normal(c)).astype(np.float32))
        sd[prefix + ".running_var"] = torch.from_numpy((0.5 + rs.random_sample(c)).astype(np.float32))
        sd[prefix + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    def linear(prefix, cout, cin, scale=1.0):
        sd[prefix + ".weight"] = torch.from_numpy(
            (rs.standard_normal((cout, cin)) * (scale / np.sqrt(cin))).astype(np.float32))
        sd[prefix + ".bias"] = torch.from_numpy((rs.standard_normal(cout) * 0.1).astype(np.float32))

    conv("conv1.conv1", C, 3, 3)
    bn("conv1.bn1", C)
    for i in range(num_residual):
        conv(f"residual_blocks.{i}.conv1", C, C, 3)
        bn(f"residual_blocks.{i}.bn1", C)
        conv(f"residual_blocks.{i}.conv2", C, C, 3)
        bn(f"residual_blocks.{i}.bn2", C)
    conv("policy_head.conv1.conv1", C, C, 3)
    bn("policy_head.conv1.bn1", C)
    linear("policy_head.fc1", A, C * g * g, scale=fc_scale)
    conv("value_head.conv1.conv1", 1, C, 1)
    bn("value_head.conv1.bn1", 1)
    linear("value_head.fc1", hidden, g * g)
    linear("value_head.fc2", 1, hidden)
    return sd


def random_positions(game: Game, n: int, seed: int, include_final: bool = False) -> List[State]:
    """Reachable positions from seeded uniform-random playouts, each truncated at a
    uniformly drawn ply (SURVEY.md §8d 'Synthetic positions')."""
    rs = np.random.RandomState(seed)
    out: List[State] = []
    while len(out) < n:
        state = game.initial_state()
        trail = [state]
        while not game.is_final(state):
            legal = game.legal_list(state)
            state = game.child(state, legal[rs.randint(len(legal))])
            trail.append(state)
        if not include_final:
            trail = trail[:-1]
        out.append(trail[rs.randint(len(trail))])
    return out
