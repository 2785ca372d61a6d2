"""torch-fp32 functional restatement of ConvolutionalNet leaf evaluation.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the one floating-point
kernel on the path, so the oracle keeps a torch fp32 reference for it
(tolerances: 1e-5 abs for the fp32 CUDA path, 1e-2 abs for bf16).

Works directly on a reference-format ``state_dict`` (key names in SURVEY.md
§8 a-N, produced by convolutionalnet.py:22-170) and on bitboard states.
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .games import HEX, HEX_SWAP, Game, State

BN_EPS = 1e-5  # nn.BatchNorm2d default used at convolutionalnet.py:82


def states_to_planes(game: Game, states: Sequence[State]) -> torch.Tensor:
    """[B,3,g,g] fp32: mover's stones, opponent's stones, constant player plane;
    Hex variants transpose the board when SECOND moves
    (othello/convolutionalnet.py:54-73, hex/convolutionalnet.py:64-94,
    hex_with_swap/convolutionalnet.py:67-97, tictactoe/convolutionalnet.py:51-68)."""
    g = game.g
    out = np.zeros((len(states), 3, g, g), dtype=np.float32)
    for b, (player, first, second) in enumerate(states):
        own, opp = (first, second) if player == 1 else (second, first)
        for y in range(g):
            for x in range(g):
                bit = y * g + x
                out[b, 0, y, x] = (own >> bit) & 1
                out[b, 1, y, x] = (opp >> bit) & 1
        out[b, 2] = float(player)
        if game.game in (HEX, HEX_SWAP) and player == 0:
            out[b, 0] = out[b, 0].T.copy()
            out[b, 1] = out[b, 1].T.copy()
    return torch.from_numpy(out)


def _conv_bn_relu(x, sd, prefix, padding):
    # ConvolutionalBlock, convolutionalnet.py:71-88 (BatchNorm in eval mode)
    x = F.conv2d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=padding)
    x = F.batch_norm(
        x, sd[prefix + "bn1.running_mean"], sd[prefix + "bn1.running_var"],
        sd[prefix + "bn1.weight"], sd[prefix + "bn1.bias"], False, 0.0, BN_EPS,
    )
    return F.relu(x)


def raw_forward(sd: Dict[str, torch.Tensor], planes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """ConvolutionalNet.forward (convolutionalnet.py:53-68): value [B,1], logits [B,A]."""
    x = _conv_bn_relu(planes, sd, "conv1.", 1)
    i = 0
    while f"residual_blocks.{i}.conv1.weight" in sd:
        p = f"residual_blocks.{i}."
        # ResidualBlock, convolutionalnet.py:109-129
        y = F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
        y = F.relu(F.batch_norm(y, sd[p + "bn1.running_mean"], sd[p + "bn1.running_var"],
                                sd[p + "bn1.weight"], sd[p + "bn1.bias"], False, 0.0, BN_EPS))
        y = F.conv2d(y, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
        y = F.batch_norm(y, sd[p + "bn2.running_mean"], sd[p + "bn2.running_var"],
                         sd[p + "bn2.weight"], sd[p + "bn2.bias"], False, 0.0, BN_EPS)
        x = F.relu(y + x)
        i += 1
    # PolicyHead, convolutionalnet.py:150-153
    p = _conv_bn_relu(x, sd, "policy_head.conv1.", 1)
    logits = F.linear(p.reshape(p.shape[0], -1), sd["policy_head.fc1.weight"], sd["policy_head.fc1.bias"])
    # ValueHead, convolutionalnet.py:165-170
    v = _conv_bn_relu(x, sd, "value_head.conv1.", 0)
    v = F.relu(F.linear(v.reshape(v.shape[0], -1), sd["value_head.fc1.weight"], sd["value_head.fc1.bias"]))
    v = F.linear(v, sd["value_head.fc2.weight"], sd["value_head.fc2.bias"])
    return v, logits


def forward(game: Game, sd: Dict[str, torch.Tensor], planes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """GameNet.forward including the Hex policy un-transpose for SECOND-to-move rows
    (hex/convolutionalnet.py:54-62, hex_with_swap/convolutionalnet.py:54-65)."""
    with torch.no_grad():
        v, logits = raw_forward(sd, planes.float())
    if game.game in (HEX, HEX_SWAP):
        g = game.g
        second = planes[:, 2, 0, 0] == 0
        logits = logits.clone()
        board = logits[:, : g * g].reshape(-1, g, g).clone()
        board[second] = board[second].transpose(1, 2).clone()
        logits[:, : g * g] = board.reshape(-1, g * g)
    return v, logits


def evaluate_batch(game: Game, sd: Dict[str, torch.Tensor], states: Sequence[State]) -> Tuple[np.ndarray, np.ndarray]:
    """GameNet.evaluate (gamenet.py:61-85) for a batch: absolute value in [0,1] and
    legal-masked, renormalised softmax priors, all in fp32."""
    planes = states_to_planes(game, states)
    v, logits = forward(game, sd, planes)
    v = torch.tanh(v[:, 0])
    sign = torch.tensor([-1.0 if s[0] == 1 else 1.0 for s in states])  # negate for FIRST (min player)
    v = (v * sign + 1) / 2
    probs = F.softmax(logits.reshape(len(states), -1), dim=1)
    mask = torch.zeros_like(probs)
    for b, s in enumerate(states):
        mask[b, game.legal_list(s)] = 1.0
    probs = probs * mask
    probs = probs / probs.sum(dim=1, keepdim=True)
    return v.numpy().astype(np.float32), probs.numpy().astype(np.float32)


def make_evaluator(sd: Dict[str, torch.Tensor]):
    """A per-state evaluator with the reference's memoisation
    (cached_state_evaluator, train.py:408-414): batch-1 forward, results widened to
    Python floats."""
    cache: Dict[Tuple[int, State], Tuple[float, List[float]]] = {}

    def evaluator(game: Game, state: State):
        key = (game.game, state)
        hit = cache.get(key)
        if hit is None:
            v, p = evaluate_batch(game, sd, [state])
            hit = (float(v[0]), p[0].tolist())
            cache[key] = hit
            evaluator.misses += 1
        return hit

    evaluator.misses = 0
    evaluator.cache = cache
    return evaluator
