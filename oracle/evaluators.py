"""Deterministic state evaluators used to pin MCTS visit counts.

TEST INFRASTRUCTURE (see oracle/__init__.py).

``uniform`` restates train.py:240-248 (the evaluator every self-play worker uses
until the first weight transfer).  ``hashed`` is a synthetic evaluator defined
purely by 64-bit integer mixing and one IEEE double division per prior, so the
CUDA engine's built-in copy (csrc/tree.cu, evaluator kind DM_EVAL_HASHED) can
reproduce it bit for bit; it exists only to give the tree non-uniform,
state-dependent priors and values in parity tests.
"""
from typing import List, Sequence, Tuple

from .games import Game, State

_M64 = (1 << 64) - 1


def mix64(x: int) -> int:
    """splitmix64 finaliser."""
    x &= _M64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & _M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & _M64
    x ^= x >> 31
    return x


def uniform(game: Game, state: State) -> Tuple[float, List[float]]:
    """The evaluator self-play uses before the first weight transfer: value 0.5, equal priors over the
    legal actions (reference train.py:240-248)."""
    legal = set(game.legal_list(state))
    n = len(legal)
    return 0.5, [1 / n if a in legal else 0.0 for a in range(game.num_actions)]


def hashed_salted(salt: int):
    """Deterministic synthetic evaluator (a test INPUT, not a restatement of reference code): priors and value
    are a hash of the state, so the reference MCTS, the oracle and the device engine can be fed identical
    numbers (the device twin is tree.cu hashed_h0 / hashed_weight)."""
    def evaluator(game: Game, state: State) -> Tuple[float, List[float]]:
        player, first, second = state
        h0 = mix64(first ^ mix64(second ^ mix64(player + 0x9E3779B97F4A7C15 * (salt + 1))))
        legal = set(game.legal_list(state))
        weights = [0] * game.num_actions
        for a in legal:
            weights[a] = (mix64(h0 + a * 0xD6E8FEB86659FD93) >> 40) % 997 + 1
        total = sum(weights)
        value = ((h0 >> 20) % 1001) / 1000.0
        return value, [float(w) / float(total) for w in weights]

    return evaluator


hashed = hashed_salted(0)
