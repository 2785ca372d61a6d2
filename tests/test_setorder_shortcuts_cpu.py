"""The two shortcuts `warp_children_order` (deep_mcts_b200/csrc/games.cuh) takes around the CPython-set emulation,
checked on the CPU against real `set` objects and the oracle's restatement of the reference's scan.  (The kernels
themselves are checked against the reference's 100 000-position digests in tests/test_games_gpu.py.)"""
import random

from oracle.games import OTHELLO_SHIFTS, make_game
from oracle.setorder import set_order
from oracle.synth import random_positions


def test_distinct_home_slots_iterate_by_key_mod_32():
    """5..18 keys < 64 that are distinct modulo 32: set(keys) -- and a set built from its iteration order, as
    MCTS.expand_node does with Othello's legal_actions (reference mcts.py:200-215 over othello/game.py:88-118) --
    iterates in ascending key & 31 whatever the insertion order."""
    rnd = random.Random(7)
    for _ in range(20000):
        n = rnd.randint(5, 18)
        keys = [home + 32 * rnd.randint(0, 1) for home in rnd.sample(range(32), n)]
        rnd.shuffle(keys)
        want = sorted(keys, key=lambda k: k & 31)
        first = set()
        for k in keys:
            first.add(k)
        assert list(first) == want
        assert list(set(list(first))) == want
        assert set_order(keys) == want


def _dir_index(dx, dy):
    return (0 if dx == 0 else 3 if dx == 1 else 6) + (0 if dy == 0 else 1 if dy == 1 else 2) - 1


def _discovery_from_the_moves(game, state):
    """What the kernel does: from every legal move's square walk the 8 rays over opponent stones to the bracketing
    own stone; that stone's ray in the opposite direction is the one that finds the move in the reference's scan,
    8 * stone cell + ray index is when, the flipped stones are the union of the bracketed runs."""
    player, first, second = state
    own, opp = (first, second) if player == 1 else (second, first)
    g = game.g
    mask = game.legal_mask(state)
    mask = mask[0] if isinstance(mask, tuple) else mask
    found = []
    for k in range(g * g):
        if not (mask >> k) & 1:
            continue
        seq, flips = None, 0
        for dx, dy in OTHELLO_SHIFTS:
            x, y, run = k % g, k // g, 0
            while True:
                x, y = x + dx, y + dy
                if not (0 <= x < g and 0 <= y < g) or not (opp >> (y * g + x)) & 1:
                    break
                run |= 1 << (y * g + x)
            if run and 0 <= x < g and 0 <= y < g and (own >> (y * g + x)) & 1:
                flips |= run
                when = (y * g + x) * 8 + _dir_index(-dx, -dy)
                seq = when if seq is None else min(seq, when)
        assert flips == game._othello_flips(state, k)
        found.append((seq, k))
    return [k for _, k in sorted(found)]


def test_othello_discovery_order_from_the_moves_own_squares():
    assert [_dir_index(dx, dy) for dx, dy in OTHELLO_SHIFTS] == list(range(8))
    for g, n in ((4, 300), (6, 400), (8, 800)):
        game = make_game("othello", g)
        for state in random_positions(game, n, seed=100 + g):
            scan = []
            for k in game._othello_discovery(state):     # the reference's scan, duplicates kept
                if k not in scan:
                    scan.append(k)
            if scan:
                assert _discovery_from_the_moves(game, state) == scan
