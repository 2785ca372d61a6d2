"""Pins the CPU oracle (oracle/) against fixtures recorded from the unmodified
reference (tests/golden/make_golden.py).  Runs without a GPU."""
import json
import random

import numpy as np
import pytest
import torch

from oracle import evaluators as oev
from oracle import net as onet
from oracle.games import make_game
from oracle.mcts import OracleMCTS
from oracle.setorder import IntSet, set_order
from oracle.synth import synthetic_state_dict

GAME_CFGS = [("tictactoe", 3), ("hex", 4), ("hex", 7), ("hex_with_swap", 5), ("hex_with_swap", 7),
             ("othello", 4), ("othello", 6), ("othello", 8)]


def test_set_order_matches_interpreter():
    rng = random.Random(0)
    for _ in range(20000):
        span = rng.choice([9, 37, 50, 65, 128])
        keys = [rng.randrange(span) for _ in range(rng.randint(1, min(span, 65)))]
        assert set_order(keys) == list(set(keys))
    for _ in range(5000):
        s, m = set(), IntSet()
        for _ in range(rng.randint(1, 40)):
            k = rng.randrange(65)
            s.add(k)
            m.add(k)
        assert m.order() == list(s)


@pytest.mark.parametrize("name,g", GAME_CFGS)
def test_game_logic_matches_reference(golden_dir, name, g):
    z = np.load(golden_dir / f"games_{name}_{g}.npz")
    game = make_game(name, g)
    off = z["offsets"]
    for i in range(len(z["player"])):
        state = (int(z["player"][i]), int(z["first"][i]), int(z["second"][i]))
        lo, hi = int(off[i]), int(off[i + 1])
        legal = [int(a) for a in z["legal"][lo:hi]]
        assert game.legal_list(state) == legal
        assert game.children_order(state) == [int(a) for a in z["order"][lo:hi]]
        assert game.is_final(state) == bool(z["is_final"][i])
        assert game.outcome(state) == float(z["outcome"][i])
        for j, a in enumerate(legal):
            want = (int(z["child_player"][lo + j]), int(z["child_first"][lo + j]), int(z["child_second"][lo + j]))
            assert game.child(state, a) == want
    assert game.initial_state() == (int(z["player"][0]), int(z["first"][0]), int(z["second"][0]))


def _evaluator(kind):
    return {None: None, "uniform": oev.uniform, "hashed": oev.hashed, "hashed7": oev.hashed_salted(7)}[kind]


def _rollout(kind):
    if kind is None:
        return None
    pick = min if kind == "min" else max
    return lambda game, state: pick(game.legal_list(state))


def test_mcts_visit_counts_match_reference(golden_dir):
    data = json.load(open(golden_dir / "mcts.json"))
    for rec in data["games"]:
        game = make_game(rec["game"], rec["grid"])
        m = OracleMCTS(game, rec["sims"], _rollout(rec["rollout"]), _evaluator(rec["evaluator"]))
        for mv in rec["moves"]:
            assert not game.is_final(m.state)
            _, stats = m.step()
            assert m.root_visit_counts() == mv["visits"]
            assert list(stats) == mv["stats"]
            assert m.root_children_actions() == mv["order"]
            assert m.visits[m.root] == mv["root_visits"]
            m.advance(mv["action"])
        assert game.is_final(m.state)
        assert game.outcome(m.state) == rec["outcome"]


def test_agent_rerooting_matches_reference(golden_dir):
    data = json.load(open(golden_dir / "mcts.json"))
    for rec in data["matches"]:
        game = make_game(rec["game"], rec["grid"])
        agents = [OracleMCTS(game, s, _rollout(ro), _evaluator(ev)) for s, ev, ro in (rec["a"], rec["b"])]
        logs = [iter(rec["log_a"]), iter(rec["log_b"])]
        state = game.initial_state()
        turn = 0
        while not game.is_final(state):
            m = agents[turn]
            m.observe(state)  # MCTSAgent.play re-rooting, mcts.py:267-283
            dist, stats = m.step()
            want = next(logs[turn])
            assert m.root_visit_counts() == want["visits"]
            assert list(stats) == want["stats"]
            assert m.root_children_actions() == want["order"]
            action = int(np.argmax(dist))
            m.advance(action)
            state = game.child(state, action)
            turn ^= 1
        assert game.outcome(state) == rec["outcome"]


NET_CFGS = [("tictactoe", 3, 3), ("hex", 7, 128), ("hex_with_swap", 7, 128), ("hex_with_swap", 5, 32),
            ("othello", 8, 128), ("othello", 6, 128)]


@pytest.mark.parametrize("name,g,c", NET_CFGS)
def test_net_matches_reference(golden_dir, name, g, c):
    z = np.load(golden_dir / f"net_{name}_{g}_c{c}.npz")
    channels, num_residual, hidden, seed = (int(v) for v in z["arch"])
    game = make_game(name, g)
    sd = synthetic_state_dict(game, channels, num_residual, hidden, seed)
    states = [(int(p), int(f), int(s)) for p, f, s in zip(z["player"], z["first"], z["second"])]
    planes = onet.states_to_planes(game, states)
    assert np.array_equal(planes.numpy(), z["planes"])
    v, logits = onet.forward(game, sd, planes)
    np.testing.assert_allclose(v.numpy(), z["raw_value"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(logits.reshape(len(states), -1).numpy(), z["raw_logits"], atol=2e-5, rtol=0)
    value, probs = onet.evaluate_batch(game, sd, states)
    # north_star tolerance for fp32 network outputs: 1e-5 absolute
    np.testing.assert_allclose(value, z["value"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(probs, z["probs"], atol=1e-5, rtol=0)


@pytest.mark.parametrize("name,g", [("tictactoe", 3), ("hex", 7), ("hex_with_swap", 7), ("othello", 8)])
def test_oracle_game_logic_matches_reference_digest_over_100k_positions(golden_dir, name, g):
    """SHA-256 over 100 000 positions of seeded random playouts (state, is_final, outcome, legal list order,
    set order, every child state) recorded from the unmodified reference managers
    (tests/golden/make_digests.py): the oracle, driven through the same walk, must reproduce the digest at
    every checkpoint."""
    import sys
    sys.path.insert(0, str(golden_dir))
    from make_digests import walk_digest
    pins = json.load(open(golden_dir / "digests.json"))
    game = make_game(name, g)
    got = walk_digest(game.initial_state, game.legal_list, game.children_order, game.child,
                      lambda s: (game.is_final(s), game.outcome(s)), max(pins["checkpoints"]), tuple(pins["checkpoints"]),
                      seed=pins["seed"])
    assert {str(k): v for k, v in got.items()} == pins["digests"][f"{name}_{g}"]


@pytest.mark.parametrize("entry_index,n_games", [(0, 16), (1, 6), (2, 6), (3, 5), (4, 1), (5, 2)])
def test_oracle_mcts_matches_reference_game_digests(golden_dir, entry_index, n_games):
    """The first games of every entry of tests/golden/mcts_digests.json (full games of the unmodified reference
    MCTS, one SHA-256 per game) replayed by the oracle search.  The GPU suite replays all 512 games."""
    import hashlib
    import struct
    import sys
    sys.path.insert(0, str(golden_dir))
    from make_mcts_digests import absorb_move, choose, game_seed
    pins = json.load(open(golden_dir / "mcts_digests.json"))
    entry = pins["plan"][entry_index]
    game = make_game(entry["game"], entry["grid"])
    for k in range(n_games):
        rng = random.Random(game_seed(entry_index, k))
        ref = OracleMCTS(game, entry["sims"], None, oev.hashed)
        h = hashlib.sha256()
        ply = 0
        while not game.is_final(ref.state):
            _, stats = ref.step()
            visits, order = ref.root_visit_counts(), ref.root_children_actions()
            action = choose(rng, ply, k % (pins["max_opening_plies"] + 1), order, visits)
            absorb_move(h, visits, stats, order, ref.visits[ref.root], action)
            ref.advance(action)
            ply += 1
        h.update(struct.pack("<f", game.outcome(ref.state)))
        assert h.hexdigest() == entry["digests"][k], (entry["game"], entry["sims"], k)
