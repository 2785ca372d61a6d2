"""GPU parity of the game-engine kernels (through the C ABI) against the golden fixtures
recorded from the reference and against the CPU oracle on larger seeded samples."""
import numpy as np
import pytest

from oracle.games import make_game
from oracle.synth import random_positions

pytestmark = pytest.mark.gpu

CFGS = [("tictactoe", 3), ("hex", 4), ("hex", 7), ("hex_with_swap", 5), ("hex_with_swap", 7),
        ("othello", 4), ("othello", 6), ("othello", 8)]


def manager_for(name, g):
    from deep_mcts_b200.hex.game import HexManager
    from deep_mcts_b200.hex_with_swap.game import HexWithSwapManager
    from deep_mcts_b200.othello.game import OthelloManager
    from deep_mcts_b200.tictactoe.game import TicTacToeManager

    return {"tictactoe": lambda: TicTacToeManager(), "hex": lambda: HexManager(g),
            "hex_with_swap": lambda: HexWithSwapManager(g), "othello": lambda: OthelloManager(g)}[name]()


@pytest.mark.parametrize("name,g", CFGS)
def test_engine_matches_reference_fixture(golden_dir, name, g):
    z = np.load(golden_dir / f"games_{name}_{g}.npz")
    m = manager_for(name, g)
    player, first, second = z["player"], z["first"], z["second"]
    lst, order, count = m.legal_batch(player, first, second)
    fin, out = m.status_batch(player, first, second)
    off = z["offsets"]
    assert np.array_equal(count, np.diff(off))
    assert np.array_equal(fin, z["is_final"].astype(bool))
    assert np.array_equal(out, z["outcome"])
    for i in range(len(player)):
        n = count[i]
        assert list(lst[i, :n]) == list(z["legal"][off[i]:off[i + 1]])
        assert list(order[i, :n]) == list(z["order"][off[i]:off[i + 1]])
        assert np.all(lst[i, n:] == 255) and np.all(order[i, n:] == 255)
    # every edge
    rep = np.repeat(np.arange(len(player)), count)
    cp, cf, cs, ok = m.child_batch(player[rep], first[rep], second[rep], z["legal"].astype(np.int32))
    assert ok.all()
    assert np.array_equal(cp, z["child_player"])
    assert np.array_equal(cf, z["child_first"])
    assert np.array_equal(cs, z["child_second"])
    ip, if_, is_ = m.initial_batch(3)
    assert (ip[0], if_[0], is_[0]) == (player[0], first[0], second[0])


@pytest.mark.parametrize("name,g", [("hex", 7), ("hex_with_swap", 6), ("othello", 8), ("othello", 6), ("hex", 8)])
def test_engine_matches_oracle_on_seeded_positions(name, g):
    game = make_game(name, g)
    states = random_positions(game, 1500, seed=11, include_final=True)
    player = np.array([s[0] for s in states], np.uint8)
    first = np.array([s[1] for s in states], np.uint64)
    second = np.array([s[2] for s in states], np.uint64)
    m = manager_for(name, g)
    lst, order, count = m.legal_batch(player, first, second)
    fin, out = m.status_batch(player, first, second)
    for i, s in enumerate(states):
        want = game.legal_list(s)
        assert list(lst[i, :count[i]]) == want
        assert list(order[i, :count[i]]) == game.children_order(s)
        assert bool(fin[i]) == game.is_final(s)
        assert float(out[i]) == game.outcome(s)
    # one random legal child per state + illegal-action detection
    rs = np.random.RandomState(3)
    has_move = count > 0  # a completely filled hex board has no legal action
    acts = np.array([lst[i, rs.randint(count[i])] if has_move[i] else 0 for i in range(len(states))], np.int32)
    cp, cf, cs, ok = m.child_batch(player, first, second, acts)
    assert np.array_equal(ok, has_move)
    for i, s in enumerate(states):
        if has_move[i]:
            assert (int(cp[i]), int(cf[i]), int(cs[i])) == game.child(s, int(acts[i]))
    bad = np.full(len(states), game.num_actions + 3, np.int32)
    assert not m.child_batch(player, first, second, bad)[3].any()


def test_single_state_manager_api():
    from deep_mcts_b200.game import Outcome, Player
    m = manager_for("othello", 6)
    s = m.initial_game_state()
    assert s.player == Player.FIRST and len(s.grid) == 6
    legal = m.legal_actions(s)
    assert sorted(legal) == [8, 13, 22, 27]
    child = m.generate_child_state(s, legal[0])
    assert child.player == Player.SECOND and not m.is_final_state(child)
    assert m.evaluate_final_state(s) == Outcome.DRAW
    with pytest.raises(AssertionError):
        m.generate_child_state(s, 0)
    assert m.action_str(m.pass_move) == "pass" and m.action_str(8) == "(2, 1)"
    assert "pass:" in m.probabilities_grid([0.0] * 37)
    assert isinstance(str(s), str) and m.copy().grid_size == 6
    # empty batch is fine
    lst, order, count = m.legal_batch(np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    assert count.shape == (0,)


@pytest.mark.parametrize("name,g,kind", [("othello", 6, "min"), ("othello", 8, "max"), ("hex", 5, "min"),
                                          ("hex_with_swap", 5, "max"), ("tictactoe", 3, "min")])
def test_deterministic_rollouts_match_oracle(name, g, kind):
    from deep_mcts_b200 import _lib
    game = make_game(name, g)
    states = random_positions(game, 300, seed=5)
    m = manager_for(name, g)
    code = _lib.DM_ROLLOUT_MIN_ACTION if kind == "min" else _lib.DM_ROLLOUT_MAX_ACTION
    out, plies = m.rollout_batch([s[0] for s in states], [s[1] for s in states], [s[2] for s in states], code)
    pick = min if kind == "min" else max
    for i, s in enumerate(states):
        n = 0
        while not game.is_final(s):
            s = game.child(s, pick(game.legal_list(s)))
            n += 1
        assert float(out[i]) == game.outcome(s) and int(plies[i]) == n


@pytest.mark.parametrize("name,g", [("othello", 6), ("hex", 5), ("hex_with_swap", 4), ("tictactoe", 3)])
def test_random_rollout_rates_within_binomial_bounds(name, g):
    """W/D/L of uniform random playouts from fixed positions: CUDA vs oracle, 5-sigma bound."""
    import random
    game = make_game(name, g)
    starts = [game.initial_state()] + random_positions(game, 2, seed=9)
    if name == "hex_with_swap":  # a state where swap is legal
        starts.append(game.child(game.initial_state(), 5))
    m = manager_for(name, g)
    n_gpu, n_cpu = 40000, 1500
    rng = random.Random(1)
    for s0 in starts:
        out, _ = m.rollout_batch([s0[0]] * n_gpu, [s0[1]] * n_gpu, [s0[2]] * n_gpu, seed=123)
        cpu = []
        for _ in range(n_cpu):
            s = s0
            while not game.is_final(s):
                s = game.child(s, rng.choice(game.legal_list(s)))
            cpu.append(game.outcome(s))
        cpu = np.array(cpu)
        for v in (0.0, 0.5, 1.0):
            p_gpu, p_cpu = float(np.mean(out == v)), float(np.mean(cpu == v))
            sigma = np.sqrt(max(p_gpu * (1 - p_gpu), 1e-4) * (1 / n_gpu + 1 / n_cpu))
            assert abs(p_gpu - p_cpu) < 5 * sigma + 1e-3, (name, s0, v, p_gpu, p_cpu)


@pytest.mark.parametrize("name,g", [("tictactoe", 3), ("hex", 7), ("hex_with_swap", 7), ("othello", 8)])
def test_device_game_logic_matches_reference_digest_over_100k_positions(golden_dir, name, g):
    """The CUDA game engines against the reference's 100 000-position digests (tests/golden/digests.json):
    the positions of the seeded walk are evaluated on the device in batches -- legal list order, set
    (child) order, is_final, outcome and every child state -- and hashed in walk order."""
    import hashlib
    import json
    import random
    import sys
    sys.path.insert(0, str(golden_dir))
    from make_digests import absorb
    pins = json.load(open(golden_dir / "digests.json"))
    n_positions, checkpoints = max(pins["checkpoints"]), set(pins["checkpoints"])
    game = make_game(name, g)
    manager = manager_for(name, g)
    # the walk (which move is played) needs only the sorted legal set; take it from the oracle
    rng = random.Random(pins["seed"])
    states = []
    while len(states) < n_positions:
        s = game.initial_state()
        while len(states) < n_positions:
            states.append(s)
            if game.is_final(s):
                break
            choices = sorted(set(game.legal_list(s)))
            s = game.child(s, choices[rng.randrange(len(choices))])
    p = np.array([s[0] for s in states], np.uint8)
    f = np.array([s[1] for s in states], np.uint64)
    s2 = np.array([s[2] for s in states], np.uint64)
    lst, order, cnt = manager.legal_batch(p, f, s2)
    final, outcome = manager.status_batch(p, f, s2)
    rep = np.repeat(np.arange(n_positions), cnt)
    acts = np.concatenate([lst[i, : cnt[i]] for i in range(n_positions)]).astype(np.int32)
    cp, cf, cs, ok = manager.child_batch(p[rep], f[rep], s2[rep], acts)
    assert ok.all()
    h = hashlib.sha256()
    got = {}
    edge = 0
    for i in range(n_positions):
        k = int(cnt[i])
        children = [(int(cp[edge + j]), int(cf[edge + j]), int(cs[edge + j])) for j in range(k)]
        edge += k
        absorb(h, states[i], bool(final[i]), float(outcome[i]), [int(a) for a in lst[i, :k]],
               [int(a) for a in order[i, :k]], children)
        if i + 1 in checkpoints:
            got[str(i + 1)] = h.hexdigest()
    assert got == pins["digests"][f"{name}_{g}"]
