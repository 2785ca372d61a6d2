"""GPU parity of the tree-search kernels: per-move root visit counts, (with, without) stats,
child order and root visits must equal the reference's (golden fixture) and the oracle's,
bit for bit, in exact mode with deterministic evaluators."""
import json

import numpy as np
import pytest

from oracle import evaluators as oev
from oracle.games import make_game
from oracle.mcts import OracleMCTS

pytestmark = pytest.mark.gpu

from test_games_gpu import manager_for  # noqa: E402


def device_callables(manager, evaluator, rollout):
    from deep_mcts_b200 import mcts as dm
    ev = {None: None, "uniform": dm.uniform_state_evaluator(manager),
          "hashed": dm.hashed_state_evaluator(manager, 0), "hashed7": dm.hashed_state_evaluator(manager, 7)}[evaluator]
    ro = {None: None, "min": dm.min_action_policy(manager), "max": dm.max_action_policy(manager)}[rollout]
    return ev, ro


def check_move(m, mv):
    dist, stats = m.step()
    visits, _ = m.pool.root_visits()
    order, count, root_visits = m.pool.root_children()
    assert [int(v) for v in visits[0]] == mv["visits"]
    assert list(stats) == mv["stats"]
    assert [int(a) for a in order[0, :count[0]]] == mv["order"]
    assert int(root_visits[0]) == mv["root_visits"]
    total = sum(mv["visits"])
    assert dist == [v / total for v in mv["visits"]]
    root = m.root
    assert list(root.children.keys()) == mv["order"] and root.visits == mv["root_visits"]


def test_visit_counts_match_reference_fixture(golden_dir):
    from deep_mcts_b200.mcts import MCTS
    data = json.load(open(golden_dir / "mcts.json"))
    for rec in data["games"]:
        manager = manager_for(rec["game"], rec["grid"])
        ev, ro = device_callables(manager, rec["evaluator"], rec["rollout"])
        m = MCTS(manager, rec["sims"], ro, ev)
        for mv in rec["moves"]:
            check_move(m, mv)
            m.advance(mv["action"])
        assert manager.is_final_state(m.state)
        assert manager.evaluate_final_state(m.state).value == rec["outcome"]


def test_host_callable_evaluator_goes_through_leaf_exchange(golden_dir):
    """An arbitrary Python state_evaluator / rollout_policy must give the same tree."""
    from deep_mcts_b200._engine import state_to_bits
    from deep_mcts_b200.mcts import MCTS
    data = json.load(open(golden_dir / "mcts.json"))
    picked = [r for r in data["games"] if (r["game"], r["grid"], r["evaluator"]) in
              {("othello", 6, "hashed"), ("hex_with_swap", 5, "uniform"), ("tictactoe", 3, "hashed")}][:4]
    picked += [r for r in data["games"] if r["rollout"] and r["game"] == "hex_with_swap"][:1]
    assert picked
    for rec in picked:
        manager = manager_for(rec["game"], rec["grid"])
        og = make_game(rec["game"], rec["grid"])
        fn = {"uniform": oev.uniform, "hashed": oev.hashed, None: None}[rec["evaluator"]]
        ev = (lambda state, fn=fn: fn(og, state_to_bits(state))) if fn else None
        ro = None
        if rec["rollout"]:
            pick = min if rec["rollout"] == "min" else max
            ro = lambda state, pick=pick: pick(manager.legal_actions(state))  # noqa: E731
        m = MCTS(manager, rec["sims"], ro, ev)
        for mv in rec["moves"]:
            check_move(m, mv)
            m.advance(mv["action"])


def test_agent_rerooting_matches_reference_fixture(golden_dir):
    from deep_mcts_b200.mcts import MCTS, MCTSAgent
    data = json.load(open(golden_dir / "mcts.json"))
    for rec in data["matches"]:
        manager = manager_for(rec["game"], rec["grid"])
        agents, logs = [], [iter(rec["log_a"]), iter(rec["log_b"])]
        for sims, evk, rok in (rec["a"], rec["b"]):
            ev, ro = device_callables(manager, evk, rok)
            agents.append(MCTSAgent(MCTS(manager, sims, ro, ev)))
        state = manager.initial_game_state()
        turn = 0
        while not manager.is_final_state(state):
            agent = agents[turn]
            agent.mcts.observe(state)
            check_move(agent.mcts, next(logs[turn]))
            # replay through the public agent API on a second copy is covered below; here advance manually
            visits, _ = agent.mcts.pool.root_visits()
            action = int(np.argmax(visits[0]))
            agent.mcts.advance(action)
            state = manager.generate_child_state(state, action)
            turn ^= 1
        assert manager.evaluate_final_state(state).value == rec["outcome"]


def test_tournament_play_outcome_matches_reference(golden_dir):
    from deep_mcts_b200.mcts import MCTS, MCTSAgent
    from deep_mcts_b200.tournament import play
    data = json.load(open(golden_dir / "mcts.json"))
    rec = data["matches"][0]
    manager = manager_for(rec["game"], rec["grid"])
    agents = []
    for sims, evk, rok in (rec["a"], rec["b"]):
        ev, ro = device_callables(manager, evk, rok)
        agents.append(MCTSAgent(MCTS(manager, sims, ro, ev)))
    outcome = play((agents[0], agents[1]), manager)
    assert outcome.value == rec["outcome"]
    assert [list(s) for s in agents[0].simulation_stats[0]] == [m["stats"] for m in rec["log_a"]]


@pytest.mark.parametrize("name,g,kind,sims", [("othello", 8, "hashed", 50), ("hex", 7, "uniform", 40),
                                               ("hex_with_swap", 6, "hashed", 30), ("othello", 6, "uniform", 60)])
def test_batched_trees_match_oracle(name, g, kind, sims):
    """64 trees searched by one kernel launch per move, each from its own forced opening."""
    from deep_mcts_b200 import mcts as dm
    game = make_game(name, g)
    manager = manager_for(name, g)
    T, moves = 64, 6
    ev = dm.uniform_state_evaluator(manager) if kind == "uniform" else dm.hashed_state_evaluator(manager, 0)
    batch = dm.BatchedMCTS(manager, T, sims, None, ev)
    refs = [OracleMCTS(game, sims, None, oev.uniform if kind == "uniform" else oev.hashed) for _ in range(T)]
    rs = np.random.RandomState(0)
    live = np.ones(T, bool)
    for ply in range(moves):
        visits, stats = batch.step()
        actions = np.full(T, -1, np.int32)
        for t, ref in enumerate(refs):
            if not live[t]:
                continue
            _, ref_stats = ref.step()
            assert [int(v) for v in visits[t]] == ref.root_visit_counts(), (t, ply)
            assert tuple(stats[t]) == tuple(ref_stats)
            legal = ref.root_children_actions()
            a = legal[rs.randint(len(legal))] if ply < 3 else int(np.argmax(visits[t]))
            ref.advance(a)
            actions[t] = a
            if game.is_final(ref.state):
                live[t] = False
        batch.advance(actions)
    c = batch.pool.counters()
    assert c["capacity_errors"] == 0 and c["simulations"] > 0


def test_value_error_when_both_callables_missing():
    from deep_mcts_b200.mcts import MCTS
    with pytest.raises(ValueError):
        MCTS(manager_for("tictactoe", 3), 10, None, None)


def test_host_noise_dirichlet_matches_oracle():
    """add_dirichlet_noise (mcts.py:233-241) with host-supplied noise, in child order."""
    from deep_mcts_b200 import mcts as dm
    game = make_game("othello", 6)
    manager = manager_for("othello", 6)
    m = dm.MCTS(manager, 40, None, dm.hashed_state_evaluator(manager), dirichlet_alpha=1.0)
    ref = OracleMCTS(game, 40, None, oev.hashed, dirichlet_alpha=1.0)
    rs = np.random.RandomState(4)
    for _ in range(5):
        n = len(game.children_order(ref.state))
        noise = rs.dirichlet([1.0] * n)
        m.pool.set_noise(noise[None, :])
        dist, stats = m.step()
        ref_dist, ref_stats = ref.step(noise=noise)
        assert dist == ref_dist and tuple(stats) == tuple(ref_stats)
        a = int(np.argmax(dist))
        m.advance(a)
        ref.advance(a)


def test_device_dirichlet_and_random_rollouts_run():
    """Stochastic device paths: sanity only (distributional parity is in test_games_gpu)."""
    from deep_mcts_b200 import mcts as dm
    manager = manager_for("hex_with_swap", 5)
    m = dm.MCTS(manager, 64, dm.random_rollout_policy(manager), None, dirichlet_alpha=0.33, seed=3)
    dist, stats = m.step()
    assert abs(sum(dist) - 1.0) < 1e-12 and sum(stats) == 64
    m2 = dm.MCTS(manager, 32, dm.random_rollout_policy(manager), dm.uniform_state_evaluator(manager),
                 rollout_share=0.5)
    dist, stats = m2.step()
    assert sum(stats) == 32


@pytest.mark.parametrize("name,g,W", [("othello", 8, 4), ("hex", 7, 3), ("hex_with_swap", 5, 2)])
def test_virtual_loss_mode_invariants_and_quality(name, g, W):
    """Throughput mode (several leaves per tree per round, virtual loss): not bit-exact by design,
    but every simulation must be accounted for and the visit distribution must stay close to the
    exact search (the synthetic evaluator is pure noise, so only a loose similarity is meaningful)."""
    from deep_mcts_b200._engine import state_to_bits
    from deep_mcts_b200 import mcts as dm
    game = make_game(name, g)
    manager = manager_for(name, g)
    sims, T = 96, 48
    ev = lambda state: oev.hashed(game, state_to_bits(state))  # noqa: E731  host evaluator -> leaf exchange
    exact = dm.BatchedMCTS(manager, T, sims, None, ev)
    fast = dm.BatchedMCTS(manager, T, sims, None, ev, leaves_per_round=W)
    # distinct openings per tree
    rs = np.random.RandomState(1)
    p, f, s = manager.initial_batch(T)
    for _ in range(3):
        lst, _, cnt = manager.legal_batch(p, f, s)
        acts = np.array([lst[i, rs.randint(cnt[i])] for i in range(T)], np.int32)
        p, f, s, ok = manager.child_batch(p, f, s, acts)
        assert ok.all()
    exact.pool.set_root_states(p, f, s)
    fast.pool.set_root_states(p, f, s)
    agree, l1 = [], []
    for ply in range(3):
        v_exact, s_exact = exact.step()
        v_fast, s_fast = fast.step()
        assert (s_fast.sum(axis=1) == sims).all() and (s_exact.sum(axis=1) == sims).all()
        assert (v_fast.sum(axis=1) >= sims - 1).all()
        pe = v_exact / v_exact.sum(axis=1, keepdims=True)
        pf = v_fast / v_fast.sum(axis=1, keepdims=True)
        agree.append((pe.argmax(axis=1) == pf.argmax(axis=1)).mean())
        l1.append(np.abs(pe - pf).sum(axis=1).mean())
        actions = np.array([int(np.argmax(v_exact[t])) for t in range(T)], np.int32)
        exact.advance(actions)
        fast.advance(actions)
    assert np.mean(agree) > 0.3, (agree, l1)
    assert np.mean(l1) < 1.3, (agree, l1)
    assert fast.pool.counters()["capacity_errors"] == 0


def test_virtual_loss_finds_forced_win():
    """Tic-tac-toe, FIRST to move with two in a row: both modes must put most visits on the winning cell."""
    from deep_mcts_b200 import mcts as dm
    from deep_mcts_b200._engine import state_to_bits
    game = make_game("tictactoe", 3)
    manager = manager_for("tictactoe", 3)
    # X X .      FIRST = X to move, wins at cell 2
    # O O .
    # . . .
    state = (1, 0b000000011, 0b000011000)
    ev = lambda st: oev.uniform(game, state_to_bits(st))  # noqa: E731
    for W in (1, 4):
        m = dm.MCTS(manager, 200, None, ev, leaves_per_round=W)
        m.pool.set_root_states([state[0]], [state[1]], [state[2]])
        m.state = manager._make_state(*state)
        dist, stats = m.step()
        assert int(np.argmax(dist)) == 2 and dist[2] > 0.5, (W, dist)


def test_virtual_loss_values_are_restored():
    """After a virtual-loss step no node may keep a virtual loss: root children means stay in [0, 1]
    and the visit total equals the simulations run (checked through a second exact step on the same tree)."""
    from deep_mcts_b200 import mcts as dm
    from deep_mcts_b200._engine import state_to_bits
    game = make_game("othello", 6)
    manager = manager_for("othello", 6)
    ev = lambda state: oev.hashed(game, state_to_bits(state))  # noqa: E731
    m = dm.MCTS(manager, 64, None, ev, leaves_per_round=4)
    dist, stats = m.step()
    assert sum(stats) == 64 and abs(sum(dist) - 1) < 1e-12
    root = m.root
    assert sum(c.visits for c in root.children.values()) == 64
    assert root.visits == 65


def test_full_size_pool_consistency_and_oracle_subset():
    """BASELINE size (4096 Othello 8x8 trees, 50 sims): 8 distinct openings x 512 copies.  Copies of
    one opening run in different warps / blocks / SMs and must produce identical results (checksum
    of checksums); one tree per opening is compared with the oracle bit for bit; totals must add up."""
    from deep_mcts_b200 import mcts as dm
    game = make_game("othello", 8)
    manager = manager_for("othello", 8)
    T, K, sims = 4096, 8, 50
    rs = np.random.RandomState(7)
    openings = []
    for _ in range(K):
        s = game.initial_state()
        for _ in range(4):
            legal = game.legal_list(s)
            s = game.child(s, legal[rs.randint(len(legal))])
        openings.append(s)
    idx = np.arange(T) % K
    p = np.array([openings[i][0] for i in idx], np.uint8)
    f = np.array([openings[i][1] for i in idx], np.uint64)
    s2 = np.array([openings[i][2] for i in idx], np.uint64)
    batch = dm.BatchedMCTS(manager, T, sims, None, dm.hashed_state_evaluator(manager, 3))
    batch.pool.set_root_states(p, f, s2)
    refs = []
    for o in openings:
        r = OracleMCTS(game, sims, None, oev.hashed_salted(3))
        r.set_root_state(o)
        refs.append(r)
    c_before = batch.pool.counters()
    for move in range(4):
        visits, stats = batch.step()
        assert (stats.sum(axis=1) == sims).all()
        for k in range(K):
            group = visits[idx == k]
            assert (group == group[0]).all(), f"copies of opening {k} diverged at move {move}"
            _, ref_stats = refs[k].step()
            assert [int(v) for v in group[0]] == refs[k].root_visit_counts()
            assert tuple(stats[k]) == tuple(ref_stats)
        actions = np.array([int(np.argmax(visits[t])) for t in range(T)], np.int32)
        for k in range(K):
            refs[k].advance(int(actions[k]))
        batch.advance(actions)
    c = batch.pool.counters()
    assert c["simulations"] - c_before["simulations"] == T * sims * 4 and c["capacity_errors"] == 0


def test_capacity_error_is_reported_not_silent():
    from deep_mcts_b200 import _lib
    from deep_mcts_b200 import mcts as dm
    manager = manager_for("othello", 8)
    m = dm.MCTS(manager, 200, None, dm.uniform_state_evaluator(manager), max_nodes_per_tree=64)
    with pytest.raises(_lib.DeepMctsError) as err:
        m.step()
    assert err.value.code == _lib.DM_ERR_CAPACITY


def test_time_budget_mode_and_illegal_action():
    from deep_mcts_b200 import mcts as dm
    manager = manager_for("hex", 5)
    m = dm.MCTS(manager, float("inf"), dm.random_rollout_policy(manager), None, time_per_move=0.2)
    dist, (with_exp, without_exp) = m.step()
    assert with_exp + without_exp >= 64 and abs(sum(dist) - 1) < 1e-9
    with pytest.raises(AssertionError):
        m.advance(0) if False else manager.generate_child_state(
            manager.generate_child_state(manager.initial_game_state(), 0), 0)
    # advancing the pool by an illegal action is counted, not applied
    before = m.pool.counters()["invalid_actions"]
    m.pool.advance([manager.num_actions + 5])
    assert m.pool.counters()["invalid_actions"] == before + 1


def test_odd_tree_counts_and_reset_subset():
    from deep_mcts_b200 import mcts as dm
    manager = manager_for("tictactoe", 3)
    b = dm.BatchedMCTS(manager, 7, 16, None, dm.hashed_state_evaluator(manager, 1))
    v1, _ = b.step()
    b.advance(np.argmax(v1, axis=1).astype(np.int32))
    b.reset([0, 3])
    p, f, s, fin, _ = b.pool.states()
    assert f[0] == 0 and s[0] == 0 and f[3] == 0 and (f[1] | s[1]) != 0
    v2, st = b.step()
    assert (st.sum(axis=1) == 16).all()
    assert (v2[0] == v1[0]).all()          # a reset tree repeats its first search exactly


@pytest.mark.parametrize("name,g,kind", [("othello", 8, "hashed"), ("hex", 7, "uniform")])
def test_arena_compaction_preserves_exactness(name, g, kind):
    """A whole game in an arena far smaller than the game's total node count: the in-place compaction
    at re-roots must keep visit counts, child order and root visits bit-identical to the oracle."""
    from deep_mcts_b200 import mcts as dm
    game = make_game(name, g)
    manager = manager_for(name, g)
    sims = 40
    ev = dm.hashed_state_evaluator(manager, 0) if kind == "hashed" else dm.uniform_state_evaluator(manager)
    # room for the reused subtree plus one move of growth, still ~10x less than the whole game creates
    # (compaction runs when the next move's worst case, (sims + 2) * actions nodes, might not fit any more)
    small = dm.MCTS(manager, sims, None, ev, max_nodes_per_tree=(sims + 2) * game.num_actions * 4)
    ref = OracleMCTS(game, sims, None, oev.hashed if kind == "hashed" else oev.uniform)
    moves = 0
    while not game.is_final(ref.state) and moves < 40:
        dist, stats = small.step()
        ref_dist, ref_stats = ref.step()
        assert dist == ref_dist and tuple(stats) == tuple(ref_stats), moves
        order, count, root_visits = small.pool.root_children()
        assert [int(a) for a in order[0, :count[0]]] == ref.root_children_actions()
        assert int(root_visits[0]) == ref.visits[ref.root]
        a = int(np.argmax(ref_dist))
        small.advance(a)
        ref.advance(a)
        moves += 1
    c = small.pool.counters()
    assert c["compactions"] >= (2 if name == "hex" else 1) and c["capacity_errors"] == 0, c
    assert c["children_created"] > small.pool._search.num_simulations  # far more nodes created than the arena holds


@pytest.mark.parametrize("name,g", [("othello", 8), ("hex", 7), ("hex_with_swap", 7), ("othello", 6)])
def test_fused_round_stepped_search_matches_oracle(name, g):
    """MCTS.step driven through dm_pool_round (expand/backup of the previous batch + next selection in one
    launch; the BASELINE board sizes run the per-game kernel instances) with a host evaluator: visit counts,
    stats and child order must equal the oracle's, move after move with subtree reuse."""
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import TreePool
    from test_selfplay_gpu import evaluate_leaves
    game = make_game(name, g)
    manager = manager_for(name, g)
    sims, T = 40, 5
    pool = TreePool(manager, T, 1 << 15, seed=2)
    pool.configure(sims, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, 0, 0.0, 0.25)
    # distinct openings
    rs = np.random.RandomState(3)
    openings = []
    for t in range(T):
        s = game.initial_state()
        for _ in range(t):
            legal = game.legal_list(s)
            s = game.child(s, legal[rs.randint(len(legal))])
        openings.append(s)
    pool.set_root_states(np.array([o[0] for o in openings], np.uint8), np.array([o[1] for o in openings], np.uint64),
                         np.array([o[2] for o in openings], np.uint64))
    refs = []
    for o in openings:
        r = OracleMCTS(game, sims, None, oev.hashed)
        r.set_root_state(o)
        refs.append(r)
    for move in range(3):
        pool.begin_step()
        pool.select()
        pri, val = evaluate_leaves(pool, game, oev.hashed, int(pool.leaf_count.item()))
        for _ in range(sims):
            pool.round(pri, val, selfplay=False)
            pri, val = evaluate_leaves(pool, game, oev.hashed, int(pool.leaf_count.item()))
        pool.expand_backup(pri, val)
        visits, stats = pool.root_visits()
        order, count, root_visits = pool.root_children()
        actions = np.zeros(T, np.int32)
        for t, ref in enumerate(refs):
            _, ref_stats = ref.step()
            assert [int(v) for v in visits[t]][: game.num_actions] == ref.root_visit_counts(), (t, move)
            assert tuple(stats[t]) == tuple(ref_stats)
            assert [int(a) for a in order[t, : count[t]]] == ref.root_children_actions()
            actions[t] = int(np.argmax(visits[t]))
            ref.advance(int(actions[t]))
        pool.advance(actions)
    assert pool.counters()["capacity_errors"] == 0


@pytest.mark.parametrize("entry_index", range(6))
def test_full_games_match_reference_mcts_digests(golden_dir, entry_index):
    """Several hundred full games of the unmodified reference MCTS (tests/golden/make_mcts_digests.py: hashed
    evaluator, 25 / 50 / 200 simulations per move, subtree reuse, forced random openings), one SHA-256 per
    game over every move's visit counts, stats, child order, root visits and action.  All games of a plan
    entry are replayed at once, one tree per game; every digest must match."""
    import hashlib
    import random
    import struct
    import sys
    from deep_mcts_b200 import mcts as dm
    sys.path.insert(0, str(golden_dir))
    from make_mcts_digests import absorb_move, choose, game_seed
    pins = json.load(open(golden_dir / "mcts_digests.json"))
    entry = pins["plan"][entry_index]
    name, g, sims, T = entry["game"], entry["grid"], entry["sims"], entry["games"]
    manager = manager_for(name, g)
    A = manager.num_actions
    batch = dm.BatchedMCTS(manager, T, sims, None, dm.hashed_state_evaluator(manager, 0))
    rngs = [random.Random(game_seed(entry_index, k)) for k in range(T)]
    hashes = [hashlib.sha256() for _ in range(T)]
    live = np.ones(T, bool)
    ply = moves = 0
    while live.any():
        ids = np.where(live)[0].astype(np.int32)
        visits, stats = batch.step(tree_ids=ids)
        order, count, root_visits = batch.pool.root_children()
        actions = np.full(T, -1, np.int32)
        for t in ids:
            v = [int(x) for x in visits[t][:A]]
            o = [int(a) for a in order[t, : count[t]]]
            a = choose(rngs[t], ply, t % (pins["max_opening_plies"] + 1), o, v)
            absorb_move(hashes[t], v, (int(stats[t][0]), int(stats[t][1])), o, int(root_visits[t]), a)
            actions[t] = a
            moves += 1
        batch.advance(actions)
        _, _, _, fin, outcome = batch.pool.states()
        for t in ids:
            if fin[t]:
                hashes[t].update(struct.pack("<f", float(outcome[t])))
                live[t] = False
        ply += 1
    assert moves == entry["moves"]
    got = [h.hexdigest() for h in hashes]
    wrong = [k for k in range(T) if got[k] != entry["digests"][k]]
    assert not wrong, f"{len(wrong)} of {T} games differ from the reference, first: game {wrong[0]}"
    assert batch.pool.counters()["capacity_errors"] == 0


def test_python_rollout_policy_next_to_a_network_evaluator_is_played():
    """evaluate_leaf with BOTH a state evaluator and a rollout policy returns (rollout + value) / 2 (mcts.py:187-196).
    A Python rollout policy next to a GameNet evaluator must go through the leaf exchange (network batch on the
    device, playout on the host) -- not silently drop the rollouts.  Deterministic policy => same visits as the oracle."""
    import numpy as np
    from deep_mcts_b200.gamenet import cached_state_evaluator
    from deep_mcts_b200.mcts import MCTS
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from oracle import net as onet
    from oracle.synth import synthetic_state_dict
    game = make_game("othello", 6)
    sd = synthetic_state_dict(game, 32, 1, 32, seed=5, conv_gain=1.0, fc_scale=1.0)
    net = ConvolutionalOthelloNet(6, num_residual=1, channels=32, value_head_hidden_units=32)
    net.load_state_dict(sd)
    net.precision = "fp32"
    manager = net.manager
    first_legal = lambda state: min(manager.legal_actions(state))       # noqa: E731  (a plain Python callable)
    m = MCTS(manager, 20, first_legal, cached_state_evaluator(net))
    oracle_eval = onet.make_evaluator(sd)
    ref = OracleMCTS(game, 20, lambda g, s: min(g.legal_list(s)), oracle_eval)
    for _ in range(2):
        dist, stats = m.step()
        ref_dist, ref_stats = ref.step()
        assert tuple(stats) == tuple(ref_stats)
        visits, _ = m.pool.root_visits()
        assert [int(v) for v in visits[0][: game.num_actions]] == ref.root_visit_counts()
        a = int(np.argmax(dist))
        m.advance(a)
        ref.advance(a)
    assert m.pool.counters()["rollout_plies"] == 0        # the playouts ran on the host, not in the device kernels


def test_virtual_loss_mode_rejects_rollout_policies():
    from deep_mcts_b200 import mcts as dm
    manager = manager_for("othello", 6)
    b = dm.BatchedMCTS(manager, 4, 16, dm.random_rollout_policy(manager), lambda s: (0.5, [1 / 37] * 37), leaves_per_round=4)
    with pytest.raises(ValueError):
        b.step()


def test_capacity_error_is_reported_once_and_the_pool_stays_usable():
    """A tree that outgrows its arena raises once (DM_ERR_CAPACITY); later counter reads succeed again, and a
    self-play tree whose root could not be expanded restarts instead of re-rooting on garbage."""
    from deep_mcts_b200 import _lib
    from deep_mcts_b200._lib import DeepMctsError
    from deep_mcts_b200.pool import TreePool
    manager = manager_for("othello", 8)
    pool = TreePool(manager, 8, 24, compaction=False)          # 24 node slots: a few expansions at most
    pool.configure(30, _lib.DM_EVAL_UNIFORM, _lib.DM_ROLLOUT_NONE, 0, 0.0, 0.25)
    pool.begin_step()
    pool.search_local()
    with pytest.raises(DeepMctsError):
        pool.counters()
    c = pool.counters()                                         # not sticky
    assert c["capacity_errors"] > 0
    pool.reset()
    pool.begin_step()
    assert pool.counters()["capacity_errors"] == c["capacity_errors"]
