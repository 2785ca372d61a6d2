"""GPU parity of the device evaluation cache (reference train.py:408-414 ``cached_state_evaluator``): a
result-transparent memo of the evaluator plus de-duplication of equal leaves inside a round.  Visit counts
must be bit-identical with the cache on and off -- against the reference's own digests for the stepped
search, and between two runs of the continuous self-play kernels."""
import hashlib
import json
import random
import struct
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_games_gpu import manager_for  # noqa: E402


class HashedDeviceEvaluator:
    """dm_eval_hashed: the DM_EVAL_HASHED test evaluator as an external one with double outputs."""

    def __init__(self, pool, salt=0):
        from deep_mcts_b200 import _lib
        self.pool, self.salt = pool, salt
        T = pool.n_trees
        self.value = torch.zeros(T * _lib.DM_MAX_LEAVES, dtype=torch.float64, device=pool.device)
        self.probs = torch.zeros((T * _lib.DM_MAX_LEAVES, _lib.DM_MAX_ACTIONS), dtype=torch.float64, device=pool.device)

    def __call__(self, dense=False):
        from deep_mcts_b200 import _lib
        pool = self.pool
        m = pool.manager
        _lib.call("dm_eval_hashed", m.game_id, m.board, pool.n_trees, pool.leaf_count.data_ptr(),
                  pool.leaf_index.data_ptr() if dense else 0, pool.leaf_first.data_ptr(), pool.leaf_second.data_ptr(),
                  pool.leaf_player.data_ptr(), self.salt, self.value.data_ptr(), self.probs.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
        return self.probs, self.value


def stepped_search(pool, ev, sims, ids):
    """One MCTS.step on the listed trees through the leaf-exchange kernels (select, then fused rounds)."""
    pool.begin_step(ids)
    pool.select()
    pri, val = ev()
    for _ in range(sims):
        pool.round(pri, val, selfplay=False)
        pri, val = ev()
    pool.expand_backup(pri, val)


@pytest.mark.parametrize("cache_log2", [None, 18, 10])
@pytest.mark.parametrize("entry_index", range(6))
def test_cached_search_matches_reference_mcts_digests(golden_dir, entry_index, cache_log2):
    """The 512 full games of the unmodified reference MCTS (tests/golden/mcts_digests.json) replayed through the
    external-evaluator kernels with the hashed evaluator computed by a separate kernel in double precision:
    without a cache, with a roomy one, and with a 1024-entry one that evicts constantly.  Every digest must
    match, so hits, in-round aliases, evictions and lost claims are all invisible in the search result."""
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import TreePool, default_node_capacity
    sys.path.insert(0, str(golden_dir))
    from make_mcts_digests import absorb_move, choose, game_seed
    pins = json.load(open(golden_dir / "mcts_digests.json"))
    entry = pins["plan"][entry_index]
    name, g, sims, T = entry["game"], entry["grid"], entry["sims"], entry["games"]
    manager = manager_for(name, g)
    A = manager.num_actions
    pool = TreePool(manager, T, default_node_capacity(manager, sims))
    pool.configure(sims, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, 0, 0.0, 0.25)
    if cache_log2 is not None:
        pool.enable_cache(cache_log2, torch.float64)
    ev = HashedDeviceEvaluator(pool)
    rngs = [random.Random(game_seed(entry_index, k)) for k in range(T)]
    hashes = [hashlib.sha256() for _ in range(T)]
    live = np.ones(T, bool)
    ply = moves = 0
    while live.any():
        ids = np.where(live)[0].astype(np.int32)
        stepped_search(pool, ev, sims, ids)
        visits, stats = pool.root_visits()
        order, count, root_visits = pool.root_children()
        actions = np.full(T, -1, np.int32)
        for t in ids:
            v = [int(x) for x in visits[t][:A]]
            o = [int(a) for a in order[t, : count[t]]]
            a = choose(rngs[t], ply, t % (pins["max_opening_plies"] + 1), o, v)
            absorb_move(hashes[t], v, (int(stats[t][0]), int(stats[t][1])), o, int(root_visits[t]), a)
            actions[t] = a
            moves += 1
        pool.advance(actions)
        _, _, _, fin, outcome = pool.states()
        for t in ids:
            if fin[t]:
                hashes[t].update(struct.pack("<f", float(outcome[t])))
                live[t] = False
        ply += 1
    assert moves == entry["moves"]
    got = [h.hexdigest() for h in hashes]
    wrong = [k for k in range(T) if got[k] != entry["digests"][k]]
    assert not wrong, f"{len(wrong)} of {T} games differ from the reference, first: game {wrong[0]}"
    c = pool.counters()
    assert c["capacity_errors"] == 0
    if cache_log2 is None:
        assert c["cache_hits"] == 0 and c["cache_aliases"] == 0
    else:
        # the games share their openings and transpose: the cache must actually have served leaves
        assert c["cache_hits"] + c["cache_aliases"] > 0.02 * c["leaf_evals"]
        assert c["cache_inserts"] <= c["leaf_evals"]


def selfplay_snapshot(name, g, T, sims, rounds, cache_log2, cutoff, alpha, seed=11, switch_at=None):
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import TreePool
    manager = manager_for(name, g)
    pool = TreePool(manager, T, 1 << 14, seed=seed)
    pool.configure(sims, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, cutoff, alpha, 0.25)
    pool.enable_selfplay(1 << 18)
    if cache_log2 is not None:
        pool.enable_cache(cache_log2, torch.float64)
    ev = HashedDeviceEvaluator(pool, salt=3)
    pri, val = ev.probs, ev.value
    live_total = 0
    for r in range(rounds):
        if r == switch_at:
            # a "weight transfer": from here on the evaluator is a different function of the state.  The batch that is
            # pending was evaluated by the old one, as in BatchedSelfPlay.sync_weights.
            ev.salt = 4
            pool.cache_clear()
        pool.round(pri, val, selfplay=True)
        pri, val = ev(dense=True)
        if r % 16 == 0:
            live_total += int(pool.leaf_count.item())
    torch.cuda.synchronize()
    c = pool.counters()
    assert c["capacity_errors"] == 0
    n = pool.positions_written()
    assert 0 < n <= pool.replay_capacity
    f = pool.replay_first[:n].cpu().numpy().view(np.uint64)
    s2 = pool.replay_second[:n].cpu().numpy().view(np.uint64)
    rows = np.concatenate([
        (f >> np.uint64(32)).astype(np.float64)[:, None], (f & np.uint64(0xffffffff)).astype(np.float64)[:, None],
        (s2 >> np.uint64(32)).astype(np.float64)[:, None], (s2 & np.uint64(0xffffffff)).astype(np.float64)[:, None],
        pool.replay_player[:n].cpu().numpy().astype(np.float64)[:, None],
        pool.replay_z[:n].cpu().numpy().astype(np.float64)[:, None],
        pool.replay_pi[:n].cpu().numpy().astype(np.float64)], axis=1)
    rows = rows[np.lexsort(rows.T[::-1])]          # the ring's write order depends on warp scheduling
    visits, stats = pool.root_visits()
    states = pool.states()
    return {"rows": rows, "visits": visits, "stats": stats, "states": states, "counters": c}


@pytest.mark.parametrize("name,g,T,sims,rounds", [("othello", 8, 512, 12, 900), ("hex", 7, 384, 14, 900),
                                                  ("hex_with_swap", 5, 256, 12, 500), ("tictactoe", 3, 256, 10, 300)])
def test_selfplay_identical_with_cache_on_and_off(name, g, T, sims, rounds):
    """Continuous self-play (dense leaf batch compacted through leaf_index, move sampling below the cutoff and
    device Dirichlet noise from the per-tree Philox streams) is a deterministic function of the seed whatever the
    cache does: every tree's root, visit counts and statistics after `rounds` rounds and the multiset of recorded
    examples must be identical with the cache off, on, and tiny."""
    base = selfplay_snapshot(name, g, T, sims, rounds, None, cutoff=6, alpha=0.5)
    assert base["counters"]["games"] > 0
    for log2 in (20, 10):
        got = selfplay_snapshot(name, g, T, sims, rounds, log2, cutoff=6, alpha=0.5)
        np.testing.assert_array_equal(got["visits"], base["visits"])
        np.testing.assert_array_equal(got["stats"], base["stats"])
        for a, b in zip(got["states"], base["states"]):
            np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(got["rows"], base["rows"])
        cb, cg = base["counters"], got["counters"]
        for k in ("simulations", "terminal_sims", "games", "moves", "depth_sum", "children_created"):
            assert cg[k] == cb[k], k
        # every non-terminal simulation got its evaluation from exactly one of the three sources
        assert cg["leaf_evals"] + cg["cache_hits"] + cg["cache_aliases"] == cb["leaf_evals"]
        assert cg["cache_hits"] + cg["cache_aliases"] > 0


def test_selfplay_engine_cache_on_off_identical_with_bf16_net():
    """BatchedSelfPlay with the tcgen05 network (Othello 8x8, CUDA-graph replay): the network is batch-independent
    bit for bit, so evaluating only the compacted cache misses must leave every tree exactly where evaluating
    every leaf leaves it."""
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from deep_mcts_b200.selfplay import BatchedSelfPlay
    out = []
    for log2 in (None, 20):
        torch.manual_seed(0)
        net = ConvolutionalOthelloNet(8)
        net.max_batch = 1024
        engine = BatchedSelfPlay(net, 1024, 12, sample_move_cutoff=8, dirichlet_alpha=1.0, replay_capacity=1 << 16,
                                 seed=9, eval_cache_log2=log2)
        engine.use_graph = True
        engine.run_rounds(260)
        torch.cuda.synchronize()
        pool = engine.pool
        out.append((pool.root_visits(), pool.states(), engine.counters()))
    (v0, s0), st0, c0 = out[0]
    (v1, s1), st1, c1 = out[1]
    np.testing.assert_array_equal(v0, v1)
    np.testing.assert_array_equal(s0, s1)
    for a, b in zip(st0, st1):
        np.testing.assert_array_equal(a, b)
    assert c0["simulations"] == c1["simulations"] and c0["moves"] == c1["moves"]
    assert c1["leaf_evals"] + c1["cache_hits"] + c1["cache_aliases"] == c0["leaf_evals"]
    assert c1["cache_hits"] + c1["cache_aliases"] > 0.05 * c0["leaf_evals"]   # 1024 games in lock-step share openings


def test_cache_clear_forgets_the_old_evaluator():
    """After dm_pool_cache_clear (weight transfer) no row of the old evaluator may be served: switching the
    evaluator mid-run gives the same trees with and without the cache."""
    base = selfplay_snapshot("othello", 6, 256, 16, 400, None, cutoff=4, alpha=0.5, switch_at=150)
    got = selfplay_snapshot("othello", 6, 256, 16, 400, 18, cutoff=4, alpha=0.5, switch_at=150)
    np.testing.assert_array_equal(got["visits"], base["visits"])
    np.testing.assert_array_equal(got["rows"], base["rows"])
    for a, b in zip(got["states"], base["states"]):
        np.testing.assert_array_equal(a, b)
    assert got["counters"]["cache_hits"] > 0
