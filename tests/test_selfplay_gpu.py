"""GPU parity of the continuous self-play kernels (move choice, re-rooting with subtree reuse,
example recording, outcome labelling, restart) against the oracle's MCTS.self_play +
create_self_play_examples semantics (reference mcts.py:96-111, train.py:258-283)."""
import numpy as np
import pytest
import torch

from oracle import evaluators as oev
from oracle.games import make_game
from oracle.mcts import OracleMCTS

pytestmark = pytest.mark.gpu

from test_games_gpu import manager_for  # noqa: E402


def evaluate_leaves(pool, game, evaluator, n, dense=False):
    """Host evaluation of the pending leaf batch.  ``dense``: the batch is the dense one of
    ``pool.round(selfplay=True)`` -- its n live slots are listed in ``pool.leaf_index`` and every result goes to
    its own slot."""
    from deep_mcts_b200 import _lib
    T = pool.n_trees
    pri = np.zeros((T, _lib.DM_MAX_ACTIONS), np.float64)
    val = np.zeros(T, np.float64)
    if n:
        slots = pool.leaf_index[:n].cpu().numpy() if dense else np.arange(n)
        p = pool.leaf_player.cpu().numpy()
        f = pool.leaf_first.cpu().numpy().view(np.uint64)
        s = pool.leaf_second.cpu().numpy().view(np.uint64)
        for i in slots:
            v, probs = evaluator(game, (int(p[i]), int(f[i]), int(s[i])))
            val[i] = v
            pri[i, : game.num_actions] = probs
    return torch.from_numpy(pri).to(pool.device), torch.from_numpy(val).to(pool.device)


def host_round(pool, game, evaluator, carry=None):
    """One self-play round with a HOST evaluator plugged into the leaf exchange.  ``carry`` (a dict)
    selects the fused form: ONE tree launch expands the batch evaluated in the previous call and
    selects the next one (dm_pool_round, dense slots: tree t's leaf is slot t)."""
    if carry is None:
        pool.selfplay_select()
        pri, val = evaluate_leaves(pool, game, evaluator, int(pool.leaf_count.item()))
        pool.expand_backup(pri, val)
        return
    if "pri" not in carry:
        carry["pri"], carry["val"] = evaluate_leaves(pool, game, evaluator, 0)
    pool.round(carry["pri"], carry["val"], selfplay=True)
    n = int(pool.leaf_count.item())
    assert 0 <= n <= pool.n_trees
    carry["pri"], carry["val"] = evaluate_leaves(pool, game, evaluator, n, dense=True)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name,g,sims", [("othello", 6, 30), ("hex_with_swap", 4, 25), ("tictactoe", 3, 20), ("othello", 8, 20),
                                         ("hex", 7, 20)])
def test_selfplay_games_match_oracle(name, g, sims, fused):
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import TreePool
    game = make_game(name, g)
    manager = manager_for(name, g)
    T = 3
    pool = TreePool(manager, T, 1 << 15, seed=5)
    pool.configure(sims, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, 0, 0.0, 0.25)   # cutoff 0, no noise: deterministic
    pool.enable_selfplay(4096)
    # oracle: one full deterministic game (all trees play the same game, then restart and repeat it)
    ref = OracleMCTS(game, sims, None, oev.hashed)
    examples = [(old, dist) for old, new, action, dist in ref.self_play()]
    outcome = game.outcome(ref.state)
    plies = len(examples)
    # enough rounds for every tree to finish one game (+ slack for root expansions / terminal sims)
    carry = {} if fused else None
    for _ in range(plies * (sims + 2) + 5):
        host_round(pool, game, oev.hashed, carry)
        if pool.counters()["games"] >= T:
            break
    c = pool.counters()
    assert c["games"] >= T and c["capacity_errors"] == 0
    written = pool.positions_written()
    assert written >= T * plies
    first = pool.replay_first[:written].cpu().numpy().view(np.uint64)
    second = pool.replay_second[:written].cpu().numpy().view(np.uint64)
    player = pool.replay_player[:written].cpu().numpy()
    pi = pool.replay_pi[:written].cpu().numpy()
    z = pool.replay_z[:written].cpu().numpy()
    z_abs = outcome * 2 - 1
    for t in range(T):   # every tree wrote one contiguous game
        for i, (state, dist) in enumerate(examples):
            r = t * plies + i
            assert (int(player[r]), int(first[r]), int(second[r])) == state
            np.testing.assert_array_equal(pi[r, : game.num_actions], np.asarray(dist, np.float32))
            assert z[r] == (z_abs if state[0] == 0 else -z_abs)   # train.py:270-282


def test_selfplay_with_device_net_runs_and_labels_consistently():
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from deep_mcts_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(0)
    net = ConvolutionalOthelloNet(4)
    engine = BatchedSelfPlay(net, 256, 8, sample_move_cutoff=4, dirichlet_alpha=1.0, replay_capacity=1 << 16, seed=3)
    engine.run_rounds(400)
    c = engine.counters()
    assert c["games"] > 100 and c["capacity_errors"] == 0 and c["moves"] > c["games"] * 4
    n = min(engine.positions_written(), 1 << 16)
    pool = engine.pool
    pi = pool.replay_pi[:n, :17].cpu().numpy()
    z = pool.replay_z[:n].cpu().numpy()
    np.testing.assert_allclose(pi.sum(axis=1), 1.0, atol=1e-5)
    assert set(np.unique(z)).issubset({-1.0, 0.0, 1.0}) and (z != 0).any()
    f, s, p, pi_b, z_b = engine.sample_batch(64)
    assert pi_b.shape == (64, 17) and z_b.shape == (64, 1) and f.is_cuda
    # multi-pool variant
    engine2 = BatchedSelfPlay(net, 256, 8, sample_move_cutoff=4, dirichlet_alpha=1.0, replay_capacity=1 << 16,
                              seed=4, n_pools=2)
    engine2.run_rounds(100)
    torch.cuda.synchronize()
    assert engine2.counters()["games"] > 10
    f2, s2, p2, pi2, z2 = engine2.sample_batch(48)
    assert pi2.shape == (48, 17) and z2.shape == (48, 1) and f2.shape == (48,)
    np.testing.assert_allclose(pi2.sum(dim=1).cpu().numpy(), 1.0, atol=1e-5)
    # sampling restricted to the last few GAMES (reference train.py:149-151 trims the buffer in games)
    f3, _, _, pi3, _ = engine2.sample_batch(16, max_games=6)
    assert f3.shape == (16,) and pi3.shape == (16, 17)


def test_pipelined_pool_graph_accounting():
    """Uneven pools replayed as one two-branch CUDA graph (pool 1 runs a phase behind pool 0): every
    replay is one round per pool, simulations are counted once, and switching between the graph and
    launch-by-launch modes completes the half-done round instead of expanding a leaf twice."""
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from deep_mcts_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(0)
    net = ConvolutionalOthelloNet(4)
    sims = 8
    engine = BatchedSelfPlay(net, 96, sims, sample_move_cutoff=4, dirichlet_alpha=1.0, replay_capacity=1 << 14,
                             seed=5, n_pools=2, pool_sizes=[64, 32])
    engine.use_graph = True
    engine.run_rounds(90)
    engine.use_graph = False
    engine.run_rounds(9)
    engine.use_graph = True
    engine.run_rounds(90)
    torch.cuda.synchronize()
    c = engine.counters()
    assert c["capacity_errors"] == 0 and c["games"] > 20
    assert c["simulations"] > 0.8 * 96 * 180      # a round is at least one simulation for (nearly) every tree
    # every finished move took exactly `sims` simulations; at most one move per tree is in progress
    assert c["moves"] * sims <= c["simulations"] <= (c["moves"] + 96) * sims
    for pool in engine.pools:
        n = min(pool.positions_written(), pool.replay_capacity)
        pi = pool.replay_pi[:n, :17].cpu().numpy()
        np.testing.assert_allclose(pi.sum(axis=1), 1.0, atol=1e-5)


def test_weight_transfer_keeps_graph_valid_and_changes_outputs():
    """Re-packing weights (the reference's weight transfer, train.py:169-171) must reach every kernel argument:
    outputs follow the new weights on the eager path AND on the CUDA-graph replay of the self-play round (the
    captured graph is dropped at the transfer), and the evaluation cache forgets the old network."""
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from deep_mcts_b200.selfplay import BatchedSelfPlay
    from oracle import net as onet
    from oracle.games import make_game
    from oracle.synth import random_positions, synthetic_state_dict
    game = make_game("othello", 6)
    states = random_positions(game, 64, seed=3)
    p = [s[0] for s in states]; f = [s[1] for s in states]; s2 = [s[2] for s in states]
    net = ConvolutionalOthelloNet(6)
    net.precision = "fp32"
    sd_a = synthetic_state_dict(game, 128, 3, 128, seed=1, conv_gain=1.0, fc_scale=1.0)
    sd_b = synthetic_state_dict(game, 128, 3, 128, seed=2, conv_gain=1.0, fc_scale=1.0)
    net.load_state_dict(sd_a)
    engine = BatchedSelfPlay(net, 128, 8, sample_move_cutoff=2, dirichlet_alpha=1.0, seed=1)
    engine.use_graph = True
    engine.run_rounds(30)
    dn = engine.device_nets[0]
    va, pa = dn.evaluate(p, f, s2)
    want_a = onet.evaluate_batch(game, sd_a, states)
    np.testing.assert_allclose(pa, want_a[1], atol=1e-5, rtol=0)
    net.load_state_dict(sd_b)
    engine.sync_weights()
    vb, pb = dn.evaluate(p, f, s2)
    want_b = onet.evaluate_batch(game, sd_b, states)
    np.testing.assert_allclose(pb, want_b[1], atol=1e-5, rtol=0)
    np.testing.assert_allclose(vb, want_b[0], atol=1e-5, rtol=0)
    assert np.abs(pa - pb).max() > 1e-3
    before = engine.counters()["simulations"]
    engine.run_rounds(30)          # the round graph is re-captured: it froze the old biases and value-head weights
    torch.cuda.synchronize()
    c = engine.counters()
    assert c["simulations"] > before and c["capacity_errors"] == 0
    # the outputs the REPLAYED round left pending must come from the new weights (by-value kernel arguments
    # included), for every leaf the round evaluated
    pool = engine.pool
    n = int(pool.leaf_count.item())
    assert n > 0
    slots = pool.leaf_index[:n].cpu().numpy()
    lp = pool.leaf_player.cpu().numpy()[slots]
    lf = pool.leaf_first.cpu().numpy().view(np.uint64)[slots]
    ls = pool.leaf_second.cpu().numpy().view(np.uint64)[slots]
    want_v, want_p = onet.evaluate_batch(game, sd_b, [(int(a), int(b), int(c2)) for a, b, c2 in zip(lp, lf, ls)])
    got_p = dn._probs[torch.from_numpy(slots).long().to(pool.device)][:, : game.num_actions].cpu().numpy()
    got_v = dn._value[torch.from_numpy(slots).long().to(pool.device)].cpu().numpy()
    np.testing.assert_allclose(got_p, want_p, atol=1e-5, rtol=0)
    np.testing.assert_allclose(got_v, want_v, atol=1e-5, rtol=0)


def test_replay_window_in_games():
    """The replay ring numbers games in ring order (one atomic hands out positions and game number): the window of
    the last N games starts where game (games - N) starts, and a sample restricted to it only returns positions
    of those games (reference train.py:149-151)."""
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import TreePool
    from deep_mcts_b200.selfplay import BatchedSelfPlay
    manager = manager_for("tictactoe", 3)
    pool = TreePool(manager, 64, 1 << 12, seed=2)
    pool.configure(6, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, 3, 0.5, 0.25)
    pool.enable_selfplay(1 << 15)
    game = make_game("tictactoe", 3)
    carry = {}
    for _ in range(260):
        host_round(pool, game, oev.uniform, carry)
    positions, games = pool.replay_progress()
    assert games == pool.counters()["games"] and games > 100 and positions <= pool.replay_capacity
    ends = pool.replay_game_end[:games].cpu().numpy()
    assert (np.diff(ends) >= 5).all() and (np.diff(ends) <= 9).all() and ends[-1] == positions   # 5..9 plies per game
    start = pool.replay_window_start(10)
    assert start == ends[games - 11]
    # positions of one game are contiguous and start at ply 0: an empty board with FIRST to move
    first = pool.replay_first[:positions].cpu().numpy()
    second = pool.replay_second[:positions].cpu().numpy()
    starts = np.concatenate([[0], ends[:-1]])
    assert (first[starts] == 0).all() and (second[starts] == 0).all()
    assert pool.replay_window_start(10 ** 6) == 0


def test_debug_timing_reports_every_warp():
    """dm_pool_debug_timing (the diagnostic behind scripts/tree_tail_probe.py): after a self-play launch every tree
    has non-zero cycle counts for the selection phase and a plausible level count."""
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import TreePool, wrap_device_buffer
    manager = manager_for("othello", 6)
    game = make_game("othello", 6)
    pool = TreePool(manager, 64, 1 << 12, seed=1)
    pool.configure(8, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, 2, 0.5, 0.25)
    pool.enable_selfplay(1 << 12)
    ptr = _lib.c_void_p()
    _lib.call("dm_pool_debug_timing", pool._h, ptr)
    dbg = wrap_device_buffer(ptr.value, (64, 4), torch.int64, pool.device)
    carry = {}
    for _ in range(12):
        host_round(pool, game, oev.hashed, carry)
    torch.cuda.synchronize()
    d = dbg.cpu().numpy()
    assert (d[:, 1] > 0).all() and (d[:, 0] >= 0).all()
    assert ((d[:, 2] & 0xffff) <= 128).all()
