"""Distributional parity of the device-sampled paths at the BASELINE sizes.  The device streams are Philox, not
Python's ``random`` / numpy's legacy generator, so these paths are compared as distributions:
  (a) Dirichlet root noise            reference mcts.py:233-241 (np.random.dirichlet)
  (b) move sampling below the cutoff  reference mcts.py:96-104  (np.random.choice(p=pi))
  (c) rollout_share mixing            reference mcts.py:187-196 (random.random() < rollout_share)
  (d) uniform random rollouts         reference mcts.py:219-226 (random.choice(legal_actions))
"""
import json

import numpy as np
import pytest
import torch

from oracle import evaluators as oev
from oracle.games import make_game
from oracle.mcts import OracleMCTS

pytestmark = pytest.mark.gpu

from test_games_gpu import manager_for  # noqa: E402

N_TREES = 100_000


def _root_state_with_children(name, g, stones):
    """A non-final state reached by `stones` seeded random plies."""
    import random
    game = make_game(name, g)
    rng = random.Random(5)
    while True:
        s = game.initial_state()
        for _ in range(stones):
            s = game.child(s, rng.choice(game.legal_list(s)))
        if not game.is_final(s):
            return game, s


@pytest.mark.parametrize("alpha", [0.33, 1.0, 2.5])
@pytest.mark.parametrize("name,g,stones,n", [("othello", 8, 0, 4), ("tictactoe", 3, 0, 9), ("hex", 6, 6, 30)])
def test_device_dirichlet_noise_distribution(name, g, stones, n, alpha):
    """10^5 draws of the device gamma/Dirichlet sampler (dirichlet_factor = 1, so a root child's prior IS its noise
    component): every draw sums to one, every component has the Dirichlet(alpha) mean 1/n and variance
    (n-1)/(n^2 (n alpha + 1)) within 5 sigma, its marginal is Beta(alpha, (n-1) alpha) (Kolmogorov-Smirnov), and
    two components have the Dirichlet covariance -1/(n^2 (n alpha + 1))."""
    from scipy import stats
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import TreePool
    game, root = _root_state_with_children(name, g, stones)
    assert len(set(game.legal_list(root))) == n
    manager = manager_for(name, g)
    T = N_TREES
    pool = TreePool(manager, T, 160, seed=12345)
    pool.configure(1, _lib.DM_EVAL_UNIFORM, _lib.DM_ROLLOUT_NONE, 0, alpha, 1.0)
    pool.set_root_states(np.full(T, root[0], np.uint8), np.full(T, root[1], np.uint64), np.full(T, root[2], np.uint64))
    pool.begin_step()
    pool.search_local()
    prior, _ = pool.root_child_stats()
    order, count, _ = pool.root_children()
    assert (count == n).all() and pool.counters()["capacity_errors"] == 0
    x = prior[:, :n]
    assert np.all(x > 0) and np.all(prior[:, n:] == 0)
    np.testing.assert_allclose(x.sum(axis=1), 1.0, atol=1e-12)
    mean, var = 1.0 / n, (n - 1) / (n * n * (n * alpha + 1))
    m = x.mean(axis=0)
    assert np.all(np.abs(m - mean) < 5 * np.sqrt(var / T)), (m, mean)
    v = x.var(axis=0)
    m4 = ((x - m) ** 4).mean(axis=0)
    sigma_v = np.sqrt(np.maximum(m4 - v * v, 1e-30) / T)
    assert np.all(np.abs(v - var) < 5 * sigma_v + 1e-9), (v, var)
    for j in (0, n - 1):
        p = stats.kstest(x[:, j], stats.beta(alpha, (n - 1) * alpha).cdf).pvalue
        assert p > 1e-4, (j, p)
    cov = np.mean((x[:, 0] - mean) * (x[:, 1] - mean))
    want = -1.0 / (n * n * (n * alpha + 1))
    prod = (x[:, 0] - mean) * (x[:, 1] - mean)
    assert abs(cov - want) < 5 * prod.std() / np.sqrt(T), (cov, want)
    # and it behaves like numpy's sampler on a statistic that mixes all components: the largest one
    ref = np.random.RandomState(0).dirichlet([alpha] * n, size=T).max(axis=1)
    assert stats.ks_2samp(x.max(axis=1), ref).pvalue > 1e-4


def _spreading_evaluator(game):
    """value 1.0 everywhere and one fixed non-uniform prior row: every visited child of a FIRST-to-move root scores
    like an unvisited one (mcts.py:45-54), so the exploration term spreads the visits over many children -- with
    ordinary evaluators the reference's first-visited child keeps nearly all of them."""
    w = np.array([1.0 + (a * 7) % 13 for a in range(game.num_actions)])
    w /= w.sum()
    return (lambda g, s: (1.0, w.tolist())), w


@pytest.mark.parametrize("name,g,sims", [("othello", 8, 50), ("hex", 7, 50), ("hex_with_swap", 7, 50)])
def test_move_sampling_follows_the_visit_distribution(name, g, sims):
    """MCTS.self_play below sample_move_cutoff (mcts.py:96-104): 10^5 trees search the same root with the same
    deterministic evaluator, so all hold the same visit vector (the oracle's); each samples its move from its own
    Philox stream.  The histogram of the chosen actions must fit pi = visits / total (chi-square), and above the
    cutoff every tree must play argmax."""
    from scipy import stats
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import TreePool
    game = make_game(name, g)
    manager = manager_for(name, g)
    evaluator, w = _spreading_evaluator(game)
    ref = OracleMCTS(game, sims, None, evaluator)
    ref.step()
    visits = np.array(ref.root_visit_counts(), dtype=np.float64)
    assert (visits > 0).sum() >= 4
    pi = visits / visits.sum()
    root = game.initial_state()
    child_of = {game.child(root, a): a for a in set(game.legal_list(root))}
    for cutoff in (1, 0):
        T = N_TREES if cutoff else 4096
        pool = TreePool(manager, T, 4096, seed=99)
        pool.configure(sims, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, cutoff, 0.0, 0.25)
        pool.enable_selfplay(1 << 10)
        # the evaluator's output is the same for every leaf: constant device buffers, no evaluation launch needed
        pri = torch.zeros((T, _lib.DM_MAX_ACTIONS), dtype=torch.float64, device=pool.device)
        pri[:, : game.num_actions] = torch.from_numpy(w).to(pool.device)
        val = torch.ones(T, dtype=torch.float64, device=pool.device)
        for _ in range(sims + 2):          # root expansion + sims rounds + the round that makes the move
            pool.round(pri, val, selfplay=True)
        c = pool.counters()
        assert c["moves"] == T and c["capacity_errors"] == 0
        p, f, s, _, _ = pool.states()
        actions = np.array([child_of[(int(a), int(b), int(d))] for a, b, d in zip(p, f, s)])
        hist = np.bincount(actions, minlength=game.num_actions).astype(np.float64)
        if cutoff == 0:
            assert hist[int(np.argmax(visits))] == T      # np.argmax: first maximum
            continue
        assert hist[pi == 0].sum() == 0
        live = pi > 0
        chi2 = (((hist - T * pi) ** 2)[live] / (T * pi[live])).sum()
        pvalue = stats.chi2.sf(chi2, int(live.sum()) - 1)
        assert pvalue > 1e-4, (chi2, pvalue, hist[live], T * pi[live])
        del pool


@pytest.mark.parametrize("share", [0.3, 0.8])
def test_rollout_share_mixes_the_right_fraction(share):
    """evaluate_leaf (mcts.py:187-196): with a state evaluator AND a rollout policy a simulation's value is
    (rollout + value) / 2 with probability rollout_share, else the evaluator's value.  One simulation per tree on
    10^5 Othello 8x8 trees: the visited child's value_sum tells which branch ran."""
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import TreePool
    manager = manager_for("othello", 8)
    T = N_TREES
    pool = TreePool(manager, T, 96, seed=7)
    pool.configure(1, _lib.DM_EVAL_HASHED, _lib.DM_ROLLOUT_RANDOM, 0, 0.0, 0.25, rollout_share=share, eval_salt=5)
    pool.begin_step()
    pool.search_local()
    visits, stats = pool.root_visits()
    _, value_sum = pool.root_child_stats()
    order, count, _ = pool.root_children()
    assert (stats[:, 0] == 1).all() and (visits.sum(axis=1) == 1).all()
    col = int(np.argmax(value_sum[0] != 0)) if (value_sum[0] != 0).any() else 0
    visited_action = int(np.argmax(visits[0]))
    assert (np.argmax(visits, axis=1) == visited_action).all()      # deterministic selection: the same child everywhere
    j = int(np.where(order[0, : count[0]] == visited_action)[0][0])
    vs = value_sum[:, j]
    values, counts = np.unique(vs, return_counts=True)
    v = values[np.argmax(counts)] if share < 0.5 else None
    # the evaluator's value for that child, from the oracle
    game = make_game("othello", 8)
    leaf = game.child(game.initial_state(), visited_action)
    want_v = oev.hashed_salted(5)(game, leaf)[0]
    if v is not None:
        assert v == want_v
    assert want_v not in (0.0, 0.5, 1.0)
    allowed = {want_v} | {(r + want_v) / 2 for r in (0.0, 0.5, 1.0)}
    assert set(values.tolist()) <= allowed
    rolled = float(np.mean(vs != want_v))
    sigma = np.sqrt(share * (1 - share) / T)
    assert abs(rolled - share) < 5 * sigma, (rolled, share)
    assert pool.counters()["rollout_plies"] > 0


@pytest.mark.parametrize("key", ["othello_8", "hex_7", "hex_with_swap_7"])
def test_random_rollout_rates_at_baseline_sizes(golden_dir, key):
    """Uniform random playouts at the BASELINE board sizes (mcts.py:219-226) from the initial state, mid-game states
    and the Hex-with-swap state in which swapping is legal: 2*10^5 CUDA playouts per state against 2*10^4 oracle
    playouts and 2*10^3 playouts of the unmodified reference (tests/golden/rollouts.json), win / draw / loss rates
    within 5-sigma binomial bounds of both."""
    data = json.load(open(golden_dir / "rollouts.json"))
    name, g = key.rsplit("_", 1)
    manager = manager_for(name, int(g))
    n_gpu = 200_000
    swap_state_seen = False
    for k, row in enumerate(data["games"][key]):
        p, f, s = row["state"]
        if name == "hex_with_swap" and p == 0 and bin(f | s).count("1") == 1:
            swap_state_seen = True
        out, plies = manager.rollout_batch(np.full(n_gpu, p, np.uint8), np.full(n_gpu, f, np.uint64),
                                           np.full(n_gpu, s, np.uint64), seed=1000 + k)
        assert set(np.unique(out).tolist()) <= {0.0, 0.5, 1.0} and plies.min() >= 1
        for src in ("oracle", "reference"):
            counts = np.array(row[src], dtype=np.float64)
            n_cpu = counts.sum()
            for i, v in enumerate((0.0, 0.5, 1.0)):
                p_gpu, p_cpu = float(np.mean(out == v)), counts[i] / n_cpu
                pooled = (p_gpu * n_gpu + counts[i]) / (n_gpu + n_cpu)
                sigma = np.sqrt(max(pooled * (1 - pooled), 1e-6) * (1 / n_gpu + 1 / n_cpu))
                assert abs(p_gpu - p_cpu) < 5 * sigma, (key, row["state"], src, v, p_gpu, p_cpu)
    assert swap_state_seen or name != "hex_with_swap"
