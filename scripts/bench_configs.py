#!/usr/bin/env python
"""Throughput of the BASELINE.json configs that are not the headline bench line.

    python scripts/bench_configs.py rollouts   # config 4: Hex-with-swap 7x7, policy-net vs random rollouts
    python scripts/bench_configs.py train      # config 5: Othello 8x8 self-play + training step (+ NCCL all-reduce)

Each prints one JSON line per measurement (rank 0).  Launch with torch.distributed.run for N > 1.
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world


def midgame_roots(manager, n, plies, seed):
    """Synthetic reachable positions: `plies` uniformly random moves from the start (device playouts)."""
    rs = np.random.RandomState(seed)
    p, f, s = manager.initial_batch(n)
    for _ in range(plies):
        lst, _, cnt = manager.legal_batch(p, f, s)
        acts = np.array([lst[i, rs.randint(cnt[i])] if cnt[i] else 0 for i in range(n)], np.int32)
        p, f, s, ok = manager.child_batch(p, f, s, acts)
    fin, _ = manager.status_batch(p, f, s)
    keep = np.where(~fin)[0]
    idx = np.resize(keep, n)
    return p[idx], f[idx], s[idx]


def timed_steps(mcts, roots, steps):
    from deep_mcts_b200.distributed import reduce_counters
    p, f, s = roots
    sims = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        mcts.pool.set_root_states(p, f, s)
        _, stats = mcts.step()
        sims += int(stats.sum())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sums, maxima = reduce_counters({"sims": sims}, {"s": dt}, "cuda")
    return sums["sims"] / maxima["s"], maxima["s"] * 1000 / steps


def rollouts(rank, world):
    from deep_mcts_b200 import mcts as dm
    from deep_mcts_b200.gamenet import cached_state_evaluator
    from deep_mcts_b200.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet
    torch.manual_seed(0)
    g = 7
    net = ConvolutionalHexWithSwapNet(g)
    net.max_batch = 8192
    manager = net.manager
    cases = [
        ("random rollouts, no evaluator (simple agent)", 4096, 400, dm.random_rollout_policy(manager), None),
        ("random rollouts + network evaluator", 4096, 50, dm.random_rollout_policy(manager), cached_state_evaluator(net)),
        ("greedy policy-net rollouts, no evaluator (complex agent)", 2048, 20, dm.greedy_policy_rollout(net), None),
        ("greedy policy-net rollouts + network evaluator", 2048, 20, dm.greedy_policy_rollout(net),
         cached_state_evaluator(net)),
        # a playout is ~40 sequential network batches over the boards still in play: more games per batch, not a
        # faster kernel, is what raises this line
        ("greedy policy-net rollouts, no evaluator (complex agent)", 8192, 20, dm.greedy_policy_rollout(net), None),
    ]
    for name, trees, sims, ro, ev in cases:
        m = dm.BatchedMCTS(manager, trees, sims, ro, ev, rollout_share=1.0, seed=rank)
        roots = midgame_roots(manager, trees, 6, seed=rank)
        timed_steps(m, roots, 1)
        value, ms = timed_steps(m, roots, 3)
        c = m.pool.counters()
        if rank == 0:
            print(json.dumps({"config": f"hex_with_swap {g}x{g}: {name}", "metric": "mcts_simulations_per_sec",
                              "value": value, "n_gpus": world, "trees_per_gpu": trees, "sims_per_step": sims,
                              "ms_per_step": ms, "rollout_plies_per_sim": c["rollout_plies"] / max(1, c["simulations"])}))
        del m


def train(rank, world, overlap=True, graph_step=True):
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from deep_mcts_b200.train import SelfPlayTrainer, TrainingConfiguration
    torch.manual_seed(0)
    net = ConvolutionalOthelloNet(8)
    cfg = TrainingConfiguration(num_games=1000, num_simulations=50, save_interval=0, evaluation_interval=0,
                                save_dir="/tmp", sample_move_cutoff=10, dirichlet_alpha=1.0, replay_buffer_max_size=5000,
                                batch_size=1024 // world, n_trees=4096, train_steps_per_block=8, transfer_interval=16,
                                uniform_until_first_transfer=False, overlap_training=overlap, graph_train_step=graph_step)
    tr = SelfPlayTrainer(net, cfg)
    for _ in range(70):           # fill the replay ring with finished games first
        tr.self_play_block()
    torch.cuda.synchronize()
    for _ in range(3):
        tr.train_step()
    # all-reduce cost alone
    from deep_mcts_b200.distributed import reduce_counters
    bucket = net.grad_bucket          # every .grad is a view into one flat buffer: the all-reduce runs in place
    assert bucket is not None and bucket.attached()
    for _ in range(3):
        bucket.allreduce()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        n_elems = bucket.allreduce()
    e1.record()
    torch.cuda.synchronize()
    allreduce_us = e0.elapsed_time(e1) * 1000 / 20
    c0 = tr.engine.counters()
    w0 = tr.engine.positions_written()
    blocks = 10
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(blocks):
        tr.block(cfg.train_steps_per_block)     # self-play block and SGD steps, overlapped on two streams
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c1 = tr.engine.counters()
    # games all start together, so endings come in bursts: derive steady-state games/s from moves/s and
    # the mean game length observed since the pools were created
    mean_len = c1["moves"] / max(1, c1["games"])
    sums, maxima = reduce_counters({"moves": c1["moves"] - c0["moves"], "sims": c1["simulations"] - c0["simulations"],
                                    "steps": blocks * cfg.train_steps_per_block}, {"s": dt, "mean_len": mean_len}, "cuda")
    if rank == 0:
        print(json.dumps({"config": "othello 8x8 full loop: self-play 4096 games/GPU + SGD (global batch 1024)",
                          "n_gpus": world, "games_per_sec": sums["moves"] / maxima["mean_len"] / maxima["s"],
                          "positions_per_sec": sums["moves"] / maxima["s"], "mean_game_length": maxima["mean_len"],
                          "sims_per_sec": sums["sims"] / maxima["s"],
                          "train_steps_per_sec": sums["steps"] / world / maxima["s"],
                          "gradient_allreduce_us": allreduce_us, "gradient_elements": n_elems,
                          "train_steps_per_block": cfg.train_steps_per_block,
                          "overlap_training": cfg.overlap_training, "graph_train_step": cfg.graph_train_step}))


def engines(rank, world):
    """Game-engine kernels alone (dm_game_*): states/s, plies/s and achieved GB/s against the HBM peak.
    Bytes per state: 17 in (two bitboards + player) + the outputs of each call."""
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.hex.game import HexManager
    from deep_mcts_b200.othello.game import OthelloManager
    n = 1 << 20
    peak = 6538.9
    try:
        peak = float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"])
    except Exception:
        pass
    for manager, plies in ((OthelloManager(8), 20), (HexManager(7), 15)):
        p, f, s = midgame_roots(manager, 4096, plies, seed=1)
        idx = np.arange(n) % 4096
        dev = manager.device
        f_d = torch.from_numpy(f[idx].view(np.int64).copy()).to(dev)
        s_d = torch.from_numpy(s[idx].view(np.int64).copy()).to(dev)
        p_d = torch.from_numpy(p[idx].copy()).to(dev)
        lst = torch.empty((n, 65), dtype=torch.uint8, device=dev)
        order = torch.empty((n, 65), dtype=torch.uint8, device=dev)
        cnt = torch.empty(n, dtype=torch.int32, device=dev)
        fin = torch.empty(n, dtype=torch.uint8, device=dev)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        plies_d = torch.empty(n, dtype=torch.int32, device=dev)
        act = torch.zeros(n, dtype=torch.int32, device=dev)
        of, os_, op, ok = torch.empty_like(f_d), torch.empty_like(s_d), torch.empty_like(p_d), torch.empty_like(fin)
        st = torch.cuda.current_stream().cuda_stream
        gid, g = manager.game_id, manager.board
        calls = {
            "legal (list + set order)": (lambda: _lib.call("dm_game_legal", gid, g, n, f_d.data_ptr(), s_d.data_ptr(), p_d.data_ptr(), lst.data_ptr(), order.data_ptr(), cnt.data_ptr(), st), 17 + 130 + 4),
            "status": (lambda: _lib.call("dm_game_status", gid, g, n, f_d.data_ptr(), s_d.data_ptr(), p_d.data_ptr(), fin.data_ptr(), out.data_ptr(), st), 17 + 5),
            "child": (lambda: _lib.call("dm_game_child", gid, g, n, f_d.data_ptr(), s_d.data_ptr(), p_d.data_ptr(), act.data_ptr(), of.data_ptr(), os_.data_ptr(), op.data_ptr(), ok.data_ptr(), st), 17 + 4 + 18),
            "random rollout": (lambda: _lib.call("dm_game_rollout", gid, g, n, f_d.data_ptr(), s_d.data_ptr(), p_d.data_ptr(), _lib.DM_ROLLOUT_RANDOM, 7, out.data_ptr(), plies_d.data_ptr(), st), 17 + 8),
        }
        _lib.call("dm_game_legal", gid, g, n, f_d.data_ptr(), s_d.data_ptr(), p_d.data_ptr(), lst.data_ptr(), 0, cnt.data_ptr(), st)
        act.copy_(lst[:, 0].to(torch.int32))
        for name, (fn, nbytes) in calls.items():
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            line = {"config": f"{type(manager).__name__}({g}) dm_game: {name}", "states": n, "ms": ms,
                    "states_per_sec": n / (ms * 1e-3), "algorithmic_bytes_per_state": nbytes,
                    "achieved_GBps": n * nbytes / (ms * 1e-3) / 1e9, "hbm_peak_GBps": peak,
                    "frac": n * nbytes / (ms * 1e-3) / 1e9 / peak}
            if name == "random rollout":
                line["plies_per_sec"] = float(plies_d.sum().item()) / (ms * 1e-3)
            if rank == 0:
                print(json.dumps(line))


if __name__ == "__main__":
    rank, world = setup()
    which = sys.argv[1] if len(sys.argv) > 1 else "rollouts"
    if which.startswith("train"):
        # train | train-serial | train-eager | train-serial-eager
        train(rank, world, overlap="serial" not in which, graph_step="eager" not in which)
    else:
        {"rollouts": rollouts, "engines": engines}[which](rank, world)
    if world > 1:
        dist.destroy_process_group()
