#!/usr/bin/env python
"""Where the joules go: the self-play round taken apart, every part run alone at 100 % duty for a few seconds with the
NVML energy counter around it (Othello 8x8, 4096 boards / trees, bf16 network).

    python scripts/energy_probe.py [--seconds 3] > profiles/energy_probe_r2.jsonl

  net_only    the network kernels back to back on 4096 reachable positions (encode .. heads)
  tree_only   the tree kernel back to back with a constant evaluator output (no network)
  full_round  the real loop (BatchedSelfPlay, CUDA graph, evaluation cache on)
For each: time per iteration, mean watts, joules per iteration, SM clock.  If the board sits at its power limit in
net_only and full_round alike, simulations/s is set by joules per simulation, not by what overlaps with what.
"""
import argparse
import json
import subprocess
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class Energy:
    def __init__(self, index=0):
        import pynvml
        pynvml.nvmlInit()
        self.nvml = pynvml
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)

    def joules(self):
        return self.nvml.nvmlDeviceGetTotalEnergyConsumption(self.h) / 1000.0

    def sm_mhz(self):
        return self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM)


def timed(fn, seconds, energy, sync_every=64):
    for _ in range(32):
        fn()
    torch.cuda.synchronize()
    n, clocks = 0, []
    j0, t0 = energy.joules(), time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(sync_every):
            fn()
        torch.cuda.synchronize()
        clocks.append(energy.sm_mhz())
        n += sync_every
    dt, dj = time.perf_counter() - t0, energy.joules() - j0
    return {"iterations": n, "us_per_iteration": 1e6 * dt / n, "mean_watts": dj / dt, "joules_per_iteration": dj / n,
            "sm_mhz_median": float(np.median(clocks))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--trees", type=int, default=4096)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
    from deep_mcts_b200.pool import TreePool
    from deep_mcts_b200.selfplay import BatchedSelfPlay
    energy = Energy(0)
    T = args.trees
    torch.manual_seed(0)
    net = ConvolutionalOthelloNet(8)
    net.precision = "bf16"
    net.max_batch = T
    engine = BatchedSelfPlay(net, T, 50, 10, 1.0, 0.25, replay_capacity=1 << 20, seed=1)
    engine.use_graph = True
    engine.run_rounds(1500)                      # mid-game positions, warm cache
    torch.cuda.synchronize()
    out = []

    # --- full round
    c0 = engine.counters()
    r = timed(lambda: engine.run_rounds(1), args.seconds, energy)
    c1 = engine.counters()
    sims = c1["simulations"] - c0["simulations"]
    evals = c1["leaf_evals"] - c0["leaf_evals"]
    r.update(part="full_round", sims_per_iteration=sims / r["iterations"], evals_per_iteration=evals / r["iterations"],
             joules_per_million_sims=1e6 * r["joules_per_iteration"] / (sims / r["iterations"]))
    out.append(r)

    # --- net only: 4096 mid-game root states through the whole network, every board live
    p, f, s, fin, _ = engine.pool.states()
    dn = engine.device_nets[0]
    dev = engine.pool.device
    d_f = torch.from_numpy(f.view(np.int64).copy()).to(dev)
    d_s = torch.from_numpy(s.view(np.int64).copy()).to(dev)
    d_p = torch.from_numpy(p.copy()).to(dev)
    g = torch.cuda.CUDAGraph()
    dn.evaluate_device(d_f, d_s, d_p, T)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        dn.evaluate_device(d_f, d_s, d_p, T)
    r = timed(g.replay, args.seconds, energy)
    flop = T * 133.66e6
    r.update(part="net_only", boards=T, tflops_useful=flop / (r["us_per_iteration"] * 1e-6) / 1e12,
             joules_per_tflop_useful=r["joules_per_iteration"] / (flop / 1e12),
             evals_per_second=T / (r["us_per_iteration"] * 1e-6))
    out.append(r)

    # --- tree only: constant evaluator output, the tree kernel back to back (self-play semantics, no cache)
    manager = net.manager
    pool = TreePool(manager, T, 1 << 15, seed=3)
    pool.configure(50, _lib.DM_EVAL_EXTERNAL, _lib.DM_ROLLOUT_NONE, 10, 1.0, 0.25)
    pool.enable_selfplay(1 << 20)
    pri = torch.full((T, _lib.DM_MAX_ACTIONS), 1.0 / 65, dtype=torch.float32, device=dev)
    val = torch.full((T,), 0.5, dtype=torch.float32, device=dev)
    for _ in range(1500):
        pool.round(pri, val, selfplay=True)
    torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        pool.round(pri, val, selfplay=True)
    c0 = pool.counters()
    r = timed(g2.replay, args.seconds, energy)
    c1 = pool.counters()
    r.update(part="tree_only", sims_per_iteration=(c1["simulations"] - c0["simulations"]) / r["iterations"])
    out.append(r)

    idle = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,power.limit", "--format=csv,noheader,nounits"],
                          capture_output=True, text=True).stdout.strip()
    for row in out:
        row["power_after_probe_w_and_limit"] = idle
        print(json.dumps(row))


if __name__ == "__main__":
    main()
