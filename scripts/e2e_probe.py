#!/usr/bin/env python
"""Where the end-to-end step's time goes: bench.py's host-driven self-play step (BatchedMCTS.step -> host move choice
-> advance -> states -> reset) with a host clock and a device synchronisation around every segment.

    python scripts/e2e_probe.py [--steps 40] [--workload othello8] > profiles/e2e_probe_r2.jsonl

The synchronisations make the step a little slower than the un-instrumented one; what matters is the split between
device search time and everything the host adds around it.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--trees", type=int, default=4096)
    ap.add_argument("--workload", default="othello8")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    from deep_mcts_b200.gamenet import cached_state_evaluator
    from deep_mcts_b200.mcts import BatchedMCTS

    torch.manual_seed(0)
    if args.workload == "othello8":
        from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
        net = ConvolutionalOthelloNet(8)
    else:
        from deep_mcts_b200.hex.convolutionalnet import ConvolutionalHexNet
        net = ConvolutionalHexNet(7)
    net.precision = "bf16"
    net.max_batch = args.trees
    T, sims, cutoff = args.trees, 50, 10
    searcher = BatchedMCTS(net.manager, T, sims, None, cached_state_evaluator(net), seed=7, dirichlet_alpha=1.0,
                           dirichlet_factor=0.25)
    h_ply = np.zeros(T, dtype=np.int64)
    rng = np.random.RandomState(1)
    seg = {}

    def clock(name, fn):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        seg[name] = seg.get(name, 0.0) + time.perf_counter() - t0
        return out

    def choose(visits):
        actions = torch.from_numpy(visits).argmax(dim=1).numpy()
        early = np.where(h_ply < cutoff)[0]
        if early.size:
            cdf = np.cumsum(visits[early].astype(np.float64), axis=1)
            draw = rng.random_sample(early.size) * cdf[:, -1]
            actions[early] = (cdf > draw[:, None]).argmax(axis=1)
        return actions.astype(np.int32)

    def step(record):
        pool = searcher.pool
        if record:
            clock("search (device: configure, begin_step, 51 rounds)", lambda: searcher.run_simulations(sims))
            visits, stats = clock("root_visits (kernel + D2H)", pool.root_visits)
            clock("counters (capacity check)", pool.counters)
            actions = clock("host move choice (numpy)", lambda: choose(visits))
            clock("advance (H2D + re-root kernel)", lambda: searcher.advance(actions))
            st = clock("states (kernel + D2H)", pool.states)
        else:
            visits, stats = searcher.step()
            actions = choose(visits)
            searcher.advance(actions)
            st = pool.states()
        h_ply[:] += 1
        fin = st[3]
        if fin.any():
            done = np.where(fin)[0].astype(np.int32)
            if record:
                clock("reset finished games (H2D + kernel)", lambda: pool.reset(done))
            else:
                pool.reset(done)
            h_ply[done] = 0
        return int(stats.sum())

    for _ in range(25):
        step(False)
    torch.cuda.synchronize()
    # un-instrumented step time
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        n += step(False)
    torch.cuda.synchronize()
    plain = time.perf_counter() - t0
    t0 = time.perf_counter()
    m = 0
    for _ in range(args.steps):
        m += step(True)
    torch.cuda.synchronize()
    instrumented = time.perf_counter() - t0
    out = {"workload": args.workload, "trees": T, "steps": args.steps,
           "plain_ms_per_step": 1e3 * plain / args.steps, "plain_sims_per_s": n / plain,
           "instrumented_ms_per_step": 1e3 * instrumented / args.steps,
           "segments_ms_per_step": {k: round(1e3 * v / args.steps, 4) for k, v in seg.items()}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
