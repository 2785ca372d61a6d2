#!/usr/bin/env python
"""Headline benchmark: batched MCTS self-play throughput (simulations/s and NN leaf evals/s).

    python bench.py --gpus N --steps K --warmup W [--impl b200|reference] [--workload othello8|hex7]

One "step" = 10 moves of every game = 10 x ``num_simulations`` (500) search rounds over the whole pool: every
round selects one leaf per tree (PUCT descent, terminal leaves handled in place), evaluates the leaves the
device evaluation cache does not know with the conv policy/value net and expands + backs up.  Workload at N=1: Othello 8x8 self-play, 4096
concurrent games, ConvolutionalOthelloNet(8) random-init weights, 50 sims/move, cutoff 10,
Dirichlet alpha 1.0 (BASELINE.json configs[2] at one GPU; --workload hex7 runs configs[1]).
Games shard across GPUs with no collective on the search path (weak scaling).

Prints ONE JSON line (see the contract in DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (game, grid, num_simulations, sample_move_cutoff, dirichlet_alpha, MFLOP per leaf per 3x3 trunk conv)
    "othello8": ("othello", 8, 50, 10, 1.0),
    "hex7": ("hex", 7, 50, 10, 0.33),
}
METRIC = "mcts_simulations_per_sec"
UNIT = "sims/s"
MOVES_PER_STEP = 10   # one step = 10 moves of every game = 10 x num_simulations search rounds over the pool


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="othello8", choices=sorted(WORKLOADS))
    ap.add_argument("--trees", type=int, default=4096, help="concurrent games per GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample per process")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=60, help="moves per game of the end-to-end leg (60 = about one whole game per tree)")
    ap.add_argument("--profile-steps", type=int, default=2, help="steps of the per-kernel CUDA-event pass")
    ap.add_argument("--pools", type=int, default=1, help="independent pools per GPU on separate streams")
    ap.add_argument("--pool-sizes", type=str, default="", help="comma-separated games per pool (default: even split)")
    ap.add_argument("--graph", type=int, default=1, help="1 = replay a captured CUDA graph of the round")
    ap.add_argument("--max-nodes", type=int, default=0, help="node arena per tree (0 = the engine's default)")
    ap.add_argument("--ncu-rounds", type=int, default=0,
                    help="profiling aid: after the warm-up run this many rounds launch by launch inside a "
                         "cudaProfilerStart/Stop range (ncu --profile-from-start off) and exit")
    ap.add_argument("--cache-log2", type=int, default=22,
                    help="log2 entries of the device evaluation cache (reference train.py:408-414); 0 = off")
    return ap.parse_args()


def ncu_dram_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused residual-block kernel (three blocks per
    launch) from the committed steady-state `ncu --set full` capture (profiles/r2g_net_kernels_ncu_full_raw.csv,
    Othello 8x8, 4096 games, evaluation cache on: ~3500 live leaves per launch), averaged over the captured
    launches; None if absent."""
    import csv
    path = ROOT / "profiles" / "r2g_net_kernels_ncu_full_raw.csv"
    if not path.exists():
        path = ROOT / "profiles" / "r2c_net_kernels_ncu_full_raw.csv"
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        name = hdr.index("Kernel Name")
        data = [r for r in rows[2:] if "k_resblock_tc" in r[name]]
        total = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(key)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            total += sum(float(r[i]) for r in data) * scale / len(data)
        return total
    except Exception:
        return None


def tree_roofline(pd, tree_ms, peaks):
    """HBM roofline of the tree kernel(s) from the run's own counters (SURVEY 8(d) byte model with
    f64 priors: child record visits 4 + value_sum 8 + prior 8 = 20 B, node header 16 B, path entry 4 B,
    state 17 B; expansion writes 26 B of node record + 17 B of stored state per child and reads 4*65 + 4 B of
    evaluator output; backup is a 12 B read + 12 B write per path node; plus the evaluation cache's probes, row
    copies and inserts).  A single pool runs both phases in one kernel (k_round)."""
    sims, depth, scanned = pd["simulations"], pd["depth_sum"], pd["children_scanned"]
    created, evals = pd["children_created"], pd["leaf_evals"]
    hits, inserts = pd.get("cache_hits", 0), pd.get("cache_inserts", 0)
    # selection: child records + headers along the path, the root state, the leaf's stored state and finality (17 B),
    # the path; evaluation-cache probe of every non-terminal leaf: 8 ways x (ctl 8 B + key 16 B)
    select_bytes = 16 * depth + 20 * scanned + 17 * sims + 17 * sims + 4 * (depth + sims) + (evals + hits) * 8 * 24
    # expansion: evaluator row in, 26 B node record + 17 B stored state per child out, backup RMW along the path;
    # cache: a hit copies a 272 B row out and back in, an insert writes a 272 B row + 24 B header
    expand_bytes = (evals + hits) * (16 + 4 * 65 + 4) + 43 * created + 24 * (depth + sims) + hits * 272 + inserts * 296
    peak = float(peaks.get("hbm_gbs", 6650.0)) if peaks else 6650.0
    out = {"bound": "hbm", "peak": peak, "unit": "GB/s",
           "note": "pointer-chasing kernels: dependent-load latency bound, far below the bandwidth roofline by construction"}
    for name, nbytes in (("select", select_bytes), ("expand_backup", expand_bytes),
                         ("tree_round", select_bytes + expand_bytes)):
        ms = tree_ms.get(name, 0.0)
        if ms <= 0:
            continue
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"algorithmic_bytes": nbytes, "ms": ms, "achieved": gbs, "frac": gbs / peak}
    return out


def workload_config(args):
    game, grid, sims, cutoff, alpha = WORKLOADS[args.workload]
    return {"game": game, "grid": grid, "num_simulations": sims, "sample_move_cutoff": cutoff,
            "dirichlet_alpha": alpha, "dirichlet_factor": 0.25, "weight_seed": 0}


def config_block(args, wl, n_gpus):
    return {
        "workload": f"{wl['game']} {wl['grid']}x{wl['grid']} self-play MCTS, {args.trees} concurrent games/GPU, "
                    f"Convolutional{wl['game'].title().replace('_', '')}Net({wl['grid']}) C=128 R=3 random-init, "
                    f"{wl['num_simulations']} sims/move, cutoff {wl['sample_move_cutoff']}, "
                    f"dirichlet {wl['dirichlet_alpha']}",
        "step": f"{MOVES_PER_STEP} moves of every game = {MOVES_PER_STEP * wl['num_simulations']} search rounds over the pool "
                "(one simulation per tree per round)",
        "eval_cache": f"device evaluation cache 2^{args.cache_log2} entries (reference train.py:408-414)" if args.cache_log2 > 0 else "off",
        "trees_per_gpu": args.trees, "pools_per_gpu": args.pools, "cuda_graph": bool(args.graph),
        "num_simulations": wl["num_simulations"],
        "parallelism": f"games sharded x{n_gpus}",
        "l2_policy": "inputs_larger_than_L2 (per-round activations 2x170 MB + GB-scale tree arenas vs 126 MB L2)",
    }


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """Times the oracle port of the reference's CPU self-play (the reference is pure Python and
    cannot travel to the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_selfplay  # noqa: the one place bench.py may execute oracle/
    wl = workload_config(args)
    procs = cpu_selfplay.default_procs()
    # a "step" of this arm is a bounded slice of CPU self-play; the whole run stays near a minute whatever K + W is
    slice_s = min(4.0, max(0.5, 60.0 / (args.warmup + args.steps)))
    slices = cpu_selfplay.run(wl, procs, args.warmup + args.steps, slice_s)
    timed = slices[args.warmup:]
    sims = sum(s[0] for s in timed)
    evals = sum(s[1] for s in timed)
    wall = sum(s[2] for s in timed)
    value = sims / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, len(timed)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args, wl, args.gpus),
        "leaf_evals_per_sec": evals / wall,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{procs} processes x 1 torch thread, {len(timed)} slices of {slice_s:.1f} s "
                                   "of full self-play games (oracle port of the reference loop)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit")

    def __init__(self, gpu_index):
        self.samples = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append((time.time(), parts))

    def mark(self):
        """Start of the timed region: samples taken before it (nvidia-smi is started during the warm-up so
        that it is already producing lines) are ignored unless nothing arrives inside the region."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        t0 = getattr(self, "t0", 0.0)
        inside = [parts for t, parts in self.samples if t >= t0]
        # a region shorter than the sampling period may catch no line: fall back to the last warm-up lines
        self.samples = inside if inside else [parts for _, parts in self.samples[-3:]]
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        def floats(col):
            out = []
            for s in self.samples:
                try:
                    out.append(float(s[col]))
                except (IndexError, ValueError):
                    pass
            return sorted(out)
        watts, limit = floats(6), floats(7)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm),
                # board power under load against its limit: the trunk is energy-bound
                "power_w": watts[len(watts) // 2] if watts else None, "power_limit_w": limit[-1] if limit else None}


# --------------------------------------------------------------------------------- b200 arm
class EnergyMeter:
    """Board energy from NVML's total-energy counter (millijoules since driver load) around a region."""

    def __init__(self, index):
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.read()
        except Exception:
            self.handle = None

    def read(self):
        if self.handle is None:
            return None
        try:
            return self.nvml.nvmlDeviceGetTotalEnergyConsumption(self.handle) / 1000.0
        except Exception:
            self.handle = None
            return None


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from deep_mcts_b200 import _lib
    from deep_mcts_b200.gamenet import cached_state_evaluator
    from deep_mcts_b200.mcts import BatchedMCTS
    from deep_mcts_b200.selfplay import BatchedSelfPlay

    wl = workload_config(args)
    torch.manual_seed(wl["weight_seed"])
    if wl["game"] == "othello":
        from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
        net = ConvolutionalOthelloNet(wl["grid"])
    else:
        from deep_mcts_b200.hex.convolutionalnet import ConvolutionalHexNet
        net = ConvolutionalHexNet(wl["grid"])
    net.precision = args.precision
    net.max_batch = args.trees
    sims, T = wl["num_simulations"], args.trees
    rounds_per_step = MOVES_PER_STEP * sims
    engine = BatchedSelfPlay(net, T, sims, wl["sample_move_cutoff"], wl["dirichlet_alpha"], wl["dirichlet_factor"],
                             replay_capacity=1 << 20, seed=1000 + rank, n_pools=args.pools,
                             pool_sizes=[int(x) for x in args.pool_sizes.split(",")] if args.pool_sizes else None,
                             eval_cache_log2=args.cache_log2 if args.cache_log2 > 0 else None,
                             max_nodes_per_tree=args.max_nodes or None)
    engine.use_graph = bool(args.graph)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: W steps of 10 moves each play most of a first game, so the timed region sees games at every
    # stage (and the round graph is captured)
    clocks = ClockSampler(local_rank)
    energy = EnergyMeter(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        engine.run_rounds(rounds_per_step)
    barrier()
    if args.ncu_rounds > 0:
        engine.use_graph = False
        engine.run_rounds(2)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        engine.run_rounds(args.ncu_rounds)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_rounds": args.ncu_rounds, "note": "profiling run, no measurement"}))
        return
    c0 = engine.counters()
    launches0 = lib.dm_launch_count() + engine.graph_launches()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.mark()
    j0 = energy.read()
    start.record()
    for _ in range(args.steps):
        engine.run_rounds(rounds_per_step)
    stop.record()
    barrier()
    j1 = energy.read()
    elapsed_ms = start.elapsed_time(stop)
    clock_info = clocks.stop() if rank == 0 else None
    launches = lib.dm_launch_count() + engine.graph_launches() - launches0
    c1 = engine.counters()
    d = {k: c1[k] - c0[k] for k in c1}

    # ---- per-kernel CUDA-event timing for the roofline.  Events cannot be timed inside a replayed
    # graph, so the same steps are run once more launch by launch (identical kernels, grids and
    # data sizes) with an event pair around every network kernel on its launch stream.
    was_graph = engine.use_graph
    engine.use_graph = False
    for pdn in engine.device_nets:
        pdn.profile(True)
    p0 = engine.counters()
    engine.tree_timers = [] if args.pools == 1 else None
    pstart, pstop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_steps = max(1, min(args.steps, args.profile_steps))
    pstart.record()
    for _ in range(prof_steps):
        engine.run_rounds(rounds_per_step)
    pstop.record()
    torch.cuda.synchronize()
    prof_elapsed_ms = pstart.elapsed_time(pstop)
    prof = {}
    for pdn in engine.device_nets:
        for k, (ms, cnt) in pdn.profile_read().items():
            prof[k] = (prof.get(k, (0.0, 0))[0] + ms, prof.get(k, (0.0, 0))[1] + cnt)
        pdn.profile(False)
    p1 = engine.counters()
    tree_ms = {}
    for name, a, b in (engine.tree_timers or []):
        tree_ms[name] = tree_ms.get(name, 0.0) + a.elapsed_time(b)
    engine.tree_timers = None
    pd = {k: p1[k] - p0[k] for k in p1}
    prof_leaf_evals = p1["leaf_evals"] - p0["leaf_evals"]
    engine.use_graph = was_graph

    # ---- end to end through the public API with HOST buffers: host-driven self-play.  Every step is one move of T
    # games played the way the reference's MCTS.self_play does it (mcts.py:96-111), batched: the games start from
    # mid-game states uploaded from pinned HOST memory; step() runs the simulations with the reference's cached
    # evaluator (cached_state_evaluator(net), train.py:408-414) and returns visit counts and statistics as host
    # arrays, the host picks the moves (argmax) into a pinned buffer, advance() copies them in and re-roots every
    # tree with subtree reuse, states() brings the successor states back to the host, finished games are reset.
    parts = [pl.states() for pl in engine.pools]
    p, f, s, fin = (np.concatenate([part[i] for part in parts]) for i in range(4))
    idx = np.resize(np.where(~fin)[0], T)
    h_first = torch.from_numpy(f[idx].view(np.int64).copy()).pin_memory()
    h_second = torch.from_numpy(s[idx].view(np.int64).copy()).pin_memory()
    h_player = torch.from_numpy(p[idx].copy()).pin_memory()
    searcher = BatchedMCTS(net.manager, T, sims, None, cached_state_evaluator(net), seed=7,
                           dirichlet_alpha=wl["dirichlet_alpha"], dirichlet_factor=wl["dirichlet_factor"])
    searcher.pool.set_root_states(h_player.numpy(), h_first.numpy().view(np.uint64), h_second.numpy().view(np.uint64))
    h_actions = torch.zeros(T, dtype=torch.int32).pin_memory()
    h_ply = np.full(T, 1 << 20, dtype=np.int64)      # the uploaded games are past their opening; restarted games count from 0
    host_rng = np.random.RandomState(2024 + rank)

    def e2e_step():
        pool = searcher.pool
        visits, stats = searcher.step()                           # D2H: visit counts + statistics
        # the host chooses the moves like MCTS.self_play (mcts.py:99-104): sampled in proportion to the visit counts
        # for the first sample_move_cutoff plies of a game, argmax (first maximum) afterwards
        actions = torch.from_numpy(visits).argmax(dim=1).numpy()   # first maximum, like np.argmax; uses the host's cores
        early = np.where(h_ply < wl["sample_move_cutoff"])[0]
        if early.size:
            cdf = np.cumsum(visits[early].astype(np.float64), axis=1)
            draw = host_rng.random_sample(early.size) * cdf[:, -1]
            actions[early] = (cdf > draw[:, None]).argmax(axis=1)
        h_actions.numpy()[:] = actions
        searcher.advance(h_actions.numpy())                       # H2D: chosen actions; re-root with subtree reuse
        h_ply[:] += 1
        np_p, np_f, np_s, np_fin, _ = pool.states()               # D2H: successor states, final flags
        if np_fin.any():
            done = np.where(np_fin)[0].astype(np.int32)
            pool.reset(done)                                      # H2D: finished games start over
            h_ply[done] = 0
        return int(stats.sum())

    for _ in range(3):
        e2e_step()
    barrier()
    ce0 = searcher.pool.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_launch0 = lib.dm_launch_count()
    e0.record()
    e2e_sims = 0
    for _ in range(args.e2e_steps):
        e2e_sims += e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    h2d = T * 4                                              # chosen actions (+ the ids of finished games)
    d2h = T * (net.manager.num_actions * 4 + 8 + 17 + 5)     # visit counts + stats, successor states + final flag/outcome
    ce1 = searcher.pool.counters()
    ce = {k: ce1[k] - ce0[k] for k in ce1}

    # ---- reduce over ranks: sums of work, max of time
    from deep_mcts_b200.distributed import reduce_counters
    joules = (j1 - j0) if (j0 is not None and j1 is not None) else 0.0
    sums, maxima = reduce_counters(
        {"sims": d["simulations"], "evals": d["leaf_evals"], "e2e_sims": e2e_sims, "launches": launches, "joules": joules},
        {"elapsed_ms": elapsed_ms, "e2e_ms": e2e_ms}, "cuda")
    tot_sims, tot_evals, tot_e2e_sims, tot_launches = sums["sims"], sums["evals"], sums["e2e_sims"], sums["launches"]
    elapsed_ms, e2e_ms = maxima["elapsed_ms"], maxima["e2e_ms"]

    if rank == 0:
        g = wl["grid"]
        flop_per_leaf_conv = 2.0 * g * g * 9 * 128 * 128  # 18.87 MFLOP at g=8, 14.45 at g=7
        conv_ms, conv_launches = prof["trunk_conv"]
        peaks = {}
        burst, sustained = 1590.0, 1590.0
        peak_src = "fallback (B200_PROFILING.md: 1.59 PFLOP/s burst, 6.65 TB/s)"
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
            burst, sustained = float(peaks["bf16_tflops"]), float(peaks["bf16_tflops_sustained"])
            peak_src = "MEASURED_PEAKS.json"
        except Exception:
            peaks = {}
        # the kernel is timed inside a loop that has kept the GPU busy for load_s seconds: sustained figure from 2 s on
        load_s = (elapsed_ms + prof_elapsed_ms) * 1e-3
        use_sustained = load_s >= 2.0
        peak_tflops = sustained if use_sustained else burst
        achieved = (prof_leaf_evals * 7 * flop_per_leaf_conv) / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        tree = tree_roofline(pd, tree_ms, peaks)
        total_j = sums["joules"]
        line = {
            "metric": METRIC, "value": tot_sims / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": config_block(args, wl, world),
            "leaf_evals_per_sec": tot_evals / (elapsed_ms * 1e-3),
            "e2e": {"value": tot_e2e_sims / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                    "leaf_evals_per_step": ce["leaf_evals"] / max(1, args.e2e_steps),
                    "cache_served_fraction": (ce["cache_hits"] + ce["cache_aliases"]) /
                                             max(1, ce["cache_hits"] + ce["cache_aliases"] + ce["leaf_evals"]),
                    "gpu_launches_direct": int(lib.dm_launch_count() - e2e_launch0),
                    "what": "host-driven self-play through the public batched MCTS API, one move of every game per step "
                            f"(the reference's MCTS.self_play loop, mcts.py:96-111): BatchedMCTS.step(): {sims} simulations "
                            "per tree with cached_state_evaluator(net), visit counts + statistics to host arrays (D2H) -> "
                            "host move choice into pinned memory (sampled from the visits for the first plies of a game, else argmax) "
                            "-> advance(actions) (H2D, re-root with subtree reuse) -> "
                            "states() (D2H); finished games restart"},
            "gpu_launches": int(tot_launches),
            "clocks": clock_info,
            "energy": {"joules_per_step": total_j / args.steps if total_j else None,
                       "joules_per_million_sims": total_j / (tot_sims / 1e6) if total_j else None,
                       "joules_per_tflop_useful": total_j / (tot_evals * 7 * flop_per_leaf_conv / 1e12) if total_j and tot_evals else None,
                       "mean_watts_per_gpu": total_j / world / (elapsed_ms * 1e-3) if total_j else None,
                       "source": "NVML total energy counter around the timed region"},
            "roofline": {
                "kernel": "k_resblock_tc<G> x1 + k_conv3x3_tc2 x1 per round (tcgen05 implicit-GEMM 3x3 convs: the three fused "
                          "residual blocks in one launch and the policy conv; seven 128->128 convolutions in two launches)",
                "bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                "frac": achieved / peak_tflops if peak_tflops else None,
                "frac_of_burst": achieved / burst, "frac_of_sustained": achieved / sustained,
                "traffic": ncu_dram_traffic_per_launch() if wl["game"] == "othello" and args.trees == 4096 else None,
                "traffic_unit": "DRAM bytes per k_resblock_tc<8> launch = three residual blocks (ncu --set full capture "
                                "under profiles/; the blocks' intermediate activations are re-read from the 126 MB L2)",
                # per fused residual block: the block input is read once (it is also the skip operand) and the
                # output written once, 128 bf16 channels per cell, plus two 9x128x128 bf16 weight sets; a launch
                # runs the trunk's three blocks over the leaves the evaluation cache did not serve
                "algorithmic_bytes_per_launch": 3 * (2 * (prof_leaf_evals / max(1, prof_steps * rounds_per_step)) * g * g * 256
                                                     + 2 * 9 * 128 * 128 * 2),
                "peak_source": f"{peak_src} bf16_tflops{'_sustained' if use_sustained else ''}: the kernels are timed after "
                               f"{load_s:.1f} s of continuous load (burst below 2 s, sustained from 2 s on)",
                "launches": conv_launches,
                "avg_launch_ms": conv_ms / conv_launches if conv_launches else None,
                "algorithmic_flop_per_launch": prof_leaf_evals * 7 * flop_per_leaf_conv / max(1, conv_launches),
                "share_of_step": conv_ms / prof_elapsed_ms if prof_elapsed_ms else None,
                "timed_in": f"{prof_steps} extra steps ({prof_steps * rounds_per_step} rounds) launched kernel by kernel right "
                            "after the timed region (CUDA events cannot be timed inside the replayed graph)",
                # second entry: the tree kernel (one launch per round), HBM-bound by construction
                "tree": tree,
            },
            "eval_cache": {"log2_entries": args.cache_log2, "hits": d["cache_hits"], "aliases": d["cache_aliases"],
                           "inserts": d["cache_inserts"], "leaf_evals": d["leaf_evals"],
                           "served_fraction": (d["cache_hits"] + d["cache_aliases"]) /
                                              max(1, d["cache_hits"] + d["cache_aliases"] + d["leaf_evals"])},
            "kernel_ms": {**{k: v[0] for k, v in prof.items()}, **tree_ms},
            "tree_stats": {"mean_depth": d["depth_sum"] / max(1, d["simulations"]),
                           "children_scanned_per_sim": d["children_scanned"] / max(1, d["simulations"]),
                           "terminal_fraction": d["terminal_sims"] / max(1, d["simulations"]),
                           "games_finished": d["games"], "moves": d["moves"],
                           "nodes_high_water": c1["nodes_high_water"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_selfplay  # noqa: cpu_baseline leg only
            procs = cpu_selfplay.default_procs()
            res = cpu_selfplay.run(wl, procs, 1, args.cpu_seconds)
            cs, ce_, cw = res[0]
            # BASELINE.md section 3 asks for the one-process figure next to the all-cores one
            one = cpu_selfplay.run({**wl, "torch_threads": procs}, 1, 1, args.cpu_seconds)[0]
            line["cpu_baseline"] = {
                "value": cs / cw, "unit": UNIT, "cores": procs, "kind": "port",
                "leaf_evals_per_sec": ce_ / cw,
                "sample": f"{procs} processes x 1 torch thread x {args.cpu_seconds:.0f} s of full self-play games "
                          "with the same net/config (oracle port of the reference loop; the Python reference "
                          "itself is not present on the GPU box)",
                "single_process": {"value": one[0] / one[2], "unit": UNIT, "torch_threads": procs,
                                   "leaf_evals_per_sec": one[1] / one[2],
                                   "sample": f"1 process, torch intra-op threads = {procs}, {args.cpu_seconds:.0f} s"},
                "calibration": "in the build container the port runs 1.19x faster per process than the unmodified "
                               "reference (331 vs 278 sims/s), see DESIGN.md section 6"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
