"""CPU self-play baseline: the oracle port of the reference's self-play loop timed on host
cores.  TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py): only bench.py's ``cpu_baseline``
and ``--impl reference`` legs run this.

Each worker process mirrors one reference self-play worker (reference deep_mcts/train.py:
228-283): batch-1 torch-fp32 forward per leaf behind the reference's memoisation
(train.py:408-414), ``torch.set_num_threads(1)``, full games with ``sample_move_cutoff`` and
Dirichlet noise.  Weights: the product's torch module constructed under
``torch.manual_seed(weight_seed)`` (random init, the BASELINE workload).
"""
import multiprocessing as mp
import os
import time
from typing import Dict, List, Tuple


def _build_state_dict(game_name: str, grid: int, weight_seed: int):
    import torch
    torch.manual_seed(weight_seed)
    if game_name == "othello":
        from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet as Net
        net = Net(grid)
    elif game_name == "hex":
        from deep_mcts_b200.hex.convolutionalnet import ConvolutionalHexNet as Net
        net = Net(grid)
    elif game_name == "hex_with_swap":
        from deep_mcts_b200.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet as Net
        net = Net(grid)
    else:
        from deep_mcts_b200.tictactoe.convolutionalnet import ConvolutionalTicTacToeNet as Net
        net = Net()
    net.net.eval()
    return {k: v.detach().clone() for k, v in net.net.state_dict().items()}


def _worker(rank: int, cfg: Dict, slices: int, seconds_per_slice: float, out_queue) -> None:
    import random

    import numpy as np
    import torch
    torch.set_num_threads(cfg.get("torch_threads", 1))
    from oracle import net as onet
    from oracle.games import make_game
    from oracle.mcts import OracleMCTS
    random.seed(rank)
    np.random.seed(rank)
    game = make_game(cfg["game"], cfg["grid"])
    sd = _build_state_dict(cfg["game"], cfg["grid"], cfg["weight_seed"])
    evaluator = onet.make_evaluator(sd)
    mcts = OracleMCTS(game, cfg["num_simulations"], None, evaluator, cfg["sample_move_cutoff"],
                      cfg["dirichlet_alpha"], cfg["dirichlet_factor"])
    ply = 0
    results = []
    for _ in range(slices):
        sims = 0
        misses0 = evaluator.misses
        start = time.perf_counter()
        while time.perf_counter() - start < seconds_per_slice:
            if game.is_final(mcts.state):
                mcts.reset()
                ply = 0
            dist, _ = mcts.step()
            sims += cfg["num_simulations"]
            if ply < cfg["sample_move_cutoff"]:
                action = int(np.random.choice(len(dist), p=dist))
            else:
                action = int(np.argmax(dist))
            mcts.advance(action)
            ply += 1
        results.append((sims, evaluator.misses - misses0, time.perf_counter() - start))
    out_queue.put((rank, results))


def run(cfg: Dict, procs: int, slices: int, seconds_per_slice: float) -> List[Tuple[int, int, float]]:
    """Per slice: (total simulations, total evaluator cache misses, max wall seconds over workers)."""
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    workers = [ctx.Process(target=_worker, args=(r, cfg, slices, seconds_per_slice, queue)) for r in range(procs)]
    for w in workers:
        w.start()
    gathered = []
    deadline = time.time() + slices * seconds_per_slice + 300
    try:
        while len(gathered) < len(workers):
            try:
                gathered.append(queue.get(timeout=5))
            except Exception:
                if time.time() > deadline or not any(w.is_alive() for w in workers):
                    if queue.empty():
                        raise RuntimeError("cpu_selfplay: a worker process died or timed out")
    finally:
        for w in workers:
            w.join(timeout=10)
            if w.is_alive():
                w.terminate()
    out = []
    for s in range(slices):
        sims = sum(res[s][0] for _, res in gathered)
        evals = sum(res[s][1] for _, res in gathered)
        wall = max(res[s][2] for _, res in gathered)
        out.append((sims, evals, wall))
    return out


def default_procs() -> int:
    return max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
