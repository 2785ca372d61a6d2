#!/usr/bin/env python
"""Which trees set the duration of the tree kernel?  Runs steady-state self-play rounds (CUDA graph off) with the
per-warp timing of dm_pool_round switched on and prints, per workload, the distribution of per-warp cycles and what
the slowest warps of each launch were doing (a move, a game end, terminal re-descents, deep paths).

    python scripts/tree_tail_probe.py [othello8|hex7] > profiles/tree_tail_probe_r2.jsonl
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from deep_mcts_b200 import _lib
    from deep_mcts_b200.pool import wrap_device_buffer
    from deep_mcts_b200.selfplay import BatchedSelfPlay
    workloads = sys.argv[1:] or ["othello8", "hex7"]
    torch.cuda.set_device(0)
    for wl in workloads:
        torch.manual_seed(0)
        if wl == "othello8":
            from deep_mcts_b200.othello.convolutionalnet import ConvolutionalOthelloNet
            net, alpha = ConvolutionalOthelloNet(8), 1.0
        else:
            from deep_mcts_b200.hex.convolutionalnet import ConvolutionalHexNet
            net, alpha = ConvolutionalHexNet(7), 0.33
        net.precision = "bf16"
        net.max_batch = 4096
        engine = BatchedSelfPlay(net, 4096, 50, 10, alpha, 0.25, replay_capacity=1 << 20, seed=1)
        engine.use_graph = True
        engine.run_rounds(3000)
        engine.use_graph = False
        pool = engine.pool
        ptr = _lib.c_void_p()
        _lib.call("dm_pool_debug_timing", pool._h, ptr)
        dbg = wrap_device_buffer(ptr.value, (4096, 4), torch.int64, pool.device)
        engine.run_rounds(3)
        torch.cuda.synchronize()
        rows = []
        for _ in range(150):
            engine.run_rounds(1)
            torch.cuda.synchronize()
            rows.append(dbg.cpu().numpy().copy())
        a = np.stack(rows)                      # [launch, tree, 4]
        total = (a[:, :, 0] + a[:, :, 1]).astype(np.float64)
        flags = a[:, :, 2]
        moves, games = (flags >> 32) & 0xffffffff, (flags >> 24) & 0xff
        terminal, levels = (flags >> 16) & 0xff, flags & 0xffff
        slowest = total.argmax(axis=1)
        idx = (np.arange(total.shape[0]), slowest)
        out = {
            "workload": wl, "launches": int(total.shape[0]),
            "warp_cycles_percentiles": {str(p): float(np.percentile(total, p)) for p in (50, 90, 99, 99.9, 100)},
            "mean_of_launch_max_cycles": float(total.max(axis=1).mean()),
            "mean_expand_cycles": float(a[:, :, 0].mean()), "mean_select_cycles": float(a[:, :, 1].mean()),
            "slowest_warp_of_a_launch": {
                "made_a_move_fraction": float((moves[idx] > 0).mean()),
                "ended_a_game_fraction": float((games[idx] > 0).mean()),
                "had_terminal_simulations_fraction": float((terminal[idx] > 0).mean()),
                "mean_levels_walked": float(levels[idx].mean()),
                "mean_expand_cycles": float(a[:, :, 0][idx].mean()), "mean_select_cycles": float(a[:, :, 1][idx].mean())},
            "all_warps": {"made_a_move_fraction": float((moves > 0).mean()), "ended_a_game_fraction": float((games > 0).mean()),
                          "had_terminal_simulations_fraction": float((terminal > 0).mean()),
                          "mean_levels_walked": float(levels.mean())},
            "cycles_by_class": {
                "plain_round": float(total[(moves == 0) & (terminal == 0)].mean()),
                "with_terminal": float(total[(moves == 0) & (terminal > 0)].mean()) if ((moves == 0) & (terminal > 0)).any() else None,
                "move": float(total[(moves > 0) & (games == 0)].mean()) if ((moves > 0) & (games == 0)).any() else None,
                "game_end": float(total[games > 0].mean()) if (games > 0).any() else None},
        }
        print(json.dumps(out), flush=True)
        del engine


if __name__ == "__main__":
    main()
