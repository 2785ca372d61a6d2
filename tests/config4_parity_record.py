#!/usr/bin/env python
"""Config 4 (BASELINE.json configs[3]): complex (greedy policy-net rollouts) vs simple (random rollouts) agents, with and
without the state evaluator, on the device (batched tournament) next to the committed oracle games
(tests/golden/config4_<case>.json).  Prints one JSON line per pairing with both win rates and their 95 % intervals.

    python tests/config4_parity_record.py [case ...] > profiles/config4_parity_r2.jsonl
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def wilson(k, n, z=1.96):
    p = k / n
    centre = (p + z * z / (2 * n)) / (1 + z * z / n)
    half = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / (1 + z * z / n)
    return [round(centre - half, 4), round(centre + half, 4)]


def main():
    from test_config4_gpu import run_device_pairing, within_binomial_bounds
    cases = sys.argv[1:] or ["small7", "full6"]
    for case in cases:
        data = json.load(open(ROOT / "tests" / "golden" / f"config4_{case}.json"))
        cfg = data["config"]
        for with_evaluator in (False, True):
            name = "with_state_evaluator" if with_evaluator else "without_state_evaluator"
            want = data["pairings"][name]
            for precision in ("fp32", "bf16") if cfg["channels"] == 128 else ("fp32",):
                t0 = time.time()
                got = run_device_pairing(cfg, with_evaluator, 240, seed=11, precision=precision)
                row = {"config": f"hex_with_swap {cfg['grid']}x{cfg['grid']}, C={cfg['channels']} R={cfg['num_residual']}, "
                                 f"{cfg['sims']} sims/move, 240 games, sides alternating", "pairing": name,
                       "device_precision": precision, "device_seconds": round(time.time() - t0, 1),
                       "oracle_cpu_seconds": round(want["cpu_seconds"], 1)}
                for side in ("first", "second"):
                    ok, p_dev, p_cpu, sigma = within_binomial_bounds(got[f"complex_wins_as_{side}"], got[f"games_as_{side}"],
                                                                     want[f"complex_wins_as_{side}"], want[f"games_as_{side}"], 3.0)
                    row[f"complex_win_rate_as_{side}"] = {
                        "device": round(p_dev, 4), "device_ci95": wilson(got[f"complex_wins_as_{side}"], got[f"games_as_{side}"]),
                        "oracle": round(p_cpu, 4), "oracle_ci95": wilson(want[f"complex_wins_as_{side}"], want[f"games_as_{side}"]),
                        "difference_in_sigmas": round(abs(p_dev - p_cpu) / sigma, 2), "within_3_sigma": bool(ok)}
                print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
