"""Config 4 (BASELINE.json configs[3], reference evaluate_complex_rollouts.py:53-90): the "complex" agent (MCTS whose
rollout policy is argmax of the policy network) against the "simple" agent (uniform random rollouts), with and
without the network as state evaluator, rollout_share = 1, alternating sides -- played on the ORACLE with a fixed
simulation budget and recorded game by game:

    python tests/golden/make_config4_golden.py <case> [procs]       (CPU hours for the large cases)

Cases (the network is oracle/synth.py's seeded synthetic state_dict, shared with the GPU test):
  small7   hex_with_swap 7x7, C=32 R=1 H=32,  50 sims, 240 games per pairing   (tests/test_config4_gpu.py)
  full6    hex_with_swap 6x6 (the reference script's size), C=128 R=3 H=128, 50 sims, 240 games per pairing
Output: tests/golden/config4_<case>.json with the outcome value of every game (0 FIRST wins, 1 SECOND wins), the agent
that moved first, and the complex agent's (wins, losses) as first and as second mover.
"""
import json
import multiprocessing as mp
import random
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

CASES = {
    "small7": dict(grid=7, channels=32, num_residual=1, hidden=32, sims=50, games=240, seed=41),
    "full6": dict(grid=6, channels=128, num_residual=3, hidden=128, sims=50, games=240, seed=31),
}


def play_games(args):
    case, with_evaluator, indices = args
    import torch
    torch.set_num_threads(1)
    from oracle import net as onet
    from oracle.games import make_game
    from oracle.mcts import OracleMCTS
    from oracle.synth import synthetic_state_dict
    cfg = CASES[case]
    game = make_game("hex_with_swap", cfg["grid"])
    sd = synthetic_state_dict(game, cfg["channels"], cfg["num_residual"], cfg["hidden"], seed=cfg["seed"],
                              conv_gain=1.0, fc_scale=1.0)
    out = []
    for k in indices:
        # fresh evaluator caches per game, like the reference's reset_caches (evaluate_complex_rollouts.py:103-109)
        ev_complex = onet.make_evaluator(sd)
        ev_simple = onet.make_evaluator(sd)
        rng = random.Random(9000 + k)
        complex_agent = OracleMCTS(game, cfg["sims"], lambda g, s: int(np.argmax(ev_complex(g, s)[1])),
                                   ev_complex if with_evaluator else None, rollout_share=1.0)
        simple_agent = OracleMCTS(game, cfg["sims"], lambda g, s: rng.choice(g.legal_list(s)),
                                  ev_simple if with_evaluator else None, rollout_share=1.0)
        agents = (complex_agent, simple_agent)
        first_mover = k % 2                      # tournament.py:66-76: sides alternate
        s = game.initial_state()
        ply = 0
        while not game.is_final(s):
            a = first_mover if ply % 2 == 0 else 1 - first_mover
            m = agents[a]
            m.observe(s)
            dist, _ = m.step()
            action = int(np.argmax(dist))
            m.advance(action)
            s = game.child(s, action)
            ply += 1
        out.append((k, first_mover, game.outcome(s), ply))
    return out


def main():
    case = sys.argv[1]
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    cfg = CASES[case]
    result = {"case": case, "config": cfg, "game": "hex_with_swap", "pairings": {}}
    with mp.Pool(procs) as pool:
        for with_evaluator in (False, True):
            t0 = time.time()
            chunks = [list(range(r, cfg["games"], procs * 4)) for r in range(procs * 4)]
            rows = sorted(sum(pool.map(play_games, [(case, with_evaluator, c) for c in chunks if c]), []))
            first_mover = [r[1] for r in rows]
            outcome = [r[2] for r in rows]
            # the complex agent is agent 0: as FIRST it wins at outcome 0, as SECOND at outcome 1
            as_first = [o for f, o in zip(first_mover, outcome) if f == 0]
            as_second = [o for f, o in zip(first_mover, outcome) if f == 1]
            name = "with_state_evaluator" if with_evaluator else "without_state_evaluator"
            result["pairings"][name] = {
                "first_mover": first_mover, "outcome": outcome, "plies": [r[3] for r in rows],
                "complex_wins_as_first": sum(o == 0.0 for o in as_first), "games_as_first": len(as_first),
                "complex_wins_as_second": sum(o == 1.0 for o in as_second), "games_as_second": len(as_second),
                "cpu_seconds": (time.time() - t0) * procs}
            print(case, name, {k: v for k, v in result["pairings"][name].items() if not isinstance(v, list)}, flush=True)
            json.dump(result, open(HERE / f"config4_{case}.json", "w"))


if __name__ == "__main__":
    main()
