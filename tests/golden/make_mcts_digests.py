"""Large MCTS pins as digests: every move's root visit counts, (with, without) statistics, child order,
root visit count and chosen action of several hundred full games of the UNMODIFIED reference MCTS
(deterministic hashed evaluator of oracle/evaluators.py, subtree reuse between moves, forced random
openings) -- SURVEY section 8(c), pin 3.

Run in the build container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_mcts_digests.py

``mcts_digests.json`` keeps one SHA-256 per game.  The GPU test replays all games of a plan entry at
once (one tree per game) and must reproduce every digest; no oracle is involved at test time.
"""
import hashlib
import json
import random
import struct
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
# (game, grid, simulations per move, games)
PLAN = [("tictactoe", 3, 25, 128), ("hex", 7, 50, 128), ("hex_with_swap", 7, 50, 96), ("othello", 8, 50, 128),
        ("othello", 8, 200, 16), ("hex", 7, 200, 16)]
MAX_OPENING_PLIES = 6


def game_seed(plan_index: int, k: int) -> int:
    return 100_000 * (plan_index + 1) + k


def absorb_move(h, visits, stats, order, root_visits, action):
    h.update(struct.pack(f"<{len(visits)}H", *visits))
    h.update(struct.pack("<HHI", stats[0], stats[1], root_visits))
    h.update(bytes(order) + bytes([action]))


def choose(rng, ply, opening_plies, order, visits):
    """Forced random opening among the root's children (sorted), then argmax of the visit counts
    (lowest action on ties, like np.argmax)."""
    if ply < opening_plies:
        return rng.choice(sorted(order))
    best = max(visits)
    return visits.index(best)


def main():
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, str(HERE))
    sys.dont_write_bytecode = True
    from deep_mcts.mcts import MCTS
    from make_golden import manager_for, ref_evaluator
    out = []
    for pi, (name, g, sims, n_games) in enumerate(PLAN):
        manager = manager_for(name, g)
        evaluator = ref_evaluator("hashed", name, g)
        digests, moves_total = [], 0
        for k in range(n_games):
            rng = random.Random(game_seed(pi, k))
            opening_plies = k % (MAX_OPENING_PLIES + 1)
            mcts = MCTS(manager, sims, None, evaluator)
            h = hashlib.sha256()
            ply = 0
            while not manager.is_final_state(mcts.state):
                _, stats = mcts.step()
                children = mcts.root.children
                visits = [children[a].visits if a in children else 0 for a in range(manager.num_actions)]
                order = list(children.keys())
                action = choose(rng, ply, opening_plies, order, visits)
                absorb_move(h, visits, stats, order, mcts.root.visits, action)
                mcts.root = children[action]
                mcts.state = manager.generate_child_state(mcts.state, action)
                ply += 1
            h.update(struct.pack("<f", manager.evaluate_final_state(mcts.state).value))
            digests.append(h.hexdigest())
            moves_total += ply
        out.append({"game": name, "grid": g, "sims": sims, "games": n_games, "moves": moves_total, "digests": digests})
        print(name, g, sims, n_games, "games", moves_total, "moves", flush=True)
    with open(HERE / "mcts_digests.json", "w") as f:
        json.dump({"max_opening_plies": MAX_OPENING_PLIES, "plan": out}, f, indent=0)


if __name__ == "__main__":
    main()
