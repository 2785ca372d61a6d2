"""Win/draw/loss counts of uniform random rollouts at the BASELINE board sizes, from the oracle AND from the
unmodified reference (run in the build container only; imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_rollout_golden.py

For every start position (initial state, seeded mid-game positions, and for Hex-with-swap the state in which the
swap is legal) it plays
  * ORACLE_N playouts of the oracle port of ``MCTS.rollout`` with ``random.choice(legal_actions(s))``
    (oracle/games.py; reference mcts.py:219-226), and
  * REFERENCE_N playouts of the reference's own managers with the same policy,
and writes tests/golden/rollouts.json: {game: [{state, oracle: [first_wins, draws, second_wins], reference: [...]}]}.
The GPU test compares the CUDA playouts' rates with both within binomial bounds.
"""
import json
import multiprocessing as mp
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

ORACLE_N = 20000
REFERENCE_N = 2000
CHUNKS = 8
GAMES = [("othello", 8, (6, 14, 30)), ("hex", 7, (5, 12, 22)), ("hex_with_swap", 7, (1, 6, 15))]


def start_positions(name, g, plies):
    from oracle.games import make_game
    game = make_game(name, g)
    rng = random.Random(4242)
    out = [game.initial_state()]
    for n in plies:
        while True:
            s = game.initial_state()
            for _ in range(n):
                if game.is_final(s):
                    break
                s = game.child(s, rng.choice(game.legal_list(s)))
            if not game.is_final(s):
                out.append(s)
                break
    return out


def oracle_chunk(args):
    name, g, state, n, seed = args
    from oracle.games import make_game
    game = make_game(name, g)
    rng = random.Random(seed)
    counts = [0, 0, 0]
    for _ in range(n):
        s = tuple(state)
        while not game.is_final(s):
            s = game.child(s, rng.choice(game.legal_list(s)))
        counts[int(game.outcome(s) * 2)] += 1
    return counts


def reference_chunk(args):
    name, g, state, n, seed = args
    from deep_mcts.game import CellState, Player
    from make_golden import manager_for
    manager = manager_for(name, g)
    state_cls = type(manager.initial_game_state())
    player, first, second = state
    grid = tuple(tuple(CellState.FIRST_PLAYER if (first >> (y * g + x)) & 1 else
                       (CellState.SECOND_PLAYER if (second >> (y * g + x)) & 1 else CellState.EMPTY)
                       for x in range(g)) for y in range(g))
    start = state_cls(Player(player), grid)
    random.seed(seed)
    counts = [0, 0, 0]
    for _ in range(n):
        s = start
        while not manager.is_final_state(s):
            s = manager.generate_child_state(s, random.choice(manager.legal_actions(s)))
        counts[int(manager.evaluate_final_state(s).value * 2)] += 1
    return counts


def total(chunks):
    return [sum(c[i] for c in chunks) for i in range(3)]


def main():
    out = {}
    with mp.Pool(CHUNKS) as pool:
        for name, g, plies in GAMES:
            rows = []
            for k, state in enumerate(start_positions(name, g, plies)):
                oracle = total(pool.map(oracle_chunk, [(name, g, state, ORACLE_N // CHUNKS, 1000 * k + c) for c in range(CHUNKS)]))
                ref = total(pool.map(reference_chunk, [(name, g, state, REFERENCE_N // CHUNKS, 77 * k + c) for c in range(CHUNKS)]))
                rows.append({"state": [int(state[0]), int(state[1]), int(state[2])], "oracle": oracle, "reference": ref})
                print(name, g, state, oracle, ref, flush=True)
            out[f"{name}_{g}"] = rows
    json.dump({"oracle_n": ORACLE_N, "reference_n": REFERENCE_N, "counts": "[FIRST wins, draws, SECOND wins]", "games": out},
              open(HERE / "rollouts.json", "w"), indent=1)


if __name__ == "__main__":
    main()
