"""Large game-logic pins as digests: SHA-256 over 100 000 positions per BASELINE game of seeded random
playouts of the UNMODIFIED reference managers (pin 1 and 2).

Run in the build container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_digests.py

For every visited position the digest absorbs: the state, is_final, the outcome, ``legal_actions``
in the reference's list order, ``list(set(legal_actions))`` (the child order MCTS.expand_node
iterates) and every child state.  ``digests.json`` keeps the running digest at a few checkpoints so
that the CPU suite can check a prefix and the GPU suite the whole walk.  The walk itself (which move
is played) is shared with the tests through ``walk_digest`` below, which only needs five callables,
so the oracle and the CUDA engines are driven through exactly the same code.
"""
import hashlib
import json
import random
import struct
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CONFIGS = [("tictactoe", 3), ("hex", 7), ("hex_with_swap", 7), ("othello", 8)]
CHECKPOINTS = (5_000, 25_000, 100_000)
SEED = 20201


def absorb(h, state, final, outcome, legal, order, children):
    h.update(struct.pack("<BQQBB", state[0], state[1], state[2], 1 if final else 0, int(round(outcome * 2))))
    h.update(struct.pack("<H", len(legal)) + bytes(legal) + bytes(order))
    for c in children:
        h.update(struct.pack("<BQQ", c[0], c[1], c[2]))


def walk_digest(initial, legal_list, set_order, child, status, n_positions, checkpoints, seed=SEED):
    """``initial() -> state``, ``legal_list(state) -> [action]``, ``set_order(state) -> [action]``,
    ``child(state, action) -> state``, ``status(state) -> (is_final, outcome value)``; states are
    ``(player, first_bitboard, second_bitboard)``.  Returns {checkpoint: hex digest}."""
    rng = random.Random(seed)
    h = hashlib.sha256()
    out = {}
    count = 0
    while count < n_positions:
        state = initial()
        while count < n_positions:
            final, outcome = status(state)
            legal = list(legal_list(state))
            children = [child(state, a) for a in legal]
            absorb(h, state, final, outcome, legal, list(set_order(state)), children)
            count += 1
            if count in checkpoints:
                out[count] = h.hexdigest()
            if final:
                break
            choices = sorted(set(legal))
            state = child(state, choices[rng.randrange(len(choices))])
    return out


def main():
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, str(HERE))
    sys.dont_write_bytecode = True
    from make_golden import manager_for, to_bits   # reference managers + state -> bitboards
    result = {}
    for name, g in CONFIGS:
        manager = manager_for(name, g)
        back = {}

        def remember(state):
            bits = to_bits(state)
            back[bits] = state
            return bits

        digests = walk_digest(
            lambda: remember(manager.initial_game_state()),
            lambda b: manager.legal_actions(back[b]),
            lambda b: list(set(manager.legal_actions(back[b]))),
            lambda b, a: remember(manager.generate_child_state(back[b], a)),
            lambda b: (manager.is_final_state(back[b]), manager.evaluate_final_state(back[b]).value),
            max(CHECKPOINTS), CHECKPOINTS)
        result[f"{name}_{g}"] = {str(k): v for k, v in digests.items()}
        print(name, g, digests, flush=True)
        back.clear()
    with open(HERE / "digests.json", "w") as f:
        json.dump({"seed": SEED, "checkpoints": list(CHECKPOINTS), "digests": result}, f, indent=1)


if __name__ == "__main__":
    main()
