"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (it imports /root/reference, which does not
exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (committed):
  games_<cfg>.npz   positions from seeded random playouts of the reference manager:
                    legal_actions list order, list(set(legal)) order, is_final,
                    outcome, and every child state
  mcts.json         per-move root visit counts / stats / child order of the
                    reference MCTS and MCTSAgent with deterministic evaluators
  net_<cfg>.npz     GameNet.forward / GameNet.evaluate outputs of the reference
                    nets on synthetic weights (oracle/synth.py)
  META.json         interpreter / library versions the fixtures were made with
(make_digests.py and make_mcts_digests.py, run the same way, add digests.json -- 100 000 positions per
BASELINE game -- and mcts_digests.json -- 512 full reference MCTS games.)

The evaluators and synthetic weights are *inputs* shared with the tests
(oracle/evaluators.py, oracle/synth.py); every recorded OUTPUT comes from the
reference's own code.
"""
import json
import os
import random
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

from deep_mcts.game import Player  # noqa: E402
from deep_mcts.mcts import MCTS, MCTSAgent  # noqa: E402
from deep_mcts.tournament import play  # noqa: E402
from deep_mcts.hex.game import HexManager, HexState  # noqa: E402
from deep_mcts.hex_with_swap.game import HexWithSwapManager  # noqa: E402
from deep_mcts.othello.game import OthelloManager, OthelloState  # noqa: E402
from deep_mcts.tictactoe.game import TicTacToeManager, TicTacToeState  # noqa: E402
from deep_mcts.hex.convolutionalnet import ConvolutionalHexNet  # noqa: E402
from deep_mcts.hex_with_swap.convolutionalnet import ConvolutionalHexWithSwapNet  # noqa: E402
from deep_mcts.othello.convolutionalnet import ConvolutionalOthelloNet  # noqa: E402
from deep_mcts.tictactoe.convolutionalnet import ConvolutionalTicTacToeNet  # noqa: E402

from oracle import evaluators as oev  # noqa: E402
from oracle.games import make_game  # noqa: E402
from oracle.synth import synthetic_state_dict  # noqa: E402

CONFIGS = [
    ("tictactoe", 3), ("hex", 4), ("hex", 7), ("hex_with_swap", 5), ("hex_with_swap", 7),
    ("othello", 4), ("othello", 6), ("othello", 8),
]


def manager_for(name, g):
    if name == "tictactoe":
        return TicTacToeManager()
    if name == "hex":
        return HexManager(g)
    if name == "hex_with_swap":
        return HexWithSwapManager(g)
    return OthelloManager(g)


def to_bits(state):
    g = len(state.grid)
    first = second = 0
    for y in range(g):
        for x in range(g):
            c = int(state.grid[y][x])
            if c == 1:
                first |= 1 << (y * g + x)
            elif c == 0:
                second |= 1 << (y * g + x)
    return int(state.player), first, second


# ------------------------------------------------------------------ game logic
def make_games(name, g, n_positions, seed):
    manager = manager_for(name, g)
    rng = random.Random(seed)
    rows = []
    while len(rows) < n_positions:
        state = manager.initial_game_state()
        while True:
            final = manager.is_final_state(state)
            legal = list(manager.legal_actions(state))
            rows.append((state, legal, final))
            if final or len(rows) >= n_positions:
                break
            state = manager.generate_child_state(state, rng.choice(legal))
    player, first, second, is_final, outcome = [], [], [], [], []
    legal_flat, legal_off, order_flat = [], [0], []
    child_player, child_first, child_second = [], [], []
    for state, legal, final in rows:
        p, f, s = to_bits(state)
        player.append(p), first.append(f), second.append(s)
        is_final.append(final)
        outcome.append(manager.evaluate_final_state(state).value)
        legal_flat += legal
        order_flat += list(set(legal))  # what MCTS.expand_node iterates (mcts.py:200)
        legal_off.append(len(legal_flat))
        for a in legal:
            cp, cf, cs = to_bits(manager.generate_child_state(state, a))
            child_player.append(cp), child_first.append(cf), child_second.append(cs)
    np.savez_compressed(
        HERE / f"games_{name}_{g}.npz",
        player=np.array(player, np.uint8), first=np.array(first, np.uint64), second=np.array(second, np.uint64),
        is_final=np.array(is_final, np.uint8), outcome=np.array(outcome, np.float32),
        legal=np.array(legal_flat, np.int16), order=np.array(order_flat, np.int16),
        offsets=np.array(legal_off, np.int32),
        child_player=np.array(child_player, np.uint8), child_first=np.array(child_first, np.uint64),
        child_second=np.array(child_second, np.uint64),
    )
    print(f"games_{name}_{g}: {len(rows)} positions, {len(legal_flat)} edges")


# ------------------------------------------------------------------------ MCTS
def ref_evaluator(kind, name, g):
    og = make_game(name, g)
    fn = {"uniform": oev.uniform, "hashed": oev.hashed, "hashed7": oev.hashed_salted(7)}[kind]

    def evaluator(state):
        return fn(og, to_bits(state))

    return evaluator


def ref_rollout_policy(kind, manager):
    if kind == "min":
        return lambda state: min(manager.legal_actions(state))
    if kind == "max":
        return lambda state: max(manager.legal_actions(state))
    raise ValueError(kind)


def mcts_game(name, g, sims, evaluator_kind, rollout_kind, opening_plies, seed):
    manager = manager_for(name, g)
    evaluator = ref_evaluator(evaluator_kind, name, g) if evaluator_kind else None
    rollout = ref_rollout_policy(rollout_kind, manager) if rollout_kind else None
    mcts = MCTS(manager, sims, rollout, evaluator)
    rng = random.Random(seed)
    moves = []
    ply = 0
    while not manager.is_final_state(mcts.state):
        dist, stats = mcts.step()
        children = mcts.root.children
        visits = [children[a].visits if a in children else 0 for a in range(manager.num_actions)]
        if ply < opening_plies:
            action = rng.choice(sorted(children.keys()))
        else:
            action = int(np.argmax(dist))
        moves.append({
            "visits": visits, "stats": list(stats), "order": list(children.keys()),
            "root_visits": mcts.root.visits, "action": action,
        })
        mcts.root = children[action]
        mcts.state = manager.generate_child_state(mcts.state, action)
        ply += 1
    return {
        "game": name, "grid": g, "sims": sims, "evaluator": evaluator_kind, "rollout": rollout_kind,
        "opening_plies": opening_plies, "seed": seed, "moves": moves,
        "outcome": manager.evaluate_final_state(mcts.state).value,
    }


def agent_match(name, g, spec_a, spec_b, seed):
    """Two MCTSAgents through tournament.play (re-rooting onto observed states, mcts.py:267-293)."""
    manager = manager_for(name, g)
    agents = []
    for sims, ev, ro in (spec_a, spec_b):
        m = MCTS(manager, sims, ref_rollout_policy(ro, manager) if ro else None,
                 ref_evaluator(ev, name, g) if ev else None)
        agents.append(MCTSAgent(m))
    logs = [[], []]
    for idx, agent in enumerate(agents):
        orig_step = agent.mcts.step

        def step(orig_step=orig_step, agent=agent, idx=idx):
            dist, stats = orig_step()
            ch = agent.mcts.root.children
            logs[idx].append({
                "visits": [ch[a].visits if a in ch else 0 for a in range(manager.num_actions)],
                "stats": list(stats), "order": list(ch.keys()), "root_visits": agent.mcts.root.visits,
            })
            return dist, stats

        agent.mcts.step = step
    outcome = play((agents[0], agents[1]), manager)
    return {"game": name, "grid": g, "a": list(spec_a), "b": list(spec_b),
            "log_a": logs[0], "log_b": logs[1], "outcome": outcome.value}


def make_mcts():
    games = []
    plan = [
        # name, g, sims, evaluator, rollout, opening plies, n games
        ("tictactoe", 3, 25, "uniform", None, 2, 4),
        ("tictactoe", 3, 25, "hashed", None, 1, 4),
        ("tictactoe", 3, 40, None, "min", 2, 2),
        ("hex", 4, 30, "uniform", None, 1, 2),
        ("hex", 7, 50, "uniform", None, 2, 2),
        ("hex", 7, 50, "hashed", None, 2, 2),
        ("hex", 5, 40, None, "max", 1, 2),
        ("hex_with_swap", 5, 40, "uniform", None, 1, 3),
        ("hex_with_swap", 7, 50, "hashed", None, 1, 2),
        ("hex_with_swap", 5, 30, "hashed", "min", 1, 2),
        ("othello", 4, 30, "uniform", None, 1, 2),
        ("othello", 6, 50, "uniform", None, 2, 2),
        ("othello", 6, 50, "hashed", None, 2, 2),
        ("othello", 8, 50, "uniform", None, 3, 2),
        ("othello", 8, 50, "hashed", None, 3, 2),
        ("othello", 8, 200, "hashed7", None, 2, 1),
        ("othello", 6, 30, None, "min", 1, 2),
        ("othello", 6, 30, "hashed", "max", 1, 1),
    ]
    seed = 0
    for name, g, sims, ev, ro, opening, n in plan:
        for _ in range(n):
            games.append(mcts_game(name, g, sims, ev, ro, opening, seed))
            print(f"mcts {name}{g} sims={sims} ev={ev} ro={ro}: {len(games[-1]['moves'])} moves")
            seed += 1
    matches = [
        agent_match("othello", 6, (30, "uniform", None), (20, "hashed", None), 0),
        agent_match("hex_with_swap", 5, (25, "hashed", None), (25, None, "min"), 1),
        agent_match("tictactoe", 3, (25, "uniform", None), (30, "hashed", None), 2),
        agent_match("hex", 5, (20, "hashed", "max"), (30, "uniform", None), 3),
    ]
    with open(HERE / "mcts.json", "w") as f:
        json.dump({"games": games, "matches": matches}, f, separators=(",", ":"))


# ------------------------------------------------------------------------- net
def make_net(name, g, channels, num_residual, hidden, n_positions, seed):
    manager = manager_for(name, g)
    og = make_game(name, g)
    sd = synthetic_state_dict(og, channels, num_residual, hidden, seed)
    if name == "tictactoe":
        net = ConvolutionalTicTacToeNet(manager, num_residual=num_residual, channels=channels,
                                        value_head_hidden_units=hidden)
    else:
        cls = {"hex": ConvolutionalHexNet, "hex_with_swap": ConvolutionalHexWithSwapNet,
               "othello": ConvolutionalOthelloNet}[name]
        net = cls(g, manager, num_residual=num_residual, channels=channels, value_head_hidden_units=hidden)
    net.load_state_dict(sd)
    rng = random.Random(seed)
    states = []
    while len(states) < n_positions:
        state = manager.initial_game_state()
        trail = []
        while not manager.is_final_state(state):
            trail.append(state)
            state = manager.generate_child_state(state, rng.choice(manager.legal_actions(state)))
        states.append(rng.choice(trail))
    net.net.eval()
    with torch.no_grad():
        raw_v, raw_logits = net.forward(net.states_to_tensor(states).float())
    values, probs = [], []
    for s in states:
        v, p = net.evaluate(s)
        values.append(v)
        probs.append(p.numpy())
    bits = [to_bits(s) for s in states]
    np.savez_compressed(
        HERE / f"net_{name}_{g}_c{channels}.npz",
        player=np.array([b[0] for b in bits], np.uint8),
        first=np.array([b[1] for b in bits], np.uint64), second=np.array([b[2] for b in bits], np.uint64),
        planes=net.states_to_tensor(states).numpy().astype(np.float32),
        raw_value=raw_v.numpy().astype(np.float32),
        raw_logits=raw_logits.reshape(len(states), -1).numpy().astype(np.float32),
        value=np.array(values, np.float32), probs=np.array(probs, np.float32),
        arch=np.array([channels, num_residual, hidden, seed], np.int32),
    )
    print(f"net_{name}_{g}_c{channels}: {len(states)} positions")


def main():
    make_games("tictactoe", 3, 400, 0)
    make_games("hex", 4, 300, 1)
    make_games("hex", 7, 500, 2)
    make_games("hex_with_swap", 5, 400, 3)
    make_games("hex_with_swap", 7, 400, 4)
    make_games("othello", 4, 300, 5)
    make_games("othello", 6, 600, 6)
    make_games("othello", 8, 900, 7)
    make_mcts()
    make_net("tictactoe", 3, 3, 3, 128, 24, 0)
    make_net("hex", 7, 128, 3, 128, 24, 1)
    make_net("hex_with_swap", 7, 128, 3, 128, 24, 2)
    make_net("hex_with_swap", 5, 32, 2, 64, 24, 3)
    make_net("othello", 8, 128, 3, 128, 32, 4)
    make_net("othello", 6, 128, 3, 128, 24, 5)
    with open(HERE / "META.json", "w") as f:
        json.dump({"python": sys.version, "torch": torch.__version__, "numpy": np.__version__,
                   "reference": "/root/reference (henribru/deep-mcts, unmodified)"}, f, indent=1)


if __name__ == "__main__":
    main()
