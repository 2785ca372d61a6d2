"""The thesis' three evaluation experiments on the batched engine.

Reference drivers: deep_mcts/evaluate_training.py (every checkpoint of a run against a random-rollout
MCTS and against its predecessor, CSV), evaluate_simple_rollouts.py (round robin between
rollout_share 0, 0.2, ..., 1.0 on a run's last checkpoint, JSON) and evaluate_complex_rollouts.py
(greedy policy-network rollouts against uniform random rollouts at an equal time budget, with and
without the state evaluator, JSON).  Result locations and schemas are the reference's:
``<root>/training/<run>.csv``, ``<root>/simple_rollouts/<run>.json``,
``<root>/complex_rollouts/{with,without}_state_evaluator/<run>.json`` next to ``<root>/saves/<run>/``.

The reference fans the run directories out over 10 worker processes, each playing its games one at
a time.  Here the games of a match advance in lock-step on the device (``compare_agents_batched`` /
``tournament_batched``), so runs are simply walked in order; the time-budgeted experiment makes every
game its own sequence of searches and uses the sequential ``compare_agents``.  The modules
``evaluate_training`` / ``evaluate_simple_rollouts`` / ``evaluate_complex_rollouts`` re-export the
entry points under the reference's names.
"""
import json
from pathlib import Path
from typing import Callable, List, Type

import pandas as pd
import torch

from .batched_tournament import AgentSpec, compare_agents_batched, tournament_batched
from .gamenet import GameNet, cached_state_evaluator
from .mcts import MCTS, MCTSAgent, greedy_policy_rollout, random_rollout_policy
from .tournament import compare_agents
from .train import EVALUATION_COLUMNS

TRAINING_COLUMNS = [c for c in EVALUATION_COLUMNS if c != "training_games"]
NO_COMPARISON = ((0.0, 0.0), (0.0, 0.0), (0.0, 0.0))


def checkpoint_files(model_dir: Path) -> List[Path]:
    """``anet-<iterations>.tar`` files of a run, oldest first."""
    return sorted((f for f in Path(model_dir).iterdir() if f.name.endswith(".tar")), key=lambda f: int(f.name[5:-4]))


def for_each_run(save_dir: Path, evaluate_run: Callable[[Path], None], last: int = 20) -> None:
    """Apply ``evaluate_run`` to the ``last`` most recent run directories under ``save_dir``."""
    for model_dir in sorted(Path(save_dir).iterdir())[-last:]:
        evaluate_run(model_dir)


def result_dir(model_dir: Path, *parts: str) -> Path:
    out = Path(model_dir).parent.parent.joinpath(*parts)
    out.mkdir(parents=True, exist_ok=True)
    return out


# ------------------------------------------------------------------------------- experiment 1
def training_models(model_dir: Path, net_class: Type[GameNet], manager, device: torch.device,
                    num_games: int = 40, num_simulations: int = 50) -> None:
    model_dir = Path(model_dir)
    random_spec = AgentSpec(2 * num_simulations, random_rollout_policy(manager), None)
    previous_spec = None
    rows = []
    for k, model_file in enumerate(checkpoint_files(model_dir)):
        print(model_dir.name, model_file.name)
        evaluator = cached_state_evaluator(net_class.from_path_full(str(model_file), manager).to(device))
        spec = AgentSpec(num_simulations, None, evaluator, epsilon=0.05)
        against_previous = NO_COMPARISON if previous_spec is None else compare_agents_batched(
            (spec, previous_spec), num_games, manager, seed=2 * k)
        previous_spec = spec
        against_random = compare_agents_batched((AgentSpec(num_simulations, None, evaluator), random_spec), num_games,
                                                manager, seed=2 * k + 1)
        rows.append([int(model_file.name[5:-4]), *(x for pair in against_random for x in pair),
                     *(x for pair in against_previous for x in pair)])
    pd.DataFrame(rows, columns=TRAINING_COLUMNS).to_csv(str(result_dir(model_dir, "training") / f"{model_dir.name}.csv"))


# ------------------------------------------------------------------------------- experiment 2
def simple_rollout_models(model_dir: Path, net_class: Type[GameNet], manager, device: torch.device,
                          num_games: int = 40, num_simulations: int = 50) -> None:
    model_dir = Path(model_dir)
    print(model_dir.name)
    model = net_class.from_path_full(str(checkpoint_files(model_dir)[-1]), manager).to(device)
    evaluator = cached_state_evaluator(model)
    # rollouts, the share draw and the value mix all happen inside the search kernels
    specs = [AgentSpec(num_simulations, random_rollout_policy(manager) if percent else None, evaluator,
                       rollout_share=percent / 100) for percent in range(0, 101, 20)]
    with open(result_dir(model_dir, "simple_rollouts") / f"{model_dir.name}.json", "w") as f:
        json.dump(tournament_batched(specs, num_games, manager), f)


# ------------------------------------------------------------------------------- experiment 3
def reset_caches(agent: MCTSAgent) -> None:
    """Per-game cache reset of the reference; only the evaluator memo exists here, the rules are
    stateless kernels."""
    if hasattr(agent.mcts.state_evaluator, "cache_clear"):
        agent.mcts.state_evaluator.cache_clear()


def complex_rollout_models(model_dir: Path, net_class: Type[GameNet], manager, device: torch.device,
                           num_games: int = 240, time_per_move: float = 1.0) -> None:
    model_dir = Path(model_dir)
    print(model_dir.name)
    model_file = checkpoint_files(model_dir)[-1]
    for use_evaluator in (True, False):
        model = net_class.from_path_full(str(model_file), manager).to(device)
        evaluator = cached_state_evaluator(model) if use_evaluator else None

        def timed_agent(rollout_policy) -> MCTSAgent:
            return MCTSAgent(MCTS(manager, float("inf"), rollout_policy, evaluator, rollout_share=1.0,
                                  time_per_move=time_per_move), reset_fn=reset_caches)

        complex_agent, simple_agent = timed_agent(greedy_policy_rollout(model)), timed_agent(random_rollout_policy(manager))
        results = compare_agents((complex_agent, simple_agent), num_games, manager)
        sub = "with_state_evaluator" if use_evaluator else "without_state_evaluator"
        with open(result_dir(model_dir, "complex_rollouts", sub) / f"{model_dir.name}.json", "w") as f:
            json.dump({"results": results, "complex_simulations": complex_agent.simulation_stats,
                       "simple_simulations": simple_agent.simulation_stats}, f)
