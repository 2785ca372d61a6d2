"""Experiment 1 evaluation: every checkpoint of a training run against a random-rollout MCTS and
against the previous checkpoint (interface and CSV schema of reference
deep_mcts/evaluate_training.py:20-134).

The reference fans the run directories out over 10 worker processes, each playing its games one
at a time; here the games of a match advance in lock-step on the device
(``compare_agents_batched``), so the directories are simply walked in order.
"""
from pathlib import Path
from typing import Type

import pandas as pd
import torch

from .batched_tournament import AgentSpec, compare_agents_batched
from .gamenet import GameNet, cached_state_evaluator
from .mcts import random_rollout_policy
from .train import EVALUATION_COLUMNS

COLUMNS = [c for c in EVALUATION_COLUMNS if c != "training_games"]


def checkpoint_files(model_dir: Path):
    """``anet-<iterations>.tar`` files of a run, oldest first (evaluate_training.py:39-42)."""
    return sorted((f for f in Path(model_dir).iterdir() if f.name.endswith(".tar")), key=lambda f: int(f.name[5:-4]))


def evaluate_training(save_dir: Path, net_class: Type[GameNet], manager, device: torch.device,
                      num_games: int = 40, num_simulations: int = 50, last: int = 20) -> None:
    for model_dir in sorted(Path(save_dir).iterdir())[-last:]:
        evaluate_models(model_dir, net_class, manager, device, num_games, num_simulations)


def evaluate_models(model_dir: Path, net_class: Type[GameNet], manager, device: torch.device,
                    num_games: int = 40, num_simulations: int = 50) -> None:
    model_dir = Path(model_dir)
    files = checkpoint_files(model_dir)
    random_spec = AgentSpec(2 * num_simulations, random_rollout_policy(manager), None)
    previous_spec = None
    rows = []
    for k, model_file in enumerate(files):
        print(model_dir.name, model_file.name)
        model = net_class.from_path_full(str(model_file), manager).to(device)
        evaluator = cached_state_evaluator(model)
        spec = AgentSpec(num_simulations, None, evaluator, epsilon=0.05)
        if previous_spec is None:
            previous_evaluation = ((0.0, 0.0), (0.0, 0.0), (0.0, 0.0))
        else:
            previous_evaluation = compare_agents_batched((spec, previous_spec), num_games, manager, seed=2 * k)
        previous_spec = spec
        random_evaluation = compare_agents_batched((AgentSpec(num_simulations, None, evaluator), random_spec),
                                                   num_games, manager, seed=2 * k + 1)
        rows.append([int(model_file.name[5:-4]),
                     *[x for pair in random_evaluation for x in pair],
                     *[x for pair in previous_evaluation for x in pair]])
    result_dir = model_dir.parent.parent / "training"
    result_dir.mkdir(exist_ok=True)
    pd.DataFrame(rows, columns=COLUMNS).to_csv(str(result_dir / f"{model_dir.name}.csv"))
