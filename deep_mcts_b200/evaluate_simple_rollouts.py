"""Experiment 2: how much weight should random rollouts get next to the network's value?  A round
robin between agents with rollout_share 0, 0.2, ..., 1.0 on the last checkpoint of each run
(interface and JSON schema of reference deep_mcts/evaluate_simple_rollouts.py:18-67).
Rollouts, the share draw and the value mix all happen inside the search kernels."""
import json
from pathlib import Path
from typing import Type

import torch

from .batched_tournament import AgentSpec, tournament_batched
from .evaluate_training import checkpoint_files
from .gamenet import GameNet, cached_state_evaluator
from .mcts import random_rollout_policy


def evaluate_simple_rollouts(save_dir: Path, net_class: Type[GameNet], manager, device: torch.device,
                             num_games: int = 40, num_simulations: int = 50, last: int = 20) -> None:
    for model_dir in sorted(Path(save_dir).iterdir())[-last:]:
        evaluate_models(model_dir, net_class, manager, device, num_games, num_simulations)


def evaluate_models(model_dir: Path, net_class: Type[GameNet], manager, device: torch.device,
                    num_games: int = 40, num_simulations: int = 50) -> None:
    model_dir = Path(model_dir)
    model_file = checkpoint_files(model_dir)[-1]
    model = net_class.from_path_full(str(model_file), manager).to(device)
    evaluator = cached_state_evaluator(model)
    specs = [AgentSpec(num_simulations, random_rollout_policy(manager) if share > 0 else None, evaluator,
                       rollout_share=share / 100)
             for share in range(0, 101, 20)]
    result_dir = model_dir.parent.parent / "simple_rollouts"
    result_dir.mkdir(exist_ok=True)
    print(model_dir.name)
    results = tournament_batched(specs, num_games, manager)
    with open(result_dir / f"{model_dir.name}.json", "w") as f:
        json.dump(results, f)
