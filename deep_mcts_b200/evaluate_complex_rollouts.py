"""Experiment 3: greedy policy-network rollouts against uniform random rollouts at an equal time
budget per move, with and without the state evaluator (interface and JSON schema of reference
deep_mcts/evaluate_complex_rollouts.py:20-109).  A wall-clock budget makes every game its own
sequence of searches, so this driver uses the sequential ``compare_agents`` with two
time-budgeted ``MCTSAgent`` s; the playouts themselves are device kernels."""
import json
from pathlib import Path
from typing import Type

import torch

from .evaluate_training import checkpoint_files
from .gamenet import GameNet, cached_state_evaluator
from .mcts import MCTS, MCTSAgent, greedy_policy_rollout, random_rollout_policy
from .tournament import compare_agents


def evaluate_complex_rollouts(save_dir: Path, net_class: Type[GameNet], manager, device: torch.device,
                              num_games: int = 240, time_per_move: float = 1.0, last: int = 20) -> None:
    for model_dir in sorted(Path(save_dir).iterdir())[-last:]:
        evaluate_models(model_dir, net_class, manager, device, num_games, time_per_move)


def reset_caches(agent: MCTSAgent) -> None:
    """Per-game cache reset of the reference (evaluate_complex_rollouts.py:103-109); only the evaluator
    memo exists here, the game rules are stateless kernels."""
    if hasattr(agent.mcts.state_evaluator, "cache_clear"):
        agent.mcts.state_evaluator.cache_clear()


def evaluate_models(model_dir: Path, net_class: Type[GameNet], manager, device: torch.device,
                    num_games: int = 240, time_per_move: float = 1.0) -> None:
    model_dir = Path(model_dir)
    print(model_dir.name)
    model_file = checkpoint_files(model_dir)[-1]
    for with_state_evaluator, dir_name in ((True, "with_state_evaluator"), (False, "without_state_evaluator")):
        model = net_class.from_path_full(str(model_file), manager).to(device)
        cached_model = cached_state_evaluator(model)
        complex_agent = MCTSAgent(
            MCTS(manager, float("inf"), greedy_policy_rollout(model), cached_model if with_state_evaluator else None,
                 rollout_share=1.0, time_per_move=time_per_move), reset_fn=reset_caches)
        simple_agent = MCTSAgent(
            MCTS(manager, float("inf"), random_rollout_policy(manager), cached_model if with_state_evaluator else None,
                 rollout_share=1.0, time_per_move=time_per_move), reset_fn=reset_caches)
        result_dir = model_dir.parent.parent / "complex_rollouts" / dir_name
        result_dir.mkdir(parents=True, exist_ok=True)
        results = compare_agents((complex_agent, simple_agent), num_games, manager)
        with open(result_dir / f"{model_dir.name}.json", "w") as f:
            json.dump({"results": results, "complex_simulations": complex_agent.simulation_stats,
                       "simple_simulations": simple_agent.simulation_stats}, f)
