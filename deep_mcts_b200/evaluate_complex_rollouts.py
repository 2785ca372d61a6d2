"""Experiment 3 under the reference's module name (deep_mcts/evaluate_complex_rollouts.py); see experiments.py."""
from functools import partial

from .experiments import complex_rollout_models as evaluate_models, for_each_run, reset_caches  # noqa: F401


def evaluate_complex_rollouts(save_dir, net_class, manager, device, num_games: int = 240, time_per_move: float = 1.0,
                              last: int = 20) -> None:
    for_each_run(save_dir, partial(evaluate_models, net_class=net_class, manager=manager, device=device,
                                   num_games=num_games, time_per_move=time_per_move), last)
