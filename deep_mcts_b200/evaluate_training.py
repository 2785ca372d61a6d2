"""Experiment 1 under the reference's module name (deep_mcts/evaluate_training.py); see experiments.py."""
from functools import partial

from .experiments import TRAINING_COLUMNS as COLUMNS, checkpoint_files, for_each_run, training_models as evaluate_models  # noqa: F401


def evaluate_training(save_dir, net_class, manager, device, num_games: int = 40, num_simulations: int = 50,
                      last: int = 20) -> None:
    for_each_run(save_dir, partial(evaluate_models, net_class=net_class, manager=manager, device=device,
                                   num_games=num_games, num_simulations=num_simulations), last)
