"""GameNet: the policy/value network wrapper of the reference API, with leaf evaluation on
the CUDA engine.

Interface of reference deep_mcts/gamenet.py:33-234: ``evaluate(state) -> (value, probs)``,
``forward``, ``train``, ``copy``, ``to``, ``save`` / ``save_full`` / ``load`` /
``from_path`` / ``from_path_full``, ``parameters``, ``load_state_dict``.
``evaluate`` (the MCTS hot path) runs on ``DeviceNet``; ``forward`` / ``train`` keep torch
autograd so the training step and its NCCL gradient all-reduce stay ordinary PyTorch.
"""
from abc import ABC, abstractmethod
from functools import lru_cache
from typing import Any, Callable, Dict, Mapping, Optional, Sequence, Tuple, Type

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ._engine import BoardManager, state_to_bits
from .device_net import DeviceNet
from .game import Action, GameManager, Player, State
from .tournament import Agent


def cross_entropy(pred: torch.Tensor, soft_targets: torch.Tensor, reduction: str = "mean") -> torch.Tensor:
    """Cross entropy against a full target distribution (visit counts), per sample then reduced."""
    pred = pred.reshape(pred.shape[0], -1)
    soft_targets = soft_targets.reshape(soft_targets.shape[0], -1)
    losses = -(soft_targets * F.log_softmax(pred, dim=1)).sum(dim=1)
    if reduction == "mean":
        return losses.mean()
    if reduction == "sum":
        return losses.sum()
    return losses


def accuracy(predictions: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    hit = predictions.reshape(predictions.shape[0], -1).argmax(dim=1) == targets.reshape(targets.shape[0], -1).argmax(dim=1)
    return hit.float().mean()


class GameNet(ABC):
    """Holds the trainable torch module, its optimizer and the device inference engine."""

    def __init__(self, net: nn.Module, manager: GameManager, optimizer_cls, optimizer_args: Tuple[Any, ...],
                 optimizer_kwargs: Mapping[str, Any]) -> None:
        self.policy_criterion = cross_entropy
        self.value_criterion = nn.MSELoss()
        self.device = torch.device("cpu")
        self.net = net
        self.manager = manager
        self.optimizer_cls = optimizer_cls
        self.optimizer_args = optimizer_args
        self.optimizer_kwargs = optimizer_kwargs
        self.optimizer = optimizer_cls(self.net.parameters(), *optimizer_args, **optimizer_kwargs)
        self._device_net: Optional[DeviceNet] = None
        self._device_net_stale = True
        self.precision: Optional[str] = None  # None = bf16 when the architecture allows, else fp32
        # optional hook run between backward() and optimizer.step(), e.g.
        # deep_mcts_b200.distributed.allreduce_gradients for data-parallel training
        self.grad_sync: Optional[Callable] = None
        # deep_mcts_b200.distributed.GradientBucket: gradients as views into one flat buffer, all-reduced in place
        self.grad_bucket = None
        self.max_batch = 4096
        # lets MCTS recognise `net.evaluate` as a device evaluator
        self._dm_net_self = self

    on_device = True

    # ------------------------------------------------------------------ device inference
    def _arch(self) -> Tuple[int, int, int]:
        """(channels, num_residual, value_head_hidden_units) of self.net."""
        return self.channels, self.num_residual, self.value_head_hidden_units

    def device_net(self) -> DeviceNet:
        if self._device_net is None:
            c, r, h = self._arch()
            self._device_net = DeviceNet(self.manager, c, r, h, self.max_batch, self.precision)
            self._device_net_stale = True
        if self._device_net_stale:
            self._device_net.load_state_dict(self.net.state_dict())
            self._device_net_stale = False
        return self._device_net

    def sync_device_weights(self) -> None:
        """Re-pack the torch weights into the CUDA kernels (the reference's weight transfer,
        train.py:169-171)."""
        self._device_net_stale = True
        self.device_net()

    def evaluate(self, state) -> Tuple[float, torch.Tensor]:
        """(value in [0,1] from SECOND's perspective, probabilities over all actions on the CPU)
        -- reference gamenet.py:61-85, computed by the CUDA kernels."""
        p, f, s = state_to_bits(state)
        value, probs = self.device_net().evaluate([p], [f], [s])
        return float(value[0]), torch.from_numpy(probs[0].copy())

    def evaluate_batch(self, player, first, second):
        return self.device_net().evaluate(player, first, second)

    def run_search(self, pool, sims: int) -> None:
        self.device_net().run_search(pool, sims)

    # --------------------------------------------------------------------- torch training
    def forward(self, states: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.net.forward(states)

    def train(self, states: torch.Tensor, probability_targets: torch.Tensor, value_targets: torch.Tensor):
        """One SGD step: soft-target cross entropy + MSE on tanh(value) (reference gamenet.py:87-110)."""
        self.net.train()
        values, logits = self.forward(states.float())
        values = torch.tanh(values)
        policy_loss = self.policy_criterion(logits, probability_targets)
        value_loss = self.value_criterion(values, value_targets)
        acc = accuracy(logits, probability_targets)
        bucket = self.grad_bucket
        if bucket is not None and bucket.attached():
            bucket.zero()                       # gradients live in one flat buffer: keep the views
        else:
            bucket = None
            self.optimizer.zero_grad()
        (policy_loss + value_loss).backward()
        if bucket is not None:
            bucket.allreduce()
        elif self.grad_sync is not None:
            self.grad_sync(self.net.parameters())
        self.optimizer.step()
        self._device_net_stale = True
        return policy_loss, value_loss, acc

    def distributions_to_tensor(self, distributions: Sequence[Sequence[float]]) -> torch.Tensor:
        targets = torch.tensor(distributions, dtype=torch.float32)
        assert targets.shape == (len(distributions), self.manager.num_actions)
        return targets

    # ------------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        torch.save(self.net.state_dict(), path)

    def save_full(self, path: str) -> None:
        torch.save(self.parameters(), path)

    def load(self, path: str) -> None:
        blob = _torch_load(path, self.device)
        self.load_state_dict(blob if path.endswith(".pth") else blob["state_dict"])

    def parameters(self) -> Dict[str, Any]:
        return {"optimizer_cls": self.optimizer_cls, "optimizer_args": self.optimizer_args,
                "optimizer_kwargs": self.optimizer_kwargs, "state_dict": self.net.state_dict()}

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self.net.load_state_dict(state_dict)
        self._device_net_stale = True

    @abstractmethod
    def states_to_tensor(self, states: Sequence[State]) -> torch.Tensor: ...

    @abstractmethod
    def copy(self): ...

    def to(self, device: torch.device):
        self.device = torch.device(device)
        self.value_criterion.to(device)
        self.net.to(device)
        self.optimizer = self.optimizer_cls(self.net.parameters(), *self.optimizer_args, **self.optimizer_kwargs)
        return self

    @classmethod
    def from_path(cls, path: str, *args: Any, **kwargs: Any):
        net = cls(*args, **kwargs)
        net.load(path)
        return net

    @classmethod
    @abstractmethod
    def from_path_full(cls, path: str, manager: Optional[GameManager] = None): ...


def _torch_load(path: str, device) -> Any:
    """torch.load for the reference's checkpoints, which pickle the optimizer class."""
    torch.serialization.add_safe_globals([torch.optim.SGD, torch.optim.Adam])
    return torch.load(path, map_location=device)


def planes_from_states(states: Sequence[State], transpose_second: bool) -> torch.Tensor:
    """[B, 3, g, g] fp32 input planes: mover's stones, opponent's stones, player id; Hex variants
    transpose the board when SECOND moves.  (Training-side twin of csrc/net_fp32.cuh write_planes.)"""
    g = len(states[0].grid)
    grids = torch.tensor([[[int(c) for c in row] for row in s.grid] for s in states], dtype=torch.int64)
    players = torch.tensor([int(s.player) for s in states], dtype=torch.int64)
    if transpose_second:
        second = players == int(Player.SECOND)
        grids[second] = grids[second].transpose(1, 2)
    me = players.view(-1, 1, 1)
    planes = torch.stack([(grids == me).float(), (grids == 1 - me).float(),
                          players.float().view(-1, 1, 1).expand(-1, g, g)], dim=1)
    return planes


def cached_state_evaluator(game_net: GameNet) -> Callable[[State], Tuple[float, Sequence[float]]]:
    """Memoised ``net.evaluate`` returning Python floats (reference deep_mcts/train.py:408-414).
    MCTS recognises it and evaluates leaves in device batches instead of calling it per state."""

    @lru_cache(2 ** 20)
    def inner(state):
        value, probabilities = game_net.evaluate(state)
        return value, probabilities.tolist()

    inner._dm_net = game_net
    return inner


class GameNetAgent(Agent):
    def __init__(self, net: GameNet) -> None:
        self.net = net

    def reset(self) -> None:
        pass


class GreedyGameNetAgent(GameNetAgent):
    """Plays the network policy's best legal move, a sampled one with probability ``epsilon`` (the reference
    declares this agent but leaves ``play`` unimplemented, gamenet.py:196-205)."""

    def __init__(self, net: GameNet, epsilon: float = 0) -> None:
        super().__init__(net)
        self.epsilon = epsilon

    def play(self, state) -> Action:
        probabilities = self.net.evaluate(state)[1].numpy().astype(np.float64)
        if self.epsilon > 0 and np.random.random_sample() < self.epsilon:
            return int(np.random.choice(len(probabilities), p=probabilities / probabilities.sum()))
        return int(np.argmax(probabilities))


class SamplingGameNetAgent(GameNetAgent):
    """Samples its move from the network's masked policy (declared, unimplemented in the reference, gamenet.py:208-210)."""

    def play(self, state) -> Action:
        probabilities = self.net.evaluate(state)[1].numpy().astype(np.float64)
        return int(np.random.choice(len(probabilities), p=probabilities / probabilities.sum()))
