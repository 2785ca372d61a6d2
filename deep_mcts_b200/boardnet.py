"""Shared body of the four per-game networks.

The reference repeats one class per game (``<game>/convolutionalnet.py``: constructor, ``copy``,
``from_path_full``, ``parameters``, ``states_to_tensor`` and, for the Hex variants, the policy
un-transpose).  The games differ only in a handful of facts, which a subclass states as class
attributes:

* ``manager_cls``      rules object built when the caller passes none
* ``extra_actions``    1 when the policy has a pass / swap entry after the board cells
* ``policy_as_grid``   policy head shaped ``[g, g]`` (Hex, tic-tac-toe) instead of flat
* ``transpose_second`` the net sees the transposed board when SECOND moves (Hex variants), so the
  board part of its policy is transposed back in ``forward``
* ``fixed_grid``       board side when it is not a constructor argument (tic-tac-toe)
* ``default_channels`` trunk width when the caller does not choose one

Constructor argument order and the saved-parameter dictionary are those of the reference classes.
"""
from typing import Any, Dict, Mapping, Optional, Sequence, Tuple

import torch

from .convolutionalnet import ConvolutionalNet
from .game import GameManager, Player
from .gamenet import GameNet, _torch_load, planes_from_states

SGD_DEFAULTS = {"lr": 0.01, "momentum": 0.9, "weight_decay": 0.001}
_ARCH_KEYS = ("num_residual", "channels", "value_head_hidden_units")


class BoardGameNet(GameNet):
    manager_cls = None
    extra_actions = 0
    policy_as_grid = False
    transpose_second = False
    fixed_grid: Optional[int] = None
    default_channels = 128

    def __init__(self, grid_size: Optional[int] = None, manager: Optional[GameManager] = None,
                 optimizer_cls=torch.optim.SGD, optimizer_args: Tuple[Any, ...] = (),
                 optimizer_kwargs: Mapping[str, Any] = SGD_DEFAULTS, num_residual: int = 3,
                 channels: Optional[int] = None, value_head_hidden_units: int = 128) -> None:
        g = self.fixed_grid if self.fixed_grid is not None else int(grid_size)
        channels = self.default_channels if channels is None else channels
        actions = g * g + self.extra_actions
        shape = (g, g) if self.policy_as_grid else (actions,)
        if manager is None:
            manager = self.manager_cls() if self.fixed_grid is not None else self.manager_cls(g)
        body = ConvolutionalNet(policy_features=actions, policy_shape=shape, grid_size=g, in_channels=3,
                                num_residual=num_residual, channels=channels,
                                value_head_hidden_units=value_head_hidden_units)
        super().__init__(body, manager, optimizer_cls, optimizer_args, optimizer_kwargs)
        self.grid_size, self.num_residual, self.channels = g, num_residual, channels
        self.value_head_hidden_units = value_head_hidden_units

    # ---- what differs between games ------------------------------------------------------
    def states_to_tensor(self, states: Sequence) -> torch.Tensor:
        return planes_from_states(states, transpose_second=self.transpose_second)

    def forward(self, states: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        values, logits = super().forward(states)
        if not self.transpose_second:
            return values, logits
        # rows whose player plane says SECOND were fed a transposed board: transpose the board part
        # of their policy back; a swap logit (last entry of a flat policy) stays where it is
        second = states[:, 2, 0, 0] == int(Player.SECOND)
        logits = logits.clone()
        if self.policy_as_grid:
            logits[second] = logits[second].transpose(1, 2)
        else:
            g, cells = self.grid_size, self.grid_size ** 2
            board = logits[second, :cells].reshape(-1, g, g).transpose(1, 2)
            logits[second, :cells] = board.flatten(start_dim=1)
        return values, logits

    # ---- construction from another net / from disk ---------------------------------------
    def _ctor_args(self) -> Dict[str, Any]:
        args = {"manager": self.manager, "optimizer_cls": self.optimizer_cls, "optimizer_args": self.optimizer_args,
                "optimizer_kwargs": self.optimizer_kwargs}
        args.update({k: getattr(self, k) for k in _ARCH_KEYS})
        if self.fixed_grid is None:
            args["grid_size"] = self.grid_size
        return args

    def copy(self):
        twin = type(self)(**self._ctor_args())
        twin.load_state_dict(self.net.state_dict())
        return twin

    @classmethod
    def from_path_full(cls, path: str, manager: Optional[GameManager] = None):
        """Rebuild a net from a ``save_full`` checkpoint (the reference's pickled dictionary)."""
        blob = _torch_load(path, torch.device("cpu"))
        args = {k: blob[k] for k in ("optimizer_cls", "optimizer_args", "optimizer_kwargs") + _ARCH_KEYS}
        if cls.fixed_grid is None:
            args["grid_size"] = blob["grid_size"]
        net = cls(manager=manager, **args)
        net.load_state_dict(blob["state_dict"])
        return net

    def parameters(self) -> Dict[str, Any]:
        out = dict(super().parameters())
        if self.fixed_grid is None:
            out["grid_size"] = self.grid_size
        out.update({k: getattr(self, k) for k in _ARCH_KEYS})
        return out
