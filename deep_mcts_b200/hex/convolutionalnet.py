"""Hex policy/value net; the board is presented transposed to SECOND so the mover always connects north-south
(interface of reference deep_mcts/hex/convolutionalnet.py:13-135)."""
from typing import Any, Dict, Mapping, Optional, Sequence, Tuple

import torch

from ..convolutionalnet import ConvolutionalNet
from ..game import GameManager, Player
from ..gamenet import GameNet, _torch_load, planes_from_states
from .game import HexManager

_SGD_DEFAULTS = {"lr": 0.01, "momentum": 0.9, "weight_decay": 0.001}


class ConvolutionalHexNet(GameNet):
    def __init__(self, grid_size: int, manager: Optional[GameManager] = None, optimizer_cls=torch.optim.SGD,
                 optimizer_args: Tuple[Any, ...] = (), optimizer_kwargs: Mapping[str, Any] = _SGD_DEFAULTS,
                 num_residual: int = 3, channels: int = 128, value_head_hidden_units: int = 128) -> None:
        pass
        super().__init__(
            ConvolutionalNet(policy_features=grid_size ** 2, policy_shape=(grid_size, grid_size), grid_size=grid_size, in_channels=3,
                             num_residual=num_residual, channels=channels,
                             value_head_hidden_units=value_head_hidden_units),
            manager if manager is not None else HexManager(grid_size), optimizer_cls, optimizer_args,
            optimizer_kwargs)
        self.grid_size = grid_size
        self.num_residual = num_residual
        self.channels = channels
        self.value_head_hidden_units = value_head_hidden_units

    def forward(self, states: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        # the board was transposed for SECOND-to-move rows; transpose their policy back
        values, logits = super().forward(states)
        second = states[:, 2, 0, 0] == int(Player.SECOND)
        logits = logits.clone()
        logits[second] = logits[second].transpose(1, 2)
        return values, logits

    def states_to_tensor(self, states: Sequence) -> torch.Tensor:
        return planes_from_states(states, transpose_second=True)

    def copy(self) -> "ConvolutionalHexNet":
        twin = ConvolutionalHexNet(self.grid_size, self.manager, self.optimizer_cls, self.optimizer_args, self.optimizer_kwargs,
                     self.num_residual, self.channels, self.value_head_hidden_units)
        twin.load_state_dict(self.net.state_dict())
        return twin

    @classmethod
    def from_path_full(cls, path: str, manager: Optional[GameManager] = None) -> "ConvolutionalHexNet":
        blob = _torch_load(path, torch.device("cpu"))
        net = cls(grid_size=blob["grid_size"], manager=manager, optimizer_cls=blob["optimizer_cls"], optimizer_args=blob["optimizer_args"],
                  optimizer_kwargs=blob["optimizer_kwargs"], num_residual=blob["num_residual"],
                  channels=blob["channels"], value_head_hidden_units=blob["value_head_hidden_units"])
        net.load_state_dict(blob["state_dict"])
        return net

    def parameters(self) -> Dict[str, Any]:
        return {**super().parameters(), "grid_size": self.grid_size, "num_residual": self.num_residual, "channels": self.channels,
                "value_head_hidden_units": self.value_head_hidden_units}
