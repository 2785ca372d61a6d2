"""Trainable torch definition of the policy/value network.

Only used for training and as the parameter container: its ``state_dict`` has the reference's
key names and shapes (reference deep_mcts/convolutionalnet.py:22-170), so thesis checkpoints
load unchanged, and ``DeviceNet`` (device_net.py) packs that state_dict into the CUDA
inference kernels.  Inference during search never runs this module.

Architecture: 3x3 conv (in -> C) + BN + ReLU; ``num_residual`` blocks of
[3x3 conv + BN + ReLU + 3x3 conv + BN, skip, ReLU]; policy head 3x3 conv + BN + ReLU + FC to
``policy_features`` logits (reshaped to ``policy_shape``); value head 1x1 conv (C -> 1) + BN +
ReLU + FC(g*g -> hidden) + ReLU + FC(hidden -> 1).
"""
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv_bn(in_channels: int, out_channels: int, kernel_size: int) -> Tuple[nn.Conv2d, nn.BatchNorm2d]:
    return (nn.Conv2d(in_channels, out_channels, kernel_size, padding=kernel_size // 2),
            nn.BatchNorm2d(out_channels))


class ConvolutionalBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, padding: int) -> None:
        super().__init__()
        assert padding == kernel_size // 2
        self.conv1, self.bn1 = _conv_bn(in_channels, out_channels, kernel_size)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.relu(self.bn1(self.conv1(x)))


class ResidualBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, padding: int) -> None:
        super().__init__()
        assert padding == kernel_size // 2
        self.conv1, self.bn1 = _conv_bn(in_channels, out_channels, kernel_size)
        self.conv2, self.bn2 = _conv_bn(out_channels, out_channels, kernel_size)
        self.projection = nn.Identity() if in_channels == out_channels else nn.Conv2d(in_channels, out_channels, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = F.relu(self.bn1(self.conv1(x)))
        return F.relu(self.bn2(self.conv2(y)) + self.projection(x))


class PolicyHead(nn.Module):
    def __init__(self, grid_size: int, in_channels: int, out_features: int, out_shape: Tuple[int, ...]) -> None:
        super().__init__()
        self.out_shape = tuple(out_shape)
        self.conv1 = ConvolutionalBlock(in_channels, in_channels, 3, 1)
        self.fc1 = nn.Linear(in_channels * grid_size ** 2, out_features)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc1(self.conv1(x).flatten(1)).reshape(-1, *self.out_shape)


class ValueHead(nn.Module):
    def __init__(self, grid_size: int, in_channels: int, hidden_units: int) -> None:
        super().__init__()
        self.conv1 = ConvolutionalBlock(in_channels, 1, 1, 0)
        self.fc1 = nn.Linear(grid_size ** 2, hidden_units)
        self.fc2 = nn.Linear(hidden_units, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc2(F.relu(self.fc1(self.conv1(x).flatten(1))))


class ConvolutionalNet(nn.Module):
    def __init__(self, policy_features: int, policy_shape: Tuple[int, ...], grid_size: int, in_channels: int,
                 num_residual: int, channels: int, value_head_hidden_units: int) -> None:
        super().__init__()
        self.conv1 = ConvolutionalBlock(in_channels, channels, 3, 1)
        self.residual_blocks = nn.ModuleList(ResidualBlock(channels, channels, 3, 1) for _ in range(num_residual))
        self.policy_head = PolicyHead(grid_size, channels, policy_features, policy_shape)
        self.value_head = ValueHead(grid_size, channels, value_head_hidden_units)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        x = self.conv1(x)
        for block in self.residual_blocks:
            x = block(x)
        return self.value_head(x), self.policy_head(x).squeeze(1)
