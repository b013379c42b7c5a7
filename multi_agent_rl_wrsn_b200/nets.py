"""The two networks the reference's trainers put on the hot path's observations, written from their shapes.

* ``UNetActor``  — ``controller/ppo/actor/UnetActor.py:61-80``: [n, 4, S, S] observation -> mean density map [n, S, S] and
  a state-independent ``log_std`` map.  Encoder 4 -> 64 -> (pool) 128 -> (pool) 256, decoder with bilinear x2 upsampling
  (``align_corners=True``) and skip concatenation 256 + 128 -> 128, 128 + 64 -> 64, 3 x 3 head to one channel;
  936 401 parameters at S = 100.
* ``CNNCritic``  — ``controller/ppo/critic/CNNCritic.py:7-48``: three 5 x 5 stride-2 convolutions 4 -> 16 -> 32 -> 64
  (100 -> 50 -> 25 -> 13), 10 816 -> 100 -> 1; 1 147 513 parameters.

Parameter names follow the reference's modules (``inc.conv.weight``, ``down1.conv_block.bn.running_mean``, ``fc1.bias`` ...)
so that ``actor.pth`` / ``critic.pth`` written by either trainer load into the other (``IPPO.py:296-309``, ``:50-64``).
Initialisation as the reference's ``utils.layer_init`` (``utils.py:20-23``): orthogonal weights with gain sqrt(2) (0.1 for
the actor's head, 1 for the critic's output), zero biases.  The convolutions are cuDNN's, as in the reference: these are the
CALLERS of the simulator, not part of it — they live here so that the rollout + update loop (``ippo.BatchedIPPO``,
``bench.py --workload ippo``) runs on a box that has no copy of the reference.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _ortho(layer, gain=math.sqrt(2.0)):
    nn.init.orthogonal_(layer.weight, gain)
    nn.init.zeros_(layer.bias)
    return layer


class _Stage(nn.Module):
    """3 x 3 convolution + batch norm + ReLU under the reference's attribute names (``conv`` / ``bn``)."""

    def __init__(self, c_in, c_out):
        super().__init__()
        self.conv = _ortho(nn.Conv2d(c_in, c_out, 3, padding=1))
        self.bn = nn.BatchNorm2d(c_out)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)), inplace=True)


class _Wrapped(nn.Module):
    """A stage reached through ``<name>.conv_block`` (the reference's Down / Up containers)."""

    def __init__(self, c_in, c_out):
        super().__init__()
        self.conv_block = _Stage(c_in, c_out)

    def forward(self, x):
        return self.conv_block(x)


class _Head(nn.Module):
    def __init__(self, c_in):
        super().__init__()
        self.conv = _ortho(nn.Conv2d(c_in, 1, 3, padding=1), 0.1)

    def forward(self, x):
        return self.conv(x)


class UNetActor(nn.Module):
    WIDTHS = (64, 128, 256)

    def __init__(self, map_size=100, in_channels=4):
        super().__init__()
        w0, w1, w2 = self.WIDTHS
        self.inc = _Stage(in_channels, w0)
        self.down1 = _Wrapped(w0, w1)
        self.down2 = _Wrapped(w1, w2)
        self.up1 = _Wrapped(w2 + w1, w1)
        self.up2 = _Wrapped(w1 + w0, w0)
        self.out_mean = _Head(w0)
        self.log_std = nn.Parameter(torch.zeros((1, 1, map_size, map_size)))

    @staticmethod
    def _merge(deep, skip):
        """bilinear x2 (align_corners), zero-pad to the skip's size, skip first in the channel concatenation"""
        up = F.interpolate(deep, scale_factor=2, mode="bilinear", align_corners=True)
        dy, dx = skip.shape[2] - up.shape[2], skip.shape[3] - up.shape[3]
        if dy or dx:
            up = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return torch.cat([skip, up], dim=1)

    def forward(self, obs):
        s0 = self.inc(obs)
        s1 = self.down1(F.max_pool2d(s0, 2))
        s2 = self.down2(F.max_pool2d(s1, 2))
        y = self.up1(self._merge(s2, s1))
        y = self.up2(self._merge(y, s0))
        mean = self.out_mean(y)
        return mean.squeeze(), self.log_std.expand_as(mean).squeeze()


class CNNCritic(nn.Module):
    def __init__(self, map_size=100, in_channels=4):
        super().__init__()
        side = map_size
        for _ in range(3):
            side = (side + 2 * 2 - 5) // 2 + 1
        self.conv1 = _ortho(nn.Conv2d(in_channels, 16, 5, stride=2, padding=2))
        self.conv2 = _ortho(nn.Conv2d(16, 32, 5, stride=2, padding=2))
        self.conv3 = _ortho(nn.Conv2d(32, 64, 5, stride=2, padding=2))
        self.fc1 = _ortho(nn.Linear(64 * side * side, 100))
        self.fc2 = _ortho(nn.Linear(100, 1), 1.0)

    def forward(self, obs):
        x = obs
        for conv in (self.conv1, self.conv2, self.conv3):
            x = F.relu(conv(x), inplace=True)
        return self.fc2(F.relu(self.fc1(x.flatten(1)), inplace=True))


def num_parameters(module):
    return sum(p.numel() for p in module.parameters())
