"""CPU port of the acoustic-model train step — TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product).

The reference's ArrayFire-CPU backend cannot be built here (no flashlight / ArrayFire sources or binaries:
SURVEY.md §0, §8c), so BASELINE.md §4 names torch-CPU (oneDNN) fp32 as the stand-in for the acoustic-model
operators and the C oracle for the criterion.  This module restates the module semantics of the arch DSL
(recipes/joint_training_vox_populi/cpc/SequentialBuilder.cpp:92-626; TDSBlock: fl/contrib TDSBlock as
documented in tools/StreamingTDSModelConverter.cpp:103-136) with plain torch CPU ops, and runs one
forward + criterion + backward + clip + SGD step: the `cpu_baseline` / `--impl reference` legs of bench.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ctc as oracle_ctc


def same_pad(T: int, k: int, s: int) -> int:
    """flashlight Conv2D SAME padding (symmetric)"""
    rem = T % s
    tot = (k - 1) - (s if rem == 0 else rem) + 1
    return max((tot + 1) // 2, 0)


class TimeConv(torch.nn.Module):
    def __init__(self, cin, cout, k, stride, relu=False, drop=0.0):
        super().__init__()
        b = math.sqrt(3.0 / (cin * k))
        self.w = torch.nn.Parameter(torch.empty(cout, cin, k, 1).uniform_(-b, b))
        self.b = torch.nn.Parameter(torch.empty(cout).uniform_(-b, b))
        self.k, self.s, self.relu, self.drop = k, stride, relu, drop

    def forward(self, x):  # x [B, C, T, W]
        p = same_pad(x.shape[2], self.k, self.s)
        y = F.conv2d(F.pad(x, (0, 0, p, p)), self.w, self.b, stride=(self.s, 1))
        if self.relu:
            y = F.relu(y)
        return F.dropout(y, self.drop, self.training)


class SampleLN(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.g = torch.nn.Parameter(torch.ones(1))
        self.b = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return F.layer_norm(x, x.shape[1:], eps=1e-5) * self.g + self.b


class TDS(torch.nn.Module):
    def __init__(self, c, k, w, drop):
        super().__init__()
        self.conv = TimeConv(c, c, k, 1, relu=True, drop=drop)
        self.ln1, self.ln2 = SampleLN(), SampleLN()
        self.l1 = torch.nn.Linear(c * w, c * w)
        self.l2 = torch.nn.Linear(c * w, c * w)
        self.drop = drop

    def forward(self, x):  # [B, C, T, W]
        z = self.ln1(x + self.conv(x))
        B, C, T, W = z.shape
        f = z.permute(0, 2, 1, 3).reshape(B, T, C * W)
        u = F.dropout(F.relu(self.l1(f)), self.drop, self.training)
        u = F.dropout(self.l2(u), self.drop, self.training)
        return self.ln2(z + u.view(B, T, C, W).permute(0, 2, 1, 3))


class Head(torch.nn.Module):
    def __init__(self, nin, nout):
        super().__init__()
        self.l = torch.nn.Linear(nin, nout)

    def forward(self, x):
        B, C, T, W = x.shape
        return self.l(x.permute(0, 2, 1, 3).reshape(B, T, C * W))


def build(arch_text: str, n_feat: int, n_label: int) -> torch.nn.Sequential:
    mods, last = [], None
    for line in arch_text.splitlines():
        line = line.split("#")[0].replace("NFEAT", str(n_feat)).replace("NLABEL", str(n_label)).split()
        if not line:
            continue
        op = line[0]
        if op in ("V", "RO", "SAUG"):
            last = None
        elif op == "C2":
            last = TimeConv(int(line[1]), int(line[2]), int(line[3]), int(line[5]))
            mods.append(last)
        elif op == "R":
            last.relu = True
        elif op == "DO":
            last.drop = float(line[1])
            last = None
        elif op == "LN":
            mods.append(SampleLN())
        elif op == "TDS":
            mods.append(TDS(int(line[1]), int(line[2]), int(line[3]), float(line[4]) if len(line) > 4 else 0.0))
        elif op == "L":
            mods.append(Head(int(line[1]), int(line[2])))
        else:
            raise ValueError(f"opcode {op} not covered by the CPU port")
    return torch.nn.Sequential(*mods)


class CpuTrainer:
    """forward + CTC (oracle) + backward + clipGradNorm + SGD on the host cores."""

    def __init__(self, arch_text, n_feat, n_label, lr=0.05, momentum=0.0, maxgradnorm=0.0, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.net = build(arch_text, n_feat, n_label).train()
        self.opt = torch.optim.SGD(self.net.parameters(), lr=lr, momentum=momentum)
        self.maxgradnorm = maxgradnorm

    def step(self, feat: np.ndarray, target: np.ndarray, scale_mode="none") -> float:
        """feat [B,1,F,T] (ArrayFire [T,F,1,B]), target [B,L] int32"""
        x = torch.from_numpy(feat).permute(0, 1, 3, 2)  # [B, C=1, T, W=F]
        emis = self.net(x)
        loss, d_emis = oracle_ctc(emis.detach().numpy(), target, scale_mode)
        self.opt.zero_grad(set_to_none=True)
        emis.backward(torch.from_numpy(d_emis) / feat.shape[0])
        if self.maxgradnorm > 0:
            torch.nn.utils.clip_grad_norm_(self.net.parameters(), self.maxgradnorm)
        self.opt.step()
        return float(loss.sum())
