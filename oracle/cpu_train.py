"""CPU port of the acoustic-model train step — TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product).

The reference's ArrayFire-CPU backend cannot be built here (no flashlight / ArrayFire sources or binaries:
SURVEY.md §0, §8c), so BASELINE.md §4 names torch-CPU (oneDNN) fp32 as the stand-in for the acoustic-model operators
and the C oracle for the criterion.  This module runs one forward + criterion + backward + clipGradNorm + SGD step of
any of the BASELINE archs (oracle/am_ref.py's restatement of the arch DSL, in float32 on the host cores): the
`cpu_baseline` / `--impl reference` legs of bench.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import am_ref
from . import asg as oracle_asg
from . import ctc as oracle_ctc


class CpuTrainer:
    def __init__(self, arch_text, n_feat, n_label, criterion="ctc", scale_mode="none", transdiag=0.0, lr=0.05, lrcrit=0.0,
                 momentum=0.0, maxgradnorm=0.0, threads=None, seed=0):
        if threads:
            torch.set_num_threads(int(threads))
        arch = am_ref.zero_dropout(arch_text)  # dropout masks cost no arithmetic worth timing; SpecAugment neither
        self.net = am_ref.RefNet(arch, n_feat, n_label, params=am_ref.init_params(arch, n_feat, n_label, seed), dtype=torch.float32)
        self.criterion, self.scale_mode, self.N = criterion, scale_mode, n_label
        self.trans = (transdiag * np.eye(n_label)).astype(np.float32) if criterion == "asg" else None
        self.opt = torch.optim.SGD(self.net.params, lr=lr, momentum=momentum)
        self.lrcrit, self.maxgradnorm = lrcrit, maxgradnorm

    def step(self, feat: np.ndarray, target: np.ndarray) -> float:
        """feat [B,1,F,T] (ArrayFire [T,F,1,B]), target [B,L] int32"""
        emis = self.net.forward(torch.from_numpy(feat))
        e = emis.detach().numpy()
        if self.criterion == "ctc":
            loss, d_emis = oracle_ctc(e, target, self.scale_mode)
            d_trans = None
        else:
            loss, d_emis, d_trans = oracle_asg(e, target, self.trans, self.scale_mode)
        self.opt.zero_grad(set_to_none=True)
        B = feat.shape[0]
        emis.backward(torch.from_numpy(d_emis) / B)
        if self.maxgradnorm > 0:
            torch.nn.utils.clip_grad_norm_(self.net.params, self.maxgradnorm)
        self.opt.step()
        if d_trans is not None and self.lrcrit:
            self.trans -= self.lrcrit * d_trans / B
        return float(np.nansum(loss))
