"""Per-GEMM device times inside a real train step (event pair around every GEMM launch), in launch order."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import TDS_CFG, make_tds_inputs
from wav2letter_b200 import capi
from wav2letter_b200.trainer import SEQ2SEQ_TDS_CTC_ARCH, Trainer
cfg = TDS_CFG
tr = Trainer(SEQ2SEQ_TDS_CTC_ARCH, cfg["F"], cfg["N"], "ctc", "none", lr=cfg["lr"], maxgradnorm=cfg["maxgradnorm"])
sets = []
for k in range(4):
    f, y = make_tds_inputs(np.random.default_rng(k), cfg)
    sets.append((torch.from_numpy(f).cuda(), torch.from_numpy(y).cuda()))
for i in range(4):
    tr.step(*sets[i % 4], True)
torch.cuda.synchronize()
prof = capi.ProfileList(1, 100)
acc = None
n = 5
for i in range(n):
    prof.arm()
    tr.step(*sets[i % 4], True)
    torch.cuda.synchronize()
    used = prof.disarm()
    t = np.array(prof.times_ms(used)) * 1e3
    acc = t if acc is None else acc + t
acc /= n
print("gemms per step", len(acc), "total us", acc.sum())
print(" ".join(f"{v:.0f}" for v in acc))
