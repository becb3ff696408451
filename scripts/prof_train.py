"""Small driver for ncu: a few TDS+CTC train steps at the bench configuration."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import TDS_CFG, make_tds_inputs
from wav2letter_b200.trainer import SEQ2SEQ_TDS_CTC_ARCH, Trainer
cfg = TDS_CFG
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tr = Trainer(SEQ2SEQ_TDS_CTC_ARCH, cfg["F"], cfg["N"], "ctc", "none", lr=cfg["lr"], maxgradnorm=cfg["maxgradnorm"])
f, y = make_tds_inputs(np.random.default_rng(0), cfg)
f, y = torch.from_numpy(f).cuda(), torch.from_numpy(y).cuda()
for _ in range(n):
    tr.step(f, y, True)
torch.cuda.synchronize()
print("done")
