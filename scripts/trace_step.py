"""Per-launch trace of one train step (bench configuration): prints every launch with its inter-event time."""
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
agg = capi.trace(lambda: tr.step(*sets[0], True))
lst = capi.trace_list()
want = sys.argv[1] if len(sys.argv) > 1 else "conv"
print("total ms", sum(v[1] for v in agg.values()))
for i, (n, ms) in enumerate(lst):
    if want in n or want == "all":
        print(i, n, f"{ms * 1e3:.1f}us")
