"""Small driver for ncu: a few fused ASG calls at the BASELINE point (B=64,T=1500,N=30)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wav2letter_b200 as w
from bench import make_asg_inputs, ASG_CFG
cfg = ASG_CFG
terms = {"asg": w.TERM_ASG, "fcc": w.TERM_FCC, "fac": w.TERM_FAC}[sys.argv[1] if len(sys.argv) > 1 else "asg"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
e, tr, y = make_asg_inputs(np.random.default_rng(0), cfg["B"], cfg["T"], cfg["N"], cfg["L"])
de, dt, dy = torch.from_numpy(e).cuda(), torch.from_numpy(tr).cuda(), torch.from_numpy(y).cuda()
for _ in range(n):
    w.asg_forward_backward(de, dy, dt, cfg["scale_mode"], None, terms)
torch.cuda.synchronize()
print("done")
