"""FAC parity (d_emis, max abs error / max abs reference) against T at fixed L/T = 1/6: linear growth = a systematic
per-step error, square-root growth = rounding noise."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import wav2letter_b200 as w
from bench import make_asg_inputs
rel = lambda a, b: float(np.abs(a - b).max() / max(1e-20, np.abs(b).max()))
out = []
for T in (188, 375, 750, 1500, 3000):
    errs = []
    for seed in range(3):
        e, tr, y = make_asg_inputs(np.random.default_rng(7 + seed), 16, T, 30, max(2, T // 6))
        l, de, dt = w.asg_forward_backward(torch.from_numpy(e).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(tr).cuda(), "none", None, w.TERM_FAC)
        torch.cuda.synchronize()
        ol, ode, odt = oracle.fac(e, y, tr, "none")
        errs.append(rel(de.cpu().numpy(), ode))
    out.append({"T": T, "d_emis_err": errs})
print(json.dumps(out))
