"""Oracle parity of the fused ASG call at the BASELINE point for several seeds (max abs error over max abs reference),
for the whole criterion and for each of its two terms."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import wav2letter_b200 as w
from bench import make_asg_inputs, ASG_CFG
cfg = dict(ASG_CFG)
out = []
rel = lambda a, b: float(np.abs(a - b).max() / max(1e-20, np.abs(b).max()))
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    e, tr, y = make_asg_inputs(np.random.default_rng(100 + seed), cfg["B"], cfg["T"], cfg["N"], cfg["L"])
    de_, dy_, dt_ = torch.from_numpy(e).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(tr).cuda()
    rec = {"seed": seed}
    for name, terms, fn in (("asg", w.TERM_ASG, lambda: oracle.asg(e, y, tr, cfg["scale_mode"])),
                            ("fac", w.TERM_FAC, lambda: oracle.fac(e, y, tr, cfg["scale_mode"])),
                            ("fcc", w.TERM_FCC, lambda: oracle.fcc(e, tr, cfg["scale_mode"], target=y))):
        l, de, dt = w.asg_forward_backward(de_, dy_, dt_, cfg["scale_mode"], None, terms)
        torch.cuda.synchronize()
        ol, ode, odt = fn()
        rec[name] = {"loss": rel(l.cpu().numpy(), ol), "d_emis": rel(de.cpu().numpy(), ode), "d_trans": rel(dt.cpu().numpy(), odt),
                     "d_emis_ref_max": float(np.abs(ode).max())}
    out.append(rec)
print(json.dumps([{"seed": r["seed"], **{k: r[k]["d_emis"] for k in ("asg", "fac", "fcc")}} for r in out]))
