"""Chain-kernel time per role at the BASELINE point (B=64, T=1500, N=30, L<=250): the chains kernel is launched with
one, two or four recursions per utterance by choosing the terms / need_grad of the call."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wav2letter_b200 as w
from wav2letter_b200 import capi
from bench import make_asg_inputs, ASG_CFG
cfg = dict(ASG_CFG)
if len(sys.argv) > 1:
    cfg["B"] = int(sys.argv[1])
e, tr, y = make_asg_inputs(np.random.default_rng(0), cfg["B"], cfg["T"], cfg["N"], cfg["L"])
de, dt, dy = torch.from_numpy(e).cuda(), torch.from_numpy(tr).cuda(), torch.from_numpy(y).cuda()
out = {}
for name, terms, grad in [("fcc_alpha", w.TERM_FCC, False), ("fac_alpha", w.TERM_FAC, False), ("fcc_alpha+beta", w.TERM_FCC, True),
                          ("fac_alpha+beta", w.TERM_FAC, True), ("asg_alpha_only", w.TERM_ASG, False), ("asg_all", w.TERM_ASG, True)]:
    for _ in range(3):
        w.asg_forward_backward(de, dy, dt, cfg["scale_mode"], None, terms, need_grad=grad)
    prof = capi.ProfileList(2, 8)
    prof.arm()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(8):
        w.asg_forward_backward(de, dy, dt, cfg["scale_mode"], None, terms, need_grad=grad)
    ev1.record()
    torch.cuda.synchronize()
    used = prof.disarm()
    ks = prof.times_ms(used)
    out[name] = {"chains_kernel_us": round(1e3 * sum(ks) / len(ks), 1), "call_us": round(1e3 * ev0.elapsed_time(ev1) / 8, 1),
                 "ns_per_step": round(1e6 * sum(ks) / len(ks) / cfg["T"], 1)}
print(json.dumps(out, indent=1))
