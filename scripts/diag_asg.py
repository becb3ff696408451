"""GPU diagnostic (not a test): error profile and timing of the ASG terms vs the oracle."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import wav2letter_b200 as w

def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()

def run(B, T, N, L, seed=7, escale=3.0):
    rng = np.random.default_rng(seed)
    e = (rng.normal(0, 1, (B, T, N)) * escale).astype(np.float32)
    tr = (4 * np.eye(N) + rng.normal(0, 0.1, (N, N))).astype(np.float32)
    y = rng.integers(0, N, (B, L)).astype(np.int32)
    for b in range(B):
        n = int(rng.integers(max(1, L // 2), L + 1)); y[b, n:] = -1
    for name, terms in (("FCC", w.TERM_FCC), ("FAC", w.TERM_FAC), ("ASG", w.TERM_ASG)):
        if terms == w.TERM_FCC: ol, ode, odt = oracle.fcc(e, tr, "none", target=y)
        elif terms == w.TERM_FAC: ol, ode, odt = oracle.fac(e, y, tr, "none")
        else: ol, ode, odt = oracle.asg(e, y, tr, "none")
        de, dy, dt = dev(e), dev(y), dev(tr)
        for _ in range(2): gl, gde, gdt = w.asg_forward_backward(de, dy, dt, "none", None, terms)
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): w.asg_forward_backward(de, dy, dt, "none", None, terms)
        b_.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b_) / 5
        gl, gde, gdt = gl.cpu().numpy(), gde.cpu().numpy(), gdt.cpu().numpy()
        err = np.abs(gde - ode)
        bt = err.max(axis=2)
        wb, wt = np.unravel_index(bt.argmax(), bt.shape)
        per_t = bt.max(axis=0)
        q = [float(per_t[int(k * (T - 1) / 8)]) for k in range(9)]
        print(f"{name} B={B} T={T} L={L} sc={escale}: {ms:.3f} ms | loss rel {np.max(np.abs(gl-ol)/np.maximum(1,np.abs(ol))):.2e} "
              f"d_emis abs {err.max():.2e} (ref max {np.abs(ode).max():.2f}) at b={wb} t={wt} | d_trans rel "
              f"{np.abs(gdt-odt).max()/max(1e-3,np.abs(odt).max()):.2e}")
        print("    err by t (9 probes):", " ".join(f"{v:.1e}" for v in q), " frames>1e-4:", int((per_t > 1e-4).sum()))

if __name__ == "__main__":
    run(4, 200, 30, 40)
    run(4, 1500, 30, 100)
    run(4, 1500, 30, 250)
    run(16, 1500, 30, 250)
    run(64, 1500, 30, 250)
