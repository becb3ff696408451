"""Cost of each GEMM epilogue variant on the stage-1 / stage-3 shapes (isolated launches, rotating operands)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wav2letter_b200 as w
capi = w.capi
nset = 4
for M, C in [(9600, 800), (2400, 1440)]:
    for b_mn in (False, True):
        As = [torch.randn(M, C, device="cuda") for _ in range(nset)]
        Bs = [torch.randn(C, C, device="cuda") for _ in range(nset)]
        Cm = torch.zeros(M, C, device="cuda")
        aux = torch.randn(M, C, device="cuda")
        bias = torch.randn(C, device="cuda")
        variants = {
            "plain": dict(),
            "bias": dict(bias=bias),
            "bias+relu": dict(bias=bias, act=1),
            "bias+relu+drop": dict(bias=bias, act=1, dropout_p=0.2, seed=7),
            "auxmask": dict(aux=aux, aux_mode=1, aux_scale=1.25),
            "accum": dict(accumulate=True),
            "auxmask+accum": dict(aux=aux, aux_mode=1, aux_scale=1.25, accumulate=True),
        }
        for name, kw in variants.items():
            for i in range(3):
                capi.gemm_tf32_ex(As[i % nset], Bs[i % nset], Cm, b_mn=b_mn, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for i in range(n):
                capi.gemm_tf32_ex(As[i % nset], Bs[i % nset], Cm, b_mn=b_mn, **kw)
            e1.record()
            torch.cuda.synchronize()
            print(f"M={M} C={C} b_mn={int(b_mn)} {name:16s} {e0.elapsed_time(e1) / n * 1e3:7.1f} us", flush=True)
