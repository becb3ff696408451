import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wav2letter_b200 as w
capi = w.capi
M, C = 9600, 800
A = torch.randn(M, C, device="cuda"); B = torch.randn(C, C, device="cuda"); Cm = torch.zeros(M, C, device="cuda"); aux = torch.randn(M, C, device="cuda")
for kw in (dict(), dict(aux=aux, aux_mode=1, aux_scale=1.25), dict(accumulate=True)):
    for i in range(2):
        capi.gemm_tf32_ex(A, B, Cm, b_mn=True, **kw)
torch.cuda.synchronize()
