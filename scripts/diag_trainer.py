"""Diagnostic: per-parameter gradient error of the C++ trainer vs a float64 torch reference, next to the error
torch's own fp32+TF32 run of the same network shows (TF32 noise floor)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from test_gpu_trainer import ARCH, TorchTDS, make_batch, rel
from wav2letter_b200.trainer import Trainer

N, B, T, L = 12, 3, 64, 5
tr = Trainer(ARCH, 80, N, "ctc", "target_sz", lr=0.0)
feat, tgt = make_batch(B, T, N, L, 1, True)
flat0 = tr.get_flat(0, 0).clone()
tr.step(feat, tgt, train=True)
grads = tr.get_flat(0, 1)
ref = TorchTDS(flat0, tr.layout(0))
logits = ref.forward(feat)
ol, ode = oracle.ctc(logits.detach().float().cpu().numpy(), tgt.cpu().numpy(), "target_sz")
logits.backward(torch.from_numpy(ode).double().cuda())

# torch fp32 with TF32 matmuls on the same graph
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
class T32(TorchTDS):
    def __init__(self, flat, layout):
        self.p = [flat[o:o+n].float().clone().requires_grad_(True) for o, n, d in layout]
r32 = T32(flat0, tr.layout(0))
import torch.nn.functional as F
x32 = feat.float()
# reuse forward but in float32: monkeypatch .double() by running with float tensors
orig_double = torch.Tensor.double
torch.Tensor.double = lambda self: self.float()
try:
    lg32 = r32.forward(feat)
finally:
    torch.Tensor.double = orig_double
lg32.backward(torch.from_numpy(ode).float().cuda())
print(f"{'param dims':28s} {'mine vs f64':>12s} {'torch tf32 vs f64':>18s}")
for (off, n, dims), p, q in zip(tr.layout(0), ref.p, r32.p):
    print(f"{str(dims):28s} {rel(grads[off:off+n], p.grad.flatten()):12.2e} {rel(q.grad.flatten(), p.grad.flatten()):18.2e}")
print("logits: mine", rel(tr.forward(feat), logits), " torch tf32", rel(lg32, logits))
