import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from test_gpu_convglu import ARCH, TorchConvGlu
from wav2letter_b200.trainer import Trainer
B, T, L, N = int(sys.argv[1]) if len(sys.argv) > 1 else 3, 50, 5, 6
tr = Trainer(ARCH, 40, N, "asg", "target_sz_sqrt", transdiag=1.0, lr=0.0, lrcrit=0.0)
g = torch.Generator(device="cuda").manual_seed(2)
feat = torch.randn((B, 1, 40, T), device="cuda", generator=g)
tgt = torch.randint(0, N, (B, L), device="cuda", generator=g, dtype=torch.int32)
flat0 = tr.get_flat(0, 0).clone()
loss = tr.step(feat, tgt, train=True)
torch.cuda.synchronize()
grads = tr.get_flat(0, 1)
ref = TorchConvGlu(flat0, tr.layout(0))
logits = ref.forward(feat)
trans = tr.get_flat(1, 0).view(N, N).cpu().numpy()
ol, ode, odt = oracle.asg(logits.detach().float().cpu().numpy(), tgt.cpu().numpy(), trans, "target_sz_sqrt")
logits.backward(torch.from_numpy(ode).double().cuda())
for (off, n, dims), p in zip(tr.layout(0), ref.p):
    mine = grads[off:off + n].double()
    den = float(p.grad.abs().max())
    print(dims, "err", float((mine - p.grad.flatten()).abs().max()) / max(den, 1e-12), "scale", den, "mine scale", float(mine.abs().max()))
