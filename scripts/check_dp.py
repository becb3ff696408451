"""torchrun --nproc-per-node 2 scripts/check_dp.py — on 2 GPUs: the all-reduced gradient arena of a 2-rank step equals
the gradient of the concatenated batch on one rank (dropout off), and parameters stay in sync after SGD."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_trainer import ARCH, make_batch
from wav2letter_b200.trainer import Trainer, init_distributed, nccl_unique_id

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
N, B, T, L = 12, 4, 64, 5
feat, tgt = make_batch(B * world, T, N, L, 11, True)           # same seed on every rank -> identical global batch
single = Trainer(ARCH, 80, N, "ctc", "none", lr=0.1, maxgradnorm=1.0)
flat0 = single.get_flat(0, 0).clone()
dist.broadcast(flat0, 0)
single.set_flat(flat0)
single.step(feat, tgt, True, total_batch=B * world)             # no communicator yet: plain single-GPU step
g_single, p_single = single.get_flat(0, 1).clone(), single.get_flat(0, 0).clone()
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.tensor(list(nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
init_distributed(rank, world, bytes(uid.cpu().tolist()))
dp = Trainer(ARCH, 80, N, "ctc", "none", lr=0.1, maxgradnorm=1.0)
dp.set_flat(flat0)
dp.sync_parameters()
sl = slice(rank * B, (rank + 1) * B)
dp.step(feat[sl].contiguous(), tgt[sl].contiguous(), True, total_batch=B * world)
g_dp, p_dp = dp.get_flat(0, 1), dp.get_flat(0, 0)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
eg, ep = rel(g_dp, g_single), rel(p_dp, p_single)
print(f"rank {rank}: grad rel diff {eg:.2e}  param rel diff {ep:.2e}", flush=True)
assert eg < 2e-3 and ep < 1e-4, (eg, ep)
dist.barrier()
if rank == 0:
    print("DP-CHECK-OK")
dist.destroy_process_group()
