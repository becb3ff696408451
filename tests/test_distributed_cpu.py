"""world_size-2 gloo test (CPU) of the data-parallel contract of the hot path: utterances are
sharded across ranks with no data-path collective, per-sample losses are local, and the only reduction is the SUM
of parameter gradients (here: the ASG transition gradient) followed by division by the global batch size
(Train.cpp:1721-1752).  The arithmetic is the oracle's — this checks the host-side sharding / reduction logic the
GPU path (bench.py, trainer) relies on, plus the rendezvous helper that ships the 128-byte NCCL id."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    B, T, N, L = 6, 40, 8, 7
    e = (rng.normal(0, 1, (B, T, N)) * 3).astype(np.float32)
    tr = (2 * np.eye(N) + rng.normal(0, 0.1, (N, N))).astype(np.float32)
    y = rng.integers(0, N, (B, L)).astype(np.int32)
    y[4, 3:] = -1
    # rank r owns utterances r, r+world, ... (dataset sharding of createDataset(..., worldRank, worldSize))
    mine = list(range(rank, B, world))
    loss, de, dtr = oracle.asg(e[mine], y[mine], tr, "target_sz")
    g = torch.from_numpy(dtr.copy())
    dist.all_reduce(g)  # reducer->add(grad) ... finalize(): sum, scale 1.0
    total = torch.tensor([float(len(mine))])
    dist.all_reduce(total)  # fl::allReduce(totalBatchSizeArr)
    g /= total
    # rendezvous payload: rank 0's 128-byte id reaches every rank intact
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    if rank == 0:
        fl, fde, fdtr = oracle.asg(e, y, tr, "target_sz")
        np.save(out, np.stack([g.numpy(), fdtr / B]))
        assert np.allclose(loss, fl[mine]) and np.allclose(de, fde[mine])
    assert uid.tolist() == list(range(128)) and total.item() == B
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_sum(tmp_path):
    out = str(tmp_path / "g.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got, want = np.load(out)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
