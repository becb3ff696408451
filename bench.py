#!/usr/bin/env python
"""bench.py — measurement contract of the repo (see the task statement, "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload asg]

Workloads
  asg   ASG criterion forward+backward at BASELINE.json's microbench point T=1500, N=30, B=64 per
        GPU (weak scaling: every rank owns its own batch of utterances; the path shards by
        utterance with no data-path collective).  A step = one fused
        w2l_asg_forward_backward call over one synthetic batch.

One JSON line on rank 0.  `value` = frames/s with inputs resident in HBM; `e2e` = the same metric
through the C ABI with HOST (pinned) buffers, H2D of emissions/targets and D2H of the losses inside
the timed region; `roofline` = achieved algorithmic GB/s of the dominant kernel (asg_chains_kernel,
timed live with CUDA events through w2l_set_profile_events) against MEASURED_PEAKS.json;
`cpu_baseline` = the oracle (a port of flashlight-0.3's CPU criterion, OpenMP over the batch like
upstream) timed on this box's host cores.  `--impl reference` times that CPU implementation as the
reference arm (the reference's own ArrayFire-CPU backend cannot be built here: DESIGN.md §oracle).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ASG_CFG = dict(T=1500, N=30, B=64, L=250, scale_mode="target_sz_sqrt", n_input_sets=16)


def asg_algorithmic_bytes(B, T, N, L):
    """read emis + write d_emis + trans/d_trans + targets + losses."""
    return 8 * B * T * N + 8 * N * N + 4 * B * L + 4 * B


def make_asg_inputs(rng, B, T, N, L):
    e = (rng.standard_normal((B, T, N), dtype=np.float32) * 3).astype(np.float32)
    tr = (4 * np.eye(N) + rng.normal(0, 0.1, (N, N))).astype(np.float32)
    y = rng.integers(0, N, (B, L)).astype(np.int32)
    lens = rng.integers(T // 8, T // 5 + 1, B)
    lens = np.minimum(lens, L)
    for b in range(B):
        y[b, lens[b]:] = -1
    return e, tr, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_asg(sample_B, cfg, threads=None, reps=3):
    """oracle timed on the host cores; returns frames/s and a description."""
    import oracle

    if threads:
        oracle.set_num_threads(threads)
    rng = np.random.default_rng(99)
    e, tr, y = make_asg_inputs(rng, sample_B, cfg["T"], cfg["N"], cfg["L"])
    oracle.asg(e[:2], y[:2], tr, cfg["scale_mode"])  # warm
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.asg(e, y, tr, cfg["scale_mode"])
        best = min(best, time.perf_counter() - t0)
    return sample_B * cfg["T"] / best, best, oracle.num_threads()


def run_reference(args, rank, world):
    """Reference arm: the CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    import oracle

    cfg = dict(ASG_CFG)
    cores = os.cpu_count() or 1
    oracle.set_num_threads(cores)
    rng = np.random.default_rng(1234)
    e, tr, y = make_asg_inputs(rng, cfg["B"], cfg["T"], cfg["N"], cfg["L"])
    for _ in range(max(1, args.warmup)):
        oracle.asg(e, y, tr, cfg["scale_mode"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.asg(e, y, tr, cfg["scale_mode"])
    dt = time.perf_counter() - t0
    fps = cfg["B"] * cfg["T"] * args.steps / dt
    line = {
        "impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 accumulation, as upstream's CPU backend)",
        "data": "synthetic",
        "config": {"workload": "ASG criterion fwd+bwd, T=1500 N=30 B=64 L<=250 (BASELINE.json configs[4] point)",
                   "note": "reference's ArrayFire-CPU backend is unbuildable here; this is the oracle port of "
                           "flashlight-0.3 lib/sequence/criterion/cpu, OpenMP over the batch like upstream"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": oracle.num_threads(), "kind": "port",
                         "sample": f"{args.steps} full batches of B=64,T=1500,N=30"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_asg(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import wav2letter_b200 as w
    from wav2letter_b200 import capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = dict(ASG_CFG)
    B, T, N, L = cfg["B"], cfg["T"], cfg["N"], cfg["L"]
    rng = np.random.default_rng(1234 + rank)
    nsets = cfg["n_input_sets"]
    host_e, host_y, dev_e, dev_y = [], [], [], []
    tr_np = None
    for _ in range(nsets):
        e, tr_np, y = make_asg_inputs(rng, B, T, N, L)
        he = torch.from_numpy(e).pin_memory()
        hy = torch.from_numpy(y).pin_memory()
        host_e.append(he)
        host_y.append(hy)
        dev_e.append(he.to(dev))
        dev_y.append(hy.to(dev))
    trans = torch.from_numpy(tr_np).to(dev)
    loss = torch.empty(B, dtype=torch.float32, device=dev)
    d_emis = torch.empty((B, T, N), dtype=torch.float32, device=dev)
    d_trans = torch.empty((N, N), dtype=torch.float32, device=dev)
    ws = torch.empty(capi.lib.w2l_asg_workspace_size(B, T, N, L), dtype=torch.uint8, device=dev)
    stage_e = torch.empty((B, T, N), dtype=torch.float32, device=dev)
    stage_y = torch.empty((B, L), dtype=torch.int32, device=dev)
    host_loss = torch.empty(B, dtype=torch.float32).pin_memory()

    def step(i):
        k = i % nsets
        w.asg_forward_backward(dev_e[k], dev_y[k], trans, cfg["scale_mode"], out=(loss, d_emis, d_trans), ws=ws)

    def step_e2e(i):
        k = i % nsets
        stage_e.copy_(host_e[k], non_blocking=True)
        stage_y.copy_(host_y[k], non_blocking=True)
        w.asg_forward_backward(stage_e, stage_y, trans, cfg["scale_mode"], out=(loss, d_emis, d_trans), ws=ws)
        host_loss.copy_(loss, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for i in range(warmup):
            fn(i)
        kern_ms = []
        evs = []
        barrier()
        sampler = ClockSampler(local_rank) if profile else None
        if sampler:
            sampler.start()
        w.reset_launch_count()
        t_start = torch.cuda.Event(enable_timing=True)
        t_stop = torch.cuda.Event(enable_timing=True)
        t_start.record()
        for i in range(steps):
            if profile:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                capi.set_profile_events(a, b)
                evs.append((a, b))
            fn(warmup + i)
        t_stop.record()
        capi.set_profile_events(None, None)
        barrier()
        launches = w.launch_count()
        clocks = sampler.stop() if sampler else None
        ms = t_start.elapsed_time(t_stop)
        for a, b in evs:
            kern_ms.append(a.elapsed_time(b))
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, kern_ms, launches, clocks

    ms, kern_ms, launches, clocks = timed(step, args.steps, args.warmup, profile=True)
    ms_e2e, _, _, _ = timed(step_e2e, args.steps, args.warmup)
    frames = B * T * world
    value = frames * args.steps / (ms * 1e-3)
    e2e = frames * args.steps / (ms_e2e * 1e-3)
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    alg = asg_algorithmic_bytes(B, T, N, L)
    kavg = sum(kern_ms) / len(kern_ms)
    achieved = alg / (kavg * 1e-3) / 1e9
    cpu_fps, cpu_s, cpu_threads = cpu_asg(B, cfg)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "asg_chains_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    line = {
        "metric": "frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ASG criterion fwd+bwd, T=1500 N=30 B=64 L<=250 per GPU (BASELINE.json configs[4] point)",
                   "criterion": "asg", "scale_mode": cfg["scale_mode"], "sharding": f"utterances, dp{world}",
                   "cold_inputs": f"rotating {nsets} input sets ({nsets * B * T * N * 4 / 1e6:.0f} MB) + "
                                  f"{ws.numel() / 1e6:.0f} MB workspace rewritten per step > 126 MB L2"},
        "asg_fwd_bwd_ms_per_batch": ms / args.steps,
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": B * T * N * 4 + B * L * 4, "d2h_bytes_per_step": B * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "asg_chains_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel_ms": kavg, "algorithmic_bytes": alg,
                     "dependent_step_ns": 1e6 * kavg / T,
                     "note": "latency-bound recursion: T dependent steps per utterance; see DESIGN.md"},
        "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                         "sample": f"1 batch B={B},T={T},N={N} (best of 3, {cpu_s:.2f} s)"},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="asg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import torch

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_asg(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
