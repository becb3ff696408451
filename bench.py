#!/usr/bin/env python
"""bench.py — measurement contract of the repo (task statement, "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload tds_ctc|asg]

Workloads
  tds_ctc (default) BASELINE.json configs[1]: the seq2seq_tds LibriSpeech TDS acoustic model with a CTC head,
          fp32 storage, one full TRAIN STEP per step — network forward, CTC, backward, NCCL gradient all-reduce
          (N > 1), division by the global batch, clipGradNorm, SGD — on a synthetic batch of B=16 utterances x
          T=1200 filterbank frames (80 bins) per GPU, 10 000 word-piece classes, targets of <= 60 tokens
          (recipes/seq2seq_tds/librispeech/train.cfg: batchsize 16, lr 0.05, maxgradnorm 15, 10k word pieces).
          Data parallel by utterance (weak scaling), gradient all-reduce is the only collective.
  asg     BASELINE.json configs[4] point: fused ASG forward+backward at T=1500, N=30, B=64 per GPU.
The default run also measures the ASG point briefly and reports it as `asg_fwd_bwd_ms_per_batch`.

One JSON line on rank 0.  `value` = frames/s with the batch resident in HBM; `e2e` = the same step through the
C ABI with HOST (pinned) batches, H2D of features/targets and D2H of the losses inside the timed region;
`roofline` = the dominant kernel (tds_ctc: gemm_tf32_kernel, achieved TFLOP/s over all GEMM launches of the
timed steps; asg: asg_chains_kernel GB/s) timed live with CUDA events through w2l_set_profile_event_list against
MEASURED_PEAKS.json; `cpu_baseline` = the CPU port (oracle/) timed on this box's host cores.
`--impl reference` times that CPU implementation as the reference arm (the reference's own ArrayFire-CPU backend
cannot be built here: DESIGN.md §2).
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
TDS_CFG = dict(B=16, T=1200, F=80, N=10000, L=60, lr=0.05, momentum=0.0, maxgradnorm=15.0, n_input_sets=8)


def asg_algorithmic_bytes(B, T, N, L):
    """read emis + write d_emis + trans/d_trans + targets + losses."""
    return 8 * B * T * N + 8 * N * N + 4 * B * L + 4 * B


def tds_gemm_flops(arch_text, B, T, n_label):
    """2*M*N*K summed over every Linear of the arch, x3 (forward, data gradient, weight gradient)."""
    total, t = 0, T
    for line in arch_text.splitlines():
        p = line.split("#")[0].replace("NLABEL", str(n_label)).split()
        if not p:
            continue
        if p[0] == "C2":
            s, k = int(p[5]), int(p[3])
            rem = t % s
            pad = max(((k - 1) - (s if rem == 0 else rem) + 1 + 1) // 2, 0)
            t = (t + 2 * pad - k) // s + 1
        elif p[0] == "TDS":
            d = int(p[1]) * int(p[3])
            total += 2 * (2 * B * t * d * d)
        elif p[0] == "L":
            total += 2 * B * t * int(p[1]) * int(p[2])
    return 3 * total, t


def make_asg_inputs(rng, B, T, N, L):
    e = (rng.standard_normal((B, T, N), dtype=np.float32) * 3).astype(np.float32)
    tr = (4 * np.eye(N) + rng.normal(0, 0.1, (N, N))).astype(np.float32)
    y = rng.integers(0, N, (B, L)).astype(np.int32)
    lens = rng.integers(T // 8, T // 5 + 1, B)
    lens = np.minimum(lens, L)
    for b in range(B):
        y[b, lens[b]:] = -1
    return e, tr, y


def make_tds_inputs(rng, cfg, B=None):
    """features [B,1,F,T] (== ArrayFire [T,F,1,B]): x ~ N(0,1) (post-LocalNorm statistics); word-piece targets."""
    B = B or cfg["B"]
    feat = rng.standard_normal((B, 1, cfg["F"], cfg["T"]), dtype=np.float32)
    tgt = rng.integers(0, cfg["N"] - 1, (B, cfg["L"])).astype(np.int32)
    lens = rng.integers(cfg["L"] // 2, cfg["L"] + 1, B)
    for b in range(B):
        tgt[b, lens[b]:] = -1
    return feat, tgt


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
            time.sleep(0.25)
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
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
# reference arm / CPU baselines (oracle/ is only ever used here, in tests/ and in smoke())
# ---------------------------------------------------------------------------------------------------------
def cpu_asg(sample_B, cfg, reps=3):
    import oracle

    rng = np.random.default_rng(99)
    e, tr, y = make_asg_inputs(rng, sample_B, cfg["T"], cfg["N"], cfg["L"])
    oracle.asg(e[:2], y[:2], tr, cfg["scale_mode"])
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.asg(e, y, tr, cfg["scale_mode"])
        best = min(best, time.perf_counter() - t0)
    return sample_B * cfg["T"] / best, best, oracle.num_threads()


def cpu_tds(sample_B, steps, warmup):
    """torch-CPU (oneDNN) + oracle CTC train step of the same arch on a bounded sample of the workload."""
    import torch

    from oracle.tds_cpu import CpuTrainer
    from wav2letter_b200.trainer import SEQ2SEQ_TDS_CTC_ARCH

    cfg = TDS_CFG
    # more than ~32 threads thrash on the many-core host (oneDNN + OpenMP oversubscription: 47 s/step with 128
    # threads against 3.4 s/step with 8 in the dev container); report the threads actually used
    cores = min(os.cpu_count() or 1, 32)
    import oracle

    oracle.set_num_threads(cores)
    tr = CpuTrainer(SEQ2SEQ_TDS_CTC_ARCH, cfg["F"], cfg["N"], cfg["lr"], cfg["momentum"], cfg["maxgradnorm"], threads=cores)
    rng = np.random.default_rng(4321)
    feat, tgt = make_tds_inputs(rng, cfg, sample_B)
    for _ in range(warmup):
        tr.step(feat, tgt)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(feat, tgt)
    dt = time.perf_counter() - t0
    return sample_B * cfg["T"] * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args, rank, world):
    """Reference arm: the CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    if args.workload == "asg":
        import oracle

        cfg = dict(ASG_CFG)
        oracle.set_num_threads(os.cpu_count() or 1)
        rng = np.random.default_rng(1234)
        e, tr, y = make_asg_inputs(rng, cfg["B"], cfg["T"], cfg["N"], cfg["L"])
        for _ in range(max(1, args.warmup)):
            oracle.asg(e, y, tr, cfg["scale_mode"])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oracle.asg(e, y, tr, cfg["scale_mode"])
        dt = (time.perf_counter() - t0) / args.steps
        fps, threads = cfg["B"] * cfg["T"] / dt, oracle.num_threads()
        workload = "ASG criterion fwd+bwd, T=1500 N=30 B=64 L<=250 (BASELINE.json configs[4] point)"
        sample = f"{args.steps} full batches of B=64,T=1500,N=30"
    else:
        sample_B = 2
        steps = max(1, min(args.steps, 4))
        fps, dt, threads = cpu_tds(sample_B, steps, min(args.warmup, 1))
        workload = ("seq2seq_tds TDS + CTC train step, fp32, T=1200 F=80 N=10000 (BASELINE.json configs[1]); "
                    f"bounded sample B={sample_B} per step")
        sample = f"{steps} train steps of B={sample_B},T=1200 (the GPU arm runs B=16 per GPU)"
    line = {
        "impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload,
                   "note": "the reference's ArrayFire-CPU backend is unbuildable here (SURVEY.md §0); this is the CPU port: "
                           "torch-CPU/oneDNN fp32 for the acoustic-model operators + the C oracle (flashlight-0.3 "
                           "lib/sequence/criterion/cpu restated) for the criterion, all host threads"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# B200 arms
# ---------------------------------------------------------------------------------------------------------
class Timed:
    def __init__(self, world, local_rank):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.world, self.dev = torch, dist, world, torch.device("cuda", local_rank)
        self.local_rank = local_rank

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, steps, warmup, sample_clocks=False, profile=None, profile_steps=3):
        """W warm-up steps, then exactly K steps between barrier+synchronize on both sides; device time, max over ranks."""
        import wav2letter_b200 as w

        torch = self.torch
        for i in range(warmup):
            fn(i)
        self.barrier()
        sampler = ClockSampler(self.local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        w.reset_launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        used = 0
        t0.record()
        for i in range(steps):
            if profile is not None and i == 0:
                profile.arm()
            fn(warmup + i)
            if profile is not None and i == profile_steps - 1:
                used = profile.disarm()
        if profile is not None and steps < profile_steps:
            used = profile.disarm()
        t1.record()
        self.barrier()
        launches = w.launch_count()
        clocks = sampler.stop() if sampler else None
        ms = t0.elapsed_time(t1)
        if self.world > 1:
            tms = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(tms, op=self.dist.ReduceOp.MAX)
            ms = float(tms.item())
        kern = profile.times_ms(used) if profile is not None else []
        return ms, kern, launches, clocks


def asg_point(tm: "Timed", rank, steps, warmup, profile=True):
    """fused ASG forward+backward at the BASELINE point; returns a dict of measurements."""
    import torch

    import wav2letter_b200 as w
    from wav2letter_b200 import capi

    cfg = dict(ASG_CFG)
    B, T, N, L = cfg["B"], cfg["T"], cfg["N"], cfg["L"]
    rng = np.random.default_rng(1234 + rank)
    nsets = cfg["n_input_sets"]
    host_e, host_y, dev_e, dev_y = [], [], [], []
    tr_np = None
    for _ in range(nsets):
        e, tr_np, y = make_asg_inputs(rng, B, T, N, L)
        he, hy = torch.from_numpy(e).pin_memory(), torch.from_numpy(y).pin_memory()
        host_e.append(he)
        host_y.append(hy)
        dev_e.append(he.to(tm.dev))
        dev_y.append(hy.to(tm.dev))
    trans = torch.from_numpy(tr_np).to(tm.dev)
    loss = torch.empty(B, dtype=torch.float32, device=tm.dev)
    d_emis = torch.empty((B, T, N), dtype=torch.float32, device=tm.dev)
    d_trans = torch.empty((N, N), dtype=torch.float32, device=tm.dev)
    ws = torch.empty(capi.lib.w2l_asg_workspace_size(B, T, N, L), dtype=torch.uint8, device=tm.dev)
    stage_e = torch.empty((B, T, N), dtype=torch.float32, device=tm.dev)
    stage_y = torch.empty((B, L), dtype=torch.int32, device=tm.dev)
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

    prof = capi.ProfileList(2, 8) if profile else None
    ms, kern, launches, clocks = tm.run(step, steps, warmup, sample_clocks=profile, profile=prof, profile_steps=min(steps, 8))
    ms_e2e, _, _, _ = tm.run(step_e2e, steps, warmup)
    kavg = sum(kern) / len(kern) if kern else None
    return dict(cfg=cfg, ms=ms / steps, ms_e2e=ms_e2e / steps, kernel_ms=kavg, launches=launches, clocks=clocks,
                ws_mb=ws.numel() / 1e6, h2d=B * T * N * 4 + B * L * 4, d2h=B * 4)


def run_asg(args, rank, world, local_rank):
    tm = Timed(world, local_rank)
    r = asg_point(tm, rank, args.steps, args.warmup)
    if rank != 0:
        return
    cfg = r["cfg"]
    B, T, N, L = cfg["B"], cfg["T"], cfg["N"], cfg["L"]
    peaks, src = measured_peaks()
    alg = asg_algorithmic_bytes(B, T, N, L)
    achieved = alg / (r["kernel_ms"] * 1e-3) / 1e9
    cpu_fps, cpu_s, cpu_threads = cpu_asg(B, cfg)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "asg_chains_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    frames = B * T * world
    line = {
        "metric": "frames_per_sec", "value": frames / (r["ms"] * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ASG criterion fwd+bwd, T=1500 N=30 B=64 L<=250 per GPU (BASELINE.json configs[4] point)",
                   "criterion": "asg", "scale_mode": cfg["scale_mode"], "sharding": f"utterances, dp{world}",
                   "cold_inputs": f"rotating {cfg['n_input_sets']} input sets ({cfg['n_input_sets'] * B * T * N * 4 / 1e6:.0f} MB) + "
                                  f"{r['ws_mb']:.0f} MB workspace rewritten per step > 126 MB L2"},
        "asg_fwd_bwd_ms_per_batch": r["ms"],
        "clocks": r["clocks"],
        "e2e": {"value": frames / (r["ms_e2e"] * 1e-3), "unit": "frames/s", "ms_per_step": r["ms_e2e"],
                "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
        "gpu_launches": r["launches"],
        "roofline": {"bound": "hbm", "kernel": "asg_chains_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": src,
                     "kernel_ms": r["kernel_ms"], "algorithmic_bytes": alg, "dependent_step_ns": 1e6 * r["kernel_ms"] / T,
                     "note": "latency-bound recursion: T dependent steps per utterance; see DESIGN.md"},
        "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                         "sample": f"1 batch B={B},T={T},N={N} (best of 3, {cpu_s:.2f} s)"},
    }
    print(json.dumps(line), flush=True)


def conv_glu_flops(arch_text, B, T, n_feat, n_label):
    """2*M*N*K over every convolution / Linear of a conv_glu arch, x3 (forward, data gradient, weight gradient)."""
    total, t = 0, T
    for line in arch_text.splitlines():
        p = line.replace("NFEAT", str(n_feat)).replace("NLABEL", str(n_label)).split()
        if p and p[0] == "WN" and p[2] == "C":
            cin, cout, kw, pad = int(p[3]), int(p[4]), int(p[5]), int(p[7])
            t = t + 2 * pad - kw + 1
            total += 2 * B * t * cout * cin * kw
        elif p and p[0] == "WN" and p[2] == "L":
            total += 2 * B * t * int(p[3]) * int(p[4])
    return 3 * total, t


def run_conv_glu(args, rank, world, local_rank):
    """conv_glu LibriSpeech 17-layer GLU model + ASG, full train step (BASELINE.json configs[2] shape, TF32 math, 1 GPU per
    rank; not the default workload).  GEMM time is measured with an event pair around every GEMM launch."""
    import torch

    from wav2letter_b200 import capi
    from wav2letter_b200.trainer import Trainer, conv_glu_librispeech_arch

    tm = Timed(world, local_rank)
    B, T, F, N, L = 8, 1000, 40, 30, 160
    arch = conv_glu_librispeech_arch()
    trainer = Trainer(arch, F, N, "asg", "target_sz_sqrt", transdiag=4.0, lr=0.1, lrcrit=0.001, maxgradnorm=0.2)
    rng = np.random.default_rng(1234 + rank)
    sets = []
    for _ in range(4):
        f = rng.standard_normal((B, 1, F, T), dtype=np.float32)
        y = rng.integers(0, N, (B, L)).astype(np.int32)
        sets.append((torch.from_numpy(f).to(tm.dev), torch.from_numpy(y).to(tm.dev)))
    loss = torch.empty(B, dtype=torch.float32, device=tm.dev)

    def step(i):
        trainer.step(sets[i % 4][0], sets[i % 4][1], True, float(B * world), loss)

    prof_steps = min(args.steps, 2)
    prof = capi.ProfileList(1, 1200 * prof_steps)
    ms, kern, launches, clocks = tm.run(step, args.steps, args.warmup, sample_clocks=True, profile=prof, profile_steps=prof_steps)
    if rank != 0:
        return
    flops, t_out = conv_glu_flops(arch, B, T, F, N)
    peaks, src = measured_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    gemm_ms = sum(kern) / prof_steps
    line = {
        "metric": "frames_per_sec", "value": B * T * world * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
        "config": {"workload": "conv_glu LibriSpeech 17-layer Conv1D+GLU acoustic model (WeightNorm) + ASG, full train step, "
                               f"B={B} x T={T} frames x {F} filterbanks per GPU, {N} letter classes "
                               "(BASELINE.json configs[2] shape; recipes/conv_glu/librispeech/network.arch)",
                   "global_batch": B * world, "output_frames": t_out, "parallelism": f"dp{world}", "params": trainer.num_params(0)},
        "final_loss_sum": float(loss.sum().item()), "clocks": clocks, "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "gemm_tf32_kernel (time convolutions as im2col-view GEMMs)",
                     "achieved": flops / (gemm_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                     "frac": flops / (gemm_ms * 1e-3) / 1e12 / peak, "traffic": None, "peak_source": src,
                     "gemm_ms_per_step": gemm_ms, "gemm_launches_per_step": len(kern) / prof_steps,
                     "algorithmic_flops_per_step": flops, "gemm_share_of_step": gemm_ms / (ms / args.steps)},
    }
    print(json.dumps(line))


def run_tds(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from wav2letter_b200 import capi
    from wav2letter_b200.trainer import SEQ2SEQ_TDS_CTC_ARCH, Trainer, init_distributed, nccl_unique_id

    tm = Timed(world, local_rank)
    cfg = dict(TDS_CFG)
    B, T, F, N, L = cfg["B"], cfg["T"], cfg["F"], cfg["N"], cfg["L"]
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=tm.dev)
        if rank == 0:
            uid.copy_(torch.tensor(list(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        init_distributed(rank, world, bytes(uid.cpu().tolist()))
    trainer = Trainer(SEQ2SEQ_TDS_CTC_ARCH, F, N, "ctc", "none", lr=cfg["lr"], momentum=cfg["momentum"], maxgradnorm=cfg["maxgradnorm"])
    trainer.sync_parameters()  # fl::allReduceParameters at the start of train(), Train.cpp:1078-1079
    rng = np.random.default_rng(1234 + rank)
    nsets = cfg["n_input_sets"]
    host_f, host_y, dev_f, dev_y = [], [], [], []
    for _ in range(nsets):
        f, y = make_tds_inputs(rng, cfg)
        hf, hy = torch.from_numpy(f).pin_memory(), torch.from_numpy(y).pin_memory()
        host_f.append(hf)
        host_y.append(hy)
        dev_f.append(hf.to(tm.dev))
        dev_y.append(hy.to(tm.dev))
    loss = torch.empty(B, dtype=torch.float32, device=tm.dev)
    stage_f = torch.empty((B, 1, F, T), dtype=torch.float32, device=tm.dev)
    stage_y = torch.empty((B, L), dtype=torch.int32, device=tm.dev)
    host_loss = torch.empty(B, dtype=torch.float32).pin_memory()
    total_batch = float(B * world)

    def step(i):
        k = i % nsets
        trainer.step(dev_f[k], dev_y[k], True, total_batch, loss)

    def step_e2e(i):
        k = i % nsets
        stage_f.copy_(host_f[k], non_blocking=True)
        stage_y.copy_(host_y[k], non_blocking=True)
        trainer.step(stage_f, stage_y, True, total_batch, loss)
        host_loss.copy_(loss, non_blocking=True)

    prof_steps = min(args.steps, 3)
    gemm_flops, t_out = tds_gemm_flops(SEQ2SEQ_TDS_CTC_ARCH, B, T, N)
    prof = capi.ProfileList(1, 100 * prof_steps)
    ms, kern, launches, clocks = tm.run(step, args.steps, args.warmup, sample_clocks=True, profile=prof, profile_steps=prof_steps)
    ms_e2e, _, _, _ = tm.run(step_e2e, args.steps, args.warmup)
    final_loss = float(loss.sum().item())
    # after the timed regions: one traced step (an event after every launch) -> warm in-situ share of each kernel
    tr = capi.trace(lambda: step(0))
    tr_total = sum(v[1] for v in tr.values()) or 1.0
    breakdown = {k: {"launches": v[0], "ms": round(v[1], 4), "share": round(v[1] / tr_total, 4)}
                 for k, v in sorted(tr.items(), key=lambda kv: -kv[1][1])}
    asg = asg_point(tm, rank, 10, 3, profile=False) if world == 1 else None
    if rank != 0:
        return
    peaks, src = measured_peaks()
    frames = B * T * world
    gemm_ms_per_step = sum(kern) / prof_steps
    achieved = gemm_flops / (gemm_ms_per_step * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    cpu_fps, cpu_s, cpu_threads = cpu_tds(2, 2, 1) if world == 1 else (None, 0.0, 0)  # CPU baseline: rank 0 at N=1 only
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    line = {
        "metric": "frames_per_sec", "value": frames * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32",
        "dtype_note": "f32 storage; dense contractions (Linear, time convolution) multiply TF32 operands on the tensor cores with f32 accumulation (what north_star asks of the TDS blocks; >= the bf16 of BASELINE configs 2-3); criterion, LayerNorm, optimizer in f32/f64",
        "data": "synthetic",
        "config": {"workload": "seq2seq_tds LibriSpeech TDS acoustic model + CTC, full train step (fwd, CTC, bwd, all-reduce, clip, SGD), "
                               f"B={B} x T={T} frames x {F} filterbanks per GPU, {N} word-piece classes, targets <= {L} "
                               "(BASELINE.json configs[1]; recipes/seq2seq_tds/librispeech/{network.arch,train.cfg})",
                   "global_batch": B * world, "frames_per_utterance": T, "output_frames": t_out, "parallelism": f"dp{world}",
                   "optimizer": f"SGD lr={cfg['lr']} momentum={cfg['momentum']} maxgradnorm={cfg['maxgradnorm']}", "dropout": 0.2,
                   "params": trainer.num_params(0),
                   "cold_inputs": f"rotating {nsets} input sets; activations + gradients of a step ({B}x{T} frames) exceed the 126 MB L2"},
        "final_loss_sum": final_loss,
        "clocks": clocks,
        "e2e": {"value": frames * args.steps / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": B * F * T * 4 + B * L * 4, "d2h_bytes_per_step": B * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "gemm_tf32_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "peak_source": src + " bf16_tflops_sustained; tf32 math has half the bf16 hardware ceiling",
                     "gemm_ms_per_step": gemm_ms_per_step, "gemm_launches_per_step": len(kern) / prof_steps,
                     "algorithmic_flops_per_step": gemm_flops, "gemm_share_of_step": gemm_ms_per_step / (ms / args.steps)},
        "step_breakdown": {"method": "one extra traced step after the timed region, a CUDA event after every launch",
                           "traced_ms": round(tr_total, 3), "kernels": dict(list(breakdown.items())[:14])},
        "cpu_baseline": ({"value": cpu_fps, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                          "sample": f"2 train steps of B=2,T={T} on torch-CPU/oneDNN + C-oracle CTC ({cpu_s:.1f} s/step)"}
                         if cpu_fps is not None else None),
    }
    if asg is not None:
        alg = asg_algorithmic_bytes(ASG_CFG["B"], ASG_CFG["T"], ASG_CFG["N"], ASG_CFG["L"])
        line["asg_fwd_bwd_ms_per_batch"] = asg["ms"]
        line["asg_point"] = {"config": "T=1500 N=30 B=64", "ms_per_batch": asg["ms"], "e2e_ms_per_batch": asg["ms_e2e"],
                             "frames_per_sec": ASG_CFG["B"] * ASG_CFG["T"] / (asg["ms"] * 1e-3),
                             "algorithmic_GBps": alg / (asg["ms"] * 1e-3) / 1e9}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="tds_ctc", choices=["tds_ctc", "asg", "conv_glu_asg"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    import torch

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        {"tds_ctc": run_tds, "asg": run_asg, "conv_glu_asg": run_conv_glu}[args.workload](args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
