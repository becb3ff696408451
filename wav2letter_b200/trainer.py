"""Python handle on the C++ training-step driver (w2l_trainer_* in include/w2l_b200.h).

The harness owns device memory (torch tensors) and the launch (torchrun); every arithmetic step of the
training loop runs inside libw2l_b200.so.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import capi
from .capi import _check, _ptr, _stream, lib


class Trainer:
    def __init__(self, arch_text: str, n_feat: int, n_label: int, criterion: str = "ctc", scale_mode="none",
                 transdiag: float = 0.0, lr: float = 0.05, lrcrit: float = 0.0, momentum: float = 0.0,
                 maxgradnorm: float = 0.0, precision: str | None = None):
        """precision: None (the thread's w2l_set_precision), "tf32", "f32" (fp32-accurate) or "bf16"."""
        mode = capi.SCALE_MODES[scale_mode] if isinstance(scale_mode, str) else int(scale_mode)
        self.h = lib.w2l_trainer_create(_stream(), arch_text.encode(), n_feat, n_label, criterion.encode(), mode,
                                        transdiag, lr, lrcrit, momentum, maxgradnorm)
        if not self.h:
            raise capi.W2LError(1, lib.w2l_last_error().decode())
        self.h = ctypes.c_void_p(self.h)
        self.n_feat, self.n_label, self.criterion = n_feat, n_label, criterion
        if precision is not None:
            self.set_precision(precision)

    def set_precision(self, precision):
        _check(lib.w2l_trainer_set_precision(self.h, capi.PRECISIONS[precision] if isinstance(precision, str) else int(precision)))

    def skipped_steps(self) -> int:
        """steps whose update the device-side guard skipped (NaN / Inf loss or gradient); synchronises"""
        n = ctypes.c_longlong(0)
        _check(lib.w2l_trainer_status(self.h, _stream(), ctypes.byref(n)))
        return int(n.value)

    def save(self, path: str):
        """own-format checkpoint: constructor arguments + parameter / momentum arenas (Train.cpp:747-800)"""
        _check(lib.w2l_trainer_save(self.h, _stream(), path.encode()))

    @classmethod
    def load(cls, path: str) -> "Trainer":
        h = lib.w2l_trainer_load(_stream(), path.encode())
        if not h:
            raise capi.W2LError(1, lib.w2l_last_error().decode())
        self = cls.__new__(cls)
        self.h = ctypes.c_void_p(h)
        self.n_feat = self.n_label = self.criterion = None
        return self

    def export_streaming(self, outdir: str, tokens_text: str | None = None):
        """arrays of every layer in the layouts of the in-tree streaming inference library + transitions.bin / tokens.txt
        (the conversions of recipes/streaming_convnets/tools/StreamingTDSModelConverter.cpp)"""
        _check(lib.w2l_trainer_export_streaming(self.h, _stream(), outdir.encode(), None if tokens_text is None else tokens_text.encode()))

    def close(self):
        if getattr(self, "h", None):
            lib.w2l_trainer_destroy(self.h)
            self.h = None

    __del__ = close

    def describe(self) -> str:
        return lib.w2l_trainer_describe(self.h).decode()

    def num_params(self, which: int = 0) -> int:
        return int(lib.w2l_trainer_num_params(self.h, which))

    def layout(self, which: int = 0):
        """[(offset, elements, dims)] of every parameter inside the flat arena (16-byte aligned slots)."""
        n = 4096
        el = (ctypes.c_longlong * n)()
        dims = (ctypes.c_longlong * (4 * n))()
        cnt = lib.w2l_trainer_param_layout(self.h, which, n, el, dims)
        out, off = [], 0
        for i in range(cnt):
            out.append((off, int(el[i]), tuple(int(dims[4 * i + d]) for d in range(4))))
            off += (int(el[i]) + 3) // 4 * 4
        return out

    def get_flat(self, which: int = 0, what: int = 0) -> torch.Tensor:
        n = self.num_params(which)
        out = torch.empty(max(n, 1), dtype=torch.float32, device="cuda")
        _check(lib.w2l_trainer_get_flat(self.h, _stream(), which, what, _ptr(out)))
        return out[:n]

    def set_flat(self, flat: torch.Tensor, which: int = 0):
        _check(lib.w2l_trainer_set_flat(self.h, _stream(), which, _ptr(flat.contiguous())))

    def step(self, features: torch.Tensor, target: torch.Tensor, train: bool = True, total_batch: float | None = None,
             loss_out: torch.Tensor | None = None) -> torch.Tensor:
        """features: CUDA float [B,1,F,T] contiguous (== ArrayFire [T,F,1,B]); target CUDA int32 [B,L]."""
        B, _, F, T = features.shape
        L = target.shape[1]
        if loss_out is None:
            loss_out = torch.empty(B, dtype=torch.float32, device=features.device)
        _check(lib.w2l_trainer_step(self.h, _stream(), B, T, _ptr(features), L, _ptr(target), _ptr(loss_out), int(train),
                                    float(total_batch if total_batch is not None else B)))
        return loss_out

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        B, _, F, T = features.shape
        cap = B * (2 * T + 64) * self.n_label  # SAME-padded even kernels grow the frame count by one each
        out = torch.empty(cap, dtype=torch.float32, device=features.device)
        tout = ctypes.c_int(0)
        _check(lib.w2l_trainer_forward(self.h, _stream(), B, T, _ptr(features), _ptr(out), cap, ctypes.byref(tout)))
        return out[: B * tout.value * self.n_label].view(B, tout.value, self.n_label)

    def sync_parameters(self):
        _check(lib.w2l_trainer_sync_parameters(self.h, _stream()))


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    _check(lib.w2l_nccl_unique_id(buf))
    return buf.raw


def init_distributed(rank: int, world: int, uid: bytes):
    _check(lib.w2l_init_distributed(rank, world, ctypes.c_char_p(uid)))


from . import archs  # noqa: E402

# recipes/seq2seq_tds/librispeech/network.arch with the CTC head BASELINE.json configs[1] asks for
# (last line `L 1440 NLABEL` instead of the seq2seq encoder's `L 1440 1024`)
SEQ2SEQ_TDS_CTC_ARCH = archs.seq2seq_tds(ctc_head=True)


def conv_glu_librispeech_arch() -> str:
    """recipes/conv_glu/librispeech/network.arch (17 WeightNorm Conv1D+GLU layers, two WeightNorm Linear layers)"""
    return archs.conv_glu_librispeech()
