"""Python handle on the token / target pipeline (w2l_text_* in include/w2l_b200.h; C++ in host/text_pipeline.cpp):
the dataset's target transform and evalOutput's path -> letters -> words -> edit distance (Train.cpp:236-254,296-316,829-872)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import capi
from .capi import lib


class TextPipeline:
    def __init__(self, tokens_text: str, lexicon_text: str = "", criterion: str = "ctc", replabel: int = 0, surround: str = "",
                 usewordpiece: bool = False, wordsep: str = "|"):
        h = lib.w2l_text_create(tokens_text.encode(), lexicon_text.encode(), criterion.encode(), int(replabel), surround.encode(),
                                int(usewordpiece), wordsep.encode())
        if not h:
            raise capi.W2LError(1, lib.w2l_last_error().decode())
        self.h = ctypes.c_void_p(h)

    def close(self):
        if getattr(self, "h", None):
            lib.w2l_text_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def num_classes(self) -> int:
        return int(lib.w2l_text_num_classes(self.h))

    def _need(self, n):
        if n < 0:
            raise capi.W2LError(1, lib.w2l_last_error().decode())
        return int(n)

    def encode(self, transcript: str) -> np.ndarray:
        n = self._need(lib.w2l_text_encode(self.h, transcript.encode(), None, 0))
        out = np.zeros(max(n, 1), np.int32)
        lib.w2l_text_encode(self.h, transcript.encode(), out.ctypes.data_as(ctypes.c_void_p), n)
        return out[:n]

    def encode_batch(self, transcripts) -> np.ndarray:
        """[B][L] int32 padded with -1 (kTargetPadValue)"""
        rows = [self.encode(t) for t in transcripts]
        L = max(1, max(len(r) for r in rows))
        out = np.full((len(rows), L), -1, np.int32)
        for b, r in enumerate(rows):
            out[b, :len(r)] = r
        return out

    def _str(self, fn, arr) -> list:
        arr = np.ascontiguousarray(arr, dtype=np.int32)
        p = arr.ctypes.data_as(ctypes.c_void_p)
        n = self._need(fn(self.h, p, arr.size, None, 0))
        buf = ctypes.create_string_buffer(n)
        fn(self.h, p, arr.size, buf, n)
        return buf.value.decode().split()

    def prediction2ltr(self, path) -> list:
        return self._str(lib.w2l_text_prediction2ltr, path)

    def target2ltr(self, target_row) -> list:
        return self._str(lib.w2l_text_target2ltr, target_row)

    def ltr2wrd(self, letters) -> list:
        s = " ".join(letters).encode()
        n = self._need(lib.w2l_text_ltr2wrd(self.h, s, None, 0))
        buf = ctypes.create_string_buffer(n)
        lib.w2l_text_ltr2wrd(self.h, s, buf, n)
        return buf.value.decode().split()


class EditDistanceMeter:
    """fl::EditDistanceMeter: add(hypothesis tokens, reference tokens); value() = [error %, n, ins %, del %, sub %]"""

    def __init__(self):
        self.acc = (ctypes.c_longlong * 4)(0, 0, 0, 0)  # n, ndel, nins, nsub

    def add(self, hyp, ref):
        capi._check(lib.w2l_edit_distance(" ".join(hyp).encode(), " ".join(ref).encode(), self.acc))

    def value(self):
        n, ndel, nins, nsub = (int(v) for v in self.acc)
        d = float(max(n, 1))
        return [100.0 * (ndel + nins + nsub) / d, n, 100.0 * nins / d, 100.0 * ndel / d, 100.0 * nsub / d]

    def raw(self):
        return tuple(int(v) for v in self.acc)
