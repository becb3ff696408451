"""numpy restatement of the reference's in-tree streaming inference modules — TEST INFRASTRUCTURE ONLY.

Follows recipes/streaming_convnets/inference/inference/module/nn/{Conv1d,LayerNorm,Linear,Relu,Residual,TDSBlock}.cpp
(the fbgemm backend's arithmetic without its fp16 weight packing): activations per frame [groups][channels/groups],
Conv1d = the same [cout/g][kw][cin/g] filter applied to every group with zero padding left / right, Linear
y[o] = sum_i x[i] W[i*nOut + o] + b[o], LayerNorm per frame (x - mean) / stddev * alpha + beta with the population
standard deviation (stddev <= 1e-? -> 1), TDSBlock = Residual(conv, relu) -> LN -> Residual(linear, relu, linear) -> LN.
Pinned in tests/test_export_cpu.py by the reference's own Conv1d and TDSBlock known answers (tests/golden/).
Used to cross-check w2l_trainer_export_streaming end to end (the converter's own 1e-2 criterion,
tools/StreamingTDSModelConverter.cpp:346-376)."""
from __future__ import annotations

import json
import os

import numpy as np


def conv1d(x, w, b, groups, kw, stride, pad_left, pad_right):
    """x [T][groups*cin_g]; w [cout_g][kw][cin_g] (shared by the groups); -> [T'][groups*cout_g]"""
    T, feat = x.shape
    cin_g = feat // groups
    cout_g = w.shape[0]
    xg = np.pad(x.reshape(T, groups, cin_g).astype(np.float64), ((pad_left, pad_right), (0, 0), (0, 0)))
    Tout = (T + pad_left + pad_right - kw) // stride + 1
    y = np.zeros((Tout, groups, cout_g))
    for t in range(Tout):
        win = xg[t * stride:t * stride + kw]  # [kw][groups][cin_g]
        y[t] = np.einsum("kgc,okc->go", win, w.astype(np.float64)) + b
    return y.reshape(Tout, groups * cout_g)


def layernorm(x, alpha, beta):
    m = x.mean(axis=1, keepdims=True)
    s = np.sqrt(np.maximum((x * x).mean(axis=1, keepdims=True) - m * m, 0.0))
    s = np.where(s <= 1e-5, 1.0, s)
    return (x - m) / s * alpha + beta


def linear(x, w, b, nin, nout):
    return x @ w.reshape(nin, nout).astype(np.float64) + b


def tds_block(x, groups, conv_w, conv_b, kw, pad_left, pad_right, ln1, lin1_w, lin1_b, lin2_w, lin2_b, ln2, inner):
    feat = x.shape[1]
    h = layernorm(x + np.maximum(conv1d(x, conv_w, conv_b, groups, kw, 1, pad_left, pad_right), 0), *ln1)
    u = linear(np.maximum(linear(h, lin1_w, lin1_b, feat, inner), 0), lin2_w, lin2_b, inner, feat)
    return layernorm(h + u, *ln2)


def run_export(outdir: str, x: np.ndarray) -> np.ndarray:
    """x [T][n_feat] (one channel per filterbank) through the exported acoustic model -> [T'][n_label]"""
    spec = json.load(open(os.path.join(outdir, "acoustic_model.json")))
    blob = np.fromfile(os.path.join(outdir, "acoustic_model.bin"), dtype=np.float32)
    assert blob.size == spec["blob_floats"]
    x = x.astype(np.float64)
    for ly in spec["layers"]:
        t = ly["type"]
        if t == "conv1d":
            g = ly["groups"]
            cin_g, cout_g = ly["cin"] // g, ly["cout"] // g
            w = blob[ly["weight"]:ly["weight"] + cout_g * ly["kw"] * cin_g].reshape(cout_g, ly["kw"], cin_g)
            b = blob[ly["bias"]:ly["bias"] + cout_g]
            x = conv1d(x, w, b, g, ly["kw"], ly["stride"], ly["pad_left"], ly["pad_right"])
        elif t == "relu":
            x = np.maximum(x, 0)
        elif t == "layernorm":
            x = layernorm(x, ly["gain"], ly["bias"])
        elif t == "linear":
            w = blob[ly["weight"]:ly["weight"] + ly["nin"] * ly["nout"]]
            b = blob[ly["bias"]:ly["bias"] + ly["nout"]]
            x = linear(x, w, b, ly["nin"], ly["nout"])
        elif t == "tds":
            g, c, kw, feat, inner = ly["groups"], ly["channels"], ly["kw"], ly["feat"], ly["inner"]
            cw = blob[ly["conv_weight"]:ly["conv_weight"] + c * kw * c].reshape(c, kw, c)
            cb = blob[ly["conv_bias"]:ly["conv_bias"] + c]
            x = tds_block(x, g, cw, cb, kw, ly["pad_left"], ly["pad_right"], ly["ln1"],
                          blob[ly["lin1_weight"]:ly["lin1_weight"] + feat * inner], blob[ly["lin1_bias"]:ly["lin1_bias"] + inner],
                          blob[ly["lin2_weight"]:ly["lin2_weight"] + inner * feat], blob[ly["lin2_bias"]:ly["lin2_bias"] + feat], ly["ln2"], inner)
        else:
            raise ValueError(t)
    return x


def read_cereal_float_vector(path: str) -> np.ndarray:
    """cereal::BinaryInputArchive of std::vector<float>: u64 size tag + raw data (SimpleStreamingASRExample.cpp:206-217)"""
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw[:8], dtype="<u8")[0])
    assert len(raw) == 8 + 4 * n
    return np.frombuffer(raw[8:], dtype="<f4").copy()
