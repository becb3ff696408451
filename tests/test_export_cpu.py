"""oracle/inference_ref.py (the numpy restatement of the in-tree streaming inference modules that checks
w2l_trainer_export_streaming) reproduces the reference's own known answers: Conv1dTest.cpp:32-70 and
TDSBlockTest.cpp:27-188, both at the reference's tolerance 1e-2."""
import os

import numpy as np

import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import conv1d_reference_golden as cg  # noqa: E402

from oracle import inference_ref as ir  # noqa: E402


def test_conv1d_reference_golden():
    x = np.array(cg.INPUT, np.float32).reshape(cg.T, cg.GROUPS * cg.CH_PER_GROUP)
    w = np.array(cg.WEIGHTS, np.float32).reshape(cg.CH_PER_GROUP, cg.KW, cg.CH_PER_GROUP)
    y = ir.conv1d(x, w, np.array(cg.BIAS), cg.GROUPS, cg.KW, 1, cg.PAD, cg.PAD)
    assert np.abs(y.reshape(-1) - np.array(cg.TARGET)).max() < 1e-2


def test_tds_block_reference_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "tds_block_golden.npz"))
    T, W, C, K = int(g["T"]), int(g["W"]), int(g["C"]), int(g["K"])
    x = g["in"].reshape(T, W * C)
    y = ir.tds_block(x, W, g["conv_weights"].reshape(C, K, C), g["conv_bias"], K, 1, 1, (float(g["ln1_weights"][0]), float(g["ln1_bias"][0])),
                     g["lin1_weights"], g["lin1_bias"], g["lin2_weights"], g["lin2_bias"], (float(g["ln2_weights"][0]), float(g["ln2_bias"][0])), W * C)
    assert np.abs(y - g["expectedOutput"].reshape(T, W * C)).max() < 1e-2
