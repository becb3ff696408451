"""own-format checkpoints and the export for the in-tree streaming inference stack.

The export is cross-checked END TO END the way the reference's converter checks its own output
(tools/StreamingTDSModelConverter.cpp:346-376: same input through the training network and through the converted
inference modules, |difference| <= 1e-2): the exported arrays are run through oracle/inference_ref.py, the numpy
restatement of inference/module/nn/*.cpp that tests/test_export_cpu.py pins with the reference's own Conv1d / TDSBlock
known answers."""
import os

import numpy as np
import pytest
import torch

from oracle import inference_ref

pytestmark = pytest.mark.gpu

STREAMING = """V -1 NFEAT 1 0
SAUG 80 27 2 100 1.0 2
PD 0 5 3
C2 1 4 10 1 2 1 0 0
R
DO 0.0
LN 1 2
TDS 4 9 80 0.0 0 1 0
PD 0 7 1
C2 4 6 10 1 2 1 0 0
R
DO 0.0
LN 1 2
TDS 6 9 80 0.0 0 1 0
TDS 6 5 80 0.0 240 0 0
RO 2 1 0 3
V 480 -1 1 0
L 480 NLABEL
V NLABEL 0 -1 1
"""


def test_streaming_export_matches_training_network(tmp_path):
    from wav2letter_b200.trainer import Trainer

    N, B, T = 12, 2, 72
    tr = Trainer(STREAMING, 80, N, "asg", "target_sz_sqrt", transdiag=1.5, lr=0.02, lrcrit=0.01, precision="f32")
    g = torch.Generator(device="cuda").manual_seed(3)
    feat = torch.randn((B, 1, 80, T), device="cuda", generator=g)
    tgt = torch.randint(0, N, (B, 5), device="cuda", generator=g, dtype=torch.int32)
    for _ in range(3):  # move the parameters (and the transitions) off their initial values
        tr.step(feat, tgt, True)
    out = str(tmp_path / "export")
    tr.export_streaming(out, "\n".join(f"t{i}" for i in range(N)) + "\n")
    emis = tr.forward(feat).cpu().numpy()  # [B,T',N]  (eval mode: SpecAugment off)
    for b in range(B):
        x = feat[b, 0].t().cpu().numpy()  # [T][80]: one channel per filterbank, the inference library's input frame
        ref = inference_ref.run_export(out, x)
        assert ref.shape == emis[b].shape
        assert np.abs(ref - emis[b]).max() <= 1e-2, np.abs(ref - emis[b]).max()  # the converter's own criterion
        assert np.abs(ref - emis[b]).max() <= 2e-4 * max(1.0, np.abs(ref).max())  # and much tighter in fp32-accurate mode
    trans = inference_ref.read_cereal_float_vector(os.path.join(out, "transitions.bin"))
    np.testing.assert_array_equal(trans, tr.get_flat(1, 0).cpu().numpy())
    assert open(os.path.join(out, "tokens.txt")).read().split() == [f"t{i}" for i in range(N)]
    tr.close()


def test_export_rejects_non_streaming_archs(tmp_path):
    from wav2letter_b200 import W2LError
    from wav2letter_b200.trainer import Trainer

    tr = Trainer("V -1 NFEAT 1 0\nC2 1 4 5 1 2 1 -1 -1\nR\nLN 3\nV 0 320 1 0\nRO 1 0 3 2\nL 320 NLABEL\n", 80, 8, "ctc")
    with pytest.raises(W2LError):  # whole-sample LayerNorm cannot stream: the converter LOG(FATAL)s on it too
        tr.export_streaming(str(tmp_path / "x"))
    tr.close()


def test_checkpoint_round_trip_continues_training_identically(tmp_path):
    from wav2letter_b200.trainer import Trainer

    N, B, T = 10, 3, 64
    arch = STREAMING.replace("SAUG 80 27 2 100 1.0 2\n", "")
    tr = Trainer(arch, 80, N, "asg", "none", transdiag=1.0, lr=0.03, lrcrit=0.01, momentum=0.6, maxgradnorm=2.0, precision="tf32")
    g = torch.Generator(device="cuda").manual_seed(9)
    feat = torch.randn((B, 1, 80, T), device="cuda", generator=g)
    tgt = torch.randint(0, N, (B, 4), device="cuda", generator=g, dtype=torch.int32)
    for _ in range(3):
        tr.step(feat, tgt, True)
    path = str(tmp_path / "model.w2lb")
    tr.save(path)
    tr2 = Trainer.load(path)
    assert torch.equal(tr.get_flat(0, 0), tr2.get_flat(0, 0)) and torch.equal(tr.get_flat(1, 0), tr2.get_flat(1, 0))
    for _ in range(3):  # momentum state travelled too: the next steps agree (not bitwise: split-K weight gradients add atomically)
        l1 = tr.step(feat, tgt, True).clone()
        l2 = tr2.step(feat, tgt, True).clone()
        assert torch.allclose(l1, l2, rtol=2e-4, atol=1e-5), (l1, l2)
    # parameters after the three further steps: equal up to the run-to-run noise of TF32 training.  Not element-wise at
    # 1e-5: the scalar LayerNorm gains / biases receive gradients that are sums with heavy cancellation over every
    # activation, so the atomically-ordered split-K sums upstream move them by ~1e-4 of the parameter scale per step.
    p1, p2 = tr.get_flat(0, 0).double(), tr2.get_flat(0, 0).double()
    assert float((p1 - p2).norm() / p1.norm()) <= 1e-4
    assert float((p1 - p2).abs().max()) <= 2e-3 * float(p1.abs().max())
    tr.close()
    tr2.close()
    with pytest.raises(Exception):
        Trainer.load(str(tmp_path / "missing.w2lb"))
