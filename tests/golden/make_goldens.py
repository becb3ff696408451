"""Generates tests/golden/criterion_goldens.npz from the pinned oracle (run once, committed
together with its output).  The reference tree holds no ASG/CTC vectors, so the
fixtures are outputs of oracle/w2l_oracle.c, which tests/test_oracle_pins.py pins by brute force,
finite differences, torch.nn.functional.ctc_loss and the TensorFlow ctc_loss_op_test vectors."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import oracle  # noqa: E402


def main():
    rng = np.random.default_rng(20260922)
    out = {}
    # ASG: WSJ-like token set N=30, ragged targets, transdiag 4 (conv_glu/librispeech/train.cfg:25)
    B, T, N, L = 4, 60, 30, 12
    e = (rng.normal(0, 1, (B, T, N)) * 3).astype(np.float32)
    tr = (4 * np.eye(N) + rng.normal(0, 0.1, (N, N))).astype(np.float32)
    y = rng.integers(0, N, (B, L)).astype(np.int32)
    y[1, 7:] = -1
    y[2, 1:] = -1
    y[3, 3:5] = y[3, 2]
    loss, de, dtr = oracle.asg(e, y, tr, "target_sz_sqrt")
    out.update(asg_emis=e, asg_trans=tr, asg_target=y, asg_loss=loss, asg_d_emis=de, asg_d_trans=dtr)
    out["asg_viterbi"] = oracle.fcc_viterbi(e, tr)
    p, idx = oracle.fac_viterbi(e, y, tr, return_index=True)
    out.update(fac_viterbi=p, fac_viterbi_idx=idx)
    # CTC: N=31 (30 tokens + blank last), repeats and padding
    N2 = 31
    e2 = (rng.normal(0, 1, (B, T, N2)) * 2).astype(np.float32)
    y2 = rng.integers(0, N2 - 1, (B, L)).astype(np.int32)
    y2[0, 4:7] = y2[0, 3]
    y2[1, 5:] = -1
    y2[2, :] = -1
    l2, d2 = oracle.ctc(e2, y2, "target_sz")
    out.update(ctc_emis=e2, ctc_target=y2, ctc_loss=l2, ctc_d_emis=d2)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "criterion_goldens.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
