"""Pins for the CPU oracle (SURVEY.md §8c "replacement pins").

The reference tree has no ASG/CTC tests (parity unpinned), so the oracle is pinned by checks
that do not depend on any recollection of upstream code:
  1. brute-force enumeration of all N^T paths / all monotone alignments (tiny sizes),
  2. posterior marginals from that enumeration == oracle gradients,
  3. central finite differences,
  4. batching invariance incl. -1 padded targets,
  5. CTC: the two TensorFlow ctc_loss_op_test vectors + torch.nn.functional.ctc_loss,
  6. identities (trans == 0 -> FCC = sum_t LSE_i e_t ; N = L = 1 -> ASG = 0).
"""
import itertools
import math

import numpy as np
import pytest
import torch

import oracle


def lse(xs):
    xs = np.asarray(xs, dtype=np.float64)
    m = xs.max()
    return m + math.log(np.exp(xs - m).sum())


def brute_fcc(e, tr):
    """log-sum over all N^T paths; returns logZ, marginals [T,N], pair marginals [N,N]."""
    T, N = e.shape
    scores, paths = [], []
    for p in itertools.product(range(N), repeat=T):
        s = e[0, p[0]]
        for t in range(1, T):
            s += e[t, p[t]] + tr[p[t], p[t - 1]]
        scores.append(s)
        paths.append(p)
    logz = lse(scores)
    w = np.exp(np.asarray(scores) - logz)
    marg = np.zeros((T, N))
    pair = np.zeros((N, N))
    for wi, p in zip(w, paths):
        for t in range(T):
            marg[t, p[t]] += wi
        for t in range(1, T):
            pair[p[t], p[t - 1]] += wi
    best = paths[int(np.argmax(scores))]
    return logz, marg, pair, best, max(scores)


def alignments(T, L):
    """all monotone l_0=0 <= ... <= l_{T-1}=L-1 with steps in {0,1}."""
    if L > T:
        return
    for adv in itertools.combinations(range(1, T), L - 1):
        adv = set(adv)
        l, out = 0, []
        for t in range(T):
            if t in adv:
                l += 1
            out.append(l)
        yield out


def brute_fac(e, y, tr):
    T, N = e.shape
    L = len(y)
    scores, als = [], []
    for a in alignments(T, L):
        s = e[0, y[a[0]]]
        for t in range(1, T):
            s += e[t, y[a[t]]] + tr[y[a[t]], y[a[t - 1]]]
        scores.append(s)
        als.append(a)
    logz = lse(scores)
    w = np.exp(np.asarray(scores) - logz)
    marg = np.zeros((T, N))
    pair = np.zeros((N, N))
    for wi, a in zip(w, als):
        for t in range(T):
            marg[t, y[a[t]]] += wi
        for t in range(1, T):
            pair[y[a[t]], y[a[t - 1]]] += wi
    return logz, marg, pair, als, scores


@pytest.mark.parametrize("T,N,seed", [(1, 3, 0), (2, 2, 1), (4, 3, 2), (5, 4, 3), (6, 3, 4)])
def test_fcc_bruteforce(T, N, seed):
    rng = np.random.default_rng(seed)
    e = rng.normal(0, 2, (T, N)).astype(np.float32)
    tr = rng.normal(0, 1, (N, N)).astype(np.float32)
    logz, marg, pair, _, _ = brute_fcc(e.astype(np.float64), tr.astype(np.float64))
    loss, de, dtr = oracle.fcc(e[None], tr)
    assert abs(loss[0] - logz) < 1e-5 * max(1, abs(logz))
    np.testing.assert_allclose(de[0], marg, atol=2e-6)
    np.testing.assert_allclose(dtr, pair, atol=2e-6)


@pytest.mark.parametrize("T,N,L,seed", [(1, 2, 1, 0), (3, 3, 1, 1), (4, 3, 2, 2), (5, 4, 3, 3), (6, 3, 3, 4),
                                        (6, 4, 6, 5), (3, 2, 3, 6)])
def test_fac_bruteforce(T, N, L, seed):
    rng = np.random.default_rng(100 + seed)
    e = rng.normal(0, 2, (T, N)).astype(np.float32)
    tr = rng.normal(0, 1, (N, N)).astype(np.float32)
    y = rng.integers(0, N, L).astype(np.int32)
    if L >= 2:
        y[1] = y[0]  # force a repeated label: scatter must accumulate
    logz, marg, pair, _, _ = brute_fac(e.astype(np.float64), y, tr.astype(np.float64))
    loss, de, dtr = oracle.fac(e[None], y[None], tr)
    assert abs(loss[0] - logz) < 1e-5 * max(1, abs(logz))
    np.testing.assert_allclose(de[0], marg, atol=2e-6)
    np.testing.assert_allclose(dtr, pair, atol=2e-6)


def test_asg_is_fcc_minus_fac_and_nonneg():
    rng = np.random.default_rng(7)
    B, T, N, L = 3, 6, 4, 3
    e = rng.normal(0, 2, (B, T, N)).astype(np.float32)
    tr = rng.normal(0, 1, (N, N)).astype(np.float32)
    y = rng.integers(0, N, (B, L)).astype(np.int32)
    la, dea, dta = oracle.asg(e, y, tr)
    lf, def_, dtf = oracle.fcc(e, tr)
    lg, deg, dtg = oracle.fac(e, y, tr)
    np.testing.assert_allclose(la, lf - lg, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dea, def_ - deg, atol=1e-6)
    np.testing.assert_allclose(dta, dtf - dtg, atol=1e-5)
    assert (la > -1e-5).all()
    for b in range(B):
        fz = brute_fcc(e[b].astype(np.float64), tr.astype(np.float64))[0]
        gz = brute_fac(e[b].astype(np.float64), y[b], tr.astype(np.float64))[0]
        assert abs(la[b] - (fz - gz)) < 1e-4


@pytest.mark.parametrize("mode", list(oracle.SCALE_MODES))
def test_scale_modes_and_dloss(mode):
    rng = np.random.default_rng(11)
    B, T, N, L = 4, 9, 5, 4
    e = rng.normal(0, 2, (B, T, N)).astype(np.float32)
    tr = rng.normal(0, 1, (N, N)).astype(np.float32)
    y = rng.integers(0, N, (B, L)).astype(np.int32)
    y[1, 2:] = -1
    y[2, 3:] = -1
    tsz = oracle.target_sizes(y, T)
    assert tsz.tolist() == [4, 2, 3, 4]
    expect = {"none": [1] * B, "input_sz": [1 / T] * B, "input_sz_sqrt": [math.sqrt(1 / T)] * B,
              "target_sz": [1 / s for s in tsz], "target_sz_sqrt": [math.sqrt(1 / s) for s in tsz]}[mode]
    l0, de0, dt0 = oracle.asg(e, y, tr, "none")
    l1, de1, dt1 = oracle.asg(e, y, tr, mode)
    np.testing.assert_allclose(l1, l0 * np.asarray(expect, np.float32), rtol=1e-5)
    np.testing.assert_allclose(de1, de0 * np.asarray(expect, np.float32)[:, None, None], atol=1e-6)
    g = rng.normal(0, 1, B).astype(np.float32)
    l2, de2, dt2 = oracle.asg(e, y, tr, mode, dloss=g)
    np.testing.assert_allclose(de2, de1 * g[:, None, None], atol=1e-6)
    # d_trans is the batch sum of per-sample grads
    acc = np.zeros_like(dt2)
    for b in range(B):
        acc += oracle.asg(e[b:b + 1], y[b:b + 1], tr, mode, dloss=g[b:b + 1])[2]
    np.testing.assert_allclose(dt2, acc, atol=1e-5)


def test_batching_invariance_with_padding():
    rng = np.random.default_rng(13)
    B, T, N, L = 5, 12, 6, 7
    e = rng.normal(0, 3, (B, T, N)).astype(np.float32)
    tr = (4 * np.eye(N) + rng.normal(0, 0.1, (N, N))).astype(np.float32)
    y = rng.integers(0, N, (B, L)).astype(np.int32)
    lens = [7, 1, 3, 5, 2]
    for b, n in enumerate(lens):
        y[b, n:] = -1
    lb, deb, _ = oracle.asg(e, y, tr, "target_sz_sqrt")
    for b, n in enumerate(lens):
        l1, de1, _ = oracle.asg(e[b:b + 1], y[b:b + 1, :n], tr, "target_sz_sqrt")
        np.testing.assert_allclose(lb[b], l1[0], rtol=1e-6)
        np.testing.assert_allclose(deb[b], de1[0], atol=1e-7)


def test_finite_differences_asg():
    rng = np.random.default_rng(17)
    B, T, N, L = 2, 7, 4, 3
    e = rng.normal(0, 1, (B, T, N)).astype(np.float32)
    tr = rng.normal(0, 0.5, (N, N)).astype(np.float32)
    y = rng.integers(0, N, (B, L)).astype(np.int32)
    _, de, dtr = oracle.asg(e, y, tr)
    eps = 1e-2
    for (b, t, i) in [(0, 0, 0), (0, 3, 2), (1, 6, 3), (1, 2, 1)]:
        ep, em = e.copy(), e.copy()
        ep[b, t, i] += eps
        em[b, t, i] -= eps
        fd = (oracle.asg(ep, y, tr, backward=False).astype(np.float64).sum()
              - oracle.asg(em, y, tr, backward=False).astype(np.float64).sum()) / (2 * eps)
        assert abs(fd - de[b, t, i]) < 2e-3
    for (i, j) in [(0, 0), (1, 2), (3, 1)]:
        tp, tm = tr.copy(), tr.copy()
        tp[i, j] += eps
        tm[i, j] -= eps
        fd = (oracle.asg(e, y, tp, backward=False).astype(np.float64).sum()
              - oracle.asg(e, y, tm, backward=False).astype(np.float64).sum()) / (2 * eps)
        assert abs(fd - dtr[i, j]) < 2e-3


def test_identities():
    rng = np.random.default_rng(19)
    B, T, N = 3, 20, 7
    e = rng.normal(0, 3, (B, T, N)).astype(np.float32)
    l = oracle.fcc(e, np.zeros((N, N), np.float32), backward=False)
    ref = np.array([sum(lse(e[b, t]) for t in range(T)) for b in range(B)])
    np.testing.assert_allclose(l, ref, rtol=1e-6)
    e1 = rng.normal(0, 3, (2, 9, 1)).astype(np.float32)
    la = oracle.asg(e1, np.zeros((2, 1), np.int32), np.full((1, 1), 0.3, np.float32), backward=False)
    np.testing.assert_allclose(la, 0, atol=1e-5)


def test_target_clamped_to_T_and_invalid_targets():
    rng = np.random.default_rng(23)
    T, N = 4, 3
    e = rng.normal(0, 1, (2, T, N)).astype(np.float32)
    tr = rng.normal(0, 1, (N, N)).astype(np.float32)
    y = np.array([[0, 1, 2, 1, 0, 2], [2, 2, 1, 0, -1, -1]], np.int32)  # L=6 > T=4 -> first T labels
    l = oracle.asg(e, y, tr, backward=False)
    l2 = oracle.asg(e, y[:, :T], tr, backward=False)
    np.testing.assert_allclose(l, l2)
    bad = np.array([[-1, -1], [0, 7]], np.int32)
    lb, de, dtr = oracle.asg(e, bad, tr)
    assert np.isnan(lb).all() and not de.any() and not dtr.any()


@pytest.mark.parametrize("T,N,seed", [(1, 3, 0), (4, 3, 1), (6, 4, 2), (7, 3, 3)])
def test_fcc_viterbi_bruteforce(T, N, seed):
    rng = np.random.default_rng(200 + seed)
    e = rng.normal(0, 2, (T, N)).astype(np.float32)
    tr = rng.normal(0, 1, (N, N)).astype(np.float32)
    _, _, _, best, _ = brute_fcc(e.astype(np.float64), tr.astype(np.float64))
    p = oracle.fcc_viterbi(e[None], tr)[0]
    assert p.tolist() == list(best)


def test_fcc_viterbi_tie_break_first_max():
    T, N = 5, 4
    e = np.zeros((1, T, N), np.float32)
    tr = np.zeros((N, N), np.float32)
    assert oracle.fcc_viterbi(e, tr)[0].tolist() == [0] * T
    e[0, 2, 3] = 1.0
    assert oracle.fcc_viterbi(e, tr)[0].tolist() == [0, 0, 3, 0, 0]


@pytest.mark.parametrize("T,N,L,seed", [(1, 2, 1, 0), (5, 3, 2, 1), (6, 4, 3, 2), (6, 3, 6, 3), (7, 4, 4, 4)])
def test_fac_viterbi_bruteforce(T, N, L, seed):
    rng = np.random.default_rng(300 + seed)
    e = rng.normal(0, 2, (T, N)).astype(np.float32)
    tr = rng.normal(0, 1, (N, N)).astype(np.float32)
    y = rng.integers(0, N, L).astype(np.int32)
    _, _, _, als, scores = brute_fac(e.astype(np.float64), y, tr.astype(np.float64))
    best = als[int(np.argmax(scores))]
    path, idx = oracle.fac_viterbi(e[None], y[None], tr, return_index=True)
    assert idx[0].tolist() == best
    assert path[0].tolist() == [int(y[l]) for l in best]


# ---- CTC -----------------------------------------------------------------------------------
TF_P0 = [[0.633766, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
         [0.111121, 0.588392, 0.278779, 0.0055756, 0.00569609, 0.010436],
         [0.0357786, 0.633813, 0.321418, 0.00249248, 0.00272882, 0.0037688],
         [0.0663296, 0.643849, 0.280111, 0.00283995, 0.0035545, 0.00331533],
         [0.458235, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]]
TF_P1 = [[0.30176, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
         [0.24082, 0.397533, 0.0557226, 0.0546814, 0.0557528, 0.19549],
         [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, 0.202456],
         [0.280884, 0.429522, 0.0326593, 0.0339046, 0.0326856, 0.190345],
         [0.423286, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]]


def test_ctc_tensorflow_known_answers():
    """TensorFlow ctc_loss_op_test vectors."""
    l0 = oracle.ctc(np.log(np.asarray(TF_P0, np.float32))[None], np.array([[0, 1, 2, 1, 0]]), backward=False)
    l1 = oracle.ctc(np.log(np.asarray(TF_P1, np.float32))[None], np.array([[0, 1, 1, 0]]), backward=False)
    assert abs(l0[0] - 3.34211) < 1e-4
    assert abs(l1[0] - 5.42262) < 1e-4


def torch_ctc(e, y, lens):
    e = torch.tensor(e, dtype=torch.float64, requires_grad=True)
    B, T, N = e.shape
    lp = torch.log_softmax(e, -1).transpose(0, 1)
    yt = torch.tensor(np.where(y < 0, 0, y), dtype=torch.long)
    loss = torch.nn.functional.ctc_loss(lp, yt, torch.full((B,), T, dtype=torch.long),
                                        torch.tensor(lens, dtype=torch.long), blank=N - 1,
                                        reduction="none", zero_infinity=False)
    loss.sum().backward()
    return loss.detach().numpy(), e.grad.numpy()


@pytest.mark.parametrize("B,T,N,L,seed", [(1, 1, 3, 1, 0), (3, 12, 5, 4, 1), (4, 30, 8, 9, 2), (2, 25, 30, 12, 3)])
def test_ctc_vs_torch(B, T, N, L, seed):
    rng = np.random.default_rng(400 + seed)
    e = rng.normal(0, 2, (B, T, N)).astype(np.float32)
    y = rng.integers(0, N - 1, (B, L)).astype(np.int32)
    lens = [L] * B
    if B > 1:
        y[1, max(1, L // 2):] = -1
        lens[1] = max(1, L // 2)
        y[0, 1:3] = y[0, 0]  # repeats
    loss, de = oracle.ctc(e, y)
    tl, tg = torch_ctc(e, y, lens)
    np.testing.assert_allclose(loss, tl, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(de, tg, atol=2e-6)


def test_ctc_empty_target_and_feasibility_clamp():
    rng = np.random.default_rng(29)
    T, N = 6, 4
    e = rng.normal(0, 1, (1, T, N)).astype(np.float32)
    l = oracle.ctc(e, np.array([[-1, -1]], np.int32), backward=False)
    lp = e[0] - np.array([lse(r) for r in e[0]])[:, None]
    assert abs(l[0] + lp[:, N - 1].sum()) < 1e-5
    # L + repeats > T -> truncated to the feasible prefix (here [1,1,1,1,2] with T=6 keeps 3 labels + 2 repeats... )
    y = np.array([[1, 1, 1, 1, 2]], np.int32)
    n = oracle.ctc_target_sizes(y, T)[0]
    reps = sum(1 for k in range(1, n) if y[0, k] == y[0, k - 1])
    assert n + reps <= T and n >= 1
    l2 = oracle.ctc(e, y, backward=False)
    l3 = oracle.ctc(e, y[:, :n], backward=False)
    assert np.isfinite(l2).all()
    np.testing.assert_allclose(l2, l3)


def test_ctc_scale_modes():
    rng = np.random.default_rng(31)
    B, T, N, L = 3, 10, 6, 4
    e = rng.normal(0, 1, (B, T, N)).astype(np.float32)
    y = rng.integers(0, N - 1, (B, L)).astype(np.int32)
    y[2, 2:] = -1
    l0, d0 = oracle.ctc(e, y, "none")
    l1, d1 = oracle.ctc(e, y, "target_sz")
    s = np.array([1 / 4, 1 / 4, 1 / 2], np.float32)
    np.testing.assert_allclose(l1, l0 * s, rtol=1e-6)
    np.testing.assert_allclose(d1, d0 * s[:, None, None], atol=1e-7)


def test_argmax_path_and_linseg():
    e = np.zeros((1, 3, 4), np.float32)
    e[0, 0, 2] = 1
    e[0, 2, 1] = e[0, 2, 3] = 5
    assert oracle.argmax_path(e)[0].tolist() == [2, 0, 1]
    out = oracle.linseg_target(np.array([[3, 1, 2, -1]], np.int32), 7)
    assert out[0].tolist() == [3, 3, 3, 1, 1, 2, 2]
