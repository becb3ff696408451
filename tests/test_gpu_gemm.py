"""GPU parity of the tcgen05/TMA GEMM (TF32 math, fp32 accumulate) against a float64 reference of
the same op.  Tolerance: TF32 keeps 10 mantissa bits per operand (TMA rounds on load), so each
product carries <= 2^-10 relative error: |C - ref| <= 1.5e-3 * (|A| |B|^T) element-wise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(M, N, K, a_mn, b_mn, bias, act, seed=0):
    import wav2letter_b200 as w

    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g)
    bv = torch.randn(N, device="cuda", generator=g) if bias else None
    C = w.capi.gemm_tf32(A, B, bv, act, a_mn, b_mn)
    torch.cuda.synchronize()
    A64 = (A.t() if a_mn else A).double()
    B64 = (B.t() if b_mn else B).double()
    ref = A64 @ B64.t()
    bound = 1.5e-3 * (A64.abs() @ B64.abs().t()) + 1e-5
    if bias:
        ref = ref + bv.double()
    if act:
        ref = ref.clamp_min(0)
    err = (C.double() - ref).abs()
    ratio = float((err / bound).max())
    assert ratio <= 1.0, f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: err/bound {ratio}, max err {float(err.max())}"
    # and it must be a real TF32-accurate product, not garbage that happens to be small
    assert float(err.max()) < 0.05 * float(ref.abs().max()) + 1e-3


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 64), (200, 136, 100), (76, 52, 36),
                                   (4000, 800, 800), (1000, 2000, 1440)])
def test_gemm_majors_and_tails(M, N, K, a_mn, b_mn):
    run(M, N, K, a_mn, b_mn, bias=False, act=0, seed=M + N + K)


def test_gemm_bias_relu_epilogue():
    run(300, 1120, 1120, False, False, bias=True, act=1)
    run(130, 100, 64, False, False, bias=True, act=0)


def test_gemm_rejects_unaligned():
    import wav2letter_b200 as w

    A = torch.randn(8, 30, device="cuda")
    B = torch.randn(8, 30, device="cuda")
    with pytest.raises(w.W2LError):
        w.capi.gemm_tf32(A, B)


@pytest.mark.parametrize("accumulate", [False, True])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(1120, 1120, 4800, True, True), (360, 360, 9600, True, True),
                                              (256, 200, 2052, False, False), (128, 128, 16 * 32, False, True)])
def test_gemm_split_k(M, N, K, a_mn, b_mn, accumulate):
    """few output tiles + long K (the weight-gradient shape): the k blocks are split over blockIdx.z and summed with
    vector atomics; same TF32 bound, plus the accumulate-onto-C semantics"""
    import wav2letter_b200 as w

    g = torch.Generator(device="cuda").manual_seed(K)
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g)
    C0 = torch.randn(M, N, device="cuda", generator=g) * 10
    C = C0.clone()
    w.capi.gemm_tf32_ex(A, B, C, a_mn=a_mn, b_mn=b_mn, accumulate=accumulate)
    torch.cuda.synchronize()
    A64 = (A.t() if a_mn else A).double()
    B64 = (B.t() if b_mn else B).double()
    ref = A64 @ B64.t() + (C0.double() if accumulate else 0)
    bound = 1.5e-3 * (A64.abs() @ B64.abs().t()) + 1e-4
    ratio = float(((C.double() - ref).abs() / bound).max())
    assert ratio <= 1.0, f"split-K M={M} N={N} K={K}: err/bound {ratio}"


@pytest.mark.parametrize("bn", [128, 160, 224, 256])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(1000, 800, 800, False, False), (700, 1120, 1120, False, True),
                                              (300, 1440, 1440, False, False), (1120, 1120, 2400, True, True),
                                              (200, 456, 100, False, True), (130, 10000, 64, False, False)])
def test_gemm_tile_widths(M, N, K, a_mn, b_mn, bn):
    """every tile width the host heuristic can pick (w2l_gemm_set_tile pins it), tails in M, N and K included"""
    import wav2letter_b200 as w

    w.capi.gemm_set_tile(bn)
    try:
        run(M, N, K, a_mn, b_mn, bias=not a_mn, act=0 if a_mn else 1, seed=bn + M)
    finally:
        w.capi.gemm_set_tile(0)


@pytest.mark.parametrize("T,Cin,Cout,kw", [(300, 40, 200, 13), (257, 100, 136, 5), (64, 244, 96, 21)])
def test_gemm_overlapping_rows_is_a_time_convolution(T, Cin, Cout, kw):
    """w2l_gemm_tf32_view: the im2col matrix of a stride-1 conv over [T][Cin] is rows of kw*Cin floats with row stride Cin
    (a zero-copy TMA view).  fwd (K-major view as A) and wgrad (MN-major view as B) against conv1d in float64."""
    import wav2letter_b200 as w

    g = torch.Generator(device="cuda").manual_seed(T)
    x = torch.randn(T, Cin, device="cuda", generator=g)
    wt = torch.randn(Cout, kw, Cin, device="cuda", generator=g) * 0.1   # arranged [co][dk*Cin + ci]
    bias = torch.randn(Cout, device="cuda", generator=g)
    Tout = T - kw + 1
    y = torch.empty(Tout, Cout, device="cuda")
    w.capi.gemm_tf32_view(x, Cin, wt.view(Cout, kw * Cin), kw * Cin, y, Tout, Cout, kw * Cin, bias=bias)
    ref = torch.nn.functional.conv1d(x.double().t().unsqueeze(0), wt.double().permute(0, 2, 1), bias.double())[0].t()
    assert float((y.double() - ref).abs().max()) < 3e-3 * float(ref.abs().max()) + 1e-3
    dy = torch.randn(Tout, Cout, device="cuda", generator=g)
    dw = torch.zeros(Cout, kw * Cin, device="cuda")
    w.capi.gemm_tf32_view(dy, Cout, x, Cin, dw, Cout, kw * Cin, Tout, a_mn=True, b_mn=True)
    cols = torch.stack([x[dk:dk + Tout] for dk in range(kw)], 1).reshape(Tout, kw * Cin).double()
    refw = dy.double().t() @ cols
    assert float((dw.double() - refw).abs().max()) < 3e-3 * float(refw.abs().max()) + 1e-3


@pytest.mark.parametrize("T,Cin,Cout,kw", [(52, 16, 40, 4), (300, 44, 200, 13)])
def test_gemm_view_data_gradient(T, Cin, Cout, kw):
    """data gradient of a stride-1 valid convolution: the view of dY padded with kw-1 zero rows in front (row t = frames
    t-(kw-1) .. t) times the flipped weights"""
    import wav2letter_b200 as w

    g = torch.Generator(device="cuda").manual_seed(T + kw)
    Tout = T - kw + 1
    dyp = torch.zeros(T + kw - 1, Cout, device="cuda")
    dyp[kw - 1:kw - 1 + Tout] = torch.randn(Tout, Cout, device="cuda", generator=g)
    wt = torch.randn(Cout, Cin, kw, device="cuda", generator=g) * 0.1   # [co][ci][dk]
    flip = wt.flip(2).permute(1, 2, 0).contiguous()                      # [ci][j][co], j = kw-1-dk
    dx = torch.empty(T, Cin, device="cuda")
    w.capi.gemm_tf32_view(dyp, Cout, flip.view(Cin, kw * Cout), kw * Cout, dx, T, Cin, kw * Cout)
    x64 = torch.zeros(1, Cin, T, dtype=torch.float64, device="cuda", requires_grad=True)
    y = torch.nn.functional.conv1d(x64, wt.double())
    y.backward(dyp[kw - 1:kw - 1 + Tout].double().t().unsqueeze(0))
    ref = x64.grad[0].t()
    assert float((dx.double() - ref).abs().max()) < 3e-3 * float(ref.abs().max()) + 1e-3
