"""Round-2 starting point: build experimental/gemm_tf32_2cta.cu on its own and compare it with torch (float64) and with
the production single-CTA kernel's timing.  NOT part of the test suite: the kernel has only been compile-checked."""
import ctypes, os, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = os.path.join(ROOT, "gpurun_out", "libexp2cta.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
                "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "wav2letter_b200", "csrc"),
                os.path.join(ROOT, "experimental", "gemm_tf32_2cta.cu"), os.path.join(ROOT, "wav2letter_b200", "csrc", "capi_common.cpp"),
                "-x", "cu", "-o", out, "-lcuda"], check=True)
lib = ctypes.CDLL(out)
vp, i = ctypes.c_void_p, ctypes.c_int
lib.w2l_exp_gemm_tf32_2cta.argtypes = [vp, i, i, i, i, i, i, vp, i, vp, i, vp, i, vp, i]
import wav2letter_b200 as w
for (M, N, K, a_mn, b_mn, bn) in [(512, 320, 256, 0, 0, 160), (9600, 800, 800, 0, 0, 160), (4800, 1120, 1120, 0, 0, 160),
                                  (2400, 1536, 1440, 0, 1, 256), (1024, 1024, 4800, 1, 1, 256)]:
    g = torch.Generator(device="cuda").manual_seed(M)
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g)
    C = torch.zeros(M, N, device="cuda")
    rc = lib.w2l_exp_gemm_tf32_2cta(torch.cuda.current_stream().cuda_stream, a_mn, b_mn, bn, M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(),
                                    B.stride(0), C.data_ptr(), N, None, 0)
    torch.cuda.synchronize()
    ref = (A.t() if a_mn else A).double() @ (B if b_mn else B.t()).double()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    t = []
    for fn in (lambda: lib.w2l_exp_gemm_tf32_2cta(torch.cuda.current_stream().cuda_stream, a_mn, b_mn, bn, M, N, K, A.data_ptr(), A.stride(0),
                                                   B.data_ptr(), B.stride(0), C.data_ptr(), N, None, 0),
               lambda: w.capi.gemm_tf32_ex(A, B, C, a_mn=bool(a_mn), b_mn=bool(b_mn))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1) / 20 * 1e3)
    print(f"M={M} N={N} K={K} majors=({a_mn},{b_mn}) bn={bn}: rc={rc} rel err {err:.2e}  2-CTA {t[0]:.1f} us  production {t[1]:.1f} us", flush=True)
