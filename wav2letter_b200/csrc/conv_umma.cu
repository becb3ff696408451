// conv_umma.cu — the TDS time convolution on the 5th-generation tensor cores (tcgen05 / TMEM), fed by TMA.
//
// A kw x 1 convolution over [B][T][C][W] activations (C = 10..27 channels, W = 80 filterbank columns innermost; arch
// opcodes C2 / TDS, cpc/SequentialBuilder.cpp:254-301) is, for one output frame t and column w,
//     Y[t][co][w] = sum_{dk,ci} Wt[co][ci][dk] * X[t*s + dk - pad][ci][w].
// GEMM view:  M = output POSITIONS (t, w)   — w is the contiguous axis, so the activation is an MN-MAJOR A operand
//             N = output channels (padded to 16 / 32)
//             K = (tap dk, input channel ci), channels padded to a multiple of 8 (one tf32 UMMA step)
// One CTA = F output frames x one 32-column chunk of W of one sample:
//   * the input window of those frames — (F-1)*s + kw frames x C rows x 32 columns — is brought in ONCE by TMA
//     (one 3-D box {32 w, C rows, 1 sample} per frame; frames outside the sample and columns >= W arrive as zeros),
//     each frame's rows padded to Cp = 8*ceil(C/8) rows in shared memory (the pad rows are zeroed once);
//   * a UMMA accumulator (M = 128) covers 4 consecutive output frames x 32 columns: the four 32-row M-atoms of its A
//     operand are the SAME shared-memory window at frame offsets s apart, so tap dk of the convolution is nothing but a
//     descriptor whose start address is dk frames further — no im2col, every input byte is staged exactly once;
//   * weights: K-major B operand [N][K] with k = dk*Cp + ci, arranged (and zero padded) once per call, loaded by TMA;
//   * the MMA thread issues kw * Cp/8 tcgen05.mma (128 x N x 8, kind::tf32) per accumulator; the accumulators of the
//     CTA's frame groups sit side by side in TMEM;
//   * epilogue: tcgen05.ld hands thread (frame, w) its N output channels; bias + ReLU + dropout (+ residual) and one
//     coalesced 128-byte store per (frame, channel).
// Roofline: HBM — the kernel reads x once (+ (kw-1)/F halo from L2) and writes y once.
// The data gradient of a stride-1 convolution is the same kernel on dy with flipped weights (channel roles swapped).
// Weight gradients and the two strided data gradients of a step stay on the mma.sync kernels (conv_mma.cu), as does
// W2L_PRECISION_F32 (3xTF32 split in registers).
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace w2l {
namespace {
using namespace umma;

constexpr int kCuThreads = 192;  // warps 0..3 epilogue, warp 4 TMA, warp 5 MMA + TMEM

struct ConvUmmaParams {
  int T, Tout, W, Cin, Cout, K, stride, pad_left;
  int Cp, Np, Kp, F, nf;  // padded channels / outputs / k extent; output frames per CTA; window frames
  const float* bias;
  const float* add;
  float* y;
  int act;
  float drop_p;
  unsigned long long seed;
};

__host__ __device__ constexpr uint32_t conv_idesc(int n) {
  // D = f32, A = B = tf32, A MN-major (bit 15), B K-major, N >> 3 at bit 17, M = 128 (>> 4) at bit 24
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int NP>  // 16 or 32 output channels (padded)
__global__ void __launch_bounds__(kCuThreads) conv_umma_fwd_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                   const __grid_constant__ CUtensorMap map_w, ConvUmmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int frame_bytes = p.Cp * 128;                       // one window frame: Cp rows of 32 columns
  unsigned char* win = smem;                                // [nf][Cp][32] fp32, 128B_ATOM_32B swizzle
  unsigned char* wsm = smem + (((size_t)p.nf * frame_bytes + 1023) & ~(size_t)1023);  // [Kp/32][NP][32] fp32, SWIZZLE_128B
  const int kblocks = p.Kp / 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + (size_t)kblocks * NP * 128);
  uint64_t* ld_full = bars;      // window + weights landed
  uint64_t* acc_full = bars + 1; // [4]: the MMAs of accumulator a are done (its epilogue overlaps the next one's MMAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* sbias = reinterpret_cast<float*>(bars + 6);  // [NP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, w0 = blockIdx.y * 32, to0 = blockIdx.x * p.F;
  const int n_acc = p.F / 4;
  const int tmem_cols = n_acc * NP <= 32 ? 32 : (n_acc * NP <= 64 ? 64 : (n_acc * NP <= 128 ? 128 : (n_acc * NP <= 256 ? 256 : 512)));

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    mbar_init(ld_full, 1);
    for (int a = 0; a < 4; ++a) mbar_init(&acc_full[a], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    // the allocation size is a run-time value of a few powers of two: one instruction per case (the operand is an immediate)
    const uint32_t slot = smem_u32(tmem_slot);
    if (tmem_cols == 32) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(slot) : "memory");
    else if (tmem_cols == 64) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(slot) : "memory");
    else if (tmem_cols == 128) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(slot) : "memory");
    else if (tmem_cols == 256) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(slot) : "memory");
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // pad rows (channels Cin..Cp-1 of every window frame) are never written by TMA: zero them (they meet zero weight
  // columns, but 0 * garbage could be NaN); bias row
  if (p.Cp > p.Cin) {
    const int pad_words = (p.Cp - p.Cin) * 32;
    for (int i = threadIdx.x; i < p.nf * pad_words; i += kCuThreads) {
      const int f = i / pad_words, r = i % pad_words;
      reinterpret_cast<float*>(win + (size_t)f * frame_bytes + (size_t)p.Cin * 128)[r] = 0.f;
    }
  }
  for (int j = threadIdx.x; j < NP; j += kCuThreads) sbias[j] = (p.bias != nullptr && j < p.Cout) ? __ldg(p.bias + j) : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the zeroed rows are read by the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // ===== TMA: weights (Kp/32 boxes of NP rows) + the window (one box of Cin rows per frame) onto one barrier =====
      mbar_expect_tx(ld_full, (uint32_t)(kblocks * NP * 128 + p.nf * p.Cin * 128));
      for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(&map_w, ld_full, wsm + (size_t)kb * NP * 128, kb * 32, 0);
      const int tin0 = to0 * p.stride - p.pad_left;
      for (int f = 0; f < p.nf; ++f) tma_load_3d(&map_x, ld_full, win + (size_t)f * frame_bytes, w0, (tin0 + f) * p.Cin, b);
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ===== MMA: accumulator a = output frames to0 + 4a .. + 3; tap dk = the same window dk frames further =====
      constexpr uint32_t idesc = conv_idesc(NP);
      mbar_wait(ld_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t win_sa = smem_u32(win), w_sa = smem_u32(wsm);
      const uint32_t lbo = (uint32_t)(p.stride * frame_bytes);  // M-atom stride: consecutive output frames
      const int ksteps = p.Cp / 8;
      for (int a = 0; a < n_acc; ++a) {
        const uint32_t d = tmem_base + (uint32_t)(a * NP);
        uint32_t accumulate = 0;
        for (int dk = 0; dk < p.K; ++dk) {
          const uint32_t fa = win_sa + (uint32_t)((4 * a * p.stride + dk) * frame_bytes);
          for (int kk = 0; kk < ksteps; ++kk) {
            const int kidx = dk * p.Cp + 8 * kk;
            const uint64_t da = make_smem_desc(fa + (uint32_t)(kk * 1024), lbo, 512, 1);  // MN-major, 128B_BASE32B: 8 k-rows = 1024 B
            const uint64_t db = make_smem_desc(w_sa + (uint32_t)((kidx >> 5) * NP * 128 + (kidx & 31) * 4), 16, 1024, 2);
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
                "}\n" ::"r"(d),
                "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                : "memory");
            accumulate = 1;
          }
        }
        umma_commit(&acc_full[a]);
      }
    }
  } else {
    // ===== epilogue: warp = frame within the accumulator's group of four, lane = column =====
    // Dropout: the backward pass reads the mask back from the stored activation (nothing is regenerated), so this kernel
    // is free to draw it its own way: ONE Philox block per thread and 8 output channels, 16 random bits per element
    // (keep iff bits >= p * 65536) — the block is keyed by (seed, element index of the first of the 8 channels).
    const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const uint32_t thresh = (uint32_t)(p.drop_p * 65536.0f);
    const int w = w0 + lane;
    const size_t cw = (size_t)p.W;
    for (int a = 0; a < n_acc; ++a) {
      mbar_wait(&acc_full[a], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int to = to0 + 4 * a + warp;
      const bool live = to < p.Tout && w < p.W;
      const size_t base = (((size_t)b * p.Tout + (size_t)(live ? to : 0)) * p.Cout) * cw + (live ? w : 0);
#pragma unroll
      for (int c0 = 0; c0 < NP; c0 += 16) {
        if (c0 >= p.Cout) break;  // warp-uniform
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * NP + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t rnd[8];  // 16 x 16 random bits
        if (p.drop_p > 0.f) {
          const unsigned long long i0 = (unsigned long long)(base + (size_t)c0 * cw), i1 = (unsigned long long)(base + (size_t)(c0 + 8) * cw);
          const uint4 r0 = philox4x32((uint32_t)i0, (uint32_t)(i0 >> 32), (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
          const uint4 r1 = philox4x32((uint32_t)i1, (uint32_t)(i1 >> 32), (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
          rnd[0] = r0.x; rnd[1] = r0.y; rnd[2] = r0.z; rnd[3] = r0.w;
          rnd[4] = r1.x; rnd[5] = r1.y; rnd[6] = r1.z; rnd[7] = r1.w;
        }
        float* yp = p.y + base + (size_t)c0 * cw;
        const float* ap = p.add != nullptr ? p.add + base + (size_t)c0 * cw : nullptr;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j < p.Cout) {  // warp-uniform
            float x = __uint_as_float(v[j]) + sbias[c0 + j];
            if (p.act == 1) x = fmaxf(x, 0.f);
            if (p.drop_p > 0.f) x *= ((rnd[j >> 1] >> (16 * (j & 1))) & 0xffffu) >= thresh ? inv_keep : 0.f;
            if (live) {
              if (ap != nullptr) x += ap[(size_t)j * cw];  // add may alias y (in-place accumulation): same element, same thread
              yp[(size_t)j * cw] = x;
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tmem_cols == 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_base) : "memory");
    else if (tmem_cols == 64) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
    else if (tmem_cols == 128) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
    else if (tmem_cols == 256) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// weights wt[Cout][Cin][K] -> wa[NP][Kp], k = dk*Cp + c (zero padded); flip: the data-gradient operator
// wa[ci][dk'*Cp + co] = wt[co][ci][K-1-dk'] (Cp = padded Cout then)
__global__ void conv_umma_arrange_kernel(int Cin, int Cout, int K, int Cp, int Np, int Kp, const float* __restrict__ wt,
                                         float* __restrict__ wa, int flip) {
  const int kin = flip ? Cout : Cin, mout = flip ? Cin : Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Np * Kp; i += gridDim.x * blockDim.x) {
    const int m = i / Kp, k = i % Kp;
    const int dk = k / Cp, c = k % Cp;
    float v = 0.f;
    if (m < mout && dk < K && c < kin) v = flip ? wt[((size_t)c * Cin + m) * K + (K - 1 - dk)] : wt[((size_t)m * Cin + c) * K + dk];
    wa[i] = v;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn conv_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

// ---- host-side entry points used by am_kernels.cu ----------------------------------------------------------
static int cu_pad8(int c) { return (c + 7) / 8 * 8; }
bool conv_umma_supported(int W, int Cin, int Cout, int K, int stride) {
  return W % 4 == 0 && W <= 96 && Cin <= 32 && Cout <= 32 && K >= 1 && K <= 64 && stride >= 1 && stride <= 4;
}
size_t conv_umma_arranged_floats(int Cin, int Cout, int K) {
  const int cp = cu_pad8(std::max(Cin, Cout));
  return (size_t)32 * (size_t)((K * cp + 31) / 32 * 32);
}

// y = dropout(act(conv(x) + bias)) (+ add); flip = 1: stride-1 data gradient (x := dy, the caller swaps the channel roles
// and passes pad_left' = K - 1 - pad_left), wt stays the forward weight tensor [wt_cout][wt_cin][K]
int conv_umma_fwd(cudaStream_t stream, int B, int T, int Tout, int W, int Cin, int Cout, int K, int stride, int pad_left, const float* x,
                  const float* wt, int wt_cin, int wt_cout, int flip, const float* bias, const float* add, float* y, int act, float drop_p,
                  unsigned long long seed, float* arranged) {
  EncodeTiledFn enc = conv_encode_fn();
  if (!enc) return fail(W2L_ERR_CUDA, "conv_umma: cuTensorMapEncodeTiled entry point not found");
  ConvUmmaParams p{};
  p.T = T;
  p.Tout = Tout;
  p.W = W;
  p.Cin = Cin;
  p.Cout = Cout;
  p.K = K;
  p.stride = stride;
  p.pad_left = pad_left;
  p.Cp = cu_pad8(Cin);
  p.Np = Cout <= 16 ? 16 : 32;
  p.Kp = (K * p.Cp + 31) / 32 * 32;
  p.bias = bias;
  p.add = add;
  p.y = y;
  p.act = act;
  p.drop_p = drop_p;
  p.seed = seed;
  // output frames per CTA: 16 (4 accumulators) when window + weights leave room for two CTAs per SM, else 8, else 4
  auto smem_for = [&](int F) {
    const size_t win = ((size_t)((F - 1) * stride + K) * p.Cp * 128 + 1023) / 1024 * 1024;
    return win + (size_t)(p.Kp / 32) * p.Np * 128 + 96 + p.Np * 4 + 1024;
  };
  int F = 16;
  if (smem_for(16) > 110 * 1024) F = 8;
  if (smem_for(F) > 220 * 1024) F = 4;
  if (smem_for(F) > 220 * 1024) return fail(W2L_ERR_UNSUPPORTED, "conv_umma: window does not fit in shared memory");
  if (Tout <= 8 && F > 8) F = 8;
  if (Tout <= 4) F = 4;
  p.F = F;
  p.nf = (F - 1) * stride + K;
  const size_t smem = smem_for(F);

  conv_umma_arrange_kernel<<<32, 256, 0, stream>>>(wt_cin, wt_cout, K, p.Cp, p.Np, p.Kp, wt, arranged, flip);
  W2L_LAUNCH_CHECK("conv_umma_arrange_kernel");

  CUtensorMap mx, mw;
  {  // activations as [B][T*Cin rows][W]: box {32 columns, Cin rows, 1 sample}; rows / columns outside arrive as zeros
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)T * (cuuint64_t)Cin, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)T * (cuuint64_t)Cin * (cuuint64_t)W * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)Cin, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(W2L_ERR_CUDA, "conv_umma: tensor map (activations) failed with code " + std::to_string((int)r));
  }
  {  // arranged weights [Np][Kp], K-major: box {32 k, Np rows}
    cuuint64_t dims[2] = {(cuuint64_t)p.Kp, (cuuint64_t)p.Np};
    cuuint64_t strides[1] = {(cuuint64_t)p.Kp * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)p.Np};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, arranged, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(W2L_ERR_CUDA, "conv_umma: tensor map (weights) failed with code " + std::to_string((int)r));
  }
  dim3 grid((Tout + F - 1) / F, (W + 31) / 32, B);
  if (p.Np == 16) {
    static bool cfg = false;
    if (!cfg) {
      W2L_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
      cfg = true;
    }
    conv_umma_fwd_kernel<16><<<grid, kCuThreads, smem, stream>>>(mx, mw, p);
  } else {
    static bool cfg = false;
    if (!cfg) {
      W2L_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
      cfg = true;
    }
    conv_umma_fwd_kernel<32><<<grid, kCuThreads, smem, stream>>>(mx, mw, p);
  }
  W2L_LAUNCH_CHECK("conv_umma_fwd_kernel");
  return W2L_OK;
}

}  // namespace w2l
