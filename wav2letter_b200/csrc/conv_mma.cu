// conv_mma.cu — tensor-core time convolution for the TDS family (sm_100a).
//
// A kw x 1 convolution with 10-27 channels over [B][T][C][W=80] activations is, per output frame, the tiny
// contraction  Y[co][w] = sum_k Wt[co][k] * X[k][w],  k = dk*Cin + ci  (K = kw*Cin = 210..378, M = Cout <= 32,
// N = W = 80).  tcgen05's minimum tile (M = 128, one CTA-wide accumulator) does not fit that shape; the
// warp-level tensor path does: mma.sync.m16n8k8 TF32 with fp32 accumulation, one warp per (frame, half of W),
// operands read from a shared-memory window of the input (each input row is staged ONCE per CTA and reused by
// every tap that touches it — the reuse the SIMT kernels in am_kernels.cu got only through L1).
//   forward / stride-1 data gradient : A = weights [Mpad][Kpad] (smem), B = input window rows (k -> row
//                                      base(t') + k because rows are laid out [tin][ci])
//   weight gradient                  : D[co][k] += dY[co][w] * X[k][w] accumulated over (b, t', w);
//                                      warps own disjoint k tiles, CTA partials reduced deterministically.
// Row pitches are chosen so every fragment load is bank-conflict free (forward: pitch = 8 mod 32 for the
// [k][w] operand; wgrad: pitch = 20 mod 32).  Operands are rounded to TF32 (cvt.rna) when staged.
// Reference op: fl::Conv2D via cuDNN (TF32 by default on Ampere+), arch opcodes C2 / TDS
// (recipes/joint_training_vox_populi/cpc/SequentialBuilder.cpp:254-301).
#include <cuda_runtime.h>

#include "common.cuh"

namespace w2l {
namespace {

constexpr int kMmaThreads = 256;  // 8 warps

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
}
__device__ __forceinline__ float drop_scale(unsigned long long seed, unsigned long long idx, float p, float inv_keep) {
  return dropout_scale(seed, idx, p, inv_keep);
}

// stage rows [row0, row0 + nrows) of the input window: smem row r = (tin - tin0) * Cin + ci, pitch `pitch`
__device__ __forceinline__ void stage_input(float* xs, int pitch, const float* __restrict__ xb, int T, int Cin, int W, int tin0,
                                            int nframes, int pad_rows) {
  const int nrows = nframes * Cin;
  const int chunks = W / 4;  // W % 4 == 0
  for (int i = threadIdx.x; i < nrows * chunks; i += blockDim.x) {
    const int r = i / chunks, c4 = i % chunks;
    const int tin = tin0 + r / Cin, ci = r % Cin;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tin >= 0 && tin < T) v = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)tin * Cin + ci) * W) + c4);
    float* dst = xs + (size_t)r * pitch + 4 * c4;
    dst[0] = to_tf32(v.x);
    dst[1] = to_tf32(v.y);
    dst[2] = to_tf32(v.z);
    dst[3] = to_tf32(v.w);
  }
  // rows read by the zero-padded tail of K must be finite
  for (int i = threadIdx.x; i < pad_rows * pitch; i += blockDim.x) xs[(size_t)nrows * pitch + i] = 0.f;
}

// ------------------------------------------------------------------------------------------
// forward: one CTA = TB = 8*UPW output frames of one sample; warp w owns frames w, w+8 (all of W).
// The K loop is chunked over ranges of kDeltaTaps taps: each chunk stages only the (TB-1)*stride + 8 input
// frames it touches (<= 100 KB for every TDS shape -> two CTAs per SM), accumulators stay in registers
// across chunks.  Weight fragments come straight from global memory (L1-resident, 50 KB at most).
// ------------------------------------------------------------------------------------------
constexpr int kDeltaTaps = 8;  // 8 * Cin is always a multiple of the MMA K (8)

template <int MT, int UPW>  // m tiles of 16 output channels (1 or 2); frames per warp
__global__ void __launch_bounds__(kMmaThreads) conv_mma_fwd_kernel(int T, int Tout, int W, int Cin, int Cout, int K, int stride,
                                                                    int pad_left, int Kpad, int apitch,
                                                                    const float* __restrict__ x, const float* __restrict__ wa,
                                                                    const float* __restrict__ bias, const float* __restrict__ add,
                                                                    float* __restrict__ y, int act, float drop_p,
                                                                    unsigned long long seed) {
  extern __shared__ __align__(16) float sm[];
  constexpr int kPitch = 88;  // 88 = 24 (mod 32): the 4 k-rows x 8 w of a B fragment hit 32 distinct banks
  constexpr int TB = 8 * UPW;
  float* xs = sm;  // [(nframes*Cin) + 8][kPitch]
  const int b = blockIdx.y, to0 = blockIdx.x * TB;
  const int nto = min(TB, Tout - to0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int wtiles = W / 8;  // <= 10
  float acc[UPW][MT][10][4];
#pragma unroll
  for (int u = 0; u < UPW; ++u)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int j = 0; j < 10; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][m][j][q] = 0.f;
  const float* xb = x + (size_t)b * T * Cin * W;
  for (int d0 = 0; d0 * Cin < Kpad; d0 += kDeltaTaps) {
    const int kbeg = d0 * Cin, kend = min(Kpad, (d0 + kDeltaTaps) * Cin);
    // frames this chunk touches: outputs tl = 0..nto-1, taps d0 .. d0+7 (clipped to the padded K range)
    const int taps = (kend - kbeg + Cin - 1) / Cin;
    const int nframes = (nto - 1) * stride + taps;
    __syncthreads();  // previous chunk fully consumed
    stage_input(xs, kPitch, xb, T, Cin, W, to0 * stride - pad_left + d0, nframes, 8);
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += 8) {
      float a[MT][4];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float* ap = wa + (size_t)(16 * m + g) * apitch + k0 + t4;
        a[m][0] = __ldg(ap);
        a[m][1] = __ldg(ap + 8 * apitch);
        a[m][2] = __ldg(ap + 4);
        a[m][3] = __ldg(ap + 8 * apitch + 4);
      }
#pragma unroll
      for (int u = 0; u < UPW; ++u) {
        const int tl = warp + 8 * u;
        if (tl < nto) {
          const float* bp = xs + (size_t)(tl * stride * Cin + (k0 - kbeg) + t4) * kPitch + g;
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            if (j < wtiles) {
              float bf[2];
              bf[0] = bp[8 * j];
              bf[1] = bp[4 * kPitch + 8 * j];
#pragma unroll
              for (int m = 0; m < MT; ++m) mma_tf32(acc[u][m][j], a[m], bf);
            }
          }
        }
      }
    }
  }
  // epilogue: d0 (co = g, w = 2 t4), d1 (g, 2 t4 + 1), d2 (g + 8, ..), d3
  const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
#pragma unroll
  for (int u = 0; u < UPW; ++u) {
    const int tl = warp + 8 * u;
    if (tl >= nto) continue;
    const int to = to0 + tl;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int co = 16 * m + g + 8 * hrow;
        if (co >= Cout) continue;
        const float bv = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          if (j < wtiles) {
            const int w = 8 * j + 2 * t4;
            const size_t idx = (((size_t)b * Tout + to) * Cout + co) * W + w;
            float v0 = acc[u][m][j][2 * hrow] + bv, v1 = acc[u][m][j][2 * hrow + 1] + bv;
            if (act == 1) {
              v0 = fmaxf(v0, 0.f);
              v1 = fmaxf(v1, 0.f);
            }
            if (drop_p > 0.f) {
              v0 *= drop_scale(seed, idx, drop_p, inv_keep);
              v1 *= drop_scale(seed, idx + 1, drop_p, inv_keep);
            }
            if (add != nullptr) {
              const float2 av = __ldg(reinterpret_cast<const float2*>(add + idx));
              v0 += av.x;
              v1 += av.y;
            }
            *reinterpret_cast<float2*>(y + idx) = make_float2(v0, v1);
          }
        }
      }
    }
  }
}

// weights wt[Cout][Cin][K] -> wa[16*MT][apitch] with k = dk*Cin + ci (TF32-rounded, zero padded);
// flip != 0 builds the stride-1 data-gradient operator: wa[ci][dk'*Cout + co] = wt[co][ci][K-1-dk']
__global__ void conv_mma_arrange_kernel(int Cin, int Cout, int K, int rows, int apitch, const float* __restrict__ wt,
                                        float* __restrict__ wa, int flip) {
  const int kin = flip ? Cout : Cin;  // channel count that plays "input" in the arranged operator
  const int mout = flip ? Cin : Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * apitch; i += gridDim.x * blockDim.x) {
    const int m = i / apitch, k = i % apitch;
    float v = 0.f;
    if (m < mout && k < K * kin) {
      const int dk = k / kin, c = k % kin;
      v = flip ? wt[((size_t)c * Cin + m) * K + (K - 1 - dk)] : wt[((size_t)m * Cin + c) * K + dk];
    }
    wa[i] = to_tf32(v);
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient: CTA = TC output frames of one sample; warps own disjoint 8-wide tiles of k
// ------------------------------------------------------------------------------------------
constexpr int kWgMaxNt = 8;  // k tiles per warp (Kc <= 8 * 8 * 8 = 512)

template <int MT>
__global__ void __launch_bounds__(kMmaThreads) conv_mma_wgrad_kernel(int T, int Tout, int W, int Cin, int Cout, int K, int stride,
                                                                      int pad_left, int TC, const float* __restrict__ x,
                                                                      const float* __restrict__ dy, float* __restrict__ partial) {
  extern __shared__ __align__(16) float sm[];
  constexpr int kPitch = 84;  // 84 = 20 (mod 32): 8 rows x 4 consecutive w of a fragment hit 32 distinct banks
  const int b = blockIdx.y;
  const int Kc = K * Cin, ktiles = (Kc + 7) / 8;
  float* xs = sm;                                                   // [(nframes*Cin) + 8][kPitch]
  float* ds = sm + ((size_t)((TC - 1) * stride + K) * Cin + 8) * kPitch;  // [TC][16*MT][kPitch]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  float acc[MT][kWgMaxNt][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < kWgMaxNt; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
  float bias_acc = 0.f;
  // a CTA walks chunks blockIdx.x, blockIdx.x + gridDim.x, ... of its sample, accumulating in registers
  for (int to0 = blockIdx.x * TC; to0 < Tout; to0 += gridDim.x * TC) {
  const int nto = min(TC, Tout - to0);
  const int nframes = (nto - 1) * stride + K;
  const int tin0 = to0 * stride - pad_left;
  __syncthreads();  // the previous chunk's reads are done
  stage_input(xs, kPitch, x + (size_t)b * T * Cin * W, T, Cin, W, tin0, nframes, 8);
  {
    const int chunks = W / 4;
    for (int i = threadIdx.x; i < nto * 16 * MT * chunks; i += blockDim.x) {
      const int c4 = i % chunks, co = (i / chunks) % (16 * MT), tl = i / (chunks * 16 * MT);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co < Cout) v = __ldg(reinterpret_cast<const float4*>(dy + (((size_t)b * Tout + to0 + tl) * Cout + co) * W) + c4);
      float* dst = ds + ((size_t)tl * 16 * MT + co) * kPitch + 4 * c4;
      dst[0] = to_tf32(v.x);
      dst[1] = to_tf32(v.y);
      dst[2] = to_tf32(v.z);
      dst[3] = to_tf32(v.w);
    }
  }
  __syncthreads();
  for (int tl = 0; tl < nto; ++tl) {
    const float* xrow = xs + (size_t)(tl * stride * Cin) * kPitch;
    const float* drow = ds + (size_t)tl * 16 * MT * kPitch;
    for (int w0 = 0; w0 < W; w0 += 8) {
      float a[MT][4];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float* ap = drow + (size_t)(16 * m + g) * kPitch + w0 + t4;  // A[co][w] = dy
        a[m][0] = ap[0];
        a[m][1] = ap[8 * kPitch];
        a[m][2] = ap[4];
        a[m][3] = ap[8 * kPitch + 4];
      }
#pragma unroll
      for (int j = 0; j < kWgMaxNt; ++j) {
        const int kt = warp + j * (kMmaThreads / 32);
        if (kt < ktiles) {
          const float* bp = xrow + (size_t)(8 * kt + g) * kPitch + w0 + t4;  // B[w][k] = x[row k][w]
          float bf[2];
          bf[0] = bp[0];
          bf[1] = bp[4];
#pragma unroll
          for (int m = 0; m < MT; ++m) mma_tf32(acc[m][j], a[m], bf);
        }
      }
    }
  }
  if (threadIdx.x < Cout) {  // bias gradient: sum over (t', w) of dy
    for (int tl = 0; tl < nto; ++tl) {
      const float* r = ds + ((size_t)tl * 16 * MT + threadIdx.x) * kPitch;
      for (int w = 0; w < W; ++w) bias_acc += r[w];
    }
  }
  }  // chunk loop
  // CTA partial: layout of the final gradient wt[co][ci][dk], then Cout bias sums
  float* out = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ((size_t)Cout * Cin * K + Cout);
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < kWgMaxNt; ++j) {
      const int kt = warp + j * (kMmaThreads / 32);
      if (kt < ktiles) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int co = 16 * m + g + 8 * (q >> 1), k = 8 * kt + 2 * t4 + (q & 1);
          if (co < Cout && k < Kc) {
            const int dk = k / Cin, ci = k % Cin;
            out[((size_t)co * Cin + ci) * K + dk] = acc[m][j][q];
          }
        }
      }
    }
  if (threadIdx.x < Cout) out[(size_t)Cout * Cin * K + threadIdx.x] = bias_acc;
}

__global__ void conv_mma_wgrad_reduce_kernel(int n_parts, int n_w, int n_b, const float* __restrict__ partial,
                                             float* __restrict__ dwt, float* __restrict__ dbias) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_w + n_b) return;
  float s0 = 0.f, s1 = 0.f;
  int q = 0;
  for (; q + 1 < n_parts; q += 2) {
    s0 += partial[(size_t)q * (n_w + n_b) + k];
    s1 += partial[(size_t)(q + 1) * (n_w + n_b) + k];
  }
  if (q < n_parts) s0 += partial[(size_t)q * (n_w + n_b) + k];
  if (k < n_w)
    dwt[k] += s0 + s1;
  else if (dbias != nullptr)
    dbias[k - n_w] += s0 + s1;
}

int apitch_for(int Kpad) {  // row pitch of the weight operand: = 4 (mod 32) -> conflict-free A fragments
  int p = Kpad + 4;
  while (p % 32 != 4) ++p;
  return p;
}

}  // namespace

// ---- host-side entry points used by am_kernels.cu's C ABI functions -------------------------------
bool conv_mma_supported(int W, int Cin, int Cout, int K, int stride) {
  return W % 8 == 0 && W <= 80 && Cin <= 32 && Cout <= 32 && K * std::max(Cin, Cout) <= 512 && stride <= K;
}
size_t conv_mma_arranged_floats(int Cin, int Cout, int K) {
  const int kin = std::max(Cin, Cout);
  const int Kpad = (K * kin + 7) / 8 * 8;
  return (size_t)32 * apitch_for(Kpad);
}

// flip = 0: y = conv(x) ; flip = 1: stride-1 data gradient (x := dy, roles of Cin/Cout swapped by the caller)
int conv_mma_fwd(cudaStream_t stream, int B, int T, int Tout, int W, int Cin, int Cout, int K, int stride, int pad_left,
                 const float* x, const float* wt, int wt_cin, int wt_cout, int flip, const float* bias, const float* add, float* y,
                 int act, float drop_p, unsigned long long seed, float* arranged) {
  const int MT = (Cout + 15) / 16;
  const int Kpad = (K * Cin + 7) / 8 * 8;
  const int apitch = apitch_for(Kpad);
  conv_mma_arrange_kernel<<<16, 256, 0, stream>>>(wt_cin, wt_cout, K, 16 * MT, apitch, wt, arranged, flip);
  W2L_LAUNCH_CHECK("conv_mma_arrange_kernel");
  // MT = 1: 16 frames per CTA (2 per warp); MT = 2: 8 frames per CTA — 80 accumulator registers either way
  const int TB = MT == 1 ? 16 : 8;
  const size_t smem = ((size_t)((TB - 1) * stride + kDeltaTaps) * Cin + 8) * 88 * 4;
  if (smem > 220 * 1024) return fail(W2L_ERR_UNSUPPORTED, "conv_mma_fwd: window does not fit in shared memory");
  dim3 grid((Tout + TB - 1) / TB, B);
  if (MT == 1) {
    if (smem > 48 * 1024) W2L_CUDA_CHECK(cudaFuncSetAttribute(conv_mma_fwd_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_mma_fwd_kernel<1, 2><<<grid, kMmaThreads, smem, stream>>>(T, Tout, W, Cin, Cout, K, stride, pad_left, Kpad, apitch, x, arranged,
                                                                   bias, add, y, act, drop_p, seed);
  } else {
    if (smem > 48 * 1024) W2L_CUDA_CHECK(cudaFuncSetAttribute(conv_mma_fwd_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_mma_fwd_kernel<2, 1><<<grid, kMmaThreads, smem, stream>>>(T, Tout, W, Cin, Cout, K, stride, pad_left, Kpad, apitch, x, arranged,
                                                                   bias, add, y, act, drop_p, seed);
  }
  W2L_LAUNCH_CHECK("conv_mma_fwd_kernel");
  return W2L_OK;
}

size_t conv_mma_wgrad_parts(int B, int Tout, int Cin, int Cout, int K, int stride, int* tc_out) {
  const int MT = (Cout + 15) / 16;
  int TC = 8;
  auto bytes_for = [&](int tc) { return (((size_t)((tc - 1) * stride + K) * Cin + 8) + (size_t)tc * 16 * MT) * 84 * 4; };
  while (TC > 1 && bytes_for(TC) > 200 * 1024) TC >>= 1;
  if (tc_out) *tc_out = TC;
  // CTAs per sample: enough to fill the chip ~twice, never more than there are chunks
  const int chunks = (Tout + TC - 1) / TC;
  const int per_sample = std::max(1, std::min(chunks, (2 * 148 + B - 1) / B));
  return (size_t)B * per_sample;
}

int conv_mma_wgrad(cudaStream_t stream, int B, int T, int Tout, int W, int Cin, int Cout, int K, int stride, int pad_left,
                   const float* x, const float* dy, float* dwt, float* dbias, float* partial) {
  const int MT = (Cout + 15) / 16;
  int TC = 8;
  const size_t parts = conv_mma_wgrad_parts(B, Tout, Cin, Cout, K, stride, &TC);
  const size_t smem = (((size_t)((TC - 1) * stride + K) * Cin + 8) + (size_t)TC * 16 * MT) * 84 * 4;
  if (smem > 220 * 1024) return fail(W2L_ERR_UNSUPPORTED, "conv_mma_wgrad: window does not fit in shared memory");
  dim3 grid((unsigned)(parts / B), B);
  if (MT == 1) {
    if (smem > 48 * 1024) W2L_CUDA_CHECK(cudaFuncSetAttribute(conv_mma_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_mma_wgrad_kernel<1><<<grid, kMmaThreads, smem, stream>>>(T, Tout, W, Cin, Cout, K, stride, pad_left, TC, x, dy, partial);
  } else {
    if (smem > 48 * 1024) W2L_CUDA_CHECK(cudaFuncSetAttribute(conv_mma_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_mma_wgrad_kernel<2><<<grid, kMmaThreads, smem, stream>>>(T, Tout, W, Cin, Cout, K, stride, pad_left, TC, x, dy, partial);
  }
  W2L_LAUNCH_CHECK("conv_mma_wgrad_kernel");
  const int n_w = Cout * Cin * K;
  conv_mma_wgrad_reduce_kernel<<<(n_w + Cout + 255) / 256, 256, 0, stream>>>((int)parts, n_w, Cout, partial, dwt, dbias);
  W2L_LAUNCH_CHECK("conv_mma_wgrad_reduce_kernel");
  return W2L_OK;
}

}  // namespace w2l
