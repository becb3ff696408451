// criterion_ctc.cu — CTC forward + backward on raw activations for sm_100a.
// Replaces ConnectionistTemporalClassificationCriterion::forward (CTCLoss constructed at
// recipes/slimIPL/src/Train.cpp:406-407, cpc/Train.cpp:524-525; upstream CUDA backend = warp-ctc).
// Conventions: internal log-softmax over N, blank = N-1 (Train.cpp:248-251), every sample runs
// the full padded T (Train.cpp:1473-1477), target made feasible by L <- min(L,T), then
// L <- min(L+R,T)-R with R adjacent repeats.
//
// Pipeline:
//   1. ctc_prep_kernel   HBM-bound: lz[b][t] = logsumexp_k e[b][t][k] (one pass over the
//                        activations); per-sample target size / scale.
//   2. ctc_chains_kernel one CTA per sample, thread per extended-target state, log domain with
//                        per-step re-centring (fp32 stays accurate for long T): alpha walk
//                        storing the lattice, then beta walk emitting per-state posteriors.
//   3. ctc_grad_kernel   HBM-bound: d_emis = coef * (softmax - occupancy), one pass: reads the
//                        activations once more, writes the gradient once; the <= 2L+1 occupied
//                        labels of a frame are subtracted afterwards.
#include <cuda_runtime.h>

#include "common.cuh"

namespace w2l {
namespace {

constexpr int kCtcThreads = 256;
constexpr int kCtcWarps = kCtcThreads / 32;
// Every per-step global read of the chains (the gathered activations e_t[z_s], the stored alpha row in the beta walk) is
// requested kCtcDepth steps ahead with cp.async into shared-memory rings; the per-frame scalars live in shared memory.
// With one-step-ahead register prefetch the step time was one DRAM latency (~0.9 us): 0.28 ms for T' = 150.
// The depth is a template parameter (8 / 4 / 2 / 1, ring = 2 x depth slots): long targets shrink the rings to fit.

__device__ __forceinline__ void ctc_cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ctc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void ctc_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

struct CtcParams {
  int B, T, N, L, Sp, scale_mode, need_grad;
  int t_smem;     // per-frame scalars (lz, alpha offsets) of a sample fit in shared memory
  const float* emis;
  const int32_t* target;
  const float* dloss;
  float* loss;
  float* d_emis;
  float* lz;      // [B][T]
  float* lat;     // [B][T][Sp] alpha-tilde, overwritten by posteriors
  double* cA;     // [B][T]
  float* psum;    // [B][T]
  int* tsz;       // [B] feasible target size
  int* valid;     // [B]
  float* coef;    // [B]
  float* scale;   // [B]
};

__global__ void __launch_bounds__(256) ctc_prep_kernel(CtcParams p, int frame_blocks) {
  const int lane = threadIdx.x & 31;
  if ((int)blockIdx.x < frame_blocks) {
    const long long nframes = (long long)p.B * p.T;
    const long long warps = (long long)frame_blocks * (blockDim.x >> 5);
    for (long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < nframes; f += warps) {
      const float* e = p.emis + f * p.N;
      float m = kNegInf, s = 0.f;
      for (int k = lane; k < p.N; k += 32) {
        const float v = __ldg(e + k);
        if (v > m) {
          s = s * __expf(m - v) + 1.0f;
          m = v;
        } else {
          s += __expf(v - m);
        }
      }
      const float gm = warp_max(m);
      s = (m == kNegInf) ? 0.f : s * __expf(m - gm);
      s = warp_sum(s);
      if (lane == 0) p.lz[f] = gm + __logf(s);
    }
    return;
  }
  int b = ((int)blockIdx.x - frame_blocks) * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  int raw = 0, n = 0, ok = 1;
  if (p.target != nullptr && p.L > 0) {
    const int32_t* y = p.target + (size_t)b * p.L;
    raw = target_size(y, p.L, p.T);
    int r = 0;
    for (int l = 1; l < raw; ++l)
      if (y[l] == y[l - 1]) ++r;
    n = min(raw + r, p.T) - r;
    if (n < 0) n = 0;
    for (int l = 0; l < n; ++l)
      if (y[l] < 0 || y[l] >= p.N - 1) ok = 0;
  }
  const float sc = scale_of(p.scale_mode, p.T, raw);
  p.tsz[b] = n;
  p.valid[b] = ok;
  p.scale[b] = sc;
  p.coef[b] = ok ? sc * (p.dloss ? p.dloss[b] : 1.0f) : 0.f;
}

__device__ __forceinline__ float lse3f(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  if (m == kNegInf) return kNegInf;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

// block-wide max of per-thread values through per-warp slots (caller supplies the barrier)
template <bool kGrad, int kCtcDepth>
__global__ void __launch_bounds__(kCtcThreads) ctc_chains_kernel(CtcParams p) {
  constexpr int kCtcRing = 2 * kCtcDepth;  // slots, power of two > depth
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the block is sized to the state count (32 .. 256 threads): idle warps would still spend issue slots on every step
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nw = blockDim.x >> 5;
  const int T = p.T, N = p.N, Sp = p.Sp, blank = N - 1;
  float* row0 = reinterpret_cast<float*>(smem_raw) + 4;  // index -2..Sp+1 valid
  float* row1 = row0 + Sp + 8;
  int32_t* z = reinterpret_cast<int32_t*>(row1 + Sp + 4);
  float* ering = reinterpret_cast<float*>(z + Sp);              // [kCtcRing][Sp] gathered activations
  float* lring = ering + (size_t)kCtcRing * Sp;                  // [kCtcRing][Sp] stored alpha rows (beta walk)
  double* cA_s = reinterpret_cast<double*>(lring + (size_t)kCtcRing * Sp);  // [T] when p.t_smem
  float* lz_s = reinterpret_cast<float*>(cA_s + (p.t_smem ? p.T : 0));     // [T] when p.t_smem
  uint8_t* skip = reinterpret_cast<uint8_t*>(lz_s + (p.t_smem ? p.T : 0));  // skip[s]: s-2 -> s allowed
  __shared__ float wmax[2][kCtcWarps];
  __shared__ float wsum[2][kCtcWarps];
  __shared__ double ll_s;
  if (!p.valid[b]) {
    if (tid == 0) p.loss[b] = NAN;
    return;
  }
  const int Lb = p.tsz[b];
  const int S = 2 * Lb + 1;
  const int32_t* yg = p.target ? p.target + (size_t)b * p.L : nullptr;
  const float* eb = p.emis + (size_t)b * T * N;
  const float* lzb = p.lz + (size_t)b * T;
  float* lat = p.lat + (size_t)b * T * Sp;
  for (int s = tid; s < Sp; s += nthr) {
    int zs = blank;
    if (s < S && (s & 1)) zs = yg[s >> 1];
    z[s] = zs;
  }
  for (int s = tid; s < Sp + 8; s += nthr) {
    row0[s - 4] = kNegInf;
    row1[s - 4] = kNegInf;
  }
  const bool ts = p.t_smem != 0;
  if (ts)
    for (int t = tid; t < T; t += nthr) lz_s[t] = lzb[t];
  __syncthreads();
  for (int s = tid; s < Sp; s += nthr) skip[s] = (s >= 2 && s < S && z[s] != blank && z[s] != z[s - 2]) ? 1 : 0;
  // t = 0
  if (tid < 2 && tid < S) row0[tid] = eb[z[tid]] - lzb[0];
  __syncthreads();
  {
    float lm = kNegInf;
    for (int s = tid; s < S; s += nthr) {
      lm = fmaxf(lm, row0[s]);
      if (kGrad) lat[s] = row0[s];
    }
    lm = warp_max(lm);
    if (lane == 0) wmax[0][warp] = lm;
    if (kGrad && tid == 0) {
      p.cA[(size_t)b * T] = 0.0;
      if (p.t_smem) cA_s[0] = 0.0;
    }
  }
  __syncthreads();
  float* rp = row0;
  float* rn = row1;
  double C = 0.0;
  int par = 0;
  // one commit group per frame: the activations e_f[z_s] of this thread's states
  auto issue_a = [&](int f) {
    if (f < T) {
      float* dst = ering + (size_t)(f & (kCtcRing - 1)) * Sp;
      const float* src = eb + (size_t)f * N;
      for (int s = tid; s < S; s += nthr) ctc_cp_async4(dst + s, src + z[s]);
    }
    ctc_cp_commit();
  };
  for (int q = 1; q <= kCtcDepth; ++q) issue_a(q);
  for (int t = 1; t < T; ++t) {
    ctc_cp_wait<kCtcDepth - 1>();  // frame t has landed (own copies: no barrier needed)
    issue_a(t + kCtcDepth);
    float d = wmax[par][0];
#pragma unroll
    for (int w = 1; w < nw; ++w) d = fmaxf(d, wmax[par][w]);
    if (!(d > -1e30f)) d = 0.f;
    C += (double)d;
    const float lzt = (ts ? lz_s[t] : lzb[t]) + d;
    float lm = kNegInf;
    const float* er = ering + (size_t)(t & (kCtcRing - 1)) * Sp;
    for (int s = tid; s < S; s += nthr) {
      const float a2 = skip[s] ? rp[s - 2] : kNegInf;
      const float v = lse3f(rp[s], rp[s - 1], a2);
      const float val = (v == kNegInf) ? kNegInf : v + (er[s] - lzt);
      rn[s] = val;
      lm = fmaxf(lm, val);
      if (kGrad) lat[(size_t)t * Sp + s] = val;
    }
    lm = warp_max(lm);
    if (lane == 0) wmax[par ^ 1][warp] = lm;
    if (kGrad && tid == 0) {
      p.cA[(size_t)b * T + t] = C;
      if (ts) cA_s[t] = C;
    }
    __syncthreads();
    float* tmp = rp;
    rp = rn;
    rn = tmp;
    par ^= 1;
  }
  ctc_cp_wait<0>();
  if (tid == 0) {
    const float a = rp[S - 1], a2 = S > 1 ? rp[S - 2] : kNegInf;
    const double ll = (double)lse2f(a, a2) + C;
    ll_s = ll;
    p.loss[b] = (float)(-(double)p.scale[b] * ll);
  }
  if (!kGrad) return;
  __syncthreads();
  const double ll = ll_s;
  if (!(ll > -1e30)) {  // infeasible (loss = +inf, e.g. -inf activations on every path): zero gradient —
    // ctc_grad_kernel zeroes the rows of samples whose `valid` flag is clear (it runs after this kernel)
    if (tid == 0) p.valid[b] = 0;
    return;
  }
  // ---- beta walk; posteriors overwrite the alpha lattice ------------------------------------------
  for (int s = tid; s < Sp + 8; s += nthr) {
    row0[s - 4] = kNegInf;
    row1[s - 4] = kNegInf;
  }
  __syncthreads();
  rp = row0;
  rn = row1;
  if (tid == 0) {
    row0[S - 1] = eb[(size_t)(T - 1) * N + z[S - 1]] - lzb[T - 1];
    if (S > 1) row0[S - 2] = eb[(size_t)(T - 1) * N + z[S - 2]] - lzb[T - 1];
  }
  __syncthreads();
  double CB = 0.0;
  par = 0;
  {
    const float K = (float)(p.cA[(size_t)b * T + T - 1] + 0.0 - ll);
    float lm = kNegInf, ps = 0.f;
    for (int s = tid; s < S; s += nthr) {
      const float bt = rp[s];
      lm = fmaxf(lm, bt);
      const float lp = eb[(size_t)(T - 1) * N + z[s]] - lzb[T - 1];
      const float q = lat[(size_t)(T - 1) * Sp + s] + bt - lp + K;
      const float post = (bt == kNegInf) ? 0.f : __expf(q);
      lat[(size_t)(T - 1) * Sp + s] = post;
      ps += post;
    }
    lm = warp_max(lm);
    ps = warp_sum(ps);
    if (lane == 0) {
      wmax[0][warp] = lm;
      wsum[0][warp] = ps;
    }
  }
  __syncthreads();
  // one commit group per frame: activations and the stored alpha row of this thread's states
  auto issue_b = [&](int f) {
    if (f >= 0) {
      const size_t slot = (size_t)(f & (kCtcRing - 1)) * Sp;
      const float* src = eb + (size_t)f * N;
      const float* lsrc = lat + (size_t)f * Sp;
      for (int s = tid; s < S; s += nthr) {
        ctc_cp_async4(ering + slot + s, src + z[s]);
        ctc_cp_async4(lring + slot + s, lsrc + s);
      }
    }
    ctc_cp_commit();
  };
  for (int q = 0; q < kCtcDepth; ++q) issue_b(T - 2 - q);
  for (int t = T - 2; t >= 0; --t) {
    ctc_cp_wait<kCtcDepth - 1>();
    issue_b(t - kCtcDepth);
    float d = wmax[par][0], psm = wsum[par][0];
#pragma unroll
    for (int w = 1; w < nw; ++w) {
      d = fmaxf(d, wmax[par][w]);
      psm += wsum[par][w];
    }
    if (tid == 0) p.psum[(size_t)b * T + t + 1] = psm;
    if (!(d > -1e30f)) d = 0.f;
    CB += (double)d;
    const float lz_t = ts ? lz_s[t] : lzb[t];
    const float K = (float)((ts ? cA_s[t] : p.cA[(size_t)b * T + t]) + CB - ll);
    float lm = kNegInf, ps = 0.f;
    const size_t slot = (size_t)(t & (kCtcRing - 1)) * Sp;
    for (int s = tid; s < S; s += nthr) {
      const float b2 = (s + 2 < S && skip[s + 2]) ? rp[s + 2] : kNegInf;
      const float v = lse3f(rp[s], rp[s + 1], b2);
      const float lp = ering[slot + s] - lz_t;
      const float val = (v == kNegInf) ? kNegInf : v + (lp - d);
      rn[s] = val;
      lm = fmaxf(lm, val);
      const float q = lring[slot + s] + val - lp + K;
      const float post = (val == kNegInf) ? 0.f : __expf(q);
      lat[(size_t)t * Sp + s] = post;
      ps += post;
    }
    lm = warp_max(lm);
    ps = warp_sum(ps);
    if (lane == 0) {
      wmax[par ^ 1][warp] = lm;
      wsum[par ^ 1][warp] = ps;
    }
    __syncthreads();
    float* tmp = rp;
    rp = rn;
    rn = tmp;
    par ^= 1;
  }
  ctc_cp_wait<0>();
  if (tid == 0) {
    float psm = 0.f;
    for (int w = 0; w < nw; ++w) psm += wsum[par][w];
    p.psum[(size_t)b * T] = psm;
  }
}

// one CTA per frame: gradient = coef * (softmax - occupancy)
__global__ void __launch_bounds__(256) ctc_grad_kernel(CtcParams p) {
  const long long f = blockIdx.x;  // frame index b*T + t
  const int b = (int)(f / p.T);
  const int N = p.N, Sp = p.Sp;
  const float* e = p.emis + f * N;
  float* de = p.d_emis + f * N;
  if (!p.valid[b]) {
    for (int k = threadIdx.x; k < N; k += blockDim.x) de[k] = 0.f;
    return;
  }
  const float coef = p.coef[b];
  const float lz = p.lz[f];
  for (int k = threadIdx.x; k < N; k += blockDim.x) de[k] = coef * __expf(__ldg(e + k) - lz);
  __syncthreads();
  const int S = 2 * p.tsz[b] + 1;
  const float ps = p.psum[f];
  const float inv = ps > 0.f ? coef / ps : 0.f;
  const float* post = p.lat + f * Sp;
  const int32_t* yg = p.target ? p.target + (size_t)b * p.L : nullptr;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float v = post[s];
    if (v != 0.f) {
      const int zs = (s & 1) ? yg[s >> 1] : N - 1;
      atomicAdd(de + zs, -v * inv);
    }
  }
}

void carve(CtcParams& p, void* ws, size_t& total) {
  Carver c(ws);
  const size_t BT = (size_t)p.B * p.T;
  p.lz = c.take<float>(BT);
  p.lat = c.take<float>(BT * p.Sp);
  p.cA = c.take<double>(BT);
  p.psum = c.take<float>(BT);
  p.tsz = c.take<int>(p.B);
  p.valid = c.take<int>(p.B);
  p.coef = c.take<float>(p.B);
  p.scale = c.take<float>(p.B);
  total = c.off;
}

int ctc_sp(int T, int L) {
  int Le = L < T ? L : T;
  if (Le < 0) Le = 0;
  return (int)align_up((size_t)(2 * Le + 1), 32);
}

}  // namespace
}  // namespace w2l

using namespace w2l;

extern "C" size_t w2l_ctc_workspace_size(int B, int T, int N, int L) {
  if (B <= 0 || T <= 0 || N <= 1) return 0;
  CtcParams p{};
  p.B = B;
  p.T = T;
  p.N = N;
  p.Sp = ctc_sp(T, L);
  size_t total = 0;
  carve(p, nullptr, total);
  return total;
}

extern "C" int w2l_ctc_forward_backward(void* stream_, int B, int T, int N, int L, int scale_mode, const float* emis,
                                        const int32_t* target, const float* dloss, float* loss, float* d_emis,
                                        void* workspace, size_t workspace_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B <= 0 || T <= 0 || N <= 1) return fail(W2L_ERR_INVALID_ARGUMENT, "ctc: B, T must be positive and N >= 2");
  if (!emis || !loss) return fail(W2L_ERR_INVALID_ARGUMENT, "ctc: null emissions/loss");
  if (L < 0 || (L > 0 && !target)) return fail(W2L_ERR_INVALID_ARGUMENT, "ctc: target pointer/width mismatch");
  if (scale_mode < 0 || scale_mode > 4) return fail(W2L_ERR_INVALID_ARGUMENT, "ctc: bad scale mode");
  CtcParams p{};
  p.B = B;
  p.T = T;
  p.N = N;
  p.L = target ? L : 0;
  p.Sp = ctc_sp(T, p.L);
  p.scale_mode = scale_mode;
  p.need_grad = d_emis != nullptr;
  p.emis = emis;
  p.target = p.L > 0 ? target : nullptr;
  p.dloss = dloss;
  p.loss = loss;
  p.d_emis = d_emis;
  size_t need = 0;
  carve(p, workspace, need);
  if (!workspace || workspace_bytes < need)
    return fail(W2L_ERR_WORKSPACE, "ctc: workspace too small (need " + std::to_string(need) + " bytes)");
  const size_t smem_base = (size_t)(2 * (p.Sp + 8) + p.Sp) * 4 + p.Sp + 64;
  int depth = 8;
  while (depth > 1 && smem_base + (size_t)4 * depth * p.Sp * 4 > 200 * 1024) depth >>= 1;
  size_t smem = smem_base + (size_t)4 * depth * p.Sp * 4;  // two rings of 2*depth slots
  if (smem > 200 * 1024) return fail(W2L_ERR_UNSUPPORTED, "ctc: target too long for the shared-memory rows");
  p.t_smem = smem + (size_t)T * 12 <= 200 * 1024 ? 1 : 0;
  if (p.t_smem) smem += (size_t)T * 12;
  const long long nframes = (long long)B * T;
  const int frame_blocks = (int)std::min<long long>((nframes + 7) / 8, 148 * 8);
  ctc_prep_kernel<<<frame_blocks + (B + 255) / 256, 256, 0, stream>>>(p, frame_blocks);
  W2L_LAUNCH_CHECK("ctc_prep_kernel");
  if (p.need_grad) {
#define W2L_CTC_LAUNCH(GRAD, D)                                                                                               \
  do {                                                                                                                         \
    if (smem > 48 * 1024)                                                                                                      \
      W2L_CUDA_CHECK(cudaFuncSetAttribute(ctc_chains_kernel<GRAD, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ctc_chains_kernel<GRAD, D><<<B, std::min(kCtcThreads, p.Sp), smem, stream>>>(p);                                                          \
  } while (0)
#define W2L_CTC_DISPATCH(GRAD)                \
  do {                                        \
    if (depth == 8) W2L_CTC_LAUNCH(GRAD, 8);  \
    else if (depth == 4) W2L_CTC_LAUNCH(GRAD, 4); \
    else if (depth == 2) W2L_CTC_LAUNCH(GRAD, 2); \
    else W2L_CTC_LAUNCH(GRAD, 1);             \
  } while (0)
    profile_kind(2);
    profile_start(stream);
    W2L_CTC_DISPATCH(true);
    profile_stop(stream);
    W2L_LAUNCH_CHECK("ctc_chains_kernel<grad>");
    ctc_grad_kernel<<<(unsigned)nframes, 256, 0, stream>>>(p);
    W2L_LAUNCH_CHECK("ctc_grad_kernel");
  } else {
    W2L_CTC_DISPATCH(false);
    W2L_LAUNCH_CHECK("ctc_chains_kernel<fwd>");
  }
  return W2L_OK;
}
