// criterion_asg.cu — fused ASG (FullConnectionCriterion - ForceAlignmentCriterion) forward +
// backward for sm_100a.  Replaces flashlight-0.3 lib/sequence/criterion/cuda/
// {FullConnectionCriterion,ForceAlignmentCriterion}.cu as reached from
// recipes/slimIPL/src/Train.cpp:408-410 (construction), :1675 (forward), :1720 (backward).
//
// Pipeline (3 launches + 1 memset, all on the caller's stream):
//   1. asg_prep_kernel    HBM-bound, parallel over frames: m_t = max_i e_t[i],
//                         X_t[i] = exp(e_t[i]-m_t) (padded to 32 lanes); per-sample target
//                         size, validity, scale*dloss.
//   2. asg_chains_kernel  latency-bound, one CTA per (sample, criterion):
//        FCC CTA: warp 0 walks alpha (t = 0..T-1), warp 1 walks beta (t = T-1..0) at the same
//                 time, in the LINEAR domain: a_t = X_t .* (M' a_{t-1}) * 2^-k with M' =
//                 exp(trans - max trans) held in registers (row i in lane i), the vector
//                 exchanged through shared memory (1 STS + 8 broadcast LDS.128), 16 FFMA2 per
//                 step, and a power-of-two rescale taken from the exponent bits of
//                 max(a_{t-2}) (lagged, so no reduction sits on the dependent chain; exact).
//        FAC CTA: 128 threads walk alpha from t=0 and 128 walk beta from t=T-1 in the LOG
//                 domain (the left-to-right band has unbounded dynamic range, a linear-domain
//                 form is not safe there), re-centred every step by the band maximum; they
//                 meet at h = T/2, the partition function is taken at the junction, and each
//                 group finishes its walk reading the other group's stored half lattice to
//                 emit occupancies (gamma = xi_stay + xi_adv) and transition statistics.
//   3. asg_grad_kernel    parallel over (sample, frame chunk): gamma_fcc = a.*b / sum,
//                         d_emis = coef*(gamma_fcc - gamma_fac), d_trans += M' .* sum_t w_t a_{t-1}^T.
#include <cuda_runtime.h>

#include "common.cuh"

namespace w2l {
namespace {

constexpr int kW = 32;            // padded FCC state width (one lane per state)
constexpr int kChainThreads = 256;
constexpr int kGroup = 128;       // threads per FAC direction
constexpr int kGroupWarps = kGroup / 32;
constexpr int kPrefetch = 8;      // frames of X prefetched ahead of the FCC chains
constexpr int kGradChunk = 32;    // frames per warp in the grad kernel

struct AsgParams {
  int B, T, N, L, Lp, scale_mode, terms, h, need_grad;
  const float* emis;
  const int32_t* target;
  const float* trans;
  const float* dloss;
  float* loss;
  float* d_emis;
  float* d_trans;
  // workspace
  float* X;       // [B][T][32]
  float* mrow;    // [B][T]
  float* A;       // [B][T][32] FCC alpha-hat
  float* Bh;      // [B][T][32] FCC beta-hat
  float* sA;      // [B][T] power-of-two scale applied at step t of the alpha walk
  float* G;       // [B][T][32] FAC occupancy per label
  float* facA;    // [B][h][Lp]    stored alpha-tilde rows, t < h
  float* facB;    // [B][T-h][Lp]  stored beta-tilde rows,  t >= h
  double* cA;     // [B][T] re-centring offsets of the FAC alpha walk
  double* cB;     // [B][T]
  double* fccLogZ;  // [B]
  double* facLogZ;  // [B]
  int* tsz;       // [B]
  int* valid;     // [B]
  float* scale;   // [B]
  float* coef;    // [B] scale * dloss
};

// ------------------------------------------------------------------------------------------
// 1. prep
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) asg_prep_kernel(AsgParams p, int frame_blocks) {
  const int lane = threadIdx.x & 31;
  if ((int)blockIdx.x < frame_blocks) {
    if (!(p.terms & W2L_TERM_FCC)) return;
    const long long nframes = (long long)p.B * p.T;
    const long long warps = (long long)frame_blocks * (blockDim.x >> 5);
    for (long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < nframes; f += warps) {
      float e = lane < p.N ? __ldg(p.emis + f * p.N + lane) : kNegInf;
      float m = warp_max(e);
      float x = lane < p.N ? __expf(e - m) : 0.0f;
      p.X[f * kW + lane] = x;
      if (lane == 0) p.mrow[f] = m;
    }
    return;
  }
  // per-sample metadata
  int b = ((int)blockIdx.x - frame_blocks) * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  int tsz = 0, ok = 1;
  if (p.target != nullptr && p.L > 0) {
    const int32_t* y = p.target + (size_t)b * p.L;
    tsz = target_size(y, p.L, p.T);
    if (p.terms & W2L_TERM_FAC) {
      if (tsz <= 0) ok = 0;
      for (int l = 0; l < tsz; ++l)
        if (y[l] < 0 || y[l] >= p.N) ok = 0;
    }
  } else if (p.terms & W2L_TERM_FAC) {
    ok = 0;
  }
  float sc = scale_of(p.scale_mode, p.T, tsz);
  p.tsz[b] = tsz;
  p.valid[b] = ok;
  p.scale[b] = sc;
  p.coef[b] = ok ? sc * (p.dloss ? p.dloss[b] : 1.0f) : 0.0f;
}

// ------------------------------------------------------------------------------------------
// 2a. FCC chains (linear domain)
// ------------------------------------------------------------------------------------------
// acc = sum_j M[j] * v[j] over the 32 shared-memory entries (broadcast LDS.128), mx = max_j v[j]
__device__ __forceinline__ void matvec32(const float (&M)[kW], const float* vsm, float& acc, float& mx) {
  const float4* v4 = reinterpret_cast<const float4*>(vsm);
  float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  float m0 = 0.f, m1 = 0.f;
#pragma unroll
  for (int q = 0; q < kW / 4; ++q) {
    float4 v = v4[q];
    float2 lo = make_float2(v.x, v.y), hi = make_float2(v.z, v.w);
    float2 mlo = make_float2(M[4 * q], M[4 * q + 1]), mhi = make_float2(M[4 * q + 2], M[4 * q + 3]);
    if (q & 1) {
      a2 = __ffma2_rn(mlo, lo, a2);
      a3 = __ffma2_rn(mhi, hi, a3);
    } else {
      a0 = __ffma2_rn(mlo, lo, a0);
      a1 = __ffma2_rn(mhi, hi, a1);
    }
    m0 = fmaxf(fmaxf(m0, v.x), v.y);
    m1 = fmaxf(fmaxf(m1, v.z), v.w);
  }
  float2 s = __fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3));
  acc = s.x + s.y;
  mx = fmaxf(m0, m1);
}

// 2^-k with k = unbiased exponent of mx (clamped); returns k through kout.  Exact.
__device__ __forceinline__ float pow2_rescale(float mx, int& kout) {
  int k = ((__float_as_int(mx) >> 23) & 0xff) - 127;
  k = max(-126, min(126, k));
  kout = k;
  return __int_as_float((127 - k) << 23);
}

template <bool kGrad>
__device__ void fcc_role(const AsgParams& p, int b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ __align__(16) float vec[2][2][kW];
  __shared__ double msum_s;
  const int T = p.T, N = p.N;
  if (warp == 2) {  // sum of the per-frame maxima, off the chains
    double s = 0.0;
    for (int t = lane; t < T; t += 32) s += (double)p.mrow[(size_t)b * T + t];
    s = warp_sum(s);
    if (lane == 0) msum_s = s;
    __threadfence_block();
    asm volatile("bar.arrive 1, 64;" ::: "memory");
    return;
  }
  if (warp > 2 || (warp == 1 && !kGrad)) return;

  // transitions: global max, then M' row (alpha walk) or column (beta walk) into registers
  float tmax = kNegInf;
  for (int j = 0; j < N; ++j) {
    float v = lane < N ? __ldg(p.trans + lane * N + j) : kNegInf;
    tmax = fmaxf(tmax, v);
  }
  tmax = warp_max(tmax);
  float M[kW];
#pragma unroll
  for (int j = 0; j < kW; ++j) {
    float v = 0.f;
    if (lane < N && j < N) v = __expf(__ldg(p.trans + (warp == 0 ? lane * N + j : j * N + lane)) - tmax);
    M[j] = v;
  }
  const float* Xb = p.X + (size_t)b * T * kW;
  float xq[kPrefetch];

  if (warp == 0) {
    // ---- alpha walk: a_t = (X_t * s_t) .* (M' a_{t-1}) ----------------------------------------
    float* Ab = p.A + (size_t)b * T * kW;
    float* sAb = p.sA + (size_t)b * T;
    float a = Xb[lane];
    if (kGrad) {
      Ab[lane] = a;
      if (lane == 0) sAb[0] = 1.0f;
    }
#pragma unroll
    for (int q = 0; q < kPrefetch; ++q) xq[q] = (1 + q < T) ? Xb[(size_t)(1 + q) * kW + lane] : 0.f;
    float s = 1.0f;
    int ksum = 0, kcur = 0;
    int buf = 0;
    for (int t0 = 1; t0 < T; t0 += kPrefetch) {
#pragma unroll
      for (int q = 0; q < kPrefetch; ++q) {
        const int t = t0 + q;
        if (t < T) {
          vec[0][buf][lane] = a;
          __syncwarp();
          float acc, mx;
          matvec32(M, vec[0][buf], acc, mx);
          const float xs = xq[q] * s;
          a = xs * acc;
          ksum += kcur;
          if (kGrad) {
            Ab[(size_t)t * kW + lane] = a;
            if (lane == 0) sAb[t] = s;
          }
          s = pow2_rescale(mx, kcur);  // applied at t+1; normalises by |a_{t-1}|
          xq[q] = (t + kPrefetch < T) ? Xb[(size_t)(t + kPrefetch) * kW + lane] : 0.f;
          buf ^= 1;
        }
      }
    }
    const float tot = warp_sum(a);
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (lane == 0) {
      p.fccLogZ[b] = msum_s + (double)(T - 1) * (double)tmax + 0.6931471805599453 * (double)ksum +
                     log((double)tot);
    }
  } else {
    // ---- beta walk: b_t = M'^T (X_{t+1} .* b_{t+1} * s) -----------------------------------------
    float* Bb = p.Bh + (size_t)b * T * kW;
    float bh = lane < N ? 1.0f : 0.0f;
    Bb[(size_t)(T - 1) * kW + lane] = bh;
#pragma unroll
    for (int q = 0; q < kPrefetch; ++q) xq[q] = (T - 1 - q >= 1) ? Xb[(size_t)(T - 1 - q) * kW + lane] : 0.f;
    float s = 1.0f;
    int kdummy;
    int buf = 0;
    for (int t0 = T - 2; t0 >= 0; t0 -= kPrefetch) {
#pragma unroll
      for (int q = 0; q < kPrefetch; ++q) {
        const int t = t0 - q;
        if (t >= 0) {
          const float u = bh * (xq[q] * s);  // uses X_{t+1}
          vec[1][buf][lane] = u;
          __syncwarp();
          float acc, mx;
          matvec32(M, vec[1][buf], acc, mx);
          bh = acc;
          Bb[(size_t)t * kW + lane] = bh;
          s = pow2_rescale(mx, kdummy);
          xq[q] = (t + 1 - kPrefetch >= 1) ? Xb[(size_t)(t + 1 - kPrefetch) * kW + lane] : 0.f;
          buf ^= 1;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// 2b. FAC chains (log domain, meet in the middle)
// ------------------------------------------------------------------------------------------
struct FacSmem {
  int32_t* y;
  float* s1;
  float* s2;
  float* rowA[2];  // index l in [-1, Lp]
  float* rowB[2];
  float* ds1[2];  // [group]
  float* ds2[2];
  float* bins;   // [2 groups][2 bufs][32]
  float* wmax;   // [2 groups][2 bufs][4]
  float* ring;   // [2 groups][4][32]
  float* rnorm;  // [2 groups][2]
  float* red;    // [16]
  double* dred;  // [4]
};

// floats: red 16, bins 128, wmax 16, ring 256, rnorm 4, y/s1/s2 3*(Lp+4), rows 4*(Lp+4), ds 4*Lp; + 4 doubles
__host__ __device__ inline size_t fac_smem_bytes(int Lp) {
  return (size_t)(16 + 128 + 16 + 256 + 4 + 3 * (Lp + 4) + 4 * (Lp + 4) + 4 * Lp) * 4 + 4 * 8;
}

__device__ inline FacSmem fac_carve(unsigned char* raw, int Lp) {
  FacSmem s;
  double* d = reinterpret_cast<double*>(raw);
  s.dred = d;
  float* f = reinterpret_cast<float*>(d + 4);
  s.red = f;
  f += 16;
  s.bins = f;
  f += 2 * 2 * 32;
  s.wmax = f;
  f += 2 * 2 * 4;
  s.ring = f;
  f += 2 * 4 * 32;
  s.rnorm = f;
  f += 4;
  s.y = reinterpret_cast<int32_t*>(f);
  f += Lp + 4;
  s.s1 = f;
  f += Lp + 4;
  s.s2 = f;
  f += Lp + 4;
  for (int k = 0; k < 2; ++k) {
    s.rowA[k] = f + 2;
    f += Lp + 4;
  }
  for (int k = 0; k < 2; ++k) {
    s.rowB[k] = f + 2;
    f += Lp + 4;
  }
  for (int k = 0; k < 2; ++k) {
    s.ds1[k] = f;
    f += Lp;
    s.ds2[k] = f;
    f += Lp;
  }
  return s;
}

__device__ __forceinline__ float max4_guard(const float* w) {
  float d = fmaxf(fmaxf(w[0], w[1]), fmaxf(w[2], w[3]));
  return (d > -1e30f) ? d : 0.0f;
}

template <bool kGrad>
__device__ void fac_role(const AsgParams& p, int b, unsigned char* smem_raw) {
  const int tid = threadIdx.x;
  const int grp = tid / kGroup;       // 0: alpha walk, 1: beta walk
  const int gt = tid % kGroup;        // thread within group
  const int gw = gt >> 5, lane = tid & 31;
  const int T = p.T, N = p.N, Lp = p.Lp;
  const int L = p.tsz[b];
  FacSmem sm = fac_carve(smem_raw, Lp);
  const float* eb = p.emis + (size_t)b * T * N;
  float* Gb = p.G + (size_t)b * T * kW;

  if (!p.valid[b]) {
    if (tid == 0) p.facLogZ[b] = (double)NAN;
    return;  // whole CTA (uniform)
  }
  const int32_t* yg = p.target + (size_t)b * p.L;
  for (int l = tid; l < Lp; l += kChainThreads) {
    int yl = l < L ? yg[l] : 0;
    sm.y[l] = yl;
    sm.s1[l] = l < L ? __ldg(p.trans + yl * N + yl) : 0.f;
    sm.s2[l] = (l < L && l > 0) ? __ldg(p.trans + yl * N + yg[l - 1]) : kNegInf;
    if (kGrad) {
      sm.ds1[0][l] = sm.ds1[1][l] = 0.f;
      sm.ds2[0][l] = sm.ds2[1][l] = 0.f;
    }
  }
  for (int l = tid; l < Lp + 4; l += kChainThreads) {
    sm.rowA[0][l - 2] = sm.rowA[1][l - 2] = kNegInf;
    sm.rowB[0][l - 2] = sm.rowB[1][l - 2] = kNegInf;
  }
  if (tid < 2 * 2 * 32) sm.bins[tid] = 0.f;
  if (tid < 4) sm.rnorm[tid] = 1.0f;
  __syncthreads();

  if (T == 1) {  // single frame: the only alignment is (0,0); L was clamped to 1
    if (tid == 0) p.facLogZ[b] = (double)eb[sm.y[0]];
    if (kGrad && tid < kW) Gb[tid] = (tid == sm.y[0]) ? 1.0f : 0.0f;
    return;
  }
  const int h = kGrad ? p.h : T;  // forward only: the alpha group walks the whole sequence
  float* row[2] = {grp == 0 ? sm.rowA[0] : sm.rowB[0], grp == 0 ? sm.rowA[1] : sm.rowB[1]};
  float* wmax = sm.wmax + grp * 8;
  float* ring = sm.ring + grp * 4 * 32;
  float* bins = sm.bins + grp * 64;
  float* rnorm = sm.rnorm + grp * 2;
  float* ds1 = sm.ds1[grp];
  float* ds2 = sm.ds2[grp];
  const int bar_id = 2 + grp;
  const int dir = grp == 0 ? 1 : -1;
  double C = 0.0;  // re-centring offset of the current row (uniform across the group)
  int cur = 0;

  // ring[t & 3][i] = e_t[i]; preload the first three frames of this group's walk
  const int t_first = grp == 0 ? 0 : T - 1;
  if (gt < N) {
    for (int q = 0; q < 3; ++q) {
      int t = t_first + dir * q;
      if (t >= 0 && t < T) ring[(t & 3) * 32 + gt] = __ldg(eb + (size_t)t * N + gt);
    }
  }
  named_barrier_sync(bar_id, kGroup);

  // ---- initial row ------------------------------------------------------------------------
  {
    float lmax = kNegInf;
    if (grp == 0) {
      if (gt == 0) {
        float v = ring[(0 & 3) * 32 + sm.y[0]];
        row[0][0] = v;
        lmax = v;
        if (kGrad) p.facA[((size_t)b * p.h + 0) * Lp + 0] = v;
      }
      if (kGrad)
        for (int l = gt + (gt == 0 ? kGroup : 0); l < L; l += kGroup) p.facA[((size_t)b * p.h + 0) * Lp + l] = kNegInf;
      if (kGrad && gt == 0) p.cA[(size_t)b * T] = 0.0;
    } else {
      if (gt == 0) {
        float v = ring[((T - 1) & 3) * 32 + sm.y[L - 1]];
        row[0][L - 1] = v;
        lmax = v;
      }
      for (int l = gt; l < L; l += kGroup)
        p.facB[((size_t)b * (T - p.h) + (T - 1 - p.h)) * Lp + l] = (l == L - 1) ? ring[((T - 1) & 3) * 32 + sm.y[L - 1]] : kNegInf;
      if (gt == 0) p.cB[(size_t)b * T + T - 1] = 0.0;
    }
    float wm = warp_max(lmax);
    if (lane == 0) wmax[0 * 4 + gw] = wm;
    named_barrier_sync(bar_id, kGroup);
  }

  // one step of this group's walk: computes row[cur^1] at frame t from row[cur]; optionally
  // stores the new row to the half lattice.  Returns nothing; state in smem.
  auto step = [&](int t, bool store) {
    const int prev = cur, nxt = cur ^ 1;
    const float delta = max4_guard(wmax + prev * 4);
    C += (double)delta;
    // stage the frame two steps ahead into the ring
    const int tp = t + 2 * dir;
    float pf = 0.f;
    const bool do_pf = gt < N && tp >= 0 && tp < T;
    if (do_pf) pf = __ldg(eb + (size_t)tp * N + gt);
    const int lo = max(0, L - (T - t)), hi = min(t, L - 1);
    const float* rp = row[prev];
    float* rn = row[nxt];
    const float* fr = ring + (t & 3) * 32;
    float lmax = kNegInf;
    for (int l = gt; l < L; l += kGroup) {
      float val = kNegInf;
      if (l >= lo && l <= hi) {
        float a0, a1;
        if (grp == 0) {
          a0 = rp[l] + (sm.s1[l] - delta);
          a1 = rp[l - 1] + (sm.s2[l] - delta);
        } else {
          a0 = rp[l] + (sm.s1[l] - delta);
          a1 = (l + 1 < L) ? rp[l + 1] + (sm.s2[l + 1] - delta) : kNegInf;
        }
        val = fr[sm.y[l]] + lse2f(a0, a1);
      }
      rn[l] = val;
      lmax = fmaxf(lmax, val);
      if (store) {
        if (grp == 0)
          p.facA[((size_t)b * p.h + t) * Lp + l] = val;
        else
          p.facB[((size_t)b * (T - p.h) + (t - p.h)) * Lp + l] = val;
      }
    }
    float wm = warp_max(lmax);
    if (lane == 0) wmax[nxt * 4 + gw] = wm;
    if (store && gt == 0) (grp == 0 ? p.cA : p.cB)[(size_t)b * T + t] = C;
    if (do_pf) ring[(tp & 3) * 32 + gt] = pf;
    cur = nxt;
  };

  // ---- phase 1: walk to the middle, storing the half lattices ---------------------------------
  if (grp == 0) {
    for (int t = 1; t < h; ++t) {
      step(t, kGrad);
      named_barrier_sync(bar_id, kGroup);
    }
  } else if (kGrad) {
    for (int t = T - 2; t >= h; --t) {
      step(t, true);
      named_barrier_sync(bar_id, kGroup);
    }
  }
  if (!kGrad) {
    if (grp == 0 && gt == 0) p.facLogZ[b] = (double)row[cur][L - 1] + C;
    return;
  }
  __syncthreads();

  // ---- junction at t = h: alpha group computes alpha_h; partition function from alpha_h + beta_h
  // beta group's current row (frame h) lives in sm.rowB[curB]; both groups ran the same number of
  // steps when T is even, one apart when odd -> publish the buffer index.
  __shared__ int curB_s;
  __shared__ double CB_h_s, logZ_s;
  if (grp == 1 && gt == 0) {
    curB_s = cur;
    CB_h_s = C;
  }
  __syncthreads();
  if (grp == 0) {
    const float* rb = sm.rowB[curB_s];
    const double C_prev = C;
    (void)C_prev;
    step(h, false);  // row[cur] = alpha-tilde_h, offset C
    // the ring slot of frame h was loaded by this group (alpha ring holds frames h-1.. h+2)
    named_barrier_sync(bar_id, kGroup);
    const float* ra = row[cur];
    const float* fr = ring + (h & 3) * 32;
    float qmax = kNegInf;
    for (int l = gt; l < L; l += kGroup) qmax = fmaxf(qmax, ra[l] + rb[l] - fr[sm.y[l]]);
    qmax = warp_max(qmax);
    if (lane == 0) sm.red[gw] = qmax;
    named_barrier_sync(bar_id, kGroup);
    qmax = fmaxf(fmaxf(sm.red[0], sm.red[1]), fmaxf(sm.red[2], sm.red[3]));
    float part = 0.f;
    for (int l = gt; l < L; l += kGroup) {
      float q = ra[l] + rb[l] - fr[sm.y[l]];
      part += (q == kNegInf) ? 0.f : __expf(q - qmax);
    }
    part = warp_sum(part);
    if (lane == 0) sm.red[4 + gw] = part;
    named_barrier_sync(bar_id, kGroup);
    const float tot = sm.red[4] + sm.red[5] + sm.red[6] + sm.red[7];
    if (gt == 0) {
      double lz = C + CB_h_s + (double)qmax + log((double)tot);
      logZ_s = lz;
      p.facLogZ[b] = lz;
    }
    // occupancy of frame h
    float* bn = bins + (h & 1) * 32;
    const float inv = 1.0f / tot;
    for (int l = gt; l < L; l += kGroup) {
      float q = ra[l] + rb[l] - fr[sm.y[l]];
      float g = (q == kNegInf) ? 0.f : __expf(q - qmax) * inv;
      if (g > 0.f) atomicAdd(&bn[sm.y[l]], g);
    }
    named_barrier_sync(bar_id, kGroup);
    if (gt < kW) {
      Gb[(size_t)h * kW + gt] = bn[gt];
      bn[gt] = 0.f;
    }
  }
  __syncthreads();
  const double logZ = logZ_s;

  // ---- phase 2: finish the walks, emitting occupancies and transition statistics --------------
  // alpha group: t = h+1 .. T-1, transitions (t-1 -> t), reads stored beta-tilde_t
  // beta  group: t = h-1 .. 0,   transitions (t -> t+1), reads stored alpha-tilde_t
  auto phase2_step = [&](int t) {
    const int prev = cur;  // row[prev]: alpha_{t-1} (grp 0) or beta_{t+1} (grp 1), offset C
    const double C_prev = C;
    const float* rp = row[prev];
    float K;
    const float* other;
    if (grp == 0) {
      K = (float)(C_prev + p.cB[(size_t)b * T + t] - logZ);
      other = p.facB + ((size_t)b * (T - p.h) + (t - p.h)) * Lp;
    } else {
      K = (float)(p.cA[(size_t)b * T + t] + C_prev - logZ);
      other = p.facA + ((size_t)b * p.h + t) * Lp;
    }
    float* bn = bins + (t & 1) * 32;
    const float rn_lag = rnorm[t & 1];  // normaliser of two steps ago (see flush below)
    for (int l = gt; l < L; l += kGroup) {
      const float o = other[l];  // written by the other group in phase 1: plain (coherent) load
      float xs, xa;
      if (grp == 0) {
        xs = __expf(rp[l] + sm.s1[l] + o + K);
        xa = __expf(rp[l - 1] + sm.s2[l] + o + K);
        ds1[l] += xs * rn_lag;
        ds2[l] += xa * rn_lag;
      } else {
        xs = __expf(o + sm.s1[l] + rp[l] + K);
        xa = (l + 1 < L) ? __expf(o + sm.s2[l + 1] + rp[l + 1] + K) : 0.f;
        ds1[l] += xs * rn_lag;
        if (l + 1 < L) ds2[l + 1] += xa * rn_lag;
      }
      const float g = xs + xa;
      if (g > 0.f) atomicAdd(&bn[sm.y[l]], g);
    }
    step(t, false);
    named_barrier_sync(bar_id, kGroup);
    // flush this frame's occupancy, renormalised so that it sums to one (removes the common-mode
    // rounding error of the fp32 log-domain walk); the normaliser is reused two steps later for the
    // transition statistics.
    if (gw == 0) {
      float v = bn[lane];
      float tot = warp_sum(v);
      float inv = tot > 0.f ? 1.0f / tot : 0.f;
      Gb[(size_t)t * kW + lane] = v * inv;
      bn[lane] = 0.f;
      if (lane == 0) rnorm[t & 1] = inv;
    }
  };
  if (grp == 0) {
    for (int t = h + 1; t < T; ++t) phase2_step(t);
  } else {
    for (int t = h - 1; t >= 0; --t) phase2_step(t);
  }
  __syncthreads();
  // transition-gradient scatter: FAC enters ASG with a minus sign
  if (p.d_trans != nullptr) {
    const float sgn = (p.terms & W2L_TERM_FCC) ? -1.0f : 1.0f;
    const float c = sgn * p.coef[b];
    for (int l = tid; l < L; l += kChainThreads) {
      const int yl = sm.y[l];
      float v1 = sm.ds1[0][l] + sm.ds1[1][l];
      if (v1 != 0.f) atomicAdd(p.d_trans + yl * N + yl, c * v1);
      if (l > 0) {
        float v2 = sm.ds2[0][l] + sm.ds2[1][l];
        if (v2 != 0.f) atomicAdd(p.d_trans + yl * N + sm.y[l - 1], c * v2);
      }
    }
  }
}

template <bool kGrad>
__global__ void __launch_bounds__(kChainThreads) asg_chains_kernel(AsgParams p, int n_fac) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if ((int)blockIdx.x < n_fac) {
    fac_role<kGrad>(p, blockIdx.x, smem_raw);
  } else {
    fcc_role<kGrad>(p, blockIdx.x - n_fac);
  }
}

// ------------------------------------------------------------------------------------------
// 3. gradient assembly (parallel)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) asg_grad_kernel(AsgParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int chunk = blockIdx.x * 4 + warp;
  const int T = p.T, N = p.N;
  const int t0 = chunk * kGradChunk, t1 = min(T, t0 + kGradChunk);
  const bool has_fcc = p.terms & W2L_TERM_FCC, has_fac = p.terms & W2L_TERM_FAC;
  const int ok = p.valid[b];
  if (chunk == 0 && lane == 0) {
    float l = NAN;
    if (ok) {
      double v = 0.0;
      if (has_fcc) v += p.fccLogZ[b];
      if (has_fac) v += (has_fcc ? -1.0 : 1.0) * p.facLogZ[b];
      l = (float)((double)p.scale[b] * v);
    }
    p.loss[b] = l;
  }
  if (t0 >= T || p.d_emis == nullptr) return;
  float* de = p.d_emis + (size_t)b * T * N;
  if (!ok) {
    for (int t = t0; t < t1; ++t)
      if (lane < N) de[(size_t)t * N + lane] = 0.f;
    return;
  }
  __shared__ __align__(16) float aprev_s[4][kW];
  const float coef = p.coef[b];
  const float sG = has_fac ? (has_fcc ? -1.f : 1.f) : 0.f;
  const float* Ab = p.A + (size_t)b * T * kW;
  const float* Bb = p.Bh + (size_t)b * T * kW;
  const float* Xb = p.X + (size_t)b * T * kW;
  const float* Gb = p.G + (size_t)b * T * kW;
  float acc[kW];
#pragma unroll
  for (int j = 0; j < kW; ++j) acc[j] = 0.f;
  for (int t = t0; t < t1; ++t) {
    float gam = 0.f;
    if (has_fcc) {
      const float a = Ab[(size_t)t * kW + lane];
      const float bh = Bb[(size_t)t * kW + lane];
      const float g = a * bh;
      const float gs = warp_sum(g);
      gam = g / gs;
      if (t >= 1 && p.d_trans != nullptr) {
        const float w = Xb[(size_t)t * kW + lane] * bh * (p.sA[(size_t)b * T + t] / gs);
        aprev_s[warp][lane] = Ab[(size_t)(t - 1) * kW + lane];
        __syncwarp();
        const float4* v4 = reinterpret_cast<const float4*>(aprev_s[warp]);
#pragma unroll
        for (int q = 0; q < kW / 4; ++q) {
          float4 v = v4[q];
          acc[4 * q + 0] = fmaf(w, v.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(w, v.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(w, v.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(w, v.w, acc[4 * q + 3]);
        }
        __syncwarp();
      }
    }
    const float gf = has_fac ? Gb[(size_t)t * kW + lane] : 0.f;
    if (lane < N) de[(size_t)t * N + lane] = coef * (gam + sG * gf);
  }
  float tmax = kNegInf;
  if (has_fcc && p.d_trans != nullptr) {
    for (int k = lane; k < N * N; k += 32) tmax = fmaxf(tmax, __ldg(p.trans + k));
    tmax = warp_max(tmax);
  }
  if (has_fcc && p.d_trans != nullptr && lane < N) {
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      if (j < N) {
        const float v = coef * acc[j] * __expf(__ldg(p.trans + lane * N + j) - tmax);
        if (v != 0.f) atomicAdd(p.d_trans + lane * N + j, v);
      }
    }
  }
}

__global__ void asg_loss_only_kernel(AsgParams p) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  float l = NAN;
  if (p.valid[b]) {
    double v = 0.0;
    if (p.terms & W2L_TERM_FCC) v += p.fccLogZ[b];
    if (p.terms & W2L_TERM_FAC) v += ((p.terms & W2L_TERM_FCC) ? -1.0 : 1.0) * p.facLogZ[b];
    l = (float)((double)p.scale[b] * v);
  }
  p.loss[b] = l;
}

void carve(AsgParams& p, void* ws, size_t& total) {
  Carver c(ws);
  const size_t BT = (size_t)p.B * p.T;
  p.X = c.take<float>(BT * kW);
  p.mrow = c.take<float>(BT);
  p.A = c.take<float>(BT * kW);
  p.Bh = c.take<float>(BT * kW);
  p.sA = c.take<float>(BT);
  p.G = c.take<float>(BT * kW);
  p.facA = c.take<float>((size_t)p.B * (p.h > 0 ? p.h : 1) * p.Lp);
  p.facB = c.take<float>((size_t)p.B * (p.T - p.h > 0 ? p.T - p.h : 1) * p.Lp);
  p.cA = c.take<double>(BT);
  p.cB = c.take<double>(BT);
  p.fccLogZ = c.take<double>(p.B);
  p.facLogZ = c.take<double>(p.B);
  p.tsz = c.take<int>(p.B);
  p.valid = c.take<int>(p.B);
  p.scale = c.take<float>(p.B);
  p.coef = c.take<float>(p.B);
  total = c.off;
}

}  // namespace
}  // namespace w2l

using namespace w2l;

extern "C" size_t w2l_asg_workspace_size(int B, int T, int N, int L) {
  if (B <= 0 || T <= 0 || N <= 0) return 0;
  AsgParams p{};
  p.B = B;
  p.T = T;
  p.N = N;
  p.L = L;
  int Le = L < T ? L : T;
  if (Le < 1) Le = 1;
  p.Lp = (int)align_up((size_t)Le, 32);
  p.h = T / 2;
  size_t total = 0;
  carve(p, nullptr, total);
  return total;
}

extern "C" int w2l_asg_forward_backward(void* stream_, int terms, int B, int T, int N, int L, int scale_mode,
                                        const float* emis, const int32_t* target, const float* trans,
                                        const float* dloss, float* loss, float* d_emis, float* d_trans,
                                        void* workspace, size_t workspace_bytes) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B <= 0 || T <= 0 || N <= 0) return fail(W2L_ERR_INVALID_ARGUMENT, "asg: B, T, N must be positive");
  if (!(terms & W2L_TERM_ASG) || (terms & ~W2L_TERM_ASG)) return fail(W2L_ERR_INVALID_ARGUMENT, "asg: bad terms");
  if (!emis || !trans || !loss) return fail(W2L_ERR_INVALID_ARGUMENT, "asg: null emissions/transitions/loss");
  if ((terms & W2L_TERM_FAC) && (!target || L <= 0))
    return fail(W2L_ERR_INVALID_ARGUMENT, "asg: ForceAlignment needs a target of width L > 0");
  if ((d_emis == nullptr) != (d_trans == nullptr))
    return fail(W2L_ERR_INVALID_ARGUMENT, "asg: d_emis and d_trans must be given together");
  if (scale_mode < 0 || scale_mode > 4) return fail(W2L_ERR_INVALID_ARGUMENT, "asg: bad scale mode");
  if (N > kW) return fail(W2L_ERR_UNSUPPORTED, "asg: N > 32 tokens is not covered by the sm_100a kernels");
  AsgParams p{};
  p.B = B;
  p.T = T;
  p.N = N;
  p.L = target ? L : 0;
  int Le = p.L < T ? p.L : T;
  if (Le < 1) Le = 1;
  p.Lp = (int)align_up((size_t)Le, 32);
  p.h = T / 2;
  p.scale_mode = scale_mode;
  p.terms = terms;
  p.need_grad = d_emis != nullptr;
  p.emis = emis;
  p.target = target;
  p.trans = trans;
  p.dloss = dloss;
  p.loss = loss;
  p.d_emis = d_emis;
  p.d_trans = d_trans;
  size_t need = 0;
  carve(p, workspace, need);
  if (!workspace || workspace_bytes < need)
    return fail(W2L_ERR_WORKSPACE, "asg: workspace too small (need " + std::to_string(need) + " bytes)");
  const size_t smem = (terms & W2L_TERM_FAC) ? fac_smem_bytes(p.Lp) : 0;
  if (smem > 220 * 1024) return fail(W2L_ERR_UNSUPPORTED, "asg: target too long for the shared-memory rows");

  if (p.need_grad) W2L_CUDA_CHECK(cudaMemsetAsync(d_trans, 0, sizeof(float) * N * N, stream));
  const long long nframes = (long long)B * T;
  int frame_blocks = (int)std::min<long long>((nframes + 7) / 8, 148 * 8);
  int meta_blocks = (B + 255) / 256;
  asg_prep_kernel<<<frame_blocks + meta_blocks, 256, 0, stream>>>(p, frame_blocks);
  W2L_LAUNCH_CHECK("asg_prep_kernel");

  const int n_fac = (terms & W2L_TERM_FAC) ? B : 0;
  const int n_fcc = (terms & W2L_TERM_FCC) ? B : 0;
  if (p.need_grad) {
    if (smem > 48 * 1024)
      W2L_CUDA_CHECK(cudaFuncSetAttribute(asg_chains_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    asg_chains_kernel<true><<<n_fac + n_fcc, kChainThreads, smem, stream>>>(p, n_fac);
    W2L_LAUNCH_CHECK("asg_chains_kernel<grad>");
    dim3 grid((T + kGradChunk * 4 - 1) / (kGradChunk * 4), B);
    asg_grad_kernel<<<grid, 128, 0, stream>>>(p);
    W2L_LAUNCH_CHECK("asg_grad_kernel");
  } else {
    if (smem > 48 * 1024)
      W2L_CUDA_CHECK(cudaFuncSetAttribute(asg_chains_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    asg_chains_kernel<false><<<n_fac + n_fcc, kChainThreads, smem, stream>>>(p, n_fac);
    W2L_LAUNCH_CHECK("asg_chains_kernel<fwd>");
    asg_loss_only_kernel<<<(B + 127) / 128, 128, 0, stream>>>(p);
    W2L_LAUNCH_CHECK("asg_loss_only_kernel");
  }
  return W2L_OK;
}
