// gemm_umma.cu — hand-written sm_100a GEMM (tcgen05 / TMEM / TMA) for the dense contractions of the acoustic
// model: TDS fully-connected layers, the output Linear, and the large-channel Conv1D layers of the conv_glu
// archs as zero-copy im2col views (reference: fl::Linear / fl::Conv2D -> af::matmul / cuDNN through cuBLAS,
// reached from recipes/slimIPL/src/Train.cpp:1470 (forward) and :1720 (backward)).
//
//   C[m][n] = act( sum_k A(m,k) * B(n,k) + bias[n] )        fp32 accumulation in TMEM
//
// Three operand kinds, one kernel template (the "precision" of BASELINE.json's configs):
//   W2L_GEMM_TF32   fp32 operands in HBM, kind::tf32 (10-bit mantissa products)  — cuDNN/cuBLAS default for fp32
//   W2L_GEMM_F32X3  fp32 operands in HBM, fp32-ACCURATE: every staged tile is split in shared memory into
//                   hi = tf32(x) and lo = tf32(x - hi) by the (otherwise idle) epilogue warps and the tensor core
//                   accumulates Al*Bh + Ah*Bl + Ah*Bh — products good to ~2^-21, i.e. SGEMM-grade (configs[1] "fp32")
//   W2L_GEMM_BF16   bf16 operands in HBM, kind::f16 — half the operand bytes through L2 -> SM, twice the MAC rate
//                   (configs[2]/[3] "bf16 convs / fp32 loss"; the reference's AMP switch, Train.cpp:211-219)
// C is fp32 or bf16 (c_bf16); bias/ReLU/dropout/mask/accumulate epilogue as before.
//
// Operand storage ("major"):
//   A K-major : A stored [M][K] (row stride lda)      A MN-major : A stored [K][M]
//   B K-major : B stored [N][K] (row stride ldb)      B MN-major : B stored [K][N]
// so one kernel family covers  forward  Y = X W^T            (A = X  K-major,  B = W  K-major)
//                              dgrad    dX = dY W            (A = dY K-major,  B = W  MN-major)
//                              wgrad    dW = dY^T X          (A = dY MN-major, B = X  MN-major)
// without any transposition pass.
//
// Structure (one CTA per 128 x BN output tile, 192 threads):
//   warp 4        TMA producer: cp.async.bulk.tensor 2D boxes (128 B inner extent, 128B swizzle) into a 3/4-stage
//                 shared-memory ring, mbarrier expect_tx
//   warp 5        TMEM allocation + single-thread tcgen05.mma.cta_group::1 issue (UMMA 128 x BN x 32 B of k, 4 per
//                 k block; 12 in F32X3 mode), tcgen05.commit onto the stage's empty barrier, final commit onto the
//                 accumulator-full barrier
//   warps 0..3    F32X3: hi/lo split of every stage (generic-proxy writes + fence.proxy.async + `ready` barrier);
//                 epilogue: tcgen05.ld 32x32b (thread = accumulator row), bias + ReLU + dropout, transposition through
//                 shared memory, coalesced global stores
// Shared-memory operand layouts are the canonical UMMA ones (cute/atom/mma_traits_sm100.hpp):
//   K-major  SW128: rows of 128 B (32 fp32 / 64 bf16 along k), 8-row groups 1024 B apart (SBO), LBO unused
//   MN-major 32-bit: SW128_BASE32B (the only MN-major layout valid for 32-bit operands; TMA swizzle 128B_ATOM_32B):
//                   boxes of [32 k-rows][128 B = 32 elements along m/n]; LBO = box size (4096 B), SBO = 512 B (4 k-rows);
//                   one UMMA (k = 8) advances the start by 1024 B
//   MN-major 16-bit: SW128: boxes of [64 k-rows][128 B = 64 elements along m/n]; LBO = box size (8192 B),
//                   SBO = 1024 B (8 k-rows); one UMMA (k = 16) advances the start by 2048 B
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace w2l {
namespace {
using namespace umma;

constexpr int BM = 128;                     // tile rows
constexpr int kRowBytes = 128;              // one swizzle row of k: 32 fp32 or 64 bf16
constexpr int kTileBytes = BM * kRowBytes;  // 16 KB of A per stage
constexpr int kGemmThreads = 192;
enum { kTf32 = W2L_GEMM_TF32, kF32x3 = W2L_GEMM_F32X3, kBf16 = W2L_GEMM_BF16 };

// The tile width BN is a template parameter (128 / 160 / 224 / 256).  The kernel is fed from L2 (~42 B/clk per SM is
// the chip-wide L2 throughput cap), so tensor-pipe time per k block scales with the operand bytes (128 + BN) * 128 B
// while the work scales with 128 * BN: wider tiles raise MAC/byte, and the host picks the BN that minimises
// waves x bytes for each shape (e.g. 160 divides 800/1120/1440 exactly).  BN <= 160: 3 stages, 2 CTAs per SM (one
// tile's epilogue overlaps the other's main loop); BN > 160: 4 stages, 1 CTA per SM.  F32X3 doubles every stage
// (hi + lo tiles): BN = 128, 3 stages, 1 CTA per SM.
__host__ __device__ constexpr int stages_for(int mode, int bn) { return mode == kF32x3 ? 3 : (bn <= 160 ? 3 : 4); }
__host__ __device__ constexpr int tmem_cols_for(int bn) { return bn <= 128 ? 128 : 256; }
__host__ __device__ constexpr size_t stage_bytes(int mode, int bn) { return (size_t)(mode == kF32x3 ? 2 : 1) * (kTileBytes + bn * kRowBytes); }
__host__ __device__ constexpr size_t smem_for(int mode, int bn) { return stages_for(mode, bn) * stage_bytes(mode, bn) + 128 + 1024 + 1024; }  // ring + barriers + bias row + alignment slack

// instruction descriptor: D = f32; A = B = tf32 (format 2, kind::tf32) or bf16 (format 1, kind::f16); M x N; majors
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | ((bf16 ? 1u : 2u) << 7) | ((bf16 ? 1u : 2u) << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <bool kIsBf16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  if (kIsBf16) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

struct GemmParams {
  int M, N, K, ldc, act;
  void* C;
  const float* bias;
  // epilogue extensions: forward dropout, backward mask read from a stored activation, C += acc
  int accumulate, aux_mode, ld_aux;  // aux_mode 0: none, 1: (aux > 0) * aux_scale, 2: (aux != 0) * aux_scale
  const void* aux;
  float aux_scale, drop_p;
  unsigned long long seed;
  int k_splits;  // > 1: blockIdx.z owns a slice of the k blocks and the epilogue adds atomically (few tiles, long K: wgrad)
  int c_bf16, aux_bf16;
  int trust_trunc;  // F32X3 experiment: rely on the tensor core ignoring the 13 low mantissa bits (hi is not written back)
};

// ---- epilogue of one 128 x BN accumulator tile (warps 0..3; TMEM lanes 32*warp .. +31) -------------------------------
// tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns per chunk).  Bias, ReLU and the dropout mask
// (one Philox block per 4 consecutive columns) are applied in that layout; the chunk is then transposed through shared
// memory (tbuf: 4 x 32 x 33 floats, conflict-free both ways) so that every global access of the rest — mask read, C read
// for accumulation, store / red — is one row x 32 consecutive columns per warp instruction: 128 B coalesced.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tmem_acc, float* tbuf_all, const float* sbias, int m0, int n0,
                                              int warp, int lane, bool have_work) {
  float* tbuf = tbuf_all + warp * (32 * 36);
  const uint32_t tb_sa = smem_u32(tbuf), sbias_sa = smem_u32(sbias);
  // vector fast path (interior chunks of 16-byte aligned rows): the 32 x 32 chunk goes through shared memory as float4 and
  // every global access is 128 bits — 8 LDS.128 + 8 STG.128 per chunk instead of 32 + 32 scalar ones with per-row address
  // arithmetic and bounds checks (the scalar epilogue was the persistent kernel's bottleneck: 1800 instructions per chunk)
  const bool vec_tile = p.k_splits == 1 && (p.ldc % 4) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 &&
                        (p.aux_mode == 0 || ((p.ld_aux % 4) == 0 && (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0));
  const int row_own = m0 + warp * 32 + lane;
  const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f;
  const bool rd_aux = p.aux_mode != 0, rd_c = p.accumulate && p.k_splits == 1;
  float* Cf = static_cast<float*>(p.C);
  __nv_bfloat16* Ch = static_cast<__nv_bfloat16*>(p.C);
  if (have_work) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= p.N) break;  // warp-uniform
      const int col = nb + lane;
      const bool col_ok = col < p.N;
      const int rows_here = min(32, p.M - (m0 + warp * 32));  // warp-uniform; may be <= 0
      uint32_t v[32];
      const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float o[32];
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 b4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(sbias_sa + (uint32_t)(c * 32 + j) * 4));
          o[j] = __uint_as_float(v[j]) + b4.x;
          o[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
          o[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
          o[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
      }
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
      }
      if (p.drop_p > 0.f) {
        // Dropout mask: the backward pass reads it back from the stored activation (nothing is regenerated), so the generator
        // only has to be a good hash of (seed, element index): one 32-bit murmur3-style finaliser per PAIR of columns, 16 bits
        // per element, keep iff bits >= p * 65536.  (Philox4x32-10 per 4 columns was 480 of the ~600 instructions of a chunk.)
        const uint32_t thresh = (uint32_t)(p.drop_p * 65536.0f);
        const uint32_t s0 = (uint32_t)p.seed, s1 = (uint32_t)(p.seed >> 32);
        const unsigned long long base_idx = ((unsigned long long)row_own * (unsigned long long)p.N + (unsigned long long)nb) >> 1;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const unsigned long long idx = base_idx + (j >> 1);
          uint32_t h = ((uint32_t)idx ^ s0) * 0x9E3779B1u + ((uint32_t)(idx >> 32) ^ s1);
          h ^= h >> 15;
          h *= 0x85EBCA77u;
          h ^= h >> 13;
          h *= 0xC2B2AE3Du;
          h ^= h >> 16;
          o[j] *= (h & 0xffffu) >= thresh ? inv_keep : 0.f;
          o[j + 1] *= (h >> 16) >= thresh ? inv_keep : 0.f;
        }
      }
      __syncwarp();  // the previous chunk's transposed reads are done
      if (vec_tile && rows_here >= 32 && nb + 32 <= p.N) {  // warp-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tb_sa + (uint32_t)(lane * 36 + 4 * j) * 4), "f"(o[4 * j]), "f"(o[4 * j + 1]),
                       "f"(o[4 * j + 2]), "f"(o[4 * j + 3])
                       : "memory");
        __syncwarp();
        const int rsub = lane >> 3, c4 = lane & 7;  // this lane: rows rsub, rsub + 4, ..., columns 4 c4 .. 4 c4 + 3 of the chunk
        const size_t row0 = (size_t)(m0 + warp * 32 + rsub);
        const int colv = nb + 4 * c4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(tb_sa + (uint32_t)((rsub + 4 * i) * 36 + 4 * c4) * 4));
          const size_t row = row0 + 4 * i;
          if (rd_aux) {
            float4 m;
            if (p.aux_bf16) {
              const uint2 mb = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p.aux) + row * p.ld_aux + colv);
              const float2 m01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mb.x)), m23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mb.y));
              m = make_float4(m01.x, m01.y, m23.x, m23.y);
            } else {
              m = *reinterpret_cast<const float4*>(static_cast<const float*>(p.aux) + row * p.ld_aux + colv);
            }
            const float sc = p.aux_scale;
            if (p.aux_mode == 1) {
              x.x *= m.x > 0.f ? sc : 0.f;
              x.y *= m.y > 0.f ? sc : 0.f;
              x.z *= m.z > 0.f ? sc : 0.f;
              x.w *= m.w > 0.f ? sc : 0.f;
            } else {
              x.x *= m.x != 0.f ? sc : 0.f;
              x.y *= m.y != 0.f ? sc : 0.f;
              x.z *= m.z != 0.f ? sc : 0.f;
              x.w *= m.w != 0.f ? sc : 0.f;
            }
          }
          if (p.c_bf16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
            *reinterpret_cast<uint2*>(Ch + row * p.ldc + colv) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
          } else {
            float4* dst = reinterpret_cast<float4*>(Cf + row * p.ldc + colv);
            if (rd_c) {
              const float4 cv = *dst;
              x.x += cv.x;
              x.y += cv.y;
              x.z += cv.z;
              x.w += cv.w;
            }
            *dst = x;
          }
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = o[j];
      // the chunk's mask (or, without a mask, its C values to accumulate onto) in the coalesced layout (lane = column,
      // one row per instruction); the lines were prefetched into L2 during the main loop
      float pre[32];
      if (rd_aux) {
        if (p.aux_bf16) {
          const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(p.aux) + (size_t)(m0 + warp * 32) * p.ld_aux + col;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) pre[rr] = (rr < rows_here && col_ok) ? __bfloat162float(src[(size_t)rr * p.ld_aux]) : 0.f;
        } else {
          const float* src = static_cast<const float*>(p.aux) + (size_t)(m0 + warp * 32) * p.ld_aux + col;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) pre[rr] = (rr < rows_here && col_ok) ? src[(size_t)rr * p.ld_aux] : 0.f;
        }
      } else if (rd_c) {
        const float* src = Cf + (size_t)(m0 + warp * 32) * p.ldc + col;
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) pre[rr] = (rr < rows_here && col_ok) ? src[(size_t)rr * p.ldc] : 0.f;
      }
      __syncwarp();
      if (p.c_bf16 && rd_aux) {  // bf16 C behind a mask: column layout (the mask was fetched in it), 2-byte stores
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          if (rr < rows_here && col_ok) {
            float x = tbuf[rr * 33 + lane];
            x *= (p.aux_mode == 1 ? pre[rr] > 0.f : pre[rr] != 0.f) ? p.aux_scale : 0.f;
            Ch[(size_t)(m0 + warp * 32 + rr) * p.ldc + col] = __float2bfloat16_rn(x);
          }
        }
      } else if (p.c_bf16) {  // bf16 C (no accumulate / split-K: checked by the host): a lane stores two adjacent columns
        const int sub = lane >> 4, cp = (lane & 15) * 2;
        const bool pair_ok = nb + cp + 1 < p.N, one_ok = nb + cp < p.N;
#pragma unroll 8
        for (int r2 = 0; r2 < 16; ++r2) {
          const int rr = 2 * r2 + sub;
          if (rr < rows_here && one_ok) {
            const float x0 = tbuf[rr * 33 + cp], x1 = tbuf[rr * 33 + cp + 1];
            __nv_bfloat16* dst = Ch + (size_t)(m0 + warp * 32 + rr) * p.ldc + nb + cp;
            if (pair_ok && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
              *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(x0, x1);
            } else {
              dst[0] = __float2bfloat16_rn(x0);
              if (pair_ok) dst[1] = __float2bfloat16_rn(x1);
            }
          }
        }
      } else if (!rd_aux && !rd_c) {  // no global reads: stream the rows out
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          if (rr < rows_here && col_ok) {
            float* dst = Cf + (size_t)(m0 + warp * 32 + rr) * p.ldc + col;
            const float x = tbuf[rr * 33 + lane];
            if (p.k_splits > 1)
              atomicAdd(dst, x);  // split-K: C was zeroed (or holds the value to accumulate onto) by the host wrapper
            else
              *dst = x;
          }
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          if (rr < rows_here && col_ok) {
            float* dst = Cf + (size_t)(m0 + warp * 32 + rr) * p.ldc + col;
            float x = tbuf[rr * 33 + lane];
            if (rd_aux) {
              x *= (p.aux_mode == 1 ? pre[rr] > 0.f : pre[rr] != 0.f) ? p.aux_scale : 0.f;
              if (rd_c) x += *dst;  // mask and accumulation together (not on the TDS path): C is read here
            } else {
              x += pre[rr];
            }
            if (p.k_splits > 1)
              atomicAdd(dst, x);
            else
              *dst = x;
          }
        }
      }
    }
  }
}

template <int kMode, bool kAMn, bool kBMn, int BN>
__global__ void __launch_bounds__(kGemmThreads, (kMode == kF32x3 || BN > 160) ? 1 : 2)
gemm_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr bool kIsBf16 = kMode == kBf16, kSplit = kMode == kF32x3;
  constexpr int ES = kIsBf16 ? 2 : 4;                 // operand element size
  constexpr int BKE = kRowBytes / ES;                 // elements of k per k block: 32 fp32 / 64 bf16
  constexpr int kMnAtom = kRowBytes / ES;             // elements of m/n per MN-major box: 32 / 64
  constexpr int kMnBoxBytes = BKE * kRowBytes;        // one MN-major box [BKE k-rows][128 B]: 4096 / 8192
  constexpr int kMnStep = kIsBf16 ? 2048 : 1024;      // MN-major start advance per UMMA (k = 16 / 8 rows)
  constexpr int kMnSbo = kIsBf16 ? 1024 : 512;        // k-row group stride (8 / 4 rows)
  constexpr int kMnLayout = kIsBf16 ? 2 : 1;          // SWIZZLE_128B / SWIZZLE_128B_BASE32B
  constexpr int kStages = stages_for(kMode, BN), kTileBytesB = BN * kRowBytes, kTmemCols = tmem_cols_for(BN);
  constexpr int kOperandBytes = kTileBytes + kTileBytesB;
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + kStages * kTileBytes;
  unsigned char* smem_lo = smem + kStages * kOperandBytes;  // F32X3: [stage][A lo | B lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes(kMode, BN));
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* ready = bars + 2 * kStages;  // F32X3 only
  uint64_t* acc_full = bars + 3 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int total_kb = (p.K + BKE - 1) / BKE;
  const int kb_per = (total_kb + p.k_splits - 1) / p.k_splits;
  const int kb_begin = blockIdx.z * kb_per;
  const int num_kb = max(0, min(total_kb, kb_begin + kb_per) - kb_begin);

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&ready[s], 128);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], kOperandBytes);
        unsigned char* sa = smem_a + s * kTileBytes;
        unsigned char* sb = smem_b + s * kTileBytesB;
        const int k0 = (kb_begin + kb) * BKE;
        if (!kAMn) {
          tma_load_2d(&map_a, &full[s], sa, k0, m0);  // box {128 B of k, 128 rows}
        } else {
#pragma unroll
          for (int j = 0; j < BM / kMnAtom; ++j) tma_load_2d(&map_a, &full[s], sa + j * kMnBoxBytes, m0 + kMnAtom * j, k0);  // box {128 B of m, BKE k}
        }
        if (!kBMn) {
          tma_load_2d(&map_b, &full[s], sb, k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / kMnAtom; ++j) tma_load_2d(&map_b, &full[s], sb + j * kMnBoxBytes, n0 + kMnAtom * j, k0);
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer (one elected thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kIsBf16, BM, BN, kAMn, kBMn);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(kSplit ? &ready[s] : &full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem_a + s * kTileBytes);
        const uint32_t sb = smem_u32(smem_b + s * kTileBytesB);
        const uint32_t la = smem_u32(smem_lo + s * kOperandBytes), lb = la + kTileBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 UMMAs of 32 B of k per 128 B swizzle row
          // K-major: +32 B per step inside the 128 B swizzle row; MN-major: + one k group of boxes' rows
          auto adesc = [&](uint32_t base) { return kAMn ? make_smem_desc(base + k * kMnStep, kMnBoxBytes, kMnSbo, kMnLayout) : make_smem_desc(base + k * 32, 16, 1024, 2); };
          auto bdesc = [&](uint32_t base) { return kBMn ? make_smem_desc(base + k * kMnStep, kMnBoxBytes, kMnSbo, kMnLayout) : make_smem_desc(base + k * 32, 16, 1024, 2); };
          if (kSplit) {
            umma<false>(tmem_base, adesc(la), bdesc(sb), idesc, (kb | k) != 0);  // Al * Bh
            umma<false>(tmem_base, adesc(sa), bdesc(lb), idesc, 1);              // Ah * Bl
            umma<false>(tmem_base, adesc(sa), bdesc(sb), idesc, 1);              // Ah * Bh
          } else {
            umma<kIsBf16>(tmem_base, adesc(sa), bdesc(sb), idesc, (kb | k) != 0);
          }
        }
        umma_commit(&empty[s]);  // frees the stage when the MMAs above have read it
      }
      umma_commit(acc_full);  // with num_kb == 0 nothing is pending: the barrier completes immediately
    }
  } else {
    // ===== warps 0..3: (F32X3: operand split,) epilogue on TMEM lanes 32*warp .. +31 =====
    // tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns per chunk).  Bias, ReLU and the
    // dropout mask (one Philox block per 4 consecutive columns) are applied in that layout; the chunk is then
    // transposed through shared memory (the idle operand ring; 33-float pitch, conflict-free both ways) so that
    // every global access of the rest — mask read, C read for accumulation, store / red — is one row x 32
    // consecutive columns per warp instruction: 128 B coalesced instead of 32 sectors.
    float* sbias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 128);  // [BN]
    if (p.bias != nullptr) {
      for (int j = threadIdx.x; j < BN; j += 128) sbias[j] = n0 + j < p.N ? __ldg(p.bias + n0 + j) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // while the main loop runs: pull the tile's mask / C lines into L2 so the epilogue's reads are L2 hits
    if (p.aux_mode != 0 || (p.accumulate && p.k_splits == 1)) {
      const int lines = (BN * 4 + 127) / 128;  // 128-byte lines per tile row (fp32 columns; bf16 masks touch half of them)
      for (int i = threadIdx.x; i < BM * lines; i += 128) {
        const int r = m0 + i / lines, cc = n0 + (i % lines) * 32;
        if (r < p.M && cc < p.N) {
          if (p.aux_mode != 0) {
            if (p.aux_bf16) {
              if (((i % lines) & 1) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const __nv_bfloat16*>(p.aux) + (size_t)r * p.ld_aux + cc));
            } else {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const float*>(p.aux) + (size_t)r * p.ld_aux + cc));
            }
          }
          if (p.accumulate && p.k_splits == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const float*>(p.C) + (size_t)r * p.ldc + cc));
        }
      }
    }
    if (kSplit) {
      // fp32-accurate mode: x = hi + lo with hi = tf32(x) (written back in place) and lo = tf32(x - hi) (second tile
      // of the stage); elementwise, so the swizzled / MN-major tile layouts are irrelevant here
      constexpr int kVecA = kTileBytes / 16 / 128, kVecB = kTileBytesB / 16 / 128;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&full[s], ph);
        uint4* ta = reinterpret_cast<uint4*>(smem_a + s * kTileBytes);
        uint4* tb = reinterpret_cast<uint4*>(smem_b + s * kTileBytesB);
        uint4* la = reinterpret_cast<uint4*>(smem_lo + s * kOperandBytes);
        uint4* lb = reinterpret_cast<uint4*>(smem_lo + s * kOperandBytes + kTileBytes);
        auto split = [](uint4* hi, uint4* lo, int i) {
          const uint4 v = hi[i];
          uint4 h, l;
          h.x = rna_tf32(__uint_as_float(v.x));
          h.y = rna_tf32(__uint_as_float(v.y));
          h.z = rna_tf32(__uint_as_float(v.z));
          h.w = rna_tf32(__uint_as_float(v.w));
          l.x = rna_tf32(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = rna_tf32(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = rna_tf32(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = rna_tf32(__uint_as_float(v.w) - __uint_as_float(h.w));
          hi[i] = h;
          lo[i] = l;
        };
#pragma unroll
        for (int i = 0; i < kVecA; ++i) split(ta, la, i * 128 + threadIdx.x);
#pragma unroll
        for (int i = 0; i < kVecB; ++i) split(tb, lb, i * 128 + threadIdx.x);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        mbar_arrive(&ready[s]);
      }
    }
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue_tile<BN>(p, tmem_base, reinterpret_cast<float*>(smem_a), sbias, m0, n0, warp, lane, num_kb > 0 || p.k_splits == 1);  // the ring is idle by now
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ================================================================================================
// Persistent variant (the default): ONE CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, so the
// CTAs running together share their A rows in L2).  What it removes from the one-tile-per-CTA kernel above:
//   * the per-tile prologue (barrier init, TMEM allocation, descriptor fetch, cold first loads): paid once per SM;
//   * the exposed epilogue: TWO accumulators live in TMEM (2 x BN columns); while warps 0..3 drain accumulator i
//     (tcgen05.ld -> bias / ReLU / dropout -> transposition -> coalesced stores), the MMA thread already accumulates the
//     next tile into accumulator i ^ 1 and the TMA warp keeps the operand ring full across the tile boundary.
// Roles: warps 0..3 epilogue, warp 4 TMA producer, warp 5 MMA issuer + TMEM owner, (F32X3 only) warps 6..13 operand split.
// Barriers: full / empty (/ ready) per ring stage with a k-block counter that runs across tiles; acc_full / acc_empty per
// accumulator with the tile counter's parity.
// ================================================================================================
__host__ __device__ constexpr int pstages_for(int mode, int bn) { return mode == kF32x3 ? 3 : (bn <= 160 ? 5 : 4); }
constexpr int kSplitThreads = 256;  // F32X3: split warps of the comparison variant (the default launches 384 threads = twelve warps, see launch_persistent_bn)
__host__ __device__ constexpr int pthreads_for(int mode) { return mode == kF32x3 ? 192 + kSplitThreads : 192; }
__host__ __device__ constexpr int acc_stride_for(int bn) { return bn <= 128 ? 128 : 256; }
constexpr int kTbufBytes = 4 * 32 * 36 * 4;  // per warp [32][36] floats: 128-bit conflict-free both ways (the scalar path uses a 33 pitch)
__host__ __device__ constexpr size_t psmem_for(int mode, int bn) {
  return pstages_for(mode, bn) * stage_bytes(mode, bn) + kTbufBytes + 256 + 2 * 1024 + 1024;  // ring + transposition + barriers + 2 bias rows + slack
}

template <int kMode, bool kAMn, bool kBMn, int BN, int kST = kSplitThreads>
__global__ void __launch_bounds__(kMode == kF32x3 ? 192 + kST : 192, 1)
gemm_umma_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmParams p, int tiles_m,
                            int tiles_n) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr bool kIsBf16 = kMode == kBf16, kSplit = kMode == kF32x3;
  constexpr int ES = kIsBf16 ? 2 : 4;
  constexpr int BKE = kRowBytes / ES;
  constexpr int kMnAtom = kRowBytes / ES;
  constexpr int kMnBoxBytes = BKE * kRowBytes;
  constexpr int kMnStep = kIsBf16 ? 2048 : 1024;
  constexpr int kMnSbo = kIsBf16 ? 1024 : 512;
  constexpr int kMnLayout = kIsBf16 ? 2 : 1;
  constexpr int kStages = pstages_for(kMode, BN), kTileBytesB = BN * kRowBytes;
  constexpr int kOperandBytes = kTileBytes + kTileBytesB;
  constexpr int kAccStride = acc_stride_for(BN), kTmemCols = 2 * kAccStride;
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + kStages * kTileBytes;
  unsigned char* smem_lo = smem + kStages * kOperandBytes;  // F32X3: [stage][A lo | B lo]
  unsigned char* tail = smem + kStages * stage_bytes(kMode, BN);
  float* tbuf = reinterpret_cast<float*>(tail);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + kTbufBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* ready = bars + 2 * kStages;
  uint64_t* acc_full = bars + 3 * kStages;       // [2]
  uint64_t* acc_empty = bars + 3 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 4);
  float* sbias_all = reinterpret_cast<float*>(tail + kTbufBytes + 256);  // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = (p.K + BKE - 1) / BKE;
  const int kb_per = (total_kb + p.k_splits - 1) / p.k_splits;
  const int tiles_mn = tiles_m * tiles_n, total_tiles = tiles_mn * p.k_splits;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&ready[s], kST);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // every role walks the same tile sequence with the same k-block counts, so the ring / accumulator phases agree
  auto tile_coords = [&](int t, int& m0, int& n0, int& kb_begin, int& num_kb) {
    const int z = t / tiles_mn, r = t - z * tiles_mn;
    m0 = (r / tiles_n) * BM;
    n0 = (r % tiles_n) * BN;
    kb_begin = z * kb_per;
    num_kb = max(0, min(total_kb, kb_begin + kb_per) - kb_begin);
  };

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0, kb_begin, num_kb;
        tile_coords(t, m0, n0, kb_begin, num_kb);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], kOperandBytes);
          unsigned char* sa = smem_a + s * kTileBytes;
          unsigned char* sb = smem_b + s * kTileBytesB;
          const int k0 = (kb_begin + kb) * BKE;
          if (!kAMn) {
            tma_load_2d(&map_a, &full[s], sa, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / kMnAtom; ++j) tma_load_2d(&map_a, &full[s], sa + j * kMnBoxBytes, m0 + kMnAtom * j, k0);
          }
          if (!kBMn) {
            tma_load_2d(&map_b, &full[s], sb, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / kMnAtom; ++j) tma_load_2d(&map_b, &full[s], sb + j * kMnBoxBytes, n0 + kMnAtom * j, k0);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kIsBf16, BM, BN, kAMn, kBMn);
      uint32_t it = 0, lt = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
        int m0, n0, kb_begin, num_kb;
        tile_coords(t, m0, n0, kb_begin, num_kb);
        const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
        mbar_wait(&acc_empty[acc], aph ^ 1);  // the epilogue has drained this accumulator (first use: passes)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_acc = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(kSplit ? &ready[s] : &full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem_a + s * kTileBytes);
          const uint32_t sb = smem_u32(smem_b + s * kTileBytesB);
          const uint32_t la = smem_u32(smem_lo + s * kOperandBytes), lb = la + kTileBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            auto adesc = [&](uint32_t base) { return kAMn ? make_smem_desc(base + k * kMnStep, kMnBoxBytes, kMnSbo, kMnLayout) : make_smem_desc(base + k * 32, 16, 1024, 2); };
            auto bdesc = [&](uint32_t base) { return kBMn ? make_smem_desc(base + k * kMnStep, kMnBoxBytes, kMnSbo, kMnLayout) : make_smem_desc(base + k * 32, 16, 1024, 2); };
            if (kSplit) {
              umma<false>(tmem_acc, adesc(la), bdesc(sb), idesc, (kb | k) != 0);  // Al * Bh
              umma<false>(tmem_acc, adesc(sa), bdesc(lb), idesc, 1);              // Ah * Bl
              umma<false>(tmem_acc, adesc(sa), bdesc(sb), idesc, 1);              // Ah * Bh
            } else {
              umma<kIsBf16>(tmem_acc, adesc(sa), bdesc(sb), idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[acc]);  // all MMAs of this tile done -> the epilogue may read the accumulator
      }
    }
  } else if (warp < 4) {
    // ===== epilogue =====
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
      int m0, n0, kb_begin, num_kb;
      tile_coords(t, m0, n0, kb_begin, num_kb);
      const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
      float* sbias = sbias_all + acc * BN;
      if (p.bias != nullptr)
        for (int j = threadIdx.x; j < BN; j += 128) sbias[j] = n0 + j < p.N ? __ldg(p.bias + n0 + j) : 0.f;
      // pull the tile's mask / C lines into L2 while its main loop runs
      if (p.aux_mode != 0 || (p.accumulate && p.k_splits == 1)) {
        const int lines = (BN * 4 + 127) / 128;
        for (int i = threadIdx.x; i < BM * lines; i += 128) {
          const int r = m0 + i / lines, cc = n0 + (i % lines) * 32;
          if (r < p.M && cc < p.N) {
            if (p.aux_mode != 0) {
              if (p.aux_bf16) {
                if (((i % lines) & 1) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const __nv_bfloat16*>(p.aux) + (size_t)r * p.ld_aux + cc));
              } else {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const float*>(p.aux) + (size_t)r * p.ld_aux + cc));
              }
            }
            if (p.accumulate && p.k_splits == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const float*>(p.C) + (size_t)r * p.ldc + cc));
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // bias row staged; every warp is past the previous tile's reads of the other row
      mbar_wait(&acc_full[acc], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      epilogue_tile<BN>(p, tmem_base + acc * kAccStride, tbuf, sbias, m0, n0, warp, lane, num_kb > 0 || p.k_splits == 1);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc_empty[acc]);  // this thread's TMEM reads of the accumulator are complete
    }
  } else if (kSplit) {
    // ===== F32X3 operand split (warps 6..13): x = hi + lo, hi = tf32(x) written back in place, lo = tf32(x - hi) in the
    // second tile of the stage; round-to-nearest by integer arithmetic (add half an ulp of the 13 dropped bits, clear them):
    // full-rate ALU ops instead of quarter-rate cvt.rna
    const int tid = threadIdx.x - 192;
    constexpr int kNVecA = kTileBytes / 16, kNVecB = kTileBytesB / 16;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int m0, n0, kb_begin, num_kb;
      tile_coords(t, m0, n0, kb_begin, num_kb);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full[s], ph);
        uint4* ta = reinterpret_cast<uint4*>(smem_a + s * kTileBytes);
        uint4* tb = reinterpret_cast<uint4*>(smem_b + s * kTileBytesB);
        uint4* la = reinterpret_cast<uint4*>(smem_lo + s * kOperandBytes);
        uint4* lb = reinterpret_cast<uint4*>(smem_lo + s * kOperandBytes + kTileBytes);
        auto rn = [](uint32_t x) { return (x + 0x1000u) & 0xffffe000u; };
        auto split = [&](uint4* hi, uint4* lo, int i) {
          const uint4 v = hi[i];
          uint4 h, l;
          if (p.trust_trunc) {  // hi = the raw value as the hardware truncates it; only lo is written
            l.x = rn(__float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xffffe000u)));
            l.y = rn(__float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xffffe000u)));
            l.z = rn(__float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xffffe000u)));
            l.w = rn(__float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xffffe000u)));
            lo[i] = l;
            return;
          }
          h.x = rn(v.x);
          h.y = rn(v.y);
          h.z = rn(v.z);
          h.w = rn(v.w);
          l.x = rn(__float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)));
          l.y = rn(__float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)));
          l.z = rn(__float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)));
          l.w = rn(__float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)));
          hi[i] = h;
          lo[i] = l;
        };
#pragma unroll
        for (int i = tid; i < kNVecA; i += kST) split(ta, la, i);
#pragma unroll
        for (int i = tid; i < kNVecB; i += kST) split(tb, lb, i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&ready[s]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2D map over a row-major matrix [rows][cols] (row stride ld elements); box {128 B of columns, box_rows}
int make_map(CUtensorMap* map, int mode, const void* ptr, long long rows, long long cols, long long ld, int box_rows, bool mn_major) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(W2L_ERR_CUDA, "gemm: cuTensorMapEncodeTiled entry point not found");
  const int es = mode == kBf16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)(kRowBytes / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  // TF32: rounding on load; F32X3: the raw fp32 bits (the split happens in shared memory)
  const CUtensorMapDataType dt = mode == kBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                               : (mode == kF32x3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32);
  const CUtensorMapSwizzle sw = (mn_major && mode != kBf16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(W2L_ERR_CUDA, "gemm: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return W2L_OK;
}

const char* kernel_name(int mode) { return mode == kBf16 ? "gemm_umma_kernel<bf16>" : (mode == kF32x3 ? "gemm_umma_kernel<f32x3>" : "gemm_umma_kernel<tf32>"); }

template <int kMode, bool kAMn, bool kBMn, int BN>
int launch_bn(cudaStream_t stream, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p) {
  constexpr size_t smem = smem_for(kMode, BN);
  static_assert(smem <= 227 * 1024, "gemm: shared-memory budget");
  static bool configured = false;
  if (!configured) {
    W2L_CUDA_CHECK(cudaFuncSetAttribute(gemm_umma_kernel<kMode, kAMn, kBMn, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.k_splits);
  profile_kind(1);
  profile_start(stream);
  gemm_umma_kernel<kMode, kAMn, kBMn, BN><<<grid, kGemmThreads, smem, stream>>>(ma, mb, p);
  profile_stop(stream);
  W2L_LAUNCH_CHECK(kernel_name(kMode));
  return W2L_OK;
}
thread_local int g_trust_trunc = 1;  // measured on B200: kind::tf32 ignores the 13 low mantissa bits, same accuracy as an explicit hi tile, 10 % faster
thread_local int g_variant = 1;  // 1: persistent kernel (default), 0: one tile per CTA (w2l_gemm_set_variant; tests compare the two)
int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}
template <int kMode, bool kAMn, bool kBMn, int BN>
int launch_persistent_bn(cudaStream_t stream, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p) {
  constexpr size_t smem = psmem_for(kMode, BN);
  static_assert(smem <= 227 * 1024, "gemm (persistent): shared-memory budget");
  static bool configured = false;
  if (!configured) {
    W2L_CUDA_CHECK(cudaFuncSetAttribute(gemm_umma_persistent_kernel<kMode, kAMn, kBMn, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int total = tiles_m * tiles_n * p.k_splits;
  profile_kind(1);
  profile_start(stream);
  if constexpr (kMode == kF32x3) {
    // twelve split warps by default (measured: GEMM time per TDS+CTC step 10.15 ms with four, 8.56 with eight, 8.13 with
    // twelve); W2L_F32X3_SPLIT_THREADS=256 selects the eight-warp variant for comparison
    static const int st = [] {
      const char* e = getenv("W2L_F32X3_SPLIT_THREADS");
      return (e && atoi(e) == 256) ? 256 : 384;
    }();
    if (st == 384) {
      static bool configured384 = false;
      if (!configured384) {
        W2L_CUDA_CHECK(cudaFuncSetAttribute(gemm_umma_persistent_kernel<kMode, kAMn, kBMn, BN, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured384 = true;
      }
      gemm_umma_persistent_kernel<kMode, kAMn, kBMn, BN, 384><<<std::min(total, sm_count()), 192 + 384, smem, stream>>>(ma, mb, p, tiles_m, tiles_n);
    } else {
      gemm_umma_persistent_kernel<kMode, kAMn, kBMn, BN><<<std::min(total, sm_count()), pthreads_for(kMode), smem, stream>>>(ma, mb, p, tiles_m, tiles_n);
    }
  } else {
    gemm_umma_persistent_kernel<kMode, kAMn, kBMn, BN><<<std::min(total, sm_count()), pthreads_for(kMode), smem, stream>>>(ma, mb, p, tiles_m, tiles_n);
  }
  profile_stop(stream);
  W2L_LAUNCH_CHECK(kernel_name(kMode));
  return W2L_OK;
}
template <int kMode, bool kAMn, bool kBMn>
int launch_mode(cudaStream_t stream, int bn, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p) {
  if (g_variant == 1) {
    if constexpr (kMode == kF32x3) {
      return launch_persistent_bn<kMode, kAMn, kBMn, 128>(stream, ma, mb, p);
    } else if constexpr (kMode == kBf16 && kBMn) {
      return bn <= 128 ? launch_persistent_bn<kMode, kAMn, kBMn, 128>(stream, ma, mb, p) : launch_persistent_bn<kMode, kAMn, kBMn, 256>(stream, ma, mb, p);
    } else {
      switch (bn) {
        case 128: return launch_persistent_bn<kMode, kAMn, kBMn, 128>(stream, ma, mb, p);
        case 160: return launch_persistent_bn<kMode, kAMn, kBMn, 160>(stream, ma, mb, p);
        case 224: return launch_persistent_bn<kMode, kAMn, kBMn, 224>(stream, ma, mb, p);
        default: return launch_persistent_bn<kMode, kAMn, kBMn, 256>(stream, ma, mb, p);
      }
    }
  }
  if constexpr (kMode == kF32x3) {
    return launch_bn<kMode, kAMn, kBMn, 128>(stream, ma, mb, p);
  } else if constexpr (kMode == kBf16 && kBMn) {  // 64-wide MN-major boxes: BN a multiple of 64
    return bn <= 128 ? launch_bn<kMode, kAMn, kBMn, 128>(stream, ma, mb, p) : launch_bn<kMode, kAMn, kBMn, 256>(stream, ma, mb, p);
  } else {
    switch (bn) {
      case 128: return launch_bn<kMode, kAMn, kBMn, 128>(stream, ma, mb, p);
      case 160: return launch_bn<kMode, kAMn, kBMn, 160>(stream, ma, mb, p);
      case 224: return launch_bn<kMode, kAMn, kBMn, 224>(stream, ma, mb, p);
      default: return launch_bn<kMode, kAMn, kBMn, 256>(stream, ma, mb, p);
    }
  }
}
template <int kMode>
int launch(cudaStream_t stream, bool a_mn, bool b_mn, int bn, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p) {
  if (!a_mn && !b_mn) return launch_mode<kMode, false, false>(stream, bn, ma, mb, p);
  if (!a_mn && b_mn) return launch_mode<kMode, false, true>(stream, bn, ma, mb, p);
  if (a_mn && b_mn) return launch_mode<kMode, true, true>(stream, bn, ma, mb, p);
  return launch_mode<kMode, true, false>(stream, bn, ma, mb, p);
}

// split-K factor for a tile count: when the tiles alone under-fill the chip and K is long (weight gradients),
// slices of >= 4 k blocks; only for plain epilogues (the partial sums are added atomically)
int splits_for(int tiles, int total_kb, int slots_per_sm, bool plain) {
  const int slots = 148 * slots_per_sm;
  if (!plain || tiles >= 148 || total_kb < 16) return 1;
  return std::max(1, std::min(std::min(slots / tiles, total_kb / 4), 32));
}
// tile width: minimise (waves over 148 SMs) x (operand bytes per tile + exposed epilogue), see the note at the top
thread_local int g_force_bn = 0;  // w2l_gemm_set_tile: tests pin the tile width
int choose_bn(int mode, bool b_mn, int M, int N, int total_kb, bool plain, int* splits_out) {
  auto allowed = [&](int bn) {
    if (mode == kF32x3) return bn == 128;
    if (mode == kBf16 && b_mn) return bn == 128 || bn == 256;
    return true;
  };
  auto slots_of = [&](int bn) { return (g_variant == 1 || mode == kF32x3 || bn > 160) ? 1 : 2; };
  if (g_force_bn && allowed(g_force_bn)) {
    const int tiles = ((N + g_force_bn - 1) / g_force_bn) * ((M + BM - 1) / BM);
    *splits_out = splits_for(tiles, total_kb, slots_of(g_force_bn), plain);
    return g_force_bn;
  }
  const int cands[4] = {128, 160, 224, 256};
  double best = 1e300;
  int best_bn = 128;
  *splits_out = 1;
  for (int bn : cands) {
    if (!allowed(bn)) continue;
    const int tiles = ((N + bn - 1) / bn) * ((M + BM - 1) / BM);
    const int splits = splits_for(tiles, total_kb, slots_of(bn), plain);
    const int kb_local = (total_kb + splits - 1) / splits;
    const double waves = std::ceil((double)tiles * splits / 148.0);
    const double cost = waves * ((double)(128 + bn) * kb_local + (bn <= 160 ? 1.0 : 3.0) * bn);
    if (cost < best - 1e-9) {
      best = cost;
      best_bn = bn;
      *splits_out = splits;
    }
  }
  return best_bn;
}

int gemm_impl(void* stream_, int mode, int a_mn_major, int b_mn_major, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
              void* C, int ldc, int c_bf16, const float* bias, int act, int accumulate, const void* aux, int ld_aux, int aux_bf16,
              int aux_mode, float aux_scale, float dropout_p, unsigned long long seed, bool allow_overlap) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (mode != kTf32 && mode != kF32x3 && mode != kBf16) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: unknown operand kind");
  if (M <= 0 || N <= 0 || K <= 0) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: M, N, K must be positive");
  if (!A || !B || !C) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: null pointer");
  if (act < 0 || act > 1) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: act must be 0 (none) or 1 (relu)");
  if (aux_mode < 0 || aux_mode > 2 || (aux_mode != 0 && (!aux || ld_aux < N)))
    return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: bad aux mask arguments");
  if (dropout_p < 0.f || dropout_p >= 1.f) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: dropout_p must be in [0, 1)");
  if (dropout_p > 0.f && N % 4 != 0) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: dropout needs N % 4 == 0 (one Philox block per 4 columns)");
  const int row_align = mode == kBf16 ? 8 : 4;  // elements per 16 bytes
  if ((lda % row_align) || (ldb % row_align) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: operand rows must be 16-byte aligned (ld % 4 == 0 for fp32, ld % 8 == 0 for bf16)");
  if (ldc < N || (!allow_overlap && (lda < (a_mn_major ? M : K) || ldb < (b_mn_major ? N : K))))
    return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: leading dimension smaller than the row length");
  if (lda <= 0 || ldb <= 0) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: non-positive leading dimension");
  if (c_bf16 && accumulate) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: accumulation needs an fp32 C");
  const int bke = kRowBytes / (mode == kBf16 ? 2 : 4);
  const int total_kb = (K + bke - 1) / bke;
  const bool plain = act == 0 && aux_mode == 0 && dropout_p == 0.f && bias == nullptr && !c_bf16;
  int splits = 1;
  const int BN = choose_bn(mode, b_mn_major != 0, M, N, total_kb, plain, &splits);
  CUtensorMap ma, mb;
  int rc;
  if (!a_mn_major)
    rc = make_map(&ma, mode, A, M, K, lda, BM, false);   // [M][K]: box {128 B of k, 128 m}
  else
    rc = make_map(&ma, mode, A, K, M, lda, bke, true);   // [K][M]: box {128 B of m, bke k}
  if (rc) return rc;
  if (!b_mn_major)
    rc = make_map(&mb, mode, B, N, K, ldb, BN, false);
  else
    rc = make_map(&mb, mode, B, K, N, ldb, bke, true);
  if (rc) return rc;
  if (splits > 1 && !accumulate) W2L_CUDA_CHECK(cudaMemset2DAsync(C, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)N, (size_t)M, stream));
  GemmParams p{M, N, K, ldc, act, C, bias, accumulate, aux_mode, ld_aux, aux, aux_scale, dropout_p, seed, splits, c_bf16, aux_bf16, g_trust_trunc};
  if (mode == kBf16) return launch<kBf16>(stream, a_mn_major, b_mn_major, BN, ma, mb, p);
  if (mode == kF32x3) return launch<kF32x3>(stream, a_mn_major, b_mn_major, BN, ma, mb, p);
  return launch<kTf32>(stream, a_mn_major, b_mn_major, BN, ma, mb, p);
}

// fp32-operand entry points follow the thread's precision setting
int f32_kind() { return current_precision() == W2L_PRECISION_F32 ? kF32x3 : kTf32; }

__global__ void __launch_bounds__(256) cast_bf16_kernel(long long n4, const float4* __restrict__ x, uint2* __restrict__ y, long long n,
                                                        const float* __restrict__ xs, __nv_bfloat16* __restrict__ ys) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 v = x[i];
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    y[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
  if (i < n - 4 * n4) ys[4 * n4 + i] = __float2bfloat16_rn(xs[4 * n4 + i]);
}
// rows of `cols` floats (row stride ld_in) -> rows of cols_p bf16 (row stride cols_p), zero-padded columns
__global__ void __launch_bounds__(256) cast_bf16_rows_kernel(long long rows, int cols, int ld_in, int cols_p, const float* __restrict__ x,
                                                             __nv_bfloat16* __restrict__ y) {
  const long long total = rows * cols_p;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_p;
    const int c = (int)(i % cols_p);
    y[i] = c < cols ? __float2bfloat16_rn(x[r * ld_in + c]) : __float2bfloat16_rn(0.f);
  }
}

}  // namespace
}  // namespace w2l

using namespace w2l;

extern "C" int w2l_gemm_set_variant(int variant) {
  if (variant < 0 || variant > 3) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: variant must be 0 (one tile per CTA) or 1 (persistent); +2: F32X3 writes its hi tile explicitly");
  g_variant = variant & 1;
  g_trust_trunc = (variant >> 1) & 1 ? 0 : 1;  // +2: write the hi tile explicitly (the conservative form)
  return W2L_OK;
}

extern "C" int w2l_gemm_set_tile(int bn) {
  if (bn != 0 && bn != 128 && bn != 160 && bn != 224 && bn != 256) return fail(W2L_ERR_INVALID_ARGUMENT, "gemm: tile width must be 0 (auto), 128, 160, 224 or 256");
  g_force_bn = bn;
  return W2L_OK;
}

extern "C" int w2l_gemm(void* stream, int kind, int a_mn_major, int b_mn_major, int M, int N, int K, const void* A, int lda, const void* B,
                        int ldb, void* C, int ldc, int c_bf16, const float* bias, int act, int accumulate, const void* aux, int ld_aux,
                        int aux_bf16, int aux_mode, float aux_scale, float dropout_p, unsigned long long seed, int allow_overlap) {
  return gemm_impl(stream, kind, a_mn_major, b_mn_major, M, N, K, A, lda, B, ldb, C, ldc, c_bf16, bias, act, accumulate, aux, ld_aux, aux_bf16,
                   aux_mode, aux_scale, dropout_p, seed, allow_overlap != 0);
}

extern "C" int w2l_gemm_tf32_ex(void* stream_, int a_mn_major, int b_mn_major, int M, int N, int K, const float* A, int lda,
                                const float* B, int ldb, float* C, int ldc, const float* bias, int act, int accumulate,
                                const float* aux, int ld_aux, int aux_mode, float aux_scale, float dropout_p,
                                unsigned long long seed) {
  return gemm_impl(stream_, f32_kind(), a_mn_major, b_mn_major, M, N, K, A, lda, B, ldb, C, ldc, 0, bias, act, accumulate, aux, ld_aux, 0, aux_mode,
                   aux_scale, dropout_p, seed, false);
}

extern "C" int w2l_gemm_tf32(void* stream_, int a_mn_major, int b_mn_major, int M, int N, int K, const float* A, int lda,
                             const float* B, int ldb, float* C, int ldc, const float* bias, int act) {
  return w2l_gemm_tf32_ex(stream_, a_mn_major, b_mn_major, M, N, K, A, lda, B, ldb, C, ldc, bias, act, 0, nullptr, 0, 0, 1.f, 0.f, 0ull);
}

// Same contraction with OVERLAPPING operand rows allowed (lda / ldb smaller than the row length): the TMA tensor map
// takes any 16-byte-multiple row stride, so the im2col matrix of a time convolution over [T][Cin] activations —
// row t = frames t .. t+kw-1, i.e. kw*Cin contiguous floats starting at frame t, row stride Cin — is a zero-copy view.
extern "C" int w2l_gemm_tf32_view(void* stream_, int a_mn_major, int b_mn_major, int M, int N, int K, const float* A, int lda,
                                  const float* B, int ldb, float* C, int ldc, const float* bias, int act, int accumulate) {
  return gemm_impl(stream_, f32_kind(), a_mn_major, b_mn_major, M, N, K, A, lda, B, ldb, C, ldc, 0, bias, act, accumulate, nullptr, 0, 0, 0, 1.f, 0.f,
                   0ull, true);
}

extern "C" int w2l_cast_bf16(void* stream_, long long n, const float* x, void* y) {
  if (n <= 0) return W2L_OK;
  if (!x || !y) return fail(W2L_ERR_INVALID_ARGUMENT, "cast_bf16: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 7) == 0);
  const long long n4 = vec ? n / 4 : 0;
  const long long threads = std::max<long long>(n4, n - 4 * n4);
  cast_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(n4, reinterpret_cast<const float4*>(x), static_cast<uint2*>(y), n, x,
                                                                         static_cast<__nv_bfloat16*>(y));
  W2L_LAUNCH_CHECK("cast_bf16_kernel");
  return W2L_OK;
}
extern "C" int w2l_cast_bf16_rows(void* stream_, long long rows, int cols, int ld_in, int cols_padded, const float* x, void* y) {
  if (rows <= 0 || cols <= 0) return W2L_OK;
  if (!x || !y || cols_padded < cols || ld_in < cols) return fail(W2L_ERR_INVALID_ARGUMENT, "cast_bf16_rows: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long total = rows * cols_padded;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  cast_bf16_rows_kernel<<<grid, 256, 0, stream>>>(rows, cols, ld_in, cols_padded, x, static_cast<__nv_bfloat16*>(y));
  W2L_LAUNCH_CHECK("cast_bf16_rows_kernel");
  return W2L_OK;
}
