// glu_wn_kernels.cu — the element-wise / per-row pieces of the Conv1D+GLU acoustic-model family (sm_100a):
//   WeightNorm   w[r][:] = g[r] * v[r][:] / ||v[r][:]||        (fl::WeightNorm around Conv2D (dim 3) and Linear (dim 0):
//                                                               both normalise per OUTPUT unit, i.e. per contiguous row
//                                                               of the [cout][cin*kw] / [out][in] weight; arch opcode
//                                                               `WN d <layer>`, cpc/SequentialBuilder.cpp:379-386)
//   GLU          y[c] = x[c] * sigmoid(x[c + C/2]) over the channel axis (`GLU 2` after a conv, `GLU 0` after the
//                Linear head; SequentialBuilder.cpp:467-473), with the following Dropout fused in
//   arrange      conv weights [cout][cin][kw] -> K-major GEMM operands [Cout_p][kw*Cin_p] (forward / weight-gradient
//                layout) and [Cin_p][kw*Cout_p] (flipped, data gradient), channel counts padded to multiples of 4
//                with zero rows/columns (TMA row strides must be multiples of 16 bytes); for a conv that feeds a GLU
//                the two halves of the output channels are padded separately so the GLU halves stay aligned.
// All HBM-bound streaming kernels: float4 where the shapes allow, one pass.
#include <cuda_runtime.h>

#include "common.cuh"

namespace w2l {
namespace {

__device__ __forceinline__ float block_sum(float v, float* red /*[33]*/) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__global__ void __launch_bounds__(256) wn_fwd_kernel(int len, const float* __restrict__ v, const float* __restrict__ g,
                                                     float* __restrict__ w, float* __restrict__ inv_norm) {
  __shared__ float red[33];
  const size_t r = blockIdx.x;
  const float* vr = v + r * len;
  float s = 0.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) s += vr[i] * vr[i];
  const float tot = block_sum(s, red);
  const float inv = rsqrtf(fmaxf(tot, 1e-30f));
  if (threadIdx.x == 0) inv_norm[r] = inv;
  const float sc = g[r] * inv;
  for (int i = threadIdx.x; i < len; i += blockDim.x) w[r * len + i] = vr[i] * sc;
}

// dg[r] += <dw, v> / ||v||;  dv += g/||v|| * (dw - v <dw, v> / ||v||^2)
__global__ void __launch_bounds__(256) wn_bwd_kernel(int len, const float* __restrict__ v, const float* __restrict__ g,
                                                     const float* __restrict__ inv_norm, const float* __restrict__ dw,
                                                     float* __restrict__ dv, float* __restrict__ dg) {
  __shared__ float red[33];
  const size_t r = blockIdx.x;
  const float* vr = v + r * len;
  const float* dr = dw + r * len;
  float s = 0.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) s += vr[i] * dr[i];
  const float dot = block_sum(s, red);
  const float inv = inv_norm[r], gi = g[r] * inv, c = dot * inv * inv;
  if (threadIdx.x == 0) dg[r] += dot * inv;
  for (int i = threadIdx.x; i < len; i += blockDim.x) dv[r * len + i] += gi * (dr[i] - vr[i] * c);
}

// row of the padded operand that holds output channel co (GLU split: the two halves are padded separately)
__device__ __forceinline__ int out_row(int co, int cout, int cout_p, int glu_split) {
  if (!glu_split) return co;
  const int h = cout / 2, hp = cout_p / 2;
  return co < h ? co : hp + (co - h);
}

// w [cout][cin][kw] -> fwd [cout_p][kw*cin_p] (k = dk*cin_p + ci), flip [cin_p][kw*cout_p] (k = j*cout_p + row(co),
// tap kw-1-j), bias -> bias_p.  Destinations are zero-filled by the host first.
__global__ void __launch_bounds__(256) conv1d_arrange_kernel(int cin, int cout, int kw, int cin_p, int cout_p, int glu_split,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ fwd, float* __restrict__ flip,
                                                             float* __restrict__ bias_p) {
  const size_t n = (size_t)cout * cin * kw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int dk = (int)(i % kw), ci = (int)((i / kw) % cin), co = (int)(i / ((size_t)kw * cin));
    const int row = out_row(co, cout, cout_p, glu_split);
    const float x = w[i];
    fwd[(size_t)row * kw * cin_p + (size_t)dk * cin_p + ci] = x;
    if (flip) flip[(size_t)ci * kw * cout_p + (size_t)(kw - 1 - dk) * cout_p + row] = x;
  }
  if (bias && bias_p)
    for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < cout; co += gridDim.x * blockDim.x)
      bias_p[out_row(co, cout, cout_p, glu_split)] = bias[co];
}

// gradient of the arranged operand back to the parameter layout: dw[co][ci][dk] += dfwd[row(co)][dk*cin_p + ci]
__global__ void __launch_bounds__(256) conv1d_unarrange_kernel(int cin, int cout, int kw, int cin_p, int cout_p, int glu_split,
                                                               const float* __restrict__ dfwd, float* __restrict__ dw) {
  const size_t n = (size_t)cout * cin * kw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int dk = (int)(i % kw), ci = (int)((i / kw) % cin), co = (int)(i / ((size_t)kw * cin));
    dw[i] += dfwd[(size_t)out_row(co, cout, cout_p, glu_split) * kw * cin_p + (size_t)dk * cin_p + ci];
  }
}

// bias gradient of a padded-row output: dbias[co] += sum_rows dy[row][out_row(co)]
__global__ void __launch_bounds__(256) conv1d_bias_grad_kernel(long long rows, int cout, int cout_p, int glu_split,
                                                               const float* __restrict__ dy, float* __restrict__ dbias) {
  __shared__ float red[33];
  const int co = blockIdx.x;
  const int col = out_row(co, cout, cout_p, glu_split);
  float s = 0.f;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) s += dy[r * cout_p + col];
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0) dbias[co] += tot;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// y[r][c] = x[r][c] * sigmoid(x[r][H + c]) * dropout(r*H + c)
__global__ void __launch_bounds__(256) glu_fwd_kernel(long long rows, int H, const float* __restrict__ x, float* __restrict__ y,
                                                      float drop_p, unsigned long long seed) {
  const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const long long n = rows * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H;
    const int c = (int)(i - r * H);
    const float a = x[r * 2 * H + c], b = x[r * 2 * H + H + c];
    float v = a * sigmoidf_(b);
    if (drop_p > 0.f) v *= dropout_scale(seed, (unsigned long long)i, drop_p, inv_keep);
    y[i] = v;
  }
}
// dx[r][c] = dy * m * sig(b) ; dx[r][H+c] = dy * m * a * sig(b) (1 - sig(b))
__global__ void __launch_bounds__(256) glu_bwd_kernel(long long rows, int H, const float* __restrict__ x, const float* __restrict__ dy,
                                                      float* __restrict__ dx, float drop_p, unsigned long long seed) {
  const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const long long n = rows * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H;
    const int c = (int)(i - r * H);
    const float a = x[r * 2 * H + c], b = x[r * 2 * H + H + c];
    const float s = sigmoidf_(b);
    float d = dy[i];
    if (drop_p > 0.f) d *= dropout_scale(seed, (unsigned long long)i, drop_p, inv_keep);
    dx[r * 2 * H + c] = d * s;
    dx[r * 2 * H + H + c] = d * a * s * (1.0f - s);
  }
}

int blocks_for_n(long long n) { return (int)std::min<long long>((n + 2047) / 2048, 148 * 8); }

}  // namespace
}  // namespace w2l

using namespace w2l;

extern "C" int w2l_weightnorm_fwd(void* stream_, int rows, int len, const float* v, const float* g, float* w, float* inv_norm) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || len <= 0 || !v || !g || !w || !inv_norm) return fail(W2L_ERR_INVALID_ARGUMENT, "weightnorm_fwd: bad arguments");
  wn_fwd_kernel<<<rows, 256, 0, stream>>>(len, v, g, w, inv_norm);
  W2L_LAUNCH_CHECK("wn_fwd_kernel");
  return W2L_OK;
}
extern "C" int w2l_weightnorm_bwd(void* stream_, int rows, int len, const float* v, const float* g, const float* inv_norm,
                                  const float* dw, float* dv, float* dg) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || len <= 0 || !v || !g || !inv_norm || !dw || !dv || !dg) return fail(W2L_ERR_INVALID_ARGUMENT, "weightnorm_bwd: bad arguments");
  wn_bwd_kernel<<<rows, 256, 0, stream>>>(len, v, g, inv_norm, dw, dv, dg);
  W2L_LAUNCH_CHECK("wn_bwd_kernel");
  return W2L_OK;
}
extern "C" int w2l_conv1d_arrange(void* stream_, int cin, int cout, int kw, int cin_p, int cout_p, int glu_split, const float* w,
                                  const float* bias, float* fwd, float* flip, float* bias_p) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (cin <= 0 || cout <= 0 || kw <= 0 || cin_p < cin || cout_p < cout || (cin_p % 4) || (cout_p % 4) || !w || !fwd)
    return fail(W2L_ERR_INVALID_ARGUMENT, "conv1d_arrange: bad arguments (padded channel counts must be multiples of 4)");
  if (glu_split && ((cout % 2) || (cout_p % 8) || cout_p / 2 < cout / 2)) return fail(W2L_ERR_INVALID_ARGUMENT, "conv1d_arrange: bad GLU split padding");
  W2L_CUDA_CHECK(cudaMemsetAsync(fwd, 0, sizeof(float) * (size_t)cout_p * kw * cin_p, stream));
  if (flip) W2L_CUDA_CHECK(cudaMemsetAsync(flip, 0, sizeof(float) * (size_t)cin_p * kw * cout_p, stream));
  if (bias_p) W2L_CUDA_CHECK(cudaMemsetAsync(bias_p, 0, sizeof(float) * (size_t)cout_p, stream));
  conv1d_arrange_kernel<<<blocks_for_n((long long)cout * cin * kw), 256, 0, stream>>>(cin, cout, kw, cin_p, cout_p, glu_split, w, bias, fwd, flip, bias_p);
  W2L_LAUNCH_CHECK("conv1d_arrange_kernel");
  return W2L_OK;
}
extern "C" int w2l_conv1d_unarrange_grad(void* stream_, int cin, int cout, int kw, int cin_p, int cout_p, int glu_split,
                                         const float* dfwd, float* dw, long long rows, const float* dy, float* dbias) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (cin <= 0 || cout <= 0 || kw <= 0 || !dfwd || !dw) return fail(W2L_ERR_INVALID_ARGUMENT, "conv1d_unarrange_grad: bad arguments");
  conv1d_unarrange_kernel<<<blocks_for_n((long long)cout * cin * kw), 256, 0, stream>>>(cin, cout, kw, cin_p, cout_p, glu_split, dfwd, dw);
  W2L_LAUNCH_CHECK("conv1d_unarrange_kernel");
  if (dbias && dy && rows > 0) {
    conv1d_bias_grad_kernel<<<cout, 256, 0, stream>>>(rows, cout, cout_p, glu_split, dy, dbias);
    W2L_LAUNCH_CHECK("conv1d_bias_grad_kernel");
  }
  return W2L_OK;
}
extern "C" int w2l_glu_fwd(void* stream_, long long rows, int half, const float* x, float* y, float dropout_p, unsigned long long seed) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || half <= 0 || !x || !y || dropout_p < 0.f || dropout_p >= 1.f) return fail(W2L_ERR_INVALID_ARGUMENT, "glu_fwd: bad arguments");
  glu_fwd_kernel<<<blocks_for_n(rows * half), 256, 0, stream>>>(rows, half, x, y, dropout_p, seed);
  W2L_LAUNCH_CHECK("glu_fwd_kernel");
  return W2L_OK;
}
extern "C" int w2l_glu_bwd(void* stream_, long long rows, int half, const float* x, const float* dy, float* dx, float dropout_p,
                           unsigned long long seed) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || half <= 0 || !x || !dy || !dx || dropout_p < 0.f || dropout_p >= 1.f) return fail(W2L_ERR_INVALID_ARGUMENT, "glu_bwd: bad arguments");
  glu_bwd_kernel<<<blocks_for_n(rows * half), 256, 0, stream>>>(rows, half, x, dy, dx, dropout_p, seed);
  W2L_LAUNCH_CHECK("glu_bwd_kernel");
  return W2L_OK;
}
