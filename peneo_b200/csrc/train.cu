// Backward of the decoder heads: the driver (peneo_heads_bwd) for both precisions, the CUDA-core fp32 kernels
// (PENEO_PREC_FP32: every configuration the reference accepts) and the element-wise / reduction kernels of the bf16
// mode (whose tensor-core kernels live in pair_bwd_tc.cu, gemm_bwd_tc.cu, gemm_tf32.cu).  What autograd does for
// model/peneo_decoder.py:349-363 in the reference, restated as explicit kernels:
//
//   g_z   = dlogits                                  (from peneo_pair_loss_bwd / peneo_pair_loss_ohem)
//   dW_out += g_z^T m ,  db_out += sum g_z           m = SiLU(u_last)
//   g_u   = (g_z W_out) * SiLU'(u)                   ... back through every hidden layer ...
//   dW_mid += g_u^T h_in , db_mid += sum g_u
//   g_s   = sum_heads g_u0 W_mid0 ;  g_v = g_s * SiLU'(a_i + b_j)
//   dA[i] = sum_{j >= i} g_v(i, j) ;  dBm[j] = sum_{i <= j} g_v(i, j)
//   then the per-token chain (combine_fc halves, shrink MLP) down to d sequence_output.
//
// Nothing of size [P, D] is kept from the forward pass: the pair activations are recomputed one
// chunk of whole pair-rows at a time (a chunk = rows i0..i1 of one document = a contiguous range
// of flat pair indices), so the workspace is O(chunk * D * num_layers), not O(P * D).
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {

namespace {

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dsilu_exact(float x) {
  const float s = sigmoid_exact(x);
  return s * (1.0f + x * (1.0f - s));
}

// ------------------------------------------------------------------------------------------------
// C[M, N] = op(A) op(B) (+ bias[N]),  written (mode 0), added (mode 1) or atomically added (mode 2,
// split-K over gridDim.z).  TA: A is stored [K, M]; else [M, K].  TB: B is stored [N, K]; else [K, N].
// Optional second output C2 = SiLU(C) (modes 0 only).  64x64 tile, BK = 16, 256 threads, 4x4/thread.
// Vector loads run along the contiguous dimension, which must be a multiple of 4 (checked by the
// launcher); the other dimension is guarded per element.
// ------------------------------------------------------------------------------------------------
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                    int64_t ldb, float* __restrict__ C, int64_t ldc, int M, int N, int K,
                                                    int kslice, int mode, const float* __restrict__ bias,
                                                    float* __restrict__ C2, int64_t ldc2, uint32_t drop_thresh,
                                                    float drop_scale, uint32_t drop_key_, uint32_t row0) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kbeg = blockIdx.z * kslice, kend = min(K, kbeg + kslice);
  float acc[4][4] = {};
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
    if (TA) {  // A[k, m]: vector along m
      const int k = tid / 16, m4 = (tid % 16) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < kend && m0 + m4 < M) v = *reinterpret_cast<const float4*>(A + (int64_t)(k0 + k) * lda + m0 + m4);
      *reinterpret_cast<float4*>(&As[k][m4]) = v;
    } else {  // A[m, k]: vector along k
      const int m = tid / 4, k4 = (tid % 4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + m < M && k0 + k4 < kend) v = *reinterpret_cast<const float4*>(A + (int64_t)(m0 + m) * lda + k0 + k4);
      As[k4 + 0][m] = v.x, As[k4 + 1][m] = v.y, As[k4 + 2][m] = v.z, As[k4 + 3][m] = v.w;
    }
    if (TB) {  // B[n, k]: vector along k
      const int n = tid / 4, k4 = (tid % 4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + n < N && k0 + k4 < kend) v = *reinterpret_cast<const float4*>(B + (int64_t)(n0 + n) * ldb + k0 + k4);
      Bs[k4 + 0][n] = v.x, Bs[k4 + 1][n] = v.y, Bs[k4 + 2][n] = v.z, Bs[k4 + 3][n] = v.w;
    } else {  // B[k, n]: vector along n
      const int k = tid / 16, n4 = (tid % 16) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < kend && n0 + n4 < N) v = *reinterpret_cast<const float4*>(B + (int64_t)(k0 + k) * ldb + n0 + n4);
      *reinterpret_cast<float4*>(&Bs[k][n4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(ar[r], br[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = m0 + ty * 4 + r;
    if (m >= M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int n = n0 + tx * 4 + c;
      if (n >= N) continue;
      float v = acc[r][c];
      float* dst = C + (int64_t)m * ldc + n;
      if (mode == 2) {
        atomicAdd(dst, v);
      } else {
        if (bias) v += bias[n];
        if (mode == 1) v += *dst;
        *dst = v;
        if (C2) {
          float h = v * sigmoid_exact(v);
          if (drop_thresh) h = drop_keep(drop_key_, drop_thresh, row0 + m, n) ? h * drop_scale : 0.f;
          C2[(int64_t)m * ldc2 + n] = h;
        }
      }
    }
  }
}

struct Gemm {
  bool ta = false, tb = false;
  const float* A = nullptr;
  int64_t lda = 0;
  const float* B = nullptr;
  int64_t ldb = 0;
  float* C = nullptr;
  int64_t ldc = 0;
  int M = 0, N = 0, K = 0;
  int mode = 0;  // 0 store, 1 add, 2 split-K atomic add
  const float* bias = nullptr;
  float* C2 = nullptr;  // optional second output: Dropout(SiLU(C))
  int64_t ldc2 = 0;
  uint32_t drop_thresh = 0, drop_key = 0, row0 = 0;  // dropout mask of C2: element (row0 + m, n)
  float drop_scale = 1.f;
  float* scratch = nullptr;  // bf16 mode: room for transposed copies of A ([M, K]) and B ([N, K]) -> K-major TF32 GEMM
};

// out[c, r] = in[r, c]   (in: [R, C] with row stride ld_in ; out: [C, R] with row stride ld_out)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int64_t ld_in, int R, int C,
                                                        float* __restrict__ out, int64_t ld_out) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32, tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int y = ty; y < 32; y += 8)
    if (r0 + y < R && c0 + tx < C) tile[y][tx] = in[(int64_t)(r0 + y) * ld_in + c0 + tx];
  __syncthreads();
  for (int y = ty; y < 32; y += 8)
    if (c0 + y < C && r0 + tx < R) out[(int64_t)(c0 + y) * ld_out + r0 + tx] = tile[tx][y];
}

// `tensor_cores`: bf16 mode — run on tcgen05 kind::tf32 (gemm_tf32.cu) when the shape qualifies; the fp32 mode
// always uses the exact CUDA-core kernel below.
int run_gemm(const Gemm& g, cudaStream_t st, bool tensor_cores = false) {
  if (g.M == 0 || g.N == 0) return PENEO_OK;
  if (tensor_cores) {
    Tf32Gemm t;
    t.ta = g.ta, t.tb = g.tb, t.A = g.A, t.lda = g.lda, t.B = g.B, t.ldb = g.ldb, t.C = g.C, t.ldc = g.ldc;
    t.M = g.M, t.N = g.N, t.K = g.K, t.mode = g.mode, t.bias = g.bias, t.C2 = g.C2, t.ldc2 = g.ldc2;
    t.drop_thresh = g.drop_thresh, t.drop_key = g.drop_key, t.row0 = g.row0, t.drop_scale = g.drop_scale;
    if (g.scratch && (g.ta || !g.tb)) {
      // kind::tf32 only takes K-major operands here: transpose what is stored the other way round (a few MB, a
      // few microseconds) instead of falling back to the CUDA-core SGEMM
      float* sc = g.scratch;
      const int64_t kp = (g.K + 3) / 4 * 4;
      if (g.ta) {
        transpose_kernel<<<dim3((g.M + 31) / 32, (g.K + 31) / 32), 256, 0, st>>>(g.A, g.lda, g.K, g.M, sc, kp);
        PENEO_CUDA_TRY(cudaGetLastError());
        t.A = sc, t.lda = kp, t.ta = false, sc += (int64_t)g.M * kp;
      }
      if (!g.tb) {
        transpose_kernel<<<dim3((g.N + 31) / 32, (g.K + 31) / 32), 256, 0, st>>>(g.B, g.ldb, g.K, g.N, sc, kp);
        PENEO_CUDA_TRY(cudaGetLastError());
        t.B = sc, t.ldb = kp, t.tb = true;
      }
    }
    if (gemm_tf32_supported(t)) return launch_gemm_tf32(t, st);
  }
  const bool ok = (g.ta ? (g.M % 4 == 0) : (g.K % 4 == 0)) && (g.tb ? (g.K % 4 == 0) : (g.N % 4 == 0)) &&
                  g.lda % 4 == 0 && g.ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(g.B) & 15) == 0;
  PENEO_REQUIRE(ok, "sgemm: operand not 4-element aligned (M=%d N=%d K=%d lda=%lld ldb=%lld ta=%d tb=%d)", g.M, g.N, g.K,
                (long long)g.lda, (long long)g.ldb, (int)g.ta, (int)g.tb);
  int splits = 1, kslice = (g.K + 15) / 16 * 16;
  const int tiles = ((g.N + 63) / 64) * ((g.M + 63) / 64);
  if (g.mode == 2) {
    splits = std::max(1, std::min((g.K + 255) / 256, (148 * 4 + tiles - 1) / tiles));
    kslice = ((g.K + splits - 1) / splits + 15) / 16 * 16;
    splits = (g.K + kslice - 1) / kslice;
  }
  if (kslice == 0) kslice = 16;
  dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, splits);
#define GO(TA, TB)                                                                                                   \
  sgemm_kernel<TA, TB><<<grid, 256, 0, st>>>(g.A, g.lda, g.B, g.ldb, g.C, g.ldc, g.M, g.N, g.K, kslice, g.mode, g.bias, \
                                             g.C2, g.ldc2, g.drop_thresh, g.drop_scale, g.drop_key, g.row0)
  if (g.ta && g.tb) GO(true, true);
  else if (g.ta) GO(true, false);
  else if (g.tb) GO(false, true);
  else GO(false, false);
#undef GO
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
// pair-chunk kernels.  A chunk is the flat pair range [p0, p0 + rows) of document b.
// ------------------------------------------------------------------------------------------------
// S[r, :] = SiLU(a_i + b_j)
__global__ void build_s_kernel(const float* __restrict__ ab, int b, int n, int d, int p0, float* __restrict__ S) {
  const int r = blockIdx.x;
  int i, j;
  pair_from_flat(p0 + r, n, i, j);
  const float* ai = ab + ((int64_t)b * n + i) * 2 * d;
  const float* bj = ab + ((int64_t)b * n + j) * 2 * d + d;
  for (int f = threadIdx.x; f < d; f += blockDim.x) S[(int64_t)r * d + f] = silu_exact(ai[f] + bj[f]);
}

// Last hidden layer + output layer of one head, 128 rows per CTA:
//   m = SiLU(U) ; G = (dz W_out) * SiLU'(U) ; dW_out += dz^T m ; db_out += sum dz ; db_mid += sum G
template <int C>
__global__ void __launch_bounds__(128) head_out_bwd_kernel(const float* __restrict__ U, const float* __restrict__ dz,
                                                           const float* __restrict__ Wout, int rows, int d,
                                                           float* __restrict__ G, float* __restrict__ dWout,
                                                           float* __restrict__ dbout, float* __restrict__ dbmid,
                                                           uint32_t drop_thresh, float drop_scale, uint32_t drop_key_,
                                                           uint32_t row0) {
  __shared__ float sdz[128][C];
  const int r0 = blockIdx.x * 128, nr = min(128, rows - r0);
  for (int e = threadIdx.x; e < nr * C; e += 128) sdz[e / C][e % C] = dz[(int64_t)r0 * C + e];
  __syncthreads();
  if (threadIdx.x < C) {
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += sdz[r][threadIdx.x];
    atomicAdd(&dbout[threadIdx.x], s);
  }
  for (int f = threadIdx.x; f < d; f += 128) {
    float w[C], aw[C], ab_ = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) w[c] = Wout[c * d + f], aw[c] = 0.f;
    for (int r = 0; r < nr; ++r) {
      const float u = U[(int64_t)(r0 + r) * d + f];
      const float s = sigmoid_exact(u);
      float m = u * s;
      const float ds = s * (1.0f + u * (1.0f - s));
      // the forward pass fed Dropout(m) to W_out: same mask on m (for dW_out) and on the gradient through it
      const float ms = drop_thresh ? (drop_keep(drop_key_, drop_thresh, row0 + r0 + r, f) ? drop_scale : 0.f) : 1.f;
      m *= ms;
      float gm = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        gm = fmaf(sdz[r][c], w[c], gm);
        aw[c] = fmaf(sdz[r][c], m, aw[c]);
      }
      const float g = gm * ms * ds;
      G[(int64_t)(r0 + r) * d + f] = g;
      ab_ += g;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) atomicAdd(&dWout[c * d + f], aw[c]);
    atomicAdd(&dbmid[f], ab_);
  }
}

// num_layers == 1: z = S W_out^T + b_out.  dS (+)= dz W_out ; dW_out += dz^T S ; db_out += sum dz
template <int C>
__global__ void __launch_bounds__(128) out_only_bwd_kernel(const float* __restrict__ S, const float* __restrict__ dz,
                                                           const float* __restrict__ Wout, int rows, int d,
                                                           float* __restrict__ dS, int accumulate,
                                                           float* __restrict__ dWout, float* __restrict__ dbout) {
  __shared__ float sdz[128][C];
  const int r0 = blockIdx.x * 128, nr = min(128, rows - r0);
  for (int e = threadIdx.x; e < nr * C; e += 128) sdz[e / C][e % C] = dz[(int64_t)r0 * C + e];
  __syncthreads();
  if (threadIdx.x < C) {
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += sdz[r][threadIdx.x];
    atomicAdd(&dbout[threadIdx.x], s);
  }
  for (int f = threadIdx.x; f < d; f += 128) {
    float w[C], aw[C];
#pragma unroll
    for (int c = 0; c < C; ++c) w[c] = Wout[c * d + f], aw[c] = 0.f;
    for (int r = 0; r < nr; ++r) {
      const int64_t at = (int64_t)(r0 + r) * d + f;
      const float s = S[at];
      float g = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        g = fmaf(sdz[r][c], w[c], g);
        aw[c] = fmaf(sdz[r][c], s, aw[c]);
      }
      dS[at] = accumulate ? dS[at] + g : g;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) atomicAdd(&dWout[c * d + f], aw[c]);
  }
}

// G = dH * mask / (1 - p) * SiLU'(U) (in place over dH; dH is the gradient w.r.t. Dropout(SiLU(U))) ;
// db += column sums.  Grid: (row blocks of 32, column blocks of 128).
__global__ void __launch_bounds__(128) act_bwd_kernel(float* __restrict__ dH, const float* __restrict__ U, int rows,
                                                      int d, int64_t ld_dh, int64_t ld_u, float* __restrict__ db,
                                                      uint32_t drop_thresh, float drop_scale, uint32_t drop_key_,
                                                      uint32_t row0) {
  const int r0 = blockIdx.x * 32, nr = min(32, rows - r0);
  const int f = blockIdx.y * 128 + threadIdx.x;
  if (f >= d) return;
  float acc = 0.f;
#pragma unroll 8
  for (int r = 0; r < nr; ++r) {
    float g = dH[(int64_t)(r0 + r) * ld_dh + f] * dsilu_exact(U[(int64_t)(r0 + r) * ld_u + f]);
    if (drop_thresh) g = drop_keep(drop_key_, drop_thresh, row0 + r0 + r, f) ? g * drop_scale : 0.f;
    dH[(int64_t)(r0 + r) * ld_dh + f] = g;
    acc += g;
  }
  if (db) atomicAdd(&db[f], acc);
}

// column sums of X[rows, d] (ld) into out[d] (atomic); same grid
__global__ void __launch_bounds__(128) colsum_kernel(const float* __restrict__ X, int rows, int d, int64_t ld,
                                                     float* __restrict__ out) {
  const int r0 = blockIdx.x * 32, nr = min(32, rows - r0);
  const int f = blockIdx.y * 128 + threadIdx.x;
  if (f >= d) return;
  float acc = 0.f;
#pragma unroll 8
  for (int r = 0; r < nr; ++r) acc += X[(int64_t)(r0 + r) * ld + f];
  atomicAdd(&out[f], acc);
}

// g_v(i, j) = dS(i, j) * SiLU'(a_i + b_j), reduced on the fly:
//   dA[t]   = sum_{j >= t} g_v(t, j)                      for pair-rows t in [i0, i1)
//   dBm[t] += sum_{i in [i0, min(t, i1 - 1)]} g_v(i, t)
// SiLU' is evaluated where it is consumed (twice per pair) instead of in a separate read-modify-write pass
// over dS.  Grid: (token t of the document, column blocks of 128); deterministic (no atomics).
__global__ void __launch_bounds__(128) gv_reduce_kernel(const float* __restrict__ dS, const float* __restrict__ ab, int b,
                                                        int n, int d, int i0, int i1, float* __restrict__ dab) {
  const int t = blockIdx.x;
  const int f = blockIdx.y * 128 + threadIdx.x;
  if (f >= d) return;
  const int p0 = row_start(i0, n);
  const float* abd = ab + (int64_t)b * n * 2 * d;  // this document's [n, 2d] projections
  float* out = dab + ((int64_t)b * n + t) * 2 * d;
  if (t >= i0 && t < i1) {
    const float at = abd[(int64_t)t * 2 * d + f];
    const float* src = dS + (int64_t)(row_start(t, n) - p0) * d + f;
    const float* bj = abd + (int64_t)t * 2 * d + d + f;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int j = t;
    for (; j + 3 < n; j += 4) {
      const int64_t o = (int64_t)(j - t);
      a0 = fmaf(src[o * d], dsilu_exact(at + bj[o * 2 * d]), a0);
      a1 = fmaf(src[(o + 1) * d], dsilu_exact(at + bj[(o + 1) * 2 * d]), a1);
      a2 = fmaf(src[(o + 2) * d], dsilu_exact(at + bj[(o + 2) * 2 * d]), a2);
      a3 = fmaf(src[(o + 3) * d], dsilu_exact(at + bj[(o + 3) * 2 * d]), a3);
    }
    for (; j < n; ++j) a0 = fmaf(src[(int64_t)(j - t) * d], dsilu_exact(at + bj[(int64_t)(j - t) * 2 * d]), a0);
    out[f] = (a0 + a1) + (a2 + a3);
  }
  const int ihi = min(t, i1 - 1);
  if (ihi >= i0) {
    const float bt = abd[(int64_t)t * 2 * d + d + f];
    float a0 = 0.f, a1 = 0.f;
    int i = i0;
    for (; i + 1 <= ihi; i += 2) {
      a0 = fmaf(dS[(int64_t)(row_start(i, n) - p0 + (t - i)) * d + f], dsilu_exact(abd[(int64_t)i * 2 * d + f] + bt), a0);
      a1 = fmaf(dS[(int64_t)(row_start(i + 1, n) - p0 + (t - i - 1)) * d + f],
                dsilu_exact(abd[(int64_t)(i + 1) * 2 * d + f] + bt), a1);
    }
    if (i <= ihi)
      a0 = fmaf(dS[(int64_t)(row_start(i, n) - p0 + (t - i)) * d + f], dsilu_exact(abd[(int64_t)i * 2 * d + f] + bt), a0);
    out[d + f] += a0 + a1;
  }
}

// Single-pass variant for the bf16 mode (whose weight gradients already use fp32 atomics): a block owns the tile
// rows [it, it + 8) x columns [jt, jt + 8) of one document's pair matrix, one thread per 4 features (float4 loads:
// every pair is one contiguous 4 d-byte read, 8 of them in flight per thread).  b_j and the 8 column sums stay in
// registers; 8 + 8 vector REDs per 64 pairs.  Tiles below the diagonal exit.
constexpr int kGvR = 8, kGvC = 8;
// One tile.  INTERIOR (every pair of the tile has i < ie, i <= j < je — all but the diagonal and the last row / column
// of tiles): no predicates, addresses advance by constants.  Per element: h = a_i / 2 + b_j / 2 (the halves are formed
// once per row / column), t = tanh(h), SiLU'(2h) = sg (1 + h (1 - t)) with sg = (1 + t) / 2.
template <bool INTERIOR>
__device__ __forceinline__ void gv_tile(const __nv_bfloat16* __restrict__ dS, const float* __restrict__ abd,
                                        float* __restrict__ out, int n, int d, int64_t ld, int p0, int it, int jt, int ie,
                                        int je) {
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // the 8 pairs (i, jt .. jt + 7) of row i: 8-byte loads, zero outside the triangle / the document
  auto load_row = [&](int i, uint2 (&v)[kGvC]) {
    // pair (i, j) lives at flat row row_start(i) + (j - i)
    const __nv_bfloat16* src = dS + ((int64_t)(row_start(i, n) - p0) + (jt - i)) * d;
#pragma unroll
    for (int c = 0; c < kGvC; ++c) {
      if (INTERIOR) v[c] = *reinterpret_cast<const uint2*>(src + c * d);
      else v[c] = (i < ie && jt + c >= i && jt + c < je) ? *reinterpret_cast<const uint2*>(src + (int64_t)c * d) : make_uint2(0u, 0u);
    }
  };
  uint2 v[2][kGvC];
  load_row(it, v[0]);
  float4 hb[kGvC], cs[kGvC];
#pragma unroll
  for (int c = 0; c < kGvC; ++c) {
    const float4 b4 = (INTERIOR || jt + c < je) ? *reinterpret_cast<const float4*>(abd + (int64_t)(jt + c) * ld + d) : zero4;
    hb[c] = make_float4(0.5f * b4.x, 0.5f * b4.y, 0.5f * b4.z, 0.5f * b4.w);
    cs[c] = zero4;
  }
  auto gvf = [](uint32_t ds_bits, float ha, float hbv) {
    const float h = ha + hbv;
    const float t = ptx::tanh_approx(h);
    const float sg = fmaf(0.5f, t, 0.5f);
    return __uint_as_float(ds_bits) * sg * fmaf(h, 1.f - t, 1.f);
  };
#pragma unroll
  for (int r = 0; r < kGvR; ++r) {
    const int i = it + r;
    if (r + 1 < kGvR) load_row(i + 1, v[(r + 1) & 1]);  // next row in flight while this one is reduced
    if (INTERIOR || i < ie) {
      const float4 a4 = *reinterpret_cast<const float4*>(abd + (int64_t)i * ld);
      const float4 ha = make_float4(0.5f * a4.x, 0.5f * a4.y, 0.5f * a4.z, 0.5f * a4.w);
      float4 rs = zero4;
#pragma unroll
      for (int c = 0; c < kGvC; ++c) {
        const uint2 w = v[r & 1][c];
        const float gx = gvf(w.x << 16, ha.x, hb[c].x), gy = gvf(w.x & 0xFFFF0000u, ha.y, hb[c].y);
        const float gz = gvf(w.y << 16, ha.z, hb[c].z), gw = gvf(w.y & 0xFFFF0000u, ha.w, hb[c].w);
        rs.x += gx, rs.y += gy, rs.z += gz, rs.w += gw;
        cs[c].x += gx, cs[c].y += gy, cs[c].z += gz, cs[c].w += gw;
      }
      atomicAdd(reinterpret_cast<float4*>(out + (int64_t)i * ld), rs);
    }
  }
#pragma unroll
  for (int c = 0; c < kGvC; ++c)
    if (INTERIOR || (jt + c < je && jt + c >= it)) atomicAdd(reinterpret_cast<float4*>(out + (int64_t)(jt + c) * ld + d), cs[c]);
}
#ifndef PENEO_GV_MINBLOCKS
#define PENEO_GV_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(128, PENEO_GV_MINBLOCKS) gv_reduce_tile_kernel(const __nv_bfloat16* __restrict__ dS,
                                                            const float* __restrict__ ab, int b, int n, int d, int i0,
                                                            int i1, float* __restrict__ dab) {
  const int f = 4 * threadIdx.x;  // blockDim.x = d / 4
  const int it = i0 + blockIdx.y * kGvR, jt = blockIdx.x * kGvC;
  const int ie = min(it + kGvR, i1), je = min(jt + kGvC, n);
  if (je <= it) return;  // every pair of the tile has j < i
  const int64_t ld = 2 * (int64_t)d;
  const float* abd = ab + (int64_t)b * n * ld + f;
  float* out = dab + (int64_t)b * n * ld + f;
  const int p0 = row_start(i0, n);
  if (it + kGvR <= i1 && jt + kGvC <= n && it + kGvR - 1 <= jt) gv_tile<true>(dS + f, abd, out, n, d, ld, p0, it, jt, ie, je);
  else gv_tile<false>(dS + f, abd, out, n, d, ld, p0, it, jt, ie, je);
}

constexpr int kD16 = 384;

// bf16 backward, output layer: T1 leaves per-CTA partial sums of dz^T M ([ctas][3][1920] fp32); db_out is the column
// sum of dz.
//   dW_out[k][c][f] += sum_cta part[cta][c][384 k + f] ;  db_out[k][c] += sum_r dz_k[r][c]
struct DzPtrs {
  const float* p[kNumHeads];
};
struct HeadPtrs {  // per-head gradient destinations, passed by value (no device-side pointer table, no sync)
  float* p[kNumHeads];
};
constexpr int kDwPartCtas = 320;  // upper bound of T1's grid (one CTA per SM) and of T1e's (two per SM)
__global__ void __launch_bounds__(128) dwout_finish_kernel(const float* __restrict__ part, int ctas,
                                                           HeadPtrs dWout) {
  const int fs = blockIdx.x * 128 + threadIdx.x;  // stacked feature index in [0, 1920)
  const int c = blockIdx.y;
  const int k = fs / kD16, f = fs - k * kD16;
  if (c >= head_classes(k)) return;
  float acc = 0.f;
  for (int b = 0; b < ctas; ++b) acc += part[((size_t)b * 3 + c) * (5 * kD16) + fs];
  atomicAdd(&dWout.p[k][c * kD16 + f], acc);
}
__global__ void __launch_bounds__(256) dbout_kernel(DzPtrs dz, int rows, HeadPtrs dbout) {
  const int k = blockIdx.y, C = head_classes(k);
  const float* p = dz.p[k];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < rows; r += (int64_t)gridDim.x * 256) {
    a0 += p[r * C], a1 += p[r * C + 1];
    if (C == 3) a2 += p[r * C + 2];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o), a1 += __shfl_xor_sync(0xffffffffu, a1, o), a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if (threadIdx.x % 32 == 0) {
    atomicAdd(&dbout.p[k][0], a0), atomicAdd(&dbout.p[k][1], a1);
    if (C == 3) atomicAdd(&dbout.p[k][2], a2);
  }
}

inline size_t fl(size_t n) { return align_up(n * sizeof(float), 1024); }

struct Plan {
  int64_t tokens;
  int chunk_rows_max;  // pair rows (flat) per chunk
  int nU, nH;          // per-chunk [rows, d] buffers for pre-activations / hidden activations
  size_t off_x, off_u1, off_y1, off_u2, off_y, off_ab, off_dab, off_dy, off_dy1;
  size_t off_S, off_dS, off_G, off_U, off_H;
  // bf16 pair part (T1 + MN-major GEMMs): S [rows, 384], G / M [rows, 1920] bf16, bf16 per-token projections
  size_t off_S16, off_Gc, off_dwpart, off_ab16, off_tokws, tokws_bytes, off_tr, off_tr2;
  size_t total;
};

Plan make_plan(const peneo_dims& dm, int prec, int batch, int n) {
  Plan p{};
  p.tokens = (int64_t)batch * n;
  const size_t T = p.tokens, d = dm.d, hid = dm.shrink ? dm.hid : 0, hin = dm.hin;
  p.nU = dm.num_layers - 1;
  p.nH = dm.num_layers >= 3 ? dm.num_layers - 2 : 0;
  const bool tc = prec == PENEO_PREC_BF16;   // fused tcgen05 pair kernels (T1 + bf16 GEMMs)
  const bool tf = prec == PENEO_PREC_TF32;   // fp32 buffers, every GEMM on tcgen05 kind::tf32 (needs K-major copies)
  const int nbuf = tc ? 4 : 3 + p.nU + p.nH + (tf ? 2 : 0);  // fp32: S, dS, G, U.., H.. (+ two transposed operands) ; bf16: dS (fp32) + S, G (bf16, 1 + 5 halves)
  // Pair buffers: up to 524 288 (bf16) / 262 144 pairs per chunk within ~3 GB (measured: per-chunk launch / wave-quantisation overheads
  // make 64 K-pair chunks 25 % slower end to end), never less than one full pair row (n pairs).  A chunk may span
  // several documents.
  int64_t rows = (int64_t)(3072ull << 20) / ((int64_t)nbuf * d * 4);
  rows = std::min<int64_t>(rows, tf ? 65536 : (tc ? 524288 : 262144));  // bf16: 524 288 pairs (3 GB) measure 2 % faster than 262 144
  if (const char* e = getenv("PENEO_BWD_CHUNK_ROWS")) rows = std::max(1, atoi(e));  // test hook: force small chunks
  rows = std::max<int64_t>(n, rows);
  rows = std::min<int64_t>(rows, (int64_t)batch * pair_count(n));
  p.chunk_rows_max = static_cast<int>(rows);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off += bytes;
    return at;
  };
  p.off_x = take(fl(T * hin));
  p.off_u1 = take(fl(T * hid)), p.off_y1 = take(fl(T * hid));
  p.off_u2 = take(fl(T * (dm.shrink ? d : 0))), p.off_y = take(fl(T * (dm.shrink ? d : 0)));
  p.off_ab = take(fl(T * 2 * d)), p.off_dab = take(fl(T * 2 * d));
  p.off_dy = take(fl(T * d)), p.off_dy1 = take(fl(T * hid));
  const size_t cb = fl((size_t)rows * d);
  p.off_dS = take(cb);
  if (tc) {
    const size_t rp = (rows + 127) / 128 * 128;
    p.off_S16 = take(align_up(rp * d * 2, 1024));
    p.off_Gc = take(align_up(rp * 5 * d * 2, 1024));
    p.off_dwpart = take(fl((size_t)kDwPartCtas * 3 * 5 * d));
    p.off_ab16 = take(align_up(T * 2 * d * 2, 1024));
    p.tokws_bytes = peneo_token_proj_workspace_bytes(&dm, PENEO_PREC_BF16, p.tokens);
    p.off_tokws = take(align_up(p.tokws_bytes, 1024));
    {
      const size_t wmax = std::max(std::max(hin, hid), 2 * d), tp = (T + 3) / 4 * 4;
      p.off_tr = take(fl(std::max(2 * tp * wmax, wmax * wmax)));  // transposed GEMM operands (activations or a weight)
    }
  } else {
    p.off_S = take(cb), p.off_G = take(cb);
    p.off_U = take(cb * p.nU), p.off_H = take(cb * p.nH);
    if (tf) {
      // K-major copies for kind::tf32: both [rows, d] operands of dW = G^T Hin (or a [d, d] weight), and the per-token chain's
      const size_t rp = (rows + 3) / 4 * 4;
      p.off_tr2 = take(fl(std::max<size_t>(2 * rp * d, d * d)));
      const size_t wmax = std::max(std::max(hin, hid), 2 * d), tp = (T + 3) / 4 * 4;
      p.off_tr = take(fl(std::max(2 * tp * wmax, wmax * wmax)));
    }
  }
  p.total = off + 1024;
  return p;
}

}  // namespace

size_t heads_bwd_workspace_bytes(const peneo_dims& dm, int prec, int batch, int n) {
  return make_plan(dm, prec, batch, n).total;
}

int launch_heads_bwd(const peneo_dims& dm, int prec, const void* pack, const void* x, int x_dtype, int64_t x_row_stride,
                     int batch, int n, const float* const dlogits[kNumHeads], const peneo_grads& gr, float* dx,
                     void* workspace, cudaStream_t st, const DropSpec* drop_in, const FusedLossBwd* fused_in) {
  const PackLayout L = pack_layout(dm, prec == PENEO_PREC_TF32 ? PENEO_PREC_FP32 : prec);
  const DropSpec drop = drop_in ? *drop_in : DropSpec{0u, 1.f, 0u, 0u};
  auto with_drop = [&](Gemm& gm, uint32_t site, uint32_t row0) {  // C2 = Dropout(SiLU(C)) of the forward pass
    gm.drop_thresh = drop.thresh, gm.drop_scale = drop.scale, gm.drop_key = drop_key(drop, site), gm.row0 = row0;
  };
  const bool tc = prec == PENEO_PREC_BF16;    // fused tcgen05 pair kernels
  const bool tcg = prec != PENEO_PREC_FP32;   // GEMMs on tensor cores (kind::tf32 for everything outside the fused kernels)
  const char* pk = static_cast<const char*>(pack);
  const Plan pl = make_plan(dm, prec, batch, n);
  char* ws = static_cast<char*>(workspace);
  const int d = dm.d, hid = dm.hid, hin = dm.hin, T = static_cast<int>(pl.tokens), NL = dm.num_layers;
  const int64_t P = pair_count(n);
  int rc;
#define TRY(x_)                              \
  do {                                       \
    if ((rc = (x_)) != PENEO_OK) return rc;  \
  } while (0)
  auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto W = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  auto zero = [&](float* p, size_t count) -> int {
    if (p && count) PENEO_CUDA_TRY(cudaMemsetAsync(p, 0, count * sizeof(float), st));
    return PENEO_OK;
  };

  // ---- zero the gradient outputs (everything below accumulates)
  if (dm.shrink) {
    TRY(zero(gr.shrink_w1, (size_t)hid * hin));
    TRY(zero(gr.shrink_b1, hid));
    TRY(zero(gr.shrink_w2, (size_t)d * hid));
    TRY(zero(gr.shrink_b2, d));
  }
  TRY(zero(gr.combine_w, (size_t)d * 2 * d));
  TRY(zero(gr.combine_b, d));
  for (int h = 0; h < kNumHeads; ++h) {
    for (int l = 0; l + 1 < NL; ++l) {
      TRY(zero(gr.mid_w[h * 8 + l], (size_t)d * d));
      TRY(zero(gr.mid_b[h * 8 + l], d));
    }
    TRY(zero(gr.out_w[h], (size_t)head_classes(h) * d));
    TRY(zero(gr.out_b[h], head_classes(h)));
  }

  // ---- recompute the per-token chain, keeping pre-activations
  const float* xin;
  int64_t ldx;
  if (x_dtype == PENEO_DT_F32 && x_row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    xin = static_cast<const float*>(x), ldx = x_row_stride;
  } else {
    TRY(launch_cast_rows(x, x_dtype, x_row_stride, F(pl.off_x), PENEO_DT_F32, T, hin, st));
    xin = F(pl.off_x), ldx = hin;
  }
  const float* y = xin;
  int64_t ldy = ldx;
  Gemm g;
  if (dm.shrink) {
    g = Gemm{}, g.tb = true, g.A = xin, g.lda = ldx, g.B = W(L.f_w1), g.ldb = hin, g.C = F(pl.off_u1), g.ldc = hid;
    g.M = T, g.N = hid, g.K = hin, g.bias = W(L.f_b1), g.C2 = F(pl.off_y1), g.ldc2 = hid;
    with_drop(g, kSiteTok0, 0);
    TRY(run_gemm(g, st, tcg));
    g = Gemm{}, g.tb = true, g.A = F(pl.off_y1), g.lda = hid, g.B = W(L.f_w2), g.ldb = hid, g.C = F(pl.off_u2), g.ldc = d;
    g.M = T, g.N = d, g.K = hid, g.bias = W(L.f_b2), g.C2 = F(pl.off_y), g.ldc2 = d;
    with_drop(g, kSiteTok1, 0);
    TRY(run_gemm(g, st, tcg));
    y = F(pl.off_y), ldy = d;
  }
  float* ab = F(pl.off_ab);
  g = Gemm{}, g.tb = true, g.A = y, g.lda = ldy, g.B = W(L.f_wc), g.ldb = 2 * d, g.C = ab, g.ldc = 2 * d;
  g.M = T, g.N = d, g.K = d;
  TRY(run_gemm(g, st, tcg));
  g.B = W(L.f_wc) + d, g.C = ab + d, g.bias = W(L.f_bc);
  TRY(run_gemm(g, st, tcg));

  // ---- pair part, one chunk of whole pair-rows at a time
  float* dab = F(pl.off_dab);
  TRY(zero(dab, (size_t)T * 2 * d));
  HeadPtrs d_outw{}, d_outb{};
  int dw_ctas = 0;
  if (tc) {
    int dev = 0;
    PENEO_CUDA_TRY(cudaGetDevice(&dev));
    PENEO_CUDA_TRY(cudaDeviceGetAttribute(&dw_ctas, cudaDevAttrMultiProcessorCount, dev));  // T1's grid never exceeds it
    if (fused_in && fused_in->save_h) dw_ctas = std::min(2 * dw_ctas, pair_bwd_elem_max_ctas());  // T1e: two CTAs per SM
    PENEO_REQUIRE(dw_ctas <= kDwPartCtas, "device has %d SMs, dW_out partial buffer holds %d", dw_ctas, kDwPartCtas);
    TRY(zero(F(pl.off_dwpart), (size_t)dw_ctas * 3 * 5 * d));
    // bf16 per-token projections (0.5-scaled A | Bm) for T1, exactly what the forward pass used (same dropout);
    // not needed when the forward pass saved s and the pre-activations (T1e reads neither a_i nor b_j)
    if (!(fused_in && fused_in->save_h))
      TRY(token_proj_fwd_bf16(dm, L, pk, x, x_dtype, x_row_stride, T, ws + pl.off_ab16, ws + pl.off_tokws, st,
                              drop.thresh ? &drop : nullptr));
    for (int h = 0; h < kNumHeads; ++h) d_outw.p[h] = gr.out_w[h], d_outb.p[h] = gr.out_b[h];
  }
  float *S = F(pl.off_S), *dS = F(pl.off_dS), *G = F(pl.off_G);
  const size_t cstride = fl((size_t)pl.chunk_rows_max * d) / sizeof(float);
  // A chunk is a contiguous range [g0, g0 + rows) of the batch-flat pair list made of whole pair-rows: segments
  // (document b, rows i0..i1) that follow each other in memory.  Everything up to dS works on the flat range; only
  // the dA / dBm reduction is per segment.
  struct Seg {
    int b, i0, i1, rows;
    int64_t off;  // first row of the segment inside the chunk
  };
  std::vector<Seg> segs;
  // saved activations: T1e once over the whole batch followed by the plain dS GEMM per chunk (default), or — PENEO_T1F=1,
  // experimental — T1f, the same transform inside the dS GEMM's operand stage (gemm_ds_fused.cu: parity-green, but
  // measured 1.46 ms per 524 288-pair chunk against 0.71 + 0.58 ms for T1e + the plain GEMM)
  const bool t1f = [] { const char* e = getenv("PENEO_T1F"); return e && atoi(e) != 0; }();
  if (tc && fused_in && fused_in->save_h && !t1f) {
    // activations saved by the forward pass: h -> G in place for the WHOLE batch in one launch (T1e needs no per-chunk
    // scratch); the chunk loop below only runs the two GEMMs and the dA / dBm reduction on views of it
    PENEO_REQUIRE((int64_t)batch * P < (1ll << 31), "heads_bwd: too many pairs for one launch of the saved-activation backward");
    FusedLossBwd fl = *fused_in;
    for (int h = 0; h < kNumHeads; ++h) fl.dbout[h] = gr.out_b[h];
    TRY(launch_pair_bwd_elem(pack, L, 0, static_cast<int>((int64_t)batch * P), fl, fused_in->save_h, F(pl.off_dwpart), nullptr, st,
                             drop.thresh ? &drop : nullptr));
  }
  int cb = 0, ci0 = 0;  // next (document, pair-row) not yet assigned to a chunk
  while (cb < batch) {
    segs.clear();
    const int64_t g0 = (int64_t)cb * P + row_start(ci0, n);
    int rows = 0;
    while (cb < batch) {
      int i1 = ci0;
      while (i1 < n && rows + (row_start(i1 + 1, n) - row_start(ci0, n)) <= pl.chunk_rows_max) ++i1;
      if (i1 == ci0) {
        if (rows > 0) break;  // chunk full
        i1 = ci0 + 1;         // (cannot happen: chunk_rows_max >= n)
      }
      const int seg_rows = row_start(i1, n) - row_start(ci0, n);  // row_start(n, n) == P
      segs.push_back(Seg{cb, ci0, i1, seg_rows, rows});
      rows += seg_rows;
      ci0 = i1;
      if (ci0 == n) ci0 = 0, ++cb;
      else break;  // the document did not fit completely: the chunk is full
    }
    {
      const int rb = (rows + 127) / 128;
      const uint32_t row0 = static_cast<uint32_t>(g0);  // batch-flat index of the chunk's first pair
      if (tc) {
        __nv_bfloat16* S16 = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_S16);
        __nv_bfloat16* Gc = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_Gc);
        const bool saved = fused_in && fused_in->save_h;
        if (saved) {  // the forward pass stored h and s for every pair: G is formed in place, no scratch copies
          S16 = fused_in->save_s + g0 * d;
          Gc = fused_in->save_h + g0 * (5 * d);
        }
        const __nv_bfloat16* ab16 = reinterpret_cast<const __nv_bfloat16*>(ws + pl.off_ab16);
        // T1: regenerate S and u, store S and G = (dz W_out) SiLU'(u), reduce dW_out += dz^T SiLU(u) into the per-CTA
        // partial sums (tcgen05, K2's structure; on CTA pairs unless PENEO_T1_PAIR=0).  With the fused loss, dz is
        // computed inside T1 from logits + tags and db_out is reduced there as well.
        static const bool t1_pair = [] { const char* e = getenv("PENEO_T1_PAIR"); return !e || atoi(e) != 0; }();
        FusedLossBwd fl{};
        if (fused_in) {
          fl = *fused_in;
          for (int h = 0; h < kNumHeads; ++h) fl.dbout[h] = gr.out_b[h];
        }
        if (saved) {
          // (G of this chunk already stands in save_h: see above)
        } else if (t1_pair || fused_in)
          TRY(launch_pair_bwd_prep_pair(pack, L, ab16, n, g0, rows, dlogits, fused_in ? &fl : nullptr, S16, Gc,
                                        F(pl.off_dwpart), st, drop.thresh ? &drop : nullptr));
        else
          TRY(launch_pair_bwd_prep(pack, L, ab16, n, g0, rows, dlogits, S16, Gc, F(pl.off_dwpart), st,
                                   drop.thresh ? &drop : nullptr));
        // dS = G W_mid (all heads in one K = 1920 GEMM) ; dW_mid += G^T S
        // (CTA-pair kernels by default; PENEO_BWD_PAIR=0 selects the single-CTA ones for A/B studies)
        static const bool gemm_pair = [] { const char* e = getenv("PENEO_BWD_PAIR"); return !e || atoi(e) != 0; }();
        if (saved && t1f)
          TRY(launch_gemm_ds_fused(pack, L, Gc, reinterpret_cast<const __nv_bfloat16*>(pk + L.wmid_full_bf16),
                                   reinterpret_cast<__nv_bfloat16*>(dS), g0, rows, fl, F(pl.off_dwpart), st,
                                   drop.thresh ? &drop : nullptr));
        else
          TRY((gemm_pair ? launch_gemm_ds_pair : launch_gemm_ds)(Gc, reinterpret_cast<const __nv_bfloat16*>(pk + L.wmid_full_bf16),
                                                                reinterpret_cast<__nv_bfloat16*>(dS), rows, st));
        float *dwm[kNumHeads], *dbm[kNumHeads];
        DzPtrs dzp{};
        for (int h = 0; h < kNumHeads; ++h) {
          dwm[h] = gr.mid_w[h * 8], dbm[h] = gr.mid_b[h * 8];
          if (!fused_in) dzp.p[h] = dlogits[h] + g0 * head_classes(h);
        }
        TRY((gemm_pair ? launch_gemm_dw_pair : launch_gemm_dw)(Gc, S16, dwm, dbm, rows, st));  // + db_mid (column sums of G)
        if (!fused_in) {  // (the fused T1 reduces db_out itself)
          dbout_kernel<<<dim3(std::min((rows + 1023) / 1024, 296), kNumHeads), 256, 0, st>>>(dzp, rows, d_outb);
          PENEO_CUDA_TRY(cudaGetLastError());
        }
      } else {
      for (const Seg& sg : segs) {
        build_s_kernel<<<sg.rows, 128, 0, st>>>(ab, sg.b, n, d, row_start(sg.i0, n), S + sg.off * d);
        PENEO_CUDA_TRY(cudaGetLastError());
      }
      for (int h = 0; h < kNumHeads; ++h) {
        const int C = head_classes(h);
        const float* dz = dlogits[h] + g0 * C;
        if (NL == 1) {
          if (C == 2)
            out_only_bwd_kernel<2><<<rb, 128, 0, st>>>(S, dz, W(L.f_out_w[h]), rows, d, dS, h > 0, gr.out_w[h], gr.out_b[h]);
          else
            out_only_bwd_kernel<3><<<rb, 128, 0, st>>>(S, dz, W(L.f_out_w[h]), rows, d, dS, h > 0, gr.out_w[h], gr.out_b[h]);
          PENEO_CUDA_TRY(cudaGetLastError());
          continue;
        }
        // forward recompute through the hidden layers: U_l = H_l W_l^T + b_l ; H_{l+1} = SiLU(U_l)
        const float* in = S;
        for (int l = 0; l + 1 < NL; ++l) {
          float* U = F(pl.off_U) + (size_t)l * cstride;
          float* Hn = (l + 2 < NL) ? F(pl.off_H) + (size_t)l * cstride : nullptr;
          g = Gemm{}, g.tb = true, g.A = in, g.lda = d, g.B = W(L.f_mid_w[h][l]), g.ldb = d, g.C = U, g.ldc = d;
          g.M = rows, g.N = d, g.K = d, g.bias = W(L.f_mid_b[h][l]), g.C2 = Hn, g.ldc2 = d;
          with_drop(g, site_head(h, l), row0);
          TRY(run_gemm(g, st, tcg));
          in = Hn;
        }
        {
          const float* U = F(pl.off_U) + (size_t)(NL - 2) * cstride;
          if (C == 2)
            head_out_bwd_kernel<2><<<rb, 128, 0, st>>>(U, dz, W(L.f_out_w[h]), rows, d, G, gr.out_w[h], gr.out_b[h],
                                                       gr.mid_b[h * 8 + NL - 2], drop.thresh, drop.scale,
                                                       drop_key(drop, site_head(h, NL - 2)), row0);
          else
            head_out_bwd_kernel<3><<<rb, 128, 0, st>>>(U, dz, W(L.f_out_w[h]), rows, d, G, gr.out_w[h], gr.out_b[h],
                                                       gr.mid_b[h * 8 + NL - 2], drop.thresh, drop.scale,
                                                       drop_key(drop, site_head(h, NL - 2)), row0);
          PENEO_CUDA_TRY(cudaGetLastError());
        }
        // G_l = gradient w.r.t. the pre-activation of hidden layer l.  Once G_l exists U_l is dead, so
        // dH_l = G_l W_l is written into U_l's buffer and turned into G_{l-1} in place.
        const float* Gcur = G;
        for (int l = NL - 2; l >= 0; --l) {
          const float* Hin = (l == 0) ? S : F(pl.off_H) + (size_t)(l - 1) * cstride;
          // dW_l[out, in] += sum_r G_l[r, out] Hin[r, in]
          float* tr2 = prec == PENEO_PREC_TF32 ? F(pl.off_tr2) : nullptr;  // K-major copies for kind::tf32
          g = Gemm{}, g.scratch = tr2, g.ta = true, g.A = Gcur, g.lda = d, g.B = Hin, g.ldb = d, g.C = gr.mid_w[h * 8 + l], g.ldc = d;
          g.M = d, g.N = d, g.K = rows, g.mode = 2;
          TRY(run_gemm(g, st, tcg));
          // dHin = G_l W_l
          g = Gemm{}, g.scratch = tr2, g.A = Gcur, g.lda = d, g.B = W(L.f_mid_w[h][l]), g.ldb = d, g.M = rows, g.N = d, g.K = d;
          if (l == 0) {
            g.C = dS, g.ldc = d, g.mode = h > 0 ? 1 : 0;
            TRY(run_gemm(g, st, tcg));
          } else {
            float* Ul = F(pl.off_U) + (size_t)l * cstride;
            g.C = Ul, g.ldc = d, g.mode = 0;
            TRY(run_gemm(g, st, tcg));
            act_bwd_kernel<<<dim3((rows + 31) / 32, (d + 127) / 128), 128, 0, st>>>(
                Ul, F(pl.off_U) + (size_t)(l - 1) * cstride, rows, d, d, d, gr.mid_b[h * 8 + l - 1], drop.thresh, drop.scale,
                drop_key(drop, site_head(h, l - 1)), row0);
            PENEO_CUDA_TRY(cudaGetLastError());
            Gcur = Ul;
          }
        }
      }
      }  // fp32 pair part
      if (tc) {
        for (const Seg& sg : segs) {
          gv_reduce_tile_kernel<<<dim3((n + kGvC - 1) / kGvC, (sg.i1 - sg.i0 + kGvR - 1) / kGvR), d / 4, 0, st>>>(
              reinterpret_cast<const __nv_bfloat16*>(dS) + sg.off * d, ab, sg.b, n, d, sg.i0, sg.i1, dab);
          PENEO_CUDA_TRY(cudaGetLastError());
        }
      } else for (const Seg& sg : segs) {
        gv_reduce_kernel<<<dim3(n, (d + 127) / 128), 128, 0, st>>>(dS + sg.off * d, ab, sg.b, n, d, sg.i0, sg.i1, dab);
        PENEO_CUDA_TRY(cudaGetLastError());
      }
    }
  }

  if (tc) {
    dwout_finish_kernel<<<dim3(15, 3), 128, 0, st>>>(F(pl.off_dwpart), dw_ctas, d_outw);
    PENEO_CUDA_TRY(cudaGetLastError());
  }

  // ---- per-token chain backward
  const int tb = (T + 31) / 32;
  float* tr = tcg ? F(pl.off_tr) : nullptr;
  // db_c = column sums of dBm
  colsum_kernel<<<dim3(tb, (d + 127) / 128), 128, 0, st>>>(dab + d, T, d, 2 * d, gr.combine_b);
  PENEO_CUDA_TRY(cudaGetLastError());
  // dW_c[:, :d] = dA^T y ; dW_c[:, d:] = dBm^T y
  g = Gemm{}, g.scratch = tr, g.ta = true, g.A = dab, g.lda = 2 * d, g.B = y, g.ldb = ldy, g.C = gr.combine_w, g.ldc = 2 * d;
  g.M = d, g.N = d, g.K = T, g.mode = 2;
  TRY(run_gemm(g, st, tcg));
  g.A = dab + d, g.C = gr.combine_w + d;
  TRY(run_gemm(g, st, tcg));
  // dy = dA W_c[:, :d] + dBm W_c[:, d:]
  float* dy = (dm.shrink || dx == nullptr) ? F(pl.off_dy) : dx;
  const int64_t lddy = (dm.shrink || dx == nullptr) ? d : hin;
  g = Gemm{}, g.scratch = tr, g.A = dab, g.lda = 2 * d, g.B = W(L.f_wc), g.ldb = 2 * d, g.C = dy, g.ldc = lddy, g.M = T, g.N = d, g.K = d;
  TRY(run_gemm(g, st, tcg));
  g.A = dab + d, g.B = W(L.f_wc) + d, g.mode = 1;
  TRY(run_gemm(g, st, tcg));
  if (dm.shrink) {
    // G2 = dy * SiLU'(u2) ; db2 ; dW2 = G2^T y1 ; dy1 = G2 W2
    act_bwd_kernel<<<dim3(tb, (d + 127) / 128), 128, 0, st>>>(dy, F(pl.off_u2), T, d, d, d, gr.shrink_b2, drop.thresh,
                                                              drop.scale, drop_key(drop, kSiteTok1), 0u);
    PENEO_CUDA_TRY(cudaGetLastError());
    g = Gemm{}, g.scratch = tr, g.ta = true, g.A = dy, g.lda = d, g.B = F(pl.off_y1), g.ldb = hid, g.C = gr.shrink_w2, g.ldc = hid;
    g.M = d, g.N = hid, g.K = T, g.mode = 2;
    TRY(run_gemm(g, st, tcg));
    float* dy1 = F(pl.off_dy1);
    g = Gemm{}, g.scratch = tr, g.A = dy, g.lda = d, g.B = W(L.f_w2), g.ldb = hid, g.C = dy1, g.ldc = hid, g.M = T, g.N = hid, g.K = d;
    TRY(run_gemm(g, st, tcg));
    // G1 = dy1 * SiLU'(u1) ; db1 ; dW1 = G1^T x ; dx = G1 W1
    act_bwd_kernel<<<dim3(tb, (hid + 127) / 128), 128, 0, st>>>(dy1, F(pl.off_u1), T, hid, hid, hid, gr.shrink_b1,
                                                                drop.thresh, drop.scale, drop_key(drop, kSiteTok0), 0u);
    PENEO_CUDA_TRY(cudaGetLastError());
    g = Gemm{}, g.scratch = tr, g.ta = true, g.A = dy1, g.lda = hid, g.B = xin, g.ldb = ldx, g.C = gr.shrink_w1, g.ldc = hin;
    g.M = hid, g.N = hin, g.K = T, g.mode = 2;
    TRY(run_gemm(g, st, tcg));
    if (dx) {
      g = Gemm{}, g.scratch = tr, g.A = dy1, g.lda = hid, g.B = W(L.f_w1), g.ldb = hin, g.C = dx, g.ldc = hin, g.M = T, g.N = hin, g.K = hid;
      TRY(run_gemm(g, st, tcg));
    }
  }
#undef TRY
  return PENEO_OK;
}

}  // namespace peneo
