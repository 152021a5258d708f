// CUDA-core fp32 kernels: the PENEO_PREC_FP32 path (exact SiLU, fp32 accumulation) for every
// configuration the reference accepts (any d, any num_layers >= 1, shrink on/off), plus the small
// utility kernels (row casts, weight packing).  The tcgen05 path lives in gemm_tc.cu / pair_heads_tc.cu.
#include "common.cuh"
#include "kernels.h"

namespace peneo {

// ------------------------------------------------------------------------------------------------
// C[M, N] (ldc) = act(A[M, K] (lda) * W[N, K]^T (ldw) + bias[N]) * out_scale
// 64x64 tile, BK = 16, 256 threads, 4x4 register tile.  K % 4 == 0, 16-byte aligned rows.
// ------------------------------------------------------------------------------------------------
template <int ACT>  // 0 none, 1 SiLU
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const float* __restrict__ A, int64_t lda,
                                                       const float* __restrict__ W, int64_t ldw,
                                                       const float* __restrict__ bias, float* __restrict__ C,
                                                       int64_t ldc, int M, int N, int K, float out_scale,
                                                       uint32_t drop_thresh, float drop_scale, uint32_t drop_key) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int lr = tid / 4, lk = (tid % 4) * 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (m0 + lr < M && k0 + lk < K) av = *reinterpret_cast<const float4*>(A + (int64_t)(m0 + lr) * lda + k0 + lk);
    if (n0 + lr < N && k0 + lk < K) bv = *reinterpret_cast<const float4*>(W + (int64_t)(n0 + lr) * ldw + k0 + lk);
    As[lk + 0][lr] = av.x, As[lk + 1][lr] = av.y, As[lk + 2][lr] = av.z, As[lk + 3][lr] = av.w;
    Bs[lk + 0][lr] = bv.x, Bs[lk + 1][lr] = bv.y, Bs[lk + 2][lr] = bv.z, Bs[lk + 3][lr] = bv.w;
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
      float v = acc[r][c] + (bias ? bias[n] : 0.f);
      if (ACT == 1) {
        v = silu_exact(v);
        if (drop_thresh) v = drop_keep(drop_key, drop_thresh, m, n) ? v * drop_scale : 0.f;
      }
      C[(int64_t)m * ldc + n] = v * out_scale;
    }
  }
}

int launch_sgemm_nt(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
                    int M, int N, int K, int act, float out_scale, cudaStream_t st, const DropSpec* drop, uint32_t site) {
  PENEO_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, "sgemm_nt: K, lda, ldw must be multiples of 4");
  if (M == 0 || N == 0) return PENEO_OK;
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  const uint32_t th = (drop && act) ? drop->thresh : 0u;
  const float sc = drop ? drop->scale : 1.f;
  const uint32_t key = drop ? drop_key(*drop, site) : 0u;
  if (act)
    sgemm_nt_kernel<1><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K, out_scale, th, sc, key);
  else
    sgemm_nt_kernel<0><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K, out_scale, 0u, 1.f, 0u);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
// Row cast: src [rows, cols] (row stride in elements) of f32 / bf16 / f16 -> dst contiguous f32 or bf16
// ------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void cast_rows_kernel(const TI* __restrict__ src, int64_t src_stride, TO* __restrict__ dst, int64_t rows,
                                 int cols) {
  const int64_t total = rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols;
    const int c = static_cast<int>(e - r * cols);
    dst[e] = static_cast<TO>(static_cast<float>(src[r * src_stride + c]));
  }
}

int launch_cast_rows(const void* src, int src_dtype, int64_t src_stride, void* dst, int dst_dtype, int64_t rows,
                     int cols, cudaStream_t st) {
  if (rows == 0) return PENEO_OK;
  const int64_t total = rows * cols;
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 16));
#define CAST_CASE(SD, TI, DD, TO)                                                                        \
  if (src_dtype == SD && dst_dtype == DD) {                                                              \
    cast_rows_kernel<TI, TO><<<blocks, 256, 0, st>>>(static_cast<const TI*>(src), src_stride,            \
                                                     static_cast<TO*>(dst), rows, cols);                 \
    PENEO_CUDA_TRY(cudaGetLastError());                                                                  \
    return PENEO_OK;                                                                                     \
  }
  CAST_CASE(PENEO_DT_F32, float, PENEO_DT_F32, float)
  CAST_CASE(PENEO_DT_BF16, __nv_bfloat16, PENEO_DT_F32, float)
  CAST_CASE(PENEO_DT_F16, __half, PENEO_DT_F32, float)
  CAST_CASE(PENEO_DT_F32, float, PENEO_DT_BF16, __nv_bfloat16)
  CAST_CASE(PENEO_DT_BF16, __nv_bfloat16, PENEO_DT_BF16, __nv_bfloat16)
  CAST_CASE(PENEO_DT_F16, __half, PENEO_DT_BF16, __nv_bfloat16)
#undef CAST_CASE
  set_error("cast_rows: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
  return PENEO_E_INVALID;
}

// ------------------------------------------------------------------------------------------------
// Backbone -> decoder seam (model/modeling_peneo.py:134-173): the decoder's input is a strided view of the backbone
// output ([B, seq + visual tokens, H] minus the CLS row / the visual tokens), then nn.Dropout (training), then the first
// GEMM.  One pass does the strip, the optional dropout and the cast to the GEMM's operand type, 8 elements per
// thread: token (b, t) is read at src + b * batch_stride + t * row_stride, written densely at dst[(b * n + t) * cols].
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kSiteInput = 2;  // dropout site of modeling_peneo.py:165 (per-token sites 0 and 1 are the shrink MLP's)
template <typename T>
__device__ __forceinline__ float to_float(T v) { return static_cast<float>(v); }
template <typename T>
__device__ __forceinline__ T from_float(float v) { return static_cast<T>(v); }

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gather_tokens_kernel(const TI* __restrict__ src, int64_t batch_stride, int64_t row_stride,
                                                            int n, TO* __restrict__ dst, int64_t tokens, int cols,
                                                            uint32_t drop_thresh, float drop_scale, uint32_t drop_key_) {
  const int groups = cols / 4;
  const int64_t total = tokens * groups;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = e / groups;
    const int c = static_cast<int>(e - tok * groups) * 4;
    const int64_t b = tok / n;
    const TI* p = src + b * batch_stride + (tok - b * n) * row_stride + c;
    float v[4] = {to_float(p[0]), to_float(p[1]), to_float(p[2]), to_float(p[3])};
    if (drop_thresh) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = drop_keep(drop_key_, drop_thresh, static_cast<uint32_t>(tok), c + k) ? v[k] * drop_scale : 0.f;
    }
    TO* q = dst + tok * cols + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = from_float<TO>(v[k]);
  }
}

int launch_gather_tokens(const void* src, int src_dtype, int64_t batch_stride, int64_t row_stride, int batch, int n, int cols,
                         void* dst, int dst_dtype, const DropSpec* drop, cudaStream_t st) {
  const int64_t tokens = (int64_t)batch * n;
  if (tokens == 0) return PENEO_OK;
  PENEO_REQUIRE(cols % 4 == 0, "gather_tokens: width must be a multiple of 4");
  const int blocks = static_cast<int>(std::min<int64_t>((tokens * (cols / 4) + 255) / 256, 148 * 16));
  const uint32_t th = drop ? drop->thresh : 0u, key = drop ? drop_key(*drop, kSiteInput) : 0u;
  const float sc = drop ? drop->scale : 1.f;
#define GATHER_CASE(SD, TI, DD, TO)                                                                                    \
  if (src_dtype == SD && dst_dtype == DD) {                                                                            \
    gather_tokens_kernel<TI, TO><<<blocks, 256, 0, st>>>(static_cast<const TI*>(src), batch_stride, row_stride, n,     \
                                                         static_cast<TO*>(dst), tokens, cols, th, sc, key);           \
    PENEO_CUDA_TRY(cudaGetLastError());                                                                                \
    return PENEO_OK;                                                                                                   \
  }
  GATHER_CASE(PENEO_DT_F32, float, PENEO_DT_F32, float)
  GATHER_CASE(PENEO_DT_BF16, __nv_bfloat16, PENEO_DT_F32, float)
  GATHER_CASE(PENEO_DT_F16, __half, PENEO_DT_F32, float)
  GATHER_CASE(PENEO_DT_F32, float, PENEO_DT_BF16, __nv_bfloat16)
  GATHER_CASE(PENEO_DT_BF16, __nv_bfloat16, PENEO_DT_BF16, __nv_bfloat16)
  GATHER_CASE(PENEO_DT_F16, __half, PENEO_DT_BF16, __nv_bfloat16)
#undef GATHER_CASE
  set_error("gather_tokens: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
  return PENEO_E_INVALID;
}

// dx[tok, c] *= mask / (1 - p) of the same input dropout (backward of the seam), in place, fp32
__global__ void __launch_bounds__(256) token_dropout_bwd_kernel(float* __restrict__ dx, int64_t tokens, int cols,
                                                                uint32_t drop_thresh, float drop_scale, uint32_t drop_key_) {
  const int64_t total = tokens * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = e / cols;
    const int c = static_cast<int>(e - tok * cols);
    dx[e] = drop_keep(drop_key_, drop_thresh, static_cast<uint32_t>(tok), c) ? dx[e] * drop_scale : 0.f;
  }
}
int launch_token_dropout_bwd(float* dx, int64_t tokens, int cols, const DropSpec* drop, cudaStream_t st) {
  if (!drop || !drop->thresh || tokens == 0) return PENEO_OK;
  const int blocks = static_cast<int>(std::min<int64_t>((tokens * cols + 255) / 256, 148 * 16));
  token_dropout_bwd_kernel<<<blocks, 256, 0, st>>>(dx, tokens, cols, drop->thresh, drop->scale, drop_key(*drop, kSiteInput));
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
// Weight packing
// ------------------------------------------------------------------------------------------------
// dst[r, c] = bf16(scale * src[r * ld + col0 + c])
__global__ void pack_bf16_kernel(const float* __restrict__ src, int64_t ld, int col0, __nv_bfloat16* __restrict__ dst,
                                 int rows, int cols, float scale) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = static_cast<int>(e / cols), c = static_cast<int>(e % cols);
    dst[e] = __float2bfloat16_rn(scale * src[(int64_t)r * ld + col0 + c]);
  }
}
// dst[c, r] = bf16(src[r * ld + c])   (transpose of a [rows, cols] matrix)
__global__ void pack_bf16_T_kernel(const float* __restrict__ src, int64_t ld, __nv_bfloat16* __restrict__ dst, int rows,
                                   int cols) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(e / rows), r = static_cast<int>(e % rows);
    dst[e] = __float2bfloat16_rn(src[(int64_t)r * ld + c]);
  }
}
__global__ void pack_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, float scale) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    dst[e] = src ? scale * src[e] : 0.f;
}
// W_out chunks for the tcgen05 kernel: dst[(c*16 + r) * 128 + f] = r < C_h ? W_out_h[r, 128*(c%3) + f] : 0
struct WoutPtrs {
  const float* w[kNumHeads];
  const float* b[kNumHeads];
};
__global__ void pack_wout_kernel(WoutPtrs p, __nv_bfloat16* __restrict__ dst, float* __restrict__ bout, int d) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 15 * 16 * 128) {
    const int f = e % 128, r = (e / 128) % 16, c = e / (128 * 16);
    const int h = c / 3, f0 = 128 * (c % 3);
    float v = 0.f;
    if (r < head_classes(h)) v = p.w[h][r * d + f0 + f];
    dst[e] = __float2bfloat16_rn(v);
  }
  if (e < kNumHeads * 4) {
    const int h = e / 4, r = e % 4;
    bout[e] = r < head_classes(h) ? p.b[h][r] : 0.f;
  }
}

// wout4[h * d + f] = (W_out_h[0][f], W_out_h[1][f], C_h == 3 ? W_out_h[2][f] : 0, 0)
__global__ void pack_wout4_kernel(WoutPtrs p, float4* __restrict__ dst, int d) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kNumHeads * d) return;
  const int h = e / d, f = e % d;
  const float* w = p.w[h];
  dst[e] = make_float4(w[f], w[d + f], head_classes(h) == 3 ? w[2 * d + f] : 0.f, 0.f);
}

int pack_weights_impl(const peneo_dims& dm, int prec, const peneo_params& P, void* pack, cudaStream_t st) {
  char* base = static_cast<char*>(pack);
  const PackLayout L = pack_layout(dm, prec);
  const int d = dm.d, hid = dm.hid, hin = dm.hin;
  auto f32 = [&](const float* src, size_t off, int64_t n, float scale) -> int {
    pack_f32_kernel<<<static_cast<int>(std::min<int64_t>((n + 255) / 256, 1184)), 256, 0, st>>>(
        src, reinterpret_cast<float*>(base + off), n, scale);
    PENEO_CUDA_TRY(cudaGetLastError());
    return PENEO_OK;
  };
  auto bf = [&](const float* src, int64_t ld, int col0, size_t off, int rows, int cols, float scale) -> int {
    const int64_t n = (int64_t)rows * cols;
    pack_bf16_kernel<<<static_cast<int>(std::min<int64_t>((n + 255) / 256, 1184)), 256, 0, st>>>(
        src, ld, col0, reinterpret_cast<__nv_bfloat16*>(base + off), rows, cols, scale);
    PENEO_CUDA_TRY(cudaGetLastError());
    return PENEO_OK;
  };
  int rc;
#define TRY(x) \
  if ((rc = (x)) != PENEO_OK) return rc
  if (prec == PENEO_PREC_FP32) {
    if (dm.shrink) {
      TRY(f32(P.shrink_w1, L.f_w1, (int64_t)hid * hin, 1.f));
      TRY(f32(P.shrink_b1, L.f_b1, hid, 1.f));
      TRY(f32(P.shrink_w2, L.f_w2, (int64_t)d * hid, 1.f));
      TRY(f32(P.shrink_b2, L.f_b2, d, 1.f));
    }
    TRY(f32(P.combine_w, L.f_wc, (int64_t)d * 2 * d, 1.f));
    TRY(f32(P.combine_b, L.f_bc, d, 1.f));
    for (int h = 0; h < kNumHeads; ++h) {
      for (int l = 0; l + 1 < dm.num_layers; ++l) {
        TRY(f32(P.mid_w[h * 8 + l], L.f_mid_w[h][l], (int64_t)d * d, 1.f));
        TRY(f32(P.mid_b[h * 8 + l], L.f_mid_b[h][l], d, 1.f));
      }
      TRY(f32(P.out_w[h], L.f_out_w[h], (int64_t)head_classes(h) * d, 1.f));
      TRY(f32(P.out_b[h], L.f_out_b[h], head_classes(h), 1.f));
    }
  } else if (!bf16_supported(dm)) {
    // unfused tensor-core forward (pair_heads_generic.cu)
    if (dm.shrink) {
      TRY(bf(P.shrink_w1, hin, 0, L.w1_bf16, hid, hin, 1.f));
      TRY(f32(P.shrink_b1, L.b1, hid, 1.f));
      TRY(bf(P.shrink_w2, hid, 0, L.w2_bf16, d, hid, 1.f));
      TRY(f32(P.shrink_b2, L.b2, d, 1.f));
    }
    TRY(bf(P.combine_w, 2 * d, 0, L.wc_bf16, d, d, 0.5f));
    TRY(bf(P.combine_w, 2 * d, d, L.wc_bf16 + (size_t)d * d * 2, d, d, 0.5f));
    TRY(f32(nullptr, L.bc_half, d, 0.f));
    TRY(f32(P.combine_b, L.bc_half + (size_t)d * 4, d, 0.5f));
    for (int h = 0; h < kNumHeads; ++h) {
      for (int l = 0; l + 1 < dm.num_layers; ++l) {
        TRY(bf(P.mid_w[h * 8 + l], d, 0, L.g_mid_w[h][l], d, d, 1.f));
        TRY(f32(P.mid_b[h * 8 + l], L.g_mid_b[h][l], d, 1.f));
      }
      const int C = head_classes(h);
      PENEO_CUDA_TRY(cudaMemsetAsync(base + L.g_out_w[h], 0, (size_t)32 * d * 2, st));
      PENEO_CUDA_TRY(cudaMemsetAsync(base + L.g_out_b[h], 0, 32 * 4, st));
      TRY(bf(P.out_w[h], d, 0, L.g_out_w[h], C, d, 1.f));
      TRY(f32(P.out_b[h], L.g_out_b[h], C, 1.f));
    }
  } else {
    TRY(bf(P.shrink_w1, hin, 0, L.w1_bf16, hid, hin, 1.f));
    TRY(f32(P.shrink_b1, L.b1, hid, 1.f));
    TRY(bf(P.shrink_w2, hid, 0, L.w2_bf16, d, hid, 1.f));
    TRY(f32(P.shrink_b2, L.b2, d, 1.f));
    TRY(bf(P.combine_w, 2 * d, 0, L.wc_bf16, d, d, 0.5f));
    TRY(bf(P.combine_w, 2 * d, d, L.wc_bf16 + (size_t)d * d * 2, d, d, 0.5f));
    TRY(f32(nullptr, L.bc_half, d, 0.f));
    TRY(f32(P.combine_b, L.bc_half + (size_t)d * 4, d, 0.5f));
    WoutPtrs wp;
    for (int h = 0; h < kNumHeads; ++h) {
      TRY(bf(P.mid_w[h * 8], d, 0, L.wmid_bf16 + (size_t)h * d * d * 2, d, d, 0.5f));
      TRY(f32(P.mid_b[h * 8], L.bmid_half + (size_t)h * d * 4, d, 0.5f));
      wp.w[h] = P.out_w[h], wp.b[h] = P.out_b[h];
    }
    pack_wout_kernel<<<(15 * 16 * 128 + 255) / 256, 256, 0, st>>>(
        wp, reinterpret_cast<__nv_bfloat16*>(base + L.wout_bf16), reinterpret_cast<float*>(base + L.bout), d);
    PENEO_CUDA_TRY(cudaGetLastError());
    // operands of the backward pass
    pack_wout4_kernel<<<(kNumHeads * d + 255) / 256, 256, 0, st>>>(wp, reinterpret_cast<float4*>(base + L.wout_f32x4), d);
    PENEO_CUDA_TRY(cudaGetLastError());
    for (int h = 0; h < kNumHeads; ++h) {
      TRY(bf(P.mid_w[h * 8], d, 0, L.wmid_full_bf16 + (size_t)h * d * d * 2, d, d, 1.f));
      pack_bf16_T_kernel<<<(d * d + 255) / 256, 256, 0, st>>>(
          P.mid_w[h * 8], d, reinterpret_cast<__nv_bfloat16*>(base + L.wmidT_bf16 + (size_t)h * d * d * 2), d, d);
      PENEO_CUDA_TRY(cudaGetLastError());
      TRY(f32(P.mid_b[h * 8], L.bmid_full + (size_t)h * d * 4, d, 1.f));
      TRY(f32(P.out_w[h], L.f_out_w[h], (int64_t)head_classes(h) * d, 1.f));
      TRY(f32(P.out_b[h], L.f_out_b[h], head_classes(h), 1.f));
    }
    TRY(f32(P.shrink_w1, L.f_w1, (int64_t)hid * hin, 1.f));
    TRY(f32(P.shrink_b1, L.f_b1, hid, 1.f));
    TRY(f32(P.shrink_w2, L.f_w2, (int64_t)d * hid, 1.f));
    TRY(f32(P.shrink_b2, L.f_b2, d, 1.f));
    TRY(f32(P.combine_w, L.f_wc, (int64_t)d * 2 * d, 1.f));
    TRY(f32(P.combine_b, L.f_bc, d, 1.f));
  }
#undef TRY
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 pair heads.  One CTA scores TP consecutive pairs of the batch-flattened pair list:
//   S[r, :] = SiLU(A[i_r] + Bm[j_r])                                (shared memory, never global)
//   per head: (num_layers-1) x { H = SiLU(H_prev W^T + b) } ; z = H W_out^T + b_out
// The last hidden layer is not stored: its SiLU output is contracted with W_out in registers.
// ------------------------------------------------------------------------------------------------
struct PairSimtArgs {
  const float* ab;  // [batch*n, 2d]: A | Bm
  int32_t batch, n, d, num_layers;
  int64_t total_pairs;  // batch * P
  int32_t pairs_per_doc;
  const float* mid_w[kNumHeads][kMaxMidLayers];
  const float* mid_b[kNumHeads][kMaxMidLayers];
  const float* out_w[kNumHeads];
  const float* out_b[kNumHeads];
  float* logits[kNumHeads];
  DropSpec drop;
};

template <int TP>
__global__ void __launch_bounds__(256) pair_heads_simt_kernel(const PairSimtArgs a) {
  constexpr int RM = TP / 16;  // rows per thread
  extern __shared__ __align__(16) float smem[];
  const int d = a.d, ldS = d + 4;
  float* S = smem;                    // [TP][ldS]
  float* H0 = S + TP * ldS;           // [TP][ldS] when num_layers >= 3
  float* H1 = H0 + (a.num_layers >= 3 ? TP * ldS : 0);  // when num_layers >= 4
  float* Bs = H1 + (a.num_layers >= 4 ? TP * ldS : 0);  // [16][68]
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16, lane = tid % 32, warp = tid / 32;
  const int64_t g0 = (int64_t)blockIdx.x * TP;

  // ---- S tile
  for (int r = warp; r < TP; r += 8) {
    const int64_t g = g0 + r;
    float* srow = S + r * ldS;
    if (g < a.total_pairs) {
      const int b = static_cast<int>(g / a.pairs_per_doc);
      const int p = static_cast<int>(g - (int64_t)b * a.pairs_per_doc);
      int i, j;
      pair_from_flat(p, a.n, i, j);
      const float* ai = a.ab + ((int64_t)b * a.n + i) * 2 * d;
      const float* bj = a.ab + ((int64_t)b * a.n + j) * 2 * d + d;
      for (int c = lane; c < d; c += 32) srow[c] = silu_exact(ai[c] + bj[c]);
    } else {
      for (int c = lane; c < d; c += 32) srow[c] = 0.f;
    }
  }
  __syncthreads();

  for (int h = 0; h < kNumHeads; ++h) {
    const int C = head_classes(h);
    float z[RM][3];
#pragma unroll
    for (int r = 0; r < RM; ++r) z[r][0] = z[r][1] = z[r][2] = 0.f;
    const float* in = S;
    if (a.num_layers == 1) {
      // z = S W_out^T : each thread takes columns tx*4 + 64*q of its rows
      for (int n0 = 0; n0 < d; n0 += 64) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = n0 + tx * 4 + c;
          if (col >= d) continue;
#pragma unroll
          for (int r = 0; r < RM; ++r) {
            const float v = in[(ty * RM + r) * ldS + col];
            for (int cc = 0; cc < C; ++cc) z[r][cc] = fmaf(a.out_w[h][cc * d + col], v, z[r][cc]);
          }
        }
      }
    }
    for (int l = 0; l + 1 < a.num_layers; ++l) {
      const bool last = (l + 2 == a.num_layers);
      float* out = (l % 2 == 0) ? H0 : H1;
      const float* W = a.mid_w[h][l];
      const float* bias = a.mid_b[h][l];
      for (int n0 = 0; n0 < d; n0 += 64) {
        float acc[RM][4];
#pragma unroll
        for (int r = 0; r < RM; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
        for (int k0 = 0; k0 < d; k0 += 16) {
          {
            const int lr = tid / 4, lk = (tid % 4) * 4;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + lr < d && k0 + lk < d) bv = *reinterpret_cast<const float4*>(W + (int64_t)(n0 + lr) * d + k0 + lk);
            Bs[(lk + 0) * 68 + lr] = bv.x, Bs[(lk + 1) * 68 + lr] = bv.y;
            Bs[(lk + 2) * 68 + lr] = bv.z, Bs[(lk + 3) * 68 + lr] = bv.w;
          }
          __syncthreads();
          const int kmax = min(16, d - k0);
          for (int kk = 0; kk < kmax; ++kk) {
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk * 68 + tx * 4]);
#pragma unroll
            for (int r = 0; r < RM; ++r) {
              const float av = in[(ty * RM + r) * ldS + k0 + kk];
              acc[r][0] = fmaf(av, b4.x, acc[r][0]);
              acc[r][1] = fmaf(av, b4.y, acc[r][1]);
              acc[r][2] = fmaf(av, b4.z, acc[r][2]);
              acc[r][3] = fmaf(av, b4.w, acc[r][3]);
            }
          }
          __syncthreads();
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = n0 + tx * 4 + c;
          if (col >= d) continue;
          const float bc = bias[col];
#pragma unroll
          for (int r = 0; r < RM; ++r) {
            float m = silu_exact(acc[r][c] + bc);
            if (a.drop.thresh) {  // Dropout after every hidden SiLU (model/peneo_decoder.py:261)
              const uint32_t g = static_cast<uint32_t>(g0 + ty * RM + r);
              m = drop_keep(drop_key(a.drop, site_head(h, l)), a.drop.thresh, g, col) ? m * a.drop.scale : 0.f;
            }
            if (last) {
              for (int cc = 0; cc < C; ++cc) z[r][cc] = fmaf(a.out_w[h][cc * d + col], m, z[r][cc]);
            } else {
              out[(ty * RM + r) * ldS + col] = m;
            }
          }
        }
      }
      __syncthreads();
      in = out;
    }
    // reduce z over the 16 threads (tx) that share the rows; they are 16 consecutive lanes
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        float v = z[r][cc];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        z[r][cc] = v;
      }
    if (tx == 0) {
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        const int64_t g = g0 + ty * RM + r;
        if (g < a.total_pairs)
          for (int cc = 0; cc < C; ++cc) a.logits[h][g * C + cc] = z[r][cc] + a.out_b[h][cc];
      }
    }
  }
}

int launch_pair_heads_simt(const peneo_dims& dm, const void* pack, const float* ab, int batch, int n,
                           float* const logits[kNumHeads], cudaStream_t st, const DropSpec* drop) {
  const PackLayout L = pack_layout(dm, PENEO_PREC_FP32);
  const char* base = static_cast<const char*>(pack);
  PENEO_REQUIRE(dm.num_layers >= 1 && dm.num_layers <= kMaxMidLayers + 1, "num_layers %d out of range", dm.num_layers);
  PENEO_REQUIRE(dm.d % 4 == 0, "decoder hidden size must be a multiple of 4");
  PairSimtArgs a{};
  a.ab = ab, a.batch = batch, a.n = n, a.d = dm.d, a.num_layers = dm.num_layers;
  a.drop = drop ? *drop : DropSpec{0u, 1.f, 0u, 0u};
  a.pairs_per_doc = static_cast<int32_t>(pair_count(n));
  a.total_pairs = (int64_t)batch * a.pairs_per_doc;
  for (int h = 0; h < kNumHeads; ++h) {
    for (int l = 0; l + 1 < dm.num_layers; ++l) {
      a.mid_w[h][l] = reinterpret_cast<const float*>(base + L.f_mid_w[h][l]);
      a.mid_b[h][l] = reinterpret_cast<const float*>(base + L.f_mid_b[h][l]);
    }
    a.out_w[h] = reinterpret_cast<const float*>(base + L.f_out_w[h]);
    a.out_b[h] = reinterpret_cast<const float*>(base + L.f_out_b[h]);
    a.logits[h] = logits[h];
  }
  if (a.total_pairs == 0) return PENEO_OK;
  const int nbuf = 1 + (dm.num_layers >= 3) + (dm.num_layers >= 4);
  auto smem_for = [&](int tp) { return (size_t)(nbuf * tp * (dm.d + 4) + 16 * 68) * sizeof(float); };
  int tp = 64;
  while (tp > 16 && smem_for(tp) > 220 * 1024) tp /= 2;
  PENEO_REQUIRE(smem_for(tp) <= 220 * 1024, "decoder hidden size %d too large for the fp32 pair kernel", dm.d);
  const size_t smem = smem_for(tp);
  const int64_t blocks = (a.total_pairs + tp - 1) / tp;
  PENEO_REQUIRE(blocks < (1ll << 31), "too many pairs for one launch");
#define LAUNCH(TP)                                                                                             \
  {                                                                                                            \
    PENEO_CUDA_TRY(cudaFuncSetAttribute(pair_heads_simt_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        static_cast<int>(smem)));                                              \
    pair_heads_simt_kernel<TP><<<static_cast<unsigned>(blocks), 256, smem, st>>>(a);                            \
  }
  if (tp == 64) LAUNCH(64) else if (tp == 32) LAUNCH(32) else LAUNCH(16)
#undef LAUNCH
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
