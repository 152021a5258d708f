// Unfused tensor-core forward of the pair heads for the configurations the fused K2 does not cover
// (PENEO_PREC_BF16 with shrink off, d != 384 or num_layers != 2).  Same math as
// model/peneo_decoder.py:149-177, 258-269, 355-363, restated per chunk of the batch-flat pair list:
//
//   S = SiLU(a_i + b_j)                                    [rows, d] bf16   (build_s_bf16_kernel)
//   per head:  H <- SiLU(H W_l^T + b_l)  for every hidden layer   gemm_tc2, bf16 output, fused activation
//              z  = H W_out^T  (W_out zero-padded to 32 rows)      gemm_tc2, fp32 output [rows, 32]
//              logits = z[:, :C] + b_out                            extract_logits_kernel
//
// The [rows, d] activations cross HBM once per layer (K2 keeps them in TMEM), so this path is HBM / GEMM-epilogue
// bound rather than at the tensor roofline — but it is 10-30x the fp32 CUDA-core kernel for d = 768.
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {
namespace {

// one CTA per pair, one thread per 8 features; ab = [tokens, 2d] bf16 : 0.5 A | 0.5 Bm
__global__ void __launch_bounds__(128) build_s_bf16_kernel(const __nv_bfloat16* __restrict__ ab, int n, int d,
                                                           int64_t pairs_per_doc, int64_t g0, __nv_bfloat16* __restrict__ S) {
  const int64_t gp = g0 + blockIdx.x;
  const int64_t b = gp / pairs_per_doc;
  int i, j;
  pair_from_flat(static_cast<int>(gp - b * pairs_per_doc), n, i, j);
  const __nv_bfloat16* ai = ab + (b * n + i) * (2 * (int64_t)d);
  const __nv_bfloat16* bj = ab + (b * n + j) * (2 * (int64_t)d) + d;
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    const uint4 av = *reinterpret_cast<const uint4*>(ai + c), bv = *reinterpret_cast<const uint4*>(bj + c);
    const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = __uint_as_float(aw[k] << 16) + __uint_as_float(bw[k] << 16);
      const float hi = __uint_as_float(aw[k] & 0xFFFF0000u) + __uint_as_float(bw[k] & 0xFFFF0000u);
      o[k] = ptx::pack_bf16x2(ptx::silu_from_half(lo), ptx::silu_from_half(hi));
    }
    *reinterpret_cast<uint4*>(S + (int64_t)blockIdx.x * d + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// logits[r, c] = z[r, c] + bias[c]  for c < C   (z: [rows, 32] fp32)
__global__ void __launch_bounds__(256) extract_logits_kernel(const float* __restrict__ z, const float* __restrict__ bias, int C,
                                                             int64_t rows, float* __restrict__ logits) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float4 v = *reinterpret_cast<const float4*>(z + r * 32);
  float* dst = logits + r * C;
  dst[0] = v.x + bias[0], dst[1] = v.y + bias[1];
  if (C == 3) dst[2] = v.z + bias[2];
}

}  // namespace

int launch_pair_heads_generic(const peneo_dims& dm, const void* pack, const PackLayout& L, const __nv_bfloat16* ab, int batch,
                              int n, float* const logits[kNumHeads], cudaStream_t st, const DropSpec* drop) {
  const char* pk = static_cast<const char*>(pack);
  const int d = dm.d, NL = dm.num_layers;
  const int64_t P = pair_count(n), total = (int64_t)batch * P;
  if (total == 0) return PENEO_OK;
  // chunk of pairs: S + two ping-pong hidden buffers (bf16) + z (fp32 [rows, 32]) within ~512 MB
  int64_t chunk = (512ll << 20) / ((int64_t)d * 2 * 3 + 128);
  chunk = std::max<int64_t>(128, std::min<int64_t>(chunk / 128 * 128, std::min<int64_t>(total, 262144)));
  const size_t act_bytes = align_up((size_t)chunk * d * 2, 1024);
  {
    // keep the stream-ordered pool's memory across calls (the default threshold of 0 hands it back to the driver at
    // every synchronisation, and mapping half a gigabyte again costs milliseconds per call)
    static bool pool_ready = false;
    if (!pool_ready) {
      int dev = 0;
      cudaMemPool_t pool;
      PENEO_CUDA_TRY(cudaGetDevice(&dev));
      PENEO_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
      uint64_t keep = ~0ull;
      PENEO_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      pool_ready = true;
    }
  }
  char* ws = nullptr;
  PENEO_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&ws), 3 * act_bytes + (size_t)chunk * 32 * 4, st));
  __nv_bfloat16* S = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* H[2] = {reinterpret_cast<__nv_bfloat16*>(ws + act_bytes), reinterpret_cast<__nv_bfloat16*>(ws + 2 * act_bytes)};
  float* z = reinterpret_cast<float*>(ws + 3 * act_bytes);
  int rc = PENEO_OK;
  for (int64_t g0 = 0; g0 < total && rc == PENEO_OK; g0 += chunk) {
    const int64_t rows = std::min(chunk, total - g0);
    build_s_bf16_kernel<<<static_cast<unsigned>(rows), std::min(128, d / 8), 0, st>>>(ab, n, d, P, g0, S);
    if (cudaGetLastError() != cudaSuccess) { rc = PENEO_E_CUDA; break; }
    for (int h = 0; h < kNumHeads && rc == PENEO_OK; ++h) {
      const __nv_bfloat16* in = S;
      for (int l = 0; l + 1 < NL && rc == PENEO_OK; ++l) {
        __nv_bfloat16* out = H[l & 1];
        // (training mode: the head's nn.Dropout after this SiLU, model/peneo_decoder.py:261, mask row = batch-flat pair)
        rc = launch_gemm_tc2(in, d, reinterpret_cast<const __nv_bfloat16*>(pk + L.g_mid_w[h][l]), d,
                             reinterpret_cast<const float*>(pk + L.g_mid_b[h][l]), out, d, rows, d, d, 0, 1, st, 1, drop,
                             site_head(h, l), static_cast<uint32_t>(g0));
        in = out;
      }
      if (rc != PENEO_OK) break;
      rc = launch_gemm_tc2(in, d, reinterpret_cast<const __nv_bfloat16*>(pk + L.g_out_w[h]), d, nullptr, z, 32, rows, 32, d, 1,
                           1, st, 0);
      if (rc != PENEO_OK) break;
      const int C = head_classes(h);
      extract_logits_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, st>>>(
          z, reinterpret_cast<const float*>(pk + L.g_out_b[h]), C, rows, logits[h] + g0 * C);
      if (cudaGetLastError() != cudaSuccess) rc = PENEO_E_CUDA;
    }
  }
  const cudaError_t fe = cudaFreeAsync(ws, st);
  if (rc == PENEO_E_CUDA) set_error("pair_heads_generic: kernel launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
  if (rc == PENEO_OK && fe != cudaSuccess) {
    set_error("pair_heads_generic: cudaFreeAsync failed (%s)", cudaGetErrorString(fe));
    rc = PENEO_E_CUDA;
  }
  return rc;
}

}  // namespace peneo
