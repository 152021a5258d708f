// tcgen05 GEMM for the per-token projections (K1):
//   C[M, N] = act(A[M, K] * W[N, K]^T + bias[N])      A, W bf16 (K contiguous), fp32 accumulate in
//   TMEM, bf16 output.  128x128 output tile per CTA, K in 64-wide TMA stages (SWIZZLE_128B),
//   warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue.
// Replaces the nn.Linear calls of shrink_projection and the per-token half of combine_fc
// (model/peneo_decoder.py:215-222, 126).
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {

// ------------------------------------------------------------------------------------------------
// tensor maps (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

static int make_tensor_map_any(void* tmap_out, CUtensorMapDataType dtype, const void* gptr, uint64_t inner, uint64_t outer,
                               uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                               CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B);

int make_tensor_map_bf16(void* tmap_out, const void* gptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                         uint32_t box_inner, uint32_t box_outer) {
  return make_tensor_map_any(tmap_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, gptr, inner, outer, row_stride_bytes, box_inner,
                             box_outer);
}
// 64-byte swizzle (box rows of 32 bf16): the per-warp store tiles of T1
int make_tensor_map_bf16_sw64(void* tmap_out, const void* gptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                              uint32_t box_inner, uint32_t box_outer) {
  return make_tensor_map_any(tmap_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, gptr, inner, outer, row_stride_bytes, box_inner,
                             box_outer, CU_TENSOR_MAP_SWIZZLE_64B);
}
int make_tensor_map_f32(void* tmap_out, const void* gptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                        uint32_t box_inner, uint32_t box_outer) {
  return make_tensor_map_any(tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, gptr, inner, outer, row_stride_bytes, box_inner,
                             box_outer);
}

static int make_tensor_map_any(void* tmap_out, CUtensorMapDataType dtype, const void* gptr, uint64_t inner, uint64_t outer,
                               uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return PENEO_E_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(static_cast<CUtensorMap*>(tmap_out), dtype, 2, const_cast<void*>(gptr),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
              box_outer);
    return PENEO_E_CUDA;
  }
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
constexpr int kGemmStages = 4;
constexpr int kGemmStageBytes = 2 * 128 * 64 * 2;  // A tile + W tile
constexpr int kGemmSmem = kGemmStages * kGemmStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

template <int ACT>
__global__ void __launch_bounds__(192, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ C, int64_t ldc, int64_t M, int N,
                   int K) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // by offset: keeps the shared address space
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGemmStages * kGemmStageBytes);
  uint64_t* full = bars;                  // [stages]
  uint64_t* empty = bars + kGemmStages;   // [stages]
  uint64_t* acc_full = bars + 2 * kGemmStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGemmStages + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n0 = blockIdx.x * 128;
  const int64_t m0 = (int64_t)blockIdx.y * 128;
  const int num_k = K / 64;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < kGemmStages; ++s) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (ptx::elect_one()) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % kGemmStages;
        const uint32_t ph = (kb / kGemmStages) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&full[s], kGemmStageBytes);
        unsigned char* st = smem + s * kGemmStageBytes;
        ptx::tma_load_2d(st, &tmA, &full[s], kb * 64, static_cast<int32_t>(m0));
        ptx::tma_load_2d(st + 128 * 64 * 2, &tmW, &full[s], kb * 64, n0);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 128);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % kGemmStages;
        const uint32_t ph = (kb / kGemmStages) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + s * kGemmStageBytes);
        const uint32_t w_addr = a_addr + 128 * 64 * 2;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          ptx::umma_ss(tmem, ptx::umma_desc_sw128(a_addr + ks * 32), ptx::umma_desc_sw128(w_addr + ks * 32), idesc,
                       (kb | ks) != 0);
        ptx::tc_commit(&empty[s]);
      }
      ptx::tc_commit(acc_full);
    }
  } else {
    const int q = warp % 4;  // TMEM lane quarter this warp may access
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    const int64_t m = m0 + q * 32 + lane;
#pragma unroll 1
    for (int piece = 0; piece < 4; ++piece) {
      uint32_t r[32];
      ptx::tmem_ld_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + piece * 32, r);
      ptx::tmem_ld_wait();
      if (m < M) {
        uint32_t packed[16];
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          float v0 = __uint_as_float(r[x]) + bias[n0 + piece * 32 + x];
          float v1 = __uint_as_float(r[x + 1]) + bias[n0 + piece * 32 + x + 1];
          if (ACT == 1) v0 = v0 / (1.f + __expf(-v0)), v1 = v1 / (1.f + __expf(-v1));
          packed[x / 2] = ptx::pack_bf16x2(v0, v1);
        }
        uint4* dst = reinterpret_cast<uint4*>(C + m * ldc + n0 + piece * 32);
#pragma unroll
        for (int v = 0; v < 4; ++v) dst[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 128);
}

int launch_gemm_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias,
                   __nv_bfloat16* C, int64_t ldc, int64_t M, int N, int K, int act, cudaStream_t st) {
  PENEO_REQUIRE(N % 128 == 0 && K % 64 == 0 && K >= 64, "gemm_tc: N %% 128 and K %% 64 required (N=%d K=%d)", N, K);
  PENEO_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "gemm_tc: leading dimensions must be multiples of 8");
  if (M == 0) return PENEO_OK;
  alignas(64) CUtensorMap tmA, tmW;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA, A, K, M, lda * 2, 64, 128)) != PENEO_OK) return rc;
  if ((rc = make_tensor_map_bf16(&tmW, W, K, N, ldw * 2, 64, 128)) != PENEO_OK) return rc;
  dim3 grid(N / 128, static_cast<unsigned>((M + 127) / 128));
  if (act) {
    PENEO_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    gemm_tc_kernel<1><<<grid, 192, kGemmSmem, st>>>(tmA, tmW, bias, C, ldc, M, N, K);
  } else {
    PENEO_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    gemm_tc_kernel<0><<<grid, 192, kGemmSmem, st>>>(tmA, tmW, bias, C, ldc, M, N, K);
  }
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
// gemm_tc2: persistent 128x128-tile tcgen05 GEMM with the output modes the backward pass needs.
//   C[M, N] (op)= A[M, K] W[N, K]^T      A, W bf16, K contiguous (TMA zero-fills the K / M / N tails)
//   OUT 0: bf16 store of (acc + bias)      OUT 1: fp32 store      OUT 2: fp32 C += acc
//   OUT 3: fp32 atomicAdd (split-K; C pre-initialised by the caller)
// One CTA per SM loops over work items (m block, n block, k split); the accumulator is double-buffered in
// TMEM so the epilogue of item i overlaps the main loop of item i + 1 — what makes the short-K (K = 384)
// GEMMs of the backward pass efficient.  warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-5 epilogue.
// ------------------------------------------------------------------------------------------------
constexpr int kG2Stages = 5;
constexpr int kG2Smem = kG2Stages * kGemmStageBytes + 1024 + 256;

// ACT: 0 = linear, 1 = SiLU, 2 = SiLU + dropout (bf16 output only).  Compile-time so that the unrolled epilogue has no
// per-element branches; SiLU is h + h tanh(h), h = v / 2 (one MUFU), as in K2.
template <int OUT, int ACT>
__global__ void __launch_bounds__(192, 1)
    gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const float* __restrict__ bias, void* __restrict__ Cv, int64_t ldc, int64_t M, int N, int K,
                    int kb_per_split, int splits, int n_blocks, int64_t num_items, int act, uint32_t drop_thresh,
                    float drop_scale, uint32_t drop_key, uint32_t drop_row0) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // by offset: keeps the shared address space
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kG2Stages * kGemmStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kG2Stages;
  uint64_t* acc_full = bars + 2 * kG2Stages;       // [2]
  uint64_t* acc_empty = bars + 2 * kG2Stages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kG2Stages + 4);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int num_k_total = (K + 63) / 64;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < kG2Stages; ++s) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
    for (int s = 0; s < 2; ++s) ptx::mbar_init(&acc_full[s], 1), ptx::mbar_init(&acc_empty[s], 4);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  auto decode = [&](int64_t item, int64_t& m0, int& n0, int& kb0, int& nk) {
    const int split = static_cast<int>(item % splits);
    const int64_t tile = item / splits;
    n0 = static_cast<int>(tile % n_blocks) * 128;
    m0 = (tile / n_blocks) * 128;
    kb0 = split * kb_per_split;
    nk = min(kb_per_split, num_k_total - kb0);
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t item = blockIdx.x; item < num_items; item += gridDim.x) {
        int64_t m0;
        int n0, kb0, nk;
        decode(item, m0, n0, kb0, nk);
        for (int kb = 0; kb < nk; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&full[s], kGemmStageBytes);
          unsigned char* st = smem + s * kGemmStageBytes;
          ptx::tma_load_2d(st, &tmA, &full[s], (kb0 + kb) * 64, static_cast<int32_t>(m0));
          ptx::tma_load_2d(st + 128 * 64 * 2, &tmW, &full[s], (kb0 + kb) * 64, n0);
          if (++s == kG2Stages) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 128);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int64_t item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        int64_t m0;
        int n0, kb0, nk;
        decode(item, m0, n0, kb0, nk);
        const int buf = it & 1;
        ptx::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t acc = tmem + 128 * buf;
        for (int kb = 0; kb < nk; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + s * kGemmStageBytes);
          const uint32_t w_addr = a_addr + 128 * 64 * 2;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_ss(acc, ptx::umma_desc_sw128(a_addr + ks * 32), ptx::umma_desc_sw128(w_addr + ks * 32), idesc,
                         (kb | ks) != 0);
          ptx::tc_commit(&empty[s]);
          if (++s == kG2Stages) s = 0, ph ^= 1;
        }
        ptx::tc_commit(&acc_full[buf]);
      }
    }
  } else {
    const int q = warp % 4;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int64_t item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      int64_t m0;
      int n0, kb0, nk;
      decode(item, m0, n0, kb0, nk);
      const int buf = it & 1;
      ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      const int64_t m = m0 + q * 32 + lane;
#pragma unroll 1
      for (int piece = 0; piece < 4; ++piece) {
        uint32_t r[32];
        ptx::tmem_ld_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + 128 * buf + piece * 32, r);
        ptx::tmem_ld_wait();
        if (piece == 3) {  // accumulator drained: hand the buffer back before the (slow) global stores
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
        }
        const int nb = n0 + piece * 32;
        if (m < M && nb < N) {
          if (OUT == 0) {
            uint32_t packed[16];
            float bv[32];
            if (bias) {  // 16-byte aligned (checked by the launcher), nb is a multiple of 32
#pragma unroll
              for (int x = 0; x < 32; x += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + nb + x));
                bv[x] = b4.x, bv[x + 1] = b4.y, bv[x + 2] = b4.z, bv[x + 3] = b4.w;
              }
            } else {
#pragma unroll
              for (int x = 0; x < 32; ++x) bv[x] = 0.f;
            }
#pragma unroll
            for (int x = 0; x < 32; x += 2) {
              float v[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                v[e] = __uint_as_float(r[x + e]) + bv[x + e];
                if (ACT >= 1) {
                  const float h = 0.5f * v[e];
                  v[e] = fmaf(h, ptx::tanh_approx(h), h);
                }
                if (ACT == 2) v[e] = drop_keep(drop_key, drop_thresh, drop_row0 + static_cast<uint32_t>(m), nb + x + e) ? v[e] * drop_scale : 0.f;
              }
              packed[x / 2] = ptx::pack_bf16x2(v[0], v[1]);
            }
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(Cv) + m * ldc + nb);
#pragma unroll
            for (int v = 0; v < 4; ++v)
              dst[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
          } else {
            float* dst = static_cast<float*>(Cv) + m * ldc + nb;
            if (OUT == 3) {
#pragma unroll
              for (int x = 0; x < 32; ++x) atomicAdd(dst + x, __uint_as_float(r[x]));
            } else {
#pragma unroll
              for (int x = 0; x < 32; x += 4) {
                float4 v = make_float4(__uint_as_float(r[x]), __uint_as_float(r[x + 1]), __uint_as_float(r[x + 2]),
                                       __uint_as_float(r[x + 3]));
                float4* d4 = reinterpret_cast<float4*>(dst + x);
                if (OUT == 2) {
                  const float4 o = *d4;
                  v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
                }
                *d4 = v;
              }
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 256);
}

int launch_gemm_tc2(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias, void* C,
                    int64_t ldc, int64_t M, int N, int K, int out_mode, int splits, cudaStream_t st, int act, const DropSpec* drop,
                    uint32_t site, uint32_t drop_row0) {
  PENEO_REQUIRE(N % 32 == 0 && K >= 1, "gemm_tc2: N %% 32 required (N=%d K=%d)", N, K);
  PENEO_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % (out_mode == 0 ? 8 : 4) == 0,
                "gemm_tc2: leading dimensions not vector aligned (lda=%lld ldw=%lld ldc=%lld)", (long long)lda, (long long)ldw,
                (long long)ldc);
  PENEO_REQUIRE(out_mode >= 0 && out_mode <= 3 && (out_mode == 3 || splits == 1), "gemm_tc2: bad output mode / splits");
  if (M == 0) return PENEO_OK;
  alignas(64) CUtensorMap tmA, tmW;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA, A, K, M, lda * 2, 64, 128)) != PENEO_OK) return rc;
  if ((rc = make_tensor_map_bf16(&tmW, W, K, N, ldw * 2, 64, 128)) != PENEO_OK) return rc;
  const int num_k = (K + 63) / 64;
  splits = std::max(1, std::min(splits, num_k));
  const int kb_per_split = (num_k + splits - 1) / splits;
  splits = (num_k + kb_per_split - 1) / kb_per_split;
  const int n_blocks = (N + 127) / 128;
  const int64_t items = ((M + 127) / 128) * n_blocks * splits;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    PENEO_CUDA_TRY(cudaGetDevice(&dev));
    PENEO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
 const int grid = static_cast<int>(std::min<int64_t>(items, sms));
  const uint32_t d_th = (drop && act) ? drop->thresh : 0u, d_key = drop ? drop_key(*drop, site) : 0u;
  const float d_sc = drop ? drop->scale : 1.f;
#define GO(OUT, ACT)                                                                                                 \
  {                                                                                                                  \
    static bool attr_set = false;                                                                                    \
    if (!attr_set) {                                                                                                 \
      PENEO_CUDA_TRY(                                                                                                \
          cudaFuncSetAttribute(gemm_tc2_kernel<OUT, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kG2Smem));    \
      attr_set = true;                                                                                               \
    }                                                                                                                \
    gemm_tc2_kernel<OUT, ACT><<<grid, 192, kG2Smem, st>>>(tmA, tmW, bias, C, ldc, M, N, K, kb_per_split, splits,      \
                                                          n_blocks, items, act, d_th, d_sc, d_key, drop_row0);       \
  }
  PENEO_REQUIRE(out_mode == 0 || !act, "gemm_tc2: the activation is fused for the bf16 output only");
  PENEO_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm_tc2: bias not 16-byte aligned");
  switch (out_mode) {
    case 0:
      if (!act) GO(0, 0) else if (!d_th) GO(0, 1) else GO(0, 2)
      break;
    case 1: GO(1, 0) break;
    case 2: GO(2, 0) break;
    default: GO(3, 0) break;
  }
#undef GO
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
