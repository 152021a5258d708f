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

int make_tensor_map_bf16(void* tmap_out, const void* gptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                         uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return PENEO_E_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(static_cast<CUtensorMap*>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
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
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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

}  // namespace peneo
