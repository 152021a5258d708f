// The two large GEMMs of the bf16 backward pass, on tcgen05 with a [128 x 384] fp32 accumulator tile in TMEM.
// Both read the row-major bf16 matrices written by T1 (pair_bwd_tc.cu) directly — no transposed copies —
// by using MN-major shared-memory descriptors where the contraction index is the ROW index in memory:
//
//   dS  [rows, 384]  = G [rows, 1920] * W_mid [1920, 384]        A K-major (row = m, contiguous = k = mid feature)
//       (stored as bf16: its only consumer is the dA / dBm reduction, a sum over up to N terms)
//                                                                B MN-major (row = k = mid feature, contiguous = n)
//   dWm [1920, 384] += G^T [1920, rows] * S [rows, 384]          A MN-major (row = k = pair, contiguous = m)
//                                                                B MN-major (row = k = pair, contiguous = n)
//
// (autograd of the five Linear(D, D) layers in model/peneo_decoder.py:258-269: grad_input = grad_output W,
// grad_weight = grad_output^T input.)  Stage = one 64-deep K block: A 16 KB + B 6 x 8 KB = 64 KB, 3 stages.
// Persistent CTAs: warp 0 TMA, warp 1 MMA (+ TMEM alloc), warps 2-5 epilogue.  The dW GEMM is split along K
// (= pairs) across CTAs and accumulated with fp32 atomics.  It also produces db_mid = column sums of G for free: a
// constant B box whose column 0 is all ones gives D[:, 384] = G^T 1 (one extra N = 16 MMA per K step).
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {
namespace gb {

constexpr int kN = 384;
constexpr int kBlk = 64 * 128;              // one [64 rows x 64 cols] swizzled box = 8 KB
constexpr int kStageBytes = 2 * kBlk + 6 * kBlk;  // 64 KB
constexpr int kStages = 3;
constexpr int kOnesOff = kStages * kStageBytes;  // dW: constant [64 k x 64 n] box, column 0 = 1 ; dS: store staging
constexpr int kSmemBytes = kStages * kStageBytes + kBlk + 1024 + 256;

struct Args {
  int64_t m_total;   // rows of the output (dS: pairs in the chunk; dW: 1920)
  int32_t k_total;   // contraction length (dS: 1920; dW: pairs in the chunk)
  int32_t kb_per_split, splits;
  int64_t num_items;  // m blocks * splits
  float* out[kNumHeads];  // dS: out[0] = dS (bf16, ld 384); dW: five [384, 384] matrices, m block -> head = mb / 3
  float* colsum[kNumHeads];  // dW only: five [384] vectors += column sums of A (db_mid)
};

template <bool A_MN>
__global__ void __launch_bounds__(192, 1)
    gemm_bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // by offset: keeps the shared address space
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kBlk);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;
  uint64_t* acc_empty = bars + 2 * kStages + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int num_kb = (a.k_total + 63) / 64;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < kStages; ++s) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, 4);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (A_MN) {
    // ones box in the MN-major SWIZZLE_128B layout: k-row r at r * 128 B, logical 16-byte chunk 0 at chunk (r & 7)
    uint4* ones = reinterpret_cast<uint4*>(smem + kOnesOff);
    for (int e = threadIdx.x; e < kBlk / 16; e += blockDim.x) {
      const int r = e / 8, ch = e % 8;
      ones[e] = make_uint4(ch == (r & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);  // bf16 1.0 in element 0
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  auto decode = [&](int64_t item, int64_t& mb, int& kb0, int& nk) {
    const int split = static_cast<int>(item % a.splits);
    mb = item / a.splits;
    kb0 = split * a.kb_per_split;
    nk = min(a.kb_per_split, num_kb - kb0);
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t item = blockIdx.x; item < a.num_items; item += gridDim.x) {
        int64_t mb;
        int kb0, nk;
        decode(item, mb, kb0, nk);
        for (int kb = 0; kb < nk; ++kb) {
          const int k0 = (kb0 + kb) * 64;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
          unsigned char* st = smem + s * kStageBytes;
          if (A_MN) {  // two [64 k-rows x 64 m-cols] boxes
            ptx::tma_load_2d(st, &tmA, &full[s], static_cast<int32_t>(mb * 128), k0);
            ptx::tma_load_2d(st + kBlk, &tmA, &full[s], static_cast<int32_t>(mb * 128 + 64), k0);
          } else {     // one [128 m-rows x 64 k-cols] box
            ptx::tma_load_2d(st, &tmA, &full[s], k0, static_cast<int32_t>(mb * 128));
          }
          for (int nb = 0; nb < 6; ++nb) ptx::tma_load_2d(st + (2 + nb) * kBlk, &tmB, &full[s], nb * 64, k0);
          if (++s == kStages) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc256 = ptx::umma_idesc_bf16_major(128, 256, A_MN, true);
      constexpr uint32_t idesc128 = ptx::umma_idesc_bf16_major(128, 128, A_MN, true);
      constexpr uint32_t idesc16 = ptx::umma_idesc_bf16_major(128, 16, A_MN, true);
      const uint32_t ones_addr = ptx::smem_u32(smem + kOnesOff);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int64_t item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
        int64_t mb;
        int kb0, nk;
        decode(item, mb, kb0, nk);
        ptx::mbar_wait(acc_empty, (it & 1) ^ 1);
        ptx::tc_fence_after();
        for (int kb = 0; kb < nk; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + s * kStageBytes);
          const uint32_t b_addr = a_addr + 2 * kBlk;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = A_MN ? ptx::umma_desc_mn_sw128(a_addr + ks * 2048, kBlk, 1024)
                                     : ptx::umma_desc_sw128(a_addr + ks * 32);
            ptx::umma_ss(tmem, ad, ptx::umma_desc_mn_sw128(b_addr + ks * 2048, kBlk, 1024), idesc256, (kb | ks) != 0);
            ptx::umma_ss(tmem + 256, ad, ptx::umma_desc_mn_sw128(b_addr + 4 * kBlk + ks * 2048, kBlk, 1024), idesc128,
                         (kb | ks) != 0);
            if (A_MN)
              ptx::umma_ss(tmem + kN, ad, ptx::umma_desc_mn_sw128(ones_addr + ks * 2048, kBlk, 1024), idesc16, (kb | ks) != 0);
          }
          ptx::tc_commit(&empty[s]);
          if (++s == kStages) s = 0, ph ^= 1;
        }
        ptx::tc_commit(acc_full);
      }
    }
  } else {
    const int q = warp % 4;
    int it = 0;
    for (int64_t item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
      int64_t mb;
      int kb0, nk;
      decode(item, mb, kb0, nk);
      ptx::mbar_wait(acc_full, it & 1);
      ptx::tc_fence_after();
      const int64_t m = mb * 128 + q * 32 + lane;
      float* dst = nullptr;
      if (A_MN) dst = a.out[static_cast<int>(m / kN)] + (m % kN) * kN;  // row m of the stacked [1920, 384] gradient
      // dS: bf16 rows through a per-warp smem tile ([32 rows][64 B], 16-byte chunks XOR-swizzled by (row >> 1) & 3) so
      // that one store instruction covers 8 rows x 64 B instead of 32 rows x 16 B
      unsigned char* ob = smem + kOnesOff + q * 2048;
      __nv_bfloat16* ds16 = reinterpret_cast<__nv_bfloat16*>(a.out[0]);
      const int64_t wrow0 = mb * 128 + q * 32;
      if (A_MN) {
        uint32_t c4[4];
        ptx::tmem_ld_x4(tmem + (static_cast<uint32_t>(q * 32) << 16) + kN, c4);
        ptx::tmem_ld_wait();
        if (m < a.m_total) atomicAdd(a.colsum[static_cast<int>(m / kN)] + (m % kN), __uint_as_float(c4[0]));
      }
#pragma unroll 1
      for (int piece = 0; piece < kN / 32; ++piece) {
        uint32_t r[32];
        ptx::tmem_ld_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + piece * 32, r);
        ptx::tmem_ld_wait();
        if (piece == kN / 32 - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(acc_empty);
        }
        if (m < a.m_total) {
          if (A_MN) {
#pragma unroll
            for (int x = 0; x < 32; x += 4)
              atomicAdd(reinterpret_cast<float4*>(dst + piece * 32 + x),
                        make_float4(__uint_as_float(r[x]), __uint_as_float(r[x + 1]), __uint_as_float(r[x + 2]),
                                    __uint_as_float(r[x + 3])));
          }
        }
        if (!A_MN) {
          __syncwarp();  // the read-back of the previous piece is complete
#pragma unroll
          for (int v = 0; v < 4; ++v)
            *reinterpret_cast<uint4*>(ob + lane * 64 + ((v ^ ((lane >> 1) & 3)) * 16)) = make_uint4(
                ptx::pack_bf16x2(__uint_as_float(r[8 * v]), __uint_as_float(r[8 * v + 1])),
                ptx::pack_bf16x2(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3])),
                ptx::pack_bf16x2(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5])),
                ptx::pack_bf16x2(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7])));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + (lane >> 2), c16 = lane & 3;
            const uint4 val = *reinterpret_cast<const uint4*>(ob + rr * 64 + ((c16 ^ ((rr >> 1) & 3)) * 16));
            if (wrow0 + rr < a.m_total) *reinterpret_cast<uint4*>(ds16 + (wrow0 + rr) * kN + piece * 32 + c16 * 8) = val;
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sms = 148;
  }
  return sms;
}

}  // namespace gb

// dS[rows, 384] (bf16) = G[rows, 1920] * Wmid[1920, 384]   (Wmid: the five [d_out, d_in] matrices stacked, bf16, unscaled)
int launch_gemm_ds(const __nv_bfloat16* G, const __nv_bfloat16* wmid_full, __nv_bfloat16* dS, int rows, cudaStream_t st) {
  using namespace gb;
  if (rows == 0) return PENEO_OK;
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA, G, 5 * kN, rows, 5 * kN * 2, 64, 128)) != PENEO_OK) return rc;
  if ((rc = make_tensor_map_bf16(&tmB, wmid_full, kN, 5 * kN, kN * 2, 64, 64)) != PENEO_OK) return rc;
  Args a{};
  a.m_total = rows, a.k_total = 5 * kN;
  a.splits = 1, a.kb_per_split = (a.k_total + 63) / 64;
  a.num_items = (rows + 127) / 128;
  a.out[0] = reinterpret_cast<float*>(dS);
  const int grid = static_cast<int>(std::min<int64_t>(a.num_items, sm_count()));
  PENEO_CUDA_TRY(cudaFuncSetAttribute(gemm_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  gemm_bwd_kernel<false><<<grid, 192, kSmemBytes, st>>>(tmA, tmB, a);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

// dWmid[h][384, 384] += (G[:, 384 h : 384 (h + 1)])^T * S        (G [rows, 1920], S [rows, 384], bf16)
// dbmid[h][384]     += column sums of G[:, 384 h : 384 (h + 1)]
int launch_gemm_dw(const __nv_bfloat16* G, const __nv_bfloat16* S, float* const dW[kNumHeads], float* const db[kNumHeads],
                   int rows, cudaStream_t st) {
  using namespace gb;
  if (rows == 0) return PENEO_OK;
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA, G, 5 * kN, rows, 5 * kN * 2, 64, 64)) != PENEO_OK) return rc;
  if ((rc = make_tensor_map_bf16(&tmB, S, kN, rows, kN * 2, 64, 64)) != PENEO_OK) return rc;
  Args a{};
  a.m_total = 5 * kN, a.k_total = rows;
  const int num_kb = (rows + 63) / 64, m_blocks = 5 * kN / 128;
  int splits = std::max(1, std::min(num_kb, sm_count() / m_blocks));  // one work item per SM, a single wave
  a.kb_per_split = (num_kb + splits - 1) / splits;
  a.splits = (num_kb + a.kb_per_split - 1) / a.kb_per_split;
  a.num_items = (int64_t)m_blocks * a.splits;
  for (int h = 0; h < kNumHeads; ++h) a.out[h] = dW[h], a.colsum[h] = db[h];
  const int grid = static_cast<int>(std::min<int64_t>(a.num_items, sm_count()));
  PENEO_CUDA_TRY(cudaFuncSetAttribute(gemm_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  gemm_bwd_kernel<true><<<grid, 192, kSmemBytes, st>>>(tmA, tmB, a);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
