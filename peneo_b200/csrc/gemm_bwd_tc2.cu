// The two large GEMMs of the bf16 backward pass on CTA pairs (cta_group::2, M = 256) — same mathematics, operand
// layouts and epilogues as gemm_bwd_tc.cu, but two CTAs of a cluster run every UMMA together: each owns one
// [128 x 384] accumulator tile (its own A rows) and supplies HALF of the B operand from its shared memory.
//
//   dS  [rows, 384]  = G [rows, 1920] * W_mid [1920, 384]        pair = two consecutive 128-row blocks of G
//   dWm [1920, 384] += G^T [1920, rows] * S [rows, 384]          pair = two consecutive 128-feature blocks, same K split
//
// Why: in the single-CTA kernel a 64-deep K block moves 64 KB through the shared-memory port twice (TMA write +
// tensor-core read) per 768 tensor cycles = 1024 wavefronts of 128 B, more than the port delivers (measured: tensor
// pipe 64 / 69 % active).  With the B operand (48 of the 64 KB) split over the pair, a CTA moves 40 KB per K block:
// 640 wavefronts per 768 cycles.
//
// Synchronisation as in pair_heads_tc2.cu: only the leader issues MMAs and waits on ITS "stage full" barrier (TMA
// bytes of both CTAs are counted there); everything the MMA releases is a tcgen05.commit multicast to both CTAs.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {
namespace gb2 {

constexpr int kN = 384;
constexpr int kBlk = 64 * 128;                    // one [64 rows x 64 cols] swizzled box = 8 KB
constexpr int kStageBytes = 2 * kBlk + 3 * kBlk;  // A 16 KB + this CTA's half of B 24 KB
constexpr int kStages = 5;
constexpr int kMiscOff = kStages * kStageBytes;   // dW: constant ones box ; dS: store staging (8 KB)
constexpr int kSmemBytes = kStages * kStageBytes + kBlk + 1024 + 256;

struct Args {
  int64_t m_total;   // rows of the output (dS: pairs in the chunk; dW: 1920)
  int32_t k_total;   // contraction length (dS: 1920; dW: pairs in the chunk)
  int32_t kb_per_split, splits;
  int64_t num_pair_items;  // (pairs of m blocks) * splits
  float* out[kNumHeads];   // dS: out[0] = dS (bf16, ld 384); dW: five [384, 384] matrices, m block -> head = mb / 3
  float* colsum[kNumHeads];  // dW only: five [384] vectors += column sums of A (db_mid)
};

template <bool A_MN>
__global__ void __launch_bounds__(192, 1)
    gemm_bwd_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kBlk);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;
  uint64_t* acc_empty = bars + 2 * kStages + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int num_kb = (a.k_total + 63) / 64;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < kStages; ++s) ptx::mbar_init(&full[s], 2), ptx::mbar_init(&empty[s], 1);
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, 8);  // 4 epilogue warps of each CTA
    ptx::fence_barrier_init();
  }
  if (A_MN) {
    // ones box in the MN-major SWIZZLE_128B layout: k-row r at r * 128 B, logical 16-byte chunk 0 at chunk (r & 7)
    uint4* ones = reinterpret_cast<uint4*>(smem + kMiscOff);
    for (int e = threadIdx.x; e < kBlk / 16; e += blockDim.x) {
      const int r = e / 8, ch = e % 8;
      ones[e] = make_uint4(ch == (r & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);  // bf16 1.0 in element 0
    }
    ptx::fence_proxy_async();
  }
  ptx::cluster_sync_all();  // both CTAs' barriers exist before the peer's TMA / remote arrives touch them
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  auto arrive_leader = [&](uint64_t* bar) {
    if (leader) ptx::mbar_arrive(bar);
    else ptx::mbar_arrive_remote(bar, 0);
  };

  const int64_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  // pair item -> (this CTA's m block, first K block, number of K blocks)
  auto decode = [&](int64_t item, int64_t& mb, int& kb0, int& nk) {
    const int split = static_cast<int>(item % a.splits);
    mb = 2 * (item / a.splits) + rank;
    kb0 = split * a.kb_per_split;
    nk = min(a.kb_per_split, num_kb - kb0);
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t item = cid; item < a.num_pair_items; item += ncl) {
        int64_t mb;
        int kb0, nk;
        decode(item, mb, kb0, nk);
        for (int kb = 0; kb < nk; ++kb) {
          const int k0 = (kb0 + kb) * 64;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          unsigned char* st = smem + s * kStageBytes;
          if (A_MN) {  // two [64 k-rows x 64 m-cols] boxes of this CTA's 128 features
            ptx::tma_load_2d_2sm(st, &tmA, &full[s], static_cast<int32_t>(mb * 128), k0);
            ptx::tma_load_2d_2sm(st + kBlk, &tmA, &full[s], static_cast<int32_t>(mb * 128 + 64), k0);
          } else {     // one [128 m-rows x 64 k-cols] box of this CTA's 128 rows
            ptx::tma_load_2d_2sm(st, &tmA, &full[s], k0, static_cast<int32_t>(mb * 128));
          }
          // this CTA's half of B: columns [128 r, 128 r + 128) for the N = 256 MMA, [256 + 64 r, + 64) for the N = 128 one
          ptx::tma_load_2d_2sm(st + 2 * kBlk, &tmB, &full[s], static_cast<int32_t>(128 * rank), k0);
          ptx::tma_load_2d_2sm(st + 3 * kBlk, &tmB, &full[s], static_cast<int32_t>(128 * rank + 64), k0);
          ptx::tma_load_2d_2sm(st + 4 * kBlk, &tmB, &full[s], static_cast<int32_t>(256 + 64 * rank), k0);
          if (leader) ptx::mbar_arrive_expect_tx(&full[s], 2 * kStageBytes);
          else ptx::mbar_arrive_remote(&full[s], 0);
          if (++s == kStages) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (leader && ptx::elect_one()) {
      constexpr uint32_t idesc256 = ptx::umma_idesc_bf16_major(256, 256, A_MN, true);
      constexpr uint32_t idesc128 = ptx::umma_idesc_bf16_major(256, 128, A_MN, true);
      constexpr uint32_t idesc16 = ptx::umma_idesc_bf16_major(256, 16, A_MN, true);
      const uint32_t ones_addr = ptx::smem_u32(smem + kMiscOff);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int64_t item = cid; item < a.num_pair_items; item += ncl, ++it) {
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
            ptx::umma_ss_2sm(tmem, ad, ptx::umma_desc_mn_sw128(b_addr + ks * 2048, kBlk, 1024), idesc256, (kb | ks) != 0);
            ptx::umma_ss_2sm(tmem + 256, ad, ptx::umma_desc_mn_sw128(b_addr + 2 * kBlk + ks * 2048, kBlk, 1024), idesc128,
                             (kb | ks) != 0);
            if (A_MN)
              ptx::umma_ss_2sm(tmem + kN, ad, ptx::umma_desc_mn_sw128(ones_addr + ks * 2048, kBlk, 1024), idesc16,
                               (kb | ks) != 0);
          }
          ptx::tc_commit_2sm(&empty[s], 3);
          if (++s == kStages) s = 0, ph ^= 1;
        }
        ptx::tc_commit_2sm(acc_full, 3);
      }
    }
  } else {
    const int q = warp % 4;
    int it = 0;
    for (int64_t item = cid; item < a.num_pair_items; item += ncl, ++it) {
      int64_t mb;
      int kb0, nk;
      decode(item, mb, kb0, nk);
      ptx::mbar_wait(acc_full, it & 1);
      ptx::tc_fence_after();
      const int64_t m = mb * 128 + q * 32 + lane;
      const bool live = m < a.m_total;
      float* dst = nullptr;
      if (A_MN && live) dst = a.out[static_cast<int>(m / kN)] + (m % kN) * kN;  // row m of the stacked [1920, 384] gradient
      unsigned char* ob = smem + kMiscOff + q * 2048;
      __nv_bfloat16* ds16 = reinterpret_cast<__nv_bfloat16*>(a.out[0]);
      const int64_t wrow0 = mb * 128 + q * 32;
      if (A_MN) {
        uint32_t c4[4];
        ptx::tmem_ld_x4(tmem + (static_cast<uint32_t>(q * 32) << 16) + kN, c4);
        ptx::tmem_ld_wait();
        if (live) atomicAdd(a.colsum[static_cast<int>(m / kN)] + (m % kN), __uint_as_float(c4[0]));
      }
#pragma unroll 1
      for (int piece = 0; piece < kN / 32; ++piece) {
        uint32_t r[32];
        ptx::tmem_ld_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + piece * 32, r);
        ptx::tmem_ld_wait();
        if (piece == kN / 32 - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(acc_empty);
        }
        if (A_MN) {
          if (live) {
#pragma unroll
            for (int x = 0; x < 32; x += 4)
              atomicAdd(reinterpret_cast<float4*>(dst + piece * 32 + x),
                        make_float4(__uint_as_float(r[x]), __uint_as_float(r[x + 1]), __uint_as_float(r[x + 2]),
                                    __uint_as_float(r[x + 3])));
          }
        } else {
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
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc_2sm(tmem, 512);
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

template <bool A_MN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& a, cudaStream_t st) {
  const int clusters = static_cast<int>(std::min<int64_t>(a.num_pair_items, sm_count() / 2));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters), cfg.blockDim = dim3(192), cfg.dynamicSmemBytes = kSmemBytes, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  PENEO_CUDA_TRY(cudaFuncSetAttribute(gemm_bwd_pair_kernel<A_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  PENEO_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_bwd_pair_kernel<A_MN>, tmA, tmB, a));
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace gb2

// dS[rows, 384] (bf16) = G[rows, 1920] * Wmid[1920, 384]   (Wmid: the five [d_out, d_in] matrices stacked, bf16, unscaled)
int launch_gemm_ds_pair(const __nv_bfloat16* G, const __nv_bfloat16* wmid_full, __nv_bfloat16* dS, int rows, cudaStream_t st) {
  using namespace gb2;
  if (rows == 0) return PENEO_OK;
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA, G, 5 * kN, rows, 5 * kN * 2, 64, 128)) != PENEO_OK) return rc;
  if ((rc = make_tensor_map_bf16(&tmB, wmid_full, kN, 5 * kN, kN * 2, 64, 64)) != PENEO_OK) return rc;
  Args a{};
  a.m_total = rows, a.k_total = 5 * kN;
  a.splits = 1, a.kb_per_split = (a.k_total + 63) / 64;
  a.num_pair_items = ((rows + 127) / 128 + 1) / 2;
  a.out[0] = reinterpret_cast<float*>(dS);
  return launch<false>(tmA, tmB, a, st);
}

// dWmid[h][384, 384] += (G[:, 384 h : 384 (h + 1)])^T * S ;  dbmid[h][384] += column sums of G[:, 384 h : 384 (h + 1)]
int launch_gemm_dw_pair(const __nv_bfloat16* G, const __nv_bfloat16* S, float* const dW[kNumHeads],
                        float* const db[kNumHeads], int rows, cudaStream_t st) {
  using namespace gb2;
  if (rows == 0) return PENEO_OK;
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA, G, 5 * kN, rows, 5 * kN * 2, 64, 64)) != PENEO_OK) return rc;
  if ((rc = make_tensor_map_bf16(&tmB, S, kN, rows, kN * 2, 64, 64)) != PENEO_OK) return rc;
  Args a{};
  a.m_total = 5 * kN, a.k_total = rows;
  const int num_kb = (rows + 63) / 64, m_pairs = (5 * kN / 128 + 1) / 2;  // 15 feature blocks -> 8 pairs (the last one half empty)
  int splits = std::max(1, std::min(num_kb, (sm_count() / 2) / m_pairs));  // one pair item per cluster, a single wave
  a.kb_per_split = (num_kb + splits - 1) / splits;
  a.splits = (num_kb + a.kb_per_split - 1) / a.kb_per_split;
  a.num_pair_items = (int64_t)m_pairs * a.splits;
  for (int h = 0; h < kNumHeads; ++h) a.out[h] = dW[h], a.colsum[h] = db[h];
  return launch<true>(tmA, tmB, a, st);
}

}  // namespace peneo
