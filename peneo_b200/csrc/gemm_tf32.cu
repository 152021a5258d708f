// fp32 GEMMs of the per-token backward chain on the tensor cores (tcgen05 kind::tf32), bf16 mode only: the operands stay
// fp32 in memory (no converted copies) and are read as TF32 (10-bit mantissa) with fp32 accumulation — inside the
// 2e-2 budget of the bf16 mode, and ~50x the throughput of the CUDA-core SGEMM these GEMMs used before.
//   C[M, N] (op)= op(A) op(B) (+ bias[N]) ; optional second output C2 = Dropout(SiLU(C))
//   ta: A stored [K, M], else [M, K] ; tb: B stored [N, K], else [K, N]
// Every combination maps onto UMMA majorness flags: an operand whose contraction index is the ROW index in memory is
// MN-major ([32 k-rows x 32 elements] swizzled boxes, LBO = 4 KB between 32-element MN blocks, SBO = 1 KB between
// 8-row atoms), otherwise K-major (one [rows x 32 elements] box).  128 x 128 fp32 tile, BK = 32, persistent CTAs with a
// double-buffered TMEM accumulator, split-K with fp32 atomics (mode 2) for the weight gradients (K = tokens).
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {
namespace gt {

constexpr int kStageBytes = 2 * 128 * 32 * 4;  // A tile + B tile, 16 KB each
constexpr int kStages = 5;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
constexpr int kBlk = 32 * 128;  // one [32 rows x 32 fp32] box = 4 KB

struct Args {
  float* C;
  int64_t ldc;
  float* C2;
  int64_t ldc2;
  const float* bias;
  int M, N, K, mode;
  int kb_per_split, splits, n_blocks;
  int64_t num_items;
  uint32_t drop_thresh, drop_key, row0;
  float drop_scale;
};

// A_MN: the A operand is MN-major (A stored [K, M]); B_MN: B is MN-major (B stored [K, N])
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
    gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // by offset: keeps the shared address space
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;       // [2]
  uint64_t* acc_empty = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int num_kb = (a.K + 31) / 32;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < kStages; ++s) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
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

  auto decode = [&](int64_t item, int& m0, int& n0, int& kb0, int& nk) {
    const int split = static_cast<int>(item % a.splits);
    const int64_t tile = item / a.splits;
    n0 = static_cast<int>(tile % a.n_blocks) * 128;
    m0 = static_cast<int>(tile / a.n_blocks) * 128;
    kb0 = split * a.kb_per_split;
    nk = min(a.kb_per_split, num_kb - kb0);
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t item = blockIdx.x; item < a.num_items; item += gridDim.x) {
        int m0, n0, kb0, nk;
        decode(item, m0, n0, kb0, nk);
        for (int kb = 0; kb < nk; ++kb) {
          const int k0 = (kb0 + kb) * 32;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
          unsigned char* sa = smem + s * kStageBytes;
          unsigned char* sb = sa + 4 * kBlk;
          if (A_MN) {
            for (int mb = 0; mb < 4; ++mb) ptx::tma_load_2d(sa + mb * kBlk, &tmA, &full[s], m0 + 32 * mb, k0);
          } else {
            ptx::tma_load_2d(sa, &tmA, &full[s], k0, m0);
          }
          if (B_MN) {
            for (int nb = 0; nb < 4; ++nb) ptx::tma_load_2d(sb + nb * kBlk, &tmB, &full[s], n0 + 32 * nb, k0);
          } else {
            ptx::tma_load_2d(sb, &tmB, &full[s], k0, n0);
          }
          if (++s == kStages) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_tf32(128, 128, A_MN, B_MN);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int64_t item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
        int m0, n0, kb0, nk;
        decode(item, m0, n0, kb0, nk);
        const int buf = it & 1;
        ptx::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t acc = tmem + 128 * buf;
        for (int kb = 0; kb < nk; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + s * kStageBytes);
          const uint32_t b_addr = a_addr + 4 * kBlk;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // K = 8 per instruction: 32 B along a K-major row, one 8-row atom when MN-major
            const uint64_t ad = A_MN ? ptx::umma_desc_mn_sw128(a_addr + ks * 1024, kBlk, 1024) : ptx::umma_desc_sw128(a_addr + ks * 32);
            const uint64_t bd = B_MN ? ptx::umma_desc_mn_sw128(b_addr + ks * 1024, kBlk, 1024) : ptx::umma_desc_sw128(b_addr + ks * 32);
            ptx::umma_ss_tf32(acc, ad, bd, idesc, (kb | ks) != 0);
          }
          ptx::tc_commit(&empty[s]);
          if (++s == kStages) s = 0, ph ^= 1;
        }
        ptx::tc_commit(&acc_full[buf]);
      }
    }
  } else {
    const int q = warp % 4;
    const bool vec_ok = a.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(a.C) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0 &&
                        (a.C2 == nullptr || (a.ldc2 % 4 == 0 && (reinterpret_cast<uintptr_t>(a.C2) & 15) == 0));
    int it = 0;
    for (int64_t item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
      int m0, n0, kb0, nk;
      decode(item, m0, n0, kb0, nk);
      const int buf = it & 1;
      ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      const int m = m0 + q * 32 + lane;
#pragma unroll 1
      for (int piece = 0; piece < 4; ++piece) {
        uint32_t r[32];
        ptx::tmem_ld_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + 128 * buf + piece * 32, r);
        ptx::tmem_ld_wait();
        if (piece == 3) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
        }
        const int nb = n0 + piece * 32;
        if (m >= a.M) continue;
        float* dst = a.C + (int64_t)m * a.ldc + nb;
        if (vec_ok && nb + 32 <= a.N) {
          // fast path (every GEMM of the decoder: N % 32 == 0, 16-byte aligned rows): straight-line float4 code, the
          // mode / bias / C2 decisions are made once per 32 columns instead of once per element
          if (a.mode == 2) {
#pragma unroll
            for (int x = 0; x < 32; x += 4)
              atomicAdd(reinterpret_cast<float4*>(dst + x),
                        make_float4(__uint_as_float(r[x]), __uint_as_float(r[x + 1]), __uint_as_float(r[x + 2]),
                                    __uint_as_float(r[x + 3])));
            continue;
          }
          float v[32];
#pragma unroll
          for (int x = 0; x < 32; ++x) v[x] = __uint_as_float(r[x]);
          if (a.bias) {
#pragma unroll
            for (int x = 0; x < 32; x += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nb + x));
              v[x] += b4.x, v[x + 1] += b4.y, v[x + 2] += b4.z, v[x + 3] += b4.w;
            }
          }
          if (a.mode == 1) {
#pragma unroll
            for (int x = 0; x < 32; x += 4) {
              const float4 o = *reinterpret_cast<const float4*>(dst + x);
              v[x] += o.x, v[x + 1] += o.y, v[x + 2] += o.z, v[x + 3] += o.w;
            }
          }
#pragma unroll
          for (int x = 0; x < 32; x += 4) *reinterpret_cast<float4*>(dst + x) = make_float4(v[x], v[x + 1], v[x + 2], v[x + 3]);
          if (a.C2) {
            float* dst2 = a.C2 + (int64_t)m * a.ldc2 + nb;
            const bool drop = a.drop_thresh != 0;
#pragma unroll
            for (int x = 0; x < 32; x += 4) {
              float h[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float hh = 0.5f * v[x + e];
                h[e] = fmaf(hh, ptx::tanh_approx(hh), hh);  // SiLU with one MUFU, as everywhere in the bf16 mode
              }
              if (drop) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  h[e] = drop_keep(a.drop_key, a.drop_thresh, a.row0 + m, nb + x + e) ? h[e] * a.drop_scale : 0.f;
              }
              *reinterpret_cast<float4*>(dst2 + x) = make_float4(h[0], h[1], h[2], h[3]);
            }
          }
          continue;
        }
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          const int n = nb + x;
          if (n >= a.N) break;
          float v = __uint_as_float(r[x]);
          if (a.mode == 2) {
            atomicAdd(dst + x, v);
          } else {
            if (a.bias) v += a.bias[n];
            if (a.mode == 1) v += dst[x];
            dst[x] = v;
            if (a.C2) {
              float h = v / (1.0f + expf(-v));
              if (a.drop_thresh) h = drop_keep(a.drop_key, a.drop_thresh, a.row0 + m, n) ? h * a.drop_scale : 0.f;
              a.C2[(int64_t)m * a.ldc2 + n] = h;
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

}  // namespace gt

bool gemm_tf32_supported(const Tf32Gemm& g) {
  // TMA: 16-byte aligned bases and row strides; the contiguous dimension must cover whole 16-byte groups
  auto ok_ptr = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  // Only the K-major x K-major combination (A [M, K], B [N, K]: every forward nn.Linear) is enabled: the selftest shows
  // that MN-major operands of 4-byte elements need the 32-byte-base swizzle atom (SWIZZLE_128B_BASE32B), which the
  // boxes loaded here do not have — those combinations fall back to the exact CUDA-core SGEMM.
  if (g.ta || !g.tb) return false;
  return g.M > 0 && g.N > 0 && g.K > 0 && ok_ptr(g.A) && ok_ptr(g.B) && g.lda % 4 == 0 && g.ldb % 4 == 0 &&
         (int64_t)g.M * g.N >= 128 * 128;  // tiny outputs are not worth a tensor-core launch
}

int launch_gemm_tf32(const Tf32Gemm& g, cudaStream_t st) {
  using namespace gt;
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  // A stored [K, M] (ta): inner = M, outer = K, box 32 x 32 ; else [M, K]: inner = K, outer = M, box 32 x 128
  if (g.ta) rc = make_tensor_map_f32(&tmA, g.A, g.M, g.K, g.lda * 4, 32, 32);
  else rc = make_tensor_map_f32(&tmA, g.A, g.K, g.M, g.lda * 4, 32, 128);
  if (rc != PENEO_OK) return rc;
  // B stored [N, K] (tb): K-major box 32 x 128 ; else [K, N]: MN-major boxes 32 x 32
  if (g.tb) rc = make_tensor_map_f32(&tmB, g.B, g.K, g.N, g.ldb * 4, 32, 128);
  else rc = make_tensor_map_f32(&tmB, g.B, g.N, g.K, g.ldb * 4, 32, 32);
  if (rc != PENEO_OK) return rc;
  Args a{};
  a.C = g.C, a.ldc = g.ldc, a.C2 = g.C2, a.ldc2 = g.ldc2, a.bias = g.bias;
  a.M = g.M, a.N = g.N, a.K = g.K, a.mode = g.mode;
  a.drop_thresh = g.drop_thresh, a.drop_key = g.drop_key, a.row0 = g.row0, a.drop_scale = g.drop_scale;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    PENEO_CUDA_TRY(cudaGetDevice(&dev));
    PENEO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int num_kb = (g.K + 31) / 32;
  a.n_blocks = (g.N + 127) / 128;
  const int tiles = ((g.M + 127) / 128) * a.n_blocks;
  int splits = 1;
  if (g.mode == 2) splits = std::max(1, std::min(num_kb, (sms + tiles - 1) / tiles));
  a.kb_per_split = (num_kb + splits - 1) / splits;
  a.splits = (num_kb + a.kb_per_split - 1) / a.kb_per_split;
  a.num_items = (int64_t)tiles * a.splits;
  const int grid = static_cast<int>(std::min<int64_t>(a.num_items, sms));
#define GO(AM, BM)                                                                                                       \
  {                                                                                                                      \
    PENEO_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32_kernel<AM, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)); \
    gemm_tf32_kernel<AM, BM><<<grid, 192, kSmemBytes, st>>>(tmA, tmB, a);                                                 \
  }
  // A is MN-major when stored [K, M]; B is MN-major when stored [K, N]
  if (g.ta && !g.tb) GO(true, true)
  else if (g.ta && g.tb) GO(true, false)
  else if (!g.ta && !g.tb) GO(false, true)
  else GO(false, false)
#undef GO
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
