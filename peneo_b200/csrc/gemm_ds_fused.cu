// T1f — the hidden-activation backward fused into the dS GEMM (saved-activation route, CTA pairs).
//
//   G  = (dz W_out) * SiLU'(2 h)          formed IN THE GEMM'S OPERAND STAGE from the saved pre-activations h
//   dS = G W_mid                          [rows, 384], K = 1920, tcgen05 cta_group::2 exactly as gemm_bwd_tc2.cu
//   dW_out += dz^T SiLU(2 h) , db_out += sum dz          reduced on the way
//
// pair_bwd_elem.cu (T1e) does the first and third line as a stand-alone HBM-bound pass (read h, write G) and the dS
// GEMM then reads G again.  Here a TMA load brings an [128 rows x 64 features] tile of h into the A-operand slot of a
// pipeline stage, eight transform warps turn it into the G tile in place (shared memory), the tensor core consumes it
// from there, and a TMA store writes it back over h for the dW_mid GEMM that follows: the GEMM's 3.75 KB per pair of G
// reads and T1e's launch disappear, the transform hides behind (or sets the pace next to) the MMAs.
//
// STATUS: experimental (PENEO_T1F=1).  Gradients equal the T1e route's, but on B200 the kernel takes 1.46 ms per
// 524 288-pair chunk against 0.71 ms (T1e) + 0.58 ms (plain dS GEMM): 16 transform warps at one 4-byte feature pair per
// lane and row issue ~1 instruction per clock and SM (~2 500 cycles per K block against ~700 for the MMAs), and the
// shared-memory atomics of the dW_out reduction (16 warps adding into the same 64 x 3 cells per K block) take 39 % of
// the warp samples.  What it needs: 16-byte cells per lane as in T1e, and dW_out through a small UMMA (T1's form) or
// warp-exclusive feature ownership instead of shared atomics.
//
// Roles (512 threads): warp 0 TMA loads, warp 1 MMA issue (leader CTA), warps 2-5 epilogue (dS -> bf16 -> HBM),
// warp 6 TMA stores of the finished G tiles, warps 8-15 transform.  Barriers per stage s:
//   a_full[s]  local   the h tile of this CTA has landed (TMA bytes)
//   b_full[s]  leader  the W_mid halves of both CTAs have landed
//   g_ready[s] leader  the transform warps of both CTAs have written G (16 arrivals)
//   t_done[s]  local   the 8 transform warps of this CTA are done with the tile -> the store warp may read it
//   empty[s]   local   the MMAs have read the stage (tcgen05.commit, multicast) AND the TMA store has read it
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {
namespace t1f {

constexpr int kN = 384, kF = 5 * kN;
constexpr int kBlk = 64 * 128;                    // one [64 rows x 64 cols] swizzled box = 8 KB
// Two rings: the h tiles come from HBM (~2 us away) and a slot stays busy until its TMA store has read it, so the A
// ring is deep; the W_mid halves come from the L2.  (One 4-stage ring for both: 2 780 cycles per K block, load-latency-bound.)
constexpr int kABytes = 2 * kBlk;  // A (h -> G): [128 rows x 64 features] = 16 KB
constexpr int kBBytes = 3 * kBlk;  // this CTA's half of B: 24 KB
constexpr int kAStages = 6, kBStages = 3;
#ifndef PENEO_T1F_WARPS
#define PENEO_T1F_WARPS 16
#endif
constexpr int kTWarps = PENEO_T1F_WARPS, kTWarp0 = 8;  // transform warps (8: 3 430 cycles per K block, latency-bound)
constexpr int kTRows = 128 / kTWarps;             // rows of a tile per transform warp
constexpr int kThreads = 32 * (kTWarp0 + kTWarps);
constexpr int kNumKb = kF / 64;                   // 30 K blocks of 64 features (six per head)

struct Smem {
  static constexpr int aring = 0;
  static constexpr int bring = aring + kAStages * kABytes;
  static constexpr int ostage = bring + kBStages * kBBytes;  // dS store staging: four epilogue warps x 2 KB
  static constexpr int dz = ostage + 4 * 2048;                   // two tiles of [5 heads][128 rows] uint4 (dz as bf16x2 broadcasts)
  static constexpr int dwout = dz + 2 * kNumHeads * 128 * 16;    // [3][1920] fp32 partial sums of dW_out
  static constexpr int misc = dwout + 3 * kF * 4;                // scale[5] | class_w[3]
  static constexpr int bars = misc + 64;
  static constexpr int total = bars + 512;
};
constexpr int bAFull = 0, bGReady = bAFull + kAStages, bTDone = bGReady + kAStages, bAEmpty = bTDone + kAStages,
              bBFull = bAEmpty + kAStages, bBEmpty = bBFull + kBStages, bAccFull = bBEmpty + kBStages, bAccEmpty = bAccFull + 1,
              bCount = bAccEmpty + 1;
static_assert(bCount * 8 + 16 <= 512, "barrier area too small");
constexpr int kSmemBytes = Smem::total + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct Args {
  int64_t m_total;  // rows (pairs) of the chunk
  int64_t g0;       // batch-flat index of the chunk's first pair
  int64_t num_pair_items;
  __nv_bfloat16* dS;  // [rows, 384]
  const float* logits[kNumHeads];
  const int64_t* tags[kNumHeads];
  const float* grad_out6;
  const double* loss_final;
  float ratio[kNumHeads], class_w[3];
  float* dbout[kNumHeads];
  const float4* wout4;  // [5][384] : (W_out[0][f], W_out[1][f], W_out[2][f] or 0, 0)
  float* dwout_part;    // [gridDim.x][3][1920], accumulated
  uint32_t drop_thresh;
  float drop_scale;
  uint32_t drop_key[kNumHeads];
};

__device__ __forceinline__ uint32_t tanh_bf16x2(uint32_t h) {
  uint32_t t;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(h));
  return t;
}

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_ds_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  float* s_dwout = reinterpret_cast<float*>(smem + Smem::dwout);
  float* s_scale = reinterpret_cast<float*>(smem + Smem::misc);
  float* s_cw = s_scale + 5;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < kAStages; ++s) {
      ptx::mbar_init(&bars[bAFull + s], 1);
      ptx::mbar_init(&bars[bGReady + s], 2 * kTWarps);
      ptx::mbar_init(&bars[bTDone + s], kTWarps);
      ptx::mbar_init(&bars[bAEmpty + s], 2);  // the MMAs (commit) and the TMA store have read the slot
    }
    for (int s = 0; s < kBStages; ++s) ptx::mbar_init(&bars[bBFull + s], 2), ptx::mbar_init(&bars[bBEmpty + s], 1);
    ptx::mbar_init(&bars[bAccFull], 1);
    ptx::mbar_init(&bars[bAccEmpty], 8);  // 4 epilogue warps of each CTA
    ptx::fence_barrier_init();
  }
  for (int e = threadIdx.x; e < 3 * kF; e += kThreads) s_dwout[e] = 0.f;
  if (threadIdx.x < kNumHeads)
    s_scale[threadIdx.x] = (a.grad_out6[5] * a.ratio[threadIdx.x] + a.grad_out6[threadIdx.x]) /
                           static_cast<float>(a.loss_final[2 * threadIdx.x + 1]);
  if (threadIdx.x >= 8 && threadIdx.x < 11) s_cw[threadIdx.x - 8] = a.class_w[threadIdx.x - 8];
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
  auto block_of = [&](int64_t item) { return 2 * item + rank; };  // this CTA's 128-row block of the pair item

  if (warp == 0) {
    // ============================== TMA loads of the h tiles ==============================
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t item = cid; item < a.num_pair_items; item += ncl) {
        const int64_t mb = block_of(item);
        for (int kb = 0; kb < kNumKb; ++kb) {
          ptx::mbar_wait(&bars[bAEmpty + s], ph ^ 1);
          // this CTA's [128 rows x 64 features] tile of h: completes on its OWN barrier (the transform warps wait there)
          ptx::mbar_arrive_expect_tx(&bars[bAFull + s], kABytes);
          ptx::tma_load_2d(smem + Smem::aring + s * kABytes, &tmA, &bars[bAFull + s], kb * 64, static_cast<int32_t>(mb * 128));
          if (++s == kAStages) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 7) {
    // ============================== TMA loads of the W_mid halves ==============================
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t item = cid; item < a.num_pair_items; item += ncl) {
        for (int kb = 0; kb < kNumKb; ++kb) {
          const int k0 = kb * 64;
          ptx::mbar_wait(&bars[bBEmpty + s], ph ^ 1);
          unsigned char* st = smem + Smem::bring + s * kBBytes;
          // this CTA's half of B: columns [128 r, 128 r + 128) for the N = 256 MMA, [256 + 64 r, + 64) for the N = 128 one
          ptx::tma_load_2d_2sm(st, &tmB, &bars[bBFull + s], static_cast<int32_t>(128 * rank), k0);
          ptx::tma_load_2d_2sm(st + kBlk, &tmB, &bars[bBFull + s], static_cast<int32_t>(128 * rank + 64), k0);
          ptx::tma_load_2d_2sm(st + 2 * kBlk, &tmB, &bars[bBFull + s], static_cast<int32_t>(256 + 64 * rank), k0);
          if (leader) ptx::mbar_arrive_expect_tx(&bars[bBFull + s], 2 * kBBytes);
          else ptx::mbar_arrive_remote(&bars[bBFull + s], 0);
          if (++s == kBStages) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issue (leader) ==============================
    if (leader && ptx::elect_one()) {
      constexpr uint32_t idesc256 = ptx::umma_idesc_bf16_major(256, 256, false, true);
      constexpr uint32_t idesc128 = ptx::umma_idesc_bf16_major(256, 128, false, true);
      int s = 0, sb = 0;
      uint32_t ph = 0, phb = 0;
      int it = 0;
      for (int64_t item = cid; item < a.num_pair_items; item += ncl, ++it) {
        ptx::mbar_wait(&bars[bAccEmpty], (it & 1) ^ 1);
        ptx::tc_fence_after();
        for (int kb = 0; kb < kNumKb; ++kb) {
          ptx::mbar_wait(&bars[bBFull + sb], phb);
          ptx::mbar_wait(&bars[bGReady + s], ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + Smem::aring + s * kABytes);
          const uint32_t b_addr = ptx::smem_u32(smem + Smem::bring + sb * kBBytes);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = ptx::umma_desc_sw128(a_addr + ks * 32);
            ptx::umma_ss_2sm(tmem, ad, ptx::umma_desc_mn_sw128(b_addr + ks * 2048, kBlk, 1024), idesc256, (kb | ks) != 0);
            ptx::umma_ss_2sm(tmem + 256, ad, ptx::umma_desc_mn_sw128(b_addr + 2 * kBlk + ks * 2048, kBlk, 1024), idesc128,
                             (kb | ks) != 0);
          }
          ptx::tc_commit_2sm(&bars[bAEmpty + s], 3);
          ptx::tc_commit_2sm(&bars[bBEmpty + sb], 3);
          if (++s == kAStages) s = 0, ph ^= 1;
          if (++sb == kBStages) sb = 0, phb ^= 1;
        }
        ptx::tc_commit_2sm(&bars[bAccFull], 3);
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ============================== epilogue: dS -> bf16 -> global ==============================
    const int q = warp % 4;
    int it = 0;
    for (int64_t item = cid; item < a.num_pair_items; item += ncl, ++it) {
      const int64_t mb = block_of(item);
      ptx::mbar_wait(&bars[bAccFull], it & 1);
      ptx::tc_fence_after();
      unsigned char* ob = smem + Smem::ostage + q * 2048;
      const int64_t wrow0 = mb * 128 + q * 32;
#pragma unroll 1
      for (int piece = 0; piece < kN / 32; ++piece) {
        uint32_t r[32];
        ptx::tmem_ld_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + piece * 32, r);
        ptx::tmem_ld_wait();
        if (piece == kN / 32 - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(&bars[bAccEmpty]);
        }
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
          if (wrow0 + rr < a.m_total) *reinterpret_cast<uint4*>(a.dS + (wrow0 + rr) * kN + piece * 32 + c16 * 8) = val;
        }
      }
    }
  } else if (warp == 6) {
    // ============================== TMA stores of the finished G tiles ==============================
    if (ptx::elect_one()) {
      int s = 0, prev = -1;
      uint32_t ph = 0;
      for (int64_t item = cid; item < a.num_pair_items; item += ncl) {
        const int64_t mb = block_of(item);
        for (int kb = 0; kb < kNumKb; ++kb) {
          ptx::mbar_wait(&bars[bTDone + s], ph);  // (the transform warps fenced the async proxy before arriving)
          ptx::tma_store_2d(&tmA, smem + Smem::aring + s * kABytes, kb * 64, static_cast<int32_t>(mb * 128));
          ptx::bulk_commit_group();
          if (prev >= 0) {  // the store issued one step earlier has read its tile: that slot may be refilled
            ptx::bulk_wait_group_read<1>();
            ptx::mbar_arrive(&bars[bAEmpty + prev]);
          }
          prev = s;
          if (++s == kAStages) s = 0, ph ^= 1;
        }
      }
      if (prev >= 0) {
        ptx::bulk_wait_group_read<0>();
        ptx::mbar_arrive(&bars[bAEmpty + prev]);
      }
      ptx::bulk_wait_group<0>();  // every G tile has been written
    }
  } else if (warp >= kTWarp0) {
    // ============================== transform: h -> G in place ==============================
    const int tw = warp - kTWarp0, tt = tw * 32 + lane;  // 0 .. 32 * kTWarps - 1
    constexpr int kDzItems = (kNumHeads * 128 + 32 * kTWarps - 1) / (32 * kTWarps);  // (row, head) items per thread
    float db0[kDzItems] = {}, db1[kDzItems] = {}, db2[kDzItems] = {};  // dz sums of this thread's (row, head) items
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int64_t item = cid; item < a.num_pair_items; item += ncl, ++it) {
      const int64_t mb = block_of(item);
      // ---- dz of the block's 128 rows x 5 heads: scale_k w[t] (softmax(z) - onehot(t)), as bf16x2 broadcasts
      uint4* dzt = reinterpret_cast<uint4*>(smem + Smem::dz) + (it & 1) * (kNumHeads * 128);
#pragma unroll
      for (int j = 0; j < kDzItems; ++j) {
        const int idx = tt + 32 * kTWarps * j;  // (head, row) = (idx / 128, idx % 128); 640 items
        if (idx < kNumHeads * 128) {
          const int hd = idx >> 7, row = idx & 127, C = head_classes(hd);
          const int64_t lr = mb * 128 + row;
          float z0 = 0.f, z1 = 0.f, z2 = 0.f;
          if (lr < a.m_total) {
            const int64_t gp = a.g0 + lr;
            const float* p = a.logits[hd] + gp * C;
            z0 = p[0], z1 = p[1];
            const float x2 = C == 3 ? p[2] : -INFINITY;
            const long long t64 = a.tags[hd][gp];
            const int tg = (t64 < 0 || t64 >= C) ? 3 : static_cast<int>(t64);  // 3 = target outside [0, C): poison with NaN
            const float mx = fmaxf(fmaxf(z0, z1), x2);
            const float e0 = __expf(z0 - mx), e1 = __expf(z1 - mx), e2 = C == 3 ? __expf(x2 - mx) : 0.f;
            const float sw = s_scale[hd] * s_cw[tg == 3 ? 0 : tg];
            const float gsc = tg == 3 ? NAN : __fdividef(sw, e0 + e1 + e2);
            z0 = gsc * e0 - (tg == 0 ? sw : 0.f);
            z1 = gsc * e1 - (tg == 1 ? sw : 0.f);
            z2 = C == 3 ? gsc * e2 - (tg == 2 ? sw : 0.f) : 0.f;
          }
          db0[j] += z0, db1[j] += z1, db2[j] += z2;  // (item j of this thread always belongs to the same head)
          dzt[idx] = make_uint4(ptx::pack_bf16x2(z0, z0), ptx::pack_bf16x2(z1, z1), ptx::pack_bf16x2(z2, z2), 0u);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kTWarps) : "memory");  // the transform warps only
      // W_out of this lane's feature pair, one K block ahead (the loads are L2 hits ~600 cycles away: issued at the top of
      // a K block they would stall every warp once per block)
      float4 np0 = a.wout4[2 * lane], np1 = a.wout4[2 * lane + 1];
      for (int kb = 0; kb < kNumKb; ++kb) {
        const int k = kb / 6;                       // head of this K block
        const int fh = (kb - 6 * k) * 64 + 2 * lane;  // this lane's feature pair inside the head
        const float4 p0 = np0, p1 = np1;
        if (kb + 1 < kNumKb) np0 = a.wout4[(kb + 1) * 64 + 2 * lane], np1 = a.wout4[(kb + 1) * 64 + 2 * lane + 1];  // ((kb + 1) * 64 = stacked feature)
        const uint32_t w0 = ptx::pack_bf16x2(p0.x, p1.x), w1 = ptx::pack_bf16x2(p0.y, p1.y), w2 = ptx::pack_bf16x2(p0.z, p1.z);
        ptx::mbar_wait(&bars[bAFull + s], ph);
        unsigned char* tile = smem + Smem::aring + s * kABytes;
        const uint4* dzk = dzt + k * 128;
        float facc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // dW_out partial sums over this warp's rows: [class][feature of the pair]
#pragma unroll 1
        for (int half = 0; half < kTRows / 8; ++half) {
        uint32_t acc0 = 0u, acc1 = 0u, acc2 = 0u;  // ... accumulated in bf16x2 over 8 rows at a time (as T1e does)
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
          const int row = tw + kTWarps * (8 * half + ii);
          // 128 B per row, 16-byte chunk index XOR (row & 7) (SWIZZLE_128B); this lane's feature pair = 4 bytes
          uint32_t* cell = reinterpret_cast<uint32_t*>(tile + row * 128 + (((lane >> 2) ^ (row & 7)) * 16) + (lane & 3) * 4);
          const uint32_t h2 = *cell;
          const uint4 dz = dzk[row];
          const uint32_t t2 = tanh_bf16x2(h2);
          uint32_t m2 = ptx::hfma2_bf16(h2, t2, h2);                            // SiLU(2 h) = h (1 + tanh h)
          const uint32_t sg2 = ptx::hfma2_bf16(0x3F003F00u, t2, 0x3F003F00u);   // sigmoid(2 h)
          const uint32_t oms2 = ptx::hfma2_bf16(0xBF00BF00u, t2, 0x3F003F00u);  // 1 - sigmoid
          uint32_t dv2 = ptx::hfma2_bf16(m2, oms2, sg2);                        // SiLU' = sg + m (1 - sg)
          if (DROP) {
            const uint32_t grow = static_cast<uint32_t>(a.g0 + mb * 128 + row);
            const float ms0 = drop_keep(a.drop_key[k], a.drop_thresh, grow, static_cast<uint32_t>(fh)) ? a.drop_scale : 0.f;
            const float ms1 = drop_keep(a.drop_key[k], a.drop_thresh, grow, static_cast<uint32_t>(fh + 1)) ? a.drop_scale : 0.f;
            const uint32_t ms2 = ptx::pack_bf16x2(ms0, ms1);
            m2 = ptx::hmul2_bf16(m2, ms2), dv2 = ptx::hmul2_bf16(dv2, ms2);
          }
          const uint32_t gm = ptx::hfma2_bf16(dz.z, w2, ptx::hfma2_bf16(dz.y, w1, ptx::hmul2_bf16(dz.x, w0)));
          *cell = ptx::hmul2_bf16(gm, dv2);
          acc0 = ptx::hfma2_bf16(dz.x, m2, acc0), acc1 = ptx::hfma2_bf16(dz.y, m2, acc1), acc2 = ptx::hfma2_bf16(dz.z, m2, acc2);
        }
        facc[0] += __uint_as_float(acc0 << 16), facc[1] += __uint_as_float(acc0 & 0xFFFF0000u);
        facc[2] += __uint_as_float(acc1 << 16), facc[3] += __uint_as_float(acc1 & 0xFFFF0000u);
        facc[4] += __uint_as_float(acc2 << 16), facc[5] += __uint_as_float(acc2 & 0xFFFF0000u);
        }
        ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core and to the TMA store
        __syncwarp();
        if (lane == 0) {
          arrive_leader(&bars[bGReady + s]);
          ptx::mbar_arrive(&bars[bTDone + s]);
        }
#ifndef PENEO_T1F_NO_DW  // (timing experiment: how much the shared-memory atomics cost)
        {
          float* dst = s_dwout + k * kN + fh;
          atomicAdd(dst, facc[0]), atomicAdd(dst + 1, facc[1]);
          atomicAdd(dst + kF, facc[2]), atomicAdd(dst + kF + 1, facc[3]);
          atomicAdd(dst + 2 * kF, facc[4]), atomicAdd(dst + 2 * kF + 1, facc[5]);
        }
#else
        if (facc[0] + facc[1] + facc[2] + facc[3] + facc[4] + facc[5] == 12345.678f) s_dwout[0] = 1.f;
#endif
        if (++s == kAStages) s = 0, ph ^= 1;
      }
    }
    // db_out: this thread's items j = 0, 1, 2 belong to heads (tt + 256 j) / 128
#pragma unroll
    for (int j = 0; j < kDzItems; ++j) {
      const int idx = tt + 32 * kTWarps * j;
      if (idx < kNumHeads * 128) {
        const int hd = idx >> 7;
        float r0 = db0[j], r1 = db1[j], r2 = db2[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          r0 += __shfl_xor_sync(0xffffffffu, r0, o), r1 += __shfl_xor_sync(0xffffffffu, r1, o);
          r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        }
        if (lane == 0) {
          atomicAdd(&a.dbout[hd][0], r0), atomicAdd(&a.dbout[hd][1], r1);
          if (head_classes(hd) == 3) atomicAdd(&a.dbout[hd][2], r2);
        }
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  {  // this CTA's dW_out partial sums -> its private slice of the global partial buffer
    float* part = a.dwout_part + static_cast<size_t>(blockIdx.x) * (3 * kF);
    for (int e = threadIdx.x; e < 3 * kF; e += kThreads) part[e] += s_dwout[e];
  }
  if (warp == 1) ptx::tmem_dealloc_2sm(tmem, 512);
}

}  // namespace t1f

// h (bf16 [rows, 1920], overwritten by G) -> dS (bf16 [rows, 384]) ; dW_out partial sums, db_out
int launch_gemm_ds_fused(const void* pack, const PackLayout& L, __nv_bfloat16* hg, const __nv_bfloat16* wmid_full,
                         __nv_bfloat16* dS, int64_t g0, int rows, const FusedLossBwd& fused, float* dwout_part,
                         cudaStream_t st, const DropSpec* drop) {
  using namespace t1f;
  if (rows == 0) return PENEO_OK;
  const char* base = static_cast<const char*>(pack);
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA, hg, kF, rows, kF * 2, 64, 128)) != PENEO_OK) return rc;
  if ((rc = make_tensor_map_bf16(&tmB, wmid_full, kN, kF, kN * 2, 64, 64)) != PENEO_OK) return rc;
  Args a{};
  a.m_total = rows, a.g0 = g0;
  a.num_pair_items = ((rows + 127) / 128 + 1) / 2;
  a.dS = dS;
  for (int h = 0; h < kNumHeads; ++h)
    a.logits[h] = fused.logits[h], a.tags[h] = fused.tags[h], a.ratio[h] = fused.ratio[h], a.dbout[h] = fused.dbout[h];
  for (int c = 0; c < 3; ++c) a.class_w[c] = fused.class_w[c];
  a.grad_out6 = fused.grad_out6, a.loss_final = fused.loss_final;
  a.wout4 = reinterpret_cast<const float4*>(base + L.wout_f32x4);
  a.dwout_part = dwout_part;
  if (drop && drop->thresh) {
    a.drop_thresh = drop->thresh, a.drop_scale = drop->scale;
    for (int h = 0; h < kNumHeads; ++h) a.drop_key[h] = drop_key(*drop, site_head(h, 0));
  }
  int dev = 0, sms = 148;
  PENEO_CUDA_TRY(cudaGetDevice(&dev));
  PENEO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int clusters = static_cast<int>(std::min<int64_t>(a.num_pair_items, sms / 2));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmemBytes, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  if (a.drop_thresh) {
    PENEO_CUDA_TRY(cudaFuncSetAttribute(gemm_ds_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    PENEO_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_ds_fused_kernel<true>, tmA, tmB, a));
  } else {
    PENEO_CUDA_TRY(cudaFuncSetAttribute(gemm_ds_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    PENEO_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_ds_fused_kernel<false>, tmA, tmB, a));
  }
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
