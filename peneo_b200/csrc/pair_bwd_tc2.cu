// T1 on CTA pairs (cta_group::2) with the loss backward fused into the pair tiles.
//
// Same job as pair_bwd_tc.cu — for every pair of a chunk [g0, g0 + rows) of the batch-flat pair list and every head k
//   s   = SiLU(a_i + b_j)                         regenerated (CUDA cores -> TMEM), also stored: S  [rows, 384]
//   u_k = W_mid,k s + b_mid,k                     tcgen05, 15 chunks of 128 mid features
//   m_k = SiLU(u_k)                               never leaves the SM: dW_out,k += dz_k^T m_k on the tensor cores
//   g_k = (dz_k W_out,k) * SiLU'(u_k)             stored: G [rows, 1920]   (dW_mid = G^T S, dS = G W_mid)
// — with three changes that the ncu profile of the single-CTA kernel asked for (shared-memory port 86 % busy:
// 2130 LSU wavefronts + 1792 TMA / tensor-core wavefronts per 4540-cycle chunk):
//   * CTA pairs: the two CTAs of a cluster run every UMMA together (M = 256); each owns one 128-pair tile and
//     supplies half of the W_mid rows of a chunk, so the W stream costs a CTA 384 + 384 wavefronts instead of 768 + 768;
//   * the W_out rows for dz W_out are read as 16-byte vectors (192 instead of 768 broadcast LDS per chunk);
//   * b_mid rides on the tensor cores exactly as in K2 (pair_heads_tc2.cu): a constant 25th K step [1, 0, .., 0] of the
//     A operand times a resident bias tile, so the accumulator already holds (u + b_mid) / 2 — no bias LDS / FADD;
//   * FUSED: dz_k is computed in registers from the stored logits, the tags and the loss normaliser
//       dz = (g_total ratio_k + g_k) / Z_k * w[t] * (softmax(z) - onehot(t))      (model/custom_loss.py:189-202)
//     so d loss / d logits never exists in HBM and db_out = sum dz is reduced here too (no separate loss-backward
//     and column-sum kernels).
//
// dW_out on a CTA pair: the two CTAs hold DIFFERENT pairs (K index), so their products cannot share a B operand.
// One N = 16 MMA does both: B[k][0:8] = dz^T of CTA 0 (its half of B: the first 8-row swizzle atom of its dz^T tile),
// B[k][8:16] = dz^T of CTA 1, A rows [0,128) = m^T of CTA 0, [128,256) = m^T of CTA 1; each CTA reads its own product
// from its own 8 columns of D (the cross terms land in the other 8 columns and are ignored).
//
// Warp roles: warp 0 TMA (W_mid stream), warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator, warps 4 .. 4 +
// kEpiWarps epilogue, then four pair-producer warps.  Barriers the MMA waits on live in the LEADER (the peer arrives remotely);
// everything the MMA releases is a tcgen05.commit multicast to both CTAs.
// TMEM columns: [0,192) s | [192,320) u buffer 0 | [320,448) u buffer 1 | [448,464) D_w 0 | [464,480) D_w 1 | [480,496) ones
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {
namespace t1p {

constexpr int D = 384;
constexpr int kChunks = 15;       // 5 heads x 3 chunks of 128 mid features
constexpr int kKChunks = 6;       // 384 / 64
constexpr int kWStages = 4;  // (shared-memory budget; 3 .. 6 stages measure the same: the MMA side is not starved)
constexpr int kWStageBytes = 64 * 64 * 2;  // 8 KB: this CTA's 64 of the chunk's 128 W_mid rows
constexpr int kStageRowBytes = D * 2;      // staging: 128 rows x 768 B
// Epilogue warps: 8 (two per scheduler, two 32-column slices of a chunk each) or 16 (one slice each).  Measured on
// B200 (262 144 pairs): 8 warps 0.538 ms, 16 warps 0.583 ms — the epilogue is not short of warps in flight.
constexpr int kEpiWarps = 8;
constexpr int kSlices = 16 / kEpiWarps;  // 32-column slices per warp and chunk
constexpr int kProdWarp0 = 4 + kEpiWarps;
constexpr int kProdRows = kEpiWarps == 8 ? 4 : 2;  // pair rows a producer warp has in flight (register budget)
constexpr int kThreads = 32 * (kProdWarp0 + 4);
constexpr int kLdG = 5 * D;                // row stride of G

constexpr uint32_t kColS = 0, kColU = 192, kColDw = 448, kColOne = 480;
constexpr int kBiasTileBytes = 64 * 128;  // [64 rows x 128 B] SWIZZLE_128B K-major: the 32-byte K step x of a row belongs to
constexpr int kBiasTiles = 4;             // chunk 4 * tile + x, its K column 0 = b_mid / 2 of that feature (see K2)
constexpr int kOutWarpBytes = 32 * 64;  // per epilogue warp: 32 rows x 32 bf16, XOR-swizzled 16-byte chunks

struct Smem {
  static constexpr int w = 0;
  static constexpr int stage = w + kWStages * kWStageBytes;
  static constexpr int mtile = stage + 128 * kStageRowBytes;  // m chunk, 2 x 2 boxes of [64 pairs x 64 features] bf16
  static constexpr int dzt = mtile + 128 * 128 * 2;           // dz^T, 2 K blocks of [16 x 64 pairs] bf16
  static constexpr int bias = dzt + 2 * 16 * 128;             // bias B-operand tiles
  static constexpr int out = bias + kBiasTiles * kBiasTileBytes;  // per epilogue warp [32 rows][64 B]
  static constexpr int wout = out + kEpiWarps * kOutWarpBytes;        // [3][960] bf16x2: W_out[c] of feature pairs (2f, 2f+1)
  static constexpr int misc = wout + 3 * (5 * D / 2) * 4;     // fused loss: scale[5] | class_w[3] | db_out[5][4]
  static constexpr int bars = misc + 128;
  static constexpr int total = bars + 512;
};
constexpr int bWFull = 0, bWEmpty = bWFull + kWStages, bUFull = bWEmpty + kWStages, bUFree = bUFull + 2,
              bSFull = bUFree + 2, bSFree = bSFull + kKChunks, bMFull = bSFree + kKChunks, bMFree = bMFull + 1,
              bDwFull = bMFree + 1, bDwFree = bDwFull + 2, bCount = bDwFree + 2;
static_assert(Smem::mtile % 1024 == 0 && Smem::dzt % 1024 == 0 && Smem::stage % 1024 == 0 && Smem::bias % 1024 == 0,
              "UMMA operand tiles need 1024-byte alignment");
static_assert(Smem::wout % 16 == 0, "W_out rows are read as 16-byte vectors");
static_assert(Smem::total + 1024 <= 227 * 1024, "shared memory budget");
static_assert(bCount * 8 + 16 <= 512, "barrier area too small");
constexpr int kSmemBytes = Smem::total + 1024;

struct Args {
  const __nv_bfloat16* ab;   // [batch*n, 768] : 0.5*A | 0.5*Bm
  const float* bmid_half;    // [1920]
  const float4* wout4;       // [5][384] : (W_out[0][f], W_out[1][f], W_out[2][f] or 0, 0)
  const float* dz[kNumHeads];  // !FUSED: d loss / d logits, fp32 [batch*P, C_h]
  // FUSED: the loss backward happens here
  const float* logits[kNumHeads];   // fp32 [batch*P, C_h] (the forward pass's output)
  const int64_t* tags[kNumHeads];   // int64 [batch*P]
  const float* grad_out6;           // device: d L / d (loss_0..4, total)
  const double* loss_final;         // device: [5][2] (sum w*nll, sum w) written by the loss forward
  float ratio[kNumHeads], class_w[3];
  float* dbout[kNumHeads];          // db_out gradients (accumulated)
  __nv_bfloat16* S;          // [rows, 384]
  __nv_bfloat16* G;          // [rows, 1920]
  float* dwout_part;         // [gridDim.x][3][1920] per-CTA partial sums of dz^T M (accumulated)
  int32_t n, pairs_per_doc;
  int64_t g0;                // first flat pair of the chunk
  int32_t rows, num_tiles;
  uint32_t drop_thresh;      // the forward pass's dropout after the hidden SiLU, regenerated (0 = none)
  float drop_scale;
  uint32_t drop_key[kNumHeads];
};

template <bool DROP, bool FUSED>
__global__ void __launch_bounds__(kThreads, 1) pair_bwd_prep_pair_kernel(const __grid_constant__ CUtensorMap tmW, const Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  float* s_scale = reinterpret_cast<float*>(smem + Smem::misc);  // [5]
  float* s_cw = s_scale + 5;                                      // [3]
  float* s_dbout = s_scale + 8;                                   // [5][4]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < kWStages; ++s) ptx::mbar_init(&bars[bWFull + s], 2), ptx::mbar_init(&bars[bWEmpty + s], 1);
    for (int s = 0; s < 2; ++s) ptx::mbar_init(&bars[bUFull + s], 1), ptx::mbar_init(&bars[bUFree + s], 2 * kEpiWarps);
    for (int s = 0; s < kKChunks; ++s) ptx::mbar_init(&bars[bSFull + s], 8), ptx::mbar_init(&bars[bSFree + s], 1);
    ptx::mbar_init(&bars[bMFull], 2 * kEpiWarps), ptx::mbar_init(&bars[bMFree], 1);
    for (int s = 0; s < 2; ++s) ptx::mbar_init(&bars[bDwFull + s], 1), ptx::mbar_init(&bars[bDwFree + s], 8);  // 4 flushing warps per CTA
    ptx::fence_barrier_init();
  }
  // bias operand tiles: zero, then K column 0 of (chunk c, row r) = b_mid / 2 of feature c * 128 + 64 * rank + r
  for (int e = threadIdx.x; e < kBiasTiles * kBiasTileBytes / 16; e += kThreads)
    reinterpret_cast<uint4*>(smem + Smem::bias)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int e = threadIdx.x; e < kChunks * 64; e += kThreads) {
    const int c = e >> 6, r = e & 63;
    const float bh = a.bmid_half[c * 128 + static_cast<int>(rank) * 64 + r];
    *reinterpret_cast<__nv_bfloat16*>(smem + Smem::bias + (c >> 2) * kBiasTileBytes + r * 128 + (((2 * (c & 3)) ^ (r & 7)) * 16)) =
        __float2bfloat16_rn(bh);
  }
  uint32_t* s_wout = reinterpret_cast<uint32_t*>(smem + Smem::wout);
  for (int e = threadIdx.x; e < 5 * D / 2; e += kThreads) {
    const float4 w0 = a.wout4[2 * e], w1 = a.wout4[2 * e + 1];
    s_wout[e] = ptx::pack_bf16x2(w0.x, w1.x);
    s_wout[5 * D / 2 + e] = ptx::pack_bf16x2(w0.y, w1.y);
    s_wout[5 * D + e] = ptx::pack_bf16x2(w0.z, w1.z);
  }
  for (int e = threadIdx.x; e < 2 * 16 * 128 / 16; e += kThreads)  // dz^T rows 3..15 stay zero for the whole kernel
    reinterpret_cast<uint4*>(smem + Smem::dzt)[e] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x < 28) {
    if (FUSED && threadIdx.x < 5)
      s_scale[threadIdx.x] = (a.grad_out6[5] * a.ratio[threadIdx.x] + a.grad_out6[threadIdx.x]) /
                             static_cast<float>(a.loss_final[2 * threadIdx.x + 1]);
    else if (threadIdx.x >= 5 && threadIdx.x < 8) s_cw[threadIdx.x - 5] = a.class_w[threadIdx.x - 5];
    else if (threadIdx.x >= 8) s_dbout[threadIdx.x - 8] = 0.f;
  }
  ptx::fence_proxy_async();
  ptx::cluster_sync_all();  // both CTAs' barriers exist before anyone (TMA of the peer, remote arrives) touches them
  if (warp == 2) {
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

  // cluster `cid` takes tile pairs cid, cid + nclusters, ...; this CTA owns tile 2 * pair + rank
  const int cid = static_cast<int>(blockIdx.x >> 1), ncl = static_cast<int>(gridDim.x >> 1);
  const int num_pairs = (a.num_tiles + 1) / 2;
  const int my_tiles = (num_pairs > cid) ? (num_pairs - 1 - cid) / ncl + 1 : 0;
  auto tile_of = [&](int it) { return 2 * (static_cast<int64_t>(cid) + static_cast<int64_t>(it) * ncl) + rank; };

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (ptx::elect_one()) {
      int ws = 0;
      uint32_t wph = 0;
      for (int it = 0; it < my_tiles; ++it)
        for (int c = 0; c < kChunks; ++c)
          for (int kc = 0; kc < kKChunks; ++kc) {
            ptx::mbar_wait(&bars[bWEmpty + ws], wph ^ 1);
            ptx::tma_load_2d_2sm(smem + Smem::w + ws * kWStageBytes, &tmW, &bars[bWFull + ws], kc * 64,
                                 c * 128 + static_cast<int>(rank) * 64);
            if (leader) ptx::mbar_arrive_expect_tx(&bars[bWFull + ws], 2 * kWStageBytes);
            else ptx::mbar_arrive_remote(&bars[bWFull + ws], 0);
            if (++ws == kWStages) ws = 0, wph ^= 1;
          }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (leader) ==============================
    if (leader && ptx::elect_one()) {
      constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(256, 128);
      constexpr uint32_t idesc_w = ptx::umma_idesc_bf16_major(256, 16, true, false);
      int ws = 0;
      uint32_t wph = 0;
      const uint32_t w_base = ptx::smem_u32(smem + Smem::w);
      const uint32_t m_base = ptx::smem_u32(smem + Smem::mtile), dz_base = ptx::smem_u32(smem + Smem::dzt);
      const uint32_t b_base = ptx::smem_u32(smem + Smem::bias);
      // D_w[gw & 1] (16 columns: 8 per CTA) = m(gw)^T-tile x dz^T : 8 K steps of 16 pairs
      auto issue_dw = [&](int gw) {
        const int wb = gw & 1;
        ptx::mbar_wait(&bars[bMFull], gw & 1);
        ptx::mbar_wait(&bars[bDwFree + wb], ((gw >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          ptx::umma_ss_2sm(tmem + kColDw + 16 * wb,
                           ptx::umma_desc_mn_sw128(m_base + (ks >> 2) * 16384 + (ks & 3) * 2048, 8192, 1024),
                           ptx::umma_desc_sw128(dz_base + (ks >> 2) * 2048 + (ks & 3) * 32), idesc_w, ks != 0);
        ptx::tc_commit_2sm(&bars[bMFree], 3);
        ptx::tc_commit_2sm(&bars[bDwFull + wb], 3);
      };
      int g = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int c = 0; c < kChunks; ++c, ++g) {
          const int buf = g & 1;
          ptx::mbar_wait(&bars[bUFree + buf], ((g >> 1) & 1) ^ 1);  // both epilogues drained chunk g - 2
          ptx::tc_fence_after();
          const uint32_t ut = tmem + kColU + 128 * buf;
          // u = [1 0 .. 0] x (b_mid / 2 tile) initialises the accumulator (the producers wrote the ones columns before
          // their first "S chunk full" arrival)
          if (g == 0) {
            ptx::mbar_wait(&bars[bSFull + 0], 0);
            ptx::tc_fence_after();
          }
          ptx::umma_ts_2sm(ut, tmem + kColOne, ptx::umma_desc_sw128(b_base + (c >> 2) * kBiasTileBytes + (c & 3) * 32), idesc1, 0);
          for (int kc = 0; kc < kKChunks; ++kc) {
            if (c == 0) ptx::mbar_wait(&bars[bSFull + kc], it & 1);
            ptx::mbar_wait(&bars[bWFull + ws], wph);
            ptx::tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_ts_2sm(ut, tmem + kColS + 32 * kc + 8 * ks, ptx::umma_desc_sw128(w_base + ws * kWStageBytes + ks * 32),
                               idesc1, 1);
            ptx::tc_commit_2sm(&bars[bWEmpty + ws], 3);
            if (c == kChunks - 1) ptx::tc_commit_2sm(&bars[bSFree + kc], 3);
            if (++ws == kWStages) ws = 0, wph ^= 1;
          }
          ptx::tc_commit_2sm(&bars[bUFull + buf], 3);
          if (g > 0) issue_dw(g - 1);  // queued behind u(g): the tensor pipe never waits for the epilogue
        }
      }
      if (g > 0) issue_dw(g - 1);
    }
  } else if (warp >= 4 && warp < kProdWarp0) {
    // ============================== epilogue ==============================
    // 16 warps: quadrant q = warp % 4 (TMEM lanes 32 q ..), column slice csel = (warp - 4) / 4 (32 of the chunk's 128
    // columns).  Four warps per scheduler: the chain TMEM load -> MUFU -> FMA -> pack -> smem of one warp is latency-
    // bound (0.28 IPC measured with two warps per scheduler), more warps in flight hide it.
    const int q = warp % 4, wsel = (warp - 4) / 4;  // this warp's slices: csel = wsel * kSlices + sl
    const int csel0 = wsel * kSlices;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int row = q * 32 + lane;
    // this lane's m row in the MN-major operand tile: box (pair half, feature half), 128 B per pair row
    unsigned char* mt0 = smem + Smem::mtile + (q >> 1) * 2 * 8192 + (row & 63) * 128;
    unsigned char* dzt = smem + Smem::dzt + (q >> 1) * 2048 + (row & 7) * 2;  // K block of this pair, byte inside a chunk
    const int dz_ch = (row & 63) >> 3;                                         // logical 16-byte chunk of this pair
    float* part = a.dwout_part + static_cast<size_t>(blockIdx.x) * (3 * kLdG) + q * 32 + lane;
    unsigned char* ob = smem + Smem::out + (warp - 4) * kOutWarpBytes;
    // D_w of chunk gw (TMEM lanes = features; this CTA's 16 of the 32 columns): add into this CTA's partial sums
    auto flush_dw = [&](int gw) {
      const int wb = gw & 1;
      ptx::mbar_wait(&bars[bDwFull + wb], (gw >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t d4[4];
      ptx::tmem_ld_x4(tmem + lane_base + kColDw + 16 * wb + 8 * rank, d4);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(&bars[bDwFree + wb]);
      float* dst = part + (gw % kChunks) * 128;
      atomicAdd(dst, __uint_as_float(d4[0]));
      atomicAdd(dst + kLdG, __uint_as_float(d4[1]));
      atomicAdd(dst + 2 * kLdG, __uint_as_float(d4[2]));
    };
    // dz of (tile, head) for this lane's pair.  The global loads are issued one chunk ahead of the first use
    // (fetch_dz: raw dz, or raw logits + tag in the FUSED form) and turned into dz where they are consumed (make_dz),
    // so that their latency never sits in the epilogue's dependency chain.
    auto fetch_dz = [&](int64_t tile_, int k_, float& z0, float& z1, float& z2, int& tg) {
      z0 = z1 = z2 = 0.f, tg = -1;
      const int64_t lr_ = tile_ * 128 + row;
      if (lr_ < a.rows) {
        const int C = head_classes(k_);
        const int64_t gp_ = a.g0 + lr_;
        const float* p = (FUSED ? a.logits[k_] : a.dz[k_]) + gp_ * C;
        z0 = p[0], z1 = p[1];
        if (C == 3) z2 = p[2];
        if (FUSED) {
          const long long t64 = a.tags[k_][gp_];
          tg = (t64 < 0 || t64 >= C) ? 3 : static_cast<int>(t64);  // 3 = target outside [0, C): poison with NaN
        }
      }
    };
    auto make_dz = [&](int k_, float& z0, float& z1, float& z2, int tg) {
      if (!FUSED) return;
      if (tg >= 0) {  // dz = scale_k w[t] (softmax(z) - onehot(t))
        const int C = head_classes(k_);
        const float x2 = C == 3 ? z2 : -INFINITY;
        const float mx = fmaxf(fmaxf(z0, z1), x2);
        const float e0 = __expf(z0 - mx), e1 = __expf(z1 - mx), e2 = C == 3 ? __expf(x2 - mx) : 0.f;
        const float sw = s_scale[k_] * s_cw[tg == 3 ? 0 : tg];
        const float gsc = tg == 3 ? NAN : __fdividef(sw, e0 + e1 + e2);
        z0 = gsc * e0 - (tg == 0 ? sw : 0.f);
        z1 = gsc * e1 - (tg == 1 ? sw : 0.f);
        z2 = C == 3 ? gsc * e2 - (tg == 2 ? sw : 0.f) : 0.f;
      }
      if (wsel == 0) {  // db_out[k] += sum over pairs of dz (one of the warps that hold this row)
        float r0 = z0, r1 = z1, r2 = z2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          r0 += __shfl_xor_sync(0xffffffffu, r0, o), r1 += __shfl_xor_sync(0xffffffffu, r1, o);
          r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        }
        if (lane == 0) atomicAdd(&s_dbout[k_ * 4], r0), atomicAdd(&s_dbout[k_ * 4 + 1], r1), atomicAdd(&s_dbout[k_ * 4 + 2], r2);
      }
    };
    int g = 0;
    float nz0 = 0.f, nz1 = 0.f, nz2 = 0.f;
    int ntg = -1;
    if (my_tiles > 0) fetch_dz(tile_of(0), 0, nz0, nz1, nz2, ntg);
    for (int it = 0; it < my_tiles; ++it) {
      const int64_t tile = tile_of(it);
      const int64_t lr = tile * 128 + row;          // row inside the chunk
      const int64_t gp = a.g0 + lr;                 // flat pair index in the batch
      float dz0 = 0.f, dz1 = 0.f, dz2 = 0.f;
      for (int c = 0; c < kChunks; ++c, ++g) {
        const int buf = g & 1, k = c / 3;
        if (c - 3 * k == 0) {
          dz0 = nz0, dz1 = nz1, dz2 = nz2;
          make_dz(k, dz0, dz1, dz2, ntg);
        }
        if (c - 3 * k == 2) {  // prefetch for the next head (or head 0 of this CTA's next tile)
          if (k + 1 < kNumHeads) fetch_dz(tile, k + 1, nz0, nz1, nz2, ntg);
          else if (it + 1 < my_tiles) fetch_dz(tile_of(it + 1), 0, nz0, nz1, nz2, ntg);
          else nz0 = nz1 = nz2 = 0.f, ntg = -1;
        }
        if (wsel == kEpiWarps / 4 - 1 && g >= 2) flush_dw(g - 2);
        if (wsel == 0 && c - 3 * k == 0) {
          // new head: dz^T (B operand of the dW_out MMA), once the MMA of the previous chunk has read the old one
          if (g > 0) ptx::mbar_wait(&bars[bMFree], (g - 1) & 1);
          const uint32_t d01 = ptx::pack_bf16x2(dz0, dz1), d2 = ptx::pack_bf16x2(dz2, 0.f);
          // row c of the K-major tile: 128 B per row, 16-byte chunk index XOR (c & 7)
          *reinterpret_cast<uint16_t*>(dzt + 0 * 128 + ((dz_ch ^ 0) * 16)) = static_cast<uint16_t>(d01);
          *reinterpret_cast<uint16_t*>(dzt + 1 * 128 + ((dz_ch ^ 1) * 16)) = static_cast<uint16_t>(d01 >> 16);
          *reinterpret_cast<uint16_t*>(dzt + 2 * 128 + ((dz_ch ^ 2) * 16)) = static_cast<uint16_t>(d2);
        }
        ptx::mbar_wait(&bars[bUFull + buf], (g >> 1) & 1);
        ptx::tc_fence_after();
        // dz of this pair as bf16x2 broadcasts: g_m = dz W_out runs on packed bf16 FMAs, two features at a time
        const uint32_t dzb0 = ptx::pack_bf16x2(dz0, dz0), dzb1 = ptx::pack_bf16x2(dz1, dz1), dzb2 = ptx::pack_bf16x2(dz2, dz2);
#pragma unroll
        for (int sl = 0; sl < kSlices; ++sl) {
          const int csel = csel0 + sl;
          const uint32_t ut = tmem + lane_base + kColU + 128 * buf + 32 * csel;
          const int f0 = c * 128 + 32 * csel;  // column in the stacked [0, 1920) feature space
          const uint32_t* wp = s_wout + f0 / 2;  // feature pairs, indexed by the stacked feature index k * 384 + f
          unsigned char* mt = mt0 + (csel >> 1) * 8192;
          uint32_t r[32];
          ptx::tmem_ld_x32(ut, r);
          ptx::tmem_ld_wait();
          if (sl == kSlices - 1) {  // accumulator slices drained: release them to the MMA warp before the math / stores
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(&bars[bUFree + buf]);
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            // W_out of the 8 features of this group: 16-byte broadcast loads
            const uint4 wq0 = *reinterpret_cast<const uint4*>(wp + 4 * v);
            const uint4 wq1 = *reinterpret_cast<const uint4*>(wp + 5 * D / 2 + 4 * v);
            const uint4 wq2 = *reinterpret_cast<const uint4*>(wp + 5 * D + 4 * v);
            const uint32_t w0a[4] = {wq0.x, wq0.y, wq0.z, wq0.w}, w1a[4] = {wq1.x, wq1.y, wq1.z, wq1.w},
                           w2a[4] = {wq2.x, wq2.y, wq2.z, wq2.w};
            uint32_t gq[4], mq[4];
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              // tanh in fp32 (one MUFU each), everything after it as packed bf16x2 FMAs: m, SiLU' and g leave as bf16
              // anyway, and the epilogue is bound by the number of instructions it issues
              const float h0 = __uint_as_float(r[8 * v + 2 * y]);  // (u + b_mid) / 2: the bias came through the MMA
              const float h1 = __uint_as_float(r[8 * v + 2 * y + 1]);
              const uint32_t t2 = ptx::pack_bf16x2(ptx::tanh_approx(h0), ptx::tanh_approx(h1));
              const uint32_t h2 = ptx::pack_bf16x2(h0, h1);
              uint32_t m2 = ptx::hfma2_bf16(h2, t2, h2);                     // SiLU(u) = h (1 + tanh h)
              const uint32_t sg2 = ptx::hfma2_bf16(0x3F003F00u, t2, 0x3F003F00u);  // sigmoid(u) = 0.5 + 0.5 tanh h
              const uint32_t oms2 = ptx::hfma2_bf16(0xBF00BF00u, t2, 0x3F003F00u); // 1 - sigmoid(u)
              uint32_t dv2 = ptx::hfma2_bf16(m2, oms2, sg2);                 // SiLU'(u) = sg + m (1 - sg)
              if (DROP) {  // m_dropped = m * mask / (1 - p) feeds W_out; its gradient carries the same factor
                const uint32_t col = (c - 3 * k) * 128 + 32 * csel + 8 * v + 2 * y;
                const float ms0 = drop_keep(a.drop_key[k], a.drop_thresh, static_cast<uint32_t>(gp), col) ? a.drop_scale : 0.f;
                const float ms1 = drop_keep(a.drop_key[k], a.drop_thresh, static_cast<uint32_t>(gp), col + 1) ? a.drop_scale : 0.f;
                const uint32_t ms2 = ptx::pack_bf16x2(ms0, ms1);
                m2 = ptx::hmul2_bf16(m2, ms2), dv2 = ptx::hmul2_bf16(dv2, ms2);
              }
              mq[y] = m2;
              const uint32_t gm = ptx::hfma2_bf16(dzb2, w2a[y], ptx::hfma2_bf16(dzb1, w1a[y], ptx::hmul2_bf16(dzb0, w0a[y])));
              gq[y] = ptx::hmul2_bf16(gm, dv2);
            }
            // G staging tile: row pitch 64 B, 16-byte chunk index XOR ((row >> 1) & 3): conflict-free writes and reads
            *reinterpret_cast<uint4*>(ob + lane * 64 + ((v ^ ((lane >> 1) & 3)) * 16)) = make_uint4(gq[0], gq[1], gq[2], gq[3]);
            // m into the MN-major operand tile (rows past the chunk end: s = 0, so m = SiLU(b_mid) is finite, and dz = 0).
            // The dW_out MMA of the previous chunk must have read the tile: it is queued behind this chunk's u MMAs, so
            // the wait sits after the first group's math, where it costs nothing.
            if (sl == 0 && v == 0 && g > 0) {
              ptx::mbar_wait(&bars[bMFree], (g - 1) & 1);
              ptx::tc_fence_after();
            }
            *reinterpret_cast<uint4*>(mt + (((4 * (csel & 1) + v) ^ (row & 7)) * 16)) = make_uint4(mq[0], mq[1], mq[2], mq[3]);
          }
          if (sl == kSlices - 1) {
            ptx::fence_proxy_async();  // generic-proxy writes (m tile, dz^T) -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) arrive_leader(&bars[bMFull]);
          } else {
            __syncwarp();
          }
          {
            // G to global memory through the staging tile: one store instruction covers 8 rows x 64 B, full sectors
            __nv_bfloat16* dstm = a.G + f0;
            const int64_t row0 = tile * 128 + q * 32;  // first chunk-row of this warp
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = 8 * i + (lane >> 2), c16 = lane & 3;
              const uint4 val = *reinterpret_cast<const uint4*>(ob + rr * 64 + ((c16 ^ ((rr >> 1) & 3)) * 16));
              if (row0 + rr < a.rows) *reinterpret_cast<uint4*>(dstm + (row0 + rr) * kLdG + c16 * 8) = val;
            }
            __syncwarp();  // the staging tile is free again
          }
        }
      }
    }
    if (wsel == kEpiWarps / 4 - 1) {
      if (g >= 2) flush_dw(g - 2);
      if (g >= 1) flush_dw(g - 1);
    }
  } else if (warp >= kProdWarp0) {
    // ============================== pair producers ==============================
    const int q = warp - kProdWarp0;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    unsigned char* stg = smem + Smem::stage;
    {  // the constant K step of the A operand: element 0 = 1.0 (bf16), the other 15 = 0; 16 columns, written once
      uint32_t one[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) one[x] = x == 0 ? 0x00003F80u : 0u;
      ptx::tmem_st_x16(tmem + lane_base + kColOne, one);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();  // ordered before this warp's first "S chunk full" arrival, which the MMA warp waits for
    }
    for (int it = 0; it < my_tiles; ++it) {
      const int64_t tile = tile_of(it);
      int64_t my_a = -1, my_b = -1;
      const int64_t my_lr = tile * 128 + q * 32 + lane;
      if (my_lr < a.rows) {
        const int64_t gp = a.g0 + my_lr;
        const int64_t b = gp / a.pairs_per_doc;
        const int p = static_cast<int>(gp - b * a.pairs_per_doc);
        int i, j;
        pair_from_flat(p, a.n, i, j);
        my_a = (b * a.n + i) * (2 * D);
        my_b = (b * a.n + j) * (2 * D) + D;
      }
      // ---- generate s rows [32q, 32q+32) into staging; lane = 4-column group (3 groups per lane)
#pragma unroll 1
      for (int rr = 0; rr < 32; rr += kProdRows) {
        uint2 av[kProdRows][3], bv[kProdRows][3];
        int64_t offa[kProdRows];
#pragma unroll
        for (int u = 0; u < kProdRows; ++u) {
          offa[u] = __shfl_sync(0xffffffffu, my_a, rr + u);
          const int64_t offb = __shfl_sync(0xffffffffu, my_b, rr + u);
#pragma unroll
          for (int mth = 0; mth < 3; ++mth) {
            const int col = 4 * (lane + 32 * mth);
            if (offa[u] >= 0) {
              av[u][mth] = __ldg(reinterpret_cast<const uint2*>(a.ab + offa[u] + col));
              bv[u][mth] = __ldg(reinterpret_cast<const uint2*>(a.ab + offb + col));
            } else {
              av[u][mth] = make_uint2(0u, 0u), bv[u][mth] = make_uint2(0u, 0u);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kProdRows; ++u) {
          const int r = q * 32 + rr + u;
#pragma unroll
          for (int mth = 0; mth < 3; ++mth) {
            const int cg = lane + 32 * mth;  // 4-column group index, 0..95
            const float a0 = __uint_as_float(av[u][mth].x << 16), a1 = __uint_as_float(av[u][mth].x & 0xFFFF0000u);
            const float a2 = __uint_as_float(av[u][mth].y << 16), a3 = __uint_as_float(av[u][mth].y & 0xFFFF0000u);
            const float b0 = __uint_as_float(bv[u][mth].x << 16), b1 = __uint_as_float(bv[u][mth].x & 0xFFFF0000u);
            const float b2 = __uint_as_float(bv[u][mth].y << 16), b3 = __uint_as_float(bv[u][mth].y & 0xFFFF0000u);
            uint2 o;
            o.x = ptx::pack_bf16x2(ptx::silu_from_half(a0 + b0), ptx::silu_from_half(a1 + b1));
            o.y = ptx::pack_bf16x2(ptx::silu_from_half(a2 + b2), ptx::silu_from_half(a3 + b3));
            const int chunk16 = (cg >> 1) ^ (r & 7);
            *reinterpret_cast<uint2*>(stg + r * kStageRowBytes + chunk16 * 16 + (cg & 1) * 8) = o;
          }
        }
      }
      __syncwarp();
      // ---- S matrix to global memory: the warp's 32 rows are 24 KB contiguous in S [rows, 384]
      {
        const int64_t row0 = tile * 128 + q * 32;
#pragma unroll 4
        for (int itc = 0; itc < 48; ++itc) {
          const int id = itc * 32 + lane, rr = id / 48, c16 = id - rr * 48;
          const int r2 = q * 32 + rr;
          const uint4 t = *reinterpret_cast<const uint4*>(stg + r2 * kStageRowBytes + ((c16 ^ (r2 & 7)) * 16));
          if (row0 + rr < a.rows) *reinterpret_cast<uint4*>(a.S + (row0 + rr) * D + c16 * 8) = t;
        }
      }
      // ---- copy into TMEM, one 64-feature K chunk at a time, as the MMA warp releases them
      const int r = q * 32 + lane;
#pragma unroll 1
      for (int kc = 0; kc < kKChunks; ++kc) {
        uint32_t v[32];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const int chunk16 = (kc * 8 + ch) ^ (r & 7);
          const uint4 t = *reinterpret_cast<const uint4*>(stg + r * kStageRowBytes + chunk16 * 16);
          v[4 * ch] = t.x, v[4 * ch + 1] = t.y, v[4 * ch + 2] = t.z, v[4 * ch + 3] = t.w;
        }
        if (it > 0) ptx::mbar_wait(&bars[bSFree + kc], (it - 1) & 1);
        ptx::tc_fence_after();
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) lo[x] = v[x], hi[x] = v[16 + x];
        ptx::tmem_st_x16(tmem + lane_base + kColS + 32 * kc, lo);
        ptx::tmem_st_x16(tmem + lane_base + kColS + 32 * kc + 16, hi);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(&bars[bSFull + kc]);
      }
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // the peer's tensor-memory reads / remote arrives are finished as well
  if (FUSED && threadIdx.x < 20) {
    const int k = threadIdx.x >> 2, c = threadIdx.x & 3;
    if (c < head_classes(k) && s_dbout[threadIdx.x] != 0.f) atomicAdd(&a.dbout[k][c], s_dbout[threadIdx.x]);
  }
  if (warp == 2) ptx::tmem_dealloc_2sm(tmem, 512);
}

}  // namespace t1p

int launch_pair_bwd_prep_pair(const void* pack, const PackLayout& L, const __nv_bfloat16* ab, int n, int64_t g0, int rows,
                              const float* const dz[kNumHeads], const FusedLossBwd* fused, __nv_bfloat16* S,
                              __nv_bfloat16* G, float* dwout_part, cudaStream_t st, const DropSpec* drop) {
  using namespace t1p;
  const char* base = static_cast<const char*>(pack);
  Args a{};
  a.ab = ab;
  a.bmid_half = reinterpret_cast<const float*>(base + L.bmid_half);
  a.wout4 = reinterpret_cast<const float4*>(base + L.wout_f32x4);
  if (fused) {
    for (int h = 0; h < kNumHeads; ++h) {
      a.logits[h] = fused->logits[h], a.tags[h] = fused->tags[h], a.ratio[h] = fused->ratio[h], a.dbout[h] = fused->dbout[h];
    }
    for (int c = 0; c < 3; ++c) a.class_w[c] = fused->class_w[c];
    a.grad_out6 = fused->grad_out6, a.loss_final = fused->loss_final;
  } else {
    for (int h = 0; h < kNumHeads; ++h) a.dz[h] = dz[h];
  }
  a.S = S, a.G = G, a.dwout_part = dwout_part;
  a.n = n;
  a.pairs_per_doc = static_cast<int32_t>(pair_count(n));
  a.g0 = g0, a.rows = rows;
  a.num_tiles = (rows + 127) / 128;
  if (drop && drop->thresh) {
    a.drop_thresh = drop->thresh, a.drop_scale = drop->scale;
    for (int h = 0; h < kNumHeads; ++h) a.drop_key[h] = drop_key(*drop, site_head(h, 0));
  }
  if (rows == 0) return PENEO_OK;
  alignas(64) CUtensorMap tmW;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmW, base + L.wmid_bf16, D, 5 * D, D * 2, 64, 64)) != PENEO_OK) return rc;
  int dev = 0, sms = 148;
  PENEO_CUDA_TRY(cudaGetDevice(&dev));
  PENEO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int pairs = (a.num_tiles + 1) / 2;
  const int grid = 2 * std::min(pairs, sms / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmemBytes, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  const bool dr = a.drop_thresh != 0;
#define GO(DR, FU)                                                                                                   \
  do {                                                                                                               \
    PENEO_CUDA_TRY(cudaFuncSetAttribute(pair_bwd_prep_pair_kernel<DR, FU>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)); \
    PENEO_CUDA_TRY(cudaLaunchKernelEx(&cfg, pair_bwd_prep_pair_kernel<DR, FU>, tmW, a));                             \
  } while (0)
  if (dr && fused) GO(true, true);
  else if (dr) GO(true, false);
  else if (fused) GO(false, true);
  else GO(false, false);
#undef GO
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
