// T1 — first kernel of the bf16 backward pass (PENEO_PREC_BF16, d = 384, num_layers = 2): the forward
// structure of K2 (pair_heads_tc.cu) with a gradient epilogue.  For every pair of a chunk [g0, g0 + rows) of
// the batch-flattened pair list and every head k:
//   s   = SiLU(a_i + b_j)                         regenerated (CUDA cores -> TMEM), also stored: S  [rows, 384]
//   u_k = W_mid,k s + b_mid,k                     tcgen05, 15 chunks of 128 mid features
//   m_k = SiLU(u_k)                               never leaves the SM: dW_out,k += dz_k^T m_k on the tensor cores
//   g_k = (dz_k W_out,k) * SiLU'(u_k)             stored: G [rows, 1920]   (dW_mid = G^T S, dS = G W_mid)
// dz_k = d loss / d logits of head k (fp32, from the loss backward).  S and G feed the GEMMs of train.cu.  What
// autograd keeps alive between forward and backward in the reference ([P, D] activations of every layer) is
// recomputed instead.
//
// dW_out: the epilogue writes the bf16 m tile of a chunk ([128 pairs x 128 features]) into shared memory in the
// MN-major SWIZZLE_128B UMMA layout (A operand: M = feature, K = pair) and dz^T ([16 x 128 pairs], K-major, rows
// 3..15 zero) as the B operand; eight N = 16 MMAs give D_w [128 features x 16] in TMEM, which four epilogue warps
// add (RED, no contention) into this CTA's private [3][1920] partial sums; train.cu reduces the partials.
//
// Warp roles as in K2: warp 0 TMA (W_mid stream), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 epilogue,
// warps 12-15 pair producers.
// TMEM columns: [0,192) s | [192,320) u buffer 0 | [320,448) u buffer 1 | [448,464) D_w 0 | [464,480) D_w 1
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {
namespace t1 {

constexpr int D = 384;
constexpr int kChunks = 15;       // 5 heads x 3 chunks of 128 mid features
constexpr int kKChunks = 6;       // 384 / 64
constexpr int kWStages = 3;
constexpr int kWStageBytes = 128 * 64 * 2;  // 16 KB
constexpr int kStageRowBytes = D * 2;       // staging: 128 rows x 768 B
constexpr int kThreads = 512;
constexpr int kLdG = 5 * D;                 // row stride of G / M

constexpr uint32_t kColS = 0, kColU = 192, kColDw = 448;
constexpr int kOutWarpBytes = 32 * 64;  // per epilogue warp: 32 rows x 32 bf16, XOR-swizzled 16-byte chunks

struct Smem {
  static constexpr int w = 0;
  static constexpr int stage = w + kWStages * kWStageBytes;
  static constexpr int mtile = stage + 128 * kStageRowBytes;  // m chunk, 2 x 2 boxes of [64 pairs x 64 features] bf16
  static constexpr int dzt = mtile + 128 * 128 * 2;           // dz^T, 2 K blocks of [16 x 64 pairs] bf16
  static constexpr int bmid = dzt + 2 * 16 * 128;             // 1920 floats
  static constexpr int out = bmid + 5 * D * 4;               // 8 epilogue warps x [32 rows][64 B]
  static constexpr int wout = out + 8 * kOutWarpBytes;       // [3][960] bf16x2: W_out[c] of feature pairs (2f, 2f+1)
  static constexpr int bars = wout + 3 * (5 * D / 2) * 4;
  static constexpr int total = bars + 512;
};
constexpr int bWFull = 0, bWEmpty = bWFull + kWStages, bUFull = bWEmpty + kWStages, bUFree = bUFull + 2,
              bSFull = bUFree + 2, bSFree = bSFull + kKChunks, bMFull = bSFree + kKChunks, bMFree = bMFull + 1,
              bDwFull = bMFree + 1, bDwFree = bDwFull + 2, bCount = bDwFree + 2;
static_assert(Smem::mtile % 1024 == 0 && Smem::dzt % 1024 == 0, "UMMA operand tiles need 1024-byte alignment");
static_assert(Smem::total + 1024 <= 227 * 1024, "shared memory budget");
static_assert(bCount * 8 + 16 <= 512, "barrier area too small");
constexpr int kSmemBytes = Smem::total + 1024;

struct Args {
  const __nv_bfloat16* ab;   // [batch*n, 768] : 0.5*A | 0.5*Bm
  const float* bmid_half;    // [1920]
  const float4* wout4;       // [5][384] : (W_out[0][f], W_out[1][f], W_out[2][f] or 0, 0)
  const float* dz[kNumHeads];  // d loss / d logits, fp32 [batch*P, C_h]
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

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 1) pair_bwd_prep_kernel(const __grid_constant__ CUtensorMap tmW, const Args a) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by offset (not by integer round trip): the compiler keeps the shared address space -> LDS / STS
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  float* s_bmid = reinterpret_cast<float*>(smem + Smem::bmid);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < kWStages; ++s) ptx::mbar_init(&bars[bWFull + s], 1), ptx::mbar_init(&bars[bWEmpty + s], 1);
    for (int s = 0; s < 2; ++s) ptx::mbar_init(&bars[bUFull + s], 1), ptx::mbar_init(&bars[bUFree + s], 8);
    for (int s = 0; s < kKChunks; ++s) ptx::mbar_init(&bars[bSFull + s], 4), ptx::mbar_init(&bars[bSFree + s], 1);
    ptx::mbar_init(&bars[bMFull], 8), ptx::mbar_init(&bars[bMFree], 1);
    for (int s = 0; s < 2; ++s) ptx::mbar_init(&bars[bDwFull + s], 1), ptx::mbar_init(&bars[bDwFree + s], 4);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  for (int e = threadIdx.x; e < 5 * D; e += kThreads) s_bmid[e] = a.bmid_half[e];
  uint32_t* s_wout = reinterpret_cast<uint32_t*>(smem + Smem::wout);
  for (int e = threadIdx.x; e < 5 * D / 2; e += kThreads) {
    const float4 w0 = a.wout4[2 * e], w1 = a.wout4[2 * e + 1];
    s_wout[e] = ptx::pack_bf16x2(w0.x, w1.x);
    s_wout[5 * D / 2 + e] = ptx::pack_bf16x2(w0.y, w1.y);
    s_wout[5 * D + e] = ptx::pack_bf16x2(w0.z, w1.z);
  }
  for (int e = threadIdx.x; e < 2 * 16 * 128 / 16; e += kThreads)  // dz^T rows 3..15 stay zero for the whole kernel
    reinterpret_cast<uint4*>(smem + Smem::dzt)[e] = make_uint4(0u, 0u, 0u, 0u);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int my_tiles = (a.num_tiles > static_cast<int>(blockIdx.x))
                           ? (a.num_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1
                           : 0;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (ptx::elect_one()) {
      int ws = 0;
      uint32_t wph = 0;
      for (int it = 0; it < my_tiles; ++it)
        for (int c = 0; c < kChunks; ++c)
          for (int kc = 0; kc < kKChunks; ++kc) {
            ptx::mbar_wait(&bars[bWEmpty + ws], wph ^ 1);
            ptx::mbar_arrive_expect_tx(&bars[bWFull + ws], kWStageBytes);
            ptx::tma_load_2d(smem + Smem::w + ws * kWStageBytes, &tmW, &bars[bWFull + ws], kc * 64, c * 128);
            if (++ws == kWStages) ws = 0, wph ^= 1;
          }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(128, 128);
      int ws = 0;
      uint32_t wph = 0;
      const uint32_t w_base = ptx::smem_u32(smem + Smem::w);
      constexpr uint32_t idesc_w = ptx::umma_idesc_bf16_major(128, 16, true, false);
      const uint32_t m_base = ptx::smem_u32(smem + Smem::mtile), dz_base = ptx::smem_u32(smem + Smem::dzt);
      // D_w[g & 1] = m(g)^T-tile x dz^T : 8 K steps of 16 pairs
      auto issue_dw = [&](int gw) {
        const int wb = gw & 1;
        ptx::mbar_wait(&bars[bMFull], gw & 1);
        ptx::mbar_wait(&bars[bDwFree + wb], ((gw >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          ptx::umma_ss(tmem + kColDw + 16 * wb,
                       ptx::umma_desc_mn_sw128(m_base + (ks >> 2) * 16384 + (ks & 3) * 2048, 8192, 1024),
                       ptx::umma_desc_sw128(dz_base + (ks >> 2) * 2048 + (ks & 3) * 32), idesc_w, ks != 0);
        ptx::tc_commit(&bars[bMFree]);
        ptx::tc_commit(&bars[bDwFull + wb]);
      };
      int g = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int c = 0; c < kChunks; ++c, ++g) {
          const int buf = g & 1;
          // the epilogue has drained this accumulator (chunk g - 2)
          ptx::mbar_wait(&bars[bUFree + buf], ((g >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t ut = tmem + kColU + 128 * buf;
          for (int kc = 0; kc < kKChunks; ++kc) {
            if (c == 0) ptx::mbar_wait(&bars[bSFull + kc], it & 1);
            ptx::mbar_wait(&bars[bWFull + ws], wph);
            ptx::tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_ts(ut, tmem + kColS + 32 * kc + 8 * ks, ptx::umma_desc_sw128(w_base + ws * kWStageBytes + ks * 32),
                           idesc1, (kc | ks) != 0);
            ptx::tc_commit(&bars[bWEmpty + ws]);
            if (c == kChunks - 1) ptx::tc_commit(&bars[bSFree + kc]);
            if (++ws == kWStages) ws = 0, wph ^= 1;
          }
          ptx::tc_commit(&bars[bUFull + buf]);
          if (g > 0) issue_dw(g - 1);  // queued behind u(g): the tensor pipe never waits for the epilogue
        }
      }
      if (g > 0) issue_dw(g - 1);
    }
  } else if (warp >= 4 && warp < 12) {
    // ============================== epilogue ==============================
    const int q = warp % 4, hsel = (warp - 4) / 4;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int row = q * 32 + lane;
    unsigned char* mt = smem + Smem::mtile + ((q >> 1) * 2 + hsel) * 8192 + (row & 63) * 128;  // this lane's m row
    unsigned char* dzt = smem + Smem::dzt + (q >> 1) * 2048 + (row & 7) * 2;  // K block of this pair, byte inside a chunk
    const int dz_ch = (row & 63) >> 3;                                         // logical 16-byte chunk of this pair
    float* part = a.dwout_part + static_cast<size_t>(blockIdx.x) * (3 * kLdG) + q * 32 + lane;
    // D_w of chunk gw (TMEM lanes = features): add into this CTA's partial sums (hsel == 1 warps, one lane quarter each)
    auto flush_dw = [&](int gw) {
      const int wb = gw & 1;
      ptx::mbar_wait(&bars[bDwFull + wb], (gw >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t d4[4];
      ptx::tmem_ld_x4(tmem + lane_base + kColDw + 16 * wb, d4);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars[bDwFree + wb]);
      float* dst = part + (gw % kChunks) * 128;
      atomicAdd(dst, __uint_as_float(d4[0]));
      atomicAdd(dst + kLdG, __uint_as_float(d4[1]));
      atomicAdd(dst + 2 * kLdG, __uint_as_float(d4[2]));
    };
    // dz of (tile, head) for this lane's pair; loaded one chunk ahead of its first use so the latency is hidden
    auto load_dz = [&](int64_t tile_, int k_, float& z0, float& z1, float& z2) {
      z0 = z1 = z2 = 0.f;
      const int64_t lr_ = tile_ * 128 + row;
      if (lr_ < a.rows) {
        const int C = head_classes(k_);
        const float* p = a.dz[k_] + (a.g0 + lr_) * C;
        z0 = p[0], z1 = p[1];
        if (C == 3) z2 = p[2];
      }
    };
    int g = 0;
    float nz0 = 0.f, nz1 = 0.f, nz2 = 0.f;
    if (my_tiles > 0) load_dz(blockIdx.x, 0, nz0, nz1, nz2);
    for (int it = 0; it < my_tiles; ++it) {
      const int64_t tile = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x;
      const int64_t lr = tile * 128 + row;          // row inside the chunk
      const int64_t gp = a.g0 + lr;                 // flat pair index in the batch
      float dz0 = 0.f, dz1 = 0.f, dz2 = 0.f;
      for (int c = 0; c < kChunks; ++c, ++g) {
        const int buf = g & 1, k = c / 3;
        if (c - 3 * k == 0) dz0 = nz0, dz1 = nz1, dz2 = nz2;
        if (c - 3 * k == 2) {  // prefetch for the next head (or head 0 of this CTA's next tile)
          if (k + 1 < kNumHeads) load_dz(tile, k + 1, nz0, nz1, nz2);
          else if (it + 1 < my_tiles) load_dz(tile + gridDim.x, 0, nz0, nz1, nz2);
        }
        if (hsel == 1 && g >= 2) flush_dw(g - 2);
        if (hsel == 0 && c - 3 * k == 0) {
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
        const uint32_t ut = tmem + lane_base + kColU + 128 * buf + 64 * hsel;
        const int f0 = c * 128 + 64 * hsel;  // column in the stacked [0, 1920) feature space
        const float* hb = s_bmid + f0;
        const uint32_t* wp = s_wout + f0 / 2;  // feature pairs, indexed by the stacked feature index k * 384 + f
        // dz of this pair as bf16x2 broadcasts: g_m = dz W_out runs on packed bf16 FMAs, two features at a time
        const uint32_t dzb0 = ptx::pack_bf16x2(dz0, dz0), dzb1 = ptx::pack_bf16x2(dz1, dz1), dzb2 = ptx::pack_bf16x2(dz2, dz2);
#pragma unroll
        for (int piece = 0; piece < 2; ++piece) {
          uint32_t r[32];
          ptx::tmem_ld_x32(ut + 32 * piece, r);
          float4 hb4[8];  // this piece's biases, fetched while the TMEM load is in flight
#pragma unroll
          for (int v = 0; v < 8; ++v) hb4[v] = *reinterpret_cast<const float4*>(hb + 32 * piece + 4 * v);
          ptx::tmem_ld_wait();
          if (piece == 1) {  // accumulator drained: release it to the MMA warp before the math / stores
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars[bUFree + buf]);
          }
          // G goes to global memory through a per-warp smem tile (lane = row in TMEM, so each lane holds 64 B of
          // its own row; the tile lets one store instruction write 8 rows x 64 B, full sectors).  The tile is
          // filled 16 B at a time as the values are produced, which keeps few of them live in registers.
          unsigned char* ob = smem + Smem::out + (warp - 4) * kOutWarpBytes;
          uint32_t mp[16];
          __syncwarp();  // the read-back of the previous piece is complete
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t gq[4];
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              const int x = 8 * v + 2 * y;
              float mv[2], dv[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int col = 32 * piece + x + e;
                const float4 hq = hb4[(x + e) / 4];
                const float hbv = ((x + e) % 4 == 0) ? hq.x : ((x + e) % 4 == 1) ? hq.y : ((x + e) % 4 == 2) ? hq.z : hq.w;
                const float h = __uint_as_float(r[x + e]) + hbv;  // u / 2
                const float t = ptx::tanh_approx(h);
                mv[e] = fmaf(h, t, h);                                  // SiLU(u) = u sigmoid(u) = h (1 + tanh h)
                const float sg = fmaf(0.5f, t, 0.5f), oms = fmaf(-0.5f, t, 0.5f);
                dv[e] = fmaf(mv[e], oms, sg);                           // SiLU'(u) = sg + m (1 - sg)
                if (DROP) {  // m_dropped = m * mask / (1 - p) feeds W_out; its gradient carries the same factor
                  const float ms = drop_keep(a.drop_key[k], a.drop_thresh, static_cast<uint32_t>(gp),
                                             (c - 3 * k) * 128 + 64 * hsel + col) ? a.drop_scale : 0.f;
                  mv[e] *= ms, dv[e] *= ms;
                }
              }
              mp[x / 2] = ptx::pack_bf16x2(mv[0], mv[1]);
              const int pi = 16 * piece + x / 2;
              const uint32_t gm = ptx::hfma2_bf16(dzb2, wp[5 * D + pi],
                                                  ptx::hfma2_bf16(dzb1, wp[5 * D / 2 + pi], ptx::hmul2_bf16(dzb0, wp[pi])));
              gq[y] = ptx::hmul2_bf16(gm, ptx::pack_bf16x2(dv[0], dv[1]));
            }
            // row pitch 64 B, 16-byte chunk index XOR ((row >> 1) & 3): conflict-free for these writes and the reads below
            *reinterpret_cast<uint4*>(ob + lane * 64 + ((v ^ ((lane >> 1) & 3)) * 16)) = make_uint4(gq[0], gq[1], gq[2], gq[3]);
          }
          __syncwarp();
          {
            __nv_bfloat16* dstm = a.G + f0 + 32 * piece;
            const int64_t row0 = tile * 128 + q * 32;  // first chunk-row of this warp
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = 8 * i + (lane >> 2), c16 = lane & 3;
              const uint4 val = *reinterpret_cast<const uint4*>(ob + rr * 64 + ((c16 ^ ((rr >> 1) & 3)) * 16));
              if (row0 + rr < a.rows) *reinterpret_cast<uint4*>(dstm + (row0 + rr) * kLdG + c16 * 8) = val;
            }
          }
          // m into the MN-major operand tile (rows past the chunk end: s = 0, so m = SiLU(b_mid) is finite, and dz = 0)
          if (piece == 0 && g > 0) {
            ptx::mbar_wait(&bars[bMFree], (g - 1) & 1);  // the dW_out MMA of the previous chunk has read the tile
            ptx::tc_fence_after();
          }
#pragma unroll
          for (int v = 0; v < 4; ++v)
            *reinterpret_cast<uint4*>(mt + (((4 * piece + v) ^ (row & 7)) * 16)) =
                make_uint4(mp[4 * v], mp[4 * v + 1], mp[4 * v + 2], mp[4 * v + 3]);
        }
        ptx::fence_proxy_async();  // generic-proxy writes (m tile, dz^T) -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars[bMFull]);
      }
    }
    if (hsel == 1) {
      if (g >= 2) flush_dw(g - 2);
      if (g >= 1) flush_dw(g - 1);
    }
  } else if (warp >= 12) {
    // ============================== pair producers ==============================
    const int q = warp - 12;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    unsigned char* stg = smem + Smem::stage;
    for (int it = 0; it < my_tiles; ++it) {
      const int64_t tile = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x;
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
      for (int rr = 0; rr < 32; rr += 4) {
        uint2 av[4][3], bv[4][3];
        int64_t offa[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
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
        for (int u = 0; u < 4; ++u) {
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
      // ---- S matrix to global memory: the warp's 32 rows are 24 KB contiguous in S [rows, 384]; read the
      //      swizzled staging tile 16 B at a time so that every store instruction covers 512 contiguous bytes
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
        if (lane == 0) ptx::mbar_arrive(&bars[bSFull + kc]);
      }
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace t1

int launch_pair_bwd_prep(const void* pack, const PackLayout& L, const __nv_bfloat16* ab, int n, int64_t g0, int rows,
                         const float* const dz[kNumHeads], __nv_bfloat16* S, __nv_bfloat16* G, float* dwout_part,
                         cudaStream_t st, const DropSpec* drop) {
  using namespace t1;
  const char* base = static_cast<const char*>(pack);
  Args a{};
  a.ab = ab;
  a.bmid_half = reinterpret_cast<const float*>(base + L.bmid_half);
  a.wout4 = reinterpret_cast<const float4*>(base + L.wout_f32x4);
  for (int h = 0; h < kNumHeads; ++h) a.dz[h] = dz[h];
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
  if ((rc = make_tensor_map_bf16(&tmW, base + L.wmid_bf16, D, 5 * D, D * 2, 64, 128)) != PENEO_OK) return rc;
  int dev = 0, sms = 148;
  PENEO_CUDA_TRY(cudaGetDevice(&dev));
  PENEO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(a.num_tiles, sms);
  auto kern = a.drop_thresh ? pair_bwd_prep_kernel<true> : pair_bwd_prep_kernel<false>;
  PENEO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  kern<<<grid, kThreads, kSmemBytes, st>>>(tmW, a);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
