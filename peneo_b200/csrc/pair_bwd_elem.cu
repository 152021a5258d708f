// T1e — the hidden-activation backward from SAVED pre-activations (no recompute GEMM).
//
// When the training forward (pair_heads_tc2.cu, SAVE instantiation) has stored h = (u + b_mid) / 2 for every pair and
// head feature (bf16 [pairs, 1920]), the backward needs no tensor-core pass to regenerate it: this kernel walks the rows
// of a chunk once and turns h into G IN PLACE,
//   m = SiLU(2 h) = h (1 + tanh h)          dW_out,k += dz_k^T m          (model/peneo_decoder.py:256-271, backward)
//   G = (dz_k W_out,k) * SiLU'(2 h)
// with dz = scale_k w[t] (softmax(z) - onehot(t)) formed from the stored logits + tags exactly as in T1's FUSED form
// (model/custom_loss.py:189-202) and db_out = sum dz reduced on the way.  HBM-bound: 3.75 KB read + 3.75 KB written per
// pair at full occupancy (T1 is bound by the issue latency of eight epilogue warps next to its MMA pipeline).
//
// Layout of the work: a thread owns 8 consecutive stacked features (16 bytes of a row) of ONE head for every row it
// sees, so W_out of those features lives in registers for the whole kernel and dW_out accumulates in registers without
// any cross-thread reduction; 240 threads cover a row (coalesced 3840-byte reads / writes).  dz of a 32-row tile is
// computed once by 160 threads (head = warp, row = lane) and broadcast through shared memory as bf16x2 pairs.
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {
namespace t1e {

constexpr int D = 384, kF = 5 * D;
constexpr int kWorkers = kF / 8;  // 240
constexpr int kThreads = 256;
constexpr int kTileRows = 32;     // rows per dz tile
constexpr int kSub = 8;           // rows a worker has in flight; dW_out accumulates in bf16x2 over kSub rows, then in fp32

struct Args {
  __nv_bfloat16* hg;  // chunk base: in h, out G; row r of the chunk at hg + r * 1920
  const float* logits[kNumHeads];
  const int64_t* tags[kNumHeads];
  const float* grad_out6;
  const double* loss_final;
  float ratio[kNumHeads], class_w[3];
  float* dbout[kNumHeads];
  const float4* wout4;  // [5][384] : (W_out[0][f], W_out[1][f], W_out[2][f] or 0, 0)
  float* dwout_part;    // [gridDim.x][3][1920], accumulated
  int64_t g0;           // batch-flat index of the chunk's first pair
  int32_t rows, num_tiles;
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
__global__ void __launch_bounds__(kThreads, 2) pair_bwd_elem_kernel(const Args a) {
  __shared__ uint32_t s_dz[kTileRows * kNumHeads * 3];  // bf16x2 (dz, dz) of [row][head][class]
  __shared__ float s_scale[kNumHeads], s_cw[3];
  const int t = threadIdx.x, lane = t % 32, warp = t / 32;
  if (t < kNumHeads)
    s_scale[t] = (a.grad_out6[5] * a.ratio[t] + a.grad_out6[t]) / static_cast<float>(a.loss_final[2 * t + 1]);
  if (t >= 8 && t < 11) s_cw[t - 8] = a.class_w[t - 8];
  const bool worker = t < kWorkers;
  const int f0 = 8 * (worker ? t : 0);  // first stacked feature of this thread
  const int k = f0 / D;                 // its head
  uint32_t w0[4], w1[4], w2[4];         // W_out[c] of the feature pairs (f0 + 2 y, f0 + 2 y + 1) as bf16x2
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    const float4 p0 = a.wout4[f0 + 2 * y], p1 = a.wout4[f0 + 2 * y + 1];
    w0[y] = ptx::pack_bf16x2(p0.x, p1.x), w1[y] = ptx::pack_bf16x2(p0.y, p1.y), w2[y] = ptx::pack_bf16x2(p0.z, p1.z);
  }
  float acc[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
  float db0 = 0.f, db1 = 0.f, db2 = 0.f;  // dz threads (t < 160): head = warp, row = lane
  __syncthreads();

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int64_t r0 = static_cast<int64_t>(tile) * kTileRows;
    // ---- dz of the tile: scale_k w[t] (softmax(z) - onehot(t)), 0 for rows past the chunk end
    if (t < 32 * kNumHeads) {
      const int hd = warp, C = head_classes(hd);
      const int64_t lr = r0 + lane;
      float z0 = 0.f, z1 = 0.f, z2 = 0.f;
      if (lr < a.rows) {
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
      db0 += z0, db1 += z1, db2 += z2;
      uint32_t* dst = s_dz + (lane * kNumHeads + hd) * 3;
      dst[0] = ptx::pack_bf16x2(z0, z0), dst[1] = ptx::pack_bf16x2(z1, z1), dst[2] = ptx::pack_bf16x2(z2, z2);
    }
    __syncthreads();
    if (worker) {
#pragma unroll 1
      for (int rs = 0; rs < kTileRows; rs += kSub) {
        if (r0 + rs >= a.rows) break;
        uint4 hv[kSub];
#pragma unroll
        for (int u = 0; u < kSub; ++u) {
          const int64_t lr = r0 + rs + u;
          hv[u] = lr < a.rows ? *reinterpret_cast<const uint4*>(a.hg + lr * kF + f0) : make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t accb[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int y = 0; y < 4; ++y) accb[c][y] = 0u;
#pragma unroll
        for (int u = 0; u < kSub; ++u) {
          const int64_t lr = r0 + rs + u;
          const uint32_t* dz = s_dz + ((rs + u) * kNumHeads + k) * 3;
          const uint32_t d0 = dz[0], d1 = dz[1], d2 = dz[2];
          const uint32_t hh[4] = {hv[u].x, hv[u].y, hv[u].z, hv[u].w};
          uint32_t gq[4];
#pragma unroll
          for (int y = 0; y < 4; ++y) {
            const uint32_t h2 = hh[y];
            const uint32_t t2 = tanh_bf16x2(h2);
            uint32_t m2 = ptx::hfma2_bf16(h2, t2, h2);                            // SiLU(2 h) = h (1 + tanh h)
            const uint32_t sg2 = ptx::hfma2_bf16(0x3F003F00u, t2, 0x3F003F00u);   // sigmoid(2 h) = 0.5 + 0.5 tanh h
            const uint32_t oms2 = ptx::hfma2_bf16(0xBF00BF00u, t2, 0x3F003F00u);  // 1 - sigmoid
            uint32_t dv2 = ptx::hfma2_bf16(m2, oms2, sg2);                        // SiLU' = sg + m (1 - sg)
            if (DROP) {  // m_dropped = m * mask / (1 - p) feeds W_out; its gradient carries the same factor
              const uint32_t col = static_cast<uint32_t>(f0 - k * D + 2 * y), grow = static_cast<uint32_t>(a.g0 + lr);
              const float ms0 = drop_keep(a.drop_key[k], a.drop_thresh, grow, col) ? a.drop_scale : 0.f;
              const float ms1 = drop_keep(a.drop_key[k], a.drop_thresh, grow, col + 1) ? a.drop_scale : 0.f;
              const uint32_t ms2 = ptx::pack_bf16x2(ms0, ms1);
              m2 = ptx::hmul2_bf16(m2, ms2), dv2 = ptx::hmul2_bf16(dv2, ms2);
            }
            const uint32_t gm = ptx::hfma2_bf16(d2, w2[y], ptx::hfma2_bf16(d1, w1[y], ptx::hmul2_bf16(d0, w0[y])));
            gq[y] = ptx::hmul2_bf16(gm, dv2);
            accb[0][y] = ptx::hfma2_bf16(d0, m2, accb[0][y]);
            accb[1][y] = ptx::hfma2_bf16(d1, m2, accb[1][y]);
            accb[2][y] = ptx::hfma2_bf16(d2, m2, accb[2][y]);
          }
          if (lr < a.rows) *reinterpret_cast<uint4*>(a.hg + lr * kF + f0) = make_uint4(gq[0], gq[1], gq[2], gq[3]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int y = 0; y < 4; ++y) {
            acc[c][2 * y] += __uint_as_float(accb[c][y] << 16);
            acc[c][2 * y + 1] += __uint_as_float(accb[c][y] & 0xFFFF0000u);
          }
      }
    }
    __syncthreads();  // s_dz is rewritten by the next tile
  }

  if (worker) {
    float* part = a.dwout_part + static_cast<size_t>(blockIdx.x) * (3 * kF) + f0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) part[c * kF + e] += acc[c][e];  // this block's private slice
  }
  if (t < 32 * kNumHeads) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      db0 += __shfl_xor_sync(0xffffffffu, db0, o), db1 += __shfl_xor_sync(0xffffffffu, db1, o);
      db2 += __shfl_xor_sync(0xffffffffu, db2, o);
    }
    if (lane == 0) {
      atomicAdd(&a.dbout[warp][0], db0), atomicAdd(&a.dbout[warp][1], db1);
      if (head_classes(warp) == 3) atomicAdd(&a.dbout[warp][2], db2);
    }
  }
}

}  // namespace t1e

int pair_bwd_elem_max_ctas() { return 2 * 160; }  // upper bound of the grid (two CTAs per SM): sizes the dW_out partial buffer

int launch_pair_bwd_elem(const void* pack, const PackLayout& L, int64_t g0, int rows, const FusedLossBwd& fused,
                         __nv_bfloat16* hg, float* dwout_part, int* ctas_out, cudaStream_t st, const DropSpec* drop) {
  using namespace t1e;
  const char* base = static_cast<const char*>(pack);
  Args a{};
  a.hg = hg;
  for (int h = 0; h < kNumHeads; ++h)
    a.logits[h] = fused.logits[h], a.tags[h] = fused.tags[h], a.ratio[h] = fused.ratio[h], a.dbout[h] = fused.dbout[h];
  for (int c = 0; c < 3; ++c) a.class_w[c] = fused.class_w[c];
  a.grad_out6 = fused.grad_out6, a.loss_final = fused.loss_final;
  a.wout4 = reinterpret_cast<const float4*>(base + L.wout_f32x4);
  a.dwout_part = dwout_part;
  a.g0 = g0, a.rows = rows, a.num_tiles = (rows + kTileRows - 1) / kTileRows;
  if (drop && drop->thresh) {
    a.drop_thresh = drop->thresh, a.drop_scale = drop->scale;
    for (int h = 0; h < kNumHeads; ++h) a.drop_key[h] = drop_key(*drop, site_head(h, 0));
  }
  int dev = 0, sms = 148;
  PENEO_CUDA_TRY(cudaGetDevice(&dev));
  PENEO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(std::min(a.num_tiles, 2 * sms), pair_bwd_elem_max_ctas());
  if (ctas_out) *ctas_out = std::min(2 * sms, pair_bwd_elem_max_ctas());
  if (rows == 0) return PENEO_OK;
  if (a.drop_thresh) pair_bwd_elem_kernel<true><<<grid, kThreads, 0, st>>>(a);
  else pair_bwd_elem_kernel<false><<<grid, kThreads, 0, st>>>(a);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
