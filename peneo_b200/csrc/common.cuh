// Shared helpers for the peneo_b200 kernels: error plumbing, pair-index arithmetic, packed
// weight layout.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/peneo_b200.h"

namespace peneo {

constexpr int kNumHeads = PENEO_NUM_HEADS;
__host__ __device__ constexpr int head_classes(int h) { return h == 0 ? 2 : 3; }

// ---- error plumbing (thread-local message, negative status codes) ---------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define PENEO_CUDA_TRY(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      peneo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PENEO_E_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define PENEO_REQUIRE(cond, ...)     \
  do {                               \
    if (!(cond)) {                   \
      peneo::set_error(__VA_ARGS__); \
      return PENEO_E_INVALID;        \
    }                                \
  } while (0)

// ---- pair index arithmetic --------------------------------------------------------------------
// Flat upper-triangular index p(i, j) = i*n - i(i-1)/2 + (j - i), i <= j < n
// (row-major order of torch.nonzero(row <= col), model/peneo_decoder.py:143-147, 171-173).
__host__ __device__ __forceinline__ int64_t pair_count(int64_t n) { return n * (n + 1) / 2; }
__host__ __device__ __forceinline__ int32_t row_start(int32_t i, int32_t n) { return i * n - (i * (i - 1)) / 2; }

__device__ __forceinline__ void pair_from_flat(int32_t p, int32_t n, int32_t& i, int32_t& j) {
  // i = largest row with row_start(i) <= p.  Float estimate, then exact integer fix-up.
  const float t = 2.0f * n + 1.0f;
  float disc = t * t - 8.0f * static_cast<float>(p);
  int32_t r = static_cast<int32_t>((t - sqrtf(fmaxf(disc, 0.0f))) * 0.5f);
  r = max(0, min(r, n - 1));
  while (r > 0 && row_start(r, n) > p) --r;
  while (r + 1 < n && row_start(r + 1, n) <= p) ++r;
  i = r;
  j = r + (p - row_start(r, n));
}

__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// Class index of a target for a C-class head.  The reference raises on a target outside [0, C)
// (F.cross_entropy, model/custom_loss.py:196-210); a kernel cannot raise, so an out-of-range target is clamped to
// class 0 for every memory access and `bad` tells the caller to poison its result with NaN — the loss (and every
// gradient derived from it) then reads NaN instead of silently touching memory outside w[] / the logits row.
__device__ __forceinline__ int checked_tag(long long t, int C, bool& bad) {
  bad = t < 0 || t >= C;
  return bad ? 0 : static_cast<int>(t);
}

// ---- dropout ------------------------------------------------------------------------------------
// The three dropout sites inside the decoder (model/peneo_decoder.py:218, 221, 261) use a counter-based mask so
// that the backward pass can regenerate it instead of storing [P, D] masks: an element is kept iff
//   mix32((row * 0x9E3779B1 + col) ^ key(seed, site)) >= p * 2^32 ,  kept values are scaled by 1 / (1 - p).
// `row` is the token index (per-token sites) or the batch-flat pair index (pair sites); `site` numbers the
// Dropout modules.  The RNG stream cannot match torch's; the mask is reproducible from (seed, site, row, col).
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
struct DropSpec {
  uint32_t thresh;  // 0 = dropout disabled
  float scale;      // 1 / (1 - p)
  uint32_t seed_lo, seed_hi;
};
__host__ __device__ __forceinline__ uint32_t drop_key(const DropSpec& d, uint32_t site) {
  return mix32(d.seed_lo ^ mix32(d.seed_hi + 0x9E3779B9u * (site + 1u)));
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t key, uint32_t thresh, uint32_t row, uint32_t col) {
  return mix32((row * 0x9E3779B1u + col) ^ key) >= thresh;
}
constexpr uint32_t kSiteTok0 = 0, kSiteTok1 = 1;
__host__ __device__ constexpr uint32_t site_head(int head, int layer) { return 16u + 8u * head + layer; }
inline DropSpec make_drop(const peneo_dropout* d) {
  DropSpec s{0u, 1.0f, 0u, 0u};
  if (d && d->p > 0.0f) {
    const double t = static_cast<double>(d->p) * 4294967296.0;
    s.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(t);
    s.scale = d->p < 1.0f ? 1.0f / (1.0f - d->p) : 0.0f;
    s.seed_lo = static_cast<uint32_t>(d->seed), s.seed_hi = static_cast<uint32_t>(d->seed >> 32);
  }
  return s;
}

// ---- packed weights ---------------------------------------------------------------------------
// One device buffer written by peneo_pack_weights; offsets in bytes from its start.
constexpr int kMaxMidLayers = 7;  // num_layers <= 8

struct PackLayout {
  // PENEO_PREC_FP32: fp32 copies in the reference's own [out, in] layouts.
  size_t f_w1, f_b1, f_w2, f_b2, f_wc, f_bc;
  size_t f_mid_w[kNumHeads][kMaxMidLayers], f_mid_b[kNumHeads][kMaxMidLayers];
  size_t f_out_w[kNumHeads], f_out_b[kNumHeads];
  // PENEO_PREC_BF16 (d == 384, num_layers == 2): converted / concatenated / pre-scaled matrices.
  size_t w1_bf16;    // [hid, hin]
  size_t b1;         // [hid] fp32
  size_t w2_bf16;    // [d, hid]
  size_t b2;         // [d] fp32
  size_t wc_bf16;    // [2d, d]   rows [0,d) = 0.5*W_c[:, :d], rows [d,2d) = 0.5*W_c[:, d:]
  size_t bc_half;    // [2d] fp32: zeros then 0.5*b_c
  size_t wmid_bf16;  // [5d, d]   0.5 * W_mid, heads stacked
  size_t bmid_half;  // [5d] fp32 0.5 * b_mid
  size_t wout_bf16;  // [15*16, 128] bf16: chunk c = (head c/3, features 128*(c%3)..), rows >= C zero
  size_t bout;       // [5*4] fp32
  // backward pass of the bf16 mode: unscaled bf16 W_mid ([5d, d], heads stacked) and its per-head transpose
  // ([5][d_in, d_out]), fp32 b_mid [5d]; the per-token chain and W_out use the fp32 fields above (f_w1 .. f_bc,
  // f_out_w, f_out_b), which the bf16 pack also carries.
  size_t wmid_full_bf16, wmidT_bf16, bmid_full;
  size_t wout_f32x4;  // [5d] float4: (W_out[0][f], W_out[1][f], W_out[2][f] or 0, 0) per stacked mid feature
  // PENEO_PREC_BF16 for every other configuration (unfused tensor-core forward, pair_heads_generic.cu): w1 / w2 / wc
  // as above (w1 / w2 only with shrink), plus per head and layer the bf16 [d, d] hidden weights (unscaled), fp32
  // biases, and the output layer zero-padded to 32 rows ([32, d] bf16, bias fp32 [32]).
  size_t g_mid_w[kNumHeads][kMaxMidLayers], g_mid_b[kNumHeads][kMaxMidLayers];
  size_t g_out_w[kNumHeads], g_out_b[kNumHeads];
  size_t total;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline bool bf16_supported(const peneo_dims& dm);
inline PackLayout pack_layout(const peneo_dims& dm, int prec) {
  PackLayout L{};
  size_t off = 0;
  const size_t d = dm.d, hid = dm.hid, hin = dm.hin;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off = align_up(off + bytes, 1024);
    return at;
  };
  if (prec == PENEO_PREC_FP32) {
    if (dm.shrink) {
      L.f_w1 = take(hid * hin * 4), L.f_b1 = take(hid * 4);
      L.f_w2 = take(d * hid * 4), L.f_b2 = take(d * 4);
    }
    L.f_wc = take(d * 2 * d * 4), L.f_bc = take(d * 4);
    for (int h = 0; h < kNumHeads; ++h) {
      for (int l = 0; l + 1 < dm.num_layers; ++l) L.f_mid_w[h][l] = take(d * d * 4), L.f_mid_b[h][l] = take(d * 4);
      L.f_out_w[h] = take(head_classes(h) * d * 4), L.f_out_b[h] = take(16);
    }
  } else if (!bf16_supported(dm)) {
    if (dm.shrink) {
      L.w1_bf16 = take(hid * hin * 2), L.b1 = take(hid * 4);
      L.w2_bf16 = take(d * hid * 2), L.b2 = take(d * 4);
    }
    L.wc_bf16 = take(2 * d * d * 2), L.bc_half = take(2 * d * 4);
    for (int h = 0; h < kNumHeads; ++h) {
      for (int l = 0; l + 1 < dm.num_layers; ++l) L.g_mid_w[h][l] = take(d * d * 2), L.g_mid_b[h][l] = take(d * 4);
      L.g_out_w[h] = take(32 * d * 2), L.g_out_b[h] = take(32 * 4);
    }
  } else {
    L.w1_bf16 = take(hid * hin * 2), L.b1 = take(hid * 4);
    L.w2_bf16 = take(d * hid * 2), L.b2 = take(d * 4);
    L.wc_bf16 = take(2 * d * d * 2), L.bc_half = take(2 * d * 4);
    L.wmid_bf16 = take(5 * d * d * 2), L.bmid_half = take(5 * d * 4);
    L.wout_bf16 = take(15 * 16 * 128 * 2), L.bout = take(5 * 4 * 4);
    L.wmid_full_bf16 = take(5 * d * d * 2), L.wmidT_bf16 = take(5 * d * d * 2), L.bmid_full = take(5 * d * 4);
    L.wout_f32x4 = take(5 * d * 16);
    L.f_w1 = take(hid * hin * 4), L.f_b1 = take(hid * 4);
    L.f_w2 = take(d * hid * 4), L.f_b2 = take(d * 4);
    L.f_wc = take(d * 2 * d * 4), L.f_bc = take(d * 4);
    for (int h = 0; h < kNumHeads; ++h) L.f_out_w[h] = take(head_classes(h) * d * 4), L.f_out_b[h] = take(16);
  }
  L.total = off;
  return L;
}

// fused tcgen05 path (K2 / T1): the shipped configuration
inline bool bf16_supported(const peneo_dims& dm) {
  return dm.shrink == 1 && dm.d == 384 && dm.hid == 768 && dm.num_layers == 2 && dm.hin % 64 == 0;
}
// unfused tensor-core forward for the other configurations (inference only): widths in whole 64-element K blocks
inline bool bf16_generic_supported(const peneo_dims& dm) {
  return !bf16_supported(dm) && dm.d % 64 == 0 && dm.hin % 64 == 0 && (!dm.shrink || dm.hid % 64 == 0) &&
         dm.num_layers >= 1 && dm.num_layers <= kMaxMidLayers + 1;
}

}  // namespace peneo
