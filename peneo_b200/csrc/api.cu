// extern "C" boundary (include/peneo_b200.h).  Argument validation, workspace carving and kernel
// sequencing only; no allocation, no global mutable state beyond the thread-local error string.
#include <cstdlib>
#include <cstring>

#include "classify.cuh"
#include "common.cuh"
#include "kernels.h"

namespace peneo {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_error; }

static int check_dims(const peneo_dims* dm) {
  PENEO_REQUIRE(dm != nullptr, "dims is NULL");
  PENEO_REQUIRE(dm->hin > 0 && dm->d > 0 && dm->num_layers >= 1 && dm->num_layers <= kMaxMidLayers + 1,
                "bad dims hin=%d d=%d num_layers=%d", dm->hin, dm->d, dm->num_layers);
  PENEO_REQUIRE(dm->hin % 4 == 0 && dm->d % 4 == 0, "hin and d must be multiples of 4");
  if (dm->shrink) {
    PENEO_REQUIRE(dm->hid > 0 && dm->hid % 4 == 0, "bad hid=%d", dm->hid);
  } else {
    PENEO_REQUIRE(dm->d == dm->hin, "without shrink the decoder width equals the input width");
  }
  return PENEO_OK;
}

static int check_prec(const peneo_dims* dm, int prec) {
  if (prec == PENEO_PREC_TF32) prec = PENEO_PREC_FP32;  // same pack, same layouts: only the backward's GEMM engine differs
  PENEO_REQUIRE(prec == PENEO_PREC_FP32 || prec == PENEO_PREC_BF16, "unknown precision mode %d", prec);
  if (prec == PENEO_PREC_BF16 && !bf16_supported(*dm) && !bf16_generic_supported(*dm)) {
    set_error("PENEO_PREC_BF16 needs hin, hid and d in multiples of 64 (fused path: shrink=1, hid=768, d=384, num_layers=2); "
              "got shrink=%d hid=%d d=%d L=%d hin=%d",
              dm->shrink, dm->hid, dm->d, dm->num_layers, dm->hin);
    return PENEO_E_INVALID;
  }
  return PENEO_OK;
}

// bf16 / tcgen05 per-token chain: x -> y1 -> y -> (0.5 A | 0.5 Bm), shared by the forward entry point and by the
// backward pass (which needs the very same bf16 projections, dropout included).
int token_proj_fwd_bf16(const peneo_dims& dm, const PackLayout& L, const char* pk, const void* x, int x_dtype,
                        int64_t x_row_stride, int64_t tokens, void* ab, char* ws, cudaStream_t st, const DropSpec* dp) {
  int rc;
  __nv_bfloat16* xc = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* y1 = reinterpret_cast<__nv_bfloat16*>(ws + align_up((size_t)tokens * dm.hin * 2, 1024));
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(y1) + align_up((size_t)tokens * dm.hid * 2, 1024));
  const __nv_bfloat16* xin;
  int64_t ldx;
  if (x_dtype == PENEO_DT_BF16 && x_row_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    xin = static_cast<const __nv_bfloat16*>(x), ldx = x_row_stride;
  } else {
    if ((rc = launch_cast_rows(x, x_dtype, x_row_stride, xc, PENEO_DT_BF16, tokens, dm.hin, st)) != PENEO_OK) return rc;
    xin = xc, ldx = dm.hin;
  }
  if (!dm.shrink)  // combine_fc reads the backbone output directly
    return launch_gemm_tc2(xin, ldx, reinterpret_cast<const __nv_bfloat16*>(pk + L.wc_bf16), dm.d,
                           reinterpret_cast<const float*>(pk + L.bc_half), static_cast<__nv_bfloat16*>(ab), 2 * dm.d, tokens,
                           2 * dm.d, dm.d, 0, 1, st, 0);
  // persistent tcgen05 GEMM chain (gemm_tc2): x -> y1 -> y -> (0.5 A | 0.5 Bm)
  if ((rc = launch_gemm_tc2(xin, ldx, reinterpret_cast<const __nv_bfloat16*>(pk + L.w1_bf16), dm.hin,
                            reinterpret_cast<const float*>(pk + L.b1), y1, dm.hid, tokens, dm.hid, dm.hin, 0, 1, st, 1, dp,
                            kSiteTok0)) != PENEO_OK)
    return rc;
  if ((rc = launch_gemm_tc2(y1, dm.hid, reinterpret_cast<const __nv_bfloat16*>(pk + L.w2_bf16), dm.hid,
                            reinterpret_cast<const float*>(pk + L.b2), y, dm.d, tokens, dm.d, dm.hid, 0, 1, st, 1, dp,
                            kSiteTok1)) != PENEO_OK)
    return rc;
  return launch_gemm_tc2(y, dm.d, reinterpret_cast<const __nv_bfloat16*>(pk + L.wc_bf16), dm.d,
                         reinterpret_cast<const float*>(pk + L.bc_half), static_cast<__nv_bfloat16*>(ab), 2 * dm.d, tokens,
                         2 * dm.d, dm.d, 0, 1, st, 0);
}

}  // namespace peneo

using namespace peneo;

extern "C" {

int peneo_abi_version(void) { return PENEO_ABI_VERSION; }
const char* peneo_last_error(void) { return get_error(); }

int peneo_device_supported(int device) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

size_t peneo_pack_bytes(const peneo_dims* dims, int prec) {
  if (check_dims(dims) != PENEO_OK || check_prec(dims, prec) != PENEO_OK) return 0;
  return pack_layout(*dims, prec).total;
}

int peneo_pack_weights(const peneo_dims* dims, int prec, const peneo_params* params, void* pack, void* stream) {
  int rc;
  if ((rc = check_dims(dims)) != PENEO_OK || (rc = check_prec(dims, prec)) != PENEO_OK) return rc;
  PENEO_REQUIRE(params && pack, "params / pack is NULL");
  PENEO_REQUIRE(params->combine_w && params->combine_b, "combine_fc parameters missing");
  if (dims->shrink)
    PENEO_REQUIRE(params->shrink_w1 && params->shrink_b1 && params->shrink_w2 && params->shrink_b2,
                  "shrink_projection parameters missing");
  for (int h = 0; h < kNumHeads; ++h) {
    PENEO_REQUIRE(params->out_w[h] && params->out_b[h], "head %d output layer missing", h);
    for (int l = 0; l + 1 < dims->num_layers; ++l)
      PENEO_REQUIRE(params->mid_w[h * 8 + l] && params->mid_b[h * 8 + l], "head %d hidden layer %d missing", h, l);
  }
  return pack_weights_impl(*dims, prec, *params, pack, static_cast<cudaStream_t>(stream));
}

size_t peneo_token_proj_workspace_bytes(const peneo_dims* dims, int prec, int64_t tokens) {
  if (check_dims(dims) != PENEO_OK || tokens < 0) return 0;
  const size_t el = prec == PENEO_PREC_BF16 ? 2 : 4;
  // x copy (cast / compaction) + y1 + y
  return align_up((size_t)tokens * dims->hin * el, 1024) + align_up((size_t)tokens * (dims->shrink ? dims->hid : 0) * el, 1024) +
         align_up((size_t)tokens * dims->d * el, 1024) + 1024;
}

int peneo_token_proj_fwd(const peneo_dims* dims, int prec, const void* pack, const void* x, int x_dtype,
                         int64_t x_row_stride, int64_t tokens, void* ab, void* workspace, const peneo_dropout* dropout,
                         void* stream) {
  int rc;
  if ((rc = check_dims(dims)) != PENEO_OK || (rc = check_prec(dims, prec)) != PENEO_OK) return rc;
  PENEO_REQUIRE(pack && x && ab && workspace, "token_proj_fwd: NULL pointer");
  PENEO_REQUIRE(tokens >= 0 && tokens < (1ll << 31), "token_proj_fwd: bad token count");
  PENEO_REQUIRE(x_row_stride >= dims->hin, "token_proj_fwd: row stride smaller than hin");
  if (tokens == 0) return PENEO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const peneo_dims& dm = *dims;
  const PackLayout L = pack_layout(dm, prec);
  const char* pk = static_cast<const char*>(pack);
  char* ws = static_cast<char*>(workspace);
  const int M = static_cast<int>(tokens);
  const DropSpec drop = make_drop(dropout);
  const DropSpec* dp = drop.thresh ? &drop : nullptr;
  if (prec != PENEO_PREC_FP32) return token_proj_fwd_bf16(dm, L, pk, x, x_dtype, x_row_stride, tokens, ab, ws, st, dp);
  {
    float* xc = reinterpret_cast<float*>(ws);
    float* y1 = reinterpret_cast<float*>(ws + align_up((size_t)tokens * dm.hin * 4, 1024));
    float* y = reinterpret_cast<float*>(reinterpret_cast<char*>(y1) + align_up((size_t)tokens * (dm.shrink ? dm.hid : 0) * 4, 1024));
    const float* xin;
    int64_t ldx;
    if (x_dtype == PENEO_DT_F32 && x_row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
      xin = static_cast<const float*>(x), ldx = x_row_stride;
    } else {
      if ((rc = launch_cast_rows(x, x_dtype, x_row_stride, xc, PENEO_DT_F32, tokens, dm.hin, st)) != PENEO_OK) return rc;
      xin = xc, ldx = dm.hin;
    }
    const float* yin = xin;
    int64_t ldy = ldx;
    if (dm.shrink) {
      if ((rc = launch_sgemm_nt(xin, ldx, reinterpret_cast<const float*>(pk + L.f_w1), dm.hin,
                                reinterpret_cast<const float*>(pk + L.f_b1), y1, dm.hid, M, dm.hid, dm.hin, 1, 1.f, st, dp,
                                kSiteTok0)) != PENEO_OK)
        return rc;
      if ((rc = launch_sgemm_nt(y1, dm.hid, reinterpret_cast<const float*>(pk + L.f_w2), dm.hid,
                                reinterpret_cast<const float*>(pk + L.f_b2), y, dm.d, M, dm.d, dm.hid, 1, 1.f, st, dp,
                                kSiteTok1)) != PENEO_OK)
        return rc;
      yin = y, ldy = dm.d;
    }
    const float* wc = reinterpret_cast<const float*>(pk + L.f_wc);
    float* abf = static_cast<float*>(ab);
    // A = y W_c[:, :d]^T ; Bm = y W_c[:, d:]^T + b_c
    if ((rc = launch_sgemm_nt(yin, ldy, wc, 2 * dm.d, nullptr, abf, 2 * dm.d, M, dm.d, dm.d, 0, 1.f, st)) != PENEO_OK) return rc;
    return launch_sgemm_nt(yin, ldy, wc + dm.d, 2 * dm.d, reinterpret_cast<const float*>(pk + L.f_bc), abf + dm.d,
                           2 * dm.d, M, dm.d, dm.d, 0, 1.f, st);
  }
}

int peneo_gather_tokens(const void* x, int x_dtype, int32_t batch, int32_t n, int32_t hin, int64_t batch_stride,
                        int64_t row_stride, void* out, int out_dtype, const peneo_dropout* in_dropout, void* stream) {
  PENEO_REQUIRE(x && out && batch >= 0 && n >= 1 && hin >= 4 && hin % 4 == 0, "gather_tokens: bad arguments");
  PENEO_REQUIRE(row_stride >= hin && batch_stride >= 0, "gather_tokens: bad strides");
  PENEO_REQUIRE(out_dtype == PENEO_DT_F32 || out_dtype == PENEO_DT_BF16, "gather_tokens: output must be fp32 or bf16");
  const DropSpec drop = make_drop(in_dropout);
  return launch_gather_tokens(x, x_dtype, batch_stride, row_stride, batch, n, hin, out, out_dtype, drop.thresh ? &drop : nullptr,
                              static_cast<cudaStream_t>(stream));
}

int peneo_token_dropout_bwd(float* dx, int64_t tokens, int32_t hin, const peneo_dropout* in_dropout, void* stream) {
  PENEO_REQUIRE(dx && tokens >= 0 && hin >= 1, "token_dropout_bwd: bad arguments");
  const DropSpec drop = make_drop(in_dropout);
  return launch_token_dropout_bwd(dx, tokens, hin, drop.thresh ? &drop : nullptr, static_cast<cudaStream_t>(stream));
}

int peneo_pair_heads_fwd(const peneo_dims* dims, int prec, const void* pack, const void* ab, int32_t batch, int32_t n,
                         float* const logits[PENEO_NUM_HEADS], const peneo_dropout* dropout, void* stream) {
  int rc;
  if ((rc = check_dims(dims)) != PENEO_OK || (rc = check_prec(dims, prec)) != PENEO_OK) return rc;
  PENEO_REQUIRE(pack && ab && logits, "pair_heads_fwd: NULL pointer");
  PENEO_REQUIRE(batch >= 0 && n >= 1 && n <= 46340, "pair_heads_fwd: bad sizes batch=%d n=%d", batch, n);
  if (batch == 0) return PENEO_OK;
  for (int h = 0; h < kNumHeads; ++h) PENEO_REQUIRE(logits[h], "pair_heads_fwd: logits[%d] is NULL", h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DropSpec drop = make_drop(dropout);
  const DropSpec* dp = drop.thresh ? &drop : nullptr;
  if (prec == PENEO_PREC_FP32)
    return launch_pair_heads_simt(*dims, pack, static_cast<const float*>(ab), batch, n, logits, st, dp);
  if (!bf16_supported(*dims))  // unfused tensor-core forward
    return launch_pair_heads_generic(*dims, pack, pack_layout(*dims, prec), static_cast<const __nv_bfloat16*>(ab), batch, n, logits, st, dp);
  // Default: the CTA-pair (cta_group::2, M = 256) variant of K2 — same results, half the W_mid traffic per SM,
  // ~3 % faster under the sustained power cap.  PENEO_K2_PAIR=0 selects the single-CTA kernel (A/B studies).
  static const bool use_pair = [] { const char* e = getenv("PENEO_K2_PAIR"); return !e || atoi(e) != 0; }();
  if (use_pair)
    return launch_pair_heads_tc_pair(pack, pack_layout(*dims, prec), static_cast<const __nv_bfloat16*>(ab), batch, n, logits, st, dp);
  return launch_pair_heads_tc(pack, pack_layout(*dims, prec), static_cast<const __nv_bfloat16*>(ab), batch, n, logits, st, dp);
}

size_t peneo_pair_heads_spots_workspace_bytes(int32_t batch, int32_t n) {
  return batch >= 1 && n >= 1 ? heads_spots_workspace_bytes(batch, n) : 0;
}

int peneo_pair_heads_spots_fwd(const peneo_dims* dims, int prec, const void* pack, const void* ab, int32_t batch, int32_t n,
                               int32_t cap, int32_t* spot_p, int32_t* spot_tag, float* spot_score, int32_t* counts,
                               void* workspace, void* stream) {
  int rc;
  if ((rc = check_dims(dims)) != PENEO_OK || (rc = check_prec(dims, prec)) != PENEO_OK) return rc;
  PENEO_REQUIRE(prec == PENEO_PREC_BF16 && bf16_supported(*dims), "pair_heads_spots_fwd: only the fused tcgen05 configuration "
                "(PENEO_PREC_BF16, shrink, hid 768, d 384, 2 layers); use pair_heads_fwd + decode_spots otherwise");
  PENEO_REQUIRE(pack && ab && spot_p && spot_tag && spot_score && counts && workspace, "pair_heads_spots_fwd: NULL pointer");
  PENEO_REQUIRE(batch >= 1 && batch <= 65535 && n >= 1 && n <= 46340 && cap >= 1, "pair_heads_spots_fwd: bad sizes batch=%d n=%d cap=%d",
                batch, n, cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t tiles = ((int64_t)batch * pair_count(n) + 127) / 128;
  const TileSpots ts = tile_spots_carve(workspace, tiles + 1);
  if ((rc = launch_pair_heads_tc_pair(pack, pack_layout(*dims, prec), static_cast<const __nv_bfloat16*>(ab), batch, n, nullptr, st,
                                      nullptr, nullptr, nullptr, &ts)) != PENEO_OK)
    return rc;
  return launch_gather_tile_spots(batch, n, workspace, cap, spot_p, spot_tag, spot_score, counts, st);
}

size_t peneo_heads_bwd_workspace_bytes(const peneo_dims* dims, int prec, int32_t batch, int32_t n) {
  if (check_dims(dims) != PENEO_OK || check_prec(dims, prec) != PENEO_OK || batch < 0 || n < 1) return 0;
  if (prec == PENEO_PREC_BF16 && !bf16_supported(*dims)) return 0;  // this configuration trains with PENEO_PREC_TF32 / FP32
  return heads_bwd_workspace_bytes(*dims, prec, batch, n);
}

int peneo_heads_bwd(const peneo_dims* dims, int prec, const void* pack, const void* x, int x_dtype, int64_t x_row_stride,
                    int32_t batch, int32_t n, const float* const dlogits[PENEO_NUM_HEADS], const peneo_grads* grads,
                    float* dx, void* workspace, const peneo_dropout* dropout, void* stream) {
  int rc;
  if ((rc = check_dims(dims)) != PENEO_OK || (rc = check_prec(dims, prec)) != PENEO_OK) return rc;
  PENEO_REQUIRE(prec != PENEO_PREC_BF16 || bf16_supported(*dims),
                "heads_bwd: PENEO_PREC_BF16 covers the fused configuration only; use PENEO_PREC_TF32 (tensor cores) or "
                "PENEO_PREC_FP32 with an fp32 pack for this one");
  PENEO_REQUIRE(pack && x && dlogits && grads && workspace, "heads_bwd: NULL pointer");
  PENEO_REQUIRE(batch >= 1 && n >= 1 && n <= 46340, "heads_bwd: bad sizes batch=%d n=%d", batch, n);
  PENEO_REQUIRE((int64_t)batch * n < (1ll << 31), "heads_bwd: too many tokens");
  PENEO_REQUIRE(x_row_stride >= dims->hin, "heads_bwd: row stride smaller than hin");
  PENEO_REQUIRE(grads->combine_w && grads->combine_b, "heads_bwd: combine_fc gradient buffers missing");
  if (dims->shrink)
    PENEO_REQUIRE(grads->shrink_w1 && grads->shrink_b1 && grads->shrink_w2 && grads->shrink_b2,
                  "heads_bwd: shrink_projection gradient buffers missing");
  for (int h = 0; h < kNumHeads; ++h) {
    PENEO_REQUIRE(dlogits[h] && grads->out_w[h] && grads->out_b[h], "heads_bwd: head %d buffers missing", h);
    for (int l = 0; l + 1 < dims->num_layers; ++l)
      PENEO_REQUIRE(grads->mid_w[h * 8 + l] && grads->mid_b[h * 8 + l], "heads_bwd: head %d layer %d buffers missing", h, l);
  }
  const DropSpec drop = make_drop(dropout);
  return launch_heads_bwd(*dims, prec, pack, x, x_dtype, x_row_stride, batch, n, dlogits, *grads, dx, workspace,
                          static_cast<cudaStream_t>(stream), drop.thresh ? &drop : nullptr);
}

size_t peneo_pair_loss_workspace_bytes(int32_t batch, int32_t n) { return pair_loss_workspace_bytes(batch, n); }

int peneo_fused_loss_supported(const peneo_dims* dims, int prec) {
  return dims && check_dims(dims) == PENEO_OK && prec == PENEO_PREC_BF16 && bf16_supported(*dims) ? 1 : 0;
}

int peneo_pair_heads_loss_fwd(const peneo_dims* dims, int prec, const void* pack, const void* ab, int32_t batch, int32_t n,
                              float* const logits[PENEO_NUM_HEADS], const int64_t* const tags[PENEO_NUM_HEADS],
                              const float* class_w_host, const float* ratio_host, float* out6, void* loss_workspace,
                              const peneo_dropout* dropout, void* stream, const peneo_saved_act* saved) {
  int rc;
  if ((rc = check_dims(dims)) != PENEO_OK || (rc = check_prec(dims, prec)) != PENEO_OK) return rc;
  PENEO_REQUIRE(peneo_fused_loss_supported(dims, prec), "pair_heads_loss_fwd: only the fused tcgen05 configuration "
                "(PENEO_PREC_BF16, shrink, hid 768, d 384, 2 layers); use pair_heads_fwd + pair_loss_fwd otherwise");
  PENEO_REQUIRE(pack && ab && logits && tags && class_w_host && out6 && loss_workspace, "pair_heads_loss_fwd: NULL pointer");
  PENEO_REQUIRE(batch >= 1 && n >= 1 && n <= 46340, "pair_heads_loss_fwd: bad sizes batch=%d n=%d", batch, n);
  for (int h = 0; h < kNumHeads; ++h) PENEO_REQUIRE(logits[h] && tags[h], "pair_heads_loss_fwd: head %d pointer is NULL", h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DropSpec drop = make_drop(dropout);
  FusedLossFwd fl{};
  for (int h = 0; h < kNumHeads; ++h) fl.tags[h] = tags[h];
  for (int c = 0; c < 3; ++c) fl.class_w[c] = class_w_host[c];
  fl.partial = pair_loss_partial_ptr(loss_workspace);
  if (saved) {
    PENEO_REQUIRE(saved->h && saved->s, "pair_heads_loss_fwd: saved->h and saved->s must both be given");
    fl.save_h = static_cast<__nv_bfloat16*>(saved->h), fl.save_s = static_cast<__nv_bfloat16*>(saved->s);
  }
  int grid = 0;
  if ((rc = launch_pair_heads_tc_pair(pack, pack_layout(*dims, prec), static_cast<const __nv_bfloat16*>(ab), batch, n, logits, st,
                                      drop.thresh ? &drop : nullptr, &fl, &grid)) != PENEO_OK)
    return rc;
  return launch_pair_loss_finalize(ratio_host, out6, loss_workspace, grid, st);
}

int peneo_heads_loss_bwd(const peneo_dims* dims, int prec, const void* pack, const void* x, int x_dtype, int64_t x_row_stride,
                         int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                         const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host, const float* ratio_host,
                         const float* grad_out6, const void* loss_workspace, const peneo_grads* grads, float* dx,
                         void* workspace, const peneo_dropout* dropout, void* stream, const peneo_saved_act* saved) {
  int rc;
  if ((rc = check_dims(dims)) != PENEO_OK || (rc = check_prec(dims, prec)) != PENEO_OK) return rc;
  PENEO_REQUIRE(peneo_fused_loss_supported(dims, prec), "heads_loss_bwd: only the fused tcgen05 configuration; use "
                "pair_loss_bwd + heads_bwd otherwise");
  PENEO_REQUIRE(pack && x && logits && tags && class_w_host && grad_out6 && loss_workspace && grads && workspace,
                "heads_loss_bwd: NULL pointer");
  PENEO_REQUIRE(batch >= 1 && n >= 1 && n <= 46340, "heads_loss_bwd: bad sizes batch=%d n=%d", batch, n);
  PENEO_REQUIRE((int64_t)batch * n < (1ll << 31), "heads_loss_bwd: too many tokens");
  PENEO_REQUIRE(x_row_stride >= dims->hin, "heads_loss_bwd: row stride smaller than hin");
  PENEO_REQUIRE(grads->combine_w && grads->combine_b && grads->shrink_w1 && grads->shrink_b1 && grads->shrink_w2 &&
                    grads->shrink_b2, "heads_loss_bwd: gradient buffers missing");
  FusedLossBwd fb{};
  for (int h = 0; h < kNumHeads; ++h) {
    PENEO_REQUIRE(logits[h] && tags[h] && grads->out_w[h] && grads->out_b[h] && grads->mid_w[h * 8] && grads->mid_b[h * 8],
                  "heads_loss_bwd: head %d buffers missing", h);
    fb.logits[h] = logits[h], fb.tags[h] = tags[h], fb.ratio[h] = ratio_host ? ratio_host[h] : 1.f;
  }
  for (int c = 0; c < 3; ++c) fb.class_w[c] = class_w_host[c];
  fb.grad_out6 = grad_out6;
  fb.loss_final = pair_loss_final_ptr(loss_workspace);
  if (saved) {
    PENEO_REQUIRE(saved->h && saved->s, "heads_loss_bwd: saved->h and saved->s must both be given");
    fb.save_h = static_cast<__nv_bfloat16*>(saved->h), fb.save_s = static_cast<__nv_bfloat16*>(saved->s);
  }
  const DropSpec drop = make_drop(dropout);
  const float* no_dz[kNumHeads] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  return launch_heads_bwd(*dims, prec, pack, x, x_dtype, x_row_stride, batch, n, no_dz, *grads, dx, workspace,
                          static_cast<cudaStream_t>(stream), drop.thresh ? &drop : nullptr, &fb);
}

int peneo_pair_loss_fwd(int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                        const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host, const float* ratio_host,
                        float* out6, void* workspace, void* stream) {
  PENEO_REQUIRE(logits && tags && class_w_host && out6 && workspace, "pair_loss_fwd: NULL pointer");
  for (int h = 0; h < kNumHeads; ++h) PENEO_REQUIRE(logits[h] && tags[h], "pair_loss_fwd: head %d pointer is NULL", h);
  return launch_pair_loss_fwd(batch, n, logits, tags, class_w_host, ratio_host, out6, workspace,
                              static_cast<cudaStream_t>(stream));
}

int peneo_pair_loss_bwd(int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                        const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host, const float* ratio_host,
                        const float* grad_out, const void* workspace, float* const dlogits[PENEO_NUM_HEADS], void* stream) {
  PENEO_REQUIRE(logits && tags && class_w_host && grad_out && workspace && dlogits, "pair_loss_bwd: NULL pointer");
  for (int h = 0; h < kNumHeads; ++h)
    PENEO_REQUIRE(logits[h] && tags[h] && dlogits[h], "pair_loss_bwd: head %d pointer is NULL", h);
  return launch_pair_loss_bwd(batch, n, logits, tags, class_w_host, ratio_host, grad_out, workspace, dlogits,
                              static_cast<cudaStream_t>(stream));
}

size_t peneo_pair_loss_ohem_workspace_bytes(int32_t batch, int32_t n) {
  return batch >= 1 && n >= 1 ? pair_loss_ohem_workspace_bytes(batch, n) : 0;
}

int peneo_pair_loss_ohem_fwd(int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                             const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host, const float* ratio_host,
                             int32_t num_hard_positive, int32_t num_hard_negative, float* out6, void* workspace,
                             void* stream) {
  PENEO_REQUIRE(logits && tags && class_w_host && out6 && workspace, "pair_loss_ohem_fwd: NULL pointer");
  for (int h = 0; h < kNumHeads; ++h) PENEO_REQUIRE(logits[h] && tags[h], "pair_loss_ohem_fwd: head %d pointer is NULL", h);
  return launch_pair_loss_ohem_fwd(batch, n, logits, tags, class_w_host, ratio_host, num_hard_positive, num_hard_negative,
                                   out6, workspace, static_cast<cudaStream_t>(stream));
}

int peneo_pair_loss_ohem_bwd(int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                             const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host, const float* ratio_host,
                             const float* grad_out, const void* workspace, float* const dlogits[PENEO_NUM_HEADS],
                             void* stream) {
  PENEO_REQUIRE(logits && tags && class_w_host && grad_out && workspace && dlogits, "pair_loss_ohem_bwd: NULL pointer");
  for (int h = 0; h < kNumHeads; ++h)
    PENEO_REQUIRE(logits[h] && tags[h] && dlogits[h], "pair_loss_ohem_bwd: head %d pointer is NULL", h);
  return launch_pair_loss_ohem_bwd(batch, n, logits, tags, class_w_host, ratio_host, grad_out, workspace, dlogits,
                                   static_cast<cudaStream_t>(stream));
}

int peneo_scatter_tags(const int32_t* spots_bijt, int64_t num_spots, int32_t batch, int32_t n, int64_t* tags,
                       void* stream) {
  PENEO_REQUIRE(tags && (spots_bijt || num_spots == 0) && batch >= 0 && n >= 1 && num_spots >= 0,
                "scatter_tags: bad arguments");
  return launch_scatter_tags(spots_bijt, num_spots, batch, n, tags, static_cast<cudaStream_t>(stream));
}

size_t peneo_decode_spots_workspace_bytes(int32_t batch, int32_t n) { return decode_spots_workspace_bytes(batch, n); }

int peneo_decode_spots(int32_t batch, int32_t n, const void* const in[PENEO_NUM_HEADS], int in_dtype, int32_t cap,
                       int32_t* spot_p, int32_t* spot_tag, float* spot_score, int32_t* counts, void* workspace,
                       void* stream) {
  PENEO_REQUIRE(in && spot_p && spot_tag && spot_score && counts && workspace, "decode_spots: NULL pointer");
  for (int h = 0; h < kNumHeads; ++h) PENEO_REQUIRE(in[h], "decode_spots: in[%d] is NULL", h);
  return launch_decode_spots(batch, n, in, in_dtype, cap, spot_p, spot_tag, spot_score, counts, workspace,
                             static_cast<cudaStream_t>(stream));
}

size_t peneo_decode_resolve_doc_ints(int32_t n, int32_t cap) { return decode_resolve_doc_ints(n, cap); }
size_t peneo_decode_resolve_workspace_bytes(int32_t batch, int32_t n) { return decode_resolve_workspace_bytes(batch, n); }

int peneo_decode_resolve(int32_t batch, int32_t n, int32_t cap, const int32_t* spot_p, const int32_t* spot_tag,
                         const float* spot_score, const int32_t* counts, int decode_gt, float score_thresh, int32_t* out,
                         void* workspace, void* stream) {
  PENEO_REQUIRE(spot_p && spot_tag && spot_score && counts && out && workspace, "decode_resolve: NULL pointer");
  return launch_decode_resolve(batch, n, cap, spot_p, spot_tag, spot_score, counts, decode_gt, score_thresh, out,
                               workspace, static_cast<cudaStream_t>(stream));
}

int peneo_selftest(uint32_t* failed_mask_host, char* report_host, size_t report_bytes) {
  return run_selftest(failed_mask_host, report_host, report_bytes);
}
int peneo_probe_rates(double* out_host, int n_out) {
  PENEO_REQUIRE(out_host && n_out > 0, "probe_rates: bad arguments");
  return run_probe_rates(out_host, n_out);
}

}  // extern "C"
