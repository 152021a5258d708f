/*
 * peneo_b200 — C ABI of the B200-native PEneo hot path (heads + pairwise loss + decode).
 *
 * The reference (ZeningLin/PEneo) is pure Python and has no FFI; the entry points below are the
 * boundary a maintainer would bind (ctypes) to replace, one for one:
 *
 *   peneo_pack_weights ............ parameter reads of PEneoDecoder        model/peneo_decoder.py:204-313
 *   peneo_token_proj_fwd .......... shrink_projection + the per-token half
 *                                   of HandshakingKernel.combine_fc         model/peneo_decoder.py:215-222, 126, 349-350
 *   peneo_gather_tokens ........... CLS / visual-token strip + dropout + cast before the decoder   model/modeling_peneo.py:134-165
 *   peneo_pair_heads_fwd .......... HandshakingKernel.forward (pair part)
 *                                   + the five classifier heads             model/peneo_decoder.py:149-177, 355-363
 *   peneo_pair_loss_fwd ........... calculate_peneo_loss / CrossEntropyLossOHEM
 *                                   (OHEM off) + the ratio-weighted sum     model/peneo_decoder.py:315-336, 375-428; model/custom_loss.py:189-202
 *   peneo_pair_loss_ohem_fwd/bwd .. CrossEntropyLossOHEM with hard-example
 *                                   selection (forward value + d loss/d logits) model/custom_loss.py:204-288
 *   peneo_pair_heads_loss_fwd ..... the two above fused: heads + loss in one sweep over the pair tiles
 *   peneo_heads_loss_bwd .......... loss backward + heads backward fused (no d loss / d logits tensor)
 *   peneo_heads_bwd ............... autograd backward of token_proj + pair_heads (implicit in the reference:
 *                                   loss.backward() through model/peneo_decoder.py:349-363)
 *   peneo_scatter_tags ............ HandshakingTaggingScheme.spots2shaking_tag4batch   model/peneo_decoder.py:34-73
 *   peneo_decode_spots ............ HandshakingTaggingScheme.get_spots_from_shaking_tag model/peneo_decoder.py:75-115
 *   peneo_pair_heads_spots_fwd .... pair heads + get_spots_from_shaking_tag fused (no logits in HBM), inference
 *   peneo_decode_resolve .......... parse_matrix_spots + the key/value chain walk of
 *                                   sample_decode_peneo                     pipeline/decode.py:9-69, 169-368
 *
 * Conventions: plain pointers and sizes only; all pointers are DEVICE pointers unless the name
 * ends in _host; every function enqueues on `stream` (a cudaStream_t passed as void*), allocates
 * nothing, and returns 0 on success or a negative PENEO_E_* code, after which
 * peneo_last_error() (thread-local) describes the failure.  No CPU fallback exists.
 */
#ifndef PENEO_B200_H
#define PENEO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PENEO_ABI_VERSION 4

#define PENEO_OK 0
#define PENEO_E_INVALID -1     /* bad argument / unsupported configuration */
#define PENEO_E_CUDA -2        /* a CUDA runtime or driver call failed */
#define PENEO_E_UNSUPPORTED -3 /* device is not sm_100 */
#define PENEO_E_OVERFLOW -4    /* an output capacity was too small (counts are still exact) */

/* arithmetic modes of the heads */
#define PENEO_PREC_FP32 0 /* CUDA-core fp32, exact SiLU: meets the 1e-3 logit tolerance */
#define PENEO_PREC_BF16 1 /* tcgen05 bf16 x bf16 -> fp32: meets the 2e-2 logit tolerance (see peneo_pack_bytes) */
#define PENEO_PREC_TF32 2 /* peneo_heads_bwd only: PENEO_PREC_FP32 pack and layouts, every GEMM on tcgen05 kind::tf32 */

/* element types of caller tensors */
#define PENEO_DT_F32 0
#define PENEO_DT_BF16 1
#define PENEO_DT_F16 2
#define PENEO_DT_I64 3

#define PENEO_NUM_HEADS 5 /* order: LE, EL-h2h, EL-t2t, LG-h2h, LG-t2t (model/peneo_decoder.py:365-372) */

int peneo_abi_version(void);
const char* peneo_last_error(void);
/* 1 if cuda device `device` can run the kernels (compute capability 10.x), else 0. */
int peneo_device_supported(int device);

/* ------------------------------------------------------------------ configuration */
typedef struct peneo_dims {
  int32_t hin;        /* input_size: backbone hidden (768) or 960 for LiLT                 */
  int32_t hid;        /* backbone_config["hidden_size"] (width of shrink layer 1)          */
  int32_t d;          /* decoder hidden: hid/2 when shrink else hin                        */
  int32_t shrink;     /* config.peneo_decoder_shrink                                       */
  int32_t num_layers; /* config.peneo_classifier_num_layers (>= 1)                         */
} peneo_dims;

/* fp32 parameters in the reference's own layouts ([out, in] row-major, nn.Linear). */
typedef struct peneo_params {
  const float* shrink_w1; /* [hid, hin]  shrink_projection.0.weight (NULL when !shrink) */
  const float* shrink_b1; /* [hid]                                                    */
  const float* shrink_w2; /* [d, hid]    shrink_projection.3.weight                   */
  const float* shrink_b2; /* [d]                                                      */
  const float* combine_w; /* [d, 2d]     handshaking_kernel.combine_fc.weight         */
  const float* combine_b; /* [d]                                                      */
  /* per head h, mid layer l < num_layers-1: [d, d] / [d]; mid_w[h * 8 + l]           */
  const float* mid_w[PENEO_NUM_HEADS * 8];
  const float* mid_b[PENEO_NUM_HEADS * 8];
  const float* out_w[PENEO_NUM_HEADS]; /* [C_h, d], C = 2,3,3,3,3 */
  const float* out_b[PENEO_NUM_HEADS]; /* [C_h]                   */
} peneo_params;

/* Dropout of the decoder's own nn.Dropout modules in training mode (model/peneo_decoder.py:218, 221, 261).
 * NULL or p == 0 means eval-mode behaviour (identity).  The mask is a pure function of (seed, site, element),
 * so the backward pass regenerates it: pass the SAME struct to the forward calls and to peneo_heads_bwd. */
typedef struct peneo_dropout {
  float p;       /* drop probability, backbone_config["hidden_dropout_prob"] */
  uint64_t seed; /* fresh per training step */
} peneo_dropout;

/* Bytes of the packed-weight buffer for (dims, prec).  0 if the combination is unsupported.
 * PENEO_PREC_BF16: hin, hid and d in multiples of 64.  shrink = 1, hid = 768 (d = 384), num_layers = 2 runs the fused
 * tcgen05 kernels, forward and backward; every other configuration runs an unfused tensor-core forward
 * (peneo_token_proj_fwd / peneo_pair_heads_fwd, dropout included) and its backward pass with PENEO_PREC_TF32
 * (peneo_heads_bwd with a PENEO_PREC_FP32 pack: recompute + gradient GEMMs on tcgen05 kind::tf32) or, exact, with
 * PENEO_PREC_FP32 on CUDA cores. */
size_t peneo_pack_bytes(const peneo_dims* dims, int prec);
/* Convert the fp32 parameters into the kernel layouts (bf16 copies, concatenated and pre-scaled
 * matrices).  Call again whenever a parameter changed. */
int peneo_pack_weights(const peneo_dims* dims, int prec, const peneo_params* params, void* pack, void* stream);

/* ------------------------------------------------------------------ heads, forward */
/* Scratch bytes needed by token_proj / pair_heads for `tokens` = B*N rows. */
size_t peneo_token_proj_workspace_bytes(const peneo_dims* dims, int prec, int64_t tokens);

/* Per-token projections.  x: [tokens, hin] of x_dtype, row stride x_row_stride elements.
 * Writes ab: [tokens, 2d]: columns [0,d) = A = y W_c[:, :d]^T, [d,2d) = Bm = y W_c[:, d:]^T + b_c
 * (fp32 for PENEO_PREC_FP32; bf16, pre-multiplied by 1/2, for PENEO_PREC_BF16).
 * If y_out != NULL also stores y (shrink output, same type as ab) for the backward pass. */
int peneo_token_proj_fwd(const peneo_dims* dims, int prec, const void* pack, const void* x, int x_dtype,
                         int64_t x_row_stride, int64_t tokens, void* ab, void* workspace, const peneo_dropout* dropout,
                         void* stream);

/* Backbone -> decoder seam (model/modeling_peneo.py:134-173): the decoder input is a strided view of the backbone output
 * (CLS row / visual tokens stripped), passed through nn.Dropout in training, then fed to the first GEMM.  One pass:
 * token (b, t) is read at x + b * batch_stride + t * row_stride (elements), optionally dropped (in_dropout, NULL = none:
 * same counter-based mask scheme as peneo_dropout, its own site), cast to out_dtype and written densely as
 * out[(b * n + t) * hin ..] — the [tokens, hin] matrix peneo_token_proj_fwd / peneo_heads_bwd read in place.
 * peneo_token_dropout_bwd applies the same mask (and 1 / (1 - p)) to d loss / d x in place. */
int peneo_gather_tokens(const void* x, int x_dtype, int32_t batch, int32_t n, int32_t hin, int64_t batch_stride,
                        int64_t row_stride, void* out, int out_dtype, const peneo_dropout* in_dropout, void* stream);
int peneo_token_dropout_bwd(float* dx, int64_t tokens, int32_t hin, const peneo_dropout* in_dropout, void* stream);

/* Pair scoring + classifier heads for `batch` documents of `n` tokens each.
 * logits[h]: fp32 [batch, n(n+1)/2, C_h], row p(i,j) = i*n - i(i-1)/2 + (j-i). */
int peneo_pair_heads_fwd(const peneo_dims* dims, int prec, const void* pack, const void* ab, int32_t batch, int32_t n,
                         float* const logits[PENEO_NUM_HEADS], const peneo_dropout* dropout, void* stream);

/* ------------------------------------------------------------------ loss */
/* Class-weighted cross entropy of the five heads over the whole batch (OHEM off):
 *   loss_h = sum_m w[t_m] nll_m / sum_m w[t_m];  out[0..4] = loss_h, out[5] = sum_h ratio_h loss_h.
 * tags[h]: int64 [batch*P].  class_w: 3 floats (LE uses the first two).  workspace: see below. */
size_t peneo_pair_loss_workspace_bytes(int32_t batch, int32_t n);
int peneo_pair_loss_fwd(int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                        const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host,
                        const float* ratio_host, float* out6, void* workspace, void* stream);
/* d L / d logits for the same definition.  grad_out6: six device floats, d L / d out[0..5] of the forward call
 * (the five sub-losses, then their ratio-weighted sum), so head h is scaled by grad_out6[5] * ratio_h + grad_out6[h]. */
int peneo_pair_loss_bwd(int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                        const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host,
                        const float* ratio_host, const float* grad_out6, const void* workspace,
                        float* const dlogits[PENEO_NUM_HEADS], void* stream);

/* Fused forms for training with the loss inside the pair tiles (PENEO_PREC_BF16, the fused tcgen05 configuration,
 * OHEM off; peneo_fused_loss_supported() tells).  What model/peneo_decoder.py:355-428 + model/custom_loss.py:189-202 do
 * in separate passes over [batch, P, C] tensors:
 *   peneo_pair_heads_loss_fwd = peneo_pair_heads_fwd + peneo_pair_loss_fwd in ONE kernel: the epilogue that writes a
 *     pair's logits also reduces its w[t] * nll and w[t] (same outputs: logits, out6, loss_workspace);
 *   peneo_heads_loss_bwd      = peneo_pair_loss_bwd + peneo_heads_bwd without d loss / d logits ever existing in
 *     memory: the backward tiles compute it in registers from logits + tags + the normalisers in loss_workspace
 *     (grad_out6 as in peneo_pair_loss_bwd), and reduce db_out on the fly.
 * loss_workspace: peneo_pair_loss_workspace_bytes(batch, n), written by the forward call, read by the backward call.
 *
 * `saved` (optional, NULL = recompute): activations the forward call stores as they pass through its registers so that
 * the backward call needs no recompute GEMM (what autograd keeps for model/peneo_decoder.py:256-271, in bf16):
 *   h  bf16 [batch * P, 5 * d]  pre-activations (W_mid s + b_mid) / 2 of the five hidden layers; the backward call
 *                               OVERWRITES it with the gradient G, so one forward's buffers serve one backward;
 *   s  bf16 [batch * P, d]      pair representations SiLU(a_i + b_j).
 * 4.5 KB per pair (18.8 GB for 32 seq-512 documents): pass NULL when that does not fit — the backward then regenerates
 * both on the tensor cores chunk by chunk (~3 GB of scratch). */
typedef struct peneo_saved_act {
  void* h;
  void* s;
} peneo_saved_act;
int peneo_fused_loss_supported(const peneo_dims* dims, int prec);
int peneo_pair_heads_loss_fwd(const peneo_dims* dims, int prec, const void* pack, const void* ab, int32_t batch, int32_t n,
                              float* const logits[PENEO_NUM_HEADS], const int64_t* const tags[PENEO_NUM_HEADS],
                              const float* class_w_host, const float* ratio_host, float* out6, void* loss_workspace,
                              const peneo_dropout* dropout, void* stream, const peneo_saved_act* saved);
/* (declared here, next to its forward half; the gradient structs are defined below) */
struct peneo_grads;
int peneo_heads_loss_bwd(const peneo_dims* dims, int prec, const void* pack, const void* x, int x_dtype,
                         int64_t x_row_stride, int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                         const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host, const float* ratio_host,
                         const float* grad_out6, const void* loss_workspace, const struct peneo_grads* grads, float* dx,
                         void* workspace, const peneo_dropout* dropout, void* stream, const peneo_saved_act* saved);

/* Same five sub-losses with online hard-example mining (num_hard_positive / num_hard_negative as in
 * PEneoConfig.peneo_ohem_num_positive / _negative; -1 = keep the whole side).  Reproduces the reference's
 * selection rule exactly, including its indexing quirk and the raw (k_pos + k_neg) divisor
 * (model/custom_loss.py:255-280).  The workspace keeps the kept-element masks for the backward call.
 * Synchronises `stream` (the side sizes are needed on the host, as in the reference). */
size_t peneo_pair_loss_ohem_workspace_bytes(int32_t batch, int32_t n);
int peneo_pair_loss_ohem_fwd(int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                             const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host,
                             const float* ratio_host, int32_t num_hard_positive, int32_t num_hard_negative, float* out6,
                             void* workspace, void* stream);
int peneo_pair_loss_ohem_bwd(int32_t batch, int32_t n, const float* const logits[PENEO_NUM_HEADS],
                             const int64_t* const tags[PENEO_NUM_HEADS], const float* class_w_host,
                             const float* ratio_host, const float* grad_out6, const void* workspace,
                             float* const dlogits[PENEO_NUM_HEADS], void* stream);

/* ------------------------------------------------------------------ heads, backward */
/* Gradient outputs, fp32, same layouts as peneo_params (all OVERWRITTEN by peneo_heads_bwd). */
typedef struct peneo_grads {
  float* shrink_w1;
  float* shrink_b1;
  float* shrink_w2;
  float* shrink_b2;
  float* combine_w;
  float* combine_b;
  float* mid_w[PENEO_NUM_HEADS * 8];
  float* mid_b[PENEO_NUM_HEADS * 8];
  float* out_w[PENEO_NUM_HEADS];
  float* out_b[PENEO_NUM_HEADS];
} peneo_grads;

/* Backward of peneo_token_proj_fwd + peneo_pair_heads_fwd for `batch` documents of `n` tokens:
 * given dlogits[h] = d loss / d logits[h] (fp32 [batch, P, C_h]) computes every parameter gradient
 * and, when dx != NULL, d loss / d x (fp32 [batch*n, hin], contiguous).  The pair activations are
 * recomputed from x chunk by chunk (nothing of size [P, d] is kept from the forward pass).
 * `pack` must be a PENEO_PREC_FP32 pack when prec == PENEO_PREC_FP32 or PENEO_PREC_TF32 (any configuration). */
size_t peneo_heads_bwd_workspace_bytes(const peneo_dims* dims, int prec, int32_t batch, int32_t n);
int peneo_heads_bwd(const peneo_dims* dims, int prec, const void* pack, const void* x, int x_dtype,
                    int64_t x_row_stride, int32_t batch, int32_t n, const float* const dlogits[PENEO_NUM_HEADS],
                    const peneo_grads* grads, float* dx, void* workspace, const peneo_dropout* dropout, void* stream);

/* ------------------------------------------------------------------ tags */
/* Dense tags from sparse spots: tags[b, p(i,j)] = tag for every (b, i, j, tag) quadruple
 * (later spots overwrite earlier ones).  tags: int64 [batch, P], zero-filled by this call. */
int peneo_scatter_tags(const int32_t* spots_bijt, int64_t num_spots, int32_t batch, int32_t n, int64_t* tags,
                       void* stream);

/* ------------------------------------------------------------------ decode */
/* Spot extraction for `batch` documents and the five heads.
 * in[h]: [batch, P, C_h] logits (PENEO_DT_F32 / BF16 / F16: softmax -> argmax -> max prob, computed
 * the way the reference does in that dtype) or [batch, P] int64 tags (PENEO_DT_I64: pred = tag,
 * score = 1).  For (b, h) the spots with pred != 0 are written in increasing p to
 *   spot_p[(b*5+h)*cap + r], spot_tag[...], spot_score[...],   r < min(count, cap)
 * and counts[b*5+h] = exact number found.  Returns PENEO_E_OVERFLOW if any count > cap is found
 * by a later peneo_decode_check (counts are read on the host by the caller). */
size_t peneo_decode_spots_workspace_bytes(int32_t batch, int32_t n);
int peneo_decode_spots(int32_t batch, int32_t n, const void* const in[PENEO_NUM_HEADS], int in_dtype, int32_t cap,
                       int32_t* spot_p, int32_t* spot_tag, float* spot_score, int32_t* counts, void* workspace,
                       void* stream);

/* Heads + spot extraction in one sweep (inference, PENEO_PREC_BF16, the fused tcgen05 configuration): what
 * peneo_pair_heads_fwd followed by peneo_decode_spots produce — the same compact (p, tag, score) lists and exact
 * counts, bit-identical scores — without the [batch, P, C] logits ever being written to or read from HBM: the pair
 * kernel's epilogue classifies every pair (model/peneo_decoder.py:98-114) and keeps only the non-zero predictions,
 * which a small gather kernel compacts in increasing p.  Feed the result to peneo_decode_resolve. */
size_t peneo_pair_heads_spots_workspace_bytes(int32_t batch, int32_t n);
int peneo_pair_heads_spots_fwd(const peneo_dims* dims, int prec, const void* pack, const void* ab, int32_t batch, int32_t n,
                               int32_t cap, int32_t* spot_p, int32_t* spot_tag, float* spot_score, int32_t* counts,
                               void* workspace, void* stream);

/* Link resolution for `batch` documents from the compact spots of peneo_decode_spots.
 * Output: one int32 record block per document at out + b * peneo_decode_resolve_doc_ints(n, cap):
 *   header[16]: n_le, n_lgh, n_lgt, n_elh, n_elt, n_kv, 0...
 *   le[2n]   : (head, tail) in the reference's dict order     (parse_matrix_spots, top-1 or GT mode)
 *   lgh[2n], lgt[2n] : same for line grouping h2h / t2t
 *   elh[2cap], elt[2cap] : filtered, tag-2-swapped (head, tail) lists in arrival order
 *   kv[4cap] : (key_head, value_head, n_key_lines, n_value_lines) for every valid pair, in order
 */
size_t peneo_decode_resolve_doc_ints(int32_t n, int32_t cap);
size_t peneo_decode_resolve_workspace_bytes(int32_t batch, int32_t n);
int peneo_decode_resolve(int32_t batch, int32_t n, int32_t cap, const int32_t* spot_p, const int32_t* spot_tag,
                         const float* spot_score, const int32_t* counts, int decode_gt, float score_thresh,
                         int32_t* out, void* workspace, void* stream);

/* ------------------------------------------------------------------ diagnostics */
/* Runs the on-device unit checks of the tcgen05 / TMA building blocks (SS- and TS-form UMMA, TMEM
 * load/store round trip, swizzled TMA tile) against host arithmetic.  Bit k of *failed_mask is
 * set when check k failed; report_host (optional, >= 2048 bytes) receives a text report. */
int peneo_selftest(uint32_t* failed_mask_host, char* report_host, size_t report_bytes);
/* Throughput probes used by bench.py to state the MUFU / FMA co-limits: returns ops/s. */
int peneo_probe_rates(double* out_host, int n_out);

#ifdef __cplusplus
}
#endif
#endif /* PENEO_B200_H */
