// Internal launcher declarations shared between the .cu files (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace peneo {

// simt_kernels.cu
int launch_sgemm_nt(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
                    int M, int N, int K, int act, float out_scale, cudaStream_t st, const DropSpec* drop = nullptr,
                    uint32_t site = 0);
int launch_cast_rows(const void* src, int src_dtype, int64_t src_stride, void* dst, int dst_dtype, int64_t rows,
                     int cols, cudaStream_t st);
// backbone -> decoder seam: strip (strided [B, n, cols] view) + optional input dropout + cast, one pass; and the mask on dx
int launch_gather_tokens(const void* src, int src_dtype, int64_t batch_stride, int64_t row_stride, int batch, int n, int cols,
                         void* dst, int dst_dtype, const DropSpec* drop, cudaStream_t st);
int launch_token_dropout_bwd(float* dx, int64_t tokens, int cols, const DropSpec* drop, cudaStream_t st);
int pack_weights_impl(const peneo_dims& dm, int prec, const peneo_params& P, void* pack, cudaStream_t st);
int launch_pair_heads_simt(const peneo_dims& dm, const void* pack, const float* ab, int batch, int n,
                           float* const logits[kNumHeads], cudaStream_t st, const DropSpec* drop = nullptr);

// api.cu : bf16 per-token chain (x -> 0.5 A | 0.5 Bm), also used by the backward pass
int token_proj_fwd_bf16(const peneo_dims& dm, const PackLayout& L, const char* pk, const void* x, int x_dtype,
                        int64_t x_row_stride, int64_t tokens, void* ab, char* ws, cudaStream_t st, const DropSpec* dp);

// gemm_tc.cu : C[M, N] = act(A[M, K] W[N, K]^T + bias) ; A, W bf16 K-contiguous; C bf16
int launch_gemm_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias,
                   __nv_bfloat16* C, int64_t ldc, int64_t M, int N, int K, int act, cudaStream_t st);

// C[M, N] (op)= A W^T ; out_mode 0 bf16 store (+bias, SiLU when act), 1 fp32 store, 2 fp32 +=, 3 fp32 atomicAdd with split-K
int launch_gemm_tc2(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias, void* C,
                    int64_t ldc, int64_t M, int N, int K, int out_mode, int splits, cudaStream_t st, int act = 0,
                    const DropSpec* drop = nullptr, uint32_t site = 0, uint32_t drop_row0 = 0);  // mask row = drop_row0 + m

// pair_heads_tc.cu
int launch_pair_heads_tc(const void* pack, const PackLayout& L, const __nv_bfloat16* ab, int batch, int n,
                         float* const logits[kNumHeads], cudaStream_t st, const DropSpec* drop = nullptr);

// pair_heads_tc2.cu : the same kernel on CTA pairs (cta_group::2, M = 256).  With `loss` != NULL the class-weighted cross
// entropy is reduced inside the tiles (per-CTA fp64 partial sums; *grid_out CTAs wrote them).
struct FusedLossFwd {
  const int64_t* tags[kNumHeads];
  float class_w[3];
  double* partial;  // [5][grid][2]
  // optional (both or none): the pre-activations (u + b_mid) / 2, bf16 [batch * P, 1920], and s, bf16 [batch * P, 384],
  // saved for a backward pass without recompute (pair_bwd_elem.cu)
  __nv_bfloat16* save_h = nullptr;
  __nv_bfloat16* save_s = nullptr;
};
struct TileSpots;  // classify.cuh
int launch_pair_heads_tc_pair(const void* pack, const PackLayout& L, const __nv_bfloat16* ab, int batch, int n,
                              float* const logits[kNumHeads], cudaStream_t st, const DropSpec* drop = nullptr,
                              const FusedLossFwd* loss = nullptr, int* grid_out = nullptr, const TileSpots* spots = nullptr);

// pair_heads_generic.cu : unfused tensor-core forward for the configurations K2 does not cover (any d % 64 == 0, any
// num_layers): S = SiLU(a_i + b_j) per chunk of pairs, one gemm_tc2 per hidden layer and head, output layer as an
// N = 32 GEMM
int launch_pair_heads_generic(const peneo_dims& dm, const void* pack, const PackLayout& L, const __nv_bfloat16* ab, int batch,
                              int n, float* const logits[kNumHeads], cudaStream_t st, const DropSpec* drop = nullptr);

// pair_bwd_tc.cu : regenerated S and G = (dz W_out) SiLU'(u) of a chunk of pairs, plus per-CTA partial sums
// [ctas][3][1920] of dz^T SiLU(u) (bf16 backward)
int launch_pair_bwd_prep(const void* pack, const PackLayout& L, const __nv_bfloat16* ab, int n, int64_t g0, int rows,
                         const float* const dz[kNumHeads], __nv_bfloat16* S, __nv_bfloat16* G, float* dwout_part,
                         cudaStream_t st, const DropSpec* drop = nullptr);

// pair_bwd_tc2.cu : T1 on CTA pairs.  With `fused` != NULL the loss backward happens inside the kernel: dz is computed
// in registers from the stored logits + tags (no d loss / d logits tensor in HBM) and db_out is reduced there too.
struct FusedLossBwd {
  const float* logits[kNumHeads];   // fp32 [batch*P, C_h]
  const int64_t* tags[kNumHeads];   // int64 [batch*P]
  const float* grad_out6;           // device: d L / d (loss_0 .. loss_4, total)
  const double* loss_final;         // device: [5][2] = (sum w nll, sum w) of the forward loss (workspace of pair_loss_fwd)
  float ratio[kNumHeads], class_w[3];
  float* dbout[kNumHeads];          // db_out gradient buffers (accumulated into)
  // optional (both or none): what the forward pass saved (FusedLossFwd::save_h / save_s); the backward then runs
  // pair_bwd_elem.cu instead of T1 and turns save_h into G in place
  __nv_bfloat16* save_h = nullptr;
  __nv_bfloat16* save_s = nullptr;
};
// pair_bwd_elem.cu : T1e, the hidden-activation backward from saved pre-activations (in place: h -> G)
int pair_bwd_elem_max_ctas();
int launch_pair_bwd_elem(const void* pack, const PackLayout& L, int64_t g0, int rows, const FusedLossBwd& fused,
                         __nv_bfloat16* hg, float* dwout_part, int* ctas_out, cudaStream_t st, const DropSpec* drop = nullptr);
// gemm_ds_fused.cu : T1f, the same transform inside the operand stage of the dS GEMM (h -> G in shared memory, dS = G W_mid
// on CTA pairs, G written back over h for the dW_mid GEMM)
int launch_gemm_ds_fused(const void* pack, const PackLayout& L, __nv_bfloat16* hg, const __nv_bfloat16* wmid_full,
                         __nv_bfloat16* dS, int64_t g0, int rows, const FusedLossBwd& fused, float* dwout_part,
                         cudaStream_t st, const DropSpec* drop = nullptr);
int launch_pair_bwd_prep_pair(const void* pack, const PackLayout& L, const __nv_bfloat16* ab, int n, int64_t g0, int rows,
                              const float* const dz[kNumHeads], const FusedLossBwd* fused, __nv_bfloat16* S,
                              __nv_bfloat16* G, float* dwout_part, cudaStream_t st, const DropSpec* drop = nullptr);

// gemm_bwd_tc.cu
int launch_gemm_ds(const __nv_bfloat16* G, const __nv_bfloat16* wmid_full, __nv_bfloat16* dS, int rows, cudaStream_t st);
int launch_gemm_dw(const __nv_bfloat16* G, const __nv_bfloat16* S, float* const dW[kNumHeads], float* const db[kNumHeads],
                   int rows, cudaStream_t st);

// gemm_bwd_tc2.cu : the same two GEMMs on CTA pairs (cta_group::2): half the B-operand traffic per SM
int launch_gemm_ds_pair(const __nv_bfloat16* G, const __nv_bfloat16* wmid_full, __nv_bfloat16* dS, int rows, cudaStream_t st);
int launch_gemm_dw_pair(const __nv_bfloat16* G, const __nv_bfloat16* S, float* const dW[kNumHeads], float* const db[kNumHeads],
                        int rows, cudaStream_t st);

// loss.cu
size_t pair_loss_workspace_bytes(int batch, int n);
int launch_pair_loss_fwd(int batch, int n, const float* const logits[kNumHeads], const int64_t* const tags[kNumHeads],
                         const float* class_w, const float* ratio, float* out6, void* ws, cudaStream_t st);
int launch_pair_loss_bwd(int batch, int n, const float* const logits[kNumHeads], const int64_t* const tags[kNumHeads],
                         const float* class_w, const float* ratio, const float* grad_out, const void* ws,
                         float* const dlogits[kNumHeads], cudaStream_t st);
// final reduction of per-CTA partial sums written by another kernel (K2 with the fused loss): out6 + the (sum w nll, sum w)
// pairs the backward pass normalises with
int launch_pair_loss_finalize(const float* ratio, float* out6, void* ws, int nblocks, cudaStream_t st);
const double* pair_loss_final_ptr(const void* ws);
double* pair_loss_partial_ptr(void* ws);
int pair_loss_max_blocks();
int launch_scatter_tags(const int32_t* spots, int64_t num_spots, int batch, int n, int64_t* tags, cudaStream_t st);

// ohem.cu
size_t pair_loss_ohem_workspace_bytes(int batch, int n);
int launch_pair_loss_ohem_fwd(int batch, int n, const float* const logits[kNumHeads], const int64_t* const tags[kNumHeads],
                              const float* class_w, const float* ratio, int num_hard_pos, int num_hard_neg, float* out6,
                              void* workspace, cudaStream_t st);
int launch_pair_loss_ohem_bwd(int batch, int n, const float* const logits[kNumHeads], const int64_t* const tags[kNumHeads],
                              const float* class_w, const float* ratio, const float* grad_out, const void* workspace,
                              float* const dlogits[kNumHeads], cudaStream_t st);

// train.cu
size_t heads_bwd_workspace_bytes(const peneo_dims& dm, int prec, int batch, int n);
struct FusedLossBwd;
int launch_heads_bwd(const peneo_dims& dm, int prec, const void* pack, const void* x, int x_dtype, int64_t x_row_stride,
                     int batch, int n, const float* const dlogits[kNumHeads], const peneo_grads& gr, float* dx,
                     void* workspace, cudaStream_t st, const DropSpec* drop = nullptr,
                     const FusedLossBwd* fused = nullptr);

// decode.cu
size_t decode_spots_workspace_bytes(int batch, int n);
int launch_decode_spots(int batch, int n, const void* const in[kNumHeads], int in_dtype, int cap, int32_t* spot_p,
                        int32_t* spot_tag, float* spot_score, int32_t* counts, void* ws, cudaStream_t st);
// ordered compaction of the per-tile spot slots K2's spots-only epilogue wrote -> the same compact lists decode_spots gives
size_t heads_spots_workspace_bytes(int batch, int n);
int launch_gather_tile_spots(int batch, int n, void* ws, int cap, int32_t* spot_p, int32_t* spot_tag, float* spot_score,
                             int32_t* counts, cudaStream_t st);
size_t decode_resolve_doc_ints(int n, int cap);
size_t decode_resolve_workspace_bytes(int batch, int n);
int launch_decode_resolve(int batch, int n, int cap, const int32_t* spot_p, const int32_t* spot_tag,
                          const float* spot_score, const int32_t* counts, int decode_gt, float score_thresh,
                          int32_t* out, void* ws, cudaStream_t st);

// selftest.cu
int run_selftest(uint32_t* failed_mask, char* report, size_t report_bytes);
int run_probe_rates(double* out, int n_out);

// tensor-map helper (gemm_tc.cu)
struct TensorMap2D;
int make_tensor_map_bf16(void* tmap_out, const void* gptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                         uint32_t box_inner, uint32_t box_outer);
int make_tensor_map_bf16_sw64(void* tmap_out, const void* gptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                              uint32_t box_inner, uint32_t box_outer);
int make_tensor_map_f32(void* tmap_out, const void* gptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                        uint32_t box_inner, uint32_t box_outer);

// gemm_tf32.cu : fp32 GEMM on tcgen05 kind::tf32.  C[M, N] (op)= op(A) op(B) (+ bias), optional C2 = Dropout(SiLU(C)).
//   ta: A stored [K, M] (else [M, K]) ; tb: B stored [N, K] (else [K, N]) ; mode 0 store, 1 add, 2 split-K atomic add
struct Tf32Gemm {
  bool ta = false, tb = false;
  const float* A = nullptr;
  int64_t lda = 0;
  const float* B = nullptr;
  int64_t ldb = 0;
  float* C = nullptr;
  int64_t ldc = 0;
  int M = 0, N = 0, K = 0, mode = 0;
  const float* bias = nullptr;
  float* C2 = nullptr;
  int64_t ldc2 = 0;
  uint32_t drop_thresh = 0, drop_key = 0, row0 = 0;
  float drop_scale = 1.f;
};
bool gemm_tf32_supported(const Tf32Gemm& g);
int launch_gemm_tf32(const Tf32Gemm& g, cudaStream_t st);

}  // namespace peneo
