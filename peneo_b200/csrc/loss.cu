// Class-weighted pairwise cross entropy (OHEM off) and its gradient w.r.t. the logits, plus the
// sparse->dense tag scatter.  Follows model/custom_loss.py:189-202 (F.cross_entropy with class
// weights, reduction="mean" => sum_m w[t_m] nll_m / sum_m w[t_m] over ALL batch*P elements, no
// padding mask) and model/peneo_decoder.py:375-428 (ratio-weighted sum of the five sub-losses).
// Reductions are deterministic: per-CTA partial sums in fp64, reduced in index order.
#include "common.cuh"
#include "kernels.h"

namespace peneo {

constexpr int kLossBlocks = 296;  // 2 x 148 SMs, grid-stride

struct LossArgs {
  const float* logits[kNumHeads];
  const int64_t* tags[kNumHeads];
  float* dlogits[kNumHeads];
  int64_t rows;  // batch * P
  float w[3];
  float ratio[kNumHeads];
  double* partial;  // [5][kLossBlocks][2]
  double* final_;   // [5][2] : sum w*nll, sum w
  float* out6;
  const float* grad_out;
};

template <int C>
__device__ __forceinline__ void row_lse(const float* x, float& lse, float (&e)[3], float& m) {
  m = x[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    e[c] = expf(x[c] - m);
    s += e[c];
  }
  lse = m + logf(s);
}

__global__ void __launch_bounds__(256) pair_loss_partial_kernel(const LossArgs a) {
  const int h = blockIdx.y, C = head_classes(h);
  double sl = 0.0, sw = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.rows; r += (int64_t)gridDim.x * blockDim.x) {
    float x[3] = {0.f, 0.f, 0.f}, e[3], lse, m;
    for (int c = 0; c < C; ++c) x[c] = a.logits[h][r * C + c];
    if (C == 2)
      row_lse<2>(x, lse, e, m);
    else
      row_lse<3>(x, lse, e, m);
    bool bad;
    const int t = checked_tag(a.tags[h][r], C, bad);
    const float w = a.w[t];
    sl += bad ? static_cast<double>(NAN) : static_cast<double>(w * (lse - x[t]));
    sw += static_cast<double>(w);
  }
  __shared__ double red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  if (lane == 0) red[0][warp] = sl, red[1][warp] = sw;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tl = 0.0, tw = 0.0;
    for (int w2 = 0; w2 < 8; ++w2) tl += red[0][w2], tw += red[1][w2];
    a.partial[((int64_t)h * gridDim.x + blockIdx.x) * 2 + 0] = tl;
    a.partial[((int64_t)h * gridDim.x + blockIdx.x) * 2 + 1] = tw;
  }
}

// One warp per head: lane l adds the partials of blocks l, l + 32, ... in order, then a fixed shuffle tree —
// deterministic, and 32 loads in flight instead of one thread walking 2 * 5 * nblocks doubles.
__global__ void __launch_bounds__(32 * kNumHeads) pair_loss_final_kernel(const LossArgs a, int nblocks) {
  __shared__ double s_loss[kNumHeads];
  const int h = threadIdx.x / 32, lane = threadIdx.x % 32;
  double tl = 0.0, tw = 0.0;
  for (int b = lane; b < nblocks; b += 32) {
    tl += a.partial[((int64_t)h * nblocks + b) * 2 + 0];
    tw += a.partial[((int64_t)h * nblocks + b) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tl += __shfl_xor_sync(0xffffffffu, tl, o);
    tw += __shfl_xor_sync(0xffffffffu, tw, o);
  }
  if (lane == 0) {
    a.final_[h * 2 + 0] = tl, a.final_[h * 2 + 1] = tw;
    const double lh = tl / tw;
    a.out6[h] = static_cast<float>(lh);
    s_loss[h] = static_cast<double>(a.ratio[h]) * lh;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int k = 0; k < kNumHeads; ++k) total += s_loss[k];
    a.out6[5] = static_cast<float>(total);
  }
}

__global__ void __launch_bounds__(256) pair_loss_bwd_kernel(const LossArgs a) {
  const int h = blockIdx.y, C = head_classes(h);
  // grad_out = d L / d (loss_0 .. loss_4, total): the total is sum_h ratio_h loss_h (model/peneo_decoder.py:399-420)
  const float scale = (a.grad_out[5] * a.ratio[h] + a.grad_out[h]) / static_cast<float>(a.final_[h * 2 + 1]);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.rows; r += (int64_t)gridDim.x * blockDim.x) {
    float x[3] = {0.f, 0.f, 0.f}, e[3], lse, m;
    for (int c = 0; c < C; ++c) x[c] = a.logits[h][r * C + c];
    if (C == 2)
      row_lse<2>(x, lse, e, m);
    else
      row_lse<3>(x, lse, e, m);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += e[c];
    bool bad;
    const int t = checked_tag(a.tags[h][r], C, bad);
    const float g = bad ? NAN : scale * a.w[t];
    for (int c = 0; c < C; ++c) a.dlogits[h][r * C + c] = g * (e[c] / s - (c == t ? 1.f : 0.f));
  }
}

size_t pair_loss_workspace_bytes(int, int) { return (size_t)(kNumHeads * kLossBlocks * 2 + kNumHeads * 2) * sizeof(double); }

static void fill_args(LossArgs& a, int batch, int n, const float* const logits[kNumHeads],
                      const int64_t* const tags[kNumHeads], const float* class_w, const float* ratio, void* ws) {
  for (int h = 0; h < kNumHeads; ++h) {
    a.logits[h] = logits[h], a.tags[h] = tags[h];
    a.ratio[h] = ratio ? ratio[h] : 1.f;
  }
  a.rows = (int64_t)batch * pair_count(n);
  for (int c = 0; c < 3; ++c) a.w[c] = class_w[c];
  a.partial = static_cast<double*>(ws);
  a.final_ = a.partial + kNumHeads * kLossBlocks * 2;
}

int launch_pair_loss_fwd(int batch, int n, const float* const logits[kNumHeads], const int64_t* const tags[kNumHeads],
                         const float* class_w, const float* ratio, float* out6, void* ws, cudaStream_t st) {
  PENEO_REQUIRE(batch >= 1 && n >= 1, "pair_loss_fwd: empty batch");
  LossArgs a{};
  fill_args(a, batch, n, logits, tags, class_w, ratio, ws);
  a.out6 = out6;
  pair_loss_partial_kernel<<<dim3(kLossBlocks, kNumHeads), 256, 0, st>>>(a);
  PENEO_CUDA_TRY(cudaGetLastError());
  pair_loss_final_kernel<<<1, 32 * kNumHeads, 0, st>>>(a, kLossBlocks);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

int launch_pair_loss_finalize(const float* ratio, float* out6, void* ws, int nblocks, cudaStream_t st) {
  PENEO_REQUIRE(nblocks >= 1 && nblocks <= kLossBlocks, "pair_loss_finalize: %d partial blocks (max %d)", nblocks, kLossBlocks);
  LossArgs a{};
  for (int h = 0; h < kNumHeads; ++h) a.ratio[h] = ratio ? ratio[h] : 1.f;
  a.partial = static_cast<double*>(ws);
  a.final_ = a.partial + kNumHeads * kLossBlocks * 2;
  a.out6 = out6;
  pair_loss_final_kernel<<<1, 32 * kNumHeads, 0, st>>>(a, nblocks);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}
const double* pair_loss_final_ptr(const void* ws) { return static_cast<const double*>(ws) + kNumHeads * kLossBlocks * 2; }
double* pair_loss_partial_ptr(void* ws) { return static_cast<double*>(ws); }
int pair_loss_max_blocks() { return kLossBlocks; }

int launch_pair_loss_bwd(int batch, int n, const float* const logits[kNumHeads], const int64_t* const tags[kNumHeads],
                         const float* class_w, const float* ratio, const float* grad_out, const void* ws,
                         float* const dlogits[kNumHeads], cudaStream_t st) {
  PENEO_REQUIRE(batch >= 1 && n >= 1, "pair_loss_bwd: empty batch");
  LossArgs a{};
  fill_args(a, batch, n, logits, tags, class_w, ratio, const_cast<void*>(ws));
  for (int h = 0; h < kNumHeads; ++h) a.dlogits[h] = dlogits[h];
  a.grad_out = grad_out;
  pair_loss_bwd_kernel<<<dim3(kLossBlocks, kNumHeads), 256, 0, st>>>(a);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
// tags[b, p(i, j)] = tag, last spot wins (model/peneo_decoder.py:68-72)
// ------------------------------------------------------------------------------------------------
// Cell of spot (b, i, j), or -1 when it lies outside the batch / the document (the host wrapper validates and
// raises like the reference's list indexing would; the kernel itself never writes out of bounds).  A spot below
// the diagonal (i > j) lands in cell 0 of its document: the reference looks it up in an N x N table that is zero
// everywhere below the diagonal (model/peneo_decoder.py:55-59, 70).
__device__ __forceinline__ int64_t scatter_cell(const int32_t* spots, int64_t s, int batch, int n, int64_t pairs) {
  const int b = spots[4 * s], i = spots[4 * s + 1], j = spots[4 * s + 2];
  if (b < 0 || b >= batch || i < 0 || i >= n || j < 0 || j >= n) return -1;
  return b * pairs + (j < i ? 0 : row_start(i, n) + (j - i));
}
__global__ void scatter_claim_kernel(const int32_t* __restrict__ spots, int64_t num, int batch, int n, int64_t pairs,
                                     long long* tags) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= num) return;
  const int64_t cell = scatter_cell(spots, s, batch, n, pairs);
  if (cell >= 0) atomicMin(&tags[cell], -(s + 1));
}
__global__ void scatter_write_kernel(const int32_t* __restrict__ spots, int64_t num, int batch, int n, int64_t pairs,
                                     long long* tags) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= num) return;
  const int64_t cell = scatter_cell(spots, s, batch, n, pairs);
  if (cell >= 0 && tags[cell] == -(s + 1)) tags[cell] = spots[4 * s + 3];
}

int launch_scatter_tags(const int32_t* spots, int64_t num_spots, int batch, int n, int64_t* tags, cudaStream_t st) {
  const int64_t pairs = pair_count(n);
  PENEO_CUDA_TRY(cudaMemsetAsync(tags, 0, (size_t)batch * pairs * sizeof(int64_t), st));
  if (num_spots == 0) return PENEO_OK;
  const unsigned blocks = static_cast<unsigned>((num_spots + 255) / 256);
  scatter_claim_kernel<<<blocks, 256, 0, st>>>(spots, num_spots, batch, n, pairs, reinterpret_cast<long long*>(tags));
  PENEO_CUDA_TRY(cudaGetLastError());
  scatter_write_kernel<<<blocks, 256, 0, st>>>(spots, num_spots, batch, n, pairs, reinterpret_cast<long long*>(tags));
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
