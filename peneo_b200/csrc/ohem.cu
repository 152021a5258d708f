// Pairwise loss with online hard-example mining — CrossEntropyLossOHEM.forward with
// num_hard_positive / num_hard_negative set (model/custom_loss.py:204-288), as PEneoDecoder builds it
// (model/peneo_decoder.py:304-313: only the num_hard_* knobs, reduction "mean", no random sampling).
//
// Per head, over all M = batch * P elements:
//   ce[m]   = w[t_m] * (logsumexp(x_m) - x_m[t_m])
//   split by t_m == 0 into negatives / positives (original order), sort each side descending
//   k_side  = min(n_side, num_hard_side)
//   k <= 0          : the side keeps everything
//   0 < k < n_side  : the side keeps sorted[idx[q]] for q < k  — the reference indexes the SORTED losses with
//                     positions of the UNSORTED array (custom_loss.py:262-263, 272-273); reproduced as is
//   loss    = (sum kept_pos + sum kept_neg) / (k_pos + k_neg)      (raw k values, -1 included)
// The sort is cub::DeviceRadixSort (the reference calls torch.sort); the kept set is returned as a
// byte mask so that the backward pass is one streaming kernel.
#include <cub/cub.cuh>

#include "common.cuh"
#include "kernels.h"

namespace peneo {

namespace {

constexpr int kBlocks = 296;

template <int C>
__device__ __forceinline__ float weighted_nll(const float* x, int t, const float* w) {
  float m = x[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) s += expf(x[c] - m);
  return w[t] * (m + logf(s) - x[t]);
}

// ce[m] and side flags (1 = positive)
__global__ void __launch_bounds__(256) ohem_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ tags,
                                                      int C, int64_t rows, float w0, float w1, float w2,
                                                      float* __restrict__ ce, int32_t* __restrict__ is_pos) {
  const float w[3] = {w0, w1, w2};
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    bool bad;  // target outside [0, C): NaN loss instead of an out-of-bounds read (common.cuh: checked_tag)
    const int t = checked_tag(tags[r], C, bad);
    float x[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) x[c] = logits[r * C + c];
    ce[r] = bad ? NAN : (C == 2 ? weighted_nll<2>(x, t, w) : weighted_nll<3>(x, t, w));
    is_pos[r] = bad || t != 0;
  }
}

// stable split into the two sides: value + original element index at the compacted position
__global__ void __launch_bounds__(256) ohem_split_kernel(const float* __restrict__ ce, const int32_t* __restrict__ is_pos,
                                                         const int32_t* __restrict__ pos_rank, int64_t rows,
                                                         float* __restrict__ val, int32_t* __restrict__ orig,
                                                         int32_t* __restrict__ iota) {
  // val / orig / iota hold two halves of `rows` entries: positives at [0, np), negatives at [rows, rows + nn),
  // each in original element order
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int pr = pos_rank[r];  // positives strictly before r
    if (is_pos[r]) {
      val[pr] = ce[r], orig[pr] = static_cast<int32_t>(r), iota[pr] = pr;
    } else {
      const int64_t nr = r - pr;  // negatives strictly before r
      val[rows + nr] = ce[r], orig[rows + nr] = static_cast<int32_t>(r), iota[rows + nr] = static_cast<int32_t>(nr);
    }
  }
}

struct SideSel {
  const float* sorted_val;    // descending
  const int32_t* sorted_idx;  // compacted index of each sorted value
  const int32_t* orig;        // compacted index -> element
  int64_t n;                  // elements on this side (device-known count mirrored on host)
  int64_t k;                  // min(n, num_hard)
};

// marks the kept elements of one side and accumulates their loss
__global__ void __launch_bounds__(256) ohem_select_kernel(SideSel s, uint8_t* __restrict__ keep, double* __restrict__ sum) {
  double acc = 0.0;
  const bool all = s.k <= 0 || s.k >= s.n;
  const int64_t cnt = all ? s.n : s.k;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = all ? q : s.sorted_idx[q];  // rank whose value is kept
    acc += static_cast<double>(s.sorted_val[r]);
    keep[s.orig[s.sorted_idx[r]]] = 1;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(sum, acc);
}

struct RatioArr {
  float r[kNumHeads];
};
__global__ void ohem_final_kernel(const double* __restrict__ sums, const double* __restrict__ denom, RatioArr ratio,
                                  float* __restrict__ out6) {
  if (threadIdx.x != 0) return;
  double total = 0.0;
  for (int h = 0; h < kNumHeads; ++h) {
    const double l = (sums[2 * h] + sums[2 * h + 1]) / denom[h];
    out6[h] = static_cast<float>(l);
    total += static_cast<double>(ratio.r[h]) * l;
  }
  out6[5] = static_cast<float>(total);
}

// dlogits = scale_h * keep * w[t] * (softmax - onehot)
__global__ void __launch_bounds__(256) ohem_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ tags,
                                                       const uint8_t* __restrict__ keep, int C, int64_t rows, float w0,
                                                       float w1, float w2, const float* __restrict__ grad_out6, int h,
                                                       float ratio, float inv_denom, float* __restrict__ dlogits) {
  const float w[3] = {w0, w1, w2};
  const float scale = (grad_out6[5] * ratio + grad_out6[h]) * inv_denom;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    float x[3] = {0.f, 0.f, 0.f}, e[3];
    for (int c = 0; c < C; ++c) x[c] = logits[r * C + c];
    float m = x[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) e[c] = expf(x[c] - m), s += e[c];
    bool bad;
    const int t = checked_tag(tags[r], C, bad);
    const float g = bad ? NAN : (keep[r] ? scale * w[t] : 0.f);
    for (int c = 0; c < C; ++c) dlogits[r * C + c] = g * (e[c] / s - (c == t ? 1.f : 0.f));
  }
}

struct Layout {
  int64_t rows;
  size_t off_keep;   // uint8 [5][rows]      (kept masks: read by the backward pass)
  size_t off_denom;  // double [5]           (k_pos + k_neg per head)
  size_t off_sums;   // double [10]
  size_t off_ce, off_flag, off_rank, off_val, off_val2, off_orig, off_idx, off_idx2;
  size_t off_tmp, tmp_bytes;
  size_t total;
};

Layout make_layout(int batch, int n) {
  Layout L{};
  L.rows = (int64_t)batch * pair_count(n);
  const size_t M = static_cast<size_t>(L.rows);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off = align_up(off + bytes, 1024);
    return at;
  };
  L.off_keep = take(kNumHeads * M);
  L.off_denom = take(kNumHeads * sizeof(double));
  L.off_sums = take(2 * kNumHeads * sizeof(double));
  L.off_ce = take(M * 4), L.off_flag = take(M * 4), L.off_rank = take(M * 4);
  // two halves: positives use [0, rows), negatives [rows, 2 rows)
  L.off_val = take(2 * M * 4), L.off_val2 = take(2 * M * 4), L.off_orig = take(2 * M * 4);
  L.off_idx = take(2 * M * 4), L.off_idx2 = take(2 * M * 4);
  L.tmp_bytes = (32u << 20) + M * 8;
  L.off_tmp = take(L.tmp_bytes);
  L.total = off;
  return L;
}

}  // namespace

size_t pair_loss_ohem_workspace_bytes(int batch, int n) { return make_layout(batch, n).total; }

int launch_pair_loss_ohem_fwd(int batch, int n, const float* const logits[kNumHeads], const int64_t* const tags[kNumHeads],
                              const float* class_w, const float* ratio, int num_hard_pos, int num_hard_neg, float* out6,
                              void* workspace, cudaStream_t st) {
  PENEO_REQUIRE(batch >= 1 && n >= 1, "pair_loss_ohem: empty batch");
  const Layout L = make_layout(batch, n);
  PENEO_REQUIRE(L.rows < (1ll << 31) / 2, "pair_loss_ohem: too many elements for 32-bit indices");
  char* ws = static_cast<char*>(workspace);
  const int64_t M = L.rows;
  uint8_t* keep = reinterpret_cast<uint8_t*>(ws + L.off_keep);
  double* denom = reinterpret_cast<double*>(ws + L.off_denom);
  double* sums = reinterpret_cast<double*>(ws + L.off_sums);
  float* ce = reinterpret_cast<float*>(ws + L.off_ce);
  int32_t* flag = reinterpret_cast<int32_t*>(ws + L.off_flag);
  int32_t* rank = reinterpret_cast<int32_t*>(ws + L.off_rank);
  float* val = reinterpret_cast<float*>(ws + L.off_val);
  float* val2 = reinterpret_cast<float*>(ws + L.off_val2);
  int32_t* orig = reinterpret_cast<int32_t*>(ws + L.off_orig);
  int32_t* idx = reinterpret_cast<int32_t*>(ws + L.off_idx);
  int32_t* idx2 = reinterpret_cast<int32_t*>(ws + L.off_idx2);
  void* tmp = ws + L.off_tmp;

  PENEO_CUDA_TRY(cudaMemsetAsync(keep, 0, (size_t)kNumHeads * M, st));
  PENEO_CUDA_TRY(cudaMemsetAsync(sums, 0, 2 * kNumHeads * sizeof(double), st));
  double h_denom[kNumHeads];
  for (int h = 0; h < kNumHeads; ++h) {
    const int C = head_classes(h);
    ohem_ce_kernel<<<kBlocks, 256, 0, st>>>(logits[h], tags[h], C, M, class_w[0], class_w[1], class_w[2], ce, flag);
    PENEO_CUDA_TRY(cudaGetLastError());
    size_t need = 0;
    PENEO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, flag, rank, static_cast<int>(M), st));
    PENEO_REQUIRE(need <= L.tmp_bytes, "pair_loss_ohem: scan scratch %zu > %zu", need, L.tmp_bytes);
    PENEO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, flag, rank, static_cast<int>(M), st));
    // number of positives: needed on the host to size the two sorts (the reference has the same
    // device->host dependency: boolean-mask indexing, custom_loss.py:234-236)
    int32_t last_rank = 0, last_flag = 0;
    PENEO_CUDA_TRY(cudaMemcpyAsync(&last_rank, rank + (M - 1), 4, cudaMemcpyDeviceToHost, st));
    PENEO_CUDA_TRY(cudaMemcpyAsync(&last_flag, flag + (M - 1), 4, cudaMemcpyDeviceToHost, st));
    PENEO_CUDA_TRY(cudaStreamSynchronize(st));
    const int64_t np = (int64_t)last_rank + last_flag, nn = M - np;
    ohem_split_kernel<<<kBlocks, 256, 0, st>>>(ce, flag, rank, M, val, orig, idx);
    PENEO_CUDA_TRY(cudaGetLastError());
    const int64_t side_n[2] = {np, nn}, side_cfg[2] = {num_hard_pos, num_hard_neg};
    int64_t ksum = 0;
    for (int side = 0; side < 2; ++side) {
      const int64_t ns = side_n[side], base = side == 0 ? 0 : M;
      const int64_t k = std::min<int64_t>(ns, side_cfg[side]);
      ksum += k;
      if (ns == 0) continue;
      SideSel s{};
      s.n = ns, s.k = k, s.orig = orig + base;
      if (k <= 0 || k >= ns) {
        // everything is kept: no sort needed; "sorted" order = compacted order
        s.sorted_val = val + base, s.sorted_idx = idx + base;
      } else {
        cub::DoubleBuffer<float> keys(val + base, val2 + base);
        cub::DoubleBuffer<int32_t> vals(idx + base, idx2 + base);
        need = 0;
        PENEO_CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, need, keys, vals, static_cast<int>(ns), 0, 32, st));
        PENEO_REQUIRE(need <= L.tmp_bytes, "pair_loss_ohem: sort scratch %zu > %zu", need, L.tmp_bytes);
        PENEO_CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(tmp, need, keys, vals, static_cast<int>(ns), 0, 32, st));
        s.sorted_val = keys.Current(), s.sorted_idx = vals.Current();
      }
      const int blocks = static_cast<int>(std::min<int64_t>(kBlocks, ((k <= 0 || k >= ns ? ns : k) + 255) / 256));
      ohem_select_kernel<<<blocks, 256, 0, st>>>(s, keep + (size_t)h * M, sums + 2 * h + side);
      PENEO_CUDA_TRY(cudaGetLastError());
    }
    h_denom[h] = static_cast<double>(ksum);
  }
  PENEO_CUDA_TRY(cudaMemcpyAsync(denom, h_denom, sizeof h_denom, cudaMemcpyHostToDevice, st));
  RatioArr ra;
  for (int h = 0; h < kNumHeads; ++h) ra.r[h] = ratio ? ratio[h] : 1.f;
  ohem_final_kernel<<<1, 32, 0, st>>>(sums, denom, ra, out6);
  PENEO_CUDA_TRY(cudaGetLastError());
  // h_denom lives on this stack frame: the H2D copy above must have consumed it before we return
  PENEO_CUDA_TRY(cudaStreamSynchronize(st));
  return PENEO_OK;
}

int launch_pair_loss_ohem_bwd(int batch, int n, const float* const logits[kNumHeads], const int64_t* const tags[kNumHeads],
                              const float* class_w, const float* ratio, const float* grad_out, const void* workspace,
                              float* const dlogits[kNumHeads], cudaStream_t st) {
  PENEO_REQUIRE(batch >= 1 && n >= 1, "pair_loss_ohem_bwd: empty batch");
  const Layout L = make_layout(batch, n);
  const char* ws = static_cast<const char*>(workspace);
  const int64_t M = L.rows;
  const uint8_t* keep = reinterpret_cast<const uint8_t*>(ws + L.off_keep);
  double h_denom[kNumHeads];
  PENEO_CUDA_TRY(cudaMemcpyAsync(h_denom, ws + L.off_denom, sizeof h_denom, cudaMemcpyDeviceToHost, st));
  PENEO_CUDA_TRY(cudaStreamSynchronize(st));
  for (int h = 0; h < kNumHeads; ++h) {
    ohem_bwd_kernel<<<kBlocks, 256, 0, st>>>(logits[h], tags[h], keep + (size_t)h * M, head_classes(h), M, class_w[0],
                                             class_w[1], class_w[2], grad_out, h, ratio ? ratio[h] : 1.f,
                                             static_cast<float>(1.0 / h_denom[h]), dlogits[h]);
    PENEO_CUDA_TRY(cudaGetLastError());
  }
  return PENEO_OK;
}

}  // namespace peneo
