// Spot classification shared by the stand-alone spot-extraction kernel (decode.cu, K3) and the spots-only epilogue of
// K2 (pair_heads_tc2.cu): softmax over C classes the way ATen does it (exp(x - max) / sum in fp32, result rounded to
// the tensor dtype), then argmax of the *probabilities* (first maximum) and its value
// (HandshakingTaggingScheme.get_spots_from_shaking_tag, model/peneo_decoder.py:98-101).
#pragma once
#include "common.cuh"

namespace peneo {

template <int DT>
__device__ __forceinline__ float round_like(float v) {
  if constexpr (DT == PENEO_DT_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  else if constexpr (DT == PENEO_DT_F16) return __half2float(__float2half_rn(v));
  else return v;
}

template <int DT, int C>
__device__ __forceinline__ void classify_vals(const float* x, int& pred, float& score) {
  float m = x[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
  float e[C], sum = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    e[c] = expf(x[c] - m);
    sum += e[c];
  }
  pred = 0;
  score = round_like<DT>(e[0] / sum);
#pragma unroll
  for (int c = 1; c < C; ++c) {
    const float pc = round_like<DT>(e[c] / sum);
    if (pc > score) score = pc, pred = c;
  }
}

// Per-tile spot slots written by K2's spots-only epilogue and gathered (ordered compaction) by decode.cu:
// slot = (tile * 5 + head) * 4 + quadrant covers the 32 consecutive pairs [128 tile + 32 quadrant, + 32) of the
// batch-flat pair list; cnt[slot] entries, in increasing pair order, at meta / score [slot * 32 ..]:
// meta = row inside the tile | tag << 8.
struct TileSpots {
  int32_t* cnt;
  int32_t* meta;
  float* score;
};
inline size_t tile_spots_bytes(int64_t tiles) { return static_cast<size_t>(tiles) * kNumHeads * 4 * (4 + 32 * 8); }
inline TileSpots tile_spots_carve(void* ws, int64_t tiles) {
  TileSpots t;
  t.cnt = static_cast<int32_t*>(ws);
  t.meta = t.cnt + tiles * kNumHeads * 4;
  t.score = reinterpret_cast<float*>(t.meta + tiles * kNumHeads * 4 * 32);
  return t;
}

}  // namespace peneo
