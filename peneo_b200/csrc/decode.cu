// Decode kernels.
//   K3 decode_spots   : per pair softmax -> argmax -> max-prob, ordered compaction of pred != 0
//                       (HandshakingTaggingScheme.get_spots_from_shaking_tag, model/peneo_decoder.py:98-114)
//   K4 decode_resolve : 1-1 map resolution (parse_matrix_spots, pipeline/decode.py:9-69) and the
//                       key/value chain walk of sample_decode_peneo (pipeline/decode.py:216-368)
// Both are HBM/latency-bound integer work: K3 reads every logit exactly once.
#include "classify.cuh"
#include "common.cuh"
#include "kernels.h"

namespace peneo {

#ifndef PENEO_K3_WARP_PAIRS
#define PENEO_K3_WARP_PAIRS 512
#endif
#ifndef PENEO_K3_UNROLL
#define PENEO_K3_UNROLL 2
#endif
constexpr int kK3Unroll = PENEO_K3_UNROLL;
constexpr int kSpotWarps = 4096 / PENEO_K3_WARP_PAIRS;  // warps per CTA: 4096 pairs (36 KB of spot buffers) per CTA
#ifndef PENEO_K3_LOOK_WINDOWS
#define PENEO_K3_LOOK_WINDOWS 1  // measured on B200: 1 window 70 us, 4 windows 76 us, 8 windows 85 us (polling traffic on the status lines costs more than the extra rounds)
#endif
constexpr int kLookWindows = PENEO_K3_LOOK_WINDOWS;  // windows of 32 predecessors read per look-back round
constexpr int kSpotWarpPairs = PENEO_K3_WARP_PAIRS;  // pairs per warp chunk (8 warps per CTA)

// ------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ float load_as_float(const void* base, int64_t idx) {
  if constexpr (DT == PENEO_DT_F32) return static_cast<const float*>(base)[idx];
  else if constexpr (DT == PENEO_DT_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[idx]);
  else if constexpr (DT == PENEO_DT_F16) return __half2float(static_cast<const __half*>(base)[idx]);
  else return 0.f;
}
template <int DT, int C>
__device__ __forceinline__ void classify(const void* base, int64_t row, int& pred, float& score) {
  float x[C];
#pragma unroll
  for (int c = 0; c < C; ++c) x[c] = load_as_float<DT>(base, row * C + c);
  classify_vals<DT, C>(x, pred, score);
}

struct SpotArgs {
  const void* in[kNumHeads];
  int32_t batch, n, pairs, cap, chunks;
  int32_t* spot_p;
  int32_t* spot_tag;
  float* spot_score;
  int32_t* counts;
  int32_t* status;  // [batch*5*chunks] : (value << 2) | flag, 0 = not ready
};

// fp32 logits, 16-byte aligned document: one warp iteration scans 128 pairs, 4 consecutive pairs per lane, read as
// C float4 (every byte of the logits crosses the LSU once, in full 16-byte requests).  Same fast reject, same
// ordered compaction (order inside the iteration = lane major, then the lane's 4 pairs).
template <int C>
__device__ __forceinline__ int scan_pairs_vec4(const float* __restrict__ doc, int pairs, int p0, int lane, int32_t* bp,
                                               float* bs, uint8_t* bt) {
  int cnt = 0;
#pragma unroll(kK3Unroll)
  for (int it = 0; it < kSpotWarpPairs / 128; ++it) {
    const int pl = p0 + it * 128 + lane * 4;
    float x[4 * C];
    if (pl + 3 < pairs) {
      const float4* src = reinterpret_cast<const float4*>(doc + (int64_t)pl * C);
#pragma unroll
      for (int v = 0; v < C; ++v) {
        const float4 t = __ldg(src + v);
        x[4 * v] = t.x, x[4 * v + 1] = t.y, x[4 * v + 2] = t.z, x[4 * v + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4 * C; ++e) x[e] = (pl + e / C < pairs) ? doc[(int64_t)pl * C + e] : 0.f;
    }
    int pred[4];
    float score[4];
    unsigned mask[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pred[k] = 0, score[k] = 1.f;
      const float x0 = x[k * C];
      float xo = x[k * C + 1];
      if (C == 3) xo = fmaxf(xo, x[k * C + 2]);
      if (pl + k < pairs && !(x0 >= xo)) classify_vals<PENEO_DT_F32, C>(x + k * C, pred[k], score[k]);
      mask[k] = __ballot_sync(0xffffffffu, pred[k] != 0);
    }
    const unsigned lt = (1u << lane) - 1u;
    int pos = cnt + __popc(mask[0] & lt) + __popc(mask[1] & lt) + __popc(mask[2] & lt) + __popc(mask[3] & lt);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (pred[k] != 0) bp[pos] = pl + k, bs[pos] = score[k], bt[pos] = static_cast<uint8_t>(pred[k]), ++pos;
    cnt += __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]);
  }
  return cnt;
}

// One WARP per 512 consecutive pairs of one (document, head) list; the warps of a CTA are independent (no block-wide
// barrier): scan -> publish the warp's count -> decoupled look-back over the preceding warp chunks of the list ->
// write the spots at the exclusive prefix.  Warp chunk w of a list runs in CTA w / 8 (blockIdx.x), i.e. chunks start
// in index order (CTAs are dispatched in block-index order, the assumption CUB's single-pass scan makes as well), so
// the look-back never waits for a warp that is not resident yet.
template <int DT>
__global__ void __launch_bounds__(32 * kSpotWarps) decode_spots_kernel(const SpotArgs a) {
  __shared__ int32_t buf_p[kSpotWarps][kSpotWarpPairs];
  __shared__ float buf_s[kSpotWarps][kSpotWarpPairs];
  __shared__ uint8_t buf_t[kSpotWarps][kSpotWarpPairs];
  const int h = blockIdx.y, b = blockIdx.z, list = b * kNumHeads + h;
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int chunk = blockIdx.x * kSpotWarps + warp;  // warp chunk inside the list
  if (chunk >= a.chunks) return;
  const int C = head_classes(h);
  const char* base = static_cast<const char*>(a.in[h]);
  const int64_t doc_row0 = (int64_t)b * a.pairs;
  int cnt = 0;
  const int p0 = chunk * kSpotWarpPairs;
  bool scanned = false;
  if constexpr (DT == PENEO_DT_F32) {
    const float* doc = reinterpret_cast<const float*>(base) + doc_row0 * C;
    if ((reinterpret_cast<uintptr_t>(doc) & 15) == 0) {
      cnt = C == 2 ? scan_pairs_vec4<2>(doc, a.pairs, p0, lane, buf_p[warp], buf_s[warp], buf_t[warp])
                   : scan_pairs_vec4<3>(doc, a.pairs, p0, lane, buf_p[warp], buf_s[warp], buf_t[warp]);
      scanned = true;
    }
  }
#pragma unroll 4
  for (int it = 0; !scanned && it < kSpotWarpPairs / 32; ++it) {
    const int p = p0 + it * 32 + lane;
    int pred = 0;
    float score = 1.f;
    if (p < a.pairs) {
      if constexpr (DT == PENEO_DT_I64) {
        pred = static_cast<int>(reinterpret_cast<const int64_t*>(base)[doc_row0 + p]);
      } else {
        // Fast reject: when the class-0 logit is >= every other logit, its probability is >= theirs after any
        // monotone rounding and argmax returns the first maximum, i.e. class 0 -> not a spot.  Only the rare
        // rows that can be spots pay for the exact softmax (exp, sum, divide, round) of the reference.
        const float x0 = load_as_float<DT>(base, (doc_row0 + p) * C);
        float xo = load_as_float<DT>(base, (doc_row0 + p) * C + 1);
        if (C == 3) xo = fmaxf(xo, load_as_float<DT>(base, (doc_row0 + p) * C + 2));
        if (!(x0 >= xo)) {
          if (C == 2)
            classify<DT, 2>(base, doc_row0 + p, pred, score);
          else
            classify<DT, 3>(base, doc_row0 + p, pred, score);
        }
      }
    }
    const bool hit = pred != 0;
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int pos = cnt + __popc(mask & ((1u << lane) - 1));
      buf_p[warp][pos] = p, buf_s[warp][pos] = score, buf_t[warp][pos] = static_cast<uint8_t>(pred);
    }
    cnt += __popc(mask);
  }
  // Decoupled look-back (single-pass scan): publish this chunk's own count first, then walk back over the
  // predecessors 32 at a time, adding their counts until one that already knows its inclusive prefix.
  // status word: (value << 2) | flag, flag 1 = chunk count, 2 = inclusive prefix, 0 = nothing yet.
  int32_t* st = a.status + (int64_t)list * a.chunks;
  if (lane == 0 && chunk + 1 < a.chunks) atomicExch(&st[chunk], (cnt << 2) | (chunk == 0 ? 2 : 1));
  int prefix = 0, look = chunk - 1;
  while (look >= 0) {
    // kLookWindows windows of 32 predecessors per round, their loads in flight together: lane 0
    // of window 0 = nearest predecessor; before the first chunk: prefix 0
    int v[kLookWindows];
#pragma unroll
    for (int w = 0; w < kLookWindows; ++w) {
      const int idx = look - 32 * w - lane;
      v[w] = idx >= 0 ? *reinterpret_cast<volatile int32_t*>(&st[idx]) : 2;
    }
    int add = 0;
    bool done = false, retry = false;
#pragma unroll
    for (int w = 0; w < kLookWindows; ++w) {
      if (done || retry) continue;  // (warp-uniform)
      const unsigned ready = __ballot_sync(0xffffffffu, v[w] != 0);
      const unsigned is_p = __ballot_sync(0xffffffffu, (v[w] & 3) == 2);
      const int first_p = is_p ? __ffs(is_p) - 1 : 32;
      const unsigned need = first_p >= 31 ? 0xffffffffu : ((1u << (first_p + 1)) - 1u);
      if ((ready & need) != need) {  // a predecessor inside the window has not published yet: read the round again
        retry = true;
        continue;
      }
      add += lane <= first_p ? (v[w] >> 2) : 0;
      done = first_p < 32;
    }
    if (retry) continue;  // (nothing of this round is kept: `add` is dropped)
    for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
    prefix += add;
    if (done) break;
    look -= 32 * kLookWindows;
  }
  if (lane == 0) {
    if (chunk > 0 && chunk + 1 < a.chunks) atomicExch(&st[chunk], ((prefix + cnt) << 2) | 2);
    if (chunk == a.chunks - 1) a.counts[list] = prefix + cnt;
  }
  __syncwarp();  // (the spot buffers of this warp were written by all its lanes)
  const int64_t out0 = (int64_t)list * a.cap;
  for (int r = lane; r < cnt; r += 32) {
    const int dst = prefix + r;
    if (dst < a.cap) {
      a.spot_p[out0 + dst] = buf_p[warp][r];
      a.spot_tag[out0 + dst] = buf_t[warp][r];
      a.spot_score[out0 + dst] = buf_s[warp][r];
    }
  }
}

size_t decode_spots_workspace_bytes(int batch, int n) {
  const int64_t pairs = pair_count(n);
  const int64_t chunks = (pairs + kSpotWarpPairs - 1) / kSpotWarpPairs;
  return static_cast<size_t>((int64_t)batch * kNumHeads * chunks * sizeof(int32_t));
}

int launch_decode_spots(int batch, int n, const void* const in[kNumHeads], int in_dtype, int cap, int32_t* spot_p,
                        int32_t* spot_tag, float* spot_score, int32_t* counts, void* ws, cudaStream_t st) {
  PENEO_REQUIRE(batch >= 0 && n >= 1 && cap >= 1, "decode_spots: bad sizes batch=%d n=%d cap=%d", batch, n, cap);
  PENEO_REQUIRE(n <= 46340, "decode_spots: n too large");
  if (batch == 0) return PENEO_OK;
  SpotArgs a{};
  for (int h = 0; h < kNumHeads; ++h) a.in[h] = in[h];
  a.batch = batch, a.n = n, a.pairs = static_cast<int32_t>(pair_count(n)), a.cap = cap;
  a.chunks = (a.pairs + kSpotWarpPairs - 1) / kSpotWarpPairs;  // warp chunks per list
  a.spot_p = spot_p, a.spot_tag = spot_tag, a.spot_score = spot_score, a.counts = counts;
  a.status = static_cast<int32_t*>(ws);
  PENEO_CUDA_TRY(cudaMemsetAsync(ws, 0, decode_spots_workspace_bytes(batch, n), st));
  PENEO_REQUIRE(batch <= 65535, "decode_spots: batch too large for one launch");
  dim3 grid((a.chunks + kSpotWarps - 1) / kSpotWarps, kNumHeads, batch);
  switch (in_dtype) {
    case PENEO_DT_F32: decode_spots_kernel<PENEO_DT_F32><<<grid, 32 * kSpotWarps, 0, st>>>(a); break;
    case PENEO_DT_BF16: decode_spots_kernel<PENEO_DT_BF16><<<grid, 32 * kSpotWarps, 0, st>>>(a); break;
    case PENEO_DT_F16: decode_spots_kernel<PENEO_DT_F16><<<grid, 32 * kSpotWarps, 0, st>>>(a); break;
    case PENEO_DT_I64: decode_spots_kernel<PENEO_DT_I64><<<grid, 32 * kSpotWarps, 0, st>>>(a); break;
    default: set_error("decode_spots: unsupported dtype %d", in_dtype); return PENEO_E_INVALID;
  }
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
// K3' : gather of the per-tile spot slots written by K2's spots-only epilogue (pair_heads_tc2.cu, classify.cuh).
// Same output as decode_spots_kernel — per (document, head) the spots in increasing p — without the logits ever
// existing in HBM.  One CTA per (chunk of 2048 tiles, head, document); each thread owns 8 consecutive tiles (32
// slots); CTA-wide exclusive scan of the slot counts, chunks chained by the same decoupled look-back as K3.
// ------------------------------------------------------------------------------------------------
constexpr int kGatherTilesPerThread = 8;
constexpr int kGatherTilesPerCta = 256 * kGatherTilesPerThread;

struct GatherArgs {
  TileSpots ts;
  int32_t batch, pairs, cap, chunks;
  int64_t total_tiles;
  int32_t* spot_p;
  int32_t* spot_tag;
  float* spot_score;
  int32_t* counts;
  int32_t* ticket;  // [batch*5]
  int32_t* status;  // [batch*5*chunks]
};

__global__ void __launch_bounds__(256) gather_tile_spots_kernel(const GatherArgs a) {
  __shared__ int32_t wsum[8];
  __shared__ int32_t s_chunk, s_base;
  const int h = blockIdx.y, b = blockIdx.z, list = b * kNumHeads + h;
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  if (threadIdx.x == 0) s_chunk = atomicAdd(&a.ticket[list], 1);
  __syncthreads();
  const int chunk = s_chunk;
  const int64_t gp0 = (int64_t)b * a.pairs, gp1 = gp0 + a.pairs;  // this document's range of the batch-flat pair list
  const int64_t doc_tile0 = gp0 / 128;
  const int64_t t0 = doc_tile0 + (int64_t)chunk * kGatherTilesPerCta + (int64_t)threadIdx.x * kGatherTilesPerThread;
  // counts of this thread's 32 slots (only entries that belong to this document)
  int cnt[kGatherTilesPerThread][4];
  int mine = 0;
#pragma unroll
  for (int t = 0; t < kGatherTilesPerThread; ++t) {
    const int64_t tile = t0 + t;
    const bool live = tile < a.total_tiles && tile * 128 < gp1;
    int4 c4 = make_int4(0, 0, 0, 0);
    if (live) c4 = *reinterpret_cast<const int4*>(a.ts.cnt + (tile * kNumHeads + h) * 4);
    int c[4] = {c4.x, c4.y, c4.z, c4.w};
    if (live && (tile * 128 < gp0 || tile * 128 + 128 > gp1)) {
      // tile straddles a document boundary: count the entries one by one
      for (int q = 0; q < 4; ++q) {
        const int64_t slot = (tile * kNumHeads + h) * 4 + q;
        int keep = 0;
        for (int e = 0; e < c[q]; ++e) {
          const int64_t gp = tile * 128 + (a.ts.meta[slot * 32 + e] & 0xFF);
          keep += gp >= gp0 && gp < gp1;
        }
        c[q] = keep;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) cnt[t][q] = c[q], mine += c[q];
  }
  // CTA-wide exclusive scan of `mine`
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int woff = 0, total = 0;
  for (int w = 0; w < 8; ++w) {
    if (w < warp) woff += wsum[w];
    total += wsum[w];
  }
  if (warp == 0) {
    // decoupled look-back over the preceding chunks of this list (status word: (value << 2) | flag, 1 = own count,
    // 2 = inclusive prefix) — see decode_spots_kernel
    int32_t* st = a.status + (int64_t)list * a.chunks;
    if (lane == 0 && chunk + 1 < a.chunks) {
      __threadfence();
      atomicExch(&st[chunk], (total << 2) | 1);
    }
    int prefix = 0, look = chunk - 1;
    while (look >= 0) {
      const int idx = look - lane;
      const int v = idx >= 0 ? *reinterpret_cast<volatile int32_t*>(&st[idx]) : 2;
      const unsigned ready = __ballot_sync(0xffffffffu, v != 0);
      const unsigned is_p = __ballot_sync(0xffffffffu, (v & 3) == 2);
      const int first_p = is_p ? __ffs(is_p) - 1 : 32;
      const unsigned need = first_p >= 31 ? 0xffffffffu : ((1u << (first_p + 1)) - 1u);
      if ((ready & need) != need) continue;
      int contrib = lane <= first_p ? (v >> 2) : 0;
      for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
      prefix += contrib;
      if (first_p < 32) break;
      look -= 32;
    }
    if (lane == 0) {
      __threadfence();
      atomicExch(&st[chunk], ((prefix + total) << 2) | 2);
      if (chunk == a.chunks - 1) a.counts[list] = prefix + total;
      s_base = prefix;
    }
  }
  __syncthreads();
  int dst = s_base + woff + (incl - mine);
  const int64_t out0 = (int64_t)list * a.cap;
#pragma unroll
  for (int t = 0; t < kGatherTilesPerThread; ++t) {
    const int64_t tile = t0 + t;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (cnt[t][q] == 0) continue;
      const int64_t slot = (tile * kNumHeads + h) * 4 + q;
      const int raw = a.ts.cnt[slot];
      for (int e = 0; e < raw; ++e) {
        const int m = a.ts.meta[slot * 32 + e];
        const int64_t gp = tile * 128 + (m & 0xFF);
        if (gp < gp0 || gp >= gp1) continue;
        if (dst < a.cap) {
          a.spot_p[out0 + dst] = static_cast<int32_t>(gp - gp0);
          a.spot_tag[out0 + dst] = m >> 8;
          a.spot_score[out0 + dst] = a.ts.score[slot * 32 + e];
        }
        ++dst;
      }
    }
  }
}

static int64_t heads_spots_tiles(int batch, int n) { return ((int64_t)batch * pair_count(n) + 127) / 128; }
static int heads_spots_chunks(int n) {  // chunks of kGatherTilesPerCta tiles that can overlap one document
  return static_cast<int>((pair_count(n) + 127) / 128 / kGatherTilesPerCta + 2);
}
size_t heads_spots_workspace_bytes(int batch, int n) {
  const int64_t tiles = heads_spots_tiles(batch, n) + 1;
  return align_up(tile_spots_bytes(tiles), 256) + (size_t)batch * kNumHeads * (1 + heads_spots_chunks(n)) * sizeof(int32_t);
}

int launch_gather_tile_spots(int batch, int n, void* ws, int cap, int32_t* spot_p, int32_t* spot_tag, float* spot_score,
                             int32_t* counts, cudaStream_t st) {
  PENEO_REQUIRE(batch >= 1 && batch <= 65535 && n >= 1 && n <= 46340 && cap >= 1, "gather_tile_spots: bad sizes");
  GatherArgs a{};
  const int64_t tiles = heads_spots_tiles(batch, n);
  a.ts = tile_spots_carve(ws, tiles + 1);
  a.batch = batch, a.pairs = static_cast<int32_t>(pair_count(n)), a.cap = cap, a.chunks = heads_spots_chunks(n);
  a.total_tiles = tiles;
  a.spot_p = spot_p, a.spot_tag = spot_tag, a.spot_score = spot_score, a.counts = counts;
  char* tail = static_cast<char*>(ws) + align_up(tile_spots_bytes(tiles + 1), 256);
  a.ticket = reinterpret_cast<int32_t*>(tail);
  a.status = a.ticket + (int64_t)batch * kNumHeads;
  PENEO_CUDA_TRY(cudaMemsetAsync(tail, 0, (size_t)batch * kNumHeads * (1 + a.chunks) * sizeof(int32_t), st));
  gather_tile_spots_kernel<<<dim3(a.chunks, kNumHeads, batch), 256, 0, st>>>(a);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
// K4 : K4a one CTA per (document, 1-1 map), K4b one CTA per document
// ------------------------------------------------------------------------------------------------
struct ResolveArgs {
  int32_t batch, n, cap, npow2, decode_gt;
  float thresh;
  const int32_t* spot_p;
  const int32_t* spot_tag;
  const float* spot_score;
  const int32_t* counts;
  int32_t* out;
  int64_t doc_ints;
  uint32_t* bitmap;  // [batch][ceil(n*n/32)]
  int64_t bitmap_words;
};

struct SpotList {
  const int32_t* p;
  const int32_t* tag;
  const float* score;
  int32_t cnt;
};

// filtered, swapped (head, tail) of spot s; returns false when the spot is dropped
__device__ __forceinline__ bool spot_edge(const SpotList& L, int s, int n, bool triu, float thresh, int& hd, int& tl,
                                          float& sc) {
  const int tag = L.tag[s];
  sc = L.score[s];
  if (tag == 0 || sc < thresh) return false;
  int i, j;
  pair_from_flat(L.p[s], n, i, j);
  if (triu && tag == 2) {
    hd = j, tl = i;
  } else {
    hd = i, tl = j;
  }
  return true;
}

__device__ void bitonic_sort_u64(unsigned long long* key, int npow2) {
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < npow2; t += blockDim.x) {
        const int ixj = t ^ j;
        if (ixj > t) {
          const unsigned long long a = key[t], b = key[ixj];
          const bool up = (t & k) == 0;
          if ((a > b) == up) key[t] = b, key[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
}

// Builds one 1-1 map.  out_pairs receives (head, tail) in the reference's dict order, map[head] = tail.
__device__ int resolve_map(const SpotList& L, int n, int npow2, bool triu, bool top, float thresh,
                           unsigned long long* best1, unsigned long long* best2, uint32_t* first, uint32_t* ord2,
                           unsigned long long* sortbuf, int32_t* map, int32_t* out_pairs, int32_t* s_count) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int t = tid; t < n; t += nt) best1[t] = 0ull, best2[t] = 0ull, first[t] = 0xFFFFFFFFu, ord2[t] = 0xFFFFFFFFu, map[t] = -1;
  for (int t = tid; t < npow2; t += nt) sortbuf[t] = ~0ull;
  if (tid == 0) *s_count = 0;
  __syncthreads();
  for (int s = tid; s < L.cnt; s += nt) {
    int hd, tl;
    float sc;
    if (!spot_edge(L, s, n, triu, thresh, hd, tl, sc)) continue;
    const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(sc)) << 32) | (0xFFFFFFFFu - s);
    atomicMax(&best1[hd], key);
    atomicMin(&first[hd], static_cast<uint32_t>(s));
  }
  __syncthreads();
  if (top) {
    // stage 2: per tail keep the best head; heads are visited in order of first arrival
    for (int hd = tid; hd < n; hd += nt) {
      if (best1[hd] == 0ull) continue;
      const int s = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(best1[hd]));
      int h2, tl;
      float sc;
      spot_edge(L, s, n, triu, thresh, h2, tl, sc);
      const unsigned long long key = (best1[hd] & 0xFFFFFFFF00000000ull) | (0xFFFFFFFFu - first[hd]);
      atomicMax(&best2[tl], key);
      atomicMin(&ord2[tl], first[hd]);
    }
    __syncthreads();
    for (int tl = tid; tl < n; tl += nt) {
      if (best2[tl] == 0ull) continue;
      const int slot = atomicAdd(s_count, 1);
      sortbuf[slot] = (static_cast<unsigned long long>(ord2[tl]) << 32) | static_cast<uint32_t>(tl);
    }
  } else {
    // GT mode: {head: [tails...]} then first element -> tail of the first-arriving spot per head
    for (int hd = tid; hd < n; hd += nt) {
      if (first[hd] == 0xFFFFFFFFu) continue;
      const int slot = atomicAdd(s_count, 1);
      sortbuf[slot] = (static_cast<unsigned long long>(first[hd]) << 32) | static_cast<uint32_t>(hd);
    }
  }
  __syncthreads();
  const int cnt = *s_count;
  bitonic_sort_u64(sortbuf, npow2);
  for (int r = tid; r < cnt; r += nt) {
    const unsigned long long e = sortbuf[r];
    int hd, tl;
    float sc;
    if (top) {
      tl = static_cast<int>(static_cast<uint32_t>(e));
      const int s = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(best2[tl]));  // first[] of the winning head
      int t2;
      spot_edge(L, s, n, triu, thresh, hd, t2, sc);
    } else {
      hd = static_cast<int>(static_cast<uint32_t>(e));
      int h2;
      spot_edge(L, static_cast<int>(e >> 32), n, triu, thresh, h2, tl, sc);
    }
    out_pairs[2 * r] = hd, out_pairs[2 * r + 1] = tl;
    map[hd] = tl;
  }
  __syncthreads();
  return cnt;
}

// ordered compaction of the filtered (head, tail) edges of a spot list; optionally sets the bitmap
__device__ int emit_edges(const SpotList& L, int n, float thresh, int32_t* out_pairs, uint32_t* bitmap, int32_t* s_warp,
                          int32_t* s_run) {
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32, nw = blockDim.x / 32;
  if (tid == 0) *s_run = 0;
  __syncthreads();
  for (int s0 = 0; s0 < L.cnt; s0 += blockDim.x) {
    const int s = s0 + tid;
    int hd = 0, tl = 0;
    float sc;
    const bool keep = s < L.cnt && spot_edge(L, s, n, true, thresh, hd, tl, sc);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    int off = *s_run;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (keep) {
      const int pos = off + __popc(mask & ((1u << lane) - 1));
      out_pairs[2 * pos] = hd, out_pairs[2 * pos + 1] = tl;
      if (bitmap) {
        const int64_t bit = (int64_t)hd * n + tl;
        atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
      }
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < nw; ++w) tot += s_warp[w];
      *s_run += tot;
    }
    __syncthreads();
  }
  return *s_run;
}

__device__ __forceinline__ void chain_walk(int hd, int tl, const int32_t* le, const int32_t* lgh, const int32_t* lgt,
                                           int& nseg, int& last_tail) {
  // pipeline/decode.py:257-296: follow line-grouping heads while LE and LG-t2t agree; 1000-step cap
  nseg = 1;
  int cur_h = hd, cur_t = tl, nxt = lgh[cur_h], ops = 0;
  while (nxt >= 0) {
    if (++ops > 1000) break;
    if (nxt == cur_h) break;
    const int nt = le[nxt];
    if (nt < 0) break;
    if (lgt[cur_t] != nt) break;
    ++nseg;
    cur_h = nxt, cur_t = nt;
    nxt = lgh[cur_h];
  }
  last_tail = cur_t;
}

// K4a: the three 1-1 maps of a document are independent -> one CTA per (document, map).
//   y = 0: line extraction (head 0) ; y = 1: line grouping tail-to-tail (head 4) ; y = 2: line grouping head-to-head (head 3)
__global__ void __launch_bounds__(256) decode_maps_kernel(const ResolveArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, np2 = a.npow2, b = blockIdx.x, which = blockIdx.y;
  unsigned long long* best1 = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned long long* best2 = best1 + n;
  unsigned long long* sortbuf = best2 + n;
  uint32_t* first = reinterpret_cast<uint32_t*>(sortbuf + np2);
  uint32_t* ord2 = first + n;
  int32_t* map = reinterpret_cast<int32_t*>(ord2 + n);
  __shared__ int32_t s_count;

  int32_t* out = a.out + (int64_t)b * a.doc_ints;
  const int h = which == 0 ? 0 : (which == 1 ? 4 : 3);
  int32_t* o_pairs = out + 16 + (which == 0 ? 0 : (which == 1 ? 4 * n : 2 * n));  // o_le | o_lgh | o_lgt order in the record
  SpotList L;
  const int64_t off = ((int64_t)b * kNumHeads + h) * a.cap;
  L.p = a.spot_p + off, L.tag = a.spot_tag + off, L.score = a.spot_score + off;
  L.cnt = min(a.counts[b * kNumHeads + h], a.cap);
  const int cnt = resolve_map(L, n, np2, which != 0, a.decode_gt == 0, a.thresh, best1, best2, first, ord2, sortbuf, map,
                              o_pairs, &s_count);
  if (threadIdx.x == 0) out[which == 0 ? 0 : (which == 1 ? 2 : 1)] = cnt;  // n_le | n_lgt | n_lgh
}

// K4b: one CTA per document: entity-linking edge lists and the key/value chain walk over the maps of K4a
__global__ void __launch_bounds__(256) decode_links_kernel(const ResolveArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, b = blockIdx.x;
  int32_t* le = reinterpret_cast<int32_t*>(smem_raw);
  int32_t* lgh = le + n;
  int32_t* lgt = lgh + n;
  __shared__ int32_t s_run, s_warp[8];

  int32_t* out = a.out + (int64_t)b * a.doc_ints;
  int32_t* o_le = out + 16;
  int32_t* o_lgh = o_le + 2 * n;
  int32_t* o_lgt = o_lgh + 2 * n;
  int32_t* o_elh = o_lgt + 2 * n;
  int32_t* o_elt = o_elh + 2 * a.cap;
  int32_t* o_kv = o_elt + 2 * a.cap;

  auto list = [&](int h) {
    SpotList L;
    const int64_t off = ((int64_t)b * kNumHeads + h) * a.cap;
    L.p = a.spot_p + off, L.tag = a.spot_tag + off, L.score = a.spot_score + off;
    L.cnt = min(a.counts[b * kNumHeads + h], a.cap);
    return L;
  };
  // dense maps back from the (head, tail) lists K4a wrote (heads are unique within a map)
  const int n_le = out[0], n_lgh = out[1], n_lgt = out[2];
  for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) le[t] = -1;
  __syncthreads();
  for (int r = threadIdx.x; r < n_le; r += blockDim.x) le[o_le[2 * r]] = o_le[2 * r + 1];
  for (int r = threadIdx.x; r < n_lgh; r += blockDim.x) lgh[o_lgh[2 * r]] = o_lgh[2 * r + 1];
  for (int r = threadIdx.x; r < n_lgt; r += blockDim.x) lgt[o_lgt[2 * r]] = o_lgt[2 * r + 1];
  __syncthreads();

  uint32_t* bitmap = a.bitmap + (int64_t)b * a.bitmap_words;
  const int n_elt = emit_edges(list(2), n, a.thresh, o_elt, bitmap, s_warp, &s_run);
  __syncthreads();
  const int n_elh = emit_edges(list(1), n, a.thresh, o_elh, nullptr, s_warp, &s_run);
  __threadfence_block();
  __syncthreads();

  // key/value pairs: one thread per entity-linking h2h edge, ordered compaction of the valid ones
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (int e0 = 0; e0 < n_elh; e0 += blockDim.x) {
    const int e = e0 + tid;
    bool ok = false;
    int kh = 0, vh = 0, nk = 0, nv = 0;
    if (e < n_elh) {
      kh = o_elh[2 * e], vh = o_elh[2 * e + 1];
      const int kt = le[kh], vt = le[vh];
      if (kt >= 0 && vt >= 0) {
        int klast, vlast;
        chain_walk(kh, kt, le, lgh, lgt, nk, klast);
        chain_walk(vh, vt, le, lgh, lgt, nv, vlast);
        const int64_t bit = (int64_t)klast * n + vlast;
        ok = (bitmap[bit >> 5] >> (bit & 31)) & 1u;
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    int off = s_run;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (ok) {
      const int pos = off + __popc(mask & ((1u << lane) - 1));
      o_kv[4 * pos] = kh, o_kv[4 * pos + 1] = vh, o_kv[4 * pos + 2] = nk, o_kv[4 * pos + 3] = nv;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp[w];
      s_run += tot;
    }
    __syncthreads();
  }
  if (tid == 0) {
    out[3] = n_elh, out[4] = n_elt, out[5] = s_run;
    for (int q = 6; q < 16; ++q) out[q] = 0;
  }
}

size_t decode_resolve_doc_ints(int n, int cap) { return 16 + 6 * (size_t)n + 8 * (size_t)cap; }
size_t decode_resolve_workspace_bytes(int batch, int n) {
  const int64_t words = ((int64_t)n * n + 31) / 32;
  return static_cast<size_t>(batch * words * 4);
}

int launch_decode_resolve(int batch, int n, int cap, const int32_t* spot_p, const int32_t* spot_tag,
                          const float* spot_score, const int32_t* counts, int decode_gt, float score_thresh,
                          int32_t* out, void* ws, cudaStream_t st) {
  PENEO_REQUIRE(n >= 1 && n <= 4096, "decode_resolve: n=%d outside [1, 4096]", n);
  PENEO_REQUIRE(cap >= 1, "decode_resolve: cap must be positive");
  if (batch == 0) return PENEO_OK;
  ResolveArgs a{};
  a.batch = batch, a.n = n, a.cap = cap, a.decode_gt = decode_gt, a.thresh = score_thresh;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  a.npow2 = np2;
  a.spot_p = spot_p, a.spot_tag = spot_tag, a.spot_score = spot_score, a.counts = counts, a.out = out;
  a.doc_ints = static_cast<int64_t>(decode_resolve_doc_ints(n, cap));
  a.bitmap = static_cast<uint32_t*>(ws);
  a.bitmap_words = ((int64_t)n * n + 31) / 32;
  PENEO_CUDA_TRY(cudaMemsetAsync(ws, 0, decode_resolve_workspace_bytes(batch, n), st));
  const size_t smem_a = (size_t)(2 * n + np2) * 8 + (size_t)3 * n * 4, smem_b = (size_t)3 * n * 4;
  PENEO_CUDA_TRY(cudaFuncSetAttribute(decode_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_a)));
  PENEO_CUDA_TRY(cudaFuncSetAttribute(decode_links_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_b)));
  decode_maps_kernel<<<dim3(batch, 3), 256, smem_a, st>>>(a);
  PENEO_CUDA_TRY(cudaGetLastError());
  decode_links_kernel<<<batch, 256, smem_b, st>>>(a);
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
