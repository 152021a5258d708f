// K2 on CTA pairs (cta_group::2) — same algorithm, warp roles and TMEM layout as pair_heads_tc.cu, but two CTAs of
// a cluster (one TPC) run every UMMA together with M = 256: each CTA owns one 128-pair tile (its S rows, accumulators
// and epilogue) and supplies HALF of the B operand (64 of the 128 W_mid rows of a chunk, 8 of the 16 W_out rows)
// from its shared memory.  Per SM this halves both the TMA writes of the W stream and the tensor core's B reads —
// the two streams that saturate the 128 B/clk shared-memory port in the single-CTA kernel.
//
// Cross-CTA synchronisation (the protocol of CUTLASS / DeepGEMM 2-SM kernels):
//   * only the leader (cluster rank 0) issues tcgen05.mma; it waits on ITS barriers for "W stage full" (both TMA
//     warps arrive, transaction bytes of both halves are counted there), "S chunk full" (producer warps of both
//     CTAs) and "m ready" (epilogue warps of both CTAs) — the peer arrives remotely (mapa + mbarrier.arrive);
//   * everything the MMA releases ("W stage empty", "S chunk free", "u full", "z full", "W_out stage empty") is a
//     tcgen05.commit multicast to the same barrier in both CTAs.
//
// b_mid rides on the tensor cores: the A operand has a constant 25th K step [1, 0, ..., 0] (eight TMEM columns written
// once) and the B operand of that step is a resident shared-memory tile whose K column 0 holds b_mid / 2 (bf16), so
// u / 2 + b_mid / 2 arrives complete in the accumulator — one more UMMA per chunk (+4 % tensor work) instead of a
// broadcast LDS.128 per four elements and an FADD per element in the epilogue warps that pace the kernel.
#include <cuda.h>

#include <cstdlib>

#include "classify.cuh"
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {
namespace k2p {

constexpr int D = 384;
// Width of a chunk of hidden features = N of the first GEMM's UMMAs.  128 (default): two accumulator buffers, the MMA
// warp needs m(g - 1) before it may issue u(g + 1).  64: FOUR buffers, the second GEMM lags three chunks behind, so the
// hand-off latencies between tensor pipe and epilogue are hidden — measured on B200: 5.73 ms against 4.44 ms per
// launch, i.e. the TS-form UMMA (A operand read from TMEM: 4 KB per instruction whatever N is) does not run N = 64 at
// full rate, which costs more than the deeper pipeline gains.  Kept as a build option for the record.
#ifndef PENEO_K2_CHUNK_COLS
#define PENEO_K2_CHUNK_COLS 128
#endif
constexpr int kChunkCols = PENEO_K2_CHUNK_COLS;
static_assert(kChunkCols == 128 || kChunkCols == 64, "chunk width");
constexpr int kHeadChunks = D / kChunkCols;        // chunks per head: 3 / 6
constexpr int kChunks = kNumHeads * kHeadChunks;   // 15 / 30
constexpr int kUBufs = 256 / kChunkCols;           // accumulator buffers in the 256 TMEM columns [192, 448): 2 / 4
constexpr int kLag = kUBufs - 1;                   // the second GEMM of chunk g - kLag is issued after the first GEMM of chunk g
constexpr int kKChunks = 6;       // 384 / 64
constexpr int kWStageBytes = (kChunkCols / 2) * 64 * 2;  // 8 / 4 KB: this CTA's half of the chunk's W_mid rows x 64 K columns
#ifndef PENEO_K2_WRING_KB
#define PENEO_K2_WRING_KB 48
#endif
constexpr int kWStages = PENEO_K2_WRING_KB * 1024 / kWStageBytes;  // 48 KB ring: 6 / 12 stages
constexpr int kOStages = 2 * kUBufs;  // (a W_out stage is released kLag chunks after it was filled: the ring must be deeper than the lag)
constexpr int kOKBlocks = kChunkCols / 64;               // 64-wide K blocks of the second GEMM per chunk
constexpr int kOStageBytes = kOKBlocks * 8 * 64 * 2;     // 2 / 1 KB: K blocks of [8 rows x 64] (this CTA's half of 16)
constexpr int kStageRowBytes = D * 2;          // staging: 128 rows x 768 B ...
constexpr int kStageBoxBytes = 128 * 128;      // ... as six [128 rows x 64 features] boxes in the TMA SWIZZLE_128B layout
constexpr int kEpiWarps = 16;      // 4 per scheduler: each takes 32 columns of a chunk (64-column chunks: of every other chunk)
constexpr int kColGroups = kChunkCols / 32;           // 32-column slices per chunk: 4 / 2
constexpr int kEpiSets = kEpiWarps / (4 * kColGroups);  // sets of warps taking chunks in turn: 1 / 2
constexpr int kProdWarp0 = 4 + kEpiWarps;
constexpr int kProdRows = 4;        // b_j rows a producer warp has in flight (register budget: 80 / thread)
constexpr int kThreads = 32 * (kProdWarp0 + 4);

constexpr uint32_t kColS = 0, kColU = 192, kColZ = 448, kColOne = 480;  // [480, 496): the constant "ones" K step of A
constexpr int kBiasRows = kChunkCols / 2;      // this CTA's half of a chunk's features
constexpr int kBiasTileBytes = kBiasRows * 128;  // [rows x 128 B] SWIZZLE_128B, K-major: the 32-byte K step x of a row belongs
                                                 // to chunk 4 * tile + x (K column 0 = b_mid / 2)
constexpr int kBiasTiles = (kChunks + 3) / 4;    // four chunks per tile

#ifdef PENEO_K2_PROFILE
// Diagnostic build (-DPENEO_K2_PROFILE): cycles the MMA-issuing thread spends waiting on each kind of barrier, summed
// over the leader CTAs: [0] S chunk full, [1] W stage full, [2] m ready, [3] z free, [4] W_out stage full, [5] total
// cycles of the issuing loop, [6] leader CTAs.  Read with peneo_debug_k2_profile().
__device__ unsigned long long g_k2_prof[8];
#define K2_WAIT(slot, call)                       \
  do {                                            \
    const long long t0_ = clock64();              \
    call;                                         \
    prof[slot] += clock64() - t0_;                \
  } while (0)
#else
#define K2_WAIT(slot, call) call
#endif

struct Smem {
  static constexpr int w = 0;
  static constexpr int o = w + kWStages * kWStageBytes;
  static constexpr int stage = o + kOStages * kOStageBytes;
  static constexpr int bias = stage + 128 * kStageRowBytes;  // bias B-operand tiles (see kBiasTileBytes)
  static constexpr int hstage = bias + kBiasTiles * kBiasTileBytes;  // SAVE: per epilogue warp one [32 rows x 64 B] TMA store tile (SWIZZLE_64B)
#ifdef PENEO_K2_NO_HSTAGE  // (timing experiment for the inference instantiations only: SAVE would overrun)
  static constexpr int bout = hstage;
#else
  static constexpr int bout = hstage + kEpiWarps * 2048;  // 20 floats
#endif
  static constexpr int loss = bout + 128;                    // 40 doubles (LOSS instantiation)
  static constexpr int bars = loss + 320;
  static constexpr int total = bars + 1024;
};
// barrier indices
constexpr int bWFull = 0, bWEmpty = bWFull + kWStages, bOFull = bWEmpty + kWStages, bOEmpty = bOFull + kOStages,
              bUFull = bOEmpty + kOStages, bMReady = bUFull + kUBufs, bZFull = bMReady + kUBufs, bZFree = bZFull + 2, bSFull = bZFree + 2,
              bSFree = bSFull + kKChunks, bCount = bSFree + kKChunks;
static_assert(bCount * 8 + 16 <= 1024, "barrier area too small");
static_assert(Smem::stage % 1024 == 0 && Smem::bias % 1024 == 0 && Smem::hstage % 1024 == 0,
              "UMMA operand / TMA store tiles need 1024-byte alignment");
constexpr int kSmemBytes = Smem::total + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct Args {
  const __nv_bfloat16* ab;  // [batch*n, 768] : 0.5*A | 0.5*Bm
  const float* bmid_half;   // [1920]
  const float* bout;        // [5][4]
  float* logits[kNumHeads];
  int32_t n, pairs_per_doc;
  int64_t total_pairs;
  int32_t num_tiles;
  uint32_t drop_thresh;  // training-mode dropout after the hidden SiLU (DROP instantiation only)
  float drop_scale;
  uint32_t drop_key[kNumHeads];
  // LOSS instantiation: the class-weighted cross entropy of model/custom_loss.py:189-202 is reduced in the tiles
  const int64_t* tags[kNumHeads];  // int64 [batch*P]
  float class_w[3];
  double* loss_partial;            // [5][gridDim.x][2] : per-CTA (sum w nll, sum w), reduced by pair_loss_final_kernel
  // SAVE instantiation (training, with LOSS): the pre-activations (u + b_mid) / 2 (bf16 [pairs, 1920]) and s (bf16
  // [pairs, 384]) are stored as they pass through the registers of the epilogue / producer warps, so that the backward
  // pass needs no recompute GEMM (pair_bwd_elem.cu)
  __nv_bfloat16* save_h;
  __nv_bfloat16* save_s;
  // SPOTS instantiation: no logits leave the SM; the epilogue classifies each pair (model/peneo_decoder.py:98-114) and
  // writes the few non-zero predictions into per-tile slots (classify.cuh), gathered in order by decode.cu
  TileSpots spots;
};

template <bool DROP, bool LOSS, bool SPOTS, bool SAVE>
__global__ void __launch_bounds__(kThreads, 1)
    pair_heads_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
                         const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmS, const Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // by offset: keeps the shared address space
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  float* s_bout = reinterpret_cast<float*>(smem + Smem::bout);
  double* s_loss = reinterpret_cast<double*>(smem + Smem::loss);  // [4 quadrants][5 heads][2]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader (issues every MMA of the pair)
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmO);
    if (SAVE) ptx::prefetch_tmap(&tmH), ptx::prefetch_tmap(&tmS);
    for (int s = 0; s < kWStages; ++s) ptx::mbar_init(&bars[bWFull + s], 2), ptx::mbar_init(&bars[bWEmpty + s], 1);
    for (int s = 0; s < kOStages; ++s) ptx::mbar_init(&bars[bOFull + s], 2), ptx::mbar_init(&bars[bOEmpty + s], 1);
    for (int s = 0; s < kUBufs; ++s) {
      ptx::mbar_init(&bars[bUFull + s], 1);
      ptx::mbar_init(&bars[bMReady + s], 2 * kEpiWarps / kEpiSets);  // the epilogue warps (of this chunk's set) of both CTAs
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bars[bZFull + s], 1);
      ptx::mbar_init(&bars[bZFree + s], 8);  // the four emitting (producer) warps of each CTA
    }
    for (int s = 0; s < kKChunks; ++s) ptx::mbar_init(&bars[bSFull + s], 8), ptx::mbar_init(&bars[bSFree + s], 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 20) s_bout[threadIdx.x] = a.bout[threadIdx.x];
  if (LOSS && threadIdx.x < 40) s_loss[threadIdx.x] = 0.0;
  // bias operand tiles: zero, then K column 0 of (chunk c, row r) = b_mid / 2 of feature c * 128 + 64 * rank + r.
  // Row r of a tile is 128 B, its 16-byte chunk index is XORed with (r & 7) (the SWIZZLE_128B pattern the MMA expects).
  for (int e = threadIdx.x; e < kBiasTiles * kBiasTileBytes / 16; e += kThreads)
    reinterpret_cast<uint4*>(smem + Smem::bias)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int e = threadIdx.x; e < kChunks * kBiasRows; e += kThreads) {
    const int c = e / kBiasRows, r = e % kBiasRows;
    const float bh = a.bmid_half[c * kChunkCols + static_cast<int>(rank) * kBiasRows + r];
    *reinterpret_cast<__nv_bfloat16*>(smem + Smem::bias + (c >> 2) * kBiasTileBytes + r * 128 + (((2 * (c & 3)) ^ (r & 7)) * 16)) =
        __float2bfloat16_rn(bh);
  }
  ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
  ptx::cluster_sync_all();  // both CTAs' barriers exist before anyone (TMA of the peer, remote arrives) touches them
  if (warp == 2) {
    ptx::tmem_alloc_2sm(tmem_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // arrive on the LEADER's copy of a barrier (local for the leader, remote for the peer)
  auto arrive_leader = [&](uint64_t* bar) {
    if (leader) ptx::mbar_arrive(bar);
    else ptx::mbar_arrive_remote(bar, 0);
  };

  // Tiles are handed out in pairs: cluster `cid` takes tile pairs cid, cid + nclusters, ...; this CTA owns tile
  // 2 * pair + rank (the last pair may have a phantom second tile: all its rows are >= total_pairs).
  const int cid = static_cast<int>(blockIdx.x >> 1), ncl = static_cast<int>(gridDim.x >> 1);
  const int num_pairs = (a.num_tiles + 1) / 2;
  const int my_tiles = (num_pairs > cid) ? (num_pairs - 1 - cid) / ncl + 1 : 0;
  auto tile_of = [&](int it) { return 2 * (static_cast<int64_t>(cid) + static_cast<int64_t>(it) * ncl) + rank; };

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (ptx::elect_one()) {
      int ws = 0, os = 0;
      uint32_t wph = 0, oph = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int c = 0; c < kChunks; ++c) {
          for (int kc = 0; kc < kKChunks; ++kc) {
            ptx::mbar_wait(&bars[bWEmpty + ws], wph ^ 1);
            // this CTA's half of the chunk's rows; both halves complete on the leader's barrier
            ptx::tma_load_2d_2sm(smem + Smem::w + ws * kWStageBytes, &tmW, &bars[bWFull + ws], kc * 64,
                                 c * kChunkCols + static_cast<int>(rank) * (kChunkCols / 2));
            if (leader) ptx::mbar_arrive_expect_tx(&bars[bWFull + ws], 2 * kWStageBytes);
            else ptx::mbar_arrive_remote(&bars[bWFull + ws], 0);
            if (++ws == kWStages) ws = 0, wph ^= 1;
          }
          ptx::mbar_wait(&bars[bOEmpty + os], oph ^ 1);
          unsigned char* od = smem + Smem::o + os * kOStageBytes;
          // W_out is packed [15 x 16 rows, 128 features]: row block = 128-feature chunk, this CTA's 8 of its 16 rows
          if (kChunkCols == 128) {
            ptx::tma_load_2d_2sm(od, &tmO, &bars[bOFull + os], 0, c * 16 + static_cast<int>(rank) * 8);
            ptx::tma_load_2d_2sm(od + 1024, &tmO, &bars[bOFull + os], 64, c * 16 + static_cast<int>(rank) * 8);
          } else {
            ptx::tma_load_2d_2sm(od, &tmO, &bars[bOFull + os], (c & 1) * 64, (c >> 1) * 16 + static_cast<int>(rank) * 8);
          }
          if (leader) ptx::mbar_arrive_expect_tx(&bars[bOFull + os], 2 * kOStageBytes);
          else ptx::mbar_arrive_remote(&bars[bOFull + os], 0);
          if (++os == kOStages) os = 0, oph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // (elect_one, not lane == 0: ptxas then treats the region as uniform and keeps descriptors in
    //  uniform registers instead of emitting a per-instruction R2UR waterfall loop)
    if (leader && ptx::elect_one()) {
      constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(256, kChunkCols);  // M = 256 across the CTA pair
      constexpr uint32_t idesc2 = ptx::umma_idesc_bf16(256, 16);
      int ws = 0, os = 0;
      uint32_t wph = 0, oph = 0;
      const uint32_t w_base = ptx::smem_u32(smem + Smem::w), o_base = ptx::smem_u32(smem + Smem::o);
      const uint32_t b_base = ptx::smem_u32(smem + Smem::bias);
#ifdef PENEO_K2_PROFILE
      long long prof[6] = {0, 0, 0, 0, 0, 0};
      const long long prof_t0 = clock64();
#endif
      // second GEMM of global chunk gp: z[head] (+)= m(gp) * W_out chunk^T
      auto mma2 = [&](int gp) {
        const int buf = gp % kUBufs, hg = gp / kHeadChunks, cpos = gp - hg * kHeadChunks;
        K2_WAIT(2, ptx::mbar_wait(&bars[bMReady + buf], (gp / kUBufs) & 1));
        if (cpos == 0 && hg >= 2) K2_WAIT(3, ptx::mbar_wait(&bars[bZFree + (hg & 1)], ((hg >> 1) & 1) ^ 1));  // z of head hg - 2 was read
        K2_WAIT(4, ptx::mbar_wait(&bars[bOFull + os], oph));
        ptx::tc_fence_after();
        const uint32_t zt = tmem + kColZ + 16 * (hg & 1);
        const uint32_t mt = tmem + kColU + kChunkCols * buf;
#ifdef PENEO_K2_ABLATE_MMA2  // timing experiment only (wrong logits): how much of the tensor pipe the N = 16 GEMM takes
        constexpr int kSteps2 = 0;
#else
        constexpr int kSteps2 = kChunkCols / 16;
#endif
#pragma unroll
        for (int ks = 0; ks < kSteps2; ++ks) {
          const uint32_t at = mt + 32 * (ks >> 1) + 8 * (ks & 1);  // m of columns [32 j, 32 j + 32) sits in TMEM columns [32 j, 32 j + 16)
          const uint64_t bd = ptx::umma_desc_sw128(o_base + os * kOStageBytes + (ks / 4) * 1024 + (ks % 4) * 32);
          ptx::umma_ts_2sm(zt, at, bd, idesc2, (cpos | ks) != 0);
        }
        ptx::tc_commit_2sm(&bars[bOEmpty + os], 3);
        if (cpos == kHeadChunks - 1) ptx::tc_commit_2sm(&bars[bZFull + (hg & 1)], 3);
        if (++os == kOStages) os = 0, oph ^= 1;
      };
      int g = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int c = 0; c < kChunks; ++c, ++g) {
          const uint32_t ut = tmem + kColU + kChunkCols * (g % kUBufs);
          // u = [1 0 .. 0] x (b_mid / 2 tile): initialises the accumulator (the ones columns were written before the
          // first "S chunk full" arrival, which the first MMA of the kernel has not passed yet — see the producers)
          if (g == 0) {
            ptx::mbar_wait(&bars[bSFull + 0], 0);
            ptx::tc_fence_after();
          }
          ptx::umma_ts_2sm(ut, tmem + kColOne, ptx::umma_desc_sw128(b_base + (c >> 2) * kBiasTileBytes + (c & 3) * 32), idesc1, 0);
          for (int kc = 0; kc < kKChunks; ++kc) {
            if (c == 0) K2_WAIT(0, ptx::mbar_wait(&bars[bSFull + kc], it & 1));
            K2_WAIT(1, ptx::mbar_wait(&bars[bWFull + ws], wph));
            ptx::tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_ts_2sm(ut, tmem + kColS + 32 * kc + 8 * ks,
                               ptx::umma_desc_sw128(w_base + ws * kWStageBytes + ks * 32), idesc1, 1);
            ptx::tc_commit_2sm(&bars[bWEmpty + ws], 3);
            if (c == kChunks - 1) ptx::tc_commit_2sm(&bars[bSFree + kc], 3);
            if (++ws == kWStages) ws = 0, wph ^= 1;
          }
          ptx::tc_commit_2sm(&bars[bUFull + (g % kUBufs)], 3);
          if (g >= kLag) mma2(g - kLag);  // (u(g + 1) overwrites the buffer of chunk g + 1 - kUBufs = g - kLag)
        }
      }
      for (int gp = g > kLag ? g - kLag : 0; gp < g; ++gp) mma2(gp);
#ifdef PENEO_K2_PROFILE
      prof[5] = clock64() - prof_t0;
      for (int i = 0; i < 6; ++i) atomicAdd(&g_k2_prof[i], static_cast<unsigned long long>(prof[i]));
      atomicAdd(&g_k2_prof[6], 1ull);
#endif
    }
  } else if (warp >= 4 && warp < kProdWarp0) {
    // ============================== epilogue ==============================
    // 16 warps: quadrant q = warp % 4 (TMEM lanes 32 q ..), column slice csel = (warp - 4) / 4 of the chunk's 128 columns.
    // Four warps per scheduler hide the TMEM-load / MUFU / barrier latencies of one another.  Nothing but the hidden
    // activation happens here: every chunk waits for its slowest epilogue warp, so the logits / loss / spot work of
    // a finished head is done by the producer warps (below), which are idle two thirds of the time.
    // (64-column chunks: two sets of eight warps take the chunks in turn — kChunks is even, so a warp's chunks have the
    //  parity of its set in every tile)
    const int q = warp % 4, csel = ((warp - 4) / 4) % kColGroups, eset = (warp - 4) / (4 * kColGroups);
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int row = q * 32 + lane;
    static_assert(kChunks % kEpiSets == 0, "chunk interleave");
    for (int it = 0; it < my_tiles; ++it) {
      const int64_t drop_row = tile_of(it) * 128 + row;
      for (int c = eset; c < kChunks; c += kEpiSets) {
        const int g = it * kChunks + c;
        const int buf = g % kUBufs;
        ptx::mbar_wait(&bars[bUFull + buf], (g / kUBufs) & 1);
        ptx::tc_fence_after();
        const uint32_t ut = tmem + lane_base + kColU + kChunkCols * buf + 32 * csel;
        uint32_t r[32];
        ptx::tmem_ld_x32(ut, r);
        ptx::tmem_ld_wait();
        uint32_t packed[16];
        uint32_t hsave[SAVE ? 16 : 1];
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          if (SAVE) {
            hsave[x / 2] = ptx::pack_bf16x2(__uint_as_float(r[x + 0]), __uint_as_float(r[x + 1]));
            hsave[x / 2 + 1] = ptx::pack_bf16x2(__uint_as_float(r[x + 2]), __uint_as_float(r[x + 3]));
          }
          // the accumulator holds (u + b_mid) / 2 (the bias came through the MMA)
          float m0 = ptx::silu_from_half(__uint_as_float(r[x + 0])), m1 = ptx::silu_from_half(__uint_as_float(r[x + 1]));
          float m2 = ptx::silu_from_half(__uint_as_float(r[x + 2])), m3 = ptx::silu_from_half(__uint_as_float(r[x + 3]));
          if (DROP) {  // nn.Dropout after the hidden SiLU (model/peneo_decoder.py:261), regenerable mask
            const uint32_t key = a.drop_key[c / kHeadChunks], grow = static_cast<uint32_t>(drop_row);
            const uint32_t col = (c % kHeadChunks) * kChunkCols + 32 * csel + x;
            m0 = drop_keep(key, a.drop_thresh, grow, col) ? m0 * a.drop_scale : 0.f;
            m1 = drop_keep(key, a.drop_thresh, grow, col + 1) ? m1 * a.drop_scale : 0.f;
            m2 = drop_keep(key, a.drop_thresh, grow, col + 2) ? m2 * a.drop_scale : 0.f;
            m3 = drop_keep(key, a.drop_thresh, grow, col + 3) ? m3 * a.drop_scale : 0.f;
          }
          packed[x / 2] = ptx::pack_bf16x2(m0, m1);
          packed[x / 2 + 1] = ptx::pack_bf16x2(m2, m3);
        }
        // m overwrites the first half of the 32 columns this warp has consumed
        ptx::tmem_st_x16(ut, packed);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(&bars[bMReady + buf]);
#ifndef PENEO_K2_ABLATE_SAVE_H  // (timing experiment only)
        if (SAVE) {
#else
        if (false) {
#endif
          // this warp's [32 pairs x 32 features] of pre-activations leave as ONE TMA store from a swizzled tile (a
          // direct STG.128 per lane touches 32 lines per instruction: measured 2.2 x slower for the whole kernel)
          unsigned char* hst = smem + Smem::hstage + (warp - 4) * 2048;
          if (lane == 0) ptx::bulk_wait_group_read<0>();  // the previous chunk's store has read the tile
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(hst + lane * 64 + ((j ^ ((lane >> 1) & 3)) * 16)) =
                make_uint4(hsave[4 * j], hsave[4 * j + 1], hsave[4 * j + 2], hsave[4 * j + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {  // rows past the end of the pair list are clipped by the tensor map
            // (evict-first: 15.7 GB of write-once data per 32 x seq-512 batch must not push W_mid / ab out of the L2;
            //  measured 6.30 -> 6.01 ms for the launch)
            ptx::tma_store_2d_hint(&tmH, hst, c * kChunkCols + 32 * csel, static_cast<int32_t>(drop_row - lane), ptx::l2_policy_evict_first());
            ptx::bulk_commit_group();
          }
        }
      }
    }
    if (SAVE && lane == 0) ptx::bulk_wait_group<0>();
  } else if (warp >= kProdWarp0) {
    // ============================== pair producers ==============================
    const int q = warp - kProdWarp0;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    unsigned char* stg = smem + Smem::stage;
    const int row = q * 32 + lane;
    // ---- logits of finished heads (head-global index hg: tile = hg / 5, head = hg % 5).  z sits in one of two TMEM
    //      accumulators; emit_ready() drains every head whose MMA2 has completed and hands the accumulator back to the
    //      MMA warp (bZFree).  It is called between the row groups of the s generation and while waiting for s chunks
    //      to be released, so a head waits a few hundred cycles at most — the accumulator is needed again three chunks
    //      (~6000 cycles) later.
    float acc_l[kNumHeads] = {0.f, 0.f, 0.f, 0.f, 0.f}, acc_w[kNumHeads] = {0.f, 0.f, 0.f, 0.f, 0.f};  // LOSS: per-lane sums
    const int total_heads = my_tiles * kNumHeads;
    int next_emit = 0;
    auto emit_z = [&](int hg) {
      ptx::tc_fence_after();
      uint32_t zr[4];
      ptx::tmem_ld_x4(tmem + lane_base + kColZ + 16 * (hg & 1), zr);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(&bars[bZFree + (hg & 1)]);  // the accumulator may be overwritten (head hg + 2)
      const int it = hg / 5, k = hg - it * 5;
      const int64_t tile = tile_of(it);
      const int64_t gp = tile * 128 + row;
      const int C = head_classes(k);
      const float z0 = __uint_as_float(zr[0]) + s_bout[k * 4], z1 = __uint_as_float(zr[1]) + s_bout[k * 4 + 1];
      const float z2 = C == 3 ? __uint_as_float(zr[2]) + s_bout[k * 4 + 2] : -INFINITY;
      if (SPOTS) {
        // spot extraction fused into the tile: same fast reject and exact softmax as decode_spots_kernel, ordered
        // compaction of the warp's 32 pairs by ballot; only the spots (a few per thousand pairs) are written
        int pred = 0;
        float score = 1.f;
        if (gp < a.total_pairs && !(z0 >= fmaxf(z1, z2))) {
          const float zz[3] = {z0, z1, z2};
          if (C == 2) classify_vals<PENEO_DT_F32, 2>(zz, pred, score);
          else classify_vals<PENEO_DT_F32, 3>(zz, pred, score);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, pred != 0);
        const int64_t slot = (tile * kNumHeads + k) * 4 + q;
        if (lane == 0) a.spots.cnt[slot] = __popc(mask);
        if (pred != 0) {
          const int64_t at = slot * 32 + __popc(mask & ((1u << lane) - 1u));
          a.spots.meta[at] = row | (pred << 8);
          a.spots.score[at] = score;
        }
      } else if (gp < a.total_pairs) {
        float* dst = a.logits[k] + gp * C;
        dst[0] = z0, dst[1] = z1;
        if (C == 3) dst[2] = z2;
        if (LOSS) {  // w[t] (logsumexp(z) - z[t]) and w[t] of this pair (model/custom_loss.py:189-202)
          const long long tag = a.tags[k][gp];
          const bool bad = tag < 0 || tag >= C;
          const int t = bad ? 0 : static_cast<int>(tag);
          const float mx = fmaxf(fmaxf(z0, z1), z2);
          const float se = __expf(z0 - mx) + __expf(z1 - mx) + (C == 3 ? __expf(z2 - mx) : 0.f);
          const float zt = t == 0 ? z0 : (t == 1 ? z1 : z2);
          const float ww = t == 0 ? a.class_w[0] : (t == 1 ? a.class_w[1] : a.class_w[2]);
#pragma unroll
          for (int kk = 0; kk < kNumHeads; ++kk)
            if (kk == k) acc_w[kk] += ww, acc_l[kk] += bad ? NAN : ww * (mx + __logf(se) - zt);
        }
      }
    };
    // warp-uniform poll: lane 0 tests the barrier, the result is broadcast, then every lane passes the (already
    // completed) wait itself so that each has observed the phase before touching tensor memory
    auto poll = [&](uint64_t* bar, uint32_t parity) {
      int ok = lane == 0 ? static_cast<int>(ptx::mbar_try_wait(bar, parity)) : 0;
      ok = __shfl_sync(0xffffffffu, ok, 0);
      if (ok) ptx::mbar_wait(bar, parity);
      return ok != 0;
    };
    auto emit_ready = [&]() {
      while (next_emit < total_heads && poll(&bars[bZFull + (next_emit & 1)], (next_emit >> 1) & 1)) {
        emit_z(next_emit);
        ++next_emit;
      }
    };
    {  // the constant K step of the A operand: element 0 = 1.0 (bf16), the other 15 = 0; 16 columns, written once
      uint32_t one[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) one[x] = x == 0 ? 0x00003F80u : 0u;
      ptx::tmem_st_x16(tmem + lane_base + kColOne, one);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();  // ordered before this warp's first "S chunk full" arrival, which the MMA warp waits for
    }
    for (int it = 0; it < my_tiles; ++it) {
      const int64_t tile = tile_of(it);
      // (row offsets of a_i / b_j for the row this lane will later copy)
      int64_t my_a = -1, my_b = -1;
      {
        const int64_t gp = tile * 128 + q * 32 + lane;
        if (gp < a.total_pairs) {
          const int64_t b = gp / a.pairs_per_doc;
          const int p = static_cast<int>(gp - b * a.pairs_per_doc);
          int i, j;
          pair_from_flat(p, a.n, i, j);
          my_a = (b * a.n + i) * (2 * D);
          my_b = (b * a.n + j) * (2 * D) + D;
        }
      }
      // ---- generate s rows [32q, 32q+32) into staging; lane = 4-column group (3 groups per lane).  Consecutive pairs
      //      share their row token i (a tile is 128 consecutive pairs of the row-major triangle), so a_i stays in
      //      registers and is reloaded only where i changes; only the b_j rows are streamed, kProdRows at a time.
      if (SAVE) {  // the TMA stores of the previous tile's s have read this warp's rows of the staging boxes
        if (lane == 0) ptx::bulk_wait_group_read<0>();
        __syncwarp();
      }
      int64_t cur_a = -2;
      float af[3][4] = {};
#pragma unroll 1
      for (int rr = 0; rr < 32; rr += kProdRows) {
        if ((rr & 7) == 0) emit_ready();
        uint2 bv[kProdRows][3];
        int64_t offa[kProdRows];
#pragma unroll
        for (int u = 0; u < kProdRows; ++u) {
          offa[u] = __shfl_sync(0xffffffffu, my_a, rr + u);
          const int64_t offb = __shfl_sync(0xffffffffu, my_b, rr + u);
#pragma unroll
          for (int mth = 0; mth < 3; ++mth)
            bv[u][mth] = offa[u] >= 0 ? __ldg(reinterpret_cast<const uint2*>(a.ab + offb + 4 * (lane + 32 * mth))) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < kProdRows; ++u) {
          const int r = q * 32 + rr + u;
          if (offa[u] != cur_a) {  // (warp-uniform) new row token: its projection, bf16 -> fp32 is a 16-bit shift
            cur_a = offa[u];
#pragma unroll
            for (int mth = 0; mth < 3; ++mth) {
              const uint2 av = cur_a >= 0 ? __ldg(reinterpret_cast<const uint2*>(a.ab + cur_a + 4 * (lane + 32 * mth))) : make_uint2(0u, 0u);
              af[mth][0] = __uint_as_float(av.x << 16), af[mth][1] = __uint_as_float(av.x & 0xFFFF0000u);
              af[mth][2] = __uint_as_float(av.y << 16), af[mth][3] = __uint_as_float(av.y & 0xFFFF0000u);
            }
          }
#pragma unroll
          for (int mth = 0; mth < 3; ++mth) {
            const int cg = lane + 32 * mth;  // 4-column group index, 0..95
            const float b0 = __uint_as_float(bv[u][mth].x << 16), b1 = __uint_as_float(bv[u][mth].x & 0xFFFF0000u);
            const float b2 = __uint_as_float(bv[u][mth].y << 16), b3 = __uint_as_float(bv[u][mth].y & 0xFFFF0000u);
            uint2 o;
            o.x = ptx::pack_bf16x2(ptx::silu_from_half(af[mth][0] + b0), ptx::silu_from_half(af[mth][1] + b1));
            o.y = ptx::pack_bf16x2(ptx::silu_from_half(af[mth][2] + b2), ptx::silu_from_half(af[mth][3] + b3));
            // box = 64-feature K chunk, 128 B per row inside it, 16-byte chunk index XOR (row & 7): the layout a TMA store
            // with SWIZZLE_128B expects, conflict-free for these writes and for the row-wise reads below
            const int chunk16 = ((cg >> 1) & 7) ^ (r & 7);
            *reinterpret_cast<uint2*>(stg + (cg >> 4) * kStageBoxBytes + r * 128 + chunk16 * 16 + (cg & 1) * 8) = o;
          }
        }
      }
#ifndef PENEO_K2_ABLATE_SAVE_S  // (timing experiment only)
      if (SAVE) {
#else
      if (false) {
#endif
        // s of this warp's 32 pairs leaves as six TMA stores straight from the staging boxes (rows past the end of the
        // pair list are clipped); the boxes are rewritten only after wait_group.read at the top of the next tile
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const uint64_t pol = ptx::l2_policy_evict_first();
          const int32_t row0 = static_cast<int32_t>(tile * 128 + q * 32);
#pragma unroll
          for (int kc = 0; kc < kKChunks; ++kc)
            ptx::tma_store_2d_hint(&tmS, stg + kc * kStageBoxBytes + q * 32 * 128, kc * 64, row0, pol);
          ptx::bulk_commit_group();
        }
      }
      __syncwarp();
      // ---- copy into TMEM, one 64-feature K chunk at a time, as the MMA warp releases them
      const int r = q * 32 + lane;
#pragma unroll 1
      for (int kc = 0; kc < kKChunks; ++kc) {
        uint32_t v[32];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint4 t = *reinterpret_cast<const uint4*>(stg + kc * kStageBoxBytes + r * 128 + ((ch ^ (r & 7)) * 16));
          v[4 * ch] = t.x, v[4 * ch + 1] = t.y, v[4 * ch + 2] = t.z, v[4 * ch + 3] = t.w;
        }
        if (it > 0) {
          while (!poll(&bars[bSFree + kc], (it - 1) & 1)) emit_ready();
        }
        ptx::tc_fence_after();
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) lo[x] = v[x], hi[x] = v[16 + x];
        ptx::tmem_st_x16(tmem + lane_base + kColS + 32 * kc, lo);
        ptx::tmem_st_x16(tmem + lane_base + kColS + 32 * kc + 16, hi);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(&bars[bSFull + kc]);
      }
      __syncwarp();
    }
    if (SAVE && lane == 0) ptx::bulk_wait_group<0>();  // every s store of this warp has been written
    while (next_emit < total_heads) {  // the heads of the last tile
      ptx::mbar_wait(&bars[bZFull + (next_emit & 1)], (next_emit >> 1) & 1);
      emit_z(next_emit);
      ++next_emit;
    }
    if (LOSS) {  // this warp's partial sums -> (quadrant, head) slots: fixed shuffle tree, fp64 from here on
#pragma unroll
      for (int k = 0; k < kNumHeads; ++k) {
        double dl = static_cast<double>(acc_l[k]), dw = static_cast<double>(acc_w[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dl += __shfl_xor_sync(0xffffffffu, dl, o), dw += __shfl_xor_sync(0xffffffffu, dw, o);
        if (lane == 0) s_loss[(q * 5 + k) * 2] = dl, s_loss[(q * 5 + k) * 2 + 1] = dw;
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // the peer's tensor-memory reads / remote arrives are finished as well
  if (LOSS && threadIdx.x < 10) {  // this CTA's partial sums, quadrants added in index order
    const int k = threadIdx.x >> 1, j = threadIdx.x & 1;
    double acc = 0.0;
    for (int qq = 0; qq < 4; ++qq) acc += s_loss[(qq * 5 + k) * 2 + j];
    a.loss_partial[(static_cast<size_t>(k) * gridDim.x + blockIdx.x) * 2 + j] = acc;
  }
  if (warp == 2) ptx::tmem_dealloc_2sm(tmem, 512);
}

}  // namespace k2p

#ifdef PENEO_K2_PROFILE
extern "C" int peneo_debug_k2_profile(unsigned long long* out_host, int reset) {
  if (cudaMemcpyFromSymbol(out_host, k2p::g_k2_prof, 8 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(k2p::g_k2_prof, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
#endif

int launch_pair_heads_tc_pair(const void* pack, const PackLayout& L, const __nv_bfloat16* ab, int batch, int n,
                              float* const logits[kNumHeads], cudaStream_t st, const DropSpec* drop, const FusedLossFwd* loss,
                              int* grid_out, const TileSpots* spots) {
  using namespace k2p;
  const char* base = static_cast<const char*>(pack);
  Args a{};
  a.ab = ab;
  a.bmid_half = reinterpret_cast<const float*>(base + L.bmid_half);
  a.bout = reinterpret_cast<const float*>(base + L.bout);
  for (int h = 0; h < kNumHeads; ++h) a.logits[h] = logits ? logits[h] : nullptr;
  if (spots) a.spots = *spots;
  a.n = n;
  a.pairs_per_doc = static_cast<int32_t>(pair_count(n));
  a.total_pairs = (int64_t)batch * a.pairs_per_doc;
  const int64_t tiles = (a.total_pairs + 127) / 128;
  PENEO_REQUIRE(tiles < (1ll << 31), "pair_heads: too many pairs for one launch");
  a.num_tiles = static_cast<int32_t>(tiles);
  if (tiles == 0) return PENEO_OK;
  alignas(64) CUtensorMap tmW, tmO, tmH, tmS;
  int rc;
  // boxes are the per-CTA halves: 64 of a chunk's 128 W_mid rows, 8 of its 16 (padded) W_out rows
  if ((rc = make_tensor_map_bf16(&tmW, base + L.wmid_bf16, D, 5 * D, D * 2, 64, kChunkCols / 2)) != PENEO_OK) return rc;
  if ((rc = make_tensor_map_bf16(&tmO, base + L.wout_bf16, 128, 15 * 16, 128 * 2, 64, 8)) != PENEO_OK) return rc;
  int dev = 0, sms = 148;
  PENEO_CUDA_TRY(cudaGetDevice(&dev));
  PENEO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t pairs = (tiles + 1) / 2;
  const int grid = 2 * static_cast<int>(std::min<int64_t>(pairs, sms / 2));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmemBytes, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  const bool dr = drop && drop->thresh;
  if (dr) {
    a.drop_thresh = drop->thresh, a.drop_scale = drop->scale;
    for (int h = 0; h < kNumHeads; ++h) a.drop_key[h] = drop_key(*drop, site_head(h, 0));
  }
  if (loss) {
    for (int h = 0; h < kNumHeads; ++h) a.tags[h] = loss->tags[h];
    for (int c = 0; c < 3; ++c) a.class_w[c] = loss->class_w[c];
    a.loss_partial = loss->partial;
    a.save_h = loss->save_h, a.save_s = loss->save_s;
    PENEO_REQUIRE((a.save_h == nullptr) == (a.save_s == nullptr), "pair_heads: save_h and save_s go together");
  }
  const bool save = loss && a.save_h;
  if (save) {
    PENEO_REQUIRE(a.total_pairs < (1ll << 31), "pair_heads: too many pairs for the saved activations");
    if ((rc = make_tensor_map_bf16_sw64(&tmH, a.save_h, 5 * D, a.total_pairs, 5 * D * 2, 32, 32)) != PENEO_OK) return rc;
    if ((rc = make_tensor_map_bf16(&tmS, a.save_s, D, a.total_pairs, D * 2, 64, 32)) != PENEO_OK) return rc;
  } else {
    tmH = tmW, tmS = tmW;  // (unused)
  }
  if (grid_out) *grid_out = grid;
#define GO(DR, LO, SP, SV)                                                                                         \
  do {                                                                                                             \
    PENEO_CUDA_TRY(cudaFuncSetAttribute(pair_heads_tc_kernel<DR, LO, SP, SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)); \
    PENEO_CUDA_TRY(cudaLaunchKernelEx(&cfg, pair_heads_tc_kernel<DR, LO, SP, SV>, tmW, tmO, tmH, tmS, a));         \
  } while (0)
  PENEO_REQUIRE(!(spots && (dr || loss)), "pair_heads: the spots-only output is an inference mode (no dropout, no loss)");
  if (spots) GO(false, false, true, false);
  else if (dr && save) GO(true, true, false, true);
  else if (save) GO(false, true, false, true);
  else if (dr && loss) GO(true, true, false, false);
  else if (dr) GO(true, false, false, false);
  else if (loss) GO(false, true, false, false);
  else GO(false, false, false, false);
#undef GO
  PENEO_CUDA_TRY(cudaGetLastError());
  return PENEO_OK;
}

}  // namespace peneo
