// On-device unit checks of the tcgen05 / TMA building blocks and small throughput probes.
// Exposed through peneo_selftest() / peneo_probe_rates(); used by tests/ and bench.py only.
#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace peneo {

// D[128, N] = A[128, K] * B[N, K]^T with A staged in TMEM (TS form), K = 64 * kblocks.
__global__ void __launch_bounds__(128, 1)
    ts_mma_test_kernel(const __grid_constant__ CUtensorMap tmB, const uint32_t* __restrict__ a_packed,
                       float* __restrict__ d_out, int N, int kblocks) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // by offset: keeps the shared address space
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_full, 1);
    ptx::mbar_init(&bar_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const int row = warp * 32 + lane;
  // A: row `row`, K/2 packed words -> TMEM columns [256, 256 + K/2)
  for (int kb = 0; kb < kblocks; ++kb) {
    uint32_t lo[16], hi[16];
    for (int x = 0; x < 16; ++x) {
      lo[x] = a_packed[row * (kblocks * 32) + kb * 32 + x];
      hi[x] = a_packed[row * (kblocks * 32) + kb * 32 + 16 + x];
    }
    ptx::tmem_st_x16(tmem + lane_base + 256 + kb * 32, lo);
    ptx::tmem_st_x16(tmem + lane_base + 256 + kb * 32 + 16, hi);
  }
  ptx::tmem_st_wait();
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::tc_fence_after();
    const uint32_t stage_bytes = N * 128;
    ptx::mbar_arrive_expect_tx(&bar_full, stage_bytes * kblocks);
    for (int kb = 0; kb < kblocks; ++kb) ptx::tma_load_2d(smem + kb * stage_bytes, &tmB, &bar_full, kb * 64, 0);
    ptx::mbar_wait(&bar_full, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
    for (int kb = 0; kb < kblocks; ++kb)
      for (int ks = 0; ks < 4; ++ks)
        ptx::umma_ts(tmem, tmem + 256 + kb * 32 + ks * 8,
                     ptx::umma_desc_sw128(ptx::smem_u32(smem + kb * stage_bytes) + ks * 32), idesc, (kb | ks) != 0);
    ptx::tc_commit(&bar_done);
  }
  ptx::mbar_wait(&bar_done, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    ptx::tmem_ld_x16(tmem + lane_base + c0, r);
    ptx::tmem_ld_wait();
    for (int x = 0; x < 16; ++x) d_out[row * N + c0 + x] = __uint_as_float(r[x]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// D[128, N] = sum_k At[k, m] Bt[k, n]: both operands MN-major, loaded as [64 k-rows x 64 cols] SWIZZLE_128B boxes.
// At: [K, 128], Bt: [K, N] row-major; K = 64 * kblocks.  `swap` exchanges the roles of LBO / SBO (diagnostic).
__global__ void __launch_bounds__(128, 1)
    mn_mma_test_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       float* __restrict__ d_out, int N, int kblocks, int swap) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // by offset: keeps the shared address space
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_full, 1);
    ptx::mbar_init(&bar_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int nbB = N / 64;                        // 64-column blocks of B
  const uint32_t blk = 64 * 128;                 // one [64 k x 64 col] box = 8 KB
  const uint32_t stage = (2 + nbB) * blk;        // per k-block: 2 A boxes then nbB B boxes
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar_full, stage * kblocks);
    for (int kb = 0; kb < kblocks; ++kb) {
      unsigned char* st = smem + kb * stage;
      for (int mb = 0; mb < 2; ++mb) ptx::tma_load_2d(st + mb * blk, &tmA, &bar_full, mb * 64, kb * 64);
      for (int nb = 0; nb < nbB; ++nb) ptx::tma_load_2d(st + (2 + nb) * blk, &tmB, &bar_full, nb * 64, kb * 64);
    }
    ptx::mbar_wait(&bar_full, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::umma_idesc_bf16_major(128, N, true, true);
    const uint32_t lbo = swap ? 1024u : blk, sbo = swap ? blk : 1024u;
    for (int kb = 0; kb < kblocks; ++kb)
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t a = ptx::smem_u32(smem + kb * stage) + ks * 2048;
        const uint32_t b = ptx::smem_u32(smem + kb * stage + 2 * blk) + ks * 2048;
        ptx::umma_ss(tmem, ptx::umma_desc_mn_sw128(a, lbo, sbo), ptx::umma_desc_mn_sw128(b, lbo, sbo), idesc,
                     (kb | ks) != 0);
      }
    ptx::tc_commit(&bar_done);
  }
  ptx::mbar_wait(&bar_done, 0);
  ptx::tc_fence_after();
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    ptx::tmem_ld_x16(tmem + lane_base + c0, r);
    ptx::tmem_ld_wait();
    for (int x = 0; x < 16; ++x) d_out[row * N + c0 + x] = __uint_as_float(r[x]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

static float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
static uint16_t bf16_bits(float v) {
  __nv_bfloat16 b = __float2bfloat16_rn(v);
  uint16_t u;
  std::memcpy(&u, &b, 2);
  return u;
}

struct Lcg {
  uint64_t s;
  float next() {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return static_cast<float>((s >> 40) & 0xFFFF) / 32768.0f - 1.0f;
  }
};

static int check_ts(int N, int kblocks, double& max_err, std::string& note) {
  const int K = 64 * kblocks;
  Lcg rng{12345ull + N * 131 + kblocks};
  std::vector<float> A(128 * K), B(N * K);
  for (auto& v : A) v = bf16_round(rng.next());
  for (auto& v : B) v = bf16_round(rng.next());
  std::vector<uint32_t> ap(128 * K / 2);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < K; k += 2)
      ap[r * (K / 2) + k / 2] = static_cast<uint32_t>(bf16_bits(A[r * K + k])) |
                                (static_cast<uint32_t>(bf16_bits(A[r * K + k + 1])) << 16);
  std::vector<uint16_t> bb(N * K);
  for (int i = 0; i < N * K; ++i) bb[i] = bf16_bits(B[i]);
  uint32_t* d_ap = nullptr;
  uint16_t* d_b = nullptr;
  float* d_d = nullptr;
  PENEO_CUDA_TRY(cudaMalloc(&d_ap, ap.size() * 4));
  PENEO_CUDA_TRY(cudaMalloc(&d_b, bb.size() * 2));
  PENEO_CUDA_TRY(cudaMalloc(&d_d, 128 * N * 4));
  PENEO_CUDA_TRY(cudaMemcpy(d_ap, ap.data(), ap.size() * 4, cudaMemcpyHostToDevice));
  PENEO_CUDA_TRY(cudaMemcpy(d_b, bb.data(), bb.size() * 2, cudaMemcpyHostToDevice));
  PENEO_CUDA_TRY(cudaMemset(d_d, 0xFF, 128 * N * 4));
  alignas(64) CUtensorMap tmB;
  int rc = make_tensor_map_bf16(&tmB, d_b, K, N, K * 2, 64, N);
  if (rc != PENEO_OK) return rc;
  const int smem = N * 128 * kblocks + 1024;
  PENEO_CUDA_TRY(cudaFuncSetAttribute(ts_mma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ts_mma_test_kernel<<<1, 128, smem>>>(tmB, d_ap, d_d, N, kblocks);
  PENEO_CUDA_TRY(cudaGetLastError());
  PENEO_CUDA_TRY(cudaDeviceSynchronize());
  std::vector<float> D(128 * N);
  PENEO_CUDA_TRY(cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_ap), cudaFree(d_b), cudaFree(d_d);
  max_err = 0.0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < N; ++c) {
      double ref = 0.0;
      for (int k = 0; k < K; ++k) ref += static_cast<double>(A[r * K + k]) * B[c * K + k];
      const double e = std::fabs(ref - D[r * N + c]);
      if (!(e <= max_err)) max_err = e;  // also catches NaN
    }
  (void)note;
  return PENEO_OK;
}

static int check_mn(int N, int kblocks, int swap, double& max_err) {
  const int K = 64 * kblocks;
  Lcg rng{4242ull + N * 17 + kblocks};
  std::vector<float> A((size_t)K * 128), B((size_t)K * N);
  for (auto& v : A) v = bf16_round(rng.next());
  for (auto& v : B) v = bf16_round(rng.next());
  std::vector<uint16_t> ab(A.size()), bb(B.size());
  for (size_t i = 0; i < A.size(); ++i) ab[i] = bf16_bits(A[i]);
  for (size_t i = 0; i < B.size(); ++i) bb[i] = bf16_bits(B[i]);
  uint16_t *d_a = nullptr, *d_b = nullptr;
  float* d_d = nullptr;
  PENEO_CUDA_TRY(cudaMalloc(&d_a, ab.size() * 2));
  PENEO_CUDA_TRY(cudaMalloc(&d_b, bb.size() * 2));
  PENEO_CUDA_TRY(cudaMalloc(&d_d, 128 * N * 4));
  PENEO_CUDA_TRY(cudaMemcpy(d_a, ab.data(), ab.size() * 2, cudaMemcpyHostToDevice));
  PENEO_CUDA_TRY(cudaMemcpy(d_b, bb.data(), bb.size() * 2, cudaMemcpyHostToDevice));
  PENEO_CUDA_TRY(cudaMemset(d_d, 0xFF, 128 * N * 4));
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA, d_a, 128, K, 128 * 2, 64, 64)) != PENEO_OK) return rc;
  if ((rc = make_tensor_map_bf16(&tmB, d_b, N, K, N * 2, 64, 64)) != PENEO_OK) return rc;
  const int smem = (2 + N / 64) * 64 * 128 * kblocks + 1024;
  PENEO_CUDA_TRY(cudaFuncSetAttribute(mn_mma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mn_mma_test_kernel<<<1, 128, smem>>>(tmA, tmB, d_d, N, kblocks, swap);
  PENEO_CUDA_TRY(cudaGetLastError());
  PENEO_CUDA_TRY(cudaDeviceSynchronize());
  std::vector<float> D(128 * N);
  PENEO_CUDA_TRY(cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_a), cudaFree(d_b), cudaFree(d_d);
  max_err = 0.0;
  for (int m = 0; m < 128; ++m)
    for (int c = 0; c < N; ++c) {
      double ref = 0.0;
      for (int k = 0; k < K; ++k) ref += static_cast<double>(A[(size_t)k * 128 + m]) * B[(size_t)k * N + c];
      const double e = std::fabs(ref - D[m * N + c]);
      if (!(e <= max_err)) max_err = e;
    }
  return PENEO_OK;
}

// TF32 GEMM (gemm_tf32.cu) in one of the four storage combinations, ragged M / N / K, against fp64 on TF32-rounded
// inputs (error budget: accumulation only).
static float tf32_round(float v) {
  uint32_t u;
  std::memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;  // round to 10 mantissa bits
  std::memcpy(&v, &u, 4);
  return v;
}
static int check_tf32(int combo, double& max_err) {
  const bool ta = combo & 1, tb = combo & 2;
  const int M = 200, N = 256, K = 100;
  const int lda = ta ? 208 : 104, ldb = tb ? 104 : 256;  // multiples of 4 >= the contiguous extent
  Lcg rng{9001ull + combo};
  std::vector<float> A((size_t)(ta ? K : M) * lda, 0.f), B((size_t)(tb ? N : K) * ldb, 0.f);
  auto a_at = [&](int m, int k) -> float& { return ta ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]; };
  auto b_at = [&](int k, int n) -> float& { return tb ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n]; };
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) a_at(m, k) = tf32_round(rng.next());
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) b_at(k, n) = tf32_round(rng.next());
  float *dA, *dB, *dC;
  PENEO_CUDA_TRY(cudaMalloc(&dA, A.size() * 4));
  PENEO_CUDA_TRY(cudaMalloc(&dB, B.size() * 4));
  PENEO_CUDA_TRY(cudaMalloc(&dC, (size_t)M * N * 4));
  PENEO_CUDA_TRY(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  PENEO_CUDA_TRY(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  PENEO_CUDA_TRY(cudaMemset(dC, 0, (size_t)M * N * 4));
  Tf32Gemm g;
  g.ta = ta, g.tb = tb, g.A = dA, g.lda = lda, g.B = dB, g.ldb = ldb, g.C = dC, g.ldc = N, g.M = M, g.N = N, g.K = K;
  g.mode = combo == 3 ? 2 : 0;  // exercise the split-K atomic path once
  int rc = launch_gemm_tf32(g, 0);
  if (rc != PENEO_OK) return rc;
  PENEO_CUDA_TRY(cudaDeviceSynchronize());
  std::vector<float> C((size_t)M * N);
  PENEO_CUDA_TRY(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA), cudaFree(dB), cudaFree(dC);
  max_err = 0.0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0.0;
      for (int k = 0; k < K; ++k) ref += static_cast<double>(a_at(m, k)) * b_at(k, n);
      const double e = std::fabs(ref - C[(size_t)m * N + n]);
      if (!(e <= max_err)) max_err = e;
    }
  return PENEO_OK;
}

static int check_ss(int M, int N, int K, double& max_err) {
  Lcg rng{777ull + M + N * 3 + K * 7};
  std::vector<float> A((size_t)M * K), W((size_t)N * K), bias(N);
  for (auto& v : A) v = bf16_round(rng.next());
  for (auto& v : W) v = bf16_round(rng.next());
  for (auto& v : bias) v = rng.next();
  std::vector<uint16_t> ab(A.size()), wb(W.size());
  for (size_t i = 0; i < A.size(); ++i) ab[i] = bf16_bits(A[i]);
  for (size_t i = 0; i < W.size(); ++i) wb[i] = bf16_bits(W[i]);
  __nv_bfloat16 *dA, *dW, *dC;
  float* dbias;
  PENEO_CUDA_TRY(cudaMalloc(&dA, ab.size() * 2));
  PENEO_CUDA_TRY(cudaMalloc(&dW, wb.size() * 2));
  PENEO_CUDA_TRY(cudaMalloc(&dC, (size_t)M * N * 2));
  PENEO_CUDA_TRY(cudaMalloc(&dbias, N * 4));
  PENEO_CUDA_TRY(cudaMemcpy(dA, ab.data(), ab.size() * 2, cudaMemcpyHostToDevice));
  PENEO_CUDA_TRY(cudaMemcpy(dW, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
  PENEO_CUDA_TRY(cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice));
  int rc = launch_gemm_tc(dA, K, dW, K, dbias, dC, N, M, N, K, 0, 0);
  if (rc != PENEO_OK) return rc;
  PENEO_CUDA_TRY(cudaDeviceSynchronize());
  std::vector<uint16_t> cb((size_t)M * N);
  PENEO_CUDA_TRY(cudaMemcpy(cb.data(), dC, cb.size() * 2, cudaMemcpyDeviceToHost));
  cudaFree(dA), cudaFree(dW), cudaFree(dC), cudaFree(dbias);
  max_err = 0.0;
  for (int r = 0; r < M; ++r)
    for (int c = 0; c < N; ++c) {
      double ref = bias[c];
      for (int k = 0; k < K; ++k) ref += static_cast<double>(A[(size_t)r * K + k]) * W[(size_t)c * K + k];
      uint32_t u = static_cast<uint32_t>(cb[(size_t)r * N + c]) << 16;
      float got;
      std::memcpy(&got, &u, 4);
      const double e = std::fabs(ref - got) / (1.0 + std::fabs(ref));
      if (!(e <= max_err)) max_err = e;
    }
  return PENEO_OK;
}

int run_selftest(uint32_t* failed_mask, char* report, size_t report_bytes) {
  uint32_t mask = 0;
  std::string rep;
  char line[256];
  struct {
    const char* name;
    int kind, a, b, c;
    double tol;
  } cases[] = {
      {"ss_gemm_128x128x64", 0, 128, 128, 64, 1e-2},   {"ss_gemm_200x256x384", 0, 200, 256, 384, 1e-2},
      {"ts_mma_n128_k64", 1, 128, 1, 0, 1e-3},          {"ts_mma_n128_k384", 1, 128, 6, 0, 2e-3},
      {"ts_mma_n16_k128", 1, 16, 2, 0, 1e-3},
      {"mn_major_n128_k64", 2, 128, 1, 0, 1e-3},         {"mn_major_n256_k192", 2, 256, 3, 0, 2e-3},
      {"tf32_gemm_nt", 3, 2, 0, 0, 1e-4},  // (the MN-major TF32 combinations are known not to work: see gemm_tf32.cu)
  };
  int idx = 0;
  for (auto& cs : cases) {
    double err = 0.0;
    std::string note;
    int rc = cs.kind == 0 ? check_ss(cs.a, cs.b, cs.c, err)
             : cs.kind == 1 ? check_ts(cs.a, cs.b, err, note)
             : cs.kind == 2 ? check_mn(cs.a, cs.b, cs.c, err)
                            : check_tf32(cs.a, err);
    if (cs.kind == 2 && rc == PENEO_OK && !(err <= cs.tol) && getenv("PENEO_SELFTEST_MN_SWAP")) {
      double err2 = 0.0;  // diagnostic: LBO / SBO exchanged
      if (check_mn(cs.a, cs.b, 1, err2) == PENEO_OK) note = " (swapped LBO/SBO: max_err=" + std::to_string(err2) + ")";
    }
    bool ok = rc == PENEO_OK && err <= cs.tol;
    if (!ok) mask |= 1u << idx;
    snprintf(line, sizeof line, "%-24s %s max_err=%.3e%s%s%s\n", cs.name, ok ? "ok  " : "FAIL", err,
             rc != PENEO_OK ? " error: " : "", rc != PENEO_OK ? get_error() : "", note.c_str());
    rep += line;
    if (rc == PENEO_E_CUDA) {  // sticky CUDA error: stop, later checks would only echo it
      for (int rest = idx + 1; rest < (int)(sizeof cases / sizeof cases[0]); ++rest) mask |= 1u << rest;
      break;
    }
    ++idx;
  }
  if (failed_mask) *failed_mask = mask;
  if (report && report_bytes) {
    strncpy(report, rep.c_str(), report_bytes - 1);
    report[report_bytes - 1] = 0;
  }
  return PENEO_OK;
}

// ------------------------------------------------------------------------------------------------
// throughput probes: ops/s of the units that co-limit K2's epilogue
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) rate_probe_kernel(float* out, int iters) {
  float v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = 0.001f * (threadIdx.x + u + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (KIND == 0) {
        v[u] = ptx::tanh_approx(v[u]);
      } else if (KIND == 1) {
        asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[u]));
      } else if (KIND == 2) {
        uint32_t w = __float_as_uint(v[u]);
        asm("tanh.approx.bf16x2 %0, %0;" : "+r"(w));
        v[u] = __uint_as_float(w);
      } else if (KIND == 3) {
        v[u] = fmaf(v[u], 1.0001f, 0.5f);
      } else {
        v[u] = ptx::silu_from_half(v[u]);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) s += v[u];
  if (s == 123.456f) out[0] = s;
}

template <int KIND>
static int probe(double ops_per_elem, double& rate) {
  float* d;
  PENEO_CUDA_TRY(cudaMalloc(&d, 4));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, iters = 4096;
  rate_probe_kernel<KIND><<<blocks, 256>>>(d, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_probe_kernel<KIND><<<blocks, 256>>>(d, iters);
  cudaEventRecord(e1);
  PENEO_CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(d);
  rate = ops_per_elem * blocks * 256.0 * 8.0 * iters / (ms * 1e-3);
  return PENEO_OK;
}

int run_probe_rates(double* out, int n_out) {
  double r[5] = {0, 0, 0, 0, 0};
  int rc;
  if ((rc = probe<0>(1.0, r[0])) != PENEO_OK) return rc;  // tanh.approx.f32 / s
  if ((rc = probe<1>(1.0, r[1])) != PENEO_OK) return rc;  // ex2.approx.f32 / s
  if ((rc = probe<2>(2.0, r[2])) != PENEO_OK) return rc;  // tanh.approx.bf16x2 elements / s
  if ((rc = probe<3>(1.0, r[3])) != PENEO_OK) return rc;  // dependent FFMA / s
  if ((rc = probe<4>(1.0, r[4])) != PENEO_OK) return rc;  // 1-MUFU SiLU / s
  for (int i = 0; i < n_out && i < 5; ++i) out[i] = r[i];
  return PENEO_OK;
}

}  // namespace peneo
