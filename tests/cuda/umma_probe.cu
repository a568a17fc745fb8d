// Bring-up probes for the tcgen05 building blocks (run on the B200 box; one probe per process so that a
// hang in one cannot take the others down):
//   ./umma_probe ss1   cta_group::1, A and B from smem (K-major, SWIZZLE_128B written by threads)
//   ./umma_probe ts1   cta_group::1, A from TMEM (tcgen05.st), B from smem
//   ./umma_probe ss2   cta_group::2 (CTA pair, M=256), SS
//   ./umma_probe ts2   cta_group::2, TS
//   ./umma_probe tma   TMA 2D fp32 tile load with SWIZZLE_128B + thread-per-row de-swizzled read
// Each prints "PROBE <name> PASS|FAIL max_err=..." and exits 0/1.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "sm100.cuh"

using namespace sm100;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

constexpr int KD = 64;  // K of the probe GEMMs (4 UMMA k-steps)

// A_rows x KD halves -> K-major SW128 tile
__device__ void fill_sw128(uint8_t* tile, const __half* src, int rows, int tid, int nthreads) {
  for (int idx = tid; idx < rows * (KD / 8); idx += nthreads) {
    const int r = idx / (KD / 8), c = idx % (KD / 8);
    *reinterpret_cast<uint4*>(tile + sw128_offset(r, c)) = *reinterpret_cast<const uint4*>(src + (size_t)r * KD + c * 8);
  }
}

// ------------------------------------------------------------------------------------------------
// cta_group::1 : D[128 x N] = A[128 x KD] * B[N x KD]^T
template <bool TS, int N>
__global__ void __launch_bounds__(128) probe_cg1(const __half* A, const __half* B, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (!TS) fill_sw128(sA, A, 128, tid, 128);
  fill_sw128(sB, B, N, tid, 128);
  fence_proxy_async();
  if (warp == 0) {
    tmem_alloc<1>(&tmem_base, 512);
    tmem_relinquish<1>();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  const uint32_t tmem_d = tb, tmem_a = tb + 256;
  if (TS) {
    // lane = row: K packed two halves per 32-bit column
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)tid * KD);
    uint32_t r[16];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = arow[h * 16 + i];
      tmem_st16(tmem_a + ((uint32_t)(warp * 32) << 16) + h * 16, r);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t da = umma_desc_k_sw128(smem_u32(sA)), db = umma_desc_k_sw128(smem_u32(sB));
    for (int s = 0; s < KD / 16; ++s) {
      if (TS)
        umma_ts<1>(tmem_d, tmem_a + s * 8, db + (uint64_t)(s * 2), idesc, s > 0);
      else
        umma_ss<1>(tmem_d, da + (uint64_t)(s * 2), db + (uint64_t)(s * 2), idesc, s > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(size_t)row * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tb, 512);
}

// ------------------------------------------------------------------------------------------------
// cta_group::2 : D[256 x N] = A[256 x KD] * B[N x KD]^T ; CTA c owns A rows [128c, 128c+128) and B rows
// [N/2 c, N/2 c + N/2); each CTA's TMEM receives its 128 rows x all N columns.
template <bool TS, int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe_cg2(const __half* A, const __half* B, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 128;
  __shared__ uint64_t bar_done;   // MMA complete (multicast commit arrives in both CTAs)
  __shared__ uint64_t bar_ready;  // leader only: both CTAs' operands are in place
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cta = cluster_ctarank();
  if (!TS) fill_sw128(sA, A + (size_t)cta * 128 * KD, 128, tid, 128);
  fill_sw128(sB, B + (size_t)cta * (N / 2) * KD, N / 2, tid, 128);
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    mbar_init(&bar_ready, 2);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc<2>(&tmem_base, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  const uint32_t tmem_d = tb, tmem_a = tb + 256;
  if (TS) {
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + ((size_t)cta * 128 + tid) * KD);
    uint32_t r[16];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = arow[h * 16 + i];
      tmem_st16(tmem_a + ((uint32_t)(warp * 32) << 16) + h * 16, r);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) mbar_arrive_cluster(&bar_ready, 0);  // tell the leader this CTA's operands are ready
  if (cta == 0 && tid == 0) {
    mbar_wait_cluster(&bar_ready, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(256, N);
    const uint64_t da = umma_desc_k_sw128(smem_u32(sA)), db = umma_desc_k_sw128(smem_u32(sB));
    for (int s = 0; s < KD / 16; ++s) {
      if (TS)
        umma_ts<2>(tmem_d, tmem_a + s * 8, db + (uint64_t)(s * 2), idesc, s > 0);
      else
        umma_ss<2>(tmem_d, da + (uint64_t)(s * 2), db + (uint64_t)(s * 2), idesc, s > 0);
    }
    umma_commit_2sm(&bar_done, 3);
  }
  mbar_wait_cluster(&bar_done, 0);
  tc_fence_after();
  const int row = (int)cta * 128 + warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(size_t)row * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 0) tmem_dealloc<2>(tb, 512);
}

// ------------------------------------------------------------------------------------------------
// TMA: fp32 [rows x 96] matrix, box 32 cols x 128 rows, SWIZZLE_128B; thread r reads row r de-swizzled.
__global__ void __launch_bounds__(128) probe_tma(const __grid_constant__ CUtensorMap tmap, float* out, int c0, int r0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, 128 * 128);
    tma_load_2d(smem, &tmap, c0, r0, &bar);
  }
  mbar_wait(&bar, 0);
  for (int c = 0; c < 8; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(smem + tid * 128 + ((c ^ (tid & 7)) << 4));
    *reinterpret_cast<float4*>(out + tid * 32 + c * 4) = v;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int run_gemm(const char* name, bool two, bool ts, int N) {
  const int M = two ? 256 : 128;
  std::vector<__half> hA((size_t)M * KD), hB((size_t)N * KD);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(7);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, (size_t)M * N * 4));
  const int smem = 128 * 128 + 256 * 128 + 1024;
#define LAUNCH(K, GRID)                                                                     \
  CK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));            \
  K<<<GRID, 128, smem>>>(dA, dB, dD);
  if (!two) {
    if (N == 128) { if (ts) { LAUNCH((probe_cg1<true, 128>), 1) } else { LAUNCH((probe_cg1<false, 128>), 1) } }
    else          { if (ts) { LAUNCH((probe_cg1<true, 256>), 1) } else { LAUNCH((probe_cg1<false, 256>), 1) } }
  } else {
    if (N == 128) { if (ts) { LAUNCH((probe_cg2<true, 128>), 2) } else { LAUNCH((probe_cg2<false, 128>), 2) } }
    else          { if (ts) { LAUNCH((probe_cg2<true, 256>), 2) } else { LAUNCH((probe_cg2<false, 256>), 2) } }
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hD((size_t)M * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < KD; ++k) ref += (double)fA[(size_t)m * KD + k] * fB[(size_t)n * KD + k];
      double e = fabs(ref - hD[(size_t)m * N + n]);
      if (!(e <= maxerr)) maxerr = e;  // NaN-propagating
    }
  const bool ok = maxerr < 1e-3;
  printf("PROBE %s N=%d %s max_err=%g\n", name, N, ok ? "PASS" : "FAIL", maxerr);
  return ok ? 0 : 1;
}

static int run_tma() {
  const int ROWS = 300, COLS = 96;
  std::vector<float> h((size_t)ROWS * COLS);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  CK(cudaMalloc(&d, h.size() * 4));
  CK(cudaMalloc(&o, 128 * 32 * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  if (!enc) { printf("PROBE tma FAIL no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tm;
  cuuint64_t dims[2] = {COLS, ROWS};
  cuuint64_t strides[1] = {COLS * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("PROBE tma FAIL encode rc=%d\n", (int)r); return 1; }
  int bad = 0;
  for (int trial = 0; trial < 2; ++trial) {
    const int c0 = trial ? 64 : 32, r0 = trial ? 256 : 128;  // second trial runs past the last row: OOB rows must be 0
    const int smem = 128 * 128 + 1024;
    CK(cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_tma<<<1, 128, smem>>>(tm, o, c0, r0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> ho(128 * 32);
    CK(cudaMemcpy(ho.data(), o, ho.size() * 4, cudaMemcpyDeviceToHost));
    for (int rr = 0; rr < 128; ++rr)
      for (int c = 0; c < 32; ++c) {
        const float ref = (r0 + rr < ROWS) ? h[(size_t)(r0 + rr) * COLS + c0 + c] : 0.f;
        if (ho[rr * 32 + c] != ref) ++bad;
      }
  }
  printf("PROBE tma %s mismatches=%d\n", bad ? "FAIL" : "PASS", bad);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  const char* p = argc > 1 ? argv[1] : "ss1";
  const int N = argc > 2 ? atoi(argv[2]) : 128;
  if (!strcmp(p, "ss1")) return run_gemm("ss1", false, false, N);
  if (!strcmp(p, "ts1")) return run_gemm("ts1", false, true, N);
  if (!strcmp(p, "ss2")) return run_gemm("ss2", true, false, N);
  if (!strcmp(p, "ts2")) return run_gemm("ts2", true, true, N);
  if (!strcmp(p, "tma")) return run_tma();
  printf("unknown probe %s\n", p);
  return 2;
}
