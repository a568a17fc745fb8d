// Tensor-pipe pacing probe (B200 box): how many SM cycles one tcgen05.mma kind::f16 costs inside a long dependent-free
// stream, by shape and operand mode, on ALL SMs at once (74 CTA pairs) -- the numbers the row-pass kernel's budget is
// built from.  Data are whatever the memories hold (timing only).
//   ./umma_rate            prints one line per case:  mode cta_group N  cycles/MMA  (MMAs issued, total cycles)
//   optional arg "ld": warps 4-11 of every CTA stream tcgen05.ld (16x256b.x8) concurrently (TMEM port contention)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

struct Args {
  int n;          // MMA N
  int ts;         // 1: A from TMEM, 0: A from smem
  int reps;       // MMAs in the stream
  int ld;         // concurrent tcgen05.ld traffic from 8 other warps
  int alt;        // 1: alternate N = n and N = 128 like the row pass (6 x N128 then 24 x n ...)
  long long* out; // [grid] cycles
};

template <int CG>
__global__ void __launch_bounds__(384, 1) rate_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cta = CG == 2 ? cluster_ctarank() : 0;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    stop = 0;
  }
  if (warp == 2) {
    tmem_alloc<CG>(&tmem_base, 512);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  if (warp == 1 && cta == 0) {
    const uint32_t idesc = umma_idesc_f16(CG == 2 ? 256 : 128, a.n), idesc128 = umma_idesc_f16(CG == 2 ? 256 : 128, 128);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
    const long long t0 = clock64();
    if (elect_one()) {      // one thread issues the whole stream (an elect per MMA costs ~170 cycles of its own)
#pragma unroll 1
      for (int i0 = 0; i0 < a.reps; i0 += 30) {
#pragma unroll
        for (int j = 0; j < 30; ++j) {
          const int i = i0 + j;
          const bool wide = a.alt && j < 6;
          const uint64_t db = umma_desc_k_sw128(sb + (uint32_t)(j & 7) * 4096u + (uint32_t)(j & 3) * 32u);
          if (a.ts) {
            if (CG == 2) umma_ts<2>(tm + (wide ? 0u : 128u), tm + 384u + (uint32_t)(j & 15) * 8u, db, wide ? idesc128 : idesc, 1u);
            else umma_ts<1>(tm + (wide ? 0u : 128u), tm + 384u + (uint32_t)(j & 15) * 8u, db, wide ? idesc128 : idesc, 1u);
          } else {
            const uint64_t da = umma_desc_k_sw128(sa + (uint32_t)(j & 1) * 16384u + (uint32_t)(j & 3) * 32u);
            if (CG == 2) umma_ss<2>(tm + (wide ? 0u : 128u), da, db, wide ? idesc128 : idesc, 1u);
            else umma_ss<1>(tm + (wide ? 0u : 128u), da, db, wide ? idesc128 : idesc, 1u);
          }
          (void)i;
        }
      }
    }
    __syncwarp();
    if (elect_one()) {
      if (CG == 2) umma_commit_2sm(&bar, 1); else umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) a.out[blockIdx.x] = t1 - t0;
    stop = 1;
    if (CG == 2 && lane == 0) {   // tell the peer CTA's load warps too
      asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(mapa(smem_u32((const void*)&stop), 1)), "r"(1) : "memory");
    }
  } else if (warp >= 4 && a.ld) {
    const uint32_t lane_addr = (uint32_t)(((warp - 4) & 3) * 32 + ((warp - 4) >> 2) * 16) << 16;
    uint32_t acc = 0;
    while (!stop) {
      uint32_t v[32];
      tmem_ld_16x256b_x8(tm + lane_addr + (uint32_t)(acc & 1u) * 64u, v);
      tmem_wait_ld();
      acc += v[lane & 31];
    }
    if (acc == 0x12345678u) a.out[0] = 0;
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tm, 512);
}

template <int CG>
void run(const char* name, int n, int ts, int ld, int alt, long long* d_out) {
  const int grid = 148, reps = 6000;
  Args a{n, ts, reps, ld, alt, d_out};
  const size_t smem = 32768 + 65536 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(d_out, 0, sizeof(long long) * grid));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int it = 0; it < 2; ++it) {
    CK(cudaLaunchKernelEx(&cfg, rate_kernel<CG>, a));
    CK(cudaDeviceSynchronize());
  }
  long long h[148];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  double sum = 0;
  int cnt = 0;
  for (int i = 0; i < grid; i += CG) { sum += (double)h[i]; ++cnt; }
  printf("%-4s cta_group::%d N=%3d%s%s  %7.1f cycles/MMA\n", name, CG, n, alt ? " alt(6xN128,24xN)" : "", ld ? " +ld traffic" : "",
         sum / cnt / reps);
}

int main(int argc, char** argv) {
  const int ld = argc > 1 && !strcmp(argv[1], "ld");
  long long* d_out;
  CK(cudaMalloc(&d_out, sizeof(long long) * 148));
  for (int n : {32, 64, 128, 256}) run<2>("TS", n, 1, ld, 0, d_out);
  for (int n : {64, 128, 256}) run<2>("SS", n, 0, ld, 0, d_out);
  for (int n : {32, 64, 128, 256}) run<1>("TS", n, 1, ld, 0, d_out);
  run<2>("TS", 64, 1, ld, 1, d_out);
  return 0;
}
