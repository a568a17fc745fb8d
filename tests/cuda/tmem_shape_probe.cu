// Prints the thread <-> (lane, column) mapping of the 16-lane tcgen05.ld/st shapes.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "sm100.cuh"
using namespace sm100;

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) { tmem_alloc<1>(&tmem_base, 64); tmem_relinquish<1>(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_base;
  // fill: value = lane * 1000 + col, lanes 0..127, cols 0..63 (32x32b: thread i <-> lane i)
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    for (int i = 0; i < 16; ++i) r[i] = (uint32_t)(tid * 1000 + c0 + i);
    tmem_st16(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
  }
  tmem_wait_st();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 1) {   // use warp 1 (lanes 32..63) to see the lane base handling
    uint32_t a[4], b[2], c[1], d[4];
    const uint32_t base0 = tb + ((uint32_t)(32) << 16), base16 = tb + ((uint32_t)(48) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base0 + 8));
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(base0 + 8));
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(c[0]) : "r"(base0 + 8));
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(base16 + 8));
    tmem_wait_ld();
    uint32_t* o = out + lane * 16;
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = a[3]; o[4] = b[0]; o[5] = b[1]; o[6] = c[0];
    o[7] = d[0]; o[8] = d[1]; o[9] = d[2]; o[10] = d[3];
    // x2 repeat of 16x256b: where do the next registers come from?
    uint32_t e[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]), "=r"(e[4]), "=r"(e[5]), "=r"(e[6]), "=r"(e[7]) : "r"(base0));
    tmem_wait_ld();
    o[11] = e[4]; o[12] = e[5]; o[13] = e[6]; o[14] = e[7];
  }
  tc_fence_before(); __syncthreads();
  // store test: 16x128b.x1 from warp 2 into cols 0..3 of lanes 64..79, then read back with 32x32b
  if (warp == 2) {
    uint32_t s0 = 500000 + lane * 10, s1 = 500000 + lane * 10 + 1;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" :: "r"(tb + ((uint32_t)64 << 16)), "r"(s0), "r"(s1));
    tmem_wait_st();
    tc_fence_before(); __syncwarp(); tc_fence_after();
    uint32_t r[16];
    tmem_ld16(tb + ((uint32_t)64 << 16), r);
    tmem_wait_ld();
    uint32_t* o = out + 32 * 16 + lane * 4;
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tb, 64);
}

int main() {
  uint32_t* d; cudaMalloc(&d, 4096 * 4); cudaMemset(d, 0, 4096 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  uint32_t h[4096]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("value = lane*1000 + col.  ld issued at lane base 32 (and 48), column base 8\n");
  for (int t = 0; t < 32; ++t) {
    uint32_t* o = h + t * 16;
    printf("t%02d 16x256b:[%u %u %u %u] 16x128b:[%u %u] 16x64b:[%u] 16x256b@48:[%u %u %u %u] x2 regs4-7@col0:[%u %u %u %u]\n", t, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], o[10], o[11], o[12], o[13], o[14]);
  }
  printf("st 16x128b.x1 by thread t of (500000+10t, +1) at lane base 64 -> readback lanes 64..95 cols 0..3\n");
  for (int t = 0; t < 32; ++t) { uint32_t* o = h + 32 * 16 + t * 4; printf("lane%02d: %u %u %u %u\n", 64 + t, o[0], o[1], o[2], o[3]); }
  return 0;
}
