// tm_gemm.cu -- batched fp32 "NT" GEMM on tcgen05 for the TransMIL / Nystrom path:
//
//      C[b] = alpha * A[b] (M x K)  *  B[b]^T (N x K)   (+ diag * I) (+ bias[col]) (+ addend) (relu)
//
// Both operands are fp32, K contiguous (nn.Linear weights, q/k/v rows, attention rows).  The reference computes every one
// of these products in IEEE fp32 (torch.matmul / einsum / nn.Linear: nystrom_attention.py:83,119-121,135,147 and
// transMIL.py:61,89), so the tensor-core path is an error-compensated TF32 split
//      a b  ~=  a_hi b_hi + a_lo b_hi + a_hi b_lo ,   a_hi = a & 0xFFFFE000 (what the MMA reads of a), a_lo = a - a_hi (exact)
// i.e. 3 kind::tf32 MMAs per product with fp32 accumulation in TMEM (~2^-21 relative).  PRECISE = false runs one
// MMA on the raw operands (plain TF32, ~2^-10).
//
// Persistent CTAs (one per SM) walk the 128 x BN output tiles; 320 threads:
//   warp 0     TMA producer: 128 x 32 fp32 boxes of A and B (SWIZZLE_128B, 4-D maps: K, rows, batch_lo, batch_hi),
//              3-stage ring that keeps running across tile boundaries
//   warp 1     MMA issuer (one elected thread), two accumulators in TMEM, tcgen05.commit frees the ring stage
//   warps 2-5  produce the lo tile of each landed stage (the raw tile is the hi operand: the MMA ignores the low bits)
//   warps 6-9  epilogue of the previous tile while the next one accumulates: thread = row (tcgen05.ld 32x32b),
//              scale / diagonal / bias / addend / activation, optional second store of the transposed tile (the
//              Moore-Penrose iteration needs every product in both orientations).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "acmil_transmil.h"
#include "gp_common.cuh"
#include "sm100.cuh"

namespace {
using namespace sm100;

constexpr int GT = 448;          // threads per CTA: TMA, MMA, 4 splitter warps, 8 epilogue warps (two per TMEM lane quarter)
constexpr int BM = 128;          // tile rows
#ifndef TM_GEMM_KC
#define TM_GEMM_KC 32
#endif
// fp32 columns per stage: 32 = one 128-byte swizzle row, 3 stages of 64 KB.  -DTM_GEMM_KC=16 builds the 64-byte-swizzle
// variant with 6 stages of 32 KB: measured 20 % SLOWER on the ViT GEMMs (fc1 478 vs 396 us), i.e. the engine is bound by
// per-stage costs (barrier round trips, proxy fence, issue), not by the number of stages in flight.
constexpr int KC = TM_GEMM_KC;
constexpr int NST = KC == 16 ? 6 : 3;          // ring stages
constexpr uint32_t TILE_BYTES = BM * KC * 4;   // 16 KB (8 KB for KC = 16)
constexpr int NST_MAX = 6;
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t smem_addr) {
  return KC == 16 ? umma_desc_k_sw64(smem_addr) : umma_desc_k_sw128(smem_addr);
}

struct GemmParams {
  CUtensorMap ta, tb;            // (K, rows, batch_lo, batch_hi) fp32
  CUtensorMap tbh, tbl;          // fp16-split mode: (K, rows) fp16 hi / lo sections of a pre-split B image
  const float* bscale;           // fp16-split mode: 1 / (power-of-two scale the image was made with), on the device
  float2* stats_out;             // chunked softmax, producer side: the epilogue stores exp(c - chunk max) and (max, sum) per
                                 //   (batch, row, 32-column chunk) at ((z * M + row) * nch_out + chunk)
  const float2* stats_in;        // consumer side: A rows are such exponentials over K; the converter rescales chunk c of a row
                                 //   by exp(max_c - row max) / row sum  (stats of nch_in = ceil(K / 32) chunks per row)
  int nch_out, nch_in;
  float* c;
  float* ct;                     // optional transposed copy: ct[b] + col * ldct + row
  const float* bias;             // [N] or null
  const float* addend;           // optional, addressed like a plain row-major C with ld = ld_add
  int M, N, K;
  int a_batched, b_batched;      // 0: the operand is shared by all batch entries
  long long c_batch_stride, ldc;
  int cbw;                       // column blocks: element (row, col) of batch b lives at
  long long cbs;                 //   c + b * c_batch_stride + (col / cbw) * cbs + row * ldc + col % cbw
  long long ct_batch_stride, ldct;
  long long add_batch_stride, ld_add;
  float alpha, beta, diag;
  int act;                       // 0 none, 1 relu, 2 exact GELU
  int bias_row;                  // bias indexed by row
  int zdiv;                      // two-level batch: z_lo = z % zdiv, z_hi = z / zdiv
  long long c_bs2, ct_bs2, add_bs2;
  int vec_ok;                    // float4 stores are legal for c
  int add_vec_ok;                // float4 loads are legal for addend
  int bias_vec_ok;               // float4 loads are legal for a column bias
  int ksplit;                    // > 1: blockIdx.z = batch * ksplit + split; raw partial tiles go to split_ws
  int k_per_split;               // multiple of KC
  float* split_ws;               // [batch * ksplit][M][N]
  int ntiles;
};

struct Bars {
  uint64_t full[NST_MAX];        // TMA bytes landed
  uint64_t split[NST_MAX];       // lo tile / TMEM operands written (4 converter warps)
  uint64_t empty[NST_MAX];       // MMAs that read the stage have completed
  uint64_t acc_full[2];          // accumulator buffer complete (tcgen05.commit)
  uint64_t acc_empty[2];         // accumulator buffer drained by the 8 epilogue warps
  uint32_t tmem;
};

__host__ __device__ constexpr uint32_t tf32_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

// D[tmem] (+)= A[tmem: lane = row, one tf32 per column] * B[smem]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// nn.GELU() (exact, erf based) on two values at once with packed fp32 arithmetic:  gelu(x) = x/2 (1 + erf(x / sqrt 2)),
// erf(|z|) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p |z|)  (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7, i.e.
// one fp32 ulp of the 1 + erf factor).  ~10 issue slots per element instead of ~30 for erff: the GELU epilogue of the MLP's
// first GEMM competes with the converter warps for issue slots.
__device__ __forceinline__ uint64_t gelu2(uint64_t x2) {
  const float x0 = lo2(x2), x1 = hi2(x2);
  const uint64_t az = pack2(fabsf(x0) * 0.70710678118654752440f, fabsf(x1) * 0.70710678118654752440f);      // |z|
  const uint64_t den = fma2(pack2(0.3275911f, 0.3275911f), az, pack2(1.f, 1.f));
  const uint64_t t = pack2(rcp_approx(lo2(den)), rcp_approx(hi2(den)));
  uint64_t pl = fma2(pack2(1.061405429f, 1.061405429f), t, pack2(-1.453152027f, -1.453152027f));
  pl = fma2(pl, t, pack2(1.421413741f, 1.421413741f));
  pl = fma2(pl, t, pack2(-0.284496736f, -0.284496736f));
  pl = fma2(pl, t, pack2(0.254829592f, 0.254829592f));
  pl = mul2(pl, t);
  const uint64_t ex = mul2(mul2(az, az), pack2(-1.4426950408889634f, -1.4426950408889634f));                  // -z^2 log2(e)
  const uint64_t e = pack2(ex2_approx(lo2(ex)), ex2_approx(hi2(ex)));
  const uint64_t erfa = fma2(mul2(pl, pack2(-1.f, -1.f)), e, pack2(1.f, 1.f));                                 // erf(|z|)
  const uint64_t erfs = pack2(copysignf(lo2(erfa), x0), copysignf(hi2(erfa), x1));
  const uint64_t hx = mul2(x2, pack2(0.5f, 0.5f));
  return fma2(erfs, hx, hx);
}

__device__ __forceinline__ float act_apply_tm(float t, int act) {
  if (act == 1) return fmaxf(t, 0.f);
  if (act == 2) return 0.5f * t * (1.f + erff(t * 0.70710678118654752440f));      // nn.GELU() (exact)
  return t;
}

// smem per stage: {A (raw fp32 = the hi operand), B hi, [B lo]}; B tiles hold BN rows.  In the precise mode the A operands
// of the MMAs live in TMEM (TS mode): the converter warps read the landed A tile once and write A (raw) and A_lo there,
// so shared memory carries the TMA fill, ONE read of A, the B split and only B's operand reads -- 128 KB per K chunk
// instead of 192 KB (SS mode with A_lo in shared memory), which is what bounds this engine (profiles/ncu_r1_gemm_summary.md).
template <int BN>
__host__ __device__ constexpr uint32_t stage_bytes(bool precise) {
  return TILE_BYTES + (uint32_t)BN * KC * 4 * (precise ? 2u : 1u);
}
template <int BN>
__host__ __device__ constexpr int ring_stages(bool precise) {      // as many as fit next to the barriers in 227 KB
  return (precise && KC == 32) ? (BN <= 64 ? 6 : 4) : NST;
}
constexpr uint32_t A_SLOT_COLS = 2 * KC;      // TMEM columns of one A operand slot: KC raw + KC lo

struct TileCoord {
  int m0, n0, bz, split;
};

// tiles: n fastest (CTAs that run concurrently share the A rows in L2), then m, then batch * ksplit
__device__ __forceinline__ TileCoord tile_coord(const GemmParams& p, int tile, int ntn, int ntm, int bn) {
  TileCoord t;
  t.n0 = (tile % ntn) * bn;
  t.m0 = ((tile / ntn) % ntm) * BM;
  const int z = tile / (ntn * ntm);
  t.bz = p.ksplit > 1 ? z / p.ksplit : z;
  t.split = p.ksplit > 1 ? z % p.ksplit : 0;
  return t;
}

// Epilogue of one 128 x BN tile for one warp: TMEM lanes 32 q .. 32 q + 31 (thread = row), the 32-column chunks `half`,
// `half` + 2, ...; scale / diagonal / bias / addend / activation on registers, float4 global traffic where the layout allows.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const TileCoord& t, int nchunk, uint32_t tacc, int q, int half,
                                              int lane, float alpha) {
  const int zlo = t.bz % p.zdiv, zhi = t.bz / p.zdiv;
  const int row = t.m0 + q * 32 + lane;
  const bool row_ok = row < p.M;
  float* crow = p.c ? p.c + (size_t)zlo * p.c_batch_stride + (size_t)zhi * p.c_bs2 + (size_t)row * p.ldc : nullptr;
  const float* arow =
      p.addend ? p.addend + (size_t)zlo * p.add_batch_stride + (size_t)zhi * p.add_bs2 + (size_t)row * p.ld_add : nullptr;
  float* ctb = p.ct ? p.ct + (size_t)zlo * p.ct_batch_stride + (size_t)zhi * p.ct_bs2 + row : nullptr;
  const int zsplit = p.ksplit > 1 ? t.bz * p.ksplit + t.split : 0;
  const float rbias = (p.bias && p.bias_row && row_ok) ? p.bias[row] : 0.f;
#pragma unroll 1
  for (int c0 = half * 32; c0 < BN; c0 += 64) {
    const int col0 = t.n0 + c0;
    if (col0 >= p.N) break;
    uint32_t v[32];
    tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + c0, v);
    tmem_wait_ld();
    const bool full = col0 + 32 <= p.N;      // warp-uniform
    if (p.ksplit > 1) {
      if (row_ok) {
        float* dst = p.split_ws + ((size_t)zsplit * p.M + row) * p.N + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) dst[j] = nchunk > 0 ? __uint_as_float(v[j]) : 0.f;
      }
      continue;
    }
    float r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = alpha * __uint_as_float(v[j]);
    if (p.diag != 0.f) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j == row) r[j] += p.diag;
    }
    if (p.bias) {
      if (p.bias_row) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] += rbias;
      } else if (full && p.bias_vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(p.bias + col0 + j);      // same address in every lane: broadcast
          r[j] += b4.x; r[j + 1] += b4.y; r[j + 2] += b4.z; r[j + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) r[j] += p.bias[col0 + j];
      }
    }
    if (row_ok) {
      if (arow) {
        if (full && p.add_vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(arow + col0 + j);
            r[j] = fmaf(p.beta, a4.x, r[j]); r[j + 1] = fmaf(p.beta, a4.y, r[j + 1]);
            r[j + 2] = fmaf(p.beta, a4.z, r[j + 2]); r[j + 3] = fmaf(p.beta, a4.w, r[j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) r[j] = fmaf(p.beta, arow[col0 + j], r[j]);
        }
      }
      if (p.stats_out) {      // softmax numerators against the chunk's own maximum; the consumer GEMM finishes the softmax
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) mx = fmaxf(mx, r[j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          r[j] = col0 + j < p.N ? __expf(r[j] - mx) : 0.f;
          sum += r[j];
        }
        p.stats_out[((size_t)t.bz * p.M + row) * p.nch_out + (col0 >> 5)] = make_float2(mx, sum);
      }
      if (p.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const uint64_t g = gelu2(pack2(r[j], r[j + 1]));
          r[j] = lo2(g);
          r[j + 1] = hi2(g);
        }
      } else if (p.act) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = act_apply_tm(r[j], p.act);
      }
      if (crow) {
        // column blocks: the chunk sits inside one block whenever its first and last column do
        const int blk0 = col0 / p.cbw, off0 = col0 - blk0 * p.cbw;
        const bool one_blk = off0 + 32 <= p.cbw;
        float* dst = crow + (size_t)blk0 * p.cbs + off0;
        if (full && one_blk && p.vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        } else if (one_blk) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) dst[j] = r[j];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col < p.N) crow[(size_t)(col / p.cbw) * p.cbs + col % p.cbw] = r[j];
          }
        }
      }
      if (ctb) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col < p.N) ctb[(size_t)col * p.ldct] = r[j];      // lanes = consecutive rows: coalesced
        }
      }
    }
  }
}

// Persistent: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the smem ring and the two TMEM
// accumulators run across tile boundaries, so TMA prefetch, MMAs and the epilogue of consecutive tiles overlap.
template <int BN, bool PRECISE>
__global__ void __launch_bounds__(GT) tm_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t B_BYTES = (uint32_t)BN * KC * 4;
  constexpr uint32_t STAGE = stage_bytes<BN>(PRECISE);
  constexpr int NS = ring_stages<BN>(PRECISE);
  constexpr uint32_t ACC_COLS = 2 * (BN < 32 ? 32 : BN);            // two accumulators
  constexpr uint32_t TM_A = ACC_COLS;                               // A operand slots behind them (precise mode)
  constexpr uint32_t TM_NEED = ACC_COLS + (PRECISE ? NS * A_SLOT_COLS : 0);
  constexpr uint32_t TM_COLS = TM_NEED <= 64 ? 64 : (TM_NEED <= 128 ? 128 : (TM_NEED <= 256 ? 256 : 512));
  static_assert(TM_NEED <= 512, "TMEM budget");
  Bars* bars = reinterpret_cast<Bars*>(smem + NS * STAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntn = (p.N + BN - 1) / BN, ntm = (p.M + BM - 1) / BM;
  const int ntiles = p.ntiles;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->split[s], 4);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], 8);
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.ta);
    tma_prefetch_desc(&p.tb);
  }
  if (warp == 1) {
    tmem_alloc<1>(&bars->tmem, TM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = bars->tmem;

  auto k_range = [&](const TileCoord& t, int& kbeg, int& nchunk) {
    kbeg = p.ksplit > 1 ? t.split * p.k_per_split : 0;
    const int kend = p.ksplit > 1 ? min(p.K, kbeg + p.k_per_split) : p.K;
    nchunk = kend > kbeg ? (kend - kbeg + KC - 1) / KC : 0;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ctr = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord t = tile_coord(p, tile, ntn, ntm, BN);
        int kbeg, nchunk;
        k_range(t, kbeg, nchunk);
        const int zlo = t.bz % p.zdiv, zhi = t.bz / p.zdiv;
        for (int c = 0; c < nchunk; ++c, ++ctr) {
          const uint32_t s = ctr % NS, ph = (ctr / NS) & 1u;
          mbar_wait(&bars->empty[s], ph ^ 1u);
          unsigned char* st = smem + s * STAGE;
          mbar_expect_tx(&bars->full[s], TILE_BYTES + B_BYTES);
          tma_load_4d(st, &p.ta, kbeg + c * KC, t.m0, p.a_batched ? zlo : 0, p.a_batched ? zhi : 0, &bars->full[s]);
          tma_load_4d(st + TILE_BYTES, &p.tb, kbeg + c * KC, t.n0, p.b_batched ? zlo : 0, p.b_batched ? zhi : 0, &bars->full[s]);
        }
      }
    }
  } else if (warp == 1) {
    {   // the whole warp runs the loop on identical values; the tcgen05 instructions sit under elect_one() so that
        // ptxas emits them directly instead of inside an ELECT / BRA.U.ANY loop per MMA
      constexpr uint32_t idesc = tf32_idesc(BM, BN);
      uint32_t ctr = 0, it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const TileCoord t = tile_coord(p, tile, ntn, ntm, BN);
        int kbeg, nchunk;
        k_range(t, kbeg, nchunk);
        const uint32_t buf = it & 1u;
        mbar_wait(&bars->acc_empty[buf], ((it >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tm + buf * (ACC_COLS / 2);
        for (int c = 0; c < nchunk; ++c, ++ctr) {
          const uint32_t s = ctr % NS, ph = (ctr / NS) & 1u;
          mbar_wait(PRECISE ? &bars->split[s] : &bars->full[s], ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + s * STAGE), b_hi = a_hi + TILE_BYTES, b_lo = b_hi + B_BYTES;
          const uint32_t ta_hi = tm + TM_A + s * A_SLOT_COLS, ta_lo = ta_hi + KC;
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KC / 8; ++k) {      // one MMA covers K = 8 tf32 = 32 bytes of the swizzle row / 8 TMEM columns
            const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
            if constexpr (PRECISE) {
              // the tensor core ignores the 13 low mantissa bits of a TF32 operand (measured: identical results with
              // and without masking), so the raw fp32 value IS the hi operand; only the lo parts are produced
              umma_tf32_ts(tacc, ta_lo + k * 8, umma_desc_k(b_hi + k * 32), idesc, acc);
              umma_tf32_ts(tacc, ta_hi + k * 8, umma_desc_k(b_lo + k * 32), idesc, 1u);
              umma_tf32_ts(tacc, ta_hi + k * 8, umma_desc_k(b_hi + k * 32), idesc, 1u);
            } else {
              umma_tf32(tacc, umma_desc_k(a_hi + k * 32), umma_desc_k(b_hi + k * 32), idesc, acc);
            }
          }
          umma_commit(&bars->empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (nchunk > 0) umma_commit(&bars->acc_full[buf]);
          else mbar_arrive(&bars->acc_full[buf]);      // empty K range (k-split tail): the epilogue writes zeros
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    if constexpr (PRECISE) {
      // Converter warps 2..5 (TMEM lane quarters 2, 3, 0, 1): thread = row of the A tile.  A: the landed fp32 row is read
      // once (de-swizzled LDS.128, conflict-free: the 8 lanes of a quarter-warp hit 8 different 16-byte chunks), the raw
      // values and lo = x - (x with the 13 low mantissa bits cleared) go to the stage's TMEM operand slot.  B: lo written
      // at the same swizzled position of the B_lo tile (linear float4 slots, layout-agnostic).
      static_assert(KC == 32, "the TMEM converter assumes 128-byte rows");
      const int qd = warp & 3;
      const int r = qd * 32 + lane;                      // A tile row == TMEM lane
      const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
      const int ct = (warp - 2) * 32 + lane;             // 0..127: linear slot owner for B
      constexpr int NVB = B_BYTES / 16;                  // float4 slots of B
      uint32_t ctr = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord t = tile_coord(p, tile, ntn, ntm, BN);
        int kbeg, nchunk;
        k_range(t, kbeg, nchunk);
        // chunked softmax, consumer side: this thread's A row holds exp(s - chunk max); finish the softmax on the way to TMEM
        const float2* srow = nullptr;
        float row_max = 0.f, row_inv = 0.f;
        if (p.stats_in && t.m0 + r < p.M) {
          srow = p.stats_in + ((size_t)t.bz * p.M + t.m0 + r) * p.nch_in;
          row_max = -INFINITY;
          for (int i = 0; i < p.nch_in; ++i) row_max = fmaxf(row_max, srow[i].x);
          float l = 0.f;
          for (int i = 0; i < p.nch_in; ++i) l += srow[i].y * __expf(srow[i].x - row_max);
          row_inv = 1.f / l;
        }
        for (int c = 0; c < nchunk; ++c, ++ctr) {
          const uint32_t s = ctr % NS, ph = (ctr / NS) & 1u;
          const float f = srow ? __expf(srow[(kbeg >> 5) + c].x - row_max) * row_inv : (p.stats_in ? 0.f : 1.f);
          mbar_wait(&bars->full[s], ph);
          const unsigned char* rowp = smem + s * STAGE + (r >> 3) * 1024 + (r & 7) * 128;
          const uint32_t ta = tm + lane_addr + TM_A + s * A_SLOT_COLS;
#pragma unroll
          for (int h = 0; h < 2; ++h) {                  // 16 columns at a time keeps the register footprint small
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 v = *reinterpret_cast<const float4*>(rowp + (((4 * h + i) ^ (r & 7)) << 4));
              if (p.stats_in) { v.x *= f; v.y *= f; v.z *= f; v.w *= f; }
              hi[4 * i] = __float_as_uint(v.x); hi[4 * i + 1] = __float_as_uint(v.y);
              hi[4 * i + 2] = __float_as_uint(v.z); hi[4 * i + 3] = __float_as_uint(v.w);
              lo[4 * i] = __float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
              lo[4 * i + 1] = __float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
              lo[4 * i + 2] = __float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
              lo[4 * i + 3] = __float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
            }
            tmem_st16(ta + 16 * h, hi);
            tmem_st16(ta + KC + 16 * h, lo);
          }
          const float4* bhi = reinterpret_cast<const float4*>(smem + s * STAGE + TILE_BYTES);
          float4* blo = reinterpret_cast<float4*>(smem + s * STAGE + TILE_BYTES + B_BYTES);
#pragma unroll 8
          for (int i = ct; i < NVB; i += 128) {
            const float4 v = bhi[i];
            float4 l;
            l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            blo[i] = l;
          }
          tmem_wait_st();
          tc_fence_before();
          fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core's async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->split[s]);
        }
      }
    }
  } else {
    // ---------------- epilogue warps 6..13: warp w may touch TMEM lanes 32 (w % 4) .. +31; the two warps of a lane
    // quarter take alternate 32-column chunks.  Thread = row: all per-element work is on registers, global traffic is
    // float4 wherever the layout allows, and the column-block split is resolved once per chunk. ----------------
    const int q = warp & 3, half = (warp - 6) >> 2;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const TileCoord t = tile_coord(p, tile, ntn, ntm, BN);
      int kbeg, nchunk;
      k_range(t, kbeg, nchunk);
      const uint32_t buf = it & 1u;
      mbar_wait(&bars->acc_full[buf], (it >> 1) & 1u);
      tc_fence_after();
      epilogue_tile<BN>(p, t, nchunk, tm + buf * (ACC_COLS / 2), q, half, lane, p.alpha);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);      // this warp's TMEM reads of the buffer are complete
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tm, TM_COLS);
}

// ------------------------------------------------------------------------------------------------------------------
// fp16-split mode (precise = 2): the products whose B operand is a WEIGHT (nn.Linear / conv weights: static, so its split
// is made once, acmil_gemm_split_b).  a b ~= a_hi b_hi + a_lo b_hi + a_hi b_lo with 11-bit hi parts and fp16 lo parts
// (b pre-scaled by a power of two so that its lo parts stay in fp16's normal range): 3 kind::f16 MMAs of K = 16 where the
// TF32 split needs 3 of K = 8 -- half the tensor-pipe time -- and per 64 columns of K a CTA moves 32 KB of A (TMA) +
// 2 x BN x 128 B of B (TMA) into shared memory, reads A once (converter) and B three times (MMAs): 144 KB against
// 256 KB for the same K in the TF32 engine, which is what bounds it (profiles/ncu_r1_gemm_summary.md).  Error: the a_lo b_lo
// term (2^-22) and the fp16 rounding of the lo parts (2^-23 of |a|, absolute floor 2^-25): fp32-level for |a| < 65504.
// Same roles as tm_gemm_kernel: warp 0 TMA, warp 1 MMA issue, warps 2-5 A converters (TMEM operand slots), warps 6-13 epilogue.
// -DTM_GEMM_PROF=1: CTA 0 of the fp16-split kernels adds clock64() spans to p.split_ws (unused in that mode; 16 x u64):
// [0] producer waits for a free stage, [1] producer total, [2] MMA warp waits for converted operands, [3] MMA warp waits for a
// drained accumulator, [4] MMA warp total, [5] converter waits for TMA bytes, [6] converter total, [7] epilogue waits for an
// accumulator, [8] epilogue total  (tests/cuda/gemm_h_prof.py)
#ifndef TM_GEMM_PROF
#define TM_GEMM_PROF 0
#endif
#if TM_GEMM_PROF
#define TPROF_DECL() long long tp_w0 = 0, tp_w1 = 0; const long long tp_start = clock64()
#define TPROF_WAIT(slot, stmt) do { const long long t0_ = clock64(); stmt; tp_w##slot += clock64() - t0_; } while (0)
#define TPROF_FLUSH(i0, i1, itot, cond)                                                                    \
  do {                                                                                                     \
    if ((cond) && p.split_ws != nullptr && blockIdx.x == 0) {                                              \
      unsigned long long* o = reinterpret_cast<unsigned long long*>(p.split_ws);                           \
      atomicAdd(o + (i0), (unsigned long long)tp_w0);                                                      \
      if ((i1) >= 0) atomicAdd(o + ((i1) >= 0 ? (i1) : 0), (unsigned long long)tp_w1);                     \
      atomicAdd(o + (itot), (unsigned long long)(clock64() - tp_start));                                   \
    }                                                                                                      \
  } while (0)
#else
#define TPROF_DECL()
#define TPROF_WAIT(slot, stmt) stmt
#define TPROF_FLUSH(i0, i1, itot, cond)
#endif
constexpr int KCH = 64;                       // K columns per stage
constexpr uint32_t H_A_SLOT = 64;             // TMEM columns per A operand slot: 32 packed hi + 32 packed lo
// threads of the fp16-split kernels: TMA, MMA, EIGHT converter warps (two per TMEM lane quarter, one 32-column half of the stage
// each: with four, the converters were the bound -- ~1 070 cycles of work per stage against 768 of MMA time, in-kernel
// counters of tests/cuda/gemm_h_prof.py), 8 epilogue warps
constexpr int GTH = 576;
constexpr int H_EPI_WARP0 = 10;
template <int BN>
__host__ __device__ constexpr uint32_t h_stage_bytes() {
  return 2u * TILE_BYTES + 2u * (uint32_t)BN * 128u;      // two 128 x 32 fp32 A tiles, B hi, B lo (BN rows of 64 halves)
}
template <int BN>
__host__ __device__ constexpr int h_stages() {
  return BN <= 64 ? 4 : 3;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));      // low half = a
  return r;
}
// two fp32 -> packed (hi, hi) and (lo, lo): hi = the top 11 significant bits (exact in fp16), lo = the rest rounded to fp16
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  const float bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  hi = pack_h2(ah, bh);
  lo = pack_h2(a - ah, b - bh);
}

template <int BN>
__global__ void __launch_bounds__(GTH) tm_gemm_h_kernel(const __grid_constant__ GemmParams p) {
  static_assert(KC == 32, "the fp16-split kernel stages A as 128-byte fp32 rows");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t BH_BYTES = (uint32_t)BN * 128u;
  constexpr uint32_t STAGE = h_stage_bytes<BN>();
  constexpr int NS = h_stages<BN>();
  constexpr uint32_t ACC_COLS = 2 * BN;
  constexpr uint32_t TM_A = ACC_COLS;
  constexpr uint32_t TM_NEED = ACC_COLS + NS * H_A_SLOT;
  constexpr uint32_t TM_COLS = TM_NEED <= 256 ? 256 : 512;
  static_assert(TM_NEED <= 512 && BN >= 32, "TMEM budget");
  Bars* bars = reinterpret_cast<Bars*>(smem + NS * STAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntn = (p.N + BN - 1) / BN, ntm = (p.M + BM - 1) / BM;
  const int ntiles = p.ntiles;
  const int nchunk = (p.K + KCH - 1) / KCH;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->split[s], 8);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], 8);
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.ta);
    tma_prefetch_desc(&p.tbh);
    tma_prefetch_desc(&p.tbl);
  }
  if (warp == 1) {
    tmem_alloc<1>(&bars->tmem, TM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = bars->tmem;

  if (warp == 0) {
    if (lane == 0) {
      TPROF_DECL();
      uint32_t ctr = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord t = tile_coord(p, tile, ntn, ntm, BN);
        const int zlo = p.a_batched ? t.bz % p.zdiv : 0, zhi = p.a_batched ? t.bz / p.zdiv : 0;
        for (int c = 0; c < nchunk; ++c, ++ctr) {
          const uint32_t s = ctr % NS, ph = (ctr / NS) & 1u;
          TPROF_WAIT(0, mbar_wait(&bars->empty[s], ph ^ 1u));
          unsigned char* st = smem + s * STAGE;
          const bool second = c * KCH + 32 < p.K;      // the K tail may end inside the first 32-column box of the chunk
          mbar_expect_tx(&bars->full[s], (second ? 2u : 1u) * TILE_BYTES + 2u * BH_BYTES);
          tma_load_4d(st, &p.ta, c * KCH, t.m0, zlo, zhi, &bars->full[s]);
          if (second) tma_load_4d(st + TILE_BYTES, &p.ta, c * KCH + 32, t.m0, zlo, zhi, &bars->full[s]);
          tma_load_2d(st + 2 * TILE_BYTES, &p.tbh, c * KCH, t.n0, &bars->full[s]);
          tma_load_2d(st + 2 * TILE_BYTES + BH_BYTES, &p.tbl, c * KCH, t.n0, &bars->full[s]);
        }
      }
      TPROF_FLUSH(0, -1, 1, true);
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
    TPROF_DECL();
    uint32_t ctr = 0, it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      TPROF_WAIT(1, mbar_wait(&bars->acc_empty[buf], ((it >> 1) & 1u) ^ 1u));      // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tm + buf * (ACC_COLS / 2);
      for (int c = 0; c < nchunk; ++c, ++ctr) {
        const uint32_t s = ctr % NS, ph = (ctr / NS) & 1u;
        TPROF_WAIT(0, mbar_wait(&bars->full[s], ph); mbar_wait(&bars->split[s], ph));      // B landed, A operands in TMEM
        tc_fence_after();
        const uint32_t b_hi = smem_u32(smem + s * STAGE + 2 * TILE_BYTES), b_lo = b_hi + BH_BYTES;
        const uint32_t ta_hi = tm + TM_A + s * H_A_SLOT, ta_lo = ta_hi + 32;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KCH / 16; ++k) {      // one MMA covers K = 16 halves = 32 bytes of the swizzle row = 8 TMEM columns
            const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
            umma_ts<1>(tacc, ta_lo + k * 8, umma_desc_k_sw128(b_hi + k * 32), idesc, acc);
            umma_ts<1>(tacc, ta_hi + k * 8, umma_desc_k_sw128(b_lo + k * 32), idesc, 1u);
            umma_ts<1>(tacc, ta_hi + k * 8, umma_desc_k_sw128(b_hi + k * 32), idesc, 1u);
          }
          umma_commit(&bars->empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bars->acc_full[buf]);
      __syncwarp();
    }
    TPROF_FLUSH(2, 3, 4, lane == 0);
  } else if (warp < H_EPI_WARP0) {
    // converters: thread = row of the A tile; the two landed 32-column fp32 tiles are read once (de-swizzled LDS.128) and
    // written to the stage's TMEM operand slot as packed fp16 pairs: columns [0, 32) hi, [32, 64) lo
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const int h = (warp - 2) >> 2;      // which 32-column half of the stage this warp converts
    TPROF_DECL();
    uint32_t ctr = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int c = 0; c < nchunk; ++c, ++ctr) {
        const uint32_t s = ctr % NS, ph = (ctr / NS) & 1u;
        TPROF_WAIT(0, mbar_wait(&bars->full[s], ph));
        const uint32_t ta = tm + lane_addr + TM_A + s * H_A_SLOT;
        {
          const unsigned char* rowp = smem + s * STAGE + h * TILE_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
          const bool landed = h == 0 || c * KCH + 32 < p.K;      // a second box wholly past K is not loaded: zeros
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = landed ? *reinterpret_cast<const float4*>(rowp + ((i ^ (r & 7)) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            split_h2(v.x, v.y, hi[2 * i], lo[2 * i]);
            split_h2(v.z, v.w, hi[2 * i + 1], lo[2 * i + 1]);
          }
          tmem_st16(ta + 16 * h, hi);
          tmem_st16(ta + 32 + 16 * h, lo);
        }
        TPROF_WAIT(1, tmem_wait_st());
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->split[s]);
      }
    }
    TPROF_FLUSH(5, 9, 6, warp == 2 && lane == 0);
  } else {
    const int q = warp & 3, half = (warp - H_EPI_WARP0) >> 2;
    const float alpha = p.alpha * (p.bscale ? __ldg(p.bscale) : 1.f);
    TPROF_DECL();
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const TileCoord t = tile_coord(p, tile, ntn, ntm, BN);
      const uint32_t buf = it & 1u;
      TPROF_WAIT(0, mbar_wait(&bars->acc_full[buf], (it >> 1) & 1u));
      tc_fence_after();
      epilogue_tile<BN>(p, t, nchunk, tm + buf * (ACC_COLS / 2), q, half, lane, alpha);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
    }
    TPROF_FLUSH(7, -1, 8, warp == H_EPI_WARP0 && lane == 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tm, TM_COLS);
}

// ------------------------------------------------------------------------------------------------------------------
// The fp16-split kernel on CTA pairs (cta_group::2): the engine is bound by the L2 -> SM delivery rate (~42 B/clk per SM,
// profiles/ncu_r2_gemm_h_summary.md), so the pair computes a 256 x BN tile of which every SM stages its own 128 rows of A but
// only HALF of B (rows [n0 + rank BN/2, + BN/2)): 48 KB per 64 columns of K instead of 64 KB at BN = 128, and at BN = 256
// the same 64 KB for twice the MMA work.  MMAs are issued by the leader CTA (M = 256: each SM's tensor core works on its own
// 128 lanes of A and D), commits are multicast to both CTAs' barriers, the converters and epilogue warps of the second CTA
// arrive on the leader's barriers.  BN = 256 leaves TMEM room for ONE accumulator (256 + 3 x 64 operand columns): the
// epilogue of a tile is then exposed, but TMA and the converters run ahead into the next tile's stages meanwhile.
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GTH, 1) tm_gemm_h2_kernel(const __grid_constant__ GemmParams p) {
  static_assert(KC == 32 && (BN == 128 || BN == 256), "tile shape");
  extern __shared__ __align__(1024) unsigned char smem_raw2[];
  unsigned char* smem = smem_raw2 + ((1024u - (smem_u32(smem_raw2) & 1023u)) & 1023u);      // same offset in both CTAs
  constexpr uint32_t BH_BYTES = (uint32_t)(BN / 2) * 128u;      // this CTA's half of B hi (and of B lo)
  constexpr uint32_t STAGE = 2u * TILE_BYTES + 2u * BH_BYTES;
  constexpr int NS = BN == 128 ? 4 : 3;
  constexpr int NACC = BN == 128 ? 2 : 1;
  constexpr uint32_t TM_A = NACC * BN;
  static_assert(TM_A + NS * H_A_SLOT <= 512, "TMEM budget");
  Bars* bars = reinterpret_cast<Bars*>(smem + NS * STAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cta = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int ntn = (p.N + BN - 1) / BN, ntm = (p.M + 2 * BM - 1) / (2 * BM);
  const int ntiles = p.ntiles;
  const int nchunk = (p.K + KCH - 1) / KCH;
  auto coord = [&](int tile) {      // this CTA's 128 rows of the pair's tile
    TileCoord t;
    t.n0 = (tile % ntn) * BN;
    t.m0 = ((tile / ntn) % ntm) * (2 * BM) + (int)cta * BM;
    t.bz = tile / (ntn * ntm);
    t.split = 0;
    return t;
  };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->split[s], 16);      // leader only: 8 converter warps of each CTA
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], 16);  // leader only: 8 epilogue warps of each CTA
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.ta);
    tma_prefetch_desc(&p.tbh);
    tma_prefetch_desc(&p.tbl);
  }
  if (warp == 1) {
    tmem_alloc<2>(&bars->tmem, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tm = bars->tmem;

  if (warp == 0) {
    if (lane == 0) {
      TPROF_DECL();
      uint32_t ctr = 0;
      for (int tile = cluster; tile < ntiles; tile += nclusters) {
        const TileCoord t = coord(tile);
        const int zlo = p.a_batched ? t.bz % p.zdiv : 0, zhi = p.a_batched ? t.bz / p.zdiv : 0;
        const bool a_ok = t.m0 < p.M;      // the pair's second half may lie wholly past M: nothing to load, zeros to TMEM
        for (int c = 0; c < nchunk; ++c, ++ctr) {
          const uint32_t s = ctr % NS, ph = (ctr / NS) & 1u;
          TPROF_WAIT(0, mbar_wait(&bars->empty[s], ph ^ 1u));
          unsigned char* st = smem + s * STAGE;
          const bool second = c * KCH + 32 < p.K;
          mbar_expect_tx(&bars->full[s], (a_ok ? (second ? 2u : 1u) * TILE_BYTES : 0u) + 2u * BH_BYTES);
          if (a_ok) {
            tma_load_4d(st, &p.ta, c * KCH, t.m0, zlo, zhi, &bars->full[s]);
            if (second) tma_load_4d(st + TILE_BYTES, &p.ta, c * KCH + 32, t.m0, zlo, zhi, &bars->full[s]);
          }
          const int nrow = t.n0 + (int)cta * (BN / 2);
          tma_load_2d(st + 2 * TILE_BYTES, &p.tbh, c * KCH, nrow, &bars->full[s]);
          tma_load_2d(st + 2 * TILE_BYTES + BH_BYTES, &p.tbl, c * KCH, nrow, &bars->full[s]);
        }
      }
      TPROF_FLUSH(0, -1, 1, true);
    }
  } else if (warp == 1) {
    if (cta == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);
      TPROF_DECL();
      uint32_t ctr = 0, it = 0;
      for (int tile = cluster; tile < ntiles; tile += nclusters, ++it) {
        const uint32_t buf = it % NACC;
        TPROF_WAIT(1, mbar_wait(&bars->acc_empty[buf], ((it / NACC) & 1u) ^ 1u));      // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tm + buf * BN;
        for (int c = 0; c < nchunk; ++c, ++ctr) {
          const uint32_t s = ctr % NS, ph = (ctr / NS) & 1u;
          TPROF_WAIT(0, mbar_wait(&bars->split[s], ph));      // both CTAs: A operands in TMEM, B halves landed
          tc_fence_after();
          const uint32_t b_hi = smem_u32(smem + s * STAGE + 2 * TILE_BYTES), b_lo = b_hi + BH_BYTES;
          const uint32_t ta_hi = tm + TM_A + s * H_A_SLOT, ta_lo = ta_hi + 32;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < KCH / 16; ++k) {
              const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
              umma_ts<2>(tacc, ta_lo + k * 8, umma_desc_k_sw128(b_hi + k * 32), idesc, acc);
              umma_ts<2>(tacc, ta_hi + k * 8, umma_desc_k_sw128(b_lo + k * 32), idesc, 1u);
              umma_ts<2>(tacc, ta_hi + k * 8, umma_desc_k_sw128(b_hi + k * 32), idesc, 1u);
            }
            umma_commit_2sm(&bars->empty[s], 3);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(&bars->acc_full[buf], 3);
        __syncwarp();
      }
      TPROF_FLUSH(2, 3, 4, lane == 0);
    }
  } else if (warp < H_EPI_WARP0) {
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const int h = (warp - 2) >> 2;      // which 32-column half of the stage this warp converts
    TPROF_DECL();
    uint32_t ctr = 0;
    for (int tile = cluster; tile < ntiles; tile += nclusters) {
      const bool a_ok = coord(tile).m0 < p.M;
      for (int c = 0; c < nchunk; ++c, ++ctr) {
        const uint32_t s = ctr % NS, ph = (ctr / NS) & 1u;
        TPROF_WAIT(0, mbar_wait(&bars->full[s], ph));
        const uint32_t ta = tm + lane_addr + TM_A + s * H_A_SLOT;
        {
          const unsigned char* rowp = smem + s * STAGE + h * TILE_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
          const bool landed = a_ok && (h == 0 || c * KCH + 32 < p.K);
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = landed ? *reinterpret_cast<const float4*>(rowp + ((i ^ (r & 7)) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            split_h2(v.x, v.y, hi[2 * i], lo[2 * i]);
            split_h2(v.z, v.w, hi[2 * i + 1], lo[2 * i + 1]);
          }
          tmem_st16(ta + 16 * h, hi);
          tmem_st16(ta + 32 + 16 * h, lo);
        }
        TPROF_WAIT(1, tmem_wait_st());
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bars->split[s], 0);
      }
    }
    TPROF_FLUSH(5, 9, 6, warp == 2 && lane == 0);
  } else {
    const int q = warp & 3, half = (warp - H_EPI_WARP0) >> 2;
    const float alpha = p.alpha * (p.bscale ? __ldg(p.bscale) : 1.f);
    TPROF_DECL();
    uint32_t it = 0;
    for (int tile = cluster; tile < ntiles; tile += nclusters, ++it) {
      const TileCoord t = coord(tile);
      const uint32_t buf = it % NACC;
      TPROF_WAIT(0, mbar_wait(&bars->acc_full[buf], (it / NACC) & 1u));
      tc_fence_after();
      epilogue_tile<BN>(p, t, nchunk, tm + buf * BN, q, half, lane, alpha);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&bars->acc_empty[buf], 0);
    }
    TPROF_FLUSH(7, -1, 8, warp == H_EPI_WARP0 && lane == 0);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 1) tmem_dealloc<2>(tm, 512);
}

// ---- pre-split image of a B operand: [256-byte header][hi: rows x ld16 fp16][lo: rows x ld16 fp16], ld16 = K rounded up to 8
struct SplitHeader {
  float inv_scale, scale;
  unsigned max_bits;             // bit pattern of max |b| (non-negative floats order like unsigned integers)
  int rows, k;
};
constexpr size_t SPLIT_HDR = 256;
inline int64_t split_ld(int k) { return ((int64_t)k + 7) / 8 * 8; }
inline size_t split_section(int rows, int k) { return ((size_t)rows * (size_t)split_ld(k) * 2 + 255) / 256 * 256; }

__global__ void __launch_bounds__(256) tm_split_max_kernel(const float* __restrict__ b, int rows, int k, long long ld, SplitHeader* h) {
  float mx = 0.f;
  const size_t total = (size_t)rows * k;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256)
    mx = fmaxf(mx, fabsf(b[(e / k) * ld + e % k]));
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(&h->max_bits, __float_as_uint(mx));
}

__global__ void __launch_bounds__(256) tm_split_write_kernel(const float* __restrict__ b, int rows, int k, long long ld, int ld16,
                                                             SplitHeader* h, uint32_t* __restrict__ hi, uint32_t* __restrict__ lo) {
  // power-of-two scale that puts max |b| into [2^13, 2^14): the lo parts (<= 2^-11 of their value) of every element within
  // 2^-13 of the maximum stay normal fp16 numbers, and the scaling itself is exact
  const float mx = __uint_as_float(h->max_bits);
  float scale = 1.f;
  if (mx > 0.f && mx < 3.0e38f) {
    int e;
    frexpf(mx, &e);
    scale = ldexpf(1.f, max(-100, min(100, 14 - e)));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    h->scale = scale;
    h->inv_scale = 1.f / scale;
    h->rows = rows;
    h->k = k;
  }
  const int pr = ld16 / 2;
  const size_t total = (size_t)rows * pr;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const size_t row = e / pr;
    const int col = (int)(e % pr) * 2;
    const float v0 = col < k ? b[row * ld + col] * scale : 0.f;
    const float v1 = col + 1 < k ? b[row * ld + col + 1] * scale : 0.f;
    uint32_t a, c;
    split_h2(v0, v1, a, c);
    hi[e] = a;
    lo[e] = c;
  }
}

// k-split: sum the partial tiles in a fixed order, then the same epilogue terms
__global__ void __launch_bounds__(256) tm_gemm_split_reduce_kernel(const __grid_constant__ GemmParams p, int batch) {
  const size_t total = (size_t)batch * p.M * p.N;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int col = (int)(e % p.N);
    const int row = (int)((e / p.N) % p.M);
    const int bz = (int)(e / ((size_t)p.M * p.N));
    float acc = 0.f;
    for (int s = 0; s < p.ksplit; ++s) acc += p.split_ws[(((size_t)bz * p.ksplit + s) * p.M + row) * p.N + col];
    float t = p.alpha * acc;
    if (p.diag != 0.f && col == row) t += p.diag;
    if (p.bias) t += p.bias[p.bias_row ? row : col];
    const size_t zlo = (size_t)(bz % p.zdiv), zhi = (size_t)(bz / p.zdiv);
    if (p.addend) t = fmaf(p.beta, p.addend[zlo * p.add_batch_stride + zhi * p.add_bs2 + (size_t)row * p.ld_add + col], t);
    t = act_apply_tm(t, p.act);
    if (p.c) p.c[zlo * p.c_batch_stride + zhi * p.c_bs2 + (size_t)(col / p.cbw) * p.cbs + (size_t)row * p.ldc + col % p.cbw] = t;
    if (p.ct) p.ct[zlo * p.ct_batch_stride + zhi * p.ct_bs2 + (size_t)col * p.ldct + row] = t;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn tm_get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeFn)ptr;
  }
  return fn;
}

int make_map(CUtensorMap* m, const float* base, int rows, int K, long long ld, int n_lo, long long stride_lo, int n_hi,
             long long stride_hi, int box_rows, const char* what) {
  EncodeFn enc = tm_get_encode();
  ACMIL_REQUIRE(enc != nullptr, ACMIL_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  ACMIL_REQUIRE(((uintptr_t)base & 15) == 0 && ld % 4 == 0 && (n_lo <= 1 || (stride_lo % 4 == 0 && stride_lo > 0)) &&
                    (n_hi <= 1 || (stride_hi % 4 == 0 && stride_hi > 0)),
                ACMIL_E_INVALID,
                "gemm: operand %s must be 16-byte aligned with leading dimension and batch strides that are positive multiples of 4 floats",
                what);
  const long long dummy = (long long)rows * ld;
  cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(n_lo < 1 ? 1 : n_lo), (cuuint64_t)(n_hi < 1 ? 1 : n_hi)};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)(n_lo > 1 ? stride_lo : dummy) * 4,
                           (cuuint64_t)(n_hi > 1 ? stride_hi : dummy) * 4};
  cuuint32_t box[4] = {KC, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, KC == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ACMIL_REQUIRE(r == CUDA_SUCCESS, ACMIL_E_CUDA, "cuTensorMapEncodeTiled failed for %s (%d)", what, (int)r);
  return ACMIL_OK;
}

template <int BN, bool PRECISE>
int launch(const GemmParams& gp, int batch, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = (size_t)ring_stages<BN>(PRECISE) * stage_bytes<BN>(PRECISE) + sizeof(Bars) + 1024;
  if (!configured) {
    ACMIL_CHECK_CUDA(cudaFuncSetAttribute(tm_gemm_kernel<BN, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    ACMIL_CHECK_CUDA(cudaGetDevice(&dev));
    ACMIL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  GemmParams gq = gp;
  const long long tiles = (long long)((gp.N + BN - 1) / BN) * ((gp.M + BM - 1) / BM) * batch * (gp.ksplit > 1 ? gp.ksplit : 1);
  ACMIL_REQUIRE(tiles < (1ll << 31), ACMIL_E_INVALID, "gemm: too many tiles");
  gq.ntiles = (int)tiles;
  tm_gemm_kernel<BN, PRECISE><<<(unsigned)std::min<long long>(tiles, n_sm), GT, smem, st>>>(gq);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  if (gp.ksplit > 1) {
    const size_t total = (size_t)batch * gp.M * gp.N;
    tm_gemm_split_reduce_kernel<<<(unsigned)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256), 256, 0, st>>>(gp, batch);
    ++g_acmil_launches;
    ACMIL_CHECK_CUDA(cudaGetLastError());
  }
  return ACMIL_OK;
}

int make_map_h(CUtensorMap* m, const void* base, int rows, int k, int64_t ld16, int box_rows, const char* what) {
  EncodeFn enc = tm_get_encode();
  ACMIL_REQUIRE(enc != nullptr, ACMIL_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  ACMIL_REQUIRE(((uintptr_t)base & 15) == 0 && ld16 % 8 == 0, ACMIL_E_INVALID, "gemm: split image section %s is not 16-byte aligned", what);
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld16 * 2};
  cuuint32_t box[2] = {(cuuint32_t)KCH, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ACMIL_REQUIRE(r == CUDA_SUCCESS, ACMIL_E_CUDA, "cuTensorMapEncodeTiled failed for %s (%d)", what, (int)r);
  return ACMIL_OK;
}

template <int BN>
int launch_h(const GemmParams& gp, int batch, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = (size_t)h_stages<BN>() * h_stage_bytes<BN>() + sizeof(Bars) + 1024;
  if (!configured) {
    ACMIL_CHECK_CUDA(cudaFuncSetAttribute(tm_gemm_h_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int dev = 0, n_sm = 0;
  ACMIL_CHECK_CUDA(cudaGetDevice(&dev));
  ACMIL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  GemmParams gq = gp;
  const long long tiles = (long long)((gp.N + BN - 1) / BN) * ((gp.M + BM - 1) / BM) * batch;
  ACMIL_REQUIRE(tiles < (1ll << 31), ACMIL_E_INVALID, "gemm: too many tiles");
  gq.ntiles = (int)tiles;
  tm_gemm_h_kernel<BN><<<(unsigned)std::min<long long>(tiles, n_sm), GTH, smem, st>>>(gq);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

template <int BN>
int launch_h2(const GemmParams& gp, int batch, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = (size_t)(BN == 128 ? 4 : 3) * (2u * TILE_BYTES + 2u * (uint32_t)(BN / 2) * 128u) + sizeof(Bars) + 1024;
  if (!configured) {
    ACMIL_CHECK_CUDA(cudaFuncSetAttribute(tm_gemm_h2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int dev = 0, n_sm = 0;
  ACMIL_CHECK_CUDA(cudaGetDevice(&dev));
  ACMIL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  GemmParams gq = gp;
  const long long tiles = (long long)((gp.N + BN - 1) / BN) * ((gp.M + 2 * BM - 1) / (2 * BM)) * batch;
  ACMIL_REQUIRE(tiles < (1ll << 31), ACMIL_E_INVALID, "gemm: too many tiles");
  gq.ntiles = (int)tiles;
  tm_gemm_h2_kernel<BN><<<2u * (unsigned)std::min<long long>(tiles, n_sm / 2), GTH, smem, st>>>(gq);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

// which tiling a pre-split weight product gets: CTA pairs with 128-wide tiles when every half of B starts inside the matrix
// and there are enough pair tiles to fill the GPU (measured on the ViT / TransMIL / ResNet weight products: 4 - 9 % faster
// than one CTA per tile; 256-wide pair tiles, with their single accumulator, are slower than both on K = 384 and only
// opt-in).  ACMIL_GEMM_PAIR = 0: never, 1 (default): 128-wide pairs, 2: 256-wide pairs where they fit
int h_tiling(int m, int n, int batch) {
  static const int allow = [] {
    const char* e = getenv("ACMIL_GEMM_PAIR");
    return e != nullptr && *e == '0' ? 0 : (e != nullptr && *e == '2' ? 2 : 1);
  }();
  if (allow == 0) return 0;
  int dev = 0, n_sm = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const long long mt = (m + 2 * BM - 1) / (2 * BM);
  auto fits = [&](int bn) {
    const int rem = n % bn;
    return n >= bn && (rem == 0 || rem > bn / 2) && (long long)((n + bn - 1) / bn) * mt * batch >= n_sm / 2;
  };
  if (allow == 2 && fits(256)) return 256;
  if (fits(128)) return 128;
  return 0;
}

}  // namespace

int tm_gemm(const acmil_gemm_desc& d, cudaStream_t st) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d.m > 0 && d.n > 0 && d.k > 0 && d.batch > 0, ACMIL_E_INVALID, "gemm: empty problem (%d x %d x %d, batch %d)", d.m,
                d.n, d.k, d.batch);
  ACMIL_REQUIRE(d.a && (d.b || d.b_split) && (d.c || d.ct), ACMIL_E_INVALID, "gemm: null operand");
  const int ksplit = d.k_split > 1 ? d.k_split : 1;
  ACMIL_REQUIRE(ksplit == 1 || d.split_ws != nullptr, ACMIL_E_INVALID, "gemm: k_split needs split_ws");
  GemmParams gp{};
  // 64-wide tiles when the problem is at most 64 columns wide, or when 128-wide tiles would leave more than half of the SMs
  // idle (the 256^3 products of the pseudo-inverse iteration: 32 tiles -> 64)
  int bn = d.n <= 64 ? 64 : 128;
  if (bn == 128) {
    static int n_sm_cached = 0;
    if (n_sm_cached == 0) {
      int dev = 0;
      ACMIL_CHECK_CUDA(cudaGetDevice(&dev));
      ACMIL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm_cached, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long t128 = (long long)((d.n + 127) / 128) * ((d.m + BM - 1) / BM) * d.batch * ksplit;
    if (2 * t128 <= n_sm_cached) bn = 64;
  }
  const int zdiv = d.batch_inner > 0 ? d.batch_inner : d.batch;
  ACMIL_REQUIRE(d.batch % zdiv == 0, ACMIL_E_INVALID, "gemm: batch %d is not a multiple of batch_inner %d", d.batch, zdiv);
  const int n_hi = d.batch / zdiv;
  const bool a_b = d.a_batch_stride != 0 || d.a_batch_stride2 != 0, b_b = d.b_batch_stride != 0 || d.b_batch_stride2 != 0;
  ACMIL_REQUIRE((!a_b || ((zdiv == 1 || d.a_batch_stride) && (n_hi == 1 || d.a_batch_stride2))) &&
                    (!b_b || ((zdiv == 1 || d.b_batch_stride) && (n_hi == 1 || d.b_batch_stride2))),
                ACMIL_E_INVALID, "gemm: a batched operand needs a non-zero stride on every batch level in use");
  // precise = 2 with a pre-split image of B (acmil_gemm_split_b): the fp16-split kernel; without an image it means precise = 1
  const bool hmode = d.precise == 2 && d.b_split != nullptr;
  int pair_bn = 0;      // fp16-split mode: 0 = one CTA per tile, 128 / 256 = CTA pairs with that tile width
  if (hmode) {
    ACMIL_REQUIRE(!b_b && ksplit == 1, ACMIL_E_INVALID, "gemm: a pre-split B operand cannot be batched or k-split");
    ACMIL_REQUIRE(d.b_split_rows >= 1 && d.b_split_row0 >= 0 && (int64_t)d.b_split_row0 + d.n <= d.b_split_rows, ACMIL_E_INVALID,
                  "gemm: rows [%d, %d) are outside the split image (%d rows)", d.b_split_row0, d.b_split_row0 + d.n, d.b_split_rows);
    ACMIL_REQUIRE(((uintptr_t)d.b_split & 255) == 0, ACMIL_E_INVALID, "gemm: the split image must be 256-byte aligned");
  }
  int rc = make_map(&gp.ta, d.a, d.m, d.k, d.lda, a_b ? zdiv : 1, d.a_batch_stride, a_b ? n_hi : 1, d.a_batch_stride2, BM, "A");
  if (rc) return rc;
  if (hmode) {
    const int64_t ld16 = split_ld(d.k);
    const unsigned char* img = reinterpret_cast<const unsigned char*>(d.b_split);
    const unsigned char* hi = img + SPLIT_HDR + (size_t)d.b_split_row0 * ld16 * 2;
    pair_bn = h_tiling(d.m, d.n, d.batch);
    const int box_rows = pair_bn ? pair_bn / 2 : bn;
    rc = make_map_h(&gp.tbh, hi, d.n, d.k, ld16, box_rows, "B hi");
    if (rc) return rc;
    rc = make_map_h(&gp.tbl, hi + split_section(d.b_split_rows, d.k), d.n, d.k, ld16, box_rows, "B lo");
    if (rc) return rc;
    gp.bscale = reinterpret_cast<const float*>(img);      // SplitHeader::inv_scale
  } else {
    ACMIL_REQUIRE(d.b != nullptr, ACMIL_E_INVALID, "gemm: null operand");
    rc = make_map(&gp.tb, d.b, d.n, d.k, d.ldb, b_b ? zdiv : 1, d.b_batch_stride, b_b ? n_hi : 1, d.b_batch_stride2, bn, "B");
    if (rc) return rc;
  }
  gp.zdiv = zdiv;
  gp.c_bs2 = d.c_batch_stride2; gp.ct_bs2 = d.ct_batch_stride2; gp.add_bs2 = d.addend_batch_stride2;
  gp.c = d.c; gp.ct = d.ct; gp.bias = d.bias; gp.addend = d.addend;
  gp.M = d.m; gp.N = d.n; gp.K = d.k;
  gp.a_batched = a_b; gp.b_batched = b_b;
  gp.c_batch_stride = d.c_batch_stride; gp.ldc = d.ldc;
  gp.cbw = d.col_block_width > 0 ? d.col_block_width : (1 << 30);      // no column blocks: one block holds every column
  gp.cbs = d.col_block_width > 0 ? d.col_block_stride : 0;
  gp.ct_batch_stride = d.ct_batch_stride; gp.ldct = d.ldct;
  gp.add_batch_stride = d.addend_batch_stride; gp.ld_add = d.ld_addend;
  gp.alpha = d.alpha; gp.beta = d.beta; gp.diag = d.diag; gp.act = d.act; gp.bias_row = d.bias_per_row;
  gp.ksplit = ksplit;
  gp.k_per_split = ((((d.k + KC - 1) / KC) + ksplit - 1) / ksplit) * KC;
  gp.split_ws = d.split_ws;

  if (d.softmax_stats_out) {
    ACMIL_REQUIRE(ksplit == 1 && d.act == 0 && d.c != nullptr && d.ct == nullptr, ACMIL_E_INVALID,
                  "gemm: softmax_stats_out needs a plain C output (no k_split, activation or transposed copy)");
    gp.stats_out = reinterpret_cast<float2*>(d.softmax_stats_out);
    gp.nch_out = (d.n + 31) / 32;
  }
  if (d.softmax_stats_in) {
    ACMIL_REQUIRE(d.precise != 0 && !hmode && KC == 32, ACMIL_E_INVALID, "gemm: softmax_stats_in needs the 3xTF32 kernel (precise = 1)");
    ACMIL_REQUIRE(((uintptr_t)d.softmax_stats_in & 7) == 0, ACMIL_E_INVALID, "gemm: softmax_stats_in must be 8-byte aligned");
    gp.stats_in = reinterpret_cast<const float2*>(d.softmax_stats_in);
    gp.nch_in = (d.k + 31) / 32;
  }
  gp.vec_ok = d.c != nullptr && ((uintptr_t)d.c & 15) == 0 && d.ldc % 4 == 0 && gp.cbw % 4 == 0 && gp.cbs % 4 == 0 &&
              d.c_batch_stride % 4 == 0 && d.c_batch_stride2 % 4 == 0;
  gp.add_vec_ok = d.addend != nullptr && ((uintptr_t)d.addend & 15) == 0 && d.ld_addend % 4 == 0 &&
                  d.addend_batch_stride % 4 == 0 && d.addend_batch_stride2 % 4 == 0;
  gp.bias_vec_ok = d.bias != nullptr && !d.bias_per_row && ((uintptr_t)d.bias & 15) == 0;
  if (hmode && pair_bn == 256) return launch_h2<256>(gp, d.batch, st);
  if (hmode && pair_bn == 128) return launch_h2<128>(gp, d.batch, st);
  if (hmode) return bn == 64 ? launch_h<64>(gp, d.batch, st) : launch_h<128>(gp, d.batch, st);
  if (d.precise) return bn == 64 ? launch<64, true>(gp, d.batch, st) : launch<128, true>(gp, d.batch, st);
  return bn == 64 ? launch<64, false>(gp, d.batch, st) : launch<128, false>(gp, d.batch, st);
}

extern "C" int acmil_gemm_nt(const acmil_gemm_desc* desc, void* stream) {
  ACMIL_REQUIRE(desc != nullptr, ACMIL_E_INVALID, "gemm: null descriptor");
  return tm_gemm(*desc, (cudaStream_t)stream);
}

extern "C" int acmil_gemm_split_bytes(int32_t rows, int32_t k, size_t* bytes) {
  ACMIL_REQUIRE(bytes != nullptr && rows >= 1 && k >= 1, ACMIL_E_INVALID, "gemm split: bad shape [%d, %d]", rows, k);
  *bytes = SPLIT_HDR + 2 * split_section(rows, k);
  return ACMIL_OK;
}

extern "C" int acmil_gemm_split_b(const float* d_b, int32_t rows, int32_t k, int64_t ldb, void* d_image, size_t image_bytes, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_b && d_image && rows >= 1 && k >= 1 && ldb >= k, ACMIL_E_INVALID, "gemm split: bad arguments");
  ACMIL_REQUIRE(((uintptr_t)d_image & 255) == 0, ACMIL_E_INVALID, "gemm split: the image must be 256-byte aligned");
  const size_t need = SPLIT_HDR + 2 * split_section(rows, k);
  ACMIL_REQUIRE(image_bytes >= need, ACMIL_E_WORKSPACE, "gemm split: image %zu < %zu bytes", image_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* img = reinterpret_cast<unsigned char*>(d_image);
  SplitHeader* h = reinterpret_cast<SplitHeader*>(img);
  ACMIL_CHECK_CUDA(cudaMemsetAsync(img, 0, SPLIT_HDR, st));
  const size_t total = (size_t)rows * k;
  const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 8);
  tm_split_max_kernel<<<grid, 256, 0, st>>>(d_b, rows, k, ldb, h);
  tm_split_write_kernel<<<grid, 256, 0, st>>>(d_b, rows, k, ldb, (int)split_ld(k), h, reinterpret_cast<uint32_t*>(img + SPLIT_HDR),
                                              reinterpret_cast<uint32_t*>(img + SPLIT_HDR + split_section(rows, k)));
  g_acmil_launches += 2;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}
