// Fused gated-attention pool row pass on the 5th-gen tensor cores (sm_100a): TMA -> smem -> fp16 hi/lo
// split -> TMEM operands -> tcgen05.mma (cta_group::2) -> TMEM accumulators -> gate / softmax / pool
// epilogue, one persistent CTA pair per two SMs.
//
// Why it looks like this (DESIGN.md has the long version):
//  * The path is HBM-bound only if the three projections run on tensor cores (106 FLOP/B), and the
//    reference is IEEE fp32, so every GEMM is done as an error-compensated fp16 split:
//        x W^T ~= x_hi W_hi^T + x_lo W_hi^T + x_hi W_lo^T     (x = x_hi + x_lo, 11 + 11 bits; fp32 accumulate)
//    which keeps ~2^-21 relative accuracy at 3 fp16 MMAs per product.
//  * hi/lo weight images need 4 B per weight = 320 KB for D_feat 384: more than one SM's shared memory.
//    A CTA pair (cta_group::2) splits the B operand (the weights) by output column, so each SM keeps
//    160 KB resident for the whole launch and nothing but x is streamed.
//  * A operands (converted x, and h = relu(xW1^T)) live in TMEM, written with tcgen05.st by the thread
//    that owns the row, so the MMAs read only B from shared memory.
//  * TMEM (512 columns): D1 h-accumulator 128 | D2 gate accumulator 128 (two halves per tile) |
//    h operand hi/lo 128 | x operand ring 4 x 32.
//
// Roles per CTA (384 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM
// allocator, warps 4-7 converters (thread = row), warps 8-11 epilogue (thread = row; each warp keeps a
// private online-softmax stream and candidate lists, so there is no cross-warp traffic per tile).
#include <cuda.h>
#include <cuda_fp16.h>

#include "gp_common.cuh"
#include "sm100.cuh"

namespace {
using namespace sm100;

constexpr int UT = 384;
constexpr int KC = 32;       // x columns per chunk (one 128-byte swizzle span of fp32)
constexpr int NSTAGE = 3;    // fp32 staging ring (TMA destination)
constexpr int NXOP = 4;      // TMEM x-operand ring
constexpr int STAGE_BYTES = 128 * KC * 4;
constexpr uint32_t TM_D1 = 0, TM_D2 = 128, TM_HHI = 256, TM_HLO = 320, TM_X = 384;
constexpr float LOG2E = 1.4426950408889634f;

#ifndef GP_UMMA_PROF
#define GP_UMMA_PROF 0
#endif
#if GP_UMMA_PROF
__device__ long long g_umma_prof[148][32];
#define PROF_T0() const long long _t0 = clock64()
#define PROF_ADD(slot) prof[slot] += clock64() - _t0
#define PROF_DECL() long long prof[16] = {0}
#define PROF_FLUSH(base) do { for (int _i = 0; _i < 8; ++_i) g_umma_prof[blockIdx.x][(base) + _i] = prof[_i]; } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF_ADD(slot) do {} while (0)
#define PROF_DECL() do {} while (0)
#define PROF_FLUSH(base) do {} while (0)
#endif

struct UmmaConsts {
  float b1[128], bv[128], bu[128], ww[KMAX][128], bw[KMAX];
  float inv_s1, inv_sv, inv_su;
};

struct UmmaParams {
  GpMainParams mp;
  UmmaConsts c;
  CUtensorMap tmap;
  const unsigned char* wimg;   // per-CTA weight images, cta_img_bytes each
  uint32_t cta_img_bytes;
  uint32_t w1_part_bytes;      // bytes of one (hi or lo) W1 half image
};

// smem carve-up (offsets from the 1024-aligned base)
struct SmemMap {
  uint32_t w1, wg, stage, tbuf, ps, cst, bars, total;
};
// per-unit record of gate constants in smem: {ww[0..KB-1], bv', bu'} padded to CREC floats, where
// bv' = -2 log2e bv and bu' = -log2e bu are the biases in the exponent domain
__host__ __device__ constexpr int cst_rec(int kb) { return kb <= 2 ? 4 : (kb <= 6 ? 8 : 10); }
__host__ __device__ inline SmemMap smem_map(int din) {
  SmemMap m;
  m.w1 = 0;
  m.wg = (uint32_t)(din / 64) * 8192u * 2u;
  m.stage = m.wg + 65536u;
  m.tbuf = m.stage + NSTAGE * STAGE_BYTES;
  m.ps = m.tbuf + 4 * 2048;
  m.cst = m.ps + 4 * 1024;
  m.bars = m.cst + 128 * 10 * 4;
  m.total = m.bars + 256;
  return m;
}

struct Bars {
  uint64_t full_x[NSTAGE], empty_x[NSTAGE];
  uint64_t xop_full[NXOP], xop_empty[NXOP];
  uint64_t d1_full, d1_empty, hop_full, d2_full, d2_empty, wload, w_ready;
  uint32_t tmem_base;
};

// position of global pair-tile g: bag s, first row inside the bag for this CTA
struct TilePos {
  int s;
  int64_t row_in_bag;
};

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// fp32 -> (fp16 hi, fp16 lo) with hi = top 11 significant bits (exact in fp16 for |v| in [2^-14, 65504])
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  const float bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  hi = pack_half2(ah, bh);
  lo = pack_half2(a - ah, b - bh);
}

template <int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UT, 1) gp_main_umma_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need a 1024-byte aligned base (same offset in both CTAs of the pair)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int DIN = p.mp.sh.d_in;
  const int NCH = DIN / KC;
  const SmemMap sm = smem_map(DIN);
  Bars* bars = reinterpret_cast<Bars*>(smem + sm.bars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cta = cluster_ctarank();
  const int cluster = blockIdx.x >> 1;
  const GpSegTable& seg = p.mp.seg;
  const int K = p.mp.sh.n_branch;

  // this cluster's run of global pair-tiles
  const int g0 = (int)(((uint32_t)cluster * (uint32_t)seg.u_total_pt) / (uint32_t)seg.u_nclusters);
  const int g1 = (int)(((uint32_t)(cluster + 1) * (uint32_t)seg.u_total_pt) / (uint32_t)seg.u_nclusters);
  const int T = g1 - g0;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars->full_x[i], 1); mbar_init(&bars->empty_x[i], 4); }
    for (int i = 0; i < NXOP; ++i) { mbar_init(&bars->xop_full[i], 8); mbar_init(&bars->xop_empty[i], 1); }
    mbar_init(&bars->d1_full, 1);
    mbar_init(&bars->d1_empty, 8);
    mbar_init(&bars->hop_full, 8);
    mbar_init(&bars->d2_full, 1);
    mbar_init(&bars->d2_empty, 8);
    mbar_init(&bars->wload, 1);
    mbar_init(&bars->w_ready, 2);
    fence_mbar_init();
    // resident weight images: one bulk copy per 16 KB
    const unsigned char* src = p.wimg + (size_t)cta * p.cta_img_bytes;
    mbar_expect_tx(&bars->wload, p.cta_img_bytes);
    for (uint32_t off = 0; off < p.cta_img_bytes; off += 16384) bulk_load(smem + off, src + off, 16384, &bars->wload);
    tma_prefetch_desc(&p.tmap);
  }
  {
    constexpr int CREC = cst_rec(KB);
    float* cst = reinterpret_cast<float*>(smem + sm.cst);
    for (int u = tid; u < 128; u += UT) {
#pragma unroll
      for (int k = 0; k < KB; ++k) cst[u * CREC + k] = p.c.ww[k][u];
      cst[u * CREC + KB] = p.c.bv[u] * (-2.f * LOG2E);
      cst[u * CREC + KB + 1] = p.c.bu[u] * (-LOG2E);
    }
  }
  if (warp == 2) {
    tmem_alloc<2>(&bars->tmem_base, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tm = bars->tmem_base;

  auto tile_pos = [&](int g, int& s_hint) -> TilePos {
    while (g >= seg.u_pt_begin[s_hint + 1]) ++s_hint;
    TilePos t;
    t.s = s_hint;
    t.row_in_bag = (int64_t)(g - seg.u_pt_begin[s_hint]) * 256 + (int64_t)cta * 128;
    return t;
  };

  // NOTE: each setmaxnreg sits at the top of a branch that never rejoins the others before the kernel's
  // tail, otherwise ptxas allocates the whole kernel for the smallest budget.
  if (warp < 4) {
  setmaxnreg_dec<64>();
  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      PROF_DECL();
      int s_hint = 0;
      uint32_t ctr = 0;
      for (int g = g0; g < g1; ++g) {
        const TilePos tp = tile_pos(g, s_hint);
        const int64_t grow = seg.row_off[tp.s] + tp.row_in_bag;
        for (int c = 0; c < NCH; ++c, ++ctr) {
          const uint32_t st = ctr % NSTAGE, ph = (ctr / NSTAGE) & 1u;
          { PROF_T0(); mbar_wait(&bars->empty_x[st], ph ^ 1u); PROF_ADD(0); }
          mbar_expect_tx(&bars->full_x[st], STAGE_BYTES);
          tma_load_2d_hint(smem + sm.stage + st * STAGE_BYTES, &p.tmap, c * KC, (int)grow, &bars->full_x[st], kEvictFirst);
        }
      }
      PROF_FLUSH(0);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA) =====================================
    if (cta == 0 && lane == 0 && T > 0) {
      PROF_DECL();
      const long long t_start = clock64();
      { PROF_T0(); mbar_wait_cluster(&bars->w_ready, 0); PROF_ADD(0); }  // both CTAs' weight images have landed
      const uint32_t idesc = umma_idesc_f16(256, 128);
      const uint32_t w1_hi = smem_u32(smem + sm.w1), w1_lo = w1_hi + p.w1_part_bytes;
      const uint32_t wg_hi = smem_u32(smem + sm.wg), wg_lo = wg_hi + 32768u;
      uint32_t xc = 0;  // x-operand chunks consumed so far (ring position / phase)
      auto g1_chunks = [&](int t, int c_begin, int c_end) {
        for (int c = c_begin; c < c_end; ++c, ++xc) {
          if (c == 0 && t > 0) {  // D1 of the previous tile must have been drained by the epilogue
            { PROF_T0(); mbar_wait_cluster(&bars->d1_empty, (uint32_t)(t - 1) & 1u); PROF_ADD(1); }
            tc_fence_after();
          }
          const uint32_t q = xc % NXOP, ph = (xc / NXOP) & 1u;
          { PROF_T0(); mbar_wait_cluster(&bars->xop_full[q], ph); PROF_ADD(2); }
          tc_fence_after();
          const uint32_t xa_hi = tm + TM_X + q * 32, xa_lo = xa_hi + 16;
          const uint32_t boff = (uint32_t)(c >> 1) * 8192u + (uint32_t)(c & 1) * 64u;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t bhi = umma_desc_k_sw128(w1_hi + boff + ks * 32), blo = umma_desc_k_sw128(w1_lo + boff + ks * 32);
            umma_ts<2>(tm + TM_D1, xa_hi + ks * 8, bhi, idesc, (c | ks) ? 1u : 0u);
            umma_ts<2>(tm + TM_D1, xa_lo + ks * 8, bhi, idesc, 1u);
            umma_ts<2>(tm + TM_D1, xa_hi + ks * 8, blo, idesc, 1u);
          }
          umma_commit_2sm(&bars->xop_empty[q], 3);
          if (c == NCH - 1) umma_commit_2sm(&bars->d1_full, 3);
        }
      };
      auto g2_half = [&](int t, int h) {
        if (h == 0) { PROF_T0(); mbar_wait_cluster(&bars->hop_full, (uint32_t)t & 1u); PROF_ADD(3); }
        const uint32_t u = 2u * (uint32_t)t + (uint32_t)h;
        if (u > 0) { PROF_T0(); mbar_wait_cluster(&bars->d2_empty, (u - 1u) & 1u); PROF_ADD(4); }
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t boff = (uint32_t)(h * 2 + (ks >> 2)) * 8192u + (uint32_t)(ks & 3) * 32u;
          const uint64_t bhi = umma_desc_k_sw128(wg_hi + boff), blo = umma_desc_k_sw128(wg_lo + boff);
          umma_ts<2>(tm + TM_D2, tm + TM_HHI + ks * 8, bhi, idesc, ks ? 1u : 0u);
          umma_ts<2>(tm + TM_D2, tm + TM_HLO + ks * 8, bhi, idesc, 1u);
          umma_ts<2>(tm + TM_D2, tm + TM_HHI + ks * 8, blo, idesc, 1u);
        }
        umma_commit_2sm(&bars->d2_full, 3);
      };
      g1_chunks(0, 0, NCH);
      for (int t = 0; t < T; ++t) {
        g2_half(t, 0);
        g2_half(t, 1);
        if (t + 1 < T) g1_chunks(t + 1, 0, NCH);
      }
#if GP_UMMA_PROF
      prof[7] = clock64() - t_start;
#endif
      PROF_FLUSH(8);
    }
    __syncwarp();
  } else if (warp == 3) {
    // this CTA's resident weights are in place -> tell the leader's MMA thread
    if (lane == 0) {
      mbar_wait(&bars->wload, 0);
      mbar_arrive_cluster(&bars->w_ready, 0);
    }
    __syncwarp();
  }
  } else if (warp < 8) {
    // ===================================== converters: fp32 staging -> fp16 hi/lo in TMEM =====================================
    setmaxnreg_dec<112>();
    const int r = (warp - 4) * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp - 4) * 32) << 16;
    uint32_t ctr = 0;
    PROF_DECL();
#if GP_UMMA_PROF
    const long long t_start = clock64();
#endif
    for (int t = 0; t < T; ++t) {
      for (int c = 0; c < NCH; ++c, ++ctr) {
        const uint32_t st = ctr % NSTAGE, ph = (ctr / NSTAGE) & 1u;
        const uint32_t q = ctr % NXOP, phq = (ctr / NXOP) & 1u;
        { PROF_T0(); mbar_wait(&bars->full_x[st], ph); PROF_ADD(0); }
        const uint8_t* rowp = smem + sm.stage + st * STAGE_BYTES + r * 128;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(rowp + ((i ^ (r & 7)) << 4));
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          split2(v[i].x, v[i].y, hi[2 * i], lo[2 * i]);
          split2(v[i].z, v[i].w, hi[2 * i + 1], lo[2 * i + 1]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->empty_x[st]);   // staging slot may be refilled
        { PROF_T0(); mbar_wait_cluster(&bars->xop_empty[q], phq ^ 1u); PROF_ADD(1); }
        tc_fence_after();
        tmem_st16(tm + lane_addr + TM_X + q * 32, hi);
        tmem_st16(tm + lane_addr + TM_X + q * 32 + 16, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bars->xop_full[q], 0);
      }
    }
#if GP_UMMA_PROF
    prof[7] = clock64() - t_start;
    if (warp == 4 && lane == 0) PROF_FLUSH(16);
#endif
  } else {
    // ===================================== epilogue: thread = row =====================================
    setmaxnreg_inc<240>();
    const int ew = warp - 8;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    float* tbuf = reinterpret_cast<float*>(smem + sm.tbuf + ew * 2048);   // [32 rows][16 feats], chunk-swizzled
    float* psw = reinterpret_cast<float*>(smem + sm.ps + ew * 1024);      // [32 rows][8]
    const int L = 128;
    const int cap = seg.n_masked_cap;
    const int jf = lane & 15, par = lane >> 4;

    float m_run[KB], l_run[KB], acc[8][KB];
    // candidate lists: lane i holds entry i of every branch
    float c_s[KB];
    int c_i[KB], c_sl[KB], c_cnt[KB];
    unsigned c_free[KB];
    int s_cur = -1, s_hint = 0, nm = 0, seg_id = 0;
    int64_t n_rows = 0;

    auto reset_stream = [&](int s) {
      s_cur = s;
      nm = seg.nm[s];
      n_rows = seg.row_off[s + 1] - seg.row_off[s];
      seg_id = seg.seg_begin[s] + (cluster - seg.u_cfirst[s]) * 8 + (int)cta * 4 + ew;
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        m_run[k] = -INFINITY;
        l_run[k] = 0.f;
        c_s[k] = -INFINITY;
        c_i[k] = 0x7fffffff;
        c_sl[k] = 0;
        c_cnt[k] = 0;
        c_free[k] = nm >= 32 ? 0xffffffffu : ((1u << nm) - 1u);
#pragma unroll
        for (int g = 0; g < 8; ++g) acc[g][k] = 0.f;
      }
    };
    auto flush_stream = [&]() {
      if (s_cur < 0) return;
      float* part = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.part) + (size_t)seg_id * K * (L + 2);
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        if (k < K) {
          if (lane == 0) {
            part[(size_t)k * (L + 2) + 0] = m_run[k];
            part[(size_t)k * (L + 2) + 1] = l_run[k];
          }
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float v = acc[g][k] + __shfl_xor_sync(0xffffffffu, acc[g][k], 16);   // even + odd rows
            if (par == 0) part[(size_t)k * (L + 2) + 2 + g * 16 + jf] = v;
          }
        }
      }
      if (cap > 0) {
        int* g_cnt = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_cnt) + (size_t)seg_id * K;
        float* g_score = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_score) + (size_t)seg_id * K * cap;
        int* g_idx = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_idx) + (size_t)seg_id * K * cap;
        int* g_slot = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_slot) + (size_t)seg_id * K * cap;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          if (k < K) {
            if (lane == 0) g_cnt[k] = c_cnt[k];
            if (lane < cap) {
              const bool live = lane < c_cnt[k];
              g_score[k * cap + lane] = live ? c_s[k] : -INFINITY;
              g_idx[k * cap + lane] = live ? c_i[k] : 0x7fffffff;
              g_slot[k * cap + lane] = live ? c_sl[k] : 0;
            }
          }
        }
      }
    };

    PROF_DECL();
#if GP_UMMA_PROF
    const long long t_start = clock64();
#endif
    for (int t = 0; t < T; ++t) {
      const TilePos tp = tile_pos(g0 + t, s_hint);
      if (tp.s != s_cur) {
        flush_stream();
        reset_stream(tp.s);
      }
      const int64_t row_in_bag = tp.row_in_bag + ew * 32 + lane;
      const bool valid = row_in_bag < n_rows;

      // ---------------- Epi1: D1 -> relu -> fp16 hi/lo operand of the gate GEMM ----------------
      { PROF_T0(); mbar_wait_cluster(&bars->d1_full, (uint32_t)t & 1u); PROF_ADD(0); }
      tc_fence_after();
#if GP_UMMA_PROF
      const long long t_e1 = clock64();
#endif
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t v[32];
        tmem_ld32(tm + lane_addr + TM_D1 + g * 32, v);
        tmem_wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = fmaxf(fmaf(__uint_as_float(v[2 * i]), p.c.inv_s1, p.c.b1[g * 32 + 2 * i]), 0.f);
          const float b = fmaxf(fmaf(__uint_as_float(v[2 * i + 1]), p.c.inv_s1, p.c.b1[g * 32 + 2 * i + 1]), 0.f);
          split2(a, b, hi[i], lo[i]);
        }
        tmem_st16(tm + lane_addr + TM_HHI + g * 16, hi);
        tmem_st16(tm + lane_addr + TM_HLO + g * 16, lo);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(&bars->hop_full, 0);
        mbar_arrive_cluster(&bars->d1_empty, 0);
      }

#if GP_UMMA_PROF
      prof[2] += clock64() - t_e1;
      const long long t_e2 = clock64();
#endif
      // ---------------- Epi2: gate + scores ----------------
      float sc_[KB];
#pragma unroll
      for (int k = 0; k < KB; ++k) sc_[k] = p.c.bw[k];
      const float* cstp = reinterpret_cast<const float*>(smem + sm.cst);
      const float cva = p.c.inv_sv * (-2.f * LOG2E), cua = p.c.inv_su * (-LOG2E);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        { PROF_T0(); mbar_wait_cluster(&bars->d2_full, (uint32_t)(2 * t + h) & 1u); PROF_ADD(1); }
        tc_fence_after();
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t zv[32], zu[32];
          tmem_ld32(tm + lane_addr + TM_D2 + sub * 32, zv);
          tmem_ld32(tm + lane_addr + TM_D2 + 64 + sub * 32, zu);
          tmem_wait_ld();
          if (sub == 1) {   // this half of D2 is in registers: the MMA warp may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&bars->d2_empty, 0);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int u = h * 64 + sub * 32 + i;
            constexpr int CREC = cst_rec(KB);
            float cr[CREC];
            if constexpr (CREC == 10) {
#pragma unroll
              for (int j = 0; j < 5; ++j) *reinterpret_cast<float2*>(&cr[2 * j]) = *reinterpret_cast<const float2*>(cstp + u * CREC + 2 * j);
            } else {
#pragma unroll
              for (int j = 0; j < CREC / 4; ++j) *reinterpret_cast<float4*>(&cr[4 * j]) = *reinterpret_cast<const float4*>(cstp + u * CREC + 4 * j);
            }
            // tanh(a) * sigmoid(b) = (1 - Ea) / ((1 + Ea)(1 + Eb)),  Ea = e^-2a, Eb = e^-b  (exponents clamped at 40);
            // scales and biases are pre-multiplied into the exponent domain
            const float ea = ex2_approx(fminf(fmaf(__uint_as_float(zv[i]), cva, cr[KB]), 57.7f));
            const float eb = ex2_approx(fminf(fmaf(__uint_as_float(zu[i]), cua, cr[KB + 1]), 57.7f));
            const float gte = (1.f - ea) * rcp_approx((1.f + ea) * (1.f + eb));
#pragma unroll
            for (int k = 0; k < KB; ++k) sc_[k] = fmaf(gte, cr[k], sc_[k]);
          }
        }
      }

#if GP_UMMA_PROF
      prof[3] += clock64() - t_e2;
      const long long t_e3 = clock64();
#endif
      // ---------------- raw scores out ----------------
      if (p.mp.a_out != nullptr && valid) {
#pragma unroll
        for (int k = 0; k < KB; ++k)
          if (k < K) p.mp.a_out[(size_t)k * p.mp.a_ld + seg.row_off[s_cur] + row_in_bag] = sc_[k];
      }

      // ---------------- candidates (top-n rows per branch stay out of the sums) + online softmax ----------------
      float pk[KB], scale[KB];
      unsigned my_new = 0u;          // bit k: this lane's row entered branch k's list this tile
      int my_slot[KB];
      float ev_w[KB];                // lane e: weight of the e-th entry evicted from branch k this tile
      int ev_slot[KB], n_ev[KB];
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        n_ev[k] = 0;
        ev_w[k] = -INFINITY;
        ev_slot[k] = 0;
        my_slot[k] = 0;
        if (k < K && nm > 0) {
          int cnt = c_cnt[k];
          const float tau = cnt == nm ? __shfl_sync(0xffffffffu, c_s[k], nm - 1) : -INFINITY;
          unsigned bal = __ballot_sync(0xffffffffu, valid && (cnt < nm || sc_[k] > tau));
          unsigned freed = 0u;
          while (bal) {
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            const float s_new = __shfl_sync(0xffffffffu, sc_[k], src);
            if (cnt == nm) {
              const float last_s = __shfl_sync(0xffffffffu, c_s[k], nm - 1);
              if (!(s_new > last_s)) continue;
              const int last_sl = __shfl_sync(0xffffffffu, c_sl[k], nm - 1);
              if (last_sl >= 0) {
                if (lane == n_ev[k]) { ev_w[k] = last_s; ev_slot[k] = last_sl; }
                freed |= 1u << last_sl;
                ++n_ev[k];
              }
            }
            const int pos = __popc(__ballot_sync(0xffffffffu, lane < cnt && c_s[k] >= s_new));
            const float up_s = __shfl_up_sync(0xffffffffu, c_s[k], 1);
            const int up_i = __shfl_up_sync(0xffffffffu, c_i[k], 1);
            const int up_sl = __shfl_up_sync(0xffffffffu, c_sl[k], 1);
            if (lane > pos) { c_s[k] = up_s; c_i[k] = up_i; c_sl[k] = up_sl; }
            if (lane == pos) { c_s[k] = s_new; c_i[k] = (int)(tp.row_in_bag + ew * 32 + src); c_sl[k] = -1 - src; }
            if (cnt < nm) ++cnt;
          }
          const bool is_new = lane < cnt && c_sl[k] < 0;
          const unsigned newmask = __ballot_sync(0xffffffffu, is_new);
          const unsigned freemask = c_free[k] | freed;
          int row_of_new = -1, slot_of_new = 0;
          if (is_new) {
            unsigned fm = freemask;
            const int rank = __popc(newmask & ((1u << lane) - 1u));
            for (int i = 0; i < rank; ++i) fm &= fm - 1;
            slot_of_new = __ffs(fm) - 1;
            row_of_new = -1 - c_sl[k];
            c_sl[k] = slot_of_new;
          }
          // tell the lane that owns each new row which slot its h row goes to
          unsigned nmk = newmask;
          while (nmk) {
            const int e = __ffs(nmk) - 1;
            nmk &= nmk - 1;
            const int rr = __shfl_sync(0xffffffffu, row_of_new, e);
            const int sl = __shfl_sync(0xffffffffu, slot_of_new, e);
            if (lane == rr) { my_new |= 1u << k; my_slot[k] = sl; }
          }
          unsigned inuse = 0u;
          for (int i = 0; i < cnt; ++i) inuse |= 1u << __shfl_sync(0xffffffffu, c_sl[k], i);
          c_free[k] = (nm >= 32 ? 0xffffffffu : ((1u << nm) - 1u)) & ~inuse;
          c_cnt[k] = cnt;
        }
        // online softmax over the rows that take part
        const bool take = valid && !((my_new >> k) & 1u);
        float tmax = take ? sc_[k] : -INFINITY;
        tmax = fmaxf(tmax, ev_w[k]);
        tmax = warp_max(tmax);
        const float m_new = fmaxf(m_run[k], tmax);
        scale[k] = (m_run[k] == -INFINITY) ? 0.f : __expf(m_run[k] - m_new);
        pk[k] = (take && m_new != -INFINITY) ? __expf(sc_[k] - m_new) : 0.f;
        float lsum = pk[k];
        if (lane < n_ev[k]) {
          ev_w[k] = __expf(ev_w[k] - m_new);
          lsum += ev_w[k];
        } else {
          ev_w[k] = 0.f;
        }
        lsum = warp_sum(lsum);
        l_run[k] = l_run[k] * scale[k] + lsum;
        m_run[k] = m_new;
#pragma unroll
        for (int g = 0; g < 8; ++g) acc[g][k] *= scale[k];
      }
      {
        float4 p0, p1;
        p0.x = pk[0];
        p0.y = KB > 1 ? pk[KB > 1 ? 1 : 0] : 0.f;
        p0.z = KB > 2 ? pk[KB > 2 ? 2 : 0] : 0.f;
        p0.w = KB > 3 ? pk[KB > 3 ? 3 : 0] : 0.f;
        p1.x = KB > 4 ? pk[KB > 4 ? 4 : 0] : 0.f;
        p1.y = KB > 5 ? pk[KB > 5 ? 5 : 0] : 0.f;
        p1.z = KB > 6 ? pk[KB > 6 ? 6 : 0] : 0.f;
        p1.w = KB > 7 ? pk[KB > 7 ? 7 : 0] : 0.f;
        *reinterpret_cast<float4*>(psw + lane * 8) = p0;
        *reinterpret_cast<float4*>(psw + lane * 8 + 4) = p1;
      }
      __syncwarp();

#if GP_UMMA_PROF
      prof[4] += clock64() - t_e3;
      const long long t_e4 = clock64();
#endif
      // entries that fell out of a list this tile rejoin the sums; their h rows were parked in scratch by
      // an earlier tile and must be read BEFORE this tile's new entries reuse the freed slots
      float* cand_h = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_h) + (size_t)seg_id * K * cap * L;
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        if (k < K) {
          for (int e = 0; e < n_ev[k]; ++e) {
            const float w = __shfl_sync(0xffffffffu, ev_w[k], e);
            const int sl = __shfl_sync(0xffffffffu, ev_slot[k], e);
            if (par == 0) {
#pragma unroll
              for (int g = 0; g < 8; ++g) acc[g][k] = fmaf(w, cand_h[((size_t)k * cap + sl) * L + g * 16 + jf], acc[g][k]);
            }
          }
        }
      }
      __syncwarp();

      // ---------------- pool: acc[k][:] += sum_rows p[row][k] h[row][:]  (h re-read from the TMEM operand) ----------------
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint32_t hh[8], hl[8];
        tmem_ld8(tm + lane_addr + TM_HHI + g * 8, hh);
        tmem_ld8(tm + lane_addr + TM_HLO + g * 8, hl);
        tmem_wait_ld();
        float f[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[i]));
          const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hl[i]));
          f[2 * i] = a.x + b.x;
          f[2 * i + 1] = a.y + b.y;
        }
        const int sw = (lane >> 1) & 3;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<float4*>(tbuf + lane * 16 + ((c ^ sw) << 2)) = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
        if (my_new) {   // park this row's h where the reduce kernel (or a later late-add) finds it
#pragma unroll
          for (int k = 0; k < KB; ++k)
            if ((my_new >> k) & 1u) {
              float* dst = cand_h + ((size_t)k * cap + my_slot[k]) * L + g * 16;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
            }
        }
        __syncwarp();
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int rr = 2 * i + par;
          const float hv = tbuf[rr * 16 + ((((jf >> 2) ^ ((rr >> 1) & 3))) << 2) + (jf & 3)];
          const float4 p0 = *reinterpret_cast<const float4*>(psw + rr * 8);
          acc[g][0] = fmaf(p0.x, hv, acc[g][0]);
          if (KB > 1) acc[g][KB > 1 ? 1 : 0] = fmaf(p0.y, hv, acc[g][KB > 1 ? 1 : 0]);
          if (KB > 2) acc[g][KB > 2 ? 2 : 0] = fmaf(p0.z, hv, acc[g][KB > 2 ? 2 : 0]);
          if (KB > 3) acc[g][KB > 3 ? 3 : 0] = fmaf(p0.w, hv, acc[g][KB > 3 ? 3 : 0]);
          if (KB > 4) {
            const float4 p1 = *reinterpret_cast<const float4*>(psw + rr * 8 + 4);
            acc[g][KB > 4 ? 4 : 0] = fmaf(p1.x, hv, acc[g][KB > 4 ? 4 : 0]);
            if (KB > 5) acc[g][KB > 5 ? 5 : 0] = fmaf(p1.y, hv, acc[g][KB > 5 ? 5 : 0]);
            if (KB > 6) acc[g][KB > 6 ? 6 : 0] = fmaf(p1.z, hv, acc[g][KB > 6 ? 6 : 0]);
            if (KB > 7) acc[g][KB > 7 ? 7 : 0] = fmaf(p1.w, hv, acc[g][KB > 7 ? 7 : 0]);
          }
        }
        __syncwarp();
      }
#if GP_UMMA_PROF
      prof[5] += clock64() - t_e4;
#endif
    }
    flush_stream();
#if GP_UMMA_PROF
    prof[7] = clock64() - t_start;
    if (warp == 8 && lane == 0) PROF_FLUSH(24);
#endif
  }

  // ---- teardown ----
  tc_fence_before();
  cluster_sync();
  if (warp == 2) tmem_dealloc<2>(tm, 512);
}

// ------------------------------------------------------------------------------------------------
// weight images: per CTA c of the pair
//   W1 part (hi, then lo): for kb in [0, DIN/64): tile [64 rows = features 64c..64c+63][64 k] K-major SWIZZLE_128B (8 KB)
//   Wg part (hi, then lo): Wg = Wv for c == 0, Wu for c == 1: for h in {0,1}, kb in {0,1}:
//                          tile [64 rows = units 64h..64h+63][64 k = features 64kb..] (8 KB)
// values are scaled by the power of two in scales[] before the split (undone in the epilogue)
__global__ void umma_absmax_kernel(const float* __restrict__ w, size_t n, float* out) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // non-negative floats order as ints
}

__device__ __forceinline__ float pow2_scale(float absmax) {
  // largest power of two S with S * absmax <= 2^14 (keeps hi and lo in the fp16 normal range)
  if (!(absmax > 0.f) || !isfinite(absmax)) return 1.f;
  int e;
  frexpf(absmax, &e);            // absmax = f * 2^e, f in [0.5, 1)
  int s = 14 - e;
  s = max(-24, min(s, 24));
  return ldexpf(1.f, s);
}

__global__ void umma_pack_kernel(acmil_gp_shape sh, acmil_gp_weights w, const float* __restrict__ absmax,
                                 unsigned char* __restrict__ img, uint32_t cta_img_bytes, uint32_t w1_part_bytes) {
  const int DIN = sh.d_in;
  const float s1 = pow2_scale(absmax[0]), sv = pow2_scale(absmax[1]), su = pow2_scale(absmax[2]);
  const size_t n_w1 = (size_t)128 * DIN, n_wg = (size_t)128 * 128;
  const size_t total = n_w1 + 2 * n_wg;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float v;
    uint32_t cta, off_hi, off_lo;
    if (i < n_w1) {
      const int n = (int)(i / DIN), k = (int)(i % DIN);          // W1[n][k]
      v = w.d_w1[i] * s1;
      cta = n >> 6;
      const uint32_t tile = (uint32_t)(k >> 6) * 8192u;
      const uint32_t inner = sw128_offset((uint32_t)(n & 63), (uint32_t)((k & 63) >> 3)) + (uint32_t)(k & 7) * 2u;
      off_hi = tile + inner;
      off_lo = w1_part_bytes + tile + inner;
    } else {
      const size_t j = i - n_w1;
      const bool is_u = j >= n_wg;
      const size_t jj = is_u ? j - n_wg : j;
      const int u = (int)(jj >> 7), k = (int)(jj & 127);          // Wg[u][k]
      v = (is_u ? w.d_wu[jj] * su : w.d_wv[jj] * sv);
      cta = is_u ? 1u : 0u;
      const uint32_t tile = (uint32_t)((u >> 6) * 2 + (k >> 6)) * 8192u;
      const uint32_t inner = sw128_offset((uint32_t)(u & 63), (uint32_t)((k & 63) >> 3)) + (uint32_t)(k & 7) * 2u;
      off_hi = 2u * w1_part_bytes + tile + inner;
      off_lo = 2u * w1_part_bytes + 32768u + tile + inner;
    }
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    unsigned char* base = img + (size_t)cta * cta_img_bytes;
    *reinterpret_cast<__half*>(base + off_hi) = hi;
    *reinterpret_cast<__half*>(base + off_lo) = lo;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeFn)ptr;
  }
  return fn;
}

template <int KB>
int launch_kb(const UmmaParams& up, int grid, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    ACMIL_CHECK_CUDA(cudaFuncSetAttribute(gp_main_umma_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  gp_main_umma_kernel<KB><<<grid, UT, smem, st>>>(up);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

}  // namespace

#if GP_UMMA_PROF
extern "C" __attribute__((visibility("default"))) int acmil_debug_umma_prof(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_umma_prof, sizeof(long long) * n);
}
#endif

int gp_umma_supported(const acmil_gp_shape& s) {
  return s.front == 1 && s.front_act == ACMIL_ACT_RELU && s.act_a == ACMIL_ACT_TANH && s.gated == 1 && s.d_inner == 128 &&
         s.d_attn == 128 && s.d_in % 64 == 0 && s.d_in >= 64 && s.d_in <= 384 && s.n_branch >= 1 && s.n_branch <= KMAX;
}

// umma section of the packed blob: [absmax: 16 B][pad to 1024][cta0 image][cta1 image]
static inline uint32_t umma_cta_img_bytes(const acmil_gp_shape& s) { return (uint32_t)(s.d_in / 64) * 8192u * 2u + 65536u; }

int gp_umma_pack(const acmil_gp_shape& s, const acmil_gp_weights& w, unsigned char* d_umma, acmil_gp_consts* consts,
                 const float* d_f32, const GpPackLayout& lay, cudaStream_t st) {
  float* d_absmax = reinterpret_cast<float*>(d_umma);
  unsigned char* img = d_umma + 1024;
  const uint32_t cta_img = umma_cta_img_bytes(s), w1_part = (uint32_t)(s.d_in / 64) * 8192u;
  ACMIL_CHECK_CUDA(cudaMemsetAsync(d_absmax, 0, 16, st));
  umma_absmax_kernel<<<64, 256, 0, st>>>(w.d_w1, (size_t)128 * s.d_in, d_absmax + 0);
  umma_absmax_kernel<<<16, 256, 0, st>>>(w.d_wv, (size_t)128 * 128, d_absmax + 1);
  umma_absmax_kernel<<<16, 256, 0, st>>>(w.d_wu, (size_t)128 * 128, d_absmax + 2);
  umma_pack_kernel<<<148, 256, 0, st>>>(s, w, d_absmax, img, cta_img, w1_part);
  g_acmil_launches += 4;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  if (consts) {
    // host copy of the small vectors (already laid out, zero-filled where absent, by the fp32 pack)
    float absmax[4] = {0, 0, 0, 0};
    ACMIL_CHECK_CUDA(cudaMemcpyAsync(absmax, d_absmax, 16, cudaMemcpyDeviceToHost, st));
    ACMIL_CHECK_CUDA(cudaMemcpyAsync(consts->b1, d_f32 + lay.b1, 128 * 4, cudaMemcpyDeviceToHost, st));
    ACMIL_CHECK_CUDA(cudaMemcpyAsync(consts->bv, d_f32 + lay.bv, 128 * 4, cudaMemcpyDeviceToHost, st));
    ACMIL_CHECK_CUDA(cudaMemcpyAsync(consts->bu, d_f32 + lay.bu, 128 * 4, cudaMemcpyDeviceToHost, st));
    ACMIL_CHECK_CUDA(cudaMemcpyAsync(consts->ww, d_f32 + lay.ww, KMAX * 128 * 4, cudaMemcpyDeviceToHost, st));
    ACMIL_CHECK_CUDA(cudaMemcpyAsync(consts->bw, d_f32 + lay.bw, KMAX * 4, cudaMemcpyDeviceToHost, st));
    ACMIL_CHECK_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 3; ++i) {
      float sc = 1.f;
      if (absmax[i] > 0.f && isfinite(absmax[i])) {
        int e;
        frexpf(absmax[i], &e);
        int sh = 14 - e;
        sh = sh < -24 ? -24 : (sh > 24 ? 24 : sh);
        sc = ldexpf(1.f, sh);
      }
      consts->inv_scale[i] = 1.f / sc;
    }
    consts->inv_scale[3] = 1.f;
    consts->valid = ACMIL_ABI_VERSION;
  }
  return ACMIL_OK;
}

// partition of the pair-tiles over the clusters + the segment table the reduce kernel consumes
int gp_umma_build_plan(const acmil_gp_batch& b, int sm_count, GpSegTable* t) {
  memset(t, 0, sizeof(*t));
  t->n_slides = b.n_slides;
  t->tile_rows = 256;
  int total_pt = 0, cap = 0;
  for (int s = 0; s < b.n_slides; ++s) {
    const int64_t n = b.row_offsets[s + 1] - b.row_offsets[s];
    if (n < 0) return -1;
    t->u_pt_begin[s] = total_pt;
    total_pt += (int)((n + 255) / 256);
    const int nm = b.n_masked > 0 ? (int)(n < b.n_masked ? n : b.n_masked) : 0;
    t->nm[s] = nm;
    if (nm > cap) cap = nm;
    t->row_off[s] = b.row_offsets[s];
    t->shard_begin[s] = b.shard_row_begin ? b.shard_row_begin[s] : 0;
  }
  t->u_pt_begin[b.n_slides] = total_pt;
  t->row_off[b.n_slides] = b.row_offsets[b.n_slides];
  t->u_total_pt = total_pt;
  if ((int64_t)total_pt * (sm_count / 2 + 1) >= (int64_t)0x7fffffff) return -1;   // 32-bit partition arithmetic
  int ncl = sm_count / 2;
  if (ncl > total_pt) ncl = total_pt;
  if (cap > 0 && ncl * 8 * cap > GP_MAX_SEG_CAND) ncl = GP_MAX_SEG_CAND / (8 * cap);   // reduce kernel's smem bound
  if (ncl < 1) ncl = 1;
  t->u_nclusters = ncl;
  t->n_masked_cap = cap;
  // segments: 8 per (cluster, bag) pair that intersects
  int seg = 0;
  for (int s = 0; s < b.n_slides; ++s) {
    t->seg_begin[s] = seg;
    const int p0 = t->u_pt_begin[s], p1 = t->u_pt_begin[s + 1];
    int first = -1, count = 0;
    for (int c = 0; c < ncl && total_pt > 0; ++c) {
      const int g0 = (int)(((uint32_t)c * (uint32_t)total_pt) / (uint32_t)ncl);
      const int g1 = (int)(((uint32_t)(c + 1) * (uint32_t)total_pt) / (uint32_t)ncl);
      if (g0 < p1 && g1 > p0) {
        if (first < 0) first = c;
        ++count;
      }
    }
    t->u_cfirst[s] = first < 0 ? 0 : first;
    seg += 8 * count;
  }
  t->seg_begin[b.n_slides] = seg;
  t->n_seg = seg;
  return 0;
}

int gp_launch_main_umma(const GpMainParams& p, const acmil_gp_consts* consts, const unsigned char* d_umma, cudaStream_t st) {
  const acmil_gp_shape& s = p.sh;
  ACMIL_REQUIRE(gp_umma_supported(s), ACMIL_E_UNSUPPORTED, "tcgen05 kernel does not support this shape");
  ACMIL_REQUIRE(consts != nullptr && consts->valid == ACMIL_ABI_VERSION, ACMIL_E_INVALID,
                "tcgen05 kernel needs the acmil_gp_consts filled by acmil_gp_pack");
  if (p.seg.u_total_pt == 0) return ACMIL_OK;
  static thread_local UmmaParams up;   // large: keep off the stack
  up.mp = p;
  memcpy(up.c.b1, consts->b1, sizeof(up.c.b1));
  memcpy(up.c.bv, consts->bv, sizeof(up.c.bv));
  memcpy(up.c.bu, consts->bu, sizeof(up.c.bu));
  memcpy(up.c.ww, consts->ww, sizeof(up.c.ww));
  memcpy(up.c.bw, consts->bw, sizeof(up.c.bw));
  up.c.inv_s1 = consts->inv_scale[0];
  up.c.inv_sv = consts->inv_scale[1];
  up.c.inv_su = consts->inv_scale[2];
  up.wimg = d_umma + 1024;
  up.cta_img_bytes = umma_cta_img_bytes(s);
  up.w1_part_bytes = (uint32_t)(s.d_in / 64) * 8192u;
  EncodeFn enc = get_encode();
  ACMIL_REQUIRE(enc != nullptr, ACMIL_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const int64_t rows = p.seg.row_off[p.seg.n_slides];
  cuuint64_t dims[2] = {(cuuint64_t)s.d_in, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)s.d_in * 4};
  cuuint32_t box[2] = {KC, 128};
  cuuint32_t es[2] = {1, 1};
  ACMIL_REQUIRE(((uintptr_t)p.x & 15) == 0, ACMIL_E_INVALID, "x must be 16-byte aligned for TMA");
  const CUresult r = enc(&up.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ACMIL_REQUIRE(r == CUDA_SUCCESS, ACMIL_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  const size_t smem = smem_map(s.d_in).total + 1024;
  const int grid = p.seg.u_nclusters * 2;
  const int K = s.n_branch;
  if (K == 1) return launch_kb<1>(up, grid, smem, st);
  if (K <= 5) return launch_kb<5>(up, grid, smem, st);
  return launch_kb<8>(up, grid, smem, st);
}
