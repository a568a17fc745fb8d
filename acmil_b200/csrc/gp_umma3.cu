// Fused gated-attention pool row pass on the 5th-gen tensor cores, second generation: TWO row tiles in flight per
// CTA pair and role-specialised epilogue warps.
//
// gp_umma.cu (one tile in flight, eight all-in-one epilogue warps) spends 13 k cycles per 256-row tile in strictly
// serial phases -- D1 -> relu/split (TMEM traffic) -> gate (MUFU-bound) -> softmax -> pool (mma.sync) -- against an
// HBM budget of 8.7 k.  Both warps of a scheduler are always in the same phase, so the MUFU-bound phase never
// overlaps the TMEM / tensor-bound ones.  Here the phases belong to different warps, one of each kind per scheduler:
//
//   warps  8-15  "G" gate warps   : thread = row, each warp half the units of a quarter; D2 -> tanh . sigmoid (3 MUFU
//                                   per unit) -> partial score dot products -> red.shared into the score tile
//   warps 16-19  "P" pool warps   : thread = row; Epi1 (D1 -> relu -> fp16 hi/lo, written IN PLACE over D1) one tile
//                                   ahead, then per tile: scores -> a_out, top-n candidates, softmax numerators,
//                                   pool on mma.sync (movmatrix transposes of the TMEM h operand)
//   warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM allocator + top-n list manager, warp 3 bag-wide
//   threshold service, warps 4-7 converters (fp32 staging -> fp16 hi/lo x operand in TMEM) -- as in gp_umma.cu.
//
// TMEM (512 columns): three D1/h buffers of 128 (tile t lives in buffer t % 3: projection accumulator, then -- after
// Epi1 -- the h operand of the gate GEMM and of the pool, per 32-feature chunk [16 columns hi | 16 columns lo]),
// one 64-column gate accumulator (a quarter = 32 V + 32 U units; single-buffered: the G warps pull it into registers
// and hand it back at once), two 32-column x-operand slots.  While the G warps work on tile t, the P warps finish
// tile t - 1 and convert tile t + 1, and the tensor core fills tile t + 2.
#include <stdlib.h>

#include "gp_umma_shared.cuh"

namespace {
using namespace sm100;
using namespace umma_shared;

constexpr int UT3 = 640;
constexpr int NXOP3 = 2;
constexpr uint32_t TM3_D2 = 384, TM3_X = 448;
__device__ __forceinline__ constexpr uint32_t tm3_dh(int b) { return 128u * (uint32_t)b; }

#ifndef GP_UMMA_PROF
#define GP_UMMA_PROF 0
#endif
#ifndef GP_EXP_NO_HMMA
#define GP_EXP_NO_HMMA 0
#endif
#ifndef GP_EXP_NOEPI
#define GP_EXP_NOEPI 0      // timing experiment: G / P warps keep the barrier protocol but skip their arithmetic
#endif
#ifndef GP_EXP_ONEMMA
#define GP_EXP_ONEMMA 0     // timing experiment: one MMA per product instead of three
#endif
#ifndef GP_L2_AHEAD
#define GP_L2_AHEAD 0       // tiles pulled into L2 ahead of the TMA loads (0: off)
#endif
#ifndef GP_MGR_SLEEP
#define GP_MGR_SLEEP 400
#endif
#ifndef GP_SVC_SLEEP
#define GP_SVC_SLEEP 2000
#endif
#if GP_UMMA_PROF
__device__ long long g_umma3_prof[148][20][8];      // [CTA][warp][slot]
#define PROF_T0() const long long _t0 = clock64()
#define PROF_ADD(slot) prof[slot] += clock64() - _t0
#define PROF_DECL() long long prof[8] = {0}
#define PROF_FLUSH() do { for (int _i = 0; _i < 8; ++_i) g_umma3_prof[blockIdx.x][warp][_i] = prof[_i]; } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF_ADD(slot) do {} while (0)
#define PROF_DECL() do {} while (0)
#define PROF_FLUSH() do {} while (0)
#endif

struct SmemMap3 {
  uint32_t w1, wg, stage, tbuf, ps, sc, cand, bars, total;
};

__host__ __device__ inline SmemMap3 smem_map3(int din, int kb) {
  SmemMap3 m;
  m.w1 = 0;
  m.wg = (uint32_t)(din / 64) * 8192u * 2u;
  m.stage = m.wg + 65536u;
  m.tbuf = m.stage + NSTAGE * STAGE_BYTES;      // first-tile selection flags [128][8] (1 KB) + appended record scores
  m.ps = m.tbuf + 7 * 1024;                     // softmax numerators of the P warps [4][32 rows][8]
  m.sc = m.ps + 4 * 1024;                       // score tile [128 rows][8]: the two G warps of a row add their halves
  m.cand = m.sc + 4 * 1024;
  m.bars = m.cand + (kb > 6 ? 128u : (uint32_t)sizeof(CandShared));
  m.total = m.bars + 512;
  return m;
}

struct Bars3 {
  uint64_t full_x[NSTAGE], empty_x[NSTAGE];
  uint64_t xop_full[NXOP3], xop_empty[NXOP3];
  uint64_t dh_full[3], h_full[3], dh_free[3];
  uint64_t d2_full, d2_empty, sc_full, sc_empty, wload, w_ready;
  uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_ld_16x128b_x4(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// ===================================== G: gate warps =====================================
// warp (lq, HALF): TMEM lanes / tile rows [32 lq, +32), thread = row; units 16 HALF + [0, 16) of every quarter.
// tanh(a) sigmoid(b) = (1 - Ea) / ((1 + Ea)(1 + Eb)), Ea = e^-2a, Eb = e^-b, on packed fp32; the score weights and
// the exponent-domain biases are warp-uniform and come straight from the kernel parameters (constant bank).
template <int KB, int HALF>
__device__ __forceinline__ void gate_warp(const UmmaParams& p, uint8_t* sc_base, Bars3* bars, uint32_t tm, int warp, int lane,
                                          int T, int K) {
  const int lq = (warp - 8) & 3;
  const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
  const int row_local = lq * 32 + lane;
  float* scp = reinterpret_cast<float*>(sc_base) + row_local * 8;
  const float cva = p.c.inv_sv * (-2.f * LOG2E), cua = p.c.inv_su * (-LOG2E);
  const uint64_t cva2 = pack2(cva, cva), cua2 = pack2(cua, cua), one2 = pack2(1.f, 1.f), mone2 = pack2(-1.f, -1.f);
  PROF_DECL();
#if GP_UMMA_PROF
  const long long t_start = clock64();
#endif
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    uint64_t s2[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) s2[k] = 0ull;
#pragma unroll
    for (int qr = 0; qr < 4; ++qr) {
      { PROF_T0(); mbar_wait_cluster(&bars->d2_full, (uint32_t)(4 * t + qr) & 1u); PROF_ADD(0); }
      tc_fence_after();
      uint32_t zv[16], zu[16];
      tmem_ld16(tm + lane_addr + TM3_D2 + HALF * 16, zv);
      tmem_ld16(tm + lane_addr + TM3_D2 + 32 + HALF * 16, zu);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&bars->d2_empty, 0);   // the quarter is in registers: the MMA warp may refill D2
#pragma unroll
      for (int j = 0; j < (GP_EXP_NOEPI ? 0 : 8); ++j) {
        const int u = qr * 32 + HALF * 16 + 2 * j;
        const uint64_t bvx2 = *reinterpret_cast<const uint64_t*>(&p.c.bvx[u]);
        const uint64_t bux2 = *reinterpret_cast<const uint64_t*>(&p.c.bux[u]);
        const uint64_t xa = fma2(pack2(__uint_as_float(zv[2 * j]), __uint_as_float(zv[2 * j + 1])), cva2, bvx2);
        const uint64_t xb = fma2(pack2(__uint_as_float(zu[2 * j]), __uint_as_float(zu[2 * j + 1])), cua2, bux2);
        // Ea's exponent is clamped so that (1 - Ea) stays finite; Eb may overflow to +inf: the quotient is then 0
        const uint64_t ea = pack2(ex2_approx(fminf(lo2(xa), 57.7f)), ex2_approx(fminf(hi2(xa), 57.7f)));
        const uint64_t eb = pack2(ex2_approx(lo2(xb)), ex2_approx(hi2(xb)));
        const uint64_t den = mul2(add2(ea, one2), add2(eb, one2));
        const uint64_t g = mul2(fma2(ea, mone2, one2), pack2(rcp_approx(lo2(den)), rcp_approx(hi2(den))));
#pragma unroll
        for (int k = 0; k < KB; ++k) s2[k] = fma2(g, *reinterpret_cast<const uint64_t*>(&p.c.ww[k][u]), s2[k]);
      }
    }
    // this warp's half of the dot products -> the score tile (two contributions per entry: order-independent)
    if (t > 0) { PROF_T0(); mbar_wait(&bars->sc_empty, (uint32_t)(t - 1) & 1u); PROF_ADD(1); }
#pragma unroll
    for (int k = 0; k < KB; ++k)
      if (k < K) atomicAdd(scp + k, lo2(s2[k]) + hi2(s2[k]));
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars->sc_full);
  }
#if GP_UMMA_PROF
  prof[7] = clock64() - t_start;
  if (lane == 0) PROF_FLUSH();
#endif
}

template <int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UT3, 1) gp_main_umma3_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int DIN = p.mp.sh.d_in;
  const int NCH = DIN / KC;
  const SmemMap3 sm = smem_map3(DIN, KB);
  Bars3* bars = reinterpret_cast<Bars3*>(smem + sm.bars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cta = cluster_ctarank();
  const int cluster = blockIdx.x >> 1;
  const GpSegTable& seg = p.mp.seg;
  const int K = p.mp.sh.n_branch;

  const int g0 = (int)(((uint32_t)cluster * (uint32_t)seg.u_total_pt) / (uint32_t)seg.u_nclusters);
  const int g1 = (int)(((uint32_t)(cluster + 1) * (uint32_t)seg.u_total_pt) / (uint32_t)seg.u_nclusters);
  const int T = g1 - g0;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars->full_x[i], 1); mbar_init(&bars->empty_x[i], 4); }
    for (int i = 0; i < NXOP3; ++i) { mbar_init(&bars->xop_full[i], 8); mbar_init(&bars->xop_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&bars->dh_full[i], 1); mbar_init(&bars->h_full[i], 8); mbar_init(&bars->dh_free[i], 8); }
    mbar_init(&bars->d2_full, 1);
    mbar_init(&bars->d2_empty, 16);
    mbar_init(&bars->sc_full, 8);
    mbar_init(&bars->sc_empty, 4);
    mbar_init(&bars->wload, 1);
    mbar_init(&bars->w_ready, 2);
    fence_mbar_init();
    const unsigned char* src = p.wimg + (size_t)cta * p.cta_img_bytes;
    mbar_expect_tx(&bars->wload, p.cta_img_bytes);
    for (uint32_t off = 0; off < p.cta_img_bytes; off += 16384) bulk_load(smem + off, src + off, 16384, &bars->wload);
    tma_prefetch_desc(&p.tmap);
  }
  for (int i = tid; i < 1024; i += UT3) reinterpret_cast<float*>(smem + sm.sc)[i] = 0.f;
  if (KB <= 6) {
    CandShared* cs0 = reinterpret_cast<CandShared*>(smem + sm.cand);
    if (tid < CAND_KMAX * 32) cs0->ls[tid >> 5][tid & 31] = INFINITY;
    if (tid < 8) { cs0->cnt[tid] = 0; cs0->app[tid] = 0; cs0->seen[tid] = 0; cs0->tau[tid] = -INFINITY; cs0->gtau[tid] = ~0ull; }
    if (tid == 0) { cs0->cur_bag = -1; cs0->epoch = 0; cs0->flush_req = 0; cs0->flush_ack = 0; cs0->rows = 0; }
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

  // NOTE: each setmaxnreg sits at the top of a branch that never rejoins the others before the kernel's tail.
  // Register budget: the CTA is launched with 640 x 96; setmaxnreg.inc can only take what setmaxnreg.dec released
  // inside the CTA: 4 x (96 - 64) + 4 x (96 - 88) = 4 x (136 - 96).
  if (warp < 4) {
  setmaxnreg_dec<64>();
  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // The staging ring holds 3 x 16 KB per SM, far less than bandwidth x HBM latency (22 B / cycle x ~3 k cycles under
    // load), so every tile is pulled into L2 GP_L2_AHEAD tiles before its loads are issued: the ring then only has to
    // cover the L2 latency.  (148 SMs x 2 tiles x 196 KB = 58 MB of the 126 MB L2.)
    if (lane == 0) {
      int s_hint = 0, s_hint2 = 0;
      uint32_t ctr = 0;
      auto prefetch_tile = [&](int g) {
        const TilePos tq = tile_pos(g, s_hint2);
        const int64_t grow = seg.row_off[tq.s] + tq.row_in_bag;
        for (int c = 0; c < NCH; ++c) tma_prefetch_l2_2d(&p.tmap, c * KC, (int)grow);
      };
      for (int g = g0; g < g1 && g < g0 + GP_L2_AHEAD; ++g) prefetch_tile(g);
      for (int g = g0; g < g1; ++g) {
        if (g + GP_L2_AHEAD < g1) prefetch_tile(g + GP_L2_AHEAD);
        const TilePos tp = tile_pos(g, s_hint);
        const int64_t grow = seg.row_off[tp.s] + tp.row_in_bag;
        for (int c = 0; c < NCH; ++c, ++ctr) {
          const uint32_t st = ctr % NSTAGE, ph = (ctr / NSTAGE) & 1u;
          mbar_wait(&bars->empty_x[st], ph ^ 1u);
          mbar_expect_tx(&bars->full_x[st], STAGE_BYTES);
          tma_load_2d_hint(smem + sm.stage + st * STAGE_BYTES, &p.tmap, c * KC, (int)grow, &bars->full_x[st], kEvictFirst);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer 1 (leader CTA): projection GEMM =====================================
    // Two issuer warps, one per GEMM, each walking its own work in order with blocking waits (warp 20 issues the gate
    // GEMM).  A single warp polling both GEMMs' barriers (mbarrier.test_wait: ~150 exposed cycles per probe, 2-3
    // probes per issued group) was itself the bottleneck of this kernel: ~550 cycles per group of 6 / 24 MMAs.
    // The whole warp runs the loop on identical values; only the tcgen05 instructions sit under elect_one().
    if (cta == 0 && T > 0) {
      PROF_DECL();
#if GP_UMMA_PROF
      const long long t_start = clock64();
#endif
      mbar_wait_cluster(&bars->w_ready, 0);      // both CTAs' weight images have landed
      const uint32_t idesc = umma_idesc_f16(256, 128);
      const uint32_t w1_hi = smem_u32(smem + sm.w1), w1_lo = w1_hi + p.w1_part_bytes;
      uint32_t xc = 0;  // x-operand chunks consumed so far (ring position / phase)
      for (int t = 0; t < T; ++t) {
        const int b = t % 3;
        const uint32_t d = tm + tm3_dh(b);
        // the buffer is free once the pool of tile t - 3 has read its h operand
        if (t >= 3) { PROF_T0(); mbar_wait_cluster(&bars->dh_free[b], (uint32_t)(t / 3 - 1) & 1u); PROF_ADD(0); }
        for (int c = 0; c < NCH; ++c, ++xc) {
          const uint32_t q = xc % NXOP3, ph = (xc / NXOP3) & 1u;
          { PROF_T0(); mbar_wait_cluster(&bars->xop_full[q], ph); PROF_ADD(1); }
          tc_fence_after();
          const uint32_t xa_hi = tm + TM3_X + q * 32, xa_lo = xa_hi + 16;
          const uint32_t boff = (uint32_t)(c >> 1) * 8192u + (uint32_t)(c & 1) * 64u;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t bhi = umma_desc_k_sw128(w1_hi + boff + ks * 32), blo = umma_desc_k_sw128(w1_lo + boff + ks * 32);
              umma_ts<2>(d, xa_hi + ks * 8, bhi, idesc, (c | ks) ? 1u : 0u);
#if !GP_EXP_ONEMMA
              umma_ts<2>(d, xa_lo + ks * 8, bhi, idesc, 1u);
              umma_ts<2>(d, xa_hi + ks * 8, blo, idesc, 1u);
#endif
            }
            umma_commit_2sm(&bars->xop_empty[q], 3);
            if (c == NCH - 1) umma_commit_2sm(&bars->dh_full[b], 3);
          }
          __syncwarp();
        }
      }
#if GP_UMMA_PROF
      prof[7] = clock64() - t_start;
      if (lane == 0) PROF_FLUSH();
#endif
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===================================== MMA issuer 2 (leader CTA): gate GEMM =====================================
    // tile t in four 32-unit quarters (N = 64: 32 V + 32 U columns) through the single D2 buffer
    if (cta == 0 && T > 0) {
      PROF_DECL();
#if GP_UMMA_PROF
      const long long t_start = clock64();
#endif
      mbar_wait_cluster(&bars->w_ready, 0);
      const uint32_t idesc64 = umma_idesc_f16(256, 64);
      const uint32_t wg_hi = smem_u32(smem + sm.wg), wg_lo = wg_hi + 32768u;
      for (int t = 0; t < T; ++t) {
        const int b = t % 3;
        const uint32_t h = tm + tm3_dh(b);
        { PROF_T0(); mbar_wait_cluster(&bars->h_full[b], (uint32_t)(t / 3) & 1u); PROF_ADD(0); }
#pragma unroll 1
        for (int qr = 0; qr < 4; ++qr) {
          const uint32_t n = 4u * (uint32_t)t + (uint32_t)qr;      // quarters issued before this one
          if (n > 0) { PROF_T0(); mbar_wait_cluster(&bars->d2_empty, (n - 1u) & 1u); PROF_ADD(1); }
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t boff = (uint32_t)(qr * 2 + (ks >> 2)) * 4096u + (uint32_t)(ks & 3) * 32u;
              const uint64_t bhi = umma_desc_k_sw128(wg_hi + boff), blo = umma_desc_k_sw128(wg_lo + boff);
              const uint32_t a_hi = h + (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u, a_lo = a_hi + 16u;
              umma_ts<2>(tm + TM3_D2, a_hi, bhi, idesc64, ks ? 1u : 0u);
#if !GP_EXP_ONEMMA
              umma_ts<2>(tm + TM3_D2, a_lo, bhi, idesc64, 1u);
              umma_ts<2>(tm + TM3_D2, a_hi, blo, idesc64, 1u);
#endif
            }
            umma_commit_2sm(&bars->d2_full, 3);
          }
          __syncwarp();
        }
      }
#if GP_UMMA_PROF
      prof[7] = clock64() - t_start;
      if (lane == 0) PROF_FLUSH();
#endif
    }
    __syncwarp();
  } else {
    // warp 3
    if (lane == 0) {      // this CTA's resident weights are in place -> tell the leader's MMA threads
      mbar_wait(&bars->wload, 0);
      mbar_arrive_cluster(&bars->w_ready, 0);
    }
    __syncwarp();
    if (KB <= CAND_KMAX && seg.n_masked_cap > 0 && T > 0) {
      // ============================ top-n list manager + bag-wide threshold service ============================
      // One warp does both jobs (see gp_umma_shared.cuh): the manager scan is cheap and latency-sensitive, so it runs
      // every round and between the chunks of the service work; the service merges ONE branch per round.
      CandShared* cs = reinterpret_cast<CandShared*>(smem + sm.cand);
      const unsigned* recs = reinterpret_cast<const unsigned*>(smem + sm.tbuf + 1024);      // [K][rec_cap] score bits
      float* g_score_all = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_score);
      const int cap = seg.n_masked_cap, rcap = seg.rec_cap;
      int dead_epoch = 0;       // lists of this epoch are final (or not booted yet): hands off
      // one manager scan; false when the kernel is finishing
      auto manager_scan = [&](bool& busy) -> bool {
        const int s = *reinterpret_cast<volatile int*>(&cs->cur_bag);
        if (s == -2) return false;
        const int ep = *reinterpret_cast<volatile int*>(&cs->epoch);
        const int fr = *reinterpret_cast<volatile int*>(&cs->flush_req);      // read BEFORE the scan: if it already
        if (ep != dead_epoch && s >= 0) {                                      // asks for ep, the scan below sees all
          const int nmk = seg.nm[s];
          const int holder = seg.seg_begin[s] + (cluster - seg.u_cfirst[s]) * 2 + (int)cta;
          for (int k = 0; k < K; ++k) {
            int seen = cs->seen[k];
            const int app = min(*reinterpret_cast<volatile int*>(&cs->app[k]), rcap);
            if (seen >= app) continue;
            busy = true;
            float mg_s = cs->ls[k][lane];
            int mg_rec = cs->lrec[k][lane], mg_cnt = cs->cnt[k];
            float mg_tau = -INFINITY;
            int mg_tau_lane = 0;
            auto find_min = [&]() {
              const unsigned key = __reduce_min_sync(0xffffffffu, ord_enc(mg_s));
              mg_tau = ord_dec(key);
              mg_tau_lane = __ffs(__ballot_sync(0xffffffffu, ord_enc(mg_s) == key)) - 1;
            };
            if (mg_cnt == nmk) find_min();
            bool changed = false;
            for (; seen < app; ++seen) {
              const unsigned bits = *reinterpret_cast<const volatile unsigned*>(&recs[k * rcap + seen]);
              if (bits == REC_EMPTY) break;       // slot taken, score not stored yet: next round
              const float s_new = __uint_as_float(bits);
              if (mg_cnt == nmk && !(s_new > mg_tau)) continue;
              const int dst = mg_cnt < nmk ? mg_cnt++ : mg_tau_lane;
              if (lane == dst) { mg_s = s_new; mg_rec = seen; }
              if (mg_cnt == nmk) find_min();
              changed = true;
            }
            if (changed) {
              cs->ls[k][lane] = mg_s;
              cs->lrec[k][lane] = mg_rec;
              if (lane < cap) g_score_all[((size_t)holder * K + k) * cap + lane] = lane < mg_cnt ? mg_s : -INFINITY;
              if (lane == 0) {
                cs->cnt[k] = mg_cnt;
                if (mg_cnt == nmk) *reinterpret_cast<volatile float*>(&cs->tau[k]) = mg_tau;
              }
            }
            if (lane == 0) cs->seen[k] = seen;
            __syncwarp();
          }
          if (fr == ep) {       // every append of this bag happened before the request: the lists are final
            __threadfence_block();
            if (lane == 0) *reinterpret_cast<volatile int*>(&cs->flush_ack) = ep;
            dead_epoch = ep;
          }
        }
        return true;
      };
      // bag-wide threshold of ONE branch: merge the mirrored top-n lists of every CTA working on bag s; the n-th best
      // of their union is the n-th best of all rows anybody has scored so far, a lower bound of the final one.
      // Entries not written in this launch are NaN (memset by the host) and ignored.
      auto service_branch = [&](int s, int k) -> bool {
        const int nmk = seg.nm[s];
        const float* g_score = g_score_all;
        const int seg0 = seg.seg_begin[s], E = (seg.seg_begin[s + 1] - seg0) * cap;
        float ls = INFINITY, tau = -INFINITY;     // lane i = entry i of the merged top-n
        int cnt = 0, tau_lane = 0;
        for (int e0 = 0; e0 < E; e0 += 256) {
          float v[8];       // 8 independent L2 reads in flight per lane
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int e = e0 + 32 * u + lane;
            v[u] = -INFINITY;
            if (e < E) {
              const int j = e / cap, i = e - j * cap;
              v[u] = __ldcg(g_score + ((size_t)(seg0 + j) * K + k) * cap + i);
              if (!(v[u] == v[u])) v[u] = -INFINITY;
            }
          }
          bool busy = false;
          if (!manager_scan(busy)) return false;      // (while the loads are in flight)
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            unsigned bal = __ballot_sync(0xffffffffu, v[u] > tau);
            while (bal) {
              const int src = __ffs(bal) - 1;
              bal &= bal - 1;
              const float s_new = __shfl_sync(0xffffffffu, v[u], src);
              if (cnt == nmk && !(s_new > tau)) continue;
              const int dst = cnt < nmk ? cnt++ : tau_lane;
              if (lane == dst) ls = s_new;
              if (cnt == nmk) {
                const unsigned key = __reduce_min_sync(0xffffffffu, ord_enc(ls));
                tau = ord_dec(key);
                tau_lane = __ffs(__ballot_sync(0xffffffffu, ord_enc(ls) == key)) - 1;
              }
            }
          }
          if (*reinterpret_cast<volatile int*>(&cs->cur_bag) != s) return true;     // bag changed / kernel finishing
        }
        if (cnt == nmk && lane == 0)
          *reinterpret_cast<volatile unsigned long long*>(&cs->gtau[k]) = ((unsigned long long)(unsigned)s << 32) | __float_as_uint(tau);
        return true;
      };
      int round = 0, kk = 0;
      while (true) {
        bool busy = false;
        if (!manager_scan(busy)) break;
        const int s = *reinterpret_cast<volatile int*>(&cs->cur_bag);
        if (s >= 0 && seg.nm[s] > 0 && (++round & 1) == 0) {
          if (!service_branch(s, kk)) break;
          kk = kk + 1 < K ? kk + 1 : 0;
        } else if (!busy) {
          // idle polls back off; a pending flush request or fresh records are served right away
          __nanosleep(*reinterpret_cast<volatile int*>(&cs->flush_req) != dead_epoch ? 40 : GP_MGR_SLEEP);
        }
      }
    }
  }
  } else if (warp < 8) {
    // ===================================== converters: fp32 staging -> fp16 hi/lo in TMEM =====================================
    setmaxnreg_dec<88>();
    const int r = (warp - 4) * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp - 4) * 32) << 16;
    uint32_t ctr = 0;
    PROF_DECL();
#if GP_UMMA_PROF
    const long long t_start = clock64();
#endif
    // Epi1 of tile t (D1 -> relu -> fp16 hi/lo operand of the gate GEMM, IN PLACE: the 32 fp32 columns of a 32-feature
    // chunk become [16 columns hi | 16 columns lo], two features per column) also runs here, in four column chunks
    // behind x chunks 1..4 of tile t + 1 (by then D1 of tile t is complete, and the MMA warp still has x operands
    // queued): these warps have the thread = row TMEM view and idle half of the time, while the pool warps are the
    // longest stage of the pipeline.  Early in the tile, because the gate GEMM of tile t waits for it.
    auto epi1_chunk = [&](int t, int c4) {
      const int b = t % 3;
      if (c4 == 0) { PROF_T0(); mbar_wait_cluster(&bars->dh_full[b], (uint32_t)(t / 3) & 1u); PROF_ADD(2); tc_fence_after(); }
      PROF_T0();
      const uint32_t base = tm + lane_addr + tm3_dh(b) + 32 * c4;
      uint32_t v[32];
      tmem_ld32(base, v);
      tmem_wait_ld();
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
#if GP_EXP_NOEPI
        hi[i] = v[2 * i]; lo[i] = v[2 * i + 1];
#else
        split2(fmaxf(__uint_as_float(v[2 * i]) * p.c.inv_s1, 0.f), fmaxf(__uint_as_float(v[2 * i + 1]) * p.c.inv_s1, 0.f),
               hi[i], lo[i]);
#endif
      }
      tmem_st16(base, hi);
      tmem_st16(base + 16, lo);
      if (c4 == 3) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bars->h_full[b], 0);
      }
      PROF_ADD(3);
    };
    for (int t = 0; t <= T; ++t) {
      if (t == T) {      // no next tile to interleave with: the last tile's Epi1 in one go
        if (T > 0) for (int c4 = 0; c4 < 4; ++c4) epi1_chunk(T - 1, c4);
        break;
      }
      int e_done = 0;
      for (int c = 0; c < NCH; ++c, ++ctr) {
        const uint32_t st = ctr % NSTAGE, ph = (ctr / NSTAGE) & 1u;
        const uint32_t q = ctr % NXOP3, phq = (ctr / NXOP3) & 1u;
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
        tmem_st16(tm + lane_addr + TM3_X + q * 32, hi);
        tmem_st16(tm + lane_addr + TM3_X + q * 32 + 16, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bars->xop_full[q], 0);
        if (t > 0 && e_done < 4 && c >= 1) epi1_chunk(t - 1, e_done++);
      }
      if (t > 0) for (; e_done < 4; ++e_done) epi1_chunk(t - 1, e_done);
    }
#if GP_UMMA_PROF
    prof[7] = clock64() - t_start;
    if (lane == 0) PROF_FLUSH();
#endif
  } else if (warp < 16) {
    // ===================================== G: gate warps =====================================
    // (the unit half is a template parameter: with compile-time unit indices the score weights and biases become
    // uniform-register operands of the packed FFMA2s, LDCU.128 on the uniform datapath; with a run-time half every one
    // of them is an indexed LDC.64 per thread -- measured 2.3x slower)
    if ((warp - 8) >> 2) gate_warp<KB, 1>(p, smem + sm.sc, bars, tm, warp, lane, T, K);
    else gate_warp<KB, 0>(p, smem + sm.sc, bars, tm, warp, lane, T, K);
  } else {
    // ===================================== P: pool warps =====================================
    setmaxnreg_inc<136>();
    const int w = warp - 16;                       // TMEM lane quarter; thread = row 32 w + lane of the tile
    const uint32_t lane_addr = (uint32_t)(w * 32) << 16;
    const int rg = lane >> 2, cp = lane & 3;       // fragment coordinates of the pool's mma.sync step
    const int row_local = w * 32 + lane;
    float* psw = reinterpret_cast<float*>(smem + sm.ps + w * 1024);       // [32 rows][8]
    float* scp = reinterpret_cast<float*>(smem + sm.sc) + row_local * 8;
    const int L = 128;
    const int cap = seg.n_masked_cap;

    // Softmax-pool state of this warp's stream (see gp_umma.cu): p' = exp(s - m_ref[k]) * 2^PSH against a per-warp
    // reference that only grows, in steps of more than REF_SLACK nats; acc = the mma.sync D fragments:
    // acc[j][0..1] = feature 16 j + rg, branches 2 cp, 2 cp + 1; acc[j][2..3] = feature 16 j + rg + 8.
    constexpr float PSH = 4.f, REF_SLACK = 7.6f;
    float l_run[KB], m_ref[KB], c_ref[KB], acc[8][4];
    CandShared* cs = reinterpret_cast<CandShared*>(smem + sm.cand);
    const int rcap = seg.rec_cap, rowcap = seg.row_cap;
    int s_cur = -1, s_hint = 0, nm = 0, seg_id = 0, cb = 0;
    bool boot = false;
    int epoch_cur = 0;
    int64_t n_rows = 0;

    auto reset_stream = [&](int s) {
      s_cur = s;
      nm = KB > 6 ? 0 : seg.nm[s];
      n_rows = seg.row_off[s + 1] - seg.row_off[s];
      seg_id = seg.seg_begin[s] + (cluster - seg.u_cfirst[s]) * 2 + (int)cta;
      cb = seg_id;
      boot = true;
      if (KB <= CAND_KMAX && w == 0 && lane == 0) *reinterpret_cast<volatile int*>(&cs->cur_bag) = s;
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        l_run[k] = 0.f;
        m_ref[k] = -INFINITY;
        c_ref[k] = INFINITY;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    };
    auto ref_scale = [](float a, float b) { return a == -INFINITY ? 0.f : ex2_approx((a - b) * LOG2E); };
    auto raise_ref = [&](int k, float v) {
      const float f = ref_scale(m_ref[k], v);
      l_run[k] *= f;
      m_ref[k] = v;
      c_ref[k] = PSH - v * LOG2E;
      if (cp == (k >> 1)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j][k & 1] *= f;
          acc[j][2 + (k & 1)] *= f;
        }
      }
    };
    // end of a bag (the 4 pool warps call this together)
    auto flush_stream = [&]() {
      if (s_cur < 0) return;
      if (nm > 0) {
        const float* rsc = reinterpret_cast<const float*>(p.mp.ws + p.mp.wl.rec_score) + (size_t)cb * K * rcap;
        const int* rix = reinterpret_cast<const int*>(p.mp.ws + p.mp.wl.rec_idx) + (size_t)cb * K * rcap;
        const int* rsl = reinterpret_cast<const int*>(p.mp.ws + p.mp.wl.rec_slot) + (size_t)cb * K * rcap;
        const float* rh = reinterpret_cast<const float*>(p.mp.ws + p.mp.wl.cand_h) + (size_t)cb * rowcap * L;
        asm volatile("bar.sync 1, 128;" ::: "memory");     // every pool warp is past the bag's last tile: no more appends
        if (w == 0 && lane == 0) *reinterpret_cast<volatile int*>(&cs->flush_req) = epoch_cur;
        while (*reinterpret_cast<volatile int*>(&cs->flush_ack) != epoch_cur) __nanosleep(50);     // manager caught up
        for (int k = w; k < K; k += 4) {     // publish which records are still in list k, hand the list to the reduce kernel
          const float mg_s = cs->ls[k][lane];
          const int mg_rec = cs->lrec[k][lane], mg_cnt = cs->cnt[k];
#pragma unroll
          for (int i = 0; i < REC_CAP / 32; ++i) {
            const unsigned m = __reduce_or_sync(0xffffffffu, (lane < mg_cnt && (mg_rec >> 5) == i) ? (1u << (mg_rec & 31)) : 0u);
            if (lane == 0) cs->active[k][i] = m;
          }
          int* g_cnt = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_cnt) + (size_t)cb * K;
          float* g_score = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_score) + (size_t)cb * K * cap;
          int* g_idx = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_idx) + (size_t)cb * K * cap;
          int* g_slot = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_slot) + (size_t)cb * K * cap;
          if (lane == 0) g_cnt[k] = mg_cnt;
          if (lane < cap) {
            const bool live = lane < mg_cnt;
            g_score[k * cap + lane] = live ? mg_s : -INFINITY;
            g_idx[k * cap + lane] = live ? rix[(size_t)k * rcap + mg_rec] : 0x7fffffff;
            g_slot[k * cap + lane] = live ? rsl[(size_t)k * rcap + mg_rec] : 0;
          }
        }
        __threadfence_block();
        asm volatile("bar.sync 1, 128;" ::: "memory");     // active[] is published
        // parked rows that did not stay in the CTA's top n rejoin the sums: warp w takes records w, w + 4, ... -- lane i
        // fetches score / h slot of one record (one round trip per 128 records), the reference moves once, then the h
        // rows are streamed two records at a time
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          if (k < K) {
            const int app = min(cs->app[k], rcap);
            for (int base = 0; base < app; base += 128) {
              const int rec = base + w + 4 * lane;
              const bool back = rec < app && !((cs->active[k][rec >> 5] >> (rec & 31)) & 1u);
              const float sc = back ? rsc[(size_t)k * rcap + rec] : -INFINITY;
              const int sl = back ? rsl[(size_t)k * rcap + rec] : 0;
              unsigned todo = __ballot_sync(0xffffffffu, back);
              if (todo) {
                const float mx = warp_max(sc);
                if (mx > m_ref[k] + REF_SLACK) raise_ref(k, mx);
                const float wgt = back ? ex2_approx(fmaf(sc, LOG2E, c_ref[k])) : 0.f;
                l_run[k] += wgt;       // per-lane partials; the lanes are folded below
                while (todo) {
                  const int s0 = __ffs(todo) - 1;
                  todo &= todo - 1;
                  const int s1 = todo ? __ffs(todo) - 1 : s0;
                  const bool two = todo != 0u;
                  todo &= todo - 1;
                  const float w0 = __shfl_sync(0xffffffffu, wgt, s0), w1 = two ? __shfl_sync(0xffffffffu, wgt, s1) : 0.f;
                  const int q0 = __shfl_sync(0xffffffffu, sl, s0), q1 = __shfl_sync(0xffffffffu, sl, s1);
                  if (cp == (k >> 1)) {
                    const float* h0 = rh + (size_t)q0 * L + rg;
                    const float* h1 = rh + (size_t)q1 * L + rg;
                    float v0[16], v1[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { v0[j] = __ldcg(h0 + 8 * j); v1[j] = __ldcg(h1 + 8 * j); }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      acc[j][k & 1] = fmaf(w0, v0[2 * j], acc[j][k & 1]);
                      acc[j][2 + (k & 1)] = fmaf(w0, v0[2 * j + 1], acc[j][2 + (k & 1)]);
                      acc[j][k & 1] = fmaf(w1, v1[2 * j], acc[j][k & 1]);
                      acc[j][2 + (k & 1)] = fmaf(w1, v1[2 * j + 1], acc[j][2 + (k & 1)]);
                    }
                  }
                }
              }
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // everybody has read app / active: reset the lists for the next bag
        for (int k = w; k < K; k += 4) {
          cs->ls[k][lane] = INFINITY;
          if (lane == 0) { cs->cnt[k] = 0; cs->app[k] = 0; cs->seen[k] = 0; cs->tau[k] = -INFINITY; }
        }
        if (w == 0 && lane == 0) cs->rows = 0;
      } else if (cap > 0 && w == 0 && lane < K) {
        reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_cnt)[(size_t)cb * K + lane] = 0;
      }
      // fold the lanes' l partials, then tree-merge the 4 warps' {m_ref, l, acc} through shared memory (tbuf + ps are
      // idle here) with the log-sum-exp rule: one record per CTA and bag
#pragma unroll
      for (int k = 0; k < KB; ++k) l_run[k] = warp_sum(l_run[k]);
      {
        float* xbuf = reinterpret_cast<float*>(smem + sm.tbuf);          // 11 KB: tbuf (7 KB) + ps (4 KB)
        constexpr int PF = KB * 130;                                     // floats of one warp partial
        static_assert(2 * PF * 4 <= 11264, "merge buffer");
        asm volatile("bar.sync 1, 128;" ::: "memory");                   // everybody is done with ps
#pragma unroll 1
        for (int stride = 2; stride >= 1; stride >>= 1) {
          if (w >= stride && w < 2 * stride) {
            float* slot = xbuf + (w - stride) * PF;
#pragma unroll
            for (int k = 0; k < KB; ++k) {
              if (lane == 0) {
                slot[k * 130] = m_ref[k];
                slot[k * 130 + 1] = l_run[k];
              }
              if (cp == (k >> 1)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  slot[k * 130 + 2 + 16 * j + rg] = acc[j][k & 1];
                  slot[k * 130 + 2 + 16 * j + rg + 8] = acc[j][2 + (k & 1)];
                }
              }
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (w < stride) {
            const float* slot = xbuf + w * PF;
#pragma unroll
            for (int k = 0; k < KB; ++k) {
              const float mo = slot[k * 130], mn = fmaxf(m_ref[k], mo);
              const float f_me = ref_scale(m_ref[k], mn), f_o = ref_scale(mo, mn);
              l_run[k] = l_run[k] * f_me + slot[k * 130 + 1] * f_o;
              m_ref[k] = mn;
              if (cp == (k >> 1)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  acc[j][k & 1] = acc[j][k & 1] * f_me + slot[k * 130 + 2 + 16 * j + rg] * f_o;
                  acc[j][2 + (k & 1)] = acc[j][2 + (k & 1)] * f_me + slot[k * 130 + 2 + 16 * j + rg + 8] * f_o;
                }
              }
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
      if (w == 0) {
        float* part = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.part) + (size_t)seg_id * K * (L + 2);
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          if (k < K) {
            if (lane == 0) {
              part[(size_t)k * (L + 2) + 0] = m_ref[k] - PSH * 0.6931471805599453f;   // l, acc are sums of exp(s - this)
              part[(size_t)k * (L + 2) + 1] = l_run[k];
            }
            if (cp == (k >> 1)) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                part[(size_t)k * (L + 2) + 2 + 16 * j + rg] = acc[j][k & 1];
                part[(size_t)k * (L + 2) + 2 + 16 * j + rg + 8] = acc[j][2 + (k & 1)];
              }
            }
          }
        }
      }
    };

    PROF_DECL();
#if GP_UMMA_PROF
    const long long t_start = clock64();
#endif
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const TilePos tp = tile_pos(g0 + t, s_hint);
      if (tp.s != s_cur) {
        { PROF_T0(); flush_stream(); PROF_ADD(5); }
        reset_stream(tp.s);
      }
      const int b = t % 3;
      const int64_t row = tp.row_in_bag + row_local;
      const bool valid = row < n_rows;

      // ---------------- scores of my row: the two G halves + the score bias; hand the tile back ----------------
      { PROF_T0(); mbar_wait(&bars->sc_full, (uint32_t)t & 1u); PROF_ADD(2); }
#if GP_UMMA_PROF
      const long long t_sp = clock64();
#endif
      float s[KB];
      {
        const float4 s0 = *reinterpret_cast<const float4*>(scp), s1 = *reinterpret_cast<const float4*>(scp + 4);
        *reinterpret_cast<float4*>(scp) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(scp + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int k = 0; k < KB; ++k) s[k] = sv[k] + p.c.bw[k];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->sc_empty);
      if (p.mp.a_out != nullptr && valid) {
        float* ao = p.mp.a_out + seg.row_off[s_cur] + row;
#pragma unroll
        for (int k = 0; k < KB; ++k)
          if (k < K) ao[(size_t)k * p.mp.a_ld] = s[k];
      }

      // ---------------- candidates: rows that beat the CTA's n-th best score are parked ----------------
      unsigned ex = 0u;       // bit k: my row is parked for branch k ...
      int myslot = 0;         // ... in this h row slot of the holder
      if (KB <= CAND_KMAX && nm > 0 && boot) {
        // First tile of the bag in this CTA: the lists are empty, every row would be a candidate.  Warp w picks the
        // tile's top n of branches w and w + 4 in one go (descending order: exactly n insertions each).
        float* rsc = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.rec_score) + (size_t)cb * K * rcap;
        int* rix = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.rec_idx) + (size_t)cb * K * rcap;
        int* rsl = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.rec_slot) + (size_t)cb * K * rcap;
        float* g_score = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_score) + (size_t)cb * K * cap;
        const float* sc_all = reinterpret_cast<const float*>(smem + sm.ps);      // [4 warps][32 rows][8]
        unsigned char* bflag = smem + sm.tbuf;                                   // [128 rows][8]: 0 or record + 1
        {
          float4 s0, s1;
          s0.x = s[0];
          s0.y = KB > 1 ? s[KB > 1 ? 1 : 0] : 0.f;
          s0.z = KB > 2 ? s[KB > 2 ? 2 : 0] : 0.f;
          s0.w = KB > 3 ? s[KB > 3 ? 3 : 0] : 0.f;
          s1.x = KB > 4 ? s[KB > 4 ? 4 : 0] : 0.f;
          s1.y = KB > 5 ? s[KB > 5 ? 5 : 0] : 0.f;
          s1.z = s1.w = 0.f;
          *reinterpret_cast<float4*>(psw + lane * 8) = s0;
          *reinterpret_cast<float4*>(psw + lane * 8 + 4) = s1;
          *reinterpret_cast<uint2*>(bflag + row_local * 8) = make_uint2(0u, 0u);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int k = w; k < K; k += 4) {
          float sv[4];
          unsigned todo = 0u;      // bit i: tile row lane + 32 i is still in the running
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            sv[i] = sc_all[i * 256 + lane * 8 + k];
            if (tp.row_in_bag + lane + 32 * i < n_rows) todo |= 1u << i;
          }
          float mg_s = INFINITY;
          int mg_rec = 0, mg_cnt = 0;
          for (int it = 0; it < nm; ++it) {
            unsigned key = 0u;
            int which = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const unsigned ki = ((todo >> i) & 1u) ? ord_enc(sv[i]) : 0u;
              if (ki > key) { key = ki; which = i; }
            }
            const unsigned best = __reduce_max_sync(0xffffffffu, key);
            if (best == 0u) break;
            // equal scores: the lower row wins, like torch.topk / the oracle (row = lane + 32 which)
            const int r = (int)__reduce_min_sync(0xffffffffu, key == best ? (unsigned)(lane + 32 * which) : 0xffffu);
            const int src = r & 31;
            if (lane == src) todo &= ~(1u << which);
            if (lane == it) { mg_s = ord_dec(best); mg_rec = it; }
            if (lane == 0) {
              rsc[(size_t)k * rcap + it] = ord_dec(best);
              rix[(size_t)k * rcap + it] = (int)tp.row_in_bag + r;
              bflag[r * 8 + k] = (unsigned char)(it + 1);
            }
            ++mg_cnt;
          }
          cs->ls[k][lane] = mg_s;
          cs->lrec[k][lane] = mg_rec;
          if (lane < cap) g_score[k * cap + lane] = lane < mg_cnt ? mg_s : -INFINITY;
          const float tau0 = ord_dec(__reduce_min_sync(0xffffffffu, ord_enc(mg_s)));
          unsigned* recs = reinterpret_cast<unsigned*>(smem + sm.tbuf + 1024) + k * rcap;
          for (int i = lane; i < rcap; i += 32) recs[i] = REC_EMPTY;       // (the manager starts behind the booted ones)
          if (lane == 0) {
            cs->cnt[k] = mg_cnt;
            cs->app[k] = mg_cnt;
            cs->seen[k] = mg_cnt;
            cs->tau[k] = mg_cnt == nm ? tau0 : -INFINITY;
          }
        }
        ++epoch_cur;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (w == 0 && lane == 0) *reinterpret_cast<volatile int*>(&cs->epoch) = epoch_cur;     // lists are live
        {
          // a selected row takes ONE h slot, which all its records (one per branch that selected it) point to
          const uint2 f = *reinterpret_cast<const uint2*>(bflag + row_local * 8);
          if ((f.x | f.y) != 0u) {
            myslot = atomicAdd(&cs->rows, 1);
#pragma unroll
            for (int k = 0; k < KB; ++k) {
              const unsigned v = ((k < 4 ? f.x : f.y) >> (8 * (k & 3))) & 0xffu;
              if (v) { ex |= 1u << k; rsl[(size_t)k * rcap + (int)v - 1] = myslot; }
            }
          }
        }
      } else if (KB <= CAND_KMAX && nm > 0) {
        float* rsc = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.rec_score) + (size_t)cb * K * rcap;
        int* rix = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.rec_idx) + (size_t)cb * K * rcap;
        int* rsl = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.rec_slot) + (size_t)cb * K * rcap;
        unsigned* recs = reinterpret_cast<unsigned*>(smem + sm.tbuf + 1024);
        unsigned hits = 0u;
#pragma unroll
        for (int k = 0; k < (KB <= CAND_KMAX ? KB : 0); ++k) {
          if (k < K) {
            const unsigned long long gq = *reinterpret_cast<volatile unsigned long long*>(&cs->gtau[k]);
            const float tau = fmaxf(*reinterpret_cast<volatile float*>(&cs->tau[k]),
                                    (unsigned)(gq >> 32) == (unsigned)s_cur ? __uint_as_float((unsigned)gq) : -INFINITY);
            if (valid && s[k] > tau) hits |= 1u << k;
          }
        }
        if (hits != 0u) {
          // one h slot per parked row, one record per (row, branch).  Out of slots / records: the bag is flagged and
          // redone by the exact FFMA kernel (rescue launch); what this kernel computes for it no longer matters
          myslot = atomicAdd(&cs->rows, 1);
          if (myslot >= rowcap) {
            reinterpret_cast<volatile int*>(p.mp.ws + p.mp.wl.flags)[s_cur] = 1;
            hits = 0u;
          }
#pragma unroll
          for (int k = 0; k < (KB <= CAND_KMAX ? KB : 0); ++k) {
            if ((hits >> k) & 1u) {
              const int rec = atomicAdd(&cs->app[k], 1);
              if (rec < rcap) {
                rsc[(size_t)k * rcap + rec] = s[k];
                rix[(size_t)k * rcap + rec] = (int)row;
                rsl[(size_t)k * rcap + rec] = myslot;
                *reinterpret_cast<volatile unsigned*>(&recs[k * rcap + rec]) = __float_as_uint(s[k]);
              } else {
                reinterpret_cast<volatile int*>(p.mp.ws + p.mp.wl.flags)[s_cur] = 1;
                hits &= ~(1u << k);
              }
            }
          }
          ex = hits;
        }
      }
      boot = false;
      // running reference: a taken row more than REF_SLACK above m_ref moves it (first tile of a stream: from -inf)
      {
        float tmax[KB];
        bool grow = false;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const bool take = valid && !((ex >> k) & 1u);
          tmax[k] = take ? s[k] : -INFINITY;
          grow |= tmax[k] > m_ref[k] + REF_SLACK;
        }
        if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
          for (int k = 0; k < KB; ++k) {
            const float v = warp_max(tmax[k]);
            if (v > m_ref[k] + REF_SLACK) raise_ref(k, v);
          }
        }
      }
      // softmax numerators p' = 2^(s log2e + c_ref)  (<= 2^15 by construction; parked / out-of-range rows: 0)
      {
        float pr[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) pr[k] = 0.f;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const bool take = valid && !((ex >> k) & 1u);
          pr[k] = take ? ex2_approx(fmaf(s[k], LOG2E, c_ref[k])) : 0.f;
          l_run[k] += pr[k];       // per-lane partials (one row per lane); folded at the end of the bag
        }
        *reinterpret_cast<float4*>(psw + lane * 8) = make_float4(pr[0], pr[1], pr[2], pr[3]);
        *reinterpret_cast<float4*>(psw + lane * 8 + 4) = make_float4(pr[4], pr[5], pr[6], pr[7]);
      }
      __syncwarp();
#if GP_UMMA_PROF
      prof[3] += clock64() - t_sp;
      const long long t_pool = clock64();
#endif
      // ---------------- pool: acc[feature][branch] += sum_rows h[row][feature] p'[row][branch] ----------------
      // per 16-row half of the warp's rows: the K dimension of mma.sync m16n8k16 with A = h^T (movmatrix transposes of
      // the packed fp16 h operand as tcgen05.ld 16x128b hands it out) and B = p'; hi/lo split as in the big GEMMs
      float* cand_h = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_h) + (size_t)cb * rowcap * L;
      const bool any_ex = __any_sync(0xffffffffu, ex != 0u);
      // eight steps (2 row halves x 4 feature chunks), software-pipelined: the TMEM loads of step i + 1 are in flight
      // while step i runs its transposes and MMAs
      {
        uint32_t hh[2][8], hl[2][8];     // regs {2i, 2i+1}: rows rg / rg + 8, features 32 c + 8 i + 2 cp + {0,1}
        auto issue_loads = [&](int step, uint32_t (&dh)[8], uint32_t (&dl)[8]) {
          const uint32_t hbase = tm + ((uint32_t)(w * 32 + (step >> 2) * 16) << 16) + tm3_dh(b) + 32 * (step & 3);
          tmem_ld_16x128b_x4(hbase, dh);
          tmem_ld_16x128b_x4(hbase + 16, dl);
        };
        uint32_t bh[2] = {0u, 0u}, bl[2] = {0u, 0u};      // B fragments of the current row half
        float* t0 = nullptr;      // parked rows of this row group -> their h slots
        float* t1 = nullptr;
        constexpr int NSTEP = GP_EXP_NOEPI ? 0 : 8;
        if (NSTEP > 0) issue_loads(0, hh[0], hl[0]);
#pragma unroll
        for (int step = 0; step < NSTEP; ++step) {
          const int hh2 = step >> 2, c = step & 3, cur = step & 1;
          if (c == 0) {      // {p'[2cp][rg], p'[2cp+1][rg]}, {p'[2cp+8][rg], p'[2cp+9][rg]} as fp16 hi / lo
            const float* pw = psw + hh2 * 128;
            const float q0 = pw[(2 * cp) * 8 + rg], q1 = pw[(2 * cp + 1) * 8 + rg];
            const float q2 = pw[(2 * cp + 8) * 8 + rg], q3 = pw[(2 * cp + 9) * 8 + rg];
            split2(q0, q1, bh[0], bl[0]);
            split2(q2, q3, bh[1], bl[1]);
            t0 = t1 = nullptr;
            if (any_ex) {
              const unsigned e0 = __shfl_sync(0xffffffffu, ex, hh2 * 16 + rg), e1 = __shfl_sync(0xffffffffu, ex, hh2 * 16 + rg + 8);
              const int sl0 = __shfl_sync(0xffffffffu, myslot, hh2 * 16 + rg), sl1 = __shfl_sync(0xffffffffu, myslot, hh2 * 16 + rg + 8);
              if (e0) t0 = cand_h + (size_t)sl0 * L;
              if (e1) t1 = cand_h + (size_t)sl1 * L;
            }
          }
          tmem_wait_ld();
          if (step + 1 < NSTEP) issue_loads(step + 1, hh[cur ^ 1], hl[cur ^ 1]);
          if (t0 != nullptr || t1 != nullptr) {   // park the rows that entered a list this tile (fp32 h = hi + lo)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 ah = __half22float2(*reinterpret_cast<const __half2*>(&hh[cur][2 * i]));
              const float2 al = __half22float2(*reinterpret_cast<const __half2*>(&hl[cur][2 * i]));
              const float2 bhf = __half22float2(*reinterpret_cast<const __half2*>(&hh[cur][2 * i + 1]));
              const float2 blf = __half22float2(*reinterpret_cast<const __half2*>(&hl[cur][2 * i + 1]));
              const int fcol = c * 32 + i * 8 + cp * 2;
              if (t0 != nullptr) *reinterpret_cast<float2*>(t0 + fcol) = make_float2(ah.x + al.x, ah.y + al.y);
              if (t1 != nullptr) *reinterpret_cast<float2*>(t1 + fcol) = make_float2(bhf.x + blf.x, bhf.y + blf.y);
            }
          }
          uint32_t ahi[2][4], alo[2][4];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            ahi[jj][0] = movmatrix_t(hh[cur][4 * jj]);
            ahi[jj][1] = movmatrix_t(hh[cur][4 * jj + 2]);
            ahi[jj][2] = movmatrix_t(hh[cur][4 * jj + 1]);
            ahi[jj][3] = movmatrix_t(hh[cur][4 * jj + 3]);
            alo[jj][0] = movmatrix_t(hl[cur][4 * jj]);
            alo[jj][1] = movmatrix_t(hl[cur][4 * jj + 2]);
            alo[jj][2] = movmatrix_t(hl[cur][4 * jj + 1]);
            alo[jj][3] = movmatrix_t(hl[cur][4 * jj + 3]);
          }
#if GP_EXP_NO_HMMA      // timing experiment only: how much tensor-pipe time do the legacy mma.sync instructions take?
          acc[c * 2][0] += __uint_as_float(ahi[0][0] ^ alo[0][1] ^ bh[0] ^ ahi[1][2] ^ alo[1][3] ^ bl[1]);
#else
          // the two accumulators of a step alternate, so that consecutive mma.sync never wait for each other
          mma_16816_f16(acc[c * 2], ahi[0], bh);
          mma_16816_f16(acc[c * 2 + 1], ahi[1], bh);
          mma_16816_f16(acc[c * 2], alo[0], bh);
          mma_16816_f16(acc[c * 2 + 1], alo[1], bh);
          mma_16816_f16(acc[c * 2], ahi[0], bl);
          mma_16816_f16(acc[c * 2 + 1], ahi[1], bl);
#endif
        }
      }
      tc_fence_before();
      __syncwarp();      // psw is rewritten by the next tile; the h operand has been read
      if (lane == 0) mbar_arrive_cluster(&bars->dh_free[b], 0);
#if GP_UMMA_PROF
      prof[4] += clock64() - t_pool;
#endif
    }
    { PROF_T0(); flush_stream(); PROF_ADD(5); }
    if (KB <= CAND_KMAX && w == 0 && lane == 0) *reinterpret_cast<volatile int*>(&cs->cur_bag) = -2;     // stops the service warps
#if GP_UMMA_PROF
    prof[7] = clock64() - t_start;
    if (lane == 0) PROF_FLUSH();
#endif
  }

  // ---- teardown ----
  tc_fence_before();
  cluster_sync();
  if (warp == 2) tmem_dealloc<2>(tm, 512);
}

template <int KB>
int launch3_kb(const UmmaParams& up, int grid, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    ACMIL_CHECK_CUDA(cudaFuncSetAttribute(gp_main_umma3_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  gp_main_umma3_kernel<KB><<<grid, UT3, smem, st>>>(up);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

}  // namespace

#if GP_UMMA_PROF
extern "C" __attribute__((visibility("default"))) int acmil_debug_umma3_prof(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_umma3_prof, sizeof(long long) * n);
}
#endif

int gp_launch_main_umma3(const UmmaParams& up, int n_branch, int grid, cudaStream_t st) {
  const int K = n_branch;
  const size_t smem = smem_map3(up.mp.sh.d_in, K == 1 ? 1 : (K <= 5 ? 5 : 8)).total + 1024;
  if (K == 1) return launch3_kb<1>(up, grid, smem, st);
  if (K <= 5) return launch3_kb<5>(up, grid, smem, st);
  return launch3_kb<8>(up, grid, smem, st);
}
