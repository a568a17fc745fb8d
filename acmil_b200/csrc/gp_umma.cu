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
//  * TMEM (512 columns): two D1 h-accumulators of 128 columns, each rewritten IN PLACE by Epi1 as the packed fp16
//    hi/lo h operand of the gate GEMM (per 64-feature half: 32 columns hi, 32 columns lo), so the projection of tile
//    t + 1 runs while the epilogue is still working on tile t | D2 gate accumulator 128 (two 64-column buffers, four
//    quarters per tile) | x operand ring 4 x 32.
//
// Roles per CTA (512 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA; warp-uniform code with
// elect.sync around the tcgen05 instructions), warp 2 TMEM allocator + top-n list manager, warp 3 bag-wide
// threshold service, warps 4-7 converters (thread = row), warps 8-15 epilogue (16 rows per warp through the
// 16-lane TMEM shapes; each warp keeps a private softmax stream {m_ref, l, acc} whose pool step runs on
// mma.sync fed by movmatrix transposes of the TMEM h operand, so there is no cross-warp traffic per tile;
// rows that may be in a branch's top n are appended to scratch records instead of being summed).
//
// Scratch in shared memory ("tbuf", 7 KB + "ps", 4 KB): per tile ps holds each warp's softmax numerators
// [16 rows][8]; tbuf holds the first-tile selection flags [128][8] (bytes 0-1023) and the appended record
// scores [K][rec_cap] (from byte 1024); at a bag's end both are reused as the 11 KB buffer of the 8-warp merge.
#include <stdlib.h>

#include <type_traits>

#include "gp_umma_shared.cuh"

namespace {
using namespace sm100;
using namespace umma_shared;

constexpr int UT = 512;
constexpr int NXOP = 4;      // TMEM x-operand ring
// D1 / h buffer b at columns 128 b; inside it, after Epi1: half hf (features 64 hf ...) hi at 64 hf, lo at 64 hf + 32
constexpr uint32_t TM_DH = 0, TM_D2 = 256, TM_X = 384;

#ifndef GP_UMMA_PROF
#define GP_UMMA_PROF 0
#endif
// polling periods (ns) of the two service warps: they share their schedulers with epilogue warps
#ifndef GP_MGR_SLEEP
#define GP_MGR_SLEEP 400
#endif
#ifndef GP_SVC_SLEEP
#define GP_SVC_SLEEP 2000
#endif
#if GP_UMMA_PROF
__device__ long long g_umma_prof[148][88];      // 0-7 TMA, 8-15 MMA, 16-23 converter, 24 + 8 e: epilogue warp e
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

// gate constants in smem, one record of 2 CREC floats per pair of adjacent units (2c, 2c+1), laid out as fp32 pairs
// for the packed FFMA2 path: {ww[k][2c], ww[k][2c+1]} for k < KB, then the bv' pair and the bu' pair, where
// bv' = -2 log2e bv and bu' = -log2e bu are the biases in the exponent domain
__host__ __device__ constexpr int cst_rec(int kb) { return kb <= 2 ? 4 : (kb <= 6 ? 8 : 10); }

// smem carve-up (offsets from the 1024-aligned base)
struct SmemMap {
  uint32_t w1, wg, stage, tbuf, ps, cst, cand, bars, total;
};

__host__ __device__ inline SmemMap smem_map(int din, int kb) {
  SmemMap m;
  m.w1 = 0;
  m.wg = (uint32_t)(din / 64) * 8192u * 2u;
  m.stage = m.wg + 65536u;
  m.tbuf = m.stage + NSTAGE * STAGE_BYTES;
  m.ps = m.tbuf + 7 * 1024;        // tbuf + ps = 11 KB scratch of the bag-end merge
  m.cst = m.ps + 8 * 512;
  m.cand = m.cst + 128 * (uint32_t)cst_rec(kb) * 4;
  m.bars = m.cand + (kb > 6 ? 128u : (uint32_t)sizeof(CandShared));   // K > 6: no masking on this kernel
  m.total = m.bars + 256;
  return m;
}

struct Bars {
  uint64_t full_x[NSTAGE], empty_x[NSTAGE];
  uint64_t xop_full[NXOP], xop_empty[NXOP];
  uint64_t d1_full[2], d1_empty[2], hop_full, d2_full[2], d2_empty[2], wload, w_ready;
  uint32_t tmem_base;
};

template <int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UT, 1) gp_main_umma_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need a 1024-byte aligned base (same offset in both CTAs of the pair)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int DIN = p.mp.sh.d_in;
  // fp32 rows: 32-column chunks (128 B per row), split into hi / lo halves by the converters; fp16 rows (x_f16): the
  // tile IS the hi operand and lo == 0, so a chunk is 64 columns (the same 128 B per row) and its x_lo MMAs are skipped
  const bool XH = p.mp.x_f16 != 0;
  const int NCH = XH ? DIN / 64 : DIN / KC;
  const SmemMap sm = smem_map(DIN, KB);
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
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->d1_full[i], 1); mbar_init(&bars->d1_empty[i], 16); }
    mbar_init(&bars->hop_full, 16);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->d2_full[i], 1); mbar_init(&bars->d2_empty[i], 16); }
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
    for (int u = tid; u < 128; u += UT) {      // record of the unit pair (2c, 2c+1): {ww[k][2c], ww[k][2c+1]}_k, bv' pair, bu' pair
      float* rec = cst + (u >> 1) * (2 * CREC) + (u & 1);
#pragma unroll
      for (int k = 0; k < KB; ++k) rec[2 * k] = p.dc->ww[k][u];
      rec[2 * KB] = p.dc->bv[u] * (-2.f * LOG2E);
      rec[2 * KB + 2] = p.dc->bu[u] * (-LOG2E);
    }
  }
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

  // NOTE: each setmaxnreg sits at the top of a branch that never rejoins the others before the kernel's
  // tail, otherwise ptxas allocates the whole kernel for the smallest budget.
  if (warp < 4) {
  setmaxnreg_dec<80>();
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
          tma_load_2d_hint(smem + sm.stage + st * STAGE_BYTES, &p.tmap, c * (XH ? 64 : KC), (int)grow, &bars->full_x[st], kEvictFirst);
        }
      }
      PROF_FLUSH(0);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA) =====================================
    // The whole warp runs the scheduler on identical values; only the tcgen05 instructions sit under elect_one().
    if (cta == 0 && T > 0) {
      PROF_DECL();
      const long long t_start = clock64();
      { PROF_T0(); mbar_wait_cluster(&bars->w_ready, 0); PROF_ADD(0); }  // both CTAs' weight images have landed
      const uint32_t idesc = umma_idesc_f16(256, 128);
      const uint32_t w1_hi = smem_u32(smem + sm.w1), w1_lo = w1_hi + p.w1_part_bytes;
      const uint32_t wg_hi = smem_u32(smem + sm.wg), wg_lo = wg_hi + 32768u;
      uint32_t xc = 0;  // x-operand chunks consumed so far (ring position / phase)
      // chunk c of tile t of the projection GEMM; returns false (nothing issued) when its inputs are not there yet
      auto g1_try = [&](int t, int c) -> bool {
        // buffer t & 1 was tile t - 2's accumulator, gate operand and pool operand: free once that tile's pool is done
        if (c == 0 && t > 1 && !mbar_test_wait(&bars->d1_empty[t & 1], (uint32_t)((t >> 1) - 1) & 1u)) return false;
        const uint32_t q = xc % NXOP, ph = (xc / NXOP) & 1u;
        if (!mbar_test_wait(&bars->xop_full[q], ph)) return false;
        tc_fence_after();
        const uint32_t xa_hi = tm + TM_X + q * 32, xa_lo = xa_hi + 16;
        const uint32_t boff = (uint32_t)(c >> 1) * 8192u + (uint32_t)(c & 1) * 64u;
        if (XH) {      // fp16 rows: chunk c = W1 k-block c, four k-steps of 16, hi operand only
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t bo = (uint32_t)c * 8192u + (uint32_t)ks * 32u;
              const uint64_t bhi = umma_desc_k_sw128(w1_hi + bo), blo = umma_desc_k_sw128(w1_lo + bo);
              umma_ts<2>(tm + TM_DH + (uint32_t)(t & 1) * 128u, xa_hi + ks * 8, bhi, idesc, (c | ks) ? 1u : 0u);
              umma_ts<2>(tm + TM_DH + (uint32_t)(t & 1) * 128u, xa_hi + ks * 8, blo, idesc, 1u);
            }
            umma_commit_2sm(&bars->xop_empty[q], 3);
            if (c == NCH - 1) umma_commit_2sm(&bars->d1_full[t & 1], 3);
          }
        } else if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t bhi = umma_desc_k_sw128(w1_hi + boff + ks * 32), blo = umma_desc_k_sw128(w1_lo + boff + ks * 32);
            umma_ts<2>(tm + TM_DH + (uint32_t)(t & 1) * 128u, xa_hi + ks * 8, bhi, idesc, (c | ks) ? 1u : 0u);
            umma_ts<2>(tm + TM_DH + (uint32_t)(t & 1) * 128u, xa_lo + ks * 8, bhi, idesc, 1u);
            umma_ts<2>(tm + TM_DH + (uint32_t)(t & 1) * 128u, xa_hi + ks * 8, blo, idesc, 1u);
          }
          umma_commit_2sm(&bars->xop_empty[q], 3);
          if (c == NCH - 1) umma_commit_2sm(&bars->d1_full[t & 1], 3);
        }
        __syncwarp();
        ++xc;
        return true;
      };
      const uint32_t idesc64 = umma_idesc_f16(256, 64);
      // gate GEMM in four 32-unit quarters (N = 64: 32 V + 32 U columns) ping-ponging between two D2 buffers,
      // so the epilogue works on one quarter while the tensor core produces the next
      auto g2_try = [&](int t, int qr) -> bool {
        const int b = qr & 1;
        if (qr == 0 && !mbar_test_wait(&bars->hop_full, (uint32_t)t & 1u)) return false;
        const uint32_t n_use = 2u * (uint32_t)t + (uint32_t)(qr >> 1);      // how often buffer b was used before
        if (n_use > 0 && !mbar_test_wait(&bars->d2_empty[b], (n_use - 1u) & 1u)) return false;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t boff = (uint32_t)(qr * 2 + (ks >> 2)) * 4096u + (uint32_t)(ks & 3) * 32u;
            const uint64_t bhi = umma_desc_k_sw128(wg_hi + boff), blo = umma_desc_k_sw128(wg_lo + boff);
            const uint32_t ha = tm + TM_DH + (uint32_t)(t & 1) * 128u + (uint32_t)(ks >> 2) * 64u + (uint32_t)(ks & 3) * 8u;
            umma_ts<2>(tm + TM_D2 + b * 64, ha, bhi, idesc64, ks ? 1u : 0u);
            umma_ts<2>(tm + TM_D2 + b * 64, ha + 32u, bhi, idesc64, 1u);
            umma_ts<2>(tm + TM_D2 + b * 64, ha, blo, idesc64, 1u);
          }
          umma_commit_2sm(&bars->d2_full[b], 3);
        }
        __syncwarp();
        return true;
      };
      // issue order: the gate quarters of tile t have priority (the epilogue is waiting for them); chunks of the
      // next tile's projection fill the gaps while the epilogue drains a D2 buffer
      for (int c = 0; c < NCH;) c += g1_try(0, c) ? 1 : 0;
      for (int t = 0; t < T; ++t) {
        int qr = 0, c = 0;
        const int nc = (t + 1 < T) ? NCH : 0;
        while (qr < 4 || c < nc) {
#if GP_UMMA_PROF
          const long long t_it = clock64();
#endif
          if (qr < 4 && g2_try(t, qr)) { ++qr; continue; }
          if (c < nc && g1_try(t + 1, c)) { ++c; continue; }
#if GP_UMMA_PROF
          {   // nothing could be issued: attribute the idle poll to what blocks each GEMM
            const long long dt = clock64() - t_it;
            if (c < nc) {
              if (c == 0 && t > 0 && !mbar_test_wait(&bars->d1_empty[(t + 1) & 1], (uint32_t)(((t + 1) >> 1) - 1) & 1u)) prof[1] += dt; else prof[2] += dt;
            } else prof[5] += dt;
            if (qr < 4) {
              if (qr == 0 && !mbar_test_wait(&bars->hop_full, (uint32_t)t & 1u)) prof[3] += dt; else prof[4] += dt;
            } else prof[6] += dt;
          }
#endif
        }
      }
#if GP_UMMA_PROF
      prof[7] = clock64() - t_start;
      if (lane == 0) PROF_FLUSH(8);
#endif
    }
    __syncwarp();
  } else if (warp == 2) {
    if (KB <= CAND_KMAX && seg.n_masked_cap > 0 && T > 0) {
      // ===================================== top-n list manager =====================================
      CandShared* cs = reinterpret_cast<CandShared*>(smem + sm.cand);
      const unsigned* recs = reinterpret_cast<const unsigned*>(smem + sm.tbuf + 1024);      // [K][rec_cap] score bits
      float* g_score_all = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_score);
      const int cap = seg.n_masked_cap, rcap = seg.rec_cap;
      int dead_epoch = 0;       // lists of this epoch are final (or not booted yet): hands off
      while (true) {
        bool busy = false;
        const int s = *reinterpret_cast<volatile int*>(&cs->cur_bag);
        if (s == -2) break;
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
        // idle polls back off; a pending flush request or fresh records are served right away
        if (!busy) __nanosleep(*reinterpret_cast<volatile int*>(&cs->flush_req) != dead_epoch ? 40 : GP_MGR_SLEEP);
      }
    }
  } else if (warp == 3) {
    // this CTA's resident weights are in place -> tell the leader's MMA thread
    if (lane == 0) {
      mbar_wait(&bars->wload, 0);
      mbar_arrive_cluster(&bars->w_ready, 0);
    }
    __syncwarp();
    if (KB <= CAND_KMAX && seg.n_masked_cap > 0 && T > 0) {
      // ===================================== bag-wide threshold service =====================================
      // merge the mirrored top-n lists of every CTA working on the current bag: the n-th best of their union is the
      // n-th best of all rows anybody has scored so far (any such row is in its own CTA's list), a lower bound of
      // the final one.  Entries not written in this launch are NaN (memset by the host) and ignored.
      CandShared* cs = reinterpret_cast<CandShared*>(smem + sm.cand);
      const float* g_score = reinterpret_cast<const float*>(p.mp.ws + p.mp.wl.cand_score);
      const int cap = seg.n_masked_cap;
      while (true) {
        const int s = *reinterpret_cast<volatile int*>(&cs->cur_bag);
        if (s == -2) break;
        const int nmk = s >= 0 ? seg.nm[s] : 0;
        if (nmk > 0) {
          const int seg0 = seg.seg_begin[s], E = (seg.seg_begin[s + 1] - seg0) * cap;
          for (int k = 0; k < K; ++k) {
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
              if (*reinterpret_cast<volatile int*>(&cs->cur_bag) != s) break;     // bag changed / kernel finishing
            }
            if (*reinterpret_cast<volatile int*>(&cs->cur_bag) != s) break;
            if (cnt == nmk && lane == 0)
              *reinterpret_cast<volatile unsigned long long*>(&cs->gtau[k]) = ((unsigned long long)(unsigned)s << 32) | __float_as_uint(tau);
          }
        }
        __nanosleep(GP_SVC_SLEEP);
      }
    }
  }
  } else if (warp < 8) {
    // ===================================== converters: fp32 staging -> fp16 hi/lo in TMEM =====================================
    setmaxnreg_dec<96>();
    const int r = (warp - 4) * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp - 4) * 32) << 16;
    uint32_t ctr = 0;
    PROF_DECL();
#if GP_UMMA_PROF
    const long long t_start = clock64();
    unsigned long long ns_start;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_start));
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
        if (XH) {      // the 128 bytes are 64 halves = the packed operand columns as they are (de-swizzled): no arithmetic
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            hi[4 * i] = __float_as_uint(v[i].x); hi[4 * i + 1] = __float_as_uint(v[i].y);
            hi[4 * i + 2] = __float_as_uint(v[i].z); hi[4 * i + 3] = __float_as_uint(v[i].w);
            lo[4 * i] = __float_as_uint(v[4 + i].x); lo[4 * i + 1] = __float_as_uint(v[4 + i].y);
            lo[4 * i + 2] = __float_as_uint(v[4 + i].z); lo[4 * i + 3] = __float_as_uint(v[4 + i].w);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            split2(v[i].x, v[i].y, hi[2 * i], lo[2 * i]);
            split2(v[i].z, v[i].w, hi[2 * i + 1], lo[2 * i + 1]);
          }
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
    {   // wall-clock span of the same loop: cycles / ns = the SM clock this kernel really ran at
      unsigned long long ns_end;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
      prof[6] = (long long)(ns_end - ns_start);
    }
    if (warp == 4 && lane == 0) PROF_FLUSH(16);
#endif
  } else {
    // ===================================== epilogue: 8 warps x 16 rows =====================================
    // Warp (q, hs) owns TMEM lanes / tile rows [32q + 16hs, +16) and reads them with the 16-lane tcgen05.ld
    // shapes: in 16x256b thread t gets rows ra = base + t/4 and rb = ra + 8, columns 8i + 2(t%4) + {0,1};
    // in 16x128b (the packed fp16 operand) it gets column 4i + t%4 of the same two rows -- i.e. the same
    // features.  Every warp is an independent stream (own m, l, acc, candidate lists): no cross-warp traffic.
    setmaxnreg_inc<168>();
    const int e_idx = warp - 8, q = e_idx & 3, hs = e_idx >> 2;
    const int lane_base = q * 32 + hs * 16;
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const int rg = lane >> 2, cp = lane & 3;
    float* psw = reinterpret_cast<float*>(smem + sm.ps + e_idx * 512);       // [16 rows][8]
    const float* cstp = reinterpret_cast<const float*>(smem + sm.cst);
    constexpr int CREC = cst_rec(KB);
    const int L = 128;
    const int cap = seg.n_masked_cap;
    const float cva = p.dc->inv_sv * (-2.f * LOG2E), cua = p.dc->inv_su * (-LOG2E), inv_s1 = p.dc->inv_s1;
    float bwk[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) bwk[k] = p.dc->bw[k];

    // Softmax-pool state of this warp's stream.  Numerators are p' = exp(s - m_ref[k]) * 2^PSH against a per-warp
    // reference m_ref[k] that only grows, and only when a taken row exceeds it by more than REF_SLACK nats (then l and
    // acc are rescaled: rare), so p' <= 2^15 always fits the fp16 hi/lo operands of the pool's tensor-core step.
    // acc is the set of mma.sync D fragments: acc[j][0..1] = feature 16j + rg, branches 2cp, 2cp + 1;
    // acc[j][2..3] = feature 16j + rg + 8.
    constexpr float PSH = 4.f, REF_SLACK = 7.6f;       // (7.6 log2e + 4 < 15)
    float l_run[KB], m_ref[KB], c_ref[KB], acc[8][4];     // l_run: per-lane partial of sum p'; c_ref = PSH - m_ref log2e
    CandShared* cs = reinterpret_cast<CandShared*>(smem + sm.cand);
    const int rcap = seg.rec_cap;
    int s_cur = -1, s_hint = 0, nm = 0, seg_id = 0, cb = 0;
    bool boot = false;     // the next tile is this CTA's first tile of the bag: the lists are empty
    int epoch_cur = 0;     // bags booted so far (same value in all epilogue warps)
    int64_t n_rows = 0;

    auto reset_stream = [&](int s) {
      s_cur = s;
      nm = KB > 6 ? 0 : seg.nm[s];
      n_rows = seg.row_off[s + 1] - seg.row_off[s];
      seg_id = seg.seg_begin[s] + (cluster - seg.u_cfirst[s]) * 2 + (int)cta;   // one segment per CTA and bag
      cb = seg_id;                     // ... which is also its candidate holder
      boot = true;
      if (KB <= CAND_KMAX && e_idx == 0 && lane == 0) *reinterpret_cast<volatile int*>(&cs->cur_bag) = s;
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        l_run[k] = 0.f;
        m_ref[k] = -INFINITY;
        c_ref[k] = INFINITY;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    };
    // exp(a - b) for references a <= b, with exp(-inf - anything) = 0
    auto ref_scale = [](float a, float b) { return a == -INFINITY ? 0.f : ex2_approx((a - b) * LOG2E); };
    // move branch k's reference up to v (warp-uniform; k is a compile-time constant at every call site)
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
    // end of a bag (all 8 epilogue warps call this together)
    auto flush_stream = [&]() {
      if (s_cur < 0) return;
      if (nm > 0) {
        const float* rsc = reinterpret_cast<const float*>(p.mp.ws + p.mp.wl.rec_score) + (size_t)cb * K * rcap;
        const int* rix = reinterpret_cast<const int*>(p.mp.ws + p.mp.wl.rec_idx) + (size_t)cb * K * rcap;
        const int* rsl = reinterpret_cast<const int*>(p.mp.ws + p.mp.wl.rec_slot) + (size_t)cb * K * rcap;
        const float* rh = reinterpret_cast<const float*>(p.mp.ws + p.mp.wl.cand_h) + (size_t)cb * seg.row_cap * L;
        asm volatile("bar.sync 1, 256;" ::: "memory");     // every warp is past the bag's last tile: no more appends
        if (e_idx == 0 && lane == 0) *reinterpret_cast<volatile int*>(&cs->flush_req) = epoch_cur;
        while (*reinterpret_cast<volatile int*>(&cs->flush_ack) != epoch_cur) __nanosleep(50);     // manager caught up
        if (e_idx < K) {     // warp k publishes which records are still in list k and hands the list to the reduce kernel
          const int k = e_idx;
          const float mg_s = cs->ls[k][lane];
          const int mg_rec = cs->lrec[k][lane], mg_cnt = cs->cnt[k];
          // Entries below the bag-wide threshold (a lower bound of the bag's final n-th best, from warp 3) cannot be in the
          // global top n: they rejoin the sums HERE, on this CTA, instead of travelling to the reduce kernel, where the
          // add-back of ~10 candidates per CTA and branch ran on one CTA per (bag, branch)
          const unsigned long long gq = *reinterpret_cast<volatile unsigned long long*>(&cs->gtau[k]);
          const float gt = (unsigned)(gq >> 32) == (unsigned)s_cur ? __uint_as_float((unsigned)gq) : -INFINITY;
          const bool keepf = lane < mg_cnt && !(mg_s < gt);
          const unsigned kbal = __ballot_sync(0xffffffffu, keepf);
          const int pos = __popc(kbal & ((1u << lane) - 1u)), nkeep = __popc(kbal);
#pragma unroll
          for (int w = 0; w < REC_CAP / 32; ++w) {
            const unsigned m = __reduce_or_sync(0xffffffffu, (keepf && (mg_rec >> 5) == w) ? (1u << (mg_rec & 31)) : 0u);
            if (lane == 0) cs->active[k][w] = m;
          }
          int* g_cnt = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_cnt) + (size_t)cb * K;
          float* g_score = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_score) + (size_t)cb * K * cap;
          int* g_idx = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_idx) + (size_t)cb * K * cap;
          int* g_slot = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_slot) + (size_t)cb * K * cap;
          if (lane == 0) g_cnt[k] = nkeep;
          if (keepf) {      // compacted: the reduce kernel reads the first g_cnt entries
            g_score[k * cap + pos] = mg_s;
            g_idx[k * cap + pos] = rix[(size_t)k * rcap + mg_rec];
            g_slot[k * cap + pos] = rsl[(size_t)k * rcap + mg_rec];
          }
          if (lane < cap && lane >= nkeep) {
            g_score[k * cap + lane] = -INFINITY;
            g_idx[k * cap + lane] = 0x7fffffff;
            g_slot[k * cap + lane] = 0;
          }
        }
        __threadfence_block();
        asm volatile("bar.sync 1, 256;" ::: "memory");     // active[] is published
        // parked rows that did not stay in the CTA's top n rejoin the sums.  Warp w takes records w, w + 8, ...: lane i
        // fetches score / h slot of record w + 8 i (one round trip for the whole list), the reference moves once, then
        // the h rows are streamed two records at a time (32 independent loads in flight per thread)
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          if (k < K) {
            const int app = min(cs->app[k], rcap);
            const int rec = e_idx + 8 * lane;
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
        asm volatile("bar.sync 1, 256;" ::: "memory");     // everybody has read app / active: reset the lists for the next bag
        if (e_idx < K) {
          cs->ls[e_idx][lane] = INFINITY;
          if (lane == 0) { cs->cnt[e_idx] = 0; cs->app[e_idx] = 0; cs->seen[e_idx] = 0; cs->tau[e_idx] = -INFINITY; }
          if (e_idx == 0 && lane == 0) cs->rows = 0;
        }
      } else if (cap > 0 && e_idx == 0 && lane < K) {
        reinterpret_cast<int*>(p.mp.ws + p.mp.wl.cand_cnt)[(size_t)cb * K + lane] = 0;
      }
      // fold the lanes' l partials, then tree-merge the 8 warps' {m_ref, l, acc} through shared memory (the buffers of
      // the tile loop are idle here) with the usual log-sum-exp rule: one record per CTA and bag
#pragma unroll
      for (int k = 0; k < KB; ++k) l_run[k] = warp_sum(l_run[k]);
      {
        float* xbuf = reinterpret_cast<float*>(smem + sm.tbuf);          // 11 KB: tbuf (7 KB) + ps (4 KB)
        constexpr int PF = KB * 130;                                     // floats of one warp partial: {m, l, acc[128]} per branch
        constexpr int MAXSLOT = (11264 / 4) / PF >= 4 ? 4 : ((11264 / 4) / PF >= 2 ? 2 : 1);
        asm volatile("bar.sync 1, 256;" ::: "memory");                   // everybody is done with ps
#pragma unroll 1
        for (int stride = 4; stride >= 1; stride >>= 1) {
#pragma unroll 1
          for (int base = 0; base < stride; base += MAXSLOT) {
            const int w_lo = stride + base, w_hi = stride + min(base + MAXSLOT, stride);
            if (e_idx >= w_lo && e_idx < w_hi) {
              float* slot = xbuf + (e_idx - w_lo) * PF;
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
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (e_idx >= base && e_idx < base + (w_hi - w_lo)) {
              const float* slot = xbuf + (e_idx - base) * PF;
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
            asm volatile("bar.sync 1, 256;" ::: "memory");
          }
        }
      }
      if (e_idx == 0) {
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
        { PROF_T0(); flush_stream(); PROF_ADD(6); }
        reset_stream(tp.s);
      }
      const int64_t row_a = tp.row_in_bag + lane_base + rg, row_b = row_a + 8;
      const bool valid_a = row_a < n_rows, valid_b = row_b < n_rows;

      // ---------------- Epi1: D1 -> relu -> fp16 hi/lo operand of the gate GEMM ----------------
      const uint32_t tm_dh = tm + lane_addr + TM_DH + (uint32_t)(t & 1) * 128u;
      { PROF_T0(); mbar_wait_cluster(&bars->d1_full[t & 1], (uint32_t)(t >> 1) & 1u); PROF_ADD(0); }
      tc_fence_after();
#if GP_UMMA_PROF
      const long long t_e1 = clock64();
#endif
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v[32];
        tmem_ld_16x256b_x8(tm_dh + hf * 64, v);
        tmem_wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          split2(fmaxf(__uint_as_float(v[4 * i]) * inv_s1, 0.f), fmaxf(__uint_as_float(v[4 * i + 1]) * inv_s1, 0.f),
                 hi[2 * i], lo[2 * i]);
          split2(fmaxf(__uint_as_float(v[4 * i + 2]) * inv_s1, 0.f), fmaxf(__uint_as_float(v[4 * i + 3]) * inv_s1, 0.f),
                 hi[2 * i + 1], lo[2 * i + 1]);
        }
        tmem_st_16x128b_x8(tm_dh + hf * 64, hi);            // in place: the 64 fp32 columns just read become 32 + 32
        tmem_st_16x128b_x8(tm_dh + hf * 64 + 32, lo);       // packed fp16 columns (this warp owns these 16 lanes)
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&bars->hop_full, 0);
#if GP_UMMA_PROF
      prof[2] += clock64() - t_e1;
      const long long t_e2 = clock64();
#endif

      // ---------------- Epi2: gate + scores (each thread: 2 rows x 32 units per tile) ----------------
      // packed fp32 (FFMA2 / FADD2 / FMUL2): each 64-bit value holds the two adjacent units a thread owns in a row
      uint64_t sa2[KB], sb2[KB];
#pragma unroll
      for (int k = 0; k < KB; ++k) sa2[k] = sb2[k] = 0ull;
      // one 32-unit quarter of the gate: V columns [0, 32), U columns [32, 64) of a D2 buffer; this thread owns
      // units unit0 + 8 ii + 2 cp + {0, 1} (ii < 4) of rows a and b
      auto gate_quarter = [&](const uint32_t (&zv)[16], const uint32_t (&zu)[16], int unit0) {
        const uint64_t cva2 = pack2(cva, cva), cua2 = pack2(cua, cua), one2 = pack2(1.f, 1.f), mone2 = pack2(-1.f, -1.f);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const float* rec = cstp + ((unit0 >> 1) + ii * 4 + cp) * (2 * CREC);
          uint64_t cr2[CREC];      // [0, KB): score weights, KB: bv', KB + 1: bu'
#pragma unroll
          for (int w = 0; w < CREC / 2; ++w) {
            const float4 v = *reinterpret_cast<const float4*>(rec + 4 * w);
            cr2[2 * w] = pack2(v.x, v.y);
            cr2[2 * w + 1] = pack2(v.z, v.w);
          }
          // tanh(a) sigmoid(b) = (1 - Ea) / ((1 + Ea)(1 + Eb)), Ea = e^-2a, Eb = e^-b.  Ea's exponent is clamped so
          // that (1 - Ea) stays finite; Eb may overflow to +inf: the quotient is then (finite) * 0 = 0, the limit.
          // D2 holds Sg z: the scales and biases are folded into the exponent-domain constants
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint64_t xa = fma2(pack2(__uint_as_float(zv[4 * ii + 2 * r]), __uint_as_float(zv[4 * ii + 2 * r + 1])), cva2, cr2[KB]);
            const uint64_t xb = fma2(pack2(__uint_as_float(zu[4 * ii + 2 * r]), __uint_as_float(zu[4 * ii + 2 * r + 1])), cua2, cr2[KB + 1]);
            const uint64_t ea = pack2(ex2_approx(fminf(lo2(xa), 57.7f)), ex2_approx(fminf(hi2(xa), 57.7f)));
            const uint64_t eb = pack2(ex2_approx(lo2(xb)), ex2_approx(hi2(xb)));
            const uint64_t den = mul2(add2(ea, one2), add2(eb, one2));
            const uint64_t g = mul2(fma2(ea, mone2, one2), pack2(rcp_approx(lo2(den)), rcp_approx(hi2(den))));
            if (r == 0) {
#pragma unroll
              for (int k = 0; k < KB; ++k) sa2[k] = fma2(g, cr2[k], sa2[k]);
            } else {
#pragma unroll
              for (int k = 0; k < KB; ++k) sb2[k] = fma2(g, cr2[k], sb2[k]);
            }
          }
        }
      };
#pragma unroll 1
      for (int qr = 0; qr < 4; ++qr) {
        const int b = qr & 1;
        { PROF_T0(); mbar_wait_cluster(&bars->d2_full[b], (uint32_t)(2 * t + (qr >> 1)) & 1u); PROF_ADD(1); }
        tc_fence_after();
        uint32_t zv[16], zu[16];
        tmem_ld_16x256b_x4(tm + lane_addr + TM_D2 + b * 64, zv);
        tmem_ld_16x256b_x4(tm + lane_addr + TM_D2 + b * 64 + 32, zu);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bars->d2_empty[b], 0);   // buffer b is in registers: the MMA warp may refill it
        gate_quarter(zv, zu, qr * 32);
      }
      // even + odd units, then the 4 threads of a row group hold disjoint unit subsets: finish the dot products
      float sa[KB], sb[KB];
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        sa[k] = lo2(sa2[k]) + hi2(sa2[k]);
        sb[k] = lo2(sb2[k]) + hi2(sb2[k]);
        sa[k] += __shfl_xor_sync(0xffffffffu, sa[k], 1);
        sb[k] += __shfl_xor_sync(0xffffffffu, sb[k], 1);
        sa[k] += __shfl_xor_sync(0xffffffffu, sa[k], 2);
        sb[k] += __shfl_xor_sync(0xffffffffu, sb[k], 2);
        sa[k] += bwk[k];
        sb[k] += bwk[k];
      }
#if GP_UMMA_PROF
      prof[3] += clock64() - t_e2;
      const long long t_e3 = clock64();
#endif

      // ---------------- raw scores out: thread (rg, cp) stores branches cp and cp + 4 of its two rows ----------------
      if (p.mp.a_out != nullptr) {
        float* ao = p.mp.a_out + seg.row_off[s_cur];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          if (k < K && (k & 3) == cp) {
            if (valid_a) ao[(size_t)k * p.mp.a_ld + row_a] = sa[k];
            if (valid_b) ao[(size_t)k * p.mp.a_ld + row_b] = sb[k];
          }
        }
      }

      // ---------------- candidates: rows that beat the CTA's n-th best score are parked ----------------
      float pa[KB], pb[KB];
      unsigned ex_a = 0u, ex_b = 0u;       // bit k: row a / b of this row group is parked for branch k
      int slot_a = 0, slot_b = 0;          // ... in this h row slot of the holder (one slot per parked row)
      const int rowcap = seg.row_cap;
      if (KB <= CAND_KMAX && nm > 0 && boot) {
        // First tile of the bag in this CTA: every row would be a candidate, so instead of 8 warps queueing at the
        // locks, warp k picks the tile's top n of branch k in one go (descending order: exactly n insertions).
        float* rsc = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.rec_score) + (size_t)cb * K * rcap;
        int* rix = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.rec_idx) + (size_t)cb * K * rcap;
        int* rsl = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.rec_slot) + (size_t)cb * K * rcap;
        float* g_score = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_score) + (size_t)cb * K * cap;
        float* sc_all = reinterpret_cast<float*>(smem + sm.ps);                 // [8 warps][16 rows][8]
        unsigned char* bflag = smem + sm.tbuf;                                   // [128 rows][8]: 0 or record + 1
        if (cp < 2) {
          float4 s0, s1;
          s0.x = cp ? sb[0] : sa[0];
          s0.y = KB > 1 ? (cp ? sb[KB > 1 ? 1 : 0] : sa[KB > 1 ? 1 : 0]) : 0.f;
          s0.z = KB > 2 ? (cp ? sb[KB > 2 ? 2 : 0] : sa[KB > 2 ? 2 : 0]) : 0.f;
          s0.w = KB > 3 ? (cp ? sb[KB > 3 ? 3 : 0] : sa[KB > 3 ? 3 : 0]) : 0.f;
          s1.x = KB > 4 ? (cp ? sb[KB > 4 ? 4 : 0] : sa[KB > 4 ? 4 : 0]) : 0.f;
          s1.y = KB > 5 ? (cp ? sb[KB > 5 ? 5 : 0] : sa[KB > 5 ? 5 : 0]) : 0.f;
          s1.z = s1.w = 0.f;
          const int prow = rg + 8 * cp;
          *reinterpret_cast<float4*>(psw + prow * 8) = s0;
          *reinterpret_cast<float4*>(psw + prow * 8 + 4) = s1;
          *reinterpret_cast<uint2*>(bflag + (lane_base + prow) * 8) = make_uint2(0u, 0u);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (e_idx < K) {
          const int k = e_idx;
          float sv[4];
          unsigned todo = 0u;      // bit i: tile row lane + 32 i is still in the running
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            sv[i] = sc_all[(((r >> 4) & 1) * 4 + (r >> 5)) * 128 + (r & 15) * 8 + k];
            if (tp.row_in_bag + r < n_rows) todo |= 1u << i;
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
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (e_idx == 0 && lane == 0) *reinterpret_cast<volatile int*>(&cs->epoch) = epoch_cur;     // lists are live
        {
          // lane (rg, cp = 0) speaks for row a, lane (rg, cp = 1) for row b: a selected row takes ONE h slot, which all
          // its records (one per branch that selected it) point to; n rows per branch <= K n <= row_cap: cannot overflow
          const uint2 f = *reinterpret_cast<const uint2*>(bflag + (lane_base + rg + 8 * (cp & 1)) * 8);
          unsigned mine = 0u;
          int myslot = 0;
          if (cp < 2 && (f.x | f.y) != 0u) {
            myslot = atomicAdd(&cs->rows, 1);
#pragma unroll
            for (int k = 0; k < KB; ++k) {
              const unsigned v = ((k < 4 ? f.x : f.y) >> (8 * (k & 3))) & 0xffu;
              if (v) { mine |= 1u << k; rsl[(size_t)k * rcap + (int)v - 1] = myslot; }
            }
          }
          ex_a = __shfl_sync(0xffffffffu, mine, lane & ~3);
          ex_b = __shfl_sync(0xffffffffu, mine, (lane & ~3) | 1);
          slot_a = __shfl_sync(0xffffffffu, myslot, lane & ~3);
          slot_b = __shfl_sync(0xffffffffu, myslot, (lane & ~3) | 1);
        }
      } else if (nm > 0) {
        float* rsc = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.rec_score) + (size_t)cb * K * rcap;
        int* rix = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.rec_idx) + (size_t)cb * K * rcap;
        int* rsl = reinterpret_cast<int*>(p.mp.ws + p.mp.wl.rec_slot) + (size_t)cb * K * rcap;
        unsigned* recs = reinterpret_cast<unsigned*>(smem + sm.tbuf + 1024);
        // lane (rg, cp = 0) speaks for row a, lane (rg, cp = 1) for row b of the row group
        const bool mine = (cp == 0 && valid_a) || (cp == 1 && valid_b);
        // all thresholds first (independent shared-memory reads), one vote for the common "nothing to park" outcome
        unsigned hits = 0u;
#pragma unroll
        for (int k = 0; k < (KB <= CAND_KMAX ? KB : 0); ++k) {
          if (k < K) {
            const unsigned long long gq = *reinterpret_cast<volatile unsigned long long*>(&cs->gtau[k]);
            const float tau = fmaxf(*reinterpret_cast<volatile float*>(&cs->tau[k]),
                                    (unsigned)(gq >> 32) == (unsigned)s_cur ? __uint_as_float((unsigned)gq) : -INFINITY);
            if (mine && (cp == 0 ? sa[k] : sb[k]) > tau) hits |= 1u << k;
          }
        }
        if (__any_sync(0xffffffffu, hits != 0u)) {
          int myslot = 0;
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
                const float sv = cp == 0 ? sa[k] : sb[k];
                const int rec = atomicAdd(&cs->app[k], 1);
                if (rec < rcap) {
                  rsc[(size_t)k * rcap + rec] = sv;
                  rix[(size_t)k * rcap + rec] = (int)(cp == 0 ? row_a : row_b);
                  rsl[(size_t)k * rcap + rec] = myslot;
                  *reinterpret_cast<volatile unsigned*>(&recs[k * rcap + rec]) = __float_as_uint(sv);
                } else {
                  reinterpret_cast<volatile int*>(p.mp.ws + p.mp.wl.flags)[s_cur] = 1;
                  hits &= ~(1u << k);
                }
              }
            }
          }
          ex_a = __shfl_sync(0xffffffffu, hits, lane & ~3);
          ex_b = __shfl_sync(0xffffffffu, hits, (lane & ~3) | 1);
          slot_a = __shfl_sync(0xffffffffu, myslot, lane & ~3);
          slot_b = __shfl_sync(0xffffffffu, myslot, (lane & ~3) | 1);
        }
      }
      boot = false;
      // running reference: a taken row more than REF_SLACK above m_ref moves it (first tile of a stream: from -inf)
      {
        float tmax[KB];
        bool grow = false;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const bool take_a = valid_a && !((ex_a >> k) & 1u), take_b = valid_b && !((ex_b >> k) & 1u);
          tmax[k] = fmaxf(take_a ? sa[k] : -INFINITY, take_b ? sb[k] : -INFINITY);
          grow |= tmax[k] > m_ref[k] + REF_SLACK;
        }
        if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
          for (int k = 0; k < KB; ++k) {
            float v = tmax[k];
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
            if (v > m_ref[k] + REF_SLACK) raise_ref(k, v);
          }
        }
      }
      // softmax numerators p' = 2^(s log2e + c_ref)  (<= 2^15 by construction; parked / out-of-range rows: 0)
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        const bool take_a = valid_a && !((ex_a >> k) & 1u), take_b = valid_b && !((ex_b >> k) & 1u);
        pa[k] = take_a ? ex2_approx(fmaf(sa[k], LOG2E, c_ref[k])) : 0.f;
        pb[k] = take_b ? ex2_approx(fmaf(sb[k], LOG2E, c_ref[k])) : 0.f;
        if (cp == 0) l_run[k] += pa[k] + pb[k];       // each row once; lanes are summed at flush
      }
      if (cp < 2) {   // p' of row a (cp 0) / row b (cp 1) -> smem, from where the pool builds its B fragments
        float4 p0, p1;
        p0.x = cp ? pb[0] : pa[0];
        p0.y = KB > 1 ? (cp ? pb[KB > 1 ? 1 : 0] : pa[KB > 1 ? 1 : 0]) : 0.f;
        p0.z = KB > 2 ? (cp ? pb[KB > 2 ? 2 : 0] : pa[KB > 2 ? 2 : 0]) : 0.f;
        p0.w = KB > 3 ? (cp ? pb[KB > 3 ? 3 : 0] : pa[KB > 3 ? 3 : 0]) : 0.f;
        p1.x = KB > 4 ? (cp ? pb[KB > 4 ? 4 : 0] : pa[KB > 4 ? 4 : 0]) : 0.f;
        p1.y = KB > 5 ? (cp ? pb[KB > 5 ? 5 : 0] : pa[KB > 5 ? 5 : 0]) : 0.f;
        p1.z = KB > 6 ? (cp ? pb[KB > 6 ? 6 : 0] : pa[KB > 6 ? 6 : 0]) : 0.f;
        p1.w = KB > 7 ? (cp ? pb[KB > 7 ? 7 : 0] : pa[KB > 7 ? 7 : 0]) : 0.f;
        const int prow = rg + 8 * cp;
        *reinterpret_cast<float4*>(psw + prow * 8) = p0;
        *reinterpret_cast<float4*>(psw + prow * 8 + 4) = p1;
      }
#if GP_UMMA_PROF
      prof[4] += clock64() - t_e3;
      const long long t_e4 = clock64();
#endif
      // ---------------- pool: acc[feature][branch] += sum_rows h[row][feature] p'[row][branch] ----------------
      // A contraction over the warp's 16 rows = the K dimension of mma.sync m16n8k16 with A = h^T (16 features x 16 rows)
      // and B = p' (16 rows x 8 branches).  tcgen05.ld 16x128b hands out the packed fp16 h operand in exactly the
      // fragment layout movmatrix expects (thread (rg, cp): rows rg / rg + 8, feature pair 8i + 2cp), so a transpose is
      // one MOVM per 8x8 block and there is no shared-memory round trip.  fp32-faithful through the same hi/lo split as
      // the big GEMMs: h_hi p_hi + h_lo p_hi + h_hi p_lo.
      float* cand_h = reinterpret_cast<float*>(p.mp.ws + p.mp.wl.cand_h) + (size_t)cb * rowcap * L;
      float* t0 = ex_a ? cand_h + (size_t)slot_a * L : nullptr;      // parked rows of this row group -> their h slots
      float* t1 = ex_b ? cand_h + (size_t)slot_b * L : nullptr;
      __syncwarp();
      uint32_t bh[2], bl[2];      // B fragments: {p'[2cp][rg], p'[2cp+1][rg]}, {p'[2cp+8][rg], p'[2cp+9][rg]} as fp16 hi / lo
      {
        const float q0 = psw[(2 * cp) * 8 + rg], q1 = psw[(2 * cp + 1) * 8 + rg];
        const float q2 = psw[(2 * cp + 8) * 8 + rg], q3 = psw[(2 * cp + 9) * 8 + rg];
        split2(q0, q1, bh[0], bl[0]);
        split2(q2, q3, bh[1], bl[1]);
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t hh[16], hl[16];     // regs {2i, 2i+1}: rows rg / rg + 8, features 64 hf + 8i + 2cp + {0,1}
        tmem_ld_16x128b_x8(tm_dh + hf * 64, hh);
        tmem_ld_16x128b_x8(tm_dh + hf * 64 + 32, hl);
        tmem_wait_ld();
        if (ex_a | ex_b) {   // park the rows that entered a list this tile (fp32 h = hi + lo)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 ah = __half22float2(*reinterpret_cast<const __half2*>(&hh[2 * i]));
            const float2 al = __half22float2(*reinterpret_cast<const __half2*>(&hl[2 * i]));
            const float2 bhf = __half22float2(*reinterpret_cast<const __half2*>(&hh[2 * i + 1]));
            const float2 blf = __half22float2(*reinterpret_cast<const __half2*>(&hl[2 * i + 1]));
            const float2 fa = make_float2(ah.x + al.x, ah.y + al.y), fb = make_float2(bhf.x + blf.x, bhf.y + blf.y);
            const int fcol = hf * 64 + i * 8 + cp * 2;
            if (t0 != nullptr) *reinterpret_cast<float2*>(t0 + fcol) = fa;
            if (t1 != nullptr) *reinterpret_cast<float2*>(t1 + fcol) = fb;
          }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          // blocks 2jj, 2jj + 1 of this half -> A fragment {blk0 rows 0-7, blk1 rows 0-7, blk0 rows 8-15, blk1 rows 8-15}
          uint32_t ahi[4], alo[4];
          ahi[0] = movmatrix_t(hh[4 * jj]);
          ahi[1] = movmatrix_t(hh[4 * jj + 2]);
          ahi[2] = movmatrix_t(hh[4 * jj + 1]);
          ahi[3] = movmatrix_t(hh[4 * jj + 3]);
          alo[0] = movmatrix_t(hl[4 * jj]);
          alo[1] = movmatrix_t(hl[4 * jj + 2]);
          alo[2] = movmatrix_t(hl[4 * jj + 1]);
          alo[3] = movmatrix_t(hl[4 * jj + 3]);
          mma_16816_f16(acc[hf * 4 + jj], ahi, bh);
          mma_16816_f16(acc[hf * 4 + jj], alo, bh);
          mma_16816_f16(acc[hf * 4 + jj], ahi, bl);
        }
      }
      tc_fence_before();
      __syncwarp();      // psw is rewritten by the next tile; the h operand has been read: its buffer may take tile t + 2
      if (lane == 0) mbar_arrive_cluster(&bars->d1_empty[t & 1], 0);
#if GP_UMMA_PROF
      prof[5] += clock64() - t_e4;
#endif
    }
    { PROF_T0(); flush_stream(); PROF_ADD(6); }
    if (KB <= CAND_KMAX && e_idx == 0 && lane == 0) *reinterpret_cast<volatile int*>(&cs->cur_bag) = -2;     // stops the service warp
#if GP_UMMA_PROF
    prof[7] = clock64() - t_start;
    if (lane == 0) PROF_FLUSH(24 + 8 * e_idx);
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
//   Wg part (hi, then lo): Wg = Wv for c == 0, Wu for c == 1: for quarter in 0..3, kb in {0,1}:
//                          tile [32 rows = units 32q..32q+31][64 k = features 64kb..] (4 KB)
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
      const uint32_t tile = (uint32_t)((u >> 5) * 2 + (k >> 6)) * 4096u;          // [quarter][k block]: 32 rows x 128 B
      const uint32_t inner = sw128_offset((uint32_t)(u & 31), (uint32_t)((k & 63) >> 3)) + (uint32_t)(k & 7) * 2u;
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

// the device-resident constants of the row pass, from the fp32 section of the packed blob (already laid out and
// zero-filled where a bias is absent) and the operand scales
__global__ void umma_consts_kernel(const float* __restrict__ f32, GpPackLayout lay, const float* __restrict__ absmax,
                                   UmmaConsts* __restrict__ out) {
  const int u = threadIdx.x;      // 128 threads
  out->bv[u] = f32[lay.bv + u];
  out->bu[u] = f32[lay.bu + u];
  for (int k = 0; k < KMAX; ++k) out->ww[k][u] = f32[lay.ww + (size_t)k * 128 + u];
  if (u < KMAX) out->bw[u] = f32[lay.bw + u];
  if (u == 0) {
    out->inv_s1 = 1.f / pow2_scale(absmax[0]);
    out->inv_sv = 1.f / pow2_scale(absmax[1]);
    out->inv_su = 1.f / pow2_scale(absmax[2]);
    out->pad = 0.f;
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
  // (no front-layer bias: Epi1 is relu(D1 / S1) only -- CLAM's Linear+bias front layer runs on the FFMA kernel)
  return s.front == 1 && s.front_bias == 0 && s.front_act == ACMIL_ACT_RELU && s.act_a == ACMIL_ACT_TANH && s.gated == 1 && s.d_inner == 128 &&
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
  UmmaConsts* dc = reinterpret_cast<UmmaConsts*>(img + 2 * (size_t)cta_img);
  umma_consts_kernel<<<1, 128, 0, st>>>(d_f32, lay, d_absmax, dc);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  if (consts) {
    // optional host copy for callers that want to look at the vectors: the only part of packing that synchronises.
    // The kernels never read it (non-finite score weights give non-finite scores, as in the reference).
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
  if (cap > 0 && ncl * 2 * cap > GP_MAX_SEG_CAND) ncl = GP_MAX_SEG_CAND / (2 * cap);   // reduce kernel's smem bound
  if (ncl < 1) ncl = 1;
  t->u_nclusters = ncl;
  t->n_masked_cap = cap;
  t->cand_div = 1;                      // one segment (and candidate holder) per CTA and bag
  t->rec_cap = cap > 0 ? REC_CAP : 0;
  t->row_cap = cap > 0 ? ROW_CAP : 0;   // one h slot per parked row, shared by the branches
  t->h_branch_stride = 0;
  // segments: 2 per (cluster, bag) pair that intersects (one per CTA)
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
    seg += 2 * count;
  }
  t->seg_begin[b.n_slides] = seg;
  t->n_seg = seg;
  return 0;
}

int gp_launch_main_umma(const GpMainParams& p, const acmil_gp_consts* consts, const unsigned char* d_umma, cudaStream_t st) {
  const acmil_gp_shape& s = p.sh;
  ACMIL_REQUIRE(gp_umma_supported(s), ACMIL_E_UNSUPPORTED, "tcgen05 kernel does not support this shape");
  (void)consts;      // the constants live in the packed blob (device); the host copy is informational
  if (p.seg.u_total_pt == 0) return ACMIL_OK;
  static thread_local UmmaParams up;   // large: keep off the stack
  up.mp = p;
  up.dc = reinterpret_cast<const UmmaConsts*>(d_umma + 1024 + 2 * (size_t)umma_cta_img_bytes(s));
  up.wimg = d_umma + 1024;
  up.cta_img_bytes = umma_cta_img_bytes(s);
  up.w1_part_bytes = (uint32_t)(s.d_in / 64) * 8192u;
  EncodeFn enc = get_encode();
  ACMIL_REQUIRE(enc != nullptr, ACMIL_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const int64_t rows = p.seg.row_off[p.seg.n_slides];
  cuuint64_t dims[2] = {(cuuint64_t)s.d_in, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)s.d_in * (p.x_f16 ? 2 : 4)};
  cuuint32_t box[2] = {p.x_f16 ? 64u : (cuuint32_t)KC, 128};      // 128 bytes per row either way
  cuuint32_t es[2] = {1, 1};
  ACMIL_REQUIRE(((uintptr_t)p.x & 15) == 0, ACMIL_E_INVALID, "x must be 16-byte aligned for TMA");
  const CUresult r = enc(&up.tmap, p.x_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         const_cast<float*>(p.x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ACMIL_REQUIRE(r == CUDA_SUCCESS, ACMIL_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  const int K = s.n_branch;
  ACMIL_REQUIRE(K <= 6 || p.seg.n_masked_cap == 0, ACMIL_E_UNSUPPORTED,
                "tcgen05 kernel: masking with more than 6 branches is not supported (use ACMIL_IMPL_FFMA)");
  const size_t smem = smem_map(s.d_in, K == 1 ? 1 : (K <= 5 ? 5 : 8)).total + 1024;
  const int grid = p.seg.u_nclusters * 2;
  if (p.seg.n_masked_cap > 0)     // one memset: per-bag overflow flags = -1 (clean), mirrored top-n lists = NaN
    ACMIL_CHECK_CUDA(cudaMemsetAsync(p.ws + p.wl.flags, 0xFF, p.wl.cand_idx - p.wl.flags, st));   // ("not written in this launch")
  if (K == 1) return launch_kb<1>(up, grid, smem, st);
  if (K <= 5) return launch_kb<5>(up, grid, smem, st);
  return launch_kb<8>(up, grid, smem, st);
}
