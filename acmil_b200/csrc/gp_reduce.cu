// Second half of the gated-attention pool: merge of per-CTA partials, top-n / mask selection,
// normalisation, classifier heads; plus the small [K, N] ops (attention stats, row softmax).
//
//   gp_reduce_kernel : per (bag, branch, 32-feature chunk): picks this rank's top-n candidates,
//                      adds every other candidate back into the sums (exact: nothing is ever
//                      subtracted), LSE-merges the segment partials -> one rank record per bag.
//   gp_finish_kernel : per bag: global top-n over the ranks' lists (= torch.topk order,
//                      transformer.py:315), masked = top[rsel] (:316-317), masked scores count as
//                      -1e9 in the softmax exactly like masked_fill (:320), afeat = softmax(A) h
//                      (:323-324), classifier heads (:325-330), -1e9 written into a_out.
#include "gp_common.cuh"

namespace {

constexpr int RT = 256;
#ifndef GP_EXP_NO_SYSFENCE
#define GP_EXP_NO_SYSFENCE 0      // timing experiment: device-scope fences in the exchange (NOT correct across GPUs)
#endif
#if GP_EXP_NO_SYSFENCE
#define GP_SYS_FENCE() __threadfence()
#else
#define GP_SYS_FENCE() __threadfence_system()
#endif
constexpr float MASK_FILL = -1e9f;

struct GpReduceParams {
  acmil_gp_shape sh;
  const unsigned char* ws;
  GpWorkspace wl;
  GpSegTable seg;
  GpRecord rec;
  float* record;  // [S][stride]
  const int* flags;   // per-bag overflow flags of the tcgen05 kernel (or NULL)
  int flag_mode;      // GP_REDUCE_*
  GpExchange x;       // n_ranks > 0: the record goes to every rank's gather buffer (NVLink stores) + flag handshake
};

// system-scope accesses for the flag handshake between GPUs
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// End of a reduce CTA when the exchange is fused in: every CTA of the step takes a ticket (also the ones that had
// nothing to do); the last one knows that all records of this rank are written (each writer fenced at system scope
// before its ticket) and raises this rank's flag on every peer.
__device__ void exchange_publish(const GpExchange& x, uint32_t epoch, bool wrote) {
  __syncthreads();      // the CTA's stores happen-before thread 0's fence (one system-scope fence per CTA, not per thread)
  if (threadIdx.x == 0) {
    if (wrote) GP_SYS_FENCE();
    const unsigned t = atomicAdd(x.ticket, 1u);
    if (t == (unsigned)x.reduce_ctas - 1u) {
      *x.ticket = 0u;
      GP_SYS_FENCE();      // ONE system-scope fence orders every record store of this rank before the flags
      for (int r = 0; r < x.n_ranks; ++r) st_relaxed_sys(x.flags[r] + x.rank, epoch + 1u);
    }
  }
}

// block-wide argmax of (score desc, idx asc); returns winner position (or -1) to every thread
__device__ int block_argbest(float s, int i, int pos, float* r_s, int* r_i, int* r_p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
    const int p2 = __shfl_xor_sync(0xffffffffu, pos, o);
    if (p2 >= 0 && (pos < 0 || cand_better(s2, i2, s, i))) { s = s2; i = i2; pos = p2; }
  }
  __syncthreads();
  if (lane == 0) { r_s[warp] = s; r_i[warp] = i; r_p[warp] = pos; }
  __syncthreads();
  if (warp == 0) {      // the warps' winners are reduced by one warp (not scanned by every thread) and handed out through r_p[0]
    float bs = lane < nw ? r_s[lane] : -INFINITY;
    int bi = lane < nw ? r_i[lane] : 0x7fffffff, bp = lane < nw ? r_p[lane] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      const int p2 = __shfl_xor_sync(0xffffffffu, bp, o);
      if (p2 >= 0 && (bp < 0 || cand_better(s2, i2, bs, bi))) { bs = s2; bi = i2; bp = p2; }
    }
    __syncwarp();
    if (lane == 0) r_p[0] = bp;
  }
  __syncthreads();
  return r_p[0];
}

__device__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < nw; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < nw; ++w) r += red[w];  // fixed order: deterministic
  return r;
}

// One CTA of 1024 threads per (bag, branch).  Latency is what matters here (a few hundred KB at most):
// warp 0 selects the rank's top-n candidates while warps 1..31 merge the segment partials, then all 32 warps
// add the unselected candidates back; each warp covers the 128 features with one float4 per lane.
// RR = threads per CTA: 1024 for the long segment / candidate lists of a whole 50k-row bag on one GPU, 256 when a bag
// has few segments (row-sharded bags, small bags), where 8 CTAs per SM finish the whole batch in one wave.
// NJ = float4 groups of a feature row per lane (1 for L <= 128): sizes the register tiles.  The 256-thread variant is
// compiled for 5 CTAs per SM so that the 640 CTAs of an 8-rank step (128 bag shards x 5 branches) fit ONE wave.
template <int RR, int NJ>
__global__ void __launch_bounds__(RR, RR == 256 ? (NJ == 1 ? 5 : 3) : 1) gp_reduce_kernel(const __grid_constant__ GpReduceParams p) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ float red[RR / 32];
  __shared__ int sel_pos[NMAX];
  __shared__ int s_nsel, s_nlive;
  __shared__ float r_s[RR / 32];
  __shared__ int r_i[RR / 32], r_p[RR / 32];

  const int L = p.sh.d_inner, K = p.sh.n_branch;
  const int k = blockIdx.x % K, s = blockIdx.x / K;
  const uint32_t epoch = p.x.n_ranks > 0 ? *p.x.epoch : 0u;
  if (p.flag_mode != GP_REDUCE_ALL) {
    const bool flagged = p.flags[s] == 1;      // the tcgen05 kernel ran out of parking slots on this bag
    if (flagged != (p.flag_mode == GP_REDUCE_FLAGGED)) {
      if (p.x.n_ranks > 0) exchange_publish(p.x, epoch, false);
      return;
    }
  }
  const int seg0 = p.seg.seg_begin[s], nseg = p.seg.seg_begin[s + 1] - seg0;
  const int nm = p.seg.nm[s], cap = p.seg.n_masked_cap, cdiv = p.seg.cand_div;
  const size_t rowcap = (size_t)p.seg.row_cap;
  const int hoff = k * p.seg.h_branch_stride;      // first h slot of this branch inside a holder
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cb0 = seg0 / cdiv, ncb = nseg / cdiv;      // candidate holders of this bag
  const int ncap = ncb * nm;                           // candidate slots of this (bag, branch); the live ones are compacted below
  const int nq = L / 4;                                // float4 groups per row (<= 128 for L <= 512)

  float* c_score = reinterpret_cast<float*>(dsm);
  int* c_idx = reinterpret_cast<int*>(c_score + ncap);
  int* c_slot = c_idx + ncap;
  int* c_sel = c_slot + ncap;                          // holder index << 1 | selected
  float4* wacc = reinterpret_cast<float4*>(dsm + (((size_t)ncap * 16 + 15) / 16) * 16);   // [32 warps][nq]

  const float* part = reinterpret_cast<const float*>(p.ws + p.wl.part);
  const int* g_cnt = reinterpret_cast<const int*>(p.ws + p.wl.cand_cnt);
  const float* g_score = reinterpret_cast<const float*>(p.ws + p.wl.cand_score);
  const int* g_idx = reinterpret_cast<const int*>(p.ws + p.wl.cand_idx);
  const int* g_slot = reinterpret_cast<const int*>(p.ws + p.wl.cand_slot);
  const float* g_h = reinterpret_cast<const float*>(p.ws + p.wl.cand_h);

  // ---- 1. live candidates of this (bag, branch) -> smem, compacted (most holder slots are empty: the row pass keeps
  // only what can still be in the bag's top n); reference point m* = max over everything ----
  if (tid == 0) s_nlive = 0;
  __syncthreads();
  float mx = -INFINITY;
  for (int c = tid; c < ncap; c += RR) {
    const int sg = c / nm, i = c % nm;
    const size_t g = ((size_t)(cb0 + sg) * K + k) * cap + i;
    if (i < g_cnt[(size_t)(cb0 + sg) * K + k]) {
      const float sc = g_score[g];
      const int pos = atomicAdd(&s_nlive, 1);
      c_score[pos] = sc;
      c_idx[pos] = g_idx[g];
      c_slot[pos] = g_slot[g];
      c_sel[pos] = sg << 1;
      mx = fmaxf(mx, sc);
    }
  }
  for (int sg = tid; sg < nseg; sg += RR) mx = fmaxf(mx, part[((size_t)(seg0 + sg) * K + k) * (L + 2)]);
  const float mstar = block_max(mx, red);      // (contains a __syncthreads: the smem lists are visible)
  const int ncand = s_nlive;

  float4 a4[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) a4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float ls = 0.f;
  // ---- 2. segment partials + this rank's top-n (sorted: score desc, index asc) ----
  // Few candidates (batches of bags, row-sharded bags): warp 0 selects alone while warps 1.. merge the partials.  Many (a
  // whole 50k-row bag on one GPU brings ~1500 per branch: the lone warp was 60 % of the kernel): all warps merge, then
  // every thread scans its own candidates and the CTA agrees on the best one per round.
  const bool wide_select = ncand > 384;
  auto merge_partials = [&](int first, int step) {
    for (int sg = first; sg < nseg; sg += step) {
      const float* pr = part + ((size_t)(seg0 + sg) * K + k) * (L + 2);
      const float m = pr[0];
      const float w = m == -INFINITY ? 0.f : expf(m - mstar);
      if (lane == 0) ls += w * pr[1];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int qd = lane + 32 * j;
        if (qd < nq) {
          const float2 u0 = *reinterpret_cast<const float2*>(pr + 2 + qd * 4);      // rows are 8-byte aligned (L + 2 floats)
          const float2 u1 = *reinterpret_cast<const float2*>(pr + 2 + qd * 4 + 2);
          a4[j].x = fmaf(w, u0.x, a4[j].x); a4[j].y = fmaf(w, u0.y, a4[j].y);
          a4[j].z = fmaf(w, u1.x, a4[j].z); a4[j].w = fmaf(w, u1.y, a4[j].w);
        }
      }
    }
  };
  int nsel = 0;
  if (wide_select) {
    merge_partials(warp, RR / 32);
    int live = 0;
    for (int c = tid; c < ncand; c += RR) live += 1;
    nsel = min(nm, (int)(block_sum((float)live, red) + 0.5f));
    for (int round = 0; round < nsel; ++round) {
      float bs = -INFINITY;
      int bi = 0x7fffffff, bp = -1;
      for (int c = tid; c < ncand; c += RR) {
        if (!(c_sel[c] & 1) && (bp < 0 || cand_better(c_score[c], c_idx[c], bs, bi))) {
          bs = c_score[c]; bi = c_idx[c]; bp = c;
        }
      }
      bp = block_argbest(bs, bi, bp, r_s, r_i, r_p);
      if (tid == 0) { sel_pos[round] = bp; c_sel[bp] |= 1; }
      __syncthreads();
    }
  } else {
    if (warp == 0) {
      int live = 0;
      for (int c = lane; c < ncand; c += 32) live += 1;
      live = (int)(warp_sum((float)live) + 0.5f);
      nsel = min(nm, live);
      for (int round = 0; round < nsel; ++round) {
        float bs = -INFINITY;
        int bi = 0x7fffffff, bp = -1;
        for (int c = lane; c < ncand; c += 32) {
          if (!(c_sel[c] & 1) && (bp < 0 || cand_better(c_score[c], c_idx[c], bs, bi))) {
            bs = c_score[c]; bi = c_idx[c]; bp = c;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
          const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
          const int p2 = __shfl_xor_sync(0xffffffffu, bp, o);
          if (p2 >= 0 && (bp < 0 || cand_better(s2, i2, bs, bi))) { bs = s2; bi = i2; bp = p2; }
        }
        if (lane == 0) { sel_pos[round] = bp; c_sel[bp] |= 1; }
        __syncwarp();
      }
    } else {
      merge_partials(warp - 1, RR / 32 - 1);
    }
    if (tid == 0) s_nsel = nsel;      // known to warp 0 only
    __syncthreads();
    nsel = s_nsel;
  }

  // ---- 3. candidates that were not selected rejoin the sums ----
  // (CPR candidates per round with all their loads in flight together: the rows sit in L2 at best)
  constexpr int CPR = NJ == 1 ? 4 : 2;
  for (int c0 = warp; c0 < ncand; c0 += CPR * (RR / 32)) {
    float wv[CPR];
    float4 uv[CPR][NJ];
#pragma unroll
    for (int t = 0; t < CPR; ++t) {
      const int c = c0 + t * (RR / 32);
      const bool on = c < ncand && !(c_sel[c] & 1);
      wv[t] = on ? expf(c_score[c] - mstar) : 0.f;
      const float* hr = g_h + ((size_t)(cb0 + (on ? (c_sel[c] >> 1) : 0)) * rowcap + hoff + (on ? c_slot[c] : 0)) * L;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int qd = lane + 32 * j;
        uv[t][j] = (on && qd < nq) ? __ldcg(reinterpret_cast<const float4*>(hr + qd * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int t = 0; t < CPR; ++t) {
      if (lane == 0) ls += wv[t];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        a4[j].x = fmaf(wv[t], uv[t][j].x, a4[j].x); a4[j].y = fmaf(wv[t], uv[t][j].y, a4[j].y);
        a4[j].z = fmaf(wv[t], uv[t][j].z, a4[j].z); a4[j].w = fmaf(wv[t], uv[t][j].w, a4[j].w);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int qd = lane + 32 * j;
    if (qd < nq) wacc[warp * nq + qd] = a4[j];
  }
  const float lstar = block_sum(ls, red);

  // with the exchange: this rank's slot of its own gather buffer, [parity][rank][bag]
  float* rec = p.x.n_ranks > 0
                   ? p.x.gather[p.x.rank] + (((size_t)(epoch & 1u) * p.x.n_ranks + p.x.rank) * p.seg.n_slides + s) * p.rec.stride()
                   : p.record + (size_t)s * p.rec.stride();
  for (int qd = tid; qd < nq; qd += RR) {      // fixed-order sum over the 32 warps: deterministic
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < RR / 32; ++w) {
      const float4 u = wacc[w * nq + qd];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    float* dst = rec + p.rec.acc() + (size_t)k * L + qd * 4;
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
  }
  // ---- 4. rank list: scores, global indices and h rows of the selected candidates ----
  const int nmc = p.rec.nmc;
  for (int i = warp; i < nmc; i += RR / 32) {
    const bool on = i < nsel;
    const int c = on ? sel_pos[i] : 0;
    const float* hr = on ? g_h + ((size_t)(cb0 + (c_sel[c] >> 1)) * rowcap + hoff + c_slot[c]) * L : nullptr;
    for (int jf = lane; jf < L; jf += 32) rec[p.rec.h() + ((size_t)k * nmc + i) * L + jf] = on ? hr[jf] : 0.f;
    if (lane == 0) {
      rec[p.rec.score() + (size_t)k * nmc + i] = on ? c_score[c] : -INFINITY;
      reinterpret_cast<int*>(rec)[p.rec.idx() + (size_t)k * nmc + i] = on ? c_idx[c] + (int)p.seg.shard_begin[s] : 0x7fffffff;
    }
  }
  if (tid == 0) {
    // an overflow flag that nobody rescues (GP_REDUCE_ALL with flags given) still fails loudly
    const bool poisoned = p.flag_mode == GP_REDUCE_ALL && p.flags != nullptr && p.flags[s] == 1;
    rec[p.rec.m() + k] = poisoned ? NAN : mstar;
    rec[p.rec.l() + k] = lstar;
    reinterpret_cast<int*>(rec)[p.rec.cnt() + k] = nsel;
  }
  if (p.x.n_ranks > 0) {
    // branch k's pieces of the record -> the same slot of every peer's gather buffer, straight over NVLink
    __syncthreads();      // (block scope: the pieces were written by this CTA)
    const size_t off = rec - p.x.gather[p.x.rank];
    const size_t piece[7][2] = {{p.rec.m() + k, 1}, {p.rec.l() + k, 1}, {p.rec.acc() + (size_t)k * L, (size_t)L},
                                {p.rec.cnt() + k, 1}, {p.rec.score() + (size_t)k * nmc, (size_t)nmc},
                                {p.rec.idx() + (size_t)k * nmc, (size_t)nmc}, {p.rec.h() + (size_t)k * nmc * L, (size_t)nmc * L}};
    // every thread reads its share of the pieces once (independent loads), then stores it to all peers
    constexpr int VMAX = 12;
    for (size_t base = 0;; base += (size_t)VMAX * RR) {      // (one round unless n_masked * d_inner is very large)
      float v[VMAX];
      size_t at[VMAX];
      int n = 0;
      size_t seen = 0;
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        for (size_t i = tid; i < piece[q][1]; i += RR) {
          const size_t ord = seen + i / RR;      // position of this element in the thread's own sequence
          if (ord >= base / RR && n < VMAX) { at[n] = piece[q][0] + i; ++n; }
        }
        seen += (piece[q][1] + RR - 1) / RR;
      }
      for (int j = 0; j < n; ++j) v[j] = __ldcg(rec + at[j]);
      for (int r = 0; r < p.x.n_ranks; ++r) {
        if (r == p.x.rank) continue;
        float* dst = p.x.gather[r] + off;
        for (int j = 0; j < n; ++j) dst[at[j]] = v[j];
      }
      if (seen * RR <= base + (size_t)VMAX * RR) break;
    }
    exchange_publish(p.x, epoch, true);
  }
}

// ------------------------------------------------------------------------------------------
// One CTA per (bag, branch): warp 0 forms the global top-n / masked set and the softmax normalisers of its branch (short,
// serial), then all 8 warps stream the rank partials and the candidate rows (the bulk of the bytes: n_ranks * n_masked
// rows of d_inner floats per branch) in a fixed assignment, so the result does not depend on timing.  The bag feature and
// the slide head need all branches: gp_heads_kernel, launched right behind.
__global__ void __launch_bounds__(RT, 5) gp_finish_kernel(const __grid_constant__ GpFinishParams p) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ float s_m, s_l;

  const int L = p.sh.d_inner, K = p.sh.n_branch, P = p.n_ranks, nmc = p.rec.nmc;
  const int s = blockIdx.x / K, k = blockIdx.x % K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t stride = p.rec.stride();
  const int ne = P * nmc;  // candidate entries of this branch (dense, holes = -inf)

  float4* wpart = reinterpret_cast<float4*>(dsm);                         // [8 warps][L / 4]
  float* af = reinterpret_cast<float*>(wpart + (size_t)(RT / 32) * (L / 4));   // [L]
  float* e_score = af + L;                                                 // [ne]
  int* e_idx = reinterpret_cast<int*>(e_score + ne);
  int* e_flag = e_idx + ne;                 // 0 dead, 1 live, 2 top (unmasked), 3 masked
  int* top_pos = e_flag + ne;               // [NMAX]

  const float* records = p.records;
  if (p.x.n_ranks > 0) {
    // records pushed by the ranks' reduce kernels: wait until every source has published this step
    const uint32_t epoch = *p.x.epoch;
    if (tid < P) {
      const uint32_t* f = p.x.flags[p.x.rank] + tid;
      while (ld_relaxed_sys(f) < epoch + 1u) __nanosleep(64);
      GP_SYS_FENCE();      // acquire side: the records behind the flag are visible to the loads below
    }
    __syncthreads();
    records = p.x.gather[p.x.rank] + (size_t)(epoch & 1u) * P * p.n_slides * stride;
  }
  auto recp = [&](int r) { return records + ((size_t)r * p.n_slides + s) * stride; };

  const int keep = p.keep[s];
  // ---- 1. global top-n and the masked subset; m*, l* (warp 0; the entries are loaded by everybody) ----
  for (int e = tid; e < ne; e += RT) {
    const int r = e / nmc, i = e % nmc;
    const float* rc = recp(r);
    const int cnt = reinterpret_cast<const int*>(rc)[p.rec.cnt() + k];
    const bool ok = i < cnt;
    e_score[e] = ok ? rc[p.rec.score() + (size_t)k * nmc + i] : -INFINITY;
    e_idx[e] = ok ? reinterpret_cast<const int*>(rc)[p.rec.idx() + (size_t)k * nmc + i] : 0x7fffffff;
    e_flag[e] = ok ? 1 : 0;
  }
  __syncthreads();
  if (warp == 0) {
    int live = 0;
    for (int e = lane; e < ne; e += 32) live += e_flag[e];
    live = (int)(warp_sum((float)live) + 0.5f);
    const int ntop = min(p.n_masked, live);
    for (int round = 0; round < ntop; ++round) {
      float bs = -INFINITY;
      int bi = 0x7fffffff, bp = -1;
      for (int e = lane; e < ne; e += 32) {
        if (e_flag[e] == 1 && (bp < 0 || cand_better(e_score[e], e_idx[e], bs, bi))) {
          bs = e_score[e]; bi = e_idx[e]; bp = e;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
        const int p2 = __shfl_xor_sync(0xffffffffu, bp, o);
        if (p2 >= 0 && (bp < 0 || cand_better(s2, i2, bs, bi))) { bs = s2; bi = i2; bp = p2; }
      }
      if (lane == 0) {
        top_pos[round] = bp;
        e_flag[bp] = 2;
        if (p.out.d_topk_idx) p.out.d_topk_idx[((size_t)s * K + k) * p.n_masked + round] = bi;
      }
      __syncwarp();
    }
    if (p.out.d_topk_idx)
      for (int i = ntop + lane; i < p.n_masked; i += 32) p.out.d_topk_idx[((size_t)s * K + k) * p.n_masked + i] = -1;
    __syncwarp();
    auto mask_one = [&](int i, int64_t sel) {      // the i-th masked entry is the top row of rank `sel`
      if (sel >= 0 && sel < ntop) {
        const int pos = top_pos[(int)sel];
        e_flag[pos] = 3;
        const int gi = e_idx[pos];
        if (p.out.d_masked_idx) p.out.d_masked_idx[((size_t)s * K + k) * p.keep_ld + i] = gi;
        const int64_t loc = (int64_t)gi - p.shard_begin[s];
        if (p.a_out && loc >= 0 && loc < p.row_off[s + 1] - p.row_off[s])
          p.a_out[(size_t)k * p.a_ld + p.row_off[s] + loc] = MASK_FILL;
      }
    };
    if (p.rand) {
      // rsel = argsort(rand[k, :ntop])[:keep] (transformer.py:316) inside the warp: lane j holds draw j, its rank is the
      // number of smaller draws (ties: lower index first); the lane of rank i provides the i-th masked entry
      const float rv = lane < ntop ? p.rand[((size_t)s * K + k) * p.rand_ld + lane] : INFINITY;
      int rank = 0;
      for (int l = 0; l < ntop; ++l) {
        const float o = __shfl_sync(0xffffffffu, rv, l);
        rank += (o < rv || (o == rv && l < lane)) ? 1 : 0;
      }
      if (lane < ntop && rank < keep) mask_one(rank, lane);
    } else {
      for (int i = lane; i < keep; i += 32) mask_one(i, p.rsel[((size_t)s * K + k) * p.keep_ld + i]);
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int r = lane; r < P; r += 32) mx = fmaxf(mx, recp(r)[p.rec.m() + k]);
    for (int e = lane; e < ne; e += 32) {
      const int f = e_flag[e];
      if (f == 1 || f == 2) mx = fmaxf(mx, e_score[e]);
      if (f == 3) mx = fmaxf(mx, MASK_FILL);
    }
    const float mstar = warp_max(mx);
    float ls = 0.f;
    for (int r = lane; r < P; r += 32) {
      const float* rc = recp(r);
      if (rc[p.rec.m() + k] != -INFINITY) ls += expf(rc[p.rec.m() + k] - mstar) * rc[p.rec.l() + k];
    }
    for (int e = lane; e < ne; e += 32) {
      const int f = e_flag[e];
      float w = 0.f;
      if (f == 1 || f == 2) w = expf(e_score[e] - mstar);
      if (f == 3) w = expf(MASK_FILL - mstar);
      e_score[e] = w;            // from here on the slot holds the entry's softmax numerator
      ls += w;
    }
    const float lstar = warp_sum(ls);
    if (lane == 0) {
      s_m = mstar;
      s_l = lstar;
      if (p.out.d_lse_m) p.out.d_lse_m[(size_t)s * K + k] = mstar;
      if (p.out.d_lse_l) p.out.d_lse_l[(size_t)s * K + k] = lstar;
    }
  }
  __syncthreads();
  const float mstar = s_m, inv_l = 1.f / s_l;

  // ---- 2. afeat of this branch: warp w takes ranks w, w + 8, ... and candidate entries w, w + 8, ...; lane l owns
  // features j0 + 4 l .. + 3 (one 16-byte load per row), four rows in flight per round ----
  const int nq = L / 4;
  for (int j0 = 0; j0 < L; j0 += 128) {
    const int jf = j0 + 4 * lane;
    const bool in = jf < L;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = warp; r0 < P; r0 += 4 * (RT / 32)) {
      float mv[4];
      float4 av[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * (RT / 32);
        const float* rc = recp(r < P ? r : 0);
        mv[u] = r < P ? __ldcg(rc + p.rec.m() + k) : -INFINITY;
        av[u] = (r < P && in) ? __ldcg(reinterpret_cast<const float4*>(rc + p.rec.acc() + (size_t)k * L + jf))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float w = mv[u] == -INFINITY ? 0.f : expf(mv[u] - mstar);
        a.x = fmaf(w, av[u].x, a.x); a.y = fmaf(w, av[u].y, a.y); a.z = fmaf(w, av[u].z, a.z); a.w = fmaf(w, av[u].w, a.w);
      }
    }
    for (int e0 = warp; e0 < ne; e0 += 4 * (RT / 32)) {
      float wv[4];
      float4 hv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * (RT / 32);
        wv[u] = e < ne ? e_score[e] : 0.f;
        const float* hr = recp((e < ne ? e : 0) / nmc) + p.rec.h() + ((size_t)k * nmc + (e < ne ? e : 0) % nmc) * L;
        hv[u] = (wv[u] != 0.f && in) ? __ldcg(reinterpret_cast<const float4*>(hr + jf)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a.x = fmaf(wv[u], hv[u].x, a.x); a.y = fmaf(wv[u], hv[u].y, a.y);
        a.z = fmaf(wv[u], hv[u].z, a.z); a.w = fmaf(wv[u], hv[u].w, a.w);
      }
    }
    if (in) wpart[warp * nq + jf / 4] = a;
  }
  __syncthreads();
  for (int qd = tid; qd < nq; qd += RT) {      // fixed-order sum over the 8 warps: deterministic
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < RT / 32; ++w) {
      const float4 u = wpart[w * nq + qd];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    t.x *= inv_l; t.y *= inv_l; t.z *= inv_l; t.w *= inv_l;
    *reinterpret_cast<float4*>(af + qd * 4) = t;
    if (p.out.d_afeat) *reinterpret_cast<float4*>(p.out.d_afeat + ((size_t)s * K + k) * L + qd * 4) = t;
  }
  __syncthreads();

  // ---- 3. this branch's logits: one warp per class ----
  const int C = p.heads.n_class;
  if (p.out.d_sub && (p.heads.n_branch_heads > 0 || p.heads.shared_head)) {
    for (int c = warp; c < C; c += RT / 32) {
      const float* w = p.heads.n_branch_heads > 0 ? p.heads.d_wc + ((size_t)k * C + c) * L : p.heads.d_ws + (size_t)c * L;
      const float b = p.heads.n_branch_heads > 0 ? p.heads.d_bc[k * C + c] : p.heads.d_bs[c];
      float t = 0.f;
      for (int jf = lane; jf < L; jf += 32) t = fmaf(w[jf], af[jf], t);
      t = warp_sum(t);
      if (lane == 0) p.out.d_sub[((size_t)s * K + k) * C + c] = t + b;
    }
  }
}

// bag feature = mean over branches of afeat (== mean_k softmax(A_k) @ h, transformer.py:328-329) and the slide head;
// one CTA per bag, behind gp_finish_kernel.  With the exchange, block 0 also moves this rank on to the next step.
__global__ void __launch_bounds__(128) gp_heads_kernel(const __grid_constant__ GpFinishParams p) {
  extern __shared__ __align__(16) unsigned char dsm[];
  float* bag = reinterpret_cast<float*>(dsm);      // [L]
  const int L = p.sh.d_inner, K = p.sh.n_branch, s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int jf = tid; jf < L; jf += 128) {
    float t = 0.f;
    for (int k = 0; k < K; ++k) t += p.out.d_afeat[((size_t)s * K + k) * L + jf];
    t /= (float)K;
    bag[jf] = t;
    if (p.out.d_bag_feat) p.out.d_bag_feat[(size_t)s * L + jf] = t;
  }
  __syncthreads();
  const int C = p.heads.n_class;
  if (p.out.d_slide && p.heads.slide_head) {
    for (int c = warp; c < C; c += 4) {
      float t = 0.f;
      for (int jf = lane; jf < L; jf += 32) t = fmaf(p.heads.d_ws[(size_t)c * L + jf], bag[jf], t);
      t = warp_sum(t);
      if (lane == 0) p.out.d_slide[(size_t)s * C + c] = t + p.heads.d_bs[c];
    }
  }
  if (p.x.n_ranks > 0 && s == 0 && tid == 0) *p.x.epoch = *p.x.epoch + 1u;      // (every finish CTA of this step has retired)
}

// ------------------------------------------------------------------------------------------
// attention statistics of one bag: gram[i][j] = sum_n P_i P_j, ent[k] = sum_n P_k log P_k
struct GpStatsParams {
  const float* a;
  int64_t a_ld;
  int K, S;
  int64_t row_off[SMAX + 1];
  const float* m;
  const float* l;
  float* gram;
  float* ent;
  float* div;
};

__global__ void __launch_bounds__(512) gp_stats_kernel(const __grid_constant__ GpStatsParams p) {
  __shared__ float red[16];
  __shared__ float g_s[KMAX * KMAX];
  const int s = blockIdx.x, K = p.K, tid = threadIdx.x;
  const int64_t r0 = p.row_off[s], n = p.row_off[s + 1] - r0;
  float mk[KMAX], il[KMAX], logl[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    mk[k] = k < K ? p.m[(size_t)s * K + k] : 0.f;
    const float l = k < K ? p.l[(size_t)s * K + k] : 1.f;
    il[k] = 1.f / l;
    logl[k] = logf(l);
  }
  float g[KMAX * (KMAX + 1) / 2], en[KMAX];
#pragma unroll
  for (int i = 0; i < KMAX * (KMAX + 1) / 2; ++i) g[i] = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) en[k] = 0.f;
  for (int64_t r = tid; r < n; r += blockDim.x) {
    float pv[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      pv[k] = 0.f;
      if (k < K) {
        const float z = p.a[(size_t)k * p.a_ld + r0 + r] - mk[k];
        pv[k] = expf(z) * il[k];
        // p log p with p == 0 contributes 0 (torch: softmax * log_softmax = 0 * (-1e9) = -0)
        en[k] += pv[k] * (z - logl[k]);
      }
    }
    int q = 0;
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
#pragma unroll
      for (int j = i; j < KMAX; ++j) {
        g[q] = fmaf(pv[i], pv[j], g[q]);
        ++q;
      }
  }
  int q = 0;
#pragma unroll
  for (int i = 0; i < KMAX; ++i)
#pragma unroll
    for (int j = i; j < KMAX; ++j) {
      const float t = block_sum(g[q++], red);
      if (tid == 0 && i < K && j < K) {
        g_s[i * KMAX + j] = t;
        g_s[j * KMAX + i] = t;
        if (p.gram) { p.gram[((size_t)s * K + i) * K + j] = t; p.gram[((size_t)s * K + j) * K + i] = t; }
      }
    }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const float t = block_sum(en[k], red);
    if (tid == 0 && k < K && p.ent) p.ent[(size_t)s * K + k] = t;
  }
  if (tid == 0 && p.div) {
    float d = 0.f;
    for (int i = 0; i < K; ++i)
      for (int j = i + 1; j < K; ++j)
        d += g_s[i * KMAX + j] / (fmaxf(sqrtf(g_s[i * KMAX + i]), 1e-8f) * fmaxf(sqrtf(g_s[j * KMAX + j]), 1e-8f)) /
             (float)(K * (K - 1) / 2);
    p.div[s] = d;
  }
}

__global__ void __launch_bounds__(512) softmax_rows_kernel(const float* __restrict__ a, int64_t a_ld, int64_t n,
                                                           float* __restrict__ out, int64_t out_ld) {
  __shared__ float red[16];
  const float* row = a + (size_t)blockIdx.x * a_ld;
  float* orow = out + (size_t)blockIdx.x * out_ld;
  float mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, row[i]);
  mx = block_max(mx, red);
  float sm = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) sm += expf(row[i] - mx);
  sm = block_sum(sm, red);
  const float inv = 1.f / sm;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) orow[i] = expf(row[i] - mx) * inv;
}

}  // namespace

int gp_launch_reduce(const GpMainParams& mp, const GpRecord& rec, float* d_record, const int* d_flags, int flag_mode,
                     const GpExchange* x, cudaStream_t st) {
  GpReduceParams p;
  memset(&p.x, 0, sizeof(p.x));
  if (x) p.x = *x;
  p.flags = d_flags;
  p.flag_mode = d_flags ? flag_mode : GP_REDUCE_ALL;
  p.sh = mp.sh;
  p.ws = mp.ws;
  p.wl = mp.wl;
  p.seg = mp.seg;
  p.rec = rec;
  p.record = d_record;
  const int S = mp.seg.n_slides, K = mp.sh.n_branch;
  if (S == 0) return ACMIL_OK;
  int max_cand = 0;
  for (int s = 0; s < S; ++s) {
    const int c = (mp.seg.seg_begin[s + 1] - mp.seg.seg_begin[s]) / mp.seg.cand_div * mp.seg.nm[s];
    if (c > max_cand) max_cand = c;
  }
  int max_seg = 0;
  for (int s = 0; s < S; ++s) max_seg = std::max(max_seg, mp.seg.seg_begin[s + 1] - mp.seg.seg_begin[s]);
  const bool small = max_seg <= 8 && max_cand <= 128;
  const int rr = small ? 256 : 1024;
  const size_t smem = (((size_t)max_cand * 16 + 15) / 16) * 16 + (size_t)(rr / 32) * mp.sh.d_inner * 4 + 32;
  ACMIL_REQUIRE(smem <= 220 * 1024, ACMIL_E_INVALID, "reduce: too many candidates per bag (%d)", max_cand);
  const int nj = (mp.sh.d_inner / 4 + 31) / 32;      // float4 groups of a feature row per lane
#define REDUCE_CASE(RRV, NJV)                                                                                              \
  {                                                                                                                        \
    static size_t configured = 48 * 1024;                                                                                  \
    if (smem > configured) {                                                                                               \
      ACMIL_CHECK_CUDA(cudaFuncSetAttribute(gp_reduce_kernel<RRV, NJV>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                            (int)smem));                                                                   \
      configured = smem;                                                                                                   \
    }                                                                                                                      \
    gp_reduce_kernel<RRV, NJV><<<S * K, RRV, smem, st>>>(p);                                                               \
  }
  if (small) {
    if (nj <= 1) REDUCE_CASE(256, 1) else if (nj <= 2) REDUCE_CASE(256, 2) else REDUCE_CASE(256, 4)
  } else {
    if (nj <= 1) REDUCE_CASE(1024, 1) else if (nj <= 2) REDUCE_CASE(1024, 2) else REDUCE_CASE(1024, 4)
  }
#undef REDUCE_CASE
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

int gp_launch_finish(const GpFinishParams& p, cudaStream_t st) {
  if (p.n_slides == 0) return ACMIL_OK;
  const int K = p.sh.n_branch, L = p.sh.d_inner;
  const size_t ne = (size_t)p.n_ranks * p.rec.nmc;
  const size_t smem = ((size_t)(RT / 32) * L + L + 3 * ne + NMAX) * 4 + 16;
  ACMIL_REQUIRE(smem <= 200 * 1024, ACMIL_E_INVALID, "finish: n_ranks * n_masked too large (%zu entries)", ne);
  ACMIL_REQUIRE(p.out.d_afeat != nullptr, ACMIL_E_INVALID, "outputs: d_afeat is required");
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    ACMIL_CHECK_CUDA(cudaFuncSetAttribute(gp_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  gp_finish_kernel<<<p.n_slides * K, RT, smem, st>>>(p);
  gp_heads_kernel<<<p.n_slides, 128, (size_t)L * 4, st>>>(p);
  g_acmil_launches += 2;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

int gp_launch_stats(const float* d_a, int64_t a_ld, int K, const int64_t* row_offsets, int S, const float* d_m,
                    const float* d_l, float* d_gram, float* d_ent, float* d_div, cudaStream_t st) {
  if (S == 0) return ACMIL_OK;
  GpStatsParams p;
  p.a = d_a; p.a_ld = a_ld; p.K = K; p.S = S;
  for (int s = 0; s <= S; ++s) p.row_off[s] = row_offsets[s];
  p.m = d_m; p.l = d_l; p.gram = d_gram; p.ent = d_ent; p.div = d_div;
  gp_stats_kernel<<<S, 512, 0, st>>>(p);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

int gp_launch_softmax_rows(const float* d_a, int64_t a_ld, int n_rows, int64_t n, float* d_out, int64_t out_ld,
                           cudaStream_t st) {
  if (n_rows == 0 || n == 0) return ACMIL_OK;
  softmax_rows_kernel<<<n_rows, 512, 0, st>>>(d_a, a_ld, n, d_out, out_ld);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}
