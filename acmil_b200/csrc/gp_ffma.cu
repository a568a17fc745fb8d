// Exact-fp32 fused gated-attention pool pass on CUDA cores (FFMA).
//
// This is the general-shape implementation (any d_in % 32 == 0, d_inner % 128 == 0, d_attn == 128,
// all activation / bias variants of SURVEY.md section 8a rows a1-a9) and the on-device fp32 yardstick
// for the tcgen05 kernel (gp_umma.cu).  One CTA owns a segment (run of 64-row tiles of one bag):
//
//   x tile --(W1 chunks streamed from L2)--> h = act(x W1^T + b1)   [64 x L]   smem       network.py:49-57
//   h --(Wv, Wu chunks)--> g = act_a(hWv^T+bv) * sigmoid(hWu^T+bu)  registers             transformer.py:261-262
//   g --> scores s = g Ww^T + bw  [64 x K]  -> a_out (raw)                                transformer.py:263-264
//   running top-n candidates per branch are kept OUT of the sums (their h rows parked in scratch)
//   online softmax over the rows: m, l, acc[k][:] += exp(s - m) h                         transformer.py:323-324
//
// Masked rows are decided later (gp_reduce.cu) from the candidate lists, so nothing is ever
// subtracted: rows that turn out not to be masked are added back exactly.
#include <cuda_fp16.h>

#include "gp_common.cuh"

namespace {

// four consecutive x values at element offset `off` of the tile: fp32 rows, or fp16 rows widened on load
__device__ __forceinline__ float4 load_x4(const float* __restrict__ xf, const __half* __restrict__ xh, size_t off) {
  if (xh == nullptr) return __ldg(reinterpret_cast<const float4*>(xf + off));
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(xh + off));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

constexpr int FR = 64;    // rows per tile
constexpr int FKC = 32;   // k chunk
constexpr int FNB = 128;  // output column block
constexpr int FT = 256;   // threads per CTA
constexpr int XLD = FKC + 4;

struct Smem {
  float* hs;   // [FR][L+4]
  float* xs;   // [FR][XLD]
  float* ws;   // [FKC][FNB]
  float* wws;  // [KMAX][128]
  float* ss;   // [FR][KMAX]  scores
  float* ps;   // [FR][KMAX]  softmax numerators
  float* m_run;
  float* l_run;
  float* scale_s;
  float* c_score;  // [KMAX][NMAX]
  int* c_idx;
  int* c_slot;
  int* c_cnt;        // [KMAX]
  unsigned* c_free;  // [KMAX]
  float* ev_w;       // [KMAX][NMAX] score, then weight, of entries evicted this tile
  int* ev_slot;      // [KMAX][NMAX]
  int* ev_cnt;       // [KMAX]
  int* ne_row;       // [KMAX][NMAX] rows of this tile that entered a list
  int* ne_slot;      // [KMAX][NMAX]
  int* ne_cnt;       // [KMAX]
  unsigned* exb;     // [KMAX][2] excluded-row bitmask of this tile
};

__host__ __device__ inline size_t smem_floats(int L) {
  return (size_t)FR * (L + 4) + FR * XLD + FKC * FNB + KMAX * 128 + 2 * FR * KMAX + 3 * KMAX +
         3 * KMAX * NMAX + 2 * KMAX + 2 * KMAX * NMAX + KMAX + 2 * KMAX * NMAX + KMAX + 2 * KMAX;
}

__device__ inline Smem carve(float* base, int L) {
  Smem s;
  float* p = base;
  s.hs = p; p += FR * (L + 4);
  s.xs = p; p += FR * XLD;
  s.ws = p; p += FKC * FNB;
  s.wws = p; p += KMAX * 128;
  s.ss = p; p += FR * KMAX;
  s.ps = p; p += FR * KMAX;
  s.m_run = p; p += KMAX;
  s.l_run = p; p += KMAX;
  s.scale_s = p; p += KMAX;
  s.c_score = p; p += KMAX * NMAX;
  s.c_idx = (int*)p; p += KMAX * NMAX;
  s.c_slot = (int*)p; p += KMAX * NMAX;
  s.c_cnt = (int*)p; p += KMAX;
  s.c_free = (unsigned*)p; p += KMAX;
  s.ev_w = p; p += KMAX * NMAX;
  s.ev_slot = (int*)p; p += KMAX * NMAX;
  s.ev_cnt = (int*)p; p += KMAX;
  s.ne_row = (int*)p; p += KMAX * NMAX;
  s.ne_slot = (int*)p; p += KMAX * NMAX;
  s.ne_cnt = (int*)p; p += KMAX;
  s.exb = (unsigned*)p; p += 2 * KMAX;
  return s;
}

// C[4][8] += A[rows ty*4..][k] * B[k][tx*8..]   over one FKC-wide chunk
__device__ __forceinline__ void chunk_fma(float (&acc)[4][8], const float* a_base, int a_ld, const float* ws, int tx) {
#pragma unroll 8
  for (int k = 0; k < FKC; ++k) {
    float a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = a_base[i * a_ld + k];
    const float4 b0 = *reinterpret_cast<const float4*>(ws + k * FNB + tx * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(ws + k * FNB + tx * 8 + 4);
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(a[i], b[c], acc[i][c]);
  }
}

// ws[k][0..127] <- wt[(k0+k)*ld + n0 .. n0+127]
__device__ __forceinline__ void load_w_chunk(float* ws, const float* __restrict__ wt, int ld, int k0, int n0, int tid) {
#pragma unroll
  for (int i = 0; i < (FKC * FNB / 4) / FT; ++i) {
    const int idx = tid + FT * i;
    const int k = idx >> 5, c4 = idx & 31;
    const float4 v = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(k0 + k) * ld + n0 + c4 * 4));
    *reinterpret_cast<float4*>(ws + k * FNB + c4 * 4) = v;
  }
}

__device__ __forceinline__ int nth_set_bit(unsigned mask, int n) {
  for (int i = 0; i < n; ++i) mask &= mask - 1;
  return __ffs(mask) - 1;
}

__global__ void __launch_bounds__(FT) gp_main_ffma_kernel(const __grid_constant__ GpMainParams p) {
  extern __shared__ __align__(16) float smem_raw[];
  const int L = p.sh.d_inner, K = p.sh.n_branch, DIN = p.sh.d_in;
  const int ldh = L + 4;
  Smem sm = carve(smem_raw, L);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;

  // ---- which segments: blockIdx.x, blockIdx.x + gridDim.x, ... (one each unless this is a rescue pass, which is launched
  // with a small grid because it normally has nothing to do) ----
  if (p.rescue_flags != nullptr) {
    // rescue pass: normally no bag is flagged -- find that out with one look at the flags instead of walking the segments
    int any = 0;
    for (int b = tid; b < p.seg.n_slides; b += FT) any |= p.rescue_flags[b] == 1;
    if (!__syncthreads_or(any)) return;
  }
  int s = 0;
  for (int seg = blockIdx.x; seg < p.seg.n_seg; seg += gridDim.x) {
  while (seg >= p.seg.seg_begin[s + 1]) ++s;
  if (p.rescue_flags != nullptr && p.rescue_flags[s] != 1) continue;   // rescue pass: only the bags the tcgen05 kernel gave up on
  __syncthreads();      // (the previous segment's last reads of shared memory are done)
  const int j = seg - p.seg.seg_begin[s];
  const int64_t row0_bag = p.seg.row_off[s];
  const int64_t n_rows = p.seg.row_off[s + 1] - row0_bag;
  const int tiles = (int)((n_rows + FR - 1) / FR);
  const int tps = p.seg.tiles_per_seg[s];
  const int t0 = j * tps, t1 = min(t0 + tps, tiles);
  const int nm = p.seg.nm[s];
  const int cap = p.seg.n_masked_cap;

  const float* __restrict__ w1t = p.pack + p.lay.w1t;
  const float* __restrict__ b1 = p.pack + p.lay.b1;
  const float* __restrict__ wvt = p.pack + p.lay.wvt;
  const float* __restrict__ wut = p.pack + p.lay.wut;
  const float* __restrict__ bv = p.pack + p.lay.bv;
  const float* __restrict__ bu = p.pack + p.lay.bu;
  const float* __restrict__ ww = p.pack + p.lay.ww;
  const float* __restrict__ bw = p.pack + p.lay.bw;
  float* cand_h = reinterpret_cast<float*>(p.ws + p.wl.cand_h) + (size_t)seg * K * cap * L;

  for (int i = tid; i < KMAX * 128; i += FT) sm.wws[i] = ww[i];
  for (int i = tid; i < 2 * FR * KMAX; i += FT) sm.ss[i] = 0.f;  // ss and ps (columns >= K stay 0)
  if (tid < KMAX) {
    sm.m_run[tid] = -INFINITY;
    sm.l_run[tid] = 0.f;
    sm.c_cnt[tid] = 0;
    sm.c_free[tid] = nm >= 32 ? 0xffffffffu : ((1u << nm) - 1u);
  }
  float pacc[2][KMAX];
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int k = 0; k < KMAX; ++k) pacc[f][k] = 0.f;
  __syncthreads();

  for (int t = t0; t < t1; ++t) {
    const int64_t trow = (int64_t)t * FR;  // first row of the tile inside the bag
    const int valid = (int)min((int64_t)FR, n_rows - trow);
    // fp32 rows, or fp16 rows widened on load (exact)
    const float* __restrict__ xt = p.x_f16 ? nullptr : p.x + (size_t)(row0_bag + trow) * DIN;
    const __half* __restrict__ xth = p.x_f16 ? reinterpret_cast<const __half*>(p.x) + (size_t)(row0_bag + trow) * DIN : nullptr;

    // ================= stage 1: h tile =================
    if (p.sh.front) {
      for (int nb = 0; nb < L / FNB; ++nb) {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
        for (int kc = 0; kc < DIN / FKC; ++kc) {
#pragma unroll
          for (int i = 0; i < (FR * FKC / 4) / FT; ++i) {
            const int idx = tid + FT * i;
            const int r = idx >> 3, kq = idx & 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < valid) v = load_x4(xt, xth, (size_t)r * DIN + kc * FKC + kq * 4);
            *reinterpret_cast<float4*>(sm.xs + r * XLD + kq * 4) = v;
          }
          load_w_chunk(sm.ws, w1t, L, kc * FKC, nb * FNB, tid);
          __syncthreads();
          chunk_fma(acc, sm.xs + (ty * 4) * XLD, XLD, sm.ws, tx);
          __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int col = nb * FNB + tx * 8 + c;
            float z = acc[i][c] + b1[col];
            sm.hs[(ty * 4 + i) * ldh + col] = act_apply(z, p.sh.front_act);
          }
      }
    } else {
      for (int idx = tid; idx < FR * (L / 4); idx += FT) {
        const int r = idx / (L / 4), c4 = idx % (L / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < valid) v = load_x4(xt, xth, (size_t)r * DIN + c4 * 4);
        *reinterpret_cast<float4*>(sm.hs + r * ldh + c4 * 4) = v;
      }
    }
    __syncthreads();

    // ================= stage 2: gate + scores =================
    {
      float accv[4][8], accu[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) accv[i][c] = accu[i][c] = 0.f;
      if (p.z != nullptr) {
        // the caller computed h Wv^T (and h Wu^T) on the tensor cores: this thread's 4 rows x 8 units come from there
        const int zc = p.sh.gated ? 2 * GP_DATTN : GP_DATTN;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = ty * 4 + i;
          if (r < valid) {
            const float* zr = p.z + (size_t)(row0_bag + trow + r) * zc + tx * 8;
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(zr)), a1 = __ldg(reinterpret_cast<const float4*>(zr + 4));
            accv[i][0] = a0.x; accv[i][1] = a0.y; accv[i][2] = a0.z; accv[i][3] = a0.w;
            accv[i][4] = a1.x; accv[i][5] = a1.y; accv[i][6] = a1.z; accv[i][7] = a1.w;
            if (p.sh.gated) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(zr + GP_DATTN)), b1 = __ldg(reinterpret_cast<const float4*>(zr + GP_DATTN + 4));
              accu[i][0] = b0.x; accu[i][1] = b0.y; accu[i][2] = b0.z; accu[i][3] = b0.w;
              accu[i][4] = b1.x; accu[i][5] = b1.y; accu[i][6] = b1.z; accu[i][7] = b1.w;
            }
          }
        }
      } else {
      for (int kc = 0; kc < L / FKC; ++kc) {
        load_w_chunk(sm.ws, wvt, GP_DATTN, kc * FKC, 0, tid);
        __syncthreads();
        chunk_fma(accv, sm.hs + (ty * 4) * ldh + kc * FKC, ldh, sm.ws, tx);
        __syncthreads();
      }
      if (p.sh.gated) {
        for (int kc = 0; kc < L / FKC; ++kc) {
          load_w_chunk(sm.ws, wut, GP_DATTN, kc * FKC, 0, tid);
          __syncthreads();
          chunk_fma(accu, sm.hs + (ty * 4) * ldh + kc * FKC, ldh, sm.ws, tx);
          __syncthreads();
        }
      }
      }
      float part[4][KMAX];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < KMAX; ++k) part[i][k] = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = tx * 8 + c;
        const float bvc = bv[col], buc = bu[col];
        float wk[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) wk[k] = sm.wws[k * 128 + col];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float g = act_apply(accv[i][c] + bvc, p.sh.act_a);
          if (p.sh.gated) g *= sigmoid_acc(accu[i][c] + buc);
#pragma unroll
          for (int k = 0; k < KMAX; ++k) part[i][k] = fmaf(g, wk[k], part[i][k]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          float v = part[i][k];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          if (tx == 0 && k < K) sm.ss[(ty * 4 + i) * KMAX + k] = v + bw[k];
        }
    }
    __syncthreads();

    // ================= stage 3: raw scores out, candidates, online softmax =================
    if (p.a_out != nullptr) {
      for (int idx = tid; idx < K * FR; idx += FT) {
        const int k = idx / FR, r = idx % FR;
        if (r < valid) p.a_out[(size_t)k * p.a_ld + row0_bag + trow + r] = sm.ss[r * KMAX + k];
      }
    }
    if (warp < K) {
      const int k = warp;
      float sv[FR / 32];
#pragma unroll
      for (int h = 0; h < FR / 32; ++h) {
        const int r = h * 32 + lane;
        sv[h] = r < valid ? sm.ss[r * KMAX + k] : -INFINITY;
      }
      int n_ev = 0, n_ne = 0;
      unsigned ex[FR / 32];
#pragma unroll
      for (int h = 0; h < FR / 32; ++h) ex[h] = 0u;
      if (nm > 0) {
        int cnt = sm.c_cnt[k];
        float e_s = lane < cnt ? sm.c_score[k * NMAX + lane] : -INFINITY;
        int e_i = lane < cnt ? sm.c_idx[k * NMAX + lane] : 0x7fffffff;
        int e_sl = lane < cnt ? sm.c_slot[k * NMAX + lane] : 0;
        unsigned freed = 0u;
#pragma unroll
        for (int h = 0; h < FR / 32; ++h) {
          const float tau = cnt == nm ? __shfl_sync(0xffffffffu, e_s, nm - 1) : -INFINITY;
          const int r = h * 32 + lane;
          unsigned bal = __ballot_sync(0xffffffffu, r < valid && (cnt < nm || sv[h] > tau));
          while (bal) {
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            const float s_new = __shfl_sync(0xffffffffu, sv[h], src);
            const int rr = h * 32 + src;
            if (cnt == nm) {
              const float last_s = __shfl_sync(0xffffffffu, e_s, nm - 1);
              if (!(s_new > last_s)) continue;  // rows arrive in increasing index: ties lose
              const int last_sl = __shfl_sync(0xffffffffu, e_sl, nm - 1);
              if (last_sl >= 0) {  // an entry parked in scratch falls out: add it back this tile
                if (lane == 0) {
                  sm.ev_w[k * NMAX + n_ev] = last_s;
                  sm.ev_slot[k * NMAX + n_ev] = last_sl;
                }
                freed |= 1u << last_sl;
                ++n_ev;
              }
            }
            const int pos = __popc(__ballot_sync(0xffffffffu, lane < cnt && e_s >= s_new));
            const float up_s = __shfl_up_sync(0xffffffffu, e_s, 1);
            const int up_i = __shfl_up_sync(0xffffffffu, e_i, 1);
            const int up_sl = __shfl_up_sync(0xffffffffu, e_sl, 1);
            if (lane > pos) { e_s = up_s; e_i = up_i; e_sl = up_sl; }
            if (lane == pos) { e_s = s_new; e_i = (int)trow + rr; e_sl = -1 - rr; }
            if (cnt < nm) ++cnt;
          }
        }
        // rows of this tile that stayed in the list: give them scratch slots, exclude them from the sums
        const bool is_new = lane < cnt && e_sl < 0;
        const unsigned newmask = __ballot_sync(0xffffffffu, is_new);
        unsigned freemask = sm.c_free[k] | freed;
        n_ne = __popc(newmask);
        if (is_new) {
          const int rank = __popc(newmask & ((1u << lane) - 1u));
          const int slot = nth_set_bit(freemask, rank);
          const int rr = -1 - e_sl;
          sm.ne_row[k * NMAX + rank] = rr;
          sm.ne_slot[k * NMAX + rank] = slot;
          e_sl = slot;
        }
        // slots in use afterwards = slots of all list entries
        unsigned inuse = 0u;
        for (int i = 0; i < cnt; ++i) inuse |= 1u << __shfl_sync(0xffffffffu, e_sl, i);
        const unsigned all = nm >= 32 ? 0xffffffffu : ((1u << nm) - 1u);
        freemask = all & ~inuse;
        __syncwarp();
        // excluded-row bitmask of this tile (rows that now sit in the list)
#pragma unroll
        for (int h = 0; h < FR / 32; ++h) {
          unsigned m_h = 0u;
          for (int i = 0; i < n_ne; ++i) {
            const int rr = sm.ne_row[k * NMAX + i];
            if ((rr >> 5) == h) m_h |= 1u << (rr & 31);
          }
          ex[h] = m_h;
        }
        if (lane < cnt) {
          sm.c_score[k * NMAX + lane] = e_s;
          sm.c_idx[k * NMAX + lane] = e_i;
          sm.c_slot[k * NMAX + lane] = e_sl;
        }
        if (lane == 0) {
          sm.c_cnt[k] = cnt;
          sm.c_free[k] = freemask;
        }
      }
      // online softmax for branch k over the rows that take part (parked rows do not, entries that
      // fell out of the list this tile do); m is the exact max of everything summed so far
      float tmax = -INFINITY;
#pragma unroll
      for (int h = 0; h < FR / 32; ++h)
        if (!((ex[h] >> lane) & 1u)) tmax = fmaxf(tmax, sv[h]);
      __syncwarp();
      if (lane < n_ev) tmax = fmaxf(tmax, sm.ev_w[k * NMAX + lane]);
      tmax = warp_max(tmax);
      const float m_old = sm.m_run[k];
      const float m_new = fmaxf(m_old, tmax);
      const float sc = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
      float lsum = 0.f;
#pragma unroll
      for (int h = 0; h < FR / 32; ++h) {
        const int r = h * 32 + lane;
        float pv = 0.f;
        if (r < valid && !((ex[h] >> lane) & 1u) && m_new != -INFINITY) pv = expf(sv[h] - m_new);
        sm.ps[r * KMAX + k] = pv;
        lsum += pv;
      }
      if (lane < n_ev) {
        const float w = expf(sm.ev_w[k * NMAX + lane] - m_new);
        sm.ev_w[k * NMAX + lane] = w;
        lsum += w;
      }
      lsum = warp_sum(lsum);
      if (lane == 0) {
        sm.l_run[k] = sm.l_run[k] * sc + lsum;
        sm.m_run[k] = m_new;
        sm.scale_s[k] = sc;
        sm.ev_cnt[k] = n_ev;
        sm.ne_cnt[k] = n_ne;
      }
    }
    __syncthreads();

    // ================= stage 4: weighted reduce =================
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int jf = tid + FT * f;
      if (jf < L) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) pacc[f][k] *= sm.scale_s[k < K ? k : 0];
        for (int r = 0; r < FR; ++r) {
          const float hv = sm.hs[r * ldh + jf];
          const float4 p0 = *reinterpret_cast<const float4*>(sm.ps + r * KMAX);
          const float4 p1 = *reinterpret_cast<const float4*>(sm.ps + r * KMAX + 4);
          pacc[f][0] = fmaf(p0.x, hv, pacc[f][0]);
          pacc[f][1] = fmaf(p0.y, hv, pacc[f][1]);
          pacc[f][2] = fmaf(p0.z, hv, pacc[f][2]);
          pacc[f][3] = fmaf(p0.w, hv, pacc[f][3]);
          pacc[f][4] = fmaf(p1.x, hv, pacc[f][4]);
          pacc[f][5] = fmaf(p1.y, hv, pacc[f][5]);
          pacc[f][6] = fmaf(p1.z, hv, pacc[f][6]);
          pacc[f][7] = fmaf(p1.w, hv, pacc[f][7]);
        }
        if (nm > 0) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
              const int ne = sm.ev_cnt[k];
              for (int e = 0; e < ne; ++e) {
                const int slot = sm.ev_slot[k * NMAX + e];
                pacc[f][k] = fmaf(sm.ev_w[k * NMAX + e], cand_h[((size_t)k * cap + slot) * L + jf], pacc[f][k]);
              }
            }
          }
        }
      }
    }
    if (nm > 0) {
      __syncthreads();  // late-adds have read their scratch rows before slots are reused
      for (int k = 0; k < K; ++k) {
        const int ne = sm.ne_cnt[k];
        for (int e = 0; e < ne; ++e) {
          const int rr = sm.ne_row[k * NMAX + e], slot = sm.ne_slot[k * NMAX + e];
          for (int jf = tid; jf < L; jf += FT) cand_h[((size_t)k * cap + slot) * L + jf] = sm.hs[rr * ldh + jf];
        }
      }
    }
    __syncthreads();
  }

  // ---- segment partial record ----
  float* part = reinterpret_cast<float*>(p.ws + p.wl.part) + (size_t)seg * K * (L + 2);
  if (tid < K) {
    part[(size_t)tid * (L + 2) + 0] = sm.m_run[tid];
    part[(size_t)tid * (L + 2) + 1] = sm.l_run[tid];
  }
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const int jf = tid + FT * f;
    if (jf < L) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) part[(size_t)k * (L + 2) + 2 + jf] = pacc[f][k];
    }
  }
  if (nm > 0) {
    int* g_cnt = reinterpret_cast<int*>(p.ws + p.wl.cand_cnt) + (size_t)seg * K;
    float* g_score = reinterpret_cast<float*>(p.ws + p.wl.cand_score) + (size_t)seg * K * cap;
    int* g_idx = reinterpret_cast<int*>(p.ws + p.wl.cand_idx) + (size_t)seg * K * cap;
    int* g_slot = reinterpret_cast<int*>(p.ws + p.wl.cand_slot) + (size_t)seg * K * cap;
    if (tid < K) g_cnt[tid] = sm.c_cnt[tid];
    for (int idx = tid; idx < K * cap; idx += FT) {
      const int k = idx / cap, i = idx % cap;
      const bool live = i < sm.c_cnt[k];
      g_score[idx] = live ? sm.c_score[k * NMAX + i] : -INFINITY;
      g_idx[idx] = live ? sm.c_idx[k * NMAX + i] : 0x7fffffff;
      g_slot[idx] = live ? sm.c_slot[k * NMAX + i] : 0;
    }
  } else if (p.seg.n_masked_cap > 0) {
    int* g_cnt = reinterpret_cast<int*>(p.ws + p.wl.cand_cnt) + (size_t)seg * K;
    if (tid < K) g_cnt[tid] = 0;
  }
  }      // segments of this CTA
}

}  // namespace

int gp_launch_main_ffma(const GpMainParams& p, cudaStream_t st) {
  const acmil_gp_shape& s = p.sh;
  ACMIL_REQUIRE(s.d_attn == GP_DATTN, ACMIL_E_UNSUPPORTED, "ffma kernel: d_attn must be 128 (got %d)", s.d_attn);
  ACMIL_REQUIRE(s.d_inner % 128 == 0 && s.d_inner >= 128 && s.d_inner <= 512, ACMIL_E_UNSUPPORTED,
                "ffma kernel: d_inner must be 128, 256, 384 or 512 (got %d)", s.d_inner);
  ACMIL_REQUIRE(s.d_in % 32 == 0, ACMIL_E_UNSUPPORTED, "ffma kernel: d_in must be a multiple of 32 (got %d)", s.d_in);
  ACMIL_REQUIRE(s.front || s.d_in == s.d_inner, ACMIL_E_INVALID, "front == 0 needs d_in == d_inner");
  if (p.seg.n_seg == 0) return ACMIL_OK;
  const size_t smem = smem_floats(s.d_inner) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    ACMIL_CHECK_CUDA(cudaFuncSetAttribute(gp_main_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  if (p.rescue_flags == nullptr) ACMIL_CHECK_CUDA(cudaMemsetAsync(p.ws + p.wl.flags, 0, SMAX * 4, st));
  int grid = p.seg.n_seg;
  if (p.rescue_flags != nullptr) {      // rescue pass: one resident wave is plenty (CTAs loop over the segments)
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (grid > sms) grid = sms;
  }
  gp_main_ffma_kernel<<<grid, FT, smem, st>>>(p);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}
