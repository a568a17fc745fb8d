// Internal layouts shared by the gated-pool kernels (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "acmil_b200.h"

#define KMAX ACMIL_MAX_BRANCH
#define NMAX ACMIL_MAX_MASKED
#define SMAX ACMIL_MAX_SLIDES
#define GP_DATTN 128            // gate hidden width supported by the kernels
#define GP_MAX_SEG_CAND 12288   // nseg(slide) * n_masked bound for the reduce kernel's smem

// ----------------------------------------------------------------------------- errors
void acmil_set_error(const char* fmt, ...);
#include <atomic>
extern std::atomic<int64_t> g_acmil_launches;      // kernels launched by this process (any host thread)

#define ACMIL_CHECK_CUDA(expr)                                                          \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      acmil_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                        \
      return ACMIL_E_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define ACMIL_REQUIRE(cond, code, ...) \
  do {                                 \
    if (!(cond)) {                     \
      acmil_set_error(__VA_ARGS__);    \
      return (code);                   \
    }                                  \
  } while (0)

// ----------------------------------------------------------------------------- packed weights
// fp32 section (FFMA kernel + reduce/finish):   all offsets in floats
//   w1t [d_in][L]   k-major copy of W1 (L contiguous)       (front only)
//   b1  [L]         zeros when the layer has no bias
//   wvt [L][128], wut [L][128]  k-major copies of Wv, Wu
//   bv, bu [128]    zeros when absent
//   ww  [KMAX][128] rows >= K zero
//   bw  [KMAX]
// umma section (tcgen05 kernel): fp16 hi/lo operand images, see gp_umma.cu
struct GpPackLayout {
  size_t w1t, b1, wvt, wut, bv, bu, ww, bw;  // float offsets
  size_t f32_floats;
  size_t umma_off;    // byte offset of the umma section (1024-aligned)
  size_t umma_bytes;
  size_t total_bytes;
};

static inline GpPackLayout gp_pack_layout(const acmil_gp_shape& s) {
  GpPackLayout l;
  size_t o = 0;
  const size_t L = (size_t)s.d_inner, D = GP_DATTN;
  l.w1t = o; o += s.front ? (size_t)s.d_in * L : 0;
  l.b1 = o;  o += L;
  l.wvt = o; o += L * D;
  l.wut = o; o += L * D;
  l.bv = o;  o += D;
  l.bu = o;  o += D;
  l.ww = o;  o += (size_t)KMAX * D;
  l.bw = o;  o += KMAX;
  l.f32_floats = o;
  l.umma_off = ((o * 4 + 1023) / 1024) * 1024;
  // [absmax, pad to 1024][per-CTA images: hi/lo fp16 of a W1 half (64 x d_in) and of Wv or Wu (128 x 128)]
  // ... followed by 8 KB for the device-resident constants (UmmaConsts, gp_umma_shared.cuh)
  l.umma_bytes = 1024 + 2 * ((size_t)((s.d_in + 63) / 64) * 8192 * 2 + 65536) + 8192;
  l.total_bytes = l.umma_off + l.umma_bytes;
  return l;
}

// ----------------------------------------------------------------------------- segments
// The rows of each bag are cut into tiles of `tile_rows`; a "segment" is a run of consecutive
// tiles of ONE bag processed by one CTA, which emits one partial record for it.
struct GpSegTable {
  int32_t n_slides;
  int32_t n_seg;                 // total segments (= grid size of the FFMA kernel)
  int32_t tile_rows;
  int32_t n_masked_cap;          // candidate slots per (segment, branch)
  int64_t row_off[SMAX + 1];     // rows of bag s: [row_off[s], row_off[s+1])
  int64_t shard_begin[SMAX];     // global index (inside the bag) of the first local row
  int32_t seg_begin[SMAX + 1];   // segments of bag s: [seg_begin[s], seg_begin[s+1])
  int32_t tiles_per_seg[SMAX];
  int32_t nm[SMAX];              // candidates tracked per branch = min(n_masked, local rows)
  // tcgen05 kernel only: bags are cut into 256-row pair-tiles, cluster c owns the global pair-tiles
  // [c * u_total_pt / u_nclusters, (c + 1) * u_total_pt / u_nclusters); every (cluster, bag) it touches
  // produces 2 segments (one per CTA; its 8 epilogue warps merge their partials in shared memory):
  // seg_begin[s] + (c - u_cfirst[s]) * 2 + cta
  // candidate bookkeeping: lists live per "candidate holder" = group of cand_div consecutive segments
  // (FFMA: every segment; tcgen05: one per CTA and bag); each holder owns cand lists of n_masked_cap
  // entries per branch, rec_cap (score, row, h-slot) records per branch and row_cap parked h rows:
  // a record's h row sits in slot  k * h_branch_stride + cand_slot  of the holder
  // (FFMA: n_masked_cap private slots per branch; tcgen05: one slot per parked ROW, shared by the branches)
  int32_t cand_div;
  int32_t rec_cap;
  int32_t row_cap;
  int32_t h_branch_stride;
  int32_t u_nclusters;
  int32_t u_total_pt;
  int32_t u_pt_begin[SMAX + 1];
  int32_t u_cfirst[SMAX];
};

static inline int gp_build_segments(const acmil_gp_batch& b, int tile_rows, int target_seg, int s_branch, GpSegTable* t) {
  memset(t, 0, sizeof(*t));
  t->n_slides = b.n_slides;
  t->tile_rows = tile_rows;
  int64_t total_tiles = 0;
  for (int s = 0; s < b.n_slides; ++s) {
    int64_t n = b.row_offsets[s + 1] - b.row_offsets[s];
    if (n < 0) return -1;
    total_tiles += (n + tile_rows - 1) / tile_rows;
  }
  int64_t tps_global = (total_tiles + target_seg - 1) / (target_seg > 0 ? target_seg : 1);
  if (tps_global < 1) tps_global = 1;
  int cap = 0, seg = 0;
  for (int s = 0; s < b.n_slides; ++s) {
    int64_t n = b.row_offsets[s + 1] - b.row_offsets[s];
    int64_t tiles = (n + tile_rows - 1) / tile_rows;
    int nm = b.n_masked > 0 ? (int)(n < b.n_masked ? n : b.n_masked) : 0;
    int64_t tps = tps_global;
    // keep nseg * nm inside the reduce kernel's smem budget
    while (nm > 0 && ((tiles + tps - 1) / tps) * nm > GP_MAX_SEG_CAND) tps *= 2;
    t->row_off[s] = b.row_offsets[s];
    t->shard_begin[s] = b.shard_row_begin ? b.shard_row_begin[s] : 0;
    t->seg_begin[s] = seg;
    t->tiles_per_seg[s] = (int)tps;
    t->nm[s] = nm;
    if (nm > cap) cap = nm;
    seg += (int)((tiles + tps - 1) / tps);
  }
  t->row_off[b.n_slides] = b.row_offsets[b.n_slides];
  t->seg_begin[b.n_slides] = seg;
  t->n_seg = seg;
  t->n_masked_cap = cap;
  t->cand_div = 1;
  t->rec_cap = cap;
  t->row_cap = cap * s_branch;
  t->h_branch_stride = cap;
  return 0;
}

// ----------------------------------------------------------------------------- workspace
// per segment and branch:  part[seg][k][L+2] = {m, l, acc[L]}
// per candidate holder cb = seg / cand_div and branch:
//   cand_cnt[cb][k]; cand_score/idx/slot[cb][k][cap] (the holder's top-n list; slot = index of the parked row)
//   rec_score/rec_idx/rec_slot[cb][k][rec_cap] (every row ever parked; tcgen05 kernel only); cand_h[cb][row_cap][L]
//   flags[s] == 1: a holder of bag s ran out of parking slots in the tcgen05 kernel -> the bag is redone by the exact
//   FFMA kernel ("rescue" launch) and the reduce kernel takes that result.  flags sits right in front of cand_score so
//   that ONE memset (0xFF: flag -1 = clean, list entries NaN = "not written in this launch") prepares both.
struct GpWorkspace {
  size_t part, cand_cnt, flags, cand_score, cand_idx, cand_slot, rec_score, rec_idx, rec_slot, cand_h;  // byte offsets
  size_t total_bytes;
};

static inline GpWorkspace gp_workspace_layout(const acmil_gp_shape& s, const GpSegTable& t) {
  GpWorkspace w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
  const size_t K = s.n_branch, L = s.d_inner;
  const size_t n_seg = t.n_seg, cap = t.n_masked_cap, rcap = t.rec_cap, rowcap = t.row_cap;
  const size_t ncb = (n_seg + t.cand_div - 1) / (t.cand_div > 0 ? t.cand_div : 1);
  w.part = take(n_seg * K * (L + 2) * 4);
  w.cand_cnt = take(ncb * K * 4);
  w.flags = take(SMAX * 4);
  w.cand_score = take(ncb * K * cap * 4);
  w.cand_idx = take(ncb * K * cap * 4);
  w.cand_slot = take(ncb * K * cap * 4);
  w.rec_score = take(ncb * K * rcap * 4);
  w.rec_idx = take(ncb * K * rcap * 4);
  w.rec_slot = take(ncb * K * rcap * 4);
  w.cand_h = take(ncb * rowcap * L * 4);
  w.total_bytes = o + 256;
  return w;
}

// ----------------------------------------------------------------------------- rank partial record
// One record per bag, all 4-byte units (ints stored bit-exact in the same buffer so that one
// all-gather moves everything):
//   m[K] l[K] acc[K][L] cnt[K](int) score[K][nmc] idx[K][nmc](int, global row in bag) h[K][nmc][L]
// (each section padded to a multiple of 4 units)
// with nmc = n_masked requested (fixed per call so that every rank agrees on the stride).
struct GpRecord {
  int K, L, nmc;
  // every section starts on a 16-byte boundary (float4 loads of the acc / h rows in the finish kernel)
  static __host__ __device__ size_t pad4(size_t n) { return (n + 3) & ~(size_t)3; }
  __host__ __device__ size_t m() const { return 0; }
  __host__ __device__ size_t l() const { return pad4((size_t)K); }
  __host__ __device__ size_t acc() const { return l() + pad4((size_t)K); }
  __host__ __device__ size_t cnt() const { return acc() + pad4((size_t)K * L); }
  __host__ __device__ size_t score() const { return cnt() + pad4((size_t)K); }
  __host__ __device__ size_t idx() const { return score() + pad4((size_t)K * nmc); }
  __host__ __device__ size_t h() const { return idx() + pad4((size_t)K * nmc); }
  __host__ __device__ size_t stride() const { return h() + pad4((size_t)K * nmc * L); }
};

static inline GpRecord gp_record(const acmil_gp_shape& s, int n_masked) {
  GpRecord r;
  r.K = s.n_branch;
  r.L = s.d_inner;
  r.nmc = n_masked > 0 ? n_masked : 0;
  return r;
}

// ----------------------------------------------------------------------------- kernel launchers
struct GpMainParams {
  acmil_gp_shape sh;
  const float* x;
  float* a_out;
  int64_t a_ld;
  const float* pack;      // fp32 section of the packed weights
  GpPackLayout lay;
  unsigned char* ws;      // workspace base
  GpWorkspace wl;
  GpSegTable seg;
  // FFMA kernel as the rescue pass of the tcgen05 kernel: only the bags with rescue_flags[s] == 1 are processed
  const int* rescue_flags;
  int x_f16;              // x holds fp16 rows (acmil_gp_batch.x_f16)
  const float* z;         // optional gate pre-activations without biases (acmil_gp_batch.d_z), FFMA kernel only
};

// which bags a reduce launch handles, by the tcgen05 kernel's per-bag overflow flag
enum { GP_REDUCE_ALL = 0, GP_REDUCE_UNFLAGGED = 1, GP_REDUCE_FLAGGED = 2 };

// device-side view of acmil_gp_exchange (NULL-able: n_ranks == 0 means "no exchange, plain local record buffer")
struct GpExchange {
  int n_ranks, rank;
  float* gather[ACMIL_MAX_PEERS];
  uint32_t* flags[ACMIL_MAX_PEERS];
  uint32_t* epoch;
  unsigned* ticket;
  int reduce_ctas;       // CTAs of all reduce launches of one step (the last one to finish publishes the flags)
};

int gp_launch_main_ffma(const GpMainParams& p, cudaStream_t st);
int gp_launch_reduce(const GpMainParams& p, const GpRecord& rec, float* d_record, const int* d_flags, int flag_mode,
                     const GpExchange* x, cudaStream_t st);

struct GpFinishParams {
  acmil_gp_shape sh;
  GpRecord rec;
  const float* records;   // [n_ranks][S][stride]
  int n_ranks;
  int n_slides;
  int n_masked;
  int32_t keep[SMAX];
  const int64_t* rsel;    // [S][K][keep_ld]
  int keep_ld;
  const float* rand;      // alternative to rsel: the uniform draws [S][K][rand_ld]; the kernel takes argsort(rand)[:keep]
  int rand_ld;
  float* a_out;
  int64_t a_ld;
  int64_t row_off[SMAX + 1];
  int64_t shard_begin[SMAX];
  acmil_gp_heads heads;
  acmil_gp_outputs out;
  GpExchange x;           // n_ranks > 0: records = x.gather[x.rank] (parity from *x.epoch), wait for the peers' flags
};
int gp_launch_finish(const GpFinishParams& p, cudaStream_t st);
int gp_launch_stats(const float* d_a, int64_t a_ld, int K, const int64_t* row_offsets, int S, const float* d_m,
                    const float* d_l, float* d_gram, float* d_ent, float* d_div, cudaStream_t st);
int gp_launch_softmax_rows(const float* d_a, int64_t a_ld, int n_rows, int64_t n, float* d_out, int64_t out_ld,
                           cudaStream_t st);

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float act_apply(float z, int act) {
  if (act == ACMIL_ACT_TANH) return tanhf(z);
  if (act == ACMIL_ACT_RELU) return fmaxf(z, 0.f);
  return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));  // exact-erf GELU (nn.GELU default)
}
__device__ __forceinline__ float sigmoid_acc(float z) { return 1.f / (1.f + expf(-z)); }
// (score desc, index asc) ordering used everywhere a top-n is formed: torch.topk order
__device__ __forceinline__ bool cand_better(float s1, int i1, float s2, int i2) {
  return (s1 > s2) || (s1 == s2 && i1 < i2);
}
#endif
