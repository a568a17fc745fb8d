// Declarations of the tcgen05 row-pass kernel (gp_umma.cu).  (A role-specialised two-tiles-in-flight variant, gp_umma3.cu,
// lived next to it for a while: parity-green but slower; git history f79081b.)
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "gp_common.cuh"
#include "sm100.cuh"

namespace umma_shared {
using namespace sm100;

constexpr int KC = 32;       // x columns per chunk (one 128-byte swizzle span of fp32)
constexpr int NSTAGE = 3;    // fp32 staging ring (TMA destination)
constexpr int STAGE_BYTES = 128 * KC * 4;
constexpr float LOG2E = 1.4426950408889634f;

// Small per-head vectors and operand scales, DEVICE-resident inside the packed blob (written by umma_consts_kernel
// at pack time, read by the row pass in its prologue): packing needs no host round trip, so an optimizer step followed
// by a re-pack stays asynchronous and graph-capturable.
struct alignas(16) UmmaConsts {
  float bv[128], bu[128], ww[KMAX][128], bw[KMAX];
  float inv_s1, inv_sv, inv_su, pad;
};
constexpr size_t UMMA_CONSTS_BYTES = 8192;   // room reserved for UmmaConsts at the end of the blob's umma section
static_assert(sizeof(UmmaConsts) <= UMMA_CONSTS_BYTES, "UmmaConsts outgrew its slot");

struct UmmaParams {
  GpMainParams mp;
  const UmmaConsts* dc;        // device copy of the constants (inside the packed blob)
  CUtensorMap tmap;
  const unsigned char* wimg;   // per-CTA weight images, cta_img_bytes each
  uint32_t cta_img_bytes;
  uint32_t w1_part_bytes;      // bytes of one (hi or lo) W1 half image
};

// Top-n tracking per CTA and bag (training-mode masking).  Rows that may belong to the bag's top n of a branch must stay
// out of the softmax sums (the reference masks them before the softmax), so they are "parked": score / row index /
// h row go to scratch records and rejoin the sums later unless they end up masked.  Nothing is ever subtracted.
//  * Epilogue warps only APPEND: a row whose score beats tau = max(CTA's n-th best, bag-wide n-th best) takes the
//    next record slot of its branch with one shared-memory atomicAdd and is parked.  No locks, no waiting; in the
//    common case no row beats tau and the check costs a few instructions per branch.
//  * Warp 2 is the list manager: it follows the appended records and keeps the CTA's top-n list of every branch
//    (entry i <-> lane i), publishes the CTA's n-th best (tau) and mirrors the list to the workspace (cand_score).
//  * Warp 3 keeps merging the mirrored lists of ALL CTAs that work on the current bag into gtau[k] = the bag-wide
//    n-th best score seen so far by anybody (a lower bound of the final one).
//  * The CTA's first tile of a bag has no threshold yet: there warp k selects the tile's top n of branch k directly
//    (two CTA barriers, once per bag and CTA).
//  * End of the bag: the manager catches up, records that are not in the final list are added back by the CTA, the
//    n survivors go to the reduce kernel.
// Which rows get parked depends on timing, so the summation order of the result does; the top-n set and the mask do not.
constexpr int REC_CAP = 256;   // records per (CTA, bag, branch)
constexpr int ROW_CAP = 512;   // parked h rows per (CTA, bag): one slot per ROW, shared by the branches
constexpr int CAND_KMAX = 6;   // masking on this kernel: K <= 6
constexpr unsigned REC_EMPTY = 0xFFFFFFFFu;     // record score not written yet (a NaN pattern no score can have)
struct CandShared {
  float ls[CAND_KMAX][32];               // list scores, +inf beyond cnt
  int lrec[CAND_KMAX][32];               // ... and their record indices
  unsigned active[CAND_KMAX][REC_CAP / 32];   // bag end: records that are still in the list
  int cnt[8];                            // live list entries
  int rows;                              // h row slots handed out for this bag (may run past row_cap: overflow)
  int app[8];                            // records appended (may run past rec_cap: overflow)
  int seen[8];                           // records the manager has looked at
  float tau[8];                          // n-th best once the list is full, else -inf
  unsigned long long gtau[8];            // bag << 32 | bits of the bag-wide n-th best so far (warp 3, one 8-byte store)
  int cur_bag;                           // bag the epilogue is working on (-1: none yet, -2: kernel is finishing)
  int epoch;                             // bumped when a bag's lists have been booted: the manager may work on them
  int flush_req, flush_ack;              // epoch whose lists the epilogue wants final / the manager has finalised
};

// order-preserving float <-> uint (0 is below every float, +inf is below every NaN)
__device__ __forceinline__ unsigned ord_enc(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_dec(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}

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


}  // namespace umma_shared
