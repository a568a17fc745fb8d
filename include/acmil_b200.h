/*
 * acmil_b200 -- C-ABI of the B200-native gated-attention MIL pooling path.
 *
 * The reference (dazhangyu123/ACMIL) is pure Python: its "operator interface" for this
 * path is the forward() of a handful of nn.Modules.  Each entry point below replaces the
 * eager-op sequence of one of those forwards; the Python modules in acmil_b200/ keep the
 * reference's constructor / forward signatures and parameter names and call these through
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 *   acmil_gp_pack      <- weights of  DimReduction.fc1            (architecture/network.py:37-57)
 *                                     Attention_Gated.{V,U,weights} (architecture/transformer.py:239-267,
 *                                                                  architecture/Attention.py:29-59,
 *                                                                  architecture/attmil.py:45-98, 100-146)
 *   acmil_gp_partial   <- ACMIL_GA.forward  lines 305-324  (architecture/transformer.py): dimreduction,
 *                         gate, stochastic top-k candidate tracking, softmax-over-N partial sums, A_out
 *   acmil_gp_finish    <- ACMIL_GA.forward  lines 311-330: top-k / random mask selection, softmax
 *                         normalisation, A @ x, per-branch Classifier_1fc, bag feature, Slide_classifier;
 *                         ABMIL.forward 277-286; Attention_with_Classifier.forward (Attention.py:67-71);
 *                         attmil AttentionGated / DAttention forward (attmil.py:84-98, 128-146)
 *   acmil_gp_attn_stats<- branch-diversity loss and attention entropy
 *                         (Step3_WSI_classification_ACMIL.py:208-214, 259)
 *   acmil_softmax_rows <- F.softmax(A, dim=1) of Attention_Gated(isNorm=True) (Attention.py:56-57)
 *
 * Conventions
 *   - every function returns 0 on success, a negative ACMIL_E_* code otherwise; the message is
 *     available from acmil_last_error() (thread-local).  No C++ exception crosses the boundary.
 *   - all pointers named d_* are DEVICE pointers into caller-owned memory (torch.empty); the
 *     library never allocates, frees or retains caller memory.  `stream` is a cudaStream_t.
 *   - a "batch" is S bags ("slides") whose rows are concatenated: x is [R_total, d_in] row-major
 *     fp32 and row_offsets[S+1] (HOST array, int64) delimits the bags (like cu_seqlens).
 *   - scores are written branch-major over the concatenation: a_out[k * a_ld + r].
 *   - there is no CPU fallback: with no CUDA device every compute entry point fails with
 *     ACMIL_E_CUDA.
 */
#ifndef ACMIL_B200_H
#define ACMIL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACMIL_ABI_VERSION 1

#if defined(__GNUC__)
#define ACMIL_API __attribute__((visibility("default")))
#else
#define ACMIL_API
#endif

enum {
  ACMIL_OK = 0,
  ACMIL_E_INVALID = -1,      /* bad argument / unsupported shape */
  ACMIL_E_CUDA = -2,         /* CUDA runtime error (message has the cudaError string) */
  ACMIL_E_WORKSPACE = -3,    /* caller buffer too small */
  ACMIL_E_UNSUPPORTED = -4   /* requested implementation cannot run this shape */
};

enum { ACMIL_ACT_TANH = 0, ACMIL_ACT_RELU = 1, ACMIL_ACT_GELU = 2 };
enum { ACMIL_IMPL_AUTO = 0, ACMIL_IMPL_FFMA = 1, ACMIL_IMPL_UMMA = 2 };

#define ACMIL_MAX_BRANCH 8     /* n_token */
#define ACMIL_MAX_MASKED 32    /* n_masked_patch */
#define ACMIL_MAX_SLIDES 128   /* bags per launch */
#define ACMIL_MAX_CLASS 16

/* Static description of one gated-attention pooling head. */
typedef struct acmil_gp_shape {
  int32_t d_in;        /* width of the rows of x (D_feat; 1024 for attmil) */
  int32_t d_inner;     /* L: width the pool runs over (D_inner; = d_in when front == 0) */
  int32_t d_attn;      /* D: hidden width of the gate (128 everywhere in the reference) */
  int32_t n_branch;    /* K = n_token */
  int32_t front;       /* 0: pool x itself (Attention.py); 1: h = act(x W1^T [+ b1]) first */
  int32_t front_bias;  /* DimReduction: 0; attmil feature layer: 1 */
  int32_t front_act;   /* ACMIL_ACT_RELU | ACMIL_ACT_GELU */
  int32_t act_a;       /* activation of the V branch: ACMIL_ACT_* */
  int32_t gated;       /* 1: multiply by sigmoid(U branch); 0: Attention2 / DAttention */
  int32_t gate_bias;   /* V/U Linear have a bias */
  int32_t score_bias;  /* attention_weights Linear has a bias */
  int32_t reserved;
} acmil_gp_shape;

/* Weights in the reference's own layout (nn.Linear: [out, in] row-major fp32, device). */
typedef struct acmil_gp_weights {
  const float* d_w1;   /* [d_inner, d_in]   or NULL when front == 0 */
  const float* d_b1;   /* [d_inner]         or NULL */
  const float* d_wv;   /* [d_attn, d_inner] */
  const float* d_bv;   /* [d_attn]          or NULL */
  const float* d_wu;   /* [d_attn, d_inner] or NULL when gated == 0 */
  const float* d_bu;   /* [d_attn]          or NULL */
  const float* d_ww;   /* [n_branch, d_attn] */
  const float* d_bw;   /* [n_branch]        or NULL */
} acmil_gp_weights;

/* HOST-side copy of the small per-head vectors (biases, score weights) and of the power-of-two operand
 * scales, filled by acmil_gp_pack.  The tcgen05 kernel receives them as kernel parameters (constant bank)
 * so that the gate epilogue needs no loads for them.  The struct is caller-owned; the library keeps no
 * copy.  Pass NULL to acmil_gp_pack / acmil_gp_partial to stay on the FFMA kernel. */
typedef struct acmil_gp_consts {
  float b1[128];
  float bv[128];
  float bu[128];
  float ww[ACMIL_MAX_BRANCH][128];
  float bw[ACMIL_MAX_BRANCH];
  float inv_scale[4];   /* 1 / 2^e applied to the W1, Wv, Wu operand images; [3] unused */
  int32_t valid;        /* set to ACMIL_ABI_VERSION by acmil_gp_pack when the tcgen05 images were built */
  int32_t reserved[3];
} acmil_gp_consts;

/* One batch of bags on one device (one rank's row shard of each bag when sharded). */
typedef struct acmil_gp_batch {
  const float* d_x;            /* [R_total, d_in] fp32 (or fp16, see x_f16) */
  const int64_t* row_offsets;  /* HOST [n_slides + 1], row_offsets[0] == 0 */
  int32_t n_slides;
  int32_t n_masked;            /* n_masked_patch requested (0 = eval / no masking) */
  /* global position of this shard inside each bag (sharded bags; NULL = shard is the bag) */
  const int64_t* shard_row_begin;  /* HOST [n_slides] or NULL */
  float* d_a_out;              /* [n_branch, a_ld] raw scores of the local rows (may be NULL) */
  int64_t a_ld;                /* leading dimension of d_a_out (>= R_total) */
  /* 1: d_x points to IEEE fp16 rows ([R_total, d_in] halves) -- the dtype the reference stores features in
   * (Step2_feature_extract.py:165) and widens before the model (Step3_WSI_classification_ACMIL.py:193).  fp16 -> fp32 is
   * exact, so results equal those of the widened input; the kernels read half the bytes and the tcgen05 kernel skips
   * the x_lo products (x_lo == 0). */
  int32_t x_f16;
  int32_t reserved;
  /* optional: the gate pre-activations of the rows in d_x WITHOUT their biases, [R_total, d_attn * (1 + gated)] fp32 (V units,
   * then U units), computed by the caller (acmil_gemm_nt on the tensor cores); the FFMA kernel then skips its own gate
   * products.  Needs front == 0 (d_x holds h) and fp32 rows; ignored by the tcgen05 kernel (ACMIL_IMPL_FFMA is forced). */
  const float* d_z;
} acmil_gp_batch;

/* Classifier heads applied by acmil_gp_finish (Classifier_1fc: nn.Linear [C, L]). */
typedef struct acmil_gp_heads {
  int32_t n_class;
  int32_t n_branch_heads;      /* K per-branch classifiers (ACMIL_GA) or 0 */
  const float* d_wc;           /* [K, C, L] stacked classifier.{i}.fc.weight, or NULL */
  const float* d_bc;           /* [K, C] */
  int32_t slide_head;          /* 1: Slide_classifier on mean-branch feature (ACMIL_GA) */
  int32_t shared_head;         /* 1: one classifier applied to every branch feature
                                     (ABMIL / Attention_with_Classifier / attmil) via d_ws */
  const float* d_ws;           /* [C, L] */
  const float* d_bs;           /* [C] */
} acmil_gp_heads;

typedef struct acmil_gp_outputs {
  float* d_sub;          /* [S, K, C]  per-branch logits (or NULL) */
  float* d_slide;        /* [S, C]     slide logits      (or NULL) */
  float* d_afeat;        /* [S, K, L]  softmax(A) @ h */
  float* d_bag_feat;     /* [S, L]     mean over branches */
  float* d_lse_m;        /* [S, K]     max score (after masking) */
  float* d_lse_l;        /* [S, K]     sum exp(score - m) */
  int64_t* d_topk_idx;   /* [S, K, n_masked]  torch.topk order (or NULL) */
  int64_t* d_masked_idx; /* [S, K, keep]      indices set to -1e9 (or NULL) */
} acmil_gp_outputs;

ACMIL_API const char* acmil_last_error(void);
ACMIL_API int acmil_abi_version(void);
/* number of CUDA devices visible to the library (0 on a CPU-only box; never fails) */
ACMIL_API int acmil_device_count(void);
/* kernels launched by this library in this process so far (bench.py's gpu_launches) */
ACMIL_API int64_t acmil_launch_count(void);

/* Profiling hook used by bench.py: while enabled, every acmil_gp_partial call brackets its row-pass
 * ("main") kernel with CUDA events on the launch stream.  acmil_prof_collect synchronises them, adds
 * up the elapsed milliseconds and resets the list. */
ACMIL_API int acmil_prof_enable(int on);
ACMIL_API int acmil_prof_collect(double* main_ms_sum, int64_t* n_launches);

/* ---- weight packing ------------------------------------------------------------------ */
ACMIL_API int acmil_gp_packed_bytes(const acmil_gp_shape* shape, size_t* bytes);
/* Packs the weights into the kernels' layouts.  When `consts` is non-NULL and the shape is supported by
 * the tcgen05 kernel it also fills *consts (one small device->host copy + stream synchronisation). */
ACMIL_API int acmil_gp_pack(const acmil_gp_shape* shape, const acmil_gp_weights* w,
                  void* d_packed, size_t packed_bytes, acmil_gp_consts* consts, void* stream);

/* ---- pass over the rows -------------------------------------------------------------- */
/* Sizes (bytes) of the scratch workspace and of the per-rank partial record that
 * acmil_gp_partial fills.  The partial record is what sharded ranks all-gather. */
ACMIL_API int acmil_gp_sizes(const acmil_gp_shape* shape, const acmil_gp_batch* batch, int impl,
                   size_t* workspace_bytes, size_t* partial_bytes);

/* Runs the fused row pass on this rank's rows: front projection, gate, raw scores -> d_a_out,
 * running top-n candidates per branch (when n_masked > 0) kept out of the sums, online-softmax
 * partial sums; then reduces the per-CTA partials into ONE record per bag in d_partial. */
ACMIL_API int acmil_gp_partial(const acmil_gp_shape* shape, const void* d_packed, const acmil_gp_consts* consts,
                     const acmil_gp_batch* batch, int impl, void* d_workspace, size_t workspace_bytes,
                     void* d_partial, size_t partial_bytes, void* stream);
/* Diagnostics (synchronises `stream`): host_flags[s] = 1 for every bag of the LAST acmil_gp_partial call on this
 * workspace that the tcgen05 kernel handed over to the exact FFMA kernel (its bounded scratch for rows that may be
 * in a branch's top n ran out -- score orders such as "ascending along the rows").  The results are the same either
 * way; this only tells how often the slow path ran. */
ACMIL_API int acmil_gp_overflow_flags(const acmil_gp_shape* shape, const acmil_gp_batch* batch, int impl,
                                      const void* d_workspace, int32_t* host_flags, void* stream);
/* 1 when ACMIL_IMPL_AUTO would pick the tcgen05 kernel for this shape (given valid consts). */
ACMIL_API int acmil_gp_umma_supported(const acmil_gp_shape* shape);

/* ---- bag sharding over the GPUs of one box: the exchange fused into the kernels ------------------------------
 * Instead of all-gathering the partial records with NCCL, the reduce kernel of acmil_gp_partial_x stores this rank's
 * records straight into every peer's gather buffer over NVLink (peer-mapped pointers, e.g. from
 * torch.distributed._symmetric_memory) and then raises a per-source flag on every peer; the finish kernel of
 * acmil_gp_finish_x waits for the n_ranks flags in-kernel.  No host synchronisation, no collective launch: the whole
 * step is plain kernel launches and can be captured in a CUDA graph.
 *   gather buffer of one rank: [2 parities][n_ranks][n_slides][record]  (step e uses parity e & 1: double-buffered, a
 *   rank can run at most one step ahead of its peers because its finish waits for their flags)
 *   flags of one rank: n_ranks words, flag[src] = number of steps src has published
 * Both live in caller-owned memory that every peer has mapped; d_epoch / d_ticket are LOCAL device words, all zero
 * before the first step (d_ticket: 4 words).  Every acmil_gp_partial_x must be followed by one acmil_gp_finish_x on the
 * same stream (it advances d_epoch). */
#define ACMIL_MAX_PEERS 16
typedef struct acmil_gp_exchange {
  int32_t n_ranks, rank;
  void* d_gather[ACMIL_MAX_PEERS];      /* gather buffer of rank r as mapped into THIS process ([rank] is the local one) */
  uint32_t* d_flags[ACMIL_MAX_PEERS];   /* flag words of rank r, mapped likewise */
  uint32_t* d_epoch;                    /* local: steps completed on this rank */
  uint32_t* d_ticket;                   /* local scratch: 4 words */
  size_t gather_bytes;                  /* bytes of ONE rank's gather buffer (>= 2 * n_ranks * partial_bytes) */
} acmil_gp_exchange;

/* acmil_gp_partial with the exchange fused in (d_partial is implied: the local gather buffer of `x`). */
ACMIL_API int acmil_gp_partial_x(const acmil_gp_shape* shape, const void* d_packed, const acmil_gp_consts* consts,
                                 const acmil_gp_batch* batch, int impl, void* d_workspace, size_t workspace_bytes,
                                 const acmil_gp_exchange* x, void* stream);
/* acmil_gp_finish / acmil_gp_finish_rand over the records pushed by every rank's acmil_gp_partial_x (exactly one of
 * d_rsel / d_rand when masking is on). */
ACMIL_API int acmil_gp_finish_x(const acmil_gp_shape* shape, const acmil_gp_batch* batch, const acmil_gp_exchange* x,
                                const int32_t* keep, const int64_t* d_rsel, const float* d_rand, int32_t rand_ld,
                                int32_t keep_ld, const acmil_gp_heads* heads, const acmil_gp_outputs* out, void* stream);

/* Merges n_ranks partial records (d_partials = n_ranks records back to back, as produced by an
 * all-gather; n_ranks == 1 for a single GPU), picks the global top-n per branch, masks
 * d_rsel-selected ones (rsel[S, K, keep] = argsort(rand)[:, :keep] drawn by the caller with
 * torch), normalises, applies the heads, and writes -1e9 into d_a_out at the masked positions
 * that fall inside this rank's shard.
 *   keep: number of masked patches per branch = int(min(n_masked, N) * mask_drop), per bag
 *         (HOST array [S]); rsel row stride is keep_ld. */
ACMIL_API int acmil_gp_finish(const acmil_gp_shape* shape, const acmil_gp_batch* batch,
                    const void* d_partials, size_t partial_bytes, int n_ranks,
                    const int32_t* keep, const int64_t* d_rsel, int32_t keep_ld,
                    const acmil_gp_heads* heads, const acmil_gp_outputs* out, void* stream);

/* Same, but the mask draw is handed over raw: d_rand[S, K, rand_ld] = the uniform numbers of
 * torch.rand(K, n) (transformer.py:316) and the kernel itself takes rsel = argsort(rand[:, :n])[:, :keep]
 * (n = min(n_masked, rows of the bag); ties: lower index first), saving the caller's sort launches.
 * d_masked_idx rows keep the stride keep_ld. */
ACMIL_API int acmil_gp_finish_rand(const acmil_gp_shape* shape, const acmil_gp_batch* batch, const void* d_partials,
                                   size_t partial_bytes, int n_ranks, const int32_t* keep, const float* d_rand,
                                   int32_t rand_ld, int32_t keep_ld, const acmil_gp_heads* heads,
                                   const acmil_gp_outputs* out, void* stream);

/* ---- small ops on the [K, N] score matrix --------------------------------------------- */
/* Per bag: gram[K,K] = sum_n P_i P_j, ent[K] = sum_n P log P with P = softmax(A_out) given the
 * (m, l) of acmil_gp_finish; div = mean pairwise cosine; all fp32. */
ACMIL_API int acmil_gp_attn_stats(const float* d_a, int64_t a_ld, int32_t n_branch, const int64_t* row_offsets,
                        int32_t n_slides, const float* d_lse_m, const float* d_lse_l,
                        float* d_gram, float* d_ent, float* d_div, void* stream);
/* out[k, n] = softmax over n of a[k, n] (one bag). */
ACMIL_API int acmil_softmax_rows(const float* d_a, int64_t a_ld, int32_t n_rows, int64_t n, float* d_out,
                       int64_t out_ld, void* stream);

/* ---- backward of the pool (recompute-based; Step3_WSI_classification_ACMIL.py:216-221 loss.backward()) ------------
 * The products over the hidden widths and over the N rows run on acmil_gemm_nt (acmil_transmil.h); these entry points
 * are the row-local kernels between them (csrc/gp_bwd.cu has the formulas).  One bag per call.
 *   z        [n, zc]  gate pre-activations, bias included: V units [0, d_attn), U units [d_attn, 2 d_attn) when gated
 *   scores   [K, n]   the forward's raw scores (-1e9 at the masked positions), leading dimension a_ld
 *   lse_m/l  [K]      softmax statistics of the forward (acmil_gp_outputs.lse_m / lse_l)
 *   g_afeat  [K, d_inner] / g_bag [d_inner] / g_scores [K, n] (ld gs_ld): upstream gradients, each may be NULL
 * outputs: dz [n, zc], dzt [zc, ldt] (transposed copy, ldt >= n), dhp [n, d_inner] (pool-path gradient of h),
 *   small [KMAX * d_attn + KMAX + 2 * d_attn] = dWw rows (row k at k * d_attn) | dbw | column sums of dz;
 *   partials: workspace of acmil_gp_bwd_workspace_floats() floats. */
typedef struct acmil_gp_bwd_gate_args {
  const float* d_h;
  const float* d_z;
  const float* d_scores;
  const float* d_lse_m;
  const float* d_lse_l;
  const float* d_afeat;
  const float* d_g_afeat;
  const float* d_g_bag;
  const float* d_g_scores;
  const float* d_ww;
  int64_t n, a_ld, gs_ld, ldt;
  int32_t d_inner, d_attn, n_branch, act_a, gated, reserved;
  float* d_dz;
  float* d_dzt;
  float* d_dhp;
  float* d_partials;
  float* d_small;
} acmil_gp_bwd_gate_args;
ACMIL_API int acmil_gp_bwd_workspace_floats(int32_t d_inner, int64_t* gate_floats, int64_t* relu_floats);
ACMIL_API int acmil_gp_bwd_gate(const acmil_gp_bwd_gate_args* args, void* stream);
/* dz1 = dh * [h > 0] (row-major copy optional: d_dz1 may be NULL), dz1t [d_inner, ldt] its transpose, db1 [d_inner] the
 * column sums (relu front layer, network.py:49-57). */
ACMIL_API int acmil_gp_bwd_relu_mask(const float* d_dh, const float* d_h, int64_t n, int32_t d_inner, float* d_dz1, float* d_dz1t,
                                     int64_t ldt, float* d_partials, float* d_db1, void* stream);
/* out[c * ldo + r] = x[r * ldx + c] */
ACMIL_API int acmil_transpose_f32(const float* d_x, int64_t ldx, int64_t rows, int32_t cols, float* d_out, int64_t ldo, void* stream);

/* Branch-diversity loss of the training loop (Step3_WSI_classification_ACMIL.py:208-214) on the [K, n] score matrix of ONE
 * bag: div = mean over branch pairs of cos(softmax(A_i), softmax(A_j)).  fwd writes ml [2 * ACMIL_MAX_BRANCH + 1] (the softmax
 * statistics m, then l, then one word of scratch), gram [ACMIL_MAX_BRANCH^2] (P P^T) and the loss; bwd writes ds = g_out * d div / d A ([K, n], ld ds_ld;
 * g_out is a DEVICE scalar: the upstream gradient).  Masked positions (-1e9) get a zero gradient like masked_fill's. */
ACMIL_API int acmil_div_loss_fwd(const float* d_a, int64_t a_ld, int32_t n_branch, int64_t n, float* d_ml, float* d_gram, float* d_div,
                                 void* stream);
ACMIL_API int acmil_div_loss_bwd(const float* d_a, int64_t a_ld, int32_t n_branch, int64_t n, const float* d_ml, const float* d_gram,
                                 const float* d_g_out, float* d_ds, int64_t ds_ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACMIL_B200_H */
