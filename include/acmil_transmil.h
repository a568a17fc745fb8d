/*
 * acmil_transmil -- C-ABI of the B200-native TransMIL / Nystrom-attention path (same library,
 * libacmil_b200.so; same conventions as acmil_b200.h: 0 / negative ACMIL_E_* return codes with
 * acmil_last_error(), d_* = caller-owned device memory, `stream` = cudaStream_t, no CPU path).
 *
 *   acmil_gemm_nt            <- every nn.Linear / einsum / matmul of the path
 *                               (nystrom_attention.py:83,119-121,135,147; transMIL.py:61,89): fp32
 *                               operands, error-compensated 3xTF32 on tcgen05 (fp32-faithful)
 *   acmil_layernorm_rows     <- nn.LayerNorm(dim) of TransLayer.norm / TransMIL.norm (transMIL.py:11,27,58,86)
 *   acmil_nystrom_workspace_bytes, acmil_nystrom_attn_fwd
 *                            <- NystromAttention.forward, mask=None, return_attn=False
 *                               (architecture/nystrom_attention.py:67-140 == pip nystrom-attention 0.0.12),
 *                               optionally fused with the pre-LayerNorm and the residual add of
 *                               TransLayer.forward (transMIL.py:25-28)
 *   acmil_ppeg_fwd           <- PPEG.forward (transMIL.py:38-45)
 *   acmil_vit_workspace_bytes, acmil_vit_fwd
 *                            <- the ViT-S/16 patch encoder of Step2_feature_extract.py (models.py:138-149,
 *                               166-179; timm 0.9.2 VisionTransformer.forward)
 */
#ifndef ACMIL_TRANSMIL_H
#define ACMIL_TRANSMIL_H

#include "acmil_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* C[b] = alpha * A[b] (m x k, row-major, ld = lda) * B[b]^T (n x k, ld = ldb)
 *        (+ diag on row == col) (+ bias[col]) (+ beta * addend[b][row * ld_addend + col]) (relu)
 * Element (row, col) of batch b is stored at
 *   d_c  + b * c_batch_stride + (col / col_block_width) * col_block_stride + row * ldc + col % col_block_width
 *   (col_block_width == 0: plain row-major), and, if d_ct != NULL, also transposed at
 *   d_ct + b * ct_batch_stride + col * ldct + row.
 * A batch stride of 0 shares the operand between batch entries.  All strides in floats; operand
 * pointers 16-byte aligned, lda / ldb multiples of 4.  precise = 1: 3xTF32 split (fp32-faithful,
 * ~2^-21), precise = 0: plain TF32 (~2^-10).  precise = 2 with b_split != NULL: B is read from a
 * pre-split fp16 hi / lo image of a weight matrix (acmil_gemm_split_b; rows b_split_row0 .. + n of an
 * image of b_split_rows rows whose K equals k; `b` is ignored, B cannot be batched or k-split) and A
 * is split to fp16 hi / lo inside the kernel: 3 kind::f16 MMAs per product, half the tensor-pipe time
 * and 56 % of the shared-memory traffic of the TF32 split, fp32-level accuracy (~2^-22) for
 * |a| < 65504 (larger values saturate, and the saturating conversion does not carry Inf / NaN of A
 * through as IEEE arithmetic would; |a| below 2^-14 keep an absolute 2^-25).  precise = 2 without
 * an image behaves like precise = 1.  k_split > 1 partitions K over that many CTAs per tile
 * (d_split_ws: k_split * batch * m * n floats; epilogue terms are applied after the reduction).
 * Two-level batches: with batch_inner > 0, entry z uses offsets (z % batch_inner) * X_batch_stride +
 * (z / batch_inner) * X_batch_stride2 for every operand X (e.g. z = head * images + image).
 * act: 0 none, 1 relu, 2 erf-based GELU (nn.GELU() default; erf by Abramowitz-Stegun 7.1.26, |error| <= 2e-7) -- applied last. */
typedef struct acmil_gemm_desc {
  const float* a;
  const float* b;
  float* c;
  float* ct;
  const float* bias;
  const float* addend;
  float* split_ws;
  int32_t m, n, k, batch;
  int64_t lda, ldb, ldc, ldct, ld_addend;
  int64_t a_batch_stride, b_batch_stride, c_batch_stride, ct_batch_stride, addend_batch_stride;
  int32_t col_block_width;
  int32_t k_split;
  int64_t col_block_stride;
  float alpha, beta, diag;
  int32_t act;
  int32_t precise;
  int32_t batch_inner;
  int32_t bias_per_row;           /* 1: bias is indexed by the output row instead of the column */
  int32_t reserved;
  int64_t a_batch_stride2, b_batch_stride2, c_batch_stride2, ct_batch_stride2, addend_batch_stride2;
  const void* b_split;            /* precise = 2: image made by acmil_gemm_split_b, or NULL */
  int32_t b_split_rows;           /* rows of the whole image */
  int32_t b_split_row0;           /* first image row of this product's B */
  /* Chunked softmax across two products (scores -> softmax -> attention x values without the softmax pass):
   * softmax_stats_out != NULL: every 32-column chunk of C stores exp(c - chunk max) and its (max, sum) pair goes to
   *   stats[((batch * m + row) * ceil(n / 32) + chunk) * 2] (floats);
   * softmax_stats_in != NULL (precise = 1 products): A holds such exponentials along K, ceil(k / 32) pairs per row in the
   *   same layout; each chunk of a row is rescaled by exp(max_c - row max) / row sum on its way to the tensor core,
   *   i.e. the product is softmax(scores) B^T. */
  float* softmax_stats_out;
  const float* softmax_stats_in;
} acmil_gemm_desc;

ACMIL_API int acmil_gemm_nt(const acmil_gemm_desc* desc, void* stream);

/* Pre-split image of a weight matrix b [rows, k] (row-major, leading dimension ldb) for precise = 2:
 * a 256-byte header (the power-of-two scale, chosen on the device from max |b|: no host sync), then the
 * fp16 hi and lo sections, rows padded to a multiple of 8 halves.  The image is a function of the
 * weight values only: rebuild it when they change.  256-byte aligned, acmil_gemm_split_bytes long. */
ACMIL_API int acmil_gemm_split_bytes(int32_t rows, int32_t k, size_t* bytes);
ACMIL_API int acmil_gemm_split_b(const float* d_b, int32_t rows, int32_t k, int64_t ldb, void* d_image, size_t image_bytes,
                                 void* stream);

/* out[r, :] = (x[r, :] - mean) / sqrt(var + eps) * w + b  (biased variance, like nn.LayerNorm). */
ACMIL_API int acmil_layernorm_rows(const float* d_x, int64_t ldx, int64_t rows, int32_t dim, const float* d_w,
                                   const float* d_b, float eps, float* d_out, int64_t ldo, void* stream);

typedef struct acmil_nystrom_shape {
  int32_t batch, n, dim;          /* x: [batch, n, dim] */
  int32_t heads, dim_head;        /* inner = heads * dim_head */
  int32_t num_landmarks;          /* m */
  int32_t pinv_iterations;
  int32_t residual;               /* depth-wise conv residual of the values */
  int32_t conv_kernel;            /* 33 */
  int32_t n_out;                  /* > 0: only the first n_out output rows are computed (TransMIL's second
                                     layer is read at the class token only, transMIL.py:86); 0 = all n */
  int32_t padded_out;             /* 1: d_out is [batch, n_pad, dim] = to_out(...) BEFORE the [:, -n:] slice,
                                     no residual (the caller applies train-mode dropout, nystrom_attention.py:56-59,138) */
  int32_t precise;                /* GEMM mode, see acmil_gemm_desc */
  int32_t reserved[4];
} acmil_nystrom_shape;

typedef struct acmil_nystrom_weights {
  const float* d_ln_w;            /* optional pre-LayerNorm (TransLayer.norm); NULL = x is used as is */
  const float* d_ln_b;
  float ln_eps;
  int32_t reserved;
  const float* d_wqkv;            /* to_qkv.weight   [3 * inner, dim] */
  const float* d_wout;            /* to_out.0.weight [dim, inner] */
  const float* d_bout;            /* to_out.0.bias   [dim] */
  const float* d_wconv;           /* res_conv.weight [heads, 1, conv_kernel, 1] */
  const void* d_split_qkv;        /* optional acmil_gemm_split_b images of to_qkv.weight (all 3 * inner rows) and */
  const void* d_split_out;        /* to_out.0.weight: used by the q, k and to_out products when precise = 2 */
} acmil_nystrom_weights;

ACMIL_API int acmil_nystrom_workspace_bytes(const acmil_nystrom_shape* shape, size_t* bytes);

/* d_out[b, i, :] = NystromAttention(LN(x))[b, i, :] + (d_residual ? d_residual[b, i, :] : 0). */
ACMIL_API int acmil_nystrom_attn_fwd(const acmil_nystrom_shape* shape, const acmil_nystrom_weights* w, const float* d_x,
                                     const float* d_residual, float* d_out, void* d_workspace, size_t workspace_bytes,
                                     void* stream);

/* x, out: [batch, 1 + gh * gw, c]; row 0 (class token) is copied; the grid rows get
 * feat + conv7(feat) + conv5(feat) + conv3(feat), depth-wise with zero padding (transMIL.py:31-45).
 * Weights in nn.Conv2d layout [c, 1, k, k], biases [c]. */
ACMIL_API int acmil_ppeg_fwd(const float* d_x, int32_t batch, int32_t gh, int32_t gw, int32_t c, const float* d_w7,
                             const float* d_b7, const float* d_w5, const float* d_b5, const float* d_w3, const float* d_b3,
                             float* d_out, void* stream);

/* ---- sequence-parallel NystromAttention: one rank's share of a padded sequence cut at landmark-group boundaries
 * (csrc/tm_ops.cu has the phase list; acmil_b200/transmil_sharded.py drives it and does the exchanges).  batch = 1. */
typedef struct acmil_nystrom_shard {
  int32_t n_loc;            /* rows of the padded sequence on this rank = m_loc * group_len (multiple of 4) */
  int32_t lead_zero;        /* leading all-zero rows (the front padding of nystrom_attention.py:72-80: rank 0 only) */
  int32_t dim, heads, dim_head;
  int32_t num_landmarks;    /* m of the whole sequence */
  int32_t m_loc;            /* landmark groups on this rank */
  int32_t group_len;        /* l = n_pad / m */
  int32_t pinv_iterations, residual, conv_kernel, precise;
  int32_t n_out;            /* phase D produces the first n_out real rows (0 = all n_loc - lead_zero) */
  int32_t head_first, head_count;   /* heads whose pseudo-inverse this rank iterates in phase B */
  int32_t halo;             /* conv_kernel / 2 columns on both sides of d_vt_ext */
} acmil_nystrom_shard;

typedef struct acmil_nystrom_shard_bufs {
  const float* d_x;         /* [n_loc - lead_zero, dim] local real rows */
  const float* d_residual;  /* [rows produced, dim] or NULL */
  float* d_out;             /* [rows produced, dim] */
  float* d_ql_loc;          /* phase A out: [heads, m_loc, dim_head] local landmark means of q (scaled) */
  float* d_kl_loc;          /* ... and of k */
  const float* d_ql;        /* [heads, m, dim_head] landmarks of the whole sequence (phases B, C) */
  const float* d_kl;        /* (phases B, D) */
  float* d_z;               /* [heads, m, m] pseudo-inverse: phase B writes its heads, phase D reads all */
  float* d_kv_part;         /* phase C out: [heads, dim_head, m] sum_n exp(s - max_loc) v */
  float* d_st_m;            /* phase C out: [heads * m] local row maxima */
  float* d_st_l;            /* phase C out: [heads * m] local row sums of exp */
  const float* d_kv;        /* phase D in: [heads, dim_head, m] = (attn3 v)^T of the whole sequence (acmil_lse_merge) */
  float* d_vt_ext;          /* [heads * dim_head, n_loc + 2 * halo]: phase A writes the middle, the caller the halos */
  void* d_workspace;
  size_t workspace_bytes;
} acmil_nystrom_shard_bufs;

ACMIL_API int acmil_nystrom_shard_workspace_bytes(const acmil_nystrom_shard* shard, size_t* bytes);
/* phase 0..3 = A..D */
ACMIL_API int acmil_nystrom_shard_phase(const acmil_nystrom_shard* shard, const acmil_nystrom_weights* w,
                                        const acmil_nystrom_shard_bufs* bufs, int32_t phase, void* stream);
/* kv = sum_p e^(m_p - M) parts_p / sum_p e^(m_p - M) l_p with parts [n_ranks][heads, dim_head, m], st_m / st_l [n_ranks][heads * m] */
ACMIL_API int acmil_lse_merge(const float* d_parts, const float* d_st_m, const float* d_st_l, int32_t n_ranks, int32_t heads,
                              int32_t dim_head, int32_t num_landmarks, float* d_kv, void* stream);
/* acmil_ppeg_fwd restricted to grid rows [y_first, y_first + ny) of one sequence; the class-token row is left alone. */
ACMIL_API int acmil_ppeg_fwd_rows(const float* d_x, int32_t gh, int32_t gw, int32_t c, const float* d_w7, const float* d_b7,
                                  const float* d_w5, const float* d_b5, const float* d_w3, const float* d_b3, float* d_out,
                                  int32_t y_first, int32_t ny, void* stream);

/* in-place softmax over `rows` rows of length len <= 1024 stored with leading dimension ld. */
ACMIL_API int acmil_softmax_rows_inplace(float* d_a, int64_t ld, int64_t rows, int32_t len, void* stream);

/* ---- ViT patch encoder (models.py:138-149 vit_small -> timm 0.9.2 VisionTransformer(img_size=224,
 * patch_size=16, embed_dim=384, num_heads=6, num_classes=0): conv patch embedding, class token + learned
 * position embedding, depth x [LN, MHSA with qkv bias, LN, MLP with exact GELU], final LN, class-token
 * feature; CustomModel.head (models.py:166-179) optionally applied to the feature). */
typedef struct acmil_vit_shape {
  int32_t batch, img, patch, in_ch;   /* images: [batch, in_ch, img, img] fp32 */
  int32_t dim, depth, heads, mlp_dim;
  int32_t n_class;                    /* 0: no head */
  int32_t precise;
  float ln_eps;                       /* 1e-6 in timm's ViT */
  int32_t reserved[5];
} acmil_vit_shape;

typedef struct acmil_vit_block_weights {
  const float *d_ln1_w, *d_ln1_b, *d_qkv_w, *d_qkv_b, *d_proj_w, *d_proj_b;
  const float *d_ln2_w, *d_ln2_b, *d_fc1_w, *d_fc1_b, *d_fc2_w, *d_fc2_b;
  /* optional acmil_gemm_split_b images of qkv.weight (all 3 * dim rows), proj / fc1 / fc2 .weight (precise = 2) */
  const void *d_split_qkv, *d_split_proj, *d_split_fc1, *d_split_fc2;
} acmil_vit_block_weights;

typedef struct acmil_vit_weights {
  const float *d_cls_token, *d_pos_embed;     /* [1,1,dim], [1, 1 + (img/patch)^2, dim] */
  const float *d_patch_w, *d_patch_b;         /* patch_embed.proj: [dim, in_ch, patch, patch], [dim] */
  const float *d_norm_w, *d_norm_b;           /* final norm */
  const float *d_head_w, *d_head_b;           /* [n_class, dim], [n_class] or NULL */
  const acmil_vit_block_weights* blocks;      /* HOST array of `depth` entries */
  const void* d_split_patch;                  /* optional split image of patch_embed.proj.weight as [dim, in_ch * patch^2] */
} acmil_vit_weights;

ACMIL_API int acmil_vit_workspace_bytes(const acmil_vit_shape* shape, size_t* bytes);
/* d_features: [batch, dim]; d_logits: [batch, n_class] or NULL. */
ACMIL_API int acmil_vit_fwd(const acmil_vit_shape* shape, const acmil_vit_weights* w, const float* d_images, float* d_features,
                            float* d_logits, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- feature-extraction loop (Step2_feature_extract.py:35-71, 164-167; datasets/dataset_h5.py:20-37) ----
 * d_img: uint8 RGB patches [batch, in_h, in_w, 3] (PIL layout) -> d_out fp32 [batch, 3, out, out] =
 * Normalize(mean, std)(ToTensor(Resize(out)(patch))); the resize is Pillow's antialiased 8-bit BILINEAR resample,
 * byte-exact.  mean3 / std3 are HOST arrays of 3 floats. */
ACMIL_API int acmil_preprocess_workspace_bytes(int32_t in_h, int32_t in_w, int32_t out_size, size_t* bytes);
ACMIL_API int acmil_preprocess_u8(const uint8_t* d_img, int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size,
                                  const float* mean3, const float* std3, float* d_out, void* d_workspace,
                                  size_t workspace_bytes, void* stream);
/* d_y[i] = (fp16) d_x[i], round to nearest even: the `.astype(np.float16)` of the feature store. */
ACMIL_API int acmil_f32_to_f16(const float* d_x, void* d_y, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACMIL_TRANSMIL_H */
