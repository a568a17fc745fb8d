// gp_bwd.cu -- the memory-bound kernels of the gated-attention pool's BACKWARD pass (recompute-based; SURVEY section 8
// row f1: the training step of Step3_WSI_classification_ACMIL.py:175-235 at kernel speed).
//
// Forward (architecture/transformer.py:305-330):  h = relu(x W1^T [+ b1]);  a = act(h Wv^T + bv);  b = sigmoid(h Wu^T + bu);
//   s = (a * b) Ww^T + bw;  s' = s with -1e9 at the masked positions;  P = softmax_N(s');  afeat = P h;  bag = mean_k afeat.
// Backward, given G = d/d afeat (+ d/d bag / K) and the gradient of the raw scores (diversity loss):
//   ds_kn  = P_kn (G_k . h_n - G_k . afeat_k) + gs_kn            (0 at masked positions: masked_fill cuts the graph there)
//   dg_n   = sum_k ds_kn Ww_k;   dzv = dg b act'(zv);   dzu = dg a b (1 - b)
//   dh_n   = sum_k P_kn G_k  +  dzv Wv + dzu Wu;   dz1 = dh * [h > 0]
//   dWw = ds g,  dbw = sum_n ds,  dWv = dzv^T h,  dWu = dzu^T h,  dbv / dbu = column sums,  dW1 = dz1^T x,  db1 = column sums.
// Every product over the hidden widths or over the N rows runs on the tcgen05 GEMM engine (tm_gemm.cu; the host side is
// acmil_b200/gp_backward.py); the kernels here do the row-local part between those GEMMs and emit the operands the
// GEMMs over N need in K-major (transposed) form, so nothing is transposed by a separate pass:
//   acmil_gp_bwd_gate       h, Z = [zv | zu], s', (m, l), afeat, G  ->  dZ = [dzv | dzu], dZ^T, pool-path dh, dWw, dbw, dbz
//   acmil_gp_bwd_relu_mask  dh, h -> dz1 (optional), dz1^T, db1
//   acmil_transpose_f32     x -> x^T (the K-major B operand of dW1 = dz1^T x)
#include <algorithm>

#include "gp_common.cuh"

namespace {

constexpr int BT = 256;          // threads per CTA
constexpr int TR = 32;           // rows per tile
constexpr int LMAX = 512;        // d_inner bound (multiple of 32)
constexpr int DA = GP_DATTN;     // gate width

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct GateArgs {
  const float *h, *z, *scores, *lse_m, *lse_l, *afeat, *g_afeat, *g_bag, *g_scores, *ww;
  long long n, a_ld, gs_ld, ldt;
  int L, K, zc, act_a, gated, ntiles;
  float *dz, *dzt, *dhp, *partials;
};
constexpr int PW = KMAX * DA + KMAX + 2 * DA;     // floats of one CTA partial: dWw rows | dbw | column sums of dZ

// exact-erf GELU derivative: Phi(z) + z phi(z)
__device__ __forceinline__ float gelu_grad(float z) {
  return 0.5f * (1.f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
}

// NF = d_inner / 32 features per lane, KB = branch bound (K <= KB).  One warp per row, 32-row tiles; the dZ tile is
// transposed through shared memory.  Shared memory is sized by the actual shape so that 4 CTAs (32 warps) fit an SM:
// the per-row chain (loads -> K warp reductions -> exp -> gate) is latency-bound and needs the warps.
template <int NF, int KB>
__global__ void __launch_bounds__(BT, 2) gp_bwd_gate_kernel(const GateArgs a) {
  extern __shared__ float sm[];
  constexpr int L = NF * 32;
  const int K = a.K, zc = a.zc;
  float* sG = sm;                          // [KB][L]
  float* sW = sG + KB * L;                 // [KB][DA]
  float* sC = sW + KB * DA;                // c[KB], m[KB], inv_l[KB]
  float* sAcc = sC + 3 * KB;               // [PW] CTA accumulator of the small reductions
  float* tile = sAcc + PW;                 // [TR][zc + 1]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < KB * L; i += BT) {
    const int k = i / L, j = i - k * L;
    sG[i] = k < K ? (a.g_afeat ? a.g_afeat[i] : 0.f) + (a.g_bag ? a.g_bag[j] / (float)K : 0.f) : 0.f;
  }
  for (int i = tid; i < KB * DA; i += BT) sW[i] = i < K * DA ? a.ww[i] : 0.f;
  for (int i = tid; i < PW; i += BT) sAcc[i] = 0.f;
  __syncthreads();
  if (warp < K) {                          // c_k = G_k . afeat_k
    float c = 0.f;
    for (int j = lane; j < L; j += 32) c = fmaf(sG[warp * L + j], a.afeat[warp * L + j], c);
    c = wsum(c);
    if (lane == 0) {
      sC[warp] = c;
      sC[KB + warp] = a.lse_m[warp];
      sC[2 * KB + warp] = 1.f / a.lse_l[warp];
    }
  }
  __syncthreads();
  float acc_ww[KB][4], acc_bz[8], acc_bw[KB];
#pragma unroll
  for (int k = 0; k < KB; ++k) {
    acc_bw[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc_ww[k][i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc_bz[i] = 0.f;

  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    const long long n0 = (long long)t * TR;
#pragma unroll 2
    for (int rr = 0; rr < TR / 8; ++rr) {      // two rows in flight per warp: their loads overlap
      const int r = warp * (TR / 8) + rr;
      const long long n = n0 + r;
      if (n < a.n) {
        float hv[NF];
        const float* hr = a.h + n * L;
#pragma unroll
        for (int i = 0; i < NF; ++i) hv[i] = hr[lane + 32 * i];
        const float* zr = a.z + n * zc;
        float zv[4], zu[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          zv[i] = zr[lane + 32 * i];
          zu[i] = a.gated ? zr[DA + lane + 32 * i] : 0.f;
        }
        float d[KB], sc[KB], gsc[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          d[k] = 0.f;
          sc[k] = k < K ? a.scores[(long long)k * a.a_ld + n] : -1e9f;
          gsc[k] = (k < K && a.g_scores) ? a.g_scores[(long long)k * a.gs_ld + n] : 0.f;
#pragma unroll
          for (int i = 0; i < NF; ++i) d[k] = fmaf(sG[k * L + lane + 32 * i], hv[i], d[k]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int k = 0; k < KB; ++k) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
        float ds[KB], pk[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const bool masked = sc[k] == -1e9f;
          pk[k] = masked ? 0.f : __expf(sc[k] - sC[KB + k]) * sC[2 * KB + k];
          ds[k] = pk[k] * (d[k] - sC[k]) + (masked ? 0.f : gsc[k]);
          acc_bw[k] += ds[k];
        }
        // pool-path gradient of h
        float* dr = a.dhp + n * L;
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          float v = 0.f;
#pragma unroll
          for (int k = 0; k < KB; ++k) v = fmaf(pk[k], sG[k * L + lane + 32 * i], v);
          dr[lane + 32 * i] = v;
        }
        // gate units
        float* dzr = a.dz + n * zc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = lane + 32 * i;
          float av, dav;
          if (a.act_a == ACMIL_ACT_TANH) { av = tanhf(zv[i]); dav = 1.f - av * av; }
          else if (a.act_a == ACMIL_ACT_RELU) { av = fmaxf(zv[i], 0.f); dav = zv[i] > 0.f ? 1.f : 0.f; }
          else { av = 0.5f * zv[i] * (1.f + erff(zv[i] * 0.70710678118654752f)); dav = gelu_grad(zv[i]); }
          const float bv = a.gated ? 1.f / (1.f + __expf(-zu[i])) : 1.f;
          const float g = av * bv;
          float dg = 0.f;
#pragma unroll
          for (int k = 0; k < KB; ++k) {
            dg = fmaf(ds[k], sW[k * DA + u], dg);
            acc_ww[k][i] = fmaf(ds[k], g, acc_ww[k][i]);
          }
          const float dzv = dg * bv * dav;
          dzr[u] = dzv;
          tile[r * (zc + 1) + u] = dzv;
          acc_bz[i] += dzv;
          if (a.gated) {
            const float dzu = dg * av * bv * (1.f - bv);
            dzr[DA + u] = dzu;
            tile[r * (zc + 1) + DA + u] = dzu;
            acc_bz[4 + i] += dzu;
          }
        }
      } else {
        for (int u = lane; u < zc; u += 32) tile[r * (zc + 1) + u] = 0.f;
      }
    }
    __syncthreads();
    // dZ^T: warp w writes units [32 w, 32 w + 32), lane = row of the tile (coalesced 128-byte rows)
    if (warp * 32 < zc && n0 + lane < a.n) {
#pragma unroll 8
      for (int uu = 0; uu < 32; ++uu) {
        const int u = warp * 32 + uu;
        a.dzt[(long long)u * a.ldt + n0 + lane] = tile[lane * (zc + 1) + u];
      }
    }
    __syncthreads();
  }
  // CTA partial of the small reductions: shared-memory accumulation over the 8 warps, once per CTA
#pragma unroll
  for (int k = 0; k < KB; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(&sAcc[k * DA + lane + 32 * i], acc_ww[k][i]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < KB; ++k) atomicAdd(&sAcc[KMAX * DA + k], acc_bw[k]);
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(&sAcc[KMAX * DA + KMAX + (i >> 2) * DA + lane + 32 * (i & 3)], acc_bz[i]);
  __syncthreads();
  float* out = a.partials + (size_t)blockIdx.x * PW;
  for (int e = tid; e < PW; e += BT) out[e] = sAcc[e];
}

// out[e] = sum over the CTA partials: one warp per output element
__global__ void __launch_bounds__(256) gp_bwd_sum_partials_kernel(const float* __restrict__ partials, int nparts, int pw,
                                                                  float* __restrict__ out) {
  const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= pw) return;
  float s = 0.f;
  for (int b = lane; b < nparts; b += 32) s += partials[(size_t)b * pw + e];
  s = wsum(s);
  if (lane == 0) out[e] = s;
}

// dz1 = dh * [h > 0] (optional row-major copy), dz1^T, column sums (db1)
template <int NF>
__global__ void __launch_bounds__(BT, NF <= 8 ? 3 : 2) gp_bwd_relu_mask_kernel(const float* __restrict__ dh, const float* __restrict__ h,
                                                                              long long n, int ntiles, float* __restrict__ dz1,
                                                                              float* __restrict__ dz1t, long long ldt,
                                                                              float* __restrict__ partials) {
  extern __shared__ float sm[];
  constexpr int L = NF * 32;
  float* tile = sm;                         // [TR][L + 1]
  float* sAcc = tile + TR * (L + 1);        // [L]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < L; i += BT) sAcc[i] = 0.f;
  float acc[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) acc[i] = 0.f;
  __syncthreads();
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long n0 = (long long)t * TR;
#pragma unroll
    for (int rr = 0; rr < TR / 8; ++rr) {
      const int r = warp * (TR / 8) + rr;
      const long long row = n0 + r;
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        float v = 0.f;
        if (row < n) {
          const float hv = h[row * L + lane + 32 * i];
          v = hv > 0.f ? dh[row * L + lane + 32 * i] : 0.f;
          if (dz1) dz1[row * L + lane + 32 * i] = v;
        }
        tile[r * (L + 1) + lane + 32 * i] = v;
        acc[i] += v;
      }
    }
    __syncthreads();
    if (n0 + lane < n)
#pragma unroll 4
      for (int f = warp; f < L; f += 8) dz1t[(long long)f * ldt + n0 + lane] = tile[lane * (L + 1) + f];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < NF; ++i) atomicAdd(&sAcc[lane + 32 * i], acc[i]);
  __syncthreads();
  for (int e = tid; e < L; e += BT) partials[(size_t)blockIdx.x * L + e] = sAcc[e];
}

__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols,
                                                            float* __restrict__ out, long long ldo) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long r = r0 + ty + 8 * j;
    tile[ty + 8 * j][tx] = (r < rows && c0 + tx < cols) ? x[r * ldx + c0 + tx] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + ty + 8 * j;
    if (c < cols && r0 + tx < rows) out[(long long)c * ldo + r0 + tx] = tile[tx][ty + 8 * j];
  }
}


// ------------------------------------------------------------------------------------------------
// Branch-diversity loss of the training script (Step3_WSI_classification_ACMIL.py:208-214):
//   P = softmax(A, dim=-1);  div = sum_{i<j} cos(P_i, P_j) / (K (K - 1) / 2),  cos = P_i.P_j / (max(|P_i|, eps) max(|P_j|, eps))
// forward: one CTA per call (A is K x N, ~1 MB): pass 1 the softmax statistics (m, l) of every branch, pass 2 the Gram
// matrix G = P P^T; backward: with n_i = max(sqrt(G_ii), eps), c = 2 / (K (K - 1)),
//   d div / d P_i = sum_j M_ij P_j,   M_ij = c / (n_i n_j) (j != i),   M_ii = -c sum_{j != i} G_ij / (n_i^3 n_j)
//   d div / d A_i[n] = P_i[n] (sum_j M_ij P_j[n] - t_i),   t_i = sum_j M_ij G_ij      (softmax backward; t_i from the Gram matrix)
__device__ __forceinline__ float block_sum_1024(float v, float* red) {
  v = wsum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 32 ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = wsum(t);
  if (threadIdx.x == 0) red[32] = t;
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max_1024(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 32 ? red[threadIdx.x] : -INFINITY;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  }
  if (threadIdx.x == 0) red[32] = t;
  __syncthreads();
  return red[32];
}

// forward, step 1: one CTA per branch: (m, l) of its softmax; CTA 0 also clears the Gram accumulator and the ticket
__global__ void __launch_bounds__(1024) div_loss_stats_kernel(const float* __restrict__ a, long long a_ld, long long n,
                                                              float* __restrict__ ml, float* __restrict__ gram, unsigned* __restrict__ ticket) {
  __shared__ float red[33];
  const int tid = threadIdx.x, k = blockIdx.x;
  if (k == 0) {
    if (tid < KMAX * KMAX) gram[tid] = 0.f;
    if (tid == 0) *ticket = 0u;
  }
  float mx = -INFINITY;
  for (long long r = tid; r < n; r += 1024) mx = fmaxf(mx, a[(long long)k * a_ld + r]);
  mx = block_max_1024(mx, red);
  float l = 0.f;
  for (long long r = tid; r < n; r += 1024) l += __expf(a[(long long)k * a_ld + r] - mx);
  l = block_sum_1024(l, red);
  if (tid == 0) { ml[k] = mx; ml[KMAX + k] = l; }
}

// forward, step 2: partial Gram matrices over row chunks, added into gram[]; the last CTA to finish writes the loss
__global__ void __launch_bounds__(256) div_loss_gram_kernel(const float* __restrict__ a, long long a_ld, int K, long long n,
                                                            const float* __restrict__ ml, float* __restrict__ gram,
                                                            unsigned* __restrict__ ticket, float* __restrict__ div) {
  __shared__ float s_m[KMAX], s_il[KMAX], s_g[KMAX * (KMAX + 1) / 2];
  __shared__ bool last;
  const int tid = threadIdx.x;
  if (tid < KMAX) { s_m[tid] = tid < K ? ml[tid] : 0.f; s_il[tid] = tid < K ? 1.f / ml[KMAX + tid] : 0.f; }
  if (tid < KMAX * (KMAX + 1) / 2) s_g[tid] = 0.f;
  __syncthreads();
  float g[KMAX * (KMAX + 1) / 2];
#pragma unroll
  for (int i = 0; i < KMAX * (KMAX + 1) / 2; ++i) g[i] = 0.f;
  for (long long r = (long long)blockIdx.x * 256 + tid; r < n; r += (long long)gridDim.x * 256) {
    float pv[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) pv[k] = k < K ? __expf(a[(long long)k * a_ld + r] - s_m[k]) * s_il[k] : 0.f;
    int q = 0;
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
#pragma unroll
      for (int j = i; j < KMAX; ++j) g[q] = fmaf(pv[i], pv[j], g[q]), ++q;
  }
  {
    int q = 0;
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
#pragma unroll
      for (int j = i; j < KMAX; ++j) {
        if (i < K && j < K) {
          const float t = wsum(g[q]);
          if ((tid & 31) == 0) atomicAdd(&s_g[q], t);
        }
        ++q;
      }
  }
  __syncthreads();
  if (tid < KMAX * (KMAX + 1) / 2) {
    int i = 0, rem = tid;      // q -> (i, j) of the upper triangle
    while (rem >= KMAX - i) { rem -= KMAX - i; ++i; }
    const int j = i + rem;
    if (i < K && j < K) {
      atomicAdd(&gram[i * KMAX + j], s_g[tid]);
      if (i != j) atomicAdd(&gram[j * KMAX + i], s_g[tid]);
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && tid == 0) {
    __threadfence();
    float d = 0.f;
    for (int i = 0; i < K; ++i)
      for (int j = i + 1; j < K; ++j)
        d += __ldcg(&gram[i * KMAX + j]) / (fmaxf(sqrtf(__ldcg(&gram[i * KMAX + i])), 1e-8f) * fmaxf(sqrtf(__ldcg(&gram[j * KMAX + j])), 1e-8f));
    *div = K > 1 ? d / ((float)K * (float)(K - 1) * 0.5f) : 0.f;
  }
}

__global__ void __launch_bounds__(256) div_loss_bwd_kernel(const float* __restrict__ a, long long a_ld, int K, long long n,
                                                           const float* __restrict__ ml, const float* __restrict__ gram,
                                                           const float* __restrict__ g_out, float* __restrict__ ds, long long ds_ld) {
  __shared__ float s_M[KMAX * KMAX], s_t[KMAX], s_m[KMAX], s_il[KMAX];
  const int tid = threadIdx.x;
  if (tid < KMAX * KMAX) {
    const int i = tid / KMAX, j = tid % KMAX;
    float v = 0.f;
    if (i < K && j < K && K > 1) {
      const float c = 2.f / ((float)K * (float)(K - 1));
      const float ni = fmaxf(sqrtf(gram[i * KMAX + i]), 1e-8f);
      if (i != j) {
        v = c / (ni * fmaxf(sqrtf(gram[j * KMAX + j]), 1e-8f));
      } else {
        float sacc = 0.f;
        for (int q = 0; q < K; ++q)
          if (q != i) sacc += gram[i * KMAX + q] / fmaxf(sqrtf(gram[q * KMAX + q]), 1e-8f);
        v = -c * sacc / (ni * ni * ni);
      }
    }
    s_M[tid] = v;
  }
  if (tid < KMAX) { s_m[tid] = ml[tid]; s_il[tid] = tid < K ? 1.f / ml[KMAX + tid] : 0.f; }
  __syncthreads();
  if (tid < KMAX) {
    float t = 0.f;
    for (int j = 0; j < K; ++j) t = fmaf(s_M[tid * KMAX + j], gram[tid * KMAX + j], t);
    s_t[tid] = t;
  }
  __syncthreads();
  const float go = *g_out;
  for (long long r = (long long)blockIdx.x * 256 + tid; r < n; r += (long long)gridDim.x * 256) {
    float pv[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) pv[k] = k < K ? __expf(a[(long long)k * a_ld + r] - s_m[k]) * s_il[k] : 0.f;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      if (i < K) {
        float v = -s_t[i];
#pragma unroll
        for (int j = 0; j < KMAX; ++j) v = fmaf(s_M[i * KMAX + j], pv[j], v);
        ds[(long long)i * ds_ld + r] = go * pv[i] * v;
      }
    }
  }
}

int sm_count_cur();
int bwd_grid() { return 4 * sm_count_cur(); }      // persistent grid of the two row kernels (<= 4 CTAs per SM)

int sm_count_cur() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
  return n;
}

}  // namespace

extern "C" {

ACMIL_API int acmil_gp_bwd_workspace_floats(int32_t d_inner, int64_t* gate_floats, int64_t* relu_floats) {
  const int grid = bwd_grid();
  if (gate_floats) *gate_floats = (int64_t)grid * PW;
  if (relu_floats) *relu_floats = (int64_t)grid * d_inner;
  return ACMIL_OK;
}

ACMIL_API int acmil_gp_bwd_gate(const acmil_gp_bwd_gate_args* g, void* stream) {
  ACMIL_REQUIRE(g != nullptr, ACMIL_E_INVALID, "gp_bwd_gate: null args");
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(g->d_inner >= 32 && g->d_inner <= LMAX && g->d_inner % 32 == 0, ACMIL_E_UNSUPPORTED,
                "gp_bwd_gate: d_inner must be a multiple of 32 in [32, %d] (got %d)", LMAX, g->d_inner);
  ACMIL_REQUIRE(g->d_attn == DA, ACMIL_E_UNSUPPORTED, "gp_bwd_gate: d_attn must be %d (got %d)", DA, g->d_attn);
  ACMIL_REQUIRE(g->n_branch >= 1 && g->n_branch <= KMAX, ACMIL_E_INVALID, "gp_bwd_gate: n_branch must be in [1, %d]", KMAX);
  ACMIL_REQUIRE(g->n >= 0 && g->d_h && g->d_z && g->d_scores && g->d_lse_m && g->d_lse_l && g->d_afeat && g->d_ww && g->d_dz &&
                    g->d_dzt && g->d_dhp && g->d_partials && g->d_small,
                ACMIL_E_INVALID, "gp_bwd_gate: null operand");
  cudaStream_t st = (cudaStream_t)stream;
  const int zc = g->gated ? 2 * DA : DA;
  GateArgs a;
  a.h = g->d_h; a.z = g->d_z; a.scores = g->d_scores; a.lse_m = g->d_lse_m; a.lse_l = g->d_lse_l; a.afeat = g->d_afeat;
  a.g_afeat = g->d_g_afeat; a.g_bag = g->d_g_bag; a.g_scores = g->d_g_scores; a.ww = g->d_ww;
  a.n = g->n; a.a_ld = g->a_ld; a.gs_ld = g->gs_ld; a.ldt = g->ldt;
  a.L = g->d_inner; a.K = g->n_branch; a.zc = zc; a.act_a = g->act_a; a.gated = g->gated;
  a.ntiles = (int)((g->n + TR - 1) / TR);
  a.dz = g->d_dz; a.dzt = g->d_dzt; a.dhp = g->d_dhp; a.partials = g->d_partials;
  const int grid = std::max(1, std::min(a.ntiles, bwd_grid()));
  const int kb = g->n_branch == 1 ? 1 : (g->n_branch <= 5 ? 5 : KMAX);
  const int nf = g->d_inner / 32;
  const size_t smem = sizeof(float) * ((size_t)kb * g->d_inner + (size_t)kb * DA + 3 * kb + PW + (size_t)TR * (2 * DA + 1));
  int rc = ACMIL_E_UNSUPPORTED;
#define GATE_CASE(NFV, KBV)                                                                                              \
  if (nf == NFV && kb == KBV) {                                                                                          \
    static bool configured = false;                                                                                      \
    if (!configured) {                                                                                                   \
      ACMIL_CHECK_CUDA(cudaFuncSetAttribute(gp_bwd_gate_kernel<NFV, KBV>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                            (int)(sizeof(float) * ((size_t)KBV * NFV * 32 + KBV * DA + 3 * KBV + PW +   \
                                                                   (size_t)TR * (2 * DA + 1)))));                       \
      configured = true;                                                                                                 \
    }                                                                                                                    \
    gp_bwd_gate_kernel<NFV, KBV><<<grid, BT, smem, st>>>(a);                                                             \
    rc = ACMIL_OK;                                                                                                       \
  }
  GATE_CASE(4, 1) GATE_CASE(4, 5) GATE_CASE(4, 8) GATE_CASE(8, 1) GATE_CASE(8, 5) GATE_CASE(8, 8)
  GATE_CASE(16, 1) GATE_CASE(16, 5) GATE_CASE(16, 8)
#undef GATE_CASE
  ACMIL_REQUIRE(rc == ACMIL_OK, ACMIL_E_UNSUPPORTED, "gp_bwd_gate: d_inner must be 128, 256 or 512 (got %d)", g->d_inner);
  gp_bwd_sum_partials_kernel<<<(PW + 7) / 8, 256, 0, st>>>(g->d_partials, grid, PW, g->d_small);
  g_acmil_launches += 2;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

ACMIL_API int acmil_gp_bwd_relu_mask(const float* d_dh, const float* d_h, int64_t n, int32_t d_inner, float* d_dz1, float* d_dz1t,
                                     int64_t ldt, float* d_partials, float* d_db1, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_inner >= 32 && d_inner <= LMAX && d_inner % 32 == 0, ACMIL_E_UNSUPPORTED,
                "gp_bwd_relu_mask: d_inner must be a multiple of 32 in [32, %d] (got %d)", LMAX, d_inner);
  ACMIL_REQUIRE(d_dh && d_h && d_dz1t && d_partials && d_db1 && n >= 0 && ldt >= n, ACMIL_E_INVALID, "gp_bwd_relu_mask: bad operand");
  cudaStream_t st = (cudaStream_t)stream;
  const int ntiles = (int)((n + TR - 1) / TR);
  const int grid = std::max(1, std::min(ntiles, bwd_grid()));
  const size_t smem = sizeof(float) * ((size_t)TR * (d_inner + 1) + d_inner);
  int rc = ACMIL_E_UNSUPPORTED;
#define RELU_CASE(NFV)                                                                                                   \
  if (d_inner == NFV * 32) {                                                                                             \
    static bool configured = false;                                                                                      \
    if (!configured) {                                                                                                   \
      ACMIL_CHECK_CUDA(cudaFuncSetAttribute(gp_bwd_relu_mask_kernel<NFV>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                            (int)(sizeof(float) * ((size_t)TR * (NFV * 32 + 1) + NFV * 32))));          \
      configured = true;                                                                                                 \
    }                                                                                                                    \
    gp_bwd_relu_mask_kernel<NFV><<<grid, BT, smem, st>>>(d_dh, d_h, n, ntiles, d_dz1, d_dz1t, ldt, d_partials);          \
    rc = ACMIL_OK;                                                                                                       \
  }
  RELU_CASE(4) RELU_CASE(8) RELU_CASE(16)
#undef RELU_CASE
  ACMIL_REQUIRE(rc == ACMIL_OK, ACMIL_E_UNSUPPORTED, "gp_bwd_relu_mask: d_inner must be 128, 256 or 512 (got %d)", d_inner);
  gp_bwd_sum_partials_kernel<<<(d_inner + 7) / 8, 256, 0, st>>>(d_partials, grid, d_inner, d_db1);
  g_acmil_launches += 2;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

ACMIL_API int acmil_transpose_f32(const float* d_x, int64_t ldx, int64_t rows, int32_t cols, float* d_out, int64_t ldo, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_x && d_out && rows >= 0 && cols > 0 && ldx >= cols && ldo >= rows, ACMIL_E_INVALID, "transpose_f32: bad operand");
  if (rows == 0) return ACMIL_OK;
  const long long gx = (rows + 31) / 32;
  ACMIL_REQUIRE(gx < (1ll << 31), ACMIL_E_INVALID, "transpose_f32: too many rows");
  transpose_f32_kernel<<<dim3((unsigned)gx, (unsigned)((cols + 31) / 32)), 256, 0, (cudaStream_t)stream>>>(d_x, ldx, rows, cols, d_out, ldo);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

ACMIL_API int acmil_div_loss_fwd(const float* d_a, int64_t a_ld, int32_t n_branch, int64_t n, float* d_ml, float* d_gram, float* d_div,
                                 void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_a && d_ml && d_gram && d_div && n > 0 && a_ld >= n && n_branch >= 1 && n_branch <= KMAX, ACMIL_E_INVALID,
                "div_loss_fwd: bad operand");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* ticket = reinterpret_cast<unsigned*>(d_ml + 2 * KMAX);      // ml is [2 KMAX + 1] 4-byte words: m | l | ticket
  div_loss_stats_kernel<<<n_branch, 1024, 0, st>>>(d_a, a_ld, n, d_ml, d_gram, ticket);
  const int grid = (int)std::max<long long>(1, std::min<long long>((n + 1023) / 1024, sm_count_cur()));
  div_loss_gram_kernel<<<grid, 256, 0, st>>>(d_a, a_ld, n_branch, n, d_ml, d_gram, ticket, d_div);
  g_acmil_launches += 2;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

ACMIL_API int acmil_div_loss_bwd(const float* d_a, int64_t a_ld, int32_t n_branch, int64_t n, const float* d_ml, const float* d_gram,
                                 const float* d_g_out, float* d_ds, int64_t ds_ld, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_a && d_ml && d_gram && d_g_out && d_ds && n > 0 && a_ld >= n && ds_ld >= n && n_branch >= 1 && n_branch <= KMAX,
                ACMIL_E_INVALID, "div_loss_bwd: bad operand");
  const int grid = (int)std::min<long long>((n + 255) / 256, 2 * sm_count_cur());
  div_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_a, a_ld, n_branch, n, d_ml, d_gram, d_g_out, d_ds, ds_ld);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

}  // extern "C"
