// tm_ops.cu -- the non-GEMM kernels of the TransMIL / Nystrom path and the composite forward:
//   NystromAttention.forward (architecture/nystrom_attention.py:67-140, mask = None, return_attn = False)
//   PPEG.forward             (architecture/transMIL.py:38-45)
//   nn.LayerNorm             (architecture/transMIL.py:11, 58)
// All products run through tm_gemm (3xTF32 on tcgen05); this file holds the memory-bound glue.
#include <algorithm>
#include <cstdlib>
#include <utility>

#include "acmil_transmil.h"
#include "gp_common.cuh"

int tm_gemm(const acmil_gemm_desc& d, cudaStream_t st);

namespace {

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row (two-pass variance like the reference's fp32 kernel).
// w == nullptr: plain copy (no norm).
__global__ void __launch_bounds__(256) tm_layernorm_kernel(const float* __restrict__ x, long long ldx, long long rows, int dim,
                                                           const float* __restrict__ w, const float* __restrict__ b, float eps,
                                                           float* __restrict__ out, long long ldo) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float* orow = out + row * ldo;
  if (w == nullptr) {
    for (int j = lane; j < dim; j += 32) orow[j] = xr[j];
    return;
  }
  float s = 0.f;
  for (int j = lane; j < dim; j += 32) s += xr[j];
  const float mean = warp_sum(s) / (float)dim;
  float v = 0.f;
  for (int j = lane; j < dim; j += 32) {
    const float t = xr[j] - mean;
    v = fmaf(t, t, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)dim + eps);
  for (int j = lane; j < dim; j += 32) orow[j] = (xr[j] - mean) * rstd * w[j] + (b ? b[j] : 0.f);
}

// ------------------------------------------------------------------------------------------
// landmarks: dst[hd][j][:] = sum_{t < l} src[hd][j * l + t][:] / l      (nystrom_attention.py:98-114)
__global__ void __launch_bounds__(256) tm_landmark_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_pad, int m,
                                                          int l, int d) {
  const int j = blockIdx.x, hd = blockIdx.y;
  const float* s = src + ((size_t)hd * n_pad + (size_t)j * l) * d;
  // threads: (dd, part); the parts' partial sums are combined through smem in a fixed order
  __shared__ float red[256];
  const int parts = 256 / d > 0 ? 256 / d : 1;
  const int dd = threadIdx.x % d, part = threadIdx.x / d;
  float acc = 0.f;
  if (part < parts)
    for (int t = part; t < l; t += parts) acc += s[(size_t)t * d + dd];
  red[threadIdx.x] = part < parts ? acc : 0.f;
  __syncthreads();
  if (threadIdx.x < d) {
    float t = 0.f;
    for (int q = 0; q < parts; ++q) t += red[q * d + threadIdx.x];
    dst[((size_t)hd * m + j) * d + threadIdx.x] = t / (float)l;
  }
}

// ------------------------------------------------------------------------------------------
// in-place softmax over rows of length len <= 32 * NI <= 1024: one warp per row, the row lives in registers
template <int NI>
__global__ void __launch_bounds__(256) tm_softmax_small_kernel_t(float* __restrict__ a, long long rows, int len, long long ld) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = a + row * ld;
  float v[NI];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = lane + 32 * i;
    v[i] = j < len ? r[j] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    v[i] = expf(v[i] - mx);      // exp(-inf) = 0 for the slots past the row
    s += v[i];
  }
  s = warp_sum(s);
#pragma unroll
  for (int i = 0; i < NI; ++i)
    if (lane + 32 * i < len) r[lane + 32 * i] = v[i] / s;
}
// launch with the smallest register tile that holds the row
static inline void tm_softmax_small_launch(float* a, long long rows, int len, long long ld, cudaStream_t st) {
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (len <= 64) tm_softmax_small_kernel_t<2><<<grid, 256, 0, st>>>(a, rows, len, ld);
  else if (len <= 128) tm_softmax_small_kernel_t<4><<<grid, 256, 0, st>>>(a, rows, len, ld);
  else if (len <= 256) tm_softmax_small_kernel_t<8><<<grid, 256, 0, st>>>(a, rows, len, ld);
  else if (len <= 512) tm_softmax_small_kernel_t<16><<<grid, 256, 0, st>>>(a, rows, len, ld);
  else tm_softmax_small_kernel_t<32><<<grid, 256, 0, st>>>(a, rows, len, ld);
}

// in-place softmax over long rows: one CTA per row, three passes (the row stays in L2)
__global__ void __launch_bounds__(1024) tm_softmax_long_kernel(float* __restrict__ a, long long len) {
  __shared__ float red[32];
  __shared__ float bc;
  float* r = a + (size_t)blockIdx.x * len;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -INFINITY;
  for (long long j = tid; j < len; j += 1024) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_max(red[lane]);
    if (lane == 0) bc = t;
  }
  __syncthreads();
  mx = bc;
  float s = 0.f;
  for (long long j = tid; j < len; j += 1024) {
    const float e = expf(r[j] - mx);
    r[j] = e;
    s += e;
  }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_sum(red[lane]);
    if (lane == 0) bc = t;
  }
  __syncthreads();
  s = bc;
  for (long long j = tid; j < len; j += 1024) r[j] = r[j] / s;
}

// the same with the row staged in shared memory (rows up to ~56 k floats: the 50 k-token attn3 rows of TransMIL): one read
// and one write of the row instead of three reads and two writes through L2
__global__ void __launch_bounds__(1024) tm_softmax_long_smem_kernel(float* __restrict__ a, long long len) {
  extern __shared__ float4 row4[];
  __shared__ float red[32];
  __shared__ float bc;
  float4* r4 = reinterpret_cast<float4*>(a + (size_t)blockIdx.x * len);
  const int n4 = (int)(len >> 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -INFINITY;
  for (int j0 = tid; j0 < n4; j0 += 8 * 1024) {      // eight independent 16-byte loads in flight per thread
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {      // volatile asm + the barrier below: ptxas otherwise pairs every load with its store
      const int j = j0 + u * 1024;
      v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (j < n4)
        asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(r4 + j));
    }
    asm volatile("" ::: "memory");
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u * 1024;
      if (j < n4) row4[j] = v[u];
      mx = fmaxf(fmaxf(mx, fmaxf(v[u].x, v[u].y)), fmaxf(v[u].z, v[u].w));
    }
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_max(red[lane]);
    if (lane == 0) bc = t;
  }
  __syncthreads();
  mx = bc;
  float s = 0.f;
  for (int j = tid; j < n4; j += 1024) {      // each thread revisits the elements it staged itself
    float4 v = row4[j];
    v.x = expf(v.x - mx); v.y = expf(v.y - mx); v.z = expf(v.z - mx); v.w = expf(v.w - mx);
    row4[j] = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_sum(red[lane]);
    if (lane == 0) bc = t;
  }
  __syncthreads();
  s = bc;
  for (int j = tid; j < n4; j += 1024) {
    float4 v = row4[j];
    v.x = v.x / s; v.y = v.y / s; v.z = v.z / s; v.w = v.w / s;
    r4[j] = v;
  }
}

int tm_softmax_long_launch(float* a, long long rows, long long len, cudaStream_t st) {
  const size_t bytes = (size_t)len * sizeof(float);
  if (len % 4 == 0 && bytes <= 220 * 1024 && ((uintptr_t)a & 15) == 0) {
    static bool configured = false;
    if (!configured) {
      ACMIL_CHECK_CUDA(cudaFuncSetAttribute(tm_softmax_long_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      configured = true;
    }
    tm_softmax_long_smem_kernel<<<(unsigned)rows, 1024, bytes, st>>>(a, len);
  } else {
    tm_softmax_long_kernel<<<(unsigned)rows, 1024, 0, st>>>(a, len);
  }
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

// ------------------------------------------------------------------------------------------
// Moore-Penrose start value (nystrom_attention.py:12-19): z0 = x^T / (max_i sum_j |x_ij| * max_j sum_i |x_ij|),
// the maxima taken over ALL batch entries and heads (torch.max of the whole tensor).
__global__ void __launch_bounds__(256) tm_pinv_sums_kernel(const float* __restrict__ x, int m, int* __restrict__ scal) {
  const float* xh = x + (size_t)blockIdx.x * m * m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float best = 0.f;
  if (blockIdx.y == 0) {      // row sums: a warp per row
    for (int i = warp; i < m; i += 8) {
      float s = 0.f;
      for (int j = lane; j < m; j += 32) s += fabsf(xh[(size_t)i * m + j]);
      best = fmaxf(best, warp_sum(s));
    }
  } else {                    // column sums: a thread per column
    for (int j = tid; j < m; j += 256) {
      float s = 0.f;
      for (int i = 0; i < m; ++i) s += fabsf(xh[(size_t)i * m + j]);
      best = fmaxf(best, s);
    }
    best = warp_max(best);
  }
  if (lane == 0) atomicMax(&scal[blockIdx.y], __float_as_int(best));      // non-negative floats order like ints
}

__global__ void __launch_bounds__(256) tm_pinv_init_kernel(const float* __restrict__ x, int m, const int* __restrict__ scal,
                                                           float* __restrict__ z, float* __restrict__ zt) {
  __shared__ float tile[32][33];
  const float denom = __int_as_float(scal[0]) * __int_as_float(scal[1]);
  const size_t base = (size_t)blockIdx.z * m * m;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + tx;
    float v = 0.f;
    if (i < m && j < m) {
      v = x[base + (size_t)i * m + j] / denom;
      zt[base + (size_t)i * m + j] = v;      // z^T = x / denom
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, i = i0 + tx;
    if (i < m && j < m) z[base + (size_t)j * m + i] = tile[tx][r];
  }
}

// ------------------------------------------------------------------------------------------
// depth-wise residual conv of the values along the sequence (nystrom_attention.py:62-65, 137-138):
//   merged[b][i - shift][c] += sum_t w[c / d][t] * v[b][c][i + t - ks/2]     with v stored transposed, vt[b][c][i]
// (shift = 0 for a whole sequence; for a sequence shard v^T lives in a halo-extended buffer whose column i is local row
// i - shift, the halo columns carrying the neighbours' values).  Tile: 32 channels x 64 positions per CTA, staged in shared
// memory with reads coalesced along i; a thread owns one channel and 8 consecutive positions: its 8 + KS - 1 inputs and the
// KS taps sit in registers (KS = 33, the reference's kernel; KS = 0 is the generic loop), writes are coalesced along c.
constexpr int CONV_TP = 64;
template <int KS>
__global__ void __launch_bounds__(256) tm_resconv_kernel(const float* __restrict__ vt, const float* __restrict__ w,
                                                         float* __restrict__ merged, int n_cols, int inner, int d, int ks,
                                                         int col0, int nrows, int shift, long long m_rows) {
  extern __shared__ float tile[];                 // [32][span], span odd
  const int half = ks / 2;
  const int width = CONV_TP + ks - 1;
  const int span = width | 1;
  const int i0 = col0 + blockIdx.x * CONV_TP, c0 = blockIdx.y * 32, bz = blockIdx.z;
  const float* vb = vt + ((size_t)bz * inner + c0) * n_cols;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cx = lane, py = warp;
  const bool c_mine = c0 + cx < inner;
  float* mb = merged + (size_t)bz * m_rows * inner + c0 + cx;
  const int end = col0 + nrows;
  // every global load of this thread is issued before anything waits: the 8 values its outputs are added to ...
  float mv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = i0 + py * 8 + j;
    mv[j] = (c_mine && i < end) ? mb[(size_t)(i - shift) * inner] : 0.f;
  }
  // ... and its share of the tile (channel rows warp, warp + 8, ..; coalesced along i)
  for (int ob = 0; ob < width; ob += 96) {      // one round for ks <= 33
    float tv[12];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int c = warp + 8 * a, o = ob + lane + 32 * b, i = i0 + o - half;
        tv[a * 3 + b] = (o < width && c0 + c < inner && i >= 0 && i < n_cols) ? vb[(size_t)c * n_cols + i] : 0.f;
      }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int c = warp + 8 * a, o = ob + lane + 32 * b;
        if (o < width) tile[c * span + o] = tv[a * 3 + b];
      }
  }
  __syncthreads();
  if (!c_mine) return;
  const float* wh = w + (size_t)((c0 + cx) / d) * ks;
  if constexpr (KS > 0) {
    float win[KS + 7];
#pragma unroll
    for (int t = 0; t < KS + 7; ++t) win[t] = tile[cx * span + py * 8 + t];
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int t = 0; t < KS; ++t) {
      const float wt = __ldg(wh + t);      // the same address in every lane of the CTA (one head per 32 channels when d % 32 == 0)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, win[j + t], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + py * 8 + j;
      if (i < end) mb[(size_t)(i - shift) * inner] = mv[j] + acc[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = py * 8 + j, i = i0 + o;
      if (i >= end) break;
      float acc = 0.f;
      for (int t = 0; t < ks; ++t) acc = fmaf(__ldg(wh + t), tile[cx * span + o + t], acc);
      mb[(size_t)(i - shift) * inner] = mv[j] + acc;
    }
  }
}

void tm_resconv_launch(dim3 grid, size_t smem, cudaStream_t st, const float* vt, const float* w, float* merged, int n_cols, int inner,
                       int d, int ks, int col0, int nrows, int shift, long long m_rows) {
  if (ks == 33) tm_resconv_kernel<33><<<grid, 256, smem, st>>>(vt, w, merged, n_cols, inner, d, ks, col0, nrows, shift, m_rows);
  else tm_resconv_kernel<0><<<grid, 256, smem, st>>>(vt, w, merged, n_cols, inner, d, ks, col0, nrows, shift, m_rows);
}

// ------------------------------------------------------------------------------------------
// PPEG (transMIL.py:38-45): out = feat + conv7(feat) + conv5(feat) + conv3(feat) on the [gh, gw] token grid,
// channels last.  The three depth-wise kernels and the identity are summed into one 7x7 stencil per channel
// (exact up to fp32 summation order).  CTA = 32 channels x (8 rows x 16 columns): the (8 + 6) x (16 + 6) input halo tile
// and the 49 combined taps of the 32 channels are staged in shared memory once (every warp load is one 128-byte line:
// 32 consecutive channels of a token; 38 independent loads per thread in flight instead of 22 per stencil row behind a
// dependent FMA block -- the direct-from-global version ran at a tenth of the HBM roofline), then a thread owns one
// channel and a run of 16 outputs along x, every input feeding up to 7 FMAs from registers.
constexpr int PPEG_TX = 16, PPEG_TY = 8;
__global__ void __launch_bounds__(256) tm_ppeg_kernel(const float* __restrict__ x, int gh, int gw, int C, const float* __restrict__ w7,
                                                      const float* __restrict__ b7, const float* __restrict__ w5,
                                                      const float* __restrict__ b5, const float* __restrict__ w3,
                                                      const float* __restrict__ b3, float* __restrict__ out, int cblocks,
                                                      int y_first, int y_end) {      // grid rows [y_first, y_end) are produced
  __shared__ float in_s[PPEG_TY + 6][PPEG_TX + 6][32];
  __shared__ float w_s[49][32];
  const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = (blockIdx.z % cblocks) * 32, bz = blockIdx.z / cblocks;
  const int c = c0 + cx;
  const bool c_ok = c < C;
  const size_t tok = (size_t)gh * gw + 1;
  const float* xb = x + (size_t)bz * tok * C;
  float* ob = out + (size_t)bz * tok * C;
  if (c_ok && blockIdx.x == 0 && blockIdx.y == 0 && ty == 0 && y_first == 0 && y_end == gh) ob[c] = xb[c];      // class token passes through
  const int y0 = y_first + blockIdx.y * PPEG_TY, x0 = blockIdx.x * PPEG_TX;
  const float* feat = xb + C;
  constexpr int NPIX = (PPEG_TY + 6) * (PPEG_TX + 6);
  for (int base = ty; base < NPIX; base += 80) {      // warp = one token's 32 channels; ten independent loads in flight per thread
    float v[10];
#pragma unroll
    for (int u = 0; u < 10; ++u) {
      const int pix = base + 8 * u;
      const int py = pix / (PPEG_TX + 6), px = pix - py * (PPEG_TX + 6);
      const int yy = y0 + py - 3, xx = x0 + px - 3;
      v[u] = (pix < NPIX && c_ok && yy >= 0 && yy < gh && xx >= 0 && xx < gw) ? feat[((size_t)yy * gw + xx) * C + c] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 10; ++u) {
      const int pix = base + 8 * u;
      if (pix < NPIX) in_s[pix / (PPEG_TX + 6)][pix % (PPEG_TX + 6)][cx] = v[u];
    }
  }
  {
    float tw[7];
#pragma unroll
    for (int u = 0; u < 7; ++u) {      // taps ty, ty + 8, ..: all loads issued before the first store
      const int k = ty + 8 * u;
      const int dy = k / 7, dx = k - dy * 7;
      float t = 0.f;
      if (c_ok && k < 49) {
        t = w7[(size_t)c * 49 + k];
        if (dy >= 1 && dy <= 5 && dx >= 1 && dx <= 5) t += w5[(size_t)c * 25 + (dy - 1) * 5 + (dx - 1)];
        if (dy >= 2 && dy <= 4 && dx >= 2 && dx <= 4) t += w3[(size_t)c * 9 + (dy - 2) * 3 + (dx - 2)];
        if (k == 24) t += 1.f;
      }
      tw[u] = t;
    }
#pragma unroll
    for (int u = 0; u < 7; ++u)
      if (ty + 8 * u < 49) w_s[ty + 8 * u][cx] = tw[u];
  }
  __syncthreads();
  const int y = y0 + ty;
  if (!c_ok || y >= y_end) return;
  const float bias = b7[c] + b5[c] + b3[c];
  float acc[PPEG_TX];
#pragma unroll
  for (int o = 0; o < PPEG_TX; ++o) acc[o] = bias;
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    float in[PPEG_TX + 6];
#pragma unroll
    for (int i = 0; i < PPEG_TX + 6; ++i) in[i] = in_s[ty + dy][i][cx];
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
      const float wk = w_s[dy * 7 + dx][cx];
#pragma unroll
      for (int o = 0; o < PPEG_TX; ++o) acc[o] = fmaf(wk, in[o + dx], acc[o]);
    }
  }
#pragma unroll
  for (int o = 0; o < PPEG_TX; ++o)
    if (x0 + o < gw) ob[(size_t)(1 + y * gw + x0 + o) * C + c] = acc[o];
}

// ------------------------------------------------------------------------------------------
struct NysLayout {      // float offsets into the workspace
  size_t xn, q, k, vt, ql, kl, a2, y, yt, t1t, t2t, za, zta, zb, ztb, kvt, wt, sbuf, merged, split, scal, stats, total;
  int n_pad, l, pad, inner, H;
  int ksplit;
};

size_t align64(size_t v) { return (v + 63) & ~(size_t)63; }

NysLayout nys_layout(const acmil_nystrom_shape& s) {
  NysLayout L{};
  const int m = s.num_landmarks;
  L.l = (s.n + m - 1) / m;
  L.n_pad = s.n % m ? m * L.l : s.n;
  L.pad = L.n_pad - s.n;
  L.inner = s.heads * s.dim_head;
  L.H = s.batch * s.heads;
  const size_t d = s.dim_head, np = L.n_pad, H = L.H;
  // K split of the (attn3 v) product: enough CTAs to cover the GPU (2 tiles per batch entry otherwise)
  L.ksplit = (int)std::max<size_t>(1, std::min<size_t>(64, (np / 32) / 48));
  size_t o = 0;
  auto take = [&](size_t n) { const size_t r = o; o += align64(n); return r; };
  L.xn = take((size_t)s.batch * np * s.dim);
  L.q = take(H * np * d);
  L.k = take(H * np * d);
  L.vt = take(H * d * np);
  L.ql = take(H * m * d);
  L.kl = take(H * m * d);
  const size_t mm = H * (size_t)m * m;
  L.a2 = take(mm); L.y = take(mm); L.yt = take(mm); L.t1t = take(mm); L.t2t = take(mm);
  L.za = take(mm); L.zta = take(mm); L.zb = take(mm); L.ztb = take(mm);
  L.kvt = take(H * d * m);
  L.wt = take(H * d * m);
  L.sbuf = take(H * np * m);
  L.merged = take((size_t)s.batch * np * L.inner);
  L.split = take((size_t)L.ksplit * H * d * m);
  L.scal = take(64);
  L.stats = take(H * np * (size_t)((m + 31) / 32) * 2);      // (max, sum) per 32-landmark chunk of every attn1 row
  L.total = o;
  return L;
}

int check_shape(const acmil_nystrom_shape& s) {
  ACMIL_REQUIRE(s.batch >= 1 && s.n >= 1 && s.dim >= 4 && s.dim % 4 == 0, ACMIL_E_INVALID,
                "nystrom: bad x shape [%d, %d, %d] (dim must be a multiple of 4)", s.batch, s.n, s.dim);
  ACMIL_REQUIRE(s.heads >= 1 && s.dim_head >= 4 && s.dim_head % 4 == 0 && s.dim_head <= 256, ACMIL_E_INVALID,
                "nystrom: dim_head %d must be a multiple of 4 in [4, 256]", s.dim_head);
  ACMIL_REQUIRE(s.num_landmarks >= 4 && s.num_landmarks % 4 == 0 && s.num_landmarks <= 1024, ACMIL_E_INVALID,
                "nystrom: num_landmarks %d must be a multiple of 4 in [4, 1024]", s.num_landmarks);
  ACMIL_REQUIRE(s.pinv_iterations >= 0 && s.n_out >= 0 && s.n_out <= s.n, ACMIL_E_INVALID, "nystrom: bad pinv_iterations / n_out");
  ACMIL_REQUIRE(!s.residual || (s.conv_kernel >= 1 && s.conv_kernel % 2 == 1 && s.conv_kernel <= 129), ACMIL_E_INVALID,
                "nystrom: residual conv kernel %d must be odd and <= 129", s.conv_kernel);
  ACMIL_REQUIRE((long long)s.batch * s.heads <= 1024, ACMIL_E_INVALID, "nystrom: batch * heads too large");
  return ACMIL_OK;
}

acmil_gemm_desc gemm0(int precise) {
  acmil_gemm_desc g{};
  g.alpha = 1.f;
  g.precise = precise;
  return g;
}

// softmax between two products folded into them (acmil_gemm_desc.softmax_stats_*); on by default (measured: + 0.7 % TransMIL,
// + 1.5 % ViT against the separate softmax pass), ACMIL_CHUNKED_SOFTMAX=0 restores the pass
bool chunked_softmax(bool dflt) {
  static const int env = [] {
    const char* e = getenv("ACMIL_CHUNKED_SOFTMAX");
    return e == nullptr || *e == 0 ? -1 : (*e != '0' ? 1 : 0);
  }();
  return env < 0 ? dflt : env != 0;
}

// B = rows [row0, row0 + g.n) of a weight matrix with `rows` rows: use its pre-split image when there is one (precise = 2)
void use_split(acmil_gemm_desc& g, const void* image, int rows, int row0) {
  if (g.precise != 2 || image == nullptr) return;
  g.b_split = image;
  g.b_split_rows = rows;
  g.b_split_row0 = row0;
}

}  // namespace

extern "C" int acmil_layernorm_rows(const float* d_x, int64_t ldx, int64_t rows, int32_t dim, const float* d_w, const float* d_b,
                                    float eps, float* d_out, int64_t ldo, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_x && d_out && rows >= 0 && dim >= 1, ACMIL_E_INVALID, "layernorm: bad arguments");
  if (rows == 0) return ACMIL_OK;
  tm_layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(d_x, ldx, rows, dim, d_w, d_b, eps, d_out, ldo);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

extern "C" int acmil_ppeg_fwd(const float* d_x, int32_t batch, int32_t gh, int32_t gw, int32_t c, const float* d_w7, const float* d_b7,
                              const float* d_w5, const float* d_b5, const float* d_w3, const float* d_b3, float* d_out,
                              void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_x && d_out && d_w7 && d_b7 && d_w5 && d_b5 && d_w3 && d_b3, ACMIL_E_INVALID, "ppeg: null pointer");
  ACMIL_REQUIRE(batch >= 1 && gh >= 1 && gw >= 1 && c >= 1 && batch <= 65535, ACMIL_E_INVALID, "ppeg: bad shape");
  const int cblocks = (c + 31) / 32;
  ACMIL_REQUIRE((long long)cblocks * batch <= 65535 && (gh + PPEG_TY - 1) / PPEG_TY <= 65535, ACMIL_E_INVALID, "ppeg: grid too large");
  dim3 grid((gw + PPEG_TX - 1) / PPEG_TX, (gh + PPEG_TY - 1) / PPEG_TY, cblocks * batch);
  tm_ppeg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x, gh, gw, c, d_w7, d_b7, d_w5, d_b5, d_w3, d_b3, d_out, cblocks, 0, gh);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

// grid rows [y_first, y_first + ny) of ONE sequence only (the class token row is not touched): the sharded TransMIL keeps a
// full-size token buffer per rank in which only its own rows and a 3-grid-row halo are valid
extern "C" int acmil_ppeg_fwd_rows(const float* d_x, int32_t gh, int32_t gw, int32_t c, const float* d_w7, const float* d_b7,
                                   const float* d_w5, const float* d_b5, const float* d_w3, const float* d_b3, float* d_out,
                                   int32_t y_first, int32_t ny, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_x && d_out && d_w7 && d_b7 && d_w5 && d_b5 && d_w3 && d_b3, ACMIL_E_INVALID, "ppeg: null pointer");
  ACMIL_REQUIRE(gh >= 1 && gw >= 1 && c >= 1 && y_first >= 0 && ny >= 0 && y_first + ny <= gh, ACMIL_E_INVALID, "ppeg rows: bad shape");
  if (ny == 0) return ACMIL_OK;
  const int cblocks = (c + 31) / 32;
  ACMIL_REQUIRE(cblocks <= 65535 && (ny + PPEG_TY - 1) / PPEG_TY <= 65535, ACMIL_E_INVALID, "ppeg: grid too large");
  dim3 grid((gw + PPEG_TX - 1) / PPEG_TX, (ny + PPEG_TY - 1) / PPEG_TY, cblocks);
  tm_ppeg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x, gh, gw, c, d_w7, d_b7, d_w5, d_b5, d_w3, d_b3, d_out, cblocks, y_first,
                                                         y_first + ny);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

extern "C" int acmil_nystrom_workspace_bytes(const acmil_nystrom_shape* shape, size_t* bytes) {
  ACMIL_REQUIRE(shape && bytes, ACMIL_E_INVALID, "nystrom: null argument");
  const int rc = check_shape(*shape);
  if (rc) return rc;
  *bytes = nys_layout(*shape).total * sizeof(float);
  return ACMIL_OK;
}

#define TM_RUN(expr)            \
  do {                          \
    const int _rc = (expr);     \
    if (_rc) return _rc;        \
  } while (0)

extern "C" int acmil_nystrom_attn_fwd(const acmil_nystrom_shape* shape, const acmil_nystrom_weights* w, const float* d_x,
                                      const float* d_residual, float* d_out, void* d_workspace, size_t workspace_bytes,
                                      void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(shape && w && d_x && d_out && d_workspace, ACMIL_E_INVALID, "nystrom: null argument");
  const acmil_nystrom_shape& s = *shape;
  TM_RUN(check_shape(s));
  ACMIL_REQUIRE(w->d_wqkv && w->d_wout && w->d_bout && (!s.residual || w->d_wconv), ACMIL_E_INVALID, "nystrom: null weight");
  ACMIL_REQUIRE(((uintptr_t)d_workspace & 255) == 0, ACMIL_E_INVALID, "nystrom: workspace must be 256-byte aligned");
  const NysLayout L = nys_layout(s);
  ACMIL_REQUIRE(workspace_bytes >= L.total * sizeof(float), ACMIL_E_WORKSPACE, "nystrom: workspace %zu < %zu bytes", workspace_bytes,
                L.total * sizeof(float));
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = reinterpret_cast<float*>(d_workspace);
  const int m = s.num_landmarks, d = s.dim_head, h = s.heads, H = L.H, np = L.n_pad, inner = L.inner, dim = s.dim;
  const int P = s.precise;
  float *xn = ws + L.xn, *q = ws + L.q, *k = ws + L.k, *vt = ws + L.vt, *ql = ws + L.ql, *kl = ws + L.kl;

  // 1. (LayerNorm and) front zero padding so that the sequence splits into m landmark groups (:72-80)
  for (int b = 0; b < s.batch; ++b) {
    if (L.pad) ACMIL_CHECK_CUDA(cudaMemsetAsync(xn + (size_t)b * np * dim, 0, (size_t)L.pad * dim * sizeof(float), st));
    TM_RUN(acmil_layernorm_rows(d_x + (size_t)b * s.n * dim, dim, s.n, dim, w->d_ln_w, w->d_ln_b, w->ln_eps,
                                xn + ((size_t)b * np + L.pad) * dim, dim, stream));
  }
  // 2. q (scaled, :93), k as [b h] n d; v transposed as [b h] d n  (:83-84)
  for (int b = 0; b < s.batch; ++b) {
    acmil_gemm_desc g = gemm0(P);
    g.a = xn + (size_t)b * np * dim; g.lda = dim; g.m = np; g.k = dim; g.batch = 1;
    g.b = w->d_wqkv; g.ldb = dim; g.n = inner;
    g.c = q + (size_t)b * h * np * d; g.ldc = d; g.col_block_width = d; g.col_block_stride = (int64_t)np * d;
    g.alpha = 1.f / sqrtf((float)d);
    use_split(g, w->d_split_qkv, 3 * inner, 0);
    TM_RUN(tm_gemm(g, st));
    g.b = w->d_wqkv + (size_t)inner * dim; g.c = k + (size_t)b * h * np * d; g.alpha = 1.f;
    use_split(g, w->d_split_qkv, 3 * inner, inner);
    TM_RUN(tm_gemm(g, st));
    acmil_gemm_desc gv = gemm0(P);
    if (P == 2 && w->d_split_qkv) {      // v = xn Wv^T on the fp16-split kernel, stored transposed by the epilogue
      gv.a = xn + (size_t)b * np * dim; gv.lda = dim; gv.m = np; gv.k = dim; gv.batch = 1;
      gv.b = w->d_wqkv + (size_t)2 * inner * dim; gv.ldb = dim; gv.n = inner;
      gv.ct = vt + (size_t)b * inner * np; gv.ldct = np;
      use_split(gv, w->d_split_qkv, 3 * inner, 2 * inner);
    } else {
      gv.a = w->d_wqkv + (size_t)2 * inner * dim; gv.lda = dim; gv.m = inner; gv.k = dim; gv.batch = 1;
      gv.b = xn + (size_t)b * np * dim; gv.ldb = dim; gv.n = np;
      gv.c = vt + (size_t)b * inner * np; gv.ldc = np;
    }
    TM_RUN(tm_gemm(gv, st));
  }
  // 3. landmarks = group means (:98-114)
  {
    dim3 grid(m, H);
    tm_landmark_kernel<<<grid, 256, 0, st>>>(q, ql, np, m, L.l, d);
    tm_landmark_kernel<<<grid, 256, 0, st>>>(k, kl, np, m, L.l, d);
    g_acmil_launches += 2;
    ACMIL_CHECK_CUDA(cudaGetLastError());
  }
  // 4. attn2 = softmax(q_l k_l^T) and its Moore-Penrose pseudo-inverse (:120, 134, 12-27)
  float* a2 = ws + L.a2;
  const int64_t mm = (int64_t)m * m;
  {
    acmil_gemm_desc g = gemm0(P);
    g.a = ql; g.lda = d; g.a_batch_stride = (int64_t)m * d; g.m = m; g.k = d; g.batch = H;
    g.b = kl; g.ldb = d; g.b_batch_stride = (int64_t)m * d; g.n = m;
    g.c = a2; g.ldc = m; g.c_batch_stride = mm;
    TM_RUN(tm_gemm(g, st));
    tm_softmax_small_launch(a2, (long long)H * m, m, m, st);
    ++g_acmil_launches;
  }
  float *z = ws + L.za, *zt = ws + L.zta, *z2 = ws + L.zb, *zt2 = ws + L.ztb;
  {
    int* scal = reinterpret_cast<int*>(ws + L.scal);
    ACMIL_CHECK_CUDA(cudaMemsetAsync(scal, 0, 8, st));
    tm_pinv_sums_kernel<<<dim3(H, 2), 256, 0, st>>>(a2, m, scal);
    tm_pinv_init_kernel<<<dim3((m + 31) / 32, (m + 31) / 32, H), 256, 0, st>>>(a2, m, scal, z, zt);
    g_acmil_launches += 2;
    ACMIL_CHECK_CUDA(cudaGetLastError());
    // z <- 1/4 z (13 I - Y (15 I - Y (7 I - Y))),  Y = x z; written as
    //   T1 = 7 Y - Y Y, T2 = 15 Y - Y T1, z' = 3.25 z - 0.25 z T2     (every product A B^T needs B transposed:
    //   each GEMM also emits the transposed result that the next one consumes)
    float *y = ws + L.y, *yt = ws + L.yt, *t1t = ws + L.t1t, *t2t = ws + L.t2t;
    auto sq = [&](const float* A, const float* Bt, float* C, float* Ct, float alpha, const float* add, float beta) {
      acmil_gemm_desc g = gemm0(P);
      g.a = A; g.lda = m; g.a_batch_stride = mm; g.m = m; g.k = m; g.batch = H;
      g.b = Bt; g.ldb = m; g.b_batch_stride = mm; g.n = m;
      g.c = C; g.ldc = m; g.c_batch_stride = mm;
      g.ct = Ct; g.ldct = m; g.ct_batch_stride = mm;
      g.alpha = alpha; g.addend = add; g.ld_addend = m; g.addend_batch_stride = mm; g.beta = beta;
      return tm_gemm(g, st);
    };
    for (int it = 0; it < s.pinv_iterations; ++it) {
      TM_RUN(sq(a2, zt, y, yt, 1.f, nullptr, 0.f));
      TM_RUN(sq(y, yt, nullptr, t1t, -1.f, y, 7.f));
      TM_RUN(sq(y, t1t, nullptr, t2t, -1.f, y, 15.f));
      TM_RUN(sq(z, t2t, z2, zt2, -0.25f, z, 3.25f));
      std::swap(z, z2);
      std::swap(zt, zt2);
    }
  }
  // 5. attn3 = softmax_n(q_l k^T); (attn3 v)^T = v^T attn3^T   (:121, 133, 135)
  float* sbuf = ws + L.sbuf;
  {
    acmil_gemm_desc g = gemm0(P);
    g.a = ql; g.lda = d; g.a_batch_stride = (int64_t)m * d; g.m = m; g.k = d; g.batch = H;
    g.b = k; g.ldb = d; g.b_batch_stride = (int64_t)np * d; g.n = np;
    g.c = sbuf; g.ldc = np; g.c_batch_stride = (int64_t)m * np;
    TM_RUN(tm_gemm(g, st));
    TM_RUN(tm_softmax_long_launch(sbuf, (long long)H * m, np, st));
    ++g_acmil_launches;
    acmil_gemm_desc g2 = gemm0(P);
    g2.a = vt; g2.lda = np; g2.a_batch_stride = (int64_t)d * np; g2.m = d; g2.k = np; g2.batch = H;
    g2.b = sbuf; g2.ldb = np; g2.b_batch_stride = (int64_t)m * np; g2.n = m;
    g2.c = ws + L.kvt; g2.ldc = m; g2.c_batch_stride = (int64_t)d * m;
    g2.k_split = L.ksplit; g2.split_ws = ws + L.split;
    TM_RUN(tm_gemm(g2, st));
  }
  // 6. W^T = (attn3 v)^T pinv^T : out = attn1 (pinv (attn3 v)) -- the cheap association of (:135)
  {
    acmil_gemm_desc g = gemm0(P);
    g.a = ws + L.kvt; g.lda = m; g.a_batch_stride = (int64_t)d * m; g.m = d; g.k = m; g.batch = H;
    g.b = z; g.ldb = m; g.b_batch_stride = mm; g.n = m;
    g.c = ws + L.wt; g.ldc = m; g.c_batch_stride = (int64_t)d * m;
    TM_RUN(tm_gemm(g, st));
  }
  // 7. rows that are needed downstream: all n_pad (padded_out), the last n, or the first n_out of those
  const int r0 = s.padded_out ? 0 : L.pad;
  const int nr = s.padded_out ? np : (s.n_out > 0 ? s.n_out : s.n);
  float* merged = ws + L.merged;
  {
    // attn1 = softmax_m(q k_l^T) (:119), out = attn1 W, heads merged as b n (h d) (:141)
    acmil_gemm_desc g = gemm0(P);
    g.a = q + (size_t)r0 * d; g.lda = d; g.a_batch_stride = (int64_t)np * d; g.m = nr; g.k = d; g.batch = H;
    g.b = kl; g.ldb = d; g.b_batch_stride = (int64_t)m * d; g.n = m;
    g.c = sbuf; g.ldc = m; g.c_batch_stride = (int64_t)nr * m;
    // softmax over the m landmarks without its own pass: the scores' epilogue leaves exp(s - chunk max) and per-chunk
    // (max, sum); the converter of the next product rescales the rows (acmil_gemm_desc.softmax_stats_*)
    float* stats = P && chunked_softmax(true) ? ws + L.stats : nullptr;
    const size_t nch = (size_t)(m + 31) / 32;
    g.softmax_stats_out = stats;
    TM_RUN(tm_gemm(g, st));
    if (!stats) {
      tm_softmax_small_launch(sbuf, (long long)H * nr, m, m, st);
      ++g_acmil_launches;
      ACMIL_CHECK_CUDA(cudaGetLastError());
    }
    for (int b = 0; b < s.batch; ++b) {
      acmil_gemm_desc g2 = gemm0(P);
      if (stats) g2.softmax_stats_in = stats + (size_t)b * h * nr * nch * 2;
      g2.a = sbuf + (size_t)b * h * nr * m; g2.lda = m; g2.a_batch_stride = (int64_t)nr * m; g2.m = nr; g2.k = m; g2.batch = h;
      g2.b = ws + L.wt + (size_t)b * h * d * m; g2.ldb = m; g2.b_batch_stride = (int64_t)d * m; g2.n = d;
      g2.c = merged + ((size_t)b * np + r0) * inner; g2.ldc = inner; g2.c_batch_stride = d;
      TM_RUN(tm_gemm(g2, st));
    }
  }
  // 8. + depth-wise conv of v (:137-138)
  if (s.residual) {
    const int ks = s.conv_kernel;
    const size_t smem = (size_t)32 * ((CONV_TP + ks - 1) | 1) * sizeof(float);
    dim3 grid((nr + CONV_TP - 1) / CONV_TP, (inner + 31) / 32, s.batch);
    tm_resconv_launch(grid, smem, st, vt, w->d_wconv, merged, np, inner, d, ks, r0, nr, 0, np);
    ++g_acmil_launches;
    ACMIL_CHECK_CUDA(cudaGetLastError());
  }
  // 9. to_out (+ bias) on the kept rows, + residual (:142-143; transMIL.py:27)
  for (int b = 0; b < s.batch; ++b) {
    acmil_gemm_desc g = gemm0(P);
    g.a = merged + ((size_t)b * np + r0) * inner; g.lda = inner; g.m = nr; g.k = inner; g.batch = 1;
    g.b = w->d_wout; g.ldb = inner; g.n = dim;
    g.bias = w->d_bout;
    const size_t out_rows = s.padded_out ? np : (s.n_out > 0 ? s.n_out : s.n);
    g.c = d_out + (size_t)b * out_rows * dim; g.ldc = dim;
    if (d_residual && !s.padded_out) {
      g.addend = d_residual + (size_t)b * s.n * dim; g.ld_addend = dim; g.beta = 1.f;
    }
    use_split(g, w->d_split_out, dim, 0);
    TM_RUN(tm_gemm(g, st));
  }
  return ACMIL_OK;
}

// ==========================================================================================
// Sequence-parallel NystromAttention (BASELINE.json configs[2]: "bag sharded 1 -> 8 B200").
// The padded sequence is cut at landmark-group boundaries: rank p owns groups [p m / P, (p + 1) m / P) = n_loc = m_loc l
// consecutive rows (rank 0's begin with the front zero padding).  Everything proportional to the sequence length is local;
// what crosses ranks is small and goes through the caller (torch.distributed / peer copies) BETWEEN the phases below:
//   phase A  LN, q / k / v of the local rows, local landmark means        -> all-gather q_l, k_l   [H, m, d]  (1 MB)
//   phase B  attn2 = softmax(q_l k_l^T) of all heads (the start value of the pseudo-inverse needs the maxima over every
//            head), Moore-Penrose iterations for heads [head_first, +count) -> all-gather pinv       [H, m, m]  (2 MB)
//   phase C  local scores q_l k_loc^T, exp against the LOCAL row maximum, (exp v) partial sums + (max, sum) per row
//                                                                          -> all-gather partials   [H, m, d + 2]
//            acmil_lse_merge: kv = sum_p e^(m_p - M) part_p / sum_p e^(m_p - M) l_p  (= attn3 v of the whole sequence)
//   phase D  W = pinv kv, attn1 of the local rows, out = attn1 W + depth-wise conv of v (16-row halo of v from both
//            neighbours, exchanged by the caller into vt_ext) -> to_out (+ bias, + residual) for the local rows
namespace {

struct ShardLayout {
  size_t xn, q, k, a2, y, yt, t1t, t2t, za, zta, zb, ztb, sbuf, wt, merged, split, scal, stats, total;
  int inner, ksplit, n_ext;
};

ShardLayout shard_layout(const acmil_nystrom_shard& s) {
  ShardLayout L{};
  L.inner = s.heads * s.dim_head;
  L.n_ext = s.n_loc + 2 * s.halo;
  const size_t H = s.heads, d = s.dim_head, m = s.num_landmarks, nl = s.n_loc;
  L.ksplit = (int)std::max<size_t>(1, std::min<size_t>(64, (nl / 32) / 48));
  size_t o = 0;
  auto take = [&](size_t n) { const size_t r = o; o += align64(n); return r; };
  L.xn = take(nl * s.dim);
  L.q = take(H * nl * d);
  L.k = take(H * nl * d);
  const size_t mm = H * m * m;
  L.a2 = take(mm); L.y = take(mm); L.yt = take(mm); L.t1t = take(mm); L.t2t = take(mm);
  L.za = take(mm); L.zta = take(mm); L.zb = take(mm); L.ztb = take(mm);
  L.sbuf = take(H * nl * m);
  L.wt = take(H * d * m);
  L.merged = take(nl * (size_t)L.inner);
  L.split = take((size_t)L.ksplit * H * d * m);
  L.scal = take(64);
  L.stats = take(H * nl * (size_t)((m + 31) / 32) * 2);
  L.total = o;
  return L;
}

int shard_check(const acmil_nystrom_shard& s) {
  ACMIL_REQUIRE(s.n_loc >= 1 && s.dim >= 4 && s.dim % 4 == 0 && s.heads >= 1 && s.dim_head >= 4 && s.dim_head % 4 == 0, ACMIL_E_INVALID,
                "nystrom shard: bad shape");
  ACMIL_REQUIRE(s.num_landmarks >= 4 && s.num_landmarks % 4 == 0 && s.num_landmarks <= 1024 && s.m_loc >= 1 &&
                    s.group_len >= 1 && s.n_loc == s.m_loc * s.group_len && s.n_loc % 4 == 0,
                ACMIL_E_INVALID, "nystrom shard: n_loc must be m_loc * group_len and a multiple of 4");
  ACMIL_REQUIRE(s.lead_zero >= 0 && s.lead_zero < s.n_loc && s.n_out >= 0 && s.n_out <= s.n_loc - s.lead_zero, ACMIL_E_INVALID,
                "nystrom shard: bad lead_zero / n_out");
  ACMIL_REQUIRE(s.head_first >= 0 && s.head_count >= 0 && s.head_first + s.head_count <= s.heads, ACMIL_E_INVALID, "nystrom shard: bad head range");
  ACMIL_REQUIRE(!s.residual || (s.conv_kernel % 2 == 1 && s.halo == s.conv_kernel / 2 && s.halo % 4 == 0), ACMIL_E_INVALID,
                "nystrom shard: halo must be conv_kernel / 2 and a multiple of 4");
  return ACMIL_OK;
}

// exp(a - rowmax) in place + (rowmax, rowsum): the local part of a softmax over a sequence that is spread over ranks
__global__ void __launch_bounds__(1024) tm_softmax_partial_kernel(float* __restrict__ a, long long len, float* __restrict__ st_m,
                                                                  float* __restrict__ st_l) {
  __shared__ float red[32];
  __shared__ float bc;
  float* r = a + (size_t)blockIdx.x * len;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -INFINITY;
  for (long long j = tid; j < len; j += 1024) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_max(red[lane]);
    if (lane == 0) bc = t;
  }
  __syncthreads();
  mx = bc;
  float s = 0.f;
  for (long long j = tid; j < len; j += 1024) {
    const float e = expf(r[j] - mx);
    r[j] = e;
    s += e;
  }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_sum(red[lane]);
    if (lane == 0) { st_m[blockIdx.x] = mx; st_l[blockIdx.x] = t; }
  }
}

// kv[h][c][j] = sum_p e^(m_p[h][j] - M) part_p[h][c][j] / sum_p e^(m_p[h][j] - M) l_p[h][j]
__global__ void __launch_bounds__(256) tm_lse_merge_kernel(const float* __restrict__ parts, const float* __restrict__ st_m,
                                                           const float* __restrict__ st_l, int P, int H, int d, int m,
                                                           float* __restrict__ kv) {
  const size_t total = (size_t)H * d * m;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int j = (int)(e % m), h = (int)(e / ((size_t)d * m));
    const size_t row = (size_t)h * m + j;
    float M = -INFINITY;
    for (int p = 0; p < P; ++p) M = fmaxf(M, st_m[(size_t)p * H * m + row]);
    float num = 0.f, den = 0.f;
    for (int p = 0; p < P; ++p) {
      const float w = expf(st_m[(size_t)p * H * m + row] - M);
      num = fmaf(w, parts[(size_t)p * total + e], num);
      den = fmaf(w, st_l[(size_t)p * H * m + row], den);
    }
    kv[e] = num / den;
  }
}

}  // namespace

extern "C" int acmil_nystrom_shard_workspace_bytes(const acmil_nystrom_shard* shard, size_t* bytes) {
  ACMIL_REQUIRE(shard && bytes, ACMIL_E_INVALID, "nystrom shard: null argument");
  const int rc = shard_check(*shard);
  if (rc) return rc;
  *bytes = shard_layout(*shard).total * sizeof(float);
  return ACMIL_OK;
}

extern "C" int acmil_lse_merge(const float* d_parts, const float* d_st_m, const float* d_st_l, int32_t n_ranks, int32_t heads,
                               int32_t dim_head, int32_t num_landmarks, float* d_kv, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_parts && d_st_m && d_st_l && d_kv && n_ranks >= 1 && heads >= 1 && dim_head >= 1 && num_landmarks >= 1, ACMIL_E_INVALID,
                "lse_merge: bad argument");
  const size_t total = (size_t)heads * dim_head * num_landmarks;
  tm_lse_merge_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 1184), 256, 0, (cudaStream_t)stream>>>(
      d_parts, d_st_m, d_st_l, n_ranks, heads, dim_head, num_landmarks, d_kv);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

extern "C" int acmil_nystrom_shard_phase(const acmil_nystrom_shard* shard, const acmil_nystrom_weights* w,
                                         const acmil_nystrom_shard_bufs* bufs, int32_t phase, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(shard && w && bufs && bufs->d_workspace, ACMIL_E_INVALID, "nystrom shard: null argument");
  const acmil_nystrom_shard& s = *shard;
  const acmil_nystrom_shard_bufs& b = *bufs;
  TM_RUN(shard_check(s));
  const ShardLayout L = shard_layout(s);
  ACMIL_REQUIRE(b.workspace_bytes >= L.total * sizeof(float), ACMIL_E_WORKSPACE, "nystrom shard: workspace %zu < %zu bytes",
                b.workspace_bytes, L.total * sizeof(float));
  ACMIL_REQUIRE(((uintptr_t)b.d_workspace & 255) == 0, ACMIL_E_INVALID, "nystrom shard: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = reinterpret_cast<float*>(b.d_workspace);
  const int m = s.num_landmarks, d = s.dim_head, H = s.heads, nl = s.n_loc, inner = L.inner, dim = s.dim, P = s.precise;
  const int n_real = nl - s.lead_zero, n_ext = L.n_ext;
  float *xn = ws + L.xn, *q = ws + L.q, *k = ws + L.k, *sbuf = ws + L.sbuf, *a2 = ws + L.a2;
  const int64_t mm = (int64_t)m * m;
  if (phase == 0) {
    ACMIL_REQUIRE(b.d_x && b.d_ql_loc && b.d_kl_loc && b.d_vt_ext && w->d_wqkv, ACMIL_E_INVALID, "nystrom shard phase A: null buffer");
    if (s.lead_zero) ACMIL_CHECK_CUDA(cudaMemsetAsync(xn, 0, (size_t)s.lead_zero * dim * sizeof(float), st));
    TM_RUN(acmil_layernorm_rows(b.d_x, dim, n_real, dim, w->d_ln_w, w->d_ln_b, w->ln_eps, xn + (size_t)s.lead_zero * dim, dim, stream));
    acmil_gemm_desc g = gemm0(P);
    g.a = xn; g.lda = dim; g.m = nl; g.k = dim; g.batch = 1;
    g.b = w->d_wqkv; g.ldb = dim; g.n = inner;
    g.c = q; g.ldc = d; g.col_block_width = d; g.col_block_stride = (int64_t)nl * d;
    g.alpha = 1.f / sqrtf((float)d);
    use_split(g, w->d_split_qkv, 3 * inner, 0);
    TM_RUN(tm_gemm(g, st));
    g.b = w->d_wqkv + (size_t)inner * dim; g.c = k; g.alpha = 1.f;
    use_split(g, w->d_split_qkv, 3 * inner, inner);
    TM_RUN(tm_gemm(g, st));
    acmil_gemm_desc gv = gemm0(P);      // v^T straight into the halo-extended buffer
    if (P == 2 && w->d_split_qkv) {
      gv.a = xn; gv.lda = dim; gv.m = nl; gv.k = dim; gv.batch = 1;
      gv.b = w->d_wqkv + (size_t)2 * inner * dim; gv.ldb = dim; gv.n = inner;
      gv.ct = b.d_vt_ext + s.halo; gv.ldct = n_ext;
      use_split(gv, w->d_split_qkv, 3 * inner, 2 * inner);
    } else {
      gv.a = w->d_wqkv + (size_t)2 * inner * dim; gv.lda = dim; gv.m = inner; gv.k = dim; gv.batch = 1;
      gv.b = xn; gv.ldb = dim; gv.n = nl;
      gv.c = b.d_vt_ext + s.halo; gv.ldc = n_ext;
    }
    TM_RUN(tm_gemm(gv, st));
    dim3 grid(s.m_loc, H);
    tm_landmark_kernel<<<grid, 256, 0, st>>>(q, b.d_ql_loc, nl, s.m_loc, s.group_len, d);
    tm_landmark_kernel<<<grid, 256, 0, st>>>(k, b.d_kl_loc, nl, s.m_loc, s.group_len, d);
    g_acmil_launches += 2;
    ACMIL_CHECK_CUDA(cudaGetLastError());
    return ACMIL_OK;
  }
  if (phase == 1) {
    ACMIL_REQUIRE(b.d_ql && b.d_kl && b.d_z, ACMIL_E_INVALID, "nystrom shard phase B: null buffer");
    acmil_gemm_desc g = gemm0(P);
    g.a = b.d_ql; g.lda = d; g.a_batch_stride = (int64_t)m * d; g.m = m; g.k = d; g.batch = H;
    g.b = b.d_kl; g.ldb = d; g.b_batch_stride = (int64_t)m * d; g.n = m;
    g.c = a2; g.ldc = m; g.c_batch_stride = mm;
    TM_RUN(tm_gemm(g, st));
    tm_softmax_small_launch(a2, (long long)H * m, m, m, st);
    ++g_acmil_launches;
    if (s.head_count == 0) return ACMIL_OK;
    float *z = ws + L.za, *zt = ws + L.zta, *z2 = ws + L.zb, *zt2 = ws + L.ztb;
    int* scal = reinterpret_cast<int*>(ws + L.scal);
    ACMIL_CHECK_CUDA(cudaMemsetAsync(scal, 0, 8, st));
    tm_pinv_sums_kernel<<<dim3(H, 2), 256, 0, st>>>(a2, m, scal);      // maxima over ALL heads (torch.max of the tensor)
    const int hc = s.head_count;
    const float* a2h = a2 + (size_t)s.head_first * mm;
    tm_pinv_init_kernel<<<dim3((m + 31) / 32, (m + 31) / 32, hc), 256, 0, st>>>(a2h, m, scal, z, zt);
    g_acmil_launches += 2;
    ACMIL_CHECK_CUDA(cudaGetLastError());
    float *y = ws + L.y, *yt = ws + L.yt, *t1t = ws + L.t1t, *t2t = ws + L.t2t;
    auto sq = [&](const float* A, const float* Bt, float* C, float* Ct, float alpha, const float* add, float beta) {
      acmil_gemm_desc g2 = gemm0(P);
      g2.a = A; g2.lda = m; g2.a_batch_stride = mm; g2.m = m; g2.k = m; g2.batch = hc;
      g2.b = Bt; g2.ldb = m; g2.b_batch_stride = mm; g2.n = m;
      g2.c = C; g2.ldc = m; g2.c_batch_stride = mm;
      g2.ct = Ct; g2.ldct = m; g2.ct_batch_stride = mm;
      g2.alpha = alpha; g2.addend = add; g2.ld_addend = m; g2.addend_batch_stride = mm; g2.beta = beta;
      return tm_gemm(g2, st);
    };
    for (int it = 0; it < s.pinv_iterations; ++it) {
      TM_RUN(sq(a2h, zt, y, yt, 1.f, nullptr, 0.f));
      TM_RUN(sq(y, yt, nullptr, t1t, -1.f, y, 7.f));
      TM_RUN(sq(y, t1t, nullptr, t2t, -1.f, y, 15.f));
      TM_RUN(sq(z, t2t, z2, zt2, -0.25f, z, 3.25f));
      std::swap(z, z2);
      std::swap(zt, zt2);
    }
    ACMIL_CHECK_CUDA(cudaMemcpyAsync(b.d_z + (size_t)s.head_first * mm, z, (size_t)hc * mm * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return ACMIL_OK;
  }
  if (phase == 2) {
    ACMIL_REQUIRE(b.d_ql && b.d_kv_part && b.d_st_m && b.d_st_l && b.d_vt_ext, ACMIL_E_INVALID, "nystrom shard phase C: null buffer");
    acmil_gemm_desc g = gemm0(P);
    g.a = b.d_ql; g.lda = d; g.a_batch_stride = (int64_t)m * d; g.m = m; g.k = d; g.batch = H;
    g.b = k; g.ldb = d; g.b_batch_stride = (int64_t)nl * d; g.n = nl;
    g.c = sbuf; g.ldc = nl; g.c_batch_stride = (int64_t)m * nl;
    TM_RUN(tm_gemm(g, st));
    tm_softmax_partial_kernel<<<(unsigned)((size_t)H * m), 1024, 0, st>>>(sbuf, nl, b.d_st_m, b.d_st_l);
    ++g_acmil_launches;
    acmil_gemm_desc g2 = gemm0(P);
    g2.a = b.d_vt_ext + s.halo; g2.lda = n_ext; g2.a_batch_stride = (int64_t)d * n_ext; g2.m = d; g2.k = nl; g2.batch = H;
    g2.b = sbuf; g2.ldb = nl; g2.b_batch_stride = (int64_t)m * nl; g2.n = m;
    g2.c = b.d_kv_part; g2.ldc = m; g2.c_batch_stride = (int64_t)d * m;
    g2.k_split = L.ksplit; g2.split_ws = ws + L.split;
    TM_RUN(tm_gemm(g2, st));
    return ACMIL_OK;
  }
  ACMIL_REQUIRE(phase == 3, ACMIL_E_INVALID, "nystrom shard: phase must be 0..3");
  ACMIL_REQUIRE(b.d_kl && b.d_z && b.d_kv && b.d_vt_ext && b.d_out && w->d_wout && w->d_bout && (!s.residual || w->d_wconv),
                ACMIL_E_INVALID, "nystrom shard phase D: null buffer");
  {
    acmil_gemm_desc g = gemm0(P);      // W^T = (attn3 v)^T pinv^T
    g.a = b.d_kv; g.lda = m; g.a_batch_stride = (int64_t)d * m; g.m = d; g.k = m; g.batch = H;
    g.b = b.d_z; g.ldb = m; g.b_batch_stride = mm; g.n = m;
    g.c = ws + L.wt; g.ldc = m; g.c_batch_stride = (int64_t)d * m;
    TM_RUN(tm_gemm(g, st));
  }
  const int r0 = s.lead_zero;
  const int nr = s.n_out > 0 ? s.n_out : n_real;
  float* merged = ws + L.merged;
  {
    acmil_gemm_desc g = gemm0(P);
    g.a = q + (size_t)r0 * d; g.lda = d; g.a_batch_stride = (int64_t)nl * d; g.m = nr; g.k = d; g.batch = H;
    g.b = b.d_kl; g.ldb = d; g.b_batch_stride = (int64_t)m * d; g.n = m;
    g.c = sbuf; g.ldc = m; g.c_batch_stride = (int64_t)nr * m;
    float* stats = P && chunked_softmax(true) ? ws + L.stats : nullptr;      // as in acmil_nystrom_attn_fwd
    g.softmax_stats_out = stats;
    TM_RUN(tm_gemm(g, st));
    if (!stats) {
      tm_softmax_small_launch(sbuf, (long long)H * nr, m, m, st);
      ++g_acmil_launches;
      ACMIL_CHECK_CUDA(cudaGetLastError());
    }
    acmil_gemm_desc g2 = gemm0(P);
    g2.softmax_stats_in = stats;
    g2.a = sbuf; g2.lda = m; g2.a_batch_stride = (int64_t)nr * m; g2.m = nr; g2.k = m; g2.batch = H;
    g2.b = ws + L.wt; g2.ldb = m; g2.b_batch_stride = (int64_t)d * m; g2.n = d;
    g2.c = merged + (size_t)r0 * inner; g2.ldc = inner; g2.c_batch_stride = d;
    TM_RUN(tm_gemm(g2, st));
  }
  if (s.residual) {
    // the conv kernel indexes merged rows and vt columns with the same i: view merged as starting `halo` rows earlier so
    // that row (halo + r) of the view = local row r = column (halo + r) of vt_ext (the first `halo` view rows are never touched)
    const int ks = s.conv_kernel;
    const size_t smem = (size_t)32 * ((CONV_TP + ks - 1) | 1) * sizeof(float);
    dim3 grid((nr + CONV_TP - 1) / CONV_TP, (inner + 31) / 32, 1);
    tm_resconv_launch(grid, smem, st, b.d_vt_ext, w->d_wconv, merged, n_ext, inner, d, ks, s.halo + r0, nr, s.halo, 0);
    ++g_acmil_launches;
    ACMIL_CHECK_CUDA(cudaGetLastError());
  }
  {
    acmil_gemm_desc g = gemm0(P);
    g.a = merged + (size_t)r0 * inner; g.lda = inner; g.m = nr; g.k = inner; g.batch = 1;
    g.b = w->d_wout; g.ldb = inner; g.n = dim;
    g.bias = w->d_bout;
    g.c = b.d_out; g.ldc = dim;
    if (b.d_residual) { g.addend = b.d_residual; g.ld_addend = dim; g.beta = 1.f; }
    use_split(g, w->d_split_out, dim, 0);
    TM_RUN(tm_gemm(g, st));
  }
  return ACMIL_OK;
}

extern "C" int acmil_softmax_rows_inplace(float* d_a, int64_t ld, int64_t rows, int32_t len, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_a && rows >= 0 && len >= 1 && len <= 1024 && ld >= len, ACMIL_E_INVALID, "softmax_rows_inplace: bad arguments");
  if (rows == 0) return ACMIL_OK;
  tm_softmax_small_launch(d_a, rows, len, ld, (cudaStream_t)stream);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

// ==========================================================================================
// ViT patch encoder (models.py:138-149 -> timm 0.9.2 VisionTransformer.forward)
namespace {

// im2col for the stride = kernel = patch convolution: col[b * np + py * g + px][c * p * p + dy * p + dx]
__global__ void __launch_bounds__(256) tm_im2col_kernel(const float* __restrict__ img, float* __restrict__ col, int B, int C, int S,
                                                        int P) {
  const int g = S / P, kk = C * P * P;
  const size_t total4 = (size_t)B * g * g * kk / 4;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total4; e += (size_t)gridDim.x * 256) {
    const int k = (int)((e * 4) % kk);
    const size_t r = (e * 4) / kk;
    const int px = (int)(r % g), py = (int)((r / g) % g), b = (int)(r / ((size_t)g * g));
    const int c = k / (P * P), dy = (k / P) % P, dx = k % P;
    const float4 v = *reinterpret_cast<const float4*>(img + (((size_t)b * C + c) * S + py * P + dy) * S + px * P + dx);
    *reinterpret_cast<float4*>(col + e * 4) = v;
  }
}

// x[b][0][:] = cls_token + pos_embed[0]
__global__ void tm_cls_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int dim,
                              long long img_stride) {
  for (int j = threadIdx.x; j < dim; j += blockDim.x) x[(size_t)blockIdx.x * img_stride + j] = cls[j] + pos[j];
}

struct VitLayout {
  size_t x, xn, qk, vt, s, merged, hid, stats, total;
  int T, Tp, np;
};

VitLayout vit_layout(const acmil_vit_shape& s) {
  VitLayout L{};
  const int g = s.img / s.patch;
  L.np = g * g;
  L.T = L.np + 1;
  L.Tp = (L.T + 3) & ~3;
  const size_t rows = (size_t)s.batch * L.Tp, dh = s.dim / s.heads;
  size_t o = 0;
  auto take = [&](size_t n) { const size_t r = o; o += align64(n); return r; };
  L.x = take(rows * s.dim);
  L.xn = take(rows * s.dim);
  L.qk = take(2 * rows * s.dim);
  L.vt = take(rows * s.dim);
  L.s = take((size_t)s.heads * s.batch * L.T * L.Tp);
  L.merged = take(rows * s.dim);
  L.hid = take(std::max(rows * s.mlp_dim, (size_t)s.batch * L.np * s.in_ch * s.patch * s.patch));
  (void)dh;
  L.stats = take((size_t)s.heads * s.batch * L.T * (size_t)((L.T + 31) / 32) * 2);      // chunked softmax of the scores
  L.total = o;
  return L;
}

int vit_check(const acmil_vit_shape& s) {
  ACMIL_REQUIRE(s.batch >= 1 && s.img >= s.patch && s.patch >= 4 && s.patch % 4 == 0 && s.img % s.patch == 0 && s.in_ch >= 1,
                ACMIL_E_INVALID, "vit: bad image shape [%d, %d, %d, %d] / patch %d", s.batch, s.in_ch, s.img, s.img, s.patch);
  ACMIL_REQUIRE(s.dim >= 4 && s.heads >= 1 && s.dim % s.heads == 0 && (s.dim / s.heads) % 4 == 0 && s.mlp_dim % 4 == 0 && s.depth >= 1,
                ACMIL_E_INVALID, "vit: bad dims (dim %d, heads %d, mlp %d, depth %d)", s.dim, s.heads, s.mlp_dim, s.depth);
  const int g = s.img / s.patch;
  ACMIL_REQUIRE(g * g + 1 <= 1024, ACMIL_E_INVALID, "vit: %d tokens exceed the 1024-token softmax kernel", g * g + 1);
  ACMIL_REQUIRE((long long)s.batch * s.heads <= 65535, ACMIL_E_INVALID, "vit: batch * heads too large (%d x %d)", s.batch, s.heads);
  return ACMIL_OK;
}

}  // namespace

extern "C" int acmil_vit_workspace_bytes(const acmil_vit_shape* shape, size_t* bytes) {
  ACMIL_REQUIRE(shape && bytes, ACMIL_E_INVALID, "vit: null argument");
  TM_RUN(vit_check(*shape));
  *bytes = vit_layout(*shape).total * sizeof(float);
  return ACMIL_OK;
}

extern "C" int acmil_vit_fwd(const acmil_vit_shape* shape, const acmil_vit_weights* w, const float* d_images, float* d_features,
                             float* d_logits, void* d_workspace, size_t workspace_bytes, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(shape && w && d_images && d_features && d_workspace && w->blocks, ACMIL_E_INVALID, "vit: null argument");
  const acmil_vit_shape& s = *shape;
  TM_RUN(vit_check(s));
  ACMIL_REQUIRE(w->d_cls_token && w->d_pos_embed && w->d_patch_w && w->d_patch_b && w->d_norm_w && w->d_norm_b, ACMIL_E_INVALID,
                "vit: null weight");
  ACMIL_REQUIRE(s.n_class == 0 || !d_logits || (w->d_head_w && w->d_head_b), ACMIL_E_INVALID, "vit: logits requested without a head");
  ACMIL_REQUIRE(((uintptr_t)d_workspace & 255) == 0 && ((uintptr_t)d_images & 15) == 0, ACMIL_E_INVALID, "vit: unaligned buffers");
  const VitLayout L = vit_layout(s);
  ACMIL_REQUIRE(workspace_bytes >= L.total * sizeof(float), ACMIL_E_WORKSPACE, "vit: workspace %zu < %zu bytes", workspace_bytes,
                L.total * sizeof(float));
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = reinterpret_cast<float*>(d_workspace);
  const int B = s.batch, D = s.dim, Hh = s.heads, dh = D / Hh, T = L.T, Tp = L.Tp, P = s.precise;
  const int64_t rows = (int64_t)B * Tp;
  float *x = ws + L.x, *xn = ws + L.xn, *qk = ws + L.qk, *vt = ws + L.vt, *S = ws + L.s, *merged = ws + L.merged, *hid = ws + L.hid;
  ACMIL_CHECK_CUDA(cudaMemsetAsync(x, 0, (size_t)rows * D * sizeof(float), st));            // pad tokens stay finite
  ACMIL_CHECK_CUDA(cudaMemsetAsync(merged, 0, (size_t)rows * D * sizeof(float), st));

  // patch embedding: Conv2d(in_ch, dim, patch, stride = patch) as im2col + GEMM, + position embedding
  const int kk = s.in_ch * s.patch * s.patch;
  {
    const size_t total4 = (size_t)B * L.np * kk / 4;
    tm_im2col_kernel<<<(unsigned)std::min<size_t>((total4 + 255) / 256, 148 * 16), 256, 0, st>>>(d_images, hid, B, s.in_ch, s.img, s.patch);
    tm_cls_kernel<<<B, 128, 0, st>>>(x, w->d_cls_token, w->d_pos_embed, D, (long long)Tp * D);
    g_acmil_launches += 2;
    ACMIL_CHECK_CUDA(cudaGetLastError());
    acmil_gemm_desc g = gemm0(P);
    g.a = hid; g.lda = kk; g.a_batch_stride = (int64_t)L.np * kk; g.m = L.np; g.k = kk; g.batch = B;
    g.b = w->d_patch_w; g.ldb = kk; g.n = D;
    g.bias = w->d_patch_b;
    g.addend = w->d_pos_embed + D; g.ld_addend = D; g.beta = 1.f;
    g.c = x + D; g.ldc = D; g.c_batch_stride = (int64_t)Tp * D;
    use_split(g, w->d_split_patch, D, 0);
    TM_RUN(tm_gemm(g, st));
  }
  const int64_t head_sz = rows * dh;      // one head of q (or k): [B * Tp][dh]
  for (int i = 0; i < s.depth; ++i) {
    const acmil_vit_block_weights& bw = w->blocks[i];
    ACMIL_REQUIRE(bw.d_ln1_w && bw.d_ln1_b && bw.d_qkv_w && bw.d_qkv_b && bw.d_proj_w && bw.d_proj_b && bw.d_ln2_w && bw.d_ln2_b &&
                      bw.d_fc1_w && bw.d_fc1_b && bw.d_fc2_w && bw.d_fc2_b,
                  ACMIL_E_INVALID, "vit: null weight in block %d", i);
    // x = x + proj(softmax(q k^T / sqrt(dh)) v),  q, k, v = qkv(LN(x))
    TM_RUN(acmil_layernorm_rows(x, D, rows, D, bw.d_ln1_w, bw.d_ln1_b, s.ln_eps, xn, D, stream));
    {
      acmil_gemm_desc g = gemm0(P);      // q, k head-major: [which * heads + head][B * Tp][dh]
      g.a = xn; g.lda = D; g.m = (int)rows; g.k = D; g.batch = 1;
      g.b = bw.d_qkv_w; g.ldb = D; g.n = 2 * D; g.bias = bw.d_qkv_b;
      g.c = qk; g.ldc = dh; g.col_block_width = dh; g.col_block_stride = head_sz;
      use_split(g, bw.d_split_qkv, 3 * D, 0);
      TM_RUN(tm_gemm(g, st));
      acmil_gemm_desc gv = gemm0(P);     // v transposed: [dim][B * Tp]
      if (P == 2 && bw.d_split_qkv) {    // v = xn Wv^T + b on the fp16-split kernel, stored transposed by the epilogue
        gv.a = xn; gv.lda = D; gv.m = (int)rows; gv.k = D; gv.batch = 1;
        gv.b = bw.d_qkv_w + (size_t)2 * D * D; gv.ldb = D; gv.n = D;
        gv.bias = bw.d_qkv_b + 2 * D;
        gv.ct = vt; gv.ldct = rows;
        use_split(gv, bw.d_split_qkv, 3 * D, 2 * D);
      } else {
        gv.a = bw.d_qkv_w + (size_t)2 * D * D; gv.lda = D; gv.m = D; gv.k = D; gv.batch = 1;
        gv.b = xn; gv.ldb = D; gv.n = (int)rows;
        gv.bias = bw.d_qkv_b + 2 * D; gv.bias_per_row = 1;
        gv.c = vt; gv.ldc = rows;
      }
      TM_RUN(tm_gemm(gv, st));
    }
    {
      acmil_gemm_desc g = gemm0(P);      // scores, batch z = head * B + image
      g.batch = Hh * B; g.batch_inner = B;
      g.a = qk; g.lda = dh; g.a_batch_stride = (int64_t)Tp * dh; g.a_batch_stride2 = head_sz; g.m = T; g.k = dh;
      g.b = qk + (size_t)Hh * head_sz; g.ldb = dh; g.b_batch_stride = (int64_t)Tp * dh; g.b_batch_stride2 = head_sz; g.n = T;
      g.c = S; g.ldc = Tp; g.c_batch_stride = (int64_t)T * Tp; g.c_batch_stride2 = (int64_t)B * T * Tp;
      g.alpha = 1.f / sqrtf((float)dh);
      float* stats = P && chunked_softmax(true) ? ws + L.stats : nullptr;      // chunked softmax instead of the softmax pass
      g.softmax_stats_out = stats;
      TM_RUN(tm_gemm(g, st));
      if (!stats) TM_RUN(acmil_softmax_rows_inplace(S, Tp, (int64_t)Hh * B * T, T, stream));
      acmil_gemm_desc g2 = gemm0(P);     // attn @ v -> heads merged as [B * Tp][dim]
      g2.softmax_stats_in = stats;
      g2.batch = Hh * B; g2.batch_inner = B;
      g2.a = S; g2.lda = Tp; g2.a_batch_stride = (int64_t)T * Tp; g2.a_batch_stride2 = (int64_t)B * T * Tp; g2.m = T; g2.k = T;
      g2.b = vt; g2.ldb = rows; g2.b_batch_stride = Tp; g2.b_batch_stride2 = (int64_t)dh * rows; g2.n = dh;
      g2.c = merged; g2.ldc = D; g2.c_batch_stride = (int64_t)Tp * D; g2.c_batch_stride2 = dh;
      TM_RUN(tm_gemm(g2, st));
      acmil_gemm_desc g3 = gemm0(P);     // proj + residual, in place
      g3.a = merged; g3.lda = D; g3.m = (int)rows; g3.k = D; g3.batch = 1;
      g3.b = bw.d_proj_w; g3.ldb = D; g3.n = D; g3.bias = bw.d_proj_b;
      g3.addend = x; g3.ld_addend = D; g3.beta = 1.f;
      g3.c = x; g3.ldc = D;
      use_split(g3, bw.d_split_proj, D, 0);
      TM_RUN(tm_gemm(g3, st));
    }
    // x = x + fc2(gelu(fc1(LN(x))))
    TM_RUN(acmil_layernorm_rows(x, D, rows, D, bw.d_ln2_w, bw.d_ln2_b, s.ln_eps, xn, D, stream));
    {
      acmil_gemm_desc g = gemm0(P);
      g.a = xn; g.lda = D; g.m = (int)rows; g.k = D; g.batch = 1;
      g.b = bw.d_fc1_w; g.ldb = D; g.n = s.mlp_dim; g.bias = bw.d_fc1_b; g.act = 2;
      g.c = hid; g.ldc = s.mlp_dim;
      use_split(g, bw.d_split_fc1, s.mlp_dim, 0);
      TM_RUN(tm_gemm(g, st));
      acmil_gemm_desc g2 = gemm0(P);
      g2.a = hid; g2.lda = s.mlp_dim; g2.m = (int)rows; g2.k = s.mlp_dim; g2.batch = 1;
      g2.b = bw.d_fc2_w; g2.ldb = s.mlp_dim; g2.n = D; g2.bias = bw.d_fc2_b;
      g2.addend = x; g2.ld_addend = D; g2.beta = 1.f;
      g2.c = x; g2.ldc = D;
      use_split(g2, bw.d_split_fc2, D, 0);
      TM_RUN(tm_gemm(g2, st));
    }
  }
  // final norm, class token (global_pool = 'token'), optional CustomModel.head
  TM_RUN(acmil_layernorm_rows(x, (int64_t)Tp * D, B, D, w->d_norm_w, w->d_norm_b, s.ln_eps, d_features, D, stream));
  if (d_logits && s.n_class > 0) {
    acmil_gemm_desc g = gemm0(P);
    g.a = d_features; g.lda = D; g.m = B; g.k = D; g.batch = 1;
    g.b = w->d_head_w; g.ldb = D; g.n = s.n_class; g.bias = w->d_head_b;
    g.c = d_logits; g.ldc = s.n_class;
    TM_RUN(tm_gemm(g, st));
  }
  return ACMIL_OK;
}
