// pre_u8.cu -- patch preprocessing of the feature-extraction loop (Step2_feature_extract.py:35-71 with
// datasets/dataset_h5.py:20-37, 207-230): PIL RGB patch -> transforms.Resize(out) (Pillow's antialiased BILINEAR resample
// for 8-bit images) -> ToTensor (/255) -> Normalize(mean, std), as ONE kernel from uint8 HWC patches to the fp32 CHW
// batch the encoder reads, plus the fp16 cast of the feature store (Step2_feature_extract.py:165).
//
// The resample follows Pillow's two-pass 8-bit algorithm exactly (src/libImaging/Resample.c: triangle filter stretched
// by the scale factor, per-output-pixel windows, coefficients normalised then rounded to 22-bit fixed point, horizontal
// pass rounded to uint8 before the vertical pass), so the bytes are identical to what the reference's CPU workers
// produce; the coefficient tables are computed on the host in double precision like Pillow does.
#include <cuda_fp16.h>

#include <cmath>
#include <vector>

#include "acmil_transmil.h"
#include "gp_common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;
constexpr int TILE_Y = 16;      // output rows per CTA

struct Coeffs {
  int ksize;
  std::vector<int> bounds;      // [out][2]: first input index, count
  std::vector<int> kk;          // [out][ksize] fixed point
};

Coeffs precompute(int in_size, int out_size) {
  Coeffs c;
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;      // bilinear: support 1
  c.ksize = (int)std::ceil(support) * 2 + 1;
  c.bounds.assign((size_t)out_size * 2, 0);
  c.kk.assign((size_t)out_size * c.ksize, 0);
  std::vector<double> k(c.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double t = (x + xmin - center + 0.5) * ss;
      if (t < 0.0) t = -t;
      const double w = t < 1.0 ? 1.0 - t : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? k[x] / ww : k[x];
      c.kk[(size_t)xx * c.ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << PRECISION_BITS)) : (int)(0.5 + v * (1 << PRECISION_BITS));
    }
    c.bounds[xx * 2] = xmin;
    c.bounds[xx * 2 + 1] = xmax;
  }
  return c;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

struct PreParams {
  const unsigned char* img;      // [B][H][W][3]
  float* out;                    // [B][3][oh][ow]
  const int* tab;                // hb[ow][2] hk[ow][ks] vb[oh][2] vk[oh][ks]
  int H, W, oh, ow, ksh, ksv;
  float mean[3], inv_unused[3], std_[3];
  int max_rows;                  // input rows a tile can need
};

__global__ void __launch_bounds__(256) pre_u8_kernel(const __grid_constant__ PreParams p) {
  extern __shared__ unsigned char sm[];
  const int* hb = p.tab;
  const int* hk = hb + p.ow * 2;
  const int* vb = hk + p.ow * p.ksh;
  const int* vk = vb + p.oh * 2;
  const int y0 = blockIdx.x * TILE_Y, y1 = min(p.oh, y0 + TILE_Y), b = blockIdx.y;
  const int r0 = vb[y0 * 2];
  const int r1 = vb[(y1 - 1) * 2] + vb[(y1 - 1) * 2 + 1];      // input rows [r0, r1)
  const int nr = r1 - r0;
  unsigned char* tmp = sm;                                     // [nr][ow][3] after the horizontal pass
  const unsigned char* src = p.img + ((size_t)b * p.H + r0) * p.W * 3;
  // horizontal pass, rounded to uint8 like ImagingResampleHorizontal_8bpc
  for (int e = threadIdx.x; e < nr * p.ow; e += 256) {
    const int r = e / p.ow, xx = e % p.ow;
    const int xmin = hb[xx * 2], cnt = hb[xx * 2 + 1];
    const int* k = hk + xx * p.ksh;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    const unsigned char* px = src + ((size_t)r * p.W + xmin) * 3;
    for (int x = 0; x < cnt; ++x) {
      const int kv = k[x];
      s0 += px[x * 3] * kv;
      s1 += px[x * 3 + 1] * kv;
      s2 += px[x * 3 + 2] * kv;
    }
    unsigned char* t = tmp + ((size_t)r * p.ow + xx) * 3;
    t[0] = (unsigned char)clip8(s0);
    t[1] = (unsigned char)clip8(s1);
    t[2] = (unsigned char)clip8(s2);
  }
  __syncthreads();
  // vertical pass + ToTensor + Normalize: thread = (row, x), writes are contiguous in x per channel plane
  const size_t plane = (size_t)p.oh * p.ow;
  for (int e = threadIdx.x; e < (y1 - y0) * p.ow; e += 256) {
    const int yy = y0 + e / p.ow, xx = e % p.ow;
    const int ymin = vb[yy * 2] - r0, cnt = vb[yy * 2 + 1];
    const int* k = vk + yy * p.ksv;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < cnt; ++y) {
      const unsigned char* t = tmp + ((size_t)(ymin + y) * p.ow + xx) * 3;
      const int kv = k[y];
      s0 += t[0] * kv;
      s1 += t[1] * kv;
      s2 += t[2] * kv;
    }
    float* o = p.out + (size_t)b * 3 * plane + (size_t)yy * p.ow + xx;
    o[0] = __fdiv_rn(__fdiv_rn((float)clip8(s0), 255.f) - p.mean[0], p.std_[0]);
    o[plane] = __fdiv_rn(__fdiv_rn((float)clip8(s1), 255.f) - p.mean[1], p.std_[1]);
    o[2 * plane] = __fdiv_rn(__fdiv_rn((float)clip8(s2), 255.f) - p.mean[2], p.std_[2]);
  }
}

__global__ void __launch_bounds__(256) f32_to_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) y[i] = __float2half_rn(x[i]);
}

size_t table_ints(int ow, int oh, int ksh, int ksv) { return (size_t)ow * (2 + ksh) + (size_t)oh * (2 + ksv); }

}  // namespace

extern "C" int acmil_preprocess_workspace_bytes(int32_t in_h, int32_t in_w, int32_t out_size, size_t* bytes) {
  ACMIL_REQUIRE(bytes && in_h >= 1 && in_w >= 1 && out_size >= 1, ACMIL_E_INVALID, "preprocess: bad shape");
  const Coeffs h = precompute(in_w, out_size), v = precompute(in_h, out_size);
  *bytes = table_ints(out_size, out_size, h.ksize, v.ksize) * sizeof(int);
  return ACMIL_OK;
}

extern "C" int acmil_preprocess_u8(const uint8_t* d_img, int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size,
                                   const float* mean3, const float* std3, float* d_out, void* d_workspace, size_t workspace_bytes,
                                   void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_img && d_out && d_workspace && mean3 && std3 && batch >= 1 && in_h >= 1 && in_w >= 1 && out_size >= 1 && batch <= 65535,
                ACMIL_E_INVALID, "preprocess: bad arguments");
  const Coeffs h = precompute(in_w, out_size), v = precompute(in_h, out_size);
  const size_t ints = table_ints(out_size, out_size, h.ksize, v.ksize);
  ACMIL_REQUIRE(workspace_bytes >= ints * sizeof(int), ACMIL_E_WORKSPACE, "preprocess: workspace too small");
  std::vector<int> tab;
  tab.reserve(ints);
  tab.insert(tab.end(), h.bounds.begin(), h.bounds.end());
  tab.insert(tab.end(), h.kk.begin(), h.kk.end());
  tab.insert(tab.end(), v.bounds.begin(), v.bounds.end());
  tab.insert(tab.end(), v.kk.begin(), v.kk.end());
  cudaStream_t st = (cudaStream_t)stream;
  ACMIL_CHECK_CUDA(cudaMemcpyAsync(d_workspace, tab.data(), ints * sizeof(int), cudaMemcpyHostToDevice, st));   // pageable: staged before return
  PreParams p{};
  p.img = d_img; p.out = d_out; p.tab = reinterpret_cast<const int*>(d_workspace);
  p.H = in_h; p.W = in_w; p.oh = out_size; p.ow = out_size; p.ksh = h.ksize; p.ksv = v.ksize;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.std_[c] = std3[c]; }
  int max_rows = 0;
  for (int y0 = 0; y0 < out_size; y0 += TILE_Y) {
    const int y1 = std::min(out_size, y0 + TILE_Y);
    max_rows = std::max(max_rows, v.bounds[(y1 - 1) * 2] + v.bounds[(y1 - 1) * 2 + 1] - v.bounds[y0 * 2]);
  }
  const size_t smem = (size_t)max_rows * out_size * 3;
  ACMIL_REQUIRE(smem <= 200 * 1024, ACMIL_E_UNSUPPORTED, "preprocess: a %d-row tile needs %zu bytes of shared memory", TILE_Y, smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    ACMIL_CHECK_CUDA(cudaFuncSetAttribute(pre_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  pre_u8_kernel<<<dim3((out_size + TILE_Y - 1) / TILE_Y, batch), 256, smem, st>>>(p);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

extern "C" int acmil_f32_to_f16(const float* d_x, void* d_y, int64_t n, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_x && d_y && n >= 0, ACMIL_E_INVALID, "f32_to_f16: bad arguments");
  if (n == 0) return ACMIL_OK;
  f32_to_f16_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(d_x, (__half*)d_y, (size_t)n);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}
