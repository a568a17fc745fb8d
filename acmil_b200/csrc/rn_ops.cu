// rn_ops.cu -- data-movement kernels of the ResNet18 patch encoder (models.py:13-77): im2col gather for
// nn.Conv2d (the contraction itself runs on the tcgen05 GEMM engine, tm_gemm.cu), MaxPool2d and the global
// average pool, all on NHWC fp32 activations.  HBM-bound gathers: one float4 (4 channels) per thread where
// the channel count allows, coalesced along the column index of the im2col matrix.
#include <algorithm>

#include "acmil_resnet.h"
#include "gp_common.cuh"

namespace {

template <bool NCHW, bool VEC4>
__global__ void __launch_bounds__(256) rn_im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int H, int W,
                                                        int C, int kh, int kw, int stride, int pad, int Ho, int Wo, int kpad) {
  constexpr int V = VEC4 ? 4 : 1;
  const int kq = kpad / V;                                  // column groups per row
  const size_t total = (size_t)B * Ho * Wo * kq;
  const int kreal = kh * kw * C;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int kk = (int)(e % kq) * V;
    const size_t r = e / kq;
    const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), b = (int)(r / ((size_t)Wo * Ho));
    float v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.f;
    if (kk < kreal) {
      const int c = kk % C, t = kk / C, kx = t % kw, ky = t / kw;
      const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        if constexpr (VEC4) {          // C % 4 == 0: the 4 columns are 4 consecutive channels of one pixel
          const float4 q = *reinterpret_cast<const float4*>(x + (((size_t)b * H + iy) * W + ix) * C + c);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else if constexpr (NCHW) {
          v[0] = x[(((size_t)b * C + c) * H + iy) * W + ix];
        } else {
          v[0] = x[(((size_t)b * H + iy) * W + ix) * C + c];
        }
      }
    }
    if constexpr (VEC4) *reinterpret_cast<float4*>(col + e * 4) = make_float4(v[0], v[1], v[2], v[3]);
    else col[e] = v[0];
  }
}

__global__ void __launch_bounds__(256) rn_maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C,
                                                         int k, int stride, int pad, int Ho, int Wo) {
  const size_t total = (size_t)B * Ho * Wo * C;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int c = (int)(e % C);
    const size_t r = e / C;
    const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), b = (int)(r / ((size_t)Wo * Ho));
    float m = -INFINITY;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= W) continue;
        m = fmaxf(m, x[(((size_t)b * H + iy) * W + ix) * C + c]);
      }
    }
    y[e] = m;
  }
}

// one CTA per image, thread = channel (strided), sequential sum over the positions in index order
__global__ void __launch_bounds__(256) rn_avgpool_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C) {
  const float* xb = x + (size_t)blockIdx.x * HW * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < HW; ++i) s += xb[(size_t)i * C + c];
    y[(size_t)blockIdx.x * C + c] = s / (float)HW;
  }
}

unsigned grid_for(size_t total) { return (unsigned)std::min<size_t>((total + 255) / 256, (size_t)148 * 16); }

}  // namespace

extern "C" int acmil_im2col(const float* d_x, float* d_col, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t kh, int32_t kw,
                            int32_t stride, int32_t pad, int32_t k_pad, int32_t nchw_in, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_x && d_col && batch > 0 && h > 0 && w > 0 && c > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0,
                ACMIL_E_INVALID, "im2col: bad arguments");
  ACMIL_REQUIRE(k_pad % 4 == 0 && k_pad >= kh * kw * c, ACMIL_E_INVALID, "im2col: k_pad %d must be a multiple of 4 >= %d", k_pad,
                kh * kw * c);
  const int ho = (h + 2 * pad - kh) / stride + 1, wo = (w + 2 * pad - kw) / stride + 1;
  ACMIL_REQUIRE(ho > 0 && wo > 0, ACMIL_E_INVALID, "im2col: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = !nchw_in && c % 4 == 0 && ((uintptr_t)d_x & 15) == 0 && ((uintptr_t)d_col & 15) == 0;
  const size_t total = (size_t)batch * ho * wo * (vec ? k_pad / 4 : k_pad);
  if (vec) rn_im2col_kernel<false, true><<<grid_for(total), 256, 0, st>>>(d_x, d_col, batch, h, w, c, kh, kw, stride, pad, ho, wo, k_pad);
  else if (nchw_in) rn_im2col_kernel<true, false><<<grid_for(total), 256, 0, st>>>(d_x, d_col, batch, h, w, c, kh, kw, stride, pad, ho, wo, k_pad);
  else rn_im2col_kernel<false, false><<<grid_for(total), 256, 0, st>>>(d_x, d_col, batch, h, w, c, kh, kw, stride, pad, ho, wo, k_pad);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

extern "C" int acmil_maxpool_nhwc(const float* d_x, float* d_y, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t kernel,
                                  int32_t stride, int32_t pad, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_x && d_y && batch > 0 && h > 0 && w > 0 && c > 0 && kernel > 0 && stride > 0 && pad >= 0 && 2 * pad <= kernel,
                ACMIL_E_INVALID, "maxpool: bad arguments");
  const int ho = (h + 2 * pad - kernel) / stride + 1, wo = (w + 2 * pad - kernel) / stride + 1;
  ACMIL_REQUIRE(ho > 0 && wo > 0, ACMIL_E_INVALID, "maxpool: empty output");
  rn_maxpool_kernel<<<grid_for((size_t)batch * ho * wo * c), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, batch, h, w, c, kernel, stride,
                                                                                            pad, ho, wo);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}

extern "C" int acmil_avgpool_nhwc(const float* d_x, float* d_y, int32_t batch, int32_t hw, int32_t c, void* stream) {
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  ACMIL_REQUIRE(d_x && d_y && batch > 0 && hw > 0 && c > 0, ACMIL_E_INVALID, "avgpool: bad arguments");
  rn_avgpool_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(d_x, d_y, hw, c);
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  return ACMIL_OK;
}
