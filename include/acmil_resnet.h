/*
 * acmil_resnet -- C-ABI of the convolutional building blocks of the ResNet18 patch encoder
 * (models.py:13-77, dispatch models.py:201-204 -- SURVEY section 8a row a13); same library and
 * conventions as acmil_b200.h.  A convolution is  im2col (this file) -> acmil_gemm_nt (3xTF32 on
 * tcgen05, acmil_transmil.h) whose epilogue applies the folded BatchNorm bias, the residual add and
 * the ReLU of BasicBlock.forward.  Activations are NHWC fp32 between layers: the GEMM output
 * [B * Ho * Wo, Cout] IS the next layer's input.
 */
#ifndef ACMIL_RESNET_H
#define ACMIL_RESNET_H

#include "acmil_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* col[(b * Ho + oy) * Wo + ox][(ky * kw + kx) * C + c] = x[b][oy * stride - pad + ky][ox * stride - pad + kx][c]
 * (0 outside the image, 0 in the k_pad - kh * kw * C padding columns).  nchw_in != 0: x is [B, C, H, W]
 * (the image batch of nn.Conv2d conv1, models.py:17), else [B, H, W, C].  k_pad % 4 == 0. */
ACMIL_API int acmil_im2col(const float* d_x, float* d_col, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t kh,
                           int32_t kw, int32_t stride, int32_t pad, int32_t k_pad, int32_t nchw_in, void* stream);

/* nn.MaxPool2d(kernel, stride, padding) on NHWC (models.py:21). */
ACMIL_API int acmil_maxpool_nhwc(const float* d_x, float* d_y, int32_t batch, int32_t h, int32_t w, int32_t c,
                                 int32_t kernel, int32_t stride, int32_t pad, void* stream);

/* nn.AdaptiveAvgPool2d(1) + flatten on NHWC (models.py:26, 72-73): y[b][c] = mean over the hw positions. */
ACMIL_API int acmil_avgpool_nhwc(const float* d_x, float* d_y, int32_t batch, int32_t hw, int32_t c, void* stream);

#ifdef __cplusplus
}
#endif
#endif
