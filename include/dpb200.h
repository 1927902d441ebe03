/*
 * dpb200 — C ABI of the B200-native DensePose R-CNN forward pass.
 *
 * The reference (dajes/DensePose-TorchScript) is pure Python and has no FFI of its own; every
 * entry point below names the reference op-level seam (file:line under /root/reference) it
 * replaces.  Conventions for all entries:
 *   - plain C linkage, plain pointers and sizes, no torch / C++ types;
 *   - every pointer is DEVICE memory unless the field says "host";
 *   - the caller owns every buffer; nothing is allocated, nothing is synchronised:
 *     all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, negative on error; dpb200_last_error() (thread local) explains;
 *   - activations are bf16 NHWC unless stated; boxes / scores / public outputs are fp32.
 */
#ifndef DPB200_H_
#define DPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPB200_ABI_VERSION 1

const char* dpb200_last_error(void);
int dpb200_abi_version(void);
/* 1 when a CUDA device of compute capability 10.x is current, else 0 (never falls back). */
int dpb200_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * conv2d / linear / deconv-phase as a tcgen05 implicit GEMM.
 * Replaces F.conv2d in Conv2d.forward (detectron2/layers/wrappers.py:104-112) with FrozenBN
 * (detectron2/layers/batch_norm.py:54-62) pre-folded into w/bias, nn.Linear in
 * FastRCNNConvFCHead.forward (detectron2/modeling/roi_heads/box_head.py:95-98) and
 * FastRCNNOutputLayers.forward (fast_rcnn.py:238-257), and one output-parity phase of
 * ConvTranspose2d(k=4,s=2,p=1) in DensePoseChartPredictor (densepose/modeling/predictors/chart.py:45-59).
 *
 *   y[n,oy,ox,co] = act( sum x[n, oy*sy+ky*dil-pad_y, ox*sx+kx*dil-pad_x, ci] * w[co,(ky*kw+kx)*cin_pad+ci]
 *                        + bias[co] + res[n, oy>>res_shift, ox>>res_shift, co] )
 */
typedef struct dpb200_conv2d_args {
  const void* x;            /* bf16, element strides below, channel stride 1                    */
  int32_t n, h, w, cin;     /* input extents                                                    */
  int64_t x_sn, x_sh, x_sw; /* input strides in elements (0 = dense NHWC)                       */
  const void* wgt;          /* bf16 packed [cout_pad][kh*kw*cin_pad] (see dpb200_pack_conv_weight) */
  int32_t cin_pad, cout_pad;
  const float* bias;        /* fp32 [cout_pad] or NULL                                          */
  int32_t kh, kw, sy, sx, pad_y, pad_x, dil;
  int32_t h_out, w_out;
  int32_t relu;
  const void* res;          /* bf16 residual or NULL                                            */
  int64_t res_sn, res_sy, res_sx;
  int32_t res_shift;        /* 1: nearest-x2 upsampled residual (FPN top-down, fpn.py:150-154)   */
  void* y;                  /* bf16 or fp32                                                     */
  int32_t y_fp32;
  int64_t y_sn, y_sy, y_sx; /* output strides in elements (0 = dense NHWC with cout_pad channels) */
  const int32_t* n_valid;   /* optional device scalar: only the first *n_valid images are computed */
  int32_t block_n, stages;  /* 0 = automatic                                                    */
  int32_t tiled;            /* 0: A operand via im2col-mode TMA; 1: via tiled 4-D boxes         */
} dpb200_conv2d_args;

int dpb200_conv2d(const dpb200_conv2d_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPB200_H_ */
