// Internal launch interface of the HBM-bound stage kernels (everything that is not a GEMM).
// All functions enqueue on `stream`, allocate nothing and return 0 / negative with dpb::set_error.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpb {

typedef __nv_bfloat16 bf16;

// ---- preprocess (defaults.py:76-89 + rcnn.py:156-181) ------------------------------------------
// src: [B, H0, W0, 3] fp32 or u8 (HWC). dst: space-to-depth stem layout [B, Hp/2, Wx, 16] bf16: the padded image
// pixel (y, x, c) sits at [y/2][x/2 + 2][((y&1)*2 + (x&1))*4 + c]; Wx = Wp/2 + 4 (two zero pixels on either side);
// everything outside the resized Hr x Wr image and channel 3 of every sub-pixel is zero.
struct PreprocessArgs {
  const void* src; int src_u8; int B, H0, W0;
  int Hr, Wr;            // resized extents floor(H0*k), floor(W0*k)
  float inv_scale;       // (float)(1.0 / k)
  int flip_rgb;          // 1: swap channel 0 and 2 (INPUT.FORMAT == RGB and bgr input)
  float mean[3], std[3];
  bf16* dst; int Hp, Wx;
  const int2* tables;    // uint8 input: fixed-point resize tables [1 + Hr + Wr] built by launch_u8_resize_tables
  int variant;           // float input: 0 = ATen's separable kernel (multi-threaded reference), 1 = its channels-last
                         // kernel (what a single-threaded reference runs for a 3-channel image)
  bf16* dst_lo;          // strict mode: low half of the bf16 hi/lo split (else null)
};
int launch_preprocess(const PreprocessArgs& a, cudaStream_t s);
// ATen's uint8 bilinear weights (int16 fixed point, double-precision centres) for both axes: tab[1 + Hr + Wr] int2.
int launch_u8_resize_tables(int2* tab, int H0, int Hr, int W0, int Wr, double scale, cudaStream_t s);

// ---- 3x3 stride-2 pad-1 max pool, NHWC bf16 (resnet.py:353) -------------------------------------
int launch_maxpool3x3s2(const bf16* x, bf16* y, int B, int H, int W, int C, cudaStream_t s);

// ---- bilinear x2 (align_corners=False), NHWC bf16 (roi_head.py:63) ------------------------------
int launch_upsample2x(const bf16* x, bf16* y, int B, int H, int W, int C, cudaStream_t s);
// out = a + up2(b) + up2(c) + up2(d)   (a: [B,H,W,C]; b,c,d: [B,H/2,W/2,C])  (roi_head.py:71-77)
int launch_decoder_merge(const bf16* a, const bf16* b, const bf16* c, const bf16* d, bf16* out, int B,
                         int H, int W, int C, cudaStream_t s);

// ---- RPN (rpn.py:319-394, proposal_utils.py:19-134) ---------------------------------------------
struct RpnLevel { const float* head; int H, W; float stride; float anchors[12]; };
struct RpnArgs {
  RpnLevel lvl[5];       // head: [B, H, W, 16] fp32 = 3 objectness logits, 12 deltas (a*4+d), 1 pad
  int B;
  int pre_topk, post_topk;
  float nms_thresh;
  float clip_x, clip_y;  // x is clamped to [0, clip_x], y to [0, clip_y]  (quirk 1: = H_pad, W_pad)
  // workspace
  float* cand_boxes;     // [B, 5, pre_topk, 4]
  float* cand_scores;    // [B, 5, pre_topk]
  int* cand_count;       // [B, 5]
  unsigned char* cand_keep;  // [B, 5, pre_topk]
  // outputs
  float* prop_boxes;     // [B, post_topk, 4]
  float* prop_scores;    // [B, post_topk]
  int* prop_count;       // [B]
};
int launch_rpn_topk_decode(const RpnArgs& a, cudaStream_t s);
int launch_rpn_nms(const RpnArgs& a, cudaStream_t s);
int launch_rpn_merge(const RpnArgs& a, cudaStream_t s);

// Generic greedy NMS on already score-sorted boxes (torchvision nms semantics); test entry.
int launch_nms_sorted(const float* boxes, int n, float thr, unsigned char* keep, cudaStream_t s);

// ---- ROIAlign (poolers.py:187-227 + torchvision roi_align, aligned=False, sampling 2) ------------
struct RoiAlignArgs {
  const bf16* feat[4]; int H[4], W[4]; float scale[4]; int n_levels;   // NHWC bf16 [B,H,W,C]
  int C;
  const float* rois;     // [R, 5] (batch index, x1, y1, x2, y2)
  const int* n_rois;     // device scalar (<= R) or null
  int R, P;
  void* out;             // [R, P, P, C] bf16 (or fp32 when out_fp32)
  int out_fp32;
};
int launch_roi_align(const RoiAlignArgs& a, cudaStream_t s);

// ---- box predictor post-processing (fast_rcnn.py:86-140, 257-326; postprocessing.py:11-61) -------
struct BoxPredictArgs {
  const float* head;        // [B*R, 16] fp32: cls logits (2), bbox deltas (4), pad
  const float* prop_boxes;  // [B, R, 4]
  const int* prop_count;    // [B]
  int B, R;
  float score_thresh, nms_thresh; int topk;
  float scale_x, scale_y, out_w, out_h;   // detector_postprocess
  float* ws_boxes;          // workspace [B, 1024, 4]
  unsigned char* ws_keep;   // workspace [B, 1024]
  // outputs
  float* det_boxes_raw;     // [B, topk, 4] network-input coordinates (feed the DensePose pooler)
  float* det_boxes;         // [B, topk, 4] original-image coordinates, clipped
  float* det_scores;        // [B, topk]
  int* det_count;           // [B]
};
int launch_box_predict(const BoxPredictArgs& a, cudaStream_t s);
// Pack per-image detections into one ROI list: rois [B*topk, 5], total count, per-image offsets [B+1].
int launch_pack_rois(const float* det_boxes_raw, const int* det_count, int B, int topk, float* rois,
                     int* total, int* offsets, cudaStream_t s);
// rois [B*R,5] from proposals (batch idx = b), count = B*R (all slots; invalid slots are zero boxes)
int launch_proposal_rois(const float* prop_boxes, const int* prop_count, int B, int R, float* rois,
                         cudaStream_t s);

// ---- GroupNorm(32) + ReLU on NHWC bf16 (deeplab.py:45,70-73,90,101) -----------------------------
// x: [R, HW, C] -> y: [R, HW, y_cstride] at channel offset baked into y; optionally broadcast a 1-pixel
// input over out_hw pixels (ASPP pooling branch, deeplab.py:109).
int launch_groupnorm_relu(const bf16* x, const float* gamma, const float* beta, bf16* y, int R, int HW,
                          int C, int y_cstride, int out_hw, const int* n_valid, cudaStream_t s);
// mean over HW: [R, HW, C] -> [R, C]  (AdaptiveAvgPool2d(1), deeplab.py:99)
int launch_avgpool(const bf16* x, bf16* y, int R, int HW, int C, const int* n_valid, cudaStream_t s);

// ---- predictor tail: bilinear x2 of the deconv output (chart.py:62-90) -> NCHW fp32 / fp16 tensors ---------------------
// low: [R, 2, 2, Cpad, S/2, S/2] fp32, the deconv output phases (py, px) as separate channel planes; channels run
// coarse[Kc], fine[25], u[25], v[25] and then the confidence heads a WC* model carries (chart_with_confidence.py:50-89:
// sigma_2[25], kappa_u[25], kappa_v[25], fine_segm_confidence[1], coarse_segm_confidence[1]). Output i is [R, ch[i], 2S, 2S];
// a null dst skips that head.
static constexpr int kMaxUpsampleOutputs = 9;
struct UpsampleOutputs { void* dst[kMaxUpsampleOutputs]; int ch[kMaxUpsampleOutputs]; int n; int total; };
int launch_predictor_upsample(const float* low, int R, int S, int Cpad, const int* n_valid, const UpsampleOutputs& outs,
                              int out_half, cudaStream_t s);

// Opt-in shared-memory attributes of the stage kernels on the current device (idempotent).
int stage_kernels_init();

// ---- strict (fp32-class) numerics: the same stage ops on bf16 hi/lo pairs (strict.cu) ---------------------------------
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi); every kernel reads both halves, computes in fp32 and splits again.
struct RoiAlignSplit { const bf16* feat_lo[4]; void* out_lo; };
int launch_maxpool3x3s2_split(const bf16* xh, const bf16* xl, bf16* yh, bf16* yl, int B, int H, int W, int C, cudaStream_t s);
int launch_upsample2x_split(const bf16* xh, const bf16* xl, bf16* yh, bf16* yl, int B, int H, int W, int C, cudaStream_t s);
// every argument is a {hi, lo} pointer pair
int launch_decoder_merge_split(const bf16* const* a, const bf16* const* b3, const bf16* const* b4, const bf16* const* b5,
                               bf16* const* out, int B, int H, int W, int C, cudaStream_t s);
int launch_roi_align_split(const RoiAlignArgs& a, const RoiAlignSplit& sp, cudaStream_t s);
int launch_groupnorm_relu_split(const bf16* xh, const bf16* xl, const float* gamma, const float* beta, bf16* yh, bf16* yl,
                                int R, int HW, int C, int y_cstride, int out_hw, const int* n_valid, cudaStream_t s);
int launch_avgpool_split(const bf16* xh, const bf16* xl, bf16* yh, bf16* yl, int R, int HW, int C, const int* n_valid, cudaStream_t s);

// ---- per-box DensePose resample (visualizer.py:10-56) -------------------------------------------
struct ResampleArgs {
  const float* coarse; const float* fine; const float* u; const float* v;   // [D, C, S, S] fp32
  int D, Kc, S;
  const int* box_wh;        // [D, 2] (w, h) already max(int, 1)  (host computed from boxes)
  const long long* offsets; // [D + 1] pixel offsets into the packed outputs
  void* labels;             // packed int64 [sum h*w] (uint8 when labels_u8)
  float* uv;                // packed fp32 [sum 2*h*w] (box i at 2*offsets[i]: u plane then v plane)
  long long total_pixels;
  int labels_u8;            // part labels are 0..24: one byte each instead of the reference's int64
};
int launch_dp_resample(const ResampleArgs& a, cudaStream_t s);

}  // namespace dpb
