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

#define DPB200_ABI_VERSION 4

const char* dpb200_last_error(void);
int dpb200_abi_version(void);
/* 1 when a CUDA device of compute capability 10.x is current, else 0 (never falls back). */
int dpb200_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * conv2d / linear / deconv-phase as a tcgen05 implicit GEMM.
 * Replaces F.conv2d in Conv2d.forward (detectron2/layers/wrappers.py:104-112) with FrozenBN
 * (detectron2/layers/batch_norm.py:54-62) pre-folded into w/bias, nn.Linear in
 * FastRCNNConvFCHead.forward (detectron2/modeling/roi_heads/box_head.py:95-98) and
 * FastRCNNOutputLayers.forward (fast_rcnn.py:238-257), and the four output-parity phases of
 * ConvTranspose2d(k=4,s=2,p=1) in DensePoseChartPredictor (densepose/modeling/predictors/chart.py:45-59),
 * one at a time or all four as the N blocks of one launch (phase_taps).
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
  int64_t y_sc;             /* output channel stride in elements (0 or 1 = NHWC; >1 = channel-planar, fp32 only) */
  int32_t epilogue;         /* 0 automatic; 1 direct global stores; 2 shared-memory slabs + TMA store (bf16
                               [M,C] outputs; the residual is then prefetched by TMA too)            */
  int32_t ks;               /* 64-channel K chunks per pipeline stage: 0 automatic (2 for N tiles <= 128), 1, 2 */
  int32_t phase_taps;       /* 1: the whole ConvTranspose2d(k=4,s=2,p=1) in one launch. wgt = the four phase
                               matrices (py,px) stacked on cout ([4*c][4*cin_pad], kh=kw=2, pad 1, stride 1);
                               N block (py,px) reads taps (ky+py, kx+px) of the pad-1 3x3 footprint and writes
                               channels [(2*py+px)*c, +c) of y (h_out = h, w_out = w).                    */
  int32_t pair;             /* CTA pairs: clusters of two CTAs run one tcgen05.mma.cta_group::2 on a 256-row tile,
                               each CTA staging half of the weight tile. 0 automatic (long-K 256-wide tiles), 1 off,
                               2 on (needs tiled == 0 and an N tile that is a multiple of 16; works with both
                               epilogues; measured neutral for the N=80 deconv phases, so not chosen there) */
  /* Strict (fp32-class) numerics, selected by x_lo != NULL: activations are pairs of bf16 tensors with x = hi + lo
   * (hi = bf16(x), lo = bf16(x - hi)); wgt is then packed per tap as three K segments [w_hi | w_lo | w_hi]
   * ([cout_pad][kh*kw*3*cin_pad]) and the kernel accumulates x_hi*w_hi + x_hi*w_lo + x_lo*w_hi in fp32 (three
   * tcgen05.mma passes into one TMEM accumulator). x_lo / res_lo / y_lo share the strides of x / res / y. */
  const void* x_lo; const void* res_lo; void* y_lo;
} dpb200_conv2d_args;

int dpb200_conv2d(const dpb200_conv2d_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage ops (HBM-bound kernels).  Each replaces the reference seam named beside it.
 */

/* DefaultPredictor.forward resize (detectron2/engine/defaults.py:85-89, F.interpolate bilinear,
 * scale_factor=k) fused with GeneralizedRCNN.preprocess_image (meta_arch/rcnn.py:156-181: normalise,
 * zero-pad to /32).  src [B,H0,W0,3] HWC fp32 (or u8); dst = space-to-depth stem layout [B,Hp/2,Wx,16] bf16:
 * padded-image pixel (y,x,c) at [y/2][x/2+2][((y&1)*2+(x&1))*4+c], Wx = Wp/2+4, channel 3 of each sub-pixel and
 * everything outside the resized image is 0 (on it the 7x7/2 stem conv is a 4x4/1 conv over 16 channels).
 * Both input types reproduce ATen's CPU kernels bit for bit: fp32 through its FMA-contracted float bilinear
 * (variant 0: the separable kernel a multi-threaded reference runs; 1: the channels-last kernel a single-threaded
 * one runs for a 3-channel image), u8 (what run.py:33-36 feeds) through its int16 fixed-point two-pass scheme,
 * whose per-axis weight tables dpb200_u8_resize_tables builds on the device. */
typedef struct dpb200_preprocess_args {
  const void* src; int32_t src_u8; int32_t b, h0, w0;
  int32_t hr, wr; float inv_scale; int32_t flip_rgb;
  float mean[3]; float std[3];
  void* dst; int32_t hp, wx;
  const void* tables;       /* u8 input: (1 + hr + wr) * 8 bytes filled by dpb200_u8_resize_tables; else NULL */
  int32_t variant;          /* fp32 input: 0 | 1 (see above)                                              */
  void* dst_lo;             /* strict numerics: low half of the bf16 hi/lo split of dst, else NULL        */
} dpb200_preprocess_args;
int dpb200_preprocess(const dpb200_preprocess_args* a, void* stream);
/* ATen's uint8 upsample_bilinear2d weights (UpSampleKernelAVXAntialias.h scheme: double-precision centres
 * scale*(i+0.5), scale = 1.0/k, int16 weights at the largest precision that keeps them below 2^15), both axes. */
int dpb200_u8_resize_tables(void* tables, int32_t h0, int32_t hr, int32_t w0, int32_t wr, double scale, void* stream);

/* F.max_pool2d(k=3,s=2,p=1) of BasicStem.forward (backbone/resnet.py:353). NHWC bf16. */
int dpb200_maxpool3x3s2(const void* x, void* y, int32_t b, int32_t h, int32_t w, int32_t c, void* stream);

/* nn.Upsample(x2, bilinear, align_corners=False) of Decoder (densepose/modeling/roi_heads/roi_head.py:63). */
int dpb200_upsample2x(const void* x, void* y, int32_t b, int32_t h, int32_t w, int32_t c, void* stream);
/* Decoder.forward sum of the four scale heads (roi_head.py:71-77): out = a + up2(b3) + up2(b4) + up2(b5). */
int dpb200_decoder_merge(const void* a, const void* b3, const void* b4, const void* b5, void* out,
                         int32_t b, int32_t h, int32_t w, int32_t c, void* stream);

/* RPN proposal selection: RPN.forward / _decode_proposals (proposal_generator/rpn.py:319-394),
 * Box2BoxTransform.apply_deltas (box_regression.py:74-112), DefaultAnchorGenerator
 * (anchor_generator.py:165-231) and find_top_rpn_proposals (proposal_utils.py:19-134) including the
 * per-level batched_nms (layers/nms.py:9-20).  head[l]: [B,H_l,W_l,16] fp32 (3 logits, 12 deltas, pad). */
typedef struct dpb200_rpn_args {
  const float* head[5]; int32_t h[5], w[5]; float stride[5]; float anchors[5][12];
  int32_t b, pre_topk, post_topk; float nms_thresh; float clip_x, clip_y;
  float* cand_boxes; float* cand_scores; int32_t* cand_count; uint8_t* cand_keep;   /* workspace */
  float* prop_boxes; float* prop_scores; int32_t* prop_count;                        /* outputs   */
} dpb200_rpn_args;
int dpb200_rpn_proposals(const dpb200_rpn_args* a, void* stream);

/* torchvision.ops.nms on boxes already sorted by descending score (layers/nms.py:20). n <= 1024.
 * keep[i] is 1 on entry for usable rows; on exit 1 for kept rows. */
int dpb200_nms_sorted(const float* boxes, int32_t n, float thr, uint8_t* keep, void* stream);

/* ROIPooler.forward (modeling/poolers.py:187-227) with assign_boxes_to_levels (poolers.py:15-51) and
 * torchvision.ops.roi_align(aligned=False, sampling_ratio=2) (layers/roi_align.py:58-65).
 * feat[l]: NHWC bf16 [B,H_l,W_l,C]; rois [R,5]; out [R,P,P,C] bf16 (fp32 if out_fp32). */
typedef struct dpb200_roi_align_args {
  const void* feat[4]; int32_t h[4], w[4]; float scale[4]; int32_t n_levels; int32_t c;
  const float* rois; const int32_t* n_rois; int32_t r, p;
  void* out; int32_t out_fp32;
} dpb200_roi_align_args;
int dpb200_roi_align(const dpb200_roi_align_args* a, void* stream);

/* FastRCNNOutputLayers.inference (roi_heads/fast_rcnn.py:257-326) + fast_rcnn_inference_single_image
 * (fast_rcnn.py:86-140) + detector_postprocess (modeling/postprocessing.py:11-61). */
typedef struct dpb200_box_predict_args {
  const float* head; const float* prop_boxes; const int32_t* prop_count; int32_t b, r;
  float score_thresh, nms_thresh; int32_t topk; float scale_x, scale_y, out_w, out_h;
  float* ws_boxes; uint8_t* ws_keep;
  float* det_boxes_raw; float* det_boxes; float* det_scores; int32_t* det_count;
} dpb200_box_predict_args;
int dpb200_box_predict(const dpb200_box_predict_args* a, void* stream);

/* nn.GroupNorm(32, C) + ReLU of DensePoseDeepLabHead (densepose/modeling/roi_heads/deeplab.py:45,70-73,90,101).
 * x [R,HW,C] bf16 -> y [R,out_hw,y_cstride] bf16 (out_hw != hw only for hw == 1: broadcast, deeplab.py:109). */
int dpb200_groupnorm_relu(const void* x, const float* gamma, const float* beta, void* y, int32_t r,
                          int32_t hw, int32_t c, int32_t y_cstride, int32_t out_hw, const int32_t* n_valid,
                          void* stream);
/* nn.AdaptiveAvgPool2d(1) of ASPPPooling (deeplab.py:99). x [R,HW,C] bf16 -> y [R,C] bf16. */
int dpb200_avgpool(const void* x, void* y, int32_t r, int32_t hw, int32_t c, const int32_t* n_valid, void* stream);

/* interp2d (bilinear x2) of DensePoseChartPredictor.forward (densepose/modeling/predictors/chart.py:62-90).
 * low [R,2,2,cpad,S/2,S/2] fp32 — the four ConvTranspose output phases (py,px) as channel planes, the layout
 * dpb200_conv2d writes with y_sc > 1 — -> n NCHW fp32 tensors: out[i] is [R,ch[i],2S,2S] and takes the next ch[i]
 * channels of low (the engine's order: coarse[kc], fine[25], u[25], v[25], then the confidence heads of a WC* model,
 * chart_with_confidence.py:50-89). A NULL out[i] skips that head. Bit-identical to ATen's CPU bilinear, including its
 * switch to the channels-last kernel when the output h + w <= 128 (the legacy 56x56 heads). */
int dpb200_predictor_upsample(const float* low, int32_t r, int32_t s, int32_t cpad, const int32_t* n_valid,
                              void* const* out, const int32_t* ch, int32_t n, void* stream);

/* DensePoseResultExtractor (visualizer.py:10-56): per-box resize + argmax + U/V gather.
 * box_wh [D,2] = (max(int(w),1), max(int(h),1)); offsets [D+1] pixel prefix sums; labels int64 packed
 * (one uint8 each when labels_u8 != 0: part labels are 0..24); uv fp32 packed (box i at 2*offsets[i]: U plane
 * then V plane). */
typedef struct dpb200_resample_args {
  const float* coarse; const float* fine; const float* u; const float* v;
  int32_t d, kc, s; const int32_t* box_wh; const int64_t* offsets;
  void* labels; float* uv; int64_t total_pixels; int32_t labels_u8;
} dpb200_resample_args;
int dpb200_dp_resample(const dpb200_resample_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole forward pass: DefaultPredictor.forward (engine/defaults.py:65-97) -> GeneralizedRCNN.inference
 * (meta_arch/rcnn.py:110-154), batched over B independent images of one size.
 */
typedef struct dpb200_model_config {
  int32_t depth;            /* 50 | 101  (MODEL.RESNETS.DEPTH)                                   */
  int32_t head;             /* 0 DensePoseV1ConvXHead, 1 DensePoseDeepLabHead                     */
  int32_t decoder_on;       /* ROI_DENSEPOSE_HEAD.DECODER_ON                                      */
  int32_t pooler_res;       /* 14 | 28                                                            */
  int32_t coarse_ch;        /* 15 | 2                                                             */
  float score_thresh, nms_test, rpn_nms;
  int32_t dets_per_image, rpn_pre_topk, rpn_post_topk;
  int32_t min_size, max_size;
  float pixel_mean[3], pixel_std[3];
  int32_t input_rgb;        /* INPUT.FORMAT == "RGB"                                              */
  int32_t extra_ch[5];      /* confidence heads a WC* model carries, in this order: sigma_2, kappa_u, kappa_v,
                               fine_segm_confidence, coarse_segm_confidence (chart_with_confidence.py:50-89): channel count
                               (25, 25, 25, 1, 1) or 0 when the head is absent. The reference builds these layers but its
                               forward drops them (chart_with_confidence.py:91-109); the engine can emit them (f4). */
  int32_t resize_variant;   /* float images: which of ATen's two CPU bilinear kernels the resize reproduces bit for bit
                               (dpb200_preprocess_args.variant): 0 = a reference running with > 1 intra-op threads,
                               1 = a single-threaded one. uint8 images have one kernel.                       */
  int32_t strict;           /* 1: fp32-class numerics end to end — activations and weights as bf16 hi/lo pairs, three
                               tensor-core passes per product, fp32 accumulate (the weights handed to
                               dpb200_model_create must then be packed with three K segments per tap); the mode in
                               which proposals, NMS keep lists and label maps are compared with the reference BY INDEX */
} dpb200_model_config;

/* One packed parameter. conv / linear / deconv-phase: data0 = bf16 [cout_pad][k*cin_pad], data1 = fp32
 * bias [cout_pad] (or NULL). GroupNorm: data0 = fp32 gamma, data1 = fp32 beta. Device pointers owned by
 * the caller and kept alive for the model's lifetime. Names: see densepose_torchscript_b200/weights.py. */
typedef struct dpb200_weight {
  const char* name; const void* data0; const void* data1; int32_t cin_pad, cout_pad;
} dpb200_weight;

typedef struct dpb200_model dpb200_model;
typedef struct dpb200_session dpb200_session;

int dpb200_model_create(const dpb200_model_config* cfg, const dpb200_weight* w, int32_t n, dpb200_model** out);
void dpb200_model_destroy(dpb200_model* m);

/* A session fixes (batch, input height, input width, input dtype) and owns the launch plan (TMA
 * descriptors, grid sizes) over a caller-provided workspace. */
size_t dpb200_session_workspace_bytes(const dpb200_model* m, int32_t b, int32_t h0, int32_t w0);
int dpb200_session_create(const dpb200_model* m, int32_t b, int32_t h0, int32_t w0, int32_t src_u8,
                          void* workspace, size_t workspace_bytes, dpb200_session** out);
void dpb200_session_destroy(dpb200_session* s);

typedef struct dpb200_forward_io {
  const void* images;       /* [B,H0,W0,3] HWC, fp32 or u8 as the session was created            */
  int32_t bgr;              /* the reference's `bgr` argument (defaults.py:65)                    */
  float* pred_boxes;        /* [B, dets_per_image, 4] original-image pixels                       */
  float* scores;            /* [B, dets_per_image]                                                */
  int32_t* det_count;       /* [B]                                                                */
  int32_t* det_offsets;     /* [B+1] start of each image's rows in the packed DensePose tensors   */
  void* coarse;             /* [B*dets_per_image, coarse_ch, 4S, 4S] packed, NCHW fp32 (fp16: out_half) */
  void* fine;               /* [B*dets_per_image, 25, 4S, 4S]                                     */
  void* u;                  /* [B*dets_per_image, 25, 4S, 4S]                                     */
  void* v;                  /* [B*dets_per_image, 25, 4S, 4S]                                     */
  int32_t out_half;         /* != 0: the four DensePose tensors are written as IEEE fp16 (what the reference's
                             * `.half()` module returns, run.py:20-29); boxes and scores stay fp32         */
  void* extra[5];           /* [B*dets_per_image, extra_ch[i], 4S, 4S] for the confidence heads the model carries (same
                             * dtype as u / v), or NULL to skip a head                                      */
} dpb200_forward_io;
/* Enqueues one forward pass. Everything is ordered after the work already on `stream` and is complete, as far as `stream`
 * can tell, when the call's last launch has run: internally the launches form a two-branch graph (the proposal / box chain
 * and the small FPN / RPN levels run on a side stream the session owns, forked from and joined back into `stream` with
 * events), so the caller only ever synchronises with `stream`. The side stream and its events are created on the first
 * run; nothing else is allocated. One run at a time per session (use one session per concurrent stream). */
int dpb200_session_run(dpb200_session* s, const dpb200_forward_io* io, void* stream);
/* enable != 0: dpb200_session_run captures its launch sequence into a CUDA graph the first time it sees an
 * io binding (all pointers + bgr) and replays the instantiated graph afterwards (one host call instead of
 * ~110 launches; instantiation is the only hidden allocation in the library). `stream` must then be a real
 * stream: on the legacy default stream (NULL) the kernels are launched directly. */
int dpb200_session_set_graph(dpb200_session* s, int32_t enable);

/* Kernel launches one run enqueues, and their algorithmic FLOPs (2*MAC of every conv/linear at full
 * detection capacity). */
int dpb200_session_launch_count(const dpb200_session* s);
/* Name ("conv:<weight>" or the stage kernel) and padded-shape 2*MAC of launch i. */
int dpb200_session_op_info(const dpb200_session* s, int32_t i, char* name, int32_t cap, double* flops);
/* Algorithmic HBM bytes of launch i at full detection capacity: every operand read once (pixels a strided
 * 1x1 conv skips excluded), every output written once; 0 for the latency-bound bookkeeping kernels. The
 * roofline numerator of the memory-bound stages. */
int dpb200_session_op_bytes(const dpb200_session* s, int32_t i, double* bytes);
/* Like dpb200_session_run but brackets every launch with CUDA events on `stream`, synchronises, and writes
 * the per-launch milliseconds to ms[0..launch_count). Measurement aid for bench.py, not the serving path. */
int dpb200_session_profile(dpb200_session* s, const dpb200_forward_io* io, void* stream, float* ms, int32_t cap);
double dpb200_session_flops(const dpb200_session* s);
/* Resized / padded extents the session computes: out[0..3] = Hr, Wr, Hp, Wp. */
void dpb200_session_geometry(const dpb200_session* s, int32_t out[4]);
/* Intermediate tensors by name (stage-parity tests): device pointer, shape (up to 4 dims, NHWC), dtype
 * (0 bf16, 1 fp32, 2 int32, 3 u8). Returns 0 if the name exists. */
int dpb200_session_tap(const dpb200_session* s, const char* name, void** ptr, int64_t shape[4], int32_t* dtype);

#ifdef __cplusplus
}
#endif
#endif /* DPB200_H_ */
