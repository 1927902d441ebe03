// C-ABI entry points (include/dpb200.h). Only plain C types cross this boundary.
// (model / session entry points live in engine.cu)
#include "../../include/dpb200.h"
#include "conv_igemm.cuh"
#include "kernels.cuh"

#include <string.h>

namespace {
int num_sms() {
  static int sms = 0;
  if (sms) return sms;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  return sms;
}
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
}  // namespace

using namespace dpb;

extern "C" {

const char* dpb200_last_error(void) { return dpb::get_error(); }
int dpb200_abi_version(void) { return DPB200_ABI_VERSION; }

int dpb200_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int dpb200_conv2d(const dpb200_conv2d_args* a, void* stream) {
  if (!a) { dpb::set_error("conv2d: null args"); return -1; }
  dpb::ConvDesc d;
  d.x = a->x; d.N = a->n; d.H = a->h; d.W = a->w; d.Cin = a->cin;
  d.x_sw = a->x_sw ? a->x_sw : a->cin;
  d.x_sh = a->x_sh ? a->x_sh : d.x_sw * a->w;
  d.x_sn = a->x_sn ? a->x_sn : d.x_sh * a->h;
  d.w = a->wgt; d.cin_pad = a->cin_pad; d.cout_pad = a->cout_pad; d.bias = a->bias;
  d.kh = a->kh; d.kw = a->kw; d.sx = a->sx; d.sy = a->sy; d.pad_x = a->pad_x; d.pad_y = a->pad_y;
  d.dil = a->dil; d.H_out = a->h_out; d.W_out = a->w_out; d.relu = a->relu;
  d.res = a->res; d.res_sn = a->res_sn; d.res_sy = a->res_sy; d.res_sx = a->res_sx;
  d.res_shift = a->res_shift;
  d.out = a->y; d.out_fp32 = a->y_fp32;
  d.out_sx = a->y_sx ? a->y_sx : a->cout_pad;
  d.out_sy = a->y_sy ? a->y_sy : d.out_sx * a->w_out;
  d.out_sn = a->y_sn ? a->y_sn : d.out_sy * a->h_out;
  d.n_valid = a->n_valid; d.block_n = a->block_n; d.stages = a->stages; d.ks = a->ks; d.phase_taps = a->phase_taps; d.pair = a->pair;
  d.im2col = a->tiled ? 0 : 1;
  d.out_sc = a->y_sc > 0 ? a->y_sc : 1;
  d.epilogue = a->epilogue;
  d.x2 = a->x_lo; d.res2 = a->res_lo; d.out2 = a->y_lo;
  dpb::ConvPlan plan;
  int r = dpb::conv_plan_build(&plan, d, num_sms());
  if (r) return r;
  return dpb::conv_plan_launch(plan, S(stream));
}

int dpb200_preprocess(const dpb200_preprocess_args* a, void* stream) {
  if (!a) { set_error("preprocess: null args"); return -1; }
  PreprocessArgs p{};
  p.src = a->src; p.src_u8 = a->src_u8; p.B = a->b; p.H0 = a->h0; p.W0 = a->w0; p.Hr = a->hr; p.Wr = a->wr;
  p.inv_scale = a->inv_scale; p.flip_rgb = a->flip_rgb;
  for (int i = 0; i < 3; ++i) { p.mean[i] = a->mean[i]; p.std[i] = a->std[i]; }
  p.dst = reinterpret_cast<bf16*>(a->dst); p.Hp = a->hp; p.Wx = a->wx;
  p.tables = reinterpret_cast<const int2*>(a->tables); p.variant = a->variant;
  p.dst_lo = reinterpret_cast<bf16*>(a->dst_lo);
  return launch_preprocess(p, S(stream));
}

int dpb200_u8_resize_tables(void* tables, int32_t h0, int32_t hr, int32_t w0, int32_t wr, double scale, void* stream) {
  if (!tables) { set_error("u8_resize_tables: null table"); return -1; }
  return launch_u8_resize_tables(reinterpret_cast<int2*>(tables), h0, hr, w0, wr, scale, S(stream));
}

int dpb200_maxpool3x3s2(const void* x, void* y, int32_t b, int32_t h, int32_t w, int32_t c, void* stream) {
  return launch_maxpool3x3s2((const bf16*)x, (bf16*)y, b, h, w, c, S(stream));
}
int dpb200_upsample2x(const void* x, void* y, int32_t b, int32_t h, int32_t w, int32_t c, void* stream) {
  return launch_upsample2x((const bf16*)x, (bf16*)y, b, h, w, c, S(stream));
}
int dpb200_decoder_merge(const void* a, const void* b3, const void* b4, const void* b5, void* out, int32_t b,
                         int32_t h, int32_t w, int32_t c, void* stream) {
  return launch_decoder_merge((const bf16*)a, (const bf16*)b3, (const bf16*)b4, (const bf16*)b5, (bf16*)out, b, h,
                              w, c, S(stream));
}

int dpb200_rpn_proposals(const dpb200_rpn_args* a, void* stream) {
  if (!a) { set_error("rpn: null args"); return -1; }
  RpnArgs r{};
  for (int l = 0; l < 5; ++l) {
    r.lvl[l].head = a->head[l]; r.lvl[l].H = a->h[l]; r.lvl[l].W = a->w[l]; r.lvl[l].stride = a->stride[l];
    memcpy(r.lvl[l].anchors, a->anchors[l], sizeof(float) * 12);
  }
  r.B = a->b; r.pre_topk = a->pre_topk; r.post_topk = a->post_topk; r.nms_thresh = a->nms_thresh;
  r.clip_x = a->clip_x; r.clip_y = a->clip_y;
  r.cand_boxes = a->cand_boxes; r.cand_scores = a->cand_scores; r.cand_count = a->cand_count;
  r.cand_keep = a->cand_keep; r.prop_boxes = a->prop_boxes; r.prop_scores = a->prop_scores;
  r.prop_count = a->prop_count;
  int rc = launch_rpn_topk_decode(r, S(stream));
  if (rc) return rc;
  rc = launch_rpn_nms(r, S(stream));
  if (rc) return rc;
  return launch_rpn_merge(r, S(stream));
}

int dpb200_nms_sorted(const float* boxes, int32_t n, float thr, uint8_t* keep, void* stream) {
  return launch_nms_sorted(boxes, n, thr, keep, S(stream));
}

int dpb200_roi_align(const dpb200_roi_align_args* a, void* stream) {
  if (!a) { set_error("roi_align: null args"); return -1; }
  RoiAlignArgs r{};
  for (int l = 0; l < 4; ++l) { r.feat[l] = (const bf16*)a->feat[l]; r.H[l] = a->h[l]; r.W[l] = a->w[l]; r.scale[l] = a->scale[l]; }
  r.n_levels = a->n_levels; r.C = a->c; r.rois = a->rois; r.n_rois = a->n_rois; r.R = a->r; r.P = a->p;
  r.out = a->out; r.out_fp32 = a->out_fp32;
  return launch_roi_align(r, S(stream));
}

int dpb200_box_predict(const dpb200_box_predict_args* a, void* stream) {
  if (!a) { set_error("box_predict: null args"); return -1; }
  BoxPredictArgs p{};
  p.head = a->head; p.prop_boxes = a->prop_boxes; p.prop_count = a->prop_count; p.B = a->b; p.R = a->r;
  p.score_thresh = a->score_thresh; p.nms_thresh = a->nms_thresh; p.topk = a->topk;
  p.scale_x = a->scale_x; p.scale_y = a->scale_y; p.out_w = a->out_w; p.out_h = a->out_h;
  p.ws_boxes = a->ws_boxes; p.ws_keep = a->ws_keep;
  p.det_boxes_raw = a->det_boxes_raw; p.det_boxes = a->det_boxes; p.det_scores = a->det_scores;
  p.det_count = a->det_count;
  return launch_box_predict(p, S(stream));
}

int dpb200_groupnorm_relu(const void* x, const float* gamma, const float* beta, void* y, int32_t r, int32_t hw,
                          int32_t c, int32_t y_cstride, int32_t out_hw, const int32_t* n_valid, void* stream) {
  return launch_groupnorm_relu((const bf16*)x, gamma, beta, (bf16*)y, r, hw, c, y_cstride, out_hw, n_valid, S(stream));
}
int dpb200_avgpool(const void* x, void* y, int32_t r, int32_t hw, int32_t c, const int32_t* n_valid, void* stream) {
  return launch_avgpool((const bf16*)x, (bf16*)y, r, hw, c, n_valid, S(stream));
}
int dpb200_predictor_upsample(const float* low, int32_t r, int32_t s, int32_t cpad, const int32_t* n_valid,
                              void* const* out, const int32_t* ch, int32_t n, void* stream) {
  if (!out || !ch || n < 1 || n > kMaxUpsampleOutputs) { set_error("predictor_upsample: bad output list"); return -1; }
  UpsampleOutputs o{};
  o.n = n;
  for (int i = 0; i < n; ++i) { o.dst[i] = out[i]; o.ch[i] = ch[i]; o.total += ch[i]; }
  return launch_predictor_upsample(low, r, s, cpad, n_valid, o, 0, S(stream));
}

int dpb200_dp_resample(const dpb200_resample_args* a, void* stream) {
  if (!a) { set_error("dp_resample: null args"); return -1; }
  ResampleArgs r{};
  r.coarse = a->coarse; r.fine = a->fine; r.u = a->u; r.v = a->v; r.D = a->d; r.Kc = a->kc; r.S = a->s;
  r.box_wh = a->box_wh; r.offsets = (const long long*)a->offsets; r.labels = a->labels;
  r.uv = a->uv; r.total_pixels = a->total_pixels; r.labels_u8 = a->labels_u8;
  return launch_dp_resample(r, S(stream));
}

}  // extern "C"
