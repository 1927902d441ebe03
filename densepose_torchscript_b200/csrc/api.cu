// C-ABI entry points (include/dpb200.h). Only plain C types cross this boundary.
#include "../../include/dpb200.h"
#include "conv_igemm.cuh"

namespace {
int num_sms() {
  static int sms = 0;
  if (sms) return sms;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  return sms;
}
}  // namespace

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
  d.n_valid = a->n_valid; d.block_n = a->block_n; d.stages = a->stages;
  d.im2col = a->tiled ? 0 : 1;
  dpb::ConvPlan plan;
  int r = dpb::conv_plan_build(&plan, d, num_sms());
  if (r) return r;
  return dpb::conv_plan_launch(plan, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
