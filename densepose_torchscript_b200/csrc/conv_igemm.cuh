// Internal (C++) interface of the tcgen05 implicit-GEMM convolution. The C-ABI wrapper lives in
// api.cu; the whole-graph forward in engine.cu builds ConvPlans once and replays them.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace dpb {

// Describes one convolution / linear layer as an implicit GEMM
//   out[n, oy, ox, co] = act( sum_{ky,kx,ci} x[n, oy*sy + ky*dil - pad_y, ox*sx + kx*dil - pad_x, ci]
//                                            * w[co, (ky*kw + kx)*cin_pad + ci] + bias[co] + res )
// x is bf16 with arbitrary element strides (channel stride must be 1), w is the packed K-major
// bf16 matrix [cout_pad][kh*kw*cin_pad] (cin_pad = multiple of 64, cout_pad = multiple of 16).
struct ConvDesc {
  const void* x = nullptr;
  int N = 1, H = 1, W = 1, Cin = 64;              // input extents seen by the TMA (W/H may be virtual)
  long long x_sn = 0, x_sh = 0, x_sw = 0;         // input strides in elements
  const void* w = nullptr;
  int cin_pad = 64, cout_pad = 16;
  const float* bias = nullptr;                    // [cout_pad] fp32 or null
  int kh = 1, kw = 1, sx = 1, sy = 1, pad_x = 0, pad_y = 0, dil = 1;
  int H_out = 1, W_out = 1;
  int relu = 0;
  const void* res = nullptr;                      // bf16 residual, or null
  long long res_sn = 0, res_sy = 0, res_sx = 0;   // residual strides (elements)
  int res_shift = 0;                              // 1: residual pixel = (oy>>1, ox>>1) (FPN top-down)
  void* out = nullptr;
  int out_fp32 = 0;
  long long out_sn = 0, out_sy = 0, out_sx = 0;   // output strides (elements)
  long long out_sc = 1;                           // output channel stride; != 1: channel-planar fp32 output
  int epilogue = 0;                               // 0 choose, 1 direct global stores, 2 smem slabs + TMA store
  const int* n_valid = nullptr;                   // device scalar: images actually present (<= N)
  int im2col = 1;                                 // A operand through im2col-mode TMA (else tiled boxes)
  int block_n = 0;                                // 0 = choose
  int stages = 0;                                 // 0 = choose
  int ks = 0;                                     // 64-channel chunks per pipeline stage: 0 = choose, 1 or 2
  // ConvTranspose2d(k4,s2,p1) as ONE launch: the four output-parity phases are the four N blocks (block_n =
  // cout_pad / 4); block (py,px) reads its 2x2 taps at window offsets (ky+py, kx+px) of a pad-1 3x3 footprint,
  // so the four CTAs working on one pixel tile share the same input rows through L2.
  int phase_taps = 0;
  int pair = 0;                                   // CTA pairs (cta_group::2, weights split over the pair): 0 choose, 1 off, 2 on
  // Strict (fp32-class) numerics: every activation is a PAIR of bf16 tensors (hi, lo) with x = hi + lo
  // (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits), weights likewise, packed per tap as three K segments
  // [w_hi | w_lo | w_hi]; the K loop then runs x_hi*w_hi + x_hi*w_lo + x_lo*w_hi into the same fp32 TMEM
  // accumulator (three tcgen05.mma passes, the lo*lo term of relative size 2^-18 is dropped).
  const void* x2 = nullptr;                       // lo half of the input (same strides as x); non-null selects the mode
  const void* res2 = nullptr;                     // lo half of the residual
  void* out2 = nullptr;                           // lo half of a bf16 output (same strides as out)
};

struct ConvKParams {
  int tw, th, tn;
  int tiles_w, tiles_h, tiles_n;
  int n_blocks, block_n;
  int H_out, W_out, N;
  int kh, kw, sx, sy, pad_x, pad_y, dil;
  int cin_chunks, stages, acc_stride;
  int ks;                                         // 64-channel K chunks per stage group (one barrier pair each)
  int nslab, res_tma;                             // staged epilogue: slab ring size, residual through TMA
  int relu, out_fp32, res_shift, im2col;
  int phase_taps;
  int nseg;                                       // K segments per tap: 1, or 3 in strict mode (x_hi*w_hi, x_hi*w_lo, x_lo*w_hi)
  const __nv_bfloat16* res2;
  void* out2;
  const float* bias;
  const __nv_bfloat16* res;
  long long res_sn, res_sy, res_sx;
  void* out;
  long long out_sn, out_sy, out_sx, out_sc;
  const int* n_valid;
};

struct ConvPlan {
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmA2;    // strict mode: the lo half of the input (else a copy of tmA)
  alignas(64) CUtensorMap tmB;
  alignas(64) CUtensorMap tmOut;   // staged epilogue: [M, cout] bf16 view of the output
  alignas(64) CUtensorMap tmRes;   // staged epilogue: [M, cout] bf16 view of the residual
  int staged = 0;
  int pair = 0;                    // launched as clusters of two CTAs (cta_group::2)
  ConvKParams p;
  int grid = 0;
  int smem = 0;
  double flops = 0;   // algorithmic 2*MAC of this layer (for reporting)
};

// Host only; no GPU work. Returns 0 or a negative error (message via dpb::set_error).
int conv_plan_build(ConvPlan* plan, const ConvDesc& d, int num_sms);
int conv_plan_launch(const ConvPlan& plan, cudaStream_t stream);
// Sets the opt-in shared-memory attribute of the conv kernel(s) on the current device (idempotent).
int conv_kernels_init();

void set_error(const char* fmt, ...);
const char* get_error();

}  // namespace dpb
