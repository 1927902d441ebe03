// Stage kernels of the strict (fp32-class) numerics mode. Every activation is a pair of NHWC bf16 tensors (hi, lo) with
// x = hi + lo exactly representable in fp32 (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits); these kernels read
// both halves, do the reference's fp32 arithmetic on x, and split the result again. They mirror the bf16 kernels of
// elementwise.cu / roi.cu operation by operation (same reference lines), but are written for clarity, not for the last
// GB/s: this is the mode in which the engine is compared with the reference BY INDEX (proposals, NMS keep lists,
// detection count, label maps), not the throughput mode.
#include "kernels.cuh"
#include "conv_igemm.cuh"
#include "ptx.cuh"
#include "device_utils.cuh"

namespace dpb {

#define DPB_CHECK_LAUNCH(name)                                                     \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      set_error("%s launch: %s", name, cudaGetErrorString(e__));                   \
      return -4;                                                                   \
    }                                                                              \
  } while (0)

static inline int grid_for(long long work, int threads, int cap = 148 * 16) {
  long long g = (work + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// 8 consecutive channels: x = hi + lo (both 16-byte loads)
__device__ __forceinline__ void load8(const uint4* __restrict__ hi, const uint4* __restrict__ lo, long long i, float* f) {
  const uint4 h = __ldg(hi + i), l = __ldg(lo + i);
  f[0] = bf16_lo(h.x) + bf16_lo(l.x); f[1] = bf16_hi(h.x) + bf16_hi(l.x);
  f[2] = bf16_lo(h.y) + bf16_lo(l.y); f[3] = bf16_hi(h.y) + bf16_hi(l.y);
  f[4] = bf16_lo(h.z) + bf16_lo(l.z); f[5] = bf16_hi(h.z) + bf16_hi(l.z);
  f[6] = bf16_lo(h.w) + bf16_lo(l.w); f[7] = bf16_hi(h.w) + bf16_hi(l.w);
}
__device__ __forceinline__ void store8(uint4* __restrict__ hi, uint4* __restrict__ lo, long long i, const float* f) {
  uint4 h, l;
  h.x = pack_bf16(f[0], f[1]); h.y = pack_bf16(f[2], f[3]); h.z = pack_bf16(f[4], f[5]); h.w = pack_bf16(f[6], f[7]);
  l.x = pack_bf16(f[0] - bf16_lo(h.x), f[1] - bf16_hi(h.x)); l.y = pack_bf16(f[2] - bf16_lo(h.y), f[3] - bf16_hi(h.y));
  l.z = pack_bf16(f[4] - bf16_lo(h.z), f[5] - bf16_hi(h.z)); l.w = pack_bf16(f[6] - bf16_lo(h.w), f[7] - bf16_hi(h.w));
  hi[i] = h; lo[i] = l;
}

// ------------------------------------------------------------------------------------ max pool (resnet.py:353)
__global__ void maxpool3x3s2_split_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl, uint4* __restrict__ yh,
                                          uint4* __restrict__ yl, int B, int H, int W, int C8, int Ho, int Wo) {
  const long long total = (long long)B * Ho * Wo * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int ox = (int)((i / C8) % Wo);
    const int oy = (int)((i / ((long long)C8 * Wo)) % Ho);
    const int b = (int)(i / ((long long)C8 * Wo * Ho));
    float m[8];
    bool first = true;
    for (int dy = -1; dy <= 1; ++dy) {
      const int iy = oy * 2 + dy;
      if (iy < 0 || iy >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int ix = ox * 2 + dx;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        load8(xh, xl, (((long long)b * H + iy) * W + ix) * C8 + c, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = first ? v[k] : fmaxf(m[k], v[k]);
        first = false;
      }
    }
    store8(yh, yl, i, m);
  }
}

int launch_maxpool3x3s2_split(const bf16* xh, const bf16* xl, bf16* yh, bf16* yl, int B, int H, int W, int C, cudaStream_t s) {
  if (C % 8) { set_error("maxpool: C %% 8 != 0"); return -1; }
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)B * Ho * Wo * (C / 8);
  maxpool3x3s2_split_kernel<<<grid_for(total, 256), 256, 0, s>>>((const uint4*)xh, (const uint4*)xl, (uint4*)yh, (uint4*)yl, B, H, W,
                                                                C / 8, Ho, Wo);
  DPB_CHECK_LAUNCH("maxpool_split");
  return 0;
}

// ------------------------------------------------------------------------------------ bilinear x2 (roi_head.py:63,71-77)
// ATen's separable fp32 expression (oracle/aten_interp.py): top = fma(lx0, p00, lx1*p01), out = fma(ly0, top, ly1*bot)
__device__ __forceinline__ void up2_taps(int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float real = 0.5f * ((float)dst + 0.5f) - 0.5f;   // exact in fp32
  if (real < 0.f) real = 0.f;
  i0 = (int)real;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = real - (float)i0;
  l0 = 1.f - l1;
}
__device__ __forceinline__ void sample_up2(const uint4* __restrict__ xh, const uint4* __restrict__ xl, int b, int h, int w,
                                           int C8, int c, int oy, int ox, float* out) {
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  up2_taps(oy, h, y0, y1, ly0, ly1);
  up2_taps(ox, w, x0, x1, lx0, lx1);
  const long long base = (long long)b * h * w * C8 + c;
  float p00[8], p01[8], p10[8], p11[8];
  load8(xh, xl, base + ((long long)y0 * w + x0) * C8, p00);
  load8(xh, xl, base + ((long long)y0 * w + x1) * C8, p01);
  load8(xh, xl, base + ((long long)y1 * w + x0) * C8, p10);
  load8(xh, xl, base + ((long long)y1 * w + x1) * C8, p11);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float top = fmaf(lx0, p00[k], __fmul_rn(lx1, p01[k]));
    const float bot = fmaf(lx0, p10[k], __fmul_rn(lx1, p11[k]));
    out[k] = fmaf(ly0, top, __fmul_rn(ly1, bot));
  }
}

__global__ void upsample2x_split_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl, uint4* __restrict__ yh,
                                        uint4* __restrict__ yl, int B, int H, int W, int C8) {
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)B * Ho * Wo * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int ox = (int)((i / C8) % Wo);
    const int oy = (int)((i / ((long long)C8 * Wo)) % Ho);
    const int b = (int)(i / ((long long)C8 * Wo * Ho));
    float v[8];
    sample_up2(xh, xl, b, H, W, C8, c, oy, ox, v);
    store8(yh, yl, i, v);
  }
}

int launch_upsample2x_split(const bf16* xh, const bf16* xl, bf16* yh, bf16* yl, int B, int H, int W, int C, cudaStream_t s) {
  if (C % 8) { set_error("upsample2x: C %% 8 != 0"); return -1; }
  const long long total = (long long)B * 4 * H * W * (C / 8);
  upsample2x_split_kernel<<<grid_for(total, 256), 256, 0, s>>>((const uint4*)xh, (const uint4*)xl, (uint4*)yh, (uint4*)yl, B, H, W, C / 8);
  DPB_CHECK_LAUNCH("upsample2x_split");
  return 0;
}

// out = ((a + up2(b3)) + up2(b4)) + up2(b5)   (reference order, roi_head.py:73-77)
__global__ void decoder_merge_split_kernel(const uint4* __restrict__ ah, const uint4* __restrict__ al,
                                           const uint4* __restrict__ b3h, const uint4* __restrict__ b3l,
                                           const uint4* __restrict__ b4h, const uint4* __restrict__ b4l,
                                           const uint4* __restrict__ b5h, const uint4* __restrict__ b5l,
                                           uint4* __restrict__ oh, uint4* __restrict__ ol, int B, int H, int W, int C8) {
  const long long total = (long long)B * H * W * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int ox = (int)((i / C8) % W);
    const int oy = (int)((i / ((long long)C8 * W)) % H);
    const int b = (int)(i / ((long long)C8 * W * H));
    float acc[8], v[8];
    load8(ah, al, i, acc);
    sample_up2(b3h, b3l, b, H / 2, W / 2, C8, c, oy, ox, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = __fadd_rn(acc[k], v[k]);
    sample_up2(b4h, b4l, b, H / 2, W / 2, C8, c, oy, ox, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = __fadd_rn(acc[k], v[k]);
    sample_up2(b5h, b5l, b, H / 2, W / 2, C8, c, oy, ox, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = __fadd_rn(acc[k], v[k]);
    store8(oh, ol, i, acc);
  }
}

int launch_decoder_merge_split(const bf16* const* a, const bf16* const* b3, const bf16* const* b4, const bf16* const* b5,
                               bf16* const* out, int B, int H, int W, int C, cudaStream_t s) {
  if (C % 8 || H % 2 || W % 2) { set_error("decoder_merge: bad shape"); return -1; }
  const long long total = (long long)B * H * W * (C / 8);
  decoder_merge_split_kernel<<<grid_for(total, 256), 256, 0, s>>>(
      (const uint4*)a[0], (const uint4*)a[1], (const uint4*)b3[0], (const uint4*)b3[1], (const uint4*)b4[0], (const uint4*)b4[1],
      (const uint4*)b5[0], (const uint4*)b5[1], (uint4*)out[0], (uint4*)out[1], B, H, W, C / 8);
  DPB_CHECK_LAUNCH("decoder_merge_split");
  return 0;
}

// ------------------------------------------------------------------------------------ ROIAlign
// torchvision roi_align (aligned=False, sampling 2) exactly as roi.cu resolves it (same sample coordinates, indices and
// weights: bit-identical to the CPU kernel given the same fp32 features), on x = hi + lo.
struct __align__(16) STap { int lo, hi; float l, h; };
__device__ __forceinline__ STap make_stap(float v, int size) {
  STap t;
  const bool dead = (v < -1.0f) || (v > (float)size);
  if (v <= 0.f) v = 0.f;
  int lo = (int)v, hi;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; }
  else hi = lo + 1;
  t.lo = dead ? -1 : lo; t.hi = hi;
  t.l = __fsub_rn(v, (float)lo);
  t.h = __fsub_rn(1.f, t.l);
  return t;
}

__global__ void __launch_bounds__(256) roi_align_split_kernel(RoiAlignArgs a, RoiAlignSplit sp) {
  __shared__ STap s_ty[64], s_tx[64];
  const int r = blockIdx.x;
  if (a.n_rois != nullptr && r >= *a.n_rois) return;
  const float* roi = a.rois + (long long)r * 5;
  const int b = (int)roi[0];
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  int lvl = 0;
  if (a.n_levels > 1) {       // poolers.py:43-51
    const float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    const float v = __fadd_rn(__fdiv_rn(sqrtf(area), 224.f), 1e-8f);
    float l = floorf(__fadd_rn(4.f, log2f(v)));
    l = fminf(fmaxf(l, 2.f), 5.f);
    lvl = (int)l - 2;
    if (lvl >= a.n_levels) lvl = a.n_levels - 1;
  }
  int H = a.H[0], W = a.W[0];                 // constant-index selection: no local-memory copy of the argument structs
  float scale = a.scale[0];
  const bf16* fbase = a.feat[0];
  const bf16* fbase_lo = sp.feat_lo[0];
#pragma unroll
  for (int l = 1; l < 4; ++l)
    if (lvl == l) { H = a.H[l]; W = a.W[l]; scale = a.scale[l]; fbase = a.feat[l]; fbase_lo = sp.feat_lo[l]; }
  const int C8 = a.C / 8;
  const uint4* fh = reinterpret_cast<const uint4*>(fbase) + (long long)b * H * W * C8;
  const uint4* fl = reinterpret_cast<const uint4*>(fbase_lo) + (long long)b * H * W * C8;
  const int P = a.P;
  {
    const float fx0 = __fmul_rn(x1, scale), fy0 = __fmul_rn(y1, scale);
    const float rw = fmaxf(__fsub_rn(__fmul_rn(x2, scale), fx0), 1.f);
    const float rh = fmaxf(__fsub_rn(__fmul_rn(y2, scale), fy0), 1.f);
    const float bw = __fdiv_rn(rw, (float)P), bh = __fdiv_rn(rh, (float)P);
    const int t = threadIdx.x;
    if (t < 2 * P) {
      const int ph = t >> 1, iy = t & 1;
      const float yy = __fadd_rn(__fadd_rn(fy0, __fmul_rn((float)ph, bh)), __fdiv_rn(__fmul_rn((float)iy + 0.5f, bh), 2.f));
      s_ty[t] = make_stap(yy, H);
    } else if (t >= 128 && t < 128 + 2 * P) {
      const int u = t - 128, pw = u >> 1, ix = u & 1;
      const float xx = __fadd_rn(__fadd_rn(fx0, __fmul_rn((float)pw, bw)), __fdiv_rn(__fmul_rn((float)ix + 0.5f, bw), 2.f));
      s_tx[u] = make_stap(xx, W);
    }
  }
  __syncthreads();
  const int chunk = threadIdx.x % C8;
  const int group = threadIdx.x / C8, groups = blockDim.x / C8;
  for (int bin = group; bin < P * P; bin += groups) {
    const int ph = bin / P, pw = bin - ph * P;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int iy = 0; iy < 2; ++iy) {
      const STap ty = s_ty[2 * ph + iy];
      for (int ix = 0; ix < 2; ++ix) {
        const STap tx = s_tx[2 * pw + ix];
        if (ty.lo < 0 || tx.lo < 0) continue;
        const float w1 = __fmul_rn(ty.h, tx.h), w2 = __fmul_rn(ty.h, tx.l), w3 = __fmul_rn(ty.l, tx.h), w4 = __fmul_rn(ty.l, tx.l);
        float q1[8], q2[8], q3[8], q4[8];
        load8(fh, fl, (long long)(ty.lo * W + tx.lo) * C8 + chunk, q1);
        load8(fh, fl, (long long)(ty.lo * W + tx.hi) * C8 + chunk, q2);
        load8(fh, fl, (long long)(ty.hi * W + tx.lo) * C8 + chunk, q3);
        load8(fh, fl, (long long)(ty.hi * W + tx.hi) * C8 + chunk, q4);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float sum = __fadd_rn(__fmul_rn(w1, q1[k]), __fmul_rn(w2, q2[k]));
          sum = __fadd_rn(sum, __fmul_rn(w3, q3[k]));
          sum = __fadd_rn(sum, __fmul_rn(w4, q4[k]));
          acc[k] = __fadd_rn(acc[k], sum);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = __fmul_rn(acc[k], 0.25f);     // / count (= 4), exact
    store8(reinterpret_cast<uint4*>(a.out), reinterpret_cast<uint4*>(sp.out_lo), ((long long)r * P * P + bin) * C8 + chunk, acc);
  }
}

int launch_roi_align_split(const RoiAlignArgs& a, const RoiAlignSplit& sp, cudaStream_t s) {
  if (a.C % 8 || 256 % (a.C / 8)) { set_error("roi_align: C/8 must divide 256"); return -1; }
  if (a.P > 32) { set_error("roi_align: pooler resolution %d > 32", a.P); return -1; }
  if (a.out_fp32) { set_error("roi_align (strict): the output is a bf16 hi/lo pair"); return -1; }
  if (a.R == 0) return 0;
  roi_align_split_kernel<<<a.R, 256, 0, s>>>(a, sp);
  DPB_CHECK_LAUNCH("roi_align_split");
  return 0;
}

// ------------------------------------------------------------------------------------ GroupNorm + ReLU, avg pool (deeplab.py)
// One CTA per ROI (two passes over x = hi + lo; fp64 group sums), like groupnorm_relu_kernel of elementwise.cu.
__global__ void __launch_bounds__(512)
groupnorm_relu_split_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl, const float* __restrict__ gamma,
                            const float* __restrict__ beta, bf16* __restrict__ yh, bf16* __restrict__ yl, int HW, int C,
                            int y_cstride, int out_hw, const int* __restrict__ n_valid) {
  const int r = blockIdx.x;
  if (n_valid != nullptr && r >= *n_valid) return;
  __shared__ double s_sum[32], s_sq[32];
  __shared__ float s_mean[32], s_rstd[32];
  if (threadIdx.x < 32) { s_sum[threadIdx.x] = 0.0; s_sq[threadIdx.x] = 0.0; }
  __syncthreads();
  const int C8 = C / 8, cpg = C / 32;
  const long long base = (long long)r * HW * C8;
  const int chunk = threadIdx.x % C8;
  const int pstart = threadIdx.x / C8, pstep = blockDim.x / C8;
  double sum = 0.0, sq = 0.0;
  for (int p = pstart; p < HW; p += pstep) {
    float f[8];
    load8(xh, xl, base + (long long)p * C8 + chunk, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { sum += (double)f[i]; sq += (double)f[i] * (double)f[i]; }
  }
  const int grp = (chunk * 8) / cpg;
  atomicAdd(&s_sum[grp], sum);
  atomicAdd(&s_sq[grp], sq);
  __syncthreads();
  if (threadIdx.x < 32) {
    const double n = (double)HW * cpg;
    const double m = s_sum[threadIdx.x] / n;
    double var = s_sq[threadIdx.x] / n - m * m;
    if (var < 0) var = 0;
    s_mean[threadIdx.x] = (float)m;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  float g[8], bt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { g[i] = gamma[chunk * 8 + i]; bt[i] = beta[chunk * 8 + i]; }
  const float mean = s_mean[grp], rstd = s_rstd[grp];
  auto put = [&](int q, const float* f) {
    const long long o = ((long long)r * out_hw + q) * y_cstride + chunk * 8;      // elements; 16-byte aligned
    store8(reinterpret_cast<uint4*>(yh + o), reinterpret_cast<uint4*>(yl + o), 0, f);
  };
  if (out_hw == HW) {
    for (int p = pstart; p < HW; p += pstep) {
      float f[8];
      load8(xh, xl, base + (long long)p * C8 + chunk, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf((f[i] - mean) * rstd * g[i] + bt[i], 0.f);
      put(p, f);
    }
  } else {      // HW == 1: normalise the single pixel and broadcast it (bilinear from 1x1, deeplab.py:109)
    float f[8];
    load8(xh, xl, base + chunk, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaxf((f[i] - mean) * rstd * g[i] + bt[i], 0.f);
    for (int p = pstart; p < out_hw; p += pstep) put(p, f);
  }
}

int launch_groupnorm_relu_split(const bf16* xh, const bf16* xl, const float* gamma, const float* beta, bf16* yh, bf16* yl,
                                int R, int HW, int C, int y_cstride, int out_hw, const int* n_valid, cudaStream_t s) {
  if (C % 256 != 0 || C > 512 * 8 || (out_hw != HW && HW != 1) || 512 % (C / 8)) { set_error("groupnorm: bad shape"); return -1; }
  if (R == 0) return 0;
  groupnorm_relu_split_kernel<<<R, 512, 0, s>>>((const uint4*)xh, (const uint4*)xl, gamma, beta, yh, yl, HW, C, y_cstride, out_hw, n_valid);
  DPB_CHECK_LAUNCH("groupnorm_relu_split");
  return 0;
}

__global__ void avgpool_split_kernel(const uint4* __restrict__ xh, const uint4* __restrict__ xl, uint4* __restrict__ yh,
                                     uint4* __restrict__ yl, int HW, int C8, const int* __restrict__ n_valid) {
  const int r = blockIdx.x;
  if (n_valid != nullptr && r >= *n_valid) return;
  // one thread per 8-channel chunk walks all pixels in order: deterministic fp32 sums
  for (int chunk = threadIdx.x; chunk < C8; chunk += blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < HW; ++p) {
      float f[8];
      load8(xh, xl, ((long long)r * HW + p) * C8 + chunk, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = acc[i] / (float)HW;
    store8(yh, yl, (long long)r * C8 + chunk, acc);
  }
}

int launch_avgpool_split(const bf16* xh, const bf16* xl, bf16* yh, bf16* yl, int R, int HW, int C, const int* n_valid, cudaStream_t s) {
  if (C % 8) { set_error("avgpool: bad C"); return -1; }
  if (R == 0) return 0;
  avgpool_split_kernel<<<R, 64, 0, s>>>((const uint4*)xh, (const uint4*)xl, (uint4*)yh, (uint4*)yl, HW, C / 8, n_valid);
  DPB_CHECK_LAUNCH("avgpool_split");
  return 0;
}

}  // namespace dpb
