// HBM-bound elementwise / resampling kernels: preprocess, maxpool, bilinear x2, decoder merge,
// GroupNorm+ReLU, average pool, predictor tail. All NHWC, 16-byte vector accesses, grid-stride.
#include "kernels.cuh"
#include "conv_igemm.cuh"
#include "ptx.cuh"
#include "device_utils.cuh"

#include <cuda_fp16.h>

namespace dpb {

static inline int grid_for(long long work, int threads, int cap = 148 * 16) {
  long long g = (work + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

#define DPB_CHECK_LAUNCH(name)                                                     \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      set_error("%s launch: %s", name, cudaGetErrorString(e__));                   \
      return -4;                                                                   \
    }                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------ preprocess
// ATen upsample_bilinear2d (align_corners=False, scale_factor given): src = s*(dst+0.5)-0.5, s = 1/k.
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1,
                                          float& l0, float& l1) {
  float real = __fsub_rn(__fmul_rn(scale, (float)dst + 0.5f), 0.5f);
  if (real < 0.f) real = 0.f;
  i0 = (int)real;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
  l0 = __fsub_rn(1.f, l1);
}

template <typename T>
__global__ void preprocess_kernel(PreprocessArgs a) {
  // one thread per full-resolution pixel slot of the space-to-depth layout: item = ((b, Y, Xc), sub = dy*2+dx)
  const int Hq = a.Hp / 2;
  const long long total = (long long)a.B * Hq * a.Wx * 4;
  const T* src = reinterpret_cast<const T*>(a.src);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int sub = (int)(i & 3);
    const int xc = (int)((i >> 2) % a.Wx);
    const int yq = (int)(((i >> 2) / a.Wx) % Hq);
    const int b = (int)((i >> 2) / ((long long)a.Wx * Hq));
    const int x = (xc - 2) * 2 + (sub & 1);
    const int y = yq * 2 + (sub >> 1);
    float v[3] = {0.f, 0.f, 0.f};
    if (x >= 0 && x < a.Wr && y < a.Hr) {
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      src_index(a.inv_scale, y, a.H0, y0, y1, ly0, ly1);
      src_index(a.inv_scale, x, a.W0, x0, x1, lx0, lx1);
      const T* base = src + (long long)b * a.H0 * a.W0 * 3;
      const T* p00 = base + ((long long)y0 * a.W0 + x0) * 3;
      const T* p01 = base + ((long long)y0 * a.W0 + x1) * 3;
      const T* p10 = base + ((long long)y1 * a.W0 + x0) * 3;
      const T* p11 = base + ((long long)y1 * a.W0 + x1) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int cs = a.flip_rgb ? 2 - c : c;
        const float top = __fadd_rn(__fmul_rn(lx0, (float)p00[cs]), __fmul_rn(lx1, (float)p01[cs]));
        const float bot = __fadd_rn(__fmul_rn(lx0, (float)p10[cs]), __fmul_rn(lx1, (float)p11[cs]));
        float r = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
        if (sizeof(T) == 1) r = fminf(fmaxf(rintf(r), 0.f), 255.f);  // uint8 images stay uint8 in the reference
        v[c] = __fdiv_rn(__fsub_rn(r, a.mean[c]), a.std[c]);
      }
    }
    uint2 o;
    o.x = pack_bf16(v[0], v[1]);
    o.y = pack_bf16(v[2], 0.f);
    reinterpret_cast<uint2*>(a.dst)[i] = o;
  }
}

int launch_preprocess(const PreprocessArgs& a, cudaStream_t s) {
  if (a.Hp % 2) { set_error("preprocess: padded height must be even"); return -1; }
  const long long total = (long long)a.B * (a.Hp / 2) * a.Wx * 4;
  const int g = grid_for(total, 256);
  if (a.src_u8) preprocess_kernel<unsigned char><<<g, 256, 0, s>>>(a);
  else preprocess_kernel<float><<<g, 256, 0, s>>>(a);
  DPB_CHECK_LAUNCH("preprocess");
  return 0;
}

// ------------------------------------------------------------------------------------ max pool
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void maxpool3x3s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H,
                                    int W, int C8, int Ho, int Wo) {
  const long long total = (long long)B * Ho * Wo * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int ox = (int)((i / C8) % Wo);
    const int oy = (int)((i / ((long long)C8 * Wo)) % Ho);
    const int b = (int)(i / ((long long)C8 * Wo * Ho));
    uint4 m;
    bool first = true;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int iy = oy * 2 + dy;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ix = ox * 2 + dx;
        if (ix < 0 || ix >= W) continue;
        const uint4 v = __ldg(x + (((long long)b * H + iy) * W + ix) * C8 + c);
        if (first) { m = v; first = false; }
        else { m.x = bf16x2_max(m.x, v.x); m.y = bf16x2_max(m.y, v.y); m.z = bf16x2_max(m.z, v.z); m.w = bf16x2_max(m.w, v.w); }
      }
    }
    y[i] = m;
  }
}

int launch_maxpool3x3s2(const bf16* x, bf16* y, int B, int H, int W, int C, cudaStream_t s) {
  if (C % 8) { set_error("maxpool: C %% 8 != 0"); return -1; }
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)B * Ho * Wo * (C / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(x),
                                                          reinterpret_cast<uint4*>(y), B, H, W, C / 8, Ho, Wo);
  DPB_CHECK_LAUNCH("maxpool");
  return 0;
}

// ------------------------------------------------------------------------------------ bilinear x2
__device__ __forceinline__ void up2_index(int dst, int in_size, int& i0, int& i1, float& l1) {
  float real = 0.5f * ((float)dst + 0.5f) - 0.5f;   // exact in fp32
  if (real < 0.f) real = 0.f;
  i0 = (int)real;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = real - (float)i0;
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
  return o;
}

// acc += bilinear sample of the half-resolution tensor `x` ([.., h, w, C8]) at output pixel (oy, ox);
// acc is four packed (lo, hi) fp32 pairs = 8 channels, all arithmetic two channels per instruction
__device__ __forceinline__ uint64_t bf16x2_to_f32x2(uint32_t q) { return pack_f32x2(q << 16, q & 0xffff0000u); }
__device__ __forceinline__ uint64_t lerp4(uint32_t q00, uint32_t q01, uint32_t q10, uint32_t q11, uint64_t HX,
                                          uint64_t LX, uint64_t HY, uint64_t LY) {
  const uint64_t top = fma_f32x2(LX, bf16x2_to_f32x2(q01), mul_f32x2(HX, bf16x2_to_f32x2(q00)));
  const uint64_t bot = fma_f32x2(LX, bf16x2_to_f32x2(q11), mul_f32x2(HX, bf16x2_to_f32x2(q10)));
  return fma_f32x2(LY, bot, mul_f32x2(HY, top));
}
__device__ __forceinline__ void add_up2(const uint4* __restrict__ x, int b, int h, int w, int C8, int c,
                                        int oy, int ox, uint64_t* acc) {
  int y0, y1, x0, x1;
  float ly, lx;
  up2_index(oy, h, y0, y1, ly);
  up2_index(ox, w, x0, x1, lx);
  const uint4* base = x + (long long)b * h * w * C8 + c;
  const uint4 v00 = __ldg(base + ((long long)y0 * w + x0) * C8);
  const uint4 v01 = __ldg(base + ((long long)y0 * w + x1) * C8);
  const uint4 v10 = __ldg(base + ((long long)y1 * w + x0) * C8);
  const uint4 v11 = __ldg(base + ((long long)y1 * w + x1) * C8);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const uint64_t HX = pack_f32x2(__float_as_uint(hx), __float_as_uint(hx)), LX = pack_f32x2(__float_as_uint(lx), __float_as_uint(lx));
  const uint64_t HY = pack_f32x2(__float_as_uint(hy), __float_as_uint(hy)), LY = pack_f32x2(__float_as_uint(ly), __float_as_uint(ly));
  acc[0] = add_f32x2(acc[0], lerp4(v00.x, v01.x, v10.x, v11.x, HX, LX, HY, LY));
  acc[1] = add_f32x2(acc[1], lerp4(v00.y, v01.y, v10.y, v11.y, HX, LX, HY, LY));
  acc[2] = add_f32x2(acc[2], lerp4(v00.z, v01.z, v10.z, v11.z, HX, LX, HY, LY));
  acc[3] = add_f32x2(acc[3], lerp4(v00.w, v01.w, v10.w, v11.w, HX, LX, HY, LY));
}
__device__ __forceinline__ uint4 pack8_f32x2(const uint64_t* acc) {
  uint4 o;
  o.x = cvt_bf16x2(acc[0]); o.y = cvt_bf16x2(acc[1]); o.z = cvt_bf16x2(acc[2]); o.w = cvt_bf16x2(acc[3]);
  return o;
}

// Both kernels below walk the output in 8x8-pixel tiles (one CTA per tile, one warp per pixel and 16-byte
// channel chunk per lane when C = 256), so the 5x5 half-resolution pixels a tile samples are fetched from L2
// once and then hit in L1 instead of being re-read by CTAs that sit a full image row apart.
__global__ void __launch_bounds__(256)
upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int C8) {
  const int Ho = 2 * H, Wo = 2 * W;
  const int tiles_x = (Wo + 7) / 8, tiles_y = (Ho + 7) / 8;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int t = blockIdx.x - b * tiles_x * tiles_y;
  const int ty0 = (t / tiles_x) * 8, tx0 = (t % tiles_x) * 8;
  for (int item = threadIdx.x; item < 64 * C8; item += blockDim.x) {
    const int c = item % C8, pix = item / C8;
    const int oy = ty0 + (pix >> 3), ox = tx0 + (pix & 7);
    if (oy >= Ho || ox >= Wo) continue;
    uint64_t acc[4] = {0, 0, 0, 0};
    add_up2(x, b, H, W, C8, c, oy, ox, acc);
    y[(((long long)b * Ho + oy) * Wo + ox) * C8 + c] = pack8_f32x2(acc);
  }
}

int launch_upsample2x(const bf16* x, bf16* y, int B, int H, int W, int C, cudaStream_t s) {
  if (C % 8) { set_error("upsample2x: C %% 8 != 0"); return -1; }
  const long long blocks = (long long)B * ((2 * H + 7) / 8) * ((2 * W + 7) / 8);
  if (blocks > 0x7fffffffLL) { set_error("upsample2x: too large"); return -1; }
  upsample2x_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(x),
                                                    reinterpret_cast<uint4*>(y), B, H, W, C / 8);
  DPB_CHECK_LAUNCH("upsample2x");
  return 0;
}

__global__ void __launch_bounds__(256)
decoder_merge_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b3, const uint4* __restrict__ b4,
                     const uint4* __restrict__ b5, uint4* __restrict__ out, int B, int H, int W, int C8) {
  const int tiles_x = (W + 7) / 8, tiles_y = (H + 7) / 8;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int t = blockIdx.x - b * tiles_x * tiles_y;
  const int ty0 = (t / tiles_x) * 8, tx0 = (t % tiles_x) * 8;
  for (int item = threadIdx.x; item < 64 * C8; item += blockDim.x) {
    const int c = item % C8, pix = item / C8;
    const int oy = ty0 + (pix >> 3), ox = tx0 + (pix & 7);
    if (oy >= H || ox >= W) continue;
    const long long i = (((long long)b * H + oy) * W + ox) * C8 + c;
    const uint4 a0 = __ldg(a + i);
    uint64_t acc[4] = {bf16x2_to_f32x2(a0.x), bf16x2_to_f32x2(a0.y), bf16x2_to_f32x2(a0.z), bf16x2_to_f32x2(a0.w)};
    // reference order: ((p2 + up(p3)) + up(p4)) + up(p5)   (roi_head.py:73-77)
    add_up2(b3, b, H / 2, W / 2, C8, c, oy, ox, acc);
    add_up2(b4, b, H / 2, W / 2, C8, c, oy, ox, acc);
    add_up2(b5, b, H / 2, W / 2, C8, c, oy, ox, acc);
    out[i] = pack8_f32x2(acc);
  }
}

int launch_decoder_merge(const bf16* a, const bf16* b, const bf16* c, const bf16* d, bf16* out, int B,
                         int H, int W, int C, cudaStream_t s) {
  if (C % 8 || H % 2 || W % 2) { set_error("decoder_merge: bad shape"); return -1; }
  const long long blocks = (long long)B * ((H + 7) / 8) * ((W + 7) / 8);
  if (blocks > 0x7fffffffLL) { set_error("decoder_merge: too large"); return -1; }
  decoder_merge_kernel<<<(unsigned)blocks, 256, 0, s>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b),
      reinterpret_cast<const uint4*>(c), reinterpret_cast<const uint4*>(d),
      reinterpret_cast<uint4*>(out), B, H, W, C / 8);
  DPB_CHECK_LAUNCH("decoder_merge");
  return 0;
}

// ------------------------------------------------------------------------------------ GroupNorm + ReLU
// One CTA per ROI. 32 groups; a 16-byte chunk (8 channels) never straddles a group (C/32 is 8 or 16).
__global__ void __launch_bounds__(512)
groupnorm_relu_kernel(const uint4* __restrict__ x, const float* __restrict__ gamma,
                      const float* __restrict__ beta, bf16* __restrict__ y, int HW, int C,
                      int y_cstride, int out_hw, const int* __restrict__ n_valid) {
  const int r = blockIdx.x;
  if (n_valid != nullptr && r >= *n_valid) return;
  __shared__ double s_sum[32], s_sq[32];
  __shared__ float s_mean[32], s_rstd[32];
  if (threadIdx.x < 32) { s_sum[threadIdx.x] = 0.0; s_sq[threadIdx.x] = 0.0; }
  __syncthreads();
  const int C8 = C / 8;
  const int cpg = C / 32;
  const uint4* xr = x + (long long)r * HW * C8;
  // each thread owns one channel chunk; pixels strided
  const int chunk = threadIdx.x % C8;
  const int pstart = threadIdx.x / C8;
  const int pstep = blockDim.x / C8;
  float sum = 0.f, sq = 0.f;
  // four independent 16-byte loads in flight per thread (a plain loop issues load -> use -> load: one ROI pass is
  // then bound by HBM latency, not bandwidth); missing tail pixels read as zeros, which add nothing
  for (int p = pstart; p < HW; p += 4 * pstep) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = p + j * pstep;
      v[j] = q < HW ? __ldg(xr + (long long)q * C8 + chunk) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
      unpack8(v[j], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { sum += f[i]; sq += f[i] * f[i]; }
    }
  }
  const int grp = (chunk * 8) / cpg;
  atomicAdd(&s_sum[grp], (double)sum);
  atomicAdd(&s_sq[grp], (double)sq);
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
  if (out_hw == HW) {
    for (int p = pstart; p < HW; p += 4 * pstep) {
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = p + j * pstep;
        v[j] = q < HW ? __ldg(xr + (long long)q * C8 + chunk) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = p + j * pstep;
        if (q >= HW) break;
        float f[8];
        unpack8(v[j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaxf((f[i] - mean) * rstd * g[i] + bt[i], 0.f);
        *reinterpret_cast<uint4*>(y + ((long long)r * out_hw + q) * y_cstride + chunk * 8) = pack8(f);
      }
    }
  } else {
    // HW == 1: normalise the single pixel and broadcast it (bilinear from 1x1 is a broadcast)
    float f[8];
    unpack8(__ldg(xr + chunk), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaxf((f[i] - mean) * rstd * g[i] + bt[i], 0.f);
    const uint4 o = pack8(f);
    for (int p = pstart; p < out_hw; p += pstep)
      *reinterpret_cast<uint4*>(y + ((long long)r * out_hw + p) * y_cstride + chunk * 8) = o;
  }
}

// Single-HBM-pass variant for the 28x28 ROI tensors: GroupNorm groups are independent, so one CTA takes one ROI and a
// 64-channel slice (4 or 8 whole groups; 128-byte rows = full cache lines), keeps its 784 x 128 B = 98 KB in shared
// memory, reduces the group statistics locally (fixed order: deterministic) and writes the normalised slice: one read
// and one write of the tensor instead of two reads (with hundreds of ROIs in flight the second read misses L2) and one
// write. Two CTAs share an SM, so one's loads overlap the other's stores.
static constexpr int kGnSliceC = 64;
__global__ void __launch_bounds__(512, 2)
groupnorm_relu_slice_kernel(const uint4* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                            bf16* __restrict__ y, int HW, int C, int y_cstride, const int* __restrict__ n_valid) {
  extern __shared__ __align__(16) uint4 gn_tile[];            // [HW][8] uint4
  __shared__ double s_part[16][8][2];                         // per warp, per 8-channel chunk: sum, sum of squares
  __shared__ float s_mean[8], s_rstd[8];                      // per chunk (chunks of one group hold the same values)
  const int slices = C / kGnSliceC;
  const int r = blockIdx.x / slices, sl = blockIdx.x - r * slices;
  if (n_valid != nullptr && r >= *n_valid) return;
  const int C8 = C / 8, cpg = C / 32;
  const int chunk = threadIdx.x & 7;                          // 16-byte chunk (8 channels) of the slice
  const int prow = threadIdx.x >> 3;                          // 64 pixel rows per pass
  const uint4* xr = x + (long long)r * HW * C8 + sl * 8 + chunk;
  float sum = 0.f, sq = 0.f;
  for (int p = prow; p < HW; p += 4 * 64) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = p + j * 64;
      v[j] = q < HW ? __ldg(xr + (long long)q * C8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = p + j * 64;
      if (q < HW) gn_tile[q * 8 + chunk] = v[j];
      float f[8];
      unpack8(v[j], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { sum += f[i]; sq += f[i] * f[i]; }
    }
  }
  // lanes l, l+8, l+16, l+24 of a warp hold the same chunk: fold them, then one fp64 partial per (warp, chunk)
  sum += __shfl_xor_sync(0xffffffffu, sum, 8);  sq += __shfl_xor_sync(0xffffffffu, sq, 8);
  sum += __shfl_xor_sync(0xffffffffu, sum, 16); sq += __shfl_xor_sync(0xffffffffu, sq, 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 8) { s_part[warp][lane][0] = (double)sum; s_part[warp][lane][1] = (double)sq; }
  __syncthreads();
  if (threadIdx.x < 8) {
    // group of this chunk: cpg/8 consecutive chunks (1 chunk when C = 256, 2 when C = 512)
    const int cpc = cpg / 8;                                  // chunks per group
    const int c0 = (threadIdx.x / cpc) * cpc;
    double ts = 0.0, tq = 0.0;
    for (int c = c0; c < c0 + cpc; ++c)
      for (int w = 0; w < 16; ++w) { ts += s_part[w][c][0]; tq += s_part[w][c][1]; }
    const double n = (double)HW * cpg;
    const double m = ts / n;
    double var = tq / n - m * m;
    if (var < 0) var = 0;
    s_mean[threadIdx.x] = (float)m;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  float g[8], bt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { g[i] = gamma[sl * kGnSliceC + chunk * 8 + i]; bt[i] = beta[sl * kGnSliceC + chunk * 8 + i]; }
  const float mean = s_mean[chunk], rstd = s_rstd[chunk];
  bf16* yr = y + (long long)r * HW * y_cstride + sl * kGnSliceC + chunk * 8;
  for (int p = prow; p < HW; p += 64) {
    float f[8];
    unpack8(gn_tile[p * 8 + chunk], f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaxf((f[i] - mean) * rstd * g[i] + bt[i], 0.f);
    *reinterpret_cast<uint4*>(yr + (long long)p * y_cstride) = pack8(f);
  }
}

int launch_groupnorm_relu(const bf16* x, const float* gamma, const float* beta, bf16* y, int R, int HW,
                          int C, int y_cstride, int out_hw, const int* n_valid, cudaStream_t s) {
  if (C % 256 != 0 || C > 512 * 8 || (out_hw != HW && HW != 1)) { set_error("groupnorm: bad shape"); return -1; }
  const int C8 = C / 8;
  int threads = 512;
  if (threads % C8) { set_error("groupnorm: C/8 must divide 512"); return -1; }
  if (out_hw == HW && C % kGnSliceC == 0 && (C / 32) % 8 == 0 && HW * 128 <= 100 * 1024 && R > 0) {
    const int smem = HW * 128;
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !done[dev]) {
      if (cudaFuncSetAttribute(groupnorm_relu_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) !=
          cudaSuccess) { set_error("groupnorm: smem attribute"); return -3; }
      done[dev] = true;
    }
    groupnorm_relu_slice_kernel<<<R * (C / kGnSliceC), 512, smem, s>>>(reinterpret_cast<const uint4*>(x), gamma, beta, y, HW,
                                                                      C, y_cstride, n_valid);
    DPB_CHECK_LAUNCH("groupnorm_relu_slice");
    return 0;
  }
  groupnorm_relu_kernel<<<R, threads, 0, s>>>(reinterpret_cast<const uint4*>(x), gamma, beta, y, HW, C,
                                              y_cstride, out_hw, n_valid);
  DPB_CHECK_LAUNCH("groupnorm_relu");
  return 0;
}

// ------------------------------------------------------------------------------------ average pool
__global__ void avgpool_kernel(const uint4* __restrict__ x, bf16* __restrict__ y, int HW, int C8,
                               const int* __restrict__ n_valid) {
  const int r = blockIdx.x;
  if (n_valid != nullptr && r >= *n_valid) return;
  extern __shared__ float s_acc[];   // [C8*8]
  for (int i = threadIdx.x; i < C8 * 8; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int chunk = threadIdx.x % C8;
  const int pstart = threadIdx.x / C8, pstep = blockDim.x / C8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = pstart; p < HW; p += pstep) {
    float f[8];
    unpack8(__ldg(x + ((long long)r * HW + p) * C8 + chunk), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += f[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(&s_acc[chunk * 8 + i], acc[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < C8 * 8; i += blockDim.x)
    y[(long long)r * C8 * 8 + i] = __float2bfloat16(s_acc[i] / (float)HW);
}

int launch_avgpool(const bf16* x, bf16* y, int R, int HW, int C, const int* n_valid, cudaStream_t s) {
  if (C % 8 || 256 % (C / 8)) { set_error("avgpool: bad C"); return -1; }
  avgpool_kernel<<<R, 256, C * sizeof(float), s>>>(reinterpret_cast<const uint4*>(x), y, HW, C / 8, n_valid);
  DPB_CHECK_LAUNCH("avgpool");
  return 0;
}

// ------------------------------------------------------------------------------------ predictor tail
// CTA (kb, roi) stages kPairs+1 consecutive low-res rows in shared memory (NHWC, as the deconv GEMM wrote them)
// and writes the 2*kPairs output rows between them. Work item = (channel c, 8 output columns): lanes run
// along c (conflict-free smem reads), every thread emits full 32-byte sectors of an NCHW plane.
// Output row 2k+1 mixes low rows (k, k+1) with weights (.75, .25), row 2k+2 with (.25, .75); k = -1 and
// k = S-1 collapse onto the edge row (ATen upsample_bilinear2d, align_corners=False, chart.py:62-74).
static constexpr int kUpPairs = 3;

__global__ void __launch_bounds__(256)
predictor_upsample_kernel(const float* __restrict__ low, int S, int Cpad, int Kc,
                          const int* __restrict__ n_valid, float* __restrict__ coarse,
                          float* __restrict__ fine, float* __restrict__ u, float* __restrict__ v) {
  const int r = blockIdx.y;
  if (n_valid != nullptr && r >= *n_valid) return;
  const int k0 = (int)blockIdx.x * kUpPairs - 1;           // first pair index handled here
  extern __shared__ __align__(16) float sm[];              // [kUpPairs + 1][S][Cpad]
  const int row_elems = S * Cpad;
  for (int j = 0; j <= kUpPairs; ++j) {
    int ry = k0 + j;
    ry = ry < 0 ? 0 : (ry > S - 1 ? S - 1 : ry);
    const float4* src = reinterpret_cast<const float4*>(low + ((long long)r * S + ry) * row_elems);
    float4* dst = reinterpret_cast<float4*>(sm + j * row_elems);
    for (int i = threadIdx.x; i < row_elems / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int So = 2 * S;
  const int C = Kc + 75;
  const int items = C * (S / 4);
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int g = it / C, c = it - g * C;
    float* dst; int cc, nc;
    if (c < Kc) { dst = coarse; cc = c; nc = Kc; }
    else if (c < Kc + 25) { dst = fine; cc = c - Kc; nc = 25; }
    else if (c < Kc + 50) { dst = u; cc = c - Kc - 25; nc = 25; }
    else { dst = v; cc = c - Kc - 50; nc = 25; }
    float* plane = dst + ((long long)r * nc + cc) * So * So + 8 * g;
    const int xm = 4 * g - 1 < 0 ? 0 : 4 * g - 1;
    const int xp = 4 * g + 4 > S - 1 ? S - 1 : 4 * g + 4;
    // horizontal pass per staged row: h[j][t] = (1-lx_t) * v[x0_t] + lx_t * v[x1_t]
    float a[6], lo[8], hi[8];
    const float lx0 = (g == 0) ? 0.f : 0.75f;               // output column 0 clamps to the edge
    auto hrow = [&](int j, float* h) {
      const float* p = sm + j * row_elems + c;
      a[0] = p[xm * Cpad];
#pragma unroll
      for (int t = 0; t < 4; ++t) a[1 + t] = p[(4 * g + t) * Cpad];
      a[5] = p[xp * Cpad];
      h[0] = (1.f - lx0) * a[0] + lx0 * a[1];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        h[1 + 2 * t] = 0.75f * a[1 + t] + 0.25f * a[2 + t];
        h[2 + 2 * t] = 0.25f * a[1 + t] + 0.75f * a[2 + t];
      }
      h[7] = 0.75f * a[4] + 0.25f * a[5];
    };
    hrow(0, lo);
#pragma unroll
    for (int j = 0; j < kUpPairs; ++j) {
      const int k = k0 + j;
      if (k > S - 1) break;
      hrow(j + 1, hi);
      if (k >= 0) {                                         // row 2k+1: (.75, .25)
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = 0.75f * lo[t] + 0.25f * hi[t];
        float4* q = reinterpret_cast<float4*>(plane + (long long)(2 * k + 1) * So);
        q[0] = make_float4(o[0], o[1], o[2], o[3]);
        q[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (k < S - 1) {                                      // row 2k+2: (.25, .75); k = -1 -> row 0 = edge row
        float o[8];
        const float ly = (k < 0) ? 0.f : 0.75f;
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = (1.f - ly) * lo[t] + ly * hi[t];
        float4* q = reinterpret_cast<float4*>(plane + (long long)(2 * k + 2) * So);
        q[0] = make_float4(o[0], o[1], o[2], o[3]);
        q[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) lo[t] = hi[t];
    }
  }
}

// Same operation on the phase-planar layout the deconv GEMM epilogue writes: low[r][py][px][c][S/2][S/2]
// holds low-res pixel (2*yy+py, 2*xx+px). No shared memory: lanes run along x, so the loads (4 B, neighbouring
// lanes adjacent) and the 16-byte stores are both coalesced; each thread walks kUpRows output-row pairs of
// one 4-column strip and carries the horizontally interpolated row between them.
static constexpr int kUpRows = 10;

// four consecutive output values: fp32 (16-byte store) or fp16 (8-byte store, round to nearest even)
__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store4(__half* p, float a, float b, float c, float d) {
  const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = o;
}

template <typename OutT>
__global__ void __launch_bounds__(256)
predictor_upsample_planar_kernel(const float* __restrict__ low, int S, int Cpad, int Kc,
                                 const int* __restrict__ n_valid, OutT* __restrict__ coarse,
                                 OutT* __restrict__ fine, OutT* __restrict__ u, OutT* __restrict__ v) {
  const int r = blockIdx.y;
  if (n_valid != nullptr && r >= *n_valid) return;
  const int Sh = S >> 1, So = 2 * S;
  const int G = S >> 1;                        // 4-column output strips per row
  const int C = Kc + 75;
  const int KB = (S + 1 + kUpRows - 1) / kUpRows;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= KB * C * G) return;
  const int g = item % G;
  const int c = (item / G) % C;
  const int kb = item / (G * C);
  OutT* dst; int cc, nc;
  if (c < Kc) { dst = coarse; cc = c; nc = Kc; }
  else if (c < Kc + 25) { dst = fine; cc = c - Kc; nc = 25; }
  else if (c < Kc + 50) { dst = u; cc = c - Kc - 25; nc = 25; }
  else { dst = v; cc = c - Kc - 50; nc = 25; }
  OutT* plane = dst + ((long long)r * nc + cc) * So * So + 4 * g;
  const long long plane_sz = (long long)Sh * Sh;
  const float* lr = low + ((long long)r * 4 * Cpad + c) * plane_sz;     // + (py*2+px)*Cpad*plane_sz
  // low columns 2g-1, 2g, 2g+1, 2g+2 (clamped): (phase px, column xx) of each
  const int xa = g == 0 ? 0 : g - 1, pa = g == 0 ? 0 : 1;
  const int xd = g == G - 1 ? Sh - 1 : g + 1, pd = g == G - 1 ? 1 : 0;
  const float lx0 = (g == 0) ? 0.f : 0.75f;    // output column 0 clamps to the edge
  auto hrow = [&](int y, float* h) {
    y = y < 0 ? 0 : (y > S - 1 ? S - 1 : y);
    const float* p0 = lr + (long long)((y & 1) * 2) * Cpad * plane_sz + (long long)(y >> 1) * Sh;
    const float* p1 = p0 + (long long)Cpad * plane_sz;
    const float a0 = __ldg((pa ? p1 : p0) + xa), a1 = __ldg(p0 + g), a2 = __ldg(p1 + g),
                a3 = __ldg((pd ? p1 : p0) + xd);
    h[0] = (1.f - lx0) * a0 + lx0 * a1;
    h[1] = 0.75f * a1 + 0.25f * a2;
    h[2] = 0.25f * a1 + 0.75f * a2;
    h[3] = 0.75f * a2 + 0.25f * a3;
  };
  const int k0 = kb * kUpRows - 1;
  float lo[4], hi[4];
  hrow(k0, lo);
#pragma unroll
  for (int j = 0; j < kUpRows; ++j) {
    const int k = k0 + j;
    if (k > S - 1) break;
    hrow(k + 1, hi);
    if (k >= 0) {
      store4(plane + (long long)(2 * k + 1) * So, 0.75f * lo[0] + 0.25f * hi[0], 0.75f * lo[1] + 0.25f * hi[1],
             0.75f * lo[2] + 0.25f * hi[2], 0.75f * lo[3] + 0.25f * hi[3]);
    }
    if (k < S - 1) {
      const float ly = (k < 0) ? 0.f : 0.75f, hy = 1.f - ly;
      store4(plane + (long long)(2 * k + 2) * So, hy * lo[0] + ly * hi[0], hy * lo[1] + ly * hi[1],
             hy * lo[2] + ly * hi[2], hy * lo[3] + ly * hi[3]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) lo[t] = hi[t];
  }
}

static constexpr size_t kUpMaxSmem = 96 * 1024;
int stage_kernels_init() {
  cudaError_t e = cudaFuncSetAttribute(predictor_upsample_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpMaxSmem);
  if (e != cudaSuccess) { set_error("predictor_upsample smem attr: %s", cudaGetErrorString(e)); return -3; }
  return 0;
}

int launch_predictor_upsample(const float* low, int R, int S, int Cpad, int Kc, const int* n_valid,
                              void* coarse, void* fine, void* u, void* v, int planar, int out_half,
                              cudaStream_t s) {
  if (planar) {
    if (S % 2 || Kc + 75 > Cpad) { set_error("predictor_upsample: bad shape S %d Cpad %d", S, Cpad); return -1; }
    if (R == 0) return 0;
    const int KB = (S + 1 + kUpRows - 1) / kUpRows;
    const int items = KB * (Kc + 75) * (S / 2);
    dim3 grid((items + 255) / 256, R);
    if (out_half)
      predictor_upsample_planar_kernel<__half><<<grid, 256, 0, s>>>(low, S, Cpad, Kc, n_valid, (__half*)coarse,
                                                                  (__half*)fine, (__half*)u, (__half*)v);
    else
      predictor_upsample_planar_kernel<float><<<grid, 256, 0, s>>>(low, S, Cpad, Kc, n_valid, (float*)coarse,
                                                                 (float*)fine, (float*)u, (float*)v);
    DPB_CHECK_LAUNCH("predictor_upsample_planar");
    return 0;
  }
  if (out_half) { set_error("predictor_upsample: fp16 output needs the phase-planar layout"); return -1; }
  if (S % 4 || Cpad % 4 || Kc + 75 > Cpad) { set_error("predictor_upsample: bad shape S %d Cpad %d", S, Cpad); return -1; }
  const size_t smem = (size_t)(kUpPairs + 1) * S * Cpad * sizeof(float);
  if (smem > kUpMaxSmem) { set_error("predictor_upsample: S %d Cpad %d needs %zu B of shared memory", S, Cpad, smem); return -1; }
  static thread_local int init_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (init_dev != dev) {
    if (stage_kernels_init()) return -3;
    init_dev = dev;
  }
  if (R == 0) return 0;
  dim3 grid((S + 1 + kUpPairs - 1) / kUpPairs, R);
  predictor_upsample_kernel<<<grid, 256, smem, s>>>(low, S, Cpad, Kc, n_valid, (float*)coarse, (float*)fine,
                                                    (float*)u, (float*)v);
  DPB_CHECK_LAUNCH("predictor_upsample");
  return 0;
}

}  // namespace dpb
