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
// Float images: ATen upsample_bilinear2d (align_corners=False, scale_factor given): src = s*(dst+0.5)-0.5, s = 1/k.
// The installed ATen CPU build contracts that expression into one FMA and evaluates the taps as
//   top = fma(lx0, p00, lx1*p01), bot = fma(lx0, p10, lx1*p11), out = fma(ly0, top, ly1*bot)     (variant 0)
// in its separable generic kernel (what a multi-threaded reference runs for the HWC->CHW permuted image); with ONE
// intra-op thread ATen picks its channels-last kernel for C == 3, whose scalar loop is
//   s = fma(w00, p00, w01*p01); s = fma(w10, p10, s); s = fma(w11, p11, s),  wij = lyi*lxj            (variant 1)
// (probed on torch 2.11 and restated in oracle/aten_interp.py; both reproduced bit for bit here).
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1,
                                          float& l0, float& l1) {
  float real = fmaf(scale, (float)dst + 0.5f, -0.5f);
  if (real < 0.f) real = 0.f;
  i0 = (int)real;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
  l0 = __fsub_rn(1.f, l1);
}

// uint8 images stay uint8 through the reference's resize (run.py:33-36 feeds torch.from_numpy(cv2 image);
// defaults.py:89). ATen resizes uint8 with the Pillow-style fixed-point scheme: per axis, double-precision triangle
// weights around center = s*(i+0.5) are normalised and quantised to int16 with the largest precision p (<= 22) that
// keeps the largest weight below 2^15; the horizontal pass rounds to uint8 ((sum + 2^(p-1)) >> p), then the vertical
// pass does the same on the rounded rows. One CTA per axis builds the table: entry i = (first tap index, w0 | w1 << 16),
// tab[0] = (precision of the y axis, precision of the x axis), y entries from tab[1], x entries from tab[1 + Hr].
__global__ void __launch_bounds__(256)
u8_resize_tables_kernel(int2* __restrict__ tab, int in_h, int out_h, int in_w, int out_w, double scale) {
  const int axis = blockIdx.x;
  const int in_size = axis ? in_w : in_h, out_size = axis ? out_w : out_h;
  int2* t = tab + 1 + (axis ? out_h : 0);
  __shared__ double s_max[256];
  __shared__ int s_prec;
  auto weights = [&](int i, long long& xmin, double& w0, double& w1) {
    const double center = __dmul_rn(scale, (double)i + 0.5);
    long long lo = (long long)__dadd_rn(__dsub_rn(center, 1.0), 0.5);
    if (lo < 0) lo = 0;
    long long hi = (long long)__dadd_rn(__dadd_rn(center, 1.0), 0.5);
    if (hi > in_size) hi = in_size;
    long long n = hi - lo;
    n = n < 0 ? 0 : (n > 2 ? 2 : n);
    double w[2] = {0.0, 0.0}, total = 0.0;
    for (int j = 0; j < (int)n; ++j) {
      const double x = fabs(__dadd_rn(__dsub_rn((double)(j + lo), center), 0.5));
      w[j] = x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
      total = __dadd_rn(total, w[j]);
    }
    if (total != 0.0) { w[0] = __ddiv_rn(w[0], total); w[1] = __ddiv_rn(w[1], total); }
    xmin = lo; w0 = w[0]; w1 = n > 1 ? w[1] : 0.0;
  };
  double mx = 0.0;
  for (int i = threadIdx.x; i < out_size; i += blockDim.x) {
    long long xmin; double w0, w1;
    weights(i, xmin, w0, w1);
    mx = fmax(mx, fmax(w0, w1));
  }
  s_max[threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) m = fmax(m, s_max[i]);
    int p = 0;
    for (; p < 22; ++p) {
      const int next_value = (int)__dadd_rn(0.5, __dmul_rn(m, (double)(1 << (p + 1))));
      if (next_value >= (1 << 15)) break;
    }
    s_prec = p;
    if (axis) tab[0].y = p; else tab[0].x = p;
  }
  __syncthreads();
  const double unit = (double)(1 << s_prec);
  for (int i = threadIdx.x; i < out_size; i += blockDim.x) {
    long long xmin; double w0, w1;
    weights(i, xmin, w0, w1);
    const int q0 = (int)__dadd_rn(0.5, __dmul_rn(w0, unit)), q1 = (int)__dadd_rn(0.5, __dmul_rn(w1, unit));   // weights are >= 0
    t[i] = make_int2((int)xmin, (q0 & 0xffff) | (q1 << 16));
  }
}

template <typename T>
__global__ void preprocess_kernel(PreprocessArgs a) {
  // one thread per full-resolution pixel slot of the space-to-depth layout: item = ((b, Y, Xc), sub = dy*2+dx);
  // grid = (slots of one row, Hp/2 rows, B images): no 64-bit division per item (28 % of the stall samples before)
  const int Hq = a.Hp / 2;
  const T* src = reinterpret_cast<const T*>(a.src);
  int prec_y = 0, prec_x = 0;
  if (sizeof(T) == 1) { const int2 hdr = a.tables[0]; prec_y = hdr.x; prec_x = hdr.y; }
  const int col = blockIdx.x * blockDim.x + threadIdx.x;      // xc * 4 + sub
  if (col < a.Wx * 4) {
    const int yq = blockIdx.y, b = blockIdx.z;
    const long long i = ((long long)b * Hq + yq) * (a.Wx * 4) + col;
    const int sub = col & 3;
    const int xc = col >> 2;
    const int x = (xc - 2) * 2 + (sub & 1);
    const int y = yq * 2 + (sub >> 1);
    float v[3] = {0.f, 0.f, 0.f};
    if (x >= 0 && x < a.Wr && y < a.Hr) {
      const T* base = src + (long long)b * a.H0 * a.W0 * 3;
      if (sizeof(T) == 1) {
        const int2 ty = a.tables[1 + y], tx = a.tables[1 + a.Hr + x];
        const int y0 = ty.x, y1 = min(y0 + 1, a.H0 - 1), x0 = tx.x, x1 = min(x0 + 1, a.W0 - 1);
        const int wy0 = ty.y & 0xffff, wy1 = ty.y >> 16, wx0 = tx.y & 0xffff, wx1 = tx.y >> 16;
        const T* p00 = base + ((long long)y0 * a.W0 + x0) * 3;
        const T* p01 = base + ((long long)y0 * a.W0 + x1) * 3;
        const T* p10 = base + ((long long)y1 * a.W0 + x0) * 3;
        const T* p11 = base + ((long long)y1 * a.W0 + x1) * 3;
        const int hx = 1 << (prec_x - 1), hy = 1 << (prec_y - 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int cs = a.flip_rgb ? 2 - c : c;
          const int top = min(((int)p00[cs] * wx0 + (int)p01[cs] * wx1 + hx) >> prec_x, 255);
          const int bot = min(((int)p10[cs] * wx0 + (int)p11[cs] * wx1 + hx) >> prec_x, 255);
          const int r = min((top * wy0 + bot * wy1 + hy) >> prec_y, 255);
          v[c] = __fdiv_rn(__fsub_rn((float)r, a.mean[c]), a.std[c]);
        }
      } else {
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        src_index(a.inv_scale, y, a.H0, y0, y1, ly0, ly1);
        src_index(a.inv_scale, x, a.W0, x0, x1, lx0, lx1);
        const T* p00 = base + ((long long)y0 * a.W0 + x0) * 3;
        const T* p01 = base + ((long long)y0 * a.W0 + x1) * 3;
        const T* p10 = base + ((long long)y1 * a.W0 + x0) * 3;
        const T* p11 = base + ((long long)y1 * a.W0 + x1) * 3;
        const float w00 = __fmul_rn(ly0, lx0), w01 = __fmul_rn(ly0, lx1), w10 = __fmul_rn(ly1, lx0), w11 = __fmul_rn(ly1, lx1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int cs = a.flip_rgb ? 2 - c : c;
          float r;
          if (a.variant == 0) {
            const float top = fmaf(lx0, (float)p00[cs], __fmul_rn(lx1, (float)p01[cs]));
            const float bot = fmaf(lx0, (float)p10[cs], __fmul_rn(lx1, (float)p11[cs]));
            r = fmaf(ly0, top, __fmul_rn(ly1, bot));
          } else {
            r = fmaf(w00, (float)p00[cs], __fmul_rn(w01, (float)p01[cs]));
            r = fmaf(w10, (float)p10[cs], r);
            r = fmaf(w11, (float)p11[cs], r);
          }
          v[c] = __fdiv_rn(__fsub_rn(r, a.mean[c]), a.std[c]);
        }
      }
    }
    uint2 o;
    o.x = pack_bf16(v[0], v[1]);
    o.y = pack_bf16(v[2], 0.f);
    reinterpret_cast<uint2*>(a.dst)[i] = o;
    if (a.dst_lo != nullptr) {      // strict mode: x = hi + lo, lo = bf16(x - hi)
      uint2 l;
      l.x = pack_bf16(__fsub_rn(v[0], bf16_lo(o.x)), __fsub_rn(v[1], bf16_hi(o.x)));
      l.y = pack_bf16(__fsub_rn(v[2], bf16_lo(o.y)), 0.f);
      reinterpret_cast<uint2*>(a.dst_lo)[i] = l;
    }
  }
}

int launch_u8_resize_tables(int2* tab, int H0, int Hr, int W0, int Wr, double scale, cudaStream_t s) {
  u8_resize_tables_kernel<<<2, 256, 0, s>>>(tab, H0, Hr, W0, Wr, scale);
  DPB_CHECK_LAUNCH("u8_resize_tables");
  return 0;
}

int launch_preprocess(const PreprocessArgs& a, cudaStream_t s) {
  if (a.Hp % 2) { set_error("preprocess: padded height must be even"); return -1; }
  if (a.src_u8 && a.tables == nullptr) { set_error("preprocess: uint8 input needs the resize tables (dpb200_u8_resize_tables)"); return -1; }
  if (a.Hp / 2 > 65535 || a.B > 65535) { set_error("preprocess: image too tall / batch too large for the launch grid"); return -1; }
  const dim3 g((a.Wx * 4 + 255) / 256, a.Hp / 2, a.B);
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
  // grid = (16-byte chunks of one output row, Ho rows, B images): 32-bit index math only
  const int col = blockIdx.x * blockDim.x + threadIdx.x;      // ox * C8 + c
  if (col >= Wo * C8) return;
  const int oy = blockIdx.y, b = blockIdx.z;
  const int ox = col / C8, c = col - ox * C8;
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
  y[((long long)b * Ho + oy) * ((long long)Wo * C8) + col] = m;
}

int launch_maxpool3x3s2(const bf16* x, bf16* y, int B, int H, int W, int C, cudaStream_t s) {
  if (C % 8) { set_error("maxpool: C %% 8 != 0"); return -1; }
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (Ho > 65535 || B > 65535) { set_error("maxpool: too tall / batch too large for the launch grid"); return -1; }
  const dim3 g((Wo * (C / 8) + 255) / 256, Ho, B);
  maxpool3x3s2_kernel<<<g, 256, 0, s>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), B, H, W, C / 8, Ho, Wo);
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

// The 2x2 output block (2*by + {0,1}, 2*bx + {0,1}) of a x2 bilinear upsample (align_corners=False) samples the 3x3
// half-resolution pixels around (by, bx): output 2*by mixes rows (by-1, by) with weights (.25, .75), output 2*by+1 rows
// (by, by+1) with (.75, .25), the same along x. One thread therefore loads 9 pixels for 4 outputs instead of 16, and
// unpacks each bf16 pair once; every output is computed by the same lerp4 expression on the same operands as
// add_up2 (bit-identical results). Interior blocks only (by >= 1, bx >= 1): the first block row / column clamps at
// the border and goes through add_up2.
__device__ __forceinline__ void add_up2_block(const uint4* __restrict__ x, int b, int h, int w, int C8, int c, int by,
                                              int bx, uint64_t (*acc)[4]) {
  const int r0 = by - 1, r1 = by, r2 = by + 1 < h ? by + 1 : h - 1;
  const int c0 = bx - 1, c1 = bx, c2 = bx + 1 < w ? bx + 1 : w - 1;
  const uint4* base = x + (long long)b * h * w * C8 + c;
  uint4 v[3][3];
  const int rr[3] = {r0, r1, r2}, cc[3] = {c0, c1, c2};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) v[i][j] = __ldg(base + ((long long)rr[i] * w + cc[j]) * C8);
  const uint64_t Q = pack_f32x2(__float_as_uint(0.25f), __float_as_uint(0.25f));
  const uint64_t T = pack_f32x2(__float_as_uint(0.75f), __float_as_uint(0.75f));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint64_t f[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint32_t q = k == 0 ? v[i][j].x : k == 1 ? v[i][j].y : k == 2 ? v[i][j].z : v[i][j].w;
        f[i][j] = bf16x2_to_f32x2(q);
      }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const uint64_t HY = dy ? T : Q, LY = dy ? Q : T;        // (hy, ly): even row (.25, .75), odd row (.75, .25)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const uint64_t HX = dx ? T : Q, LX = dx ? Q : T;
        const uint64_t top = fma_f32x2(LX, f[dy][dx + 1], mul_f32x2(HX, f[dy][dx]));
        const uint64_t bot = fma_f32x2(LX, f[dy + 1][dx + 1], mul_f32x2(HX, f[dy + 1][dx]));
        acc[dy * 2 + dx][k] = add_f32x2(acc[dy * 2 + dx][k], fma_f32x2(LY, bot, mul_f32x2(HY, top)));
      }
    }
  }
}

// Both kernels below walk the output in 8x8-pixel tiles (one CTA per tile; a work item is one 2x2 output block x one
// 16-byte channel chunk), so the 5x5 half-resolution pixels a tile samples are fetched from L2 once and then hit in
// L1 instead of being re-read by CTAs that sit a full image row apart.
__global__ void __launch_bounds__(256)
upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int C8) {
  const int Ho = 2 * H, Wo = 2 * W;
  const int tiles_x = (Wo + 7) / 8, tiles_y = (Ho + 7) / 8;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int t = blockIdx.x - b * tiles_x * tiles_y;
  const int ty0 = (t / tiles_x) * 8, tx0 = (t % tiles_x) * 8;
  for (int item = threadIdx.x; item < 16 * C8; item += blockDim.x) {
    const int c = item % C8, blk = item / C8;
    const int by = (ty0 >> 1) + (blk >> 2), bx = (tx0 >> 1) + (blk & 3);
    if (by >= H || bx >= W) continue;
    uint64_t acc[4][4] = {};
    if (by >= 1 && bx >= 1) {
      add_up2_block(x, b, H, W, C8, c, by, bx, acc);
    } else {
#pragma unroll
      for (int o = 0; o < 4; ++o) add_up2(x, b, H, W, C8, c, 2 * by + (o >> 1), 2 * bx + (o & 1), acc[o]);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
      y[(((long long)b * Ho + 2 * by + (o >> 1)) * Wo + 2 * bx + (o & 1)) * C8 + c] = pack8_f32x2(acc[o]);
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

#ifndef DPB_MERGE_MINB
#define DPB_MERGE_MINB 3
#endif
__global__ void __launch_bounds__(256, DPB_MERGE_MINB)
decoder_merge_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b3, const uint4* __restrict__ b4,
                     const uint4* __restrict__ b5, uint4* __restrict__ out, int B, int H, int W, int C8) {
  const int tiles_x = (W + 7) / 8, tiles_y = (H + 7) / 8;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int t = blockIdx.x - b * tiles_x * tiles_y;
  const int ty0 = (t / tiles_x) * 8, tx0 = (t % tiles_x) * 8;
  const int h = H / 2, w = W / 2;
  for (int item = threadIdx.x; item < 16 * C8; item += blockDim.x) {
    const int c = item % C8, blk = item / C8;
    const int by = (ty0 >> 1) + (blk >> 2), bx = (tx0 >> 1) + (blk & 3);
    if (by >= h || bx >= w) continue;
    uint64_t acc[4][4];
    long long idx[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      idx[o] = (((long long)b * H + 2 * by + (o >> 1)) * W + 2 * bx + (o & 1)) * C8 + c;
      const uint4 a0 = __ldg(a + idx[o]);
      acc[o][0] = bf16x2_to_f32x2(a0.x); acc[o][1] = bf16x2_to_f32x2(a0.y);
      acc[o][2] = bf16x2_to_f32x2(a0.z); acc[o][3] = bf16x2_to_f32x2(a0.w);
    }
    // reference order: ((p2 + up(p3)) + up(p4)) + up(p5)   (roi_head.py:73-77)
    const uint4* br[3] = {b3, b4, b5};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (by >= 1 && bx >= 1) {
        add_up2_block(br[j], b, h, w, C8, c, by, bx, acc);
      } else {
#pragma unroll
        for (int o = 0; o < 4; ++o) add_up2(br[j], b, h, w, C8, c, 2 * by + (o >> 1), 2 * bx + (o & 1), acc[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) out[idx[o]] = pack8_f32x2(acc[o]);
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
// interp2d (chart.py:62-74): F.interpolate(scale_factor=2, bilinear, align_corners=False) of the four deconv outputs,
// reading the phase-planar layout the deconv GEMM epilogue writes: low[r][py][px][c][S/2][S/2] holds low-res pixel
// (2*yy+py, 2*xx+px). No shared memory: lanes run along x, so the loads (4 B, neighbouring lanes adjacent) and the
// 16-byte stores are both coalesced; each thread walks kUpRows output-row pairs of one 4-column strip.
// The fp32 results reproduce ATen's CPU kernel bit for bit (oracle/aten_interp.py): output 2m+1 = taps (m, m+1) with
// weights (.75, .25), output 2m = taps (m-1, m) with (.25, .75), edges clamped, and
//   output h + w > 128 (S = 56):  separable, top = fma(lx0, p00, lx1*p01), out = fma(ly0, top, ly1*bot);
//   output h + w <= 128 (S = 28, the legacy heads): ATen's channels-last kernel on wij = lyi*lxj — channels below
//   C - C % 8 of each tensor: s = fma(w11,p11, w10*p10); s = fma(w01,p01,s); s = fma(w00,p00,s); the tail channels:
//   s = fma(w00,p00, w01*p01); s = fma(w10,p10,s); s = fma(w11,p11,s).
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

// l0 * a + l1 * b the way ATen's separable kernel rounds it
__device__ __forceinline__ float lerp_sep(float l0, float a, float l1, float b) { return fmaf(l0, a, __fmul_rn(l1, b)); }
// the channels-last kernel's four-tap sum (vec: the 8-lane vector expression, else the scalar tail)
__device__ __forceinline__ float lerp_cl(bool vec, float w00, float p00, float w01, float p01, float w10, float p10,
                                         float w11, float p11) {
  float s;
  if (vec) {
    s = fmaf(w11, p11, __fmul_rn(w10, p10));
    s = fmaf(w01, p01, s);
    return fmaf(w00, p00, s);
  }
  s = fmaf(w00, p00, __fmul_rn(w01, p01));
  s = fmaf(w10, p10, s);
  return fmaf(w11, p11, s);
}

template <typename OutT, bool kSmall>
__global__ void __launch_bounds__(256)
predictor_upsample_planar_kernel(const float* __restrict__ low, int S, int Cpad, const int* __restrict__ n_valid,
                                 const UpsampleOutputs outs) {
  const int r = blockIdx.y;
  if (n_valid != nullptr && r >= *n_valid) return;
  const int Sh = S >> 1, So = 2 * S;
  const int G = S >> 1;                        // 4-column output strips per row
  const int C = outs.total;
  const int KB = (S + 1 + kUpRows - 1) / kUpRows;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= KB * C * G) return;
  const int g = item % G;
  const int c = (item / G) % C;
  const int kb = item / (G * C);
  // channel c of the fused deconv output belongs to one output tensor (coarse, fine, u, v, then the confidence heads).
  // Fully unrolled with constant indices: indexing the by-value parameter struct with a run-time index would make
  // every thread copy it to local memory (120 B of stack, +28 % DRAM traffic when this kernel first did that).
  OutT* dst = nullptr;
  int cc = 0, nc = 1, c0 = 0;
#pragma unroll
  for (int i = 0; i < kMaxUpsampleOutputs; ++i) {
    if (i < outs.n) {
      const int n_i = outs.ch[i];
      if (c >= c0 && c < c0 + n_i) { dst = reinterpret_cast<OutT*>(outs.dst[i]); cc = c - c0; nc = n_i; }
      c0 += n_i;
    }
  }
  if (dst == nullptr) return;                  // a head the caller does not want
  const bool vec = cc < nc - (nc & 7);
  OutT* plane = dst + ((long long)r * nc + cc) * So * So + 4 * g;
  const long long plane_sz = (long long)Sh * Sh;
  const float* lr = low + ((long long)r * 4 * Cpad + c) * plane_sz;     // + (py*2+px)*Cpad*plane_sz
  // low columns 2g-1, 2g, 2g+1, 2g+2 (clamped): (phase px, column xx) of each
  const int xa = g == 0 ? 0 : g - 1, pa = g == 0 ? 0 : 1;
  const int xd = g == G - 1 ? Sh - 1 : g + 1, pd = g == G - 1 ? 1 : 0;
  // horizontal weights of the strip's four output columns: (l0, l1) on taps (a0,a1), (a1,a2), (a1,a2), (a2,a3)
  const float e0 = g == 0 ? 1.f : 0.25f, e1 = g == 0 ? 0.f : 0.75f;       // output column 0 clamps to the edge
  auto taps = [&](int y, float* t) {
    y = y < 0 ? 0 : (y > S - 1 ? S - 1 : y);
    const float* p0 = lr + (long long)((y & 1) * 2) * Cpad * plane_sz + (long long)(y >> 1) * Sh;
    const float* p1 = p0 + (long long)Cpad * plane_sz;
    t[0] = __ldg((pa ? p1 : p0) + xa); t[1] = __ldg(p0 + g); t[2] = __ldg(p1 + g); t[3] = __ldg((pd ? p1 : p0) + xd);
  };
  auto hrow = [&](const float* t, float* h) {
    h[0] = lerp_sep(e0, t[0], e1, t[1]);
    h[1] = lerp_sep(0.75f, t[1], 0.25f, t[2]);
    h[2] = lerp_sep(0.25f, t[1], 0.75f, t[2]);
    h[3] = lerp_sep(0.75f, t[2], 0.25f, t[3]);
  };
  // one output row from the taps of low rows (lo, hi) with vertical weights (ly0, ly1)
  auto emit = [&](int orow, const float* tl, const float* th, const float* hl, const float* hh, float ly0, float ly1) {
    float o[4];
    if (kSmall) {
      o[0] = lerp_cl(vec, __fmul_rn(ly0, e0), tl[0], __fmul_rn(ly0, e1), tl[1], __fmul_rn(ly1, e0), th[0], __fmul_rn(ly1, e1), th[1]);
      o[1] = lerp_cl(vec, ly0 * 0.75f, tl[1], ly0 * 0.25f, tl[2], ly1 * 0.75f, th[1], ly1 * 0.25f, th[2]);
      o[2] = lerp_cl(vec, ly0 * 0.25f, tl[1], ly0 * 0.75f, tl[2], ly1 * 0.25f, th[1], ly1 * 0.75f, th[2]);
      o[3] = lerp_cl(vec, ly0 * 0.75f, tl[2], ly0 * 0.25f, tl[3], ly1 * 0.75f, th[2], ly1 * 0.25f, th[3]);
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) o[t] = lerp_sep(ly0, hl[t], ly1, hh[t]);
    }
    store4(plane + (long long)orow * So, o[0], o[1], o[2], o[3]);
  };
  const int k0 = kb * kUpRows - 1;
  float tl[4], th[4], hl[4], hh[4];
  taps(k0, tl);
  hrow(tl, hl);
#pragma unroll
  for (int j = 0; j < kUpRows; ++j) {
    const int k = k0 + j;
    if (k > S - 1) break;
    taps(k + 1, th);
    hrow(th, hh);
    if (k >= 0) emit(2 * k + 1, tl, th, hl, hh, 0.75f, 0.25f);                    // row 2k+1: taps (k, k+1), weights (.75, .25)
    if (k < S - 1) emit(2 * k + 2, tl, th, hl, hh, k < 0 ? 1.f : 0.25f, k < 0 ? 0.f : 0.75f);   // row 2k+2; k = -1 -> row 0 = edge row
#pragma unroll
    for (int t = 0; t < 4; ++t) { tl[t] = th[t]; hl[t] = hh[t]; }
  }
}

int stage_kernels_init() { return 0; }

int launch_predictor_upsample(const float* low, int R, int S, int Cpad, const int* n_valid, const UpsampleOutputs& outs,
                              int out_half, cudaStream_t s) {
  int total = 0;
  for (int i = 0; i < outs.n; ++i) total += outs.ch[i];
  if (S % 2 || outs.n < 1 || outs.n > kMaxUpsampleOutputs || total != outs.total || total > Cpad) {
    set_error("predictor_upsample: bad shape S %d Cpad %d channels %d", S, Cpad, total);
    return -1;
  }
  if (R == 0) return 0;
  const int KB = (S + 1 + kUpRows - 1) / kUpRows;
  const int items = KB * total * (S / 2);
  dim3 grid((items + 255) / 256, R);
  const bool small = 4 * S <= 128;      // ATen switches kernels on output h + w = 2S + 2S
#define DPB_UP(T, SM) predictor_upsample_planar_kernel<T, SM><<<grid, 256, 0, s>>>(low, S, Cpad, n_valid, outs)
  if (out_half) { if (small) DPB_UP(__half, true); else DPB_UP(__half, false); }
  else { if (small) DPB_UP(float, true); else DPB_UP(float, false); }
#undef DPB_UP
  DPB_CHECK_LAUNCH("predictor_upsample_planar");
  return 0;
}

}  // namespace dpb
