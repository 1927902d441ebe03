// Per-box DensePose result extraction on the device (replaces the host loop of visualizer.py:10-56):
// bilinear resize of coarse / fine / U / V to the integer box size, 25-way part argmax gated by the
// coarse argmax, and the U/V gather at the winning part. One thread per output pixel.
#include "kernels.cuh"
#include "conv_igemm.cuh"

namespace dpb {

// ATen's CPU upsample_bilinear2d, reproduced bit for bit (probed on torch 2.11, restated in oracle/aten_interp.py):
//  * source index: one FMA, real = fma(scale, dst + 0.5, -0.5), scale = float(in) / float(out);
//  * output h + w > 128: the separable generic kernel,
//        top = fma(lx0, p00, lx1*p01); bot = fma(lx0, p10, lx1*p11); out = fma(ly0, top, ly1*bot);
//  * output h + w <= 128: the channels-last kernel (wij = lyi*lxj rounded), 8 channels per vector; channels below
//    C - C % 8 take the vector expression, the rest the scalar tail — the two associate differently:
//        vector:  s = fma(w11, p11, w10*p10); s = fma(w01, p01, s); s = fma(w00, p00, s)
//        tail:    s = fma(w00, p00, w01*p01); s = fma(w10, p10, s); s = fma(w11, p11, s)
__device__ __forceinline__ void size_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0,
                                           float& l1) {
  float real = fmaf(scale, (float)dst + 0.5f, -0.5f);
  if (real < 0.f) real = 0.f;
  i0 = (int)real;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = (i0 + 1 < in_size - 1) ? i0 + 1 : in_size - 1;
  l1 = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
  l0 = __fsub_rn(1.f, l1);
}

__global__ void __launch_bounds__(256) dp_resample_kernel(ResampleArgs a) {
  const int S = a.S;
  const long long plane = (long long)S * S;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < a.total_pixels;
       p += (long long)gridDim.x * blockDim.x) {
    // box lookup: offsets is ascending, D <= ~1000
    int lo = 0, hi = a.D;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.offsets[mid] <= p) lo = mid; else hi = mid;
    }
    const int d = lo;
    const int w = a.box_wh[2 * d], h = a.box_wh[2 * d + 1];
    const long long local = p - a.offsets[d];
    const int oy = (int)(local / w), ox = (int)(local - (long long)oy * w);
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    size_index((float)S / (float)h, oy, S, y0, y1, ly0, ly1);
    size_index((float)S / (float)w, ox, S, x0, x1, lx0, lx1);
    const long long o00 = (long long)y0 * S + x0, o01 = (long long)y0 * S + x1;
    const long long o10 = (long long)y1 * S + x0, o11 = (long long)y1 * S + x1;
    const bool small = h + w <= 128;
    const float w00 = __fmul_rn(ly0, lx0), w01 = __fmul_rn(ly0, lx1), w10 = __fmul_rn(ly1, lx0), w11 = __fmul_rn(ly1, lx1);
    // channel c of a tensor with C channels
    auto sample = [&](const float* pl, int c, int C) -> float {
      const float p00 = __ldg(pl + o00), p01 = __ldg(pl + o01), p10 = __ldg(pl + o10), p11 = __ldg(pl + o11);
      if (!small) {
        const float top = fmaf(lx0, p00, __fmul_rn(lx1, p01));
        const float bot = fmaf(lx0, p10, __fmul_rn(lx1, p11));
        return fmaf(ly0, top, __fmul_rn(ly1, bot));
      }
      float s;
      if (c < C - (C & 7)) {
        s = fmaf(w11, p11, __fmul_rn(w10, p10));
        s = fmaf(w01, p01, s);
        return fmaf(w00, p00, s);
      }
      s = fmaf(w00, p00, __fmul_rn(w01, p01));
      s = fmaf(w10, p10, s);
      return fmaf(w11, p11, s);
    };
    // coarse argmax (first maximum wins, like torch.argmax)
    const float* cbase = a.coarse + (long long)d * a.Kc * plane;
    int carg = 0;
    float cbest = sample(cbase, 0, a.Kc);
    for (int c = 1; c < a.Kc; ++c) {
      const float v = sample(cbase + c * plane, c, a.Kc);
      if (v > cbest) { cbest = v; carg = c; }
    }
    const float* fbase = a.fine + (long long)d * 25 * plane;
    int farg = 0;
    float fbest = sample(fbase, 0, 25);
    for (int c = 1; c < 25; ++c) {
      const float v = sample(fbase + c * plane, c, 25);
      if (v > fbest) { fbest = v; farg = c; }
    }
    const int label = (carg > 0) ? farg : 0;
    if (a.labels_u8) reinterpret_cast<unsigned char*>(a.labels)[p] = (unsigned char)label;
    else reinterpret_cast<long long*>(a.labels)[p] = (long long)label;
    float uu = 0.f, vv = 0.f;
    if (label > 0) {
      uu = sample(a.u + ((long long)d * 25 + label) * plane, label, 25);
      vv = sample(a.v + ((long long)d * 25 + label) * plane, label, 25);
    }
    const long long hw = (long long)w * h;
    float* uvb = a.uv + 2 * a.offsets[d];
    uvb[local] = uu;
    uvb[hw + local] = vv;
  }
}

int launch_dp_resample(const ResampleArgs& a, cudaStream_t s) {
  if (a.total_pixels == 0 || a.D == 0) return 0;
  long long g = (a.total_pixels + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  dp_resample_kernel<<<(int)g, 256, 0, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("dp_resample launch: %s", cudaGetErrorString(e)); return -4; }
  return 0;
}

}  // namespace dpb
