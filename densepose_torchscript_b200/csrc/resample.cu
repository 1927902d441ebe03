// Per-box DensePose result extraction on the device (replaces the host loop of visualizer.py:10-56):
// bilinear resize of coarse / fine / U / V to the integer box size, 25-way part argmax gated by the
// coarse argmax, and the U/V gather at the winning part. One thread per output pixel.
#include "kernels.cuh"
#include "conv_igemm.cuh"

namespace dpb {

__device__ __forceinline__ void size_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0,
                                           float& l1) {
  // ATen area_pixel_compute_source_index (align_corners=False) + guard_index_and_lambda
  float real = __fsub_rn(__fmul_rn(scale, (float)dst + 0.5f), 0.5f);
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
    auto sample = [&](const float* pl) -> float {
      const float top = __fadd_rn(__fmul_rn(lx0, __ldg(pl + o00)), __fmul_rn(lx1, __ldg(pl + o01)));
      const float bot = __fadd_rn(__fmul_rn(lx0, __ldg(pl + o10)), __fmul_rn(lx1, __ldg(pl + o11)));
      return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    };
    // coarse argmax (first maximum wins, like torch.argmax)
    const float* cbase = a.coarse + (long long)d * a.Kc * plane;
    int carg = 0;
    float cbest = sample(cbase);
    for (int c = 1; c < a.Kc; ++c) {
      const float v = sample(cbase + c * plane);
      if (v > cbest) { cbest = v; carg = c; }
    }
    const float* fbase = a.fine + (long long)d * 25 * plane;
    int farg = 0;
    float fbest = sample(fbase);
    for (int c = 1; c < 25; ++c) {
      const float v = sample(fbase + c * plane);
      if (v > fbest) { fbest = v; farg = c; }
    }
    const int label = (carg > 0) ? farg : 0;
    if (a.labels_u8) reinterpret_cast<unsigned char*>(a.labels)[p] = (unsigned char)label;
    else reinterpret_cast<long long*>(a.labels)[p] = (long long)label;
    float uu = 0.f, vv = 0.f;
    if (label > 0) {
      uu = sample(a.u + ((long long)d * 25 + label) * plane);
      vv = sample(a.v + ((long long)d * 25 + label) * plane);
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
