// Block-level device helpers shared by rpn.cu and roi.cu: order-preserving float keys, bitonic sort,
// block scan, and the greedy NMS on score-sorted boxes (torchvision nms semantics).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpb {

__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// Descending bitonic sort of N (power of two) 64-bit keys in shared memory by the whole block.
static __device__ void bitonic_sort_desc(unsigned long long* keys, int N) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool desc = ((i & k) == 0);
          if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// Block-wide inclusive scan of one unsigned per thread (blockDim.x == 1024).
__device__ __forceinline__ unsigned block_scan_incl(unsigned v, unsigned* warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  if (warp > 0) v += warp_sums[warp - 1];
  __syncthreads();
  return v;
}

// ---------------------------------------------------------------------------------------- NMS
// boxes sorted by descending score; alive[] marks usable entries on entry; keep[] on exit.
// Shared: mask [n][32] u32 (bit j of row i: j > i and IoU(i, j) > thr), boxes [n] float4, areas [n].
static __device__ void nms_sorted_block(const float4* __restrict__ g_boxes, int n, float thr,
                                 unsigned char* __restrict__ keep_io, uint32_t* smem) {
  uint32_t* mask = smem;                                  // n * 32
  float4* boxes = reinterpret_cast<float4*>(smem + 1024 * 32);
  float* areas = reinterpret_cast<float*>(boxes + 1024);
  uint32_t* alive_words = reinterpret_cast<uint32_t*>(areas + 1024);   // 32
  if (threadIdx.x < 32) alive_words[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 bx = g_boxes[i];
    boxes[i] = bx;
    areas[i] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
    if (keep_io[i]) atomicOr(&alive_words[i >> 5], 1u << (i & 31));
  }
  __syncthreads();
  const int nw = (n + 31) >> 5;
  {
    // One warp per (row i, 32-column word w >= i/32): lane = column, the word is a ballot. Box reads are
    // conflict-free (consecutive float4 per lane, row box broadcast from registers).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = warp; i < n; i += nwarps) {
      const float4 bi = boxes[i];
      const float ai = areas[i];
      for (int w = i >> 5; w < nw; ++w) {
        const int j = w * 32 + lane;
        bool sup = false;
        if (j > i && j < n) {
          const float4 bj = boxes[j];
          const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
          const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
          const float w_ = fmaxf(0.f, __fsub_rn(xx2, xx1)), h_ = fmaxf(0.f, __fsub_rn(yy2, yy1));
          const float inter = __fmul_rn(w_, h_);
          const float uni = __fsub_rn(__fadd_rn(ai, areas[j]), inter);
          // torchvision: inter / union > thr with an IEEE division. The quotient is only computed when the
          // multiplied-out comparison is within 4e-6 of the boundary (or thr * union is not a normal positive
          // number); everywhere else the two agree, so the mask is bit-identical at a fraction of the cost.
          const float tu = __fmul_rn(thr, uni);
          const bool normal = tu > 1.0e-30f && tu < 3.0e38f;     // no denormal / overflow / NaN corner
          if (normal && inter > __fmul_rn(tu, 1.000004f)) sup = true;
          else if (normal && inter < __fmul_rn(tu, 0.999996f)) sup = false;
          else sup = __fdiv_rn(inter, uni) > thr;
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, sup);
        if (lane == 0) mask[i * 32 + w] = bits;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    uint32_t removed = ~alive_words[lane];
    uint32_t kept = 0;
    for (int i = 0; i < n; ++i) {
      const uint32_t r = __shfl_sync(0xffffffffu, removed, i >> 5);
      if (!((r >> (i & 31)) & 1u)) {
        if (lane == (i >> 5)) kept |= 1u << (i & 31);
        if (lane >= (i >> 5) && lane < nw) removed |= mask[i * 32 + lane];   // words left of the diagonal are never written
      }
    }
    alive_words[lane] = kept;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    keep_io[i] = (alive_words[i >> 5] >> (i & 31)) & 1u;
}


static constexpr int kNmsSmemBytes = 1024 * 32 * 4 + 1024 * 16 + 1024 * 4 + 32 * 4;

}  // namespace dpb
