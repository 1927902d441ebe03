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

// ---------------------------------------------------------------------------------------- clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address -> the same offset in the shared memory of CTA `rank` of this cluster
__device__ __forceinline__ uint32_t cluster_map_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_cluster_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_cluster_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t atom_add_shared_cluster_u32(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ double ld_shared_cluster_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
// the two halves of a cluster barrier (every thread of every CTA executes both, in this order)
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// all threads of all CTAs of the cluster; release/acquire orders shared::cluster and global accesses
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------- NMS
// Greedy NMS (torchvision semantics) of n <= 1024 boxes sorted by descending score; keep_io[] marks usable
// entries on entry and the kept ones on exit.
//   phase 1  IoU bit-mask: bit j of row i = (j > i and IoU(i, j) > thr), one warp per (row, 32-column word).
//            With NC > 1 the kernel must have been launched as clusters of NC CTAs (pick_nms_cluster): the rows are dealt round-robin over
//            the NC*32 warps of the cluster and every word is written straight into the shared memory of
//            CTA 0 (distributed shared memory), so the O(n^2) part uses NC SMs per problem.
//   phase 2  one warp of CTA 0 resolves the greedy order 32 boxes at a time: the 32x32 diagonal block is
//            gathered into registers and walked as a pure ALU chain, then the rows of the kept boxes are
//            OR-ed into the removed set with independent shared loads.
// Shared (CTA 0): mask [1024][32] u32, boxes [1024] float4, areas [1024], alive words [32].
// Returns true in the CTA that holds the result (CTA 0 of the cluster); the other CTAs are done.
static __device__ bool nms_sorted_block(const float4* __restrict__ g_boxes, int n, float thr,
                                        unsigned char* __restrict__ keep_io, uint32_t* smem, const int NC) {
  uint32_t* mask = smem;                                  // n * 32
  float4* boxes = reinterpret_cast<float4*>(smem + 1024 * 32);
  float* areas = reinterpret_cast<float*>(boxes + 1024);
  uint32_t* alive_words = reinterpret_cast<uint32_t*>(areas + 1024);   // 32
  const uint32_t rank = NC > 1 ? cluster_ctarank() : 0u;
  // distributed shared memory of a CTA may only be touched once that CTA runs: arrive now, wait just before the first
  // remote store of the mask phase (the box loads in between hide the barrier)
  if (NC > 1) cluster_arrive();
  if (threadIdx.x < 32) alive_words[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 bx = g_boxes[i];
    boxes[i] = bx;
    areas[i] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
    if (keep_io[i]) atomicOr(&alive_words[i >> 5], 1u << (i & 31));
  }
  __syncthreads();
  if (NC > 1) cluster_wait();
  const int nw = (n + 31) >> 5;
  {
    // lane = column, the word is a ballot. Box reads are conflict-free (consecutive float4 per lane, row box
    // broadcast from registers).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t mask0 = NC > 1 ? cluster_map_shared((uint32_t)__cvta_generic_to_shared(mask), 0u) : 0u;
    for (int i = (int)rank * nwarps + warp; i < n; i += NC * nwarps) {
      const float4 bi = boxes[i];
      const float ai = areas[i];
      for (int w = i >> 5; w < nw; ++w) {
        const int j = w * 32 + lane;
        const int jc = j < n ? j : n - 1;                      // clamped: every lane computes, invalid ones are masked
        const float4 bj = boxes[jc];
        const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
        const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
        const float w_ = fmaxf(0.f, __fsub_rn(xx2, xx1)), h_ = fmaxf(0.f, __fsub_rn(yy2, yy1));
        const float inter = __fmul_rn(w_, h_);
        const float uni = __fsub_rn(__fadd_rn(ai, areas[jc]), inter);
        // torchvision: inter / union > thr with an IEEE division. The quotient is only computed when the
        // multiplied-out comparison is within 4e-6 of the boundary (or thr * union is not a normal positive
        // number); everywhere else the two agree, so the mask is bit-identical at a fraction of the cost.
        const float tu = __fmul_rn(thr, uni);
        const bool valid = j > i && j < n;
        const bool normal = tu > 1.0e-30f && tu < 3.0e38f;     // no denormal / overflow / NaN corner
        bool sup = normal && inter > __fmul_rn(tu, 1.000004f);
        const bool unsure = valid && !sup && !(normal && inter < __fmul_rn(tu, 0.999996f));
        if (__any_sync(0xffffffffu, unsure)) {                 // warp-uniform and rare
          if (unsure) sup = __fdiv_rn(inter, uni) > thr;
        }
        sup = sup && valid;
        const uint32_t bits = __ballot_sync(0xffffffffu, sup);
        if (lane == 0) {
          if (NC > 1) st_shared_cluster_u32(mask0 + (uint32_t)(i * 32 + w) * 4u, bits);
          else mask[i * 32 + w] = bits;
        }
      }
    }
  }
  if (NC > 1) cluster_sync_all();
  else __syncthreads();
  if (rank != 0) return false;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    uint32_t removed = ~alive_words[lane];      // word `lane` of the removed set (bits past n are set)
    uint32_t kept = 0;
    for (int w = 0; w < nw; ++w) {
      uint32_t rw = __shfl_sync(0xffffffffu, removed, w);          // warp-uniform
      const int row = w * 32 + lane;
      const uint32_t diag = (row < n) ? mask[row * 32 + w] : 0u;    // suppression inside this word (bits j > i)
      // word `lane` of the 32 rows of this block, loaded before the scan needs them: the loads do not depend on
      // which boxes survive, so their latency hides behind the serial chain below (words left of the diagonal
      // are never written: those lanes load nothing)
      const bool later = lane > w && lane < nw;
      uint32_t r[32];
#pragma unroll
      for (int b = 0; b < 32; ++b) r[b] = later ? mask[(w * 32 + b) * 32 + lane] : 0u;
      // greedy order inside the word: box b survives iff its bit is still clear when its turn comes; a row only
      // has bits j > b, so bit b is final after step b and the kept set is the complement of the final word.
      // Two dependent ALU operations per step (test, predicated OR).
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        const uint32_t db = __shfl_sync(0xffffffffu, diag, b);
        if (!(rw & (1u << b))) rw |= db;
      }
      const uint32_t kw = ~rw;
      if (lane == w) kept = kw;
      // rows of the kept boxes suppress later words
      uint32_t acc = 0;
#pragma unroll
      for (int b = 0; b < 32; ++b) acc |= (kw >> b) & 1u ? r[b] : 0u;
      removed |= acc;
    }
    alive_words[lane] = kept;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    keep_io[i] = (alive_words[i >> 5] >> (i & 31)) & 1u;
  return true;
}


// Largest cluster size <= 8 (the portable maximum) that keeps problems * NC CTAs within one wave of the device.
static inline int pick_nms_cluster(int problems, int num_sms) {
  for (int nc = 8; nc > 1; --nc)
    if (problems * nc <= num_sms) return nc;
  return 1;
}

static constexpr int kNmsSmemBytes = 1024 * 32 * 4 + 1024 * 16 + 1024 * 4 + 32 * 4;

}  // namespace dpb
