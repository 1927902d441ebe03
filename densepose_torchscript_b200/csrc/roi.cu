// ROIAlign over the FPN pyramid, and the box-head tail (softmax, decode, threshold, NMS, top-k,
// detector_postprocess) — all on the device with device-side counts, no host sync.
// Reference: poolers.py:15-51,187-227; layers/roi_align.py:49-65 (-> torchvision roi_align, aligned=False);
// fast_rcnn.py:86-140,257-326; postprocessing.py:11-61; structures.py:107-140.
#include "kernels.cuh"
#include "conv_igemm.cuh"
#include "device_utils.cuh"
#include "ptx.cuh"

namespace dpb {

#define DPB_CHECK_LAUNCH(name)                                                     \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      set_error("%s launch: %s", name, cudaGetErrorString(e__));                   \
      return -4;                                                                   \
    }                                                                              \
  } while (0)


// ------------------------------------------------------------------------------------ ROIAlign
// torchvision bilinear pre-calc for one sample coordinate (aligned=False): source rows/cols lo, hi and the
// weights l (towards hi) and h = 1 - l. lo < 0 marks a sample outside [-1, size] (contributes nothing).
struct __align__(16) Tap {
  int lo, hi;
  float l, h;
};
__device__ __forceinline__ Tap make_tap(float v, int size) {
  Tap t;
  const bool dead = (v < -1.0f) || (v > (float)size);
  if (v <= 0.f) v = 0.f;
  int lo = (int)v;
  int hi;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; }
  else hi = lo + 1;
  t.lo = dead ? -1 : lo; t.hi = hi;
  t.l = __fsub_rn(v, (float)lo);
  t.h = __fsub_rn(1.f, t.l);
  return t;
}

// acc += ((w1*v1 + w2*v2) + w3*v3) + w4*v4 for one bf16 pair of each of the four taps; every product and
// sum individually rounded (torchvision's CPU kernel has no FMA contraction). Two channels per instruction:
// the products are mul.rn.f32x2 and every sum is fma.rn.f32x2(p, ONE, s) = round(p * 1 + s) = round(p + s).
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 despite the rounding modifiers; it cannot
// contract into an FMA whose multiplier is a run-time value, so ONE = (1.0f, 1.0f) arrives as a kernel
// argument and the SASS is FMUL2 + FFMA2(p, 1.0, s): bit-identical to separately rounded scalar code.
__device__ __forceinline__ uint64_t prod2(uint64_t w, uint32_t q) {
  return mul_f32x2(w, pack_f32x2(q << 16, q & 0xffff0000u));
}
__device__ __forceinline__ uint64_t tap_accum(uint64_t acc, uint32_t q1, uint32_t q2, uint32_t q3, uint32_t q4,
                                              uint64_t w1, uint64_t w2, uint64_t w3, uint64_t w4, uint64_t one) {
  uint64_t sum = fma_f32x2(prod2(w2, q2), one, prod2(w1, q1));
  sum = fma_f32x2(prod2(w3, q3), one, sum);
  sum = fma_f32x2(prod2(w4, q4), one, sum);
  return fma_f32x2(sum, one, acc);
}

// One block per ROI. The 2P sample coordinates per axis are resolved once into shared-memory tap tables;
// then a group of C/8 threads (16 bytes of channels each) walks the bins: 16 independent 16-byte loads in
// flight per thread, packed f32x2 arithmetic, one 16-byte store.
__global__ void __launch_bounds__(256) roi_align_kernel(RoiAlignArgs a, float one_f) {
  __shared__ Tap s_ty[64], s_tx[64];
  const int r = blockIdx.x;
  if (a.n_rois != nullptr && r >= *a.n_rois) return;
  const float* roi = a.rois + (long long)r * 5;
  const int b = (int)roi[0];
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  int lvl = 0;
  if (a.n_levels > 1) {
    // poolers.py:43-51: floor(4 + log2(sqrt(area)/224 + 1e-8)) clamped to [2, 5]
    const float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    const float v = __fadd_rn(__fdiv_rn(sqrtf(area), 224.f), 1e-8f);
    float l = floorf(__fadd_rn(4.f, log2f(v)));
    l = fminf(fmaxf(l, 2.f), 5.f);
    lvl = (int)l - 2;
    if (lvl >= a.n_levels) lvl = a.n_levels - 1;
  }
  // per-level parameters picked with constant indices: indexing the by-value argument struct with a run-time
  // `lvl` would make every thread copy its arrays to local memory (a 128-byte stack frame in round 1)
  int H = a.H[0], W = a.W[0];
  float scale = a.scale[0];
  const bf16* fbase = a.feat[0];
#pragma unroll
  for (int l = 1; l < 4; ++l)
    if (lvl == l) { H = a.H[l]; W = a.W[l]; scale = a.scale[l]; fbase = a.feat[l]; }
  const int C8 = a.C / 8;
  const uint4* feat = reinterpret_cast<const uint4*>(fbase) + (long long)b * H * W * C8;
  const int P = a.P;
  {
    const float fx0 = __fmul_rn(x1, scale), fy0 = __fmul_rn(y1, scale);
    const float rw = fmaxf(__fsub_rn(__fmul_rn(x2, scale), fx0), 1.f);
    const float rh = fmaxf(__fsub_rn(__fmul_rn(y2, scale), fy0), 1.f);
    const float bw = __fdiv_rn(rw, (float)P), bh = __fdiv_rn(rh, (float)P);
    const int t = threadIdx.x;
    if (t < 2 * P) {            // y samples: index 2*ph + iy
      const int ph = t >> 1, iy = t & 1;
      const float yy = __fadd_rn(__fadd_rn(fy0, __fmul_rn((float)ph, bh)),
                                 __fdiv_rn(__fmul_rn((float)iy + 0.5f, bh), 2.f));
      s_ty[t] = make_tap(yy, H);
    } else if (t >= 128 && t < 128 + 2 * P) {
      const int u = t - 128, pw = u >> 1, ix = u & 1;
      const float xx = __fadd_rn(__fadd_rn(fx0, __fmul_rn((float)pw, bw)),
                                 __fdiv_rn(__fmul_rn((float)ix + 0.5f, bw), 2.f));
      s_tx[u] = make_tap(xx, W);
    }
  }
  __syncthreads();

  const int chunk = threadIdx.x % C8;
  const int group = threadIdx.x / C8, groups = blockDim.x / C8;
  const uint4* fc = feat + chunk;
  const uint64_t one = pack_f32x2(__float_as_uint(one_f), __float_as_uint(one_f));
  // gridDim.y CTAs share one ROI (contiguous slices of its bins): with one CTA per ROI the 800 DensePose ROIs of a
  // batch are 1.35 waves of the 592 resident CTAs, i.e. the second wave runs at a third of the machine
  const int per = (P * P + (int)gridDim.y - 1) / (int)gridDim.y;
  const int bin_end = min(P * P, ((int)blockIdx.y + 1) * per);
  for (int bin = (int)blockIdx.y * per + group; bin < bin_end; bin += groups) {
    const int ph = bin / P, pw = bin - ph * P;
    uint64_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;    // four (+0, +0) pairs = 8 channels
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
      const Tap ty = s_ty[2 * ph + iy];
#pragma unroll
      for (int ix = 0; ix < 2; ++ix) {
        const Tap tx = s_tx[2 * pw + ix];
        if (ty.lo < 0 || tx.lo < 0) continue;
        const float f1 = __fmul_rn(ty.h, tx.h), f2 = __fmul_rn(ty.h, tx.l);
        const float f3 = __fmul_rn(ty.l, tx.h), f4 = __fmul_rn(ty.l, tx.l);
        const uint64_t w1 = pack_f32x2(__float_as_uint(f1), __float_as_uint(f1)), w2 = pack_f32x2(__float_as_uint(f2), __float_as_uint(f2));
        const uint64_t w3 = pack_f32x2(__float_as_uint(f3), __float_as_uint(f3)), w4 = pack_f32x2(__float_as_uint(f4), __float_as_uint(f4));
        const uint4 q1 = __ldg(fc + (ty.lo * W + tx.lo) * C8);
        const uint4 q2 = __ldg(fc + (ty.lo * W + tx.hi) * C8);
        const uint4 q3 = __ldg(fc + (ty.hi * W + tx.lo) * C8);
        const uint4 q4 = __ldg(fc + (ty.hi * W + tx.hi) * C8);
        acc0 = tap_accum(acc0, q1.x, q2.x, q3.x, q4.x, w1, w2, w3, w4, one);
        acc1 = tap_accum(acc1, q1.y, q2.y, q3.y, q4.y, w1, w2, w3, w4, one);
        acc2 = tap_accum(acc2, q1.z, q2.z, q3.z, q4.z, w1, w2, w3, w4, one);
        acc3 = tap_accum(acc3, q1.w, q2.w, q3.w, q4.w, w1, w2, w3, w4, one);
      }
    }
    // / count (= 4): a power of two, so the multiply is the exactly rounded quotient as well
    const uint64_t qq = pack_f32x2(__float_as_uint(0.25f), __float_as_uint(0.25f));
    acc0 = mul_f32x2(acc0, qq); acc1 = mul_f32x2(acc1, qq); acc2 = mul_f32x2(acc2, qq); acc3 = mul_f32x2(acc3, qq);   // no add follows
    const long long o = ((long long)r * P * P + bin) * a.C + chunk * 8;
    if (a.out_fp32) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<float*>(a.out) + o);
      dst[0] = make_uint4((uint32_t)acc0, (uint32_t)(acc0 >> 32), (uint32_t)acc1, (uint32_t)(acc1 >> 32));
      dst[1] = make_uint4((uint32_t)acc2, (uint32_t)(acc2 >> 32), (uint32_t)acc3, (uint32_t)(acc3 >> 32));
    } else {
      uint4 v;
      v.x = cvt_bf16x2(acc0); v.y = cvt_bf16x2(acc1); v.z = cvt_bf16x2(acc2); v.w = cvt_bf16x2(acc3);
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(a.out) + o) = v;
    }
  }
}

int launch_roi_align(const RoiAlignArgs& a, cudaStream_t s) {
  if (a.C % 8 || 256 % (a.C / 8)) { set_error("roi_align: C/8 must divide 256"); return -1; }
  if (a.P > 32) { set_error("roi_align: pooler resolution %d > 32", a.P); return -1; }
  if (a.R == 0) return 0;
  const int parts = a.P >= 14 ? 4 : 1;             // large poolers: several CTAs per ROI (wave quantisation)
  roi_align_kernel<<<dim3(a.R, parts), 256, 0, s>>>(a, 1.0f);   // 1.0f as a run-time value: see tap_accum
  DPB_CHECK_LAUNCH("roi_align");
  return 0;
}

// ------------------------------------------------------------------------------------ box predictor tail

__global__ void __launch_bounds__(1024)
box_predict_kernel(BoxPredictArgs a, int nc) {
  extern __shared__ uint32_t bp_smem[];
  // layout: [nms region (kNms)] [keys 1024 u64] [boxes 1024 float4] [scores 1024 f32]
  const int kNms = 1024 * 32 * 4 + 1024 * 16 + 1024 * 4 + 32 * 4;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(bp_smem) + kNms);
  float4* sbox = reinterpret_cast<float4*>(keys + 1024);
  float* sscore = reinterpret_cast<float*>(sbox + 1024);
  __shared__ unsigned s_ncand, s_warp[32], s_total;
  // a cluster of nc CTAs per image: every CTA repeats the (cheap) score / decode / sort prologue, the IoU mask
  // of the NMS is split over the cluster, CTA 0 finishes
  const int b = blockIdx.x / nc, t = threadIdx.x;
  if (t == 0) s_ncand = 0;
  __syncthreads();
  const int np = a.prop_count[b] < a.R ? a.prop_count[b] : a.R;
  unsigned long long key = 0ull;
  if (t < np) {
    const float* h = a.head + ((long long)b * a.R + t) * 16;
    const float l0 = h[0], l1 = h[1];
    // F.softmax(dim=-1): exp(x - max) / sum  (fast_rcnn.py:325)
    const float m = fmaxf(l0, l1);
    const float e0 = expf(__fsub_rn(l0, m)), e1 = expf(__fsub_rn(l1, m));
    const float sum = __fadd_rn(e0, e1);
    const float p0 = __fdiv_rn(e0, sum), p1 = __fdiv_rn(e1, sum);
    const float4 pb = reinterpret_cast<const float4*>(a.prop_boxes)[(long long)b * a.R + t];
    float d[4] = {h[2], h[3], h[4], h[5]}, box[4];
    // Box2BoxTransform weights (10, 10, 5, 5): fast_rcnn.py:300-303
    {
      const float scale_clamp = 4.135166556742356f;
      const float widths = __fsub_rn(pb.z, pb.x), heights = __fsub_rn(pb.w, pb.y);
      const float ctr_x = __fadd_rn(pb.x, __fmul_rn(0.5f, widths));
      const float ctr_y = __fadd_rn(pb.y, __fmul_rn(0.5f, heights));
      const float dx = __fdiv_rn(d[0], 10.f), dy = __fdiv_rn(d[1], 10.f);
      const float dw = fminf(__fdiv_rn(d[2], 5.f), scale_clamp), dh = fminf(__fdiv_rn(d[3], 5.f), scale_clamp);
      const float pcx = __fadd_rn(__fmul_rn(dx, widths), ctr_x);
      const float pcy = __fadd_rn(__fmul_rn(dy, heights), ctr_y);
      const float pw = __fmul_rn(expf(dw), widths), ph = __fmul_rn(expf(dh), heights);
      box[0] = __fsub_rn(pcx, __fmul_rn(0.5f, pw)); box[1] = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
      box[2] = __fadd_rn(pcx, __fmul_rn(0.5f, pw)); box[3] = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
    }
    const bool finite = isfinite(box[0]) && isfinite(box[1]) && isfinite(box[2]) && isfinite(box[3]) &&
                        isfinite(p0) && isfinite(p1);
    sbox[t] = make_float4(box[0], box[1], box[2], box[3]);
    sscore[t] = p0;
    if (finite && p0 > a.score_thresh) {   // fast_rcnn.py:105-118 (clip at :113 is a discarded no-op)
      key = ((unsigned long long)f2ord(p0) << 32) | (0xFFFFFFFFu - (uint32_t)t);
      atomicAdd(&s_ncand, 1u);
    }
  }
  keys[t] = key;
  __syncthreads();
  bitonic_sort_desc(keys, 1024);
  const int ncand = (int)s_ncand;
  float4* gb = reinterpret_cast<float4*>(a.ws_boxes) + (long long)b * 1024;
  unsigned char* gk = a.ws_keep + (long long)b * 1024;
  int src = -1;
  if (t < ncand) {
    src = (int)(0xFFFFFFFFu - (uint32_t)(keys[t] & 0xFFFFFFFFull));
    gb[t] = sbox[src];
    gk[t] = 1;
  } else {
    gk[t] = 0;
  }
  __syncthreads();
  if (!nms_sorted_block(gb, ncand, a.nms_thresh, gk, bp_smem, nc)) return;   // batched_nms with a single class
  __syncthreads();
  // first `topk` kept, in score order
  const unsigned flag = (t < ncand && gk[t]) ? 1u : 0u;
  unsigned incl = flag;
  {
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      s_warp[lane] = w;
      if (lane == 31) s_total = w;
    }
    __syncthreads();
    if (warp > 0) incl += s_warp[warp - 1];
  }
  const int pos = (int)incl - 1;
  if (flag && pos < a.topk) {
    const float4 bx = sbox[src];
    const long long o = (long long)b * a.topk + pos;
    reinterpret_cast<float4*>(a.det_boxes_raw)[o] = bx;
    // detector_postprocess: scale_boxes then clip to (H_orig, W_orig)  (postprocessing.py:43-54)
    float4 ob;
    ob.x = fminf(fmaxf(__fmul_rn(bx.x, a.scale_x), 0.f), a.out_w);
    ob.y = fminf(fmaxf(__fmul_rn(bx.y, a.scale_y), 0.f), a.out_h);
    ob.z = fminf(fmaxf(__fmul_rn(bx.z, a.scale_x), 0.f), a.out_w);
    ob.w = fminf(fmaxf(__fmul_rn(bx.w, a.scale_y), 0.f), a.out_h);
    reinterpret_cast<float4*>(a.det_boxes)[o] = ob;
    a.det_scores[o] = sscore[src];
  }
  if (t == 0) a.det_count[b] = (int)s_total < a.topk ? (int)s_total : a.topk;
}

static constexpr int kBoxPredictSmem = (1024 * 32 * 4 + 1024 * 16 + 1024 * 4 + 32 * 4) + 1024 * 8 + 1024 * 16 + 1024 * 4;

int launch_box_predict(const BoxPredictArgs& a, cudaStream_t s) {
  if (a.R > 1024) { set_error("box_predict: R > 1024"); return -1; }
  static bool done[64] = {};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(box_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kBoxPredictSmem);
    if (e != cudaSuccess) { set_error("box_predict smem: %s", cudaGetErrorString(e)); return -3; }
    done[dev] = true;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int nc = pick_nms_cluster(a.B, sms > 0 ? sms : 148);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.B * nc); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = kBoxPredictSmem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, box_predict_kernel, a, nc);
  if (e != cudaSuccess) { set_error("box_predict launch: %s", cudaGetErrorString(e)); return -4; }
  return 0;
}

__global__ void pack_rois_kernel(const float4* __restrict__ boxes, const int* __restrict__ count, int B,
                                 int topk, float* __restrict__ rois, int* __restrict__ total,
                                 int* __restrict__ offsets) {
  __shared__ int s_off[1025];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) { s_off[b] = acc; acc += count[b]; }
    s_off[B] = acc;
    *total = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= B; i += blockDim.x) offsets[i] = s_off[i];
  for (int i = threadIdx.x; i < B * topk; i += blockDim.x) {
    const int b = i / topk, t = i - b * topk;
    if (t < count[b]) {
      const float4 bx = boxes[i];
      float* o = rois + (long long)(s_off[b] + t) * 5;
      o[0] = (float)b; o[1] = bx.x; o[2] = bx.y; o[3] = bx.z; o[4] = bx.w;
    }
  }
}

int launch_pack_rois(const float* det_boxes_raw, const int* det_count, int B, int topk, float* rois,
                     int* total, int* offsets, cudaStream_t s) {
  if (B > 1024) { set_error("pack_rois: B > 1024"); return -1; }
  pack_rois_kernel<<<1, 256, 0, s>>>(reinterpret_cast<const float4*>(det_boxes_raw), det_count, B, topk,
                                     rois, total, offsets);
  DPB_CHECK_LAUNCH("pack_rois");
  return 0;
}

__global__ void proposal_rois_kernel(const float4* __restrict__ boxes, const int* __restrict__ count, int B,
                                     int R, float* __restrict__ rois) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * R; i += gridDim.x * blockDim.x) {
    const int b = i / R, t = i - b * R;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < count[b]) bx = boxes[i];
    float* o = rois + (long long)i * 5;
    o[0] = (float)b; o[1] = bx.x; o[2] = bx.y; o[3] = bx.z; o[4] = bx.w;
  }
}

int launch_proposal_rois(const float* prop_boxes, const int* prop_count, int B, int R, float* rois,
                         cudaStream_t s) {
  const int g = (B * R + 255) / 256;
  proposal_rois_kernel<<<g, 256, 0, s>>>(reinterpret_cast<const float4*>(prop_boxes), prop_count, B, R, rois);
  DPB_CHECK_LAUNCH("proposal_rois");
  return 0;
}

}  // namespace dpb
