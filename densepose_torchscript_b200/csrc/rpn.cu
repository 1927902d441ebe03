// RPN proposal selection on the device, no host sync:
//   rpn_topk_decode : per (image, level) exact top-k of the objectness logits (radix select + bitonic
//                     sort), anchor generation + Box2BoxTransform.apply_deltas for the selected anchors
//                     only, (swapped-extent) clip, finite / non-empty flags.
//   rpn_nms         : per (image, level) greedy NMS (IoU bit-mask in shared memory + warp scan).
//   rpn_merge       : per image merge of the kept candidates, sort by score, first post_topk.
// Reference: proposal_generator/rpn.py:319-394, proposal_utils.py:19-134, box_regression.py:74-112,
// anchor_generator.py:165-231, structures.py:107-122; torchvision nms semantics.
#include "kernels.cuh"
#include "conv_igemm.cuh"
#include "device_utils.cuh"

#include <math.h>

namespace dpb {

#define DPB_CHECK_LAUNCH(name)                                                     \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      set_error("%s launch: %s", name, cudaGetErrorString(e__));                   \
      return -4;                                                                   \
    }                                                                              \
  } while (0)

__device__ __forceinline__ void decode_box(const float* __restrict__ d, float ax1, float ay1, float ax2,
                                           float ay2, float wx, float wy, float ww, float wh,
                                           float* out) {
  // box_regression.py:86-110, every operation rounded separately (no FMA) like the eager reference
  const float scale_clamp = 4.135166556742356f;   // log(1000/16)
  const float widths = __fsub_rn(ax2, ax1), heights = __fsub_rn(ay2, ay1);
  const float ctr_x = __fadd_rn(ax1, __fmul_rn(0.5f, widths));
  const float ctr_y = __fadd_rn(ay1, __fmul_rn(0.5f, heights));
  const float dx = __fdiv_rn(d[0], wx), dy = __fdiv_rn(d[1], wy);
  float dw = __fdiv_rn(d[2], ww), dh = __fdiv_rn(d[3], wh);
  dw = fminf(dw, scale_clamp);
  dh = fminf(dh, scale_clamp);
  const float pcx = __fadd_rn(__fmul_rn(dx, widths), ctr_x);
  const float pcy = __fadd_rn(__fmul_rn(dy, heights), ctr_y);
  const float pw = __fmul_rn(expf(dw), widths);
  const float ph = __fmul_rn(expf(dh), heights);
  out[0] = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  out[1] = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  out[2] = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
  out[3] = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
}

// A cluster of nc CTAs per (level, image): a p2 level is 1.6 M logits that one CTA would stream four times (three
// histogram passes and the collection); here every CTA takes a contiguous slice of the pixels, the per-pass
// histograms stay in each CTA's shared memory and every CTA sums them through distributed shared memory (so all of
// them walk the same radix path without a broadcast), and the selected keys are appended to CTA 0's list with
// cluster-scope atomics. CTA 0 sorts, decodes and writes.
__global__ void __launch_bounds__(1024)
rpn_topk_decode_kernel(RpnArgs a, int nc) {
  const int lvl = blockIdx.x / nc, b = blockIdx.y;
  const uint32_t rank = nc > 1 ? cluster_ctarank() : 0u;
  const RpnLevel& L = a.lvl[lvl];
  const int n = L.H * L.W * 3;
  const int k = n < a.pre_topk ? n : a.pre_topk;
  const float* head = L.head + (long long)b * L.H * L.W * 16;

  __shared__ unsigned hist[2048];
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned long long keys[1024];
  __shared__ unsigned s_bin, s_above, s_cnt, s_eq_taken;

  const int npix = L.H * L.W;
  const float4* head4 = reinterpret_cast<const float4*>(head);
  auto key_at = [&](int i) -> uint32_t { return f2ord(__ldg(head + (long long)(i / 3) * 16 + (i % 3))); };
  // this CTA's pixels
  const int per = (npix + nc - 1) / nc;
  const int px_lo = (int)rank * per, px_hi = min(npix, px_lo + per);
  const uint32_t hist_sa = (uint32_t)__cvta_generic_to_shared(hist);
  // merged count of one bin: the sum over the cluster's histograms
  auto merged = [&](int bin) -> unsigned {
    if (nc == 1) return hist[bin];
    unsigned c = 0;
    for (int r = 0; r < nc; ++r) c += ld_shared_cluster_u32(cluster_map_shared(hist_sa + 4u * (uint32_t)bin, (uint32_t)r));
    return c;
  };

  if (threadIdx.x == 0) { s_cnt = 0; s_eq_taken = 0; }
  keys[threadIdx.x] = 0ull;

  uint32_t prefix = 0, mask = 0;
  unsigned remaining = (unsigned)k;
  unsigned eq_total = 0;
  const int shifts[3] = {21, 10, 0};
  const int bits[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; ++pass) {
    const int nb = 1 << bits[pass];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    // four independent 16-byte loads per thread are issued before any is used, otherwise every pass is a chain of
    // dependent L2 round trips
    for (int px0 = px_lo + threadIdx.x; px0 < px_hi; px0 += 4 * blockDim.x) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int px = px0 + j * blockDim.x;
        if (px < px_hi) v[j] = __ldg(head4 + (long long)px * 4);   // (logit a0, a1, a2, first delta): one 16-byte load
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (px0 + j * (int)blockDim.x >= px_hi) break;
        const uint32_t u0 = f2ord(v[j].x), u1 = f2ord(v[j].y), u2 = f2ord(v[j].z);
        if ((u0 & mask) == prefix) atomicAdd(&hist[(u0 >> shifts[pass]) & (nb - 1)], 1u);
        if ((u1 & mask) == prefix) atomicAdd(&hist[(u1 >> shifts[pass]) & (nb - 1)], 1u);
        if ((u2 & mask) == prefix) atomicAdd(&hist[(u2 >> shifts[pass]) & (nb - 1)], 1u);
      }
    }
    if (nc > 1) cluster_sync_all();          // every CTA's histogram is complete and visible
    else __syncthreads();
    // reversed bins: thread t owns reversed positions 2t, 2t+1 (bin = nb-1-pos)
    unsigned c0 = 0, c1 = 0;
    const int p0 = 2 * threadIdx.x, p1 = p0 + 1;
    if (p0 < nb) c0 = merged(nb - 1 - p0);
    if (p1 < nb) c1 = merged(nb - 1 - p1);
    const unsigned incl = block_scan_incl(c0 + c1, warp_sums);
    const unsigned before = incl - (c0 + c1);
    if (before < remaining && remaining <= incl) {
      if (remaining <= before + c0) { s_bin = nb - 1 - p0; s_above = before; }
      else { s_bin = nb - 1 - p1; s_above = before + c0; }
    }
    __syncthreads();
    prefix |= (s_bin << shifts[pass]);
    mask |= ((uint32_t)(nb - 1) << shifts[pass]);
    remaining -= s_above;
    if (pass == 2) eq_total = merged((int)s_bin);       // elements equal to the k-th key
    if (nc > 1) cluster_sync_all();          // all remote reads of this pass are done before any histogram is reset
    else __syncthreads();
  }
  // prefix == k-th largest key T; `remaining` of the elements equal to T are needed.
  const uint32_t T = prefix;
  const bool take_all_eq = (eq_total == remaining);
  const uint32_t cnt_sa = cluster_map_shared((uint32_t)__cvta_generic_to_shared(&s_cnt), 0u);
  const uint32_t keys_sa = cluster_map_shared((uint32_t)__cvta_generic_to_shared(keys), 0u);
  for (int px0 = px_lo + threadIdx.x; px0 < px_hi; px0 += 4 * blockDim.x) {
    float4 v4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int px = px0 + j * blockDim.x;
      if (px < px_hi) v4[j] = __ldg(head4 + (long long)px * 4);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int px = px0 + j * blockDim.x;
      if (px >= px_hi) break;
      const uint32_t us[3] = {f2ord(v4[j].x), f2ord(v4[j].y), f2ord(v4[j].z)};
#pragma unroll
      for (int an = 0; an < 3; ++an) {
        const uint32_t u = us[an];
        if (u > T || (take_all_eq && u == T)) {
          // the list lives in CTA 0 (its order does not matter: it is sorted by (score, ~index) below)
          const unsigned long long key = ((unsigned long long)u << 32) | (0xFFFFFFFFu - (uint32_t)(px * 3 + an));
          if (nc > 1) {
            const unsigned pos = atom_add_shared_cluster_u32(cnt_sa, 1u);
            if (pos < 1024) st_shared_cluster_u64(keys_sa + 8u * pos, key);
          } else {
            const unsigned pos = atomicAdd(&s_cnt, 1u);
            if (pos < 1024) keys[pos] = key;
          }
        }
      }
    }
  }
  if (nc > 1) {
    cluster_sync_all();                      // every key has landed in CTA 0
    if (rank != 0) return;
  } else {
    __syncthreads();
  }
  if (!take_all_eq) {
    // excess ties: keep the lowest-index ones, chunk by chunk in index order (CTA 0 walks the whole level: rare)
    for (int base = 0; base < n; base += blockDim.x) {
      const int i = base + threadIdx.x;
      const unsigned flag = (i < n && key_at(i) == T) ? 1u : 0u;
      const unsigned incl = block_scan_incl(flag, warp_sums);
      const unsigned taken = s_eq_taken;
      if (flag && taken + incl <= remaining) {
        const unsigned pos = atomicAdd(&s_cnt, 1u);
        if (pos < 1024) keys[pos] = ((unsigned long long)T << 32) | (0xFFFFFFFFu - (uint32_t)i);
      }
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) s_eq_taken = taken + incl;
      __syncthreads();
      if (s_eq_taken >= remaining) break;
    }
  }
  __syncthreads();
  bitonic_sort_desc(keys, 1024);

  const int t = threadIdx.x;
  const long long slot = ((long long)b * 5 + lvl) * a.pre_topk + t;
  if (t < k) {
    const unsigned long long key = keys[t];
    const float score = ord2f((uint32_t)(key >> 32));
    const int i = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
    const int an = i % 3, pix = i / 3;
    const int x = pix % L.W, y = pix / L.W;
    const float sx = (float)x * L.stride, sy = (float)y * L.stride;
    const float ax1 = __fadd_rn(sx, L.anchors[an * 4 + 0]), ay1 = __fadd_rn(sy, L.anchors[an * 4 + 1]);
    const float ax2 = __fadd_rn(sx, L.anchors[an * 4 + 2]), ay2 = __fadd_rn(sy, L.anchors[an * 4 + 3]);
    float d[4], box[4];
    // deltas live at channels 3 + an*4 .. +3 (not 16-byte aligned): scalar loads
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = __ldg(head + (long long)pix * 16 + 3 + an * 4 + j);
    decode_box(d, ax1, ay1, ax2, ay2, 1.f, 1.f, 1.f, 1.f, box);
    const bool finite = isfinite(box[0]) && isfinite(box[1]) && isfinite(box[2]) && isfinite(box[3]) &&
                        isfinite(score);
    // clip_boxes(boxes, image_size) with image_size = (W_pad, H_pad): x in [0, H_pad], y in [0, W_pad]
    box[0] = fminf(fmaxf(box[0], 0.f), a.clip_x);
    box[1] = fminf(fmaxf(box[1], 0.f), a.clip_y);
    box[2] = fminf(fmaxf(box[2], 0.f), a.clip_x);
    box[3] = fminf(fmaxf(box[3], 0.f), a.clip_y);
    const bool nonempty = (__fsub_rn(box[2], box[0]) >= 0.f) && (__fsub_rn(box[3], box[1]) >= 0.f);
    reinterpret_cast<float4*>(a.cand_boxes)[slot] = make_float4(box[0], box[1], box[2], box[3]);
    a.cand_scores[slot] = score;
    a.cand_keep[slot] = (finite && nonempty) ? 1 : 0;
  } else if (t < a.pre_topk) {
    a.cand_keep[slot] = 0;
  }
  if (t == 0) a.cand_count[b * 5 + lvl] = k;
}

static constexpr int kNmsSmem = 1024 * 32 * 4 + 1024 * 16 + 1024 * 4 + 32 * 4;

__global__ void __launch_bounds__(1024) rpn_nms_kernel(RpnArgs a, int nc) {
  extern __shared__ uint32_t nms_smem[];
  const int lvl = blockIdx.x / nc, b = blockIdx.y;       // a cluster of nc CTAs (along x) per (level, image)
  const int n = a.cand_count[b * 5 + lvl];
  const long long base = ((long long)b * 5 + lvl) * a.pre_topk;
  nms_sorted_block(reinterpret_cast<const float4*>(a.cand_boxes) + base, n, a.nms_thresh,
                   a.cand_keep + base, nms_smem, nc);
}

__global__ void __launch_bounds__(1024)
nms_sorted_kernel(const float4* boxes, int n, float thr, unsigned char* keep, int nc) {
  extern __shared__ uint32_t nms_smem[];
  nms_sorted_block(boxes, n, thr, keep, nms_smem, nc);
}

// opt-in dynamic shared memory, once per (kernel, device)
static int ensure_smem(const void* fn, int bytes, bool* done_per_device) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (done_per_device[dev]) return 0;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { set_error("smem attr: %s", cudaGetErrorString(e)); return -3; }
  done_per_device[dev] = true;
  return 0;
}

static int device_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// launch `fn` with clusters of nc CTAs along x
template <typename... KArgs, typename... Args>
static cudaError_t launch_clustered(void (*fn)(KArgs...), dim3 grid, int threads, int smem, int nc, cudaStream_t s,
                                    Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, fn, args...);
}

int launch_rpn_topk_decode(const RpnArgs& a, cudaStream_t s) {
  if (a.pre_topk > 1024) { set_error("rpn: pre_topk > 1024 unsupported"); return -1; }
  const int nc = pick_nms_cluster(5 * a.B, device_sms());
  cudaError_t e = launch_clustered(rpn_topk_decode_kernel, dim3(5 * nc, a.B), 1024, 0, nc, s, a, nc);
  if (e != cudaSuccess) { set_error("rpn_topk_decode launch: %s", cudaGetErrorString(e)); return -4; }
  return 0;
}

int launch_rpn_nms(const RpnArgs& a, cudaStream_t s) {
  static bool done[64] = {};
  if (ensure_smem((const void*)rpn_nms_kernel, kNmsSmem, done)) return -3;
  const int nc = pick_nms_cluster(5 * a.B, device_sms());
  cudaError_t e = launch_clustered(rpn_nms_kernel, dim3(5 * nc, a.B), 1024, kNmsSmem, nc, s, a, nc);
  if (e != cudaSuccess) { set_error("rpn_nms launch: %s", cudaGetErrorString(e)); return -4; }
  return 0;
}

int launch_nms_sorted(const float* boxes, int n, float thr, unsigned char* keep, cudaStream_t s) {
  if (n > 1024) { set_error("nms_sorted: n > 1024"); return -1; }
  static bool done[64] = {};
  if (ensure_smem((const void*)nms_sorted_kernel, kNmsSmem, done)) return -3;
  const int nc = 8;
  cudaError_t e = launch_clustered(nms_sorted_kernel, dim3(nc), 1024, kNmsSmem, nc, s,
                                   reinterpret_cast<const float4*>(boxes), n, thr, keep, nc);
  if (e != cudaSuccess) { set_error("nms_sorted launch: %s", cudaGetErrorString(e)); return -4; }
  return 0;
}

// ---------------------------------------------------------------------------------------- merge
// proposal_utils.py:118-126 after the per-level NMS: torchvision's batched_nms returns the kept candidates of all
// levels sorted by score, keep[:post_topk]. Every level's candidate list is already sorted (score descending, ties by
// ascending anchor index), so no sort is needed: the final rank of a kept candidate is the number of kept candidates
// ahead of it = (kept ones before it in its own level) + per other level (kept ones among the first p entries, p found
// by binary search on that level's sorted scores; equal scores: the lower level goes first, as the (score, ~index)
// keys of the previous bitonic sort ordered them). One CTA per image, 1024 threads, ~5 k candidates.
__global__ void __launch_bounds__(1024) rpn_merge_kernel(RpnArgs a) {
  __shared__ uint32_t s_ord[5][1024];       // order-preserving score keys, 0 = empty slot (sorts last)
  __shared__ uint16_t s_pref[5][1024];      // exclusive prefix count of kept candidates
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_total[5];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  uint32_t flags = 0;                         // bit l: candidate (level l, position t) is kept
  for (int l = 0; l < 5; ++l) {
    const long long slot = ((long long)b * 5 + l) * a.pre_topk + t;
    const bool valid = t < a.pre_topk && t < a.cand_count[b * 5 + l];
    s_ord[l][t] = valid ? f2ord(a.cand_scores[slot]) : 0u;
    const uint32_t k = (valid && a.cand_keep[slot]) ? 1u : 0u;
    flags |= k << l;
    // block-wide exclusive scan of k
    uint32_t incl = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      s_warp[lane] = w;
      if (lane == 31) s_total[l] = w;
    }
    __syncthreads();
    s_pref[l][t] = (uint16_t)(incl - k + (warp > 0 ? s_warp[warp - 1] : 0u));
    __syncthreads();
  }
  for (int l = 0; l < 5; ++l) {
    if (!((flags >> l) & 1u)) continue;
    const uint32_t key = s_ord[l][t];
    uint32_t rank = s_pref[l][t];
    for (int m = 0; m < 5; ++m) {
      if (m == l) continue;
      // first position p of level m whose candidate does NOT come before (key, l): entries are descending
      int lo = 0, hi = 1024;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t o = s_ord[m][mid];
        const bool before = o > key || (o == key && m < l);
        if (before) lo = mid + 1; else hi = mid;
      }
      rank += lo < 1024 ? (uint32_t)s_pref[m][lo] : s_total[m];
    }
    if (rank < (uint32_t)a.post_topk) {
      const long long slot = ((long long)b * 5 + l) * a.pre_topk + t;
      reinterpret_cast<float4*>(a.prop_boxes)[(long long)b * a.post_topk + rank] = reinterpret_cast<const float4*>(a.cand_boxes)[slot];
      a.prop_scores[(long long)b * a.post_topk + rank] = a.cand_scores[slot];
    }
  }
  const uint32_t kept = s_total[0] + s_total[1] + s_total[2] + s_total[3] + s_total[4];
  const int n = (int)kept < a.post_topk ? (int)kept : a.post_topk;
  for (int i = n + t; i < a.post_topk; i += blockDim.x) {
    reinterpret_cast<float4*>(a.prop_boxes)[(long long)b * a.post_topk + i] = make_float4(0.f, 0.f, 0.f, 0.f);
    a.prop_scores[(long long)b * a.post_topk + i] = 0.f;
  }
  if (t == 0) a.prop_count[b] = n;
}

int launch_rpn_merge(const RpnArgs& a, cudaStream_t s) {
  if (a.pre_topk > 1024) { set_error("rpn_merge: pre_topk > 1024"); return -1; }
  rpn_merge_kernel<<<a.B, 1024, 0, s>>>(a);
  DPB_CHECK_LAUNCH("rpn_merge");
  return 0;
}

}  // namespace dpb
