// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Replaces every F.conv2d / nn.Linear / ConvTranspose2d-phase on the reference hot path
// (detectron2/layers/wrappers.py:104-112 Conv2d.forward, box_head.py:95-98, chart.py:76-90).
//
// GEMM view: M = output pixels (128 consecutive (n, oy, ox) pixels per tile through im2col-mode TMA, or a
// tw x th x tn pixel box through tiled TMA), N = Cout tile (block_n <= 256), K = taps x Cin in 64-channel chunks.
// Per chunk the producer issues one TMA load of the shifted input window (zero fill outside the image implements
// the padding) and one 2-D load of the packed weights, both 128B-swizzled; one elected thread issues four
// tcgen05.mma (K = 16 each) into a double-buffered TMEM accumulator; eight epilogue warps drain TMEM with tcgen05.ld
// and fuse bias / residual / nearest-x2 top-down add / ReLU / dtype conversion, leaving through swizzled shared
// memory slabs + TMA stores (bf16 matrix outputs) or direct stores (fp32 / strided / channel-planar outputs).
//
// Warp roles (384 threads, persistent, 1 CTA per SM):
//   warp 0 : TMA producer              warp 1 : MMA issuer (CTA pairs: the leader CTA only)
//   warp 2 : TMEM alloc / dealloc, residual prefetch (TMA -> slab)
//   warp 3 : TMA store issuer          warps 4..11 : epilogue (TMEM lane quarter = warp % 4, column half = (warp-4)/4)
// Variants selected per layer by conv_plan_build: staged / direct epilogue, CTA pairs (cta_group::2), deconv phases
// as N blocks (phase_taps). Every launch is a programmatic dependent launch.
#include "conv_igemm.cuh"
#include "ptx.cuh"
#include "device_utils.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace dpb {

static thread_local char g_err[512];
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static constexpr int kThreads = 384;          // TMA / MMA / residual / store warps + 8 epilogue warps
static constexpr int kEpiThreads = 256;
static constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16
static constexpr int kSlabBytes = 128 * 128;  // epilogue slab: 128 rows x 64 bf16, 128B-swizzled

// One thread's 32 accumulator columns of a slab row: + bias (+ residual read back from the slab) -> relu ->
// bf16 -> four 16-byte swizzled shared stores. Adds are packed f32x2 (IEEE rn, same results as scalar adds).
template <bool kRes>
__device__ __forceinline__ void epi_slab_half(const uint32_t* v, const float4* bv, uint32_t srow, uint32_t sw,
                                              int hh, bool relu) {
#pragma unroll
  for (int c8 = 0; c8 < 4; ++c8) {
    const float4 b0 = bv[2 * c8], b1 = bv[2 * c8 + 1];
    uint64_t a01 = add_f32x2(pack_f32x2(v[c8 * 8 + 0], v[c8 * 8 + 1]), pack_f32x2(__float_as_uint(b0.x), __float_as_uint(b0.y)));
    uint64_t a23 = add_f32x2(pack_f32x2(v[c8 * 8 + 2], v[c8 * 8 + 3]), pack_f32x2(__float_as_uint(b0.z), __float_as_uint(b0.w)));
    uint64_t a45 = add_f32x2(pack_f32x2(v[c8 * 8 + 4], v[c8 * 8 + 5]), pack_f32x2(__float_as_uint(b1.x), __float_as_uint(b1.y)));
    uint64_t a67 = add_f32x2(pack_f32x2(v[c8 * 8 + 6], v[c8 * 8 + 7]), pack_f32x2(__float_as_uint(b1.z), __float_as_uint(b1.w)));
    const uint32_t saddr = srow + ((((uint32_t)(hh * 4 + c8)) ^ sw) << 4);
    if (kRes) {
      const uint4 r = ld_shared_v4(saddr);
      a01 = add_f32x2(a01, pack_f32x2(r.x << 16, r.x & 0xffff0000u));
      a23 = add_f32x2(a23, pack_f32x2(r.y << 16, r.y & 0xffff0000u));
      a45 = add_f32x2(a45, pack_f32x2(r.z << 16, r.z & 0xffff0000u));
      a67 = add_f32x2(a67, pack_f32x2(r.w << 16, r.w & 0xffff0000u));
    }
    uint4 o;
    if (relu) {
      o.x = cvt_bf16x2_relu(a01); o.y = cvt_bf16x2_relu(a23);
      o.z = cvt_bf16x2_relu(a45); o.w = cvt_bf16x2_relu(a67);
    } else {
      o.x = cvt_bf16x2(a01); o.y = cvt_bf16x2(a23);
      o.z = cvt_bf16x2(a45); o.w = cvt_bf16x2(a67);
    }
    st_shared_v4(saddr, o);
  }
}

// kStaged: bf16 NHWC outputs leave through shared-memory slabs and TMA stores (and the residual arrives by
// TMA into the same slab), so the epilogue warps only touch TMEM and shared memory. Otherwise every
// epilogue thread stores its own row straight to global memory (fp32 / strided / planar outputs).
//
// kPair (staged, im2col only): two CTAs of a cluster work as one cta_group::2 MMA unit on a 256-row x block_n tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (block_n/2 weight rows), so the pair reads the
// weights from L2 once instead of twice; the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 for both,
// every CTA drains its own TMEM lanes (its 128 output rows). full[] lives in the leader (both CTAs' TMA loads
// count their bytes there), empty[] / tmem_full[] are signalled in both CTAs by multicast commits, and the leader's
// tmem_empty[] collects one arrive per epilogue warp of both CTAs.
template <bool kStaged, bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                  const __grid_constant__ CUtensorMap tmA2, const ConvKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // warp index made provably warp-uniform, so the role loops below compile to uniform-datapath code
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();      // the next kernel may start its own prologue as soon as SMs free up

  // 1024-byte aligned stage buffers (SWIZZLE_128B atoms are 8 rows x 128 B)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t b_bytes = (uint32_t)(kPair ? p.block_n >> 1 : p.block_n) * 128u;
  const uint32_t stage_bytes = kABytes + ((b_bytes + 1023u) & ~1023u);
  const uint32_t slab_base = smem_base + (uint32_t)(p.stages * p.ks) * stage_bytes;
  const uint32_t bar_base = slab_base + (uint32_t)p.nslab * kSlabBytes;
  // barriers: full[s], empty[s], tmem_full[2], tmem_empty[2], slab res_full / ready / free [nslab];
  // then the TMEM base address slot
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * p.stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * p.stages + 2 + s); };
  auto sres_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * p.stages + 4 + s); };
  auto sready_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * p.stages + 4 + p.nslab + s); };
  auto sfree_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * p.stages + 4 + 2 * p.nslab + s); };
  const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * p.stages + 4 + 3 * p.nslab);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const uint32_t tmem_cols = 2u * (uint32_t)p.acc_stride;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.nseg > 1) prefetch_tmap(&tmA2);
    if (kStaged) {
      prefetch_tmap(&tmOut);
      if (p.res_tma) prefetch_tmap(&tmRes);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kPair ? 2 * (kEpiThreads / 32) : kEpiThreads);
    }
    for (int s = 0; s < p.nslab; ++s) {
      mbar_init(sres_bar(s), 1);
      mbar_init(sready_bar(s), kEpiThreads);
      mbar_init(sfree_bar(s), 1);
    }
    fence_mbar_init();
  }
  if (kPair) {
    cluster_sync_divergent();                       // both CTAs' barriers exist before anything remote can touch them
    if (warp == 2) tmem_alloc_pair(tmem_slot, tmem_cols);
    tc_fence_before();
    cluster_sync_divergent();
  } else {
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlaps the tail of the
  // previous kernel in the stream; its results (and n_valid) are only touched after this wait.
  griddep_wait();

  const int nvalid = p.n_valid ? min(*p.n_valid, p.N) : p.N;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  // pair mode: a "tile" is a pair of consecutive 128-row tiles (this CTA takes the rank-th of them)
  const int total_tiles = (kPair ? (m_tiles + 1) >> 1 : m_tiles) * p.n_blocks;
  const int t_begin = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int k_iters = p.kh * p.kw * p.cin_chunks * p.nseg;
  const uint32_t a_tx = p.im2col ? (uint32_t)kABytes : (uint32_t)(p.tw * p.th * p.tn) * 128u;
  const int hw_out = p.H_out * p.W_out;
  const int m_valid = nvalid * hw_out;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp runs the loop
    // with warp-uniform values; one elected lane issues the copies)
    int s = 0;
    uint32_t ph = 0;
    const int n_stages = p.stages, cin_chunks = p.cin_chunks, kw = p.kw, dil = p.dil, block_n = p.block_n;
    const int ks = p.ks, n_groups = (k_iters + ks - 1) / ks;
    const bool im2col = p.im2col != 0;
    for (int tile = t_begin; tile < total_tiles; tile += t_step) {
      const int n_blk = tile % p.n_blocks;
      int mt = tile / p.n_blocks;
      if (kPair) {
        if (mt * 256 >= m_valid) continue;
        mt = 2 * mt + (int)rank;
      }
      int ix0, iy0, in0;
      if (im2col) {
        // column = 128 consecutive output pixels in (n, oy, ox) order
        const int m0 = mt * 128;
        if (!kPair && m0 >= m_valid) continue;
        in0 = m0 / hw_out;
        const int rem = m0 - in0 * hw_out;
        const int oy0 = rem / p.W_out;
        ix0 = (rem - oy0 * p.W_out) * p.sx - p.pad_x;
        iy0 = oy0 * p.sy - p.pad_y;
      } else {
        const int twi = mt % p.tiles_w;
        const int thi = (mt / p.tiles_w) % p.tiles_h;
        const int tni = mt / (p.tiles_w * p.tiles_h);
        if (tni * p.tn >= nvalid) continue;
        ix0 = twi * p.tw * p.sx - p.pad_x;
        iy0 = thi * p.th * p.sy - p.pad_y;
        in0 = tni * p.tn;
      }
      // flat k loop over stage groups of ks 64-channel chunks (one barrier pair per group): (ky, kx, cc)
      // advance incrementally; the loop-invariant launch parameters live in registers
      const int b_row = n_blk * block_n + (kPair ? (int)rank * (block_n >> 1) : 0);
      // deconv phases: block (py,px) = (n_blk >> 1, n_blk & 1) starts its 2x2 taps at (py,px) of the 3x3 footprint
      const int pyo = p.phase_taps ? (n_blk >> 1) * dil : 0, pxo = p.phase_taps ? (n_blk & 1) * dil : 0;
      // K order per tap: cin chunks fastest, then the strict-mode segment (hi*w_hi, hi*w_lo, lo*w_hi), then (kx, ky);
      // the weight matrix is packed in exactly this order, so its column just advances by 64 per chunk
      int cc = 0, seg = 0, kx = 0, ky = 0, kcol = 0;
      const int nseg = p.nseg;
      auto advance = [&]() {
        if (++cc == cin_chunks) { cc = 0; if (++seg == nseg) { seg = 0; if (++kx == kw) { kx = 0; ++ky; } } }
        kcol += 64;
      };
      for (int g = 0; g < n_groups; ++g) {
        mbar_wait(bar_base + 8u * (uint32_t)(n_stages + s), ph ^ 1u);           // empty[s]
        const int nsub = (ks == 2 && g * 2 + 1 < k_iters) ? 2 : 1;
        // coordinates of the (up to) two chunks of this group, computed by the whole warp
        const int c0 = cc * 64, ox0 = kx * dil + pxo, oy0 = ky * dil + pyo, kc0 = kcol;
        const CUtensorMap* tm0 = seg == 2 ? &tmA2 : &tmA;
        advance();
        const int c1 = cc * 64, ox1 = kx * dil + pxo, oy1 = ky * dil + pyo, kc1 = kcol;
        const CUtensorMap* tm1 = seg == 2 ? &tmA2 : &tmA;
        if (nsub == 2) advance();
        if (kPair) {
          if (elect_one()) {
            const uint32_t a_dst = smem_base + (uint32_t)s * stage_bytes;
            const uint32_t fb = bar_base + 8u * (uint32_t)s;                      // full[s]
            if (rank == 0) mbar_expect_tx(fb, 2u * (a_tx + b_bytes));             // both CTAs' bytes land here
            const uint32_t fb0 = mapa_shared(fb, 0);
            tma_load_im2col_4d_pair(a_dst, tm0, fb0, c0, ix0, iy0, in0, (uint16_t)ox0, (uint16_t)oy0);
            tma_load_2d_pair(a_dst + kABytes, &tmB, fb0, kc0, b_row);
          }
        } else if (elect_one()) {
          const uint32_t a_dst = smem_base + (uint32_t)(s * ks) * stage_bytes;
          const uint32_t fb = bar_base + 8u * (uint32_t)s;                        // full[s]
          mbar_expect_tx(fb, (uint32_t)nsub * (a_tx + b_bytes));
          if (im2col) tma_load_im2col_4d(a_dst, tm0, fb, c0, ix0, iy0, in0, (uint16_t)ox0, (uint16_t)oy0);
          else tma_load_4d(a_dst, tm0, fb, c0, ix0 + ox0, iy0 + oy0, in0);
          tma_load_2d(a_dst + kABytes, &tmB, fb, kc0, b_row);
          if (nsub == 2) {
            if (im2col) tma_load_im2col_4d(a_dst + stage_bytes, tm1, fb, c1, ix0, iy0, in0, (uint16_t)ox1, (uint16_t)oy1);
            else tma_load_4d(a_dst + stage_bytes, tm1, fb, c1, ix0 + ox1, iy0 + oy1, in0);
            tma_load_2d(a_dst + stage_bytes + kABytes, &tmB, fb, kc1, b_row);
          }
        }
        __syncwarp();
        if (++s == n_stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane;
    // pair mode: the leader CTA only)
    const uint32_t idesc = umma_idesc_bf16(kPair ? 256 : 128, p.block_n);
    const uint32_t desc_lo0 = (uint32_t)(umma_desc_sw128(smem_base) & 0xffffffffu);
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO | version | SWIZZLE_128B
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    const int n_stages = p.stages, ks = p.ks, n_groups = (k_iters + ks - 1) / ks;
    for (int tile = t_begin; tile < total_tiles; tile += t_step) {
      const int mt = tile / p.n_blocks;
      if (kPair) {
        if (mt * 256 >= m_valid) continue;
      } else if (p.im2col) {
        if (mt * 128 >= m_valid) continue;
      } else {
        if ((mt / (p.tiles_w * p.tiles_h)) * p.tn >= nvalid) continue;
      }
      mbar_wait(tempty_bar(as), aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.acc_stride);
      for (int g = 0; g < n_groups; ++g) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const int nsub = (ks == 2 && g * 2 + 1 < k_iters) ? 2 : 1;
        if (elect_one()) {
          // descriptor low word = 16-byte-unit address | LBO: stepping a chunk or a K=16 slice is an add
          uint32_t a_lo = desc_lo0 + (uint32_t)(s * ks) * (stage_bytes >> 4);
          for (int j = 0; j < nsub; ++j) {
            const uint32_t b_lo = a_lo + (kABytes >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kPair)
                umma_bf16_pair(d_tmem, ((uint64_t)kDescHi << 32) | (a_lo + 2u * k), ((uint64_t)kDescHi << 32) | (b_lo + 2u * k),
                               idesc, (g > 0 || j > 0 || k > 0) ? 1u : 0u);
              else
                umma_bf16(d_tmem, ((uint64_t)kDescHi << 32) | (a_lo + 2u * k), ((uint64_t)kDescHi << 32) | (b_lo + 2u * k),
                          idesc, (g > 0 || j > 0 || k > 0) ? 1u : 0u);
            }
            a_lo += stage_bytes >> 4;
          }
          // frees the group's smem (in both CTAs of a pair) when these MMAs retire
          if (kPair) umma_commit_pair(empty_bar(s));
          else umma_commit(empty_bar(s));
        }
        __syncwarp();
        if (++s == n_stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) {                             // accumulator complete -> epilogue (of both CTAs)
        if (kPair) umma_commit_pair(tfull_bar(as));
        else umma_commit(tfull_bar(as));
      }
      __syncwarp();
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  } else if (kStaged && warp == 2 && lane == 0 && p.res_tma) {
    // ------------------------------------------------------------ residual prefetch (TMA -> slab)
    const int slabs_per_tile = p.block_n >> 6;
    int slot = 0;
    uint32_t sph = 0;
    for (int tile = t_begin; tile < total_tiles; tile += t_step) {
      const int n_blk = tile % p.n_blocks;
      int mt = tile / p.n_blocks;
      if (kPair) {
        if (mt * 256 >= m_valid) continue;
        mt = 2 * mt + (int)rank;
      }
      const int m0 = mt * 128;
      if (!kPair && m0 >= m_valid) continue;
      for (int j = 0; j < slabs_per_tile; ++j) {
        mbar_wait(sfree_bar(slot), sph ^ 1u);
        mbar_expect_tx(sres_bar(slot), (uint32_t)kSlabBytes);
        tma_load_2d(slab_base + (uint32_t)slot * kSlabBytes, &tmRes, sres_bar(slot),
                    n_blk * p.block_n + j * 64, m0);
        if (++slot == p.nslab) { slot = 0; sph ^= 1u; }
      }
    }
  } else if (kStaged && warp == 3 && lane == 0) {
    // ------------------------------------------------------------ TMA store issuer (slab -> global)
    const int slabs_per_tile = p.block_n >> 6;
    int slot = 0, prev = -1;
    uint32_t sph = 0;
    for (int tile = t_begin; tile < total_tiles; tile += t_step) {
      const int n_blk = tile % p.n_blocks;
      int mt = tile / p.n_blocks;
      if (kPair) {
        if (mt * 256 >= m_valid) continue;
        mt = 2 * mt + (int)rank;
      }
      const int m0 = mt * 128;
      if (!kPair && m0 >= m_valid) continue;
      for (int j = 0; j < slabs_per_tile; ++j) {
        mbar_wait(sready_bar(slot), sph);
        // (pair mode: the second CTA's tile can lie wholly past the valid rows; it still runs the barrier protocol)
        if (!kPair || m0 < m_valid)
          tma_store_2d(&tmOut, slab_base + (uint32_t)slot * kSlabBytes, n_blk * p.block_n + j * 64, m0);
        tma_store_commit();
        if (prev >= 0) {
          tma_store_wait_read<1>();     // every store but the one just issued has drained its slab
          mbar_arrive(sfree_bar(prev));
        }
        prev = slot;
        if (++slot == p.nslab) { slot = 0; sph ^= 1u; }
      }
    }
    if (prev >= 0) {
      tma_store_wait_read<0>();
      mbar_arrive(sfree_bar(prev));
    }
    tma_store_wait_all();
  } else if (kStaged && warp >= 4) {
    // ------------------------------------------------------------ epilogue through shared-memory slabs
    // 8 warps: TMEM lane quarter q = warp % 4; warps 4-7 take channels 0-31 of every 64-channel slab,
    // warps 8-11 channels 32-63. Per slab a thread does one tcgen05.ld.x32, bias / residual adds as packed
    // f32x2, a fused relu + bf16x2 convert and four 16-byte swizzled shared stores.
    const int q = warp & 3;
    const int hh = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int slabs_per_tile = p.block_n >> 6;
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const bool has_res = p.res != nullptr;
    int as = 0, slot = 0;
    uint32_t aph = 0, sph = 0;
    for (int tile = t_begin; tile < total_tiles; tile += t_step) {
      const int n_blk = tile % p.n_blocks;
      int mt = tile / p.n_blocks;
      if (kPair) {
        if (mt * 256 >= m_valid) continue;
        mt = 2 * mt + (int)rank;
      }
      const int m0 = mt * 128;
      if (!kPair && m0 >= m_valid) continue;
      const int c_base = n_blk * p.block_n + hh * 32;
      uint32_t res_off = 0;                       // element offset of this row's residual pixel (gathered mode)
      if (has_res && !p.res_tma) {               // top-down add (res_shift) or an irregular residual view
        int m = m0 + row;
        if (m > p.N * hw_out - 1) m = p.N * hw_out - 1;
        const int on = m / hw_out;
        const int rem = m - on * hw_out;
        const int oy = rem / p.W_out, ox = rem - oy * p.W_out;
        res_off = (uint32_t)(on * p.res_sn + (long long)(oy >> p.res_shift) * p.res_sy +
                             (long long)(ox >> p.res_shift) * p.res_sx) + (uint32_t)c_base;
      }
      bool acc_ready = false;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.acc_stride + hh * 32);
      // gathered residual: rows r = 8*i + lane/4 (i = 0..3), 16-byte chunk lane%4 of this warp's 64-byte half.
      // The loads for slab j+1 are issued before slab j is processed, so only the first slab of a tile
      // exposes their latency (and that one overlaps the wait for the accumulator).
      const bool gather = has_res && !p.res_tma;
      uint32_t goff[4];
      uint4 gv[4];
      if (gather) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          goff[i] = __shfl_sync(0xffffffffu, res_off, 8 * i + (lane >> 2));
          gv[i] = __ldg(reinterpret_cast<const uint4*>(p.res + goff[i]) + (lane & 3));
        }
      }
      for (int j = 0; j < slabs_per_tile; ++j) {
        const uint32_t slab = slab_base + (uint32_t)slot * kSlabBytes;
        float4 bv[8];
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + c_base + j * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = __ldg(b4 + i);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (p.res_tma) {
          mbar_wait(sres_bar(slot), sph);
        } else {
          mbar_wait(sfree_bar(slot), sph ^ 1u);
          if (gather) {
            // the warp copies its 32 rows x 64 B into the slab with coalesced 64-byte row segments (4 lanes
            // per row, 8 rows per request); every thread then reads its own row back like a TMA residual
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t rr = (uint32_t)(q * 32 + 8 * i + (lane >> 2));
              st_shared_v4(slab + rr * 128u + ((((uint32_t)(hh * 4 + (lane & 3))) ^ (rr & 7u)) << 4), gv[i]);
            }
            __syncwarp();
            if (j + 1 < slabs_per_tile) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                gv[i] = __ldg(reinterpret_cast<const uint4*>(p.res + goff[i] + (j + 1) * 64) + (lane & 3));
            }
          }
        }
        if (!acc_ready) {
          mbar_wait(tfull_bar(as), aph);
          tc_fence_after();
          acc_ready = true;
        }
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)(j * 64), v);
        tmem_ld_wait();
        const uint32_t srow = slab + row_off;
        if (has_res) epi_slab_half<true>(v, bv, srow, sw, hh, p.relu != 0);
        else epi_slab_half<false>(v, bv, srow, sw, hh, p.relu != 0);
        fence_proxy_async();            // generic-proxy slab writes -> visible to the TMA store
        mbar_arrive(sready_bar(slot));
        if (++slot == p.nslab) { slot = 0; sph ^= 1u; }
      }
      tc_fence_before();
      if (kPair) {                                  // one arrive per warp on the LEADER's barrier
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));
      } else {
        mbar_arrive(tempty_bar(as));
      }
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  } else if (!kStaged && warp >= 4) {
    // ------------------------------------------------------------ epilogue, direct global stores
    // 8 warps: lane quarter q = warp % 4; the two warps of a quarter alternate 16-column chunks
    const int q = warp & 3;
    const int hh = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int rows_in_tile = p.im2col ? 128 : p.tw * p.th * p.tn;
    const int rx = p.im2col ? 0 : row % p.tw;
    const int ry = p.im2col ? 0 : (row / p.tw) % p.th;
    const int rn = p.im2col ? 0 : row / (p.tw * p.th);
    int as = 0;
    uint32_t aph = 0;
    for (int tile = t_begin; tile < total_tiles; tile += t_step) {
      const int n_blk = tile % p.n_blocks;
      int mt = tile / p.n_blocks;
      if (kPair) {
        if (mt * 256 >= m_valid) continue;
        mt = 2 * mt + (int)rank;
      }
      int ox, oy, on;
      if (p.im2col) {
        const int m0 = mt * 128;
        if (!kPair && m0 >= m_valid) continue;
        const int m = m0 + row;
        on = m / hw_out;
        const int rem = m - on * hw_out;
        oy = rem / p.W_out;
        ox = rem - oy * p.W_out;
      } else {
        const int twi = mt % p.tiles_w;
        const int thi = (mt / p.tiles_w) % p.tiles_h;
        const int tni = mt / (p.tiles_w * p.tiles_h);
        if (tni * p.tn >= nvalid) continue;
        ox = twi * p.tw + rx;
        oy = thi * p.th + ry;
        on = tni * p.tn + rn;
      }
      const bool valid = row < rows_in_tile && ox < p.W_out && oy < p.H_out && on < nvalid;
      const int c_base = n_blk * p.block_n;
      const long long out_off = on * p.out_sn + oy * p.out_sy + ox * p.out_sx + c_base;
      const long long out_off_planar = on * p.out_sn + oy * p.out_sy + ox * p.out_sx + (long long)c_base * p.out_sc;
      const __nv_bfloat16* res_ptr = nullptr;
      const __nv_bfloat16* res2_ptr = nullptr;
      if (p.res != nullptr && valid) {
        const long long ro = on * p.res_sn + (long long)(oy >> p.res_shift) * p.res_sy +
                             (long long)(ox >> p.res_shift) * p.res_sx + c_base;
        res_ptr = p.res + ro;
        if (p.res2 != nullptr) res2_ptr = p.res2 + ro;
      }
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.acc_stride);
      for (int c0 = hh * 16; c0 < p.block_n; c0 += 32) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(t_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + c_base + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b = __ldg(b4 + i);
              f[4 * i + 0] += b.x; f[4 * i + 1] += b.y; f[4 * i + 2] += b.z; f[4 * i + 3] += b.w;
            }
          }
          if (res2_ptr != nullptr) {
            // strict mode: the residual is hi + lo (exact in fp32), added once like the reference's fp32 add
            const uint4* r4 = reinterpret_cast<const uint4*>(res_ptr + c0);
            const uint4* l4 = reinterpret_cast<const uint4*>(res2_ptr + c0);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint4 r = __ldg(r4 + i), l = __ldg(l4 + i);
              f[8 * i + 0] += bf16_lo(r.x) + bf16_lo(l.x); f[8 * i + 1] += bf16_hi(r.x) + bf16_hi(l.x);
              f[8 * i + 2] += bf16_lo(r.y) + bf16_lo(l.y); f[8 * i + 3] += bf16_hi(r.y) + bf16_hi(l.y);
              f[8 * i + 4] += bf16_lo(r.z) + bf16_lo(l.z); f[8 * i + 5] += bf16_hi(r.z) + bf16_hi(l.z);
              f[8 * i + 6] += bf16_lo(r.w) + bf16_lo(l.w); f[8 * i + 7] += bf16_hi(r.w) + bf16_hi(l.w);
            }
          } else if (res_ptr != nullptr) {
            const uint4* r4 = reinterpret_cast<const uint4*>(res_ptr + c0);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint4 r = __ldg(r4 + i);
              f[8 * i + 0] += bf16_lo(r.x); f[8 * i + 1] += bf16_hi(r.x);
              f[8 * i + 2] += bf16_lo(r.y); f[8 * i + 3] += bf16_hi(r.y);
              f[8 * i + 4] += bf16_lo(r.z); f[8 * i + 5] += bf16_hi(r.z);
              f[8 * i + 6] += bf16_lo(r.w); f[8 * i + 7] += bf16_hi(r.w);
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
          }
          if (p.out_sc != 1) {
            // channel-planar fp32 output: lanes are consecutive pixels, so each store is one 128-byte row
            float* o = reinterpret_cast<float*>(p.out) + out_off_planar + (long long)c0 * p.out_sc;
#pragma unroll
            for (int i = 0; i < 16; ++i) o[(long long)i * p.out_sc] = f[i];
          } else if (p.out_fp32) {
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } else {
            uint4* o4 =
                reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + out_off + c0);
            uint4* l4 = p.out2 == nullptr ? nullptr :
                reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + out_off + c0);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint4 o;
              o.x = pack_bf16(f[8 * i + 0], f[8 * i + 1]);
              o.y = pack_bf16(f[8 * i + 2], f[8 * i + 3]);
              o.z = pack_bf16(f[8 * i + 4], f[8 * i + 5]);
              o.w = pack_bf16(f[8 * i + 6], f[8 * i + 7]);
              o4[i] = o;
              if (l4 != nullptr) {            // strict mode: lo = bf16(x - hi)
                uint4 l;
                l.x = pack_bf16(f[8 * i + 0] - bf16_lo(o.x), f[8 * i + 1] - bf16_hi(o.x));
                l.y = pack_bf16(f[8 * i + 2] - bf16_lo(o.y), f[8 * i + 3] - bf16_hi(o.y));
                l.z = pack_bf16(f[8 * i + 4] - bf16_lo(o.z), f[8 * i + 5] - bf16_hi(o.z));
                l.w = pack_bf16(f[8 * i + 6] - bf16_lo(o.w), f[8 * i + 7] - bf16_hi(o.w));
                l4[i] = l;
              }
            }
          }
        }
      }
      tc_fence_before();
      if (kPair) {                                  // one arrive per warp on the LEADER's barrier
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));
      } else {
        mbar_arrive(tempty_bar(as));
      }
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncwarp();
  if (kPair) cluster_sync_divergent();             // the peer's shared memory / TMEM / barriers stay alive until both are done
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int encode_tiled_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr) {
  PFN_encodeTiled fn = get_encode_tiled();
  if (!fn) return -1;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu,%llu box "
              "%u,%u,%u,%u)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0);
    return -2;
  }
  return 0;
}

typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                     cuuint32_t, cuuint32_t, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_im2col_bf16(CUtensorMap* tm, const void* base, const uint64_t* dims,
                       const uint64_t* strides_bytes, const int* lower, const int* upper,
                       uint32_t channels, uint32_t pixels, const uint32_t* estr, int swizzle128) {
  static PFN_encodeIm2col fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeIm2col unavailable (%s)", cudaGetErrorString(e));
      return -1;
    }
    fn = reinterpret_cast<PFN_encodeIm2col>(p);
  }
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                  strides_bytes, lower, upper, channels, pixels, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed: CUresult %d (dims %llu,%llu,%llu,%llu strides "
              "%llu,%llu,%llu lower %d,%d upper %d,%d estr %u,%u)",
              (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)dims[2], (unsigned long long)dims[3],
              (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1],
              (unsigned long long)strides_bytes[2], lower[0], lower[1], upper[0], upper[1], estr[1],
              estr[2]);
    return -2;
  }
  return 0;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Pick the (tw, th, tn) pixel box that minimises the number of 128-row MMA tiles.
static void choose_tile(int W, int H, int N, int sx, int sy, int* tw_o, int* th_o, int* tn_o) {
  long long best_cost = -1;
  int btw = 1, bth = 1, btn = 1;
  const int max_tw = 256 / sx < 128 ? 256 / sx : 128;
  const int max_th = 256 / sy < 128 ? 256 / sy : 128;
  for (int tw = 1; tw <= W && tw <= max_tw; ++tw) {
    for (int th = 1; th <= H && th <= max_th && tw * th <= 128; ++th) {
      int tn = 1;
      if (tw == W && th == H) {
        tn = 128 / (tw * th);
        if (tn > N) tn = N;
        if (tn < 1) tn = 1;
      }
      long long cost = (long long)ceil_div(W, tw) * ceil_div(H, th) * ceil_div(N, tn);
      // tie-break: wider boxes (longer contiguous runs)
      if (best_cost < 0 || cost < best_cost || (cost == best_cost && tw > btw)) {
        best_cost = cost; btw = tw; bth = th; btn = tn;
      }
    }
  }
  *tw_o = btw; *th_o = bth; *tn_o = btn;
}

int conv_plan_build(ConvPlan* plan, const ConvDesc& d, int num_sms) {
  if (!d.x || !d.w || !d.out) { set_error("conv: null pointer"); return -1; }
  if (d.cin_pad % 64 != 0 || d.cout_pad % 16 != 0) {
    set_error("conv: cin_pad %d must be a multiple of 64 and cout_pad %d of 16", d.cin_pad,
              d.cout_pad);
    return -1;
  }
  ConvKParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  // N tile
  int block_n = d.block_n;
  if (block_n == 0) {
    if (d.cout_pad <= 256) block_n = d.cout_pad;
    else {
      // largest multiple of 16 that divides cout_pad and is <= 256
      block_n = 16;
      for (int b = 256; b >= 16; b -= 16)
        if (d.cout_pad % b == 0) { block_n = b; break; }
    }
  }
  if (d.block_n == 0 && !d.phase_taps) {
    // Few-tile launches (batch 1: res4 / res5 / FPN p5 / FC1 have 8-33 row tiles): a 256-wide N tile leaves most SMs
    // idle while each busy one walks the whole K loop at 512 cycles per chunk. Narrower N tiles put 2-4x the CTAs to
    // work and a 64-wide tile's chunk costs ~330 cycles (its A load), so the launch's critical path shrinks by up to a
    // half; the extra A traffic is free while the machine is not full. Results do not depend on the N tile.
    long long m_tiles;
    if (d.im2col) {
      m_tiles = ((long long)d.N * d.H_out * d.W_out + 127) / 128;
    } else {
      int tw, th, tn;
      choose_tile(d.W_out, d.H_out, d.N, d.sx, d.sy, &tw, &th, &tn);
      m_tiles = (long long)ceil_div(d.W_out, tw) * ceil_div(d.H_out, th) * ceil_div(d.N, tn);
    }
    while (block_n >= 128 && (block_n / 2) % 64 == 0 && m_tiles * (d.cout_pad / block_n) * 2 <= num_sms) block_n /= 2;
  }
  if (d.phase_taps) {
    if (d.cout_pad % 64 != 0 || d.kh != 2 || d.kw != 2 || d.pad_x != 1 || d.pad_y != 1 || d.sx != 1 || d.sy != 1) {
      set_error("conv: phase_taps needs k=2, pad=1, stride 1 and cout_pad = 4 x (multiple of 16)");
      return -1;
    }
    block_n = d.cout_pad / 4;
  }
  if (block_n % 16 != 0 || block_n > 256 || d.cout_pad % block_n != 0) {
    set_error("conv: bad block_n %d for cout_pad %d", block_n, d.cout_pad);
    return -1;
  }
  p.block_n = block_n;
  p.n_blocks = d.cout_pad / block_n;
  p.acc_stride = block_n <= 32 ? 32 : block_n <= 64 ? 64 : block_n <= 128 ? 128 : 256;
  p.im2col = d.im2col;
  if (d.im2col) {
    const long long m_total = (long long)d.N * d.H_out * d.W_out;   // (re-declared below for the epilogue)
    if (m_total >= (1LL << 31) - 128) { set_error("conv: too many output pixels"); return -1; }
    p.tw = 128; p.th = 1; p.tn = 1;
    p.tiles_w = (int)((m_total + 127) / 128); p.tiles_h = 1; p.tiles_n = 1;
  } else {
    choose_tile(d.W_out, d.H_out, d.N, d.sx, d.sy, &p.tw, &p.th, &p.tn);
    p.tiles_w = ceil_div(d.W_out, p.tw);
    p.tiles_h = ceil_div(d.H_out, p.th);
    p.tiles_n = ceil_div(d.N, p.tn);
  }
  p.H_out = d.H_out; p.W_out = d.W_out; p.N = d.N;
  p.kh = d.kh; p.kw = d.kw; p.sx = d.sx; p.sy = d.sy; p.pad_x = d.pad_x; p.pad_y = d.pad_y;
  p.dil = d.dil;
  p.phase_taps = d.phase_taps;
  const bool strict = d.x2 != nullptr;
  p.nseg = strict ? 3 : 1;
  p.res2 = reinterpret_cast<const __nv_bfloat16*>(d.res2);
  p.out2 = d.out2;
  if (strict && d.res != nullptr && d.res2 == nullptr) { set_error("conv: strict mode needs both halves of the residual"); return -1; }
  if (strict && !d.out_fp32 && d.out2 == nullptr) { set_error("conv: strict mode needs both halves of a bf16 output"); return -1; }
  p.cin_chunks = d.cin_pad / 64;
  p.relu = d.relu; p.out_fp32 = d.out_fp32; p.res_shift = d.res_shift;
  p.bias = d.bias;
  p.res = reinterpret_cast<const __nv_bfloat16*>(d.res);
  p.res_sn = d.res_sn; p.res_sy = d.res_sy; p.res_sx = d.res_sx;
  p.out = d.out; p.out_sn = d.out_sn; p.out_sy = d.out_sy; p.out_sx = d.out_sx;
  p.out_sc = d.out_sc > 0 ? d.out_sc : 1;
  p.n_valid = d.n_valid;
  if (p.out_sc != 1 && !d.out_fp32) { set_error("conv: planar output must be fp32"); return -1; }

  const int k_iters = d.kh * d.kw * p.cin_chunks * p.nseg;
  const long long m_total = (long long)d.N * d.H_out * d.W_out;
  // Staged epilogue (smem slabs + TMA): bf16 output that is a plain [M, C] matrix (row stride out_sx) in
  // im2col row order. Measured faster than direct stores for every shape on the path (3x on the K=64 1x1
  // convs, +4% on the K=4608 head convs), so it is the default whenever the output qualifies.
  const bool out_matrix = d.im2col && !d.out_fp32 && p.out_sc == 1 && block_n % 64 == 0 &&
                          d.out_sy == d.out_sx * d.W_out && d.out_sn == d.out_sy * d.H_out &&
                          d.out_sx % 8 == 0 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0;
  bool staged = out_matrix && d.epilogue != 1 && !strict;     // strict mode writes two tensors: direct stores
  if (d.epilogue == 2 && !out_matrix) { set_error("conv: staged epilogue needs a bf16 [M, C] output"); return -1; }
  const bool res_matrix = d.res != nullptr && d.res_shift == 0 && d.res_sy == d.res_sx * d.W_out &&
                          d.res_sn == d.res_sy * d.H_out && d.res_sx % 8 == 0 &&
                          (reinterpret_cast<uintptr_t>(d.res) & 15) == 0;
  plan->staged = staged ? 1 : 0;
  p.res_tma = (staged && res_matrix) ? 1 : 0;
  p.nslab = 0;
  // slab ring of the staged epilogue (same-box sweep over 2 / 3 / 4 / 6 / 8 slabs on the path's shapes, round 2): four
  // slabs, or two where the shared memory is better spent on pipeline stages - the long-K convs and the narrow 3x3s
  // (res2's 64->64: one more A stage is worth 14 %). Eight residual-prefetch slabs for the K <= 2 bottleneck outputs
  // (round 1's choice) lose 5-8 % to four.
  if (staged) p.nslab = (k_iters > 16 || (k_iters > 8 && block_n <= 128)) ? 2 : 4;

  // CTA pairs (cta_group::2): staged im2col convs whose N tile splits into two halves of a multiple of 16 rows.
  // Chosen automatically for the long-K, 256-wide tiles (the 3x3 256->256 / 512->512 convs), where halving the
  // weight traffic from L2 pays; d.pair forces it on (2) or off (1).
  const bool pair_ok = d.im2col && block_n % 16 == 0 && block_n >= 32 && (num_sms & ~1) >= 2 && !strict;
  if (d.pair == 2 && !pair_ok) { set_error("conv: pair mode needs an im2col conv with block_n %% 16 == 0"); return -1; }
  // (measured on the path's shapes: short-K tiles lose to the pair's extra barrier traffic; K >= 18 chunks gains 6-10 %)
  const bool pair = d.pair == 2 || (d.pair == 0 && pair_ok && staged && block_n == 256 && k_iters >= 18 && m_total >= 4096);
  plan->pair = pair ? 1 : 0;
  const int b_rows = pair ? block_n / 2 : block_n;
  const int b_bytes = (b_rows * 128 + 1023) & ~1023;
  const int stage_bytes = kABytes + b_bytes;
  // Narrow N tiles do little tensor work per 64-channel chunk (2*N clocks), less than one trip of the producer /
  // MMA-issuer loops costs: group two chunks behind one barrier pair so the per-trip overhead is paid half as often.
  // (only when at least three such groups fit beside the epilogue slabs: with two the producer cannot run ahead)
  const int fixed_bytes = 1024 /*align slack*/ + p.nslab * kSlabBytes + 8 * (2 * 8 + 4 + 3 * p.nslab) + 16;
  int ks = (block_n <= 128 && k_iters >= 2 && (227 * 1024 - fixed_bytes) / (2 * stage_bytes) >= 3) ? 2 : 1;
  if (d.ks == 1 || d.ks == 2) ks = d.ks;
  if (ks == 2 && k_iters < 2) ks = 1;
  if (pair) ks = 1;
  p.ks = ks;
  const int n_groups = (k_iters + ks - 1) / ks;
  int stages = d.stages;
  if (stages == 0) {
    stages = (227 * 1024 - fixed_bytes) / (ks * stage_bytes);
    if (stages > 8) stages = 8;
    if (stages > n_groups + 1) stages = n_groups + 1;
    if (stages < 2) stages = 2;
  }
  p.stages = stages;
  plan->smem = stages * ks * stage_bytes + p.nslab * kSlabBytes + 1024 /*align slack*/ +
               8 * (2 * stages + 4 + 3 * p.nslab) + 16;
  if (plan->smem > 227 * 1024) { set_error("conv: %d B of shared memory needed (block_n %d, stages %d x %d chunks, slabs %d)", plan->smem, block_n, stages, ks, p.nslab); return -1; }

  // A: 4-D (C, W, H, N) view of the input; box = (64 ch, tw, th, tn) output pixels, traversal
  // strides implement the convolution stride.
  {
    uint64_t dims[4] = {(uint64_t)d.Cin, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.N};
    uint64_t strides[3] = {(uint64_t)d.x_sw * 2, (uint64_t)d.x_sh * 2, (uint64_t)d.x_sn * 2};
    uint32_t box[4] = {64, (uint32_t)(p.tw * d.sx), (uint32_t)(p.th * d.sy), (uint32_t)p.tn};
    uint32_t estr[4] = {1, (uint32_t)d.sx, (uint32_t)d.sy, 1};
    // degenerate outer dims must still carry a legal (multiple of 16 B) stride
    for (int i = 0; i < 3; ++i)
      if (strides[i] == 0) strides[i] = (i == 0 ? (uint64_t)d.Cin * 2 : strides[i - 1]);
    int r;
    if (d.im2col) {
      // Bounding box [lower, dim-1+upper] holds exactly the W_out x H_out base pixels (spaced by
      // the traversal stride); the per-tap offsets (kx*dil, ky*dil) are given at load time.
      int lower[2] = {-d.pad_x, -d.pad_y};
      int upper[2] = {(d.W_out - 1) * d.sx - d.pad_x - (d.W - 1),
                      (d.H_out - 1) * d.sy - d.pad_y - (d.H - 1)};
      r = encode_im2col_bf16(&plan->tmA, d.x, dims, strides, lower, upper, 64, 128, estr, 1);
    } else {
      r = encode_tiled_bf16(&plan->tmA, d.x, 4, dims, strides, box, estr);
    }
    if (r) return r;
    plan->tmA2 = plan->tmA;
    if (strict) {
      if (d.im2col) {
        int lower[2] = {-d.pad_x, -d.pad_y};
        int upper[2] = {(d.W_out - 1) * d.sx - d.pad_x - (d.W - 1),
                        (d.H_out - 1) * d.sy - d.pad_y - (d.H - 1)};
        r = encode_im2col_bf16(&plan->tmA2, d.x2, dims, strides, lower, upper, 64, 128, estr, 1);
      } else {
        r = encode_tiled_bf16(&plan->tmA2, d.x2, 4, dims, strides, box, estr);
      }
      if (r) return r;
    }
  }
  {
    const uint64_t K = (uint64_t)d.kh * d.kw * p.nseg * d.cin_pad;
    uint64_t dims[2] = {K, (uint64_t)d.cout_pad};
    uint64_t strides[1] = {K * 2};
    uint32_t box[2] = {64, (uint32_t)b_rows};
    uint32_t estr[2] = {1, 1};
    int r = encode_tiled_bf16(&plan->tmB, d.w, 2, dims, strides, box, estr);
    if (r) return r;
  }
  if (staged) {
    uint64_t dims[2] = {(uint64_t)d.cout_pad, (uint64_t)m_total};
    uint32_t box[2] = {64, 128};
    uint32_t estr[2] = {1, 1};
    uint64_t so[1] = {(uint64_t)d.out_sx * 2};
    int r = encode_tiled_bf16(&plan->tmOut, d.out, 2, dims, so, box, estr);
    if (r) return r;
    if (p.res_tma) {
      uint64_t sr[1] = {(uint64_t)d.res_sx * 2};
      r = encode_tiled_bf16(&plan->tmRes, d.res, 2, dims, sr, box, estr);
      if (r) return r;
    } else {
      plan->tmRes = plan->tmOut;
    }
  } else {
    plan->tmOut = plan->tmB;   // unused by the direct epilogue; any valid descriptor
    plan->tmRes = plan->tmB;
  }
  const long long total_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n * p.n_blocks;
  plan->grid = (int)(total_tiles < num_sms ? total_tiles : num_sms);
  if (plan->grid < 1) plan->grid = 1;
  if (pair) {
    const long long pair_tiles = (((long long)p.tiles_w + 1) / 2) * p.n_blocks;
    const long long ctas = 2 * pair_tiles;
    plan->grid = (int)(ctas < (num_sms & ~1) ? ctas : (num_sms & ~1));
  }
  plan->flops = 2.0 * d.N * d.H_out * d.W_out * (double)d.cout_pad * d.kh * d.kw * d.Cin;
  return 0;
}

int conv_kernels_init() {
  cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<false, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_igemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_igemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_igemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024);
  if (e != cudaSuccess) { set_error("conv: smem attr: %s", cudaGetErrorString(e)); return -3; }
  return 0;
}

int conv_plan_launch(const ConvPlan& plan, cudaStream_t stream) {
  static thread_local int init_dev = -1;     // per host thread: the device the attribute was last set on
  int dev = 0;
  cudaGetDevice(&dev);
  if (init_dev != dev) {
    if (conv_kernels_init()) return -3;
    init_dev = dev;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)plan.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)plan.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  // programmatic dependent launch (the kernel calls griddepcontrol.wait before it reads anything)
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (plan.pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)na;
  cudaError_t le;
  if (plan.pair && plan.staged)
    le = cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true, true>, plan.tmA, plan.tmB, plan.tmOut, plan.tmRes, plan.tmA2, plan.p);
  else if (plan.pair)
    le = cudaLaunchKernelEx(&cfg, conv_igemm_kernel<false, true>, plan.tmA, plan.tmB, plan.tmOut, plan.tmRes, plan.tmA2, plan.p);
  else if (plan.staged)
    le = cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true, false>, plan.tmA, plan.tmB, plan.tmOut, plan.tmRes, plan.tmA2, plan.p);
  else
    le = cudaLaunchKernelEx(&cfg, conv_igemm_kernel<false, false>, plan.tmA, plan.tmB, plan.tmOut, plan.tmRes, plan.tmA2, plan.p);
  if (le != cudaSuccess) { set_error("conv launch: %s", cudaGetErrorString(le)); return -4; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("conv launch: %s", cudaGetErrorString(e)); return -4; }
  return 0;
}

}  // namespace dpb
