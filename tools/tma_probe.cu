// Bring-up probe (not part of the product): dumps what TMA im2col / strided / overlapping-stride
// tensor maps actually deliver to shared memory on a B200, so the conv producer's coordinate
// conventions are pinned by observation.  Build: see tools/build_probe.sh.  Run under gpurun.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../densepose_torchscript_b200/csrc/ptx.cuh"

namespace dpb {
int encode_tiled_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr);
int encode_im2col_bf16(CUtensorMap* tm, const void* base, const uint64_t* dims,
                       const uint64_t* strides_bytes, const int* lower, const int* upper,
                       uint32_t channels, uint32_t pixels, const uint32_t* estr, int swizzle128);
const char* get_error();
}  // namespace dpb
using namespace dpb;

// mode 0: tiled 4d; mode 1: im2col 4d
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, int mode, int c0, int c1, int c2,
                             int c3, int offw, int offh, int bytes, __nv_bfloat16* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t dst = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* dst_ptr = smem + (dst - smem_u32(smem));
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x)
    reinterpret_cast<__nv_bfloat16*>(dst_ptr)[i] = __float2bfloat16(-9.0f);
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
    mbar_expect_tx(smem_u32(&bar), bytes);
    if (mode == 0) tma_load_4d(dst, &tm, smem_u32(&bar), c0, c1, c2, c3);
    else tma_load_im2col_4d(dst, &tm, smem_u32(&bar), c0, c1, c2, c3, (uint16_t)offw, (uint16_t)offh);
    mbar_wait(smem_u32(&bar), 0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x)
    out[i] = reinterpret_cast<__nv_bfloat16*>(dst_ptr)[i];
}

static __nv_bfloat16* d_out;
static std::vector<__nv_bfloat16> h_out;

// undo the 128B swizzle for printing: 16-byte chunk index ^= (row & 7)
static float at(int row, int ch) {
  int chunk = ch / 8, within = ch % 8;
  int phys = (chunk ^ (row & 7)) * 8 + within;
  return __bfloat162float(h_out[row * 64 + phys]);
}

static void run(const char* title, const CUtensorMap& tm, int mode, int c0, int c1, int c2, int c3,
                int offw, int offh, int rows, int chans_to_print) {
  const int bytes = rows * 128;
  cudaMemset(d_out, 0, 256 * 128);
  probe_kernel<<<1, 128, 40 * 1024>>>(tm, mode, c0, c1, c2, c3, offw, offh, bytes, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("== %s  coords(%d,%d,%d,%d) off(%d,%d): %s\n", title, c0, c1, c2, c3, offw, offh,
         cudaGetErrorString(e));
  if (e != cudaSuccess) exit(3);
  cudaMemcpy(h_out.data(), d_out, bytes, cudaMemcpyDeviceToHost);
  for (int r = 0; r < rows; ++r) {
    printf("[%2d:", r);
    for (int c = 0; c < chans_to_print; ++c) printf(" %g", at(r, c));
    printf("] ");
    if (r % 4 == 3) printf("\n");
  }
  printf("\n");
}

int main() {
  cudaMalloc(&d_out, 256 * 128);
  h_out.resize(256 * 64);
  // ---- A: NHWC tensor N=2,H=5,W=7,C=64 : ch0=n+1, ch1=y+1, ch2=x+1
  const int N = 2, H = 5, W = 7, C = 64;
  std::vector<__nv_bfloat16> h(N * H * W * C, __float2bfloat16(0.f));
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        size_t o = ((size_t)(n * H + y) * W + x) * C;
        h[o] = __float2bfloat16(n + 1.f); h[o + 1] = __float2bfloat16(y + 1.f);
        h[o + 2] = __float2bfloat16(x + 1.f);
        for (int c = 3; c < C; ++c) h[o + c] = __float2bfloat16((float)c);
      }
  __nv_bfloat16* d_x;
  cudaMalloc(&d_x, h.size() * 2 + 4096);
  cudaMemset(d_x, 0, h.size() * 2 + 4096);
  cudaMemcpy(d_x, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  uint64_t dims[4] = {C, W, H, N};
  uint64_t strides[3] = {C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  {
    CUtensorMap tm;
    int lower[2] = {-1, -1}, upper[2] = {-1, -1};
    uint32_t estr[4] = {1, 1, 1, 1};
    int r = encode_im2col_bf16(&tm, d_x, dims, strides, lower, upper, 64, 32, estr, 1);
    printf("encode im2col 3x3p1: %d %s\n", r, r ? get_error() : "");
    if (!r) {
      run("im2col 3x3 pad1 tap(0,0)", tm, 1, 0, -1, -1, 0, 0, 0, 32, 3);
      run("im2col 3x3 pad1 tap(1,1)", tm, 1, 0, -1, -1, 0, 1, 1, 32, 3);
      run("im2col 3x3 pad1 tap(2,2)", tm, 1, 0, -1, -1, 0, 2, 2, 32, 3);
      run("im2col 3x3 pad1 midrow tap(1,1) base=(ox3,oy1)", tm, 1, 0, 2, 0, 0, 1, 1, 32, 3);
      run("im2col 3x3 pad1 tail n=1 base=(ox0,oy3) tap(1,1)", tm, 1, 0, -1, 2, 1, 1, 1, 32, 3);
    }
  }
  {
    CUtensorMap tm;  // 1x1 stride 2 on H=5? use W=7,H=5 -> W_out=4,H_out=3 ; upper=(W_out-1)*2-(W-1)
    int lower[2] = {0, 0}, upper[2] = {(4 - 1) * 2 - (W - 1), (3 - 1) * 2 - (H - 1)};
    uint32_t estr[4] = {1, 2, 2, 1};
    int r = encode_im2col_bf16(&tm, d_x, dims, strides, lower, upper, 64, 32, estr, 1);
    printf("encode im2col 1x1s2: %d %s\n", r, r ? get_error() : "");
    if (!r) run("im2col 1x1 stride2 (expect x=1,3,5,7 y=1,3,5)", tm, 1, 0, 0, 0, 0, 0, 0, 32, 3);
  }
  {
    CUtensorMap tm;  // 3x3 dilation 2, pad 2
    int lower[2] = {-2, -2}, upper[2] = {-2, -2};
    uint32_t estr[4] = {1, 1, 1, 1};
    int r = encode_im2col_bf16(&tm, d_x, dims, strides, lower, upper, 64, 32, estr, 1);
    printf("encode im2col 3x3 d2 p2: %d %s\n", r, r ? get_error() : "");
    if (!r) run("im2col dil2 tap(2,1)->off(4,2)", tm, 1, 0, -2, -2, 0, 4, 2, 32, 3);
  }
  {
    CUtensorMap tm;  // tiled, traversal stride 2: box (64, 4*2, 3*2, 1)
    uint32_t box[4] = {64, 8, 6, 1}, estr[4] = {1, 2, 2, 1};
    int r = encode_tiled_bf16(&tm, d_x, 4, dims, strides, box, estr);
    printf("encode tiled s2: %d %s\n", r, r ? get_error() : "");
    if (!r) run("tiled stride2 box 4x3 at (0,0)", tm, 0, 0, 0, 0, 0, 0, 0, 12, 3);
    if (!r) run("tiled stride2 box 4x3 at (-1,-1)", tm, 0, 0, -1, -1, 0, 0, 0, 12, 3);
  }
  // ---- D: overlapping-window virtual tensor (stem trick). image [H=4][Wp=40][4ch]: ch0=y+1, ch1=x+1
  {
    const int Hs = 4, Wp = 40;
    std::vector<__nv_bfloat16> hs(Hs * Wp * 4 + 256, __float2bfloat16(0.f));
    for (int y = 0; y < Hs; ++y)
      for (int x = 0; x < Wp; ++x) {
        hs[(y * Wp + x) * 4 + 0] = __float2bfloat16(y + 1.f);
        hs[(y * Wp + x) * 4 + 1] = __float2bfloat16(x + 1.f);
      }
    __nv_bfloat16* d_s;
    cudaMalloc(&d_s, hs.size() * 2);
    cudaMemcpy(d_s, hs.data(), hs.size() * 2, cudaMemcpyHostToDevice);
    // virtual: dim0 = 64 elements (16 px * 4 ch) contiguous, dim1 = ox (stride 2 px = 16 B), dim2 = y
    uint64_t vd[4] = {64, 12, Hs, 1};
    uint64_t vs[3] = {16, (uint64_t)Wp * 8, (uint64_t)Hs * Wp * 8};
    CUtensorMap tm;
    uint32_t box[4] = {64, 8, 2, 1}, estr[4] = {1, 1, 1, 1};
    int r = encode_tiled_bf16(&tm, d_s, 4, vd, vs, box, estr);
    printf("encode overlapping tiled: %d %s\n", r, r ? get_error() : "");
    if (!r) run("overlap tiled box(8 ox,2 y) at ox=2,y=1: expect row r: y=2+(r/8), x0=2*(2+r%8)+1", tm, 0, 0, 2, 1, 0, 0, 0, 16, 8);
    int lower[2] = {0, -3}, upper[2] = {0, (2 - 1) * 2 - 3 - (Hs - 1)};
    uint32_t e2[4] = {1, 1, 2, 1};
    CUtensorMap tm2;
    r = encode_im2col_bf16(&tm2, d_s, vd, vs, lower, upper, 64, 16, e2, 1);
    printf("encode overlapping im2col (7x1 s2 p3 in y): %d %s\n", r, r ? get_error() : "");
    if (!r) run("overlap im2col tap ky=3 (expect y=1,3 rows; x0=2*ox+1)", tm2, 1, 0, 0, -3, 0, 0, 3, 16, 8);
    if (!r) run("overlap im2col tap ky=0 (expect first 12 rows zero (y=-3), then y=... )", tm2, 1, 0, 0, -3, 0, 0, 0, 16, 8);
  }
  printf("PROBE DONE\n");
  return 0;
}
