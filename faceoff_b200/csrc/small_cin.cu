// First VGG16 convolution of the LPIPS trunk (reference models/lpips.py:96-103,119-127): ScalingLayer + Conv2d(3 -> 64, 3x3,
// pad 1) + ReLU on the fp32 NCHW image, producing the bf16 channels-last activation the rest of the trunk consumes.
//
// With 3 input channels the layer has no arithmetic to speak of (K = 27) and writes 128 bytes per pixel: it is a pure
// store-bandwidth problem.  The former path materialised an im2col matrix [pixels, 32] in HBM (4 GB written and read per
// 960-frame image set) and ran a K = 32 GEMM whose epilogue reached 3.4 TB/s: 5.8 ms per image set for 8.75 GB of
// algorithmic traffic.  Here a CTA builds the [128 pixels x 32] A tile in shared memory straight from the image rows
// (coalesced loads, scaling applied once per loaded element, 64-byte-swizzled K-major layout written by the threads like
// vq_assign's loader), the weights stay resident in shared memory, ONE tcgen05 MMA pair produces the tile in TMEM, and the
// epilogue stages bias + ReLU + bf16 in shared memory so that the 16 KB of the tile (128 consecutive pixels x 64 channels)
// leave as fully coalesced 16-byte stores.  No pipeline inside the CTA: several CTAs per SM overlap each other's phases.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace fo {

constexpr int kFcThreads = 256;
constexpr int kFcTile = 256;          // output pixels per tile: one run of an image row, two M = 128 MMA tiles
constexpr int kFcStagePitch = 264;    // floats per staged image row (258 used: the tile's columns -1 .. +256)

struct FirstConvParams {
  const float* x;        // [n, 3, h, w] fp32
  const float* weight;   // [64, 3, 3, 3] fp32 (PyTorch layout)
  const float* bias;     // [64]
  const float* shift;    // [3] or null   (ScalingLayer: (x - shift) / scale)
  const float* scale;    // [3] or null
  __nv_bfloat16* out;    // relu(conv) bf16 channels-last [n, h, w, 64]
  int n, h, w;
  int tiles_x;           // ceil(w / 256)
  int total_tiles;       // n * h * tiles_x (< 2^31, checked by the launcher)
  FastDiv fd_tx, fd_h;   // tile -> (image, row, column tile) without integer divisions
};

__global__ void __launch_bounds__(kFcThreads, 3)
vgg_first_conv_kernel(const FirstConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // the output staging tile (32 KB) overlays the A tiles and the staged image rows, which are dead by the time the
  // epilogue runs.  Three CTAs per SM, not four: at 64 registers per thread the nine prefetched image values of the next
  // tile were spilled to local memory right behind their loads (ncu: STL on the long scoreboard was the top stall);
  // 80 registers keep them in flight: 2.88 -> 2.48 ms per 960 frames
  uint8_t* sA = smem;                       // 2 x [128 x 64 B]  pixels x k (27 valid), 64B swizzle
  float* sStage = reinterpret_cast<float*>(sA + 2 * 128 * 64);      // [3 ch x 3 rows][kFcStagePitch] (9.3 KB of 16 KB)
  uint8_t* sOut = smem;                     // [256 x 128 B]     pixels x 64 bf16, 16-byte chunks XOR-ed with (row & 7)
  uint8_t* sB = smem + kFcTile * 128;       // [64 x 64 B]       output channels x k
  float* sBias = reinterpret_cast<float*>(sB + 64 * 64);            // [64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBias + 64);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 128);
    tmem_relinquish();
  } else if (tid == 32) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  pdl_trigger();   // programmatic dependent launch (common.cuh): global memory only after pdl_wait()
  pdl_wait();
  // weights -> K-major bf16 tile, k = (ky * 3 + kx) * 3 + ch (zero for k >= 27)
  for (int i = tid; i < 64 * 32; i += kFcThreads) {
    const int co = i >> 5, k = i & 31;
    const float v = k < 27 ? p.weight[co * 27 + (k % 3) * 9 + k / 3] : 0.f;
    const uint32_t off = (uint32_t)co * 64 + ((((uint32_t)k >> 3) ^ (((uint32_t)co >> 1) & 3)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(v);
  }
  if (tid < 64) sBias[tid] = p.bias != nullptr ? p.bias[tid] : 0.f;
  // ScalingLayer as one FMA per loaded element: (v - shift) / scale = v * (1 / scale) - shift / scale (the result is rounded to
  // bf16 right after; the verification mode does not use this kernel)
  float sc_a[3] = {1.f, 1.f, 1.f}, sc_b[3] = {0.f, 0.f, 0.f};
  if (p.shift != nullptr)
    for (int c = 0; c < 3; ++c) { sc_a[c] = 1.f / p.scale[c]; sc_b[c] = -p.shift[c] / p.scale[c]; }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  const uint64_t a_desc = make_smem_desc(smem_u32(sA), 64, 16);
  const uint64_t b_desc = make_smem_desc(smem_u32(sB), 64, 16);
  uint32_t phase = 0;

  // Image values are fetched into registers one tile ahead: thread t owns column x0 - 1 + t of the nine (channel, row)
  // lines (threads 0 / 1 also the two right-most columns), no per-element index arithmetic.  The registers hold the RAW
  // values; the ScalingLayer FMA is applied when they are staged one iteration later -- an FMA right behind the load made
  // the warp wait for its loads before it could start on the current tile (ncu: long-scoreboard stalls on those FFMAs were
  // the top stall reason).  Offsets inside an image are 32-bit (3 * h * w < 2^31, checked by the launcher).
  float pre[9], pre_tail[9];
  uint32_t pre_ok = 0;      // bit r: pre[r] is inside the image (else the conv's zero padding), bit 9 + r: pre_tail[r]
  auto decode = [&](int tile, int& n, int& y, int& x0) {
    int rest, tx;
    p.fd_tx.divmod(tile, rest, tx);
    p.fd_h.divmod(rest, n, y);
    x0 = tx * kFcTile;
  };
  auto fetch = [&](int tile) {
    int n, y, x0;
    decode(tile, n, y, x0);
    const int ix = x0 - 1 + tid, ixt = x0 + 255 + tid;       // tail: columns 256, 257 of the patch (tid < 2)
    const bool okx = (unsigned)ix < (unsigned)p.w, okt = tid < 2 && (unsigned)ixt < (unsigned)p.w;
    const float* img = p.x + (size_t)n * 3 * p.h * p.w;
    const int plane = p.h * p.w;
    pre_ok = 0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y - 1 + ky;
      const bool oky = (unsigned)iy < (unsigned)p.h;
      const int ro = (oky ? iy : 0) * p.w;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const int r = ch * 3 + ky;
        const float* row = img + ch * plane + ro;
        pre[r] = (oky && okx) ? __ldg(row + ix) : 0.f;
        pre_tail[r] = (oky && okt) ? __ldg(row + ixt) : 0.f;
        pre_ok |= (uint32_t)(oky && okx) << r;
        pre_ok |= (uint32_t)(oky && okt) << (9 + r);
      }
    }
  };
  if ((int)blockIdx.x < p.total_tiles) fetch(blockIdx.x);
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    int n, y, x0;
    decode(tile, n, y, x0);
    // ---- the 3 x 3 image rows of the tile (258 columns each) -> shared memory; then start on the next tile's values
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int ch = r / 3;
      sStage[r * kFcStagePitch + tid] = (pre_ok >> r) & 1u ? fmaf(pre[r], sc_a[ch], sc_b[ch]) : 0.f;
      if (tid < 2) sStage[r * kFcStagePitch + 256 + tid] = (pre_ok >> (9 + r)) & 1u ? fmaf(pre_tail[r], sc_a[ch], sc_b[ch]) : 0.f;
    }
    __syncthreads();
    if (tile + (int)gridDim.x < p.total_tiles) fetch(tile + gridDim.x);
    // ---- A tiles: thread = pixel, all 32 k values (four 16-byte chunks)
    {
      const int px = tid;
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int tap = k / 3, ch = k - tap * 3, ky = tap / 3, kx = tap - ky * 3;
        v[k] = k < 27 ? sStage[(ch * 3 + ky) * kFcStagePitch + px + kx] : 0.f;
      }
      uint8_t* dst = sA + (px >> 7) * (128 * 64) + (px & 127) * 64;
      const uint32_t sw = ((uint32_t)(px & 127) >> 1) & 3;
#pragma unroll
      for (int chunk = 0; chunk < 4; ++chunk) {
        uint4 o;
        o.x = pack_bf16x2(v[chunk * 8 + 0], v[chunk * 8 + 1]); o.y = pack_bf16x2(v[chunk * 8 + 2], v[chunk * 8 + 3]);
        o.z = pack_bf16x2(v[chunk * 8 + 4], v[chunk * 8 + 5]); o.w = pack_bf16x2(v[chunk * 8 + 6], v[chunk * 8 + 7]);
        *reinterpret_cast<uint4*>(dst + (((uint32_t)chunk ^ sw) << 4)) = o;
      }
    }
    fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const uint64_t am = a_desc + (uint64_t)((m * 128 * 64) >> 4);
        umma_bf16(tmem + m * 64, am, b_desc, idesc, 0);
        umma_bf16(tmem + m * 64, am + 2, b_desc + 2, idesc, 1);
      }
      umma_commit(bar);
    }
    __syncwarp();   // warp 0 reconverges before anyone waits
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: warp w -> M tile w / 4, TMEM lane quarter w % 4 (lane = pixel), all 64 channels
    {
      const int quarter = warp & 3, m = warp >> 2;
      const int row = m * 128 + quarter * 32 + lane;
#pragma unroll
      for (int hcol = 0; hcol < 2; ++hcol) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + m * 64 + hcol * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b0 = *reinterpret_cast<const float4*>(sBias + hcol * 32 + q * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(sBias + hcol * 32 + q * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]) + bb[e];
          uint4 o;   // ReLU inside the packed conversion (cvt.rn.relu.bf16x2.f32)
          o.x = pack_bf16x2_relu(f[0], f[1]); o.y = pack_bf16x2_relu(f[2], f[3]);
          o.z = pack_bf16x2_relu(f[4], f[5]); o.w = pack_bf16x2_relu(f[6], f[7]);
          const int chunk = hcol * 4 + q;
          *reinterpret_cast<uint4*>(sOut + row * 128 + ((chunk ^ (row & 7)) << 4)) = o;
        }
      }
      tc_fence_before();
    }
    __syncthreads();
    // ---- the tile's 256 pixels x 128 B are contiguous in the channels-last output: coalesced 16-byte stores
    {
      uint4* dst = reinterpret_cast<uint4*>(p.out + (((size_t)n * p.h + y) * p.w + x0) * 64);
#pragma unroll
      for (int i = tid; i < kFcTile * 8; i += kFcThreads) {
        const int row = i >> 3, chunk = i & 7;
        if (x0 + row < p.w) dst[i] = *reinterpret_cast<const uint4*>(sOut + row * 128 + ((chunk ^ (row & 7)) << 4));
      }
    }
    __syncthreads();   // sOut overlays sStage / sA, which the next iteration writes first
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// =====================================================================================================================
// Data gradient of the first VGG conv (Conv2d 3 -> 64, 3x3, pad 1; reference models/lpips.py:119-127 slice1, the gradient
// that leaves the LPIPS trunk towards the decoder).  As an implicit GEMM it is the worst shape there is: N = 3 (padded to
// 16) output channels against K = 9 taps x 64, i.e. 36 MMAs of 128 x 16 x 16 per 128 pixels, each paying the 64-cycle
// A-operand read -- 5.3 ms per 960 frames for 0.2 TFLOP (+ 0.4 ms to unpack the bf16 channels-last result to NCHW).
// Here the taps move from K to N:
//     T[p][tap][c] = sum_co dy[p][co] * W[co][c][tap]          ONE K = 64 GEMM per pixel tile, N = 9 taps x (3 + 1 pad) = 36 (48)
//     dx[q][c]     = sum_tap T[q - shift(tap)][tap][c]         shift-add inside the tile
// A CTA takes a T tile of 8 rows x 32 columns of pixels (TMA box with halo, out-of-range pixels zero-filled = the padding;
// two M = 128 MMAs x 4 K steps), copies the 36 accumulator columns to shared memory (fp32), and every thread then sums the
// nine shifted 16-byte (tap) entries of one of the 6 x 30 interior output pixels and writes fp32 NCHW directly, divided by
// the ScalingLayer's scale.  The next tile's TMA load is issued as soon as the MMAs have consumed the current one.
constexpr int kDgThreads = 256;
constexpr int kDgTR = 8, kDgTC = 32;          // T tile (pixels)
constexpr int kDgOR = kDgTR - 2, kDgOC = kDgTC - 2;   // outputs per tile
constexpr int kDgPitch = 36;                  // floats per T pixel: 9 taps x 4
constexpr int kDgBRows = 40;                  // B rows kept in shared memory (N = 48: rows 40..47 alias sT, their columns are unused)
struct FirstDgradParams {
  const float* weight;   // [64][3][3][3]
  const float* scale;    // [3] or null: dx /= scale[c]
  float* dx;             // fp32 NCHW [n][3][h][w]
  int n, h, w, tiles_x, tiles_y, total_tiles;
  FastDiv fd_tx, fd_ty;
};

__global__ void __launch_bounds__(kDgThreads, 3)
vgg_first_dgrad_kernel(const FirstDgradParams p, const __grid_constant__ CUtensorMap map_dy) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                    // [256 pixels x 128 B] dy tile, 128B swizzle (TMA)
  uint8_t* sB = sA + kDgTR * kDgTC * 128;                // [40 (48) x 128 B] W2[n = tap * 4 + c][co], 128B swizzle
  float* sT = reinterpret_cast<float*>(sB + kDgBRows * 128);   // [256 pixels][36]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sT + kDgTR * kDgTC * kDgPitch);
  uint64_t* mma_bar = full_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mma_bar + 1);
  float* sScale = reinterpret_cast<float*>(tmem_ptr + 2);      // [3]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    if (lane == 0) tma_prefetch_desc(&map_dy);
    __syncwarp();
    tmem_alloc(tmem_ptr, 128);
    tmem_relinquish();
  } else if (tid == 32) {
    mbar_init(full_bar, 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  pdl_trigger();   // programmatic dependent launch (common.cuh): global memory only after pdl_wait()
  pdl_wait();
  for (int i = tid; i < kDgBRows * 64; i += kDgThreads) {
    const int n = i >> 6, k = i & 63, tap = n >> 2, c = n & 3;
    const float v = (tap < 9 && c < 3) ? p.weight[(k * 3 + c) * 9 + tap] : 0.f;
    const uint32_t off = (uint32_t)n * 128 + ((((uint32_t)k >> 3) ^ ((uint32_t)n & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(v);
  }
  if (tid < 3) sScale[tid] = p.scale != nullptr ? p.scale[tid] : 1.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t idesc = make_idesc_bf16(128, 48, 0, 0);
  const uint64_t a_desc = make_smem_desc(smem_u32(sA), 128, 16);
  const uint64_t b_desc = make_smem_desc(smem_u32(sB), 128, 16);
  auto decode = [&](int tile, int& n, int& y0, int& x0) {
    int rest, tx, ty;
    p.fd_tx.divmod(tile, rest, tx);
    p.fd_ty.divmod(rest, n, ty);
    y0 = ty * kDgOR;
    x0 = tx * kDgOC;
  };
  auto issue = [&](int tile) {   // one thread: the T tile starts one pixel above / left of its first output
    int n, y0, x0;
    decode(tile, n, y0, x0);
    mbar_expect_tx(full_bar, kDgTR * kDgTC * 128);
    tma_load_5d(sA, &map_dy, full_bar, 0, x0 - 1, y0 - 1, n, 0);
  };
  if (tid == 0 && (int)blockIdx.x < p.total_tiles) issue(blockIdx.x);
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    int n, y0, x0;
    decode(tile, n, y0, x0);
    if (tid == 0) {
      mbar_wait(full_bar, phase);
      tc_fence_after();
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          umma_bf16(tmem + m * 64, a_desc + (uint32_t)((m * 128 * 128 + j * 32) >> 4), b_desc + (uint32_t)((j * 32) >> 4), idesc,
                    j != 0);
      umma_commit(mma_bar);
    }
    __syncwarp();
    mbar_wait(mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    // the MMAs have consumed sA: fetch the next tile while this one is post-processed
    if (tid == 0 && tile + (int)gridDim.x < p.total_tiles) issue(tile + gridDim.x);
    // ---- accumulators -> sT: warp w owns M tile w / 4, TMEM lane quarter w % 4 (lane = pixel)
    {
      const int quarter = warp & 3, m = warp >> 2;
      float* dst = sT + (m * 128 + quarter * 32 + lane) * kDgPitch;
#pragma unroll
      for (int c0 = 0; c0 < 48; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + m * 64 + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c0 + 4 * q < kDgPitch)
            *reinterpret_cast<float4*>(dst + c0 + 4 * q) = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
      }
      tc_fence_before();
    }
    __syncthreads();
    // ---- shift-add: thread -> one interior pixel, nine 16-byte reads (pixel pitch 144 B: conflict-free quarter-warps)
    if (tid < kDgOR * kDgOC) {
      const int r = tid / kDgOC, cc = tid - r * kDgOC;
      const int y = y0 + r, x = x0 + cc;
      if (y < p.h && x < p.w) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 t = *reinterpret_cast<const float4*>(sT + ((r + 2 - ky) * kDgTC + (cc + 2 - kx)) * kDgPitch + (ky * 3 + kx) * 4);
            a0 += t.x; a1 += t.y; a2 += t.z;
          }
        float* o = p.dx + (((size_t)n * 3) * p.h + y) * p.w + x;
        const size_t plane = (size_t)p.h * p.w;
        o[0] = a0 / sScale[0];
        o[plane] = a1 / sScale[1];
        o[2 * plane] = a2 / sScale[2];
      }
    }
    __syncthreads();   // sT and the accumulators are rewritten by the next iteration
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

cudaError_t launch_vgg_first_dgrad(const CUtensorMap* map_dy, int n, int h, int w, const float* weight, const float* scale,
                                   float* dx, int num_sms, cudaStream_t st) {
  FirstDgradParams p;
  p.weight = weight; p.scale = scale; p.dx = dx; p.n = n; p.h = h; p.w = w;
  p.tiles_x = (w + kDgOC - 1) / kDgOC;
  p.tiles_y = (h + kDgOR - 1) / kDgOR;
  const long long total = (long long)n * p.tiles_x * p.tiles_y;
  if (total > 0x7fffffffLL) return cudaErrorInvalidValue;
  p.total_tiles = (int)total;
  p.fd_tx = make_fastdiv(p.tiles_x);
  p.fd_ty = make_fastdiv(p.tiles_y);
  const size_t smem = (size_t)kDgTR * kDgTC * 128 + kDgBRows * 128 + (size_t)kDgTR * kDgTC * kDgPitch * sizeof(float) + 16 + 8 + 16 + 1008;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(vgg_first_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = p.total_tiles < num_sms * 3 ? p.total_tiles : num_sms * 3;
  return launch_k(vgg_first_dgrad_kernel, dim3(grid), dim3(kDgThreads), smem, st, 1, p, *map_dy);
}

// =====================================================================================================================
// Image-side layers of the VQVAE (reference models/vqvae_conv3d_latent.py:109 first Conv2d(6 -> 64, 4, stride 2, pad 1);
// :154-156 last ConvTranspose2d(64 -> 6, 4, stride 2, pad 1)) WITHOUT an explicit im2col matrix.  Two kernels share one tile
// builder: the [128 output pixels x K] im2col tile (k = ch * 16 + ky * 4 + kx, the order of the PyTorch weight; K = 16 * C
// <= 128, two 64-wide chunks of 128-byte rows, 128B swizzle) is assembled in shared memory straight from the image rows:
// a thread owns one output pixel and half of the C * 4 image rows; per row it loads the aligned pair (2x, 2x + 1) with one
// 8-byte load and takes its outer taps (2x - 1, 2x + 2) from the neighbouring lanes by shuffle (lanes 0 / 31 load them).
//   s2conv_kernel   D[pixel, 64] = tile x W^T : the first conv's forward (bias + ReLU) and -- with the transposed-conv weight,
//                   whose layout [64, C, 4, 4] is exactly a Conv2d weight -- the last layer's DATA gradient (ReLU gate of the
//                   layer input, optional addend).  The tile is the K-major A operand.
//   s2wgrad_kernel  dW^T[k, 64] = sum_pixels tile^T x Y : both layers' WEIGHT gradients (Y = dy of the first conv / the input
//                   activation of the last layer).  The same tile is the MN-major A operand (M = k), Y the MN-major B operand;
//                   the accumulator stays in TMEM over all tiles of a (persistent) CTA; pad slot k = 16 * C holds 1.0, so
//                   that row of the result is the column sum of Y = the bias gradient.
// Phases of a tile run back to back inside a CTA (the image rows of the next tile are already in flight during the epilogue
// / the MMAs); three CTAs per SM overlap each other.
// =====================================================================================================================
constexpr int kS2Threads = 256;

struct S2Params {
  const float* x;          // image-side tensor [n, ca, H, W] fp32 (first c channels used)
  int n, ca, c, H, W;      // H, W even; output grid (H/2, W/2)
  int tiles_x;             // ceil((W/2) / 128)
  int total_tiles;         // n * (H/2) * tiles_x
  FastDiv fd_tx, fd_ho;    // dividers by tiles_x and H/2
  // s2conv
  const float* weight;     // [64, c, 4, 4] fp32
  const float* bias;       // [64] or null
  const __nv_bfloat16* mask;    // [n, H/2, W/2, 64] or null: result zeroed where mask <= 0
  const __nv_bfloat16* addend;  // same layout or null
  __nv_bfloat16* out;      // [n, H/2, W/2, 64]
  int relu;                // apply ReLU to the result (first conv) -- the data-gradient form writes the raw value
  // s2wgrad
  const __nv_bfloat16* y;  // [n, H/2, W/2, 64] bf16
  float* partial;          // [gridDim.x][128][64] fp32
};

struct S2Tile { int n, oy, ox0; };
__device__ __forceinline__ S2Tile s2_decode(const S2Params& p, int tile) {
  int ty, tx;
  p.fd_tx.divmod(tile, ty, tx);
  S2Tile t;
  p.fd_ho.divmod(ty, t.n, t.oy);
  t.ox0 = tx * 128;
  return t;
}

// Image values of one output pixel for the 2 * C image rows of one half of the tile row: half h owns the filter rows
// ky = 2h, 2h + 1 of every channel, i.e. the 16-byte unit u = 2 * ch + h (k = 8u .. 8u + 7 = ch * 16 + ky * 4 + kx)
template <int C>
struct S2Rows {
  float2 mid[2 * C];   // [ch * 2 + (ky & 1)]: input columns 2x, 2x + 1
  float edge[2 * C];   // lane 0: column 2x - 1; lane 31: column 2x + 2 (other lanes get them from their neighbours)
};

template <int C>
__device__ __forceinline__ void s2_load_rows(const S2Params& p, const S2Tile& t, int px, int half, int lane, S2Rows<C>& rw) {
  const int ix = 2 * (t.ox0 + px);
  const bool okm = ix < p.W;
  const int ex = lane == 0 ? ix - 1 : ix + 2;
  const bool oke = (lane == 0 || lane == 31) && (unsigned)ex < (unsigned)p.W;
  const int iy0 = 2 * t.oy - 1 + 2 * half;                   // H even: only iy0 = -1 (half 0) and iy0 + 1 = H (half 1) fall outside
  const bool ok0 = iy0 >= 0, ok1 = iy0 + 1 < p.H;
  const bool m0 = ok0 && okm, m1 = ok1 && okm, e0 = ok0 && oke, e1 = ok1 && oke;
  const size_t plane = (size_t)p.H * p.W;
  const float* r0 = p.x + (size_t)t.n * p.ca * plane + (size_t)(ok0 ? iy0 : 0) * p.W;
  const float* r1 = p.x + (size_t)t.n * p.ca * plane + (size_t)(ok1 ? iy0 + 1 : 0) * p.W;
#pragma unroll
  for (int ch = 0; ch < C; ++ch) {
    rw.mid[2 * ch] = m0 ? __ldg(reinterpret_cast<const float2*>(r0 + ix)) : make_float2(0.f, 0.f);
    rw.mid[2 * ch + 1] = m1 ? __ldg(reinterpret_cast<const float2*>(r1 + ix)) : make_float2(0.f, 0.f);
    rw.edge[2 * ch] = e0 ? __ldg(r0 + ex) : 0.f;
    rw.edge[2 * ch + 1] = e1 ? __ldg(r1 + ex) : 0.f;
    r0 += plane;
    r1 += plane;
  }
}

template <int C>
__device__ __forceinline__ void s2_build_tile(const S2Rows<C>& rw, uint8_t* sT, int px, int half, int lane) {
  uint8_t* row = sT + px * 128;
  const int sw = (px & 7) ^ half;          // unit u = 2 ch + half sits at slot (u & 7) ^ (px & 7) of chunk u >> 3 = ch >> 2
#pragma unroll
  for (int ch = 0; ch < C; ++ch) {
    uint32_t w[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 m = rw.mid[2 * ch + h];
      float left = __shfl_up_sync(0xffffffffu, m.y, 1);
      float right = __shfl_down_sync(0xffffffffu, m.x, 1);
      if (lane == 0) left = rw.edge[2 * ch + h];
      if (lane == 31) right = rw.edge[2 * ch + h];
      w[2 * h] = pack_bf16x2(left, m.x);
      w[2 * h + 1] = pack_bf16x2(m.y, right);
    }
    *reinterpret_cast<uint4*>(row + (ch >> 2) * (128 * 128) + ((((2 * ch) & 7) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// Units 2C .. 15 of every tile row are constant: zero, and 1.0 at k = 16 * C when ONES (the bias-gradient row of s2wgrad)
template <int C, bool ONES>
__device__ __forceinline__ void s2_fill_pad(uint8_t* sT, int tid) {
  for (int i = tid; i < 128 * (16 - 2 * C); i += kS2Threads) {
    const int px = i / (16 - 2 * C), u = 2 * C + i % (16 - 2 * C);
    const uint4 v = make_uint4((ONES && u == 2 * C) ? 0x00003F80u : 0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sT + (u >> 3) * (128 * 128) + px * 128 + (((u & 7) ^ (px & 7)) << 4)) = v;
  }
}

// EPI: the epilogue has a mask and / or an addend.  Their tiles are then fetched with coalesced 16-byte loads while the MMAs
// run, the accumulator (+ bias) is staged in fp32, and the gate / sum / single bf16 rounding happen in the store phase.
template <int C, bool EPI, int OCC>
__global__ void __launch_bounds__(kS2Threads, OCC)
s2conv_kernel(const S2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sT = smem;                               // 2 x [128 x 128 B] tile
  // output staging: bf16 [128 x 128 B] over the chunk the MMAs of this C do not need to survive (C = 6: chunk 0, rewritten
  // by every tile; C = 3: chunk 1, never read), or fp32 [128 x 256 B] over both (EPI; the constant pad units are refilled)
  uint8_t* sOut = EPI ? sT : sT + (C == 3 ? 128 * 128 : 0);
  uint8_t* sB = sT + 2 * 128 * 128;                 // 2 x [64 x 128 B] weights (k chunks), 128B swizzle
  float* sBias = reinterpret_cast<float*>(sB + 2 * 64 * 128);   // [64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBias + 64);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp & 1, px = (warp >> 1) * 32 + lane;
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 64);
    tmem_relinquish();
  } else if (tid == 32) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  pdl_trigger();   // programmatic dependent launch (common.cuh): global memory only after pdl_wait()
  s2_fill_pad<C, false>(sT, tid);
  pdl_wait();
  for (int i = tid; i < 64 * 128; i += kS2Threads) {
    const int co = i >> 7, k = i & 127;
    const float v = k < 16 * C ? p.weight[co * 16 * C + k] : 0.f;
    const uint32_t off = (uint32_t)(k >> 6) * (64 * 128) + co * 128 + (((((uint32_t)k & 63) >> 3) ^ ((uint32_t)co & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(v);
  }
  if (tid < 64) sBias[tid] = p.bias != nullptr ? p.bias[tid] : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  const uint64_t a_desc = make_smem_desc(smem_u32(sT), 128, 16);
  const uint64_t b_desc = make_smem_desc(smem_u32(sB), 128, 16);
  constexpr int kK16 = (16 * C + 15) / 16;          // 16-wide K slices that hold data (6 for C = 6)
  uint32_t phase = 0;
  const int Wo = p.W / 2, Ho = p.H / 2;

  S2Rows<C> rw;
  S2Tile t_next = s2_decode(p, blockIdx.x);
  if ((int)blockIdx.x < p.total_tiles) s2_load_rows<C>(p, t_next, px, half, lane, rw);
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const S2Tile t = t_next;
    const size_t pix0 = ((size_t)t.n * Ho + t.oy) * Wo + t.ox0;
    const int live_rows = Wo - t.ox0 < 128 ? Wo - t.ox0 : 128;
    s2_build_tile<C>(rw, sT, px, half, lane);
    if (EPI && C == 6) s2_fill_pad<C, false>(sT, tid);   // the fp32 staging tile of the previous iteration covered them
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < kK16; ++j) {
        const uint32_t off16 = (uint32_t)((j >> 2) * (128 * 128) + (j & 3) * 32) >> 4;
        const uint32_t boff16 = (uint32_t)((j >> 2) * (64 * 128) + (j & 3) * 32) >> 4;
        umma_bf16(tmem, a_desc + off16, b_desc + boff16, idesc, j != 0);
      }
      umma_commit(bar);
    }
    __syncwarp();
    // coalesced epilogue operands: thread -> (row, 16-byte chunk) = i >> 3, i & 7 for i = tid + 256 * j
    uint4 em[4], ea[4];
    if (EPI) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = tid + j * kS2Threads;
        const bool live = (i >> 3) < live_rows;
        em[j] = ea[j] = make_uint4(0u, 0u, 0u, 0u);
        if (p.mask != nullptr && live) em[j] = __ldg(reinterpret_cast<const uint4*>(p.mask + pix0 * 64) + i);
        if (p.addend != nullptr && live) ea[j] = __ldg(reinterpret_cast<const uint4*>(p.addend + pix0 * 64) + i);
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: warp w -> TMEM lane quarter w % 4 (lane = pixel), column half w / 4
    {
      const int quarter = warp & 3, hcol = warp >> 2;
      const int row = quarter * 32 + lane;
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + hcol * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b0 = *reinterpret_cast<const float4*>(sBias + hcol * 32 + q * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(sBias + hcol * 32 + q * 8 + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]) + bb[e];
        const int chunk = hcol * 4 + q;       // 8 channels
        if (EPI) {                            // fp32 rows of 256 B = 16 units; unit ^ (row & 7) keeps quarter-warps conflict-free
          uint8_t* r = sOut + row * 256;
          *reinterpret_cast<float4*>(r + (((2 * chunk) ^ (row & 7)) << 4)) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(r + (((2 * chunk + 1) ^ (row & 7)) << 4)) = make_float4(f[4], f[5], f[6], f[7]);
        } else {
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          uint4 o;
          o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(sOut + row * 128 + ((chunk ^ (row & 7)) << 4)) = o;
        }
      }
      tc_fence_before();
    }
    __syncthreads();
    if (tile + (int)gridDim.x < p.total_tiles) {      // in flight during the store phase
      t_next = s2_decode(p, tile + gridDim.x);
      s2_load_rows<C>(p, t_next, px, half, lane, rw);
    }
    {
      uint4* dst = reinterpret_cast<uint4*>(p.out + pix0 * 64);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = tid + j * kS2Threads;
        const int row = i >> 3, chunk = i & 7;
        if (row >= live_rows) continue;
        if (EPI) {
          const uint8_t* r = sOut + row * 256;
          const float4 lo = *reinterpret_cast<const float4*>(r + (((2 * chunk) ^ (row & 7)) << 4));
          const float4 hi = *reinterpret_cast<const float4*>(r + (((2 * chunk + 1) ^ (row & 7)) << 4));
          float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          if (p.mask != nullptr) {
            const uint32_t w[4] = {em[j].x, em[j].y, em[j].z, em[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (!(bf16lo(w[e]) > 0.f)) f[2 * e] = 0.f;
              if (!(bf16hi(w[e]) > 0.f)) f[2 * e + 1] = 0.f;
            }
          }
          {
            const uint32_t w[4] = {ea[j].x, ea[j].y, ea[j].z, ea[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { f[2 * e] += bf16lo(w[e]); f[2 * e + 1] += bf16hi(w[e]); }
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          uint4 o;
          o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
          dst[i] = o;
        } else {
          dst[i] = *reinterpret_cast<const uint4*>(sOut + row * 128 + ((chunk ^ (row & 7)) << 4));
        }
      }
    }
    __syncthreads();          // the staging tile overlays sT, which the next iteration rewrites
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

template <int C>
__global__ void __launch_bounds__(kS2Threads, 3)
s2wgrad_kernel(const S2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sT = smem;                               // 2 x [128 pixels x 128 B]: MN-major A operand, M = k
  uint8_t* sY = sT + 2 * 128 * 128;                 // [128 pixels x 128 B]: MN-major B operand, N = 64
  uint64_t* bar = reinterpret_cast<uint64_t*>(sY + 128 * 128);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp & 1, px = (warp >> 1) * 32 + lane;
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 64);
    tmem_relinquish();
  } else if (tid == 32) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  pdl_trigger();   // programmatic dependent launch (common.cuh): global memory only after pdl_wait()
  s2_fill_pad<C, true>(sT, tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
  const uint64_t a_desc = make_smem_desc(smem_u32(sT), 128, 128 * 128);   // LBO = distance between the two 64-wide M chunks
  const uint64_t b_desc = make_smem_desc(smem_u32(sY), 128, 0);
  uint32_t phase = 0;
  const int Wo = p.W / 2, Ho = p.H / 2;
  // contiguous range of tiles per CTA
  const int per = (p.total_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per;
  const int t_end = t_begin + per < p.total_tiles ? t_begin + per : p.total_tiles;
  bool first = true;
  for (int tile = t_begin; tile < t_end; ++tile) {
    const S2Tile t = s2_decode(p, tile);
    S2Rows<C> rw;
    s2_load_rows<C>(p, t, px, half, lane, rw);
    uint4 yv[4];    // Y tile: 128 pixels x 128 B, coalesced 16-byte loads; pixels past the row end are zero
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.y + (((size_t)t.n * Ho + t.oy) * Wo + t.ox0) * 64);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = tid + j * kS2Threads;
        yv[j] = make_uint4(0u, 0u, 0u, 0u);
        if (t.ox0 + (i >> 3) < Wo) yv[j] = __ldg(src + i);
      }
    }
    if (!first) {                       // the previous tile's MMAs still read sT / sY while this tile's loads are in flight
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
    }
    s2_build_tile<C>(rw, sT, px, half, lane);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = tid + j * kS2Threads;
      const int row = i >> 3, chunk = i & 7;
      *reinterpret_cast<uint4*>(sY + row * 128 + ((chunk ^ (row & 7)) << 4)) = yv[j];
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)    // 16 pixels per MMA = 16 rows of 128 B
        umma_bf16(tmem, a_desc + (uint32_t)((kk * 16 * 128) >> 4), b_desc + (uint32_t)((kk * 16 * 128) >> 4), idesc, !(first && kk == 0));
      umma_commit(bar);
    }
    __syncwarp();
    first = false;
  }
  if (!first) {
    mbar_wait(bar, phase);
    tc_fence_after();
  }
  // partial[cta][k][n]: warps -> lane quarters (k rows) w % 4, column halves w / 4
  {
    const int quarter = warp & 3, hcol = warp >> 2;
    const int row = quarter * 32 + lane;
    uint32_t v[32];
    if (!first) {
      tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + hcol * 32, v);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0u;
    }
    float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.x * 128 + row) * 64 + hcol * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

// One block per k: dweight[n64][k] (+)= sum_cta partial[cta][k][n64] for k < 16 * C;  dbias[n64] (+)= the same sum at k = 16 * C
__global__ void __launch_bounds__(1024)
s2wgrad_finalize_kernel(const float* __restrict__ partial, int ctas, int c, float* __restrict__ dweight, int accumulate,
                        float* __restrict__ dbias, int dbias_accumulate) {
  __shared__ float red[16][64];
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int k = blockIdx.x, n = threadIdx.x & 63, part = threadIdx.x >> 6;
  float s = 0.f;
  for (int j = part; j < ctas; j += 16) s += partial[((size_t)j * 128 + k) * 64 + n];
  red[part][n] = s;
  __syncthreads();
  if (part != 0) return;
#pragma unroll
  for (int j = 1; j < 16; ++j) s += red[j][n];
  if (k < 16 * c) {
    float* d = dweight + (size_t)n * 16 * c + k;
    *d = accumulate ? *d + s : s;
  } else if (dbias != nullptr) {
    dbias[n] = dbias_accumulate ? dbias[n] + s : s;
  }
}

static size_t s2_smem_bytes(bool wgrad) {
  return (wgrad ? 3 * 128 * 128 : 2 * 128 * 128 + 2 * 64 * 128 + 64 * sizeof(float)) + 64 + 1024;
}
static int s2_occ() {   // resident CTAs per SM the conv kernel is compiled for (experiment knob FO_S2_OCC = 2 | 3)
  static int occ = 0;
  if (occ == 0) {
    const char* e = getenv("FO_S2_OCC");
    occ = (e != nullptr && atoi(e) == 3) ? 3 : 2;
  }
  return occ;
}
int s2_grid(int num_sms) { return num_sms * 3; }

static S2Params s2_params(const float* x, int n, int ca, int c, int H, int W) {
  S2Params p = {};
  p.x = x; p.n = n; p.ca = ca; p.c = c; p.H = H; p.W = W;
  p.tiles_x = (W / 2 + 127) / 128;
  p.total_tiles = n * (H / 2) * p.tiles_x;
  p.fd_tx = make_fastdiv(p.tiles_x);
  p.fd_ho = make_fastdiv(H / 2);
  return p;
}

template <typename K>
static cudaError_t s2_launch(K kernel, const S2Params& p, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return launch_k(kernel, dim3(grid), dim3(kS2Threads), smem, st, 1, p);
}

cudaError_t launch_s2conv(const float* x, int n, int ca, int c, int H, int W, const float* weight, const float* bias,
                          const void* mask, const void* addend, void* out, int relu, int num_sms, cudaStream_t st) {
  if ((long long)n * (H / 2) * ((W / 2 + 127) / 128) > 0x7fffffffLL) return cudaErrorInvalidValue;
  S2Params p = s2_params(x, n, ca, c, H, W);
  p.weight = weight; p.bias = bias; p.mask = (const __nv_bfloat16*)mask; p.addend = (const __nv_bfloat16*)addend;
  p.out = (__nv_bfloat16*)out; p.relu = relu;
  const size_t smem = s2_smem_bytes(false);
  const int occ = s2_occ();
  const int grid = p.total_tiles < num_sms * occ ? p.total_tiles : num_sms * occ;
  const bool epi = mask != nullptr || addend != nullptr;
#define S2CONV(C_, E_) (occ == 3 ? s2_launch(s2conv_kernel<C_, E_, 3>, p, grid, smem, st) : s2_launch(s2conv_kernel<C_, E_, 2>, p, grid, smem, st))
  if (c == 6) return epi ? S2CONV(6, true) : S2CONV(6, false);
  if (c == 3) return epi ? S2CONV(3, true) : S2CONV(3, false);
#undef S2CONV
  return cudaErrorInvalidValue;
}

cudaError_t launch_s2wgrad(const float* x, int n, int ca, int c, int H, int W, const void* y, float* dweight, int accumulate,
                           float* dbias, int dbias_accumulate, float* workspace, int num_sms, cudaStream_t st) {
  if ((long long)n * (H / 2) * ((W / 2 + 127) / 128) > 0x7fffffffLL) return cudaErrorInvalidValue;
  S2Params p = s2_params(x, n, ca, c, H, W);
  p.y = (const __nv_bfloat16*)y; p.partial = workspace;
  const size_t smem = s2_smem_bytes(true);
  const int grid = p.total_tiles < s2_grid(num_sms) ? p.total_tiles : s2_grid(num_sms);
  cudaError_t e;
  if (c == 6) e = s2_launch(s2wgrad_kernel<6>, p, grid, smem, st);
  else if (c == 3) e = s2_launch(s2wgrad_kernel<3>, p, grid, smem, st);
  else return cudaErrorInvalidValue;
  if (e != cudaSuccess) return e;
  return launch_k(s2wgrad_finalize_kernel, dim3(16 * c + 1), dim3(1024), 0, st, 1, workspace, grid, c, dweight, accumulate, dbias,
                  dbias_accumulate);
}

cudaError_t launch_vgg_first_conv(const float* x, int n, int h, int w, const float* weight, const float* bias,
                                  const float* shift, const float* scale, void* out, int num_sms, cudaStream_t st) {
  FirstConvParams p;
  p.x = x; p.weight = weight; p.bias = bias; p.shift = shift; p.scale = scale; p.out = (__nv_bfloat16*)out;
  p.n = n; p.h = h; p.w = w;
  p.tiles_x = (w + kFcTile - 1) / kFcTile;
  if ((long long)n * h * p.tiles_x > 0x7fffffffLL || 3LL * h * w > 0x7fffffffLL) return cudaErrorInvalidValue;
  p.total_tiles = n * h * p.tiles_x;
  p.fd_tx = make_fastdiv(p.tiles_x);
  p.fd_h = make_fastdiv(h);
  const size_t smem = kFcTile * 128 + 64 * 64 + 64 * sizeof(float) + 64 + 1024;
  static_assert(2 * 128 * 64 + 9 * kFcStagePitch * sizeof(float) <= kFcTile * 128, "A tiles + staged rows fit under sOut");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(vgg_first_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  long long grid = p.total_tiles;
  const long long cap = (long long)num_sms * 3;
  if (grid > cap) grid = cap;
  return launch_k(vgg_first_conv_kernel, dim3((int)grid), dim3(kFcThreads), smem, st, 1, p);
}

}  // namespace fo
