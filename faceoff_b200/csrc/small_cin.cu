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
  long long total_tiles; // n * h * tiles_x
};

__global__ void __launch_bounds__(kFcThreads, 4)
vgg_first_conv_kernel(const FirstConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // the output staging tile (32 KB) overlays the A tiles and the staged image rows, which are dead by the time the
  // epilogue runs (one extra barrier per tile buys a fourth CTA per SM)
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
  // lines (threads 0 / 1 also the two right-most columns), no per-element index arithmetic.
  float pre[9], pre_tail[9];
  auto fetch = [&](long long tile) {
    const int tx = (int)(tile % p.tiles_x);
    const long long ty = tile / p.tiles_x;
    const int y = (int)(ty % p.h);
    const int n = (int)(ty / p.h);
    const int x0 = tx * kFcTile;
    const int ix = x0 - 1 + tid, ixt = x0 + 255 + tid;       // tail: columns 256, 257 of the patch (tid < 2)
    const bool okx = (unsigned)ix < (unsigned)p.w, okt = tid < 2 && (unsigned)ixt < (unsigned)p.w;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int ch = r / 3, ky = r - ch * 3;
      const int iy = y - 1 + ky;
      const bool oky = (unsigned)iy < (unsigned)p.h;
      const float* row = p.x + (((size_t)n * 3 + ch) * p.h + (oky ? iy : 0)) * p.w;
      pre[r] = (oky && okx) ? fmaf(__ldg(row + ix), sc_a[ch], sc_b[ch]) : 0.f;
      pre_tail[r] = (oky && okt) ? fmaf(__ldg(row + ixt), sc_a[ch], sc_b[ch]) : 0.f;
    }
  };
  if ((long long)blockIdx.x < p.total_tiles) fetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int tx = (int)(tile % p.tiles_x);
    const long long ty = tile / p.tiles_x;
    const int y = (int)(ty % p.h);
    const int n = (int)(ty / p.h);
    const int x0 = tx * kFcTile;
    // ---- the 3 x 3 image rows of the tile (258 columns each) -> shared memory; then start on the next tile's values
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      sStage[r * kFcStagePitch + tid] = pre[r];
      if (tid < 2) sStage[r * kFcStagePitch + 256 + tid] = pre_tail[r];
    }
    __syncthreads();
    if (tile + gridDim.x < p.total_tiles) fetch(tile + gridDim.x);
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
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[q * 8 + e]) + bb[e], 0.f);
          uint4 o;
          o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
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

cudaError_t launch_vgg_first_conv(const float* x, int n, int h, int w, const float* weight, const float* bias,
                                  const float* shift, const float* scale, void* out, int num_sms, cudaStream_t st) {
  FirstConvParams p;
  p.x = x; p.weight = weight; p.bias = bias; p.shift = shift; p.scale = scale; p.out = (__nv_bfloat16*)out;
  p.n = n; p.h = h; p.w = w;
  p.tiles_x = (w + kFcTile - 1) / kFcTile;
  p.total_tiles = (long long)n * h * p.tiles_x;
  const size_t smem = kFcTile * 128 + 64 * 64 + 64 * sizeof(float) + 64 + 1024;
  static_assert(2 * 128 * 64 + 9 * kFcStagePitch * sizeof(float) <= kFcTile * 128, "A tiles + staged rows fit under sOut");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(vgg_first_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  long long grid = p.total_tiles;
  const long long cap = (long long)num_sms * 4;
  if (grid > cap) grid = cap;
  vgg_first_conv_kernel<<<(int)grid, kFcThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace fo
