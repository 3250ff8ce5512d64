// Split ("hi|lo") activations: helpers that move fp32 values in and out of the hi|lo bf16 pair layout
// (hi = bf16(v) in channels [0, cp), lo = bf16(v - hi) in [cp, 2cp); see fo_conv_t.split_out in include/faceoff_b200.h),
// and fp32 channels-last max-pool kernels for the LPIPS trunk in that mode.  Test infrastructure for tight parity checks
// of the tensor-core kernels against an fp64 CPU restatement; the split kernel also feeds the split-bf16 GEMMs of the
// discriminator path (faceoff_b200/mocoganhd/layers.py).  Never used by the bf16 VQVAE / LPIPS path.
#include "common.cuh"
#include "kernels.h"

namespace fo {

// element (n, ch, p) of the fp32 tensor lives at x[n * sn + ch * sc + p * sp]
__global__ void split_f32_kernel(const float* __restrict__ x, int n, int c, int hw, long long sn, long long sc,
                                 long long sp, __nv_bfloat16* __restrict__ out, int cp) {
  const size_t total = (size_t)n * hw * cp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cp);
    const size_t r = i / cp;
    const size_t nn = r / hw, p = r % hw;
    const float v = ch < c ? x[nn * sn + ch * sc + p * sp] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16(v);
    out[r * 2 * cp + ch] = hi;
    out[r * 2 * cp + cp + ch] = __float2bfloat16(v - __bfloat162float(hi));
  }
}
__global__ void merge_f32_kernel(const __nv_bfloat16* __restrict__ in, int n, int c, int hw, int cp,
                                 float* __restrict__ out, long long sn, long long sc, long long sp) {
  const size_t total = (size_t)n * hw * c;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const size_t r = i / c;
    const size_t nn = r / hw, p = r % hw;
    out[nn * sn + ch * sc + p * sp] = __bfloat162float(in[r * 2 * cp + ch]) + __bfloat162float(in[r * 2 * cp + cp + ch]);
  }
}

// fp32 channels-last 2x2/2 max pool; backward routes to the first maximum of the window (torch's rule) and applies the
// ReLU gate of the post-ReLU input (x > 0), like the bf16 kernel.
__global__ void maxpool2_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int h, int w, int c) {
  const int ho = h / 2, wo = w / 2;
  const size_t total = (size_t)n * ho * wo * c;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    size_t r = i / c;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const size_t nn = r / ho;
    float m = -INFINITY;
    for (int k = 0; k < 4; ++k)
      m = fmaxf(m, x[((nn * h + 2 * oy + (k >> 1)) * w + 2 * ox + (k & 1)) * c + ch]);
    y[i] = m;
  }
}
__global__ void maxpool2_bwd_f32_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                        const float* __restrict__ dy, float* __restrict__ dx, int n, int h, int w, int c) {
  const int ho = h / 2, wo = w / 2;
  const size_t total = (size_t)n * ho * wo * c;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    size_t r = i / c;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const size_t nn = r / ho;
    (void)y;   // the maximum is recomputed: y went through a split/merge round trip in the caller
    float m = -INFINITY;
    for (int k = 0; k < 4; ++k)
      m = fmaxf(m, x[((nn * h + 2 * oy + (k >> 1)) * w + 2 * ox + (k & 1)) * c + ch]);
    const float g = dy[i];
    bool taken = false;
    for (int k = 0; k < 4; ++k) {
      const size_t idx = ((nn * h + 2 * oy + (k >> 1)) * w + 2 * ox + (k & 1)) * c + ch;
      const float v = x[idx];
      const bool hit = !taken && v == m;
      taken = taken || hit;
      dx[idx] = (hit && v > 0.f) ? g : 0.f;
    }
  }
}

static int blocks_for(size_t total) {
  size_t b = (total + 255) / 256;
  return (int)(b < 1 ? 1 : b > 148 * 16 ? 148 * 16 : b);
}
cudaError_t launch_split_f32(const float* x, int n, int c, int hw, long long sn, long long sc, long long sp, void* out,
                             int cp, cudaStream_t st) {
  split_f32_kernel<<<blocks_for((size_t)n * hw * cp), 256, 0, st>>>(x, n, c, hw, sn, sc, sp, (__nv_bfloat16*)out, cp);
  return cudaGetLastError();
}
cudaError_t launch_merge_f32(const void* in, int n, int c, int hw, int cp, float* out, long long sn, long long sc,
                             long long sp, cudaStream_t st) {
  merge_f32_kernel<<<blocks_for((size_t)n * hw * c), 256, 0, st>>>((const __nv_bfloat16*)in, n, c, hw, cp, out, sn, sc, sp);
  return cudaGetLastError();
}
cudaError_t launch_maxpool2_f32(const float* x, float* y, int n, int h, int w, int c, cudaStream_t st) {
  maxpool2_f32_kernel<<<blocks_for((size_t)n * (h / 2) * (w / 2) * c), 256, 0, st>>>(x, y, n, h, w, c);
  return cudaGetLastError();
}
cudaError_t launch_maxpool2_bwd_f32(const float* x, const float* y, const float* dy, float* dx, int n, int h, int w, int c,
                                    cudaStream_t st) {
  maxpool2_bwd_f32_kernel<<<blocks_for((size_t)n * (h / 2) * (w / 2) * c), 256, 0, st>>>(x, y, dy, dx, n, h, w, c);
  return cudaGetLastError();
}

}  // namespace fo
