// Bandwidth-bound helpers around the implicit-GEMM kernels: layout packing, weight re-packing,
// split-K reduction, bias-gradient column sums, ReLU, 2x2 max-pool.  All vectorised (16-byte accesses on the
// channels-last side) and sized as multiples of the SM count.
#include "common.cuh"
#include "kernels.h"

namespace fo {

// ---------------------------------------------------------------- NCHW fp32 -> channels-last bf16
// block handles 32 pixels of one image; smem transpose so both sides are coalesced.
__global__ void pack_nchw_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int c, int hw, int cs,
                                 const float* __restrict__ shift, const float* __restrict__ scale) {
  __shared__ float tile[32][33];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int cc = ty; cc < cs && cc < 32; cc += 8) {
    float v = 0.f;
    if (cc < c && p0 + tx < hw) {
      v = x[((size_t)n * c + cc) * hw + p0 + tx];
      if (shift != nullptr) v = (v - shift[cc]) / scale[cc];
    }
    tile[cc][tx] = v;
  }
  __syncthreads();
  // write: each pixel has cs (<=32) channels contiguous
  for (int i = threadIdx.x; i < 32 * cs; i += blockDim.x) {
    const int px = i / cs, cc = i % cs;
    if (p0 + px < hw) out[((size_t)n * hw + p0 + px) * cs + cc] = __float2bfloat16(tile[cc][px]);
  }
}

// generic (any c): one thread per (pixel, 8-channel group) reading strided planes; used for c > 32
__global__ void pack_nchw_wide_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int c, int hw,
                                      int cs, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % cs);
    const size_t np = i / cs;
    const size_t n = np / hw, px = np % hw;
    float v = cc < c ? x[(n * c + cc) * hw + px] : 0.f;
    out[i] = __float2bfloat16(v);
  }
}

__global__ void unpack_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int c, int hw,
                                   int cs) {
  // block: 32 pixels x up to 32 channels per pass through smem
  __shared__ float tile[32][33];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  for (int cb = 0; cb < c; cb += 32) {
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
      const int px = i >> 5, cc = i & 31;
      float v = 0.f;
      if (cb + cc < c && p0 + px < hw) v = __bfloat162float(x[((size_t)n * hw + p0 + px) * cs + cb + cc]);
      tile[cc][px] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
      const int cc = i >> 5, px = i & 31;
      if (cb + cc < c && p0 + px < hw) out[((size_t)n * c + cb + cc) * hw + p0 + px] = tile[cc][px];
    }
    __syncthreads();
  }
}

__global__ void relu_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, size_t nvec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = x[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // bf16 pair: clear negative halves
      uint32_t u = w[e];
      if (u & 0x8000u) u &= 0xFFFF0000u;
      if (u & 0x80000000u) u &= 0x0000FFFFu;
      w[e] = u;
    }
    y[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---------------------------------------------------------------- weight re-pack
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, PackParams pp) {
  const size_t total = (size_t)pp.npad * pp.ktot;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / pp.ktot);
    const int col = (int)(i % pp.ktot);
    const int step = col / pp.kc, j = col % pp.kc;
    const PackStep s = pp.steps[step];
    float v = 0.f;
    if (n < pp.cout && j < s.valid) {
      const size_t k = (size_t)s.wk0 + j;
      const size_t idx = pp.n_axis == 0 ? ((size_t)n * pp.dimB + k) * pp.taps + s.tap
                                        : (k * pp.dimB + n) * pp.taps + s.tap;
      v = w[idx];
      if (pp.n_scale != nullptr) v *= pp.n_scale[n];
    }
    out[i] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------- split-K reduce + scatter to PyTorch layout
__global__ void wgrad_finalize_kernel(FinalizeParams fp) {
  const size_t per_tap = (size_t)fp.MC * fp.NC;
  const size_t total = (size_t)fp.taps * fp.m_real * fp.n_real;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // n fastest for coalesced partial reads
    const int n = (int)(i % fp.n_real);
    const int m = (int)((i / fp.n_real) % fp.m_real);
    const int tap = (int)(i / ((size_t)fp.n_real * fp.m_real));
    const float* src = fp.partial + (size_t)tap * per_tap + (size_t)m * fp.NC + n;
    float acc = 0.f;
    for (int s = 0; s < fp.splits; ++s) acc += src[(size_t)s * fp.taps * per_tap];
    const size_t idx = fp.m_axis == 0 ? ((size_t)m * fp.dimB + (fp.q_w_off + n)) * fp.taps + tap
                                      : ((size_t)(fp.q_w_off + n) * fp.dimB + m) * fp.taps + tap;
    if (fp.accumulate) fp.dweight[idx] += acc; else fp.dweight[idx] = acc;
  }
}

// ---------------------------------------------------------------- column sums (bias gradients)
// x bf16 [rows, cs]; each block strides over rows; thread owns an 8-channel vector lane.
__global__ void colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, size_t rows, int cs, float* __restrict__ part) {
  extern __shared__ float sm[];  // [rows_per_iter][cs]
  const int vecs = cs / 8;
  const int rpi = blockDim.x / vecs;  // rows per iteration
  const int lane_v = threadIdx.x % vecs;
  const int r_in = threadIdx.x / vecs;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r_in < rpi) {
    for (size_t r = (size_t)blockIdx.x * rpi + r_in; r < rows; r += (size_t)gridDim.x * rpi) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + r * cs) + lane_v);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] += bf16lo(w[e]);
        acc[2 * e + 1] += bf16hi(w[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sm[r_in * cs + lane_v * 8 + e] = acc[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cs; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rpi; ++r) s += sm[r * cs + c];
    part[(size_t)blockIdx.x * cs + c] = s;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int nblocks, int cs, int c_off, int c,
                                    float* __restrict__ out, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += part[(size_t)b * cs + c_off + i];
  if (accumulate) out[i] += s; else out[i] = s;
}

// ---------------------------------------------------------------- 2x2 max pool, channels-last bf16
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__global__ void maxpool2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int vecs) {
  const int ho = h / 2, wo = w / 2;
  const size_t total = (size_t)n * ho * wo * vecs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    size_t r = i / vecs;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const size_t nn = r / ho;
    const size_t base = ((nn * h + 2 * oy) * w + 2 * ox) * vecs + v;
    const uint4 a = __ldg(x + base), b = __ldg(x + base + vecs), c = __ldg(x + base + (size_t)w * vecs),
                d = __ldg(x + base + (size_t)w * vecs + vecs);
    uint4 o;
    o.x = bf16x2_max(bf16x2_max(a.x, b.x), bf16x2_max(c.x, d.x));
    o.y = bf16x2_max(bf16x2_max(a.y, b.y), bf16x2_max(c.y, d.y));
    o.z = bf16x2_max(bf16x2_max(a.z, b.z), bf16x2_max(c.z, d.z));
    o.w = bf16x2_max(bf16x2_max(a.w, b.w), bf16x2_max(c.w, d.w));
    y[i] = o;
  }
}
// dx[pos] = (x[pos] == y && first such position in (row-major) window order) ? dy : 0, then optional ReLU mask of x
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                                    const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int n, int h,
                                    int w, int cs) {
  const int ho = h / 2, wo = w / 2;
  const size_t total = (size_t)n * ho * wo * cs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cs);
    size_t r = i / cs;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const size_t nn = r / ho;
    const float yo = __bfloat162float(y[i]);
    const __nv_bfloat16 g = dy[i];
    const __nv_bfloat16 zero = __float2bfloat16(0.f);
    bool taken = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t idx = ((nn * h + 2 * oy + (k >> 1)) * w + 2 * ox + (k & 1)) * cs + c;
      const float xv = __bfloat162float(x[idx]);
      const bool hit = !taken && (xv == yo);
      // x is a post-ReLU activation: x == 0 means the ReLU gate is closed, so no gradient flows
      dx[idx] = (hit && xv > 0.f) ? g : zero;
      taken = taken || hit;
    }
  }
}

// ---------------------------------------------------------------- launchers
static inline int grid_for(size_t work_items, int threads, int num_sms, int per_sm = 8) {
  size_t blocks = (work_items + threads - 1) / threads;
  size_t cap = (size_t)num_sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

cudaError_t launch_pack_nchw(const float* x, void* out, int n, int c, int hw, int cs, const float* shift,
                             const float* scale, int num_sms, cudaStream_t st) {
  if (cs <= 32) {
    dim3 grid((hw + 31) / 32, n);
    pack_nchw_kernel<<<grid, 256, 0, st>>>(x, (__nv_bfloat16*)out, c, hw, cs, shift, scale);
  } else {
    const size_t total = (size_t)n * hw * cs;
    pack_nchw_wide_kernel<<<grid_for(total, 256, num_sms), 256, 0, st>>>(x, (__nv_bfloat16*)out, c, hw, cs, total);
  }
  return cudaGetLastError();
}
cudaError_t launch_unpack_nchw(const void* x, float* out, int n, int c, int hw, int cs, cudaStream_t st) {
  dim3 grid((hw + 31) / 32, n);
  unpack_nchw_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, out, c, hw, cs);
  return cudaGetLastError();
}
cudaError_t launch_relu(const void* x, void* y, size_t numel, int num_sms, cudaStream_t st) {
  const size_t nvec = numel / 8;
  relu_kernel<<<grid_for(nvec, 256, num_sms), 256, 0, st>>>((const uint4*)x, (uint4*)y, nvec);
  return cudaGetLastError();
}
cudaError_t launch_pack_weights(const float* w, void* out, const PackParams& pp, int num_sms, cudaStream_t st) {
  const size_t total = (size_t)pp.npad * pp.ktot;
  pack_weights_kernel<<<grid_for(total, 256, num_sms), 256, 0, st>>>(w, (__nv_bfloat16*)out, pp);
  return cudaGetLastError();
}
cudaError_t launch_wgrad_finalize(const FinalizeParams& fp, int num_sms, cudaStream_t st) {
  const size_t total = (size_t)fp.taps * fp.m_real * fp.n_real;
  wgrad_finalize_kernel<<<grid_for(total, 256, num_sms), 256, 0, st>>>(fp);
  return cudaGetLastError();
}
int colsum_blocks(int num_sms) { return num_sms * 4; }
cudaError_t launch_colsum(const void* x, size_t rows, int cs, int c_off, int c, float* out, int accumulate,
                          float* workspace, int num_sms, cudaStream_t st) {
  const int threads = 256;
  const int vecs = cs / 8;
  const int rpi = threads / vecs;
  int nblocks = colsum_blocks(num_sms);
  const size_t need = (rows + rpi - 1) / rpi;
  if ((size_t)nblocks > need) nblocks = (int)(need ? need : 1);
  colsum_partial_kernel<<<nblocks, threads, (size_t)rpi * cs * sizeof(float), st>>>((const __nv_bfloat16*)x, rows, cs,
                                                                                     workspace);
  colsum_final_kernel<<<(c + 127) / 128, 128, 0, st>>>(workspace, nblocks, cs, c_off, c, out, accumulate);
  return cudaGetLastError();
}
cudaError_t launch_maxpool2(const void* x, void* y, int n, int h, int w, int cs, int num_sms, cudaStream_t st) {
  const int vecs = cs / 8;
  const size_t total = (size_t)n * (h / 2) * (w / 2) * vecs;
  maxpool2_kernel<<<grid_for(total, 256, num_sms), 256, 0, st>>>((const uint4*)x, (uint4*)y, n, h, w, vecs);
  return cudaGetLastError();
}
cudaError_t launch_maxpool2_bwd(const void* x, const void* y, const void* dy, void* dx, int n, int h, int w, int cs,
                                int num_sms, cudaStream_t st) {
  const size_t total = (size_t)n * (h / 2) * (w / 2) * cs;
  maxpool2_bwd_kernel<<<grid_for(total, 256, num_sms), 256, 0, st>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, n, h, w, cs);
  return cudaGetLastError();
}

}  // namespace fo
