// Bandwidth-bound helpers around the implicit-GEMM kernels: layout packing, weight re-packing,
// split-K reduction, bias-gradient column sums, ReLU, 2x2 max-pool.  All vectorised (16-byte accesses on the
// channels-last side) and sized as multiples of the SM count.
#include "common.cuh"
#include "kernels.h"

namespace fo {

// ---------------------------------------------------------------- NCHW fp32 -> channels-last bf16
// one thread per pixel: per-plane reads are coalesced across the warp, each thread writes its CS channels as 16-byte
// vectors (a warp writes 32*CS*2 contiguous bytes).
template <int CS>
__global__ void pack_nchw_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int c, int hw,
                                 const float* __restrict__ shift, const float* __restrict__ scale, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / hw, px = i % hw;
    float v[CS];
#pragma unroll
    for (int cc = 0; cc < CS; ++cc) {
      v[cc] = 0.f;
      if (cc < c) {
        v[cc] = __ldg(x + (n * c + cc) * hw + px);
        if (shift != nullptr) v[cc] = (v[cc] - shift[cc]) / scale[cc];
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + i * CS);
#pragma unroll
    for (int q = 0; q < CS / 8; ++q) {
      uint4 o;
      o.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
      o.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
      o.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
      o.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
      dst[q] = o;
    }
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

// one thread per pixel: reads its cs channels as 16-byte vectors, writes c planes (coalesced across the warp)
template <int CS>
__global__ void unpack_nchw_small_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int c, int hw,
                                         size_t total) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / hw, px = i % hw;
    float v[CS];
    const uint4* src = reinterpret_cast<const uint4*>(x + i * CS);
#pragma unroll
    for (int q = 0; q < CS / 8; ++q) {
      const uint4 u = __ldg(src + q);
      v[8 * q + 0] = bf16lo(u.x); v[8 * q + 1] = bf16hi(u.x); v[8 * q + 2] = bf16lo(u.y); v[8 * q + 3] = bf16hi(u.y);
      v[8 * q + 4] = bf16lo(u.z); v[8 * q + 5] = bf16hi(u.z); v[8 * q + 6] = bf16lo(u.w); v[8 * q + 7] = bf16hi(u.w);
    }
#pragma unroll
    for (int cc = 0; cc < CS; ++cc)
      if (cc < c) out[(n * c + cc) * hw + px] = v[cc];
  }
}

__global__ void unpack_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int c, int hw,
                                   int cs) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
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
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
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
// One launch reduces the weight partials AND (blocks past main_blocks) the bias column-sum partials of a wgrad_igemm launch.
// Up to 148 splits per output element: a thread that walks them alone is a chain of ~40 dependent L2 round trips (the former
// kernel: 18 us per launch whatever the batch, 11 us more for the bias -- 1.3 ms of every step, i.e. 12 % of a 4-clip
// step).  Here the 8 warps of a block take the splits s = warp, warp + 8, ... of the same 32 consecutive elements (lanes along
// n: coalesced), four independent accumulators each, and warp 0 adds the 8 sums in a fixed order: deterministic.
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(FinalizeParams fp, int main_blocks, const float* __restrict__ bias_part, int bias_rows,
                      int bias_c, float* __restrict__ dbias, int dbias_accumulate) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if ((int)blockIdx.x >= main_blocks) {
    // bias gradient: dbias[c] (+)= sum_rows bias_part[row][c]   (rows = passes * splits, row pitch = MC)
    const int i = ((int)blockIdx.x - main_blocks) * 32 + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (i < bias_c) {
      const float* src = bias_part + i;
      int r = w;
      for (; r + 24 < bias_rows; r += 32) {
        a0 += src[(size_t)r * fp.MC];
        a1 += src[(size_t)(r + 8) * fp.MC];
        a2 += src[(size_t)(r + 16) * fp.MC];
        a3 += src[(size_t)(r + 24) * fp.MC];
      }
      for (; r < bias_rows; r += 8) a0 += src[(size_t)r * fp.MC];
    }
    red[w][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (w == 0 && i < bias_c) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += red[k][lane];
      if (dbias_accumulate) dbias[i] += acc; else dbias[i] = acc;
    }
    return;
  }
  const size_t per_tap = (size_t)fp.MC * fp.NC;
  const size_t split_stride = (size_t)fp.taps * per_tap;
  const size_t total = (size_t)fp.taps * fp.m_real * fp.n_real;
  if (fp.splits <= 16) {
    // few splits (many passes, e.g. Conv3d: 27 taps x 128 x 128 elements over 16 splits): one thread per element
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)main_blocks * blockDim.x) {
      const int n = (int)(i % fp.n_real);
      const int m = (int)((i / fp.n_real) % fp.m_real);
      const int tap = (int)(i / ((size_t)fp.n_real * fp.m_real));
      const float* src = fp.partial + (size_t)tap * per_tap + (size_t)m * fp.NC + n;
      float acc = 0.f;
      for (int sp = 0; sp < fp.splits; ++sp) acc += src[(size_t)sp * split_stride];
      const int wt = fp.tap_index[tap];
      const size_t idx = fp.m_axis == 0 ? ((size_t)m * fp.dimB + (fp.q_w_off + n)) * fp.taps + wt
                                        : ((size_t)(fp.q_w_off + n) * fp.dimB + m) * fp.taps + wt;
      if (fp.accumulate) fp.dweight[idx] += acc; else fp.dweight[idx] = acc;
    }
    return;
  }
  for (size_t base = (size_t)blockIdx.x * 32; base < total; base += (size_t)main_blocks * 32) {
    const size_t i = base + lane;
    // n fastest for coalesced partial reads
    const int n = (int)(i % fp.n_real);
    const int m = (int)((i / fp.n_real) % fp.m_real);
    const int tap = (int)(i / ((size_t)fp.n_real * fp.m_real));
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (i < total) {
      const float* src = fp.partial + (size_t)tap * per_tap + (size_t)m * fp.NC + n;
      int sp = w;
      for (; sp + 24 < fp.splits; sp += 32) {
        a0 += src[(size_t)sp * split_stride];
        a1 += src[(size_t)(sp + 8) * split_stride];
        a2 += src[(size_t)(sp + 16) * split_stride];
        a3 += src[(size_t)(sp + 24) * split_stride];
      }
      for (; sp < fp.splits; sp += 8) a0 += src[(size_t)sp * split_stride];
    }
    red[w][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (w == 0 && i < total) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += red[k][lane];
      const int wt = fp.tap_index[tap];
      const size_t idx = fp.m_axis == 0 ? ((size_t)m * fp.dimB + (fp.q_w_off + n)) * fp.taps + wt
                                        : ((size_t)(fp.q_w_off + n) * fp.dimB + m) * fp.taps + wt;
      if (fp.accumulate) fp.dweight[idx] += acc; else fp.dweight[idx] = acc;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- column sums (bias gradients)
// x bf16 [rows, cs]; each block strides over rows; thread owns an 8-channel vector lane.
__global__ void colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, size_t rows, int cs, float* __restrict__ part) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  extern __shared__ float sm[];  // [rows_per_iter][cs]
  const int vecs = cs / 8;
  const int rpi = blockDim.x / vecs;  // rows per iteration
  const int lane_v = threadIdx.x % vecs;
  const int r_in = threadIdx.x / vecs;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r_in < rpi) {
    const size_t stride = (size_t)gridDim.x * rpi;
    size_t r = (size_t)blockIdx.x * rpi + r_in;
    // four independent 16-byte loads in flight per thread
    for (; r + 3 * stride < rows; r += 4 * stride) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u * stride) * cs) + lane_v);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += bf16lo(w[e]);
          acc[2 * e + 1] += bf16hi(w[e]);
        }
      }
    }
    for (; r < rows; r += stride) {
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
// 32 channels per CTA, 32 slices of the partial rows per channel (coalesced over channels), eight independent loads in
// flight per thread (the rows are ~1200 L2 round trips: the former 8-slice, 4-accumulator version took 30 us per launch
// whatever the batch), fixed-order sums: deterministic
__global__ void __launch_bounds__(1024)
colsum_final_kernel(const float* __restrict__ part, int nblocks, int cs, int c_off, int c, float* __restrict__ out,
                    int accumulate) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  __shared__ float red[32][33];
  const int ci = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ci;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (i < c) {
    const float* col = part + c_off + i;
    int b = sl;
    for (; b + 7 * 32 < nblocks; b += 8 * 32) {
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += col[(size_t)(b + u * 32) * cs];
    }
    for (; b < nblocks; b += 32) a[0] += col[(size_t)b * cs];
  }
  red[sl][ci] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (sl == 0 && i < c) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) s += red[k][ci];
    if (accumulate) out[i] += s; else out[i] = s;
  }
}

// ---------------------------------------------------------------- 2x2 max pool, channels-last bf16
__global__ void maxpool2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int vecs) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
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
// dx[pos] = (x[pos] == y && first such position in (row-major) window order && x[pos] > 0) ? dy : 0  (pool_bwd_pair, common.cuh)
__global__ void maxpool2_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ y, const uint4* __restrict__ dy,
                                    uint4* __restrict__ dx, int n, int h, int w, int vecs) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int ho = h / 2, wo = w / 2;
  const size_t total = (size_t)n * ho * wo * vecs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    size_t r = i / vecs;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const size_t nn = r / ho;
    const uint4 yo = __ldg(y + i), g = __ldg(dy + i);
    // packed form of pool_bwd_pair (common.cuh): m_k = (x_k == y) and no earlier window position matched; x_k > 0 is y > 0
    // there (the per-half scalar form was instruction bound: 4.6 TB/s)
    size_t idx[4];
    uint4 xv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      idx[k] = ((nn * h + 2 * oy + (k >> 1)) * w + 2 * ox + (k & 1)) * vecs + v;
      xv[k] = __ldg(x + idx[k]);
    }
    const uint32_t yw[4] = {yo.x, yo.y, yo.z, yo.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t x0 = e == 0 ? xv[0].x : e == 1 ? xv[0].y : e == 2 ? xv[0].z : xv[0].w;
      const uint32_t x1 = e == 0 ? xv[1].x : e == 1 ? xv[1].y : e == 2 ? xv[1].z : xv[1].w;
      const uint32_t x2 = e == 0 ? xv[2].x : e == 1 ? xv[2].y : e == 2 ? xv[2].z : xv[2].w;
      const uint32_t x3 = e == 0 ? xv[3].x : e == 1 ? xv[3].y : e == 2 ? xv[3].z : xv[3].w;
      const uint32_t gated = gw[e] & bf16x2_gt_mask(yw[e], 0u);
      const uint32_t m0 = bf16x2_eq_mask(x0, yw[e]);
      const uint32_t m1 = bf16x2_eq_mask(x1, yw[e]) & ~m0;
      const uint32_t m01 = m0 | m1;
      const uint32_t m2 = bf16x2_eq_mask(x2, yw[e]) & ~m01;
      const uint32_t m3 = bf16x2_eq_mask(x3, yw[e]) & ~(m01 | m2);
      o[0][e] = gated & m0; o[1][e] = gated & m1; o[2][e] = gated & m2; o[3][e] = gated & m3;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) dx[idx[k]] = make_uint4(o[k][0], o[k][1], o[k][2], o[k][3]);
  }
}

// ---------------------------------------------------------------- uint8 frames -> normalised fp32 planes
// The reference's loader turns each uint8 HWC frame into a tensor with torchvision's ToTensor (x / 255, HWC -> CHW) and
// Normalize(0.5, 0.5) ((v - 0.5) / 0.5) on the CPU (TemporalAlignment/dataset.py:235-249) and concatenates source and
// background on the channel axis (utils.py:29-38).  Here the uint8 frames cross PCIe (4x fewer bytes) and this kernel does
// the same arithmetic, in the same order and rounding (IEEE division, subtraction, division), into channels
// c_off..c_off+2 of an NCHW tensor [n, c_total, hw].  4 pixels (12 bytes) per thread, 16-byte stores per plane.
__global__ void u8hwc_to_nchw_kernel(const uint32_t* __restrict__ x, float* __restrict__ out, int hw4, size_t total4,
                                     int c_total, int c_off, float mean, float stdv) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / hw4, p4 = i % hw4;
    const uint32_t w0 = __ldg(x + 3 * i), w1 = __ldg(x + 3 * i + 1), w2 = __ldg(x + 3 * i + 2);
    // bytes: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
    const uint32_t b[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24,
                            w1 & 255u, (w1 >> 8) & 255u, (w1 >> 16) & 255u, w1 >> 24,
                            w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)b[k], 255.f), mean), stdv);
    float* o = out + (n * c_total + c_off) * (size_t)hw4 * 4 + p4 * 4;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      *reinterpret_cast<float4*>(o + (size_t)ch * hw4 * 4) = make_float4(v[ch], v[3 + ch], v[6 + ch], v[9 + ch]);
  }
}

// ---------------------------------------------------------------- im2col / col2im for the 6-channel 4x4 stride-2 layers
// The first Conv2d (6->64, k4 s2 p1) and the last ConvTranspose2d (64->6) have too few channels for an efficient
// implicit GEMM (32-byte TMA rows, N=16 MMAs).  Their image-side operand is therefore laid out as an explicit
// [pixels, 16 taps x 8 channels = 128] bf16 matrix so that the layer becomes a plain K=128 (or N=128) GEMM on the same
// tcgen05 kernel.  k = (ky*4+kx)*8 + c, value = x[c][2*oy+ky-1][2*ox+kx-1] (zero outside / for c >= C).
__global__ void __launch_bounds__(256)
im2col4x4s2_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int ca, int c, int h, int w) {
  // block: 64 output pixels of one output row; input window 4 rows x 130 columns x c channels.
  // Warp cc stages channel cc (lanes run along the row: coalesced, no per-element index arithmetic).
  __shared__ float sm[8][4][132];
  const int wo = w / 2;
  const int n = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * 64;
  const int ix0 = 2 * ox0 - 1, iy0 = 2 * oy - 1;
  const int cc = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    float v[4][5];
    const float* xc = x + ((size_t)n * ca + cc) * h * w;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int iy = iy0 + r;
      const bool row_ok = cc < c && iy >= 0 && iy < h;
      const float* xr = xc + (size_t)iy * w;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int col = lane + 32 * k, ix = ix0 + col;
        v[r][k] = (row_ok && col < 130 && ix >= 0 && ix < w) ? __ldg(xr + ix) : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int col = lane + 32 * k;
        if (col < 130) sm[cc][r][col] = v[r][k];
      }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = threadIdx.x + it * 256;
    const int px = i >> 4, tap = i & 15;
    const int ox = ox0 + px;
    if (ox >= wo) continue;
    const int ky = tap >> 2, kx = tap & 3;
    const int col = 2 * px + kx;
    uint4 o;
    o.x = pack_bf16x2(sm[0][ky][col], sm[1][ky][col]);
    o.y = pack_bf16x2(sm[2][ky][col], sm[3][ky][col]);
    o.z = pack_bf16x2(sm[4][ky][col], sm[5][ky][col]);
    o.w = pack_bf16x2(sm[6][ky][col], sm[7][ky][col]);
    reinterpret_cast<uint4*>(out + (((size_t)n * (h / 2) + oy) * wo + ox) * 128)[tap] = o;
  }
}

// 3x3 pad-1 im2col for c <= 3 input channels (first VGG16 conv of LPIPS, reference models/lpips.py:119-127) with the
// ScalingLayer folded in: out[n][y][x][(ky*3+kx)*3 + ch] = (x[ch][y+ky-1][x+kx-1] - shift[ch]) / scale[ch], zero outside
// the image and for k >= 27 (K padded to 32 -> 64-byte rows, one K=32 GEMM step instead of nine 32-byte-row TMA boxes).
// Block = 256 threads = a strip of 256 pixels x kIm2Rows image rows.  Load phase: one coalesced row load per (channel,
// input row) into shared memory (scaling applied once per loaded element).  Store phase: a thread owns ONE 16-byte piece
// (8 of the 32 k values; the 8 shared-memory offsets are computed once) and walks over pixels, so that a warp instruction
// writes 8 pixels x 64 B = 512 contiguous bytes.
constexpr int kIm2Rows = 4;
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int c, int h, int w,
                 const float* __restrict__ shift, const float* __restrict__ scale) {
  constexpr int P = 260;   // row pitch in floats (258 used)
  __shared__ float sm[3 * (kIm2Rows + 2) * P];
  const int n = blockIdx.z, y0 = blockIdx.y * kIm2Rows, x0 = blockIdx.x * 256;
  const int tid = threadIdx.x;
  for (int cr = 0; cr < 3 * (kIm2Rows + 2); ++cr) {
    const int cc = cr / (kIm2Rows + 2), r = cr % (kIm2Rows + 2);
    const int iy = y0 - 1 + r;
    const bool row_ok = cc < c && iy >= 0 && iy < h;
    const float* xr = x + (((size_t)n * c + (cc < c ? cc : 0)) * h + (row_ok ? iy : 0)) * w;
    const float sh = (shift != nullptr && cc < c) ? shift[cc] : 0.f, sc = (scale != nullptr && cc < c) ? scale[cc] : 1.f;
    auto fetch = [&](int j) {           // shared-memory column j <-> image column x0 - 1 + j
      const int ix = x0 - 1 + j;
      float v = 0.f;
      if (row_ok && ix >= 0 && ix < w) {
        v = __ldg(xr + ix);
        if (shift != nullptr) v = (v - sh) / sc;
      }
      sm[cr * P + j] = v;
    };
    fetch(tid + 1);
    if (tid < 2) fetch(tid == 0 ? 0 : 257);
  }
  __syncthreads();
  const int piece = tid & 3, pq = tid >> 2;
  int offs[8];
  bool live[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = piece * 8 + e;
    const int tap = k / 3, cc = k % 3;
    live[e] = k < 27;
    offs[e] = live[e] ? (cc * (kIm2Rows + 2) + tap / 3) * P + tap % 3 : 0;
  }
  for (int rr = 0; rr < kIm2Rows; ++rr) {
    const int y = y0 + rr;
    if (y >= h) break;
    __nv_bfloat16* orow = out + (((size_t)n * h + y) * w + x0) * 32;
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      const int px = pq + 64 * k4;
      if (x0 + px >= w) continue;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = live[e] ? sm[offs[e] + rr * P + px] : 0.f;
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      reinterpret_cast<uint4*>(orow + (size_t)px * 32)[piece] = o;
    }
  }
}

// out[n][co][oy][ox] = bias[co] + sum of the (up to) 4 col entries that map to this output pixel
// grid (ceil(wo / 256), ho, n): one thread per output pixel, no index divisions
__global__ void __launch_bounds__(256)
col2im4x4s2_kernel(const __nv_bfloat16* __restrict__ col, const float* __restrict__ bias, float* __restrict__ out, int c,
                   int hi, int wi) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int ho = 2 * hi, wo = 2 * wi;
  const int ox = blockIdx.x * 256 + threadIdx.x;
  const int oy = blockIdx.y;
  const size_t n = blockIdx.z;
  if (ox >= wo) return;
  const int hy = oy >> 1, hx = ox >> 1;
  // oy = 2*iy - 1 + ky : even oy -> (iy=hy, ky=1), (hy-1, 3) ; odd oy -> (hy+1, 0), (hy, 2)
  const int iy_[2] = {(oy & 1) ? hy + 1 : hy, (oy & 1) ? hy : hy - 1};
  const int ky_[2] = {(oy & 1) ? 0 : 1, (oy & 1) ? 2 : 3};
  const int ix_[2] = {(ox & 1) ? hx + 1 : hx, (ox & 1) ? hx : hx - 1};
  const int kx_[2] = {(ox & 1) ? 0 : 1, (ox & 1) ? 2 : 3};
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    if (iy_[a] < 0 || iy_[a] >= hi) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      if (ix_[b] < 0 || ix_[b] >= wi) continue;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(col + ((n * hi + iy_[a]) * wi + ix_[b]) * 128) +
                            (ky_[a] * 4 + kx_[b]));
      acc[0] += bf16lo(u.x); acc[1] += bf16hi(u.x); acc[2] += bf16lo(u.y); acc[3] += bf16hi(u.y);
      acc[4] += bf16lo(u.z); acc[5] += bf16hi(u.z); acc[6] += bf16lo(u.w); acc[7] += bf16hi(u.w);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (e < c) out[((n * c + e) * ho + oy) * wo + ox] = acc[e] + (bias != nullptr ? bias[e] : 0.f);
}

// out[c] (+)= sum over n, hw of x[n][c][hw]   (bias gradient of the last layer, NCHW fp32 gradient)
__global__ void chansum_nchw_kernel(const float* __restrict__ x, int n, int ca, int c, int hw, float* __restrict__ out) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  __shared__ float red[32];
  const int cc = blockIdx.y;
  float acc = 0.f;
  const int hw4 = hw / 4;   // hw is a multiple of 4 (checked by the launcher)
  const size_t total = (size_t)n * hw4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t nn = i / hw4, p4 = i % hw4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (nn * ca + cc) * hw) + p4);
    acc += (v.x + v.y) + (v.z + v.w);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(out + cc, s);
  }
  (void)c;
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
  if (cs == 16 || cs == 32) {
    const size_t total = (size_t)n * hw;
    const int grid = grid_for(total, 256, num_sms, 16);
    if (cs == 16) pack_nchw_kernel<16><<<grid, 256, 0, st>>>(x, (__nv_bfloat16*)out, c, hw, shift, scale, total);
    else pack_nchw_kernel<32><<<grid, 256, 0, st>>>(x, (__nv_bfloat16*)out, c, hw, shift, scale, total);
  } else {
    const size_t total = (size_t)n * hw * cs;
    pack_nchw_wide_kernel<<<grid_for(total, 256, num_sms), 256, 0, st>>>(x, (__nv_bfloat16*)out, c, hw, cs, total);
  }
  return cudaGetLastError();
}
cudaError_t launch_unpack_nchw(const void* x, float* out, int n, int c, int hw, int cs, cudaStream_t st) {
  if (cs == 16) {
    const size_t total = (size_t)n * hw;
    (void)launch_k(unpack_nchw_small_kernel<16>, dim3(grid_for(total, 256, 148, 16)), dim3(256), 0, st, 1, (const __nv_bfloat16*)x, out, c, hw, total);
    return cudaGetLastError();
  }
  dim3 grid((hw + 31) / 32, n);
  (void)launch_k(unpack_nchw_kernel, grid, dim3(256), 0, st, 1, (const __nv_bfloat16*)x, out, c, hw, cs);
  return cudaGetLastError();
}
// ---------------------------------------------------------------- fused multi-tensor Adam
// Reference: optim.Adam(model.parameters(), lr=3e-4) train_faceoff_perceptual.py:190 (torch defaults: betas (0.9, 0.999),
// eps 1e-8, weight_decay 0, no amsgrad).  One launch updates every parameter tensor: `table` (device) lists the tensors,
// `chunks` (device) maps each 16,384-element chunk to (tensor, first element).  Same arithmetic, in the same order,
// as torch's single-tensor path: m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).
struct AdamTensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long long numel;
};
constexpr int kAdamChunk = 16384;
__global__ void __launch_bounds__(256)
adam_kernel(const AdamTensor* __restrict__ table, const int2* __restrict__ chunks, int n_chunks, float lr, float beta1,
            float beta2, float eps, float weight_decay, float bias_c1, float sqrt_bias_c2, float grad_scale) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int2 ch = chunks[c];
    const AdamTensor t = table[ch.x];
    const long long start = (long long)ch.y * kAdamChunk;
    const long long n = t.numel - start < kAdamChunk ? t.numel - start : kAdamChunk;
    const float step_size = lr / bias_c1;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const long long j = start + i;
      float g = t.grad[j] * grad_scale;
      const float p = t.param[j];
      if (weight_decay != 0.f) g = fmaf(weight_decay, p, g);
      // torch: exp_avg.lerp_(grad, 1 - beta1) ; exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
      const float m0 = t.exp_avg[j];
      const float m = m0 + (g - m0) * (1.f - beta1);
      const float v = t.exp_avg_sq[j] * beta2 + (1.f - beta2) * g * g;
      t.exp_avg[j] = m;
      t.exp_avg_sq[j] = v;
      const float denom = sqrtf(v) / sqrt_bias_c2 + eps;
      t.param[j] = p - step_size * (m / denom);
    }
  }
}
cudaError_t launch_adam(const void* table, const void* chunks, int n_chunks, float lr, float beta1, float beta2,
                        float eps, float weight_decay, float bias_c1, float sqrt_bias_c2, float grad_scale, int num_sms,
                        cudaStream_t st) {
  int blocks = n_chunks < num_sms * 8 ? n_chunks : num_sms * 8;
  if (blocks < 1) return cudaSuccess;
  adam_kernel<<<blocks, 256, 0, st>>>((const AdamTensor*)table, (const int2*)chunks, n_chunks, lr, beta1, beta2, eps,
                                      weight_decay, bias_c1, sqrt_bias_c2, grad_scale);
  return cudaGetLastError();
}

cudaError_t launch_u8hwc_to_nchw(const void* x, float* out, int n, int hw, int c_total, int c_off, float mean, float stdv,
                                 int num_sms, cudaStream_t st) {
  const size_t total4 = (size_t)n * (hw / 4);
  (void)launch_k(u8hwc_to_nchw_kernel, dim3(grid_for(total4, 256, num_sms, 16)), dim3(256), 0, st, 1, (const uint32_t*)x, out, hw / 4, total4, c_total,
                                                                           c_off, mean, stdv);
  return cudaGetLastError();
}
cudaError_t launch_relu(const void* x, void* y, size_t numel, int num_sms, cudaStream_t st) {
  const size_t nvec = numel / 8;
  (void)launch_k(relu_kernel, dim3(grid_for(nvec, 256, num_sms)), dim3(256), 0, st, 1, (const uint4*)x, (uint4*)y, nvec);
  return cudaGetLastError();
}
cudaError_t launch_pack_weights(const float* w, void* out, const PackParams& pp, int num_sms, cudaStream_t st) {
  const size_t total = (size_t)pp.npad * pp.ktot;
  pack_weights_kernel<<<grid_for(total, 256, num_sms), 256, 0, st>>>(w, (__nv_bfloat16*)out, pp);
  return cudaGetLastError();
}
cudaError_t launch_wgrad_finalize(const FinalizeParams& fp, const float* bias_part, int bias_rows, int bias_c, float* dbias,
                                  int dbias_accumulate, int num_sms, cudaStream_t st) {
  const size_t total = (size_t)fp.taps * fp.m_real * fp.n_real;
  size_t mb = fp.splits <= 16 ? (total + 255) / 256 : (total + 31) / 32;   // (thread | 8 warps x 32 lanes) per element(s)
  if (mb > (size_t)num_sms * 8) mb = (size_t)num_sms * 8;
  const int bias_blocks = dbias != nullptr ? (bias_c + 31) / 32 : 0;
  return launch_k(wgrad_finalize_kernel, dim3((unsigned)mb + bias_blocks), dim3(256), 0, st, 1, fp, (int)mb, bias_part, bias_rows,
                  bias_c, dbias, dbias_accumulate);
}
int colsum_blocks(int num_sms) { return num_sms * 8; }
cudaError_t launch_colsum(const void* x, size_t rows, int cs, int c_off, int c, float* out, int accumulate,
                          float* workspace, int num_sms, cudaStream_t st) {
  const int threads = 256;
  const int vecs = cs / 8;
  const int rpi = threads / vecs;
  int nblocks = colsum_blocks(num_sms);
  const size_t need = (rows + rpi - 1) / rpi;
  if ((size_t)nblocks > need) nblocks = (int)(need ? need : 1);
  (void)launch_k(colsum_partial_kernel, dim3(nblocks), dim3(threads), (size_t)rpi * cs * sizeof(float), st, 1, (const __nv_bfloat16*)x, rows, cs,
                                                                                     workspace);
  (void)launch_k(colsum_final_kernel, dim3((c + 31) / 32), dim3(1024), 0, st, 1, workspace, nblocks, cs, c_off, c, out, accumulate);
  return cudaGetLastError();
}
cudaError_t launch_maxpool2(const void* x, void* y, int n, int h, int w, int cs, int num_sms, cudaStream_t st) {
  const int vecs = cs / 8;
  const size_t total = (size_t)n * (h / 2) * (w / 2) * vecs;
  (void)launch_k(maxpool2_kernel, dim3(grid_for(total, 256, num_sms)), dim3(256), 0, st, 1, (const uint4*)x, (uint4*)y, n, h, w, vecs);
  return cudaGetLastError();
}
cudaError_t launch_maxpool2_bwd(const void* x, const void* y, const void* dy, void* dx, int n, int h, int w, int cs,
                                int num_sms, cudaStream_t st) {
  const int vecs = cs / 8;
  const size_t total = (size_t)n * (h / 2) * (w / 2) * vecs;
  (void)launch_k(maxpool2_bwd_kernel, dim3(grid_for(total, 256, num_sms, 16)), dim3(256), 0, st, 1, (const uint4*)x, (const uint4*)y,
                                                                        (const uint4*)dy, (uint4*)dx, n, h, w, vecs);
  return cudaGetLastError();
}

cudaError_t launch_im2col4x4s2(const float* x, void* out, int n, int ca, int c, int h, int w, cudaStream_t st) {
  dim3 grid((w / 2 + 63) / 64, h / 2, n);
  im2col4x4s2_kernel<<<grid, 256, 0, st>>>(x, (__nv_bfloat16*)out, ca, c, h, w);
  return cudaGetLastError();
}
cudaError_t launch_im2col3x3(const float* x, void* out, int n, int c, int h, int w, const float* shift,
                             const float* scale, cudaStream_t st) {
  dim3 grid((w + 255) / 256, (h + kIm2Rows - 1) / kIm2Rows, n);
  im2col3x3_kernel<<<grid, 256, 0, st>>>(x, (__nv_bfloat16*)out, c, h, w, shift, scale);
  return cudaGetLastError();
}
cudaError_t launch_col2im4x4s2(const void* col, const float* bias, float* out, int n, int c, int hi, int wi, int num_sms,
                               cudaStream_t st) {
  (void)num_sms;
  dim3 grid((2 * wi + 255) / 256, 2 * hi, n);
  (void)launch_k(col2im4x4s2_kernel, grid, dim3(256), 0, st, 1, (const __nv_bfloat16*)col, bias, out, c, hi, wi);
  return cudaGetLastError();
}
cudaError_t launch_chansum_nchw(const float* x, int n, int ca, int c, int hw, float* out, int accumulate, int num_sms,
                                cudaStream_t st) {
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * c, st);
    if (e != cudaSuccess) return e;
  }
  if (hw % 4 != 0) return cudaErrorInvalidValue;
  dim3 grid(num_sms * 2, c);
  (void)launch_k(chansum_nchw_kernel, grid, dim3(512), 0, st, 1, x, n, ca, c, hw, out);
  return cudaGetLastError();
}

}  // namespace fo
