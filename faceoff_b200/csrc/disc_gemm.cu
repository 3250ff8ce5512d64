// Tensor-core path of the MoCoGAN-HD discriminator convolutions (reference
// TemporalAlignment/models/mocoganhd_content_disc.py:49-165, mocoganhd_video_disc.py:55-176: Conv2d / Conv3d k4, stride 2 or 1,
// pad 2, on odd-sized maps 256 -> 129 -> 65 -> 33 -> 34 -> 35).
//
// These shapes (pad 2, odd extents, 4x4x4 filters) do not fit the parity-view forms of the implicit-GEMM planner, and the
// reference discriminators are fp32.  So each convolution becomes plain GEMMs on the SAME tcgen05 kernel (conv_igemm, 1x1
// form) over an explicit im2col matrix, in the hi|lo split-bf16 arithmetic of the verification mode (hi*w_hi + lo*w_hi +
// hi*w_lo, fp32 accumulation: ~2^-16 relative per product, i.e. fp32-accurate at 1/3 of the bf16 tensor rate):
//   forward   y[n, co, p]   = bias + sum_k col[n*P + p, k] * W[co, k]          col  = dim2col_pairs   (bf16 pairs [N*P, 2*Kp])
//   dgrad     dcol[p, k]    = sum_co dy[p, co] * W[co, k] ; dx = dcol2im(dcol)  (fp32 [N*P, Kp] -> NCDHW gather)
//   wgrad     dW[co, k]     = sum_p dy^T[co, p] * colT[k, p]                    colT = dim2col_t       (fp32 [chunks][K][Pc])
// with k = ((ci * kd + a) * kh + b) * kw + c, the order of the PyTorch weight [Cout, Cin, kD, kH, kW], so that W is the
// parameter itself.  The kernels here only move data (they are bandwidth-trivial next to the GEMMs: <= 0.3 GB per layer).
#include "common.cuh"
#include "kernels.h"

namespace fo {

struct KDecode {
  int ci, a, b, c;
};
__device__ __forceinline__ KDecode decode_k(const DConvParams& p, int k) {
  KDecode r;
  r.c = k % p.kw; k /= p.kw;
  r.b = k % p.kh; k /= p.kh;
  r.a = k % p.kd;
  r.ci = k / p.kd;
  return r;
}
__device__ __forceinline__ float im2col_value(const DConvParams& p, const float* __restrict__ x, int n, int od, int oh, int ow,
                                              int k, int K) {
  if (k >= K) return 0.f;
  const KDecode t = decode_k(p, k);
  const int id = od * p.sd - p.pd + t.a, ih = oh * p.sh - p.ph + t.b, iw = ow * p.sw - p.pw + t.c;
  if ((unsigned)id >= (unsigned)p.id || (unsigned)ih >= (unsigned)p.ih || (unsigned)iw >= (unsigned)p.iw) return 0.f;
  return __ldg(x + ((((size_t)n * p.cin + t.ci) * p.id + id) * p.ih + ih) * p.iw + iw);
}

// col[row = n*P + pos][part * kp + k], part = 0 .. PARTS-1: the im2col value split into PARTS bf16 terms (x0 = bf16(v),
// x1 = bf16(v - x0), x2 = bf16(v - x0 - x1)); one thread per (row, 8 consecutive k): PARTS 16-byte stores.
// PARTS = 2 (16 mantissa bits) serves the gradient GEMMs; the FORWARD uses 3 (24 bits): its rounding decides on which
// side of zero a LeakyReLU input falls, and a forward error of 2^-16 moved weight gradients by up to 5e-2 (measured
// against an fp64 restatement), while 2^-16 in the two gradient GEMMs stays at 1e-5.
template <int PARTS>
__global__ void __launch_bounds__(256)
dim2col_pairs_kernel(const DConvParams p, const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int K, int kp,
                     size_t total) {
  const int k8n = kp / 8;
  const int P = p.od * p.oh * p.ow;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % k8n) * 8;
    const size_t row = i / k8n;
    const int n = (int)(row / P);
    int pos = (int)(row % P);
    const int ow = pos % p.ow; pos /= p.ow;
    const int oh = pos % p.oh;
    const int od = pos / p.oh;
    float v[8];
    KDecode t = decode_k(p, k0);
    const int bd = od * p.sd - p.pd, bh = oh * p.sh - p.ph, bw = ow * p.sw - p.pw;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = 0.f;
      const int id = bd + t.a, ih = bh + t.b, iw = bw + t.c;
      if (k0 + e < K && (unsigned)id < (unsigned)p.id && (unsigned)ih < (unsigned)p.ih && (unsigned)iw < (unsigned)p.iw)
        v[e] = __ldg(x + ((((size_t)n * p.cin + t.ci) * p.id + id) * p.ih + ih) * p.iw + iw);
      if (++t.c == p.kw) { t.c = 0; if (++t.b == p.kh) { t.b = 0; if (++t.a == p.kd) { t.a = 0; ++t.ci; } } }
    }
    __nv_bfloat16* dst = col + row * (size_t)(PARTS * kp) + k0;
#pragma unroll
    for (int part = 0; part < PARTS; ++part) {
      uint4 h;
      h.x = pack_bf16x2(v[0], v[1]); h.y = pack_bf16x2(v[2], v[3]); h.z = pack_bf16x2(v[4], v[5]); h.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(dst + (size_t)part * kp) = h;
      v[0] -= bf16lo(h.x); v[1] -= bf16hi(h.x); v[2] -= bf16lo(h.y); v[3] -= bf16hi(h.y);
      v[4] -= bf16lo(h.z); v[5] -= bf16hi(h.z); v[6] -= bf16lo(h.w); v[7] -= bf16hi(h.w);
    }
  }
}

// Weight-side operand of the weight-gradient GEMM, written directly in conv_igemm's packed layout for a 1x1 form with the
// sources (hi, lo, hi): out[chunk = q / pc][k][3 * pc] bf16 = (hi | hi | lo) of the im2col value of global position
// q = n*P + pos (zero for q >= N*P and for k >= K), i.e. what fo_conv_pack_weights would produce from (w_hi, w_hi, w_lo).
// One thread per (chunk, k, 8 consecutive positions); the position is decoded once and advanced like an odometer.
__global__ void __launch_bounds__(256)
dim2col_t_kernel(const DConvParams p, const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int K, int kp, int pc,
                 size_t total) {
  const int P = p.od * p.oh * p.ow;
  const size_t rows_total = (size_t)p.n * P;
  const int j8n = pc / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j0 = (int)(i % j8n) * 8;
    const size_t r = i / j8n;
    const int k = (int)(r % kp);
    const size_t chunk = r / kp;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    size_t q = chunk * pc + j0;
    if (k < K && q < rows_total) {
      const KDecode t = decode_k(p, k);
      int n = (int)(q / P);
      int pos = (int)(q % P);
      int ow = pos % p.ow; pos /= p.ow;
      int oh = pos % p.oh;
      int od = pos / p.oh;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (q + e < rows_total) {
          const int id = od * p.sd - p.pd + t.a, ih = oh * p.sh - p.ph + t.b, iw = ow * p.sw - p.pw + t.c;
          if ((unsigned)id < (unsigned)p.id && (unsigned)ih < (unsigned)p.ih && (unsigned)iw < (unsigned)p.iw)
            v[e] = __ldg(x + ((((size_t)n * p.cin + t.ci) * p.id + id) * p.ih + ih) * p.iw + iw);
        }
        if (++ow == p.ow) { ow = 0; if (++oh == p.oh) { oh = 0; if (++od == p.od) { od = 0; ++n; } } }
      }
    }
    uint4 hi, lo;
    hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]); hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
    lo.x = pack_bf16x2(v[0] - bf16lo(hi.x), v[1] - bf16hi(hi.x)); lo.y = pack_bf16x2(v[2] - bf16lo(hi.y), v[3] - bf16hi(hi.y));
    lo.z = pack_bf16x2(v[4] - bf16lo(hi.z), v[5] - bf16hi(hi.z)); lo.w = pack_bf16x2(v[6] - bf16lo(hi.w), v[7] - bf16hi(hi.w));
    __nv_bfloat16* dst = out + (chunk * kp + k) * (size_t)(3 * pc) + j0;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + pc) = hi;
    *reinterpret_cast<uint4*>(dst + 2 * pc) = lo;
  }
}

// dx[n][ci][id][ih][iw] = sum over the taps (a, b, c) that reach this input element of dcol[n*P + pos(od, oh, ow)][k]
// (gather form: every dcol element is read exactly once, no atomics); one thread per input element
__global__ void __launch_bounds__(256)
dcol2im_kernel(const DConvParams p, const float* __restrict__ dcol, long long ld, float* __restrict__ dx, size_t total) {
  const int P = p.od * p.oh * p.ow;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int iw = (int)(r % p.iw); r /= p.iw;
    const int ih = (int)(r % p.ih); r /= p.ih;
    const int id = (int)(r % p.id); r /= p.id;
    const int ci = (int)(r % p.cin);
    const int n = (int)(r / p.cin);
    float acc = 0.f;
    for (int a = 0; a < p.kd; ++a) {
      const int td = id + p.pd - a;
      if (td < 0 || td % p.sd != 0) continue;
      const int od = td / p.sd;
      if (od >= p.od) continue;
      for (int b = 0; b < p.kh; ++b) {
        const int th = ih + p.ph - b;
        if (th < 0 || th % p.sh != 0) continue;
        const int oh = th / p.sh;
        if (oh >= p.oh) continue;
        for (int c = 0; c < p.kw; ++c) {
          const int tw = iw + p.pw - c;
          if (tw < 0 || tw % p.sw != 0) continue;
          const int ow = tw / p.sw;
          if (ow >= p.ow) continue;
          const size_t row = (size_t)n * P + ((size_t)od * p.oh + oh) * p.ow + ow;
          const int k = ((ci * p.kd + a) * p.kh + b) * p.kw + c;
          acc += __ldg(dcol + row * ld + k);
        }
      }
    }
    dx[i] = acc;
  }
}

// dbias[c] = sum over n and positions of dy[n][c][pos]; one block per channel
__global__ void __launch_bounds__(256)
dconv_dbias_kernel(const float* __restrict__ dy, int n, int c, int plane, float* __restrict__ dbias) {
  __shared__ float red[8];
  const int ch = blockIdx.x;
  float acc = 0.f;
  for (int s = 0; s < n; ++s) {
    const float* src = dy + ((size_t)s * c + ch) * plane;
    for (int i = threadIdx.x; i < plane; i += blockDim.x) acc += __ldg(src + i);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    dbias[ch] = s;
  }
}

static inline int blocks_for(size_t total, int num_sms) {
  size_t b = (total + 255) / 256;
  const size_t cap = (size_t)num_sms * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}
cudaError_t launch_dim2col_pairs(const DConvParams& p, const float* x, void* col, int kp, int parts, int num_sms,
                                 cudaStream_t st) {
  const int K = p.cin * p.kd * p.kh * p.kw;
  const size_t total = (size_t)p.n * p.od * p.oh * p.ow * (kp / 8);
  if (parts == 3)
    dim2col_pairs_kernel<3><<<blocks_for(total, num_sms), 256, 0, st>>>(p, x, (__nv_bfloat16*)col, K, kp, total);
  else
    dim2col_pairs_kernel<2><<<blocks_for(total, num_sms), 256, 0, st>>>(p, x, (__nv_bfloat16*)col, K, kp, total);
  return cudaGetLastError();
}
cudaError_t launch_dim2col_t(const DConvParams& p, const float* x, void* out, int kp, int pc, int chunks, int num_sms,
                             cudaStream_t st) {
  const int K = p.cin * p.kd * p.kh * p.kw;
  const size_t total = (size_t)chunks * kp * (pc / 8);
  dim2col_t_kernel<<<blocks_for(total, num_sms), 256, 0, st>>>(p, x, (__nv_bfloat16*)out, K, kp, pc, total);
  return cudaGetLastError();
}
cudaError_t launch_dcol2im(const DConvParams& p, const float* dcol, long long ld, float* dx, int num_sms, cudaStream_t st) {
  const size_t total = (size_t)p.n * p.cin * p.id * p.ih * p.iw;
  dcol2im_kernel<<<blocks_for(total, num_sms), 256, 0, st>>>(p, dcol, ld, dx, total);
  return cudaGetLastError();
}
cudaError_t launch_dconv_dbias(const float* dy, int n, int c, int plane, float* dbias, cudaStream_t st) {
  dconv_dbias_kernel<<<c, 256, 0, st>>>(dy, n, c, plane, dbias);
  return cudaGetLastError();
}

}  // namespace fo
