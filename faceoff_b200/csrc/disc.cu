// MoCoGAN-HD discriminator step (reference TemporalAlignment/models/mocoganhd_content_disc.py:8-165,
// mocoganhd_video_disc.py:8-176, mocoganhd_losses.py:109-126): kernel-4 Conv2d / Conv3d with stride 2 or 1 and padding 2
// on odd-sized maps (256 -> 129 -> 65 -> 33 -> 34 -> 35), InstanceNorm (track_running_stats) + LeakyReLU(0.2),
// AvgPool(3, stride 2, pad 1, count_include_pad=False) between the scales, relativistic average LSGAN loss.
//
// First implementation of this row: fp32 NCDHW (PyTorch-native layout, no conversions), the convolutions are tiled
// fp32 implicit GEMMs on the CUDA cores (128 x 128 x 8 tiles, 8 x 8 outputs per thread, operands gathered on the fly with
// zero padding).  The discriminators see ONE 2-frame pair / 11-frame clip per step (~0.15 TFLOP forward), so this path is
// ~3 % of the FLOPs of the batch-32 VQVAE+LPIPS step; moving these (odd-sized, pad-2, k4) forms onto the tcgen05 planner
// is listed as the next step in DESIGN.md.  Everything is exact fp32 arithmetic (parity rtol 1e-4 vs the reference).
#include "common.cuh"
#include "kernels.h"

namespace fo {

// ------------------------------------------------------------------------------------------ convolution (fp32)
__device__ __forceinline__ void decode_pos(int m, const DConvParams& p, bool out_side, int& n, int& d, int& h, int& w) {
  const int W = out_side ? p.ow : p.iw, H = out_side ? p.oh : p.ih, D = out_side ? p.od : p.id;
  w = m % W; m /= W;
  h = m % H; m /= H;
  d = m % D;
  n = m / D;
}

// MODE 0 (forward):  C[m = output position][n = co] = sum_k x[pos (+) tap][ci] * w[co][ci][tap],   k = (ci, tap)
// MODE 1 (dgrad):    C[m = input position][n = ci]  = sum_k dy[(pos + pad - tap) / stride][co] * w[co][ci][tap], k = (co, tap)
// 128 x 128 x 8 tiles, 256 threads, 8 x 8 outputs per thread (64 FMAs per 16 shared-memory operand reads).  The k -> (channel,
// tap) decode of a K slice is done once per slice by 8 threads into shared memory; every loader thread owns a fixed tile
// row (its position is decoded once), so a gathered element costs three adds, the bounds test and one address.
template <int MODE>
__global__ void __launch_bounds__(256)
dconv_gemm_kernel(const DConvParams p, const float* __restrict__ a_src, const float* __restrict__ wgt,
                  const float* __restrict__ bias, float* __restrict__ out) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int kt_ch[BK], kt_d[BK], kt_h[BK], kt_w[BK], kt_tap[BK];
  const int taps = p.kd * p.kh * p.kw;
  // dgrad of a strided convolution: an input position only receives the taps t = (i + p) mod s (per dimension).  The
  // input positions are therefore split into sd*sh*sw residue classes (blockIdx.z); within a class the K loop runs over the
  // matching taps only (1/8 of them for a stride-2 Conv3d) instead of multiplying zeros.
  int cd_ = 0, ch_ = 0, cw_ = 0, Jd = 1, Jh = 1, Jw = 1, t0d = 0, t0h = 0, t0w = 0, Td = p.kd, Th = p.kh, Tw = p.kw;
  if (MODE == 1) {
    const int cls = blockIdx.z;
    cw_ = cls % p.sw; ch_ = (cls / p.sw) % p.sh; cd_ = cls / (p.sw * p.sh);
    Jd = cd_ < p.id ? (p.id - cd_ + p.sd - 1) / p.sd : 0;
    Jh = ch_ < p.ih ? (p.ih - ch_ + p.sh - 1) / p.sh : 0;
    Jw = cw_ < p.iw ? (p.iw - cw_ + p.sw - 1) / p.sw : 0;
    t0d = (cd_ + p.pd) % p.sd; t0h = (ch_ + p.ph) % p.sh; t0w = (cw_ + p.pw) % p.sw;
    Td = t0d < p.kd ? (p.kd - t0d + p.sd - 1) / p.sd : 0;
    Th = t0h < p.kh ? (p.kh - t0h + p.sh - 1) / p.sh : 0;
    Tw = t0w < p.kw ? (p.kw - t0w + p.sw - 1) / p.sw : 0;
  }
  const int ctaps = MODE == 0 ? taps : Td * Th * Tw;      // taps walked by the K loop
  const int M = MODE == 0 ? p.n * p.od * p.oh * p.ow : p.n * Jd * Jh * Jw;
  const int N = MODE == 0 ? p.cout : p.cin;
  const int K = (MODE == 0 ? p.cin : p.cout) * ctaps;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= M) return;
  const int tid = threadIdx.x;
  // loaders: thread t fetches tile row / column (t % 128) for k = t / 128 + 2 j, j = 0..3
  const int lrow = tid & 127, kq = tid >> 7;
  const int am = m0 + lrow;
  const bool am_ok = am < M;
  int an = 0, ad = 0, ah = 0, aw = 0;
  auto decode_m = [&](int m, int& n, int& d, int& h, int& w) {
    if (MODE == 0) {
      decode_pos(m, p, true, n, d, h, w);
    } else {   // position inside the residue class -> input coordinates
      w = cw_ + p.sw * (m % Jw); m /= Jw;
      h = ch_ + p.sh * (m % Jh); m /= Jh;
      d = cd_ + p.sd * (m % Jd);
      n = m / Jd;
    }
  };
  if (am_ok) decode_m(am, an, ad, ah, aw);
  // base coordinates of the gather: forward o * s - p (+ tap); dgrad (i + p - t0) / s (- tap index inside the class)
  const int bd = MODE == 0 ? ad * p.sd - p.pd : (ad + p.pd - t0d) / p.sd;
  const int bh = MODE == 0 ? ah * p.sh - p.ph : (ah + p.ph - t0h) / p.sh;
  const int bw = MODE == 0 ? aw * p.sw - p.pw : (aw + p.pw - t0w) / p.sw;
  const long long src_plane = MODE == 0 ? (long long)p.id * p.ih * p.iw : (long long)p.od * p.oh * p.ow;
  const int src_c = MODE == 0 ? p.cin : p.cout;
  const float* a_img = a_src + (long long)an * src_c * src_plane;
  const int bn = n0 + lrow;
  const bool bn_ok = bn < N;
  const int tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads, 8 x 8 outputs each (rows ty*8.., columns tx*8..)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += BK) {
    if (tid < BK) {
      const int k = k0 + tid;
      int ch = -1, td = 0, th = 0, tw = 0, tap = 0;
      if (k < K) {
        ch = k / ctaps;
        const int ct = k % ctaps;
        if (MODE == 0) {
          tap = ct;
          tw = ct % p.kw;
          th = (ct / p.kw) % p.kh;
          td = ct / (p.kw * p.kh);
        } else {   // tap index inside the class -> filter tap t0 + s * t'
          tw = ct % Tw;
          th = (ct / Tw) % Th;
          td = ct / (Tw * Th);
          tap = ((t0d + p.sd * td) * p.kh + (t0h + p.sh * th)) * p.kw + (t0w + p.sw * tw);
        }
      }
      kt_ch[tid] = ch; kt_d[tid] = td; kt_h[tid] = th; kt_w[tid] = tw; kt_tap[tid] = tap;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kl = kq + 2 * j;
      const int ch = kt_ch[kl];
      float av = 0.f, bv = 0.f;
      if (ch >= 0) {
        if (am_ok) {
          if (MODE == 0) {
            const int id = bd + kt_d[kl], ih = bh + kt_h[kl], iw = bw + kt_w[kl];
            if ((unsigned)id < (unsigned)p.id && (unsigned)ih < (unsigned)p.ih && (unsigned)iw < (unsigned)p.iw)
              av = __ldg(a_img + (long long)ch * src_plane + ((long long)id * p.ih + ih) * p.iw + iw);
          } else {
            const int od = bd - kt_d[kl], oh = bh - kt_h[kl], ow = bw - kt_w[kl];
            if ((unsigned)od < (unsigned)p.od && (unsigned)oh < (unsigned)p.oh && (unsigned)ow < (unsigned)p.ow)
              av = __ldg(a_img + (long long)ch * src_plane + ((long long)od * p.oh + oh) * p.ow + ow);
          }
        }
        if (bn_ok) {
          // weight [cout][cin][taps]: forward B[k = (ci, tap)][n = co]; dgrad B[k = (co, tap)][n = ci]
          bv = MODE == 0 ? __ldg(wgt + ((long long)bn * p.cin + ch) * taps + kt_tap[kl])
                         : __ldg(wgt + ((long long)ch * p.cin + bn) * taps + kt_tap[kl]);
        }
      }
      As[kl][lrow] = av;
      Bs[kl][lrow] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]), b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const long long out_plane = MODE == 0 ? (long long)p.od * p.oh * p.ow : (long long)p.id * p.ih * p.iw;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    int n, d, h, w;
    decode_m(m, n, d, h, w);
    const int W = MODE == 0 ? p.ow : p.iw, H = MODE == 0 ? p.oh : p.ih;
    const long long sp = ((long long)d * H + h) * W + w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + tx * 8 + j;
      if (c >= N) continue;
      float v = acc[i][j];
      if (MODE == 0 && bias != nullptr) v += bias[c];
      out[((long long)n * N + c) * out_plane + sp] = v;
    }
  }
}

// weight gradient: dw[co][k = (ci, tap)] += sum_pos dy[pos][co] * x[pos (+) tap][ci]; positions split over blockIdx.z.
// 128 (co) x 128 (k) x 8 (positions) tiles; a loader thread owns one k column (its (ci, tap) decoded once) resp. one co
// row; the 8 positions of a slice are decoded once per slice into shared memory.
__global__ void __launch_bounds__(256)
dconv_wgrad_kernel(const DConvParams p, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                   int pos_per_split) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ float As[BK][BM + 4];   // [pos][co]
  __shared__ float Bs[BK][BN + 4];   // [pos][k]
  __shared__ int pt_n[BK], pt_d[BK], pt_h[BK], pt_w[BK];
  const int taps = p.kd * p.kh * p.kw;
  const int K = p.cin * taps;
  const int P = p.n * p.od * p.oh * p.ow;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int r_begin = blockIdx.z * pos_per_split, r_end = min(P, r_begin + pos_per_split);
  const int tid = threadIdx.x;
  const int col = tid & 127, rq = tid >> 7;
  const int co = m0 + col;
  const int k = n0 + col;
  const bool k_ok = k < K, co_ok = co < p.cout;
  const int ci = k_ok ? k / taps : 0, tap = k_ok ? k % taps : 0;
  const int tw = tap % p.kw - p.pw, th = (tap / p.kw) % p.kh - p.ph, td = tap / (p.kw * p.kh) - p.pd;
  const int tx = tid & 15, ty = tid >> 4;
  const long long oplane = (long long)p.od * p.oh * p.ow, iplane = (long long)p.id * p.ih * p.iw;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
    if (tid < BK) {
      const int r = r0 + tid;
      int n = -1, d = 0, h = 0, w = 0;
      if (r < r_end) decode_pos(r, p, true, n, d, h, w);
      pt_n[tid] = n; pt_d[tid] = d; pt_h[tid] = h; pt_w[tid] = w;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rl = rq + 2 * j;
      const int n = pt_n[rl];
      float av = 0.f, bv = 0.f;
      if (n >= 0) {
        const int d = pt_d[rl], h = pt_h[rl], w = pt_w[rl];
        if (co_ok) av = __ldg(dy + ((long long)n * p.cout + co) * oplane + ((long long)d * p.oh + h) * p.ow + w);
        if (k_ok) {
          const int id = d * p.sd + td, ih = h * p.sh + th, iw = w * p.sw + tw;
          if ((unsigned)id < (unsigned)p.id && (unsigned)ih < (unsigned)p.ih && (unsigned)iw < (unsigned)p.iw)
            bv = __ldg(x + ((long long)n * p.cin + ci) * iplane + ((long long)id * p.ih + ih) * p.iw + iw);
        }
      }
      As[rl][col] = av;
      Bs[rl][col] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]), b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = m0 + ty * 8 + i;
    if (c >= p.cout) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = n0 + tx * 8 + j;
      if (kk < K) atomicAdd(dw + (long long)c * K + kk, acc[i][j]);
    }
  }
}

// per-(n, c) plane sums: out[c] (+)= sum over n and the plane of x[n][c][:]  (bias gradient)
__global__ void plane_sum_kernel(const float* __restrict__ x, int n, int c, long long plane, float* __restrict__ out) {
  __shared__ float red[32];
  const int cc = blockIdx.x;
  float acc = 0.f;
  for (int nn = 0; nn < n; ++nn) {
    const float* px = x + ((long long)nn * c + cc) * plane;
    for (long long i = threadIdx.x; i < plane; i += blockDim.x) acc += px[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    out[cc] += s;
  }
}

// ------------------------------------------------------------------------------------------ InstanceNorm + LeakyReLU
// One block per (n, c) plane.  Training: instance statistics (biased variance for the normalisation, unbiased for the
// running estimate, momentum as nn.InstanceNorm*d(track_running_stats=True)); eval: the running statistics.
// y = lrelu((x - mean) * rstd); slope = 1 disables the activation.  save[2 * plane] = mean, rstd (for backward).
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  return s;
}
__global__ void __launch_bounds__(512)
instnorm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long plane, int c, float eps, float slope,
                    int training, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                    int n_batch, float* __restrict__ save) {
  __shared__ float red[32];
  const long long pl = blockIdx.x;
  const int ch = (int)(pl % c);
  const float* px = x + pl * plane;
  float mean, rstd;
  if (training) {
    float s = 0.f;
    for (long long i = threadIdx.x; i < plane; i += blockDim.x) s += px[i];
    mean = block_sum(s, red) / (float)plane;
    float q = 0.f;
    for (long long i = threadIdx.x; i < plane; i += blockDim.x) {
      const float d = px[i] - mean;
      q += d * d;
    }
    const float ss = block_sum(q, red);
    const float var = ss / (float)plane;
    rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0 && running_mean != nullptr) {
      // the running statistics average the instances of the batch (PyTorch reshapes to [1, N*C, ..] and means over N)
      const float unbiased = plane > 1 ? ss / (float)(plane - 1) : var;
      atomicAdd(running_mean + ch, momentum * mean / (float)n_batch);
      atomicAdd(running_var + ch, momentum * unbiased / (float)n_batch);
    }
  } else {
    mean = running_mean[ch];
    rstd = rsqrtf(running_var[ch] + eps);
  }
  if (threadIdx.x == 0 && save != nullptr) {
    save[2 * pl] = mean;
    save[2 * pl + 1] = rstd;
  }
  float* py = y + pl * plane;
  for (long long i = threadIdx.x; i < plane; i += blockDim.x) {
    const float v = (px[i] - mean) * rstd;
    py[i] = v > 0.f ? v : slope * v;
  }
}
// scale the running statistics by (1 - momentum) before the forward kernel adds momentum * batch mean
__global__ void running_decay_kernel(float* __restrict__ rm, float* __restrict__ rv, int c, float keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) {
    rm[i] *= keep;
    rv[i] *= keep;
  }
}
// dx for y = lrelu(xhat), xhat = (x - mean) * rstd:  g = dy * lrelu'(y);  training: dx = rstd * (g - mean(g) - xhat * mean(g xhat))
__global__ void __launch_bounds__(512)
instnorm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, long long plane,
                    float slope, int training, const float* __restrict__ save) {
  __shared__ float red[32];
  const long long pl = blockIdx.x;
  const float rstd = save[2 * pl + 1];
  const float* py = y + pl * plane;
  const float* pg = dy + pl * plane;
  float* pd = dx + pl * plane;
  const float inv_slope = 1.f / slope;
  float s1 = 0.f, s2 = 0.f;
  if (training) {
    for (long long i = threadIdx.x; i < plane; i += blockDim.x) {
      const float yy = py[i];
      const float g = yy > 0.f ? pg[i] : slope * pg[i];
      const float xh = yy > 0.f ? yy : yy * inv_slope;
      s1 += g;
      s2 += g * xh;
    }
    s1 = block_sum(s1, red) / (float)plane;
    s2 = block_sum(s2, red) / (float)plane;
  }
  for (long long i = threadIdx.x; i < plane; i += blockDim.x) {
    const float yy = py[i];
    const float g = yy > 0.f ? pg[i] : slope * pg[i];
    const float xh = yy > 0.f ? yy : yy * inv_slope;
    pd[i] = training ? rstd * (g - s1 - xh * s2) : rstd * g;
  }
}
// plain LeakyReLU (first layer of each discriminator: no norm): y = lrelu(x); backward dx = dy * lrelu'(y)
__global__ void lrelu_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, float slope) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = v > 0.f ? v : slope * v;
  }
}
__global__ void lrelu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, size_t n,
                                 float slope) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : slope * dy[i];
}

// ------------------------------------------------------------------------------------------ AvgPool(3, pad 1, count_include_pad=False)
// kernel (kd, 3, 3) with kd in {1, 3}, stride (sd, 2, 2), padding (kd / 2, 1, 1); the divisor is the number of in-range
// elements of the window.  x [planes = n*c][id][ih][iw] -> y [planes][od][oh][ow].
__global__ void avgpool3_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int id, int ih, int iw,
                                int od, int oh, int ow, int kd, int sd, int sh, int sw) {
  const long long total = planes * od * oh * ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int w = (int)(r % ow); r /= ow;
    const int h = (int)(r % oh); r /= oh;
    const int d = (int)(r % od);
    const long long pl = r / od;
    const float* px = x + pl * id * ih * iw;
    float s = 0.f;
    int cnt = 0;
    for (int a = 0; a < kd; ++a) {
      const int zd = d * sd - kd / 2 + a;
      if (zd < 0 || zd >= id) continue;
      for (int b = 0; b < 3; ++b) {
        const int zh = h * sh - 1 + b;
        if (zh < 0 || zh >= ih) continue;
        for (int c = 0; c < 3; ++c) {
          const int zw = w * sw - 1 + c;
          if (zw < 0 || zw >= iw) continue;
          s += px[((long long)zd * ih + zh) * iw + zw];
          ++cnt;
        }
      }
    }
    y[i] = s / (float)cnt;
  }
}
// gather form of the gradient: dx[pos] = sum over the windows containing pos of dy[window] / count(window)
__global__ void avgpool3_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long planes, int id, int ih,
                                    int iw, int od, int oh, int ow, int kd, int sd, int sh, int sw) {
  const long long total = planes * id * ih * iw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int w = (int)(r % iw); r /= iw;
    const int h = (int)(r % ih); r /= ih;
    const int d = (int)(r % id);
    const long long pl = r / id;
    const float* pg = dy + pl * od * oh * ow;
    float s = 0.f;
    for (int a = 0; a < kd; ++a) {
      const int qd = d + kd / 2 - a;
      if (qd < 0 || qd % sd != 0 || qd / sd >= od) continue;
      const int o_d = qd / sd;
      const int d_lo = max(0, o_d * sd - kd / 2), d_hi = min(id - 1, o_d * sd - kd / 2 + kd - 1);
      for (int b = 0; b < 3; ++b) {
        const int qh = h + 1 - b;
        if (qh < 0 || qh % sh != 0 || qh / sh >= oh) continue;
        const int o_h = qh / sh;
        const int h_lo = max(0, o_h * sh - 1), h_hi = min(ih - 1, o_h * sh + 1);
        for (int c = 0; c < 3; ++c) {
          const int qw = w + 1 - c;
          if (qw < 0 || qw % sw != 0 || qw / sw >= ow) continue;
          const int o_w = qw / sw;
          const int w_lo = max(0, o_w * sw - 1), w_hi = min(iw - 1, o_w * sw + 1);
          const int cnt = (d_hi - d_lo + 1) * (h_hi - h_lo + 1) * (w_hi - w_lo + 1);
          s += pg[((long long)o_d * oh + o_h) * ow + o_w] / (float)cnt;
        }
      }
    }
    dx[i] = s;
  }
}

// ------------------------------------------------------------------------------------------ relativistic average LSGAN
// loss = mean((a - mean(b) - t)^2)   (mocoganhd_losses.py:109-126 with nn.MSELoss);  single block.
// forward: out[0] = loss, out[1] = mean(b).  backward (g = upstream scalar): da = g * 2 (a - mb - t) / n,
// db = -g * (2 / n) * sum(a - mb - t) / m  (the same value for every element of b).
__global__ void __launch_bounds__(1024)
ralsgan_kernel(const float* __restrict__ a, int n, const float* __restrict__ b, int m, float target, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < m; i += blockDim.x) s += b[i];
  const float mb = block_sum(s, red) / (float)m;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = a[i] - mb - target;
    q += d * d;
  }
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    out[0] = q / (float)n;
    out[1] = mb;
  }
}
__global__ void __launch_bounds__(1024)
ralsgan_bwd_kernel(const float* __restrict__ a, int n, int m, float target, const float* __restrict__ fwd,
                   const float* __restrict__ g, float* __restrict__ da, float* __restrict__ db) {
  __shared__ float red[32];
  const float mb = fwd[1], gg = g[0] * 2.f / (float)n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = a[i] - mb - target;
    if (da != nullptr) da[i] = gg * d;
    s += d;
  }
  s = block_sum(s, red);
  if (db != nullptr) {
    const float v = -gg * s / (float)m;
    for (int i = threadIdx.x; i < m; i += blockDim.x) db[i] = v;
  }
}

// ------------------------------------------------------------------------------------------ launchers
static int nblocks(size_t total, int per_block = 256, int cap = 148 * 16) {
  size_t b = (total + per_block - 1) / per_block;
  return (int)(b < 1 ? 1 : b > (size_t)cap ? cap : b);
}
cudaError_t launch_dconv_fwd(const DConvParams& p, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  const long long M = (long long)p.n * p.od * p.oh * p.ow;
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)((p.cout + 127) / 128));
  dconv_gemm_kernel<0><<<grid, 256, 0, st>>>(p, x, w, bias, y);
  return cudaGetLastError();
}
cudaError_t launch_dconv_dgrad(const DConvParams& p, const float* dy, const float* w, float* dx, cudaStream_t st) {
  // one residue class of input positions per blockIdx.z; grid.x covers the largest class (class (0, 0, 0))
  const long long M = (long long)p.n * ((p.id + p.sd - 1) / p.sd) * ((p.ih + p.sh - 1) / p.sh) * ((p.iw + p.sw - 1) / p.sw);
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)((p.cin + 127) / 128), (unsigned)(p.sd * p.sh * p.sw));
  dconv_gemm_kernel<1><<<grid, 256, 0, st>>>(p, dy, w, nullptr, dx);
  return cudaGetLastError();
}
cudaError_t launch_dconv_wgrad(const DConvParams& p, const float* x, const float* dy, float* dw, float* dbias, int num_sms,
                               cudaStream_t st) {
  const int taps = p.kd * p.kh * p.kw;
  const long long K = (long long)p.cin * taps;
  const int P = p.n * p.od * p.oh * p.ow;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * p.cout * K, st);
  if (e != cudaSuccess) return e;
  const int tiles = (int)(((p.cout + 127) / 128) * ((K + 127) / 128));
  int splits = (4 * num_sms + tiles - 1) / tiles;
  const int max_splits = (P + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int per = (P + splits - 1) / splits;
  per = (per + 7) / 8 * 8;
  splits = (P + per - 1) / per;
  dim3 grid((unsigned)((p.cout + 127) / 128), (unsigned)((K + 127) / 128), (unsigned)splits);
  dconv_wgrad_kernel<<<grid, 256, 0, st>>>(p, x, dy, dw, per);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (dbias != nullptr) {
    if ((e = cudaMemsetAsync(dbias, 0, sizeof(float) * p.cout, st)) != cudaSuccess) return e;
    plane_sum_kernel<<<p.cout, 256, 0, st>>>(dy, p.n, p.cout, (long long)p.od * p.oh * p.ow, dbias);
  }
  return cudaGetLastError();
}
cudaError_t launch_instnorm_fwd(const float* x, float* y, int n, int c, long long plane, float eps, float slope, int training,
                                float momentum, float* running_mean, float* running_var, float* save, cudaStream_t st) {
  if (training && running_mean != nullptr) running_decay_kernel<<<(c + 255) / 256, 256, 0, st>>>(running_mean, running_var, c, 1.f - momentum);
  instnorm_fwd_kernel<<<(unsigned)((long long)n * c), 512, 0, st>>>(x, y, plane, c, eps, slope, training, momentum, running_mean,
                                                                   running_var, n, save);
  return cudaGetLastError();
}
cudaError_t launch_instnorm_bwd(const float* y, const float* dy, float* dx, int n, int c, long long plane, float slope,
                                int training, const float* save, cudaStream_t st) {
  instnorm_bwd_kernel<<<(unsigned)((long long)n * c), 512, 0, st>>>(y, dy, dx, plane, slope, training, save);
  return cudaGetLastError();
}
cudaError_t launch_lrelu(const float* x, float* y, size_t n, float slope, cudaStream_t st) {
  lrelu_kernel<<<nblocks(n), 256, 0, st>>>(x, y, n, slope);
  return cudaGetLastError();
}
cudaError_t launch_lrelu_bwd(const float* y, const float* dy, float* dx, size_t n, float slope, cudaStream_t st) {
  lrelu_bwd_kernel<<<nblocks(n), 256, 0, st>>>(y, dy, dx, n, slope);
  return cudaGetLastError();
}
cudaError_t launch_avgpool3(const float* x, float* y, long long planes, int id, int ih, int iw, int od, int oh, int ow, int kd,
                            int sd, int sh, int sw, cudaStream_t st) {
  avgpool3_kernel<<<nblocks((size_t)planes * od * oh * ow), 256, 0, st>>>(x, y, planes, id, ih, iw, od, oh, ow, kd, sd, sh, sw);
  return cudaGetLastError();
}
cudaError_t launch_avgpool3_bwd(const float* dy, float* dx, long long planes, int id, int ih, int iw, int od, int oh, int ow,
                                int kd, int sd, int sh, int sw, cudaStream_t st) {
  avgpool3_bwd_kernel<<<nblocks((size_t)planes * id * ih * iw), 256, 0, st>>>(dy, dx, planes, id, ih, iw, od, oh, ow, kd, sd, sh,
                                                                              sw);
  return cudaGetLastError();
}
cudaError_t launch_ralsgan(const float* a, int n, const float* b, int m, float target, float* out, cudaStream_t st) {
  ralsgan_kernel<<<1, 1024, 0, st>>>(a, n, b, m, target, out);
  return cudaGetLastError();
}
cudaError_t launch_ralsgan_bwd(const float* a, int n, int m, float target, const float* fwd, const float* g, float* da,
                               float* db, cudaStream_t st) {
  ralsgan_bwd_kernel<<<1, 1024, 0, st>>>(a, n, m, target, fwd, g, da, db);
  return cudaGetLastError();
}

}  // namespace fo
