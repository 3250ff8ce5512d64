// LPIPS head (reference models/lpips.py:80-93,155-161): per tap, channel-normalise both feature maps, squared
// difference, 1x1 "lin" conv to one channel, spatial mean -- fused into ONE bandwidth-bound pass that reads
// f0 and f1 exactly once (bf16 channels-last) and emits a scalar per image.  The reference materialises >= 10
// full-size temporaries per tap.
#include "common.cuh"
#include "kernels.h"

namespace fo {

constexpr float kLpipsEps = 1e-10f;  // models/lpips.py:155 -- added to the norm, outside the sqrt

// LPP lanes cooperate on one pixel; each lane holds VPL 16-byte vectors (8 channels) of f0 and f1.
// SPLIT (verification mode, fo_conv_t.split_out): a pixel holds c hi values followed by c lo values; the feature is hi + lo.
template <int VPL, bool SPLIT>
__device__ __forceinline__ void load_pix(const __nv_bfloat16* p, int lpp, int sub, int c, float (&v)[VPL * 8]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const uint4 u = __ldg(q + i * lpp + sub);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[i * 8 + 2 * e] = bf16lo(w[e]);
      v[i * 8 + 2 * e + 1] = bf16hi(w[e]);
    }
    if (SPLIT) {
      const uint4 ul = __ldg(q + c / 8 + i * lpp + sub);
      const uint32_t wl[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[i * 8 + 2 * e] += bf16lo(wl[e]);
        v[i * 8 + 2 * e + 1] += bf16hi(wl[e]);
      }
    }
  }
}
// Both kernels were instruction bound, not memory bound (~130 / ~230 issue slots per 16-byte vector pair: IEEE sqrt and
// division sequences, scalar FMAs, per-element selects).  Now: packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2, sm_100),
// sqrt.approx + rcp.approx (2 MUFU; relative error ~2^-22, far inside the 1e-5 parity bound), the packed bf16 vectors stay
// in registers until used, U pixels per thread and iteration.
template <int VPL, bool SPLIT>
struct PixRaw {
  uint4 hi[VPL];
  uint4 lo[SPLIT ? VPL : 1];
};
template <int VPL, bool SPLIT>
__device__ __forceinline__ void load_raw(const __nv_bfloat16* p, int lpp, int sub, int c, PixRaw<VPL, SPLIT>& r) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    r.hi[i] = __ldg(q + i * lpp + sub);
    if (SPLIT) r.lo[i] = __ldg(q + c / 8 + i * lpp + sub);
  }
}
// channel pairs (2e, 2e+1) of vector i -> v[i * 4 + e]
template <int VPL, bool SPLIT>
__device__ __forceinline__ void unpack_raw(const PixRaw<VPL, SPLIT>& r, float2 (&v)[VPL * 4]) {
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const uint32_t w[4] = {r.hi[i].x, r.hi[i].y, r.hi[i].z, r.hi[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) v[i * 4 + e] = make_float2(bf16lo(w[e]), bf16hi(w[e]));
    if (SPLIT) {
      const uint32_t wl[4] = {r.lo[i].x, r.lo[i].y, r.lo[i].z, r.lo[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) v[i * 4 + e] = __fadd2_rn(v[i * 4 + e], make_float2(bf16lo(wl[e]), bf16hi(wl[e])));
    }
  }
}
__device__ __forceinline__ float group_sum(float s, int lpp) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    if (o < lpp) s += __shfl_xor_sync(0xffffffffu, s, o);   // (a compile-time lpp leaves straight-line shuffles)
  return s;
}
__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <int N>
__device__ __forceinline__ float sum_sq(const float2 (&v)[N]) {
  float2 s = __fmul2_rn(v[0], v[0]);
#pragma unroll
  for (int j = 1; j < N; ++j) s = __ffma2_rn(v[j], v[j], s);
  return s.x + s.y;
}

template <int VPL, bool SPLIT, int U>
__global__ void lpips_tap_kernel(const __nv_bfloat16* __restrict__ f0, const __nv_bfloat16* __restrict__ f1,
                                 const float* __restrict__ w, int hw, int c, int lpp, float* __restrict__ out) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  __shared__ float red[32];
  const int n = blockIdx.y;
  const int ppw = 32 / lpp;                        // pixels per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % lpp, pw = lane / lpp;
  float2 wv[VPL * 4];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      wv[i * 4 + e] = make_float2(__ldg(w + (i * lpp + sub) * 8 + 2 * e), __ldg(w + (i * lpp + sub) * 8 + 2 * e + 1));
  float2 acc2 = make_float2(0.f, 0.f);
  const int pix_per_block = (blockDim.x >> 5) * ppw;
  for (int p0 = blockIdx.x * pix_per_block * U; p0 < hw; p0 += gridDim.x * pix_per_block * U) {
    PixRaw<VPL, SPLIT> ra[U], rb[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = p0 + u * pix_per_block + warp * ppw + pw;
      ok[u] = pix < hw;
      const size_t off = ((size_t)n * hw + (ok[u] ? pix : 0)) * c * (SPLIT ? 2 : 1);
      load_raw<VPL, SPLIT>(f0 + off, lpp, sub, c, ra[u]);
      load_raw<VPL, SPLIT>(f1 + off, lpp, sub, c, rb[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float2 a[VPL * 4], b[VPL * 4];
      unpack_raw<VPL, SPLIT>(ra[u], a);
      unpack_raw<VPL, SPLIT>(rb[u], b);
      const float s0 = group_sum(sum_sq(a), lpp), s1 = group_sum(sum_sq(b), lpp);
      const float i0 = rcp_fast(sqrt_fast(s0) + kLpipsEps), i1 = ok[u] ? -rcp_fast(sqrt_fast(s1) + kLpipsEps) : 0.f;
      const float2 i0v = make_float2(ok[u] ? i0 : 0.f, ok[u] ? i0 : 0.f), i1v = make_float2(i1, i1);
#pragma unroll
      for (int j = 0; j < VPL * 4; ++j) {
        const float2 t = __ffma2_rn(a[j], i0v, __fmul2_rn(b[j], i1v));   // a/n0 - b/n1 (0 for pixels past the end)
        acc2 = __ffma2_rn(__fmul2_rn(wv[j], t), t, acc2);
      }
    }
  }
  float acc = warp_sum(acc2.x + acc2.y);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(out + n, s / (float)hw);
  }
}

// The same tap for a feature map that feeds a 2x2 max pool: the unit of work is a pooling window, and the window maximum of
// f0 is written as the pooled tensor (bit-identical to maxpool2) -- the pool kernel's re-read of f0 goes away.
// (LPP = lanes per pixel is a template parameter here: with a run-time value the three shuffle reductions per pixel stay
// loops -- SHFL + FADD + SHF + ISETP + BRA per step, ~12 % of the instructions of these issue-bound kernels)
template <int VPL, int LPP>
__global__ void __launch_bounds__(256, VPL == 1 ? 3 : 1)
lpips_tap_pool_kernel(const __nv_bfloat16* __restrict__ f0, const __nv_bfloat16* __restrict__ f1, const float* __restrict__ w,
                      int h, int wd, int c, float* __restrict__ out, __nv_bfloat16* __restrict__ pooled,
                      __nv_bfloat16* __restrict__ pooled1) {
  constexpr int lpp = LPP;
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  __shared__ float red[32];
  const int n = blockIdx.y;
  const int ppw = 32 / lpp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % lpp, pw = lane / lpp;
  const int hw = h * wd, wo = wd / 2, nwin = (h / 2) * wo;
  float2 wv[VPL * 4];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      wv[i * 4 + e] = make_float2(__ldg(w + (i * lpp + sub) * 8 + 2 * e), __ldg(w + (i * lpp + sub) * 8 + 2 * e + 1));
  float2 acc2 = make_float2(0.f, 0.f);
  const int win_per_block = (blockDim.x >> 5) * ppw;
  for (int w0 = blockIdx.x * win_per_block; w0 < nwin; w0 += gridDim.x * win_per_block) {
    const int win = w0 + warp * ppw + pw;
    const bool ok = win < nwin;
    const int wvn = ok ? win : 0;
    const int wy = wvn / wo, wx = wvn - wy * wo;
    PixRaw<VPL, false> ra[4], rb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pix = (2 * wy + (u >> 1)) * wd + 2 * wx + (u & 1);
      const size_t off = ((size_t)n * hw + pix) * c;
      load_raw<VPL, false>(f0 + off, lpp, sub, c, ra[u]);
      load_raw<VPL, false>(f1 + off, lpp, sub, c, rb[u]);
    }
    if (ok) {
      uint4* dst = reinterpret_cast<uint4*>(pooled + ((size_t)n * nwin + win) * c);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        uint4 y;
        y.x = bf16x2_max(bf16x2_max(ra[0].hi[i].x, ra[1].hi[i].x), bf16x2_max(ra[2].hi[i].x, ra[3].hi[i].x));
        y.y = bf16x2_max(bf16x2_max(ra[0].hi[i].y, ra[1].hi[i].y), bf16x2_max(ra[2].hi[i].y, ra[3].hi[i].y));
        y.z = bf16x2_max(bf16x2_max(ra[0].hi[i].z, ra[1].hi[i].z), bf16x2_max(ra[2].hi[i].z, ra[3].hi[i].z));
        y.w = bf16x2_max(bf16x2_max(ra[0].hi[i].w, ra[1].hi[i].w), bf16x2_max(ra[2].hi[i].w, ra[3].hi[i].w));
        dst[i * lpp + sub] = y;
      }
      if (pooled1 != nullptr) {   // the other side's pool (the fixed side's trunk runs in lockstep: lpips.py)
        uint4* dst1 = reinterpret_cast<uint4*>(pooled1 + ((size_t)n * nwin + win) * c);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          uint4 y;
          y.x = bf16x2_max(bf16x2_max(rb[0].hi[i].x, rb[1].hi[i].x), bf16x2_max(rb[2].hi[i].x, rb[3].hi[i].x));
          y.y = bf16x2_max(bf16x2_max(rb[0].hi[i].y, rb[1].hi[i].y), bf16x2_max(rb[2].hi[i].y, rb[3].hi[i].y));
          y.z = bf16x2_max(bf16x2_max(rb[0].hi[i].z, rb[1].hi[i].z), bf16x2_max(rb[2].hi[i].z, rb[3].hi[i].z));
          y.w = bf16x2_max(bf16x2_max(rb[0].hi[i].w, rb[1].hi[i].w), bf16x2_max(rb[2].hi[i].w, rb[3].hi[i].w));
          dst1[i * lpp + sub] = y;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 a[VPL * 4], b[VPL * 4];
      unpack_raw<VPL, false>(ra[u], a);
      unpack_raw<VPL, false>(rb[u], b);
      const float s0 = group_sum(sum_sq(a), lpp), s1 = group_sum(sum_sq(b), lpp);
      const float i0 = rcp_fast(sqrt_fast(s0) + kLpipsEps), i1 = ok ? -rcp_fast(sqrt_fast(s1) + kLpipsEps) : 0.f;
      const float2 i0v = make_float2(ok ? i0 : 0.f, ok ? i0 : 0.f), i1v = make_float2(i1, i1);
#pragma unroll
      for (int j = 0; j < VPL * 4; ++j) {
        const float2 t = __ffma2_rn(a[j], i0v, __fmul2_rn(b[j], i1v));   // a/n0 - b/n1 (0 for windows past the end)
        acc2 = __ffma2_rn(__fmul2_rn(wv[j], t), t, acc2);
      }
    }
  }
  float acc = warp_sum(acc2.x + acc2.y);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(out + n, s / (float)hw);
  }
}

// d/df0 of the tap value, times g[n], gated by the ReLU that produced f0, plus optional addend (pool gradient).
//   a = f0/n0, n0 = |f0| + eps ;  u_c = (2/hw) w_c (a_c - b_c)
//   dL/df0_j = u_j / n0 - (sum_c u_c f0_c) f0_j / (n0^2 |f0|)
template <int VPL, bool SPLIT, int U>
__global__ void lpips_tap_bwd_kernel(const __nv_bfloat16* __restrict__ f0, const __nv_bfloat16* __restrict__ f1,
                                     const float* __restrict__ w, const float* __restrict__ g, int hw, int c, int lpp,
                                     __nv_bfloat16* __restrict__ d_f0, const __nv_bfloat16* __restrict__ addend) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int n = blockIdx.y;
  const int ppw = 32 / lpp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % lpp, pw = lane / lpp;
  const float gn = g[n] * 2.f / (float)hw;
  float2 gw[VPL * 4];   // (2 g / hw) w_c
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      gw[i * 4 + e] = make_float2(gn * __ldg(w + (i * lpp + sub) * 8 + 2 * e), gn * __ldg(w + (i * lpp + sub) * 8 + 2 * e + 1));
  const int pix_per_block = (blockDim.x >> 5) * ppw;
  for (int p0 = blockIdx.x * pix_per_block * U; p0 < hw; p0 += gridDim.x * pix_per_block * U) {
    // all loads of the U pixels first (f0, f1, pool-gradient addend)
    PixRaw<VPL, SPLIT> ra[U], rb[U], rad[U];
    bool ok[U];
    size_t off[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = p0 + u * pix_per_block + warp * ppw + pw;
      ok[u] = pix < hw;
      off[u] = ((size_t)n * hw + (ok[u] ? pix : 0)) * c * (SPLIT ? 2 : 1);
      load_raw<VPL, SPLIT>(f0 + off[u], lpp, sub, c, ra[u]);
      load_raw<VPL, SPLIT>(f1 + off[u], lpp, sub, c, rb[u]);
      if (addend != nullptr) load_raw<VPL, SPLIT>(addend + off[u], lpp, sub, c, rad[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float2 a[VPL * 4], b[VPL * 4];
      unpack_raw<VPL, SPLIT>(ra[u], a);
      unpack_raw<VPL, SPLIT>(rb[u], b);
      const float s0 = group_sum(sum_sq(a), lpp), s1 = group_sum(sum_sq(b), lpp);
      const float r0 = sqrt_fast(s0);
      const float i0 = rcp_fast(r0 + kLpipsEps), i1 = -rcp_fast(sqrt_fast(s1) + kLpipsEps);
      const float2 i0v = make_float2(i0, i0), i1v = make_float2(i1, i1);
      float2 uu[VPL * 4];
      float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < VPL * 4; ++j) {
        uu[j] = __fmul2_rn(gw[j], __ffma2_rn(a[j], i0v, __fmul2_rn(b[j], i1v)));
        dot2 = __ffma2_rn(uu[j], a[j], dot2);
      }
      const float dot = group_sum(dot2.x + dot2.y, lpp);
      const float k2 = r0 > 0.f ? -dot * i0 * i0 * rcp_fast(r0) : 0.f;
      const float2 k2v = make_float2(k2, k2);
      float2 ad[VPL * 4];
      if (addend != nullptr) {
        unpack_raw<VPL, SPLIT>(rad[u], ad);
      } else {
#pragma unroll
        for (int j = 0; j < VPL * 4; ++j) ad[j] = make_float2(0.f, 0.f);
      }
      if (ok[u]) {
        uint4* dst = reinterpret_cast<uint4*>(d_f0 + off[u]);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          float2 o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = i * 4 + e;
            // u / n0 - k f0 + addend, gated by the ReLU of the tap (the addend is already gated)
            const float2 gr = __ffma2_rn(uu[j], i0v, __ffma2_rn(k2v, a[j], ad[j]));
            o[e] = make_float2(a[j].x > 0.f ? gr.x : 0.f, a[j].y > 0.f ? gr.y : 0.f);
          }
          uint4 ov;
          ov.x = pack_bf16x2(o[0].x, o[0].y); ov.y = pack_bf16x2(o[1].x, o[1].y);
          ov.z = pack_bf16x2(o[2].x, o[2].y); ov.w = pack_bf16x2(o[3].x, o[3].y);
          dst[i * lpp + sub] = ov;
          if (SPLIT) {   // lo = bf16(value - hi)
            const uint32_t wv2[4] = {ov.x, ov.y, ov.z, ov.w};
            uint4 ol;
            ol.x = pack_bf16x2(o[0].x - bf16lo(wv2[0]), o[0].y - bf16hi(wv2[0]));
            ol.y = pack_bf16x2(o[1].x - bf16lo(wv2[1]), o[1].y - bf16hi(wv2[1]));
            ol.z = pack_bf16x2(o[2].x - bf16lo(wv2[2]), o[2].y - bf16hi(wv2[2]));
            ol.w = pack_bf16x2(o[3].x - bf16lo(wv2[3]), o[3].y - bf16hi(wv2[3]));
            dst[c / 8 + i * lpp + sub] = ol;
          }
        }
      }
    }
  }
}

// The same gradient for a tap that feeds a 2x2 max pool (relu1_2 .. relu4_3 of the VGG trunk, reference models/lpips.py:
// 115-152), with the pool's backward folded in: the unit of work is a pooling WINDOW.  A lane group loads f0 / f1 of the four
// pixels and the pooled gradient dy of the window, recomputes the window maximum from f0 (bit equality on bf16, first
// position in row-major window order wins, closed gate at 0 -- exactly maxpool2_bwd), and adds that gradient where the
// unfused path read it back as `addend`.  Results are bit-identical to maxpool2_bwd + lpips_tap_bwd(addend); the traffic per
// pixel drops from 6.5 to 3.25 feature-map units (no dx write + read-back, no re-read of f0, no pooled y).
template <int VPL, int LPP>
__global__ void __launch_bounds__(256, VPL == 1 ? 3 : 1)
lpips_tap_bwd_pool_kernel(const __nv_bfloat16* __restrict__ f0, const __nv_bfloat16* __restrict__ f1,
                          const float* __restrict__ w, const float* __restrict__ g, int h, int wd, int c,
                          __nv_bfloat16* __restrict__ d_f0, const __nv_bfloat16* __restrict__ pool_dy) {
  constexpr int lpp = LPP;
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int n = blockIdx.y;
  const int ppw = 32 / lpp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % lpp, pw = lane / lpp;
  const int hw = h * wd, wo = wd / 2, nwin = (h / 2) * wo;
  const float gn = g[n] * 2.f / (float)hw;
  float2 gw[VPL * 4];   // (2 g / hw) w_c
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      gw[i * 4 + e] = make_float2(gn * __ldg(w + (i * lpp + sub) * 8 + 2 * e), gn * __ldg(w + (i * lpp + sub) * 8 + 2 * e + 1));
  const int win_per_block = (blockDim.x >> 5) * ppw;
  for (int w0 = blockIdx.x * win_per_block; w0 < nwin; w0 += gridDim.x * win_per_block) {
    const int win = w0 + warp * ppw + pw;
    const bool ok = win < nwin;
    const int wv = ok ? win : 0;
    const int wy = wv / wo, wx = wv - wy * wo;
    PixRaw<VPL, false> ra[4], rb[4], rg;
    size_t off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pix = (2 * wy + (u >> 1)) * wd + 2 * wx + (u & 1);
      off[u] = ((size_t)n * hw + pix) * c;
      load_raw<VPL, false>(f0 + off[u], lpp, sub, c, ra[u]);
      load_raw<VPL, false>(f1 + off[u], lpp, sub, c, rb[u]);
    }
    load_raw<VPL, false>(pool_dy + ((size_t)n * nwin + wv) * c, lpp, sub, c, rg);
    // pool gradient of the four pixels (bf16 pairs), packed: m_u = (x_u == max) and no earlier position matched; the gate
    // (x_u > 0) is max > 0 wherever x_u == max.  (~18 instructions per 32-bit word for the whole window; the per-half
    // scalar form of maxpool2_bwd made this kernel instruction bound: 3.1 TB/s.)
    PixRaw<VPL, false> pg[4];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const uint32_t x0[4] = {ra[0].hi[i].x, ra[0].hi[i].y, ra[0].hi[i].z, ra[0].hi[i].w};
      const uint32_t x1[4] = {ra[1].hi[i].x, ra[1].hi[i].y, ra[1].hi[i].z, ra[1].hi[i].w};
      const uint32_t x2[4] = {ra[2].hi[i].x, ra[2].hi[i].y, ra[2].hi[i].z, ra[2].hi[i].w};
      const uint32_t x3[4] = {ra[3].hi[i].x, ra[3].hi[i].y, ra[3].hi[i].z, ra[3].hi[i].w};
      const uint32_t gg[4] = {rg.hi[i].x, rg.hi[i].y, rg.hi[i].z, rg.hi[i].w};
      uint32_t o[4][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t y = bf16x2_max(bf16x2_max(x0[e], x1[e]), bf16x2_max(x2[e], x3[e]));
        const uint32_t gated = gg[e] & bf16x2_gt_mask(y, 0u);
        const uint32_t m0 = bf16x2_eq_mask(x0[e], y);
        const uint32_t m1 = bf16x2_eq_mask(x1[e], y) & ~m0;
        const uint32_t m01 = m0 | m1;
        const uint32_t m2 = bf16x2_eq_mask(x2[e], y) & ~m01;
        const uint32_t m3 = ~(m01 | m2);          // the maximum is one of the four
        o[0][e] = gated & m0;
        o[1][e] = gated & m1;
        o[2][e] = gated & m2;
        o[3][e] = gated & m3;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) pg[u].hi[i] = make_uint4(o[u][0], o[u][1], o[u][2], o[u][3]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 a[VPL * 4], b[VPL * 4], ad[VPL * 4];
      unpack_raw<VPL, false>(ra[u], a);
      unpack_raw<VPL, false>(rb[u], b);
      unpack_raw<VPL, false>(pg[u], ad);
      const float s0 = group_sum(sum_sq(a), lpp), s1 = group_sum(sum_sq(b), lpp);
      const float r0 = sqrt_fast(s0);
      const float i0 = rcp_fast(r0 + kLpipsEps), i1 = -rcp_fast(sqrt_fast(s1) + kLpipsEps);
      const float2 i0v = make_float2(i0, i0), i1v = make_float2(i1, i1);
      float2 uu[VPL * 4];
      float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < VPL * 4; ++j) {
        uu[j] = __fmul2_rn(gw[j], __ffma2_rn(a[j], i0v, __fmul2_rn(b[j], i1v)));
        dot2 = __ffma2_rn(uu[j], a[j], dot2);
      }
      const float dot = group_sum(dot2.x + dot2.y, lpp);
      const float k2 = r0 > 0.f ? -dot * i0 * i0 * rcp_fast(r0) : 0.f;
      const float2 k2v = make_float2(k2, k2);
      if (ok) {
        uint4* dst = reinterpret_cast<uint4*>(d_f0 + off[u]);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          float2 o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = i * 4 + e;
            const float2 gr = __ffma2_rn(uu[j], i0v, __ffma2_rn(k2v, a[j], ad[j]));
            o[e] = make_float2(a[j].x > 0.f ? gr.x : 0.f, a[j].y > 0.f ? gr.y : 0.f);
          }
          uint4 ov;
          ov.x = pack_bf16x2(o[0].x, o[0].y); ov.y = pack_bf16x2(o[1].x, o[1].y);
          ov.z = pack_bf16x2(o[2].x, o[2].y); ov.w = pack_bf16x2(o[3].x, o[3].y);
          dst[i * lpp + sub] = ov;
        }
      }
    }
  }
}

static void lpips_geometry(int c, int& lpp, int& vpl) {
  const int vecs = c / 8;
  lpp = vecs < 32 ? vecs : 32;
  vpl = vecs / lpp;
}

cudaError_t launch_lpips_tap(const void* f0, const void* f1, const float* w, int n, int hw, int c, float* out,
                             int num_sms, cudaStream_t st, int split) {
  int lpp, vpl;
  lpips_geometry(c, lpp, vpl);
  const int threads = 256;
  const int pix_per_block = (threads / 32) * (32 / lpp);
  int bx = (hw + pix_per_block - 1) / pix_per_block;
  const int cap = (num_sms * 8 + n - 1) / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
  const __nv_bfloat16 *a = (const __nv_bfloat16*)f0, *b = (const __nv_bfloat16*)f1;
  if (vpl == 1 && !split) (void)launch_k(lpips_tap_kernel<1, false, 2>, grid, dim3(threads), 0, st, 1, a, b, w, hw, c, lpp, out);
  else if (vpl == 2 && !split) (void)launch_k(lpips_tap_kernel<2, false, 2>, grid, dim3(threads), 0, st, 1, a, b, w, hw, c, lpp, out);
  else if (vpl == 1) (void)launch_k(lpips_tap_kernel<1, true, 1>, grid, dim3(threads), 0, st, 1, a, b, w, hw, c, lpp, out);
  else if (vpl == 2) (void)launch_k(lpips_tap_kernel<2, true, 1>, grid, dim3(threads), 0, st, 1, a, b, w, hw, c, lpp, out);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
cudaError_t launch_lpips_tap_bwd(const void* f0, const void* f1, const float* w, const float* g, int n, int hw, int c,
                                 void* d_f0, const void* addend, int num_sms, cudaStream_t st, int split) {
  int lpp, vpl;
  lpips_geometry(c, lpp, vpl);
  const int threads = 256;
  const int pix_per_block = (threads / 32) * (32 / lpp);
  int bx = (hw + pix_per_block - 1) / pix_per_block;
  const int cap = (num_sms * 8 + n - 1) / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
  const __nv_bfloat16 *a = (const __nv_bfloat16*)f0, *b = (const __nv_bfloat16*)f1;
  __nv_bfloat16* d = (__nv_bfloat16*)d_f0;
  const __nv_bfloat16* ad = (const __nv_bfloat16*)addend;
  if (vpl == 1 && !split) (void)launch_k(lpips_tap_bwd_kernel<1, false, 2>, grid, dim3(threads), 0, st, 1, a, b, w, g, hw, c, lpp, d, ad);
  else if (vpl == 2 && !split) (void)launch_k(lpips_tap_bwd_kernel<2, false, 1>, grid, dim3(threads), 0, st, 1, a, b, w, g, hw, c, lpp, d, ad);
  else if (vpl == 1) (void)launch_k(lpips_tap_bwd_kernel<1, true, 1>, grid, dim3(threads), 0, st, 1, a, b, w, g, hw, c, lpp, d, ad);
  else if (vpl == 2) (void)launch_k(lpips_tap_bwd_kernel<2, true, 1>, grid, dim3(threads), 0, st, 1, a, b, w, g, hw, c, lpp, d, ad);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_lpips_tap_pool(const void* f0, const void* f1, const float* w, int n, int h, int wd, int c, float* out,
                                  void* pooled, void* pooled1, int num_sms, cudaStream_t st) {
  int lpp, vpl;
  lpips_geometry(c, lpp, vpl);
  const int threads = 256;
  const int win_per_block = (threads / 32) * (32 / lpp);
  const int nwin = (h / 2) * (wd / 2);
  int bx = (nwin + win_per_block - 1) / win_per_block;
  const int cap = (num_sms * 8 + n - 1) / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
  const __nv_bfloat16 *a = (const __nv_bfloat16*)f0, *b = (const __nv_bfloat16*)f1;
  __nv_bfloat16 *y = (__nv_bfloat16*)pooled, *y1 = (__nv_bfloat16*)pooled1;
  if (vpl == 1 && lpp == 8) return launch_k(lpips_tap_pool_kernel<1, 8>, grid, dim3(threads), 0, st, 1, a, b, w, h, wd, c, out, y, y1);
  if (vpl == 1 && lpp == 16) return launch_k(lpips_tap_pool_kernel<1, 16>, grid, dim3(threads), 0, st, 1, a, b, w, h, wd, c, out, y, y1);
  if (vpl == 1 && lpp == 32) return launch_k(lpips_tap_pool_kernel<1, 32>, grid, dim3(threads), 0, st, 1, a, b, w, h, wd, c, out, y, y1);
  if (vpl == 2 && lpp == 32) return launch_k(lpips_tap_pool_kernel<2, 32>, grid, dim3(threads), 0, st, 1, a, b, w, h, wd, c, out, y, y1);
  return cudaErrorInvalidValue;
}
cudaError_t launch_lpips_tap_bwd_pool(const void* f0, const void* f1, const float* w, const float* g, int n, int h, int wd,
                                      int c, void* d_f0, const void* pool_dy, int num_sms, cudaStream_t st) {
  int lpp, vpl;
  lpips_geometry(c, lpp, vpl);
  const int threads = 256;
  const int win_per_block = (threads / 32) * (32 / lpp);
  const int nwin = (h / 2) * (wd / 2);
  int bx = (nwin + win_per_block - 1) / win_per_block;
  const int cap = (num_sms * 8 + n - 1) / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
  const __nv_bfloat16 *a = (const __nv_bfloat16*)f0, *b = (const __nv_bfloat16*)f1, *pd = (const __nv_bfloat16*)pool_dy;
  __nv_bfloat16* d = (__nv_bfloat16*)d_f0;
  if (vpl == 1 && lpp == 8) return launch_k(lpips_tap_bwd_pool_kernel<1, 8>, grid, dim3(threads), 0, st, 1, a, b, w, g, h, wd, c, d, pd);
  if (vpl == 1 && lpp == 16) return launch_k(lpips_tap_bwd_pool_kernel<1, 16>, grid, dim3(threads), 0, st, 1, a, b, w, g, h, wd, c, d, pd);
  if (vpl == 1 && lpp == 32) return launch_k(lpips_tap_bwd_pool_kernel<1, 32>, grid, dim3(threads), 0, st, 1, a, b, w, g, h, wd, c, d, pd);
  if (vpl == 2 && lpp == 32) return launch_k(lpips_tap_bwd_pool_kernel<2, 32>, grid, dim3(threads), 0, st, 1, a, b, w, g, h, wd, c, d, pd);
  return cudaErrorInvalidValue;
}

// sum((a[:, :c] - b)^2), NCHW fp32
// sum((a[:, :c] - b)^2) and, with GRAD, grad[n, ca, hw] = g * (a[:, :c] - b) (channels >= c: 0), 16-byte accesses.
// a: [n, ca, hw] fp32 NCHW, b: [n, c, hw]; hw % 4 == 0.
template <bool GRAD>
__global__ void __launch_bounds__(256)
mse_kernel(const float4* __restrict__ a, const float4* __restrict__ b, int ca, int c, int hw4, size_t total4,
           float* __restrict__ out, const float* __restrict__ gscale, float scale, float4* __restrict__ grad) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  __shared__ float red[8];
  float acc = 0.f;
  const float g = GRAD ? __ldg(gscale) * scale : 0.f;
  // GRAD walks a's index space (it must also zero the channels beyond c), the plain sum walks b's
  const size_t per_n = (size_t)(GRAD ? ca : c) * hw4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < total4; i0 += 4 * stride) {
    float4 va[4], vb[4];
    size_t ia[4];
    bool live[4], in_c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = i0 + u * stride;
      live[u] = i < total4;
      const size_t n = live[u] ? i / per_n : 0, r = live[u] ? i % per_n : 0;
      in_c[u] = live[u] && r < (size_t)c * hw4;
      ia[u] = n * (size_t)ca * hw4 + r;
      va[u] = vb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in_c[u]) {
        va[u] = __ldg(a + ia[u]);
        vb[u] = __ldg(b + n * (size_t)c * hw4 + r);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 d = make_float4(va[u].x - vb[u].x, va[u].y - vb[u].y, va[u].z - vb[u].z, va[u].w - vb[u].w);
      if (GRAD) {
        if (live[u]) grad[ia[u]] = make_float4(g * d.x, g * d.y, g * d.z, g * d.w);
      } else {
        acc += (d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w);
      }
    }
  }
  if (!GRAD) {
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
      atomicAdd(out, s);
    }
  }
}
cudaError_t launch_mse(const float* a, const float* b, int n, int ca, int c, int hw, float* sum_out, int num_sms,
                       cudaStream_t st) {
  const size_t total4 = (size_t)n * c * (hw / 4);
  size_t blocks = (total4 + 1023) / 1024;
  if (blocks > (size_t)num_sms * 8) blocks = (size_t)num_sms * 8;
  if (blocks < 1) blocks = 1;
  (void)launch_k(mse_kernel<false>, dim3((int)blocks), dim3(256), 0, st, 1, (const float4*)a, (const float4*)b, ca, c, hw / 4, total4, sum_out,
                                                 nullptr, 0.f, nullptr);
  return cudaGetLastError();
}
cudaError_t launch_mse_grad(const float* a, const float* b, int n, int ca, int c, int hw, const float* gscale,
                            float scale, float* grad, int num_sms, cudaStream_t st) {
  const size_t total4 = (size_t)n * ca * (hw / 4);
  size_t blocks = (total4 + 1023) / 1024;
  if (blocks > (size_t)num_sms * 8) blocks = (size_t)num_sms * 8;
  if (blocks < 1) blocks = 1;
  (void)launch_k(mse_kernel<true>, dim3((int)blocks), dim3(256), 0, st, 1, (const float4*)a, (const float4*)b, ca, c, hw / 4, total4, nullptr,
                                                gscale, scale, (float4*)grad);
  return cudaGetLastError();
}

}  // namespace fo
