// Vector quantiser kernels (reference models/vqvae_conv3d_latent.py:33-83).
//
//  vq_prep        codebook fp32 [dim, n_embed] -> bf16 hi/lo split [n_embed, 2*dim] (K-major GEMM B operand),
//                 transposed fp32 copy [n_embed, dim] (gather operand) and |e|^2.
//  vq_assign      nearest code per row (:48-54).  The distance matrix is a dense contraction, so it runs on
//                 tcgen05: A = x split into bf16 hi + lo on the fly (fp32 rows are loaded coalesced, split in
//                 registers, written into 128B-swizzled smem tiles), B = codebook hi/lo tiles streamed by TMA,
//                 three MMAs per K block (hi*hi + lo*hi + hi*lo ~ 2^-17 relative error), accumulators
//                 double-buffered in TMEM.  The epilogue keeps (best, second best, argbest) per row in registers
//                 -- the [rows, n_embed] distance matrix never exists in HBM.  Rows whose top-2 gap is inside
//                 the error band are re-evaluated exactly (fp64 accumulation of the fp32 data) by vq_refine,
//                 which makes embed_ind bit-exact w.r.t. the reference outside true near-ties.
//  vq_gather_stats gather + straight-through + commitment loss + EMA statistics in one pass (:55-61,77-78).
//  vq_ema         EMA + renormalisation (:66-75).     vq_backward   grad of :77-78.
#include "common.cuh"
#include "kernels.h"

namespace fo {

// =============================================================================== prep
__global__ void vq_prep_kernel(const float* __restrict__ embed, int dim, int n_embed, __nv_bfloat16* __restrict__ e_split,
                               float* __restrict__ e_t, float* __restrict__ e_norm2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_embed) return;
  float s = 0.f;
  for (int d = 0; d < dim; ++d) {
    const float v = embed[(size_t)d * n_embed + k];
    s += v * v;
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
    e_split[(size_t)k * 2 * dim + d] = hi;
    e_split[(size_t)k * 2 * dim + dim + d] = lo;
    e_t[(size_t)k * dim + d] = v;
  }
  e_norm2[k] = s;
}
__global__ void vq_e2max_kernel(float* e_norm2, int n_embed) {
  // e_norm2[n_embed] = max_k |e_k|^2 (error-band scale for vq_assign)
  __shared__ float sm[32];
  float m = 0.f;
  for (int k = threadIdx.x; k < n_embed; k += blockDim.x) m = fmaxf(m, e_norm2[k]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, sm[i]);
    e_norm2[n_embed] = m;
  }
}

cudaError_t launch_vq_prep(const float* embed, int dim, int n_embed, void* e_split, float* e_t, float* e_norm2,
                           cudaStream_t st) {
  vq_prep_kernel<<<(n_embed + 127) / 128, 128, 0, st>>>(embed, dim, n_embed, (__nv_bfloat16*)e_split, e_t, e_norm2);
  vq_e2max_kernel<<<1, 256, 0, st>>>(e_norm2, n_embed);
  return cudaGetLastError();
}

// =============================================================================== assign (tcgen05)
constexpr int kVqThreads = 32 * 14;  // warp0 TMA(B), warp1 MMA, warps 2-5 loader/convert, warps 6-13 epilogue (two
                                     // column halves x four TMEM lane quarters)
constexpr int kVqNT = 256;           // codes per accumulator
constexpr int kVqBStages = 4;        // ring of [256 x 64] bf16 B tiles (32 KB each)
constexpr float kVqBand = 1.5e-4f;   // |error of dist_k| <= kVqBand * |x| * |e_k|  (2.5x the split-bf16 bound 2*3*2^-17)

struct VqAssignParams {
  const float* x;
  size_t rows;
  int dim, n_embed;
  int n_tiles;      // ceil(n_embed / 256)
  int kchunks;      // dim / 64
  int a_bufs;       // 1 or 2
  int row_tiles;    // ceil(rows / 128)
  const float* e_norm2;  // [n_embed + 1]
  long long* embed_ind;
  int* flag_count;
  int* flag_rows;
};

__global__ void __launch_bounds__(kVqThreads, 1)
vq_assign_kernel(const __grid_constant__ VqAssignParams p, const __grid_constant__ CUtensorMap map_e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_sub = 128 * 128;                          // one [128 x 64] bf16 swizzled sub-tile
  const int a_buf_bytes = 2 * p.kchunks * a_sub;        // hi chunks then lo chunks
  const int b_tile = kVqNT * 128;                       // [256 x 64] bf16
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.a_bufs * a_buf_bytes;
  float* sE2 = reinterpret_cast<float*>(sB + (size_t)kVqBStages * b_tile);  // [n_tiles*256]
  float* sEN = sE2 + p.n_tiles * kVqNT;                                      // [n_tiles*256]  |e_k|
  float* sX2 = sEN + p.n_tiles * kVqNT;                                      // [4][128] (epilogue may lag the loader by 3 tiles)
  float* sMerge = sX2 + 4 * 128;                                             // [2][128][4] column-half hand-over
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMerge + 2 * 128 * 4);
  uint64_t* b_full = bars;                    // [kVqBStages]
  uint64_t* b_empty = b_full + kVqBStages;    // [kVqBStages]
  uint64_t* a_full = b_empty + kVqBStages;    // [2]
  uint64_t* a_empty = a_full + 2;             // [2]
  uint64_t* t_full = a_empty + 2;             // [2]
  uint64_t* t_empty = t_full + 2;             // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) tma_prefetch_desc(&map_e);
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < kVqBStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_full[b], 128);
      mbar_init(&a_empty[b], 1);
      mbar_init(&t_full[b], 1);
      mbar_init(&t_empty[b], 8);
    }
    fence_mbar_init();
  }
  // |e|^2 to smem (padded codes get +inf so they never win)
  for (int k = threadIdx.x; k < p.n_tiles * kVqNT; k += blockDim.x) {
    sE2[k] = k < p.n_embed ? p.e_norm2[k] : __int_as_float(0x7f800000);
    sEN[k] = k < p.n_embed ? sqrtf(p.e_norm2[k]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // warp-uniform copy (the shuffle lets the compiler keep MMA operands in uniform registers)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    // ---------------------------------------------------------------- B producer (codebook tiles via TMA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x) {
        for (int nt = 0; nt < p.n_tiles; ++nt)
          for (int kc = 0; kc < p.kchunks; ++kc)
            for (int part = 0; part < 2; ++part) {  // 0: hi, 1: lo
              mbar_wait(&b_empty[stage], phase ^ 1);
              mbar_expect_tx(&b_full[stage], b_tile);
              tma_load_2d(sB + (size_t)stage * b_tile, &map_e, &b_full[stage], part * p.dim + kc * 64, nt * kVqNT);
              if (++stage == kVqBStages) { stage = 0; phase ^= 1; }
            }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, kVqNT, 0, 0);
    const uint64_t desc_base = make_smem_desc(0, 128, 16);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0, acc_it = 0;
    for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x, ++it) {
      const int ab = it % p.a_bufs;
      mbar_wait(&a_full[ab], (it / p.a_bufs) & 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + (size_t)ab * a_buf_bytes);
      for (int nt = 0; nt < p.n_tiles; ++nt, ++acc_it) {
        const int tb = acc_it & 1;
        mbar_wait(&t_empty[tb], ((acc_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + tb * kVqNT;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const uint64_t a_hi = desc_base + (((a_base + kc * a_sub) & 0x3FFFF) >> 4);
          const uint64_t a_lo = desc_base + (((a_base + (p.kchunks + kc) * a_sub) & 0x3FFFF) >> 4);
          // B hi tile: hi*hi and lo*hi
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t sb = desc_base + ((smem_u32(sB + (size_t)stage * b_tile) & 0x3FFFF) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_hi + 2 * kk, sb + 2 * kk, idesc, (kc | kk) != 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_lo + 2 * kk, sb + 2 * kk, idesc, 1);
            umma_commit(&b_empty[stage]);
          }
          __syncwarp();
          if (++stage == kVqBStages) { stage = 0; phase ^= 1; }
          // B lo tile: hi*lo
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t sb = desc_base + ((smem_u32(sB + (size_t)stage * b_tile) & 0x3FFFF) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_hi + 2 * kk, sb + 2 * kk, idesc, 1);
            umma_commit(&b_empty[stage]);
          }
          __syncwarp();
          if (++stage == kVqBStages) { stage = 0; phase ^= 1; }
        }
        if (lane == 0) umma_commit(&t_full[tb]);
        __syncwarp();
      }
      if (lane == 0) umma_commit(&a_empty[ab]);
      __syncwarp();
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- loader: fp32 rows -> bf16 hi/lo swizzled tiles
    const int t = threadIdx.x - 64;         // 0..127
    const int q4 = p.dim / 4;               // float4 per row
    const int rows_per_iter = 128 / q4;     // 8 (dim 64) or 4 (dim 128)
    int it = 0;
    for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x, ++it) {
      const int ab = it % p.a_bufs;
      mbar_wait(&a_empty[ab], ((it / p.a_bufs) & 1) ^ 1);
      uint8_t* a_base = sA + (size_t)ab * a_buf_bytes;
      const size_t row0 = (size_t)rt * 128;
      // all global loads of the tile are issued before the first use (16 x 16 B in flight per thread): the loop would
      // otherwise expose one DRAM latency per iteration
      for (int i0 = 0; i0 < q4; i0 += 16) {
        float4 vv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int idx = (i0 + u) * 128 + t;
          const int r = idx / q4, q = idx % q4;
          vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row0 + r < p.rows) vv[u] = __ldg(reinterpret_cast<const float4*>(p.x + (row0 + r) * p.dim) + q);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int idx = (i0 + u) * 128 + t;
          const int r = idx / q4, q = idx % q4;
          const float4 v = vv[u];
          // |x|^2 of the row: reduce over the q4 lanes that share it
          float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          for (int o = q4 >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 32);
          if (q == 0) sX2[(it & 3) * 128 + r] = s;
          const __nv_bfloat16 h0 = __float2bfloat16(v.x), h1 = __float2bfloat16(v.y), h2 = __float2bfloat16(v.z),
                              h3 = __float2bfloat16(v.w);
          const float l0 = v.x - __bfloat162float(h0), l1 = v.y - __bfloat162float(h1),
                      l2 = v.z - __bfloat162float(h2), l3 = v.w - __bfloat162float(h3);
          const int k = q * 4;
          const int kc = k >> 6, kin = k & 63;
          const uint32_t off = (uint32_t)r * 128 + ((((uint32_t)kin >> 3) ^ ((uint32_t)r & 7)) << 4) + (kin & 7) * 2;
          uint2 hv, lv;
          hv.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          hv.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
          lv.x = pack_bf16x2(l0, l1);
          lv.y = pack_bf16x2(l2, l3);
          *reinterpret_cast<uint2*>(a_base + kc * a_sub + off) = hv;
          *reinterpret_cast<uint2*>(a_base + (p.kchunks + kc) * a_sub + off) = lv;
        }
      }
      (void)rows_per_iter;
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(&a_full[ab]);
    }
  } else {
    // ---------------------------------------------------------------- epilogue: running top-2 argmin per row
    const int quarter = warp & 3;
    const int half = (warp - 6) >> 2;        // which 128 columns of every 256-column accumulator this warp scans
    const int row = quarter * 32 + lane;
    const float INF = __int_as_float(0x7f800000);
    int it = 0, acc_it = 0;
    for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x, ++it) {
      // best = smallest distance so far; other_lb = smallest LOWER bound among all other codes, where code k's
      // distance is only known to +- kVqBand * |x| * |e_k| (split-bf16 product error, scaled per code so that dead
      // codes with huge norms -- the EMA renormalisation blows unused codes up, reference :70-75 -- do not widen it)
      // Four independent running minima (columns j, j+4, ...): one chain would serialise 512 dependent compare/select
      // steps per row; they are merged after the last code tile.
      float best[4], berr[4], other[4];
      int bidx[4];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) { best[ch] = INF; berr[ch] = 0.f; other[ch] = INF; bidx[ch] = 0; }
      float x2 = 0.f, cx = 0.f;
      for (int nt = 0; nt < p.n_tiles; ++nt, ++acc_it) {
        const int tb = acc_it & 1;
        mbar_wait(&t_full[tb], (acc_it >> 1) & 1);
        tc_fence_after();
        if (nt == 0) {
          x2 = sX2[(it & 3) * 128 + row];   // written before a_full, which precedes t_full
          cx = kVqBand * sqrtf(x2);
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + tb * kVqNT;
        const float* e2 = sE2 + nt * kVqNT;
        const float* en = sEN + nt * kVqNT;
#pragma unroll 1
        for (int c = half * (kVqNT / 2); c < (half + 1) * (kVqNT / 2); c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int ch = j & 3;
            // same association as the reference: (|x|^2 - 2 x.e) + |e|^2   (:49-53)
            const float d = (x2 - 2.f * __uint_as_float(v[j])) + e2[c + j];
            const float err = cx * en[c + j];
            const bool lt = d < best[ch];   // strict: the lower index wins ties inside a chain
            // the loser of the comparison only contributes its lower bound
            other[ch] = fminf(other[ch], lt ? best[ch] - berr[ch] : d - err);
            best[ch] = lt ? d : best[ch];
            berr[ch] = lt ? err : berr[ch];
            bidx[ch] = lt ? nt * kVqNT + c + j : bidx[ch];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[tb]);
      }
      // merge the chains (ties -> lowest index, the reference's first-max rule :54)
      float bestv = best[0], best_err = berr[0], other_lb = other[0];
      int besti = bidx[0];
#pragma unroll
      for (int ch = 1; ch < 4; ++ch) {
        const bool lt = best[ch] < bestv || (best[ch] == bestv && bidx[ch] < besti);
        other_lb = fminf(other_lb, lt ? bestv - best_err : best[ch] - berr[ch]);
        other_lb = fminf(other_lb, other[ch]);
        bestv = lt ? best[ch] : bestv;
        best_err = lt ? berr[ch] : best_err;
        besti = lt ? bidx[ch] : besti;
      }
      // hand the upper column half over to the lower one (named barrier 1: the 256 epilogue threads)
      float* mg = sMerge + ((it & 1) * 128 + row) * 4;
      if (half == 1) {
        mg[0] = bestv; mg[1] = best_err; mg[2] = other_lb; mg[3] = __int_as_float(besti);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 1) continue;
      {
        const float b1 = mg[0], e1 = mg[1], o1 = mg[2];
        const int i1 = __float_as_int(mg[3]);
        const bool lt = b1 < bestv || (b1 == bestv && i1 < besti);   // ties -> lowest index
        other_lb = fminf(other_lb, lt ? bestv - best_err : b1 - e1);
        other_lb = fminf(other_lb, o1);
        bestv = lt ? b1 : bestv;
        best_err = lt ? e1 : best_err;
        besti = lt ? i1 : besti;
      }
      const size_t grow = (size_t)rt * 128 + row;
      if (grow < p.rows) {
        p.embed_ind[grow] = besti;
        if (!(other_lb > bestv + best_err)) {   // ambiguous within the error bound (also catches NaN)
          const int slot = atomicAdd(p.flag_count, 1);
          p.flag_rows[slot] = (int)grow;   // capacity = rows
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// exact re-evaluation of flagged rows: fp64 accumulation, first minimum wins (reference tie rule :54).
// One warp per row; each lane scans codes lane, lane+32, ... of the transposed codebook e_t [n_embed][dim] with 16-byte
// loads and four independent fp64 accumulators.
__global__ void vq_refine_kernel(const float* __restrict__ x, const float* __restrict__ e_t, int dim, int n_embed,
                                 const int* __restrict__ flag_count, const int* __restrict__ flag_rows,
                                 long long* __restrict__ embed_ind) {
  extern __shared__ float sx[];  // [warps][dim]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int total = *flag_count;
  float* myx = sx + warp * dim;
  for (int f = blockIdx.x * nwarps + warp; f < total; f += gridDim.x * nwarps) {
    const int row = flag_rows[f];
    __syncwarp();
    for (int d = lane; d < dim; d += 32) myx[d] = x[(size_t)row * dim + d];
    __syncwarp();
    double best = 1e300;
    int besti = 0x7fffffff;
    for (int k = lane; k < n_embed; k += 32) {
      const float4* er = reinterpret_cast<const float4*>(e_t + (size_t)k * dim);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      for (int q = 0; q < dim / 4; ++q) {
        const float4 e = __ldg(er + q);
        const float4 xv = *reinterpret_cast<const float4*>(myx + 4 * q);
        // e^2 - 2 x e, term by term (|x|^2 is common to all codes)
        a0 += (double)e.x * ((double)e.x - 2.0 * (double)xv.x);
        a1 += (double)e.y * ((double)e.y - 2.0 * (double)xv.y);
        a2 += (double)e.z * ((double)e.z - 2.0 * (double)xv.z);
        a3 += (double)e.w * ((double)e.w - 2.0 * (double)xv.w);
      }
      const double dist = (a0 + a1) + (a2 + a3);
      if (dist < best) { best = dist; besti = k; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) embed_ind[row] = besti;
  }
}

size_t vq_assign_smem_bytes(int dim, int n_embed) {
  const int kchunks = dim / 64;
  const int a_bufs = dim <= 64 ? 2 : 1;
  const int n_tiles = (n_embed + kVqNT - 1) / kVqNT;
  return (size_t)a_bufs * 2 * kchunks * 128 * 128 + (size_t)kVqBStages * kVqNT * 128 + (size_t)n_tiles * kVqNT * 8 +
         4 * 128 * 4 + 2 * 128 * 4 * 4 + (2 * kVqBStages + 8) * 8 + 16 + 1024;
}
size_t vq_assign_workspace_bytes(size_t rows, int dim) { return 256 + rows * sizeof(int); }

cudaError_t launch_vq_assign(const float* x, size_t rows, int dim, int n_embed, const float* e_t,
                             const void* e_split, const float* e_norm2, int64_t* embed_ind, int* n_flagged,
                             void* workspace, const CUtensorMap* map_e, int num_sms, cudaStream_t st) {
  (void)e_split;
  VqAssignParams p;
  p.x = x; p.rows = rows; p.dim = dim; p.n_embed = n_embed;
  p.n_tiles = (n_embed + kVqNT - 1) / kVqNT;
  p.kchunks = dim / 64;
  p.a_bufs = dim <= 64 ? 2 : 1;
  p.row_tiles = (int)((rows + 127) / 128);
  p.e_norm2 = e_norm2;
  p.embed_ind = reinterpret_cast<long long*>(embed_ind);
  p.flag_count = reinterpret_cast<int*>(workspace);
  p.flag_rows = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  cudaError_t e = cudaMemsetAsync(p.flag_count, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  const int grid = p.row_tiles < num_sms ? p.row_tiles : num_sms;
  vq_assign_kernel<<<grid, kVqThreads, vq_assign_smem_bytes(dim, n_embed), st>>>(p, *map_e);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int warps = 8;
  vq_refine_kernel<<<num_sms * 2, warps * 32, warps * dim * sizeof(float), st>>>(
      x, e_t, dim, n_embed, p.flag_count, p.flag_rows, reinterpret_cast<long long*>(embed_ind));
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (n_flagged != nullptr) e = cudaMemcpyAsync(n_flagged, p.flag_count, sizeof(int), cudaMemcpyDeviceToDevice, st);
  return e;
}
cudaError_t init_vq() {
  return cudaFuncSetAttribute(vq_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
}

// =============================================================================== gather + ST + loss + EMA stats
// LPR lanes share a row (float4 each).  Per-CTA statistics are privatised in shared memory
// ([n_embed][dim] fp32 + counts) when they fit, then flushed with one global atomic per entry.
template <bool SMEM_STATS>
__global__ void vq_gather_stats_kernel(const float* __restrict__ x, const long long* __restrict__ ind, size_t rows,
                                       int dim, int n_embed, const float* __restrict__ e_t, float* __restrict__ q_f32,
                                       __nv_bfloat16* __restrict__ q_bf16, float* __restrict__ diff_sum,
                                       float* __restrict__ counts, float* __restrict__ embed_sum) {
  extern __shared__ float sm[];  // SMEM_STATS: [n_embed*dim] sums, [n_embed] counts
  __shared__ float red[32];
  const bool stats = counts != nullptr;
  if (SMEM_STATS && stats) {
    for (int i = threadIdx.x; i < n_embed * (dim + 1); i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
  }
  float* s_sum = sm;
  float* s_cnt = sm + (size_t)n_embed * dim;
  const int q4 = dim / 4;
  const size_t total = rows * q4;
  float dacc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / q4;
    const int q = (int)(i % q4);
    const int k = (int)ind[r];
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + r * dim) + q);
    const float4 ev = __ldg(reinterpret_cast<const float4*>(e_t + (size_t)k * dim) + q);
    // (:77) diff = (quantize - input)^2 ; (:78) quantize = input + (quantize - input)
    const float t0 = ev.x - xv.x, t1 = ev.y - xv.y, t2 = ev.z - xv.z, t3 = ev.w - xv.w;
    dacc += t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3;
    const float4 qv = make_float4(xv.x + t0, xv.y + t1, xv.z + t2, xv.w + t3);
    if (q_f32 != nullptr) reinterpret_cast<float4*>(q_f32 + r * dim)[q] = qv;
    if (q_bf16 != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(qv.x, qv.y);
      o.y = pack_bf16x2(qv.z, qv.w);
      reinterpret_cast<uint2*>(q_bf16 + r * dim)[q] = o;
    }
    if (stats) {
      if (SMEM_STATS) {
        float* d = s_sum + (size_t)k * dim + q * 4;
        atomicAdd(d + 0, xv.x); atomicAdd(d + 1, xv.y); atomicAdd(d + 2, xv.z); atomicAdd(d + 3, xv.w);
        if (q == 0) atomicAdd(s_cnt + k, 1.f);
      } else {
        const int d0 = q * 4;
        atomicAdd(embed_sum + (size_t)(d0 + 0) * n_embed + k, xv.x);
        atomicAdd(embed_sum + (size_t)(d0 + 1) * n_embed + k, xv.y);
        atomicAdd(embed_sum + (size_t)(d0 + 2) * n_embed + k, xv.z);
        atomicAdd(embed_sum + (size_t)(d0 + 3) * n_embed + k, xv.w);
        if (q == 0) atomicAdd(counts + k, 1.f);
      }
    }
  }
  // block reduce of the commitment-loss partial
  dacc = warp_sum(dacc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(diff_sum, s);
  }
  if (SMEM_STATS && stats) {
    for (int i = threadIdx.x; i < n_embed * dim; i += blockDim.x) {
      const float v = s_sum[i];
      if (v != 0.f) {
        const int k = i / dim, d = i % dim;
        atomicAdd(embed_sum + (size_t)d * n_embed + k, v);
      }
    }
    for (int k = threadIdx.x; k < n_embed; k += blockDim.x) {
      const float v = s_cnt[k];
      if (v != 0.f) atomicAdd(counts + k, v);
    }
  }
}

cudaError_t launch_vq_gather_stats(const float* x, const int64_t* ind, size_t rows, int dim, int n_embed,
                                   const float* e_t, float* q_f32, void* q_bf16, float* diff_sum, float* counts,
                                   float* embed_sum, int num_sms, cudaStream_t st) {
  const size_t smem = (size_t)n_embed * (dim + 1) * sizeof(float);
  const size_t total = rows * (dim / 4);
  size_t blocks = (total + 255) / 256;
  if (smem <= 200 * 1024) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(vq_gather_stats_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kMaxDynSmem - 1024);  // the kernel also has 128 B of static smem
      if (e != cudaSuccess) return e;
      configured = true;
    }
    if (blocks > (size_t)num_sms) blocks = num_sms;  // one CTA per SM (smem-limited)
    vq_gather_stats_kernel<true><<<(int)blocks, 1024, smem, st>>>(x, (const long long*)ind, rows, dim, n_embed, e_t,
                                                                   q_f32, (__nv_bfloat16*)q_bf16, diff_sum, counts,
                                                                   embed_sum);
  } else {
    if (blocks > (size_t)num_sms * 8) blocks = (size_t)num_sms * 8;
    vq_gather_stats_kernel<false><<<(int)blocks, 256, 0, st>>>(x, (const long long*)ind, rows, dim, n_embed, e_t, q_f32,
                                                                (__nv_bfloat16*)q_bf16, diff_sum, counts, embed_sum);
  }
  return cudaGetLastError();
}

// =============================================================================== EMA (:66-75)
__global__ void vq_ema_kernel(float* __restrict__ embed, float* __restrict__ cluster_size, float* __restrict__ embed_avg,
                              const float* __restrict__ counts, const float* __restrict__ embed_sum, int dim,
                              int n_embed, float decay, float eps) {
  __shared__ float red[32];
  __shared__ float n_total;
  float part = 0.f;
  for (int k = threadIdx.x; k < n_embed; k += blockDim.x) {
    const float cs = cluster_size[k] * decay + counts[k] * (1.f - decay);
    cluster_size[k] = cs;
    part += cs;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    n_total = s;
  }
  __syncthreads();
  const float n = n_total;
  const float denom = n + n_embed * eps;
  for (int i = threadIdx.x; i < dim * n_embed; i += blockDim.x) {
    const int k = i % n_embed;
    const float ea = embed_avg[i] * decay + embed_sum[i] * (1.f - decay);
    embed_avg[i] = ea;
    const float cs = (cluster_size[k] + eps) / denom * n;
    embed[i] = ea / cs;
  }
}
cudaError_t launch_vq_ema(float* embed, float* cluster_size, float* embed_avg, const float* counts,
                          const float* embed_sum, int dim, int n_embed, float decay, float eps, cudaStream_t st) {
  vq_ema_kernel<<<1, 1024, 0, st>>>(embed, cluster_size, embed_avg, counts, embed_sum, dim, n_embed, decay, eps);
  return cudaGetLastError();
}

// =============================================================================== backward of :77-78
__global__ void vq_backward_kernel(const void* __restrict__ g_q, int g_is_bf16, int g_cs, int g_c_off,
                                   const float* __restrict__ g_diff, const float* __restrict__ x,
                                   const long long* __restrict__ ind, const float* __restrict__ e_t, size_t rows,
                                   int dim, float* __restrict__ gx_f32, __nv_bfloat16* __restrict__ gx_bf16) {
  const int q4 = dim / 4;
  const size_t total = rows * q4;
  const float scale = g_diff != nullptr ? (*g_diff) * 2.f / (float)((double)rows * dim) : 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / q4;
    const int q = (int)(i % q4);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g_q != nullptr) {
      if (g_is_bf16) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g_q) + r * g_cs +
                                                             g_c_off) + q);
        g = make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y));
      } else {
        g = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g_q) + r * g_cs + g_c_off) + q);
      }
    }
    if (scale != 0.f) {
      const int k = (int)ind[r];
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + r * dim) + q);
      const float4 ev = __ldg(reinterpret_cast<const float4*>(e_t + (size_t)k * dim) + q);
      g.x += scale * (xv.x - ev.x); g.y += scale * (xv.y - ev.y);
      g.z += scale * (xv.z - ev.z); g.w += scale * (xv.w - ev.w);
    }
    if (gx_f32 != nullptr) reinterpret_cast<float4*>(gx_f32 + r * dim)[q] = g;
    if (gx_bf16 != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(g.x, g.y);
      o.y = pack_bf16x2(g.z, g.w);
      reinterpret_cast<uint2*>(gx_bf16 + r * dim)[q] = o;
    }
  }
}
cudaError_t launch_vq_backward(const void* g_q, int g_q_is_bf16, int g_cs, int g_c_off, const float* g_diff,
                               const float* x, const int64_t* ind, const float* e_t, size_t rows, int dim,
                               int n_embed, float* gx_f32, void* gx_bf16, int num_sms, cudaStream_t st) {
  (void)n_embed;
  const size_t total = rows * (dim / 4);
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)num_sms * 8) blocks = (size_t)num_sms * 8;
  if (blocks < 1) blocks = 1;
  vq_backward_kernel<<<(int)blocks, 256, 0, st>>>(g_q, g_q_is_bf16, g_cs, g_c_off, g_diff, x, (const long long*)ind,
                                                   e_t, rows, dim, gx_f32, (__nv_bfloat16*)gx_bf16);
  return cudaGetLastError();
}

}  // namespace fo
