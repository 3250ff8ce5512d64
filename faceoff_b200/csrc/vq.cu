// Vector quantiser kernels (reference models/vqvae_conv3d_latent.py:33-83).
//
//  vq_prep        codebook fp32 [dim, n_embed] -> bf16 hi/lo split [n_embed, 2*dim] (K-major GEMM B operand) followed by
//                 the augmented K slice [n_pad, 16] (-|e|^2/2 as three bf16 terms, |e| rounded up), transposed fp32 copy
//                 [n_embed, dim] (gather operand) and |e|^2.
//  vq_assign      nearest code per row (:48-54).  The distance matrix is a dense contraction, so it runs on
//                 tcgen05: A = x split into bf16 hi + lo on the fly (fp32 rows are loaded coalesced, split in
//                 registers, written into 128B-swizzled smem tiles), B = codebook hi/lo tiles streamed by TMA,
//                 three MMAs per K block (hi*hi + lo*hi + hi*lo ~ 2^-17 relative error) plus one K=16 MMA per code
//                 tile that adds -|e|^2/2 and half the error band, accumulators double-buffered in TMEM.  The epilogue
//                 keeps the two largest scores and the arg-best per row in registers -- the [rows, n_embed] distance
//                 matrix never exists in HBM.  Rows whose top-2 gap is inside the error band are re-evaluated exactly
//                 (fp32 filter, then fp64 accumulation of the fp32 data) by vq_refine, which makes embed_ind bit-exact
//                 w.r.t. the reference outside true near-ties.
//  vq_gather_stats gather + straight-through + commitment loss + EMA statistics in one pass (:55-61,77-78).
//  vq_ema         EMA + renormalisation (:66-75).     vq_backward   grad of :77-78.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace fo {

// =============================================================================== prep
constexpr int kVqNTc = 256;   // codes per accumulator tile (= kVqNT below)
size_t vq_split_elems(int dim, int n_embed) {
  const size_t n_pad = (size_t)(n_embed + kVqNTc - 1) / kVqNTc * kVqNTc;
  return (size_t)n_embed * 2 * dim + n_pad * 16;
}
// e_split holds the split codebook [n_embed][2*dim] followed by the AUGMENTED K slice [n_pad][16] (n_pad = n_embed rounded
// up to 256): per code {-h, -m, -l, en_up, 0 x 12} with |e_k|^2 / 2 = h + m + l (three bf16 terms: fp32-exact) and en_up =
// |e_k| rounded up to bf16.  Against the row vector {1, 1, 1, cx/2, 0...} one extra K=16 MMA turns the accumulator into
// x.e_k - |e_k|^2/2 + (cx/2)|e_k| = -(lower bound of the distance)/2, so the epilogue needs no per-code arithmetic.
// Pad codes get -1e30: they can never win.
__global__ void vq_prep_kernel(const float* __restrict__ embed, int dim, int n_embed, int n_pad,
                               __nv_bfloat16* __restrict__ e_split, float* __restrict__ e_t, float* __restrict__ e_norm2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pad) return;
  __nv_bfloat16* aug = e_split + (size_t)n_embed * 2 * dim + (size_t)k * 16;
  if (k >= n_embed) {
    aug[0] = __float2bfloat16(-1e30f);
    for (int j = 1; j < 16; ++j) aug[j] = __float2bfloat16(0.f);
    return;
  }
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
  const float hv = -0.5f * s;
  const __nv_bfloat16 h = __float2bfloat16(hv);
  const float r1 = hv - __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16(r1);
  const __nv_bfloat16 l = __float2bfloat16(r1 - __bfloat162float(m));
  aug[0] = h; aug[1] = m; aug[2] = l;
  aug[3] = __float2bfloat16_ru(sqrtf(s));
  for (int j = 4; j < 16; ++j) aug[j] = __float2bfloat16(0.f);
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
  const int n_pad = (n_embed + kVqNTc - 1) / kVqNTc * kVqNTc;
  vq_prep_kernel<<<(n_pad + 127) / 128, 128, 0, st>>>(embed, dim, n_embed, n_pad, (__nv_bfloat16*)e_split, e_t, e_norm2);
  vq_e2max_kernel<<<1, 256, 0, st>>>(e_norm2, n_embed);
  return cudaGetLastError();
}

// =============================================================================== assign (tcgen05)
constexpr int kVqThreads = 32 * 14;  // warp0 TMA(B), warp1 MMA, warps 2-5 loader/convert, warps 6-13 epilogue (two
                                     // column halves x four TMEM lane quarters)
constexpr int kVqNT = 256;           // codes per accumulator
constexpr int kVqBStages = 3;        // ring of [256 x 64] bf16 B tiles (32 KB each)
constexpr float kVqBand = 1.5e-4f;   // |error of dist_k| <= kVqBand * |x| * |e_k|  (2.5x the split-bf16 bound 2*3*2^-17)
constexpr float kVqBig = 1e38f;      // -kVqBig = "no candidate yet" (finite: the index tag must not make a NaN)

struct VqAssignParams {
  const float* x;
  size_t rows;
  int dim, n_embed;
  int n_tiles;      // ceil(n_embed / 256)
  int a_bufs;       // 1 or 2
  int row_tiles;    // ceil(rows / 128)
  const float* e_norm2;  // [n_embed + 1]
  long long* embed_ind;
  int* flag_count;
  int* flag_rows;
  float* flag_u;    // per flagged row: certain upper bound of the winner's |e|^2 - 2 x.e
};

// tcgen05.wait::ld that also names the destination registers, so that no consumer can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),
                 "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),
                 "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// One halving step of a transposed butterfly reduction: N values per lane, lanes `off` apart exchange the half they do
// not keep.  After the steps off = 8N/16 ... the lane whose index bits select value u holds its full sum in s[0].
template <int N>
__device__ __forceinline__ void vq_halve(float (&s)[16], int lane, int off) {
  const bool upper = (lane & off) != 0;
#pragma unroll
  for (int j = 0; j < N / 2; ++j) {
    const float keep = upper ? s[j + N / 2] : s[j];
    const float send = upper ? s[j] : s[j + N / 2];
    s[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
  }
}

// bf16-rounded-up half band of a row: cx/2 with cx = kVqBand * |x| (the row's entry of the augmented K slice)
__device__ __forceinline__ __nv_bfloat16 vq_half_band(float x2) { return __float2bfloat16_ru(0.5f * kVqBand * sqrtf(x2)); }

// Scan 32 accumulator columns of one row.  Thanks to the augmented K slice the accumulator IS v_k = -(lower bound of code
// k's distance)/2, so no per-code arithmetic is needed: what remains is finding the largest v, its column, and whether any
// OTHER code lies within the winner's error band.  FMNMX and LOP3 share the ALU pipe (one warp instruction per two cycles
// per scheduler), and a running top-2 costs ~3.5 such operations per code -- with two epilogue warps per scheduler that
// was MORE than the tensor time of a code tile.  So the runner-up is not tracked: per code
//   ALU pipe : tag the column into 5 mantissa bits (LOP3) and a 3-input max tree (0.5 FMNMX3)            ~1.5 ops
//   FMA pipe : count the codes above thr = v_max - band(max) with a saturated FMA (exactly 0 or 1) + add     2 ops
// The two pipes issue in parallel.  When the running maximum moves by more than the new winner's band the count restarts
// (every earlier code is then certainly out of reach); when it moves by less, the old maximum stays counted and the
// row ends up flagged -- which is exactly right.  count == 1 <=> the winner is certain.
constexpr float kVqCountScale = 1.152921504606847e18f;   // 2^60: (t - thr) * 2^60 saturates to exactly 0 or 1
constexpr float kVqTagSlack = 1.6e-5f;                   // 4x the value shift of two index tags (2 * 2^-19 relative)
__device__ __forceinline__ float vq_max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ float vq_threshold(float m1, float band1) { return fmaf(-kVqTagSlack, fabsf(m1), m1 - band1); }
__device__ __forceinline__ void vq_scan32(const uint32_t (&v)[32], uint32_t tag_mask, float& m1, int& midx, float& cnt,
                                          float cxp, const float* sEN, int code0) {
  float t[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) t[j] = __uint_as_float((v[j] & tag_mask) | (uint32_t)j);   // ONE LOP3 each (mask in a register)
  float a[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = vq_max3(t[3 * i], t[3 * i + 1], t[3 * i + 2]);
  a[10] = fmaxf(t[30], t[31]);
  const float m = fmaxf(vq_max3(vq_max3(a[0], a[1], a[2]), vq_max3(a[3], a[4], a[5]), vq_max3(a[6], a[7], a[8])),
                        fmaxf(a[9], a[10]));
  const bool changed = m > m1;
  const float old = m1;
  m1 = fmaxf(m1, m);
  midx = changed ? code0 + (int)(__float_as_uint(m) & 31u) : midx;
  const float thr = vq_threshold(m1, cxp * sEN[midx]);
  cnt = (changed && old < thr) ? 0.f : cnt;
  const float thr_s = -thr * kVqCountScale;
  float c0 = 0.f, c1 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    c0 += __saturatef(fmaf(t[j], kVqCountScale, thr_s));
    c1 += __saturatef(fmaf(t[j + 1], kVqCountScale, thr_s));
  }
  cnt += c0 + c1;
}

template <int DIM>
__global__ void __launch_bounds__(kVqThreads, 1)   // 128 registers: 4 warps of a 16 K-register SM sub-partition
vq_assign_kernel(const __grid_constant__ VqAssignParams p, const __grid_constant__ CUtensorMap map_e,
                 const __grid_constant__ CUtensorMap map_x) {
  constexpr int KCH = DIM / 64;                         // 64-wide K chunks
  constexpr int Q4 = DIM / 4;                           // float4 (= loader lanes) per row: 16 or 32
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic instead of a pointer round-trip keeps the shared address space known: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int a_sub = 128 * 128;                      // one [128 x 64] bf16 swizzled sub-tile
  constexpr int a_buf_bytes = 2 * KCH * a_sub;          // hi chunks then lo chunks
  constexpr int b_tile = kVqNT * 128;                   // [256 x 64] bf16
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.a_bufs * a_buf_bytes;
  constexpr int ax_tile = 128 * 32;                     // augmented slice of the rows: [128 x 16] bf16, 32B swizzle
  constexpr int bx_tile = kVqNT * 32;                   // augmented slice of a code tile: [256 x 16] bf16
  uint8_t* sAx = sB + (size_t)kVqBStages * b_tile;      // [a_bufs]
  uint8_t* sBx = sAx + (size_t)p.a_bufs * ax_tile;      // [2]
  float* sEN = reinterpret_cast<float*>(sBx + 2 * bx_tile);                  // [n_tiles*256]  |e_k| rounded up to bf16
  float* sX2 = sEN + p.n_tiles * kVqNT;                                      // [4][128] (epilogue may lag the loader by 3 tiles)
  float* sMerge = sX2 + 4 * 128;                                             // [2][128][4] column-half hand-over
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMerge + 2 * 128 * 4);
  uint64_t* b_full = bars;                    // [kVqBStages]
  uint64_t* b_empty = b_full + kVqBStages;    // [kVqBStages]
  uint64_t* bx_full = b_empty + kVqBStages;   // [2]
  uint64_t* bx_empty = bx_full + 2;           // [2]
  uint64_t* a_full = bx_empty + 2;            // [2]
  uint64_t* a_empty = a_full + 2;             // [2]
  uint64_t* t_full = a_empty + 2;             // [2]
  uint64_t* t_empty = t_full + 2;             // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&map_e);
      tma_prefetch_desc(&map_x);
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < kVqBStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bx_full[b], 1);
      mbar_init(&bx_empty[b], 1);
      mbar_init(&a_full[b], 128);
      mbar_init(&a_empty[b], 1);
      mbar_init(&t_full[b], 1);
      mbar_init(&t_empty[b], 8);
    }
    fence_mbar_init();
  }
  // |e_k| exactly as the augmented slice holds it (rounded up to bf16): the winner's band is recomputed from it
  for (int k = threadIdx.x; k < p.n_tiles * kVqNT; k += blockDim.x)
    sEN[k] = k < p.n_embed ? __bfloat162float(__float2bfloat16_ru(sqrtf(p.e_norm2[k]))) : 0.f;
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
      int xit = 0;
      for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x) {
        for (int nt = 0; nt < p.n_tiles; ++nt, ++xit) {
          {   // augmented slice of this code tile
            const int xs = xit & 1;
            mbar_wait_backoff(&bx_empty[xs], ((xit >> 1) & 1) ^ 1, 64);
            mbar_expect_tx(&bx_full[xs], bx_tile);
            tma_load_2d(sBx + xs * bx_tile, &map_x, &bx_full[xs], 0, nt * kVqNT);
          }
          for (int kc = 0; kc < KCH; ++kc)
            for (int part = 0; part < 2; ++part) {  // 0: hi, 1: lo
              mbar_wait_backoff(&b_empty[stage], phase ^ 1, 64);
              mbar_expect_tx(&b_full[stage], b_tile);
              tma_load_2d(sB + (size_t)stage * b_tile, &map_e, &b_full[stage], part * DIM + kc * 64, nt * kVqNT);
              if (++stage == kVqBStages) { stage = 0; phase ^= 1; }
            }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, kVqNT, 0, 0);
    const uint64_t desc_base = make_smem_desc(0, 128, 16);
    const uint64_t desc32_base = make_smem_desc(0, 32, 16);   // the K=16 augmented slice: 32-byte rows, 32B swizzle
    int stage = 0;
    uint32_t phase = 0;
    int it = 0, acc_it = 0;
    for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x, ++it) {
      const int ab = it % p.a_bufs;
      mbar_wait_backoff(&a_full[ab], (it / p.a_bufs) & 1, 32);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + (size_t)ab * a_buf_bytes);
      for (int nt = 0; nt < p.n_tiles; ++nt, ++acc_it) {
        const int tb = acc_it & 1;
        mbar_wait_backoff(&t_empty[tb], ((acc_it >> 1) & 1) ^ 1, 32);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + tb * kVqNT;
#pragma unroll
        for (int kc = 0; kc < KCH; ++kc) {
          const uint64_t a_hi = desc_base + (((a_base + kc * a_sub) & 0x3FFFF) >> 4);
          const uint64_t a_lo = desc_base + (((a_base + (KCH + kc) * a_sub) & 0x3FFFF) >> 4);
          // B hi tile: hi*hi and lo*hi
          mbar_wait_backoff(&b_full[stage], phase, 20);
          tc_fence_after();
          if (elect_one()) {
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
          mbar_wait_backoff(&b_full[stage], phase, 20);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t sb = desc_base + ((smem_u32(sB + (size_t)stage * b_tile) & 0x3FFFF) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_hi + 2 * kk, sb + 2 * kk, idesc, 1);
            umma_commit(&b_empty[stage]);
          }
          __syncwarp();
          if (++stage == kVqBStages) { stage = 0; phase ^= 1; }
        }
        {   // augmented K slice: + (-|e|^2/2 + (cx/2)|e|)
          const int xs = acc_it & 1;
          mbar_wait_backoff(&bx_full[xs], (acc_it >> 1) & 1, 20);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ax = desc32_base + ((smem_u32(sAx + (size_t)ab * ax_tile) & 0x3FFFF) >> 4);
            const uint64_t bxd = desc32_base + ((smem_u32(sBx + xs * bx_tile) & 0x3FFFF) >> 4);
            umma_bf16(d_tmem, ax, bxd, idesc, 1);
            umma_commit(&bx_empty[xs]);
          }
          __syncwarp();
        }
        if (lane == 0) umma_commit(&t_full[tb]);
        __syncwarp();
      }
      if (lane == 0) umma_commit(&a_empty[ab]);
      __syncwarp();
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- loader: fp32 rows -> bf16 hi/lo swizzled tiles
    // Q4 consecutive lanes read one row (coalesced 16-byte loads); a thread's 16 loads of a batch belong to 16 rows.
    const int t = threadIdx.x - 64;         // 0..127
    const int q = t & (Q4 - 1);             // float4 column of this lane
    const int rsub = t / Q4;                // row within the 128/Q4 rows one load instruction covers
    constexpr int RPI = 128 / Q4;           // rows per load instruction: 8 or 4
    const int kc = (q * 4) >> 6, kin = (q * 4) & 63;
    int it = 0;
    for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x, ++it) {
      const int ab = it % p.a_bufs;
      const size_t row0 = (size_t)rt * 128;
      uint8_t* a_base = sA + (size_t)ab * a_buf_bytes;
      bool waited = false;
#pragma unroll 1
      for (int i0 = 0; i0 < Q4; i0 += 16) {
        // all 16 loads are in flight before the first use; they do not depend on the smem buffer being free
        float4 vv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int r = (i0 + u) * RPI + rsub;
          vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row0 + r < p.rows) vv[u] = __ldg(reinterpret_cast<const float4*>(p.x + (row0 + r) * DIM) + q);
        }
        if (!waited) {
          mbar_wait_backoff(&a_empty[ab], ((it / p.a_bufs) & 1) ^ 1, 64);
          waited = true;
        }
        float s[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int r = (i0 + u) * RPI + rsub;
          const float4 v = vv[u];
          s[u] = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
          // hi = bf16(v) (packed pair conversion), lo = bf16(v - hi)
          uint2 hv, lv;
          hv.x = pack_bf16x2(v.x, v.y);
          hv.y = pack_bf16x2(v.z, v.w);
          lv.x = pack_bf16x2(v.x - __uint_as_float(hv.x << 16), v.y - __uint_as_float(hv.x & 0xFFFF0000u));
          lv.y = pack_bf16x2(v.z - __uint_as_float(hv.y << 16), v.w - __uint_as_float(hv.y & 0xFFFF0000u));
          const uint32_t off = (uint32_t)r * 128 + ((((uint32_t)kin >> 3) ^ ((uint32_t)r & 7)) << 4) + (kin & 7) * 2;
          *reinterpret_cast<uint2*>(a_base + kc * a_sub + off) = hv;
          *reinterpret_cast<uint2*>(a_base + (KCH + kc) * a_sub + off) = lv;
        }
        // |x|^2 of the 16 rows: transposed butterfly over the Q4 lanes of a row (15 or 16 shuffles instead of 64+)
        int xr = -1;   // row whose |x|^2 this lane ends up holding
        if (Q4 == 32) {
          vq_halve<16>(s, lane, 16); vq_halve<8>(s, lane, 8); vq_halve<4>(s, lane, 4); vq_halve<2>(s, lane, 2);
          s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
          if ((q & 1) == 0) xr = (i0 + (q >> 1)) * RPI + rsub;
        } else {
          vq_halve<16>(s, lane, 8); vq_halve<8>(s, lane, 4); vq_halve<4>(s, lane, 2); vq_halve<2>(s, lane, 1);
          xr = (i0 + q) * RPI + rsub;
        }
        if (xr >= 0) {
          sX2[(it & 3) * 128 + xr] = s[0];
          // augmented slice of the row: {1, 1, 1, cx/2, 0 x 12} as a 32-byte row, 16-byte chunks swizzled by row bit 2
          const uint32_t hb = (uint32_t)__bfloat16_as_ushort(vq_half_band(s[0]));
          uint8_t* rowp = sAx + (size_t)ab * ax_tile + xr * 32;
          const int c0 = (xr >> 2) & 1;
          *reinterpret_cast<uint4*>(rowp + c0 * 16) = make_uint4(0x3F803F80u, 0x3F80u | (hb << 16), 0u, 0u);
          *reinterpret_cast<uint4*>(rowp + (c0 ^ 1) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(&a_full[ab]);
    }
  } else {
    // ---------------------------------------------------------------- epilogue: two largest scores per row
    // The accumulator holds v_k = x.e_k - |e_k|^2/2 + band_k/2 with band_k = cx' * en'_k >= kVqBand * |x| * |e_k| (both
    // factors rounded up to bf16; the split-bf16 product error of 2 x.e_k stays below it; it is scaled per code so that
    // dead codes with huge norms -- the EMA renormalisation blows unused codes up, reference :70-75 -- do not widen it).
    // -2 v_k is a LOWER bound of the distance (|x|^2 dropped) and -2 v_k + 2 band_k an upper bound, so the code k1 with
    // the largest v is the certain winner iff v_k1 - v_k2 > band_k1 for the runner-up k2; every other row goes to the
    // exact re-check together with U = -2 v_k1 + 2 band_k1.
    const int quarter = warp & 3;
    const int half = (warp - 6) >> 2;        // which 128 columns of every 256-column accumulator this warp scans
    const int row = quarter * 32 + lane;
    // 0xffffffe0, derived from a launch parameter so that it is not constant-folded (see vq_scan32)
    const uint32_t tag_mask = 0xffffffe0u | ((uint32_t)p.n_tiles >> 30);
    int it = 0, acc_it = 0;
    for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x, ++it) {
      float m1 = -kVqBig, cnt = 0.f;
      int ridx = 0;
      // written before a_full, which precedes t_full; same rounding as the loader's augmented row
      float cxp = 0.f;
      for (int nt = 0; nt < p.n_tiles; ++nt, ++acc_it) {
        const int tb = acc_it & 1;
        mbar_wait(&t_full[tb], (acc_it >> 1) & 1);
        tc_fence_after();
        if (nt == 0) cxp = 2.f * __bfloat162float(vq_half_band(sX2[(it & 3) * 128 + row]));
        const int c0 = half * (kVqNT / 2);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + tb * kVqNT + c0;
        const int code0 = nt * kVqNT + c0;
        // software pipeline over the four 32-column chunks: the next TMEM read is in flight while one is scanned
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
        tmem_ld_wait32(va);
        tmem_ld32(taddr + 32, vb);
        vq_scan32(va, tag_mask, m1, ridx, cnt, cxp, sEN, code0);
        tmem_ld_wait32(vb);
        tmem_ld32(taddr + 64, va);
        vq_scan32(vb, tag_mask, m1, ridx, cnt, cxp, sEN, code0 + 32);
        tmem_ld_wait32(va);
        tmem_ld32(taddr + 96, vb);
        vq_scan32(va, tag_mask, m1, ridx, cnt, cxp, sEN, code0 + 64);
        tmem_ld_wait32(vb);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[tb]);   // the accumulator is in registers: release it before the last scan
        vq_scan32(vb, tag_mask, m1, ridx, cnt, cxp, sEN, code0 + 96);
      }
      // hand the upper column half over to the lower one.  The upper half only ARRIVES on named barrier 1 (256 = the
      // epilogue threads) and goes on with the next row tile; sMerge is double-buffered by row-tile parity, and a warp
      // cannot be two row tiles ahead of its partner (both must release every accumulator before it is reused).
      float* mg = sMerge + ((it & 1) * 128 + row) * 4;
      if (half == 1) {
        mg[0] = m1; mg[1] = cnt; mg[2] = __int_as_float(ridx);
        asm volatile("bar.arrive 1, 256;" ::: "memory");
        continue;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float lose;   // the best value of the half that did not win
      {
        const float o1 = mg[0], ocnt = mg[1];
        const int oi = __float_as_int(mg[2]);
        const bool gt = o1 > m1;
        lose = fminf(o1, m1);
        cnt = gt ? ocnt : cnt;
        ridx = gt ? oi : ridx;
        m1 = fmaxf(m1, o1);
      }
      const size_t grow = (size_t)rt * 128 + row;
      if (grow < p.rows) {
        p.embed_ind[grow] = ridx;
        const float band1 = cxp * sEN[ridx];
        const float thr = vq_threshold(m1, band1);
        // certain iff the winner is the only code above its threshold in its own half and the other half stays below it
        // (the negations also catch NaN rows: every comparison fails, the count is 0)
        if (!(cnt == 1.f) || !(lose < thr)) {
          const int slot = atomicAdd(p.flag_count, 1);
          p.flag_rows[slot] = (int)grow;   // capacity = rows
          p.flag_u[slot] = -2.f * m1 + 2.f * band1 + 2.f * kVqTagSlack * fabsf(m1);
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
// A 128-thread CTA takes kVqRefineRows flagged rows at a time and streams the transposed codebook e_t [n_embed][dim]
// once for all of them (thread t scans codes t, t+128, ...).  Each code is first scored in fp32; only codes whose fp32
// score can still reach the row's upper bound U (left by the assign kernel: the certain upper bound of the winner's
// distance) are evaluated in fp64 -- typically two or three per row.
constexpr int kVqRefineThreads = 128;
constexpr int kVqRefineRows = 8;

template <int DIM>
__device__ __forceinline__ double vq_exact_score(const float* __restrict__ er, const float* xs) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
  for (int q = 0; q < DIM / 4; ++q) {
    const float4 e = __ldg(reinterpret_cast<const float4*>(er) + q);
    const float4 xv = *reinterpret_cast<const float4*>(xs + 4 * q);
    // e^2 - 2 x e, term by term (|x|^2 is common to all codes)
    a0 += (double)e.x * ((double)e.x - 2.0 * (double)xv.x);
    a1 += (double)e.y * ((double)e.y - 2.0 * (double)xv.y);
    a2 += (double)e.z * ((double)e.z - 2.0 * (double)xv.z);
    a3 += (double)e.w * ((double)e.w - 2.0 * (double)xv.w);
  }
  return (a0 + a1) + (a2 + a3);
}

template <int DIM>
__global__ void __launch_bounds__(kVqRefineThreads)
vq_refine_kernel(const float* __restrict__ x, const float* __restrict__ e_t, const float* __restrict__ e_norm2,
                 int n_embed, const int* __restrict__ flag_count, const int* __restrict__ flag_rows,
                 const float* __restrict__ flag_u, long long* __restrict__ embed_ind) {
  constexpr int R = kVqRefineRows;
  __shared__ __align__(16) float sx[R][DIM];
  __shared__ float s_u[R], s_xn[R];
  __shared__ int s_row[R];
  __shared__ double s_best[R][kVqRefineThreads / 32];
  __shared__ int s_idx[R][kVqRefineThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = *flag_count;
  // fp32 score error: (DIM + 4) roundings of relative size 2^-24 on terms bounded by |e|^2 + 2|x||e|  (x2 margin)
  const float ctol = 2.f * (float)(DIM + 4) * 5.9604645e-8f;
  for (int f0 = blockIdx.x * R; f0 < total; f0 += gridDim.x * R) {
    __syncthreads();   // previous group fully consumed
    for (int i = threadIdx.x; i < R * (DIM / 4); i += kVqRefineThreads) {
      const int r = i / (DIM / 4), q = i % (DIM / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f0 + r < total) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)flag_rows[f0 + r] * DIM) + q);
      reinterpret_cast<float4*>(&sx[r][0])[q] = v;
    }
    __syncthreads();
    if (threadIdx.x < R) {
      const int r = threadIdx.x;
      float n2 = 0.f;
      for (int d = 0; d < DIM; ++d) n2 = fmaf(sx[r][d], sx[r][d], n2);
      s_xn[r] = sqrtf(n2) * 1.0001f;
      const bool live = f0 + r < total;
      s_row[r] = live ? flag_rows[f0 + r] : -1;
      const float u = live ? flag_u[f0 + r] : -__int_as_float(0x7f800000);
      s_u[r] = (u == u) ? u : __int_as_float(0x7f800000);   // a NaN bound (NaN row) checks every code
    }
    __syncthreads();
    double best[R];
    int besti[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { best[r] = 1e300; besti[r] = 0x7fffffff; }
    for (int k = threadIdx.x; k < n_embed; k += kVqRefineThreads) {
      const float4* er = reinterpret_cast<const float4*>(e_t + (size_t)k * DIM);
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll 1
      for (int q0 = 0; q0 < DIM / 4; q0 += 16) {
        float4 ev[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) ev[u] = __ldg(er + q0 + u);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(&sx[r][4 * (q0 + u)]);
            acc[r] = fmaf(ev[u].x, xv.x, acc[r]);
            acc[r] = fmaf(ev[u].y, xv.y, acc[r]);
            acc[r] = fmaf(ev[u].z, xv.z, acc[r]);
            acc[r] = fmaf(ev[u].w, xv.w, acc[r]);
          }
        }
      }
      const float e2 = __ldg(e_norm2 + k);
      const float en = sqrtf(e2);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float d32 = fmaf(-2.f, acc[r], e2);
        const float tol = ctol * fmaf(2.f * s_xn[r], en, e2);
        if (d32 - tol <= s_u[r]) {
          const double dist = vq_exact_score<DIM>(reinterpret_cast<const float*>(er), &sx[r][0]);
          if (dist < best[r]) { best[r] = dist; besti[r] = k; }   // k ascends: the first minimum is kept
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double bv = best[r];
      int bi = besti[r];
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < bv || (ob == bv && oi < bi)) { bv = ob; bi = oi; }
      }
      if (lane == 0) { s_best[r][warp] = bv; s_idx[r][warp] = bi; }
    }
    __syncthreads();
    if (threadIdx.x < R && s_row[threadIdx.x] >= 0) {
      const int r = threadIdx.x;
      double bv = s_best[r][0];
      int bi = s_idx[r][0];
#pragma unroll
      for (int w = 1; w < kVqRefineThreads / 32; ++w)
        if (s_best[r][w] < bv || (s_best[r][w] == bv && s_idx[r][w] < bi)) { bv = s_best[r][w]; bi = s_idx[r][w]; }
      embed_ind[s_row[r]] = bi == 0x7fffffff ? 0 : bi;
    }
  }
}

size_t vq_assign_smem_bytes(int dim, int n_embed) {
  const int kchunks = dim / 64;
  const int a_bufs = dim <= 64 ? 2 : 1;
  const int n_tiles = (n_embed + kVqNT - 1) / kVqNT;
  return (size_t)a_bufs * 2 * kchunks * 128 * 128 + (size_t)kVqBStages * kVqNT * 128 + (size_t)a_bufs * 128 * 32 +
         2 * (size_t)kVqNT * 32 + (size_t)n_tiles * kVqNT * 4 + 4 * 128 * 4 + 2 * 128 * 4 * 4 +
         (2 * kVqBStages + 12) * 8 + 16 + 1024;
}
size_t vq_assign_workspace_bytes(size_t rows, int dim) { return 256 + rows * (sizeof(int) + sizeof(float)); }

cudaError_t launch_vq_assign(const float* x, size_t rows, int dim, int n_embed, const float* e_t,
                             const void* e_split, const float* e_norm2, int64_t* embed_ind, int* n_flagged,
                             void* workspace, const CUtensorMap* map_e, const CUtensorMap* map_x, int num_sms,
                             cudaStream_t st) {
  (void)e_split;
  VqAssignParams p;
  p.x = x; p.rows = rows; p.dim = dim; p.n_embed = n_embed;
  p.n_tiles = (n_embed + kVqNT - 1) / kVqNT;
  p.a_bufs = dim <= 64 ? 2 : 1;
  p.row_tiles = (int)((rows + 127) / 128);
  p.e_norm2 = e_norm2;
  p.embed_ind = reinterpret_cast<long long*>(embed_ind);
  p.flag_count = reinterpret_cast<int*>(workspace);
  p.flag_rows = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  p.flag_u = reinterpret_cast<float*>(p.flag_rows + rows);
  cudaError_t e = cudaMemsetAsync(p.flag_count, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  const int grid = p.row_tiles < num_sms ? p.row_tiles : num_sms;
  if (dim == 64)
    vq_assign_kernel<64><<<grid, kVqThreads, vq_assign_smem_bytes(dim, n_embed), st>>>(p, *map_e, *map_x);
  else
    vq_assign_kernel<128><<<grid, kVqThreads, vq_assign_smem_bytes(dim, n_embed), st>>>(p, *map_e, *map_x);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  long long* ind = reinterpret_cast<long long*>(embed_ind);
  if (dim == 64)
    vq_refine_kernel<64><<<num_sms * 4, kVqRefineThreads, 0, st>>>(x, e_t, e_norm2, n_embed, p.flag_count, p.flag_rows,
                                                                   p.flag_u, ind);
  else
    vq_refine_kernel<128><<<num_sms * 4, kVqRefineThreads, 0, st>>>(x, e_t, e_norm2, n_embed, p.flag_count, p.flag_rows,
                                                                    p.flag_u, ind);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (n_flagged != nullptr) e = cudaMemcpyAsync(n_flagged, p.flag_count, sizeof(int), cudaMemcpyDeviceToDevice, st);
  return e;
}
// Codebooks the tensor-core kernel is not specialised for (dim other than 64 / 128, n_embed not a multiple of 16 or above
// 8192; the reference's Quantize takes any, models/vqvae_conv3d_latent.py:34-45): every row is scored against every code in
// fp64 exactly like the flagged rows above (e^2 - 2 x e accumulated term by term, first minimum wins) -- bit-exact by
// construction, CUDA cores only.  A CTA takes 8 rows at a time (dynamic shared memory [8][dim]); thread t scans codes t,
// t + 128, ... and reads each code row once for the 8 rows.
__global__ void __launch_bounds__(kVqRefineThreads)
vq_assign_generic_kernel(const float* __restrict__ x, const float* __restrict__ e_t, size_t rows, int dim, int n_embed,
                         long long* __restrict__ embed_ind) {
  constexpr int R = kVqRefineRows;
  extern __shared__ __align__(16) float sxg[];   // [R][dim]
  __shared__ double s_best[R][kVqRefineThreads / 32];
  __shared__ int s_idx[R][kVqRefineThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (size_t r0 = (size_t)blockIdx.x * R; r0 < rows; r0 += (size_t)gridDim.x * R) {
    __syncthreads();   // previous group fully consumed
    for (int i = threadIdx.x; i < R * dim; i += kVqRefineThreads) {
      const int r = i / dim, d = i % dim;
      sxg[i] = r0 + r < rows ? x[(r0 + r) * dim + d] : 0.f;
    }
    __syncthreads();
    double best[R];
    int besti[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { best[r] = 1e300; besti[r] = 0x7fffffff; }
    for (int k = threadIdx.x; k < n_embed; k += kVqRefineThreads) {
      const float* er = e_t + (size_t)k * dim;
      double acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.0;
      for (int d = 0; d < dim; ++d) {
        const double e = (double)__ldg(er + d);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] += e * (e - 2.0 * (double)sxg[r * dim + d]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (acc[r] < best[r]) { best[r] = acc[r]; besti[r] = k; }   // k ascends: the first minimum is kept
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double bv = best[r];
      int bi = besti[r];
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < bv || (ob == bv && oi < bi)) { bv = ob; bi = oi; }
      }
      if (lane == 0) { s_best[r][warp] = bv; s_idx[r][warp] = bi; }
    }
    __syncthreads();
    if (threadIdx.x < R && r0 + threadIdx.x < rows) {
      const int r = threadIdx.x;
      double bv = s_best[r][0];
      int bi = s_idx[r][0];
#pragma unroll
      for (int w = 1; w < kVqRefineThreads / 32; ++w)
        if (s_best[r][w] < bv || (s_best[r][w] == bv && s_idx[r][w] < bi)) { bv = s_best[r][w]; bi = s_idx[r][w]; }
      embed_ind[r0 + r] = bi == 0x7fffffff ? 0 : bi;   // a NaN row compares false everywhere: index 0 like the fast path
    }
  }
}
bool vq_assign_is_generic(int dim, int n_embed) { return (dim != 64 && dim != 128) || n_embed % 16 != 0 || n_embed > 8192; }
cudaError_t launch_vq_assign_generic(const float* x, size_t rows, int dim, int n_embed, const float* e_t, int64_t* embed_ind,
                                     int* n_flagged, int num_sms, cudaStream_t st) {
  const size_t smem = (size_t)kVqRefineRows * dim * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  size_t blocks = (rows + kVqRefineRows - 1) / kVqRefineRows;
  if (blocks > (size_t)num_sms * 16) blocks = (size_t)num_sms * 16;
  vq_assign_generic_kernel<<<(int)blocks, kVqRefineThreads, smem, st>>>(x, e_t, rows, dim, n_embed,
                                                                        reinterpret_cast<long long*>(embed_ind));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (n_flagged != nullptr) e = cudaMemsetAsync(n_flagged, 0, sizeof(int), st);
  return e;
}

cudaError_t init_vq() {
  cudaError_t e = cudaFuncSetAttribute(vq_assign_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(vq_assign_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
}

// =============================================================================== gather + ST + loss + EMA stats
// LPR lanes share a row (float4 each).  Per-CTA statistics are privatised in shared memory
// ([n_embed][dim] fp32 + counts) when they fit, then flushed with one global atomic per entry.
template <bool SMEM_STATS>
__global__ void __launch_bounds__(512)
vq_gather_stats_kernel(const float* __restrict__ x, const long long* __restrict__ ind, size_t rows,
                                       int dim, int n_embed, const float* __restrict__ e_t, float* __restrict__ q_f32,
                                       __nv_bfloat16* __restrict__ q_bf16, float* __restrict__ diff_sum,
                                       float* __restrict__ counts, float* __restrict__ embed_sum,
                                       float* __restrict__ sum_t) {
  extern __shared__ float sm[];  // SMEM_STATS: [n_embed*dim] sums, [n_embed] counts
  __shared__ float red[32];
  const bool stats = counts != nullptr;
  if (SMEM_STATS && stats) {
    for (int i = threadIdx.x; i < n_embed * (dim + 1); i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
  }
  float* s_sum = sm;
  float* s_cnt = sm + (size_t)n_embed * dim;
  const int q4 = dim / 4;                 // lanes per row (16 or 32; dim is 64 or 128 on this path, any multiple of 4 works)
  float dacc = 0.f;
  // A group of q4 lanes walks kSeg CONSECUTIVE rows; statistics of a run of rows with the same code are summed in
  // registers and flushed with one atomic per component when the code changes.  Neighbouring latent pixels mostly share
  // their code (and an untrained / collapsed codebook uses a handful of codes), which is exactly when per-row atomics on
  // the same few addresses would serialise.
  constexpr int kSeg = 32;
  const int groups_per_block = blockDim.x / q4;
  const int q = threadIdx.x % q4;
  const size_t group = (size_t)blockIdx.x * groups_per_block + threadIdx.x / q4;
  const size_t n_groups = (size_t)gridDim.x * groups_per_block;
  const size_t n_seg = (rows + kSeg - 1) / kSeg;
  auto flush = [&](int k, const float4& a, float cnt) {
    if (k < 0) return;
    if (SMEM_STATS) {
      float* d = s_sum + (size_t)k * dim + q * 4;
      atomicAdd(d + 0, a.x); atomicAdd(d + 1, a.y); atomicAdd(d + 2, a.z); atomicAdd(d + 3, a.w);
      if (q == 0) atomicAdd(s_cnt + k, cnt);
    } else if (sum_t != nullptr) {
      // codebooks whose statistics do not fit in shared memory: one 16-byte vector atomic per lane into a TRANSPOSED
      // scratch [n_embed][dim] (the dim floats of a code are contiguous: 2-4 lines per flushed run instead of dim lines
      // n_embed floats apart); vq_stats_transpose_add folds it into embed_sum afterwards
      atomicAdd(reinterpret_cast<float4*>(sum_t + (size_t)k * dim) + q, a);
      if (q == 0) atomicAdd(counts + k, cnt);
    } else {
      const int d0 = q * 4;
      atomicAdd(embed_sum + (size_t)(d0 + 0) * n_embed + k, a.x);
      atomicAdd(embed_sum + (size_t)(d0 + 1) * n_embed + k, a.y);
      atomicAdd(embed_sum + (size_t)(d0 + 2) * n_embed + k, a.z);
      atomicAdd(embed_sum + (size_t)(d0 + 3) * n_embed + k, a.w);
      if (q == 0) atomicAdd(counts + k, cnt);
    }
  };
  const bool lane_live = (int)threadIdx.x < groups_per_block * q4;   // blockDim need not be a multiple of q4
  for (size_t seg = lane_live ? group : n_seg; seg < n_seg; seg += n_groups) {
    const size_t r0 = seg * kSeg;
    int cur_k = -1;
    float4 run = make_float4(0.f, 0.f, 0.f, 0.f);
    float run_n = 0.f;
#pragma unroll 1
    for (int j0 = 0; j0 < kSeg; j0 += 4) {
      int kk[4];
      float4 xv[4], ev[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {   // four rows in flight
        const size_t r = r0 + j0 + u;
        kk[u] = r < rows ? (int)ind[r] : -1;
        xv[u] = ev[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kk[u] >= 0) {
          xv[u] = __ldg(reinterpret_cast<const float4*>(x + r * dim) + q);
          ev[u] = __ldg(reinterpret_cast<const float4*>(e_t + (size_t)kk[u] * dim) + q);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (kk[u] < 0) continue;
        const size_t r = r0 + j0 + u;
        // (:77) diff = (quantize - input)^2 ; (:78) quantize = input + (quantize - input)
        const float t0 = ev[u].x - xv[u].x, t1 = ev[u].y - xv[u].y, t2 = ev[u].z - xv[u].z, t3 = ev[u].w - xv[u].w;
        dacc += t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3;
        const float4 qv = make_float4(xv[u].x + t0, xv[u].y + t1, xv[u].z + t2, xv[u].w + t3);
        if (q_f32 != nullptr) reinterpret_cast<float4*>(q_f32 + r * dim)[q] = qv;
        if (q_bf16 != nullptr) {
          uint2 o;
          o.x = pack_bf16x2(qv.x, qv.y);
          o.y = pack_bf16x2(qv.z, qv.w);
          reinterpret_cast<uint2*>(q_bf16 + r * dim)[q] = o;
        }
        if (stats) {
          if (kk[u] != cur_k) {
            flush(cur_k, run, run_n);
            cur_k = kk[u];
            run = make_float4(0.f, 0.f, 0.f, 0.f);
            run_n = 0.f;
          }
          run.x += xv[u].x; run.y += xv[u].y; run.z += xv[u].z; run.w += xv[u].w;
          run_n += 1.f;
        }
      }
    }
    if (stats) flush(cur_k, run, run_n);
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

// embed_sum[d][k] += sum_t[k][d]   (tiny: n_embed * dim floats; tiled through shared memory, both sides coalesced)
__global__ void vq_stats_transpose_add_kernel(const float* __restrict__ sum_t, float* __restrict__ embed_sum, int dim,
                                              int n_embed) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, d = d0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < n_embed && d < dim) ? sum_t[(size_t)k * dim + d] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, k = k0 + threadIdx.x;
    if (d < dim && k < n_embed) embed_sum[(size_t)d * n_embed + k] += tile[threadIdx.x][r];
  }
}

size_t vq_gather_scratch_bytes(int dim, int n_embed) {
  const size_t smem = (size_t)n_embed * (dim + 1) * sizeof(float);
  static const char* force = getenv("FO_VQ_SMEM_STATS");
  return (smem <= 200 * 1024 && !(force && atoi(force) == 0)) ? 0 : (size_t)n_embed * dim * sizeof(float);
}

cudaError_t launch_vq_gather_stats(const float* x, const int64_t* ind, size_t rows, int dim, int n_embed,
                                   const float* e_t, float* q_f32, void* q_bf16, float* diff_sum, float* counts,
                                   float* embed_sum, float* scratch, int num_sms, cudaStream_t st) {
  size_t smem = (size_t)n_embed * (dim + 1) * sizeof(float);
  const size_t total = rows * (dim / 4);
  size_t blocks = (total + 255) / 256;
  {
    static const char* force = getenv("FO_VQ_SMEM_STATS");   // experiments only: 0 = always the global vector-atomic path
    if (force && atoi(force) == 0 && scratch != nullptr) smem = (size_t)1 << 30;
  }
  if (smem <= 200 * 1024) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(vq_gather_stats_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kMaxDynSmem - 1024);  // the kernel also has 128 B of static smem
      if (e != cudaSuccess) return e;
      configured = true;
    }
    if (blocks > (size_t)num_sms) blocks = num_sms;  // one CTA per SM (smem-limited)
    vq_gather_stats_kernel<true><<<(int)blocks, 512, smem, st>>>(x, (const long long*)ind, rows, dim, n_embed, e_t,
                                                                   q_f32, (__nv_bfloat16*)q_bf16, diff_sum, counts,
                                                                   embed_sum, nullptr);
  } else {
    const bool use_t = scratch != nullptr && counts != nullptr && dim % 4 == 0;
    if (use_t) {
      cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)n_embed * dim * sizeof(float), st);
      if (e != cudaSuccess) return e;
    }
    if (blocks > (size_t)num_sms * 8) blocks = (size_t)num_sms * 8;
    vq_gather_stats_kernel<false><<<(int)blocks, 256, 0, st>>>(x, (const long long*)ind, rows, dim, n_embed, e_t, q_f32,
                                                                (__nv_bfloat16*)q_bf16, diff_sum, counts, embed_sum,
                                                                use_t ? scratch : nullptr);
    if (use_t) {
      dim3 grid((n_embed + 31) / 32, (dim + 31) / 32), block(32, 8);
      vq_stats_transpose_add_kernel<<<grid, block, 0, st>>>(scratch, embed_sum, dim, n_embed);
    }
  }
  return cudaGetLastError();
}

// =============================================================================== EMA (:66-75)
__global__ void vq_ema_kernel(float* __restrict__ embed, float* __restrict__ cluster_size, float* __restrict__ embed_avg,
                              const float* __restrict__ counts, const float* __restrict__ embed_sum, int dim,
                              int n_embed, float decay, float omd, float eps) {
  __shared__ float red[32];
  __shared__ float n_total;
  float part = 0.f;
  for (int k = threadIdx.x; k < n_embed; k += blockDim.x) {
    const float cs = cluster_size[k] * decay + counts[k] * omd;
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
    const float ea = embed_avg[i] * decay + embed_sum[i] * omd;
    embed_avg[i] = ea;
    const float cs = (cluster_size[k] + eps) / denom * n;
    embed[i] = ea / cs;
  }
}
cudaError_t launch_vq_ema(float* embed, float* cluster_size, float* embed_avg, const float* counts,
                          const float* embed_sum, int dim, int n_embed, float decay, float one_minus_decay, float eps,
                          cudaStream_t st) {
  vq_ema_kernel<<<1, 1024, 0, st>>>(embed, cluster_size, embed_avg, counts, embed_sum, dim, n_embed, decay,
                                    one_minus_decay, eps);
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
