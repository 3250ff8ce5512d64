// Parameter blocks shared by the host planners (api.cu) and the tcgen05 implicit-GEMM kernels.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace fo {

constexpr int kMaxKSteps = 256;   // 27 taps x 4 channel chunks, VGG 9 x 8, x3 in the split (hi|lo) verification mode
constexpr int kMaxAMaps = 6;       // sources of one conv: torch.cat of 2, x3 in the split (hi|lo) verification mode
constexpr int kMaxDynSmem = 232448;  // 227 KB opt-in limit per CTA on sm_100
constexpr int kMaxGroups = 4;     // sub-pixel (parity) classes of a stride-2 transposed conv

// One K step of the implicit GEMM: which activation map, which channel chunk, which spatial shift.
struct KStep {
  int32_t c0;      // coordinate on map dim 0 (channels; the GEMM forms of the discriminator path reach > 2^15)
  int8_t d1, d2, d3;  // coordinate deltas on map dims 1..3
  uint8_t map;     // A-operand tensor map index
};

// Division by a launch-time constant as multiply-high + shift (dividends < 2^31): the tile decode runs once per tile in the
// producer and in every epilogue warp, where a hardware-less integer division costs ~25 dependent instructions.
struct FastDiv {
  uint32_t mul, shr, div;
#if defined(__CUDACC__)
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
    q = div == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shr);
    r = n - q * (int)div;
  }
#endif
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.div = (uint32_t)d;
  if (d <= 1) { f.div = 1; f.mul = 0; f.shr = 0; return f; }
  int lg = 0;
  while ((1u << lg) < (uint32_t)d) ++lg;           // ceil(log2 d)
  const unsigned p = 31 + lg;
  f.mul = (uint32_t)(((1ull << p) + (uint32_t)d - 1) / (uint32_t)d);
  f.shr = p - 32;
  return f;
}

// conv_igemm: D[128 pixels, NT] = sum_k A_k[128, KC] * W_k[NT, KC]^T   (+ fused epilogue)
struct ConvParams {
  // M tiling: a tile is a box of box[0] x box[1] x box[2] x box[3] positions on map dims 1..4, product 128*MT
  int tile_cnt[4];    // tiles along map dims 1..4 (dim 1 fastest)
  int tile_step[4];   // coordinate step per tile on dims 1..4
  int box[4];         // in-tile extents on dims 1..4
  int lim[4];         // valid output extents on dims 1..4 (partial tiles are predicated)
  int MT;             // M sub-tiles of 128 rows per CTA tile (1 or 2); they share every B tile
  int TPS;            // filter taps per pipeline stage (1, or 3 = the three vertical taps read from ONE halo'd A box)
  int a_bytes;        // bytes of the A box of one stage
  int sub_off;        // byte offset of M sub-tile 1 inside the A box
  int tap_off;        // byte offset between consecutive taps inside the A box (one image row)
  int a_ops;          // the A box is fetched by a_ops TMA operations of a_op_rows image rows each (more ops in flight)
  int a_op_rows;
  int dbg_skip_mma;   // debug: do not issue MMAs (measures the pure TMA streaming rate)
  int backoff_ns;     // sleep between mbarrier probes of the long waits (0 = plain polling)
  int cta_pair;       // 1: CTA pairs (cluster of 2, tcgen05 cta_group::2): tiles 2i / 2i+1 share every B tile
  int pos_tiles;      // position tiles per (group, n-tile) = product of tile_cnt
  FastDiv fd_pos, fd_nt, fd_cnt[4];   // dividers by pos_tiles, n_tiles, tile_cnt[d]
  int e_bufs;         // epilogue operand prefetch: 0 = off, else MT (one buffer set per M sub-tile / warp group)
  int e_depth;        // 1: rows of a tile are fetched while its MMAs run; 2: one tile ahead (double-buffered)
  int e_mask, e_add;  // which of the two epilogue operands are prefetched (shared memory permitting)
  int e_box[4];       // extents on map dims 1..4 of the 32 rows owned by one epilogue warp (an aligned sub-box of the tile)
  int e_cols;         // channels per prefetch box: 64 (128-byte rows, 128B swizzle) or 32 (64-byte rows, 64B swizzle)
  int n_tiles;        // N tiles of width NT
  int NT;             // columns per N tile (multiple of 16, <= 256)
  int KC;             // channels per K step: 16 / 32 / 64  (row bytes 32 / 64 / 128 = swizzle mode)
  int num_ksteps;     // per group
  int groups;
  int stages;
  int total_tiles;    // groups * n_tiles * prod(tile_cnt)
  // epilogue
  long long out_stride[4];       // element strides of dims 1..4 in the output tensors
  long long out_off[kMaxGroups]; // per-group element offset
  int out_cstride;               // channel stride of out_f32 (1 = channels-last)
  int c_store;                   // channels actually stored (<= n_tiles*NT)
  const float* bias;             // [>= n_tiles*NT] or null
  const __nv_bfloat16* mask;     // zero the result where mask <= 0 (same layout as out_bf16) or null
  const __nv_bfloat16* addend;   // added after masking (same layout) or null
  __nv_bfloat16* out_bf16;       // raw result or null
  __nv_bfloat16* out_relu;       // relu(result) or null
  float* out_f32;                // raw result in fp32 (channel stride out_cstride) or null
  int relu_f32;                  // apply relu to out_f32 too
  int acc_f32;                   // out_f32 += result (K split over several launches: the discriminator GEMMs)
  int split_off;                 // > 0: bf16 tensors of the epilogue are hi|lo pairs, lo part split_off channels after hi
  KStep ksteps[kMaxKSteps];
};

struct ConvMaps {
  CUtensorMap a[kMaxAMaps];
  CUtensorMap b;
  CUtensorMap e[2][kMaxGroups];   // epilogue operand maps ([0] mask, [1] addend; one per sub-pixel group), see e_bufs
};

// wgrad_igemm: dW[g][m][n] = sum_pixels P[pix][m] * Q[pix (+) shift_g][n]; both operands MN-major.
constexpr int kMaxWgTaps = 9;  // accumulators resident in TMEM per CTA pass (taps_per_pass * NC <= 512 columns)
struct WgTap {
  int16_t c0;
  int8_t d1, d2, d3;
  uint8_t map;
  int16_t pad;
};
struct WgradParams {
  int tile_cnt[4];
  int tile_step[4];
  int box[3];            // product = pixels per K step (64)
  int MC;                // P-side channels (rows of dW tile): 64 or 128 (UMMA M)
  int p_chunks;          // TMA boxes on the P side (MC/chunk width)
  int p_rowb;            // bytes per pixel row in a P box (32/64/128)
  int NC;                // Q-side channels per tap (UMMA N), multiple of 16, <= 128
  int q_chunks;
  int q_rowb;
  int kpix;              // pixels per pipeline stage (64 or 128); kpix/16 MMAs per tap group.  128 whenever two stages
                         // fit: a stage costs ~1.4 k cycles of fixed hand-shake latency whatever its size (measured: 1x1
                         // layers 4.2 -> 5.7 TB/s, 4x4 stride-2 layers 2.13 -> 1.15 ms when going from 64 to 128 pixels)
  int halo;              // 1: ONE Q box of R+2 image rows holds the 3 vertical taps of a filter column
  int q_tap_off;         // halo: byte offset between consecutive taps inside the Q box (one image row)
  int q_box_bytes;       // bytes of one Q chunk box
  int taps_per_pass;     // <= kMaxWgTaps
  int q_loads;           // Q boxes fetched per stage (each q_chunks TMA operations)
  int taps_per_load;     // taps served by one Q box: 3 with halo, else 1 (taps_per_pass = q_loads * taps_per_load)
  int mma_group;         // G consecutive taps share ONE MMA of N = G * NC columns: their Q boxes lie at a uniform byte
                         // stride in the stage, which an MN-major descriptor expresses as the chunk stride (LBO); the
                         // narrow-N layers (NC <= 64) otherwise pay the A-operand read (64 cycles per MMA) once per tap
  int passes;
  int splits;            // split-K factor over pixel tiles
  int total_ptiles;      // prod(tile_cnt)
  int stages;
  int chain_tiles;       // pixel tiles accumulated into a TMEM accumulator before it is written out and restarted (see
                         // wgrad_igemm.cu "accumulation chains"); a CTA's range is cut into n_flush such chunks
  int n_flush;           // chunks per split = partial slots per split
  int p_c0;              // channel offset of the P side inside its map
  int dbg_skip_mma;      // debug: do not issue MMAs (pure TMA streaming rate)
  int backoff_ns;        // sleep between mbarrier probes of the long waits (0 = plain polling)
  float* partial;        // [splits * n_flush][passes*taps_per_pass][MC][NC] fp32
  float* bias_partial;   // [passes][splits * n_flush][MC] partial column sums of P (= bias gradient) or null
  WgTap taps[64];        // passes * taps_per_pass entries (padded entries have map = 255)
};
// Accumulation chains of the weight gradient (wgrad_igemm.cu): a split's range of `per` pixel tiles (kpix / 16 accumulating
// MMAs each) is cut into n_flush balanced chunks of chain_tiles tiles so that no TMEM accumulator sums more than chain_mmas
// MMAs (0 = unbounded).  Chunk c of a CTA with n_my tiles covers [min(n_my, c * chain_tiles), min(n_my, (c + 1) * chain_tiles)).
inline void wgrad_plan_chains(int per, int kpix, int chain_mmas, int* n_flush, int* chain_tiles) {
  const int mmas_per_tile = kpix / 16;
  int nf = chain_mmas > 0 ? (int)(((long long)per * mmas_per_tile + chain_mmas - 1) / chain_mmas) : 1;
  if (nf < 1) nf = 1;
  int ct = (per + nf - 1) / nf;
  if (ct < 1) ct = 1;
  *n_flush = nf;
  *chain_tiles = ct;
}

struct WgradMaps {
  CUtensorMap p;
  CUtensorMap q[kMaxAMaps];
};

}  // namespace fo
