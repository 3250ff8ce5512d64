// Implicit-GEMM convolution for sm_100a: TMA-fed tcgen05.mma with TMEM accumulators.
//
// One kernel covers every forward / data-gradient convolution on the FaceOff hot path
// (reference models/vqvae_conv3d_latent.py:86-190 Conv2d k1/k3/k4s2, ConvTranspose2d k4s2, Conv3d k3;
// models/lpips.py:115-152 VGG16 3x3).  The kernel knows nothing about convolutions: the host planner
// (api.cu) describes the op as
//   * an M tiling of the output into boxes of 128 positions on a <=5-D channels-last tensor map,
//   * a list of K steps (tensor map, channel chunk, spatial shift).  Out-of-range coordinates are
//     zero-filled by TMA, which implements the conv padding; stride-2 convs address a parity
//     (space-to-depth) view of the same memory; torch.cat inputs are just K steps on another map,
//   * a packed bf16 weight matrix [Cout][K steps * KC] in the same K-step order.
//
// Warp roles (320 threads, 1 CTA / SM, persistent over tiles):
//   warp 0      TMA producer      (one A box + TPS B boxes [NT x KC] per stage, 128B/64B/32B swizzle)
//   warp 1      MMA issuer        (one elected lane issues tcgen05.mma kind::f16 BF16xBF16->FP32, M=128 per CTA, N=NT)
//   warps 2..9  epilogue          (tcgen05.ld -> +bias -> *mask -> +addend -> relu -> bf16/fp32 stores); warps 2-5 own
//                                 M sub-tile 0, warps 6-9 sub-tile 1 (two warps per scheduler hide each other's latency)
// TMEM holds two sets of accumulators (2 x MT x NT columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// What bounds the tensor-bound layers is the SM's shared-memory port, shared by TMA writes and UMMA operand reads (an
// SS-mode 128x128x16 MMA reads 4 KB of A + 4 KB of B in its 64 math cycles; measured: Conv3d with TMA only 1.21 ms, MMAs
// only 1.97 ms, both 2.44 ms), so operand reuse is what matters:
//   MT = 2   the CTA tile is 256 output positions = two M=128 MMAs per B tile (B traffic / 2);
//   TPS = 3  "halo" stages for 3x3(x3) filters: the A box holds R+2 image rows; the three vertical taps are the SAME
//            shared-memory box read through descriptors offset by one image row (S*rowb bytes, a multiple of the
//            1024-byte swizzle atom), so A traffic drops by 3R/(R+2);
//   CG = 2   CTA pairs (cta_group::2, M = 256 over two SMs): each CTA loads half of every B tile (long-K layers).
#include "common.cuh"
#include "igemm.cuh"

namespace fo {

constexpr int kConvThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue of sub-tile 0, warps 6-9 of sub-tile 1

__device__ __forceinline__ void decode_tile(const ConvParams& p, int tile, int& g, int& nt, int (&base)[4]) {
  int t;
  if (p.cta_pair) {
    // position tile fastest: tiles 2i and 2i+1 (the two CTAs of a pair) share the n-tile and the group
    int rest;
    p.fd_pos.divmod(tile, rest, t);
    int gq;
    p.fd_nt.divmod(rest, gq, nt);
    t += gq * p.pos_tiles;
  } else {
    // n-tile fastest so CTAs that share an A tile run concurrently and hit L2
    p.fd_nt.divmod(tile, t, nt);
  }
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    int i;
    p.fd_cnt[d].divmod(t, t, i);
    base[d] = i * p.tile_step[d];
  }
  g = t;
}

template <int NCH>
__device__ __forceinline__ void epilogue_chunk(const ConvParams& p, const uint32_t (&v)[NCH], int col0, bool valid,
                                               long long off) {
  // col0: first output channel of this chunk (global channel index); off: element offset of channel 0
  float f[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) f[j] = __uint_as_float(v[j]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) f[j] += __ldg(p.bias + col0 + j);
  }
  if (!valid) return;
  const bool full = (col0 + NCH <= p.c_store);
  if (p.mask != nullptr && full) {
    const uint4* mp = reinterpret_cast<const uint4*>(p.mask + off + col0);
#pragma unroll
    for (int q = 0; q < NCH / 8; ++q) {
      uint4 m = __ldg(mp + q);
      uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (!(bf16lo(w[e]) > 0.f)) f[q * 8 + 2 * e] = 0.f;
        if (!(bf16hi(w[e]) > 0.f)) f[q * 8 + 2 * e + 1] = 0.f;
      }
    }
  }
  if (p.addend != nullptr && full) {
    // split mode: the addend is a hi|lo pair, both halves are added
    for (int part = 0; part < (p.split_off > 0 ? 2 : 1); ++part) {
      const uint4* ap = reinterpret_cast<const uint4*>(p.addend + off + part * p.split_off + col0);
#pragma unroll
      for (int q = 0; q < NCH / 8; ++q) {
        uint4 m = __ldg(ap + q);
        uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          f[q * 8 + 2 * e] += bf16lo(w[e]);
          f[q * 8 + 2 * e + 1] += bf16hi(w[e]);
        }
      }
    }
  }
  if (p.out_f32 != nullptr) {
    if (p.out_cstride == 1 && full) {
      float4* op = reinterpret_cast<float4*>(p.out_f32 + off + col0);
#pragma unroll
      for (int q = 0; q < NCH / 4; ++q) {
        float4 o = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        if (p.acc_f32) {
          const float4 old = op[q];
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        if (p.relu_f32) {
          o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        }
        op[q] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        if (col0 + j < p.c_store) {
          float* dst = p.out_f32 + off + (long long)(col0 + j) * p.out_cstride;
          float o = f[j];
          if (p.acc_f32) o += *dst;
          *dst = p.relu_f32 ? fmaxf(o, 0.f) : o;
        }
      }
    }
  }
  if (full) {
    // bf16 outputs: value (and, in split mode, the bf16 rounding residual lo = bf16(v - hi) split_off channels further)
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* dst = which == 0 ? p.out_bf16 : p.out_relu;
      if (dst == nullptr) continue;
      float v[NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) v[j] = which == 0 ? f[j] : fmaxf(f[j], 0.f);
      for (int part = 0; part < (p.split_off > 0 ? 2 : 1); ++part) {
        uint4* op = reinterpret_cast<uint4*>(dst + off + part * p.split_off + col0);
#pragma unroll
        for (int q = 0; q < NCH / 8; ++q) {
          uint4 o;
          o.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
          o.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
          o.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
          o.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
          op[q] = o;
          if (p.split_off > 0) {
            const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[8 * q + 2 * e] -= bf16lo(w[e]);
              v[8 * q + 2 * e + 1] -= bf16hi(w[e]);
            }
          }
        }
      }
    }
  }
}

// Coalesced variant for full 32-column chunks of channels-last bf16 tensors.  TMEM gives each lane one pixel row, but
// a pixel's channels are contiguous in memory, so lane-per-row accesses touch 32 different lines per instruction.
// Every tensor access therefore goes through a per-warp shared-memory transpose: 8 rows x 64 B per instruction
// (full 32-byte sectors, 4 lanes per row).  stg: this warp's staging buffer, 32 rows x 80 B (16 B pad: conflict-free).
// roff / rvalid: element offset and validity of row i*8 + lane/4 (i = 0..3), shuffled once per tile.
constexpr int kStgPitch = 80;
__device__ __forceinline__ void stage_load_rows(const __nv_bfloat16* src, int col0, const long long (&roff)[4],
                                                const bool (&rvalid)[4], uint8_t* stg, int lane, uint32_t (&w)[16]) {
  const int piece = lane & 3, r0 = lane >> 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (rvalid[i]) v = __ldg(reinterpret_cast<const uint4*>(src + roff[i] + col0) + piece);
    *reinterpret_cast<uint4*>(stg + (i * 8 + r0) * kStgPitch + piece * 16) = v;
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 v = *reinterpret_cast<const uint4*>(stg + lane * kStgPitch + q * 16);
    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
  }
  __syncwarp();
}
__device__ __forceinline__ void stage_store_rows(__nv_bfloat16* dst, int col0, const long long (&roff)[4],
                                                 const bool (&rvalid)[4], uint8_t* stg, int lane,
                                                 const uint32_t (&w)[16]) {
  const int piece = lane & 3, r0 = lane >> 2;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(stg + lane * kStgPitch + q * 16) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 v = *reinterpret_cast<const uint4*>(stg + (i * 8 + r0) * kStgPitch + piece * 16);
    if (rvalid[i]) *(reinterpret_cast<uint4*>(dst + roff[i] + col0) + piece) = v;
  }
  __syncwarp();
}

// Prefetched mask / addend tile of one warp: [column block][32 rows][e_cols * 2 bytes], written by TMA with the 128-byte
// (e_cols = 64) or 64-byte (e_cols = 32) swizzle.  Lane L reads the 32 channels starting at column ccol of ITS row L:
// four 16-byte chunks whose physical position is chunk ^ (row & 7) resp. chunk ^ ((row >> 1) & 3) -- conflict-free.
__device__ __forceinline__ void load_prefetched(const uint8_t* ebase, int ccol, int e_cols, int lane, uint32_t (&w)[16]) {
  // e_cols is 64 or 32: shifts, not divisions (this runs once per operand and 32-column chunk)
  const int sh = e_cols == 64 ? 6 : 5;
  const int rowb = e_cols * 2;
  const int cb = ccol >> sh;
  const uint8_t* row = ebase + cb * (32 * rowb) + lane * rowb;
  const int k0 = (ccol & (e_cols - 1)) >> 3;                 // first logical 16-byte chunk inside the row
  const int sw = e_cols == 64 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 v = *reinterpret_cast<const uint4*>(row + (((k0 + q) ^ sw) << 4));
    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
  }
}

// e_mask / e_add: this lane's row in the prefetched (cp.async) copies of the mask / addend tile, or null.
__device__ __forceinline__ void epilogue_chunk32_coalesced(const ConvParams& p, const uint32_t (&v)[32], int col0,
                                                           int ccol, const long long (&roff)[4],
                                                           const bool (&rvalid)[4], uint8_t* stg, int lane,
                                                           const uint8_t* e_mask, const uint8_t* e_add,
                                                           const float* s_bias) {
  // e_mask / e_add: base of this warp's prefetched tile (see load_prefetched), or null
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if (s_bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = *reinterpret_cast<const float4*>(s_bias + ccol + 4 * q);
      f[4 * q] += b.x; f[4 * q + 1] += b.y; f[4 * q + 2] += b.z; f[4 * q + 3] += b.w;
    }
  } else if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + col0 + j);
  }
  uint32_t w[16];
  if (p.mask != nullptr) {
    if (e_mask != nullptr) load_prefetched(e_mask, ccol, p.e_cols, lane, w);
    else stage_load_rows(p.mask, col0, roff, rvalid, stg, lane, w);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      if (!(bf16lo(w[e]) > 0.f)) f[2 * e] = 0.f;
      if (!(bf16hi(w[e]) > 0.f)) f[2 * e + 1] = 0.f;
    }
  }
  if (p.addend != nullptr) {
    if (e_add != nullptr) load_prefetched(e_add, ccol, p.e_cols, lane, w);
    else stage_load_rows(p.addend, col0, roff, rvalid, stg, lane, w);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      f[2 * e] += bf16lo(w[e]);
      f[2 * e + 1] += bf16hi(w[e]);
    }
  }
  if (p.out_bf16 != nullptr) {
#pragma unroll
    for (int e = 0; e < 16; ++e) w[e] = pack_bf16x2(f[2 * e], f[2 * e + 1]);
    if (p.dbg_skip_mma != 3) stage_store_rows(p.out_bf16, col0, roff, rvalid, stg, lane, w);
    else if (w[0] == 0x12345678u) p.out_bf16[0] = __float2bfloat16(1.f);   // debug: keep the math alive, no stores
  }
  if (p.out_relu != nullptr) {
#pragma unroll
    for (int e = 0; e < 16; ++e) w[e] = pack_bf16x2(fmaxf(f[2 * e], 0.f), fmaxf(f[2 * e + 1], 0.f));
    if (p.dbg_skip_mma != 3) stage_store_rows(p.out_relu, col0, roff, rvalid, stg, lane, w);
    else if (w[0] == 0x12345678u) p.out_relu[0] = __float2bfloat16(1.f);
  }
}

template <int CG>
__device__ __forceinline__ void umma_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (CG == 2) umma_bf16_pair(tmem_d, adesc, bdesc, idesc, acc);
  else umma_bf16(tmem_d, adesc, bdesc, idesc, acc);
}
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint64_t* bar) {
  if (CG == 2) umma_commit_pair(bar);
  else umma_commit(bar);
}

// CG = 1: one CTA per tile.  CG = 2: the CTAs of a cluster of two work on tiles 2i / 2i+1 with cta_group::2 MMAs (M = 256):
// each CTA loads its own A box and HALF of every B tile, so B costs half the L2->SM traffic and half the shared-memory
// reads per SM.  Rank 0 issues the MMAs for both; see common.cuh "CTA pairs".
template <int CG>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ ConvParams p, const __grid_constant__ ConvMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A tile | B tile)] | barriers | tmem ptr
  // (offset arithmetic, not a pointer round-trip: the compiler keeps the shared address space -> LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int rowb = p.KC * 2;
  const int a_bytes = p.a_bytes;
  const int b_bytes = (p.NT / CG) * rowb;   // this CTA's share of a B tile
  const int stage_bytes = (a_bytes + p.TPS * b_bytes + 1023) & ~1023;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;
  const int tile0 = CG == 2 ? (int)(blockIdx.x & ~1u) + (int)cta_rank : (int)blockIdx.x;   // first tile of this CTA
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;                  // [stages]
  uint64_t* empty_bar = bars + p.stages;      // [stages]
  uint64_t* tfull_bar = bars + 2 * p.stages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint64_t* e_bar = tempty_bar + 2;           // [8 epilogue warps][2 slots]: operand prefetch (TMA) completion
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(e_bar + 16);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr + 4);        // [256] bias of the (single) N tile
  uint8_t* stg_all = reinterpret_cast<uint8_t*>(s_bias + 256);   // 8 epilogue warps x 32 rows x 80 B
  uint8_t* e_all = stg_all + 8 * 32 * kStgPitch;                // prefetched mask / addend tiles (see epilogue)
  e_all += (1024u - (smem_u32(e_all) & 1023u)) & 1023u;         // TMA destinations with the 128-byte swizzle

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * p.MT * p.NT) tmem_cols <<= 1;

  pdl_trigger();   // the next kernel of the stream may start its prologue (common.cuh)
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < kMaxAMaps; ++i) tma_prefetch_desc(&maps.a[i]);
      tma_prefetch_desc(&maps.b);
      if (p.e_bufs > 0)
        for (int i = 0; i < 2 * kMaxGroups; ++i) tma_prefetch_desc(&maps.e[i / kMaxGroups][i % kMaxGroups]);
    }
    __syncwarp();
    if (CG == 2) {
      tmem_alloc_pair(tmem_ptr, tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr, tmem_cols);
      tmem_relinquish();
    }
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 16; ++i) mbar_init(&e_bar[i], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4 * p.MT * CG);   // CG == 2: the epilogue warps of both CTAs arrive at rank 0
    }
    fence_mbar_init();
  }
  // everything above overlaps the tail of the preceding kernel; global memory is touched only from here on
  pdl_wait();
  const bool bias_in_smem = p.bias != nullptr && p.n_tiles == 1;
  if (bias_in_smem)
    for (int i = threadIdx.x; i < p.NT; i += blockDim.x) s_bias[i] = p.bias[i];
  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // the peer's barriers must be initialised before anything is sent to them
  else __syncthreads();
  tc_fence_after();
  // warp-uniform copy (the shuffle lets the compiler keep MMA operands in uniform registers)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // One lane per TMA operation of a stage (lanes 0 .. a_ops-1: the A box, the next TPS lanes: the B tiles): the
    // operands of every operation are computed in parallel and the operations are issued back to back.  A single
    // thread issuing them one after the other is a dependent chain of ~230 cycles per operation, which bounded the
    // layers whose stages are short (4x4 stride-2: 2 operations per 512 MMA cycles); see also wgrad_igemm.cu.
    const int n_ops = p.a_ops + p.TPS;
    if (lane < n_ops) {
      const bool is_a = lane < p.a_ops;
      const int j = lane - p.a_ops;                 // B tile index (tap) of this lane
      const int a_op_bytes = a_bytes / p.a_ops;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < p.total_tiles; tile += gridDim.x) {
        int g, nt, base[4];
        decode_tile(p, tile, g, nt, base);
        const KStep* ks = p.ksteps + g * p.num_ksteps;
        const int kcol0 = g * p.num_ksteps * p.TPS * p.KC;
        for (int k = 0; k < p.num_ksteps; ++k) {
          mbar_wait_ns(&empty_bar[stage], phase ^ 1, p.backoff_ns);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          if (p.dbg_skip_mma == 2) {   // debug: MMAs over whatever is in shared memory, no loads
            if (lane == 0) mbar_arrive(&full_bar[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            continue;
          }
          const KStep s = ks[k];
          if (CG == 2) {
            // both CTAs' loads complete on rank 0's barrier, which rank 0 arms for the bytes of the pair
            if (cta_rank == 0 && lane == 0) mbar_expect_tx(&full_bar[stage], 2 * (a_bytes + p.TPS * b_bytes));
            const uint32_t fb = mapa_u32(&full_bar[stage], 0);
            if (is_a)
              tma_load_5d_pair(sa + lane * a_op_bytes, &maps.a[s.map], fb, s.c0, base[0] + s.d1,
                               base[1] + s.d2 + lane * p.a_op_rows, base[2] + s.d3, base[3]);
            else
              tma_load_2d_pair(sb + j * b_bytes, &maps.b, fb, kcol0 + (k * p.TPS + j) * p.KC,
                               nt * p.NT + (int)cta_rank * (p.NT / 2));
          } else {
            if (lane == 0) mbar_expect_tx(&full_bar[stage], a_bytes + p.TPS * b_bytes);
            if (is_a)
              tma_load_5d(sa + lane * a_op_bytes, &maps.a[s.map], &full_bar[stage], s.c0, base[0] + s.d1,
                          base[1] + s.d2 + lane * p.a_op_rows, base[2] + s.d3, base[3]);
            else
              tma_load_2d(sb + j * b_bytes, &maps.b, &full_bar[stage], kcol0 + (k * p.TPS + j) * p.KC, nt * p.NT);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (rank 0 of a pair)
    if (cta_rank == 0) {
    const uint32_t idesc = make_idesc_bf16(128 * CG, p.NT, 0, 0);
    const uint64_t desc_base = make_smem_desc(0, rowb, 16);   // start-address field left at 0
    const int n_taps = p.dbg_skip_mma == 1 ? 0 : p.TPS;
    const int kk_n = p.KC / 16;
    const int b_step16 = b_bytes >> 4, sub16 = p.sub_off >> 4, tap16 = p.tap_off >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = tile0; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait_ns(&tempty_bar[buf], ((it >> 1) & 1) ^ 1, p.backoff_ns);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * p.MT * p.NT;
      for (int k = 0; k < p.num_ksteps; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          // The single issuing thread must stay well under one MMA time (64 cycles at N=128) per instruction, so
          // descriptors are formed by adding a pre-shifted byte offset to a per-stage base instead of re-encoding.
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t a_base = desc_base + ((sa & 0x3FFFF) >> 4);
          const uint64_t b_base = a_base + (a_bytes >> 4);
          for (int j = 0; j < n_taps; ++j) {
            const uint64_t bj = b_base + (uint32_t)(j * b_step16);
            const uint64_t a0 = a_base + (uint32_t)(j * tap16);
            const uint32_t acc = (k | j) != 0;
            if (p.MT == 2) {
              // alternate the two sub-tile accumulators: consecutive MMAs never wait for each other's accumulate
              const uint64_t a1 = a0 + (uint32_t)sub16;
              const uint32_t d0 = d_tmem, d1 = d_tmem + p.NT;
              umma_cg<CG>(d0, a0, bj, idesc, acc);
              umma_cg<CG>(d1, a1, bj, idesc, acc);
              if (kk_n >= 2) {
                umma_cg<CG>(d0, a0 + 2, bj + 2, idesc, 1);
                umma_cg<CG>(d1, a1 + 2, bj + 2, idesc, 1);
              }
              if (kk_n == 4) {
                umma_cg<CG>(d0, a0 + 4, bj + 4, idesc, 1);
                umma_cg<CG>(d1, a1 + 4, bj + 4, idesc, 1);
                umma_cg<CG>(d0, a0 + 6, bj + 6, idesc, 1);
                umma_cg<CG>(d1, a1 + 6, bj + 6, idesc, 1);
              }
            } else {
              umma_cg<CG>(d_tmem, a0, bj, idesc, acc);
              if (kk_n >= 2) umma_cg<CG>(d_tmem, a0 + 2, bj + 2, idesc, 1);
              if (kk_n == 4) {
                umma_cg<CG>(d_tmem, a0 + 4, bj + 4, idesc, 1);
                umma_cg<CG>(d_tmem, a0 + 6, bj + 6, idesc, 1);
              }
            }
          }
          umma_commit_cg<CG>(&empty_bar[stage]);
          if (k == p.num_ksteps - 1) umma_commit_cg<CG>(&tfull_bar[buf]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    }
  } else if (warp - 2 < 4 * p.MT) {
    // ------------------------------------------------------------------ epilogue (warps 2..5: sub-tile 0, 6..9: 1)
    const int my_m = (warp - 2) >> 2;       // M sub-tile owned by this warp
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;    // pixel row inside the 128-row sub-tile
    uint8_t* stg = stg_all + (warp - 2) * (32 * kStgPitch);
    // bf16 channels-last tensors only (fp32 outputs keep the per-lane path)
    const bool coalesced = p.out_f32 == nullptr && p.out_cstride == 1 && p.split_off == 0;
    // Epilogue operand prefetch.  The mask / addend tiles are the only *loads* of the epilogue; with 4 warps of
    // dependent load->use chains they would be latency bound (~8 KB in flight per SM).  Each warp therefore copies its 32
    // rows of every such tile into shared memory with cp.async BEFORE waiting for the accumulator, i.e. overlapped with
    // the MMAs of the tile (fully coalesced: 16 lanes x 16 B per row), and later reads its own row back (pitch
    // NT*2+16 B: conflict-free).
    const int e_tile_bytes = 32 * p.NT * 2;      // one operand, this warp's 32 rows
    const int n_e = p.e_mask + p.e_add;
    const bool prefetch = coalesced && p.e_bufs > 0 && n_e > 0;
    const int ew = (my_m << 2) | quarter;         // this warp's index among the (up to) 8 epilogue warps
    // layout: [depth slot][epilogue warp (4 * MT of them)][tensor][column block][32 rows][e_cols * 2 B]
    auto e_ptr = [&](int slot, int tensor) -> uint8_t* {
      return e_all + ((size_t)(slot * 4 * p.MT + ew) * n_e + tensor) * e_tile_bytes;
    };
    const bool ahead = prefetch && p.e_depth == 2;   // tiles of step i+1 are fetched while step i is processed
    // One lane asks the TMA unit for the warp's rows of every prefetched operand (NT / e_cols boxes each); rows outside
    // the tensor arrive as zeros (gate closed / nothing added; the stores of those rows are predicated anyway).
    // (the previous cp.async version spent ~45 instructions per 16 bytes on index arithmetic and shuffles: ncu)
    const int rg0 = my_m * 128 + quarter * 32;      // first row of this warp inside a tile
    const int eb1 = rg0 % p.box[0];
    const int eb2 = (rg0 / p.box[0]) % p.box[1];
    const int eb3 = (rg0 / (p.box[0] * p.box[1])) % p.box[2];
    const int eb4 = rg0 / (p.box[0] * p.box[1] * p.box[2]);
    uint32_t e_phase = 0;                           // bit s: parity to wait for on slot s
    auto issue_prefetch = [&](int slot, int g, const int (&base)[4], int ncol0) {
      if (lane == 0) {
        uint64_t* bar = &e_bar[ew * 2 + slot];
        mbar_expect_tx(bar, n_e * e_tile_bytes);
        int t = 0;
        for (int which = 0; which < 2; ++which) {
          if (!(which == 0 ? p.e_mask : p.e_add)) continue;
          uint8_t* dst = e_ptr(slot, t++);
          for (int cb = 0; cb * p.e_cols < p.NT; ++cb)
            tma_load_5d(dst + cb * (32 * p.e_cols * 2), &maps.e[which][g], bar, ncol0 + cb * p.e_cols, base[0] + eb1,
                        base[1] + eb2, base[2] + eb3, base[3] + eb4);
        }
      }
      __syncwarp();
    };
    // position of this lane's row inside a tile: the same for every tile
    const int rg_ = my_m * 128 + row;
    const int rb1 = rg_ % p.box[0];
    const int rb2 = (rg_ / p.box[0]) % p.box[1];
    const int rb3 = (rg_ / (p.box[0] * p.box[1])) % p.box[2];
    const int rb4 = rg_ / (p.box[0] * p.box[1] * p.box[2]);
    auto row_geometry = [&](const int (&base)[4], int g, int m, long long& off, bool& valid) {
      (void)m;
      const int c1 = base[0] + rb1, c2 = base[1] + rb2, c3 = base[2] + rb3, c4 = base[3] + rb4;
      valid = (c1 < p.lim[0]) && (c2 < p.lim[1]) && (c3 < p.lim[2]) && (c4 < p.lim[3]);
      off = p.out_off[g] + c1 * p.out_stride[0] + c2 * p.out_stride[1] + c3 * p.out_stride[2] + c4 * p.out_stride[3];
    };
    int it = 0;
    const uint32_t tempty_remote[2] = {CG == 2 ? mapa_u32(&tempty_bar[0], 0) : 0u,
                                       CG == 2 ? mapa_u32(&tempty_bar[1], 0) : 0u};
    // geometry of the tile decoded last (the prefetch of tile i+1 already needs it: carried into the next iteration)
    int cg = 0, cnt_ = 0, cbase[4] = {0, 0, 0, 0};
    long long coff = 0;
    bool cvalid = false;
    if (tile0 < p.total_tiles) {
      decode_tile(p, tile0, cg, cnt_, cbase);
      row_geometry(cbase, cg, my_m, coff, cvalid);
      if (ahead) issue_prefetch(0, cg, cbase, cnt_ * p.NT);   // first tile's rows
    }
    for (int tile = tile0; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int g = cg, nt = cnt_;
      int base[4] = {cbase[0], cbase[1], cbase[2], cbase[3]};
      const int ncol0 = nt * p.NT;
      const long long off = coff;
      const bool valid = cvalid;
      const int eslot = ahead ? (it & 1) : 0;
      {
        const int tn = tile + gridDim.x;
        if (tn < p.total_tiles) {
          decode_tile(p, tn, cg, cnt_, cbase);
          row_geometry(cbase, cg, my_m, coff, cvalid);
          if (ahead) issue_prefetch((it + 1) & 1, cg, cbase, cnt_ * p.NT);   // slot last read by this warp one tile ago
        }
      }
      if (!ahead && prefetch) issue_prefetch(0, g, base, ncol0);   // buffer was last read by this warp in the previous tile
      mbar_wait_ns(&tfull_bar[buf], (it >> 1) & 1, p.backoff_ns);
      tc_fence_after();
      {
        const int m = my_m;
        const uint8_t* e_mask = nullptr;
        const uint8_t* e_add = nullptr;
        if (prefetch) {
          mbar_wait(&e_bar[ew * 2 + eslot], (e_phase >> eslot) & 1u);
          e_phase ^= 1u << eslot;
          int t = 0;
          if (p.e_mask) e_mask = e_ptr(eslot, t++);
          if (p.e_add) e_add = e_ptr(eslot, t);
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (buf * p.MT + m) * p.NT;
        // rows handled by this lane in the transposed (coalesced) accesses
        long long roff[4];
        bool rvalid[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          roff[i] = __shfl_sync(0xffffffffu, off, i * 8 + (lane >> 2));
          rvalid[i] = __shfl_sync(0xffffffffu, (int)valid, i * 8 + (lane >> 2)) != 0;
        }
        int c = p.dbg_skip_mma == 4 ? p.NT : 0;   // debug 4: no epilogue work at all
        for (; c + 32 <= p.NT; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          if (coalesced && ncol0 + c + 32 <= p.c_store)
            epilogue_chunk32_coalesced(p, v, ncol0 + c, c, roff, rvalid, stg, lane, e_mask, e_add,
                                       bias_in_smem ? s_bias : nullptr);
          else epilogue_chunk<32>(p, v, ncol0 + c, valid, off);
        }
        for (; c + 16 <= p.NT; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
          epilogue_chunk<16>(p, v, ncol0 + c, valid, off);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_remote[buf]);
        else mbar_arrive(&tempty_bar[buf]);
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // rank 0's MMAs read the peer's shared memory and write its TMEM until the end
  else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

size_t conv_smem_bytes(const ConvParams& p) {
  const int rowb = p.KC * 2;
  const int stage_bytes = (p.a_bytes + p.TPS * (p.NT / (p.cta_pair ? 2 : 1)) * rowb + 1023) & ~1023;
  const int n_e = p.e_mask + p.e_add;
  return (size_t)p.stages * stage_bytes + (2 * p.stages + 4 + 16) * 8 + 16 + 1024 + 8 * 32 * 80 + 1024 +
         (size_t)p.e_bufs * (p.e_depth > 1 ? p.e_depth : 1) * n_e * 128 * (p.NT * 2 + 16) + 1024;
}

cudaError_t launch_conv_igemm(const ConvParams& p, const ConvMaps& maps, int num_sms, cudaStream_t stream) {
  const size_t smem = conv_smem_bytes(p);
  int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  if (p.cta_pair) {
    grid &= ~1;
    return launch_k(conv_igemm_kernel<2>, dim3(grid), dim3(kConvThreads), smem, stream, 2, p, maps);
  }
  return launch_k(conv_igemm_kernel<1>, dim3(grid), dim3(kConvThreads), smem, stream, 1, p, maps);
}

cudaError_t init_conv_igemm() {
  cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_igemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
}

}  // namespace fo
