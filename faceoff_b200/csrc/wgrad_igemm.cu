// Weight-gradient implicit GEMM for sm_100a (tcgen05 + TMA + TMEM), split-K over pixels.
//
//   dW[tap][m][n] = sum over pixels  P[pix][m] * Q[pix (+) shift_tap][n]
//
// P and Q are channels-last activations/gradients; the reduction (K) dimension is the pixel index,
// so both operands are "MN-major" for the tensor core: a TMA box {channels<=64, 64 pixels} lands in
// shared memory as 64 rows (pixels) of 32/64/128 B and is consumed directly through an MN-major UMMA
// descriptor (LBO = distance between 64-channel boxes, SBO = 8 rows).  The spatial shift of Q per
// filter tap is a TMA coordinate offset with hardware zero fill = the conv padding.
// (reference autograd of nn.Conv2d / ConvTranspose2d / Conv3d in models/vqvae_conv3d_latent.py:86-190.)
//
// grid = (splits, passes).  A CTA owns `taps_per_pass` taps (one 128 x NC fp32 accumulator each, all
// resident in TMEM) and a contiguous range of kpix-pixel tiles; per tile it loads P once and Q either once per
// tap, or (3x3(x3) filters, "halo") ONCE as a box of R+2 image rows whose three vertical taps are read through
// descriptors offset by one image row -- the same operand-reuse trick as conv_igemm.cu, which turns the kernel from
// TMA-bound into MMA-bound.  Narrow Q sides (NC <= 64) would pay the A-operand read of an MMA (64 cycles whatever N
// is) once per tap: there `mma_group` taps share one MMA of N = G * NC columns (see WgradParams::mma_group).  Partials go to partial[split][slot][m][n]; wgrad_finalize (elementwise.cu) reduces the
// splits in a fixed order (deterministic) and scatters into the PyTorch weight layout.
//
// Accumulation chains.  tcgen05 adds the products of an MMA to the fp32 accumulator with TRUNCATION: a sum of signed terms
// shrinks by ~1.5e-8 of its magnitude per accumulating MMA (tests/gpu_accum_bias.py: Conv3d weight gradient vs an fp64 GEMM,
// norm error -7.4e-6 at 480 MMAs per accumulator, -6.1e-5 at 3840, -2.5e-4 at 15360 = 32 clips in one launch; the fp32 sum of
// per-clip launches stays at -7.4e-6).  A forward convolution's chain is its K loop (<= 288 MMAs); a weight gradient's is
// the CTA's whole pixel range.  So the range is cut into chunks of `chain_tiles` pixel tiles (<= 4096 MMAs per accumulator):
// after each chunk the accumulators are written to their own partial slot (a "virtual split") and restarted, and the fp32
// round-to-nearest finalize adds the slots.  Cost: the pipeline drains at a chunk boundary (the epilogue warps double as Q
// producers) and the finalize reads more slots; only the launches with long chains have more than one chunk.  Measured at
// 32 clips (Conv3d: 4 chunks): norm error -2.45e-4 -> -6.2e-5, launch +2 %, step unchanged (2048: -3.1e-5, launch +7 %).
#include "common.cuh"
#include "igemm.cuh"

namespace fo {

constexpr int kWgThreads = 192;

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_igemm_kernel(const __grid_constant__ WgradParams p, const __grid_constant__ WgradMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic, not a pointer round-trip: the compiler keeps the shared address space -> LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int p_chunk_bytes = p.kpix * p.p_rowb;
  const int q_chunk_bytes = p.q_box_bytes;
  const int p_bytes = p.p_chunks * p_chunk_bytes;
  const int q_bytes = p.q_chunks * q_chunk_bytes;  // per Q load group (per tap, or per stage with halo)
  const int q_loads = p.q_loads;
  const int stage_bytes = (p_bytes + q_loads * q_bytes + 1023) & ~1023;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* done_bar = bars + 2 * p.stages;      // MMAs of a chunk complete -> epilogue
  uint64_t* tfree_bar = done_bar + 1;            // accumulators read out -> MMA warp may restart them
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tfree_bar + 1);
  // all-ones MN-major B tile [kpix pixels x 16] (32-byte rows): one extra N=16 MMA per K step turns the tensor core
  // into a column-sum unit, so the bias gradient sum_pix dy[pix][m] comes for free with the weight gradient
  uint8_t* ones_tile = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(tmem_ptr + 4) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.taps_per_pass * p.NC + (p.bias_partial != nullptr ? 16 : 0)) tmem_cols <<= 1;

  pdl_trigger();   // programmatic dependent launch: see common.cuh
  const int split = blockIdx.x;
  const int pass = blockIdx.y;
  // the column-sum MMAs are spread round-robin over the passes (pass p takes pixel tiles i with i % passes == p) so no
  // CTA becomes a straggler
  const bool do_bias = p.bias_partial != nullptr;
  const int bias_col = p.taps_per_pass * p.NC;   // TMEM column of the bias accumulator
  if (do_bias) {
    const uint32_t one2 = 0x3F803F80u;  // bf16 (1.0, 1.0)
    for (int i = threadIdx.x; i < p.kpix * 32 / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(ones_tile)[i] = make_uint4(one2, one2, one2, one2);
    fence_proxy_async();
  }
  // contiguous range of pixel tiles for this split
  const int per = (p.total_ptiles + p.splits - 1) / p.splits;
  const int t_begin = split * per;
  const int t_end = min(p.total_ptiles, t_begin + per);
  const int n_my = max(0, t_end - t_begin);

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.p);
      for (int i = 0; i < kMaxAMaps; ++i) tma_prefetch_desc(&maps.q[i]);
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, tmem_cols);
    tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    mbar_init(tfree_bar, 4);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();      // the preceding kernel has completed: global memory (TMA loads, partial stores) from here on
  // warp-uniform copy (the shuffle lets the compiler keep MMA operands in uniform registers)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const WgTap* taps = p.taps + pass * p.taps_per_pass;

  // ---------------------------------------------------------------- producers
  // One elected thread issuing every TMA operation of a stage is a serial chain of ~230 cycles per operation (tap table
  // reads, uniform-register moves, UTMALDG) plus the tile decode: measured 1.4 k cycles per 3-operation stage whatever the
  // number of stages in flight -- the 1x1 and stride-2 layers were bound by this thread, not by memory.  So the work is
  // spread: warp 0 arms the barrier and loads P, the (otherwise idle) epilogue warps 2.. load one Q box each, and the tile
  // coordinates are carried incrementally (no divisions per tile).
  const int q_role = warp - 2;   // Q box fetched by this warp (epilogue warps), if < q_loads
  const bool is_producer = (warp == 0 || (q_role >= 0 && q_role < q_loads)) && n_my > 0;
  // producer state (lane 0 of a producer warp), carried across the chunks of the CTA's range
  int pr_stage = 0;
  uint32_t pr_phase = 0;
  int idx[4] = {0, 0, 0, 0}, base[4] = {0, 0, 0, 0};
  if (is_producer && lane == 0) {
    int r = t_begin;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int c = p.tile_cnt[d];
      idx[d] = r % c;
      r /= c;
      base[d] = idx[d] * p.tile_step[d];
    }
  }
  // loads the pixel tiles [t0, t1) of this CTA's range
  auto produce = [&](int t0, int t1) {
    if (is_producer && lane == 0) {
      const int p_cw = p.p_rowb / 2, q_cw = p.q_rowb / 2;
      const WgTap w = taps[(warp == 0 ? 0 : q_role) * p.taps_per_load];
      const CUtensorMap* qmap = &maps.q[w.map];
      const int q_off = p_bytes + (warp == 0 ? 0 : q_role) * q_bytes;
      const int tx_bytes = p_bytes + q_loads * q_bytes;
      for (int t = t0; t < t1; ++t) {
        mbar_wait_ns(&empty_bar[pr_stage], pr_phase ^ 1, p.backoff_ns);
        uint8_t* sp = smem + (size_t)pr_stage * stage_bytes;
        if (warp == 0) {
          mbar_expect_tx(&full_bar[pr_stage], tx_bytes);
          for (int c = 0; c < p.p_chunks; ++c)
            tma_load_5d(sp + c * p_chunk_bytes, &maps.p, &full_bar[pr_stage], p.p_c0 + c * p_cw, base[0], base[1], base[2],
                        base[3]);
        } else {
          uint8_t* sq = sp + q_off;
          for (int c = 0; c < p.q_chunks; ++c)
            tma_load_5d(sq + c * q_chunk_bytes, qmap, &full_bar[pr_stage], w.c0 + c * q_cw, base[0] + w.d1, base[1] + w.d2,
                        base[2] + w.d3, base[3]);
        }
        if (++pr_stage == p.stages) { pr_stage = 0; pr_phase ^= 1; }
        // next tile: odometer over the tile grid
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          if (++idx[d] < p.tile_cnt[d]) { base[d] += p.tile_step[d]; break; }
          idx[d] = 0;
          base[d] = 0;
        }
      }
    }
    __syncwarp();
  };
  if (warp == 0) {
    produce(t_begin, t_end);   // warp 0 only produces (it runs ahead over chunk boundaries; the stage ring throttles it)
  } else if (warp == 1) {
    // G taps per MMA (N = G * NC): the taps' Q data lie tap_stride bytes apart (inside a halo box: one image row; else
    // one Q box), which the MN-major descriptor takes as the stride between its NC-wide chunks
    const int G = p.mma_group, tpl = p.taps_per_load;
    const uint32_t tap_stride = tpl > 1 ? (uint32_t)p.q_tap_off : (uint32_t)q_bytes;
    const uint32_t idesc = make_idesc_bf16(128, p.NC * G, 1, 1);
    const uint64_t a_desc_base = make_smem_desc(0, p.p_rowb, p.p_chunks > 1 ? p_chunk_bytes : 0);
    const uint64_t b_desc_base = make_smem_desc(0, p.q_rowb, p.q_chunks > 1 ? q_chunk_bytes : G > 1 ? tap_stride : 0);
    const uint32_t qload16 = q_bytes >> 4, qtap16 = p.q_tap_off >> 4, pk16 = (16 * p.p_rowb) >> 4, qk16 = (16 * p.q_rowb) >> 4;
    const int n_taps = p.dbg_skip_mma ? 0 : p.taps_per_pass;
    const int kk_n = p.kpix / 16;
    const uint32_t idesc_bias = make_idesc_bf16(128, 16, 1, 1);
    const uint64_t ones_desc = make_smem_desc(smem_u32(ones_tile), 32, 0);
    int stage = 0;
    uint32_t phase = 0;
    int chunk = 0;
    for (int c0 = 0; c0 < n_my; c0 += p.chain_tiles, ++chunk) {
      const int c1 = min(n_my, c0 + p.chain_tiles);
      if (chunk > 0) {   // the epilogue has read the previous chunk's accumulators out
        mbar_wait_ns(tfree_bar, (chunk - 1) & 1, p.backoff_ns);
        tc_fence_after();
      }
      uint32_t bias_started = 0;
      for (int i = c0; i < c1; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const bool bias_tile = do_bias && (i % p.passes) == pass;
        if (elect_one()) {
          const uint32_t sp = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t a0 = a_desc_base + ((sp & 0x3FFFF) >> 4);
          const uint64_t b0 = b_desc_base + (((sp + p_bytes) & 0x3FFFF) >> 4);
          for (int tp = 0; tp < n_taps; tp += G) {
            const uint64_t bt = b0 + (uint32_t)((tp / tpl) * qload16 + (tp % tpl) * qtap16);
            const uint32_t dt = tmem_base + tp * p.NC;
            // 16 pixels per MMA = two 8-row swizzle groups = 16 * rowb bytes
            umma_bf16(dt, a0, bt, idesc, i != c0);
            umma_bf16(dt, a0 + pk16, bt + qk16, idesc, 1);
            umma_bf16(dt, a0 + 2 * pk16, bt + 2 * qk16, idesc, 1);
            umma_bf16(dt, a0 + 3 * pk16, bt + 3 * qk16, idesc, 1);
            if (kk_n == 8) {
              umma_bf16(dt, a0 + 4 * pk16, bt + 4 * qk16, idesc, 1);
              umma_bf16(dt, a0 + 5 * pk16, bt + 5 * qk16, idesc, 1);
              umma_bf16(dt, a0 + 6 * pk16, bt + 6 * qk16, idesc, 1);
              umma_bf16(dt, a0 + 7 * pk16, bt + 7 * qk16, idesc, 1);
            }
          }
          if (bias_tile) {
            for (int kk = 0; kk < kk_n; ++kk)
              umma_bf16(tmem_base + bias_col, a0 + kk * pk16, ones_desc + kk * 32, idesc_bias, bias_started | (uint32_t)(kk != 0));
          }
          umma_commit(&empty_bar[stage]);
          if (i == c1 - 1) umma_commit(done_bar);
        }
        if (bias_tile) bias_started = 1;
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // Q producer (lane 0 of the first q_loads of these warps) + epilogue, chunk by chunk:
    //   TMEM -> partial[split * n_flush + chunk][pass*tpp + tp][m][n]
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    for (int chunk = 0; chunk < p.n_flush; ++chunk) {
      const int c0 = min(n_my, chunk * p.chain_tiles), c1 = min(n_my, c0 + p.chain_tiles);
      const bool live = c1 > c0;     // chunks past the end of a short range store zeros
      produce(t_begin + c0, t_begin + c1);
      if (live) {
        mbar_wait_ns(done_bar, chunk & 1, p.backoff_ns);
        tc_fence_after();
      }
      const size_t slot = (size_t)split * p.n_flush + chunk;
      for (int tp = 0; tp < p.taps_per_pass; ++tp) {
        const int tap = pass * p.taps_per_pass + tp;
        float* dst = p.partial + ((slot * (p.passes * p.taps_per_pass) + tap) * p.MC + row) * p.NC;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + tp * p.NC;
        for (int c = 0; c < p.NC; c += 16) {
          uint32_t v[16];
          if (live) {
            tmem_ld16(taddr + c, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
          }
          if (row < p.MC) {
            float4* o = reinterpret_cast<float4*>(dst + c);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              o[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
          }
        }
      }
      if (do_bias) {
        uint32_t v[16];
        // did this chunk issue a column-sum MMA?  (tiles i of the CTA's range with i % passes == pass)
        const int first = c0 + ((pass - c0 % p.passes) + p.passes) % p.passes;
        if (live && first < c1) {
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + bias_col, v);
          tmem_ld_wait();
        } else {
          v[0] = 0u;
        }
        if (row < p.MC) p.bias_partial[((size_t)pass * p.splits * p.n_flush + slot) * p.MC + row] = __uint_as_float(v[0]);
      }
      if (live) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tfree_bar);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

size_t wgrad_smem_bytes(const WgradParams& p) {
  const int p_bytes = p.p_chunks * p.kpix * p.p_rowb;
  const int q_bytes = p.q_chunks * p.q_box_bytes;
  const int stage_bytes = (p_bytes + p.q_loads * q_bytes + 1023) & ~1023;
  return (size_t)p.stages * stage_bytes + (2 * p.stages + 2) * 8 + 16 + 1024 + 1024 + 128 * 32;
}

cudaError_t launch_wgrad_igemm(const WgradParams& p, const WgradMaps& maps, cudaStream_t stream) {
  dim3 grid(p.splits, p.passes);
  return launch_k(wgrad_igemm_kernel, grid, dim3(kWgThreads), wgrad_smem_bytes(p), stream, 1, p, maps);
}

cudaError_t init_wgrad_igemm() {
  return cudaFuncSetAttribute(wgrad_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
}

}  // namespace fo
