// C ABI (include/faceoff_b200.h) + host-side planners that turn a convolution description into the generic
// implicit-GEMM launch (tile boxes, K-step table, TMA tensor maps, packed-weight order).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/faceoff_b200.h"
#include "kernels.h"

using namespace fo;

// ------------------------------------------------------------------------------------------ errors / init
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return fail(FO_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;
namespace fo { int g_pdl = 1; }   // programmatic dependent launch (common.cuh); FO_PDL=0 turns it off
static int g_pair_ok = 1;   // clusters of two CTAs can be scheduled on every TPC
static std::once_flag g_once;
static int g_init_rc = FO_ERR_NO_DEVICE;
static char g_init_err[512] = "fo_init not called";

static void do_init() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    snprintf(g_init_err, sizeof(g_init_err), "no CUDA device: faceoff_b200 has no CPU path");
    g_init_rc = FO_ERR_NO_DEVICE;
    return;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  if (prop.major != 10) {
    snprintf(g_init_err, sizeof(g_init_err), "device %s is sm_%d%d; kernels are built for sm_100a only", prop.name,
             prop.major, prop.minor);
    g_init_rc = FO_ERR_NO_DEVICE;
    return;
  }
  g_num_sms = prop.multiProcessorCount;
  {
    const char* e = getenv("FO_PDL");
    if (e != nullptr) fo::g_pdl = atoi(e) != 0;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    snprintf(g_init_err, sizeof(g_init_err), "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    g_init_rc = FO_ERR_CUDA;
    return;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  if ((e = init_conv_igemm()) != cudaSuccess || (e = init_wgrad_igemm()) != cudaSuccess ||
      (e = init_vq()) != cudaSuccess) {
    snprintf(g_init_err, sizeof(g_init_err), "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    g_init_rc = FO_ERR_CUDA;
    return;
  }
  g_init_rc = FO_OK;
  g_init_err[0] = 0;
}

extern "C" const char* fo_last_error(void) { return g_err; }
extern "C" int fo_version(void) { return 100; }
extern "C" int fo_init(void) {
  std::call_once(g_once, do_init);
  if (g_init_rc != FO_OK) return fail(g_init_rc, "%s", g_init_err);
  return FO_OK;
}
#define REQUIRE_INIT()                       \
  do {                                       \
    int _rc = fo_init();                     \
    if (_rc != FO_OK) return _rc;            \
  } while (0)

// ------------------------------------------------------------------------------------------ tensor maps
static CUtensorMapSwizzle swizzle_for(int rowb) {
  return rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}
// bf16 tensor, `rank` dims (dim 0 contiguous), strides in elements for dims 1..rank-1
static int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_el,
                      const uint32_t* box, int rowb) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_el[i] * 2;
    if (box[i] > 256 || box[i] == 0) return fail(FO_ERR_INVALID, "TMA box dim %d = %u out of range", i, box[i]);
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(FO_ERR_INVALID, "tensor base not 16-byte aligned");
  for (int i = 0; i + 1 < rank; ++i)
    if (gs[i] % 16 != 0) return fail(FO_ERR_INVALID, "TMA stride %d (%llu B) not a multiple of 16", i, (unsigned long long)gs[i]);
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(rowb), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(FO_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
                (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
                (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
                rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
  }
  return FO_OK;
}

// ------------------------------------------------------------------------------------------ tile boxes
static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
static int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }
// choose box extents (powers of two, product == positions) over the tile domain ext[0..3]
static void choose_box(const int ext[4], int positions, int box[4], int max_first = 128) {
  int rem = positions;
  int last_used = 0;
  for (int d = 0; d < 4; ++d) {
    int b = 1;
    if (ext[d] > 1 && rem > 1) {
      const int pf = pow2_floor(ext[d]);
      b = (ext[d] % pf == 0) ? pf : pow2_ceil(ext[d]);
      if (b > rem) b = rem;
      if (d == 0 && b > max_first) b = max_first;   // keep >= 2 image rows per 256-position tile (halo stages)
      last_used = d;
    }
    box[d] = b;
    rem /= b;
  }
  if (rem > 1) box[last_used] *= rem;  // tiny tensors: over-size the box, TMA zero-fills / stores are predicated
}

static const int kDownD[4] = {-1, 0, 0, 1};   // 4x4 s2 p1: input index 2*o + k - 1  ->  half-res offset
static const int kDownPar[4] = {1, 0, 1, 0};  //                                        and parity

// ------------------------------------------------------------------------------------------ conv planner
struct ConvPlan {
  ConvParams p;
  ConvMaps maps;
  PackParams pack;
  int npad, ktot;
};

static int plan_conv(const fo_conv_t* c, ConvPlan* out, bool need_maps) {
  ConvParams& p = out->p;
  memset(&p, 0, sizeof(p));
  memset(&out->pack, 0, sizeof(out->pack));
  if (c->n_src < 1 || c->n_src > kMaxAMaps) return fail(FO_ERR_INVALID, "n_src must be 1..%d", kMaxAMaps);
  const bool s1 = c->form == FO_FORM_S1 || c->form == FO_FORM_S1_DGRAD;
  if (!s1 && c->form != FO_FORM_DOWN && c->form != FO_FORM_UP) return fail(FO_ERR_INVALID, "bad form %d", c->form);
  if (c->ndim != 2 && !(c->ndim == 3 && s1)) return fail(FO_ERR_INVALID, "ndim %d unsupported for form %d", c->ndim, c->form);
  if (s1 && c->ksize != 1 && c->ksize != 3) return fail(FO_ERR_INVALID, "ksize %d unsupported", c->ksize);
  if (c->form == FO_FORM_DOWN && ((c->h | c->w) & 1)) return fail(FO_ERR_INVALID, "stride-2 conv needs even h, w");
  // K chunk
  int kc = 64;
  for (int s = 0; s < c->n_src; ++s) {
    const fo_src_t& src = c->src[s];
    if (src.cs % 8 != 0 || src.c_off % 8 != 0) return fail(FO_ERR_INVALID, "source storage channels must be multiples of 8");
    const int cpad = (src.c + 15) / 16 * 16;
    if (src.c_off + cpad > src.cs) return fail(FO_ERR_INVALID, "source slice [%d,+%d) exceeds storage %d", src.c_off, cpad, src.cs);
    while (cpad % kc != 0) kc /= 2;
  }
  {
    const char* e = getenv("FO_FORCE_KC");   // experiments only
    if (e && (atoi(e) == 32 || atoi(e) == 16) && atoi(e) < kc) kc = atoi(e);
  }
  if (kc < 16) return fail(FO_ERR_INVALID, "channel counts must be multiples of 16 after padding");
  p.KC = kc;
  const int rowb = kc * 2;
  // N tiling
  const int cout_pad = (c->cout + 15) / 16 * 16;
  // N tiles of 128 columns whenever cout is a multiple of 128: two M=128 sub-tiles x 128 columns fill TMEM (2 x 2 x 128)
  // and allow the halo / 256-position tiling below; the A tile is re-read per N tile from L2.
  if (cout_pad <= 128) { p.NT = cout_pad; p.n_tiles = 1; }
  else if (cout_pad % 128 == 0) { p.NT = 128; p.n_tiles = cout_pad / 128; }
  else if (cout_pad <= 256) { p.NT = cout_pad; p.n_tiles = 1; }
  else return fail(FO_ERR_INVALID, "cout %d unsupported", c->cout);
  out->npad = p.NT * p.n_tiles;

  // tile domain (map dims 1..4) and output addressing
  int ext[4];
  const long long ocs = c->out_cs;
  const bool nchw = c->out_f32 != nullptr && c->out_f32_nchw;
  if (nchw && (c->out_bf16 || c->out_relu || c->mask || c->addend || c->ndim != 2))
    return fail(FO_ERR_INVALID, "NCHW fp32 output cannot be combined with channels-last outputs");
  p.groups = 1;
  if (s1 || c->form == FO_FORM_UP) {
    if (c->ndim == 3) { ext[0] = c->w; ext[1] = c->h; ext[2] = c->d; ext[3] = c->n; }
    else { ext[0] = c->w; ext[1] = c->h; ext[2] = c->n; ext[3] = 1; }
  } else {  // DOWN: (wx, py, hy, f)
    ext[0] = c->w / 2; ext[1] = 1; ext[2] = c->h / 2; ext[3] = c->n;
  }
  // CTA tile: 128 positions, or 256 (two M=128 sub-tiles sharing every B tile) when the problem is large enough;
  // 3x3(x3) stride-1 filters additionally read their three vertical taps from one halo'd A box (see conv_igemm.cu).
  int box[4];
  p.MT = 1;
  p.TPS = 1;
  bool halo = false;
  const int smem_avail = kMaxDynSmem - 3328 - 1024 - 8 * 32 * 80;
  const int n_e_plan = (c->mask != nullptr ? 1 : 0) + (c->addend != nullptr ? 1 : 0);
  const bool want_prefetch = n_e_plan > 0 && !nchw && c->out_f32 == nullptr && p.NT % 32 == 0;
  const int e_one_plan = n_e_plan * 128 * (p.NT * 2 + 16);
  // candidate modes, best first: 256-position tiles (MT = 2) when the problem is large, else 128; each with halo
  // stages when the filter is 3x3(x3) and a tile covers whole image rows of one frame.  A mode is skipped when the
  // epilogue's operand prefetch (mask / addend present) would not fit next to two pipeline stages.
  struct Mode { bool ok; int halo; int box[4]; } modes[3] = {};   // halo = extra image rows in the A box (0, 1, 2)
  for (int mt = 1; mt <= (p.NT <= 128 ? 2 : 1); ++mt) {
    Mode& md = modes[mt];
    // halo-eligible filters prefer 4 image rows per tile (box of 6 rows: 1.5x read amplification instead of 2x)
    const bool halo_form = (s1 && c->ksize == 3) || (c->form == FO_FORM_UP && c->ndim == 2);
    choose_box(ext, 128 * mt, md.box, halo_form ? 32 * mt : 128);
    long long tiles = 1;
    bool exact = true;
    for (int d = 0; d < 4; ++d) {
      tiles *= (ext[d] + md.box[d] - 1) / md.box[d];
      if (md.box[d] > 1 && md.box[d] > pow2_ceil(ext[d])) exact = false;
    }
    const int sms = g_num_sms > 0 ? g_num_sms : 148;
    md.ok = mt == 1 || (exact && tiles * p.n_tiles >= 2LL * sms);
    const int tap_off = md.box[0] * rowb;
    const bool rows_ok = md.box[2] == 1 && md.box[3] == 1 && md.box[0] * md.box[1] == 128 * mt && tap_off % 1024 == 0 &&
                         md.box[0] <= ext[0] && md.box[1] <= ext[1];
    // 3x3(x3): the three vertical taps share one box of R+2 rows; 4x4 stride-2 transposed: the two vertical taps of a
    // sub-pixel group share one box of R+1 rows
    md.halo = !rows_ok ? 0 : (s1 && c->ksize == 3) ? 2 : (c->form == FO_FORM_UP && c->ndim == 2) ? 1 : 0;
    const int stage_b = (md.box[0] * (md.box[1] + md.halo) * rowb + (md.halo + 1) * p.NT * rowb + 1023) & ~1023;
    if (md.halo && 2 * stage_b > smem_avail) md.halo = 0;   // halo box too large: plain stages
  }
  int pick = modes[2].ok ? 2 : 1;
  {
    const char* e = getenv("FO_FORCE_MT");   // experiments only
    if (e && atoi(e) == 2 && modes[2].ok) pick = 2;
    if (e && atoi(e) == 1) pick = 1;
  }
  p.MT = pick;
  const int halo_extra = modes[pick].halo;
  halo = halo_extra > 0;
  p.TPS = halo_extra + 1;
  for (int d = 0; d < 4; ++d) box[d] = modes[pick].box[d];
  (void)n_e_plan; (void)want_prefetch; (void)e_one_plan;
  for (int d = 0; d < 4; ++d) {
    p.box[d] = box[d];
    p.tile_step[d] = box[d];
    p.tile_cnt[d] = (ext[d] + box[d] - 1) / box[d];
    p.lim[d] = ext[d];
  }
  {   // the 32 rows of one epilogue warp form an aligned sub-box of the tile (all extents are powers of two)
    int rem = 32;
    for (int d = 0; d < 4; ++d) {
      const int sdim = box[d] < rem ? box[d] : rem;
      p.e_box[d] = sdim;
      rem /= sdim;
    }
    p.e_cols = p.NT % 64 == 0 ? 64 : 32;
    if (rem != 1) p.e_cols = 0;   // (cannot happen with power-of-two boxes) -> no operand prefetch
  }
  p.sub_off = 128 * rowb;
  p.tap_off = halo ? box[0] * rowb : 0;
  p.a_bytes = (halo ? box[0] * (box[1] + halo_extra) : 128 * p.MT) * rowb;
  // split the A box over dim 2 (image rows) into several TMA operations
  const int a_rows = box[1] + halo_extra;
  p.a_ops = 1;
  {
    const char* e = getenv("FO_A_OPS");
    int want = e ? atoi(e) : 1;
    if (want > 1 && a_rows % want == 0 && box[2] == 1 && box[3] == 1 && ((a_rows / want) * box[0] * rowb) % 1024 == 0)
      p.a_ops = want;
    const char* sk = getenv("FO_SKIP_MMA");
    p.dbg_skip_mma = sk ? atoi(sk) : 0;
    const char* bo = getenv("FO_BACKOFF");
    p.backoff_ns = bo ? atoi(bo) : 0;
  }
  p.a_op_rows = a_rows / p.a_ops;
  if (s1) {
    const long long W = c->w, H = c->h, D = c->d;
    if (c->ndim == 3) {
      p.out_stride[0] = ocs; p.out_stride[1] = W * ocs; p.out_stride[2] = H * W * ocs; p.out_stride[3] = D * H * W * ocs;
    } else if (!nchw) {
      p.out_stride[0] = ocs; p.out_stride[1] = W * ocs; p.out_stride[2] = H * W * ocs; p.out_stride[3] = 0;
    } else {
      p.out_stride[0] = 1; p.out_stride[1] = W; p.out_stride[2] = (long long)c->cout * H * W; p.out_stride[3] = 0;
      p.out_cstride = (int)(H * W);
    }
  } else if (c->form == FO_FORM_DOWN) {
    const long long Wo = c->w / 2, Ho = c->h / 2;
    if (!nchw) { p.out_stride[0] = ocs; p.out_stride[1] = 0; p.out_stride[2] = Wo * ocs; p.out_stride[3] = Ho * Wo * ocs; }
    else { p.out_stride[0] = 1; p.out_stride[1] = 0; p.out_stride[2] = Wo; p.out_stride[3] = (long long)c->cout * Ho * Wo; p.out_cstride = (int)(Ho * Wo); }
  } else {  // UP
    const long long Wo = 2LL * c->w, Ho = 2LL * c->h;
    p.groups = 4;
    if (!nchw) {
      p.out_stride[0] = 2 * ocs; p.out_stride[1] = 2 * Wo * ocs; p.out_stride[2] = Ho * Wo * ocs; p.out_stride[3] = 0;
      for (int g = 0; g < 4; ++g) p.out_off[g] = ((g >> 1) * Wo + (g & 1)) * ocs;
    } else {
      p.out_stride[0] = 2; p.out_stride[1] = 2 * Wo; p.out_stride[2] = (long long)c->cout * Ho * Wo; p.out_stride[3] = 0;
      for (int g = 0; g < 4; ++g) p.out_off[g] = (g >> 1) * Wo + (g & 1);
      p.out_cstride = (int)(Ho * Wo);
    }
  }
  if (!nchw) p.out_cstride = 1;
  p.c_store = nchw ? c->cout : (int)(ocs < out->npad ? ocs : out->npad);
  if (!nchw && (c->out_bf16 || c->out_relu || c->out_f32) && (ocs % 8 != 0))
    return fail(FO_ERR_INVALID, "out_cs must be a multiple of 8");

  // K steps (= pipeline stages) and the matching packed-weight column blocks (TPS per stage)
  int nk = 0;  // stages per group
  KStep* ks = p.ksteps;
  PackStep* ps = out->pack.steps;
  int n_stage = 0, n_pack = 0;
  bool ok = true;
  auto push_stage = [&](int map, int c0, int d1, int d2, int d3) {
    if (n_stage >= kMaxKSteps) { ok = false; return; }
    ks[n_stage].c0 = c0; ks[n_stage].d1 = (int8_t)d1; ks[n_stage].d2 = (int8_t)d2; ks[n_stage].d3 = (int8_t)d3;
    ks[n_stage].map = (uint8_t)map;
    ++n_stage;
  };
  auto push_pack = [&](int tap, int wk0, int valid) {
    if (n_pack >= kMaxKSteps) { ok = false; return; }
    ps[n_pack].tap = (int16_t)tap; ps[n_pack].wk0 = (int16_t)wk0; ps[n_pack].valid = (int16_t)valid; ps[n_pack].pad = 0;
    ++n_pack;
  };
  // all channel chunks of source `map` for one stage; taps[] lists the TPS filter taps served by the stage
  auto chunks = [&](int map, int pix_c0, int d1, int d2, int d3, const int* taps) {
    const fo_src_t& src = c->src[map];
    int wbase = 0;
    for (int s = 0; s < map; ++s) wbase += c->src[s].c;
    const int cpad = (src.c + 15) / 16 * 16;
    for (int ch = 0; ch < cpad; ch += kc) {
      const int valid = src.c - ch < kc ? src.c - ch : kc;
      push_stage(map, pix_c0 + src.c_off + ch, d1, d2, d3);
      for (int j = 0; j < p.TPS; ++j) push_pack(taps[j], wbase + ch, valid);
    }
  };
  if (s1) {
    const int k = c->ksize, pad = (k - 1) / 2;
    const int kd_n = c->ndim == 3 ? k : 1;
    const int sign = c->form == FO_FORM_S1 ? 1 : -1;
    for (int kd = 0; kd < kd_n; ++kd)
      for (int kh = 0; kh < (halo ? 1 : k); ++kh)
        for (int kw = 0; kw < k; ++kw) {
          const int dw = sign * (kw - pad), dd = c->ndim == 3 ? sign * (kd - pad) : 0;
          if (halo) {
            // box starts one image row above the tile; sub-box j holds the rows for vertical offset j-1
            int taps[3];
            for (int j = 0; j < 3; ++j) taps[j] = (kd * k + (sign > 0 ? j : 2 - j)) * k + kw;
            for (int s = 0; s < c->n_src; ++s) chunks(s, 0, dw, -1, dd, taps);
          } else {
            const int tap = (kd * k + kh) * k + kw;
            for (int s = 0; s < c->n_src; ++s) chunks(s, 0, dw, sign * (kh - pad), dd, &tap);
          }
        }
    nk = n_stage;
  } else if (c->form == FO_FORM_DOWN) {
    for (int ky = 0; ky < 4; ++ky)
      for (int kx = 0; kx < 4; ++kx)
        for (int s = 0; s < c->n_src; ++s) {
          const int tap = ky * 4 + kx;
          chunks(s, kDownPar[kx] * c->src[s].cs, kDownD[kx], kDownPar[ky], kDownD[ky], &tap);
        }
    nk = n_stage;
  } else {  // UP: out[2i-1+k] += in[i] w[k]; output parity p: k in {1,3} (p=0: i = h, h-1) or {0,2} (p=1: i = h+1, h)
    static const int kk[2][2] = {{1, 3}, {0, 2}};
    static const int dd[2][2] = {{0, -1}, {1, 0}};
    for (int g = 0; g < 4; ++g) {
      const int py = g >> 1, px = g & 1;
      if (halo) {
        // vertical pair of the group in one box: rows start at min(dd[py]); slot 0 = upper input row
        const int start = py == 0 ? -1 : 0;
        int taps[2];
        for (int b = 0; b < 2; ++b) {
          // slot j reads input row offset start + j: py=0 -> (-1: ky=3, 0: ky=1); py=1 -> (0: ky=2, +1: ky=0)
          taps[0] = (py == 0 ? 3 : 2) * 4 + kk[px][b];
          taps[1] = (py == 0 ? 1 : 0) * 4 + kk[px][b];
          for (int s = 0; s < c->n_src; ++s) chunks(s, 0, dd[px][b], start, 0, taps);
        }
      } else {
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b)
            for (int s = 0; s < c->n_src; ++s) {
              const int tap = kk[py][a] * 4 + kk[px][b];
              chunks(s, 0, dd[px][b], dd[py][a], 0, &tap);
            }
      }
      if (g == 0) nk = n_stage;
    }
  }
  if (!ok) return fail(FO_ERR_INVALID, "too many K steps (> %d)", kMaxKSteps);
  p.num_ksteps = nk;
  out->ktot = n_pack * kc;
  // CTA pairs (cta_group::2): two neighbouring position tiles share every B tile; needs an even tile count per
  // (group, n-tile) so that both CTAs of a pair always have work with the same K-step list
  p.pos_tiles = p.tile_cnt[0] * p.tile_cnt[1] * p.tile_cnt[2] * p.tile_cnt[3];
  p.fd_pos = make_fastdiv(p.pos_tiles);
  p.fd_nt = make_fastdiv(p.n_tiles);
  for (int d = 0; d < 4; ++d) p.fd_cnt[d] = make_fastdiv(p.tile_cnt[d]);
  {
    // worth it for the long-K layers (3x3 / 4x4 / 3x3x3 filters over >= 64 channels); the short 1x1 layers are bound by
    // their epilogue and only pay for the pair's extra synchronisation.  FO_CTA_PAIR=0 disables, =2 forces (debug).
    const char* e = getenv("FO_CTA_PAIR");
    const int want = e ? atoi(e) : 1;
    const bool heavy = p.num_ksteps * p.TPS * kc >= 576;
    p.cta_pair = (want == 2 || (want == 1 && heavy)) && p.pos_tiles % 2 == 0 && p.NT % 32 == 0 &&
                 g_num_sms % 2 == 0 && g_pair_ok && (p.dbg_skip_mma == 0 || p.dbg_skip_mma >= 3);
  }
  const int stage_bytes = (p.a_bytes + p.TPS * (p.NT / (p.cta_pair ? 2 : 1)) * rowb + 1023) & ~1023;
  // shared memory: pipeline stages + (as far as it fits next to two stages) the epilogue's prefetched addend / mask rows
  const int avail = kMaxDynSmem - 3328 - 1024 - 8 * 32 * 80;
  const int e_tensor = p.MT * 128 * (p.NT * 2 + 16);   // one operand, all sub-tiles
  p.e_bufs = 0; p.e_mask = 0; p.e_add = 0;
  if (!nchw && c->out_f32 == nullptr && p.NT % 32 == 0 && !c->split_out && p.e_cols != 0 && c->out_cs % 8 == 0) {
    int room = avail - 2 * stage_bytes;
    if (c->addend != nullptr && room >= e_tensor) { p.e_add = 1; room -= e_tensor; }
    if (c->mask != nullptr && room >= e_tensor) { p.e_mask = 1; room -= e_tensor; }
    if (p.e_add || p.e_mask) p.e_bufs = p.MT;
    // short-K layers (1x1, 32-channel inputs): the tile's MMAs are too short to hide the fetch, so fetch one tile ahead
    // when a second copy fits next to two pipeline stages
    const char* ed = getenv("FO_E_DEPTH");
    const int want_depth = ed ? atoi(ed) : 2;
    if (p.e_bufs && want_depth == 2 && p.NT >= 64 && room >= (p.e_add + p.e_mask) * e_tensor) p.e_depth = 2;   // two stages stay
  }
  if (p.e_depth < 1) p.e_depth = 1;
  int stages = (avail - (p.e_add + p.e_mask) * e_tensor * p.e_depth) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return fail(FO_ERR_INVALID, "tile does not fit in shared memory");
  p.stages = stages;
  p.total_tiles = p.groups * p.n_tiles * p.tile_cnt[0] * p.tile_cnt[1] * p.tile_cnt[2] * p.tile_cnt[3];

  // weight pack description
  PackParams& pk = out->pack;
  pk.npad = out->npad; pk.ktot = out->ktot; pk.kc = kc; pk.cout = c->cout;
  pk.taps = s1 ? (c->ndim == 3 ? c->ksize * c->ksize * c->ksize : c->ksize * c->ksize) : 16;

  // epilogue pointers
  p.bias = c->bias;
  p.mask = (const __nv_bfloat16*)c->mask;
  p.addend = (const __nv_bfloat16*)c->addend;
  p.out_bf16 = (__nv_bfloat16*)c->out_bf16;
  p.out_relu = (__nv_bfloat16*)c->out_relu;
  p.out_f32 = c->out_f32;
  p.relu_f32 = c->relu_f32;
  p.acc_f32 = c->out_f32_accumulate;
  if (p.acc_f32 && c->out_f32 == nullptr) return fail(FO_ERR_INVALID, "out_f32_accumulate needs out_f32");
  p.split_off = 0;
  if (c->split_out) {
    if (nchw || (c->out_cs & 31) != 0 || c->out_cs / 2 < out->npad)
      return fail(FO_ERR_INVALID, "split_out needs channels-last outputs with out_cs = 2 * padded cout");
    p.split_off = c->out_cs / 2;
  }

  if (!need_maps) return FO_OK;
  // tensor maps
  for (int s = 0; s < c->n_src; ++s) {
    const fo_src_t& src = c->src[s];
    uint64_t dims[5], str[5];
    uint32_t bx[5];
    const uint64_t cs = src.cs;
    if (c->form == FO_FORM_DOWN) {
      dims[0] = 2 * cs; dims[1] = c->w / 2; dims[2] = 2; dims[3] = c->h / 2; dims[4] = c->n;
      str[0] = 1; str[1] = 2 * cs; str[2] = (uint64_t)c->w * cs; str[3] = 2ULL * c->w * cs; str[4] = (uint64_t)c->h * c->w * cs;
    } else if (c->ndim == 3) {
      dims[0] = cs; dims[1] = c->w; dims[2] = c->h; dims[3] = c->d; dims[4] = c->n;
      str[0] = 1; str[1] = cs; str[2] = (uint64_t)c->w * cs; str[3] = (uint64_t)c->h * c->w * cs; str[4] = (uint64_t)c->d * c->h * c->w * cs;
    } else {
      dims[0] = cs; dims[1] = c->w; dims[2] = c->h; dims[3] = c->n; dims[4] = 1;
      str[0] = 1; str[1] = cs; str[2] = (uint64_t)c->w * cs; str[3] = (uint64_t)c->h * c->w * cs; str[4] = (uint64_t)c->n * c->h * c->w * cs;
    }
    bx[0] = kc; bx[1] = box[0]; bx[2] = p.a_op_rows; bx[3] = box[2]; bx[4] = box[3];
    int rc = encode_map(&out->maps.a[s], src.ptr, 5, dims, str, bx, rowb);
    if (rc != FO_OK) return rc;
  }
  for (int s = c->n_src; s < kMaxAMaps; ++s) out->maps.a[s] = out->maps.a[0];
  // epilogue operand maps (mask / addend have the layout of the bf16 outputs): tile domain = map dims 1..4, element
  // strides = the epilogue's out_stride; UP form: one map per sub-pixel group (base shifted by out_off[g])
  for (int which = 0; which < 2; ++which) {
    const void* base = which == 0 ? c->mask : c->addend;
    const bool on = p.e_bufs > 0 && (which == 0 ? p.e_mask : p.e_add);
    for (int g = 0; g < kMaxGroups; ++g) {
      if (!on || g >= p.groups) { out->maps.e[which][g] = out->maps.a[0]; continue; }
      uint64_t dims[5], str[5];
      uint32_t bx[5];
      dims[0] = (uint64_t)ocs; str[0] = 1; bx[0] = (uint32_t)p.e_cols;
      uint64_t fallback = (uint64_t)ocs;
      for (int d = 0; d < 4; ++d) {
        dims[d + 1] = (uint64_t)p.lim[d];
        // a dummy dimension (extent 1) carries stride 0 in out_stride: give it any legal stride
        str[d + 1] = p.out_stride[d] > 0 ? (uint64_t)p.out_stride[d] : fallback;
        if (p.out_stride[d] > 0) fallback = (uint64_t)p.out_stride[d] * (uint64_t)p.lim[d];
        bx[d + 1] = (uint32_t)p.e_box[d];
      }
      int rc = encode_map(&out->maps.e[which][g], (const uint8_t*)base + (size_t)p.out_off[g] * 2, 5, dims, str, bx,
                          p.e_cols * 2);
      if (rc != FO_OK) return rc;
    }
  }
  {
    uint64_t dims[2] = {(uint64_t)out->ktot, (uint64_t)out->npad};
    uint64_t str[2] = {1, (uint64_t)out->ktot};
    uint32_t bx[2] = {(uint32_t)kc, (uint32_t)(p.NT / (p.cta_pair ? 2 : 1))};   // a pair's CTAs load half a B tile each
    int rc = encode_map(&out->maps.b, c->wpacked, 2, dims, str, bx, rowb);
    if (rc != FO_OK) return rc;
  }
  return FO_OK;
}

extern "C" size_t fo_conv_wpacked_bytes(const fo_conv_t* c) {
  static thread_local ConvPlan plan;
  if (plan_conv(c, &plan, false) != FO_OK) return 0;
  return (size_t)plan.npad * plan.ktot * 2;
}

extern "C" int fo_conv_pack_weights(const fo_conv_t* c, const float* weight, int dimA, int dimB, int n_axis,
                                    const float* n_scale, void* wpacked, fo_stream_t stream) {
  REQUIRE_INIT();
  static thread_local ConvPlan plan;
  int rc = plan_conv(c, &plan, false);
  if (rc != FO_OK) return rc;
  (void)dimA;
  plan.pack.dimB = dimB;
  plan.pack.n_axis = n_axis;
  plan.pack.n_scale = n_scale;
  CUDA_TRY(launch_pack_weights(weight, wpacked, plan.pack, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}

// Plans (tile geometry, K-step table, encoded tensor maps) are cached per thread, keyed on the descriptor bytes: PyTorch's
// caching allocator hands out the same addresses every step, so in steady state a launch costs one hash lookup instead of
// the planner, ~6 getenv calls and up to 15 cuTensorMapEncodeTiled calls.  (The forward runs on the main thread, the
// backward on the autograd thread: two caches, no lock.)  Descriptors come zero-initialised from the bindings, so the
// padding bytes compare equal.
template <typename Plan>
struct PlanCache {
  std::unordered_map<std::string, std::unique_ptr<Plan>> map;
  template <typename Desc, typename Fn>
  int get(const Desc* d, Fn&& build, Plan** out) {
    std::string key(reinterpret_cast<const char*>(d), sizeof(Desc));
    auto it = map.find(key);
    if (it != map.end()) { *out = it->second.get(); return FO_OK; }
    std::unique_ptr<Plan> plan(new Plan);
    const int rc = build(plan.get());
    if (rc != FO_OK) return rc;
    if (map.size() >= 4096) map.clear();   // shapes / addresses keep changing: start over rather than grow without bound
    *out = plan.get();
    map.emplace(std::move(key), std::move(plan));
    return FO_OK;
  }
};

extern "C" int fo_conv_run(const fo_conv_t* c, fo_stream_t stream) {
  REQUIRE_INIT();
  static thread_local PlanCache<ConvPlan> cache;
  ConvPlan* plan = nullptr;
  int rc = cache.get(c, [&](ConvPlan* pl) { return plan_conv(c, pl, true); }, &plan);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_conv_igemm(plan->p, plan->maps, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}

// ------------------------------------------------------------------------------------------ wgrad planner
struct WgradPlan {
  WgradParams p;
  WgradMaps maps;
  FinalizeParams fin;
  int taps;
};

static int side_geometry(int c, int* rowb, int* chunks) {
  const int cpad = (c + 15) / 16 * 16;
  if (cpad == 16) { *rowb = 32; *chunks = 1; }
  else if (cpad == 32) { *rowb = 64; *chunks = 1; }
  else if (cpad == 64) { *rowb = 128; *chunks = 1; }
  else if (cpad == 128) { *rowb = 128; *chunks = 2; }
  else return -1;
  return cpad;
}

static int plan_wgrad(const fo_wgrad_t* g, WgradPlan* out, bool need_maps) {
  WgradParams& p = out->p;
  memset(&p, 0, sizeof(p));
  if (g->form != FO_FORM_S1 && g->form != FO_FORM_DOWN) return fail(FO_ERR_INVALID, "wgrad form must be S1 or DOWN");
  const bool s1 = g->form == FO_FORM_S1;
  if (g->ndim != 2 && !(g->ndim == 3 && s1)) return fail(FO_ERR_INVALID, "wgrad ndim unsupported");
  const int mc = side_geometry(g->p.c, &p.p_rowb, &p.p_chunks);
  const int nc = side_geometry(g->q.c, &p.q_rowb, &p.q_chunks);
  if (mc < 0 || nc < 0) return fail(FO_ERR_INVALID, "wgrad channel counts must pad to 16/32/64/128 (got %d, %d)", g->p.c, g->q.c);
  if (g->p.c_off + mc > g->p.cs || g->q.c_off + nc > g->q.cs) return fail(FO_ERR_INVALID, "wgrad slice exceeds storage");
  p.MC = mc; p.NC = nc; p.p_c0 = g->p.c_off;
  int ext[4];
  if (s1 && g->ndim == 3) { ext[0] = g->w; ext[1] = g->h; ext[2] = g->d; ext[3] = g->n; }
  else if (s1) { ext[0] = g->w; ext[1] = g->h; ext[2] = g->n; ext[3] = 1; }
  else { ext[0] = g->w; ext[1] = 1; ext[2] = g->h; ext[3] = g->n; }  // P low-res (w,h) with a dummy parity dim
  // pixels per stage: 128 (= 8 MMAs per tap) with a halo'd Q box for 3x3(x3) filters when a 128-pixel tile fits in one
  // image, else 64
  int box[4];
  p.kpix = 64;
  p.halo = 0;
  if (s1 && g->ksize == 3) {
    int b2[4];
    // narrow, tall tiles: the halo'd Q box has R + 2 image rows for R rows of pixels, and this kernel's 3x3(x3) forms are
    // bound by L2->SM traffic (Conv3d: P 32 KB + Q box per 1536 MMA cycles) -- 16 x 8 pixels read 1.25x, 64 x 2 read 2x
    int boxw = 16;
    {
      const char* e = getenv("FO_WG_BOXW");   // experiments only
      if (e && atoi(e) >= 8) boxw = atoi(e);
    }
    choose_box(ext, 128, b2, boxw);
    if (b2[2] == 1 && b2[3] == 1 && b2[0] * b2[1] == 128 && b2[0] <= ext[0] && b2[1] <= ext[1] &&
        (b2[0] * p.q_rowb) % 1024 == 0 && (b2[0] * p.p_rowb) % 1024 == 0) {
      p.kpix = 128;
      p.halo = 1;
      for (int d = 0; d < 4; ++d) box[d] = b2[d];
    }
  }
  if (!p.halo) {
    // 128 pixels per stage whenever the tile exists and two stages fit (see WgradParams::kpix); FO_WG_KPIX=64 forces the
    // small stages (experiments)
    const char* kp = getenv("FO_WG_KPIX");
    int b2[4];
    choose_box(ext, 128, b2);
    const int q_loads_max = s1 ? (g->ksize == 3 ? 3 : 1) : 4;
    const int stage128 = (p.p_chunks * 128 * p.p_rowb + q_loads_max * p.q_chunks * 128 * p.q_rowb + 1023) & ~1023;
    if (!(kp && atoi(kp) == 64) && b2[3] == 1 && b2[0] * b2[1] * b2[2] == 128 && 2 * stage128 <= kMaxDynSmem - 2048 - 8192) {
      p.kpix = 128;
      for (int d = 0; d < 4; ++d) box[d] = b2[d];
    } else {
      choose_box(ext, 64, box);
    }
  }
  p.total_ptiles = 1;
  for (int d = 0; d < 4; ++d) {
    p.tile_step[d] = box[d];
    p.tile_cnt[d] = (ext[d] + box[d] - 1) / box[d];
    p.total_ptiles *= p.tile_cnt[d];
    if (d < 3) p.box[d] = box[d];
  }
  if (box[3] != 1) return fail(FO_ERR_INVALID, "wgrad: tensor too small for a 64-pixel tile");
  p.q_tap_off = p.halo ? box[0] * p.q_rowb : 0;
  p.q_box_bytes = (p.halo ? box[0] * (box[1] + 2) : p.kpix) * p.q_rowb;
  // taps; slot order = pass-major.  tap_index maps a slot to the PyTorch filter tap.
  int nt = 0;
  FinalizeParams& f = out->fin;
  const char* wg_env = getenv("FO_WG_GROUP");   // experiments only: 0 = one MMA per tap, no merged passes
  const bool grouping = !(wg_env && atoi(wg_env) == 0);
  const int avail_smem = kMaxDynSmem - 2048 - 8192;   // 8 KB: all-ones tile of the fused bias gradient + alignment
  const int p_bytes = p.p_chunks * p.kpix * p.p_rowb, q_bytes = p.q_chunks * p.q_box_bytes;
  if (s1) {
    const int k = g->ksize, pad = (k - 1) / 2;
    if (k != 1 && k != 3) return fail(FO_ERR_INVALID, "wgrad ksize unsupported");
    const int kd_n = g->ndim == 3 ? k : 1;
    const int sg = g->q_shift_sign < 0 ? -1 : 1;
    if (p.halo) {
      // pass = (kd, kw); the three slots of a filter column kw are the vertical offsets -1, 0, +1 read from one box that
      // starts one row above the tile; slot j serves filter row kh with sg * (kh - pad) == j - 1.
      // Narrow Q sides (9 accumulators of NC columns fit in TMEM next to the bias column): the three filter columns
      // share a pass too -- three Q boxes per stage, P is read once instead of three times.
      const bool merge_kw = grouping && g->ndim == 2 && 9 * nc + 16 <= 512 && 2 * ((p_bytes + 3 * q_bytes + 1023) & ~1023) <= avail_smem;
      for (int kd = 0; kd < kd_n; ++kd)
        for (int kw = 0; kw < k; ++kw)
          for (int j = 0; j < k; ++j) {
            const int kh = pad + sg * (j - 1);
            WgTap& t = p.taps[nt];
            t.c0 = (int16_t)g->q.c_off; t.d1 = (int8_t)(sg * (kw - pad)); t.d2 = (int8_t)(-1);
            t.d3 = (int8_t)(g->ndim == 3 ? sg * (kd - pad) : 0); t.map = 0; t.pad = 0;
            f.tap_index[nt] = (int8_t)((kd * k + kh) * k + kw);
            ++nt;
          }
      p.taps_per_load = 3;
      p.q_loads = merge_kw ? 3 : 1;
    } else {
      for (int kd = 0; kd < kd_n; ++kd)
        for (int kh = 0; kh < k; ++kh)
          for (int kw = 0; kw < k; ++kw) {
            WgTap& t = p.taps[nt];
            t.c0 = (int16_t)g->q.c_off; t.d1 = (int8_t)(sg * (kw - pad)); t.d2 = (int8_t)(sg * (kh - pad));
            t.d3 = (int8_t)(g->ndim == 3 ? sg * (kd - pad) : 0); t.map = 0; t.pad = 0;
            f.tap_index[nt] = (int8_t)nt;
            ++nt;
          }
      p.taps_per_load = 1;
      p.q_loads = k == 1 ? 1 : 3;
    }
  } else {
    for (int ky = 0; ky < 4; ++ky)
      for (int kx = 0; kx < 4; ++kx) {
        WgTap& t = p.taps[nt];
        t.c0 = (int16_t)(kDownPar[kx] * g->q.cs + g->q.c_off); t.d1 = (int8_t)kDownD[kx]; t.d2 = (int8_t)kDownPar[ky];
        t.d3 = (int8_t)kDownD[ky]; t.map = 0; t.pad = 0;
        f.tap_index[nt] = (int8_t)nt;
        ++nt;
      }
    p.taps_per_load = 1;
    p.q_loads = 4;
  }
  p.taps_per_pass = p.q_loads * p.taps_per_load;
  // taps that share one MMA: a whole halo box (3 taps, one image row apart) or all the boxes of a stage (q_bytes apart);
  // needs one chunk per tap (NC <= 64) and N = G * NC <= 256
  p.mma_group = 1;
  if (grouping && p.q_chunks == 1) {
    int gsz = p.taps_per_load > 1 ? p.taps_per_load : p.q_loads;
    while (gsz > 1 && (gsz * nc > 256 || (p.taps_per_load > 1 ? p.taps_per_load : p.q_loads) % gsz != 0)) --gsz;
    p.mma_group = gsz;
  }
  out->taps = nt;
  p.passes = nt / p.taps_per_pass;
  const int stage_bytes = (p_bytes + p.q_loads * q_bytes + 1023) & ~1023;
  int stages = avail_smem / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return fail(FO_ERR_INVALID, "wgrad stage does not fit");
  p.stages = stages;
  int splits = (g_num_sms > 0 ? g_num_sms : 148) / p.passes;
  if (splits < 1) splits = 1;
  if (splits > p.total_ptiles) splits = p.total_ptiles;
  p.splits = splits;
  {
    // accumulation chains (wgrad_igemm.cu): at most ~4096 accumulating MMAs per TMEM accumulator, then a new partial slot
    int chain_mmas = 4096;
    const char* e = getenv("FO_WG_CHAIN");   // experiments: 0 = one chain per CTA
    if (e != nullptr) chain_mmas = atoi(e);
    const int per = (p.total_ptiles + splits - 1) / splits;
    wgrad_plan_chains(per, p.kpix, chain_mmas, &p.n_flush, &p.chain_tiles);
  }
  p.partial = (float*)g->workspace;
  {
    const char* sk = getenv("FO_SKIP_MMA");
    p.dbg_skip_mma = sk ? atoi(sk) : 0;
    const char* bo = getenv("FO_BACKOFF");
    p.backoff_ns = bo ? atoi(bo) : 0;
    const char* st = getenv("FO_WG_STAGES");
    if (st && atoi(st) >= 2 && atoi(st) <= p.stages) p.stages = atoi(st);
  }

  f.partial = p.partial; f.dweight = g->dweight; f.splits = splits * p.n_flush; f.taps = nt; f.MC = mc; f.NC = nc;
  f.m_real = g->p.c; f.n_real = g->q.c; f.dimB = g->dimB; f.m_axis = g->m_axis; f.q_w_off = g->q_w_off;
  f.accumulate = g->accumulate;
  if (!need_maps) return FO_OK;

  uint64_t dims[5], str[5];
  uint32_t bx[5];
  // P map
  {
    const uint64_t cs = g->p.cs;
    if (s1 && g->ndim == 3) {
      dims[0] = cs; dims[1] = g->w; dims[2] = g->h; dims[3] = g->d; dims[4] = g->n;
      str[0] = 1; str[1] = cs; str[2] = (uint64_t)g->w * cs; str[3] = (uint64_t)g->h * g->w * cs; str[4] = (uint64_t)g->d * g->h * g->w * cs;
    } else if (s1) {
      dims[0] = cs; dims[1] = g->w; dims[2] = g->h; dims[3] = g->n; dims[4] = 1;
      str[0] = 1; str[1] = cs; str[2] = (uint64_t)g->w * cs; str[3] = (uint64_t)g->h * g->w * cs; str[4] = (uint64_t)g->n * g->h * g->w * cs;
    } else {
      dims[0] = cs; dims[1] = g->w; dims[2] = 1; dims[3] = g->h; dims[4] = g->n;
      str[0] = 1; str[1] = cs; str[2] = (uint64_t)g->w * cs; str[3] = (uint64_t)g->w * cs; str[4] = (uint64_t)g->h * g->w * cs;
    }
    bx[0] = p.p_rowb / 2; bx[1] = box[0]; bx[2] = box[1]; bx[3] = box[2]; bx[4] = 1;
    int rc = encode_map(&out->maps.p, g->p.ptr, 5, dims, str, bx, p.p_rowb);
    if (rc != FO_OK) return rc;
  }
  {
    const uint64_t cs = g->q.cs;
    if (s1 && g->ndim == 3) {
      dims[0] = cs; dims[1] = g->w; dims[2] = g->h; dims[3] = g->d; dims[4] = g->n;
      str[0] = 1; str[1] = cs; str[2] = (uint64_t)g->w * cs; str[3] = (uint64_t)g->h * g->w * cs; str[4] = (uint64_t)g->d * g->h * g->w * cs;
    } else if (s1) {
      dims[0] = cs; dims[1] = g->w; dims[2] = g->h; dims[3] = g->n; dims[4] = 1;
      str[0] = 1; str[1] = cs; str[2] = (uint64_t)g->w * cs; str[3] = (uint64_t)g->h * g->w * cs; str[4] = (uint64_t)g->n * g->h * g->w * cs;
    } else {  // Q is the hi-res tensor [n, 2h, 2w, cs] in parity view
      const uint64_t W2 = 2ULL * g->w, H2 = 2ULL * g->h;
      dims[0] = 2 * cs; dims[1] = g->w; dims[2] = 2; dims[3] = g->h; dims[4] = g->n;
      str[0] = 1; str[1] = 2 * cs; str[2] = W2 * cs; str[3] = 2 * W2 * cs; str[4] = H2 * W2 * cs;
    }
    bx[0] = p.q_rowb / 2; bx[1] = box[0]; bx[2] = box[1] + (p.halo ? 2 : 0); bx[3] = box[2]; bx[4] = 1;
    int rc = encode_map(&out->maps.q[0], g->q.ptr, 5, dims, str, bx, p.q_rowb);
    if (rc != FO_OK) return rc;
    for (int i = 1; i < kMaxAMaps; ++i) out->maps.q[i] = out->maps.q[0];
  }
  return FO_OK;
}

extern "C" size_t fo_wgrad_workspace_bytes(const fo_wgrad_t* g) {
  if (fo_init() != FO_OK) return 0;
  static thread_local PlanCache<WgradPlan> cache;   // (the query precedes the run and carries no workspace yet: its own cache)
  WgradPlan* planp = nullptr;
  if (cache.get(g, [&](WgradPlan* pl) { return plan_wgrad(g, pl, false); }, &planp) != FO_OK) return 0;
  const WgradPlan& plan = *planp;
  const size_t slots = (size_t)plan.p.splits * plan.p.n_flush;
  return slots * plan.taps * plan.p.MC * plan.p.NC * sizeof(float) + (size_t)plan.p.passes * slots * plan.p.MC * sizeof(float);
}

extern "C" int fo_wgrad_run(const fo_wgrad_t* g, fo_stream_t stream) {
  REQUIRE_INIT();
  static thread_local PlanCache<WgradPlan> cache;
  WgradPlan* planp = nullptr;
  int rc = cache.get(g, [&](WgradPlan* pl) { return plan_wgrad(g, pl, true); }, &planp);
  if (rc != FO_OK) return rc;
  WgradPlan& plan = *planp;
  const size_t slots = (size_t)plan.p.splits * plan.p.n_flush;
  const size_t main_bytes = slots * plan.taps * plan.p.MC * plan.p.NC * sizeof(float);
  const size_t need = main_bytes + (size_t)plan.p.passes * slots * plan.p.MC * sizeof(float);
  if (g->workspace == nullptr || g->workspace_bytes < need)
    return fail(FO_ERR_INVALID, "wgrad workspace too small: %zu < %zu", g->workspace_bytes, need);
  plan.p.bias_partial = nullptr;
  if (g->dbias != nullptr) {
    if (plan.p.taps_per_pass * plan.p.NC + 16 > 512) return fail(FO_ERR_INVALID, "wgrad: no TMEM room for the fused bias gradient");
    plan.p.bias_partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g->workspace) + main_bytes);
  }
  CUDA_TRY(launch_wgrad_igemm(plan.p, plan.maps, (cudaStream_t)stream));
  // one launch: split reduction + scatter of the weight gradient, and the bias gradient's column-sum partials
  CUDA_TRY(launch_wgrad_finalize(plan.fin, plan.p.bias_partial, plan.p.passes * (int)slots, g->p.c, g->dbias,
                                 g->dbias_accumulate, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}

// ------------------------------------------------------------------------------------------ elementwise
extern "C" int fo_pack_nchw(const float* x, void* out, int n, int c, int hw, int cs, const float* shift,
                            const float* scale, fo_stream_t stream) {
  REQUIRE_INIT();
  if (cs % 8 != 0 || c > cs) return fail(FO_ERR_INVALID, "pack: bad channel counts %d/%d", c, cs);
  if (cs != 16 && cs != 32 && shift != nullptr) return fail(FO_ERR_INVALID, "pack: shift/scale only for cs 16 / 32");
  CUDA_TRY(launch_pack_nchw(x, out, n, c, hw, cs, shift, scale, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_unpack_nchw(const void* x, float* out, int n, int c, int hw, int cs, fo_stream_t stream) {
  REQUIRE_INIT();
  CUDA_TRY(launch_unpack_nchw(x, out, n, c, hw, cs, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_u8hwc_to_nchw(const void* x, float* out, int n, int hw, int c_total, int c_off, float mean, float stdv,
                                fo_stream_t stream) {
  REQUIRE_INIT();
  if (hw % 4 != 0 || c_off < 0 || c_off + 3 > c_total) return fail(FO_ERR_INVALID, "u8hwc_to_nchw: hw %% 4 == 0 and c_off + 3 <= c_total required");
  if ((reinterpret_cast<uintptr_t>(x) & 3) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0)
    return fail(FO_ERR_INVALID, "u8hwc_to_nchw: input must be 4-byte and output 16-byte aligned");
  if ((size_t)n * hw == 0) return FO_OK;
  CUDA_TRY(launch_u8hwc_to_nchw(x, out, n, hw, c_total, c_off, mean, stdv, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_relu(const void* x, void* y, size_t numel, fo_stream_t stream) {
  REQUIRE_INIT();
  if (numel % 8 != 0) return fail(FO_ERR_INVALID, "relu: numel must be a multiple of 8");
  CUDA_TRY(launch_relu(x, y, numel, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_adam_chunk_elems(void) { return 16384; }
extern "C" int fo_adam_step(const fo_adam_tensor_t* table_dev, const int* chunks_dev, int n_chunks, float lr, float beta1,
                            float beta2, float eps, float weight_decay, int step, float grad_scale, fo_stream_t stream) {
  REQUIRE_INIT();
  if (step < 1) return fail(FO_ERR_INVALID, "adam: step counts from 1");
  if (n_chunks <= 0) return FO_OK;
  // bias corrections in double like torch (1 - beta ** step), then rounded once
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float sbc2 = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  CUDA_TRY(launch_adam(table_dev, chunks_dev, n_chunks, lr, beta1, beta2, eps, weight_decay, bc1, sbc2, grad_scale,
                       g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" size_t fo_colsum_workspace_bytes(int cs) {
  if (fo_init() != FO_OK) return 0;
  return (size_t)colsum_blocks(g_num_sms) * cs * sizeof(float);
}
extern "C" int fo_colsum(const void* x, size_t rows, int cs, int c_off, int c, float* out, int accumulate,
                         void* workspace, size_t workspace_bytes, fo_stream_t stream) {
  REQUIRE_INIT();
  if (cs % 8 != 0 || cs > 2048) return fail(FO_ERR_INVALID, "colsum: cs must be a multiple of 8");
  if (workspace_bytes < fo_colsum_workspace_bytes(cs)) return fail(FO_ERR_INVALID, "colsum workspace too small");
  CUDA_TRY(launch_colsum(x, rows, cs, c_off, c, out, accumulate, (float*)workspace, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_maxpool2(const void* x, void* y, int n, int h, int w, int cs, fo_stream_t stream) {
  REQUIRE_INIT();
  if ((h | w) & 1 || cs % 8) return fail(FO_ERR_INVALID, "maxpool2: even h, w and cs %% 8 == 0 required");
  CUDA_TRY(launch_maxpool2(x, y, n, h, w, cs, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_maxpool2_bwd(const void* x, const void* y, const void* dy, void* dx, int n, int h, int w, int cs,
                               fo_stream_t stream) {
  REQUIRE_INIT();
  CUDA_TRY(launch_maxpool2_bwd(x, y, dy, dx, n, h, w, cs, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}

extern "C" int fo_im2col4x4s2(const float* x, void* out, int n, int ca, int c, int h, int w, fo_stream_t stream) {
  REQUIRE_INIT();
  if (c > 8 || c > ca || ((h | w) & 1)) return fail(FO_ERR_INVALID, "im2col4x4s2: need c <= 8, even h and w");
  CUDA_TRY(launch_im2col4x4s2(x, out, n, ca, c, h, w, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_im2col3x3(const float* x, void* out, int n, int c, int h, int w, const float* shift, const float* scale,
                            fo_stream_t stream) {
  REQUIRE_INIT();
  if (c > 3) return fail(FO_ERR_INVALID, "im2col3x3: need c <= 3");
  CUDA_TRY(launch_im2col3x3(x, out, n, c, h, w, shift, scale, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_vgg_first_conv(const float* x, int n, int h, int w, const float* weight, const float* bias,
                                 const float* shift, const float* scale, void* out_relu, fo_stream_t stream) {
  REQUIRE_INIT();
  if (n < 1 || h < 1 || w < 1) return fail(FO_ERR_INVALID, "vgg_first_conv: bad extents");
  if ((shift == nullptr) != (scale == nullptr)) return fail(FO_ERR_INVALID, "vgg_first_conv: shift and scale go together");
  if ((reinterpret_cast<uintptr_t>(out_relu) & 15) != 0) return fail(FO_ERR_INVALID, "vgg_first_conv: output must be 16-byte aligned");
  CUDA_TRY(launch_vgg_first_conv(x, n, h, w, weight, bias, shift, scale, out_relu, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_vgg_first_dgrad(const void* dy, int n, int h, int w, const float* weight, const float* scale, float* dx,
                                  fo_stream_t stream) {
  REQUIRE_INIT();
  if (n < 1 || h < 1 || w < 1) return fail(FO_ERR_INVALID, "vgg_first_dgrad: bad extents");
  // dy: bf16 channels-last [n, h, w, 64]; one TMA box = 8 rows x 32 pixels x 64 channels (out-of-range pixels read as zero)
  CUtensorMap map;
  const uint64_t dims[5] = {64, (uint64_t)w, (uint64_t)h, (uint64_t)n, 1};
  const uint64_t str[5] = {1, 64, (uint64_t)w * 64, (uint64_t)h * w * 64, (uint64_t)n * h * w * 64};
  const uint32_t bx[5] = {64, 32, 8, 1, 1};
  int rc = encode_map(&map, dy, 5, dims, str, bx, 128);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_vgg_first_dgrad(&map, n, h, w, weight, scale, dx, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
static int s2_check(int n, int ca, int c, int H, int W) {
  if (n < 1 || (c != 3 && c != 6) || ca < c || H < 2 || W < 2 || ((H | W) & 1))
    return fail(FO_ERR_INVALID, "s2conv: needs c in {3, 6}, ca >= c and even H, W (got c=%d ca=%d H=%d W=%d)", c, ca, H, W);
  return FO_OK;
}
extern "C" int fo_s2conv(const float* x, int n, int ca, int c, int H, int W, const float* weight, const float* bias,
                         const void* mask, const void* addend, void* out, int relu, fo_stream_t stream) {
  REQUIRE_INIT();
  int rc = s2_check(n, ca, c, H, W);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_s2conv(x, n, ca, c, H, W, weight, bias, mask, addend, out, relu, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" size_t fo_s2wgrad_workspace_bytes(void) {
  if (fo_init() != FO_OK) return 0;
  return (size_t)s2_grid(g_num_sms) * 128 * 64 * sizeof(float);
}
extern "C" int fo_s2wgrad(const float* x, int n, int ca, int c, int H, int W, const void* y, float* dweight, int accumulate,
                          float* dbias, int dbias_accumulate, void* workspace, size_t workspace_bytes, fo_stream_t stream) {
  REQUIRE_INIT();
  int rc = s2_check(n, ca, c, H, W);
  if (rc != FO_OK) return rc;
  if (workspace == nullptr || workspace_bytes < fo_s2wgrad_workspace_bytes()) return fail(FO_ERR_INVALID, "s2wgrad: workspace too small");
  CUDA_TRY(launch_s2wgrad(x, n, ca, c, H, W, y, dweight, accumulate, dbias, dbias_accumulate, (float*)workspace, g_num_sms,
                          (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_col2im4x4s2(const void* col, const float* bias, float* out, int n, int c, int hi, int wi,
                              fo_stream_t stream) {
  REQUIRE_INIT();
  if (c > 8) return fail(FO_ERR_INVALID, "col2im4x4s2: need c <= 8");
  CUDA_TRY(launch_col2im4x4s2(col, bias, out, n, c, hi, wi, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_chansum_nchw(const float* x, int n, int ca, int c, int hw, float* out, int accumulate,
                               fo_stream_t stream) {
  REQUIRE_INIT();
  CUDA_TRY(launch_chansum_nchw(x, n, ca, c, hw, out, accumulate, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}

// ------------------------------------------------------------------------------------------ VQ
extern "C" int fo_vq_prep(const float* embed, int dim, int n_embed, void* e_split, float* e_t, float* e_norm2,
                          fo_stream_t stream) {
  REQUIRE_INIT();
  CUDA_TRY(launch_vq_prep(embed, dim, n_embed, e_split, e_t, e_norm2, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" size_t fo_vq_split_elems(int dim, int n_embed) { return vq_split_elems(dim, n_embed); }
extern "C" size_t fo_vq_assign_workspace_bytes(size_t rows, int dim) { return vq_assign_workspace_bytes(rows, dim); }
extern "C" int fo_vq_assign(const float* x, size_t rows, int dim, int n_embed, const float* e_t, const void* e_split,
                            const float* e_norm2, int64_t* embed_ind, int* n_flagged, void* workspace,
                            size_t workspace_bytes, fo_stream_t stream) {
  REQUIRE_INIT();
  if (dim <= 0 || n_embed <= 0) return fail(FO_ERR_INVALID, "vq_assign: dim and n_embed must be positive");
  if (vq_assign_is_generic(dim, n_embed)) {
    // any other codebook shape the reference's Quantize accepts: exact fp64 scan on CUDA cores (vq.cu)
    if (dim > 1536) return fail(FO_ERR_INVALID, "vq_assign: dim %d above the generic kernel's 1536", dim);
    if (rows == 0) return FO_OK;
    CUDA_TRY(launch_vq_assign_generic(x, rows, dim, n_embed, e_t, embed_ind, n_flagged, g_num_sms, (cudaStream_t)stream));
    return FO_OK;
  }
  if (workspace_bytes < vq_assign_workspace_bytes(rows, dim)) return fail(FO_ERR_INVALID, "vq_assign workspace too small");
  if (rows == 0) return FO_OK;
  CUtensorMap map_e;
  uint64_t dims[2] = {(uint64_t)2 * dim, (uint64_t)n_embed};
  uint64_t str[2] = {1, (uint64_t)2 * dim};
  uint32_t bx[2] = {64, 256};
  int rc = encode_map(&map_e, e_split, 2, dims, str, bx, 128);
  if (rc != FO_OK) return rc;
  // the augmented K slice [n_pad][16] bf16 behind the split codebook (see fo_vq_prep): 32-byte rows, 32B swizzle
  CUtensorMap map_x;
  const uint64_t n_pad = ((uint64_t)n_embed + 255) / 256 * 256;
  uint64_t xdims[2] = {16, n_pad};
  uint64_t xstr[2] = {1, 16};
  uint32_t xbx[2] = {16, 256};
  rc = encode_map(&map_x, (const uint8_t*)e_split + (size_t)n_embed * 2 * dim * 2, 2, xdims, xstr, xbx, 32);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_vq_assign(x, rows, dim, n_embed, e_t, e_split, e_norm2, embed_ind, n_flagged, workspace, &map_e,
                            &map_x, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" size_t fo_vq_gather_scratch_bytes(int dim, int n_embed) { return vq_gather_scratch_bytes(dim, n_embed); }
extern "C" int fo_vq_gather_stats(const float* x, const int64_t* embed_ind, size_t rows, int dim, int n_embed,
                                  const float* e_t, float* q_f32, void* q_bf16, float* diff_sum, float* counts,
                                  float* embed_sum, float* scratch, fo_stream_t stream) {
  REQUIRE_INIT();
  if (dim % 4 != 0) return fail(FO_ERR_INVALID, "vq: dim must be a multiple of 4");
  if (rows == 0) return FO_OK;
  if (scratch != nullptr && (reinterpret_cast<uintptr_t>(scratch) & 15) != 0) return fail(FO_ERR_INVALID, "vq: scratch must be 16-byte aligned");
  CUDA_TRY(launch_vq_gather_stats(x, embed_ind, rows, dim, n_embed, e_t, q_f32, q_bf16, diff_sum, counts, embed_sum,
                                  scratch, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_vq_ema(float* embed, float* cluster_size, float* embed_avg, const float* counts,
                         const float* embed_sum, int dim, int n_embed, float decay, float one_minus_decay, float eps,
                         fo_stream_t stream) {
  REQUIRE_INIT();
  CUDA_TRY(launch_vq_ema(embed, cluster_size, embed_avg, counts, embed_sum, dim, n_embed, decay, one_minus_decay, eps,
                         (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_vq_backward(const void* g_q, int g_q_is_bf16, int g_cs, int g_c_off, const float* g_diff,
                              const float* x, const int64_t* embed_ind, const float* e_t, size_t rows, int dim,
                              int n_embed, float* gx_f32, void* gx_bf16, fo_stream_t stream) {
  REQUIRE_INIT();
  if (rows == 0) return FO_OK;
  CUDA_TRY(launch_vq_backward(g_q, g_q_is_bf16, g_cs, g_c_off, g_diff, x, embed_ind, e_t, rows, dim, n_embed, gx_f32,
                              gx_bf16, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}

// ------------------------------------------------------------------------------------------ LPIPS / MSE
extern "C" int fo_lpips_tap(const void* f0, const void* f1, const float* w, int n, int hw, int c, float* out,
                            fo_stream_t stream) {
  REQUIRE_INIT();
  if (c % 64 != 0 || c > 512) return fail(FO_ERR_INVALID, "lpips_tap: c must be 64..512, multiple of 64");
  CUDA_TRY(launch_lpips_tap(f0, f1, w, n, hw, c, out, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_lpips_tap_bwd(const void* f0, const void* f1, const float* w, const float* g, int n, int hw, int c,
                                void* d_f0, const void* addend, fo_stream_t stream) {
  REQUIRE_INIT();
  if (c % 64 != 0 || c > 512) return fail(FO_ERR_INVALID, "lpips_tap_bwd: c must be 64..512, multiple of 64");
  CUDA_TRY(launch_lpips_tap_bwd(f0, f1, w, g, n, hw, c, d_f0, addend, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_lpips_tap_pool(const void* f0, const void* f1, const float* w, int n, int h, int wd, int c, float* out,
                                 void* pooled, void* pooled1, fo_stream_t stream) {
  REQUIRE_INIT();
  if (c % 64 != 0 || c > 512) return fail(FO_ERR_INVALID, "lpips_tap_pool: c must be 64..512, multiple of 64");
  if (h < 2 || wd < 2 || ((h | wd) & 1)) return fail(FO_ERR_INVALID, "lpips_tap_pool: even h, w required");
  CUDA_TRY(launch_lpips_tap_pool(f0, f1, w, n, h, wd, c, out, pooled, pooled1, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_lpips_tap_bwd_pool(const void* f0, const void* f1, const float* w, const float* g, int n, int h, int wd,
                                     int c, void* d_f0, const void* pool_dy, fo_stream_t stream) {
  REQUIRE_INIT();
  if (c % 64 != 0 || c > 512) return fail(FO_ERR_INVALID, "lpips_tap_bwd_pool: c must be 64..512, multiple of 64");
  if (h < 2 || wd < 2 || ((h | wd) & 1)) return fail(FO_ERR_INVALID, "lpips_tap_bwd_pool: even h, w required");
  CUDA_TRY(launch_lpips_tap_bwd_pool(f0, f1, w, g, n, h, wd, c, d_f0, pool_dy, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_lpips_tap_split(const void* f0, const void* f1, const float* w, int n, int hw, int c, float* out,
                                  fo_stream_t stream) {
  REQUIRE_INIT();
  if (c % 64 != 0 || c > 512) return fail(FO_ERR_INVALID, "lpips_tap: c must be 64..512, multiple of 64");
  CUDA_TRY(launch_lpips_tap(f0, f1, w, n, hw, c, out, g_num_sms, (cudaStream_t)stream, 1));
  return FO_OK;
}
extern "C" int fo_lpips_tap_bwd_split(const void* f0, const void* f1, const float* w, const float* g, int n, int hw, int c,
                                      void* d_f0, const void* addend, fo_stream_t stream) {
  REQUIRE_INIT();
  if (c % 64 != 0 || c > 512) return fail(FO_ERR_INVALID, "lpips_tap_bwd: c must be 64..512, multiple of 64");
  CUDA_TRY(launch_lpips_tap_bwd(f0, f1, w, g, n, hw, c, d_f0, addend, g_num_sms, (cudaStream_t)stream, 1));
  return FO_OK;
}
// ------------------------------------------------------------------------------------------ discriminators (8(f1))
static int to_dconv(const fo_dconv_t* d, DConvParams* p) {
  if (d->n < 1 || d->cin < 1 || d->cout < 1 || d->kd < 1 || d->kh < 1 || d->kw < 1 || d->sd < 1 || d->sh < 1 || d->sw < 1)
    return fail(FO_ERR_INVALID, "dconv: bad geometry");
  if (d->od != (d->id + 2 * d->pd - d->kd) / d->sd + 1 || d->oh != (d->ih + 2 * d->ph - d->kh) / d->sh + 1 ||
      d->ow != (d->iw + 2 * d->pw - d->kw) / d->sw + 1)
    return fail(FO_ERR_INVALID, "dconv: output extents do not match (i + 2p - k) / s + 1");
  if ((long long)d->n * d->od * d->oh * d->ow > 0x7fffffffLL || (long long)d->n * d->id * d->ih * d->iw > 0x7fffffffLL)
    return fail(FO_ERR_INVALID, "dconv: more than 2^31 positions");
  p->n = d->n; p->cin = d->cin; p->id = d->id; p->ih = d->ih; p->iw = d->iw;
  p->cout = d->cout; p->od = d->od; p->oh = d->oh; p->ow = d->ow;
  p->kd = d->kd; p->kh = d->kh; p->kw = d->kw; p->sd = d->sd; p->sh = d->sh; p->sw = d->sw;
  p->pd = d->pd; p->ph = d->ph; p->pw = d->pw;
  return FO_OK;
}
extern "C" int fo_dconv_fwd(const fo_dconv_t* d, const float* x, const float* w, const float* bias, float* y,
                            fo_stream_t stream) {
  REQUIRE_INIT();
  DConvParams p;
  int rc = to_dconv(d, &p);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_dconv_fwd(p, x, w, bias, y, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_dconv_dgrad(const fo_dconv_t* d, const float* dy, const float* w, float* dx, fo_stream_t stream) {
  REQUIRE_INIT();
  DConvParams p;
  int rc = to_dconv(d, &p);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_dconv_dgrad(p, dy, w, dx, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_dconv_wgrad(const fo_dconv_t* d, const float* x, const float* dy, float* dw, float* dbias,
                              fo_stream_t stream) {
  REQUIRE_INIT();
  DConvParams p;
  int rc = to_dconv(d, &p);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_dconv_wgrad(p, x, dy, dw, dbias, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_dconv_im2col_pairs(const fo_dconv_t* d, const float* x, void* col, int kp, int parts, fo_stream_t stream) {
  REQUIRE_INIT();
  DConvParams p = {};
  int rc = to_dconv(d, &p);
  if (rc != FO_OK) return rc;
  if (kp % 8 != 0 || kp < d->cin * d->kd * d->kh * d->kw) return fail(FO_ERR_INVALID, "im2col_pairs: kp must be a multiple of 8 and >= K");
  if (parts != 2 && parts != 3) return fail(FO_ERR_INVALID, "im2col_pairs: parts must be 2 or 3");
  CUDA_TRY(launch_dim2col_pairs(p, x, col, kp, parts, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_dconv_im2col_t(const fo_dconv_t* d, const float* x, void* out, int kp, int pc, int chunks,
                                 fo_stream_t stream) {
  REQUIRE_INIT();
  DConvParams p = {};
  int rc = to_dconv(d, &p);
  if (rc != FO_OK) return rc;
  if (pc < 8 || pc % 8 != 0 || (long long)pc * chunks < (long long)d->n * d->od * d->oh * d->ow)
    return fail(FO_ERR_INVALID, "im2col_t: pc (multiple of 8) * chunks must cover all positions");
  if (kp < d->cin * d->kd * d->kh * d->kw) return fail(FO_ERR_INVALID, "im2col_t: kp < K");
  CUDA_TRY(launch_dim2col_t(p, x, out, kp, pc, chunks, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_dconv_col2im(const fo_dconv_t* d, const float* dcol, long long ld, float* dx, fo_stream_t stream) {
  REQUIRE_INIT();
  DConvParams p;
  int rc = to_dconv(d, &p);
  if (rc != FO_OK) return rc;
  if (ld < (long long)d->cin * d->kd * d->kh * d->kw) return fail(FO_ERR_INVALID, "col2im: ld < K");
  CUDA_TRY(launch_dcol2im(p, dcol, ld, dx, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_dconv_dbias(const fo_dconv_t* d, const float* dy, float* dbias, fo_stream_t stream) {
  REQUIRE_INIT();
  DConvParams p = {};
  int rc = to_dconv(d, &p);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_dconv_dbias(dy, p.n, p.cout, p.od * p.oh * p.ow, dbias, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_instnorm_fwd(const float* x, float* y, int n, int c, long long plane, float eps, float slope, int training,
                               float momentum, float* running_mean, float* running_var, float* save, fo_stream_t stream) {
  REQUIRE_INIT();
  if (n < 1 || c < 1 || plane < 1 || slope == 0.f) return fail(FO_ERR_INVALID, "instnorm: bad arguments");
  if (!training && (running_mean == nullptr || running_var == nullptr))
    return fail(FO_ERR_INVALID, "instnorm: eval mode needs the running statistics");
  CUDA_TRY(launch_instnorm_fwd(x, y, n, c, plane, eps, slope, training, momentum, running_mean, running_var, save,
                               (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_instnorm_bwd(const float* y, const float* dy, float* dx, int n, int c, long long plane, float slope,
                               int training, const float* save, fo_stream_t stream) {
  REQUIRE_INIT();
  if (n < 1 || c < 1 || plane < 1 || slope == 0.f || save == nullptr) return fail(FO_ERR_INVALID, "instnorm_bwd: bad arguments");
  CUDA_TRY(launch_instnorm_bwd(y, dy, dx, n, c, plane, slope, training, save, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_lrelu(const float* x, float* y, size_t numel, float slope, fo_stream_t stream) {
  REQUIRE_INIT();
  if (numel == 0) return FO_OK;
  CUDA_TRY(launch_lrelu(x, y, numel, slope, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_lrelu_bwd(const float* y, const float* dy, float* dx, size_t numel, float slope, fo_stream_t stream) {
  REQUIRE_INIT();
  if (numel == 0) return FO_OK;
  CUDA_TRY(launch_lrelu_bwd(y, dy, dx, numel, slope, (cudaStream_t)stream));
  return FO_OK;
}
static int check_pool(int id, int ih, int iw, int od, int oh, int ow, int kd, int sd, int sh, int sw) {
  if ((kd != 1 && kd != 3) || sd < 1 || sh < 1 || sw < 1) return fail(FO_ERR_INVALID, "avgpool3: kd must be 1 or 3");
  if (od != (id + 2 * (kd / 2) - kd) / sd + 1 || oh != (ih + 2 - 3) / sh + 1 || ow != (iw + 2 - 3) / sw + 1)
    return fail(FO_ERR_INVALID, "avgpool3: output extents do not match");
  return FO_OK;
}
extern "C" int fo_avgpool3(const float* x, float* y, long long planes, int id, int ih, int iw, int od, int oh, int ow, int kd,
                           int sd, int sh, int sw, fo_stream_t stream) {
  REQUIRE_INIT();
  int rc = check_pool(id, ih, iw, od, oh, ow, kd, sd, sh, sw);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_avgpool3(x, y, planes, id, ih, iw, od, oh, ow, kd, sd, sh, sw, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_avgpool3_bwd(const float* dy, float* dx, long long planes, int id, int ih, int iw, int od, int oh, int ow,
                               int kd, int sd, int sh, int sw, fo_stream_t stream) {
  REQUIRE_INIT();
  int rc = check_pool(id, ih, iw, od, oh, ow, kd, sd, sh, sw);
  if (rc != FO_OK) return rc;
  CUDA_TRY(launch_avgpool3_bwd(dy, dx, planes, id, ih, iw, od, oh, ow, kd, sd, sh, sw, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_ralsgan(const float* a, int n, const float* b, int m, float target, float* out, fo_stream_t stream) {
  REQUIRE_INIT();
  if (n < 1 || m < 1) return fail(FO_ERR_INVALID, "ralsgan: empty prediction");
  CUDA_TRY(launch_ralsgan(a, n, b, m, target, out, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_ralsgan_bwd(const float* a, int n, int m, float target, const float* fwd, const float* g, float* da,
                              float* db, fo_stream_t stream) {
  REQUIRE_INIT();
  if (n < 1 || m < 1) return fail(FO_ERR_INVALID, "ralsgan: empty prediction");
  CUDA_TRY(launch_ralsgan_bwd(a, n, m, target, fwd, g, da, db, (cudaStream_t)stream));
  return FO_OK;
}

// ------------------------------------------------------------------------------------------ verification mode helpers
extern "C" int fo_split_f32(const float* x, int n, int c, int hw, long long sn, long long sc, long long sp, void* out,
                            int cp, fo_stream_t stream) {
  REQUIRE_INIT();
  if (cp % 16 != 0 || c > cp) return fail(FO_ERR_INVALID, "split_f32: cp must be a multiple of 16 and >= c");
  if ((size_t)n * hw == 0) return FO_OK;
  CUDA_TRY(launch_split_f32(x, n, c, hw, sn, sc, sp, out, cp, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_merge_f32(const void* in, int n, int c, int hw, int cp, float* out, long long sn, long long sc,
                            long long sp, fo_stream_t stream) {
  REQUIRE_INIT();
  if (cp % 16 != 0 || c > cp) return fail(FO_ERR_INVALID, "merge_f32: cp must be a multiple of 16 and >= c");
  if ((size_t)n * hw == 0) return FO_OK;
  CUDA_TRY(launch_merge_f32(in, n, c, hw, cp, out, sn, sc, sp, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_maxpool2_f32(const float* x, float* y, int n, int h, int w, int c, fo_stream_t stream) {
  REQUIRE_INIT();
  if ((h | w) & 1) return fail(FO_ERR_INVALID, "maxpool2_f32: even h, w required");
  CUDA_TRY(launch_maxpool2_f32(x, y, n, h, w, c, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_maxpool2_bwd_f32(const float* x, const float* y, const float* dy, float* dx, int n, int h, int w, int c,
                                   fo_stream_t stream) {
  REQUIRE_INIT();
  if ((h | w) & 1) return fail(FO_ERR_INVALID, "maxpool2_bwd_f32: even h, w required");
  CUDA_TRY(launch_maxpool2_bwd_f32(x, y, dy, dx, n, h, w, c, (cudaStream_t)stream));
  return FO_OK;
}

extern "C" int fo_mse(const float* a, const float* b, int n, int ca, int c, int hw, float* sum_out, fo_stream_t stream) {
  REQUIRE_INIT();
  if (hw % 4 != 0 || c > ca) return fail(FO_ERR_INVALID, "mse: hw must be a multiple of 4 and c <= ca");
  if ((size_t)n * c * hw == 0) return FO_OK;
  CUDA_TRY(launch_mse(a, b, n, ca, c, hw, sum_out, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
extern "C" int fo_mse_grad(const float* a, const float* b, int n, int ca, int c, int hw, const float* gscale,
                           float scale, float* grad, fo_stream_t stream) {
  REQUIRE_INIT();
  if (hw % 4 != 0 || c > ca) return fail(FO_ERR_INVALID, "mse_grad: hw must be a multiple of 4 and c <= ca");
  if ((size_t)n * ca * hw == 0) return FO_OK;
  CUDA_TRY(launch_mse_grad(a, b, n, ca, c, hw, gscale, scale, grad, g_num_sms, (cudaStream_t)stream));
  return FO_OK;
}
