/* faceoff_b200 -- C ABI of the B200-native FaceOff training-step hot path.
 *
 * The reference (skymanaditya1/FaceOff) is pure Python/PyTorch and has no FFI of its own; its "plugin
 * boundary" for this path is the nn.Module surface of models/vqvae_conv3d_latent.py, models/lpips.py,
 * loss.py:VQLPIPS and the distributed/ package.  the Python package faceoff_b200 re-implements that
 * surface and calls ONLY the functions below (via ctypes): plain pointers, sizes and a cudaStream_t.
 * PyTorch owns every buffer; nothing here allocates persistent device memory.
 *
 * All functions return 0 on success, else a non-zero code; fo_last_error() gives the message
 * (thread-local).  There is no CPU fallback: every entry point fails with FO_ERR_NO_DEVICE when no
 * sm_100 device is present.
 *
 * Layouts: activations are channels-last bf16: [N, (D,) H, W, Cs] with Cs (storage channels per pixel) a
 * multiple of 16 (logical channels beyond C are zero).  Weights stay in the PyTorch fp32 layouts
 * ([Cout,Cin,k..] for Conv, [Cin,Cout,kh,kw] for ConvTranspose2d) and are re-packed to bf16 K-step order by
 * fo_conv_pack_weights into caller-provided workspace.
 */
#ifndef FACEOFF_B200_H
#define FACEOFF_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fo_stream_t; /* cudaStream_t */

enum {
  FO_OK = 0,
  FO_ERR_INVALID = 1,   /* bad argument / unsupported geometry */
  FO_ERR_CUDA = 2,      /* CUDA runtime or driver error */
  FO_ERR_NO_DEVICE = 3, /* no sm_100 GPU: there is no CPU path */
};

const char* fo_last_error(void);
int fo_version(void);
/* Must be called once per process after the CUDA context exists (sets kernel attributes, resolves
 * cuTensorMapEncodeTiled). */
int fo_init(void);

/* ---------------------------------------------------------------------------------------------
 * Convolution family (implicit GEMM on tcgen05).  Replaces, for this path, torch.nn.functional
 * conv2d / conv_transpose2d / conv3d forward and their autograd (cuDNN fprop / dgrad / wgrad) as called
 * from reference models/vqvae_conv3d_latent.py:86-190,208-217 and models/lpips.py:115-152.
 * ------------------------------------------------------------------------------------------- */

/* "form" of the implicit GEMM */
enum {
  FO_FORM_S1 = 0,      /* stride-1 conv, k in {1,3}, pad (k-1)/2, 2-D or 3-D              (Conv2d/Conv3d forward) */
  FO_FORM_S1_DGRAD = 1,/* same geometry with mirrored taps                                 (their data gradient)   */
  FO_FORM_DOWN = 2,    /* 4x4 stride-2 pad-1 gather form: out[o] = sum in[2o+k-1] w[k]     (Conv2d s2 forward, ConvTranspose2d dgrad) */
  FO_FORM_UP = 3,      /* 4x4 stride-2 pad-1 scatter form: out[2i-1+k] += in[i] w[k]       (ConvTranspose2d forward, Conv2d s2 dgrad) */
};

typedef struct {
  const void* ptr; /* bf16 channels-last */
  int c;           /* logical channels taken from this source */
  int cs;          /* storage channels per pixel */
  int c_off;       /* first channel inside the pixel (for reading a slice of a wider tensor) */
} fo_src_t;

typedef struct {
  int form;
  int ndim;  /* 2 or 3 (3 only for FO_FORM_S1 / FO_FORM_S1_DGRAD) */
  int ksize; /* 1 or 3 for S1 forms; ignored (4) for DOWN / UP */
  /* input-side extents (the tensor the A operand is read from).  2-D: n = frames, d = 1. */
  int n, d, h, w;
  int n_src; /* 1..6: torch.cat([src0, src1, ...], dim=1) folded into the K loop (the hi|lo verification mode lists
                each tensor three times: hi, lo, hi -- see split_out) */
  fo_src_t src[6];
  int cout; /* logical output channels (N of the GEMM) */
  /* packed weights produced by fo_conv_pack_weights for the SAME descriptor */
  const void* wpacked;
  /* fused epilogue: r = acc + bias; r = mask>0 ? r : 0; r += addend; outputs below (any subset) */
  const float* bias;      /* [cout] fp32 or NULL */
  const void* mask;       /* bf16, layout of out_bf16, or NULL */
  const void* addend;     /* bf16, layout of out_bf16, or NULL */
  void* out_bf16;         /* r            -> bf16 channels-last [.., out_cs] or NULL */
  void* out_relu;         /* relu(r)      -> bf16 channels-last [.., out_cs] or NULL */
  float* out_f32;         /* r (or relu)  -> fp32, channels-last [.., out_cs] (out_f32_nchw=0) or NCHW [n,cout,h,w] (=1) */
  int out_cs;             /* storage channels of the channels-last outputs */
  int out_f32_nchw;
  int relu_f32;
  /* Verification mode (tests only): every bf16 tensor of the epilogue (mask, addend, out_bf16, out_relu) holds an fp32
   * value as a pair of bf16 numbers, hi = bf16(v) in channels [0, out_cs/2) and lo = bf16(v - hi) in [out_cs/2, out_cs).
   * Together with sources listed as (hi, lo, hi) against weights packed as (w_hi, w_hi, w_lo) the SAME kernels compute
   * the convolution to ~2^-16 relative accuracy, which lets parity tests compare gradients with an fp64 CPU restatement at 1e-4
   * instead of the bf16 tolerance.  0 = off (the product path). */
  int split_out;
  /* out_f32 += result instead of out_f32 = result: a GEMM whose K dimension is split over several launches (the
   * discriminator convolutions below).  0 = overwrite. */
  int out_f32_accumulate;
} fo_conv_t;

/* Bytes of packed-weight workspace needed for this descriptor. */
size_t fo_conv_wpacked_bytes(const fo_conv_t* c);
/* Re-pack fp32 PyTorch-layout weights [A][B][taps] into bf16 K-step order.
 * n_axis: which of the first two weight axes is the GEMM N (output) axis: Conv fwd 0, Conv dgrad 1,
 * ConvTranspose fwd 1, ConvTranspose dgrad 0.  dimA/dimB are the sizes of those two axes.
 * n_scale: optional fp32 [cout] multiplier per output channel (folds LPIPS' 1/scale into the first VGG dgrad). */
int fo_conv_pack_weights(const fo_conv_t* c, const float* weight, int dimA, int dimB, int n_axis,
                         const float* n_scale, void* wpacked, fo_stream_t stream);
int fo_conv_run(const fo_conv_t* c, fo_stream_t stream);

/* Weight gradient: dW[m][n][tap] = sum_pix P[pix][m] * Q[pix (+) tap][n].
 * form FO_FORM_S1 (P = dy, Q = x), FO_FORM_DOWN (P = dy low-res, Q = x hi-res) -> Conv weight [m=cout][n=cin][taps];
 * FO_FORM_UP (P = x low-res, Q = dy hi-res) -> ConvTranspose2d weight [m=cin][n=cout][taps]. */
typedef struct {
  int form;
  int ndim, ksize;
  int n, d, h, w;       /* extents of P (the non-shifted operand) */
  fo_src_t p, q;
  float* dweight;       /* fp32 PyTorch-layout gradient tensor (written, or accumulated if accumulate != 0) */
  int dimA, dimB;       /* sizes of the first two axes of dweight */
  int m_axis;           /* which weight axis the P channels index (0 or 1) */
  int q_w_off;          /* channel offset of Q's channels on the other weight axis (torch.cat second source) */
  int q_shift_sign;     /* +1: Q is read at pix + tap (P = dy, Q = x); -1: at pix - tap (P = x, Q = dy; lets the wider
                           tensor be the M side of the MMA).  0 is treated as +1. */
  int accumulate;
  float* dbias;         /* optional: fp32 [p.c] bias gradient = column sums of P (fused: one extra N=16 MMA per K step) */
  int dbias_accumulate;
  void* workspace;      /* split-K partials */
  size_t workspace_bytes;
} fo_wgrad_t;
size_t fo_wgrad_workspace_bytes(const fo_wgrad_t* g);
int fo_wgrad_run(const fo_wgrad_t* g, fo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Layout / elementwise helpers
 * ------------------------------------------------------------------------------------------- */
/* NCHW fp32 -> channels-last bf16 [n, hw, cs]; channels >= c zero; optional per-channel (x - shift)/scale
 * (LPIPS ScalingLayer, reference models/lpips.py:96-103). */
int fo_pack_nchw(const float* x, void* out, int n, int c, int hw, int cs, const float* shift, const float* scale,
                 fo_stream_t stream);
/* channels-last bf16 [n, hw, cs] -> NCHW fp32 [n, c, hw] */
int fo_unpack_nchw(const void* x, float* out, int n, int c, int hw, int cs, fo_stream_t stream);
/* Input pipeline on the GPU: uint8 HWC frames [n, hw, 3] -> channels c_off..c_off+2 of an fp32 NCHW tensor
 * [n, c_total, hw] as ((x / 255) - mean) / std -- torchvision ToTensor + Normalize(0.5, 0.5) of the reference's loader
 * (TemporalAlignment/dataset.py:235-249), bit-exact (IEEE fp32 division / subtraction in the same order); calling it for
 * the source frames (c_off 0) and the background frames (c_off 3) of a 6-channel tensor is utils.process_data's
 * torch.cat([source, background]) (utils.py:29-38).  hw must be a multiple of 4. */
int fo_u8hwc_to_nchw(const void* x, float* out, int n, int hw, int c_total, int c_off, float mean, float stdv,
                     fo_stream_t stream);
/* y = relu(x) on bf16 */
int fo_relu(const void* x, void* y, size_t numel, fo_stream_t stream);
/* out[c] (+)= sum_rows x[row][c_off + c], x bf16 [rows, cs]  (bias gradients) */
int fo_colsum(const void* x, size_t rows, int cs, int c_off, int c, float* out, int accumulate, void* workspace,
              size_t workspace_bytes, fo_stream_t stream);
size_t fo_colsum_workspace_bytes(int cs);
/* Explicit im2col of a k4 s2 p1 window for tiny channel counts (c <= 8): x NCHW fp32 [n, ca, h, w] (first c channels)
 * -> bf16 [n, h/2, w/2, 128], k = (ky*4+kx)*8 + ch.  Turns the first Conv2d(6->64,4,2,1) (reference
 * models/vqvae_conv3d_latent.py:109) and the gradient of the last ConvTranspose2d(64->6,4,2,1) (:154-156) into plain
 * K=128 GEMMs on the tcgen05 kernel. */
int fo_im2col4x4s2(const float* x, void* out, int n, int ca, int c, int h, int w, fo_stream_t stream);
/* 3x3 pad-1 im2col for the first VGG16 conv of LPIPS (c <= 3): x NCHW fp32 -> bf16 [n, h, w, 32], k = (ky*3+kx)*3 + ch,
 * with the ScalingLayer (x - shift) / scale folded in (reference models/lpips.py:96-103,119). */
int fo_im2col3x3(const float* x, void* out, int n, int c, int h, int w, const float* shift, const float* scale,
                 fo_stream_t stream);
/* The same layer in one kernel (the product path): ScalingLayer + Conv2d(3 -> 64, 3x3, pad 1) + ReLU from the fp32 NCHW
 * image x [n, 3, h, w] to the bf16 channels-last activation out_relu [n, h, w, 64]; weight [64, 3, 3, 3] / bias [64] are
 * the PyTorch parameters (reference models/lpips.py:96-103,119-127).  The A tile is built in shared memory, no im2col
 * matrix reaches HBM. */
int fo_vgg_first_conv(const float* x, int n, int h, int w, const float* weight, const float* bias, const float* shift,
                      const float* scale, void* out_relu, fo_stream_t stream);
/* Data gradient of that convolution (the gradient leaving the LPIPS trunk): dy bf16 channels-last [n,h,w,64] (gradient
 * w.r.t. the pre-ReLU output), weight fp32 [64,3,3,3] -> dx fp32 NCHW [n,3,h,w], divided by scale[c] if scale != NULL
 * (ScalingLayer backward, models/lpips.py:96-103).  One K = 64 GEMM per pixel tile with the nine taps on the N axis + an
 * in-tile shift-add (csrc/small_cin.cu). */
int fo_vgg_first_dgrad(const void* dy, int n, int h, int w, const float* weight, const float* scale, float* dx,
                       fo_stream_t stream);
/* Image-side layers of the VQVAE without an im2col matrix (csrc/small_cin.cu; reference models/vqvae_conv3d_latent.py:109
 * first Conv2d(c -> 64, 4, stride 2, pad 1), :154-156 last ConvTranspose2d(64 -> c, 4, stride 2, pad 1)); c in {3, 6}.
 * x: fp32 NCHW [n, ca, H, W] (first c channels used), H and W even; all bf16 tensors are channels-last [n, H/2, W/2, 64].
 *   fo_s2conv : out = f(conv4x4s2(x, weight [64, c, 4, 4]) + bias), f = (mask > 0 ? . : 0) + addend, then ReLU if relu != 0.
 *               First conv forward: bias, relu = 1.  Data gradient of the last ConvTranspose2d w.r.t. its 64-channel input:
 *               x = the fp32 NCHW output gradient, weight = the ConvTranspose2d parameter [64, c, 4, 4] as it is,
 *               mask = the layer input (ReLU gate), relu = 0.
 *   fo_s2wgrad: dweight [64, c, 4, 4] (+)= sum_pixels y[pixel, :] (x) im2col(x)[pixel, :], dbias [64] (+)= sum_pixels y
 *               (dbias may be NULL).  First conv: y = dy.  Last ConvTranspose2d: y = the layer input, x = the output gradient. */
int fo_s2conv(const float* x, int n, int ca, int c, int H, int W, const float* weight, const float* bias, const void* mask,
              const void* addend, void* out, int relu, fo_stream_t stream);
size_t fo_s2wgrad_workspace_bytes(void);
int fo_s2wgrad(const float* x, int n, int ca, int c, int H, int W, const void* y, float* dweight, int accumulate, float* dbias,
               int dbias_accumulate, void* workspace, size_t workspace_bytes, fo_stream_t stream);
/* Inverse scatter for the last ConvTranspose2d: col bf16 [n, hi, wi, 128] (k = tap*8 + co) + bias -> NCHW fp32
 * [n, c, 2hi, 2wi]. */
int fo_col2im4x4s2(const void* col, const float* bias, float* out, int n, int c, int hi, int wi, fo_stream_t stream);
/* out[ch] (+)= sum_{n,hw} x[n][ch][hw] for ch < c, x NCHW fp32 [n, ca, hw] (bias gradient of the last layer) */
int fo_chansum_nchw(const float* x, int n, int ca, int c, int hw, float* out, int accumulate, fo_stream_t stream);
/* 2x2/2 max pool on channels-last bf16 [n,h,w,cs] and its gradient (reference VGG trunk models/lpips.py:115-152) */
int fo_maxpool2(const void* x, void* y, int n, int h, int w, int cs, fo_stream_t stream);
/* dx is the gradient w.r.t. the PRE-ReLU conv output feeding the pool (x is post-ReLU: x == 0 closes the gate). */
int fo_maxpool2_bwd(const void* x, const void* y, const void* dy, void* dx, int n, int h, int w, int cs,
                    fo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Vector quantiser (reference models/vqvae_conv3d_latent.py:33-83)
 * ------------------------------------------------------------------------------------------- */
/* Codebook prep: embed fp32 [dim, n_embed] -> e_split bf16, fo_vq_split_elems(dim, n_embed) elements: the split codebook
 * [n_embed, 2*dim] (hi | lo halves, the GEMM B operand) followed by the augmented K slice [n_pad, 16] (n_pad = n_embed
 * rounded up to 256; per code -|e|^2/2 as three bf16 terms and |e| rounded up: one extra K=16 MMA folds |e|^2 and the error
 * band into the accumulator); e_t fp32 [n_embed, dim] (transposed copy: the gather operand; keep it for fo_vq_backward
 * because the EMA update overwrites embed inside forward), e_norm2 fp32 [n_embed + 1] (|e_k|^2, then max_k |e_k|^2). */
size_t fo_vq_split_elems(int dim, int n_embed);
int fo_vq_prep(const float* embed, int dim, int n_embed, void* e_split, float* e_t, float* e_norm2,
               fo_stream_t stream);
/* Nearest-code assignment (:48-54).  x fp32 [rows, dim].  embed_ind int64 [rows].
 * e_t / e_split / e_norm2 come from fo_vq_prep.  Tensor-core distances (bf16 split, error-bounded per code) + exact
 * (fp64-accumulated) re-evaluation of the rows that are ambiguous within the bound; n_flagged (device int32, optional) counts those rows.
 * That kernel takes dim 64 / 128 and n_embed % 16 == 0, <= 8192; every other shape the reference's Quantize accepts (:34-45;
 * dim <= 1536) is scored exactly in fp64 on CUDA cores (same first-minimum rule, n_flagged = 0, workspace unused). */
int fo_vq_assign(const float* x, size_t rows, int dim, int n_embed, const float* e_t, const void* e_split,
                 const float* e_norm2, int64_t* embed_ind, int* n_flagged, void* workspace, size_t workspace_bytes,
                 fo_stream_t stream);
size_t fo_vq_assign_workspace_bytes(size_t rows, int dim);
/* Gather + straight-through + commitment loss + EMA statistics (:55-61,77-78), one pass:
 *   quantize = x + (E[:, ind] - x)         -> q_f32 (optional) and q_bf16 (optional, channels-last copy)
 *   diff_sum += sum (E[:, ind] - x)^2      (caller divides by rows*dim)
 *   counts[k] += #rows assigned to k ; embed_sum[d][k] += sum of x rows assigned to k   (if counts != NULL)
 * scratch: fo_vq_gather_scratch_bytes(dim, n_embed) bytes (0 when the per-CTA statistics fit in shared memory), 16-byte
 * aligned, or NULL.  Large codebooks accumulate into it in transposed order with 16-byte vector atomics and fold it into
 * embed_sum at the end; with NULL they fall back to scalar atomics on embed_sum (same result, slower). */
size_t fo_vq_gather_scratch_bytes(int dim, int n_embed);
int fo_vq_gather_stats(const float* x, const int64_t* embed_ind, size_t rows, int dim, int n_embed,
                       const float* e_t, float* q_f32, void* q_bf16, float* diff_sum, float* counts,
                       float* embed_sum, float* scratch, fo_stream_t stream);
/* EMA update + renormalisation (:66-75), in place on the three buffers; counts / embed_sum are the
 * (all-reduced) statistics.  one_minus_decay is passed separately because the reference forms ``alpha = 1 - decay`` in
 * double precision before rounding it to fp32 (1.f - 0.99f differs from (float)0.01 by 1e-6 relative). */
int fo_vq_ema(float* embed, float* cluster_size, float* embed_avg, const float* counts, const float* embed_sum,
              int dim, int n_embed, float decay, float one_minus_decay, float eps, fo_stream_t stream);
/* Backward of :77-78: gx = g_q + g_diff * 2 (x - q) / (rows*dim).  g_q fp32 or bf16 (g_q_is_bf16). */
int fo_vq_backward(const void* g_q, int g_q_is_bf16, int g_cs, int g_c_off, const float* g_diff, const float* x,
                   const int64_t* embed_ind, const float* e_t, size_t rows, int dim, int n_embed, float* gx_f32,
                   void* gx_bf16, fo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * LPIPS head (reference models/lpips.py:80-93,155-161) and MSE (train_faceoff_perceptual.py:37-39)
 * ------------------------------------------------------------------------------------------- */
/* One tap: per image n, out[n] += (1/hw) * sum_pix sum_c w[c] * (f0/(|f0|+eps) - f1/(|f1|+eps))^2.
 * f0, f1 channels-last bf16 [n, hw, c]. */
int fo_lpips_tap(const void* f0, const void* f1, const float* w, int n, int hw, int c, float* out, fo_stream_t stream);
/* Gradient wrt f0: d_f0 (bf16, same layout), scaled by g[n] (fp32 per image), masked by f0 > 0 (ReLU tap). */
int fo_lpips_tap_bwd(const void* f0, const void* f1, const float* w, const float* g, int n, int hw, int c, void* d_f0,
                     const void* addend, fo_stream_t stream);
/* fo_lpips_tap for feature maps [n, h, w, c] that feed a 2x2/2 max pool: also writes pooled = maxpool2(f0) and, if
 * pooled1 != NULL, pooled1 = maxpool2(f1) ([n, h/2, w/2, c], bit-identical to fo_maxpool2): the pools do not read the
 * feature maps again. */
int fo_lpips_tap_pool(const void* f0, const void* f1, const float* w, int n, int h, int wd, int c, float* out, void* pooled,
                      void* pooled1, fo_stream_t stream);
/* The same gradient for a tap that feeds a 2x2/2 max pool (relu1_2 .. relu4_3), with the pool's backward folded in:
 * f0, f1, d_f0 channels-last bf16 [n, h, w, c]; pool_dy = gradient w.r.t. the pooled tensor [n, h/2, w/2, c].
 * Bit-identical to fo_maxpool2_bwd followed by fo_lpips_tap_bwd(addend = its result), half the HBM traffic. */
int fo_lpips_tap_bwd_pool(const void* f0, const void* f1, const float* w, const float* g, int n, int h, int wd, int c,
                          void* d_f0, const void* pool_dy, fo_stream_t stream);
/* Verification mode (tests only; see fo_conv_t.split_out): the same two kernels on hi|lo pair tensors [n, hw, 2c]
 * (feature = hi + lo; d_f0 and addend are pairs too). */
int fo_lpips_tap_split(const void* f0, const void* f1, const float* w, int n, int hw, int c, float* out,
                       fo_stream_t stream);
int fo_lpips_tap_bwd_split(const void* f0, const void* f1, const float* w, const float* g, int n, int hw, int c,
                           void* d_f0, const void* addend, fo_stream_t stream);
/* Verification mode helpers: fp32 <-> hi|lo bf16 pairs (out / in: bf16 channels-last [n*hw, 2*cp], hi in [0, cp), lo in
 * [cp, 2cp)).  Element (n, ch, p) of the fp32 tensor is x[n*sn + ch*sc + p*sp] (NCHW: sn = C*hw, sc = hw, sp = 1;
 * channels-last: sn = hw*Cs, sc = 1, sp = Cs).  fp32 channels-last 2x2/2 max pool + gradient (first maximum wins, ReLU
 * gate x > 0 fused like fo_maxpool2_bwd) for the LPIPS trunk in that mode. */
int fo_split_f32(const float* x, int n, int c, int hw, long long sn, long long sc, long long sp, void* out, int cp,
                 fo_stream_t stream);
int fo_merge_f32(const void* in, int n, int c, int hw, int cp, float* out, long long sn, long long sc, long long sp,
                 fo_stream_t stream);
int fo_maxpool2_f32(const float* x, float* y, int n, int h, int w, int c, fo_stream_t stream);
int fo_maxpool2_bwd_f32(const float* x, const float* y, const float* dy, float* dx, int n, int h, int w, int c,
                        fo_stream_t stream);

/* Reconstruction loss (reference train_faceoff_perceptual.py:38-40: nn.MSELoss()(out[:, :3], gt)).
 * fo_mse: *sum_out += sum((a[:, :c] - b)^2) over NCHW fp32 a [n, ca, hw] and b [n, c, hw] (hw % 4 == 0); the caller
 * zeroes sum_out and divides by n*c*hw.
 * fo_mse_grad: grad [n, ca, hw] = (*gscale) * scale * (a[:, :c] - b), channels >= c zero -- the gradient of the mean with
 * scale = 2 / (n*c*hw) and *gscale the upstream gradient (a device scalar: no host synchronisation). */
int fo_mse(const float* a, const float* b, int n, int ca, int c, int hw, float* sum_out, fo_stream_t stream);
int fo_mse_grad(const float* a, const float* b, int n, int ca, int c, int hw, const float* gscale, float scale,
                float* grad, fo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * MoCoGAN-HD discriminator step (reference TemporalAlignment/models/mocoganhd_content_disc.py:8-165,
 * mocoganhd_video_disc.py:8-176, mocoganhd_losses.py:109-126, disc_trainers/train_vqvae_perceptual_mocoganhd_disc.py:160-333).
 * fp32 tensors in the PyTorch-native NC(D)HW layout.  Replaces, for this path, torch's conv2d / conv3d (+ autograd),
 * instance_norm, leaky_relu, avg_pool2d / avg_pool3d and the MSELoss of the relativistic average LSGAN.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int n, cin, id, ih, iw; /* input  [n, cin, id, ih, iw] (2-D convolution: id = kd = sd = 1, pd = 0) */
  int cout, od, oh, ow;   /* output [n, cout, od, oh, ow], o = (i + 2 p - k) / s + 1 (checked) */
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
} fo_dconv_t;
/* y = conv(x, w) + bias;  w [cout, cin, kd, kh, kw], bias [cout] or NULL */
int fo_dconv_fwd(const fo_dconv_t* d, const float* x, const float* w, const float* bias, float* y, fo_stream_t stream);
/* dx = gradient of the convolution w.r.t. its input */
int fo_dconv_dgrad(const fo_dconv_t* d, const float* dy, const float* w, float* dx, fo_stream_t stream);
/* dw (overwritten) = gradient w.r.t. the weight, dbias (overwritten, optional) = sum of dy over n and positions */
int fo_dconv_wgrad(const fo_dconv_t* d, const float* x, const float* dy, float* dw, float* dbias, fo_stream_t stream);
/* Tensor-core path of the same convolutions: im2col + GEMMs on fo_conv_run (1x1 form) in the hi|lo split-bf16 arithmetic
 * (split_out / verification mode: fp32-accurate).  These four only move data; k = ((ci * kd + a) * kh + b) * kw + c.
 *   fo_dconv_im2col_pairs : col [n * P, parts * kp] bf16, term j = bf16(v - sum of the earlier terms) in [j kp, (j+1) kp);
 *                           parts = 2 (16 mantissa bits) or 3 (24 bits: the forward GEMM); kp % 8 == 0, kp >= K
 *   fo_dconv_im2col_t     : the transposed matrix as the WEIGHT operand of the weight-gradient GEMM, already in
 *                           fo_conv_run's packed layout for the sources (hi, lo, hi): out [chunks][kp][3 * pc] bf16 =
 *                           (hi | hi | lo), chunk = (n * P + pos) / pc, zero past the last position and for k >= K
 *   fo_dconv_col2im       : dx (NCDHW) = gather of dcol [n * P, ld] fp32 over the taps that reach each input element
 *   fo_dconv_dbias        : dbias [cout] = sum of dy over n and positions */
int fo_dconv_im2col_pairs(const fo_dconv_t* d, const float* x, void* col, int kp, int parts, fo_stream_t stream);
int fo_dconv_im2col_t(const fo_dconv_t* d, const float* x, void* out, int kp, int pc, int chunks, fo_stream_t stream);
int fo_dconv_col2im(const fo_dconv_t* d, const float* dcol, long long ld, float* dx, fo_stream_t stream);
int fo_dconv_dbias(const fo_dconv_t* d, const float* dy, float* dbias, fo_stream_t stream);
/* InstanceNorm (affine=False) fused with LeakyReLU(slope) (slope = 1: no activation): y = lrelu((x - mean) / sqrt(var + eps)),
 * per (n, c) plane of `plane` elements.  training != 0: instance statistics; running_mean / running_var (optional)
 * are updated with `momentum` like nn.InstanceNorm*d(track_running_stats=True) (unbiased variance, averaged over n).
 * training == 0: the running statistics are used.  save [2 * n * c] receives (mean, rstd) for the backward kernel. */
int fo_instnorm_fwd(const float* x, float* y, int n, int c, long long plane, float eps, float slope, int training,
                    float momentum, float* running_mean, float* running_var, float* save, fo_stream_t stream);
int fo_instnorm_bwd(const float* y, const float* dy, float* dx, int n, int c, long long plane, float slope, int training,
                    const float* save, fo_stream_t stream);
int fo_lrelu(const float* x, float* y, size_t numel, float slope, fo_stream_t stream);
int fo_lrelu_bwd(const float* y, const float* dy, float* dx, size_t numel, float slope, fo_stream_t stream);
/* AvgPool with kernel (kd, 3, 3), kd in {1, 3}, stride (sd, sh, sw), padding (kd / 2, 1, 1), count_include_pad=False
 * (reference mocoganhd_content_disc.py:74-77, mocoganhd_video_disc.py:80-89).  x [planes, id, ih, iw]. */
int fo_avgpool3(const float* x, float* y, long long planes, int id, int ih, int iw, int od, int oh, int ow, int kd, int sd,
                int sh, int sw, fo_stream_t stream);
int fo_avgpool3_bwd(const float* dy, float* dx, long long planes, int id, int ih, int iw, int od, int oh, int ow, int kd,
                    int sd, int sh, int sw, fo_stream_t stream);
/* Relativistic average LSGAN term (mocoganhd_losses.py:109-126): out[0] = mean((a - mean(b) - target)^2), out[1] = mean(b).
 * Backward: da [n] and db [m] (either may be NULL) for the upstream gradient *g (device scalar). */
int fo_ralsgan(const float* a, int n, const float* b, int m, float target, float* out, fo_stream_t stream);
int fo_ralsgan_bwd(const float* a, int n, int m, float target, const float* fwd, const float* g, float* da, float* db,
                   fo_stream_t stream);

/* Optimizer step (reference optim.Adam(model.parameters(), lr=3e-4), train_faceoff_perceptual.py:190,
 * torch.optim.Adam semantics without amsgrad): one launch over all parameter tensors.
 * table_dev : DEVICE array of n tensors; chunks_dev : DEVICE array of n_chunks (tensor index, chunk index) int pairs, a
 * chunk being fo_adam_chunk_elems() consecutive elements of that tensor.  step counts from 1.  grad_scale multiplies the
 * gradients first (e.g. 1/world after a SUM all-reduce). */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long long numel;
} fo_adam_tensor_t;
int fo_adam_chunk_elems(void);
int fo_adam_step(const fo_adam_tensor_t* table_dev, const int* chunks_dev, int n_chunks, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int step, float grad_scale, fo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FACEOFF_B200_H */
