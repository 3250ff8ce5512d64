// Internal launcher declarations shared between the .cu translation units and api.cu.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "igemm.cuh"

namespace fo {

struct PackStep {
  int16_t tap;    // linear filter-tap index in the PyTorch weight
  int16_t wk0;    // first index on the weight's K axis
  int16_t valid;  // valid channels in this chunk (rest are zero)
  int16_t pad;
};
struct PackParams {
  int npad, ktot, kc, cout;
  int dimB, taps, n_axis;
  const float* n_scale;
  PackStep steps[kMaxKSteps];
};
struct FinalizeParams {
  const float* partial;
  float* dweight;
  int splits, taps, MC, NC, m_real, n_real;
  int dimB, m_axis, q_w_off, accumulate;
  int8_t tap_index[64];  // partial slot -> filter tap index in the PyTorch weight
};

// conv_igemm.cu / wgrad_igemm.cu
size_t conv_smem_bytes(const ConvParams& p);
cudaError_t launch_conv_igemm(const ConvParams& p, const ConvMaps& maps, int num_sms, cudaStream_t stream);
cudaError_t init_conv_igemm();
size_t wgrad_smem_bytes(const WgradParams& p);
cudaError_t launch_wgrad_igemm(const WgradParams& p, const WgradMaps& maps, cudaStream_t stream);
cudaError_t init_wgrad_igemm();

// elementwise.cu
cudaError_t launch_pack_nchw(const float* x, void* out, int n, int c, int hw, int cs, const float* shift,
                             const float* scale, int num_sms, cudaStream_t st);
cudaError_t launch_unpack_nchw(const void* x, float* out, int n, int c, int hw, int cs, cudaStream_t st);
cudaError_t launch_relu(const void* x, void* y, size_t numel, int num_sms, cudaStream_t st);
cudaError_t launch_u8hwc_to_nchw(const void* x, float* out, int n, int hw, int c_total, int c_off, float mean, float stdv,
                                 int num_sms, cudaStream_t st);
cudaError_t launch_adam(const void* table, const void* chunks, int n_chunks, float lr, float beta1, float beta2,
                        float eps, float weight_decay, float bias_c1, float sqrt_bias_c2, float grad_scale, int num_sms,
                        cudaStream_t st);
cudaError_t launch_pack_weights(const float* w, void* out, const PackParams& pp, int num_sms, cudaStream_t st);
cudaError_t launch_wgrad_finalize(const FinalizeParams& fp, const float* bias_part, int bias_rows, int bias_c, float* dbias,
                                  int dbias_accumulate, int num_sms, cudaStream_t st);
int colsum_blocks(int num_sms);
cudaError_t launch_colsum(const void* x, size_t rows, int cs, int c_off, int c, float* out, int accumulate,
                          float* workspace, int num_sms, cudaStream_t st);
cudaError_t launch_maxpool2(const void* x, void* y, int n, int h, int w, int cs, int num_sms, cudaStream_t st);
cudaError_t launch_maxpool2_bwd(const void* x, const void* y, const void* dy, void* dx, int n, int h, int w, int cs,
                                int num_sms, cudaStream_t st);

cudaError_t launch_im2col4x4s2(const float* x, void* out, int n, int ca, int c, int h, int w, cudaStream_t st);
cudaError_t launch_im2col3x3(const float* x, void* out, int n, int c, int h, int w, const float* shift,
                             const float* scale, cudaStream_t st);
cudaError_t launch_col2im4x4s2(const void* col, const float* bias, float* out, int n, int c, int hi, int wi, int num_sms,
                               cudaStream_t st);
cudaError_t launch_chansum_nchw(const float* x, int n, int ca, int c, int hw, float* out, int accumulate, int num_sms,
                                cudaStream_t st);

// small_cin.cu
cudaError_t launch_vgg_first_conv(const float* x, int n, int h, int w, const float* weight, const float* bias,
                                  const float* shift, const float* scale, void* out, int num_sms, cudaStream_t st);

cudaError_t launch_vgg_first_dgrad(const CUtensorMap* map_dy, int n, int h, int w, const float* weight, const float* scale,
                                   float* dx, int num_sms, cudaStream_t st);

int s2_grid(int num_sms);
cudaError_t launch_s2conv(const float* x, int n, int ca, int c, int H, int W, const float* weight, const float* bias,
                          const void* mask, const void* addend, void* out, int relu, int num_sms, cudaStream_t st);
cudaError_t launch_s2wgrad(const float* x, int n, int ca, int c, int H, int W, const void* y, float* dweight, int accumulate,
                           float* dbias, int dbias_accumulate, float* workspace, int num_sms, cudaStream_t st);

// vq.cu
cudaError_t launch_vq_prep(const float* embed, int dim, int n_embed, void* e_split, float* e_t, float* e_norm2,
                           cudaStream_t st);
size_t vq_assign_workspace_bytes(size_t rows, int dim);
size_t vq_split_elems(int dim, int n_embed);
cudaError_t launch_vq_assign(const float* x, size_t rows, int dim, int n_embed, const float* e_t,
                             const void* e_split, const float* e_norm2, int64_t* embed_ind, int* n_flagged,
                             void* workspace, const CUtensorMap* map_e, const CUtensorMap* map_x, int num_sms,
                             cudaStream_t st);
bool vq_assign_is_generic(int dim, int n_embed);
cudaError_t launch_vq_assign_generic(const float* x, size_t rows, int dim, int n_embed, const float* e_t, int64_t* embed_ind,
                                     int* n_flagged, int num_sms, cudaStream_t st);
cudaError_t init_vq();
cudaError_t launch_vq_gather_stats(const float* x, const int64_t* ind, size_t rows, int dim, int n_embed,
                                   const float* e_t, float* q_f32, void* q_bf16, float* diff_sum, float* counts,
                                   float* embed_sum, float* scratch, int num_sms, cudaStream_t st);
size_t vq_gather_scratch_bytes(int dim, int n_embed);
cudaError_t launch_vq_ema(float* embed, float* cluster_size, float* embed_avg, const float* counts,
                          const float* embed_sum, int dim, int n_embed, float decay, float one_minus_decay, float eps,
                          cudaStream_t st);
cudaError_t launch_vq_backward(const void* g_q, int g_q_is_bf16, int g_cs, int g_c_off, const float* g_diff,
                               const float* x, const int64_t* ind, const float* e_t, size_t rows, int dim,
                               int n_embed, float* gx_f32, void* gx_bf16, int num_sms, cudaStream_t st);

// lpips.cu
cudaError_t launch_lpips_tap(const void* f0, const void* f1, const float* w, int n, int hw, int c, float* out,
                             int num_sms, cudaStream_t st, int split = 0);
cudaError_t launch_lpips_tap_bwd(const void* f0, const void* f1, const float* w, const float* g, int n, int hw, int c,
                                 void* d_f0, const void* addend, int num_sms, cudaStream_t st, int split = 0);

cudaError_t launch_lpips_tap_pool(const void* f0, const void* f1, const float* w, int n, int h, int wd, int c, float* out,
                                  void* pooled, void* pooled1, int num_sms, cudaStream_t st);
cudaError_t launch_lpips_tap_bwd_pool(const void* f0, const void* f1, const float* w, const float* g, int n, int h, int wd,
                                      int c, void* d_f0, const void* pool_dy, int num_sms, cudaStream_t st);

// disc.cu (MoCoGAN-HD discriminators): fp32 NCDHW direct convolution + norm / pool / loss kernels
struct DConvParams {
  int n, cin, id, ih, iw;       // input  [n, cin, id, ih, iw]   (2-D: id = 1)
  int cout, od, oh, ow;         // output [n, cout, od, oh, ow]
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
};
cudaError_t launch_dconv_fwd(const DConvParams& p, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
cudaError_t launch_dconv_dgrad(const DConvParams& p, const float* dy, const float* w, float* dx, cudaStream_t st);
cudaError_t launch_dconv_wgrad(const DConvParams& p, const float* x, const float* dy, float* dw, float* dbias, int num_sms,
                               cudaStream_t st);
// disc_gemm.cu: data movement of the tensor-core (im2col + split-bf16 GEMM) path of the discriminator convolutions
cudaError_t launch_dim2col_pairs(const DConvParams& p, const float* x, void* col, int kp, int parts, int num_sms,
                                 cudaStream_t st);
cudaError_t launch_dim2col_t(const DConvParams& p, const float* x, void* out, int kp, int pc, int chunks, int num_sms,
                             cudaStream_t st);
cudaError_t launch_dcol2im(const DConvParams& p, const float* dcol, long long ld, float* dx, int num_sms, cudaStream_t st);
cudaError_t launch_dconv_dbias(const float* dy, int n, int c, int plane, float* dbias, cudaStream_t st);
cudaError_t launch_instnorm_fwd(const float* x, float* y, int n, int c, long long plane, float eps, float slope, int training,
                                float momentum, float* running_mean, float* running_var, float* save, cudaStream_t st);
cudaError_t launch_instnorm_bwd(const float* y, const float* dy, float* dx, int n, int c, long long plane, float slope,
                                int training, const float* save, cudaStream_t st);
cudaError_t launch_lrelu(const float* x, float* y, size_t n, float slope, cudaStream_t st);
cudaError_t launch_lrelu_bwd(const float* y, const float* dy, float* dx, size_t n, float slope, cudaStream_t st);
cudaError_t launch_avgpool3(const float* x, float* y, long long planes, int id, int ih, int iw, int od, int oh, int ow, int kd,
                            int sd, int sh, int sw, cudaStream_t st);
cudaError_t launch_avgpool3_bwd(const float* dy, float* dx, long long planes, int id, int ih, int iw, int od, int oh, int ow,
                                int kd, int sd, int sh, int sw, cudaStream_t st);
cudaError_t launch_ralsgan(const float* a, int n, const float* b, int m, float target, float* out, cudaStream_t st);
cudaError_t launch_ralsgan_bwd(const float* a, int n, int m, float target, const float* fwd, const float* g, float* da,
                               float* db, cudaStream_t st);

// precise.cu (verification mode)
cudaError_t launch_split_f32(const float* x, int n, int c, int hw, long long sn, long long sc, long long sp, void* out,
                             int cp, cudaStream_t st);
cudaError_t launch_merge_f32(const void* in, int n, int c, int hw, int cp, float* out, long long sn, long long sc,
                             long long sp, cudaStream_t st);
cudaError_t launch_maxpool2_f32(const float* x, float* y, int n, int h, int w, int c, cudaStream_t st);
cudaError_t launch_maxpool2_bwd_f32(const float* x, const float* y, const float* dy, float* dx, int n, int h, int w, int c,
                                    cudaStream_t st);
cudaError_t launch_mse(const float* a, const float* b, int n, int ca, int c, int hw, float* sum_out, int num_sms,
                       cudaStream_t st);
cudaError_t launch_mse_grad(const float* a, const float* b, int n, int ca, int c, int hw, const float* gscale,
                            float scale, float* grad, int num_sms, cudaStream_t st);

}  // namespace fo
