// Internal (C++) interfaces between the kernel files and the engine.
#pragma once
#include "common.cuh"

namespace ddn {

struct ConvGeom {
  int N, Hin, Win, Cin;   // gather source (NHWC)
  int Hout, Wout, Cout;   // destination (NHWC)
  int KH, KW;
  int stride;             // destination -> source coordinate multiplier
  int ups;                // source up-sampling (transposed conv); 1 for forward
  int pad, dil;
  int cin_log2;
};

// conv_simt.cu
int conv_geom_init(ConvGeom* g, int N, int Hin, int Win, int Cin, int Hout, int Wout, int Cout,
                   int KH, int KW, int stride, int ups, int pad, int dil);
int launch_pack_weights(const float* w, float* wp, int Cout, int Cin, int CinP, int KH, int KW, int dgrad, cudaStream_t st);
int launch_unpack_wgrad(const float* dwp, float* dw, int Cout, int Cin, int CinP, int KH, int KW, cudaStream_t st);
int launch_conv_gather_f32(const float* in, const float* wp, const float* addend, float* out, const ConvGeom& g, cudaStream_t st);
int launch_conv_wgrad_f32(const float* in, const float* dy, float* dwp, const ConvGeom& g, cudaStream_t st);

// bn.cu -- NHWC tensors viewed as [M = N*H*W][C]
// partial buffer: float [2][nblk][C]; stats: float mean[C], invstd[C]
int bn_partial_blocks(int64_t M, int C);
int launch_bn_stats(const float* x, int64_t M, int C, float* partial, float* mean, float* invstd,
                    float* running_mean, float* running_var, float momentum, float eps, cudaStream_t st);
int launch_bn_stats_finalize(const float* partial, int nblk, int64_t M, int C, float* mean, float* invstd,
                             float* running_mean, float* running_var, float momentum, float eps, cudaStream_t st);
int launch_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps,
                         float* mean, float* invstd, cudaStream_t st);
// eval-mode BN folded to y = x*scale + shift (scale = gamma/sqrt(rv+eps), shift = beta - rm*scale)
int launch_bn_fold(const float* running_mean, const float* running_var, const float* gamma, const float* beta, int C, float eps,
                   float* scale, float* shift, cudaStream_t st);
// y = relu?( (x-mean)*invstd*gamma+beta + res ), res = r (identity) or (r-rmean)*rinvstd*rgamma+rbeta
struct BnApplyArgs {
  const float* x; const float* mean; const float* invstd; const float* gamma; const float* beta;
  const float* r; const float* rmean; const float* rinvstd; const float* rgamma; const float* rbeta;
  float* y; int64_t M; int C; int relu;
  __nv_bfloat16* hi; __nv_bfloat16* lo;   // optional: also emit y as bf16 hi/lo planes (tensor-core operands)
};
int launch_bn_apply(const BnApplyArgs& a, cudaStream_t st);
// backward of y = relu?(bn(x) + res): g = dy*(y>0); sums -> dgamma,dbeta; dx; optional g_out (= d res)
struct BnBwdArgs {
  const float* dy; const float* y; const float* x; const float* mean; const float* invstd; const float* gamma;
  float* dx; float* dgamma; float* dbeta; float* g_out; float* partial; int64_t M; int C; int relu; int training;
  __nv_bfloat16* dx_hi; __nv_bfloat16* dx_lo;   // optional: emit dx as bf16 hi/lo planes (dx itself may then be null)
  const __nv_bfloat16* y_hi;                    // optional: bf16 hi plane of y, read for the ReLU mask instead of y
};
int launch_bn_backward(const BnBwdArgs& a, cudaStream_t st);

// stem: conv1 raw [N,Hc,Wc,64] -> bn+relu+maxpool3x3/2 -> y [N,Hp,Wp,64], argmax uint8
int launch_stem_bn_relu_pool(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                             float* y, uint8_t* argmax, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo,
                             int N, int Hc, int Wc, int C, cudaStream_t st);
// dy_pool [N,Hp,Wp,C] -> g [N,Hc,Wc,C] = d(relu out) * (bn(x) > 0)   (pre-BN-backward gradient)
int launch_stem_pool_relu_backward(const float* dy_pool, const uint8_t* argmax, const float* x, const float* mean,
                                   const float* invstd, const float* gamma, const float* beta, float* g,
                                   int N, int Hc, int Wc, int C, cudaStream_t st);

// head.cu
int launch_nchw_to_nhwc4(const float* x, float* y, int N, int H, int W, cudaStream_t st);
int launch_fc_forward(const float* feat, const float* w, const float* bias, float* low, int64_t Mimg, int N, int C, int D, cudaStream_t st);
int launch_fc_backward(const float* dlow, const float* feat, const float* w, float* dfeat, float* dw, float* dbias,
                       int64_t Mimg, int N, int C, int D, cudaStream_t st);
int launch_upsample_fwd(const float* x, float* y, int NC, int h, int w, int H, int W, cudaStream_t st);
int launch_upsample_bwd(const float* dy, float* dx, int NC, int h, int w, int H, int W, cudaStream_t st);
int launch_fill_zero(void* p, size_t bytes, cudaStream_t st);

}  // namespace ddn
