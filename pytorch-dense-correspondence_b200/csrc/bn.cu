// BatchNorm2d (+residual, +ReLU) forward/backward on NHWC fp32 viewed as [M][C], and the stem's fused
// BN + ReLU + MaxPool(3,2,1).  All HBM-bound elementwise / column-reduction kernels.
//
// Reference semantics: nn.BatchNorm2d / nn.ReLU(inplace) / nn.MaxPool2d(3,2,1) as wired by
// PSD/vision/torchvision/models/resnet.py:53-69 (BasicBlock.forward) and :231-236 (stem):
// training = batch mean and BIASED variance over (N,H,W), eps inside the sqrt, running statistics updated
// with momentum and the UNBIASED variance; eval = running statistics.
#include "conv.cuh"

namespace ddn {

constexpr int BN_THREADS = 256;

// x -> (hi, lo) bf16 with x ~= hi + lo; 4 values -> two 8-byte stores
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t i4, float4 v) {
  __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
  __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
  uint2 ho; ho.x = *reinterpret_cast<uint32_t*>(&a); ho.y = *reinterpret_cast<uint32_t*>(&b);
  reinterpret_cast<uint2*>(hi)[i4] = ho;
  if (lo) {
    __nv_bfloat162 c = __halves2bfloat162(__float2bfloat16_rn(v.x - __bfloat162float(h0)), __float2bfloat16_rn(v.y - __bfloat162float(h1)));
    __nv_bfloat162 d = __halves2bfloat162(__float2bfloat16_rn(v.z - __bfloat162float(h2)), __float2bfloat16_rn(v.w - __bfloat162float(h3)));
    uint2 l2; l2.x = *reinterpret_cast<uint32_t*>(&c); l2.y = *reinterpret_cast<uint32_t*>(&d);
    reinterpret_cast<uint2*>(lo)[i4] = l2;
  }
}

static inline int bn_rows_per_iter(int C) { return BN_THREADS / (C / 4); }

int bn_partial_blocks(int64_t M, int C) {
  int rpi = bn_rows_per_iter(C);
  int64_t target = (int64_t)num_sms() * 4;
  int64_t iters = std::max<int64_t>(4, ceil_div(M, (int64_t)rpi * target));
  return (int)ceil_div(M, (int64_t)rpi * iters);
}
static inline int64_t bn_rows_per_block(int64_t M, int C) {
  int rpi = bn_rows_per_iter(C);
  int64_t target = (int64_t)num_sms() * 4;
  int64_t iters = std::max<int64_t>(4, ceil_div(M, (int64_t)rpi * target));
  return (int64_t)rpi * iters;
}

// column sums of v0(x) and v1(x) over a row chunk -> partial[0][blk][C], partial[1][blk][C]
// MODE 0: (x, x^2)      MODE 1: (g, g*xhat) with g = dy * (y > 0 if relu)
// 4 packed bf16 values > 0 ?  (sign bit clear and not +/-0)
__device__ __forceinline__ void mask_from_bf16x4(uint2 h, float4& g) {
  if ((h.x & 0x8000u) || !(h.x & 0x7fffu)) g.x = 0.f;
  if ((h.x & 0x80000000u) || !(h.x & 0x7fff0000u)) g.y = 0.f;
  if ((h.y & 0x8000u) || !(h.y & 0x7fffu)) g.z = 0.f;
  if ((h.y & 0x80000000u) || !(h.y & 0x7fff0000u)) g.w = 0.f;
}

template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_colsum_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                 const __nv_bfloat16* __restrict__ y_hi,
                 const float* __restrict__ mean, const float* __restrict__ invstd,
                 int64_t M, int C, int64_t rows_per_block, int relu, float* __restrict__ partial) {
  const int q = C >> 2;
  const int cq = threadIdx.x % q, rr = threadIdx.x / q, rpi = BN_THREADS / q;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float4 s0 = make_float4(0, 0, 0, 0), s1 = make_float4(0, 0, 0, 0);
  float4 mu = make_float4(0, 0, 0, 0), is = make_float4(1, 1, 1, 1);
  if (MODE == 1) {
    mu = reinterpret_cast<const float4*>(mean)[cq];
    is = reinterpret_cast<const float4*>(invstd)[cq];
  }
  for (int64_t r = r0 + rr; r < r1; r += rpi) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C) + cq);
    if (MODE == 0) {
      s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
      s1.x = fmaf(v.x, v.x, s1.x); s1.y = fmaf(v.y, v.y, s1.y); s1.z = fmaf(v.z, v.z, s1.z); s1.w = fmaf(v.w, v.w, s1.w);
    } else {
      float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * C) + cq);
      if (relu) {
        if (y_hi) {     // the bf16 "hi" plane of y has y's sign and zero-ness: half the bytes of the fp32 copy
          mask_from_bf16x4(__ldg(reinterpret_cast<const uint2*>(y_hi + r * C) + cq), g);
        } else {
          float4 o = __ldg(reinterpret_cast<const float4*>(y + r * C) + cq);
          g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
        }
      }
      s0.x += g.x; s0.y += g.y; s0.z += g.z; s0.w += g.w;
      s1.x = fmaf(g.x, (v.x - mu.x) * is.x, s1.x); s1.y = fmaf(g.y, (v.y - mu.y) * is.y, s1.y);
      s1.z = fmaf(g.z, (v.z - mu.z) * is.z, s1.z); s1.w = fmaf(g.w, (v.w - mu.w) * is.w, s1.w);
    }
  }
  __shared__ float4 sh0[BN_THREADS], sh1[BN_THREADS];
  sh0[threadIdx.x] = s0; sh1[threadIdx.x] = s1;
  __syncthreads();
  if (rr == 0) {
    for (int k = 1; k < rpi; ++k) {
      float4 a = sh0[k * q + cq], b = sh1[k * q + cq];
      s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
      s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
    }
    int64_t nblk = gridDim.x;
    reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * C)[cq] = s0;
    reinterpret_cast<float4*>(partial + (nblk + blockIdx.x) * C)[cq] = s1;
  }
}

// Column sums of the per-block partials in fp64.  Block = 4 channels (one 16-byte segment per partial row) x 64 slices;
// every thread walks nblk/64 rows with 8 independent loads in flight, so even ~1200 partial rows are ~5 dependent rounds
// (these kernels are pure load latency: 2-128 CTAs, a few KB to a few MB of partials).
constexpr int FIN_CH = 4, FIN_SLICES = 64;
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partial, int nblk, int C, int c, int slice,
                                                double& s, double& ss) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, q0 = 0, q1 = 0, q2 = 0, q3 = 0;
  if (c < C) {
    const float* ps = partial + c;
    const float* pq = partial + (int64_t)nblk * C + c;
    int b = slice;
    for (; b + 3 * FIN_SLICES < nblk; b += 4 * FIN_SLICES) {
      float a0 = __ldg(ps + (int64_t)b * C), a1 = __ldg(ps + (int64_t)(b + FIN_SLICES) * C);
      float a2 = __ldg(ps + (int64_t)(b + 2 * FIN_SLICES) * C), a3 = __ldg(ps + (int64_t)(b + 3 * FIN_SLICES) * C);
      float b0 = __ldg(pq + (int64_t)b * C), b1 = __ldg(pq + (int64_t)(b + FIN_SLICES) * C);
      float b2 = __ldg(pq + (int64_t)(b + 2 * FIN_SLICES) * C), b3 = __ldg(pq + (int64_t)(b + 3 * FIN_SLICES) * C);
      s0 += (double)a0; s1 += (double)a1; s2 += (double)a2; s3 += (double)a3;
      q0 += (double)b0; q1 += (double)b1; q2 += (double)b2; q3 += (double)b3;
    }
    for (; b < nblk; b += FIN_SLICES) { s0 += (double)__ldg(ps + (int64_t)b * C); q0 += (double)__ldg(pq + (int64_t)b * C); }
  }
  s = (s0 + s1) + (s2 + s3); ss = (q0 + q1) + (q2 + q3);
  __shared__ double sh[2][FIN_SLICES][FIN_CH];
  const int ch = threadIdx.x & (FIN_CH - 1);
  sh[0][slice][ch] = s; sh[1][slice][ch] = ss;
  __syncthreads();
  if (slice == 0) {
#pragma unroll 8
    for (int k = 1; k < FIN_SLICES; ++k) { s += sh[0][k][ch]; ss += sh[1][k][ch]; }
  }
}

__global__ void __launch_bounds__(256)
bn_stats_finalize_kernel(const float* __restrict__ partial, int nblk, int64_t M, int C,
                         float* __restrict__ mean, float* __restrict__ invstd,
                         float* __restrict__ running_mean, float* __restrict__ running_var,
                         float momentum, float eps) {
  const int c = blockIdx.x * FIN_CH + (threadIdx.x & (FIN_CH - 1)), slice = threadIdx.x / FIN_CH;
  double s, ss;
  reduce_partials(partial, nblk, C, c, slice, s, ss);
  if (slice != 0 || c >= C) return;
  double mu = s / (double)M;
  double var = ss / (double)M - mu * mu;
  if (var < 0) var = 0;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mu);
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
  }
}

// eval-mode BatchNorm as one multiply-add per element: y = x * scale + shift
__global__ void bn_fold_kernel(const float* __restrict__ rm, const float* __restrict__ rv, const float* __restrict__ gamma,
                               const float* __restrict__ beta, int C, float eps, float* __restrict__ scale, float* __restrict__ shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * (1.0f / sqrtf(rv[c] + eps));
  scale[c] = sc;
  shift[c] = fmaf(-rm[c], sc, beta[c]);
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, int C, float eps,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rm[c];
  invstd[c] = 1.0f / sqrtf(rv[c] + eps);
}

__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(BnApplyArgs a) {
  extern __shared__ float sm[];   // scale[C], mean[C], beta[C] (+ same for residual BN)
  float* sc = sm; float* mu = sm + a.C; float* be = sm + 2 * a.C;
  float* rsc = sm + 3 * a.C; float* rmu = sm + 4 * a.C; float* rbe = sm + 5 * a.C;
  const bool res_bn = a.r && a.rmean;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    sc[c] = a.gamma[c] * a.invstd[c]; mu[c] = a.mean[c]; be[c] = a.beta[c];
    if (res_bn) { rsc[c] = a.rgamma[c] * a.rinvstd[c]; rmu[c] = a.rmean[c]; rbe[c] = a.rbeta[c]; }
  }
  __syncthreads();
  const int q = a.C >> 2;
  const int64_t total = a.M * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) << 2;
    float4 v = __ldg(reinterpret_cast<const float4*>(a.x) + i);
    v.x = fmaf(v.x - mu[c], sc[c], be[c]); v.y = fmaf(v.y - mu[c + 1], sc[c + 1], be[c + 1]);
    v.z = fmaf(v.z - mu[c + 2], sc[c + 2], be[c + 2]); v.w = fmaf(v.w - mu[c + 3], sc[c + 3], be[c + 3]);
    if (a.r) {
      float4 r = __ldg(reinterpret_cast<const float4*>(a.r) + i);
      if (res_bn) {
        r.x = fmaf(r.x - rmu[c], rsc[c], rbe[c]); r.y = fmaf(r.y - rmu[c + 1], rsc[c + 1], rbe[c + 1]);
        r.z = fmaf(r.z - rmu[c + 2], rsc[c + 2], rbe[c + 2]); r.w = fmaf(r.w - rmu[c + 3], rsc[c + 3], rbe[c + 3]);
      }
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    reinterpret_cast<float4*>(a.y)[i] = v;
    if (a.hi) store_split4(a.hi, a.lo, i, v);
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C,
                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * FIN_CH + (threadIdx.x & (FIN_CH - 1)), slice = threadIdx.x / FIN_CH;
  double s, ss;
  reduce_partials(partial, nblk, C, c, slice, s, ss);
  if (slice != 0 || c >= C) return;
  dbeta[c] = (float)s;
  dgamma[c] = (float)ss;
}

// dx = gamma*invstd*(g - dbeta/M - xhat*dgamma/M)  (training)   |   gamma*invstd*g  (eval)
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_apply_kernel(BnBwdArgs a) {
  extern __shared__ float sm[];
  float* k1 = sm; float* mu = sm + a.C; float* is = sm + 2 * a.C; float* mb = sm + 3 * a.C; float* mg = sm + 4 * a.C;
  const float invM = (float)(1.0 / (double)a.M);
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    k1[c] = a.gamma[c] * a.invstd[c]; mu[c] = a.mean[c]; is[c] = a.invstd[c];
    mb[c] = a.training ? a.dbeta[c] * invM : 0.f;
    mg[c] = a.training ? a.dgamma[c] * invM : 0.f;
  }
  __syncthreads();
  const int q = a.C >> 2;
  const int64_t total = a.M * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) << 2;
    float4 g = __ldg(reinterpret_cast<const float4*>(a.dy) + i);
    if (a.relu) {
      if (a.y_hi) {
        mask_from_bf16x4(__ldg(reinterpret_cast<const uint2*>(a.y_hi) + i), g);
      } else {
        float4 o = __ldg(reinterpret_cast<const float4*>(a.y) + i);
        g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
      }
    }
    if (a.g_out) reinterpret_cast<float4*>(a.g_out)[i] = g;
    float4 v = __ldg(reinterpret_cast<const float4*>(a.x) + i);
    float4 d;
    d.x = k1[c] * (g.x - mb[c] - (v.x - mu[c]) * is[c] * mg[c]);
    d.y = k1[c + 1] * (g.y - mb[c + 1] - (v.y - mu[c + 1]) * is[c + 1] * mg[c + 1]);
    d.z = k1[c + 2] * (g.z - mb[c + 2] - (v.z - mu[c + 2]) * is[c + 2] * mg[c + 2]);
    d.w = k1[c + 3] * (g.w - mb[c + 3] - (v.w - mu[c + 3]) * is[c + 3] * mg[c + 3]);
    if (a.dx) reinterpret_cast<float4*>(a.dx)[i] = d;
    if (a.dx_hi) store_split4(a.dx_hi, a.dx_lo, i, d);
  }
}

// ------------------------------------------------------------------------------------------------ stem
// y[n,hp,wp,c] = max over the 3x3/2 pad-1 window of relu(bn(x));  argmax = first maximum in (r,s) scan order
__global__ void __launch_bounds__(256)
stem_bn_relu_pool_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ y, uint8_t* __restrict__ argmax, __nv_bfloat16* __restrict__ y_hi,
                         __nv_bfloat16* __restrict__ y_lo, int N, int Hc, int Wc, int C, int Hp, int Wp) {
  const int q = C >> 2;
  const int64_t total = (int64_t)N * Hp * Wp * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) << 2; int64_t t = i / q;
    int wp = (int)(t % Wp); t /= Wp;
    int hp = (int)(t % Hp); int n = (int)(t / Hp);
    float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
    float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    float4 sc = make_float4(ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w);
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 arg = make_uchar4(0, 0, 0, 0);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int h = 2 * hp - 1 + r;
      if (h < 0 || h >= Hc) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int w = 2 * wp - 1 + s;
        if (w < 0 || w >= Wc) continue;
        float4 v = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)n * Hc + h) * Wc + w) * C + c));
        v.x = fmaxf(fmaf(v.x - mu.x, sc.x, be.x), 0.f); v.y = fmaxf(fmaf(v.y - mu.y, sc.y, be.y), 0.f);
        v.z = fmaxf(fmaf(v.z - mu.z, sc.z, be.z), 0.f); v.w = fmaxf(fmaf(v.w - mu.w, sc.w, be.w), 0.f);
        unsigned char k = (unsigned char)(r * 3 + s);
        if (v.x > best.x) { best.x = v.x; arg.x = k; }
        if (v.y > best.y) { best.y = v.y; arg.y = k; }
        if (v.z > best.z) { best.z = v.z; arg.z = k; }
        if (v.w > best.w) { best.w = v.w; arg.w = k; }
      }
    }
    reinterpret_cast<float4*>(y)[i] = best;
    reinterpret_cast<uchar4*>(argmax)[i] = arg;
    if (y_hi) store_split4(y_hi, y_lo, i, best);
  }
}

// g[n,h,w,c] = (bn(x) > 0) * sum over pooling windows whose argmax is (h,w) of dy_pool
__global__ void __launch_bounds__(256)
stem_pool_relu_bwd_kernel(const float* __restrict__ dyp, const uint8_t* __restrict__ argmax, const float* __restrict__ x,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ g,
                          int N, int Hc, int Wc, int C, int Hp, int Wp) {
  const int q = C >> 2;
  const int64_t total = (int64_t)N * Hc * Wc * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % q) << 2; int64_t t = i / q;
    int w = (int)(t % Wc); t /= Wc;
    int h = (int)(t % Hc); int n = (int)(t / Hc);
    float4 acc = make_float4(0, 0, 0, 0);
    int hp0 = h >> 1, hp1 = (h + 1) >> 1, wp0 = w >> 1, wp1 = (w + 1) >> 1;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      int hp = a ? hp1 : hp0;
      if ((a && hp1 == hp0) || hp >= Hp) continue;
      int r = h - (2 * hp - 1);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        int wp = b ? wp1 : wp0;
        if ((b && wp1 == wp0) || wp >= Wp) continue;
        int s = w - (2 * wp - 1);
        unsigned char k = (unsigned char)(r * 3 + s);
        int64_t pi = (((int64_t)n * Hp + hp) * Wp + wp) * q + (c >> 2);
        uchar4 am = reinterpret_cast<const uchar4*>(argmax)[pi];
        float4 d = __ldg(reinterpret_cast<const float4*>(dyp) + pi);
        if (am.x == k) acc.x += d.x;
        if (am.y == k) acc.y += d.y;
        if (am.z == k) acc.z += d.z;
        if (am.w == k) acc.w += d.w;
      }
    }
    float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
    float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    if (!(fmaf(v.x - mu.x, ga.x * is.x, be.x) > 0.f)) acc.x = 0.f;
    if (!(fmaf(v.y - mu.y, ga.y * is.y, be.y) > 0.f)) acc.y = 0.f;
    if (!(fmaf(v.z - mu.z, ga.z * is.z, be.z) > 0.f)) acc.z = 0.f;
    if (!(fmaf(v.w - mu.w, ga.w * is.w, be.w) > 0.f)) acc.w = 0.f;
    reinterpret_cast<float4*>(g)[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ launchers
static int check_c(int C) {
  DDN_CHECK_ARG(C >= 4 && C % 4 == 0 && (C / 4) <= BN_THREADS && BN_THREADS % (C / 4) == 0,
                "BatchNorm kernels need C in {4..1024} with 256 %% (C/4) == 0 (got %d)", C);
  return 0;
}
static int ew_blocks(int64_t total) { return (int)std::min<int64_t>(ceil_div(total, BN_THREADS), (int64_t)num_sms() * 8); }

int launch_bn_stats(const float* x, int64_t M, int C, float* partial, float* mean, float* invstd,
                    float* running_mean, float* running_var, float momentum, float eps, cudaStream_t st) {
  DDN_TRY(check_c(C));
  int nblk = bn_partial_blocks(M, C);
  DDN_LAUNCH(bn_colsum_kernel<0>, nblk, BN_THREADS, 0, st, x, nullptr, nullptr, nullptr, nullptr, nullptr, M, C,
             bn_rows_per_block(M, C), 0, partial);
  DDN_LAUNCH(bn_stats_finalize_kernel, (int)ceil_div(C, FIN_CH), 256, 0, st, partial, nblk, M, C, mean, invstd,
             running_mean, running_var, momentum, eps);
  return 0;
}

int launch_bn_stats_finalize(const float* partial, int nblk, int64_t M, int C, float* mean, float* invstd,
                             float* running_mean, float* running_var, float momentum, float eps, cudaStream_t st) {
  DDN_LAUNCH(bn_stats_finalize_kernel, (int)ceil_div(C, FIN_CH), 256, 0, st, partial, nblk, M, C, mean, invstd,
             running_mean, running_var, momentum, eps);
  return 0;
}

int launch_bn_eval_stats(const float* rm, const float* rv, int C, float eps, float* mean, float* invstd, cudaStream_t st) {
  DDN_LAUNCH(bn_eval_stats_kernel, (int)ceil_div(C, 128), 128, 0, st, rm, rv, C, eps, mean, invstd);
  return 0;
}

int launch_bn_fold(const float* rm, const float* rv, const float* gamma, const float* beta, int C, float eps,
                   float* scale, float* shift, cudaStream_t st) {
  DDN_LAUNCH(bn_fold_kernel, (int)ceil_div(C, 128), 128, 0, st, rm, rv, gamma, beta, C, eps, scale, shift);
  return 0;
}

int launch_bn_apply(const BnApplyArgs& a, cudaStream_t st) {
  DDN_TRY(check_c(a.C));
  DDN_LAUNCH(bn_apply_kernel, ew_blocks(a.M * (a.C / 4)), BN_THREADS, 6 * a.C * sizeof(float), st, a);
  return 0;
}

int launch_bn_backward(const BnBwdArgs& a, cudaStream_t st) {
  DDN_TRY(check_c(a.C));
  int nblk = bn_partial_blocks(a.M, a.C);
  DDN_LAUNCH(bn_colsum_kernel<1>, nblk, BN_THREADS, 0, st, a.x, a.dy, a.y, a.y_hi, a.mean, a.invstd, a.M, a.C,
             bn_rows_per_block(a.M, a.C), a.relu, a.partial);
  DDN_LAUNCH(bn_bwd_finalize_kernel, (int)ceil_div(a.C, FIN_CH), 256, 0, st, a.partial, nblk, a.C, a.dgamma, a.dbeta);
  DDN_LAUNCH(bn_bwd_apply_kernel, ew_blocks(a.M * (a.C / 4)), BN_THREADS, 5 * a.C * sizeof(float), st, a);
  return 0;
}

int launch_stem_bn_relu_pool(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                             float* y, uint8_t* argmax, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo,
                             int N, int Hc, int Wc, int C, cudaStream_t st) {
  int Hp = (Hc - 1) / 2 + 1, Wp = (Wc - 1) / 2 + 1;
  int64_t total = (int64_t)N * Hp * Wp * (C / 4);
  DDN_LAUNCH(stem_bn_relu_pool_kernel, ew_blocks(total), 256, 0, st, x, mean, invstd, gamma, beta, y, argmax, y_hi, y_lo, N, Hc, Wc, C, Hp, Wp);
  return 0;
}

int launch_stem_pool_relu_backward(const float* dy_pool, const uint8_t* argmax, const float* x, const float* mean,
                                   const float* invstd, const float* gamma, const float* beta, float* g,
                                   int N, int Hc, int Wc, int C, cudaStream_t st) {
  int Hp = (Hc - 1) / 2 + 1, Wp = (Wc - 1) / 2 + 1;
  int64_t total = (int64_t)N * Hc * Wc * (C / 4);
  DDN_LAUNCH(stem_pool_relu_bwd_kernel, ew_blocks(total), 256, 0, st, dy_pool, argmax, x, mean, invstd, gamma, beta, g,
             N, Hc, Wc, C, Hp, Wp);
  return 0;
}

}  // namespace ddn
