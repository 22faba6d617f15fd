// Batched best-match search in descriptor space.
// Replaces the host-side numpy scan of DenseCorrespondenceNetwork.find_best_match
// (dense_correspondence/network/dense_correspondence_network.py:488-525: norm_diffs = sqrt(sum((res_b - d)^2, axis=2));
// argmin; called ~100x per image pair by evaluation.py:993,1047 after a D2H copy of both descriptor images) with one
// launch for Q query descriptors against a descriptor image that stays on the device.
// HBM/L2-bound: every query streams the [P, D] image once (1.23*D MB); queries of one launch share it through L2.
#include "common.cuh"

namespace ddn {

constexpr int MATCH_THREADS = 256;
constexpr int MATCH_MAXD = 32;

// res_b element (p, c) at p*sp + c*sc.  best[q] = packed (float bits of squared distance << 32 | pixel index), pre-set to ~0.
// Squared distances are non-negative, so their IEEE bit patterns order like the values; equal distances resolve to the
// smallest pixel index, which is numpy.argmin's first-minimum rule.
__global__ void __launch_bounds__(MATCH_THREADS)
best_match_kernel(const float* __restrict__ res_b, int64_t sp, int64_t sc, int64_t P, int D,
                  const float* __restrict__ queries, int Q, int pixels_per_block,
                  unsigned long long* __restrict__ best, float* __restrict__ norm_diffs,
                  const float* __restrict__ mask, unsigned long long* __restrict__ best_masked) {
  pdl_prologue();
  const int q = blockIdx.y;
  __shared__ float qd[MATCH_MAXD];
  if (threadIdx.x < D) qd[threadIdx.x] = queries[(int64_t)q * D + threadIdx.x];
  __syncthreads();
  const int64_t p0 = (int64_t)blockIdx.x * pixels_per_block;
  const int64_t p1 = min(P, p0 + pixels_per_block);
  unsigned long long loc = ~0ull, locm = ~0ull;
  for (int64_t p = p0 + threadIdx.x; p < p1; p += MATCH_THREADS) {
    float s = 0.f;
    for (int c = 0; c < D; ++c) {
      float d = __ldg(res_b + p * sp + c * sc) - qd[c];
      s = fmaf(d, d, s);
    }
    if (norm_diffs) norm_diffs[(int64_t)q * P + p] = sqrtf(s);
    unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(uint32_t)p;
    loc = key < loc ? key : loc;
    if (mask) {      // evaluation.py:1052-1059: argmin(norm_diffs + (1 - mask_b) * 1e6), in fp32 like numpy on float32 arrays
      const float dm = sqrtf(s) + (1.0f - __ldg(mask + p)) * 1e6f;
      unsigned long long km = ((unsigned long long)__float_as_uint(dm) << 32) | (unsigned long long)(uint32_t)p;
      locm = km < locm ? km : locm;
    }
  }
  if (mask) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, locm, o);
      locm = other < locm ? other : locm;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(best_masked + q, locm);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, loc, o);
    loc = other < loc ? other : loc;
  }
  __shared__ unsigned long long sm[MATCH_THREADS / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = loc;
  __syncthreads();
  if (threadIdx.x < 32) {
    loc = threadIdx.x < MATCH_THREADS / 32 ? sm[threadIdx.x] : ~0ull;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, loc, o);
      loc = other < loc ? other : loc;
    }
    if (threadIdx.x == 0) atomicMin(best + q, loc);
  }
}

__global__ void best_match_finish_kernel(const unsigned long long* __restrict__ best, int Q, int W,
                                         int64_t* __restrict__ uv, float* __restrict__ diff, int is_distance) {
  pdl_prologue();
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  unsigned long long k = best[q];
  uint32_t p = (uint32_t)(k & 0xffffffffull);
  uv[2 * q + 0] = p % W;            // u = column
  uv[2 * q + 1] = p / W;            // v = row
  const float v = __uint_as_float((uint32_t)(k >> 32));
  diff[q] = is_distance ? v : sqrtf(v);
}

}  // namespace ddn

using namespace ddn;

extern "C" int ddn_find_best_match(const float* res_b, int64_t stride_p, int64_t stride_c, int H, int W, int D,
                                   const float* queries, int Q, int64_t* best_uv, float* best_diff, float* norm_diffs,
                                   const float* mask_b, int64_t* best_uv_masked, float* best_diff_masked,
                                   void* scratch /* 2 x Q x 8 bytes */, void* stream) {
  DDN_CHECK_ARG(res_b && queries && best_uv && best_diff && scratch, "null argument");
  DDN_CHECK_ARG(!mask_b || (best_uv_masked && best_diff_masked), "a mask needs the masked outputs");
  DDN_CHECK_ARG(H > 0 && W > 0 && D >= 1 && D <= MATCH_MAXD && Q >= 1 && Q <= 65535 && (int64_t)H * W < (1ll << 32), "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t P = (int64_t)H * W;
  DDN_CUDA(cudaMemsetAsync(scratch, 0xff, sizeof(unsigned long long) * Q * 2, st));
  unsigned long long* best = reinterpret_cast<unsigned long long*>(scratch);
  int blocks_x = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(P, 2048), ceil_div((int64_t)num_sms() * 4, Q)));
  int ppb = (int)(ceil_div(ceil_div(P, blocks_x), MATCH_THREADS) * MATCH_THREADS);
  blocks_x = (int)ceil_div(P, ppb);
  dim3 grid(blocks_x, Q);
  DDN_LAUNCH(best_match_kernel, grid, MATCH_THREADS, 0, st, res_b, stride_p, stride_c, P, D, queries, Q, ppb,
             best, norm_diffs, mask_b, best + Q);
  DDN_LAUNCH(best_match_finish_kernel, (Q + 127) / 128, 128, 0, st, best, Q, W, best_uv, best_diff, 0);
  if (mask_b)
    DDN_LAUNCH(best_match_finish_kernel, (Q + 127) / 128, 128, 0, st, best + Q, Q, W, best_uv_masked, best_diff_masked, 1);
  return 0;
}
