// Non-match sampling on the device.
// Replaces the CPU pipeline of dense_correspondence/correspondence_tools/correspondence_finder.py:276-405
// (create_non_correspondences: nonzero(mask) -> rand*len -> floor -> index_select -> (u, v)) followed by
// dense_correspondence/dataset/spartan_dataset_masked.py:841-858 (create_non_matches: every match repeated k times on the
// A side) and :1255-1264 (flatten_uv_tensor: n = u + W*v), i.e. it emits the two int64 index tensors the loss consumes
// directly, so up to 2 x 1.5 M x 8 bytes per pair never cross PCIe.
// The reference's "perturb non-matches that are too close to a match" step is a no-op upstream (`ones = torch.zeros_like`
// at correspondence_finder.py:354 makes need_to_be_perturbed identically zero); it is reproduced as that no-op.
// The uniform random numbers are an INPUT (torch.rand on the device), which makes the op bit-reproducible against the
// restated reference given the same numbers.
#include "common.cuh"

namespace ddn {

constexpr int SAMP_THREADS = 256;
constexpr int SAMP_PER_BLOCK = 1024;     // pixels per block in the compaction passes

__global__ void __launch_bounds__(SAMP_THREADS)
mask_count_kernel(const float* __restrict__ mask, int64_t P, int* __restrict__ block_counts) {
  const int64_t base = (int64_t)blockIdx.x * SAMP_PER_BLOCK;
  int c = 0;
#pragma unroll
  for (int i = 0; i < SAMP_PER_BLOCK / SAMP_THREADS; ++i) {
    int64_t p = base + i * SAMP_THREADS + threadIdx.x;
    c += (p < P && mask[p] != 0.f) ? 1 : 0;
  }
  c = warp_sum(c);
  __shared__ int s[SAMP_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < SAMP_THREADS / 32; ++i) t += s[i];
    block_counts[blockIdx.x] = t;
  }
}

// exclusive scan of up to 8192 block counts by one block; total -> counts[nblk]
__global__ void __launch_bounds__(1024)
mask_scan_kernel(int* __restrict__ counts, int nblk) {
  __shared__ int s[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nblk ? counts[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      int t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    int incl = s[threadIdx.x];
    if (i < nblk) counts[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[nblk] = carry;
}

// ascending list of the nonzero pixels (== torch.nonzero order)
__global__ void __launch_bounds__(SAMP_THREADS)
mask_compact_kernel(const float* __restrict__ mask, int64_t P, const int* __restrict__ block_offsets, int* __restrict__ nz) {
  const int64_t base = (int64_t)blockIdx.x * SAMP_PER_BLOCK;
  __shared__ int warp_tot[SAMP_PER_BLOCK / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  bool f[SAMP_PER_BLOCK / SAMP_THREADS];
  int rank[SAMP_PER_BLOCK / SAMP_THREADS];
#pragma unroll
  for (int i = 0; i < SAMP_PER_BLOCK / SAMP_THREADS; ++i) {
    int64_t p = base + i * SAMP_THREADS + threadIdx.x;
    f[i] = p < P && mask[p] != 0.f;
    unsigned b = __ballot_sync(0xffffffffu, f[i]);
    rank[i] = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[i * (SAMP_THREADS / 32) + wid] = __popc(b);
  }
  __syncthreads();
  // segment order inside the block: (i, wid) ascending == pixel order
#pragma unroll
  for (int i = 0; i < SAMP_PER_BLOCK / SAMP_THREADS; ++i) {
    if (!f[i]) continue;
    int seg = i * (SAMP_THREADS / 32) + wid, before = 0;
    for (int k = 0; k < seg; ++k) before += warp_tot[k];
    nz[block_offsets[blockIdx.x] + before + rank[i]] = (int)(base + i * SAMP_THREADS + threadIdx.x);
  }
}

__global__ void __launch_bounds__(SAMP_THREADS)
sample_non_matches_kernel(const int* __restrict__ nz, const int* __restrict__ total, const float* __restrict__ rand_u,
                          const float* __restrict__ rand_v, int64_t n, int H, int W, const int64_t* __restrict__ matches_a,
                          int64_t k, int64_t* __restrict__ out_a, int64_t* __restrict__ out_b) {
  const int L = total ? total[0] : 0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t b;
    if (L > 0) {
      int r = (int)floorf(rand_u[j] * (float)L);      // torch.rand(n) * len(mask_b_indices_flat) -> floor -> long
      if (r >= L) r = L - 1;                          // fp32 rounding at rand ~ 1: the reference's index_select would raise here
      b = nz[r];
    } else {                                           // no / empty mask: pytorch_rand_select_pixel (finder.py:64-75)
      int u = (int)floorf(rand_u[j] * (float)W), v = (int)floorf(rand_v[j] * (float)H);
      if (u >= W) u = W - 1;
      if (v >= H) v = H - 1;
      b = (int64_t)u + (int64_t)W * v;
    }
    out_b[j] = b;
    if (out_a) out_a[j] = matches_a[j / k];
  }
}

}  // namespace ddn

using namespace ddn;

extern "C" size_t ddn_sample_non_matches_scratch_bytes(int H, int W) {
  int64_t P = (int64_t)H * W;
  return sizeof(int) * (size_t)(P + ceil_div(P, SAMP_PER_BLOCK) + 8) + 256;
}

extern "C" int ddn_sample_non_matches(const float* mask, int H, int W, const float* rand_u, const float* rand_v, int64_t n,
                                      const int64_t* matches_a, int64_t non_matches_per_match, int64_t* out_a, int64_t* out_b,
                                      void* scratch, size_t scratch_bytes, void* stream) {
  DDN_CHECK_ARG(rand_u && rand_v && out_b && H > 0 && W > 0 && n >= 0 && (int64_t)H * W < (1ll << 31), "bad arguments");
  DDN_CHECK_ARG(!out_a || (matches_a && non_matches_per_match >= 1), "A-side output needs matches_a and non_matches_per_match");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t P = (int64_t)H * W;
  int* nz = nullptr; int* total = nullptr;
  if (mask) {
    DDN_CHECK_ARG(scratch && scratch_bytes >= ddn_sample_non_matches_scratch_bytes(H, W), "scratch too small");
    const int nblk = (int)ceil_div(P, SAMP_PER_BLOCK);
    DDN_CHECK_ARG(nblk <= 1 << 20, "image too large");
    int* counts = reinterpret_cast<int*>(scratch);      // [nblk + 1]
    nz = counts + nblk + 8;
    DDN_LAUNCH(mask_count_kernel, nblk, SAMP_THREADS, 0, st, mask, P, counts);
    DDN_LAUNCH(mask_scan_kernel, 1, 1024, 0, st, counts, nblk);
    DDN_LAUNCH(mask_compact_kernel, nblk, SAMP_THREADS, 0, st, mask, P, counts, nz);
    total = counts + nblk;
  }
  int blocks = (int)std::min<int64_t>(ceil_div(n, SAMP_THREADS), (int64_t)num_sms() * 8);
  DDN_LAUNCH(sample_non_matches_kernel, blocks, SAMP_THREADS, 0, st, nz, total, rand_u, rand_v, n, H, W, matches_a,
             non_matches_per_match, out_a, out_b);
  return 0;
}
