// Shared pieces of the pixelwise-contrastive-loss kernels (loss.cu: gathers from the full-resolution descriptor image;
// loss_lowres.cu: gathers through the bilinear upsample from the low-resolution map).
#pragma once
#include "common.cuh"

namespace ddn {

struct DevTerm {
  const int64_t* ia;
  const int64_t* ib;
  const int64_t* gt;
  const int64_t* len;      // per-pair true counts of a ragged batch (rows padded with -1), or null
  const int64_t* len_gt;
  int64_t n, n_gt;
  int64_t k;           // n / n_gt (py2 integer division, pcl.py:321) when the batch is not ragged
  int kind, flags;
  float margin, m_pixel, inv_m_pixel;
  int block_begin;     // first blockIdx.x of this term
};
struct DevTerms {
  int n_terms;
  int total_blocks;
  DevTerm t[DDN_MAX_TERMS];
};

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_ITEMS = 4;   // index pairs per thread
constexpr int LOSS_MAXD = 32;

__device__ __forceinline__ int find_term(const DevTerms& T, int bx) {
  int t = 0;
#pragma unroll
  for (int i = 1; i < DDN_MAX_TERMS; ++i)
    if (i < T.n_terms && bx >= T.t[i].block_begin) t = i;
  return t;
}

__device__ __forceinline__ float pixel_weight(const DevTerm& tm, int64_t b, int64_t j, int64_t nb, int W) {
  // l2_pixel_loss: 1/M_pixel * clamp(||uv_gt - uv||_2, max=M_pixel), uv = (n % W, n // W)   (pcl.py:321-331,349-351)
  int64_t k = tm.k;
  if (tm.len || tm.len_gt) {      // ragged: this pair's own non-matches-per-match
    const int64_t ng = tm.len_gt ? tm.len_gt[b] : tm.n_gt;
    k = ng > 0 ? (tm.len ? tm.len[b] : tm.n) / ng : 1;
    if (k < 1) k = 1;
  }
  int64_t gi = j / k;
  if (gi >= tm.n_gt) gi = tm.n_gt - 1;
  int64_t g = tm.gt[b * tm.n_gt + gi];
  float du = (float)(g % W - nb % W);
  float dv = (float)(g / W - nb / W);
  float nrm = sqrtf(du * du + dv * dv);
  return tm.inv_m_pixel * fminf(nrm, tm.m_pixel);
}

static inline int build_terms(const ddn_loss_term* th, int n_terms, DevTerms* T) {
  DDN_CHECK_ARG(th && n_terms >= 1 && n_terms <= DDN_MAX_TERMS, "n_terms must be in [1,%d]", DDN_MAX_TERMS);
  T->n_terms = n_terms;
  int blk = 0;
  for (int i = 0; i < n_terms; ++i) {
    const ddn_loss_term& h = th[i];
    DDN_CHECK_ARG(h.n >= 0 && (h.n == 0 || (h.idx_a && h.idx_b)), "term %d: null indices", i);
    DDN_CHECK_ARG(h.kind >= DDN_TERM_MATCH && h.kind <= DDN_TERM_HINGE_INV, "term %d: bad kind", i);
    DevTerm& d = T->t[i];
    d.ia = h.idx_a; d.ib = h.idx_b; d.gt = h.gt_b; d.n = h.n; d.n_gt = h.n_gt; d.k = 1; d.len = h.len; d.len_gt = h.len_gt;
    d.kind = h.kind; d.flags = h.flags; d.margin = h.margin; d.m_pixel = h.m_pixel;
    d.inv_m_pixel = h.m_pixel != 0.f ? (float)(1.0 / (double)h.m_pixel) : 0.f;
    if (h.flags & DDN_TERM_PIXEL_WEIGHT) {
      DDN_CHECK_ARG(h.kind != DDN_TERM_MATCH && h.gt_b && h.n_gt > 0 && ((h.len || h.len_gt) || (h.n % h.n_gt == 0 && h.n >= h.n_gt)),
                    "term %d: pixel weight needs gt_b with n a positive multiple of n_gt", i);
      d.k = h.n >= h.n_gt ? h.n / h.n_gt : 1;
    }
    d.block_begin = blk;
    blk += (int)ceil_div(h.n, LOSS_THREADS * LOSS_ITEMS);
  }
  T->total_blocks = blk;
  return 0;
}

static inline int check_common(const float* pa, const float* pb, int B, int64_t P, int D, int W) {
  DDN_CHECK_ARG(pa && pb, "null descriptor image");
  DDN_CHECK_ARG(B >= 1 && B <= 65535 && P >= 1 && D >= 1 && D <= LOSS_MAXD && W >= 1, "bad B/P/D/W (D<=%d)", LOSS_MAXD);
  return 0;
}


}  // namespace ddn
