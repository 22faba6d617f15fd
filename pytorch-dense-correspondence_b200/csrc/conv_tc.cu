// tcgen05 implicit-GEMM convolution for sm_100a: the dilated 3x3 / 1x1 stride-1 convolutions of Resnet34_8s
// (layers 1-4: 99% of the FLOPs), forward and data-gradient, on the 5th-generation tensor cores.
//
//   D[128 pixels x BLOCK_N channels] (fp32, TMEM) += A[128 pixels x 64 ch] (smem) * B[BLOCK_N x 64 ch]^T (smem)
//
// * Activations are NHWC bf16 planes; one CTA owns an 8x16-pixel output tile of one image.  For filter tap
//   (r,s) and 64-channel chunk c the A tile is ONE 4-D TMA box load at (c, w0+(s-1)*dil, h0+(r-1)*dil, n): TMA's
//   out-of-bounds zero fill *is* the convolution padding, so there is no im2col buffer and no halo logic.
// * Weights are [Cout][tap*Cin + ci] bf16 (K-major); a [BLOCK_N x 64] box per k-block.
// * Both operands land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes directly.
// * Precision: DDN_PRECISION_BF16X3 keeps fp32-equivalent results by splitting every operand x = hi + lo
//   (both bf16) and issuing hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator (3 MMAs per k-step);
//   DDN_PRECISION_BF16 issues hi*hi only.
// * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one thread),
//   warps 2-5 = epilogue (tcgen05.ld -> registers -> fp32 NHWC global, optional fused addend).
//   smem ring of kStages {A_hi,A_lo,B_hi,B_lo} slots with full/empty mbarriers; tcgen05.commit frees slots.
//
// Reference op replaced: nn.Conv2d via conv3x3 (PSD/vision/torchvision/models/resnet.py:20-37,45,48) and the
// stride-1 1x1 downsample convs (resnet.py:210-214), plus their autograd data gradient.
#include <cuda.h>

#include "conv.cuh"
#include "conv_tc.cuh"

namespace ddn {

constexpr int TC_TH = 8, TC_TW = 16;          // output tile: 8 rows x 16 cols = 128 pixels = UMMA M
constexpr int TC_BLOCK_K = 64;                // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = 128 * TC_BLOCK_K * 2;   // 16 KB

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, both K-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ the kernel
struct TcConvParams {
  float* out;            // [N,H,W,Cout] fp32
  const float* addend;   // optional, same shape
  int N, H, W, Cin, Cout;
  int taps_w;            // 1 or 3 (k x k filter)
  int dil;
  int tiles_h, tiles_w;
};

template <int BLOCK_N, int NPROD>   // NPROD = 1 (bf16) or 3 (bf16x3)
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const TcConvParams p) {
  constexpr int NSPLIT = NPROD == 3 ? 2 : 1;                  // operand planes per matrix
  constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
  constexpr int STAGE_BYTES = NSPLIT * (TC_A_BYTES + B_BYTES);
  constexpr int STAGES = (200 * 1024) / STAGE_BYTES >= 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  constexpr uint32_t IDESC = make_idesc_bf16(128, BLOCK_N);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int t = blockIdx.x;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h; const int n = t / p.tiles_h;
  const int h0 = th * TC_TH, w0 = tw * TC_TW;
  const int co0 = blockIdx.y * BLOCK_N;
  const int cin_chunks = p.Cin / TC_BLOCK_K;
  const int num_kb = p.taps_w * p.taps_w * cin_chunks;
  const int half = p.taps_w >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_b_hi);
    if (NSPLIT == 2) { tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_b_lo); }
  }
  if (warp == 1) {   // TMEM allocation (whole warp), BLOCK_N fp32 columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(BLOCK_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        const int tap = kb / cin_chunks, cc = kb - tap * cin_chunks;
        const int r = tap / p.taps_w, sx = tap - r * p.taps_w;
        const int hh = h0 + (r - half) * p.dil, ww = w0 + (sx - half) * p.dil;
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        const uint32_t bar = smem_u32(&full_bar[s]);
        mbar_expect_tx(bar, STAGE_BYTES);
        tma_load_4d(smem_u32(st), &tm_a_hi, bar, cc * TC_BLOCK_K, ww, hh, n);
        tma_load_2d(smem_u32(st + NSPLIT * TC_A_BYTES), &tm_b_hi, bar, kb * TC_BLOCK_K, co0);
        if (NSPLIT == 2) {
          tma_load_4d(smem_u32(st + TC_A_BYTES), &tm_a_lo, bar, cc * TC_BLOCK_K, ww, hh, n);
          tma_load_2d(smem_u32(st + 2 * TC_A_BYTES + B_BYTES), &tm_b_lo, bar, kb * TC_BLOCK_K, co0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t a_hi = make_kmajor_sw128_desc(st);
        const uint64_t b_hi = make_kmajor_sw128_desc(st + NSPLIT * TC_A_BYTES);
        const uint64_t a_lo = make_kmajor_sw128_desc(st + TC_A_BYTES);
        const uint64_t b_lo = make_kmajor_sw128_desc(st + 2 * TC_A_BYTES + B_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);       // 16 bf16 = 32 bytes along K inside the swizzle row
          if (NPROD == 3) {
            umma_bf16(tmem_base, a_hi + adv, b_lo + adv, IDESC, (kb | k) != 0);
            umma_bf16(tmem_base, a_lo + adv, b_hi + adv, IDESC, 1);
            umma_bf16(tmem_base, a_hi + adv, b_hi + adv, IDESC, 1);
          } else {
            umma_bf16(tmem_base, a_hi + adv, b_hi + adv, IDESC, (kb | k) != 0);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));       // frees the smem slot once these MMAs have read it
      }
      umma_commit(smem_u32(&tmem_full_bar));        // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global (fp32 NHWC), 4 warps x 32 lanes = 128 rows =====
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int h = h0 + row / TC_TW, w = w0 + row % TC_TW;
    const bool ok = h < p.H && w < p.W;
    mbar_wait(smem_u32(&tmem_full_bar), 0);
    tc_fence_after();
    const size_t pix = ((size_t)n * p.H + (ok ? h : 0)) * p.W + (ok ? w : 0);
    float* o = p.out + pix * p.Cout + co0;
    const float* ad = p.addend ? p.addend + pix * p.Cout + co0 : nullptr;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      if (ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 f = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          if (ad) {
            float4 a = __ldg(reinterpret_cast<const float4*>(ad + c * 32 + j));
            f.x += a.x; f.y += a.y; f.z += a.z; f.w += a.w;
          }
          *reinterpret_cast<float4*>(o + c * 32 + j) = f;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BLOCK_N));
  }
}

// ------------------------------------------------------------------------------------------------ operand preparation
// x fp32 -> hi = bf16(x), lo = bf16(x - hi)      (n multiple of 4)
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                  int64_t n4, int want_lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
    uint2 ho; ho.x = *reinterpret_cast<uint32_t*>(&a); ho.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(hi)[i] = ho;
    if (want_lo) {
      __nv_bfloat162 c = __halves2bfloat162(__float2bfloat16_rn(v.x - __bfloat162float(h0)), __float2bfloat16_rn(v.y - __bfloat162float(h1)));
      __nv_bfloat162 d = __halves2bfloat162(__float2bfloat16_rn(v.z - __bfloat162float(h2)), __float2bfloat16_rn(v.w - __bfloat162float(h3)));
      uint2 lo2; lo2.x = *reinterpret_cast<uint32_t*>(&c); lo2.y = *reinterpret_cast<uint32_t*>(&d);
      reinterpret_cast<uint2*>(lo)[i] = lo2;
    }
  }
}

// w [Cout][Cin][k][k] fp32 ->  fwd:   B[co][(r*k+s)*Cin + ci]
//                             dgrad: B[ci][(r'*k+s')*Cout + co]  with (r,s) = (k-1-r', k-1-s')
__global__ void pack_weights_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                       int Cout, int Cin, int k, int dgrad, int want_lo) {
  const int64_t total = (int64_t)Cout * Cin * k * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int co, ci, r, s;
    if (!dgrad) {
      ci = (int)(i % Cin); int64_t q = i / Cin;
      s = (int)(q % k); q /= k;
      r = (int)(q % k); co = (int)(q / k);
    } else {
      co = (int)(i % Cout); int64_t q = i / Cout;
      s = k - 1 - (int)(q % k); q /= k;
      r = k - 1 - (int)(q % k); ci = (int)(q / k);
    }
    float v = w[(((int64_t)co * Cin + ci) * k + r) * k + s];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (want_lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return DDN_EUNSUPPORTED; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {TC_BLOCK_K, TC_TW, TC_TH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return DDN_EINVAL; }
  return 0;
}
static int make_weight_map(CUtensorMap* m, const void* base, int rows, int K, int block_n) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return DDN_EUNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {TC_BLOCK_K, (cuuint32_t)block_n};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return DDN_EINVAL; }
  return 0;
}

bool tc_available() { return true; }

bool tc_conv_supported(int Cin, int Cout, int k, int stride, int pad, int dil, int H, int W) {
  if (stride != 1 || Cin % 64 || Cout % 64) return false;
  if (k == 3) return pad == dil;
  if (k == 1) return pad == 0;
  return false;
}

// staging layout inside the tc workspace: act_hi | act_lo | w_hi | w_lo
static const size_t kMaxWeightElems = (size_t)9 * 512 * 512;
size_t tc_workspace_bytes(size_t max_act_elems) {
  // max_act_elems: the largest N*H*W*C tensor any supported conv reads (forward input or dY)
  return 2 * align_up(max_act_elems * 2, 1024) + 2 * align_up(kMaxWeightElems * 2, 1024) + 2048;
}

template <int BLOCK_N, int NPROD>
static int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                     const TcConvParams& p, cudaStream_t st) {
  constexpr int NSPLIT = NPROD == 3 ? 2 : 1;
  constexpr int STAGE_BYTES = NSPLIT * (TC_A_BYTES + BLOCK_N * TC_BLOCK_K * 2);
  constexpr int STAGES = (200 * 1024) / STAGE_BYTES >= 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    DDN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((unsigned)(p.N * p.tiles_h * p.tiles_w), (unsigned)(p.Cout / BLOCK_N));
  DDN_LAUNCH((conv_tc_kernel<BLOCK_N, NPROD>), grid, TC_THREADS, smem, st, a_hi, a_lo, b_hi, b_lo, p);
  return 0;
}

// out[N,H,W,Cout] = conv(in[N,H,W,Cin]; packed weights) (+ addend); weights come in the reference layout and are
// packed for the forward (dgrad = 0) or the data-gradient (dgrad = 1; then Cin/Cout are those of the ORIGINAL conv and
// `in` is dY [N,H,W,Cout], `out` is dX [N,H,W,Cin]).
static int tc_run(const float* in, const float* w_oihw, float* out, const float* addend, int N, int H, int W,
                  int Cin, int Cout, int k, int dil, int dgrad, int precision, void* ws, size_t ws_bytes, cudaStream_t st) {
  const double fl = 2.0 * N * H * W * (double)Cout * k * k * Cin;
  const int gin = dgrad ? Cout : Cin, gout = dgrad ? Cin : Cout;     // channels of the GEMM's input / output tensors
  const size_t act = (size_t)N * H * W * gin;
  const size_t wel = (size_t)Cout * Cin * k * k;
  const size_t act_b = align_up(act * 2, 1024), w_b = align_up(kMaxWeightElems * 2, 1024);
  DDN_CHECK_ARG(ws != nullptr, "tc workspace missing");
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<uintptr_t>(ws), 1024));
  if ((size_t)(base - (char*)ws) + 2 * act_b + 2 * w_b > ws_bytes) {
    set_error("tcgen05 conv workspace too small (%zu needed)", 2 * act_b + 2 * w_b + 1024);
    return DDN_EWORKSPACE;
  }
  DDN_CHECK_ARG(wel <= kMaxWeightElems, "weight tensor larger than the staging buffer");
  __nv_bfloat16* a_hi = (__nv_bfloat16*)base;
  __nv_bfloat16* a_lo = (__nv_bfloat16*)(base + act_b);
  __nv_bfloat16* b_hi = (__nv_bfloat16*)(base + 2 * act_b);
  __nv_bfloat16* b_lo = (__nv_bfloat16*)(base + 2 * act_b + w_b);
  const int want_lo = precision == DDN_PRECISION_BF16X3;
  {
    int64_t n4 = (int64_t)(act / 4);
    int blocks = (int)std::min<int64_t>(ceil_div(n4, 256), (int64_t)num_sms() * 8);
    DDN_LAUNCH(split_bf16_kernel, blocks, 256, 0, st, in, a_hi, a_lo, n4, want_lo);
    int wblocks = (int)std::min<int64_t>(ceil_div((int64_t)wel, 256), 4096);
    DDN_LAUNCH(pack_weights_tc_kernel, wblocks, 256, 0, st, w_oihw, b_hi, b_lo, Cout, Cin, k, dgrad, want_lo);
  }
  const int block_n = gout % 128 == 0 ? 128 : 64;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  DDN_TRY(make_act_map(&ma_hi, a_hi, N, H, W, gin));
  DDN_TRY(make_act_map(&ma_lo, want_lo ? a_lo : a_hi, N, H, W, gin));
  DDN_TRY(make_weight_map(&mb_hi, b_hi, gout, k * k * gin, block_n));
  DDN_TRY(make_weight_map(&mb_lo, want_lo ? b_lo : b_hi, gout, k * k * gin, block_n));
  TcConvParams p;
  p.out = out; p.addend = addend; p.N = N; p.H = H; p.W = W; p.Cin = gin; p.Cout = gout; p.taps_w = k; p.dil = dil;
  p.tiles_h = (int)ceil_div(H, TC_TH); p.tiles_w = (int)ceil_div(W, TC_TW);
  ProfScope ps(dgrad ? PROF_CONV_DGRAD_TC : PROF_CONV_FWD_TC, fl, st);   // times the MMA kernel only
  if (want_lo) {
    if (block_n == 128) return launch_tc<128, 3>(ma_hi, ma_lo, mb_hi, mb_lo, p, st);
    return launch_tc<64, 3>(ma_hi, ma_lo, mb_hi, mb_lo, p, st);
  }
  if (block_n == 128) return launch_tc<128, 1>(ma_hi, ma_lo, mb_hi, mb_lo, p, st);
  return launch_tc<64, 1>(ma_hi, ma_lo, mb_hi, mb_lo, p, st);
}

int tc_conv_forward(const float* x, const float* w, float* y, int N, int H, int W, int Cin, int Cout, int k, int pad, int dil,
                    int precision, void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)pad;
  return tc_run(x, w, y, nullptr, N, H, W, Cin, Cout, k, dil, 0, precision, ws, ws_bytes, st);
}

// data gradient on the tensor cores; weight gradient on the fp32 SIMT kernel for now (dwp_scratch: K x Cout floats)
int tc_conv_backward(const float* x, const float* w, const float* dy, float* dx, const float* dx_addend, float* dw,
                     int N, int H, int W, int Cin, int Cout, int k, int pad, int dil, int precision,
                     void* ws, size_t ws_bytes, float* dwp_scratch, cudaStream_t st) {
  const double fl = 2.0 * N * H * W * (double)Cout * k * k * Cin;
  ConvGeom g;
  DDN_TRY(conv_geom_init(&g, N, H, W, Cin, H, W, Cout, k, k, 1, 1, pad, dil));
  DDN_TRY(launch_fill_zero(dwp_scratch, sizeof(float) * (size_t)k * k * Cin * Cout, st));
  {
    ProfScope ps(PROF_CONV_WGRAD_SIMT, fl, st);
    DDN_TRY(launch_conv_wgrad_f32(x, dy, dwp_scratch, g, st));
  }
  DDN_TRY(launch_unpack_wgrad(dwp_scratch, dw, Cout, Cin, Cin, k, k, st));
  if (dx) DDN_TRY(tc_run(dy, w, dx, dx_addend, N, H, W, Cin, Cout, k, dil, 1, precision, ws, ws_bytes, st));
  return 0;
}

}  // namespace ddn
