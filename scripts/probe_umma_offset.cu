// Hardware probe (not part of the library): does tcgen05.mma accept 128-byte-swizzled operands whose descriptor start address is
// NOT aligned to the 1024-byte swizzle atom, i.e. a window that starts at an arbitrary 128-byte row of a larger staged tile?
// That is what reusing one halo tile of activations for all 9 taps of a 3x3 convolution needs:
//   K-major A (forward / data gradient): 128 MMA rows = 16 groups of 8 consecutive pixels, groups `pitch` rows apart
//                                        (SBO = pitch * 128 B), first row p0 = any row of the halo tile;
//   MN-major operands (weight gradient): K = 16 consecutive pixel rows starting at any row p0.
// Each variant is run with the descriptor's base-offset field = 0 and = (start >> 7) & 7.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probe_umma_offset scripts/probe_umma_offset.cu && scripts/probe_umma_offset
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D1;\n\tbra W1;\n\tD1:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo, int base_mode) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  if (base_mode) d |= (uint64_t)((addr >> 7) & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__host__ __device__ inline float val_a(int p, int k) { return (float)(((p * 7 + k * 3) % 17) - 8); }
__host__ __device__ inline float val_b(int n, int k) { return (float)(((n * 5 + k) % 13) - 6); }

constexpr int ROWS = 400;      // rows of the staged "halo" tile

// mode 0: K-major.  A rows = pixels (128 B = 64 k), window: row(m) = p0 + (m / 8) * pitch + (m % 8); B = [64 n][64 k], aligned.
// mode 1: MN-major. A = [pixel rows][2 atoms of 64 m], B = [pixel rows][64 n]; K = 64 pixel rows p0 .. p0 + 63 (4 MMAs of 16 rows).
// mode 2: MN-major, the two 64-wide M atoms of A are the SAME staged rows shifted by `shift` rows (LBO = shift * 128 B: two filter
//         taps stacked along M), K groups `pitch` rows apart (SBO = pitch * 128 B); B = [pixel rows][64 n], aligned, SBO 1024.
__global__ void __launch_bounds__(128) probe_kernel(int mode, int p0, int pitch, int base_mode, int shift, float* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                          // mode 0: ROWS x 128 B; mode 1: 2 atoms x ROWS x 128 B
  uint8_t* sb = smem + 2 * ROWS * 128;         // mode 0: 64 x 128 B;   mode 1: ROWS x 128 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  auto put = [](uint8_t* base, int row, int col, float v) {     // address-based 128B swizzle (base is 1024-aligned)
    const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)col & 7u) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(v);
  };
  if (mode == 0) {
    for (int i = tid; i < ROWS * 64; i += 128) put(sa, i / 64, i % 64, val_a(i / 64, i % 64));
    for (int i = tid; i < 64 * 64; i += 128) put(sb, i / 64, i % 64, val_b(i / 64, i % 64));
  } else if (mode == 2) {
    for (int i = tid; i < ROWS * 64; i += 128) put(sa, i / 64, i % 64, val_a(i / 64, i % 64));
    for (int i = tid; i < ROWS * 64; i += 128) put(sb, i / 64, i % 64, val_b(i % 64, i / 64));      // B[k][n] = val_b(n, k)
  } else {
    for (int i = tid; i < ROWS * 128; i += 128) { const int p = i / 128, m = i % 128; put(sa + (m / 64) * ROWS * 128, p, m % 64, val_a(p, m)); }
    for (int i = tid; i < ROWS * 64; i += 128) put(sb, i / 64, i % 64, val_b(i % 64, i / 64));      // B[p][n] = val_b(n, p)
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    if (mode == 0) {
      const uint32_t a0 = smem_u32(sa) + (uint32_t)p0 * 128u, b0 = smem_u32(sb);
      for (int k = 0; k < 4; ++k)
        umma(tmem, desc_sw128(a0 + k * 32, 16, (uint32_t)pitch * 128u, base_mode), desc_sw128(b0 + k * 32, 16, 1024, 0), idesc_bf16(128, 64, false), k != 0);
    } else if (mode == 2) {
      const uint32_t a0 = smem_u32(sa) + (uint32_t)p0 * 128u, b0 = smem_u32(sb);
      for (int k = 0; k < 4; ++k)      // k-step = 16 K rows = 2 groups of 8, `pitch` rows apart in A, 8 rows apart in B
        umma(tmem, desc_sw128(a0 + k * 2 * pitch * 128, (uint32_t)shift * 128u, (uint32_t)pitch * 128u, base_mode),
             desc_sw128(b0 + k * 16 * 128, ROWS * 128, 1024, 0), idesc_bf16(128, 64, true), k != 0);
    } else {
      const uint32_t a0 = smem_u32(sa) + (uint32_t)p0 * 128u, b0 = smem_u32(sb) + (uint32_t)p0 * 128u;
      for (int k = 0; k < 4; ++k)
        umma(tmem, desc_sw128(a0 + k * 16 * 128, ROWS * 128, 1024, base_mode), desc_sw128(b0 + k * 16 * 128, ROWS * 128, 1024, base_mode),
             idesc_bf16(128, 64, true), k != 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  for (int c = 0; c < 2; ++c) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[tid * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 64 * sizeof(float));
  const size_t smem = 3 * ROWS * 128 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> out(128 * 64);
  int bad_total = 0;
  const int p0s[] = {0, 8, 1, 3, 10, 21};
  const int pitches[] = {8, 10, 12, 18, 24};
  const int shifts[] = {1, 8, 10, 0};
  for (int mode = 0; mode < 3; ++mode)
    for (int base_mode = 0; base_mode < (mode == 2 ? 1 : 2); ++base_mode)
      for (int p0 : p0s)
        for (int pitch : pitches)
        for (int shift : shifts) {
          if (mode != 2 && shift != 1) continue;
          if (mode == 1 && pitch != 8) continue;
          if (p0 + 15 * pitch + 8 > ROWS) continue;
          cudaMemset(d_out, 0, out.size() * sizeof(float));
          probe_kernel<<<1, 128, smem>>>(mode, p0, pitch, base_mode, shift, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mode %d base %d p0 %d pitch %d: CUDA error %s\n", mode, base_mode, p0, pitch, cudaGetErrorString(e)); return 1; }
          cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
          int bad = 0; double maxerr = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
              double ref = 0;
              if (mode == 0) { const int row = p0 + (m / 8) * pitch + (m % 8); for (int k = 0; k < 64; ++k) ref += (double)val_a(row, k) * val_b(n, k); }
              else if (mode == 2) { for (int k = 0; k < 64; ++k) ref += (double)val_a(p0 + (k / 8) * pitch + (k % 8) + (m / 64) * shift, m % 64) * val_b(n, k); }
              else { for (int k = 0; k < 64; ++k) ref += (double)val_a(p0 + k, m) * val_b(n, p0 + k); }
              const double err = fabs(ref - out[m * 64 + n]);
              if (err > 1e-3) ++bad;
              if (err > maxerr) maxerr = err;
            }
          printf("%s base_offset=%s p0=%2d pitch=%2d shift=%2d : %s (%d / 8192 wrong, max err %.1f)\n", mode == 0 ? "K-major " : mode == 1 ? "MN-major" : "MN-stack",
                 base_mode ? "auto" : "0   ", p0, pitch, shift, bad ? "MISMATCH" : "ok", bad, maxerr);
          bad_total += bad != 0;
        }
  printf("variants with mismatches: %d\n", bad_total);
  return 0;
}
