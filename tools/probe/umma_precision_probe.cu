// 3xTF32 precision of one tcgen05.mma k-step (M128 N32 K8) through the three operand paths of the cost-volume backward:
//   0: A, B K-major SWIZZLE_NONE tiles in shared memory      1: A from tensor memory (tcgen05.st), B K-major
//   2: A, B MN-major SWIZZLE_128B_BASE32B tiles (the product reduces over the 8 tile ROWS: D[i][o] = sum_r A[r][i] B[r][o])
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return ((uint64_t)layout << 61) | (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

// A [128][8], B [32][8] row-major fp32 in global; out [3][128][32]
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  unsigned char *Ak_hi = base, *Ak_lo = base + 4096, *Bk_hi = base + 8192, *Bk_lo = base + 9216;        // K-major tiles
  unsigned char *Am_hi = base + 16384, *Am_lo = base + 16384 + 4096, *Bm_hi = base + 16384 + 8192, *Bm_lo = base + 16384 + 12288;   // MN-major: 8 rows x 128 B (+ groups)
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 4; i += 128) ((uint32_t*)base)[i] = 0u;
  __syncthreads();
  // K-major: element (row, k) at (k>>2)*128 + (row>>3)*256 + (row&7)*16 + (k&3)*4
  for (int k = 0; k < 8; k++) {
    const float v = A[tid * 8 + k]; const uint32_t h = to_tf32(v), l = to_tf32(v - __uint_as_float(h));
    const uint32_t off = (k >> 2) * 128 + (tid >> 3) * 256 + (tid & 7) * 16 + (k & 3) * 4;
    *(uint32_t*)(Ak_hi + off) = h; *(uint32_t*)(Ak_lo + off) = l;
    if (tid < 32) {
      const float w = B[tid * 8 + k]; const uint32_t bh = to_tf32(w), bl = to_tf32(w - __uint_as_float(bh));
      *(uint32_t*)(Bk_hi + off) = bh; *(uint32_t*)(Bk_lo + off) = bl;
    }
  }
  // MN-major path computes D2[i][o] = sum_{r<8} A[i][r] * B[o][r]  (same numbers: tile row r = k, tile column = mn index)
  // element (row r, col c) at (c>>5)*1024 + r*128 + ((((c&31)>>3) ^ (r&3))<<5) + (c&7)*4      [groups of 32 columns at LBO = 1024]
  for (int k = 0; k < 8; k++) {
    const float v = A[tid * 8 + k]; const uint32_t h = to_tf32(v), l = to_tf32(v - __uint_as_float(h));
    const int c = tid, r = k;
    const uint32_t off = (c >> 5) * 1024 + r * 128 + ((((c & 31) >> 3) ^ (r & 3)) << 5) + (c & 7) * 4;
    *(uint32_t*)(Am_hi + off) = h; *(uint32_t*)(Am_lo + off) = l;
    if (tid < 32) {
      const float w = B[tid * 8 + k]; const uint32_t bh = to_tf32(w), bl = to_tf32(w - __uint_as_float(bh));
      *(uint32_t*)(Bm_hi + off) = bh; *(uint32_t*)(Bm_lo + off) = bl;
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base, t_row = tmem + ((uint32_t)(warp * 32) << 16);
  {   // A operand into TMEM columns 32..39 (hi) and 40..47 (lo)
    uint32_t h[8], l[8];
    for (int k = 0; k < 8; k++) { const float v = A[tid * 8 + k]; h[k] = to_tf32(v); l[k] = to_tf32(v - __uint_as_float(h[k])); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(t_row + 32u), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(t_row + 40u), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t phase = 0;
  for (int path = 0; path < 3; path++) {
    if (tid == 0) {
      const uint32_t idk = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      if (path == 0) {
        mma_ss(tmem, desc(smem_u32(Ak_lo), 128, 256, 0), desc(smem_u32(Bk_hi), 128, 256, 0), idk, 0u);
        mma_ss(tmem, desc(smem_u32(Ak_hi), 128, 256, 0), desc(smem_u32(Bk_lo), 128, 256, 0), idk, 1u);
        mma_ss(tmem, desc(smem_u32(Ak_hi), 128, 256, 0), desc(smem_u32(Bk_hi), 128, 256, 0), idk, 1u);
      } else if (path == 1) {
        mma_ts(tmem, tmem + 40u, desc(smem_u32(Bk_hi), 128, 256, 0), idk, 0u);
        mma_ts(tmem, tmem + 32u, desc(smem_u32(Bk_lo), 128, 256, 0), idk, 1u);
        mma_ts(tmem, tmem + 32u, desc(smem_u32(Bk_hi), 128, 256, 0), idk, 1u);
      } else {
        const uint32_t idm = idk | (1u << 15) | (1u << 16);
        mma_ss(tmem, desc(smem_u32(Am_lo), 1024, 512, 1), desc(smem_u32(Bm_hi), 1024, 512, 1), idm, 0u);
        mma_ss(tmem, desc(smem_u32(Am_hi), 1024, 512, 1), desc(smem_u32(Bm_lo), 1024, 512, 1), idm, 1u);
        mma_ss(tmem, desc(smem_u32(Am_hi), 1024, 512, 1), desc(smem_u32(Bm_hi), 1024, 512, 1), idm, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(t_row));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; i++) out[((size_t)path * 128 + tid) * 32 + i] = __uint_as_float(r[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  static float A[128 * 8], B[32 * 8], out[3][128][32];
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 32768.0f - 1.0f + ((s >> 24) & 0xff) * 1e-6f; };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  float *dA, *dB, *dO;
  cudaMalloc(&dA, sizeof(A)); cudaMalloc(&dB, sizeof(B)); cudaMalloc(&dO, sizeof(out));
  cudaMemcpy(dA, A, sizeof(A), cudaMemcpyHostToDevice); cudaMemcpy(dB, B, sizeof(B), cudaMemcpyHostToDevice);
  const size_t smem = 32768 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dA, dB, dO);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  cudaMemcpy(out, dO, sizeof(out), cudaMemcpyDeviceToHost);
  const char* names[3] = {"SS K-major", "TS (A in TMEM)", "SS MN-major BASE32B"};
  for (int p = 0; p < 3; p++) {
    double emax = 0, e1 = 0;
    for (int m = 0; m < 128; m++)
      for (int n = 0; n < 32; n++) {
        double ref = 0, ref1 = 0;
        for (int k = 0; k < 8; k++) ref += (double)A[m * 8 + k] * B[n * 8 + k];
        emax = fmax(emax, fabs(out[p][m][n] - ref));
      }
    printf("%-22s max |err| = %.3e   D[3][5] = %.8f\n", names[p], emax, out[p][3][5]);
    (void)e1;
  }
  return 0;
}
