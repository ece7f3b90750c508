// Probe of how tcgen05.mma (kind::tf32, SWIZZLE_NONE) addresses shared memory for K-major and MN-major operands.
// The probed operand's shared-memory region holds word i = float(i) (exact in tf32 below 2048); the other operand is a
// K-major "identity" so that D shows which word the hardware used for element (mn, k).  Build & run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/umma_layout_probe tools/probe/umma_layout_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Variant { uint32_t lbo, sbo, major, probe_a, layout; };   // probe_a: 1 = probe the A operand, 0 = probe B (N = 32)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
  return ((uint64_t)layout << 61) | (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128) probe(const Variant* vars, int nvar, float* out /* [nvar][128][32] */) {
  extern __shared__ __align__(128) unsigned char raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  float* ident = (float*)base;                    // 128 rows x 8 k, canonical K-major: (k>>2)*128 + (row>>3)*256 + (row&7)*16 + (k&3)*4
  float* region = (float*)(base + 4096);          // 8192 words = 32 KB
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 1024; i += 128) ident[i] = 0.f;
  __syncthreads();
  if (tid < 8) { const int row = tid, k = tid; ident[((k >> 2) * 128 + (row >> 3) * 256 + (row & 7) * 16 + (k & 3) * 4) / 4] = 1.f; }
  for (int i = tid; i < 8192; i += 128) region[i] = (float)(i & 2047);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  uint32_t phase = 0;
  for (int v = 0; v < nvar; v++) {
    const Variant var = vars[v];
    if (var.probe_a == 2) {     // every thread writes its lane of the A operand: A(m, k) = 8 m + k at TMEM columns 16..23
      const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 16u;
      uint32_t w[8];
      for (int k = 0; k < 8; k++) w[k] = __float_as_uint((float)(8 * tid + k));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) {
        const uint64_t d_id = make_desc(smem_u32(ident), 128, 256);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem), "r"(tmem + 16u), "l"(d_id), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
    } else if (tid == 0) {
      const uint64_t d_id = make_desc(smem_u32(ident), 128, 256);
      const uint64_t d_pr = make_desc(smem_u32(region), var.lbo, var.sbo, var.layout);
      uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      idesc |= var.major << (var.probe_a ? 15 : 16);
      const uint64_t da = var.probe_a ? d_pr : d_id, db = var.probe_a ? d_id : d_pr;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; i++) out[((size_t)v * 128 + tid) * 32 + i] = __uint_as_float(r[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  const Variant h[] = {
      {2048, 512, 1, 0, 1},    // B MN-major 128B_BASE32B, LBO != SBO
      {512, 2048, 1, 0, 1},
      {2048, 512, 1, 1, 1},    // A MN-major 128B_BASE32B
      {512, 2048, 1, 1, 1},
      {0, 0, 0, 2, 0},         // A from TMEM (tcgen05.st), B = K-major identity
  };
  const int nvar = sizeof(h) / sizeof(h[0]);
  Variant* dv; float* dout;
  cudaMalloc(&dv, sizeof(h)); cudaMemcpy(dv, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, (size_t)nvar * 128 * 32 * 4); cudaMemset(dout, 0xff, (size_t)nvar * 128 * 32 * 4);
  const size_t smem = 4096 + 32768 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dv, nvar, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static float out[16][128][32];
  cudaMemcpy(out, dout, (size_t)nvar * 128 * 32 * 4, cudaMemcpyDeviceToHost);
  for (int v = 0; v < nvar; v++) {
    printf("variant %d: lbo=%u sbo=%u major=%s probe=%s layout=%u\n", v, h[v].lbo, h[v].sbo, h[v].major ? "MN" : "K", h[v].probe_a ? "A" : "B", h[v].layout);
    if (h[v].probe_a == 0) {          // D[m=k][n] = B(n, k): print word index for n = 0..11, k = 0..7
      for (int k = 0; k < 8; k++) { printf("  k=%d:", k); for (int n = 0; n < 12; n++) printf(" %5.0f", out[v][k][n]); printf(" | n=16:%5.0f n=31:%5.0f\n", out[v][k][16], out[v][k][31]); }
    } else {                      // D[m][n=k] = A(m, k): print m = 0..11 and a few more
      for (int k = 0; k < 8; k++) { printf("  k=%d:", k); for (int m = 0; m < 12; m++) printf(" %5.0f", out[v][m][k]); printf(" | m=16:%5.0f m=32:%5.0f m=64:%5.0f m=127:%5.0f\n", out[v][16][k], out[v][32][k], out[v][64][k], out[v][127][k]); }
    }
  }
  return 0;
}
