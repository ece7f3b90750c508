// tc_tf32.cuh -- the small tcgen05 / TMEM vocabulary the 3xTF32 kernels of the PTF GRU share (ptf.cu: forward,
// ptf_gru_bwd.cu: the data / weight gradient products of its backward).  Operands are staged by the threads themselves in the
// canonical no-swizzle K-major layout (8 x 16 B core matrices, LBO 128 B, SBO 256 B); accumulators live in TMEM.
#pragma once
#include "common.cuh"

namespace fs {
namespace tc {

constexpr int kRound = 16;                                   // K columns per operand round = 2 UMMA k-steps of 8
__host__ __device__ constexpr int bTile(int n) { return 2 * (n / 8) * 256; }     // bytes of one operand round of n rows (n x 16 x 4)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
// activation-side tf32 split by TRUNCATION: hi = top 19 bits of x, lo = top 19 bits of (x - hi) (the subtraction is exact).
// cvt.rna.tf32.f32 is three instructions on sm_100a (FSETP + IADD + LOP3): the rounded split cost 7 instructions per element
// and ~15 % of the cost-volume / GRU kernels; the truncated one costs 3.  x = hi + lo holds to 2^-20 |x| (2^-22 rounded):
// the dropped lo*lo term and the split error stay ~1e-6 relative, inside the 1e-4 budget.  Weights keep the rounded split.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da),
               "l"(db), "r"(id), "r"(acc) : "memory");
}
// One elected lane of a converged warp (elect.sync): inside `if (warp == 0 && elect_one())` ptxas knows that exactly one lane runs,
// keeps descriptors and addresses in uniform registers and emits 1-2 instructions per tcgen05.mma; under `if (tid == 0)` every MMA
// was wrapped in an ELECT / R2UR / BRA.U.ANY uniformisation loop plus the descriptor arithmetic: 9-17 instructions per MMA, all on
// the one warp every other warp of the CTA then waits for at the next barrier (cuobjdump -sass, r2).
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0u;
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(bar),
               "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// byte offset of (row, k) inside one operand round (k in 0..15)
__host__ __device__ inline uint32_t op_off(int row, int k, int rows) {
  return (uint32_t)((k >> 3) * ((rows / 8) * 256) + (row >> 3) * 256 + ((k >> 2) & 1) * 128 + (row & 7) * 16 + (k & 3) * 4);
}

// four consecutive K values (k0 % 4 == 0) of one row of a 128-row A round: one 16-byte store per hi / lo tile
__device__ __forceinline__ void store_a4(unsigned char* hi, unsigned char* lo, int row, int k0, float v0, float v1, float v2, float v3) {
  uint4 h, l;
  split_tf32(v0, h.x, l.x); split_tf32(v1, h.y, l.y); split_tf32(v2, h.z, l.z); split_tf32(v3, h.w, l.w);
  const uint32_t off = op_off(row, k0, 128);
  *reinterpret_cast<uint4*>(hi + off) = h; *reinterpret_cast<uint4*>(lo + off) = l;
}

}  // namespace tc
}  // namespace fs
