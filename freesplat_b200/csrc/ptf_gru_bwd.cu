// ptf_gru_bwd.cu -- the matrix products of the PTF GRU's backward pass (training: autograd through
// networks.py:201-214) on the tensor cores: tcgen05.mma kind::tf32 with the 3xTF32 split (a_lo b_hi + a_hi b_lo + a_hi b_hi,
// fp32 accumulation in TMEM), the same arithmetic as the forward kernel (ptf.cu, gru::ptf_gru_tc_kernel).
//
//   gru_bwd_data_kernel    dX[M,N]  (op)=  dY[M,64] . W[64,N]              one 128-pair row tile per CTA
//        the six "gradient w.r.t. the layer input" products; epilogues: plain store, ReLU mask (H > 0) of the layer
//        below, or accumulate into dX (the two first-layer products that land in the same dA1).
//   gru_bwd_weights_kernel G[128,Np] += [Y0 | Y1]^T . [X0 | X1 | 1]        reduction over the M pairs, persistent CTAs
//        two 64-row output gradients stacked on the 128 MMA rows against up to two input matrices side by side and a column
//        of ones: one launch yields two weight gradients (the diagonal blocks, or both row blocks when they share the input)
//        and both bias gradients (the ones column).
//
// Both operands of the weight-gradient product and the weight operand of the data product are needed TRANSPOSED (the
// reduction index -- pairs, or W's rows -- must be the contiguous K index of the shared-memory operand): stage_transposed()
// reads the matrices with lanes along their contiguous dimension (coalesced 128-byte rows) and four consecutive K rows per
// thread, so that every store is one conflict-free 16-byte word of the canonical K-major layout.
#include "common.cuh"
#include "tc_tf32.cuh"

namespace fs {
namespace grubwd {
using namespace tc;

constexpr int kNT = 128;                 // threads per CTA = rows of one MMA = TMEM lanes
constexpr int kKd = 64;                  // reduction length of every data product (all hidden / output widths are 64)
constexpr int kChunk = 32;               // pairs per weight-gradient step (2 operand rounds = 4 UMMA k-steps)

// Operand rows [row0, row0 + nf) of a K-major operand with `rows` rows for NQ K-quads q = q_first + b * q_stride: element
// (row0 + f, 4 q + e) = src[(k_begin + 4 q + e) * ld + f], zero where that row index reaches k_end.  hi / lo: regions of rounds
// of bTile(rows) bytes.  All NQ x MAXIT x 4 loads (read-only path) are issued before the first store: one memory latency per
// call -- with a load-split-store loop per element group the staging was a chain of ~22 dependent latencies (275 us per
// weight-gradient launch, profiles/r2_ptf_train_profile_tc.txt, first version).
template <int MAXIT, int NQ>
__device__ __forceinline__ void stage_transposed(unsigned char* hi, unsigned char* lo, int rows, int row0, const float* __restrict__ src, int ld,
                                                 int nf, long long k_begin, long long k_end, int q_first, int q_stride, int lane) {
  float v[NQ][MAXIT][4];
#pragma unroll
  for (int b = 0; b < NQ; b++) {
    const long long k0 = k_begin + 4 * (q_first + b * q_stride);
    const float* s0 = src + k0 * ld;
#pragma unroll
    for (int i = 0; i < MAXIT; i++) {
      const int f = lane + 32 * i;
#pragma unroll
      for (int e = 0; e < 4; e++) v[b][i][e] = (f < nf && k0 + e < k_end) ? __ldg(s0 + (long long)e * ld + f) : 0.f;
    }
  }
#pragma unroll
  for (int b = 0; b < NQ; b++) {
    const int q = q_first + b * q_stride;
    const uint32_t tile_off = (uint32_t)(q >> 2) * (uint32_t)bTile(rows);
#pragma unroll
    for (int i = 0; i < MAXIT; i++) {
      const int f = lane + 32 * i;
      if (f < nf) {
        uint4 h, l;
        split_tf32(v[b][i][0], h.x, l.x); split_tf32(v[b][i][1], h.y, l.y); split_tf32(v[b][i][2], h.z, l.z); split_tf32(v[b][i][3], h.w, l.w);
        const uint32_t off = tile_off + op_off(row0 + f, (q & 3) * 4, rows);
        *reinterpret_cast<uint4*>(hi + off) = h; *reinterpret_cast<uint4*>(lo + off) = l;
      }
    }
  }
}

struct Ctl {                             // tail of the dynamic shared memory
  unsigned long long bar;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t setup_tmem(Ctl& c, uint32_t cols, int tid, int warp) {
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&c.bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&c.tmem_base)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return c.tmem_base;
}
__device__ __forceinline__ void publish_operands() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ data gradient
// shared memory: A hi | A lo (4 rounds x bTile(128) each) | B hi | B lo (4 rounds x bTile(Np) each) | Ctl
__global__ void __launch_bounds__(kNT) gru_bwd_data_kernel(int M, int N, int Np, int mode, const float* __restrict__ A, int lda,
                                                           const float* __restrict__ W, const float* __restrict__ mask, int ldm,
                                                           float* __restrict__ C, int ldc, uint32_t tmem_cols) {
  extern __shared__ __align__(128) unsigned char gb_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gb_smem_raw) + 127) & ~(uintptr_t)127);
  constexpr int kRounds = kKd / kRound;                              // 4
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + kRounds * bTile(128);
  unsigned char* b_hi = a_lo + kRounds * bTile(128);
  unsigned char* b_lo = b_hi + kRounds * bTile(Np);
  Ctl& ctl = *reinterpret_cast<Ctl*>(b_lo + kRounds * bTile(Np));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m = (long long)blockIdx.x * kNT + tid;
  const bool active = m < M;
  // padding rows of the weight operand (N..Np) must be zero
  if (Np > N)
    for (int r = 0; r < kRounds; r++)
      for (int i = tid; i < (Np - N) * kRound; i += kNT) {
        const uint32_t off = (uint32_t)r * bTile(Np) + op_off(N + i / kRound, i % kRound, Np);
        *reinterpret_cast<uint32_t*>(b_hi + off) = 0u; *reinterpret_cast<uint32_t*>(b_lo + off) = 0u;
      }
  const uint32_t tmem = setup_tmem(ctl, tmem_cols, tid, warp);
  // B = W^T: operand row n, K index k = W[k][n]
  stage_transposed<6, 2>(b_hi, b_lo, Np, 0, W, N, N, 0, kKd, warp, 4, lane);          // K quads warp, warp + 4
  stage_transposed<6, 2>(b_hi, b_lo, Np, 0, W, N, N, 0, kKd, warp + 8, 4, lane);      //         warp + 8, warp + 12
  // A: this thread's row of dY
  {
    const float4* ap = reinterpret_cast<const float4*>(A + m * lda);
#pragma unroll
    for (int k4 = 0; k4 < kKd / 4; k4++) {
      const float4 v = active ? __ldg(ap + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
      store_a4(a_hi + (k4 >> 2) * bTile(128), a_lo + (k4 >> 2) * bTile(128), tid, (k4 & 3) * 4, v.x, v.y, v.z, v.w);
    }
  }
  publish_operands();
  const uint32_t bar = smem_u32(&ctl.bar);
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t aH = smem_u32(a_hi), aL = smem_u32(a_lo), bH = smem_u32(b_hi), bL = smem_u32(b_lo);
    const uint32_t id = idesc(Np);
    for (int r = 0; r < kRounds; r++)
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const uint32_t ao = r * bTile(128) + s * (128 / 8) * 256, bo = r * bTile(Np) + s * (Np / 8) * 256;
        mma_tf32(tmem, make_desc(aL + ao), make_desc(bH + bo), id, (r > 0 || s > 0) ? 1u : 0u);
        mma_tf32(tmem, make_desc(aH + ao), make_desc(bL + bo), id, 1u);
        mma_tf32(tmem, make_desc(aH + ao), make_desc(bH + bo), id, 1u);
      }
    commit(bar);
  }
  // mode 1 (N == 64): the ReLU mask of this thread's row as 64 bits, fetched while the MMAs run
  uint32_t keep_bits[2] = {0u, 0u};
  if (mode == 1 && active) {
    const float4* mp = reinterpret_cast<const float4*>(mask + m * ldm);
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const float4 h = __ldg(mp + q);
      const uint32_t b4 = (h.x > 0.f ? 1u : 0u) | (h.y > 0.f ? 2u : 0u) | (h.z > 0.f ? 4u : 0u) | (h.w > 0.f ? 8u : 0u);
      keep_bits[q >> 3] |= b4 << (4 * (q & 7));
    }
  }
  mbar_wait(bar, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(t_row + (uint32_t)c0, v);                               // warp-collective: every lane takes part
    if (!active) continue;
    float* cp = C + m * ldc + c0;
    if (mode == 1) {
      const uint32_t bits = keep_bits[(c0 >> 5) & 1] >> (c0 & 16);
#pragma unroll
      for (int e = 0; e < 16; e++) v[e] = ((bits >> e) & 1u) ? v[e] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (c0 + 4 * q >= N) break;                                     // N % 4 == 0
      if (mode == 2)                                                  // accumulate: fire-and-forget 16-byte reduction
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]),
                     "f"(v[4 * q + 3]) : "memory");
      else
        *reinterpret_cast<float4*>(cp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------------------ weight gradient
// shared memory: A hi | A lo (2 rounds x bTile(128)) | B hi | B lo (2 rounds x bTile(Np)) | Ctl
__global__ void __launch_bounds__(kNT) gru_bwd_weights_kernel(int M, const float* __restrict__ Y0, int ldy0, const float* __restrict__ Y1, int ldy1,
                                                              const float* __restrict__ X0, int ldx0, int nx0, const float* __restrict__ X1,
                                                              int ldx1, int nx1, int Np, float* __restrict__ G, uint32_t tmem_cols) {
  extern __shared__ __align__(128) unsigned char gb_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gb_smem_raw) + 127) & ~(uintptr_t)127);
  constexpr int kRounds = kChunk / kRound;                           // 2
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + kRounds * bTile(128);
  unsigned char* b_hi = a_lo + kRounds * bTile(128);
  unsigned char* b_lo = b_hi + kRounds * bTile(Np);
  Ctl& ctl = *reinterpret_cast<Ctl*>(b_lo + kRounds * bTile(Np));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = (M + kChunk - 1) / kChunk;
  if ((int)blockIdx.x >= nchunks) return;                            // CTA-uniform, before any barrier / TMEM allocation
  // zero both operands once (unused Y1 rows and the padding rows of B stay zero), then the row of ones: hi = 1, lo = 0
  {
    const int words = (2 * kRounds * bTile(128) + 2 * kRounds * bTile(Np)) / 16;
    for (int i = tid; i < words; i += kNT) reinterpret_cast<uint4*>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const int ones = nx0 + nx1;
    for (int k = tid; k < kChunk; k += kNT)
      *reinterpret_cast<float*>(b_hi + (k >> 4) * bTile(Np) + op_off(ones, k & 15, Np)) = 1.0f;
  }
  const uint32_t tmem = setup_tmem(ctl, tmem_cols, tid, warp);
  const uint32_t bar = smem_u32(&ctl.bar);
  uint32_t phase = 0u;
  bool first = true;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    if (!first) {                                                     // the previous step's MMAs have read the operands
      mbar_wait(bar, phase); phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const long long k0 = (long long)chunk * kChunk;
    // K quads warp and warp + 4 of the 32-pair step, every source in flight at once (up to 2 x 11 x 4 loads per thread)
    {
      constexpr int kIt = 11;                                         // 352 virtual rows / 32 lanes: [Y0 | Y1 | X0 | X1]
      float v[2][kIt][4];
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const long long kq = k0 + 4 * (warp + 4 * b);
#pragma unroll
        for (int i = 0; i < kIt; i++) {
          const float* p = nullptr; int ld = 0;
          if (i < 2) { p = Y0 + kq * ldy0 + (lane + 32 * i); ld = ldy0; }
          else if (i < 4) { if (Y1) { p = Y1 + kq * ldy1 + (lane + 32 * (i - 2)); ld = ldy1; } }
          else {
            const int xf = lane + 32 * (i - 4);
            if (xf < nx0) { p = X0 + kq * ldx0 + xf; ld = ldx0; }
            else if (xf < nx0 + nx1) { p = X1 + kq * ldx1 + (xf - nx0); ld = ldx1; }
          }
#pragma unroll
          for (int e = 0; e < 4; e++) v[b][i][e] = (p != nullptr && kq + e < M) ? __ldg(p + (long long)e * ld) : 0.f;
        }
      }
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int q = warp + 4 * b;
#pragma unroll
        for (int i = 0; i < kIt; i++) {
          const int vf = lane + 32 * i;
          if (i < 4 ? (i < 2 || Y1 != nullptr) : (vf - 128 < nx0 + nx1)) {
            uint4 h, l;
            split_tf32(v[b][i][0], h.x, l.x); split_tf32(v[b][i][1], h.y, l.y); split_tf32(v[b][i][2], h.z, l.z); split_tf32(v[b][i][3], h.w, l.w);
            if (i < 4) {
              const uint32_t off = (uint32_t)(q >> 2) * (uint32_t)bTile(128) + op_off(vf, (q & 3) * 4, 128);
              *reinterpret_cast<uint4*>(a_hi + off) = h; *reinterpret_cast<uint4*>(a_lo + off) = l;
            } else {
              const uint32_t off = (uint32_t)(q >> 2) * (uint32_t)bTile(Np) + op_off(vf - 128, (q & 3) * 4, Np);
              *reinterpret_cast<uint4*>(b_hi + off) = h; *reinterpret_cast<uint4*>(b_lo + off) = l;
            }
          }
        }
      }
    }
    publish_operands();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aH = smem_u32(a_hi), aL = smem_u32(a_lo), bH = smem_u32(b_hi), bL = smem_u32(b_lo);
      const uint32_t id = idesc(Np);
#pragma unroll
      for (int r = 0; r < kRounds; r++)
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const uint32_t ao = r * bTile(128) + s * (128 / 8) * 256, bo = r * bTile(Np) + s * (Np / 8) * 256;
          mma_tf32(tmem, make_desc(aL + ao), make_desc(bH + bo), id, (!first || r > 0 || s > 0) ? 1u : 0u);
          mma_tf32(tmem, make_desc(aH + ao), make_desc(bL + bo), id, 1u);
          mma_tf32(tmem, make_desc(aH + ao), make_desc(bH + bo), id, 1u);
        }
      commit(bar);
    }
    first = false;
  }
  mbar_wait(bar, phase);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // this CTA's partial sums -> G (row = TMEM lane = tid); one 16-byte reduction per four columns
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  const bool row_used = tid < 64 || Y1 != nullptr;
  for (int c0 = 0; c0 < Np; c0 += 16) {
    float v[16];
    tmem_ld16(t_row + (uint32_t)c0, v);
    if (!row_used) continue;
    float* gp = G + (size_t)tid * Np + c0;
#pragma unroll
    for (int q = 0; q < 4; q++)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gp + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]),
                   "f"(v[4 * q + 3]) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

static uint32_t tmem_cols_for(int n) { uint32_t c = 32; while ((int)c < n) c <<= 1; return c; }

}  // namespace grubwd

int launch_ptf_gru_bwd_data(const FsGruBwdDataArgs& a, cudaStream_t s) {
  using namespace grubwd;
  if (a.M == 0) return FS_OK;
  const int Np = (a.N + 15) / 16 * 16;
  const size_t smem = 2 * (kKd / tc::kRound) * (size_t)(tc::bTile(128) + tc::bTile(Np)) + sizeof(Ctl) + 128;
  if (int rc = check_cuda(cudaFuncSetAttribute(gru_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(gru_bwd_data_kernel)")) return rc;
  gru_bwd_data_kernel<<<(a.M + kNT - 1) / kNT, kNT, smem, s>>>(a.M, a.N, Np, a.mode, a.A, a.lda, a.W, a.mask ? a.mask : a.A, a.mask ? a.ldm : a.lda,
                                                               a.C, a.ldc, tmem_cols_for(Np));
  return check_cuda(cudaGetLastError(), "gru_bwd_data_kernel");
}

int launch_ptf_gru_bwd_weights(const FsGruBwdWeightsArgs& a, cudaStream_t s) {
  using namespace grubwd;
  const int Np = a.ldg;
  if (int rc = check_cuda(cudaMemsetAsync(a.G, 0, (size_t)128 * Np * sizeof(float), s), "cudaMemsetAsync(G)")) return rc;
  if (a.M == 0) return FS_OK;
  const size_t smem = 2 * (kChunk / tc::kRound) * (size_t)(tc::bTile(128) + tc::bTile(Np)) + sizeof(Ctl) + 128;
  if (int rc = check_cuda(cudaFuncSetAttribute(gru_bwd_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(gru_bwd_weights_kernel)")) return rc;
  static int sms = 0;
  if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const int nchunks = (a.M + kChunk - 1) / kChunk;
  const int grid = nchunks < 2 * sms ? nchunks : 2 * sms;            // two 88 KB / 256-TMEM-column CTAs per SM
  gru_bwd_weights_kernel<<<grid, kNT, smem, s>>>(a.M, a.Y0, a.ldy0, a.Y1, a.ldy1, a.X0, a.ldx0, a.nx0, a.X1, a.ldx1, a.nx1, Np, a.G,
                                                 tmem_cols_for(Np));
  return check_cuda(cudaGetLastError(), "gru_bwd_weights_kernel");
}

}  // namespace fs
