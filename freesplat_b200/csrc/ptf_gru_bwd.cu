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
// grid = (persistent row-tile CTAs, column blocks of <= kColBlock output columns).  A CTA stages its block of W^T once and then
// walks over 128-pair row tiles: load the dY rows, 24 MMAs, epilogue.  112 KB of shared memory and 128 TMEM columns: two CTAs
// per SM overlap one another's load / MMA / epilogue phases (the first version -- one tile per CTA, all N columns, 154 KB, one
// CTA per SM -- took 100 us per launch with 4 warps per SM, all of it latency).
// shared memory: A hi | A lo (4 rounds x bTile(128) each) | B hi | B lo (4 rounds x bTile(Npc) each) | Ctl
constexpr int kColBlock = 96;
struct DataEpilogue {                    // mode 3 (the update-gate chain rule fused into the dU product, see the header)
  const float* h; int ldh;               // A1 (its first 64 columns are the global latent h)
  const float* r_lin;                    // [M,64]
  float* dr_lin;                         // [M,64]
};

template <int MODE>
__global__ void __launch_bounds__(kNT) gru_bwd_data_kernel(int M, int N, const float* __restrict__ A, int lda,
                                                           const float* __restrict__ W, const float* __restrict__ mask, int ldm,
                                                           float* __restrict__ C, int ldc, DataEpilogue ep, uint32_t tmem_cols) {
  extern __shared__ __align__(128) unsigned char gb_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gb_smem_raw) + 127) & ~(uintptr_t)127);
  constexpr int kRounds = kKd / kRound;                              // 4
  const int n0 = blockIdx.y * kColBlock;                             // this CTA's output columns [n0, n0 + Nc)
  const int Nc = min(kColBlock, N - n0), Npc = (Nc + 15) & ~15;
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + kRounds * bTile(128);
  unsigned char* b_hi = a_lo + kRounds * bTile(128);
  unsigned char* b_lo = b_hi + kRounds * bTile(Npc);
  Ctl& ctl = *reinterpret_cast<Ctl*>(b_lo + kRounds * bTile(Npc));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (M + kNT - 1) / kNT;
  if ((int)blockIdx.x >= ntiles) return;                             // CTA-uniform, before any barrier / TMEM allocation
  // padding rows of the weight operand (Nc..Npc) must be zero
  if (Npc > Nc)
    for (int r = 0; r < kRounds; r++)
      for (int i = tid; i < (Npc - Nc) * kRound; i += kNT) {
        const uint32_t off = (uint32_t)r * bTile(Npc) + op_off(Nc + i / kRound, i % kRound, Npc);
        *reinterpret_cast<uint32_t*>(b_hi + off) = 0u; *reinterpret_cast<uint32_t*>(b_lo + off) = 0u;
      }
  const uint32_t tmem = setup_tmem(ctl, tmem_cols, tid, warp);
  // B = (W[:, n0 : n0 + Nc])^T: operand row n, K index k = W[k][n0 + n]; K quads warp, warp + 4, warp + 8, warp + 12
  stage_transposed<kColBlock / 32, 4>(b_hi, b_lo, Npc, 0, W + n0, N, Nc, 0, kKd, warp, 4, lane);
  const uint32_t bar = smem_u32(&ctl.bar);
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0u;
  // Software pipeline over the row tiles: the dY rows (and the ReLU mask rows) of tile t + 1 are requested right after the MMAs of
  // tile t are issued and arrive while those run and tile t's epilogue drains -- the global-memory latency is paid once per CTA,
  // not once per tile (ncu of the un-pipelined loop: 12.4 long-scoreboard stalls per issue, 8 warps per SM, 2 TB/s).
  float4 av[kKd / 4], mk[kKd / 4];
  uint32_t keep_bits[2] = {0u, 0u};
  auto request_rows = [&](int tile) {
    const long long m = (long long)tile * kNT + tid;
    const bool ok = m < M;
    const float4* ap = reinterpret_cast<const float4*>(A + m * lda);
    const float4* mp = reinterpret_cast<const float4*>(mask + m * ldm);
#pragma unroll
    for (int k4 = 0; k4 < kKd / 4; k4++) av[k4] = ok ? __ldg(ap + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 1) {
#pragma unroll
      for (int q = 0; q < 16; q++) mk[q] = ok ? __ldg(mp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto mask_bits = [&]() {                                            // MODE 1 (N == 64): the ReLU mask of this thread's row as 64 bits
    keep_bits[0] = keep_bits[1] = 0u;
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const uint32_t b4 = (mk[q].x > 0.f ? 1u : 0u) | (mk[q].y > 0.f ? 2u : 0u) | (mk[q].z > 0.f ? 4u : 0u) | (mk[q].w > 0.f ? 8u : 0u);
      keep_bits[q >> 3] |= b4 << (4 * (q & 7));
    }
  };
  request_rows(blockIdx.x);
  if (MODE == 1) mask_bits();
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long m = (long long)tile * kNT + tid;
    const bool active = m < M;
    // A: this thread's row of dY (the previous tile's MMAs have completed: its epilogue waited for them)
#pragma unroll
    for (int k4 = 0; k4 < kKd / 4; k4++)
      store_a4(a_hi + (k4 >> 2) * bTile(128), a_lo + (k4 >> 2) * bTile(128), tid, (k4 & 3) * 4, av[k4].x, av[k4].y, av[k4].z, av[k4].w);
    publish_operands();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aH = smem_u32(a_hi), aL = smem_u32(a_lo), bH = smem_u32(b_hi), bL = smem_u32(b_lo);
      const uint32_t id = idesc(Npc);
#pragma unroll
      for (int r = 0; r < kRounds; r++)
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const uint32_t ao = r * bTile(128) + s * (128 / 8) * 256, bo = r * bTile(Npc) + s * (Npc / 8) * 256;
          mma_tf32(tmem, make_desc(aL + ao), make_desc(bH + bo), id, (r > 0 || s > 0) ? 1u : 0u);
          mma_tf32(tmem, make_desc(aH + ao), make_desc(bL + bo), id, 1u);
          mma_tf32(tmem, make_desc(aH + ao), make_desc(bH + bo), id, 1u);
        }
      commit(bar);
    }
    // MODE 3, gate columns: h and r_lin of this tile's rows, requested while the MMAs run
    float4 hv[16], rv[16];
    const bool gate = MODE == 3 && n0 == 0;
    if (gate) {
      const float4* hp = reinterpret_cast<const float4*>(ep.h + m * ep.ldh);
      const float4* rp = reinterpret_cast<const float4*>(ep.r_lin + m * 64);
#pragma unroll
      for (int q = 0; q < 16; q++) {
        hv[q] = active ? __ldg(hp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        rv[q] = active ? __ldg(rp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (active) {                                                   // dA1[:, 64:88] = 0: e_h enters only through the first layers
#pragma unroll
        for (int q = 0; q < 6; q++) *reinterpret_cast<float4*>(C + m * ldc + 64 + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const int next = tile + gridDim.x;
    if (MODE != 3 && next < ntiles) request_rows(next);               // (MODE 3 holds h / r_lin in those registers until its epilogue)
    mbar_wait(bar, phase); phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int pc = 0; pc < kColBlock / 16; pc++) {
      const int c0 = 16 * pc;
      if (c0 >= Nc) break;                                            // CTA-uniform
      float v[16];
      tmem_ld16(t_row + (uint32_t)c0, v);                             // warp-collective: every lane takes part
      if (!active) continue;
      const int c = n0 + c0;                                          // first output column of this piece
      if (MODE == 3) {
        // P = dU.  c < 64: dr_lin = P h r (1 - r), dA1[:, c] += P r  (r = sigmoid(r_lin));  c >= 64: dA1[:, c + 24] = P
        if (gate && pc < 4) {
          float4* dp = reinterpret_cast<float4*>(ep.dr_lin + m * 64 + c);
          float* cp = C + m * ldc + c;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const float4 h = hv[(pc & 3) * 4 + q], rl = rv[(pc & 3) * 4 + q];
            const float r0 = 1.0f / (1.0f + expf(-rl.x)), r1 = 1.0f / (1.0f + expf(-rl.y)), r2 = 1.0f / (1.0f + expf(-rl.z)),
                        r3 = 1.0f / (1.0f + expf(-rl.w));
            dp[q] = make_float4(v[4 * q] * h.x * r0 * (1.0f - r0), v[4 * q + 1] * h.y * r1 * (1.0f - r1), v[4 * q + 2] * h.z * r2 * (1.0f - r2),
                                v[4 * q + 3] * h.w * r3 * (1.0f - r3));
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * q), "f"(v[4 * q] * r0), "f"(v[4 * q + 1] * r1),
                         "f"(v[4 * q + 2] * r2), "f"(v[4 * q + 3] * r3) : "memory");
          }
        } else {
          float* cp = C + m * ldc + c + 24;
#pragma unroll
          for (int q = 0; q < 4; q++)
            if (c0 + 4 * q < Nc) *reinterpret_cast<float4*>(cp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        continue;
      }
      float* cp = C + m * ldc + c;
      if (MODE == 1) {
        const uint32_t bits = keep_bits[(c0 >> 5) & 1] >> (c0 & 16);
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = ((bits >> e) & 1u) ? v[e] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (c0 + 4 * q >= Nc) break;                                  // N % 4 == 0
        if (MODE == 2)                                                // accumulate: fire-and-forget 16-byte reduction
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]),
                       "f"(v[4 * q + 3]) : "memory");
        else
          *reinterpret_cast<float4*>(cp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    if (MODE == 3 && next < ntiles) request_rows(next);
    if (MODE == 1 && next < ntiles) mask_bits();
    // the next tile's MMAs overwrite the accumulator: order this tile's TMEM reads before them
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
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
  // K quads warp and warp + 4 of a 32-pair step, every source in flight at once (2 x 11 x 4 loads per thread).  The rows of step
  // t + 1 are requested right after the MMAs of step t are issued: their latency overlaps the MMAs instead of following them.
  constexpr int kIt = 11;                                             // 352 virtual rows / 32 lanes: [Y0 | Y1 | X0 | X1]
  float v[2][kIt][4];
  auto request = [&](int chunk) {
    const long long k0 = (long long)chunk * kChunk;
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
  };
  auto deposit = [&]() {
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
  };
  request(blockIdx.x);
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    if (!first) {                                                     // the previous step's MMAs have read the operands
      mbar_wait(bar, phase); phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    deposit();
    publish_operands();
    if (warp == 0 && elect_one()) {
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
    if (chunk + (int)gridDim.x < nchunks) request(chunk + gridDim.x);
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
  const int ncb = (a.N + kColBlock - 1) / kColBlock;
  const int Npc = ((a.N < kColBlock ? a.N : kColBlock) + 15) / 16 * 16;                       // widest column block
  const size_t smem = 2 * (kKd / tc::kRound) * (size_t)(tc::bTile(128) + tc::bTile(Npc)) + sizeof(Ctl) + 128;
  static int sms = 0;
  if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const int ntiles = (a.M + kNT - 1) / kNT;
  int gx = (2 * sms + ncb - 1) / ncb;                                                           // two resident CTAs per SM in total
  if (gx > ntiles) gx = ntiles;
  const DataEpilogue ep{a.h, a.ldh, a.r_lin, a.dr_lin};
  auto go = [&](auto kernel) -> int {
    if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(gru_bwd_data_kernel)")) return rc;
    kernel<<<dim3(gx, ncb), kNT, smem, s>>>(a.M, a.N, a.A, a.lda, a.W, a.mask ? a.mask : a.A, a.mask ? a.ldm : a.lda, a.C, a.ldc, ep,
                                            tmem_cols_for(Npc));
    return FS_OK;
  };
  int rc = a.mode == 0 ? go(gru_bwd_data_kernel<0>) : a.mode == 1 ? go(gru_bwd_data_kernel<1>) : a.mode == 2 ? go(gru_bwd_data_kernel<2>)
                                                                                                  : go(gru_bwd_data_kernel<3>);
  if (rc) return rc;
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
