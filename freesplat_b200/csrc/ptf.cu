// ptf.cu -- Pixel-wise Triplet Fusion index / merge kernels.
//
// Replaces the torch op chain of EncoderFreeSplat.fuse_gaussians
// (/root/reference/src/model/encoder/encoder_freesplat.py:431-522): per fused view the reference runs
// ~40 ops (matmul, round, scatter_reduce amin, isin x2, 12 full-state torch.cat).  Here one step is
//   ptf_project_kernel : q = E_i^-1 X_j, pixel = round_half_even(q.xy/q.z * f + c), z-buffer by atomicMin
//   ptf_match_kernel   : match_j = valid & (z_j == zbuf[p]) & fuse_pix[p];  append_p = !fuse_pix[p];
//                        per-block counts of kept / matched / appended items
//   ptf_offsets_kernel : one block scans the per-block counts -> block offsets and the new total
//   (GRU on the matched pairs: plain GEMMs through cuBLAS, inputs gathered by ptf_gru_gather_kernel)
//   ptf_compact_kernel : writes the new state in the reference's order
//                        [kept (ascending j)] ++ [fused (ascending j)] ++ [unmatched pixels of view i (raster order)]
// Compiled with -fmad=false: the arithmetic is the canonical order of oracle/ptf.py because merged
// coordinates feed the index decisions of the following views (bit-exact indices across all steps).
//
// HBM-bound: algorithmic bytes per step = 16 N (project) + 8 HW (z-buffer) + 344 N_out + 280 HW.
#include <cstdlib>

#include "common.cuh"
#include "tc_tf32.cuh"

namespace fs {

constexpr int kPtfThreads = 256;
constexpr int kPtfItems = 512;    // items per block in the flag scans (2 per thread; 1024 left the merge kernel below one wave)

static inline int ptf_blocks(int n_upper, int HW) {
  const int m = n_upper > HW ? n_upper : HW;
  return (m + kPtfItems - 1) / kPtfItems;
}

// counters (int32[8]) of one step: [0] N_in  [1] n_keep  [2] n_match  [3] n_append  [4] N_out
__global__ void __launch_bounds__(kPtfThreads) ptf_project_kernel(FsPtfArgs a) {
  const int N = a.counts_in[0];
  const int j = blockIdx.x * kPtfThreads + threadIdx.x;
  if (j >= N) return;
  const float x = a.coords[3 * (size_t)j], y = a.coords[3 * (size_t)j + 1], z = a.coords[3 * (size_t)j + 2];
  const float* __restrict__ E = a.E_inv;    // row-major 4x4
  const float qx = fmaf(E[3], 1.0f, fmaf(E[2], z, fmaf(E[1], y, E[0] * x)));
  const float qy = fmaf(E[7], 1.0f, fmaf(E[6], z, fmaf(E[5], y, E[4] * x)));
  const float qz = fmaf(E[11], 1.0f, fmaf(E[10], z, fmaf(E[9], y, E[8] * x)));
  const float fx = a.K_px[0], fy = a.K_px[4], cx = a.K_px[2], cy = a.K_px[5];
  const float u = (qx / qz) * fx + cx;
  const float v = (qy / qz) * fy + cy;
  const float col = rintf(u), row = rintf(v);
  const bool valid = (row >= 0.f) && (row < (float)a.H) && (col >= 0.f) && (col < (float)a.W) && (qz > 0.f);
  int pix = -1;
  if (valid) {
    pix = (int)col + (int)row * a.W;
    atomicMin(a.zbuf + pix, __float_as_uint(qz));      // qz > 0: IEEE bits are monotonic
  }
  a.pix[j] = pix;
  a.zeta[j] = qz;
}

__device__ __forceinline__ bool fuse_pixel(const FsPtfArgs& a, int p) {
  const float zb = __uint_as_float(a.zbuf[p]);
  const float d = a.v_depth[p];
  return fabsf(zb - d) < fmaxf(d * 0.05f, a.depth_thres);
}

// flags: bit0 = matched (for j < N) ; for pixels p < HW: app[p] = 1 if appended
__global__ void __launch_bounds__(kPtfThreads) ptf_match_kernel(FsPtfArgs a) {
  __shared__ int s_cnt[3];
  const int N = a.counts_in[0], HW = a.H * a.W;
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int keep = 0, mat = 0, app = 0;
  const int base = blockIdx.x * kPtfItems;
#pragma unroll
  for (int r = 0; r < kPtfItems / kPtfThreads; r++) {
    const int k = base + r * kPtfThreads + threadIdx.x;
    if (k < N) {
      const int p = a.pix[k];
      bool m = false;
      if (p >= 0) m = (__float_as_uint(a.zeta[k]) == a.zbuf[p]) && fuse_pixel(a, p);
      a.match[k] = m ? 1 : 0;
      mat += m; keep += !m;
    }
    if (k < HW) {
      const bool ap = !fuse_pixel(a, k);
      a.append[k] = ap ? 1 : 0;
      app += ap;
    }
  }
  // block reduction of the three counts
  for (int o = 16; o > 0; o >>= 1) {
    keep += __shfl_xor_sync(0xffffffffu, keep, o); mat += __shfl_xor_sync(0xffffffffu, mat, o);
    app += __shfl_xor_sync(0xffffffffu, app, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], keep); atomicAdd(&s_cnt[1], mat); atomicAdd(&s_cnt[2], app); }
  __syncthreads();
  if (threadIdx.x < 3) a.block_counts[3 * (size_t)blockIdx.x + threadIdx.x] = s_cnt[threadIdx.x];
}

// single block: exclusive scan of block_counts[nb][3] in place; writes the step's counters
__global__ void __launch_bounds__(1024) ptf_offsets_kernel(FsPtfArgs a, int nb) {
  __shared__ int warp_sums[3][32];
  __shared__ int carry[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 3) carry[tid] = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int k = base + tid;
    int c[3], incl[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { c[q] = (k < nb) ? a.block_counts[3 * (size_t)k + q] : 0; incl[q] = c[q]; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (int q = 0; q < 3; q++) { const int t = __shfl_up_sync(0xffffffffu, incl[q], o); if (lane >= o) incl[q] += t; }
    }
    if (lane == 31) { warp_sums[0][warp] = incl[0]; warp_sums[1][warp] = incl[1]; warp_sums[2][warp] = incl[2]; }
    __syncthreads();
    if (warp < 3) {
      int w = warp_sums[warp][lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_sums[warp][lane] = w;
    }
    __syncthreads();
    int endv[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
      endv[q] = carry[q] + (warp ? warp_sums[q][warp - 1] : 0) + incl[q];
      if (k < nb) a.block_counts[3 * (size_t)k + q] = endv[q] - c[q];
    }
    __syncthreads();
    if (tid == 1023) { carry[0] = endv[0]; carry[1] = endv[1]; carry[2] = endv[2]; }
    __syncthreads();
  }
  if (tid == 0) {
    a.counts_out[0] = a.counts_in[0];
    a.counts_out[1] = carry[0]; a.counts_out[2] = carry[1]; a.counts_out[3] = carry[2];
    a.counts_out[4] = carry[0] + carry[1] + carry[2];
  }
}

// rank of a set flag inside its block (items laid out r*256 + tid, like ptf_match_kernel)
struct BlockRanks {
  int rank[kPtfItems / kPtfThreads];
};

template <typename FlagFn>
__device__ __forceinline__ void block_ranks(FlagFn flag, int (&rank)[kPtfItems / kPtfThreads], int* s_warp /*[rounds][8]*/) {
  constexpr int R = kPtfItems / kPtfThreads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned ball[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    ball[r] = __ballot_sync(0xffffffffu, flag(r));
    if (lane == 0) s_warp[r * 8 + warp] = __popc(ball[r]);
  }
  __syncthreads();
  int run = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    int before = run;
    for (int w = 0; w < 8; w++) {
      const int c = s_warp[r * 8 + w];
      if (w < warp) before += c;
      run += c;
    }
    rank[r] = before + __popc(ball[r] & ((1u << lane) - 1u));
  }
  __syncthreads();
}

// list of matched pairs for the GRU: pair_j[m], pair_p[m]  (m = rank among matched, ascending j)
__global__ void __launch_bounds__(kPtfThreads) ptf_pairs_kernel(FsPtfArgs a) {
  __shared__ int s_warp[(kPtfItems / kPtfThreads) * 8];
  const int N = a.counts_in[0];
  const int base = blockIdx.x * kPtfItems;
  if (base >= N) return;
  int rank[kPtfItems / kPtfThreads];
  block_ranks([&](int r) { const int k = base + r * kPtfThreads + threadIdx.x; return k < N && a.match[k]; }, rank, s_warp);
  const int off = a.block_counts[3 * (size_t)blockIdx.x + 1];
#pragma unroll
  for (int r = 0; r < kPtfItems / kPtfThreads; r++) {
    const int k = base + r * kPtfThreads + threadIdx.x;
    if (k < N && a.match[k]) { a.pair_j[off + rank[r]] = k; a.pair_p[off + rank[r]] = a.pix[k]; }
  }
}

__device__ __forceinline__ float wmean(float g, float w0, float v, float w1, float ws) { return (g * w0 + v * w1) / ws; }

// writes the new state.  One block handles kPtfItems old items and kPtfItems pixels of view i.
__global__ void __launch_bounds__(kPtfThreads) ptf_compact_kernel(FsPtfArgs a) {
  constexpr int R = kPtfItems / kPtfThreads;
  __shared__ int s_warp[R * 8];
  const int N = a.counts_in[0], HW = a.H * a.W, F = a.F;
  const int n_keep = a.counts_out[1], n_match = a.counts_out[2];
  const int base = blockIdx.x * kPtfItems;
  if (base >= max(N, HW)) return;            // the grid is sized by an upper bound of N (no host read of the counters)
  const int off_keep = a.block_counts[3 * (size_t)blockIdx.x + 0];
  const int off_mat = a.block_counts[3 * (size_t)blockIdx.x + 1];
  const int off_app = a.block_counts[3 * (size_t)blockIdx.x + 2];
  int rk[R], rm[R], ra[R];
  block_ranks([&](int r) { const int k = base + r * kPtfThreads + threadIdx.x; return k < N && !a.match[k]; }, rk, s_warp);
  block_ranks([&](int r) { const int k = base + r * kPtfThreads + threadIdx.x; return k < N && a.match[k]; }, rm, s_warp);
  block_ranks([&](int r) { const int k = base + r * kPtfThreads + threadIdx.x; return k < HW && a.append[k]; }, ra, s_warp);
  // destination of every item of this block (or -1), staged so that feature rows can be copied by whole warps
  __shared__ int s_dst_old[kPtfItems], s_dst_px[kPtfItems], s_pair[kPtfItems];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int t = r * kPtfThreads + threadIdx.x, k = base + t;
    int d_old = -1, d_px = -1, pr = -1;
    if (k < N) {
      if (a.match[k]) { d_old = n_keep + off_mat + rm[r]; pr = off_mat + rm[r]; }
      else d_old = off_keep + rk[r];
    }
    if (k < HW && a.append[k]) d_px = n_keep + n_match + off_app + ra[r];
    s_dst_old[t] = d_old; s_dst_px[t] = d_px; s_pair[t] = pr;
    if (a.map_old != nullptr && k < N) a.map_old[k] = d_old;
    if (a.map_px != nullptr && k < HW) a.map_px[k] = d_px;
    // ---- scalar / small fields: one thread per item ----
    if (d_old >= 0) {
      if (pr < 0) {
        a.o_coords[3 * (size_t)d_old] = a.coords[3 * (size_t)k]; a.o_coords[3 * (size_t)d_old + 1] = a.coords[3 * (size_t)k + 1];
        a.o_coords[3 * (size_t)d_old + 2] = a.coords[3 * (size_t)k + 2];
        a.o_dens[d_old] = a.dens[k]; a.o_wemb[d_old] = a.wemb[k]; a.o_depth[d_old] = a.depth[k];
        {
          const float4* se = reinterpret_cast<const float4*>(a.ext + 16 * (size_t)k);
          float4* de = reinterpret_cast<float4*>(a.o_ext + 16 * (size_t)d_old);
          const float4 e0 = se[0], e1 = se[1], e2 = se[2], e3 = se[3];
          de[0] = e0; de[1] = e1; de[2] = e2; de[3] = e3;
        }
      } else {
        const int p = a.pix[k];
        const float w0 = a.dens[k], w1 = a.v_dens[p], ws = w0 + w1;
#pragma unroll
        for (int e = 0; e < 3; e++) a.o_coords[3 * (size_t)d_old + e] = wmean(a.coords[3 * (size_t)k + e], w0, a.v_coords[3 * (size_t)p + e], w1, ws);
        a.o_dens[d_old] = ws;
        a.o_wemb[d_old] = a.wemb[k] + a.v_wemb[p];
        a.o_depth[d_old] = wmean(a.depth[k], w0, a.v_depth[p], w1, ws);
#pragma unroll
        for (int e = 0; e < 16; e++) a.o_ext[16 * (size_t)d_old + e] = wmean(a.ext[16 * (size_t)k + e], w0, a.v_ext[e], w1, ws);
      }
    }
    if (d_px >= 0) {
      a.o_coords[3 * (size_t)d_px] = a.v_coords[3 * (size_t)k]; a.o_coords[3 * (size_t)d_px + 1] = a.v_coords[3 * (size_t)k + 1];
      a.o_coords[3 * (size_t)d_px + 2] = a.v_coords[3 * (size_t)k + 2];
      a.o_dens[d_px] = a.v_dens[k]; a.o_wemb[d_px] = a.v_wemb[k]; a.o_depth[d_px] = a.v_depth[k];
      {
        const float4* se = reinterpret_cast<const float4*>(a.v_ext);
        float4* de = reinterpret_cast<float4*>(a.o_ext + 16 * (size_t)d_px);
        de[0] = se[0]; de[1] = se[1]; de[2] = se[2]; de[3] = se[3];
      }
    }
  }
  __syncthreads();
  // ---- feature rows: 16-byte chunks, one (row, chunk) pair per thread and iteration (independent loads in flight) ----
  if ((F & 3) == 0) {
    const int cpr = F >> 2;                                   // float4 chunks per row
    const float4* feats4 = reinterpret_cast<const float4*>(a.feats);
    const float4* gru4 = reinterpret_cast<const float4*>(a.gru_out);
    const float4* vfeats4 = reinterpret_cast<const float4*>(a.v_feats);
    float4* out4 = reinterpret_cast<float4*>(a.o_feats);
#pragma unroll 4
    for (int idx = threadIdx.x; idx < kPtfItems * cpr; idx += kPtfThreads) {
      const int t = idx / cpr, ch = idx - t * cpr;
      const int d_old = s_dst_old[t];
      if (d_old >= 0) {
        const int pr = s_pair[t];
        const float4 v = pr >= 0 ? gru4[(size_t)pr * cpr + ch] : feats4[(size_t)(base + t) * cpr + ch];
        out4[(size_t)d_old * cpr + ch] = v;
      }
    }
#pragma unroll 4
    for (int idx = threadIdx.x; idx < kPtfItems * cpr; idx += kPtfThreads) {
      const int t = idx / cpr, ch = idx - t * cpr;
      const int d_px = s_dst_px[t];
      if (d_px >= 0) out4[(size_t)d_px * cpr + ch] = vfeats4[(size_t)(base + t) * cpr + ch];
    }
  } else {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = warp; t < kPtfItems; t += 8) {
      const int k = base + t;
      const int d_old = s_dst_old[t], d_px = s_dst_px[t], pr = s_pair[t];
      if (d_old >= 0) {
        const float* src = pr >= 0 ? a.gru_out + (size_t)pr * F : a.feats + (size_t)k * F;
        for (int e = lane; e < F; e += 32) a.o_feats[(size_t)d_old * F + e] = src[e];
      }
      if (d_px >= 0) {
        const float* src = a.v_feats + (size_t)k * F;
        for (int e = lane; e < F; e += 32) a.o_feats[(size_t)d_px * F + e] = src[e];
      }
    }
  }
}

// ---- backward of the merge (training).  Same block shape as ptf_compact_kernel: 512 old items and 512 pixels per block;
// scalar fields by one thread per item, feature rows as 16-byte chunks by the whole block (coalesced).
__global__ void __launch_bounds__(kPtfThreads) ptf_merge_bwd_kernel(FsPtfMergeBwdArgs a) {
  constexpr int R = kPtfItems / kPtfThreads;
  const int N = a.N, HW = a.H * a.W, F = a.F;
  const int base = blockIdx.x * kPtfItems;
  __shared__ int s_src_old[kPtfItems], s_src_px[kPtfItems], s_pair[kPtfItems];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int t = r * kPtfThreads + threadIdx.x, k = base + t;
    int so = -1, sp = -1, pr = -1;
    if (k < N) {
      const int d = a.map_old[k];
      so = d;
      auto G = [&](const float* g, size_t idx) { return g ? g[idx] : 0.f; };
      if (!a.match[k]) {
#pragma unroll
        for (int e = 0; e < 3; e++) a.d_coords[3 * (size_t)k + e] = G(a.g_coords, 3 * (size_t)d + e);
        a.d_dens[k] = G(a.g_dens, d); a.d_wemb[k] = G(a.g_wemb, d); a.d_depth[k] = G(a.g_depth, d);
#pragma unroll
        for (int e = 0; e < 16; e++) a.d_ext[16 * (size_t)k + e] = G(a.g_ext, 16 * (size_t)d + e);
      } else {
        pr = d - a.n_keep;
        so = -2;                                        // fused row: the latent gradient goes to the GRU, not to feats[k]
        const int p = a.pix[k];
        const float w0 = a.dens[k], w1 = a.v_dens[p], ws = w0 + w1;
        const float r0 = w0 / ws, r1 = w1 / ws;
        float gw0 = G(a.g_dens, d), gw1 = gw0;
#pragma unroll
        for (int e = 0; e < 3; e++) {
          const float g = G(a.g_coords, 3 * (size_t)d + e);
          const float x0 = a.coords[3 * (size_t)k + e], x1 = a.v_coords[3 * (size_t)p + e];
          const float o = wmean(x0, w0, x1, w1, ws);
          a.d_coords[3 * (size_t)k + e] = g * r0;
          atomicAdd(a.dv_coords + 3 * (size_t)p + e, g * r1);
          gw0 += g * (x0 - o) / ws; gw1 += g * (x1 - o) / ws;
        }
        {
          const float g = G(a.g_depth, d);
          const float x0 = a.depth[k], x1 = a.v_depth[p];
          const float o = wmean(x0, w0, x1, w1, ws);
          a.d_depth[k] = g * r0;
          atomicAdd(a.dv_depth + p, g * r1);
          gw0 += g * (x0 - o) / ws; gw1 += g * (x1 - o) / ws;
        }
#pragma unroll
        for (int e = 0; e < 16; e++) {
          const float g = G(a.g_ext, 16 * (size_t)d + e);
          const float x0 = a.ext[16 * (size_t)k + e], x1 = a.v_ext[e];
          const float o = wmean(x0, w0, x1, w1, ws);
          a.d_ext[16 * (size_t)k + e] = g * r0;
          gw0 += g * (x0 - o) / ws; gw1 += g * (x1 - o) / ws;
        }
        a.d_dens[k] = gw0;
        atomicAdd(a.dv_dens + p, gw1);
        const float gw = G(a.g_wemb, d);
        a.d_wemb[k] = gw;
        atomicAdd(a.dv_wemb + p, gw);
      }
    }
    if (k < HW) {
      const int d = a.map_px[k];
      sp = d;
      if (d >= 0) {                                     // appended pixel: plain copies (never a partner of a fused row)
#pragma unroll
        for (int e = 0; e < 3; e++) a.dv_coords[3 * (size_t)k + e] = a.g_coords ? a.g_coords[3 * (size_t)d + e] : 0.f;
        a.dv_dens[k] = a.g_dens ? a.g_dens[d] : 0.f;
        a.dv_wemb[k] = a.g_wemb ? a.g_wemb[d] : 0.f;
        a.dv_depth[k] = a.g_depth ? a.g_depth[d] : 0.f;
      }
    }
    s_src_old[t] = so; s_src_px[t] = sp; s_pair[t] = pr;
  }
  __syncthreads();
  const int cpr = F >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(a.g_feats);
  float4* d4 = reinterpret_cast<float4*>(a.d_feats);
  float4* dv4 = reinterpret_cast<float4*>(a.dv_feats);
  float4* dg4 = reinterpret_cast<float4*>(a.d_gru);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int idx = threadIdx.x; idx < kPtfItems * cpr; idx += kPtfThreads) {
    const int t = idx / cpr, ch = idx - t * cpr;
    const int k = base + t;
    if (k < N) {
      const int so = s_src_old[t];
      const int row = so == -2 ? a.n_keep + s_pair[t] : so;
      const float4 g = g4 ? g4[(size_t)row * cpr + ch] : zero;
      if (so == -2) { dg4[(size_t)s_pair[t] * cpr + ch] = g; d4[(size_t)k * cpr + ch] = zero; }
      else d4[(size_t)k * cpr + ch] = g;
    }
    if (k < HW) {
      const int sp = s_src_px[t];
      dv4[(size_t)k * cpr + ch] = (sp >= 0 && g4) ? g4[(size_t)sp * cpr + ch] : zero;
    }
  }
}

int launch_ptf_merge_bwd(const FsPtfMergeBwdArgs& a, cudaStream_t s) {
  const int HW = a.H * a.W;
  int rc;
  // the view-side scalars of matched pixels accumulate with atomics (z-buffer ties: several globals per pixel)
  if ((rc = check_cuda(cudaMemsetAsync(a.dv_coords, 0, (size_t)HW * 3 * sizeof(float), s), "memset dv_coords"))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(a.dv_dens, 0, (size_t)HW * sizeof(float), s), "memset dv_dens"))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(a.dv_wemb, 0, (size_t)HW * sizeof(float), s), "memset dv_wemb"))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(a.dv_depth, 0, (size_t)HW * sizeof(float), s), "memset dv_depth"))) return rc;
  const int nb = ptf_blocks(a.N, HW);
  ptf_merge_bwd_kernel<<<nb, kPtfThreads, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "ptf_merge_bwd_kernel");
}

// =====================================================================================================================
// Append-only pool (inference fold).  ptf_compact_kernel rewrites the WHOLE state (344 B per Gaussian, read + write) at every
// fold step although a step only changes the matched rows and adds the unmatched pixels of view i.  Here the state lives in
// place: row r of the pool never moves; a fused Gaussian is updated where it is, new ones are appended behind the last row.
// The ORDER contract of the reference ([kept] ++ [fused] ++ [appended], per step) is carried by an index instead of by the
// data: `phys[k]` = pool row of logical position k.  A step's order update is a stable partition of that index by the step's
// match flags (N ints instead of 344 N bytes); it depends only on the flags, so the host runs it on a side stream under the
// following steps.  ONE gather at the end materialises the state in logical order.  Arithmetic per row is unchanged (wmean),
// so the result is bit-identical to the compacting fold.
// ---------------------------------------------------------------------------------------------------------------------
// in-place merge of the matched pairs (pair m: pool row j = pair_j[m], pixel p = pair_p[m]).  Grid-stride (M is only known on the
// device; a grid sized by its upper bound was ~300 waves of empty blocks per step): first one THREAD per pair for the scalar
// fields (independent loads across the threads; a warp-per-pair version with dependent scalar loads took 0.12 ms per step, as
// long as the compaction it replaces), then one thread per 16-byte chunk of the fused latents.
__global__ void __launch_bounds__(256) ptf_pool_fuse_kernel(FsPtfArgs a, float* __restrict__ feats, float* __restrict__ coords,
                                                            float* __restrict__ dens, float* __restrict__ wemb, float* __restrict__ ext,
                                                            float* __restrict__ depth) {
  const int M = a.counts_out[2];
  const int gtid = blockIdx.x * 256 + threadIdx.x, gsize = gridDim.x * 256;
  for (int m = gtid; m < M; m += gsize) {
    const int j = a.pair_j[m], p = a.pair_p[m];
    const float w0 = dens[j], w1 = a.v_dens[p], ws = w0 + w1;
    float c[3];
#pragma unroll
    for (int e = 0; e < 3; e++) c[e] = wmean(coords[3 * (size_t)j + e], w0, a.v_coords[3 * (size_t)p + e], w1, ws);
    const float z = wmean(depth[j], w0, a.v_depth[p], w1, ws);
    const float we = wemb[j] + a.v_wemb[p];
    float4* ej = reinterpret_cast<float4*>(ext + 16 * (size_t)j);
    const float4* ev = reinterpret_cast<const float4*>(a.v_ext);
    float4 E[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float4 g = ej[q], v = ev[q];
      E[q] = make_float4(wmean(g.x, w0, v.x, w1, ws), wmean(g.y, w0, v.y, w1, ws), wmean(g.z, w0, v.z, w1, ws), wmean(g.w, w0, v.w, w1, ws));
    }
#pragma unroll
    for (int e = 0; e < 3; e++) coords[3 * (size_t)j + e] = c[e];
    depth[j] = z; wemb[j] = we; dens[j] = ws;
#pragma unroll
    for (int q = 0; q < 4; q++) ej[q] = E[q];
  }
  const int cpr = a.F >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(a.gru_out);
  float4* f4 = reinterpret_cast<float4*>(feats);
  const long long total = (long long)M * cpr;
  for (long long idx = gtid; idx < total; idx += gsize) {
    const int m = (int)(idx / cpr), ch = (int)(idx - (long long)m * cpr);
    f4[(size_t)a.pair_j[m] * cpr + ch] = g4[idx];
  }
}

// unmatched pixels of view i -> rows N, N+1, ... (raster order); block = kPtfItems pixels
__global__ void __launch_bounds__(kPtfThreads) ptf_pool_append_kernel(FsPtfArgs a, float* __restrict__ feats, float* __restrict__ coords,
                                                                      float* __restrict__ dens, float* __restrict__ wemb, float* __restrict__ ext,
                                                                      float* __restrict__ depth) {
  constexpr int R = kPtfItems / kPtfThreads;
  __shared__ int s_warp[R * 8];
  __shared__ int s_dst[kPtfItems];
  const int N = a.counts_in[0], HW = a.H * a.W, F = a.F;
  const int base = blockIdx.x * kPtfItems;
  if (base >= HW) return;
  const int off_app = a.block_counts[3 * (size_t)blockIdx.x + 2];
  int ra[R];
  block_ranks([&](int r) { const int k = base + r * kPtfThreads + threadIdx.x; return k < HW && a.append[k]; }, ra, s_warp);
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int t = r * kPtfThreads + threadIdx.x, k = base + t;
    int d = -1;
    if (k < HW && a.append[k]) {
      d = N + off_app + ra[r];
      coords[3 * (size_t)d] = a.v_coords[3 * (size_t)k]; coords[3 * (size_t)d + 1] = a.v_coords[3 * (size_t)k + 1];
      coords[3 * (size_t)d + 2] = a.v_coords[3 * (size_t)k + 2];
      dens[d] = a.v_dens[k]; wemb[d] = a.v_wemb[k]; depth[d] = a.v_depth[k];
      const float4* se = reinterpret_cast<const float4*>(a.v_ext);
      float4* de = reinterpret_cast<float4*>(ext + 16 * (size_t)d);
      de[0] = se[0]; de[1] = se[1]; de[2] = se[2]; de[3] = se[3];
    }
    s_dst[t] = d;
  }
  __syncthreads();
  const int cpr = F >> 2;
  const float4* vf4 = reinterpret_cast<const float4*>(a.v_feats);
  float4* out4 = reinterpret_cast<float4*>(feats);
#pragma unroll 4
  for (int idx = threadIdx.x; idx < kPtfItems * cpr; idx += kPtfThreads) {
    const int t = idx / cpr, ch = idx - t * cpr;
    const int d = s_dst[t];
    if (d >= 0) out4[(size_t)d * cpr + ch] = vf4[(size_t)(base + t) * cpr + ch];
  }
}

// ---- order index: stable partition of phys[0..N) by the step's match flags, then the appended rows ----
// counts[0] = N (rows before the step), counts[1] = n_keep, counts[3] = n_append
__global__ void __launch_bounds__(kPtfThreads) ptf_order_count_kernel(const int* __restrict__ counts, const int* __restrict__ phys,
                                                                      const uint8_t* __restrict__ match, int* __restrict__ blk) {
  __shared__ int s_cnt;
  const int N = counts[0];
  const int base = blockIdx.x * kPtfItems;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int c = 0;
#pragma unroll
  for (int r = 0; r < kPtfItems / kPtfThreads; r++) {
    const int k = base + r * kPtfThreads + threadIdx.x;
    if (k < N) c += match[phys ? phys[k] : k] ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) blk[blockIdx.x] = s_cnt;
}

__global__ void __launch_bounds__(1024) ptf_order_scan_kernel(int* __restrict__ blk, int nb) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int k = base + tid;
    const int c = k < nb ? blk[k] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int endv = carry + (warp ? warp_sums[warp - 1] : 0) + incl;
    if (k < nb) blk[k] = endv - c;                          // exclusive offset of the block's matched rows
    __syncthreads();
    if (tid == 1023) carry = endv;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPtfThreads) ptf_order_scatter_kernel(const int* __restrict__ counts, const int* __restrict__ phys,
                                                                        const uint8_t* __restrict__ match, const int* __restrict__ blk,
                                                                        int* __restrict__ phys_out) {
  constexpr int R = kPtfItems / kPtfThreads;
  __shared__ int s_warp[R * 8];
  const int N = counts[0], n_keep = counts[1], n_app = counts[3];
  const int base = blockIdx.x * kPtfItems;
  if (base >= N + n_app) return;
  int row[R];
  bool mt[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int k = base + r * kPtfThreads + threadIdx.x;
    row[r] = k < N ? (phys ? phys[k] : k) : -1;
    mt[r] = k < N && match[row[r]];
  }
  int rm[R];
  block_ranks([&](int r) { return mt[r]; }, rm, s_warp);
  const int off_m = base < N ? blk[blockIdx.x] : 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int k = base + r * kPtfThreads + threadIdx.x;
    if (k < N) {
      // matched rows before position k (exclusive): off_m + rm[r]; kept rows before it: k - that
      const int before = off_m + rm[r];
      phys_out[mt[r] ? n_keep + before : k - before] = row[r];
    } else if (k < N + n_app) {
      phys_out[k] = k;                                      // appended rows sit in raster order behind the old ones, pool row = position
    }
  }
}

// final materialisation: logical position k <- pool row phys[k]
__global__ void __launch_bounds__(kPtfThreads) ptf_pool_gather_kernel(const int* __restrict__ n_dev, const int* __restrict__ phys, int F,
                                                                      const float* __restrict__ feats, const float* __restrict__ coords,
                                                                      const float* __restrict__ dens, const float* __restrict__ wemb,
                                                                      const float* __restrict__ ext, const float* __restrict__ depth,
                                                                      float* __restrict__ o_feats, float* __restrict__ o_coords,
                                                                      float* __restrict__ o_dens, float* __restrict__ o_wemb,
                                                                      float* __restrict__ o_ext, float* __restrict__ o_depth) {
  __shared__ int s_src[kPtfItems];
  const int N = n_dev[0];
  const int base = blockIdx.x * kPtfItems;
  if (base >= N) return;
#pragma unroll
  for (int r = 0; r < kPtfItems / kPtfThreads; r++) {
    const int t = r * kPtfThreads + threadIdx.x, k = base + t;
    int src = -1;
    if (k < N) {
      src = phys[k];
      o_coords[3 * (size_t)k] = coords[3 * (size_t)src]; o_coords[3 * (size_t)k + 1] = coords[3 * (size_t)src + 1];
      o_coords[3 * (size_t)k + 2] = coords[3 * (size_t)src + 2];
      o_dens[k] = dens[src]; o_wemb[k] = wemb[src]; o_depth[k] = depth[src];
      const float4* se = reinterpret_cast<const float4*>(ext + 16 * (size_t)src);
      float4* de = reinterpret_cast<float4*>(o_ext + 16 * (size_t)k);
      const float4 e0 = se[0], e1 = se[1], e2 = se[2], e3 = se[3];
      de[0] = e0; de[1] = e1; de[2] = e2; de[3] = e3;
    }
    s_src[t] = src;
  }
  __syncthreads();
  const int cpr = F >> 2;
  const float4* f4 = reinterpret_cast<const float4*>(feats);
  float4* out4 = reinterpret_cast<float4*>(o_feats);
#pragma unroll 4
  for (int idx = threadIdx.x; idx < kPtfItems * cpr; idx += kPtfThreads) {
    const int t = idx / cpr, ch = idx - t * cpr;
    const int src = s_src[t];
    if (src >= 0) out4[(size_t)(base + t) * cpr + ch] = f4[(size_t)src * cpr + ch];
  }
}

int launch_ptf_pool_update(const FsPtfArgs& a, cudaStream_t s) {
  // the pool IS the state: a.feats ... a.depth are updated in place (the const of the struct refers to the compacting fold)
  float* feats = const_cast<float*>(a.feats); float* coords = const_cast<float*>(a.coords); float* dens = const_cast<float*>(a.dens);
  float* wemb = const_cast<float*>(a.wemb); float* ext = const_cast<float*>(a.ext); float* depth = const_cast<float*>(a.depth);
  const int HW = a.H * a.W;
  if (a.n_upper > 0) {
    const int blocks = min((a.n_upper + 255) / 256, 148 * 8);
    ptf_pool_fuse_kernel<<<blocks, 256, 0, s>>>(a, feats, coords, dens, wemb, ext, depth);
    if (int rc = check_cuda(cudaGetLastError(), "ptf_pool_fuse_kernel")) return rc;
  }
  ptf_pool_append_kernel<<<(HW + kPtfItems - 1) / kPtfItems, kPtfThreads, 0, s>>>(a, feats, coords, dens, wemb, ext, depth);
  return check_cuda(cudaGetLastError(), "ptf_pool_append_kernel");
}

int launch_ptf_pool_order(int n_upper, const int* counts, const int* phys_in, const uint8_t* match, int* blk, int* phys_out, cudaStream_t s) {
  const int nb = (n_upper + kPtfItems - 1) / kPtfItems;
  if (nb <= 0) return FS_OK;
  ptf_order_count_kernel<<<nb, kPtfThreads, 0, s>>>(counts, phys_in, match, blk);
  if (int rc = check_cuda(cudaGetLastError(), "ptf_order_count_kernel")) return rc;
  ptf_order_scan_kernel<<<1, 1024, 0, s>>>(blk, nb);
  if (int rc = check_cuda(cudaGetLastError(), "ptf_order_scan_kernel")) return rc;
  ptf_order_scatter_kernel<<<nb, kPtfThreads, 0, s>>>(counts, phys_in, match, blk, phys_out);
  return check_cuda(cudaGetLastError(), "ptf_order_scatter_kernel");
}

int launch_ptf_pool_gather(int n_upper, const int* n_dev, const int* phys, int F, const float* feats, const float* coords, const float* dens,
                           const float* wemb, const float* ext, const float* depth, float* o_feats, float* o_coords, float* o_dens,
                           float* o_wemb, float* o_ext, float* o_depth, cudaStream_t s) {
  const int nb = (n_upper + kPtfItems - 1) / kPtfItems;
  if (nb <= 0) return FS_OK;
  ptf_pool_gather_kernel<<<nb, kPtfThreads, 0, s>>>(n_dev, phys, F, feats, coords, dens, wemb, ext, depth, o_feats, o_coords, o_dens, o_wemb,
                                                    o_ext, o_depth);
  return check_cuda(cudaGetLastError(), "ptf_pool_gather_kernel");
}

__global__ void ptf_fill_kernel(uint32_t* p, uint32_t v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
static int ptf_fill_u32(uint32_t* p, uint32_t v, int n, cudaStream_t s) {
  if (n <= 0) return FS_OK;
  ptf_fill_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, v, n);
  return check_cuda(cudaGetLastError(), "ptf_fill_kernel");
}

// ---- GRU glue (networks.py:201-214) around the cuBLAS GEMMs: gathers, positional encodings, concatenations and
// gates of the matched pairs in three launches instead of ~30 element-wise torch kernels per fused view ----
__device__ __forceinline__ void pe6(float x, float* o) {   // encoder_freesplat.py:62-77: (sin, cos)(x * 2^f), f = 0..5
  float s = x;
#pragma unroll
  for (int f = 0; f < 6; f++) { o[2 * f] = sinf(s); o[2 * f + 1] = cosf(s); s = s * 2.0f; }
}

// A1[m] = [hidden(F) | PE(v_dens[p], wemb[j]) (24) | input(F) | PE(dens[j], v_wemb[p]) (24)]     (concat_input of the GRU)
__global__ void __launch_bounds__(256) ptf_gru_inputs_kernel(int M, int F, const int* __restrict__ pair_j, const int* __restrict__ pair_p,
                                                             const float* __restrict__ feats, const float* __restrict__ dens,
                                                             const float* __restrict__ wemb, const float* __restrict__ v_feats,
                                                             const float* __restrict__ v_dens, const float* __restrict__ v_wemb,
                                                             float* __restrict__ A1) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);      // one warp per pair
  if (m >= M) return;
  const int j = pair_j[m], p = pair_p[m];
  const int ld = 2 * F + 48;
  float* row = A1 + (size_t)m * ld;
  for (int e = lane; e < F; e += 32) { row[e] = feats[(size_t)j * F + e]; row[F + 24 + e] = v_feats[(size_t)p * F + e]; }
  if (lane < 4) {
    // lane 0: PE(v_dens[p]) ; 1: PE(wemb[j])  -> e_h    ; 2: PE(dens[j]) ; 3: PE(v_wemb[p]) -> e_in
    const float x = lane == 0 ? v_dens[p] : lane == 1 ? wemb[j] : lane == 2 ? dens[j] : v_wemb[p];
    float o[12];
    pe6(x, o);
    float* dst = row + (lane < 2 ? F + 12 * lane : 2 * F + 24 + 12 * (lane - 2));
#pragma unroll
    for (int k = 0; k < 12; k++) dst[k] = o[k];
  }
}

// U[m] = [sigmoid(r_lin[m]) * hidden | input_feat_1]  (update_feat)
__global__ void __launch_bounds__(256) ptf_gru_update_kernel(int M, int F, const float* __restrict__ A1, const float* __restrict__ r_lin,
                                                             float* __restrict__ U) {
  const int ld1 = 2 * F + 48, ldu = 2 * F + 24;
  const size_t total = (size_t)M * ldu;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const size_t m = i / ldu; const int c = (int)(i - m * ldu);
  float v;
  if (c < F) v = A1[m * ld1 + c] * (1.0f / (1.0f + expf(-r_lin[m * F + c])));
  else v = A1[m * ld1 + F + 24 + (c - F)];
  U[i] = v;
}

// out = (1 - z) * hidden + z * tanh(q_lin),  z = sigmoid(z_lin)
__global__ void __launch_bounds__(256) ptf_gru_output_kernel(int M, int F, const float* __restrict__ A1, const float* __restrict__ z_lin,
                                                             const float* __restrict__ q_lin, float* __restrict__ out) {
  const size_t total = (size_t)M * F;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const size_t m = i / F; const int c = (int)(i - m * F);
  const float z = 1.0f / (1.0f + expf(-z_lin[i]));
  const float h = A1[m * (2 * F + 48) + c];
  out[i] = (1.0f - z) * h + z * tanhf(q_lin[i]);
}

// ---- backward of the three glue kernels (training: the Linear layers in between stay GEMMs of the caller) ----
// (1) out = (1-z) h + z q:   dz_lin = g (q - h) z (1 - z) ;  dq_lin = g z (1 - q^2) ;  dA1[:, :F] = g (1 - z)   (direct path to h)
__global__ void __launch_bounds__(256) ptf_gru_output_bwd_kernel(int M, int F, const float* __restrict__ A1, const float* __restrict__ z_lin,
                                                                 const float* __restrict__ q_lin, const float* __restrict__ g_out,
                                                                 float* __restrict__ dz_lin, float* __restrict__ dq_lin, float* __restrict__ dA1) {
  const size_t total = (size_t)M * F;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const size_t m = i / F; const int c = (int)(i - m * F);
  const int ld1 = 2 * F + 48;
  const float z = 1.0f / (1.0f + expf(-z_lin[i])), q = tanhf(q_lin[i]), h = A1[m * ld1 + c], g = g_out[i];
  dz_lin[i] = g * (q - h) * z * (1.0f - z);
  dq_lin[i] = g * z * (1.0f - q * q);
  dA1[m * ld1 + c] = g * (1.0f - z);
}

// (2) U = [sigmoid(r_lin) h | x | e_in]:  dr_lin = dU[:, :F] h r (1 - r) ;  dA1[:, :F] += dU[:, :F] r ;  dA1[:, F:F+24] = 0 (e_h enters
//     only through the first layers) ;  dA1[:, F+24:] = dU[:, F:]
__global__ void __launch_bounds__(256) ptf_gru_update_bwd_kernel(int M, int F, const float* __restrict__ A1, const float* __restrict__ r_lin,
                                                                 const float* __restrict__ dU, float* __restrict__ dr_lin, float* __restrict__ dA1) {
  const int ld1 = 2 * F + 48, ldu = 2 * F + 24;
  const size_t total = (size_t)M * ld1;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const size_t m = i / ld1; const int c = (int)(i - m * ld1);
  if (c < F) {
    const float r = 1.0f / (1.0f + expf(-r_lin[m * F + c])), h = A1[i], du = dU[m * ldu + c];
    dr_lin[m * F + c] = du * h * r * (1.0f - r);
    dA1[i] += du * r;
  } else if (c < F + 24) {
    dA1[i] = 0.f;
  } else {
    dA1[i] = dU[m * ldu + (c - 24)];
  }
}

// (3) gathers / positional encodings:  d feats[j] = dA1[:, :F] (every global is matched at most once: plain rows) ; view-side
//     rows and scalars accumulate with atomics (z-buffer ties: several globals per pixel) ;
//     PE(x) = (sin 2^f x, cos 2^f x)_f  ->  dx = sum_f 2^f (cos(2^f x) d_sin_f - sin(2^f x) d_cos_f)
__global__ void __launch_bounds__(256) ptf_gru_inputs_bwd_kernel(int M, int F, const int* __restrict__ pair_j, const int* __restrict__ pair_p,
                                                                 const float* __restrict__ dens, const float* __restrict__ wemb,
                                                                 const float* __restrict__ v_dens, const float* __restrict__ v_wemb,
                                                                 const float* __restrict__ dA1, float* __restrict__ d_feats,
                                                                 float* __restrict__ d_dens, float* __restrict__ d_wemb,
                                                                 float* __restrict__ dv_feats, float* __restrict__ dv_dens,
                                                                 float* __restrict__ dv_wemb) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);      // one warp per pair
  if (m >= M) return;
  const int j = pair_j[m], p = pair_p[m];
  const int ld = 2 * F + 48;
  const float* row = dA1 + (size_t)m * ld;
  for (int e = lane; e < F; e += 32) {
    d_feats[(size_t)j * F + e] = row[e];
    atomicAdd(dv_feats + (size_t)p * F + e, row[F + 24 + e]);
  }
  if (lane < 4) {
    const float x = lane == 0 ? v_dens[p] : lane == 1 ? wemb[j] : lane == 2 ? dens[j] : v_wemb[p];
    const float* g = row + (lane < 2 ? F + 12 * lane : 2 * F + 24 + 12 * (lane - 2));
    float acc = 0.f, s = x, sc = 1.0f;
#pragma unroll
    for (int f = 0; f < 6; f++) { acc += sc * (cosf(s) * g[2 * f] - sinf(s) * g[2 * f + 1]); s = s * 2.0f; sc = sc * 2.0f; }
    if (lane == 0) atomicAdd(dv_dens + p, acc);
    else if (lane == 1) d_wemb[j] = acc;
    else if (lane == 2) d_dens[j] = acc;
    else atomicAdd(dv_wemb + p, acc);
  }
}

int launch_ptf_gru_output_bwd(int M, int F, const float* A1, const float* z_lin, const float* q_lin, const float* g_out, float* dz_lin,
                              float* dq_lin, float* dA1, cudaStream_t s) {
  if (M <= 0) return FS_OK;
  const size_t total = (size_t)M * F;
  ptf_gru_output_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(M, F, A1, z_lin, q_lin, g_out, dz_lin, dq_lin, dA1);
  return check_cuda(cudaGetLastError(), "ptf_gru_output_bwd_kernel");
}
int launch_ptf_gru_update_bwd(int M, int F, const float* A1, const float* r_lin, const float* dU, float* dr_lin, float* dA1, cudaStream_t s) {
  if (M <= 0) return FS_OK;
  const size_t total = (size_t)M * (2 * F + 48);
  ptf_gru_update_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(M, F, A1, r_lin, dU, dr_lin, dA1);
  return check_cuda(cudaGetLastError(), "ptf_gru_update_bwd_kernel");
}
int launch_ptf_gru_inputs_bwd(int M, int F, const int* pj, const int* pp, const float* dens, const float* wemb, const float* v_dens,
                              const float* v_wemb, const float* dA1, float* d_feats, float* d_dens, float* d_wemb, float* dv_feats,
                              float* dv_dens, float* dv_wemb, cudaStream_t s) {
  if (M <= 0) return FS_OK;
  ptf_gru_inputs_bwd_kernel<<<(M + 7) / 8, 256, 0, s>>>(M, F, pj, pp, dens, wemb, v_dens, v_wemb, dA1, d_feats, d_dens, d_wemb, dv_feats,
                                                         dv_dens, dv_wemb);
  return check_cuda(cudaGetLastError(), "ptf_gru_inputs_bwd_kernel");
}

int launch_ptf_gru_inputs(int M, int F, const int* pj, const int* pp, const float* feats, const float* dens, const float* wemb,
                          const float* v_feats, const float* v_dens, const float* v_wemb, float* A1, cudaStream_t s) {
  if (M <= 0) return FS_OK;
  ptf_gru_inputs_kernel<<<(M + 7) / 8, 256, 0, s>>>(M, F, pj, pp, feats, dens, wemb, v_feats, v_dens, v_wemb, A1);
  return check_cuda(cudaGetLastError(), "ptf_gru_inputs_kernel");
}
int launch_ptf_gru_update(int M, int F, const float* A1, const float* r_lin, float* U, cudaStream_t s) {
  if (M <= 0) return FS_OK;
  const size_t total = (size_t)M * (2 * F + 24);
  ptf_gru_update_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(M, F, A1, r_lin, U);
  return check_cuda(cudaGetLastError(), "ptf_gru_update_kernel");
}
int launch_ptf_gru_output(int M, int F, const float* A1, const float* z_lin, const float* q_lin, float* out, cudaStream_t s) {
  if (M <= 0) return FS_OK;
  const size_t total = (size_t)M * F;
  ptf_gru_output_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(M, F, A1, z_lin, q_lin, out);
  return check_cuda(cudaGetLastError(), "ptf_gru_output_kernel");
}



// ---- per-view constants of the fold: E_i^-1 and the pixel-space intrinsics (encoder_freesplat.py:445-454) ----
// The reference calls extrinsic.inverse() (LAPACK on the CPU, cuSOLVER / MAGMA on the GPU: the two differ in the last
// bits).  The projected coordinates feed rounding and z-buffer decisions, so the inverse is part of the canonical
// arithmetic: cofactor expansion over 2x2 sub-determinants in fp64 (no fused multiply-add: -fmad=false), one rounding
// to fp32.  oracle/ptf.py::canonical_inverse is the same expression tree.
__global__ void ptf_view_setup_kernel(int V, int H, int W, const float* __restrict__ ext, const float* __restrict__ K,
                                      float* __restrict__ E_inv, float* __restrict__ K_px) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = (double)ext[16 * (size_t)v + k];
  const double s0 = a[0] * a[5] - a[4] * a[1], s1 = a[0] * a[6] - a[4] * a[2], s2 = a[0] * a[7] - a[4] * a[3];
  const double s3 = a[1] * a[6] - a[5] * a[2], s4 = a[1] * a[7] - a[5] * a[3], s5 = a[2] * a[7] - a[6] * a[3];
  const double c5 = a[10] * a[15] - a[14] * a[11], c4 = a[9] * a[15] - a[13] * a[11], c3 = a[9] * a[14] - a[13] * a[10];
  const double c2 = a[8] * a[15] - a[12] * a[11], c1 = a[8] * a[14] - a[12] * a[10], c0 = a[8] * a[13] - a[12] * a[9];
  const double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  const double id = 1.0 / det;
  double b[16];
  b[0] = (a[5] * c5 - a[6] * c4 + a[7] * c3) * id;
  b[1] = (-a[1] * c5 + a[2] * c4 - a[3] * c3) * id;
  b[2] = (a[13] * s5 - a[14] * s4 + a[15] * s3) * id;
  b[3] = (-a[9] * s5 + a[10] * s4 - a[11] * s3) * id;
  b[4] = (-a[4] * c5 + a[6] * c2 - a[7] * c1) * id;
  b[5] = (a[0] * c5 - a[2] * c2 + a[3] * c1) * id;
  b[6] = (-a[12] * s5 + a[14] * s2 - a[15] * s1) * id;
  b[7] = (a[8] * s5 - a[10] * s2 + a[11] * s1) * id;
  b[8] = (a[4] * c4 - a[5] * c2 + a[7] * c0) * id;
  b[9] = (-a[0] * c4 + a[1] * c2 - a[3] * c0) * id;
  b[10] = (a[12] * s4 - a[13] * s2 + a[15] * s0) * id;
  b[11] = (-a[8] * s4 + a[9] * s2 - a[11] * s0) * id;
  b[12] = (-a[4] * c3 + a[5] * c1 - a[6] * c0) * id;
  b[13] = (a[0] * c3 - a[1] * c1 + a[2] * c0) * id;
  b[14] = (-a[12] * s3 + a[13] * s1 - a[14] * s0) * id;
  b[15] = (a[8] * s3 - a[9] * s1 + a[10] * s0) * id;
#pragma unroll
  for (int k = 0; k < 16; k++) E_inv[16 * (size_t)v + k] = (float)b[k];
  const float fw = (float)W, fh = (float)H;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const float x = K[9 * (size_t)v + k];
    K_px[9 * (size_t)v + k] = k < 3 ? x * fw : (k < 6 ? x * fh : x);
  }
}

int launch_ptf_view_setup(int V, int H, int W, const float* ext, const float* K, float* E_inv, float* K_px, cudaStream_t s) {
  if (V <= 0) return FS_OK;
  ptf_view_setup_kernel<<<(V + 63) / 64, 64, 0, s>>>(V, H, W, ext, K, E_inv, K_px);
  return check_cuda(cudaGetLastError(), "ptf_view_setup_kernel");
}

int launch_ptf_match(const FsPtfArgs& a, cudaStream_t s) {
  const int HW = a.H * a.W;
  int rc;
  // z-buffer initialised to 1e4 = 0x461C4000 (encoder_freesplat.py:463)
  if ((rc = ptf_fill_u32(a.zbuf, 0x461C4000u, HW, s))) return rc;
  if (a.n_upper > 0) {
    ptf_project_kernel<<<(a.n_upper + kPtfThreads - 1) / kPtfThreads, kPtfThreads, 0, s>>>(a);
    if ((rc = check_cuda(cudaGetLastError(), "ptf_project_kernel"))) return rc;
  }
  const int nb = ptf_blocks(a.n_upper, HW);
  ptf_match_kernel<<<nb, kPtfThreads, 0, s>>>(a);
  if ((rc = check_cuda(cudaGetLastError(), "ptf_match_kernel"))) return rc;
  ptf_offsets_kernel<<<1, 1024, 0, s>>>(a, nb);
  if ((rc = check_cuda(cudaGetLastError(), "ptf_offsets_kernel"))) return rc;
  ptf_pairs_kernel<<<nb, kPtfThreads, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "ptf_pairs_kernel");
}

int launch_ptf_merge(const FsPtfArgs& a, cudaStream_t s) {
  const int nb = ptf_blocks(a.n_upper, a.H * a.W);
  ptf_compact_kernel<<<nb, kPtfThreads, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "ptf_compact_kernel");
}

// ============================================================================================================
// GRU of the matched pairs on the tensor cores (networks.py:188-214): out = (1-z) h + z tanh(n), with
//   [r,z] = sigmoid(MLP_{r,z}([h | e_h | x | e_in])),  n = MLP_n([r*h | x | e_in]),  MLP = Linear-ReLU-Linear.
// cuBLAS fp32 ran these skinny GEMMs (N = 64) at ~15 TFLOP/s and needed ~10 element-wise launches around them
// (1.5 ms per fused view at 0.26 M pairs, profiles/r1_bench_ops.json).  Here one CTA owns 128 pairs:
//   * thread = pair = TMEM lane; K is streamed in 16-column rounds, double-buffered in shared memory: the A operand
//     round is built by the threads themselves (gather of the latents, positional encodings, ReLU / gate outputs of the
//     previous layer read back from TMEM), the B operand round is a straight 16-byte copy of weights pre-split into
//     tf32 hi/lo and pre-arranged in the canonical K-major core-matrix layout by ptf_gru_prep_kernel;
//   * 3xTF32 (Ahi*Bhi + Ahi*Blo + Alo*Bhi), fp32 accumulation in TMEM: fp32-accurate (tests: <= 1e-4 vs the reference);
//   * layer 1 computes r and z hidden layers together (N = 128); TMEM columns: D1 [0,128) -> reused by D3 [0,64), D4 [64,128);
//     D2r [128,192), D2z [192,256).
namespace gru {
using namespace tc;

constexpr int kF = 64, kE = 24, kK1 = 2 * kF + 2 * kE /*176*/, kK3 = 2 * kF + kE /*152*/, kK3p = 160;
constexpr int kR1 = kK1 / kRound /*11*/, kR2 = kF / kRound /*4*/, kR3 = kK3p / kRound /*10*/, kR4 = kF / kRound /*4*/;
constexpr int kATile = 2 * (128 / 8) * 256;                  // one A round (128 rows x 16 cols): 8 KB per hi / lo

// ---- weight preparation: every B round as [hi tile | lo tile] in the canonical layout, rounds concatenated ----
// order: L1 (N=128: rows 0..63 = mlp_r[0], 64..127 = mlp_z[0]; 11 rounds) | L2r (N=64, 4) | L2z (4) | L3 (N=64, 10, K padded) | L4 (4)
constexpr size_t kOffL1 = 0, kOffL2r = kOffL1 + (size_t)kR1 * 2 * bTile(128), kOffL2z = kOffL2r + (size_t)kR2 * 2 * bTile(64),
                 kOffL3 = kOffL2z + (size_t)kR2 * 2 * bTile(64), kOffL4 = kOffL3 + (size_t)kR3 * 2 * bTile(64),
                 kWBytes = kOffL4 + (size_t)kR4 * 2 * bTile(64);

__global__ void ptf_gru_prep_kernel(const float* Wr0, const float* Wz0, const float* Wr2, const float* Wz2, const float* Wn0, const float* Wn2,
                                    unsigned char* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  auto put = [&](size_t base, int rows, int round, int row, int kk, float v) {
    const uint32_t hi = to_tf32(v), lo = to_tf32(v - __uint_as_float(hi));
    unsigned char* tile = out + base + (size_t)round * 2 * bTile(rows);
    *reinterpret_cast<uint32_t*>(tile + op_off(row, kk, rows)) = hi;
    *reinterpret_cast<uint32_t*>(tile + bTile(rows) + op_off(row, kk, rows)) = lo;
  };
  if (t < 128 * kK1) { const int n = t / kK1, k = t - n * kK1; put(kOffL1, 128, k / kRound, n, k % kRound, n < 64 ? Wr0[n * kK1 + k] : Wz0[(n - 64) * kK1 + k]); }
  if (t < 64 * kF) { const int n = t / kF, k = t - n * kF; put(kOffL2r, 64, k / kRound, n, k % kRound, Wr2[t]); put(kOffL2z, 64, k / kRound, n, k % kRound, Wz2[t]);
                     put(kOffL4, 64, k / kRound, n, k % kRound, Wn2[t]); }
  if (t < 64 * kK3p) { const int n = t / kK3p, k = t - n * kK3p; put(kOffL3, 64, k / kRound, n, k % kRound, k < kK3 ? Wn0[n * kK3 + k] : 0.f); }
}

// kT = row tiles of 128 pairs per CTA.  kT = 2 (256 pairs, all 512 TMEM columns, one CTA per SM): every weight round is copied
// from L2 once and multiplied into BOTH tiles -- half the L2 -> shared weight traffic (352 KB per 128 pairs at kT = 1, ~0.6 GB
// per fold step) and half the block barriers per pair.
template <int kT>
struct __align__(128) Smem {
  unsigned char A[kT][2][2][kATile];    // [tile][buffer][hi|lo]   (a second A tile for layer 2's z half lives in Az)
  unsigned char Az[kT][2][2][kATile];
  unsigned char B[2][2 * bTile(128)];   // [buffer][hi tile | lo tile]  (layer 2: r round then z round, 64 rows each)
  float b_r0[kF], b_z0[kF], b_r2[kF], b_z2[kF], b_n0[kF], b_n2[kF];
  unsigned long long bar_free[2], bar_done;
  uint32_t tmem_base;
};

// gates through MUFU.EX2 / MUFU.RCP (|rel err| ~ 1e-6 each, far inside the 1e-4 budget of the latents; the library expf +
// IEEE division + tanhf versions were ~35 % of the kernel's instructions)
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigm(float x) { return rcpa(1.0f + ex2a(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2a(-2.8853900817779268f * fminf(fmaxf(x, -20.f), 20.f));      // exp(-2x)
  return (1.0f - e) * rcpa(1.0f + e);
}

template <int kT>
__global__ void __launch_bounds__(128 * kT) ptf_gru_tc_kernel(int M_host, const int* __restrict__ M_dev, const int* __restrict__ pair_j, const int* __restrict__ pair_p,
                                                         const float* __restrict__ feats, const float* __restrict__ dens,
                                                         const float* __restrict__ wemb, const float* __restrict__ v_feats,
                                                         const float* __restrict__ v_dens, const float* __restrict__ v_wemb,
                                                         const unsigned char* __restrict__ W, const float* __restrict__ biases /*6 x 64*/,
                                                         float* __restrict__ out, float* __restrict__ save /*6 x [M_host,64] or NULL*/,
                                                         float* __restrict__ save_a1 /*[M,176] or NULL*/) {
  extern __shared__ __align__(128) unsigned char gru_smem_raw[];
  Smem<kT>& sm = *reinterpret_cast<Smem<kT>*>(gru_smem_raw);
  constexpr int kNT = 128 * kT;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile = tid >> 7, row = tid & 127;             // this thread's pair = row `row` of row tile `tile`
  const int M = M_dev ? min(*M_dev, M_host) : M_host;        // M_host is only the grid's upper bound when M_dev is given
  if ((int)blockIdx.x * kNT >= M) return;                  // CTA-uniform, before any barrier / TMEM allocation
  const int m = blockIdx.x * kNT + tid;
  // training: the six intermediate activations the backward needs, [Hr | Hz | r_lin | z_lin | Hn | q_lin] as six [M,64] matrices
  // (fs_ptf_gru_backward reads them instead of recomputing the six layers)
  auto keep = [&](int which, int col, float v0, float v1, float v2, float v3) {
    *reinterpret_cast<float4*>(save + ((size_t)which * M_host + m) * kF + col) = make_float4(v0, v1, v2, v3);
  };
  const bool active = m < M;
  if (tid < kF) { sm.b_r0[tid] = biases[tid]; sm.b_z0[tid] = biases[64 + tid]; sm.b_r2[tid] = biases[128 + tid]; sm.b_z2[tid] = biases[192 + tid];
                  sm.b_n0[tid] = biases[256 + tid]; sm.b_n2[tid] = biases[320 + tid]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_free[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_free[1])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_done)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(256u * kT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  // a warp reaches the 32 TMEM lanes of its quarter (warp % 4); tile t owns columns [256 t, 256 t + 256)
  const uint32_t t_row = tmem + 256u * (uint32_t)tile + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t bar_free[2] = {smem_u32(&sm.bar_free[0]), smem_u32(&sm.bar_free[1])};
  const uint32_t bar_done = smem_u32(&sm.bar_done);
  uint32_t ph_free[2] = {0u, 0u}, ph_done = 0u;
  int uses[2] = {0, 0};                                    // rounds issued on each buffer so far

  const int j = active ? pair_j[m] : 0, p = active ? pair_p[m] : 0;
  float h[kF];
  {
    const float4* hp = reinterpret_cast<const float4*>(feats + (size_t)j * kF);
#pragma unroll
    for (int q = 0; q < kF / 4; q++) { const float4 v = __ldg(hp + q); h[4 * q] = v.x; h[4 * q + 1] = v.y; h[4 * q + 2] = v.z; h[4 * q + 3] = v.w; }
  }
  // the view-i latent of the pair is consumed in rounds 5-9 of layer 1 and again in layer 3: fetched once, up front, next to h
  // (ncu r1k: the scalar loads at the point of use were the long-scoreboard stall of the A builds)
  float x[kF];
  {
    const float4* xp = reinterpret_cast<const float4*>(v_feats + (size_t)p * kF);
#pragma unroll
    for (int q = 0; q < kF / 4; q++) { const float4 v = __ldg(xp + q); x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w; }
  }
  const float pe_src[4] = {v_dens[p], wemb[j], dens[j], v_wemb[p]};     // e_h = PE(v_dens, wemb) ; e_in = PE(dens, v_wemb)

  // generic "run one round": wait until buffer free, let `fill` write the A tile(s), copy the B round, issue the MMAs
  auto begin_round = [&](int buf) {
    if (uses[buf] > 0) { mbar_wait(bar_free[buf], ph_free[buf]); ph_free[buf] ^= 1u; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  // B rounds travel global -> shared with cp.async (16 bytes, L2 only): the copy is issued BEFORE the threads build the A
  // round and completes underneath it (ncu r1f: the synchronous LDG -> STS copy was the long-scoreboard stall of this kernel)
  auto copy_b = [&](int buf, const unsigned char* src, int bytes) {
    const uint32_t d0 = smem_u32(sm.B[buf]);
    for (int k = tid; k < bytes / 16; k += kNT)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (uint32_t)k * 16u), "l"(src + (size_t)k * 16) : "memory");
  };
  auto publish = [&]() {
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  };
  // 24-wide positional encoding of (a, b): [sin(2^k a), cos(2^k a)]_{k<6} ++ the same for b (encoder_freesplat.py:62-77).
  // One sincosf per value, the five octaves by angle doubling (sin 2t = 2 sin t cos t, cos 2t = 1 - 2 sin^2 t): the 48
  // separate sinf / cosf calls with their range-reduction branches were ~20 % of this kernel's instructions (ncu r1f);
  // the doubling error grows to ~2^5 ulp(1) = 4e-6 absolute at the top octave, far inside the 1e-4 budget of the latents
  // (the encodings only feed the GRU, never an index decision).
  auto pe24_all = [&](float a, float b, float (&o)[kE]) {
    float sa, ca, sb, cb;
    sincosf(a, &sa, &ca); sincosf(b, &sb, &cb);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      o[2 * k] = sa; o[2 * k + 1] = ca; o[12 + 2 * k] = sb; o[12 + 2 * k + 1] = cb;
      const float sa2 = 2.0f * sa * ca, ca2 = fmaf(-2.0f * sa, sa, 1.0f), sb2 = 2.0f * sb * cb, cb2 = fmaf(-2.0f * sb, sb, 1.0f);
      sa = sa2; ca = ca2; sb = sb2; cb = cb2;
    }
  };
  float e_h[kE], e_in[kE];
  pe24_all(pe_src[0], pe_src[1], e_h);
  pe24_all(pe_src[2], pe_src[3], e_in);
  // column c of concat_input (176) / of update_feat's tail; r_gate only used for layer 3
  auto a1_col = [&](int c) -> float {
    if (c < kF) return h[c];
    if (c < kF + kE) return e_h[c - kF];
    if (c < 2 * kF + kE) return x[c - kF - kE];
    return e_in[c - 2 * kF - kE];
  };

  int round_no = 0;
  // (all round loops are fully unrolled: column indices, buffer numbers and mbarrier phases are compile-time constants,
  //  so h[] stays in registers)
  // ------------------------------------------------ layer 1: [128 x 176] x [176 x 128]
#pragma unroll
  for (int r = 0; r < kR1; r++, round_no++) {
    const int buf = round_no & 1;
    begin_round(buf);
    copy_b(buf, W + kOffL1 + (size_t)r * 2 * bTile(128), 2 * bTile(128));
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = r * kRound + 4 * q;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; e++) v[e] = a1_col(c + e);
      store_a4(sm.A[tile][buf][0], sm.A[tile][buf][1], row, 4 * q, v[0], v[1], v[2], v[3]);
      if (save_a1 && active) *reinterpret_cast<float4*>(save_a1 + (size_t)m * kK1 + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
    publish();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t bH = smem_u32(sm.B[buf]), bL = bH + bTile(128);
#pragma unroll
      for (int t = 0; t < kT; t++) {
        const uint32_t aH = smem_u32(sm.A[t][buf][0]), aL = smem_u32(sm.A[t][buf][1]), td = tmem + 256u * t;
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const uint32_t ao = s * (128 / 8) * 256, bo = s * (128 / 8) * 256;
          mma_tf32(td, make_desc(aL + ao), make_desc(bH + bo), idesc(128), (r > 0 || s > 0) ? 1u : 0u);
          mma_tf32(td, make_desc(aH + ao), make_desc(bL + bo), idesc(128), 1u);
          mma_tf32(td, make_desc(aH + ao), make_desc(bH + bo), idesc(128), 1u);
        }
      }
      commit(bar_free[buf]);
      if (r == kR1 - 1) commit(bar_done);
    }
    uses[buf]++;
  }
  mbar_wait(bar_done, ph_done); ph_done ^= 1u;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ------------------------------------------------ layer 2: relu(D1 + b) -> r_lin (D2r), z_lin (D2z)
#pragma unroll
  for (int r = 0; r < kR2; r++, round_no++) {
    const int buf = round_no & 1;
    begin_round(buf);
    {   // B: r round (hi|lo, 64 rows) then z round
      const unsigned char* s1 = W + kOffL2r + (size_t)r * 2 * bTile(64);
      const unsigned char* s2 = W + kOffL2z + (size_t)r * 2 * bTile(64);
      const uint32_t d0 = smem_u32(sm.B[buf]);
      const int n16 = 2 * bTile(64) / 16;
      for (int k = tid; k < n16; k += kNT) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (uint32_t)k * 16u), "l"(s1 + (size_t)k * 16) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (uint32_t)(n16 + k) * 16u), "l"(s2 + (size_t)k * 16) : "memory");
      }
    }
    float hr[16], hz[16];
    tmem_ld16(t_row + (uint32_t)(r * kRound), hr);
    tmem_ld16(t_row + (uint32_t)(64 + r * kRound), hz);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      float a[4], b[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        a[e] = fmaxf(hr[4 * q + e] + sm.b_r0[r * kRound + 4 * q + e], 0.f);
        b[e] = fmaxf(hz[4 * q + e] + sm.b_z0[r * kRound + 4 * q + e], 0.f);
      }
      store_a4(sm.A[tile][buf][0], sm.A[tile][buf][1], row, 4 * q, a[0], a[1], a[2], a[3]);
      store_a4(sm.Az[tile][buf][0], sm.Az[tile][buf][1], row, 4 * q, b[0], b[1], b[2], b[3]);
      if (save && active) { keep(0, r * kRound + 4 * q, a[0], a[1], a[2], a[3]); keep(1, r * kRound + 4 * q, b[0], b[1], b[2], b[3]); }
    }
    publish();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t bR = smem_u32(sm.B[buf]), bZ = bR + 2 * bTile(64);
#pragma unroll
      for (int t = 0; t < kT; t++) {
        const uint32_t arH = smem_u32(sm.A[t][buf][0]), arL = smem_u32(sm.A[t][buf][1]), azH = smem_u32(sm.Az[t][buf][0]),
                       azL = smem_u32(sm.Az[t][buf][1]), td = tmem + 256u * t;
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const uint32_t ao = s * (128 / 8) * 256, bo = s * (64 / 8) * 256;
          const uint32_t acc = (r > 0 || s > 0) ? 1u : 0u;
          mma_tf32(td + 128u, make_desc(arL + ao), make_desc(bR + bo), idesc(64), acc);
          mma_tf32(td + 128u, make_desc(arH + ao), make_desc(bR + bTile(64) + bo), idesc(64), 1u);
          mma_tf32(td + 128u, make_desc(arH + ao), make_desc(bR + bo), idesc(64), 1u);
          mma_tf32(td + 192u, make_desc(azL + ao), make_desc(bZ + bo), idesc(64), acc);
          mma_tf32(td + 192u, make_desc(azH + ao), make_desc(bZ + bTile(64) + bo), idesc(64), 1u);
          mma_tf32(td + 192u, make_desc(azH + ao), make_desc(bZ + bo), idesc(64), 1u);
        }
      }
      commit(bar_free[buf]);
      if (r == kR2 - 1) commit(bar_done);
    }
    uses[buf]++;
  }
  mbar_wait(bar_done, ph_done); ph_done ^= 1u;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ------------------------------------------------ layer 3: U = [sigmoid(r_lin + b) * h | x | e_in | 0] (160) x [160 x 64] -> D3 [0,64)
#pragma unroll
  for (int r = 0; r < kR3; r++, round_no++) {
    const int buf = round_no & 1;
    begin_round(buf);
    copy_b(buf, W + kOffL3 + (size_t)r * 2 * bTile(64), 2 * bTile(64));
    float rg[16];
    if (r < kR2) tmem_ld16(t_row + (uint32_t)(128 + r * kRound), rg);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int c = r * kRound + 4 * q + e;           // column of update_feat
        if (r < kR2) v[e] = sigm(rg[4 * q + e] + sm.b_r2[c]) * h[c];
        else if (c < 2 * kF) v[e] = x[c - kF];
        else if (c < kK3) v[e] = e_in[c - 2 * kF];
        else v[e] = 0.f;
      }
      store_a4(sm.A[tile][buf][0], sm.A[tile][buf][1], row, 4 * q, v[0], v[1], v[2], v[3]);
      if (r < kR2 && save && active) {
        const int c = r * kRound + 4 * q;
        keep(2, c, rg[4 * q] + sm.b_r2[c], rg[4 * q + 1] + sm.b_r2[c + 1], rg[4 * q + 2] + sm.b_r2[c + 2], rg[4 * q + 3] + sm.b_r2[c + 3]);
      }
    }
    publish();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t bH = smem_u32(sm.B[buf]), bL = bH + bTile(64);
#pragma unroll
      for (int t = 0; t < kT; t++) {
        const uint32_t aH = smem_u32(sm.A[t][buf][0]), aL = smem_u32(sm.A[t][buf][1]), td = tmem + 256u * t;
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const uint32_t ao = s * (128 / 8) * 256, bo = s * (64 / 8) * 256;
          mma_tf32(td, make_desc(aL + ao), make_desc(bH + bo), idesc(64), (r > 0 || s > 0) ? 1u : 0u);
          mma_tf32(td, make_desc(aH + ao), make_desc(bL + bo), idesc(64), 1u);
          mma_tf32(td, make_desc(aH + ao), make_desc(bH + bo), idesc(64), 1u);
        }
      }
      commit(bar_free[buf]);
      if (r == kR3 - 1) commit(bar_done);
    }
    uses[buf]++;
  }
  mbar_wait(bar_done, ph_done); ph_done ^= 1u;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ------------------------------------------------ layer 4: relu(D3 + b) x [64 x 64] -> D4 [64,128)
#pragma unroll
  for (int r = 0; r < kR4; r++, round_no++) {
    const int buf = round_no & 1;
    begin_round(buf);
    copy_b(buf, W + kOffL4 + (size_t)r * 2 * bTile(64), 2 * bTile(64));
    float hn[16];
    tmem_ld16(t_row + (uint32_t)(r * kRound), hn);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = r * kRound + 4 * q;
      const float v0 = fmaxf(hn[4 * q] + sm.b_n0[c], 0.f), v1 = fmaxf(hn[4 * q + 1] + sm.b_n0[c + 1], 0.f),
                  v2 = fmaxf(hn[4 * q + 2] + sm.b_n0[c + 2], 0.f), v3 = fmaxf(hn[4 * q + 3] + sm.b_n0[c + 3], 0.f);
      store_a4(sm.A[tile][buf][0], sm.A[tile][buf][1], row, 4 * q, v0, v1, v2, v3);
      if (save && active) keep(4, c, v0, v1, v2, v3);
    }
    publish();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t bH = smem_u32(sm.B[buf]), bL = bH + bTile(64);
#pragma unroll
      for (int t = 0; t < kT; t++) {
        const uint32_t aH = smem_u32(sm.A[t][buf][0]), aL = smem_u32(sm.A[t][buf][1]), td = tmem + 256u * t;
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const uint32_t ao = s * (128 / 8) * 256, bo = s * (64 / 8) * 256;
          mma_tf32(td + 64u, make_desc(aL + ao), make_desc(bH + bo), idesc(64), (r > 0 || s > 0) ? 1u : 0u);
          mma_tf32(td + 64u, make_desc(aH + ao), make_desc(bL + bo), idesc(64), 1u);
          mma_tf32(td + 64u, make_desc(aH + ao), make_desc(bH + bo), idesc(64), 1u);
        }
      }
      commit(bar_free[buf]);
      if (r == kR4 - 1) commit(bar_done);
    }
    uses[buf]++;
  }
  mbar_wait(bar_done, ph_done); ph_done ^= 1u;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ------------------------------------------------ gates
#pragma unroll
  for (int r = 0; r < kR2; r++) {
    float zl[16], ql[16];
    tmem_ld16(t_row + (uint32_t)(192 + r * kRound), zl);
    tmem_ld16(t_row + (uint32_t)(64 + r * kRound), ql);
    if (active) {
      float4* op = reinterpret_cast<float4*>(out + (size_t)m * kF + r * kRound);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int c = r * kRound + 4 * q + e;
          const float z = sigm(zl[4 * q + e] + sm.b_z2[c]);
          o[e] = (1.0f - z) * h[c] + z * tanh_fast(ql[4 * q + e] + sm.b_n2[c]);
        }
        op[q] = make_float4(o[0], o[1], o[2], o[3]);
        if (save) {
          const int c = r * kRound + 4 * q;
          keep(3, c, zl[4 * q] + sm.b_z2[c], zl[4 * q + 1] + sm.b_z2[c + 1], zl[4 * q + 2] + sm.b_z2[c + 2], zl[4 * q + 3] + sm.b_z2[c + 3]);
          keep(5, c, ql[4 * q] + sm.b_n2[c], ql[4 * q + 1] + sm.b_n2[c + 1], ql[4 * q + 2] + sm.b_n2[c + 2], ql[4 * q + 3] + sm.b_n2[c + 3]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u * kT) : "memory");
}

}  // namespace gru

int launch_ptf_gru_tc(const FsPtfGruArgs& a, cudaStream_t s) {
  if (a.M <= 0) return FS_OK;
  int rc;
  if (!(a.flags & 1)) {
    gru::ptf_gru_prep_kernel<<<(128 * gru::kK1 + 255) / 256, 256, 0, s>>>(a.W_r0, a.W_z0, a.W_r2, a.W_z2, a.W_n0, a.W_n2, a.wscratch);
    if ((rc = check_cuda(cudaGetLastError(), "ptf_gru_prep_kernel"))) return rc;
  }
  // 128 pairs per CTA, two CTAs per SM.  FS_GRU_TILES=2 selects the 256-pair variant (two row tiles share every weight round,
  // one CTA per SM): bit-identical, but 4 % SLOWER on B200 (4.55 vs 4.37 ms, 10-view fold) -- the kernel is bound by the per-thread
  // A builds, not by the L2 weight stream, and two independent CTAs overlap those better than one wide one.
  static int tiles = -1;
  if (tiles < 0) { const char* e = getenv("FS_GRU_TILES"); tiles = (e && e[0] == '2') ? 2 : 1; }
  if (tiles == 2) {
    const size_t smem = sizeof(gru::Smem<2>) + 128;
    if ((rc = check_cuda(cudaFuncSetAttribute(gru::ptf_gru_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "cudaFuncSetAttribute(ptf_gru_tc_kernel<2>)"))) return rc;
    gru::ptf_gru_tc_kernel<2><<<(a.M + 255) / 256, 256, smem, s>>>(a.M, a.M_dev, a.pair_j, a.pair_p, a.feats, a.dens, a.wemb, a.v_feats, a.v_dens,
                                                                    a.v_wemb, a.wscratch, a.biases, a.out, a.save, a.save_a1);
  } else {
    const size_t smem = sizeof(gru::Smem<1>) + 128;
    if ((rc = check_cuda(cudaFuncSetAttribute(gru::ptf_gru_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "cudaFuncSetAttribute(ptf_gru_tc_kernel<1>)"))) return rc;
    gru::ptf_gru_tc_kernel<1><<<(a.M + 127) / 128, 128, smem, s>>>(a.M, a.M_dev, a.pair_j, a.pair_p, a.feats, a.dens, a.wemb, a.v_feats, a.v_dens,
                                                                    a.v_wemb, a.wscratch, a.biases, a.out, a.save, a.save_a1);
  }
  return check_cuda(cudaGetLastError(), "ptf_gru_tc_kernel");
}

size_t ptf_gru_wscratch_bytes() { return gru::kWBytes; }

}  // namespace fs
