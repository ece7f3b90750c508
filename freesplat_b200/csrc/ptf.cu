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
#include "common.cuh"

namespace fs {

constexpr int kPtfThreads = 256;
constexpr int kPtfItems = 1024;   // items per block in the flag scans (4 per thread)

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

static inline int ptf_blocks(int n_upper, int HW) {
  const int m = n_upper > HW ? n_upper : HW;
  return (m + kPtfItems - 1) / kPtfItems;
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

}  // namespace fs
