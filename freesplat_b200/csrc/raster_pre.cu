// raster_pre.cu -- per-Gaussian stages of the rasterizer (sm_100a).
//   preprocess_kernel       : SURVEY §8a R1  (upstream forward.cu::preprocessCUDA) + per-tile counting
//   preprocess_bwd_kernel   : SURVEY §8a R8+R9 (upstream backward.cu::computeCov2DCUDA, preprocessCUDA)
//   mark_visible_kernel     : upstream rasterizer_impl.cu::checkFrustum
// Compiled with -fmad=false: the only fused operations are the explicit fmaf()
// calls of raster_math.cuh, which makes radius / tile-rect / depth-key decisions
// bit-identical to the CPU oracle (oracle/raster_oracle.c).
//
// HBM-bound streaming kernels: one thread per (view, Gaussian), 256-thread blocks,
// grid = ceil(P/256) x V.  Algorithmic bytes per Gaussian and view:
//   read 12 (mean) + 24 (cov6) + 4 (opacity) + 12*M (SH)  = 148 B at M=9
//   write 48 (record) + 24 (cov used) + 4 (radius) + 4 (tiles) + 1 (clamp) = 81 B
#include "common.cuh"
#include "raster_math.cuh"
#include "raster_scan.cuh"

namespace fs {

// Shared-memory layout of preprocess_kernel (dynamic):
//   float sh[256 * sh_stride]   SH rows of the block's 256 Gaussians, staged with coalesced loads
//                               (a per-thread walk over a 108-byte-strided row costs 27 L1 wavefronts
//                               per load instruction); sh_stride is odd => conflict-free reads.
// Each thread projects ITS Gaussian into all V views: the 148 input bytes are read once, not V times.
// DEG = active SH degree (compile time: the 3*(DEG+1)^2 coefficient loads become straight-line shared-memory
// reads with immediate offsets; the generic loop cost ~300 instructions of predicates / address math per view).
template <int DEG>
__global__ void __launch_bounds__(kThreads, 4) preprocess_kernel(FsRasterFwdArgs a, int gx, int gy, int sh_stride, int fused_scan) {
  extern __shared__ float4 smem4[];
  float* s_sh = reinterpret_cast<float*>(smem4);
  const int tid = threadIdx.x;
  const int block_base = blockIdx.x * kThreads;
  const int i = block_base + tid;
  const int nblk = min(kThreads, a.P - block_base);        // Gaussians handled by this block
  const int M3 = a.M * 3;
  if (a.shs) {
    const float* src = a.shs + (size_t)block_base * M3;
    const int total = nblk * M3;
    // (g, c) = divmod(k, M3) maintained incrementally: a runtime integer division per element
    // was the single largest cost of this kernel (ncu r1a: IABS/IMAD/ISETP chains).
    // The copies are asynchronous (cp.async, 4 bytes each: rows are 4*M3 bytes apart in HBM, sh_stride words apart in
    // shared memory): they complete while the threads load and project their Gaussians; the wait sits right before
    // the first colour evaluation (ncu r1g: the load->store dependency of the synchronous version was 15 % of all stalls).
    const int qstep = kThreads / M3, rstep = kThreads - qstep * M3;
    int g = tid / M3, c = tid - g * M3;
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_sh);
    for (int k = tid; k < total; k += kThreads) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_base + (uint32_t)(g * sh_stride + c) * 4u), "l"(src + k) : "memory");
      g += qstep; c += rstep;
      if (c >= M3) { c -= M3; g++; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const bool active = i < a.P;
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, opacity = 0.f;
  float c6in[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
    m0 = __ldg(a.means3D + 3 * (size_t)i + 0); m1 = __ldg(a.means3D + 3 * (size_t)i + 1); m2 = __ldg(a.means3D + 3 * (size_t)i + 2);
    opacity = __ldg(a.opacities + i);
    if (a.cov3D_precomp && a.cov_stride == 9) {
      // full row-major 3x3 matrices (the reference's Gaussians.covariances): upper triangle = cov6
      const float* cp = a.cov3D_precomp + 9 * (size_t)i;
      c6in[0] = __ldg(cp); c6in[1] = __ldg(cp + 1); c6in[2] = __ldg(cp + 2); c6in[3] = __ldg(cp + 4); c6in[4] = __ldg(cp + 5);
      c6in[5] = __ldg(cp + 8);
    } else if (a.cov3D_precomp) {
      const float2* cp = reinterpret_cast<const float2*>(a.cov3D_precomp + 6 * (size_t)i);
      const float2 c01 = __ldg(cp), c23 = __ldg(cp + 1), c45 = __ldg(cp + 2);
      c6in[0] = c01.x; c6in[1] = c01.y; c6in[2] = c23.x; c6in[3] = c23.y; c6in[4] = c45.x; c6in[5] = c45.y;
    } else {
      float sc[3] = {__ldg(a.scales + 3 * (size_t)i), __ldg(a.scales + 3 * (size_t)i + 1), __ldg(a.scales + 3 * (size_t)i + 2)};
      float q[4] = {__ldg(a.rotations + 4 * (size_t)i), __ldg(a.rotations + 4 * (size_t)i + 1),
                    __ldg(a.rotations + 4 * (size_t)i + 2), __ldg(a.rotations + 4 * (size_t)i + 3)};
      fsm::cov3d_from_scale_rot(sc, a.scale_modifier, q, c6in);
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();                                            // every thread's SH copies have landed
  const bool binned = use_bins(a);
  for (int v = 0; v < a.V; v++) {
    const size_t vi = (size_t)v * a.P + i;
    const float* __restrict__ view = a.views + (size_t)v * kViewFloats;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = make_float4(0.f, 0.f, -3.0e38f, -3.0e38f);
    int my_tiles = 0, rx0 = 0, ry0 = 0, rw = 1;
    if (active) {
      const float* __restrict__ proj = view + 16;
      const float tanx = view[38], tany = view[39], sscale = view[40];
      const float mean[3] = {m0 * sscale, m1 * sscale, m2 * sscale};
      const float s2 = sscale * sscale;
      float c6[6];
#pragma unroll
      for (int k = 0; k < 6; k++) c6[k] = c6in[k] * s2;
      const fsm::Projected pr = fsm::project_gaussian(mean, c6, view, proj, tanx, tany, a.H, a.W);
      if (a.cov3D) {
        float2* __restrict__ covo = reinterpret_cast<float2*>(a.cov3D + 6 * vi);
        covo[0] = make_float2(c6[0], c6[1]); covo[1] = make_float2(c6[2], c6[3]); covo[2] = make_float2(c6[4], c6[5]);
      }
      int clampmask = 0;
      uint32_t ntiles = 0;
      if (pr.radius > 0) {
        float rgb[3];
        if (a.colors_precomp) {
          rgb[0] = __ldg(a.colors_precomp + 3 * (size_t)i); rgb[1] = __ldg(a.colors_precomp + 3 * (size_t)i + 1);
          rgb[2] = __ldg(a.colors_precomp + 3 * (size_t)i + 2);
        } else {
          constexpr int NC = (DEG + 1) * (DEG + 1);                      // only the active coefficients are used
          float sh[NC * 3];                                              // sh[coef*3 + ch]
          const float* shp = s_sh + tid * sh_stride;
          if (a.sh_layout) {       // [P,3,M] (the reference's Gaussians.harmonics): channel-major rows
            const float* p0 = shp; const float* p1 = shp + a.M; const float* p2 = shp + 2 * a.M;
#pragma unroll
            for (int k = 0; k < NC; k++) { sh[3 * k] = p0[k]; sh[3 * k + 1] = p1[k]; sh[3 * k + 2] = p2[k]; }
          } else {
#pragma unroll
            for (int k = 0; k < NC * 3; k++) sh[k] = shp[k];
          }
          clampmask = fsm::sh_to_rgb(DEG, mean, view + 32, sh, rgb);
        }
        float hx, hy;
        fsm::alpha_extent(pr.con_x, pr.con_y, pr.con_z, opacity, &hx, &hy);
        if (hx < 0.f) { hx = -3.0e38f; hy = -3.0e38f; }
        r0 = make_float4(pr.px, pr.py, pr.con_x, pr.con_y);
        r1 = make_float4(pr.con_z, opacity, rgb[0], rgb[1]);
        r2 = make_float4(rgb[2], pr.depth, hx, hy);
        ntiles = (uint32_t)((pr.x1 - pr.x0) * (pr.y1 - pr.y0));
        rx0 = pr.x0; ry0 = pr.y0; rw = pr.x1 - pr.x0;
      }
      a.radii[vi] = pr.radius;
      if (a.tiles_touched) a.tiles_touched[vi] = ntiles;
      a.clamped[vi] = (uint8_t)clampmask;
      my_tiles = (int)ntiles;
    }
    // per-tile population count (tile ranges come from a scan over these counters).  Neighbouring
    // Gaussians mostly land in the same tile, so lanes are grouped by tile (match.any) and one lane
    // per group issues a single add: same-address atomics no longer serialise 32 deep.
    {
      uint32_t* __restrict__ cnt = a.tile_count + (size_t)v * gx * gy;
      const int maxn = __reduce_max_sync(0xffffffffu, my_tiles);
      int tx = rx0, ty = ry0;
      if (!binned) {
        for (int k = 0; k < maxn; k++) {
          const bool on = k < my_tiles;
          const int tile = on ? ty * gx + tx : -1 - (int)(threadIdx.x & 31);
          const unsigned grp = __match_any_sync(0xffffffffu, tile);
          if (on && (int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(cnt + tile, (uint32_t)__popc(grp));
          if (++tx == rx0 + rw) { tx = rx0; ty++; }
        }
      } else {
        // direct binning: the counting atomic returns the group's first slot in the tile's fixed-capacity bin and every lane of
        // the group stores its key there (same key as scatter_kernel: depth bits | Gaussian index)
        const int lane = (int)(threadIdx.x & 31);
        const unsigned long long key = ((unsigned long long)__float_as_uint(r2.y) << 32) | (unsigned long long)(uint32_t)i;
        unsigned long long* __restrict__ vb = reinterpret_cast<unsigned long long*>(a.bins) + (size_t)v * gx * gy * (size_t)a.bin_cap;
        // up to four tiles per pass: their slot atomics are in flight together and are consumed afterwards (one L2 round trip per
        // pass instead of one per tile); passes beyond the warp's maximum are skipped by warp-uniform branches
        const unsigned lt_mask = (1u << lane) - 1u;
        const uint32_t cap = (uint32_t)a.bin_cap;
        for (int k0 = 0; k0 < maxn; k0 += 4) {
          unsigned grp[4];
          uint32_t base[4];
          int tl[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            grp[u] = 0u; base[u] = 0u; tl[u] = -1;
            if (k0 + u < maxn) {                                     // warp-uniform
              const bool on = k0 + u < my_tiles;
              tl[u] = on ? ty * gx + tx : -1 - lane;
              grp[u] = __match_any_sync(0xffffffffu, tl[u]);
              if (on && lane == __ffs(grp[u]) - 1) base[u] = atomicAdd(cnt + tl[u], (uint32_t)__popc(grp[u]));
              if (++tx == rx0 + rw) { tx = rx0; ty++; }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (k0 + u < maxn) {                                     // warp-uniform
              const uint32_t b = __shfl_sync(0xffffffffu, base[u], __ffs(grp[u]) - 1);
              if (tl[u] >= 0) {
                const uint32_t slot = b + (uint32_t)__popc(grp[u] & lt_mask);
                if (slot < cap) vb[(size_t)tl[u] * (size_t)cap + slot] = key;
                else a.tile_cursor[(size_t)a.V * gx * gy] = 1u;      // a bin is full: this call falls back to scan + scatter
              }
            }
          }
        }
      }
    }
    // records: three 16-byte stores per thread (48-byte stride across the warp; the L2 merges the partial sectors).
    // Staging them through shared memory for fully coalesced stores cost two block barriers per view and was slower
    // (51.0 -> 46.4 us, measured).
    if (active) {
      float4* __restrict__ dst = reinterpret_cast<float4*>(a.rec) + 3 * vi;
      dst[0] = r0; dst[1] = r1; dst[2] = r2;
    }
  }
  // ---- the LAST CTA to finish turns the tile counters into ranges (R2 + R5): no separate one-block launch.  Every CTA makes its
  //      counter updates visible (fence), takes a ticket from status[3] (zeroed with the counters by the call's memset); the CTA
  //      that draws the last ticket scans (ld.global.cg reads) and resets the ticket. ----
  if (fused_scan) {
    __shared__ int s_last;
    __syncthreads();                       // the counter updates of all warps of this CTA happen-before thread 0's fence
    if (tid == 0) {
      __threadfence();                     // cumulative: orders them before the ticket (one fence per CTA: a fence in every thread
                                           // made each CTA wait for its own record stores to drain, +15 us on the kernel)
      s_last = (atomicAdd(a.status + 3, 1u) == gridDim.x - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      tile_scan_block<kThreads>(a.tile_count, a.ranges, a.status, a.V * gx * gy, (long long)a.capacity);
    }
  }
}

__global__ void __launch_bounds__(kThreads) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                                const float* __restrict__ view,
                                                                uint8_t* __restrict__ vis) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= P) return;
  const float s = view[40];
  const float z = fsm::dot4row(view, 2, means3D[3 * (size_t)i] * s, means3D[3 * (size_t)i + 1] * s,
                               means3D[3 * (size_t)i + 2] * s);
  vis[i] = z > 0.2f ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Backward of the per-Gaussian stages.  One thread per Gaussian loops over the V views, so the
// sums over views are deterministic and need no atomics.
// SH rows (values in, gradients out) go through shared memory: the block moves its 256 rows with coalesced accesses
// (cp.async in, float runs out) and every thread accumulates the gradient of ITS row over the views in shared memory; the
// first version read-modify-wrote the 108-byte-strided global rows once per view (ncu r1f: 0.12 of the HBM peak,
// long-scoreboard bound).
__global__ void __launch_bounds__(kThreads) preprocess_bwd_kernel(FsRasterBwdArgs a, int sh_stride) {
  extern __shared__ float s_bwd[];
  float* s_sh = s_bwd;                                     // [256][sh_stride] SH coefficients of the block's Gaussians
  float* s_gsh = s_bwd + kThreads * sh_stride;             // [256][sh_stride] their gradients
  const int tid = threadIdx.x;
  const int block_base = blockIdx.x * kThreads;
  const int i = block_base + tid;
  const int nblk = min(kThreads, a.P - block_base);
  const int M3 = a.M * 3;
  const bool use_sh = a.shs && a.dL_dshs;
  // fused reduce-scatter, vector path: full block, one owner rank, in-place layouts, staging large enough (13 floats per row)
  const bool vec_peer = a.peer_delta != nullptr && use_sh && nblk == kThreads && (a.shard_rows % kThreads) == 0 && a.cov_stride == 9 &&
                        a.dL_dcov3D != nullptr && sh_stride >= 13 && ((nblk * M3) & 3) == 0;
  if (use_sh) {
    const float* src = a.shs + (size_t)block_base * M3;
    const int total = nblk * M3;
    const int qstep = kThreads / M3, rstep = kThreads - qstep * M3;
    int g = tid / M3, c = tid - g * M3;
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_sh);
    for (int k = tid; k < total; k += kThreads) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_base + (uint32_t)(g * sh_stride + c) * 4u), "l"(src + k) : "memory");
      g += qstep; c += rstep;
      if (c >= M3) { c -= M3; g++; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    float* mine = s_gsh + tid * sh_stride;
    for (int k = 0; k < M3; k++) mine[k] = 0.f;
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  if (i < a.P) {
  const int H = a.H, W = a.W;
  float g_mean[3] = {0.f, 0.f, 0.f}, g_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, g_op = 0.f, g_col[3] = {0.f, 0.f, 0.f};
  const float m0 = a.means3D[3 * (size_t)i], m1 = a.means3D[3 * (size_t)i + 1], m2 = a.means3D[3 * (size_t)i + 2];
  float c6in[6];
  if (a.cov3D_precomp && a.cov_stride == 9) {
    const float* cp = a.cov3D_precomp + 9 * (size_t)i;
    c6in[0] = cp[0]; c6in[1] = cp[1]; c6in[2] = cp[2]; c6in[3] = cp[4]; c6in[4] = cp[5]; c6in[5] = cp[8];
  } else if (a.cov3D_precomp) {
#pragma unroll
    for (int k = 0; k < 6; k++) c6in[k] = a.cov3D_precomp[6 * (size_t)i + k];
  } else {
    fsm::cov3d_from_scale_rot(a.scales + 3 * (size_t)i, a.scale_modifier, a.rotations + 4 * (size_t)i, c6in);
  }

  for (int v = 0; v < a.V; v++) {
    const size_t vi = (size_t)v * a.P + i;
    float* g2d = a.dL_dmeans2D + 3 * vi;
    if (!(a.radii[vi] > 0)) { g2d[0] = 0.f; g2d[1] = 0.f; g2d[2] = 0.f; continue; }
    const float* __restrict__ view = a.views + (size_t)v * kViewFloats;
    const float* __restrict__ pm = view + 16;
    const float tanx = view[38], tany = view[39], ss = view[40];
    const float fx = (float)W / (2.0f * tanx), fy = (float)H / (2.0f * tany);
    const float4* gs = reinterpret_cast<const float4*>(a.dL_dscreen + 12 * vi);
    const float4 ga = gs[0], gb = gs[1], gc = gs[2];
    // ga = (mean2D.x, mean2D.y, conic.x, conic.y)  gb = (conic.w, opacity, r, g)  gc = (b, depth, -, -)
    g2d[0] = ga.x; g2d[1] = ga.y; g2d[2] = 0.f;
    g_op += gb.y;
    const float mean[3] = {m0 * ss, m1 * ss, m2 * ss};
    float c6[6];
    {
      const float s2c = ss * ss;
#pragma unroll
      for (int k = 0; k < 6; k++) c6[k] = c6in[k] * s2c;     // same arithmetic as the forward pass
    }
    const float pvx = fsm::dot4row(view, 0, mean[0], mean[1], mean[2]);
    const float pvy = fsm::dot4row(view, 1, mean[0], mean[1], mean[2]);
    const float pvz = fsm::dot4row(view, 2, mean[0], mean[1], mean[2]);
    const fsm::Cov2D cv = fsm::cov2d(pvx, pvy, pvz, fx, fy, tanx, tany, c6, view);
    const float xg = cv.clampx ? 0.f : 1.f, yg = cv.clampy ? 0.f : 1.f;
    const float ca = cv.a, cb = cv.b, cc = cv.c;
    const float gxx = ga.z, gxy = ga.w, gyy = gb.x;
    const float denom = ca * cc - cb * cb;
    const float d2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    const float* A3 = cv.Ta; const float* B3 = cv.Tb;
    float gcv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (d2inv != 0.f) {
      dL_da = d2inv * (-cc * cc * gxx + 2.f * cb * cc * gxy + (denom - ca * cc) * gyy);
      dL_dc = d2inv * (-ca * ca * gyy + 2.f * ca * cb * gxy + (denom - ca * cc) * gxx);
      dL_db = d2inv * 2.f * (cb * cc * gxx - (denom + 2.f * cb * cb) * gxy + ca * cb * gyy);
      gcv[0] = A3[0] * A3[0] * dL_da + A3[0] * B3[0] * dL_db + B3[0] * B3[0] * dL_dc;
      gcv[3] = A3[1] * A3[1] * dL_da + A3[1] * B3[1] * dL_db + B3[1] * B3[1] * dL_dc;
      gcv[5] = A3[2] * A3[2] * dL_da + A3[2] * B3[2] * dL_db + B3[2] * B3[2] * dL_dc;
      gcv[1] = 2.f * A3[0] * A3[1] * dL_da + (A3[0] * B3[1] + A3[1] * B3[0]) * dL_db + 2.f * B3[0] * B3[1] * dL_dc;
      gcv[2] = 2.f * A3[0] * A3[2] * dL_da + (A3[0] * B3[2] + A3[2] * B3[0]) * dL_db + 2.f * B3[0] * B3[2] * dL_dc;
      gcv[4] = 2.f * A3[2] * A3[1] * dL_da + (A3[1] * B3[2] + A3[2] * B3[1]) * dL_db + 2.f * B3[1] * B3[2] * dL_dc;
    }
    const float s2 = ss * ss;
#pragma unroll
    for (int k = 0; k < 6; k++) g_cov[k] += gcv[k] * s2;
    const float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
    float Sa[3], Sb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      Sa[k] = S[3 * k] * A3[0] + S[3 * k + 1] * A3[1] + S[3 * k + 2] * A3[2];
      Sb[k] = S[3 * k] * B3[0] + S[3 * k + 1] * B3[1] + S[3 * k + 2] * B3[2];
    }
    float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float dTa = 2.f * Sa[k] * dL_da + Sb[k] * dL_db;
      const float dTb = 2.f * Sb[k] * dL_dc + Sa[k] * dL_db;
      const float R0 = view[k * 4 + 0], R1 = view[k * 4 + 1], R2 = view[k * 4 + 2];
      dJ00 += R0 * dTa; dJ02 += R2 * dTa; dJ11 += R1 * dTb; dJ12 += R2 * dTb;
    }
    const float tz = 1.f / cv.tz, tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = xg * -fx * tz2 * dJ02;
    const float dty = yg * -fy * tz2 * dJ12;
    const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * cv.tx) * tz3 * dJ02 + (2.f * fy * cv.ty) * tz3 * dJ12;
    float dm[3];
#pragma unroll
    for (int k = 0; k < 3; k++) dm[k] = view[4 * k] * dtx + view[4 * k + 1] * dty + view[4 * k + 2] * dtz;
    if (a.has_depth_grad) {
#pragma unroll
      for (int k = 0; k < 3; k++) dm[k] += view[4 * k + 2] * gc.y;
    }
    // projection term (screen-space mean gradient)
    const float phx = fsm::dot4row(pm, 0, mean[0], mean[1], mean[2]);
    const float phy = fsm::dot4row(pm, 1, mean[0], mean[1], mean[2]);
    const float phw = fsm::dot4row(pm, 3, mean[0], mean[1], mean[2]);
    const float mw = 1.0f / (phw + 0.0000001f);
    const float mul1 = phx * mw * mw, mul2 = phy * mw * mw;
    dm[0] += (pm[0] * mw - pm[3] * mul1) * ga.x + (pm[1] * mw - pm[3] * mul2) * ga.y;
    dm[1] += (pm[4] * mw - pm[7] * mul1) * ga.x + (pm[5] * mw - pm[7] * mul2) * ga.y;
    dm[2] += (pm[8] * mw - pm[11] * mul1) * ga.x + (pm[9] * mw - pm[11] * mul2) * ga.y;

    const float grgb[3] = {gb.z, gb.w, gc.x};
    if (a.colors_precomp) {
      g_col[0] += grgb[0]; g_col[1] += grgb[1]; g_col[2] += grgb[2];
    } else if (use_sh) {
      const float* campos = view + 32;
      const float dox = mean[0] - campos[0], doy = mean[1] - campos[1], doz = mean[2] - campos[2];
      const float sum2 = dox * dox + doy * doy + doz * doz;
      const float len = sqrtf(sum2);
      const float x = dox / len, y = doy / len, z = doz / len;
      const float* sh = s_sh + tid * sh_stride;
      float* gsh = s_gsh + tid * sh_stride;
      const uint8_t cm = a.clamped[vi];
      const int D = a.sh_degree;
      float ddx = 0.f, ddy = 0.f, ddz = 0.f;
      float basis[16];
      basis[0] = fsm::kShC0;
      if (D > 0) {
        basis[1] = -fsm::kShC1 * y; basis[2] = fsm::kShC1 * z; basis[3] = -fsm::kShC1 * x;
        if (D > 1) {
          const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
          basis[4] = fsm::kShC2_0 * xy; basis[5] = fsm::kShC2_1 * yz; basis[6] = fsm::kShC2_2 * (2.f * zz - xx - yy);
          basis[7] = fsm::kShC2_3 * xz; basis[8] = fsm::kShC2_4 * (xx - yy);
          if (D > 2) {
            basis[9] = fsm::kShC3_0 * y * (3.f * xx - yy); basis[10] = fsm::kShC3_1 * xy * z;
            basis[11] = fsm::kShC3_2 * y * (4.f * zz - xx - yy); basis[12] = fsm::kShC3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
            basis[13] = fsm::kShC3_4 * x * (4.f * zz - xx - yy); basis[14] = fsm::kShC3_5 * z * (xx - yy);
            basis[15] = fsm::kShC3_6 * x * (xx - 3.f * yy);
          }
        }
      }
      const int nb = (D + 1) * (D + 1);
      for (int ch = 0; ch < 3; ch++) {
        const float g = ((cm >> ch) & 1) ? 0.f : grgb[ch];
        for (int k = 0; k < nb; k++) {
          float* dst = a.sh_layout ? gsh + ch * a.M + k : gsh + k * 3 + ch;
          *dst += basis[k] * g;
        }
        float rx = 0.f, ry = 0.f, rz = 0.f;
#define SHV(k) (a.sh_layout ? sh[ch * a.M + (k)] : sh[(k) * 3 + ch])
        if (D > 0) {
          rx = -fsm::kShC1 * SHV(3); ry = -fsm::kShC1 * SHV(1); rz = fsm::kShC1 * SHV(2);
          if (D > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            rx += fsm::kShC2_0 * y * SHV(4) + fsm::kShC2_2 * 2.f * -x * SHV(6) + fsm::kShC2_3 * z * SHV(7) + fsm::kShC2_4 * 2.f * x * SHV(8);
            ry += fsm::kShC2_0 * x * SHV(4) + fsm::kShC2_1 * z * SHV(5) + fsm::kShC2_2 * 2.f * -y * SHV(6) + fsm::kShC2_4 * 2.f * -y * SHV(8);
            rz += fsm::kShC2_1 * y * SHV(5) + fsm::kShC2_2 * 2.f * 2.f * z * SHV(6) + fsm::kShC2_3 * x * SHV(7);
            if (D > 2) {
              rx += fsm::kShC3_0 * SHV(9) * 3.f * 2.f * xy + fsm::kShC3_1 * SHV(10) * yz + fsm::kShC3_2 * SHV(11) * -2.f * xy +
                    fsm::kShC3_3 * SHV(12) * -3.f * 2.f * xz + fsm::kShC3_4 * SHV(13) * (-3.f * xx + 4.f * zz - yy) +
                    fsm::kShC3_5 * SHV(14) * 2.f * xz + fsm::kShC3_6 * SHV(15) * 3.f * (xx - yy);
              ry += fsm::kShC3_0 * SHV(9) * 3.f * (xx - yy) + fsm::kShC3_1 * SHV(10) * xz + fsm::kShC3_2 * SHV(11) * (-3.f * yy + 4.f * zz - xx) +
                    fsm::kShC3_3 * SHV(12) * -3.f * 2.f * yz + fsm::kShC3_4 * SHV(13) * -2.f * xy + fsm::kShC3_5 * SHV(14) * -2.f * yz +
                    fsm::kShC3_6 * SHV(15) * -3.f * 2.f * xy;
              rz += fsm::kShC3_1 * SHV(10) * xy + fsm::kShC3_2 * SHV(11) * 4.f * 2.f * yz + fsm::kShC3_3 * SHV(12) * 3.f * (2.f * zz - xx - yy) +
                    fsm::kShC3_4 * SHV(13) * 4.f * 2.f * xz + fsm::kShC3_5 * SHV(14) * (xx - yy);
            }
          }
        }
#undef SHV
        ddx += rx * g; ddy += ry * g; ddz += rz * g;
      }
      const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dm[0] += ((sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * inv32;
      dm[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * inv32;
      dm[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * inv32;
    }
    g_mean[0] += dm[0] * ss; g_mean[1] += dm[1] * ss; g_mean[2] += dm[2] * ss;
  }
  if (a.peer_delta != nullptr && vec_peer) {
    // fused reduce-scatter, vector path: the block's rows are parked in shared memory (the SH input staging is dead by now)
    // and leave below as coalesced 16-byte reductions
    __syncthreads();          // every thread is past its last read of the SH staging (vec_peer: the block is full, so all 256 get here)
    float* st = s_sh;
    st[3 * tid] = g_mean[0]; st[3 * tid + 1] = g_mean[1]; st[3 * tid + 2] = g_mean[2];
    float* sc = st + 3 * kThreads + 9 * tid;
    sc[0] = g_cov[0]; sc[1] = g_cov[1]; sc[2] = g_cov[2]; sc[3] = 0.f; sc[4] = g_cov[3]; sc[5] = g_cov[4]; sc[6] = 0.f; sc[7] = 0.f; sc[8] = g_cov[5];
    st[12 * kThreads + tid] = g_op;
  } else if (a.peer_delta != nullptr) {
    // fused reduce-scatter: this rank's partial sums go straight into the owner rank's buffers over NVLink (fire-and-forget
    // reductions on peer-mapped memory; the buffers were zeroed on every rank and a cross-rank barrier follows the kernel)
    const long long delta = a.peer_delta[min(i / a.shard_rows, a.world - 1)];
    auto peer = [&](float* p) { return reinterpret_cast<float*>(reinterpret_cast<char*>(p) + delta); };
    float* gm = peer(a.dL_dmeans3D + 3 * (size_t)i);
    atomicAdd(gm, g_mean[0]); atomicAdd(gm + 1, g_mean[1]); atomicAdd(gm + 2, g_mean[2]);
    atomicAdd(peer(a.dL_dopacities + i), g_op);
    if (a.dL_dcov3D && a.cov_stride == 9) {
      float* g = peer(a.dL_dcov3D + 9 * (size_t)i);
      atomicAdd(g, g_cov[0]); atomicAdd(g + 1, g_cov[1]); atomicAdd(g + 2, g_cov[2]); atomicAdd(g + 4, g_cov[3]); atomicAdd(g + 5, g_cov[4]);
      atomicAdd(g + 8, g_cov[5]);
    } else if (a.dL_dcov3D) {
      float* g = peer(a.dL_dcov3D + 6 * (size_t)i);
#pragma unroll
      for (int k = 0; k < 6; k++) atomicAdd(g + k, g_cov[k]);
    }
  } else {
  a.dL_dmeans3D[3 * (size_t)i] = g_mean[0]; a.dL_dmeans3D[3 * (size_t)i + 1] = g_mean[1]; a.dL_dmeans3D[3 * (size_t)i + 2] = g_mean[2];
  a.dL_dopacities[i] = g_op;
  if (a.dL_dcov3D && a.cov_stride == 9) {      // gradient of the upper-triangle gather: lower triangle gets 0
    float* g = a.dL_dcov3D + 9 * (size_t)i;
    g[0] = g_cov[0]; g[1] = g_cov[1]; g[2] = g_cov[2]; g[3] = 0.f; g[4] = g_cov[3]; g[5] = g_cov[4]; g[6] = 0.f; g[7] = 0.f; g[8] = g_cov[5];
  } else if (a.dL_dcov3D) {
#pragma unroll
    for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * (size_t)i + k] = g_cov[k];
  }
  }
  if (a.dL_dcolors) { a.dL_dcolors[3 * (size_t)i] = g_col[0]; a.dL_dcolors[3 * (size_t)i + 1] = g_col[1]; a.dL_dcolors[3 * (size_t)i + 2] = g_col[2]; }
  if (a.scales && a.rotations && a.dL_dscales && a.dL_drotations) {
    const float* q = a.rotations + 4 * (size_t)i;
    const float* s = a.scales + 3 * (size_t)i;
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float Rm[9] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                         2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                         2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
    const float mod = a.scale_modifier;
    const float sv[3] = {mod * s[0], mod * s[1], mod * s[2]};
    const float G[9] = {g_cov[0], 0.5f * g_cov[1], 0.5f * g_cov[2], 0.5f * g_cov[1], g_cov[3], 0.5f * g_cov[4],
                        0.5f * g_cov[2], 0.5f * g_cov[4], g_cov[5]};
    float dR[9];
    float gsc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int aa = 0; aa < 3; aa++)
#pragma unroll
      for (int bb = 0; bb < 3; bb++) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) acc += G[3 * aa + k] * (Rm[3 * k + bb] * sv[bb]);
        const float dM = 2.f * acc;
        gsc[bb] += Rm[3 * aa + bb] * dM;
        dR[3 * aa + bb] = dM * sv[bb];
      }
    a.dL_dscales[3 * (size_t)i] = mod * gsc[0]; a.dL_dscales[3 * (size_t)i + 1] = mod * gsc[1]; a.dL_dscales[3 * (size_t)i + 2] = mod * gsc[2];
    a.dL_drotations[4 * (size_t)i + 0] = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
    a.dL_drotations[4 * (size_t)i + 1] = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2.f * x * dR[8]);
    a.dL_drotations[4 * (size_t)i + 2] = 2.f * (-2.f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2.f * y * dR[8]);
    a.dL_drotations[4 * (size_t)i + 3] = 2.f * (-2.f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
  }
  }   // i < a.P
  if (use_sh) {                       // SH gradient rows of the block: contiguous in HBM
    __syncthreads();
    float* dst = a.dL_dshs + (size_t)block_base * M3;
    const int total = nblk * M3;
    const int qstep = kThreads / M3, rstep = kThreads - qstep * M3;
    int g = tid / M3, c = tid - g * M3;
    const bool fused = a.peer_delta != nullptr;
    if (vec_peer) {
      // one owner for the whole block (shard_rows is a multiple of the block size): 16-byte reductions, 32 lanes -> 512
      // contiguous bytes per instruction, for all four gradient tensors of the block's 256 Gaussians
      const long long delta = a.peer_delta[min(block_base / a.shard_rows, a.world - 1)];
      auto red4 = [&](float* p, float v0, float v1, float v2, float v3) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<char*>(p) + delta), "f"(v0), "f"(v1), "f"(v2), "f"(v3)
                     : "memory");
      };
      for (int k4 = tid; k4 < total / 4; k4 += kThreads) {             // SH rows: 256 x M3 floats, gathered from the padded staging
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) { const int k = 4 * k4 + e, gg = k / M3; v[e] = s_gsh[gg * sh_stride + (k - gg * M3)]; }
        red4(dst + 4 * k4, v[0], v[1], v[2], v[3]);
      }
      const float4* st4 = reinterpret_cast<const float4*>(s_sh);
      float* dm = a.dL_dmeans3D + 3 * (size_t)block_base;
      for (int k4 = tid; k4 < 3 * kThreads / 4; k4 += kThreads) { const float4 q = st4[k4]; red4(dm + 4 * k4, q.x, q.y, q.z, q.w); }
      float* dc = a.dL_dcov3D + 9 * (size_t)block_base;
      for (int k4 = tid; k4 < 9 * kThreads / 4; k4 += kThreads) { const float4 q = st4[3 * kThreads / 4 + k4]; red4(dc + 4 * k4, q.x, q.y, q.z, q.w); }
      float* dop = a.dL_dopacities + block_base;
      for (int k4 = tid; k4 < kThreads / 4; k4 += kThreads) { const float4 q = st4[12 * kThreads / 4 + k4]; red4(dop + 4 * k4, q.x, q.y, q.z, q.w); }
    } else
    for (int k = tid; k < total; k += kThreads) {
      const float v = s_gsh[g * sh_stride + c];
      if (fused) {
        const long long delta = a.peer_delta[min((block_base + g) / a.shard_rows, a.world - 1)];
        atomicAdd(reinterpret_cast<float*>(reinterpret_cast<char*>(dst + k) + delta), v);     // coalesced: 32 lanes, 128 contiguous bytes
      } else {
        dst[k] = v;
      }
      g += qstep; c += rstep;
      if (c >= M3) { c -= M3; g++; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Camera records: what the reference adapter assembles with ~30 tiny torch launches per call
// (cuda_splatting.py:64-87 + geometry/projection.py:233-247; 0.76 ms of host-driven launches, 3x the whole
// raster pipeline) as ONE launch, one thread per view, evaluated in fp64 and rounded once to fp32.
__device__ inline void inv3(const double* m, double* o) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

__global__ void camera_records_kernel(int V, const float* __restrict__ extrinsics, const float* __restrict__ intrinsics,
                                      const float* __restrict__ near, const float* __restrict__ far, const float* __restrict__ bg,
                                      int scale_invariant, float* __restrict__ views) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double E[16], K[9];
  for (int k = 0; k < 16; k++) E[k] = (double)extrinsics[16 * v + k];
  for (int k = 0; k < 9; k++) K[k] = (double)intrinsics[9 * v + k];
  double nr = (double)near[v], fr = (double)far[v];
  const float scale_f = scale_invariant ? 1.0f / near[v] : 1.0f;     // fp32, exactly as `scale = 1 / near`
  const double scale = (double)scale_f;
  if (scale_invariant) {
    // the reference scales in fp32: extrinsics[:3,3] * scale, near * scale, far * scale
    E[3] = (double)((float)E[3] * scale_f); E[7] = (double)((float)E[7] * scale_f); E[11] = (double)((float)E[11] * scale_f);
    nr = (double)(near[v] * scale_f); fr = (double)(far[v] * scale_f);
  }
  // ---- fov from the normalised intrinsics (get_fov) ----
  double Ki[9];
  inv3(K, Ki);
  auto dir = [&](double x, double y, double* o) {
    o[0] = Ki[0] * x + Ki[1] * y + Ki[2]; o[1] = Ki[3] * x + Ki[4] * y + Ki[5]; o[2] = Ki[6] * x + Ki[7] * y + Ki[8];
    const double n = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
    o[0] /= n; o[1] /= n; o[2] /= n;
  };
  double l[3], r[3], t[3], b[3];
  dir(0.0, 0.5, l); dir(1.0, 0.5, r); dir(0.5, 0.0, t); dir(0.5, 1.0, b);
  const double fov_x = acos(l[0] * r[0] + l[1] * r[1] + l[2] * r[2]);
  const double fov_y = acos(t[0] * b[0] + t[1] * b[1] + t[2] * b[2]);
  const double tx = tan(0.5 * fov_x), ty = tan(0.5 * fov_y);
  // ---- projection matrix (get_projection_matrix) ----
  const double top = ty * nr, right = tx * nr;
  double P[16] = {0};
  P[0] = 2 * nr / (2 * right); P[5] = 2 * nr / (2 * top);
  P[2] = 0.0; P[6] = 0.0;                       // symmetric frustum: (right+left), (top+bottom) = 0
  P[14] = 1.0; P[10] = fr / (fr - nr); P[11] = -(fr * nr) / (fr - nr);
  // ---- world->camera = inverse of the (rigid or general affine) camera-to-world matrix ----
  double R[9] = {E[0], E[1], E[2], E[4], E[5], E[6], E[8], E[9], E[10]}, Ri[9];
  inv3(R, Ri);
  double Vw[16] = {Ri[0], Ri[1], Ri[2], -(Ri[0] * E[3] + Ri[1] * E[7] + Ri[2] * E[11]),
                   Ri[3], Ri[4], Ri[5], -(Ri[3] * E[3] + Ri[4] * E[7] + Ri[5] * E[11]),
                   Ri[6], Ri[7], Ri[8], -(Ri[6] * E[3] + Ri[7] * E[7] + Ri[8] * E[11]),
                   0, 0, 0, 1};
  float* o = views + (size_t)v * kViewFloats;
  // records hold the TRANSPOSES (flat m[c*4+r] = M[r][c]); full projection = P * Vw
  for (int rr = 0; rr < 4; rr++)
    for (int c = 0; c < 4; c++) {
      o[c * 4 + rr] = (float)Vw[rr * 4 + c];
      double acc = 0.0;
      for (int k = 0; k < 4; k++) acc += P[rr * 4 + k] * Vw[k * 4 + c];
      o[16 + c * 4 + rr] = (float)acc;
    }
  o[32] = (float)E[3]; o[33] = (float)E[7]; o[34] = (float)E[11];
  o[35] = bg[3 * v]; o[36] = bg[3 * v + 1]; o[37] = bg[3 * v + 2];
  o[38] = (float)tx; o[39] = (float)ty; o[40] = scale_f;
  for (int k = 41; k < kViewFloats; k++) o[k] = 0.f;
}

int launch_camera_records(int V, const float* ext, const float* K, const float* near, const float* far, const float* bg,
                          int scale_invariant, float* views, cudaStream_t s) {
  if (V <= 0) return FS_OK;
  camera_records_kernel<<<(V + 63) / 64, 64, 0, s>>>(V, ext, K, near, far, bg, scale_invariant, views);
  return check_cuda(cudaGetLastError(), "camera_records_kernel");
}

int launch_preprocess(const FsRasterFwdArgs& a, cudaStream_t s) {
  const int gx = tiles_x(a.W), gy = tiles_y(a.H);
  const size_t nt = (size_t)a.V * gx * gy;
  int rc;
  const bool fused = scan_fused_into_preprocess(a);
  if (fused) {                                       // [counters | cursors | status]: one memset, status[3] = ticket of the fused scan
    if ((rc = check_cuda(cudaMemsetAsync(a.tile_count, 0, (2 * nt + 4) * 4, s), "memset tile counters"))) return rc;
  } else if (a.tile_cursor == a.tile_count + nt) {   // adjacent scratch (the Python wrapper allocates it that way): one memset
    // (+ the bin-overflow flag and the scan-ready word behind the cursors when the call bins directly)
    if ((rc = check_cuda(cudaMemsetAsync(a.tile_count, 0, (2 * nt + (use_bins(a) ? 2 : 0)) * 4, s), "memset tile counters"))) return rc;
  } else {
    if ((rc = check_cuda(cudaMemsetAsync(a.tile_count, 0, nt * 4, s), "memset tile_count"))) return rc;
    if ((rc = check_cuda(cudaMemsetAsync(a.tile_cursor, 0, nt * 4, s), "memset tile_cursor"))) return rc;
  }
  if (a.P > 0) {
    const int sh_stride = (a.M * 3) | 1;
    const size_t smem = a.shs ? (size_t)kThreads * sh_stride * sizeof(float) : 0;
    void (*kern)(FsRasterFwdArgs, int, int, int, int) =
        a.sh_degree == 0 ? preprocess_kernel<0> : a.sh_degree == 1 ? preprocess_kernel<1> : a.sh_degree == 2 ? preprocess_kernel<2> : preprocess_kernel<3>;
    if (smem > 48 * 1024) {
      if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                           "cudaFuncSetAttribute(preprocess_kernel)"))) return rc;
    }
    kern<<<(a.P + kThreads - 1) / kThreads, kThreads, smem, s>>>(a, gx, gy, sh_stride, fused ? 1 : 0);
    if ((rc = check_cuda(cudaGetLastError(), "preprocess_kernel"))) return rc;
  }
  return FS_OK;
}

int launch_preprocess_bwd(const FsRasterBwdArgs& a, cudaStream_t s) {
  if (a.P <= 0) return FS_OK;
  const int sh_stride = (a.M * 3) | 1;
  const size_t smem = (a.shs && a.dL_dshs) ? (size_t)2 * kThreads * sh_stride * sizeof(float) : 0;
  if (smem > 48 * 1024) {
    if (int rc = check_cuda(cudaFuncSetAttribute(preprocess_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(preprocess_bwd_kernel)")) return rc;
  }
  preprocess_bwd_kernel<<<(a.P + kThreads - 1) / kThreads, kThreads, smem, s>>>(a, sh_stride);
  return check_cuda(cudaGetLastError(), "preprocess_bwd_kernel");
}

int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* vis, cudaStream_t s) {
  if (P <= 0) return FS_OK;
  mark_visible_kernel<<<(P + kThreads - 1) / kThreads, kThreads, 0, s>>>(P, means3D, view, vis);
  return check_cuda(cudaGetLastError(), "mark_visible_kernel");
}

}  // namespace fs
