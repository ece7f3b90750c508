// raster_scan.cuh -- exclusive scan of the V*tiles tile counters -> `ranges` (this IS identifyTileRanges), the instance total R
// and the overflow flag in `status`, plus a counting sort of the tiles into 32 population buckets (heaviest-first processing
// order for the render kernel, written over the consumed counters).  One block of NT threads, 4096 / NT consecutive counters per
// thread and pass: 640x480 x 3 views (3600 tiles) is ONE pass for 1024 and for 256 threads alike.
#pragma once
#include "common.cuh"

namespace fs {

__device__ __forceinline__ int order_bucket(uint32_t c) { return 31 - (int)min(31u, c >> 6); }

// `fused`: the scan runs inside preprocess: status[3] is the ticket counter of that kernel and the counters were written with
// atomics by other CTAs: read them with ld.global.cg (L2), never through this SM's L1.
template <int NT>
__device__ __noinline__ void tile_scan_block(uint32_t* __restrict__ count, uint32_t* __restrict__ ranges, uint32_t* __restrict__ status, int n,
                                             long long capacity) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry_s;
  __shared__ uint32_t hist[32], bbase[32];
  constexpr int NW = NT / 32;
  constexpr int IT = 4096 / NT;                      // counters per thread and pass
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0ull;
  if (tid < 32) hist[tid] = 0u;
  __syncthreads();
  const unsigned long long cap = (unsigned long long)capacity;
  for (int base = 0; base < n; base += IT * NT) {
    const int k0 = base + IT * tid;
    uint32_t c[IT];
#pragma unroll
    for (int e = 0; e < IT; e++) c[e] = (k0 + e < n) ? __ldcg(count + k0 + e) : 0u;
    unsigned long long s4 = 0ull;
#pragma unroll
    for (int e = 0; e < IT; e++) s4 += c[e];
    unsigned long long incl = s4;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = lane < NW ? warp_sums[lane] : 0ull;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long carry = carry_s;
    const unsigned long long warp_excl = warp ? warp_sums[warp - 1] : 0ull;
    unsigned long long st = carry + warp_excl + incl - s4;
#pragma unroll
    for (int e = 0; e < IT; e++) {
      const unsigned long long end = st + c[e];
      if (k0 + e < n) {
        // clamp so that a too-small workspace can never be overrun (the call reports overflow)
        ranges[2 * (k0 + e)] = (uint32_t)(st < cap ? st : cap);
        ranges[2 * (k0 + e) + 1] = (uint32_t)(end < cap ? end : cap);
        atomicAdd(&hist[order_bucket((uint32_t)((end < cap ? end : cap) - (st < cap ? st : cap)))], 1u);
      }
      st = end;
    }
    __syncthreads();
    if (tid == NT - 1) carry_s = st;
    __syncthreads();
  }
  if (tid == 0) {
    const unsigned long long R = carry_s;
    status[0] = (uint32_t)(R & 0xffffffffull);
    status[1] = (uint32_t)(R >> 32);
    status[2] = (R > cap || R > 0xffffffffull) ? 1u : 0u;
    status[3] = 0u;
  }
  // ---- tile order: counting sort of the tiles by population bucket, heaviest first (order inside a bucket is
  //      arbitrary: it only decides which CTA starts earlier).  `count` is dead by now and receives the order. ----
  if (warp == 0) {
    const uint32_t h = hist[lane];
    uint32_t incl = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    bbase[lane] = incl - h;
    hist[lane] = 0u;
  }
  __syncthreads();
  for (int k = tid; k < n; k += NT) {
    const int b = order_bucket(ranges[2 * k + 1] - ranges[2 * k]);
    count[bbase[b] + atomicAdd(&hist[b], 1u)] = (uint32_t)k;
  }
}

// The scan is fused into preprocess when there is at least one Gaussian (= at least one preprocess CTA) and the caller's `status`
// buffer sits right behind the tile cursors, so that the call's ONE memset also zeroes the ticket counter status[3].  The rule
// depends only on the arguments, not on `stages`: a caller that runs the stages as separate calls (bench.py's per-stage events)
// gets the scan with the preprocess call and only the scatter with the binning call.
__host__ inline bool scan_fused_into_preprocess(const FsRasterFwdArgs& a) {
  const size_t nt = (size_t)a.V * tiles_x(a.W) * tiles_y(a.H);
  return a.P > 0 && a.status == a.tile_cursor + nt && a.tile_cursor == a.tile_count + nt;
}

// Direct binning (FsRasterFwdArgs::bins): decided from the arguments alone, so that the three stage launches of one call agree.
__host__ __device__ inline bool use_bins(const FsRasterFwdArgs& a) {
  const size_t nt = (size_t)a.V * tiles_x(a.W) * tiles_y(a.H);
  return a.bins != nullptr && a.bin_cap > 0 && a.bin_cap <= kSortSmemKeys && a.P > 0 && !(a.stages & FS_STAGE_RENDER_PACKED) &&
         a.tile_cursor == a.tile_count + nt && a.status != a.tile_cursor + nt;
}

}  // namespace fs
