// raster_bin.cu -- tile binning re-designed for B200.
//
// Upstream: InclusiveSum -> D2H read of R -> duplicateWithKeys -> 64-bit global radix sort
// (6 passes over R pairs) -> identifyTileRanges.  Here the (tile | depth) sort is split along
// its two fields, which removes the global sort and the host sync:
//   1. preprocess_kernel already counted the population of every tile (atomic histogram);
//   2. tile_scan_kernel      : exclusive scan of the V*tiles counters -> `ranges` (this IS
//                              identifyTileRanges) and the total R, kept on the device;
//   3. scatter_kernel        : every (Gaussian, tile) instance is appended to its tile's range
//                              as key = depth_bits<<32 | gaussian_index (arbitrary order);
//   4. the render kernel     : the CTA that blends a tile first sorts that tile's range in SHARED MEMORY
//                              (raster_sort.cuh: bitonic network, 64-bit keys; ranges above kSortSmemKeys sort in
//                              global) and writes `point_list`: no separate sort launch, and the barrier-bound sort
//                              of one tile overlaps the issue-bound blending of the other tiles resident on the SM.
//   The scan kernel also emits a heaviest-first tile order (32 population buckets): long tiles start first,
//   which trims the tail of the render launch.
// Sorting by (depth_bits, gaussian_index) reproduces exactly the order of upstream's stable
// radix sort on (tile<<32 | depth_bits) with values emitted in Gaussian order, so
// `point_list` and `ranges` are bit-identical to the reference's (tests/test_raster_gpu.py).
// HBM traffic per instance: 8 B scatter + 8 B read + 8 B + 4 B write, against ~6*24 B upstream.
#include "common.cuh"
#include "raster_math.cuh"
#include "raster_scan.cuh"

namespace fs {

// ---- 2. scan over tile counters -------------------------------------------------------------
// The scan itself lives in raster_scan.cuh (tile_scan_block<NT>): it is run by the LAST CTA of preprocess_kernel to finish
// its tile counting (raster_pre.cu: one launch and ~9 us less per step), and by this stand-alone kernel when a call has no
// Gaussians or runs the binning stage on its own.
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t* __restrict__ count, uint32_t* __restrict__ ranges,
                                                         uint32_t* __restrict__ status, int n, long long capacity) {
  tile_scan_block<1024>(count, ranges, status, n, capacity);
}

// ---- 3. scatter instances into their tile ranges --------------------------------------------
// One 256-thread group `t256` handles Gaussians [256 * bx, +256) of view v.
__device__ __forceinline__ void scatter_group(const FsRasterFwdArgs& a, int gx, int gy, int v, int bx, int t256) {
  const int lane = t256 & 31;
  const int i = bx * kThreads + t256;
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  unsigned long long key = 0ull;
  if (i < a.P) {
    const size_t vi = (size_t)v * a.P + i;
    // issue every load before the first use (radius, centre, depth are independent)
    const float4* rec = reinterpret_cast<const float4*>(a.rec) + 3 * vi;
    const int r = __ldg(a.radii + vi);
    const float2 xy = __ldg(reinterpret_cast<const float2*>(rec));
    const float depth = __ldg(reinterpret_cast<const float*>(rec + 2) + 1);
    if (r > 0) {
      fsm::get_rect(xy.x, xy.y, r, gx, gy, &x0, &y0, &x1, &y1);
      key = ((unsigned long long)__float_as_uint(depth) << 32) | (unsigned long long)(uint32_t)i;
    }
  }
  const int my_tiles = (x1 - x0) * (y1 - y0);
  const int maxn = __reduce_max_sync(0xffffffffu, my_tiles);
  const size_t tbase = (size_t)v * gx * gy;
  int tx = x0, ty = y0;
  // warp-aggregated slot allocation: lanes bound for the same tile share one atomic
  for (int k = 0; k < maxn; k++) {
    const bool on = k < my_tiles;
    const int tile = on ? ty * gx + tx : -1 - lane;
    const unsigned grp = __match_any_sync(0xffffffffu, tile);
    const int leader = __ffs(grp) - 1;
    uint32_t base = 0;
    if (on && lane == leader) base = atomicAdd(a.tile_cursor + tbase + tile, (uint32_t)__popc(grp));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (on) {
      const uint32_t start = __ldcg(a.ranges + 2 * (tbase + tile));
      a.keybuf[start + base + (uint32_t)__popc(grp & ((1u << lane) - 1u))] = key;
    }
    if (++tx == x1) { tx = x0; ty++; }
  }
}

__global__ void __launch_bounds__(kThreads) scatter_kernel(FsRasterFwdArgs a, int gx, int gy) {
  if (a.status[2]) return;
  scatter_group(a, gx, gy, blockIdx.y, blockIdx.x, threadIdx.x);
}

// Direct binning: ONE launch for the tile scan and the (rare) fallback scatter.  CTA 0 scans; the other CTAs leave at once unless a
// bin overflowed (flag word behind the cursors), in which case they wait for the scan (ready word next to the flag: the grid is at
// most one CTA per SM, so CTA 0 is resident and the wait cannot deadlock) and rebuild the compact key ranges, grid-striding over the
// (view, 256-Gaussian group) pairs.  status[3] reports the fallback to the host.
__global__ void __launch_bounds__(1024) scan_scatter_kernel(FsRasterFwdArgs a, int gx, int gy, int nblk_x) {
  const int nt = a.V * gx * gy;
  uint32_t* flag = a.tile_cursor + nt;                     // [0] a bin overflowed   [1] the scan has finished
  __shared__ uint32_t s_over;
  if (blockIdx.x == 0) {
    tile_scan_block<1024>(a.tile_count, a.ranges, a.status, nt, a.capacity);
    __syncthreads();                                       // every range / status store of the CTA is issued
    if (threadIdx.x == 0) {
      s_over = __ldcg(flag);
      a.status[3] = s_over;
      __threadfence();
      atomicExch(flag + 1, 1u);
    }
    __syncthreads();
    if (s_over == 0u) return;
  } else {
    if (threadIdx.x == 0) {
      s_over = __ldcg(flag);
      if (s_over) {
        while (atomicAdd(flag + 1, 0u) == 0u) {}
        __threadfence();
      }
    }
    __syncthreads();
    if (s_over == 0u) return;
  }
  if (__ldcg(a.status + 2)) return;                        // R exceeds the workspace: the call reports overflow
  const int groups = nblk_x * a.V;
  for (int gidx = (int)blockIdx.x * 4 + (int)(threadIdx.x >> 8); gidx < groups; gidx += (int)gridDim.x * 4) {
    const int v = gidx / nblk_x;
    scatter_group(a, gx, gy, v, gidx - v * nblk_x, (int)(threadIdx.x & 255));
  }
}

int launch_binning(const FsRasterFwdArgs& a, cudaStream_t s) {
  const int gx = tiles_x(a.W), gy = tiles_y(a.H);
  const int nt = a.V * gx * gy;
  int rc;
  if (use_bins(a)) {                                       // (implies a separate status buffer: no fused scan)
    const int nblk_x = (a.P + kThreads - 1) / kThreads;
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    scan_scatter_kernel<<<sms > 0 ? sms : 1, 1024, 0, s>>>(a, gx, gy, nblk_x);
    return check_cuda(cudaGetLastError(), "scan_scatter_kernel");
  }
  // the scan normally ran inside preprocess (last CTA); stand-alone only if that stage was not part of this call sequence
  if (!scan_fused_into_preprocess(a)) {
    tile_scan_kernel<<<1, 1024, 0, s>>>(a.tile_count, a.ranges, a.status, nt, (long long)a.capacity);
    if ((rc = check_cuda(cudaGetLastError(), "tile_scan_kernel"))) return rc;
  }
  if (a.P > 0) {
    dim3 grid((a.P + kThreads - 1) / kThreads, a.V);
    scatter_kernel<<<grid, kThreads, 0, s>>>(a, gx, gy);
    if ((rc = check_cuda(cudaGetLastError(), "scatter_kernel"))) return rc;
  }
  return FS_OK;
}

}  // namespace fs
