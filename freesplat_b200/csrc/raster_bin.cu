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

namespace fs {

// ---- 2. scan over tile counters -------------------------------------------------------------
// `count` is consumed by the scan and then OVERWRITTEN with the tile processing order (heaviest bucket first).
__device__ __forceinline__ int order_bucket(uint32_t c) { return 31 - (int)min(31u, c >> 6); }

__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t* __restrict__ count, uint32_t* __restrict__ ranges,
                                                         uint32_t* __restrict__ status, int n, long long capacity) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry_s;
  __shared__ uint32_t hist[32], bbase[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0ull;
  if (tid < 32) hist[tid] = 0u;
  __syncthreads();
  const unsigned long long cap = (unsigned long long)capacity;
  // four consecutive counters per thread: 640x480 x 3 views (3600 tiles) is ONE pass of the block
  for (int base = 0; base < n; base += 4096) {
    const int k0 = base + 4 * tid;
    uint32_t c[4];
#pragma unroll
    for (int e = 0; e < 4; e++) c[e] = (k0 + e < n) ? count[k0 + e] : 0u;
    const unsigned long long s4 = (unsigned long long)c[0] + c[1] + c[2] + c[3];
    unsigned long long incl = s4;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_sums[lane];
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
    for (int e = 0; e < 4; e++) {
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
    if (tid == 1023) carry_s = st;
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
  for (int k = tid; k < n; k += 1024) {
    const int b = order_bucket(ranges[2 * k + 1] - ranges[2 * k]);
    count[bbase[b] + atomicAdd(&hist[b], 1u)] = (uint32_t)k;
  }
}

// ---- 3. scatter instances into their tile ranges --------------------------------------------
__global__ void __launch_bounds__(kThreads) scatter_kernel(FsRasterFwdArgs a, int gx, int gy) {
  if (a.status[2]) return;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  const int v = blockIdx.y;
  const int lane = threadIdx.x & 31;
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
      const uint32_t start = __ldg(a.ranges + 2 * (tbase + tile));
      a.keybuf[start + base + (uint32_t)__popc(grp & ((1u << lane) - 1u))] = key;
    }
    if (++tx == x1) { tx = x0; ty++; }
  }
}

int launch_binning(const FsRasterFwdArgs& a, cudaStream_t s) {
  const int gx = tiles_x(a.W), gy = tiles_y(a.H);
  const int nt = a.V * gx * gy;
  int rc;
  tile_scan_kernel<<<1, 1024, 0, s>>>(a.tile_count, a.ranges, a.status, nt, (long long)a.capacity);
  if ((rc = check_cuda(cudaGetLastError(), "tile_scan_kernel"))) return rc;
  if (a.P > 0) {
    dim3 grid((a.P + kThreads - 1) / kThreads, a.V);
    scatter_kernel<<<grid, kThreads, 0, s>>>(a, gx, gy);
    if ((rc = check_cuda(cudaGetLastError(), "scatter_kernel"))) return rc;
  }
  return FS_OK;
}

}  // namespace fs
