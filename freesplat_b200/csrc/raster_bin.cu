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
//   4. tile_sort_kernel      : one CTA per tile sorts its range in SHARED MEMORY (bitonic
//                              network, 64-bit keys; ranges above kSortSmemKeys sort in global).
// Sorting by (depth_bits, gaussian_index) reproduces exactly the order of upstream's stable
// radix sort on (tile<<32 | depth_bits) with values emitted in Gaussian order, so
// `point_list` and `ranges` are bit-identical to the reference's (tests/test_raster_gpu.py).
// HBM traffic per instance: 8 B scatter + 8 B read + 8 B + 4 B write, against ~6*24 B upstream.
#include "common.cuh"
#include "raster_math.cuh"

namespace fs {

// ---- 2. scan over tile counters -------------------------------------------------------------
__global__ void __launch_bounds__(1024) tile_scan_kernel(const uint32_t* __restrict__ count, uint32_t* __restrict__ ranges,
                                                         uint32_t* __restrict__ status, int n, long long capacity) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0ull;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int k = base + tid;
    const unsigned long long c = (k < n) ? (unsigned long long)count[k] : 0ull;
    unsigned long long incl = c;
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
    const unsigned long long end = carry + warp_excl + incl;
    if (k < n) {
      // clamp so that a too-small workspace can never be overrun (the call reports overflow)
      const unsigned long long cap = (unsigned long long)capacity;
      const unsigned long long st = end - c;
      ranges[2 * k] = (uint32_t)(st < cap ? st : cap);
      ranges[2 * k + 1] = (uint32_t)(end < cap ? end : cap);
    }
    __syncthreads();
    if (tid == 1023) carry_s = end;
    __syncthreads();
  }
  if (tid == 0) {
    const unsigned long long R = carry_s;
    status[0] = (uint32_t)(R & 0xffffffffull);
    status[1] = (uint32_t)(R >> 32);
    status[2] = (R > (unsigned long long)capacity || R > 0xffffffffull) ? 1u : 0u;
    status[3] = 0u;
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

// ---- 4. per-tile sort -------------------------------------------------------------------------
// Bitonic network with ascending comparators only (first step of every merge is the mirrored
// "flip" step), so virtual +inf padding above n needs no storage.  Compare-exchange c of a stage
// touches only the aligned 64-element region [64*(c/32), +64) whenever the stage's block size is
// <= 64, and warp w always owns CEs {32w..32w+31} (+256m): such stages only need __syncwarp().
// For N = 512 that leaves 9 block-wide barriers out of 45 stages.  All index math is shifts/masks.
__device__ __forceinline__ void stage_sync(bool block_wide) {
  if (block_wide) __syncthreads(); else __syncwarp();
}

template <typename KeyPtr>
__device__ __forceinline__ void bitonic_sort_block(KeyPtr keys, int n, int tid) {
  int logN = 0;
  while ((1 << logN) < n) logN++;
  const int halfN = (1 << logN) >> 1;
  int prevB = 1 << 30;                                   // "previous stage" before the first one: block-wide
  for (int lk = 1; lk <= logN; lk++) {
    const int k = 1 << lk, half = k >> 1;
    stage_sync(k > 64 || prevB > 64);
    for (int c = tid; c < halfN; c += kThreads) {        // flip step
      const int b = c >> (lk - 1), off = c & (half - 1);
      const int i = (b << lk) + off, l = (b << lk) + (k - 1 - off);
      if (l < n) {
        const unsigned long long ki = keys[i], kl = keys[l];
        if (ki > kl) { keys[i] = kl; keys[l] = ki; }
      }
    }
    prevB = k;
    for (int lj = lk - 2; lj >= 0; lj--) {
      const int j = 1 << lj, B = j << 1;
      stage_sync(B > 64 || prevB > 64);
      for (int c = tid; c < halfN; c += kThreads) {
        const int b = c >> lj, off = c & (j - 1);
        const int i = (b << (lj + 1)) + off, l = i + j;
        if (l < n) {
          const unsigned long long ki = keys[i], kl = keys[l];
          if (ki > kl) { keys[i] = kl; keys[l] = ki; }
        }
      }
      prevB = B;
    }
  }
  __syncthreads();
}

// Fully unrolled network for N = 2^LOGN <= 512 keys: one compare-exchange per thread and stage, all
// shifts/masks compile-time constants (the generic loop spent ~64 instructions per stage, ncu r1b).
template <int LOGN>
__device__ __forceinline__ void bitonic_sort_fixed(unsigned long long* keys, int n, int tid) {
  constexpr int HALF = (1 << LOGN) >> 1;
  static_assert(HALF <= kThreads, "one compare-exchange per thread");
  const bool has_ce = tid < HALF;
  int prevB = 1 << 30;
#pragma unroll
  for (int lk = 1; lk <= LOGN; lk++) {
    const int k = 1 << lk, half = k >> 1;
    stage_sync(k > 64 || prevB > 64);
    if (has_ce) {
      const int b = tid >> (lk - 1), off = tid & (half - 1);
      const int i = (b << lk) + off, l = (b << lk) + (k - 1 - off);
      if (l < n) {
        const unsigned long long ki = keys[i], kl = keys[l];
        if (ki > kl) { keys[i] = kl; keys[l] = ki; }
      }
    }
    prevB = k;
#pragma unroll
    for (int lj = lk - 2; lj >= 0; lj--) {
      const int j = 1 << lj, B = j << 1;
      stage_sync(B > 64 || prevB > 64);
      if (has_ce) {
        const int b = tid >> lj, off = tid & (j - 1);
        const int i = (b << (lj + 1)) + off, l = i + j;
        if (l < n) {
          const unsigned long long ki = keys[i], kl = keys[l];
          if (ki > kl) { keys[i] = kl; keys[l] = ki; }
        }
      }
      prevB = B;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) tile_sort_kernel(const uint32_t* __restrict__ ranges, unsigned long long* __restrict__ keybuf,
                                                             uint32_t* __restrict__ point_list, const uint32_t* __restrict__ status) {
  if (status[2]) return;
  extern __shared__ unsigned long long skeys[];
  const int t = blockIdx.x;
  const uint32_t start = ranges[2 * t], end = ranges[2 * t + 1];
  const int n = (int)(end - start);
  if (n <= 0) return;
  const int tid = threadIdx.x;
  unsigned long long* g = keybuf + start;
  if (n <= kSortSmemKeys) {
    for (int k = tid; k < n; k += kThreads) skeys[k] = g[k];
    if (n <= 32) bitonic_sort_fixed<5>(skeys, n, tid);
    else if (n <= 64) bitonic_sort_fixed<6>(skeys, n, tid);
    else if (n <= 128) bitonic_sort_fixed<7>(skeys, n, tid);
    else if (n <= 256) bitonic_sort_fixed<8>(skeys, n, tid);
    else if (n <= 512) bitonic_sort_fixed<9>(skeys, n, tid);
    else bitonic_sort_block(skeys, n, tid);
    for (int k = tid; k < n; k += kThreads) {
      const unsigned long long key = skeys[k];
      g[k] = key;
      point_list[start + k] = (uint32_t)(key & 0xffffffffull);
    }
  } else {
    bitonic_sort_block(g, n, tid);             // rare: very crowded tile, sort in L2/HBM
    for (int k = tid; k < n; k += kThreads) point_list[start + k] = (uint32_t)(g[k] & 0xffffffffull);
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
    tile_sort_kernel<<<nt, kThreads, kSortSmemKeys * 8, s>>>(a.ranges, reinterpret_cast<unsigned long long*>(a.keybuf),
                                                              a.point_list, a.status);
    if ((rc = check_cuda(cudaGetLastError(), "tile_sort_kernel"))) return rc;
  }
  return FS_OK;
}

}  // namespace fs
