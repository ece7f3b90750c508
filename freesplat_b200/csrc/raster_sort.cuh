// raster_sort.cuh -- block-wide bitonic sort of one tile's 64-bit (depth_bits << 32 | gaussian) keys
// (the depth half of upstream's global radix sort, see raster_bin.cu).  Used by the render kernel, which
// sorts its own tile in shared memory before blending it.
#pragma once
#include "common.cuh"

namespace fs {

// Bitonic network with ascending comparators only (first step of every merge is the mirrored
// "flip" step), so virtual +inf padding above n needs no storage.  Compare-exchange c of a stage
// touches only the aligned 64-element region [64*(c/32), +64) whenever the stage's block size is
// <= 64, and warp w always owns CEs {32w..32w+31} (+256m): such stages only need __syncwarp().
// For N = 512 that leaves 9 block-wide barriers out of 45 stages.  All index math is shifts/masks.
__device__ __forceinline__ void stage_sync(bool block_wide) {
  if (block_wide) __syncthreads(); else __syncwarp();
}

template <typename KeyPtr, int NT = kThreads>
__device__ __forceinline__ void bitonic_sort_block(KeyPtr keys, int n, int tid) {
  int logN = 0;
  while ((1 << logN) < n) logN++;
  const int halfN = (1 << logN) >> 1;
  int prevB = 1 << 30;                                   // "previous stage" before the first one: block-wide
  for (int lk = 1; lk <= logN; lk++) {
    const int k = 1 << lk, half = k >> 1;
    stage_sync(k > 64 || prevB > 64);
    for (int c = tid; c < halfN; c += NT) {               // flip step
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
      for (int c = tid; c < halfN; c += NT) {
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
template <int LOGN, int NT = kThreads>
__device__ __forceinline__ void bitonic_sort_fixed(unsigned long long* keys, int n, int tid) {
  constexpr int HALF = (1 << LOGN) >> 1;
  constexpr int PER = (HALF + NT - 1) / NT;          // compare-exchanges per thread and stage (1 at NT = 256, up to 2 at NT = 128)
  int prevB = 1 << 30;
#pragma unroll
  for (int lk = 1; lk <= LOGN; lk++) {
    const int k = 1 << lk, half = k >> 1;
    stage_sync(k > 64 || prevB > 64);
#pragma unroll
    for (int m = 0; m < PER; m++) {
      const int c = tid + m * NT;
      if (c < HALF) {
        const int b = c >> (lk - 1), off = c & (half - 1);
        const int i = (b << lk) + off, l = (b << lk) + (k - 1 - off);
        if (l < n) {
          const unsigned long long ki = keys[i], kl = keys[l];
          if (ki > kl) { keys[i] = kl; keys[l] = ki; }
        }
      }
    }
    prevB = k;
#pragma unroll
    for (int lj = lk - 2; lj >= 0; lj--) {
      const int j = 1 << lj, B = j << 1;
      stage_sync(B > 64 || prevB > 64);
#pragma unroll
      for (int m = 0; m < PER; m++) {
        const int c = tid + m * NT;
        if (c < HALF) {
          const int b = c >> lj, off = c & (j - 1);
          const int i = (b << (lj + 1)) + off, l = i + j;
          if (l < n) {
            const unsigned long long ki = keys[i], kl = keys[l];
            if (ki > kl) { keys[i] = kl; keys[l] = ki; }
          }
        }
      }
      prevB = B;
    }
  }
  __syncthreads();
}

}  // namespace fs
