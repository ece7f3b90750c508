// raster_render.cu -- per-tile alpha blending and its gradient (SURVEY §8a R6, R7;
// upstream forward.cu::renderCUDA / backward.cu::renderCUDA, with the depth channel of
// the `-w-depth` fork: D += z*alpha*T).
//
// One CTA = one 16x16 tile of one view (grid = tiles x V), 8 warps; warp w owns an 8x4-pixel
// sub-tile.  Per batch of 256 sorted instances the CTA stages the 48-byte projected records
// (three coalesced 16-byte loads per instance, gathered by `point_list`) in shared memory;
// every warp then tests 32 records at a time against its sub-tile with one ballot (the
// record's conservative alpha>=1/255 extent, see raster_math.cuh::alpha_extent) and only
// walks the set bits, front to back.  All 32 lanes of a warp process the SAME instance
// (shared-memory broadcast reads), so the backward pass reduces the nine per-Gaussian
// partial derivatives with a butterfly of warp shuffles before touching memory.
//
// Per pixel and instance the arithmetic is the canonical order of oracle/raster_oracle.c:
//   t = fma(cb, dy, ca*dx) ; power = fma(cc*dy, dy, t*dx) ; alpha = min(.99, o*exp(power)).
// exp() is MUFU.EX2 (ex2.approx.ftz) of power*log2(e): |rel err| < 1e-6.
#include "common.cuh"
#include "raster_sort.cuh"
#include "raster_scan.cuh"

namespace fs {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------- forward
// Explicit shared-memory addressing (ld.shared with a 32-bit address): the compiler's generic path
// re-derived the shared window base (S2R SR_CgaCtaId + LEA chain) on every loop iteration.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kLog2e = 1.4426950408889634f;

// Staged per-instance data (56 B): the exponent coefficients are pre-multiplied by log2(e) so that
// the per-pixel chain is  t = fma(cb,dy,ca*dx); p2 = fma(cc*dy,dy,t*dx); alpha = min(.99, o*ex2(p2)).
// instances staged per batch: 2 per thread.  A tile of up to 512 instances (the mean of config 2 is 345) is staged ONCE and its
// warps then blend to the end without meeting at another barrier (256-instance batches: two more block barriers per batch, and
// every warp waited at each of them for the sub-tile with the most overlapping instances)
constexpr int kBatch = 2 * kThreads;
struct __align__(16) RenderSmem {
  float4 T[kBatch];   // x, y, hx, hy         (sub-tile overlap test)
  float4 A[kBatch];   // x, y, ca, cb         (hot loop)
  float4 C[kBatch];   // r, g, b, depth       (only when the pixel is actually hit)
  float2 B[kBatch];   // cc, opacity          (hot loop)
};

// The kernel first SORTS its tile: the tile's unsorted (depth_bits << 32 | gaussian) keys are pulled into
// shared memory, sorted with the block-wide bitonic network and written back together with `point_list` (the backward
// pass and the parity tests read both); the same shared memory is then reused for the blending batches.  Tiles are taken
// in the heaviest-first order the scan kernel left in `order`.
static_assert(sizeof(RenderSmem) <= (size_t)kSortSmemKeys * 8, "blend staging must fit in the sort buffer");

__global__ void __launch_bounds__(kThreads) render_fwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, unsigned long long* keybuf, uint32_t* point_list,
    const float4* __restrict__ rec, const float* __restrict__ views, const uint32_t* __restrict__ status, int P, int H, int W,
    int gx, int ntiles, float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, const unsigned long long* __restrict__ bins, int bin_cap, const uint32_t* __restrict__ bin_flag) {
  if (status[2]) return;
  extern __shared__ unsigned long long skeys[];
  RenderSmem& sm = *reinterpret_cast<RenderSmem*>(skeys);
  const int t_flat = (int)order[blockIdx.x];
  const int v = t_flat / ntiles, tile = t_flat - v * ntiles;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = tile % gx, tile_y = tile / gx;
  const int x0 = tile_x * 16 + (warp & 1) * 8, y0 = tile_y * 16 + (warp >> 1) * 4;
  const int px = x0 + (lane & 7), py = y0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float fx0 = (float)x0, fx1 = (float)(x0 + 7), fy0 = (float)y0, fy1 = (float)(y0 + 3);
  const uint2 range = ranges[t_flat];
  const float4* __restrict__ rec_v = rec + (size_t)v * P * 3;
  // ---- sort this tile's keys by (depth bits, Gaussian index) ----
  {
    const int n = (int)(range.y - range.x);
    if (n > 0) {
      unsigned long long* g = keybuf + range.x;
      // direct binning: the unsorted keys sit in the tile's bin (unless a bin overflowed and the call fell back to the scatter pass)
      const unsigned long long* src = (bins != nullptr && *bin_flag == 0u) ? bins + (size_t)t_flat * (size_t)bin_cap : g;
      if (n <= kSortSmemKeys) {
        for (int k = tid; k < n; k += kThreads) skeys[k] = src[k];
        if (n <= 32) bitonic_sort_fixed<5>(skeys, n, tid);
        else if (n <= 64) bitonic_sort_fixed<6>(skeys, n, tid);
        else if (n <= 128) bitonic_sort_fixed<7>(skeys, n, tid);
        else if (n <= 256) bitonic_sort_fixed<8>(skeys, n, tid);
        else if (n <= 512) bitonic_sort_fixed<9>(skeys, n, tid);
        else bitonic_sort_block(skeys, n, tid);
        for (int k = tid; k < n; k += kThreads) {
          const unsigned long long key = skeys[k];
          g[k] = key;
          point_list[range.x + k] = (uint32_t)(key & 0xffffffffull);
        }
      } else {
        bitonic_sort_block(g, n, tid);             // rare: very crowded tile, sort in L2/HBM
        for (int k = tid; k < n; k += kThreads) point_list[range.x + k] = (uint32_t)(g[k] & 0xffffffffull);
      }
    }
    __syncthreads();                               // point_list visible to the whole CTA; shared memory free for reuse
  }
  const uint32_t aT = (uint32_t)__cvta_generic_to_shared(sm.T), aA = (uint32_t)__cvta_generic_to_shared(sm.A),
                 aC = (uint32_t)__cvta_generic_to_shared(sm.C), aB = (uint32_t)__cvta_generic_to_shared(sm.B);

  // T > 0: pixel still accumulating.  T < 0: finished, |T| is its final transmittance.
  float T = inside ? 1.f : -1.f;
  float C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;
  uint32_t last = 0;

  for (uint32_t base = range.x; base < range.y; base += kBatch) {
    if (__syncthreads_count(T < 0.f) == kThreads) break;   // barrier also protects the smem reuse
    const int n = min((int)kBatch, (int)(range.y - base));
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
      const int sl = tid + h2 * kThreads;
      if (sl < n) {
        const uint32_t id = point_list[base + sl];
        const float4* r = rec_v + 3 * (size_t)id;
        const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
        sm.T[sl] = make_float4(r0.x, r0.y, r2.z, r2.w);
        sm.A[sl] = make_float4(r0.x, r0.y, (-0.5f * r0.z) * kLog2e, (-r0.w) * kLog2e);
        sm.B[sl] = make_float2((-0.5f * r1.x) * kLog2e, r1.y);
        sm.C[sl] = make_float4(r1.z, r1.w, r2.x, r2.y);
      }
    }
    __syncthreads();
    if (__all_sync(kFull, T < 0.f)) continue;
    const uint32_t pos0 = base - range.x + 1u;      // contributor index of staged slot 0
    for (int g = 0; g < n; g += 32) {
      const int j = g + lane;
      bool hit = false;
      if (j < n) {
        const float4 q = lds128(aT + (uint32_t)j * 16u);
        hit = (q.x - q.z <= fx1) && (q.x + q.z >= fx0) && (q.y - q.w <= fy1) && (q.y + q.w >= fy0);
      }
      unsigned m = __ballot_sync(kFull, hit);
      while (m) {
        const uint32_t jj = (uint32_t)(g + __ffs(m) - 1);
        m &= m - 1;
        const float4 a = lds128(aA + jj * 16u);
        const float2 bo = lds64(aB + jj * 8u);
        const float dx = a.x - pxf, dy = a.y - pyf;
        const float t = fmaf(a.w, dy, a.z * dx);
        const float p2 = fmaf(bo.x * dy, dy, t * dx);
        const float alpha = fminf(0.99f, bo.y * ex2_approx(p2));
        // branch-free accumulate: ~94% of the walked instances touch at least one lane, so a
        // divergent branch here only added BSSY/BSYNC and register shuffling (ncu r1b).
        const float4 c = lds128(aC + jj * 16u);
        const bool pass = (p2 <= 0.f) && (alpha >= 1.f / 255.f) && (T > 0.f);
        const float test_T = T * (1.f - alpha);
        const bool fin = pass && (test_T < 0.0001f);
        const bool acc = pass && !fin;
        const float w = acc ? alpha * T : 0.f;
        C0 = fmaf(c.x, w, C0); C1 = fmaf(c.y, w, C1); C2 = fmaf(c.z, w, C2);
        Dp = fmaf(c.w, w, Dp);
        T = fin ? -T : (acc ? test_T : T);
        last = acc ? pos0 + jj : last;
      }
      if (__all_sync(kFull, T < 0.f)) break;
    }
  }
  if (inside) {
    const float T_fin = fabsf(T);
    const float* bg = views + (size_t)v * kViewFloats + 35;
    const size_t HW = (size_t)H * W, pix = (size_t)py * W + px;
    float* oc = out_color + (size_t)v * 3 * HW;
    oc[pix] = fmaf(T_fin, bg[0], C0);
    oc[HW + pix] = fmaf(T_fin, bg[1], C1);
    oc[2 * HW + pix] = fmaf(T_fin, bg[2], C2);
    out_depth[(size_t)v * HW + pix] = Dp;
    final_T[(size_t)v * HW + pix] = T_fin;
    n_contrib[(size_t)v * HW + pix] = last;
  }
}

// ------------------------------------------------------------------------------- forward, two pixels per lane
// Variant of render_fwd_kernel for the issue-bound regime (ncu r1: 75 % of the issue slots busy, 39 instructions per walked
// (warp, instance)): a CTA of 128 threads renders the 16x16 tile, warp w owns an 8x8 sub-tile and every lane TWO vertically
// adjacent pixels, (px, py) and (px, py + 1).  The two HALVES of a warp (lanes 0-15: rows 0-3, lanes 16-31: rows 4-7 of the
// sub-tile) keep the 8x4 culling granularity of the one-pixel kernel: each half walks ITS OWN list of overlapping instances
// (one ballot per half box; a lane reads the instance its half is at: two broadcast addresses per shared-memory load), so
// one loop iteration retires an (8x4 region, instance) pair for both halves at once.
// The pair shares dx; dy, the exponent, alpha, the transmittance test and the four
// accumulations run on Blackwell's packed fp32 pipe (fma.rn.f32x2 / mul.rn.f32x2 / add.rn.f32x2 -> FFMA2 / FMUL2 / FADD2):
// one instruction for both pixels.  Per-instance scalars (cb, cc, opacity, colour) enter as (s, s) pairs, which the SASS
// encodes as a broadcast operand (`R.F32`) of the packed instruction: no duplicate staging, no MOVs.
// Per pixel the arithmetic (and therefore every bit of the result) is that of
// render_fwd_kernel:  t = fma(cb,dy,ca*dx); p = fma(cc*dy,dy,t*dx); alpha = min(.99, o*ex2(p)); T' = T*(1-alpha); C += c*(alpha*T).
using f2 = unsigned long long;
__device__ __forceinline__ f2 pk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// (lo, hi) += (s, s) * (wlo, whi): the loop-carried accumulators stay 32-bit values for the register allocator (64-bit
// loop-carried asm operands cost two MOVs per accumulator and iteration: the destination pair was never coalesced)
__device__ __forceinline__ void fma2_acc(float& lo, float& hi, float s, float wlo, float whi) {
  asm("{\n\t.reg .b64 a, w, c;\n\tmov.b64 a, {%2, %2};\n\tmov.b64 w, {%3, %4};\n\tmov.b64 c, {%0, %1};\n\t"
      "fma.rn.f32x2 c, a, w, c;\n\tmov.b64 {%0, %1}, c;\n\t}" : "+f"(lo), "+f"(hi) : "f"(s), "f"(wlo), "f"(whi));
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void lds128_2(uint32_t addr, f2& a, f2& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ f2 lds64_1(uint32_t addr) { f2 a; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a) : "r"(addr)); return a; }

constexpr int kThreads2 = 128;
constexpr int kSortSmemKeys2 = 2048;      // 16 KB: up to 14 CTAs of this kernel fit one SM; more crowded tiles sort in global memory

struct __align__(16) Render2Smem {
  float4 T[kThreads2];    // x, y, hx, hy           (sub-tile overlap test)
  float4 L0[kThreads2];   // x, ca, y, cb
  float4 L1[kThreads2];   // cc, o, r, g
  float2 L2[kThreads2];   // b, depth
};
static_assert(sizeof(Render2Smem) <= (size_t)kSortSmemKeys2 * 8, "blend staging must fit in the sort buffer");

__global__ void __launch_bounds__(kThreads2, 8) render_fwd2_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, unsigned long long* keybuf, uint32_t* point_list,
    const float4* __restrict__ rec, const float* __restrict__ views, const uint32_t* __restrict__ status, int P, int H, int W,
    int gx, int ntiles, float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib) {
  if (status[2]) return;
  extern __shared__ unsigned long long skeys[];
  Render2Smem& sm = *reinterpret_cast<Render2Smem*>(skeys);
  const int t_flat = (int)order[blockIdx.x];
  const int v = t_flat / ntiles, tile = t_flat - v * ntiles;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = tile % gx, tile_y = tile / gx;
  const int x0 = tile_x * 16 + (warp & 1) * 8, y0 = tile_y * 16 + (warp >> 1) * 8;
  const int half = lane >> 4, hl = lane & 15;
  const int px = x0 + (hl & 7), pyA = y0 + half * 4 + (hl >> 3) * 2, pyB = pyA + 1;
  const bool insideA = px < W && pyA < H, insideB = px < W && pyB < H;
  const float pxf = (float)px;
  const f2 npy2 = pk2(-(float)pyA, -(float)pyB);
  const float fx0 = (float)x0, fx1 = (float)(x0 + 7);
  const float fyA0 = (float)y0, fyA1 = (float)(y0 + 3), fyB0 = (float)(y0 + 4), fyB1 = (float)(y0 + 7);
  const uint2 range = ranges[t_flat];
  const float4* __restrict__ rec_v = rec + (size_t)v * P * 3;
  // ---- sort this tile's keys by (depth bits, Gaussian index) ----
  {
    const int n = (int)(range.y - range.x);
    if (n > 0) {
      unsigned long long* g = keybuf + range.x;
      if (n <= kSortSmemKeys2) {
        for (int k = tid; k < n; k += kThreads2) skeys[k] = g[k];
        if (n <= 32) bitonic_sort_fixed<5, kThreads2>(skeys, n, tid);
        else if (n <= 64) bitonic_sort_fixed<6, kThreads2>(skeys, n, tid);
        else if (n <= 128) bitonic_sort_fixed<7, kThreads2>(skeys, n, tid);
        else if (n <= 256) bitonic_sort_fixed<8, kThreads2>(skeys, n, tid);
        else if (n <= 512) bitonic_sort_fixed<9, kThreads2>(skeys, n, tid);
        else bitonic_sort_block<unsigned long long*, kThreads2>(skeys, n, tid);
        for (int k = tid; k < n; k += kThreads2) {
          const unsigned long long key = skeys[k];
          g[k] = key;
          point_list[range.x + k] = (uint32_t)(key & 0xffffffffull);
        }
      } else {
        bitonic_sort_block<unsigned long long*, kThreads2>(g, n, tid);
        for (int k = tid; k < n; k += kThreads2) point_list[range.x + k] = (uint32_t)(g[k] & 0xffffffffull);
      }
    }
    __syncthreads();
  }
  const uint32_t aT = (uint32_t)__cvta_generic_to_shared(sm.T), a0 = (uint32_t)__cvta_generic_to_shared(sm.L0),
                 a1 = (uint32_t)__cvta_generic_to_shared(sm.L1), a2 = (uint32_t)__cvta_generic_to_shared(sm.L2);

  // T > 0: pixel still accumulating.  T < 0: finished, |T| is its final transmittance.
  float TA = insideA ? 1.f : -1.f, TB = insideB ? 1.f : -1.f;
  float c0A = 0.f, c0B = 0.f, c1A = 0.f, c1B = 0.f, c2A = 0.f, c2B = 0.f, dA = 0.f, dB = 0.f;
  uint32_t lastA = 0, lastB = 0;
  const f2 one2 = pk2(1.f, 1.f), mone2 = pk2(-1.f, -1.f);

  for (uint32_t base = range.x; base < range.y; base += kThreads2) {
    if (__syncthreads_count(TA < 0.f && TB < 0.f) == kThreads2) break;   // barrier also protects the smem reuse
    const int n = min((int)kThreads2, (int)(range.y - base));
    if (tid < n) {
      const uint32_t id = point_list[base + tid];
      const float4* r = rec_v + 3 * (size_t)id;
      const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
      const float ca = (-0.5f * r0.z) * kLog2e, cb = (-r0.w) * kLog2e, cc = (-0.5f * r1.x) * kLog2e;
      sm.T[tid] = make_float4(r0.x, r0.y, r2.z, r2.w);
      sm.L0[tid] = make_float4(r0.x, ca, r0.y, cb);
      sm.L1[tid] = make_float4(cc, r1.y, r1.z, r1.w);
      sm.L2[tid] = make_float2(r2.x, r2.y);
    }
    __syncthreads();
    if (__all_sync(kFull, TA < 0.f && TB < 0.f)) continue;
    const uint32_t pos0 = base - range.x + 1u;      // contributor index of staged slot 0
    for (int g = 0; g < n; g += 32) {
      const int j = g + lane;
      bool hitA = false, hitB = false;
      if (j < n) {
        const float4 q = lds128(aT + (uint32_t)j * 16u);
        const bool hx = (q.x - q.z <= fx1) && (q.x + q.z >= fx0);
        hitA = hx && (q.y - q.w <= fyA1) && (q.y + q.w >= fyA0);
        hitB = hx && (q.y - q.w <= fyB1) && (q.y + q.w >= fyB0);
      }
      const unsigned mA = __ballot_sync(kFull, hitA), mB = __ballot_sync(kFull, hitB);
      unsigned m = half ? mB : mA;                    // the instances THIS half of the warp still has to walk
      while (__any_sync(kFull, m != 0u)) {
        const bool valid = m != 0u;
        const uint32_t jj = (uint32_t)g + (valid ? (uint32_t)(__ffs(m) - 1) : 0u);
        m &= m - 1;
        const float4 q0 = lds128(a0 + jj * 16u);     // x, ca, y, cb
        const float4 q1 = lds128(a1 + jj * 16u);     // cc, o, r, g
        const float dx = q0.x - pxf;
        const float mm = q0.y * dx;
        const f2 dy2 = add2(pk2(q0.z, q0.z), npy2);
        const f2 t2 = fma2(pk2(q0.w, q0.w), dy2, pk2(mm, mm));
        const f2 p2 = fma2(mul2(pk2(q1.x, q1.x), dy2), dy2, mul2(t2, pk2(dx, dx)));
        float pA, pB;
        unpk2(p2, pA, pB);
        const f2 e2 = pk2(ex2_approx(pA), ex2_approx(pB));
        float alA, alB;
        unpk2(mul2(pk2(q1.y, q1.y), e2), alA, alB);
        alA = fminf(0.99f, alA); alB = fminf(0.99f, alB);
        const f2 al2 = pk2(alA, alB);
        const f2 T2 = pk2(TA, TB);
        float ttA, ttB;
        unpk2(mul2(T2, fma2(al2, mone2, one2)), ttA, ttB);          // T * (1 - alpha)
        const bool passA = valid && (pA <= 0.f) && (alA >= 1.f / 255.f) && (TA > 0.f);
        const bool passB = valid && (pB <= 0.f) && (alB >= 1.f / 255.f) && (TB > 0.f);
        const bool finA = passA && (ttA < 0.0001f), finB = passB && (ttB < 0.0001f);
        const bool accA = passA && !finA, accB = passB && !finB;
        float wA, wB;
        unpk2(mul2(al2, T2), wA, wB);
        wA = accA ? wA : 0.f; wB = accB ? wB : 0.f;
        const float2 q2 = lds64(a2 + jj * 8u);       // b, depth
        fma2_acc(c0A, c0B, q1.z, wA, wB); fma2_acc(c1A, c1B, q1.w, wA, wB);
        fma2_acc(c2A, c2B, q2.x, wA, wB); fma2_acc(dA, dB, q2.y, wA, wB);
        TA = finA ? -TA : (accA ? ttA : TA);
        TB = finB ? -TB : (accB ? ttB : TB);
        lastA = accA ? pos0 + jj : lastA;
        lastB = accB ? pos0 + jj : lastB;
      }
      if (__all_sync(kFull, TA < 0.f && TB < 0.f)) break;
    }
  }
  const float* bg = views + (size_t)v * kViewFloats + 35;
  const size_t HW = (size_t)H * W;
  float* oc = out_color + (size_t)v * 3 * HW;
  if (insideA) {
    const float T_fin = fabsf(TA);
    const size_t pix = (size_t)pyA * W + px;
    oc[pix] = fmaf(T_fin, bg[0], c0A); oc[HW + pix] = fmaf(T_fin, bg[1], c1A); oc[2 * HW + pix] = fmaf(T_fin, bg[2], c2A);
    out_depth[(size_t)v * HW + pix] = dA; final_T[(size_t)v * HW + pix] = T_fin; n_contrib[(size_t)v * HW + pix] = lastA;
  }
  if (insideB) {
    const float T_fin = fabsf(TB);
    const size_t pix = (size_t)pyB * W + px;
    oc[pix] = fmaf(T_fin, bg[0], c0B); oc[HW + pix] = fmaf(T_fin, bg[1], c1B); oc[2 * HW + pix] = fmaf(T_fin, bg[2], c2B);
    out_depth[(size_t)v * HW + pix] = dB; final_T[(size_t)v * HW + pix] = T_fin; n_contrib[(size_t)v * HW + pix] = lastB;
  }
}

// ------------------------------------------------------------------------------- backward
// Butterfly reduction of 8 per-lane values across the warp: 9 shuffles instead of 40.
// On return lane L holds, in r, the warp-wide sum of value number (L >> 2) (all 4 lanes of a
// quad hold the same total).
__device__ __forceinline__ float warp_reduce8(float v0, float v1, float v2, float v3, float v4, float v5, float v6,
                                              float v7, int lane) {
  // step 1 (xor 16): keep 4 of 8
  const bool hi16 = lane & 16;
  float a0 = hi16 ? v4 : v0, a1 = hi16 ? v5 : v1, a2 = hi16 ? v6 : v2, a3 = hi16 ? v7 : v3;
  float s0 = hi16 ? v0 : v4, s1 = hi16 ? v1 : v5, s2 = hi16 ? v2 : v6, s3 = hi16 ? v3 : v7;
  a0 += __shfl_xor_sync(kFull, s0, 16); a1 += __shfl_xor_sync(kFull, s1, 16);
  a2 += __shfl_xor_sync(kFull, s2, 16); a3 += __shfl_xor_sync(kFull, s3, 16);
  // step 2 (xor 8): keep 2 of 4
  const bool hi8 = lane & 8;
  float b0 = hi8 ? a2 : a0, b1 = hi8 ? a3 : a1;
  float t0 = hi8 ? a0 : a2, t1 = hi8 ? a1 : a3;
  b0 += __shfl_xor_sync(kFull, t0, 8); b1 += __shfl_xor_sync(kFull, t1, 8);
  // step 3 (xor 4): keep 1 of 2
  const bool hi4 = lane & 4;
  float c0 = hi4 ? b1 : b0;
  const float u0 = hi4 ? b0 : b1;
  c0 += __shfl_xor_sync(kFull, u0, 4);
  // steps 4,5: finish within the quad
  c0 += __shfl_xor_sync(kFull, c0, 2);
  c0 += __shfl_xor_sync(kFull, c0, 1);
  return c0;   // value index = (bit4?4:0) + (bit3?2:0) + (bit2?1:0)
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}


template <bool kDepthGrad>
__global__ void __launch_bounds__(kThreads) render_bwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const float4* __restrict__ rec,
    const float* __restrict__ views, const uint32_t* __restrict__ status, int P, int H, int W, int gx, int ntiles,
    const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor,
    const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha_out, float* __restrict__ dL_dscreen) {
  if (status[2]) return;
  __shared__ float4 sA[kThreads], sB[kThreads], sC[kThreads];
  __shared__ uint32_t sId[kThreads];
  const int tile = blockIdx.x, v = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = tile % gx, tile_y = tile / gx;
  const int x0 = tile_x * 16 + (warp & 1) * 8, y0 = tile_y * 16 + (warp >> 1) * 4;
  const int px = x0 + (lane & 7), py = y0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float fx0 = (float)x0, fx1 = (float)(x0 + 7), fy0 = (float)y0, fy1 = (float)(y0 + 3);
  const uint2 range = ranges[(size_t)v * ntiles + tile];
  const int total = (int)(range.y - range.x);
  if (total <= 0) return;
  const float4* __restrict__ rec_v = rec + (size_t)v * P * 3;
  const size_t HW = (size_t)H * W, pix = (size_t)py * W + px;
  const float* bg = views + (size_t)v * kViewFloats + 35;

  const float T_final = inside ? final_T[(size_t)v * HW + pix] : 0.f;
  float T = T_final;
  const int last = inside ? (int)n_contrib[(size_t)v * HW + pix] : 0;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f, dLd = 0.f;
  if (inside) {
    const float* g = dL_dcolor + (size_t)v * 3 * HW;
    dLp0 = g[pix]; dLp1 = g[HW + pix]; dLp2 = g[2 * HW + pix];
    if (kDepthGrad) dLd = dL_ddepth[(size_t)v * HW + pix];
  }
  // out_color = C + T_final * bg and out_alpha = 1 - T_final: both reach alpha_j only through T_final
  // (d T_final / d alpha_j = -T_final / (1 - alpha_j)), so a gradient on the alpha output folds into the background term
  float bg_dot = bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2;
  if (dL_dalpha_out != nullptr && inside) bg_dot -= dL_dalpha_out[(size_t)v * HW + pix];
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f, last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ld = 0.f;
  const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;
  // highest list position (1-based) any pixel of this warp contributed to
  const int warp_last = __reduce_max_sync(kFull, last);

  // walk the tile's list back to front in batches of 256 (batch 0 holds the LAST instances)
  for (int done_cnt = 0; done_cnt < total; done_cnt += kThreads) {
    const int n = min((int)kThreads, total - done_cnt);
    // staged slot k (0..n-1) holds list position pos = total-1-done_cnt-k  (0-based, front=0)
    __syncthreads();
    if (tid < n) {
      const uint32_t id = point_list[range.x + (uint32_t)(total - 1 - done_cnt - tid)];
      const float4* r = rec_v + 3 * (size_t)id;
      sA[tid] = __ldg(r); sB[tid] = __ldg(r + 1); sC[tid] = __ldg(r + 2);
      sId[tid] = id;
    }
    __syncthreads();
    const int pos_hi = total - 1 - done_cnt;          // list position of slot 0
    if (pos_hi - (n - 1) < warp_last) {               // some slot in this batch can matter to this warp
      for (int g = 0; g < n; g += 32) {
        const int j = g + lane;
        bool hit = false;
        if (j < n) {
          const float4 a = sA[j];
          const float4 c = sC[j];
          hit = (a.x - c.z <= fx1) && (a.x + c.z >= fx0) && (a.y - c.w <= fy1) && (a.y + c.w >= fy0) &&
                (pos_hi - j < warp_last);
        }
        unsigned m = __ballot_sync(kFull, hit);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int jj = g + b;
          const int pos = pos_hi - jj;                // 0-based list position; contributor index = pos+1
          const float4 a = sA[jj], bb = sB[jj], c = sC[jj];
          float g_mx = 0.f, g_my = 0.f, g_cx = 0.f, g_cy = 0.f, g_cw = 0.f, g_op = 0.f, g_r = 0.f, g_g = 0.f, g_b = 0.f, g_d = 0.f;
          if (pos < last) {
            const float dx = a.x - pxf, dy = a.y - pyf;
            // the SAME expressions as the forward pass (coefficients pre-multiplied by log2(e), MUFU.EX2): alpha and the
            // alpha >= 1/255 decision are bit-identical to the ones that produced final_T / n_contrib
            const float ca = (-0.5f * a.z) * kLog2e, cb = (-a.w) * kLog2e, cc = (-0.5f * bb.x) * kLog2e;
            const float t = fmaf(cb, dy, ca * dx);
            const float power = fmaf(cc * dy, dy, t * dx);
            if (power <= 0.f) {
              const float G = ex2_approx(power);
              const float alpha = fminf(0.99f, bb.y * G);
              if (alpha >= 1.f / 255.f) {
                const float r1ma = __frcp_rn(1.f - alpha);     // one correctly rounded reciprocal serves both divisions
                T = T * r1ma;
                const float w = alpha * T;
                float dL_dalpha;
                acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0; lc0 = bb.z;
                acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1; lc1 = bb.w;
                acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2; lc2 = c.x;
                dL_dalpha = (bb.z - acc0) * dLp0 + (bb.w - acc1) * dLp1 + (c.x - acc2) * dLp2;
                g_r = w * dLp0; g_g = w * dLp1; g_b = w * dLp2;
                if (kDepthGrad) {
                  accd = last_alpha * ld + (1.f - last_alpha) * accd; ld = c.y;
                  dL_dalpha += (c.y - accd) * dLd;
                  g_d = w * dLd;
                }
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final * r1ma) * bg_dot;
                const float dL_dG = bb.y * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                g_mx = dL_dG * (-gdx * a.z - gdy * a.w) * ddelx_dx;
                g_my = dL_dG * (-gdy * bb.x - gdx * a.w) * ddely_dy;
                g_cx = -0.5f * gdx * dx * dL_dG;
                g_cy = -0.5f * gdx * dy * dL_dG;
                g_cw = -0.5f * gdy * dy * dL_dG;
                g_op = G * dL_dalpha;
              }
            }
          }
          const float r8 = warp_reduce8(g_mx, g_my, g_cx, g_cy, g_cw, g_op, g_r, g_g, lane);
          const float rb = warp_sum(g_b);
          // ONE reduction instruction per walked instance: lanes 0,4,..,28 carry components 0..7, lane 1 component 8 (b),
          // lane 2 component 9 (depth) -> ten neighbouring floats of the instance's 48-byte gradient record, i.e. two
          // sectors for the L2 atomic unit.  The first version accumulated in shared memory (two CAS loops per walked
          // instance: fp32 shared atomics are compare-and-swap on this part) and flushed per batch behind two more block
          // barriers: ~190 instructions per walked instance and 30 % of the stall samples on the barrier (ncu r1f).
          float val = r8;
          int comp = (lane & 3) == 0 ? (lane >> 2) : -1;
          if (lane == 1) { val = rb; comp = 8; }
          if (kDepthGrad) {
            const float rd = warp_sum(g_d);
            if (lane == 2) { val = rd; comp = 9; }
          }
          if (comp >= 0 && val != 0.f) atomicAdd(dL_dscreen + ((size_t)v * P + sId[jj]) * 12 + comp, val);
        }
      }
    }
  }
}

int launch_render_fwd(const FsRasterFwdArgs& a, cudaStream_t s) {
  const int gx = tiles_x(a.W), gy = tiles_y(a.H);
  if (a.stages & FS_STAGE_RENDER_PACKED) {      // two pixels per lane on the packed fp32 pipe (measured 136 vs 129 us: not the default)
    render_fwd2_kernel<<<gx * gy * a.V, kThreads2, kSortSmemKeys2 * 8, s>>>(
        reinterpret_cast<const uint2*>(a.ranges), a.tile_count /* holds the tile order after binning */,
        reinterpret_cast<unsigned long long*>(a.keybuf), a.point_list, reinterpret_cast<const float4*>(a.rec), a.views, a.status, a.P,
        a.H, a.W, gx, gx * gy, a.out_color, a.out_depth, a.final_T, a.n_contrib);
    return check_cuda(cudaGetLastError(), "render_fwd2_kernel");
  }
  const bool binned = use_bins(a);
  render_fwd_kernel<<<gx * gy * a.V, kThreads, kSortSmemKeys * 8, s>>>(
      reinterpret_cast<const uint2*>(a.ranges), a.tile_count /* holds the tile order after binning */,
      reinterpret_cast<unsigned long long*>(a.keybuf), a.point_list, reinterpret_cast<const float4*>(a.rec), a.views, a.status, a.P,
      a.H, a.W, gx, gx * gy, a.out_color, a.out_depth, a.final_T, a.n_contrib,
      binned ? reinterpret_cast<const unsigned long long*>(a.bins) : nullptr, a.bin_cap, a.tile_cursor + (size_t)a.V * gx * gy);
  return check_cuda(cudaGetLastError(), "render_fwd_kernel");
}

int launch_render_bwd(const FsRasterBwdArgs& a, cudaStream_t s) {
  const int gx = tiles_x(a.W), gy = tiles_y(a.H);
  int rc;
  if ((rc = check_cuda(cudaMemsetAsync(a.dL_dscreen, 0, (size_t)a.V * a.P * 12 * sizeof(float), s), "memset dL_dscreen"))) return rc;
  dim3 grid(gx * gy, a.V);
  if (a.has_depth_grad && a.dL_ddepth)
    render_bwd_kernel<true><<<grid, kThreads, 0, s>>>(reinterpret_cast<const uint2*>(a.ranges), a.point_list,
                                                       reinterpret_cast<const float4*>(a.rec), a.views, a.status, a.P, a.H, a.W,
                                                       gx, gx * gy, a.final_T, a.n_contrib, a.dL_dcolor, a.dL_ddepth, a.dL_dalpha, a.dL_dscreen);
  else
    render_bwd_kernel<false><<<grid, kThreads, 0, s>>>(reinterpret_cast<const uint2*>(a.ranges), a.point_list,
                                                        reinterpret_cast<const float4*>(a.rec), a.views, a.status, a.P, a.H, a.W,
                                                        gx, gx * gy, a.final_T, a.n_contrib, a.dL_dcolor, a.dL_ddepth, a.dL_dalpha, a.dL_dscreen);
  return check_cuda(cudaGetLastError(), "render_bwd_kernel");
}

}  // namespace fs
