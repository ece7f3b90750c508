// cost_volume.cu -- fused plane-sweep feature volume.
//
// Replaces AVGFeatureVolumeManager.build_cost_volume
// (/root/reference/src/model/encoder/modules/cost_volume.py:429-619): per depth plane the reference
// runs BackprojectDepth -> Project3D -> F.grid_sample (materialising K warped copies of the source
// features) -> dot / masked means -> 3-layer MLP, ~15 launches x D planes.  Here one kernel computes
//   out[b,d,v,u] = MLP([mean_k valid warped features (48), mean_k dot (1)])
// directly: every thread owns one reference pixel, keeps its 48 reference features in registers and
// walks the planes of its chunk; per (plane, source) it evaluates the homography, gathers the four
// bilinear taps of the 48 channels straight from the NCHW source map (neighbouring threads read
// neighbouring texels: coalesced sectors, L1/L2 resident) and feeds the 49-vector to the MLP whose
// weights sit transposed in shared memory (broadcast reads).  Nothing but the [B,D,H,W] result
// is written: algorithmic HBM bytes per reference view = (1+K)*C*H'W'*4 + D*H'W'*4.
//
// Sampling arithmetic follows the reference chain exactly (geometry_utils.py:50-89,
// cost_volume.py:536-549, torch grid_sampler un-normalisation, align_corners=False):
//   X = z_d * (invK . (u+.5, v+.5, 1)) ; cam = P[:3,:3].X + P[:3,3] ; depth = cam.z + 1e-8
//   s = |cam.z| > 1e-8 ? 1/depth : 1 ; (px,py) = cam.xy * s ; g = 2*px*(1/W) - 1 ; ix = ((g+1)*W-1)/2
// validity of a source = (dot*[depth>0] != 0), the reference's exact-zero test (cost_volume.py:589-598).
#include "common.cuh"

namespace fs {

constexpr int kCvC = 48;        // matching channels (encoder_freesplat.py:160)
constexpr int kCvHid = 32;      // MLP hidden width (cost_volume.py:423-426)
constexpr int kCvThreads = 128;
constexpr int kCvIn = kCvC + 1;

struct CvSmem {
  float W0t[kCvIn][kCvHid];     // transposed: [input][output]
  float W1t[kCvHid][kCvHid];
  float W2[kCvHid];
  float b0[kCvHid];
  float b1[kCvHid];
  float b2;
  float proj[16 * 12];          // up to 16 sources x (3x4)
};

struct Taps {
  int o00, o01, o10, o11;       // element offsets inside one channel plane (clamped)
  float w00, w01, w10, w11;     // bilinear weights, zero for out-of-bounds taps
};

// returns false when no tap can be in bounds (or the point is behind the source camera)
__device__ __forceinline__ bool make_taps(const float* __restrict__ P, float X0, float X1, float X2, int H, int W,
                                          float uvx, float uvy, Taps& t) {
  const float cx = fmaf(P[2], X2, fmaf(P[1], X1, P[0] * X0)) + P[3];
  const float cy = fmaf(P[6], X2, fmaf(P[5], X1, P[4] * X0)) + P[7];
  const float cz = fmaf(P[10], X2, fmaf(P[9], X1, P[8] * X0)) + P[11];
  const float depth = cz + 1e-8f;
  if (!(depth > 0.f)) return false;
  const float s = fabsf(cz) > 1e-8f ? 1.0f / depth : 1.0f;
  const float px = cx * s, py = cy * s;
  const float gxn = 2.0f * px * uvx - 1.0f, gyn = 2.0f * py * uvy - 1.0f;
  const float ix = ((gxn + 1.0f) * (float)W - 1.0f) * 0.5f;
  const float iy = ((gyn + 1.0f) * (float)H - 1.0f) * 0.5f;
  if (!(ix >= -1.0f && ix < (float)W && iy >= -1.0f && iy < (float)H)) return false;
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix;
  const float wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
  const bool vx0 = x0 >= 0, vx1 = x1 <= W - 1, vy0 = y0 >= 0, vy1 = y1 <= H - 1;
  const int cx0 = max(x0, 0), cx1 = min(x1, W - 1), cy0 = max(y0, 0), cy1 = min(y1, H - 1);
  t.o00 = cy0 * W + cx0; t.o01 = cy0 * W + cx1; t.o10 = cy1 * W + cx0; t.o11 = cy1 * W + cx1;
  t.w00 = (vx0 && vy0) ? wx0 * wy0 : 0.f;
  t.w01 = (vx1 && vy0) ? wx1 * wy0 : 0.f;
  t.w10 = (vx0 && vy1) ? wx0 * wy1 : 0.f;
  t.w11 = (vx1 && vy1) ? wx1 * wy1 : 0.f;
  return true;
}

// Accumulates the warped features of the sources in `use_mask` into fsum and returns through
// zero_mask the sources whose dot product came out exactly 0 although geometrically valid.
__device__ __forceinline__ void gather_sources(const float* __restrict__ src_b, const float* __restrict__ proj, int K, int H, int W,
                                               size_t HW, float X0, float X1, float X2, float uvx, float uvy,
                                               const float (&cur)[kCvC], unsigned use_mask, float (&fsum)[kCvC], float& dsum,
                                               unsigned& geo_mask, unsigned& zero_mask) {
#pragma unroll
  for (int c = 0; c < kCvC; c++) fsum[c] = 0.f;
  dsum = 0.f; geo_mask = 0u; zero_mask = 0u;
  for (int k = 0; k < K; k++) {
    if (!((use_mask >> k) & 1u)) continue;
    Taps t;
    if (!make_taps(proj + 12 * k, X0, X1, X2, H, W, uvx, uvy, t)) continue;
    geo_mask |= 1u << k;
    const float* __restrict__ s = src_b + (size_t)k * kCvC * HW;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < kCvC; c++) {
      const float* __restrict__ sc = s + (size_t)c * HW;
      const float w = fmaf(t.w11, __ldg(sc + t.o11), fmaf(t.w10, __ldg(sc + t.o10), fmaf(t.w01, __ldg(sc + t.o01), t.w00 * __ldg(sc + t.o00))));
      dot = fmaf(w, cur[c], dot);
      fsum[c] += w;
    }
    if (dot == 0.f) zero_mask |= 1u << k;
    dsum += dot;
  }
}

// ---- channel-packed source maps ---------------------------------------------------------------
// The NCHW gather costs one 32-bit load + its address arithmetic per (tap, channel): ~4300 instructions per
// warp and voxel row (ncu r1).  cv_pack_kernel re-lays every source map as [C/4][H*W][4] once per call, so a tap
// fetches four channels with ONE 16-byte load (neighbouring lanes -> neighbouring 16-byte words: 512 contiguous
// bytes per warp instruction): 4x fewer load and address instructions for the same FMA count.
__global__ void __launch_bounds__(256) cv_pack_kernel(const float* __restrict__ src, float4* __restrict__ dst, size_t HW, size_t total) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;            // over maps * 12 * HW
  if (i >= total) return;
  const size_t pix = i % HW, mg = i / HW;                              // mg = map*12 + group
  const float* s = src + (mg * 4) * HW + pix;                          // channel 4*group of that map
  dst[i] = make_float4(__ldg(s), __ldg(s + HW), __ldg(s + 2 * HW), __ldg(s + 3 * HW));
}

__device__ __forceinline__ void gather_sources_packed(const float4* __restrict__ srcp_b, const float* __restrict__ proj, int K, int H,
                                                      int W, size_t HW, float X0, float X1, float X2, float uvx, float uvy,
                                                      const float (&cur)[kCvC], unsigned use_mask, float (&fsum)[kCvC], float& dsum,
                                                      unsigned& geo_mask, unsigned& zero_mask) {
#pragma unroll
  for (int c = 0; c < kCvC; c++) fsum[c] = 0.f;
  dsum = 0.f; geo_mask = 0u; zero_mask = 0u;
  for (int k = 0; k < K; k++) {
    if (!((use_mask >> k) & 1u)) continue;
    Taps t;
    if (!make_taps(proj + 12 * k, X0, X1, X2, H, W, uvx, uvy, t)) continue;
    geo_mask |= 1u << k;
    const float4* __restrict__ s = srcp_b + (size_t)k * (kCvC / 4) * HW;
    float dot = 0.f;
#pragma unroll
    for (int g = 0; g < kCvC / 4; g++) {
      const float4* __restrict__ sg = s + (size_t)g * HW;
      const float4 a = __ldg(sg + t.o00), b = __ldg(sg + t.o01), c = __ldg(sg + t.o10), d = __ldg(sg + t.o11);
      const float w0 = fmaf(t.w11, d.x, fmaf(t.w10, c.x, fmaf(t.w01, b.x, t.w00 * a.x)));
      const float w1 = fmaf(t.w11, d.y, fmaf(t.w10, c.y, fmaf(t.w01, b.y, t.w00 * a.y)));
      const float w2 = fmaf(t.w11, d.z, fmaf(t.w10, c.z, fmaf(t.w01, b.z, t.w00 * a.z)));
      const float w3 = fmaf(t.w11, d.w, fmaf(t.w10, c.w, fmaf(t.w01, b.w, t.w00 * a.w)));
      dot = fmaf(w0, cur[4 * g], dot); dot = fmaf(w1, cur[4 * g + 1], dot);
      dot = fmaf(w2, cur[4 * g + 2], dot); dot = fmaf(w3, cur[4 * g + 3], dot);
      fsum[4 * g] += w0; fsum[4 * g + 1] += w1; fsum[4 * g + 2] += w2; fsum[4 * g + 3] += w3;
    }
    if (dot == 0.f) zero_mask |= 1u << k;
    dsum += dot;
  }
}

// Same loop body as gather_sources_packed for the sources [k0, k1), ACCUMULATING into fsum / dsum / the masks (the caller zeroes
// them): the software-pipelined forward kernel gathers one plane in two instalments, underneath the two MMA round trips of the
// previous plane.  Sources are still visited in ascending k, so the sums are bit-identical to the one-call version.
__device__ __forceinline__ void gather_accum_packed(const float4* __restrict__ srcp_b, const float* __restrict__ proj, int k0, int k1, int H,
                                                    int W, size_t HW, float X0, float X1, float X2, float uvx, float uvy,
                                                    const float (&cur)[kCvC], unsigned use_mask, float (&fsum)[kCvC], float& dsum,
                                                    unsigned& geo_mask, unsigned& zero_mask) {
  for (int k = k0; k < k1; k++) {
    if (!((use_mask >> k) & 1u)) continue;
    Taps t;
    if (!make_taps(proj + 12 * k, X0, X1, X2, H, W, uvx, uvy, t)) continue;
    geo_mask |= 1u << k;
    const float4* __restrict__ s = srcp_b + (size_t)k * (kCvC / 4) * HW;
    float dot = 0.f;
#pragma unroll
    for (int g = 0; g < kCvC / 4; g++) {
      const float4* __restrict__ sg = s + (size_t)g * HW;
      const float4 a = __ldg(sg + t.o00), b = __ldg(sg + t.o01), c = __ldg(sg + t.o10), d = __ldg(sg + t.o11);
      const float w0 = fmaf(t.w11, d.x, fmaf(t.w10, c.x, fmaf(t.w01, b.x, t.w00 * a.x)));
      const float w1 = fmaf(t.w11, d.y, fmaf(t.w10, c.y, fmaf(t.w01, b.y, t.w00 * a.y)));
      const float w2 = fmaf(t.w11, d.z, fmaf(t.w10, c.z, fmaf(t.w01, b.z, t.w00 * a.z)));
      const float w3 = fmaf(t.w11, d.w, fmaf(t.w10, c.w, fmaf(t.w01, b.w, t.w00 * a.w)));
      dot = fmaf(w0, cur[4 * g], dot); dot = fmaf(w1, cur[4 * g + 1], dot);
      dot = fmaf(w2, cur[4 * g + 2], dot); dot = fmaf(w3, cur[4 * g + 3], dot);
      fsum[4 * g] += w0; fsum[4 * g + 1] += w1; fsum[4 * g + 2] += w2; fsum[4 * g + 3] += w3;
    }
    if (dot == 0.f) zero_mask |= 1u << k;
    dsum += dot;
  }
}

// L1 prefetch (CCTL.PF1) of the two tap rows of every channel group of the sources in [k0, k1) at plane position X: the lanes of a
// warp cover a contiguous run of texels, so the 32 requests of one instruction fall into ~5 lines.
__device__ __forceinline__ void prefetch_taps_packed(const float4* __restrict__ srcp_b, const float* __restrict__ proj, int k0, int k1, int H,
                                                     int W, size_t HW, float X0, float X1, float X2, float uvx, float uvy) {
  for (int k = k0; k < k1; k++) {
    Taps t;
    if (!make_taps(proj + 12 * k, X0, X1, X2, H, W, uvx, uvy, t)) continue;
    const float4* __restrict__ s = srcp_b + (size_t)k * (kCvC / 4) * HW;
#pragma unroll
    for (int g = 0; g < kCvC / 4; g++) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(s + (size_t)g * HW + t.o00));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(s + (size_t)g * HW + t.o11));
    }
  }
}

// CTA -> 32x4 pixel patch (warp = one 32-pixel row segment, the 4 warps = 4 consecutive rows): rows y and y+1 share
// their bilinear tap rows, which lifts the L1 hit rate of the gather over a 128-pixel run of a single row.
__device__ __forceinline__ bool patch_pixel(int block, int tid, int H, int W, int& u, int& v) {
  const int px = (W + 31) >> 5;
  const int by = block / px, bx = block - by * px;
  u = bx * 32 + (tid & 31); v = by * 4 + (tid >> 5);
  return u < W && v < H;
}
__host__ __device__ inline unsigned patch_blocks(int H, int W) { return (unsigned)(((W + 31) >> 5) * ((H + 3) >> 2)); }

__device__ __forceinline__ float leaky(float x) { return fmaxf(x, 0.01f * x); }   // == (x > 0 ? x : 0.01 x) for every x, one instruction less

__global__ void __launch_bounds__(kCvThreads) cost_volume_fwd_kernel(FsCostVolumeArgs a, int planes_per_block) {
  __shared__ CvSmem sm;
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int H = a.H, W = a.W, K = a.K;
  const size_t HW = (size_t)H * W;
  // ---- stage MLP weights (transposed) and this view's projection matrices ----
  {
    const float* w = a.mlp;
    for (int k = tid; k < kCvHid * kCvIn; k += kCvThreads) { const int o = k / kCvIn, i = k - o * kCvIn; sm.W0t[i][o] = w[k]; }
    w += kCvHid * kCvIn;
    for (int k = tid; k < kCvHid; k += kCvThreads) sm.b0[k] = w[k];
    w += kCvHid;
    for (int k = tid; k < kCvHid * kCvHid; k += kCvThreads) { const int o = k / kCvHid, i = k - o * kCvHid; sm.W1t[i][o] = w[k]; }
    w += kCvHid * kCvHid;
    for (int k = tid; k < kCvHid; k += kCvThreads) sm.b1[k] = w[k];
    w += kCvHid;
    for (int k = tid; k < kCvHid; k += kCvThreads) sm.W2[k] = w[k];
    w += kCvHid;
    if (tid == 0) sm.b2 = w[0];
    for (int k = tid; k < K * 12; k += kCvThreads) sm.proj[k] = a.proj[(size_t)b * K * 12 + k];
  }
  __syncthreads();
  int u, v;
  if (!patch_pixel(blockIdx.x, tid, H, W, u, v)) return;
  const int p = v * W + u;
  const float uvx = 1.0f / (float)W, uvy = 1.0f / (float)H;
  const float* ik = a.cur_invK + (size_t)b * 9;
  const float pu = (float)u + 0.5f, pv = (float)v + 0.5f;
  const float r0 = fmaf(ik[1], pv, ik[0] * pu) + ik[2];
  const float r1 = fmaf(ik[4], pv, ik[3] * pu) + ik[5];
  const float r2 = fmaf(ik[7], pv, ik[6] * pu) + ik[8];
  float cur[kCvC];
  {
    const float* cb = a.cur_feats + (size_t)b * kCvC * HW + p;
#pragma unroll
    for (int c = 0; c < kCvC; c++) cur[c] = __ldg(cb + (size_t)c * HW);
  }
  const float4* src_b = reinterpret_cast<const float4*>(a.src_packed) + (size_t)b * K * (kCvC / 4) * HW;
  const unsigned all = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
  const int d0 = blockIdx.y * planes_per_block, d1 = min(a.D, d0 + planes_per_block);
  for (int d = d0; d < d1; d++) {
    const float zd = __ldg(a.planes + d);
    const float X0 = zd * r0, X1 = zd * r1, X2 = zd * r2;
    float x[kCvC];
    float dsum;
    unsigned geo, zero;
    gather_sources_packed(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, all, x, dsum, geo, zero);
    unsigned valid = geo & ~zero;
    if (zero) {   // measure-zero case: redo with the exact set of valid sources (keeps the reference's dot != 0 rule)
      float ds2; unsigned g2, z2;
      gather_sources_packed(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, valid, x, ds2, g2, z2);
    }
    const float rn = 1.0f / ((float)__popc(valid) + 1e-8f);
    // ---- MLP 49 -> 32 -> 32 -> 1, LeakyReLU(0.01) ----
    float h1[kCvHid];
#pragma unroll
    for (int o = 0; o < kCvHid; o++) h1[o] = sm.b0[o];
#pragma unroll
    for (int i = 0; i < kCvC; i++) {
      const float xi = x[i] * rn;
#pragma unroll
      for (int o = 0; o < kCvHid; o++) h1[o] = fmaf(sm.W0t[i][o], xi, h1[o]);
    }
    {
      const float xi = dsum * rn;
#pragma unroll
      for (int o = 0; o < kCvHid; o++) h1[o] = fmaf(sm.W0t[kCvC][o], xi, h1[o]);
    }
    float h2[kCvHid];
#pragma unroll
    for (int o = 0; o < kCvHid; o++) h2[o] = sm.b1[o];
#pragma unroll
    for (int i = 0; i < kCvHid; i++) {
      const float hi = leaky(h1[i]);
#pragma unroll
      for (int o = 0; o < kCvHid; o++) h2[o] = fmaf(sm.W1t[i][o], hi, h2[o]);
    }
    float y = sm.b2;
#pragma unroll
    for (int i = 0; i < kCvHid; i++) y = fmaf(sm.W2[i], leaky(h2[i]), y);
    a.out[((size_t)b * a.D + d) * HW + p] = y;
  }
}

// =========================================================================== tensor-core forward
// Same gather, but the 49->32->32 contraction runs on the 5th-generation tensor cores:
//   * each CTA owns 128 rows (= 128 reference pixels at one plane) -> UMMA M = 128, N = 32, K = 8 (tf32);
//   * thread r writes row r of the A operand into shared memory in the canonical K-major /
//     SWIZZLE_NONE core-matrix layout ([k-step][row-group of 8][k-chunk of 16 B][row][16 B]:
//     SBO = 256 B, LBO = 128 B), weights (B operand, nn.Linear layout is already K-major) likewise;
//   * fp32 accuracy by the 3xTF32 split: x = hi + lo (both tf32), D = Ahi*Bhi + Ahi*Blo + Alo*Bhi accumulated in
//     TMEM (fp32), |error| ~ 2^-22: inside the 1e-4 parity budget (tests/test_cost_volume_gpu.py);
//   * one elected thread issues the tcgen05.mma batch and commits to an mbarrier; every warp then reads ITS
//     32 TMEM lanes with tcgen05.ld.32x32b.x32 -- lane = row, so "thread = row" holds on both sides;
//   * layer 2 re-uses the A buffer; the last layer (32 -> 1) stays on the CUDA cores (32 FMA per row).
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
// activation-side tf32 split by TRUNCATION: hi = top 19 bits of x, lo = top 19 bits of (x - hi) (the subtraction is exact).
// cvt.rna.tf32.f32 is three instructions on sm_100a (FSETP + IADD + LOP3): the rounded split cost 7 instructions per element
// and ~15 % of the cost-volume / GRU kernels; the truncated one costs 3.  x = hi + lo holds to 2^-20 |x| (2^-22 rounded):
// the dropped lo*lo term and the split error stay ~1e-6 relative, inside the 1e-4 budget.  Weights keep the rounded split.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  constexpr uint64_t kLbo = 128 >> 4, kSbo = 256 >> 4;
  return (uint64_t)((saddr >> 4) & 0x3fffu) | (kLbo << 16) | (kSbo << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, N = 32, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
// One elected lane of a converged warp (elect.sync): inside `if (warp == 0 && elect_one())` ptxas knows that exactly one lane runs,
// keeps descriptors and addresses in uniform registers and emits 1-2 instructions per tcgen05.mma; under `if (tid == 0)` every MMA
// was wrapped in an ELECT / R2UR / BRA.U.ANY uniformisation loop plus the descriptor arithmetic: 9-17 instructions per MMA, all on
// the one warp every other warp of the CTA then waits for at the next barrier (cuobjdump -sass, r2).
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0u;
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

constexpr int kK1 = 56, kSteps1 = kK1 / 8, kSteps2 = kCvHid / 8;   // layer-1 K padded 49 -> 56
constexpr int kABytesPerStep = 128 / 8 * 256;                      // 16 row groups x 256 B = 4096
constexpr int kBBytesPerStep = kCvHid / 8 * 256;                   // 4 row groups x 256 B = 1024

struct __align__(128) Smem {
  unsigned char A_hi[kSteps1 * kABytesPerStep];     // 28 KB (layer 2 re-uses the first 16 KB)
  unsigned char A_lo[kSteps1 * kABytesPerStep];
  unsigned char B0_hi[kSteps1 * kBBytesPerStep];    // 7 KB
  unsigned char B0_lo[kSteps1 * kBBytesPerStep];
  unsigned char B1_hi[kSteps2 * kBBytesPerStep];    // 4 KB
  unsigned char B1_lo[kSteps2 * kBBytesPerStep];
  float b0[kCvHid], b1[kCvHid], W2[kCvHid];
  float b2;
  float proj[16 * 12];
  unsigned long long bar;
  uint32_t tmem_base;
};

// byte offset of element (row, k) inside an operand tile (k counted in tf32 elements)
__device__ __forceinline__ uint32_t op_off(int row, int k, int bytes_per_step) {
  return (uint32_t)((k >> 3) * bytes_per_step + (row >> 3) * 256 + ((k >> 2) & 1) * 128 + (row & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ void split_store4(unsigned char* hi, unsigned char* lo, uint32_t off, float v0, float v1, float v2, float v3) {
  uint4 h, l;
  split_tf32(v0, h.x, l.x); split_tf32(v1, h.y, l.y); split_tf32(v2, h.z, l.z); split_tf32(v3, h.w, l.w);
  *reinterpret_cast<uint4*>(hi + off) = h;
  *reinterpret_cast<uint4*>(lo + off) = l;
}

// MODE 0: gather -> layer 1 -> layer 2 strictly in sequence per plane (round 1).
// MODE 1: software-pipelined over the planes: the gather of plane d+1 is issued in two instalments (sources [0, K/2) and
//         [K/2, K)) right behind the tcgen05.commit of layer 1 resp. layer 2 of plane d, so the two MMA -> TMEM -> register round
//         trips of a CTA are covered by its own loads instead of only by the other CTA of the SM.  Same arithmetic in the same
//         order: bit-identical output.
// MODE 2: MODE 1 + L1 prefetch of plane d+2's tap lines.
template <int MODE>
__global__ void __launch_bounds__(kCvThreads) cost_volume_fwd_tc_kernel(FsCostVolumeArgs a, int planes_per_block) {
  extern __shared__ __align__(128) unsigned char tc_smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(tc_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z;
  const int H = a.H, W = a.W, K = a.K;
  const size_t HW = (size_t)H * W;
  // ---- one-time setup: weights as tf32 hi/lo B tiles, biases, projections, mbarrier, TMEM ----
  {
    const float* w0 = a.mlp;                                // [32][49]
    const float* pb0 = w0 + kCvHid * kCvIn;
    const float* w1 = pb0 + kCvHid;                         // [32][32]
    // every load of the thread is issued before the first split / store (one exposed L2 round trip instead of 22)
    float v0[kCvHid * kK1 / kCvThreads], v1[kCvHid * kCvHid / kCvThreads];
#pragma unroll
    for (int q = 0; q < kCvHid * kK1 / kCvThreads; q++) {
      const int e = tid + q * kCvThreads, n = e / kK1, k = e - n * kK1;
      v0[q] = k < kCvIn ? __ldg(w0 + n * kCvIn + k) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < kCvHid * kCvHid / kCvThreads; q++) v1[q] = __ldg(w1 + tid + q * kCvThreads);
#pragma unroll
    for (int q = 0; q < kCvHid * kK1 / kCvThreads; q++) {
      const int e = tid + q * kCvThreads, n = e / kK1, k = e - n * kK1;
      const uint32_t hi = to_tf32(v0[q]), lo = to_tf32(v0[q] - __uint_as_float(hi));
      const uint32_t off = op_off(n, k, kBBytesPerStep);
      *reinterpret_cast<uint32_t*>(sm.B0_hi + off) = hi; *reinterpret_cast<uint32_t*>(sm.B0_lo + off) = lo;
    }
#pragma unroll
    for (int q = 0; q < kCvHid * kCvHid / kCvThreads; q++) {
      const int e = tid + q * kCvThreads, n = e / kCvHid, k = e - n * kCvHid;
      const uint32_t hi = to_tf32(v1[q]), lo = to_tf32(v1[q] - __uint_as_float(hi));
      const uint32_t off = op_off(n, k, kBBytesPerStep);
      *reinterpret_cast<uint32_t*>(sm.B1_hi + off) = hi; *reinterpret_cast<uint32_t*>(sm.B1_lo + off) = lo;
    }
    const float* pb1 = w1 + kCvHid * kCvHid;
    const float* w2 = pb1 + kCvHid;
    for (int k = tid; k < kCvHid; k += kCvThreads) { sm.b0[k] = pb0[k]; sm.b1[k] = pb1[k]; sm.W2[k] = w2[k]; }
    if (tid == 0) sm.b2 = w2[kCvHid];
    for (int k = tid; k < K * 12; k += kCvThreads) sm.proj[k] = a.proj[(size_t)b * K * 12 + k];
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(64u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const uint32_t tmem = sm.tmem_base;
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);     // this warp's 32 TMEM lanes
  const uint32_t bar = smem_u32(&sm.bar);
  const uint32_t aA_hi = smem_u32(sm.A_hi), aA_lo = smem_u32(sm.A_lo);
  const uint32_t aB0_hi = smem_u32(sm.B0_hi), aB0_lo = smem_u32(sm.B0_lo), aB1_hi = smem_u32(sm.B1_hi), aB1_lo = smem_u32(sm.B1_lo);
  uint32_t phase = 0;

  int u, v;
  const bool active = patch_pixel(blockIdx.x, tid, H, W, u, v);
  if (!active) { u = 0; v = 0; }
  const int p = v * W + u;
  const int pc = p;
  const float uvx = 1.0f / (float)W, uvy = 1.0f / (float)H;
  const float* ik = a.cur_invK + (size_t)b * 9;
  const float pu = (float)u + 0.5f, pv = (float)v + 0.5f;
  const float r0 = fmaf(ik[1], pv, ik[0] * pu) + ik[2];
  const float r1 = fmaf(ik[4], pv, ik[3] * pu) + ik[5];
  const float r2 = fmaf(ik[7], pv, ik[6] * pu) + ik[8];
  float cur[kCvC];
  {
    const float* cb = a.cur_feats + (size_t)b * kCvC * HW + pc;
#pragma unroll
    for (int c = 0; c < kCvC; c++) cur[c] = __ldg(cb + (size_t)c * HW);
  }
  const float4* src_b = reinterpret_cast<const float4*>(a.src_packed) + (size_t)b * K * (kCvC / 4) * HW;
  const unsigned all = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
  const int d0 = blockIdx.y * planes_per_block, d1 = min(a.D, d0 + planes_per_block);
  const uint32_t my_off = (uint32_t)((tid >> 3) * 256 + (tid & 7) * 16);   // row part of op_off

  if constexpr (MODE == 0) {
  for (int d = d0; d < d1; d++) {
    const float zd = __ldg(a.planes + d);
    const float X0 = zd * r0, X1 = zd * r1, X2 = zd * r2;
    {
      float x[kCvC];
      float dsum;
      unsigned geo, zero;
      gather_sources_packed(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, all, x, dsum, geo, zero);
      const unsigned valid = geo & ~zero;
      if (zero) { float ds2; unsigned g2, z2; gather_sources_packed(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, valid, x, ds2, g2, z2); }
      const float rn = 1.0f / ((float)__popc(valid) + 1e-8f);
      // ---- A operand of layer 1: row tid, k = 0..55 (x*rn, dot*rn, zero padding), hi/lo tf32 ----
#pragma unroll
      for (int q = 0; q < kCvC / 4; q++) {
        const uint32_t off = (uint32_t)((q >> 1) * kABytesPerStep + (q & 1) * 128) + my_off;
        split_store4(sm.A_hi, sm.A_lo, off, x[4 * q] * rn, x[4 * q + 1] * rn, x[4 * q + 2] * rn, x[4 * q + 3] * rn);
      }
      split_store4(sm.A_hi, sm.A_lo, (uint32_t)(6 * kABytesPerStep) + my_off, dsum * rn, 0.f, 0.f, 0.f);
      split_store4(sm.A_hi, sm.A_lo, (uint32_t)(6 * kABytesPerStep + 128) + my_off, 0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < kSteps1; s++) {
        const uint64_t dah = make_desc(aA_hi + s * kABytesPerStep), dal = make_desc(aA_lo + s * kABytesPerStep);
        const uint64_t dbh = make_desc(aB0_hi + s * kBBytesPerStep), dbl = make_desc(aB0_lo + s * kBBytesPerStep);
        mma_tf32(tmem, dal, dbh, s > 0 ? 1u : 0u);
        mma_tf32(tmem, dah, dbl, 1u);
        mma_tf32(tmem, dah, dbh, 1u);
      }
      commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      float h[kCvHid];
      tmem_ld32(t_row, h);
      // ---- A operand of layer 2: leaky(h + b0), K = 32 ----
#pragma unroll
      for (int q = 0; q < kCvHid / 4; q++) {
        const uint32_t off = (uint32_t)((q >> 1) * kABytesPerStep + (q & 1) * 128) + my_off;
        split_store4(sm.A_hi, sm.A_lo, off, leaky(h[4 * q] + sm.b0[4 * q]), leaky(h[4 * q + 1] + sm.b0[4 * q + 1]),
                     leaky(h[4 * q + 2] + sm.b0[4 * q + 2]), leaky(h[4 * q + 3] + sm.b0[4 * q + 3]));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < kSteps2; s++) {
        const uint64_t dah = make_desc(aA_hi + s * kABytesPerStep), dal = make_desc(aA_lo + s * kABytesPerStep);
        const uint64_t dbh = make_desc(aB1_hi + s * kBBytesPerStep), dbl = make_desc(aB1_lo + s * kBBytesPerStep);
        mma_tf32(tmem + 32u, dal, dbh, s > 0 ? 1u : 0u);
        mma_tf32(tmem + 32u, dah, dbl, 1u);
        mma_tf32(tmem + 32u, dah, dbh, 1u);
      }
      commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      float h[kCvHid];
      tmem_ld32(t_row + 32u, h);
      float y = sm.b2;
#pragma unroll
      for (int i = 0; i < kCvHid; i++) y = fmaf(sm.W2[i], leaky(h[i] + sm.b1[i]), y);
      if (active) a.out[((size_t)b * a.D + d) * HW + p] = y;
    }
  }
  } else {
  // ---------------------------------------------------------------- software-pipelined plane loop
  float x[kCvC];
  float dsum = 0.f, rn = 0.f;
  unsigned geo = 0u, zero = 0u;
  const int Ka = (K + 1) >> 1;
  // exact `dot != 0` rule + 1 / n for the plane whose sums are in x / dsum
  auto finalize = [&](float X0, float X1, float X2) {
    const unsigned valid = geo & ~zero;
    if (zero) {
#pragma unroll
      for (int c = 0; c < kCvC; c++) x[c] = 0.f;
      float ds2 = 0.f; unsigned g2 = 0u, z2 = 0u;
      gather_accum_packed(src_b, sm.proj, 0, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, valid, x, ds2, g2, z2);
    }
    rn = 1.0f / ((float)__popc(valid) + 1e-8f);
  };
  if (d0 < d1) {                                             // prologue: the first plane, gathered in one go
    const float zd = __ldg(a.planes + d0);
    const float X0 = zd * r0, X1 = zd * r1, X2 = zd * r2;
#pragma unroll
    for (int c = 0; c < kCvC; c++) x[c] = 0.f;
    gather_accum_packed(src_b, sm.proj, 0, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, all, x, dsum, geo, zero);
    finalize(X0, X1, X2);
  }
  for (int d = d0; d < d1; d++) {
    // ---- A operand of layer 1 from the sums of plane d (x is dead afterwards) ----
#pragma unroll
    for (int q = 0; q < kCvC / 4; q++) {
      const uint32_t off = (uint32_t)((q >> 1) * kABytesPerStep + (q & 1) * 128) + my_off;
      split_store4(sm.A_hi, sm.A_lo, off, x[4 * q] * rn, x[4 * q + 1] * rn, x[4 * q + 2] * rn, x[4 * q + 3] * rn);
    }
    split_store4(sm.A_hi, sm.A_lo, (uint32_t)(6 * kABytesPerStep) + my_off, dsum * rn, 0.f, 0.f, 0.f);
    split_store4(sm.A_hi, sm.A_lo, (uint32_t)(6 * kABytesPerStep + 128) + my_off, 0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < kSteps1; s++) {
        const uint64_t dah = make_desc(aA_hi + s * kABytesPerStep), dal = make_desc(aA_lo + s * kABytesPerStep);
        const uint64_t dbh = make_desc(aB0_hi + s * kBBytesPerStep), dbl = make_desc(aB0_lo + s * kBBytesPerStep);
        mma_tf32(tmem, dal, dbh, s > 0 ? 1u : 0u);
        mma_tf32(tmem, dah, dbl, 1u);
        mma_tf32(tmem, dah, dbh, 1u);
      }
      commit(bar);
    }
    // ---- first instalment of plane d+1, underneath the layer-1 MMAs ----
    const bool next = d + 1 < d1;
    float nX0 = 0.f, nX1 = 0.f, nX2 = 0.f;
    if (next) {
      const float zd = __ldg(a.planes + d + 1);
      nX0 = zd * r0; nX1 = zd * r1; nX2 = zd * r2;
#pragma unroll
      for (int c = 0; c < kCvC; c++) x[c] = 0.f;
      dsum = 0.f; geo = 0u; zero = 0u;
      gather_accum_packed(src_b, sm.proj, 0, Ka, H, W, HW, nX0, nX1, nX2, uvx, uvy, cur, all, x, dsum, geo, zero);
    }
    mbar_wait(bar, phase); phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      float h[kCvHid];
      tmem_ld32(t_row, h);
#pragma unroll
      for (int q = 0; q < kCvHid / 4; q++) {
        const uint32_t off = (uint32_t)((q >> 1) * kABytesPerStep + (q & 1) * 128) + my_off;
        split_store4(sm.A_hi, sm.A_lo, off, leaky(h[4 * q] + sm.b0[4 * q]), leaky(h[4 * q + 1] + sm.b0[4 * q + 1]),
                     leaky(h[4 * q + 2] + sm.b0[4 * q + 2]), leaky(h[4 * q + 3] + sm.b0[4 * q + 3]));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < kSteps2; s++) {
        const uint64_t dah = make_desc(aA_hi + s * kABytesPerStep), dal = make_desc(aA_lo + s * kABytesPerStep);
        const uint64_t dbh = make_desc(aB1_hi + s * kBBytesPerStep), dbl = make_desc(aB1_lo + s * kBBytesPerStep);
        mma_tf32(tmem + 32u, dal, dbh, s > 0 ? 1u : 0u);
        mma_tf32(tmem + 32u, dah, dbl, 1u);
        mma_tf32(tmem + 32u, dah, dbh, 1u);
      }
      commit(bar);
    }
    // ---- second instalment of plane d+1, underneath the layer-2 MMAs ----
    if (next) {
      gather_accum_packed(src_b, sm.proj, Ka, K, H, W, HW, nX0, nX1, nX2, uvx, uvy, cur, all, x, dsum, geo, zero);
      finalize(nX0, nX1, nX2);
      if constexpr (MODE == 2) {
        if (d + 2 < d1) {
          const float zd = __ldg(a.planes + d + 2);
          prefetch_taps_packed(src_b, sm.proj, 0, K, H, W, HW, zd * r0, zd * r1, zd * r2, uvx, uvy);
        }
      }
    }
    mbar_wait(bar, phase); phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      float h[kCvHid];
      tmem_ld32(t_row + 32u, h);
      float y = sm.b2;
#pragma unroll
      for (int i = 0; i < kCvHid; i++) y = fmaf(sm.W2[i], leaky(h[i] + sm.b1[i]), y);
      if (active) a.out[((size_t)b * a.D + d) * HW + p] = y;
    }
  }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

}  // namespace tc

static int launch_pack(const FsCostVolumeArgs& a, cudaStream_t s) {
  const size_t HW = (size_t)a.H * a.W;
  const size_t total = (size_t)a.B * a.K * (kCvC / 4) * HW;
  cv_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a.src_feats, reinterpret_cast<float4*>(a.src_packed), HW, total);
  return check_cuda(cudaGetLastError(), "cv_pack_kernel");
}

int launch_cost_volume_fwd(const FsCostVolumeArgs& a, cudaStream_t s) {
  const size_t HW = (size_t)a.H * a.W;
  const int ppb = 16;      // planes per CTA: 4 / 8 / 16 / 32 measured 1.59 / 1.44 / 1.39 / 1.40 ms (config 3)
  if (int rc = launch_pack(a, s)) return rc;
  (void)HW;
  dim3 grid(patch_blocks(a.H, a.W), (unsigned)((a.D + ppb - 1) / ppb), (unsigned)a.B);
  if (a.mlp_mode == 1) {   // fp32 CUDA-core MLP: validation path for the tensor-core kernel
    cost_volume_fwd_kernel<<<grid, kCvThreads, 0, s>>>(a, ppb);
    return check_cuda(cudaGetLastError(), "cost_volume_fwd_kernel");
  }
  const size_t smem = sizeof(tc::Smem) + 128;
  // mlp_mode 0 = pipelined (default), 2 = strictly sequential planes (round 1), 3 = pipelined + L1 prefetch
  auto kern = a.mlp_mode == 2 ? tc::cost_volume_fwd_tc_kernel<0> : (a.mlp_mode == 3 ? tc::cost_volume_fwd_tc_kernel<2> : tc::cost_volume_fwd_tc_kernel<1>);
  if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(cost_volume_fwd_tc_kernel)")) return rc;
  kern<<<grid, kCvThreads, smem, s>>>(a, ppb);
  return check_cuda(cudaGetLastError(), "cost_volume_fwd_tc_kernel");
}

}  // namespace fs

// =================================================================================== backward
// Gradients w.r.t. cur_feats, src_feats and the MLP parameters (poses / intrinsics get none in the
// reference's training path).  Per block: 128 reference pixels x a chunk of planes.
//   phase A (thread = pixel): recompute the forward row, back-propagate through the MLP in registers
//            -> dx[49]; scatter d(warped) to dL_dsrc with red.global.add (neighbouring lanes hit
//            neighbouring texels), accumulate dL_dcur in registers; park the row's (dz1, x, dz2, a1, g*a2, g)
//            in shared memory;
//   phase B (thread = (output o, input slice)): dW += sum over the 128 rows of the outer products,
//            operands broadcast from shared memory as float4, ~21 accumulators per thread kept in
//            registers across all planes; one atomicAdd per parameter and block at the end.
namespace fs {

constexpr int kRowStride = 196;                  // floats per parked row (16-byte aligned segments)
constexpr int kOffDz1 = 0, kOffX = 32, kOffDz2 = 84, kOffA1 = 116, kOffGa2 = 148, kOffG = 180;

struct CvBwdSmem {
  float W0t[kCvIn][kCvHid];
  float W1t[kCvHid][kCvHid];
  float W0[kCvHid][kCvIn + 3];                   // row-major copy for dx = W0^T dz1 (padded to 52)
  float W1[kCvHid][kCvHid];
  float W2[kCvHid];
  float b0[kCvHid];
  float b1[kCvHid];
  float b2;
  float proj[16 * 12];
};

// inverse of cv_pack_kernel: [maps][C/4][HW][4] -> NCHW
__global__ void __launch_bounds__(256) cv_unpack_kernel(const float4* __restrict__ src, float* __restrict__ dst, size_t HW, size_t total) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const size_t pix = i % HW, mg = i / HW;
  const float4 v = src[i];
  float* d = dst + (mg * 4) * HW + pix;
  d[0] = v.x; d[HW] = v.y; d[2 * HW] = v.z; d[3 * HW] = v.w;
}

// one 16-byte vector reduction (REDG.E.ADD.F32x4): four channels of one tap
__device__ __forceinline__ void red_add_v4(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// the same reduction under a predicate instead of a branch (48 per thread and plane in the tensor-core backward)
__device__ __forceinline__ void red_add_v4_if(bool on, float4* addr, float a, float b, float c, float d) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t"
      "}\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"((int)on) : "memory");
}

__device__ __forceinline__ float dleaky(float z) { return z > 0.f ? 1.f : 0.01f; }

__global__ void __launch_bounds__(kCvThreads) cost_volume_bwd_kernel(FsCostVolumeArgs a, int planes_per_block) {
  extern __shared__ __align__(16) unsigned char cv_smem_raw[];
  CvBwdSmem& sm = *reinterpret_cast<CvBwdSmem*>(cv_smem_raw);
  float* rows = reinterpret_cast<float*>(cv_smem_raw + ((sizeof(CvBwdSmem) + 15) / 16) * 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int H = a.H, W = a.W, K = a.K;
  const size_t HW = (size_t)H * W;
  {
    const float* w = a.mlp;
    for (int k = tid; k < kCvHid * kCvIn; k += kCvThreads) { const int o = k / kCvIn, i = k - o * kCvIn; sm.W0t[i][o] = w[k]; sm.W0[o][i] = w[k]; }
    for (int k = tid; k < kCvHid * 3; k += kCvThreads) sm.W0[k / 3][kCvIn + (k % 3)] = 0.f;
    w += kCvHid * kCvIn;
    for (int k = tid; k < kCvHid; k += kCvThreads) sm.b0[k] = w[k];
    w += kCvHid;
    for (int k = tid; k < kCvHid * kCvHid; k += kCvThreads) { const int o = k / kCvHid, i = k - o * kCvHid; sm.W1t[i][o] = w[k]; sm.W1[o][i] = w[k]; }
    w += kCvHid * kCvHid;
    for (int k = tid; k < kCvHid; k += kCvThreads) sm.b1[k] = w[k];
    w += kCvHid;
    for (int k = tid; k < kCvHid; k += kCvThreads) sm.W2[k] = w[k];
    w += kCvHid;
    if (tid == 0) sm.b2 = w[0];
    for (int k = tid; k < K * 12; k += kCvThreads) sm.proj[k] = a.proj[(size_t)b * K * 12 + k];
  }
  __syncthreads();
  int u, v;
  const bool active = patch_pixel(blockIdx.x, tid, H, W, u, v);
  if (!active) { u = 0; v = 0; }
  const int p = v * W + u;
  const int pc = p;
  const float uvx = 1.0f / (float)W, uvy = 1.0f / (float)H;
  const float* ik = a.cur_invK + (size_t)b * 9;
  const float pu = (float)u + 0.5f, pv = (float)v + 0.5f;
  const float r0 = fmaf(ik[1], pv, ik[0] * pu) + ik[2];
  const float r1 = fmaf(ik[4], pv, ik[3] * pu) + ik[5];
  const float r2 = fmaf(ik[7], pv, ik[6] * pu) + ik[8];
  float cur[kCvC], dcur[kCvC];
  {
    const float* cb = a.cur_feats + (size_t)b * kCvC * HW + pc;
#pragma unroll
    for (int c = 0; c < kCvC; c++) { cur[c] = __ldg(cb + (size_t)c * HW); dcur[c] = 0.f; }
  }
  const float4* src_b = reinterpret_cast<const float4*>(a.src_packed) + (size_t)b * K * (kCvC / 4) * HW;
  float4* dsrc_b = reinterpret_cast<float4*>(a.dsrc_packed) + (size_t)b * K * (kCvC / 4) * HW;
  const unsigned all = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
  const int d0 = blockIdx.y * planes_per_block, d1 = min(a.D, d0 + planes_per_block);
  float* myrow = rows + (size_t)tid * kRowStride;

  // phase-B accumulators.  warp 0..2: dW0[o=lane][16*warp .. +16) and dW1[lane][4*warp .. +4);
  // warp 3: dW0[lane][48], dW1[lane][12..32), db0[lane], db1[lane], dW2[lane] (+ db2 in lane 0)
  float accA[16], accB[20], acc_b0 = 0.f, acc_b1 = 0.f, acc_w2 = 0.f, acc_b2 = 0.f;
#pragma unroll
  for (int k = 0; k < 16; k++) accA[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 20; k++) accB[k] = 0.f;

  for (int d = d0; d < d1; d++) {
    // ------------------------------------------------------------------ phase A
    float g = 0.f;
    if (active) g = __ldg(a.dL_dout + ((size_t)b * a.D + d) * HW + p);
    {
      const float zd = __ldg(a.planes + d);
      const float X0 = zd * r0, X1 = zd * r1, X2 = zd * r2;
      float x[kCvC];
      float dsum;
      unsigned geo, zero;
      gather_sources_packed(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, all, x, dsum, geo, zero);
      const unsigned valid = geo & ~zero;
      if (zero) { float ds2; unsigned g2, z2; gather_sources_packed(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, valid, x, ds2, g2, z2); }
      const float rn = 1.0f / ((float)__popc(valid) + 1e-8f);
      float z1[kCvHid], z2v[kCvHid];
#pragma unroll
      for (int o = 0; o < kCvHid; o++) z1[o] = sm.b0[o];
#pragma unroll
      for (int i = 0; i < kCvC; i++) {
        x[i] = x[i] * rn;
#pragma unroll
        for (int o = 0; o < kCvHid; o++) z1[o] = fmaf(sm.W0t[i][o], x[i], z1[o]);
      }
      const float xdot = dsum * rn;
#pragma unroll
      for (int o = 0; o < kCvHid; o++) z1[o] = fmaf(sm.W0t[kCvC][o], xdot, z1[o]);
#pragma unroll
      for (int o = 0; o < kCvHid; o++) z2v[o] = sm.b1[o];
#pragma unroll
      for (int i = 0; i < kCvHid; i++) {
        const float hi = leaky(z1[i]);
#pragma unroll
        for (int o = 0; o < kCvHid; o++) z2v[o] = fmaf(sm.W1t[i][o], hi, z2v[o]);
      }
      // park x, a1, g*a2, g ; then dz2, dz1
#pragma unroll
      for (int i = 0; i < kCvC; i++) myrow[kOffX + i] = x[i];
      myrow[kOffX + kCvC] = xdot; myrow[kOffX + 49] = 0.f; myrow[kOffX + 50] = 0.f; myrow[kOffX + 51] = 0.f;
#pragma unroll
      for (int i = 0; i < kCvHid; i++) { myrow[kOffA1 + i] = leaky(z1[i]); myrow[kOffGa2 + i] = g * leaky(z2v[i]); }
      myrow[kOffG] = g;
      // dz2 = g * W2 * lk'(z2)
#pragma unroll
      for (int i = 0; i < kCvHid; i++) { z2v[i] = g * sm.W2[i] * dleaky(z2v[i]); myrow[kOffDz2 + i] = z2v[i]; }
      // da1[i] = sum_o W1[o][i] dz2[o] ; dz1 = da1 * lk'(z1)
      float da1[kCvHid];
#pragma unroll
      for (int i = 0; i < kCvHid; i++) da1[i] = 0.f;
#pragma unroll
      for (int o = 0; o < kCvHid; o++) {
#pragma unroll
        for (int i = 0; i < kCvHid; i++) da1[i] = fmaf(sm.W1[o][i], z2v[o], da1[i]);
      }
#pragma unroll
      for (int i = 0; i < kCvHid; i++) { z1[i] = da1[i] * dleaky(z1[i]); myrow[kOffDz1 + i] = z1[i]; }
      // dx[i] = sum_o W0[o][i] dz1[o]   (x[] reused as dx[])
      float dxdot = 0.f;
#pragma unroll
      for (int i = 0; i < kCvC; i++) x[i] = 0.f;
#pragma unroll
      for (int o = 0; o < kCvHid; o++) {
#pragma unroll
        for (int i = 0; i < kCvC; i++) x[i] = fmaf(sm.W0[o][i], z1[o], x[i]);
        dxdot = fmaf(sm.W0[o][kCvC], z1[o], dxdot);
      }
      // ---- back through the masked means: dw_k[c] = [valid_k] dx[c]/n + [geo_k] cur[c] dxdot/n
      if (active && g != 0.f) {
        const float gd = dxdot * rn;
        for (int k = 0; k < K; k++) {
          if (!((geo >> k) & 1u)) continue;
          Taps t;
          make_taps(sm.proj + 12 * k, X0, X1, X2, H, W, uvx, uvy, t);
          const float fv = ((valid >> k) & 1u) ? rn : 0.f;
          const float4* __restrict__ s = src_b + (size_t)k * (kCvC / 4) * HW;
          float4* __restrict__ ds = dsrc_b + (size_t)k * (kCvC / 4) * HW;
#pragma unroll
          for (int g4 = 0; g4 < kCvC / 4; g4++) {
            // d dot_k / d cur[c] = w_k[c]  (re-gathered: cheaper than keeping K x 48 registers)
            const float4* __restrict__ sg = s + (size_t)g4 * HW;
            const float4 qa = __ldg(sg + t.o00), qb = __ldg(sg + t.o01), qc = __ldg(sg + t.o10), qd = __ldg(sg + t.o11);
            const float wk[4] = {fmaf(t.w11, qd.x, fmaf(t.w10, qc.x, fmaf(t.w01, qb.x, t.w00 * qa.x))),
                                 fmaf(t.w11, qd.y, fmaf(t.w10, qc.y, fmaf(t.w01, qb.y, t.w00 * qa.y))),
                                 fmaf(t.w11, qd.z, fmaf(t.w10, qc.z, fmaf(t.w01, qb.z, t.w00 * qa.z))),
                                 fmaf(t.w11, qd.w, fmaf(t.w10, qc.w, fmaf(t.w01, qb.w, t.w00 * qa.w)))};
            float gw[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int c = 4 * g4 + e;
              gw[e] = fmaf(x[c], fv, cur[c] * gd);
              dcur[c] = fmaf(gd, wk[e], dcur[c]);
            }
            float4* dg = ds + (size_t)g4 * HW;     // d(warped) scattered with one 16-byte reduction per tap
            if (t.w00 != 0.f) red_add_v4(dg + t.o00, t.w00 * gw[0], t.w00 * gw[1], t.w00 * gw[2], t.w00 * gw[3]);
            if (t.w01 != 0.f) red_add_v4(dg + t.o01, t.w01 * gw[0], t.w01 * gw[1], t.w01 * gw[2], t.w01 * gw[3]);
            if (t.w10 != 0.f) red_add_v4(dg + t.o10, t.w10 * gw[0], t.w10 * gw[1], t.w10 * gw[2], t.w10 * gw[3]);
            if (t.w11 != 0.f) red_add_v4(dg + t.o11, t.w11 * gw[0], t.w11 * gw[1], t.w11 * gw[2], t.w11 * gw[3]);
          }
        }
      }
      if (!active || g == 0.f) {   // rows that must not contribute to the parameter gradients
        if (!active) {
#pragma unroll
          for (int i = 0; i < kCvHid; i++) { myrow[kOffDz1 + i] = 0.f; myrow[kOffDz2 + i] = 0.f; myrow[kOffGa2 + i] = 0.f; }
          myrow[kOffG] = 0.f;
        }
      }
    }
    __syncthreads();
    // ------------------------------------------------------------------ phase B
    if (warp < 3) {
      for (int r = 0; r < kCvThreads; r++) {
        const float* row = rows + (size_t)r * kRowStride;
        const float dz1 = row[kOffDz1 + lane], dz2 = row[kOffDz2 + lane];
        const float4* xr = reinterpret_cast<const float4*>(row + kOffX + 16 * warp);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4 xv = xr[q];
          accA[4 * q + 0] = fmaf(dz1, xv.x, accA[4 * q + 0]); accA[4 * q + 1] = fmaf(dz1, xv.y, accA[4 * q + 1]);
          accA[4 * q + 2] = fmaf(dz1, xv.z, accA[4 * q + 2]); accA[4 * q + 3] = fmaf(dz1, xv.w, accA[4 * q + 3]);
        }
        const float4 av = *reinterpret_cast<const float4*>(row + kOffA1 + 4 * warp);
        accB[0] = fmaf(dz2, av.x, accB[0]); accB[1] = fmaf(dz2, av.y, accB[1]);
        accB[2] = fmaf(dz2, av.z, accB[2]); accB[3] = fmaf(dz2, av.w, accB[3]);
      }
    } else {
      for (int r = 0; r < kCvThreads; r++) {
        const float* row = rows + (size_t)r * kRowStride;
        const float dz1 = row[kOffDz1 + lane], dz2 = row[kOffDz2 + lane];
        accA[0] = fmaf(dz1, row[kOffX + kCvC], accA[0]);
        acc_b0 += dz1; acc_b1 += dz2;
        acc_w2 += row[kOffGa2 + lane];
        acc_b2 += row[kOffG];
        const float4* ar = reinterpret_cast<const float4*>(row + kOffA1 + 12);
#pragma unroll
        for (int q = 0; q < 5; q++) {
          const float4 av = ar[q];
          accB[4 * q + 0] = fmaf(dz2, av.x, accB[4 * q + 0]); accB[4 * q + 1] = fmaf(dz2, av.y, accB[4 * q + 1]);
          accB[4 * q + 2] = fmaf(dz2, av.z, accB[4 * q + 2]); accB[4 * q + 3] = fmaf(dz2, av.w, accB[4 * q + 3]);
        }
      }
    }
    __syncthreads();
  }
  // ---- flush dL_dcur (several plane chunks add to the same pixel) ----
  if (active) {
    float* dc = a.dL_dcur + (size_t)b * kCvC * HW + p;
#pragma unroll
    for (int c = 0; c < kCvC; c++) if (dcur[c] != 0.f) atomicAdd(dc + (size_t)c * HW, dcur[c]);
  }
  // ---- flush parameter gradients: packed layout W0[32,49] b0[32] W1[32,32] b1[32] W2[32] b2[1] ----
  float* gW0 = a.dL_dmlp;
  float* gb0 = gW0 + kCvHid * kCvIn;
  float* gW1 = gb0 + kCvHid;
  float* gb1 = gW1 + kCvHid * kCvHid;
  float* gW2 = gb1 + kCvHid;
  float* gb2 = gW2 + kCvHid;
  if (warp < 3) {
#pragma unroll
    for (int k = 0; k < 16; k++) atomicAdd(gW0 + lane * kCvIn + 16 * warp + k, accA[k]);
#pragma unroll
    for (int k = 0; k < 4; k++) atomicAdd(gW1 + lane * kCvHid + 4 * warp + k, accB[k]);
  } else {
    atomicAdd(gW0 + lane * kCvIn + kCvC, accA[0]);
#pragma unroll
    for (int k = 0; k < 20; k++) atomicAdd(gW1 + lane * kCvHid + 12 + k, accB[k]);
    atomicAdd(gb0 + lane, acc_b0); atomicAdd(gb1 + lane, acc_b1); atomicAdd(gW2 + lane, acc_w2);
    if (lane == 0) atomicAdd(gb2, acc_b2);
  }
}


// =========================================================================== tensor-core backward
// The fp32 kernel above keeps 1 CTA x 4 warps per SM (121 KB of parked rows, 255 registers + spills) and spends
// its time in dependent FMA chains (ncu r1: FMA pipe 10.7 %, 6 % warps active).  Here every contraction of the
// backward pass is a tcgen05 MMA (3xTF32, fp32-accurate):
//
//   Z1  = X  . W0^T          (M128 N32 K56)  X = [x*rn (48) | dot*rn | 1 | 0..]  -> the bias b0 rides in column 49
//   Z2  = A1 . W1^T          (M128 N32 K32)  A1 = leaky(Z1)
//   dA1 = dZ2 . W1           (M128 N32 K32)  dZ2 = g * W2 * leaky'(Z2)
//   dX  = dZ1 . W0           (M128 N64 K32)  dZ1 = dA1 * leaky'(Z1)
//   U1 += [X | A1]^T . dZ2   (M128 N32, K = the 128 rows)   rows 64..95 = dW1^T, row 49 (the ones column) = db1
//   U0 += [X | A1]^T . dZ1   (M128 N32, K = the 128 rows)   rows 0..48 = dW0^T, row 49 = db0
//
// Row-wise products take their A operand (row = TMEM lane = reference pixel) straight from tensor memory, where the
// threads put it with tcgen05.st; weights are K-major SWIZZLE_NONE tiles in shared memory (both orientations).
// The two products that reduce over the ROWS need MN-major operands: for tf32 the tensor core accepts those only in
// the SWIZZLE_128B_BASE32B layout (measured with tools/probe/umma_layout_probe.cu: every other layout type reads as
// zeros): element (row, col) at  (col/32)*LBO + (row/4)*SBO + (row%4)*128 + ((((col%32)/8) ^ (row%4))*32 + (col%8)*4,
// here dense (SBO = 512, LBO = 16 KB): a thread writes its 128-byte row with the 32-byte units permuted by row%4.
// U0/U1 stay in TMEM across all planes of the CTA and are flushed once.
// 256 threads: thread t and t+128 share reference pixel t&127 and split the 48 channels (24 each) and every
// 32-column TMEM read (16 each) -> 8 warps per SM and no spills, instead of 4 warps that spill.
namespace tcb {

using tc::commit;
using tc::elect_one;
using tc::mbar_wait;
using tc::smem_u32;
using tc::to_tf32;

constexpr int kThreads = 256;
constexpr int kHalfC = kCvC / 2;                 // channels per thread
constexpr uint32_t kGroup = 128 * 128;           // one MN-major group: 128 rows x 32 columns x 4 B
// tensor-memory columns: accumulators ...
constexpr uint32_t kColZ1 = 0, kColZ2 = 32, kColDA1 = 64, kColDX = 96, kColU0 = 160, kColU1 = 192;
// ... and A operands (hi / lo halves of the 3xTF32 split)
constexpr uint32_t kColXhi = 224, kColXlo = 280, kColHhi = 336, kColHlo = 368, kTmemCols = 512;
constexpr int kMaxK = 16;

struct __align__(1024) Smem {
  // MN-major tiles.  The M = 128 operand window is 4 groups from Xg0: [X cols 0..31][X cols 32..63][A1][whatever follows];
  // the 4th group only feeds accumulator rows 96..127, which nobody reads.
  unsigned char X_hi[2 * kGroup], A1_hi[kGroup];
  unsigned char X_lo[2 * kGroup], A1_lo[kGroup];
  unsigned char dZ_hi[kGroup], dZ_lo[kGroup];        // dZ2 (U1's B operand)
  unsigned char dZb_hi[kGroup], dZb_lo[kGroup];      // dZ1 (U0's B operand): its own tile, so that U1 / U0 can run BEHIND the chain
  // K-major SWIZZLE_NONE weight tiles, [k-step][n/8][2 chunks][8 rows][16 B]
  unsigned char W0f_hi[7 * 1024], W0f_lo[7 * 1024];     // n = out (32), k = in (56):  Z1 = X W0^T
  unsigned char W1f_hi[4 * 1024], W1f_lo[4 * 1024];     // n = out, k = in:            Z2 = A1 W1^T
  unsigned char W1t_hi[4 * 1024], W1t_lo[4 * 1024];     // n = in, k = out:            dA1 = dZ2 W1
  unsigned char W0t_hi[4 * 2048], W0t_lo[4 * 2048];     // n = in (64), k = out (32):  dX = dZ1 W0
  float dotp[2][kMaxK][128];                     // per-half partial dot products of one plane
  float b1[kCvHid], W2[kCvHid];
  float proj[kMaxK * 12];
  unsigned long long bar_chain, bar_dw;
  uint32_t tmem_base;
};

// K-major SWIZZLE_NONE descriptor of a weight tile k-step (LBO = 128: the two 16-byte k-chunks, SBO = 256: 8-row groups)
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
// MN-major SWIZZLE_128B_BASE32B descriptor (layout type 1): LBO = 32-column groups, SBO = 4-row groups
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(kGroup >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// D = f32, A = B = tf32, M = 128
__host__ __device__ constexpr uint32_t idesc(uint32_t n, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t id, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(id), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t id, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(id), "r"(accumulate) : "memory");
}
// Descriptors as (low word, constant high word): the tiles are 1024-byte aligned inside the 227 KB window, so the 14-bit start
// address field never overflows and the k-step of a tile is a plain add on the low word.  (Building every descriptor from the
// byte address -- shift, mask, or -- cost ~17 uniform-datapath instructions per MMA: with 96 weight-gradient MMAs per plane the
// issuing thread arrived ~1.5 us late at the next plane's barrier, ncu r2: 9 % barrier stalls.)
__device__ __forceinline__ uint32_t dlo_k(uint32_t saddr) { return (saddr >> 4) | ((128u >> 4) << 16); }
__device__ __forceinline__ uint32_t dlo_mn(uint32_t saddr) { return (saddr >> 4) | ((kGroup >> 4) << 16); }
constexpr uint32_t kDhiK = (256u >> 4) | (1u << 14);                     // SBO | version 1 (bit 46)
constexpr uint32_t kDhiMn = (512u >> 4) | (1u << 14) | (1u << 29);       // SBO | version 1 | SWIZZLE_128B_BASE32B (bit 61)
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// 3xTF32 product of one k-step: lo*hi + hi*lo + hi*hi  (b_hi / b_lo: descriptor low words of the k-step)
__device__ __forceinline__ void mma3_ts(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t id, uint32_t accumulate) {
  mma_ts(tmem_d, a_lo, mk_desc(b_hi, kDhiK), id, accumulate);
  mma_ts(tmem_d, a_hi, mk_desc(b_lo, kDhiK), id, 1u);
  mma_ts(tmem_d, a_hi, mk_desc(b_hi, kDhiK), id, 1u);
}
__device__ __forceinline__ void mma3_mn(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t id, uint32_t accumulate) {
  mma_ss(tmem_d, mk_desc(a_lo, kDhiMn), mk_desc(b_hi, kDhiMn), id, accumulate);
  mma_ss(tmem_d, mk_desc(a_hi, kDhiMn), mk_desc(b_lo, kDhiMn), id, 1u);
  mma_ss(tmem_d, mk_desc(a_hi, kDhiMn), mk_desc(b_hi, kDhiMn), id, 1u);
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld1_issue(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]) {     // N columns (multiple of 8) of this thread's lane
  uint32_t r[N / 8][8];
#pragma unroll
  for (int q = 0; q < N / 8; q++) tmem_ld8_issue(taddr + (uint32_t)(8 * q), r[q]);
  tmem_ld_wait();
#pragma unroll
  for (int q = 0; q < N / 8; q++)
#pragma unroll
    for (int e = 0; e < 8; e++) v[8 * q + e] = __uint_as_float(r[q][e]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&w)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
// Splits 8 consecutive columns (col0 multiple of 8) of this thread's row into tf32 hi / lo and puts them (a) into the
// TMEM A-operand columns t_hi / t_lo (already offset to col0) and (b) into the MN-major tile pair at mn_hi / mn_lo.
__device__ __forceinline__ void put8(const float (&v)[8], uint32_t t_hi, uint32_t t_lo, bool to_tmem, unsigned char* mn_hi,
                                     unsigned char* mn_lo, int row, int col0) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int e = 0; e < 8; e++) tc::split_tf32(v[e], h[e], l[e]);
  if (to_tmem) { tmem_st8(t_hi, h); tmem_st8(t_lo, l); }
  const uint32_t off = (uint32_t)((col0 >> 5) * kGroup + row * 128 + ((((col0 & 31) >> 3) ^ (row & 3)) << 5));
  *reinterpret_cast<uint4*>(mn_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(mn_hi + off + 16) = make_uint4(h[4], h[5], h[6], h[7]);
  *reinterpret_cast<uint4*>(mn_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  *reinterpret_cast<uint4*>(mn_lo + off + 16) = make_uint4(l[4], l[5], l[6], l[7]);
}
__device__ __forceinline__ void sync_for_mma() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
}
// byte offset of element (n, k) in a K-major SWIZZLE_NONE weight tile with N rows
__device__ __forceinline__ uint32_t wk_off(int n, int k, int N) {
  return (uint32_t)((k >> 3) * (N / 8 * 256) + (n >> 3) * 256 + ((k >> 2) & 1) * 128 + (n & 7) * 16 + (k & 3) * 4);
}
__device__ __forceinline__ void put_w(unsigned char* hi, unsigned char* lo, uint32_t off, float v) {
  const uint32_t h = to_tf32(v);
  *reinterpret_cast<uint32_t*>(hi + off) = h;
  *reinterpret_cast<uint32_t*>(lo + off) = to_tf32(v - __uint_as_float(h));
}

// This thread's half of the gather: channel groups [g0, g0 + 6) of the sources in use_mask.  Partial dot products go
// to dots[k * 128] (0 for sources that are not geometrically valid).
__device__ __forceinline__ void gather_half(const float4* __restrict__ srcp_b, const float* __restrict__ proj, int K, int H, int W,
                                            size_t HW, float X0, float X1, float X2, float uvx, float uvy, int g0,
                                            const float (&cur)[kHalfC], unsigned use_mask, float (&fsum)[kHalfC],
                                            float* __restrict__ dots, unsigned& geo_mask) {
#pragma unroll
  for (int c = 0; c < kHalfC; c++) fsum[c] = 0.f;
  geo_mask = 0u;
  for (int k = 0; k < K; k++) {
    float dot = 0.f;
    Taps t;
    if (((use_mask >> k) & 1u) && make_taps(proj + 12 * k, X0, X1, X2, H, W, uvx, uvy, t)) {
      geo_mask |= 1u << k;
      const float4* __restrict__ s = srcp_b + ((size_t)k * (kCvC / 4) + g0) * HW;
#pragma unroll
      for (int g = 0; g < kHalfC / 4; g++) {
        const float4* __restrict__ sg = s + (size_t)g * HW;
        const float4 a = __ldg(sg + t.o00), b = __ldg(sg + t.o01), c = __ldg(sg + t.o10), d = __ldg(sg + t.o11);
        const float w0 = fmaf(t.w11, d.x, fmaf(t.w10, c.x, fmaf(t.w01, b.x, t.w00 * a.x)));
        const float w1 = fmaf(t.w11, d.y, fmaf(t.w10, c.y, fmaf(t.w01, b.y, t.w00 * a.y)));
        const float w2 = fmaf(t.w11, d.z, fmaf(t.w10, c.z, fmaf(t.w01, b.z, t.w00 * a.z)));
        const float w3 = fmaf(t.w11, d.w, fmaf(t.w10, c.w, fmaf(t.w01, b.w, t.w00 * a.w)));
        dot = fmaf(w0, cur[4 * g], dot); dot = fmaf(w1, cur[4 * g + 1], dot);
        dot = fmaf(w2, cur[4 * g + 2], dot); dot = fmaf(w3, cur[4 * g + 3], dot);
        fsum[4 * g] += w0; fsum[4 * g + 1] += w1; fsum[4 * g + 2] += w2; fsum[4 * g + 3] += w3;
      }
    }
    if (dots) dots[k * 128] = dot;
  }
}

__global__ void __launch_bounds__(kThreads, 1) cost_volume_bwd_tc_kernel(FsCostVolumeArgs a, int planes_per_block) {
  extern __shared__ unsigned char tcb_smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(tcb_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, half = tid >> 7;                 // reference pixel (= tile row = TMEM lane) and channel half
  const int b = blockIdx.z;
  const int H = a.H, W = a.W, K = a.K;
  const size_t HW = (size_t)H * W;
  // ---- one-time setup ----
  {
    const float* w0 = a.mlp;                                // [32][49]
    const float* pb0 = w0 + kCvHid * kCvIn;
    const float* w1 = pb0 + kCvHid;                         // [32][32]
    const float* pb1 = w1 + kCvHid * kCvHid;
    const float* w2 = pb1 + kCvHid;
    // all 12 loads of a thread are issued before the first split / store (the rolled loops waited for one L2 round trip per
    // iteration: ~10 % of the kernel's stall samples sat in this prologue, ncu r2)
    float v0[kCvHid * 64 / kThreads], v1[kCvHid * kCvHid / kThreads];
#pragma unroll
    for (int q = 0; q < kCvHid * 64 / kThreads; q++) {      // o = e>>6, i = e&63: columns 0..48 weights, column 49 = b0, rest 0
      const int e = tid + q * kThreads, o = e >> 6, i = e & 63;
      v0[q] = i < kCvIn ? __ldg(w0 + o * kCvIn + i) : (i == kCvIn ? __ldg(pb0 + o) : 0.f);
    }
#pragma unroll
    for (int q = 0; q < kCvHid * kCvHid / kThreads; q++) v1[q] = __ldg(w1 + tid + q * kThreads);
#pragma unroll
    for (int q = 0; q < kCvHid * 64 / kThreads; q++) {
      const int e = tid + q * kThreads, o = e >> 6, i = e & 63;
      if (i < 56) put_w(sm.W0f_hi, sm.W0f_lo, wk_off(o, i, 32), v0[q]);
      put_w(sm.W0t_hi, sm.W0t_lo, wk_off(i, o, 64), i < kCvIn ? v0[q] : 0.f);     // dX needs no bias column
    }
#pragma unroll
    for (int q = 0; q < kCvHid * kCvHid / kThreads; q++) {
      const int e = tid + q * kThreads, o = e >> 5, i = e & 31;
      put_w(sm.W1f_hi, sm.W1f_lo, wk_off(o, i, 32), v1[q]);
      put_w(sm.W1t_hi, sm.W1t_lo, wk_off(i, o, 32), v1[q]);
    }
    for (int k = tid; k < kCvHid; k += kThreads) { sm.b1[k] = pb1[k]; sm.W2[k] = w2[k]; }
    for (int k = tid; k < K * 12; k += kThreads) sm.proj[k] = a.proj[(size_t)b * K * 12 + k];
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_chain)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(smem_u32(&sm.bar_dw)) : "memory");   // two issuing threads
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const uint32_t tmem = sm.tmem_base;
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);     // this warp's 32 TMEM lanes
  if (half == 1) {                                          // constant columns 56..63 of the MN-major X tile (columns 52..55: see S1)
    const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    put8(z8, 0u, 0u, false, sm.X_hi, sm.X_lo, r, 56);
  }
  const uint32_t bar_chain = smem_u32(&sm.bar_chain), bar_dw = smem_u32(&sm.bar_dw);
  // descriptor low words of the first k-step of every tile (+64 per 1024-byte step, +128 per 2048-byte step)
  const uint32_t aX_hi = dlo_mn(smem_u32(sm.X_hi)), aX_lo = dlo_mn(smem_u32(sm.X_lo)), aDZ_hi = dlo_mn(smem_u32(sm.dZ_hi)), aDZ_lo = dlo_mn(smem_u32(sm.dZ_lo));
  const uint32_t aDZb_hi = dlo_mn(smem_u32(sm.dZb_hi)), aDZb_lo = dlo_mn(smem_u32(sm.dZb_lo));
  const uint32_t aW0f_hi = dlo_k(smem_u32(sm.W0f_hi)), aW0f_lo = dlo_k(smem_u32(sm.W0f_lo)), aW1f_hi = dlo_k(smem_u32(sm.W1f_hi)), aW1f_lo = dlo_k(smem_u32(sm.W1f_lo));
  const uint32_t aW1t_hi = dlo_k(smem_u32(sm.W1t_hi)), aW1t_lo = dlo_k(smem_u32(sm.W1t_lo)), aW0t_hi = dlo_k(smem_u32(sm.W0t_hi)), aW0t_lo = dlo_k(smem_u32(sm.W0t_lo));
  uint32_t ph_chain = 0, ph_dw = 0;
  constexpr uint32_t kIdK32 = idesc(32, 0, 0), kIdK64 = idesc(64, 0, 0), kIdMN32 = idesc(32, 1, 1);

  int u, v;
  const bool active = patch_pixel(blockIdx.x, r, H, W, u, v);
  if (!active) { u = 0; v = 0; }
  const int p = v * W + u;
  const float uvx = 1.0f / (float)W, uvy = 1.0f / (float)H;
  const float* ik = a.cur_invK + (size_t)b * 9;
  const float pu = (float)u + 0.5f, pv = (float)v + 0.5f;
  const float r0 = fmaf(ik[1], pv, ik[0] * pu) + ik[2];
  const float r1 = fmaf(ik[4], pv, ik[3] * pu) + ik[5];
  const float r2 = fmaf(ik[7], pv, ik[6] * pu) + ik[8];
  const int c0 = half * kHalfC, g0 = half * (kHalfC / 4);   // first channel / channel group of this thread
  float cur[kHalfC], dcur[kHalfC];
  {
    const float* cb = a.cur_feats + ((size_t)b * kCvC + c0) * HW + p;
#pragma unroll
    for (int c = 0; c < kHalfC; c++) { cur[c] = __ldg(cb + (size_t)c * HW); dcur[c] = 0.f; }
  }
  float acc_w2[16], acc_b2 = 0.f;                           // dW2[16*half + j], db2 (half 0) summed over this thread's planes
#pragma unroll
  for (int j = 0; j < 16; j++) acc_w2[j] = 0.f;
  const float4* src_b = reinterpret_cast<const float4*>(a.src_packed) + (size_t)b * K * (kCvC / 4) * HW;
  float4* dsrc_b = reinterpret_cast<float4*>(a.dsrc_packed) + (size_t)b * K * (kCvC / 4) * HW;
  const unsigned all = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
  const int d0 = blockIdx.y * planes_per_block, d1 = min(a.D, d0 + planes_per_block);
  float* my_dots = &sm.dotp[half][0][r];

  for (int d = d0; d < d1; d++) {
    const bool first = d == d0;
    float g = 0.f;
    if (active) g = __ldg(a.dL_dout + ((size_t)b * a.D + d) * HW + p);
    const float zd = __ldg(a.planes + d);
    const float X0 = zd * r0, X1 = zd * r1, X2 = zd * r2;
    // ------------------------------------------------------------ S1: gather, X operand, Z1 = X W0^T (+ b0)
    unsigned geo, valid;
    float rn;
    float xs[kHalfC];                                       // x * rn of this thread's channels (live until S5)
    {
      gather_half(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, g0, cur, all, xs, my_dots, geo);
      __syncthreads();
      float dsum = 0.f;
      unsigned zero = 0u;
      for (int k = 0; k < K; k++) {
        const float dot = sm.dotp[0][k][r] + sm.dotp[1][k][r];
        if (((geo >> k) & 1u) && dot == 0.f) zero |= 1u << k;
        dsum += dot;
      }
      valid = geo & ~zero;
      if (zero) { unsigned g2; gather_half(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, g0, cur, valid, xs, nullptr, g2); }
      rn = 1.0f / ((float)__popc(valid) + 1e-8f);
#pragma unroll
      for (int c = 0; c < kHalfC; c++) xs[c] *= rn;
      if (!first) { mbar_wait(bar_dw, ph_dw); ph_dw ^= 1u; }          // U1 / U0 of the previous plane have finished reading the X, A1 and dZ tiles
#pragma unroll
      for (int q = 0; q < kHalfC / 8; q++) {
        const float v8[8] = {xs[8 * q], xs[8 * q + 1], xs[8 * q + 2], xs[8 * q + 3], xs[8 * q + 4], xs[8 * q + 5], xs[8 * q + 6], xs[8 * q + 7]};
        put8(v8, t_row + kColXhi + (uint32_t)(c0 + 8 * q), t_row + kColXlo + (uint32_t)(c0 + 8 * q), true, sm.X_hi, sm.X_lo, r, c0 + 8 * q);
      }
      if (half == 0) {
        const float v8[8] = {dsum * rn, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        put8(v8, t_row + kColXhi + 48u, t_row + kColXlo + 48u, true, sm.X_hi, sm.X_lo, r, 48);
      }
    }
    sync_for_mma();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < 7; s++)
        mma3_ts(tmem + kColZ1, tmem + kColXhi + 8u * s, tmem + kColXlo + 8u * s, aW0f_hi + s * 64, aW0f_lo + s * 64, kIdK32, s > 0 ? 1u : 0u);
      commit(bar_chain);
    }
    mbar_wait(bar_chain, ph_chain); ph_chain ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ------------------------------------------------------------ S2: A1 = leaky(Z1), Z2 = A1 W1^T
    {
      float z[16];
      tmem_ld<16>(t_row + kColZ1 + 16u * half, z);
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const float v8[8] = {leaky(z[8 * q]), leaky(z[8 * q + 1]), leaky(z[8 * q + 2]), leaky(z[8 * q + 3]),
                             leaky(z[8 * q + 4]), leaky(z[8 * q + 5]), leaky(z[8 * q + 6]), leaky(z[8 * q + 7])};
        put8(v8, t_row + kColHhi + (uint32_t)(16 * half + 8 * q), t_row + kColHlo + (uint32_t)(16 * half + 8 * q), true, sm.A1_hi, sm.A1_lo, r, 16 * half + 8 * q);
      }
    }
    sync_for_mma();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < 4; s++)
        mma3_ts(tmem + kColZ2, tmem + kColHhi + 8u * s, tmem + kColHlo + 8u * s, aW1f_hi + s * 64, aW1f_lo + s * 64, kIdK32, s > 0 ? 1u : 0u);
      commit(bar_chain);
    }
    mbar_wait(bar_chain, ph_chain); ph_chain ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ------------------------------------------------------------ S3: dZ2 = g W2 leaky'(Z2); dA1 = dZ2 W1; U1 += [X|A1]^T dZ2
    {
      float z[16];
      tmem_ld<16>(t_row + kColZ2 + 16u * half, z);
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const float z2 = z[j] + sm.b1[16 * half + j];
        acc_w2[j] = fmaf(g, leaky(z2), acc_w2[j]);
        z[j] = g * sm.W2[16 * half + j] * dleaky(z2);
      }
      if (half == 0) acc_b2 += g;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const float v8[8] = {z[8 * q], z[8 * q + 1], z[8 * q + 2], z[8 * q + 3], z[8 * q + 4], z[8 * q + 5], z[8 * q + 6], z[8 * q + 7]};
        put8(v8, t_row + kColHhi + (uint32_t)(16 * half + 8 * q), t_row + kColHlo + (uint32_t)(16 * half + 8 * q), true, sm.dZ_hi, sm.dZ_lo, r, 16 * half + 8 * q);
      }
    }
    sync_for_mma();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < 4; s++)
        mma3_ts(tmem + kColDA1, tmem + kColHhi + 8u * s, tmem + kColHlo + 8u * s, aW1t_hi + s * 64, aW1t_lo + s * 64, kIdK32, s > 0 ? 1u : 0u);
      commit(bar_chain);               // (U1 is issued behind the dX product below: the dependent chain never queues behind it)
    }
    mbar_wait(bar_chain, ph_chain); ph_chain ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ------------------------------------------------------------ S4: dZ1 = dA1 leaky'(Z1); dX = dZ1 W0; U0 += [X|A1]^T dZ1
    {
      float da[16], z[16];
      tmem_ld<16>(t_row + kColDA1 + 16u * half, da);
      tmem_ld<16>(t_row + kColZ1 + 16u * half, z);
#pragma unroll
      for (int j = 0; j < 16; j++) da[j] *= dleaky(z[j]);
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const float v8[8] = {da[8 * q], da[8 * q + 1], da[8 * q + 2], da[8 * q + 3], da[8 * q + 4], da[8 * q + 5], da[8 * q + 6], da[8 * q + 7]};
        put8(v8, t_row + kColHhi + (uint32_t)(16 * half + 8 * q), t_row + kColHlo + (uint32_t)(16 * half + 8 * q), true, sm.dZb_hi, sm.dZb_lo, r, 16 * half + 8 * q);
      }
    }
    sync_for_mma();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < 4; s++)
        mma3_ts(tmem + kColDX, tmem + kColHhi + 8u * s, tmem + kColHlo + 8u * s, aW0t_hi + s * 128, aW0t_lo + s * 128, kIdK64, s > 0 ? 1u : 0u);
      commit(bar_chain);
      // the two weight-gradient products (2 x 48 MMAs) run behind the chain, underneath the scatter of this plane and the gather
      // of the next one; bar_dw (two arrivals) guards the X / A1 / dZ tiles they read (waited in S1 of the next plane).  They are
      // issued by two threads of different warps (~10 instructions per MMA: one thread issuing all 96 reached the next plane's
      // barrier ~1.5 us after the other warps).
#pragma unroll
      for (int s = 0; s < 16; s++)     // k = 8 tile rows per step
        mma3_mn(tmem + kColU1, aX_hi + s * 64, aX_lo + s * 64, aDZ_hi + s * 64, aDZ_lo + s * 64, kIdMN32, (first && s == 0) ? 0u : 1u);
      commit(bar_dw);
    } else if (warp == 4 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < 16; s++)
        mma3_mn(tmem + kColU0, aX_hi + s * 64, aX_lo + s * 64, aDZb_hi + s * 64, aDZb_lo + s * 64, kIdMN32, (first && s == 0) ? 0u : 1u);
      commit(bar_dw);
    }
    mbar_wait(bar_chain, ph_chain); ph_chain ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ------------------------------------------------------------ S5: dX -> dL_dcur (registers), dL_dsrc (vector reductions)
    float dx[kHalfC];
    uint32_t dd;
    {                                // tcgen05.ld is warp-collective: issued by every lane, outside the divergent part
      uint32_t q0[8], q1[8], q2[8];
      tmem_ld8_issue(t_row + kColDX + (uint32_t)c0, q0);
      tmem_ld8_issue(t_row + kColDX + (uint32_t)c0 + 8u, q1);
      tmem_ld8_issue(t_row + kColDX + (uint32_t)c0 + 16u, q2);
      tmem_ld1_issue(t_row + kColDX + (uint32_t)kCvC, dd);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 8; e++) { dx[e] = __uint_as_float(q0[e]); dx[8 + e] = __uint_as_float(q1[e]); dx[16 + e] = __uint_as_float(q2[e]); }
    }
    const bool row_on = active && g != 0.f;
    const float dxdot = __uint_as_float(dd);
    const float gd = dxdot * rn;
    if (row_on) {
      // d dot_k / d cur[c] = w_k[c] and gd is the same for every source: gd * sum_k w_k[c] = dxdot * (x[c] * rn)
      if (valid == geo) {
#pragma unroll
        for (int c = 0; c < kHalfC; c++) dcur[c] = fmaf(dxdot, xs[c], dcur[c]);
      } else {                       // a source whose dot product is exactly 0 still feeds dL_dcur: gather the geometric set again
        float wsum[kHalfC]; unsigned g2;
        gather_half(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, g0, cur, geo, wsum, nullptr, g2);
#pragma unroll
        for (int c = 0; c < kHalfC; c++) dcur[c] = fmaf(gd, wsum[c], dcur[c]);
      }
    }
    // back through the masked means: d(warped_k)[c] = [valid_k] dx[c] / n + [geo_k] cur[c] dxdot / n, scattered with 16-byte vector
    // reductions.  The lanes of a warp are 32 neighbouring pixels of one row: where the sampling step is one texel, lane L's RIGHT
    // taps are lane L+1's LEFT taps.  Those pairs are merged in registers (shuffle + add) and leave as ONE reduction: the scatter
    // was 25 % of this kernel's stall samples with 11 GB of reduction traffic into L2 per backward (ncu r2), this halves it.
    // Every lane runs the loop (the shuffles are warp-wide); lanes without a contribution carry zero weights and offset -1.
    for (int k = 0; k < K; k++) {
      const bool on_k = row_on && ((geo >> k) & 1u);
      if (!__any_sync(0xffffffffu, on_k)) continue;                        // warp-uniform
      Taps t;
      if (!(on_k && make_taps(sm.proj + 12 * k, X0, X1, X2, H, W, uvx, uvy, t))) {
        t.o00 = t.o01 = t.o10 = t.o11 = -1;
        t.w00 = t.w01 = t.w10 = t.w11 = 0.f;
      }
      const int nb00 = __shfl_down_sync(0xffffffffu, t.o00, 1), nb10 = __shfl_down_sync(0xffffffffu, t.o10, 1);
      const bool send_top = lane < 31 && t.w01 != 0.f && nb00 == t.o01;    // my right taps go to lane + 1 ...
      const bool send_bot = lane < 31 && t.w11 != 0.f && nb10 == t.o11;
      const bool recv_top = __shfl_up_sync(0xffffffffu, (int)send_top, 1) != 0 && lane > 0;   // ... and lane - 1's come to me
      const bool recv_bot = __shfl_up_sync(0xffffffffu, (int)send_bot, 1) != 0 && lane > 0;
      const float fv = (on_k && ((valid >> k) & 1u)) ? rn : 0.f;
      const float gk = on_k ? gd : 0.f;
      float4* __restrict__ ds = dsrc_b + ((size_t)k * (kCvC / 4) + g0) * HW;
#pragma unroll
      for (int q = 0; q < kHalfC / 4; q++) {
        float c00[4], c01[4], c10[4], c11[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float gw = fmaf(dx[4 * q + e], fv, cur[4 * q + e] * gk);
          c00[e] = t.w00 * gw; c01[e] = t.w01 * gw; c10[e] = t.w10 * gw; c11[e] = t.w11 * gw;
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float rt = __shfl_up_sync(0xffffffffu, c01[e], 1), rb = __shfl_up_sync(0xffffffffu, c11[e], 1);
          if (recv_top) c00[e] += rt;
          if (recv_bot) c10[e] += rb;
        }
        float4* dg = ds + (size_t)q * HW;
        red_add_v4_if(t.w00 != 0.f || recv_top, dg + t.o00, c00[0], c00[1], c00[2], c00[3]);
        red_add_v4_if(t.w01 != 0.f && !send_top, dg + t.o01, c01[0], c01[1], c01[2], c01[3]);
        red_add_v4_if(t.w10 != 0.f || recv_bot, dg + t.o10, c10[0], c10[1], c10[2], c10[3]);
        red_add_v4_if(t.w11 != 0.f && !send_bot, dg + t.o11, c11[0], c11[1], c11[2], c11[3]);
      }
    }
  }
  // ---- flush dL_dcur (the plane chunks of one pixel live in different CTAs) ----
  if (active) {
    float* dc = a.dL_dcur + ((size_t)b * kCvC + c0) * HW + p;
#pragma unroll
    for (int c = 0; c < kHalfC; c++) if (dcur[c] != 0.f) atomicAdd(dc + (size_t)c * HW, dcur[c]);
  }
  // ---- parameter gradients: packed layout W0[32,49] b0[32] W1[32,32] b1[32] W2[32] b2[1] ----
  float* gW0 = a.dL_dmlp;
  float* gb0 = gW0 + kCvHid * kCvIn;
  float* gW1 = gb0 + kCvHid;
  float* gb1 = gW1 + kCvHid * kCvHid;
  float* gW2 = gb1 + kCvHid;
  float* gb2 = gW2 + kCvHid;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    float s = acc_w2[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) atomicAdd(gW2 + 16 * half + j, s);
  }
  if (half == 0) {
    float s = acc_b2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) atomicAdd(gb2, s);
  }
  if (d1 > d0) {
    mbar_wait(bar_dw, ph_dw); ph_dw ^= 1u;                  // the last U0 chain
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float w[kCvHid];
    if (half == 0) {                                        // lane r of U0 = column r of [X | A1]: input i = r, r == 49: the ones column
      tmem_ld<32>(t_row + kColU0, w);
      if (r < kCvIn) {
#pragma unroll
        for (int o = 0; o < kCvHid; o++) atomicAdd(gW0 + o * kCvIn + r, w[o]);
      } else if (r == kCvIn) {
#pragma unroll
        for (int o = 0; o < kCvHid; o++) atomicAdd(gb0 + o, w[o]);
      }
    } else {
      tmem_ld<32>(t_row + kColU1, w);
      if (r >= 64 && r < 64 + kCvHid) {
#pragma unroll
        for (int o = 0; o < kCvHid; o++) atomicAdd(gW1 + o * kCvHid + (r - 64), w[o]);
      } else if (r == kCvIn) {
#pragma unroll
        for (int o = 0; o < kCvHid; o++) atomicAdd(gb1 + o, w[o]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace tcb

int launch_cost_volume_bwd(const FsCostVolumeArgs& a, cudaStream_t s) {
  const size_t HW = (size_t)a.H * a.W;
  const int ppb = 16;
  int rc;
  const size_t n_src = (size_t)a.B * a.K * kCvC * HW, n_cur = (size_t)a.B * kCvC * HW;
  const size_t n_mlp = kCvHid * kCvIn + kCvHid + kCvHid * kCvHid + kCvHid + kCvHid + 1;
  if ((rc = launch_pack(a, s))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(a.dsrc_packed, 0, n_src * sizeof(float), s), "memset dsrc_packed"))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(a.dL_dcur, 0, n_cur * sizeof(float), s), "memset dL_dcur"))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(a.dL_dmlp, 0, n_mlp * sizeof(float), s), "memset dL_dmlp"))) return rc;
  dim3 grid(patch_blocks(a.H, a.W), (unsigned)((a.D + ppb - 1) / ppb), (unsigned)a.B);
  if (a.mlp_mode == 1) {   // fp32 CUDA-core kernel: validation path for the tensor-core kernel
    const size_t smem = ((sizeof(CvBwdSmem) + 15) / 16) * 16 + (size_t)kCvThreads * kRowStride * sizeof(float);
    if ((rc = check_cuda(cudaFuncSetAttribute(cost_volume_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "cudaFuncSetAttribute(cost_volume_bwd_kernel)"))) return rc;
    cost_volume_bwd_kernel<<<grid, kCvThreads, smem, s>>>(a, ppb);
    if ((rc = check_cuda(cudaGetLastError(), "cost_volume_bwd_kernel"))) return rc;
  } else {
    const size_t smem = sizeof(tcb::Smem) + 1024;
    if ((rc = check_cuda(cudaFuncSetAttribute(tcb::cost_volume_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "cudaFuncSetAttribute(cost_volume_bwd_tc_kernel)"))) return rc;
    tcb::cost_volume_bwd_tc_kernel<<<grid, tcb::kThreads, smem, s>>>(a, ppb);
    if ((rc = check_cuda(cudaGetLastError(), "cost_volume_bwd_tc_kernel"))) return rc;
  }
  const size_t total = n_src / 4;
  cv_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(a.dsrc_packed), a.dL_dsrc, HW, total);
  return check_cuda(cudaGetLastError(), "cv_unpack_kernel");
}

}  // namespace fs
