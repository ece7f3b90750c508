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

// CTA -> 32x4 pixel patch (warp = one 32-pixel row segment, the 4 warps = 4 consecutive rows): rows y and y+1 share
// their bilinear tap rows, which lifts the L1 hit rate of the gather over a 128-pixel run of a single row.
__device__ __forceinline__ bool patch_pixel(int block, int tid, int H, int W, int& u, int& v) {
  const int px = (W + 31) >> 5;
  const int by = block / px, bx = block - by * px;
  u = bx * 32 + (tid & 31); v = by * 4 + (tid >> 5);
  return u < W && v < H;
}
__host__ __device__ inline unsigned patch_blocks(int H, int W) { return (unsigned)(((W + 31) >> 5) * ((H + 3) >> 2)); }

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.01f * x; }

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
  h.x = to_tf32(v0); h.y = to_tf32(v1); h.z = to_tf32(v2); h.w = to_tf32(v3);
  l.x = to_tf32(v0 - __uint_as_float(h.x)); l.y = to_tf32(v1 - __uint_as_float(h.y));
  l.z = to_tf32(v2 - __uint_as_float(h.z)); l.w = to_tf32(v3 - __uint_as_float(h.w));
  *reinterpret_cast<uint4*>(hi + off) = h;
  *reinterpret_cast<uint4*>(lo + off) = l;
}

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
    for (int e = tid; e < kCvHid * kK1; e += kCvThreads) {
      const int n = e / kK1, k = e - n * kK1;
      const float v = k < kCvIn ? w0[n * kCvIn + k] : 0.f;
      const uint32_t hi = to_tf32(v), lo = to_tf32(v - __uint_as_float(hi));
      const uint32_t off = op_off(n, k, kBBytesPerStep);
      *reinterpret_cast<uint32_t*>(sm.B0_hi + off) = hi; *reinterpret_cast<uint32_t*>(sm.B0_lo + off) = lo;
    }
    const float* pb0 = w0 + kCvHid * kCvIn;
    const float* w1 = pb0 + kCvHid;                         // [32][32]
    for (int e = tid; e < kCvHid * kCvHid; e += kCvThreads) {
      const int n = e / kCvHid, k = e - n * kCvHid;
      const float v = w1[e];
      const uint32_t hi = to_tf32(v), lo = to_tf32(v - __uint_as_float(hi));
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
  const int ppb = 8;
  if (int rc = launch_pack(a, s)) return rc;
  (void)HW;
  dim3 grid(patch_blocks(a.H, a.W), (unsigned)((a.D + ppb - 1) / ppb), (unsigned)a.B);
  if (a.mlp_mode == 1) {   // fp32 CUDA-core MLP: validation path for the tensor-core kernel
    cost_volume_fwd_kernel<<<grid, kCvThreads, 0, s>>>(a, ppb);
    return check_cuda(cudaGetLastError(), "cost_volume_fwd_kernel");
  }
  const size_t smem = sizeof(tc::Smem) + 128;
  if (int rc = check_cuda(cudaFuncSetAttribute(tc::cost_volume_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(cost_volume_fwd_tc_kernel)")) return rc;
  tc::cost_volume_fwd_tc_kernel<<<grid, kCvThreads, smem, s>>>(a, ppb);
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
  const size_t smem = ((sizeof(CvBwdSmem) + 15) / 16) * 16 + (size_t)kCvThreads * kRowStride * sizeof(float);
  if ((rc = check_cuda(cudaFuncSetAttribute(cost_volume_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(cost_volume_bwd_kernel)"))) return rc;
  (void)HW;
  dim3 grid(patch_blocks(a.H, a.W), (unsigned)((a.D + ppb - 1) / ppb), (unsigned)a.B);
  cost_volume_bwd_kernel<<<grid, kCvThreads, smem, s>>>(a, ppb);
  if ((rc = check_cuda(cudaGetLastError(), "cost_volume_bwd_kernel"))) return rc;
  const size_t total = n_src / 4;
  cv_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(a.dsrc_packed), a.dL_dsrc, HW, total);
  return check_cuda(cudaGetLastError(), "cv_unpack_kernel");
}

}  // namespace fs
