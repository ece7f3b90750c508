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
  const int p = blockIdx.x * kCvThreads + tid;
  if (p >= (int)HW) return;
  const int v = p / W, u = p - v * W;
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
  const float* src_b = a.src_feats + (size_t)b * K * kCvC * HW;
  const unsigned all = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
  const int d0 = blockIdx.y * planes_per_block, d1 = min(a.D, d0 + planes_per_block);
  for (int d = d0; d < d1; d++) {
    const float zd = __ldg(a.planes + d);
    const float X0 = zd * r0, X1 = zd * r1, X2 = zd * r2;
    float x[kCvC];
    float dsum;
    unsigned geo, zero;
    gather_sources(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, all, x, dsum, geo, zero);
    unsigned valid = geo & ~zero;
    if (zero) {   // measure-zero case: redo with the exact set of valid sources (keeps the reference's dot != 0 rule)
      float ds2; unsigned g2, z2;
      gather_sources(src_b, sm.proj, K, H, W, HW, X0, X1, X2, uvx, uvy, cur, valid, x, ds2, g2, z2);
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

int launch_cost_volume_fwd(const FsCostVolumeArgs& a, cudaStream_t s) {
  const size_t HW = (size_t)a.H * a.W;
  const int ppb = 8;
  dim3 grid((unsigned)((HW + kCvThreads - 1) / kCvThreads), (unsigned)((a.D + ppb - 1) / ppb), (unsigned)a.B);
  cost_volume_fwd_kernel<<<grid, kCvThreads, 0, s>>>(a, ppb);
  return check_cuda(cudaGetLastError(), "cost_volume_fwd_kernel");
}

}  // namespace fs

namespace fs {
int launch_cost_volume_bwd(const FsCostVolumeArgs& a, cudaStream_t s) {
  (void)a; (void)s;
  set_error("fs_cost_volume_backward: not implemented yet");
  return FS_ERR_UNSUPPORTED;
}
}  // namespace fs
