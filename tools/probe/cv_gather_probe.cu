// cv_gather_probe.cu -- measurement behind DESIGN.md §4 "cost-volume gather": how fast can the bilinear 4-tap x 48-channel
// gather of the plane-sweep cost volume be fed
//   (A) the product's way: per-thread 16-byte loads from the channel-packed source maps [K][12][H*W][4] through L1 / L2,
//   (B) north_star's way: a TMA box (cp.async.bulk.tensor.4d) of the source patch that the CTA's 128 pixels x a chunk of
//       planes can touch, staged in shared memory (double-buffered: the next box is in flight while the current one is read),
//       then the same taps as LDS.128.
// Both kernels do the same arithmetic per (pixel, plane, source): 4 taps x 12 channel groups, bilinear weights, 48 fma into the
// warped feature vector, dot with the reference features; the MLP is left out (it is identical in both).  Geometry: config 3's
// (120 x 160 feature maps, 128 planes uniform in inverse depth between 0.5 and 15 m, focal 144 px, baselines 0.25 m / 0.5 m:
// 0.55 / 1.1 px of disparity per plane), patch = 32 x 4 pixels per 128-thread CTA, 16 planes per CTA as in the product kernel.
// Dummy dynamic shared memory reproduces the product's occupancy: A runs at 2 CTAs / SM (80 KB of operand tiles each), B at
// 1 CTA / SM (80 KB + two boxes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/cv_gather_probe tools/probe/cv_gather_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int H = 120, W = 160, D = 128, K = 2, G = 12, PPB = 16, PC = 4;
constexpr int BW = 48, BH = 6;                       // TMA box in texels: 32 + slide of PC planes + bilinear + alignment, 4 + 1 + 1
constexpr int kBoxBytes = G * BH * BW * 16;          // 55 296 B
constexpr float kFocal = 144.f, kNear = 0.5f, kFar = 15.f;

__host__ __device__ inline float plane_invdepth(int d) { return 1.f / kNear + (1.f / kFar - 1.f / kNear) * ((float)d / (float)(D - 1)); }
// affine model of the homography for a fronto-parallel plane and a sideways baseline: source x = u + f B / z, source y = v
__host__ __device__ inline float src_x(float u, int d, int k) { return u + kFocal * (0.25f * (float)(k + 1)) * plane_invdepth(d) - 40.f * (float)(k + 1); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Taps { int x0, y0; float w00, w01, w10, w11; };
__device__ __forceinline__ Taps make_taps(float ix, float iy) {
  Taps t;
  const float xf = floorf(ix), yf = floorf(iy);
  t.x0 = (int)xf; t.y0 = (int)yf;
  const float wx1 = ix - xf, wx0 = 1.f - wx1, wy1 = iy - yf, wy0 = 1.f - wy1;
  t.w00 = wx0 * wy0; t.w01 = wx1 * wy0; t.w10 = wx0 * wy1; t.w11 = wx1 * wy1;
  return t;
}

// ------------------------------------------------------------------------------------------------ (A) global loads
__global__ void __launch_bounds__(128) gather_ldg(const float4* __restrict__ src, const float* __restrict__ cur, float* __restrict__ out) {
  const int tid = threadIdx.x;
  const int px = W / 32;
  const int by = blockIdx.x / px, bx = blockIdx.x - by * px;
  const int u = bx * 32 + (tid & 31), v = by * 4 + (tid >> 5);
  float c[48];
#pragma unroll
  for (int i = 0; i < 48; i++) c[i] = cur[(size_t)i * H * W + v * W + u];
  const int d0 = blockIdx.y * PPB;
  for (int d = d0; d < d0 + PPB; d++) {
    float acc = 0.f;
    for (int k = 0; k < K; k++) {
      const Taps t = make_taps(src_x((float)u, d, k), (float)v + 0.25f);
      const bool inb = t.x0 >= 0 && t.x0 + 1 < W && t.y0 >= 0 && t.y0 + 1 < H;
      if (!inb) continue;
      const float4* s = src + (size_t)k * G * H * W + t.y0 * W + t.x0;
      float x[48];
#pragma unroll
      for (int g = 0; g < G; g++) {
        const float4* sg = s + (size_t)g * H * W;
        const float4 a = __ldg(sg), b = __ldg(sg + 1), cc = __ldg(sg + W), e = __ldg(sg + W + 1);
        x[4 * g] = t.w00 * a.x + t.w01 * b.x + t.w10 * cc.x + t.w11 * e.x;
        x[4 * g + 1] = t.w00 * a.y + t.w01 * b.y + t.w10 * cc.y + t.w11 * e.y;
        x[4 * g + 2] = t.w00 * a.z + t.w01 * b.z + t.w10 * cc.z + t.w11 * e.z;
        x[4 * g + 3] = t.w00 * a.w + t.w01 * b.w + t.w10 * cc.w + t.w11 * e.w;
      }
#pragma unroll
      for (int i = 0; i < 48; i++) acc = fmaf(x[i], c[i], acc);
    }
    out[(size_t)d * H * W + v * W + u] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ (B) TMA boxes
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(bar),
               "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128) gather_tma(const __grid_constant__ CUtensorMap tmap, const float4* __restrict__ src,
                                                  const float* __restrict__ cur, float* __restrict__ out, int double_buffer) {
  extern __shared__ __align__(128) unsigned char raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)raw + 127) & ~(uintptr_t)127);
  float4* box[2] = {reinterpret_cast<float4*>(base), reinterpret_cast<float4*>(base + kBoxBytes)};
  __shared__ unsigned long long bars[2];
  const int tid = threadIdx.x;
  const int px = W / 32;
  const int by = blockIdx.x / px, bx = blockIdx.x - by * px;
  const int u0 = bx * 32, v0 = by * 4;
  const int u = u0 + (tid & 31), v = v0 + (tid >> 5);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float c[48];
#pragma unroll
  for (int i = 0; i < 48; i++) c[i] = cur[(size_t)i * H * W + v * W + u];
  const int d0 = blockIdx.y * PPB;
  // work units: (chunk of PC planes, source k); the box covers what the patch touches over the chunk's planes
  constexpr int kUnits = (PPB / PC) * K;
  auto unit_origin = [&](int unit, int& ox, int& oy, int& kk, int& dc) {
    dc = d0 + (unit / K) * PC; kk = unit % K;
    // the disparity shrinks with the plane index: the left-most sample of the chunk belongs to its LAST plane
    ox = (int)floorf(src_x((float)u0, dc + PC - 1, kk)); oy = v0;
  };
  auto issue = [&](int unit, int buf) {
    int ox, oy, kk, dc;
    unit_origin(unit, ox, oy, kk, dc);
    const uint32_t bar = smem_u32(&bars[buf]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kBoxBytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(box[buf])),
                 "l"(&tmap), "r"(0), "r"(ox), "r"(oy), "r"(kk * G), "r"(bar) : "memory");
  };
  uint32_t phase[2] = {0u, 0u};
  float acc[PC];
  if (tid == 0) issue(0, 0);
  for (int unit = 0; unit < kUnits; unit++) {
    const int buf = double_buffer ? (unit & 1) : 0;
    if (double_buffer && tid == 0 && unit + 1 < kUnits) issue(unit + 1, buf ^ 1);     // next box in flight under this unit's gather
    int ox, oy, kk, dc;
    unit_origin(unit, ox, oy, kk, dc);
    if (kk == 0) {
#pragma unroll
      for (int i = 0; i < PC; i++) acc[i] = 0.f;
    }
    mbar_wait(smem_u32(&bars[buf]), phase[buf]); phase[buf] ^= 1u;
    const float4* bxp = box[buf];
#pragma unroll
    for (int i = 0; i < PC; i++) {
      const int d = dc + i;
      const Taps t = make_taps(src_x((float)u, d, kk), (float)v + 0.25f);
      const bool inb = t.x0 >= 0 && t.x0 + 1 < W && t.y0 >= 0 && t.y0 + 1 < H;
      const int lx = t.x0 - ox, ly = t.y0 - oy;
      if (!inb || lx < 0 || lx + 1 >= BW || ly < 0 || ly + 1 >= BH) continue;         // (the product would fall back to global loads)
      const float4* s = bxp + ly * BW + lx;
      float x[48];
#pragma unroll
      for (int g = 0; g < G; g++) {
        const float4* sg = s + g * BH * BW;
        const float4 a = sg[0], b = sg[1], cc = sg[BW], e = sg[BW + 1];
        x[4 * g] = t.w00 * a.x + t.w01 * b.x + t.w10 * cc.x + t.w11 * e.x;
        x[4 * g + 1] = t.w00 * a.y + t.w01 * b.y + t.w10 * cc.y + t.w11 * e.y;
        x[4 * g + 2] = t.w00 * a.z + t.w01 * b.z + t.w10 * cc.z + t.w11 * e.z;
        x[4 * g + 3] = t.w00 * a.w + t.w01 * b.w + t.w10 * cc.w + t.w11 * e.w;
      }
      float dot = 0.f;
#pragma unroll
      for (int q = 0; q < 48; q++) dot = fmaf(x[q], c[q], dot);
      acc[i] += dot;
    }
    __syncthreads();                                   // everyone is done with this box before it is refilled
    if (!double_buffer && tid == 0 && unit + 1 < kUnits) issue(unit + 1, 0);
    if (kk == K - 1) {
#pragma unroll
      for (int i = 0; i < PC; i++) out[(size_t)(dc + i) * H * W + v * W + u] = acc[i];
    }
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
  const size_t HW = (size_t)H * W;
  std::vector<float> h_src((size_t)K * G * HW * 4), h_cur(48 * HW);
  srand(1);
  for (auto& x : h_src) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto& x : h_cur) x = (float)rand() / RAND_MAX - 0.5f;
  float4* d_src; float *d_cur, *d_outA, *d_outB;
  CK(cudaMalloc(&d_src, h_src.size() * 4)); CK(cudaMalloc(&d_cur, h_cur.size() * 4));
  CK(cudaMalloc(&d_outA, (size_t)D * HW * 4)); CK(cudaMalloc(&d_outB, (size_t)D * HW * 4));
  CK(cudaMemcpy(d_src, h_src.data(), h_src.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_cur, h_cur.data(), h_cur.size() * 4, cudaMemcpyHostToDevice));
  // tensor map over the packed maps: [4 floats][W][H][K * 12 groups]
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)K * G};
  const cuuint64_t gstride[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)HW * 16};
  const cuuint32_t bdim[4] = {4, BW, BH, G};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d_src, gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); return 1; }
  dim3 grid((W / 32) * (H / 4), D / PPB);
  const int smemA = 80 * 1024, smemB1 = 80 * 1024 + kBoxBytes + 256, smemB2 = 80 * 1024 + 2 * kBoxBytes + 256;
  CK(cudaFuncSetAttribute(gather_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, smemA));
  CK(cudaFuncSetAttribute(gather_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smemB2));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float msA = 0, msB1 = 0, msB2 = 0, msA0 = 0;
  const int reps = 20;
  for (int variant = 0; variant < 4; variant++) {
    for (int it = 0; it < reps + 3; it++) {
      if (it == 3) CK(cudaEventRecord(e0));
      if (variant == 0) gather_ldg<<<grid, 128, smemA>>>(d_src, d_cur, d_outA);
      else if (variant == 1) gather_ldg<<<grid, 128, 0>>>(d_src, d_cur, d_outA);
      else gather_tma<<<grid, 128, variant == 2 ? smemB1 : smemB2>>>(tmap, d_src, d_cur, d_outB, variant == 3);
    }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    (variant == 0 ? msA : variant == 1 ? msA0 : variant == 2 ? msB1 : msB2) = ms / reps;
  }
  std::vector<float> a((size_t)D * HW), b((size_t)D * HW);
  CK(cudaMemcpy(a.data(), d_outA, a.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), d_outB, b.size() * 4, cudaMemcpyDeviceToHost));
  double maxd = 0, maxa = 0; size_t nz = 0;
  for (size_t i = 0; i < a.size(); i++) { maxd = fmax(maxd, fabs((double)a[i] - b[i])); maxa = fmax(maxa, fabs((double)a[i])); nz += a[i] != 0.f; }
  const double rows = (double)D * HW * K;
  printf("{\"workload\": \"1 reference view, K=2, 48x120x160, D=128: gather + bilinear + dot only\", \"rows\": %.0f,\n"
         " \"ldg_2cta_per_sm_ms\": %.4f, \"ldg_max_occupancy_ms\": %.4f, \"tma_single_buffer_1cta_ms\": %.4f, \"tma_double_buffer_1cta_ms\": %.4f,\n"
         " \"box_bytes\": %d, \"max_abs_diff\": %.3g, \"max_abs\": %.3g, \"nonzero_frac\": %.3f}\n",
         rows, msA, msA0, msB1, msB2, kBoxBytes, maxd, maxa, (double)nz / a.size());
  return 0;
}
