// baseline/upstream_proxy.cu -- NOT PRODUCT CODE.  An upstream-STRUCTURED proxy of the reference's CUDA rasterizer,
// written from the algorithm description in SURVEY.md §2.2 / Appendix A (the real extension,
// JonathonLuiten/diff-gaussian-rasterization-w-depth, is not in /root/reference, not installed and not fetchable):
//   per view:  preprocess -> cub::DeviceScan::InclusiveSum -> blocking D2H of num_rendered -> duplicateWithKeys
//              -> cub::DeviceRadixSort::SortPairs (64-bit tile|depth keys) -> identifyTileRanges
//              -> render: one 16x16 CTA per tile, every thread walks EVERY instance of the tile (no culling),
//                 256-instance batches of (id, xy, conic_opacity) in shared memory, rgb/depth read per thread.
// It exists only so that tools/bench_proxy.py can put a "2023-style CUDA recompiled for sm_100" number next to
// ours on the same B200 (labelled as a proxy everywhere it is quoted).  Per-Gaussian math is shared with the
// product header so both produce the same image (checked by the tool).
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../freesplat_b200/csrc/raster_math.cuh"

namespace {

struct Buffers {
  int P = 0, H = 0, W = 0;
  size_t cap = 0;
  float *depths = nullptr, *xy = nullptr, *conic_o = nullptr, *rgb = nullptr;
  int* radii = nullptr;
  uint32_t *tiles = nullptr, *offsets = nullptr;
  uint64_t *keys = nullptr, *keys_sorted = nullptr;
  uint32_t *vals = nullptr, *vals_sorted = nullptr;
  uint2* ranges = nullptr;
  void *scan_tmp = nullptr, *sort_tmp = nullptr;
  size_t scan_bytes = 0, sort_bytes = 0;
} B;

__global__ void pre_kernel(int P, int M, int deg, int H, int W, const float* means, const float* cov6, const float* opac,
                           const float* shs, const float* view /*48 floats*/, float* depths, float* xy, float* conic_o, float* rgb,
                           int* radii, uint32_t* tiles) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  radii[i] = 0; tiles[i] = 0;
  const float s = view[40];
  float mean[3] = {means[3 * i] * s, means[3 * i + 1] * s, means[3 * i + 2] * s};
  float c6[6];
  for (int k = 0; k < 6; k++) c6[k] = cov6[6 * i + k] * (s * s);
  fsm::Projected pr = fsm::project_gaussian(mean, c6, view, view + 16, view[38], view[39], H, W);
  if (pr.radius == 0) return;
  float sh[48];
  for (int k = 0; k < 48; k++) sh[k] = k < M * 3 ? shs[(size_t)i * M * 3 + k] : 0.f;
  float c[3];
  fsm::sh_to_rgb(deg, mean, view + 32, sh, c);
  depths[i] = pr.depth; radii[i] = pr.radius; xy[2 * i] = pr.px; xy[2 * i + 1] = pr.py;
  conic_o[4 * i] = pr.con_x; conic_o[4 * i + 1] = pr.con_y; conic_o[4 * i + 2] = pr.con_z; conic_o[4 * i + 3] = opac[i];
  rgb[3 * i] = c[0]; rgb[3 * i + 1] = c[1]; rgb[3 * i + 2] = c[2];
  tiles[i] = (pr.x1 - pr.x0) * (pr.y1 - pr.y0);
}

__global__ void dup_kernel(int P, int gx, int gy, const float* xy, const float* depths, const uint32_t* offsets, const int* radii,
                           uint64_t* keys, uint32_t* vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P || radii[i] <= 0) return;
  uint32_t off = i ? offsets[i - 1] : 0;
  int x0, y0, x1, y1;
  fsm::get_rect(xy[2 * i], xy[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++) {
      keys[off] = ((uint64_t)(y * gx + x) << 32) | __float_as_uint(depths[i]);
      vals[off] = i; off++;
    }
}

__global__ void ranges_kernel(int R, const uint64_t* keys, uint2* ranges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const uint32_t t = keys[i] >> 32;
  if (i == 0) ranges[t].x = 0;
  else { const uint32_t p = keys[i - 1] >> 32; if (p != t) { ranges[p].y = i; ranges[t].x = i; } }
  if (i == R - 1) ranges[t].y = R;
}

__global__ void __launch_bounds__(256) render_kernel(int H, int W, const uint2* ranges, const uint32_t* point_list, const float* xy,
                                                     const float* conic_o, const float* rgb, const float* depths, const float* bg,
                                                     float* out_color, float* out_depth) {
  __shared__ int ids[256];
  __shared__ float2 sxy[256];
  __shared__ float4 sco[256];
  const int gx = (W + 15) / 16;
  const int px = blockIdx.x * 16 + threadIdx.x, py = blockIdx.y * 16 + threadIdx.y;
  const int tid = threadIdx.y * 16 + threadIdx.x;
  const bool inside = px < W && py < H;
  const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
  bool done = !inside;
  float T = 1.f, C[3] = {0, 0, 0}, D = 0.f;
  int todo = range.y - range.x;
  for (int base = range.x; base < (int)range.y; base += 256, todo -= 256) {
    if (__syncthreads_count(done) == 256) break;
    if (base + tid < (int)range.y) {
      const int id = point_list[base + tid];
      ids[tid] = id; sxy[tid] = make_float2(xy[2 * id], xy[2 * id + 1]);
      sco[tid] = make_float4(conic_o[4 * id], conic_o[4 * id + 1], conic_o[4 * id + 2], conic_o[4 * id + 3]);
    }
    __syncthreads();
    for (int j = 0; !done && j < min(256, todo); j++) {
      const float dx = sxy[j].x - (float)px, dy = sxy[j].y - (float)py;
      const float4 co = sco[j];
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.f) continue;
      const float alpha = fminf(0.99f, co.w * expf(power));
      if (alpha < 1.f / 255.f) continue;
      const float test_T = T * (1.f - alpha);
      if (test_T < 0.0001f) { done = true; continue; }
      const int id = ids[j];
      for (int ch = 0; ch < 3; ch++) C[ch] += rgb[3 * id + ch] * alpha * T;
      D += depths[id] * alpha * T;
      T = test_T;
    }
  }
  if (inside) {
    const size_t HW = (size_t)H * W, pix = (size_t)py * W + px;
    for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = C[ch] + T * bg[ch];
    out_depth[pix] = D;
  }
}

}  // namespace

extern "C" int proxy_setup(int P, int H, int W, size_t cap) {
  B.P = P; B.H = H; B.W = W; B.cap = cap;
  const int tiles = ((W + 15) / 16) * ((H + 15) / 16);
  cudaMalloc(&B.depths, P * 4); cudaMalloc(&B.xy, P * 8); cudaMalloc(&B.conic_o, P * 16); cudaMalloc(&B.rgb, P * 12);
  cudaMalloc(&B.radii, P * 4); cudaMalloc(&B.tiles, P * 4); cudaMalloc(&B.offsets, P * 4);
  cudaMalloc(&B.keys, cap * 8); cudaMalloc(&B.keys_sorted, cap * 8); cudaMalloc(&B.vals, cap * 4); cudaMalloc(&B.vals_sorted, cap * 4);
  cudaMalloc(&B.ranges, tiles * 8);
  cub::DeviceScan::InclusiveSum(nullptr, B.scan_bytes, B.tiles, B.offsets, P);
  cub::DeviceRadixSort::SortPairs(nullptr, B.sort_bytes, B.keys, B.keys_sorted, B.vals, B.vals_sorted, (int)cap, 0, 48);
  cudaMalloc(&B.scan_tmp, B.scan_bytes); cudaMalloc(&B.sort_tmp, B.sort_bytes);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// renders ONE view; returns num_rendered (or -1).  All pointers device.  cov6 [P,6], shs [P,M,3], view: 48-float record.
extern "C" long long proxy_render_view(int M, int deg, const float* means, const float* cov6, const float* opac, const float* shs,
                                       const float* view, float* out_color, float* out_depth, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int P = B.P, H = B.H, W = B.W, gx = (W + 15) / 16, gy = (H + 15) / 16;
  pre_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, M, deg, H, W, means, cov6, opac, shs, view, B.depths, B.xy, B.conic_o, B.rgb, B.radii, B.tiles);
  cub::DeviceScan::InclusiveSum(B.scan_tmp, B.scan_bytes, B.tiles, B.offsets, P, s);
  uint32_t R = 0;
  cudaMemcpyAsync(&R, B.offsets + P - 1, 4, cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);                                   // the per-view host sync of the upstream design
  if (R > B.cap) return -1;
  dup_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, gx, gy, B.xy, B.depths, B.offsets, B.radii, B.keys, B.vals);
  int bits = 0; while ((1 << bits) < gx * gy) bits++;
  cub::DeviceRadixSort::SortPairs(B.sort_tmp, B.sort_bytes, B.keys, B.keys_sorted, B.vals, B.vals_sorted, (int)R, 0, 32 + bits, s);
  cudaMemsetAsync(B.ranges, 0, (size_t)gx * gy * 8, s);
  if (R) ranges_kernel<<<(R + 255) / 256, 256, 0, s>>>((int)R, B.keys_sorted, B.ranges);
  render_kernel<<<dim3(gx, gy), dim3(16, 16), 0, s>>>(H, W, B.ranges, B.vals_sorted, B.xy, B.conic_o, B.rgb, B.depths, view + 35, out_color, out_depth);
  return cudaGetLastError() == cudaSuccess ? (long long)R : -1;
}
