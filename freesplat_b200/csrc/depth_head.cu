// depth_head.cu -- tail of the depth-regression head.
//
// Replaces the last lines of DepthDecoder.forward
// (/root/reference/src/model/encoder/modules/networks.py:130-152): per scale
//     planes = softmax(logits, dim=1) ; E = sum_d candi[d] * planes[d] ; depth = exp(E)  (or 1/E)
// and, for scale 0,
//     depth_up   = exp( bilinear x2 (align_corners=True) of E )
//     weights_up = max_d ( bilinear x2 (align_corners=True) of planes[d] )
// The reference materialises softmax [B,D,h,w], its x2 upsampling [B,D,2h,2w] (4x the logits) and reduces it
// again: ~12 passes over the logits' bytes.  Here the logits are read from HBM ONCE:
//   * depth_head_up_kernel: CTA = 32x4 coarse pixels (+1 halo row / column).  The [D][5][36] logit tile is staged
//     in shared memory by TMA box loads (cp.async.bulk.tensor.3d, one mbarrier per 32-plane chunk; out-of-image
//     elements are zero-filled by the hardware), normalised in place (two sweeps per pixel), and every coarse cell
//     then produces its ~2x2 output pixels straight from the tile: per PAIR of planes 8 shared loads, 16 packed
//     FFMA2 / FMUL2 (fma.rn.f32x2) and 4 three-input maxima (FMNMX3).
//     256 threads = 128 cells x 2 halves of the plane range; 2 CTAs per SM (93 KB each) overlap load and math.
//   * depth_head_expect_kernel: scales 1..3 (no upsampling): thread per pixel, online softmax, coalesced.
// Algorithmic HBM bytes: 4*D*h*w read + (2 + 8) * 4*h*w written per view.
//
// Bilinear rule = aten's (UpSampleKernel.cpp compute_indices_weights, align_corners=True): scale = (in-1)/(out-1)
// in fp32, src = scale*dst, i0 = min(int(src), in-1), lam = clamp(src-i0, 0, 1), i1 = min(i0+1, in-1).
#include <cuda.h>
#include <cstring>

#include "common.cuh"

namespace fs {
namespace dh {

constexpr int kCoreX = 32, kCoreY = 4;
constexpr int kBoxX = 36, kBoxY = 5;          // core + 1 halo, x padded to a multiple of 16 bytes
constexpr int kPlane = kBoxX * kBoxY;         // 180 floats per plane of the tile
constexpr int kChunkD = 32;                   // planes per TMA box
constexpr int kMaxD = 128;
constexpr int kThreads = 256;
constexpr uint32_t kChunkBytes = kChunkD * kPlane * sizeof(float);

struct __align__(128) Smem {
  float tile[kMaxD * kPlane];                 // TMA destination: [plane][5][36]
  float candi[kMaxD];
  float rs[kPlane];                           // 1 / sum_d exp(l - max) per tile pixel
  float E[kPlane];                            // expectation per tile pixel
  float wmax[4][128];                         // partial maxima of the upper plane half
  unsigned long long bar[kMaxD / kChunkD];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "DH_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DH_DONE;\n\t"
      "bra DH_WAIT;\n\t"
      "DH_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Blackwell packed fp32 pairs (FFMA2 / FMUL2) and the 3-input maximum (FMNMX3): phase 2 evaluates two planes per instruction
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float max3(float a, unsigned long long v) {
  float lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(lo), "f"(hi));
  return r;
}

// source index / interpolation weight of output index o (aten rule, align_corners=True)
__device__ __forceinline__ int src_index(float scale, int o, int n_in) { return min((int)(scale * (float)o), n_in - 1); }
__device__ __forceinline__ float src_lambda(float scale, int o, int i0) { return fminf(fmaxf(scale * (float)o - (float)i0, 0.f), 1.f); }
// smallest output index in [0, n_out] whose source index is >= i  (n_out when there is none)
__device__ __forceinline__ int first_out(float scale, int i, int n_in, int n_out) {
  if (i <= 0) return 0;
  if (i >= n_in || !(scale > 0.f)) return n_out;
  int o = min(max((int)((float)i / scale), 0), n_out);
  while (o > 0 && src_index(scale, o - 1, n_in) >= i) o--;
  while (o < n_out && src_index(scale, o, n_in) < i) o++;
  return o;
}

template <bool kTma>
__global__ void __launch_bounds__(kThreads, 2) depth_head_up_kernel(const __grid_constant__ CUtensorMap tmap, FsDepthHeadArgs a) {
  extern __shared__ unsigned char dh_smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(dh_smem_raw) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x;
  const int b = blockIdx.z, ty0 = blockIdx.y * kCoreY, tx0 = blockIdx.x * kCoreX;
  const int D = a.D, h = a.h, w = a.w, Ho = 2 * a.h, Wo = 2 * a.w;
  const int nchunk = (D + kChunkD - 1) / kChunkD;

  // ---------------------------------------------------------------- stage the logit tile
  if (kTma) {
    if (tid == 0) {
      for (int c = 0; c < nchunk; c++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar[c])) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (int c = 0; c < nchunk; c++) {
        const uint32_t bar = smem_u32(&sm.bar[c]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kChunkBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                smem_u32(sm.tile + c * kChunkD * kPlane)),
            "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(bar), "r"(tx0), "r"(ty0), "r"(b * D + c * kChunkD)
            : "memory");
      }
    }
  } else {
    const float* lb = a.logits + (size_t)b * D * h * w;
    for (int e = tid; e < D * kPlane; e += kThreads) {
      const int d = e / kPlane, i = e - d * kPlane, yy = i / kBoxX, xx = i - yy * kBoxX;
      const int gy = ty0 + yy, gx = tx0 + xx;
      sm.tile[e] = (gy < h && gx < w) ? __ldg(lb + ((size_t)d * h + gy) * w + gx) : 0.f;
    }
  }
  for (int d = tid; d < D; d += kThreads) sm.candi[d] = a.candi[d];
  __syncthreads();                                                     // barrier inits, candi (and the LDG-staged tile) visible

  // ---------------------------------------------------------------- phase 1: softmax statistics per tile pixel
  {
    const int yy = tid / kBoxX, xx = tid - yy * kBoxX;
    const int gy = ty0 + yy, gx = tx0 + xx;
    const bool mine = tid < kPlane && xx <= kCoreX && gy < h && gx < w;
    float* tp = sm.tile + tid;
    float m0 = -INFINITY, m1 = -INFINITY;                                // two chains: the sweeps are latency bound
    for (int c = 0; c < nchunk; c++) {
      if (kTma) mbar_wait(smem_u32(&sm.bar[c]), 0u);
      if (mine) {
        const int dend = min(D, (c + 1) * kChunkD);
        const float* q = tp + c * kChunkD * kPlane;
        int d = c * kChunkD;
#pragma unroll 4
        for (; d + 2 <= dend; d += 2, q += 2 * kPlane) { m0 = fmaxf(m0, q[0]); m1 = fmaxf(m1, q[kPlane]); }
        if (d < dend) m0 = fmaxf(m0, q[0]);
      }
    }
    if (mine) {
      // e = exp(l - m) as ex2(l*log2e - m*log2e): one FFMA + one MUFU.EX2 per plane (|rel err| < 1e-6, inside the 1e-4 budget)
      constexpr float kLog2e = 1.4426950408889634f;
      const float m = fmaxf(m0, m1), mb = -m * kLog2e;
      float s0 = 0.f, s1 = 0.f, a0 = 0.f, a1 = 0.f;
      float* q = tp;
      int d = 0;
#pragma unroll 4
      for (; d + 2 <= D; d += 2, q += 2 * kPlane) {
        float e0, e1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(q[0], kLog2e, mb)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(q[kPlane], kLog2e, mb)));
        s0 += e0; s1 += e1;
        a0 = fmaf(sm.candi[d], e0, a0); a1 = fmaf(sm.candi[d + 1], e1, a1);
        q[0] = e0; q[kPlane] = e1;
      }
      if (d < D) {
        float e0;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(q[0], kLog2e, mb)));
        s0 += e0; a0 = fmaf(sm.candi[d], e0, a0); q[0] = e0;
      }
      const float s = s0 + s1, acc = a0 + a1;
      const float rs = 1.0f / s, E = acc * rs;
      sm.rs[tid] = rs; sm.E[tid] = E;
      if (yy < kCoreY && xx < kCoreX) {
        const size_t o = ((size_t)b * h + gy) * w + gx;
        a.expect[o] = E;
        a.depth[o] = a.log_planes ? expf(E) : 1.0f / E;
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 2: the output pixels of every coarse cell
  const int cell = tid & 127, half = tid >> 7;
  const int cy = cell >> 5, cx = cell & 31;
  const int h1 = ty0 + cy, w1 = tx0 + cx;
  const bool cell_ok = h1 < h && w1 < w;
  const float rh = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f, rw = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  int Y0 = 0, X0 = 0, nr = 0, nc = 0;
  if (cell_ok) {
    Y0 = first_out(rh, h1, h, Ho); nr = first_out(rh, h1 + 1, h, Ho) - Y0;
    X0 = first_out(rw, w1, w, Wo); nc = first_out(rw, w1 + 1, w, Wo) - X0;
  }
  const int h1p = (h1 < h - 1) ? 1 : 0, w1p = (w1 < w - 1) ? 1 : 0;
  const int i00 = cy * kBoxX + cx, i01 = i00 + w1p, i10 = i00 + h1p * kBoxX, i11 = i10 + w1p;
  const int dsplit = (D + 1) >> 1;
  const int dbeg = half ? dsplit : 0, dend = half ? D : dsplit;
  // a cell owns 2x2 output pixels, rarely 3 along an axis ((out-1)/(in-1) is slightly above 2): 2x2 sub-blocks
  for (int it = 0; it < 4; it++) {
    const int ry = (it >> 1) * 2, rx = (it & 1) * 2;
    const bool mine = cell_ok && ry < nr && rx < nc;
    if (it > 0 && !__syncthreads_or(mine)) continue;                 // CTA-uniform: extra sub-blocks are rare
    float wgt[4][4], mx[4] = {0.f, 0.f, 0.f, 0.f};
    float ly[2], lx[2];
    if (mine) {
      const float r00 = sm.rs[i00], r01 = sm.rs[i01], r10 = sm.rs[i10], r11 = sm.rs[i11];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        ly[j] = src_lambda(rh, min(Y0 + ry + j, Ho - 1), h1);
        lx[j] = src_lambda(rw, min(X0 + rx + j, Wo - 1), w1);
      }
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const float wy1 = ly[p >> 1], wy0 = 1.f - wy1, wx1 = lx[p & 1], wx0 = 1.f - wx1;
        wgt[p][0] = wy0 * wx0 * r00; wgt[p][1] = wy0 * wx1 * r01; wgt[p][2] = wy1 * wx0 * r10; wgt[p][3] = wy1 * wx1 * r11;
      }
      const float* tp = sm.tile;
      unsigned long long wq[4][4];                        // tap weights duplicated into both halves of a pair
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int t = 0; t < 4; t++) wq[p][t] = pack2(wgt[p][t], wgt[p][t]);
      int d = dbeg;
#pragma unroll 2
      for (; d + 2 <= dend; d += 2) {                     // planes d and d+1 per packed instruction
        const unsigned long long e00 = pack2(tp[d * kPlane + i00], tp[(d + 1) * kPlane + i00]);
        const unsigned long long e01 = pack2(tp[d * kPlane + i01], tp[(d + 1) * kPlane + i01]);
        const unsigned long long e10 = pack2(tp[d * kPlane + i10], tp[(d + 1) * kPlane + i10]);
        const unsigned long long e11 = pack2(tp[d * kPlane + i11], tp[(d + 1) * kPlane + i11]);
#pragma unroll
        for (int p = 0; p < 4; p++)
          mx[p] = max3(mx[p], fma2(wq[p][3], e11, fma2(wq[p][2], e10, fma2(wq[p][1], e01, mul2(wq[p][0], e00)))));
      }
      for (; d < dend; d++) {
        const float e00 = tp[d * kPlane + i00], e01 = tp[d * kPlane + i01], e10 = tp[d * kPlane + i10], e11 = tp[d * kPlane + i11];
#pragma unroll
        for (int p = 0; p < 4; p++)
          mx[p] = fmaxf(mx[p], fmaf(wgt[p][3], e11, fmaf(wgt[p][2], e10, fmaf(wgt[p][1], e01, wgt[p][0] * e00))));
      }
      if (half == 1) {
#pragma unroll
        for (int p = 0; p < 4; p++) sm.wmax[p][cell] = mx[p];
      }
    }
    __syncthreads();
    if (mine && half == 0) {
      const float E00 = sm.E[i00], E01 = sm.E[i01], E10 = sm.E[i10], E11 = sm.E[i11];
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const int Y = Y0 + ry + (p >> 1), X = X0 + rx + (p & 1);
        if (ry + (p >> 1) < nr && rx + (p & 1) < nc) {
          const float wy1 = ly[p >> 1], wy0 = 1.f - wy1, wx1 = lx[p & 1], wx0 = 1.f - wx1;
          const float fine = wy0 * (wx0 * E00 + wx1 * E01) + wy1 * (wx0 * E10 + wx1 * E11);
          const size_t o = ((size_t)b * Ho + Y) * Wo + X;
          a.depth_up[o] = a.log_planes ? expf(fine) : 1.0f / fine;
          a.weights_up[o] = D > 1 ? fmaxf(mx[p], sm.wmax[p][cell]) : mx[p];
        }
      }
    }
  }
}

// scales without upsampling: thread per pixel, one pass (online softmax), loads coalesced across the warp
__global__ void __launch_bounds__(256) depth_head_expect_kernel(FsDepthHeadArgs a) {
  const size_t HW = (size_t)a.h * a.w;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= HW) return;
  const int b = blockIdx.y;
  const float* lp = a.logits + (size_t)b * a.D * HW + i;
  float m = -INFINITY, s = 0.f, acc = 0.f;
  int d = 0;
  for (; d + 8 <= a.D; d += 8) {
    float l[8];
#pragma unroll
    for (int j = 0; j < 8; j++) l[j] = __ldg(lp + (size_t)(d + j) * HW);
    float mn = m;
#pragma unroll
    for (int j = 0; j < 8; j++) mn = fmaxf(mn, l[j]);
    const float sc = __expf(m - mn);                 // 0 on the first group (m = -inf)
    s *= sc; acc *= sc; m = mn;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float e = __expf(l[j] - m);
      s += e; acc = fmaf(__ldg(a.candi + d + j), e, acc);
    }
  }
  for (; d < a.D; d++) {
    const float l = __ldg(lp + (size_t)d * HW);
    const float mn = fmaxf(m, l);
    const float sc = __expf(m - mn), e = __expf(l - mn);
    s = s * sc + e; acc = fmaf(__ldg(a.candi + d), e, acc * sc); m = mn;
  }
  const float E = acc / s;
  a.expect[(size_t)b * HW + i] = E;
  a.depth[(size_t)b * HW + i] = a.log_planes ? expf(E) : 1.0f / E;
}

// ---------------------------------------------------------------------------------------------------- backward (training)
// The reference differentiates the tail through softmax, the expectation, exp / reciprocal, two bilinear x2 upsamplings and
// a max over the upsampled planes (networks.py:130-152).  Here: (1) per coarse pixel the softmax statistics (max, sum exp,
// expectation) are recomputed in one coalesced pass; (2) if depth_weights received a gradient, the plane that attains the
// maximum of every fine pixel is found (4 taps x D planes from L2); (3) one thread per coarse pixel gathers what reaches it
// -- its own expectation / depth gradients, the <= 6x6 fine pixels whose bilinear taps include it (aten's align_corners
// rule, the same fp32 arithmetic as the forward) -- and writes the whole logit column:
//   d logit_d = p_d (candi_d dE + dP_d - sum_j p_j (candi_j dE + dP_j)),  dP_d = sum over fine pixels with arg max d of g w_tap.
__device__ __forceinline__ void up_taps(int dst, int n_in, int n_out, int& i0, int& i1, float& w0, float& w1) {
  const float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
  const float src = scale * (float)dst;
  i0 = min((int)src, n_in - 1);
  const float lam = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  i1 = min(i0 + 1, n_in - 1);
  w0 = 1.f - lam; w1 = lam;
}

__global__ void __launch_bounds__(256) depth_head_stats_kernel(FsDepthHeadBwdArgs a) {
  const size_t HW = (size_t)a.h * a.w;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= HW) return;
  const int b = blockIdx.y;
  const float* lp = a.logits + (size_t)b * a.D * HW + i;
  float m = -INFINITY;
  for (int d = 0; d < a.D; d++) m = fmaxf(m, __ldg(lp + (size_t)d * HW));
  float s = 0.f, acc = 0.f;
  for (int d = 0; d < a.D; d++) {
    const float e = expf(__ldg(lp + (size_t)d * HW) - m);
    s += e; acc = fmaf(__ldg(a.candi + d), e, acc);
  }
  float* st = a.stats + 3 * ((size_t)b * HW + i);
  st[0] = m; st[1] = s; st[2] = acc / s;
}

__global__ void __launch_bounds__(256) depth_head_argmax_kernel(FsDepthHeadBwdArgs a) {
  const int Ho = 2 * a.h, Wo = 2 * a.w;
  const size_t HWo = (size_t)Ho * Wo, HW = (size_t)a.h * a.w;
  const size_t f = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (f >= HWo) return;
  const int b = blockIdx.y;
  const int Y = (int)(f / Wo), X = (int)(f - (size_t)Y * Wo);
  int y0, y1, x0, x1; float wy0, wy1, wx0, wx1;
  up_taps(Y, a.h, Ho, y0, y1, wy0, wy1);
  up_taps(X, a.w, Wo, x0, x1, wx0, wx1);
  const size_t t[4] = {(size_t)y0 * a.w + x0, (size_t)y0 * a.w + x1, (size_t)y1 * a.w + x0, (size_t)y1 * a.w + x1};
  float mm[4], ws[4];
  const float wt[4] = {wy0 * wx0, wy0 * wx1, wy1 * wx0, wy1 * wx1};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float* st = a.stats + 3 * ((size_t)b * HW + t[k]);
    mm[k] = st[0]; ws[k] = wt[k] / st[1];
  }
  const float* lp = a.logits + (size_t)b * a.D * HW;
  float best = -INFINITY;
  int arg = 0;
  for (int d = 0; d < a.D; d++) {
    // bilinear of the softmax planes in aten's order: (top-left wx0 + top-right wx1) wy0 + (bottom ...) wy1
    const float* l = lp + (size_t)d * HW;
    const float p00 = expf(__ldg(l + t[0]) - mm[0]) * ws[0], p01 = expf(__ldg(l + t[1]) - mm[1]) * ws[1];
    const float p10 = expf(__ldg(l + t[2]) - mm[2]) * ws[2], p11 = expf(__ldg(l + t[3]) - mm[3]) * ws[3];
    const float v = (p00 + p01) + (p10 + p11);
    if (v > best) { best = v; arg = d; }
  }
  a.argmax_up[(size_t)b * HWo + f] = (uint8_t)arg;
}

__global__ void __launch_bounds__(128) depth_head_bwd_kernel(FsDepthHeadBwdArgs a) {
  const size_t HW = (size_t)a.h * a.w;
  const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= HW) return;
  const int b = blockIdx.y;
  const int y = (int)(i / a.w), x = (int)(i - (size_t)y * a.w);
  const float* stb = a.stats + 3 * (size_t)b * HW;
  const float m = stb[3 * i], s = stb[3 * i + 1], E = stb[3 * i + 2];
  float dE = a.g_expect ? a.g_expect[(size_t)b * HW + i] : 0.f;
  if (a.g_depth) dE += a.g_depth[(size_t)b * HW + i] * (a.log_planes ? expf(E) : -1.0f / (E * E));
  // contributions of the fine grid (scale 0 only)
  constexpr int kMaxList = 36;
  int ld[kMaxList];
  float lc[kMaxList];
  int nl = 0;
  if (a.upsample && (a.g_depth_up || a.g_weights_up)) {
    const int Ho = 2 * a.h, Wo = 2 * a.w;
    const size_t HWo = (size_t)Ho * Wo;
    for (int Y = max(0, 2 * y - 2); Y <= min(Ho - 1, 2 * y + 3); Y++) {
      int y0, y1; float wy0, wy1;
      up_taps(Y, a.h, Ho, y0, y1, wy0, wy1);
      const float wy = (y0 == y ? wy0 : 0.f) + (y1 == y ? wy1 : 0.f);
      if (wy == 0.f && y0 != y && y1 != y) continue;
      for (int X = max(0, 2 * x - 2); X <= min(Wo - 1, 2 * x + 3); X++) {
        int x0, x1; float wx0, wx1;
        up_taps(X, a.w, Wo, x0, x1, wx0, wx1);
        if (x0 != x && x1 != x) continue;
        const float wx = (x0 == x ? wx0 : 0.f) + (x1 == x ? wx1 : 0.f);
        const float wgt = wy * wx;
        const size_t f = (size_t)b * HWo + (size_t)Y * Wo + X;
        if (a.g_depth_up) {
          const float fine = (stb[3 * ((size_t)y0 * a.w + x0) + 2] * wx0 + stb[3 * ((size_t)y0 * a.w + x1) + 2] * wx1) * wy0 +
                             (stb[3 * ((size_t)y1 * a.w + x0) + 2] * wx0 + stb[3 * ((size_t)y1 * a.w + x1) + 2] * wx1) * wy1;
          dE += a.g_depth_up[f] * (a.log_planes ? expf(fine) : -1.0f / (fine * fine)) * wgt;
        }
        if (a.g_weights_up && nl < kMaxList) {
          const float g = a.g_weights_up[f] * wgt;
          if (g != 0.f) { ld[nl] = (int)a.argmax_up[f]; lc[nl] = g; nl++; }
        }
      }
    }
  }
  const float* lp = a.logits + (size_t)b * a.D * HW + i;
  float* dp = a.d_logits + (size_t)b * a.D * HW + i;
  const float inv_s = 1.0f / s;
  float S = dE * E;
  for (int k = 0; k < nl; k++) {
    const float p = expf(__ldg(lp + (size_t)ld[k] * HW) - m) * inv_s;
    lc[k] *= p;                      // p_{d*} g w
    S += lc[k];
  }
  for (int d = 0; d < a.D; d++) {
    const float p = expf(__ldg(lp + (size_t)d * HW) - m) * inv_s;
    dp[(size_t)d * HW] = p * (__ldg(a.candi + d) * dE - S);
  }
  for (int k = 0; k < nl; k++) dp[(size_t)ld[k] * HW] += lc[k];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_tile_map(const FsDepthHeadArgs& a, CUtensorMap* out) {
  static EncodeTiledFn fn = nullptr;                 // resolved once through the runtime: no link-time libcuda dependency
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (int rc = check_cuda(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q), "cudaGetDriverEntryPoint")) return rc;
    if (q != cudaDriverEntryPointSuccess || !p) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return FS_ERR_CUDA; }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t gdim[3] = {(cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.B * a.D};
  const cuuint64_t gstride[2] = {(cuuint64_t)a.w * 4, (cuuint64_t)a.h * a.w * 4};
  const cuuint32_t box[3] = {kBoxX, kBoxY, kChunkD};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.logits), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return FS_ERR_CUDA; }
  return FS_OK;
}

}  // namespace dh

int launch_depth_head_bwd(const FsDepthHeadBwdArgs& a, cudaStream_t s) {
  using namespace dh;
  if (a.B == 0) return FS_OK;
  const size_t HW = (size_t)a.h * a.w;
  dim3 g1((unsigned)((HW + 255) / 256), (unsigned)a.B);
  depth_head_stats_kernel<<<g1, 256, 0, s>>>(a);
  int rc;
  if ((rc = check_cuda(cudaGetLastError(), "depth_head_stats_kernel"))) return rc;
  if (a.upsample && a.g_weights_up) {
    dim3 g2((unsigned)((4 * HW + 255) / 256), (unsigned)a.B);
    depth_head_argmax_kernel<<<g2, 256, 0, s>>>(a);
    if ((rc = check_cuda(cudaGetLastError(), "depth_head_argmax_kernel"))) return rc;
  }
  dim3 g3((unsigned)((HW + 127) / 128), (unsigned)a.B);
  depth_head_bwd_kernel<<<g3, 128, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "depth_head_bwd_kernel");
}

int launch_depth_head(const FsDepthHeadArgs& a, cudaStream_t s) {
  using namespace dh;
  if (a.B == 0) return FS_OK;
  if (!a.upsample) {
    dim3 grid((unsigned)(((size_t)a.h * a.w + 255) / 256), (unsigned)a.B);
    depth_head_expect_kernel<<<grid, 256, 0, s>>>(a);
    return check_cuda(cudaGetLastError(), "depth_head_expect_kernel");
  }
  if (a.D > kMaxD) { set_error("depth head: at most %d planes in the fused upsampling kernel (got %d)", kMaxD, a.D); return FS_ERR_INVALID_ARG; }
  const size_t smem = sizeof(Smem) + 128;
  dim3 grid((unsigned)((a.w + kCoreX - 1) / kCoreX), (unsigned)((a.h + kCoreY - 1) / kCoreY), (unsigned)a.B);
  // TMA needs 16-byte aligned rows; tensors smaller than one box take the LDG-staged variant of the same kernel
  const bool tma_ok = a.tile_mode == 0 && (a.w % 4) == 0 && (reinterpret_cast<uintptr_t>(a.logits) % 16) == 0 && a.w >= kBoxX &&
                      a.h >= kBoxY && (long long)a.B * a.D >= kChunkD;
  int rc;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  if (tma_ok) {
    if ((rc = encode_tile_map(a, &tmap))) return rc;
    if ((rc = check_cuda(cudaFuncSetAttribute(depth_head_up_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "cudaFuncSetAttribute(depth_head_up_kernel)"))) return rc;
    depth_head_up_kernel<true><<<grid, dh::kThreads, smem, s>>>(tmap, a);
  } else {
    if ((rc = check_cuda(cudaFuncSetAttribute(depth_head_up_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "cudaFuncSetAttribute(depth_head_up_kernel)"))) return rc;
    depth_head_up_kernel<false><<<grid, dh::kThreads, smem, s>>>(tmap, a);
  }
  return check_cuda(cudaGetLastError(), "depth_head_up_kernel");
}

}  // namespace fs
