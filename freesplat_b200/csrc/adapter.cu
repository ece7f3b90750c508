// adapter.cu -- Gaussian head: the reference turns the fused per-Gaussian latents into
// rasterizer inputs with ~25 element-wise torch launches
// (src/model/encoder/common/gaussian_adapter.py:151-172 + common/gaussians.py:8-44, called at
//  src/model/encoder/encoder_freesplat.py:376-386: sigmoid scale mapping x depth x pixel multiplier, quaternion
//  normalisation, SH mask, R S S^T R^T, rotation into the world by the (averaged) camera-to-world matrix).
// One HBM-bound kernel here: read 34 + 1 + 1 + 16 + 3 floats, write 3 + 9 + 27 + 1 + 3 + 4 per Gaussian, written in the
// layouts the raster kernels read in place ([N,3,3] covariances, [N,3,d_sh] harmonics).
#include "common.cuh"

namespace fs {

__global__ void __launch_bounds__(256) gaussian_head_kernel(FsAdapterArgs a) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= a.N) return;
  const int dsh = (a.sh_degree + 1) * (a.sh_degree + 1);
  const int din = 7 + 3 * dsh;
  const float* raw = a.raw + (size_t)i * din;
  const float depth = a.depths[i];
  // multiplier = 0.1 * sum(inv(K[:2,:2]) @ (1/w, 1/h))   (get_scale_multiplier, gaussian_adapter.py:203-214)
  const float k00 = a.K[0], k01 = a.K[1], k10 = a.K[3], k11 = a.K[4];
  const float idet = 1.0f / (k00 * k11 - k01 * k10);
  const float pw = 1.0f / (float)a.W, ph = 1.0f / (float)a.H;
  const float mult = 0.1f * ((k11 * pw - k01 * ph) * idet) + 0.1f * ((-k10 * pw + k00 * ph) * idet);
  float s[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float sg = 1.0f / (1.0f + expf(-raw[k]));
    s[k] = (a.scale_min + (a.scale_max - a.scale_min) * sg) * depth * mult;
    a.scales[3 * (size_t)i + k] = s[k];
  }
  float q[4] = {raw[3], raw[4], raw[5], raw[6]};
  const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + a.eps;
#pragma unroll
  for (int k = 0; k < 4; k++) { q[k] = q[k] / nrm; a.rotations[4 * (size_t)i + k] = q[k]; }
  // quaternion_to_matrix (xyzw order)
  const float qi = q[0], qj = q[1], qk = q[2], qr = q[3];
  const float two_s = 2.0f / (qi * qi + qj * qj + qk * qk + qr * qr + 1e-8f);
  const float R[9] = {1 - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
                      two_s * (qi * qj + qk * qr), 1 - two_s * (qi * qi + qk * qk), two_s * (qj * qk - qi * qr),
                      two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi * qi + qj * qj)};
  // M = C R S  (C = camera-to-world rotation block);  covariance = M M^T
  const float* E = a.ext + 16 * (size_t)i;
  float M[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) M[3 * r + c] = (E[4 * r] * R[c] + E[4 * r + 1] * R[3 + c] + E[4 * r + 2] * R[6 + c]) * s[c];
  float* cov = a.covariances + 9 * (size_t)i;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) cov[3 * r + c] = M[3 * r] * M[3 * c] + M[3 * r + 1] * M[3 * c + 1] + M[3 * r + 2] * M[3 * c + 2];
  // harmonics [3][d_sh] * mask(degree)
  float* sh = a.harmonics + (size_t)i * 3 * dsh;
  for (int c = 0; c < 3; c++)
    for (int d = 0; d < dsh; d++) {
      const int deg = d == 0 ? 0 : (d < 4 ? 1 : (d < 9 ? 2 : 3));
      const float mask = deg == 0 ? 1.0f : 0.1f * (deg == 1 ? 0.25f : deg == 2 ? 0.0625f : 0.015625f);
      sh[c * dsh + d] = raw[7 + c * dsh + d] * mask;
    }
  a.means[3 * (size_t)i] = a.coords[3 * (size_t)i]; a.means[3 * (size_t)i + 1] = a.coords[3 * (size_t)i + 1];
  a.means[3 * (size_t)i + 2] = a.coords[3 * (size_t)i + 2];
  a.opacities_out[i] = a.opacities[i];
}

// Backward of gaussian_head_kernel (training path; the reference gets it from autograd through the ~25 torch ops of
// gaussian_adapter.py:151-172 / gaussians.py:8-44).  One thread per Gaussian re-derives the forward intermediates and
// applies the chain rule by hand:
//   cov = M M^T, M = C R S      ->  dM = (G + G^T) M ;  ds_c = sum_r dM[r][c] (CR)[r][c] ;  d(CR) = dM S ;
//                                   dR = C^T d(CR) ;  dC = d(CR) R^T
//   R(q) = I + two_s B(q), two_s = 2 / (|q|^2 + 1e-8)   (pytorch3d quaternion_to_matrix, xyzw)
//   q = raw_q / (|raw_q| + eps) ;  s = (smin + (smax - smin) sigmoid(raw_s)) * depth * mult
// Gradients w.r.t. raw [N,7+3 d_sh], depths, opacities, coords and the rotation block of the per-Gaussian c2w matrix
// (PTF's density-weighted average: the reference back-propagates through it into the densities).
__global__ void __launch_bounds__(256) gaussian_head_bwd_kernel(FsAdapterBwdArgs a) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= a.N) return;
  const int dsh = (a.sh_degree + 1) * (a.sh_degree + 1);
  const int din = 7 + 3 * dsh;
  const float* raw = a.raw + (size_t)i * din;
  float* d_raw = a.d_raw + (size_t)i * din;
  const float depth = a.depths[i];
  const float k00 = a.K[0], k01 = a.K[1], k10 = a.K[3], k11 = a.K[4];
  const float idet = 1.0f / (k00 * k11 - k01 * k10);
  const float pw = 1.0f / (float)a.W, ph = 1.0f / (float)a.H;
  const float mult = 0.1f * ((k11 * pw - k01 * ph) * idet) + 0.1f * ((-k10 * pw + k00 * ph) * idet);
  float s[3], sg[3], base[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    sg[k] = 1.0f / (1.0f + expf(-raw[k]));
    base[k] = a.scale_min + (a.scale_max - a.scale_min) * sg[k];
    s[k] = base[k] * depth * mult;
  }
  const float rq[4] = {raw[3], raw[4], raw[5], raw[6]};
  const float nq = sqrtf(rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3]);
  const float den = nq + a.eps;
  const float q[4] = {rq[0] / den, rq[1] / den, rq[2] / den, rq[3] / den};
  const float qi = q[0], qj = q[1], qk = q[2], qr = q[3];
  const float n2 = qi * qi + qj * qj + qk * qk + qr * qr + 1e-8f;
  const float two_s = 2.0f / n2;
  // B(q): R = I + two_s * B
  const float B[9] = {-(qj * qj + qk * qk), qi * qj - qk * qr, qi * qk + qj * qr,
                      qi * qj + qk * qr, -(qi * qi + qk * qk), qj * qk - qi * qr,
                      qi * qk - qj * qr, qj * qk + qi * qr, -(qi * qi + qj * qj)};
  float R[9];
#pragma unroll
  for (int e = 0; e < 9; e++) R[e] = ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f) + two_s * B[e];
  const float* E = a.ext + 16 * (size_t)i;
  float CR[9], M[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      CR[3 * r + c] = E[4 * r] * R[c] + E[4 * r + 1] * R[3 + c] + E[4 * r + 2] * R[6 + c];
      M[3 * r + c] = CR[3 * r + c] * s[c];
    }
  // upstream gradient of the covariance (any of the g_* may be NULL)
  float G[9];
#pragma unroll
  for (int e = 0; e < 9; e++) G[e] = a.g_cov ? a.g_cov[9 * (size_t)i + e] : 0.f;
  float dM[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; k++) acc += (G[3 * r + k] + G[3 * k + r]) * M[3 * k + c];
      dM[3 * r + c] = acc;
    }
  float ds[3], dCR[9];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    ds[c] = dM[c] * CR[c] + dM[3 + c] * CR[3 + c] + dM[6 + c] * CR[6 + c];
    if (a.g_scales) ds[c] += a.g_scales[3 * (size_t)i + c];
#pragma unroll
    for (int r = 0; r < 3; r++) dCR[3 * r + c] = dM[3 * r + c] * s[c];
  }
  // dR = C^T dCR ; dC = dCR R^T
  float dR[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) dR[3 * r + c] = E[r] * dCR[c] + E[4 + r] * dCR[3 + c] + E[8 + r] * dCR[6 + c];
  if (a.d_ext) {
    float* dE = a.d_ext + 16 * (size_t)i;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int c = 0; c < 4; c++) {
        float v = 0.f;
        if (r < 3 && c < 3) v = dCR[3 * r] * R[3 * c] + dCR[3 * r + 1] * R[3 * c + 1] + dCR[3 * r + 2] * R[3 * c + 2];
        dE[4 * r + c] = v;
      }
  }
  // dR -> dq (normalised quaternion): R = I + two_s B
  float dB[9], dtwo = 0.f;
#pragma unroll
  for (int e = 0; e < 9; e++) { dB[e] = dR[e] * two_s; dtwo += dR[e] * B[e]; }
  float dq[4];
  dq[0] = (dB[1] + dB[3]) * qj + (dB[2] + dB[6]) * qk + (dB[7] - dB[5]) * qr - 2.f * qi * (dB[4] + dB[8]);
  dq[1] = (dB[1] + dB[3]) * qi + (dB[5] + dB[7]) * qk + (dB[2] - dB[6]) * qr - 2.f * qj * (dB[0] + dB[8]);
  dq[2] = (dB[2] + dB[6]) * qi + (dB[5] + dB[7]) * qj + (dB[3] - dB[1]) * qr - 2.f * qk * (dB[0] + dB[4]);
  dq[3] = -dB[1] * qk + dB[2] * qj + dB[3] * qk - dB[5] * qi - dB[6] * qj + dB[7] * qi;
  const float dn2 = dtwo * (-2.0f / (n2 * n2));
#pragma unroll
  for (int k = 0; k < 4; k++) {
    dq[k] += dn2 * 2.f * q[k];
    if (a.g_rotations) dq[k] += a.g_rotations[4 * (size_t)i + k];
  }
  // q = rq / (|rq| + eps):  d rq_j = dq_j / den - (sum_i dq_i rq_i) rq_j / (|rq| den^2)
  const float dot = dq[0] * rq[0] + dq[1] * rq[1] + dq[2] * rq[2] + dq[3] * rq[3];
  const float f = nq > 0.f ? dot / (nq * den * den) : 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++) d_raw[3 + k] = dq[k] / den - f * rq[k];
  float d_depth = 0.f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    d_raw[k] = ds[k] * (a.scale_max - a.scale_min) * sg[k] * (1.0f - sg[k]) * depth * mult;
    d_depth += ds[k] * base[k] * mult;
  }
  a.d_depths[i] = d_depth;
  for (int c = 0; c < 3; c++)
    for (int d = 0; d < dsh; d++) {
      const int deg = d == 0 ? 0 : (d < 4 ? 1 : (d < 9 ? 2 : 3));
      const float mask = deg == 0 ? 1.0f : 0.1f * (deg == 1 ? 0.25f : deg == 2 ? 0.0625f : 0.015625f);
      d_raw[7 + c * dsh + d] = a.g_harmonics ? a.g_harmonics[(size_t)i * 3 * dsh + c * dsh + d] * mask : 0.f;
    }
#pragma unroll
  for (int k = 0; k < 3; k++) a.d_coords[3 * (size_t)i + k] = a.g_means ? a.g_means[3 * (size_t)i + k] : 0.f;
  a.d_opacities[i] = a.g_opacities ? a.g_opacities[i] : 0.f;
}

int launch_gaussian_head_bwd(const FsAdapterBwdArgs& a, cudaStream_t s) {
  if (a.N <= 0) return FS_OK;
  gaussian_head_bwd_kernel<<<(a.N + 255) / 256, 256, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "gaussian_head_bwd_kernel");
}

// Depth back-projection of the context views: GaussianAdapter.forward(fusion=True)
// (gaussian_adapter.py:175-189) -> Create_from_depth_map.project (:48-68): per pixel (i, j)
//   cam = ((j - cx) / fx * z, (i - cy) / fy * z, z, 1),  world = c2w . cam
// with the pixel-space intrinsics of view 0 (rows scaled by w and h in fp32, :179-181).  The reference loops over
// views in Python with ~15 torch launches each; one launch here.  The world coordinates feed the index decisions of
// PTF, so the arithmetic is the canonical order of oracle/adapter.py::backproject (round-to-nearest intrinsics, explicit
// fma chain = torch's CPU sgemm order for [4,4] @ [4,N]): bit-identical to the reference on the golden cases.
__global__ void __launch_bounds__(256) backproject_kernel(FsBackprojectArgs a) {
  const int HW = a.H * a.W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const int v = blockIdx.y;
  const int i = p / a.W, j = p - i * a.W;
  const float fx = __fmul_rn(a.K[0], (float)a.W), cx = __fmul_rn(a.K[2], (float)a.W);
  const float fy = __fmul_rn(a.K[4], (float)a.H), cy = __fmul_rn(a.K[5], (float)a.H);
  const float z = a.depth[(size_t)v * HW + p];
  const float x = __fmul_rn(__fdiv_rn(__fsub_rn((float)j, cx), fx), z);
  const float y = __fmul_rn(__fdiv_rn(__fsub_rn((float)i, cy), fy), z);
  const float* E = a.c2w + 16 * (size_t)v;
  float* o = a.means + 3 * ((size_t)v * HW + p);
#pragma unroll
  for (int r = 0; r < 3; r++)
    o[r] = __fmaf_rn(E[4 * r + 3], 1.0f, __fmaf_rn(E[4 * r + 2], z, __fmaf_rn(E[4 * r + 1], y, __fmul_rn(E[4 * r], x))));
}

// backward of backproject_kernel w.r.t. the depth map: world = c2w . (x z, y z, z, 1) with x = (j - cx)/fx, y = (i - cy)/fy
//   d depth = sum_r g[r] (E[r][0] x + E[r][1] y + E[r][2])
__global__ void __launch_bounds__(256) backproject_bwd_kernel(FsBackprojectArgs a, const float* __restrict__ g_means, float* __restrict__ d_depth) {
  const int HW = a.H * a.W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const int v = blockIdx.y;
  const int i = p / a.W, j = p - i * a.W;
  const float fx = __fmul_rn(a.K[0], (float)a.W), cx = __fmul_rn(a.K[2], (float)a.W);
  const float fy = __fmul_rn(a.K[4], (float)a.H), cy = __fmul_rn(a.K[5], (float)a.H);
  const float x = ((float)j - cx) / fx, y = ((float)i - cy) / fy;
  const float* E = a.c2w + 16 * (size_t)v;
  const float* g = g_means + 3 * ((size_t)v * HW + p);
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < 3; r++) acc += g[r] * (E[4 * r] * x + E[4 * r + 1] * y + E[4 * r + 2]);
  d_depth[(size_t)v * HW + p] = acc;
}

int launch_backproject_bwd(const FsBackprojectArgs& a, const float* g_means, float* d_depth, cudaStream_t s) {
  if (a.V <= 0) return FS_OK;
  dim3 grid((unsigned)((a.H * a.W + 255) / 256), (unsigned)a.V);
  backproject_bwd_kernel<<<grid, 256, 0, s>>>(a, g_means, d_depth);
  return check_cuda(cudaGetLastError(), "backproject_bwd_kernel");
}

int launch_backproject(const FsBackprojectArgs& a, cudaStream_t s) {
  if (a.V <= 0) return FS_OK;
  dim3 grid((unsigned)((a.H * a.W + 255) / 256), (unsigned)a.V);
  backproject_kernel<<<grid, 256, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "backproject_kernel");
}

// Vertex table of the .ply export (src/model/ply_export.py:26-92): per Gaussian
//   xyz = R ((mean - median) / scale_factor), normals 0, DC band of the harmonics, opacity, log(scale / scale_factor),
//   rotation = quaternion of  R . matrix(q)  in (w, x, y, z) order -- scipy's from_quat / from_matrix rules in fp64, as the
// reference evaluates them (oracle/ply.py).  Rows leave through shared memory as contiguous float runs (17 floats per row).
__global__ void __launch_bounds__(256) ply_vertices_kernel(FsPlyArgs a) {
  __shared__ float rows[256 * 17];
  const int base = blockIdx.x * 256, n = base + threadIdx.x;
  if (n < a.N) {
    float* o = rows + threadIdx.x * 17;
    const float sf = a.scale_factor;
    float m[3];
#pragma unroll
    for (int k = 0; k < 3; k++) m[k] = (a.means[3 * (size_t)n + k] - a.shift[k]) / sf;
#pragma unroll
    for (int r = 0; r < 3; r++) o[r] = a.R[3 * r] * m[0] + a.R[3 * r + 1] * m[1] + a.R[3 * r + 2] * m[2];
    o[3] = 0.f; o[4] = 0.f; o[5] = 0.f;
    const float* sh = a.harmonics + (size_t)n * 3 * a.d_sh;
    o[6] = sh[0]; o[7] = sh[a.d_sh]; o[8] = sh[2 * a.d_sh];
    o[9] = a.opacities[n];
#pragma unroll
    for (int k = 0; k < 3; k++) o[10 + k] = logf(a.scales[3 * (size_t)n + k] / sf);
    // quaternion (x, y, z, w) -> matrix -> rotate -> quaternion, fp64
    double q[4] = {(double)a.rotations[4 * (size_t)n], (double)a.rotations[4 * (size_t)n + 1], (double)a.rotations[4 * (size_t)n + 2],
                   (double)a.rotations[4 * (size_t)n + 3]};
    const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double x = q[0] / qn, y = q[1] / qn, z = q[2] / qn, w = q[3] / qn;
    const double Mq[9] = {x * x - y * y - z * z + w * w, 2 * (x * y - z * w), 2 * (x * z + y * w),
                          2 * (x * y + z * w), -x * x + y * y - z * z + w * w, 2 * (y * z - x * w),
                          2 * (x * z - y * w), 2 * (y * z + x * w), -x * x - y * y + z * z + w * w};
    double A[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        A[3 * r + c] = (double)a.R[3 * r] * Mq[c] + (double)a.R[3 * r + 1] * Mq[3 + c] + (double)a.R[3 * r + 2] * Mq[6 + c];
    const double tr = A[0] + A[4] + A[8];
    int c = 0;
    double best = A[0];
    if (A[4] > best) { best = A[4]; c = 1; }
    if (A[8] > best) { best = A[8]; c = 2; }
    if (tr > best) c = 3;
    double qo[4];
    if (c != 3) {
      const int i = c, j = (c + 1) % 3, k = (c + 2) % 3;
      qo[i] = 1 - tr + 2 * A[3 * i + i]; qo[j] = A[3 * j + i] + A[3 * i + j]; qo[k] = A[3 * k + i] + A[3 * i + k];
      qo[3] = A[3 * k + j] - A[3 * j + k];
    } else {
      qo[0] = A[7] - A[5]; qo[1] = A[2] - A[6]; qo[2] = A[3] - A[1]; qo[3] = 1 + tr;
    }
    const double on = sqrt(qo[0] * qo[0] + qo[1] * qo[1] + qo[2] * qo[2] + qo[3] * qo[3]);
    o[13] = (float)(qo[3] / on); o[14] = (float)(qo[0] / on); o[15] = (float)(qo[1] / on); o[16] = (float)(qo[2] / on);
  }
  __syncthreads();
  const int cnt = min(256, a.N - base) * 17;
  float* dst = a.table + (size_t)base * 17;
  for (int k = threadIdx.x; k < cnt; k += 256) dst[k] = rows[k];
}

// Evaluation image dump (src/misc/image_io.py:36-53 `prep_image`, called by `save_image` from
// model_wrapper.py:382-416): float images in [0,1] -> uint8 HWC, batch concatenated along the width ("b c h w -> c h (b w)"),
// single-channel images repeated to 3 channels, value = uint8(clip(x, 0, 1) * 255) (truncation, as torch's .type(torch.uint8)).
// One pass on the device: the D2H copy that follows moves 1 byte per sample instead of 4.
__global__ void __launch_bounds__(256) image_u8_kernel(int B, int C, int H, int W, const float* __restrict__ img, uint8_t* __restrict__ out) {
  const int Co = C == 1 ? 3 : C;
  const size_t total = (size_t)H * B * W;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;       // output pixel (y, b*W + x)
  if (i >= total) return;
  const int y = (int)(i / ((size_t)B * W));
  const int bx = (int)(i - (size_t)y * B * W);
  const int b = bx / W, x = bx - b * W;
  const float* src = img + ((size_t)b * C * H + y) * W + x;
  uint8_t* dst = out + i * Co;
  for (int c = 0; c < Co; c++) {
    const float v = src[(size_t)(C == 1 ? 0 : c) * H * W];
    const float q = fminf(fmaxf(v, 0.f), 1.f) * 255.f;
    dst[c] = (uint8_t)q;
  }
}

int launch_image_u8(int B, int C, int H, int W, const float* img, uint8_t* out, cudaStream_t s) {
  const size_t total = (size_t)H * B * W;
  if (total == 0) return FS_OK;
  image_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(B, C, H, W, img, out);
  return check_cuda(cudaGetLastError(), "image_u8_kernel");
}

int launch_ply_vertices(const FsPlyArgs& a, cudaStream_t s) {
  if (a.N <= 0) return FS_OK;
  ply_vertices_kernel<<<(a.N + 255) / 256, 256, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "ply_vertices_kernel");
}

int launch_gaussian_head(const FsAdapterArgs& a, cudaStream_t s) {
  if (a.N <= 0) return FS_OK;
  gaussian_head_kernel<<<(a.N + 255) / 256, 256, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "gaussian_head_kernel");
}

}  // namespace fs
