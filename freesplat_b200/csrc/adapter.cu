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

int launch_gaussian_head(const FsAdapterArgs& a, cudaStream_t s) {
  if (a.N <= 0) return FS_OK;
  gaussian_head_kernel<<<(a.N + 255) / 256, 256, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "gaussian_head_kernel");
}

}  // namespace fs
