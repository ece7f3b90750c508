// raster_math.cuh -- per-Gaussian projection math shared by the preprocess and
// preprocess-backward kernels (upstream
// cuda_rasterizer/forward.cu::preprocessCUDA, computeCov2D, computeColorFromSH).
//
// Canonical floating-point order: every fused multiply-add is an explicit fmaf();
// the translation unit that includes this for the *forward* is compiled with
// -fmad=false so nothing else is contracted.  Division and sqrt are IEEE
// (nvcc defaults -prec-div=true -prec-sqrt=true, no -use_fast_math), so all
// integer decisions (radius, tile rect, depth key) are reproducible bit for bit.
//
// FS_HD lets tests/host_harness compile the very same functions for the host.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define FS_HD __host__ __device__ __forceinline__
#else
#define FS_HD static inline
#endif

namespace fsm {

constexpr float kShC0 = 0.28209479177387814f;
constexpr float kShC1 = 0.4886025119029199f;
constexpr float kShC2_0 = 1.0925484305920792f, kShC2_1 = -1.0925484305920792f,
                kShC2_2 = 0.31539156525252005f, kShC2_3 = -1.0925484305920792f,
                kShC2_4 = 0.5462742152960396f;
constexpr float kShC3_0 = -0.5900435899266435f, kShC3_1 = 2.890611442640554f,
                kShC3_2 = -0.4570457994644658f, kShC3_3 = 0.3731763325901154f,
                kShC3_4 = -0.4570457994644658f, kShC3_5 = 1.445305721320277f,
                kShC3_6 = -0.5900435899266435f;

struct Vec3 { float x, y, z; };

FS_HD float dot4row(const float* m, int r, float x, float y, float z) {
  // ((m[r]*x + m[4+r]*y) + m[8+r]*z) + m[12+r]
  return fmaf(m[8 + r], z, fmaf(m[4 + r], y, m[r] * x)) + m[12 + r];
}

FS_HD float ndc2pix(float v, int S) {
  return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5);
}

FS_HD int imin_(int a, int b) { return a < b ? a : b; }
FS_HD int imax_(int a, int b) { return a > b ? a : b; }

FS_HD void get_rect(float px, float py, int r, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
  const float fr = (float)r;
  *x0 = imin_(gx, imax_(0, (int)((px - fr) / 16.0f)));
  *y0 = imin_(gy, imax_(0, (int)((py - fr) / 16.0f)));
  *x1 = imin_(gx, imax_(0, (int)(((px + fr) + 15.0f) / 16.0f)));
  *y1 = imin_(gy, imax_(0, (int)(((py + fr) + 15.0f) / 16.0f)));
}

// Sigma = R(q) diag(mod*s)^2 R(q)^T, q = (r,x,y,z) not normalised (upstream computeCov3D).
FS_HD void cov3d_from_scale_rot(const float* s, float mod, const float* q, float* c6) {
  const float r = q[0], x = q[1], y = q[2], z = q[3];
  const float R[9] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                      2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                      2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
  const float sx = mod * s[0], sy = mod * s[1], sz = mod * s[2];
  float M[9];
  for (int i = 0; i < 3; i++) { M[3 * i] = R[3 * i] * sx; M[3 * i + 1] = R[3 * i + 1] * sy; M[3 * i + 2] = R[3 * i + 2] * sz; }
  int k = 0;
  for (int i = 0; i < 3; i++)
    for (int j = i; j < 3; j++)
      c6[k++] = fmaf(M[3 * i + 2], M[3 * j + 2], fmaf(M[3 * i + 1], M[3 * j + 1], M[3 * i] * M[3 * j]));
}

struct Cov2D {
  float a, b, c;        // cov2D (+0.3 dilation on a and c)
  float Ta[3], Tb[3];   // rows of J*R
  float tx, ty, tz;     // clamped camera-space point
  int clampx, clampy;   // 1 if the 1.3*tanfov clamp was active
};

FS_HD Cov2D cov2d(float pvx, float pvy, float pvz, float fx, float fy, float tanx, float tany,
                  const float* c6, const float* view) {
  Cov2D o;
  const float limx = 1.3f * tanx, limy = 1.3f * tany;
  const float txtz = pvx / pvz, tytz = pvy / pvz;
  o.clampx = (txtz < -limx || txtz > limx);
  o.clampy = (tytz < -limy || tytz > limy);
  const float tx = fminf(limx, fmaxf(-limx, txtz)) * pvz;
  const float ty = fminf(limy, fmaxf(-limy, tytz)) * pvz;
  const float tz = pvz;
  const float J00 = fx / tz, J11 = fy / tz;
  const float J02 = -(fx * tx) / (tz * tz), J12 = -(fy * ty) / (tz * tz);
  for (int i = 0; i < 3; i++) {
    const float R0 = view[i * 4 + 0], R1 = view[i * 4 + 1], R2 = view[i * 4 + 2];
    o.Ta[i] = fmaf(R2, J02, R0 * J00);
    o.Tb[i] = fmaf(R2, J12, R1 * J11);
  }
  const float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
  float sa[3], sb[3];
  for (int k = 0; k < 3; k++) {
    sa[k] = fmaf(S[3 * k + 2], o.Ta[2], fmaf(S[3 * k + 1], o.Ta[1], S[3 * k] * o.Ta[0]));
    sb[k] = fmaf(S[3 * k + 2], o.Tb[2], fmaf(S[3 * k + 1], o.Tb[1], S[3 * k] * o.Tb[0]));
  }
  o.a = fmaf(o.Ta[2], sa[2], fmaf(o.Ta[1], sa[1], o.Ta[0] * sa[0])) + 0.3f;
  o.b = fmaf(o.Ta[2], sb[2], fmaf(o.Ta[1], sb[1], o.Ta[0] * sb[0]));
  o.c = fmaf(o.Tb[2], sb[2], fmaf(o.Tb[1], sb[1], o.Tb[0] * sb[0])) + 0.3f;
  o.tx = tx; o.ty = ty; o.tz = tz;
  return o;
}

// SH -> RGB for one Gaussian.  sh: [M][3].  Returns clamp bitmask (bit c = channel c clamped).
FS_HD int sh_to_rgb(int deg, const float* mean, const float* campos, const float* sh, float* rgb) {
  const float dx = mean[0] - campos[0], dy = mean[1] - campos[1], dz = mean[2] - campos[2];
  const float len = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
  const float x = dx / len, y = dy / len, z = dz / len;
  int mask = 0;
  for (int c = 0; c < 3; c++) {
    float res = kShC0 * sh[c];
    if (deg > 0) {
      res = fmaf(-(kShC1 * y), sh[3 + c], res);
      res = fmaf(kShC1 * z, sh[6 + c], res);
      res = fmaf(-(kShC1 * x), sh[9 + c], res);
      if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        res = fmaf(kShC2_0 * xy, sh[12 + c], res);
        res = fmaf(kShC2_1 * yz, sh[15 + c], res);
        res = fmaf(kShC2_2 * ((2.0f * zz - xx) - yy), sh[18 + c], res);
        res = fmaf(kShC2_3 * xz, sh[21 + c], res);
        res = fmaf(kShC2_4 * (xx - yy), sh[24 + c], res);
        if (deg > 2) {
          res = fmaf(kShC3_0 * y * (3.0f * xx - yy), sh[27 + c], res);
          res = fmaf(kShC3_1 * xy * z, sh[30 + c], res);
          res = fmaf(kShC3_2 * y * ((4.0f * zz - xx) - yy), sh[33 + c], res);
          res = fmaf(kShC3_3 * z * ((2.0f * zz - 3.0f * xx) - 3.0f * yy), sh[36 + c], res);
          res = fmaf(kShC3_4 * x * ((4.0f * zz - xx) - yy), sh[39 + c], res);
          res = fmaf(kShC3_5 * z * (xx - yy), sh[42 + c], res);
          res = fmaf(kShC3_6 * x * (xx - 3.0f * yy), sh[45 + c], res);
        }
      }
    }
    res += 0.5f;
    if (res < 0.f) mask |= (1 << c);
    rgb[c] = fmaxf(res, 0.f);
  }
  return mask;
}

// Result of projecting one Gaussian into one view (A.1 steps 1-10).
struct Projected {
  int radius;            // 0 => culled
  int x0, y0, x1, y1;    // tile rect [x0,x1) x [y0,y1)
  float px, py;          // pixel-space mean
  float con_x, con_y, con_z;
  float depth;           // p_view.z
  float c6[6];           // covariance actually used (after scene_scale^2)
};

// mean: world-space mean ALREADY multiplied by scene_scale; c6: covariance already scaled.
FS_HD Projected project_gaussian(const float* mean, const float* c6, const float* view, const float* proj,
                                 float tanx, float tany, int H, int W) {
  Projected o;
  o.radius = 0; o.x0 = o.y0 = o.x1 = o.y1 = 0; o.px = o.py = 0.f;
  o.con_x = o.con_y = o.con_z = 0.f; o.depth = 0.f;
  for (int k = 0; k < 6; k++) o.c6[k] = c6[k];
  const int gx = (W + 15) / 16, gy = (H + 15) / 16;
  const float fx = (float)W / (2.0f * tanx), fy = (float)H / (2.0f * tany);
  const float pvx = dot4row(view, 0, mean[0], mean[1], mean[2]);
  const float pvy = dot4row(view, 1, mean[0], mean[1], mean[2]);
  const float pvz = dot4row(view, 2, mean[0], mean[1], mean[2]);
  if (pvz <= 0.2f) return o;
  const float phx = dot4row(proj, 0, mean[0], mean[1], mean[2]);
  const float phy = dot4row(proj, 1, mean[0], mean[1], mean[2]);
  const float phw = dot4row(proj, 3, mean[0], mean[1], mean[2]);
  const float pw = 1.0f / (phw + 0.0000001f);
  const float ppx = phx * pw, ppy = phy * pw;
  const Cov2D cv = cov2d(pvx, pvy, pvz, fx, fy, tanx, tany, c6, view);
  const float a = cv.a, b = cv.b, c = cv.c;
  const float det = fmaf(-b, b, a * c);
  if (det == 0.0f) return o;
  const float det_inv = 1.f / det;
  const float mid = 0.5f * (a + c);
  const float sq = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
  const float l1 = mid + sq, l2 = mid - sq;
  const int r = (int)ceilf(3.f * sqrtf(fmaxf(l1, l2)));
  const float px = ndc2pix(ppx, W), py = ndc2pix(ppy, H);
  int x0, y0, x1, y1;
  get_rect(px, py, r, gx, gy, &x0, &y0, &x1, &y1);
  if ((x1 - x0) * (y1 - y0) == 0) return o;
  o.radius = r; o.x0 = x0; o.y0 = y0; o.x1 = x1; o.y1 = y1;
  o.px = px; o.py = py;
  o.con_x = c * det_inv; o.con_y = -b * det_inv; o.con_z = a * det_inv;
  o.depth = pvz;
  return o;
}

// Conservative half-extents of {pixels where o*exp(power) >= 1/255 can hold}.
// NOT part of the reference arithmetic: it only prunes work whose result the
// per-pixel alpha test would discard anyway (see DESIGN.md "exact culling").
FS_HD void alpha_extent(float con_x, float con_y, float con_z, float opacity, float* hx, float* hy) {
  const float BIG = 3.0e38f;
  const float tau = logf(255.0f * opacity) + 0.01f;       // alpha>=1/255  <=>  -power <= ln(255 o)
  const float detq = con_x * con_z - con_y * con_y;
  if (!(opacity > 0.f) || !(tau == tau)) { *hx = -1.f; *hy = -1.f; return; }
  if (tau < 0.f) { *hx = -1.f; *hy = -1.f; return; }        // can never reach 1/255
  if (!(detq > 0.f) || !(con_x > 0.f) || !(con_z > 0.f)) { *hx = BIG; *hy = BIG; return; }
  const float ex = sqrtf(2.0f * tau * con_z / detq), ey = sqrtf(2.0f * tau * con_x / detq);
  *hx = (ex == ex && ex < 1.0e30f) ? ex * 1.001f + 0.01f : BIG;
  *hy = (ey == ey && ey < 1.0e30f) ? ey * 1.001f + 0.01f : BIG;
}

}  // namespace fsm
