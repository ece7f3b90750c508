/*
 * oracle/raster_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP, fp32) of the tile-based differentiable
 * Gaussian rasterizer that FreeSplat calls through
 *   /root/reference/src/model/decoder/cuda_splatting.py:100-127
 * (module `diff_gaussian_rasterization_depth`, requirements.txt:17:
 *  git+https://github.com/JonathonLuiten/diff-gaussian-rasterization-w-depth,
 *  un-pinned, NOT vendored in /root/reference and not installed in this image).
 *
 * PARITY UNPINNED: the arithmetic lives in that absent third-party CUDA
 * extension; the reference holds no test / golden vector for it.
 * This file restates the published algorithm (upstream
 * files cuda_rasterizer/{forward,backward,rasterizer_impl}.cu, auxiliary.h) and
 * is itself cross-checked in tests/ against an independent dense PyTorch
 * renderer + autograd (tests/dense_torch_raster.py).  Facts pinned by the
 * FreeSplat call site: 4-tuple return, depth rank-2 [H,W] un-normalised
 * (cuda_splatting.py:120-128, decoder_splatting_cuda.py:60-62), SH layout
 * [P,M,3] (:75,:123), cov6 order xx,xy,xz,yy,yz,zz (:116,:126), matrices passed
 * transposed (:85-87).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Canonical floating-point op order: every a*b+c that is meant to be fused is
 * written as fmaf(); compile with -ffp-contract=off so nothing else is.  The
 * CUDA kernels (freesplat_b200/csrc) use the same order with -fmad=false in the
 * per-Gaussian stages so that all integer outputs (radii, tiles_touched,
 * offsets, keys, point_list, ranges) are bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BLOCK_X 16
#define BLOCK_Y 16

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f,
                               0.31539156525252005f, -1.0925484305920792f,
                               0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,
                               -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* ---------- small helpers (Appendix A.0) ---------- */
static inline void xform4x3(const float* m, float x, float y, float z, float* o) {
  /* ((m0*x + m4*y) + m8*z) + m12, products fused into the running sum */
  o[0] = fmaf(m[8], z, fmaf(m[4], y, m[0] * x)) + m[12];
  o[1] = fmaf(m[9], z, fmaf(m[5], y, m[1] * x)) + m[13];
  o[2] = fmaf(m[10], z, fmaf(m[6], y, m[2] * x)) + m[14];
}
static inline void xform4x4(const float* m, float x, float y, float z, float* o) {
  o[0] = fmaf(m[8], z, fmaf(m[4], y, m[0] * x)) + m[12];
  o[1] = fmaf(m[9], z, fmaf(m[5], y, m[1] * x)) + m[13];
  o[2] = fmaf(m[10], z, fmaf(m[6], y, m[2] * x)) + m[14];
  o[3] = fmaf(m[11], z, fmaf(m[7], y, m[3] * x)) + m[15];
}
static inline float ndc2pix(float v, int S) {
  return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5);
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

static inline void get_rect(float px, float py, int r, int gx, int gy, int* rmin, int* rmax) {
  float fr = (float)r;
  rmin[0] = imin(gx, imax(0, (int)((px - fr) / (float)BLOCK_X)));
  rmin[1] = imin(gy, imax(0, (int)((py - fr) / (float)BLOCK_Y)));
  rmax[0] = imin(gx, imax(0, (int)(((px + fr) + (float)(BLOCK_X - 1)) / (float)BLOCK_X)));
  rmax[1] = imin(gy, imax(0, (int)(((py + fr) + (float)(BLOCK_Y - 1)) / (float)BLOCK_Y)));
}

/* quaternion (r,x,y,z), NOT normalised (upstream) -> Sigma = R S^2 R^T, 6 unique */
static void cov3d_from_scale_rot(const float* s, float mod, const float* q, float* cov6) {
  float r = q[0], x = q[1], y = q[2], z = q[3];
  /* R = standard rotation matrix of q (row-major) */
  float R[9] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
  float sx = mod * s[0], sy = mod * s[1], sz = mod * s[2];
  /* M = R * S  (columns scaled);  Sigma = M M^T */
  float M[9];
  for (int i = 0; i < 3; i++) { M[3 * i + 0] = R[3 * i + 0] * sx; M[3 * i + 1] = R[3 * i + 1] * sy; M[3 * i + 2] = R[3 * i + 2] * sz; }
  int k = 0;
  for (int i = 0; i < 3; i++)
    for (int j = i; j < 3; j++)
      cov6[k++] = fmaf(M[3 * i + 2], M[3 * j + 2], fmaf(M[3 * i + 1], M[3 * j + 1], M[3 * i + 0] * M[3 * j + 0]));
}

/* EWA projection, A.1 step 4.  Returns a,b,c (cov2D with +0.3 dilation) and
 * optionally the intermediate rows (Ta = J R row 0, Tb = J R row 1). */
static void cov2d(const float* t_in, float fx, float fy, float tanx, float tany,
                  const float* c6, const float* view, float* abc, float* Ta, float* Tb,
                  float* t_out, int* clampx, int* clampy) {
  float tx = t_in[0], ty = t_in[1], tz = t_in[2];
  float limx = 1.3f * tanx, limy = 1.3f * tany;
  float txtz = tx / tz, tytz = ty / tz;
  if (clampx) *clampx = (txtz < -limx || txtz > limx);
  if (clampy) *clampy = (tytz < -limy || tytz > limy);
  tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
  ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
  float J00 = fx / tz, J11 = fy / tz;
  float J02 = -(fx * tx) / (tz * tz), J12 = -(fy * ty) / (tz * tz);
  /* R[r][c] = view[c*4+r] */
  float a[3], b[3];
  for (int i = 0; i < 3; i++) {
    float R0 = view[i * 4 + 0], R1 = view[i * 4 + 1], R2 = view[i * 4 + 2];
    a[i] = fmaf(R2, J02, R0 * J00);
    b[i] = fmaf(R2, J12, R1 * J11);
  }
  float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
  float sa[3], sb[3];
  for (int k = 0; k < 3; k++) {
    sa[k] = fmaf(S[3 * k + 2], a[2], fmaf(S[3 * k + 1], a[1], S[3 * k + 0] * a[0]));
    sb[k] = fmaf(S[3 * k + 2], b[2], fmaf(S[3 * k + 1], b[1], S[3 * k + 0] * b[0]));
  }
  abc[0] = fmaf(a[2], sa[2], fmaf(a[1], sa[1], a[0] * sa[0])) + 0.3f;
  abc[1] = fmaf(a[2], sb[2], fmaf(a[1], sb[1], a[0] * sb[0]));
  abc[2] = fmaf(b[2], sb[2], fmaf(b[1], sb[1], b[0] * sb[0])) + 0.3f;
  if (Ta) { memcpy(Ta, a, 12); memcpy(Tb, b, 12); }
  if (t_out) { t_out[0] = tx; t_out[1] = ty; t_out[2] = tz; }
}

/* SH -> RGB, A.1 step 9 */
static void sh_to_rgb(int deg, int M, const float* mean, const float* campos,
                      const float* sh /*[M][3]*/, float* rgb, int* clamped) {
  float dx = mean[0] - campos[0], dy = mean[1] - campos[1], dz = mean[2] - campos[2];
  float len = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
  float x = dx / len, y = dy / len, z = dz / len;
  (void)M;
  for (int c = 0; c < 3; c++) {
    float res = SH_C0 * sh[0 * 3 + c];
    if (deg > 0) {
      res = fmaf(-(SH_C1 * y), sh[1 * 3 + c], res);
      res = fmaf(SH_C1 * z, sh[2 * 3 + c], res);
      res = fmaf(-(SH_C1 * x), sh[3 * 3 + c], res);
      if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        res = fmaf(SH_C2[0] * xy, sh[4 * 3 + c], res);
        res = fmaf(SH_C2[1] * yz, sh[5 * 3 + c], res);
        res = fmaf(SH_C2[2] * ((2.0f * zz - xx) - yy), sh[6 * 3 + c], res);
        res = fmaf(SH_C2[3] * xz, sh[7 * 3 + c], res);
        res = fmaf(SH_C2[4] * (xx - yy), sh[8 * 3 + c], res);
        if (deg > 2) {
          res = fmaf(SH_C3[0] * y * (3.0f * xx - yy), sh[9 * 3 + c], res);
          res = fmaf(SH_C3[1] * xy * z, sh[10 * 3 + c], res);
          res = fmaf(SH_C3[2] * y * ((4.0f * zz - xx) - yy), sh[11 * 3 + c], res);
          res = fmaf(SH_C3[3] * z * ((2.0f * zz - 3.0f * xx) - 3.0f * yy), sh[12 * 3 + c], res);
          res = fmaf(SH_C3[4] * x * ((4.0f * zz - xx) - yy), sh[13 * 3 + c], res);
          res = fmaf(SH_C3[5] * z * (xx - yy), sh[14 * 3 + c], res);
          res = fmaf(SH_C3[6] * x * (xx - 3.0f * yy), sh[15 * 3 + c], res);
        }
      }
    }
    res += 0.5f;
    clamped[c] = res < 0.f;
    rgb[c] = fmaxf(res, 0.f);
  }
}

/* ------------------------------------------------------------------------- */
/* Stage 1: preprocess + inclusive scan.  Returns R = number of tile instances.
 * All per-Gaussian outputs have length P (xy: 2P, conic_opacity: 4P, rgb: 3P,
 * clamped: 3P, cov3D: 6P).                                                    */
int64_t fso_raster_preprocess(
    int P, int D, int M, int H, int W, float tanfovx, float tanfovy, float scale_modifier,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* campos,
    /* out */ int32_t* radii, float* depths, float* xy, float* conic_opacity, float* rgb,
    int32_t* clamped, float* cov3D, uint32_t* tiles_touched, uint32_t* offsets) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
  const float focal_x = (float)W / (2.0f * tanfovx), focal_y = (float)H / (2.0f * tanfovy);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    radii[i] = 0; tiles_touched[i] = 0; depths[i] = 0.f;
    xy[2 * i] = xy[2 * i + 1] = 0.f;
    for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0.f;
    for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0.f; clamped[3 * i + k] = 0; }
    for (int k = 0; k < 6; k++) cov3D[6 * i + k] = 0.f;
    const float* m = means3D + 3 * i;
    float pv[3];
    xform4x3(viewmatrix, m[0], m[1], m[2], pv);
    if (pv[2] <= 0.2f) continue;
    float ph[4];
    xform4x4(projmatrix, m[0], m[1], m[2], ph);
    float pw = 1.0f / (ph[3] + 0.0000001f);
    float ppx = ph[0] * pw, ppy = ph[1] * pw;
    float c6[6];
    if (cov3D_precomp) memcpy(c6, cov3D_precomp + 6 * i, 24);
    else cov3d_from_scale_rot(scales + 3 * i, scale_modifier, rotations + 4 * i, c6);
    memcpy(cov3D + 6 * i, c6, 24);
    float abc[3];
    cov2d(pv, focal_x, focal_y, tanfovx, tanfovy, c6, viewmatrix, abc, NULL, NULL, NULL, NULL, NULL);
    float a = abc[0], b = abc[1], c = abc[2];
    float det = fmaf(-b, b, a * c);
    if (det == 0.0f) continue;
    float det_inv = 1.f / det;
    float con0 = c * det_inv, con1 = -b * det_inv, con2 = a * det_inv;
    float mid = 0.5f * (a + c);
    float sq = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
    float l1 = mid + sq, l2 = mid - sq;
    int my_radius = (int)ceilf(3.f * sqrtf(fmaxf(l1, l2)));
    float px = ndc2pix(ppx, W), py = ndc2pix(ppy, H);
    int rmin[2], rmax[2];
    get_rect(px, py, my_radius, gx, gy, rmin, rmax);
    int area = (rmax[0] - rmin[0]) * (rmax[1] - rmin[1]);
    if (area == 0) continue;
    if (colors_precomp) {
      for (int k = 0; k < 3; k++) rgb[3 * i + k] = colors_precomp[3 * i + k];
    } else {
      int cl[3];
      sh_to_rgb(D, M, m, campos, shs + (size_t)i * M * 3, rgb + 3 * i, cl);
      for (int k = 0; k < 3; k++) clamped[3 * i + k] = cl[k];
    }
    depths[i] = pv[2];
    radii[i] = my_radius;
    xy[2 * i] = px; xy[2 * i + 1] = py;
    conic_opacity[4 * i + 0] = con0; conic_opacity[4 * i + 1] = con1;
    conic_opacity[4 * i + 2] = con2; conic_opacity[4 * i + 3] = opacities[i];
    tiles_touched[i] = (uint32_t)area;
  }
  uint64_t run = 0;
  for (int i = 0; i < P; i++) { run += tiles_touched[i]; offsets[i] = (uint32_t)run; }
  return (int64_t)run;
}

/* Stage 2: duplicateWithKeys + stable radix sort + identifyTileRanges (A.2). */
static void radix_sort_pairs(uint64_t* k, uint32_t* v, uint64_t* k2, uint32_t* v2, int64_t n) {
  for (int pass = 0; pass < 8; pass++) {
    int sh = pass * 8;
    int64_t cnt[257];
    memset(cnt, 0, sizeof cnt);
    for (int64_t i = 0; i < n; i++) cnt[((k[i] >> sh) & 255) + 1]++;
    int skip = 0;
    for (int d = 0; d < 256; d++) if (cnt[d + 1] == n) skip = 1;
    if (skip) continue;
    for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
    for (int64_t i = 0; i < n; i++) { int64_t p = cnt[(k[i] >> sh) & 255]++; k2[p] = k[i]; v2[p] = v[i]; }
    memcpy(k, k2, n * 8); memcpy(v, v2, n * 4);
  }
}

void fso_raster_bin(int P, int H, int W, int64_t R, const int32_t* radii, const float* depths,
                    const float* xy, const uint32_t* offsets,
                    /* out */ uint64_t* keys_sorted, uint32_t* point_list, uint32_t* ranges /* tiles*2 */) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
  uint64_t* k2 = (uint64_t*)malloc((size_t)(R ? R : 1) * 8);
  uint32_t* v2 = (uint32_t*)malloc((size_t)(R ? R : 1) * 4);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (radii[i] <= 0) continue;
    uint32_t off = i ? offsets[i - 1] : 0;
    int rmin[2], rmax[2];
    get_rect(xy[2 * i], xy[2 * i + 1], radii[i], gx, gy, rmin, rmax);
    uint32_t dbits; memcpy(&dbits, depths + i, 4);
    for (int y = rmin[1]; y < rmax[1]; y++)
      for (int x = rmin[0]; x < rmax[0]; x++) {
        uint64_t key = (uint64_t)(y * gx + x);
        key = (key << 32) | dbits;
        keys_sorted[off] = key; point_list[off] = (uint32_t)i; off++;
      }
  }
  radix_sort_pairs(keys_sorted, point_list, k2, v2, R);
  memset(ranges, 0, (size_t)gx * gy * 2 * 4);
  for (int64_t i = 0; i < R; i++) {
    uint32_t t = (uint32_t)(keys_sorted[i] >> 32);
    if (i == 0) ranges[2 * t] = 0;
    else {
      uint32_t pt = (uint32_t)(keys_sorted[i - 1] >> 32);
      if (pt != t) { ranges[2 * pt + 1] = (uint32_t)i; ranges[2 * t] = (uint32_t)i; }
    }
    if (i == R - 1) ranges[2 * t + 1] = (uint32_t)R;
  }
  free(k2); free(v2);
}

/* Stage 3: per-tile front-to-back blend (A.3).  Canonical per-pixel op order:
 *   dx = x_j - px ; dy = y_j - py
 *   ca = -0.5*con.x ; cb = -con.y ; cc = -0.5*con.z      (exact scalings)
 *   t = fma(cb, dy, ca*dx) ; power = fma(cc*dy, dy, t*dx)
 *   alpha = min(0.99, o*expf(power)) ; w = alpha*T ; C = fma(rgb, w, C) ...     */
void fso_raster_render(int H, int W, const float* bg, const uint32_t* ranges,
                       const uint32_t* point_list, const float* xy, const float* conic_opacity,
                       const float* rgb, const float* depths,
                       /* out */ float* out_color /*3HW*/, float* out_depth /*HW*/,
                       float* final_T /*HW*/, uint32_t* n_contrib /*HW*/) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int ty = 0; ty < gy; ty++)
    for (int tx = 0; tx < gx; tx++) {
      uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
      for (int ly = 0; ly < BLOCK_Y; ly++)
        for (int lx = 0; lx < BLOCK_X; lx++) {
          int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
          if (px >= W || py >= H) continue;
          float pxf = (float)px, pyf = (float)py;
          float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;
          uint32_t contributor = 0, last = 0;
          for (uint32_t j = r0; j < r1; j++) {
            contributor++;
            uint32_t id = point_list[j];
            float dx = xy[2 * id] - pxf, dy = xy[2 * id + 1] - pyf;
            const float* co = conic_opacity + 4 * id;
            float ca = -0.5f * co[0], cb = -co[1], cc = -0.5f * co[2];
            float t = fmaf(cb, dy, ca * dx);
            float power = fmaf(cc * dy, dy, t * dx);
            if (power > 0.f) continue;
            float alpha = fminf(0.99f, co[3] * expf(power));
            if (alpha < 1.f / 255.f) continue;
            float test_T = T * (1.f - alpha);
            if (test_T < 0.0001f) break; /* done: nothing further is accumulated */
            float w = alpha * T;
            C0 = fmaf(rgb[3 * id + 0], w, C0);
            C1 = fmaf(rgb[3 * id + 1], w, C1);
            C2 = fmaf(rgb[3 * id + 2], w, C2);
            Dp = fmaf(depths[id], w, Dp);
            T = test_T;
            last = contributor;
          }
          size_t pix = (size_t)py * W + px;
          final_T[pix] = T; n_contrib[pix] = last;
          out_color[0 * (size_t)H * W + pix] = fmaf(T, bg[0], C0);
          out_color[1 * (size_t)H * W + pix] = fmaf(T, bg[1], C1);
          out_color[2 * (size_t)H * W + pix] = fmaf(T, bg[2], C2);
          out_depth[pix] = Dp;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Backward (A.4).  Per-term values in fp32; per-Gaussian sums accumulated in
 * fp64 so that the oracle is independent of thread order.                    */
static inline void atomic_addd(double* p, double v) {
#pragma omp atomic
  *p += v;
}

void fso_raster_backward(
    int P, int D, int M, int H, int W, float tanfovx, float tanfovy, float scale_modifier,
    const float* bg, const float* means3D, const float* shs, const float* colors_precomp,
    const float* scales, const float* rotations, const float* cov3D /*6P as used fwd*/,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    const int32_t* radii, const float* xy, const float* conic_opacity, const float* rgb,
    const float* depths, const int32_t* clamped, const uint32_t* ranges, const uint32_t* point_list,
    const float* final_T, const uint32_t* n_contrib,
    const float* dL_dcolor /*3HW*/, const float* dL_ddepth /*HW or NULL*/, const float* dL_dalpha_out /*HW or NULL*/,
    /* out */ float* dL_dmean2D /*3P*/, float* dL_dconic /*4P: x,y,_,w*/, float* dL_dopacity /*P*/,
    float* dL_drgb /*3P*/, float* dL_dmean3D /*3P*/, float* dL_dcov3D /*6P*/, float* dL_dsh /*P*M*3*/,
    float* dL_dscale /*3P*/, float* dL_drot /*4P*/) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
  const float focal_x = (float)W / (2.0f * tanfovx), focal_y = (float)H / (2.0f * tanfovy);
  /* accumulators: mean2D.x,.y, conic.x,.y,.w, opacity, rgb[3], (depth) */
  double* acc = (double*)calloc((size_t)P * 10, sizeof(double));
  const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int ty = 0; ty < gy; ty++)
    for (int tx = 0; tx < gx; tx++) {
      uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
      for (int ly = 0; ly < BLOCK_Y; ly++)
        for (int lx = 0; lx < BLOCK_X; lx++) {
          int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
          if (px >= W || py >= H) continue;
          size_t pix = (size_t)py * W + px;
          float pxf = (float)px, pyf = (float)py;
          const float T_final = final_T[pix];
          float T = T_final;
          uint32_t contributor = r1 - r0;
          const uint32_t last = n_contrib[pix];
          float accum[3] = {0, 0, 0}, accum_d = 0.f, last_alpha = 0.f, last_c[3] = {0, 0, 0}, last_d = 0.f;
          float dLp[3] = {dL_dcolor[pix], dL_dcolor[HW + pix], dL_dcolor[2 * HW + pix]};
          float dLd = dL_ddepth ? dL_ddepth[pix] : 0.f;
          const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;
          for (uint32_t jj = r1; jj > r0; jj--) {
            uint32_t j = jj - 1;
            contributor--;
            if (contributor >= last) continue;
            uint32_t id = point_list[j];
            float dx = xy[2 * id] - pxf, dy = xy[2 * id + 1] - pyf;
            const float* co = conic_opacity + 4 * id;
            float ca = -0.5f * co[0], cb = -co[1], cc = -0.5f * co[2];
            float t = fmaf(cb, dy, ca * dx);
            float power = fmaf(cc * dy, dy, t * dx);
            if (power > 0.f) continue;
            float G = expf(power);
            float alpha = fminf(0.99f, co[3] * G);
            if (alpha < 1.f / 255.f) continue;
            T = T / (1.f - alpha);
            float w = alpha * T;
            float dL_dalpha = 0.f;
            for (int ch = 0; ch < 3; ch++) {
              float c = rgb[3 * id + ch];
              accum[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * accum[ch];
              last_c[ch] = c;
              dL_dalpha += (c - accum[ch]) * dLp[ch];
              atomic_addd(&acc[(size_t)id * 10 + 6 + ch], (double)(w * dLp[ch]));
            }
            if (dL_ddepth) { /* extension: gradient through the depth channel */
              float dz = depths[id];
              accum_d = last_alpha * last_d + (1.f - last_alpha) * accum_d;
              last_d = dz;
              dL_dalpha += (dz - accum_d) * dLd;
              atomic_addd(&acc[(size_t)id * 10 + 9], (double)(w * dLd));
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            float bg_dot = 0.f;
            for (int ch = 0; ch < 3; ch++) bg_dot += bg[ch] * dLp[ch];
            /* extension: the op's 4th output 1 - final_T depends on alpha_j only through final_T, like the background term */
            if (dL_dalpha_out) bg_dot -= dL_dalpha_out[pix];
            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
            float dL_dG = co[3] * dL_dalpha;
            float gdx = G * dx, gdy = G * dy;
            float dG_ddelx = -gdx * co[0] - gdy * co[1];
            float dG_ddely = -gdy * co[2] - gdx * co[1];
            atomic_addd(&acc[(size_t)id * 10 + 0], (double)(dL_dG * dG_ddelx * ddelx_dx));
            atomic_addd(&acc[(size_t)id * 10 + 1], (double)(dL_dG * dG_ddely * ddely_dy));
            atomic_addd(&acc[(size_t)id * 10 + 2], (double)(-0.5f * gdx * dx * dL_dG));
            atomic_addd(&acc[(size_t)id * 10 + 3], (double)(-0.5f * gdx * dy * dL_dG));
            atomic_addd(&acc[(size_t)id * 10 + 4], (double)(-0.5f * gdy * dy * dL_dG));
            atomic_addd(&acc[(size_t)id * 10 + 5], (double)(G * dL_dalpha));
          }
        }
    }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    dL_dmean2D[3 * i + 0] = (float)acc[(size_t)i * 10 + 0];
    dL_dmean2D[3 * i + 1] = (float)acc[(size_t)i * 10 + 1];
    dL_dmean2D[3 * i + 2] = 0.f;
    dL_dconic[4 * i + 0] = (float)acc[(size_t)i * 10 + 2];
    dL_dconic[4 * i + 1] = (float)acc[(size_t)i * 10 + 3];
    dL_dconic[4 * i + 2] = 0.f;
    dL_dconic[4 * i + 3] = (float)acc[(size_t)i * 10 + 4];
    dL_dopacity[i] = (float)acc[(size_t)i * 10 + 5];
    for (int ch = 0; ch < 3; ch++) dL_drgb[3 * i + ch] = (float)acc[(size_t)i * 10 + 6 + ch];
    for (int k = 0; k < 3; k++) dL_dmean3D[3 * i + k] = 0.f;
    for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = 0.f;
    if (dL_dsh) for (int k = 0; k < M * 3; k++) dL_dsh[(size_t)i * M * 3 + k] = 0.f;
    if (dL_dscale) for (int k = 0; k < 3; k++) dL_dscale[3 * i + k] = 0.f;
    if (dL_drot) for (int k = 0; k < 4; k++) dL_drot[4 * i + k] = 0.f;
    if (!(radii[i] > 0)) continue;
    const float* m = means3D + 3 * i;
    const float* c6 = cov3D + 6 * i;
    /* ---- computeCov2D backward ---- */
    float pv[3];
    xform4x3(viewmatrix, m[0], m[1], m[2], pv);
    float abc[3], a3[3], b3[3], tcl[3];
    int clx, cly;
    cov2d(pv, focal_x, focal_y, tanfovx, tanfovy, c6, viewmatrix, abc, a3, b3, tcl, &clx, &cly);
    float xg = clx ? 0.f : 1.f, yg = cly ? 0.f : 1.f;
    float a = abc[0], b = abc[1], c = abc[2];
    float gxx = dL_dconic[4 * i + 0], gxy = dL_dconic[4 * i + 1], gyy = dL_dconic[4 * i + 3];
    float denom = a * c - b * b;
    float d2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    if (d2inv != 0.f) {
      dL_da = d2inv * (-c * c * gxx + 2.f * b * c * gxy + (denom - a * c) * gyy);
      dL_dc = d2inv * (-a * a * gyy + 2.f * a * b * gxy + (denom - a * c) * gxx);
      dL_db = d2inv * 2.f * (b * c * gxx - (denom + 2.f * b * b) * gxy + a * b * gyy);
      float* g = dL_dcov3D + 6 * i;
      g[0] = a3[0] * a3[0] * dL_da + a3[0] * b3[0] * dL_db + b3[0] * b3[0] * dL_dc;
      g[3] = a3[1] * a3[1] * dL_da + a3[1] * b3[1] * dL_db + b3[1] * b3[1] * dL_dc;
      g[5] = a3[2] * a3[2] * dL_da + a3[2] * b3[2] * dL_db + b3[2] * b3[2] * dL_dc;
      g[1] = 2.f * a3[0] * a3[1] * dL_da + (a3[0] * b3[1] + a3[1] * b3[0]) * dL_db + 2.f * b3[0] * b3[1] * dL_dc;
      g[2] = 2.f * a3[0] * a3[2] * dL_da + (a3[0] * b3[2] + a3[2] * b3[0]) * dL_db + 2.f * b3[0] * b3[2] * dL_dc;
      g[4] = 2.f * a3[2] * a3[1] * dL_da + (a3[1] * b3[2] + a3[2] * b3[1]) * dL_db + 2.f * b3[1] * b3[2] * dL_dc;
    }
    float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
    float Sa[3], Sb[3];
    for (int k = 0; k < 3; k++) {
      Sa[k] = S[3 * k] * a3[0] + S[3 * k + 1] * a3[1] + S[3 * k + 2] * a3[2];
      Sb[k] = S[3 * k] * b3[0] + S[3 * k + 1] * b3[1] + S[3 * k + 2] * b3[2];
    }
    float dTa[3], dTb[3];
    for (int k = 0; k < 3; k++) {
      dTa[k] = 2.f * Sa[k] * dL_da + Sb[k] * dL_db;
      dTb[k] = 2.f * Sb[k] * dL_dc + Sa[k] * dL_db;
    }
    float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
    for (int k = 0; k < 3; k++) {
      float R0 = viewmatrix[k * 4 + 0], R1 = viewmatrix[k * 4 + 1], R2 = viewmatrix[k * 4 + 2];
      dJ00 += R0 * dTa[k]; dJ02 += R2 * dTa[k];
      dJ11 += R1 * dTb[k]; dJ12 += R2 * dTb[k];
    }
    float tz = 1.f / tcl[2], tz2 = tz * tz, tz3 = tz2 * tz;
    float dtx = xg * -focal_x * tz2 * dJ02;
    float dty = yg * -focal_y * tz2 * dJ12;
    float dtz = -focal_x * tz2 * dJ00 - focal_y * tz2 * dJ11 + (2.f * focal_x * tcl[0]) * tz3 * dJ02 +
                (2.f * focal_y * tcl[1]) * tz3 * dJ12;
    float dm[3];
    for (int k = 0; k < 3; k++)
      dm[k] = viewmatrix[4 * k + 0] * dtx + viewmatrix[4 * k + 1] * dty + viewmatrix[4 * k + 2] * dtz;
    /* ---- extension: depth channel -> mean (p_view.z = row 2 of view) ---- */
    if (dL_ddepth) {
      float gz = (float)acc[(size_t)i * 10 + 9];
      for (int k = 0; k < 3; k++) dm[k] += viewmatrix[4 * k + 2] * gz;
    }
    /* ---- preprocess backward: projection term ---- */
    float ph[4];
    xform4x4(projmatrix, m[0], m[1], m[2], ph);
    float mw = 1.0f / (ph[3] + 0.0000001f);
    float mul1 = ph[0] * mw * mw, mul2 = ph[1] * mw * mw;
    float g2x = dL_dmean2D[3 * i], g2y = dL_dmean2D[3 * i + 1];
    const float* pm = projmatrix;
    dm[0] += (pm[0] * mw - pm[3] * mul1) * g2x + (pm[1] * mw - pm[3] * mul2) * g2y;
    dm[1] += (pm[4] * mw - pm[7] * mul1) * g2x + (pm[5] * mw - pm[7] * mul2) * g2y;
    dm[2] += (pm[8] * mw - pm[11] * mul1) * g2x + (pm[9] * mw - pm[11] * mul2) * g2y;
    /* ---- SH backward ---- */
    if (!colors_precomp && shs && dL_dsh) {
      const float* sh = shs + (size_t)i * M * 3;
      float* gsh = dL_dsh + (size_t)i * M * 3;
      float dox = m[0] - campos[0], doy = m[1] - campos[1], doz = m[2] - campos[2];
      float len = sqrtf(dox * dox + doy * doy + doz * doz);
      float x = dox / len, y = doy / len, z = doz / len;
      float gL[3];
      for (int ch = 0; ch < 3; ch++) gL[ch] = clamped[3 * i + ch] ? 0.f : dL_drgb[3 * i + ch];
      float ddx = 0, ddy = 0, ddz = 0; /* dL/ddir */
      for (int ch = 0; ch < 3; ch++) {
        float g = gL[ch];
        const float* s = sh + ch; /* stride 3 */
#define SHV(k) s[(k) * 3]
        float* go = gsh + ch;
        go[0] = SH_C0 * g;
        float rx = 0, ry = 0, rz = 0;
        if (D > 0) {
          go[1 * 3] = -SH_C1 * y * g; go[2 * 3] = SH_C1 * z * g; go[3 * 3] = -SH_C1 * x * g;
          rx = -SH_C1 * SHV(3); ry = -SH_C1 * SHV(1); rz = SH_C1 * SHV(2);
          if (D > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy_ = x * y, yz = y * z, xz = x * z;
            go[4 * 3] = SH_C2[0] * xy_ * g; go[5 * 3] = SH_C2[1] * yz * g;
            go[6 * 3] = SH_C2[2] * (2.f * zz - xx - yy) * g;
            go[7 * 3] = SH_C2[3] * xz * g; go[8 * 3] = SH_C2[4] * (xx - yy) * g;
            rx += SH_C2[0] * y * SHV(4) + SH_C2[2] * 2.f * -x * SHV(6) + SH_C2[3] * z * SHV(7) + SH_C2[4] * 2.f * x * SHV(8);
            ry += SH_C2[0] * x * SHV(4) + SH_C2[1] * z * SHV(5) + SH_C2[2] * 2.f * -y * SHV(6) + SH_C2[4] * 2.f * -y * SHV(8);
            rz += SH_C2[1] * y * SHV(5) + SH_C2[2] * 2.f * 2.f * z * SHV(6) + SH_C2[3] * x * SHV(7);
            if (D > 2) {
              go[9 * 3] = SH_C3[0] * y * (3.f * xx - yy) * g;
              go[10 * 3] = SH_C3[1] * xy_ * z * g;
              go[11 * 3] = SH_C3[2] * y * (4.f * zz - xx - yy) * g;
              go[12 * 3] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * g;
              go[13 * 3] = SH_C3[4] * x * (4.f * zz - xx - yy) * g;
              go[14 * 3] = SH_C3[5] * z * (xx - yy) * g;
              go[15 * 3] = SH_C3[6] * x * (xx - 3.f * yy) * g;
              rx += SH_C3[0] * SHV(9) * 3.f * 2.f * xy_ + SH_C3[1] * SHV(10) * yz + SH_C3[2] * SHV(11) * -2.f * xy_ +
                    SH_C3[3] * SHV(12) * -3.f * 2.f * xz + SH_C3[4] * SHV(13) * (-3.f * xx + 4.f * zz - yy) +
                    SH_C3[5] * SHV(14) * 2.f * xz + SH_C3[6] * SHV(15) * 3.f * (xx - yy);
              ry += SH_C3[0] * SHV(9) * 3.f * (xx - yy) + SH_C3[1] * SHV(10) * xz + SH_C3[2] * SHV(11) * (-3.f * yy + 4.f * zz - xx) +
                    SH_C3[3] * SHV(12) * -3.f * 2.f * yz + SH_C3[4] * SHV(13) * -2.f * xy_ + SH_C3[5] * SHV(14) * -2.f * yz +
                    SH_C3[6] * SHV(15) * -3.f * 2.f * xy_;
              rz += SH_C3[1] * SHV(10) * xy_ + SH_C3[2] * SHV(11) * 4.f * 2.f * yz + SH_C3[3] * SHV(12) * 3.f * (2.f * zz - xx - yy) +
                    SH_C3[4] * SHV(13) * 4.f * 2.f * xz + SH_C3[5] * SHV(14) * (xx - yy);
            }
          }
        }
#undef SHV
        ddx += rx * g; ddy += ry * g; ddz += rz * g;
      }
      /* dnormvdv(dir_orig, dL_ddir) */
      float sum2 = dox * dox + doy * doy + doz * doz;
      float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dm[0] += ((sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * inv32;
      dm[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * inv32;
      dm[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * inv32;
    }
    for (int k = 0; k < 3; k++) dL_dmean3D[3 * i + k] = dm[k];
    /* ---- computeCov3D backward (only when scales/rotations were given) ---- */
    if (scales && rotations && dL_dscale && dL_drot) {
      const float* q = rotations + 4 * i;
      const float* s = scales + 3 * i;
      float r = q[0], x = q[1], y = q[2], z = q[3];
      float Rm[9] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                     2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                     2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
      float sv[3] = {scale_modifier * s[0], scale_modifier * s[1], scale_modifier * s[2]};
      const float* g = dL_dcov3D + 6 * i;
      /* symmetric dL/dSigma with off-diagonals halved */
      float G[9] = {g[0], 0.5f * g[1], 0.5f * g[2], 0.5f * g[1], g[3], 0.5f * g[4], 0.5f * g[2], 0.5f * g[4], g[5]};
      /* Sigma = M M^T, M = R diag(sv):  dL/dM = 2 G M */
      float Mx[9], dM[9];
      for (int a_ = 0; a_ < 3; a_++) for (int b_ = 0; b_ < 3; b_++) Mx[3 * a_ + b_] = Rm[3 * a_ + b_] * sv[b_];
      for (int a_ = 0; a_ < 3; a_++) for (int b_ = 0; b_ < 3; b_++) {
        float acc_ = 0; for (int k = 0; k < 3; k++) acc_ += G[3 * a_ + k] * Mx[3 * k + b_];
        dM[3 * a_ + b_] = 2.f * acc_;
      }
      /* M[a][b] = R[a][b]*sv[b] */
      float dR[9];
      for (int b_ = 0; b_ < 3; b_++) {
        float acc_ = 0;
        for (int a_ = 0; a_ < 3; a_++) { acc_ += Rm[3 * a_ + b_] * dM[3 * a_ + b_]; dR[3 * a_ + b_] = dM[3 * a_ + b_] * sv[b_]; }
        dL_dscale[3 * i + b_] = scale_modifier * acc_;
      }
      /* dR -> dq for the (unnormalised) formula above */
      float dr = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
      float dxq = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2.f * x * dR[8]);
      float dyq = 2.f * (-2.f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2.f * y * dR[8]);
      float dzq = 2.f * (-2.f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
      dL_drot[4 * i + 0] = dr; dL_drot[4 * i + 1] = dxq; dL_drot[4 * i + 2] = dyq; dL_drot[4 * i + 3] = dzq;
    }
  }
  free(acc);
}

int fso_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void fso_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
