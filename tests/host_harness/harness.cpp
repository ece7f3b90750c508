// TEST INFRASTRUCTURE: compiles the PRODUCT's per-Gaussian math header
// (freesplat_b200/csrc/raster_math.cuh) for the host so that its integer decisions can be
// compared bit-for-bit with oracle/raster_oracle.c without a GPU.  Not linked into the product.
#include "../../freesplat_b200/csrc/raster_math.cuh"
#include <cstring>

extern "C" void fsh_project(int P, int H, int W, float tanx, float tany, const float* means, const float* cov6,
                            const float* view, const float* proj, const float* opac, int deg, int M, const float* shs,
                            const float* campos,
                            int* radii, int* rect, float* xy, float* conic, float* depth, float* rgb, int* clamp,
                            float* ext) {
  for (int i = 0; i < P; i++) {
    fsm::Projected p = fsm::project_gaussian(means + 3 * i, cov6 + 6 * i, view, proj, tanx, tany, H, W);
    radii[i] = p.radius;
    rect[4 * i] = p.x0; rect[4 * i + 1] = p.y0; rect[4 * i + 2] = p.x1; rect[4 * i + 3] = p.y1;
    xy[2 * i] = p.px; xy[2 * i + 1] = p.py;
    conic[3 * i] = p.con_x; conic[3 * i + 1] = p.con_y; conic[3 * i + 2] = p.con_z;
    depth[i] = p.depth;
    float sh[48];
    for (int k = 0; k < 48; k++) sh[k] = k < M * 3 ? shs[(size_t)i * M * 3 + k] : 0.f;
    clamp[i] = fsm::sh_to_rgb(deg, means + 3 * i, campos, sh, rgb + 3 * i);
    fsm::alpha_extent(p.con_x, p.con_y, p.con_z, opac[i], ext + 2 * i, ext + 2 * i + 1);
  }
}
