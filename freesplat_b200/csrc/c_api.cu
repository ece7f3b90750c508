// c_api.cu -- extern "C" entry points of libfreesplat_b200.so (see include/freesplat_b200.h).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace fs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return FS_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return FS_ERR_CUDA;
}

}  // namespace fs

using namespace fs;

extern "C" {

int fs_abi_version(void) { return FS_ABI_VERSION; }
int fs_struct_size(int32_t which) {
  switch (which) {
    case 0: return (int)sizeof(FsRasterFwdArgs);
    case 1: return (int)sizeof(FsRasterBwdArgs);
    case 2: return (int)sizeof(FsCostVolumeArgs);
    case 3: return (int)sizeof(FsPtfArgs);
    case 4: return (int)sizeof(FsPtfGruArgs);
    case 12: return (int)sizeof(FsGruBwdDataArgs);
    case 13: return (int)sizeof(FsGruBwdWeightsArgs);
    case 5: return (int)sizeof(FsAdapterArgs);
    case 6: return (int)sizeof(FsDepthHeadArgs);
    case 7: return (int)sizeof(FsBackprojectArgs);
    case 8: return (int)sizeof(FsPlyArgs);
    case 9: return (int)sizeof(FsPtfMergeBwdArgs);
    case 10: return (int)sizeof(FsAdapterBwdArgs);
    case 11: return (int)sizeof(FsDepthHeadBwdArgs);
    default: return -1;
  }
}
const char* fs_last_error(void) { return g_err; }

int fs_device_sm_count(void) {
  int dev = 0, n = 0;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (int rc = check_cuda(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;
  return n;
}

#define FS_REQUIRE(cond, msg)            \
  do {                                   \
    if (!(cond)) {                       \
      set_error("invalid argument: %s", msg); \
      return FS_ERR_INVALID_ARG;         \
    }                                    \
  } while (0)

int fs_raster_forward(const FsRasterFwdArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr, "args is NULL");
  FS_REQUIRE(a->P >= 0 && a->V >= 1 && a->V <= 65535 && a->H >= 1 && a->W >= 1, "bad sizes");
  FS_REQUIRE((a->shs != nullptr) != (a->colors_precomp != nullptr), "provide exactly one of shs / colors_precomp");
  FS_REQUIRE((a->cov3D_precomp != nullptr) != (a->scales != nullptr && a->rotations != nullptr),
             "provide exactly one of cov3D_precomp / (scales, rotations)");
  FS_REQUIRE(a->sh_degree >= 0 && a->sh_degree <= 3, "sh_degree must be 0..3");
  FS_REQUIRE(a->shs == nullptr || ((a->sh_degree + 1) * (a->sh_degree + 1) <= a->M && a->M <= 16), "M too small for sh_degree (or > 16)");
  FS_REQUIRE(a->capacity >= 0 && a->capacity <= 0xffffffffll, "capacity out of range");
  FS_REQUIRE(a->views && a->out_color && a->out_depth && a->final_T && a->n_contrib && a->radii && a->rec && a->clamped && a->tile_count && a->tile_cursor && a->ranges && a->status,
             "NULL buffer");
  FS_REQUIRE(a->P == 0 || (a->means3D && a->opacities), "NULL input");
  FS_REQUIRE(a->capacity == 0 || (a->keybuf && a->point_list), "NULL key buffers");
  FS_REQUIRE(a->bins == nullptr || (a->bin_cap >= 1 && a->bin_cap <= 4096), "bin_cap must be 1..4096 when bins is given");
  FS_REQUIRE((long long)tiles_x(a->W) * tiles_y(a->H) * a->V < (1ll << 31), "too many tiles");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int rc;
  const int stages = (a->stages & 7) ? a->stages : (a->stages | FS_STAGE_PREPROCESS | FS_STAGE_BINNING | FS_STAGE_RENDER);
  if ((stages & FS_STAGE_PREPROCESS) && (rc = launch_preprocess(*a, s))) return rc;
  if ((stages & FS_STAGE_BINNING) && (rc = launch_binning(*a, s))) return rc;
  if ((stages & FS_STAGE_RENDER) && (rc = launch_render_fwd(*a, s))) return rc;
  return FS_OK;
}

int fs_raster_backward(const FsRasterBwdArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr, "args is NULL");
  FS_REQUIRE(a->P >= 0 && a->V >= 1 && a->V <= 65535 && a->H >= 1 && a->W >= 1, "bad sizes");
  FS_REQUIRE((a->cov3D_precomp != nullptr) != (a->scales != nullptr && a->rotations != nullptr),
             "provide exactly one of cov3D_precomp / (scales, rotations)");
  FS_REQUIRE(a->views && a->rec && a->radii && a->clamped && a->ranges && a->final_T && a->n_contrib &&
                 a->status && a->dL_dcolor && a->dL_dscreen && a->dL_dmeans2D && a->dL_dmeans3D && a->dL_dopacities,
             "NULL buffer");
  FS_REQUIRE(!a->has_depth_grad || a->dL_ddepth, "has_depth_grad set but dL_ddepth is NULL");
  FS_REQUIRE(a->peer_delta == nullptr || (a->world >= 1 && a->shard_rows >= 1), "peer_delta given but world / shard_rows are not");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = launch_render_bwd(*a, s))) return rc;
  if ((rc = launch_preprocess_bwd(*a, s))) return rc;
  return FS_OK;
}

int fs_mark_visible(int32_t P, const float* means3D, const float* view, uint8_t* visible, void* stream) {
  FS_REQUIRE(P >= 0 && (P == 0 || (means3D && view && visible)), "bad arguments");
  return launch_mark_visible(P, means3D, view, visible, reinterpret_cast<cudaStream_t>(stream));
}

int fs_camera_records(int32_t V, const float* extrinsics, const float* intrinsics, const float* near, const float* far,
                      const float* bg, int32_t scale_invariant, float* views, void* stream) {
  FS_REQUIRE(V >= 0 && (V == 0 || (extrinsics && intrinsics && near && far && bg && views)), "bad arguments");
  return launch_camera_records(V, extrinsics, intrinsics, near, far, bg, scale_invariant, views, reinterpret_cast<cudaStream_t>(stream));
}

static int check_cv(const FsCostVolumeArgs* a) {
  FS_REQUIRE(a != nullptr, "args is NULL");
  FS_REQUIRE(a->B >= 1 && a->K >= 1 && a->K <= 16 && a->H >= 1 && a->W >= 1 && a->D >= 1, "bad sizes (K must be 1..16)");
  FS_REQUIRE(a->C == 48, "matching feature width must be 48 (encoder_freesplat.py:160)");
  FS_REQUIRE((long long)a->H * a->W < (1ll << 30) && a->B <= 65535 && a->D <= 65535 * 8, "feature map too large");
  FS_REQUIRE(a->cur_feats && a->src_feats && a->proj && a->cur_invK && a->planes && a->mlp, "NULL input");
  return FS_OK;
}

int fs_cost_volume_forward(const FsCostVolumeArgs* a, void* stream) {
  if (int rc = check_cv(a)) return rc;
  FS_REQUIRE(a->out != nullptr && a->src_packed != nullptr, "out / src_packed is NULL");
  return launch_cost_volume_fwd(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_cost_volume_backward(const FsCostVolumeArgs* a, void* stream) {
  if (int rc = check_cv(a)) return rc;
  FS_REQUIRE(a->dL_dout && a->dL_dcur && a->dL_dsrc && a->dL_dmlp && a->src_packed && a->dsrc_packed, "NULL gradient / scratch buffer");
  return launch_cost_volume_bwd(*a, reinterpret_cast<cudaStream_t>(stream));
}

static int check_ptf(const FsPtfArgs* a) {
  FS_REQUIRE(a != nullptr, "args is NULL");
  FS_REQUIRE(a->H >= 1 && a->W >= 1 && a->F >= 1 && a->n_upper >= 0, "bad sizes");
  FS_REQUIRE((long long)a->H * a->W < (1ll << 30), "image too large");
  FS_REQUIRE(a->coords && a->counts_in && a->v_depth && a->E_inv && a->K_px && a->zbuf && a->pix && a->zeta && a->match &&
                 a->append && a->block_counts && a->pair_j && a->pair_p && a->counts_out, "NULL buffer");
  return FS_OK;
}

int fs_ptf_view_setup(int32_t V, int32_t H, int32_t W, const float* extrinsics, const float* intrinsics, float* E_inv, float* K_px,
                      void* stream) {
  FS_REQUIRE(V >= 0 && H >= 1 && W >= 1 && (V == 0 || (extrinsics && intrinsics && E_inv && K_px)), "bad arguments");
  return launch_ptf_view_setup(V, H, W, extrinsics, intrinsics, E_inv, K_px, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ptf_match(const FsPtfArgs* a, void* stream) {
  if (int rc = check_ptf(a)) return rc;
  return launch_ptf_match(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ptf_merge(const FsPtfArgs* a, void* stream) {
  if (int rc = check_ptf(a)) return rc;
  FS_REQUIRE(a->feats && a->dens && a->wemb && a->ext && a->depth && a->v_feats && a->v_coords && a->v_dens && a->v_wemb &&
                 a->v_ext && a->o_feats && a->o_coords && a->o_dens && a->o_wemb && a->o_ext && a->o_depth, "NULL state buffer");
  return launch_ptf_merge(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ptf_pool_update(const FsPtfArgs* a, void* stream) {
  if (int rc = check_ptf(a)) return rc;
  FS_REQUIRE(a->feats && a->dens && a->wemb && a->ext && a->depth && a->v_feats && a->v_coords && a->v_dens && a->v_wemb && a->v_ext,
             "NULL state buffer");
  FS_REQUIRE((a->F & 3) == 0 && (a->n_upper == 0 || a->gru_out), "F must be a multiple of 4; gru_out is NULL");
  return launch_ptf_pool_update(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ptf_pool_order(int32_t n_upper, const int32_t* counts, const int32_t* phys_in, const uint8_t* match, int32_t* block_scratch,
                      int32_t* phys_out, void* stream) {
  FS_REQUIRE(n_upper >= 0 && (n_upper == 0 || (counts && match && block_scratch && phys_out)), "bad arguments");
  return launch_ptf_pool_order(n_upper, counts, phys_in, match, block_scratch, phys_out, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ptf_pool_gather(int32_t n_upper, const int32_t* n_dev, const int32_t* phys, int32_t F, const float* feats, const float* coords,
                       const float* dens, const float* wemb, const float* ext, const float* depth, float* o_feats, float* o_coords,
                       float* o_dens, float* o_wemb, float* o_ext, float* o_depth, void* stream) {
  FS_REQUIRE(n_upper >= 0 && F >= 4 && (F & 3) == 0, "bad sizes (F must be a multiple of 4)");
  FS_REQUIRE(n_upper == 0 || (n_dev && phys && feats && coords && dens && wemb && ext && depth && o_feats && o_coords && o_dens && o_wemb &&
                              o_ext && o_depth), "NULL buffer");
  return launch_ptf_pool_gather(n_upper, n_dev, phys, F, feats, coords, dens, wemb, ext, depth, o_feats, o_coords, o_dens, o_wemb, o_ext,
                                o_depth, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ptf_merge_backward(const FsPtfMergeBwdArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->H >= 1 && a->W >= 1 && a->F >= 4 && (a->F & 3) == 0 && a->N >= 0 && a->n_keep >= 0 && a->n_match >= 0,
             "bad sizes (F must be a multiple of 4)");
  FS_REQUIRE(a->coords && a->dens && a->ext && a->depth && a->v_coords && a->v_dens && a->v_depth && a->v_ext && a->match && a->pix &&
                 a->map_old && a->map_px, "NULL forward state");
  FS_REQUIRE(a->d_feats && a->d_coords && a->d_dens && a->d_wemb && a->d_ext && a->d_depth && a->dv_feats && a->dv_coords && a->dv_dens &&
                 a->dv_wemb && a->dv_depth && (a->n_match == 0 || a->d_gru), "NULL output buffer");
  return launch_ptf_merge_bwd(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ptf_gru_inputs(int32_t M, int32_t F, const int32_t* pair_j, const int32_t* pair_p, const float* feats, const float* dens,
                      const float* wemb, const float* v_feats, const float* v_dens, const float* v_wemb, float* A1, void* stream) {
  FS_REQUIRE(M >= 0 && F >= 1 && (M == 0 || (pair_j && pair_p && feats && dens && wemb && v_feats && v_dens && v_wemb && A1)), "bad arguments");
  return launch_ptf_gru_inputs(M, F, pair_j, pair_p, feats, dens, wemb, v_feats, v_dens, v_wemb, A1, reinterpret_cast<cudaStream_t>(stream));
}
int fs_ptf_gru_update(int32_t M, int32_t F, const float* A1, const float* r_lin, float* U, void* stream) {
  FS_REQUIRE(M >= 0 && F >= 1 && (M == 0 || (A1 && r_lin && U)), "bad arguments");
  return launch_ptf_gru_update(M, F, A1, r_lin, U, reinterpret_cast<cudaStream_t>(stream));
}
int fs_ptf_gru_output(int32_t M, int32_t F, const float* A1, const float* z_lin, const float* q_lin, float* out, void* stream) {
  FS_REQUIRE(M >= 0 && F >= 1 && (M == 0 || (A1 && z_lin && q_lin && out)), "bad arguments");
  return launch_ptf_gru_output(M, F, A1, z_lin, q_lin, out, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ptf_gru_output_backward(int32_t M, int32_t F, const float* A1, const float* z_lin, const float* q_lin, const float* g_out, float* dz_lin,
                               float* dq_lin, float* dA1, void* stream) {
  FS_REQUIRE(M >= 0 && F >= 1 && (M == 0 || (A1 && z_lin && q_lin && g_out && dz_lin && dq_lin && dA1)), "bad arguments");
  return launch_ptf_gru_output_bwd(M, F, A1, z_lin, q_lin, g_out, dz_lin, dq_lin, dA1, reinterpret_cast<cudaStream_t>(stream));
}
int fs_ptf_gru_update_backward(int32_t M, int32_t F, const float* A1, const float* r_lin, const float* dU, float* dr_lin, float* dA1, void* stream) {
  FS_REQUIRE(M >= 0 && F >= 1 && (M == 0 || (A1 && r_lin && dU && dr_lin && dA1)), "bad arguments");
  return launch_ptf_gru_update_bwd(M, F, A1, r_lin, dU, dr_lin, dA1, reinterpret_cast<cudaStream_t>(stream));
}
int fs_ptf_gru_inputs_backward(int32_t M, int32_t F, const int32_t* pair_j, const int32_t* pair_p, const float* dens, const float* wemb,
                               const float* v_dens, const float* v_wemb, const float* dA1, float* d_feats, float* d_dens, float* d_wemb,
                               float* dv_feats, float* dv_dens, float* dv_wemb, void* stream) {
  FS_REQUIRE(M >= 0 && F >= 1 && (M == 0 || (pair_j && pair_p && dens && wemb && v_dens && v_wemb && dA1 && d_feats && d_dens && d_wemb &&
                                           dv_feats && dv_dens && dv_wemb)), "bad arguments");
  return launch_ptf_gru_inputs_bwd(M, F, pair_j, pair_p, dens, wemb, v_dens, v_wemb, dA1, d_feats, d_dens, d_wemb, dv_feats, dv_dens, dv_wemb,
                                   reinterpret_cast<cudaStream_t>(stream));
}

int fs_gaussian_head(const FsAdapterArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->N >= 0 && a->H >= 1 && a->W >= 1 && a->sh_degree >= 0 && a->sh_degree <= 3, "bad arguments");
  FS_REQUIRE(a->N == 0 || (a->raw && a->depths && a->opacities && a->coords && a->ext && a->K && a->means && a->covariances &&
                           a->harmonics && a->opacities_out && a->scales && a->rotations), "NULL buffer");
  return launch_gaussian_head(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_gaussian_head_backward(const FsAdapterBwdArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->N >= 0 && a->H >= 1 && a->W >= 1 && a->sh_degree >= 0 && a->sh_degree <= 3, "bad arguments");
  FS_REQUIRE(a->N == 0 || (a->raw && a->depths && a->ext && a->K && a->d_raw && a->d_depths && a->d_opacities && a->d_coords), "NULL buffer");
  return launch_gaussian_head_bwd(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_depth_head_backward(const FsDepthHeadBwdArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->B >= 0 && a->D >= 1 && a->h >= 1 && a->w >= 1 && a->B <= 65535, "bad sizes");
  FS_REQUIRE(a->B == 0 || (a->logits && a->candi && a->stats && a->d_logits), "NULL buffer");
  FS_REQUIRE(a->B == 0 || !(a->upsample && a->g_weights_up) || (a->argmax_up && a->D <= 256), "g_weights_up needs argmax_up scratch and D <= 256");
  return launch_depth_head_bwd(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_image_u8(int32_t B, int32_t C, int32_t H, int32_t W, const float* images, uint8_t* out, void* stream) {
  FS_REQUIRE(B >= 0 && H >= 0 && W >= 0 && (C == 1 || C == 3 || C == 4), "bad sizes (1, 3 or 4 channels)");
  FS_REQUIRE((long long)B * H * W == 0 || (images && out), "NULL buffer");
  return launch_image_u8(B, C, H, W, images, out, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ply_vertices(const FsPlyArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->N >= 0 && a->d_sh >= 1, "bad sizes");
  FS_REQUIRE(a->N == 0 || (a->means && a->scales && a->rotations && a->harmonics && a->opacities && a->table), "NULL buffer");
  FS_REQUIRE(a->scale_factor > 0.f, "scale_factor must be positive");
  return launch_ply_vertices(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_backproject(const FsBackprojectArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->V >= 0 && a->V <= 65535 && a->H >= 1 && a->W >= 1 && (long long)a->H * a->W < (1ll << 30), "bad sizes");
  FS_REQUIRE(a->V == 0 || (a->depth && a->K && a->c2w && a->means), "NULL buffer");
  return launch_backproject(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_backproject_backward(const FsBackprojectArgs* a, const float* g_means, float* d_depth, void* stream) {
  FS_REQUIRE(a != nullptr && a->V >= 0 && a->V <= 65535 && a->H >= 1 && a->W >= 1 && (long long)a->H * a->W < (1ll << 30), "bad sizes");
  FS_REQUIRE(a->V == 0 || (a->K && a->c2w && g_means && d_depth), "NULL buffer");
  return launch_backproject_bwd(*a, g_means, d_depth, reinterpret_cast<cudaStream_t>(stream));
}

int fs_depth_head(const FsDepthHeadArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->B >= 0 && a->D >= 1 && a->h >= 1 && a->w >= 1 && a->B <= 65535, "bad sizes");
  FS_REQUIRE((long long)a->h * a->w < (1ll << 29), "feature map too large");
  FS_REQUIRE(a->B == 0 || (a->logits && a->candi && a->expect && a->depth), "NULL buffer");
  FS_REQUIRE(a->B == 0 || !a->upsample || (a->depth_up && a->weights_up), "upsample set but depth_up / weights_up is NULL");
  return launch_depth_head(*a, reinterpret_cast<cudaStream_t>(stream));
}

int fs_ptf_gru(const FsPtfGruArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->M >= 0, "bad arguments");
  FS_REQUIRE(a->M == 0 || (a->pair_j && a->pair_p && a->feats && a->dens && a->wemb && a->v_feats && a->v_dens && a->v_wemb && a->W_r0 &&
                           a->W_z0 && a->W_r2 && a->W_z2 && a->W_n0 && a->W_n2 && a->biases && a->wscratch && a->out), "NULL buffer");
  return launch_ptf_gru_tc(*a, reinterpret_cast<cudaStream_t>(stream));
}
int64_t fs_ptf_gru_wscratch_bytes(void) { return (int64_t)ptf_gru_wscratch_bytes(); }

int fs_ptf_gru_bwd_data(const FsGruBwdDataArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->M >= 0 && a->N >= 4 && a->N <= 192 && a->N % 4 == 0, "bad sizes (N % 4 == 0, 4 <= N <= 192)");
  FS_REQUIRE(a->mode >= 0 && a->mode <= 3, "mode must be 0 (store), 1 (ReLU mask), 2 (accumulate) or 3 (update-gate epilogue)");
  FS_REQUIRE(a->mode != 3 || (a->N == 152 && a->ldc >= 176 && (a->M == 0 || (a->h && a->r_lin && a->dr_lin && a->ldh >= 64 && a->ldh % 4 == 0))),
             "mode 3 needs N == 152, ldc >= 176, h / r_lin / dr_lin");
  FS_REQUIRE(a->lda >= 64 && a->lda % 4 == 0 && a->ldc >= a->N && a->ldc % 4 == 0, "bad leading dimensions");
  FS_REQUIRE(a->M == 0 || (a->A && a->W && a->C), "NULL buffer");
  FS_REQUIRE(a->mode != 1 || (a->N == 64 && (a->M == 0 || (a->mask && a->ldm >= 64 && a->ldm % 4 == 0))),
             "mode 1 needs N == 64 and a mask with ldm >= 64, ldm % 4 == 0");
  return launch_ptf_gru_bwd_data(*a, reinterpret_cast<cudaStream_t>(stream));
}
int fs_ptf_gru_bwd_weights(const FsGruBwdWeightsArgs* a, void* stream) {
  FS_REQUIRE(a != nullptr && a->M >= 0 && a->nx0 >= 1 && a->nx1 >= 0, "bad sizes");
  FS_REQUIRE(a->ldg % 16 == 0 && a->ldg <= 256 && a->nx0 + a->nx1 < a->ldg, "ldg % 16 == 0, nx0 + nx1 < ldg <= 256");
  FS_REQUIRE(a->G != nullptr, "G is NULL");
  FS_REQUIRE(a->M == 0 || (a->Y0 && a->X0 && a->ldy0 >= 64 && a->ldx0 >= a->nx0), "NULL buffer / bad leading dimension");
  FS_REQUIRE(a->M == 0 || !a->Y1 || a->ldy1 >= 64, "ldy1 < 64");
  FS_REQUIRE(a->M == 0 || ((a->X1 != nullptr) == (a->nx1 > 0) && (!a->X1 || a->ldx1 >= a->nx1)), "X1 / nx1 mismatch");
  return launch_ptf_gru_bwd_weights(*a, reinterpret_cast<cudaStream_t>(stream));
}

// ---- CUDA graphs: a launch sequence of fs_* calls with static arguments, replayed by ONE cudaGraphLaunch ----
int fs_graph_capture_begin(void** stream_out) {
  FS_REQUIRE(stream_out != nullptr, "stream_out is NULL");
  cudaStream_t s = nullptr;
  if (int rc = check_cuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreateWithFlags")) return rc;
  // thread-local mode: allocator activity of other host threads does not invalidate the capture
  if (int rc = check_cuda(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture")) {
    cudaStreamDestroy(s);
    return rc;
  }
  *stream_out = s;
  return FS_OK;
}

int fs_graph_capture_end(void* stream, void** graph_exec_out) {
  FS_REQUIRE(stream != nullptr && graph_exec_out != nullptr, "NULL argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaGraph_t g = nullptr;
  cudaGraphExec_t ge = nullptr;
  int rc = check_cuda(cudaStreamEndCapture(s, &g), "cudaStreamEndCapture");
  if (!rc) rc = check_cuda(cudaGraphInstantiate(&ge, g, 0), "cudaGraphInstantiate");
  if (g) cudaGraphDestroy(g);
  cudaStreamDestroy(s);
  if (rc) return rc;
  *graph_exec_out = ge;
  return FS_OK;
}

int fs_graph_launch(void* graph_exec, void* stream) {
  FS_REQUIRE(graph_exec != nullptr, "graph is NULL");
  return check_cuda(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), reinterpret_cast<cudaStream_t>(stream)), "cudaGraphLaunch");
}

int fs_graph_destroy(void* graph_exec) {
  if (graph_exec == nullptr) return FS_OK;
  return check_cuda(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec)), "cudaGraphExecDestroy");
}

}  // extern "C"
