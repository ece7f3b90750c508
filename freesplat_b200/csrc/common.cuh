// common.cuh -- shared declarations of the sm_100a kernels behind the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/freesplat_b200.h"

namespace fs {

constexpr int kThreads = 256;
constexpr int kViewFloats = FS_VIEW_FLOATS;
constexpr int kRecFloats = FS_REC_FLOATS;
// keys sorted in shared memory by one block (64-bit keys); larger tiles sort in global memory.
constexpr int kSortSmemKeys = 4096;

// error plumbing (thread-local message, see c_api.cu)
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// launchers (each returns FsStatus)
int launch_preprocess(const FsRasterFwdArgs& a, cudaStream_t s);     // raster_pre.cu
int launch_binning(const FsRasterFwdArgs& a, cudaStream_t s);        // raster_bin.cu
int launch_render_fwd(const FsRasterFwdArgs& a, cudaStream_t s);     // raster_render.cu
int launch_render_bwd(const FsRasterBwdArgs& a, cudaStream_t s);     // raster_render.cu
int launch_preprocess_bwd(const FsRasterBwdArgs& a, cudaStream_t s); // raster_pre.cu
int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* vis, cudaStream_t s);

int launch_camera_records(int V, const float* ext, const float* K, const float* near, const float* far, const float* bg,
                          int scale_invariant, float* views, cudaStream_t s);   // raster_pre.cu
int launch_cost_volume_fwd(const FsCostVolumeArgs& a, cudaStream_t s);  // cost_volume.cu
int launch_cost_volume_bwd(const FsCostVolumeArgs& a, cudaStream_t s);  // cost_volume.cu

int launch_ptf_view_setup(int V, int H, int W, const float* ext, const float* K, float* E_inv, float* K_px, cudaStream_t s);  // ptf.cu
int launch_ptf_match(const FsPtfArgs& a, cudaStream_t s);   // ptf.cu
int launch_ptf_merge(const FsPtfArgs& a, cudaStream_t s);   // ptf.cu
int launch_ptf_merge_bwd(const FsPtfMergeBwdArgs& a, cudaStream_t s);   // ptf.cu
int launch_ptf_pool_update(const FsPtfArgs& a, cudaStream_t s);   // ptf.cu
int launch_ptf_pool_order(int n_upper, const int* counts, const int* phys_in, const uint8_t* match, int* blk, int* phys_out, cudaStream_t s);
int launch_ptf_pool_gather(int n_upper, const int* n_dev, const int* phys, int F, const float* feats, const float* coords, const float* dens,
                           const float* wemb, const float* ext, const float* depth, float* o_feats, float* o_coords, float* o_dens,
                           float* o_wemb, float* o_ext, float* o_depth, cudaStream_t s);
int launch_gaussian_head(const FsAdapterArgs& a, cudaStream_t s);   // adapter.cu
int launch_backproject(const FsBackprojectArgs& a, cudaStream_t s);   // adapter.cu
int launch_backproject_bwd(const FsBackprojectArgs& a, const float* g_means, float* d_depth, cudaStream_t s);   // adapter.cu
int launch_ply_vertices(const FsPlyArgs& a, cudaStream_t s);          // adapter.cu
int launch_image_u8(int B, int C, int H, int W, const float* img, uint8_t* out, cudaStream_t s);   // adapter.cu
int launch_depth_head(const FsDepthHeadArgs& a, cudaStream_t s);    // depth_head.cu
int launch_depth_head_bwd(const FsDepthHeadBwdArgs& a, cudaStream_t s);   // depth_head.cu
int launch_gaussian_head_bwd(const FsAdapterBwdArgs& a, cudaStream_t s);  // adapter.cu
int launch_ptf_gru_tc(const FsPtfGruArgs& a, cudaStream_t s);
int launch_ptf_gru_bwd_data(const FsGruBwdDataArgs& a, cudaStream_t s);
int launch_ptf_gru_bwd_weights(const FsGruBwdWeightsArgs& a, cudaStream_t s);
size_t ptf_gru_wscratch_bytes();
int launch_ptf_gru_inputs(int M, int F, const int* pj, const int* pp, const float* feats, const float* dens, const float* wemb,
                          const float* v_feats, const float* v_dens, const float* v_wemb, float* A1, cudaStream_t s);
int launch_ptf_gru_update(int M, int F, const float* A1, const float* r_lin, float* U, cudaStream_t s);
int launch_ptf_gru_output(int M, int F, const float* A1, const float* z_lin, const float* q_lin, float* out, cudaStream_t s);

int launch_ptf_gru_output_bwd(int M, int F, const float* A1, const float* z_lin, const float* q_lin, const float* g_out, float* dz_lin,
                              float* dq_lin, float* dA1, cudaStream_t s);
int launch_ptf_gru_update_bwd(int M, int F, const float* A1, const float* r_lin, const float* dU, float* dr_lin, float* dA1, cudaStream_t s);
int launch_ptf_gru_inputs_bwd(int M, int F, const int* pj, const int* pp, const float* dens, const float* wemb, const float* v_dens,
                              const float* v_wemb, const float* dA1, float* d_feats, float* d_dens, float* d_wemb, float* dv_feats,
                              float* dv_dens, float* dv_wemb, cudaStream_t s);

__host__ __device__ inline int tiles_x(int W) { return (W + FS_TILE - 1) / FS_TILE; }
__host__ __device__ inline int tiles_y(int H) { return (H + FS_TILE - 1) / FS_TILE; }

}  // namespace fs
