/*
 * freesplat_b200.h -- C ABI of libfreesplat_b200.so (sm_100a).
 *
 * Drop-in boundary for FreeSplat's data-parallel hot path.
 * Plain pointers and sizes only: every pointer below is a DEVICE pointer unless
 * its name ends in _host; `stream` is a cudaStream_t passed as void*.  The
 * library allocates nothing on the hot path: the caller (PyTorch's caching
 * allocator in freesplat_b200/*.py) owns all buffers.  Every entry point returns
 * 0 on success or a negative FsStatus; no C++ exception crosses the boundary.
 *
 * What each entry point replaces in the reference:
 *   fs_raster_forward / fs_raster_backward / fs_mark_visible
 *       -> the pybind `_C.rasterize_gaussians`, `_C.rasterize_gaussians_backward`,
 *          `_C.mark_visible` of the third-party module imported at
 *          /root/reference/src/model/decoder/cuda_splatting.py:5-8 and called at
 *          :114-127 (module source un-vendored, requirements.txt:17).
 *   fs_cost_volume_forward / fs_cost_volume_backward
 *       -> AVGFeatureVolumeManager.build_cost_volume,
 *          /root/reference/src/model/encoder/modules/cost_volume.py:429-619
 *          (called at src/model/encoder/encoder_freesplat.py:280-288).
 *   fs_ptf_*  -> the index/merge part of EncoderFreeSplat.fuse_gaussians,
 *          /root/reference/src/model/encoder/encoder_freesplat.py:431-522.
 */
#ifndef FREESPLAT_B200_H_
#define FREESPLAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS_ABI_VERSION 4
#define FS_TILE 16            /* BLOCK_X == BLOCK_Y == 16 (upstream config.h) */
#define FS_VIEW_FLOATS 48     /* floats per FsView record                     */
#define FS_REC_FLOATS 12      /* floats per projected-Gaussian record         */

typedef enum FsStatus {
  FS_OK = 0,
  FS_ERR_INVALID_ARG = -1,
  FS_ERR_CUDA = -2,        /* a CUDA runtime call failed: see fs_last_error() */
  FS_ERR_UNSUPPORTED = -3
} FsStatus;

/* Per-view camera record, FS_VIEW_FLOATS floats, device memory.
 *  [0:16)  viewmatrix  (world->camera, flat, TRANSPOSED as the reference passes it,
 *                       cuda_splatting.py:85-87:  m[c*4+r] = M[r][c])
 *  [16:32) projmatrix  (full projection P*V, same flat convention)
 *  [32:35) campos      [35:38) bg colour
 *  [38] tanfovx  [39] tanfovy
 *  [40] scene_scale: means are multiplied by it and covariances by its square before
 *       projection (the `scale_invariant` rescale of cuda_splatting.py:64-71 folded into
 *       preprocess; 1.0 for the plain drop-in op)
 *  [41:48) reserved (0)                                                            */

/* Projected-Gaussian record written by preprocess, FS_REC_FLOATS floats
 * (three float4 so the tile kernels fetch it with 3 coalesced 16-byte loads):
 *  [0] x  [1] y (pixel)   [2] conic.x  [3] conic.y
 *  [4] conic.z  [5] opacity  [6] r  [7] g
 *  [8] b  [9] depth (p_view.z)  [10] hx  [11] hy
 * hx,hy: conservative half-extent (pixels) of the region where alpha >= 1/255 can
 * hold; used only to skip work that the reference's per-pixel test would reject. */

/* ------------------------------------------------------------------ raster */
typedef struct FsRasterFwdArgs {
  /* sizes */
  int32_t P;             /* Gaussians                                            */
  int32_t V;             /* views rendered by this call (1 for the drop-in op)   */
  int32_t H, W;          /* image size                                           */
  int32_t sh_degree;     /* active SH degree D (0..3)                            */
  int32_t M;             /* SH coefficients per channel in `shs` (0 if colours)  */
  float scale_modifier;
  int32_t prefiltered;   /* accepted for API parity; unused (as upstream)        */
  int32_t stages;        /* bit mask of FS_STAGE_*; 0 means all.  Lets the caller put
                            CUDA events between stages (bench.py roofline timing)   */
  int32_t sh_layout;     /* 0: shs is [P,M,3] (upstream op) ; 1: [P,3,M] (the reference's
                            Gaussians.harmonics, read in place: no transpose copy)       */
  int32_t cov_stride;    /* 6: cov3D_precomp is [P,6] ; 9: full row-major [P,3,3]         */
  int64_t capacity;      /* tile-instance capacity of keybuf / point_list        */
  /* inputs */
  const float* means3D;        /* [P,3]                                          */
  const float* shs;            /* [P,M,3] or NULL                                */
  const float* colors_precomp; /* [P,3]  or NULL                                 */
  const float* opacities;      /* [P]                                            */
  const float* scales;         /* [P,3]  or NULL                                 */
  const float* rotations;      /* [P,4]  or NULL  (r,x,y,z)                      */
  const float* cov3D_precomp;  /* [P,6]  or NULL  (xx,xy,xz,yy,yz,zz)            */
  const float* views;          /* [V,FS_VIEW_FLOATS]                             */
  /* outputs */
  float* out_color;      /* [V,3,H,W]                                            */
  float* out_depth;      /* [V,H,W]   sum_j z_j alpha_j T_j (un-normalised)      */
  float* final_T;        /* [V,H,W]   transmittance (4th return = 1-final_T)     */
  uint32_t* n_contrib;   /* [V,H,W]                                              */
  int32_t* radii;        /* [V,P]                                                */
  /* state kept for backward / parity checks */
  float* rec;            /* [V,P,FS_REC_FLOATS]                                  */
  float* cov3D;          /* [V,P,6] covariance actually used (scaled); NULL = not stored
                            (backward recomputes it; only the parity tests ask for it) */
  uint32_t* tiles_touched; /* [V,P] or NULL (parity tests only)                  */
  uint8_t* clamped;      /* [V,P] bit c set <=> SH colour channel c clamped at 0 */
  uint32_t* tile_count;  /* [V*tiles]   scratch (zeroed by the call; after binning: the heaviest-first tile order) */
  uint32_t* tile_cursor; /* [V*tiles]   scratch (zeroed by the call)             */
  uint32_t* ranges;      /* [V*tiles,2] absolute [start,end) into point_list     */
  uint64_t* keybuf;      /* [capacity]  (depth_bits<<32 | gaussian) per instance,
                            sorted ascending inside each tile range on return    */
  uint32_t* point_list;  /* [capacity]  Gaussian index per sorted instance       */
  uint32_t* status;      /* [4] {R_lo, R_hi, overflow(R>capacity), bin fallback taken (direct binning; else 0)}.  If tile_count, tile_cursor and status are
                            ONE buffer [counters | cursors | status] the tile scan runs inside the preprocess kernel
                            (its last CTA; status[3] is the ticket): one launch fewer per step                         */
  uint64_t* bins;        /* optional [V*tiles*bin_cap] scratch (NULL = off): DIRECT BINNING.  preprocess appends every instance key
                            to the fixed-capacity bin of its tile while it counts (the counting atomic returns the slot), the render
                            kernel sorts its tile straight out of the bin: the scatter pass (a second walk over all Gaussians with an
                            atomic round trip per tile) does not run.  If any tile holds more than bin_cap instances, a flag word
                            (tile_cursor[V*tiles], zeroed with the counters) is raised and the call falls back, on the device and
                            inside the same launch sequence, to scan + scatter: results are identical either way.  Requires
                            tile_cursor == tile_count + V*tiles with one spare word behind it, a separate `status`, and
                            bin_cap <= 4096.                                                                              */
  int32_t bin_cap;       /* keys per tile bin                                                                             */
} FsRasterFwdArgs;

#define FS_STAGE_PREPROCESS 1  /* per-Gaussian projection + tile counting            */
#define FS_STAGE_BINNING 2     /* tile scan (+ tile order), scatter                  */
#define FS_STAGE_RENDER 4      /* per-tile depth sort + alpha blend (one kernel)     */
#define FS_STAGE_RENDER_PACKED 8 /* modifier: render with render_fwd2_kernel (two pixels per lane, FFMA2 / FMUL2 packed fp32,
                                    half-warp-independent instance lists); bit-identical results, see DESIGN.md 4           */

typedef struct FsRasterBwdArgs {
  int32_t P, V, H, W, sh_degree, M;
  float scale_modifier;
  int32_t has_depth_grad;      /* 0: dL_ddepth ignored (reference behaviour)     */
  int32_t sh_layout;           /* as in FsRasterFwdArgs; dL_dshs uses the same layout           */
  int32_t cov_stride;          /* 6 or 9; dL_dcov3D uses the same stride                        */
  const float* means3D; const float* shs; const float* colors_precomp;
  const float* opacities; const float* scales; const float* rotations;
  const float* cov3D_precomp;  /* exactly one of cov3D_precomp / (scales, rotations), as in forward */
  const float* views;
  /* forward state */
  const float* rec; const int32_t* radii; const uint8_t* clamped;
  const uint32_t* ranges; const uint32_t* point_list;
  const float* final_T; const uint32_t* n_contrib; const uint32_t* status;
  /* upstream gradients */
  const float* dL_dcolor;      /* [V,3,H,W]                                      */
  const float* dL_ddepth;      /* [V,H,W] or NULL                                */
  const float* dL_dalpha;      /* [V,H,W] or NULL: gradient w.r.t. the 4th return value (1 - final_T) */
  /* scratch: per-(view,Gaussian) screen-space gradients, zeroed by the call     */
  float* dL_dscreen;           /* [V,P,12]: mean2D.xy, conic.xyw, opacity, rgb, depth, pad */
  /* outputs (summed over the V views)                                           */
  float* dL_dmeans2D;          /* [V,P,3]  (x,y,0) per view, as upstream returns  */
  float* dL_dmeans3D;          /* [P,3]                                          */
  float* dL_dcov3D;            /* [P,6]  or NULL when scales/rotations are used  */
  float* dL_dshs;              /* [P,M,3] or NULL                                */
  float* dL_dcolors;           /* [P,3]  or NULL                                 */
  float* dL_dopacities;        /* [P]                                            */
  float* dL_dscales;           /* [P,3]  or NULL                                 */
  float* dL_drotations;        /* [P,4]  or NULL                                 */
  /* Fused reduce-scatter over NVLink (optional, multi-GPU training: every rank back-propagates ITS target views into the
   * replicated Gaussian set).  With peer_delta != NULL the per-Gaussian sums dL_dmeans3D / dL_dcov3D / dL_dshs /
   * dL_dopacities are not stored locally but ADDED (red.global.add.f32 on peer-mapped memory) into the copy of these
   * buffers that lives on rank owner(i) = min(i / shard_rows, world - 1).  peer_delta[r] = byte distance from this rank's
   * buffers to rank r's (one symmetric allocation with the same layout on every rank; 0 for r = own rank).  The buffers
   * must be zero on every rank before the first rank launches, and a cross-rank barrier must follow before they are read:
   * the transfer then overlaps the kernel instead of following it as a separate all-reduce.                              */
  const int64_t* peer_delta;   /* [world] device array, or NULL (plain local stores)                                   */
  int32_t shard_rows;          /* Gaussians per owner rank                                                              */
  int32_t world;
} FsRasterBwdArgs;

/* ------------------------------------------------------------- cost volume */
/* Fused plane-sweep feature volume (cost_volume.py:429-619).  C must be 48 and the MLP
 * 49 -> 32 -> 32 -> 1 (cost_volume.py:423-426, encoder_freesplat.py:157-160).               */
typedef struct FsCostVolumeArgs {
  int32_t B;               /* reference views (b*v of the encoder)                           */
  int32_t K;               /* source views per reference view (<= 16)                        */
  int32_t C, H, W;         /* matching channels (48) and feature-map size                    */
  int32_t D;               /* depth planes                                                   */
  const float* cur_feats;  /* [B,C,H,W]                                                      */
  const float* src_feats;  /* [B,K,C,H,W]                                                    */
  const float* proj;       /* [B,K,3,4]  rows 0..2 of src_Ks @ src_extrinsics (geometry_utils.py:78) */
  const float* cur_invK;   /* [B,3,3]    upper-left block of cur_invK                        */
  const float* planes;     /* [D]        plane depths (cost_volume.py:98-134)                */
  const float* mlp;        /* packed nn.Linear weights: W0[32,49] b0[32] W1[32,32] b1[32] W2[1,32] b2[1] */
  float* out;              /* [B,D,H,W]                                                      */
  /* backward only */
  const float* dL_dout;    /* [B,D,H,W]                                                      */
  float* dL_dcur;          /* [B,C,H,W]   (written)                                          */
  float* dL_dsrc;          /* [B,K,C,H,W] (zeroed by the call, then accumulated)             */
  float* dL_dmlp;          /* packed like `mlp` (zeroed by the call, then accumulated)       */
  float* src_packed;       /* scratch [B,K,C/4,H*W,4]: channel-packed copy of src_feats written by the call
                              (one 16-byte load per tap and 4 channels in the gather)         */
  float* dsrc_packed;      /* backward scratch [B,K,C/4,H*W,4]: dL_dsrc accumulated with 16-byte vector reductions,
                              unpacked to NCHW at the end of the call                           */
  int32_t mlp_mode;        /* 0 = tcgen05 tensor cores (3xTF32, fp32-accurate; forward software-pipelined over the
                              planes), 1 = fp32 CUDA cores (validation of mode 0); forward only: 2 = tensor cores with
                              strictly sequential planes (the round-1 kernel), 3 = mode 0 + L1 prefetch of the next
                              plane's tap lines (measured slower; kept for the record).  Modes 0/2/3 are bit-identical. */
} FsCostVolumeArgs;

int fs_cost_volume_forward(const FsCostVolumeArgs* args, void* stream);
int fs_cost_volume_backward(const FsCostVolumeArgs* args, void* stream);

/* --------------------------------------------------------------------- PTF */
/* One fold step of EncoderFreeSplat.fuse_gaussians (encoder_freesplat.py:443-519): global state of
 * N Gaussians (SoA) + view i  ->  new state.  fs_ptf_match projects / z-buffers / matches and leaves
 * the matched pairs (pair_j, pair_p) and the counters on the device; the caller runs the GRU
 * (networks.py:188-214, plain GEMMs) on those pairs and then calls fs_ptf_merge, which writes the new
 * state in the reference's order: [kept] ++ [fused] ++ [unmatched pixels of view i].            */
typedef struct FsPtfArgs {
  int32_t H, W;            /* image size of view i (pixel grid of the candidates)            */
  int32_t F;               /* latent feature width (64)                                      */
  int32_t n_upper;         /* host-side upper bound of N (grid sizing); N itself is counts_in[0] */
  float depth_thres;       /* 0.1 (encoder_freesplat.py:432)                                 */
  /* current global state, capacity >= n_upper */
  const float* feats;      /* [N,F]  */
  const float* coords;     /* [N,3]  */
  const float* dens;       /* [N]    */
  const float* wemb;       /* [N]    */
  const float* ext;        /* [N,16] per-Gaussian (averaged) camera-to-world matrix          */
  const float* depth;      /* [N]    */
  const int32_t* counts_in;/* [>=1]  counts_in[0] = N                                        */
  /* view i */
  const float* v_feats;    /* [HW,F] */
  const float* v_coords;   /* [HW,3] */
  const float* v_dens;     /* [HW]   */
  const float* v_wemb;     /* [HW]   */
  const float* v_depth;    /* [HW]   predicted depth map of view i                           */
  const float* v_ext;      /* [16]   camera-to-world of view i                               */
  const float* E_inv;      /* [16]   its inverse (row-major)                                 */
  const float* K_px;       /* [9]    pixel-space intrinsics of view i                        */
  /* scratch */
  uint32_t* zbuf;          /* [HW]   */
  int32_t* pix;            /* [n_upper] */
  float* zeta;             /* [n_upper] */
  uint8_t* match;          /* [n_upper] */
  uint8_t* append;         /* [HW]   */
  int32_t* block_counts;   /* [3*ceil(max(n_upper,HW)/1024)] */
  int32_t* pair_j;         /* [n_upper] matched global index, ascending                      */
  int32_t* pair_p;         /* [n_upper] partner pixel of view i                              */
  int32_t* counts_out;     /* [8] {N_in, n_keep, n_match, n_append, N_out, ...}              */
  /* merge inputs / outputs */
  const float* gru_out;    /* [n_match,F] fused features of the matched pairs                */
  float* o_feats; float* o_coords; float* o_dens; float* o_wemb; float* o_ext; float* o_depth;
  /* optional (training): where every input row went.  map_old[j] = output row of global Gaussian j (kept or fused),
   * map_px[p] = output row of appended pixel p or -1.  NULL = not written.                                          */
  int32_t* map_old;        /* [n_upper] */
  int32_t* map_px;         /* [HW]      */
} FsPtfArgs;

int fs_ptf_match(const FsPtfArgs* args, void* stream);
int fs_ptf_merge(const FsPtfArgs* args, void* stream);

/* Append-only pool (inference fold): instead of fs_ptf_merge rewriting the whole state every step, the state arrays of `args`
 * (feats ... depth, capacity >= N + H*W) are updated IN PLACE: fused Gaussians where they are (pair m: row pair_j[m]), the unmatched
 * pixels of view i appended at rows N, N+1, ... (raster order).  The reference's order contract ([kept] ++ [fused] ++ [appended]
 * per step) is tracked by an index: fs_ptf_pool_order turns phys_in (logical position -> pool row of the previous step; NULL =
 * identity) into phys_out by a stable partition with the step's `match` flags (indexed by pool row) and the step's counters
 * `counts` = fs_ptf_match's counts_out; fs_ptf_pool_gather materialises the state in logical order once, at the end.        */
int fs_ptf_pool_update(const FsPtfArgs* args, void* stream);
int fs_ptf_pool_order(int32_t n_upper, const int32_t* counts, const int32_t* phys_in, const uint8_t* match, int32_t* block_scratch,
                      int32_t* phys_out, void* stream);
int fs_ptf_pool_gather(int32_t n_upper, const int32_t* n_dev, const int32_t* phys, int32_t F, const float* feats, const float* coords,
                       const float* dens, const float* wemb, const float* ext, const float* depth, float* o_feats, float* o_coords,
                       float* o_dens, float* o_wemb, float* o_ext, float* o_depth, void* stream);

/* Backward of fs_ptf_merge (training; the reference gets it from autograd through its index / cat ops,
 * encoder_freesplat.py:492-519): gradients of the new state -> gradients of the old state, of view i's candidates and of
 * the GRU output rows.  Kept / appended rows are copies; a fused row j (partner pixel p) splits by the density weights
 * r0 = w0/(w0+w1), r1 = w1/(w0+w1), and the densities receive the derivative of the weighted means.  Several globals may
 * share a partner pixel (z-buffer ties): the view-side scalars accumulate with atomics (zeroed by the call).
 * Any g_* input may be NULL (= zero).                                                                              */
typedef struct FsPtfMergeBwdArgs {
  int32_t H, W, F;
  int32_t N;               /* old state size */
  int32_t n_keep, n_match; /* counters of the step (counts_out[1], [2])                                             */
  /* forward inputs */
  const float* coords; const float* dens; const float* ext; const float* depth;       /* old state [N,...]          */
  const float* v_coords; const float* v_dens; const float* v_depth; const float* v_ext;/* view i                    */
  const uint8_t* match;    /* [N]  */
  const int32_t* pix;      /* [N]  partner pixel of a matched global                                                */
  const int32_t* map_old;  /* [N]  */
  const int32_t* map_px;   /* [HW] */
  /* upstream gradients w.r.t. the NEW state [N_out, ...] */
  const float* g_feats; const float* g_coords; const float* g_dens; const float* g_wemb; const float* g_ext; const float* g_depth;
  /* outputs */
  float* d_feats;  float* d_coords;  float* d_dens;  float* d_wemb;  float* d_ext;  float* d_depth;     /* [N, ...]  */
  float* dv_feats; float* dv_coords; float* dv_dens; float* dv_wemb; float* dv_depth;                   /* [HW, ...] */
  float* d_gru;            /* [n_match, F] gradient of the GRU output rows (NULL if n_match == 0)                   */
} FsPtfMergeBwdArgs;
int fs_ptf_merge_backward(const FsPtfMergeBwdArgs* args, void* stream);
/* Per-view constants of the fold for all V views in one launch: E_inv[v] = extrinsics[v]^-1 (encoder_freesplat.py:454;
 * canonical arithmetic: fp64 cofactor expansion, one rounding to fp32) and the pixel-space intrinsics K_px[v]
 * (rows 0 / 1 of the normalised K times W / H, :445-447).  extrinsics [V,16], intrinsics [V,9] -> E_inv [V,16], K_px [V,9]. */
int fs_ptf_view_setup(int32_t V, int32_t H, int32_t W, const float* extrinsics, const float* intrinsics, float* E_inv,
                      float* K_px, void* stream);
/* Element-wise glue of the GRU (networks.py:201-214) for the M matched pairs; the Linear layers in between are
 * plain GEMMs run by the caller (cuBLAS).  A1 [M,2F+48] = [hidden | PE(v_dens,wemb) | input | PE(dens,v_wemb)],
 * U [M,2F+24] = [sigmoid(r_lin)*hidden | input | PE(dens,v_wemb)], out [M,F] = (1-z)*hidden + z*tanh(q_lin).          */
int fs_ptf_gru_inputs(int32_t M, int32_t F, const int32_t* pair_j, const int32_t* pair_p, const float* feats, const float* dens,
                      const float* wemb, const float* v_feats, const float* v_dens, const float* v_wemb, float* A1, void* stream);
int fs_ptf_gru_update(int32_t M, int32_t F, const float* A1, const float* r_lin, float* U, void* stream);
int fs_ptf_gru_output(int32_t M, int32_t F, const float* A1, const float* z_lin, const float* q_lin, float* out, void* stream);

/* Backward of the three glue steps (training path: forward = fs_ptf_gru on the tensor cores, backward = recompute with
 * these kernels around the caller's GEMMs).  dA1 [M,2F+48] is built up across the calls: output_backward writes its [:, :F]
 * (direct path to the hidden latent), update_backward adds to [:, :F] and writes the rest, the caller adds the first layers'
 * products, inputs_backward scatters it: d_feats / d_dens / d_wemb rows of the matched globals are plain stores (the caller
 * zero-fills the buffers), dv_* of view i accumulate with atomics.                                                      */
int fs_ptf_gru_output_backward(int32_t M, int32_t F, const float* A1, const float* z_lin, const float* q_lin, const float* g_out, float* dz_lin,
                               float* dq_lin, float* dA1, void* stream);
int fs_ptf_gru_update_backward(int32_t M, int32_t F, const float* A1, const float* r_lin, const float* dU, float* dr_lin, float* dA1, void* stream);
int fs_ptf_gru_inputs_backward(int32_t M, int32_t F, const int32_t* pair_j, const int32_t* pair_p, const float* dens, const float* wemb,
                               const float* v_dens, const float* v_wemb, const float* dA1, float* d_feats, float* d_dens, float* d_wemb,
                               float* dv_feats, float* dv_dens, float* dv_wemb, void* stream);

/* The whole GRU (networks.py:188-214, latent width 64, 24-wide weight embeddings) for the M matched pairs on the
 * tensor cores (tcgen05, 3xTF32): gathers, positional encodings, the six Linear layers and the gates in one kernel.
 * W_* are the nn.Linear weights ([out,in] row-major): mlp_r[0] [64,176], mlp_z[0] [64,176], mlp_r[2] / mlp_z[2] [64,64],
 * mlp_n[0] [64,152], mlp_n[2] [64,64]; biases = [b_r0 | b_z0 | b_r2 | b_z2 | b_n0 | b_n2] (6 x 64);
 * wscratch: fs_ptf_gru_wscratch_bytes() bytes of device scratch (tf32-split, pre-tiled weights).                */
typedef struct FsPtfGruArgs {
  int32_t M;                      /* number of pairs, or an UPPER BOUND of it when M_dev is given (sizes the grid)            */
  int32_t flags;                  /* bit 0: wscratch already holds the prepared weights (skip the preparation kernel)       */
  const int32_t* pair_j; const int32_t* pair_p;
  const float* feats; const float* dens; const float* wemb;          /* global state  */
  const float* v_feats; const float* v_dens; const float* v_wemb;    /* view i        */
  const float* W_r0; const float* W_z0; const float* W_r2; const float* W_z2; const float* W_n0; const float* W_n2;
  const float* biases;
  unsigned char* wscratch;
  float* out;                                                         /* [M,64]        */
  const int32_t* M_dev;           /* optional: the pair count on the DEVICE (fs_ptf_match's counts_out[2]): no host read   */
  float* save;                    /* optional (training; M must then be exact): six [M,64] matrices [Hr | Hz | r_lin | z_lin | Hn |
                                     q_lin], the post-ReLU hidden layers and pre-gate outputs the backward consumes            */
  float* save_a1;                 /* optional (training): the first-layer input [M,176] = [h | e_h | x | e_in]                  */
} FsPtfGruArgs;
int fs_ptf_gru(const FsPtfGruArgs* args, void* stream);
int64_t fs_ptf_gru_wscratch_bytes(void);

/* The matrix products of the GRU's backward pass (training; the reference: autograd through networks.py:201-214 inside the fold
 * of encoder_freesplat.py:485-506) on the tensor cores, 3xTF32 like fs_ptf_gru.  Everything element-wise around them is
 * fs_ptf_gru_{inputs,update,output_backward,update_backward,inputs_backward}.
 *
 * fs_ptf_gru_bwd_data:    C[M,N] (op)= A[M,64] . W[64,N]   (W = an nn.Linear weight [out = 64, in = N], row-major, ld = N)
 *   mode 0: C = P;   mode 1 (N == 64): C = P where mask > 0, else 0 (mask [M,64] = the post-ReLU activation of the layer below);
 *   mode 2: C += P (16-byte reductions; C must not be written by anything else meanwhile);
 *   mode 3 (N == 152, the product dU = dHn . W_n0): the update-gate chain rule of U = [sigmoid(r_lin) h | x | e_in] in the
 *           epilogue instead of a stored dU: dr_lin = P[:, :64] h r (1 - r), C[:, :64] += P[:, :64] r, C[:, 64:88] = 0,
 *           C[:, 88:176] = P[:, 64:152]  (C = dA1 [M,176], h = the first 64 columns of A1, r = sigmoid(r_lin)).
 *   N % 4 == 0, N <= 192; lda, ldc, ldm, ldh % 4 == 0 (16-byte aligned rows).                                            */
typedef struct FsGruBwdDataArgs {
  int32_t M, N, mode;
  int32_t lda, ldc, ldm;
  const float* A; const float* W; const float* mask;
  float* C;
  const float* h; const float* r_lin; float* dr_lin;      /* mode 3 only: A1 (ldh), r_lin [M,64], dr_lin [M,64] */
  int32_t ldh, reserved;
} FsGruBwdDataArgs;
int fs_ptf_gru_bwd_data(const FsGruBwdDataArgs* args, void* stream);
/* fs_ptf_gru_bwd_weights: G[128,ldg] = [Y0 | Y1]^T . [X0 | X1 | 1]  summed over the M pairs (G is overwritten).
 *   Y0, Y1: [M,64] output gradients (Y1 may be NULL: rows 64..127 of G stay zero); X0 [M,nx0], X1 [M,nx1] (may be NULL) the
 *   layer inputs (nx0 + nx1 <= 224); column nx0 + nx1 of G is the sum of the Y rows (the bias gradients).  ldg % 16 == 0,
 *   nx0 + nx1 < ldg <= 256.
 *   Weight gradients of two layers come out of one call as the blocks G[0:64, 0:nx0] and G[64:128, nx0:nx0+nx1] (or both row
 *   blocks against X0 when the layers share their input).  Summation order over CTAs is not fixed (fp32 atomics).           */
typedef struct FsGruBwdWeightsArgs {
  int32_t M, ldy0, ldy1, ldx0, nx0, ldx1, nx1, ldg;
  const float* Y0; const float* Y1; const float* X0; const float* X1;
  float* G;
} FsGruBwdWeightsArgs;
int fs_ptf_gru_bwd_weights(const FsGruBwdWeightsArgs* args, void* stream);

/* ------------------------------------------------------------ Gaussian head */
/* GaussianAdapter.forward(fusion=False, coords=...) (gaussian_adapter.py:136-200) as one kernel.               */
typedef struct FsAdapterArgs {
  int32_t N, H, W, sh_degree;
  float scale_min, scale_max, eps;
  const float* raw;        /* [N, 7+3*d_sh]: scales(3) | quaternion xyzw (4) | SH (3 x d_sh)          */
  const float* depths;     /* [N]                                                                    */
  const float* opacities;  /* [N]                                                                    */
  const float* coords;     /* [N,3]                                                                  */
  const float* ext;        /* [N,16] per-Gaussian camera-to-world                                    */
  const float* K;          /* [9] normalised intrinsics                                              */
  float* means;            /* [N,3]  */
  float* covariances;      /* [N,3,3]*/
  float* harmonics;        /* [N,3,d_sh] */
  float* opacities_out;    /* [N]    */
  float* scales;           /* [N,3]  */
  float* rotations;        /* [N,4]  */
} FsAdapterArgs;
int fs_gaussian_head(const FsAdapterArgs* args, void* stream);

/* Backward of fs_gaussian_head (training): upstream gradients of the head's outputs (any may be NULL = zero) -> gradients
 * of raw, depths, opacities, coords and the per-Gaussian c2w matrices (the reference: autograd through
 * gaussian_adapter.py:151-172).                                                                                     */
typedef struct FsAdapterBwdArgs {
  int32_t N, H, W, sh_degree;
  float scale_min, scale_max, eps;
  const float* raw; const float* depths; const float* ext; const float* K;       /* forward inputs                 */
  const float* g_means;      /* [N,3]      */
  const float* g_cov;        /* [N,3,3]    */
  const float* g_harmonics;  /* [N,3,d_sh] */
  const float* g_opacities;  /* [N]        */
  const float* g_scales;     /* [N,3]      */
  const float* g_rotations;  /* [N,4]      */
  float* d_raw;              /* [N,7+3*d_sh] */
  float* d_depths;           /* [N]        */
  float* d_opacities;        /* [N]        */
  float* d_coords;           /* [N,3]      */
  float* d_ext;              /* [N,16] or NULL */
} FsAdapterBwdArgs;
int fs_gaussian_head_backward(const FsAdapterBwdArgs* args, void* stream);

/* ------------------------------------------------------------ .ply vertex table */
/* The per-Gaussian part of export_ply (src/model/ply_export.py:26-92): [N,17] float rows
 * (x y z nx ny nz f_dc_0..2 opacity scale_0..2 rot_0..3) ready to be written after a binary_little_endian header.   */
typedef struct FsPlyArgs {
  int32_t N, d_sh;
  float scale_factor;      /* means.abs().quantile(0.95, dim=0).max() of the median-shifted means     */
  float shift[3];          /* means.median(dim=0).values                                              */
  float R[9];              /* viewer rotation @ inverse camera rotation, row-major (ply_export.py:43-63) */
  const float* means;      /* [N,3]  */
  const float* scales;     /* [N,3]  */
  const float* rotations;  /* [N,4] xyzw */
  const float* harmonics;  /* [N,3,d_sh] */
  const float* opacities;  /* [N]    */
  float* table;            /* [N,17] */
} FsPlyArgs;
int fs_ply_vertices(const FsPlyArgs* args, void* stream);

/* Evaluation image dump (src/misc/image_io.py:36-53 prep_image, used by save_image at model_wrapper.py:382-416): images
 * [B,C,H,W] float (C = 1, 3 or 4) -> uint8 [H, B*W, C'] (C' = 3 for C = 1), uint8(clip(x,0,1)*255), batch side by side.   */
int fs_image_u8(int32_t B, int32_t C, int32_t H, int32_t W, const float* images, uint8_t* out, void* stream);

/* ------------------------------------------------------------ depth back-projection */
/* GaussianAdapter.forward(fusion=True) (gaussian_adapter.py:175-189 -> Create_from_depth_map.project :48-68): world
 * coordinates of every pixel of the V context views from their depth maps, one launch.                          */
typedef struct FsBackprojectArgs {
  int32_t V, H, W, reserved;
  const float* depth;      /* [V,H,W]                                                                 */
  const float* K;          /* [9] NORMALISED intrinsics of view 0 (the reference uses intrinsics[i,0]) */
  const float* c2w;        /* [V,16] camera-to-world                                                  */
  float* means;            /* [V,H*W,3]                                                               */
} FsBackprojectArgs;
int fs_backproject(const FsBackprojectArgs* args, void* stream);
/* Backward w.r.t. the depth maps (training): g_means [V,H*W,3] -> d_depth [V,H,W]; `depth` / `means` of args are unused. */
int fs_backproject_backward(const FsBackprojectArgs* args, const float* g_means, float* d_depth, void* stream);

/* ------------------------------------------------------------ depth-regression head tail */
/* Tail of DepthDecoder.forward (modules/networks.py:130-152) for one scale: softmax over the D planes, expectation of
 * the plane candidates, depth = exp(E) (log_planes) or 1/E; with `upsample` also the x2 bilinear (align_corners=True)
 * outputs of scale 0: depth_pred_s-1 and depth_weights = max_d upsampled planes.  One pass over the logits.     */
typedef struct FsDepthHeadArgs {
  int32_t B, D, h, w;      /* logits [B,D,h,w]                                                       */
  int32_t log_planes;      /* 1: depth = exp(E) (ScanNet / Replica configs), 0: depth = 1/E (RE10K)  */
  int32_t upsample;        /* 1: scale 0 (fused x2 outputs, D <= 128), 0: scales 1..3                */
  int32_t tile_mode;       /* 0: TMA box loads of the logit tiles, 1: LDG staging (validation path)  */
  int32_t reserved;
  const float* logits;
  const float* candi;      /* [D] plane candidates (DepthDecoder.depth_candi_curr)                   */
  float* expect;           /* [B,h,w]   log_depth_pred_s{i}                                          */
  float* depth;            /* [B,h,w]   depth_pred_s{i}                                              */
  float* depth_up;         /* [B,2h,2w] depth_pred_s-1   (upsample only)                             */
  float* weights_up;       /* [B,2h,2w] depth_weights    (upsample only)                             */
} FsDepthHeadArgs;
int fs_depth_head(const FsDepthHeadArgs* args, void* stream);

/* Backward of fs_depth_head (training): gradients of the tail's outputs (any may be NULL = zero) -> gradient of the plane
 * logits.  The reference: autograd through networks.py:130-152 (softmax, expectation, exp / reciprocal, bilinear x2,
 * max over the upsampled planes).  `stats` and `argmax_up` are scratch written by the call.                           */
typedef struct FsDepthHeadBwdArgs {
  int32_t B, D, h, w;
  int32_t log_planes, upsample, reserved0, reserved1;
  const float* logits;       /* [B,D,h,w]                                                         */
  const float* candi;        /* [D]                                                               */
  const float* g_expect;     /* [B,h,w]   d/d log_depth_pred_s{i}                                 */
  const float* g_depth;      /* [B,h,w]   d/d depth_pred_s{i}                                     */
  const float* g_depth_up;   /* [B,2h,2w] d/d depth_pred_s-1   (upsample only)                    */
  const float* g_weights_up; /* [B,2h,2w] d/d depth_weights    (upsample only)                    */
  float* stats;              /* scratch [B,h,w,3]: max, sum exp, expectation                      */
  uint8_t* argmax_up;        /* scratch [B,2h,2w] (needed when g_weights_up is given; D <= 256)   */
  float* d_logits;           /* [B,D,h,w]                                                         */
} FsDepthHeadBwdArgs;
int fs_depth_head_backward(const FsDepthHeadBwdArgs* args, void* stream);

/* ------------------------------------------------------------ CUDA graphs */
/* A launch sequence whose arguments (pointers, sizes) do not change between steps -- e.g. fs_raster_forward on a static
 * workspace -- can be recorded once and replayed with ONE cudaGraphLaunch: no host-side launch gaps between its kernels.
 *   fs_graph_capture_begin(&s);  fs_raster_forward(&args, s); ...;  fs_graph_capture_end(s, &g);   (s is consumed)
 *   fs_graph_launch(g, stream) per step;  fs_graph_destroy(g).
 * (The reference has no counterpart: its op blocks on a D2H read of the instance count between its launches.)          */
int fs_graph_capture_begin(void** stream_out);
int fs_graph_capture_end(void* stream, void** graph_exec_out);
int fs_graph_launch(void* graph_exec, void* stream);
int fs_graph_destroy(void* graph_exec);

int fs_abi_version(void);
/* sizeof() of the argument structs as compiled (0: FsRasterFwdArgs, 1: FsRasterBwdArgs, 2: FsCostVolumeArgs, 3: FsPtfArgs,
 * 4: FsPtfGruArgs, 5: FsAdapterArgs, 6: FsDepthHeadArgs, 7: FsBackprojectArgs, 8: FsPlyArgs, 9: FsPtfMergeBwdArgs, 10: FsAdapterBwdArgs, 11: FsDepthHeadBwdArgs,
 * 12: FsGruBwdDataArgs, 13: FsGruBwdWeightsArgs; -1 otherwise) so that a foreign-language binding can verify its layout.      */
int fs_struct_size(int32_t which);
const char* fs_last_error(void);      /* thread-local, valid until the next call  */
int fs_device_sm_count(void);         /* negative FsStatus on failure             */

int fs_raster_forward(const FsRasterFwdArgs* args, void* stream);
int fs_raster_backward(const FsRasterBwdArgs* args, void* stream);
/* Builds the [V,FS_VIEW_FLOATS] camera records from camera-to-world extrinsics [V,4,4], normalised intrinsics
 * [V,3,3], near/far [V] and background colours [V,3] -- the work of cuda_splatting.py:64-87 (scale-invariant
 * rescale, get_fov, get_projection_matrix, inverse, transposes) in one launch.                       */
int fs_camera_records(int32_t V, const float* extrinsics, const float* intrinsics, const float* near, const float* far,
                      const float* bg, int32_t scale_invariant, float* views, void* stream);
/* visible[i] = (p_view.z > 0.2) for one view (upstream mark_visible / checkFrustum) */
int fs_mark_visible(int32_t P, const float* means3D, const float* view /*FS_VIEW_FLOATS*/,
                    uint8_t* visible, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FREESPLAT_B200_H_ */
