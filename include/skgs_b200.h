/*
 * skgs_b200.h - C ABI of the B200-native SK_GS hot path (libskgs_b200.so).
 *
 * Plain pointers and sizes only (no torch types).  All pointers are DEVICE pointers unless the name ends in `_host`.
 * Every entry point enqueues work on `stream` (a cudaStream_t passed as void*), never synchronises the host,
 * returns 0 on success or a negative skgs_status and records a message retrievable with skgs_last_error().
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   skgs_raster_forward        <- my_ext/_C/src/nerf/gaussian_rasterizer_forward.cu:260-315 `rasterize_gaussians`
 *                                 (RasterizeGaussiansCUDA -> Rasterizer::forward :157-250) and the un-vendored
 *                                 diff_gaussian_rasterization `_C.rasterize_gaussians` behind
 *                                 networks/renderer/gaussian_render_origin.py:36-52
 *   skgs_raster_backward       <- my_ext/_C/src/nerf/gaussian_rasterizer_backwrad.cu:200-261
 *                                 `rasterize_gaussians_backward` (Rasterizer::backward :148-198)
 *   skgs_raster_layout         <- GeometryState/ImageState/BinningState::fromChunk + required<T>()
 *                                 my_ext/_C/src/nerf/gaussian_rasterizer_imp.cu:39-73, include/gaussian_render.h:111-158
 *   skgs_fk_lbs_forward/backward <- networks/sk_gs.py:1069-1150 (`kinematic` post-MLP + `skeleton_warp_SE3` :193-206 +
 *                                 `calc_LBS_weight` :751-774 + LBS blend :1147-1149); there is no native boundary in the
 *                                 reference (lietorch + pytorch3d ops), this is the new one
 *   skgs_assemble_forward/backward <- output assembly networks/sk_gs.py:1162-1163,1192,1202-1203 with the activations
 *                                 of networks/gaussian_splatting.py:155-160
 */
#ifndef SKGS_B200_H_
#define SKGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SKGS_API __attribute__((visibility("default")))
#else
#define SKGS_API
#endif

typedef enum skgs_status {
  SKGS_OK = 0,
  SKGS_ERR_INVALID_ARG = -1,
  SKGS_ERR_CUDA = -2,
  SKGS_ERR_UNSUPPORTED = -3,
  SKGS_ERR_WORKSPACE = -4
} skgs_status;

/* thread-local message of the last failing call on this thread ("" if none) */
SKGS_API const char* skgs_last_error(void);
/* library version, increases whenever the ABI changes */
SKGS_API int skgs_abi_version(void);
/* compute capability this library was built for (100 = sm_100a) */
SKGS_API int skgs_built_for_sm(void);
/* number of kernels launched by this process through the library so far (bench.py's `gpu_launches`) */
SKGS_API uint64_t skgs_launch_count(void);

/* Optional per-kernel device timing: while enabled, every kernel launched through the library is bracketed by a
 * CUDA-event pair on its own stream.  skgs_profile_enable(0|1) also clears the records; skgs_profile_collect
 * synchronises the recorded events and writes one line per kernel name: "<name> <launches> <total microseconds>". */
SKGS_API void skgs_profile_enable(int on);
SKGS_API int skgs_profile_collect(char* buf, size_t cap);

/* ---------------------------------------------------------------------------------------------------------------
 * Rasterizer
 * ------------------------------------------------------------------------------------------------------------- */

/* Mirrors GaussianRasterizationSettings (networks/gaussian_splatting.py:271-284; networks/renderer/gaussian_render.py:34-48).
 * Matrices are the torch row-major tensors viewmatrix = Tw2v.T and projmatrix = (Tv2c @ Tw2v).T (16 floats each). */
typedef struct skgs_raster_settings {
  int32_t image_height;
  int32_t image_width;
  float tanfovx;
  float tanfovy;
  float scale_modifier;
  int32_t sh_degree;        /* active degree D, 0..3 */
  int32_t quat_wxyz;        /* 1: rotations are (w,x,y,z) (upstream boundary B1); 0: (x,y,z,w) (in-tree boundary B2) */
  int32_t prefiltered;      /* accepted for API parity; culled points are simply skipped */
  int32_t debug;            /* bit 0: accepted for API parity; bit 1: stop after binning (tests); bit 2 (backward): recover
                               T_final as 1 - (1 - T) like the reference's in-tree extension does (parity tests only) */
  const float* viewmatrix;  /* device [16] */
  const float* projmatrix;  /* device [16] */
  const float* campos;      /* device [3] */
  const float* bg;          /* device [3] or NULL (= black) */
} skgs_raster_settings;

/* Byte offsets of every array inside the three caller-owned arenas (all 256-byte aligned).
 * geom arena (per Gaussian, written by forward, read by backward):                                   */
typedef struct skgs_raster_layout {
  size_t geom_bytes, binning_bytes, img_bytes;
  /* geom */
  size_t header;         /* skgs_raster_header */
  size_t means2D;        /* float2 [P]   pixel-space centre */
  size_t depths;         /* float  [P]   view-space z */
  size_t cov3D;          /* float  [P][6] */
  size_t conic_opacity;  /* float4 [P]   (conic a, b, c, opacity) */
  size_t rgbd;           /* float4 [P]   (r, g, b, depth) */
  size_t cull;           /* float4 [P]   footprint-culling record of the compositing kernels: (pmin, -B/C, -B/A, -) with
                                         pmin = -log(255 opacity) - margin, the exponent below which alpha < 1/255 */
  size_t clamped;        /* uint8  [P]   bit c set <=> colour channel c was clamped at 0 */
  size_t tiles_touched;  /* uint32 [P] */
  size_t point_offsets;  /* uint32 [P]   inclusive prefix sum of tiles_touched */
  size_t scan_state;     /* uint64 [ceil(P/256)+1] look-back words of the fused scan */
  size_t geom_grads;     /* float  [P][12] packed backward accumulators: mean2D.xy conic.abc opacity depth - rgb -
                                         (zero outside the composite-bwd -> preprocess-bwd window) */
  /* binning (capacity R_cap entries) */
  size_t keys;           /* uint64 [R_cap]  (tile << 32) | depth bits: emission order, after the forward SORTED by
                                            (tile, depth), stable - the list cub::DeviceRadixSort::SortPairs produces in
                                            the reference (gaussian_rasterizer_forward.cu:224-229) */
  size_t vals;           /* uint32 [R_cap]  Gaussian ids, same order: after the forward the reference's point_list */
  size_t tile_pairs;     /* uint64 [R_cap]  scratch: the per-tile segments (depth bits << 32 | Gaussian id) */
  size_t tile_grid;      /* int32 [gy+1][gx+1] x 32: 2-D difference grid of the tile rectangles (one 128-byte line per
                                            cell); its prefix sum is the number of keys per tile */
  size_t tile_cursors;   /* uint32 [tiles]  fill level of every tile segment during the scatter pass */
  /* img */
  size_t ranges;         /* uint2  [tiles]  [start, end) of every screen tile in the sorted list, (0, 0) if empty */
  size_t n_contrib;      /* uint32 [H*W] */
  size_t final_T;        /* float  [H*W] */
  size_t tile_order;     /* uint4  [tiles]  (tile, start, end, -) by decreasing list length: work order of compositing */
  size_t work_counters;  /* uint32 [8]      work tickets: [0] forward / [1] backward compositing, [2..4] per-tile sort */
} skgs_raster_layout;

/* Lives at geom + layout.header; written on the device, never read by the library on the host.
 * After the forward the sorted lists are layout.keys / layout.vals. */
typedef struct skgs_raster_header {
  uint32_t num_rendered;  /* R = sum tiles_touched (may exceed R_cap) */
  uint32_t num_visible;   /* Gaussians with radius > 0 */
  uint32_t scan_ticket;   /* internal */
  uint32_t overflow;      /* 1 if R > R_cap: the image of this call is INVALID, re-run with a larger binning arena */
  uint32_t reserved[28];
} skgs_raster_header;

SKGS_API int skgs_raster_layout_query(int32_t P, int32_t W, int32_t H, int64_t R_cap, skgs_raster_layout* out);

/* Forward: preprocess (+ fused prefix sum + key emission + tile-rectangle counting, ONE kernel) -> tile plan (ranges,
 * work order) -> scatter into per-tile segments -> per-tile sort in shared memory -> per-tile compositing.
 * Exactly one of shs / colors_precomp and one of (scales, rotations) / cov3D_precomp must be non-NULL (same rule as
 * networks/renderer/gaussian_render.py:250-255).
 *   means3D [P][3], shs [P][M][3], colors_precomp [P][3], opacities [P], scales [P][3], rotations [P][4], cov3D_precomp [P][6]
 *   out_color [3][H][W], out_depth [H][W], out_alpha [H][W] (= 1 - T), radii int32 [P]
 *   num_rendered_host: optional PINNED host uint32[4] that receives the first four header words
 *   {R, num_visible, -, overflow} by an async copy enqueued right after the preprocess kernel.
 * P == 0 is legal (image = background). */
SKGS_API int skgs_raster_forward(const skgs_raster_settings* s, int32_t P, int32_t M, const float* means3D,
                                 const float* shs, const float* colors_precomp, const float* opacities,
                                 const float* scales, const float* rotations, const float* cov3D_precomp, void* geom,
                                 void* binning, int64_t R_cap, void* img, float* out_color, float* out_depth,
                                 float* out_alpha, int32_t* radii, uint32_t* num_rendered_host, void* stream);

/* Forward split in two, for callers that look at R between the stages (the reference sizes its binning buffer from a
 * blocking read of R, gaussian_rasterizer_forward.cu:208-213):
 *   _geometry: preprocess + scan (+ async copy of the header words to num_rendered_host).  With a binning arena
 *              (binning != NULL, sized for R_cap entries, plus the img arena) the same kernel also emits the keys;
 *              with binning == NULL nothing is emitted and R_cap / img are ignored.
 *   _render  : [key emission from the stored geometry unless keys_emitted] + sort + ranges + composite, for the R_cap
 *              the binning arena was sized for.  May be re-run on the same geometry with a larger arena
 *              (keys_emitted = 0) after an overflow. */
SKGS_API int skgs_raster_forward_geometry(const skgs_raster_settings* s, int32_t P, int32_t M, const float* means3D,
                                          const float* shs, const float* colors_precomp, const float* opacities,
                                          const float* scales, const float* rotations, const float* cov3D_precomp,
                                          void* geom, int32_t* radii, void* binning, int64_t R_cap, void* img,
                                          uint32_t* num_rendered_host, void* stream);
SKGS_API int skgs_raster_forward_render(const skgs_raster_settings* s, int32_t P, void* geom, void* binning,
                                        int64_t R_cap, int64_t R_hint, void* img, const int32_t* radii,
                                        int32_t keys_emitted, float* out_color, float* out_depth, float* out_alpha,
                                        uint32_t* num_rendered_host, void* stream);

/* Backward.  dL_dcolor [3][H][W] is required; dL_ddepth / dL_dalpha [H][W] may be NULL.
 * Outputs (each may be NULL when the corresponding input was absent): dL_dmeans3D [P][3], dL_dmeans2D [P][3]
 * (screen-space gradient used by densification, z = 0), dL_dsh [P][M][3], dL_dcolors [P][3], dL_dopacity [P],
 * dL_dscales [P][3], dL_drotations [P][4] (same quaternion layout as the input), dL_dcov3D [P][6].
 * All outputs are fully written (zeros for culled Gaussians), no pre-zeroing needed. */
SKGS_API int skgs_raster_backward(const skgs_raster_settings* s, int32_t P, int32_t M, const float* means3D,
                                  const float* shs, const float* colors_precomp, const float* scales,
                                  const float* rotations, const float* cov3D_precomp, const int32_t* radii, void* geom,
                                  const void* binning, int64_t R_cap, const void* img, const float* dL_dcolor,
                                  const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D,
                                  float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors, float* dL_dopacity,
                                  float* dL_dscales, float* dL_drotations, float* dL_dcov3D, void* stream);

/* skgs_raster_backward with the assembly backward (skgs_assemble_backward) fused into its per-Gaussian kernel: the
 * gradients of the assembled Gaussians are chained on the spot to the canonical parameters - dL_dscaling = dL/dscales *
 * exp(_scaling), dL_drotation = normalize-backward at (_rotation + d_rot), dL_dopacity = dL/dopacity * s(1 - s) - and
 * dL/dscales, dL/drotations, dL/dopacity never touch HBM.  SH colours + scale / rotation inputs only (the SK_GS
 * training step); rotations (x,y,z,w).  The LBS backward then takes dL_dd_xyz = dL_dxyz, dL_dd_rot = dL_drotation,
 * dL_dd_scale.  d_rot may be NULL (static stage).  dL_dd_xyz / dL_dd_rot (may be NULL): private copies of dL_dxyz /
 * dL_drotation for a caller whose all-reduce overwrites those two (in an arena) while the LBS backward still needs the
 * local values. */
SKGS_API int skgs_raster_assemble_backward(const skgs_raster_settings* s, int32_t P, int32_t M, const float* means3D,
                                           const float* shs, const float* scales, const float* rotations,
                                           const int32_t* radii, void* geom, const void* binning, int64_t R_cap,
                                           const void* img, const float* dL_dcolor, const float* dL_ddepth,
                                           const float* dL_dalpha, const float* scaling, const float* rotation,
                                           const float* opacity_logit, const float* d_rot, float* dL_dxyz,
                                           float* dL_dmeans2D, float* dL_dsh, float* dL_dscaling, float* dL_drotation,
                                           float* dL_dopacity, float* dL_dd_scale, float* dL_dd_xyz,
                                           float* dL_dd_rot, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Skeleton forward kinematics + linear blend skinning (+ optional output assembly)
 * ------------------------------------------------------------------------------------------------------------- */
typedef enum skgs_lbs_mode {
  SKGS_LBS_W = 0,               /* softmax(gather(sp_W, idx))                      networks/sk_gs.py:767-768 */
  SKGS_LBS_KERNEL = 1,          /* exp(-d2/(2 r^2)) + 1e-7, L1-normalised          :760-766 */
  SKGS_LBS_WEIGHTED_KERNEL = 2, /* ... * sigmoid(weight)                           :763-764 */
  SKGS_LBS_DIST = 3             /* softmax(-d2 / temperature)                      :769-770 */
} skgs_lbs_mode;

typedef struct skgs_skeleton {
  int32_t M;               /* joints, 1..1024 */
  int32_t L;               /* levels of the binary-lifting table */
  int32_t root;
  int32_t K;               /* nearest joints per Gaussian, 1..8, <= M */
  int32_t mode;            /* skgs_lbs_mode */
  float temperature;       /* SKGS_LBS_DIST only */
  const float* joints;     /* [M][3] */
  const float* sk_r;       /* [M][4] xyzw local joint rotation (normalised inside, like lietorch) */
  const float* sk_r_delta; /* NULL, or [M][3] axis-angle / [M][4] quaternion repose delta (networks/sk_gs.py:1087-1088) */
  int32_t sk_r_delta_dim;  /* 3 or 4 */
  const float* sk_d_rot;   /* [M][4] */
  const float* sk_d_scale; /* [M][3] */
  const float* g_tr;       /* [7] (t, q xyzw) global transform or NULL (identity) */
  const int32_t* parents;  /* [M][L], parents[:, l] = 2^l-th ancestor, parents[root, :] = root */
  const float* sp_W;       /* [P][M]   (mode W) */
  const float* sp_radius;  /* [M] log-radius (kernel modes) */
  const float* sp_weight;  /* [M] logit (weighted_kernel) */
} skgs_skeleton;

/* Forward.  xyz [P][3] (treated as constant: the reference detaches it, networks/sk_gs.py:1113).
 * Outputs: d_xyz [P][3], d_rot [P][4], d_scale [P][3], sk_T [M][7] (t, q xyzw), weights [P][K], indices int64 [P][K].
 * workspace: device scratch of skgs_fk_lbs_workspace_bytes(M) bytes (the joint table: forward kinematics run ONCE, in
 * a one-CTA kernel, the per-Gaussian kernel copies the table into shared memory). */
SKGS_API size_t skgs_fk_lbs_workspace_bytes(int32_t M);
SKGS_API int skgs_fk_lbs_forward(const skgs_skeleton* sk, int32_t P, const float* xyz, float* d_xyz, float* d_rot,
                                 float* d_scale, float* sk_T, float* weights, int64_t* indices, void* workspace,
                                 void* stream);

/* Backward.  Incoming: dL_dd_xyz [P][3], dL_dd_rot [P][4], dL_dd_scale [P][3] (any may be NULL = zero), plus optional
 * direct gradients on the auxiliary outputs dL_dsk_T [M][7], dL_dweights [P][K] (NULL = zero).
 * Outgoing (NULL = not wanted): dL_djoints [M][3], dL_dsk_r [M][4], dL_dsk_d_rot [M][4], dL_dsk_d_scale [M][3],
 * dL_dg_tr [7], dL_dsp_W [P][M] (dense, K non-zeros per row), dL_dsp_W_knn [P][K] (the same K values in KNN order -
 * the compact form data-parallel ranks exchange, since the KNN pattern is identical on every rank),
 * dL_dsp_radius [M], dL_dsp_weight [M].
 * workspace: device scratch of skgs_fk_lbs_workspace_bytes(M) bytes. */
SKGS_API int skgs_fk_lbs_backward(const skgs_skeleton* sk, int32_t P, const float* xyz, const float* sk_T,
                                  const float* weights, const int64_t* indices, const float* dL_dd_xyz,
                                  const float* dL_dd_rot, const float* dL_dd_scale, const float* dL_dsk_T,
                                  const float* dL_dweights, float* dL_djoints, float* dL_dsk_r, float* dL_dsk_d_rot,
                                  float* dL_dsk_d_scale, float* dL_dg_tr, float* dL_dsp_W, float* dL_dsp_W_knn,
                                  float* dL_dsp_radius, float* dL_dsp_weight, void* workspace, void* stream);

/* Output assembly: points = _xyz + d_xyz, scales = exp(_scaling) + d_scale, rotations = normalize(_rotation + d_rot)
 * (eps 1e-12), opacity = sigmoid(_opacity).  d_* may be NULL (static stage). */
SKGS_API int skgs_assemble_forward(int32_t P, const float* xyz, const float* scaling, const float* rotation,
                                   const float* opacity, const float* d_xyz, const float* d_rot, const float* d_scale,
                                   float* points, float* scales, float* rotations, float* opacities, void* stream);
SKGS_API int skgs_assemble_backward(int32_t P, const float* scaling, const float* rotation, const float* opacity,
                                    const float* d_rot, const float* dL_dpoints, const float* dL_dscales,
                                    const float* dL_drotations, const float* dL_dopacities, float* dL_dxyz,
                                    float* dL_dscaling, float* dL_drotation, float* dL_dopacity, float* dL_dd_xyz,
                                    float* dL_dd_rot, float* dL_dd_scale, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * sp-stage LBS: the superpoint twin of skgs_fk_lbs_*.  Replaces `warp` (networks/sk_gs.py:776-828)
 * and `calc_LBS_weight` (:751-774) as `sp_stage` calls them (:830-856): the per-superpoint SE3 comes from the
 * deformation network (sp_t = d_xyz, sp_r = normalize(d_rotation + bias)) instead of from forward kinematics.
 *   d_points   = sum_k w_k (R(sp_r_k) p + t_k) - p          method LBS   : t = sp_t
 *                                                           method LBS_C : t = sp_t + c + R(sp_r)(-c), c = sp_points
 *                R(sp_r_a) p + t_a - p, a = argmax_k w_k     method LARGEST (t as LBS)
 *   d_rotation = sum_k w_k sp_rot_k   (sp_rot == NULL: sp_r, :818-821);   d_scales = sum_k w_k sp_scale_k (NULL: 0)
 *   spT [M][7] = (t, sp_r)
 * Gradients w.r.t. sp_r through the rigid action are the tangent-space-projected ones lietorch's FromVec backward
 * returns (the component along sp_r is removed); the blend d_rotation = sum w sp_r contributes its plain gradient.
 * ------------------------------------------------------------------------------------------------------------- */
typedef enum skgs_warp_method { SKGS_WARP_LBS = 0, SKGS_WARP_LBS_C = 1, SKGS_WARP_LARGEST = 2 } skgs_warp_method;

typedef struct skgs_superpoints {
  int32_t M;               /* superpoints, 1..1024 (exps/default.yaml:25: 512) */
  int32_t K;               /* nearest superpoints per Gaussian, 1..8 */
  int32_t mode;            /* skgs_lbs_mode */
  int32_t method;          /* skgs_warp_method */
  float temperature;       /* SKGS_LBS_DIST only */
  const float* sp_points;  /* [M][3] */
  const float* sp_t;       /* [M][3] */
  const float* sp_r;       /* [M][4] xyzw (unit; normalised again inside) */
  const float* sp_rot;     /* [M][4] residual rotation blended into d_rotation, or NULL (blend sp_r) */
  const float* sp_scale;   /* [M][3] residual scale, or NULL */
  const float* sp_W;       /* [P][M]   (mode W) */
  const float* sp_radius;  /* [M] log-radius (kernel modes) */
  const float* sp_weight;  /* [M] logit (weighted_kernel) */
} skgs_superpoints;

SKGS_API size_t skgs_sp_lbs_workspace_bytes(int32_t M);
/* points [P][3] (constant: :834 detaches).  Outputs: d_points [P][3], d_rotation [P][4], d_scales [P][3], spT [M][7]
 * (may be NULL), weights [P][K], indices int64 [P][K]. */
SKGS_API int skgs_sp_lbs_forward(const skgs_superpoints* sp, int32_t P, const float* points, float* d_points,
                                 float* d_rotation, float* d_scales, float* spT, float* weights, int64_t* indices,
                                 void* workspace, void* stream);
/* Incoming gradients may be NULL (= zero); dL_dspT [M][7] / dL_dweights [P][K] are direct gradients on the auxiliary
 * outputs.  Outgoing (NULL = not wanted): dL_dsp_points [M][3] (through the weight function's distances and, LBS_C,
 * the centred rotation), dL_dsp_t [M][3], dL_dsp_r [M][4], dL_dsp_rot [M][4], dL_dsp_scale [M][3], dL_dsp_W [P][M]
 * (dense) or dL_dsp_W_knn [P][K] (compact), dL_dsp_radius [M], dL_dsp_weight [M]. */
SKGS_API int skgs_sp_lbs_backward(const skgs_superpoints* sp, int32_t P, const float* points, const float* spT,
                                  const float* weights, const int64_t* indices, const float* dL_dd_points,
                                  const float* dL_dd_rotation, const float* dL_dd_scales, const float* dL_dspT,
                                  const float* dL_dweights, float* dL_dsp_points, float* dL_dsp_t, float* dL_dsp_r,
                                  float* dL_dsp_rot, float* dL_dsp_scale, float* dL_dsp_W, float* dL_dsp_W_knn,
                                  float* dL_dsp_radius, float* dL_dsp_weight, void* workspace, void* stream);

/* The per-Gaussian forward of one view in TWO launches instead of four: forward kinematics once (one CTA), then one
 * kernel that does K nearest joints + skinning weights + linear blend + output assembly + preprocess + prefix sum + key
 * emission.  Same device functions and build flags as skgs_fk_lbs_forward -> skgs_assemble_forward ->
 * skgs_raster_forward_geometry, whose results it reproduces bit for bit (it is an execution plan of the same
 * operators: networks/sk_gs.py:1109-1150 + :1192,1202-1203 + gaussian_rasterizer_forward.cu:157-215 back to back).
 * Everything the three backward entry points need is written: weights / indices [P][K], d_rot [P][4], points / scales /
 * rotations (x,y,z,w) / opacities, sk_T [M][7], the geom arena, radii; keys are emitted into `binning` (R_cap entries):
 * continue with skgs_raster_forward_render(keys_emitted = 1).  LBS mode, K, joints ... come from `sk`; the skinning
 * table sk->sp_W is [P][M].  lbs_workspace: skgs_fk_lbs_workspace_bytes(sk->M) bytes. */
SKGS_API int skgs_deform_forward_geometry(const skgs_skeleton* sk, const skgs_raster_settings* s, int32_t P,
                                          int32_t M_sh, const float* xyz, const float* scaling, const float* rotation,
                                          const float* opacity_logit, const float* shs, float* points, float* scales,
                                          float* rotations, float* opacities, float* d_rot, float* weights,
                                          int64_t* indices, float* sk_T, void* lbs_workspace, void* geom, int32_t* radii,
                                          void* binning, int64_t R_cap, void* img, uint32_t* num_rendered_host,
                                          void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Widening rows: the step after the path (photometric loss) and the step after that (Adam)
 * ------------------------------------------------------------------------------------------------------------- */
/* loss = w_image * mean|I-G| (method 0; method 1: mean (I-G)^2) + w_ssim * (1 - mean SSIM_11x11(I, G)) and its gradient
 * w.r.t. the rendered image, times grad_scale.  Replaces ImageLoss.forward (networks/losses/image_loss.py:20-33, unmasked),
 * SSIM_Loss.forward + _ssim (networks/losses/ssim.py:27-62) and their autograd, called at networks/sk_gs.py:1528-1529
 * with the weights of exps/default.yaml:83-84.
 *   image  [3,H,W] channel-major (the rasterizer's output);
 *   target [3,H,W] if target_pixel_stride == 0, else pixel-major [H,W,stride] with stride 3 or 4 (RGB of an RGBA image);
 *   workspace: skgs_image_loss_workspace_bytes(H, W) bytes;
 *   loss_terms float[3] (device): {pixel term, 1 - mean SSIM, weighted total};
 *   dL_dimage [3,H,W] or NULL (forward only). */
SKGS_API size_t skgs_image_loss_workspace_bytes(int32_t H, int32_t W);
SKGS_API int skgs_image_loss(int32_t H, int32_t W, const float* image, const float* target, int32_t target_pixel_stride,
                             int32_t method, float w_image, float w_ssim, float grad_scale, void* workspace,
                             float* loss_terms, float* dL_dimage, void* stream);

/* One Adam step over up to SKGS_ADAM_MAX_TENSORS parameter tensors in ONE launch (torch.optim.Adam semantics, no weight
 * decay, no amsgrad - the optimizer the reference builds at networks/gaussian_splatting.py:445-453 with
 * exps/default.yaml:121-125 eps 1e-15, betas (0.9, 0.999)):
 *   m += (g - m)(1 - beta1);  v = beta2 v + (1 - beta2) g^2;
 *   p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps).
 * A tensor with `knn_indices != NULL` has its gradient in compact form: `grad` is [rows, K], `knn_indices` int64
 * [rows, K] names the columns of the [rows, cols] parameter that receive it (the skinning-weight table sp_W, whose
 * gradient has K non-zeros per row, networks/sk_gs.py:767-768); all other entries of the row see g = 0, exactly as the
 * dense torch optimizer does.
 * `dynamic_hyper` (device float[1 + 2 count], may be NULL) = {sqrt(1 - beta2^step), then per tensor lr_i / (1 - beta1^step)
 * and lr2_i / (1 - beta1^step)}: when
 * given it overrides `step` and `lr`, so that a captured CUDA graph can be replayed with the values of the current
 * iteration (uploaded by the caller) instead of the ones frozen at capture time.
 * `skip_if_nonzero` (device uint32, may be NULL): when the word is non-zero at execution time the whole step is a no-op.
 * Point it at skgs_raster_header.overflow of the render that produced the gradients: a fixed-capacity (CUDA-graph) render
 * whose binning arena overflowed yields an invalid image and invalid gradients, which must not reach the parameters. */
#define SKGS_ADAM_MAX_TENSORS 16
typedef struct skgs_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;              /* rows * cols for a compact-gradient tensor */
  double lr;
  double lr2;                 /* see period */
  int32_t period;             /* 0: every element uses lr.  > 0: element i uses lr if i % period < split, else lr2 - two
                                 interleaved param groups stored as one array, e.g. the SH coefficients [P,16,3] whose
                                 first 3 of every 48 floats are f_dc (lr_feature) and the rest f_rest (lr_feature / 20,
                                 networks/gaussian_splatting.py:448-449); removes the per-step cat(f_dc, f_rest) */
  int32_t split;
  int32_t cols;               /* compact gradient only */
  int32_t K;                  /* compact gradient only */
  const int64_t* knn_indices; /* NULL = dense gradient */
} skgs_adam_tensor;
SKGS_API int skgs_adam_step(const skgs_adam_tensor* tensors /* host */, int32_t count, int32_t step, double beta1,
                            double beta2, double eps, float grad_scale, const float* dynamic_hyper,
                            const uint32_t* skip_if_nonzero, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Densification bookkeeping, networks/gaussian_splatting.py:503-703.
 *
 * skgs_densify_stats: the per-step statistics of adaptive_control (:669-675 -> add_densification_stats :503-513):
 *   for every Gaussian with radii > 0:  max_radii2D = max(max_radii2D, radii);  grad_accum += |viewspace_grad[:2]|;
 *   denom += 1.   viewspace_grad is [P][grad_stride] (the rasterizer's dL/dmeans2D, stride 3; multi-view steps pass the
 *   SUM over views and the MAX of the radii, :509-512, :670-671).  `skip_if_nonzero`: as in skgs_adam_step.
 *
 * skgs_densify_plan + skgs_densify_apply: densify (:640-645 = clone :624-638 then split :589-622) and / or prune
 *   (:653-660) in ONE gather instead of four rounds of torch.cat / mask indexing over every parameter and both Adam
 *   moments (change_optimizer :515-563).  The plan decides per existing Gaussian
 *     g = grad_accum / denom (NaN -> 0);   hot = g >= grad_threshold;   smax = max exp(scaling)
 *     clone  = hot && smax <= densify_extent         (densify_percent_dense * cameras_extent)
 *     split  = hot && smax >  densify_extent         (N = 2 samples, the original is dropped)
 *     pruned = sigmoid(opacity) < min_opacity || (max_screen_size > 0 && (max_radii2D > max_screen_size ||
 *              smax > prune_extent))                 evaluated on the element's OWN values: clones inherit them,
 *              samples carry scaling log(exp(s) / 1.6); max_radii2D counts as 0 when do_densify (the reference zeroes
 *              it in densification_postfix :586 before prune runs)
 *   and lays the survivors out in the reference's final order
 *     [ kept originals | clones | split samples n = 0 | split samples n = 1 ]   (each in index order)
 *   as src[d] (source Gaussian of slot d), kind[d] (0 kept, 1 clone, 2 / 3 sample n = 0 / 1) and noise_row[d]
 *   (n * n_selected + rank of the source among the split-selected: the row of `samples` at :601-603).  src / kind /
 *   noise_row need room for 2 P entries.  `counts` (device) is written by the plan; the host reads it to size the new
 *   arrays and the noise table ([2 n_selected][3] standard normal: torch.normal(0, std) = noise * std).
 *   The apply pass gathers every listed tensor (row width `width`) and its Adam moments: kept rows keep their moments,
 *   new rows start at zero (:548-552); role XYZ rows of samples become R(normalize(rotation)) (noise * exp(scaling)) +
 *   xyz (:600-610), role SCALING rows of samples log(exp(s) / 1.6) (:611).
 * skgs_opacity_reset: reset_opacity (:662-665) - opacity = logit(min(sigmoid(opacity), cap)), moments zeroed.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct skgs_densify_config {
  int32_t do_densify;     /* clone + split */
  int32_t do_prune;
  float grad_threshold;   /* densify_grad_threshold (exps/default.yaml:70) */
  float densify_extent;   /* densify_percent_dense * cameras_extent */
  float min_opacity;      /* prune_opacity_threshold */
  float max_screen_size;  /* prune_max_screen_size, <= 0: the two size tests are off (size_threshold None, :688-691) */
  float prune_extent;     /* prune_percent_dense * cameras_extent */
} skgs_densify_config;

typedef struct skgs_densify_counts {
  uint32_t n_keep, n_clone, n_split, n_selected, n_new, reserved[3];
} skgs_densify_counts;

typedef enum skgs_densify_role { SKGS_DENSIFY_ROLE_COPY = 0, SKGS_DENSIFY_ROLE_XYZ = 1, SKGS_DENSIFY_ROLE_SCALING = 2 }
    skgs_densify_role;

#define SKGS_DENSIFY_MAX_TENSORS 16
typedef struct skgs_densify_tensor {
  const float* in;    /* [P][width] */
  float* out;         /* [P_new][width] */
  const float* m_in;  /* Adam exp_avg (or NULL) */
  float* m_out;
  const float* v_in;  /* Adam exp_avg_sq (or NULL) */
  float* v_out;
  int32_t width;
  int32_t role;       /* skgs_densify_role */
} skgs_densify_tensor;

SKGS_API int skgs_densify_stats(int32_t P, const int32_t* radii, const float* viewspace_grad, int32_t grad_stride,
                                float* max_radii2D, float* grad_accum, float* denom, const uint32_t* skip_if_nonzero,
                                void* stream);
SKGS_API size_t skgs_densify_workspace_bytes(int32_t P);
SKGS_API int skgs_densify_plan(const skgs_densify_config* cfg, int32_t P, const float* scaling, const float* opacity,
                               const float* grad_accum, const float* denom, const float* max_radii2D, int32_t* src,
                               uint8_t* kind, int32_t* noise_row, skgs_densify_counts* counts /* device */,
                               void* workspace, void* stream);
SKGS_API int skgs_densify_apply(const skgs_densify_tensor* tensors /* host */, int32_t count, int32_t P_new,
                                const int32_t* src, const uint8_t* kind, const int32_t* noise_row,
                                const float* scaling /* source [P][3] */, const float* rotation /* source [P][4] xyzw */,
                                const float* noise /* [2 n_selected][3] */, void* stream);
SKGS_API int skgs_opacity_reset(int32_t P, float* opacity, float* exp_avg, float* exp_avg_sq, float cap, void* stream);

/* Joint-rotation network of the `sk` stage, the step before forward kinematics:
 * joints [M,3], time t -> sk_r [M,4] (unit quaternion xyzw), d_rot [M,4], d_scale [M,3].
 * Replaces SimpleDeformationNetwork.forward (networks/sk_gs.py:134-164: freq encoders freqencoder.cu:7-31, MLP_with_skips
 * my_ext/blocks/mlp.py:44-85) plus the head of `kinematic` (sk_gs.py:1074-1076) and their autograd.
 * Parameters live in ONE flat array `theta`: for hidden layer i = 0..depth-1 the Linear weight [width, in_i] (row-major,
 * torch layout) then its bias [width]; then the three heads as one [11, in] weight and [11] bias
 * (rows 0-3 rotation, 4-7 d_rot, 8-10 d_scale).  in_0 = enc = 3 (1 + 2 degree_p) + (1 + 2 degree_t); in_i = width
 * (+ enc if bit i-1 of skip_mask is set: the encoded input is concatenated AFTER the ReLU of layer i-1).
 * `t` is a DEVICE pointer to one float (so that a captured graph can be replayed with a new time).
 * The backward uses the activations the forward left in `workspace` (same pointer, no call in between) and ASSIGNS
 * dL_dtheta [param_count] and dL_djoints [M,3] (the part that flows through the network input; may be NULL). */
typedef struct skgs_joint_mlp {
  int32_t M;
  int32_t degree_p;      /* 10 (exps/default.yaml:50) */
  int32_t degree_t;      /* 6  (exps/default.yaml:52) */
  int32_t width;         /* 256 */
  int32_t depth;         /* 8 */
  int32_t skip_mask;     /* skips [4] -> 1 << 4 */
  int32_t n_out;         /* 11 */
  int32_t rotation_head; /* 1: sk_r = normalize(out[0:4] + (0,0,0,1)), F.normalize eps 1e-12; 0: raw outputs */
  const float* theta;
} skgs_joint_mlp;
SKGS_API int skgs_joint_mlp_layout(const skgs_joint_mlp* net, int64_t* weight_offsets /* [depth+1] */,
                                   int64_t* bias_offsets /* [depth+1] */, int32_t* in_dims /* [depth+1] */,
                                   int64_t* param_count);
SKGS_API size_t skgs_joint_mlp_workspace_bytes(const skgs_joint_mlp* net);
SKGS_API int skgs_joint_mlp_forward(const skgs_joint_mlp* net, const float* joints, const float* t, float* sk_r,
                                    float* d_rot, float* d_scale, void* workspace, void* stream);
SKGS_API int skgs_joint_mlp_backward(const skgs_joint_mlp* net, const float* dL_dsk_r, const float* dL_dd_rot,
                                     const float* dL_dd_scale, float* dL_dtheta, float* dL_djoints, void* workspace,
                                     void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange
 * ------------------------------------------------------------------------------------------------------------- */
/* In-switch all-reduce (SUM, in place) of a flat fp32 arena that lives in symmetric memory bound to one NVLS multicast
 * address (`multicast_ptr`, e.g. from torch.distributed._symmetric_memory).  Rank `rank` reduces and re-broadcasts the
 * rank-th slice with multimem.ld_reduce / multimem.st.  The caller brackets the call with cross-GPU barriers on the
 * same stream.  The reference has no counterpart (its DDP wiring is unused, my_ext/framework.py:339-357). */
SKGS_API int skgs_multimem_allreduce(void* multicast_ptr, int64_t numel, int32_t rank, int32_t world, void* stream);


/* Multi-view steps: the reference loops over the views of a step and autograd SUMS their gradients
 * (networks/sk_gs.py:1220); the MAX of the screen radii over the views feeds max-radius tracking
 * (networks/gaussian_splatting.py:638-640).  dst[i] += src[i] over a flat fp32 arena (both 16-byte aligned) and
 * dst[i] = max(dst[i], src[i]) over int32 radii. */
SKGS_API int skgs_accumulate_f32(float* dst, const float* src, int64_t numel, void* stream);
SKGS_API int skgs_max_i32(int32_t* dst, const int32_t* src, int64_t numel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SKGS_B200_H_ */
