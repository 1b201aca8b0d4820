/*
 * gpnerf_abi.h – C ABI of libgpnerf_b200.so: the B200 (sm_100a) kernels of
 * GP-NeRF's geometry-guided progressive volume-rendering hot path.
 *
 * The reference (sail-sg/GP-Nerf) has no native code and therefore no FFI; it
 * reaches the GPU through ATen/cuBLAS/cuDNN calls issued from Python.  Each
 * entry point below replaces the group of torch calls named in its comment
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - all tensors are dense, row-major, fp32 unless stated; indices are int32;
 *  - nothing is allocated: the caller owns all buffers, sizes them for the
 *    worst case (see gpnerf_workspace_bytes) and keeps them alive until the
 *    stream has drained;
 *  - data-dependent sizes (rays, surviving points) stay on the device in
 *    `counters` (int32[GPNERF_N_COUNTERS]); kernels downstream read them there,
 *    so a frame is one sync-free stream of launches;
 *  - `stream` is a cudaStream_t passed as void*; functions are thread-safe when
 *    called with distinct streams and buffers;
 *  - return value: 0 = launched, <0 = GPNERF_E_* (never throws, never exits).
 */
#ifndef GPNERF_ABI_H
#define GPNERF_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPNERF_ABI_VERSION 2
#define GPNERF_MAX_PEERS 8
#define GPNERF_MAX_VIEWS 8
#define GPNERF_N_LEVELS 4
#define GPNERF_LEVEL_CH 32 /* head.sigma.outdims = [32]*4, configs/default.py:92 */

enum {
  GPNERF_OK = 0,
  GPNERF_E_ARG = -1,     /* bad argument (null pointer, size out of range)   */
  GPNERF_E_CUDA = -2,    /* a CUDA runtime call failed; see gpnerf_last_error */
  GPNERF_E_UNSUPPORTED = -3
};

/* slots of the device-side `counters` array */
enum {
  GPNERF_CNT_PIX = 0,   /* pixels set in the voxel-projection mask           */
  GPNERF_CNT_RAYS = 1,  /* rays kept by the box test (this rank's share)     */
  GPNERF_CNT_P1 = 2,    /* sample points surviving the occupancy test        */
  GPNERF_CNT_P2 = 3,    /* points surviving the density test                 */
  GPNERF_N_COUNTERS = 8
};

/* Per-frame constants, filled on the host.  Mirrors what render_rays reads
 * from `batch` / `sp_input` (libs/renders/demo_render.py:96-113, 393-427). */
typedef struct gpnerf_frame {
  /* SMPL pose: p_smpl = (p_world - Th) · R          (BaseRender.py:52-60)  */
  float R[9];              /* row-major 3x3 rotation matrix (batch['Rh'])   */
  float Th[3];
  float bounds_min[3];     /* batch['bounds'][0,0] – SMPL-frame min (x,y,z) */
  float voxel_size[3];     /* cfg.dataset.voxel_size (0.005)                */
  int32_t out_sh[3];       /* padded full-res volume shape (d,h,w)          */
  int32_t level_dims[GPNERF_N_LEVELS][3]; /* (D,H,W) of every dense level   */
  /* target camera                                   (demo_render.py:177-207) */
  float target_pose[12];   /* 3x4 row-major [R|T], world→camera              */
  float target_K[9];
  float target_K_inv[9];
  int32_t H, W;            /* target image size (reference hard-codes 512)   */
  /* source views                                    (BaseRender.py:233-323) */
  int32_t n_views;
  float src_KE[GPNERF_MAX_VIEWS][16]; /* K(4x4)·E(4x4), row-major, per view  */
  int32_t src_h, src_w;    /* source image size                              */
  int32_t feat_h, feat_w;  /* encoder feature-map size                       */
  /* sampling */
  int32_t n_samples;       /* cfg.train.n_samples (64)                       */
  int32_t neg_ray;         /* THuman convention (BaseRender.py:165-168)      */
  float mask_threshold;    /* 0.1 (demo_render.py:155)                       */
  /* ray sharding over ranks: pixel tiles of `tile_px` consecutive row-major
   * pixels are dealt diagonally; rank r keeps the tiles t with
   * (t + (t*tile_px)/W) % world == r */
  int32_t rank, world, tile_px;
  int32_t reserved_[2];
  /* Optional DEVICE address of a copy of this very struct.  When non-zero the
   * kernels read the per-frame values (pose, cameras, bounds …) from there
   * instead of from their by-value launch parameter, so a CUDA graph captured
   * once replays with the values of the current frame; shapes (H, W, dims,
   * n_views, n_samples) must not change between capture and replay. */
  uint64_t self_dev;
} gpnerf_frame_t;

/* Head weights, all device pointers, nn.Linear layout [out][in] row-major, as
 * they sit in the reference state_dict (libs/nerfheads/trainhead.py:39-41,
 * 85-110).  `rgb0` has 32·n_views input columns. */
typedef struct gpnerf_head_weights {
  const float *geo_w, *geo_b;                 /* sigmahead.out_geometry_fc.0   [64,128] */
  const float *den_w[4], *den_b[4];           /* rgbhead.out_geometry_fc.{0,2,4,6}: [64,134] [32,64] [16,32] [1,16] */
  const float *base_w[2], *base_b[2];         /* rgbhead.base_fc.{0,2}: [64,105] [32,64] */
  const float *vis_w[2], *vis_b[2];           /* rgbhead.vis_fc.{0,2}:  [32,32] [32,32]  */
  const float *rgb_w[3], *rgb_b[3];           /* rgbhead.rgb_fc.{0,2,4}: [32,32V] [16,32] [3,16] */
  /* bf16 UMMA-operand image of all of the above, produced by gpnerf_k3_pack_weights
   * (gpnerf_k3_packed_weight_bytes() bytes, 128-byte aligned); only read when
   * precision = 1, may be NULL otherwise. */
  const void *tc_image;
} gpnerf_head_weights_t;

/* Where K5 publishes this rank's finished pixel tiles when a frame (or a batch
 * of frames) is spread over the GPUs of one NVLink box: destination images in
 * this rank's and its peers' exchange buffers (peer pointers obtained with
 * gpnerf_peer_alloc / gpnerf_peer_open), and the arrival flags that replace a
 * collective.  Lives in DEVICE memory and is written once; the frame counter it
 * points to advances on the device.  The reference has no counterpart: its
 * renders leave the GPU through `.cpu().numpy()` (demo_render.py:346-353). */
typedef struct gpnerf_peer {
  int32_t n_dst;                             /* images every owned tile is written to (entry 0 = own) */
  int32_t n_flag;                            /* peers whose arrival flag is set when the frame is done */
  /* destinations, double-buffered: frame number `seq` uses half (seq & 1)                            */
  uint64_t dst_img[2][GPNERF_MAX_PEERS];     /* float   [H*W*3]                                        */
  uint64_t dst_hit[2][GPNERF_MAX_PEERS];     /* uint8_t [H*W]                                          */
  uint64_t dst_flag[GPNERF_MAX_PEERS];       /* int32_t*: this rank's slot in each peer's flag array   */
  uint64_t ticket;                           /* int32_t*: local CTA counter, zero between launches     */
  uint64_t seq;                              /* int32_t*: local frame counter; K5 renders frame
                                                *seq + 1 and its last CTA stores that number back – the
                                                host never touches it, so frames can be queued ahead   */
} gpnerf_peer_t;

int gpnerf_abi_version(void);
const char *gpnerf_last_error(void);
/* number of SMs of the current device (grid sizing is a multiple of it) */
int gpnerf_sm_count(void);
/* sizeof of the ABI structs as this library was compiled (0 gpnerf_frame_t,
 * 1 gpnerf_head_weights_t, 2 gpnerf_peer_t): a binding checks its own layout
 * against these before the first call. */
int gpnerf_struct_bytes(int which);

/* ---- K0: layout of the upstream products ------------------------------- */
/* One dense level NCDHW → NDHWC (one line per voxel; `storage`: 0 = 128 B fp32,
 * 1 = 64 B bf16, 2 = 64 B fp16 saturated at ±65504 – the storage the
 * tensor-core path gathers from and interpolates with HFMA2) plus the
 * fp32 per-voxel channel sum that SparseConvNet.encode reduces
 * (libs/nerfheads/networks/SparseConvNet.py:135-136).
 * With `pad` the output is written inside a one-voxel border,
 * [D+2][H+2][W+2][32], which the caller zeroes once: gathers from it need no
 * bounds tests (the border is grid_sample's zeros padding). */
int gpnerf_k0_level_to_channels_last(const float *ncdhw, int D, int H, int W, int storage, int pad,
                                     void *ndhwc, float *chan_sum, void *stream);
/* The same for everything the tensor-core path gathers from, in ONE launch: the
 * 4 levels (+ their channel sums) and the V encoder maps [V,32,fh,fw], all to
 * fp16 (storage 2) inside the zero border (pad 1).  featmaps may be NULL. */
int gpnerf_k0_products_to_f16(const float *const levels[GPNERF_N_LEVELS],
                              const int32_t level_dims[GPNERF_N_LEVELS][3], const float *featmaps,
                              int V, int fh, int fw, void *const levels_out[GPNERF_N_LEVELS],
                              float *const chan_sums[GPNERF_N_LEVELS], void *featmaps_out, void *stream);
/* SURVEY §8f row 1, first step – the pyramid's outputs without the dense detour:
 * the active rows of every level (feats[l] float[n_rows[l]][32], indices[l]
 * int32[n_rows[l]][idx_cols], last three columns = (d,h,w); what a
 * SparseConvTensor holds before .dense(), SparseConvNet.py:110) scattered straight
 * into the zero-bordered fp16 volumes + channel sums (both cleared first).
 * n_rows_dev (may be NULL): per level a DEVICE pointer to the live row count,
 * n_rows[l] then being the capacity of the arrays (gpnerf_sc_* outputs). */
int gpnerf_k0_sparse_to_f16(const float *const feats[GPNERF_N_LEVELS],
                            const int32_t *const indices[GPNERF_N_LEVELS],
                            const int32_t n_rows[GPNERF_N_LEVELS],
                            const int32_t *const n_rows_dev[GPNERF_N_LEVELS], int idx_cols,
                            const int32_t level_dims[GPNERF_N_LEVELS][3],
                            void *const levels_out[GPNERF_N_LEVELS],
                            float *const chan_sums[GPNERF_N_LEVELS], void *stream);
/* The same scatter into the fp32 channel-last volumes of the exact-arithmetic path
 * (float[D][H][W][32], no border): lets `render.file B200Render` run in fp32 from the
 * pyramid's rows (what a dataset batch leads to) without a dense NCDHW tensor. */
int gpnerf_k0_sparse_to_f32(const float *const feats[GPNERF_N_LEVELS],
                            const int32_t *const indices[GPNERF_N_LEVELS],
                            const int32_t n_rows[GPNERF_N_LEVELS],
                            const int32_t *const n_rows_dev[GPNERF_N_LEVELS], int idx_cols,
                            const int32_t level_dims[GPNERF_N_LEVELS][3],
                            void *const levels_out[GPNERF_N_LEVELS],
                            float *const chan_sums[GPNERF_N_LEVELS], void *stream);
/* masks3d on the level-1 grid = Σ_levels nearest-upsampled channel sums
 * (SparseConvNet.py:137-139). */
int gpnerf_k0_build_masks3d(const float *const chan_sum[GPNERF_N_LEVELS],
                            const gpnerf_frame_t *frame_host, float *masks3d, void *stream);
/* encoder maps [V,C=32,h,w] → [V,h,w,32]; images [V,3,H,W] → [V,H,W,4] (RGB,
 * pad); with `unnormalize` the [-1,1] inputs become x*0.5+0.5 on the way
 * (BaseRender.py:231), otherwise they are taken as already in [0,1]. */
int gpnerf_k0_featmaps_to_channels_last(const float *nchw, int V, int h, int w, int storage, int pad,
                                        void *nhwc, void *stream);   /* pad: [V][h+2][w+2][32] */
int gpnerf_k0_images_to_rgbx(const float *nchw, int V, int H, int W, int unnormalize, int pad,
                             float *rgbx, void *stream);              /* pad: [V][H+2][W+2][4]  */

/* ---- K1: pixel mask, rays, box intersection ---------------------------- */
/* demo_render.py:166-200: occupied voxels → world → can_bounds (min/max,
 * z∓0.05) and the 4-neighbour pixel mask of their projection.
 * can_bounds float[12]: [0..5] = (min xyz, max xyz), [6..11] = scratch for the
 * ordered-int atomics; pix_mask float[H*W] (1.0 / 0.0). */
int gpnerf_k1_voxel_pixel_mask(const float *masks3d, const gpnerf_frame_t *frame_host,
                               float *can_bounds, float *pix_mask, void *stream);
/* demo_render.py:200-239: rays through masked pixels, 6-plane test with
 * exactly two hits, near/far.  Kept rays are written in ascending pixel order
 * for this rank's tiles.  ray_pix int32[H*W], rays_d float[H*W*3], near/far
 * float[H*W], rays_o float[3]; counters[PIX], counters[RAYS] are set. */
int gpnerf_k1_rays_bbox(const float *pix_mask, const float *can_bounds,
                        const gpnerf_frame_t *frame_host, int32_t *ray_pix, float *rays_o,
                        float *rays_d, float *near, float *far, int32_t *counters,
                        void *workspace, int32_t *tile_ray_begin, void *stream);
/* tile_ray_begin (may be NULL): int32[ceil(H*W/tile_px) + 1], CSR offsets of the
 * rays of every pixel tile (K5 walks its tiles with them). */

/* The dataset path's rays: libs/datasets/data_utils.py:47-63
 * get_rays + the test-split branch of sample_ray (:331-337) + get_near_far
 * (:96-130) for every pixel of an H×W view, with numpy's dtype promotions (fp64
 * pixel→world, fp32 rays, fp64 box test).  HOST inputs (computed by the caller
 * with the same numpy calls the reference makes): K_inv = inv(K) and
 * R_inv = inv(R) row-major 3×3, origin = -R_inv·T, bounds = (min xyz, max xyz)
 * of the fp32 box widened by ∓0.01 in fp64.  Outputs (device): mask_at_box
 * uint8[H*W]; for the n_rays[0] rays that hit the box exactly twice, in ascending
 * pixel order: ray_pix, ray_o / ray_d [n][3], near / far [n].  Buffers sized H*W;
 * workspace as for the other compactions. */
int gpnerf_k1_dataset_rays(const double *K_inv_host, const double *R_inv_host,
                           const double *origin_host, const double *bounds_host, int H, int W,
                           int32_t *ray_pix, float *ray_o, float *ray_d, float *near, float *far,
                           uint8_t *mask_at_box, int32_t *n_rays, void *workspace, void *stream);

/* ---- K2: occupancy test, gathers --------------------------------------- */
/* demo_render.py:59-94, 270-283: sample S depths per ray (t_vals = linspace,
 * optional jitter t_rand[R*S] or NULL), world→SMPL, trilinear tap of masks3d,
 * valid = ascending flat indices (ray*S+sample) of points with tap > 0.
 * If masks3d is NULL every point is kept (BaseRender semantics).
 * n_rays_max bounds the grid; the live count is counters[RAYS].
 * Outputs valid int32[n_rays_max*S], counters[P1]. */
int gpnerf_k2_occupancy_compact(const float *masks3d, const float *rays_o, const float *rays_d,
                                const float *near, const float *far, const float *t_vals,
                                const float *t_rand, const gpnerf_frame_t *frame_host,
                                int n_rays_max, int32_t *valid, float *z_vals,
                                int32_t *counters, void *workspace, int32_t *ray_pt_begin,
                                void *stream);
/* ray_pt_begin (may be NULL): int32[n_rays_max + 1], CSR offsets of the
 * surviving points of every ray (ray_pt_begin[n_rays] = counters[P1]). */
/* Where the two gathers take their sample points from (`point_kind`):
 *   0  ray-parametrised: flat index valid[i] → (ray, sample), p = o + d·z;
 *      the live count is counters[P1] (n_points_max bounds the grid);
 *   1  explicit world points[n][3]   – the Projector.compute(xyz, …) API;
 *   2  explicit normalised grid coordinates[n][3] in [-1,1] – the
 *      SparseConvNet.forward(x, grid_coords) API (volume gather only).
 * For kinds 1/2 pass counters = NULL and n_points_max = n. */

/* SparseConvNet.py:111-122: 4-level trilinear gather (align_corners, zeros);
 * vol_feat float[n][128] level-major. */
int gpnerf_k2_gather_volume(const float *const levels_ndhwc[GPNERF_N_LEVELS], int point_kind,
                            const int32_t *valid, const float *rays_o, const float *rays_d,
                            const float *z_vals, const float *points,
                            const gpnerf_frame_t *frame_host, int n_points_max,
                            const int32_t *counters, float *vol_feat, void *stream);
/* Projector.compute + fused_mean_variance (demo_render.py:560-609,
 * trainhead.py:20-24): rgb_feat float[n][V][35] (RGB first), mask
 * float[n][V], meanvar float[n][70] = (mean 35 | var 35). */
int gpnerf_k2_project_gather_meanvar(const float *images_rgbx, const float *featmaps_nhwc,
                                     int point_kind, const int32_t *valid, const float *rays_o,
                                     const float *rays_d, const float *z_vals, const float *points,
                                     const gpnerf_frame_t *frame_host, int n_points_max,
                                     const int32_t *counters, float *rgb_feat, float *mask,
                                     float *meanvar, void *stream);
/* fused_mean_variance alone (trainhead.py:20-24): rgb_feat [n][V][35] → [n][70] */
int gpnerf_k2_mean_variance(const float *rgb_feat, int n_views, int n_points, float *meanvar,
                            void *stream);

/* ---- K3: heads ---------------------------------------------------------- */
/* For both heads the live row count is counters[counter_slot], or exactly
 * n_points_max when counters is NULL. */

/* trainhead.py:39-41,61-76,102-110,133-137: sigma_feat = ELU(Linear 128→64);
 * σ = ReLU(MLP 134→64→32→16→1 on [sigma_feat|mean|var]); σ = 0 where no view
 * is valid.  input_kind 0: `feat_in` = vol_feat float[n][128];
 * input_kind 1: `feat_in` = sigma_feat float[n][64] (NeRFRGBHead.forward API).
 * sigma float[n]; sigma_feat_out float[n][64] or NULL.
 * precision: 0 = fp32 CUDA cores (parity), 1 = bf16 tcgen05 tensor cores. */
int gpnerf_k3_density_mlp(const float *feat_in, int input_kind, const float *meanvar,
                          const float *mask, const gpnerf_head_weights_t *weights_host, int n_views,
                          int n_points_max, const int32_t *counters, int counter_slot, float *sigma,
                          float *sigma_feat_out, int precision, void *stream);
/* trainhead.py:128-145 colour trunk on the rows listed in valid1 (indices into
 * rgb_feat/meanvar; NULL = all rows in order); rgb float[rows][3] is written at
 * those rows. */
int gpnerf_k3_color_mlp(const float *rgb_feat, const float *meanvar, const int32_t *valid1,
                        const gpnerf_head_weights_t *weights_host, int n_views, int n_points_max,
                        const int32_t *counters, int counter_slot, float *rgb, int precision,
                        void *stream);

/* Pack the fp32 head weights into the bf16 K-major operand layout the tcgen05
 * kernels consume (once per weight update). n_views in 1..4. */
int64_t gpnerf_k3_packed_weight_bytes(void);
int gpnerf_k3_pack_weights(const gpnerf_head_weights_t *weights_host, int n_views, void *image,
                           void *stream);

/* ---- K2+K3 fused, bf16 tensor-core path ---------------------------------- */
/* Gathers (SparseConvNet.py:111-122, BaseRender.py:283-363), mean/variance
 * (trainhead.py:20-24) and the density head (trainhead.py:39-41,102-110,
 * 133-137) in one kernel: gathered features go straight into the shared-memory
 * operand tiles of the tcgen05 GEMM chain.  Inputs are the fp16 channel-last
 * zero-bordered products of K0 (storage = 2, pad = 1) and the padded fp32 RGBx
 * images; points are the
 * ray-parametrised survivors `valid` (count = counters[P1]).  Outputs σ
 * float[P1] and one bf16 record of gpnerf_k23_record_bytes(V) bytes per point
 * ([mean|var] and the V per-view rows in operand order) for
 * gpnerf_k3_color_mlp_records.  n_views in 1..4. */
int64_t gpnerf_k23_record_bytes(int n_views);
int gpnerf_k23_gather_density_tc(const void *const levels_f16[GPNERF_N_LEVELS],
                                 const void *featmaps_f16, const float *images_rgbx,
                                 const int32_t *valid, const float *rays_o, const float *rays_d,
                                 const float *z_vals, const gpnerf_frame_t *frame_host,
                                 const gpnerf_head_weights_t *weights_host, int n_points_max,
                                 const int32_t *counters, float *sigma, void *records, float *alpha,
                                 void *k4_workspace, void *stream);
/* With alpha / k4_workspace (both or neither) the kernel also writes K4's
 * α = 1-exp(-σ) and its survivor flags straight into the compaction workspace:
 * follow with gpnerf_k4_compact_alpha(sigma = NULL, …). */
/* Tile hand-off to the colour head (the default of the tensor-core path): the same kernel, which additionally
 * leaves the colour head's inputs behind – per 128-point tile of the P1 list one block of
 * gpnerf_k23_tile_record_bytes(V) bytes, already in the tcgen05 operand layouts of
 * gpnerf_k3_color_tiles_tc (mean|var features, the V per-view feature rows, RGB mean/var, per-view RGB:
 * BaseRender.py:283-363, trainhead.py:20-24), so that a point's pixel-aligned features are gathered once per
 * frame.  tile_records: ceil(n_points_max / 128) blocks, 1024-byte aligned, ZERO-FILLED ONCE by the caller
 * (columns no view owns are never written).  rgb_in (may be NULL) float[P1][V][3] receives the per-view RGB
 * taps (BaseRender's rgb_in_map input).  alpha / k4_workspace as above. */
int64_t gpnerf_k23_tile_record_bytes(int n_views);
int gpnerf_k23_gather_density_tiles_tc(const void *const levels_f16[GPNERF_N_LEVELS],
                                       const void *featmaps_f16, const float *images_rgbx,
                                       const int32_t *valid, const float *rays_o, const float *rays_d,
                                       const float *z_vals, const gpnerf_frame_t *frame_host,
                                       const gpnerf_head_weights_t *weights_host, int n_points_max,
                                       const int32_t *counters, float *sigma, void *tile_records,
                                       float *rgb_in, float *alpha, void *k4_workspace, void *stream);
/* Colour trunk (trainhead.py:85-100, 118-145) on the tiles of the P1 list, fed from the tile records above
 * (one bulk copy per tile; no gather).  rgb float[P1][3] is written for every point of every processed tile.
 * k4_workspace (may be NULL): the workspace gpnerf_k23_gather_density_tiles_tc wrote the progressive step's
 * survivor flags into – tiles without a survivor are skipped (demo_render.py:312-333: colour only for the
 * survivors; here at tile granularity, K5 ignores the colour of a culled point).  Count =
 * counters[counter_slot] (the P1 count).  n_views in 1..4. */
int gpnerf_k3_color_tiles_tc(const void *tile_records, const void *k4_workspace,
                             const gpnerf_head_weights_t *weights_host, int n_views, int n_points_max,
                             int32_t *counters, int counter_slot, float *rgb, void *stream);
/* With k4_workspace the kernel also leaves the progressive step's survivor count in counters[GPNERF_CNT_P2]
 * (popcount of the flags; the ordered list itself – gpnerf_k4_compact_alpha(sigma = NULL) – is only needed by
 * callers that want valid1).  counters[6], counters[7] are scratch: zero before the first call, zero after each. */
/* Colour trunk (trainhead.py:128-145) on the record rows listed in valid1. */
int gpnerf_k3_color_mlp_records(const void *records, const int32_t *valid1,
                                const gpnerf_head_weights_t *weights_host, int n_views,
                                int n_points_max, const int32_t *counters, int counter_slot,
                                float *rgb, void *stream);

/* Colour trunk (trainhead.py:85-100, 118-145) for the points listed in valid1 (indices into
 * the P1 arrays; NULL = all P1 points in order), gathering its own inputs: the V-view
 * pixel-aligned features and RGB taps (BaseRender.py:283-363) and their mean / variance
 * (trainhead.py:20-24) from the fp16 channel-last zero-bordered feature maps and the padded
 * RGBx images of K0 – no per-point record is written upstream of the progressive step.
 * rgb float[P1][3] is written at the P1 index of every processed point; rgb_in (may be
 * NULL) float[P1][V][3] receives the per-view RGB taps (BaseRender's rgb_in_map input).
 * Count = counters[counter_slot].  n_views in 1..4. */
int gpnerf_k3_color_gather_tc(const void *featmaps_f16, const float *images_rgbx, const int32_t *valid,
                              const int32_t *valid1, const float *rays_o, const float *rays_d,
                              const float *z_vals, const gpnerf_frame_t *frame_host,
                              const gpnerf_head_weights_t *weights_host, int n_points_max,
                              const int32_t *counters, int counter_slot, float *rgb, float *rgb_in,
                              void *stream);

/* ---- K7: the sparse-conv geometry encoder -------------- */
/* libs/nerfheads/networks/SparseConvNet.py:21-124 without spconv (1.2.1 @ abf0acf3, not in the reference tree,
 * does not build for sm_100): SubMConv3d / SparseConv3d(3, 2, padding 1) as their published semantics – a dense
 * cross-correlation restricted to the active output sites, weights [kd][kh][kw][in][out] – followed by the
 * (folded, inference-form) BatchNorm1d scale/shift and ReLU.  Rows = active sites; counts stay on the device.
 *  index_input : voxel coords [n][cols] (d,h,w last) → de-duplicated sites (the smallest row id owns a voxel):
 *                owners int32[n] (input rows, ascending), coords_out int32[n][3], idx_vol int32[D*H*W]
 *                (row id per site, 0x7f7f7f7f = none), n_out[0].
 *  gather_rows : feat_out[j] = feat_in[rows[j]] for j < n_dev[0].
 *  strided_sites: sites of the next level (stride 2) + its index volume; out_lin = scratch int32[n_out_max].
 *  neighbours  : nbr[k*n_out_max + o] = input row at o·stride − 1 + k (k = (kd*3+kh)*3+kw), −1 if none; int32[27*n_out_max].
 *                One table per (output site list, stride, input level); shared by the convolutions on it.
 *  conv        : out_feat[o] = relu(scale ⊙ Σ_k W[k]ᵀ·in[nbr[k][o]] + shift).
 * workspace: gpnerf_workspace_bytes(max(n, voxels of the level being compacted)). */
int gpnerf_sc_index_input(const int32_t *coords, int cols, int n, int D, int H, int W, int32_t *idx_vol,
                          int32_t *owners, int32_t *coords_out, int32_t *n_out, void *workspace,
                          void *stream);
int gpnerf_sc_gather_rows(const float *feat_in, int C, const int32_t *rows, const int32_t *n_dev,
                          int n_max, float *feat_out, void *stream);
int gpnerf_sc_strided_sites(const int32_t *in_coords, const int32_t *n_in_dev, int n_in_max, int Do,
                            int Ho, int Wo, int32_t *out_lin, int32_t *out_coords,
                            int32_t *out_idx_vol, int32_t *n_out_dev, void *workspace, void *stream);
int gpnerf_sc_neighbours(const int32_t *out_coords, const int32_t *n_out_dev, int n_out_max, int stride,
                         const int32_t *in_idx_vol, int Di, int Hi, int Wi, const int32_t *n_in_dev,
                         int32_t *nbr, void *stream);
int gpnerf_sc_conv(const float *in_feat, int c_in, const int32_t *nbr, const int32_t *n_out_dev,
                   int n_out_max, const float *weight, const float *scale, const float *shift,
                   int c_out, float *out_feat, void *stream);

/* The same convolution on tensor cores (tcgen05.mma.kind::tf32, three-term hi/lo split: fp32-grade accuracy).
 * Features travel split between the layers: a row is [hi (c) | lo (c)] floats, hi = the value with its 13 low
 * mantissa bits cleared (TF32-exact), lo = the remainder.  gather_rows_split makes the first layer's rows;
 * conv_tc writes out_split [n][2*c_out] and, if out_full != NULL, plain fp32 rows [n][c_out] as well.
 * w_packed: per tap k two UMMA B-operand images of W[k]^T ([c_out x c_in], K-major, 8x16-byte core matrices),
 * hi then lo; float[27][2][c_out*c_in] – element (n, c) at ((n/8)*(c_in/4)*32 + (c/4)*32 + (n%8)*4 + c%4). */
int gpnerf_sc_gather_rows_split(const float *feat_in, int C, const int32_t *rows, const int32_t *n_dev,
                                int n_max, float *feat_out, void *stream);
int gpnerf_sc_conv_tc(const float *in_split, int c_in, const int32_t *nbr, const int32_t *n_out_dev,
                      int n_out_max, const float *w_packed, const float *scale, const float *shift,
                      int c_out, float *out_split, float *out_full, void *stream);

/* ---- K8: SMPL-code attention (trainhead.py:48-51; MultiHeadAttention.py:40-98, sum=False) ---- */
/* out[i] = W_fc · concat_h( softmax_v( (W_q·code[i])_h/√d_k · (W_k·feat[i,v])_h ) · (W_v·feat[i,v])_h ).
 * code [n][d_model]; feat[i,v] = feats + i*vertex_stride + v*view_stride (floats, kv_dim of them – lets the
 * caller pass the [n][V][35] rows of gpnerf_k2_project_gather_meanvar with an offset of 3); weights in
 * torch.nn.Linear layout [out][in]; out [n][d_model].  (d_model, kv_dim, n_head·d_k) ∈ {(16,32,16),(32,32,32)},
 * n_views ≤ 8, n_head ≤ 8. */
int gpnerf_attn_smpl_code(const float *code, const float *feats, long long view_stride,
                          long long vertex_stride, int n, int n_views, const float *w_q,
                          const float *w_k, const float *w_v, const float *w_fc, int d_model,
                          int kv_dim, int n_head, int d_k, float *out, void *stream);

/* ---- K9: the image encoder's layers between its convolutions (UNet.py:32-51,120-130,204-216) ---- */
/* Tensors are channels-last, dtype 0 = fp32, 1 = bf16, 2 = fp16.
 * Output geometry (y, H, W, y_pad, y_ctot, y_coff): y is [N][H+2*y_pad][W+2*y_pad][y_ctot]; the call writes
 * channels [y_coff, y_coff+C); y_pad = 1 also fills the one-pixel border with the reflection of the interior
 * (= F.pad(mode="reflect") for the next 3x3 convolution).
 * instance_norm_act: x dense [N][H][W][C]; residual NULL or [N][H+2*res_pad][W+2*res_pad][C];
 *   y = act((x - mean_nc) * rsqrt(var_nc + eps) * gamma_c + beta_c [+ residual]), biased variance over H*W,
 *   act 0 none | 1 ReLU | 2 ELU; scratch: 256 + scratch_nc*24 bytes (8-byte aligned, scratch_nc >= N*C, N <= 64),
 *   zeroed ONCE by the caller – the kernels leave its accumulators zeroed again, so consecutive norms on one
 *   stream share it without a memset in between; y may alias x when y_pad = 0 and y_ctot = C.
 * resample_pad: src [N][Hs+2*src_pad][Ws+2*src_pad][C]; mode 0 copy (H = Hs, W = Ws), mode 1 bilinear
 *   (align_corners = True) to H x W, mode 2 every second pixel (H = ceil(Hs/2), W = ceil(Ws/2)).
 * C, y_ctot, y_coff multiples of 4 (fp32) / 8 (16-bit); for instance_norm_act C/4 resp. C/8 divides 256. */
int gpnerf_k9_instance_norm_act(const void *x, const void *residual, int res_pad, int dtype, int N, int H,
                                int W, int C, const float *gamma, const float *beta, float eps, int act,
                                void *scratch, int scratch_nc, void *y, int y_pad, int y_ctot, int y_coff,
                                void *stream);
int gpnerf_k9_resample_pad(const void *src, int dtype, int N, int Hs, int Ws, int src_pad, int C, int mode,
                           void *y, int H, int W, int y_pad, int y_ctot, int y_coff, void *stream);

/* ---- K4: progressive step ---------------------------------------------- */
/* demo_render.py:312-317: α = 1-exp(-σ); valid1 = ascending indices (into the
 * P1 arrays) with α > 1e-14; counters[P2]. */
/* sigma == NULL: α and the flags come from gpnerf_k23_gather_density_tc (above);
 * only the compaction runs. */
int gpnerf_k4_compact_alpha(const float *sigma, int n_points_max, int32_t *counters,
                            float *alpha, int32_t *valid1, void *workspace, void *stream);

/* ---- K5: compositing ---------------------------------------------------- */
/* demo_render.py:335-353 without the dense scatter: per ray, walk its surviving
 * points (CSR offsets ray_pt_begin from K2), T = Π(1-α+1e-10) exclusive by a
 * warp product scan, rgb_map = Σ α·T·rgb.  One CTA per pixel tile of this rank
 * (tile_ray_begin from K1); the finished tile – zeros where no ray – goes with
 * coalesced stores to pred_img float[H*W*3] / hit_mask uint8[H*W] (the
 * reference's host-side `pred_img[mask_at_box] = rgb_map`), so neither needs
 * clearing.  With `peer_dev` (DEVICE pointer to a gpnerf_peer_t) the tile is
 * written to every image listed there instead – this rank's and its peers',
 * over NVLink – and the frame's sequence number is published in the peers'
 * arrival flags when the last tile has left; gpnerf_peer_wait is the matching
 * wait.  t_min > 0 stops a ray once T < t_min (off at 0: exact). */
int gpnerf_k5_composite(const float *alpha, const float *rgb, const int32_t *ray_pix,
                        const int32_t *tile_ray_begin, const int32_t *ray_pt_begin,
                        const gpnerf_frame_t *frame_host, float t_min, float *rgb_map,
                        float *pred_img, uint8_t *hit_mask, const gpnerf_peer_t *peer_dev,
                        void *stream);
/* Blocks the stream until flags[k] >= *peer_dev->seq (the frame K5 just
 * published) for every k != self. */
int gpnerf_peer_wait(const int32_t *flags, int n_flags, int self, const gpnerf_peer_t *peer_dev,
                     void *stream);
/* Peer (CUDA IPC) memory for the exchange buffers.  HOST-side calls: alloc =
 * cudaMalloc + zero + cudaIpcGetMemHandle (handle_host receives
 * gpnerf_peer_handle_bytes() bytes to send to the other ranks), open =
 * cudaIpcOpenMemHandle on a handle received from a peer. */
int gpnerf_peer_handle_bytes(void);
int gpnerf_peer_alloc(int64_t bytes, void **dev_ptr_host, void *handle_host);
int gpnerf_peer_open(const void *handle_host, void **dev_ptr_host);
int gpnerf_peer_close(void *dev_ptr);
int gpnerf_peer_free(void *dev_ptr);
/* Renderer.raw2outputs (BaseRender.py:75-107,147): dense [R][S] path.
 * raw float[R][S][4] (rgb,σ); rgb_in float[R][S][V][3] or NULL.  Outputs
 * rgb_map[R][3], disp/acc/depth[R], weights[R][S], rgb_in_map[R][V][3]. */
int gpnerf_k5_raw2outputs(const float *raw, const float *z_vals, const float *rgb_in, int n_rays,
                          int n_samples, int n_views, int neg, float *rgb_map, float *disp,
                          float *acc, float *depth, float *weights, float *rgb_in_map, void *stream);

/* Backward of gpnerf_k5_raw2outputs (training, BaseRender.py:75-107 under
 * autograd): upstream gradients of rgb_map [R][3], disp/acc/depth [R], weights
 * [R][S], rgb_in_map [R][3V] (any may be NULL = zero) → d_raw float[R][S][4]
 * (∂L/∂rgb, ∂L/∂σ per sample). */
int gpnerf_k5_raw2outputs_bwd(const float *raw, const float *z_vals, const float *rgb_in, int n_rays,
                              int n_samples, int n_views, int neg, const float *g_rgb_map,
                              const float *g_disp, const float *g_acc, const float *g_depth,
                              const float *g_weights, const float *g_rgb_in_map, float *d_raw,
                              void *stream);

/* ---- K6: training path (layer-wise fp32; BASELINE configs[3]) ------------- */
/* Backward of the gathers: scatter-add of the weighted upstream gradients into
 * channel-last fp32 gradient buffers (caller zeroes them). Same point sources
 * as the forward gathers. d_vol_feat float[n][128]; d_rgb_feat float[n][V][35]
 * (only the 32 feature channels propagate: the images carry no gradient). */
int gpnerf_k2_gather_volume_bwd(float *const d_levels_ndhwc[GPNERF_N_LEVELS], int point_kind,
                                const int32_t *valid, const float *rays_o, const float *rays_d,
                                const float *z_vals, const float *points,
                                const gpnerf_frame_t *frame_host, int n_points_max,
                                const int32_t *counters, const float *d_vol_feat, void *stream);
int gpnerf_k2_project_gather_bwd(float *d_featmaps_nhwc, int point_kind, const int32_t *valid,
                                 const float *rays_o, const float *rays_d, const float *z_vals,
                                 const float *points, const gpnerf_frame_t *frame_host,
                                 int n_points_max, const int32_t *counters, const float *d_rgb_feat,
                                 void *stream);
/* Y[p][0..N) = epi(X'[p][0..K)·B + bias) for p < P, rows strided by ldx / ldy;
 * X' = in_scale·X, times ELU'(in_aux) element-wise when in_aux is given (in_aux
 * = the saved ELU output the incoming gradient has to pass through);
 * B[k][n] = w_is_kn ? W[k*ldw+n] : W[n*ldw+k] (nn.Linear layout → forward; its
 * transpose → ∂L/∂X).  epilogue: 0 none, 1 ELU, 2 ReLU, 3 sigmoid, 4 multiply
 * by ELU'(aux[p][n]).  add_pre / add_post add the previous content of Y before
 * / after the epilogue.  precision: 0 = fp32 FMA chains on CUDA cores (the
 * parity path), 1 = tcgen05.mma.kind::tf32 with fp32 accumulators in TMEM. */
int gpnerf_k6_linear(const float *X, int ldx, int K, float in_scale, const float *in_aux,
                     int ld_in_aux, const float *W, int ldw, int w_is_kn, int N, const float *bias,
                     int epilogue, const float *aux, int ld_aux, float *Y, int ldy, int add_pre,
                     int add_post, long long P, int precision, void *stream);
/* dW[n*ldw+k] += Σ_p dY'[p][n]·X[p][k]·in_scale ; db[n] += Σ_p dY'[p][n] (db may
 * be NULL); dY' = dY ⊙ ELU'(dy_aux) when dy_aux is given */
int gpnerf_k6_grad_weights(const float *X, int ldx, int K, float in_scale, const float *dY, int ldy,
                           int N, const float *dy_aux, int ld_dy_aux, float *dW, int ldw, float *db,
                           long long P, int precision, void *stream);
/* raw[p] = (rgb[p], σ[p]) with σ = s_relu[p], forced to 0 where Σ_v mask[p][v] < 1
 * (trainhead.py:136-137,162); and the matching split of ∂L/∂raw into the
 * pre-sigmoid / pre-ReLU gradients */
int gpnerf_k6_assemble_raw(const float *rgb, const float *s_relu, const float *mask, int n_views,
                           long long P, float *raw, void *stream);
int gpnerf_k6_raw_grad_split(const float *d_raw, const float *rgb, const float *s_relu,
                             const float *mask, int n_views, long long P, float *d_rgb_pre,
                             float *d_s_pre, void *stream);
/* fused_mean_variance backward: d_rgb_feat += dmean/V + dvar·2(x−mean)/V */
int gpnerf_k6_meanvar_bwd(const float *rgb_feat, const float *meanvar, const float *d_meanvar,
                          int n_views, long long P, float *d_rgb_feat, void *stream);
/* [batch][n][32] channel-last → [batch][32][n] (gradients back to NC(D)HW) */
int gpnerf_k6_from_channels_last(const float *nxc, int batch, long long n, float *cxn, void *stream);

/* bytes of scratch `workspace` needed by the compaction passes for up to
 * n_items flags */
int64_t gpnerf_workspace_bytes(int64_t n_items);

#ifdef __cplusplus
}
#endif
#endif /* GPNERF_ABI_H */
