// K1 – occupied-voxel pixel mask, ray generation, SMPL-box intersection.
// Follows libs/renders/demo_render.py:166-239 op for op (see common.cuh on the
// exact-arithmetic convention); the integer results (pixel mask, kept-ray list)
// are bit-identical to the oracle.  Integer/byte HBM-bound work: one pass over
// the level-1 occupancy grid (4 B/voxel) and one over the H·W pixel mask.
#include "common.cuh"

namespace gpnerf {

__device__ __forceinline__ unsigned enc_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned e) {
  unsigned u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  return __uint_as_float(u);
}

__global__ void init_bounds(unsigned* enc) {
  if (threadIdx.x < 3) enc[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) enc[threadIdx.x] = 0u;
}

// demo_render.py:166-200.  One thread per level-1 voxel.
__global__ void __launch_bounds__(256) voxel_pixel_mask(const float* __restrict__ masks3d,
                                                        const __grid_constant__ gpnerf_frame_t fparam,
                                                        unsigned* __restrict__ enc_bounds,
                                                        float* __restrict__ pix_mask) {
  GPNERF_LOAD_FRAME(fparam)
  const int D = f.level_dims[0][0], H = f.level_dims[0][1], W = f.level_dims[0][2];
  const long long n = (long long)D * H * W;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n;
       v += (long long)gridDim.x * blockDim.x) {
    if (!(__ldg(masks3d + v) > f.mask_threshold)) continue;
    int w = (int)(v % W);
    int h = (int)((v / W) % H);
    int d = (int)(v / ((long long)W * H));
    // mask_xyz = (x,y,z) index · 2 ; → SMPL frame → world
    float sx = xadd(xmul((float)(2 * w), f.voxel_size[0]), f.bounds_min[0]);
    float sy = xadd(xmul((float)(2 * h), f.voxel_size[1]), f.bounds_min[1]);
    float sz = xadd(xmul((float)(2 * d), f.voxel_size[2]), f.bounds_min[2]);
    float wp[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      wp[c] = xadd(dot3(sx, f.R[c * 3 + 0], sy, f.R[c * 3 + 1], sz, f.R[c * 3 + 2]), f.Th[c]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lo[c] = fminf(lo[c], wp[c]);
      hi[c] = fmaxf(hi[c], wp[c]);
    }
    // project into the target view
    float cam[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      cam[c] = xadd(dot3(wp[0], f.target_pose[c * 4 + 0], wp[1], f.target_pose[c * 4 + 1], wp[2],
                         f.target_pose[c * 4 + 2]),
                    f.target_pose[c * 4 + 3]);
    float px = dot3(cam[0], f.target_K[0], cam[1], f.target_K[1], cam[2], f.target_K[2]);
    float py = dot3(cam[0], f.target_K[3], cam[1], f.target_K[4], cam[2], f.target_K[5]);
    float pz = dot3(cam[0], f.target_K[6], cam[1], f.target_K[7], cam[2], f.target_K[8]);
    long long x0 = __float2ll_rz(xdiv(px, pz)), y0 = __float2ll_rz(xdiv(py, pz));  // .long()
    long long x1 = x0 + 1, y1 = y0 + 1;
    const long long Wm = f.W - 1, Hm = f.H - 1;
    x0 = min(max(x0, 0ll), Wm);
    x1 = min(max(x1, 0ll), Wm);
    y0 = min(max(y0, 0ll), Hm);
    y1 = min(max(y1, 0ll), Hm);
    pix_mask[y0 * f.W + x0] = 1.0f;
    pix_mask[y1 * f.W + x0] = 1.0f;
    pix_mask[y0 * f.W + x1] = 1.0f;
    pix_mask[y1 * f.W + x1] = 1.0f;
  }
  // block reduction of the world-space bounds, one atomic set per CTA
  __shared__ float red[6][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
    if (lane == 0) {
      red[c][wid] = lo[c];
      red[3 + c][wid] = hi[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = red[threadIdx.x][0];
    for (int k = 1; k < 8; ++k)
      v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][k]) : fmaxf(v, red[threadIdx.x][k]);
    if (threadIdx.x < 3) {
      if (v != INFINITY) atomicMin(enc_bounds + threadIdx.x, enc_ordered(v));
    } else {
      if (v != -INFINITY) atomicMax(enc_bounds + threadIdx.x, enc_ordered(v));
    }
  }
}

// demo_render.py:171-175: min/max of the voxel points, z∓0.05
__global__ void finalize_bounds(const unsigned* __restrict__ enc, float* __restrict__ can_bounds) {
  int t = threadIdx.x;
  if (t >= 6) return;
  unsigned e = enc[t];
  float v;
  if (t < 3) v = (e == 0xffffffffu) ? INFINITY : dec_ordered(e);
  else v = (e == 0u) ? -INFINITY : dec_ordered(e);
  if (t == 2) v = xsub(v, 0.05f);
  if (t == 5) v = xadd(v, 0.05f);
  can_bounds[t] = v;
}

// camera centre: (−Rᵀ) @ T                                (demo_render.py:202)
__device__ __forceinline__ void camera_origin(const gpnerf_frame_t& f, float o[3]) {
  const float* P = f.target_pose;
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = dot3(-P[0 * 4 + i], P[3], -P[1 * 4 + i], P[7], -P[2 * 4 + i], P[11]);
}

struct RayBox {
  bool hit;
  float d[3];
  float near, far;
};

// demo_render.py:204-239 for one pixel
__device__ __forceinline__ RayBox ray_through_pixel(int p, const gpnerf_frame_t& f,
                                                    const float* __restrict__ cb, const float o[3]) {
  RayBox r;
  const float fi = (float)(p % f.W), fj = (float)(p / f.W);
  const float* Ki = f.target_K_inv;
  const float* P = f.target_pose;
  float q[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    q[c] = xsub(dot3(fi, Ki[c * 3 + 0], fj, Ki[c * 3 + 1], 1.0f, Ki[c * 3 + 2]), P[c * 4 + 3]);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    r.d[c] = xsub(dot3(q[0], P[0 * 4 + c], q[1], P[1 * 4 + c], q[2], P[2 * 4 + c]), o[c]);
  const float eps = 1e-6f;
  float lo[3], hi[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    lo[c] = xsub(cb[c], eps);
    hi[c] = xadd(cb[3 + c], eps);
  }
  int n_hit = 0;
  float hp[2][3];
#pragma unroll
  for (int m = 0; m < 6; ++m) {
    const int a = m % 3;
    float t = xdiv(xsub(cb[m], o[a]), r.d[a]);
    float px = xadd(xmul(t, r.d[0]), o[0]);
    float py = xadd(xmul(t, r.d[1]), o[1]);
    float pz = xadd(xmul(t, r.d[2]), o[2]);
    bool in = (px >= lo[0]) && (px <= hi[0]) && (py >= lo[1]) && (py <= hi[1]) && (pz >= lo[2]) &&
              (pz <= hi[2]);
    if (in) {
      if (n_hit < 2) {
        hp[n_hit][0] = px;
        hp[n_hit][1] = py;
        hp[n_hit][2] = pz;
      }
      ++n_hit;
    }
  }
  r.hit = (n_hit == 2);
  r.near = r.far = 0.0f;
  if (r.hit) {
    float nd = norm3(r.d[0], r.d[1], r.d[2]);
    float d0 = xdiv(norm3(xsub(hp[0][0], o[0]), xsub(hp[0][1], o[1]), xsub(hp[0][2], o[2])), nd);
    float d1 = xdiv(norm3(xsub(hp[1][0], o[0]), xsub(hp[1][1], o[1]), xsub(hp[1][2], o[2])), nd);
    if (f.neg_ray) d1 = -d1;
    r.near = fminf(d0, d1);
    r.far = fmaxf(d0, d1);
  }
  return r;
}

// one thread per pixel: flag = in voxel mask ∧ exactly-two-hits ∧ owned tile
__global__ void __launch_bounds__(256) ray_flags(const float* __restrict__ pix_mask,
                                                 const float* __restrict__ can_bounds,
                                                 const __grid_constant__ gpnerf_frame_t fparam,
                                                 uint32_t* __restrict__ words,
                                                 int32_t* __restrict__ counters,
                                                 float* __restrict__ rays_o) {
  GPNERF_LOAD_FRAME(fparam)
  const int n = f.H * f.W;
  const int n_warp_items = (n + 31) & ~31;
  float o[3];
  camera_origin(f, o);
  if (blockIdx.x == 0 && threadIdx.x < 3) rays_o[threadIdx.x] = o[threadIdx.x];
  float cb[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) cb[k] = __ldg(can_bounds + k);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_warp_items; p += gridDim.x * blockDim.x) {
    bool masked = (p < n) && (__ldg(pix_mask + p) == 1.0f);
    bool keep = false;
    // diagonal tile deal (gpnerf_b200/shard.py): owner = (tile + row of the tile's first pixel) % world
    const int tile = p / f.tile_px;
    if (masked && ((tile + (tile * f.tile_px) / f.W) % f.world == f.rank)) keep = ray_through_pixel(p, f, cb, o).hit;
    unsigned mb = __ballot_sync(0xffffffffu, masked);
    unsigned kb = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) {
      words[p >> 5] = kb;
      if (mb) atomicAdd(counters + GPNERF_CNT_PIX, __popc(mb));
    }
  }
}

__global__ void __launch_bounds__(256) ray_finalize(const int32_t* __restrict__ ray_pix,
                                                    const float* __restrict__ can_bounds,
                                                    const __grid_constant__ gpnerf_frame_t fparam,
                                                    const int32_t* __restrict__ counters,
                                                    float* __restrict__ rays_d,
                                                    float* __restrict__ near,
                                                    float* __restrict__ far) {
  GPNERF_LOAD_FRAME(fparam)
  const int n = __ldg(counters + GPNERF_CNT_RAYS);
  float o[3];
  camera_origin(f, o);
  float cb[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) cb[k] = __ldg(can_bounds + k);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    RayBox r = ray_through_pixel(__ldg(ray_pix + i), f, cb, o);
    rays_d[i * 3 + 0] = r.d[0];
    rays_d[i * 3 + 1] = r.d[1];
    rays_d[i * 3 + 2] = r.d[2];
    near[i] = r.near;
    far[i] = r.far;
  }
}


// ---------------------------------------------------------------------------
// Row f3: the dataset path's rays on the GPU – what the reference's
// CPU DataLoader workers compute per frame with numpy
// (libs/datasets/data_utils.py:47-63 get_rays, :331-337 sample_ray test split,
// :96-130 get_near_far).  Same dtype promotions: pixel → world in fp64, rays cast
// to fp32, box test in fp64 on those fp32 rays (bounds ± 0.01 folded in by the
// host, |d| < 1e-5 → 1e-5, eps 1e-6, exactly two hits, both depths signed by the
// first hit's side).  numpy never contracts a*b+c: explicit _rn intrinsics below.
// ---------------------------------------------------------------------------
struct DsRayArgs {
  double Kinv[9], Rinv[9], origin[3], bounds[6];   // bounds = (min xyz, max xyz) already widened by 0.01
  int H, W;
};
struct DsRay {
  float d[3], near, far;
  bool at_box;
};
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

__device__ DsRay dataset_ray(const DsRayArgs& a, int pix) {
  const double x = (double)(float)(pix % a.W), y = (double)(float)(pix / a.W);
  double pc[3], pw[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)      // np.dot(xy1, inv(K).T): dgemm accumulates k ascending with FMAs
    pc[c] = fma(1.0, a.Kinv[c * 3 + 2], fma(y, a.Kinv[c * 3 + 1], dmul(x, a.Kinv[c * 3 + 0])));
#pragma unroll
  for (int c = 0; c < 3; ++c)
    pw[c] = dadd(fma(pc[2], a.Rinv[c * 3 + 2], fma(pc[1], a.Rinv[c * 3 + 1], dmul(pc[0], a.Rinv[c * 3 + 0]))), a.origin[c]);
  DsRay r;
  float o32[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    r.d[c] = (float)dsub(pw[c], a.origin[c]);
    if (fabsf(r.d[c]) < 1e-5f) r.d[c] = 1e-5f;
    o32[c] = (float)a.origin[c];
  }
  const double eps = 1e-6;
  double hit[2][3];
  int n_hit = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int c = k % 3;
    const double t = dsub(a.bounds[k], (double)o32[c]) / (double)r.d[c];
    double p[3];
    bool in = true;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      p[q] = dadd(dmul(t, (double)r.d[q]), (double)o32[q]);
      in = in && (p[q] >= dsub(a.bounds[q], eps)) && (p[q] <= dadd(a.bounds[3 + q], eps));
    }
    if (in) {
      if (n_hit < 2) {
        hit[n_hit][0] = p[0]; hit[n_hit][1] = p[1]; hit[n_hit][2] = p[2];
      }
      ++n_hit;
    }
  }
  r.at_box = n_hit == 2;
  r.near = r.far = 0.0f;
  if (r.at_box) {
    const float nrm32 = __fsqrt_rn(xadd(xadd(xmul(r.d[0], r.d[0]), xmul(r.d[1], r.d[1])), xmul(r.d[2], r.d[2])));
    double dist[2], dot = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double vx = dsub(hit[h][0], (double)o32[0]), vy = dsub(hit[h][1], (double)o32[1]);
      const double vz = dsub(hit[h][2], (double)o32[2]);
      dist[h] = __dsqrt_rn(dadd(dadd(dmul(vx, vx), dmul(vy, vy)), dmul(vz, vz)));
      if (h == 0) dot = dadd(dadd(dmul(vx, (double)r.d[0]), dmul(vy, (double)r.d[1])), dmul(vz, (double)r.d[2]));
    }
    const double sign = dot < 0.0 ? -1.0 : 1.0;
    const double d0 = dmul(dist[0] / (double)nrm32, sign), d1 = dmul(dist[1] / (double)nrm32, sign);
    r.near = (float)fmin(d0, d1);
    r.far = (float)fmax(d0, d1);
  }
  return r;
}

__global__ void __launch_bounds__(256) dataset_ray_flags(const __grid_constant__ DsRayArgs a, uint32_t* __restrict__ words,
                                                         uint8_t* __restrict__ mask_at_box) {
  const int n = a.H * a.W, n_pad = (n + 31) & ~31;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pad; p += gridDim.x * blockDim.x) {
    bool keep = false;
    if (p < n) {
      keep = dataset_ray(a, p).at_box;
      mask_at_box[p] = keep ? 1 : 0;
    }
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) words[p >> 5] = b;
  }
}
__global__ void __launch_bounds__(256) dataset_ray_finalize(const __grid_constant__ DsRayArgs a,
                                                            const int32_t* __restrict__ ray_pix,
                                                            const int32_t* __restrict__ n_rays,
                                                            float* __restrict__ ray_o, float* __restrict__ ray_d,
                                                            float* __restrict__ near, float* __restrict__ far) {
  const int n = __ldg(n_rays);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const DsRay r = dataset_ray(a, __ldg(ray_pix + i));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ray_o[i * 3 + c] = (float)a.origin[c];
      ray_d[i * 3 + c] = r.d[c];
    }
    near[i] = r.near;
    far[i] = r.far;
  }
}

}  // namespace gpnerf

using namespace gpnerf;

static int check_frame(const gpnerf_frame_t* f) {
  if (!f) return 0;
  if (f->H <= 0 || f->W <= 0 || (long long)f->H * f->W > (1ll << 30)) return 0;
  if (f->world < 1 || f->rank < 0 || f->rank >= f->world || f->tile_px < 1) return 0;
  for (int k = 0; k < GPNERF_N_LEVELS; ++k)
    for (int j = 0; j < 3; ++j)
      if (f->level_dims[k][j] <= 0) return 0;
  return 1;
}

extern "C" {

int gpnerf_k1_voxel_pixel_mask(const float* masks3d, const gpnerf_frame_t* f, float* can_bounds,
                               float* pix_mask, void* stream) {
  GPNERF_REQUIRE(masks3d && can_bounds && pix_mask && check_frame(f));
  cudaStream_t st = (cudaStream_t)stream;
  // the ordered-int min/max scratch lives behind the 6 output floats' caller-owned
  // buffer: can_bounds must hold 12 floats (6 results + 6 scratch words)
  unsigned* enc = reinterpret_cast<unsigned*>(can_bounds + 6);
  cudaError_t e = cudaMemsetAsync(pix_mask, 0, sizeof(float) * (size_t)f->H * f->W, st);
  if (e != cudaSuccess) {
    set_error("memset pix_mask", e);
    return GPNERF_E_CUDA;
  }
  init_bounds<<<1, 32, 0, st>>>(enc);
  long long n = (long long)f->level_dims[0][0] * f->level_dims[0][1] * f->level_dims[0][2];
  long long blocks = (n + 255) / 256;
  int grid = (int)(blocks < (long long)sm_count() * 4 ? blocks : sm_count() * 4);
  voxel_pixel_mask<<<grid, 256, 0, st>>>(masks3d, *f, enc, pix_mask);
  finalize_bounds<<<1, 32, 0, st>>>(enc, can_bounds);
  return check_launch("k1_voxel_pixel_mask");
}

int gpnerf_k1_rays_bbox(const float* pix_mask, const float* can_bounds, const gpnerf_frame_t* f,
                        int32_t* ray_pix, float* rays_o, float* rays_d, float* near, float* far,
                        int32_t* counters, void* workspace, int32_t* tile_ray_begin, void* stream) {
  GPNERF_REQUIRE(pix_mask && can_bounds && ray_pix && rays_o && rays_d && near && far && counters &&
                 workspace && check_frame(f));
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)f->H * f->W;
  CompactWs ws = carve_workspace(workspace, n);
  cudaError_t e = cudaMemsetAsync(counters + GPNERF_CNT_PIX, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) {
    set_error("memset counters", e);
    return GPNERF_E_CUDA;
  }
  long long blocks = (n + 255) / 256;
  int grid = (int)(blocks < (long long)sm_count() * 8 ? blocks : sm_count() * 8);
  ray_flags<<<grid, 256, 0, st>>>(pix_mask, can_bounds, *f, ws.words, counters, rays_o);
  int rc = compact_launch(ws, nullptr, 1, n, n, ray_pix, counters + GPNERF_CNT_RAYS, st, f->tile_px, tile_ray_begin);
  if (rc != GPNERF_OK) return rc;
  ray_finalize<<<grid, 256, 0, st>>>(ray_pix, can_bounds, *f, counters, rays_d, near, far);
  return check_launch("k1_rays_bbox");
}

int gpnerf_k1_dataset_rays(const double* K_inv_host, const double* R_inv_host, const double* origin_host,
                           const double* bounds_host, int H, int W, int32_t* ray_pix, float* ray_o, float* ray_d,
                           float* near, float* far, uint8_t* mask_at_box, int32_t* n_rays, void* workspace,
                           void* stream) {
  GPNERF_REQUIRE(K_inv_host && R_inv_host && origin_host && bounds_host && ray_pix && ray_o && ray_d && near && far &&
                 mask_at_box && n_rays && workspace);
  GPNERF_REQUIRE(H > 0 && W > 0 && (long long)H * W < (1ll << 30));
  DsRayArgs a;
  for (int i = 0; i < 9; ++i) { a.Kinv[i] = K_inv_host[i]; a.Rinv[i] = R_inv_host[i]; }
  for (int i = 0; i < 3; ++i) a.origin[i] = origin_host[i];
  for (int i = 0; i < 6; ++i) a.bounds[i] = bounds_host[i];
  a.H = H; a.W = W;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)H * W;
  CompactWs ws = carve_workspace(workspace, n);
  long long blocks = (n + 255) / 256;
  int grid = (int)(blocks < (long long)sm_count() * 8 ? blocks : sm_count() * 8);
  dataset_ray_flags<<<grid, 256, 0, st>>>(a, ws.words, mask_at_box);
  int rc = compact_launch(ws, nullptr, 1, n, n, ray_pix, n_rays, st);
  if (rc != GPNERF_OK) return rc;
  dataset_ray_finalize<<<grid, 256, 0, st>>>(a, ray_pix, n_rays, ray_o, ray_d, near, far);
  return check_launch("k1_dataset_rays");
}

}  // extern "C"
