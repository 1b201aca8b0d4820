// K2 – occupancy test (+ first compaction) and the two gathers.
//
//  * occupancy_flags : demo_render.py:59-94, 270-281 – one thread per sample
//    point; trilinear tap of masks3d (32 B/point, L2-resident 3 MB grid).
//  * gather_volume   : SparseConvNet.py:111-122 – 4 levels × 8 corners × one
//    128-byte channel line per surviving point (4,096 B/point requested).
//  * project_gather  : BaseRender.py:283-363 + trainhead.py:20-24 – V views ×
//    4 corners × (16 B RGBx + 128 B feature line) and the mean/variance over
//    views, fused (1,680 B/point requested at V=3).
//
// Gather kernels use 8 lanes per point, 4 channels (one float4) per lane: a
// corner is one 128-byte coalesced request per 8-lane group, and the index
// arithmetic is shared by 4 points per warp instruction.  Bound: L2/HBM
// bandwidth.
#include "common.cuh"

namespace gpnerf {

__global__ void __launch_bounds__(256) occupancy_flags(const float* __restrict__ masks3d,
                                                       const float* __restrict__ rays_o,
                                                       const float* __restrict__ rays_d,
                                                       const float* __restrict__ near,
                                                       const float* __restrict__ far,
                                                       const float* __restrict__ t_vals,
                                                       const float* __restrict__ t_rand,
                                                       const __grid_constant__ gpnerf_frame_t fparam,
                                                       const int32_t* __restrict__ counters,
                                                       uint32_t* __restrict__ words,
                                                       float* __restrict__ z_vals, int n_rays_max) {
  GPNERF_LOAD_FRAME(fparam)
  const int S = f.n_samples;
  const long long n = (long long)min(__ldg(counters + GPNERF_CNT_RAYS), n_rays_max) * S;   // never past the buffers
  const long long n_pad = (n + 31) & ~31ll;
  const float o[3] = {__ldg(rays_o), __ldg(rays_o + 1), __ldg(rays_o + 2)};
  const int D = f.level_dims[0][0], H = f.level_dims[0][1], W = f.level_dims[0][2];
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pad;
       p += (long long)gridDim.x * blockDim.x) {
    bool keep = false;
    if (p < n) {
      const int r = (int)(p / S), s = (int)(p - (long long)r * S);
      const float d[3] = {__ldg(rays_d + r * 3), __ldg(rays_d + r * 3 + 1), __ldg(rays_d + r * 3 + 2)};
      const float z = sample_depth(__ldg(near + r), __ldg(far + r), t_vals, s, S, t_rand, p);
      z_vals[p] = z;
      if (masks3d == nullptr) {
        keep = true;
      } else {
        Tri t = trilinear_setup(world_to_grid(f, point_on_ray(o, d, z)), D, H, W);
        // all 8 taps are issued up front from clamped addresses (8 loads in flight instead of 8 dependent
        // branches); an out-of-range corner contributes the exact +0 it is skipped with in the reference sum
        float m[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int zc = min(max(t.z0 + (c >> 2), 0), D - 1), yc = min(max(t.y0 + ((c >> 1) & 1), 0), H - 1);
          const int xc = min(max(t.x0 + (c & 1), 0), W - 1);
          m[c] = __ldg(masks3d + ((zc * H + yc) * W + xc));
        }
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc = xadd(acc, ((t.inb >> c) & 1u) ? xmul(m[c], t.w[c]) : 0.0f);
        keep = acc > 0.0f;
      }
    }
    unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) words[p >> 5] = b;
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct LevelPtrs {
  const float* p[GPNERF_N_LEVELS];
};

// Where a kernel takes its sample points from:
//  kind 0: ray-parametrised – flat index valid[i] → (ray, sample), p = o + d·z
//  kind 1: explicit world points[i][3]            (Projector.compute API)
//  kind 2: explicit normalised grid coords[i][3]  (SparseConvNet.forward API)
struct PointSrc {
  const int32_t* valid;
  const float *rays_o, *rays_d, *z_vals, *points;
  int kind, S;
};
__device__ __forceinline__ Vec3 fetch_world_point(const PointSrc& ps, long long i) {
  if (ps.kind != 0) {
    Vec3 p;
    p.x = __ldg(ps.points + i * 3);
    p.y = __ldg(ps.points + i * 3 + 1);
    p.z = __ldg(ps.points + i * 3 + 2);
    return p;
  }
  const int q = __ldg(ps.valid + i);
  const int r = q / ps.S;
  const float o[3] = {__ldg(ps.rays_o), __ldg(ps.rays_o + 1), __ldg(ps.rays_o + 2)};
  const float d[3] = {__ldg(ps.rays_d + r * 3), __ldg(ps.rays_d + r * 3 + 1), __ldg(ps.rays_d + r * 3 + 2)};
  return point_on_ray(o, d, __ldg(ps.z_vals + q));
}

__global__ void __launch_bounds__(256) gather_volume(LevelPtrs lv, PointSrc ps,
                                                     const __grid_constant__ gpnerf_frame_t fparam,
                                                     const int32_t* __restrict__ count_ptr, int n_const,
                                                     float* __restrict__ vol_feat) {
  GPNERF_LOAD_FRAME(fparam)
  const int n = count_ptr ? __ldg(count_ptr) : n_const;
  const int sub = threadIdx.x & 7;  // which float4 of the 128-byte line
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long n_groups = ((long long)gridDim.x * blockDim.x) >> 3;
  for (long long i = group; i < n; i += n_groups) {
    const Vec3 wp = fetch_world_point(ps, i);
    const Vec3 g = (ps.kind == 2) ? wp : world_to_grid(f, wp);
#pragma unroll
    for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
      const int D = f.level_dims[l][0], H = f.level_dims[l][1], W = f.level_dims[l][2];
      const Tri t = trilinear_setup(g, D, H, W);
      float4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if ((t.inb >> c) & 1u) {
          long long idx = ((long long)(t.z0 + (c >> 2)) * H + (t.y0 + ((c >> 1) & 1))) * W + (t.x0 + (c & 1));
          v[c] = ld4(lv.p[l] + idx * 32 + sub * 4);
        } else {
          v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if ((t.inb >> c) & 1u) {  // ATen adds in-bounds corners only, mul then add
          acc.x = xadd(acc.x, xmul(v[c].x, t.w[c]));
          acc.y = xadd(acc.y, xmul(v[c].y, t.w[c]));
          acc.z = xadd(acc.z, xmul(v[c].z, t.w[c]));
          acc.w = xadd(acc.w, xmul(v[c].w, t.w[c]));
        }
      }
      *reinterpret_cast<float4*>(vol_feat + i * 128 + l * 32 + sub * 4) = acc;
    }
  }
}

// bilinear tap of a channel-last map, ATen's vectorised CPU kernel order:
// w = x-floor(x), e = 1-w, n = y-floor(y), s = 1-n; nw=s·e, ne=s·w, sw=n·e, se=n·w
struct Bil {
  int x0, y0;
  float nw, ne, sw, se;
  unsigned inb;
};
__device__ __forceinline__ Bil bilinear_setup(float nx, float ny, int Wm, int Hm) {
  float ix = xmul(xadd(nx, 1.0f), (float)(Wm - 1) * 0.5f);
  float iy = xmul(xadd(ny, 1.0f), (float)(Hm - 1) * 0.5f);
  float fx = floorf(ix), fy = floorf(iy);
  Bil b;
  b.x0 = (int)fminf(fmaxf(fx, -2.0f), (float)Wm + 1.0f);
  b.y0 = (int)fminf(fmaxf(fy, -2.0f), (float)Hm + 1.0f);
  float w = xsub(ix, fx), e = xsub(1.0f, w), n = xsub(iy, fy), s = xsub(1.0f, n);
  b.nw = xmul(s, e);
  b.ne = xmul(s, w);
  b.sw = xmul(n, e);
  b.se = xmul(n, w);
  bool finite = (ix == ix) && (iy == iy);
  bool x0ok = b.x0 >= 0 && b.x0 < Wm, x1ok = b.x0 + 1 >= 0 && b.x0 + 1 < Wm;
  bool y0ok = b.y0 >= 0 && b.y0 < Hm, y1ok = b.y0 + 1 >= 0 && b.y0 + 1 < Hm;
  b.inb = finite ? ((x0ok && y0ok) | ((x1ok && y0ok) << 1) | ((x0ok && y1ok) << 2) | ((x1ok && y1ok) << 3)) : 0u;
  return b;
}
__device__ __forceinline__ float4 bil_tap(const float* __restrict__ base, int pitch_floats, int Wm,
                                          const Bil& b) {
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long i00 = ((long long)b.y0 * Wm + b.x0) * pitch_floats;
  float4 a = (b.inb & 1u) ? ld4(base + i00) : z;
  float4 c = (b.inb & 2u) ? ld4(base + i00 + pitch_floats) : z;
  float4 d = (b.inb & 4u) ? ld4(base + i00 + (long long)Wm * pitch_floats) : z;
  float4 e = (b.inb & 8u) ? ld4(base + i00 + (long long)(Wm + 1) * pitch_floats) : z;
  float4 r;
  r.x = a.x * b.nw + c.x * b.ne + d.x * b.sw + e.x * b.se;
  r.y = a.y * b.nw + c.y * b.ne + d.y * b.sw + e.y * b.se;
  r.z = a.z * b.nw + c.z * b.ne + d.z * b.sw + e.z * b.se;
  r.w = a.w * b.nw + c.w * b.ne + d.w * b.sw + e.w * b.se;
  return r;
}

template <int V>
__global__ void __launch_bounds__(256) project_gather_meanvar(
    const float* __restrict__ images_rgbx, const float* __restrict__ featmaps, PointSrc ps,
    const __grid_constant__ gpnerf_frame_t fparam, const int32_t* __restrict__ count_ptr, int n_const,
    float* __restrict__ rgb_feat, float* __restrict__ mask, float* __restrict__ meanvar) {
  GPNERF_LOAD_FRAME(fparam)
  const int n = count_ptr ? __ldg(count_ptr) : n_const;
  const int sub = threadIdx.x & 7;
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long n_groups = ((long long)gridDim.x * blockDim.x) >> 3;
  const float wm1 = xsub((float)f.src_w, 1.0f), hm1 = xsub((float)f.src_h, 1.0f);
  const long long img_stride = (long long)f.src_h * f.src_w * 4;
  const long long map_stride = (long long)f.feat_h * f.feat_w * 32;
  constexpr int CF = 35;
  for (long long i = group; i < n; i += n_groups) {
    const Vec3 pt = fetch_world_point(ps, i);
    float4 feat[V];
    float4 rgb[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float* KE = f.src_KE[v];
      float qx = dot4(KE[0], pt.x, KE[1], pt.y, KE[2], pt.z, KE[3], 1.0f);
      float qy = dot4(KE[4], pt.x, KE[5], pt.y, KE[6], pt.z, KE[7], 1.0f);
      float qz = dot4(KE[8], pt.x, KE[9], pt.y, KE[10], pt.z, KE[11], 1.0f);
      float px = fminf(fmaxf(xdiv(qx, qz), -1e6f), 1e6f);
      float py = fminf(fmaxf(xdiv(qy, qz), -1e6f), 1e6f);
      bool front = f.neg_ray ? (qz < 0.0f) : (qz > 0.0f);
      bool inb = (px <= wm1) && (px >= 0.0f) && (py <= hm1) && (py >= 0.0f);
      float nx = xsub(xdiv(xmul(2.0f, px), wm1), 1.0f);
      float ny = xsub(xdiv(xmul(2.0f, py), hm1), 1.0f);
      Bil bi = bilinear_setup(nx, ny, f.src_w, f.src_h);
      Bil bf = bilinear_setup(nx, ny, f.feat_w, f.feat_h);
      rgb[v] = bil_tap(images_rgbx + v * img_stride, 4, f.src_w, bi);
      feat[v] = bil_tap(featmaps + v * map_stride + sub * 4, 32, f.feat_w, bf);
      float* row = rgb_feat + (i * V + v) * CF;
      row[3 + sub * 4 + 0] = feat[v].x;
      row[3 + sub * 4 + 1] = feat[v].y;
      row[3 + sub * 4 + 2] = feat[v].z;
      row[3 + sub * 4 + 3] = feat[v].w;
      if (sub == 0) {
        row[0] = rgb[v].x;
        row[1] = rgb[v].y;
        row[2] = rgb[v].z;
        mask[i * V + v] = (inb && front) ? 1.0f : 0.0f;
      }
    }
    // population mean / variance over views, unmasked (trainhead.py:20-24)
    float4 m = feat[0], mr = rgb[0];
#pragma unroll
    for (int v = 1; v < V; ++v) {
      m.x += feat[v].x; m.y += feat[v].y; m.z += feat[v].z; m.w += feat[v].w;
      mr.x += rgb[v].x; mr.y += rgb[v].y; mr.z += rgb[v].z;
    }
    const float fv = (float)V;
    m.x = xdiv(m.x, fv); m.y = xdiv(m.y, fv); m.z = xdiv(m.z, fv); m.w = xdiv(m.w, fv);
    mr.x = xdiv(mr.x, fv); mr.y = xdiv(mr.y, fv); mr.z = xdiv(mr.z, fv);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), qr = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float t;
      t = feat[v].x - m.x; q.x += t * t;
      t = feat[v].y - m.y; q.y += t * t;
      t = feat[v].z - m.z; q.z += t * t;
      t = feat[v].w - m.w; q.w += t * t;
      t = rgb[v].x - mr.x; qr.x += t * t;
      t = rgb[v].y - mr.y; qr.y += t * t;
      t = rgb[v].z - mr.z; qr.z += t * t;
    }
    float* mv = meanvar + i * 2 * CF;
    mv[3 + sub * 4 + 0] = m.x;
    mv[3 + sub * 4 + 1] = m.y;
    mv[3 + sub * 4 + 2] = m.z;
    mv[3 + sub * 4 + 3] = m.w;
    mv[CF + 3 + sub * 4 + 0] = xdiv(q.x, fv);
    mv[CF + 3 + sub * 4 + 1] = xdiv(q.y, fv);
    mv[CF + 3 + sub * 4 + 2] = xdiv(q.z, fv);
    mv[CF + 3 + sub * 4 + 3] = xdiv(q.w, fv);
    if (sub == 0) {
      mv[0] = mr.x; mv[1] = mr.y; mv[2] = mr.z;
      mv[CF + 0] = xdiv(qr.x, fv); mv[CF + 1] = xdiv(qr.y, fv); mv[CF + 2] = xdiv(qr.z, fv);
    }
  }
}

// fused_mean_variance on its own (trainhead.py:20-24): rgb_feat [n][V][35] → [n][70]
__global__ void __launch_bounds__(256) mean_variance(const float* __restrict__ rgb_feat, int V, long long n,
                                                     float* __restrict__ meanvar) {
  const long long total = n * 35;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / 35;
    const int c = (int)(t - i * 35);
    const float* x = rgb_feat + i * V * 35 + c;
    float m = __ldg(x);
    for (int v = 1; v < V; ++v) m += __ldg(x + v * 35);
    m = xdiv(m, (float)V);
    float q = 0.0f;
    for (int v = 0; v < V; ++v) {
      float d = __ldg(x + v * 35) - m;
      q += d * d;
    }
    meanvar[i * 70 + c] = m;
    meanvar[i * 70 + 35 + c] = xdiv(q, (float)V);
  }
}

// ---------------------------------------------------------------------------
// Backward of the two gathers (training): grid_sample's gradient w.r.t. its
// input = scatter-add of weight·upstream into the sampled corners.  Same point
// sources, same weights (trilinear_setup / bilinear_setup) as the forward
// kernels above; 16-byte vector atomics into channel-last fp32 gradient
// buffers (sums are order-dependent at the ulp level, like torch's).
// ---------------------------------------------------------------------------
struct LevelGradPtrs {
  float* p[GPNERF_N_LEVELS];
};

__global__ void __launch_bounds__(256) scatter_volume_bwd(LevelGradPtrs lg, PointSrc ps,
                                                          const __grid_constant__ gpnerf_frame_t fparam,
                                                          const int32_t* __restrict__ count_ptr, int n_const,
                                                          const float* __restrict__ d_vol_feat) {
  GPNERF_LOAD_FRAME(fparam)
  const int n = count_ptr ? __ldg(count_ptr) : n_const;
  const int sub = threadIdx.x & 7;
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long n_groups = ((long long)gridDim.x * blockDim.x) >> 3;
  for (long long i = group; i < n; i += n_groups) {
    const Vec3 wp = fetch_world_point(ps, i);
    const Vec3 g = (ps.kind == 2) ? wp : world_to_grid(f, wp);
#pragma unroll
    for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
      const int D = f.level_dims[l][0], H = f.level_dims[l][1], W = f.level_dims[l][2];
      const Tri t = trilinear_setup(g, D, H, W);
      const float4 d = ld4(d_vol_feat + i * 128 + l * 32 + sub * 4);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if ((t.inb >> c) & 1u) {
          long long idx = ((long long)(t.z0 + (c >> 2)) * H + (t.y0 + ((c >> 1) & 1))) * W + (t.x0 + (c & 1));
          const float w = t.w[c];
          atomicAdd(reinterpret_cast<float4*>(lg.p[l] + idx * 32 + sub * 4),
                    make_float4(d.x * w, d.y * w, d.z * w, d.w * w));
        }
      }
    }
  }
}

template <int V>
__global__ void __launch_bounds__(256) scatter_features_bwd(float* __restrict__ d_featmaps, PointSrc ps,
                                                            const __grid_constant__ gpnerf_frame_t fparam,
                                                            const int32_t* __restrict__ count_ptr, int n_const,
                                                            const float* __restrict__ d_rgb_feat) {
  GPNERF_LOAD_FRAME(fparam)
  const int n = count_ptr ? __ldg(count_ptr) : n_const;
  const int sub = threadIdx.x & 7;
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long n_groups = ((long long)gridDim.x * blockDim.x) >> 3;
  const float wm1 = xsub((float)f.src_w, 1.0f), hm1 = xsub((float)f.src_h, 1.0f);
  const long long map_stride = (long long)f.feat_h * f.feat_w * 32;
  for (long long i = group; i < n; i += n_groups) {
    const Vec3 pt = fetch_world_point(ps, i);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float* KE = f.src_KE[v];
      float qx = dot4(KE[0], pt.x, KE[1], pt.y, KE[2], pt.z, KE[3], 1.0f);
      float qy = dot4(KE[4], pt.x, KE[5], pt.y, KE[6], pt.z, KE[7], 1.0f);
      float qz = dot4(KE[8], pt.x, KE[9], pt.y, KE[10], pt.z, KE[11], 1.0f);
      float px = fminf(fmaxf(xdiv(qx, qz), -1e6f), 1e6f);
      float py = fminf(fmaxf(xdiv(qy, qz), -1e6f), 1e6f);
      float nx = xsub(xdiv(xmul(2.0f, px), wm1), 1.0f);
      float ny = xsub(xdiv(xmul(2.0f, py), hm1), 1.0f);
      const Bil b = bilinear_setup(nx, ny, f.feat_w, f.feat_h);
      const float* dr = d_rgb_feat + (i * V + v) * 35 + 3 + sub * 4;
      const float4 d = make_float4(__ldg(dr), __ldg(dr + 1), __ldg(dr + 2), __ldg(dr + 3));
      float* base = d_featmaps + v * map_stride + sub * 4;
      const float wts[4] = {b.nw, b.ne, b.sw, b.se};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if ((b.inb >> c) & 1u) {
          const long long idx = ((long long)(b.y0 + (c >> 1)) * f.feat_w + (b.x0 + (c & 1))) * 32;
          const float w = wts[c];
          atomicAdd(reinterpret_cast<float4*>(base + idx), make_float4(d.x * w, d.y * w, d.z * w, d.w * w));
        }
      }
    }
  }
}

static int persistent_grid(int ctas_per_sm) { return sm_count() * ctas_per_sm; }

}  // namespace gpnerf

using namespace gpnerf;

extern "C" {

int gpnerf_k2_occupancy_compact(const float* masks3d, const float* rays_o, const float* rays_d,
                                const float* near, const float* far, const float* t_vals,
                                const float* t_rand, const gpnerf_frame_t* f, int n_rays_max,
                                int32_t* valid, float* z_vals, int32_t* counters, void* workspace,
                                int32_t* ray_pt_begin, void* stream) {
  GPNERF_REQUIRE(rays_o && rays_d && near && far && t_vals && f && valid && z_vals && counters && workspace);
  GPNERF_REQUIRE(n_rays_max > 0 && f->n_samples > 0 && (long long)n_rays_max * f->n_samples < (1ll << 31));
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_max = (long long)n_rays_max * f->n_samples;
  CompactWs ws = carve_workspace(workspace, n_max);
  long long blocks = (n_max + 255) / 256;
  int grid = (int)(blocks < (long long)persistent_grid(8) ? blocks : persistent_grid(8));
  occupancy_flags<<<grid, 256, 0, st>>>(masks3d, rays_o, rays_d, near, far, t_vals, t_rand, *f,
                                        counters, ws.words, z_vals, n_rays_max);
  return compact_launch(ws, counters + GPNERF_CNT_RAYS, f->n_samples, 0, n_max, valid,
                        counters + GPNERF_CNT_P1, st, f->n_samples, ray_pt_begin);
}

static int make_point_src(PointSrc* ps, int point_kind, const int32_t* valid, const float* rays_o,
                          const float* rays_d, const float* z_vals, const float* points,
                          const gpnerf_frame_t* f) {
  ps->valid = valid; ps->rays_o = rays_o; ps->rays_d = rays_d; ps->z_vals = z_vals; ps->points = points;
  ps->kind = point_kind; ps->S = f->n_samples;
  if (point_kind == 0) return valid && rays_o && rays_d && z_vals && f->n_samples > 0;
  if (point_kind == 1 || point_kind == 2) return points != nullptr;
  return 0;
}

int gpnerf_k2_gather_volume(const float* const levels[GPNERF_N_LEVELS], int point_kind,
                            const int32_t* valid, const float* rays_o, const float* rays_d,
                            const float* z_vals, const float* points, const gpnerf_frame_t* f,
                            int n_points_max, const int32_t* counters, float* vol_feat, void* stream) {
  GPNERF_REQUIRE(levels && f && vol_feat && n_points_max > 0);
  PointSrc ps;
  GPNERF_REQUIRE(make_point_src(&ps, point_kind, valid, rays_o, rays_d, z_vals, points, f));
  LevelPtrs lv;
  for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
    GPNERF_REQUIRE(levels[l] != nullptr);
    lv.p[l] = levels[l];
  }
  long long blocks = ((long long)n_points_max * 8 + 255) / 256;
  int grid = (int)(blocks < (long long)persistent_grid(8) ? blocks : persistent_grid(8));
  gather_volume<<<grid, 256, 0, (cudaStream_t)stream>>>(lv, ps, *f, counters ? counters + GPNERF_CNT_P1 : nullptr,
                                                        n_points_max, vol_feat);
  return check_launch("k2_gather_volume");
}

int gpnerf_k2_project_gather_meanvar(const float* images_rgbx, const float* featmaps, int point_kind,
                                     const int32_t* valid, const float* rays_o, const float* rays_d,
                                     const float* z_vals, const float* points, const gpnerf_frame_t* f,
                                     int n_points_max, const int32_t* counters, float* rgb_feat,
                                     float* mask, float* meanvar, void* stream) {
  GPNERF_REQUIRE(images_rgbx && featmaps && f && rgb_feat && mask && meanvar && n_points_max > 0);
  PointSrc ps;
  GPNERF_REQUIRE(point_kind != 2 && make_point_src(&ps, point_kind, valid, rays_o, rays_d, z_vals, points, f));
  const int32_t* cp = counters ? counters + GPNERF_CNT_P1 : nullptr;
  long long blocks = ((long long)n_points_max * 8 + 255) / 256;
  int grid = (int)(blocks < (long long)persistent_grid(8) ? blocks : persistent_grid(8));
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_V(VV)                                                                                   \
  case VV:                                                                                             \
    project_gather_meanvar<VV><<<grid, 256, 0, st>>>(images_rgbx, featmaps, ps, *f, cp, n_points_max,  \
                                                     rgb_feat, mask, meanvar);                         \
    break;
  switch (f->n_views) {
    LAUNCH_V(1) LAUNCH_V(2) LAUNCH_V(3) LAUNCH_V(4) LAUNCH_V(5) LAUNCH_V(6) LAUNCH_V(7) LAUNCH_V(8)
    default:
      set_error("n_views out of range", cudaSuccess);
      return GPNERF_E_ARG;
  }
#undef LAUNCH_V
  return check_launch("k2_project_gather_meanvar");
}

int gpnerf_k2_mean_variance(const float* rgb_feat, int n_views, int n_points, float* meanvar, void* stream) {
  GPNERF_REQUIRE(rgb_feat && meanvar && n_points > 0 && n_views >= 1 && n_views <= GPNERF_MAX_VIEWS);
  long long blocks = ((long long)n_points * 35 + 255) / 256;
  int grid = (int)(blocks < (long long)persistent_grid(8) ? blocks : persistent_grid(8));
  mean_variance<<<grid, 256, 0, (cudaStream_t)stream>>>(rgb_feat, n_views, n_points, meanvar);
  return check_launch("k2_mean_variance");
}

int gpnerf_k2_gather_volume_bwd(float* const d_levels_ndhwc[GPNERF_N_LEVELS], int point_kind, const int32_t* valid,
                                const float* rays_o, const float* rays_d, const float* z_vals, const float* points,
                                const gpnerf_frame_t* f, int n_points_max, const int32_t* counters,
                                const float* d_vol_feat, void* stream) {
  GPNERF_REQUIRE(d_levels_ndhwc && f && d_vol_feat && n_points_max > 0);
  PointSrc ps;
  GPNERF_REQUIRE(make_point_src(&ps, point_kind, valid, rays_o, rays_d, z_vals, points, f));
  LevelGradPtrs lg;
  for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
    GPNERF_REQUIRE(d_levels_ndhwc[l] != nullptr);
    lg.p[l] = d_levels_ndhwc[l];
  }
  long long blocks = ((long long)n_points_max * 8 + 255) / 256;
  int grid = (int)(blocks < (long long)persistent_grid(8) ? blocks : persistent_grid(8));
  scatter_volume_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(lg, ps, *f, counters ? counters + GPNERF_CNT_P1 : nullptr,
                                                             n_points_max, d_vol_feat);
  return check_launch("k2_gather_volume_bwd");
}

int gpnerf_k2_project_gather_bwd(float* d_featmaps_nhwc, int point_kind, const int32_t* valid, const float* rays_o,
                                 const float* rays_d, const float* z_vals, const float* points,
                                 const gpnerf_frame_t* f, int n_points_max, const int32_t* counters,
                                 const float* d_rgb_feat, void* stream) {
  GPNERF_REQUIRE(d_featmaps_nhwc && f && d_rgb_feat && n_points_max > 0);
  PointSrc ps;
  GPNERF_REQUIRE(point_kind != 2 && make_point_src(&ps, point_kind, valid, rays_o, rays_d, z_vals, points, f));
  const int32_t* cp = counters ? counters + GPNERF_CNT_P1 : nullptr;
  long long blocks = ((long long)n_points_max * 8 + 255) / 256;
  int grid = (int)(blocks < (long long)persistent_grid(8) ? blocks : persistent_grid(8));
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_V(VV)                                                                                         \
  case VV:                                                                                                   \
    scatter_features_bwd<VV><<<grid, 256, 0, st>>>(d_featmaps_nhwc, ps, *f, cp, n_points_max, d_rgb_feat);   \
    break;
  switch (f->n_views) {
    LAUNCH_V(1) LAUNCH_V(2) LAUNCH_V(3) LAUNCH_V(4) LAUNCH_V(5) LAUNCH_V(6) LAUNCH_V(7) LAUNCH_V(8)
    default:
      set_error("n_views out of range", cudaSuccess);
      return GPNERF_E_ARG;
  }
#undef LAUNCH_V
  return check_launch("k2_project_gather_bwd");
}

}  // extern "C"
