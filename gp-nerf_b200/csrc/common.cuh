// Shared device helpers for libgpnerf_b200.so (sm_100a only).
//
// "Exact chain" arithmetic: every value that feeds an integer result of the
// reference (pixel mask, ray list, box hits, occupancy survivors) is computed
// with explicit round-to-nearest intrinsics in the order the reference's torch
// CPU ops round them (oracle/gpnerf_oracle.py header), so nvcc can neither
// contract mul+add into FMA nor reassociate.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gpnerf_abi.h"

namespace gpnerf {

void set_error(const char* what, cudaError_t err);
int check_launch(const char* what);
int sm_count();

#define GPNERF_REQUIRE(cond)            \
  do {                                  \
    if (!(cond)) {                      \
      gpnerf::set_error(#cond, cudaSuccess); \
      return GPNERF_E_ARG;              \
    }                                   \
  } while (0)

__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// Row of a tiny-K matmul the way ATen's CPU sgemm rounds it: first product
// rounded, then one FMA per further term, k ascending.
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  return xfma(a2, b2, xfma(a1, b1, xmul(a0, b0)));
}
__device__ __forceinline__ float dot4(float a0, float b0, float a1, float b1, float a2, float b2,
                                      float a3, float b3) {
  return xfma(a3, b3, dot3(a0, b0, a1, b1, a2, b2));
}
__device__ __forceinline__ float norm3(float x, float y, float z) {
  return __fsqrt_rn(xfma(z, z, xfma(y, y, xmul(x, x))));
}

struct Vec3 {
  float x, y, z;
};

// Every kernel that reads per-frame constants copies gpnerf_frame_t into shared
// memory first: from the device-resident copy when frame.self_dev is set (so a
// captured CUDA graph sees the values of the *current* frame on replay), else
// from its by-value kernel parameter.  Must be reached by all threads.
__device__ __forceinline__ void load_frame(gpnerf_frame_t* dst, const gpnerf_frame_t& param) {
  const uint32_t* src = param.self_dev ? reinterpret_cast<const uint32_t*>(param.self_dev)
                                       : reinterpret_cast<const uint32_t*>(&param);
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  const int nth = blockDim.x * blockDim.y;
  for (int i = tid; i < (int)(sizeof(gpnerf_frame_t) / 4); i += nth) reinterpret_cast<uint32_t*>(dst)[i] = src[i];
  __syncthreads();
}
#define GPNERF_LOAD_FRAME(param)        \
  __shared__ gpnerf_frame_t f_shared__; \
  gpnerf::load_frame(&f_shared__, param); \
  const gpnerf_frame_t& f = f_shared__;

// Depth of sample s on a ray (BaseRender.py:37-48): z = near·(1−t) + far·t,
// optionally jittered inside its stratum with a host-drawn t_rand.
__device__ __forceinline__ float plain_depth(float near, float far, const float* __restrict__ t_vals, int s) {
  float t = __ldg(t_vals + s);
  return xadd(xmul(near, xsub(1.0f, t)), xmul(far, t));
}
__device__ __forceinline__ float sample_depth(float near, float far, const float* __restrict__ t_vals, int s,
                                              int S, const float* __restrict__ t_rand, long long flat) {
  float z = plain_depth(near, far, t_vals, s);
  if (t_rand != nullptr) {
    float lower = z, upper = z;
    if (s > 0) lower = xmul(0.5f, xadd(z, plain_depth(near, far, t_vals, s - 1)));
    if (s < S - 1) upper = xmul(0.5f, xadd(plain_depth(near, far, t_vals, s + 1), z));
    z = xadd(lower, xmul(xsub(upper, lower), __ldg(t_rand + flat)));
  }
  return z;
}

// pts = o + d·z  (separate mul and add)
__device__ __forceinline__ Vec3 point_on_ray(const float* __restrict__ o, const float* __restrict__ d, float z) {
  Vec3 p;
  p.x = xadd(o[0], xmul(d[0], z));
  p.y = xadd(o[1], xmul(d[1], z));
  p.z = xadd(o[2], xmul(d[2], z));
  return p;
}

// world → SMPL frame → normalised grid coordinate → continuous (x,y,z) index
// into a level with dims (Dk,Hk,Wk).  BaseRender.py:52-73 + ATen
// grid_sampler_unnormalize(align_corners): ((c+1)/2)*(size-1).
__device__ __forceinline__ Vec3 world_to_grid(const gpnerf_frame_t& f, Vec3 p) {
  float qx = xsub(p.x, f.Th[0]), qy = xsub(p.y, f.Th[1]), qz = xsub(p.z, f.Th[2]);
  float cx = dot3(qx, f.R[0], qy, f.R[3], qz, f.R[6]);
  float cy = dot3(qx, f.R[1], qy, f.R[4], qz, f.R[7]);
  float cz = dot3(qx, f.R[2], qy, f.R[5], qz, f.R[8]);
  Vec3 g;
  g.x = xsub(xmul(xdiv(xdiv(xsub(cx, f.bounds_min[0]), f.voxel_size[0]), (float)f.out_sh[2]), 2.0f), 1.0f);
  g.y = xsub(xmul(xdiv(xdiv(xsub(cy, f.bounds_min[1]), f.voxel_size[1]), (float)f.out_sh[1]), 2.0f), 1.0f);
  g.z = xsub(xmul(xdiv(xdiv(xsub(cz, f.bounds_min[2]), f.voxel_size[2]), (float)f.out_sh[0]), 2.0f), 1.0f);
  return g;
}
__device__ __forceinline__ float unnormalize(float c, int size) {
  return xmul(xdiv(xadd(c, 1.0f), 2.0f), (float)(size - 1));
}

// Trilinear tap bookkeeping in ATen's grid_sampler_3d order:
// corners (z0,y0,x0),(z0,y0,x1),(z0,y1,x0),(z0,y1,x1),(z1,...) ; weights are
// left-to-right products.
struct Tri {
  int x0, y0, z0;
  float w[8];
  unsigned inb;  // bit c set when corner c lies inside the volume
};
__device__ __forceinline__ Tri trilinear_setup(Vec3 g, int D, int H, int W) {
  float ix = unnormalize(g.x, W), iy = unnormalize(g.y, H), iz = unnormalize(g.z, D);
  float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  Tri t;
  // clamp before the int conversion so far-outside points cannot overflow
  t.x0 = (int)fminf(fmaxf(fx, -2.0f), (float)W + 1.0f);
  t.y0 = (int)fminf(fmaxf(fy, -2.0f), (float)H + 1.0f);
  t.z0 = (int)fminf(fmaxf(fz, -2.0f), (float)D + 1.0f);
  float wx1 = xsub(ix, fx), wx0 = xsub(xadd(fx, 1.0f), ix);
  float wy1 = xsub(iy, fy), wy0 = xsub(xadd(fy, 1.0f), iy);
  float wz1 = xsub(iz, fz), wz0 = xsub(xadd(fz, 1.0f), iz);
  t.w[0] = xmul(xmul(wx0, wy0), wz0);
  t.w[1] = xmul(xmul(wx1, wy0), wz0);
  t.w[2] = xmul(xmul(wx0, wy1), wz0);
  t.w[3] = xmul(xmul(wx1, wy1), wz0);
  t.w[4] = xmul(xmul(wx0, wy0), wz1);
  t.w[5] = xmul(xmul(wx1, wy0), wz1);
  t.w[6] = xmul(xmul(wx0, wy1), wz1);
  t.w[7] = xmul(xmul(wx1, wy1), wz1);
  bool finite = (ix == ix) && (iy == iy) && (iz == iz);
  unsigned inb = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int x = t.x0 + (c & 1), y = t.y0 + ((c >> 1) & 1), z = t.z0 + (c >> 2);
    bool ok = finite && x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D;
    inb |= (ok ? 1u : 0u) << c;
  }
  t.inb = inb;
  return t;
}

// ---------------------------------------------------------------------------
// Ordered stream compaction (ascending index order, deterministic):
//   producer kernel  : one ballot word per 32 items            → words[]
//   compact_tile_sums: popcount per tile of 256 words           → tile_sums[]
//   compact_scan     : single-CTA exclusive scan of tile sums   → tile_offs[], total
//   compact_expand   : per tile, scan words and write indices   → out[]
// The live item count may sit on the device (n_src[0]*mult) or be a constant.
// ---------------------------------------------------------------------------
constexpr int kTileWords = 256;

struct CompactWs {
  uint32_t* words;
  int32_t* tile_sums;
  int32_t* tile_offs;
};
CompactWs carve_workspace(void* ws, int64_t n_items_max);
// launches the three passes; out_count_slot receives the total.  With `row_begin` the items are also
// seen as rows of `row_len` consecutive items (a ray's samples, a tile's pixels) and
// row_begin[r] = number of survivors before row r (CSR offsets, row_begin[n_rows] = total) comes out of
// the same passes.
int compact_launch(const CompactWs& ws, const int32_t* n_src, int mult, int64_t n_const,
                   int64_t n_items_max, int32_t* out_idx, int32_t* out_count, cudaStream_t st,
                   int row_len = 0, int32_t* row_begin = nullptr);

// Live item count of a pass: a device-side counter times `mult`, clamped to the capacity `n_const` the buffers
// were sized for (n_const > 0), or the constant itself.
__device__ __forceinline__ long long live_count(const int32_t* n_src, int mult, long long n_const) {
  if (n_src == nullptr) return n_const;
  const long long n = (long long)__ldg(n_src) * mult;
  return (n_const > 0 && n > n_const) ? n_const : n;
}

}  // namespace gpnerf
