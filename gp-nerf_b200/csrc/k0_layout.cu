// K0 – re-lay the upstream products for gathering.
//
// The reference samples channel-first volumes/maps with grid_sample
// (SparseConvNet.py:111-122, BaseRender.py:346-356): every trilinear corner
// costs 32 strided 4-byte reads.  Here each voxel / feature pixel becomes one
// contiguous 128-byte line (32 fp32 channels), so a corner is one coalesced
// transaction.  The same pass produces the per-voxel channel sum that
// SparseConvNet.encode reduces into masks3d (SparseConvNet.py:135-139).
// Pure HBM streaming: 4 B read + 4 B written per element.
#include "common.cuh"

namespace gpnerf {

// in: [32][n] (channel-first), out: [n][32]; optional chan_sum[n] = Σ_c in[c][v]
// (c ascending).  Block (32,8), tile = 32 channels × 32 positions.
__global__ void __launch_bounds__(256) to_channels_last_32(const float* __restrict__ in,
                                                           long long n, long long batch_stride_in,
                                                           float* __restrict__ out,
                                                           float* __restrict__ chan_sum) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long n_tiles = (n + 31) / 32;
  const float* src = in + (long long)blockIdx.y * batch_stride_in;
  float* dst = out + (long long)blockIdx.y * n * 32;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long v0 = t * 32;
#pragma unroll
    for (int c = ty; c < 32; c += 8) {
      long long v = v0 + tx;
      tile[c][tx] = (v < n) ? __ldg(src + (long long)c * n + v) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int vv = ty; vv < 32; vv += 8) {
      long long v = v0 + vv;
      if (v < n) dst[v * 32 + tx] = tile[tx][vv];
    }
    if (chan_sum != nullptr && ty == 0) {
      long long v = v0 + tx;
      if (v < n) {
        float s = tile[0][tx];
#pragma unroll
        for (int c = 1; c < 32; ++c) s = xadd(s, tile[c][tx]);
        chan_sum[v] = s;
      }
    }
    __syncthreads();
  }
}

struct MaskArgs {
  const float* cs[GPNERF_N_LEVELS];
  int dims[GPNERF_N_LEVELS][3];
};

// masks3d[d,h,w] on the level-1 grid = Σ_k cs_k[nearest(d,h,w)]
// nearest source index as F.interpolate(mode='nearest'): floor(dst·in/out).
__global__ void __launch_bounds__(256) build_masks3d(MaskArgs a, float* __restrict__ masks3d) {
  const int D = a.dims[0][0], H = a.dims[0][1], W = a.dims[0][2];
  const long long n = (long long)D * H * W;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n;
       v += (long long)gridDim.x * blockDim.x) {
    int w = (int)(v % W);
    int h = (int)((v / W) % H);
    int d = (int)(v / ((long long)W * H));
    float s = __ldg(a.cs[0] + v);
#pragma unroll
    for (int k = 1; k < GPNERF_N_LEVELS; ++k) {
      const int Dk = a.dims[k][0], Hk = a.dims[k][1], Wk = a.dims[k][2];
      int sd = min((int)floorf(d * ((float)Dk / (float)D)), Dk - 1);
      int sh = min((int)floorf(h * ((float)Hk / (float)H)), Hk - 1);
      int sw = min((int)floorf(w * ((float)Wk / (float)W)), Wk - 1);
      s = xadd(s, __ldg(a.cs[k] + ((long long)sd * Hk + sh) * Wk + sw));
    }
    masks3d[v] = s;
  }
}

// [V][3][H][W] in [-1,1] → [V][H][W][4] = (x·0.5+0.5, 0); separate mul/add like
// the reference's `src_imgs * 0.5 + 0.5` (BaseRender.py:231).
__global__ void __launch_bounds__(256) images_to_rgbx(const float* __restrict__ in, int V,
                                                      long long hw, int unnormalize,
                                                      float4* __restrict__ out) {
  const long long n = (long long)V * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    long long v = i / hw, p = i % hw;
    const float* base = in + v * 3 * hw + p;
    float4 o;
    o.x = __ldg(base);
    o.y = __ldg(base + hw);
    o.z = __ldg(base + 2 * hw);
    if (unnormalize) {
      o.x = xadd(xmul(o.x, 0.5f), 0.5f);
      o.y = xadd(xmul(o.y, 0.5f), 0.5f);
      o.z = xadd(xmul(o.z, 0.5f), 0.5f);
    }
    o.w = 0.0f;
    out[i] = o;
  }
}

static int grid_for(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  long long cap = (long long)sm_count() * 16;
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace gpnerf

using namespace gpnerf;

extern "C" {

int gpnerf_k0_level_to_channels_last(const float* ncdhw, int D, int H, int W, float* ndhwc,
                                     float* chan_sum, void* stream) {
  GPNERF_REQUIRE(ncdhw && ndhwc && D > 0 && H > 0 && W > 0);
  long long n = (long long)D * H * W;
  dim3 grid(grid_for((n + 31) / 32, 1), 1), block(32, 8);
  to_channels_last_32<<<grid, block, 0, (cudaStream_t)stream>>>(ncdhw, n, 0, ndhwc, chan_sum);
  return check_launch("k0_level_to_channels_last");
}

int gpnerf_k0_build_masks3d(const float* const chan_sum[GPNERF_N_LEVELS],
                            const gpnerf_frame_t* f, float* masks3d, void* stream) {
  GPNERF_REQUIRE(chan_sum && f && masks3d);
  MaskArgs a;
  for (int k = 0; k < GPNERF_N_LEVELS; ++k) {
    GPNERF_REQUIRE(chan_sum[k] != nullptr);
    a.cs[k] = chan_sum[k];
    for (int j = 0; j < 3; ++j) a.dims[k][j] = f->level_dims[k][j];
  }
  long long n = (long long)a.dims[0][0] * a.dims[0][1] * a.dims[0][2];
  build_masks3d<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, masks3d);
  return check_launch("k0_build_masks3d");
}

int gpnerf_k0_featmaps_to_channels_last(const float* nchw, int V, int h, int w, float* nhwc,
                                        void* stream) {
  GPNERF_REQUIRE(nchw && nhwc && V > 0 && V <= GPNERF_MAX_VIEWS && h > 0 && w > 0);
  long long n = (long long)h * w;
  dim3 grid(grid_for((n + 31) / 32, 1), V), block(32, 8);
  to_channels_last_32<<<grid, block, 0, (cudaStream_t)stream>>>(nchw, n, 32 * n, nhwc, nullptr);
  return check_launch("k0_featmaps_to_channels_last");
}

int gpnerf_k0_images_to_rgbx(const float* nchw, int V, int H, int W, int unnormalize, float* rgbx,
                             void* stream) {
  GPNERF_REQUIRE(nchw && rgbx && V > 0 && V <= GPNERF_MAX_VIEWS && H > 0 && W > 0);
  long long hw = (long long)H * W;
  images_to_rgbx<<<grid_for(V * hw, 256), 256, 0, (cudaStream_t)stream>>>(
      nchw, V, hw, unnormalize, reinterpret_cast<float4*>(rgbx));
  return check_launch("k0_images_to_rgbx");
}

}  // extern "C"
