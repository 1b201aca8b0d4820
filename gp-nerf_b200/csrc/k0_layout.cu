// K0 – re-lay the upstream products for gathering.
//
// The reference samples channel-first volumes/maps with grid_sample
// (SparseConvNet.py:111-122, BaseRender.py:346-356): every trilinear corner
// costs 32 strided 4-byte reads.  Here each voxel / feature pixel becomes one
// contiguous 128-byte line (32 fp32 channels), so a corner is one coalesced
// transaction.  The same pass produces the per-voxel channel sum that
// SparseConvNet.encode reduces into masks3d (SparseConvNet.py:135-139).
// Pure HBM streaming: 4 B read + 4 B written per element.
#include <string.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace gpnerf {

// in: [32][n] (channel-first) → out: [n][32] as fp32 (128-byte lines) or as
// 16-bit bf16 / fp16 (64-byte lines; fp16 saturates at ±65504 and is what the
// fused tensor-core kernel interpolates with HFMA2); optional
// chan_sum[n] = Σ_c in[c][v] (c ascending, fp32).
// CTA = 256 threads, tile = 32 channels × 128 positions: float4 loads along the
// positions (512 B per channel row), transposed through shared memory, 32 B
// (bf16) or 64 B (fp32) stored per thread.
// Where position v = (z*H + y)*W + x of batch b lands in the output: dense, or
// inside a one-element zero border (pad) so that gathers need no bounds tests.
struct OutMap {
  int H, W;                 // source extents of the two fastest dims
  long long sz, sy, off;    // output strides of z and y, offset of (0,0,0)
  long long batch;          // output positions per batch entry
  int pad;
  __device__ __forceinline__ long long operator()(long long b, long long v) const {
    if (!pad) return b * batch + v;
    const int x = (int)(v % W);
    const long long t = v / W;
    const int y = (int)(t % H);
    const long long z = t / H;
    return b * batch + z * sz + (long long)y * sy + x + off;
  }
};

// FMT: 0 fp32, 1 bf16, 2 fp16
template <int FMT>
__global__ void __launch_bounds__(256) to_channels_last_32(const float* __restrict__ in, long long n,
                                                           long long batch_stride_in, OutMap om,
                                                           void* __restrict__ out_v,
                                                           float* __restrict__ chan_sum) {
  constexpr int TV = 128, LD = TV + 4;
  __shared__ __align__(16) float tile[32 * LD];
  const int tid = threadIdx.x;
  const long long n_tiles = (n + TV - 1) / TV;
  const float* src = in + (long long)blockIdx.y * batch_stride_in;
  const bool vec_ok = (n % 4 == 0);
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long v0 = t * TV;
    {
      const int q = tid & 31;            // which float4 of the 128 positions
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = (tid >> 5) + 8 * i;
        const long long v = v0 + 4 * q;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* p = src + (long long)c * n + v;
        if (vec_ok && v + 3 < n) {
          x = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (v < n) x.x = __ldg(p);
          if (v + 1 < n) x.y = __ldg(p + 1);
          if (v + 2 < n) x.z = __ldg(p + 2);
          if (v + 3 < n) x.w = __ldg(p + 3);
        }
        *reinterpret_cast<float4*>(&tile[c * LD + 4 * q]) = x;
      }
    }
    __syncthreads();
    {
      const int vv = tid & 127, h = tid >> 7;     // position, channel half
      const long long v = v0 + vv;
      if (v < n) {
        float x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = tile[(h * 16 + k) * LD + vv];
        if constexpr (FMT == 2) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out_v) + om(blockIdx.y, v) * 32 + h * 16);
          float y[16];    // saturated copy: x itself still feeds the exact channel sum below
#pragma unroll
          for (int k = 0; k < 16; ++k) y[k] = fminf(fmaxf(x[k], -65504.0f), 65504.0f);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            uint4 o;
            __half2 b0 = __floats2half2_rn(y[k * 8 + 0], y[k * 8 + 1]);
            __half2 b1 = __floats2half2_rn(y[k * 8 + 2], y[k * 8 + 3]);
            __half2 b2 = __floats2half2_rn(y[k * 8 + 4], y[k * 8 + 5]);
            __half2 b3 = __floats2half2_rn(y[k * 8 + 6], y[k * 8 + 7]);
            o.x = *reinterpret_cast<uint32_t*>(&b0);
            o.y = *reinterpret_cast<uint32_t*>(&b1);
            o.z = *reinterpret_cast<uint32_t*>(&b2);
            o.w = *reinterpret_cast<uint32_t*>(&b3);
            dst[k] = o;
          }
        } else if constexpr (FMT == 1) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_v) +
                                                om(blockIdx.y, v) * 32 + h * 16);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            uint4 o;
            __nv_bfloat162 b0 = __floats2bfloat162_rn(x[k * 8 + 0], x[k * 8 + 1]);
            __nv_bfloat162 b1 = __floats2bfloat162_rn(x[k * 8 + 2], x[k * 8 + 3]);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(x[k * 8 + 4], x[k * 8 + 5]);
            __nv_bfloat162 b3 = __floats2bfloat162_rn(x[k * 8 + 6], x[k * 8 + 7]);
            o.x = *reinterpret_cast<uint32_t*>(&b0);
            o.y = *reinterpret_cast<uint32_t*>(&b1);
            o.z = *reinterpret_cast<uint32_t*>(&b2);
            o.w = *reinterpret_cast<uint32_t*>(&b3);
            dst[k] = o;
          }
        } else {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_v) +
                                                  om(blockIdx.y, v) * 32 + h * 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) dst[k] = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
        }
        if (chan_sum != nullptr && h == 0) {
          float s = x[0];
#pragma unroll
          for (int k = 1; k < 16; ++k) s = xadd(s, x[k]);
#pragma unroll
          for (int k = 16; k < 32; ++k) s = xadd(s, tile[k * LD + vv]);
          chan_sum[v] = s;
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// One launch for everything the tensor-core path gathers from: the 4 dense
// levels and the V encoder maps, channel-first fp32 → channel-last fp16 inside
// a zero border, plus the per-voxel channel sums of the levels.
//   tile = 32 channels × 256 positions, 256 threads
//   load : warp w reads channels w, w+8, w+16, w+24 – two float4 per lane and
//          channel (1 KB contiguous per warp instruction), 8 loads in flight
//   smem : idx(c, p) = c·256 + (p ^ ((c >> 3) << 3)): float4 writes along p and
//          the 8-channel reads of the store phase are both conflict-free
//   store: 4 lanes per voxel, 16 B each → 8 voxels = 512 contiguous bytes per
//          warp instruction
// ---------------------------------------------------------------------------
constexpr int kMaxJobs = GPNERF_N_LEVELS + 1;
struct LayoutJob {
  const float* src;      // [batch][32][n]
  __half* dst;           // [batch][(D+2)][(H+2)][(W+2)][32] (pad_z) or [batch][(H+2)][(W+2)][32]
  float* chan_sum;       // [n] or NULL
  unsigned n, H, W;      // positions per batch entry, extents of the two fastest dims
  unsigned sz, sy, off;  // padded strides of z, y and offset of element (0,0,0), in voxels
  unsigned batch, batch_stride_out;   // entries, padded voxels per entry
  unsigned tile0, tiles_per_batch;    // first global tile of the job, tiles per batch entry
};
struct LayoutJobs {
  LayoutJob j[kMaxJobs];
  int n_jobs;
  unsigned n_tiles;
};

__global__ void __launch_bounds__(256) products_to_channels_last_f16(const __grid_constant__ LayoutJobs jobs) {
  constexpr int TP = 256;
  __shared__ __align__(16) float tile[32 * TP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto job_of = [&](unsigned t) {
    int ji = 0;
#pragma unroll
    for (int k = 1; k < kMaxJobs; ++k)
      if (k < jobs.n_jobs && t >= jobs.j[k].tile0) ji = k;
    return ji;
  };
  // a tile's 32 channels × 256 positions as 8 float4 per thread (channel warp + 8i, positions hf·128 + 4·lane …)
  auto load_tile = [&](unsigned t, float4 (&x)[8]) {
    const LayoutJob& J = jobs.j[job_of(t)];
    const unsigned lt = t - J.tile0;
    const unsigned b = lt / J.tiles_per_batch;
    const unsigned v0 = (lt - b * J.tiles_per_batch) * TP;
    const float* src = J.src + (size_t)b * 32 * J.n;
    const bool vec_ok = (J.n % 4 == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = warp + 8 * i;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const unsigned v = v0 + hf * 128 + 4 * lane;
        const float* g = src + (size_t)c * J.n + v;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec_ok && v + 3 < J.n) {
          q = __ldg(reinterpret_cast<const float4*>(g));
        } else {
          if (v < J.n) q.x = __ldg(g);
          if (v + 1 < J.n) q.y = __ldg(g + 1);
          if (v + 2 < J.n) q.z = __ldg(g + 2);
          if (v + 3 < J.n) q.w = __ldg(g + 3);
        }
        x[i * 2 + hf] = q;
      }
    }
  };
  float4 x[8];
  if (blockIdx.x < jobs.n_tiles) load_tile(blockIdx.x, x);
  for (unsigned t = blockIdx.x; t < jobs.n_tiles; t += gridDim.x) {
    const LayoutJob& J = jobs.j[job_of(t)];
    const unsigned lt = t - J.tile0;
    const unsigned b = lt / J.tiles_per_batch;
    const unsigned v0 = (lt - b * J.tiles_per_batch) * TP;
    // ---- registers → shared memory (swizzled), then the next tile's loads go out: they are in flight while this
    //      tile is summed and stored (the kernel is bound by DRAM latency × bytes in flight, not by bandwidth)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = warp + 8 * i;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const unsigned p = hf * 128 + 4 * lane;
        *reinterpret_cast<float4*>(&tile[c * TP + (p ^ ((c >> 3) << 3))]) = x[i * 2 + hf];
      }
    }
    if (t + gridDim.x < jobs.n_tiles) load_tile(t + gridDim.x, x);
    __syncthreads();
    // ---- channel sums (exact: c ascending, one rounding per add)
    if (J.chan_sum != nullptr && v0 + tid < J.n) {
      float sum = tile[tid];
#pragma unroll
      for (int c = 1; c < 32; ++c) sum = xadd(sum, tile[c * TP + (tid ^ ((c >> 3) << 3))]);
      J.chan_sum[v0 + tid] = sum;
    }
    // ---- store phase: warp w owns positions [32w, 32w+32), 8 voxels per step
    {
      const int i = lane >> 2, j = lane & 3;
      __half* dst = J.dst + (size_t)b * J.batch_stride_out * 32;
#pragma unroll
      for (int step = 0; step < 4; ++step) {
        const unsigned p = warp * 32 + step * 8 + i;
        const unsigned v = v0 + p;
        if (v < J.n) {
          float y[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = 8 * j + k;
            y[k] = fminf(fmaxf(tile[c * TP + (p ^ (j << 3))], -65504.0f), 65504.0f);
          }
          const unsigned x = v % J.W, r = v / J.W;
          const unsigned yy = r % J.H, zz = r / J.H;
          const size_t o = (size_t)zz * J.sz + (size_t)yy * J.sy + x + J.off;
          uint4 q;
          __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
          __half2 h2 = __floats2half2_rn(y[4], y[5]), h3 = __floats2half2_rn(y[6], y[7]);
          q.x = *reinterpret_cast<uint32_t*>(&h0);
          q.y = *reinterpret_cast<uint32_t*>(&h1);
          q.z = *reinterpret_cast<uint32_t*>(&h2);
          q.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(dst + o * 32 + j * 8) = q;
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// SURVEY §8f row 1, first step: the sparse-conv pyramid's outputs WITHOUT the
// dense detour.  The reference materialises every level with
// SparseConvTensor.dense() (SparseConvNet.py:110; an NCDHW fp32 tensor that is
// >95 % zeros) only to sample it; here the active rows (features [N,32] fp32 +
// voxel indices [N,idx_cols] int32, the last three columns = d,h,w) are scattered
// straight into the zero-bordered channel-last fp16 volumes the fused kernel
// gathers from, together with the per-voxel channel sums of masks3d.
// 8 lanes per row: one float4 load and one 8-byte store per lane.
// ---------------------------------------------------------------------------
struct SparseJob {
  const float* feat;
  const int32_t* idx;
  void* dst;
  float* chan_sum;
  const int32_t* n_dev;   // optional: the live row count sits on the device (n is then the capacity)
  int n, D, H, W;
  int row0;          // first global row of the job
};
struct SparseJobs {
  SparseJob j[GPNERF_N_LEVELS];
  int n_rows, idx_cols;
};
// F32 = true: the same scatter into the fp32 channel-last volumes (no border) of the exact-arithmetic path.
template <bool F32>
__global__ void __launch_bounds__(256) sparse_rows_scatter(const __grid_constant__ SparseJobs jobs) {
  const int lane = threadIdx.x & 31, q = lane & 7, sub = lane >> 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int r0 = warp * 4; r0 < jobs.n_rows; r0 += n_warps * 4) {
    const int r = r0 + sub;
    int ji = 0;
#pragma unroll
    for (int k = 1; k < GPNERF_N_LEVELS; ++k)
      if (r >= jobs.j[k].row0) ji = k;
    const SparseJob& J = jobs.j[ji];
    const int lr = r - J.row0;
    const bool ok = r < jobs.n_rows && lr < (J.n_dev ? min(J.n, __ldg(J.n_dev)) : J.n);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int d = 0, h = 0, w = 0;
    if (ok) {
      v = __ldg(reinterpret_cast<const float4*>(J.feat + (size_t)lr * 32) + q);
      const int32_t* ip = J.idx + (size_t)lr * jobs.idx_cols + (jobs.idx_cols - 3);
      d = __ldg(ip); h = __ldg(ip + 1); w = __ldg(ip + 2);
    }
    const bool inside = ok && d >= 0 && d < J.D && h >= 0 && h < J.H && w >= 0 && w < J.W;
    // channel sum, c ascending with one rounding per add (the order torch.sum(dim=0) and K0 use)
    float sum = 0.0f;
#pragma unroll
    for (int qq = 0; qq < 8; ++qq) {
      const int src = (lane & 24) | qq;
      const float a = __shfl_sync(0xffffffffu, v.x, src), b = __shfl_sync(0xffffffffu, v.y, src);
      const float c = __shfl_sync(0xffffffffu, v.z, src), e = __shfl_sync(0xffffffffu, v.w, src);
      sum = (qq == 0) ? a : xadd(sum, a);
      sum = xadd(xadd(xadd(sum, b), c), e);
    }
    if (inside && F32) {
      const size_t vox = ((size_t)d * J.H + h) * J.W + w;
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(J.dst) + vox * 32 + q * 4) = v;
      if (q == 0) J.chan_sum[vox] = sum;
    } else if (inside) {
      const size_t vox = ((size_t)(d + 1) * (J.H + 2) + (h + 1)) * (J.W + 2) + (w + 1);
      __half2 lo = __floats2half2_rn(fminf(fmaxf(v.x, -65504.f), 65504.f), fminf(fmaxf(v.y, -65504.f), 65504.f));
      __half2 hi = __floats2half2_rn(fminf(fmaxf(v.z, -65504.f), 65504.f), fminf(fmaxf(v.w, -65504.f), 65504.f));
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(J.dst) + vox * 32 + q * 4) = o;
      if (q == 0) J.chan_sum[((size_t)d * J.H + h) * J.W + w] = sum;
    }
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
                                                      long long hw, int unnormalize, OutMap om,
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
    out[om(v, p)] = o;
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

static OutMap make_map(int H, int W, int pad, int pad_z, long long n) {
  OutMap om;
  om.H = H; om.W = W; om.pad = pad;
  om.sy = W + 2;
  om.sz = (long long)(H + 2) * (W + 2);
  om.off = (pad_z ? om.sz : 0) + om.sy + 1;
  om.batch = pad ? 0 : n;   // callers set the padded batch stride
  return om;
}

int gpnerf_k0_level_to_channels_last(const float* ncdhw, int D, int H, int W, int storage, int pad,
                                     void* ndhwc, float* chan_sum, void* stream) {
  GPNERF_REQUIRE(ncdhw && ndhwc && D > 0 && H > 0 && W > 0);
  long long n = (long long)D * H * W;
  dim3 grid(grid_for((n + 127) / 128, 1), 1);
  OutMap om = make_map(H, W, pad, 1, n);
  GPNERF_REQUIRE(storage >= 0 && storage <= 2);
  if (storage == 2)
    to_channels_last_32<2><<<grid, 256, 0, (cudaStream_t)stream>>>(ncdhw, n, 0, om, ndhwc, chan_sum);
  else if (storage == 1)
    to_channels_last_32<1><<<grid, 256, 0, (cudaStream_t)stream>>>(ncdhw, n, 0, om, ndhwc, chan_sum);
  else
    to_channels_last_32<0><<<grid, 256, 0, (cudaStream_t)stream>>>(ncdhw, n, 0, om, ndhwc, chan_sum);
  return check_launch("k0_level_to_channels_last");
}

int gpnerf_k0_products_to_f16(const float* const levels[GPNERF_N_LEVELS], const int32_t level_dims[GPNERF_N_LEVELS][3],
                              const float* featmaps, int V, int fh, int fw, void* const levels_out[GPNERF_N_LEVELS],
                              float* const chan_sums[GPNERF_N_LEVELS], void* featmaps_out, void* stream) {
  GPNERF_REQUIRE(levels && level_dims && levels_out && chan_sums);
  GPNERF_REQUIRE(featmaps == nullptr || (featmaps_out && V > 0 && V <= GPNERF_MAX_VIEWS && fh > 0 && fw > 0));
  LayoutJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  unsigned tile = 0;
  int nj = 0;
  for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
    GPNERF_REQUIRE(levels[l] && levels_out[l] && chan_sums[l]);
    const int D = level_dims[l][0], H = level_dims[l][1], W = level_dims[l][2];
    GPNERF_REQUIRE(D > 0 && H > 0 && W > 0 && (long long)(D + 2) * (H + 2) * (W + 2) < (1ll << 31));
    LayoutJob& J = jobs.j[nj++];
    J.src = levels[l]; J.dst = reinterpret_cast<__half*>(levels_out[l]); J.chan_sum = chan_sums[l];
    J.n = (unsigned)D * H * W; J.H = H; J.W = W;
    J.sy = W + 2; J.sz = (unsigned)(H + 2) * (W + 2); J.off = J.sz + J.sy + 1;
    J.batch = 1; J.batch_stride_out = 0;
    J.tile0 = tile; J.tiles_per_batch = (J.n + 255) / 256;
    tile += J.tiles_per_batch;
  }
  if (featmaps != nullptr) {
    LayoutJob& J = jobs.j[nj++];
    J.src = featmaps; J.dst = reinterpret_cast<__half*>(featmaps_out); J.chan_sum = nullptr;
    J.n = (unsigned)fh * fw; J.H = fh; J.W = fw;
    J.sy = fw + 2; J.sz = 0; J.off = J.sy + 1;
    J.batch = V; J.batch_stride_out = (unsigned)(fh + 2) * (fw + 2);
    J.tile0 = tile; J.tiles_per_batch = (J.n + 255) / 256;
    tile += J.tiles_per_batch * V;
  }
  jobs.n_jobs = nj;
  jobs.n_tiles = tile;
  const int cap = sm_count() * 8;
  const int grid = (int)(tile < (unsigned)cap ? tile : (unsigned)cap);
  products_to_channels_last_f16<<<grid, 256, 0, (cudaStream_t)stream>>>(jobs);
  return check_launch("k0_products_to_f16");
}

static int sparse_scatter_launch(bool f32, const float* const feats[GPNERF_N_LEVELS], const int32_t* const indices[GPNERF_N_LEVELS],
                            const int32_t n_rows[GPNERF_N_LEVELS], const int32_t* const n_rows_dev[GPNERF_N_LEVELS],
                            int idx_cols, const int32_t level_dims[GPNERF_N_LEVELS][3],
                            void* const levels_out[GPNERF_N_LEVELS], float* const chan_sums[GPNERF_N_LEVELS],
                            void* stream) {
  const int pad = f32 ? 0 : 1;
  const size_t voxel_bytes = f32 ? 128 : 64;
  GPNERF_REQUIRE(feats && indices && n_rows && level_dims && levels_out && chan_sums && idx_cols >= 3 && idx_cols <= 4);
  cudaStream_t st = (cudaStream_t)stream;
  SparseJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  int row = 0;
  for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
    const int D = level_dims[l][0], H = level_dims[l][1], W = level_dims[l][2];
    GPNERF_REQUIRE(levels_out[l] && chan_sums[l] && n_rows[l] >= 0 && (n_rows[l] == 0 || (feats[l] && indices[l])));
    GPNERF_REQUIRE(D > 0 && H > 0 && W > 0 && (long long)(D + 2 * pad) * (H + 2 * pad) * (W + 2 * pad) < (1ll << 31));
    // yesterday's active sites go away with the whole volume: 64 B per voxel at HBM speed (≈10 µs for all levels)
    cudaError_t e = cudaMemsetAsync(levels_out[l], 0, (size_t)(D + 2 * pad) * (H + 2 * pad) * (W + 2 * pad) * voxel_bytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(chan_sums[l], 0, (size_t)D * H * W * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("memset sparse level", e);
      return GPNERF_E_CUDA;
    }
    SparseJob& J = jobs.j[l];
    J.feat = feats[l]; J.idx = indices[l]; J.dst = levels_out[l]; J.chan_sum = chan_sums[l];
    J.n_dev = n_rows_dev ? n_rows_dev[l] : nullptr;
    J.n = n_rows[l]; J.D = D; J.H = H; J.W = W; J.row0 = row;
    row += (n_rows[l] + 3) & ~3;          // a warp's 4 rows never straddle two levels
  }
  jobs.n_rows = row;
  jobs.idx_cols = idx_cols;
  if (row == 0) return GPNERF_OK;
  const long long blocks = ((long long)row * 8 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (f32)
    sparse_rows_scatter<true><<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(jobs);
  else
    sparse_rows_scatter<false><<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(jobs);
  return check_launch("k0_sparse_scatter");
}

int gpnerf_k0_sparse_to_f16(const float* const feats[GPNERF_N_LEVELS], const int32_t* const indices[GPNERF_N_LEVELS],
                            const int32_t n_rows[GPNERF_N_LEVELS], const int32_t* const n_rows_dev[GPNERF_N_LEVELS],
                            int idx_cols, const int32_t level_dims[GPNERF_N_LEVELS][3],
                            void* const levels_out[GPNERF_N_LEVELS], float* const chan_sums[GPNERF_N_LEVELS],
                            void* stream) {
  return sparse_scatter_launch(false, feats, indices, n_rows, n_rows_dev, idx_cols, level_dims, levels_out, chan_sums,
                               stream);
}

int gpnerf_k0_sparse_to_f32(const float* const feats[GPNERF_N_LEVELS], const int32_t* const indices[GPNERF_N_LEVELS],
                            const int32_t n_rows[GPNERF_N_LEVELS], const int32_t* const n_rows_dev[GPNERF_N_LEVELS],
                            int idx_cols, const int32_t level_dims[GPNERF_N_LEVELS][3],
                            void* const levels_out[GPNERF_N_LEVELS], float* const chan_sums[GPNERF_N_LEVELS],
                            void* stream) {
  return sparse_scatter_launch(true, feats, indices, n_rows, n_rows_dev, idx_cols, level_dims, levels_out, chan_sums,
                               stream);
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

int gpnerf_k0_featmaps_to_channels_last(const float* nchw, int V, int h, int w, int storage, int pad,
                                        void* nhwc, void* stream) {
  GPNERF_REQUIRE(nchw && nhwc && V > 0 && V <= GPNERF_MAX_VIEWS && h > 0 && w > 0);
  long long n = (long long)h * w;
  dim3 grid(grid_for((n + 127) / 128, 1), V);
  OutMap om = make_map(h, w, pad, 0, n);
  if (pad) om.batch = (long long)(h + 2) * (w + 2);
  GPNERF_REQUIRE(storage >= 0 && storage <= 2);
  if (storage == 2)
    to_channels_last_32<2><<<grid, 256, 0, (cudaStream_t)stream>>>(nchw, n, 32 * n, om, nhwc, nullptr);
  else if (storage == 1)
    to_channels_last_32<1><<<grid, 256, 0, (cudaStream_t)stream>>>(nchw, n, 32 * n, om, nhwc, nullptr);
  else
    to_channels_last_32<0><<<grid, 256, 0, (cudaStream_t)stream>>>(nchw, n, 32 * n, om, nhwc, nullptr);
  return check_launch("k0_featmaps_to_channels_last");
}

int gpnerf_k0_images_to_rgbx(const float* nchw, int V, int H, int W, int unnormalize, int pad,
                             float* rgbx, void* stream) {
  GPNERF_REQUIRE(nchw && rgbx && V > 0 && V <= GPNERF_MAX_VIEWS && H > 0 && W > 0);
  long long hw = (long long)H * W;
  images_to_rgbx<<<grid_for(V * hw, 256), 256, 0, (cudaStream_t)stream>>>(
      nchw, V, hw, unnormalize, [&] {
        OutMap om = make_map(H, W, pad, 0, hw);
        if (pad) om.batch = (long long)(H + 2) * (W + 2);
        return om;
      }(), reinterpret_cast<float4*>(rgbx));
  return check_launch("k0_images_to_rgbx");
}

}  // extern "C"
