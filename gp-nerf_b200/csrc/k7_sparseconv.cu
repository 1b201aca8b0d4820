// K7 – SURVEY §8f row 1: the sparse-conv geometry encoder (libs/nerfheads/networks/SparseConvNet.py:21-124)
// without spconv.  The reference builds it from spconv 1.2.1 (git abf0acf3, README.md:27-33): SubMConv3d
// (3x3x3, output sites = input sites) and SparseConv3d (3x3x3, stride 2, padding 1, output sites = every
// site reached by an input site), both bias-free, each followed by BatchNorm1d over the active sites and
// ReLU.  spconv is not in the reference tree and does not build for sm_100; its published semantics – the
// convolution equals a dense cross-correlation restricted to the active output sites, weights stored
// [kd][kh][kw][in][out] – are what this file implements (oracle: dense conv3d emulation,
// oracle/gpnerf_oracle.py sparse_conv_net; parity against spconv itself is unpinned).
//
// Inference form: BatchNorm folded into a per-channel scale/shift by the caller (running statistics).
// Data: per level a row list (coords [n][3] = d,h,w; features [n][C] fp32) plus a dense int32 index volume
// (row id of the site, "none" = 0x7f7f7f7f) for neighbour look-ups.  All row counts stay on the device.
//
//   sc_index_input    coords of the SMPL voxels → de-duplicated ascending site list + index volume
//                     (several vertices fall into one 5 mm voxel; the smallest row id owns the site)
//   sc_strided_sites  active sites of the next (stride-2) level from the current ones
//   sc_neighbours     per convolution geometry: the 27 neighbour rows of every output site (tap-major table)
//   sc_conv           gather (27 taps through the table) – GEMM – scale/shift – ReLU, fp32 FFMA, cp.async stages
#include <string.h>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace gpnerf {

// index volumes are filled by cudaMemset(0x7f): an untouched voxel reads 0x7f7f7f7f, which is >= any row count and
// therefore "no site" for sc_neighbours

struct Dims3 {
  int D, H, W;
};

__global__ void __launch_bounds__(256) sc_claim_sites(const int32_t* __restrict__ coords, int cols, int n, Dims3 g,
                                                      int32_t* __restrict__ idx_vol) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int32_t* c = coords + (size_t)i * cols + (cols - 3);
    const int d = __ldg(c), h = __ldg(c + 1), w = __ldg(c + 2);
    if (d >= 0 && d < g.D && h >= 0 && h < g.H && w >= 0 && w < g.W) atomicMin(idx_vol + ((size_t)d * g.H + h) * g.W + w, i);
  }
}
// one flag per input row: does it own its voxel?
__global__ void __launch_bounds__(256) sc_owner_flags(const int32_t* __restrict__ coords, int cols, int n, Dims3 g,
                                                      const int32_t* __restrict__ idx_vol, uint32_t* __restrict__ words) {
  const int n_pad = (n + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
    bool own = false;
    if (i < n) {
      const int32_t* c = coords + (size_t)i * cols + (cols - 3);
      const int d = __ldg(c), h = __ldg(c + 1), w = __ldg(c + 2);
      own = d >= 0 && d < g.D && h >= 0 && h < g.H && w >= 0 && w < g.W &&
            __ldg(idx_vol + ((size_t)d * g.H + h) * g.W + w) == i;
    }
    const unsigned b = __ballot_sync(0xffffffffu, own);
    if ((threadIdx.x & 31) == 0) words[i >> 5] = b;
  }
}
// owners (ascending input row) → their coords, and the index volume re-labelled with the compact row id
__global__ void __launch_bounds__(256) sc_finish_input(const int32_t* __restrict__ coords, int cols, Dims3 g,
                                                       const int32_t* __restrict__ owners, const int32_t* __restrict__ n_own,
                                                       int32_t* __restrict__ coords_out, int32_t* __restrict__ idx_vol) {
  const int n = __ldg(n_own);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int32_t* c = coords + (size_t)__ldg(owners + j) * cols + (cols - 3);
    const int d = __ldg(c), h = __ldg(c + 1), w = __ldg(c + 2);
    coords_out[j * 3 + 0] = d; coords_out[j * 3 + 1] = h; coords_out[j * 3 + 2] = w;
    idx_vol[((size_t)d * g.H + h) * g.W + w] = j;
  }
}
__global__ void __launch_bounds__(256) sc_gather_rows(const float* __restrict__ in, int C, const int32_t* __restrict__ rows,
                                                      const int32_t* __restrict__ n_dev, float* __restrict__ out) {
  const long long n = (long long)__ldg(n_dev) * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / C), c = (int)(t - (long long)j * C);
    out[t] = __ldg(in + (size_t)__ldg(rows + j) * C + c);
  }
}

// SparseConv3d(3, stride 2, padding 1) site rule: input site i reaches output o = (i + 1 - k) / 2 for every tap k
// in {0,1,2}^3 for which the division is exact and o lies inside the output grid.
__global__ void __launch_bounds__(256) sc_mark_strided(const int32_t* __restrict__ in_coords, const int32_t* __restrict__ n_in,
                                                       Dims3 go, uint32_t* __restrict__ words) {
  const int n = __ldg(n_in);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * 27; t += gridDim.x * blockDim.x) {
    const int i = t / 27, k = t - i * 27;
    const int kd = k / 9, kh = (k / 3) % 3, kw = k % 3;
    const int d = __ldg(in_coords + i * 3) + 1 - kd, h = __ldg(in_coords + i * 3 + 1) + 1 - kh;
    const int w = __ldg(in_coords + i * 3 + 2) + 1 - kw;
    if (d < 0 || h < 0 || w < 0 || (d & 1) || (h & 1) || (w & 1)) continue;
    const int od = d >> 1, oh = h >> 1, ow = w >> 1;
    if (od >= go.D || oh >= go.H || ow >= go.W) continue;
    const unsigned lin = (unsigned)((od * go.H + oh) * go.W + ow);
    atomicOr(words + (lin >> 5), 1u << (lin & 31));
  }
}
__global__ void __launch_bounds__(256) sc_finish_sites(const int32_t* __restrict__ lin, const int32_t* __restrict__ n_dev,
                                                       Dims3 g, int32_t* __restrict__ coords_out,
                                                       int32_t* __restrict__ idx_vol) {
  const int n = __ldg(n_dev);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int v = __ldg(lin + j);
    coords_out[j * 3 + 0] = v / (g.H * g.W);
    coords_out[j * 3 + 1] = (v / g.W) % g.H;
    coords_out[j * 3 + 2] = v % g.W;
    idx_vol[v] = j;
  }
}

// Neighbour table of a convolution: nbr[k][o] = input row under tap k of output site o (−1: inactive or outside),
// nbr_k(o) = o·stride − 1 + k.  Tap-major so the convolution reads a tile's 64 ids per tap coalesced.  Built
// once per (site list, stride) and shared by the convolutions on it (both layers of a double_conv).
__global__ void __launch_bounds__(256) sc_neighbours(const int32_t* __restrict__ out_coords, const int32_t* __restrict__ n_out_dev,
                                                     int n_out_max, int stride, const int32_t* __restrict__ in_idx, Dims3 gi,
                                                     const int32_t* __restrict__ n_in_dev, int32_t* __restrict__ nbr) {
  const int n_out = __ldg(n_out_dev), n_in = __ldg(n_in_dev);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_out * 27; t += gridDim.x * blockDim.x) {
    const int k = t / n_out, o = t - k * n_out;
    const int d = __ldg(out_coords + o * 3) * stride - 1 + k / 9;
    const int h = __ldg(out_coords + o * 3 + 1) * stride - 1 + (k / 3) % 3;
    const int w = __ldg(out_coords + o * 3 + 2) * stride - 1 + k % 3;
    int row = -1;
    if (d >= 0 && d < gi.D && h >= 0 && h < gi.H && w >= 0 && w < gi.W) {
      row = __ldg(in_idx + ((size_t)d * gi.H + h) * gi.W + w);
      if (row < 0 || row >= n_in) row = -1;
    }
    nbr[(size_t)k * n_out_max + o] = row;
  }
}

// out[o][co] = relu(scale[co] · Σ_k Σ_ci W[k][ci][co] · in[nbr[k][o]][ci] + shift[co])        (fp32 FFMA)
// A cluster of 3 CTAs owns a tile of 64 output sites × COUT channels; CTA r of the cluster multiplies the nine
// taps kd = r (the pyramid's lower levels have only 10–250 tiles: splitting the taps over the cluster puts
// 3× the CTAs on the machine and cuts the serial tap chain from 27 to 9), then CTAs 1 and 2 park their
// partial sums in shared memory and CTA 0 adds them over DSMEM in a fixed order (deterministic) and writes
// the rows.  Within a CTA: 128 threads, each RM sites × 4 channels; per active tap the 64 neighbour rows
// (16-byte cp.async, zero-filled where there is no neighbour) and W[k] land in one of kScStages shared-memory
// stages while earlier taps are being multiplied; taps no site of the tile has are skipped.  A rows are read
// by LDS.128 broadcast (the lanes of a row group share an address), W rows as 128 contiguous bytes per
// quarter-warp: no bank conflicts.
constexpr int kScTile = 64, kScStages = 3, kScThreads = 128, kScSplit = 3, kScTaps = 27 / kScSplit;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int n = valid ? 16 : 0;                        // src-size 0: 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(n));
}

template <int CIN, int COUT>
__global__ void __cluster_dims__(kScSplit, 1, 1) __launch_bounds__(kScThreads)
    sc_conv(const float* __restrict__ in_feat, const int32_t* __restrict__ nbr, int n_out_max,
            const int32_t* __restrict__ n_out_dev, const float* __restrict__ W, const float* __restrict__ scale,
            const float* __restrict__ shift, float* __restrict__ out_feat) {
  constexpr int TN = COUT / 4, TM = kScThreads / TN, RM = kScTile / TM;
  static_assert(COUT <= CIN * kScStages, "partial sums are parked in the A stages");
  __shared__ __align__(16) float As[kScStages][kScTile][CIN];
  __shared__ __align__(16) float Ws[kScStages][CIN][COUT];
  __shared__ int ids[kScTaps][kScTile];
  __shared__ int tap_any[kScTaps];
  __shared__ int taps[kScTaps];
  __shared__ int n_taps_s;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, tn = tid % TN, tm = tid / TN;
  const int n_out = __ldg(n_out_dev);
  const int n_tiles = (n_out + kScTile - 1) / kScTile;
  float* park = &As[0][0][0];                          // [kScTile][COUT] partial sums of this CTA
  for (int tile = blockIdx.x / kScSplit; tile < n_tiles; tile += gridDim.x / kScSplit) {
    const int o0 = tile * kScTile;
    if (tid < kScTaps) tap_any[tid] = 0;
    __syncthreads();
    for (int t = tid; t < kScTaps * kScTile; t += kScThreads) {
      const int kk = t / kScTile, r = t - kk * kScTile;
      const int row = o0 + r < n_out ? __ldg(nbr + (size_t)(crank * kScTaps + kk) * n_out_max + o0 + r) : -1;
      ids[kk][r] = row;
      if (row >= 0) tap_any[kk] = 1;                   // benign race: every writer stores 1
    }
    __syncthreads();
    if (tid == 0) {
      int n = 0;
      for (int kk = 0; kk < kScTaps; ++kk)
        if (tap_any[kk]) taps[n++] = kk;
      n_taps_s = n;
    }
    __syncthreads();
    const int n_taps = n_taps_s;
    auto stage_in = [&](int j) {                       // tap taps[j] → stage j % kScStages
      if (j < n_taps) {
        const int kk = taps[j], sidx = j % kScStages;
        constexpr int CH = CIN / 4;                    // 16-byte chunks per row
        for (int t = tid; t < kScTile * CH; t += kScThreads) {
          const int r = t / CH, c = t - r * CH;
          const int row = ids[kk][r];
          cp_async16(&As[sidx][r][c * 4], in_feat + (size_t)(row >= 0 ? row : 0) * CIN + c * 4, row >= 0);
        }
        const float* wk = W + (size_t)(crank * kScTaps + kk) * CIN * COUT;
        for (int t = tid; t < CIN * COUT / 4; t += kScThreads) cp_async16(&Ws[sidx][0][0] + t * 4, wk + t * 4, true);
      }
      asm volatile("cp.async.commit_group;");          // one group per call, empty past the last tap
    };
    float acc[RM][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
    for (int j = 0; j < kScStages - 1; ++j) stage_in(j);
    for (int j = 0; j < n_taps; ++j) {
      stage_in(j + kScStages - 1);                     // its stage was read at iteration j-1 (barrier below)
      asm volatile("cp.async.wait_group %0;" ::"n"(kScStages - 1));
      __syncthreads();
      const int sidx = j % kScStages;
#pragma unroll
      for (int ci = 0; ci < CIN; ci += 4) {
        float4 a[RM], w[4];
#pragma unroll
        for (int i = 0; i < RM; ++i) a[i] = *reinterpret_cast<const float4*>(&As[sidx][tm * RM + i][ci]);
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const float4*>(&Ws[sidx][ci + q][tn * 4]);
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[i][0] = fmaf(av[q], w[q].x, acc[i][0]);
            acc[i][1] = fmaf(av[q], w[q].y, acc[i][1]);
            acc[i][2] = fmaf(av[q], w[q].z, acc[i][2]);
            acc[i][3] = fmaf(av[q], w[q].w, acc[i][3]);
          }
        }
      }
      __syncthreads();
    }
    asm volatile("cp.async.wait_group 0;");
    if (crank != 0) {
#pragma unroll
      for (int i = 0; i < RM; ++i)
        *reinterpret_cast<float4*>(park + (tm * RM + i) * COUT + tn * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    cluster.sync();                                    // partial sums of CTAs 1, 2 are visible
    if (crank == 0) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + tn), sh = __ldg(reinterpret_cast<const float4*>(shift) + tn);
      const float* p1 = cluster.map_shared_rank(park, 1);
      const float* p2 = cluster.map_shared_rank(park, 2);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        const int o = o0 + tm * RM + i;
        if (o < n_out) {
          const float4 b1 = *reinterpret_cast<const float4*>(p1 + (tm * RM + i) * COUT + tn * 4);
          const float4 b2 = *reinterpret_cast<const float4*>(p2 + (tm * RM + i) * COUT + tn * 4);
          float4 y;
          y.x = fmaxf(fmaf((acc[i][0] + b1.x) + b2.x, sc.x, sh.x), 0.0f);
          y.y = fmaxf(fmaf((acc[i][1] + b1.y) + b2.y, sc.y, sh.y), 0.0f);
          y.z = fmaxf(fmaf((acc[i][2] + b1.z) + b2.z, sc.z, sh.z), 0.0f);
          y.w = fmaxf(fmaf((acc[i][3] + b1.w) + b2.w, sc.w, sh.w), 0.0f);
          *reinterpret_cast<float4*>(out_feat + (size_t)o * COUT + tn * 4) = y;
        }
      }
    }
    cluster.sync();                                    // parked sums were read: the stages are free again
  }
}

static int grid1d(long long items, int per_block) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 8;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

}  // namespace gpnerf

using namespace gpnerf;

extern "C" {

int gpnerf_sc_index_input(const int32_t* coords, int cols, int n, int D, int H, int W, int32_t* idx_vol,
                          int32_t* owners, int32_t* coords_out, int32_t* n_out, void* workspace, void* stream) {
  GPNERF_REQUIRE(coords && idx_vol && owners && coords_out && n_out && workspace && n > 0 && cols >= 3 && cols <= 4);
  GPNERF_REQUIRE(D > 0 && H > 0 && W > 0 && (long long)D * H * W < (1ll << 31));
  cudaStream_t st = (cudaStream_t)stream;
  const Dims3 g{D, H, W};
  cudaError_t e = cudaMemsetAsync(idx_vol, 0x7f, (size_t)D * H * W * sizeof(int32_t), st);
  if (e != cudaSuccess) {
    set_error("memset index volume", e);
    return GPNERF_E_CUDA;
  }
  CompactWs ws = carve_workspace(workspace, n);
  sc_claim_sites<<<grid1d(n, 256), 256, 0, st>>>(coords, cols, n, g, idx_vol);
  sc_owner_flags<<<grid1d(n, 256), 256, 0, st>>>(coords, cols, n, g, idx_vol, ws.words);
  int rc = compact_launch(ws, nullptr, 1, n, n, owners, n_out, st);
  if (rc != GPNERF_OK) return rc;
  sc_finish_input<<<grid1d(n, 256), 256, 0, st>>>(coords, cols, g, owners, n_out, coords_out, idx_vol);
  return check_launch("sc_index_input");
}

int gpnerf_sc_gather_rows(const float* feat_in, int C, const int32_t* rows, const int32_t* n_dev, int n_max,
                          float* feat_out, void* stream) {
  GPNERF_REQUIRE(feat_in && rows && n_dev && feat_out && C > 0 && n_max > 0);
  sc_gather_rows<<<grid1d((long long)n_max * C, 256), 256, 0, (cudaStream_t)stream>>>(feat_in, C, rows, n_dev, feat_out);
  return check_launch("sc_gather_rows");
}

int gpnerf_sc_strided_sites(const int32_t* in_coords, const int32_t* n_in_dev, int n_in_max, int Do, int Ho, int Wo,
                            int32_t* out_lin, int32_t* out_coords, int32_t* out_idx_vol, int32_t* n_out_dev,
                            void* workspace, void* stream) {
  GPNERF_REQUIRE(in_coords && n_in_dev && out_lin && out_coords && out_idx_vol && n_out_dev && workspace && n_in_max > 0);
  GPNERF_REQUIRE(Do > 0 && Ho > 0 && Wo > 0 && (long long)Do * Ho * Wo < (1ll << 31));
  cudaStream_t st = (cudaStream_t)stream;
  const Dims3 g{Do, Ho, Wo};
  const long long nv = (long long)Do * Ho * Wo;
  CompactWs ws = carve_workspace(workspace, nv);
  cudaError_t e = cudaMemsetAsync(ws.words, 0, (size_t)((nv + 31) / 32) * sizeof(uint32_t), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(out_idx_vol, 0x7f, (size_t)nv * sizeof(int32_t), st);
  if (e != cudaSuccess) {
    set_error("memset strided sites", e);
    return GPNERF_E_CUDA;
  }
  sc_mark_strided<<<grid1d((long long)n_in_max * 27, 256), 256, 0, st>>>(in_coords, n_in_dev, g, ws.words);
  int rc = compact_launch(ws, nullptr, 1, nv, nv, out_lin, n_out_dev, st);
  if (rc != GPNERF_OK) return rc;
  sc_finish_sites<<<grid1d(nv < (long long)n_in_max * 8 ? nv : (long long)n_in_max * 8, 256), 256, 0, st>>>(
      out_lin, n_out_dev, g, out_coords, out_idx_vol);
  return check_launch("sc_strided_sites");
}

int gpnerf_sc_neighbours(const int32_t* out_coords, const int32_t* n_out_dev, int n_out_max, int stride,
                         const int32_t* in_idx_vol, int Di, int Hi, int Wi, const int32_t* n_in_dev, int32_t* nbr,
                         void* stream) {
  GPNERF_REQUIRE(out_coords && n_out_dev && in_idx_vol && n_in_dev && nbr && n_out_max > 0);
  GPNERF_REQUIRE((stride == 1 || stride == 2) && Di > 0 && Hi > 0 && Wi > 0);
  const Dims3 gi{Di, Hi, Wi};
  sc_neighbours<<<grid1d((long long)n_out_max * 27, 256), 256, 0, (cudaStream_t)stream>>>(
      out_coords, n_out_dev, n_out_max, stride, in_idx_vol, gi, n_in_dev, nbr);
  return check_launch("sc_neighbours");
}

int gpnerf_sc_conv(const float* in_feat, int c_in, const int32_t* nbr, const int32_t* n_out_dev, int n_out_max,
                   const float* weight, const float* scale, const float* shift, int c_out, float* out_feat,
                   void* stream) {
  GPNERF_REQUIRE(in_feat && nbr && n_out_dev && weight && scale && shift && out_feat && n_out_max > 0);
  cudaStream_t st = (cudaStream_t)stream;
#define GPNERF_SC(CI, CO)                                                                                      \
  if (c_in == CI && c_out == CO) {                                                                             \
    sc_conv<CI, CO><<<grid1d(n_out_max, kScTile) * kScSplit, kScThreads, 0, st>>>(                             \
        in_feat, nbr, n_out_max, n_out_dev, weight, scale, shift, out_feat);                                   \
    return check_launch("sc_conv");                                                                            \
  }
  GPNERF_SC(16, 16) GPNERF_SC(16, 32) GPNERF_SC(32, 32) GPNERF_SC(32, 16)
#undef GPNERF_SC
  set_error("sc_conv supports channel widths 16 and 32", cudaSuccess);
  return GPNERF_E_UNSUPPORTED;
}

}  // extern "C"
