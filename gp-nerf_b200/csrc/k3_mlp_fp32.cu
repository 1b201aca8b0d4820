// K3 (fp32 parity path) – the density and colour heads on CUDA cores.
//
// libs/nerfheads/trainhead.py:39-41 (128→64 ELU), :102-110 (134→64→32→16→1,
// ELU×3, ReLU), :85-100,128-145 (colour trunk).  One thread owns one sample
// point; layer inputs sit transposed in shared memory (xs[k][point], conflict
// free), weights sit transposed in shared memory (Wt[k][n]) and are read as
// broadcast float4.  Every output neuron is one sequential FMA chain over k
// (first product rounded, bias added last) – the rounding ATen's CPU sgemm
// produces for these shapes – so σ and hence the progressive step's survivor
// list match the oracle bit for bit wherever expm1 agrees.
//
// This is the reference-precision path (precision = 0).  The bf16 tcgen05
// path lives in k3_mlp_tc.cu.
#include "common.cuh"

namespace gpnerf {

constexpr int TP = 128;        // points per CTA tile = threads per CTA
constexpr int XS = TP + 1;     // padded row stride of the transposed tiles

__device__ __forceinline__ float elu1(float x) { return x > 0.0f ? x : expm1f(x); }

// Wt[k][n] ← W[n][k]
__device__ __forceinline__ void load_wt(float* __restrict__ dst, const float* __restrict__ w, int N, int K) {
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    int n = i / K, k = i - n * K;
    dst[k * N + n] = __ldg(w + i);
  }
}
__device__ __forceinline__ void load_vec(float* __restrict__ dst, const float* __restrict__ b, int N) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) dst[i] = __ldg(b + i);
}

// acc[n] = Σ_k x[k]·Wt[k][n]  (sequential FMA chain), x[k] = xs[row(k)][tid]/div
template <int K, int N, class RowFn>
__device__ __forceinline__ void dense(const float* __restrict__ xs, RowFn row, const float* __restrict__ wt,
                                      float (&acc)[N], float div_by) {
  static_assert(N % 4 == 0, "N must be a multiple of 4");
#pragma unroll
  for (int n = 0; n < N; ++n) acc[n] = 0.0f;
  const int tid = threadIdx.x;
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    float x = xs[row(k) * XS + tid];
    if (div_by != 1.0f) x = xdiv(x, div_by);
    const float4* w4 = reinterpret_cast<const float4*>(wt + k * N);
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      float4 w = w4[q];
      acc[4 * q + 0] = xfma(x, w.x, acc[4 * q + 0]);
      acc[4 * q + 1] = xfma(x, w.y, acc[4 * q + 1]);
      acc[4 * q + 2] = xfma(x, w.z, acc[4 * q + 2]);
      acc[4 * q + 3] = xfma(x, w.w, acc[4 * q + 3]);
    }
  }
}

struct Ident {
  __device__ __forceinline__ int operator()(int k) const { return k; }
};

template <int N>
__device__ __forceinline__ void bias_elu_store(float (&acc)[N], const float* __restrict__ b,
                                               float* __restrict__ xs, int row0) {
#pragma unroll
  for (int n = 0; n < N; ++n) xs[(row0 + n) * XS + threadIdx.x] = elu1(xadd(acc[n], b[n]));
}

// coalesced tile load: rows [pt][K] in global → xs[row0+k][pt]
__device__ __forceinline__ void load_tile_T(float* __restrict__ xs, int row0, const float* __restrict__ src,
                                            long long first, int n_valid, int K, const int32_t* __restrict__ index) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int pt = wid; pt < TP; pt += nw) {
    if (pt < n_valid) {
      long long rowi = index ? (long long)__ldg(index + first + pt) : first + pt;
      const float* s = src + rowi * K;
      for (int k = lane; k < K; k += 32) xs[(row0 + k) * XS + pt] = __ldg(s + k);
    } else {
      for (int k = lane; k < K; k += 32) xs[(row0 + k) * XS + pt] = 0.0f;
    }
  }
}

struct DensityW {
  const float *geo_w, *geo_b, *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3;
};

// smem floats: weights 8192+8576+2048+512+16, biases 64+64+32+16+1(+pad), tiles 128·XS + 134·XS
constexpr int DEN_W_FLOATS = 8192 + 8576 + 2048 + 512 + 16 + 64 + 64 + 32 + 16 + 4;
constexpr int DEN_SMEM = (DEN_W_FLOATS + 128 * XS + 134 * XS) * 4;

__global__ void __launch_bounds__(TP, 1) density_mlp_fp32(const float* __restrict__ vol_feat,
                                                          const float* __restrict__ meanvar,
                                                          const float* __restrict__ mask, DensityW w,
                                                          int V, int input_kind, int n_const,
                                                          const int32_t* __restrict__ count_ptr,
                                                          float* __restrict__ sigma,
                                                          float* __restrict__ sigma_feat) {
  extern __shared__ __align__(16) float sm[];
  float* wt_geo = sm;                 // [128][64]
  float* wt0 = wt_geo + 8192;         // [134][64]
  float* wt1 = wt0 + 8576;            // [64][32]
  float* wt2 = wt1 + 2048;            // [32][16]
  float* wt3 = wt2 + 512;             // [16]
  float* b_geo = wt3 + 16;
  float* b0 = b_geo + 64;
  float* b1 = b0 + 64;
  float* b2 = b1 + 32;
  float* b3 = b2 + 16;
  float* bufA = sm + DEN_W_FLOATS;    // 128 rows
  float* bufB = bufA + 128 * XS;      // 134 rows
  if (input_kind == 0) load_wt(wt_geo, w.geo_w, 64, 128);
  load_wt(wt0, w.w0, 64, 134);
  load_wt(wt1, w.w1, 32, 64);
  load_wt(wt2, w.w2, 16, 32);
  load_vec(wt3, w.w3, 16);
  if (input_kind == 0) load_vec(b_geo, w.geo_b, 64);
  load_vec(b0, w.b0, 64);
  load_vec(b1, w.b1, 32);
  load_vec(b2, w.b2, 16);
  load_vec(b3, w.b3, 1);
  const int n = count_ptr ? __ldg(count_ptr) : n_const;
  const int n_tiles = (n + TP - 1) / TP;
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long first = (long long)tile * TP;
    const int n_valid = min(TP, n - (int)first);
    __syncthreads();  // previous tile fully consumed (and weights visible on the first pass)
    if (input_kind == 0) load_tile_T(bufA, 0, vol_feat, first, n_valid, 128, nullptr);
    else load_tile_T(bufB, 0, vol_feat, first, n_valid, 64, nullptr);   // rows already are sigma_feat
    load_tile_T(bufB, 64, meanvar, first, n_valid, 70, nullptr);
    __syncthreads();
    float acc[64];
    if (input_kind == 0) {
      // sigmahead.out_geometry_fc: 128 → 64, ELU
      dense<128, 64>(bufA, Ident(), wt_geo, acc, 1.0f);
      bias_elu_store<64>(acc, b_geo, bufB, 0);
    }
    if (sigma_feat != nullptr && tid < n_valid) {
      // row-major [P1][64]; strided per thread, small
      float* o = sigma_feat + (first + tid) * 64;
#pragma unroll
      for (int k = 0; k < 64; ++k) o[k] = bufB[k * XS + tid];
    }
    // (own column only: no barrier needed between layers)
    dense<134, 64>(bufB, Ident(), wt0, acc, 1.0f);
    bias_elu_store<64>(acc, b0, bufA, 0);
    float acc1[32];
    dense<64, 32>(bufA, Ident(), wt1, acc1, 1.0f);
    bias_elu_store<32>(acc1, b1, bufB, 0);
    float acc2[16];
    dense<32, 16>(bufB, Ident(), wt2, acc2, 1.0f);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float h = elu1(xadd(acc2[k], b2[k]));
      s = (k == 0) ? xmul(h, wt3[0]) : xfma(h, wt3[k], s);
    }
    s = fmaxf(xadd(s, b3[0]), 0.0f);
    if (tid < n_valid) {
      float nv = 0.0f;
      for (int v = 0; v < V; ++v) nv += __ldg(mask + (first + tid) * V + v);
      sigma[first + tid] = (nv < 1.0f) ? 0.0f : s;
    }
  }
}

struct ColorW {
  const float *bw0, *bb0, *bw1, *bb1, *vw0, *vb0, *vw1, *vb1, *rw0, *rb0, *rw1, *rb1, *rw2, *rb2;
};

template <int V>
struct ColorCfg {
  static constexpr int W_FLOATS = 105 * 64 + 64 * 32 + 32 * 32 * 2 + 32 * V * 32 + 32 * 16 + 16 * 4 + 64 +
                                  32 * 4 + 16 + 4;
  static constexpr int IN_ROWS = 70 + 35 * V;
  static constexpr int SMEM = (W_FLOATS + (IN_ROWS + 64) * XS) * 4;
};

template <int V>
__global__ void __launch_bounds__(TP, 1) color_mlp_fp32(const float* __restrict__ rgb_feat,
                                                        const float* __restrict__ meanvar,
                                                        const int32_t* __restrict__ valid1, ColorW w,
                                                        const int32_t* __restrict__ count_ptr, int n_const,
                                                        float* __restrict__ rgb) {
  extern __shared__ __align__(16) float sm[];
  float* wt_b0 = sm;                       // [105][64]
  float* wt_b1 = wt_b0 + 105 * 64;         // [64][32]
  float* wt_v0 = wt_b1 + 64 * 32;          // [32][32]
  float* wt_v1 = wt_v0 + 32 * 32;          // [32][32]
  float* wt_r0 = wt_v1 + 32 * 32;          // [32V][32]
  float* wt_r1 = wt_r0 + 32 * V * 32;      // [32][16]
  float* wt_r2 = wt_r1 + 32 * 16;          // [16][4] (3 used)
  float* bb0 = wt_r2 + 16 * 4;
  float* bb1 = bb0 + 64;
  float* vb0 = bb1 + 32;
  float* vb1 = vb0 + 32;
  float* rb0 = vb1 + 32;
  float* rb1 = rb0 + 32;
  float* rb2 = rb1 + 16;
  float* in = sm + ColorCfg<V>::W_FLOATS;   // rows: [0,70) mean|var, [70+35v, +35) view v
  float* tmp = in + ColorCfg<V>::IN_ROWS * XS;  // 64 rows
  load_wt(wt_b0, w.bw0, 64, 105);
  load_wt(wt_b1, w.bw1, 32, 64);
  load_wt(wt_v0, w.vw0, 32, 32);
  load_wt(wt_v1, w.vw1, 32, 32);
  load_wt(wt_r0, w.rw0, 32, 32 * V);
  load_wt(wt_r1, w.rw1, 16, 32);
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    int k = i >> 2, n = i & 3;
    wt_r2[i] = (n < 3) ? __ldg(w.rw2 + n * 16 + k) : 0.0f;
  }
  load_vec(bb0, w.bb0, 64);
  load_vec(bb1, w.bb1, 32);
  load_vec(vb0, w.vb0, 32);
  load_vec(vb1, w.vb1, 32);
  load_vec(rb0, w.rb0, 32);
  load_vec(rb1, w.rb1, 16);
  load_vec(rb2, w.rb2, 3);
  const int n = count_ptr ? __ldg(count_ptr) : n_const;
  const int n_tiles = (n + TP - 1) / TP;
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long first = (long long)tile * TP;
    const int n_valid = min(TP, n - (int)first);
    __syncthreads();
    load_tile_T(in, 0, meanvar, first, n_valid, 70, valid1);
    load_tile_T(in, 70, rgb_feat, first, n_valid, 35 * V, valid1);
    __syncthreads();
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      const int vrow = 70 + 35 * v;
      float acc[64];
      // base_fc.0 on [mean | var | rgb_feat_v] (trainhead.py:131,139)
      dense<105, 64>(in, [vrow](int k) { return k < 70 ? k : vrow + (k - 70); }, wt_b0, acc, 1.0f);
      bias_elu_store<64>(acc, bb0, tmp, 0);
      float h[32];
      dense<64, 32>(tmp, Ident(), wt_b1, h, 1.0f);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        h[k] = elu1(xadd(h[k], bb1[k]));
        in[(vrow + k) * XS + tid] = h[k];   // view rows are consumed: reuse them for x_v
      }
      // vis_fc on x/V, residual (trainhead.py:140-141)
      float a2[32];
      dense<32, 32>(in, [vrow](int k) { return vrow + k; }, wt_v0, a2, (float)V);
      bias_elu_store<32>(a2, vb0, tmp, 0);
      dense<32, 32>(tmp, Ident(), wt_v1, a2, 1.0f);
#pragma unroll
      for (int k = 0; k < 32; ++k) in[(vrow + k) * XS + tid] = xadd(h[k], elu1(xadd(a2[k], vb1[k])));
    }
    // rgb_fc on the view-major concat (trainhead.py:143)
    float r0[32];
    dense<32 * V, 32>(in, [](int k) { return 70 + 35 * (k >> 5) + (k & 31); }, wt_r0, r0, 1.0f);
    bias_elu_store<32>(r0, rb0, tmp, 0);
    float r1[16];
    dense<32, 16>(tmp, Ident(), wt_r1, r1, 1.0f);
    float out[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float hk = elu1(xadd(r1[k], rb1[k]));
#pragma unroll
      for (int c = 0; c < 3; ++c) out[c] = (k == 0) ? xmul(hk, wt_r2[c]) : xfma(hk, wt_r2[k * 4 + c], out[c]);
    }
    if (tid < n_valid) {
      long long row = valid1 ? (long long)__ldg(valid1 + first + tid) : first + tid;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float y = xadd(out[c], rb2[c]);
        rgb[row * 3 + c] = 1.0f / (1.0f + expf(-y));
      }
    }
  }
}

}  // namespace gpnerf

using namespace gpnerf;

int gpnerf_density_mlp_tc(const float* vol_feat, const float* meanvar, const float* mask,
                          const gpnerf_head_weights_t* w, int n_views, int input_kind, int n_points_max,
                          const int32_t* count_ptr, float* sigma, float* sigma_feat, cudaStream_t st);
int gpnerf_color_mlp_tc(const float* rgb_feat, const float* meanvar, const int32_t* valid1,
                        const gpnerf_head_weights_t* w, int n_views, int n_points_max,
                        const int32_t* count_ptr, float* rgb, cudaStream_t st);

template <int V>
static int launch_color(const float* rgb_feat, const float* meanvar, const int32_t* valid1, ColorW cw,
                        const int32_t* count_ptr, int n_points_max, float* rgb, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(color_mlp_fp32<V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ColorCfg<V>::SMEM);
    if (e != cudaSuccess) {
      set_error("color smem attribute", e);
      return GPNERF_E_CUDA;
    }
    attr_set = true;
  }
  int tiles = (n_points_max + TP - 1) / TP;
  int grid = tiles < sm_count() ? tiles : sm_count();
  color_mlp_fp32<V><<<grid, TP, ColorCfg<V>::SMEM, st>>>(rgb_feat, meanvar, valid1, cw, count_ptr,
                                                         n_points_max, rgb);
  return check_launch("k3_color_mlp");
}

extern "C" {

int gpnerf_k3_density_mlp(const float* vol_feat, int input_kind, const float* meanvar, const float* mask,
                          const gpnerf_head_weights_t* w, int n_views, int n_points_max,
                          const int32_t* counters, int counter_slot, float* sigma, float* sigma_feat,
                          int precision, void* stream) {
  GPNERF_REQUIRE(vol_feat && meanvar && mask && w && sigma && n_points_max > 0);
  GPNERF_REQUIRE(input_kind == 0 || input_kind == 1);
  GPNERF_REQUIRE(counter_slot >= 0 && counter_slot < GPNERF_N_COUNTERS && n_views >= 1 && n_views <= GPNERF_MAX_VIEWS);
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t* count_ptr = counters ? counters + counter_slot : nullptr;
  if (precision == 1)
    return gpnerf_density_mlp_tc(vol_feat, meanvar, mask, w, n_views, input_kind, n_points_max,
                                 count_ptr, sigma, sigma_feat, st);
  GPNERF_REQUIRE(precision == 0);
  DensityW dw{w->geo_w, w->geo_b, w->den_w[0], w->den_b[0], w->den_w[1], w->den_b[1],
              w->den_w[2], w->den_b[2], w->den_w[3], w->den_b[3]};
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(density_mlp_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, DEN_SMEM);
    if (e != cudaSuccess) {
      set_error("density smem attribute", e);
      return GPNERF_E_CUDA;
    }
    attr_set = true;
  }
  int tiles = (n_points_max + TP - 1) / TP;
  int grid = tiles < sm_count() ? tiles : sm_count();
  density_mlp_fp32<<<grid, TP, DEN_SMEM, st>>>(vol_feat, meanvar, mask, dw, n_views, input_kind,
                                               n_points_max, count_ptr, sigma, sigma_feat);
  return check_launch("k3_density_mlp");
}

int gpnerf_k3_color_mlp(const float* rgb_feat, const float* meanvar, const int32_t* valid1,
                        const gpnerf_head_weights_t* w, int n_views, int n_points_max,
                        const int32_t* counters, int counter_slot, float* rgb, int precision,
                        void* stream) {
  GPNERF_REQUIRE(rgb_feat && meanvar && w && rgb && n_points_max > 0);
  GPNERF_REQUIRE(counter_slot >= 0 && counter_slot < GPNERF_N_COUNTERS);
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t* count_ptr = counters ? counters + counter_slot : nullptr;
  if (precision == 1)
    return gpnerf_color_mlp_tc(rgb_feat, meanvar, valid1, w, n_views, n_points_max, count_ptr, rgb, st);
  GPNERF_REQUIRE(precision == 0);
  ColorW cw{w->base_w[0], w->base_b[0], w->base_w[1], w->base_b[1], w->vis_w[0], w->vis_b[0],
            w->vis_w[1], w->vis_b[1], w->rgb_w[0], w->rgb_b[0], w->rgb_w[1], w->rgb_b[1],
            w->rgb_w[2], w->rgb_b[2]};
  switch (n_views) {
    case 1: return launch_color<1>(rgb_feat, meanvar, valid1, cw, count_ptr, n_points_max, rgb, st);
    case 2: return launch_color<2>(rgb_feat, meanvar, valid1, cw, count_ptr, n_points_max, rgb, st);
    case 3: return launch_color<3>(rgb_feat, meanvar, valid1, cw, count_ptr, n_points_max, rgb, st);
    case 4: return launch_color<4>(rgb_feat, meanvar, valid1, cw, count_ptr, n_points_max, rgb, st);
    default:
      set_error("fp32 colour head supports 1..4 source views", cudaSuccess);
      return GPNERF_E_UNSUPPORTED;
  }
}

}  // extern "C"
