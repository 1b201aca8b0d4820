// K6 – training path (BASELINE configs[3]): layer-wise fp32 kernels for the
// forward with saved activations and for the backward of the two heads
// (trainhead.py:39-41, 85-110, 118-145 under autograd), of
// fused_mean_variance (trainhead.py:20-24) and of the two gathers
// (SparseConvNet.py:111-122, BaseRender.py:346-358: grid_sample backward =
// scatter-add of the weighted upstream gradient into the sampled corners).
//
// First correct training path: every Linear is one launch over all P points
// (row = point), fp32 on CUDA cores, activations round-trip through HBM.  The
// fused / tensor-core kernels of the inference path are not used here; fusing
// the backward the same way is the next step.
//
//   linear_rows   : Y = epi(X·B + bias)          forward layers and ∂L/∂X
//   grad_weights  : dW += dYᵀ·X, db += Σ dY      ∂L/∂W, ∂L/∂b
//   assemble_raw / raw_grad_split : (rgb, masked σ) ↔ their pre-activation gradients
//   meanvar_bwd   : ∂L/∂rgb_feat from ∂L/∂[mean|var]
//   from_channels_last_32 : gradient volumes back to the reference's NC(D)HW
// (the scatter kernels live next to their gathers in k2_gather.cu).
#include "common.cuh"

namespace gpnerf {

enum { EPI_NONE = 0, EPI_ELU = 1, EPI_RELU = 2, EPI_SIGMOID = 3, EPI_MUL_DELU = 4 };

struct LinArgs {
  const float* X; int ldx; int K; float in_scale;   // X[p][k] = X[p*ldx + k] * in_scale
  const float* in_aux; int ld_in_aux;                // optional: X[p][k] *= ELU'(in_aux[p][k])
  const float* W; int ldw; int w_is_kn;              // B[k][n] = w_is_kn ? W[k*ldw+n] : W[n*ldw+k]
  int N;
  const float* bias;                                 // [N] or NULL
  int epi;                                           // EPI_*
  const float* aux; int ld_aux;                      // EPI_MUL_DELU: multiply by ELU'(aux[p][n]) (aux = ELU output)
  float* Y; int ldy;
  int add_pre;                                       // v += Y[p][n] before the epilogue function
  int add_post;                                      // v += Y[p][n] after it (residual / gradient accumulation)
  long long P;
};

constexpr int LTP = 128;          // points per tile = threads per CTA
constexpr int LXS = LTP + 1;

__device__ __forceinline__ float epi_apply(float v, int epi, float aux) {
  switch (epi) {
    case EPI_ELU: return v > 0.0f ? v : expm1f(v);
    case EPI_RELU: return fmaxf(v, 0.0f);
    case EPI_SIGMOID: return 1.0f / (1.0f + expf(-v));
    case EPI_MUL_DELU: return v * (aux > 0.0f ? 1.0f : aux + 1.0f);   // ELU'(pre) from the ELU output
    default: return v;
  }
}

// one thread per point; B (zero-padded to a multiple of 4 columns) and the
// transposed X tile sit in shared memory; NB accumulators per column pass
template <int NB>
__global__ void __launch_bounds__(LTP) linear_rows(LinArgs a) {
  extern __shared__ __align__(16) float lsm[];
  const int K = a.K, N = a.N, NP = (N + 3) & ~3;
  float* bt = lsm;               // [K][NP]
  float* xs = lsm + K * NP;      // [K][LXS]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < K * NP; i += LTP) {
    const int k = i / NP, n = i - k * NP;
    bt[i] = (n < N) ? __ldg(a.w_is_kn ? a.W + (long long)k * a.ldw + n : a.W + (long long)n * a.ldw + k) : 0.0f;
  }
  const long long n_tiles = (a.P + LTP - 1) / LTP;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long first = tile * LTP;
    const int n_valid = (int)min((long long)LTP, a.P - first);
    __syncthreads();
    for (int pt = wid; pt < LTP; pt += LTP / 32) {
      const float* src = a.X + (first + pt) * a.ldx;
      const float* ia = a.in_aux ? a.in_aux + (first + pt) * a.ld_in_aux : nullptr;
      for (int k = lane; k < K; k += 32) {
        float x = 0.0f;
        if (pt < n_valid) {
          x = __ldg(src + k) * a.in_scale;
          if (ia) {
            const float h = __ldg(ia + k);
            x *= (h > 0.0f ? 1.0f : h + 1.0f);
          }
        }
        xs[k * LXS + pt] = x;
      }
    }
    __syncthreads();
    for (int nb = 0; nb < NP; nb += NB) {
      float acc[NB];
#pragma unroll
      for (int n = 0; n < NB; ++n) acc[n] = 0.0f;
#pragma unroll 2
      for (int k = 0; k < K; ++k) {
        const float x = xs[k * LXS + tid];
        const float4* w4 = reinterpret_cast<const float4*>(bt + k * NP + nb);
#pragma unroll
        for (int q = 0; q < NB / 4; ++q) {
          if (nb + 4 * q < NP) {
            const float4 w = w4[q];
            acc[4 * q + 0] = fmaf(x, w.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(x, w.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(x, w.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(x, w.w, acc[4 * q + 3]);
          }
        }
      }
      if (tid < n_valid) {
        float* y = a.Y + (first + tid) * a.ldy;
        const float* ax = a.aux ? a.aux + (first + tid) * a.ld_aux : nullptr;
#pragma unroll
        for (int n = 0; n < NB; ++n) {
          const int col = nb + n;
          if (col < N) {
            float v = acc[n] + (a.bias ? __ldg(a.bias + col) : 0.0f);
            if (a.add_pre) v += y[col];
            v = epi_apply(v, a.epi, ax ? __ldg(ax + col) : 0.0f);
            if (a.add_post) v += y[col];
            y[col] = v;
          }
        }
      }
    }
  }
}

// dW[n][k] (+)= Σ_p dY[p][n]·X[p][k]·in_scale ; db[n] (+)= Σ_p dY[p][n]
// 256 threads = 16 (n) × 16 (k) ; register tile 4 × 9 ; chunks of 64 points
struct GwArgs {
  const float* X; int ldx; int K; float in_scale;
  const float* dY; int ldy; int N;
  const float* dy_aux; int ld_dy_aux;                // optional: dY[p][n] *= ELU'(dy_aux[p][n])
  float* dW; int ldw; float* db; long long P;
};
constexpr int GCH = 64;
__global__ void __launch_bounds__(256) grad_weights(GwArgs a) {
  extern __shared__ __align__(16) float gsm[];
  const int K = a.K, N = a.N;
  float* xs = gsm;                 // [GCH][K+1]
  float* ds = gsm + GCH * (K + 1); // [GCH][N+1]
  const int tid = threadIdx.x, tn = tid >> 4, tk = tid & 15;
  float acc[4][9];
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[i][j] = 0.0f;
  const long long n_chunks = (a.P + GCH - 1) / GCH;
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const long long first = ch * GCH;
    const int n_valid = (int)min((long long)GCH, a.P - first);
    __syncthreads();
    for (int i = tid; i < GCH * K; i += 256) {
      const int p = i / K, k = i - p * K;
      xs[p * (K + 1) + k] = (p < n_valid) ? __ldg(a.X + (first + p) * a.ldx + k) * a.in_scale : 0.0f;
    }
    for (int i = tid; i < GCH * N; i += 256) {
      const int p = i / N, n = i - p * N;
      float d = 0.0f;
      if (p < n_valid) {
        d = __ldg(a.dY + (first + p) * a.ldy + n);
        if (a.dy_aux) {
          const float h = __ldg(a.dy_aux + (first + p) * a.ld_dy_aux + n);
          d *= (h > 0.0f ? 1.0f : h + 1.0f);
        }
      }
      ds[p * (N + 1) + n] = d;
    }
    __syncthreads();
    for (int p = 0; p < GCH; ++p) {
      float dy[4], x[9];
#pragma unroll
      for (int i = 0; i < 4; ++i) dy[i] = (tn + 16 * i < N) ? ds[p * (N + 1) + tn + 16 * i] : 0.0f;
#pragma unroll
      for (int j = 0; j < 9; ++j) x[j] = (tk + 16 * j < K) ? xs[p * (K + 1) + tk + 16 * j] : 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bacc[i] += dy[i];
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[i][j] = fmaf(dy[i], x[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = tn + 16 * i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int k = tk + 16 * j;
      if (k < K) atomicAdd(a.dW + (long long)n * a.ldw + k, acc[i][j]);
    }
    if (a.db != nullptr && tk == 0) atomicAdd(a.db + n, bacc[i]);
  }
}

// raw[p] = (rgb[p], σ[p]) with σ = relu_out[p] forced to 0 where no source view is valid
// (trainhead.py:136-137, 162)
__global__ void __launch_bounds__(256) assemble_raw(const float* __restrict__ rgb, const float* __restrict__ s_relu,
                                                    const float* __restrict__ mask, int V, long long P,
                                                    float4* __restrict__ raw) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    float nv = 0.0f;
    for (int v = 0; v < V; ++v) nv += __ldg(mask + p * V + v);
    raw[p] = make_float4(__ldg(rgb + p * 3), __ldg(rgb + p * 3 + 1), __ldg(rgb + p * 3 + 2),
                         nv < 1.0f ? 0.0f : __ldg(s_relu + p));
  }
}
// its backward: ∂L/∂(pre-sigmoid rgb) and ∂L/∂(pre-ReLU σ)
__global__ void __launch_bounds__(256) raw_grad_split(const float4* __restrict__ d_raw, const float* __restrict__ rgb,
                                                      const float* __restrict__ s_relu,
                                                      const float* __restrict__ mask, int V, long long P,
                                                      float* __restrict__ d_rgb_pre, float* __restrict__ d_s_pre) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const float4 d = __ldg(d_raw + p);
    const float r = __ldg(rgb + p * 3), g = __ldg(rgb + p * 3 + 1), b = __ldg(rgb + p * 3 + 2);
    d_rgb_pre[p * 3 + 0] = d.x * r * (1.0f - r);
    d_rgb_pre[p * 3 + 1] = d.y * g * (1.0f - g);
    d_rgb_pre[p * 3 + 2] = d.z * b * (1.0f - b);
    float nv = 0.0f;
    for (int v = 0; v < V; ++v) nv += __ldg(mask + p * V + v);
    d_s_pre[p] = (nv < 1.0f || !(__ldg(s_relu + p) > 0.0f)) ? 0.0f : d.w;
  }
}

// fused_mean_variance backward: rgb_feat [P][V][35], d_meanvar [P][70] (mean | var) →
// d_rgb_feat[p][v][c] += dmean/V + dvar·2(x−mean)/V
__global__ void __launch_bounds__(256) meanvar_bwd(const float* __restrict__ rgb_feat,
                                                   const float* __restrict__ meanvar,
                                                   const float* __restrict__ d_meanvar, int V, long long P,
                                                   float* __restrict__ d_rgb_feat) {
  const long long total = P * 35;
  const float inv_v = 1.0f / (float)V;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long p = t / 35;
    const int c = (int)(t - p * 35);
    const float m = __ldg(meanvar + p * 70 + c);
    const float dm = __ldg(d_meanvar + p * 70 + c) * inv_v;
    const float dv = __ldg(d_meanvar + p * 70 + 35 + c) * 2.0f * inv_v;
    for (int v = 0; v < V; ++v) {
      const long long idx = (p * V + v) * 35 + c;
      d_rgb_feat[idx] += dm + dv * (__ldg(rgb_feat + idx) - m);
    }
  }
}

// [n][32] (channel-last) → [32][n] (channel-first), fp32 – the inverse of K0
__global__ void __launch_bounds__(256) from_channels_last_32(const float* __restrict__ in, long long n,
                                                             float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long n_tiles = (n + 31) / 32;
  const float* src = in + (long long)blockIdx.y * n * 32;
  float* dst = out + (long long)blockIdx.y * n * 32;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long v0 = t * 32;
#pragma unroll
    for (int vv = ty; vv < 32; vv += 8) tile[vv][tx] = (v0 + vv < n) ? __ldg(src + (v0 + vv) * 32 + tx) : 0.0f;
    __syncthreads();
#pragma unroll
    for (int c = ty; c < 32; c += 8)
      if (v0 + tx < n) dst[(long long)c * n + v0 + tx] = tile[tx][c];
    __syncthreads();
  }
}

static int grid_cap(long long blocks, int per_sm) {
  long long cap = (long long)sm_count() * per_sm;
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace gpnerf

using namespace gpnerf;

namespace gpnerf {   // k6_train_tc.cu
int linear_tc_launch(const float* X, int ldx, int K, float in_scale, const float* in_aux, int ld_in_aux,
                     const float* W, int ldw, int w_is_kn, int N, const float* bias, int epilogue, const float* aux,
                     int ld_aux, float* Y, int ldy, int add_pre, int add_post, long long P, cudaStream_t st);
int grad_weights_tc_launch(const float* X, int ldx, int K, float in_scale, const float* dY, int ldy, int N,
                           const float* dy_aux, int ld_dy_aux, float* dW, int ldw, float* db, long long P,
                           cudaStream_t st);
}  // namespace gpnerf

extern "C" {

int gpnerf_k6_linear(const float* X, int ldx, int K, float in_scale, const float* in_aux, int ld_in_aux,
                     const float* W, int ldw, int w_is_kn, int N, const float* bias, int epilogue, const float* aux,
                     int ld_aux, float* Y, int ldy, int add_pre, int add_post, long long P, int precision,
                     void* stream) {
  GPNERF_REQUIRE(X && W && Y && K > 0 && K <= 160 && N > 0 && N <= 160 && P > 0 && ldx >= K && ldy >= N);
  GPNERF_REQUIRE(epilogue >= 0 && epilogue <= EPI_MUL_DELU && (epilogue != EPI_MUL_DELU || aux != nullptr));
  GPNERF_REQUIRE(precision == 0 || precision == 1);
  if (precision == 1)
    return linear_tc_launch(X, ldx, K, in_scale, in_aux, ld_in_aux, W, ldw, w_is_kn, N, bias, epilogue, aux, ld_aux,
                            Y, ldy, add_pre, add_post, P, (cudaStream_t)stream);
  LinArgs a{X, ldx, K, in_scale, in_aux, ld_in_aux, W, ldw, w_is_kn, N, bias, epilogue, aux, ld_aux, Y, ldy,
            add_pre, add_post, P};
  const int NP = (N + 3) & ~3;
  const size_t smem = (size_t)(K * NP + K * LXS) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_cap((P + LTP - 1) / LTP, 2);
#define GPNERF_LIN(NB)                                                                                         \
  do {                                                                                                         \
    static bool set_##NB = false;                                                                              \
    if (!set_##NB) {                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(linear_rows<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (e != cudaSuccess) { set_error("linear_rows smem attribute", e); return GPNERF_E_CUDA; }             \
      set_##NB = true;                                                                                         \
    }                                                                                                          \
    linear_rows<NB><<<grid, LTP, smem, st>>>(a);                                                               \
  } while (0)
  if (NP <= 4) GPNERF_LIN(4);
  else if (NP <= 16) GPNERF_LIN(16);
  else if (NP <= 32) GPNERF_LIN(32);
  else GPNERF_LIN(64);
#undef GPNERF_LIN
  return check_launch("k6_linear");
}

int gpnerf_k6_grad_weights(const float* X, int ldx, int K, float in_scale, const float* dY, int ldy, int N,
                           const float* dy_aux, int ld_dy_aux, float* dW, int ldw, float* db, long long P,
                           int precision, void* stream) {
  GPNERF_REQUIRE(X && dY && dW && K > 0 && K <= 144 && N > 0 && N <= 64 && P > 0 && ldw >= K);
  GPNERF_REQUIRE(precision == 0 || precision == 1);
  if (precision == 1)
    return grad_weights_tc_launch(X, ldx, K, in_scale, dY, ldy, N, dy_aux, ld_dy_aux, dW, ldw, db, P,
                                  (cudaStream_t)stream);
  GwArgs a{X, ldx, K, in_scale, dY, ldy, N, dy_aux, ld_dy_aux, dW, ldw, db, P};
  const size_t smem = (size_t)(GCH * (K + 1) + GCH * (N + 1)) * sizeof(float);
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(grad_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) { set_error("grad_weights smem attribute", e); return GPNERF_E_CUDA; }
    set = true;
  }
  grad_weights<<<grid_cap((P + GCH - 1) / GCH, 2), 256, smem, (cudaStream_t)stream>>>(a);
  return check_launch("k6_grad_weights");
}

int gpnerf_k6_assemble_raw(const float* rgb, const float* s_relu, const float* mask, int n_views, long long P,
                           float* raw, void* stream) {
  GPNERF_REQUIRE(rgb && s_relu && mask && raw && P > 0 && n_views >= 1 && n_views <= GPNERF_MAX_VIEWS);
  assemble_raw<<<grid_cap((P + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(rgb, s_relu, mask, n_views, P,
                                                                               reinterpret_cast<float4*>(raw));
  return check_launch("k6_assemble_raw");
}

int gpnerf_k6_raw_grad_split(const float* d_raw, const float* rgb, const float* s_relu, const float* mask,
                             int n_views, long long P, float* d_rgb_pre, float* d_s_pre, void* stream) {
  GPNERF_REQUIRE(d_raw && rgb && s_relu && mask && d_rgb_pre && d_s_pre && P > 0 && n_views >= 1 &&
                 n_views <= GPNERF_MAX_VIEWS);
  raw_grad_split<<<grid_cap((P + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(d_raw), rgb, s_relu, mask, n_views, P, d_rgb_pre, d_s_pre);
  return check_launch("k6_raw_grad_split");
}

int gpnerf_k6_meanvar_bwd(const float* rgb_feat, const float* meanvar, const float* d_meanvar, int n_views,
                          long long P, float* d_rgb_feat, void* stream) {
  GPNERF_REQUIRE(rgb_feat && meanvar && d_meanvar && d_rgb_feat && P > 0 && n_views >= 1 && n_views <= GPNERF_MAX_VIEWS);
  meanvar_bwd<<<grid_cap((P * 35 + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(rgb_feat, meanvar, d_meanvar,
                                                                                   n_views, P, d_rgb_feat);
  return check_launch("k6_meanvar_bwd");
}

int gpnerf_k6_from_channels_last(const float* nxc, int batch, long long n, float* cxn, void* stream) {
  GPNERF_REQUIRE(nxc && cxn && batch > 0 && n > 0);
  dim3 grid(grid_cap((n + 31) / 32, 16), batch);
  from_channels_last_32<<<grid, 256, 0, (cudaStream_t)stream>>>(nxc, n, cxn);
  return check_launch("k6_from_channels_last");
}

}  // extern "C"
