// K6 on tensor cores – the training path's two GEMM shapes as tcgen05.mma.kind::tf32
// kernels (fp32 data in HBM, TF32 operands, fp32 accumulators in TMEM):
//
//   linear_tc        Y[P x N] = epi(X[P x K] · B[K x N] + bias)      forward layers and ∂L/∂X
//                    (trainhead.py:39-41, 85-110, 118-145 and their autograd transposes)
//   grad_weights_tc  dW[N x K] += dYᵀ · X,  db[N] += Σ_p dY          ∂L/∂W, ∂L/∂b
//                    (contraction over the POINTS: both operands are staged transposed)
//
// Same semantics, argument for argument, as linear_rows / grad_weights in k6_train.cu (the fp32 CUDA-core
// parity kernels); selected by `precision = 1` in gpnerf_k6_linear / gpnerf_k6_grad_weights.
//
// Operand layout: K-major, no swizzle, 32-bit elements – core matrix = 8 rows x 16 bytes (4 values):
//     byte(r, k) = (r/8)·SBO + (k/4)·128 + (r%8)·16 + (k%4)·4 ;   one MMA consumes K = 8 (two core matrices).
// Tiles are staged by the CTA's own threads straight from the row-major fp32 arrays: lane = (row%8, k%4),
// so a warp instruction reads 8 rows x 16 contiguous bytes from global memory and writes 32 distinct banks
// (no alignment requirement on the rows: ldx = 134, column offsets … all occur).  For grad_weights the
// same mapping with the roles of "row" and "k" swapped transposes the tile on the fly.
#include "tc_common.cuh"
#include "common.cuh"

namespace gpnerf {
using namespace tc;

enum { TCE_NONE = 0, TCE_ELU = 1, TCE_RELU = 2, TCE_SIGMOID = 3, TCE_MUL_DELU = 4 };
constexpr uint32_t kFmtTF32 = 2;

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[128 x N] (+)= A[128 x Kp] · B[N x Kp]^T, Kp a multiple of 8
__device__ __forceinline__ void issue_gemm_tf32(uint32_t a_addr, uint32_t a_sbo, uint32_t b_addr, uint32_t b_sbo,
                                                int Kp, int N, uint32_t tmem_d, bool accumulate, uint64_t* bar) {
  const uint32_t idesc = make_idesc(128, N, kFmtTF32);
  for (int k8 = 0; k8 < Kp / 8; ++k8) {
    const uint64_t ad = make_smem_desc(a_addr + k8 * 2 * kLBO, kLBO, a_sbo);
    const uint64_t bd = make_smem_desc(b_addr + k8 * 2 * kLBO, kLBO, b_sbo);
    umma_tf32(tmem_d, ad, bd, idesc, (accumulate || k8 > 0) ? 1u : 0u);
  }
  umma_commit(bar);
}
__device__ __forceinline__ uint32_t off32(int r, int k, uint32_t sbo) {
  return (uint32_t)(r >> 3) * sbo + (uint32_t)(k >> 2) * kLBO + (uint32_t)(r & 7) * 16 + (uint32_t)(k & 3) * 4;
}
__device__ __forceinline__ float delu_from_out(float h) { return h > 0.0f ? 1.0f : h + 1.0f; }   // ELU'(pre) from ELU(pre)
__device__ __forceinline__ float tce_apply(float v, int epi, float aux) {
  switch (epi) {
    case TCE_ELU: return v > 0.0f ? v : expm1f(v);
    case TCE_RELU: return fmaxf(v, 0.0f);
    case TCE_SIGMOID: return 1.0f / (1.0f + expf(-v));
    case TCE_MUL_DELU: return v * delu_from_out(aux);
    default: return v;
  }
}
__device__ __forceinline__ void sync_round() {
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}
__host__ __device__ constexpr uint32_t tmem_cols_for(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256)); }

struct LinTcArgs {
  const float* X; int ldx; int K; float in_scale;
  const float* in_aux; int ld_in_aux;
  const float* W; int ldw; int w_is_kn;
  int N;
  const float* bias;
  int epi;
  const float* aux; int ld_aux;
  float* Y; int ldy;
  int add_pre, add_post;
  long long P;
  int Kp, Np;      // K rounded up to 8, N rounded up to 16
};

__global__ void __launch_bounds__(256, 2) linear_tc(LinTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Kp = a.Kp, Np = a.Np, K = a.K, N = a.N;
  const uint32_t sbo = (uint32_t)(Kp / 4) * kLBO;          // both operands are [rows x Kp]
  uint8_t* Bt = smem;                                       // [Np x Kp]
  uint8_t* At = smem + (size_t)(Np / 8) * sbo;              // [128 x Kp]
  uint64_t* bar = reinterpret_cast<uint64_t*>(At + 16 * sbo);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tmem_cols_for(Np);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  // B[n][k] = W^T or W, zero padded
  for (int i = tid; i < Np * Kp; i += blockDim.x) {
    const int n = i / Kp, k = i - n * Kp;
    float v = 0.0f;
    if (n < N && k < K) v = __ldg(a.w_is_kn ? a.W + (long long)k * a.ldw + n : a.W + (long long)n * a.ldw + k);
    *reinterpret_cast<float*>(Bt + off32(n, k, sbo)) = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t a_addr = smem_u32(At), b_addr = smem_u32(Bt);
  const int row = tid & 127, half = tid >> 7;
  const int r8 = lane >> 2, e = lane & 3;
  uint32_t phase = 0;
  const long long n_tiles = (a.P + 127) / 128;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long first = tile * 128;
    // ---- stage the X tile (input scaling / ELU' of the saved activation applied on the way)
    for (int rg = warp; rg < 16; rg += 8) {
      const long long p = first + rg * 8 + r8;
      const bool okp = p < a.P;
      const float* xr = a.X + p * a.ldx;
      const float* ar = a.in_aux ? a.in_aux + p * a.ld_in_aux : nullptr;
      uint8_t* dst = At + (uint32_t)rg * sbo + (uint32_t)r8 * 16 + (uint32_t)e * 4;
      // batches of 8 independent loads before their stores (the loop is latency bound otherwise)
      for (int kc0 = 0; kc0 < Kp / 4; kc0 += 8) {
        float x[8], h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = 4 * (kc0 + j) + e;
          const bool ok = okp && k < K;
          x[j] = ok ? __ldg(xr + k) : 0.0f;
          h[j] = (ok && ar) ? __ldg(ar + k) : 1.0f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (kc0 + j < Kp / 4)
            *reinterpret_cast<float*>(dst + (uint32_t)(kc0 + j) * kLBO) = x[j] * a.in_scale * delu_from_out(h[j]);
      }
    }
    sync_round();
    if (tid == 0) issue_gemm_tf32(a_addr, sbo, b_addr, sbo, Kp, Np, tmem, false, bar);
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    // ---- epilogue: thread (row, half) takes the 16-column blocks half, half+2, …
    const long long p = first + row;
    for (int c0 = half * 16; c0 < Np; c0 += 32) {
      uint32_t r[16];
      tmem_ld16(t_row + c0, r);
      tmem_wait_ld();
      if (p < a.P) {
        float* y = a.Y + p * a.ldy;
        const float* ax = a.aux ? a.aux + p * a.ld_aux : nullptr;
        const bool rd_y = a.add_pre || a.add_post;
        float yv[16], av[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {           // all loads before the first store to y
          const int col = c0 + j;
          yv[j] = (rd_y && col < N) ? y[col] : 0.0f;
          av[j] = (ax && col < N) ? __ldg(ax + col) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = c0 + j;
          if (col < N) {
            float v = __uint_as_float(r[j]) + (a.bias ? __ldg(a.bias + col) : 0.0f);
            if (a.add_pre) v += yv[j];
            v = tce_apply(v, a.epi, av[j]);
            if (a.add_post) v += yv[j];
            y[col] = v;
          }
        }
      }
    }
    // TMEM reads of this tile are ordered before the next tile's MMAs by the next sync_round
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

struct GwTcArgs {
  const float* X; int ldx; int K; float in_scale;
  const float* dY; int ldy; int N;
  const float* dy_aux; int ld_dy_aux;
  float* dW; int ldw; float* db; long long P;
  int Kp;          // (K + 1 "ones" row for the bias gradient) rounded up to 16
};
constexpr int GPT = 64;      // points per chunk = contraction length per round

__global__ void __launch_bounds__(256, 2) grad_weights_tc(GwTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Kp = a.Kp, K = a.K, N = a.N;
  constexpr uint32_t sbo = (uint32_t)(GPT / 4) * kLBO;      // operands are [rows x GPT points]
  uint8_t* At = smem;                                       // dY^T : [128 x GPT] (rows >= N stay zero)
  uint8_t* Bt = smem + 16 * sbo;                            // X^T  : [Kp x GPT] (row K = ones)
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bt + (size_t)(Kp / 8) * sbo);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tmem_cols_for(Kp);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  for (int i = tid; i < (int)(16 * sbo / 16); i += blockDim.x) reinterpret_cast<uint4*>(At)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t a_addr = smem_u32(At), b_addr = smem_u32(Bt);
  const int f8 = lane >> 2, e = lane & 3;                   // feature within its group of 8, point within its group of 4
  uint32_t phase = 0;
  bool any = false;
  const long long n_chunks = (a.P + GPT - 1) / GPT;
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const long long first = ch * GPT;
    // ---- dY^T: feature groups over the warps, 16 point-quads each
    for (int g = warp; g < (N + 7) / 8; g += 8) {
      const int n = 8 * g + f8;
      uint8_t* dst = At + (uint32_t)g * sbo + (uint32_t)f8 * 16 + (uint32_t)e * 4;
      float d[GPT / 4], h[GPT / 4];
#pragma unroll
      for (int pc = 0; pc < GPT / 4; ++pc) {       // all loads of the group first
        const long long p = first + 4 * pc + e;
        const bool ok = p < a.P && n < N;
        d[pc] = ok ? __ldg(a.dY + p * a.ldy + n) : 0.0f;
        h[pc] = (ok && a.dy_aux) ? __ldg(a.dy_aux + p * a.ld_dy_aux + n) : 1.0f;
      }
#pragma unroll
      for (int pc = 0; pc < GPT / 4; ++pc)
        *reinterpret_cast<float*>(dst + (uint32_t)pc * kLBO) = d[pc] * delu_from_out(h[pc]);
    }
    // ---- X^T (+ the ones row that yields db)
    for (int g = warp; g < Kp / 8; g += 8) {
      const int k = 8 * g + f8;
      uint8_t* dst = Bt + (uint32_t)g * sbo + (uint32_t)f8 * 16 + (uint32_t)e * 4;
      float x[GPT / 4];
#pragma unroll
      for (int pc = 0; pc < GPT / 4; ++pc) {
        const long long p = first + 4 * pc + e;
        x[pc] = 0.0f;
        if (p < a.P) {
          if (k < K) x[pc] = __ldg(a.X + p * a.ldx + k) * a.in_scale;
          else if (k == K) x[pc] = 1.0f;
        }
      }
#pragma unroll
      for (int pc = 0; pc < GPT / 4; ++pc) *reinterpret_cast<float*>(dst + (uint32_t)pc * kLBO) = x[pc];
    }
    sync_round();
    if (tid == 0) issue_gemm_tf32(a_addr, sbo, b_addr, sbo, GPT, Kp, tmem, any, bar);
    any = true;
    mbar_wait(bar, phase);      // the operand tiles are rewritten next round
    phase ^= 1u;
    tc_fence_after();
  }
  if (any) {
    const int n = tid & 127, half = tid >> 7;
    for (int c0 = half * 16; c0 < Kp; c0 += 32) {
      uint32_t r[16];
      tmem_ld16(t_row + c0, r);
      tmem_wait_ld();
      if (n < N) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = c0 + j;
          if (k < K) atomicAdd(a.dW + (long long)n * a.ldw + k, __uint_as_float(r[j]));
          else if (k == K && a.db != nullptr) atomicAdd(a.db + n, __uint_as_float(r[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

static int grid_for_tiles(long long tiles) {
  const long long cap = 2ll * sm_count();
  return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

int linear_tc_launch(const float* X, int ldx, int K, float in_scale, const float* in_aux, int ld_in_aux,
                     const float* W, int ldw, int w_is_kn, int N, const float* bias, int epilogue, const float* aux,
                     int ld_aux, float* Y, int ldy, int add_pre, int add_post, long long P, cudaStream_t st) {
  LinTcArgs a{X, ldx, K, in_scale, in_aux, ld_in_aux, W, ldw, w_is_kn, N, bias, epilogue, aux, ld_aux, Y, ldy,
              add_pre, add_post, P, (K + 7) & ~7, (N + 15) & ~15};
  const size_t smem = (size_t)(a.Np + 128) * a.Kp * 4 + 64;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("linear_tc smem attribute", e);
      return GPNERF_E_CUDA;
    }
    set = true;
  }
  linear_tc<<<grid_for_tiles((P + 127) / 128), 256, smem, st>>>(a);
  return check_launch("k6_linear (tcgen05 tf32)");
}

int grad_weights_tc_launch(const float* X, int ldx, int K, float in_scale, const float* dY, int ldy, int N,
                           const float* dy_aux, int ld_dy_aux, float* dW, int ldw, float* db, long long P,
                           cudaStream_t st) {
  GwTcArgs a{X, ldx, K, in_scale, dY, ldy, N, dy_aux, ld_dy_aux, dW, ldw, db, P, (K + 1 + 15) & ~15};
  const size_t smem = (size_t)(128 + a.Kp) * GPT * 4 + 64;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(grad_weights_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("grad_weights_tc smem attribute", e);
      return GPNERF_E_CUDA;
    }
    set = true;
  }
  grad_weights_tc<<<grid_for_tiles((P + GPT - 1) / GPT), 256, smem, st>>>(a);
  return check_launch("k6_grad_weights (tcgen05 tf32)");
}

}  // namespace gpnerf
