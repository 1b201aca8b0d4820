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
// branch-free forms for the epilogues (a per-element switch compiles to an indirect branch plus out-of-line
// expm1f / expf calls: ≈400 cycles per element, the whole epilogue of a tile 12 k cycles)
template <int EPI>
__device__ __forceinline__ float tce_apply_t(float v, float aux) {
  if (EPI == TCE_ELU) return v > 0.0f ? v : ex2_ftz(v * kLog2e) - 1.0f;
  if (EPI == TCE_RELU) return fmaxf(v, 0.0f);
  if (EPI == TCE_SIGMOID) return __fdividef(1.0f, 1.0f + ex2_ftz(-v * kLog2e));
  if (EPI == TCE_MUL_DELU) return v * delu_from_out(aux);
  return v;
}
__device__ __forceinline__ void sync_round() {
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}
__host__ __device__ constexpr uint32_t tmem_cols_for(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256)); }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
  int nbuf;        // X tiles in flight (2..4)
  int vec_y;       // Y (and aux) rows are 16-byte aligned and N % 4 == 0: float4 epilogue
};

#ifdef GPNERF_K6_TRACE
__device__ long long g_k6_trace[2048];
__device__ int g_k6_trace_n;
__device__ unsigned long long g_k6_span[1024];
__device__ __forceinline__ unsigned long long k6_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define K6_STAMP() do { if (tid == 0 && blockIdx.x == 0 && g_k6_trace_n < 2040) g_k6_trace[g_k6_trace_n++] = clock64(); } while (0)
#else
#define K6_STAMP() do {} while (0)
#endif
// ---- asynchronous copies global → shared (LDGSTS): no registers held while the data is in flight, so a thread
// keeps a whole tile (or several) in flight; src-size 0 writes zeros (rows beyond P, padding columns)
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n) {      // at most n groups still pending
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}

// Y[P x N] = epi(in_scale · (X ⊙ ELU'(in_aux)) · B + bias), one 128-row tile per round, persistent CTAs.
// Streaming kernel (≈100 B read + ≈100 B written per point and layer, a few hundred FLOP): what matters is bytes in
// flight.  The X tiles go global → shared with cp.async straight into the tcgen05 operand layout, `nbuf` − 1 tiles
// ahead of the one being multiplied; the accumulator is double buffered in TMEM, so the epilogue of tile i − 1
// (TMEM → registers → global) overlaps the GEMM of tile i and the loads of tiles i + 1 …
// VEC: X rows (and in_aux rows) are 16-byte aligned and K % 4 == 0 – 16-byte copies (one core-matrix row each).
constexpr int kStgStride = 20;        // floats per patch row (16 + 4: conflict-free 16-byte row writes)
constexpr size_t kStgBytes = 8 * 32 * kStgStride * 4;
template <bool VEC, int EPI>
__global__ void __launch_bounds__(256, 2) linear_tc(LinTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Kp = a.Kp, Np = a.Np, K = a.K, N = a.N, nbuf = a.nbuf;
  const uint32_t sbo = (uint32_t)(Kp / 4) * kLBO;          // both operands are [rows x Kp]
  const uint32_t tile_bytes = 16 * sbo;
  uint8_t* Bt = smem;                                       // [Np x Kp]
  uint8_t* At = smem + (size_t)(Np / 8) * sbo;              // nbuf x [128 x Kp]
  uint8_t* Xt = At + (size_t)nbuf * tile_bytes;             // nbuf x [128 x Kp]: in_aux tiles (only with in_aux)
  float* stg = reinterpret_cast<float*>(Xt + (a.in_aux ? (size_t)nbuf * tile_bytes : 0));      // 8 warps x [32 x 16] patches
  uint64_t* bar = reinterpret_cast<uint64_t*>(stg + 8 * 32 * kStgStride);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t acc_cols = tmem_cols_for(Np);
  K6_STAMP();
#ifdef GPNERF_K6_TRACE
  if (tid == 0 && blockIdx.x < 512) g_k6_span[2 * blockIdx.x] = k6_gtime();
#endif
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 2 * acc_cols);
  const long long n_tiles = (a.P + 127) / 128;
  const uint32_t a_addr = smem_u32(At), x_addr = smem_u32(Xt), b_addr = smem_u32(Bt);
  // all copies of tile `t` into buffer `b` (rows beyond P and columns beyond K are zero-filled)
  auto issue_loads = [&](long long t, int b) {
    const long long first = t * 128;
    const uint32_t ab = a_addr + (uint32_t)b * tile_bytes, xb = x_addr + (uint32_t)b * tile_bytes;
    if (VEC) {
      const int r8 = lane & 7, kq = lane >> 3;
      for (int rg = warp; rg < 16; rg += 8) {
        const long long p = first + rg * 8 + r8;
        const bool okp = p < a.P;
        const float* xr = a.X + (okp ? p : 0) * a.ldx;
        const float* ar = a.in_aux ? a.in_aux + (okp ? p : 0) * a.ld_in_aux : nullptr;
        const uint32_t off = (uint32_t)rg * sbo + (uint32_t)r8 * 16;
        for (int kc = kq; kc < Kp / 4; kc += 4) {
          const bool ok = okp && 4 * kc < K;
          cp_async16(ab + off + (uint32_t)kc * kLBO, xr + (ok ? 4 * kc : 0), ok);
          if (ar) cp_async16(xb + off + (uint32_t)kc * kLBO, ar + (ok ? 4 * kc : 0), ok);
        }
      }
    } else {
      const int r8 = lane >> 2, e = lane & 3;
      for (int rg = warp; rg < 16; rg += 8) {
        const long long p = first + rg * 8 + r8;
        const bool okp = p < a.P;
        const float* xr = a.X + (okp ? p : 0) * a.ldx;
        const float* ar = a.in_aux ? a.in_aux + (okp ? p : 0) * a.ld_in_aux : nullptr;
        const uint32_t off = (uint32_t)rg * sbo + (uint32_t)r8 * 16 + (uint32_t)e * 4;
        for (int kc = 0; kc < Kp / 4; ++kc) {
          const int k = 4 * kc + e;
          const bool ok = okp && k < K;
          cp_async4(ab + off + (uint32_t)kc * kLBO, xr + (ok ? k : 0), ok);
          if (ar) cp_async4(xb + off + (uint32_t)kc * kLBO, ar + (ok ? k : 0), ok);
        }
      }
    }
  };
  // X ⊙ ELU'(in_aux) on the elements this thread copied itself (visible to it after its own wait_group)
  auto scale_by_aux = [&](int b) {
    uint8_t* ab = At + (size_t)b * tile_bytes;
    const uint8_t* xb = Xt + (size_t)b * tile_bytes;
    if (VEC) {
      const int r8 = lane & 7, kq = lane >> 3;
      for (int rg = warp; rg < 16; rg += 8) {
        const uint32_t off = (uint32_t)rg * sbo + (uint32_t)r8 * 16;
        for (int kc = kq; kc < Kp / 4; kc += 4) {
          float4 x = *reinterpret_cast<float4*>(ab + off + (uint32_t)kc * kLBO);
          const float4 h = *reinterpret_cast<const float4*>(xb + off + (uint32_t)kc * kLBO);
          x.x *= delu_from_out(h.x); x.y *= delu_from_out(h.y); x.z *= delu_from_out(h.z); x.w *= delu_from_out(h.w);
          *reinterpret_cast<float4*>(ab + off + (uint32_t)kc * kLBO) = x;
        }
      }
    } else {
      const int r8 = lane >> 2, e = lane & 3;
      for (int rg = warp; rg < 16; rg += 8) {
        const uint32_t off = (uint32_t)rg * sbo + (uint32_t)r8 * 16 + (uint32_t)e * 4;
        for (int kc = 0; kc < Kp / 4; ++kc) {
          float* x = reinterpret_cast<float*>(ab + off + (uint32_t)kc * kLBO);
          *x *= delu_from_out(*reinterpret_cast<const float*>(xb + off + (uint32_t)kc * kLBO));
        }
      }
    }
  };
  auto tile_of = [&](int i) { return (long long)blockIdx.x + (long long)i * gridDim.x; };
  // B[n][k] = W^T or W, zero padded: asynchronous copies as well (a loop of dependent load → store pairs cost
  // 19 k cycles per CTA here), in the same group as the first X tile
  for (int i = tid; i < Np * Kp; i += blockDim.x) {
    const int n = i / Kp, k = i - n * Kp;
    const bool ok = n < N && k < K;
    const float* src = a.W;
    if (ok) src = a.w_is_kn ? a.W + (long long)k * a.ldw + n : a.W + (long long)n * a.ldw + k;
    cp_async4(b_addr + off32(n, k, sbo), src, ok);
  }
  // the first nbuf − 1 tiles are requested before anything else
  for (int j = 0; j < nbuf - 1; ++j) {
    if (tile_of(j) < n_tiles) issue_loads(tile_of(j), j);
    cp_async_commit();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  K6_STAMP();
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int row = tid & 127, half = tid >> 7;
  uint32_t ph[2] = {0u, 0u};

  auto epilogue = [&](long long tile, int ab) {
    const uint32_t t_acc = t_row + (uint32_t)ab * acc_cols;
    for (int c0 = half * 16; c0 < Np; c0 += 32) {
      uint32_t r[16];
      K6_STAMP();
      tmem_ld16(t_acc + c0, r);
      tmem_wait_ld();
      K6_STAMP();
      const bool rd_y = a.add_pre || a.add_post;
      if (a.vec_y) {
        // Thread = row holds 16 consecutive columns; written like that, a warp store touches 32 different lines (32 L1
        // wavefronts for 512 bytes).  The warp's [32 x 16] block goes through a private shared-memory patch instead
        // and leaves – and its aux / y operands arrive – as 4 instructions of 8 rows x 64 contiguous bytes.
        float* patch = stg + warp * (32 * kStgStride);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = c0 + 4 * q;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias && col < N) bv = make_float4(__ldg(a.bias + col), __ldg(a.bias + col + 1), __ldg(a.bias + col + 2), __ldg(a.bias + col + 3));
          *reinterpret_cast<float4*>(patch + lane * kStgStride + 4 * q) =
              make_float4(fmaf(__uint_as_float(r[4 * q]), a.in_scale, bv.x), fmaf(__uint_as_float(r[4 * q + 1]), a.in_scale, bv.y),
                          fmaf(__uint_as_float(r[4 * q + 2]), a.in_scale, bv.z), fmaf(__uint_as_float(r[4 * q + 3]), a.in_scale, bv.w));
        }
        __syncwarp();
        const int col = c0 + 4 * (lane & 3);
        const long long p0 = tile * 128 + (warp & 3) * 32 + (lane >> 2);
        float4 yv[4], av[4];
#pragma unroll
        for (int sI = 0; sI < 4; ++sI) {          // all operand loads first
          const long long pp = p0 + 8 * sI;
          const bool ok = pp < a.P && col < N;
          yv[sI] = av[sI] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && rd_y) yv[sI] = *reinterpret_cast<const float4*>(a.Y + pp * a.ldy + col);
          if (ok && a.aux) av[sI] = __ldg(reinterpret_cast<const float4*>(a.aux + pp * a.ld_aux + col));
        }
#pragma unroll
        for (int sI = 0; sI < 4; ++sI) {
          const long long pp = p0 + 8 * sI;
          if (!(pp < a.P && col < N)) continue;
          const float4 t = *reinterpret_cast<const float4*>(patch + (8 * sI + (lane >> 2)) * kStgStride + 4 * (lane & 3));
          float v[4] = {t.x, t.y, t.z, t.w};
          const float yy[4] = {yv[sI].x, yv[sI].y, yv[sI].z, yv[sI].w}, aa[4] = {av[sI].x, av[sI].y, av[sI].z, av[sI].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (a.add_pre) v[j] += yy[j];
            v[j] = tce_apply_t<EPI>(v[j], aa[j]);
            if (a.add_post) v[j] += yy[j];
          }
          *reinterpret_cast<float4*>(a.Y + pp * a.ldy + col) = make_float4(v[0], v[1], v[2], v[3]);
        }
      } else {
        // rows of Y that are not 16-byte aligned (the 70- / 35-float gather gradients): same patch, read back as
        // 16 instructions of 2 rows x 16 consecutive floats
        float* patch = stg + warp * (32 * kStgStride);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float b4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) b4[j] = (a.bias && c0 + 4 * q + j < N) ? __ldg(a.bias + c0 + 4 * q + j) : 0.0f;
          *reinterpret_cast<float4*>(patch + lane * kStgStride + 4 * q) =
              make_float4(fmaf(__uint_as_float(r[4 * q]), a.in_scale, b4[0]), fmaf(__uint_as_float(r[4 * q + 1]), a.in_scale, b4[1]),
                          fmaf(__uint_as_float(r[4 * q + 2]), a.in_scale, b4[2]), fmaf(__uint_as_float(r[4 * q + 3]), a.in_scale, b4[3]));
        }
        __syncwarp();
        const int col = c0 + (lane & 15);
        const long long p0 = tile * 128 + (warp & 3) * 32 + (lane >> 4);
#pragma unroll 4
        for (int sI = 0; sI < 16; ++sI) {
          const long long pp = p0 + 2 * sI;
          if (!(pp < a.P && col < N)) continue;
          float v = patch[(2 * sI + (lane >> 4)) * kStgStride + (lane & 15)];
          const float yy = rd_y ? a.Y[pp * a.ldy + col] : 0.0f;
          const float aa = a.aux ? __ldg(a.aux + pp * a.ld_aux + col) : 0.0f;
          if (a.add_pre) v += yy;
          v = tce_apply_t<EPI>(v, aa);
          if (a.add_post) v += yy;
          a.Y[pp * a.ldy + col] = v;
        }
      }
    }
  };

  int i = 0;
  for (; tile_of(i) < n_tiles; ++i) {
    const int b = i % nbuf, ab = i & 1;
    K6_STAMP();
    if (i >= 1) {                      // GEMM of tile i − 1 done: its X buffer is free, its accumulator is complete
      mbar_wait(bar + (ab ^ 1), ph[ab ^ 1]);
      ph[ab ^ 1] ^= 1u;
      tc_fence_after();
    }
    K6_STAMP();
    {
      const int j = i + nbuf - 1;      // (its buffer is the one tile i − 1 has just released)
      if (tile_of(j) < n_tiles) issue_loads(tile_of(j), j % nbuf);
      cp_async_commit();
    }
    K6_STAMP();
    cp_async_wait_pending(nbuf - 1);   // this thread's copies of tile i have landed
    K6_STAMP();
    if (a.in_aux) scale_by_aux(b);
    sync_round();
    K6_STAMP();
    if (tid == 0)
      issue_gemm_tf32(a_addr + (uint32_t)b * tile_bytes, sbo, b_addr, sbo, Kp, Np, tmem + (uint32_t)ab * acc_cols, false, bar + ab);
    if (i >= 1) epilogue(tile_of(i - 1), ab ^ 1);
    K6_STAMP();
  }
  if (i >= 1) {
    const int ab = (i - 1) & 1;
    mbar_wait(bar + ab, ph[ab]);
    tc_fence_after();
    epilogue(tile_of(i - 1), ab);
  }
  cp_async_wait_pending(0);
  K6_STAMP();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 2 * acc_cols);
  K6_STAMP();
#ifdef GPNERF_K6_TRACE
  if (tid == 0 && blockIdx.x < 512) g_k6_span[2 * blockIdx.x + 1] = k6_gtime();
#endif
}

struct GwTcArgs {
  const float* X; int ldx; int K; float in_scale;
  const float* dY; int ldy; int N;
  const float* dy_aux; int ld_dy_aux;
  float* dW; int ldw; float* db; long long P;
  int Kp;          // (K + 1 "ones" row for the bias gradient) rounded up to 16
  int nbuf;        // operand chunks in flight (2..3)
};
constexpr int GPT = 64;      // points per chunk = contraction length per round

// dW[N x K] += in_scale · (dY ⊙ ELU'(dy_aux))ᵀ · X, db[N] += Σ_p dY ⊙ ELU'(dy_aux): contraction over the points,
// 64 per round, accumulated in ONE TMEM accumulator over all rounds of the CTA and added to dW / db with atomics
// at the end.  Both operands are staged transposed (feature-major) with 4-byte cp.async copies straight into the
// tcgen05 operand layout – any row stride works – `nbuf` − 1 chunks ahead of the one being multiplied.
__global__ void __launch_bounds__(256, 2) grad_weights_tc(GwTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Kp = a.Kp, K = a.K, N = a.N, nbuf = a.nbuf;
  constexpr uint32_t sbo = (uint32_t)(GPT / 4) * kLBO;      // operands are [rows x GPT points]
  const uint32_t a_bytes = 16 * sbo, b_bytes = (uint32_t)(Kp / 8) * sbo, h_bytes = a.dy_aux ? a_bytes : 0u;
  const uint32_t buf_bytes = a_bytes + b_bytes + h_bytes;   // dY^T [128 x GPT] | X^T [Kp x GPT] (row K = ones) | aux^T
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)nbuf * buf_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tmem_cols_for(Kp);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  // rows >= N of every dY^T tile stay zero
  for (int b = 0; b < nbuf; ++b)
    for (int i = tid; i < (int)(a_bytes / 16); i += blockDim.x)
      reinterpret_cast<uint4*>(smem + (size_t)b * buf_bytes)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const int f8 = lane >> 2, e = lane & 3;                   // feature within its group of 8, point within its group of 4
  const int nA = (N + 7) / 8, nB = Kp / 8;                  // feature groups of the two operands, dealt to the warps together
  const long long n_chunks = (a.P + GPT - 1) / GPT;
  auto chunk_of = [&](int i) { return (long long)blockIdx.x + (long long)i * gridDim.x; };
  const uint32_t s_addr = smem_u32(smem);
  auto issue_loads = [&](long long ch, int b) {
    const long long first = ch * GPT;
    const uint32_t base = s_addr + (uint32_t)b * buf_bytes;
    for (int t = warp; t < nA + nB; t += 8) {
      if (t < nA) {
        const int n = 8 * t + f8;
        const uint32_t dst = base + (uint32_t)t * sbo + (uint32_t)f8 * 16 + (uint32_t)e * 4;
#pragma unroll 4
        for (int pc = 0; pc < GPT / 4; ++pc) {
          const long long p = first + 4 * pc + e;
          const bool ok = p < a.P && n < N;
          cp_async4(dst + (uint32_t)pc * kLBO, a.dY + (ok ? p * a.ldy + n : 0), ok);
          if (a.dy_aux) cp_async4(dst + a_bytes + b_bytes + (uint32_t)pc * kLBO, a.dy_aux + (ok ? p * a.ld_dy_aux + n : 0), ok);
        }
      } else {
        const int g = t - nA, k = 8 * g + f8;
        const uint32_t dst = base + a_bytes + (uint32_t)g * sbo + (uint32_t)f8 * 16 + (uint32_t)e * 4;
        if (k == K) continue;                                // the ones row is written after the wait (plain stores)
#pragma unroll 4
        for (int pc = 0; pc < GPT / 4; ++pc) {
          const long long p = first + 4 * pc + e;
          const bool ok = p < a.P && k < K;
          cp_async4(dst + (uint32_t)pc * kLBO, a.X + (ok ? p * a.ldx + k : 0), ok);
        }
      }
    }
  };
  // what cp.async cannot do, on the elements this thread owns: the ones row of X^T and dY ⊙ ELU'(dy_aux)
  auto fix_up = [&](long long ch, int b) {
    const long long first = ch * GPT;
    uint8_t* base = smem + (size_t)b * buf_bytes;
    for (int t = warp; t < nA + nB; t += 8) {
      if (t < nA) {
        if (!a.dy_aux) continue;
        uint8_t* dst = base + (uint32_t)t * sbo + (uint32_t)f8 * 16 + (uint32_t)e * 4;
        for (int pc = 0; pc < GPT / 4; ++pc) {
          float* d = reinterpret_cast<float*>(dst + (uint32_t)pc * kLBO);
          *d *= delu_from_out(*reinterpret_cast<const float*>(dst + a_bytes + b_bytes + (uint32_t)pc * kLBO));
        }
      } else {
        const int g = t - nA, k = 8 * g + f8;
        if (k != K) continue;
        uint8_t* dst = base + a_bytes + (uint32_t)g * sbo + (uint32_t)f8 * 16 + (uint32_t)e * 4;
        for (int pc = 0; pc < GPT / 4; ++pc)
          *reinterpret_cast<float*>(dst + (uint32_t)pc * kLBO) = (first + 4 * pc + e < a.P) ? 1.0f : 0.0f;
      }
    }
  };
  for (int j = 0; j < nbuf - 1; ++j) {
    if (chunk_of(j) < n_chunks) issue_loads(chunk_of(j), j);
    cp_async_commit();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;
  int i = 0;
  for (; chunk_of(i) < n_chunks; ++i) {
    const int b = i % nbuf;
    if (i >= 1) {                      // GEMM of chunk i − 1 done: its buffer is free
      mbar_wait(bar, phase);
      phase ^= 1u;
      tc_fence_after();
    }
    {
      const int j = i + nbuf - 1;
      if (chunk_of(j) < n_chunks) issue_loads(chunk_of(j), j % nbuf);
      cp_async_commit();
    }
    cp_async_wait_pending(nbuf - 1);
    fix_up(chunk_of(i), b);
    sync_round();
    if (tid == 0) {
      const uint32_t base = s_addr + (uint32_t)b * buf_bytes;
      issue_gemm_tf32(base, sbo, base + a_bytes, sbo, GPT, Kp, tmem, i > 0, bar);
    }
  }
  cp_async_wait_pending(0);
  if (i >= 1) {
    mbar_wait(bar, phase);
    tc_fence_after();
    const int n = tid & 127, half = tid >> 7;
    for (int c0 = half * 16; c0 < Kp; c0 += 32) {
      uint32_t r[16];
      tmem_ld16(t_row + c0, r);
      tmem_wait_ld();
      if (n < N) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = c0 + j;
          if (k < K) atomicAdd(a.dW + (long long)n * a.ldw + k, __uint_as_float(r[j]) * a.in_scale);
          else if (k == K && a.db != nullptr) atomicAdd(a.db + n, __uint_as_float(r[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

int linear_tc_launch(const float* X, int ldx, int K, float in_scale, const float* in_aux, int ld_in_aux,
                     const float* W, int ldw, int w_is_kn, int N, const float* bias, int epilogue, const float* aux,
                     int ld_aux, float* Y, int ldy, int add_pre, int add_post, long long P, cudaStream_t st) {
  LinTcArgs a{X, ldx, K, in_scale, in_aux, ld_in_aux, W, ldw, w_is_kn, N, bias, epilogue, aux, ld_aux, Y, ldy,
              add_pre, add_post, P, (K + 7) & ~7, (N + 15) & ~15, 2, 0};
  const bool vec_x = aligned16(X) && ldx % 4 == 0 && K % 4 == 0 &&
                     (in_aux == nullptr || (aligned16(in_aux) && ld_in_aux % 4 == 0));
  a.vec_y = aligned16(Y) && ldy % 4 == 0 && N % 4 == 0 && (aux == nullptr || (aligned16(aux) && ld_aux % 4 == 0)) &&
            (bias == nullptr || true);
  // shared memory: weights + nbuf X tiles (+ nbuf in_aux tiles); two CTAs per SM when at least two tiles each fit
  const size_t tile = (size_t)128 * a.Kp * 4 * (in_aux ? 2 : 1), wb = (size_t)a.Np * a.Kp * 4;
  const size_t budget = 220 * 1024;
  int ctas = 2;
  const long long fixed = (long long)wb + (long long)kStgBytes + 256;
  long long nbuf = ((long long)(budget / 2) - fixed) / (long long)tile;
  if (nbuf < 2) {
    ctas = 1;
    nbuf = ((long long)budget - fixed) / (long long)tile;
  }
  if (nbuf < 2) {
    set_error("k6_linear (tcgen05 tf32): K too large for two X tiles in shared memory", cudaSuccess);
    return GPNERF_E_UNSUPPORTED;
  }
  a.nbuf = (int)(nbuf > 4 ? 4 : nbuf);
  const size_t smem = wb + (size_t)a.nbuf * tile + kStgBytes + 64;
  const long long tiles = (P + 127) / 128, cap = (long long)ctas * sm_count();
  const int grid = (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
  using Kern = void (*)(LinTcArgs);
  static const Kern kerns[2][5] = {
      {linear_tc<false, 0>, linear_tc<false, 1>, linear_tc<false, 2>, linear_tc<false, 3>, linear_tc<false, 4>},
      {linear_tc<true, 0>, linear_tc<true, 1>, linear_tc<true, 2>, linear_tc<true, 3>, linear_tc<true, 4>}};
  static bool set = false;
  if (!set) {
    for (int v = 0; v < 2; ++v)
      for (int e = 0; e < 5; ++e) {
        cudaError_t err = cudaFuncSetAttribute(kerns[v][e], cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (err != cudaSuccess) {
          set_error("linear_tc smem attribute", err);
          return GPNERF_E_CUDA;
        }
      }
    set = true;
  }
  if (epilogue < 0 || epilogue > 4) {
    set_error("k6_linear: unknown epilogue", cudaSuccess);
    return GPNERF_E_ARG;
  }
  kerns[vec_x ? 1 : 0][epilogue]<<<grid, 256, smem, st>>>(a);
  return check_launch("k6_linear (tcgen05 tf32)");
}

// Same contraction for 16-byte aligned operands (the activations and their gradients: every layer but the ones that
// read the 70- / 35-float gather outputs): 4-byte cp.async copies cost the LSU ≈30 cycles per 128 bytes, so the
// operands are fetched with 16-byte loads into registers – one chunk ahead of the chunk being multiplied – and
// scattered into the transposed operand layout with 4-byte shared-memory stores; dY ⊙ ELU'(dy_aux) is applied in
// registers.  Two operand buffers, two barriers: the GEMM of chunk i − 1 overlaps the staging of chunk i.
__global__ void __launch_bounds__(256, 2) grad_weights_vec(GwTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Kp = a.Kp, K = a.K, N = a.N;
  constexpr uint32_t sbo = (uint32_t)(GPT / 4) * kLBO;
  const uint32_t a_bytes = 16 * sbo, b_bytes = (uint32_t)(Kp / 8) * sbo, buf_bytes = a_bytes + b_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * (size_t)buf_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t ncols = tmem_cols_for(Kp);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  // rows >= N of dY^T and rows >= K of X^T stay zero (row K, the ones row, is rewritten per chunk)
  for (int i = tid; i < (int)(2 * buf_bytes / 16); i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  const long long n_chunks = (a.P + GPT - 1) / GPT;
  auto chunk_of = [&](int i) { return (long long)blockIdx.x + (long long)i * gridDim.x; };
  const int qa = N / 4, qb = K / 4;                         // 16-byte pieces per point of dY and of X
  const int na = (GPT * qa + 255) / 256, nb = (GPT * qb + 255) / 256;        // pieces per thread: <= 4, <= 8
  float4 dyv[4], hv[4], xv[8];
  auto load_regs = [&](long long ch) {
    const long long first = ch * GPT;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + 256 * it, pl = idx / qa, q = idx - pl * qa;
      dyv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      hv[it] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (it < na && pl < GPT && first + pl < a.P) {
        dyv[it] = __ldg(reinterpret_cast<const float4*>(a.dY + (first + pl) * a.ldy) + q);
        if (a.dy_aux) hv[it] = __ldg(reinterpret_cast<const float4*>(a.dy_aux + (first + pl) * a.ld_dy_aux) + q);
      }
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = tid + 256 * it, pl = idx / qb, q = idx - pl * qb;
      xv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (it < nb && pl < GPT && first + pl < a.P) xv[it] = __ldg(reinterpret_cast<const float4*>(a.X + (first + pl) * a.ldx) + q);
    }
  };
  auto store_regs = [&](long long ch, int b) {
    uint8_t* At = smem + (size_t)b * buf_bytes;
    uint8_t* Bt = At + a_bytes;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + 256 * it, pl = idx / qa, q = idx - pl * qa;
      if (it < na && pl < GPT) {
        const float v[4] = {dyv[it].x * delu_from_out(hv[it].x), dyv[it].y * delu_from_out(hv[it].y),
                            dyv[it].z * delu_from_out(hv[it].z), dyv[it].w * delu_from_out(hv[it].w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float*>(At + off32(4 * q + j, pl, sbo)) = v[j];
      }
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = tid + 256 * it, pl = idx / qb, q = idx - pl * qb;
      if (it < nb && pl < GPT) {
        const float v[4] = {xv[it].x, xv[it].y, xv[it].z, xv[it].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float*>(Bt + off32(4 * q + j, pl, sbo)) = v[j];
      }
    }
    if (tid < GPT) *reinterpret_cast<float*>(Bt + off32(K, tid, sbo)) = (ch * GPT + tid < a.P) ? 1.0f : 0.0f;   // ones row → db
  };
  if (chunk_of(0) < n_chunks) load_regs(chunk_of(0));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t s_addr = smem_u32(smem);
  uint32_t ph[2] = {0u, 0u};
  int i = 0;
  for (; chunk_of(i) < n_chunks; ++i) {
    const int b = i & 1;
    if (i >= 2) {                      // GEMM of chunk i − 2 done: its buffer is free
      mbar_wait(bar + b, ph[b]);
      ph[b] ^= 1u;
    }
    store_regs(chunk_of(i), b);
    if (chunk_of(i + 1) < n_chunks) load_regs(chunk_of(i + 1));      // in flight while this chunk is multiplied
    sync_round();
    if (tid == 0) {
      const uint32_t base = s_addr + (uint32_t)b * buf_bytes;
      issue_gemm_tf32(base, sbo, base + a_bytes, sbo, GPT, Kp, tmem, i > 0, bar + b);
    }
  }
  if (i >= 1) {
    // everything issued has completed when the last commit has arrived (tcgen05.commit covers all earlier MMAs)
    const int b = (i - 1) & 1;
    mbar_wait(bar + b, ph[b]);
    tc_fence_after();
    const int n = tid & 127, half = tid >> 7;
    for (int c0 = half * 16; c0 < Kp; c0 += 32) {
      uint32_t r[16];
      tmem_ld16(t_row + c0, r);
      tmem_wait_ld();
      if (n < N) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = c0 + j;
          if (k < K) atomicAdd(a.dW + (long long)n * a.ldw + k, __uint_as_float(r[j]) * a.in_scale);
          else if (k == K && a.db != nullptr) atomicAdd(a.db + n, __uint_as_float(r[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

int grad_weights_tc_launch(const float* X, int ldx, int K, float in_scale, const float* dY, int ldy, int N,
                           const float* dy_aux, int ld_dy_aux, float* dW, int ldw, float* db, long long P,
                           cudaStream_t st) {
  GwTcArgs a{X, ldx, K, in_scale, dY, ldy, N, dy_aux, ld_dy_aux, dW, ldw, db, P, (K + 1 + 15) & ~15, 2};
  const bool vec = aligned16(X) && ldx % 4 == 0 && K % 4 == 0 && K <= 128 && aligned16(dY) && ldy % 4 == 0 && N % 4 == 0 &&
                   N <= 64 && (dy_aux == nullptr || (aligned16(dy_aux) && ld_dy_aux % 4 == 0));
  if (vec) {
    const size_t smem_v = 2 * (size_t)(128 + a.Kp) * GPT * 4 + 64;
    static bool set_v = false;
    if (!set_v) {
      cudaError_t e = cudaFuncSetAttribute(grad_weights_vec, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
      if (e != cudaSuccess) {
        set_error("grad_weights_vec smem attribute", e);
        return GPNERF_E_CUDA;
      }
      set_v = true;
    }
    const long long chunks_v = (P + GPT - 1) / GPT, cap_v = (smem_v <= 110 * 1024 ? 2ll : 1ll) * sm_count();
    const int grid_v = (int)(chunks_v < cap_v ? (chunks_v > 0 ? chunks_v : 1) : cap_v);
    grad_weights_vec<<<grid_v, 256, smem_v, st>>>(a);
    return check_launch("k6_grad_weights (tcgen05 tf32, vector staging)");
  }
  const size_t buf = (size_t)(128 + a.Kp + (dy_aux ? 128 : 0)) * GPT * 4;
  const size_t budget = 220 * 1024;
  int ctas = 2;
  long long nbuf = ((long long)(budget / 2) - 256) / (long long)buf;
  if (nbuf < 2) {
    ctas = 1;
    nbuf = ((long long)budget - 256) / (long long)buf;
  }
  if (nbuf < 2) {
    set_error("k6_grad_weights (tcgen05 tf32): K too large for two operand chunks in shared memory", cudaSuccess);
    return GPNERF_E_UNSUPPORTED;
  }
  a.nbuf = (int)(nbuf > 3 ? 3 : nbuf);
  const size_t smem = (size_t)a.nbuf * buf + 64;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(grad_weights_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) {
      set_error("grad_weights_tc smem attribute", e);
      return GPNERF_E_CUDA;
    }
    set = true;
  }
  const long long chunks = (P + GPT - 1) / GPT, cap = (long long)ctas * sm_count();
  const int grid = (int)(chunks < cap ? (chunks > 0 ? chunks : 1) : cap);
  grad_weights_tc<<<grid, 256, smem, st>>>(a);
  return check_launch("k6_grad_weights (tcgen05 tf32)");
}

#ifdef GPNERF_K6_TRACE
extern "C" int gpnerf_debug_k6_trace(long long* out, int reset) {
  int n = 0;
  cudaMemcpyFromSymbol(&n, g_k6_trace_n, sizeof(int));
  if (out) cudaMemcpyFromSymbol(out, g_k6_trace, sizeof(long long) * 2048);
  if (reset) { int z = 0; cudaMemcpyToSymbol(g_k6_trace_n, &z, sizeof(int)); }
  return n;
}
extern "C" int gpnerf_debug_k6_span(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_k6_span, sizeof(unsigned long long) * 1024);
}
#endif

}  // namespace gpnerf
