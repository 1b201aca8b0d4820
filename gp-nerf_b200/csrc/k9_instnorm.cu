// K9 – SURVEY §8f row 2 (image encoder, libs/encoders/UNet.py:133-234): the normalisation layers between
// the cuDNN convolutions.  Every convolution of the ResUNet is followed by InstanceNorm2d(affine, no running
// statistics; UNet.py:32,36,120,162) and then ReLU (UNet.py:42,51), "+ identity, ReLU" (UNet.py:50-51) or ELU
// (UNet.py:123).  In torch that is 2–4 elementwise passes per layer; here two:
//
//   in_stats : Σx, Σx² per (image, channel) over H·W – fp32 partials per thread, fp64 atomics into a scratch that
//              is zero on entry and zeroed again by its reader (no memset node per norm); the last block
//              of an image folds mean, variance, γ, β into one (scale, shift) pair per channel
//   in_apply : y = act(scale·x + shift [+ residual]),  act ∈ {none, ReLU, ELU}; may run in place; can
//              write into a reflect-bordered buffer / a channel slice of a concatenation buffer (PadGeom)
//   resample_pad : skip-connection copy or bilinear ×2 upsampling into such a buffer
//
// Activations are channels-last ([N,H,W,C], what cuDNN's tensor-core convolutions want), fp16, bf16 or fp32.
// HBM-bound: in_stats reads the tensor once, in_apply reads it (+ residual) and writes it once.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace gpnerf {

template <typename T>
struct Vec;
template <>
struct Vec<float> {                       // 4 channels per 16-byte access
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  __device__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<__nv_bfloat16> {               // 8 channels per 16-byte access
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <>
struct Vec<__half> {                      // 8 channels per 16-byte access
  static constexpr int N = 8;
  __device__ static void load(const __half* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  __device__ static void store(__half* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// grid (chunks, N); a thread owns one channel group (VN channels) and strides over the pixels of its chunk.
template <typename T>
__global__ void __launch_bounds__(256) in_stats(const T* __restrict__ x, int HW, int C, int px_per_block,
                                                double* __restrict__ sums /* [N][C][2] */, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, float eps, double inv_hw,
                                                float2* __restrict__ ab, unsigned* __restrict__ tickets) {
  constexpr int VN = Vec<T>::N;
  const int groups = C / VN;                            // ≤ 256 / … checked by the caller: 256 % groups == 0
  const int g = threadIdx.x % groups, lane_px = threadIdx.x / groups, px_step = blockDim.x / groups;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * px_per_block, p1 = min(HW, p0 + px_per_block);
  float s[VN], q[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) s[i] = q[i] = 0.0f;
  const T* base = x + (size_t)n * HW * C + g * VN;
  for (int p = p0 + lane_px; p < p1; p += px_step) {
    float v[VN];
    Vec<T>::load(base + (size_t)p * C, v);
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      s[i] += v[i];
      q[i] = fmaf(v[i], v[i], q[i]);
    }
  }
  // rows of the block that share a channel group meet in shared memory: one fp64 atomic per (block, channel)
  __shared__ float red[256][2 * VN + 1];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    red[threadIdx.x][2 * i] = s[i];
    red[threadIdx.x][2 * i + 1] = q[i];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < groups * 2 * VN; t += blockDim.x) {
    const int gg = t / (2 * VN), j = t - gg * 2 * VN;
    double acc = 0.0;
    for (int r = 0; r < px_step; ++r) acc += (double)red[r * groups + gg][j];
    atomicAdd(sums + ((size_t)n * C + gg * VN) * 2 + j, acc);
  }
  // the last block of an image turns the sums into y = a·x + b per channel (fp64 once per channel, not per thread)
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(tickets + n, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double m = __ldcg(sums + ((size_t)n * C + c) * 2) * inv_hw;
    const double var = fmax(fma(__ldcg(sums + ((size_t)n * C + c) * 2 + 1), inv_hw, -m * m), 0.0);   // biased, as instance_norm
    const float a = __ldg(gamma + c) * rsqrtf((float)var + eps);
    ab[(size_t)n * C + c] = make_float2(a, __ldg(beta + c) - (float)m * a);
    sums[((size_t)n * C + c) * 2] = 0.0;               // self-cleaning scratch: the next norm finds zeros again
    sums[((size_t)n * C + c) * 2 + 1] = 0.0;
  }
  if (threadIdx.x == 0) tickets[n] = 0;
}

// Where a kernel writes: a channels-last tensor [N][H+2p][W+2p][Ctot] of which this producer owns the channels
// [coff, coff+C).  p = 1: the one-pixel border is filled with the reflection of the interior (what
// F.pad(mode="reflect") would produce for the following 3×3 convolution), so the convolution runs unpadded
// on the buffer and no separate padding pass (and no NCHW round trip) exists.
struct PadGeom {
  int H, W, pad, Ctot, coff;
};

template <typename T>
__device__ __forceinline__ void store_reflect(T* __restrict__ y, const PadGeom& g, int n, int h, int w,
                                              int c, const float (&v)[Vec<T>::N]) {
  const int Hp = g.H + 2 * g.pad, Wp = g.W + 2 * g.pad;
  int rows[3], cols[3], nr = 0, nc = 0;
  rows[nr++] = h + g.pad;
  cols[nc++] = w + g.pad;
  if (g.pad == 1) {
    if (h == 1) rows[nr++] = 0;
    if (h == g.H - 2) rows[nr++] = g.H + 1;
    if (w == 1) cols[nc++] = 0;
    if (w == g.W - 2) cols[nc++] = g.W + 1;
  }
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j)
      Vec<T>::store(y + (((size_t)n * Hp + rows[i]) * Wp + cols[j]) * g.Ctot + g.coff + c, v);
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256) in_apply(const T* __restrict__ x, const T* __restrict__ residual, int res_pad,
                                                const float2* __restrict__ ab, int C, PadGeom out, T* __restrict__ y) {
  constexpr int VN = Vec<T>::N;
  const int groups = C / VN, HW = out.H * out.W;
  const int n = blockIdx.y;
  // a thread keeps its channel group (groups divides the block size)
  const int g = threadIdx.x % groups;
  float a[VN], b[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    const float2 t = __ldg(ab + (size_t)n * C + g * VN + i);
    a[i] = t.x;
    b[i] = t.y;
  }
  const int px_step = gridDim.x * (blockDim.x / groups);
  for (int px = blockIdx.x * (blockDim.x / groups) + threadIdx.x / groups; px < HW; px += px_step) {
    const int h = px / out.W, w = px - h * out.W;
    float v[VN], r[VN];
    Vec<T>::load(x + ((size_t)n * HW + px) * C + g * VN, v);
    if (residual)
      Vec<T>::load(residual + (((size_t)n * (out.H + 2 * res_pad) + h + res_pad) * (out.W + 2 * res_pad) + w + res_pad) * C +
                       g * VN, r);
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      float o = fmaf(v[i], a[i], b[i]);
      if (residual) o += r[i];
      if (ACT == 1) o = fmaxf(o, 0.0f);
      if (ACT == 2) o = o > 0.0f ? o : expm1f(o);
      v[i] = o;
    }
    store_reflect<T>(y, out, n, h, w, g * VN, v);
  }
}

// mode 2: every second pixel (the input of a 1×1 stride-2 convolution, UNet.py:187-190);
// mode 0: copy (skip connection into its channel slice of the concatenation buffer, UNet.py:204-216 with equal
// sizes); mode 1: bilinear ×2, align_corners=True (UNet.py:128; torch's upsample_bilinear2d arithmetic in fp32).
// src: channels-last [N][Hs+2ps][Ws+2ps][C]; dst geometry as above (H, W = output size).
template <typename T, int MODE>
__global__ void __launch_bounds__(256) resample_pad(const T* __restrict__ src, int Hs, int Ws, int src_pad, int C,
                                                    PadGeom out, T* __restrict__ y) {
  constexpr int VN = Vec<T>::N;
  const int groups = C / VN;
  const size_t per_img = (size_t)out.H * out.W * groups;
  const int n = blockIdx.y;
  const int Wsp = Ws + 2 * src_pad;
  const T* base = src + (size_t)n * (Hs + 2 * src_pad) * Wsp * C;
  const float rh = out.H > 1 ? (float)(Hs - 1) / (float)(out.H - 1) : 0.0f;
  const float rw = out.W > 1 ? (float)(Ws - 1) / (float)(out.W - 1) : 0.0f;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < per_img; t += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(t % groups), px = (int)(t / groups);
    const int h = px / out.W, w = px - h * out.W;
    float v[VN];
    if (MODE == 0) {
      Vec<T>::load(base + ((size_t)(h + src_pad) * Wsp + w + src_pad) * C + g * VN, v);
    } else if (MODE == 2) {
      Vec<T>::load(base + ((size_t)(2 * h + src_pad) * Wsp + 2 * w + src_pad) * C + g * VN, v);
    } else {
      const float h1r = rh * h, w1r = rw * w;
      const int h1 = (int)h1r, w1 = (int)w1r;
      const int hp = h1 < Hs - 1 ? 1 : 0, wp = w1 < Ws - 1 ? 1 : 0;
      const float h1l = h1r - h1, h0l = 1.0f - h1l, w1l = w1r - w1, w0l = 1.0f - w1l;
      float a[VN], b[VN], c[VN], d[VN];
      const T* p = base + ((size_t)(h1 + src_pad) * Wsp + w1 + src_pad) * C + g * VN;
      Vec<T>::load(p, a);
      Vec<T>::load(p + (size_t)wp * C, b);
      Vec<T>::load(p + (size_t)hp * Wsp * C, c);
      Vec<T>::load(p + ((size_t)hp * Wsp + wp) * C, d);
#pragma unroll
      for (int i = 0; i < VN; ++i) v[i] = h0l * (w0l * a[i] + w1l * b[i]) + h1l * (w0l * c[i] + w1l * d[i]);
    }
    store_reflect<T>(y, out, n, h, w, g * VN, v);
  }
}

static dim3 apply_grid(int N, size_t per_img) {
  int blocks = (int)((per_img + 255) / 256);
  const int cap = (sm_count() * 8 + N - 1) / N;
  if (blocks > cap) blocks = cap;
  return dim3(blocks < 1 ? 1 : blocks, N);
}

template <typename T>
static int launch_norm(const void* x, const void* residual, int res_pad, int N, int C, const float* gamma,
                       const float* beta, float eps, int act, void* scratch, int scratch_nc, PadGeom out, void* y,
                       cudaStream_t st) {
  constexpr int VN = Vec<T>::N;
  const int groups = C / VN, HW = out.H * out.W;
  if (C % VN || groups > 256 || 256 % groups || out.Ctot % VN || out.coff % VN) {
    set_error("instance_norm: channel counts must be multiples of the vector width, C/width dividing 256", cudaSuccess);
    return GPNERF_E_UNSUPPORTED;
  }
  // scratch (zeroed once by the caller, self-cleaning afterwards – no memset node per norm):
  //   unsigned tickets[64] | double sums[scratch_nc * 2] | float2 ab[scratch_nc]
  unsigned* tickets = reinterpret_cast<unsigned*>(scratch);
  double* sums = reinterpret_cast<double*>(tickets + 64);
  float2* ab = reinterpret_cast<float2*>(sums + (size_t)scratch_nc * 2);
  // enough blocks to fill the machine, at least 8 pixels per thread row to amortise the atomics
  const int px_step = 256 / groups;
  int chunks = (sm_count() * 4 + N - 1) / N;
  int px_per_block = (HW + chunks - 1) / chunks;
  if (px_per_block < px_step * 8) px_per_block = px_step * 8;
  chunks = (HW + px_per_block - 1) / px_per_block;
  in_stats<T><<<dim3(chunks, N), 256, 0, st>>>((const T*)x, HW, C, px_per_block, sums, gamma, beta, eps, 1.0 / HW, ab, tickets);
  const dim3 grid = apply_grid(N, (size_t)HW * groups);
  const T* xx = (const T*)x;
  const T* rr = (const T*)residual;
  switch (act) {
    case 0: in_apply<T, 0><<<grid, 256, 0, st>>>(xx, rr, res_pad, ab, C, out, (T*)y); break;
    case 1: in_apply<T, 1><<<grid, 256, 0, st>>>(xx, rr, res_pad, ab, C, out, (T*)y); break;
    default: in_apply<T, 2><<<grid, 256, 0, st>>>(xx, rr, res_pad, ab, C, out, (T*)y); break;
  }
  return check_launch("instance_norm_act");
}

template <typename T>
static int launch_resample(const void* src, int N, int Hs, int Ws, int src_pad, int C, int mode, PadGeom out, void* y,
                           cudaStream_t st) {
  constexpr int VN = Vec<T>::N;
  if (C % VN || out.Ctot % VN || out.coff % VN) {
    set_error("resample_pad: channel counts must be multiples of the vector width", cudaSuccess);
    return GPNERF_E_UNSUPPORTED;
  }
  const dim3 grid = apply_grid(N, (size_t)out.H * out.W * (C / VN));
  if (mode == 0)
    resample_pad<T, 0><<<grid, 256, 0, st>>>((const T*)src, Hs, Ws, src_pad, C, out, (T*)y);
  else if (mode == 2)
    resample_pad<T, 2><<<grid, 256, 0, st>>>((const T*)src, Hs, Ws, src_pad, C, out, (T*)y);
  else
    resample_pad<T, 1><<<grid, 256, 0, st>>>((const T*)src, Hs, Ws, src_pad, C, out, (T*)y);
  return check_launch("resample_pad");
}

}  // namespace gpnerf

using namespace gpnerf;

extern "C" {

int gpnerf_k9_instance_norm_act(const void* x, const void* residual, int res_pad, int dtype, int N, int H, int W, int C,
                                const float* gamma, const float* beta, float eps, int act, void* scratch, int scratch_nc,
                                void* y, int y_pad, int y_ctot, int y_coff, void* stream) {
  GPNERF_REQUIRE(x && gamma && beta && scratch && y && N > 0 && N <= 64 && H > 0 && W > 0 && C > 0 && act >= 0 && act <= 2);
  GPNERF_REQUIRE((long long)N * C <= scratch_nc);
  GPNERF_REQUIRE(dtype >= 0 && dtype <= 2 && (y_pad == 0 || (y_pad == 1 && H >= 2 && W >= 2)) && (res_pad == 0 || res_pad == 1));
  GPNERF_REQUIRE(y_ctot >= y_coff + C && y_coff >= 0);
  cudaStream_t st = (cudaStream_t)stream;
  const PadGeom out{H, W, y_pad, y_ctot, y_coff};
  if (dtype == 0) return launch_norm<float>(x, residual, res_pad, N, C, gamma, beta, eps, act, scratch, scratch_nc, out, y, st);
  if (dtype == 1) return launch_norm<__nv_bfloat16>(x, residual, res_pad, N, C, gamma, beta, eps, act, scratch, scratch_nc, out, y, st);
  return launch_norm<__half>(x, residual, res_pad, N, C, gamma, beta, eps, act, scratch, scratch_nc, out, y, st);
}

int gpnerf_k9_resample_pad(const void* src, int dtype, int N, int Hs, int Ws, int src_pad, int C, int mode, void* y,
                           int H, int W, int y_pad, int y_ctot, int y_coff, void* stream) {
  GPNERF_REQUIRE(src && y && N > 0 && Hs > 0 && Ws > 0 && C > 0 && H > 0 && W > 0 && mode >= 0 && mode <= 2);
  GPNERF_REQUIRE(dtype >= 0 && dtype <= 2 && (y_pad == 0 || (y_pad == 1 && H >= 2 && W >= 2)) && (src_pad == 0 || src_pad == 1));
  GPNERF_REQUIRE(y_ctot >= y_coff + C && y_coff >= 0 && (mode != 0 || (H == Hs && W == Ws)) &&
                 (mode != 2 || (H == (Hs + 1) / 2 && W == (Ws + 1) / 2)));
  cudaStream_t st = (cudaStream_t)stream;
  const PadGeom out{H, W, y_pad, y_ctot, y_coff};
  if (dtype == 0) return launch_resample<float>(src, N, Hs, Ws, src_pad, C, mode, out, y, st);
  if (dtype == 1) return launch_resample<__nv_bfloat16>(src, N, Hs, Ws, src_pad, C, mode, out, y, st);
  return launch_resample<__half>(src, N, Hs, Ws, src_pad, C, mode, out, y, st);
}

}  // extern "C"
