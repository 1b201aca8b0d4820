// K3 (bf16 tensor-core path, precision = 1) – the density and colour heads as
// chains of tcgen05.mma GEMMs with fp32 accumulators in TMEM.
//
// trainhead.py:39-41 (128→64), :102-110 (134→64→32→16→1), :85-100,128-145
// (colour trunk).  One CTA = 128 threads = one tile of 128 sample points
// (UMMA M = 128: accumulator row i lives in TMEM lane i and is read back by
// thread i).  Per layer:
//     one thread issues K/16 tcgen05.mma (A = activations, B = weights, both
//     bf16 K-major in shared memory) and commits them to an mbarrier;
//     all threads wait, tcgen05.ld their accumulator row, add bias, apply ELU,
//     round to bf16 and store the row as the next layer's A operand.
// Activations never leave the SM; weights are packed once into the UMMA
// operand layout (gpnerf_k3_pack_weights) and fetched with one TMA bulk copy
// per CTA.  The last layer of each head (16→1, 16→3) runs on CUDA cores in the
// epilogue.  Two CTAs per SM overlap one tile's epilogue with the other's MMAs.
#include "common.cuh"
#include "tc_common.cuh"

namespace gpnerf {
using namespace tc;

// ---------------------------------------------------------------------------
// packed weight images
// ---------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t op_sbo(int Kp) { return (uint32_t)(Kp / 8) * kLBO; }
__host__ __device__ constexpr uint32_t op_bytes(int rows, int Kp) { return (uint32_t)(rows / 8) * op_sbo(Kp); }

struct DenImg {   // density head
  static constexpr uint32_t Wg = 0;                                 // [64 x 128]
  static constexpr uint32_t W0 = Wg + op_bytes(64, 128);            // [64 x 144] (134 padded)
  static constexpr uint32_t W1 = W0 + op_bytes(64, 144);            // [32 x 64]
  static constexpr uint32_t W2 = W1 + op_bytes(32, 64);             // [16 x 32]
  static constexpr uint32_t F32 = W2 + op_bytes(16, 32);            // floats below
  static constexpr int bg = 0, b0 = 64, b1 = 128, b2 = 160, w3 = 176, b3 = 192, NF = 196;
  static constexpr uint32_t BYTES = F32 + NF * 4;
};
static_assert(DenImg::BYTES % 16 == 0, "bulk copy needs 16-byte multiples");

template <int V>
struct ColImg {   // colour head; base_fc.0 is split into its [mean|var] and per-view column blocks
  static constexpr uint32_t Wb0a = 0;                               // [64 x 80]  cols 0..69 of base_fc.0
  static constexpr uint32_t Wb0b = Wb0a + op_bytes(64, 80);         // [64 x 48]  cols 70..104
  static constexpr uint32_t Wb1 = Wb0b + op_bytes(64, 48);          // [32 x 64]
  static constexpr uint32_t Wv0 = Wb1 + op_bytes(32, 64);           // [32 x 32]
  static constexpr uint32_t Wv1 = Wv0 + op_bytes(32, 32);           // [32 x 32]
  static constexpr uint32_t Wr0 = Wv1 + op_bytes(32, 32);           // [32 x 32V]
  static constexpr uint32_t Wr1 = Wr0 + op_bytes(32, 32 * V);       // [16 x 32]
  static constexpr uint32_t F32 = Wr1 + op_bytes(16, 32);
  static constexpr int bb0 = 0, bb1 = 64, vb0 = 96, vb1 = 128, rb0 = 160, rb1 = 192, rw2 = 208, rb2 = 256, NF = 260;
  static constexpr uint32_t BYTES = F32 + NF * 4;
};

constexpr uint32_t kColImgMax = ColImg<4>::BYTES;
constexpr uint32_t kImageBytes = ((DenImg::BYTES + 127) / 128) * 128 + ((kColImgMax + 127) / 128) * 128;
constexpr uint32_t kColImgOffset = ((DenImg::BYTES + 127) / 128) * 128;

// W fp32 [N][ldw] columns [c0, c0+K) → bf16 operand [N x Kp] at dst
__device__ void pack_operand(uint8_t* dst, const float* __restrict__ W, int N, int ldw, int c0, int K, int Kp) {
  const uint32_t sbo = op_sbo(Kp);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * Kp; i += gridDim.x * blockDim.x) {
    int n = i / Kp, k = i - n * Kp;
    float v = (k < K) ? __ldg(W + (long long)n * ldw + c0 + k) : 0.0f;
    *reinterpret_cast<__nv_bfloat16*>(dst + chunk_off(n, k >> 3, sbo) + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
}
__device__ void pack_floats(float* dst, const float* __restrict__ src, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = __ldg(src + i);
}

struct RawW {
  const float *geo_w, *geo_b, *den_w[4], *den_b[4], *base_w[2], *base_b[2], *vis_w[2], *vis_b[2], *rgb_w[3], *rgb_b[3];
};

template <int V>
__device__ void pack_color(uint8_t* c, const RawW& w) {
  using I = ColImg<V>;
  pack_operand(c + I::Wb0a, w.base_w[0], 64, 105, 0, 70, 80);
  pack_operand(c + I::Wb0b, w.base_w[0], 64, 105, 70, 35, 48);
  pack_operand(c + I::Wb1, w.base_w[1], 32, 64, 0, 64, 64);
  pack_operand(c + I::Wv0, w.vis_w[0], 32, 32, 0, 32, 32);
  pack_operand(c + I::Wv1, w.vis_w[1], 32, 32, 0, 32, 32);
  pack_operand(c + I::Wr0, w.rgb_w[0], 32, 32 * V, 0, 32 * V, 32 * V);
  pack_operand(c + I::Wr1, w.rgb_w[1], 16, 32, 0, 32, 32);
  float* f = reinterpret_cast<float*>(c + I::F32);
  pack_floats(f + I::bb0, w.base_b[0], 64);
  pack_floats(f + I::bb1, w.base_b[1], 32);
  pack_floats(f + I::vb0, w.vis_b[0], 32);
  pack_floats(f + I::vb1, w.vis_b[1], 32);
  pack_floats(f + I::rb0, w.rgb_b[0], 32);
  pack_floats(f + I::rb1, w.rgb_b[1], 16);
  pack_floats(f + I::rw2, w.rgb_w[2], 48);
  pack_floats(f + I::rb2, w.rgb_b[2], 3);
}

__global__ void pack_weights_kernel(RawW w, int V, uint8_t* image) {
  uint8_t* d = image;
  pack_operand(d + DenImg::Wg, w.geo_w, 64, 128, 0, 128, 128);
  pack_operand(d + DenImg::W0, w.den_w[0], 64, 134, 0, 134, 144);
  pack_operand(d + DenImg::W1, w.den_w[1], 32, 64, 0, 64, 64);
  pack_operand(d + DenImg::W2, w.den_w[2], 16, 32, 0, 32, 32);
  float* f = reinterpret_cast<float*>(d + DenImg::F32);
  pack_floats(f + DenImg::bg, w.geo_b, 64);
  pack_floats(f + DenImg::b0, w.den_b[0], 64);
  pack_floats(f + DenImg::b1, w.den_b[1], 32);
  pack_floats(f + DenImg::b2, w.den_b[2], 16);
  pack_floats(f + DenImg::w3, w.den_w[3], 16);
  pack_floats(f + DenImg::b3, w.den_b[3], 1);
  uint8_t* c = image + kColImgOffset;
  switch (V) {
    case 1: pack_color<1>(c, w); break;
    case 2: pack_color<2>(c, w); break;
    case 3: pack_color<3>(c, w); break;
    default: pack_color<4>(c, w); break;
  }
}

// ---------------------------------------------------------------------------
// shared pieces of the two head kernels
// ---------------------------------------------------------------------------
// 8 consecutive fp32 of a row (zero beyond kmax) → one bf16 chunk
__device__ __forceinline__ void load_chunk(const float* __restrict__ row, int k0, int kmax, float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (row != nullptr && k0 + j < kmax) ? __ldg(row + k0 + j) : 0.0f;
}

// accumulator row (N columns starting at `col`) → bias + ELU → bf16 row of the next operand
template <int N>
__device__ __forceinline__ void epilogue_elu_to_tile(uint32_t taddr, const float* __restrict__ bias,
                                                     uint8_t* tile, uint32_t sbo, int row, int kc0) {
  static_assert(N % 16 == 0, "");
  if constexpr (N % 32 == 0) {
#pragma unroll
    for (int c = 0; c < N / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = elu_fast(__uint_as_float(r[j * 8 + e]) + bias[c * 32 + j * 8 + e]);
        st_chunk(tile, chunk_off(row, kc0 + c * 4 + j, sbo), v);
      }
    }
  } else {
    uint32_t r[16];
    tmem_ld16(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = elu_fast(__uint_as_float(r[j * 8 + e]) + bias[j * 8 + e]);
      st_chunk(tile, chunk_off(row, kc0 + j, sbo), v);
    }
  }
}

// make the operand stores visible to the tensor core, order the TMEM reads
// before the next MMAs, and meet
__device__ __forceinline__ void round_sync() {
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

__device__ __forceinline__ void wait_round(uint64_t* bar, uint32_t& phase) {
  mbar_wait(bar, phase);
  phase ^= 1u;
  tc_fence_after();
}

// ---------------------------------------------------------------------------
// density head
// ---------------------------------------------------------------------------
struct DenSmem {
  static constexpr uint32_t IMG = 0;
  static constexpr uint32_t A0 = ((DenImg::BYTES + 127) / 128) * 128;     // [128 x 128] / later [128 x 64]
  static constexpr uint32_t A1 = A0 + op_bytes(128, 128);                 // [128 x 144] / later [128 x 32]
  static constexpr uint32_t BAR = A1 + op_bytes(128, 144);                // 2 mbarriers + tmem ptr
  static constexpr uint32_t BYTES = BAR + 64;
};

__global__ void __launch_bounds__(128, 2) density_mlp_tc(const float* __restrict__ feat_in, int input_kind,
                                                         const float* __restrict__ meanvar,
                                                         const float* __restrict__ mask, int V,
                                                         const uint8_t* __restrict__ image,
                                                         const int32_t* __restrict__ count_ptr, int n_const,
                                                         float* __restrict__ sigma,
                                                         float* __restrict__ sigma_feat_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* img = smem + DenSmem::IMG;
  uint8_t* A0 = smem + DenSmem::A0;
  uint8_t* A1 = smem + DenSmem::A1;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + DenSmem::BAR);
  uint64_t* bar_m = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const float* fl = reinterpret_cast<const float*>(img + DenImg::F32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(bar_w, DenImg::BYTES);
    bulk_g2s(img, image, DenImg::BYTES, bar_w);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's lane quarter
  mbar_wait(bar_w, 0);                                             // weights landed

  const uint32_t a0 = smem_u32(A0), a1 = smem_u32(A1), wimg = smem_u32(img);
  uint32_t phase = 0;
  const int n = count_ptr ? __ldg(count_ptr) : n_const;
  const int n_tiles = (n + 127) / 128;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long first = (long long)tile * 128;
    const int n_valid = min(128, n - (int)first);
    // ---- stage the tile: fp32 rows → bf16 operands (8 rows x 4 chunks per warp step)
    {
      const int rsub = lane & 7, csub = lane >> 3;
      for (int rg = warp; rg < 16; rg += 4) {
        const int r = rg * 8 + rsub;
        const bool ok = r < n_valid;
        const float* frow = ok ? feat_in + (first + r) * (input_kind == 0 ? 128 : 64) : nullptr;
        const float* mrow = ok ? meanvar + (first + r) * 70 : nullptr;
        float v[8];
        if (input_kind == 0) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int kc = it * 4 + csub;
            load_chunk(frow, kc * 8, 128, v);
            st_chunk(A0, chunk_off(r, kc, op_sbo(128)), v);
          }
        } else {
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int kc = it * 4 + csub;
            load_chunk(frow, kc * 8, 64, v);
            st_chunk(A1, chunk_off(r, kc, op_sbo(144)), v);
          }
        }
        for (int kc = csub; kc < 10; kc += 4) {       // [mean|var] → columns 64..143 (70 real, 10 zero)
          load_chunk(mrow, kc * 8, 70, v);
          st_chunk(A1, chunk_off(r, 8 + kc, op_sbo(144)), v);
        }
      }
    }
    const int row = tid;
    if (input_kind == 0) {
      // sigmahead.out_geometry_fc: [128 x 128] · Wgᵀ → 64, ELU → columns 0..63 of A1
      round_sync();
      if (tid == 0) issue_gemm(a0, op_sbo(128), wimg + DenImg::Wg, op_sbo(128), 128, 64, tmem, bar_m);
      wait_round(bar_m, phase);
      epilogue_elu_to_tile<64>(t_row, fl + DenImg::bg, A1, op_sbo(144), row, 0);
    }
    if (sigma_feat_out != nullptr && input_kind == 0 && row < n_valid) {
      // module API only: sigma_feat as the bf16 values the next layer consumes
      float* o = sigma_feat_out + (first + row) * 64;
      for (int kc = 0; kc < 8; ++kc) {
        uint4 q = *reinterpret_cast<const uint4*>(A1 + chunk_off(row, kc, op_sbo(144)));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f2 = __bfloat1622float2(h[e]);
          o[kc * 8 + e * 2] = f2.x;
          o[kc * 8 + e * 2 + 1] = f2.y;
        }
      }
    }
    // out_geometry_fc.0: [128 x 144] · W0ᵀ → 64, ELU → A0 as [128 x 64]
    round_sync();
    if (tid == 0) issue_gemm(a1, op_sbo(144), wimg + DenImg::W0, op_sbo(144), 144, 64, tmem, bar_m);
    wait_round(bar_m, phase);
    epilogue_elu_to_tile<64>(t_row, fl + DenImg::b0, A0, op_sbo(64), row, 0);
    // .2: [128 x 64] → 32, ELU → A1 as [128 x 32]
    round_sync();
    if (tid == 0) issue_gemm(a0, op_sbo(64), wimg + DenImg::W1, op_sbo(64), 64, 32, tmem, bar_m);
    wait_round(bar_m, phase);
    epilogue_elu_to_tile<32>(t_row, fl + DenImg::b1, A1, op_sbo(32), row, 0);
    // .4: [128 x 32] → 16, ELU ; .6: 16 → 1 on CUDA cores, ReLU, no-valid-view fill
    round_sync();
    if (tid == 0) issue_gemm(a1, op_sbo(32), wimg + DenImg::W2, op_sbo(32), 32, 16, tmem, bar_m);
    wait_round(bar_m, phase);
    {
      uint32_t r[16];
      tmem_ld16(t_row, r);
      tmem_wait_ld();
      float s = fl[DenImg::b3];
#pragma unroll
      for (int k = 0; k < 16; ++k) s = fmaf(elu_fast(__uint_as_float(r[k]) + fl[DenImg::b2 + k]), fl[DenImg::w3 + k], s);
      s = fmaxf(s, 0.0f);
      if (row < n_valid) {
        float nv = 0.0f;
        for (int v = 0; v < V; ++v) nv += __ldg(mask + (first + row) * V + v);
        sigma[first + row] = (nv < 1.0f) ? 0.0f : s;
      }
    }
    // the next tile's staging overwrites A0/A1: their last MMA readers completed
    // (bar_m), and the TMEM reads above are ordered by the next round_sync.
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------------------
// colour head
// ---------------------------------------------------------------------------
template <int V>
struct ColSmem {
  static constexpr uint32_t IMG = 0;
  static constexpr uint32_t G = ((ColImg<V>::BYTES + 127) / 128) * 128;  // [128 x 80] mean|var ; later Y_v / Z
  static constexpr uint32_t H = G + op_bytes(128, 80);                   // [128 x 64] base_fc hidden
  static constexpr uint32_t F = H + op_bytes(128, 64);                   // V x [128 x 48] ; later Xs_v / flat
  static constexpr uint32_t BAR = F + V * op_bytes(128, 48);
  static constexpr uint32_t BYTES = BAR + 64;
  static constexpr uint32_t TMEM_COLS = (V <= 2) ? 64 : 128;
};

template <int V>
__global__ void __launch_bounds__(128, 2) color_mlp_tc(const float* __restrict__ rgb_feat,
                                                       const float* __restrict__ meanvar,
                                                       const int32_t* __restrict__ valid1,
                                                       const uint8_t* __restrict__ image,
                                                       const int32_t* __restrict__ count_ptr, int n_const,
                                                       float* __restrict__ rgb) {
  using I = ColImg<V>;
  using S = ColSmem<V>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* img = smem + S::IMG;
  uint8_t* G = smem + S::G;
  uint8_t* Hb = smem + S::H;
  uint8_t* F = smem + S::F;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + S::BAR);
  uint64_t* bar_m = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const float* fl = reinterpret_cast<const float*>(img + I::F32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(bar_w, I::BYTES);
    bulk_g2s(img, image, I::BYTES, bar_w);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  mbar_wait(bar_w, 0);

  const uint32_t g_a = smem_u32(G), h_a = smem_u32(Hb), f_a = smem_u32(F), wimg = smem_u32(img);
  constexpr uint32_t F_STRIDE = op_bytes(128, 48);     // per-view feature operand
  constexpr uint32_t X_STRIDE = op_bytes(128, 32);     // per-view [128 x 32] operand
  uint32_t phase = 0;
  const int n = count_ptr ? __ldg(count_ptr) : n_const;
  const int n_tiles = (n + 127) / 128;
  const float inv_v = 1.0f / (float)V;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long first = (long long)tile * 128;
    const int n_valid = min(128, n - (int)first);
    // ---- stage: [mean|var] → G (70 → 80), per-view [rgb|feat] → F_v (35 → 48)
    {
      const int rsub = lane & 7, csub = lane >> 3;
      for (int rg = warp; rg < 16; rg += 4) {
        const int r = rg * 8 + rsub;
        const float *mrow = nullptr, *frow = nullptr;
        if (r < n_valid) {
          const long long src = valid1 ? (long long)__ldg(valid1 + first + r) : first + r;
          mrow = meanvar + src * 70;
          frow = rgb_feat + src * (V * 35);
        }
        float v[8];
        for (int kc = csub; kc < 10; kc += 4) {
          load_chunk(mrow, kc * 8, 70, v);
          st_chunk(G, chunk_off(r, kc, op_sbo(80)), v);
        }
        for (int t = csub; t < 6 * V; t += 4) {
          const int vw = t / 6, kc = t - vw * 6;
          load_chunk(frow ? frow + vw * 35 : nullptr, kc * 8, 35, v);
          st_chunk(F + vw * F_STRIDE, chunk_off(r, kc, op_sbo(48)), v);
        }
      }
    }
    const int row = tid;
    float x[V][32];                                   // base_fc output per view (residual branch)
#pragma unroll
    for (int vw = 0; vw < V; ++vw) {
      // base_fc.0 on [mean|var|feat_v]: two accumulating GEMMs (K = 80 + 48) → 64
      round_sync();
      if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, 64);
        for (int k16 = 0; k16 < 5; ++k16)
          umma_bf16(tmem, make_smem_desc(g_a + k16 * 2 * kLBO, kLBO, op_sbo(80)),
                    make_smem_desc(wimg + I::Wb0a + k16 * 2 * kLBO, kLBO, op_sbo(80)), idesc, k16 > 0);
        for (int k16 = 0; k16 < 3; ++k16)
          umma_bf16(tmem, make_smem_desc(f_a + vw * F_STRIDE + k16 * 2 * kLBO, kLBO, op_sbo(48)),
                    make_smem_desc(wimg + I::Wb0b + k16 * 2 * kLBO, kLBO, op_sbo(48)), idesc, 1u);
        umma_commit(bar_m);
      }
      wait_round(bar_m, phase);
      epilogue_elu_to_tile<64>(t_row, fl + I::bb0, Hb, op_sbo(64), row, 0);
      // base_fc.2: 64 → 32, ELU ; keep x_v in registers, stage x_v / V for vis_fc
      round_sync();
      if (tid == 0) issue_gemm(h_a, op_sbo(64), wimg + I::Wb1, op_sbo(64), 64, 32, tmem, bar_m);
      wait_round(bar_m, phase);
      {
        uint32_t r[32];
        tmem_ld32(t_row, r);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) x[vw][k] = elu_fast(__uint_as_float(r[k]) + fl[I::bb1 + k]);
      }
    }
    // all F_v consumed (last base_fc.0 committed and waited) → reuse F for Xs_v = x_v / V
#pragma unroll
    for (int vw = 0; vw < V; ++vw)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = x[vw][j * 8 + e] * inv_v;
        st_chunk(F + vw * X_STRIDE, chunk_off(row, j, op_sbo(32)), v);
      }
    // vis_fc.0 for all views at once: V GEMMs into V column blocks
    round_sync();
    if (tid == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 32);
      for (int vw = 0; vw < V; ++vw)
        for (int k16 = 0; k16 < 2; ++k16)
          umma_bf16(tmem + vw * 32, make_smem_desc(f_a + vw * X_STRIDE + k16 * 2 * kLBO, kLBO, op_sbo(32)),
                    make_smem_desc(wimg + I::Wv0 + k16 * 2 * kLBO, kLBO, op_sbo(32)), idesc, k16 > 0);
      umma_commit(bar_m);
    }
    wait_round(bar_m, phase);
#pragma unroll
    for (int vw = 0; vw < V; ++vw)   // Y_v → G region (G and H are free now)
      epilogue_elu_to_tile<32>(t_row + vw * 32, fl + I::vb0, G + vw * X_STRIDE, op_sbo(32), row, 0);
    // vis_fc.2, residual, flatten view-major → flat [128 x 32V] in F
    round_sync();
    if (tid == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 32);
      for (int vw = 0; vw < V; ++vw)
        for (int k16 = 0; k16 < 2; ++k16)
          umma_bf16(tmem + vw * 32, make_smem_desc(g_a + vw * X_STRIDE + k16 * 2 * kLBO, kLBO, op_sbo(32)),
                    make_smem_desc(wimg + I::Wv1 + k16 * 2 * kLBO, kLBO, op_sbo(32)), idesc, k16 > 0);
      umma_commit(bar_m);
    }
    wait_round(bar_m, phase);
#pragma unroll
    for (int vw = 0; vw < V; ++vw) {
      uint32_t r[32];
      tmem_ld32(t_row + vw * 32, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          v[e] = x[vw][j * 8 + e] + elu_fast(__uint_as_float(r[j * 8 + e]) + fl[I::vb1 + j * 8 + e]);
        st_chunk(F, chunk_off(row, vw * 4 + j, op_sbo(32 * V)), v);
      }
    }
    // rgb_fc.0: 32V → 32, ELU → Z in G
    round_sync();
    if (tid == 0) issue_gemm(f_a, op_sbo(32 * V), wimg + I::Wr0, op_sbo(32 * V), 32 * V, 32, tmem, bar_m);
    wait_round(bar_m, phase);
    epilogue_elu_to_tile<32>(t_row, fl + I::rb0, G, op_sbo(32), row, 0);
    // rgb_fc.2: 32 → 16, ELU ; rgb_fc.4: 16 → 3 on CUDA cores ; sigmoid
    round_sync();
    if (tid == 0) issue_gemm(g_a, op_sbo(32), wimg + I::Wr1, op_sbo(32), 32, 16, tmem, bar_m);
    wait_round(bar_m, phase);
    {
      uint32_t r[16];
      tmem_ld16(t_row, r);
      tmem_wait_ld();
      float o0 = fl[I::rb2], o1 = fl[I::rb2 + 1], o2 = fl[I::rb2 + 2];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float h = elu_fast(__uint_as_float(r[k]) + fl[I::rb1 + k]);
        o0 = fmaf(h, fl[I::rw2 + k], o0);
        o1 = fmaf(h, fl[I::rw2 + 16 + k], o1);
        o2 = fmaf(h, fl[I::rw2 + 32 + k], o2);
      }
      if (row < n_valid) {
        const long long dst = valid1 ? (long long)__ldg(valid1 + first + row) : first + row;
        rgb[dst * 3 + 0] = 1.0f / (1.0f + __expf(-o0));
        rgb[dst * 3 + 1] = 1.0f / (1.0f + __expf(-o1));
        rgb[dst * 3 + 2] = 1.0f / (1.0f + __expf(-o2));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, S::TMEM_COLS);
}

}  // namespace gpnerf

using namespace gpnerf;

int gpnerf_density_mlp_tc(const float* feat_in, const float* meanvar, const float* mask,
                          const gpnerf_head_weights_t* w, int n_views, int input_kind, int n_points_max,
                          const int32_t* count_ptr, float* sigma, float* sigma_feat, cudaStream_t st) {
  if (w->tc_image == nullptr) {
    set_error("weights->tc_image is NULL: call gpnerf_k3_pack_weights first", cudaSuccess);
    return GPNERF_E_ARG;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(density_mlp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, DenSmem::BYTES);
    if (e != cudaSuccess) {
      set_error("density_mlp_tc smem attribute", e);
      return GPNERF_E_CUDA;
    }
    attr_set = true;
  }
  int tiles = (n_points_max + 127) / 128;
  int grid = tiles < 2 * sm_count() ? tiles : 2 * sm_count();
  density_mlp_tc<<<grid, 128, DenSmem::BYTES, st>>>(feat_in, input_kind, meanvar, mask, n_views,
                                                    reinterpret_cast<const uint8_t*>(w->tc_image), count_ptr,
                                                    n_points_max, sigma, sigma_feat);
  return check_launch("k3_density_mlp (tcgen05)");
}

template <int V>
static int launch_color_tc(const float* rgb_feat, const float* meanvar, const int32_t* valid1, const uint8_t* image,
                           int n_points_max, const int32_t* count_ptr, float* rgb, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(color_mlp_tc<V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ColSmem<V>::BYTES);
    if (e != cudaSuccess) {
      set_error("color_mlp_tc smem attribute", e);
      return GPNERF_E_CUDA;
    }
    attr_set = true;
  }
  int tiles = (n_points_max + 127) / 128;
  int grid = tiles < 2 * sm_count() ? tiles : 2 * sm_count();
  color_mlp_tc<V><<<grid, 128, ColSmem<V>::BYTES, st>>>(rgb_feat, meanvar, valid1, image + kColImgOffset, count_ptr,
                                                        n_points_max, rgb);
  return check_launch("k3_color_mlp (tcgen05)");
}

int gpnerf_color_mlp_tc(const float* rgb_feat, const float* meanvar, const int32_t* valid1,
                        const gpnerf_head_weights_t* w, int n_views, int n_points_max, const int32_t* count_ptr,
                        float* rgb, cudaStream_t st) {
  if (w->tc_image == nullptr) {
    set_error("weights->tc_image is NULL: call gpnerf_k3_pack_weights first", cudaSuccess);
    return GPNERF_E_ARG;
  }
  const uint8_t* image = reinterpret_cast<const uint8_t*>(w->tc_image);
  switch (n_views) {
    case 1: return launch_color_tc<1>(rgb_feat, meanvar, valid1, image, n_points_max, count_ptr, rgb, st);
    case 2: return launch_color_tc<2>(rgb_feat, meanvar, valid1, image, n_points_max, count_ptr, rgb, st);
    case 3: return launch_color_tc<3>(rgb_feat, meanvar, valid1, image, n_points_max, count_ptr, rgb, st);
    case 4: return launch_color_tc<4>(rgb_feat, meanvar, valid1, image, n_points_max, count_ptr, rgb, st);
    default:
      set_error("tcgen05 colour head supports 1..4 source views", cudaSuccess);
      return GPNERF_E_UNSUPPORTED;
  }
}

extern "C" {

int64_t gpnerf_k3_packed_weight_bytes(void) { return (int64_t)kImageBytes; }

int gpnerf_k3_pack_weights(const gpnerf_head_weights_t* w, int n_views, void* image, void* stream) {
  GPNERF_REQUIRE(w && image && n_views >= 1 && n_views <= 4);
  GPNERF_REQUIRE(w->geo_w && w->geo_b);
  RawW r;
  r.geo_w = w->geo_w; r.geo_b = w->geo_b;
  for (int i = 0; i < 4; ++i) { GPNERF_REQUIRE(w->den_w[i] && w->den_b[i]); r.den_w[i] = w->den_w[i]; r.den_b[i] = w->den_b[i]; }
  for (int i = 0; i < 2; ++i) { GPNERF_REQUIRE(w->base_w[i] && w->base_b[i] && w->vis_w[i] && w->vis_b[i]);
    r.base_w[i] = w->base_w[i]; r.base_b[i] = w->base_b[i]; r.vis_w[i] = w->vis_w[i]; r.vis_b[i] = w->vis_b[i]; }
  for (int i = 0; i < 3; ++i) { GPNERF_REQUIRE(w->rgb_w[i] && w->rgb_b[i]); r.rgb_w[i] = w->rgb_w[i]; r.rgb_b[i] = w->rgb_b[i]; }
  pack_weights_kernel<<<32, 256, 0, (cudaStream_t)stream>>>(r, n_views, reinterpret_cast<uint8_t*>(image));
  return check_launch("k3_pack_weights");
}

}  // extern "C"
