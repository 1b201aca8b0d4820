// K3 (bf16 tensor-core path, precision = 1) – the density and colour heads as
// chains of tcgen05.mma GEMMs with fp32 accumulators in TMEM.
//
// trainhead.py:39-41 (128→64), :102-110 (134→64→32→16→1), :85-100,128-145
// (colour trunk).  One CTA works on a tile of 128 sample points (UMMA M = 128:
// accumulator row i lives in TMEM lane i).  Per layer:
//     one thread issues K/16 tcgen05.mma (A = activations, B = weights, both
//     bf16 K-major in shared memory) and commits them to an mbarrier;
//     all threads wait, tcgen05.ld their part of the accumulator row, add bias,
//     apply ELU, round to bf16 and store it as the next layer's A operand.
// Activations never leave the SM; weights are packed once into the UMMA
// operand layout (gpnerf_k3_pack_weights) and fetched with one TMA bulk copy
// per CTA.  The last layer of each head (16→1, 16→3) runs on CUDA cores in the
// epilogue.  Two CTAs per SM overlap one tile's epilogue with the other's MMAs.
//
// This file: weight packing, the stand-alone density head (operator API) and
// the colour head (operator API from fp32 rows, engine path from the bf16
// records the fused gather+density kernel of k23_fused_tc.cu leaves behind).
#include <stdlib.h>

#include "tc_heads.cuh"

namespace gpnerf {

// W fp32 [N][ldw] → 16-bit operand [N x Kp] at dst.  Column j of the operand is
//   scale · W[:, colmap(j)]            colmap(j) >= 0   (scale applies to columns j >= scaled_from only)
//   0                                  colmap(j) == kZero
//   bf16 hi / lo halves of c · bias    colmap(j) == kBiasHi / kBiasLo
// stored as fp16 for j < n_f16_cols and as bf16 otherwise.
constexpr int kZero = -1, kBiasHi = -2, kBiasLo = -3;
template <class Map>
__device__ void pack_operand(uint8_t* dst, const float* __restrict__ W, const float* __restrict__ bias, int N,
                             int ldw, int Kp, float scale, int scaled_from, int n_f16_cols, Map colmap) {
  const uint32_t sbo = op_sbo(Kp);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * Kp; i += gridDim.x * blockDim.x) {
    int n = i / Kp, k = i - n * Kp;
    const int src = colmap(k);
    float v = 0.0f;
    if (src >= 0) {
      v = (k >= scaled_from ? scale : 1.0f) * __ldg(W + (long long)n * ldw + src);
    } else if (src == kBiasHi || src == kBiasLo) {
      const float b = kLog2e * __ldg(bias + n);
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      v = (src == kBiasHi) ? hi : b - hi;
    }
    uint8_t* at = dst + chunk_off(n, k >> 3, sbo) + (k & 7) * 2;
    if (k < n_f16_cols)
      *reinterpret_cast<__half*>(at) = __float2half_rn(v);
    else
      *reinterpret_cast<__nv_bfloat16*>(at) = __float2bfloat16_rn(v);
  }
}
__device__ void pack_floats(float* dst, const float* __restrict__ src, int n, float scale) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = scale * __ldg(src + i);
}
// identity columns [0, K) followed by a bias K block (ones in columns K+6, K+7 of the A side)
struct IdBiasMap {
  int K;
  __device__ __forceinline__ int operator()(int j) const {
    return j < K ? j : (j == K + kBiasColHi ? kBiasHi : (j == K + kBiasColLo ? kBiasLo : kZero));
  }
};
// G-tile order with the bias in the tile's constant-one columns 70, 71
__device__ __forceinline__ int gmap_bias(int j) { return j == 70 ? kBiasHi : (j == 71 ? kBiasLo : gmap(j)); }

struct RawW {
  const float *geo_w, *geo_b, *den_w[4], *den_b[4], *base_w[2], *base_b[2], *vis_w[2], *vis_b[2], *rgb_w[3], *rgb_b[3];
};

template <int V>
__device__ void pack_color(uint8_t* c, const RawW& w) {
  using I = ColImg<V>;
  pack_operand(c + I::Wb0a, w.base_w[0], w.base_b[0], 64, 105, 80, kLog2e, 0, 0, [](int j) { return gmap_bias(j); });
  pack_operand(c + I::Wb0b, w.base_w[0], nullptr, 64, 105, 48, kLog2e, 0, 0,
               [](int j) { int s = fmap(j); return s < 0 ? kZero : 70 + s; });
  pack_operand(c + I::Wb1, w.base_w[1], w.base_b[1], 32, 64, 80, 1.0f, 0, 0, IdBiasMap{64});
  pack_operand(c + I::Wv0, w.vis_w[0], w.vis_b[0], 32, 32, 48, 1.0f, 0, 0, IdBiasMap{32});
  pack_operand(c + I::Wv1, w.vis_w[1], w.vis_b[1], 32, 32, 48, 1.0f, 0, 0, IdBiasMap{32});
  pack_operand(c + I::Wr0, w.rgb_w[0], w.rgb_b[0], 32, 32 * V, 32 * V + 16, 1.0f, 0, 0, IdBiasMap{32 * V});
  pack_operand(c + I::Wr1, w.rgb_w[1], w.rgb_b[1], 16, 32, 48, 1.0f, 0, 0, IdBiasMap{32});
  for (int v = 0; v < V; ++v)
    pack_operand(c + I::Wb0r + v * op_bytes(64, 16), w.base_w[0], nullptr, 64, 105, 16, kLog2e, 0, 0,
                 [v](int j) { return (j >= 3 * v && j < 3 * v + 3) ? 70 + (j - 3 * v) : kZero; });
  pack_operand(c + I::Wr0x, w.rgb_w[0], nullptr, 32, 32 * V, 32 * V, (float)V, 0, 0, [](int j) { return j; });
  float* f = reinterpret_cast<float*>(c + I::F32);
  pack_floats(f + I::rw2, w.rgb_w[2], 48, 1.0f / kLog2e);
  pack_floats(f + I::rb2, w.rgb_b[2], 3, 1.0f);
  pack_floats(f + I::rb0c, w.rgb_b[0], 32, kLog2e);
  pack_floats(f + I::w1f, w.rgb_w[1], 512, 1.0f / kLog2e);
  pack_floats(f + I::b1f, w.rgb_b[1], 16, 1.0f);
  pack_floats(f + I::rw2u, w.rgb_w[2], 48, 1.0f);
}

__global__ void pack_weights_kernel(RawW w, int V, uint8_t* image) {
  uint8_t* d = image;
  pack_operand(d + DenImg::Wg, w.geo_w, w.geo_b, 64, 128, 144, kLog2e, 0, 128, IdBiasMap{128});
  pack_operand(d + DenImg::W0, w.den_w[0], w.den_b[0], 64, 134, 144, kLog2e, 64, 0, [](int j) {
    if (j < 64) return j;                       // sigma_feat columns: inputs arrive scaled, weights as they are
    int s = gmap_bias(j - 64);
    return s >= 0 ? 64 + s : s;
  });
  pack_operand(d + DenImg::W1, w.den_w[1], w.den_b[1], 32, 64, 80, 1.0f, 0, 0, IdBiasMap{64});
  pack_operand(d + DenImg::W2, w.den_w[2], w.den_b[2], 16, 32, 48, 1.0f, 0, 0, IdBiasMap{32});
  float* f = reinterpret_cast<float*>(d + DenImg::F32);
  pack_floats(f + DenImg::w3, w.den_w[3], 16, 1.0f / kLog2e);
  pack_floats(f + DenImg::b3, w.den_b[3], 1, 1.0f);
  uint8_t* c = image + kColImgOffset;
  switch (V) {
    case 1: pack_color<1>(c, w); break;
    case 2: pack_color<2>(c, w); break;
    case 3: pack_color<3>(c, w); break;
    default: pack_color<4>(c, w); break;
  }
}

// ---------------------------------------------------------------------------
// stand-alone density head (operator API: fp32 rows in the reference layout)
// ---------------------------------------------------------------------------
struct DenSmem {
  static constexpr uint32_t IMG = 0;
  static constexpr uint32_t A0 = ((DenImg::BYTES + 127) / 128) * 128;     // [128 x 128] / later [128 x 64] + [128 x 32]
  static constexpr uint32_t A1 = A0 + op_bytes(128, 128);                 // [128 x 144]
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
  uint8_t* A2 = A0 + op_bytes(128, 64);            // [128 x 32] behind the [128 x 64] tile
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

  const uint32_t a0 = smem_u32(A0), a1 = smem_u32(A1), a2 = smem_u32(A2), wimg = smem_u32(img);
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
            load_chunk_mapped(frow, kc, IdMap{128}, v);
            st_chunk_f16(A0, chunk_off(r, kc, op_sbo(128)), v);      // the Wg operand pair is fp16
          }
        } else {
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int kc = it * 4 + csub;
            load_chunk_mapped(frow, kc, IdMap{64}, v);
            st_chunk(A1, chunk_off(r, kc, op_sbo(144)), v);
          }
        }
        for (int kc = csub; kc < 10; kc += 4) {       // [mean|var] in G order → columns 64..143
          load_chunk_mapped(mrow, kc, GMap(), v);
          if (kc == 8) v[kBiasColHi] = v[kBiasColLo] = 1.0f;    // the tile's constant-one columns
          st_chunk(A1, chunk_off(r, 8 + kc, op_sbo(144)), v);
        }
      }
    }
    const int row = tid;
    const uint32_t ones = a1 + 8 * 2 * kLBO;          // last K block of A1: (…, 1, 1 | 0×8)
    if (input_kind == 0) {
      // sigmahead.out_geometry_fc: [128 x 128] · Wgᵀ → 64, ELU → columns 0..63 of A1
      round_sync();
      if (tid == 0)
        issue_gemm_bias(a0, op_sbo(128), wimg + DenImg::Wg, 128, 64, ones, op_sbo(144), tmem, bar_m, kFmtF16);
      wait_round(bar_m, phase);
      epilogue_elu_to_tile<64>(t_row, A1, op_sbo(144), row, 0);
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
          o[kc * 8 + e * 2] = f2.x * (1.0f / kLog2e);          // activations are kept times log2(e)
          o[kc * 8 + e * 2 + 1] = f2.y * (1.0f / kLog2e);
        }
      }
    }
    // out_geometry_fc.0: [128 x 144] · W0ᵀ → 64, ELU → A0 as [128 x 64]
    round_sync();
    if (tid == 0) issue_gemm(a1, op_sbo(144), wimg + DenImg::W0, op_sbo(144), 144, 64, tmem, bar_m);
    wait_round(bar_m, phase);
    epilogue_elu_to_tile<64>(t_row, A0, op_sbo(64), row, 0);
    // .2: [128 x 64] → 32, ELU → A2 as [128 x 32]
    round_sync();
    if (tid == 0) issue_gemm_bias(a0, op_sbo(64), wimg + DenImg::W1, 64, 32, ones, op_sbo(144), tmem, bar_m);
    wait_round(bar_m, phase);
    epilogue_elu_to_tile<32>(t_row, A2, op_sbo(32), row, 0);
    // .4: [128 x 32] → 16, ELU ; .6: 16 → 1 on CUDA cores, ReLU, no-valid-view fill
    round_sync();
    if (tid == 0) issue_gemm_bias(a2, op_sbo(32), wimg + DenImg::W2, 32, 16, ones, op_sbo(144), tmem, bar_m);
    wait_round(bar_m, phase);
    {
      uint32_t r[16];
      tmem_ld16(t_row, r);
      tmem_wait_ld();
      float s = fl[DenImg::b3];
#pragma unroll
      for (int k = 0; k < 16; ++k) s = fmaf(elu_scaled(__uint_as_float(r[k])), fl[DenImg::w3 + k], s);
      s = fmaxf(s, 0.0f);
      if (row < n_valid) {
        float nv = 0.0f;
        for (int v = 0; v < V; ++v) nv += __ldg(mask + (first + row) * V + v);
        sigma[first + row] = (nv < 1.0f) ? 0.0f : s;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------------------
// colour head: 256 threads, thread t ↔ row t % 128, column half t / 128
// ---------------------------------------------------------------------------
template <int V>
struct ColSmem {
  static constexpr uint32_t IMG = 0;
  static constexpr uint32_t G = ((ColImg<V>::BYTES + 127) / 128) * 128;  // [128 x 80] mean|var ; later Y_v / Z
  static constexpr uint32_t H = G + op_bytes(128, 80);                   // [128 x 64] base_fc hidden
  static constexpr uint32_t F = H + op_bytes(128, 64);                   // V x [128 x 48] ; later Xs_v / flat
  static constexpr uint32_t ONES = F + V * op_bytes(128, 48);            // [128 x 16] constant (…, 1, 1 | 0×8): bias block
  static constexpr uint32_t BAR = ONES + op_bytes(128, 16);
  static constexpr uint32_t BYTES = BAR + 64;
  static constexpr uint32_t TMEM_COLS = V == 1 ? 64 : (V == 2 ? 128 : 256);     // V blocks of 64 columns
};

// FROM_REC: rows come from the bf16 records of the fused kernel (rec_chunks(V)
// chunks per point, already in operand order); otherwise from the reference-
// layout fp32 arrays rgb_feat [n][V][35], meanvar [n][70].
template <int V, bool FROM_REC>
__global__ void __launch_bounds__(256, 2) color_mlp_tc(const float* __restrict__ rgb_feat,
                                                       const float* __restrict__ meanvar,
                                                       const uint4* __restrict__ rec,
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
  const int row = tid & 127, half = tid >> 7;

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
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  mbar_wait(bar_w, 0);

  const uint32_t g_a = smem_u32(G), h_a = smem_u32(Hb), f_a = smem_u32(F), wimg = smem_u32(img);
  const uint32_t ones = smem_u32(smem + S::ONES);
  constexpr uint32_t ONES_SBO = op_sbo(16);
  if (tid < 128) {     // written once: made visible to the tensor core by the first round_sync
    uint8_t* o = smem + S::ONES;
    *reinterpret_cast<uint4*>(o + chunk_off(tid, 0, ONES_SBO)) = make_uint4(0u, 0u, 0u, 0x3F803F80u);   // bf16 1.0 in columns 6, 7
    *reinterpret_cast<uint4*>(o + chunk_off(tid, 1, ONES_SBO)) = make_uint4(0u, 0u, 0u, 0u);
  }
  constexpr uint32_t F_STRIDE = op_bytes(128, 48);     // per-view feature operand
  constexpr uint32_t X_STRIDE = op_bytes(128, 32);     // per-view [128 x 32] operand
  constexpr uint32_t HB_STRIDE = op_bytes(128, 64);    // per-view base_fc hidden tile
  static_assert(V * HB_STRIDE <= S::ONES - S::G, "hidden tiles must fit the G|H|F pool");
  constexpr int RC = rec_chunks(V);
  uint32_t phase = 0;
  const int n = count_ptr ? __ldg(count_ptr) : n_const;
  const int n_tiles = (n + 127) / 128;
  const float inv_v = 1.0f / (float)V;
  // source row (record index) of tile row r, -1 past the end
  auto src_of = [&](long long first_row, int r) -> int {
    if (first_row + r >= n) return -1;
    return valid1 ? __ldg(valid1 + first_row + r) : (int)(first_row + r);
  };
  // rows this thread stages: r = (warp + 8 j) * 8 + (lane & 7), j = 0, 1.  Their source indices are
  // fetched one tile ahead and their records are pulled into L2 while the current tile computes.
  int src_cur[2], src_nxt[2] = {-1, -1};
#pragma unroll
  for (int j = 0; j < 2; ++j) src_cur[j] = src_of((long long)blockIdx.x * 128, (warp + 8 * j) * 8 + (lane & 7));
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long first = (long long)tile * 128;
    const int n_valid = min(128, n - (int)first);
    const bool has_next = tile + (int)gridDim.x < n_tiles;
    if (has_next) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
        src_nxt[j] = src_of((long long)(tile + gridDim.x) * 128, (warp + 8 * j) * 8 + (lane & 7));
    }
    // ---- stage G (80 cols) and F_v (48 cols); lanes = 8 rows x 4 chunks
    {
      const int rsub = lane & 7, csub = lane >> 3;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int rg = warp + 8 * j;
        const int r = rg * 8 + rsub;
        const long long src = src_cur[j];     // 64-bit for the row offsets below
        if constexpr (FROM_REC) {
          // RC stored chunks + (1 + V) padding chunks to zero; all loads are issued before the stores
          constexpr int NT = (RC + 1 + V + 3) / 4;
          uint4 q[NT];
#pragma unroll
          for (int it = 0; it < NT; ++it) {
            const int c = csub + 4 * it;
            q[it] = (c < RC && src >= 0) ? __ldg(rec + src * RC + c) : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int it = 0; it < NT; ++it) {
            const int c = csub + 4 * it;
            if (c >= RC + 1 + V) continue;
            uint8_t* dst;
            if (c < 9) {
              dst = G + chunk_off(r, c, op_sbo(80));
            } else if (c < RC) {
              const int vw = (c - 9) / 5, kc = (c - 9) - vw * 5;
              dst = F + vw * F_STRIDE + chunk_off(r, kc, op_sbo(48));
            } else if (c == RC) {
              dst = G + chunk_off(r, 9, op_sbo(80));
            } else {
              dst = F + (c - RC - 1) * F_STRIDE + chunk_off(r, 5, op_sbo(48));
            }
            *reinterpret_cast<uint4*>(dst) = q[it];
          }
        } else {
          const float* mrow = src >= 0 ? meanvar + src * 70 : nullptr;
          const float* frow = src >= 0 ? rgb_feat + src * (V * 35) : nullptr;
          float v[8];
          for (int kc = csub; kc < 10; kc += 4) {
            load_chunk_mapped(mrow, kc, GMap(), v);
            if (kc == 8) v[kBiasColHi] = v[kBiasColLo] = 1.0f;      // constant-one columns 70, 71 (base_fc.0 bias)
            st_chunk(G, chunk_off(r, kc, op_sbo(80)), v);
          }
          for (int t = csub; t < 6 * V; t += 4) {
            const int vw = t / 6, kc = t - vw * 6;
            load_chunk_mapped(frow ? frow + vw * 35 : nullptr, kc, FMap(), v);
            st_chunk(F + vw * F_STRIDE, chunk_off(r, kc, op_sbo(48)), v);
          }
        }
      }
    }
    float x[V][16];                                   // this thread's half of base_fc's output, per view
    // base_fc.0 on [mean|var|feat_v] for ALL views in one round: per view two accumulating GEMMs
    // (K = 80 + 48) → 64, into V blocks of 64 TMEM columns (fewer issue→commit→wait→sync round trips
    // per tile: the kernel is bound by their latency, not by the tensor pipe)
    round_sync();
    if (tid == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64);
#pragma unroll
      for (int vw = 0; vw < V; ++vw) {
        for (int k16 = 0; k16 < 5; ++k16)
          umma_bf16(tmem + vw * 64, make_smem_desc(g_a + k16 * 2 * kLBO, kLBO, op_sbo(80)),
                    make_smem_desc(wimg + I::Wb0a + k16 * 2 * kLBO, kLBO, op_sbo(80)), idesc, k16 > 0);
        for (int k16 = 0; k16 < 3; ++k16)
          umma_bf16(tmem + vw * 64, make_smem_desc(f_a + vw * F_STRIDE + k16 * 2 * kLBO, kLBO, op_sbo(48)),
                    make_smem_desc(wimg + I::Wb0b + k16 * 2 * kLBO, kLBO, op_sbo(48)), idesc, 1u);
      }
      umma_commit(bar_m);
    }
    wait_round(bar_m, phase);
    // hidden tiles Hb_v [128 x 64] over the (now consumed) G | H | F pool
#pragma unroll
    for (int vw = 0; vw < V; ++vw) epi32_to_tile(t_row + vw * 64, half * 32, G + vw * HB_STRIDE, op_sbo(64), row, 0);
    // base_fc.2 for all views: 64 → 32, ELU ; keep x_v in registers
    round_sync();
    if (tid == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 32);
#pragma unroll
      for (int vw = 0; vw < V; ++vw) {
        for (int k16 = 0; k16 < 4; ++k16)
          umma_bf16(tmem + vw * 32, make_smem_desc(g_a + vw * HB_STRIDE + k16 * 2 * kLBO, kLBO, op_sbo(64)),
                    make_smem_desc(wimg + I::Wb1 + k16 * 2 * kLBO, kLBO, op_sbo(80)), idesc, k16 > 0);
        umma_bf16(tmem + vw * 32, make_smem_desc(ones, kLBO, ONES_SBO),
                  make_smem_desc(wimg + I::Wb1 + 4 * 2 * kLBO, kLBO, op_sbo(80)), idesc, 1u);
      }
      umma_commit(bar_m);
    }
    wait_round(bar_m, phase);
#pragma unroll
    for (int vw = 0; vw < V; ++vw) {
      uint32_t r[16];
      tmem_ld16(t_row + vw * 32 + half * 16, r);
      tmem_wait_ld();
#pragma unroll
      for (int k = 0; k < 16; ++k) x[vw][k] = elu_scaled(__uint_as_float(r[k]));
    }
    if constexpr (FROM_REC) {
      if (has_next) {     // pull the next tile's records into L2 (one 128-byte line per lane of the row quad)
        const int csub = lane >> 3;
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (src_nxt[j] >= 0 && csub * 128 < RC * 16) {
            const char* p = reinterpret_cast<const char*>(rec + (long long)src_nxt[j] * RC) + csub * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
          }
      }
    }
    // all F_v consumed → reuse F for Xs_v = x_v / V (vis_fc input, trainhead.py:140)
#pragma unroll
    for (int vw = 0; vw < V; ++vw)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = x[vw][j * 8 + e] * inv_v;
        st_chunk(F + vw * X_STRIDE, chunk_off(row, half * 2 + j, op_sbo(32)), v);
      }
    // vis_fc.0 for all views at once: V GEMMs into V column blocks
    round_sync();
    if (tid == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 32);
      for (int vw = 0; vw < V; ++vw) {
        for (int k16 = 0; k16 < 2; ++k16)
          umma_bf16(tmem + vw * 32, make_smem_desc(f_a + vw * X_STRIDE + k16 * 2 * kLBO, kLBO, op_sbo(32)),
                    make_smem_desc(wimg + I::Wv0 + k16 * 2 * kLBO, kLBO, op_sbo(48)), idesc, k16 > 0);
        umma_bf16(tmem + vw * 32, make_smem_desc(ones, kLBO, ONES_SBO),
                  make_smem_desc(wimg + I::Wv0 + 2 * 2 * kLBO, kLBO, op_sbo(48)), idesc, 1u);
      }
      umma_commit(bar_m);
    }
    wait_round(bar_m, phase);
#pragma unroll
    for (int vw = 0; vw < V; ++vw)   // Y_v → G region (G and H are free now)
      epi16_to_tile(t_row + vw * 32, half * 16, G + vw * X_STRIDE, op_sbo(32), row, 0);
    // vis_fc.2, residual, flatten view-major → flat [128 x 32V] in F
    round_sync();
    if (tid == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 32);
      for (int vw = 0; vw < V; ++vw) {
        for (int k16 = 0; k16 < 2; ++k16)
          umma_bf16(tmem + vw * 32, make_smem_desc(g_a + vw * X_STRIDE + k16 * 2 * kLBO, kLBO, op_sbo(32)),
                    make_smem_desc(wimg + I::Wv1 + k16 * 2 * kLBO, kLBO, op_sbo(48)), idesc, k16 > 0);
        umma_bf16(tmem + vw * 32, make_smem_desc(ones, kLBO, ONES_SBO),
                  make_smem_desc(wimg + I::Wv1 + 2 * 2 * kLBO, kLBO, op_sbo(48)), idesc, 1u);
      }
      umma_commit(bar_m);
    }
    wait_round(bar_m, phase);
#pragma unroll
    for (int vw = 0; vw < V; ++vw) {
      uint32_t r[16];
      tmem_ld16(t_row + vw * 32 + half * 16, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          v[e] = x[vw][j * 8 + e] + elu_scaled(__uint_as_float(r[j * 8 + e]));
        st_chunk(F, chunk_off(row, vw * 4 + half * 2 + j, op_sbo(32 * V)), v);
      }
    }
    // rgb_fc.0: 32V → 32, ELU → Z in G
    round_sync();
    if (tid == 0) issue_gemm_bias(f_a, op_sbo(32 * V), wimg + I::Wr0, 32 * V, 32, ones, ONES_SBO, tmem, bar_m);
    wait_round(bar_m, phase);
    epi16_to_tile(t_row, half * 16, G, op_sbo(32), row, 0);
    // rgb_fc.2: 32 → 16, ELU ; rgb_fc.4: 16 → 3 on CUDA cores ; sigmoid
    round_sync();
    if (tid == 0) issue_gemm_bias(g_a, op_sbo(32), wimg + I::Wr1, 32, 16, ones, ONES_SBO, tmem, bar_m);
    wait_round(bar_m, phase);
    if (half == 0) {
      uint32_t r[16];
      tmem_ld16(t_row, r);
      tmem_wait_ld();
      float o0 = fl[I::rb2], o1 = fl[I::rb2 + 1], o2 = fl[I::rb2 + 2];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float h = elu_scaled(__uint_as_float(r[k]));
        o0 = fmaf(h, fl[I::rw2 + k], o0);
        o1 = fmaf(h, fl[I::rw2 + 16 + k], o1);
        o2 = fmaf(h, fl[I::rw2 + 32 + k], o2);
      }
      if (row < n_valid) {
        const long long dst = valid1 ? (long long)__ldg(valid1 + first + row) : first + row;
        rgb[dst * 3 + 0] = sigmoid_fast(o0);
        rgb[dst * 3 + 1] = sigmoid_fast(o1);
        rgb[dst * 3 + 2] = sigmoid_fast(o2);
      }
    }
    src_cur[0] = src_nxt[0];
    src_cur[1] = src_nxt[1];
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

template <int V, bool FROM_REC>
static int launch_color_tc(const float* rgb_feat, const float* meanvar, const void* rec, const int32_t* valid1,
                           const uint8_t* image, int n_points_max, const int32_t* count_ptr, float* rgb,
                           cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(color_mlp_tc<V, FROM_REC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ColSmem<V>::BYTES);
    if (e != cudaSuccess) {
      set_error("color_mlp_tc smem attribute", e);
      return GPNERF_E_CUDA;
    }
    attr_set = true;
  }
  int tiles = (n_points_max + 127) / 128;
  static const int per_sm = getenv("GPNERF_COLOR_CTAS_PER_SM") ? atoi(getenv("GPNERF_COLOR_CTAS_PER_SM")) : 2;   // experiment knob
  int grid = tiles < per_sm * sm_count() ? tiles : per_sm * sm_count();
  color_mlp_tc<V, FROM_REC><<<grid, 256, ColSmem<V>::BYTES, st>>>(rgb_feat, meanvar,
                                                                  reinterpret_cast<const uint4*>(rec), valid1,
                                                                  image + kColImgOffset, count_ptr, n_points_max,
                                                                  rgb);
  return check_launch("k3_color_mlp (tcgen05)");
}

// rec != NULL selects the record-fed variant
int gpnerf_color_mlp_tc_any(const float* rgb_feat, const float* meanvar, const void* rec, const int32_t* valid1,
                            const gpnerf_head_weights_t* w, int n_views, int n_points_max,
                            const int32_t* count_ptr, float* rgb, cudaStream_t st) {
  if (w->tc_image == nullptr) {
    set_error("weights->tc_image is NULL: call gpnerf_k3_pack_weights first", cudaSuccess);
    return GPNERF_E_ARG;
  }
  const uint8_t* image = reinterpret_cast<const uint8_t*>(w->tc_image);
#define GPNERF_COLOR_CASE(VV)                                                                                  \
  case VV:                                                                                                     \
    return rec ? launch_color_tc<VV, true>(rgb_feat, meanvar, rec, valid1, image, n_points_max, count_ptr, rgb, st) \
               : launch_color_tc<VV, false>(rgb_feat, meanvar, rec, valid1, image, n_points_max, count_ptr, rgb, st);
  switch (n_views) {
    GPNERF_COLOR_CASE(1) GPNERF_COLOR_CASE(2) GPNERF_COLOR_CASE(3) GPNERF_COLOR_CASE(4)
    default:
      set_error("tcgen05 colour head supports 1..4 source views", cudaSuccess);
      return GPNERF_E_UNSUPPORTED;
  }
#undef GPNERF_COLOR_CASE
}

int gpnerf_color_mlp_tc(const float* rgb_feat, const float* meanvar, const int32_t* valid1,
                        const gpnerf_head_weights_t* w, int n_views, int n_points_max, const int32_t* count_ptr,
                        float* rgb, cudaStream_t st) {
  return gpnerf_color_mlp_tc_any(rgb_feat, meanvar, nullptr, valid1, w, n_views, n_points_max, count_ptr, rgb, st);
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
