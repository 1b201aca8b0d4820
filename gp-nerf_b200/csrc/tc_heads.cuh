// Shared definitions of the tcgen05 head kernels: packed-weight image layout,
// operand column orders, epilogue helpers.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace gpnerf {
using namespace tc;

__host__ __device__ constexpr uint32_t op_sbo(int Kp) { return (uint32_t)(Kp / 8) * kLBO; }
__host__ __device__ constexpr uint32_t op_bytes(int rows, int Kp) { return (uint32_t)(rows / 8) * op_sbo(Kp); }

// ---------------------------------------------------------------------------
// Operand column orders.  The gather kernels own 8 consecutive feature channels
// per lane, i.e. exactly one 16-byte bf16 chunk; so the 32 feature channels
// come first and the 3 RGB channels after them (the reference order is RGB
// first, BaseRender.py:358).  Weights are packed with the same permutation, so
// the products are unchanged.
//   G tile (80 cols): mean_feat 32 | var_feat 32 | mean_rgb 3, var_rgb 3, 1, 1 | 0×8
//   F tile (48 cols): feat 32 | rgb 3, 0×5 | 0×8
// Columns 70, 71 of the G tile are the constant 1.0: the layers that read the G
// tile carry their bias (bf16 hi + lo) in those two columns of the weight
// operand, and the tile's last K block (columns 64..79) doubles as the "ones"
// A block of issue_gemm_bias for the other layers of the density head.
// ---------------------------------------------------------------------------
constexpr int kBiasColHi = 6, kBiasColLo = 7;   // position of the ones inside a bias K block
// column j of the G tile → index into the reference's [mean 35 | var 35] row (or -1)
__host__ __device__ inline int gmap(int j) {
  if (j < 32) return 3 + j;
  if (j < 64) return 35 + 3 + (j - 32);
  if (j < 67) return j - 64;
  if (j < 70) return 35 + (j - 67);
  return -1;
}
// column j of an F tile → index into the reference's [rgb 3 | feat 32] row (or -1)
__host__ __device__ inline int fmap(int j) {
  if (j < 32) return 3 + j;
  if (j < 35) return j - 32;
  return -1;
}

// ---------------------------------------------------------------------------
// packed weight images (bf16 UMMA operands + fp32 biases / last layers)
// ---------------------------------------------------------------------------
// Scaling (tc_common.cuh, elu_scaled): operands that multiply raw gathered features are packed
// times c = log2(e), every bias times c, the CUDA-core last layers divided by c.
struct DenImg {   // density head
  static constexpr uint32_t Wg = 0;                                 // [64 x 144] = c·Wg as FP16 (128) | bias block
  static constexpr uint32_t W0 = Wg + op_bytes(64, 144);            // [64 x 144] = sigma_feat 64 | G order 80 (bias in 70,71)
  static constexpr uint32_t W1 = W0 + op_bytes(64, 144);            // [32 x 80]  = 64 | bias block
  static constexpr uint32_t W2 = W1 + op_bytes(32, 80);             // [16 x 48]  = 32 | bias block
  static constexpr uint32_t F32 = W2 + op_bytes(16, 48);            // floats below
  static constexpr int w3 = 0, b3 = 16, NF = 20;                    // w3 / c, b3
  static constexpr uint32_t BYTES = F32 + NF * 4;
};
static_assert(DenImg::BYTES % 16 == 0, "bulk copy needs 16-byte multiples");

template <int V>
struct ColImg {   // colour head; base_fc.0 split into its [mean|var] (G order) and per-view (F order) blocks
  static constexpr uint32_t Wb0a = 0;                               // [64 x 80]  c·W, bias in columns 70, 71
  static constexpr uint32_t Wb0b = Wb0a + op_bytes(64, 80);         // [64 x 48]  c·W
  static constexpr uint32_t Wb1 = Wb0b + op_bytes(64, 48);          // [32 x 80]  = 64 | bias block
  static constexpr uint32_t Wv0 = Wb1 + op_bytes(32, 80);           // [32 x 48]
  static constexpr uint32_t Wv1 = Wv0 + op_bytes(32, 48);           // [32 x 48]
  static constexpr uint32_t Wr0 = Wv1 + op_bytes(32, 48);           // [32 x (32V+16)]
  static constexpr uint32_t Wr1 = Wr0 + op_bytes(32, 32 * V + 16);  // [16 x 48]
  // warp-specialised colour head (k3_color_ws.cu): the per-view RGB inputs of base_fc.0 sit in ONE shared
  // [128 x 16] operand (column 3v + c = channel c of view v), so every view has its own [64 x 16] weight block that
  // picks its three columns; and rgb_fc.0's residual input x_v is fed as x_v / V (the vis_fc input that is in TMEM
  // anyway) against weights scaled by V
  static constexpr uint32_t Wb0r = Wr1 + op_bytes(16, 48);          // V x [64 x 16]  c·W[:, 70 + c] in columns 3v + c
  static constexpr uint32_t Wr0x = Wb0r + V * op_bytes(64, 16);     // [32 x 32V]    V · rgb_fc.0.weight
  static constexpr uint32_t F32 = Wr0x + op_bytes(32, 32 * V);
  static constexpr int rw2 = 0, rb2 = 48;                           // rgb_fc.4 weights / c, bias
  // CUDA-core tail of the warp-specialised colour head: c·rgb_fc.0.bias (added in the epilogue), rgb_fc.2 as fp32
  // weights / c [16][32] + bias, rgb_fc.4 weights unscaled [3][16]
  static constexpr int rb0c = 52, w1f = 84, b1f = 596, rw2u = 612, NF = 660;
  static constexpr uint32_t BYTES = F32 + NF * 4;
};

constexpr uint32_t kColImgOffset = ((DenImg::BYTES + 127) / 128) * 128;
constexpr uint32_t kImageBytes = kColImgOffset + ((ColImg<4>::BYTES + 127) / 128) * 128;

// Colour-stage record written by the fused gather+density kernel, one per P1
// point: the G tile's 9 non-zero chunks followed by 5 non-zero chunks per view.
__host__ __device__ constexpr int rec_chunks(int V) { return 9 + 5 * V; }

// Colour-stage tile record (round 2, tile hand-off): what the fused gather → density kernel leaves behind for
// the colour head, one block per 128-point tile of the P1 list, laid out exactly as the colour head's shared-
// memory stage (tcgen05 operand layouts), so that the colour head fetches a tile with ONE bulk copy
// (cp.async.bulk) and feeds it to the tensor core untouched:
//   G64   [128 x 64] bf16  mean_feat | var_feat                  K-major SWIZZLE_128B
//   FF    views 2j, 2j+1 share a [128 x 64] SWIZZLE_128B block (view v at columns 32 (v % 2) …); an odd last view
//         has a [128 x 32] block in the unswizzled core-matrix layout (8 KB)
//   TAIL  [128 x 16] bf16  mean rgb 3, var rgb 3, 1, 1 | 0 x 8   core-matrix layout
//   RGBS  [128 x 16] bf16  column 3v + c = channel c of view v   core-matrix layout
// Columns that no view owns are never written: the buffer must be zero-filled once by whoever allocates it.
template <int V>
struct RecTile {
  static constexpr uint32_t G64 = 0;
  static constexpr uint32_t FF = 16384;
  static constexpr uint32_t FF_ODD = FF + (V / 2) * 16384;            // the odd last view's block
  static constexpr uint32_t TAIL = FF_ODD + (V & 1) * 8192;
  static constexpr uint32_t RGBS = TAIL + 4096;
  static constexpr uint32_t BYTES = RGBS + 4096;                      // V = 1..4: 32, 40, 48, 56 KB
  static constexpr uint32_t kOddSbo = op_sbo(32);
  // byte offset (inside the tile) of 16-byte chunk c (0..3) of view v's 32 features of row r
  __device__ __forceinline__ static uint32_t ff_off(int v, int r, int c) {
    if ((V & 1) && v == V - 1) return FF_ODD + chunk_off(r, c, kOddSbo);
    return FF + (uint32_t)(v >> 1) * 16384u + sw128_off(r, (v & 1) * 4 + c);
  }
};
__host__ __device__ constexpr uint32_t rec_tile_bytes(int V) { return 16384u + (V / 2) * 16384u + (V & 1) * 8192u + 8192u; }

struct FusedArgs {
  const __half* lv[GPNERF_N_LEVELS];
  const __half* feat;            // [V][fh+2][fw+2][32]
  const float4* rgbx;            // [V][H+2][W+2] (r,g,b,·) in [0,1]
  const int32_t* valid;
  const float *rays_o, *rays_d, *z_vals;
  const int32_t* counters;
  const uint8_t* image;          // packed weights
  float* sigma;
  uint4* rec;
  float* alpha;              // optional: K4's α = 1 − exp(−σ) …
  uint32_t* alpha_words;     // … and its survivor flags (one ballot word per 32 points), written here
  uint8_t* rec_tiles;        // optional: colour-stage tile records (RecTile<V>), one per 128 points
  float* rgb_in;             // optional [P1][V][3]: the per-view RGB taps (BaseRender's rgb_in_map input)
  int debug;                 // profiling experiments only (GPNERF_FUSED_DEBUG): 1 = producers skip the gathers
};

// ---------------------------------------------------------------------------
// epilogue helpers
// ---------------------------------------------------------------------------
// accumulator columns [c0, c0+32) of this thread's row → scaled ELU → 4 bf16 chunks
__device__ __forceinline__ void epi32_to_tile(uint32_t taddr, int c0, uint8_t* tile, uint32_t sbo, int row, int kc0) {
  uint32_t r[32];
  tmem_ld32(taddr + c0, r);
  tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = elu_scaled(__uint_as_float(r[j * 8 + e]));
    st_chunk(tile, chunk_off(row, kc0 + (c0 >> 3) + j, sbo), v);
  }
}
__device__ __forceinline__ void epi16_to_tile(uint32_t taddr, int c0, uint8_t* tile, uint32_t sbo, int row, int kc0) {
  uint32_t r[16];
  tmem_ld16(taddr + c0, r);
  tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = elu_scaled(__uint_as_float(r[j * 8 + e]));
    st_chunk(tile, chunk_off(row, kc0 + (c0 >> 3) + j, sbo), v);
  }
}
template <int N>
__device__ __forceinline__ void epilogue_elu_to_tile(uint32_t taddr, uint8_t* tile, uint32_t sbo, int row, int kc0) {
  static_assert(N % 32 == 0, "");
#pragma unroll
  for (int c = 0; c < N / 32; ++c) epi32_to_tile(taddr, c * 32, tile, sbo, row, kc0);
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

// 8 fp32 of a reference-layout row picked through a column map → one chunk
template <class Map>
__device__ __forceinline__ void load_chunk_mapped(const float* __restrict__ row, int kc, Map map, float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int src = map(kc * 8 + j);
    v[j] = (row != nullptr && src >= 0) ? __ldg(row + src) : 0.0f;
  }
}
struct GMap {
  __device__ __forceinline__ int operator()(int j) const { return gmap(j); }
};
struct FMap {
  __device__ __forceinline__ int operator()(int j) const { return fmap(j); }
};
struct IdMap {
  int kmax;
  __device__ __forceinline__ int operator()(int j) const { return j < kmax ? j : -1; }
};

}  // namespace gpnerf
