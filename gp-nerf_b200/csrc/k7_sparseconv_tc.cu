// K7 on tensor cores: the sparse convolution's gather-GEMM as tcgen05.mma.kind::tf32 with a three-term split
// (SparseConvNet.py:21-124; same arguments and result as sc_conv in k7_sparseconv.cu, which stays as the
// plain-fp32 parity kernel).
//
//   out[o] = relu(scale ⊙ Σ_k W[k]ᵀ · in[nbr_k(o)] + shift)          per output site o, 27 taps k
//
// fp32 accuracy from TF32 operands: x = x_hi + x_lo with x_hi = x with its 13 low mantissa bits cleared (exactly
// representable in TF32) and x_lo = x − x_hi (exact in fp32; the tensor core keeps its 11 leading bits), and
//     x·w ≈ x_hi·w_hi + x_lo·w_hi + x_hi·w_lo            (dropped: x_lo·w_lo ≈ 2⁻²²·|x·w|)
// accumulated in fp32 in TMEM – three MMAs per K-step instead of one.  The weights are split and laid out as UMMA
// B operands once on the host side (sparseconv.SparseConvNet._fold); the gathered rows are split on their way
// from registers to shared memory.
//
// One CTA = 256 threads = a tile of 128 output sites; a cluster of 3 CTAs splits the 27 taps (9 each: kd = rank)
// and reduces the three partial tiles over DSMEM in a fixed order, as the fp32 kernel does.  Per active tap:
// the 128 neighbour rows (16-byte loads issued one tap ahead, zero rows where there is no neighbour) and the
// tap's two weight images go to one of two shared-memory stages; thread 0 issues 3·CIN/8 MMAs
// (M=128, N=COUT, K=8) and commits to the stage's mbarrier, which gates the stage's reuse.  Taps no site of
// the tile has are skipped.
//
// Operand layout (32-bit elements, K-major, no swizzle; core matrix = 8 rows × 16 bytes):
//     byte(r, k) = (r/8)·SBO + (k/4)·128 + (r%8)·16 + (k%4)·4,   SBO = (CIN/4)·128.
// Lane mapping of the staging: lane%8 = row within its group of 8, lane/8 = 16-byte chunk, so a quarter-warp
// writes 128 contiguous bytes (no bank conflicts) and reads 8 rows × one chunk from global memory.
#include <cooperative_groups.h>

#include "tc_common.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace gpnerf {
using namespace tc;

namespace {

constexpr uint32_t kFmtTF32x = 2;
constexpr int kTile = 128, kThreads = 256, kSplit = 3, kTaps = 27 / kSplit, kStages = 2;

__device__ __forceinline__ void umma_tf32x(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int CIN, int COUT>
struct ScTc {
  static constexpr int CH = CIN / 4;                              // 16-byte chunks per row
  static constexpr uint32_t SBO = (uint32_t)CH * kLBO;            // bytes between 8-row groups (A and B alike)
  static constexpr uint32_t A_BYTES = kTile / 8 * SBO;            // one [128 × CIN] fp32 operand
  static constexpr uint32_t B_BYTES = COUT / 8 * SBO;             // one [COUT × CIN] fp32 operand
  static constexpr uint32_t STAGE = 2 * A_BYTES + 2 * B_BYTES;    // A_hi | A_lo | B_hi | B_lo
  static constexpr int A_ITEMS = kTile * CH / kThreads;           // float4 per thread and tap (4 | 2)
  static constexpr int B_F4 = 2 * (int)B_BYTES / 16;              // float4 in a tap's packed weight image
  static constexpr size_t SMEM = (size_t)kStages * STAGE + kTaps * kTile * 4 + 256;
};

template <int CIN, int COUT>
__global__ void __cluster_dims__(kSplit, 1, 1) __launch_bounds__(kThreads, 2)
    sc_conv_tc(const float* __restrict__ in_feat, const int32_t* __restrict__ nbr, int n_out_max,
               const int32_t* __restrict__ n_out_dev, const float* __restrict__ w_packed,
               const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out_feat) {
  using L = ScTc<CIN, COUT>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* stages = smem;
  int* ids = reinterpret_cast<int*>(smem + kStages * L::STAGE);                      // [kTaps][kTile]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ids + kTaps * kTile);                 // [kStages] stage free, [kStages] = all done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kStages + 1);
  int* tap_any = reinterpret_cast<int*>(tmem_slot + 1);                              // [kTaps]
  int* taps = tap_any + kTaps;                                                       // [kTaps]
  int* n_taps_s = taps + kTaps;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s <= kStages; ++s) mbar_init(bars + s, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t idesc = make_idesc(128, COUT, kFmtTF32x);
  const int n_out = __ldg(n_out_dev);
  const int n_tiles = (n_out + kTile - 1) / kTile;
  uint32_t stage_uses[kStages] = {0, 0};       // how often each stage has been committed so far (all threads agree)
  uint32_t done_phase = 0;
  float* park = reinterpret_cast<float*>(stages);                                    // [kTile][COUT] partial sums

  // staging coordinates of this thread: item i covers row rg*8 + lane%8, chunk c
  int st_row[L::A_ITEMS];
  uint32_t st_off[L::A_ITEMS];
#pragma unroll
  for (int i = 0; i < L::A_ITEMS; ++i) {
    const int blk = (i * kThreads + tid) >> 5;                   // 32-item block
    const int rg = L::CH == 8 ? blk >> 1 : blk, c = (L::CH == 8 ? (blk & 1) * 4 : 0) + (lane >> 3);
    st_row[i] = rg * 8 + (lane & 7);
    st_off[i] = (uint32_t)rg * L::SBO + (uint32_t)c * kLBO + (uint32_t)(lane & 7) * 16;
    st_row[i] |= c << 16;                                         // chunk in the high half
  }

  for (int tile = blockIdx.x / kSplit; tile < n_tiles; tile += gridDim.x / kSplit) {
    const int o0 = tile * kTile;
    if (tid < kTaps) tap_any[tid] = 0;
    __syncthreads();
    for (int t = tid; t < kTaps * kTile; t += kThreads) {
      const int kk = t / kTile, r = t - kk * kTile;
      const int row = o0 + r < n_out ? __ldg(nbr + (size_t)(crank * kTaps + kk) * n_out_max + o0 + r) : -1;
      ids[t] = row;
      if (row >= 0) tap_any[kk] = 1;                   // benign race: every writer stores 1
    }
    __syncthreads();
    if (tid == 0) {
      int n = 0;
      for (int kk = 0; kk < kTaps; ++kk)
        if (tap_any[kk]) taps[n++] = kk;
      *n_taps_s = n;
    }
    __syncthreads();
    const int n_taps = *n_taps_s;

    float4 xa[L::A_ITEMS];
    auto load_tap = [&](int j) {                       // global → registers (rows of tap taps[j])
      const int kk = taps[j];
#pragma unroll
      for (int i = 0; i < L::A_ITEMS; ++i) {
        const int row = ids[kk * kTile + (st_row[i] & 0xffff)];
        xa[i] = row >= 0 ? __ldg(reinterpret_cast<const float4*>(in_feat + (size_t)row * CIN) + (st_row[i] >> 16))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (n_taps > 0) load_tap(0);
    for (int j = 0; j < n_taps; ++j) {
      const int s = j % kStages;
      uint8_t* sb = stages + (size_t)s * L::STAGE;
      if (stage_uses[s] > 0) mbar_wait(bars + s, (stage_uses[s] - 1) & 1u);        // MMAs that read the stage are done
      tc_fence_after();
      // registers → A_hi | A_lo
#pragma unroll
      for (int i = 0; i < L::A_ITEMS; ++i) {
        const float v[4] = {xa[i].x, xa[i].y, xa[i].z, xa[i].w};
        float hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          hi[q] = __uint_as_float(__float_as_uint(v[q]) & 0xffffe000u);
          lo[q] = v[q] - hi[q];
        }
        *reinterpret_cast<float4*>(sb + st_off[i]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(sb + L::A_BYTES + st_off[i]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      // the tap's weight images (already in operand layout): B_hi | B_lo
      {
        const float4* wsrc = reinterpret_cast<const float4*>(w_packed) + (size_t)(crank * kTaps + taps[j]) * L::B_F4;
        float4* wdst = reinterpret_cast<float4*>(sb + 2 * L::A_BYTES);
        for (int t = tid; t < L::B_F4; t += kThreads) wdst[t] = __ldg(wsrc + t);
      }
      if (j + 1 < n_taps) load_tap(j + 1);             // in flight across the barrier and the MMA issue
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) {
        const uint32_t a_hi = smem_u32(sb), a_lo = a_hi + L::A_BYTES;
        const uint32_t b_hi = a_hi + 2 * L::A_BYTES, b_lo = b_hi + L::B_BYTES;
#pragma unroll
        for (int k8 = 0; k8 < CIN / 8; ++k8) {
          const uint32_t ko = (uint32_t)k8 * 2 * kLBO;
          const uint64_t ah = make_smem_desc(a_hi + ko, kLBO, L::SBO), al = make_smem_desc(a_lo + ko, kLBO, L::SBO);
          const uint64_t bh = make_smem_desc(b_hi + ko, kLBO, L::SBO), bl = make_smem_desc(b_lo + ko, kLBO, L::SBO);
          umma_tf32x(tmem, ah, bh, idesc, (j > 0 || k8 > 0) ? 1u : 0u);
          umma_tf32x(tmem, al, bh, idesc, 1u);
          umma_tf32x(tmem, ah, bl, idesc, 1u);
        }
        umma_commit(bars + s);
        if (j == n_taps - 1) umma_commit(bars + kStages);          // everything of this tile
      }
      stage_uses[s] += 1;
    }
    // ---- epilogue: thread (row, half) owns 16 columns of its row
    const int row = tid & 127, half = tid >> 7;
    const bool has_cols = half * 16 < COUT;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.0f;
    if (n_taps > 0) {
      mbar_wait(bars + kStages, done_phase);
      done_phase ^= 1u;
      tc_fence_after();
      if (has_cols) {
        uint32_t r[16];
        tmem_ld16(t_row + half * 16, r);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = __uint_as_float(r[c]);
      }
    }
    tc_fence_before();
    __syncthreads();                                   // all MMAs done, all TMEM reads done: stages are free
    if (crank != 0 && has_cols) {
#pragma unroll
      for (int c = 0; c < 16; c += 4)
        *reinterpret_cast<float4*>(park + row * COUT + half * 16 + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
    }
    cluster.sync();                                    // partial sums of CTAs 1, 2 are visible
    if (crank == 0 && has_cols) {
      const int o = o0 + row;
      if (o < n_out) {
        const float* p1 = cluster.map_shared_rank(park, 1) + row * COUT + half * 16;
        const float* p2 = cluster.map_shared_rank(park, 2) + row * COUT + half * 16;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float4 b1 = *reinterpret_cast<const float4*>(p1 + c), b2 = *reinterpret_cast<const float4*>(p2 + c);
          const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + half * 16 + c));
          const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + half * 16 + c));
          float4 y;
          y.x = fmaxf(fmaf((acc[c] + b1.x) + b2.x, sc.x, sh.x), 0.0f);
          y.y = fmaxf(fmaf((acc[c + 1] + b1.y) + b2.y, sc.y, sh.y), 0.0f);
          y.z = fmaxf(fmaf((acc[c + 2] + b1.z) + b2.z, sc.z, sh.z), 0.0f);
          y.w = fmaxf(fmaf((acc[c + 3] + b1.w) + b2.w, sc.w, sh.w), 0.0f);
          *reinterpret_cast<float4*>(out_feat + (size_t)o * COUT + half * 16 + c) = y;
        }
      }
    }
    cluster.sync();                                    // parked sums were read: the stages may be refilled
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

template <int CIN, int COUT>
int launch(const float* in_feat, const int32_t* nbr, const int32_t* n_out_dev, int n_out_max, const float* w_packed,
           const float* scale, const float* shift, float* out_feat, cudaStream_t st) {
  using L = ScTc<CIN, COUT>;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(sc_conv_tc<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    if (e != cudaSuccess) {
      set_error("sc_conv_tc smem attribute", e);
      return GPNERF_E_CUDA;
    }
    set = true;
  }
  long long tiles = ((long long)n_out_max + kTile - 1) / kTile;
  const long long cap = (long long)sm_count() * 2;
  if (tiles > cap) tiles = cap;
  if (tiles < 1) tiles = 1;
  sc_conv_tc<CIN, COUT><<<(int)tiles * kSplit, kThreads, L::SMEM, st>>>(in_feat, nbr, n_out_max, n_out_dev, w_packed,
                                                                        scale, shift, out_feat);
  return check_launch("sc_conv_tc");
}

}  // namespace
}  // namespace gpnerf

using namespace gpnerf;

extern "C" {

int gpnerf_sc_conv_tc(const float* in_feat, int c_in, const int32_t* nbr, const int32_t* n_out_dev, int n_out_max,
                      const float* w_packed, const float* scale, const float* shift, int c_out, float* out_feat,
                      void* stream) {
  GPNERF_REQUIRE(in_feat && nbr && n_out_dev && w_packed && scale && shift && out_feat && n_out_max > 0);
  cudaStream_t st = (cudaStream_t)stream;
#define GPNERF_SC(CI, CO) \
  if (c_in == CI && c_out == CO) return launch<CI, CO>(in_feat, nbr, n_out_dev, n_out_max, w_packed, scale, shift, out_feat, st);
  GPNERF_SC(16, 16) GPNERF_SC(16, 32) GPNERF_SC(32, 32) GPNERF_SC(32, 16)
#undef GPNERF_SC
  set_error("sc_conv_tc supports channel widths 16 and 32", cudaSuccess);
  return GPNERF_E_UNSUPPORTED;
}

}  // extern "C"
