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
// Between the layers the features travel already split: a row is [x_hi (C) | x_lo (C)], written by the producing
// layer's epilogue, so the gather is a plain 16-byte asynchronous copy (cp.async, zero-fill where a site has no
// neighbour) straight into the two A operands – no register staging, two taps in flight beyond the current one.
//
// One CTA = 256 threads = a tile of 128 output sites, all 27 taps (those no site of the tile has are skipped).
// Per active tap the 128 neighbour rows and the tap's two weight images land in one of four shared-memory
// stages; thread 0 issues 3·CIN/8 MMAs (M=128, N=COUT, K=8) and commits to the stage's mbarrier, which gates the
// stage's refill.  Epilogue: TMEM → scale/shift → ReLU → split → rows of the next layer (+ plain fp32 rows for
// the renderer where the layer closes a pyramid level).
//
// Operand layout (32-bit elements, K-major, no swizzle; core matrix = 8 rows × 16 bytes):
//     byte(r, k) = (r/8)·SBO + (k/4)·128 + (r%8)·16 + (k%4)·4,   SBO = (CIN/4)·128.
// Lane mapping of the staging: lane%8 = row within its group of 8, lane/8 = 16-byte chunk, so a quarter-warp
// writes 128 contiguous bytes (no bank conflicts) and reads 8 rows × one chunk from global memory.
#include "tc_common.cuh"
#include "common.cuh"

namespace gpnerf {
using namespace tc;

namespace {

constexpr uint32_t kFmtTF32x = 2;
constexpr int kTile = 128, kThreads = 256, kNTaps = 27, kStages = 4, kAhead = kStages - 2;   // taps in flight beyond the current one

__device__ __forceinline__ void umma_tf32x(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_addr, const void* gmem, bool valid) {
  const int n = valid ? 16 : 0;                        // src-size 0: 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

template <int CIN, int COUT>
struct ScTc {
  static constexpr int CH = CIN / 4;                              // 16-byte chunks per operand row
  static constexpr uint32_t SBO = (uint32_t)CH * kLBO;            // bytes between 8-row groups (A and B alike)
  static constexpr uint32_t A_BYTES = kTile / 8 * SBO;            // one [128 × CIN] fp32 operand
  static constexpr uint32_t B_BYTES = COUT / 8 * SBO;             // one [COUT × CIN] fp32 operand
  static constexpr uint32_t STAGE = 2 * A_BYTES + 2 * B_BYTES;    // A_hi | A_lo | B_hi | B_lo
  static constexpr int A_ITEMS = 2 * kTile * CH / kThreads;       // 16-byte copies per thread and tap (8 | 4)
  static constexpr int B_F4 = 2 * (int)B_BYTES / 16;              // float4 in a tap's packed weight image
  static constexpr size_t SMEM = (size_t)kStages * STAGE + kNTaps * kTile * 4 + 512;
};

// in_split / out_split: rows of 2·C floats, [TF32-exact part | remainder] (written by the producing layer's
// epilogue, so the gather is a plain asynchronous copy into the two A operands)
template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1)
    sc_conv_tc(const float* __restrict__ in_split, const int32_t* __restrict__ nbr, int n_out_max,
               const int32_t* __restrict__ n_out_dev, const float* __restrict__ w_packed,
               const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out_split,
               float* __restrict__ out_full) {
  using L = ScTc<CIN, COUT>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* stages = smem;
  int* ids = reinterpret_cast<int*>(smem + kStages * L::STAGE);                      // [27][kTile]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ids + kNTaps * kTile);                // [kStages] stage free, [kStages] = tile done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kStages + 1);
  int* tap_any = reinterpret_cast<int*>(tmem_slot + 1);                              // [27]
  int* taps = tap_any + kNTaps;                                                      // [27]
  int* n_taps_s = taps + kNTaps;
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
  const uint32_t stages_u32 = smem_u32(stages);
  const uint64_t d_ah = make_smem_desc(stages_u32, kLBO, L::SBO), d_al = make_smem_desc(stages_u32 + L::A_BYTES, kLBO, L::SBO);
  const uint64_t d_bh = make_smem_desc(stages_u32 + 2 * L::A_BYTES, kLBO, L::SBO);
  const uint64_t d_bl = make_smem_desc(stages_u32 + 2 * L::A_BYTES + L::B_BYTES, kLBO, L::SBO);
  const int n_out = __ldg(n_out_dev);
  const int n_tiles = (n_out + kTile - 1) / kTile;
  uint32_t commits = 0;                          // MMAs committed so far by this CTA (stage = commit index % kStages)
  uint32_t done_phase = 0;

  // this thread's copies of a tap: item i → (part hi|lo, row, chunk); a quarter-warp fills 128 contiguous bytes
  int it_row[L::A_ITEMS];
  uint32_t it_dst[L::A_ITEMS], it_src[L::A_ITEMS];
#pragma unroll
  for (int i = 0; i < L::A_ITEMS; ++i) {
    const int blk = (i * kThreads + tid) >> 5;                         // 32-copy block: 8 rows × 4 chunks
    constexpr int BLK_PER_PART = kTile / 8 * (L::CH / 4);
    const int part = blk / BLK_PER_PART, b2 = blk - part * BLK_PER_PART;
    const int rg = b2 / (L::CH / 4), c = (b2 % (L::CH / 4)) * 4 + (lane >> 3);
    it_row[i] = rg * 8 + (lane & 7);
    it_dst[i] = (uint32_t)part * L::A_BYTES + (uint32_t)rg * L::SBO + (uint32_t)c * kLBO + (uint32_t)(lane & 7) * 16;
    it_src[i] = (uint32_t)(part * CIN + c * 4);                        // float offset inside the 2·CIN row
  }

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int o0 = tile * kTile;
    if (tid < kNTaps) tap_any[tid] = 0;
    __syncthreads();
    for (int t = tid; t < kNTaps * kTile; t += kThreads) {
      const int kk = t / kTile, r = t - kk * kTile;
      const int row = o0 + r < n_out ? __ldg(nbr + (size_t)kk * n_out_max + o0 + r) : -1;
      ids[t] = row;
      if (row >= 0) tap_any[kk] = 1;                   // benign race: every writer stores 1
    }
    __syncthreads();
    if (tid == 0) {
      int n = 0;
      for (int kk = 0; kk < kNTaps; ++kk)
        if (tap_any[kk]) taps[n++] = kk;
      *n_taps_s = n;
    }
    __syncthreads();
    const int n_taps = *n_taps_s;

    auto issue = [&](int j, uint32_t stage) {          // tap taps[j] → shared-memory stage (asynchronous copies)
      if (j < n_taps) {
        const int kk = taps[j];
        const uint32_t sb = stages_u32 + stage * L::STAGE;
#pragma unroll
        for (int i = 0; i < L::A_ITEMS; ++i) {
          const int row = ids[kk * kTile + it_row[i]];
          cp_async16_zfill(sb + it_dst[i], in_split + (size_t)(row >= 0 ? row : 0) * (2 * CIN) + it_src[i], row >= 0);
        }
        const float4* wsrc = reinterpret_cast<const float4*>(w_packed) + (size_t)kk * L::B_F4;
        for (int t = tid; t < L::B_F4; t += kThreads) cp_async16_zfill(sb + 2 * L::A_BYTES + (uint32_t)t * 16, wsrc + t, true);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");          // one group per call, empty past the last tap
    };
    // stage of tap j = (commits_at_tile_start + j) % kStages
    const uint32_t c0 = commits;
    // A stage is refilled two taps after it was multiplied (kStages = kAhead + 2): the wait below is for a commit
    // that is one whole iteration old, so the MMA → mbarrier → wake-up latency stays off the tap chain.  Stages
    // touched before the first wait held taps of the previous tile, all older than its tile-done barrier.
    for (int j = 0; j < kAhead; ++j) issue(j, (c0 + j) % kStages);
    for (int j = 0; j < n_taps; ++j) {
      const uint32_t s = (c0 + j) % kStages;
      if (j >= 2) {
        const uint32_t g = c0 + j - 2;                               // global index of the commit to wait for
        mbar_wait(bars + g % kStages, (g / kStages) & 1u);
        tc_fence_after();
      }
      issue(j + kAhead, (c0 + j + kAhead) % kStages);
      asm volatile("cp.async.wait_group %0;" ::"n"(kAhead) : "memory");
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) {
        // descriptors differ from the stage-0, k=0 ones only in the address field (bits 0-13, 16-byte units): the
        // issuing thread is on the critical path between two barriers, so it adds instead of rebuilding them
        const uint64_t so = (uint64_t)((s * L::STAGE) >> 4);
#pragma unroll
        for (int k8 = 0; k8 < CIN / 8; ++k8) {
          const uint64_t o = so + (uint64_t)((k8 * 2 * kLBO) >> 4);
          umma_tf32x(tmem, d_ah + o, d_bh + o, idesc, (j > 0 || k8 > 0) ? 1u : 0u);
          umma_tf32x(tmem, d_al + o, d_bh + o, idesc, 1u);
          umma_tf32x(tmem, d_ah + o, d_bl + o, idesc, 1u);
        }
        umma_commit(bars + s);
        if (j == n_taps - 1) umma_commit(bars + kStages);          // everything of this tile
      }
    }
    commits += (uint32_t)n_taps;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
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
    const int o = o0 + row;
    if (has_cols && o < n_out) {
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + half * 16 + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + half * 16 + c));
        float4 y;
        y.x = fmaxf(fmaf(acc[c], sc.x, sh.x), 0.0f);
        y.y = fmaxf(fmaf(acc[c + 1], sc.y, sh.y), 0.0f);
        y.z = fmaxf(fmaf(acc[c + 2], sc.z, sh.z), 0.0f);
        y.w = fmaxf(fmaf(acc[c + 3], sc.w, sh.w), 0.0f);
        const float4 hi = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
        float* dst = out_split + (size_t)o * (2 * COUT) + half * 16 + c;
        *reinterpret_cast<float4*>(dst) = hi;
        *reinterpret_cast<float4*>(dst + COUT) = make_float4(y.x - hi.x, y.y - hi.y, y.z - hi.z, y.w - hi.w);
        if (out_full) *reinterpret_cast<float4*>(out_full + (size_t)o * COUT + half * 16 + c) = y;
      }
    }
    tc_fence_before();
    __syncthreads();                                   // TMEM reads done before the next tile's first MMA
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// x0 rows of the first layer: feat_out[j] = split(feat_in[rows[j]])
__global__ void __launch_bounds__(256) sc_gather_rows_split(const float* __restrict__ in, int C, const int32_t* __restrict__ rows,
                                                            const int32_t* __restrict__ n_dev, float* __restrict__ out) {
  const long long n = (long long)__ldg(n_dev) * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / C), c = (int)(t - (long long)j * C);
    const float v = __ldg(in + (size_t)__ldg(rows + j) * C + c), hi = tf32_hi(v);
    out[(size_t)j * 2 * C + c] = hi;
    out[(size_t)j * 2 * C + C + c] = v - hi;
  }
}

template <int CIN, int COUT>
int launch(const float* in_split, const int32_t* nbr, const int32_t* n_out_dev, int n_out_max, const float* w_packed,
           const float* scale, const float* shift, float* out_split, float* out_full, cudaStream_t st) {
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
  sc_conv_tc<CIN, COUT><<<(int)tiles, kThreads, L::SMEM, st>>>(in_split, nbr, n_out_max, n_out_dev, w_packed, scale, shift,
                                                               out_split, out_full);
  return check_launch("sc_conv_tc");
}

}  // namespace
}  // namespace gpnerf

using namespace gpnerf;

extern "C" {

int gpnerf_sc_conv_tc(const float* in_split, int c_in, const int32_t* nbr, const int32_t* n_out_dev, int n_out_max,
                      const float* w_packed, const float* scale, const float* shift, int c_out, float* out_split,
                      float* out_full, void* stream) {
  GPNERF_REQUIRE(in_split && nbr && n_out_dev && w_packed && scale && shift && out_split && n_out_max > 0);
  cudaStream_t st = (cudaStream_t)stream;
#define GPNERF_SC(CI, CO)        \
  if (c_in == CI && c_out == CO) \
    return launch<CI, CO>(in_split, nbr, n_out_dev, n_out_max, w_packed, scale, shift, out_split, out_full, st);
  GPNERF_SC(16, 16) GPNERF_SC(16, 32) GPNERF_SC(32, 32) GPNERF_SC(32, 16)
#undef GPNERF_SC
  set_error("sc_conv_tc supports channel widths 16 and 32", cudaSuccess);
  return GPNERF_E_UNSUPPORTED;
}

int gpnerf_sc_gather_rows_split(const float* feat_in, int C, const int32_t* rows, const int32_t* n_dev, int n_max,
                                float* feat_out, void* stream) {
  GPNERF_REQUIRE(feat_in && rows && n_dev && feat_out && C > 0 && n_max > 0);
  long long b = ((long long)n_max * C + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  sc_gather_rows_split<<<(int)(b < cap ? b : cap), 256, 0, (cudaStream_t)stream>>>(feat_in, C, rows, n_dev, feat_out);
  return check_launch("sc_gather_rows_split");
}

}  // extern "C"
