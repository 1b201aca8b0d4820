// K3 colour head, tile-fed (tensor-core path): trainhead.py:85-100, 118-145 on the 128-point tiles of the P1 list,
// fed by the tile records the fused gather → density kernel leaves behind (RecTile<V>, tc_heads.cuh).
//
// Why tiles of the P1 list and not the compacted survivors of the progressive step (demo_render.py:312-333): the
// per-view pixel-aligned features of a point are gathered ONCE, by the fused kernel, for the density head's
// mean / variance inputs; the producers of that kernel are what bounds the frame (L1 data pipe + load latency).
// A colour head that re-gathers them for the survivors (k3_color_ws.cu, round 2's first design) pays the same
// latency-bound gather a second time and is bound by it (0.287 ms).  Here the tile arrives with one bulk copy
// (cp.async.bulk, 48 KB for 3 views) already in the tcgen05 operand layouts, no thread touches it, and the head
// runs at the speed of its GEMM round trips.  The price: the head also processes the points that die in the
// progressive step inside a tile that has survivors (≈10 % of P1 on the benchmark frame; K5 ignores their
// colour, k4_k5_progressive.cu) – tiles without a single survivor are skipped – and the record is HBM traffic
// (384 B written + read per point, ≈0.1 ms of HBM time per frame, overlapped with the heads' compute).
//
// One persistent CTA per SM:
//   loader      1 thread: for every live tile, waits for a free stage and issues the bulk copy of its record.
//   chains      4 x 4 warps, thread = row = TMEM lane, 128 TMEM columns per chain.  Lane 0 of a chain's first warp
//               issues the chain's tcgen05.mma itself (no separate issuer to wake): per view base_fc.0 (operands
//               from the stage), base_fc.2, vis_fc.0, vis_fc.2 with the hidden activations read from TMEM
//               (tcgen05.mma with a TMEM A operand), rgb_fc.0 accumulated view by view; the epilogues between
//               them (tcgen05.ld → scaled ELU → bf16 → tcgen05.st) are the chain's own warps, synchronised with
//               named barriers.  rgb_fc.2 / rgb_fc.4 + sigmoid on CUDA cores.
// rgb_fc.0's input x_v + vis(x_v) is never formed: W·(x_v + e_v) = (V·W)·(x_v / V) + W·e_v (see k3_color_ws.cu).
#include <stdlib.h>
#include "tc_heads.cuh"

namespace gpnerf {

struct ColorTilesArgs {
  const uint8_t* rec;            // [tiles] RecTile<V>
  const uint32_t* alive_words;   // optional: the progressive step's survivor flags, 4 words per tile (all zero = skip)
  const int32_t* count_ptr;      // number of P1 points
  int32_t* counters;             // with alive_words: counters[P2] receives the number of survivors (slots 6, 7: scratch)
  const uint8_t* image;          // packed colour weights (ColImg<V>)
  float* rgb;                    // [P1][3]
};

namespace ctl {
constexpr int kChains = 4;
constexpr int kLoadWarp = 4 * kChains;
constexpr int kThreads = (4 * kChains + 1) * 32;      // 544
constexpr uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }
template <int V>
struct Smem {
  static constexpr int kStages = V <= 3 ? 3 : 2;
  static constexpr uint32_t IMG = 0;
  static constexpr uint32_t STAGE0 = align_up(ColImg<V>::BYTES, 1024);
  static constexpr uint32_t STAGE_BYTES = RecTile<V>::BYTES;
  static constexpr uint32_t ONES = STAGE0 + kStages * STAGE_BYTES;    // [128 x 16] constant (…, 1, 1 | 0 x 8): bias rows
  static constexpr uint32_t MISC = ONES + 4096;
  static constexpr uint32_t BYTES = MISC + 256;
};
static_assert(Smem<3>::BYTES + 1024 <= 227 * 1024 && Smem<4>::BYTES + 1024 <= 227 * 1024, "one CTA per SM");
constexpr uint32_t kSbo16 = op_sbo(16);
// TMEM columns of a chain (128): an A operand in TMEM must start on a 32-column boundary (r02_ts_probe.txt)
//   0.. 63  accumulators of base_fc.0 (64), base_fc.2 / vis_fc.0 / vis_fc.2 (32 each, columns 0..31)
//  32.. 47  x_v / V        (written after base_fc.0's accumulator has been consumed, read by vis_fc.0 and rgb_fc.0)
//  64.. 95  H_v (64 values), then Y_v, E_v (32 values each)
//  96..127  rgb_fc.0 accumulator (all views)
constexpr uint32_t ACC = 0, XS = 32, ACT = 64, ACC5 = 96, kChainCols = 128;

__device__ __forceinline__ void wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test(bar, parity)) {
    __nanosleep(32);
    if (++spins > 8000000u) __trap();        // a protocol bug must surface as a trap, not as a hung GPU
  }
}
}  // namespace ctl

template <int V>
__global__ void __launch_bounds__(ctl::kThreads, 1) color_tiles_ws(ColorTilesArgs a) {
  using namespace ctl;
  using I = ColImg<V>;
  using S = Smem<V>;
  using R = RecTile<V>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* img = smem + S::IMG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::MISC);
  uint64_t* bar_w = bars;
  uint64_t* full = bars + 1;                 // [kStages] bulk copy landed
  uint64_t* empty = full + kStages;          // [kStages] tcgen05.commit: the last GEMM that reads the stage is done
  uint64_t* m2e = empty + kStages;           // [kChains] tcgen05.commit → the chain's warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(m2e + kChains);
  volatile int* next_start = reinterpret_cast<volatile int*>(smem + S::MISC + 192);
  const float* fl = reinterpret_cast<const float*>(img + I::F32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int c = 0; c < kChains; ++c) mbar_init(m2e + c, 1);
    *next_start = 0;
    fence_mbar_init();
    mbar_arrive_expect_tx(bar_w, I::BYTES);
    bulk_g2s(img, a.image, I::BYTES, bar_w);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid < 128) {
    uint8_t* o = smem + S::ONES;
    *reinterpret_cast<uint4*>(o + chunk_off(tid, 0, kSbo16)) = make_uint4(0u, 0u, 0u, 0x3F803F80u);   // bf16 1.0 in columns 6, 7
    *reinterpret_cast<uint4*>(o + chunk_off(tid, 1, kSbo16)) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n = __ldg(a.count_ptr);
  const int n_tiles = (n + 127) / 128;
  const int G = gridDim.x;
  // a tile none of whose points survived the progressive step needs no colour
  auto live = [&](int tile) -> bool {
    if (a.alive_words == nullptr) return true;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(a.alive_words) + tile);
    return (w.x | w.y | w.z | w.w) != 0u;
  };

  if (warp == kLoadWarp) {
    // =========================================================== loader
    if (lane == 0) {
      int i = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += G) {
        if (!live(tile)) continue;
        const int s = i % kStages;
        wait_backoff(empty + s, ((i / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(full + s, R::BYTES);
        bulk_g2s(smem + S::STAGE0 + s * S::STAGE_BYTES, a.rec + (size_t)tile * R::BYTES, R::BYTES, full + s);
        ++i;
      }
    } else if (lane >= 4 && lane < 8 && a.alive_words != nullptr && a.counters != nullptr) {
      // the progressive step's survivor count (counters[P2], demo_render.py:312-317) without a compaction pass: four
      // idle lanes count the flag bits of this CTA's tiles; the last CTA to finish publishes the total and re-arms
      // the two scratch slots (zero at allocation, zero again when the kernel ends)
      int cnt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += G) cnt += __popc(__ldg(a.alive_words + tile * 4 + (lane - 4)));
      cnt += __shfl_xor_sync(0xf0u, cnt, 1);
      cnt += __shfl_xor_sync(0xf0u, cnt, 2);
      if (lane == 4) {
        int* acc = a.counters + GPNERF_N_COUNTERS - 2;
        int* ticket = a.counters + GPNERF_N_COUNTERS - 1;
        atomicAdd(acc, cnt);
        __threadfence();
        if (atomicAdd(ticket, 1) == G - 1) {
          __threadfence();
          a.counters[GPNERF_CNT_P2] = atomicExch(acc, 0);
          *ticket = 0;
        }
      }
    }
  } else {
    // =========================================================== chains
    const int c = warp >> 2, wq = warp & 3;
    const int row = wq * 32 + lane;
    const bool leader = wq == 0;
    const uint32_t tb_issue = tmem + c * kChainCols;                              // lane field 0: MMA addresses
    const uint32_t tb = tb_issue + ((uint32_t)(wq * 32) << 16);                  // this warp's 32 TMEM lanes
    mbar_wait(bar_w, 0);
    const uint32_t wimg = smem_u32(img);
    const uint64_t ones_d = make_smem_desc(smem_u32(smem + S::ONES), kLBO, kSbo16);
    const uint32_t id64 = make_idesc_bf16(128, 64), id32 = make_idesc_bf16(128, 32);
    auto bdesc = [&](uint32_t w_off, int k16, int Kp) {
      return make_smem_desc(wimg + w_off + k16 * 2 * kLBO, kLBO, op_sbo(Kp));
    };
    // base_fc.0 of view v: [mean|var] (80 columns, bias in 70/71) + feat_v (32) + rgb_v (its 16-column block) → 64
    auto issue_r1 = [&](uint32_t stage, int v) {
      for (int k16 = 0; k16 < 4; ++k16)
        umma_bf16(tb_issue + ACC, make_smem_desc_sw128(stage + R::G64 + k16 * 32), bdesc(I::Wb0a, k16, 80), id64, k16 > 0);
      umma_bf16(tb_issue + ACC, make_smem_desc(stage + R::TAIL, kLBO, kSbo16), bdesc(I::Wb0a, 4, 80), id64, 1u);
      for (int k16 = 0; k16 < 2; ++k16) {
        const uint64_t ad = ((V & 1) && v == V - 1)
                                ? make_smem_desc(stage + R::FF_ODD + k16 * 2 * kLBO, kLBO, R::kOddSbo)
                                : make_smem_desc_sw128(stage + R::FF + (v >> 1) * 16384 + ((v & 1) * 2 + k16) * 32);
        umma_bf16(tb_issue + ACC, ad, bdesc(I::Wb0b, k16, 48), id64, 1u);
      }
      umma_bf16(tb_issue + ACC, make_smem_desc(stage + R::RGBS, kLBO, kSbo16),
                make_smem_desc(wimg + I::Wb0r + v * op_bytes(64, 16), kLBO, kSbo16), id64, 1u);
    };
    uint32_t ph = 0;
    // the chain's GEMMs of this round have completed: the leader polls the mbarrier, the other three warps block
    // on a named barrier (a poll is a shared-memory access; a parked warp costs nothing)
    auto wait_acc = [&]() {
      mbar_wait(m2e + c, ph);          // mbarrier.try_wait with a suspend-time hint: the warp sleeps in hardware
      ph ^= 1u;
      tc_fence_after();
    };
    // the chain's warps have written their activations to TMEM / read the accumulator: the leader may issue
    auto meet = [&]() {
      tc_fence_before();
      if (leader) {
        asm volatile("bar.sync %0, 128;" ::"r"(2 + 2 * c) : "memory");
        tc_fence_after();
      } else {
        asm volatile("bar.arrive %0, 128;" ::"r"(2 + 2 * c) : "memory");
      }
    };
    // 32 accumulator columns → scaled ELU (times `scale`) → 16 packed bf16 pairs → 16 TMEM columns of the next operand
    auto epi32 = [&](uint32_t col, uint32_t dst_col, float scale) {
      uint32_t r[32], pk[16];
      tmem_ld32(tb + col, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        pk[j] = pack_bf16x2(scale * elu_scaled(__uint_as_float(r[2 * j])), scale * elu_scaled(__uint_as_float(r[2 * j + 1])));
      tmem_st16(tb + dst_col, pk);
    };
    const float inv_views = 1.0f / (float)V;
    // start of the chain's next tile: stage full → base_fc.0 of view 0.  Tiles start strictly in order (a parity
    // wait on full[s] tells only two consecutive phases apart; with more chains than stages a chain could otherwise
    // take an earlier tenant's phase for its own).  `block` = false: only if it can start right now.
    auto try_start = [&](int i, bool block) -> bool {      // leader warp only; lane 0 decides and issues
      const int s = i % kStages;
      if (!block) {
        int ready = 0;
        if (lane == 0) ready = (*next_start == i && mbar_test(full + s, (i / kStages) & 1)) ? 1 : 0;
        if (!__shfl_sync(0xffffffffu, ready, 0)) return false;
      }
      if (lane == 0) {
        uint32_t spins = 0;
        while (*next_start != i) {
          __nanosleep(64);
          if (++spins > 8000000u) __trap();
        }
        wait_backoff(full + s, (i / kStages) & 1);
        tc_fence_after();
        issue_r1(smem_u32(smem + S::STAGE0 + s * S::STAGE_BYTES), 0);
        umma_commit(m2e + c);
        if (V == 1) umma_commit(empty + s);
        __threadfence_block();
        *next_start = i + 1;
      }
      __syncwarp();
      return true;
    };
    int i = 0;                     // CTA-local index over the live tiles
    bool started = false;          // the current tile's first GEMM has been issued already (by the previous tile's tail)
    for (int tile = blockIdx.x; tile < n_tiles; tile += G) {
      if (!live(tile)) continue;
      if (i % kChains != c) {
        ++i;
        continue;
      }
      const int s = i % kStages;
      const uint32_t stage = smem_u32(smem + S::STAGE0 + s * S::STAGE_BYTES);
      if (leader && !started) try_start(i, true);
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        // ---- base_fc.0 → H_v (64 values) ; then base_fc.2 (64 → 32)
        wait_acc();
        epi32(ACC, ACT, 1.0f);
        epi32(ACC + 32, ACT + 16, 1.0f);
        tmem_wait_st();
        meet();
        if (leader && lane == 0) {
          for (int k16 = 0; k16 < 4; ++k16) umma_ts(tb_issue + ACC, tb_issue + ACT + k16 * 8, bdesc(I::Wb1, k16, 80), id32, k16 > 0);
          umma_bf16(tb_issue + ACC, ones_d, bdesc(I::Wb1, 4, 80), id32, 1u);
          umma_commit(m2e + c);
        }
        // ---- base_fc.2 → x_v, stored as x_v / V (vis_fc input, trainhead.py:140) ; then vis_fc.0 (32 → 32) and this
        // view's residual share of rgb_fc.0: (V·W_v)·(x_v / V)
        wait_acc();
        epi32(ACC, XS, inv_views);
        tmem_wait_st();
        meet();
        if (leader && lane == 0) {
          for (int k16 = 0; k16 < 2; ++k16) umma_ts(tb_issue + ACC, tb_issue + XS + k16 * 8, bdesc(I::Wv0, k16, 48), id32, k16 > 0);
          umma_bf16(tb_issue + ACC, ones_d, bdesc(I::Wv0, 2, 48), id32, 1u);
          umma_commit(m2e + c);
          for (int k16 = 0; k16 < 2; ++k16)
            umma_ts(tb_issue + ACC5, tb_issue + XS + k16 * 8, bdesc(I::Wr0x, 2 * v + k16, 32 * V), id32, (v > 0 || k16 > 0));
        }
        // ---- vis_fc.0 → Y_v ; then vis_fc.2 (32 → 32)
        wait_acc();
        epi32(ACC, ACT, 1.0f);
        tmem_wait_st();
        meet();
        if (leader && lane == 0) {
          for (int k16 = 0; k16 < 2; ++k16) umma_ts(tb_issue + ACC, tb_issue + ACT + k16 * 8, bdesc(I::Wv1, k16, 48), id32, k16 > 0);
          umma_bf16(tb_issue + ACC, ones_d, bdesc(I::Wv1, 2, 48), id32, 1u);
          umma_commit(m2e + c);
        }
        // ---- vis_fc.2 → E_v ; then its share of rgb_fc.0 (bias added by the final epilogue) and the next view's
        // base_fc.0: its accumulator (columns 0..63) overlaps vis_fc.2's and x_v / V, all consumed by now
        wait_acc();
        epi32(ACC, ACT, 1.0f);
        tmem_wait_st();
        meet();
        if (leader && lane == 0) {
          for (int k16 = 0; k16 < 2; ++k16)
            umma_ts(tb_issue + ACC5, tb_issue + ACT + k16 * 8, bdesc(I::Wr0, 2 * v + k16, 32 * V + 16), id32, 1u);
          if (v + 1 < V) {
            issue_r1(stage, v + 1);
            if (v + 2 == V) umma_commit(empty + s);      // the last GEMM that reads the stage
          }
          umma_commit(m2e + c);
        }
      }
      // ---- rgb_fc.0 (bias added here) → z ; the chain's next tile is started before the CUDA-core tail if its stage
      // has landed ; rgb_fc.2 (32 → 16) and rgb_fc.4 (16 → 3) on CUDA cores ; sigmoid
      wait_acc();
      float z[32];
      {
        uint32_t r[32];
        tmem_ld32(tb + ACC5, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = elu_scaled(__uint_as_float(r[j]) + fl[I::rb0c + j]);
      }
      meet();
      // the chain's next live tile, if any
      int i_next = -1;
      {
        int ii = i + 1;
        for (int t2 = tile + G; t2 < n_tiles; t2 += G) {
          if (!live(t2)) continue;
          if (ii % kChains == c) {
            i_next = ii;
            break;
          }
          ++ii;
        }
      }
      started = false;
      if (leader && i_next >= 0) started = try_start(i_next, false);
      float o0 = fl[I::rb2], o1 = fl[I::rb2 + 1], o2 = fl[I::rb2 + 2];
#pragma unroll 4
      for (int nn = 0; nn < 16; ++nn) {
        float hsum = fl[I::b1f + nn];
        const float4* wrow = reinterpret_cast<const float4*>(fl + I::w1f + nn * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 wv = wrow[j4];
          hsum = fmaf(wv.x, z[4 * j4], fmaf(wv.y, z[4 * j4 + 1], fmaf(wv.z, z[4 * j4 + 2], fmaf(wv.w, z[4 * j4 + 3], hsum))));
        }
        const float h = hsum > 0.0f ? hsum : ex2_ftz(hsum * kLog2e) - 1.0f;
        o0 = fmaf(h, fl[I::rw2u + nn], o0);
        o1 = fmaf(h, fl[I::rw2u + 16 + nn], o1);
        o2 = fmaf(h, fl[I::rw2u + 32 + nn], o2);
      }
      const long long p = (long long)tile * 128 + row;
      if (p < n) {
        a.rgb[p * 3 + 0] = sigmoid_fast(o0);
        a.rgb[p * 3 + 1] = sigmoid_fast(o1);
        a.rgb[p * 3 + 2] = sigmoid_fast(o2);
      }
      ++i;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace gpnerf

using namespace gpnerf;

template <int V>
static int launch_color_tiles(const ColorTilesArgs& a, int n_points_max, cudaStream_t st) {
  static bool attr_set = false;
  constexpr uint32_t bytes = ctl::Smem<V>::BYTES + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(color_tiles_ws<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
      set_error("color_tiles_ws smem attribute", e);
      return GPNERF_E_CUDA;
    }
    attr_set = true;
  }
  const int tiles = (n_points_max + 127) / 128;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  color_tiles_ws<V><<<grid, ctl::kThreads, bytes, st>>>(a);
  return check_launch("k3_color_tiles_ws");
}

extern "C" {

int64_t gpnerf_k23_tile_record_bytes(int n_views) {
  return (n_views >= 1 && n_views <= 4) ? (int64_t)rec_tile_bytes(n_views) : (int64_t)GPNERF_E_ARG;
}

int gpnerf_k3_color_tiles_tc(const void* tile_records, const void* k4_workspace, const gpnerf_head_weights_t* w,
                             int n_views, int n_points_max, int32_t* counters, int counter_slot, float* rgb,
                             void* stream) {
  GPNERF_REQUIRE(tile_records && w && counters && rgb);
  GPNERF_REQUIRE(n_points_max > 0 && counter_slot >= 0 && counter_slot < GPNERF_N_COUNTERS);
  GPNERF_REQUIRE(w->tc_image != nullptr);
  ColorTilesArgs a;
  a.rec = reinterpret_cast<const uint8_t*>(tile_records);
  a.alive_words = k4_workspace ? carve_workspace(const_cast<void*>(k4_workspace), n_points_max).words : nullptr;
  a.count_ptr = counters + counter_slot;
  a.counters = counters;
  a.image = reinterpret_cast<const uint8_t*>(w->tc_image) + kColImgOffset;
  a.rgb = rgb;
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_views) {
    case 1: return launch_color_tiles<1>(a, n_points_max, st);
    case 2: return launch_color_tiles<2>(a, n_points_max, st);
    case 3: return launch_color_tiles<3>(a, n_points_max, st);
    case 4: return launch_color_tiles<4>(a, n_points_max, st);
    default:
      set_error("tcgen05 colour head supports 1..4 source views", cudaSuccess);
      return GPNERF_E_UNSUPPORTED;
  }
}

}  // extern "C"
