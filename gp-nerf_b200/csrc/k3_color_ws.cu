// K3 colour head, warp-specialised (tensor-core path): trainhead.py:85-100, 118-145 for the survivors of the
// progressive step (demo_render.py:312-333) – and for them only: the head gathers its own inputs (the V-view
// pixel-aligned features, BaseRender.py:283-363, and their mean / variance, trainhead.py:20-24) from the
// L2-resident feature maps, so nothing is written per point upstream of the progressive step (round 1 wrote a
// 384-byte record for every P1 point, 600 MB of DRAM traffic per frame, most of it for points that die).
//
// One persistent CTA per SM, three roles:
//   producers   8 warps: survivor t → P1 index valid1[t] → sample (ray, z) → projection into the V views →
//               bilinear feature / RGB taps (HFMA2 on fp16 storage, 4 lanes x 8 channels per point) → the operand
//               tiles of a stage: [mean|var] 64 columns and the per-view features (SWIZZLE_128B), the RGB
//               mean/var + bias columns and the per-view RGB block (16 columns each).
//   MMA issuer  1 thread: for each of the 3 colour chains (one 128-point tile each) the layer GEMMs, view by
//               view: base_fc.0 (128 → 64, operands from the stage), base_fc.2 (64 → 32), vis_fc.0/.2 (32 → 32),
//               then rgb_fc.0 (32V → 32, accumulated view by view as soon as a view's inputs exist) and rgb_fc.2
//               (32 → 16).  Hidden activations are read from TMEM (tcgen05.mma with a TMEM A operand).  The next
//               view's base_fc.0 is issued while the current view's epilogues run.
//   epilogues   3 x 4 warps, thread = row = TMEM lane: tcgen05.ld, scaled ELU, bf16, tcgen05.st as the next A
//               operand; rgb_fc.4 (16 → 3) + sigmoid on CUDA cores.
// rgb_fc.0's input x_v + vis(x_v) is never formed: W·(x_v + e_v) = (V·W)·(x_v / V) + W·e_v, and x_v / V is the
// vis_fc input that sits in TMEM anyway – no residual registers, two more small MMAs per view.
#include <stdlib.h>
#include "tc_heads.cuh"

namespace gpnerf {

struct ColorWsArgs {
  const __half* feat;            // [V][fh+2][fw+2][32]
  const float4* rgbx;            // [V][H+2][W+2] (r,g,b,·) in [0,1]
  const int32_t* valid;          // P1 list: flat sample indices
  const int32_t* valid1;         // optional: indices into the P1 arrays (the survivors); NULL = all of them
  const float *rays_o, *rays_d, *z_vals;
  const int32_t* count_ptr;
  const uint8_t* image;          // packed colour weights (ColImg<V>)
  float* rgb;                    // [P1][3], written at the P1 index of every processed point
  float* rgb_in;                 // optional [P1][V][3]: the per-view RGB taps (BaseRender's rgb_in_map input)
};


namespace cws {
constexpr int kProdWarps = 8, kChains = 3, kStages = 2;
constexpr int kPasses = 16 / kProdWarps;          // passes of 8 points per producer warp and tile
constexpr int kThreads = (kProdWarps + 4 * kChains + kChains) * 32;  // 736
constexpr int kMmaWarp0 = kProdWarps + 4 * kChains;                  // one MMA-issuing warp per chain
constexpr int kChainCols = 160;                                      // TMEM columns per chain
constexpr uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }
template <int V>
struct Smem {
  static constexpr uint32_t IMG = 0;
  static constexpr uint32_t STAGE0 = align_up(ColImg<V>::BYTES, 1024);
  static constexpr uint32_t G64 = 0;                      // [128 x 64] bf16 mean_feat | var_feat, SWIZZLE_128B
  static constexpr uint32_t FF = 16384;                   // 2 x [128 x 64] bf16: view v in block v / 2, columns 32 (v % 2) …
  static constexpr uint32_t TAIL = FF + 2 * 16384;        // [128 x 16] (mean/var rgb, 1, 1 | 0 x 8), core-matrix layout
  static constexpr uint32_t RGBS = TAIL + 4096;           // [128 x 16] column 3v + c = channel c of view v
  static constexpr uint32_t STAGE_BYTES = RGBS + 4096;    // 56 KB
  static constexpr uint32_t ONES = STAGE0 + kStages * STAGE_BYTES;
  static constexpr uint32_t PLAN = ONES + 4096;           // kProdWarps x 4 entries x (8 points + 1 pad) x 16 bytes
  static constexpr uint32_t DST = PLAN + kProdWarps * 576;          // kStages x 128 int32: P1 index of every row
  static constexpr uint32_t MISC = DST + kStages * 512;
  static constexpr uint32_t BYTES = MISC + 256;
};
static_assert(Smem<4>::BYTES + 1024 + 1024 <= 227 * 1024, "one CTA per SM");
constexpr uint32_t kSbo16 = op_sbo(16);
}  // namespace cws

__device__ __forceinline__ __half2 cw_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ void cw_hfma8(__half2 (&acc)[4], const uint4& q, __half2 w) {
  acc[0] = __hfma2(cw_h2(q.x), w, acc[0]);
  acc[1] = __hfma2(cw_h2(q.y), w, acc[1]);
  acc[2] = __hfma2(cw_h2(q.z), w, acc[2]);
  acc[3] = __hfma2(cw_h2(q.w), w, acc[3]);
}
__device__ __forceinline__ uint4 cw_pack8(const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  return q;
}
// see k23_fused_ws.cu: 8x4 → 4x8 lane transpose, then a conflict-free swizzled store
__device__ __forceinline__ void cw_st_rows8_sw(uint8_t* kblock, int row8, int c0, uint4 v, int lane) {
  const int src = (lane & 7) * 4 + (lane >> 3);
  v.x = __shfl_sync(0xffffffffu, v.x, src);
  v.y = __shfl_sync(0xffffffffu, v.y, src);
  v.z = __shfl_sync(0xffffffffu, v.z, src);
  v.w = __shfl_sync(0xffffffffu, v.w, src);
  *reinterpret_cast<uint4*>(kblock + sw128_off(row8 + (lane & 7), c0 + (lane >> 3))) = v;
}
__device__ __forceinline__ void cw_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test(bar, parity)) {
    __nanosleep(40);
    if (++spins > 4000000u) __trap();
  }
}

template <int V>
__global__ void __launch_bounds__(cws::kThreads, 1) color_mlp_ws(ColorWsArgs a, const __grid_constant__ gpnerf_frame_t fparam) {
  using namespace cws;
  using I = ColImg<V>;
  using S = Smem<V>;
  GPNERF_LOAD_FRAME(fparam)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* img = smem + S::IMG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::MISC);
  uint64_t* bar_w = bars;
  uint64_t* full = bars + 1;                 // [kStages] producers → MMA issuer
  uint64_t* empty = full + kStages;          // [kStages] tcgen05.commit → producers
  uint64_t* m2e = empty + kStages;           // [kChains] tcgen05.commit → epilogue group
  uint64_t* e2m = m2e + kChains;             // [kChains] epilogue group → MMA issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e2m + kChains);
  const float* fl = reinterpret_cast<const float*>(img + I::F32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, kProdWarps);
      mbar_init(empty + s, 1);
    }
    for (int c = 0; c < kChains; ++c) {
      mbar_init(m2e + c, 1);
      mbar_init(e2m + c, 128);
    }
    *reinterpret_cast<volatile int*>(smem + S::MISC + 192) = 0;       // next tile (CTA-local index) allowed to start
    fence_mbar_init();
    mbar_arrive_expect_tx(bar_w, I::BYTES);
    bulk_g2s(img, a.image, I::BYTES, bar_w);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // constants of the operand tiles: the bias rows, the zero halves of the 16-column tiles and – for an odd view
  // count – the unused half of the last feature block
  for (int t = tid; t < 128 * (1 + kStages); t += kThreads) {
    const int r = t & 127, which = t >> 7;
    if (which == 0) {
      uint8_t* o = smem + S::ONES;
      *reinterpret_cast<uint4*>(o + chunk_off(r, 0, kSbo16)) = make_uint4(0u, 0u, 0u, 0x3F803F80u);   // bf16 1.0 in columns 6, 7
      *reinterpret_cast<uint4*>(o + chunk_off(r, 1, kSbo16)) = make_uint4(0u, 0u, 0u, 0u);
    } else {
      uint8_t* st = smem + S::STAGE0 + (which - 1) * S::STAGE_BYTES;
      *reinterpret_cast<uint4*>(st + S::TAIL + chunk_off(r, 1, kSbo16)) = make_uint4(0u, 0u, 0u, 0u);
      if (V & 1) {
#pragma unroll
        for (int cch = 4; cch < 8; ++cch)
          *reinterpret_cast<uint4*>(st + S::FF + ((V - 1) >> 1) * 16384 + sw128_off(r, cch)) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n = __ldg(a.count_ptr);
  const int n_tiles = (n + 127) / 128;
  const int G = gridDim.x;

  if (warp < kProdWarps) {
    // =========================================================== producers
    const int grp = lane >> 2, sub = lane & 3;
    uint4* plan_base = reinterpret_cast<uint4*>(smem + S::PLAN + warp * 576) + grp;
#define plan(e) plan_base[(e) * 9]
    const float ox = __ldg(a.rays_o), oy = __ldg(a.rays_o + 1), oz = __ldg(a.rays_o + 2);
    const int Sn = f.n_samples;
    const float sfx = (float)(f.feat_w - 1) / (float)(f.src_w - 1), sfy = (float)(f.feat_h - 1) / (float)(f.src_h - 1);
    const float wm1 = (float)(f.src_w - 1), hm1 = (float)(f.src_h - 1);
    const int img_stride_p = (f.src_h + 2) * (f.src_w + 2);
    const int map_stride_q = (f.feat_h + 2) * (f.feat_w + 2) * 4;
    // survivor → P1 index → sample: three dependent global loads, fetched one tile ahead
    auto fetch_i = [&](long long first_row, int r) -> int {
      if (first_row + r >= n) return -1;
      return a.valid1 ? __ldg(a.valid1 + first_row + r) : (int)(first_row + r);
    };
    auto position = [&](int i, float& x, float& y, float& z) {
      x = y = z = 0.f;
      if (i >= 0) {
        const int q = __ldg(a.valid + i);
        const int ray = q / Sn;
        const float zz = __ldg(a.z_vals + q);
        x = fmaf(__ldg(a.rays_d + ray * 3 + 0), zz, ox);
        y = fmaf(__ldg(a.rays_d + ray * 3 + 1), zz, oy);
        z = fmaf(__ldg(a.rays_d + ray * 3 + 2), zz, oz);
      }
    };
    int icur[kPasses], inext[kPasses];
    float px[kPasses], py[kPasses], pz[kPasses];
#pragma unroll
    for (int j = 0; j < kPasses; ++j) {
      icur[j] = fetch_i((long long)blockIdx.x * 128, (warp * kPasses + j) * 8 + grp);
      position(icur[j], px[j], py[j], pz[j]);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += G, ++it) {
      const int s = it % kStages;
      uint8_t* stage = smem + S::STAGE0 + s * S::STAGE_BYTES;
      int32_t* dst_idx = reinterpret_cast<int32_t*>(smem + S::DST + s * 512);
#pragma unroll
      for (int j = 0; j < kPasses; ++j) inext[j] = fetch_i((long long)(tile + G) * 128, (warp * kPasses + j) * 8 + grp);
      cw_wait(empty + s, ((it / kStages) & 1) ^ 1);
#pragma unroll 1
      for (int j = 0; j < kPasses; ++j) {
        const int r = (warp * kPasses + j) * 8 + grp;
        const int pi = icur[j];
        const bool ok = pi >= 0;
        const float ppx = px[j], ppy = py[j], ppz = pz[j];
        if (sub < V) {
          const float* KE = f.src_KE[sub];
          const float qx = fmaf(KE[0], ppx, fmaf(KE[1], ppy, fmaf(KE[2], ppz, KE[3])));
          const float qy = fmaf(KE[4], ppx, fmaf(KE[5], ppy, fmaf(KE[6], ppz, KE[7])));
          const float qz = fmaf(KE[8], ppx, fmaf(KE[9], ppy, fmaf(KE[10], ppz, KE[11])));
          const float inv = 1.0f / qz;
          const float ux2 = qx * inv, uy2 = qy * inv;
          uint4 e;
          {
            const float ix = fminf(fmaxf(ux2 * sfx, -1.0f), (float)f.feat_w);
            const float iy = fminf(fmaxf(uy2 * sfy, -1.0f), (float)f.feat_h);
            const int x0 = min((int)floorf(ix), f.feat_w - 1), y0 = min((int)floorf(iy), f.feat_h - 1);
            e.x = ok ? (uint32_t)(sub * map_stride_q + (y0 + 1) * (f.feat_w + 2) * 4 + (x0 + 1) * 4) : 0u;
            e.z = pack_f16x2(ix - (float)x0, iy - (float)y0);
          }
          {
            const float ix = fminf(fmaxf(ux2, -1.0f), (float)f.src_w), iy = fminf(fmaxf(uy2, -1.0f), (float)f.src_h);
            const int x0 = min((int)floorf(ix), f.src_w - 1), y0 = min((int)floorf(iy), f.src_h - 1);
            e.y = ok ? (uint32_t)(sub * img_stride_p + (y0 + 1) * (f.src_w + 2) + (x0 + 1)) : 0u;
            e.w = pack_f16x2(ix - (float)x0, iy - (float)y0);
          }
          plan(sub) = e;
        }
        if (sub == 0) dst_idx[r] = pi;
        __syncwarp();
        // ---- V source views: feature taps by all 4 lanes (8 channels each)
        float fv[V][8];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const uint4 e = plan(v);
          const int dy = (f.feat_w + 2) * 4;
          const uint4* p0 = reinterpret_cast<const uint4*>(a.feat) + (e.x + sub);
          uint4 q[4];
          q[0] = __ldg(p0);
          q[1] = __ldg(p0 + 4);
          q[2] = __ldg(p0 + dy);
          q[3] = __ldg(p0 + dy + 4);
          const float2 wf = __half22float2(cw_h2(e.z));
          const float wx = wf.x, wy = wf.y;
          __half2 acc[4] = {__float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f)};
          cw_hfma8(acc, q[0], __float2half2_rn((1.0f - wx) * (1.0f - wy)));
          cw_hfma8(acc, q[1], __float2half2_rn(wx * (1.0f - wy)));
          cw_hfma8(acc, q[2], __float2half2_rn((1.0f - wx) * wy));
          cw_hfma8(acc, q[3], __float2half2_rn(wx * wy));
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float2 t = __half22float2(acc[jj]);
            fv[v][2 * jj] = t.x;
            fv[v][2 * jj + 1] = t.y;
          }
          cw_st_rows8_sw(stage + S::FF + (v >> 1) * 16384, r & ~7, (v & 1) * 4, cw_pack8(fv[v]), lane);
        }
        // ---- RGB taps: lane `sub` takes view `sub`
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (sub < V) {
          const uint4 e = plan(sub);
          const int dy = f.src_w + 2;
          const float4* p0 = a.rgbx + e.y;
          const float4 t0 = __ldg(p0), t1 = __ldg(p0 + 1), t2 = __ldg(p0 + dy), t3 = __ldg(p0 + dy + 1);
          const float2 wr = __half22float2(cw_h2(e.w));
          const float wx = wr.x, wy = wr.y;
          const float w0 = (1.0f - wx) * (1.0f - wy), w1 = wx * (1.0f - wy), w2 = (1.0f - wx) * wy, w3 = wx * wy;
          c0 = fmaf(t3.x, w3, fmaf(t2.x, w2, fmaf(t1.x, w1, t0.x * w0)));
          c1 = fmaf(t3.y, w3, fmaf(t2.y, w2, fmaf(t1.y, w1, t0.y * w0)));
          c2 = fmaf(t3.z, w3, fmaf(t2.z, w2, fmaf(t1.z, w1, t0.z * w0)));
          if (a.rgb_in != nullptr && ok) {
            float* o = a.rgb_in + ((long long)pi * V + sub) * 3;
            o[0] = c0; o[1] = c1; o[2] = c2;
          }
        }
        {
          const float inv_v = 1.0f / (float)V;
          float mean[8], var[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float m = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) m += fv[v][e];
            m *= inv_v;
            float sq = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) sq = fmaf(fv[v][e] - m, fv[v][e] - m, sq);
            mean[e] = m;
            var[e] = sq * inv_v;
          }
          cw_st_rows8_sw(stage + S::G64, r & ~7, 0, cw_pack8(mean), lane);
          cw_st_rows8_sw(stage + S::G64, r & ~7, 4, cw_pack8(var), lane);
          // per-view RGB block: lane 0 of the point collects (r,g,b) of the V views from lanes 0..V-1
          float rgbv[12];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            rgbv[3 * v + 0] = __shfl_sync(0xffffffffu, c0, (lane & ~3) + v);
            rgbv[3 * v + 1] = __shfl_sync(0xffffffffu, c1, (lane & ~3) + v);
            rgbv[3 * v + 2] = __shfl_sync(0xffffffffu, c2, (lane & ~3) + v);
          }
          float m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
          for (int v = 0; v < V; ++v) { m0 += rgbv[3 * v]; m1 += rgbv[3 * v + 1]; m2 += rgbv[3 * v + 2]; }
          m0 *= inv_v; m1 *= inv_v; m2 *= inv_v;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int v = 0; v < V; ++v) {
            s0 = fmaf(rgbv[3 * v] - m0, rgbv[3 * v] - m0, s0);
            s1 = fmaf(rgbv[3 * v + 1] - m1, rgbv[3 * v + 1] - m1, s1);
            s2 = fmaf(rgbv[3 * v + 2] - m2, rgbv[3 * v + 2] - m2, s2);
          }
          if (sub == 0) {
            const float t[8] = {m0, m1, m2, s0 * inv_v, s1 * inv_v, s2 * inv_v, 1.0f, 1.0f};
            *reinterpret_cast<uint4*>(stage + S::TAIL + chunk_off(r, 0, kSbo16)) = cw_pack8(t);
            const float u0[8] = {rgbv[0], rgbv[1], rgbv[2], rgbv[3], rgbv[4], rgbv[5], rgbv[6], rgbv[7]};
            const float u1[8] = {rgbv[8], V > 3 ? rgbv[9] : 0.f, V > 3 ? rgbv[10] : 0.f, V > 3 ? rgbv[11] : 0.f, 0.f, 0.f, 0.f, 0.f};
            *reinterpret_cast<uint4*>(stage + S::RGBS + chunk_off(r, 0, kSbo16)) = cw_pack8(u0);
            *reinterpret_cast<uint4*>(stage + S::RGBS + chunk_off(r, 1, kSbo16)) = cw_pack8(u1);
          }
        }
        __syncwarp();
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full + s);
#pragma unroll
      for (int j = 0; j < kPasses; ++j) {
        icur[j] = inext[j];
        position(icur[j], px[j], py[j], pz[j]);
      }
    }
#undef plan
  } else if (warp >= kMmaWarp0) {
    // =========================================================== MMA issuers: one thread per chain
    // (a single issuer serving all chains through a polling state machine was the bottleneck of the first
    // version: ~40 services per three tiles, each a few hundred cycles of descriptor arithmetic on one thread)
    if (lane == 0) {
      const int c = warp - kMmaWarp0;
      volatile int* next_start = reinterpret_cast<volatile int*>(smem + S::MISC + 192);
      mbar_wait(bar_w, 0);
      const uint32_t wimg = smem_u32(img);
      const uint64_t ones_d = make_smem_desc(smem_u32(smem + S::ONES), kLBO, kSbo16);
      const uint32_t id64 = make_idesc_bf16(128, 64), id32 = make_idesc_bf16(128, 32);
      auto bdesc = [&](uint32_t w_off, int k16, int Kp) {
        return make_smem_desc(wimg + w_off + k16 * 2 * kLBO, kLBO, op_sbo(Kp));
      };
      // per-chain TMEM columns.  An A operand in TMEM must start on a 32-column boundary (an operand at column 112
      // was read as garbage for part of the rows): H, Y and E share columns 64.., each dead before the next is written
      constexpr uint32_t ACC1 = 0, ACC2 = 0, ACC3 = 32, ACC4 = 0, ACC5 = 128;
      constexpr uint32_t H_ = 64, XS = 96, Y_ = 64, E_ = 64;
      const uint32_t tb = tmem + c * kChainCols;
      // base_fc.0 of view v: [mean|var] (80 columns, bias in 70/71) + feat_v (32) + rgb_v (16-column block) → 64
      auto issue_r1 = [&](uint32_t stage, int v) {
        for (int k16 = 0; k16 < 4; ++k16)
          umma_bf16(tb + ACC1, make_smem_desc_sw128(stage + S::G64 + k16 * 32), bdesc(I::Wb0a, k16, 80), id64, k16 > 0);
        umma_bf16(tb + ACC1, make_smem_desc(stage + S::TAIL, kLBO, kSbo16), bdesc(I::Wb0a, 4, 80), id64, 1u);
        for (int k16 = 0; k16 < 2; ++k16)
          umma_bf16(tb + ACC1, make_smem_desc_sw128(stage + S::FF + (v >> 1) * 16384 + ((v & 1) * 2 + k16) * 32),
                    bdesc(I::Wb0b, k16, 48), id64, 1u);
        umma_bf16(tb + ACC1, make_smem_desc(stage + S::RGBS, kLBO, kSbo16),
                  make_smem_desc(wimg + I::Wb0r + v * op_bytes(64, 16), kLBO, kSbo16), id64, 1u);
      };
      uint32_t eph = 0;
      auto wait_epi = [&]() {          // the chain's epilogue group has published its activations
        uint32_t spins = 0;
        while (!mbar_test(e2m + c, eph)) {
          __nanosleep(32);
          if (++spins > 8000000u) __trap();
        }
        eph ^= 1u;
        tc_fence_after();
      };
      for (int i = c; blockIdx.x + (long long)i * G < n_tiles; i += kChains) {
        const int s = i % kStages;
        const uint32_t stage = smem_u32(smem + S::STAGE0 + s * S::STAGE_BYTES);
        // tiles start strictly in order: a parity wait on full[s] tells only two consecutive phases apart, and with
        // fewer stages than chains a chain would otherwise see "its" phase of a stage complete one tenant early
        {
          uint32_t spins = 0;
          while (*next_start != i) {
            __nanosleep(64);
            if (++spins > 8000000u) __trap();
          }
        }
        if (i >= kChains) wait_epi();                // chain free: the final epilogue of its previous tile is through
        cw_wait(full + s, (i / kStages) & 1);
        tc_fence_after();
        issue_r1(stage, 0);
        umma_commit(m2e + c);
        __threadfence_block();
        *next_start = i + 1;
#pragma unroll 1
        for (int v = 0; v < V; ++v) {
          // after H_v: base_fc.2 (64 → 32)
          wait_epi();
          for (int k16 = 0; k16 < 4; ++k16) umma_ts(tb + ACC2, tb + H_ + k16 * 8, bdesc(I::Wb1, k16, 80), id32, k16 > 0);
          umma_bf16(tb + ACC2, ones_d, bdesc(I::Wb1, 4, 80), id32, 1u);
          umma_commit(m2e + c);
          // one view only: base_fc.0 was the last reader of the stage (and the epilogue has fetched its rows'
          // destinations from it by now)
          if (V == 1) umma_commit(empty + s);
          // after Xs_v: vis_fc.0 (32 → 32) and this view's residual share of rgb_fc.0: (V·W_v)·(x_v / V)
          wait_epi();
          for (int k16 = 0; k16 < 2; ++k16) umma_ts(tb + ACC3, tb + XS + k16 * 8, bdesc(I::Wv0, k16, 48), id32, k16 > 0);
          umma_bf16(tb + ACC3, ones_d, bdesc(I::Wv0, 2, 48), id32, 1u);
          umma_commit(m2e + c);
          for (int k16 = 0; k16 < 2; ++k16)
            umma_ts(tb + ACC5, tb + XS + k16 * 8, bdesc(I::Wr0x, 2 * v + k16, 32 * V), id32, (v > 0 || k16 > 0));
          // after Y_v: vis_fc.2 (32 → 32)
          wait_epi();
          for (int k16 = 0; k16 < 2; ++k16) umma_ts(tb + ACC4, tb + Y_ + k16 * 8, bdesc(I::Wv1, k16, 48), id32, k16 > 0);
          umma_bf16(tb + ACC4, ones_d, bdesc(I::Wv1, 2, 48), id32, 1u);
          umma_commit(m2e + c);
          // after E_v: its share of rgb_fc.0 (the bias is added by the epilogue); then the next view's base_fc.0 – its
          // accumulator, columns 0..63, overlaps those of base_fc.2 / vis_fc.0 / vis_fc.2, all consumed by now
          wait_epi();
          for (int k16 = 0; k16 < 2; ++k16)
            umma_ts(tb + ACC5, tb + E_ + k16 * 8, bdesc(I::Wr0, 2 * v + k16, 32 * V + 16), id32, 1u);
          if (v + 1 < V) {
            issue_r1(stage, v + 1);
            if (v + 2 == V) umma_commit(empty + s);      // the last GEMM that reads the stage
          }
          umma_commit(m2e + c);
        }
      }
    }
  } else {
    // =========================================================== epilogues
    const int c = (warp - kProdWarps) >> 2;
    const int wq = (warp - kProdWarps) & 3;
    const int row = wq * 32 + lane;
    const uint32_t tb = tmem + c * kChainCols + ((uint32_t)(wq * 32) << 16);
    constexpr uint32_t ACC1 = 0, ACC2 = 0, ACC3 = 32, ACC4 = 0, ACC5 = 128;
    constexpr uint32_t H_ = 64, XS = 96, Y_ = 64, E_ = 64;
    mbar_wait(bar_w, 0);
    uint32_t ph = 0;
    // one warp of the group polls the mbarrier; the other three block on a named barrier (no polling)
    const bool leader = wq == 0;
    auto wait_acc = [&]() {
      if (leader) cw_wait(m2e + c, ph);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + c) : "memory");
      ph ^= 1u;
      tc_fence_after();
    };
    auto publish = [&]() {
      tc_fence_before();
      mbar_arrive(e2m + c);
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
    for (int i = c; blockIdx.x + (long long)i * G < n_tiles; i += kChains) {
      const int s = i % kStages;
      int dst = -1;
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        // ---- base_fc.0 → H_v (64 columns)
        wait_acc();
        if (v == 0) dst = reinterpret_cast<const int32_t*>(smem + S::DST + s * 512)[row];   // (the stage is still held)
        epi32(ACC1, H_, 1.0f);
        epi32(ACC1 + 32, H_ + 16, 1.0f);
        tmem_wait_st();
        publish();
        // ---- base_fc.2 → x_v ; stored as x_v / V (vis_fc input, trainhead.py:140)
        wait_acc();
        epi32(ACC2, XS, inv_views);
        tmem_wait_st();
        publish();
        // ---- vis_fc.0 → Y_v
        wait_acc();
        epi32(ACC3, Y_, 1.0f);
        tmem_wait_st();
        publish();
        // ---- vis_fc.2 → E_v
        wait_acc();
        epi32(ACC4, E_, 1.0f);
        tmem_wait_st();
        publish();
      }
      // ---- rgb_fc.0 (bias added here) → z ; rgb_fc.2 (32 → 16) and rgb_fc.4 (16 → 3) on CUDA cores ; sigmoid.
      // (The two small tail layers are 560 FMAs per point: cheaper than two more MMA round trips per tile – and bias
      // MMAs at these accumulator positions returned garbage for reasons not understood, see profiles/r02 notes.)
      wait_acc();
      float z[32];
      {
        uint32_t r[32];
        tmem_ld32(tb + ACC5, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = elu_scaled(__uint_as_float(r[j]) + fl[I::rb0c + j]);
      }
      publish();                                   // chain free: its next tile may overwrite the accumulators
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
      if (dst >= 0) {
        a.rgb[(long long)dst * 3 + 0] = sigmoid_fast(o0);
        a.rgb[(long long)dst * 3 + 1] = sigmoid_fast(o1);
        a.rgb[(long long)dst * 3 + 2] = sigmoid_fast(o2);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace gpnerf

using namespace gpnerf;

template <int V>
static int launch_color_ws(const ColorWsArgs& a, const gpnerf_frame_t* f, int n_points_max, cudaStream_t st) {
  static bool attr_set = false;
  constexpr uint32_t bytes = cws::Smem<V>::BYTES + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(color_mlp_ws<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
      set_error("color_mlp_ws smem attribute", e);
      return GPNERF_E_CUDA;
    }
    attr_set = true;
  }
  const int tiles = (n_points_max + 127) / 128;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  color_mlp_ws<V><<<grid, cws::kThreads, bytes, st>>>(a, *f);
  return check_launch("k3_color_mlp_ws");
}

extern "C" {


int gpnerf_k3_color_gather_tc(const void* featmaps_f16, const float* images_rgbx, const int32_t* valid,
                              const int32_t* valid1, const float* rays_o, const float* rays_d, const float* z_vals,
                              const gpnerf_frame_t* f, const gpnerf_head_weights_t* w, int n_points_max,
                              const int32_t* counters, int counter_slot, float* rgb, float* rgb_in, void* stream) {
  GPNERF_REQUIRE(featmaps_f16 && images_rgbx && valid && rays_o && rays_d && z_vals && f && w && counters && rgb);
  GPNERF_REQUIRE(n_points_max > 0 && counter_slot >= 0 && counter_slot < GPNERF_N_COUNTERS);
  GPNERF_REQUIRE(w->tc_image != nullptr && f->n_samples > 0 && f->src_w > 1 && f->src_h > 1);
  ColorWsArgs a;
  a.feat = reinterpret_cast<const __half*>(featmaps_f16);
  a.rgbx = reinterpret_cast<const float4*>(images_rgbx);
  a.valid = valid; a.valid1 = valid1; a.rays_o = rays_o; a.rays_d = rays_d; a.z_vals = z_vals;
  a.count_ptr = counters + counter_slot;
  a.image = reinterpret_cast<const uint8_t*>(w->tc_image) + kColImgOffset;
  a.rgb = rgb;
  a.rgb_in = rgb_in;
  cudaStream_t st = (cudaStream_t)stream;
  switch (f->n_views) {
    case 1: return launch_color_ws<1>(a, f, n_points_max, st);
    case 2: return launch_color_ws<2>(a, f, n_points_max, st);
    case 3: return launch_color_ws<3>(a, f, n_points_max, st);
    case 4: return launch_color_ws<4>(a, f, n_points_max, st);
    default:
      set_error("tcgen05 colour head supports 1..4 source views", cudaSuccess);
      return GPNERF_E_UNSUPPORTED;
  }
}

}  // extern "C"
