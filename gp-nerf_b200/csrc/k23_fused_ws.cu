// K2+K3 fused, warp-specialised (tensor-core path): gathers → density head for tiles of 128 surviving sample
// points, one persistent CTA per SM, three roles that overlap tile by tile:
//
//   producers   8 warps.  Warp w owns rows 16w..16w+15 of every tile (two passes of 8 points x 4 lanes): sample
//               position, per-level / per-view tap plan (one lane per level and view, exchanged through a 1 KB
//               per-warp scratch), the 4-level trilinear gather (SparseConvNet.py:111-122) and the V-view bilinear
//               gathers with mean / variance (BaseRender.py:283-363, trainhead.py:20-24), HFMA2 on fp16 storage.
//               Results land in one of kStages shared-memory stages as tcgen05 A operands in the K-major
//               SWIZZLE_128B layout (the tensor core fetches a [128 x 16] slice of it in ≈70 cycles, of the
//               unswizzled core-matrix layout in ≈125: profiles/r02_ts_probe.txt).
//   MMA issuer  1 thread.  Polls the stage-full and operand-ready barriers of the two density chains round robin
//               and issues the layer GEMMs (trainhead.py:39-41, 102-110): 128→64 (fp16 x fp16, operands from the
//               stage), 144→64, 64→32, 32→16 (bf16) with the hidden activations read straight from TMEM
//               (tcgen05.mma with a TMEM A operand): they never touch shared memory.  tcgen05.commit signals the
//               chain's epilogue warps and, after the second GEMM, hands the stage back to the producers.
//   epilogues   2 x 4 warps, one group per chain, thread = row = TMEM lane: tcgen05.ld the accumulator, scaled ELU
//               (tc_common.cuh), round to bf16, tcgen05.st it back as the next layer's A operand; last layer
//               16→1 + ReLU + no-valid-view fill on CUDA cores, α and the progressive step's survivor flags.
//
// While chain 0's tile waits for a GEMM, chain 1's tile is in its epilogue and the producers are two tiles ahead:
// the gather (L1-bound) never waits for the head (latency-bound) as it did in the monolithic kernel
// (k23_fused_tc.cu, kept behind GPNERF_FUSED_IMPL=monolithic).
// With `rec_tiles` the producers also leave the colour head's inputs behind, one block per tile in the colour head's
// own operand layouts (RecTile<V>, tc_heads.cuh; consumed by k3_color_tiles.cu with one bulk copy per tile): the
// per-view features are gathered once per point and frame.  Without it nothing is written per point but σ / α
// (k3_color_ws.cu then gathers the colour head's inputs again for the survivors of the progressive step).
#include <stdlib.h>
#include "tc_heads.cuh"

namespace gpnerf {

namespace ws {
constexpr int kProdWarps = 16, kChains = 2, kStages = 3;
constexpr int kPasses = 16 / kProdWarps;          // passes of 8 points per producer warp and tile
constexpr int kThreads = (kProdWarps + 4 * kChains + 1) * 32;        // 800
constexpr int kMmaWarp = kProdWarps + 4 * kChains;

constexpr uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }
struct Smem {
  static constexpr uint32_t IMG = 0;
  static constexpr uint32_t STAGE0 = align_up(DenImg::BYTES, 1024);
  static constexpr uint32_t A0 = 0;                       // [128 x 128] fp16, 2 K blocks of [128 x 64] SWIZZLE_128B
  static constexpr uint32_t G64 = 32768;                  // [128 x 64] bf16 (mean_feat | var_feat), SWIZZLE_128B
  static constexpr uint32_t TAIL = G64 + 16384;           // [128 x 16] bf16 (mean/var rgb, 1, 1 | 0 x 8), core-matrix layout
  static constexpr uint32_t STAGE_BYTES = TAIL + 4096;    // 52 KB: every stage stays 1024-byte aligned
  static constexpr uint32_t ONES = STAGE0 + kStages * STAGE_BYTES;   // [128 x 16] constant (…, 1, 1 | 0 x 8): bias rows
  static constexpr uint32_t PLAN = ONES + 4096;           // kProdWarps x 8 entries x (8 points + 1 pad) x 16 bytes
  static constexpr uint32_t NV = PLAN + kProdWarps * 1280;          // kStages x 128 bytes: valid views per row
  static constexpr uint32_t MISC = NV + kStages * 128;
  static constexpr uint32_t BYTES = MISC + 256;
};
static_assert(Smem::BYTES + 1024 + 1024 <= 227 * 1024, "one CTA per SM");
constexpr uint32_t kTailSbo = op_sbo(16);
}  // namespace ws

__device__ __forceinline__ uint4 ws_pack8(const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  return q;
}
__device__ __forceinline__ __half2 ws_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t ws_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ void ws_hfma8(__half2 (&acc)[4], const uint4& q, __half2 w) {
  acc[0] = __hfma2(ws_h2(q.x), w, acc[0]);
  acc[1] = __hfma2(ws_h2(q.y), w, acc[1]);
  acc[2] = __hfma2(ws_h2(q.z), w, acc[2]);
  acc[3] = __hfma2(ws_h2(q.w), w, acc[3]);
}
// Lane L = 4·i + sub holds chunk (c0 + sub) of row (row8 + i), i = 0..7.  After an 8x4 → 4x8 lane transpose every
// quarter-warp stores one chunk column of 8 consecutive rows: with the swizzle those are 8 different 16-byte slots
// of 8 different 128-byte rows – no bank conflicts.
__device__ __forceinline__ void ws_st_rows8_sw(uint8_t* kblock, int row8, int c0, uint4 v, int lane) {
  const int src = (lane & 7) * 4 + (lane >> 3);
  v.x = __shfl_sync(0xffffffffu, v.x, src);
  v.y = __shfl_sync(0xffffffffu, v.y, src);
  v.z = __shfl_sync(0xffffffffu, v.z, src);
  v.w = __shfl_sync(0xffffffffu, v.w, src);
  *reinterpret_cast<uint4*>(kblock + sw128_off(row8 + (lane & 7), c0 + (lane >> 3))) = v;
}

// Wait for a barrier phase.  Every poll is a shared-memory access that competes with the gathers for the L1 data
// pipe (the first version of this kernel spent a third of its instructions polling): back off between polls.
// Bounded, so that a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void ws_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test(bar, parity)) {
    __nanosleep(100);
    if (++spins > 4000000u) __trap();
  }
}

template <int V>
__global__ void __launch_bounds__(ws::kThreads, 1) gather_density_ws(FusedArgs a, const __grid_constant__ gpnerf_frame_t fparam) {
  using namespace ws;
  GPNERF_LOAD_FRAME(fparam)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* img = smem + Smem::IMG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::MISC);
  uint64_t* bar_w = bars;                    // weights landed
  uint64_t* full = bars + 1;                 // [kStages] producers → MMA issuer (+ epilogues)
  uint64_t* empty = full + kStages;          // [kStages] tcgen05.commit → producers
  uint64_t* m2e = empty + kStages;           // [kChains] tcgen05.commit → epilogue group
  uint64_t* e2m = m2e + kChains;             // [kChains] epilogue group → MMA issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e2m + kChains);
  float* xf = reinterpret_cast<float*>(smem + Smem::MISC + 128);      // 12 floats: u = A·p + B
  const float* fl = reinterpret_cast<const float*>(img + DenImg::F32);
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
    fence_mbar_init();
    mbar_arrive_expect_tx(bar_w, DenImg::BYTES);
    bulk_g2s(img, a.image, DenImg::BYTES, bar_w);
    // normalised volume coordinate u_c = ((p − Th)·R[:,c] − bmin_c) / (voxel_c · out_sh_c), c = x,y,z
    for (int c = 0; c < 3; ++c) {
      const double scale = 1.0 / ((double)f.voxel_size[c] * (double)f.out_sh[2 - c]);
      double b = -(double)f.bounds_min[c];
      for (int k = 0; k < 3; ++k) {
        xf[c * 4 + k] = (float)((double)f.R[k * 3 + c] * scale);
        b -= (double)f.Th[k] * (double)f.R[k * 3 + c];
      }
      xf[c * 4 + 3] = (float)(b * scale);
    }
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  // constants of the operand tiles: the bias rows (…, 1, 1 | 0 x 8) and the zero half of every stage's tail tile
  if (tid < 128) {
    uint8_t* o = smem + Smem::ONES;
    *reinterpret_cast<uint4*>(o + chunk_off(tid, 0, kTailSbo)) = make_uint4(0u, 0u, 0u, 0x3F803F80u);   // bf16 1.0 in columns 6, 7
    *reinterpret_cast<uint4*>(o + chunk_off(tid, 1, kTailSbo)) = make_uint4(0u, 0u, 0u, 0u);
    for (int s = 0; s < kStages; ++s)
      *reinterpret_cast<uint4*>(smem + Smem::STAGE0 + s * Smem::STAGE_BYTES + Smem::TAIL + chunk_off(tid, 1, kTailSbo)) =
          make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n = __ldg(a.counters + GPNERF_CNT_P1);
  const int n_tiles = (n + 127) / 128;
  const int G = gridDim.x;

  if (warp < kProdWarps) {
    // =========================================================== producers
    const int grp = lane >> 2, sub = lane & 3;
    // plan scratch of the warp, entry-major: entry e of point g at e·9 + g (16-byte units; 8 points read one entry as
    // 128 contiguous bytes, the odd stride spreads the writes over the banks)
    uint4* plan_base = reinterpret_cast<uint4*>(smem + Smem::PLAN + warp * 1280) + grp;
#define plan(e) plan_base[(e) * 9]
    const float ox = __ldg(a.rays_o), oy = __ldg(a.rays_o + 1), oz = __ldg(a.rays_o + 2);
    const int S = f.n_samples;
    const float sfx = (float)(f.feat_w - 1) / (float)(f.src_w - 1), sfy = (float)(f.feat_h - 1) / (float)(f.src_h - 1);
    const float wm1 = (float)(f.src_w - 1), hm1 = (float)(f.src_h - 1);
    const int img_stride_p = (f.src_h + 2) * (f.src_w + 2);          // padded image, float4 units
    const int map_stride_q = (f.feat_h + 2) * (f.feat_w + 2) * 4;    // padded map, uint4 units
    using RT = RecTile<V>;
    auto fetch_q = [&](long long first_row, int r) -> int {
      return (first_row + r < n) ? __ldg(a.valid + first_row + r) : -1;
    };
    // position of this lane's point in pass j of the tile: fetched one tile ahead (valid → (ray, z) → p is a
    // chain of two dependent global loads)
    float px[kPasses], py[kPasses], pz[kPasses];
    int qn[kPasses];
#pragma unroll
    for (int j = 0; j < kPasses; ++j) {
      const int q = fetch_q((long long)blockIdx.x * 128, (warp * kPasses + j) * 8 + grp);
      px[j] = py[j] = pz[j] = 0.f;
      if (q >= 0) {
        const int ray = q / S;
        const float z = __ldg(a.z_vals + q);
        px[j] = fmaf(__ldg(a.rays_d + ray * 3 + 0), z, ox);
        py[j] = fmaf(__ldg(a.rays_d + ray * 3 + 1), z, oy);
        pz[j] = fmaf(__ldg(a.rays_d + ray * 3 + 2), z, oz);
      }
    }
    mbar_wait(bar_w, 0);          // (the producers do not read the weights; keeps the prologue ordered)
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += G, ++it) {
      const long long first = (long long)tile * 128;
      const int n_valid = min(128, n - (int)first);
      const int s = it % kStages;
      uint8_t* stage = smem + Smem::STAGE0 + s * Smem::STAGE_BYTES;
      // colour-stage record of this tile (tc_heads.cuh): written with streaming stores, read back by one bulk copy
      uint8_t* rt = a.rec_tiles ? a.rec_tiles + (size_t)tile * RT::BYTES : nullptr;
#pragma unroll
      for (int j = 0; j < kPasses; ++j) qn[j] = fetch_q((long long)(tile + G) * 128, (warp * kPasses + j) * 8 + grp);
      // the stage is free once the second GEMM of its previous tenant has completed
      ws_wait(empty + s, ((it / kStages) & 1) ^ 1);
#pragma unroll 1
      for (int j = 0; j < ((a.debug & 1) ? 0 : kPasses); ++j) {
        const int r = (warp * kPasses + j) * 8 + grp;
        const bool ok = r < n_valid;
        const float ppx = px[j], ppy = py[j], ppz = pz[j];
        // ---- plan: lane `sub` plans level `sub` and (sub < V) view `sub` of the point
        {
          const float ux = fmaf(xf[0], ppx, fmaf(xf[1], ppy, fmaf(xf[2], ppz, xf[3])));
          const float uy = fmaf(xf[4], ppx, fmaf(xf[5], ppy, fmaf(xf[6], ppz, xf[7])));
          const float uz = fmaf(xf[8], ppx, fmaf(xf[9], ppy, fmaf(xf[10], ppz, xf[11])));
          // the level is stored inside a one-voxel zero border: after clamping the continuous index to
          // [-1, size] all 8 corners are addressable and out-of-range corners read zeros (= zeros padding)
          const int D = f.level_dims[sub][0], H = f.level_dims[sub][1], W = f.level_dims[sub][2];
          const float ix = fminf(fmaxf(ux * (float)(W - 1), -1.0f), (float)W);
          const float iy = fminf(fmaxf(uy * (float)(H - 1), -1.0f), (float)H);
          const float iz = fminf(fmaxf(uz * (float)(D - 1), -1.0f), (float)D);
          const int x0 = min((int)floorf(ix), W - 1), y0 = min((int)floorf(iy), H - 1), z0 = min((int)floorf(iz), D - 1);
          const int dy = (W + 2) * 4, dz = (H + 2) * dy;               // strides in uint4 (16 B) units
          uint4 e;
          e.x = ok ? (uint32_t)((z0 + 1) * dz + (y0 + 1) * dy + (x0 + 1) * 4) : 0u;
          e.y = __float_as_uint(ix - (float)x0);
          e.z = __float_as_uint(iy - (float)y0);
          e.w = __float_as_uint(iz - (float)z0);
          plan(sub) = e;
        }
        int nv = 0;
        if (sub < V) {
          const float* KE = f.src_KE[sub];
          const float qx = fmaf(KE[0], ppx, fmaf(KE[1], ppy, fmaf(KE[2], ppz, KE[3])));
          const float qy = fmaf(KE[4], ppx, fmaf(KE[5], ppy, fmaf(KE[6], ppz, KE[7])));
          const float qz = fmaf(KE[8], ppx, fmaf(KE[9], ppy, fmaf(KE[10], ppz, KE[11])));
          const float inv = 1.0f / qz;
          const float ux2 = qx * inv, uy2 = qy * inv;
          const bool front = f.neg_ray ? (qz < 0.0f) : (qz > 0.0f);
          const bool inbv = (ux2 <= wm1) && (ux2 >= 0.0f) && (uy2 <= hm1) && (uy2 >= 0.0f);
          nv = (front && inbv) ? 1 : 0;
          uint4 e;
          {  // feature map tap (align_corners: pixel p ↦ p·(Wm−1)/(w−1)); map stored inside a zero border
            const float ix = fminf(fmaxf(ux2 * sfx, -1.0f), (float)f.feat_w);
            const float iy = fminf(fmaxf(uy2 * sfy, -1.0f), (float)f.feat_h);
            const int x0 = min((int)floorf(ix), f.feat_w - 1), y0 = min((int)floorf(iy), f.feat_h - 1);
            e.x = ok ? (uint32_t)(sub * map_stride_q + (y0 + 1) * (f.feat_w + 2) * 4 + (x0 + 1) * 4) : 0u;
            e.z = pack_f16x2(ix - (float)x0, iy - (float)y0);
          }
          {  // RGB tap at the image's own resolution
            const float ix = fminf(fmaxf(ux2, -1.0f), (float)f.src_w), iy = fminf(fmaxf(uy2, -1.0f), (float)f.src_h);
            const int x0 = min((int)floorf(ix), f.src_w - 1), y0 = min((int)floorf(iy), f.src_h - 1);
            e.y = ok ? (uint32_t)(sub * img_stride_p + (y0 + 1) * (f.src_w + 2) + (x0 + 1)) : 0u;
            e.w = pack_f16x2(ix - (float)x0, iy - (float)y0);
          }
          plan(4 + sub) = e;
        }
        nv += __shfl_xor_sync(0xffffffffu, nv, 1);
        nv += __shfl_xor_sync(0xffffffffu, nv, 2);
        if (sub == 0) smem[Smem::NV + s * 128 + r] = (uint8_t)nv;
        __syncwarp();
        // ---- 4-level trilinear gather → A0 chunk (level*4 + sub), fp16
#pragma unroll
        for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
          const uint4 e = plan(l);
          const int W = f.level_dims[l][2], H = f.level_dims[l][1];
          const int dy = (W + 2) * 4, dz = (H + 2) * dy;
          const uint4* p0 = reinterpret_cast<const uint4*>(a.lv[l]) + (e.x + sub);
          uint4 q[8];
          q[0] = __ldg(p0);
          q[1] = __ldg(p0 + 4);
          q[2] = __ldg(p0 + dy);
          q[3] = __ldg(p0 + dy + 4);
          q[4] = __ldg(p0 + dz);
          q[5] = __ldg(p0 + dz + 4);
          q[6] = __ldg(p0 + dz + dy);
          q[7] = __ldg(p0 + dz + dy + 4);
          const float wx1 = __uint_as_float(e.y), wy1 = __uint_as_float(e.z), wz1 = __uint_as_float(e.w);
          const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1, wz0 = 1.0f - wz1;
          const float w00 = wy0 * wz0, w10 = wy1 * wz0, w01 = wy0 * wz1, w11 = wy1 * wz1;
          __half2 acc[4] = {__float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f)};
          ws_hfma8(acc, q[0], __float2half2_rn(wx0 * w00));
          ws_hfma8(acc, q[1], __float2half2_rn(wx1 * w00));
          ws_hfma8(acc, q[2], __float2half2_rn(wx0 * w10));
          ws_hfma8(acc, q[3], __float2half2_rn(wx1 * w10));
          ws_hfma8(acc, q[4], __float2half2_rn(wx0 * w01));
          ws_hfma8(acc, q[5], __float2half2_rn(wx1 * w01));
          ws_hfma8(acc, q[6], __float2half2_rn(wx0 * w11));
          ws_hfma8(acc, q[7], __float2half2_rn(wx1 * w11));
          ws_st_rows8_sw(stage + Smem::A0 + (l >> 1) * 16384, r & ~7, (l & 1) * 4,
                         make_uint4(ws_u32(acc[0]), ws_u32(acc[1]), ws_u32(acc[2]), ws_u32(acc[3])), lane);
        }
        // ---- V source views: feature taps by all 4 lanes (8 channels each), mean / variance
        float fv[V][8];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const uint4 e = plan(4 + v);
          const int dy = (f.feat_w + 2) * 4;
          const uint4* p0 = reinterpret_cast<const uint4*>(a.feat) + (e.x + sub);
          uint4 q[4];
          q[0] = __ldg(p0);
          q[1] = __ldg(p0 + 4);
          q[2] = __ldg(p0 + dy);
          q[3] = __ldg(p0 + dy + 4);
          const float2 wf = __half22float2(ws_h2(e.z));
          const float wx = wf.x, wy = wf.y;
          __half2 acc[4] = {__float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f)};
          ws_hfma8(acc, q[0], __float2half2_rn((1.0f - wx) * (1.0f - wy)));
          ws_hfma8(acc, q[1], __float2half2_rn(wx * (1.0f - wy)));
          ws_hfma8(acc, q[2], __float2half2_rn((1.0f - wx) * wy));
          ws_hfma8(acc, q[3], __float2half2_rn(wx * wy));
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float2 t = __half22float2(acc[jj]);
            fv[v][2 * jj] = t.x;
            fv[v][2 * jj + 1] = t.y;
          }
          if (rt != nullptr) __stcs(reinterpret_cast<uint4*>(rt + RT::ff_off(v, r, sub)), ws_pack8(fv[v]));
        }
        // ---- RGB taps: lane `sub` takes view `sub` (fp32 images, fp32 arithmetic)
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (sub < V) {
          const uint4 e = plan(4 + sub);
          const int dy = f.src_w + 2;
          const float4* p0 = a.rgbx + e.y;
          const float4 t0 = __ldg(p0), t1 = __ldg(p0 + 1), t2 = __ldg(p0 + dy), t3 = __ldg(p0 + dy + 1);
          const float2 wr = __half22float2(ws_h2(e.w));
          const float wx = wr.x, wy = wr.y;
          const float w0 = (1.0f - wx) * (1.0f - wy), w1 = wx * (1.0f - wy), w2 = (1.0f - wx) * wy, w3 = wx * wy;
          c0 = fmaf(t3.x, w3, fmaf(t2.x, w2, fmaf(t1.x, w1, t0.x * w0)));
          c1 = fmaf(t3.y, w3, fmaf(t2.y, w2, fmaf(t1.y, w1, t0.y * w0)));
          c2 = fmaf(t3.z, w3, fmaf(t2.z, w2, fmaf(t1.z, w1, t0.z * w0)));
          if (a.rgb_in != nullptr && ok) {
            float* o = a.rgb_in + ((first + r) * V + sub) * 3;
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
          const uint4 qm = ws_pack8(mean), qv = ws_pack8(var);
          ws_st_rows8_sw(stage + Smem::G64, r & ~7, 0, qm, lane);
          ws_st_rows8_sw(stage + Smem::G64, r & ~7, 4, qv, lane);
          if (rt != nullptr) {
            __stcs(reinterpret_cast<uint4*>(rt + RT::G64 + sw128_off(r, sub)), qm);
            __stcs(reinterpret_cast<uint4*>(rt + RT::G64 + sw128_off(r, 4 + sub)), qv);
          }
          // RGB mean / variance over the views: lane 0 of the point collects (r,g,b) of the V views from lanes 0..V-1
          float rgbv[12];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            rgbv[3 * v + 0] = __shfl_sync(0xffffffffu, c0, (lane & ~3) + v);
            rgbv[3 * v + 1] = __shfl_sync(0xffffffffu, c1, (lane & ~3) + v);
            rgbv[3 * v + 2] = __shfl_sync(0xffffffffu, c2, (lane & ~3) + v);
          }
          if (sub == 0) {
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
            // columns 70, 71 of the [mean|var] operand are the constant 1.0 that carries the layer's bias
            const float t[8] = {m0, m1, m2, s0 * inv_v, s1 * inv_v, s2 * inv_v, 1.0f, 1.0f};
            const uint4 qc = ws_pack8(t);
            *reinterpret_cast<uint4*>(stage + Smem::TAIL + chunk_off(r, 0, kTailSbo)) = qc;
            if (rt != nullptr) {
              __stcs(reinterpret_cast<uint4*>(rt + RT::TAIL + chunk_off(r, 0, kTailSbo)), qc);
              // per-view RGB block of base_fc.0's input: column 3v + c = channel c of view v
              const float u0[8] = {rgbv[0], rgbv[1], rgbv[2], rgbv[3], rgbv[4], rgbv[5], rgbv[6], rgbv[7]};
              __stcs(reinterpret_cast<uint4*>(rt + RT::RGBS + chunk_off(r, 0, kTailSbo)), ws_pack8(u0));
              if (V > 2) {
                const float u1[8] = {rgbv[8], V > 3 ? rgbv[9] : 0.f, V > 3 ? rgbv[10] : 0.f, V > 3 ? rgbv[11] : 0.f, 0.f, 0.f, 0.f, 0.f};
                __stcs(reinterpret_cast<uint4*>(rt + RT::RGBS + chunk_off(r, 1, kTailSbo)), ws_pack8(u1));
              }
            }
          }
        }
        __syncwarp();               // the next pass rewrites this warp's plan scratch
      }
      // hand the stage to the tensor core: every lane publishes its own stores to the async proxy, then one
      // arrival per warp
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full + s);
      // next tile's positions (their `valid` entries were requested at the top of this tile)
#pragma unroll
      for (int j = 0; j < kPasses; ++j) {
        px[j] = py[j] = pz[j] = 0.f;
        if (qn[j] >= 0) {
          const int ray = qn[j] / S;
          const float z = __ldg(a.z_vals + qn[j]);
          px[j] = fmaf(__ldg(a.rays_d + ray * 3 + 0), z, ox);
          py[j] = fmaf(__ldg(a.rays_d + ray * 3 + 1), z, oy);
          pz[j] = fmaf(__ldg(a.rays_d + ray * 3 + 2), z, oz);
        }
      }
    }
#undef plan
  } else if (warp == kMmaWarp) {
    // =========================================================== MMA issuer (one thread)
    if (lane == 0) {
      mbar_wait(bar_w, 0);
      const uint32_t wimg = smem_u32(img), ones = smem_u32(smem + Smem::ONES);
      const uint64_t ones_d = make_smem_desc(ones, kLBO, kTailSbo);
      const uint32_t id64h = make_idesc(128, 64, kFmtF16), id64 = make_idesc_bf16(128, 64);
      const uint32_t id32 = make_idesc_bf16(128, 32), id16 = make_idesc_bf16(128, 16);
      auto bdesc = [&](uint32_t w_off, int k16, int Kp) {
        return make_smem_desc(wimg + w_off + k16 * 2 * kLBO, kLBO, op_sbo(Kp));
      };
      int step[kChains] = {0, 0}, li[kChains] = {0, 1};
      uint32_t e2m_ph[kChains] = {0, 0};
      bool done[kChains] = {false, false};
      uint32_t idle = 0;
      int next_start = 0;          // tiles start strictly in order (a parity wait tells only consecutive phases apart)
      if (a.debug & 2) {           // profiling experiment: no head at all, stages handed straight back
        for (int i = 0; blockIdx.x + (long long)i * G < n_tiles; ++i) {
          ws_wait(full + i % kStages, (i / kStages) & 1);
          mbar_arrive(empty + i % kStages);
        }
        done[0] = done[1] = true;
      }
      while (!(done[0] && done[1])) {
        bool progressed = false;
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
          if (done[c]) continue;
          const int i = li[c];
          if (blockIdx.x + (long long)i * G >= n_tiles) {
            done[c] = true;
            progressed = true;
            continue;
          }
          const int s = i % kStages;
          const uint32_t stage = smem_u32(smem + Smem::STAGE0 + s * Smem::STAGE_BYTES);
          const uint32_t acc = tmem + c * 128, act = tmem + c * 128 + 64;
          if (step[c] == 0) {
            // chain free (the final epilogue of its previous tile has read the accumulator) and stage full
            if (i != next_start) continue;
            if (i >= kChains && !mbar_test(e2m + c, e2m_ph[c])) continue;
            if (!mbar_test(full + s, (i / kStages) & 1)) continue;
            if (i >= kChains) e2m_ph[c] ^= 1u;
            tc_fence_after();
            // sigmahead.out_geometry_fc: [128 x 128] fp16 · Wgᵀ → 64
            for (int k16 = 0; k16 < 8; ++k16)
              umma_bf16(acc, make_smem_desc_sw128(stage + Smem::A0 + (k16 >> 2) * 16384 + (k16 & 3) * 32),
                        bdesc(DenImg::Wg, k16, 144), id64h, k16 > 0);
            umma_bf16(acc, ones_d, bdesc(DenImg::Wg, 8, 144), id64, 1u);
            umma_commit(m2e + c);
            step[c] = 1;
            ++next_start;
          } else {
            if (!mbar_test(e2m + c, e2m_ph[c])) continue;
            e2m_ph[c] ^= 1u;
            tc_fence_after();
            if (step[c] == 1) {
              // out_geometry_fc.0: [sigma_feat (TMEM) | mean,var (stage)] 144 → 64
              for (int k16 = 0; k16 < 4; ++k16) umma_ts(acc, act + k16 * 8, bdesc(DenImg::W0, k16, 144), id64, k16 > 0);
              for (int k16 = 0; k16 < 4; ++k16)
                umma_bf16(acc, make_smem_desc_sw128(stage + Smem::G64 + k16 * 32), bdesc(DenImg::W0, 4 + k16, 144), id64, 1u);
              umma_bf16(acc, make_smem_desc(stage + Smem::TAIL, kLBO, kTailSbo), bdesc(DenImg::W0, 8, 144), id64, 1u);
              umma_commit(m2e + c);
              umma_commit(empty + s);           // both GEMMs that read the stage have completed when this arrives
            } else if (step[c] == 2) {
              // .2: 64 → 32
              for (int k16 = 0; k16 < 4; ++k16) umma_ts(acc, act + k16 * 8, bdesc(DenImg::W1, k16, 80), id32, k16 > 0);
              umma_bf16(acc, ones_d, bdesc(DenImg::W1, 4, 80), id32, 1u);
              umma_commit(m2e + c);
            } else {
              // .4: 32 → 16
              for (int k16 = 0; k16 < 2; ++k16) umma_ts(acc, act + k16 * 8, bdesc(DenImg::W2, k16, 48), id16, k16 > 0);
              umma_bf16(acc, ones_d, bdesc(DenImg::W2, 2, 48), id16, 1u);
              umma_commit(m2e + c);
            }
            step[c] = (step[c] + 1) & 3;
            if (step[c] == 0) li[c] += kChains;
          }
          progressed = true;
        }
        if (progressed) {
          idle = 0;
        } else {
          __nanosleep(100);
          if (++idle > 8000000u) __trap();        // a protocol bug must surface as a trap, not as a hung GPU
        }
      }
    }
  } else {
    // =========================================================== epilogues
    const int c = (warp - kProdWarps) >> 2;
    const int row = ((warp - kProdWarps) & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(((warp - kProdWarps) & 3) * 32) << 16;
    const uint32_t acc = tmem + c * 128 + lane_sel, act = tmem + c * 128 + 64 + lane_sel;
    mbar_wait(bar_w, 0);
    uint32_t ph = 0;
    // one warp of the group polls the mbarrier (every poll is a shared-memory access on the L1 data pipe the
    // gathers are bound by); the other three block on a named barrier, which costs nothing while they wait
    const bool leader = ((warp - kProdWarps) & 3) == 0;
    auto wait_acc = [&]() {
      if (leader) ws_wait(m2e + c, ph);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + c) : "memory");
      ph ^= 1u;
      tc_fence_after();
    };
    auto publish = [&]() {          // the activations are in TMEM / the accumulator has been read: next GEMM may go
      tc_fence_before();
      mbar_arrive(e2m + c);
    };
    // accumulator columns [c0, c0+32) → scaled ELU → 16 packed bf16 pairs
    auto epi32 = [&](int c0, uint32_t (&pk)[16]) {
      uint32_t r[32];
      tmem_ld32(acc + c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        pk[j] = pack_bf16x2(elu_scaled(__uint_as_float(r[2 * j])), elu_scaled(__uint_as_float(r[2 * j + 1])));
    };
    for (int i = c; blockIdx.x + (long long)i * G < n_tiles && !(a.debug & 2); i += kChains) {
      const long long first = ((long long)blockIdx.x + (long long)i * G) * 128;
      const int n_valid = min(128, n - (int)first);
      const int s = i % kStages;
      uint32_t pk[16];
      // ---- sigma_feat
      wait_acc();
      // (full[s] has completed – the GEMM that just finished was issued after it – and the MMA issuer's acquire of it
      // is ordered before the commit the leader observed: the producers' plain stores to NV are visible)
      const int nv = smem[Smem::NV + s * 128 + row];
      epi32(0, pk);
      tmem_st16(act, pk);
      epi32(32, pk);
      tmem_st16(act + 16, pk);
      tmem_wait_st();
      publish();
      // ---- out_geometry_fc.0
      wait_acc();
      epi32(0, pk);
      tmem_st16(act, pk);
      epi32(32, pk);
      tmem_st16(act + 16, pk);
      tmem_wait_st();
      publish();
      // ---- .2
      wait_acc();
      epi32(0, pk);
      tmem_st16(act, pk);
      tmem_wait_st();
      publish();
      // ---- .4 ; .6: 16 → 1 on CUDA cores, ReLU, no-valid-view fill, progressive-step test
      wait_acc();
      uint32_t r16[16];
      tmem_ld16(acc, r16);
      tmem_wait_ld();
      publish();                                   // chain free: its next tile's first GEMM may overwrite the accumulator
      float sg = fl[DenImg::b3];
#pragma unroll
      for (int k = 0; k < 16; ++k) sg = fmaf(elu_scaled(__uint_as_float(r16[k])), fl[DenImg::w3 + k], sg);
      sg = fmaxf(sg, 0.0f);
      sg = (nv < 1) ? 0.0f : sg;
      if (row < n_valid) a.sigma[first + row] = sg;
      if (a.alpha != nullptr) {
        // demo_render.py:312-317 (as alpha_flags computes it): the warp holds 32 consecutive points = one flag word
        const float al = xsub(1.0f, expf(-sg));
        const bool keep = row < n_valid && al > 1e-14f;
        if (row < n_valid) a.alpha[first + row] = al;
        const unsigned bits = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) a.alpha_words[(first + row) >> 5] = bits;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace gpnerf

using namespace gpnerf;

template <int V>
static int launch_ws(const FusedArgs& a, const gpnerf_frame_t* f, int n_points_max, cudaStream_t st) {
  static bool attr_set = false;
  constexpr uint32_t bytes = ws::Smem::BYTES + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gather_density_ws<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
      set_error("gather_density_ws smem attribute", e);
      return GPNERF_E_CUDA;
    }
    attr_set = true;
  }
  const int tiles = (n_points_max + 127) / 128;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gather_density_ws<V><<<grid, ws::kThreads, bytes, st>>>(a, *f);
  return check_launch("k23_gather_density_ws");
}

int launch_fused_ws(const FusedArgs& a, const gpnerf_frame_t* f, int n_points_max, cudaStream_t st) {
  switch (f->n_views) {
    case 1: return launch_ws<1>(a, f, n_points_max, st);
    case 2: return launch_ws<2>(a, f, n_points_max, st);
    case 3: return launch_ws<3>(a, f, n_points_max, st);
    case 4: return launch_ws<4>(a, f, n_points_max, st);
    default:
      set_error("fused tcgen05 path supports 1..4 source views", cudaSuccess);
      return GPNERF_E_UNSUPPORTED;
  }
}
