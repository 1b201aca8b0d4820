// K2+K3 fused (tensor-core path): for each tile of 128 surviving sample points,
// gather the 4-level geometry volume and the V source views, aggregate
// mean/variance, and run the density head – without the gathered features ever
// leaving the SM.
//
//   plan phase    one thread per (point, {levels | views}): sample position, world→grid
//                 transform, per-level corner offset + fractions, per-view projection,
//                 in-front/in-bounds test, tap offsets + fractions.  16 bytes per level and
//                 per view, parked in the (still unused) sigma_feat columns of the A1 tile.
//   gather phase  (SparseConvNet.py:111-122, BaseRender.py:283-363, trainhead.py:20-24):
//                 4 lanes per point, 8 channels (one 16-byte load) per lane and corner.
//                 Volumes and feature maps are stored channel-last in FP16 inside a zero
//                 border; the interpolation is HFMA2 on the packed pairs (no unpacking,
//                 11-bit mantissa: finer than the bf16 the result is rounded to anyway).
//                 Each lane's result is exactly one 16-byte chunk of a UMMA A operand,
//                 stored straight into shared memory.  RGB taps: lane v takes view v.
//   density phase (trainhead.py:39-41, 102-110, 133-137): four tcgen05.mma rounds
//                 128→64 (fp16 × fp16), 144→64, 64→32, 32→16 (bf16) with TMEM accumulators,
//                 biases folded into the GEMMs, activations kept times log2(e)
//                 (tc_common.cuh, elu_scaled); 16→1 + ReLU + no-valid-view fill on CUDA cores.
// By-product: one 16·(9+5V)-byte bf16 record per point ([mean|var] and the V
// per-view feature rows, already in operand order) for the colour head, which
// only the points that survive the progressive step will read back.
//
// Index arithmetic is one affine map per point (FMAs): unlike the fp32 path this
// kernel does not reproduce the reference's rounding sequence – the integer
// results (which points exist) were fixed upstream by the exact K1/K2 kernels.
// 256 threads, 2 CTAs/SM: one CTA's gather overlaps the other's MMA/epilogue.
#include <stdlib.h>
#include <string.h>
#include "tc_heads.cuh"

namespace gpnerf {

struct FusedSmem {
  static constexpr uint32_t IMG = 0;
  static constexpr uint32_t A0 = ((DenImg::BYTES + 127) / 128) * 128;   // [128 x 128] fp16; later [128 x 64] + [128 x 32]
  static constexpr uint32_t A1 = A0 + op_bytes(128, 128);               // [128 x 144] = sigma_feat | G ; plan in cols 0..63
  static constexpr uint32_t MISC = A1 + op_bytes(128, 144);             // barriers, tmem slot, flags, transform
  static constexpr uint32_t BYTES = MISC + 512;
};
static_assert(2 * (FusedSmem::BYTES + 1024) <= 228 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  return q;
}
__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 dup_h2(float w) { return __float2half2_rn(w); }
// acc += q (8 fp16 channels) * w
__device__ __forceinline__ void hfma8(__half2 (&acc)[4], const uint4& q, __half2 w) {
  acc[0] = __hfma2(as_h2(q.x), w, acc[0]);
  acc[1] = __hfma2(as_h2(q.y), w, acc[1]);
  acc[2] = __hfma2(as_h2(q.z), w, acc[2]);
  acc[3] = __hfma2(as_h2(q.w), w, acc[3]);
}

// Operand-tile store for the gather lanes.  Lane L = 4·i + sub holds chunk (kc0 + sub) of row (row8 + i),
// i = 0..7.  All K chunks of one row fall into the same 4 banks (LBO = 128 B), so storing from this lane
// order is a 4-way conflict in every quarter-warp; after an 8x4 → 4x8 lane transpose (4 shuffles) every
// quarter-warp writes one chunk column of 8 consecutive rows = 128 contiguous bytes.
__device__ __forceinline__ void st_chunk_rows8(uint8_t* tile, uint32_t sbo, int row8, int kc0, uint4 v, int lane) {
  const int src = (lane & 7) * 4 + (lane >> 3);
  v.x = __shfl_sync(0xffffffffu, v.x, src);
  v.y = __shfl_sync(0xffffffffu, v.y, src);
  v.z = __shfl_sync(0xffffffffu, v.z, src);
  v.w = __shfl_sync(0xffffffffu, v.w, src);
  *reinterpret_cast<uint4*>(tile + chunk_off(row8 + (lane & 7), kc0 + (lane >> 3), sbo)) = v;
}

template <int V>
__global__ void __launch_bounds__(256, 2) gather_density_tc(FusedArgs a, const __grid_constant__ gpnerf_frame_t fparam) {
  GPNERF_LOAD_FRAME(fparam)
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* img = smem + FusedSmem::IMG;
  uint8_t* A0 = smem + FusedSmem::A0;
  uint8_t* A1 = smem + FusedSmem::A1;
  uint8_t* A2 = A0 + op_bytes(128, 64);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + FusedSmem::MISC);
  uint64_t* bar_m = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  float* xf = reinterpret_cast<float*>(smem + FusedSmem::MISC + 32);      // 12 floats: u = A·p + B
  uint8_t* nvalid = smem + FusedSmem::MISC + 128;                         // 128 flags
  const float* fl = reinterpret_cast<const float*>(img + DenImg::F32);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
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
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  // G chunk 9 (operand columns 136..143) is padding: zero once, nothing overwrites it
  if (tid < 128) *reinterpret_cast<uint4*>(A1 + chunk_off(tid, 17, op_sbo(144))) = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  mbar_wait(bar_w, 0);

  const uint32_t a0 = smem_u32(A0), a1 = smem_u32(A1), a2 = smem_u32(A2), wimg = smem_u32(img);
  const uint32_t ones = a1 + 8 * 2 * kLBO;          // last K block of A1: (mean/var rgb, 1, 1 | 0×8)
  const float ox = __ldg(a.rays_o), oy = __ldg(a.rays_o + 1), oz = __ldg(a.rays_o + 2);
  const int S = f.n_samples;
  const float sfx = (float)(f.feat_w - 1) / (float)(f.src_w - 1), sfy = (float)(f.feat_h - 1) / (float)(f.src_h - 1);
  const float wm1 = (float)(f.src_w - 1), hm1 = (float)(f.src_h - 1);
  const int img_stride_p = (f.src_h + 2) * (f.src_w + 2);          // padded image, float4 units
  const int map_stride_q = (f.feat_h + 2) * (f.feat_w + 2) * 4;    // padded map, uint4 units
  constexpr int RC = rec_chunks(V);
  constexpr uint32_t SBO1 = op_sbo(144);
  const int grp = tid >> 2, sub = tid & 3;
  const int row = tid & 127, half = tid >> 7;
  uint32_t phase = 0;
  const int n = __ldg(a.counters + GPNERF_CNT_P1);
  const int n_tiles = (n + 127) / 128;

  // Sample position of this thread's plan row, fetched one tile ahead: valid → (ray, z) → p is a chain of
  // two dependent global loads that would otherwise sit at the head of every tile.
  auto fetch_q = [&](long long first_row) -> int {
    return (first_row + row < n) ? __ldg(a.valid + first_row + row) : -1;
  };
  auto fetch_point = [&](int q, float& px, float& py, float& pz) {
    px = py = pz = 0.f;
    if (q >= 0) {
      const int ray = q / S;
      const float z = __ldg(a.z_vals + q);
      px = fmaf(__ldg(a.rays_d + ray * 3 + 0), z, ox);
      py = fmaf(__ldg(a.rays_d + ray * 3 + 1), z, oy);
      pz = fmaf(__ldg(a.rays_d + ray * 3 + 2), z, oz);
    }
  };
  float px, py, pz;
  fetch_point(fetch_q((long long)blockIdx.x * 128), px, py, pz);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long first = (long long)tile * 128;
    const int n_valid = min(128, n - (int)first);
    // =================== plan phase ===================
    // thread (row, half): half 0 plans the 4 volume levels of point `row`, half 1 its V views.
    // Plan entry = 16 bytes in chunk kc of the row's (unused until the first epilogue) sigma_feat
    // columns of A1: kc = level, or 4 + view.  Rows past the tile's end plan all-zero taps at offset 0
    // (the zero border), so the gather phase needs no row guard.
    {
      const bool ok = row < n_valid;
      uint8_t* prow = A1 + chunk_off(row, 0, SBO1);
      if (half == 0) {
        const float ux = fmaf(xf[0], px, fmaf(xf[1], py, fmaf(xf[2], pz, xf[3])));
        const float uy = fmaf(xf[4], px, fmaf(xf[5], py, fmaf(xf[6], pz, xf[7])));
        const float uz = fmaf(xf[8], px, fmaf(xf[9], py, fmaf(xf[10], pz, xf[11])));
#pragma unroll
        for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
          // the level is stored inside a one-voxel zero border: after clamping the continuous index to
          // [-1, size] all 8 corners are addressable and out-of-range corners read zeros (= zeros padding)
          const int D = f.level_dims[l][0], H = f.level_dims[l][1], W = f.level_dims[l][2];
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
          *reinterpret_cast<uint4*>(prow + l * kLBO) = e;
        }
      } else {
        int nv = 0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float* KE = f.src_KE[v];
          const float qx = fmaf(KE[0], px, fmaf(KE[1], py, fmaf(KE[2], pz, KE[3])));
          const float qy = fmaf(KE[4], px, fmaf(KE[5], py, fmaf(KE[6], pz, KE[7])));
          const float qz = fmaf(KE[8], px, fmaf(KE[9], py, fmaf(KE[10], pz, KE[11])));
          const float inv = 1.0f / qz;
          const float ux2 = qx * inv, uy2 = qy * inv;
          const bool front = f.neg_ray ? (qz < 0.0f) : (qz > 0.0f);
          const bool inbv = (ux2 <= wm1) && (ux2 >= 0.0f) && (uy2 <= hm1) && (uy2 >= 0.0f);
          nv += (front && inbv) ? 1 : 0;
          uint4 e;
          {  // feature map tap (align_corners: pixel p ↦ p·(Wm−1)/(w−1)); map stored inside a zero border
            const float ix = fminf(fmaxf(ux2 * sfx, -1.0f), (float)f.feat_w);
            const float iy = fminf(fmaxf(uy2 * sfy, -1.0f), (float)f.feat_h);
            const int x0 = min((int)floorf(ix), f.feat_w - 1), y0 = min((int)floorf(iy), f.feat_h - 1);
            e.x = ok ? (uint32_t)(v * map_stride_q + (y0 + 1) * (f.feat_w + 2) * 4 + (x0 + 1) * 4) : 0u;
            e.z = pack_f16x2(ix - (float)x0, iy - (float)y0);
          }
          {  // RGB tap at the image's own resolution
            const float ix = fminf(fmaxf(ux2, -1.0f), (float)f.src_w), iy = fminf(fmaxf(uy2, -1.0f), (float)f.src_h);
            const int x0 = min((int)floorf(ix), f.src_w - 1), y0 = min((int)floorf(iy), f.src_h - 1);
            e.y = ok ? (uint32_t)(v * img_stride_p + (y0 + 1) * (f.src_w + 2) + (x0 + 1)) : 0u;
            e.w = pack_f16x2(ix - (float)x0, iy - (float)y0);
          }
          *reinterpret_cast<uint4*>(prow + (4 + v) * kLBO) = e;
        }
        nvalid[row] = (uint8_t)nv;
      }
    }
    __syncthreads();
    const int q_next = fetch_q((long long)(tile + gridDim.x) * 128);     // lands during the gather phase
    // =================== gather phase ===================
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int r = pass * 64 + grp;
      const bool ok = r < n_valid;
      const uint8_t* prow = A1 + chunk_off(r, 0, SBO1);
      // ---- 4-level trilinear gather → A0 chunk (level*4 + sub), fp16
#pragma unroll
      for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
        const uint4 e = *reinterpret_cast<const uint4*>(prow + l * kLBO);
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
        __half2 acc[4] = {dup_h2(0.f), dup_h2(0.f), dup_h2(0.f), dup_h2(0.f)};
        hfma8(acc, q[0], dup_h2(wx0 * w00));
        hfma8(acc, q[1], dup_h2(wx1 * w00));
        hfma8(acc, q[2], dup_h2(wx0 * w10));
        hfma8(acc, q[3], dup_h2(wx1 * w10));
        hfma8(acc, q[4], dup_h2(wx0 * w01));
        hfma8(acc, q[5], dup_h2(wx1 * w01));
        hfma8(acc, q[6], dup_h2(wx0 * w11));
        hfma8(acc, q[7], dup_h2(wx1 * w11));
        st_chunk_rows8(A0, op_sbo(128), r & ~7, l * 4,
                       make_uint4(as_u32(acc[0]), as_u32(acc[1]), as_u32(acc[2]), as_u32(acc[3])), tid & 31);
      }
      // ---- V source views: feature taps by all 4 lanes (8 channels each), mean / variance
      float fv[V][8];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const uint4 e = *reinterpret_cast<const uint4*>(prow + (4 + v) * kLBO);
        const int dy = (f.feat_w + 2) * 4;
        const uint4* p0 = reinterpret_cast<const uint4*>(a.feat) + (e.x + sub);
        uint4 q[4];
        q[0] = __ldg(p0);
        q[1] = __ldg(p0 + 4);
        q[2] = __ldg(p0 + dy);
        q[3] = __ldg(p0 + dy + 4);
        const float2 wf = __half22float2(as_h2(e.z));
        const float wx = wf.x, wy = wf.y;
        __half2 acc[4] = {dup_h2(0.f), dup_h2(0.f), dup_h2(0.f), dup_h2(0.f)};
        hfma8(acc, q[0], dup_h2((1.0f - wx) * (1.0f - wy)));
        hfma8(acc, q[1], dup_h2(wx * (1.0f - wy)));
        hfma8(acc, q[2], dup_h2((1.0f - wx) * wy));
        hfma8(acc, q[3], dup_h2(wx * wy));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(acc[j]);
          fv[v][2 * j] = t.x;
          fv[v][2 * j + 1] = t.y;
        }
        if (ok) a.rec[(first + r) * RC + 9 + v * 5 + sub] = pack8(fv[v]);
      }
      // ---- RGB taps: lane `sub` takes view `sub` (fp32 images, fp32 arithmetic)
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (sub < V) {
        const uint4 e = *reinterpret_cast<const uint4*>(prow + (4 + sub) * kLBO);
        const int dy = f.src_w + 2;
        const float4* p0 = a.rgbx + e.y;
        const float4 t0 = __ldg(p0), t1 = __ldg(p0 + 1), t2 = __ldg(p0 + dy), t3 = __ldg(p0 + dy + 1);
        const float2 wr = __half22float2(as_h2(e.w));
        const float wx = wr.x, wy = wr.y;
        const float w0 = (1.0f - wx) * (1.0f - wy), w1 = wx * (1.0f - wy), w2 = (1.0f - wx) * wy, w3 = wx * wy;
        c0 = fmaf(t3.x, w3, fmaf(t2.x, w2, fmaf(t1.x, w1, t0.x * w0)));
        c1 = fmaf(t3.y, w3, fmaf(t2.y, w2, fmaf(t1.y, w1, t0.y * w0)));
        c2 = fmaf(t3.z, w3, fmaf(t2.z, w2, fmaf(t1.z, w1, t0.z * w0)));
        if (ok) {
          const float t[8] = {c0, c1, c2, 0.f, 0.f, 0.f, 0.f, 0.f};
          a.rec[(first + r) * RC + 9 + sub * 5 + 4] = pack8(t);
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
          float s = 0.f;
#pragma unroll
          for (int v = 0; v < V; ++v) s = fmaf(fv[v][e] - m, fv[v][e] - m, s);
          mean[e] = m;
          var[e] = s * inv_v;
        }
        const uint4 qm = pack8(mean), qv = pack8(var);
        st_chunk_rows8(A1, SBO1, r & ~7, 8, qm, tid & 31);
        st_chunk_rows8(A1, SBO1, r & ~7, 12, qv, tid & 31);
        if (ok) {
          uint4* rp = a.rec + (first + r) * RC;
          rp[sub] = qm;
          rp[4 + sub] = qv;
        }
        // RGB mean / variance over the views: butterfly over the point's 4 lanes (lanes >= V hold 0)
        float m0 = c0, m1 = c1, m2 = c2;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          m0 += __shfl_xor_sync(0xffffffffu, m0, o);
          m1 += __shfl_xor_sync(0xffffffffu, m1, o);
          m2 += __shfl_xor_sync(0xffffffffu, m2, o);
        }
        m0 *= inv_v; m1 *= inv_v; m2 *= inv_v;
        const bool mine = sub < V;
        float s0 = mine ? (c0 - m0) * (c0 - m0) : 0.f, s1 = mine ? (c1 - m1) * (c1 - m1) : 0.f;
        float s2 = mine ? (c2 - m2) * (c2 - m2) : 0.f;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (sub == 0) {
          // columns 70, 71 of the G tile are the constant 1.0 that carries the layers' biases
          const float t[8] = {m0, m1, m2, s0 * inv_v, s1 * inv_v, s2 * inv_v, 1.0f, 1.0f};
          const uint4 qc = pack8(t);
          *reinterpret_cast<uint4*>(A1 + chunk_off(r, 16, SBO1)) = qc;
          if (ok) a.rec[(first + r) * RC + 8] = qc;
        }
      }
    }
    fetch_point(q_next, px, py, pz);                                     // lands during the density phase
    // =================== density phase ===================
    // sigmahead.out_geometry_fc: [128 x 128] · Wgᵀ → 64, ELU → columns 0..63 of A1 (over the plan)
    round_sync();
    if (tid == 0) issue_gemm_bias(a0, op_sbo(128), wimg + DenImg::Wg, 128, 64, ones, SBO1, tmem, bar_m, kFmtF16);
    wait_round(bar_m, phase);
    epi32_to_tile(t_row, half * 32, A1, SBO1, row, 0);
    // out_geometry_fc.0: [128 x 144] → 64, ELU → A0 as [128 x 64]
    round_sync();
    if (tid == 0) issue_gemm(a1, SBO1, wimg + DenImg::W0, SBO1, 144, 64, tmem, bar_m);
    wait_round(bar_m, phase);
    epi32_to_tile(t_row, half * 32, A0, op_sbo(64), row, 0);
    // .2: [128 x 64] → 32, ELU → A2
    round_sync();
    if (tid == 0) issue_gemm_bias(a0, op_sbo(64), wimg + DenImg::W1, 64, 32, ones, SBO1, tmem, bar_m);
    wait_round(bar_m, phase);
    epi16_to_tile(t_row, half * 16, A2, op_sbo(32), row, 0);
    // .4: [128 x 32] → 16, ELU ; .6 + ReLU + fill on CUDA cores
    round_sync();
    if (tid == 0) issue_gemm_bias(a2, op_sbo(32), wimg + DenImg::W2, 32, 16, ones, SBO1, tmem, bar_m);
    wait_round(bar_m, phase);
    if (half == 0) {
      uint32_t r16[16];
      tmem_ld16(t_row, r16);
      tmem_wait_ld();
      float s = fl[DenImg::b3];
#pragma unroll
      for (int k = 0; k < 16; ++k) s = fmaf(elu_scaled(__uint_as_float(r16[k])), fl[DenImg::w3 + k], s);
      s = fmaxf(s, 0.0f);
      s = (nvalid[row] < 1) ? 0.0f : s;
      if (row < n_valid) a.sigma[first + row] = s;
      if (a.alpha != nullptr) {
        // the progressive step's test (demo_render.py:312-317, as alpha_flags computes it), fused: the 4
        // warps of this half hold the tile's 128 consecutive points = 4 flag words
        const float al = xsub(1.0f, expf(-s));
        const bool keep = row < n_valid && al > 1e-14f;
        if (row < n_valid) a.alpha[first + row] = al;
        const unsigned bits = __ballot_sync(0xffffffffu, keep);
        if ((tid & 31) == 0) a.alpha_words[(first + row) >> 5] = bits;
      }
    }
    // the next plan phase rewrites A1/nvalid: order it after every thread's
    // reads of this tile (TMEM reads are fenced by the next round_sync)
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

}  // namespace gpnerf

using namespace gpnerf;

int gpnerf_color_mlp_tc_any(const float* rgb_feat, const float* meanvar, const void* rec, const int32_t* valid1,
                            const gpnerf_head_weights_t* w, int n_views, int n_points_max,
                            const int32_t* count_ptr, float* rgb, cudaStream_t st);

int launch_fused_ws(const gpnerf::FusedArgs& a, const gpnerf_frame_t* f, int n_points_max, cudaStream_t st);

template <int V>
static int launch_fused(const FusedArgs& a, const gpnerf_frame_t* f, int n_points_max, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gather_density_tc<V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FusedSmem::BYTES);
    if (e != cudaSuccess) {
      set_error("gather_density_tc smem attribute", e);
      return GPNERF_E_CUDA;
    }
    attr_set = true;
  }
  int tiles = (n_points_max + 127) / 128;
  static const int per_sm = getenv("GPNERF_FUSED_CTAS_PER_SM") ? atoi(getenv("GPNERF_FUSED_CTAS_PER_SM")) : 2;   // experiment knob
  int grid = tiles < per_sm * sm_count() ? tiles : per_sm * sm_count();
  gather_density_tc<V><<<grid, 256, FusedSmem::BYTES, st>>>(a, *f);
  return check_launch("k23_gather_density_tc");
}

extern "C" {

int64_t gpnerf_k23_record_bytes(int n_views) { return 16ll * rec_chunks(n_views); }

static int fused_args(FusedArgs& a, const void* const levels_f16[GPNERF_N_LEVELS], const void* featmaps_f16,
                      const float* images_rgbx, const int32_t* valid, const float* rays_o, const float* rays_d,
                      const float* z_vals, const gpnerf_frame_t* f, const gpnerf_head_weights_t* w, int n_points_max,
                      const int32_t* counters, float* sigma, float* alpha, void* k4_workspace) {
  GPNERF_REQUIRE(levels_f16 && featmaps_f16 && images_rgbx && valid && rays_o && rays_d && z_vals && f && w &&
                 counters && sigma && n_points_max > 0);
  GPNERF_REQUIRE(w->tc_image != nullptr && f->n_samples > 0 && f->src_w > 1 && f->src_h > 1);
  for (int l = 0; l < GPNERF_N_LEVELS; ++l) {
    GPNERF_REQUIRE(levels_f16[l] != nullptr);
    a.lv[l] = reinterpret_cast<const __half*>(levels_f16[l]);
  }
  a.feat = reinterpret_cast<const __half*>(featmaps_f16);
  a.rgbx = reinterpret_cast<const float4*>(images_rgbx);
  a.valid = valid; a.rays_o = rays_o; a.rays_d = rays_d; a.z_vals = z_vals;
  a.counters = counters;
  a.image = reinterpret_cast<const uint8_t*>(w->tc_image);
  a.sigma = sigma;
  a.rec = nullptr;
  a.rec_tiles = nullptr;
  a.rgb_in = nullptr;
  static const int dbg = getenv("GPNERF_FUSED_DEBUG") ? atoi(getenv("GPNERF_FUSED_DEBUG")) : 0;
  a.debug = dbg;
  GPNERF_REQUIRE((alpha == nullptr) == (k4_workspace == nullptr));
  a.alpha = alpha;
  a.alpha_words = alpha ? carve_workspace(k4_workspace, n_points_max).words : nullptr;
  return GPNERF_OK;
}

int gpnerf_k23_gather_density_tc(const void* const levels_f16[GPNERF_N_LEVELS], const void* featmaps_f16,
                                 const float* images_rgbx, const int32_t* valid, const float* rays_o,
                                 const float* rays_d, const float* z_vals, const gpnerf_frame_t* f,
                                 const gpnerf_head_weights_t* w, int n_points_max, const int32_t* counters,
                                 float* sigma, void* records, float* alpha, void* k4_workspace, void* stream) {
  FusedArgs a;
  int rc = fused_args(a, levels_f16, featmaps_f16, images_rgbx, valid, rays_o, rays_d, z_vals, f, w, n_points_max,
                      counters, sigma, alpha, k4_workspace);
  if (rc != GPNERF_OK) return rc;
  a.rec = reinterpret_cast<uint4*>(records);
  cudaStream_t st = (cudaStream_t)stream;
  // Per-point records (round 1's hand-off to gpnerf_k3_color_mlp_records) are written by round 1's kernel only: one
  // 256-thread CTA does plan → gather → 4 MMA rounds serially, two CTAs per SM.  GPNERF_FUSED_IMPL=monolithic selects
  // it without records as well; the default is the warp-specialised kernel of k23_fused_ws.cu.
  static const bool monolithic = getenv("GPNERF_FUSED_IMPL") && !strcmp(getenv("GPNERF_FUSED_IMPL"), "monolithic");
  if (!monolithic && records == nullptr) return launch_fused_ws(a, f, n_points_max, st);
  GPNERF_REQUIRE(records != nullptr);       // the round-1 kernel always writes its colour records
  switch (f->n_views) {
    case 1: return launch_fused<1>(a, f, n_points_max, st);
    case 2: return launch_fused<2>(a, f, n_points_max, st);
    case 3: return launch_fused<3>(a, f, n_points_max, st);
    case 4: return launch_fused<4>(a, f, n_points_max, st);
    default:
      set_error("fused tcgen05 path supports 1..4 source views", cudaSuccess);
      return GPNERF_E_UNSUPPORTED;
  }
}

int gpnerf_k23_gather_density_tiles_tc(const void* const levels_f16[GPNERF_N_LEVELS], const void* featmaps_f16,
                                       const float* images_rgbx, const int32_t* valid, const float* rays_o,
                                       const float* rays_d, const float* z_vals, const gpnerf_frame_t* f,
                                       const gpnerf_head_weights_t* w, int n_points_max, const int32_t* counters,
                                       float* sigma, void* tile_records, float* rgb_in, float* alpha,
                                       void* k4_workspace, void* stream) {
  FusedArgs a;
  int rc = fused_args(a, levels_f16, featmaps_f16, images_rgbx, valid, rays_o, rays_d, z_vals, f, w, n_points_max,
                      counters, sigma, alpha, k4_workspace);
  if (rc != GPNERF_OK) return rc;
  GPNERF_REQUIRE(tile_records != nullptr);
  a.rec_tiles = reinterpret_cast<uint8_t*>(tile_records);
  a.rgb_in = rgb_in;
  return launch_fused_ws(a, f, n_points_max, (cudaStream_t)stream);
}

int gpnerf_k3_color_mlp_records(const void* records, const int32_t* valid1, const gpnerf_head_weights_t* w,
                                int n_views, int n_points_max, const int32_t* counters, int counter_slot,
                                float* rgb, void* stream) {
  GPNERF_REQUIRE(records && w && rgb && n_points_max > 0 && counter_slot >= 0 && counter_slot < GPNERF_N_COUNTERS);
  return gpnerf_color_mlp_tc_any(nullptr, nullptr, records, valid1, w, n_views, n_points_max,
                                 counters ? counters + counter_slot : nullptr, rgb, (cudaStream_t)stream);
}

}  // extern "C"
