// K4 – GP-NeRF's progressive step (demo_render.py:312-317): α = 1−exp(−σ) and
// the ordered compaction of the points whose α survives, so that only they
// reach the colour head.  4 B read + 4 B written per point, 4 B per survivor.
//
// K5 – front-to-back compositing.
//   composite_tiles  : demo_render.py:335-353.  `valid` is ascending in
//     ray·S+sample, so the survivors of a ray are one contiguous run (CSR
//     offsets from the compactions); a warp walks the run 32 samples at a time
//     with a shuffle product-scan for the transmittance.  Samples that did not
//     survive have α = 0 and contribute the factor fl32(1+1e-10) = 1, so
//     skipping them is exact.  Reads 16 B per surviving sample; publishes whole
//     pixel tiles, locally and into peer GPUs' images.
//   raw2outputs      : BaseRender.py:75-107,147 dense [R][S] variant with the
//     auxiliary maps the training loss consumes.
#include <string.h>

#include "common.cuh"

namespace gpnerf {

__global__ void __launch_bounds__(256) alpha_flags(const float* __restrict__ sigma,
                                                   const int32_t* __restrict__ counters,
                                                   float* __restrict__ alpha,
                                                   uint32_t* __restrict__ words) {
  const int n = __ldg(counters + GPNERF_CNT_P1);
  const int n_pad = (n + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
    bool keep = false;
    if (i < n) {
      float a = xsub(1.0f, expf(-__ldg(sigma + i)));
      alpha[i] = a;
      keep = a > 1e-14f;
    }
    unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) words[i >> 5] = b;
  }
}

// One CTA per pixel tile of this rank (frame.rank/world/tile_px: the diagonal deal of K1), one warp
// per ray of the tile.  `tile_ray_begin` / `ray_pt_begin` are the CSR offsets the two compactions leave
// behind (rays of a tile, surviving points of a ray), so nothing is searched.  The finished tile
// (tile_px RGB triples + hit bytes, zeros where no ray) is staged in shared memory and written with
// coalesced stores to every destination: the local image and – when the frame is sharded over GPUs –
// the same tile of every peer's image, straight over NVLink (peer pointers from gpnerf_peer_t).  The
// last CTA to finish publishes the frame's sequence number in each peer's arrival flag
// (fence.sys + st.release.sys); gpnerf_peer_wait is the matching acquire.  Every pixel of an owned tile
// is written, so no destination needs a memset.
constexpr int kMaxTilePx = 256;
__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) composite_tiles(const float* __restrict__ alpha,
                                                       const float* __restrict__ rgb,
                                                       const int32_t* __restrict__ ray_pix,
                                                       const int32_t* __restrict__ tile_ray_begin,
                                                       const int32_t* __restrict__ ray_pt_begin,
                                                       const __grid_constant__ gpnerf_frame_t fparam, float t_min,
                                                       float* __restrict__ rgb_map, float* __restrict__ pred_img,
                                                       uint8_t* __restrict__ hit_mask,
                                                       const gpnerf_peer_t* __restrict__ peer_dev) {
  GPNERF_LOAD_FRAME(fparam)
  __shared__ float tile_rgb[kMaxTilePx * 3];
  __shared__ __align__(4) uint8_t tile_hit[kMaxTilePx];
  __shared__ gpnerf_peer_t pr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (peer_dev != nullptr) {
    for (int i = tid; i < (int)(sizeof(gpnerf_peer_t) / 4); i += blockDim.x)
      reinterpret_cast<uint32_t*>(&pr)[i] = reinterpret_cast<const uint32_t*>(peer_dev)[i];
    __syncthreads();
  }
  // destinations: without a peer block the plain local pointers; with one, everything listed there
  // (entry 0 is this rank's own buffer of the current frame by convention)
  const int n_dst = peer_dev ? pr.n_dst : 1;
  // frame number of this launch: stable while the kernel runs (only its last CTA advances the counter)
  const int32_t seq = peer_dev ? *reinterpret_cast<const volatile int32_t*>(pr.seq) + 1 : 0;
  const int hb = seq & 1;
  const int tile_px = f.tile_px, n_px = f.H * f.W;
  const int n_tiles = (n_px + tile_px - 1) / tile_px;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    if ((t + (int)(((long long)t * tile_px) / f.W)) % f.world != f.rank) continue;   // not this rank's tile
    const int px0 = t * tile_px, npix = min(tile_px, n_px - px0);
    for (int i = tid; i < npix * 3; i += blockDim.x) tile_rgb[i] = 0.0f;
    for (int i = tid; i < (npix + 3) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tile_hit)[i] = 0u;
    __syncthreads();
    const int r0 = __ldg(tile_ray_begin + t), r1 = __ldg(tile_ray_begin + t + 1);
    const int n_warps = (int)(blockDim.x >> 5);
    // this warp's rays are r0 + warp + j·n_warps: lane j fetches the CSR offsets of ray j up front (one
    // round trip for all of them; a tile has at most tile_px <= 256 rays, i.e. <= 32 per warp)
    int my_begin = 0, my_end = 0;
    {
      const int rj = r0 + warp + lane * n_warps;
      if (rj < r1) {
        my_begin = __ldg(ray_pt_begin + rj);
        my_end = __ldg(ray_pt_begin + rj + 1);
      }
    }
    int j = 0;
    for (int r = r0 + warp; r < r1; r += n_warps, ++j) {
      const int begin = __shfl_sync(0xffffffffu, my_begin, j), end = __shfl_sync(0xffffffffu, my_end, j);
      float T = 1.0f;            // transmittance carried between 32-sample chunks
      float cr = 0.f, cg = 0.f, cb = 0.f;
      for (int base = begin; base < end; base += 32) {
        const int i = base + lane;
        float a = 0.0f, pr_ = 0.f, pg = 0.f, pb = 0.f;
        if (i < end) {
          // colour exists only for points that passed K4 (a > 1e-14); the loads are issued together with
          // α's and discarded otherwise (the rows of culled points hold stale finite-or-not bits)
          a = __ldg(alpha + i);
          const float q0 = __ldg(rgb + (long long)i * 3), q1 = __ldg(rgb + (long long)i * 3 + 1);
          const float q2 = __ldg(rgb + (long long)i * 3 + 2);
          const bool has = a > 1e-14f;
          pr_ = has ? q0 : 0.f;
          pg = has ? q1 : 0.f;
          pb = has ? q2 : 0.f;
        }
        float fct = xadd(xsub(1.0f, a), 1e-10f);
        float incl = fct;        // inclusive product scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float tt = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl *= tt;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float wgt = a * (T * excl);
        cr += wgt * pr_;
        cg += wgt * pg;
        cb += wgt * pb;
        T *= __shfl_sync(0xffffffffu, incl, 31);
        if (T < t_min) break;    // early termination (off when t_min == 0)
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cr += __shfl_xor_sync(0xffffffffu, cr, o);
        cg += __shfl_xor_sync(0xffffffffu, cg, o);
        cb += __shfl_xor_sync(0xffffffffu, cb, o);
      }
      if (lane == 0) {
        rgb_map[(long long)r * 3 + 0] = cr;
        rgb_map[(long long)r * 3 + 1] = cg;
        rgb_map[(long long)r * 3 + 2] = cb;
        const int p = __ldg(ray_pix + r) - px0;
        tile_rgb[p * 3 + 0] = cr;
        tile_rgb[p * 3 + 1] = cg;
        tile_rgb[p * 3 + 2] = cb;
        tile_hit[p] = 1;
      }
    }
    __syncthreads();
    for (int k = 0; k < n_dst; ++k) {
      float* img = peer_dev ? reinterpret_cast<float*>(pr.dst_img[hb][k]) : pred_img;
      uint8_t* hit = peer_dev ? reinterpret_cast<uint8_t*>(pr.dst_hit[hb][k]) : hit_mask;
      for (int i = tid; i < npix * 3; i += blockDim.x) img[(long long)px0 * 3 + i] = tile_rgb[i];
      if ((npix & 3) == 0 && (px0 & 3) == 0) {
        for (int i = tid; i < npix / 4; i += blockDim.x)
          reinterpret_cast<uint32_t*>(hit + px0)[i] = reinterpret_cast<const uint32_t*>(tile_hit)[i];
      } else {
        for (int i = tid; i < npix; i += blockDim.x) hit[px0 + i] = tile_hit[i];
      }
    }
    __syncthreads();
  }
  if (peer_dev != nullptr) {
    __threadfence_system();      // this thread's peer stores before the CTA's ticket
    __syncthreads();
    if (tid == 0) {
      int32_t* ticket = reinterpret_cast<int32_t*>(pr.ticket);
      const int done = atomicAdd(ticket, 1);
      if (done == (int)gridDim.x - 1) {     // every CTA has read the counter and written its tiles
        *ticket = 0;                        // ready for the next launch (graph replay)
        *reinterpret_cast<volatile int32_t*>(pr.seq) = seq;
        __threadfence_system();
        for (int k = 0; k < pr.n_flag; ++k) st_release_sys(reinterpret_cast<int32_t*>(pr.dst_flag[k]), seq);
      }
    }
  }
}

// Matching acquire: thread k waits until flags[k] has reached the frame's sequence number (peers k
// publish it from composite_tiles when all their tiles have landed here).  Bounded: traps after ~30 s
// (a peer that is still capturing its graph or paging its image in must not take the job down).
__global__ void peer_wait_kernel(const int32_t* __restrict__ flags, int n_flags, int self,
                                 const gpnerf_peer_t* __restrict__ peer_dev) {
  const int k = threadIdx.x;
  if (k >= n_flags || k == self) return;
  const int32_t seq = *reinterpret_cast<const volatile int32_t*>(peer_dev->seq);   // K5 ran before us
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flags + k) - seq) < 0) {
    __nanosleep(200);
    if (clock64() - t0 > 60000000000ll) __trap();
  }
}

// one warp per ray over the dense [S] samples
__global__ void __launch_bounds__(256) raw2outputs_kernel(const float* __restrict__ raw,
                                                          const float* __restrict__ z_vals,
                                                          const float* __restrict__ rgb_in, int n_rays,
                                                          int S, int V, int neg,
                                                          float* __restrict__ rgb_map,
                                                          float* __restrict__ disp, float* __restrict__ acc,
                                                          float* __restrict__ depth,
                                                          float* __restrict__ weights,
                                                          float* __restrict__ rgb_in_map) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n_rays; r += n_warps) {
    float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, dsum = 0.f, asum = 0.f;
    float vin[GPNERF_MAX_VIEWS * 3];
#pragma unroll
    for (int k = 0; k < GPNERF_MAX_VIEWS * 3; ++k) vin[k] = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int j = base + lane;              // position in compositing order
      const int s = neg ? (S - 1 - j) : j;    // flipped rgb/σ when neg (BaseRender.py:86-88)
      float a = 0.f, pr = 0.f, pg = 0.f, pb = 0.f, z = 0.f;
      if (j < S) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(raw + ((long long)r * S + s) * 4));
        a = xsub(1.0f, expf(-q.w));
        pr = q.x; pg = q.y; pb = q.z;
        z = __ldg(z_vals + (long long)r * S + j);   // z_vals are NOT flipped by the reference
      }
      float fct = xadd(xsub(1.0f, a), 1e-10f);
      float incl = fct;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= t;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float wgt = a * (T * excl);
      if (j < S) {
        weights[(long long)r * S + j] = wgt;
        cr += wgt * pr; cg += wgt * pg; cb += wgt * pb;
        dsum += wgt * z; asum += wgt;
        if (rgb_in != nullptr) {
          // rgb_in is not flipped either (BaseRender.py:147)
          const float* ri = rgb_in + ((long long)r * S + j) * V * 3;
          for (int k = 0; k < V * 3; ++k) vin[k] += wgt * __ldg(ri + k);
        }
      }
      T *= __shfl_sync(0xffffffffu, incl, 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cr += __shfl_xor_sync(0xffffffffu, cr, o);
      cg += __shfl_xor_sync(0xffffffffu, cg, o);
      cb += __shfl_xor_sync(0xffffffffu, cb, o);
      dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
      asum += __shfl_xor_sync(0xffffffffu, asum, o);
    }
    if (rgb_in != nullptr) {
      for (int k = 0; k < V * 3; ++k) {
        float t = vin[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) rgb_in_map[(long long)r * V * 3 + k] = t;
      }
    }
    if (lane == 0) {
      rgb_map[(long long)r * 3 + 0] = cr;
      rgb_map[(long long)r * 3 + 1] = cg;
      rgb_map[(long long)r * 3 + 2] = cb;
      depth[r] = dsum;
      acc[r] = asum;
      const float q = dsum / asum;                       // 0/0 → NaN on empty rays, as in the reference
      disp[r] = 1.0f / ((q != q) ? q : fmaxf(1e-10f, q));  // torch.max propagates NaN (BaseRender.py:101)
    }
  }
}


// Backward of raw2outputs (training path).  With G_j = ∂L/∂w_j collected from
// every consumer of the weights (rgb_map, depth, acc, disp, the returned
// weights themselves, rgb_in_map):
//   ∂L/∂c_j = w_j · g_rgb
//   ∂L/∂σ_j = (1−α_j) · [ G_j·T_j − (Σ_{k>j} G_k·w_k) / (1−α_j+1e-10) ]
// One warp per ray: forward product-scan recomputed, suffix sums by a reverse
// shuffle scan, 32 samples at a time (two passes over the ray).
__global__ void __launch_bounds__(256) raw2outputs_bwd_kernel(
    const float* __restrict__ raw, const float* __restrict__ z_vals, const float* __restrict__ rgb_in, int n_rays,
    int S, int V, int neg, const float* __restrict__ g_rgb_map, const float* __restrict__ g_disp,
    const float* __restrict__ g_acc, const float* __restrict__ g_depth, const float* __restrict__ g_weights,
    const float* __restrict__ g_rgb_in_map, float* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n_rays; r += n_warps) {
    const float gr = g_rgb_map ? __ldg(g_rgb_map + r * 3) : 0.f, gg = g_rgb_map ? __ldg(g_rgb_map + r * 3 + 1) : 0.f,
                gb = g_rgb_map ? __ldg(g_rgb_map + r * 3 + 2) : 0.f;
    float gd = g_depth ? __ldg(g_depth + r) : 0.f, ga = g_acc ? __ldg(g_acc + r) : 0.f;
    // pass 1: forward sums needed by disp = 1 / max(1e-10, depth/acc)
    if (g_disp != nullptr) {
      float T = 1.0f, dsum = 0.f, asum = 0.f;
      for (int base = 0; base < S; base += 32) {
        const int j = base + lane, s = neg ? (S - 1 - j) : j;
        float a = 0.f, z = 0.f;
        if (j < S) {
          a = xsub(1.0f, expf(-__ldg(raw + ((long long)r * S + s) * 4 + 3)));
          z = __ldg(z_vals + (long long)r * S + j);
        }
        float incl = xadd(xsub(1.0f, a), 1e-10f);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl *= t;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float w = a * (T * excl);
        dsum += w * z;
        asum += w;
        T *= __shfl_sync(0xffffffffu, incl, 31);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        asum += __shfl_xor_sync(0xffffffffu, asum, o);
      }
      const float q = dsum / asum;
      if (q > 1e-10f) {             // disp = acc/depth ; the clamped / NaN branch has zero gradient
        const float gq = -__ldg(g_disp + r) / (q * q);
        gd += gq / asum;
        ga += -gq * dsum / (asum * asum);
      }
    }
    // pass 2: chunks in reverse order, carrying the suffix sum Σ_{k>j} G_k w_k; the transmittance at a
    // chunk start is the product of the factors of all earlier chunks, recomputed per chunk (S/32 ≤ 8)
    float suffix = 0.f;
    const int n_chunks = (S + 31) / 32;
    for (int ch = n_chunks - 1; ch >= 0; --ch) {
      float T0 = 1.0f;
      for (int c2 = 0; c2 < ch; ++c2) {
        const int j = c2 * 32 + lane, s = neg ? (S - 1 - j) : j;
        float fct = 1.0f;
        if (j < S) fct = xadd(xsub(1.0f, xsub(1.0f, expf(-__ldg(raw + ((long long)r * S + s) * 4 + 3)))), 1e-10f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fct *= __shfl_xor_sync(0xffffffffu, fct, o);
        T0 *= fct;
      }
      const int j = ch * 32 + lane, s = neg ? (S - 1 - j) : j;
      float a = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, G = 0.f;
      if (j < S) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(raw + ((long long)r * S + s) * 4));
        a = xsub(1.0f, expf(-q.w));
        cr = q.x; cg = q.y; cb = q.z;
        G = gr * cr + gg * cg + gb * cb + gd * __ldg(z_vals + (long long)r * S + j) + ga;
        if (g_weights) G += __ldg(g_weights + (long long)r * S + j);
        if (g_rgb_in_map && rgb_in) {
          const float* ri = rgb_in + ((long long)r * S + j) * V * 3;
          const float* gi = g_rgb_in_map + (long long)r * V * 3;
          for (int k = 0; k < V * 3; ++k) G += __ldg(gi + k) * __ldg(ri + k);
        }
      }
      const float fct = xadd(xsub(1.0f, a), 1e-10f);
      float incl = fct;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= t;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float Tj = T0 * excl, w = a * Tj;
      // inclusive suffix sum of G·w inside the chunk (reverse scan), plus the carry of later chunks
      float gw = G * w, suf = gw;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_down_sync(0xffffffffu, suf, o);
        if (lane + o < 32) suf += t;
      }
      const float after = suf - gw + suffix;          // Σ_{k>j} G_k w_k
      if (j < S) {
        float* o = d_raw + ((long long)r * S + s) * 4;
        o[0] = w * gr;
        o[1] = w * gg;
        o[2] = w * gb;
        o[3] = (1.0f - a) * (G * Tj - after / fct);
      }
      suffix += __shfl_sync(0xffffffffu, suf, 0);
    }
  }
}

}  // namespace gpnerf

using namespace gpnerf;

extern "C" {

int gpnerf_k4_compact_alpha(const float* sigma, int n_points_max, int32_t* counters, float* alpha,
                            int32_t* valid1, void* workspace, void* stream) {
  GPNERF_REQUIRE(counters && alpha && valid1 && workspace && n_points_max > 0);
  cudaStream_t st = (cudaStream_t)stream;
  CompactWs ws = carve_workspace(workspace, n_points_max);
  if (sigma != nullptr) {     // else α and the flag words were written by gpnerf_k23_gather_density_tc
    long long blocks = ((long long)n_points_max + 255) / 256;
    int grid = (int)(blocks < (long long)sm_count() * 8 ? blocks : sm_count() * 8);
    alpha_flags<<<grid, 256, 0, st>>>(sigma, counters, alpha, ws.words);
  }
  return compact_launch(ws, counters + GPNERF_CNT_P1, 1, 0, n_points_max, valid1,
                        counters + GPNERF_CNT_P2, st);
}

int gpnerf_k5_composite(const float* alpha, const float* rgb, const int32_t* ray_pix,
                        const int32_t* tile_ray_begin, const int32_t* ray_pt_begin, const gpnerf_frame_t* f,
                        float t_min, float* rgb_map, float* pred_img, uint8_t* hit_mask,
                        const gpnerf_peer_t* peer_dev, void* stream) {
  GPNERF_REQUIRE(alpha && rgb && ray_pix && tile_ray_begin && ray_pt_begin && f && rgb_map);
  GPNERF_REQUIRE(peer_dev != nullptr || (pred_img && hit_mask));
  GPNERF_REQUIRE(f->H > 0 && f->W > 0 && f->tile_px >= 1 && f->tile_px <= kMaxTilePx && f->world >= 1);
  const long long n_px = (long long)f->H * f->W;
  const long long n_tiles = (n_px + f->tile_px - 1) / f->tile_px;
  const long long cap = (long long)sm_count() * 8;
  const int grid = (int)(n_tiles < cap ? n_tiles : cap);
  composite_tiles<<<grid, 256, 0, (cudaStream_t)stream>>>(alpha, rgb, ray_pix, tile_ray_begin, ray_pt_begin, *f,
                                                          t_min, rgb_map, pred_img, hit_mask, peer_dev);
  return check_launch("k5_composite");
}

int gpnerf_peer_wait(const int32_t* flags, int n_flags, int self, const gpnerf_peer_t* peer_dev, void* stream) {
  GPNERF_REQUIRE(flags && peer_dev && n_flags >= 1 && n_flags <= GPNERF_MAX_PEERS);
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, n_flags, self, peer_dev);
  return check_launch("peer_wait");
}

// ---- peer (IPC) memory: one allocation per rank, opened by every other rank of the box -------------
int gpnerf_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int gpnerf_peer_alloc(int64_t bytes, void** dev_ptr_host, void* handle_host) {
  GPNERF_REQUIRE(bytes > 0 && dev_ptr_host && handle_host);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_host), p);
  if (e != cudaSuccess) {
    set_error("peer_alloc", e);
    if (p) cudaFree(p);
    return GPNERF_E_CUDA;
  }
  *dev_ptr_host = p;
  return GPNERF_OK;
}

int gpnerf_peer_open(const void* handle_host, void** dev_ptr_host) {
  GPNERF_REQUIRE(handle_host && dev_ptr_host);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("peer_open (cudaIpcOpenMemHandle)", e);
    return GPNERF_E_CUDA;
  }
  *dev_ptr_host = p;
  return GPNERF_OK;
}

int gpnerf_peer_close(void* dev_ptr) {
  GPNERF_REQUIRE(dev_ptr);
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) {
    set_error("peer_close", e);
    return GPNERF_E_CUDA;
  }
  return GPNERF_OK;
}

int gpnerf_peer_free(void* dev_ptr) {
  GPNERF_REQUIRE(dev_ptr);
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) {
    set_error("peer_free", e);
    return GPNERF_E_CUDA;
  }
  return GPNERF_OK;
}

int gpnerf_k5_raw2outputs(const float* raw, const float* z_vals, const float* rgb_in, int n_rays,
                          int n_samples, int n_views, int neg, float* rgb_map, float* disp, float* acc,
                          float* depth, float* weights, float* rgb_in_map, void* stream) {
  GPNERF_REQUIRE(raw && z_vals && rgb_map && disp && acc && depth && weights && n_rays > 0 && n_samples > 0);
  GPNERF_REQUIRE(n_views >= 0 && n_views <= GPNERF_MAX_VIEWS && (rgb_in == nullptr || rgb_in_map != nullptr));
  long long blocks = ((long long)n_rays * 32 + 255) / 256;
  int grid = (int)(blocks < (long long)sm_count() * 8 ? blocks : sm_count() * 8);
  raw2outputs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(raw, z_vals, rgb_in, n_rays, n_samples,
                                                             n_views, neg, rgb_map, disp, acc, depth,
                                                             weights, rgb_in_map);
  return check_launch("k5_raw2outputs");
}

int gpnerf_k5_raw2outputs_bwd(const float* raw, const float* z_vals, const float* rgb_in, int n_rays,
                              int n_samples, int n_views, int neg, const float* g_rgb_map, const float* g_disp,
                              const float* g_acc, const float* g_depth, const float* g_weights,
                              const float* g_rgb_in_map, float* d_raw, void* stream) {
  GPNERF_REQUIRE(raw && z_vals && d_raw && n_rays > 0 && n_samples > 0);
  GPNERF_REQUIRE(n_views >= 0 && n_views <= GPNERF_MAX_VIEWS && (g_rgb_in_map == nullptr || rgb_in != nullptr));
  long long blocks = ((long long)n_rays * 32 + 255) / 256;
  int grid = (int)(blocks < (long long)sm_count() * 8 ? blocks : sm_count() * 8);
  raw2outputs_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(raw, z_vals, rgb_in, n_rays, n_samples, n_views,
                                                                 neg, g_rgb_map, g_disp, g_acc, g_depth, g_weights,
                                                                 g_rgb_in_map, d_raw);
  return check_launch("k5_raw2outputs_bwd");
}

}  // extern "C"
