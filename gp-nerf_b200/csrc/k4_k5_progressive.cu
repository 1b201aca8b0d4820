// K4 – GP-NeRF's progressive step (demo_render.py:312-317): α = 1−exp(−σ) and
// the ordered compaction of the points whose α survives, so that only they
// reach the colour head.  4 B read + 4 B written per point, 4 B per survivor.
//
// K5 – front-to-back compositing.
//   composite_rays   : demo_render.py:335-353.  `valid` is ascending in
//     ray·S+sample, so the survivors of a ray are one contiguous run (CSR by
//     binary search); a warp walks the run 32 samples at a time with a
//     shuffle product-scan for the transmittance.  Samples that did not
//     survive have α = 0 and contribute the factor fl32(1+1e-10) = 1, so
//     skipping them is exact.  Reads 16 B per surviving sample.
//   raw2outputs      : BaseRender.py:75-107,147 dense [R][S] variant with the
//     auxiliary maps the training loss consumes.
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

__device__ __forceinline__ int lower_bound(const int32_t* __restrict__ a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// one warp per ray
__global__ void __launch_bounds__(256) composite_rays(const int32_t* __restrict__ valid,
                                                      const float* __restrict__ alpha,
                                                      const float* __restrict__ rgb,
                                                      const int32_t* __restrict__ ray_pix,
                                                      const int32_t* __restrict__ counters, int S,
                                                      float t_min, float* __restrict__ rgb_map,
                                                      float* __restrict__ pred_img,
                                                      uint8_t* __restrict__ hit_mask) {
  const int n_rays = __ldg(counters + GPNERF_CNT_RAYS);
  const int n_pts = __ldg(counters + GPNERF_CNT_P1);
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n_rays; r += n_warps) {
    int b0 = 0, b1 = 0;
    if (lane == 0) b0 = lower_bound(valid, n_pts, r * S);
    if (lane == 1) b1 = lower_bound(valid, n_pts, (r + 1) * S);
    const int begin = __shfl_sync(0xffffffffu, b0, 0), end = __shfl_sync(0xffffffffu, b1, 1);
    float T = 1.0f;            // transmittance carried between 32-sample chunks
    float cr = 0.f, cg = 0.f, cb = 0.f;
    for (int base = begin; base < end; base += 32) {
      const int i = base + lane;
      float a = 0.0f, pr = 0.f, pg = 0.f, pb = 0.f;
      if (i < end) {
        a = __ldg(alpha + i);
        if (a > 1e-14f) {      // colour exists only for points that passed K4
          pr = __ldg(rgb + (long long)i * 3);
          pg = __ldg(rgb + (long long)i * 3 + 1);
          pb = __ldg(rgb + (long long)i * 3 + 2);
        }
      }
      float fct = xadd(xsub(1.0f, a), 1e-10f);
      float incl = fct;        // inclusive product scan
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= t;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float wgt = a * (T * excl);
      cr += wgt * pr;
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
      const int p = __ldg(ray_pix + r);
      pred_img[(long long)p * 3 + 0] = cr;
      pred_img[(long long)p * 3 + 1] = cg;
      pred_img[(long long)p * 3 + 2] = cb;
      hit_mask[p] = 1;
    }
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
  GPNERF_REQUIRE(sigma && counters && alpha && valid1 && workspace && n_points_max > 0);
  cudaStream_t st = (cudaStream_t)stream;
  CompactWs ws = carve_workspace(workspace, n_points_max);
  long long blocks = ((long long)n_points_max + 255) / 256;
  int grid = (int)(blocks < (long long)sm_count() * 8 ? blocks : sm_count() * 8);
  alpha_flags<<<grid, 256, 0, st>>>(sigma, counters, alpha, ws.words);
  return compact_launch(ws, counters + GPNERF_CNT_P1, 1, 0, n_points_max, valid1,
                        counters + GPNERF_CNT_P2, st);
}

int gpnerf_k5_composite(const int32_t* valid, const float* alpha, const float* rgb,
                        const int32_t* ray_pix, const gpnerf_frame_t* f, int n_rays_max,
                        const int32_t* counters, float t_min, float* rgb_map, float* pred_img,
                        uint8_t* hit_mask, void* stream) {
  GPNERF_REQUIRE(valid && alpha && rgb && ray_pix && f && counters && rgb_map && pred_img && hit_mask && n_rays_max > 0);
  cudaStream_t st = (cudaStream_t)stream;
  size_t npx = (size_t)f->H * f->W;
  cudaError_t e = cudaMemsetAsync(pred_img, 0, npx * 3 * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(hit_mask, 0, npx, st);
  if (e != cudaSuccess) {
    set_error("memset pred_img", e);
    return GPNERF_E_CUDA;
  }
  long long blocks = ((long long)n_rays_max * 32 + 255) / 256;
  int grid = (int)(blocks < (long long)sm_count() * 8 ? blocks : sm_count() * 8);
  composite_rays<<<grid, 256, 0, st>>>(valid, alpha, rgb, ray_pix, counters, f->n_samples, t_min,
                                       rgb_map, pred_img, hit_mask);
  return check_launch("k5_composite");
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
