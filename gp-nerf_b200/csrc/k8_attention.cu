// K8 – SURVEY §8f row 1, the step before the sparse convolution: every SMPL vertex's latent code attends
// to the pixel-aligned features of that vertex in the V source views
// (libs/nerfheads/trainhead.py:48-51 → libs/nerfheads/networks/MultiHeadAttention.py:40-98 with
// sum=False, mask=None: no residual, no LayerNorm, no dropout on this path).
//
//   q = W_q·code / sqrt(d_k)            [n_head·d_k]
//   k_v = W_k·feat_v,  val_v = W_v·feat_v   (v = 0…V-1)
//   a_hv = softmax_v(q_h·k_vh),  o_h = Σ_v a_hv·val_vh,  out = W_fc·o        [d_model]
//
// One thread per vertex, the four weight matrices (6 KB) in shared memory.  6890 vertices × V ≤ 8 views:
// ≈ 25 MFLOP and 3 MB per frame – latency-, not throughput-bound; the point of the kernel is that the
// whole chain (project → gather → attend → sparse conv) stays on the stream with no library call.
// Two passes over the views (logits, then values) keep the softmax in torch's form exp(l − max)/Σ.
#include "common.cuh"

namespace gpnerf {

template <int DM, int KV, int HD>
__global__ void __launch_bounds__(64) smpl_code_attention(const float* __restrict__ code, const float* __restrict__ feats,
                                                          long long view_stride, long long vertex_stride, int n, int V,
                                                          const float* __restrict__ w_q, const float* __restrict__ w_k,
                                                          const float* __restrict__ w_v, const float* __restrict__ w_fc,
                                                          int n_head, float inv_temp, float* __restrict__ out) {
  __shared__ float sq[HD * DM], sk[HD * KV], sv[HD * KV], sf[DM * HD];
  for (int i = threadIdx.x; i < HD * DM; i += blockDim.x) sq[i] = w_q[i], sf[i] = w_fc[i];
  for (int i = threadIdx.x; i < HD * KV; i += blockDim.x) sk[i] = w_k[i], sv[i] = w_v[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int dk = HD / n_head;
  float q[HD];
  {
    float c[DM];
#pragma unroll
    for (int j = 0; j < DM; ++j) c[j] = __ldg(code + (size_t)i * DM + j);
#pragma unroll
    for (int o = 0; o < HD; ++o) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < DM; ++j) a = fmaf(sq[o * DM + j], c[j], a);
      q[o] = a;
    }
  }
  // torch divides q by the temperature before the product (MultiHeadAttention.py:29)
#pragma unroll
  for (int o = 0; o < HD; ++o) q[o] *= inv_temp;
  const float* f0 = feats + (size_t)i * vertex_stride;
  float logit[8][8];                         // [head][view]; n_head ≤ 8, V ≤ 8 (checked by the caller)
  float f[KV];
  for (int v = 0; v < V; ++v) {
#pragma unroll
    for (int j = 0; j < KV; ++j) f[j] = __ldg(f0 + v * view_stride + j);
    for (int h = 0; h < n_head; ++h) {
      float l = 0.f;
      for (int d = 0; d < dk; ++d) {
        const int o = h * dk + d;
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < KV; ++j) a = fmaf(sk[o * KV + j], f[j], a);
        l = fmaf(q[o], a, l);
      }
      logit[h][v] = l;
    }
  }
  for (int h = 0; h < n_head; ++h) {
    float m = logit[h][0];
    for (int v = 1; v < V; ++v) m = fmaxf(m, logit[h][v]);
    float s = 0.f;
    for (int v = 0; v < V; ++v) {
      logit[h][v] = expf(logit[h][v] - m);
      s += logit[h][v];
    }
    const float r = 1.f / s;
    for (int v = 0; v < V; ++v) logit[h][v] *= r;
  }
  float o_acc[HD];
#pragma unroll
  for (int o = 0; o < HD; ++o) o_acc[o] = 0.f;
  for (int v = 0; v < V; ++v) {
#pragma unroll
    for (int j = 0; j < KV; ++j) f[j] = __ldg(f0 + v * view_stride + j);
    for (int h = 0; h < n_head; ++h) {
      const float a_hv = logit[h][v];
      for (int d = 0; d < dk; ++d) {
        const int o = h * dk + d;
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < KV; ++j) a = fmaf(sv[o * KV + j], f[j], a);
        o_acc[o] = fmaf(a_hv, a, o_acc[o]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < DM; ++j) {
    float a = 0.f;
#pragma unroll
    for (int o = 0; o < HD; ++o) a = fmaf(sf[j * HD + o], o_acc[o], a);
    out[(size_t)i * DM + j] = a;
  }
}

}  // namespace gpnerf

using namespace gpnerf;

extern "C" {

int gpnerf_attn_smpl_code(const float* code, const float* feats, long long view_stride, long long vertex_stride, int n,
                          int n_views, const float* w_q, const float* w_k, const float* w_v, const float* w_fc,
                          int d_model, int kv_dim, int n_head, int d_k, float* out, void* stream) {
  GPNERF_REQUIRE(code && feats && w_q && w_k && w_v && w_fc && out && n > 0);
  GPNERF_REQUIRE(n_views >= 1 && n_views <= 8 && n_head >= 1 && n_head <= 8 && d_k >= 1);
  GPNERF_REQUIRE(view_stride >= kv_dim && vertex_stride >= view_stride * n_views);
  const float inv_temp = 1.f / sqrtf((float)d_k);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (n + 63) / 64;
#define GPNERF_ATTN(DM, KV, HD)                                                                                    \
  if (d_model == DM && kv_dim == KV && n_head * d_k == HD) {                                                       \
    smpl_code_attention<DM, KV, HD><<<grid, 64, 0, st>>>(code, feats, view_stride, vertex_stride, n, n_views, w_q, \
                                                         w_k, w_v, w_fc, n_head, inv_temp, out);                   \
    return check_launch("attn_smpl_code");                                                                         \
  }
  GPNERF_ATTN(16, 32, 16)      // the reference's configuration (configs/*.yaml: code_dim 16, 4 heads, 32-ch features)
  GPNERF_ATTN(32, 32, 32)
#undef GPNERF_ATTN
  set_error("attn_smpl_code: (d_model, kv_dim, n_head*d_k) must be (16,32,16) or (32,32,32)", cudaSuccess);
  return GPNERF_E_ARG;
}

}  // extern "C"
