// Probe (not part of the library): does tcgen05.mma accept its A operand from TMEM in the layout
// "lane = row, 32-bit column j = (k = 2j | k = 2j+1 << 16)", written by tcgen05.st.32x32b?  And what is the
// issue → commit → wake-up round trip of a small GEMM?  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -I gp-nerf_b200/csrc -I include -o gpurun_out/ts_probe tools/probes/ts_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_common.cuh"

using namespace gpnerf::tc;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr int K = 32, N = 16;
__device__ __host__ inline float aval(int r, int k) { return (float)((r * 7 + k * 3) % 17 - 8) * 0.125f; }
__device__ __host__ inline float bval(int n, int k) { return (float)((n * 5 + k * 11) % 13 - 6) * 0.25f; }

// mode 0: A from TMEM (TS); mode 1: A from shared memory (SS, the library's layout) – the control
__global__ void __launch_bounds__(128) probe(int mode, int n_mma_rep, float* out, long long* cycles) {
  __shared__ __align__(128) uint8_t sB[N * K * 2];
  __shared__ __align__(128) uint8_t sA[128 * K * 2];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 128);
  for (int i = tid; i < N * K; i += 128) {
    int n = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + chunk_off(n, k >> 3, (K / 8) * kLBO) + (k & 7) * 2) = __float2bfloat16_rn(bval(n, k));
  }
  for (int k = 0; k < K; ++k)
    *reinterpret_cast<__nv_bfloat16*>(sA + chunk_off(tid, k >> 3, (K / 8) * kLBO) + (k & 7) * 2) = __float2bfloat16_rn(aval(tid, k));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  // A into TMEM columns 64..79: 16 packed pairs per row
  uint32_t pk[16];
  for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(aval(tid, 2 * j), aval(tid, 2 * j + 1));
  tmem_st16(t_row + 64, pk);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    t0 = clock64();
    for (int rep = 0; rep < n_mma_rep; ++rep)
      for (int k16 = 0; k16 < K / 16; ++k16) {
        const uint64_t bd = make_smem_desc(smem_u32(sB) + k16 * 2 * kLBO, kLBO, (K / 8) * kLBO);
        if (mode == 0)
          umma_ts(tmem, tmem + 64 + k16 * 8, bd, idesc, (k16 > 0 || rep > 0) ? 1u : 0u);
        else
          umma_bf16(tmem, make_smem_desc(smem_u32(sA) + k16 * 2 * kLBO, kLBO, (K / 8) * kLBO), bd, idesc,
                    (k16 > 0 || rep > 0) ? 1u : 0u);
      }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
  uint32_t r[16];
  tmem_ld16(t_row, r);
  tmem_wait_ld();
  for (int n = 0; n < N; ++n) out[tid * N + n] = __uint_as_float(r[n]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
  float* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 128 * N * 4); cudaMalloc(&d_cyc, 8);
  std::vector<float> ref(128 * N), got(128 * N);
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += aval(r, k) * bval(n, k); ref[r * N + n] = s; }
  for (int mode = 0; mode < 2; ++mode)
    for (int rep : {1, 4, 16}) {
      long long cyc = 0;
      for (int it = 0; it < 3; ++it) probe<<<1, 128>>>(mode, rep, d_out, d_cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d rep %d: CUDA error %s\n", mode, rep, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(got.data(), d_out, 128 * N * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
      double err = 0;
      for (int i = 0; i < 128 * N; ++i) err = fmax(err, fabs(got[i] - rep * ref[i]));
      printf("mode %s  %2d x K=32 (N=16)  max|err| = %g   issue->wake %lld cycles  (got[5]=%g ref=%g)\n", mode == 0 ? "TS" : "SS", rep, err, cyc,
             got[5], rep * ref[5]);
    }
  return 0;
}
