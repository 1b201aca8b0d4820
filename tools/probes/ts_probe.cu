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

constexpr int K = 32, N = 16, NB = 64;     // NB: rows of the B operand in shared memory (N <= NB)
__device__ __host__ inline float aval(int r, int k) { return (float)((r * 7 + k * 3) % 17 - 8) * 0.125f; }
__device__ __host__ inline float bval(int n, int k) { return (float)((n * 5 + k * 11) % 13 - 6) * 0.25f; }

// mode 0: A from TMEM (TS); mode 1: A from shared memory (SS, the library's layout) – the control
// mode 2: as mode 0 but the repetitions alternate between two accumulators (is the per-MMA cost a dependency
// latency on the accumulator or an issue interval?); n_cols = N of the instruction (16 or 64)
__global__ void __launch_bounds__(128) probe(int mode, int n_mma_rep, int n_cols, float* out, long long* cycles) {
  __shared__ __align__(128) uint8_t sB[NB * K * 2];
  __shared__ __align__(128) uint8_t sA[128 * K * 2];
  __shared__ __align__(1024) uint8_t sAsw[128 * 128];      // mode 3: rows of 128 B (64 bf16), 16-byte chunks XOR (row % 8)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 256);
  for (int i = tid; i < NB * K; i += 128) {
    int n = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + chunk_off(n, k >> 3, (K / 8) * kLBO) + (k & 7) * 2) = __float2bfloat16_rn(bval(n, k));
  }
  for (int k = 0; k < K; ++k)
    *reinterpret_cast<__nv_bfloat16*>(sA + chunk_off(tid, k >> 3, (K / 8) * kLBO) + (k & 7) * 2) = __float2bfloat16_rn(aval(tid, k));
  for (int k = 0; k < 64; ++k)
    *reinterpret_cast<__nv_bfloat16*>(sAsw + (tid >> 3) * 1024 + (tid & 7) * 128 + (((k >> 3) ^ (tid & 7)) << 4) + (k & 7) * 2) =
        __float2bfloat16_rn(k < K ? aval(tid, k) : 0.f);
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
  tmem_st16(t_row + 96, pk);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n_cols);
    t0 = clock64();
    for (int rep = 0; rep < n_mma_rep; ++rep)
      for (int k16 = 0; k16 < K / 16; ++k16) {
        const uint64_t bd = make_smem_desc(smem_u32(sB) + k16 * 2 * kLBO, kLBO, (K / 8) * kLBO);
        if (mode == 0)
          umma_ts(tmem, tmem + 64 + k16 * 8, bd, idesc, (k16 > 0 || rep > 0) ? 1u : 0u);
        else if (mode == 3) {
          // K-major SWIZZLE_128B: SBO = 1024 B between 8-row groups, LBO unused (1), layout_type 2 in bits 61..63;
          // a K = 16 step advances the start address by 32 B inside the swizzle atom
          uint64_t ad = make_smem_desc(smem_u32(sAsw) + k16 * 32, 16, 1024) | ((uint64_t)2 << 61);
          umma_bf16(tmem, ad, bd, idesc, (k16 > 0 || rep > 0) ? 1u : 0u);
        } else if (mode == 2)
          umma_ts(tmem + (rep & 1) * 128, tmem + 64 + 32 + k16 * 8, bd, idesc, (k16 > 0 || rep > 1) ? 1u : 0u);
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
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
  float* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 128 * N * 4); cudaMalloc(&d_cyc, 8);
  std::vector<float> ref(128 * N), got(128 * N);
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += aval(r, k) * bval(n, k); ref[r * N + n] = s; }
  for (int mode = 0; mode < 4; ++mode)
   for (int ncols : {16, 64})
    for (int rep : {1, 4, 16}) {
      long long cyc = 0;
      for (int it = 0; it < 3; ++it) probe<<<1, 128>>>(mode, rep, ncols, d_out, d_cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d rep %d: CUDA error %s\n", mode, rep, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(got.data(), d_out, 128 * N * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
      double err = 0;
      const int scale = mode == 2 ? (rep + 1) / 2 : rep;     // mode 2: accumulator 0 holds every second repetition
      for (int i = 0; i < 128 * N; ++i) err = fmax(err, fabs(got[i] - scale * ref[i]));
      printf("mode %s  %2d x K=32 (N=%d)  max|err| = %g   issue->wake %lld cycles\n",
             mode == 0 ? "TS" : (mode == 1 ? "SS" : (mode == 2 ? "TS, two accumulators" : "SS, A in SWIZZLE_128B")), rep, ncols, err, cyc);
    }
  return 0;
}
