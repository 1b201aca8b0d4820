// tcgen05 / TMEM / mbarrier / bulk-copy helpers (inline PTX, sm_100a).
//
// Operand layout used by every UMMA in this library: K-major, no swizzle
// ("interleave").  A [rows x K] bf16 operand is stored as 8x8 core matrices of
// 128 contiguous bytes (8 rows x 16 B):
//     byte(r, k) = (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2,   LBO = 128 B.
// The tiles are written by this CTA's own threads (gather results, epilogue
// activations), never by a row-major TMA box, so no swizzle is needed: a warp's
// 16-byte-per-thread stores (8 rows x 4 chunks) and the tensor core's 128-byte
// core-matrix reads are both bank-conflict free.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpnerf {
namespace tc {

constexpr uint32_t kLBO = 128;  // bytes between core matrices along K

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)      // suspend-time hint (ns): sleep, do not poll
      : "memory");
  return ok;
}
// Bounded wait: a descriptor bug must surface as a trap, not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 4000000u) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

// ---- TMA bulk copy global → shared (1-D), completes on an mbarrier ----------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- proxies / fences -------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32-bit, 16 / 32 consecutive columns: thread i of the warp receives
// columns [col, col+n) of TMEM lane (lane_base + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// tcgen05.st: thread i of the warp writes columns [col, col+n) of TMEM lane (lane_base + i).  Used for the
// activations between the layers of a head: a [128 x K] bf16 A operand lives in TMEM as K/2 32-bit columns,
// column j = (k = 2j | k = 2j+1 << 16) – the layout tcgen05.mma reads when its A operand is a TMEM address
// (verified bit-exact on a B200: tools/probes/ts_probe.cu, profiles/r02_ts_probe.txt).
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// non-blocking phase test (the MMA-issuing thread polls several barriers round robin)
__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- UMMA descriptors and issue --------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor for kind::f16: (bf16 | fp16) x (bf16 | fp16) → fp32, both K-major
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1;
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, uint32_t fmt) {
  return (1u << 4)                         // c_format  = F32
         | (fmt << 7)                      // a_format
         | (fmt << 10)                     // b_format
         | ((uint32_t)(N >> 3) << 17)      // n_dim
         | ((uint32_t)(M >> 4) << 24);     // m_dim
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) { return make_idesc(M, N, kFmtBF16); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (".ts" form): taddr_a = lane 0 / first column of the packed [128 x 16] slice
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major SWIZZLE_128B operand: rows of 128 bytes (64 16-bit elements), the 16-byte chunk c of row r stored at
// chunk position c ^ (r % 8); 8-row groups 1024 bytes apart (SBO), tile base 1024-byte aligned; a K = 16 step
// advances the start address by 32 bytes inside the swizzle atom.  The tensor core fetches a [128 x 16] slice of
// this layout in ≈70 cycles, the unswizzled core-matrix layout in ≈125 (profiles/r02_ts_probe.txt).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  return make_smem_desc(smem_addr, 16, 1024) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {       // byte offset of chunk c (0..7) of row r
  return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[128 x N] (+)= A[128 x Kp] · B[N x Kp]^T ; one thread issues Kp/16 MMAs
// and commits them to `bar`.
__device__ __forceinline__ void issue_gemm(uint32_t a_addr, uint32_t a_sbo, uint32_t b_addr, uint32_t b_sbo,
                                           int Kp, int N, uint32_t tmem_d, uint64_t* bar) {
  const uint32_t idesc = make_idesc_bf16(128, N);
  for (int k16 = 0; k16 < Kp / 16; ++k16) {
    const uint64_t ad = make_smem_desc(a_addr + k16 * 2 * kLBO, kLBO, a_sbo);
    const uint64_t bd = make_smem_desc(b_addr + k16 * 2 * kLBO, kLBO, b_sbo);
    umma_bf16(tmem_d, ad, bd, idesc, k16 > 0 ? 1u : 0u);
  }
  umma_commit(bar);
}
// Same with the layer's bias folded in: B is [N x (Kp+16)], its last K block holds the bias (hi, lo
// bf16 split) in columns 6 and 7; `ones_addr` is a [128 x 16] bf16 A block whose columns 6, 7 are 1.0
// and whose columns 8..15 are zero (columns 0..5 meet zeros in B).  `fmt` is the format of the first
// Kp columns of A and B (the bias block is always bf16).
__device__ __forceinline__ void issue_gemm_bias(uint32_t a_addr, uint32_t a_sbo, uint32_t b_addr, int Kp, int N,
                                                uint32_t ones_addr, uint32_t ones_sbo, uint32_t tmem_d,
                                                uint64_t* bar, uint32_t fmt = kFmtBF16) {
  const uint32_t idesc = make_idesc(128, N, fmt);
  const uint32_t b_sbo = (uint32_t)((Kp + 16) / 8) * kLBO;
  for (int k16 = 0; k16 < Kp / 16; ++k16) {
    const uint64_t ad = make_smem_desc(a_addr + k16 * 2 * kLBO, kLBO, a_sbo);
    const uint64_t bd = make_smem_desc(b_addr + k16 * 2 * kLBO, kLBO, b_sbo);
    umma_bf16(tmem_d, ad, bd, idesc, k16 > 0 ? 1u : 0u);
  }
  umma_bf16(tmem_d, make_smem_desc(ones_addr, kLBO, ones_sbo),
            make_smem_desc(b_addr + (Kp / 16) * 2 * kLBO, kLBO, b_sbo), make_idesc_bf16(128, N), 1u);
  umma_commit(bar);
}

// byte offset of the 16-byte chunk holding k = 8*kc .. 8*kc+7 of row r
__device__ __forceinline__ uint32_t chunk_off(int r, int kc, uint32_t sbo) {
  return (uint32_t)(r >> 3) * sbo + (uint32_t)kc * kLBO + (uint32_t)(r & 7) * 16;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_chunk(uint8_t* tile, uint32_t off, const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile + off) = q;
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_chunk_f16(uint8_t* tile, uint32_t off, const float (&v)[8]) {
  uint4 q;
  q.x = pack_f16x2(v[0], v[1]);
  q.y = pack_f16x2(v[2], v[3]);
  q.z = pack_f16x2(v[4], v[5]);
  q.w = pack_f16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile + off) = q;
}

// ELU for activations that are about to be rounded to bf16
// (one FMUL + MUFU.EX2 with flush-to-zero: the denormal-preserving __expf costs three more
// instructions per element, and these epilogues are issue bound)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float elu_fast(float x) {
  const float e = ex2_ftz(x * 1.4426950408889634f) - 1.0f;
  return x > 0.0f ? x : e;
}
// The head kernels keep every hidden activation multiplied by c = log2(e): the first layer's
// weights and all biases are packed pre-scaled (the bias sits in the B operand, see
// issue_gemm_bias), so the accumulator already holds y = c·(W·a + b) and
//     c·ELU(y / c) = y > 0 ? y : c·2^y − c
// is MUFU.EX2 + FFMA + compare/select: no bias add, no pre-multiply.  The last (CUDA-core) layer of
// each head is packed with its weights divided by c.
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float elu_scaled(float y) {
  const float e = fmaf(ex2_ftz(y), kLog2e, -kLog2e);
  return y > 0.0f ? y : e;
}
__device__ __forceinline__ float sigmoid_fast(float x) {
  return __frcp_rn(1.0f + ex2_ftz(-1.4426950408889634f * x));
}

}  // namespace tc
}  // namespace gpnerf
