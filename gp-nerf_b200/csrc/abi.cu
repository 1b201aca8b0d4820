// Library-wide plumbing: error reporting, device query and the ordered
// stream-compaction passes shared by K1 (pixels→rays), K2 (occupancy) and K4
// (density).  HBM-bound integer work: 1 bit/item in, 4 B/survivor out.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gpnerf {

static thread_local char g_err[256] = "";

void set_error(const char* what, cudaError_t err) {
  if (err != cudaSuccess)
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(err));
  else
    snprintf(g_err, sizeof(g_err), "invalid argument: %s", what);
}

int check_launch(const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(what, err);
    return GPNERF_E_CUDA;
  }
  return GPNERF_OK;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached = n;
  }
  return cached;
}

static inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

CompactWs carve_workspace(void* ws, int64_t n_items_max) {
  int64_t n_words = div_up(n_items_max, 32);
  int64_t n_tiles = div_up(n_words, kTileWords);
  CompactWs c;
  c.words = reinterpret_cast<uint32_t*>(ws);
  int64_t words_padded = div_up(n_words, 64) * 64;
  c.tile_sums = reinterpret_cast<int32_t*>(c.words + words_padded);
  c.tile_offs = c.tile_sums + div_up(n_tiles, 64) * 64;
  return c;
}

// --- pass B1: per-tile popcount (tile = kTileWords words, one per thread) ------
__global__ void __launch_bounds__(kTileWords) compact_tile_sums(const uint32_t* __restrict__ words,
                                                                const int32_t* n_src, int mult,
                                                                long long n_const,
                                                                int32_t* __restrict__ tile_sums) {
  const long long n_items = live_count(n_src, mult, n_const);
  const long long n_words = (n_items + 31) >> 5;
  const long long n_tiles = (n_words + kTileWords - 1) / kTileWords;
  __shared__ int warp_part[kTileWords / 32];
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long w = tile * kTileWords + threadIdx.x;
    int local = (w < n_words) ? __popc(__ldg(words + w)) : 0;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int k = 0; k < kTileWords / 32; ++k) s += warp_part[k];
      tile_sums[tile] = s;
    }
    __syncthreads();
  }
}

// --- pass B2: single-CTA exclusive scan of the tile sums -------------------
__global__ void __launch_bounds__(1024) compact_scan(const int32_t* __restrict__ tile_sums,
                                                     const int32_t* n_src, int mult,
                                                     long long n_const,
                                                     int32_t* __restrict__ tile_offs,
                                                     int32_t* __restrict__ out_count, int row_len,
                                                     int32_t* __restrict__ row_begin) {
  const long long n_items = live_count(n_src, mult, n_const);
  const long long n_words = (n_items + 31) >> 5;
  const int n_tiles = (int)((n_words + kTileWords - 1) / kTileWords);
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < n_tiles; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < n_tiles) ? tile_sums[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_tot[lane];
      int wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + warp_tot[wid] + incl - v;
    if (i < n_tiles) tile_offs[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *out_count = carry_s;
    if (row_begin != nullptr) row_begin[(n_items + row_len - 1) / row_len] = carry_s;   // CSR sentinel
  }
}

// --- pass C: expand bit words into ascending indices -----------------------
__global__ void __launch_bounds__(kTileWords) compact_expand(const uint32_t* __restrict__ words,
                                                             const int32_t* __restrict__ tile_offs,
                                                             const int32_t* n_src, int mult,
                                                             long long n_const,
                                                             int32_t* __restrict__ out_idx, int row_len,
                                                             int32_t* __restrict__ row_begin) {
  const long long n_items = live_count(n_src, mult, n_const);
  const long long n_words = (n_items + 31) >> 5;
  const long long n_tiles = (n_words + kTileWords - 1) / kTileWords;
  constexpr int NW = kTileWords / 32;
  __shared__ int word_off[kTileWords];
  __shared__ uint32_t word_bits[kTileWords];
  __shared__ int warp_tot[NW];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long w = tile * kTileWords + threadIdx.x;
    const uint32_t bits = (w < n_words) ? __ldg(words + w) : 0u;
    const int cnt = __popc(bits);
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int wbase = 0;
    for (int k = 0; k < wid; ++k) wbase += warp_tot[k];
    word_off[threadIdx.x] = tile_offs[tile] + wbase + incl - cnt;
    word_bits[threadIdx.x] = bits;
    __syncthreads();
    // a warp expands one word at a time: coalesced index writes
    for (int wl = wid; wl < kTileWords; wl += NW) {
      const uint32_t b = word_bits[wl];
      const unsigned item = (unsigned)((tile * kTileWords + wl) * 32 + lane);     // item counts stay below 2^31
      const int rank = __popc(b & ((1u << lane) - 1u));
      if (row_begin != nullptr && item < (unsigned)n_items && item % (unsigned)row_len == 0u)
        row_begin[item / (unsigned)row_len] = word_off[wl] + rank;
      if (b == 0u) continue;
      if ((b >> lane) & 1u) out_idx[word_off[wl] + rank] = (int32_t)item;
    }
    __syncthreads();
  }
}

// --- single pass: tile popcount, decoupled look-back over the tile prefixes, expansion ----------------------
// One launch instead of three (tile sums → single-CTA scan → expand): a tile (kTileWords words = 8,192 items) is
// claimed through a ticket, so tiles are processed in an order in which every predecessor has been started; the CTA
// publishes its aggregate, looks back over its predecessors' descriptors ((status << 32) | value: 1 = aggregate,
// 2 = inclusive prefix) until it meets an inclusive prefix, publishes its own and expands its words into ascending
// indices.  Descriptors and the ticket live in the workspace (tile_sums | tile_offs as one uint64 array) and are
// cleared by a memset node in front of every launch.
__global__ void __launch_bounds__(kTileWords) compact_single_pass(const uint32_t* __restrict__ words,
                                                                  unsigned long long* __restrict__ desc, int* __restrict__ ticket,
                                                                  const int32_t* n_src, int mult, long long n_const,
                                                                  int32_t* __restrict__ out_idx, int32_t* __restrict__ out_count,
                                                                  int row_len, int32_t* __restrict__ row_begin) {
  const long long n_items = live_count(n_src, mult, n_const);
  const long long n_words = (n_items + 31) >> 5;
  const int n_tiles = (int)((n_words + kTileWords - 1) / kTileWords);
  constexpr int NW = kTileWords / 32;
  __shared__ int word_off[kTileWords];
  __shared__ uint32_t word_bits[kTileWords];
  __shared__ int warp_tot[NW];
  __shared__ int tile_s, excl_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (n_tiles == 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      *out_count = 0;
      if (row_begin != nullptr) row_begin[0] = 0;
    }
    return;
  }
  for (;;) {
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1);
    __syncthreads();
    const int tile = tile_s;
    if (tile >= n_tiles) return;
    const long long w = (long long)tile * kTileWords + threadIdx.x;
    const uint32_t bits = (w < n_words) ? __ldg(words + w) : 0u;
    const int cnt = __popc(bits);
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int k = 0; k < NW; ++k) {
      if (k < wid) wbase += warp_tot[k];
      total += warp_tot[k];
    }
    // ---- publish the aggregate, look back (warp 0), publish the inclusive prefix
    if (wid == 0) {
      volatile unsigned long long* vd = desc;
      if (lane == 0) {
        const unsigned long long mine = ((unsigned long long)(tile == 0 ? 2u : 1u) << 32) | (unsigned)total;
        atomicExch(desc + tile, mine);
      }
      int excl = 0;
      if (tile > 0) {
        int look = tile - 1;
        for (;;) {          // 32 predecessors at a time, nearest first
          const int t = look - lane;
          unsigned long long d = 0;
          if (t >= 0) {
            do {
              d = vd[t];
            } while ((d >> 32) == 0ull);
          } else {
            d = 2ull << 32;               // before tile 0: an empty inclusive prefix
          }
          const unsigned full = __ballot_sync(0xffffffffu, (d >> 32) == 2ull);
          const int first_full = full ? (__ffs(full) - 1) : 32;
          int v = (lane <= first_full) ? (int)(unsigned)d : 0;
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          excl += v;
          if (full) break;
          look -= 32;
        }
        if (lane == 0) {
          __threadfence();
          atomicExch(desc + tile, (2ull << 32) | (unsigned)(excl + total));
        }
      }
      if (lane == 0) {
        excl_s = excl;
        if (tile == n_tiles - 1) {
          *out_count = excl + total;
          if (row_begin != nullptr) row_begin[(n_items + row_len - 1) / row_len] = excl + total;   // CSR sentinel
        }
      }
    }
    word_off[threadIdx.x] = wbase + incl - cnt;
    word_bits[threadIdx.x] = bits;
    __syncthreads();
    const int tile_off = excl_s;
    // ---- expansion: a warp expands one word at a time (coalesced index writes)
    for (int wl = wid; wl < kTileWords; wl += NW) {
      const uint32_t b = word_bits[wl];
      const unsigned item = (unsigned)(((long long)tile * kTileWords + wl) * 32 + lane);     // item counts stay below 2^31
      const int rank = __popc(b & ((1u << lane) - 1u));
      if (row_begin != nullptr && item < (unsigned)n_items && item % (unsigned)row_len == 0u)
        row_begin[item / (unsigned)row_len] = tile_off + word_off[wl] + rank;
      if (b == 0u) continue;
      if ((b >> lane) & 1u) out_idx[tile_off + word_off[wl] + rank] = (int32_t)item;
    }
    __syncthreads();
  }
}

int compact_launch(const CompactWs& ws, const int32_t* n_src, int mult, int64_t n_const,
                   int64_t n_items_max, int32_t* out_idx, int32_t* out_count, cudaStream_t st, int row_len,
                   int32_t* row_begin) {
  if (row_begin == nullptr || row_len <= 0) {
    row_begin = nullptr;
    row_len = 1;
  }
  if (n_src != nullptr) n_const = n_items_max;      // a device-side count is clamped to the buffers' capacity
  int64_t n_words = div_up(n_items_max, 32);
  int64_t n_tiles = div_up(n_words, kTileWords);
  int grid = (int)(n_tiles < (int64_t)sm_count() * 8 ? (n_tiles > 0 ? n_tiles : 1) : sm_count() * 8);
  static const bool three_pass = getenv("GPNERF_COMPACT_IMPL") && !strcmp(getenv("GPNERF_COMPACT_IMPL"), "three_pass");
  if (three_pass) {
    compact_tile_sums<<<grid, kTileWords, 0, st>>>(ws.words, n_src, mult, n_const, ws.tile_sums);
    compact_scan<<<1, 1024, 0, st>>>(ws.tile_sums, n_src, mult, n_const, ws.tile_offs, out_count, row_len, row_begin);
    compact_expand<<<grid, kTileWords, 0, st>>>(ws.words, ws.tile_offs, n_src, mult, n_const, out_idx, row_len,
                                                row_begin);
    return check_launch("compact");
  }
  // descriptors (one uint64 per tile, in the tile_sums | tile_offs area) + the ticket behind them
  const int64_t n_pad = div_up(n_tiles, 64) * 64;
  unsigned long long* desc = reinterpret_cast<unsigned long long*>(ws.tile_sums);
  int* ticket = reinterpret_cast<int*>(desc + n_pad);
  cudaError_t e = cudaMemsetAsync(desc, 0, (size_t)n_pad * 8 + 64, st);
  if (e != cudaSuccess) {
    set_error("memset compaction descriptors", e);
    return GPNERF_E_CUDA;
  }
  compact_single_pass<<<grid, kTileWords, 0, st>>>(ws.words, desc, ticket, n_src, mult, n_const, out_idx, out_count, row_len,
                                                   row_begin);
  return check_launch("compact");
}

}  // namespace gpnerf

extern "C" {

int gpnerf_abi_version(void) { return GPNERF_ABI_VERSION; }
const char* gpnerf_last_error(void) { return gpnerf::g_err; }
int gpnerf_sm_count(void) { return gpnerf::sm_count(); }
int gpnerf_struct_bytes(int which) {
  switch (which) {
    case 0: return (int)sizeof(gpnerf_frame_t);
    case 1: return (int)sizeof(gpnerf_head_weights_t);
    case 2: return (int)sizeof(gpnerf_peer_t);
    default: return GPNERF_E_ARG;
  }
}

int64_t gpnerf_workspace_bytes(int64_t n_items) {
  if (n_items < 0) return GPNERF_E_ARG;
  int64_t n_words = gpnerf::div_up(n_items, 32);
  int64_t n_tiles = gpnerf::div_up(n_words, gpnerf::kTileWords);
  return (gpnerf::div_up(n_words, 64) * 64 + 2 * gpnerf::div_up(n_tiles, 64) * 64 + 64 + 64) * 4;
}

}  // extern "C"
