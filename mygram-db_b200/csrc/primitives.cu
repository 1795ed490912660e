// primitives.cu — device-wide building blocks written for this library:
// an exclusive scan (u32 -> u64) and a stable LSD radix sort of (u64, u32) pairs.
//
// Both are HBM-bound streaming kernels: grids are sized from the data (many
// waves over 148 SMs), every global access is a coalesced full-warp access on
// the read side, and the only shared-memory traffic is per-CTA histograms.
#include "mgx_internal.cuh"

namespace mgx {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;                          // per thread
constexpr int kScanTile = kScanThreads * kScanItems;    // 4096 per CTA

__device__ __forceinline__ uint64_t warp_inclusive_scan_u64(uint64_t v) {
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= static_cast<unsigned>(d)) {
      v += o;
    }
  }
  return v;
}

// Block-wide exclusive scan of one u64 per thread; returns the exclusive prefix
// and (to every thread) the block total. smem: one u64 per warp.
template <int THREADS>
__device__ __forceinline__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t* total, uint64_t* smem) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const uint64_t inc = warp_inclusive_scan_u64(v);
  if (lane == 31) {
    smem[warp] = inc;
  }
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < THREADS / 32 ? smem[lane] : 0;
    w = warp_inclusive_scan_u64(w);
    if (lane < THREADS / 32) {
      smem[lane] = w;
    }
  }
  __syncthreads();
  const uint64_t warp_prefix = warp == 0 ? 0 : smem[warp - 1];
  *total = smem[THREADS / 32 - 1];
  __syncthreads();
  return warp_prefix + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                                   uint64_t* __restrict__ block_sums) {
  __shared__ uint64_t smem[kScanThreads / 32];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile;
  uint64_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t i = base + static_cast<uint64_t>(k) * kScanThreads + threadIdx.x;
    if (i < n) {
      sum += in[i];
    }
  }
  uint64_t total;
  (void)block_exclusive_scan_u64<kScanThreads>(sum, &total, smem);
  if (threadIdx.x == 0) {
    block_sums[blockIdx.x] = total;
  }
}

// One CTA turns the block sums into exclusive block offsets (in place) and
// writes the grand total to *total_out.
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(uint64_t* __restrict__ block_sums, uint64_t n_blocks,
                                                               uint64_t* __restrict__ total_out) {
  __shared__ uint64_t smem[32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0) {
    carry_s = 0;
  }
  __syncthreads();
  for (uint64_t base = 0; base < n_blocks; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    const uint64_t v = i < n_blocks ? block_sums[i] : 0;
    uint64_t total;
    const uint64_t ex = block_exclusive_scan_u64<1024>(v, &total, smem);
    const uint64_t carry = carry_s;
    if (i < n_blocks) {
      block_sums[i] = carry + ex;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      carry_s = carry + total;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *total_out = carry_s;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                                  const uint64_t* __restrict__ block_offsets,
                                                                  uint64_t* __restrict__ out) {
  __shared__ uint64_t smem[kScanThreads / 32];
  // Thread t owns kScanItems CONSECUTIVE items so the scan order is the array order.
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
  uint32_t v[kScanItems];
  uint64_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t i = base + k;
    v[k] = i < n ? in[i] : 0u;
    sum += v[k];
  }
  uint64_t total;
  uint64_t prefix = block_exclusive_scan_u64<kScanThreads>(sum, &total, smem) + block_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t i = base + k;
    if (i < n) {
      out[i] = prefix;
    }
    prefix += v[k];
  }
}

// ------------------------------------------------------------------ radix sort
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 16;                                // items per thread
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortTile = kSortThreads * kSortRounds;          // 4096 items per CTA
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n,
                                                                  int shift, uint32_t mask, uint32_t* __restrict__ hist,
                                                                  uint32_t n_tiles) {
  __shared__ uint32_t bins[kRadix];
  bins[threadIdx.x] = 0;  // kSortThreads == kRadix
  __syncthreads();
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kSortTile;
#pragma unroll 4
  for (int r = 0; r < kSortRounds; ++r) {
    const uint64_t i = base + static_cast<uint64_t>(r) * kSortThreads + threadIdx.x;
    if (i < n) {
      atomicAdd(&bins[static_cast<uint32_t>(keys[i] >> shift) & mask], 1u);
    }
  }
  __syncthreads();
  hist[static_cast<uint64_t>(threadIdx.x) * n_tiles + blockIdx.x] = bins[threadIdx.x];
}

// Global digit histograms of every pass in ONE read of the keys: a pass whose digit is the same for all
// keys is a pure copy and is skipped (e.g. the high bits of each 21-bit code-point field).
constexpr int kMaxPasses = 9;
struct PassList {
  int n;
  int shift[kMaxPasses];
  uint32_t mask[kMaxPasses];
};
__global__ void __launch_bounds__(256) radix_global_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n,
                                                                PassList pl,
                                                                unsigned long long* __restrict__ hist /*[passes][256]*/) {
  const int passes = pl.n;
  __shared__ uint32_t bins[kMaxPasses][kRadix];
  for (int p = 0; p < passes; ++p) {
    bins[p][threadIdx.x] = 0;
  }
  __syncthreads();
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t k = keys[i];
    for (int p = 0; p < passes; ++p) {
      atomicAdd(&bins[p][static_cast<uint32_t>(k >> pl.shift[p]) & pl.mask[p]], 1u);
    }
  }
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    const uint32_t c = bins[p][threadIdx.x];
    if (c != 0) {
      atomicAdd(&hist[p * kRadix + threadIdx.x], static_cast<unsigned long long>(c));
    }
  }
}

// Stable scatter with a shared-memory reorder. Warp w owns the contiguous items [w*512, (w+1)*512) of the tile,
// visited in 16 rounds of 32 consecutive items, so (warp, round, lane) order == input order; ranks are assigned in
// that order. Items are first placed at their tile-local sorted position in shared memory; the tile is then written
// out run by run, so consecutive threads store to consecutive global addresses.
struct ScatterSmem {
  uint64_t keys[kSortTile];
  uint32_t vals[kSortTile];
  uint32_t warp_cnt[kSortWarps][kRadix];
  uint32_t local_start[kRadix];
  uint64_t global_base[kRadix];  // 64-bit: a pass over more than 2^32 pairs must not wrap
  uint32_t scan_tmp[kSortWarps];
};

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                                     const uint32_t* __restrict__ vals_in,
                                                                     uint64_t* __restrict__ keys_out,
                                                                     uint32_t* __restrict__ vals_out, uint64_t n,
                                                                     int shift, uint32_t mask,
                                                                     const uint64_t* __restrict__ hist_scan,
                                                                     uint32_t n_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ScatterSmem& sm = *reinterpret_cast<ScatterSmem*>(smem_raw);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) {
    sm.warp_cnt[w][threadIdx.x] = 0;
  }
  __syncthreads();

  const uint64_t tile_base = static_cast<uint64_t>(blockIdx.x) * kSortTile;
  const uint32_t tile_n = static_cast<uint32_t>(n - tile_base < kSortTile ? n - tile_base : kSortTile);
  const uint32_t warp_off = warp * (32 * kSortRounds);
  uint64_t key[kSortRounds];
  uint32_t val[kSortRounds];
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const uint32_t i = warp_off + r * 32 + lane;
    const bool valid = i < tile_n;
    key[r] = valid ? keys_in[tile_base + i] : 0;
    val[r] = valid ? vals_in[tile_base + i] : 0;
  }
  // phase 1: per-warp digit counts
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const bool valid = warp_off + r * 32 + lane < tile_n;
    const uint32_t d = valid ? (static_cast<uint32_t>(key[r] >> shift) & mask) : 0x1FFu;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (valid && (peers & lt_mask) == 0) {
      sm.warp_cnt[warp][d] += __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  // per digit (thread d): tile count, exclusive scan over digits -> tile-local start; per-warp offsets
  {
    const unsigned d = threadIdx.x;
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      cnt += sm.warp_cnt[w][d];
    }
    uint32_t inc = cnt;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, s);
      if (lane >= static_cast<unsigned>(s)) {
        inc += o;
      }
    }
    if (lane == 31) {
      sm.scan_tmp[warp] = inc;
    }
    __syncthreads();
    uint32_t prefix = 0;
    for (unsigned w = 0; w < warp; ++w) {
      prefix += sm.scan_tmp[w];
    }
    const uint32_t start = prefix + inc - cnt;
    sm.local_start[d] = start;
    sm.global_base[d] = hist_scan[static_cast<uint64_t>(d) * n_tiles + blockIdx.x];
    uint32_t running = start;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = sm.warp_cnt[w][d];
      sm.warp_cnt[w][d] = running;
      running += c;
    }
  }
  __syncthreads();
  // phase 2: rank inside the tile and place into shared memory
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const bool valid = warp_off + r * 32 + lane < tile_n;
    const uint32_t d = valid ? (static_cast<uint32_t>(key[r] >> shift) & mask) : 0x1FFu;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    uint32_t base = 0;
    if (valid) {
      base = sm.warp_cnt[warp][d];
    }
    __syncwarp();
    if (valid && (peers & lt_mask) == 0) {
      sm.warp_cnt[warp][d] = base + __popc(peers);
    }
    __syncwarp();
    if (valid) {
      const uint32_t pos = base + __popc(peers & lt_mask);
      sm.keys[pos] = key[r];
      sm.vals[pos] = val[r];
    }
  }
  __syncthreads();
  // phase 3: coalesced write-out, run by run
  for (uint32_t i = threadIdx.x; i < tile_n; i += kSortThreads) {
    const uint64_t k = sm.keys[i];
    const uint32_t d = static_cast<uint32_t>(k >> shift) & mask;
    const uint64_t pos = sm.global_base[d] + (i - sm.local_start[d]);
    keys_out[pos] = k;
    vals_out[pos] = sm.vals[i];
  }
}


// ------------------------------------------------------------------ one-sweep pass (decoupled look-back + bulk copies)
// One kernel per pass instead of histogram -> scan -> scatter: a pass reads every pair once and writes it once.
//  * The digit histograms of ALL passes come from one read of the keys (radix_global_hist_kernel); their exclusive
//    scans (digit_scan_kernel) are the global start of every digit.
//  * Persistent CTAs (two per SM) claim tiles in order from a device counter. A tile publishes its 256 digit counts
//    (flag AGGREGATE), walks back over its predecessors' entries until it meets an inclusive PREFIX, and publishes its
//    own inclusive prefix: the classic decoupled look-back, one chain per digit, one thread per chain. Tiles are
//    claimed in increasing order by resident CTAs, so the tile a chain waits for is always running or finished.
//  * The tile's keys and values arrive by cp.async.bulk (the TMA unit's 1-D bulk copy) into a two-stage shared-memory
//    ring, completion signalled on an mbarrier: the copy of tile i+1 runs while tile i is ranked and written. A stage
//    is read into registers once and then re-used as the reorder buffer of the same tile.
constexpr int kOsStages = 2;
constexpr unsigned long long kOsFlagAggregate = 1ULL << 62;
constexpr unsigned long long kOsFlagPrefix = 2ULL << 62;
constexpr unsigned long long kOsValueMask = (1ULL << 54) - 1;  // bits 54..61: pass tag (entries of other passes read as empty)

struct OsSmem {
  alignas(128) uint64_t keys[kOsStages][kSortTile];
  alignas(128) uint32_t vals[kOsStages][kSortTile];
  uint32_t warp_cnt[kSortWarps][kRadix];
  uint64_t global_base[kRadix];
  uint32_t local_start[kRadix];
  uint32_t scan_tmp[kSortWarps];
  alignas(8) uint64_t bar[kOsStages];
  uint32_t tile_id[kOsStages];
  uint32_t bulk[kOsStages];  // 1: the stage is filled by a bulk copy (full tiles), 0: read straight from global
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src_global), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

__global__ void __launch_bounds__(kRadix) digit_scan_kernel(const unsigned long long* __restrict__ hist,
                                                            unsigned long long* __restrict__ base) {
  __shared__ uint64_t smem[kRadix / 32];
  const unsigned long long v = hist[blockIdx.x * kRadix + threadIdx.x];
  uint64_t total;
  base[blockIdx.x * kRadix + threadIdx.x] = block_exclusive_scan_u64<kRadix>(v, &total, smem);
}

// Lanes of the warp whose digit equals mine. match.any walks the distinct values of the warp one after the other
// (about 30 of them when an 8-bit digit is uniformly distributed); eight ballots cost the same whatever the digits are.
// The host picks per pass: BALLOT when a warp's 32 digits are expected to hold more than kBallotDistinct different
// values (measured at 10M documents: uniform low bytes 8.9 -> 7.9 ms per pass with ballots, the skewed high bytes of
// CJK code points 5.6 -> 7.5 ms, so those keep match.any).
template <bool BALLOT>
__device__ __forceinline__ unsigned digit_peers(uint32_t d) {
  if (!BALLOT) {
    return __match_any_sync(0xffffffffu, d);
  }
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 9; ++b) {  // 8 digit bits + the "invalid item" bit
    const unsigned set = __ballot_sync(0xffffffffu, (d >> b) & 1u);
    peers &= ((d >> b) & 1u) ? set : ~set;
  }
  return peers;
}
constexpr double kBallotDistinct = 20.0;
constexpr int kLookBackWindow = 4;  // predecessor entries a look-back fetches side by side

#ifdef MGX_OS_TIMING
__device__ unsigned long long g_os_timing[8];
#define OS_T(i)                                         \
  if (threadIdx.x == 0) {                               \
    const long long now_ = clock64();                   \
    atomicAdd(&g_os_timing[i], static_cast<unsigned long long>(now_ - t_last_)); \
    t_last_ = now_;                                     \
  }
#else
#define OS_T(i)
#endif

template <bool BALLOT>
__global__ void __launch_bounds__(kSortThreads, 2)
radix_onesweep_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                      uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint64_t n, int shift,
                      uint32_t mask, const unsigned long long* __restrict__ digit_base,
                      unsigned long long* __restrict__ tile_state, uint32_t* __restrict__ tile_counter,
                      uint32_t n_tiles, unsigned long long pass_tag) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  OsSmem& sm = *reinterpret_cast<OsSmem*>(smem_raw);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint32_t warp_off = warp * (32 * kSortRounds);

  // thread 0 claims tiles and starts their copies
  auto claim_and_load = [&](int stage) {
    const uint32_t id = atomicAdd(tile_counter, 1u);
    sm.tile_id[stage] = id;
    sm.bulk[stage] = 0;
    if (id < n_tiles) {
      const uint64_t base = static_cast<uint64_t>(id) * kSortTile;
      if (n - base >= kSortTile) {
        sm.bulk[stage] = 1;
        // the stage was last written through the generic proxy (the reorder of an earlier tile)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&sm.bar[stage], kSortTile * 12u);
        bulk_load(sm.keys[stage], keys_in + base, kSortTile * 8u, &sm.bar[stage]);
        bulk_load(sm.vals[stage], vals_in + base, kSortTile * 4u, &sm.bar[stage]);
      }
    }
  };
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    claim_and_load(0);
  }
  __syncthreads();
  uint32_t phases = 0;  // bit s: parity the next wait on stage s uses
#ifdef MGX_OS_TIMING
  long long t_last_ = clock64();
#endif

  for (int it = 0;; ++it) {
    const int st = it & 1;
    const uint32_t tile = sm.tile_id[st];
    if (tile >= n_tiles) {
      break;
    }
    if (threadIdx.x == 0) {
      claim_and_load(st ^ 1);  // the other stage was released by the barrier that ended the previous round
    }
    const uint64_t tile_base = static_cast<uint64_t>(tile) * kSortTile;
    const uint32_t tile_n = static_cast<uint32_t>(n - tile_base < kSortTile ? n - tile_base : kSortTile);
    const bool bulk = sm.bulk[st] != 0;
    uint64_t key[kSortRounds];
    uint32_t val[kSortRounds];
    if (bulk) {
      mbar_wait(&sm.bar[st], (phases >> st) & 1u);
      phases ^= 1u << st;
#pragma unroll
      for (int r = 0; r < kSortRounds; ++r) {
        const uint32_t i = warp_off + r * 32 + lane;
        key[r] = sm.keys[st][i];
        val[r] = sm.vals[st][i];
      }
    } else {
#pragma unroll
      for (int r = 0; r < kSortRounds; ++r) {
        const uint32_t i = warp_off + r * 32 + lane;
        const bool valid = i < tile_n;
        key[r] = valid ? keys_in[tile_base + i] : 0;
        val[r] = valid ? vals_in[tile_base + i] : 0;
      }
    }
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      sm.warp_cnt[w][threadIdx.x] = 0;
    }
    __syncthreads();  // counters cleared; every thread holds its items, so the stage can take the reordered tile
    OS_T(0)  // wait for the tile + registers

    // phase 1: per-warp digit counts
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const bool valid = warp_off + r * 32 + lane < tile_n;
      const uint32_t d = valid ? (static_cast<uint32_t>(key[r] >> shift) & mask) : 0x1FFu;
      const unsigned peers = digit_peers<BALLOT>(d);
      if (valid && (peers & lt_mask) == 0) {
        sm.warp_cnt[warp][d] += __popc(peers);
      }
      __syncwarp();
    }
    __syncthreads();
    OS_T(1)  // phase 1
    uint32_t my_cnt = 0;
    {
      // thread d: tile count of digit d (published at once as this tile's AGGREGATE), the tile-local start and the
      // per-warp offsets
      const unsigned d = threadIdx.x;
      uint32_t cnt = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        cnt += sm.warp_cnt[w][d];
      }
      volatile unsigned long long* const my_state = tile_state + static_cast<uint64_t>(tile) * kRadix + d;
      my_cnt = cnt;
      if (tile + 1 < n_tiles) {  // nobody looks back at the last tile
        *my_state = kOsFlagAggregate | pass_tag | cnt;
      }
      uint32_t inc = cnt;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= static_cast<unsigned>(s)) {
          inc += o;
        }
      }
      if (lane == 31) {
        sm.scan_tmp[warp] = inc;
      }
      __syncthreads();
      uint32_t prefix = 0;
      for (unsigned w = 0; w < warp; ++w) {
        prefix += sm.scan_tmp[w];
      }
      const uint32_t start = prefix + inc - cnt;
      sm.local_start[d] = start;
      uint32_t running = start;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t c = sm.warp_cnt[w][d];
        sm.warp_cnt[w][d] = running;
        running += c;
      }
    }
    __syncthreads();
    OS_T(2)  // digit scan
    // phase 2: rank inside the tile and place into the stage. Nothing here needs the global digit starts, so the
    // predecessors get this long to publish before the look-back below starts waiting for them; the entries of the
    // nearest kLookBackWindow predecessors are fetched NOW, side by side, and looked at after the ranking (a
    // look-back step is a dependent L2 round trip: one at a time they were 30 % of a pass's stall samples).
    unsigned long long lb_pre[kLookBackWindow];
#pragma unroll
    for (int j = 0; j < kLookBackWindow; ++j) {
      lb_pre[j] = tile > static_cast<uint32_t>(j)
                      ? *reinterpret_cast<volatile const unsigned long long*>(tile_state + static_cast<uint64_t>(tile - 1 - j) * kRadix +
                                                                           threadIdx.x)
                      : 0ULL;
    }
    uint64_t* const skeys = sm.keys[st];
    uint32_t* const svals = sm.vals[st];
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const bool valid = warp_off + r * 32 + lane < tile_n;
      const uint32_t d = valid ? (static_cast<uint32_t>(key[r] >> shift) & mask) : 0x1FFu;
      const unsigned peers = digit_peers<BALLOT>(d);
      uint32_t base = 0;
      if (valid) {
        base = sm.warp_cnt[warp][d];
      }
      __syncwarp();
      if (valid && (peers & lt_mask) == 0) {
        sm.warp_cnt[warp][d] = base + __popc(peers);
      }
      __syncwarp();
      if (valid) {
        const uint32_t pos = base + __popc(peers & lt_mask);
        skeys[pos] = key[r];
        svals[pos] = val[r];
      }
    }
    OS_T(3)  // phase 2 (thread 0's share)
    {
      // look back: thread d sums the counts of digit d over the preceding tiles until it meets an inclusive prefix
      const unsigned d = threadIdx.x;
      const uint32_t cnt = my_cnt;
      volatile unsigned long long* const my_state = tile_state + static_cast<uint64_t>(tile) * kRadix + d;
      unsigned long long before = 0;
      uint32_t p = tile;  // the next entry to look at is p - 1
      bool found = tile == 0;
      while (!found) {
#pragma unroll
        for (int j = 0; j < kLookBackWindow; ++j) {
          if (!found) {
            --p;
            volatile const unsigned long long* const ps = tile_state + static_cast<uint64_t>(p) * kRadix + d;
            unsigned long long v = lb_pre[j];
            while ((v & (3ULL << 62)) == 0 || (v & (0xFFULL << 54)) != pass_tag) {
              __nanosleep(20);
              v = *ps;
            }
            before += v & kOsValueMask;
            found = (v & kOsFlagPrefix) != 0 || p == 0;
          }
        }
        if (!found) {  // next window, again side by side
#pragma unroll
          for (int j = 0; j < kLookBackWindow; ++j) {
            lb_pre[j] = p > static_cast<uint32_t>(j)
                            ? *reinterpret_cast<volatile const unsigned long long*>(tile_state + static_cast<uint64_t>(p - 1 - j) * kRadix + d)
                            : 0ULL;
          }
        }
      }
      if (tile + 1 < n_tiles) {
        *my_state = kOsFlagPrefix | pass_tag | (before + cnt);
      }
      sm.global_base[d] = digit_base[d] + before - sm.local_start[d];  // position of tile item i: global_base[d] + i
    }
    OS_T(4)  // look-back (thread 0's digit)
    __syncthreads();
    OS_T(5)  // waiting for the other digits' look-backs
    // phase 3: coalesced write-out, run by run
    for (uint32_t i = threadIdx.x; i < tile_n; i += kSortThreads) {
      const uint64_t k = skeys[i];
      const uint32_t d = static_cast<uint32_t>(k >> shift) & mask;
      const uint64_t pos = sm.global_base[d] + i;
      keys_out[pos] = k;
      vals_out[pos] = svals[i];
    }
    __syncthreads();  // the stage is free: the next round's claim may start a copy into it
    OS_T(6)  // phase 3
  }
}

}  // namespace

namespace {
// Up to three short independent scans in ONE launch: CTA i walks array i in chunks of 4096 with a running carry.
// The planning stage of a query batch scans three arrays of ~10^4 elements; as three-kernel device-wide scans
// they cost nine launches of a few microseconds each.
constexpr int kSmallScanThreads = 1024;
__global__ void __launch_bounds__(kSmallScanThreads) small_scans_kernel(SmallScanJobs jobs) {
  __shared__ uint64_t s_warp[kSmallScanThreads / 32];
  __shared__ uint64_t s_carry;
  const uint32_t* __restrict__ in = jobs.in[blockIdx.x];
  uint64_t* __restrict__ out = jobs.out[blockIdx.x];
  const uint64_t n = jobs.n[blockIdx.x];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    s_carry = 0;
  }
  __syncthreads();
  for (uint64_t base = 0; base < n; base += 4ULL * kSmallScanThreads) {
    const uint64_t i0 = base + 4ULL * threadIdx.x;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = i0 + k < n ? in[i0 + k] : 0u;
    }
    const uint64_t mine = static_cast<uint64_t>(v[0]) + v[1] + v[2] + v[3];
    uint64_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint64_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= static_cast<unsigned>(d)) {
        inc += o;
      }
    }
    if (lane == 31) {
      s_warp[warp] = inc;
    }
    __syncthreads();
    uint64_t prefix = s_carry;
    for (unsigned w = 0; w < warp; ++w) {
      prefix += s_warp[w];
    }
    uint64_t run = prefix + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) {
        out[i0 + k] = run;
      }
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == kSmallScanThreads - 1) {
      s_carry = prefix + inc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[n] = s_carry;
  }
}
}  // namespace

void exclusive_scans_small(const SmallScanJobs& jobs, int n_jobs, cudaStream_t stream) {
  if (n_jobs <= 0) {
    return;
  }
  small_scans_kernel<<<static_cast<unsigned>(n_jobs), kSmallScanThreads, 0, stream>>>(jobs);
  MGX_LAUNCH_CHECK();
}

size_t scan_scratch_elems(uint64_t n) { return static_cast<size_t>((n + kScanTile - 1) / kScanTile + 2); }

void exclusive_scan_u32_u64(const uint32_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_scratch,
                            cudaStream_t stream) {
  // out[n] (the total) is written by the block-sums kernel
  const uint64_t n_blocks = (n + kScanTile - 1) / kScanTile;
  uint64_t* d_block_sums = d_scratch;
  if (n_blocks > 0) {
    scan_reduce_kernel<<<static_cast<unsigned>(n_blocks), kScanThreads, 0, stream>>>(d_in, n, d_block_sums);
    MGX_LAUNCH_CHECK();
  }
  scan_block_sums_kernel<<<1, 1024, 0, stream>>>(d_block_sums, n_blocks, d_out + n);
  MGX_LAUNCH_CHECK();
  if (n_blocks > 0) {
    scan_apply_kernel<<<static_cast<unsigned>(n_blocks), kScanThreads, 0, stream>>>(d_in, n, d_block_sums, d_out);
    MGX_LAUNCH_CHECK();
  }
}

namespace {
struct SortScratch {
  size_t hist_off, hist_scan_off, ghist_off, scan_off, dbase_off, counter_off, state_off, total;
};
bool sort_classic() {  // MGX_SORT=classic: the three-kernel passes (A/B measurements)
  const char* e = std::getenv("MGX_SORT");
  return e != nullptr && std::strcmp(e, "classic") == 0;
}
SortScratch sort_scratch_layout(uint64_t n) {
  const uint64_t n_tiles = (n + kSortTile - 1) / kSortTile;
  const uint64_t hist_len = static_cast<uint64_t>(kRadix) * n_tiles;
  SortScratch L{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~static_cast<size_t>(255);
    return at;
  };
  L.hist_off = take(hist_len * sizeof(uint32_t));
  L.hist_scan_off = take((hist_len + 1) * sizeof(uint64_t));
  L.ghist_off = take(static_cast<size_t>(kMaxPasses) * kRadix * sizeof(unsigned long long));
  L.scan_off = take(scan_scratch_elems(hist_len) * sizeof(uint64_t));
  L.dbase_off = take(static_cast<size_t>(kMaxPasses) * kRadix * sizeof(unsigned long long));
  L.counter_off = take(static_cast<size_t>(kMaxPasses) * sizeof(uint32_t));
  L.state_off = take(hist_len * sizeof(unsigned long long));  // one-sweep look-back entries, [tile][digit]
  L.total = off + 256;
  return L;
}
}  // namespace

size_t radix_sort_scratch_bytes(uint64_t n) { return sort_scratch_layout(n).total; }

SortResult radix_sort_pairs(uint64_t* d_keys_a, uint32_t* d_vals_a, uint64_t* d_keys_b, uint32_t* d_vals_b, uint64_t n,
                            int key_bits, int bit_base, uint8_t* d_scratch, cudaStream_t stream) {
  SortResult cur{d_keys_a, d_vals_a};
  SortResult alt{d_keys_b, d_vals_b};
  if (n == 0) {
    return cur;
  }
  static bool attr_set = false;
  if (!attr_set) {
    MGX_CUDA(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(sizeof(ScatterSmem))));
    MGX_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(sizeof(OsSmem))));
    MGX_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(sizeof(OsSmem))));
    attr_set = true;
  }
  const bool classic = sort_classic();
  const uint32_t n_tiles = static_cast<uint32_t>((n + kSortTile - 1) / kSortTile);
  const uint64_t hist_len = static_cast<uint64_t>(kRadix) * n_tiles;
  // Digits are aligned to the 21-bit code-point fields of the packed key (8 + 8 + 5 bits per field), so that the
  // constant high bits of a field (zero for every BMP code point) form a digit of their own and the pass is skipped.
  PassList pl{};
  for (int f = 0; f * 21 < key_bits; ++f) {
    const int widths[3] = {8, 8, 5};
    int off = 0;
    for (int k = 0; k < 3; ++k) {
      pl.shift[pl.n] = bit_base + f * 21 + off;
      pl.mask[pl.n] = (1u << widths[k]) - 1u;
      ++pl.n;
      off += widths[k];
    }
  }
  const int passes = pl.n;
  const SortScratch L = sort_scratch_layout(n);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_scratch) + 255) & ~static_cast<uintptr_t>(255));
  uint32_t* d_hist = reinterpret_cast<uint32_t*>(base + L.hist_off);
  uint64_t* d_hist_scan = reinterpret_cast<uint64_t*>(base + L.hist_scan_off);
  unsigned long long* d_ghist = reinterpret_cast<unsigned long long*>(base + L.ghist_off);
  uint64_t* d_scan_scratch = reinterpret_cast<uint64_t*>(base + L.scan_off);
  MGX_CUDA(cudaMemsetAsync(d_ghist, 0, kMaxPasses * kRadix * sizeof(unsigned long long), stream));
  PhaseTrace trace(stream);
  int sm_count = 148;
  int dev = 0;
  MGX_CUDA(cudaGetDevice(&dev));
  MGX_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  radix_global_hist_kernel<<<sm_count * 8, 256, 0, stream>>>(cur.keys, n, pl, d_ghist);
  MGX_LAUNCH_CHECK();
  std::vector<unsigned long long> ghist(static_cast<size_t>(kMaxPasses) * kRadix, 0);
  MGX_CUDA(cudaMemcpyAsync(ghist.data(), d_ghist, ghist.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  trace.mark("  sort: global hist");
  unsigned long long* d_dbase = reinterpret_cast<unsigned long long*>(base + L.dbase_off);
  uint32_t* d_counters = reinterpret_cast<uint32_t*>(base + L.counter_off);
  unsigned long long* d_state = reinterpret_cast<unsigned long long*>(base + L.state_off);
  int per_sm = 2;
  if (!classic) {
    digit_scan_kernel<<<passes, kRadix, 0, stream>>>(d_ghist, d_dbase);
    MGX_LAUNCH_CHECK();
    // one memset per sort: the entries carry the pass number, so the passes do not clear them in between
    MGX_CUDA(cudaMemsetAsync(d_counters, 0, kMaxPasses * sizeof(uint32_t), stream));
    MGX_CUDA(cudaMemsetAsync(d_state, 0, hist_len * sizeof(unsigned long long), stream));
    MGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, radix_onesweep_kernel<true>, kSortThreads,
                                                           sizeof(OsSmem)));
    per_sm = std::max(per_sm, 1);
  }
  for (int p = 0; p < passes; ++p) {
    bool trivial = false;
    for (int d = 0; d < kRadix; ++d) {
      trivial = trivial || ghist[static_cast<size_t>(p) * kRadix + d] == n;
    }
    if (trivial) {
      continue;  // every key has the same digit here: the pass would be a stable copy
    }
    const int shift = pl.shift[p];
    const uint32_t mask = pl.mask[p];
    if (!classic) {
      const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(n_tiles, static_cast<uint64_t>(sm_count) * per_sm));
      // expected number of different digit values among 32 keys: sum over d of 1 - (1 - share_d)^32
      double distinct = 0.0;
      for (int d = 0; d < kRadix; ++d) {
        const double share = static_cast<double>(ghist[static_cast<size_t>(p) * kRadix + d]) / static_cast<double>(n);
        distinct += 1.0 - std::pow(1.0 - share, 32.0);
      }
      static const char* force = std::getenv("MGX_OS_RANK");  // "ballot" / "match": A/B measurements
      const bool ballot = force != nullptr ? std::strcmp(force, "ballot") == 0 : distinct > kBallotDistinct;
      auto kernel = ballot ? radix_onesweep_kernel<true> : radix_onesweep_kernel<false>;
      kernel<<<grid, kSortThreads, sizeof(OsSmem), stream>>>(
          cur.keys, cur.vals, alt.keys, alt.vals, n, shift, mask, d_dbase + static_cast<size_t>(p) * kRadix, d_state,
          d_counters + p, n_tiles, static_cast<unsigned long long>(p + 1) << 54);
      MGX_LAUNCH_CHECK();
      std::swap(cur, alt);
      trace.mark("  sort: one-sweep pass");
#ifdef MGX_OS_TIMING
      {
        unsigned long long t[8] = {0};
        MGX_CUDA(cudaMemcpyFromSymbol(t, g_os_timing, sizeof(t)));
        unsigned long long zero[8] = {0};
        MGX_CUDA(cudaMemcpyToSymbol(g_os_timing, zero, sizeof(zero)));
        double tot = 0;
        for (int i = 0; i < 7; ++i) tot += static_cast<double>(t[i]);
        fprintf(stderr, "    [os timing %s] load %.1f%% phase1 %.1f%% scan %.1f%% phase2 %.1f%% lookback %.1f%% lb-others %.1f%% phase3 %.1f%%\n",
                ballot ? "ballot" : "match", 100 * t[0] / tot, 100 * t[1] / tot, 100 * t[2] / tot, 100 * t[3] / tot,
                100 * t[4] / tot, 100 * t[5] / tot, 100 * t[6] / tot);
      }
#endif
      continue;
    }
    radix_hist_kernel<<<n_tiles, kSortThreads, 0, stream>>>(cur.keys, n, shift, mask, d_hist, n_tiles);
    MGX_LAUNCH_CHECK();
    exclusive_scan_u32_u64(d_hist, d_hist_scan, hist_len, d_scan_scratch, stream);
    radix_scatter_kernel<<<n_tiles, kSortThreads, sizeof(ScatterSmem), stream>>>(cur.keys, cur.vals, alt.keys, alt.vals,
                                                                                  n, shift, mask, d_hist_scan, n_tiles);
    MGX_LAUNCH_CHECK();
    std::swap(cur, alt);
    trace.mark("  sort: pass");
  }
  return cur;
}

}  // namespace mgx
