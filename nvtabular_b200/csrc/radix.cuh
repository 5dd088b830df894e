// radix.cuh — hand-written LSD radix-sort passes (sm_100a), the ordering primitive of
// the engine: K3's sort-based group-by for high-cardinality columns (sortagg.cuh), K4's
// (size desc, key asc) vocabulary ordering, and the region partition of the lookup build
// (vocab.cu).  No library code: the reference gets the same orderings from cuDF
// sort_values (nvtabular/ops/categorify.py:1300,1316).
//
// One STABLE pass over n elements E (uint32 keys, or uint64 packed pairs) by a digit of
// up to 11 bits = three kernels, all stream-ordered, no host round trip:
//
//   rx_hist_kernel     one CTA per tile of 8192 elements: digit histogram in shared memory,
//                      written digit-major to tile_hist[digit][tile]
//   rx_rowscan_kernel  one CTA per digit: exclusive scan over the tiles of its row; the
//                      last CTA to finish scans the digit totals (bin bases)
//   rx_scatter_kernel  one CTA per tile: STABLE ranks (per-warp private histograms;
//                      lanes holding the same digit in one 32-element round are ordered
//                      with MATCH.ANY, rounds and warps by construction), the tile is
//                      staged digit-sorted in shared memory and written out so that
//                      consecutive lanes write consecutive addresses of a bin's run.
//
// The element count may live on the device (n_ptr): the sort-based group-by drops the
// null rows in its first pass and only the device knows how many keys are left.
#pragma once

namespace nvtb {

constexpr int kRxThreads = 512;
constexpr int kRxWarps = kRxThreads / 32;
constexpr int kRxTileElems = 8192;          // 16 elements per thread: 64 registers, 2-3 CTAs per SM
constexpr int kRxMaxStableBits = 11;

template <typename E> struct RxTile {
  static constexpr int kElems = kRxTileElems;
  static constexpr int kItems = kElems / kRxThreads;
  static constexpr int kBytes = kElems * (int)sizeof(E);           // u32: 32 KB, u64: 64 KB of staging
};

// digit of an element: bits [shift, shift + bits) of it, optionally complemented
// (descending order)
struct BitsDigit {
  int shift;
  unsigned mask;
  unsigned flip;
  template <typename E> __device__ __forceinline__ unsigned operator()(E e) const {
    return ((unsigned)(e >> shift) & mask) ^ flip;
  }
};

__device__ __forceinline__ int64_t rx_count(const uint32_t* n_ptr, int64_t n_max) {
  if (n_ptr == nullptr) return n_max;
  const int64_t n = (int64_t)*n_ptr;
  return n < n_max ? n : n_max;
}

// exclusive scan of vals[0..P) (shared memory, P % T == 0, blocked: thread t owns
// [t*P/T, (t+1)*P/T)) into out[0..P); returns the total to every thread
template <int T>
__device__ __forceinline__ uint32_t rx_block_excl_scan(const uint32_t* vals, uint32_t* out, int P,
                                                       uint32_t* warp_sums /*[T/32 + 1]*/) {
  const int per = P / T;
  uint32_t local = 0;
  for (int j = 0; j < per; ++j) local += vals[threadIdx.x * per + j];
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += y;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = threadIdx.x < T / 32 ? warp_sums[threadIdx.x] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (threadIdx.x >= o) wi += y;
    }
    if (threadIdx.x < T / 32) warp_sums[threadIdx.x] = wi - w;
    if (threadIdx.x == T / 32 - 1) warp_sums[T / 32] = wi;
  }
  __syncthreads();
  uint32_t run = warp_sums[threadIdx.x >> 5] + incl - local;
  for (int j = 0; j < per; ++j) {
    const uint32_t v = vals[threadIdx.x * per + j];
    out[threadIdx.x * per + j] = run;
    run += v;
  }
  const uint32_t total = warp_sums[T / 32];
  __syncthreads();
  return total;
}

// ---------------------------------------------------------------------------------------
// (1) per-tile digit histograms
// ---------------------------------------------------------------------------------------
template <typename E, typename DigitFn>
__global__ void __launch_bounds__(kRxThreads)
rx_hist_kernel(const E* __restrict__ src, const uint32_t* __restrict__ n_ptr, int64_t n_max,
               DigitFn fn, int nbins, uint32_t* __restrict__ tile_hist, int T) {
  extern __shared__ __align__(16) unsigned char rx_smem[];
  uint32_t* hist = reinterpret_cast<uint32_t*>(rx_smem);
  constexpr int kElems = RxTile<E>::kElems;
  const int64_t n = rx_count(n_ptr, n_max);
  const int tile = blockIdx.x;
  for (int d = threadIdx.x; d < nbins; d += kRxThreads) hist[d] = 0u;
  __syncthreads();
  const int64_t base = (int64_t)tile * kElems;
  if (base < n) {
    const int cnt = (int)((n - base < kElems) ? n - base : kElems);
    constexpr int kVec = 16 / (int)sizeof(E);           // elements per 128-bit load
    const uint4* v4 = reinterpret_cast<const uint4*>(src + base);
    const int nvec = cnt / kVec;
    for (int i = threadIdx.x; i < nvec; i += kRxThreads) {
      const uint4 q = __ldg(v4 + i);
      if constexpr (sizeof(E) == 4) {
        atomicAdd(&hist[fn((E)q.x)], 1u); atomicAdd(&hist[fn((E)q.y)], 1u);
        atomicAdd(&hist[fn((E)q.z)], 1u); atomicAdd(&hist[fn((E)q.w)], 1u);
      } else {
        atomicAdd(&hist[fn((E)(((uint64_t)q.y << 32) | q.x))], 1u);
        atomicAdd(&hist[fn((E)(((uint64_t)q.w << 32) | q.z))], 1u);
      }
    }
    for (int i = nvec * kVec + threadIdx.x; i < cnt; i += kRxThreads) atomicAdd(&hist[fn(src[base + i])], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < nbins; d += kRxThreads) tile_hist[(int64_t)d * T + tile] = hist[d];
}

// ---------------------------------------------------------------------------------------
// (2) row scans + bin bases.  tile_hist[d][.] becomes the exclusive prefix over tiles;
// bin_base[d] the exclusive prefix over digits of the row totals.  `done` is a device
// counter that must be 0 on entry and is left at 0.
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kRxThreads)
rx_rowscan_kernel(uint32_t* __restrict__ tile_hist, int T, int nbins, uint32_t* __restrict__ row_tot,
                         uint32_t* __restrict__ bin_base, unsigned int* done) {
  __shared__ uint32_t ws[kRxWarps + 1];
  __shared__ uint32_t s_carry;
  __shared__ int s_last;
  const int d = blockIdx.x;
  uint32_t* row = tile_hist + (int64_t)d * T;
  if (threadIdx.x == 0) s_carry = 0u;
  __syncthreads();
  // chunks of kRxThreads * 4 entries, carried sequentially (T is a few thousand)
  constexpr int kPer = 4;
  for (int c0 = 0; c0 < T; c0 += kRxThreads * kPer) {
    uint32_t v[kPer];
    uint32_t local = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = c0 + threadIdx.x * kPer + j;
      v[j] = i < T ? row[i] : 0u;
      local += v[j];
    }
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = threadIdx.x < kRxWarps ? ws[threadIdx.x] : 0u;
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
        if (threadIdx.x >= o) wi += y;
      }
      if (threadIdx.x < kRxWarps) ws[threadIdx.x] = wi - w;
      if (threadIdx.x == kRxWarps - 1) ws[kRxWarps] = wi;
    }
    __syncthreads();
    uint32_t run = s_carry + ws[threadIdx.x >> 5] + incl - local;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = c0 + threadIdx.x * kPer + j;
      if (i < T) row[i] = run;
      run += v[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += ws[kRxWarps];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    row_tot[d] = s_carry;
    __threadfence();
    s_last = (atomicAdd(done, 1u) == (unsigned)(nbins - 1)) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {                       // the last row: exclusive scan of the row totals
    __threadfence();
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    for (int c0 = 0; c0 < nbins; c0 += kRxThreads) {
      const int i = c0 + threadIdx.x;
      const uint32_t v = i < nbins ? *reinterpret_cast<volatile uint32_t*>(row_tot + i) : 0u;
      uint32_t incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += y;
      }
      if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
      __syncthreads();
      if (threadIdx.x < 32) {
        uint32_t w = threadIdx.x < kRxWarps ? ws[threadIdx.x] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
          if (threadIdx.x >= o) wi += y;
        }
        if (threadIdx.x < kRxWarps) ws[threadIdx.x] = wi - w;
        if (threadIdx.x == kRxWarps - 1) ws[kRxWarps] = wi;
      }
      __syncthreads();
      if (i < nbins) bin_base[i] = s_carry + ws[threadIdx.x >> 5] + incl - v;
      __syncthreads();
      if (threadIdx.x == 0) s_carry += ws[kRxWarps];
      __syncthreads();
    }
    if (threadIdx.x == 0) *done = 0u;
  }
}

// ---------------------------------------------------------------------------------------
// (3) stable scatter
// ---------------------------------------------------------------------------------------
template <typename E, typename DigitFn>
__global__ void __launch_bounds__(kRxThreads, 2)
rx_scatter_kernel(const E* __restrict__ src, E* __restrict__ dst, const uint32_t* __restrict__ n_ptr,
                  int64_t n_max, DigitFn fn, int nbins, const uint32_t* __restrict__ tile_hist,
                  const uint32_t* __restrict__ bin_base, int T) {
  extern __shared__ __align__(16) unsigned char rx_smem[];
  constexpr int kElems = RxTile<E>::kElems;
  constexpr int kItems = RxTile<E>::kItems;
  const int P = nbins < kRxThreads ? kRxThreads : nbins;      // padded bin count for the block scan
  E* stage = reinterpret_cast<E*>(rx_smem);                                   // [kElems]
  uint32_t* tot = reinterpret_cast<uint32_t*>(rx_smem + RxTile<E>::kBytes);   // [P] digit totals of the tile
  uint32_t* bs = tot + P;                                                      // [P] staged start of every digit
  uint32_t* delta = bs + P;                                                    // [P] global start - staged start
  uint16_t* wh = reinterpret_cast<uint16_t*>(delta + P);                      // [kRxWarps][nbins]
  __shared__ uint32_t ws[kRxWarps + 1];

  const int64_t n = rx_count(n_ptr, n_max);
  const int tile = blockIdx.x;
  const int64_t base = (int64_t)tile * kElems;
  if (base >= n) return;
  const int cnt = (int)((n - base < kElems) ? n - base : kElems);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int i = threadIdx.x; i < kRxWarps * nbins / 2; i += kRxThreads) reinterpret_cast<uint32_t*>(wh)[i] = 0u;
  for (int i = threadIdx.x; i < P; i += kRxThreads) tot[i] = 0u;

  // warp w owns elements [w * 32 * kItems, (w + 1) * 32 * kItems) of the tile; round r is
  // the 32 consecutive elements starting at r * 32: input order == (warp, round, lane)
  E e[kItems];
  const int wbase = warp * 32 * kItems;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int i = wbase + r * 32 + lane;
    e[r] = i < cnt ? src[base + i] : (E)0;
  }
  __syncthreads();
  uint16_t rank[kItems];
  uint16_t* whw = wh + warp * nbins;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int i = wbase + r * 32 + lane;
    const bool valid = i < cnt;
    const unsigned act = __ballot_sync(0xFFFFFFFFu, valid);
    rank[r] = 0;
    if (valid) {
      const unsigned d = fn(e[r]);
      const unsigned m = __match_any_sync(act, d);
      const int leader = __ffs(m) - 1;
      unsigned old = 0;
      if (lane == leader) { old = whw[d]; whw[d] = (uint16_t)(old + __popc(m)); }
      old = __shfl_sync(m, old, leader);
      rank[r] = (uint16_t)(old + __popc(m & lt_mask));
    }
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive scan over the warps (in place) and the tile total
  {
    const int per = P / kRxThreads;
    for (int j = 0; j < per; ++j) {
      const int d = threadIdx.x * per + j;
      if (d < nbins) {
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < kRxWarps; ++w) {
          const unsigned c = wh[w * nbins + d];
          wh[w * nbins + d] = (uint16_t)run;
          run += c;
        }
        tot[d] = run;
      }
    }
  }
  __syncthreads();
  rx_block_excl_scan<kRxThreads>(tot, bs, P, ws);
  for (int d = threadIdx.x; d < nbins; d += kRxThreads)
    delta[d] = tile_hist[(int64_t)d * T + tile] + bin_base[d] - bs[d];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int i = wbase + r * 32 + lane;
    if (i < cnt) {
      const unsigned d = fn(e[r]);
      stage[bs[d] + whw[d] + rank[r]] = e[r];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cnt; i += kRxThreads) {
    const E v = stage[i];
    dst[(int64_t)delta[fn(v)] + i] = v;
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct RxScratch {
  uint32_t* tile_hist;   // [nbins_max * T_max]
  uint32_t* row_tot;     // [nbins_max]
  uint32_t* bin_base;    // [nbins_max]
  unsigned int* done;    // zero
};

template <typename E> inline int rx_tiles(int64_t n_max) {
  return (int)((n_max + RxTile<E>::kElems - 1) / RxTile<E>::kElems);
}
template <typename E> inline size_t rx_scratch_bytes(int64_t n_max, int max_bits) {
  const size_t nb = (size_t)1 << max_bits;
  return sizeof(uint32_t) * (nb * (size_t)rx_tiles<E>(n_max) + 2 * nb) + 256;
}
inline RxScratch rx_scratch_carve(void* p, int T, int max_bits) {
  RxScratch s;
  const size_t nb = (size_t)1 << max_bits;
  s.done = reinterpret_cast<unsigned int*>(p);                       // first 256 bytes (memset to 0 once)
  s.row_tot = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(p) + 256);
  s.bin_base = s.row_tot + nb;
  s.tile_hist = s.bin_base + nb;
  (void)T;
  return s;
}

template <typename E> inline int rx_scatter_smem(int nbins) {
  const int P = nbins < kRxThreads ? kRxThreads : nbins;
  return RxTile<E>::kBytes + 3 * 4 * P + 2 * kRxWarps * nbins;
}

// one stable pass src -> dst by `fn` (digits < 2^bits, bits <= kRxMaxStableBits)
template <typename E, typename DigitFn>
static int rx_pass(const E* src, E* dst, const uint32_t* n_ptr, int64_t n_max, DigitFn fn, int bits,
                   const RxScratch& sc, cudaStream_t st) {
  NVTB_REQUIRE(bits >= 1 && bits <= kRxMaxStableBits, "radix digit width out of range");
  if (n_max <= 0) return NVTB_OK;
  const int nbins = 1 << bits;
  const int T = rx_tiles<E>(n_max);
  const int smem = rx_scatter_smem<E>(nbins);
  static thread_local int attr_bytes = 0;
  if (smem > attr_bytes) {
    NVTB_CUDA_OK(cudaFuncSetAttribute(rx_scatter_kernel<E, DigitFn>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      rx_scatter_smem<E>(1 << kRxMaxStableBits)));
    attr_bytes = rx_scatter_smem<E>(1 << kRxMaxStableBits);
  }
  rx_hist_kernel<E, DigitFn><<<T, kRxThreads, sizeof(uint32_t) * nbins, st>>>(src, n_ptr, n_max, fn, nbins, sc.tile_hist, T);
  NVTB_LAUNCH_OK();
  rx_rowscan_kernel<<<nbins, kRxThreads, 0, st>>>(sc.tile_hist, T, nbins, sc.row_tot, sc.bin_base, sc.done);
  NVTB_LAUNCH_OK();
  rx_scatter_kernel<E, DigitFn><<<T, kRxThreads, smem, st>>>(src, dst, n_ptr, n_max, fn, nbins, sc.tile_hist, sc.bin_base, T);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

// Stable LSD sort of `n` elements by bits [lo_bit, hi_bit) of each element (ascending, or
// descending on that bit field).  a holds the input; the result ends in a (return 0) or
// b (return 1) — *result_in_b.  Scratch from rx_scratch_bytes(n_max, kRxMaxStableBits).
template <typename E>
static int rx_sort_bits(E* a, E* b, const uint32_t* n_ptr, int64_t n_max, int lo_bit, int hi_bit, bool descending,
                        void* scratch, cudaStream_t st, int* result_in_b) {
  *result_in_b = 0;
  if (n_max <= 0 || hi_bit <= lo_bit) return NVTB_OK;
  const int total = hi_bit - lo_bit;
  // <= 10 bits per pass: the scatter kernel then fits two CTAs per SM (108 KB each)
  constexpr int kPrefBits = 10;
  const int passes = (total + kPrefBits - 1) / kPrefBits;
  const int per = (total + passes - 1) / passes;
  const RxScratch sc = rx_scratch_carve(scratch, rx_tiles<E>(n_max), kRxMaxStableBits);
  E* src = a;
  E* dst = b;
  int bit = lo_bit;
  for (int p = 0; p < passes; ++p) {
    const int w = (hi_bit - bit < per) ? hi_bit - bit : per;
    BitsDigit fn{bit, (1u << w) - 1u, descending ? (1u << w) - 1u : 0u};
    int rc = rx_pass<E, BitsDigit>(src, dst, n_ptr, n_max, fn, w, sc, st);
    if (rc) return rc;
    E* t = src; src = dst; dst = t;
    bit += w;
    *result_in_b ^= 1;
  }
  return NVTB_OK;
}

}  // namespace nvtb
