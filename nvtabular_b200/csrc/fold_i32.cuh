// fold_i32.cuh — K3 fast path for int32 keys without payload (every Categorify column of
// the Criteo workload).  Included by hashagg.cu after the global-table primitives.
//
// Why it looks like this (profiles/microbench_atomics_r1.csv, 2^26 keys on one B200):
//   shared-memory RED            52 us   (vs 42 us for just streaming the keys)
//   global RED, L2-resident     370-430 us
//   sector probe + global RED   850-1000 us; > 1400 us once the table leaves L2
// so a row that reaches the global table costs ~20x a row absorbed in shared memory.
// The reference's answer to the same problem is a per-partition cuDF groupby followed by
// a concat+groupby tree (nvtabular/ops/categorify.py:955-1137); here:
//
//   DIRECT mode  (expected distinct keys <= what one SM's shared memory holds):
//     fold_i32_kernel streams the column once; every CTA owns a find-or-claim table in
//     shared memory (4-way buckets, 28 672 slots = 224 KB, filled to <= 20 %) and flushes one
//     (key, count) pair per distinct key into the resident global table at the end.
//   PARTS mode  (more distinct keys): one 512..4096-way hash partition of the keys
//     (part_hist_kernel -> part_scan_kernel -> part_scatter_kernel, staged through
//     shared memory so runs are written, not single words), then the SAME fold kernel
//     runs once per partition: a partition holds ~U/P distinct keys, which fit.
//   Either way a key that finds its shared bucket full goes straight to the global
//   table (and from there, if the table is too small, to the overflow arena): the
//   cardinality estimate only ever costs time, never correctness.
//
// The shared table stores h = fold_hash(key), a BIJECTION of the 32-bit key, instead of
// the key: the top lg(P) bits of h are the partition, the following bits pick the bucket,
// and the flush recovers the key with fold_unhash().
#pragma once

namespace nvtb {

// fold_hash / fold_unhash / kFoldEmpty: common.cuh

constexpr int kFoldThreadsDirect = 1024;   // 1 CTA / SM, 224 KB table
constexpr int kFoldThreadsParts = 512;     // 2 CTAs / SM, 110 KB tables
constexpr int kFoldWays = 4;                       // slots per bucket = one 128-bit shared load
constexpr int kFoldSlotBytes = 10;                 // hash + count + one entry of the live list
constexpr unsigned kFoldBucketsDirect = 5728;      // x 4 slots x 10 B = 224 KB
constexpr unsigned kFoldBucketsParts = 2816;       // 110 KB
// 4-way buckets without displacement overflow for ~0.2 % of the keys at load 0.2 and ~2 % at
// load 0.4; an overflowing key costs every one of its rows the divergent slow path
constexpr double kFoldMaxLoad = 0.2;
constexpr int kPartThreads = 512;
constexpr int kPartGroups = 4;                                   // 8-row groups per thread per tile
constexpr int kPartTile = kPartThreads * 8 * kPartGroups;        // 16384 rows
constexpr int kMaxParts = 4096;
constexpr int kMinParts = 512;

// rows [i, i+8) of an int32 column as one lane's group: values, valid bits, in-range bits
struct Rows8 { int32_t v[8]; unsigned m; unsigned lv; };

__device__ __forceinline__ void load_rows8(const int32_t* __restrict__ keys,
                                           const uint8_t* __restrict__ mask, int64_t i,
                                           int64_t end, Rows8& r, bool aligned = true) {
  if (i + 8 <= end && aligned) {
    ld_rows8<int32_t>(keys + i, r.v);
    r.lv = 0xFFu;
    r.m = valid8(mask, i);
  } else if (i < end) {
    r.lv = (end - i >= 8) ? 0xFFu : ((1u << (unsigned)(end - i)) - 1u);
    r.m = valid8(mask, i) & r.lv;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = (i + k < end) ? keys[i + k] : 0;
  } else {
    r.lv = 0u; r.m = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = 0;
  }
}

// ---------------------------------------------------------------------------------------
// shared-memory find-or-claim table
// ---------------------------------------------------------------------------------------
struct FoldTable {
  uint32_t* hk;    // [4 * nb] hashes, bucket b = hk[4b .. 4b+3]; kFoldEmpty = free
  uint32_t* cnt;   // [4 * nb]
  uint16_t* live;  // [4 * nb] slots claimed since the last flush, in claim order
  unsigned* n_live;
  unsigned nb;
  int lg;          // partition bits already consumed at the top of h

  __device__ __forceinline__ unsigned bucket(uint32_t h) const { return __umulhi(h << lg, nb); }

  template <int T> __device__ __forceinline__ void clear() {
    for (unsigned s = threadIdx.x; s < kFoldWays * nb; s += T) { hk[s] = kFoldEmpty; cnt[s] = 0u; }
  }

  // claim a slot of bucket b for h (or find it there after losing a race); false = full.
  // A fresh claim is appended to the live list: flush and re-clear then touch only the
  // slots in use (a partition of a high-cardinality column fills ~15 % of its table, and a
  // flush that walks every slot serialises ~14 dependent DRAM round trips per thread).
  static __device__ __forceinline__ bool claim(uint32_t* hk, uint32_t* cnt, uint16_t* live,
                                               unsigned* n_live, uint32_t h, unsigned b) {
    bool done = false;
#pragma unroll
    for (int j = 0; j < kFoldWays; ++j) {
      if (!done) {
        const unsigned s = kFoldWays * b + j;
        uint32_t cur = *reinterpret_cast<volatile uint32_t*>(hk + s);
        if (cur == kFoldEmpty) {
          cur = atomicCAS(hk + s, kFoldEmpty, h);
          if (cur == kFoldEmpty) live[atomicAdd(n_live, 1u)] = (uint16_t)s;
        }
        if (cur == kFoldEmpty || cur == h) {
          atomicAdd(cnt + s, 1u);
          done = true;
        }
      }
    }
    return done;
  }

  // candidates of h's bucket (issued early, resolved later: several LDS in flight per lane)
  __device__ __forceinline__ uint4 peek(uint32_t h) const {
    uint4 c;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(hk + kFoldWays * bucket(h));
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "r"(addr));
    return c;
  }
  // fast path only: true = h sat in its bucket and was counted
  __device__ __forceinline__ bool hit(uint32_t h, uint4 c) {
    if (h == kFoldEmpty) return false;            // would "match" a free slot
    const int j = (c.x == h) ? 0 : (c.y == h) ? 1 : (c.z == h) ? 2 : (c.w == h) ? 3 : -1;
    if (j < 0) return false;
    atomicAdd(cnt + kFoldWays * bucket(h) + j, 1u);
    return true;
  }
};

// Everything that is not "h already sits in its shared bucket": claim a shared slot, or -
// bucket full, or h is the one reserved value - update the global table directly.  ONE
// out-of-line copy, called from a cold block after the fast loop, so that the values the
// fast loop keeps in registers are not saved and restored around a call per row.
// Returns the number of keys it added to the GLOBAL table (0 or 1).
static __device__ __noinline__ unsigned fold_slow(const FoldTable& ft, uint32_t h, const Table& t,
                                                  const Arena& arena, Counters* ctr) {
  if (h != kFoldEmpty && FoldTable::claim(ft.hk, ft.cnt, ft.live, ft.n_live, h, ft.bucket(h))) return 0u;
  unsigned n_new = 0;
  const long long key = (long long)(int32_t)fold_unhash(h);
  Probe<true> pr;
  probe_first<true>(t, key, pr);
  upsert_or_spill<true>(t, arena, ctr, key, 1, pr, n_new);
  return n_new;
}

// One work unit = a contiguous row range folded into a freshly cleared shared table and
// flushed.  DIRECT: unit u = rows [u*chunk, (u+1)*chunk) of the column (with its validity
// mask).  PARTS: unit p = partition p = rows [starts[p], ends[p]) of the partition buffer.
template <int T, int MINB, bool PREHASHED>
__global__ void __launch_bounds__(T, MINB)
fold_i32_kernel(const int32_t* __restrict__ keys, const uint8_t* __restrict__ mask, int64_t n,
                const uint32_t* __restrict__ starts, const uint32_t* __restrict__ ends,
                int n_units, int64_t chunk, int lg, unsigned nb, int aligned,
                Table t, Counters* ctr, Arena arena) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FoldTable ft;
  ft.hk = reinterpret_cast<uint32_t*>(smem_raw);
  ft.cnt = ft.hk + kFoldWays * nb;
  ft.live = reinterpret_cast<uint16_t*>(ft.cnt + kFoldWays * nb);
  ft.nb = nb;
  ft.lg = lg;
  __shared__ unsigned int s_null, s_new, s_live;
  ft.n_live = &s_live;
  if (threadIdx.x == 0) { s_null = 0u; s_new = 0u; s_live = 0u; }
  unsigned n_null = 0, n_new = 0;
  ft.clear<T>();      // once: a flush hands the table back empty

  // 8 rows of one lane.  The hash replaces the key in registers (fold_unhash recovers it
  // where needed); bucket candidates are fetched four rows at a time; rows that are not
  // plain hits are only FLAGGED in the fast loop and resolved afterwards.
  auto fold8 = [&](Rows8& r) {
    n_null += __popc(r.lv & ~r.m);
    unsigned pend = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint4 c[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!PREHASHED) r.v[4 * half + k] = (int32_t)fold_hash((uint32_t)r.v[4 * half + k]);
        c[k] = ft.peek((uint32_t)r.v[4 * half + k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (((r.m >> (4 * half + k)) & 1u) && !ft.hit((uint32_t)r.v[4 * half + k], c[k]))
          pend |= 1u << (4 * half + k);
    }
    if (pend) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if ((pend >> k) & 1u)
          n_new += fold_slow(ft, (uint32_t)r.v[k], t, arena, ctr);
    }
  };

  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    int64_t r0, r1;
    if (starts != nullptr) { r0 = (int64_t)starts[unit]; r1 = (int64_t)ends[unit]; }
    else { r0 = (int64_t)unit * chunk; r1 = r0 + chunk < n ? r0 + chunk : n; }
    __syncthreads();
    if (aligned) {
      constexpr int64_t step = (int64_t)T * 8;
      for (int64_t base = r0; base < r1; base += 2 * step) {
        Rows8 a, b;
        load_rows8(keys, mask, base + (int64_t)threadIdx.x * 8, r1, a);
        load_rows8(keys, mask, base + step + (int64_t)threadIdx.x * 8, r1, b);
        fold8(a);
        fold8(b);
      }
    } else {
      for (int64_t i = r0 + threadIdx.x; i < r1; i += T) {
        if (!valid1(mask, i)) { n_null++; continue; }
        const int32_t key = keys[i];
        const uint32_t h = PREHASHED ? (uint32_t)key : fold_hash((uint32_t)key);
        if (!ft.hit(h, ft.peek(h))) n_new += fold_slow(ft, h, t, arena, ctr);
      }
    }
    __syncthreads();
    // flush: one global update per distinct key of the unit, two first probes in flight per
    // thread; every flushed slot is handed back empty
    const unsigned n_live = s_live;
    for (unsigned i0 = 0; i0 < n_live; i0 += T * 2) {
      long long k[2];
      unsigned c[2];
      Probe<true> pr[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const unsigned i = i0 + j * T + threadIdx.x;
        c[j] = 0u;
        k[j] = 0;
        if (i < n_live) {
          const unsigned s = ft.live[i];
          c[j] = ft.cnt[s];
          k[j] = (long long)(int32_t)fold_unhash(ft.hk[s]);
          ft.hk[s] = kFoldEmpty;
          ft.cnt[s] = 0u;
          probe_first<true>(t, k[j], pr[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (c[j]) upsert_or_spill<true>(t, arena, ctr, k[j], (int64_t)c[j], pr[j], n_new);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_live = 0u;
  }
  if (n_null) atomicAdd(&s_null, n_null);
  if (n_new) atomicAdd(&s_new, n_new);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_null) atomicAdd(&ctr->size[0], (unsigned long long)s_null);
    if (s_new) atomicAdd(&ctr->n_unique, (unsigned long long)s_new);
  }
}

// ---------------------------------------------------------------------------------------
// hash partition of the valid keys of a column
// ---------------------------------------------------------------------------------------
// What a key becomes in the partition buffer, and which partition it goes to:
//   PartHashTop  h = fold_hash(key), partition = top lg bits of h   (shared-memory fold)
//   PartKeyLow   u = key ^ 2^31 (unsigned order == signed order), partition = LOW lg bits
//                of u: the first, order-free pass of the LSD radix sort of sortagg.cuh
struct PartHashTop {
  int lg;
  __device__ __forceinline__ uint32_t xform(uint32_t k) const { return fold_hash(k); }
  __device__ __forceinline__ uint32_t bin(uint32_t v) const { return v >> (32 - lg); }
};
struct PartKeyLow {
  int lg;
  __device__ __forceinline__ uint32_t xform(uint32_t k) const { return k ^ 0x80000000u; }
  __device__ __forceinline__ uint32_t bin(uint32_t v) const { return v & ((1u << lg) - 1u); }
};

// (1) partition sizes; the nulls are counted here and dropped by the scatter
template <typename Pol>
__global__ void __launch_bounds__(kPartThreads)
part_hist_kernel(const int32_t* __restrict__ keys, const uint8_t* __restrict__ mask, int64_t n,
                 Pol pol, uint32_t* __restrict__ total, Counters* ctr, int aligned) {
  const int lg = pol.lg;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* cnt = reinterpret_cast<uint32_t*>(smem_raw);
  const int P = 1 << lg;
  for (int d = threadIdx.x; d < P; d += kPartThreads) cnt[d] = 0u;
  __shared__ unsigned int s_null;
  if (threadIdx.x == 0) s_null = 0u;
  __syncthreads();
  unsigned n_null = 0;
  const int64_t n_tiles = (n + kPartTile - 1) / kPartTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    Rows8 r[kPartGroups];
#pragma unroll
    for (int g = 0; g < kPartGroups; ++g)
      load_rows8(keys, mask, tile * kPartTile + ((int64_t)g * kPartThreads + threadIdx.x) * 8, n, r[g], aligned != 0);
#pragma unroll
    for (int g = 0; g < kPartGroups; ++g) {
      n_null += __popc(r[g].lv & ~r[g].m);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if ((r[g].m >> k) & 1u) atomicAdd(&cnt[pol.bin(pol.xform((uint32_t)r[g].v[k]))], 1u);
    }
  }
  if (n_null) atomicAdd(&s_null, n_null);
  __syncthreads();
  for (int d = threadIdx.x; d < P; d += kPartThreads)
    if (cnt[d]) atomicAdd(&total[d], cnt[d]);
  if (threadIdx.x == 0 && s_null && ctr != nullptr) atomicAdd(&ctr->size[0], (unsigned long long)s_null);
}

// exclusive scan of `vals[0..P)` held in shared memory, P % T == 0; every value is first
// rounded up to a multiple of `round` (1 = none).  Result in out[0..P); returns nothing.
template <int T>
__device__ __forceinline__ void block_excl_scan(const uint32_t* vals, uint32_t* out, int P,
                                                uint32_t round, uint32_t* warp_sums /*[T/32]*/) {
  const int per = P / T;
  uint32_t local = 0;
  for (int j = 0; j < per; ++j) {
    const uint32_t v = vals[threadIdx.x * per + j];
    local += (v + round - 1) / round * round;
  }
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
  }
  __syncthreads();
  uint32_t run = warp_sums[threadIdx.x >> 5] + incl - local;
  for (int j = 0; j < per; ++j) {
    const uint32_t v = vals[threadIdx.x * per + j];
    out[threadIdx.x * per + j] = run;
    run += (v + round - 1) / round * round;
  }
  __syncthreads();
}

// (2) partition starts (rounded up to multiples of `round` rows: 8 rows = 32 bytes for the
// fold kernel's 256-bit loads, 1 = dense for the radix sort), write cursors = starts;
// n_total (may be NULL) receives the end of the last partition
static __global__ void __launch_bounds__(kPartThreads)
part_scan_kernel(const uint32_t* __restrict__ total, int lg, uint32_t* __restrict__ starts,
                 uint32_t* __restrict__ cursor, uint32_t round, uint32_t* __restrict__ n_total) {
  __shared__ uint32_t v[kMaxParts];
  __shared__ uint32_t o[kMaxParts];
  __shared__ uint32_t ws[kPartThreads / 32];
  const int P = 1 << lg;
  for (int d = threadIdx.x; d < P; d += kPartThreads) v[d] = total[d];
  __syncthreads();
  block_excl_scan<kPartThreads>(v, o, P, round, ws);
  for (int d = threadIdx.x; d < P; d += kPartThreads) { starts[d] = o[d]; cursor[d] = o[d]; }
  if (n_total != nullptr && threadIdx.x == 0) *n_total = o[P - 1] + (v[P - 1] + round - 1) / round * round;
}

// (3) scatter (the buffer receives h = fold_hash(key), which the fold kernel consumes as
// is).  Per tile of 16 384 rows: count per partition (shared RED), reserve the
// tile's run in every partition with ONE global atomic per non-empty (tile, partition),
// bin the keys in shared memory, then copy the staged tile out so that consecutive lanes
// write consecutive words of a run.  Order inside a partition is irrelevant (counting).
template <typename Pol>
__global__ void __launch_bounds__(kPartThreads, 2)
part_scatter_kernel(const int32_t* __restrict__ keys, const uint8_t* __restrict__ mask, int64_t n,
                    Pol pol, uint32_t* __restrict__ cursor, int32_t* __restrict__ out, int aligned) {
  const int lg = pol.lg;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* stage = reinterpret_cast<int32_t*>(smem_raw);                 // [kPartTile]
  uint32_t* cnt = reinterpret_cast<uint32_t*>(stage + kPartTile);        // [P] counts, then running cursors
  uint32_t* delta = cnt + (1 << lg);                                     // [P] global start - staged start
  __shared__ uint32_t ws[kPartThreads / 32];
  __shared__ uint32_t s_total;
  const int P = 1 << lg;
  const int64_t n_tiles = (n + kPartTile - 1) / kPartTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int d = threadIdx.x; d < P; d += kPartThreads) cnt[d] = 0u;
    __syncthreads();
    Rows8 r[kPartGroups];
#pragma unroll
    for (int g = 0; g < kPartGroups; ++g)
      load_rows8(keys, mask, tile * kPartTile + ((int64_t)g * kPartThreads + threadIdx.x) * 8, n, r[g], aligned != 0);
#pragma unroll
    for (int g = 0; g < kPartGroups; ++g)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        r[g].v[k] = (int32_t)pol.xform((uint32_t)r[g].v[k]);      // the buffer holds hashes / biased keys
        if ((r[g].m >> k) & 1u) atomicAdd(&cnt[pol.bin((uint32_t)r[g].v[k])], 1u);
      }
    __syncthreads();
    // staged offsets (exclusive scan of the counts, in place via `delta` as scratch)
    block_excl_scan<kPartThreads>(cnt, delta, P, 1u, ws);
    for (int d = threadIdx.x; d < P; d += kPartThreads) {
      const uint32_t c = cnt[d], off = delta[d];
      uint32_t g0 = 0;
      if (c) g0 = atomicAdd(&cursor[d], c);
      delta[d] = g0 - off;            // modulo 2^32: global index = delta + staged index
      cnt[d] = off;                   // running staged cursor
      if (d == P - 1) s_total = off + c;
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < kPartGroups; ++g)
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if ((r[g].m >> k) & 1u) {
          const uint32_t p = atomicAdd(&cnt[pol.bin((uint32_t)r[g].v[k])], 1u);
          stage[p] = r[g].v[k];
        }
    __syncthreads();
    const uint32_t total = s_total;
    for (uint32_t j = threadIdx.x; j < total; j += kPartThreads) {
      const int32_t hv = stage[j];
      out[delta[pol.bin((uint32_t)hv)] + j] = hv;
    }
    __syncthreads();
  }
}

}  // namespace nvtb
