// hashagg.cu — K3: device hash aggregation  groupby(key, dropna=False).agg(...)
//
// Replaces the reference's per-partition cuDF groupby + concat/groupby tree
// (nvtabular/ops/categorify.py:955-1137, graph built at :1344-1540) with ONE
// resident open-addressing table per column group that every batch is folded
// into.  Design (B200-first, not a translation of cuDF's groupby, which sizes a
// fresh 2N-slot table per partition):
//
//   * two slot layouts (see `struct Table`): wide {int64 key, int64 size} (+ optional
//     per-slot payload {sum, sumsq, min, max} per cont column) and narrow 8-byte packed
//     slots in 4-way sector buckets for int32 keys without payload.
//   * int32 keys without payload (every Categorify column of the Criteo workload) take
//     the path of fold_i32.cuh: rows are absorbed in SHARED memory (find-or-claim tables
//     of 110-224 KB per CTA), after a one-pass hash partition of the column when the
//     expected number of distinct keys exceeds what one table holds; the global table
//     only sees one (key, count) pair per distinct key per CTA / partition.
//   * other keys: ONE launch per (column, batch); insert_keys_kernel streams the key
//     column with 256-bit loads and first folds rows into a per-CTA shared-memory table
//     (64 KB) - a CTA whose first tile shows < 25 % reuse bypasses it; insert_agg_kernel
//     (payload) updates the global table row by row.
//   * the table is sized from a CARDINALITY ESTIMATE (exact distinct count of
//     the first 2^20 rows, inverted through U = K(1 - exp(-s/K)), or the
//     caller's hint), not from the worst case.  Correctness never depends on
//     the estimate: a probe gives up after 128 buckets and the refused
//     (key, count) pair goes to an overflow ARENA sized for the launch.  The
//     next call on the handle ("settle") reads the counters back, grows the
//     table if the arena is non-empty and merges the arena into it.
//   * the null key (dropna=False) and the one key equal to the EMPTY sentinel
//     (INT64_MIN) live in two "special" groups outside the table.
//
// Throughput bound: the shared-memory pipe and L2 atomics / random 32-B sector traffic,
// not the HBM stream; see DESIGN.md "K3 design and its real roofline".
#include <algorithm>
#include <mutex>
#include <new>

#include "common.cuh"

namespace nvtb {

constexpr int kSmemProbes = 4;
constexpr int kMaxProbes = 128;    // global probe limit before a pair is refused
constexpr int64_t kSampleRows = (int64_t)1 << 20;
constexpr int64_t kMinCapacity = 1 << 16;
constexpr int kInsertCtasPerSm = 3;

// min/max are kept as order-preserving int64 images of the double so that the
// native 64-bit atomicMin/atomicMax can be used.
__host__ __device__ __forceinline__ int64_t enc_ordered(double x) {
#ifdef __CUDA_ARCH__
  int64_t b = __double_as_longlong(x);
#else
  int64_t b; memcpy(&b, &x, 8);
#endif
  return b >= 0 ? b : (b ^ 0x7FFFFFFFFFFFFFFFll);
}
__host__ __device__ __forceinline__ double dec_ordered(int64_t e) {
  int64_t b = e >= 0 ? e : (e ^ 0x7FFFFFFFFFFFFFFFll);
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double x; memcpy(&x, &b, 8); return x;
#endif
}
// INT64_MAX / INT64_MIN decode to NaN: "no value seen" == pandas NaN min/max.
constexpr int64_t kMinInit = INT64_MAX;
constexpr int64_t kMaxInit = INT64_MIN;

// Two slot layouts:
//   wide   (16 B)  {int64 key, int64 size}; empty key = INT64_MIN
//   narrow ( 8 B)  (uint32 size << 32) | uint32 key; empty = 0 (a live slot has
//                  size >= 1).  Used for int32 keys without payload while the
//                  handle has seen < 2^32 rows: half the footprint (tables of a
//                  few million keys stay L2-resident), a new key costs ONE
//                  64-bit CAS that also deposits the count, a known key one
//                  32-bit RED.
struct Table {
  int64_t* slots;   // wide: [2*capacity]; narrow: [capacity] packed words
  double* vals;     // [capacity * 4 * n_agg] or nullptr (wide only)
  int64_t capacity; // power of two
  int n_agg;
  int narrow;
};

struct Counters {
  unsigned long long n_unique;   // distinct keys in the table
  unsigned long long size[2];    // special groups: [0] null key, [1] INT64_MIN key
  unsigned long long ovf_count;  // pairs refused into the arena by the pending launch
  unsigned long long max_count;  // sorted accumulator (sortagg.cuh): largest group size seen
};

struct Arena {
  int64_t* keys;
  int64_t* sizes;
  double* vals;      // [cap * 4 * n_agg] or nullptr
  int64_t cap;
  int slot;          // index in the shared arena pool, -1 = private allocation
};

}  // namespace nvtb

struct nvtb_hashagg {
  nvtb::Table t;
  nvtb::Counters* ctr;       // device
  double* special_vals;      // device [2][4*n_agg]
  nvtb::Counters* mailbox;   // pinned host
  cudaEvent_t ev;
  bool pending;              // a launch whose counters have not been read back
  nvtb::Arena arena;         // arena of the pending launch
  cudaStream_t pending_stream;
  int64_t u_known;           // distinct keys at the last settle
  int64_t rows_total;        // rows folded in so far
  double k_est;              // cardinality estimate (0 = none yet)
  double predicted;          // distinct keys expected after the batch being prepared
  int64_t hint;
  int n_agg;
  bool mailbox_valid;        // mailbox == device counters (no launch / host edit since the readback)
  // sorted accumulator (sortagg.cuh), mode == 1: acc[acc_cur] holds ctr->n_unique packed pairs
  int mode;                  // 0 = hash table, 1 = sorted runs
  uint64_t* acc[2];
  int64_t acc_cap[2];
  int acc_cur;
  uint32_t* d_n;             // device uint32[4]: [1] valid keys of the batch, [2] distinct keys of the batch
  // staging of a sorted accumulator: batches are only COPIED (keys + validity bytes) until
  // NVTB_STAGE_ROWS rows are waiting or somebody reads the handle; one sort + run-length
  // encode + merge then takes all of them (a merge per batch re-reads and re-writes the
  // whole accumulator: 4.2 -> 5.7 ms per 6.25e7-row batch at 1.4e8 accumulated keys)
  int32_t* stage_keys;
  uint8_t* stage_mask;
  int64_t stage_cap;         // rows
  int64_t stage_rows;        // rows waiting (a multiple of 8 except after the last batch)
  int64_t stage_hint;        // rows the previous fits staged: the buffer grows to hold one whole fit
  cudaEvent_t stage_ev;
  cudaStream_t stage_last;
};

namespace nvtb {

__device__ __forceinline__ void vals_combine(double* __restrict__ dst,
                                             double sum, double sumsq,
                                             double mn, double mx) {
  atomicAdd(dst + 0, sum);
  atomicAdd(dst + 1, sumsq);
  if (mn == mn) atomicMin(reinterpret_cast<long long*>(dst + 2), (long long)enc_ordered(mn));
  if (mx == mx) atomicMax(reinterpret_cast<long long*>(dst + 3), (long long)enc_ordered(mx));
}

// append one refused pair (rare path: a plain atomic, safe under any divergence)
__device__ __forceinline__ int64_t arena_claim(Counters* ctr) {
  return (int64_t)atomicAdd(&ctr->ovf_count, 1ull);
}

// A probe = the table words fetched for one key.
//   narrow: one 32-byte SECTOR = a 4-way bucket of packed slots, fetched with a single
//           256-bit load (LDG.E.256).  A key displaced by collisions is still found on
//           the inlined fast path unless its bucket holds > 4 keys; with one slot per
//           probe ~load% of ALL rows took the divergent out-of-line path.
//   wide:   one 16-byte slot (key word only), linear probing.
template <bool NARROW> struct Probe;
template <> struct Probe<true>  { int64_t b; unsigned long long w[4]; };
template <> struct Probe<false> { int64_t b; unsigned long long w[1]; };

template <bool NARROW>
__device__ __forceinline__ void probe_load(const Table& t, int64_t b, Probe<NARROW>& p) {
  p.b = b;
  if constexpr (NARROW) {
    const unsigned long long* a = reinterpret_cast<const unsigned long long*>(t.slots) + 4 * b;
    asm volatile("ld.global.cg.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=l"(p.w[0]), "=l"(p.w[1]), "=l"(p.w[2]), "=l"(p.w[3]) : "l"(a), "l"(l2_evict_last()));
  } else {
    p.w[0] = __ldcg(reinterpret_cast<const unsigned long long*>(t.slots) + 2 * b);
  }
}

template <bool NARROW>
__device__ __forceinline__ int64_t probe_home(const Table& t, int64_t key) {
  if constexpr (NARROW)   // int32 keys: 32-bit mixer, low bits pick the bucket
    return (int64_t)((uint64_t)table_mix32((uint32_t)(int32_t)key) & (uint64_t)((t.capacity >> 2) - 1));
  return (int64_t)(table_mix64((uint64_t)key) & (uint64_t)(t.capacity - 1));
}

template <bool NARROW>
__device__ __forceinline__ void probe_first(const Table& t, int64_t key, Probe<NARROW>& p) {
  probe_load<NARROW>(t, probe_home<NARROW>(t, key), p);
}

// Everything that is not "the key sits in the prefetched bucket": claim an empty slot
// (64-bit CAS; on the narrow layout the CAS also deposits the count), walk to the next
// bucket, give up after kMaxProbes buckets (-> -1, the pair is refused).  ONE out-of-line
// copy keeps the 16-row-unrolled kernel small enough for the instruction cache.
// There is no load-factor guard: a table that is too small simply fills up, probes start
// failing, refused pairs go to the arena and settle() regrows the table.
template <bool NARROW>
__device__ __noinline__ int64_t upsert_slow(const Table& t, int64_t key, int64_t add,
                                            Probe<NARROW> p, unsigned& n_new) {
  unsigned long long* base = reinterpret_cast<unsigned long long*>(t.slots);
  const int64_t mask = NARROW ? (t.capacity >> 2) - 1 : t.capacity - 1;
  int64_t result = -1;
  bool done = false;
#pragma unroll 1
  for (int probe = 0; probe < kMaxProbes && !done; ++probe) {
    if constexpr (NARROW) {
      const unsigned long long want =
          ((unsigned long long)(unsigned)add << 32) | (unsigned long long)(unsigned)key;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!done) {
          unsigned long long cur = p.w[j];
          if (cur == 0ull) {
            cur = atomicCAS(base + 4 * p.b + j, 0ull, want);
            if (cur == 0ull) { ++n_new; result = 4 * p.b + j; done = true; }   // count deposited
          }
          if (!done && (unsigned)cur == (unsigned)key) {        // cur != 0 here
            atomicAdd(reinterpret_cast<unsigned*>(base + 4 * p.b + j) + 1, (unsigned)add);
            result = 4 * p.b + j; done = true;
          }
        }
      }
    } else {
      bool match = ((long long)p.w[0] == key);
      if ((long long)p.w[0] == kEmptyKey) {
        const unsigned long long prev = atomicCAS(base + 2 * p.b, (unsigned long long)kEmptyKey,
                                                  (unsigned long long)key);
        if ((long long)prev == kEmptyKey) { ++n_new; match = true; }
        else match = ((long long)prev == key);
      }
      if (match) {
        atomicAdd(base + 2 * p.b + 1, (unsigned long long)add);
        result = p.b; done = true;
      }
    }
    if (!done) probe_load<NARROW>(t, (p.b + 1) & mask, p);
  }
  return result;
}

// Fold (key, add) into the table starting from a PREFETCHED first probe: callers issue
// the first-probe loads of several keys back to back and only then resolve them, so a
// thread keeps several table sectors in flight.  Returns the slot, or -1 (refused).
template <bool NARROW>
__device__ __forceinline__ int64_t upsert_add(const Table& t, int64_t key, int64_t add,
                                              const Probe<NARROW>& p, unsigned& n_new) {
  unsigned long long* base = reinterpret_cast<unsigned long long*>(t.slots);
  if constexpr (NARROW) {
    int hit = -1;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (p.w[j] != 0ull && (unsigned)p.w[j] == (unsigned)key) hit = j;
    if (hit >= 0) {
      atomicAdd(reinterpret_cast<unsigned*>(base + 4 * p.b + hit) + 1, (unsigned)add);
      return 4 * p.b + hit;
    }
  } else {
    if ((long long)p.w[0] == key) {
      atomicAdd(base + 2 * p.b + 1, (unsigned long long)add);
      return p.b;
    }
  }
  return upsert_slow<NARROW>(t, key, add, p, n_new);
}

// one-key convenience form (cold paths: merge, scalar tails)
template <bool NARROW>
__device__ __forceinline__ int64_t upsert_one(const Table& t, int64_t key, int64_t add, unsigned& n_new) {
  Probe<NARROW> p;
  probe_first<NARROW>(t, key, p);
  return upsert_add<NARROW>(t, key, add, p, n_new);
}

template <bool NARROW>
__device__ __forceinline__ void upsert_or_spill(const Table& t, const Arena& a, Counters* ctr,
                                                int64_t key, int64_t add, const Probe<NARROW>& p,
                                                unsigned& n_new) {
  if (upsert_add<NARROW>(t, key, add, p, n_new) < 0) {
    const int64_t o = arena_claim(ctr);
    if (o < a.cap) { a.keys[o] = key; a.sizes[o] = add; }
  }
}

__global__ void table_init_kernel(Table t) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
       s < t.capacity; s += stride) {
    if (t.narrow) { t.slots[s] = 0; continue; }
    t.slots[2 * s] = kEmptyKey;
    t.slots[2 * s + 1] = 0;
    for (int j = 0; j < t.n_agg; ++j) {
      double* v = t.vals + (s * t.n_agg + j) * 4;
      v[0] = 0.0; v[1] = 0.0;
      reinterpret_cast<int64_t*>(v)[2] = kMinInit;
      reinterpret_cast<int64_t*>(v)[3] = kMaxInit;
    }
  }
}

__global__ void arm_launch_kernel(Counters* ctr) {
  ctr->ovf_count = 0ull;
}

__global__ void special_init_kernel(Counters* ctr, double* special_vals, int n_agg) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ctr->n_unique = 0; ctr->size[0] = 0; ctr->size[1] = 0; ctr->ovf_count = 0; ctr->max_count = 0;
    for (int g = 0; g < 2; ++g)
      for (int j = 0; j < n_agg; ++j) {
        double* v = special_vals + (g * n_agg + j) * 4;
        v[0] = 0.0; v[1] = 0.0;
        reinterpret_cast<int64_t*>(v)[2] = kMinInit;
        reinterpret_cast<int64_t*>(v)[3] = kMaxInit;
      }
  }
}

}  // namespace nvtb
#include "fold_i32.cuh"
namespace nvtb {
constexpr int kExportPerThread = 4;
}
#include "sortagg.cuh"
#include "bucketagg.cuh"
namespace nvtb {

// ---------------------------------------------------------------------------
// insert, keys only (Categorify)
// ---------------------------------------------------------------------------
// Per-CTA shared-memory pre-aggregation table.
//   int32 keys: 8192 packed slots (count << 32 | key; 0 = empty), 64 KB
//   int64 keys: 4096 slots of {int64 key, uint32 count}, 48 KB
template <typename KeyT> struct SmemAgg;

template <> struct SmemAgg<int32_t> {
  // 4096 two-way buckets: one 128-bit shared load fetches both candidate slots of a key,
  // so a key displaced by a collision is still found on the inlined fast path (with
  // one-slot-per-probe linear probing ~load% of ALL rows took the divergent slow path).
  static constexpr int kBuckets = 4096;
  static constexpr int kSlots = 2 * kBuckets;
  static constexpr int kBytes = kSlots * 8;
  unsigned long long* w;
  __device__ __forceinline__ explicit SmemAgg(unsigned char* raw) : w(reinterpret_cast<unsigned long long*>(raw)) {}
  __device__ __forceinline__ void clear() {
    for (int s = threadIdx.x; s < kSlots; s += kThreads) w[s] = 0ull;
  }
  // claim one of the two slots of bucket b for `key` (or find it there); false = bucket full
  static __device__ __noinline__ bool fold_slow(unsigned long long* w, unsigned key, unsigned b) {
    bool done = false;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (!done) {
        unsigned long long* p = w + 2 * b + j;
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(p);
        if (cur == 0ull) {
          cur = atomicCAS(p, 0ull, (1ull << 32) | (unsigned long long)key);
          if (cur == 0ull) done = true;                          // claimed with count 1
        }
        if (!done && cur != 0ull && (unsigned)cur == key) {
          atomicAdd(reinterpret_cast<unsigned*>(p) + 1, 1u);
          done = true;
        }
      }
    }
    return done;
  }
  // fold one key; true = absorbed.  `hits` counts keys that were already present.
  __device__ __forceinline__ bool fold(long long k, unsigned& hits) {
    const unsigned key = (unsigned)(int)k;
    const unsigned b = (table_mix32(key) >> 20) & (kBuckets - 1);   // top bits; the global bucket uses the low bits
    unsigned long long x, y;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(w + 2 * b);
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(addr));
    const bool mx = (x != 0ull) && ((unsigned)x == key);
    const bool my = (y != 0ull) && ((unsigned)y == key);
    if (mx || my) {
      atomicAdd(reinterpret_cast<unsigned*>(w + 2 * b + (mx ? 0 : 1)) + 1, 1u);
      hits++;
      return true;
    }
    return fold_slow(w, key, b);
  }
  __device__ __forceinline__ bool get(int s, long long* key, unsigned* cnt) const {
    const unsigned long long cur = w[s];
    *key = (long long)(int)(unsigned)cur;
    *cnt = (unsigned)(cur >> 32);
    return cur != 0ull;
  }
};

template <> struct SmemAgg<int64_t> {
  static constexpr int kSlots = 4096;
  static constexpr int kBytes = kSlots * 12;
  long long* keys;
  unsigned* cnt;
  __device__ __forceinline__ explicit SmemAgg(unsigned char* raw)
      : keys(reinterpret_cast<long long*>(raw)), cnt(reinterpret_cast<unsigned*>(raw + sizeof(long long) * kSlots)) {}
  __device__ __forceinline__ void clear() {
    for (int s = threadIdx.x; s < kSlots; s += kThreads) { keys[s] = kEmptyKey; cnt[s] = 0u; }
  }
  static __device__ __noinline__ bool fold_slow(long long* keys, unsigned* cnt, long long k, unsigned s) {
    bool done = false;
#pragma unroll 1
    for (int p = 0; p < kSmemProbes && !done; ++p) {
      long long cur = *reinterpret_cast<volatile long long*>(&keys[s]);
      if (cur == kEmptyKey)
        cur = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&keys[s]),
                                   (unsigned long long)kEmptyKey, (unsigned long long)k);
      if (cur == kEmptyKey || cur == k) {
        atomicAdd(&cnt[s], 1u);
        done = true;
      }
      s = (s + 1) & (kSlots - 1);
    }
    return done;
  }
  __device__ __forceinline__ bool fold(long long k, unsigned& hits) {
    const unsigned s = (unsigned)(table_mix64((uint64_t)k) >> 40) & (kSlots - 1);
    if (*reinterpret_cast<volatile long long*>(&keys[s]) == k) {
      atomicAdd(&cnt[s], 1u);
      hits++;
      return true;
    }
    return fold_slow(keys, cnt, k, s);
  }
  __device__ __forceinline__ bool get(int s, long long* key, unsigned* c) const {
    *key = keys[s];
    *c = cnt[s];
    return keys[s] != kEmptyKey;
  }
};

template <typename KeyT, bool NARROW>
__global__ void __launch_bounds__(kThreads, 3)
insert_keys_kernel(const KeyT* __restrict__ keys,
                   const uint8_t* __restrict__ mask, int64_t n, Table t,
                   Counters* ctr, Arena arena) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemAgg<KeyT> sm(smem_raw);
  __shared__ unsigned long long s_null, s_min;
  __shared__ unsigned int s_hits, s_rows, s_new;
  __shared__ int s_bypass;
  sm.clear();
  if (threadIdx.x == 0) { s_null = 0ull; s_min = 0ull; s_hits = 0u; s_rows = 0u; s_new = 0u; s_bypass = 0; }
  __syncthreads();

  unsigned int n_null = 0, n_min = 0, n_new = 0;
  bool bypass = false;
  const bool aligned = is_aligned32(keys);

  // 8 rows of one lane: (1) shared-memory fold, (2) first-probe loads of every key
  // that still has to reach the global table, issued back to back, (3) resolve
  auto fold8 = [&](const KeyT (&v)[kRows], unsigned m, unsigned& hits, unsigned& rows) {
    unsigned pend = 0;
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const bool valid = (m >> k) & 1u;
      const long long key = (long long)v[k];
      const bool is_min = valid && sizeof(KeyT) == 8 && key == kEmptyKey;
      n_null += valid ? 0u : 1u;
      n_min += is_min ? 1u : 0u;
      if (valid && !is_min) {
        rows++;
        if (bypass || !sm.fold(key, hits)) pend |= 1u << k;
      }
    }
    // global phase in two halves of 4 keys: 4 sector loads in flight per thread
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const unsigned hp = (pend >> (4 * half)) & 0xFu;
      if (hp != 0) {
        Probe<NARROW> pr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((hp >> k) & 1u) probe_first<NARROW>(t, (long long)v[4 * half + k], pr[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((hp >> k) & 1u)
            upsert_or_spill<NARROW>(t, arena, ctr, (long long)v[4 * half + k], 1, pr[k], n_new);
      }
    }
  };

  const int64_t n_tiles = (n + kTile - 1) / kTile;
  bool first = true;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kTile;
    unsigned hits = 0, rows = 0;
    if (aligned && base + kTile <= n) {
      KeyT v[kGroups][kRows];
      unsigned m[kGroups];
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        const int64_t i = base + (int64_t)g * (kThreads * kRows) + (int64_t)threadIdx.x * kRows;
        ld_rows8<KeyT>(keys + i, v[g]);
        m[g] = valid8(mask, i);
      }
#pragma unroll
      for (int g = 0; g < kGroups; ++g) fold8(v[g], m[g], hits, rows);
    } else {
      const int64_t end = (base + kTile < n) ? base + kTile : n;
      for (int64_t i = base + threadIdx.x; i < end; i += kThreads) {
        if (!valid1(mask, i)) { n_null++; continue; }
        const long long key = (long long)keys[i];
        if (sizeof(KeyT) == 8 && key == kEmptyKey) { n_min++; continue; }
        rows++;
        if (!bypass && sm.fold(key, hits)) continue;
        Probe<NARROW> pr;
        probe_first<NARROW>(t, key, pr);
        upsert_or_spill<NARROW>(t, arena, ctr, key, 1, pr, n_new);
      }
    }
    if (first) {
      // after its first tile the CTA decides whether shared-memory folding pays:
      // < 25 % of the rows met an already-present key  =>  go straight to global
      first = false;
      if (hits) atomicAdd(&s_hits, hits);
      if (rows) atomicAdd(&s_rows, rows);
      __syncthreads();
      if (threadIdx.x == 0) s_bypass = (s_hits * 4u < s_rows) ? 1 : 0;
      __syncthreads();
      bypass = (s_bypass != 0);
    }
  }

  if (n_null) atomicAdd(&s_null, (unsigned long long)n_null);
  if (n_min) atomicAdd(&s_min, (unsigned long long)n_min);
  __syncthreads();
  // flush the CTA-local aggregates: one global update per distinct key per CTA
  for (int s0 = 0; s0 < SmemAgg<KeyT>::kSlots; s0 += kThreads * 4) {
    long long k[4];
    unsigned c[4];
    bool live[4];
    Probe<NARROW> pr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      live[j] = sm.get(s0 + j * kThreads + threadIdx.x, &k[j], &c[j]);
      if (live[j]) probe_first<NARROW>(t, k[j], pr[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (live[j]) upsert_or_spill<NARROW>(t, arena, ctr, k[j], (int64_t)c[j], pr[j], n_new);
  }
  if (n_new) atomicAdd(&s_new, n_new);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_null) atomicAdd(&ctr->size[0], s_null);
    if (s_min) atomicAdd(&ctr->size[1], s_min);
    if (s_new) atomicAdd(&ctr->n_unique, (unsigned long long)s_new);
  }
}

// ---------------------------------------------------------------------------
// insert with continuous payload (JoinGroupby / TargetEncoding)
// ---------------------------------------------------------------------------
constexpr int kMaxAgg = 8;
struct AggCols {
  const void* data[kMaxAgg];
  const uint8_t* mask[kMaxAgg];
  int32_t dtype[kMaxAgg];
};

__device__ __forceinline__ bool load_agg(const AggCols& a, int j, int64_t i,
                                         double* out) {
  if (!valid1(a.mask[j], i)) return false;
  double v;
  switch (a.dtype[j]) {
    case NVTB_I32: v = (double)((const int32_t*)a.data[j])[i]; break;
    case NVTB_I64: v = (double)((const int64_t*)a.data[j])[i]; break;
    case NVTB_F32: v = (double)((const float*)a.data[j])[i]; break;
    case NVTB_U8:  v = (double)((const uint8_t*)a.data[j])[i]; break;
    default:       v = ((const double*)a.data[j])[i]; break;
  }
  if (v != v) return false;  // NaN == null
  *out = v;
  return true;
}

template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
insert_agg_kernel(const KeyT* __restrict__ keys,
                  const uint8_t* __restrict__ mask, AggCols agg, int64_t n,
                  Table t, Counters* ctr, double* special_vals, Arena arena) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned n_new = 0;
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const bool valid = valid1(mask, i);
    const long long k = valid ? (long long)keys[i] : 0;
    double* vdst = nullptr;
    if (!valid) {
      atomicAdd(&ctr->size[0], 1ull);
      vdst = special_vals;
    } else if (sizeof(KeyT) == 8 && k == kEmptyKey) {
      atomicAdd(&ctr->size[1], 1ull);
      vdst = special_vals + (int64_t)t.n_agg * 4;
    } else {
      const int64_t slot = upsert_one<false>(t, k, 1, n_new);
      if (slot >= 0) {
        vdst = t.vals + slot * t.n_agg * 4;
      } else {
        const int64_t o = arena_claim(ctr);
        if (o < arena.cap) {
          arena.keys[o] = k;
          arena.sizes[o] = 1;
          for (int j = 0; j < t.n_agg; ++j) {
            double v;
            double* w = arena.vals + (o * t.n_agg + j) * 4;
            if (load_agg(agg, j, i, &v)) { w[0] = v; w[1] = v * v; w[2] = v; w[3] = v; }
            else { w[0] = 0.0; w[1] = 0.0; w[2] = nan; w[3] = nan; }
          }
        }
        continue;
      }
    }
    for (int j = 0; j < t.n_agg; ++j) {
      double v;
      if (load_agg(agg, j, i, &v)) vals_combine(vdst + j * 4, v, v * v, v, v);
    }
  }
  if (n_new) atomicAdd(&ctr->n_unique, (unsigned long long)n_new);
}

// merge pre-aggregated rows (other GPUs' partials, or a drained arena)
template <bool NARROW>
__global__ void __launch_bounds__(kThreads)
merge_kernel(const int64_t* __restrict__ keys, const int64_t* __restrict__ sizes,
             const double* __restrict__ vals, int64_t n, Table t, Counters* ctr,
             double* special_vals, Arena arena) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned n_new = 0;
  // four rows per thread per step: their first probes are issued back to back
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    long long k[4];
    bool live[4];
    Probe<NARROW> pr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = i0 + j * stride;
      live[j] = i < n;
      k[j] = live[j] ? keys[i] : 0;
      if (live[j] && (NARROW || k[j] != kEmptyKey)) probe_first<NARROW>(t, k[j], pr[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!live[j]) continue;
      const int64_t i = i0 + j * stride;
      double* vdst;
      if (!NARROW && k[j] == kEmptyKey) {
        atomicAdd(&ctr->size[1], (unsigned long long)sizes[i]);
        vdst = special_vals + (int64_t)t.n_agg * 4;
      } else {
        const int64_t slot = upsert_add<NARROW>(t, k[j], sizes[i], pr[j], n_new);
        if (slot < 0) {
          const int64_t o = arena_claim(ctr);
          if (o < arena.cap) {
            arena.keys[o] = k[j];
            arena.sizes[o] = sizes[i];
            if (vals != nullptr)
              for (int q = 0; q < t.n_agg * 4; ++q)
                arena.vals[o * t.n_agg * 4 + q] = vals[i * t.n_agg * 4 + q];
          }
          continue;
        }
        vdst = NARROW ? nullptr : t.vals + slot * t.n_agg * 4;
      }
      if (vals != nullptr && vdst != nullptr)
        for (int q = 0; q < t.n_agg; ++q) {
          const double* v = vals + (i * t.n_agg + q) * 4;
          vals_combine(vdst + q * 4, v[0], v[1], v[2], v[3]);
        }
    }
  }
  if (n_new) atomicAdd(&ctr->n_unique, (unsigned long long)n_new);
}

// rehash an old table into a new one (larger, and/or narrow -> wide).  Keys are
// distinct: plain stores after the claim; n_unique is unchanged.  The new table is
// at most half full, so the probe loop terminates.
__global__ void __launch_bounds__(kThreads)
rehash_kernel(Table old_t, Table new_t) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t mask = new_t.capacity - 1;
  unsigned long long* nb = reinterpret_cast<unsigned long long*>(new_t.slots);
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
       s < old_t.capacity; s += stride) {
    long long k, sz;
    if (old_t.narrow) {
      const unsigned long long w = (unsigned long long)old_t.slots[s];
      if (w == 0ull) continue;
      k = (long long)(int)(unsigned)w;
      sz = (long long)(w >> 32);
    } else {
      k = old_t.slots[2 * s];
      if (k == kEmptyKey) continue;
      sz = old_t.slots[2 * s + 1];
    }
    int64_t slot = (int64_t)(table_mix64((uint64_t)k) & (uint64_t)mask);
    if (new_t.narrow) {
      const unsigned long long want = ((unsigned long long)sz << 32) | (unsigned long long)(unsigned)k;
      const int64_t bmask = (new_t.capacity >> 2) - 1;
      int64_t b = (int64_t)((uint64_t)table_mix32((uint32_t)(int32_t)k) & (uint64_t)bmask);
      bool placed = false;
      while (!placed) {
        for (int j = 0; j < 4 && !placed; ++j)
          placed = (atomicCAS(nb + 4 * b + j, 0ull, want) == 0ull);
        b = (b + 1) & bmask;
      }
    } else {
      while ((long long)atomicCAS(nb + 2 * slot, (unsigned long long)kEmptyKey, (unsigned long long)k) != kEmptyKey)
        slot = (slot + 1) & mask;
      new_t.slots[2 * slot + 1] = sz;
      if (!old_t.narrow)
        for (int j = 0; j < old_t.n_agg * 4; ++j)
          new_t.vals[slot * old_t.n_agg * 4 + j] = old_t.vals[s * old_t.n_agg * 4 + j];
    }
  }
}

// compaction: table -> dense (unordered) arrays.  A CTA compacts 1024 slots at a time and
// reserves its output range with ONE atomic: a per-warp atomicAdd on the single cursor
// (500 k same-address atomics for a 16 M-slot table) serialised at ~1 ns each and cost
// ~20x the time the scan itself needs.
__global__ void __launch_bounds__(kThreads)
export_kernel(Table t, int64_t* __restrict__ keys_out,
              int64_t* __restrict__ sizes_out, double* __restrict__ vals_out,
              unsigned long long* cursor) {
  __shared__ unsigned s_warp[kThreads / 32];
  __shared__ unsigned long long s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int64_t kChunk = (int64_t)kThreads * kExportPerThread;
  // capacity is a power of two >= 65536: every chunk is full
  for (int64_t c0 = (int64_t)blockIdx.x * kChunk; c0 < t.capacity; c0 += (int64_t)gridDim.x * kChunk) {
    long long k[kExportPerThread], sz[kExportPerThread];
    unsigned live = 0;
#pragma unroll
    for (int j = 0; j < kExportPerThread; ++j) {
      const int64_t s = c0 + (int64_t)j * kThreads + threadIdx.x;
      sz[j] = 0;
      if (t.narrow) {
        const unsigned long long w = (unsigned long long)t.slots[s];
        if (w != 0ull) live |= 1u << j;
        k[j] = (long long)(int)(unsigned)w;
        sz[j] = (long long)(w >> 32);
      } else {
        k[j] = t.slots[2 * s];
        if (k[j] != kEmptyKey) { live |= 1u << j; sz[j] = t.slots[2 * s + 1]; }
      }
    }
    const unsigned mine = __popc(live);
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned tot = 0;
      for (int w = 0; w < kThreads / 32; ++w) { const unsigned x = s_warp[w]; s_warp[w] = tot; tot += x; }
      s_base = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    int64_t o = (int64_t)s_base + s_warp[warp] + (incl - mine);
#pragma unroll
    for (int j = 0; j < kExportPerThread; ++j) {
      if (!((live >> j) & 1u)) continue;
      keys_out[o] = k[j];
      if (sizes_out) sizes_out[o] = sz[j];
      if (vals_out) {
        const int64_t s = c0 + (int64_t)j * kThreads + threadIdx.x;
        for (int q = 0; q < t.n_agg; ++q) {
          const double* v = t.vals + (s * t.n_agg + q) * 4;
          double* w = vals_out + (o * t.n_agg + q) * 4;
          w[0] = v[0]; w[1] = v[1];
          w[2] = dec_ordered(reinterpret_cast<const int64_t*>(v)[2]);
          w[3] = dec_ordered(reinterpret_cast<const int64_t*>(v)[3]);
        }
      }
      ++o;
    }
    __syncthreads();     // s_warp / s_base are reused by the next chunk
  }
}

__global__ void decode_special_kernel(const double* special_vals, int n_agg,
                                      double* out) {
  const int i = threadIdx.x;
  if (i < 2 * n_agg) {
    const double* v = special_vals + i * 4;
    double* w = out + i * 4;
    w[0] = v[0]; w[1] = v[1];
    w[2] = dec_ordered(reinterpret_cast<const int64_t*>(v)[2]);
    w[3] = dec_ordered(reinterpret_cast<const int64_t*>(v)[3]);
  }
}

// ---------------------------------------------------------------------------
// owner partition / gathers / key packing
// ---------------------------------------------------------------------------
__device__ __forceinline__ int owner_of(int64_t key, int n_parts) {
  // bits disjoint from both the global and the smem slot bits
  return (int)((table_mix64((uint64_t)key) >> 52) % (uint64_t)n_parts);
}

__global__ void __launch_bounds__(kThreads)
owner_count_kernel(const int64_t* __restrict__ keys, int64_t n, int n_parts,
                   unsigned long long* counts) {
  __shared__ unsigned int sc[64];
  if (threadIdx.x < 64) sc[threadIdx.x] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(&sc[owner_of(keys[i], n_parts)], 1u);
  __syncthreads();
  if (threadIdx.x < n_parts && sc[threadIdx.x])
    atomicAdd(&counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
}

// exclusive prefix of the per-owner counts -> write cursors (n_parts <= 64: one thread)
__global__ void owner_prefix_kernel(const unsigned long long* __restrict__ counts, int n_parts,
                                    unsigned long long* __restrict__ cursors) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long acc = 0;
    for (int p = 0; p < n_parts; ++p) { cursors[p] = acc; acc += counts[p]; }
  }
}

// scatter pass: each CTA ranks its chunk of rows per owner in shared memory and reserves
// ONE contiguous range per owner with a single global atomic, instead of one global
// atomic per row on only `n_parts` addresses (which serialised at ~1 row/ns).
__global__ void __launch_bounds__(kThreads)
owner_scatter_kernel(const int64_t* __restrict__ keys, int64_t n, int n_parts,
                     unsigned long long* cursors, int64_t* __restrict__ perm) {
  constexpr int kPer = 8;                       // rows per thread per chunk
  __shared__ unsigned int s_cnt[64];
  __shared__ unsigned long long s_base[64];
  const int64_t chunk = (int64_t)kThreads * kPer;
  const int64_t n_chunks = (n + chunk - 1) / chunk;
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    int own[kPer];
    unsigned rank[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int64_t i = c * chunk + (int64_t)j * kThreads + threadIdx.x;
      own[j] = -1;
      if (i < n) {
        own[j] = owner_of(keys[i], n_parts);
        rank[j] = atomicAdd(&s_cnt[own[j]], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < n_parts && s_cnt[threadIdx.x])
      s_base[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int64_t i = c * chunk + (int64_t)j * kThreads + threadIdx.x;
      if (own[j] >= 0) perm[s_base[own[j]] + rank[j]] = i;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads)
gather_i64_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ perm,
                  int64_t n, int64_t* __restrict__ dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = src[perm[i]];
}

__global__ void __launch_bounds__(kThreads)
gather_f64_rows_kernel(const double* __restrict__ src, const int64_t* __restrict__ perm,
                       int64_t n, int w, double* __restrict__ dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t total = n * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / w, c = i - r * w;
    dst[i] = src[perm[r] * w + c];
  }
}

__global__ void __launch_bounds__(kThreads)
pack_keys2_kernel(const int32_t* __restrict__ a, const uint8_t* __restrict__ ma,
                  const int32_t* __restrict__ b, const uint8_t* __restrict__ mb,
                  int64_t n, int64_t* __restrict__ out, uint8_t* __restrict__ vout) {
  // one thread per 8 rows so each thread owns one validity byte
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n8 = (n + 7) / 8;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += stride) {
    unsigned vb = 0;
    for (int k = 0; k < 8; ++k) {
      const int64_t i = g * 8 + k;
      if (i >= n) break;
      const bool va = valid1(ma, i), vb_ = valid1(mb, i);
      const int32_t x = va ? a[i] : INT32_MIN;
      const int32_t y = vb_ ? b[i] : INT32_MIN;
      out[i] = (int64_t)(((uint64_t)(uint32_t)x << 32) |
                         (uint64_t)((uint32_t)y ^ 0x80000000u));
      if (va || vb_) vb |= 1u << k;
    }
    if (vout) vout[g] = (uint8_t)vb;
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int plain_grid(int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)sm_count() * 8));
}

static int table_alloc(Table* t, int64_t capacity, int n_agg, bool narrow, cudaStream_t st) {
  t->capacity = capacity;
  t->n_agg = n_agg;
  t->narrow = narrow ? 1 : 0;
  t->slots = nullptr;
  t->vals = nullptr;
  NVTB_CUDA_OK(cudaMallocAsync(&t->slots, sizeof(int64_t) * (narrow ? 1 : 2) * capacity, st));
  if (n_agg > 0)
    NVTB_CUDA_OK(cudaMallocAsync(&t->vals, sizeof(double) * 4 * n_agg * capacity, st));
  table_init_kernel<<<plain_grid(capacity), kThreads, 0, st>>>(*t);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

static int table_free(Table* t, cudaStream_t st) {
  if (t->slots) NVTB_CUDA_OK(cudaFreeAsync(t->slots, st));
  if (t->vals) NVTB_CUDA_OK(cudaFreeAsync(t->vals, st));
  t->slots = nullptr; t->vals = nullptr;
  return NVTB_OK;
}

static int64_t next_pow2(int64_t v) {
  int64_t p = kMinCapacity;
  while (p < v) p <<= 1;
  return p;
}

static int settle(nvtb_hashagg* h);

static int arena_alloc(Arena* a, int64_t cap, int n_agg, cudaStream_t st) {
  a->cap = cap; a->keys = nullptr; a->sizes = nullptr; a->vals = nullptr; a->slot = -1;
  if (cap <= 0) return NVTB_OK;
  NVTB_CUDA_OK(cudaMallocAsync(&a->keys, sizeof(int64_t) * cap, st));
  NVTB_CUDA_OK(cudaMallocAsync(&a->sizes, sizeof(int64_t) * cap, st));
  if (n_agg > 0) NVTB_CUDA_OK(cudaMallocAsync(&a->vals, sizeof(double) * 4 * n_agg * cap, st));
  return NVTB_OK;
}

// The arena of a launch must be able to hold one pair per input row, but is
// almost never touched.  Two pooled arenas (double buffering) serve every
// handle: acquiring a slot that still belongs to an unsettled launch settles
// that launch first, so at most two arenas exist however many columns are fitted.
struct ArenaSlot {
  int64_t* keys; int64_t* sizes; double* vals;
  int64_t cap_alloc; int64_t vals_alloc;
  nvtb_hashagg* owner;
  uint64_t stamp;
};
static ArenaSlot g_slots[2] = {};
static uint64_t g_stamp = 0;
static std::recursive_mutex g_arena_mu;

static int arena_acquire(nvtb_hashagg* h, Arena* out, int64_t cap, int n_agg) {
  std::lock_guard<std::recursive_mutex> lk(g_arena_mu);
  int pick = -1;
  for (int i = 0; i < 2; ++i)
    if (g_slots[i].owner == nullptr) { pick = i; break; }
  if (pick < 0) {
    pick = g_slots[0].stamp < g_slots[1].stamp ? 0 : 1;
    int rc = settle(g_slots[pick].owner);   // releases the slot
    if (rc) return rc;
  }
  ArenaSlot& sl = g_slots[pick];
  const int64_t need_vals = cap * 4 * n_agg;
  if (sl.cap_alloc < cap) {
    NVTB_CUDA_OK(cudaDeviceSynchronize());
    if (sl.keys) cudaFree(sl.keys);
    if (sl.sizes) cudaFree(sl.sizes);
    sl.keys = nullptr; sl.sizes = nullptr; sl.cap_alloc = 0;
    NVTB_CUDA_OK(cudaMalloc(&sl.keys, sizeof(int64_t) * cap));
    NVTB_CUDA_OK(cudaMalloc(&sl.sizes, sizeof(int64_t) * cap));
    sl.cap_alloc = cap;
  }
  if (sl.vals_alloc < need_vals) {
    NVTB_CUDA_OK(cudaDeviceSynchronize());
    if (sl.vals) cudaFree(sl.vals);
    sl.vals = nullptr; sl.vals_alloc = 0;
    NVTB_CUDA_OK(cudaMalloc(&sl.vals, sizeof(double) * need_vals));
    sl.vals_alloc = need_vals;
  }
  sl.owner = h;
  sl.stamp = ++g_stamp;
  out->keys = sl.keys; out->sizes = sl.sizes; out->vals = n_agg > 0 ? sl.vals : nullptr;
  out->cap = cap; out->slot = pick;
  return NVTB_OK;
}

static int arena_free(Arena* a, cudaStream_t st) {
  if (a->slot >= 0) {
    std::lock_guard<std::recursive_mutex> lk(g_arena_mu);
    g_slots[a->slot].owner = nullptr;
  } else {
    if (a->keys) NVTB_CUDA_OK(cudaFreeAsync(a->keys, st));
    if (a->sizes) NVTB_CUDA_OK(cudaFreeAsync(a->sizes, st));
    if (a->vals) NVTB_CUDA_OK(cudaFreeAsync(a->vals, st));
  }
  a->keys = nullptr; a->sizes = nullptr; a->vals = nullptr; a->cap = 0; a->slot = -1;
  return NVTB_OK;
}

// distinct keys expected among `rows` uniform draws from K values
static double expected_unique(double K, double rows) {
  if (K <= 0) return 0;
  return K * -std::expm1(-rows / K);
}

// invert U = K (1 - exp(-s/K)) for K (bisection); U >= 0.98 s  =>  "all distinct"
static double estimate_cardinality(double U, double s) {
  if (U <= 0) return 1.0;
  if (U >= 0.98 * s) return 1e18;
  double lo = U, hi = 1e18;
  for (int it = 0; it < 200; ++it) {
    const double mid = std::sqrt(lo * hi);
    if (expected_unique(mid, s) < U) lo = mid; else hi = mid;
    if (hi / lo < 1.0 + 1e-9) break;
  }
  return lo;
}

// move to a table of `new_cap` slots and/or the other layout (narrow -> wide only)
static int grow_to(nvtb_hashagg* h, int64_t new_cap, cudaStream_t st, bool force_wide = false) {
  const bool narrow = h->t.narrow && !force_wide;
  new_cap = std::max(new_cap, h->t.capacity);
  if (new_cap == h->t.capacity && narrow == (bool)h->t.narrow) return NVTB_OK;
  Table nt;
  int rc = table_alloc(&nt, new_cap, h->n_agg, narrow, st);
  if (rc) return rc;
  if (h->u_known > 0 || h->rows_total > 0) {
    rehash_kernel<<<plain_grid(h->t.capacity), kThreads, 0, st>>>(h->t, nt);
    NVTB_LAUNCH_OK();
  }
  rc = table_free(&h->t, st);
  if (rc) return rc;
  h->t = nt;
  return NVTB_OK;
}

static int launch_merge(nvtb_hashagg* h, const int64_t* keys, const int64_t* sizes,
                        const double* vals, int64_t n, const Arena& arena, cudaStream_t st) {
  const int grid = plain_grid(n);
  arm_launch_kernel<<<1, 1, 0, st>>>(h->ctr);
  NVTB_LAUNCH_OK();
  if (h->t.narrow)
    merge_kernel<true><<<grid, kThreads, 0, st>>>(keys, sizes, vals, n, h->t, h->ctr, h->special_vals, arena);
  else
    merge_kernel<false><<<grid, kThreads, 0, st>>>(keys, sizes, vals, n, h->t, h->ctr, h->special_vals, arena);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

// wait for the pending launch, read its counters, and fold a non-empty arena
// back in after growing the table.  Leaves no pending state.
static int settle(nvtb_hashagg* h) {
  if (h == nullptr) return NVTB_OK;
  while (h->pending) {
    cudaStream_t st = h->pending_stream;
    NVTB_CUDA_OK(cudaEventSynchronize(h->ev));
    h->pending = false;
    h->mailbox_valid = true;
    const Counters c = *h->mailbox;
    h->u_known = (int64_t)c.n_unique;
    const int64_t ovf = (int64_t)std::min<unsigned long long>(c.ovf_count, (unsigned long long)h->arena.cap);
    Arena old = h->arena;
    h->arena = Arena{nullptr, nullptr, nullptr, 0, -1};
    if (ovf > 0) {
      // every refused pair may be a new key: size for all of them at load <= 0.25
      int rc = grow_to(h, next_pow2(4 * (h->u_known + ovf)), st);
      if (rc) return rc;
      // a pair that is still refused (128
      // probes) lands in a fresh private arena and is settled by the loop
      rc = arena_alloc(&h->arena, ovf, h->n_agg, st);
      if (rc) return rc;
      rc = launch_merge(h, old.keys, old.sizes, old.vals, ovf, h->arena, st);
      if (rc) return rc;
      NVTB_CUDA_OK(cudaMemcpyAsync(h->mailbox, h->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
      NVTB_CUDA_OK(cudaEventRecord(h->ev, st));
      h->pending = true;
      h->pending_stream = st;
    }
    int rc = arena_free(&old, st);
    if (rc) return rc;
  }
  if (h->rows_total > 0) {
    // draws = VALID rows: the null rows of the sample are not draws from the key distribution.
    // (Counting them made a column with 8 % nulls look 8 % "duplicated": a 4e7-key column was
    // estimated at 5.7e6 keys from its first 2^20 rows, its first batch filled the table and
    // took the 128-bucket probe path — three 23-55 ms launches in the cold fit of the Criteo bench.)
    int64_t draws = h->rows_total;
    if (h->mailbox_valid) draws -= (int64_t)h->mailbox->size[0];
    h->k_est = estimate_cardinality((double)std::max<int64_t>(h->u_known, 1), (double)std::max<int64_t>(draws, 1));
  }
  return NVTB_OK;
}

// after a launch: async readback of the counters
static int post(nvtb_hashagg* h, cudaStream_t st) {
  NVTB_CUDA_OK(cudaMemcpyAsync(h->mailbox, h->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  NVTB_CUDA_OK(cudaEventRecord(h->ev, st));
  h->pending = true;
  h->mailbox_valid = false;
  h->pending_stream = st;
  return NVTB_OK;
}

// distinct keys expected in the table after `rows` more rows (-> h->predicted)
static void predict(nvtb_hashagg* h, int64_t rows) {
  double predicted;
  if (h->hint > 0) {
    predicted = (double)std::max<int64_t>(h->hint, h->u_known);
  } else if (h->k_est > 0) {
    predicted = expected_unique(h->k_est, (double)(h->rows_total + rows));
    predicted = std::max(predicted, (double)h->u_known);
  } else {
    predicted = (double)h->u_known + (double)rows;   // no information: worst case
  }
  predicted = std::min(predicted, (double)h->u_known + (double)rows);
  h->predicted = predicted;
}

// size the table for `rows` more rows: load <= 0.5 for the ESTIMATED distinct count
static int prepare(nvtb_hashagg* h, int64_t rows, cudaStream_t st) {
  predict(h, rows);
  const int64_t want = next_pow2((int64_t)(2.5 * h->predicted) + 1);
  return grow_to(h, want, st);
}

struct PartScratch { void* ptr; size_t bytes; cudaEvent_t ev; cudaStream_t last; bool used; };
static PartScratch g_part = {nullptr, 0, nullptr, nullptr, false};

// int32 keys, narrow table, no payload: shared-memory fold, hash-partitioned first when the
// expected number of distinct keys exceeds what one SM's shared memory holds (fold_i32.cuh)
static int launch_fold_i32(nvtb_hashagg* h, const int32_t* kp, const uint8_t* mp, int64_t m,
                           cudaStream_t st) {
  static bool attrs = false;
  constexpr int kDirectSmem = (int)(kFoldBucketsDirect * kFoldWays * kFoldSlotBytes);
  constexpr int kPartsSmem = (int)(kFoldBucketsParts * kFoldWays * kFoldSlotBytes);
  constexpr int kScatterSmemMax = kPartTile * 4 + 2 * 4 * kMaxParts;
  if (!attrs) {
    NVTB_CUDA_OK(cudaFuncSetAttribute(fold_i32_kernel<kFoldThreadsDirect, 1, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kDirectSmem));
    NVTB_CUDA_OK(cudaFuncSetAttribute(fold_i32_kernel<kFoldThreadsParts, 2, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kPartsSmem));
    NVTB_CUDA_OK(cudaFuncSetAttribute(part_scatter_kernel<PartHashTop>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmemMax));
    attrs = true;
  }
  const int sms = sm_count();
  const bool aligned = is_aligned32(kp);
  // distinct keys this batch is expected to show (the table may already hold some of them)
  const double fresh = std::min(h->predicted, (double)m);
  const bool parts = aligned && m >= ((int64_t)1 << 18) && m < (int64_t)0xFFFF0000ll &&
                     fresh > kFoldMaxLoad * (double)(kFoldWays * kFoldBucketsDirect);
  if (!parts) {
    constexpr int64_t kStep = (int64_t)kFoldThreadsDirect * 8;
    int64_t units = std::min<int64_t>(sms, (m + kStep - 1) / kStep);
    int64_t chunk = ((m + units - 1) / units + kStep - 1) / kStep * kStep;
    units = (m + chunk - 1) / chunk;
    fold_i32_kernel<kFoldThreadsDirect, 1, false><<<(int)units, kFoldThreadsDirect, kDirectSmem, st>>>(
        kp, mp, m, nullptr, nullptr, (int)units, chunk, 0, kFoldBucketsDirect, aligned ? 1 : 0,
        h->t, h->ctr, h->arena);
    NVTB_LAUNCH_OK();
    return NVTB_OK;
  }
  int lg = 9;
  while ((1 << lg) < kMinParts) ++lg;
  while ((1 << lg) < kMaxParts &&
         fresh / (double)(1 << lg) > kFoldMaxLoad * (double)(kFoldWays * kFoldBucketsParts)) ++lg;
  const int P = 1 << lg;
  // partition buffer: ONE grow-only device allocation shared by every handle (a fresh
  // 268 MB cudaMallocAsync per column occasionally costs tens of ms when the pool has to
  // map new memory).  Uses are ordered by an event when the stream changes.
  uint32_t* meta = nullptr;      // total[P] | starts[P] | cursor[P]
  int32_t* buf = nullptr;
  {
    std::lock_guard<std::recursive_mutex> lk(g_arena_mu);
    const size_t need = sizeof(uint32_t) * 3 * kMaxParts + sizeof(int32_t) * (size_t)(m + 8 * (int64_t)kMaxParts);
    if (g_part.bytes < need) {
      NVTB_CUDA_OK(cudaDeviceSynchronize());
      if (g_part.ptr) cudaFree(g_part.ptr);
      g_part.ptr = nullptr; g_part.bytes = 0;
      NVTB_CUDA_OK(cudaMalloc(&g_part.ptr, need));
      g_part.bytes = need;
    }
    if (g_part.ev == nullptr) NVTB_CUDA_OK(cudaEventCreateWithFlags(&g_part.ev, cudaEventDisableTiming));
    if (g_part.used && g_part.last != st) NVTB_CUDA_OK(cudaStreamWaitEvent(st, g_part.ev, 0));
    meta = reinterpret_cast<uint32_t*>(g_part.ptr);
    buf = reinterpret_cast<int32_t*>(meta + 3 * kMaxParts);
  }
  NVTB_CUDA_OK(cudaMemsetAsync(meta, 0, sizeof(uint32_t) * P, st));
  const int64_t tiles = (m + kPartTile - 1) / kPartTile;
  part_hist_kernel<PartHashTop><<<(int)std::min<int64_t>(tiles, 3 * sms), kPartThreads, 4 * P, st>>>(
      kp, mp, m, PartHashTop{lg}, meta, h->ctr, 1);
  NVTB_LAUNCH_OK();
  part_scan_kernel<<<1, kPartThreads, 0, st>>>(meta, lg, meta + P, meta + 2 * P, 8u, nullptr);
  NVTB_LAUNCH_OK();
  part_scatter_kernel<PartHashTop><<<(int)std::min<int64_t>(tiles, 2 * sms), kPartThreads, kPartTile * 4 + 2 * 4 * P, st>>>(
      kp, mp, m, PartHashTop{lg}, meta + 2 * P, buf, 1);
  NVTB_LAUNCH_OK();
  fold_i32_kernel<kFoldThreadsParts, 2, true><<<std::min(P, 2 * sms), kFoldThreadsParts, kPartsSmem, st>>>(
      buf, nullptr, m + 8 * (int64_t)P, meta + P, meta + 2 * P, P, 0, lg, kFoldBucketsParts, 1,
      h->t, h->ctr, h->arena);
  NVTB_LAUNCH_OK();
  {
    std::lock_guard<std::recursive_mutex> lk(g_arena_mu);
    NVTB_CUDA_OK(cudaEventRecord(g_part.ev, st));
    g_part.used = true;
    g_part.last = st;
  }
  return NVTB_OK;
}


// ---------------------------------------------------------------------------------------
// sorted accumulator (sortagg.cuh): host side
// ---------------------------------------------------------------------------------------
// expected distinct keys above which an int32 column leaves the hash table for the sorted
// accumulator: the table (2.5 slots x 8 B per key) then no longer stays in the 126 MB L2
static int64_t runs_min_keys() {
  const char* e = getenv("NVTB_RUNS_MIN_KEYS");      // read every time: tests flip it
  int64_t v = e ? atoll(e) : ((int64_t)8 << 20);
  return v < 1 ? 1 : v;
}

// ONE grow-only device scratch for the sort pipeline, shared by every handle of the process
// (one process per GPU); uses on different streams are ordered by an event, like g_part
struct SortScratch { void* ptr; size_t bytes; cudaEvent_t ev; cudaStream_t last; bool used; };
static SortScratch g_sort = {nullptr, 0, nullptr, nullptr, false};

struct SortCarve {
  unsigned int* rx_zero;     // 256 B kept at zero (radix "done" counter)
  void* rx;                  // radix scratch (starts at rx_zero)
  uint32_t* part_meta;       // total[P] | starts[P] | cursor[P]
  uint32_t* keys_a;
  uint32_t* keys_b;
  uint64_t* rle;             // [m + 1]
  uint32_t* tile_heads;      // [ceil(m / kRleTile)]
  uint2* splits;             // [MT + 1]
  uint32_t* tile_out;        // [MT]
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int sort_scratch_acquire(int64_t m, int64_t n_pairs_sort, int64_t mt, cudaStream_t st, SortCarve* c) {
  const int P = 1 << kSortLowBits;
  const size_t rx_bytes = align_up(std::max(rx_scratch_bytes<uint32_t>(m, kRxMaxStableBits - 1),
                                            rx_scratch_bytes<uint64_t>(n_pairs_sort, kRxMaxStableBits - 1)), 256);
  // radix path: total[P] | starts[P] | cursor[P]; bucket path (bucketagg.cuh): the same three with
  // 8192 entries + distinct[8192] + {min, max, lo, shift, flag}
  const size_t meta_bytes = align_up(sizeof(uint32_t) * (size_t)std::max(3 * P, 4 * kBkParts + 64), 256);
  const size_t keys_bytes = align_up(sizeof(uint32_t) * (size_t)(m + 64), 256);
  const size_t rle_bytes = align_up(sizeof(uint64_t) * (size_t)(m + 2), 256);
  const size_t heads_bytes = align_up(sizeof(uint32_t) * (size_t)((m + kRleTile - 1) / kRleTile + 1), 256);
  const size_t splits_bytes = align_up(sizeof(uint2) * (size_t)(mt + 2), 256);
  const size_t out_bytes = align_up(sizeof(uint32_t) * (size_t)(mt + 2), 256);
  const size_t need = rx_bytes + meta_bytes + 2 * keys_bytes + rle_bytes + heads_bytes + splits_bytes + out_bytes;
  std::lock_guard<std::recursive_mutex> lk(g_arena_mu);
  if (g_sort.bytes < need) {
    NVTB_CUDA_OK(cudaDeviceSynchronize());
    if (g_sort.ptr) cudaFree(g_sort.ptr);
    g_sort.ptr = nullptr; g_sort.bytes = 0;
    const size_t want = need + need / 8;
    NVTB_CUDA_OK(cudaMalloc(&g_sort.ptr, want));
    g_sort.bytes = want;
    NVTB_CUDA_OK(cudaMemsetAsync(g_sort.ptr, 0, 256, st));
    g_sort.used = false;
  }
  if (g_sort.ev == nullptr) NVTB_CUDA_OK(cudaEventCreateWithFlags(&g_sort.ev, cudaEventDisableTiming));
  if (g_sort.used && g_sort.last != st) NVTB_CUDA_OK(cudaStreamWaitEvent(st, g_sort.ev, 0));
  char* p = reinterpret_cast<char*>(g_sort.ptr);
  c->rx_zero = reinterpret_cast<unsigned int*>(p);
  c->rx = p;                                         p += rx_bytes;
  c->part_meta = reinterpret_cast<uint32_t*>(p);     p += meta_bytes;
  c->keys_a = reinterpret_cast<uint32_t*>(p);        p += keys_bytes;
  c->keys_b = reinterpret_cast<uint32_t*>(p);        p += keys_bytes;
  c->rle = reinterpret_cast<uint64_t*>(p);           p += rle_bytes;
  c->tile_heads = reinterpret_cast<uint32_t*>(p);    p += heads_bytes;
  c->splits = reinterpret_cast<uint2*>(p);           p += splits_bytes;
  c->tile_out = reinterpret_cast<uint32_t*>(p);
  return NVTB_OK;
}

static int sort_scratch_release(cudaStream_t st) {
  std::lock_guard<std::recursive_mutex> lk(g_arena_mu);
  NVTB_CUDA_OK(cudaEventRecord(g_sort.ev, st));
  g_sort.used = true;
  g_sort.last = st;
  return NVTB_OK;
}

// make sure acc[which] can hold `pairs` packed pairs (contents are NOT preserved)
static int acc_reserve(nvtb_hashagg* h, int which, int64_t pairs, cudaStream_t st) {
  if (h->acc_cap[which] >= pairs) return NVTB_OK;
  if (h->acc[which]) NVTB_CUDA_OK(cudaFreeAsync(h->acc[which], st));
  h->acc[which] = nullptr; h->acc_cap[which] = 0;
  const int64_t want = pairs + pairs / 8 + 1024;
  NVTB_CUDA_OK(cudaMallocAsync(&h->acc[which], sizeof(uint64_t) * (size_t)want, st));
  h->acc_cap[which] = want;
  return NVTB_OK;
}

// hash table -> sorted accumulator (the handle must be settled and its table narrow).
// One-off: export the u_known (key, count) pairs packed and sort them by key.
static int table_to_runs(nvtb_hashagg* h, cudaStream_t st) {
  const int64_t nu = h->u_known;
  if (h->d_n == nullptr) {
    NVTB_CUDA_OK(cudaMalloc(&h->d_n, sizeof(uint32_t) * 4));
    NVTB_CUDA_OK(cudaMemsetAsync(h->d_n, 0, sizeof(uint32_t) * 4, st));
  }
  h->acc_cur = 0;
  if (nu > 0) {
    int rc = acc_reserve(h, 0, nu, st);
    if (rc) return rc;
    rc = acc_reserve(h, 1, nu, st);
    if (rc) return rc;
    SortCarve c;
    rc = sort_scratch_acquire(64, nu, 1, st, &c);
    if (rc) return rc;
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(c.part_meta);
    NVTB_CUDA_OK(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st));
    table_to_pairs_kernel<<<plain_grid(h->t.capacity / kExportPerThread), kThreads, 0, st>>>(
        h->t, h->acc[0], cursor, &h->ctr->max_count);
    NVTB_LAUNCH_OK();
    int in_b = 0;
    rc = rx_sort_bits<uint64_t>(h->acc[0], h->acc[1], nullptr, nu, 32, 64, false, c.rx, st, &in_b);
    if (rc) return rc;
    h->acc_cur = in_b;
    rc = sort_scratch_release(st);
    if (rc) return rc;
  }
  // the table itself is no longer used: keep a minimal one so that the handle stays uniform
  int rc = table_free(&h->t, st);
  if (rc) return rc;
  rc = table_alloc(&h->t, kMinCapacity, 0, true, st);
  if (rc) return rc;
  h->mode = 1;
  return NVTB_OK;
}

// fold one batch of int32 keys into the sorted accumulator (handle settled: u_known exact)
static int launch_runs_insert(nvtb_hashagg* h, const int32_t* kp, const uint8_t* mp, int64_t m, cudaStream_t st) {
  static bool attrs = false;
  constexpr int kScatterSmem = kPartTile * 4 + 2 * 4 * (1 << kSortLowBits);
  if (!attrs) {
    NVTB_CUDA_OK(cudaFuncSetAttribute(part_scatter_kernel<PartKeyLow>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem));
    NVTB_CUDA_OK(cudaFuncSetAttribute(merge_write_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeSmem)));
    NVTB_CUDA_OK(cudaFuncSetAttribute(merge_write_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeSmem)));
    attrs = true;
  }
  const int64_t ua = h->u_known;
  const int other = h->acc_cur ^ 1;
  int rc = NVTB_OK;
  const int64_t mt = (ua + m + kMergeTile - 1) / kMergeTile;
  SortCarve c;
  rc = sort_scratch_acquire(m, 64, mt, st, &c);
  if (rc) return rc;
  const int sms = sm_count();
  const int P = 1 << kSortLowBits;
  const int aligned = is_aligned32(kp) ? 1 : 0;
  uint32_t* n_valid = h->d_n + 1;
  uint32_t* n_batch = h->d_n + 2;
  const int64_t tiles = (m + kPartTile - 1) / kPartTile;
  const uint64_t* A = h->acc[h->acc_cur];
  // ---- bucket path (bucketagg.cuh): range partition + direct-address counting, no sort ----------
  bool done = false;
  const char* path_env = getenv("NVTB_SORT_PATH");
  if (!(path_env && strcmp(path_env, "radix") == 0)) {
    static bool bk_attrs = false;
    constexpr int kBkScatterSmem = kPartTile * 4 + 2 * 4 * kBkParts;
    if (!bk_attrs) {
      NVTB_CUDA_OK(cudaFuncSetAttribute(part_hist_kernel<PartRange>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kBkParts));
      NVTB_CUDA_OK(cudaFuncSetAttribute(part_scatter_kernel<PartRange>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBkScatterSmem));
      NVTB_CUDA_OK(cudaFuncSetAttribute(bk_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBkCountSmem));
      NVTB_CUDA_OK(cudaFuncSetAttribute(bk_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBkEmitSmem));
      bk_attrs = true;
    }
    uint32_t* total = c.part_meta;                    // -> exclusive starts after the scan
    uint32_t* cursor = c.part_meta + kBkParts;
    uint32_t* distinct = c.part_meta + 2 * kBkParts;  // -> output offsets after the scan
    uint32_t* mm = c.part_meta + 4 * kBkParts;        // {min, max}
    uint32_t* par = mm + 2;                           // {lo, shift}
    unsigned int* flag = reinterpret_cast<unsigned int*>(mm + 4);
    const uint32_t mm_init[6] = {0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u};
    NVTB_CUDA_OK(cudaMemcpyAsync(mm, mm_init, sizeof(mm_init), cudaMemcpyHostToDevice, st));
    NVTB_CUDA_OK(cudaMemsetAsync(total, 0, sizeof(uint32_t) * kBkParts, st));
    bk_minmax_kernel<<<(int)std::min<int64_t>(tiles, 4 * sms), kPartThreads, 0, st>>>(kp, mp, m, mm, aligned);
    NVTB_LAUNCH_OK();
    bk_params_kernel<<<1, 1, 0, st>>>(mm, par);
    NVTB_LAUNCH_OK();
    const PartRange pol{kBkLgParts, par};
    part_hist_kernel<PartRange><<<(int)std::min<int64_t>(tiles, 3 * sms), kPartThreads, 4 * kBkParts, st>>>(
        kp, mp, m, pol, total, h->ctr, aligned);
    NVTB_LAUNCH_OK();
    scan_tiles_kernel<<<1, kRunThreads, 0, st>>>(total, kBkParts, n_valid, nullptr);
    NVTB_LAUNCH_OK();
    NVTB_CUDA_OK(cudaMemcpyAsync(cursor, total, sizeof(uint32_t) * kBkParts, cudaMemcpyDeviceToDevice, st));
    part_scatter_kernel<PartRange><<<(int)std::min<int64_t>(tiles, sms), kPartThreads, kBkScatterSmem, st>>>(
        kp, mp, m, pol, cursor, reinterpret_cast<int32_t*>(c.keys_a), aligned);
    NVTB_LAUNCH_OK();
    bk_count_kernel<<<kBkParts, kBkThreads, kBkCountSmem, st>>>(c.keys_a, total, n_valid, par, distinct);
    NVTB_LAUNCH_OK();
    scan_tiles_kernel<<<1, kRunThreads, 0, st>>>(distinct, kBkParts, n_batch, ua == 0 ? &h->ctr->n_unique : nullptr);
    NVTB_LAUNCH_OK();
    // the accumulator is sized for what the batch really holds (its distinct keys are known now),
    // not for the worst case of all rows distinct: 1.3 GB instead of 2.2 GB per buffer at 2.5e8 rows
    uint32_t nb_h = 0;
    NVTB_CUDA_OK(cudaMemcpyAsync(&nb_h, n_batch, sizeof(nb_h), cudaMemcpyDeviceToHost, st));
    NVTB_CUDA_OK(cudaStreamSynchronize(st));
    rc = acc_reserve(h, other, ua + (int64_t)nb_h, st);
    if (rc) return rc;
    // the batch's pairs go straight into the accumulator when it is empty, else to the merge input
    uint64_t* B = (ua == 0) ? h->acc[other] : c.rle;
    bk_emit_kernel<<<kBkParts, kBkThreads, kBkEmitSmem, st>>>(c.keys_a, total, n_valid, par, distinct, B, flag,
                                                               &h->ctr->max_count);
    NVTB_LAUNCH_OK();
    unsigned int flag_h = 0;
    NVTB_CUDA_OK(cudaMemcpyAsync(&flag_h, flag, sizeof(flag_h), cudaMemcpyDeviceToHost, st));
    NVTB_CUDA_OK(cudaStreamSynchronize(st));
    if (flag_h == 0) {
      if (ua > 0) {
        merge_split_kernel<<<(int)((mt + 1 + 255) / 256), 256, 0, st>>>(A, (uint32_t)ua, B, n_batch, (int)mt, c.splits);
        NVTB_LAUNCH_OK();
        merge_count_kernel<<<(int)mt, kRunThreads, 0, st>>>(A, B, c.splits, c.tile_out);
        NVTB_LAUNCH_OK();
        scan_tiles_kernel<<<1, kRunThreads, 0, st>>>(c.tile_out, (int)mt, nullptr, &h->ctr->n_unique);
        NVTB_LAUNCH_OK();
        merge_write_kernel<true><<<(int)mt, kRunThreads, sizeof(MergeSmem), st>>>(A, B, c.splits, c.tile_out, h->acc[other],
                                                                                  &h->ctr->max_count);
        NVTB_LAUNCH_OK();
      }
      done = true;
    }
    // flag set: some window holds more than kBkDupCap duplicated values -> the radix pipeline
    // below redoes the batch (the nulls have been counted already)
  }
  if (!done) {
    rc = acc_reserve(h, other, ua + m, st);
    if (rc) return rc;
    Counters* null_ctr = (path_env && strcmp(path_env, "radix") == 0) ? h->ctr : nullptr;
    // (1) LSD radix sort of the valid keys (as key ^ 2^31)
    NVTB_CUDA_OK(cudaMemsetAsync(c.part_meta, 0, sizeof(uint32_t) * P, st));
    part_hist_kernel<PartKeyLow><<<(int)std::min<int64_t>(tiles, 3 * sms), kPartThreads, 4 * P, st>>>(
        kp, mp, m, PartKeyLow{kSortLowBits}, c.part_meta, null_ctr, aligned);
    NVTB_LAUNCH_OK();
    part_scan_kernel<<<1, kPartThreads, 0, st>>>(c.part_meta, kSortLowBits, c.part_meta + P, c.part_meta + 2 * P, 1u, n_valid);
    NVTB_LAUNCH_OK();
    part_scatter_kernel<PartKeyLow><<<(int)std::min<int64_t>(tiles, 2 * sms), kPartThreads, kScatterSmem, st>>>(
        kp, mp, m, PartKeyLow{kSortLowBits}, c.part_meta + 2 * P, reinterpret_cast<int32_t*>(c.keys_a), aligned);
    NVTB_LAUNCH_OK();
    int in_b = 0;
    rc = rx_sort_bits<uint32_t>(c.keys_a, c.keys_b, n_valid, m, kSortLowBits, 32, false, c.rx, st, &in_b);
    if (rc) return rc;
    const uint32_t* sorted = in_b ? c.keys_b : c.keys_a;
    // (2) run heads
    const int rt = (int)((m + kRleTile - 1) / kRleTile);
    rle_count_kernel<<<rt, kRunThreads, 0, st>>>(sorted, n_valid, m, c.tile_heads);
    NVTB_LAUNCH_OK();
    scan_tiles_kernel<<<1, kRunThreads, 0, st>>>(c.tile_heads, rt, n_batch, nullptr);
    NVTB_LAUNCH_OK();
    rle_write_kernel<<<rt, kRunThreads, 0, st>>>(sorted, n_valid, m, c.tile_heads, n_batch, c.rle);
    NVTB_LAUNCH_OK();
    // (3) merge with the accumulator
    merge_split_kernel<<<(int)((mt + 1 + 255) / 256), 256, 0, st>>>(A, (uint32_t)ua, c.rle, n_batch, (int)mt, c.splits);
    NVTB_LAUNCH_OK();
    merge_count_kernel<<<(int)mt, kRunThreads, 0, st>>>(A, c.rle, c.splits, c.tile_out);
    NVTB_LAUNCH_OK();
    scan_tiles_kernel<<<1, kRunThreads, 0, st>>>(c.tile_out, (int)mt, nullptr, &h->ctr->n_unique);
    NVTB_LAUNCH_OK();
    merge_write_kernel<false><<<(int)mt, kRunThreads, sizeof(MergeSmem), st>>>(A, c.rle, c.splits, c.tile_out, h->acc[other],
                                                                               &h->ctr->max_count);
    NVTB_LAUNCH_OK();
  }
  h->acc_cur = other;
  return sort_scratch_release(st);
}

// rows a sorted accumulator stages before it sorts (NVTB_STAGE_ROWS, default 2^28 = 1 GiB of keys)
static int64_t stage_cap_rows() {
  const char* e = getenv("NVTB_STAGE_ROWS");
  int64_t v = e ? atoll(e) : ((int64_t)1 << 28);
  if (v < 0) v = 0;
  if (v > (int64_t)0xF0000000ll) v = (int64_t)0xF0000000ll;
  return v / 64 * 64;
}

// sort + run-length encode + merge everything that is staged (handle settled on entry; leaves a
// pending launch)
static int post(nvtb_hashagg* h, cudaStream_t st);
static int stage_flush(nvtb_hashagg* h, cudaStream_t st) {
  if (h->stage_rows == 0) return NVTB_OK;
  int rc = settle(h);
  if (rc) return rc;
  if (h->stage_last != st && h->stage_ev) NVTB_CUDA_OK(cudaStreamWaitEvent(st, h->stage_ev, 0));
  rc = launch_runs_insert(h, h->stage_keys, h->stage_mask, h->stage_rows, st);
  if (rc) return rc;
  h->stage_rows = 0;
  NVTB_CUDA_OK(cudaEventRecord(h->stage_ev, st));
  h->stage_last = st;
  return post(h, st);
}

// append one batch to the staging buffers (stage_rows % 8 == 0; the caller flushed when the
// batch does not fit behind the waiting rows)
static int stage_append(nvtb_hashagg* h, const int32_t* kp, const uint8_t* mp, int64_t m, cudaStream_t st) {
  if (h->stage_ev == nullptr) NVTB_CUDA_OK(cudaEventCreateWithFlags(&h->stage_ev, cudaEventDisableTiming));
  if (h->stage_last != nullptr && h->stage_last != st) NVTB_CUDA_OK(cudaStreamWaitEvent(st, h->stage_ev, 0));
  const int64_t cap_max = stage_cap_rows();
  const bool grow_for_fit = h->stage_rows == 0 && h->stage_cap < std::min<int64_t>(cap_max, h->stage_hint);
  if (h->stage_rows + m > h->stage_cap || grow_for_fit) {
    NVTB_REQUIRE(h->stage_rows == 0, "staging buffer resized while rows are waiting");
    if (h->stage_keys) NVTB_CUDA_OK(cudaFreeAsync(h->stage_keys, st));
    if (h->stage_mask) NVTB_CUDA_OK(cudaFreeAsync(h->stage_mask, st));
    h->stage_keys = nullptr; h->stage_mask = nullptr; h->stage_cap = 0;
    // sized for what the fit has shown so far (a small fit must not pay for 1 GiB), doubling
    const int64_t cap = cap_max;
    int64_t want = std::max<int64_t>((int64_t)1 << 22, next_pow2(2 * (h->rows_total + m)));
    want = std::max<int64_t>(want, (h->stage_hint + 63) / 64 * 64);      // one flush per fit from the second fit on
    want = std::max<int64_t>(std::min<int64_t>(want, cap), m);
    NVTB_CUDA_OK(cudaMallocAsync(&h->stage_keys, sizeof(int32_t) * (size_t)(want + 64), st));
    NVTB_CUDA_OK(cudaMallocAsync(&h->stage_mask, (size_t)(want / 8 + 64), st));
    h->stage_cap = want;
  }
  NVTB_CUDA_OK(cudaMemcpyAsync(h->stage_keys + h->stage_rows, kp, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToDevice, st));
  uint8_t* md = h->stage_mask + (h->stage_rows >> 3);
  const size_t mbytes = (size_t)((m + 7) >> 3);
  if (mp) NVTB_CUDA_OK(cudaMemcpyAsync(md, mp, mbytes, cudaMemcpyDeviceToDevice, st));
  else    NVTB_CUDA_OK(cudaMemsetAsync(md, 0xFF, mbytes, st));
  h->stage_rows += m;
  NVTB_CUDA_OK(cudaEventRecord(h->stage_ev, st));
  h->stage_last = st;
  return NVTB_OK;
}

template <typename KeyT>
static int launch_insert(nvtb_hashagg* h, const KeyT* kp, const uint8_t* mp, const AggCols& ac,
                         int64_t m, cudaStream_t st) {
  int rc = arena_acquire(h, &h->arena, m, h->n_agg);
  if (rc) return rc;
  arm_launch_kernel<<<1, 1, 0, st>>>(h->ctr);
  NVTB_LAUNCH_OK();
  if (h->n_agg == 0 && sizeof(KeyT) == 4 && h->t.narrow) {
    rc = launch_fold_i32(h, reinterpret_cast<const int32_t*>(kp), mp, m, st);
    if (rc) return rc;
  } else if (h->n_agg == 0) {
    const int grid = scan_grid(m, kInsertCtasPerSm);
    constexpr int kSmemBytes = SmemAgg<KeyT>::kBytes;
    if (h->t.narrow) {
      NVTB_CUDA_OK(cudaFuncSetAttribute(insert_keys_kernel<KeyT, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      insert_keys_kernel<KeyT, true><<<grid, kThreads, kSmemBytes, st>>>(kp, mp, m, h->t, h->ctr, h->arena);
    } else {
      NVTB_CUDA_OK(cudaFuncSetAttribute(insert_keys_kernel<KeyT, false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      insert_keys_kernel<KeyT, false><<<grid, kThreads, kSmemBytes, st>>>(kp, mp, m, h->t, h->ctr, h->arena);
    }
  } else {
    const int grid = plain_grid(m);
    insert_agg_kernel<KeyT><<<grid, kThreads, 0, st>>>(kp, mp, ac, m, h->t, h->ctr,
                                                       h->special_vals, h->arena);
  }
  NVTB_LAUNCH_OK();
  h->rows_total += m;
  return post(h, st);
}

// view of a handle for nvtb_vocab_build_from_hashagg (vocab.cu): synchronises on the handle's
// pending launch.  *pairs == nullptr when the handle is a hash table (the caller exports).
int hashagg_sorted_view(nvtb_hashagg* h, const uint64_t** pairs, int64_t* n_unique, int64_t* null_size,
                        uint64_t* max_count, int* is_i32_table, cudaStream_t st) {
  int64_t nu = 0, ns = 0;
  int rc = nvtb_hashagg_size(h, &nu, &ns, (void*)st);
  if (rc) return rc;
  *n_unique = nu;
  *null_size = ns;
  *max_count = (uint64_t)h->mailbox->max_count;
  *pairs = h->mode == 1 ? h->acc[h->acc_cur] : nullptr;
  *is_i32_table = (h->mode == 0 && h->t.narrow) ? 1 : 0;
  return NVTB_OK;
}

}  // namespace nvtb

using namespace nvtb;

extern "C" {

int nvtb_hashagg_create(nvtb_hashagg_t** out, int n_agg, int64_t capacity_hint) {
  NVTB_REQUIRE(out != nullptr, "out is NULL");
  ensure_pool_configured();
  NVTB_REQUIRE(n_agg >= 0 && n_agg <= kMaxAgg, "n_agg must be in [0, 8]");
  nvtb_hashagg* h = new (std::nothrow) nvtb_hashagg();
  NVTB_REQUIRE(h != nullptr, "host allocation failed");
  memset(h, 0, sizeof(*h));
  h->n_agg = n_agg;
  h->hint = capacity_hint > 0 ? capacity_hint : 0;
  cudaStream_t st = 0;
  int rc = table_alloc(&h->t, next_pow2(std::max<int64_t>((int64_t)(2.5 * capacity_hint), kMinCapacity)), n_agg, false, st);
  if (rc) { delete h; return rc; }
  NVTB_CUDA_OK(cudaMalloc(&h->ctr, sizeof(Counters)));
  NVTB_CUDA_OK(cudaMalloc(&h->special_vals, sizeof(double) * 8 * (n_agg > 0 ? n_agg : 1)));
  special_init_kernel<<<1, 32, 0, st>>>(h->ctr, h->special_vals, n_agg);
  NVTB_LAUNCH_OK();
  NVTB_CUDA_OK(cudaMallocHost(&h->mailbox, sizeof(Counters)));
  NVTB_CUDA_OK(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
  NVTB_CUDA_OK(cudaStreamSynchronize(st));
  *out = h;
  return NVTB_OK;
}

int nvtb_hashagg_reset(nvtb_hashagg_t* h, void* stream) {
  NVTB_REQUIRE(h != nullptr, "NULL handle");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = settle(h);
  if (rc) return rc;
  // the table keeps its capacity (and the estimate its value): a second fit over
  // similar data needs no growth and no sampling pass
  h->hint = std::max<int64_t>(h->hint, h->u_known);
  if (h->mode == 0) {      // a sorted accumulator is emptied by zeroing its counters below
    table_init_kernel<<<plain_grid(h->t.capacity), kThreads, 0, st>>>(h->t);
    NVTB_LAUNCH_OK();
  }
  special_init_kernel<<<1, 32, 0, st>>>(h->ctr, h->special_vals, h->n_agg);
  NVTB_LAUNCH_OK();
  h->u_known = 0;
  h->rows_total = 0;
  h->stage_hint = std::max<int64_t>(h->stage_hint, h->rows_total);
  h->stage_rows = 0;           // batches still waiting belong to the fit that is being discarded
  h->mailbox_valid = false;
  return NVTB_OK;
}

int nvtb_hashagg_destroy(nvtb_hashagg_t* h) {
  if (h == nullptr) return NVTB_OK;
  settle(h);              // hands a pooled arena back
  cudaDeviceSynchronize();
  if (h->t.slots) cudaFree(h->t.slots);
  if (h->t.vals) cudaFree(h->t.vals);
  if (h->ctr) cudaFree(h->ctr);
  if (h->special_vals) cudaFree(h->special_vals);
  if (h->mailbox) cudaFreeHost(h->mailbox);
  if (h->ev) cudaEventDestroy(h->ev);
  if (h->acc[0]) cudaFree(h->acc[0]);
  if (h->acc[1]) cudaFree(h->acc[1]);
  if (h->d_n) cudaFree(h->d_n);
  if (h->stage_keys) cudaFree(h->stage_keys);
  if (h->stage_mask) cudaFree(h->stage_mask);
  if (h->stage_ev) cudaEventDestroy(h->stage_ev);
  delete h;
  return NVTB_OK;
}

int nvtb_hashagg_insert(nvtb_hashagg_t* h, const nvtb_col_t* key,
                        const nvtb_col_t* agg_cols, int64_t n, void* stream) {
  NVTB_REQUIRE(h != nullptr && key != nullptr, "NULL handle/key");
  NVTB_REQUIRE(n >= 0, "n < 0");
  NVTB_REQUIRE(key->dtype == NVTB_I32 || key->dtype == NVTB_I64,
               "key dtype must be int32 or int64");
  NVTB_REQUIRE(h->n_agg == 0 || agg_cols != nullptr, "agg_cols is NULL");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(key->data != nullptr, "key data is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  AggCols ac;
  memset(&ac, 0, sizeof(ac));
  for (int j = 0; j < h->n_agg; ++j) {
    NVTB_REQUIRE(agg_cols[j].data != nullptr, "agg column data is NULL");
    NVTB_REQUIRE(agg_cols[j].dtype >= NVTB_I32 && agg_cols[j].dtype <= NVTB_U8, "bad agg dtype");
    ac.data[j] = agg_cols[j].data; ac.mask[j] = agg_cols[j].validity; ac.dtype[j] = agg_cols[j].dtype;
  }
  const size_t ksz = dtype_size(key->dtype);
  {
    int rc = settle(h);
    if (rc) return rc;
    const bool can_narrow = h->n_agg == 0 && key->dtype == NVTB_I32 &&
                            h->rows_total + n < (int64_t)0xFFFFFFF0ll;
    if (h->mode == 1) {
      // sorted accumulator: nothing to switch
    } else if (h->rows_total == 0 && h->u_known == 0 && can_narrow && !h->t.narrow) {
      // empty table: switch to the 8-byte layout in place
      const int64_t cap = h->t.capacity;
      rc = table_free(&h->t, st);
      if (rc) return rc;
      rc = table_alloc(&h->t, cap, 0, true, st);
      if (rc) return rc;
    } else if (h->t.narrow && !can_narrow) {
      rc = grow_to(h, h->t.capacity, st, /*force_wide=*/true);   // int64 keys or >= 2^32 rows
      if (rc) return rc;
    }
  }
  // batches: [sample of 2^20 rows when nothing is known about the cardinality] + the rest
  int64_t off = 0;
  while (off < n) {
    int rc = settle(h);
    if (rc) return rc;
    int64_t m = n - off;
    const bool blind = (h->mode == 0 && h->hint == 0 && h->k_est == 0);
    if (blind && m > 4 * kSampleRows) m = kSampleRows;
    const void* kp = (const char*)key->data + off * ksz;
    const uint8_t* mp = key->validity ? key->validity + (off >> 3) : nullptr;  // off % 8 == 0
    if (h->mode == 0) {
      // a column whose table would leave the L2 moves to the sorted accumulator (sortagg.cuh)
      predict(h, m);
      const bool eligible = h->n_agg == 0 && key->dtype == NVTB_I32 && h->t.narrow &&
                            h->rows_total + n < (int64_t)0xFFFFFFF0ll && m < (int64_t)0xFFFF0000ll;
      if (eligible && h->predicted > (double)runs_min_keys()) {
        rc = table_to_runs(h, st);
        if (rc) return rc;
      }
    }
    if (h->mode == 1) {
      NVTB_REQUIRE(h->n_agg == 0 && key->dtype == NVTB_I32, "a sorted accumulator takes int32 keys without payload");
      NVTB_REQUIRE(h->rows_total + m < (int64_t)0xFFFFFFF0ll, "more than 2^32 rows in one int32 accumulator");
      int64_t mm = m < (int64_t)0xFFFF0000ll ? m : (int64_t)0x80000000ll;
      const int64_t cap = stage_cap_rows();
      if (cap >= 64 && mm <= cap && (h->stage_rows & 7) == 0) {
        // stage: copy now, sort later together with the other batches of this fit
        if (h->stage_rows + mm > h->stage_cap && h->stage_rows > 0) {
          rc = stage_flush(h, st);
          if (rc) return rc;
        }
        rc = stage_append(h, (const int32_t*)kp, mp, mm, st);
        if (rc) return rc;
        h->rows_total += mm;
        off += mm;
        continue;
      }
      rc = stage_flush(h, st);            // ragged tail waiting, or a batch larger than the stage
      if (rc) return rc;
      rc = settle(h);
      if (rc) return rc;
      rc = launch_runs_insert(h, (const int32_t*)kp, mp, mm, st);
      if (rc) return rc;
      h->rows_total += mm;
      rc = post(h, st);
      if (rc) return rc;
      off += mm;
      continue;
    }
    rc = prepare(h, m, st);
    if (rc) return rc;
    AggCols a2 = ac;
    for (int j = 0; j < h->n_agg; ++j) {
      a2.data[j] = (const char*)ac.data[j] + off * dtype_size(ac.dtype[j]);
      a2.mask[j] = ac.mask[j] ? ac.mask[j] + (off >> 3) : nullptr;
    }
    if (key->dtype == NVTB_I32) rc = launch_insert<int32_t>(h, (const int32_t*)kp, mp, a2, m, st);
    else                        rc = launch_insert<int64_t>(h, (const int64_t*)kp, mp, a2, m, st);
    if (rc) return rc;
    off += m;
  }
  return NVTB_OK;
}

int nvtb_hashagg_merge(nvtb_hashagg_t* h, const int64_t* keys,
                       const int64_t* sizes, const double* vals, int64_t n,
                       void* stream) {
  NVTB_REQUIRE(h != nullptr && n >= 0, "NULL handle or n < 0");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(keys != nullptr && sizes != nullptr, "NULL keys/sizes");
  NVTB_REQUIRE(h->n_agg == 0 || vals != nullptr, "vals is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = settle(h);
  if (rc) return rc;
  if (h->mode == 1) {
    set_error("nvtb_hashagg_merge: the handle holds a sorted accumulator (raw int32 rows only)");
    return NVTB_ESTATE;
  }
  // pre-aggregated rows: distinct keys within the batch => every row may be new; sizes are
  // arbitrary int64 => wide layout
  rc = grow_to(h, next_pow2((int64_t)(2.5 * (double)(h->u_known + n)) + 1), st, /*force_wide=*/true);
  if (rc) return rc;
  rc = arena_acquire(h, &h->arena, n, h->n_agg);
  if (rc) return rc;
  rc = launch_merge(h, keys, sizes, vals, n, h->arena, st);
  if (rc) return rc;
  h->rows_total += n;
  return post(h, st);
}

int nvtb_hashagg_add_null_group(nvtb_hashagg_t* h, int64_t size, const double* vals_host) {
  NVTB_REQUIRE(h != nullptr && size >= 0, "NULL handle or size < 0");
  // tiny, synchronous: used once per rank in the cross-GPU merge
  int rc = settle(h);
  if (rc) return rc;
  Counters s;
  NVTB_CUDA_OK(cudaDeviceSynchronize());
  NVTB_CUDA_OK(cudaMemcpy(&s, h->ctr, sizeof(s), cudaMemcpyDeviceToHost));
  s.size[0] += (unsigned long long)size;
  NVTB_CUDA_OK(cudaMemcpy(h->ctr, &s, sizeof(s), cudaMemcpyHostToDevice));
  h->mailbox_valid = false;
  if (h->n_agg > 0 && vals_host != nullptr) {
    double cur[4 * kMaxAgg];
    NVTB_CUDA_OK(cudaMemcpy(cur, h->special_vals, sizeof(double) * 4 * h->n_agg, cudaMemcpyDeviceToHost));
    for (int j = 0; j < h->n_agg; ++j) {
      cur[j * 4 + 0] += vals_host[j * 4 + 0];
      cur[j * 4 + 1] += vals_host[j * 4 + 1];
      int64_t mn, mx;
      memcpy(&mn, &cur[j * 4 + 2], 8);
      memcpy(&mx, &cur[j * 4 + 3], 8);
      const double a = vals_host[j * 4 + 2], b = vals_host[j * 4 + 3];
      if (a == a) mn = std::min<int64_t>(mn, enc_ordered(a));
      if (b == b) mx = std::max<int64_t>(mx, enc_ordered(b));
      memcpy(&cur[j * 4 + 2], &mn, 8);
      memcpy(&cur[j * 4 + 3], &mx, 8);
    }
    NVTB_CUDA_OK(cudaMemcpy(h->special_vals, cur, sizeof(double) * 4 * h->n_agg, cudaMemcpyHostToDevice));
  }
  return NVTB_OK;
}

int nvtb_hashagg_size(nvtb_hashagg_t* h, int64_t* n_unique, int64_t* null_size, void* stream) {
  NVTB_REQUIRE(h != nullptr, "NULL handle");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = settle(h);
  if (rc) return rc;
  if (h->stage_rows > 0) {   // a sorted accumulator with batches still waiting: sort them in now
    rc = stage_flush(h, st);
    if (rc) return rc;
    rc = settle(h);
    if (rc) return rc;
  }
  if (!h->mailbox_valid) {   // e.g. right after create/reset/add_null_group: read the counters
    NVTB_CUDA_OK(cudaMemcpyAsync(h->mailbox, h->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    NVTB_CUDA_OK(cudaStreamSynchronize(st));
    h->mailbox_valid = true;
  }
  const Counters s = *h->mailbox;
  h->u_known = (int64_t)s.n_unique;
  if (n_unique) *n_unique = (int64_t)s.n_unique + (s.size[1] ? 1 : 0);
  if (null_size) *null_size = (int64_t)s.size[0];
  return NVTB_OK;
}

int nvtb_hashagg_export(nvtb_hashagg_t* h, int64_t* keys_out, int64_t* sizes_out,
                        double* vals_out, double* null_vals_host, void* stream) {
  NVTB_REQUIRE(h != nullptr, "NULL handle");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t nu = 0, ns = 0;
  int rc = nvtb_hashagg_size(h, &nu, &ns, stream);
  if (rc) return rc;
  const Counters s = *h->mailbox;
  double dec[8 * kMaxAgg];
  if (h->n_agg > 0) {
    double* d_dec = nullptr;
    NVTB_CUDA_OK(cudaMallocAsync(&d_dec, sizeof(double) * 8 * h->n_agg, st));
    decode_special_kernel<<<1, 64, 0, st>>>(h->special_vals, h->n_agg, d_dec);
    NVTB_LAUNCH_OK();
    NVTB_CUDA_OK(cudaMemcpyAsync(dec, d_dec, sizeof(double) * 8 * h->n_agg, cudaMemcpyDeviceToHost, st));
    NVTB_CUDA_OK(cudaFreeAsync(d_dec, st));
    NVTB_CUDA_OK(cudaStreamSynchronize(st));
    if (null_vals_host) memcpy(null_vals_host, dec, sizeof(double) * 4 * h->n_agg);
  }
  if (nu == 0) return NVTB_OK;
  NVTB_REQUIRE(keys_out != nullptr, "keys_out is NULL");
  NVTB_REQUIRE(h->n_agg == 0 || vals_out != nullptr, "vals_out is NULL");
  if (h->mode == 1) {      // sorted accumulator: unpack (the rows come out in key order)
    runs_unpack_kernel<<<plain_grid(nu), kThreads, 0, st>>>(h->acc[h->acc_cur], nu, keys_out, sizes_out);
    NVTB_LAUNCH_OK();
    return NVTB_OK;
  }
  unsigned long long* cursor = nullptr;
  NVTB_CUDA_OK(cudaMallocAsync(&cursor, sizeof(unsigned long long), st));
  NVTB_CUDA_OK(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st));
  export_kernel<<<plain_grid(h->t.capacity / kExportPerThread), kThreads, 0, st>>>(h->t, keys_out, sizes_out, vals_out, cursor);
  NVTB_LAUNCH_OK();
  NVTB_CUDA_OK(cudaFreeAsync(cursor, st));
  if (s.size[1]) {  // the INT64_MIN key lives outside the table: append it last
    const int64_t o = (int64_t)s.n_unique;
    const int64_t k = kEmptyKey, sz = (int64_t)s.size[1];
    NVTB_CUDA_OK(cudaMemcpyAsync(keys_out + o, &k, 8, cudaMemcpyHostToDevice, st));
    if (sizes_out) NVTB_CUDA_OK(cudaMemcpyAsync(sizes_out + o, &sz, 8, cudaMemcpyHostToDevice, st));
    if (vals_out && h->n_agg > 0)
      NVTB_CUDA_OK(cudaMemcpyAsync(vals_out + o * h->n_agg * 4, dec + 4 * h->n_agg,
                                   sizeof(double) * 4 * h->n_agg, cudaMemcpyHostToDevice, st));
    NVTB_CUDA_OK(cudaStreamSynchronize(st));  // host temporaries above
  }
  return NVTB_OK;
}

// Stable LSD radix sort of bits [lo_bit, hi_bit) (radix.cuh), exposed for tests and for
// callers that order their own device arrays.  The result ends in `data` or in `tmp`
// (*result_in_tmp_host).
static int radix_sort_entry(void* data, void* tmp, int64_t n, int elem_bytes, int lo_bit, int hi_bit,
                            int descending, int* result_in_tmp_host, void* stream) {
  NVTB_REQUIRE(n >= 0 && result_in_tmp_host != nullptr, "bad n / NULL result flag");
  NVTB_REQUIRE(lo_bit >= 0 && hi_bit <= 8 * elem_bytes && lo_bit <= hi_bit, "bad bit range");
  *result_in_tmp_host = 0;
  if (n == 0 || lo_bit == hi_bit) return NVTB_OK;
  NVTB_REQUIRE(data != nullptr && tmp != nullptr, "NULL data/tmp");
  NVTB_REQUIRE((reinterpret_cast<uintptr_t>(data) & 15u) == 0 && (reinterpret_cast<uintptr_t>(tmp) & 15u) == 0,
               "data/tmp must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  void* scratch = nullptr;
  const size_t bytes = elem_bytes == 4 ? rx_scratch_bytes<uint32_t>(n, kRxMaxStableBits - 1)
                                       : rx_scratch_bytes<uint64_t>(n, kRxMaxStableBits - 1);
  NVTB_CUDA_OK(cudaMallocAsync(&scratch, bytes, st));
  NVTB_CUDA_OK(cudaMemsetAsync(scratch, 0, 256, st));
  int rc;
  if (elem_bytes == 4)
    rc = rx_sort_bits<uint32_t>((uint32_t*)data, (uint32_t*)tmp, nullptr, n, lo_bit, hi_bit, descending != 0, scratch, st,
                                result_in_tmp_host);
  else
    rc = rx_sort_bits<uint64_t>((uint64_t*)data, (uint64_t*)tmp, nullptr, n, lo_bit, hi_bit, descending != 0, scratch, st,
                                result_in_tmp_host);
  NVTB_CUDA_OK(cudaFreeAsync(scratch, st));
  return rc;
}

int nvtb_radix_sort_u32(uint32_t* data, uint32_t* tmp, int64_t n, int lo_bit, int hi_bit, int descending,
                        int* result_in_tmp_host, void* stream) {
  return radix_sort_entry(data, tmp, n, 4, lo_bit, hi_bit, descending, result_in_tmp_host, stream);
}

int nvtb_radix_sort_u64(uint64_t* data, uint64_t* tmp, int64_t n, int lo_bit, int hi_bit, int descending,
                        int* result_in_tmp_host, void* stream) {
  return radix_sort_entry(data, tmp, n, 8, lo_bit, hi_bit, descending, result_in_tmp_host, stream);
}

// ---------------------------------------------------------------------------------------
// sorted-pair primitives of the cross-GPU vocabulary merge (nvtabular_b200/dist.py)
// ---------------------------------------------------------------------------------------
// Turn an int32 key-count handle into a sorted accumulator (no-op when it already is one), so
// that every rank of a fit holds the same representation of a high-cardinality column.
int nvtb_hashagg_to_sorted(nvtb_hashagg_t* h, void* stream) {
  NVTB_REQUIRE(h != nullptr, "NULL handle");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = settle(h);
  if (rc) return rc;
  if (h->mode == 1) return NVTB_OK;
  int64_t nu = 0, ns = 0;
  rc = nvtb_hashagg_size(h, &nu, &ns, stream);
  if (rc) return rc;
  if (h->n_agg != 0 || (h->u_known > 0 && !h->t.narrow) || h->mailbox->size[1] != 0 ||
      h->rows_total >= (int64_t)0xFFFFFFF0ll) {
    set_error("nvtb_hashagg_to_sorted: only int32 key-count tables below 2^32 rows can become sorted accumulators");
    return NVTB_ESTATE;
  }
  return table_to_runs(h, st);
}

// packed pairs (key ^ 2^31) << 32 | count of a sorted accumulator, in key order.  out == NULL:
// only *n_host is set.
int nvtb_hashagg_export_packed(nvtb_hashagg_t* h, uint64_t* out, int64_t* n_host, void* stream) {
  NVTB_REQUIRE(h != nullptr && n_host != nullptr, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t nu = 0, ns = 0;
  int rc = nvtb_hashagg_size(h, &nu, &ns, stream);
  if (rc) return rc;
  if (h->mode != 1) {
    set_error("nvtb_hashagg_export_packed: the handle is not a sorted accumulator");
    return NVTB_ESTATE;
  }
  *n_host = h->u_known;
  if (out != nullptr && h->u_known > 0)
    NVTB_CUDA_OK(cudaMemcpyAsync(out, h->acc[h->acc_cur], sizeof(uint64_t) * (size_t)h->u_known,
                                 cudaMemcpyDeviceToDevice, st));
  return NVTB_OK;
}

int nvtb_pairs_lower_bounds(const uint64_t* pairs, int64_t n, const uint32_t* bounds_dev, int m,
                            int64_t* out_dev, void* stream) {
  NVTB_REQUIRE(n >= 0 && m >= 0, "negative size");
  if (m == 0) return NVTB_OK;
  NVTB_REQUIRE(bounds_dev != nullptr && out_dev != nullptr && (n == 0 || pairs != nullptr), "NULL argument");
  pairs_lower_bound_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      pairs, n, bounds_dev, m, reinterpret_cast<long long*>(out_dev));
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

// out = merge of two key-sorted, key-unique packed-pair arrays, counts of equal keys added.
// `out` must hold na + nb pairs; the merged length comes back on the host (one stream sync).
int nvtb_pairs_merge(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, uint64_t* out,
                     int64_t* n_out_host, void* stream) {
  NVTB_REQUIRE(na >= 0 && nb >= 0 && n_out_host != nullptr, "bad sizes / NULL n_out");
  NVTB_REQUIRE(na + nb < (int64_t)0xFFFF0000ll, "more than 2^32 pairs in one merge");
  cudaStream_t st = (cudaStream_t)stream;
  *n_out_host = 0;
  if (na + nb == 0) return NVTB_OK;
  NVTB_REQUIRE(out != nullptr, "NULL out");
  if (na == 0 || nb == 0) {
    NVTB_CUDA_OK(cudaMemcpyAsync(out, na ? a : b, sizeof(uint64_t) * (size_t)(na + nb), cudaMemcpyDeviceToDevice, st));
    *n_out_host = na + nb;
    return NVTB_OK;
  }
  static bool attrs = false;
  if (!attrs) {
    NVTB_CUDA_OK(cudaFuncSetAttribute(merge_write_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeSmem)));
    attrs = true;
  }
  const int64_t mt = (na + nb + kMergeTile - 1) / kMergeTile;
  // scratch: splits[mt + 1] | tile_out[mt] | ub | n_unique | max_count
  char* scratch = nullptr;
  const size_t sp_bytes = align_up(sizeof(uint2) * (size_t)(mt + 2), 256);
  const size_t to_bytes = align_up(sizeof(uint32_t) * (size_t)(mt + 2), 256);
  NVTB_CUDA_OK(cudaMallocAsync(&scratch, sp_bytes + to_bytes + 256, st));
  uint2* splits = reinterpret_cast<uint2*>(scratch);
  uint32_t* tile_out = reinterpret_cast<uint32_t*>(scratch + sp_bytes);
  uint32_t* ub_dev = reinterpret_cast<uint32_t*>(scratch + sp_bytes + to_bytes);
  unsigned long long* nu_dev = reinterpret_cast<unsigned long long*>(scratch + sp_bytes + to_bytes + 8);
  unsigned long long* mx_dev = nu_dev + 1;
  const uint32_t ub_h = (uint32_t)nb;
  NVTB_CUDA_OK(cudaMemsetAsync(ub_dev, 0, 64, st));
  NVTB_CUDA_OK(cudaMemcpyAsync(ub_dev, &ub_h, sizeof(ub_h), cudaMemcpyHostToDevice, st));
  merge_split_kernel<<<(int)((mt + 1 + 255) / 256), 256, 0, st>>>(a, (uint32_t)na, b, ub_dev, (int)mt, splits);
  NVTB_LAUNCH_OK();
  merge_count_kernel<<<(int)mt, kRunThreads, 0, st>>>(a, b, splits, tile_out);
  NVTB_LAUNCH_OK();
  scan_tiles_kernel<<<1, kRunThreads, 0, st>>>(tile_out, (int)mt, nullptr, nu_dev);
  NVTB_LAUNCH_OK();
  merge_write_kernel<true><<<(int)mt, kRunThreads, sizeof(MergeSmem), st>>>(a, b, splits, tile_out, out, mx_dev);
  NVTB_LAUNCH_OK();
  unsigned long long nu_h = 0;
  NVTB_CUDA_OK(cudaMemcpyAsync(&nu_h, nu_dev, sizeof(nu_h), cudaMemcpyDeviceToHost, st));
  NVTB_CUDA_OK(cudaFreeAsync(scratch, st));
  NVTB_CUDA_OK(cudaStreamSynchronize(st));
  *n_out_host = (int64_t)nu_h;
  return NVTB_OK;
}

// contiguous segments of src to their destinations: seg_src[nseg + 1] ascending prefix (device),
// seg_dst[nseg] (device)
int nvtb_segment_copy_u64(const uint64_t* src, uint64_t* dst, const int64_t* seg_src_dev, const int64_t* seg_dst_dev,
                          int nseg, int64_t n, void* stream) {
  NVTB_REQUIRE(nseg >= 0 && n >= 0, "negative size");
  if (n == 0 || nseg == 0) return NVTB_OK;
  NVTB_REQUIRE(src && dst && seg_src_dev && seg_dst_dev, "NULL argument");
  segment_copy_kernel<<<plain_grid((n + 3) / 4), kThreads, 0, (cudaStream_t)stream>>>(
      src, dst, reinterpret_cast<const long long*>(seg_src_dev), reinterpret_cast<const long long*>(seg_dst_dev), nseg, n);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

// Sort / run-length encode / merge the batches a sorted accumulator has staged (no-op for a hash
// table or when nothing is waiting).  Reads of the handle do this implicitly; a caller that
// wants the cost attributed to the group-by (bench.py) calls it at the end of the last batch.
int nvtb_hashagg_flush(nvtb_hashagg_t* h, void* stream) {
  NVTB_REQUIRE(h != nullptr, "NULL handle");
  if (h->mode != 1 || h->stage_rows == 0) return NVTB_OK;
  return stage_flush(h, (cudaStream_t)stream);
}

int nvtb_hashagg_mode(nvtb_hashagg_t* h, int* mode_host) {
  NVTB_REQUIRE(h != nullptr && mode_host != nullptr, "NULL argument");
  *mode_host = h->mode;
  return NVTB_OK;
}

int nvtb_partition_by_owner(const int64_t* keys, int64_t n, int n_parts,
                            int64_t* perm_out, int64_t* part_counts_host, void* stream) {
  NVTB_REQUIRE(n >= 0 && n_parts >= 1 && n_parts <= 64, "n_parts must be in [1, 64]");
  NVTB_REQUIRE(part_counts_host != nullptr, "part_counts_host is NULL");
  for (int p = 0; p < n_parts; ++p) part_counts_host[p] = 0;
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(keys != nullptr && perm_out != nullptr, "NULL keys/perm");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* d = nullptr;
  NVTB_CUDA_OK(cudaMallocAsync(&d, sizeof(unsigned long long) * 128, st));
  NVTB_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * 128, st));
  const int grid = plain_grid(n);
  owner_count_kernel<<<grid, kThreads, 0, st>>>(keys, n, n_parts, d);
  NVTB_LAUNCH_OK();
  unsigned long long hc[64];
  NVTB_CUDA_OK(cudaMemcpyAsync(hc, d, sizeof(unsigned long long) * n_parts, cudaMemcpyDeviceToHost, st));
  NVTB_CUDA_OK(cudaStreamSynchronize(st));
  unsigned long long cur[64], acc = 0;
  for (int p = 0; p < n_parts; ++p) { cur[p] = acc; acc += hc[p]; part_counts_host[p] = (int64_t)hc[p]; }
  NVTB_CUDA_OK(cudaMemcpyAsync(d + 64, cur, sizeof(unsigned long long) * n_parts, cudaMemcpyHostToDevice, st));
  owner_scatter_kernel<<<grid, kThreads, 0, st>>>(keys, n, n_parts, d + 64, perm_out);
  NVTB_LAUNCH_OK();
  NVTB_CUDA_OK(cudaStreamSynchronize(st));  // `cur` is a host temporary
  NVTB_CUDA_OK(cudaFreeAsync(d, st));
  return NVTB_OK;
}

int nvtb_partition_by_owner_async(const int64_t* keys, int64_t n, int n_parts,
                                  int64_t* perm_out, int64_t* part_counts_dev, void* stream) {
  NVTB_REQUIRE(n >= 0 && n_parts >= 1 && n_parts <= 64, "n_parts must be in [1, 64]");
  NVTB_REQUIRE(part_counts_dev != nullptr, "part_counts_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  NVTB_CUDA_OK(cudaMemsetAsync(part_counts_dev, 0, sizeof(int64_t) * n_parts, st));
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(keys != nullptr && perm_out != nullptr, "NULL keys/perm");
  unsigned long long* d = nullptr;     // [64] write cursors
  NVTB_CUDA_OK(cudaMallocAsync(&d, sizeof(unsigned long long) * 64, st));
  const int grid = plain_grid(n);
  owner_count_kernel<<<grid, kThreads, 0, st>>>(keys, n, n_parts, reinterpret_cast<unsigned long long*>(part_counts_dev));
  NVTB_LAUNCH_OK();
  owner_prefix_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const unsigned long long*>(part_counts_dev), n_parts, d);
  NVTB_LAUNCH_OK();
  owner_scatter_kernel<<<grid, kThreads, 0, st>>>(keys, n, n_parts, d, perm_out);
  NVTB_LAUNCH_OK();
  NVTB_CUDA_OK(cudaFreeAsync(d, st));
  return NVTB_OK;
}

int nvtb_gather_i64(const int64_t* src, const int64_t* perm, int64_t n, int64_t* dst, void* stream) {
  NVTB_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(src && perm && dst, "NULL pointer");
  gather_i64_kernel<<<plain_grid(n), kThreads, 0, (cudaStream_t)stream>>>(src, perm, n, dst);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

int nvtb_gather_f64_rows(const double* src, const int64_t* perm, int64_t n, int row_width,
                         double* dst, void* stream) {
  NVTB_REQUIRE(n >= 0 && row_width >= 1, "bad n/row_width");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(src && perm && dst, "NULL pointer");
  gather_f64_rows_kernel<<<plain_grid(n * row_width), kThreads, 0, (cudaStream_t)stream>>>(src, perm, n, row_width, dst);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

int nvtb_pack_keys2(const nvtb_col_t* a, const nvtb_col_t* b, int64_t n,
                    int64_t* keys_out, uint8_t* validity_out, void* stream) {
  NVTB_REQUIRE(a && b && n >= 0, "NULL column or n < 0");
  NVTB_REQUIRE(a->dtype == NVTB_I32 && b->dtype == NVTB_I32, "pack_keys2 needs int32 columns");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(a->data && b->data && keys_out, "NULL data");
  pack_keys2_kernel<<<plain_grid((n + 7) / 8), kThreads, 0, (cudaStream_t)stream>>>(
      (const int32_t*)a->data, a->validity, (const int32_t*)b->data, b->validity, n, keys_out, validity_out);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

}  // extern "C"
