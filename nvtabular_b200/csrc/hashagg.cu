// hashagg.cu — K3: device hash aggregation  groupby(key, dropna=False).agg(...)
//
// Replaces the reference's per-partition cuDF groupby + concat/groupby tree
// (nvtabular/ops/categorify.py:955-1137, graph built at :1344-1540) with ONE
// resident open-addressing table per column group that every batch is folded
// into.  Design (B200-first, not a translation of cuDF's groupby):
//
//   * table slot = {int64 key, int64 size} (16 B, one 32-B sector holds two) in
//     HBM/L2; optional per-slot payload {sum, sumsq, min, max} per cont column.
//   * the insert kernel streams the key column with 256-bit loads and first
//     folds rows into a per-CTA shared-memory table (4096 slots, ATOMS), so
//     hot keys (Zipf heads, low-cardinality columns) cost one global atomic
//     per CTA instead of one per row; misses go straight to the global table
//     (RED.ADD after a key CAS).
//   * load factor is kept <= 0.5 BY CONSTRUCTION: before a batch of B rows is
//     launched the table has capacity >= 2*(U + B) where U is an upper bound on
//     the distinct keys so far (true count read back asynchronously through a
//     pinned mailbox, plus rows launched since).  Linear probing therefore
//     always terminates and no overflow path exists.  Growth = rehash kernel.
//   * the null key (dropna=False) and the one key equal to the EMPTY sentinel
//     (INT64_MIN) live in two "special" groups outside the table.
//
// Throughput bound: shared/L2 atomic units (~1 atomic/clk/SM), not HBM; see
// DESIGN.md "K3 roofline".
#include <algorithm>
#include <new>

#include "common.cuh"

namespace nvtb {

constexpr int kSmemSlots = 4096;   // per-CTA pre-aggregation table
constexpr int kSmemProbes = 4;
constexpr int kInsertSmemBytes = kSmemSlots * (int)(sizeof(long long) + sizeof(unsigned int));
constexpr int64_t kChunkRows = (int64_t)1 << 23;  // rows per launch
constexpr int64_t kMinCapacity = 1 << 10;

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

struct Table {
  int64_t* slots;   // [2*capacity] {key, size}
  double* vals;     // [capacity * 4 * n_agg] or nullptr
  int64_t capacity; // power of two
  int n_agg;
};

// counters (device, unsigned long long): 0 = distinct keys in table,
// special groups: [0] = null key, [1] = INT64_MIN key
struct Special {
  unsigned long long n_unique;
  unsigned long long size[2];
};

}  // namespace nvtb

struct nvtb_hashagg {
  nvtb::Table t;
  nvtb::Special* ctr;        // device
  double* special_vals;      // device [2][4*n_agg]
  nvtb::Special* mailbox;    // pinned host
  cudaEvent_t mailbox_ev;
  bool mailbox_pending;
  int64_t u_known;           // distinct keys at the last completed readback
  int64_t rows_since;        // rows launched after that readback was enqueued
  int64_t rows_at_enqueue;   // rows_since value that the pending readback covers
  int n_agg;
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

// find-or-claim the slot of `key` (key != kEmptyKey).  Load factor <= 0.5 is
// guaranteed by the host, so the loop terminates.
__device__ __forceinline__ int64_t table_find_or_insert(const Table& t,
                                                        int64_t key,
                                                        Special* ctr) {
  const int64_t mask = t.capacity - 1;
  int64_t slot = (int64_t)(table_mix64((uint64_t)key) & (uint64_t)mask);
  while (true) {
    long long* kp = reinterpret_cast<long long*>(t.slots + 2 * slot);
    long long cur = __ldcg(kp);
    if (cur == key) return slot;
    if (cur == kEmptyKey) {
      long long prev = (long long)atomicCAS(
          reinterpret_cast<unsigned long long*>(kp),
          (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (prev == kEmptyKey) {
        atomicAdd(&ctr->n_unique, 1ull);
        return slot;
      }
      if (prev == key) return slot;
    }
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ void table_add_size(const Table& t, int64_t slot,
                                               int64_t add) {
  atomicAdd(reinterpret_cast<unsigned long long*>(t.slots + 2 * slot + 1),
            (unsigned long long)add);
}

__global__ void table_init_kernel(Table t) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
       s < t.capacity; s += stride) {
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

__global__ void special_init_kernel(Special* ctr, double* special_vals, int n_agg) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ctr->n_unique = 0; ctr->size[0] = 0; ctr->size[1] = 0;
    for (int g = 0; g < 2; ++g)
      for (int j = 0; j < n_agg; ++j) {
        double* v = special_vals + (g * n_agg + j) * 4;
        v[0] = 0.0; v[1] = 0.0;
        reinterpret_cast<int64_t*>(v)[2] = kMinInit;
        reinterpret_cast<int64_t*>(v)[3] = kMaxInit;
      }
  }
}

// ---------------------------------------------------------------------------
// insert, keys only (Categorify): smem pre-aggregation + global table
// ---------------------------------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
insert_keys_kernel(const KeyT* __restrict__ keys,
                   const uint8_t* __restrict__ mask, int64_t n, Table t,
                   Special* ctr) {
  // 48 KB of dynamic shared memory: keys[4096] (8 B) then counts[4096] (4 B)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  long long* skeys = reinterpret_cast<long long*>(smem_raw);
  unsigned int* scnt = reinterpret_cast<unsigned int*>(smem_raw + sizeof(long long) * kSmemSlots);
  __shared__ unsigned long long s_null, s_min;
  for (int s = threadIdx.x; s < kSmemSlots; s += kThreads) {
    skeys[s] = kEmptyKey;
    scnt[s] = 0u;
  }
  if (threadIdx.x == 0) { s_null = 0ull; s_min = 0ull; }
  __syncthreads();

  unsigned int n_null = 0, n_min = 0;
  const bool aligned = is_aligned32(keys);
  for_each_row<KeyT>(keys, mask, n, aligned, [&](int64_t, KeyT x, bool valid) {
    if (!valid) { n_null++; return; }
    const long long k = (long long)x;
    if (sizeof(KeyT) == 8 && k == kEmptyKey) { n_min++; return; }
    const uint64_t h = table_mix64((uint64_t)k);
    // upper hash bits pick the smem slot so that it is independent of the
    // global slot (low bits)
    unsigned s = (unsigned)(h >> 40) & (kSmemSlots - 1);
#pragma unroll
    for (int p = 0; p < kSmemProbes; ++p) {
      long long cur = *reinterpret_cast<volatile long long*>(&skeys[s]);
      if (cur == kEmptyKey) {
        cur = (long long)atomicCAS(
            reinterpret_cast<unsigned long long*>(&skeys[s]),
            (unsigned long long)kEmptyKey, (unsigned long long)k);
        if (cur == kEmptyKey) cur = k;
      }
      if (cur == k) {
        atomicAdd(&scnt[s], 1u);
        return;
      }
      s = (s + 1) & (kSmemSlots - 1);
    }
    // smem neighbourhood full: straight to the global table
    const int64_t slot = table_find_or_insert(t, k, ctr);
    table_add_size(t, slot, 1);
  });

  if (n_null) atomicAdd(&s_null, (unsigned long long)n_null);
  if (n_min) atomicAdd(&s_min, (unsigned long long)n_min);
  __syncthreads();
  // flush the CTA-local aggregates: one global update per distinct key per CTA
  for (int s = threadIdx.x; s < kSmemSlots; s += kThreads) {
    const long long k = skeys[s];
    if (k != kEmptyKey) {
      const int64_t slot = table_find_or_insert(t, k, ctr);
      table_add_size(t, slot, (int64_t)scnt[s]);
    }
  }
  if (threadIdx.x == 0) {
    if (s_null) atomicAdd(&ctr->size[0], s_null);
    if (s_min) atomicAdd(&ctr->size[1], s_min);
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
                  Table t, Special* ctr, double* special_vals) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const bool valid = valid1(mask, i);
    const long long k = valid ? (long long)keys[i] : 0;
    double* vdst;
    if (!valid) {
      atomicAdd(&ctr->size[0], 1ull);
      vdst = special_vals;
    } else if (sizeof(KeyT) == 8 && k == kEmptyKey) {
      atomicAdd(&ctr->size[1], 1ull);
      vdst = special_vals + (int64_t)t.n_agg * 4;
    } else {
      const int64_t slot = table_find_or_insert(t, k, ctr);
      table_add_size(t, slot, 1);
      vdst = t.vals + slot * t.n_agg * 4;
    }
    for (int j = 0; j < t.n_agg; ++j) {
      double v;
      if (load_agg(agg, j, i, &v)) vals_combine(vdst + j * 4, v, v * v, v, v);
    }
  }
}

// merge pre-aggregated rows (other GPUs' partials, or an old table on growth)
__global__ void __launch_bounds__(kThreads)
merge_kernel(const int64_t* __restrict__ keys, const int64_t* __restrict__ sizes,
             const double* __restrict__ vals, int64_t n, Table t, Special* ctr,
             double* special_vals) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const long long k = keys[i];
    double* vdst;
    if (k == kEmptyKey) {
      atomicAdd(&ctr->size[1], (unsigned long long)sizes[i]);
      vdst = special_vals + (int64_t)t.n_agg * 4;
    } else {
      const int64_t slot = table_find_or_insert(t, k, ctr);
      table_add_size(t, slot, sizes[i]);
      vdst = t.vals + slot * t.n_agg * 4;
    }
    if (vals != nullptr)
      for (int j = 0; j < t.n_agg; ++j) {
        const double* v = vals + (i * t.n_agg + j) * 4;
        vals_combine(vdst + j * 4, v[0], v[1], v[2], v[3]);
      }
  }
}

// rehash an old table into a new one (keys are distinct: plain stores after
// the claim; n_unique is carried over by the host)
__global__ void __launch_bounds__(kThreads)
rehash_kernel(Table old_t, Table new_t, Special* scratch_ctr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
       s < old_t.capacity; s += stride) {
    const long long k = old_t.slots[2 * s];
    if (k == kEmptyKey) continue;
    const int64_t slot = table_find_or_insert(new_t, k, scratch_ctr);
    new_t.slots[2 * slot + 1] = old_t.slots[2 * s + 1];
    for (int j = 0; j < old_t.n_agg * 4; ++j)
      new_t.vals[slot * old_t.n_agg * 4 + j] = old_t.vals[s * old_t.n_agg * 4 + j];
  }
}

// compaction: table -> dense (unordered) arrays
__global__ void __launch_bounds__(kThreads)
export_kernel(Table t, int64_t* __restrict__ keys_out,
              int64_t* __restrict__ sizes_out, double* __restrict__ vals_out,
              unsigned long long* cursor) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // capacity is a power of two >= 1024, so every warp runs the same trip count
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
       s < t.capacity; s += stride) {
    const long long k = t.slots[2 * s];
    const bool live = (k != kEmptyKey);
    const unsigned ballot = __ballot_sync(0xffffffffu, live);
    if (ballot == 0) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) {
      const int64_t o = (int64_t)base + __popc(ballot & ((1u << lane) - 1u));
      keys_out[o] = k;
      if (sizes_out) sizes_out[o] = t.slots[2 * s + 1];
      if (vals_out)
        for (int j = 0; j < t.n_agg; ++j) {
          const double* v = t.vals + (s * t.n_agg + j) * 4;
          double* w = vals_out + (o * t.n_agg + j) * 4;
          w[0] = v[0]; w[1] = v[1];
          w[2] = dec_ordered(reinterpret_cast<const int64_t*>(v)[2]);
          w[3] = dec_ordered(reinterpret_cast<const int64_t*>(v)[3]);
        }
    }
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

__global__ void __launch_bounds__(kThreads)
owner_scatter_kernel(const int64_t* __restrict__ keys, int64_t n, int n_parts,
                     unsigned long long* cursors, int64_t* __restrict__ perm) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int o = owner_of(keys[i], n_parts);
    const unsigned long long pos = atomicAdd(&cursors[o], 1ull);
    perm[pos] = i;
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

static int table_alloc(Table* t, int64_t capacity, int n_agg, cudaStream_t st) {
  t->capacity = capacity;
  t->n_agg = n_agg;
  t->slots = nullptr;
  t->vals = nullptr;
  NVTB_CUDA_OK(cudaMallocAsync(&t->slots, sizeof(int64_t) * 2 * capacity, st));
  if (n_agg > 0)
    NVTB_CUDA_OK(cudaMallocAsync(&t->vals, sizeof(double) * 4 * n_agg * capacity, st));
  int grid = (int)std::min<int64_t>((capacity + kThreads - 1) / kThreads,
                                    (int64_t)sm_count() * 8);
  table_init_kernel<<<grid, kThreads, 0, st>>>(*t);
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

// poll / wait the asynchronous distinct-count readback
static int mailbox_poll(nvtb_hashagg* h, bool block) {
  if (!h->mailbox_pending) return NVTB_OK;
  cudaError_t e = block ? cudaEventSynchronize(h->mailbox_ev)
                        : cudaEventQuery(h->mailbox_ev);
  if (e == cudaErrorNotReady) return NVTB_OK;
  if (e != cudaSuccess) {
    set_error("mailbox event failed: %s", cudaGetErrorString(e));
    return NVTB_ECUDA;
  }
  h->u_known = (int64_t)h->mailbox->n_unique;
  h->rows_since -= h->rows_at_enqueue;
  h->mailbox_pending = false;
  return NVTB_OK;
}

static int mailbox_post(nvtb_hashagg* h, cudaStream_t st) {
  if (h->mailbox_pending) return NVTB_OK;  // one readback in flight at a time
  NVTB_CUDA_OK(cudaMemcpyAsync(h->mailbox, h->ctr, sizeof(Special),
                               cudaMemcpyDeviceToHost, st));
  NVTB_CUDA_OK(cudaEventRecord(h->mailbox_ev, st));
  h->mailbox_pending = true;
  h->rows_at_enqueue = h->rows_since;
  return NVTB_OK;
}

// make sure the table can take `batch` more rows with load factor <= 0.5
static int ensure_capacity(nvtb_hashagg* h, int64_t batch, cudaStream_t st) {
  int rc = mailbox_poll(h, false);
  if (rc) return rc;
  int64_t need = 2 * (h->u_known + h->rows_since + batch);
  if (need <= h->t.capacity) return NVTB_OK;
  // the bound is stale: get the true distinct count before paying for growth
  if (!h->mailbox_pending) { rc = mailbox_post(h, st); if (rc) return rc; }
  rc = mailbox_poll(h, true);
  if (rc) return rc;
  if (h->rows_since > 0) {  // rows were launched after that readback: redo it
    rc = mailbox_post(h, st); if (rc) return rc;
    rc = mailbox_poll(h, true); if (rc) return rc;
  }
  need = 2 * (h->u_known + h->rows_since + batch);
  if (need <= h->t.capacity) return NVTB_OK;
  // grow: 4x headroom over what is known to be needed
  const int64_t new_cap = next_pow2(std::max<int64_t>(need, 4 * (h->u_known + batch)));
  Table nt;
  rc = table_alloc(&nt, new_cap, h->n_agg, st);
  if (rc) return rc;
  Special* scratch = nullptr;
  NVTB_CUDA_OK(cudaMallocAsync(&scratch, sizeof(Special), st));
  NVTB_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(Special), st));
  int grid = (int)std::min<int64_t>((h->t.capacity + kThreads - 1) / kThreads,
                                    (int64_t)sm_count() * 8);
  rehash_kernel<<<grid, kThreads, 0, st>>>(h->t, nt, scratch);
  NVTB_LAUNCH_OK();
  NVTB_CUDA_OK(cudaFreeAsync(scratch, st));
  rc = table_free(&h->t, st);
  if (rc) return rc;
  h->t = nt;
  return NVTB_OK;
}

}  // namespace nvtb

using namespace nvtb;

extern "C" {

int nvtb_hashagg_create(nvtb_hashagg_t** out, int n_agg, int64_t capacity_hint) {
  NVTB_REQUIRE(out != nullptr, "out is NULL");
  NVTB_REQUIRE(n_agg >= 0 && n_agg <= kMaxAgg, "n_agg must be in [0, 8]");
  nvtb_hashagg* h = new (std::nothrow) nvtb_hashagg();
  NVTB_REQUIRE(h != nullptr, "host allocation failed");
  memset(h, 0, sizeof(*h));
  h->n_agg = n_agg;
  cudaStream_t st = 0;
  int rc = table_alloc(&h->t, next_pow2(std::max<int64_t>(2 * capacity_hint, kMinCapacity)), n_agg, st);
  if (rc) { delete h; return rc; }
  NVTB_CUDA_OK(cudaMalloc(&h->ctr, sizeof(Special)));
  NVTB_CUDA_OK(cudaMalloc(&h->special_vals, sizeof(double) * 8 * (n_agg > 0 ? n_agg : 1)));
  special_init_kernel<<<1, 32, 0, st>>>(h->ctr, h->special_vals, n_agg);
  NVTB_LAUNCH_OK();
  NVTB_CUDA_OK(cudaMallocHost(&h->mailbox, sizeof(Special)));
  NVTB_CUDA_OK(cudaEventCreateWithFlags(&h->mailbox_ev, cudaEventDisableTiming));
  NVTB_CUDA_OK(cudaStreamSynchronize(st));
  *out = h;
  return NVTB_OK;
}

int nvtb_hashagg_destroy(nvtb_hashagg_t* h) {
  if (h == nullptr) return NVTB_OK;
  cudaDeviceSynchronize();
  if (h->t.slots) cudaFree(h->t.slots);
  if (h->t.vals) cudaFree(h->t.vals);
  if (h->ctr) cudaFree(h->ctr);
  if (h->special_vals) cudaFree(h->special_vals);
  if (h->mailbox) cudaFreeHost(h->mailbox);
  if (h->mailbox_ev) cudaEventDestroy(h->mailbox_ev);
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
  for (int64_t off = 0; off < n; off += kChunkRows) {
    const int64_t m = std::min<int64_t>(kChunkRows, n - off);
    int rc = ensure_capacity(h, m, st);
    if (rc) return rc;
    const void* kp = (const char*)key->data + off * ksz;
    const uint8_t* mp = key->validity ? key->validity + (off >> 3) : nullptr;  // off % 8 == 0
    if (h->n_agg == 0) {
      const int grid = scan_grid(m, 4);
      if (key->dtype == NVTB_I32) {
        NVTB_CUDA_OK(cudaFuncSetAttribute(insert_keys_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kInsertSmemBytes));
        insert_keys_kernel<int32_t><<<grid, kThreads, kInsertSmemBytes, st>>>((const int32_t*)kp, mp, m, h->t, h->ctr);
      } else {
        NVTB_CUDA_OK(cudaFuncSetAttribute(insert_keys_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kInsertSmemBytes));
        insert_keys_kernel<int64_t><<<grid, kThreads, kInsertSmemBytes, st>>>((const int64_t*)kp, mp, m, h->t, h->ctr);
      }
    } else {
      AggCols a2 = ac;
      for (int j = 0; j < h->n_agg; ++j) {
        a2.data[j] = (const char*)ac.data[j] + off * dtype_size(ac.dtype[j]);
        a2.mask[j] = ac.mask[j] ? ac.mask[j] + (off >> 3) : nullptr;
      }
      const int grid = (int)std::min<int64_t>((m + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
      if (key->dtype == NVTB_I32)
        insert_agg_kernel<int32_t><<<grid, kThreads, 0, st>>>((const int32_t*)kp, mp, a2, m, h->t, h->ctr, h->special_vals);
      else
        insert_agg_kernel<int64_t><<<grid, kThreads, 0, st>>>((const int64_t*)kp, mp, a2, m, h->t, h->ctr, h->special_vals);
    }
    NVTB_LAUNCH_OK();
    h->rows_since += m;
    rc = mailbox_post(h, st);
    if (rc) return rc;
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
  for (int64_t off = 0; off < n; off += kChunkRows) {
    const int64_t m = std::min<int64_t>(kChunkRows, n - off);
    int rc = ensure_capacity(h, m, st);
    if (rc) return rc;
    const int grid = (int)std::min<int64_t>((m + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
    merge_kernel<<<grid, kThreads, 0, st>>>(keys + off, sizes + off,
                                           vals ? vals + off * h->n_agg * 4 : nullptr,
                                           m, h->t, h->ctr, h->special_vals);
    NVTB_LAUNCH_OK();
    h->rows_since += m;
    rc = mailbox_post(h, st);
    if (rc) return rc;
  }
  return NVTB_OK;
}

int nvtb_hashagg_add_null_group(nvtb_hashagg_t* h, int64_t size, const double* vals_host) {
  NVTB_REQUIRE(h != nullptr && size >= 0, "NULL handle or size < 0");
  // tiny, synchronous: used once per rank in the cross-GPU merge
  Special s;
  NVTB_CUDA_OK(cudaDeviceSynchronize());
  NVTB_CUDA_OK(cudaMemcpy(&s, h->ctr, sizeof(s), cudaMemcpyDeviceToHost));
  s.size[0] += (unsigned long long)size;
  NVTB_CUDA_OK(cudaMemcpy(h->ctr, &s, sizeof(s), cudaMemcpyHostToDevice));
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
  int rc = mailbox_poll(h, true);
  if (rc) return rc;
  Special s;
  NVTB_CUDA_OK(cudaMemcpyAsync(h->mailbox, h->ctr, sizeof(Special), cudaMemcpyDeviceToHost, st));
  NVTB_CUDA_OK(cudaStreamSynchronize(st));
  s = *h->mailbox;
  h->u_known = (int64_t)s.n_unique;
  h->rows_since = 0;
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
  const Special s = *h->mailbox;
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
  unsigned long long* cursor = nullptr;
  NVTB_CUDA_OK(cudaMallocAsync(&cursor, sizeof(unsigned long long), st));
  NVTB_CUDA_OK(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st));
  int grid = (int)std::min<int64_t>((h->t.capacity + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
  export_kernel<<<grid, kThreads, 0, st>>>(h->t, keys_out, sizes_out, vals_out, cursor);
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
  const int grid = (int)std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
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

int nvtb_gather_i64(const int64_t* src, const int64_t* perm, int64_t n, int64_t* dst, void* stream) {
  NVTB_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(src && perm && dst, "NULL pointer");
  const int grid = (int)std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
  gather_i64_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(src, perm, n, dst);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

int nvtb_gather_f64_rows(const double* src, const int64_t* perm, int64_t n, int row_width,
                         double* dst, void* stream) {
  NVTB_REQUIRE(n >= 0 && row_width >= 1, "bad n/row_width");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(src && perm && dst, "NULL pointer");
  const int64_t total = n * row_width;
  const int grid = (int)std::min<int64_t>((total + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
  gather_f64_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(src, perm, n, row_width, dst);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

int nvtb_pack_keys2(const nvtb_col_t* a, const nvtb_col_t* b, int64_t n,
                    int64_t* keys_out, uint8_t* validity_out, void* stream) {
  NVTB_REQUIRE(a && b && n >= 0, "NULL column or n < 0");
  NVTB_REQUIRE(a->dtype == NVTB_I32 && b->dtype == NVTB_I32, "pack_keys2 needs int32 columns");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(a->data && b->data && keys_out, "NULL data");
  const int64_t n8 = (n + 7) / 8;
  const int grid = (int)std::min<int64_t>((n8 + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
  pack_keys2_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      (const int32_t*)a->data, a->validity, (const int32_t*)b->data, b->validity, n, keys_out, validity_out);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

}  // extern "C"
