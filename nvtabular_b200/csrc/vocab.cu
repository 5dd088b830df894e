// vocab.cu — K4 vocabulary ordering/cut, K5 lookup build + encode, and the
// group-statistics gather used by JoinGroupby / TargetEncoding.
//
// Reference behaviour restated (the reference does this with cuDF sort_values,
// merge and a second sort back to row order):
//   _write_uniques / _save_encodings  nvtabular/ops/categorify.py:1149-1337, 719-822
//   _encode                           nvtabular/ops/categorify.py:1558-1807
//   JoinGroupby.transform             nvtabular/ops/join_groupby.py:175-217
//
// Ordering rule: (size desc, key asc) — the stable form of the reference's
// sort_values(key) followed by sort_values(size, ascending=False)
// (categorify.py:1300,1316).  Implemented as two stable LSD
// radix sorts (cub::DeviceRadixSort — library code, on U distinct keys, not on
// the N-row stream).  The encode is a single in-order probe pass: no join, no
// sort back to row order.
#include <algorithm>
#include <cub/cub.cuh>
#include <new>

#include "common.cuh"

namespace nvtb {

// read-only lookup table: slot = {key, position}; immutable after build so the
// probes go through the read-only (L1-cacheable) path.
struct Lookup {
  int64_t* slots;     // [2*capacity]
  int64_t capacity;   // power of two, >= 2 * n
  int64_t min_key_pos;  // position of key INT64_MIN (the EMPTY sentinel) or -1
};

__global__ void lookup_init_kernel(int64_t* slots, int64_t capacity) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < capacity; s += stride) {
    slots[2 * s] = kEmptyKey;
    slots[2 * s + 1] = INT64_MAX;
  }
}

// position = index in `keys` (smallest index for a repeated key).  The INT64_MIN
// key cannot be stored (sentinel): its position goes to *min_key_pos.
__global__ void lookup_build_kernel(const int64_t* __restrict__ keys, int64_t n,
                                    int64_t* slots, int64_t capacity,
                                    long long* min_key_pos) {
  const int64_t mask = capacity - 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long k = keys[i];
    if (k == kEmptyKey) { *min_key_pos = i; continue; }
    int64_t slot = (int64_t)(table_mix64((uint64_t)k) & (uint64_t)mask);
    while (true) {
      long long prev = (long long)atomicCAS(
          reinterpret_cast<unsigned long long*>(slots + 2 * slot),
          (unsigned long long)kEmptyKey, (unsigned long long)k);
      if (prev == kEmptyKey || prev == k) {
        // duplicate keys (possible in a user vocab): the first position wins
        atomicMin(reinterpret_cast<long long*>(slots + 2 * slot + 1), (long long)i);
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}

__device__ __forceinline__ int64_t lookup_find(const Lookup& t, int64_t key) {
  if (key == kEmptyKey) return t.min_key_pos;
  const int64_t mask = t.capacity - 1;
  int64_t slot = (int64_t)(table_mix64((uint64_t)key) & (uint64_t)mask);
  while (true) {
    const longlong2 kv = __ldg(reinterpret_cast<const longlong2*>(t.slots + 2 * slot));
    if (kv.x == key) return kv.y;
    if (kv.x == kEmptyKey) return -1;
    slot = (slot + 1) & mask;
  }
}

// number of leading rows with size >= threshold in a size-descending array
__global__ void count_ge_kernel(const int64_t* __restrict__ sizes, int64_t n,
                                int64_t threshold, long long* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (sizes[i] >= threshold && (i == n - 1 || sizes[i + 1] < threshold)) *out = i + 1;
  }
}

// ---------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------
constexpr int kMaxHashCols = 8;
struct HashCols {
  const void* data[kMaxHashCols];
  const uint8_t* mask[kMaxHashCols];
  int32_t dtype[kMaxHashCols];
  int32_t ncols;
};

__device__ __forceinline__ uint64_t hash_cols_at(const HashCols& hc, int64_t i) {
  uint64_t h = 0;
  for (int c = 0; c < hc.ncols; ++c) {
    uint64_t bits;
    if (!valid1(hc.mask[c], i)) {
      bits = kNaNBits;
    } else {
      switch (hc.dtype[c]) {
        case NVTB_I32: bits = value_bits<int32_t>(((const int32_t*)hc.data[c])[i]); break;
        case NVTB_I64: bits = value_bits<int64_t>(((const int64_t*)hc.data[c])[i]); break;
        case NVTB_F32: bits = value_bits<float>(((const float*)hc.data[c])[i]); break;
        case NVTB_F64: bits = value_bits<double>(((const double*)hc.data[c])[i]); break;
        case NVTB_H64: h ^= (uint64_t)((const int64_t*)hc.data[c])[i]; continue;  // already a hash
        default:       bits = value_bits<uint8_t>(((const uint8_t*)hc.data[c])[i]); break;
      }
    }
    h ^= pandas_mix64(bits);
  }
  return h;
}

struct EncodeParams {
  int64_t null_label, oov_label, first_label;
  uint64_t num_buckets;  // <= 1: single OOV index
};

template <typename KeyT, typename OutT>
__global__ void __launch_bounds__(kThreads)
encode_kernel(const KeyT* __restrict__ keys, const uint8_t* __restrict__ mask,
              int64_t n, Lookup t, EncodeParams p, HashCols hc,
              OutT* __restrict__ out) {
  const bool aligned = is_aligned32(keys) && is_aligned32(out);
  map_rows<KeyT, OutT>(keys, mask, out, n, aligned,
                       [&](int64_t i, KeyT x, bool valid) -> OutT {
                         if (!valid) return (OutT)p.null_label;
                         const int64_t pos = lookup_find(t, (int64_t)x);
                         if (pos >= 0) return (OutT)(p.first_label + pos);
                         int64_t lab = p.oov_label;
                         if (p.num_buckets > 1) {
                           const uint64_t h = hc.ncols > 0
                                                  ? hash_cols_at(hc, i)
                                                  : pandas_mix64(value_bits<KeyT>(x));
                           lab += (int64_t)(h % p.num_buckets);
                         }
                         return (OutT)lab;
                       });
}

// ---------------------------------------------------------------------------
// group-statistics gather
// ---------------------------------------------------------------------------
constexpr int kMaxGatherCols = 16;
struct GatherOut {
  void* out[kMaxGatherCols];
  double miss[kMaxGatherCols];
  int32_t col[kMaxGatherCols];
  int32_t dtype[kMaxGatherCols];
  int32_t ncols;
};

template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
gather_stats_kernel(const KeyT* __restrict__ keys, const uint8_t* __restrict__ mask,
                    int64_t n, Lookup t, int64_t null_row,
                    const double* __restrict__ stats, int width, GatherOut go) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t row = valid1(mask, i) ? lookup_find(t, (int64_t)keys[i]) : null_row;
    for (int j = 0; j < go.ncols; ++j) {
      const double v = row >= 0 ? __ldg(stats + row * width + go.col[j]) : go.miss[j];
      switch (go.dtype[j]) {
        case NVTB_I32: ((int32_t*)go.out[j])[i] = (int32_t)v; break;
        case NVTB_I64: ((int64_t*)go.out[j])[i] = (int64_t)v; break;
        case NVTB_F32: ((float*)go.out[j])[i] = (float)v; break;
        default:       ((double*)go.out[j])[i] = v; break;
      }
    }
  }
}

static int64_t pow2_at_least(int64_t v) {
  int64_t p = 16;
  while (p < v) p <<= 1;
  return p;
}

static int lookup_create(Lookup* t, const int64_t* keys, int64_t n, bool may_have_dups,
                         cudaStream_t st) {
  t->capacity = pow2_at_least(2 * n);
  t->min_key_pos = -1;
  t->slots = nullptr;
  NVTB_CUDA_OK(cudaMallocAsync(&t->slots, sizeof(int64_t) * 2 * t->capacity, st));
  const int g0 = (int)std::min<int64_t>((t->capacity + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
  lookup_init_kernel<<<g0, kThreads, 0, st>>>(t->slots, t->capacity);
  NVTB_LAUNCH_OK();
  if (n > 0) {
    long long* d_min = nullptr;
    NVTB_CUDA_OK(cudaMallocAsync(&d_min, sizeof(long long), st));
    NVTB_CUDA_OK(cudaMemsetAsync(d_min, 0xFF, sizeof(long long), st));  // -1
    const int g1 = (int)std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
    lookup_build_kernel<<<g1, kThreads, 0, st>>>(keys, n, t->slots, t->capacity, d_min);
    NVTB_LAUNCH_OK();
    (void)may_have_dups;
    long long h_min = -1;
    NVTB_CUDA_OK(cudaMemcpyAsync(&h_min, d_min, sizeof(long long), cudaMemcpyDeviceToHost, st));
    NVTB_CUDA_OK(cudaStreamSynchronize(st));
    NVTB_CUDA_OK(cudaFreeAsync(d_min, st));
    t->min_key_pos = h_min;
  }
  return NVTB_OK;
}

}  // namespace nvtb

struct nvtb_vocab {
  nvtb::Lookup t;
  int64_t* keys;   // device [n_kept], label order
  int64_t* sizes;  // device [n_kept] or nullptr
  nvtb_vocab_info_t info;
};

struct nvtb_groupstats {
  nvtb::Lookup t;
  double* stats;  // device [n_groups * width]
  int64_t n_groups;
  int width;
  int64_t null_row;
};

using namespace nvtb;

extern "C" {

int nvtb_vocab_build(nvtb_vocab_t** out, const int64_t* keys, const int64_t* sizes,
                     int64_t n, int64_t null_size, int64_t freq_threshold,
                     int64_t max_size, int64_t num_buckets, void* stream) {
  NVTB_REQUIRE(out != nullptr && n >= 0, "out NULL or n < 0");
  NVTB_REQUIRE(n == 0 || (keys && sizes), "NULL keys/sizes");
  NVTB_REQUIRE(!(freq_threshold > 0 && max_size > 0),
               "cannot use freq_threshold together with max_size");
  const int64_t oov_count = num_buckets > 0 ? num_buckets : 1;
  // categorify.py:1206-1211
  NVTB_REQUIRE(!(max_size > 0 && max_size < oov_count + 2),
               "`max_size` can never be less than the maximum of `num_buckets + 2` and `3`");
  cudaStream_t st = (cudaStream_t)stream;
  nvtb_vocab* v = new (std::nothrow) nvtb_vocab();
  NVTB_REQUIRE(v != nullptr, "host allocation failed");
  memset(v, 0, sizeof(*v));
  v->info.n_total = n;
  v->info.null_size = null_size;
  int64_t n_keep = n;
  if (n > 0) {
    // (1) key asc, (2) stable size desc  =>  (size desc, key asc)
    int64_t *k1 = nullptr, *s1 = nullptr, *k2 = nullptr, *s2 = nullptr;
    NVTB_CUDA_OK(cudaMallocAsync(&k1, sizeof(int64_t) * n, st));
    NVTB_CUDA_OK(cudaMallocAsync(&s1, sizeof(int64_t) * n, st));
    NVTB_CUDA_OK(cudaMallocAsync(&k2, sizeof(int64_t) * n, st));
    NVTB_CUDA_OK(cudaMallocAsync(&s2, sizeof(int64_t) * n, st));
    size_t tmp_a = 0, tmp_b = 0;
    NVTB_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_a, keys, k1, sizes, s1, n, 0, 64, st));
    NVTB_CUDA_OK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_b, s1, s2, k1, k2, n, 0, 64, st));
    size_t tmp_bytes = std::max(tmp_a, tmp_b);
    void* tmp = nullptr;
    NVTB_CUDA_OK(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, st));
    NVTB_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, k1, sizes, s1, n, 0, 64, st));
    NVTB_CUDA_OK(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, s1, s2, k1, k2, n, 0, 64, st));
    NVTB_CUDA_OK(cudaFreeAsync(k1, st));
    NVTB_CUDA_OK(cudaFreeAsync(s1, st));

    // cut (categorify.py:766-785)
    long long* d_scalars = nullptr;  // [0]=n_ge, [1]=sum kept, [2]=sum all
    NVTB_CUDA_OK(cudaMallocAsync(&d_scalars, sizeof(long long) * 4, st));
    NVTB_CUDA_OK(cudaMemsetAsync(d_scalars, 0, sizeof(long long) * 4, st));
    if (freq_threshold > 0) {
      const int g = (int)std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
      count_ge_kernel<<<g, kThreads, 0, st>>>(s2, n, freq_threshold, d_scalars);
      NVTB_LAUNCH_OK();
      long long n_ge = 0;
      NVTB_CUDA_OK(cudaMemcpyAsync(&n_ge, d_scalars, sizeof(long long), cudaMemcpyDeviceToHost, st));
      NVTB_CUDA_OK(cudaStreamSynchronize(st));
      n_keep = n_ge;
    } else if (max_size > 0) {
      n_keep = std::min<int64_t>(n, max_size - (oov_count + 2));
    }
    size_t rb = 0;
    NVTB_CUDA_OK(cub::DeviceReduce::Sum(nullptr, rb, s2, (long long*)nullptr, n, st));
    if (rb > tmp_bytes) {
      NVTB_CUDA_OK(cudaFreeAsync(tmp, st));
      tmp_bytes = rb;
      NVTB_CUDA_OK(cudaMallocAsync(&tmp, tmp_bytes, st));
    }
    if (n_keep > 0)
      NVTB_CUDA_OK(cub::DeviceReduce::Sum(tmp, tmp_bytes, s2, d_scalars + 1, n_keep, st));
    NVTB_CUDA_OK(cub::DeviceReduce::Sum(tmp, tmp_bytes, s2, d_scalars + 2, n, st));
    long long sums[2] = {0, 0};
    NVTB_CUDA_OK(cudaMemcpyAsync(sums, d_scalars + 1, sizeof(long long) * 2, cudaMemcpyDeviceToHost, st));
    NVTB_CUDA_OK(cudaStreamSynchronize(st));
    v->info.unique_size = sums[0];
    v->info.oov_size = sums[1] - sums[0];
    NVTB_CUDA_OK(cudaFreeAsync(d_scalars, st));
    NVTB_CUDA_OK(cudaFreeAsync(tmp, st));

    if (n_keep > 0) {
      NVTB_CUDA_OK(cudaMallocAsync(&v->keys, sizeof(int64_t) * n_keep, st));
      NVTB_CUDA_OK(cudaMallocAsync(&v->sizes, sizeof(int64_t) * n_keep, st));
      NVTB_CUDA_OK(cudaMemcpyAsync(v->keys, k2, sizeof(int64_t) * n_keep, cudaMemcpyDeviceToDevice, st));
      NVTB_CUDA_OK(cudaMemcpyAsync(v->sizes, s2, sizeof(int64_t) * n_keep, cudaMemcpyDeviceToDevice, st));
    }
    NVTB_CUDA_OK(cudaFreeAsync(k2, st));
    NVTB_CUDA_OK(cudaFreeAsync(s2, st));
  }
  v->info.n_kept = n_keep;
  int rc = lookup_create(&v->t, v->keys, n_keep, false, st);
  if (rc) { nvtb_vocab_destroy(v); return rc; }
  NVTB_CUDA_OK(cudaStreamSynchronize(st));
  *out = v;
  return NVTB_OK;
}

int nvtb_vocab_from_arrays(nvtb_vocab_t** out, const int64_t* keys, const int64_t* sizes,
                           int64_t n, void* stream) {
  NVTB_REQUIRE(out != nullptr && n >= 0, "out NULL or n < 0");
  NVTB_REQUIRE(n == 0 || keys, "NULL keys");
  cudaStream_t st = (cudaStream_t)stream;
  nvtb_vocab* v = new (std::nothrow) nvtb_vocab();
  NVTB_REQUIRE(v != nullptr, "host allocation failed");
  memset(v, 0, sizeof(*v));
  v->info.n_kept = n;
  v->info.n_total = n;
  if (n > 0) {
    NVTB_CUDA_OK(cudaMallocAsync(&v->keys, sizeof(int64_t) * n, st));
    NVTB_CUDA_OK(cudaMemcpyAsync(v->keys, keys, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, st));
    if (sizes) {
      NVTB_CUDA_OK(cudaMallocAsync(&v->sizes, sizeof(int64_t) * n, st));
      NVTB_CUDA_OK(cudaMemcpyAsync(v->sizes, sizes, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, st));
    }
  }
  int rc = lookup_create(&v->t, v->keys, n, true, st);
  if (rc) { nvtb_vocab_destroy(v); return rc; }
  NVTB_CUDA_OK(cudaStreamSynchronize(st));
  *out = v;
  return NVTB_OK;
}

int nvtb_vocab_destroy(nvtb_vocab_t* v) {
  if (v == nullptr) return NVTB_OK;
  // stream-ordered frees on the legacy default stream: ordered after every kernel
  // that may still probe the table, without a device-wide host sync
  if (v->t.slots) cudaFreeAsync(v->t.slots, 0);
  if (v->keys) cudaFreeAsync(v->keys, 0);
  if (v->sizes) cudaFreeAsync(v->sizes, 0);
  delete v;
  return NVTB_OK;
}

int nvtb_vocab_info(const nvtb_vocab_t* v, nvtb_vocab_info_t* info) {
  NVTB_REQUIRE(v != nullptr && info != nullptr, "NULL argument");
  *info = v->info;
  return NVTB_OK;
}

int nvtb_vocab_export(const nvtb_vocab_t* v, int64_t* keys_out, int64_t* sizes_out, void* stream) {
  NVTB_REQUIRE(v != nullptr, "NULL vocab");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = v->info.n_kept;
  if (n == 0) return NVTB_OK;
  if (keys_out)
    NVTB_CUDA_OK(cudaMemcpyAsync(keys_out, v->keys, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, st));
  if (sizes_out) {
    NVTB_REQUIRE(v->sizes != nullptr, "vocab has no sizes");
    NVTB_CUDA_OK(cudaMemcpyAsync(sizes_out, v->sizes, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, st));
  }
  return NVTB_OK;
}

int nvtb_encode_apply(const nvtb_vocab_t* v, const nvtb_col_t* key, int64_t n,
                      int64_t null_label, int64_t oov_label, int64_t first_label,
                      uint64_t num_buckets, const nvtb_col_t* hash_cols, int n_hash_cols,
                      void* out, int out_dtype, void* stream) {
  NVTB_REQUIRE(v != nullptr && key != nullptr && n >= 0, "NULL argument or n < 0");
  NVTB_REQUIRE(key->dtype == NVTB_I32 || key->dtype == NVTB_I64, "key dtype must be int32 or int64");
  NVTB_REQUIRE(out_dtype == NVTB_I32 || out_dtype == NVTB_I64, "out_dtype must be int32 or int64");
  NVTB_REQUIRE(n_hash_cols >= 0 && n_hash_cols <= kMaxHashCols, "n_hash_cols must be in [0, 8]");
  NVTB_REQUIRE(n_hash_cols == 0 || hash_cols != nullptr, "hash_cols is NULL");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(key->data && out, "NULL data/out");
  HashCols hc;
  memset(&hc, 0, sizeof(hc));
  hc.ncols = n_hash_cols;
  for (int c = 0; c < n_hash_cols; ++c) {
    NVTB_REQUIRE(hash_cols[c].data != nullptr && hash_cols[c].dtype >= NVTB_I32 && hash_cols[c].dtype <= NVTB_H64,
                 "bad hash column");
    hc.data[c] = hash_cols[c].data; hc.mask[c] = hash_cols[c].validity; hc.dtype[c] = hash_cols[c].dtype;
  }
  EncodeParams p{null_label, oov_label, first_label, num_buckets};
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = scan_grid(n, 8);
  if (key->dtype == NVTB_I32) {
    if (out_dtype == NVTB_I64)
      encode_kernel<int32_t, int64_t><<<grid, kThreads, 0, st>>>((const int32_t*)key->data, key->validity, n, v->t, p, hc, (int64_t*)out);
    else
      encode_kernel<int32_t, int32_t><<<grid, kThreads, 0, st>>>((const int32_t*)key->data, key->validity, n, v->t, p, hc, (int32_t*)out);
  } else {
    if (out_dtype == NVTB_I64)
      encode_kernel<int64_t, int64_t><<<grid, kThreads, 0, st>>>((const int64_t*)key->data, key->validity, n, v->t, p, hc, (int64_t*)out);
    else
      encode_kernel<int64_t, int32_t><<<grid, kThreads, 0, st>>>((const int64_t*)key->data, key->validity, n, v->t, p, hc, (int32_t*)out);
  }
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

int nvtb_groupstats_create(nvtb_groupstats_t** out, const int64_t* keys, int64_t n_groups,
                           const double* stats, int width, int64_t null_row, void* stream) {
  NVTB_REQUIRE(out != nullptr && n_groups >= 0 && width >= 1, "bad arguments");
  NVTB_REQUIRE(n_groups == 0 || (keys && stats), "NULL keys/stats");
  NVTB_REQUIRE(null_row >= -1, "null_row must be >= -1");
  cudaStream_t st = (cudaStream_t)stream;
  nvtb_groupstats* g = new (std::nothrow) nvtb_groupstats();
  NVTB_REQUIRE(g != nullptr, "host allocation failed");
  memset(g, 0, sizeof(*g));
  g->n_groups = n_groups; g->width = width; g->null_row = null_row;
  // the stats matrix may have more rows than keys (the null group's row)
  const int64_t n_rows = std::max<int64_t>(n_groups, null_row + 1);
  if (n_rows > 0) {
    NVTB_CUDA_OK(cudaMallocAsync(&g->stats, sizeof(double) * n_rows * width, st));
    NVTB_CUDA_OK(cudaMemcpyAsync(g->stats, stats, sizeof(double) * n_rows * width, cudaMemcpyDeviceToDevice, st));
  }
  int rc = lookup_create(&g->t, keys, n_groups, false, st);
  if (rc) { nvtb_groupstats_destroy(g); return rc; }
  NVTB_CUDA_OK(cudaStreamSynchronize(st));
  *out = g;
  return NVTB_OK;
}

int nvtb_groupstats_destroy(nvtb_groupstats_t* g) {
  if (g == nullptr) return NVTB_OK;
  if (g->t.slots) cudaFreeAsync(g->t.slots, 0);
  if (g->stats) cudaFreeAsync(g->stats, 0);
  delete g;
  return NVTB_OK;
}

int nvtb_groupstats_gather(const nvtb_groupstats_t* g, const nvtb_col_t* key, int64_t n,
                           const int* cols, int ncols_out, const double* miss_vals,
                           void* const* out, const int* out_dtypes, void* stream) {
  NVTB_REQUIRE(g != nullptr && key != nullptr && n >= 0, "NULL argument or n < 0");
  NVTB_REQUIRE(key->dtype == NVTB_I32 || key->dtype == NVTB_I64, "key dtype must be int32 or int64");
  NVTB_REQUIRE(ncols_out >= 1 && ncols_out <= kMaxGatherCols, "ncols_out must be in [1, 16]");
  NVTB_REQUIRE(cols && miss_vals && out && out_dtypes, "NULL argument");
  if (n == 0) return NVTB_OK;
  NVTB_REQUIRE(key->data != nullptr, "key data is NULL");
  GatherOut go;
  memset(&go, 0, sizeof(go));
  go.ncols = ncols_out;
  for (int j = 0; j < ncols_out; ++j) {
    NVTB_REQUIRE(cols[j] >= 0 && cols[j] < g->width, "stats column out of range");
    NVTB_REQUIRE(out[j] != nullptr, "out column is NULL");
    NVTB_REQUIRE(out_dtypes[j] >= NVTB_I32 && out_dtypes[j] <= NVTB_F64, "bad out dtype");
    go.out[j] = out[j]; go.miss[j] = miss_vals[j]; go.col[j] = cols[j]; go.dtype[j] = out_dtypes[j];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)sm_count() * 8);
  if (key->dtype == NVTB_I32)
    gather_stats_kernel<int32_t><<<grid, kThreads, 0, st>>>((const int32_t*)key->data, key->validity, n, g->t, g->null_row, g->stats, g->width, go);
  else
    gather_stats_kernel<int64_t><<<grid, kThreads, 0, st>>>((const int64_t*)key->data, key->validity, n, g->t, g->null_row, g->stats, g->width, go);
  NVTB_LAUNCH_OK();
  return NVTB_OK;
}

}  // extern "C"
