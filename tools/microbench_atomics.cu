// Microbenchmarks behind the K3 (hash-aggregate) design decisions in DESIGN.md: how fast can
// one B200 count 2^26 int32 keys with each primitive?  Build + run:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/microbench_atomics.cu -o /tmp/mb && /tmp/mb
// Output (one line per variant): name, distinct keys, table bytes, us per 2^26 rows, Grows/s.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h;
}

__global__ void gen_keys(int32_t* keys, int64_t n, uint32_t k, uint32_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r = mix32((uint32_t)i * 2654435761u + seed);
    uint32_t id = (uint32_t)(((uint64_t)r * k) >> 32);
    keys[i] = (int32_t)((id * 2654435761u) & 0x7FFFFFFFu);
  }
}

__device__ __forceinline__ void ld8(const int32_t* p, int32_t (&v)[8]) {
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// 0: read only (HBM floor of this access pattern)
__global__ void __launch_bounds__(256) k_read(const int32_t* keys, int64_t n, unsigned long long* out) {
  unsigned acc = 0;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 8; i < n; i += (int64_t)gridDim.x * blockDim.x * 8) {
    int32_t v[8]; ld8(keys + i, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += mix32((uint32_t)v[j]);
  }
  if (acc == 0x12345u) out[0] = acc;
}

// A: one global RED (no return) per row, 64-bit or 32-bit counters, direct-addressed by hash
template <typename C>
__global__ void __launch_bounds__(256) k_redg(const int32_t* keys, int64_t n, C* table, uint32_t mask) {
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 8; i < n; i += (int64_t)gridDim.x * blockDim.x * 8) {
    int32_t v[8]; ld8(keys + i, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&table[mix32((uint32_t)v[j]) & mask], (C)1);
  }
}

// B: the real global path: 32-byte sector load of a 4-way bucket, compare, RED on the hit word
__global__ void __launch_bounds__(256) k_probe_red(const int32_t* keys, int64_t n, unsigned long long* table, uint32_t bmask) {
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 8; i < n; i += (int64_t)gridDim.x * blockDim.x * 8) {
    int32_t v[8]; ld8(keys + i, v);
    ulonglong4 w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned long long* p = table + 4ull * (mix32((uint32_t)v[j]) & bmask);
      asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[j].x), "=l"(w[j].y), "=l"(w[j].z), "=l"(w[j].w) : "l"(p));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      unsigned long long* p = table + 4ull * (mix32((uint32_t)v[j]) & bmask);
      const unsigned key = (unsigned)v[j];
      int s = ((unsigned)w[j].y == key) ? 1 : ((unsigned)w[j].z == key) ? 2 : ((unsigned)w[j].w == key) ? 3 : 0;
      atomicAdd(reinterpret_cast<unsigned*>(p + s) + 1, 1u);
    }
  }
}

// C: shared-memory atomics, one per row, 8192-word table per CTA (no flush: throughput only)
template <int MODE>   // 0 = ATOMS.ADD u32, 1 = plain LDS+STS (racy), 2 = match_any + leader ATOMS
__global__ void __launch_bounds__(256, 3) k_smem(const int32_t* keys, int64_t n, unsigned* out) {
  __shared__ unsigned tab[8192];
  for (int s = threadIdx.x; s < 8192; s += 256) tab[s] = 0;
  __syncthreads();
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 8; i < n; i += (int64_t)gridDim.x * blockDim.x * 8) {
    int32_t v[8]; ld8(keys + i, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned s = mix32((uint32_t)v[j]) >> 19;
      if (MODE == 0) atomicAdd(&tab[s], 1u);
      else if (MODE == 1) tab[s] = tab[s] + 1u;
      else {
        const unsigned grp = __match_any_sync(0xFFFFFFFFu, s);
        if ((threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&tab[s], (unsigned)__popc(grp));
      }
    }
  }
  __syncthreads();
  unsigned acc = 0;
  for (int s = threadIdx.x; s < 8192; s += 256) acc += tab[s];
  if (acc == 0x12345u) out[0] = acc;
}

// D: warp-private counting without atomics: each warp sorts nothing; every lane owns the
// table words with (slot & 31) == lane and receives keys by 32 shuffles (owner-computes)
__global__ void __launch_bounds__(256, 3) k_owner(const int32_t* keys, int64_t n, unsigned* out) {
  __shared__ unsigned tab[8192];
  for (int s = threadIdx.x; s < 8192; s += 256) tab[s] = 0;
  __syncthreads();
  // 8-bit digit histogram per CTA with match_any (the cub onesweep ranking primitive)
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 8; i < n; i += (int64_t)gridDim.x * blockDim.x * 8) {
    int32_t v[8]; ld8(keys + i, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned d = mix32((uint32_t)v[j]) >> 24;
      const unsigned grp = __match_any_sync(0xFFFFFFFFu, d);
      if ((threadIdx.x & 31) == __ffs(grp) - 1) tab[(threadIdx.x >> 5) * 256 + d] += (unsigned)__popc(grp);
      __syncwarp();
    }
  }
  __syncthreads();
  unsigned acc = 0;
  for (int s = threadIdx.x; s < 8192; s += 256) acc += tab[s];
  if (acc == 0x12345u) out[0] = acc;
}

template <typename F>
static float time_us(F&& launch, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int r = 0; r < reps; ++r) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms * 1e3f / reps;
}

int main() {
  const int64_t n = 1ll << 26;
  int32_t* keys; CK(cudaMalloc(&keys, n * 4));
  unsigned long long* table; CK(cudaMalloc(&table, (1ull << 27) * 8));     // up to 2^27 words (1 GiB)
  unsigned* out; CK(cudaMalloc(&out, 64));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = sms * 8;
  const uint32_t cards[] = {3u, 63u, 1543u, 7120u, 39043u, 100000u, 2000000u, 4450000u, 30000000u};
  printf("variant,distinct,table_bytes,us,Grows_per_s\n");
  for (uint32_t k : cards) {
    gen_keys<<<grid, 256>>>(keys, n, k, 17u);
    CK(cudaDeviceSynchronize());
    auto rep = [&](const char* name, uint64_t tb, float us) {
      printf("%s,%u,%llu,%.1f,%.1f\n", name, k, (unsigned long long)tb, us, n / us * 1e-3);
      fflush(stdout);
    };
    rep("read_only", 0, time_us([&] { k_read<<<grid, 256>>>(keys, n, table); }));
    // table sized ~2x distinct, power of two, >= 2^13 words
    uint32_t cap = 1u << 13;
    while (cap < 2u * k) cap <<= 1;
    CK(cudaMemset(table, 0, (size_t)cap * 8));
    rep("redg_u64_direct", (uint64_t)cap * 8, time_us([&] { k_redg<unsigned long long><<<grid, 256>>>(keys, n, table, cap - 1); }));
    rep("redg_u32_direct", (uint64_t)cap * 4, time_us([&] { k_redg<unsigned><<<grid, 256>>>(keys, n, reinterpret_cast<unsigned*>(table), cap - 1); }));
    rep("probe32B_then_redg", (uint64_t)cap * 8, time_us([&] { k_probe_red<<<grid, 256>>>(keys, n, table, cap / 4 - 1); }));
    rep("smem_atoms", 32768, time_us([&] { k_smem<0><<<sms * 3, 256>>>(keys, n, out); }));
    rep("smem_plain_rmw_racy", 32768, time_us([&] { k_smem<1><<<sms * 3, 256>>>(keys, n, out); }));
    rep("smem_match_any_leader", 32768, time_us([&] { k_smem<2><<<sms * 3, 256>>>(keys, n, out); }));
    rep("warp_digit_hist_match", 32768, time_us([&] { k_owner<<<sms * 3, 256>>>(keys, n, out); }));
  }
  return 0;
}
