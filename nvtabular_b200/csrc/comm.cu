// comm.cu — the cross-GPU collectives of the fit path behind the C-ABI:
// nvtb_comm_t wraps an ncclComm_t (created here from a unique id the host runtime distributes,
// or provided by the caller) and exposes exactly the exchanges the path has:
//   * moments:            all-reduce of {count, sum, sumsq | min | max} per column
//   * group-by partials:  all-to-all of variable-sized row blocks (grouped ncclSend/ncclRecv)
//   * vocabulary shards:  all-gather of equal-sized blocks
// They replace the dask tree reduction over TCP/UCX and the shared-filesystem "broadcast" of
// reference nvtabular/ops/categorify.py:1399-1540, 1627-1643 and ops/moments.py:34-57.
// NCCL is resolved at run time (dlopen): the process uses the instance the host framework has
// already loaded (torch bundles libnccl.so.2), never a second copy.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>
#include <new>

#include "common.cuh"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the instance already in the process
  if (h == nullptr) {
    const char* p = getenv("NVTB_NCCL_LIB");
    if (p && *p) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
  }
  if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) return;
  g_nccl.lib = h;
#define NVTB_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name)
  NVTB_SYM(GetUniqueId, "ncclGetUniqueId");
  NVTB_SYM(CommInitRank, "ncclCommInitRank");
  NVTB_SYM(CommDestroy, "ncclCommDestroy");
  NVTB_SYM(AllReduce, "ncclAllReduce");
  NVTB_SYM(AllGather, "ncclAllGather");
  NVTB_SYM(Send, "ncclSend");
  NVTB_SYM(Recv, "ncclRecv");
  NVTB_SYM(GroupStart, "ncclGroupStart");
  NVTB_SYM(GroupEnd, "ncclGroupEnd");
  NVTB_SYM(GetErrorString, "ncclGetErrorString");
#undef NVTB_SYM
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce &&
              g_nccl.AllGather && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd;
}

int need_nccl() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) {
    nvtb::set_error("NCCL is not available in this process (libnccl.so.2 not found; set NVTB_NCCL_LIB)");
    return NVTB_ENCCL;
  }
  return NVTB_OK;
}

#define NVTB_NCCL_OK(expr)                                                                   \
  do {                                                                                        \
    ncclResult_t r_ = (expr);                                                                 \
    if (r_ != ncclSuccess) {                                                                  \
      nvtb::set_error("%s failed: %s (%s:%d)", #expr,                                         \
                      g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error", __FILE__, __LINE__); \
      return NVTB_ENCCL;                                                                      \
    }                                                                                         \
  } while (0)

}  // namespace

struct nvtb_comm {
  ncclComm_t comm;
  int rank, world;
  bool owned;
};

extern "C" {

int nvtb_comm_available(void) { return need_nccl() == NVTB_OK ? 1 : 0; }

int nvtb_comm_unique_id(uint8_t* id_out128) {
  NVTB_REQUIRE(id_out128 != nullptr, "NULL id buffer");
  int rc = need_nccl();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NVTB_NCCL_OK(g_nccl.GetUniqueId(&id));
  memcpy(id_out128, &id, sizeof(id));
  return NVTB_OK;
}

int nvtb_comm_create(nvtb_comm_t** out, const uint8_t* id128, int rank, int world) {
  NVTB_REQUIRE(out != nullptr && id128 != nullptr && world >= 1 && rank >= 0 && rank < world, "bad arguments");
  int rc = need_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c;
  NVTB_NCCL_OK(g_nccl.CommInitRank(&c, world, id, rank));
  nvtb_comm* h = new (std::nothrow) nvtb_comm{c, rank, world, true};
  NVTB_REQUIRE(h != nullptr, "host allocation failed");
  *out = h;
  return NVTB_OK;
}

int nvtb_comm_wrap(nvtb_comm_t** out, void* nccl_comm, int rank, int world) {
  NVTB_REQUIRE(out != nullptr && nccl_comm != nullptr && world >= 1 && rank >= 0 && rank < world, "bad arguments");
  int rc = need_nccl();
  if (rc) return rc;
  nvtb_comm* h = new (std::nothrow) nvtb_comm{(ncclComm_t)nccl_comm, rank, world, false};
  NVTB_REQUIRE(h != nullptr, "host allocation failed");
  *out = h;
  return NVTB_OK;
}

int nvtb_comm_destroy(nvtb_comm_t* c) {
  if (c == nullptr) return NVTB_OK;
  if (c->owned && g_nccl.ok) g_nccl.CommDestroy(c->comm);
  delete c;
  return NVTB_OK;
}

int nvtb_comm_rank(const nvtb_comm_t* c, int* rank, int* world) {
  NVTB_REQUIRE(c != nullptr, "NULL comm");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return NVTB_OK;
}

// in-place all-reduce of n doubles (op: 0 sum, 1 min, 2 max) / int64s
int nvtb_comm_allreduce_f64(nvtb_comm_t* c, double* buf, int64_t n, int op, void* stream) {
  NVTB_REQUIRE(c != nullptr && n >= 0 && op >= 0 && op <= 2, "bad arguments");
  if (n == 0) return NVTB_OK;
  const ncclRedOp_t o = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
  NVTB_NCCL_OK(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat64, o, c->comm, (cudaStream_t)stream));
  return NVTB_OK;
}

int nvtb_comm_allreduce_i64(nvtb_comm_t* c, int64_t* buf, int64_t n, int op, void* stream) {
  NVTB_REQUIRE(c != nullptr && n >= 0 && op >= 0 && op <= 2, "bad arguments");
  if (n == 0) return NVTB_OK;
  const ncclRedOp_t o = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
  NVTB_NCCL_OK(g_nccl.AllReduce(buf, buf, (size_t)n, ncclInt64, o, c->comm, (cudaStream_t)stream));
  return NVTB_OK;
}

// moments accumulator [ncols][5] = {count, sum, sumsq, min, max}: three all-reduces over strided
// views are avoided by reducing the whole block with each operator into scratch and stitching
// (ncols is 13: the block is 520 bytes)
__global__ void moments_stitch_kernel(double* acc, const double* mn, const double* mx, int ncols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncols) { acc[c * 5 + 3] = mn[c * 5 + 3]; acc[c * 5 + 4] = mx[c * 5 + 4]; }
}

int nvtb_moments_allreduce(nvtb_comm_t* c, double* acc_dev, int ncols, void* stream) {
  NVTB_REQUIRE(c != nullptr && acc_dev != nullptr && ncols >= 0, "bad arguments");
  if (ncols == 0 || c->world == 1) return NVTB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ncols * 5;
  double* tmp = nullptr;
  NVTB_CUDA_OK(cudaMallocAsync(&tmp, sizeof(double) * 2 * n, st));
  NVTB_NCCL_OK(g_nccl.GroupStart());
  NVTB_NCCL_OK(g_nccl.AllReduce(acc_dev, tmp, n, ncclFloat64, ncclMin, c->comm, st));
  NVTB_NCCL_OK(g_nccl.AllReduce(acc_dev, tmp + n, n, ncclFloat64, ncclMax, c->comm, st));
  NVTB_NCCL_OK(g_nccl.GroupEnd());
  NVTB_NCCL_OK(g_nccl.AllReduce(acc_dev, acc_dev, n, ncclFloat64, ncclSum, c->comm, st));
  moments_stitch_kernel<<<(ncols + 63) / 64, 64, 0, st>>>(acc_dev, tmp, tmp + n, ncols);
  NVTB_LAUNCH_OK();
  NVTB_CUDA_OK(cudaFreeAsync(tmp, st));
  return NVTB_OK;
}

// all-gather of one block of `bytes` bytes per rank
int nvtb_comm_allgather(nvtb_comm_t* c, const void* send, void* recv, int64_t bytes, void* stream) {
  NVTB_REQUIRE(c != nullptr && bytes >= 0, "bad arguments");
  if (bytes == 0) return NVTB_OK;
  NVTB_REQUIRE(send != nullptr && recv != nullptr, "NULL buffer");
  NVTB_NCCL_OK(g_nccl.AllGather(send, recv, (size_t)bytes, ncclUint8, c->comm, (cudaStream_t)stream));
  return NVTB_OK;
}

// all-to-all of variable-sized blocks: send_counts_host[r] elements of elem_bytes go to rank r
// (consecutive in `send`), recv_counts_host[r] arrive from rank r (consecutive in `recv`)
int nvtb_comm_alltoallv(nvtb_comm_t* c, const void* send, const int64_t* send_counts_host, void* recv,
                        const int64_t* recv_counts_host, int elem_bytes, void* stream) {
  NVTB_REQUIRE(c != nullptr && send_counts_host != nullptr && recv_counts_host != nullptr && elem_bytes > 0,
               "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const char* s = (const char*)send;
  char* r = (char*)recv;
  NVTB_NCCL_OK(g_nccl.GroupStart());
  for (int p = 0; p < c->world; ++p) {
    const size_t sb = (size_t)send_counts_host[p] * elem_bytes, rb = (size_t)recv_counts_host[p] * elem_bytes;
    if (sb) NVTB_NCCL_OK(g_nccl.Send(s, sb, ncclUint8, p, c->comm, st));
    if (rb) NVTB_NCCL_OK(g_nccl.Recv(r, rb, ncclUint8, p, c->comm, st));
    s += sb;
    r += rb;
  }
  NVTB_NCCL_OK(g_nccl.GroupEnd());
  return NVTB_OK;
}

}  // extern "C"
