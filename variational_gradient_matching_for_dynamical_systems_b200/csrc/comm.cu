// Multi-GPU exchange steps of the hot path as thin NCCL wrappers (include/vgm_b200.h "vgm_comm_*").
//
// The path has exactly two exchanges per coordinate-ascent iteration:
//   * the sum over ODE rows k of the parameter statistics, proxies_for_ode_parameters_and_states.py:64-79
//     (`local_mean += ...`, `local_scaling += ...`): rows live on different ranks -> one all-reduce of p + p^2 doubles;
//   * the sweep-start snapshot of the neighbours' boundary states, proxies...py:191 (`state_proxy.values` copied once
//     per sweep; the coefficients of state u read states u-3 .. u+3 for Lorenz96) -> every rank publishes the owned
//     rows somebody else reads, one all-gather, every rank picks its halo rows.
// NCCL is resolved at run time with dlopen (the copy already loaded into the process -- e.g. the one PyTorch brings --
// or libnccl.so.2 from the loader path), so libvgm_b200.so has no link-time dependency on it and loads on a machine
// without NCCL; the entry points then return VGM_ENCCL.  The pack / unpack around the all-gather are two small
// kernels on the caller's stream: nothing here synchronises the device.
#include <dlfcn.h>
#include <stdlib.h>

#include <mutex>

#include "vgm_common.cuh"

namespace vgm {

// The handful of NCCL declarations used here (stable across NCCL 2.x; nccl.h: ncclUniqueId is 128 opaque bytes,
// ncclFloat64 = 8, ncclSum = 0, ncclSuccess = 0).
struct NcclUniqueId {
  char internal[VGM_COMM_ID_BYTES];
};
typedef void* NcclComm;
struct NcclApi {
  void* handle = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

static void nccl_load() {
  // prefer the instance that is already in the process (one NCCL per process: the caller's framework and this
  // library then share topology detection and proxy threads, and a framework imported LATER does not find an older
  // libnccl.so.2 bound to the soname); VGM_NCCL_LIB names a specific file.
  void* h = nullptr;
  if (const char* path = getenv("VGM_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return;
  g_nccl.handle = h;
#define VGM_NCCL_SYM(field, name)                                             \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                  \
  if (!g_nccl.field) return;
  VGM_NCCL_SYM(GetVersion, "ncclGetVersion")
  VGM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  VGM_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  VGM_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  VGM_NCCL_SYM(AllReduce, "ncclAllReduce")
  VGM_NCCL_SYM(AllGather, "ncclAllGather")
  VGM_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef VGM_NCCL_SYM
  g_nccl.ok = true;
}

static int nccl_api() {
  std::call_once(g_nccl_once, nccl_load);
  if (!g_nccl.ok) {
    set_error("NCCL is not available: libnccl.so.2 could not be loaded (%s)", g_nccl.handle ? "missing symbol" : "dlopen failed");
    return VGM_ENCCL;
  }
  return VGM_OK;
}

#define VGM_NCCL_OK(expr)                                                                              \
  do {                                                                                                 \
    int _r = (expr);                                                                                   \
    if (_r != 0) {                                                                                     \
      vgm::set_error("%s:%d: %s -> NCCL error %d (%s)", __FILE__, __LINE__, #expr, _r, vgm::g_nccl.GetErrorString(_r)); \
      return VGM_ENCCL;                                                                                \
    }                                                                                                  \
  } while (0)

// One 16-byte piece per thread; grid.y = row.  Rows [n_mine, n_pub) of the published block are padding (every rank
// contributes the same count to the all-gather) and are zero-filled so the buffer never holds stale data.
__global__ void __launch_bounds__(256) halo_pack_kernel(const double* __restrict__ X, int ld, const int* __restrict__ pub_rows,
                                                        int n_mine, double* __restrict__ pub) {
  const int j = blockIdx.y;
  const double2* src = (j < n_mine) ? reinterpret_cast<const double2*>(X + (size_t)pub_rows[j] * ld) : nullptr;
  double2* dst = reinterpret_cast<double2*>(pub + (size_t)j * ld);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ld / 2; c += gridDim.x * blockDim.x)
    dst[c] = src ? src[c] : make_double2(0.0, 0.0);
}

__global__ void __launch_bounds__(256) halo_unpack_kernel(double* __restrict__ X, int ld, const int* __restrict__ halo_rows,
                                                          const int* __restrict__ halo_src, const double* __restrict__ gathered) {
  const int j = blockIdx.y;
  const double2* src = reinterpret_cast<const double2*>(gathered + (size_t)halo_src[j] * ld);
  double2* dst = reinterpret_cast<double2*>(X + (size_t)halo_rows[j] * ld);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ld / 2; c += gridDim.x * blockDim.x) dst[c] = src[c];
}

}  // namespace vgm

struct vgm_comm {
  vgm::NcclComm nccl;
  int rank, world, device;
};

extern "C" int vgm_comm_nccl_version(int* version) {
  VGM_TRY(vgm::nccl_api());
  VGM_REQUIRE(version != nullptr, "null output");
  VGM_NCCL_OK(vgm::g_nccl.GetVersion(version));
  return VGM_OK;
}

extern "C" int vgm_comm_unique_id(void* id_host) {
  VGM_TRY(vgm::nccl_api());
  VGM_REQUIRE(id_host != nullptr, "null id buffer");
  vgm::NcclUniqueId id;
  VGM_NCCL_OK(vgm::g_nccl.GetUniqueId(&id));
  memcpy(id_host, id.internal, VGM_COMM_ID_BYTES);
  return VGM_OK;
}

extern "C" int vgm_comm_init(const void* id_host, int rank, int world, vgm_comm** comm) {
  VGM_TRY(vgm::nccl_api());
  VGM_REQUIRE(id_host && comm, "null argument");
  VGM_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank outside [0, world)");
  vgm::NcclUniqueId id;
  memcpy(id.internal, id_host, VGM_COMM_ID_BYTES);
  vgm_comm* c = new vgm_comm();
  c->rank = rank;
  c->world = world;
  if (cudaGetDevice(&c->device) != cudaSuccess) {
    delete c;
    vgm::set_error("vgm_comm_init: no current CUDA device");
    return VGM_ECUDA;
  }
  int r = vgm::g_nccl.CommInitRank(&c->nccl, world, id, rank);
  if (r != 0) {
    delete c;
    vgm::set_error("ncclCommInitRank(rank %d of %d) -> NCCL error %d (%s)", rank, world, r, vgm::g_nccl.GetErrorString(r));
    return VGM_ENCCL;
  }
  *comm = c;
  return VGM_OK;
}

extern "C" int vgm_comm_free(vgm_comm* comm) {
  if (!comm) return VGM_OK;
  int r = vgm::g_nccl.ok ? vgm::g_nccl.CommDestroy(comm->nccl) : 0;
  delete comm;
  if (r != 0) {
    vgm::set_error("ncclCommDestroy -> NCCL error %d", r);
    return VGM_ENCCL;
  }
  return VGM_OK;
}

extern "C" int vgm_comm_rank(const vgm_comm* comm, int* rank, int* world) {
  VGM_REQUIRE(comm != nullptr, "null communicator");
  if (rank) *rank = comm->rank;
  if (world) *world = comm->world;
  return VGM_OK;
}

extern "C" int vgm_comm_allreduce_stats(vgm_comm* comm, double* stats, int n, vgm_stream_t stream) {
  VGM_REQUIRE(comm != nullptr && stats != nullptr, "null argument");
  VGM_REQUIRE(n > 0, "empty statistics");
  VGM_NCCL_OK(vgm::g_nccl.AllReduce(stats, stats, (size_t)n, vgm::NCCL_FLOAT64, vgm::NCCL_SUM, comm->nccl, (cudaStream_t)stream));
  return VGM_OK;
}

extern "C" size_t vgm_comm_halo_ws_bytes(int world, int n_pub, int ld) {
  if (world < 1 || n_pub < 0 || ld < 0) return 0;
  return (size_t)(world + 1) * (size_t)n_pub * (size_t)ld * sizeof(double);
}

extern "C" int vgm_comm_halo_exchange(vgm_comm* comm, double* X, int ld, const int* pub_rows, int n_pub_mine, int n_pub,
                                      const int* halo_rows, const int* halo_src, int n_halo, void* ws, size_t ws_bytes,
                                      vgm_stream_t stream) {
  VGM_REQUIRE(comm != nullptr && X != nullptr, "null argument");
  VGM_REQUIRE(ld > 0 && (ld % 2) == 0 && ((uintptr_t)X % 16) == 0, "rows must be 16-byte aligned (ld even)");
  VGM_REQUIRE(n_pub >= 0 && n_pub_mine >= 0 && n_pub_mine <= n_pub && n_halo >= 0, "bad counts");
  if (n_pub == 0) return VGM_OK;                       /* trajectory batches: nobody reads anybody's rows */
  VGM_REQUIRE(n_pub_mine == 0 || pub_rows != nullptr, "null pub_rows");
  VGM_REQUIRE(n_halo == 0 || (halo_rows != nullptr && halo_src != nullptr), "null halo tables");
  VGM_REQUIRE(ws != nullptr && ((uintptr_t)ws % 16) == 0 && ws_bytes >= vgm_comm_halo_ws_bytes(comm->world, n_pub, ld),
              "workspace too small (vgm_comm_halo_ws_bytes) or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  double* pub = static_cast<double*>(ws);
  double* gathered = pub + (size_t)n_pub * ld;
  const int gx = vgm::ceil_div(ld / 2, 256);
  vgm::halo_pack_kernel<<<dim3(gx, n_pub), 256, 0, s>>>(X, ld, pub_rows, n_pub_mine, pub);
  VGM_LAUNCH_OK();
  VGM_NCCL_OK(vgm::g_nccl.AllGather(pub, gathered, (size_t)n_pub * ld, vgm::NCCL_FLOAT64, comm->nccl, s));
  if (n_halo > 0) {
    vgm::halo_unpack_kernel<<<dim3(gx, n_halo), 256, 0, s>>>(X, ld, halo_rows, halo_src, gathered);
    VGM_LAUNCH_OK();
  }
  return VGM_OK;
}
