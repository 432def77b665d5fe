// nla_mg.cuh -- single-process multi-GPU entry points (nla_mg_*, include/nextla_b200.h): one host thread drives every GPU of the box.
// Included at the end of nla_api.cu (it uses the per-GPU entry points defined there).
//
// The reference has no multi-device path.  The right-hand-side vectors of unified_rectrxm! are independent, so the one
// distributed strategy is: B sharded by RHS vector across the GPUs, A replicated -- broadcast from the GPU that holds it over
// NVLink 5 / NVSwitch with NCCL, in column panels in the order the schedule consumes them, every GPU's solve gated panel by panel
// (nla_rectrxm_gated) so that the broadcast hides behind the compute.  No other exchange.  The library owns the NCCL communicators
// (ncclCommInitAll), two streams per GPU (compute, broadcast), the per-panel events and the replicas of A.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that already carries NCCL -- torch, NCCL.jl -- shares its copy, a bare
// host gets the system library, and the single-GPU entry points never need it.
#pragma once
#include <dlfcn.h>

#include <thread>

namespace {

typedef void* nccl_comm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok() const { return CommInitAll && CommDestroy && Broadcast && GroupStart && GroupEnd; }
};

static bool load_nccl(NcclApi& api) {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return false;
  api.CommInitAll = (int (*)(nccl_comm_t*, int, const int*))dlsym(api.lib, "ncclCommInitAll");
  api.CommDestroy = (int (*)(nccl_comm_t))dlsym(api.lib, "ncclCommDestroy");
  api.Broadcast = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(api.lib, "ncclBroadcast");
  api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
  api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
  api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
  return api.ok();
}

constexpr int NCCL_UINT8 = 1;   // ncclUint8: the panels travel as bytes, whatever the element type

}  // namespace

struct nla_mg_context {
  uint32_t magic;
  int ngpu;
  std::vector<int> devices;
  std::vector<nla_handle_t> handles;
  NcclApi nccl;
  std::vector<nccl_comm_t> comms;
  std::vector<cudaStream_t> cstream, bstream, ustream;    // compute, broadcast, upload (host variant)
  std::vector<std::vector<cudaEvent_t>> panel_ev;         // [gpu][panel]: panel has arrived on that GPU
  std::vector<std::vector<cudaEvent_t>> up_ev;            // [gpu][panel]: panel uploaded from the host (host variant)
  std::vector<cudaEvent_t> done_ev;                       // [gpu]: previous call finished with the replica
  std::vector<void*> replica; std::vector<size_t> replica_bytes;
  int last_nccl, last_cuda;
};

static const uint32_t NLA_MG_MAGIC = 0x4e4c4d47u;  // "NLMG"
static inline bool mg_valid(nla_mg_t mg) { return mg && mg->magic == NLA_MG_MAGIC; }

#define MG_CUDA(mg, call)                                        \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) { (mg)->last_cuda = (int)e__; return NLA_ERR_CUDA; } \
  } while (0)
#define MG_NCCL(mg, call)                                        \
  do {                                                           \
    int r__ = (call);                                            \
    if (r__ != 0) { (mg)->last_nccl = r__; return NLA_ERR_NCCL; } \
  } while (0)

static void mg_panel_geometry(int64_t n, int64_t panels, int64_t& pc, int64_t& npan) {
  const int64_t gran = n >= 1024 * panels ? 1024 : 128;   // multiples of the block-inverse order let Float32/Float16 prepare panel by panel
  pc = (n + panels - 1) / panels;
  pc = (pc + gran - 1) / gran * gran;
  npan = (n + pc - 1) / pc;
}

static int mg_ensure_events(nla_mg_t mg, int64_t npan) {
  for (int i = 0; i < mg->ngpu; i++) {
    MG_CUDA(mg, cudaSetDevice(mg->devices[i]));
    while ((int64_t)mg->panel_ev[i].size() < npan) {
      cudaEvent_t e, u;
      MG_CUDA(mg, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      MG_CUDA(mg, cudaEventCreateWithFlags(&u, cudaEventDisableTiming));
      mg->panel_ev[i].push_back(e); mg->up_ev[i].push_back(u);
    }
  }
  return NLA_OK;
}

// replica of A on GPU i: n columns of pitch `ld` elements
static int mg_ensure_replica(nla_mg_t mg, int i, size_t bytes) {
  if (mg->replica_bytes[i] >= bytes) return NLA_OK;
  MG_CUDA(mg, cudaSetDevice(mg->devices[i]));
  if (mg->replica[i]) { MG_CUDA(mg, cudaStreamSynchronize(mg->cstream[i])); MG_CUDA(mg, cudaFree(mg->replica[i])); }
  mg->replica[i] = nullptr; mg->replica_bytes[i] = 0;
  MG_CUDA(mg, cudaMalloc(&mg->replica[i], bytes));
  mg->replica_bytes[i] = bytes;
  return NLA_OK;
}

extern "C" {

int nla_mg_create(nla_mg_t* out, int ngpu, const int* devices) {
  if (!out) return NLA_ERR_NULL_POINTER;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return NLA_ERR_NO_DEVICE;
  if (ngpu <= 0 || ngpu > count) return NLA_ERR_NO_DEVICE;
  for (int i = 0; i < ngpu; i++) {
    const int d = devices ? devices[i] : i;
    if (d < 0 || d >= count) return NLA_ERR_NO_DEVICE;
    for (int j = 0; j < i; j++) if ((devices ? devices[j] : j) == d) return NLA_ERR_INVALID_DIM;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  nla_mg_context* mg = new (std::nothrow) nla_mg_context();
  if (!mg) return NLA_ERR_UNSUPPORTED;
  mg->magic = NLA_MG_MAGIC; mg->ngpu = ngpu; mg->last_nccl = 0; mg->last_cuda = 0;
  mg->devices.resize(ngpu); mg->handles.assign(ngpu, nullptr); mg->comms.assign(ngpu, nullptr);
  mg->cstream.assign(ngpu, nullptr); mg->bstream.assign(ngpu, nullptr); mg->ustream.assign(ngpu, nullptr);
  mg->panel_ev.resize(ngpu); mg->up_ev.resize(ngpu); mg->done_ev.assign(ngpu, nullptr);
  mg->replica.assign(ngpu, nullptr); mg->replica_bytes.assign(ngpu, 0);
  int rc = NLA_OK;
  for (int i = 0; i < ngpu && rc == NLA_OK; i++) {
    mg->devices[i] = devices ? devices[i] : i;
    rc = nla_create(&mg->handles[i], mg->devices[i]);
    if (rc != NLA_OK) break;
    if (cudaSetDevice(mg->devices[i]) != cudaSuccess || cudaStreamCreateWithFlags(&mg->cstream[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&mg->bstream[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&mg->ustream[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&mg->done_ev[i], cudaEventDisableTiming) != cudaSuccess)
      rc = NLA_ERR_CUDA;
  }
  if (rc == NLA_OK && ngpu > 1) {
    if (!load_nccl(mg->nccl)) rc = NLA_ERR_NCCL;
    else if ((mg->last_nccl = mg->nccl.CommInitAll(mg->comms.data(), ngpu, mg->devices.data())) != 0) rc = NLA_ERR_NCCL;
  }
  cudaSetDevice(prev);
  if (rc != NLA_OK) { nla_mg_destroy(mg); return rc; }
  *out = mg;
  return NLA_OK;
}

int nla_mg_destroy(nla_mg_t mg) {
  if (!mg_valid(mg)) return NLA_ERR_INVALID_HANDLE;
  int prev = 0;
  cudaGetDevice(&prev);
  for (int i = 0; i < mg->ngpu; i++) {
    cudaSetDevice(mg->devices[i]);
    cudaDeviceSynchronize();
    if (mg->comms[i] && mg->nccl.CommDestroy) mg->nccl.CommDestroy(mg->comms[i]);
    for (auto e : mg->panel_ev[i]) cudaEventDestroy(e);
    for (auto e : mg->up_ev[i]) cudaEventDestroy(e);
    if (mg->done_ev[i]) cudaEventDestroy(mg->done_ev[i]);
    if (mg->cstream[i]) cudaStreamDestroy(mg->cstream[i]);
    if (mg->bstream[i]) cudaStreamDestroy(mg->bstream[i]);
    if (mg->ustream[i]) cudaStreamDestroy(mg->ustream[i]);
    if (mg->replica[i]) cudaFree(mg->replica[i]);
    if (mg->handles[i]) nla_destroy(mg->handles[i]);
  }
  cudaSetDevice(prev);
  mg->magic = 0;
  delete mg;
  return NLA_OK;
}

int nla_mg_device_count(nla_mg_t mg) { return mg_valid(mg) ? mg->ngpu : -1; }
nla_handle_t nla_mg_handle(nla_mg_t mg, int i) { return (mg_valid(mg) && i >= 0 && i < mg->ngpu) ? mg->handles[i] : nullptr; }
void* nla_mg_stream(nla_mg_t mg, int i) { return (mg_valid(mg) && i >= 0 && i < mg->ngpu) ? (void*)mg->cstream[i] : nullptr; }
int nla_mg_last_nccl_error(nla_mg_t mg) { return mg_valid(mg) ? mg->last_nccl : -1; }

int nla_mg_sync(nla_mg_t mg) {
  if (!mg_valid(mg)) return NLA_ERR_INVALID_HANDLE;
  int prev = 0;
  cudaGetDevice(&prev);
  int rc = NLA_OK;
  for (int i = 0; i < mg->ngpu; i++) {
    if (cudaSetDevice(mg->devices[i]) != cudaSuccess || cudaStreamSynchronize(mg->cstream[i]) != cudaSuccess ||
        cudaStreamSynchronize(mg->bstream[i]) != cudaSuccess) { mg->last_cuda = (int)cudaGetLastError(); rc = NLA_ERR_CUDA; }
  }
  cudaSetDevice(prev);
  return rc;
}

}  // extern "C"

// Broadcast of A in panels, consumption order.  src_of[p] = index of the GPU that holds panel p; bufs[i] = A on GPU i (pitch ld elements).
// wait_up: the root's broadcast stream first waits for its upload event of that panel (host variant).
static int mg_broadcast_panels(nla_mg_t mg, const std::vector<int64_t>& order, const std::vector<int>& src_of, const std::vector<void*>& bufs, int64_t n,
                               int64_t ld, size_t es, int64_t pc, bool wait_up) {
  for (int64_t p : order) {
    const int64_t c0 = p * pc, c1 = std::min(n, (p + 1) * pc);
    const size_t off = (size_t)c0 * ld * es, bytes = ((size_t)(c1 - c0 - 1) * ld + n) * es;   // whole columns of the panel; the last one without its padding
    const int root = src_of[(size_t)p];
    if (wait_up) {
      MG_CUDA(mg, cudaSetDevice(mg->devices[root]));
      MG_CUDA(mg, cudaStreamWaitEvent(mg->bstream[root], mg->up_ev[root][(size_t)p], 0));
    }
    if (mg->ngpu > 1) {
      MG_NCCL(mg, mg->nccl.GroupStart());
      for (int i = 0; i < mg->ngpu; i++) {
        char* b = (char*)bufs[i] + off;
        int r = mg->nccl.Broadcast(b, b, bytes, NCCL_UINT8, root, mg->comms[i], mg->bstream[i]);
        if (r != 0) { mg->nccl.GroupEnd(); mg->last_nccl = r; return NLA_ERR_NCCL; }
      }
      MG_NCCL(mg, mg->nccl.GroupEnd());
    }
    for (int i = 0; i < mg->ngpu; i++) {
      MG_CUDA(mg, cudaSetDevice(mg->devices[i]));
      MG_CUDA(mg, cudaEventRecord(mg->panel_ev[i][(size_t)p], mg->bstream[i]));
    }
  }
  return NLA_OK;
}

static int mg_check_args(nla_mg_t mg, char side, char uplo, char trans, char func, int dtype, int64_t n, const void* A, int64_t lda,
                         void* const* B_shards, const int64_t* shard_m, const int64_t* ldb) {
  if (!mg_valid(mg)) return NLA_ERR_INVALID_HANDLE;
  if (!B_shards || !shard_m || !ldb) return NLA_ERR_NULL_POINTER;
  Problem P;
  for (int i = 0; i < mg->ngpu; i++) {
    if (shard_m[i] < 0) return NLA_ERR_INVALID_DIM;
    int rc = make_problem(P, side, uplo, trans, func, dtype, n, shard_m[i], 1.0, A, lda, B_shards[i] ? B_shards[i] : (void*)16,
                          std::max<int64_t>(ldb[i], 1));
    if (rc != NLA_OK) return rc;
    if (n > 0 && shard_m[i] > 0 && !B_shards[i]) return NLA_ERR_NULL_POINTER;
  }
  return NLA_OK;
}

extern "C" {

int nla_mg_rectrxm(nla_mg_t mg, char side, char uplo, char trans, char func, int dtype, int64_t n, double alpha, int root, const void* A_root,
                   int64_t lda, void* const* B_shards, const int64_t* shard_m, const int64_t* ldb) {
  int rc = mg_check_args(mg, side, uplo, trans, func, dtype, n, A_root, lda, B_shards, shard_m, ldb);
  if (rc != NLA_OK) return rc;
  if (root < 0 || root >= mg->ngpu) return NLA_ERR_INVALID_DIM;
  if (n == 0) return NLA_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
  const size_t es = dtype_size(dtype);
  int64_t pc, npan;
  mg_panel_geometry(n, 8, pc, npan);
  std::vector<int64_t> order((size_t)npan);
  const int64_t cnt = nla_panel_order(side, uplo, trans, func, n, pc, order.data(), npan);
  if (cnt != npan) return NLA_ERR_UNSUPPORTED;
  if ((rc = mg_ensure_events(mg, npan)) != NLA_OK) return rc;
  std::vector<void*> bufs((size_t)mg->ngpu);
  for (int i = 0; i < mg->ngpu; i++) {
    if (i == root) { bufs[i] = const_cast<void*>(A_root); continue; }
    if ((rc = mg_ensure_replica(mg, i, (size_t)lda * n * es)) != NLA_OK) return rc;
    bufs[i] = mg->replica[i];
  }
  // the broadcast may only overwrite a replica once the previous call has finished reading it
  for (int i = 0; i < mg->ngpu; i++) {
    MG_CUDA(mg, cudaSetDevice(mg->devices[i]));
    MG_CUDA(mg, cudaStreamWaitEvent(mg->bstream[i], mg->done_ev[i], 0));
  }
  std::vector<int> src_of((size_t)npan, root);
  if ((rc = mg_broadcast_panels(mg, order, src_of, bufs, n, lda, es, pc, false)) != NLA_OK) return rc;
  for (int i = 0; i < mg->ngpu; i++) {
    MG_CUDA(mg, cudaSetDevice(mg->devices[i]));
    if (shard_m[i] > 0) {
      rc = nla_rectrxm_gated(mg->handles[i], side, uplo, trans, func, dtype, n, shard_m[i], alpha, bufs[i], lda, B_shards[i], ldb[i],
                             (void*)mg->cstream[i], pc, npan, (void* const*)mg->panel_ev[i].data());
      if (rc != NLA_OK) return rc;
    } else {
      MG_CUDA(mg, cudaStreamWaitEvent(mg->cstream[i], mg->panel_ev[i][(size_t)order.back()], 0));
    }
    MG_CUDA(mg, cudaEventRecord(mg->done_ev[i], mg->cstream[i]));
  }
  return NLA_OK;
}

int nla_mg_rectrxm_host(nla_mg_t mg, char side, char uplo, char trans, char func, int dtype, int64_t n, double alpha, const void* A_host,
                        int64_t lda, void* const* B_host_shards, const int64_t* shard_m, const int64_t* ldb) {
  int rc = mg_check_args(mg, side, uplo, trans, func, dtype, n, A_host, lda, B_host_shards, shard_m, ldb);
  if (rc != NLA_OK) return rc;
  if (n == 0) return NLA_OK;
  if (!A_host) return NLA_ERR_NULL_POINTER;
  int prev = 0;
  cudaGetDevice(&prev);
  struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
  const size_t es = dtype_size(dtype);
  int64_t pc, npan;
  mg_panel_geometry(n, 8, pc, npan);
  std::vector<int64_t> order((size_t)npan);
  if (nla_panel_order(side, uplo, trans, func, n, pc, order.data(), npan) != npan) return NLA_ERR_UNSUPPORTED;
  if ((rc = mg_ensure_events(mg, npan)) != NLA_OK) return rc;
  const int64_t ld = (n + 15) & ~15ll;   // device pitch of the replicas (16-byte rule of TMA for every element size)
  std::vector<void*> bufs((size_t)mg->ngpu);
  for (int i = 0; i < mg->ngpu; i++) {
    if ((rc = mg_ensure_replica(mg, i, (size_t)ld * n * es)) != NLA_OK) return rc;
    bufs[i] = mg->replica[i];
    MG_CUDA(mg, cudaSetDevice(mg->devices[i]));
    MG_CUDA(mg, cudaStreamWaitEvent(mg->bstream[i], mg->done_ev[i], 0));
    MG_CUDA(mg, cudaStreamWaitEvent(mg->ustream[i], mg->done_ev[i], 0));
  }
  // uploads: the k-th consumed panel goes through GPU k mod N's PCIe link (only the referenced trapezoid), all links at once
  std::vector<int> src_of((size_t)npan, 0);
  const bool a_lower = uplo == 'L';
  for (int64_t k = 0; k < npan; k++) {
    const int64_t p = order[(size_t)k];
    const int g = (int)(k % mg->ngpu);
    src_of[(size_t)p] = g;
    const int64_t c0 = p * pc, c1 = std::min(n, (p + 1) * pc);
    const int64_t r0 = a_lower ? c0 : 0, r1 = a_lower ? n : c1;
    MG_CUDA(mg, cudaSetDevice(mg->devices[g]));
    MG_CUDA(mg, cudaMemcpy2DAsync((char*)bufs[g] + ((size_t)c0 * ld + r0) * es, (size_t)ld * es, (const char*)A_host + ((size_t)c0 * lda + r0) * es,
                                  (size_t)lda * es, (size_t)(r1 - r0) * es, (size_t)(c1 - c0), cudaMemcpyHostToDevice, mg->ustream[g]));
    MG_CUDA(mg, cudaEventRecord(mg->up_ev[g][(size_t)p], mg->ustream[g]));
  }
  if ((rc = mg_broadcast_panels(mg, order, src_of, bufs, n, ld, es, pc, true)) != NLA_OK) return rc;
  // every GPU streams its own shard of B through the host pipeline (synchronous per GPU): one host thread per GPU
  std::vector<int> rcs((size_t)mg->ngpu, NLA_OK);
  std::vector<std::thread> workers;
  for (int i = 0; i < mg->ngpu; i++) {
    if (shard_m[i] <= 0) continue;
    workers.emplace_back([&, i]() {
      rcs[(size_t)i] = nla_rectrxm_hostb_gated(mg->handles[i], side, uplo, trans, func, dtype, n, shard_m[i], alpha, bufs[i], ld, B_host_shards[i],
                                               ldb[i], pc, npan, (void* const*)mg->panel_ev[i].data());
    });
  }
  for (auto& w : workers) w.join();
  for (int i = 0; i < mg->ngpu; i++) {
    if (rcs[(size_t)i] != NLA_OK) return rcs[(size_t)i];
    MG_CUDA(mg, cudaSetDevice(mg->devices[i]));
    MG_CUDA(mg, cudaStreamSynchronize(mg->bstream[i]));
    MG_CUDA(mg, cudaEventRecord(mg->done_ev[i], mg->cstream[i]));
  }
  return NLA_OK;
}

}  // extern "C"
