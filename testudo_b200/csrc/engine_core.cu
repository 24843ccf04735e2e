// Engine core: contexts (one per GPU, all driven from ONE process), worker threads, the NCCL clique, and the
// host-facing entry points that shard over the GPUs (include/testudo_b200.h). No kernels live in this unit and there is
// no CPU arithmetic path: without a CUDA device tb200_init fails and every entry point returns TB200_E_STATE.
//
// Multi-GPU model (the reference is one process that fans rows out internally, src/sqrt_pst.rs:121-125):
//   * tb200_init_devices(devices, n) creates one context per GPU. devices[0] is the PRIMARY device: every
//     single-device entry point (the *_dev calls, MIPP, PST openings) runs there.
//   * A single large MSM is split into contiguous POINT RANGES, row commitments into ROW RANGES over the replicated
//     SRS tables, a pairing product into PAIR RANGES. Each GPU's share is driven by its own worker thread (its own
//     current device, its own -- possibly blocking -- copies from the caller's host buffers), uploads are chunked and
//     overlap the compute of the previous chunk.
//   * The per-GPU partial results (96-B points / 576-B Miller products) are combined by ONE ncclAllGather over NVLink
//     inside the call, then summed / multiplied + final-exponentiated on the primary device. Row commitments need no
//     collective: every GPU writes its rows straight into the caller's output buffer.
#include <dlfcn.h>
#include <nccl.h>
#include <sched.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>

#include "engine.h"

namespace tbe {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
std::mutex g_mu;
Engine E;

Engine::~Engine() {
  for (auto& d : devs) {
    if (!d || !d->worker.joinable()) continue;
    {
      std::lock_guard<std::mutex> lk(d->wmu);
      d->quit = true;
      d->wcv.notify_all();
    }
    d->worker.join();
  }
  for (auto& d : devs) d.release();  // contexts are left to the driver's own teardown
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

int need_ready() {
  if (!E.ready) return fail(TB200_E_STATE, "tb200_init has not been called (or failed): no CUDA context");
  return 0;
}

// ---- arena -----------------------------------------------------------------------------------------------------------
int Arena::reserve(size_t bytes) {
  if (bytes <= cap) return 0;
  if (base) {
    cudaError_t e = cudaFree(base);  // implicit device synchronisation: earlier users of the scratch have finished
    if (e != cudaSuccess) return (int)e;
    base = nullptr;
    cap = 0;
    used = false;
  }
  size_t want = bytes + (bytes >> 3);
  cudaError_t e = cudaMalloc((void**)&base, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    e = cudaMalloc((void**)&base, bytes);
    want = bytes;
    if (e != cudaSuccess) return (int)e;
  }
  cap = want;
  return 0;
}
int Arena::acquire(cudaStream_t st) {
  if (!last_use) CU(cudaEventCreateWithFlags(&last_use, cudaEventDisableTiming));
  if (used) CU(cudaStreamWaitEvent(st, last_use, 0));
  return 0;
}
int Arena::release(cudaStream_t st) {
  if (!last_use) CU(cudaEventCreateWithFlags(&last_use, cudaEventDisableTiming));
  CU(cudaEventRecord(last_use, st));
  used = true;
  return 0;
}
void Arena::destroy() {
  if (base) cudaFree(base);
  if (last_use) cudaEventDestroy(last_use);
  *this = Arena();
}

// ---- stage timers (labels mirror the reference's Timer names where one exists) -------------------------------------------
int mark(Ctx& g, cudaStream_t st, const char* name) {
  if (!g.profiling) return 0;
  size_t idx = g.marks.size();
  if (idx >= g.ev_pool.size()) {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    g.ev_pool.push_back(e);
  }
  CU(cudaEventRecord(g.ev_pool[idx], st));
  g.marks.push_back({name, g.ev_pool[idx]});
  return 0;
}
int finish_marks(Ctx& g, cudaStream_t st) {
  if (!g.profiling || g.marks.empty()) return 0;
  CU(cudaStreamSynchronize(st));
  g.stage_ms.clear();
  for (size_t i = 1; i < g.marks.size(); i++) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, g.marks[i - 1].ev, g.marks[i].ev));
    g.stage_ms[g.marks[i].name] += ms;
  }
  float tot = 0;
  CU(cudaEventElapsedTime(&tot, g.marks.front().ev, g.marks.back().ev));
  g.stage_ms["total"] = tot;
  g.marks.clear();
  return 0;
}

// ---- NUMA topology from sysfs (placement of pinned buffers and of the worker threads; see tb200_host_alloc_sharded) ------
int numa_node_of(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return -1;
  for (char* c = bus; *c; c++) *c = (char)tolower(*c);
  std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}
// CPUs of a node ("0-31,64-95") intersected with the CPUs this process may run on; empty when unknown
std::vector<int> cpus_of_node(int node) {
  std::vector<int> out;
  if (node < 0) return out;
  char path[96];
  snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
  FILE* f = fopen(path, "r");
  if (!f) return out;
  char buf[4096] = {0};
  if (!fgets(buf, sizeof buf, f)) buf[0] = 0;
  fclose(f);
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return out;
  for (char* p = buf; *p && *p != '\n';) {
    char* end;
    long a = strtol(p, &end, 10), b = a;
    if (end == p) break;
    p = end;
    if (*p == '-') {
      b = strtol(p + 1, &end, 10);
      p = end;
    }
    for (long c = a; c <= b && c < CPU_SETSIZE; c++)
      if (CPU_ISSET((int)c, &allowed)) out.push_back((int)c);
    if (*p == ',') p++;
  }
  return out;
}
bool bind_this_thread(const std::vector<int>& cpus) {
  if (cpus.empty()) return false;
  cpu_set_t set;
  CPU_ZERO(&set);
  for (int c : cpus) CPU_SET(c, &set);
  return sched_setaffinity(0, sizeof set, &set) == 0;
}
// ---- worker threads ------------------------------------------------------------------------------------------------------
static void worker_main(Ctx* c) {
  cudaSetDevice(c->device);
  bind_this_thread(cpus_of_node(numa_node_of(c->device)));  // copies are issued from the GPU's own socket
  std::unique_lock<std::mutex> lk(c->wmu);
  for (;;) {
    c->wcv.wait(lk, [c] { return c->job_ready || c->quit; });
    if (c->quit) return;
    c->job_ready = false;
    std::function<int()> fn = std::move(c->job);
    lk.unlock();
    g_err.clear();
    int rc = fn();
    lk.lock();
    c->job_rc = rc;
    c->job_err = rc ? g_err : std::string();
    c->job_done = true;
    c->wcv.notify_all();
  }
}

int for_each_device(const std::function<int(Ctx&)>& fn, int first, int count) {
  const int n = count < 0 ? ndev() - first : count;
  // slots other than `first` run on their workers; `first` runs here (the caller's thread, primary device current)
  for (int i = 1; i < n; i++) {
    Ctx* c = E.devs[first + i].get();
    if (c->slot == 0) continue;
    std::lock_guard<std::mutex> lk(c->wmu);
    c->job = [c, &fn] { return fn(*c); };
    c->job_done = false;
    c->job_ready = true;
    c->wcv.notify_all();
  }
  int rc = 0;
  std::string err;
  {
    Ctx& c0 = *E.devs[first];
    if (c0.slot == 0) {
      rc = fn(c0);
      if (rc) err = g_err;
    } else {  // a sub-range that does not start at the primary: run its head on its worker too
      {
        std::lock_guard<std::mutex> lk(c0.wmu);
        c0.job = [&c0, &fn] { return fn(c0); };
        c0.job_done = false;
        c0.job_ready = true;
        c0.wcv.notify_all();
      }
      std::unique_lock<std::mutex> lk(c0.wmu);
      c0.wcv.wait(lk, [&c0] { return c0.job_done; });
      rc = c0.job_rc;
      err = c0.job_err;
    }
  }
  for (int i = 1; i < n; i++) {
    Ctx* c = E.devs[first + i].get();
    std::unique_lock<std::mutex> lk(c->wmu);
    c->wcv.wait(lk, [c] { return c->job_done; });
    if (c->job_rc && !rc) {
      rc = c->job_rc;
      err = c->job_err;
    }
  }
  if (rc) g_err = err;
  return rc;
}

// ---- NCCL (single-process clique; loaded at run time so the library itself has no link-time dependency on it) -------------
struct Nccl {
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
} nccl;

static int nccl_load() {
  if (E.nccl_lib) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(TB200_E_STATE, "multi-GPU needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
#define SYM(field, name)                                                                \
  *(void**)(&nccl.field) = dlsym(h, name);                                               \
  if (!nccl.field) return fail(TB200_E_STATE, "NCCL symbol %s not found", name)
  SYM(CommInitAll, "ncclCommInitAll");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllGather, "ncclAllGather");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  E.nccl_lib = h;
  return 0;
}
#define NCCLCHECK(expr)                                                                                   \
  do {                                                                                                    \
    ncclResult_t r__ = (expr);                                                                            \
    if (r__ != ncclSuccess) return fail(1000 + (int)r__, "%s failed: %s", #expr, nccl.GetErrorString(r__)); \
  } while (0)

static int nccl_init_clique() {
  const int n = ndev();
  if (n < 2) return 0;
  if (E.devs[0]->nccl_comm) {  // the clique changes when devices are added: rebuild
    for (auto& d : E.devs)
      if (d->nccl_comm) {
        nccl.CommDestroy((ncclComm_t)d->nccl_comm);
        d->nccl_comm = nullptr;
      }
  }
  if (int rc = nccl_load()) return rc;
  std::vector<int> ids(n);
  std::vector<ncclComm_t> comms(n);
  for (int i = 0; i < n; i++) ids[i] = E.devs[i]->device;
  NCCLCHECK(nccl.CommInitAll(comms.data(), n, ids.data()));
  for (int i = 0; i < n; i++) E.devs[i]->nccl_comm = comms[i];
  CU(cudaSetDevice(E.devs[0]->device));
  return 0;
}

int all_gather(const std::vector<void*>& d_send, const std::vector<void*>& d_recv, size_t bytes) {
  const int n = ndev();
  if (n < 2) return fail(TB200_E_STATE, "all_gather needs at least two devices");
  NCCLCHECK(nccl.GroupStart());
  for (int i = 0; i < n; i++) {
    ncclResult_t r = nccl.AllGather(d_send[i], d_recv[i], bytes, ncclUint8, (ncclComm_t)E.devs[i]->nccl_comm,
                                    E.devs[i]->stream);
    if (r != ncclSuccess) {
      nccl.GroupEnd();
      return fail(1000 + (int)r, "ncclAllGather failed: %s", nccl.GetErrorString(r));
    }
  }
  NCCLCHECK(nccl.GroupEnd());
  g_launches += n;  // one NCCL kernel per device
  CU(cudaSetDevice(E.devs[0]->device));
  return 0;
}

// ---- context lifecycle ---------------------------------------------------------------------------------------------------
static int ctx_create(Ctx& g, int device) {
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  g.device = device;
  g.sms = prop.multiProcessorCount;
  {  // keep freed staging buffers in the pool: with the default threshold (0) every synchronisation returns them to
     // the OS and the next host-facing call pays hundreds of ms to map gigabytes again
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = ~0ull;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&g.ev_points, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&g.stream2, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&g.ev_join, cudaEventDisableTiming));
  {
    int lo = 0, hi = 0;  // numerically lowest = greatest priority
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int i = 0; i < 2; i++) CU(cudaStreamCreateWithPriority(&g.split_stream[i], cudaStreamNonBlocking, hi));
    for (int i = 0; i < 4; i++) CU(cudaEventCreateWithFlags(&g.ev_split[i], cudaEventDisableTiming));
  }
  CU(cudaMalloc((void**)&g.d_result, Ctx::RES_BYTES));
  CU(cudaMallocHost((void**)&g.h_result, Ctx::RES_BYTES));
  g.profiling = E.profiling;
  return 0;
}
static void ctx_destroy(Ctx& g) {
  if (g.worker.joinable()) {
    {
      std::lock_guard<std::mutex> lk(g.wmu);
      g.quit = true;
      g.wcv.notify_all();
    }
    g.worker.join();
  }
  cudaSetDevice(g.device);
  cudaDeviceSynchronize();
  g.arena.destroy();
  g.arena2.destroy();
  for (int i = 0; i < 2; i++) {
    g.split_arena[i].destroy();
    if (g.split_stream[i]) cudaStreamDestroy(g.split_stream[i]);
  }
  for (int i = 0; i < 4; i++)
    if (g.ev_split[i]) cudaEventDestroy(g.ev_split[i]);
  for (int i = 0; i < Ctx::SIDE; i++) {
    if (g.side_stream[i]) {
      cudaStreamDestroy(g.side_stream[i]);
      cudaEventDestroy(g.side_done[i]);
    }
    g.side_arena[i].destroy();
  }
  if (g.pair_stream) {
    cudaStreamDestroy(g.pair_stream);
    cudaEventDestroy(g.ev_pair);
    cudaEventDestroy(g.ev_pair2);
  }
  for (auto e : g.ev_pool) cudaEventDestroy(e);
  for (auto e : g.chunk_ev) cudaEventDestroy(e);
  if (g.d_result) cudaFree(g.d_result);
  if (g.h_result) cudaFreeHost(g.h_result);
  if (g.stream2) cudaStreamDestroy(g.stream2);
  if (g.ev_join) cudaEventDestroy(g.ev_join);
  if (g.stream) cudaStreamDestroy(g.stream);
  if (g.copy_stream) cudaStreamDestroy(g.copy_stream);
  if (g.ev_points) cudaEventDestroy(g.ev_points);
}

static int init_devices_locked(const int* devices, int n) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(e != cudaSuccess ? (int)e : TB200_E_STATE,
                "no CUDA device available (%s): testudo_b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (n < 1 || n > 64 || !devices) return fail(TB200_E_ARG, "bad device list (n = %d)", n);
  std::vector<int> want(devices, devices + n);
  if (want[0] < 0) CU(cudaGetDevice(&want[0]));
  for (int i = 0; i < n; i++) {
    if (want[i] < 0 || want[i] >= count) return fail(TB200_E_ARG, "device %d does not exist (%d devices)", want[i], count);
    for (int j = 0; j < i; j++)
      if (want[j] == want[i]) return fail(TB200_E_ARG, "device %d listed twice", want[i]);
  }
  if (E.ready) {
    // idempotent; a later call may ADD devices behind the same primary (tb200_init(d) followed by
    // tb200_init_devices({d, ...}))
    if (want[0] != E.devs[0]->device)
      return fail(TB200_E_STATE, "already initialised on device %d: call tb200_shutdown before choosing another primary",
                  E.devs[0]->device);
    for (int i = 0; i < (int)E.devs.size() && i < n; i++)
      if (E.devs[i]->device != want[i])
        return fail(TB200_E_STATE, "device list differs from the initialised one at position %d", i);
    if (n <= (int)E.devs.size()) return 0;
  } else {
    if (const char* m = getenv("TB200_HOST_CHUNK_MIN")) E.host_chunk_min = (size_t)atoll(m);  // tuning aids
    if (const char* m = getenv("TB200_SHARD_MIN")) E.shard_min = (size_t)atoll(m);
    if (const char* m = getenv("TB200_PASS_ENTRIES_MAX")) E.pass_entries_max = std::max<uint64_t>(1024, strtoull(m, nullptr, 10));
    if (const char* m = getenv("TB200_ACC_MODE")) {
      int v = atoi(m);
      E.acc_mode = (v == 3 || v == 4) ? v : 0;
    }
  }
  const size_t have = E.devs.size();
  for (int i = (int)have; i < n; i++) {
    std::unique_ptr<Ctx> c(new Ctx());
    c->slot = i;
    int rc = ctx_create(*c, want[i]);
    if (rc) {
      ctx_destroy(*c);
      if (!E.ready) {
        for (auto& d : E.devs) ctx_destroy(*d);
        E.devs.clear();
      }
      return rc;
    }
    if (i > 0) c->worker = std::thread(worker_main, c.get());
    E.devs.push_back(std::move(c));
  }
  CU(cudaSetDevice(E.devs[0]->device));
  E.ready = true;
  if (E.devs.size() > 1) {
    int rc = nccl_init_clique();
    if (rc) return rc;
  }
  return 0;
}

}  // namespace tbe

using namespace tbe;

// split [0, n) into ndev contiguous ranges (the first n % ndev ranges are one longer)
static inline void shard_range(size_t n, int nd, int slot, size_t* lo, size_t* hi) {
  const size_t q = n / nd, r = n % nd;
  *lo = slot * q + std::min<size_t>(slot, r);
  *hi = *lo + q + ((size_t)slot < r ? 1 : 0);
}

extern "C" {

int tb200_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (E.ready && (device < 0 || device == E.devs[0]->device)) return 0;
  return init_devices_locked(&device, 1);
}
int tb200_init_devices(const int* devices, int ndevices) {
  std::lock_guard<std::mutex> lk(g_mu);
  return init_devices_locked(devices, ndevices);
}
int tb200_device_count(void) { return E.ready ? ndev() : 0; }

void tb200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!E.ready) return;
  for (auto& d : E.devs)
    if (d->nccl_comm) {
      nccl.CommDestroy((ncclComm_t)d->nccl_comm);
      d->nccl_comm = nullptr;
    }
  for (auto it = E.devs.rbegin(); it != E.devs.rend(); ++it) ctx_destroy(**it);
  E.devs.clear();
  E.ready = false;
}

const char* tb200_last_error(void) { return g_err.c_str(); }
uint64_t tb200_launch_count(void) { return g_launches.load(); }
void tb200_reset_launch_count(void) { g_launches = 0; }
void tb200_set_profiling(int enabled) {
  E.profiling = enabled != 0;
  for (auto& d : E.devs) d->profiling = E.profiling;
}
double tb200_stage_ms(const char* stage) {
  if (!E.ready) return -1.0;
  auto& m = primary().stage_ms;
  auto it = m.find(stage ? stage : "");
  return it == m.end() ? -1.0 : it->second;
}
int tb200_last_geometry(int* c, int* windows, uint64_t* entries, uint64_t* buckets, int* segment) {
  if (!E.ready) return TB200_E_STATE;
  Ctx& g = primary();
  if (c) *c = g.last_c;
  if (windows) *windows = g.last_W;
  if (entries) *entries = g.last_entries;
  if (buckets) *buckets = g.last_buckets;
  if (segment) *segment = g.last_K;
  return 0;
}
void tb200_set_pairing_coop_max(int n) { E.pairing_coop_max = n < 0 ? 0 : n; }
void tb200_set_pairing_team(int lanes) { E.pairing_team = (lanes == 96 || lanes == 64 || lanes == 33 || lanes == 32) ? lanes : 0; }
void tb200_set_window_bits(int c) { E.forced_c = (c >= 3 && c <= 22) ? c : 0; }
void tb200_set_accumulate_mode(int mode) { E.acc_mode = (mode == 3 || mode == 4) ? mode : 0; }
void tb200_set_pass_entries_max(uint64_t entries) {
  E.pass_entries_max = entries ? std::min<uint64_t>(std::max<uint64_t>(entries, 1024), (1ull << 32) - 1024) : (1ull << 32) - 1024;
}
void tb200_set_shard_min(size_t units) { E.shard_min = units ? units : (size_t(1) << 18); }
void tb200_set_commit_pipeline(int enabled) { E.commit_pipeline = enabled ? 1 : 0; }
void tb200_set_msm_overlap(int enabled) { E.msm_overlap = enabled ? 1 : 0; }
int tb200_set_host_upload(int pace, const int* sixteenths, int count) {
  if (count < 0 || count > 16 || (count && !sixteenths)) return fail(TB200_E_ARG, "at most 16 chunks");
  int sum = 0;
  for (int i = 0; i < count; i++) {
    if (sixteenths[i] < 1) return fail(TB200_E_ARG, "chunk sizes are positive sixteenths");
    sum += sixteenths[i];
  }
  if (count && sum != 16) return fail(TB200_E_ARG, "chunk sizes must add up to 16 sixteenths (got %d)", sum);
  E.host_upload_pace = pace ? 1 : 0;
  E.host_chunk_count = count;
  for (int i = 0; i < count; i++) E.host_chunk_frac[i] = sixteenths[i];
  return 0;
}
void tb200_set_small_msm_max(int n) { E.small_msm_max = n < 0 ? 1024 : std::min(n, 1024); }

// ---- device / pinned host buffers for hosts without a CUDA runtime of their own ------------------------------------------
int tb200_dev_alloc(size_t bytes, void** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || bytes == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  CU(cudaMalloc(out, bytes));
  return 0;
}
int tb200_dev_free(void* p) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(primary().device));
  CU(cudaStreamSynchronize(primary().stream));
  CU(cudaFree(p));
  return 0;
}
int tb200_dev_upload(void* d_dst, const void* h_src, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_dst || !h_src) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(primary().device));
  CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, primary().stream));
  CU(cudaStreamSynchronize(primary().stream));
  return 0;
}
int tb200_dev_download(void* h_dst, const void* d_src, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h_dst || !d_src) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(primary().device));
  CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, primary().stream));
  CU(cudaStreamSynchronize(primary().stream));
  return 0;
}
int tb200_stream_sync(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(primary().device));
  CU(cudaStreamSynchronize(primary().stream));
  return 0;
}
// ---- NUMA placement of pinned host buffers ----------------------------------------------------------------------------
// Eight GPUs pulling 2 GiB each from pinned pages that all sit on one socket are limited by that socket's memory
// controllers and the inter-socket link, not by PCIe (round 1: end-to-end efficiency 0.72 at 8 GPUs). Pinned pages are
// physically placed when they are first touched, so the shard a GPU will read is touched by a thread bound to the CPUs
// of that GPU's NUMA node and only then page-locked. Everything degrades to a plain pinned allocation when sysfs
// exposes no topology (numa_node = -1, single node, restricted cpuset).
namespace {
std::map<void*, size_t> g_mapped;  // buffers handed out by tb200_host_alloc_sharded / _near: base -> mapped bytes

// first touch of [p, p + bytes) from threads bound to `node` (unbound if the node is unknown)
void touch_on_node(char* p, size_t bytes, int node) {
  const std::vector<int> cpus = cpus_of_node(node);
  const int nt = (int)std::max<size_t>(1, std::min<size_t>(8, bytes >> 26));
  std::vector<std::thread> ts;
  for (int t = 0; t < nt; t++)
    ts.emplace_back([=, &cpus] {
      bind_this_thread(cpus);
      const size_t a = bytes * t / nt, b = bytes * (t + 1) / nt;
      for (size_t o = a; o < b; o += 4096) p[o] = 0;
    });
  for (auto& t : ts) t.join();
}
// shares[i] bytes placed next to device slot dev_slot[i], one contiguous page-locked mapping
int host_alloc_placed(const std::vector<size_t>& shares, const std::vector<int>& dev_slot, void** out) {
  size_t bytes = 0;
  for (size_t s : shares) bytes += s;
  const size_t mapped = (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
  void* p = mmap(nullptr, mapped, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) return fail(TB200_E_LIMIT, "mmap of %zu bytes failed", mapped);
  madvise(p, mapped, MADV_HUGEPAGE);
  size_t off = 0;
  for (size_t i = 0; i < shares.size(); i++) {
    // page-granular boundaries: the page that straddles two shares goes with the first
    const size_t end = (i + 1 == shares.size()) ? mapped : std::min(mapped, (off + shares[i] + 4095) & ~size_t(4095));
    const size_t start = (off + 4095) & ~size_t(4095);
    if (end > start) touch_on_node((char*)p + start, end - start, numa_node_of(E.devs[dev_slot[i]]->device));
    off += shares[i];
  }
  cudaError_t e = cudaHostRegister(p, mapped, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    munmap(p, mapped);
    return fail((int)e, "cudaHostRegister of %zu bytes failed: %s", mapped, cudaGetErrorString(e));
  }
  g_mapped[p] = mapped;
  *out = p;
  return 0;
}
}  // namespace

int tb200_host_alloc(size_t bytes, void** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || bytes == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  CU(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return 0;
}
int tb200_host_alloc_near(size_t bytes, int device_slot, void** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || bytes == 0 || device_slot < 0 || device_slot >= ndev()) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  return host_alloc_placed({bytes}, {device_slot}, out);
}
int tb200_host_alloc_sharded(size_t units, size_t unit_bytes, void** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || units == 0 || unit_bytes == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  std::vector<size_t> shares;
  std::vector<int> slots;
  for (int i = 0; i < ndev(); i++) {  // the same split the sharded entry points use
    size_t lo, hi;
    shard_range(units, ndev(), i, &lo, &hi);
    shares.push_back((hi - lo) * unit_bytes);
    slots.push_back(i);
  }
  return host_alloc_placed(shares, slots, out);
}
int tb200_device_numa_node(int device_slot) {
  if (!E.ready || device_slot < 0 || device_slot >= ndev()) return -1;
  return numa_node_of(E.devs[device_slot]->device);
}
int tb200_host_free(void* p) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  auto it = g_mapped.find(p);
  if (it != g_mapped.end()) {
    cudaError_t e = cudaHostUnregister(p);
    munmap(p, it->second);
    g_mapped.erase(it);
    if (e != cudaSuccess) return fail((int)e, "cudaHostUnregister failed: %s", cudaGetErrorString(e));
    return 0;
  }
  CU(cudaFreeHost(p));
  return 0;
}
int tb200_host_register(void* p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!p || bytes == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  CU(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return 0;
}
int tb200_host_unregister(void* p) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaHostUnregister(p));
  return 0;
}

// ---- single variable-base MSM from host buffers, sharded by point range ---------------------------------------------------
static void free_all(Ctx& g, std::vector<void*>& v) {
  for (void* p : v) cudaFreeAsync(p, g.stream);
  v.clear();
}

int tb200_msm_g1(const uint64_t* bases_xy, const uint64_t* scalars, size_t n, unsigned flags, uint64_t out_xy[12]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out_xy || (n && (!bases_xy || !scalars))) return fail(TB200_E_ARG, "null pointer");
  Ctx& g0 = primary();
  CU(cudaSetDevice(g0.device));
  const int nd = (ndev() > 1 && n >= (size_t)ndev() * E.shard_min) ? ndev() : 1;
  std::vector<std::vector<void*>> to_free(nd);
  int rc;
  if (nd == 1) {
    rc = msm_host_enqueue(g0, bases_xy, scalars, n, flags, to_free[0]);
  } else {
    // contiguous point ranges, one partial point per GPU (nothing replicated: SURVEY.md 8e)
    rc = for_each_device([&](Ctx& g) {
      size_t lo, hi;
      shard_range(n, nd, g.slot, &lo, &hi);
      return msm_host_enqueue(g, bases_xy + 12 * lo, scalars + 4 * lo, hi - lo, flags, to_free[g.slot], nd);
    });
    if (rc == 0) {
      std::vector<void*> send(nd), recv(nd);
      for (int i = 0; i < nd; i++) {
        send[i] = E.devs[i]->d_result;
        recv[i] = E.devs[i]->d_result + Ctx::RES_GATHER;
      }
      rc = all_gather(send, recv, 96);
    }
    if (rc == 0) rc = g1_sum_dev(g0, g0.d_result + Ctx::RES_GATHER, (size_t)nd, g0.d_result, g0.stream);
  }
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(g0.h_result, g0.d_result, 96, cudaMemcpyDeviceToHost, g0.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g0.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
    else memcpy(out_xy, g0.h_result, 96);
  }
  // the host buffers are only borrowed for the duration of the call: every device must have finished reading them
  for (int i = 0; i < nd; i++) {
    Ctx& g = *E.devs[i];
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.copy_stream);
    cudaStreamSynchronize(g.stream);
    free_all(g, to_free[i]);
  }
  cudaSetDevice(g0.device);
  if (rc == 0) rc = finish_marks(g0, g0.stream);
  return rc;
}

// Same with the inputs RESIDENT on the GPUs: device i holds n[i] points / scalars (any split; n[i] may be 0). The
// partial points are all-gathered and summed as above; the result comes back to the host (96 bytes).
int tb200_msm_g1_sharded_dev(const void* const* d_bases_xy, const void* const* d_scalars, const size_t* n, unsigned flags,
                             uint64_t out_xy[12]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out_xy || !d_bases_xy || !d_scalars || !n) return fail(TB200_E_ARG, "null pointer");
  Ctx& g0 = primary();
  CU(cudaSetDevice(g0.device));
  const int nd = ndev();
  for (int i = 0; i < nd; i++)
    if (n[i] && (!d_bases_xy[i] || !d_scalars[i])) return fail(TB200_E_ARG, "null device pointer for device slot %d", i);
  int rc = for_each_device([&](Ctx& g) {
    CU(cudaSetDevice(g.device));
    const void* b = n[g.slot] ? d_bases_xy[g.slot] : (const void*)g.d_result;
    const void* s = n[g.slot] ? d_scalars[g.slot] : (const void*)g.d_result;
    return msm_dev(g, b, s, n[g.slot], flags, g.d_result, g.stream, nullptr, nullptr, false);
  });
  if (rc == 0 && nd > 1) {
    std::vector<void*> send(nd), recv(nd);
    for (int i = 0; i < nd; i++) {
      send[i] = E.devs[i]->d_result;
      recv[i] = E.devs[i]->d_result + Ctx::RES_GATHER;
    }
    rc = all_gather(send, recv, 96);
    if (rc == 0) rc = g1_sum_dev(g0, g0.d_result + Ctx::RES_GATHER, (size_t)nd, g0.d_result, g0.stream);
  }
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(g0.h_result, g0.d_result, 96, cudaMemcpyDeviceToHost, g0.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g0.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
    else memcpy(out_xy, g0.h_result, 96);
  }
  for (int i = 1; i < nd; i++) {
    cudaSetDevice(E.devs[i]->device);
    cudaStreamSynchronize(E.devs[i]->stream);
  }
  cudaSetDevice(g0.device);
  if (rc == 0) rc = finish_marks(g0, g0.stream);
  return rc;
}

// ---- SRS (replicated on every GPU) ---------------------------------------------------------------------------------------
static int srs_load_locked(const uint64_t* bases_xy, size_t n, const uint64_t* extra_xy, size_t n_extra, int window_bits,
                           tb200_srs_t* out) {
  if (!out || !bases_xy || n == 0 || n + n_extra >= (1u << 24) || (n_extra && !extra_xy))
    return fail(TB200_E_ARG, "bad SRS arguments (n = %zu)", n);
  tb200_srs* s = new tb200_srs();
  s->n = (uint32_t)(n + n_extra);
  s->extra = (uint32_t)n_extra;
  s->c = (window_bits >= 3 && window_bits <= 16) ? window_bits : pick_c_batch(s->n);
  s->W = (253 + s->c) / s->c;  // num_windows(c): one spare bit for the signed-digit carry
  s->table.assign(ndev(), nullptr);
  std::vector<uint64_t> joined;
  const uint64_t* src = bases_xy;
  if (n_extra) {
    joined.resize((size_t)s->n * 12);
    memcpy(joined.data(), bases_xy, n * 96);
    memcpy(joined.data() + n * 12, extra_xy, n_extra * 96);
    src = joined.data();
  }
  int rc = for_each_device([&](Ctx& g) { return srs_build_table(g, src, s->n, s->c, s->W, &s->table[g.slot]); });
  cudaSetDevice(primary().device);
  if (rc) {
    for (int i = 0; i < ndev(); i++)
      if (s->table[i]) {
        cudaSetDevice(E.devs[i]->device);
        cudaFree(s->table[i]);
      }
    cudaSetDevice(primary().device);
    delete s;
    return rc;
  }
  *out = s;
  return 0;
}
int tb200_srs_load(const uint64_t* bases_xy, size_t n, int window_bits, tb200_srs_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(primary().device));
  return srs_load_locked(bases_xy, n, nullptr, 0, window_bits, out);
}
int tb200_srs_load_blinded(const uint64_t* bases_xy, size_t n, const uint64_t h_xy[12], int window_bits, tb200_srs_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h_xy) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(primary().device));
  return srs_load_locked(bases_xy, n, h_xy, 1, window_bits, out);
}
int tb200_srs_free(tb200_srs_t srs) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!srs) return fail(TB200_E_ARG, "null SRS handle");
  if (E.ready) {
    for (int i = 0; i < ndev() && i < (int)srs->table.size(); i++) {
      cudaSetDevice(E.devs[i]->device);
      cudaStreamSynchronize(E.devs[i]->stream);
      if (srs->table[i]) cudaFree(srs->table[i]);
    }
    cudaSetDevice(primary().device);
  }
  delete srs;
  return 0;
}
size_t tb200_srs_size(tb200_srs_t srs) { return srs ? srs->n - srs->extra : 0; }

// ---- row commitments from host buffers, sharded by row range ------------------------------------------------------------------
namespace {

// Where the scalars of a batch live on the host: a strided view of one matrix (tb200_msm_g1_batch) or one heap buffer per
// row (tb200_msm_g1_batch_ptrs: what `Polynomial::commit` holds in self.polys, src/sqrt_pst.rs:48-62). `blinds`, if set,
// is one extra scalar per row for the blinding column appended to the SRS (Hyrax, src/commitments.rs:80-86).
struct RowSource {
  const uint64_t* base = nullptr;
  long long rs = 0, cs = 0;
  const uint64_t* const* ptrs = nullptr;
  const uint64_t* blinds = nullptr;
  size_t rows = 0, cols = 0;  // cols = scalars per row WITHOUT the blind
};

// rows [a, b) of one device's share, as they are laid out in its staging buffer
struct RowChunk {
  size_t a, b;
};

// chunks of growing size: the first upload is the only one the GPU waits for, every later one hides behind the
// previous chunk's compute (57 us of integer work vs 5-25 us of transfer per 8192-column row)
std::vector<RowChunk> row_chunks(size_t rows, int sms, bool small_tail) {
  std::vector<RowChunk> v;
  const size_t min_chunk = (size_t)std::max(sms, 1);  // >= one CTA per SM for the per-row sort kernel
  if (rows < 4 * min_chunk) {
    v.push_back({0, rows});
    return v;
  }
  // The LAST chunks are small as well: whatever follows the row MSMs of a chunk (its Miller loops in
  // tb200_sqrt_pst_commit, the download of its rows) overlaps the compute of the chunks behind it -- except for the last
  // one. Miller kernels that share the SMs with the accumulation run ~3x slower than alone (their dependent IMAD chains
  // queue behind a saturated integer pipe), so the chunk BEFORE the last is small too: a 4460-pair Miller kernel behind
  // a 148-row tail left 10+ ms exposed (measured: t cost 22 ms instead of 12).
  // Measured (profiles/r02_summary.md): the pipelined pairing LOSES 9-12 ms at 4096-8192 rows per GPU -- every Miller
  // CTA that becomes resident displaces an accumulation CTA (4 x 128 threads x 128 registers fill the register file) --
  // so it is opt-in (tb200_set_commit_pipeline) and the small tail chunks exist only then.
  const size_t tail = small_tail ? min_chunk : 0;
  size_t pre = small_tail ? std::max(min_chunk, rows / 16) : 0;
  if (rows < tail + pre + 2 * min_chunk) pre = 0;
  const size_t body = rows - tail - pre;
  size_t len = std::max(min_chunk, rows / 16), a = 0;
  while (a < body) {
    size_t b = std::min(body, a + len);
    if (body - b < min_chunk) b = body;
    v.push_back({a, b});
    a = b;
    len *= 2;
  }
  if (pre) v.push_back({body, body + pre});
  if (tail) v.push_back({body + pre, rows});
  return v;
}

// Pairing stage of tb200_sqrt_pst_commit, pipelined: once the rows of a chunk exist, their Miller loops against the
// matching slice of the G2 key run on the context's pair stream NEXT TO the row MSMs of the following chunk (the Miller
// kernels are latency-bound and leave the integer pipe to the accumulation); one partial Miller product per chunk.
struct PairPipe {
  const uint4* d_h = nullptr;    // this share's slice of h_vec on the device (rows [r0, r1)), nullptr: no pairing stage
  cudaEvent_t h_ready = nullptr; // recorded on the copy stream once d_h has arrived
  uint4* d_parts = nullptr;      // one 576-byte partial product per chunk
  int nparts = 0;
};
int pair_stream_ready(Ctx& g) {
  if (g.pair_stream) return 0;
  CU(cudaStreamCreateWithFlags(&g.pair_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&g.ev_pair, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&g.ev_pair2, cudaEventDisableTiming));
  return 0;
}
int pair_chunk(Ctx& g, PairPipe* pp, const uint4* d_rows, size_t a, size_t nr) {
  if (!pp || !pp->d_h || nr == 0) return 0;
  if (int rc = pair_stream_ready(g)) return rc;
  CU(cudaEventRecord(g.ev_pair, g.stream));               // the rows of this chunk exist
  CU(cudaStreamWaitEvent(g.pair_stream, g.ev_pair, 0));
  if (pp->nparts == 0) CU(cudaStreamWaitEvent(g.pair_stream, pp->h_ready, 0));
  const bool prof = g.profiling;
  g.profiling = false;                                    // stage marks belong to the main stream
  int rc = pairing_products(g, d_rows + 6 * a, pp->d_h + 12 * a, (uint32_t)nr, 0, 1, pp->d_parts + 36 * (size_t)pp->nparts,
                            g.pair_stream, nullptr, nullptr, false);
  g.profiling = prof;
  pp->nparts++;
  return rc;
}

int batch_share_enqueue(Ctx& g, const tb200_srs* srs, const RowSource& src, size_t r0, size_t r1, unsigned flags,
                        uint4** d_out_p, std::vector<void*>& to_free, PairPipe* pp = nullptr) {
  CU(cudaSetDevice(g.device));
  const size_t my_rows = r1 - r0, cols = src.cols, colsT = cols + (src.blinds ? 1 : 0);
  uint4* d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, std::max<size_t>(my_rows, 1) * 96, g.stream));
  to_free.push_back(d_o);
  *d_out_p = d_o;
  if (my_rows == 0) return 0;
  if (pp && pp->d_h) {
    CU(cudaMallocAsync((void**)&pp->d_parts, (row_chunks(my_rows, g.sms, true).size() + 1) * 576, g.stream));
    to_free.push_back(pp->d_parts);
  }
  if (colsT == 0) {
    int rc = batch_dev(g, srs->table[g.slot], srs->c, srs->W, srs->n, (const uint32_t*)d_o, my_rows, 0, 0, 1, flags, d_o,
                       g.stream);
    return rc ? rc : pair_chunk(g, pp, d_o, 0, my_rows);
  }
  const bool col_major = !src.ptrs && src.rs == 1 && src.rows > 1;  // rows are the unit-stride dimension
  const bool simple = src.ptrs || col_major || (src.cs == 1);
  if (!simple) {
    // arbitrary strides: upload the whole extent of this share and let the kernels apply the strides
    const size_t extent = (my_rows - 1) * (size_t)src.rs + (cols - 1) * (size_t)src.cs + 1;
    if (src.blinds) return fail(TB200_E_ARG, "blinds need unit row or column stride");
    uint4* d_s = nullptr;
    CU(cudaMallocAsync((void**)&d_s, extent * 32, g.stream));
    to_free.push_back(d_s);
    CU(cudaMemcpyAsync(d_s, src.base + 4 * (long long)r0 * src.rs, extent * 32, cudaMemcpyHostToDevice, g.stream));
    int rc = batch_dev(g, srs->table[g.slot], srs->c, srs->W, srs->n, (const uint32_t*)d_s, my_rows, cols, src.rs, src.cs,
                       flags, d_o, g.stream);
    return rc ? rc : pair_chunk(g, pp, d_o, 0, my_rows);
  }
  uint4* d_s = nullptr;
  CU(cudaMallocAsync((void**)&d_s, my_rows * colsT * 32, g.stream));
  to_free.push_back(d_s);
  CU(cudaEventRecord(g.ev_points, g.stream));  // the allocation exists
  CU(cudaStreamWaitEvent(g.copy_stream, g.ev_points, 0));
  const std::vector<RowChunk> chunks = row_chunks(my_rows, g.sms, pp && pp->d_h);
  std::vector<cudaEvent_t>& evs = g.chunk_ev;
  while (evs.size() < chunks.size()) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    evs.push_back(e);
  }
  for (size_t k = 0; k < chunks.size(); k++) {
    const size_t a = chunks[k].a, b = chunks[k].b, nr = b - a;
    char* dst = (char*)d_s + a * colsT * 32;  // every chunk is a dense block of nr * colsT scalars
    if (src.ptrs) {
      for (size_t i = a; i < b;) {  // runs of rows that happen to be adjacent in memory go in one copy
        const uint64_t* p = src.ptrs[r0 + i];
        if (!p) return fail(TB200_E_ARG, "row pointer %zu is null", r0 + i);
        size_t j = i + 1;
        if (!src.blinds)
          while (j < b && src.ptrs[r0 + j] == p + 4 * cols * (j - i)) j++;
        if (src.blinds) {
          CU(cudaMemcpyAsync(dst + (i - a) * colsT * 32, p, cols * 32, cudaMemcpyHostToDevice, g.copy_stream));
        } else {
          CU(cudaMemcpyAsync(dst + (i - a) * cols * 32, p, (j - i) * cols * 32, cudaMemcpyHostToDevice, g.copy_stream));
        }
        i = j;
      }
      if (src.blinds)
        CU(cudaMemcpy2DAsync(dst + cols * 32, colsT * 32, src.blinds + 4 * (r0 + a), 32, 32, nr, cudaMemcpyHostToDevice,
                             g.copy_stream));
    } else if (col_major) {
      // source element (row i, col j) at base + (i + j * cs); device block [colsT][nr]
      CU(cudaMemcpy2DAsync(dst, nr * 32, src.base + 4 * (r0 + a), (size_t)src.cs * 32, nr * 32, cols, cudaMemcpyHostToDevice,
                           g.copy_stream));
      if (src.blinds)
        CU(cudaMemcpyAsync(dst + cols * nr * 32, src.blinds + 4 * (r0 + a), nr * 32, cudaMemcpyHostToDevice, g.copy_stream));
    } else {  // row-major with row stride rs >= cols (or a single row, whose stride does not matter)
      const uint64_t* from = src.base + 4 * (long long)(r0 + a) * src.rs;
      if (nr == 1 || ((size_t)src.rs == cols && !src.blinds))
        CU(cudaMemcpyAsync(dst, from, nr * cols * 32, cudaMemcpyHostToDevice, g.copy_stream));
      else if ((size_t)src.rs < cols)
        return fail(TB200_E_ARG, "row stride %lld is smaller than cols %zu", src.rs, cols);
      else
        CU(cudaMemcpy2DAsync(dst, colsT * 32, from, (size_t)src.rs * 32, cols * 32, nr, cudaMemcpyHostToDevice,
                             g.copy_stream));
      if (src.blinds)
        CU(cudaMemcpy2DAsync(dst + cols * 32, colsT * 32, src.blinds + 4 * (r0 + a), 32, 32, nr, cudaMemcpyHostToDevice,
                             g.copy_stream));
    }
    CU(cudaEventRecord(evs[k], g.copy_stream));
    CU(cudaStreamWaitEvent(g.stream, evs[k], 0));
    const long long rs = col_major ? 1 : (long long)colsT, cs = col_major ? (long long)nr : 1;
    int rc = batch_dev(g, srs->table[g.slot], srs->c, srs->W, srs->n, (const uint32_t*)dst, nr, colsT, rs, cs, flags,
                       d_o + 6 * a, g.stream);
    if (rc) return rc;
    rc = pair_chunk(g, pp, d_o, a, nr);
    if (rc) return rc;
  }
  return 0;
}

struct ShardPlan {
  int nd;
  std::vector<size_t> lo, hi;
};
ShardPlan plan_rows(size_t rows, size_t cols) {
  ShardPlan p;
  // sharding pays once every GPU gets a few hundred thousand scalar products; tiny batches stay on the primary
  p.nd = (ndev() > 1 && rows >= (size_t)ndev() && rows * std::max<size_t>(cols, 1) >= (size_t)ndev() * E.shard_min) ? ndev() : 1;
  p.lo.resize(p.nd);
  p.hi.resize(p.nd);
  for (int i = 0; i < p.nd; i++) shard_range(rows, p.nd, i, &p.lo[i], &p.hi[i]);
  return p;
}

// rows -> out_xy (host). If h_vec is given also t = prod_i e(C_i, h_i) (src/sqrt_pst.rs:131-144) -> out_t.
int commit_rows_locked(tb200_srs_t srs, const RowSource& src, unsigned flags, uint64_t* out_xy, const uint64_t* h_vec,
                       uint64_t* out_t) {
  Ctx& g0 = primary();
  CU(cudaSetDevice(g0.device));
  static const bool trace = getenv("TB200_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  const size_t rows = src.rows;
  const ShardPlan sp = plan_rows(rows, src.cols);
  const int nd = sp.nd;
  std::vector<std::vector<void*>> to_free(nd);
  std::vector<uint4*> d_out(nd, nullptr);
  int rc = for_each_device(
      [&](Ctx& g) {
        const size_t lo = sp.lo[g.slot], hi = sp.hi[g.slot];
        uint4* d_h = nullptr;
        g.marks.clear();
        if (h_vec && hi > lo) {  // this share's slice of the G2 key travels while the rows are being committed
          CU(cudaSetDevice(g.device));
          CU(cudaMallocAsync((void**)&d_h, (hi - lo) * 192, g.stream));
          to_free[g.slot].push_back(d_h);
          CU(cudaEventRecord(g.ev_join, g.stream));
          CU(cudaStreamWaitEvent(g.copy_stream, g.ev_join, 0));
          CU(cudaMemcpyAsync(d_h, h_vec + 24 * lo, (hi - lo) * 192, cudaMemcpyHostToDevice, g.copy_stream));
          CU(cudaEventRecord(g.ev_join, g.copy_stream));
        }
        PairPipe pipe;
        pipe.d_h = d_h;
        pipe.h_ready = g.ev_join;
        const bool pipelined = h_vec && E.commit_pipeline != 0;
        int r = batch_share_enqueue(g, srs, src, lo, hi, flags, &d_out[g.slot], to_free[g.slot], pipelined ? &pipe : nullptr);
        if (r) return r;
        if (trace) fprintf(stderr, "[tb200] slot %d: rows enqueued at %.2f ms\n", g.slot, since());
        if (hi > lo)
          CU(cudaMemcpyAsync(out_xy + 12 * lo, d_out[g.slot], (hi - lo) * 96, cudaMemcpyDeviceToHost, g.stream));
        if (h_vec) {
          // this share's partial Miller product = the product of its chunks' partials (no final exponentiation unless
          // this is the only GPU), 576 B at d_result + RES_PART
          if (!pipelined) {  // all Miller loops of the share behind its row MSMs
            if (d_h) CU(cudaStreamWaitEvent(g.stream, g.ev_join, 0));
            r = pairing_products(g, d_out[g.slot], d_h, (uint32_t)(hi - lo), 0, 1, g.d_result + Ctx::RES_PART, g.stream, nullptr,
                                 nullptr, nd == 1);
          } else {
            if (pipe.nparts) {
              CU(cudaEventRecord(g.ev_pair2, g.pair_stream));
              CU(cudaStreamWaitEvent(g.stream, g.ev_pair2, 0));
            }
            r = pairing_products(g, nullptr, nullptr, (uint32_t)pipe.nparts, 0, 1, g.d_result + Ctx::RES_PART, g.stream,
                                 nullptr, pipe.d_parts, nd == 1);
          }
          if (r) return r;
        }
        return 0;
      },
      0, nd);
  if (rc == 0 && h_vec) {
    if (nd > 1) {
      std::vector<void*> send(ndev()), recv(ndev());
      // devices beyond nd (none today: nd is 1 or ndev) would have to contribute the neutral element
      for (int i = 0; i < ndev(); i++) {
        send[i] = E.devs[i]->d_result + Ctx::RES_PART;
        recv[i] = E.devs[i]->d_result + Ctx::RES_GATHER;
      }
      if (rc == 0) rc = all_gather(send, recv, 576);
      if (rc == 0)
        rc = pairing_products(g0, nullptr, nullptr, (uint32_t)nd, 0, 1, g0.d_result + Ctx::RES_PART, g0.stream, nullptr,
                              g0.d_result + Ctx::RES_GATHER, true);
    }
    if (rc == 0) {
      cudaError_t e = cudaMemcpyAsync(out_t, g0.d_result + Ctx::RES_PART, 576, cudaMemcpyDeviceToHost, g0.stream);
      if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
    }
  }
  if (trace) fprintf(stderr, "[tb200] all shares enqueued at %.2f ms\n", since());
  for (int i = 0; i < nd; i++) {
    Ctx& g = *E.devs[i];
    cudaSetDevice(g.device);
    cudaError_t e = cudaStreamSynchronize(g.copy_stream);
    if (trace) fprintf(stderr, "[tb200] slot %d: uploads done at %.2f ms\n", i, since());
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (trace) fprintf(stderr, "[tb200] slot %d: compute done at %.2f ms\n", i, since());
    if (e != cudaSuccess && rc == 0) rc = fail((int)e, "device %d failed: %s", g.device, cudaGetErrorString(e));
    free_all(g, to_free[i]);
  }
  if (trace) fprintf(stderr, "[tb200] freed at %.2f ms\n", since());
  cudaSetDevice(g0.device);
  if (rc == 0) rc = finish_marks(g0, g0.stream);
  return rc;
}

int check_batch_args(tb200_srs_t srs, size_t rows, size_t cols, const void* out, bool blinds) {
  if (!srs || (rows && !out)) return fail(TB200_E_ARG, "null pointer");
  const size_t want = srs->n - srs->extra;
  if (cols != want && cols != 0)
    return fail(TB200_E_ARG, "cols (%zu) must equal the SRS size (%zu): window tables are laid out per SRS", cols, want);
  if (blinds && !srs->extra) return fail(TB200_E_ARG, "blinds need an SRS loaded with tb200_srs_load_blinded");
  if (!blinds && srs->extra && cols)
    return fail(TB200_E_ARG, "this SRS carries a blinding column: pass the blinds (tb200_msm_g1_batch_blinded)");
  if ((int)srs->table.size() < ndev())
    return fail(TB200_E_STATE, "the SRS was loaded before devices were added: reload it");
  return 0;
}

}  // namespace

int tb200_msm_g1_batch(tb200_srs_t srs, const uint64_t* scalars, size_t rows, size_t cols, ptrdiff_t row_stride,
                       ptrdiff_t col_stride, unsigned flags, uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (int rc = check_batch_args(srs, rows, cols, out_xy, false)) return rc;
  if (rows && cols && !scalars) return fail(TB200_E_ARG, "null pointer");
  if (rows == 0) return 0;
  if (row_stride < 0 || col_stride < 0) return fail(TB200_E_ARG, "negative strides are not supported");
  RowSource src;
  src.base = scalars;
  src.rs = row_stride;
  src.cs = col_stride;
  src.rows = rows;
  src.cols = cols;
  return commit_rows_locked(srs, src, flags, out_xy, nullptr, nullptr);
}

int tb200_msm_g1_batch_ptrs(tb200_srs_t srs, const uint64_t* const* row_ptrs, size_t rows, size_t cols, unsigned flags,
                            uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (int rc = check_batch_args(srs, rows, cols, out_xy, false)) return rc;
  if (rows && !row_ptrs) return fail(TB200_E_ARG, "null pointer");
  if (rows == 0) return 0;
  RowSource src;
  src.ptrs = row_ptrs;
  src.rows = rows;
  src.cols = cols;
  return commit_rows_locked(srs, src, flags, out_xy, nullptr, nullptr);
}

// Hyrax rows with blinds (src/dense_mlpoly.rs:315-329 -> PedersenCommit::commit_slice, src/commitments.rs:80-86):
// out[i] = MSM(G, row_i) + blinds[i] * h, h being the extra column of an SRS loaded with tb200_srs_load_blinded
int tb200_msm_g1_batch_blinded(tb200_srs_t srs, const uint64_t* scalars, size_t rows, size_t cols, const uint64_t* blinds,
                               unsigned flags, uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (int rc = check_batch_args(srs, rows, cols, out_xy, true)) return rc;
  if (rows && (!blinds || (cols && !scalars))) return fail(TB200_E_ARG, "null pointer");
  if (rows == 0) return 0;
  RowSource src;
  src.base = scalars;
  src.rs = (long long)cols;
  src.cs = 1;
  src.blinds = blinds;
  src.rows = rows;
  src.cols = cols;
  return commit_rows_locked(srs, src, flags, out_xy, nullptr, nullptr);
}

// `Polynomial::commit` in one call (src/sqrt_pst.rs:117-149): the row commitments AND the IPP commitment
// t = prod_i e(C_i, h_vec[i]); the rows never leave their GPU between the two stages.
int tb200_sqrt_pst_commit(tb200_srs_t srs, const uint64_t* const* row_ptrs, size_t rows, size_t cols, unsigned flags,
                          const uint64_t* h_vec, uint64_t* out_rows_xy, uint64_t out_t[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (int rc = check_batch_args(srs, rows, cols, out_rows_xy, false)) return rc;
  if (!row_ptrs || !h_vec || !out_t || rows == 0) return fail(TB200_E_ARG, "null pointer / no rows");
  if (rows >= (1u << 26)) return fail(TB200_E_LIMIT, "too many rows");
  RowSource src;
  src.ptrs = row_ptrs;
  src.rows = rows;
  src.cols = cols;
  return commit_rows_locked(srs, src, flags, out_rows_xy, h_vec, out_t);
}
int tb200_sqrt_pst_commit_strided(tb200_srs_t srs, const uint64_t* scalars, size_t rows, size_t cols, ptrdiff_t row_stride,
                                  ptrdiff_t col_stride, unsigned flags, const uint64_t* h_vec, uint64_t* out_rows_xy,
                                  uint64_t out_t[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (int rc = check_batch_args(srs, rows, cols, out_rows_xy, false)) return rc;
  if (!scalars || !h_vec || !out_t || rows == 0) return fail(TB200_E_ARG, "null pointer / no rows");
  if (row_stride < 0 || col_stride < 0) return fail(TB200_E_ARG, "negative strides are not supported");
  if (rows >= (1u << 26)) return fail(TB200_E_LIMIT, "too many rows");
  RowSource src;
  src.base = scalars;
  src.rs = row_stride;
  src.cs = col_stride;
  src.rows = rows;
  src.cols = cols;
  return commit_rows_locked(srs, src, flags, out_rows_xy, h_vec, out_t);
}

// ---- pairing products from host buffers, sharded by pair range --------------------------------------------------------------
// mode 0: full product (Miller loops, product, final exponentiation); 1: partial Miller product (no final exponentiation)
static int pairing_host_locked(const uint64_t* g1_xy, const uint64_t* g2, size_t n, uint64_t out[72], int mode) {
  Ctx& g0 = primary();
  CU(cudaSetDevice(g0.device));
  const int nd = (mode == 0 && ndev() > 1 && n >= (size_t)ndev() * 256) ? ndev() : 1;
  std::vector<std::vector<void*>> to_free(nd);
  int rc = for_each_device(
      [&](Ctx& g) {
        size_t lo, hi;
        shard_range(n, nd, g.slot, &lo, &hi);
        CU(cudaSetDevice(g.device));
        g.marks.clear();
        uint4 *d_p = nullptr, *d_q = nullptr;
        if (hi > lo) {
          CU(cudaMallocAsync((void**)&d_p, (hi - lo) * 96, g.stream));
          to_free[g.slot].push_back(d_p);
          CU(cudaMallocAsync((void**)&d_q, (hi - lo) * 192, g.stream));
          to_free[g.slot].push_back(d_q);
          CU(cudaMemcpyAsync(d_p, g1_xy + 12 * lo, (hi - lo) * 96, cudaMemcpyHostToDevice, g.stream));
          CU(cudaMemcpyAsync(d_q, g2 + 24 * lo, (hi - lo) * 192, cudaMemcpyHostToDevice, g.stream));
        }
        return pairing_products(g, d_p, d_q, (uint32_t)(hi - lo), 0, 1, g.d_result + Ctx::RES_PART, g.stream, nullptr, nullptr,
                                nd == 1 && mode == 0);
      },
      0, nd);
  if (rc == 0 && nd > 1) {
    std::vector<void*> send(ndev()), recv(ndev());
    for (int i = 0; i < ndev(); i++) {
      send[i] = E.devs[i]->d_result + Ctx::RES_PART;
      recv[i] = E.devs[i]->d_result + Ctx::RES_GATHER;
    }
    if (rc == 0) rc = all_gather(send, recv, 576);
    if (rc == 0)
      rc = pairing_products(g0, nullptr, nullptr, (uint32_t)nd, 0, 1, g0.d_result + Ctx::RES_PART, g0.stream, nullptr,
                            g0.d_result + Ctx::RES_GATHER, true);
  }
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out, g0.d_result + Ctx::RES_PART, 576, cudaMemcpyDeviceToHost, g0.stream);
    if (e != cudaSuccess) rc = fail((int)e, "pairing result copy failed: %s", cudaGetErrorString(e));
  }
  for (int i = 0; i < nd; i++) {
    Ctx& g = *E.devs[i];
    cudaSetDevice(g.device);
    cudaError_t e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess && rc == 0) rc = fail((int)e, "device %d failed: %s", g.device, cudaGetErrorString(e));
    free_all(g, to_free[i]);
  }
  cudaSetDevice(g0.device);
  if (rc == 0) rc = finish_marks(g0, g0.stream);
  return rc;
}

int tb200_multi_pairing(const uint64_t* g1_xy, const uint64_t* g2, size_t n, uint64_t out[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && (!g1_xy || !g2))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "too many pairs");
  return pairing_host_locked(g1_xy, g2, n, out, 0);
}
int tb200_miller_product(const uint64_t* g1_xy, const uint64_t* g2, size_t n, uint64_t out[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && (!g1_xy || !g2))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "too many pairs");
  return pairing_host_locked(g1_xy, g2, n, out, 1);
}
// `products` independent pairing products of `pairs_each` pairs in ONE pass of the pairing engine (the Miller loops of
// all products in one launch, one product tree per segment, the final exponentiations side by side): a verifier's five
// products cost the latency of one. Shorter products are padded with identity pairs, which contribute 1.
int tb200_multi_pairing_batch(const uint64_t* g1_xy, const uint64_t* g2, size_t products, size_t pairs_each, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (products == 0) return 0;
  if (!out || (pairs_each && (!g1_xy || !g2))) return fail(TB200_E_ARG, "null pointer");
  if (products > 4096 || pairs_each >= (1u << 20) || products * pairs_each >= (1u << 22))
    return fail(TB200_E_LIMIT, "batch of pairing products too large");
  if (pairs_each == 0 && products > 32) return fail(TB200_E_LIMIT, "more than 32 empty products");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  g.marks.clear();
  const size_t n = products * pairs_each;
  uint4 *d_p = nullptr, *d_q = nullptr, *d_o = nullptr;
  std::vector<void*> to_free;
  CU(cudaMallocAsync((void**)&d_o, products * 576, g.stream));
  to_free.push_back(d_o);
  if (n) {
    CU(cudaMallocAsync((void**)&d_p, n * 96, g.stream));
    to_free.push_back(d_p);
    CU(cudaMallocAsync((void**)&d_q, n * 192, g.stream));
    to_free.push_back(d_q);
    CU(cudaMemcpyAsync(d_p, g1_xy, n * 96, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_q, g2, n * 192, cudaMemcpyHostToDevice, g.stream));
  }
  int rc = pairing_products(g, d_p, d_q, (uint32_t)n, 0, (uint32_t)products, d_o, g.stream, nullptr, nullptr, true);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out, d_o, products * 576, cudaMemcpyDeviceToHost, g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "pairing result copy failed: %s", cudaGetErrorString(e));
  }
  cudaError_t e = cudaStreamSynchronize(g.stream);
  if (e != cudaSuccess && rc == 0) rc = fail((int)e, "device %d failed: %s", g.device, cudaGetErrorString(e));
  free_all(g, to_free);
  if (rc == 0) rc = finish_marks(g, g.stream);
  return rc;
}

}  // extern "C"
