// Host-side internals shared by the engine's translation units (no kernels here).
//
//   engine_core.cu     contexts (one per GPU, all driven from ONE process), worker threads, NCCL, the multi-GPU
//                      sharding of the host-facing entry points, device buffers, profiling
//   engine_g1.cu       G1 kernels + the MSM pipeline (sort -> accumulate -> fix-up -> reduce -> finalize)
//   engine_g2.cu       G2 kernels (the pipeline's point-touching stages over Fq2)
//   engine_pairing.cu  Fq12 tower, Miller loops, final exponentiation, endomorphism folds
//
// Each .cu holds the kernels it launches (whole-program device code, no -rdc), so the three heavy units compile in
// parallel; the handful of cross-unit launches go through the functions declared at the end of this header.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/testudo_b200.h"

namespace tbe {

// ---- errors --------------------------------------------------------------------------------------------------------
extern thread_local std::string g_err;
int fail(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
extern std::mutex g_mu;  // serialises the C ABI (the reference calls it from many rayon workers)

#define CU(expr)                                                                                                     \
  do {                                                                                                               \
    cudaError_t e__ = (expr);                                                                                        \
    if (e__ != cudaSuccess)                                                                                          \
      return ::tbe::fail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);     \
  } while (0)
#define LAUNCH(kernel, grid, block, stream, ...)           \
  do {                                                     \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__); \
    ::tbe::g_launches++;                                   \
    CU(cudaGetLastError());                                \
  } while (0)
#define LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)     \
  do {                                                          \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
    ::tbe::g_launches++;                                        \
    CU(cudaGetLastError());                                     \
  } while (0)

inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// Stream-ordered scratch of one host-facing call: everything allocated through it is released when the call returns,
// on the error paths (the CU() early returns) as well -- after a synchronisation there, since a failed call may still have
// copies in flight on the stream.
struct StreamScratch {
  cudaStream_t st;
  std::vector<void*> ptrs;
  bool done = false;   // set once the stream has been synchronised by the call itself
  explicit StreamScratch(cudaStream_t s) : st(s) {}
  StreamScratch(const StreamScratch&) = delete;
  StreamScratch& operator=(const StreamScratch&) = delete;
  template <class T>
  cudaError_t alloc(T** out, size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 16, st);
    if (e == cudaSuccess) ptrs.push_back(p);
    *out = static_cast<T*>(p);
    return e;
  }
  ~StreamScratch() {
    if (!done) cudaStreamSynchronize(st);
    for (void* p : ptrs) cudaFreeAsync(p, st);
  }
};

// ---- grow-only device arena: the pipeline carves its scratch out of one allocation (no cudaMalloc on the hot path)
// Several streams may use one arena one after the other (the *_dev entry points take a caller stream): `last_use`
// is recorded when a pipeline has been enqueued and awaited by the next user's stream before it touches the scratch.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  cudaEvent_t last_use = nullptr;
  bool used = false;
  int reserve(size_t bytes);          // may cudaFree / cudaMalloc (both synchronise the device)
  int acquire(cudaStream_t st);       // order `st` behind the previous user of the scratch
  int release(cudaStream_t st);       // record the end of this use
  void reset() { off = 0; }
  void destroy();
  template <class T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
  static size_t pad(size_t bytes) { return (bytes + 255) & ~size_t(255); }
};

struct Stage {
  const char* name;
  cudaEvent_t ev;
};

// ---- one context per GPU -------------------------------------------------------------------------------------------
struct Ctx {
  int slot = 0;     // index into Engine::devs (0 = primary)
  int device = -1;  // CUDA ordinal
  int sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // uploads of host-facing calls overlap the sort / accumulate stages
  cudaEvent_t ev_points = nullptr;
  cudaStream_t stream2 = nullptr;      // second pipeline for independent small MSMs (MIPP cross commitments)
  cudaEvent_t ev_join = nullptr;
  // large single MSMs: phases of different window ranges / point-range chunks side by side (engine_g1.cu)
  cudaStream_t split_stream[2] = {};   // high priority: their memory-bound sort kernels slip in next to an accumulation
  Arena split_arena[2];
  cudaEvent_t ev_split[4] = {};
  cudaStream_t pair_stream = nullptr;  // pairing products of a MIPP round, next to its cross MSMs
  cudaEvent_t ev_pair = nullptr, ev_pair2 = nullptr;
  Arena arena, arena2;
  static constexpr int SIDE = 8;       // side pipelines: the per-variable MSMs of a PST opening
  cudaStream_t side_stream[SIDE] = {};
  cudaEvent_t side_done[SIDE] = {};
  Arena side_arena[SIDE];
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<Stage> marks;
  std::map<std::string, double> stage_ms;
  int last_c = 0, last_W = 0, last_K = 0;
  uint64_t last_entries = 0, last_buckets = 0;
  // 64 KiB staging for host-facing calls (uint4 units): [0, RES_PART) results of the call, RES_PART: this GPU's partial
  // Miller product (576 B), RES_GATHER..: the all-gathered partials of every GPU (<= 64 x 576 B)
  static constexpr int RES_PART = 128, RES_GATHER = 256, RES_BYTES = 65536;
  uint4* d_result = nullptr;
  uint4* h_result = nullptr;  // pinned, same size
  std::vector<cudaEvent_t> chunk_ev;  // upload-chunk events of the host-facing batch path
  // worker thread: the per-GPU share of a sharded call runs here (its own current device, its own blocking copies)
  std::thread worker;
  std::mutex wmu;
  std::condition_variable wcv;
  std::function<int()> job;
  bool job_ready = false, job_done = false, quit = false;
  int job_rc = 0;
  std::string job_err;
  void* nccl_comm = nullptr;  // ncclComm_t of this device in the single-process clique
};

struct Engine {
  bool ready = false;
  std::vector<std::unique_ptr<Ctx>> devs;
  // tuning (process-wide)
  int forced_c = 0;
  size_t host_chunk_min = size_t(1) << 21;  // host-facing single MSMs: chunked upload/compute overlap from here
  size_t shard_min = size_t(1) << 18;       // host-facing single MSMs are sharded over the GPUs from ndev * this
  int pairing_team = 0;                     // lanes of a cooperative Fq12 team: 0 = by size (engine_pairing.cu), 32 / 33 (two pairs per warp) / 64 / 96 (pipelined Miller) forced
  int pairing_coop_max = 8192;              // Miller loops: CTA-per-pair up to this many pairs, lanes-per-pair above
  int host_upload_pace = 1;                 // host-facing single MSMs: upload chunk k once chunk k-2 is accumulated
  int host_chunk_count = 0;                 // tuning: chunk sizes in sixteenths of the points (0 = built-in schedule)
  int host_chunk_frac[16] = {};
  int msm_overlap = 0;                      // large single MSMs: window ranges / chunk sorts on side streams (engine_g1.cu);
                                            // measured neutral (84.2 -> 84.0 ms resident, 92.6 -> 91.4 ms from host): opt-in
  int mipp_cross_small_max = 64;            // tb200_mipp_cross_all: Straus path for the cross MSMs up to this many points
  int small_msm_max = 1024;                  // single G1 MSMs of up to this many points run in one CTA (kernels_small.cuh)
  int commit_pipeline = 0;                  // tb200_sqrt_pst_commit: Miller loops of a row chunk next to the next chunk's MSMs
  int acc_mode = 0;                         // 0 / 4: fused-Y3 XYZZ segments (default); 3: plain CIOS products
  uint64_t pass_entries_max = (1ull << 32) - 1024;  // sorted entries one pipeline pass can index (tests lower it)
  bool profiling = false;
  void* nccl_lib = nullptr;
  ~Engine();  // process exit without tb200_shutdown: stop the worker threads (no CUDA calls: the runtime may be gone)
};
extern Engine E;
inline Ctx& primary() { return *E.devs[0]; }
inline int ndev() { return (int)E.devs.size(); }
int need_ready();

// run `fn(slot)` for every device on its worker thread (slot 0 on the calling thread) and wait; returns the first
// failure (message moved to the caller's thread-local error string)
int for_each_device(const std::function<int(Ctx&)>& fn, int first = 0, int count = -1);
// all-gather `bytes` per device: d_send[slot] -> d_recv[slot] (ndev * bytes) on every device's main stream; NCCL over
// NVLink, enqueued from the calling thread inside one group
int all_gather(const std::vector<void*>& d_send, const std::vector<void*>& d_recv, size_t bytes);

int mark(Ctx& g, cudaStream_t st, const char* name);
int finish_marks(Ctx& g, cudaStream_t st);

// ---- handles ---------------------------------------------------------------------------------------------------------
}  // namespace tbe

struct tb200_srs {
  uint32_t n = 0;
  int c = 0, W = 0;
  std::vector<uint4*> table;  // per device slot: W * n affine points (replicated: SURVEY.md 8e)
  uint32_t extra = 0;         // columns appended behind the SRS proper (Hyrax blinding base h): n includes them
};
struct tb200_mipp {
  uint32_t n = 0;  // current length
  unsigned flags = 0;
  uint4* a = nullptr;             // n0 affine points
  uint32_t* y = nullptr;          // n0 scalars (8 limbs)
  uint32_t* scal = nullptr;       // 16 limbs staging for c, c_inv, one slot per round (the folds are only enqueued)
  uint32_t* scal_host = nullptr;  // pinned, same shape
  uint32_t* digits = nullptr;     // the two 128-bit halves of c over the G1 endomorphism, 8 words per round
  int round = 0;
  // two-phase fold (kernels_pairing.cuh): the multiples 2^j a_r[i] of the NEXT fold are computed on `pre_st` while the
  // round's cross values are; nullptr = one-phase fold (vector too long for the table)
  uint4* mult = nullptr;
  cudaStream_t pre_st = nullptr;
  cudaEvent_t ev_pre = nullptr, ev_fold = nullptr;
  uint16_t* sel = nullptr;        // per round: the selection list of the fold scalar (glv_host.h), device / pinned host
  uint16_t* sel_host = nullptr;
};
// the G2 commitment key of MIPP (m_h, src/mipp.rs:43,114): folded on its own stream, overlapping the G1 rounds
struct tb200_mipp_g2 {
  uint32_t n = 0;
  unsigned flags = 0;
  uint4* h = nullptr;             // n0 G2 affine points (12 uint4 each)
  uint32_t* scal = nullptr;       // device staging: one 8-limb scalar per round
  uint32_t* digits = nullptr;     // its four base-x digits, same shape
  uint32_t* scal_host = nullptr;  // pinned, same shape
  cudaStream_t st = nullptr;      // own stream: the folds overlap the G1 rounds on the library's streams
  int round = 0;
  uint4* mult = nullptr;          // two-phase fold, as in tb200_mipp
  cudaStream_t pre_st = nullptr;
  cudaEvent_t ev_pre = nullptr, ev_fold = nullptr;
  uint16_t* sel = nullptr;
  uint16_t* sel_host = nullptr;
};

namespace tbe {

// ---- engine_g1.cu: the pipeline -------------------------------------------------------------------------------------
// single MSM over device pointers on `st` (G1 or G2 points); result: 96 / 192 bytes at d_out
int msm_dev(Ctx& g, const void* d_bases, const void* d_scalars, size_t n, unsigned flags, void* d_out, cudaStream_t st,
            cudaEvent_t points_ready = nullptr, Arena* arena = nullptr, bool finish = true, bool g2 = false);
// host-facing single MSM of THIS device's share: uploads (chunked, overlapped), computes, leaves the affine result at
// g.d_result (device) -- the caller decides whether to download or all-gather it. Enqueue only; no synchronisation.
// `sharing` = the number of GPUs that pull from the same host during this call (picks the upload chunk schedule)
int msm_host_enqueue(Ctx& g, const uint64_t* bases_xy, const uint64_t* scalars, size_t n, unsigned flags,
                     std::vector<void*>& to_free, int sharing = 1);
// shared-base batch over device scalars; rows are processed in chunks that respect the per-pass limits
int batch_dev(Ctx& g, const uint4* table, int c, int W, uint32_t srs_n, const uint32_t* d_scalars, size_t rows,
              size_t cols, long long rs, long long cs, unsigned flags, uint4* d_out, cudaStream_t st);
int srs_build_table(Ctx& g, const uint64_t* bases_xy, size_t n, int c, int W, uint4** out_table);
int pick_c_batch(uint64_t cols);
int g1_sum_dev(Ctx& g, const void* d_pts, size_t n, void* d_out, cudaStream_t st);
// unit-test scaffolding: blocking copies around one launch on the primary's stream
int with_buffers(const void* a, size_t abytes, const void* b, size_t bbytes, void* o1, size_t o1bytes, void* o2,
                 size_t o2bytes, const std::function<int(char*, char*, char*, char*)>& launch);

// ---- engine_g2.cu: G2 stages of the pipeline --------------------------------------------------------------------------
int g2_accumulate(cudaStream_t st, uint32_t S_max, const uint32_t* entries, const uint32_t* starts, uint32_t B, uint32_t K,
                  const uint4* points, uint4* buckets, uint4* heads, int32_t* head_bucket);
int g2_fixup_round(cudaStream_t st, uint32_t S_max, const uint32_t* starts, uint32_t B, uint32_t K, uint32_t round,
                   uint4* heads, const int32_t* head_bucket);
int g2_fixup_final(cudaStream_t st, uint32_t S_max, const uint32_t* starts, uint32_t B, uint32_t K, uint4* buckets,
                   const uint4* heads, const int32_t* head_bucket);
int g2_reduce_pass_coop(cudaStream_t st, const uint4* inS, const uint4* inW, const uint32_t* level0, uint4* outS,
                        uint4* outW, uint32_t L, int log2_ell, uint64_t n);   // engine_pairing.cu
int g2_reduce_pass(cudaStream_t st, const uint4* inS, const uint4* inW, const uint32_t* level0, uint4* outS, uint4* outW,
                   uint32_t L, int log2_ell, uint64_t n);
int g2_finalize_single(cudaStream_t st, const uint4* group_w, int W, int c, uint4* d_out);

// ---- engine_pairing.cu ---------------------------------------------------------------------------------------------------
// window combine of a single G2 MSM over the twisted Frobenius (4 W parallel 64-doubling chains + a tree)
int g2_finalize_single_glv(cudaStream_t st, const uint4* group_w, int W, int c, uint4* fin_scratch, uint4* d_out);
// a_l + c a_r over the G1 endomorphism: digits of c (8 words at d_digits) then the fold of `split` elements
int g1_fold_glv(cudaStream_t st, const uint32_t* d_scaler, int mont, uint32_t* d_digits, uint4* a, uint32_t split);
// the same fold in two phases: multiples of a[first .. first + count) (independent of the scalar) / the fold itself
constexpr size_t FOLD_MULT_BYTES_MAX = size_t(3) << 30;      // per handle; longer vectors use the one-phase fold
inline size_t g1_fold_mult_bytes(size_t n) { return (n / 2) * 127 * 192; }
inline size_t g2_fold_mult_bytes(size_t n) { return (n / 2) * 64 * 384; }
int g1_fold_pre(cudaStream_t st, const uint4* a, uint32_t first, uint32_t count, uint4* mult);
int g1_fold_apply(cudaStream_t st, const uint16_t* d_sel, uint4* a, uint32_t split, const uint4* mult);
// Miller loops of n pairs -> product tree -> (optionally) final exponentiation; see engine_pairing.cu
int pairing_products(Ctx& g, const uint4* d_g1, const uint4* d_g2, uint32_t n, uint32_t xor_mask, uint32_t segs,
                     uint4* d_out, cudaStream_t st, cudaEvent_t after_miller, const uint4* d_gt_in = nullptr,
                     bool final_exp = true);

}  // namespace tbe
