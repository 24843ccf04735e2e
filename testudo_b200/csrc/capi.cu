// C ABI of the engine (include/testudo_b200.h): context, workspace arena, pipeline orchestration.
// Everything numerical happens in the kernels of kernels.cuh; there is no CPU arithmetic path in this file --
// if no CUDA device is present tb200_init fails and every other entry point returns TB200_E_STATE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/testudo_b200.h"
#include "kernels_affine.cuh"
#include "kernels_g2.cuh"
#include "kernels_pairing.cuh"
#include "kernels_smem.cuh"

using namespace tb;

namespace {

thread_local std::string g_err;
std::mutex g_mu;
uint64_t g_launches = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(expr)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define LAUNCH(kernel, grid, block, stream, ...)                          \
  do {                                                                    \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                \
    g_launches++;                                                         \
    CU(cudaGetLastError());                                               \
  } while (0)

inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// grow-only device arena: MSM calls carve their scratch out of one allocation (no cudaMalloc on the hot path)
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (base) {
      cudaError_t e = cudaFree(base);
      if (e != cudaSuccess) return (int)e;
      base = nullptr;
      cap = 0;
    }
    size_t want = bytes + (bytes >> 3);
    cudaError_t e = cudaMalloc((void**)&base, want);
    if (e != cudaSuccess) {
      e = cudaMalloc((void**)&base, bytes);
      want = bytes;
      if (e != cudaSuccess) return (int)e;
    }
    cap = want;
    return 0;
  }
  void reset() { off = 0; }
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

struct Ctx {
  bool ready = false;
  int device = -1;
  int sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // base upload of host-facing calls overlaps the digit/sort stages
  cudaEvent_t ev_points = nullptr;
  cudaStream_t stream2 = nullptr;      // second pipeline for independent small MSMs (MIPP cross commitments)
  cudaEvent_t ev_join = nullptr;
  cudaStream_t pair_stream = nullptr;  // pairing products of a MIPP round, next to its cross MSMs (created on first use)
  cudaEvent_t ev_pair = nullptr, ev_pair2 = nullptr;
  Arena arena;
  Arena arena2;
  // side pipelines for batches of independent small MSMs (the per-variable MSMs of a PST opening)
  static constexpr int SIDE = 8;
  cudaStream_t side_stream[SIDE] = {};
  cudaEvent_t side_done[SIDE] = {};
  Arena side_arena[SIDE];
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<Stage> marks;
  std::map<std::string, double> stage_ms;
  int forced_c = 0;
  size_t host_chunk_min = size_t(1) << 21;  // tb200_msm_g1 with host buffers: chunked upload/compute overlap from here
  int pairing_coop_max = 2048;  // Miller loops: warp-per-pair up to this many pairs, thread-per-pair above
  int acc_mode = 0;  // 0 = automatic (= 3), 1 = XYZZ segments, register operands (k_accumulate), 2 = batched-affine
                     // rounds, 3 = XYZZ segments, shared-memory operand slots (k_accumulate_s)
  // geometry of the last call
  int last_c = 0, last_W = 0, last_K = 0;
  uint64_t last_entries = 0, last_buckets = 0;
  uint4* d_result = nullptr;  // 96-byte staging for host-facing calls
  uint4* h_result = nullptr;  // pinned
} g;

int mark(cudaStream_t st, const char* name) {
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
int finish_marks(cudaStream_t st) {
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

// ---- window selection ---------------------------------------------------------------------------------------
// cost in mixed-addition units: accumulation entries + ~4 adds per bucket for the hierarchical reduction
int pick_c_single(uint64_t n) {
  if (g.forced_c) return g.forced_c;
  int best = 3;
  double bestc = 1e300;
  for (int c = 3; c <= 20; c++) {
    int W = num_windows(c);
    double cost = (double)W * ((double)n + 4.0 * (double)(1u << (c - 1)));
    int top_bits = SCALAR_BITS - (W - 1) * c;  // payload bits of the top window
    if (top_bits < 4) cost += 0.5 * (double)n;  // degenerate top window: contended atomics, one giant bucket
    if (cost < bestc) {
      bestc = cost;
      best = c;
    }
  }
  return best;
}
int pick_c_batch(uint64_t cols) {
  int best = 3;
  double bestc = 1e300;
  for (int c = 3; c <= 16; c++) {
    int W = num_windows(c);
    double cost = (double)W * (double)cols + 4.0 * (double)(1u << (c - 1));
    if (cost < bestc) {
      bestc = cost;
      best = c;
    }
  }
  return best;
}

// ---- the pipeline --------------------------------------------------------------------------------------------
struct Plan {
  MsmGeom geo;
  uint64_t M_max, B;
  uint32_t K, S_max, ntiles;
  std::vector<uint32_t> Ls;  // reduction fan-in per level
  size_t bytes;
  bool g2 = false;           // points are G2 (Fq2 coordinates: 192-byte affine, 384-byte XYZZ); single MSMs only
  bool affine;               // batched-affine rounds instead of k_accumulate
  uint64_t N1, N2;           // upper bounds of the round-0 / round-1 output counts
};

// a single MSM processed as point-range chunks that accumulate into ONE persistent bucket array
struct ChunkCtl {
  uint32_t ref_base;  // global index of the chunk's first point
  uint4* buckets;     // B * 192 B, all zero (= identity) before the first chunk
  bool last;          // run the reduction / finalisation after this chunk
};

int make_plan(Plan& p, uint32_t rows, uint32_t cols, long long rs, long long cs, int c, int batch, unsigned flags,
              bool g2 = false) {
  MsmGeom& q = p.geo;
  p.g2 = g2;
  const size_t pw = g2 ? 2 : 1;  // point width relative to G1
  q.ref_base = 0;
  q.rows = rows;
  q.cols = cols;
  q.row_stride = rs;
  q.col_stride = cs;
  q.c = c;
  q.W = num_windows(c);
  q.nb = 1u << (c - 1);
  q.batch = batch;
  q.groups = batch ? rows : (uint32_t)q.W;
  q.mont = (flags & TB200_SCALARS_MONT) ? 1 : 0;
  p.M_max = (uint64_t)rows * cols * q.W;
  p.B = (uint64_t)q.groups * q.nb;
  if (p.M_max >= (1ull << 32) - 512 || p.B >= (1ull << 31))
    return fail(TB200_E_LIMIT, "MSM too large for one pass: %llu entries, %llu buckets",
                (unsigned long long)p.M_max, (unsigned long long)p.B);
  uint64_t target_threads = (uint64_t)g.sms * 3 * ACC_THREADS * 4;
  uint64_t K = (p.M_max + target_threads - 1) / target_threads;
  p.K = (uint32_t)std::min<uint64_t>(256, std::max<uint64_t>(4, K));
  p.S_max = cdiv(std::max<uint64_t>(p.M_max, 1), p.K);
  p.ntiles = (uint32_t)(p.B / SCAN_TILE + 1);
  p.Ls.clear();
  // Fan-in 32 at level 0 (throughput-bound: millions of buckets). The levels above it of a SINGLE MSM hold few
  // elements and are latency-bound (2L - 1 sequential additions + log2(ell) doublings per thread): fan-in 8 there
  // (measured at 2^24, c = 20: 3.5 -> 1.9 ms for the upper levels). Batches keep 32: thousands of rows fill the GPU.
  for (uint32_t n = q.nb; n > 1;) {
    uint32_t L = std::min<uint32_t>(n, (batch || p.Ls.empty()) ? 32 : 8);
    p.Ls.push_back(L);
    n /= L;
  }
  // batched-affine accumulation pays off once the GPU is saturated and buckets hold several points
  p.N1 = (p.M_max + p.B) / 2 + 1;
  p.N2 = (p.N1 + p.B) / 2 + 1;
  // Opt-in only. Measured on B200 at 2^24 (profiles/r01_summary.md): the rounds move ~105 GB through HBM and pay
  // one Fermat inversion per thread per round, which outweighs the 30% fewer multiplications; the XYZZ segment
  // kernel (85% of the integer pipe) stays the automatic choice.
  p.affine = g.acc_mode == 2 && !g2;
  // arena size
  size_t b = 0;
  b += Arena::pad((p.B + 1) * 4) * 2;  // counts, starts
  b += Arena::pad(p.B * 4);            // cursors
  b += Arena::pad((size_t)p.ntiles * 4 + 4);
  b += Arena::pad(std::max<uint64_t>(p.M_max, 1) * 4);  // entries
  if (p.affine) {
    b += Arena::pad((p.B + 1) * 4) * 3;                 // sizes, two offset arrays
    b += Arena::pad(p.N1 * 4) + Arena::pad(p.N1 * 48) + Arena::pad(p.N1);  // pair index, prefix scratch, kinds
    b += Arena::pad(p.N1 * 96) + Arena::pad(p.N2 * 96); // ping-pong point arrays
  } else {
    b += Arena::pad(p.B * 192 * pw);                    // buckets
    b += Arena::pad((size_t)p.S_max * 192 * pw) + Arena::pad((size_t)p.S_max * 4);
  }
  uint64_t n = p.B;
  for (uint32_t L : p.Ls) {
    n /= L;
    b += Arena::pad(n * 192 * pw) * 2;
  }
  b += Arena::pad((size_t)q.groups * 192 * pw) * 2 + 4096;
  if (g2) b += Arena::pad((size_t)4 * 96 * 384);        // window-combine partial sums (k_finalize_single_g2_glv)
  p.bytes = b;
  return 0;
}

// Runs the whole pipeline on `st`. d_points: affine points indexed by entry refs. d_out: groups*96 B (batch) or 96 B.
int run_pipeline(const Plan& p, const uint32_t* d_scalars, const uint4* d_points, uint4* d_out, cudaStream_t st,
                 cudaEvent_t points_ready = nullptr, Arena* arena_p = nullptr, const ChunkCtl* chunk = nullptr) {
  MsmGeom q = p.geo;
  if (chunk) q.ref_base = chunk->ref_base;
  const size_t pw = p.g2 ? 2 : 1;
  Arena& arena = arena_p ? *arena_p : g.arena;
  int rc = arena.reserve(p.bytes);
  if (rc) return fail(rc, "workspace allocation of %zu bytes failed: %s", p.bytes, cudaGetErrorString((cudaError_t)rc));
  arena.reset();
  uint32_t* counts = arena.take<uint32_t>(p.B + 1);
  uint32_t* starts = arena.take<uint32_t>(p.B + 1);
  uint32_t* cursors = arena.take<uint32_t>(p.B);
  uint32_t* tile_sums = arena.take<uint32_t>(p.ntiles + 1);
  uint32_t* entries = arena.take<uint32_t>(std::max<uint64_t>(p.M_max, 1));
  uint4 *buckets = nullptr, *heads = nullptr;
  int32_t* head_bucket = nullptr;
  uint32_t *sizes = nullptr, *offA = nullptr, *offB = nullptr, *pidx = nullptr;
  uint4 *scratch = nullptr, *ptsA = nullptr, *ptsB = nullptr;
  uint8_t* kinds = nullptr;
  if (p.affine) {
    sizes = arena.take<uint32_t>(p.B + 1);
    offA = arena.take<uint32_t>(p.B + 1);
    offB = arena.take<uint32_t>(p.B + 1);
    pidx = arena.take<uint32_t>(p.N1);
    scratch = arena.take<uint4>(p.N1 * 3);
    kinds = arena.take<uint8_t>(p.N1);
    ptsA = arena.take<uint4>(p.N1 * 6);
    ptsB = arena.take<uint4>(p.N2 * 6);
  } else {
    buckets = chunk ? chunk->buckets : arena.take<uint4>(p.B * 12 * pw);
    heads = arena.take<uint4>((size_t)p.S_max * 12 * pw);
    head_bucket = arena.take<int32_t>(p.S_max);
  }

  g.last_c = q.c;
  g.last_W = q.W;
  g.last_K = (int)p.K;
  g.last_entries = p.M_max;
  g.last_buckets = p.B;

  const uint64_t items = (uint64_t)q.rows * q.cols;
  if (items == 0) {
    uint32_t cnt = q.batch ? q.rows : 1;
    if (cnt) LAUNCH(k_write_identity, cdiv(cnt * 6 * pw, 128), 128, st, d_out, (uint32_t)(cnt * pw));
    return 0;
  }
  if (mark(st, "begin")) return 1;
  const uint32_t dig_grid = (uint32_t)std::min<uint64_t>(cdiv(items, 256), (uint64_t)g.sms * 16);
  // batches with enough rows to fill the GPU sort each row inside one CTA's shared memory
  const size_t row_smem = (size_t)q.nb * 4;
  const bool row_sort = q.batch && q.rows >= (uint32_t)g.sms && row_smem <= 160 * 1024;
  // >= 116 KB of dynamic shared memory => one CTA per SM (see k_batch_digits)
  const size_t row_smem_launch = std::max<size_t>(row_smem, 116 * 1024);
  if (row_sort) {
    CU(cudaFuncSetAttribute(k_batch_digits<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem_launch));
    CU(cudaFuncSetAttribute(k_batch_digits<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem_launch));
    k_batch_digits<false><<<q.rows, 1024, row_smem_launch, st>>>(d_scalars, q, counts, (uint32_t*)nullptr);
    g_launches++;
    CU(cudaGetLastError());
  } else {
    CU(cudaMemsetAsync(counts, 0, (p.B + 1) * 4, st));
    LAUNCH(k_digits<false>, dig_grid, 256, st, d_scalars, q, counts, (uint32_t*)nullptr, -1);
  }
  if (mark(st, "digits")) return 1;
  LAUNCH(k_scan_tile_sums, p.ntiles, SCAN_THREADS, st, counts, (uint32_t)p.B, tile_sums);
  LAUNCH(k_scan_tile_offsets, 1, SCAN_THREADS, st, tile_sums, p.ntiles, tile_sums + p.ntiles);
  LAUNCH(k_scan_apply, p.ntiles, SCAN_THREADS, st, counts, (uint32_t)p.B, tile_sums, starts, cursors);
  if (mark(st, "scan")) return 1;
  if (row_sort) {
    k_batch_digits<true><<<q.rows, 1024, row_smem_launch, st>>>(d_scalars, q, starts, entries);
    g_launches++;
    CU(cudaGetLastError());
  } else {
    // large single MSMs: one pass per window keeps the writes of a pass inside an L2-sized slice of entries[]
    const bool per_window = !q.batch && q.c >= 19 && (uint64_t)q.cols * 4 * q.W > (64ull << 20);
    if (per_window) {
      for (int w = 0; w < q.W; w++) LAUNCH(k_digits<true>, dig_grid, 256, st, d_scalars, q, cursors, entries, w);
    } else {
      LAUNCH(k_digits<true>, dig_grid, 256, st, d_scalars, q, cursors, entries, -1);
    }
  }
  if (mark(st, "scatter")) return 1;
  if (points_ready) CU(cudaStreamWaitEvent(st, points_ready, 0));  // bases may still be in flight until here
  const uint4 *inS = nullptr, *inW = nullptr;
  uint64_t n = p.B;
  int log2_ell = 0;
  size_t first_level = 0;
  if (p.affine) {
    // ---- batched-affine rounds: every bucket shrinks from n to ceil(n/2) points per round ------------------------
    uint32_t* d_max = tile_sums + p.ntiles;  // one spare word next to the scan scratch
    CU(cudaMemsetAsync(d_max, 0, 4, st));
    LAUNCH(k_max_size, (uint32_t)std::min<uint64_t>(cdiv(p.B, 256), (uint64_t)g.sms * 8), 256, st, starts,
           (uint32_t)p.B, d_max);
    uint32_t maxn = 0;
    CU(cudaMemcpyAsync(&maxn, d_max, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));  // the only host round trip of the pipeline: the number of rounds
    const uint32_t* off_in = starts;
    uint32_t* off_bufs[2] = {offA, offB};
    uint4* pt_bufs[2] = {ptsA, ptsB};
    const uint4* cur = nullptr;
    uint64_t n_in_bound = p.M_max;
    int round = 0;
    for (; maxn > 1; maxn = (maxn + 1) / 2, round++) {
      uint32_t* off_out = off_bufs[round & 1];
      uint4* nxt = pt_bufs[round & 1];
      const uint64_t n_out_bound = (n_in_bound + p.B) / 2 + 1;
      LAUNCH(k_half_sizes, cdiv(p.B, 256), 256, st, off_in, (uint32_t)p.B, sizes);
      LAUNCH(k_scan_tile_sums, p.ntiles, SCAN_THREADS, st, sizes, (uint32_t)p.B, tile_sums);
      LAUNCH(k_scan_tile_offsets, 1, SCAN_THREADS, st, tile_sums, p.ntiles, tile_sums + p.ntiles);
      LAUNCH(k_scan_apply, p.ntiles, SCAN_THREADS, st, sizes, (uint32_t)p.B, tile_sums, off_out, cursors);
      LAUNCH(k_pair_index, cdiv(cdiv(n_out_bound, 8), 256), 256, st, off_in, off_out, (uint32_t)p.B, pidx);
      // outputs per thread: enough threads for ~2 waves, at least 64 outputs to amortise the inversion
      uint64_t threads_target = (uint64_t)g.sms * 3 * 128 * 2;
      uint32_t T = (uint32_t)std::min<uint64_t>(512, std::max<uint64_t>(64, n_out_bound / threads_target));
      uint32_t warps = cdiv(n_out_bound, 32ull * T);
      if (round == 0)
        LAUNCH(k_affine_round<true>, cdiv((uint64_t)warps * 32, 128), 128, st, pidx, off_out, (uint32_t)p.B, T, entries,
               d_points, cur, nxt, scratch, kinds);
      else
        LAUNCH(k_affine_round<false>, cdiv((uint64_t)warps * 32, 128), 128, st, pidx, off_out, (uint32_t)p.B, T,
               entries, d_points, cur, nxt, scratch, kinds);
      off_in = off_out;
      cur = nxt;
      n_in_bound = n_out_bound;
    }
    if (mark(st, "accumulate")) return 1;
    if (mark(st, "fixup")) return 1;
    // level 0 of the reduction reads the (at most one) affine point of every bucket
    const uint32_t L0 = p.Ls.empty() ? 1 : p.Ls[0];
    if (round == 0) {
      // every bucket already has <= 1 entry: materialise the points once through a copy round
      uint32_t* off_out = off_bufs[0];
      LAUNCH(k_half_sizes, cdiv(p.B, 256), 256, st, off_in, (uint32_t)p.B, sizes);
      LAUNCH(k_scan_tile_sums, p.ntiles, SCAN_THREADS, st, sizes, (uint32_t)p.B, tile_sums);
      LAUNCH(k_scan_tile_offsets, 1, SCAN_THREADS, st, tile_sums, p.ntiles, tile_sums + p.ntiles);
      LAUNCH(k_scan_apply, p.ntiles, SCAN_THREADS, st, sizes, (uint32_t)p.B, tile_sums, off_out, cursors);
      LAUNCH(k_pair_index, cdiv(cdiv(p.N1, 8), 256), 256, st, off_in, off_out, (uint32_t)p.B, pidx);
      LAUNCH(k_affine_round<true>, cdiv(cdiv(p.N1, 32ull * 64) * 32ull, 128), 128, st, pidx, off_out, (uint32_t)p.B,
             64u, entries, d_points, cur, ptsA, scratch, kinds);
      off_in = off_out;
      cur = ptsA;
    }
    n /= L0;
    uint4* outS = arena.take<uint4>(n * 12);
    uint4* outW = arena.take<uint4>(n * 12);
    LAUNCH(k_reduce_pass0_affine, cdiv(n, 128), 128, st, cur, off_in, outS, outW, L0, n);
    inS = outS;
    inW = outW;
    for (uint32_t v = L0; v > 1; v >>= 1) log2_ell++;
    first_level = 1;
  } else {
  // M is only known on the device (starts[B]); launch for the upper bound, surplus threads exit immediately
  if (p.g2) {
    LAUNCH(k_accumulate_g2, cdiv(p.S_max, 64), 64, st, entries, starts, (uint32_t)p.B, p.K, d_points, buckets, heads,
           head_bucket);
  } else if (g.acc_mode == 0 || g.acc_mode >= 3 || chunk) {  // operands in shared-memory slots (kernels_smem.cuh)
    const dim3 grid(cdiv(p.S_max, ACCS_THREADS));
    // mode 3: plain CIOS products; 0 / 4: Y3 as one fused sum of two products (default); 5: + Karatsuba singles
#define TB_ACCS(V)                                                                                                   \
  do {                                                                                                               \
    CU(cudaFuncSetAttribute(k_accumulate_s<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACCS_SMEM));             \
    k_accumulate_s<V><<<grid, ACCS_THREADS, ACCS_SMEM, st>>>(entries, starts, (uint32_t)p.B, p.K, d_points, buckets, \
                                                             heads, head_bucket, chunk ? 1 : 0);                    \
  } while (0)
    if (g.acc_mode == 3) TB_ACCS(0);
    else if (g.acc_mode == 5) TB_ACCS(3);
    else TB_ACCS(2);
#undef TB_ACCS
    g_launches++;
    CU(cudaGetLastError());
  } else {
    LAUNCH(k_accumulate, cdiv(p.S_max, ACC_THREADS), ACC_THREADS, st, entries, starts, (uint32_t)p.B, p.K, d_points,
           buckets, heads, head_bucket);
  }
  if (mark(st, "accumulate")) return 1;
  // A bucket holds at most one entry per (column, window) of its group: cols entries (single MSM: one window per
  // group) or cols * W (batch row), i.e. it spans at most that many / K + 1 segments.
  const uint64_t max_bucket = q.batch ? (uint64_t)q.cols * q.W : (uint64_t)q.cols;
  const uint64_t max_span = std::min<uint64_t>(p.S_max, max_bucket / p.K + 2);
  if (p.g2) {
    for (uint32_t round = 0; (1ull << round) < max_span; round++)
      LAUNCH(k_fixup_round_g2, cdiv(p.S_max, 64), 64, st, starts, (uint32_t)p.B, p.K, round, heads, head_bucket);
    LAUNCH(k_fixup_final_g2, cdiv(p.S_max, 64), 64, st, starts, (uint32_t)p.B, p.K, buckets, heads, head_bucket);
  } else {
    for (uint32_t round = 0; (1ull << round) < max_span; round++)
      LAUNCH(k_fixup_round, cdiv(p.S_max, 128), 128, st, starts, (uint32_t)p.B, p.K, round, heads, head_bucket);
    LAUNCH(k_fixup_final, cdiv(p.S_max, 128), 128, st, starts, (uint32_t)p.B, p.K, buckets, heads, head_bucket);
  }
  if (mark(st, "fixup")) return 1;
  if (chunk && !chunk->last) return 0;  // later point-range chunks continue in the same buckets
  inS = buckets;
  }
  // hierarchical bucket reduction
  // chunked runs: emptiness is per chunk; the persistent buckets carry the identity (all zero) instead
  const uint32_t* level0 = (p.affine || chunk) ? nullptr : starts;
  for (size_t li = first_level; li < p.Ls.size(); li++) {
    const uint32_t L = p.Ls[li];
    n /= L;
    uint4* outS = arena.take<uint4>(n * 12 * pw);
    uint4* outW = arena.take<uint4>(n * 12 * pw);
    if (p.g2) LAUNCH(k_reduce_pass_g2, cdiv(n, 64), 64, st, inS, inW, level0, outS, outW, L, log2_ell, n);
    else LAUNCH(k_reduce_pass, cdiv(n, 128), 128, st, inS, inW, level0, outS, outW, L, log2_ell, n);
    inS = outS;
    inW = outW;
    level0 = nullptr;
    for (uint32_t v = L; v > 1; v >>= 1) log2_ell++;
  }
  const uint4* group_w = inW;  // nb >= 4 (c >= 3): at least one reduction level has run
  if (mark(st, "reduce")) return 1;
  if (p.g2 && q.W <= 96) {
    // window combine over the twisted Frobenius: 4 W parallel 64-doubling chains + a tree instead of ~250 serial doublings
    uint4* fin = arena.take<uint4>((size_t)4 * q.W * 24);
    LAUNCH(k_finalize_single_g2_glv, 1, 384, st, group_w, q.W, q.c, fin, d_out);
  } else if (p.g2) LAUNCH(k_finalize_single_g2, 1, 32, st, group_w, q.W, q.c, d_out);
  else if (q.batch) LAUNCH(k_finalize_batch, cdiv(q.groups, 128), 128, st, group_w, q.groups, d_out);
  else LAUNCH(k_finalize_single, 1, 32, st, group_w, q.W, q.c, d_out);
  if (mark(st, "finalize")) return 1;
  return 0;
}

int need_ready() {
  if (!g.ready) return fail(TB200_E_STATE, "tb200_init has not been called (or failed): no CUDA context");
  return 0;
}

}  // namespace

struct tb200_srs {
  uint32_t n = 0;
  int c = 0, W = 0;
  uint4* table = nullptr;  // W * n affine points
};
struct tb200_mipp {
  uint32_t n = 0;       // current length
  unsigned flags = 0;
  uint4* a = nullptr;   // n0 affine points
  uint32_t* y = nullptr;  // n0 scalars (8 limbs)
  uint32_t* scal = nullptr;  // 16 limbs staging for c, c_inv, one slot per round (the folds are only enqueued)
  uint32_t* scal_host = nullptr;  // pinned, same shape
  uint32_t* digits = nullptr;     // the two 128-bit halves of c over the G1 endomorphism (k_glv2_digits), 8 words per round
  int round = 0;
};

// the G2 commitment key of MIPP (m_h, src/mipp.rs:43,114): folded on its own stream, overlapping the G1 rounds
struct tb200_mipp_g2 {
  uint32_t n = 0;
  unsigned flags = 0;
  uint4* h = nullptr;          // n0 G2 affine points (12 uint4 each)
  uint32_t* scal = nullptr;    // device staging: one 8-limb scalar per round
  uint32_t* digits = nullptr;  // its four base-x digits (k_glv4_digits), same shape
  uint32_t* scal_host = nullptr;  // pinned, same shape (every round has its own slot: the copies are asynchronous)
  cudaStream_t st = nullptr;      // own stream: the folds overlap the G1 rounds on the library's streams
  int round = 0;
};

extern "C" {

int tb200_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g.ready) return 0;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(e != cudaSuccess ? (int)e : TB200_E_STATE,
                "no CUDA device available (%s): testudo_b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0) CU(cudaGetDevice(&device));
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  g.device = device;
  g.sms = prop.multiProcessorCount;
  {  // keep freed staging buffers in the pool: with the default threshold (0) every synchronisation returns them
     // to the OS and the next host-facing call pays hundreds of ms to map gigabytes again
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
  CU(cudaMalloc((void**)&g.d_result, 16384));
  CU(cudaMallocHost((void**)&g.h_result, 16384));
  if (const char* m = getenv("TB200_HOST_CHUNK_MIN")) g.host_chunk_min = (size_t)atoll(m);  // tuning aid
  if (const char* m = getenv("TB200_ACC_MODE")) {  // tuning aid: same effect as tb200_set_accumulate_mode
    int v = atoi(m);
    g.acc_mode = (v >= 0 && v <= 5) ? v : 0;
  }
  g.ready = true;
  return 0;
}

void tb200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g.ready) return;
  cudaSetDevice(g.device);
  cudaStreamSynchronize(g.stream);
  cudaStreamSynchronize(g.stream2);
  if (g.arena.base) cudaFree(g.arena.base);
  if (g.arena2.base) cudaFree(g.arena2.base);
  for (int i = 0; i < g.SIDE; i++) {
    if (g.side_stream[i]) {
      cudaStreamSynchronize(g.side_stream[i]);
      cudaStreamDestroy(g.side_stream[i]);
      cudaEventDestroy(g.side_done[i]);
      g.side_stream[i] = nullptr;
    }
    if (g.side_arena[i].base) cudaFree(g.side_arena[i].base);
    g.side_arena[i] = Arena();
  }
  g.arena = Arena();
  g.arena2 = Arena();
  cudaStreamDestroy(g.stream2);
  cudaEventDestroy(g.ev_join);
  if (g.pair_stream) {
    cudaStreamSynchronize(g.pair_stream);
    cudaStreamDestroy(g.pair_stream);
    cudaEventDestroy(g.ev_pair);
    cudaEventDestroy(g.ev_pair2);
    g.pair_stream = nullptr;
  }
  for (auto e : g.ev_pool) cudaEventDestroy(e);
  g.ev_pool.clear();
  cudaFree(g.d_result);
  cudaFreeHost(g.h_result);
  cudaStreamDestroy(g.stream);
  cudaStreamDestroy(g.copy_stream);
  cudaEventDestroy(g.ev_points);
  g.ready = false;
}

const char* tb200_last_error(void) { return g_err.c_str(); }
uint64_t tb200_launch_count(void) { return g_launches; }
void tb200_reset_launch_count(void) { g_launches = 0; }
void tb200_set_profiling(int enabled) { g.profiling = enabled != 0; }
double tb200_stage_ms(const char* stage) {
  auto it = g.stage_ms.find(stage ? stage : "");
  return it == g.stage_ms.end() ? -1.0 : it->second;
}
int tb200_last_geometry(int* c, int* windows, uint64_t* entries, uint64_t* buckets, int* segment) {
  if (c) *c = g.last_c;
  if (windows) *windows = g.last_W;
  if (entries) *entries = g.last_entries;
  if (buckets) *buckets = g.last_buckets;
  if (segment) *segment = g.last_K;
  return 0;
}
void tb200_set_pairing_coop_max(int n) { g.pairing_coop_max = n < 0 ? 0 : n; }
void tb200_set_window_bits(int c) { g.forced_c = (c >= 3 && c <= 22) ? c : 0; }
void tb200_set_accumulate_mode(int mode) { g.acc_mode = (mode >= 0 && mode <= 5) ? mode : 0; }

// ---- single MSM -------------------------------------------------------------------------------------------------
static int msm_dev_locked(const void* d_bases, const void* d_scalars, size_t n, unsigned flags, void* d_out,
                          cudaStream_t st, cudaEvent_t points_ready = nullptr, Arena* arena = nullptr,
                          bool finish = true, bool g2 = false) {
  if (n >= (1ull << 31)) return fail(TB200_E_LIMIT, "n = %zu exceeds 2^31 - 1 points per call", n);
  if (((uintptr_t)d_bases | (uintptr_t)d_scalars | (uintptr_t)d_out) & 15)
    return fail(TB200_E_ARG, "device pointers must be 16-byte aligned");
  Plan p;
  int c = pick_c_single(std::max<size_t>(n, 1));
  int rc = make_plan(p, 1, (uint32_t)n, 0, 1, c, 0, flags, g2);
  while (rc == TB200_E_LIMIT && c > 3) rc = make_plan(p, 1, (uint32_t)n, 0, 1, --c, 0, flags, g2);
  if (rc) return rc;
  g.marks.clear();
  rc = run_pipeline(p, (const uint32_t*)d_scalars, (const uint4*)d_bases, (uint4*)d_out, st, points_ready, arena);
  if (rc) return rc;
  return finish ? finish_marks(st) : 0;
}

int tb200_msm_g1_dev(const void* d_bases_xy, const void* d_scalars, size_t n, unsigned flags, void* d_out_xy,
                     void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out_xy || (n && (!d_bases_xy || !d_scalars))) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(g.device));
  return msm_dev_locked(d_bases_xy, d_scalars, n, flags, d_out_xy, stream ? (cudaStream_t)stream : g.stream);
}

// Host-facing single MSM for large n: point-range chunks of growing size (1/8, 1/4, 1/4, 3/8 of the points). Chunk
// k+1 is uploaded on the copy stream while chunk k is sorted and accumulated; every chunk accumulates into the SAME
// persistent bucket array (k_accumulate_s continues from the stored bucket, no extra group operations), and the
// reduction / finalisation run once after the last chunk. Without it the accumulation waits for the whole 96 n byte
// base upload (~37 ms at 2^24 over PCIe Gen5) with the GPU idle.
static int msm_host_chunked(const uint64_t* bases_xy, const uint64_t* scalars, size_t n, unsigned flags,
                            uint64_t out_xy[12]) {
  constexpr int C = 4;
  const size_t unit = ((n + 7) / 8 + 31) & ~size_t(31);
  const size_t cut[C + 1] = {0, std::min(n, unit), std::min(n, 3 * unit), std::min(n, 5 * unit), n};
  const int c = pick_c_single(n);
  Plan plans[C];
  size_t B = 0;
  for (int k = 0; k < C; k++) {
    int rc = make_plan(plans[k], 1, (uint32_t)(cut[k + 1] - cut[k]), 0, 1, c, 0, flags);
    if (rc) return rc;
    B = plans[k].B;
  }
  uint4 *d_b = nullptr, *d_s = nullptr, *d_buckets = nullptr;
  CU(cudaMallocAsync((void**)&d_b, n * 96, g.stream));
  CU(cudaMallocAsync((void**)&d_s, n * 32, g.stream));
  CU(cudaMallocAsync((void**)&d_buckets, B * 192, g.stream));
  CU(cudaMemsetAsync(d_buckets, 0, B * 192, g.stream));
  CU(cudaEventRecord(g.ev_points, g.stream));  // allocations exist
  CU(cudaStreamWaitEvent(g.copy_stream, g.ev_points, 0));
  cudaEvent_t ev_s[C], ev_b[C];
  for (int k = 0; k < C; k++) {
    const size_t lo = cut[k], cnt = cut[k + 1] - cut[k];
    CU(cudaEventCreateWithFlags(&ev_s[k], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev_b[k], cudaEventDisableTiming));
    CU(cudaMemcpyAsync((char*)d_s + lo * 32, (const char*)scalars + lo * 32, cnt * 32, cudaMemcpyHostToDevice,
                       g.copy_stream));
    CU(cudaEventRecord(ev_s[k], g.copy_stream));
    CU(cudaMemcpyAsync((char*)d_b + lo * 96, (const char*)bases_xy + lo * 96, cnt * 96, cudaMemcpyHostToDevice,
                       g.copy_stream));
    CU(cudaEventRecord(ev_b[k], g.copy_stream));
  }
  g.marks.clear();
  int rc = 0;
  for (int k = 0; k < C && rc == 0; k++) {
    CU(cudaStreamWaitEvent(g.stream, ev_s[k], 0));
    ChunkCtl ctl{(uint32_t)cut[k], d_buckets, k == C - 1};
    rc = run_pipeline(plans[k], (const uint32_t*)d_s + 8 * cut[k], d_b, g.d_result, g.stream, ev_b[k], nullptr, &ctl);
  }
  if (rc == 0) rc = finish_marks(g.stream);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(g.h_result, g.d_result, 96, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
    else memcpy(out_xy, g.h_result, 96);
  } else {
    cudaStreamSynchronize(g.stream);
    cudaStreamSynchronize(g.copy_stream);
  }
  for (int k = 0; k < C; k++) {
    cudaEventDestroy(ev_s[k]);
    cudaEventDestroy(ev_b[k]);
  }
  cudaFreeAsync(d_b, g.stream);
  cudaFreeAsync(d_s, g.stream);
  cudaFreeAsync(d_buckets, g.stream);
  return rc;
}

int tb200_msm_g1(const uint64_t* bases_xy, const uint64_t* scalars, size_t n, unsigned flags, uint64_t out_xy[12]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out_xy || (n && (!bases_xy || !scalars))) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(g.device));
  // large inputs: overlap the upload with the accumulation (every chunk is non-empty and keeps the GPU busy)
  if (n >= g.host_chunk_min && n < (size_t(1) << 31) && g.acc_mode != 1 && g.acc_mode != 2)
    return msm_host_chunked(bases_xy, scalars, n, flags, out_xy);
  uint4 *d_b = nullptr, *d_s = nullptr;
  if (n) {
    CU(cudaMallocAsync((void**)&d_b, n * 96, g.stream));
    CU(cudaMallocAsync((void**)&d_s, n * 32, g.stream));
    // the digit/sort stages only need the scalars; the 3x larger base upload runs on the copy stream and is
    // awaited right before the accumulation kernel
    CU(cudaEventRecord(g.ev_points, g.stream));  // allocations done
    CU(cudaStreamWaitEvent(g.copy_stream, g.ev_points, 0));
    CU(cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_b, bases_xy, n * 96, cudaMemcpyHostToDevice, g.copy_stream));
    CU(cudaEventRecord(g.ev_points, g.copy_stream));
  }
  int rc = msm_dev_locked(d_b ? d_b : g.d_result, d_s ? d_s : g.d_result, n, flags, g.d_result, g.stream,
                          n ? g.ev_points : nullptr);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(g.h_result, g.d_result, 96, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
    else memcpy(out_xy, g.h_result, 96);
  }
  if (d_b) cudaFreeAsync(d_b, g.stream);
  if (d_s) cudaFreeAsync(d_s, g.stream);
  return rc;
}

// ---- G2 (MultilinearPC::open, commit_g2, G2 compress) -------------------------------------------
int tb200_msm_g2_dev(const void* d_bases, const void* d_scalars, size_t n, unsigned flags, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || (n && (!d_bases || !d_scalars))) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(g.device));
  return msm_dev_locked(d_bases, d_scalars, n, flags, d_out, stream ? (cudaStream_t)stream : g.stream, nullptr, nullptr,
                        true, true);
}

int tb200_msm_g2(const uint64_t* bases, const uint64_t* scalars, size_t n, unsigned flags, uint64_t out[24]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && (!bases || !scalars))) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(g.device));
  uint4 *d_b = nullptr, *d_s = nullptr;
  if (n) {
    CU(cudaMallocAsync((void**)&d_b, n * 192, g.stream));
    CU(cudaMallocAsync((void**)&d_s, n * 32, g.stream));
    CU(cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_b, bases, n * 192, cudaMemcpyHostToDevice, g.stream));
  }
  int rc = msm_dev_locked(d_b ? d_b : g.d_result, d_s ? d_s : g.d_result, n, flags, g.d_result, g.stream, nullptr,
                          nullptr, true, true);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(g.h_result, g.d_result, 192, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
    else memcpy(out, g.h_result, 192);
  }
  if (d_b) cudaFreeAsync(d_b, g.stream);
  if (d_s) cudaFreeAsync(d_s, g.stream);
  return rc;
}

int tb200_compress_g2(uint64_t* vec, size_t split, const uint64_t scaler[4], unsigned flags) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!vec || !scaler) return fail(TB200_E_ARG, "null pointer");
  if (split == 0) return 0;
  if (split >= (1u << 26)) return fail(TB200_E_LIMIT, "split too large");
  CU(cudaSetDevice(g.device));
  uint4* d_v = nullptr;
  uint32_t* d_k = nullptr;
  CU(cudaMallocAsync((void**)&d_v, 2 * split * 192, g.stream));
  CU(cudaMallocAsync((void**)&d_k, 32, g.stream));
  CU(cudaMemcpyAsync(d_v, vec, 2 * split * 192, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d_k, scaler, 32, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_compress_g2, cdiv(split, 64), 64, g.stream, d_v, (uint32_t)split, d_k,
         (flags & TB200_SCALARS_MONT) ? 1 : 0);
  CU(cudaMemcpyAsync(vec, d_v, split * 192, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  cudaFreeAsync(d_v, g.stream);
  cudaFreeAsync(d_k, g.stream);
  return 0;
}

// ---- MultilinearPC::open (G2 proofs) / open_g1 (G1 proofs): quotient loop on the device, one MSM per variable ---------
static int pst_open_locked(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                           unsigned flags, uint64_t* proofs, bool g2) {
  if (!evals || !point || !level_bases || !proofs) return fail(TB200_E_ARG, "null pointer");
  if (nv == 0) return 0;
  if (nv > 28) return fail(TB200_E_LIMIT, "nv = %zu: at most 2^28 evaluations", nv);
  for (size_t i = 0; i < nv; i++)
    if (!level_bases[i]) return fail(TB200_E_ARG, "level_bases[%zu] is null", i);
  const size_t n = size_t(1) << nv, pt = g2 ? 192 : 96;
  uint32_t *d_r0 = nullptr, *d_r1 = nullptr, *d_q = nullptr, *d_p = nullptr;
  uint4 *d_bases = nullptr, *d_proofs = nullptr;
  // level i occupies [off_i, off_i + 2^(nv-i)) of d_q / d_bases, off_i = 2^(nv+1) - 2^(nv-i+1)
  CU(cudaMallocAsync((void**)&d_r0, n * 32, g.stream));
  CU(cudaMallocAsync((void**)&d_r1, std::max<size_t>(n / 2, 1) * 32, g.stream));
  CU(cudaMallocAsync((void**)&d_q, 2 * n * 32, g.stream));
  CU(cudaMallocAsync((void**)&d_p, nv * 32, g.stream));
  CU(cudaMallocAsync((void**)&d_bases, 2 * n * pt, g.stream));
  CU(cudaMallocAsync((void**)&d_proofs, nv * pt, g.stream));
  CU(cudaMemcpyAsync(d_r0, evals, n * 32, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d_p, point, nv * 32, cudaMemcpyHostToDevice, g.stream));
  if (!(flags & TB200_SCALARS_MONT)) {
    LAUNCH(k_fr_to_mont, cdiv(n, 128), 128, g.stream, d_r0, (uint32_t)n);
    LAUNCH(k_fr_to_mont, cdiv(nv, 128), 128, g.stream, d_p, (uint32_t)nv);
  }
  // the quotient loop is a cheap sequential chain; the nv MSMs that consume it are independent of each other and
  // latency-bound (Horner chain + inversion), so they run concurrently on side streams with their own workspaces
  std::vector<size_t> off(nv);
  uint32_t *r_in = d_r0, *r_out = d_r1;
  size_t o = 0;
  for (size_t i = 0; i < nv; i++) {
    const size_t half = size_t(1) << (nv - i - 1);
    off[i] = o;
    CU(cudaMemcpyAsync((char*)d_bases + o * pt, level_bases[i], 2 * half * pt, cudaMemcpyHostToDevice, g.stream));
    LAUNCH(k_pst_level, cdiv(half, 128), 128, g.stream, r_in, (uint32_t)half, d_p + 8 * i, r_out, d_q + 8 * o);
    std::swap(r_in, r_out);
    o += 2 * half;
  }
  CU(cudaEventRecord(g.ev_join, g.stream));
  const bool prof = g.profiling;
  g.profiling = false;
  int rc = 0;
  for (size_t i = 0; i < nv && rc == 0; i++) {
    const int sl = (int)(i % g.SIDE);
    if (!g.side_stream[sl]) {
      CU(cudaStreamCreateWithFlags(&g.side_stream[sl], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&g.side_done[sl], cudaEventDisableTiming));
    }
    if (i < (size_t)g.SIDE) CU(cudaStreamWaitEvent(g.side_stream[sl], g.ev_join, 0));
    rc = msm_dev_locked((char*)d_bases + off[i] * pt, d_q + 8 * off[i], size_t(1) << (nv - i), TB200_SCALARS_MONT,
                        (char*)d_proofs + i * pt, g.side_stream[sl], nullptr, &g.side_arena[sl], false, g2);
  }
  g.profiling = prof;
  for (int sl = 0; sl < g.SIDE && sl < (int)nv; sl++) {
    if (!g.side_stream[sl]) continue;
    cudaEventRecord(g.side_done[sl], g.side_stream[sl]);
    cudaStreamWaitEvent(g.stream, g.side_done[sl], 0);
  }
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(proofs, d_proofs, nv * pt, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "proof copy failed: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(g.stream);
  }
  g.marks.clear();
  cudaFreeAsync(d_r0, g.stream);
  cudaFreeAsync(d_r1, g.stream);
  cudaFreeAsync(d_q, g.stream);
  cudaFreeAsync(d_p, g.stream);
  cudaFreeAsync(d_bases, g.stream);
  cudaFreeAsync(d_proofs, g.stream);
  return rc;
}
int tb200_pst_open_g1(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                      unsigned flags, uint64_t* proofs) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(g.device));
  return pst_open_locked(evals, nv, point, level_bases, flags, proofs, false);
}
int tb200_pst_open_g2(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                      unsigned flags, uint64_t* proofs) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(g.device));
  return pst_open_locked(evals, nv, point, level_bases, flags, proofs, true);
}

// ---- SRS / batch ----------------------------------------------------------------------------------------------------
int tb200_srs_load(const uint64_t* bases_xy, size_t n, int window_bits, tb200_srs_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || !bases_xy || n == 0 || n >= (1u << 24)) return fail(TB200_E_ARG, "bad SRS arguments (n = %zu)", n);
  CU(cudaSetDevice(g.device));
  tb200_srs* s = new tb200_srs();
  s->n = (uint32_t)n;
  s->c = (window_bits >= 3 && window_bits <= 16) ? window_bits : pick_c_batch(n);
  s->W = num_windows(s->c);
  uint4* d_b = nullptr;
  cudaError_t e = cudaMalloc((void**)&s->table, (size_t)s->W * n * 96);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_b, n * 96);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, bases_xy, n * 96, cudaMemcpyHostToDevice, g.stream);
  if (e != cudaSuccess) {
    if (s->table) cudaFree(s->table);
    if (d_b) cudaFree(d_b);
    delete s;
    return fail((int)e, "SRS upload failed: %s", cudaGetErrorString(e));
  }
  k_srs_tables<<<cdiv(n, 128), 128, 0, g.stream>>>(d_b, (uint32_t)n, s->c, s->W, s->table);
  g_launches++;
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
  cudaFree(d_b);
  if (e != cudaSuccess) {
    cudaFree(s->table);
    delete s;
    return fail((int)e, "SRS table kernel failed: %s", cudaGetErrorString(e));
  }
  *out = s;
  return 0;
}
int tb200_srs_free(tb200_srs_t srs) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!srs) return fail(TB200_E_ARG, "null SRS handle");
  if (g.ready) {
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.stream);
    cudaFree(srs->table);
  }
  delete srs;
  return 0;
}
size_t tb200_srs_size(tb200_srs_t srs) { return srs ? srs->n : 0; }

// rows are processed in chunks that respect the per-pass limits of the pipeline
static int batch_dev_locked(tb200_srs_t srs, const uint32_t* d_scalars, size_t rows, size_t cols, long long rs,
                            long long cs, unsigned flags, uint4* d_out, cudaStream_t st) {
  if (cols > srs->n) return fail(TB200_E_ARG, "cols = %zu exceeds the SRS size %u", cols, srs->n);
  if (((uintptr_t)d_scalars | (uintptr_t)d_out) & 15) return fail(TB200_E_ARG, "device pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  const uint64_t per_row = (uint64_t)std::max<size_t>(cols, 1) * srs->W;
  const uint64_t nb = 1ull << (srs->c - 1);
  uint64_t chunk = std::min<uint64_t>({(uint64_t)rows, ((1ull << 31) - 1) / per_row, ((1ull << 30)) / nb});
  // batched-affine rounds keep ~100 B of scratch per sorted entry: bound a chunk to ~2.7e8 entries (~30 GB)
  if (g.acc_mode == 2) chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(1, (1ull << 28) / per_row));
  if (chunk == 0) return fail(TB200_E_LIMIT, "a single row exceeds the per-pass limits");
  g.marks.clear();
  for (size_t r0 = 0; r0 < rows; r0 += chunk) {
    uint32_t nr = (uint32_t)std::min<uint64_t>(chunk, rows - r0);
    Plan p;
    int rc = make_plan(p, nr, (uint32_t)cols, rs, cs, srs->c, 1, flags);
    if (rc) return rc;
    // entry refs are w * cols + j and the callers guarantee cols == srs->n, the stride of the window tables
    rc = run_pipeline(p, d_scalars + 8 * (long long)r0 * rs, srs->table, d_out + 6 * r0, st);
    if (rc) return rc;
  }
  return finish_marks(st);
}

int tb200_msm_g1_batch_dev(tb200_srs_t srs, const void* d_scalars, size_t rows, size_t cols, ptrdiff_t row_stride,
                           ptrdiff_t col_stride, unsigned flags, void* d_out_xy, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!srs || (rows && (!d_out_xy || (cols && !d_scalars)))) return fail(TB200_E_ARG, "null pointer");
  if (cols != srs->n && cols != 0)
    return fail(TB200_E_ARG, "cols (%zu) must equal the SRS size (%u): window tables are laid out per SRS", cols,
                srs->n);
  CU(cudaSetDevice(g.device));
  return batch_dev_locked(srs, (const uint32_t*)d_scalars, rows, cols, row_stride, col_stride, flags,
                          (uint4*)d_out_xy, stream ? (cudaStream_t)stream : g.stream);
}

int tb200_msm_g1_batch(tb200_srs_t srs, const uint64_t* scalars, size_t rows, size_t cols, ptrdiff_t row_stride,
                       ptrdiff_t col_stride, unsigned flags, uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!srs || (rows && (!out_xy || (cols && !scalars)))) return fail(TB200_E_ARG, "null pointer");
  if (cols != srs->n && cols != 0)
    return fail(TB200_E_ARG, "cols (%zu) must equal the SRS size (%u): window tables are laid out per SRS", cols,
                srs->n);
  if (rows == 0) return 0;
  if (row_stride < 0 || col_stride < 0) return fail(TB200_E_ARG, "negative strides are not supported");
  CU(cudaSetDevice(g.device));
  // extent of the strided view in scalars
  size_t extent = cols ? (rows - 1) * (size_t)row_stride + (cols - 1) * (size_t)col_stride + 1 : 0;
  uint4 *d_s = nullptr, *d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, rows * 96, g.stream));
  if (extent) {
    CU(cudaMallocAsync((void**)&d_s, extent * 32, g.stream));
    CU(cudaMemcpyAsync(d_s, scalars, extent * 32, cudaMemcpyHostToDevice, g.stream));
  }
  int rc = batch_dev_locked(srs, (const uint32_t*)(d_s ? d_s : d_o), rows, cols, row_stride, col_stride, flags, d_o,
                            g.stream);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out_xy, d_o, rows * 96, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
  }
  if (d_s) cudaFreeAsync(d_s, g.stream);
  cudaFreeAsync(d_o, g.stream);
  return rc;
}

int tb200_msm_g1_batch_ptrs(tb200_srs_t srs, const uint64_t* const* row_ptrs, size_t rows, size_t cols,
                            unsigned flags, uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!srs || (rows && (!out_xy || !row_ptrs))) return fail(TB200_E_ARG, "null pointer");
  if (cols != srs->n && cols != 0)
    return fail(TB200_E_ARG, "cols (%zu) must equal the SRS size (%u)", cols, srs->n);
  if (rows == 0) return 0;
  CU(cudaSetDevice(g.device));
  uint4 *d_s = nullptr, *d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, rows * 96, g.stream));
  if (cols) {
    CU(cudaMallocAsync((void**)&d_s, rows * cols * 32, g.stream));
    for (size_t i = 0; i < rows; i++) {
      if (!row_ptrs[i]) {
        cudaFreeAsync(d_s, g.stream);
        cudaFreeAsync(d_o, g.stream);
        return fail(TB200_E_ARG, "row pointer %zu is null", i);
      }
      CU(cudaMemcpyAsync((char*)d_s + i * cols * 32, row_ptrs[i], cols * 32, cudaMemcpyHostToDevice, g.stream));
    }
  }
  int rc = batch_dev_locked(srs, (const uint32_t*)(d_s ? d_s : d_o), rows, cols, (long long)cols, 1, flags, d_o,
                            g.stream);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out_xy, d_o, rows * 96, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
  }
  if (d_s) cudaFreeAsync(d_s, g.stream);
  cudaFreeAsync(d_o, g.stream);
  return rc;
}

// ---- MIPP -------------------------------------------------------------------------------------------------------
int tb200_mipp_g1_begin(const uint64_t* a_xy, const uint64_t* y, size_t n, unsigned flags, tb200_mipp_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || !a_xy || !y || n == 0 || (n & (n - 1)) || n >= (1u << 28))
    return fail(TB200_E_ARG, "MIPP vectors must have a power-of-two length (got %zu)", n);
  CU(cudaSetDevice(g.device));
  tb200_mipp* h = new tb200_mipp();
  h->n = (uint32_t)n;
  h->flags = flags;
  cudaError_t e = cudaMalloc((void**)&h->a, n * 96);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->y, n * 32);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->scal, 64 * 64);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->digits, 64 * 32);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&h->scal_host, 64 * 64);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->a, a_xy, n * 96, cudaMemcpyHostToDevice, g.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->y, y, n * 32, cudaMemcpyHostToDevice, g.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
  if (e != cudaSuccess) {
    cudaFree(h->a);
    cudaFree(h->y);
    cudaFree(h->scal);
    cudaFree(h->digits);
    cudaFreeHost(h->scal_host);
    delete h;
    return fail((int)e, "MIPP upload failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return 0;
}
size_t tb200_mipp_g1_len(tb200_mipp_t h) { return h ? h->n : 0; }

int tb200_mipp_g1_cross(tb200_mipp_t h, uint64_t comm_u_l[12], uint64_t comm_u_r[12]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h || !comm_u_l || !comm_u_r) return fail(TB200_E_ARG, "null pointer");
  if (h->n < 2) return fail(TB200_E_STATE, "MIPP vectors are already folded to length 1");
  CU(cudaSetDevice(g.device));
  const uint32_t split = h->n / 2;
  // comm_u_l = MSM(a[:split], y[split:]), comm_u_r = MSM(a[split:], y[:split])   (src/mipp.rs:82-84)
  // the two MSMs are independent and latency-bound (sequential Horner + inversion tail): run them concurrently on
  // two streams with separate workspaces (the reference runs them as two rayon tasks, src/mipp.rs:77-85)
  const bool prof = g.profiling;
  g.profiling = false;
  CU(cudaEventRecord(g.ev_join, g.stream));                     // the folds are only enqueued: stream2 must follow them
  CU(cudaStreamWaitEvent(g.stream2, g.ev_join, 0));
  int rc = msm_dev_locked(h->a, h->y + 8 * (size_t)split, split, h->flags, g.d_result, g.stream, nullptr, nullptr, false);
  if (rc == 0)
    rc = msm_dev_locked(h->a + 6 * (size_t)split, h->y, split, h->flags, g.d_result + 6, g.stream2, nullptr, &g.arena2,
                        false);
  g.profiling = prof;
  if (rc) return rc;
  CU(cudaEventRecord(g.ev_join, g.stream2));
  CU(cudaStreamWaitEvent(g.stream, g.ev_join, 0));
  CU(cudaMemcpyAsync(g.h_result, g.d_result, 192, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  memcpy(comm_u_l, g.h_result, 96);
  memcpy(comm_u_r, (char*)g.h_result + 96, 96);
  return 0;
}

int tb200_mipp_g1_fold(tb200_mipp_t h, const uint64_t c[4], const uint64_t c_inv[4]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h || !c || !c_inv) return fail(TB200_E_ARG, "null pointer");
  if (h->n < 2) return fail(TB200_E_STATE, "MIPP vectors are already folded to length 1");
  CU(cudaSetDevice(g.device));
  if (h->round >= 64) return fail(TB200_E_LIMIT, "too many rounds");
  const uint32_t split = h->n / 2;
  // enqueue only: every later use of a / y is ordered behind the folds on the library's stream
  uint32_t* hs = h->scal_host + 16 * h->round;
  uint32_t* ds = h->scal + 16 * h->round;
  memcpy(hs, c, 32);
  memcpy(hs + 8, c_inv, 32);
  CU(cudaMemcpyAsync(ds, hs, 64, cudaMemcpyHostToDevice, g.stream));
  const int mont = (h->flags & TB200_SCALARS_MONT) ? 1 : 0;
  // a_l + c a_r over the G1 endomorphism (kernels_pairing.cuh): 127 doublings instead of 253
  LAUNCH(k_glv2_digits, 1, 32, g.stream, ds, mont, h->digits + 8 * h->round);
  LAUNCH(k_compress_g1_glv, cdiv(split, 128), 128, g.stream, h->a, split, h->digits + 8 * h->round);
  LAUNCH(k_compress_fr, cdiv(split, 128), 128, g.stream, h->y, split, ds + 8, mont);
  h->round++;
  h->n = split;
  return 0;
}

int tb200_mipp_g1_read(tb200_mipp_t h, uint64_t* a_xy, uint64_t* y) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h) return fail(TB200_E_ARG, "null handle");
  CU(cudaSetDevice(g.device));
  if (a_xy) CU(cudaMemcpyAsync(a_xy, h->a, (size_t)h->n * 96, cudaMemcpyDeviceToHost, g.stream));
  if (y) CU(cudaMemcpyAsync(y, h->y, (size_t)h->n * 32, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return 0;
}
int tb200_mipp_g1_end(tb200_mipp_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!h) return fail(TB200_E_ARG, "null handle");
  if (g.ready) {
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.stream);
    cudaFree(h->a);
    cudaFree(h->y);
    cudaFree(h->scal);
    cudaFree(h->digits);
    cudaFreeHost(h->scal_host);
  }
  delete h;
  return 0;
}

int tb200_mipp_g2_begin(const uint64_t* h_vec, size_t n, unsigned flags, tb200_mipp_g2_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h_vec || !out || n == 0) return fail(TB200_E_ARG, "bad arguments");
  if (n & (n - 1)) return fail(TB200_E_ARG, "MIPP vectors must have a power-of-two length (n = %zu)", n);
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "vector too long");
  CU(cudaSetDevice(g.device));
  tb200_mipp_g2* m = new tb200_mipp_g2();
  m->n = (uint32_t)n;
  m->flags = flags;
  cudaError_t e = cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking);
  cudaStream_t m_st = m->st;
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->h, n * 192);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->scal, 64 * 32);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->digits, 64 * 32);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&m->scal_host, 64 * 32);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->h, h_vec, n * 192, cudaMemcpyHostToDevice, m_st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(m_st);   // h_vec is only borrowed for the duration of the call
  if (e != cudaSuccess) {
    cudaFree(m->h);
    cudaFree(m->scal);
    cudaFree(m->digits);
    cudaFreeHost(m->scal_host);
    if (m->st) cudaStreamDestroy(m->st);
    delete m;
    return fail((int)e, "mipp_g2_begin failed: %s", cudaGetErrorString(e));
  }
  *out = m;
  return 0;
}
size_t tb200_mipp_g2_len(tb200_mipp_g2_t h) { return h ? h->n : 0; }
/* h[i] <- h[i] + c_inv * h[split + i]; returns after ENQUEUEING on the handle's own stream */
int tb200_mipp_g2_fold(tb200_mipp_g2_t h, const uint64_t c_inv[4]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h || !c_inv) return fail(TB200_E_ARG, "null pointer");
  if (h->n < 2) return fail(TB200_E_STATE, "MIPP vector is already folded to length 1");
  if (h->round >= 64) return fail(TB200_E_LIMIT, "too many rounds");
  CU(cudaSetDevice(g.device));
  const uint32_t split = h->n / 2;
  cudaStream_t m_st = h->st;
  memcpy(h->scal_host + 8 * h->round, c_inv, 32);
  CU(cudaMemcpyAsync(h->scal + 8 * h->round, h->scal_host + 8 * h->round, 32, cudaMemcpyHostToDevice, m_st));
  // 4-dimensional decomposition over the twisted Frobenius (kernels_pairing.cuh): 64 doublings instead of 253
  LAUNCH(k_glv4_digits, 1, 32, m_st, h->scal + 8 * h->round, (h->flags & TB200_SCALARS_MONT) ? 1 : 0, h->digits + 8 * h->round);
  LAUNCH(k_compress_g2_glv4w, cdiv(split, 32), 128, m_st, h->h, split, h->digits + 8 * h->round);
  h->round++;
  h->n = split;
  return 0;
}
/* the current vector (len() points); waits for the enqueued folds */
int tb200_mipp_g2_read(tb200_mipp_g2_t h, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h || !out) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(g.device));
  cudaStream_t m_st = h->st;
  CU(cudaMemcpyAsync(out, h->h, (size_t)h->n * 192, cudaMemcpyDeviceToHost, m_st));
  CU(cudaStreamSynchronize(m_st));
  return 0;
}
int tb200_mipp_g2_end(tb200_mipp_g2_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!h) return fail(TB200_E_ARG, "null handle");
  if (g.ready) {
    cudaSetDevice(g.device);
    cudaStreamSynchronize(h->st);
    cudaFree(h->h);
    cudaFree(h->scal);
    cudaFree(h->digits);
    cudaFreeHost(h->scal_host);
    cudaStreamDestroy(h->st);
  }
  delete h;
  return 0;
}

// ---- pairing products -----------------------------------------------------------------------------------------------
// Miller values of `n` pairs (g2 index = j ^ xor_mask) -> `segs` products (segment s = pairs [s n/segs, (s+1) n/segs))
// -> final exponentiation of each, written to d_out (segs x 576 B). Everything is enqueued on `st`; the scratch is
// stream-ordered. `after_miller`, if given, is recorded once the Miller kernel (the only reader of g1 / g2) is enqueued.
// `d_gt_in` != nullptr: skip the Miller stage, the n inputs are Fq12 values (per-rank Miller products to be combined).
// `final_exp` = false: stop after the product tree (a partial Miller product for a sharded pairing product).
static int pairing_products_locked(const uint4* d_g1, const uint4* d_g2, uint32_t n, uint32_t xor_mask, uint32_t segs,
                                   uint4* d_out, cudaStream_t st, cudaEvent_t after_miller,
                                   const uint4* d_gt_in = nullptr, bool final_exp = true) {
  if (n == 0) {
    LAUNCH(k_fq12_set_one, 1, 32, st, d_out, segs);
    return 0;
  }
  uint32_t len = n / segs;
  uint4 *buf_a = nullptr, *buf_b = nullptr;
  CU(cudaMallocAsync((void**)&buf_a, (size_t)n * 576, st));
  CU(cudaMallocAsync((void**)&buf_b, (size_t)segs * cdiv(len, FQ12_FAN) * 576 + 576, st));
  g.marks.clear();
  if (mark(st, "begin")) return 1;
  // below ~2 waves of resident warps one WARP per pair (latency 10 -> ~3 ms); above, one thread per pair
  const bool coop = n <= (uint32_t)g.pairing_coop_max;
  if (d_gt_in) CU(cudaMemcpyAsync(buf_a, d_gt_in, (size_t)n * 576, cudaMemcpyDeviceToDevice, st));
  else if (coop) LAUNCH(k_miller_coop, n, W12_THREADS, st, d_g1, d_g2, xor_mask, buf_a);
  else LAUNCH(k_miller, cdiv(n, 32), 32, st, d_g1, d_g2, n, xor_mask, buf_a);
  if (after_miller) CU(cudaEventRecord(after_miller, st));
  if (mark(st, "miller")) return 1;
  uint4 *cur = buf_a, *nxt = buf_b;
  while (len > 1) {
    const uint32_t m = cdiv(len, FQ12_FAN);
    if ((uint64_t)m * segs <= 4096) LAUNCH(k_fq12_prod_level_coop, dim3(m, segs), W12_THREADS, st, cur, len, m, nxt);
    else LAUNCH(k_fq12_prod_level, dim3(cdiv(m, 32), segs), 32, st, cur, len, m, nxt);
    std::swap(cur, nxt);
    len = m;
  }
  if (mark(st, "gt_product")) return 1;
  if (final_exp) LAUNCH(k_final_exp, segs, W12_THREADS, st, cur, d_out);
  else CU(cudaMemcpyAsync(d_out, cur, (size_t)segs * 576, cudaMemcpyDeviceToDevice, st));
  if (mark(st, "final_exp")) return 1;
  CU(cudaFreeAsync(buf_a, st));
  CU(cudaFreeAsync(buf_b, st));
  return finish_marks(st);
}

int tb200_multi_pairing_dev(const void* d_g1_xy, const void* d_g2, size_t n, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || (n && (!d_g1_xy || !d_g2))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "too many pairs");
  CU(cudaSetDevice(g.device));
  return pairing_products_locked((const uint4*)d_g1_xy, (const uint4*)d_g2, (uint32_t)n, 0, 1, (uint4*)d_out,
                                 stream ? (cudaStream_t)stream : g.stream, nullptr);
}

int tb200_multi_pairing(const uint64_t* g1_xy, const uint64_t* g2, size_t n, uint64_t out[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && (!g1_xy || !g2))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "too many pairs");
  CU(cudaSetDevice(g.device));
  uint4 *d_p = nullptr, *d_q = nullptr, *d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, 576, g.stream));
  if (n) {
    CU(cudaMallocAsync((void**)&d_p, n * 96, g.stream));
    CU(cudaMallocAsync((void**)&d_q, n * 192, g.stream));
    CU(cudaMemcpyAsync(d_p, g1_xy, n * 96, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_q, g2, n * 192, cudaMemcpyHostToDevice, g.stream));
  }
  int rc = pairing_products_locked(d_p, d_q, (uint32_t)n, 0, 1, d_o, g.stream, nullptr);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out, d_o, 576, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "pairing result copy failed: %s", cudaGetErrorString(e));
  }
  if (d_p) cudaFreeAsync(d_p, g.stream);
  if (d_q) cudaFreeAsync(d_q, g.stream);
  cudaFreeAsync(d_o, g.stream);
  return rc;
}

// Sharded pairing product (SURVEY.md 8e pattern: per-GPU partial, one all-gather, combine): the Miller product of a
// slice of the pairs WITHOUT the final exponentiation ...
int tb200_miller_product(const uint64_t* g1_xy, const uint64_t* g2, size_t n, uint64_t out[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && (!g1_xy || !g2))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "too many pairs");
  CU(cudaSetDevice(g.device));
  uint4 *d_p = nullptr, *d_q = nullptr, *d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, 576, g.stream));
  if (n) {
    CU(cudaMallocAsync((void**)&d_p, n * 96, g.stream));
    CU(cudaMallocAsync((void**)&d_q, n * 192, g.stream));
    CU(cudaMemcpyAsync(d_p, g1_xy, n * 96, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_q, g2, n * 192, cudaMemcpyHostToDevice, g.stream));
  }
  int rc = pairing_products_locked(d_p, d_q, (uint32_t)n, 0, 1, d_o, g.stream, nullptr, nullptr, false);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out, d_o, 576, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
  }
  if (d_p) cudaFreeAsync(d_p, g.stream);
  if (d_q) cudaFreeAsync(d_q, g.stream);
  cudaFreeAsync(d_o, g.stream);
  return rc;
}
int tb200_miller_product_dev(const void* d_g1_xy, const void* d_g2, size_t n, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || (n && (!d_g1_xy || !d_g2))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "too many pairs");
  CU(cudaSetDevice(g.device));
  return pairing_products_locked((const uint4*)d_g1_xy, (const uint4*)d_g2, (uint32_t)n, 0, 1, (uint4*)d_out,
                                 stream ? (cudaStream_t)stream : g.stream, nullptr, nullptr, false);
}
// ... and the combination: the product of `n` such partial values, then ONE final exponentiation
int tb200_gt_product_final_exp(const uint64_t* parts, size_t n, uint64_t out[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && !parts)) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 20)) return fail(TB200_E_LIMIT, "too many partial products");
  CU(cudaSetDevice(g.device));
  uint4 *d_in = nullptr, *d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, 576, g.stream));
  if (n) {
    CU(cudaMallocAsync((void**)&d_in, n * 576, g.stream));
    CU(cudaMemcpyAsync(d_in, parts, n * 576, cudaMemcpyHostToDevice, g.stream));
  }
  int rc = pairing_products_locked(nullptr, nullptr, (uint32_t)n, 0, 1, d_o, g.stream, nullptr, d_in, true);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out, d_o, 576, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
  }
  if (d_in) cudaFreeAsync(d_in, g.stream);
  cudaFreeAsync(d_o, g.stream);
  return rc;
}
int tb200_gt_product_final_exp_dev(const void* d_parts, size_t n, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || (n && !d_parts)) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 20)) return fail(TB200_E_LIMIT, "too many partial products");
  CU(cudaSetDevice(g.device));
  return pairing_products_locked(nullptr, nullptr, (uint32_t)n, 0, 1, (uint4*)d_out,
                                 stream ? (cudaStream_t)stream : g.stream, nullptr, (const uint4*)d_parts, true);
}

int tb200_mipp_pairing_cross(tb200_mipp_t a, tb200_mipp_g2_t h, uint64_t comm_t_l[72], uint64_t comm_t_r[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !h || !comm_t_l || !comm_t_r) return fail(TB200_E_ARG, "null pointer");
  if (a->n != h->n) return fail(TB200_E_ARG, "MIPP vectors differ in length (%u vs %u)", a->n, h->n);
  if (a->n < 2) return fail(TB200_E_STATE, "MIPP vectors are already folded to length 1");
  CU(cudaSetDevice(g.device));
  const uint32_t n = a->n, split = n / 2;
  // the G2 key is folded on its own stream: wait for the folds enqueued so far, and make later folds (which rewrite h
  // in place) wait for this round's Miller kernel
  CU(cudaEventRecord(g.ev_join, h->st));
  CU(cudaStreamWaitEvent(g.stream, g.ev_join, 0));
  uint4* d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, 2 * 576, g.stream));
  int rc = pairing_products_locked(a->a, h->h, n, split, 2, d_o, g.stream, g.ev_join);
  if (rc == 0) {
    cudaError_t e = cudaStreamWaitEvent(h->st, g.ev_join, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(comm_t_l, d_o, 576, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(comm_t_r, d_o + 36, 576, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "pairing result copy failed: %s", cudaGetErrorString(e));
  }
  cudaFreeAsync(d_o, g.stream);
  return rc;
}

// One MIPP round's four values in one call: the two cross MSMs (library streams) and the two cross pairing products
// (their own stream) run side by side; a single host synchronisation at the end.
int tb200_mipp_cross_all(tb200_mipp_t a, tb200_mipp_g2_t h, uint64_t comm_u_l[12], uint64_t comm_u_r[12],
                         uint64_t comm_t_l[72], uint64_t comm_t_r[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !h || !comm_u_l || !comm_u_r || !comm_t_l || !comm_t_r) return fail(TB200_E_ARG, "null pointer");
  if (a->n != h->n) return fail(TB200_E_ARG, "MIPP vectors differ in length (%u vs %u)", a->n, h->n);
  if (a->n < 2) return fail(TB200_E_STATE, "MIPP vectors are already folded to length 1");
  CU(cudaSetDevice(g.device));
  if (!g.pair_stream) {
    CU(cudaStreamCreateWithFlags(&g.pair_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&g.ev_pair, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g.ev_pair2, cudaEventDisableTiming));
  }
  const uint32_t n = a->n, split = n / 2;
  const bool prof = g.profiling;
  g.profiling = false;
  // pairing stream: behind the G1 folds (library stream) and the G2 folds (the key's stream)
  CU(cudaEventRecord(g.ev_pair, g.stream));
  CU(cudaStreamWaitEvent(g.pair_stream, g.ev_pair, 0));
  CU(cudaStreamWaitEvent(g.stream2, g.ev_pair, 0));             // the second cross MSM reads the folded a, y too
  CU(cudaEventRecord(g.ev_pair2, h->st));
  CU(cudaStreamWaitEvent(g.pair_stream, g.ev_pair2, 0));
  uint4* d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, 2 * 576, g.pair_stream));
  int rc = pairing_products_locked(a->a, h->h, n, split, 2, d_o, g.pair_stream, g.ev_pair);
  if (rc == 0) {
    cudaError_t e = cudaStreamWaitEvent(h->st, g.ev_pair, 0);   // later G2 folds rewrite h: behind this round's Miller kernel
    if (e != cudaSuccess) rc = fail((int)e, "event wait failed: %s", cudaGetErrorString(e));
  }
  // cross MSMs as in tb200_mipp_g1_cross
  if (rc == 0)
    rc = msm_dev_locked(a->a, a->y + 8 * (size_t)split, split, a->flags, g.d_result, g.stream, nullptr, nullptr, false);
  if (rc == 0)
    rc = msm_dev_locked(a->a + 6 * (size_t)split, a->y, split, a->flags, g.d_result + 6, g.stream2, nullptr, &g.arena2,
                        false);
  g.profiling = prof;
  if (rc == 0) {
    cudaError_t e = cudaEventRecord(g.ev_join, g.stream2);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g.stream, g.ev_join, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g.h_result, g.d_result, 192, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(comm_t_l, d_o, 576, cudaMemcpyDeviceToHost, g.pair_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(comm_t_r, d_o + 36, 576, cudaMemcpyDeviceToHost, g.pair_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.pair_stream);
    if (e != cudaSuccess) rc = fail((int)e, "MIPP round failed: %s", cudaGetErrorString(e));
    else {
      memcpy(comm_u_l, g.h_result, 96);
      memcpy(comm_u_r, (char*)g.h_result + 96, 96);
    }
  } else {
    cudaStreamSynchronize(g.stream);
    cudaStreamSynchronize(g.stream2);
    cudaStreamSynchronize(g.pair_stream);
  }
  cudaFreeAsync(d_o, g.pair_stream);
  return rc;
}

int tb200_gt_pow(const uint64_t* bases, const uint64_t* exps, size_t n, unsigned flags, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (n == 0) return 0;
  if (!bases || !exps || !out) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 22)) return fail(TB200_E_LIMIT, "too many elements");
  CU(cudaSetDevice(g.device));
  uint4 *d_b = nullptr, *d_o = nullptr;
  uint32_t* d_e = nullptr;
  CU(cudaMallocAsync((void**)&d_b, n * 576, g.stream));
  CU(cudaMallocAsync((void**)&d_o, n * 576, g.stream));
  CU(cudaMallocAsync((void**)&d_e, n * 32, g.stream));
  CU(cudaMemcpyAsync(d_b, bases, n * 576, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d_e, exps, n * 32, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_fq12_pow, cdiv(n, 32), 32, g.stream, d_b, d_e, (uint32_t)n, (flags & TB200_SCALARS_MONT) ? 1 : 0, d_o);
  CU(cudaMemcpyAsync(out, d_o, n * 576, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  cudaFreeAsync(d_b, g.stream);
  cudaFreeAsync(d_o, g.stream);
  cudaFreeAsync(d_e, g.stream);
  return 0;
}

int tb200_compress_g1(uint64_t* vec_xy, size_t split, const uint64_t scaler[4], unsigned flags) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!vec_xy || !scaler) return fail(TB200_E_ARG, "null pointer");
  if (split == 0) return 0;
  if (split >= (1u << 27)) return fail(TB200_E_LIMIT, "split too large");
  CU(cudaSetDevice(g.device));
  uint4* d_v = nullptr;
  uint32_t* d_k = nullptr;
  CU(cudaMallocAsync((void**)&d_v, 2 * split * 96, g.stream));
  CU(cudaMallocAsync((void**)&d_k, 32, g.stream));
  CU(cudaMemcpyAsync(d_v, vec_xy, 2 * split * 96, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d_k, scaler, 32, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_compress_g1, cdiv(split, 128), 128, g.stream, d_v, (uint32_t)split, d_k,
         (flags & TB200_SCALARS_MONT) ? 1 : 0);
  CU(cudaMemcpyAsync(vec_xy, d_v, split * 96, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  cudaFreeAsync(d_v, g.stream);
  cudaFreeAsync(d_k, g.stream);
  return 0;
}

// ---- device buffers for hosts without a CUDA runtime of their own (the Rust/C++ side keeps Z resident) -----------------
int tb200_dev_alloc(size_t bytes, void** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || bytes == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(g.device));
  CU(cudaMalloc(out, bytes));
  return 0;
}
int tb200_dev_free(void* p) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(g.device));
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaFree(p));
  return 0;
}
int tb200_dev_upload(void* d_dst, const void* h_src, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_dst || !h_src) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(g.device));
  CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return 0;
}
int tb200_dev_download(void* h_dst, const void* d_src, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h_dst || !d_src) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(g.device));
  CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return 0;
}
int tb200_stream_sync(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(g.device));
  CU(cudaStreamSynchronize(g.stream));
  return 0;
}

// ---- sqrt_pst scalar work on the device -------------------------------------------------------------------------------
int tb200_fr_chis(const uint64_t* b, size_t m, uint64_t* chis_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!chis_out || (m && !b) || m > 28) return fail(TB200_E_ARG, "bad arguments (m = %zu)", m);
  CU(cudaSetDevice(g.device));
  const size_t n = size_t(1) << m;
  uint32_t *d_b = nullptr, *d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_b, std::max<size_t>(m, 1) * 32, g.stream));
  CU(cudaMallocAsync((void**)&d_o, n * 32, g.stream));
  if (m) CU(cudaMemcpyAsync(d_b, b, m * 32, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_fr_chis, cdiv(n, 128), 128, g.stream, d_b, (uint32_t)m, d_o);
  CU(cudaMemcpyAsync(chis_out, d_o, n * 32, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  cudaFreeAsync(d_b, g.stream);
  cudaFreeAsync(d_o, g.stream);
  return 0;
}
int tb200_fr_matvec_dev(const void* d_Z, size_t rows, size_t cols, const void* d_v, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || ((rows && cols) && (!d_Z || !d_v))) return fail(TB200_E_ARG, "null pointer");
  if (rows >= (1ull << 31) || cols >= (1ull << 31)) return fail(TB200_E_LIMIT, "matrix too large");
  if (((uintptr_t)d_Z | (uintptr_t)d_v | (uintptr_t)d_out) & 15) return fail(TB200_E_ARG, "device pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  CU(cudaSetDevice(g.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  LAUNCH(k_fr_matvec, cdiv(rows * 32, 256), 256, st, (const uint32_t*)d_Z, (uint32_t)rows, (uint32_t)cols,
         (const uint32_t*)d_v, (uint32_t*)d_out);
  return 0;
}
int tb200_fr_matvec(const uint64_t* Z, size_t rows, size_t cols, const uint64_t* v, uint64_t* out) {
  uint32_t *d_z = nullptr, *d_v = nullptr, *d_o = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (need_ready()) return TB200_E_STATE;
    if (!out || !Z || !v || rows == 0 || cols == 0) return fail(TB200_E_ARG, "bad arguments");
    CU(cudaSetDevice(g.device));
    CU(cudaMallocAsync((void**)&d_z, rows * cols * 32, g.stream));
    CU(cudaMallocAsync((void**)&d_v, cols * 32, g.stream));
    CU(cudaMallocAsync((void**)&d_o, rows * 32, g.stream));
    CU(cudaMemcpyAsync(d_z, Z, rows * cols * 32, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_v, v, cols * 32, cudaMemcpyHostToDevice, g.stream));
  }
  int rc = tb200_fr_matvec_dev(d_z, rows, cols, d_v, d_o, nullptr);
  std::lock_guard<std::mutex> lk(g_mu);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out, d_o, rows * 32, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
  }
  cudaFreeAsync(d_z, g.stream);
  cudaFreeAsync(d_v, g.stream);
  cudaFreeAsync(d_o, g.stream);
  return rc;
}

// ---- group utilities ---------------------------------------------------------------------------------------------
int tb200_g1_sum_dev(const void* d_pts_xy, size_t n, void* d_out_xy, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out_xy || (n && !d_pts_xy)) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 20)) return fail(TB200_E_LIMIT, "tb200_g1_sum is meant for a handful of partial results");
  CU(cudaSetDevice(g.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  LAUNCH(k_g1_sum, 1, 32, st, (const uint4*)d_pts_xy, (uint32_t)n, (uint4*)d_out_xy);
  return 0;
}
int tb200_g1_sum(const uint64_t* pts_xy, size_t n, uint64_t out_xy[12]) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (need_ready()) return TB200_E_STATE;
    if (!out_xy || (n && !pts_xy)) return fail(TB200_E_ARG, "null pointer");
    if (n > 128) return fail(TB200_E_LIMIT, "tb200_g1_sum (host form) takes at most 128 points");
    CU(cudaSetDevice(g.device));
    if (n) CU(cudaMemcpyAsync(g.d_result + 6, pts_xy, n * 96, cudaMemcpyHostToDevice, g.stream));
    LAUNCH(k_g1_sum, 1, 32, g.stream, g.d_result + 6, (uint32_t)n, g.d_result);
    CU(cudaMemcpyAsync(g.h_result, g.d_result, 96, cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    memcpy(out_xy, g.h_result, 96);
  }
  return 0;
}
int tb200_g1_outer_sum_dev(const void* d_a_xy, size_t na, const void* d_b_xy, size_t nb, void* d_out_xy,
                           void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_a_xy || !d_b_xy || !d_out_xy || na == 0 || nb == 0) return fail(TB200_E_ARG, "bad arguments");
  if ((uint64_t)na * nb >= (1ull << 31)) return fail(TB200_E_LIMIT, "outer sum too large");
  CU(cudaSetDevice(g.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  LAUNCH(k_g1_outer_sum, cdiv((uint64_t)na * nb, 128), 128, st, (const uint4*)d_a_xy, (uint32_t)na,
         (const uint4*)d_b_xy, (uint32_t)nb, (uint4*)d_out_xy);
  return 0;
}

// ---- microbenchmarks / unit-test hooks ------------------------------------------------------------------------------
int tb200_int_pipe_peak(int kind, int iters, double* out_per_s) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out_per_s || iters <= 0 || kind < 0 || kind > 2) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(g.device));
  const int threads = kind == 2 ? 128 : 256;
  const int blocks = g.sms * (kind == 2 ? 3 : 8);
  void* sink = nullptr;
  CU(cudaMalloc(&sink, (size_t)blocks * threads * 8));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {  // first repetition warms up
    CU(cudaEventRecord(e0, g.stream));
    if (kind == 2) LAUNCH(k_fq_mul_peak, blocks, threads, g.stream, iters, 12345u, (uint32_t*)sink);
    else LAUNCH(k_int_pipe, blocks, threads, g.stream, kind, iters, 12345u, (uint64_t*)sink);
    CU(cudaEventRecord(e1, g.stream));
    CU(cudaStreamSynchronize(g.stream));
  }
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  double ops = (double)blocks * threads * (double)iters * (kind == 2 ? 2.0 : 64.0);
  *out_per_s = ops / (ms * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

}  // extern "C"

template <class F>
static int with_buffers(const void* a, size_t abytes, const void* b, size_t bbytes, void* o1, size_t o1bytes,
                        void* o2, size_t o2bytes, F&& launch) {
  char *d_a = nullptr, *d_b = nullptr, *d_o1 = nullptr, *d_o2 = nullptr;
  CU(cudaMalloc((void**)&d_a, std::max<size_t>(abytes, 16)));
  CU(cudaMalloc((void**)&d_b, std::max<size_t>(bbytes, 16)));
  CU(cudaMalloc((void**)&d_o1, std::max<size_t>(o1bytes, 16)));
  CU(cudaMalloc((void**)&d_o2, std::max<size_t>(o2bytes, 16)));
  CU(cudaMemcpy(d_a, a, abytes, cudaMemcpyHostToDevice));
  if (b) CU(cudaMemcpy(d_b, b, bbytes, cudaMemcpyHostToDevice));
  int rc = launch(d_a, d_b, d_o1, d_o2);
  if (rc == 0) {
    CU(cudaStreamSynchronize(g.stream));
    CU(cudaMemcpy(o1, d_o1, o1bytes, cudaMemcpyDeviceToHost));
    if (o2) CU(cudaMemcpy(o2, d_o2, o2bytes, cudaMemcpyDeviceToHost));
  }
  cudaFree(d_a);
  cudaFree(d_b);
  cudaFree(d_o1);
  cudaFree(d_o2);
  return rc;
}

extern "C" {

int tb200_test_fq_mul(const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !b || !out || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(g.device));
  return with_buffers(a, n * 48, b, n * 48, out, n * 48, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    LAUNCH(k_test_fq_mul, cdiv(n, 128), 128, g.stream, (const uint32_t*)da, (const uint32_t*)db, (uint32_t)n,
           (uint32_t*)d1);
    return 0;
  });
}
int tb200_test_fq_addsub(const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out_add, uint64_t* out_sub) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !b || !out_add || !out_sub || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(g.device));
  return with_buffers(a, n * 48, b, n * 48, out_add, n * 48, out_sub, n * 48,
                      [&](char* da, char* db, char* d1, char* d2) {
                        LAUNCH(k_test_fq_addsub, cdiv(n, 128), 128, g.stream, (const uint32_t*)da,
                               (const uint32_t*)db, (uint32_t)n, (uint32_t*)d1, (uint32_t*)d2);
                        return 0;
                      });
}
int tb200_test_g1_add(const uint64_t* p_xy, const uint64_t* q_xy, size_t n, uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!p_xy || !q_xy || !out_xy || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(g.device));
  return with_buffers(p_xy, n * 96, q_xy, n * 96, out_xy, n * 96, nullptr, 0,
                      [&](char* da, char* db, char* d1, char*) {
                        LAUNCH(k_test_g1_add, cdiv(n, 64), 64, g.stream, (const uint4*)da, (const uint4*)db,
                               (uint32_t)n, (uint4*)d1);
                        return 0;
                      });
}
int tb200_test_g2_add(const uint64_t* p, const uint64_t* q, size_t n, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!p || !q || !out || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(g.device));
  return with_buffers(p, n * 192, q, n * 192, out, n * 192, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    LAUNCH(k_test_g2_add, cdiv(n, 64), 64, g.stream, (const uint4*)da, (const uint4*)db, (uint32_t)n, (uint4*)d1);
    return 0;
  });
}
int tb200_test_g2_mul(const uint64_t* p, const uint64_t* k, size_t n, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!p || !k || !out || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(g.device));
  return with_buffers(p, n * 192, k, n * 32, out, n * 192, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    LAUNCH(k_test_g2_mul, cdiv(n, 64), 64, g.stream, (const uint4*)da, (const uint32_t*)db, (uint32_t)n, (uint4*)d1);
    return 0;
  });
}
int tb200_test_g1_mul(const uint64_t* p_xy, const uint64_t* k, size_t n, uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!p_xy || !k || !out_xy || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(g.device));
  return with_buffers(p_xy, n * 96, k, n * 32, out_xy, n * 96, nullptr, 0,
                      [&](char* da, char* db, char* d1, char*) {
                        LAUNCH(k_test_g1_mul, cdiv(n, 64), 64, g.stream, (const uint4*)da, (const uint32_t*)db,
                               (uint32_t)n, (uint4*)d1);
                        return 0;
                      });
}

int tb200_test_fq12_op(int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !b || !out || n == 0 || op < 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(g.device));
  return with_buffers(a, n * 576, b, n * 576, out, n * 576, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    if (op >= 20 && op < 100)
      LAUNCH(k_test_w12_op, (uint32_t)n, W12_THREADS, g.stream, op, (const uint4*)da, (const uint4*)db, (uint4*)d1);
    else
      LAUNCH(k_test_fq12_op, cdiv(n, 32), 32, g.stream, op, (const uint4*)da, (const uint4*)db, (uint32_t)n, (uint4*)d1);
    return 0;
  });
}

}  // extern "C"
