// G1 unit of the engine: the kernels of kernels.cuh / kernels_smem.cuh and everything that launches
// them -- the MSM pipeline (digits -> scan -> scatter -> accumulate -> fix-up -> reduce -> finalize), the shared-base
// batch, the SRS window tables, the device-resident MIPP G1 loop, the PST quotient loop and the Fr helpers of sqrt_pst.
// There is no CPU arithmetic path in this file (nor anywhere in the library).
#include <algorithm>

#include "engine.h"
#include "glv_host.h"
#include "kernels_smem.cuh"
#include "kernels_small.cuh"

using namespace tb;

namespace tbe {

// ---- window selection ---------------------------------------------------------------------------------------
// cost in mixed-addition units: accumulation entries + ~4 adds per bucket for the hierarchical reduction
static int pick_c_single(uint64_t n) {
  if (E.forced_c) return E.forced_c;
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
};

// a single MSM processed as point-range chunks that accumulate into ONE persistent bucket array
struct ChunkCtl {
  uint32_t ref_base;  // global index of the chunk's first point
  uint4* buckets;     // B * 192 B, all zero (= identity) before the first chunk
  bool last;          // run the reduction / finalisation after this chunk
};

static int make_plan(const Ctx& g, Plan& p, uint32_t rows, uint32_t cols, long long rs, long long cs, int c, int batch,
                     unsigned flags, bool g2 = false, int w_lo = 0, int w_hi = -1) {
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
  q.w_lo = w_lo;
  q.w_hi = w_hi < 0 ? q.W : w_hi;
  q.groups = batch ? rows : (uint32_t)(q.w_hi - q.w_lo);
  q.mont = (flags & TB200_SCALARS_MONT) ? 1 : 0;
  p.M_max = (uint64_t)rows * cols * (batch ? q.W : (q.w_hi - q.w_lo));
  p.B = (uint64_t)q.groups * q.nb;
  if (p.M_max > E.pass_entries_max || p.B >= (1ull << 31))
    return fail(TB200_E_LIMIT, "MSM too large for one pass: %llu entries, %llu buckets", (unsigned long long)p.M_max,
                (unsigned long long)p.B);
  uint64_t target_threads = (uint64_t)g.sms * 3 * ACCS_THREADS * 4;
  uint64_t K = (p.M_max + target_threads - 1) / target_threads;
  p.K = (uint32_t)std::min<uint64_t>(256, std::max<uint64_t>(4, K));
  p.S_max = cdiv(std::max<uint64_t>(p.M_max, 1), p.K);
  p.ntiles = (uint32_t)(p.B / SCAN_TILE + 1);
  p.Ls.clear();
  // Fan-in 32 at level 0 (throughput-bound: millions of buckets). The levels above it of a SINGLE MSM hold few
  // elements and are latency-bound (2L - 1 sequential additions + log2(ell) doublings per thread): fan-in 8 there
  // (measured at 2^24, c = 20: 3.5 -> 1.9 ms for the upper levels). Batches keep 32: thousands of rows fill the GPU.
  // Small single MSMs (G2 always: sqrt(n)-sized, ~50 us per addition) are latency-bound at level 0 as well: with
  // fan-in 8 everywhere a 2^13-point G2 MSM reduces in 54 sequential group operations instead of 99.
  const bool small_single = !batch && p.B < (1ull << 17);
  for (uint32_t n = q.nb; n > 1;) {
    uint32_t L = std::min<uint32_t>(n, (batch || (p.Ls.empty() && !small_single)) ? 32 : 8);
    p.Ls.push_back(L);
    n /= L;
  }
  size_t b = 0;
  b += Arena::pad((p.B + 1) * 4) * 2;  // counts, starts
  b += Arena::pad(p.B * 4);            // cursors
  b += Arena::pad((size_t)p.ntiles * 4 + 4);
  b += Arena::pad(std::max<uint64_t>(p.M_max, 1) * 4);  // entries
  b += Arena::pad(p.B * 192 * pw);                      // buckets
  b += Arena::pad((size_t)p.S_max * 192 * pw) + Arena::pad((size_t)p.S_max * 4) + Arena::pad(64 * 4);
  uint64_t n = p.B;
  for (uint32_t L : p.Ls) {
    n /= L;
    b += Arena::pad(n * 192 * pw) * 2;
  }
  b += Arena::pad((size_t)q.groups * 192 * pw) * 2 + 4096;
  if (g2) b += Arena::pad((size_t)4 * 96 * 384);  // window-combine partial sums (k_finalize_single_g2_glv)
  p.bytes = b;
  return 0;
}

// The pipeline in four phases, so that callers can put them on different streams (msm_dev: window ranges side by side;
// msm_host_enqueue: the sort of chunk k+1 next to the accumulation of chunk k). `arena` holds the phase's scratch; the
// caller has acquired it for `st` (Arena::acquire) and releases it after the last phase that uses the buffers.
struct PipeBufs {
  uint32_t *counts = nullptr, *starts = nullptr, *cursors = nullptr, *tile_sums = nullptr, *entries = nullptr;
  uint4 *buckets = nullptr, *heads = nullptr;
  int32_t* head_bucket = nullptr;
  uint32_t* need = nullptr;   // 64 words: which fix-up rounds have work (k_fixup_round)
};

// digits -> scan -> scatter. Needs only the scalars.
static int pipe_sort(Ctx& g, const Plan& p, const MsmGeom& q, const uint32_t* d_scalars, cudaStream_t st, Arena& arena,
                     const ChunkCtl* chunk, PipeBufs& b) {
  const size_t pw = p.g2 ? 2 : 1;
  int rc = arena.reserve(p.bytes);
  if (rc) return fail(rc, "workspace allocation of %zu bytes failed: %s", p.bytes, cudaGetErrorString((cudaError_t)rc));
  arena.reset();
  b.counts = arena.take<uint32_t>(p.B + 1);
  b.starts = arena.take<uint32_t>(p.B + 1);
  b.cursors = arena.take<uint32_t>(p.B);
  b.tile_sums = arena.take<uint32_t>(p.ntiles + 1);
  b.entries = arena.take<uint32_t>(std::max<uint64_t>(p.M_max, 1));
  b.buckets = chunk ? chunk->buckets : arena.take<uint4>(p.B * 12 * pw);
  b.heads = arena.take<uint4>((size_t)p.S_max * 12 * pw);
  b.head_bucket = arena.take<int32_t>(p.S_max);
  b.need = arena.take<uint32_t>(64);
  const uint64_t items = (uint64_t)q.rows * q.cols;
  if (mark(g, st, "begin")) return 1;
  const uint32_t dig_grid = (uint32_t)std::min<uint64_t>(cdiv(items, 256), (uint64_t)g.sms * 16);
  // batches with enough rows to fill the GPU sort each row inside one CTA's shared memory
  const size_t row_smem = (size_t)q.nb * 4;
  const bool row_sort = q.batch && q.rows >= (uint32_t)g.sms && row_smem <= 160 * 1024;
  // >= 116 KB of dynamic shared memory => one CTA per SM (see k_batch_digits)
  const size_t row_smem_launch = std::max<size_t>(row_smem, 116 * 1024);
  if (row_sort) {
    CU(cudaFuncSetAttribute(k_batch_digits<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem_launch));
    CU(cudaFuncSetAttribute(k_batch_digits<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem_launch));
    LAUNCH_SMEM(k_batch_digits<false>, q.rows, 1024, row_smem_launch, st, d_scalars, q, b.counts, (uint32_t*)nullptr);
  } else {
    CU(cudaMemsetAsync(b.counts, 0, (p.B + 1) * 4, st));
    LAUNCH(k_digits<false>, dig_grid, 256, st, d_scalars, q, b.counts, (uint32_t*)nullptr, -1);
  }
  if (mark(g, st, "digits")) return 1;
  LAUNCH(k_scan_tile_sums, p.ntiles, SCAN_THREADS, st, b.counts, (uint32_t)p.B, b.tile_sums);
  LAUNCH(k_scan_tile_offsets, 1, SCAN_THREADS, st, b.tile_sums, p.ntiles, b.tile_sums + p.ntiles);
  LAUNCH(k_scan_apply, p.ntiles, SCAN_THREADS, st, b.counts, (uint32_t)p.B, b.tile_sums, b.starts, b.cursors);
  if (mark(g, st, "scan")) return 1;
  if (row_sort) {
    LAUNCH_SMEM(k_batch_digits<true>, q.rows, 1024, row_smem_launch, st, d_scalars, q, b.starts, b.entries);
  } else {
    // large single MSMs: one pass per window keeps the writes of a pass inside an L2-sized slice of entries[]
    const bool per_window = !q.batch && q.c >= 19 && (uint64_t)q.cols * 4 * q.W > (64ull << 20);
    if (per_window) {
      for (int w = q.w_lo; w < q.w_hi; w++) LAUNCH(k_digits<true>, dig_grid, 256, st, d_scalars, q, b.cursors, b.entries, w);
    } else {
      LAUNCH(k_digits<true>, dig_grid, 256, st, d_scalars, q, b.cursors, b.entries, -1);
    }
  }
  if (mark(g, st, "scatter")) return 1;
  return 0;
}

// accumulate -> fix-up. Needs the points (`points_ready`, if given, is awaited first).
static int pipe_accumulate(Ctx& g, const Plan& p, const MsmGeom& q, const PipeBufs& b, const uint4* d_points, cudaStream_t st,
                           cudaEvent_t points_ready, const ChunkCtl* chunk) {
  if (points_ready) CU(cudaStreamWaitEvent(st, points_ready, 0));  // bases may still be in flight until here
  // M is only known on the device (starts[B]); launch for the upper bound, surplus threads exit immediately
  if (p.g2) {
    if (int r = g2_accumulate(st, p.S_max, b.entries, b.starts, (uint32_t)p.B, p.K, d_points, b.buckets, b.heads, b.head_bucket))
      return r;
  } else {
    // operands in shared-memory slots (kernels_smem.cuh). acc_mode 3: plain CIOS products; 0 / 4: Y3 as one fused sum
    // of two products (default)
    const dim3 grid(cdiv(p.S_max, ACCS_THREADS));
    if (E.acc_mode == 3) {
      CU(cudaFuncSetAttribute(k_accumulate_s<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACCS_SMEM));
      LAUNCH_SMEM(k_accumulate_s<0>, grid, ACCS_THREADS, ACCS_SMEM, st, b.entries, b.starts, (uint32_t)p.B, p.K, d_points,
                  b.buckets, b.heads, b.head_bucket, chunk ? 1 : 0);
    } else {
      CU(cudaFuncSetAttribute(k_accumulate_s<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACCS_SMEM));
      LAUNCH_SMEM(k_accumulate_s<2>, grid, ACCS_THREADS, ACCS_SMEM, st, b.entries, b.starts, (uint32_t)p.B, p.K, d_points,
                  b.buckets, b.heads, b.head_bucket, chunk ? 1 : 0);
    }
  }
  if (mark(g, st, "accumulate")) return 1;
  // A bucket holds at most one entry per (column, window) of its group: cols entries (single MSM: one window per
  // group) or cols * W (batch row), i.e. it spans at most that many / K + 1 segments.
  const uint64_t max_bucket = q.batch ? (uint64_t)q.cols * q.W : (uint64_t)q.cols;
  const uint64_t max_span = std::min<uint64_t>(p.S_max, max_bucket / p.K + 2);
  if (p.g2) {
    for (uint32_t round = 0; (1ull << round) < max_span; round++)
      if (int r = g2_fixup_round(st, p.S_max, b.starts, (uint32_t)p.B, p.K, round, b.heads, b.head_bucket)) return r;
    if (int r = g2_fixup_final(st, p.S_max, b.starts, (uint32_t)p.B, p.K, b.buckets, b.heads, b.head_bucket)) return r;
  } else {
    CU(cudaMemsetAsync(b.need, 0, 64 * 4, st));
    for (uint32_t round = 0; (1ull << round) < max_span && round < 62; round++)
      LAUNCH(k_fixup_round, cdiv(p.S_max, 128), 128, st, b.starts, (uint32_t)p.B, p.K, round, b.heads, b.head_bucket, b.need);
    LAUNCH(k_fixup_final, cdiv(p.S_max, 128), 128, st, b.starts, (uint32_t)p.B, p.K, b.buckets, b.heads, b.head_bucket);
  }
  if (mark(g, st, "fixup")) return 1;
  return 0;
}

// hierarchical bucket reduction: one (S, W) pair per group; *group_w = the groups' sums (XYZZ, q.groups entries)
static int pipe_reduce(Ctx& g, const Plan& p, const MsmGeom& q, const PipeBufs& b, cudaStream_t st, Arena& arena,
                       const ChunkCtl* chunk, const uint4** group_w) {
  const size_t pw = p.g2 ? 2 : 1;
  // chunked runs: emptiness is per chunk; the persistent buckets carry the identity (all zero) instead
  const uint4 *inS = b.buckets, *inW = nullptr;
  uint64_t n = p.B;
  int log2_ell = 0;
  const uint32_t* level0 = chunk ? nullptr : b.starts;
  for (size_t li = 0; li < p.Ls.size(); li++) {
    const uint32_t L = p.Ls[li];
    n /= L;
    uint4* outS = arena.take<uint4>(n * 12 * pw);
    uint4* outW = arena.take<uint4>(n * 12 * pw);
    if (p.g2) {
      if (int r = g2_reduce_pass(st, inS, inW, level0, outS, outW, L, log2_ell, n)) return r;
    } else {
      LAUNCH(k_reduce_pass, cdiv(n, 128), 128, st, inS, inW, level0, outS, outW, L, log2_ell, n);
    }
    inS = outS;
    inW = outW;
    level0 = nullptr;
    for (uint32_t v = L; v > 1; v >>= 1) log2_ell++;
  }
  *group_w = inW;  // nb >= 4 (c >= 3): at least one reduction level has run
  (void)q;
  if (mark(g, st, "reduce")) return 1;
  return 0;
}

// window combine (single MSM: `W` window sums at group_w) or per-row normalisation (batch)
static int pipe_finalize(Ctx& g, const Plan& p, const MsmGeom& q, const uint4* group_w, uint4* d_out, cudaStream_t st,
                         Arena& arena) {
  if (p.g2 && q.W <= 96) {
    // window combine over the twisted Frobenius: 4 W parallel 64-doubling chains + a tree instead of ~250 serial doublings
    uint4* fin = arena.take<uint4>((size_t)4 * q.W * 24);
    if (int r = g2_finalize_single_glv(st, group_w, q.W, q.c, fin, d_out)) return r;
  } else if (p.g2) {
    if (int r = g2_finalize_single(st, group_w, q.W, q.c, d_out)) return r;
  } else if (q.batch) {
    LAUNCH(k_finalize_batch, cdiv(q.groups, 128), 128, st, group_w, q.groups, d_out);
  } else {
    LAUNCH(k_finalize_single, 1, 32, st, group_w, q.W, q.c, d_out);
  }
  if (mark(g, st, "finalize")) return 1;
  return 0;
}

static void note_geometry(Ctx& g, const Plan& p, uint64_t entries, uint64_t buckets) {
  g.last_c = p.geo.c;
  g.last_W = p.geo.W;
  g.last_K = (int)p.K;
  g.last_entries = entries;
  g.last_buckets = buckets;
}

// Runs the whole pipeline on `st`. d_points: affine points indexed by entry refs. d_out: groups*96 B (batch) or 96 B.
// The caller has acquired `arena` for `st` (Arena::acquire) and releases it afterwards.
static int run_pipeline(Ctx& g, const Plan& p, const uint32_t* d_scalars, const uint4* d_points, uint4* d_out,
                        cudaStream_t st, cudaEvent_t points_ready, Arena& arena, const ChunkCtl* chunk = nullptr) {
  MsmGeom q = p.geo;
  if (chunk) q.ref_base = chunk->ref_base;
  const size_t pw = p.g2 ? 2 : 1;
  note_geometry(g, p, p.M_max, p.B);
  const uint64_t items = (uint64_t)q.rows * q.cols;
  if (items == 0) {
    uint32_t cnt = q.batch ? q.rows : 1;
    if (cnt) LAUNCH(k_write_identity, cdiv(cnt * 6 * pw, 128), 128, st, d_out, (uint32_t)(cnt * pw));
    return 0;
  }
  PipeBufs b;
  if (int rc = pipe_sort(g, p, q, d_scalars, st, arena, chunk, b)) return rc;
  if (int rc = pipe_accumulate(g, p, q, b, d_points, st, points_ready, chunk)) return rc;
  if (chunk && !chunk->last) return 0;  // later point-range chunks continue in the same buckets
  const uint4* group_w = nullptr;
  if (int rc = pipe_reduce(g, p, q, b, st, arena, chunk, &group_w)) return rc;
  return pipe_finalize(g, p, q, group_w, d_out, st, arena);
}

// ---- single MSM over device pointers ------------------------------------------------------------------------------
int msm_dev(Ctx& g, const void* d_bases, const void* d_scalars, size_t n, unsigned flags, void* d_out, cudaStream_t st,
            cudaEvent_t points_ready, Arena* arena_p, bool finish, bool g2) {
  if (n >= (1ull << 31)) return fail(TB200_E_LIMIT, "n = %zu exceeds 2^31 - 1 points per call", n);
  if (((uintptr_t)d_bases | (uintptr_t)d_scalars | (uintptr_t)d_out) & 15)
    return fail(TB200_E_ARG, "device pointers must be 16-byte aligned");
  g.marks.clear();
  if (!g2 && n >= 1 && n <= (size_t)E.small_msm_max && !E.forced_c) {   // a forced window width means: the pipeline
    // small-n fast path (kernels_small.cuh): a quad of lanes per point, 8 points per one-warp CTA, no sort pipeline
    const uint32_t nblk = cdiv(n, SMALL_QUADS);
    uint4* scratch = nullptr;
    if (nblk > 1) {
      CU(cudaMallocAsync((void**)&scratch, ((size_t)nblk * 12 + 1) * 16, st));
      CU(cudaMemsetAsync(scratch + (size_t)nblk * 12, 0, 16, st));
    }
    if (points_ready) CU(cudaStreamWaitEvent(st, points_ready, 0));
    if (mark(g, st, "begin")) return 1;
    LAUNCH(k_msm_small, nblk, 4 * SMALL_QUADS, st, (const uint4*)d_bases, (const uint32_t*)d_scalars, (uint32_t)n,
           (flags & TB200_SCALARS_MONT) ? 1 : 0, scratch, (uint4*)d_out, 0u, (const uint32_t*)nullptr);
    if (scratch) CU(cudaFreeAsync(scratch, st));
    if (mark(g, st, "accumulate")) return 1;
    g.last_c = 4;
    g.last_W = SMALL_DIGITS;
    g.last_K = 0;
    g.last_entries = n * SMALL_DIGITS;
    g.last_buckets = 0;
    return finish ? finish_marks(g, st) : 0;
  }
  Arena& arena = arena_p ? *arena_p : g.arena;
  const int c = pick_c_single(std::max<size_t>(n, 1));
  const uint64_t W = (uint64_t)num_windows(c);
  const uint64_t pass_pts = std::max<uint64_t>(E.pass_entries_max / W, 1);
  int rc = arena.acquire(st);
  if (rc) return rc;
  const bool overlap = E.msm_overlap && !g2 && !arena_p && !g.profiling && n >= (size_t(1) << 20) && n <= pass_pts && W >= 6;
  if (overlap) {
    // Three window ranges on three streams. The sort stages (HBM / L2 bound) of a range and the latency-bound tail of
    // its reduction hide behind another range's accumulation (integer-pipe bound); the first range is ONE window so
    // that an accumulation is running ~2 ms into the call. The side streams have high priority: their short CTAs take
    // the SM slots an accumulation frees before its own next wave does.
    const int Wn = (int)W, cut[4] = {0, 1, 1 + (Wn - 1) / 2, Wn};
    cudaStream_t ss[3] = {st, g.split_stream[0], g.split_stream[1]};
    Arena* as[3] = {&arena, &g.split_arena[0], &g.split_arena[1]};
    uint4* gw_all = nullptr;
    CU(cudaMallocAsync((void**)&gw_all, (size_t)Wn * 192, st));
    CU(cudaEventRecord(g.ev_split[0], st));              // inputs of this call are ordered behind st's earlier work
    Plan plans[3];
    for (int r = 0; r < 3 && rc == 0; r++) {
      rc = make_plan(g, plans[r], 1, (uint32_t)n, 0, 1, c, 0, flags, false, cut[r], cut[r + 1]);
      if (rc) break;
      if (r > 0) {
        CU(cudaStreamWaitEvent(ss[r], g.ev_split[0], 0));
        rc = as[r]->acquire(ss[r]);
        if (rc) break;
      }
      const MsmGeom q = plans[r].geo;
      PipeBufs b;
      // staggered: range r is sorted WHILE range r-1 accumulates (its sort waits for the sort of r-1 to finish, i.e.
      // for that accumulation to start); started together, all three sorts would run up front with nothing to hide behind
      if (r > 0) CU(cudaStreamWaitEvent(ss[r], g.ev_split[3], 0));
      rc = pipe_sort(g, plans[r], q, (const uint32_t*)d_scalars, ss[r], *as[r], nullptr, b);
      if (rc == 0 && r < 2) CU(cudaEventRecord(g.ev_split[3], ss[r]));
      if (rc == 0) rc = pipe_accumulate(g, plans[r], q, b, (const uint4*)d_bases, ss[r], points_ready, nullptr);
      const uint4* gw = nullptr;
      if (rc == 0) rc = pipe_reduce(g, plans[r], q, b, ss[r], *as[r], nullptr, &gw);
      if (rc) break;
      CU(cudaMemcpyAsync(gw_all + 12 * (size_t)cut[r], gw, (size_t)q.groups * 192, cudaMemcpyDeviceToDevice, ss[r]));
      if (r > 0) {
        rc = as[r]->release(ss[r]);
        if (rc) break;
        CU(cudaEventRecord(g.ev_split[r], ss[r]));
        CU(cudaStreamWaitEvent(st, g.ev_split[r], 0));
      }
    }
    if (rc == 0) {
      note_geometry(g, plans[1], (uint64_t)n * W, W << (c - 1));
      rc = pipe_finalize(g, plans[0], plans[0].geo, gw_all, (uint4*)d_out, st, arena);
    }
    cudaFreeAsync(gw_all, st);
  } else if (n <= pass_pts) {
    Plan p;
    rc = make_plan(g, p, 1, (uint32_t)n, 0, 1, c, 0, flags, g2);
    if (rc) return rc;
    rc = run_pipeline(g, p, (const uint32_t*)d_scalars, (const uint4*)d_bases, (uint4*)d_out, st, points_ready, arena);
  } else {
    // more sorted entries than one pass can index (n W > 2^32): point-range passes over ONE persistent bucket array
    // (the accumulate kernel continues from the stored bucket), reduction / finalisation once after the last pass
    if (g2) return fail(TB200_E_LIMIT, "G2 MSM of %zu points exceeds the per-pass entry limit", n);
    const uint64_t cap = pass_pts >= 32 ? (pass_pts & ~31ull) : pass_pts;  // 32-aligned pass boundaries
    const uint64_t npass = (n + cap - 1) / cap;
    const uint64_t per = std::min<uint64_t>(cap, (((n + npass - 1) / npass) + 31) & ~31ull);
    const uint64_t B = W << (c - 1);
    uint4* d_buckets = nullptr;
    CU(cudaMallocAsync((void**)&d_buckets, B * 192, st));
    CU(cudaMemsetAsync(d_buckets, 0, B * 192, st));
    for (uint64_t lo = 0; lo < n && rc == 0; lo += per) {
      const uint64_t cnt = std::min<uint64_t>(per, n - lo);
      Plan p;
      rc = make_plan(g, p, 1, (uint32_t)cnt, 0, 1, c, 0, flags, false);
      if (rc) break;
      ChunkCtl ctl{(uint32_t)lo, d_buckets, lo + cnt >= n};
      rc = run_pipeline(g, p, (const uint32_t*)d_scalars + 8 * lo, (const uint4*)d_bases, (uint4*)d_out, st,
                        lo == 0 ? points_ready : nullptr, arena, &ctl);
    }
    cudaFreeAsync(d_buckets, st);
  }
  if (rc) return rc;
  rc = arena.release(st);
  if (rc) return rc;
  return finish ? finish_marks(g, st) : 0;
}

// Host-facing single MSM of one device's share. Large n: point-range chunks of growing size (1/8, 1/4, 1/4, 3/8 of the
// points); chunk k+1 is uploaded on the copy stream while chunk k is sorted and accumulated; every chunk accumulates into
// the SAME persistent bucket array (k_accumulate_s continues from the stored bucket, no extra group operations), and the
// reduction / finalisation run once after the last chunk. Without it the accumulation waits for the whole 96 n byte
// base upload (~37 ms at 2^24 over PCIe Gen5) with the GPU idle.
int msm_host_enqueue(Ctx& g, const uint64_t* bases_xy, const uint64_t* scalars, size_t n, unsigned flags,
                     std::vector<void*>& to_free, int sharing) {
  CU(cudaSetDevice(g.device));
  if (n >= (size_t(1) << 31)) return fail(TB200_E_LIMIT, "n = %zu exceeds 2^31 - 1 points per call", n);
  if (n == 0) return msm_dev(g, g.d_result, g.d_result, 0, flags, g.d_result, g.stream, nullptr, nullptr, false);
  uint4 *d_b = nullptr, *d_s = nullptr;
  CU(cudaMallocAsync((void**)&d_b, n * 96, g.stream));
  to_free.push_back(d_b);
  CU(cudaMallocAsync((void**)&d_s, n * 32, g.stream));
  to_free.push_back(d_s);
  const int c = pick_c_single(n);
  const uint64_t W = (uint64_t)num_windows(c);
  const bool chunked = n >= E.host_chunk_min && (uint64_t)n * W <= E.pass_entries_max;
  if (!chunked) {
    // the digit / sort stages only need the scalars; the 3x larger base upload runs on the copy stream and is awaited
    // right before the accumulation kernel
    CU(cudaEventRecord(g.ev_points, g.stream));  // allocations done
    CU(cudaStreamWaitEvent(g.copy_stream, g.ev_points, 0));
    CU(cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_b, bases_xy, n * 96, cudaMemcpyHostToDevice, g.copy_stream));
    CU(cudaEventRecord(g.ev_points, g.copy_stream));
    return msm_dev(g, d_b, d_s, n, flags, g.d_result, g.stream, g.ev_points, nullptr, false);
  }
  // Chunk sizes in sixteenths of the points. The first chunk is small (its upload is the only one the GPU waits for),
  // the LAST chunks are small too: when several GPUs share the host link the call is upload-bound (8 GPUs of this box
  // pull 23-35 GB/s each instead of 55, profiles/r02_h2d_probe_8gpu.txt) and everything after the last byte has
  // arrived -- the last chunk's sort and accumulation -- is exposed. Below 2^23 points four chunks keep the number of
  // pipeline passes (fix-up rounds, launch gaps) down.
  // With four or more GPUs pulling from one host the call IS upload-bound on this box (measured at 8 GPUs, 2^24 points
  // each: 122.7 ms with the schedule above, 111.9 with ten chunks, 107.7 with sixteen equal ones, all paced): equal
  // sixteenths keep every upload behind a compute step of the same size and the exposed tail at one sixteenth. Alone
  // on the link the extra passes cost 2.8 ms (95.4 vs 92.6 ms), so one or two GPUs keep the seven-chunk schedule.
  static const int big[] = {1, 2, 4, 4, 3, 1, 1}, small[] = {2, 4, 4, 6}, shared[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
  const bool many = n >= (size_t(1) << 23);
  constexpr int CMAX = 16;
  int C = many ? (sharing >= 4 ? 16 : 7) : 4;
  const int* frac = many ? (sharing >= 4 ? shared : big) : small;
  if (many && E.host_chunk_count > 0) {   // tuning hook: tb200_set_host_upload
    C = E.host_chunk_count;
    frac = E.host_chunk_frac;
  }
  const size_t unit = ((n + 15) / 16 + 31) & ~size_t(31);
  size_t cut[CMAX + 1] = {0};
  for (int k = 0, acc = 0; k < C; k++) {
    acc += frac[k];
    cut[k + 1] = (k == C - 1) ? n : std::min(n, (size_t)acc * unit);
  }
  Plan plans[CMAX];
  size_t B = 0;
  for (int k = 0; k < C; k++) {
    int rc = make_plan(g, plans[k], 1, (uint32_t)(cut[k + 1] - cut[k]), 0, 1, c, 0, flags);
    if (rc) return rc;
    B = plans[k].B;
  }
  uint4* d_buckets = nullptr;
  CU(cudaMallocAsync((void**)&d_buckets, B * 192, g.stream));
  to_free.push_back(d_buckets);
  CU(cudaMemsetAsync(d_buckets, 0, B * 192, g.stream));
  CU(cudaEventRecord(g.ev_points, g.stream));  // allocations exist
  CU(cudaStreamWaitEvent(g.copy_stream, g.ev_points, 0));
  while (g.chunk_ev.size() < 3 * (size_t)C) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g.chunk_ev.push_back(e);
  }
  // PACED uploads (double buffering): chunk k is uploaded once chunk k-2 has been accumulated, not as early as the copy
  // engine could. A GPU with a fast host link otherwise drains its whole 2 GiB up front and takes bandwidth from the
  // GPUs behind a slower link exactly when those are the ones the call waits for (8 GPUs of this box: 35 vs 23 GB/s
  // when all pull at once, but 29 GB/s for the slow four once the fast four pull at their compute rate only).
  auto upload = [&](int k, bool pace) -> int {
    const size_t lo = cut[k], cnt = cut[k + 1] - cut[k];
    if (pace && k >= 2) CU(cudaStreamWaitEvent(g.copy_stream, g.chunk_ev[2 * C + k - 2], 0));
    CU(cudaMemcpyAsync((char*)d_s + lo * 32, (const char*)scalars + lo * 32, cnt * 32, cudaMemcpyHostToDevice,
                       g.copy_stream));
    CU(cudaEventRecord(g.chunk_ev[2 * k], g.copy_stream));
    CU(cudaMemcpyAsync((char*)d_b + lo * 96, (const char*)bases_xy + lo * 96, cnt * 96, cudaMemcpyHostToDevice,
                       g.copy_stream));
    CU(cudaEventRecord(g.chunk_ev[2 * k + 1], g.copy_stream));
    return 0;
  };
  g.marks.clear();
  if (!E.msm_overlap || g.profiling) {
    if (int rc = g.arena.acquire(g.stream)) return rc;
    for (int k = 0; k < C; k++) {
      if (int rc = upload(k, E.host_upload_pace != 0)) return rc;
      CU(cudaStreamWaitEvent(g.stream, g.chunk_ev[2 * k], 0));
      ChunkCtl ctl{(uint32_t)cut[k], d_buckets, k == C - 1};
      int rc = run_pipeline(g, plans[k], (const uint32_t*)d_s + 8 * cut[k], d_b, g.d_result, g.stream, g.chunk_ev[2 * k + 1],
                            g.arena, &ctl);
      if (rc) return rc;
      CU(cudaEventRecord(g.chunk_ev[2 * C + k], g.stream));   // chunk k is in the buckets
    }
    return g.arena.release(g.stream);
  }
  for (int k = 0; k < C; k++)   // the overlap variant orders its accumulations itself: uploads as early as possible
    if (int rc = upload(k, false)) return rc;
  // Chunks alternate between the main stream and a high-priority side stream: the sort of chunk k+1 (memory bound) runs
  // next to the accumulation of chunk k (integer-pipe bound). The accumulations themselves stay in chunk order -- they
  // update the same persistent buckets -- through the ev_split events.
  CU(cudaEventRecord(g.ev_split[2], g.stream));          // the bucket array is zeroed
  CU(cudaStreamWaitEvent(g.split_stream[0], g.ev_split[2], 0));
  note_geometry(g, plans[C - 1], (uint64_t)n * W, B);
  for (int k = 0; k < C; k++) {
    cudaStream_t s = (k & 1) ? g.split_stream[0] : g.stream;
    Arena& ar = (k & 1) ? g.split_arena[0] : g.arena;
    CU(cudaStreamWaitEvent(s, g.chunk_ev[2 * k], 0));
    if (int rc = ar.acquire(s)) return rc;
    ChunkCtl ctl{(uint32_t)cut[k], d_buckets, k == C - 1};
    MsmGeom q = plans[k].geo;
    q.ref_base = ctl.ref_base;
    PipeBufs b;
    if (int rc = pipe_sort(g, plans[k], q, (const uint32_t*)d_s + 8 * cut[k], s, ar, &ctl, b)) return rc;
    if (k > 0) CU(cudaStreamWaitEvent(s, g.ev_split[(k - 1) & 1], 0));   // the previous chunk's buckets are final
    if (int rc = pipe_accumulate(g, plans[k], q, b, d_b, s, g.chunk_ev[2 * k + 1], &ctl)) return rc;
    if (k == C - 1) {
      const uint4* gw = nullptr;
      if (int rc = pipe_reduce(g, plans[k], q, b, s, ar, &ctl, &gw)) return rc;
      if (int rc = pipe_finalize(g, plans[k], q, gw, g.d_result, s, ar)) return rc;
    }
    if (int rc = ar.release(s)) return rc;
    CU(cudaEventRecord(g.ev_split[k & 1], s));
  }
  if ((C - 1) & 1) CU(cudaStreamWaitEvent(g.stream, g.ev_split[(C - 1) & 1], 0));   // the result is ordered into g.stream
  return 0;
}

// ---- shared-base batch ------------------------------------------------------------------------------------------------
// rows are processed in chunks that respect the per-pass limits of the pipeline
int batch_dev(Ctx& g, const uint4* table, int c, int W, uint32_t srs_n, const uint32_t* d_scalars, size_t rows, size_t cols,
              long long rs, long long cs, unsigned flags, uint4* d_out, cudaStream_t st) {
  if (cols > srs_n) return fail(TB200_E_ARG, "cols = %zu exceeds the SRS size %u", cols, srs_n);
  if (((uintptr_t)d_scalars | (uintptr_t)d_out) & 15) return fail(TB200_E_ARG, "device pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  const uint64_t per_row = (uint64_t)std::max<size_t>(cols, 1) * W;
  const uint64_t nb = 1ull << (c - 1);
  const uint64_t chunk = std::min<uint64_t>(
      {(uint64_t)rows, std::min<uint64_t>(E.pass_entries_max, (1ull << 31) - 1) / per_row, (1ull << 30) / nb});
  if (chunk == 0) return fail(TB200_E_LIMIT, "a single row exceeds the per-pass limits");
  if (int rc = g.arena.acquire(st)) return rc;
  for (size_t r0 = 0; r0 < rows; r0 += chunk) {
    uint32_t nr = (uint32_t)std::min<uint64_t>(chunk, rows - r0);
    Plan p;
    int rc = make_plan(g, p, nr, (uint32_t)cols, rs, cs, c, 1, flags);
    if (rc) return rc;
    // entry refs are w * cols + j and the callers guarantee cols == srs_n, the stride of the window tables
    rc = run_pipeline(g, p, d_scalars + 8 * (long long)r0 * rs, table, d_out + 6 * r0, st, nullptr, g.arena);
    if (rc) return rc;
  }
  return g.arena.release(st);
}

int srs_build_table(Ctx& g, const uint64_t* bases_xy, size_t n, int c, int W, uint4** out_table) {
  CU(cudaSetDevice(g.device));
  uint4 *table = nullptr, *d_b = nullptr;
  cudaError_t e = cudaMalloc((void**)&table, (size_t)W * n * 96);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_b, n * 96);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, bases_xy, n * 96, cudaMemcpyHostToDevice, g.stream);
  if (e == cudaSuccess) {
    k_srs_tables<<<cdiv(n, 128), 128, 0, g.stream>>>(d_b, (uint32_t)n, c, W, table);
    g_launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
  if (d_b) cudaFree(d_b);
  if (e != cudaSuccess) {
    if (table) cudaFree(table);
    return fail((int)e, "SRS table build failed on device %d: %s", g.device, cudaGetErrorString(e));
  }
  *out_table = table;
  return 0;
}

int g1_sum_dev(Ctx& g, const void* d_pts, size_t n, void* d_out, cudaStream_t st) {
  (void)g;
  LAUNCH(k_g1_sum, 1, 32, st, (const uint4*)d_pts, (uint32_t)n, (uint4*)d_out);
  return 0;
}

static int fr_fold(cudaStream_t st, uint32_t* y, uint32_t split, const uint32_t* d_c_inv, int mont) {
  LAUNCH(k_compress_fr, cdiv(split, 128), 128, st, y, split, d_c_inv, mont);
  return 0;
}

}  // namespace tbe

using namespace tbe;

extern "C" {

int tb200_msm_g1_dev(const void* d_bases_xy, const void* d_scalars, size_t n, unsigned flags, void* d_out_xy,
                     void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out_xy || (n && (!d_bases_xy || !d_scalars))) return fail(TB200_E_ARG, "null pointer");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  return msm_dev(g, d_bases_xy, d_scalars, n, flags, d_out_xy, stream ? (cudaStream_t)stream : g.stream);
}

int tb200_msm_g1_batch_dev(tb200_srs_t srs, const void* d_scalars, size_t rows, size_t cols, ptrdiff_t row_stride,
                           ptrdiff_t col_stride, unsigned flags, void* d_out_xy, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!srs || (rows && (!d_out_xy || (cols && !d_scalars)))) return fail(TB200_E_ARG, "null pointer");
  if (cols != srs->n && cols != 0)
    return fail(TB200_E_ARG, "cols (%zu) must equal the SRS size (%u): window tables are laid out per SRS", cols, srs->n);
  if (row_stride < 0 || col_stride < 0) return fail(TB200_E_ARG, "negative strides are not supported");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  g.marks.clear();
  int rc = batch_dev(g, srs->table[0], srs->c, srs->W, srs->n, (const uint32_t*)d_scalars, rows, cols, row_stride,
                     col_stride, flags, (uint4*)d_out_xy, st);
  return rc ? rc : finish_marks(g, st);
}

// ---- MultilinearPC::open (G2 proofs) / open_g1 (G1 proofs): quotient loop on the device, one MSM per variable ---------
}  // extern "C"

// One PST opening in flight: everything is ENQUEUED by pst_open_enqueue (uploads and the quotient loop on the job's own
// stream, the per-variable MSMs on the context's side streams), pst_open_finish waits and downloads the proofs. The
// blocking entry points run the two back to back; tb200_pst_open_g2_begin / _end expose them separately so that the G2
// opening of q (src/sqrt_pst.rs:225), which does not depend on the MIPP transcript, runs NEXT TO the MIPP rounds.
struct tb200_pst_open {
  cudaStream_t st = nullptr;
  cudaEvent_t ev = nullptr;
  uint32_t *d_r0 = nullptr, *d_r1 = nullptr, *d_q = nullptr, *d_p = nullptr;
  uint4 *d_bases = nullptr, *d_proofs = nullptr;
  size_t nv = 0, pt = 0;
};

namespace {
void pst_open_release(tb200_pst_open* j) {
  if (!j) return;
  if (j->st) {
    for (void* p : {(void*)j->d_r0, (void*)j->d_r1, (void*)j->d_q, (void*)j->d_p, (void*)j->d_bases, (void*)j->d_proofs})
      if (p) cudaFreeAsync(p, j->st);
    cudaStreamDestroy(j->st);
  }
  if (j->ev) cudaEventDestroy(j->ev);
  delete j;
}

int pst_open_enqueue(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                     unsigned flags, bool g2, tb200_pst_open** out) {
  if (!evals || !point || !level_bases || !out) return fail(TB200_E_ARG, "null pointer");
  if (nv == 0 || nv > 28) return fail(nv ? TB200_E_LIMIT : TB200_E_ARG, "nv = %zu: between 1 and 28 variables", nv);
  for (size_t i = 0; i < nv; i++)
    if (!level_bases[i]) return fail(TB200_E_ARG, "level_bases[%zu] is null", i);
  Ctx& g = primary();
  const size_t n = size_t(1) << nv, pt = g2 ? 192 : 96;
  tb200_pst_open* j = new tb200_pst_open();
  j->nv = nv;
  j->pt = pt;
  auto bail = [&](int rc) {
    if (j->st) cudaStreamSynchronize(j->st);
    pst_open_release(j);
    return rc;
  };
#define CUJ(expr)                                                                                                    \
  do {                                                                                                               \
    cudaError_t e__ = (expr);                                                                                        \
    if (e__ != cudaSuccess) return bail(fail((int)e__, "%s failed: %s", #expr, cudaGetErrorString(e__)));            \
  } while (0)
  CUJ(cudaStreamCreateWithFlags(&j->st, cudaStreamNonBlocking));
  CUJ(cudaEventCreateWithFlags(&j->ev, cudaEventDisableTiming));
  cudaStream_t st = j->st;
  // level i occupies [off_i, off_i + 2^(nv-i)) of d_q / d_bases, off_i = 2^(nv+1) - 2^(nv-i+1)
  CUJ(cudaMallocAsync((void**)&j->d_r0, n * 32, st));
  CUJ(cudaMallocAsync((void**)&j->d_r1, std::max<size_t>(n / 2, 1) * 32, st));
  CUJ(cudaMallocAsync((void**)&j->d_q, 2 * n * 32, st));
  CUJ(cudaMallocAsync((void**)&j->d_p, nv * 32, st));
  CUJ(cudaMallocAsync((void**)&j->d_bases, 2 * n * pt, st));
  CUJ(cudaMallocAsync((void**)&j->d_proofs, nv * pt, st));
  CUJ(cudaMemcpyAsync(j->d_r0, evals, n * 32, cudaMemcpyHostToDevice, st));
  CUJ(cudaMemcpyAsync(j->d_p, point, nv * 32, cudaMemcpyHostToDevice, st));
  if (!(flags & TB200_SCALARS_MONT)) {
    k_fr_to_mont<<<cdiv(n, 128), 128, 0, st>>>(j->d_r0, (uint32_t)n);
    k_fr_to_mont<<<cdiv(nv, 128), 128, 0, st>>>(j->d_p, (uint32_t)nv);
    g_launches += 2;
  }
  // the quotient loop is a cheap sequential chain; the nv MSMs that consume it are independent of each other and
  // latency-bound (Horner chain + inversion), so they run concurrently on side streams with their own workspaces
  std::vector<size_t> off(nv);
  uint32_t *r_in = j->d_r0, *r_out = j->d_r1;
  size_t o = 0;
  for (size_t i = 0; i < nv; i++) {
    const size_t half = size_t(1) << (nv - i - 1);
    off[i] = o;
    CUJ(cudaMemcpyAsync((char*)j->d_bases + o * pt, level_bases[i], 2 * half * pt, cudaMemcpyHostToDevice, st));
    k_pst_level<<<cdiv(half, 128), 128, 0, st>>>(r_in, (uint32_t)half, j->d_p + 8 * i, r_out, j->d_q + 8 * o);
    g_launches++;
    std::swap(r_in, r_out);
    o += 2 * half;
  }
  CUJ(cudaGetLastError());
  CUJ(cudaEventRecord(j->ev, st));
  const bool prof = g.profiling;
  g.profiling = false;
  int rc = 0;
  for (size_t i = 0; i < nv && rc == 0; i++) {
    const int sl = (int)(i % Ctx::SIDE);
    if (!g.side_stream[sl]) {
      cudaError_t e = cudaStreamCreateWithFlags(&g.side_stream[sl], cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g.side_done[sl], cudaEventDisableTiming);
      if (e != cudaSuccess) {
        rc = fail((int)e, "side stream creation failed: %s", cudaGetErrorString(e));
        break;
      }
    }
    if (i < (size_t)Ctx::SIDE) cudaStreamWaitEvent(g.side_stream[sl], j->ev, 0);
    rc = msm_dev(g, (char*)j->d_bases + off[i] * pt, j->d_q + 8 * off[i], size_t(1) << (nv - i), TB200_SCALARS_MONT,
                 (char*)j->d_proofs + i * pt, g.side_stream[sl], nullptr, &g.side_arena[sl], false, g2);
  }
  g.profiling = prof;
  g.marks.clear();
  for (int sl = 0; sl < Ctx::SIDE && sl < (int)nv; sl++) {
    if (!g.side_stream[sl]) continue;
    cudaEventRecord(g.side_done[sl], g.side_stream[sl]);
    cudaStreamWaitEvent(st, g.side_done[sl], 0);
  }
  if (rc) return bail(rc);
#undef CUJ
  *out = j;
  return 0;
}

int pst_open_finish(tb200_pst_open* j, uint64_t* proofs) {
  cudaError_t e = cudaSuccess;
  if (proofs) e = cudaMemcpyAsync(proofs, j->d_proofs, j->nv * j->pt, cudaMemcpyDeviceToHost, j->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(j->st);
  int rc = e == cudaSuccess ? 0 : fail((int)e, "proof copy failed: %s", cudaGetErrorString(e));
  pst_open_release(j);
  return rc;
}

int pst_open_locked(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                    unsigned flags, uint64_t* proofs, bool g2) {
  if (!proofs) return fail(TB200_E_ARG, "null pointer");
  if (nv == 0) return (evals && point && level_bases) ? 0 : fail(TB200_E_ARG, "null pointer");
  tb200_pst_open* j = nullptr;
  if (int rc = pst_open_enqueue(evals, nv, point, level_bases, flags, g2, &j)) return rc;
  return pst_open_finish(j, proofs);
}
}  // namespace

extern "C" {

/* asynchronous form: _begin returns once everything is enqueued (the host buffers may be reused when it returns: the
 * uploads are staged), _end waits and writes the nv proofs */
int tb200_pst_open_g2_begin(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                            unsigned flags, tb200_pst_open_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(primary().device));
  return pst_open_enqueue(evals, nv, point, level_bases, flags, true, out);
}
int tb200_pst_open_g1_begin(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                            unsigned flags, tb200_pst_open_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(primary().device));
  return pst_open_enqueue(evals, nv, point, level_bases, flags, false, out);
}
int tb200_pst_open_end(tb200_pst_open_t h, uint64_t* proofs) {
  if (!h) return fail(TB200_E_ARG, "null handle");
  if (need_ready()) return TB200_E_STATE;
  // no library lock while waiting: the point of the asynchronous form is that other calls run meanwhile
  {
    std::lock_guard<std::mutex> lk(g_mu);
    CU(cudaSetDevice(primary().device));
  }
  return pst_open_finish(h, proofs);
}

int tb200_pst_open_g1(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                      unsigned flags, uint64_t* proofs) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(primary().device));
  return pst_open_locked(evals, nv, point, level_bases, flags, proofs, false);
}
int tb200_pst_open_g2(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                      unsigned flags, uint64_t* proofs) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  CU(cudaSetDevice(primary().device));
  return pst_open_locked(evals, nv, point, level_bases, flags, proofs, true);
}

// A ragged batch of independent small MSMs in ONE launch of the Straus kernel: row i takes the next row_len[i] points
// (0 .. 1024 each). The verifier's G1 work -- the 28-point UC fold, the 13-point fold of check_2 and the nv + 1 two-point
// rows of check -- is three independent MSMs of different lengths: one launch, the latency of one.
int tb200_msm_g1_rows(const uint64_t* bases_xy, const uint64_t* scalars, const size_t* row_len, size_t rows, unsigned flags,
                      uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (rows == 0) return 0;
  if (!row_len || !out_xy) return fail(TB200_E_ARG, "null pointer");
  if (rows > (1u << 16)) return fail(TB200_E_LIMIT, "tb200_msm_g1_rows takes at most 65536 rows");
  size_t n = 0;
  std::vector<uint32_t> segs;
  for (size_t r = 0; r < rows; r++) {
    const size_t len = row_len[r];
    if (len > (size_t)SMALL_MAX_POINTS) return fail(TB200_E_LIMIT, "row %zu has %zu points: at most %d per row", r, len, SMALL_MAX_POINTS);
    const uint32_t ctas = std::max<uint32_t>(1, cdiv(len, SMALL_QUADS)), cta0 = (uint32_t)(segs.size() / SMALL_SEG_WORDS);
    for (uint32_t k = 0; k < ctas; k++) {
      const uint32_t cnt = (uint32_t)std::min<size_t>(SMALL_QUADS, len - std::min<size_t>(len, (size_t)k * SMALL_QUADS));
      const uint32_t d[SMALL_SEG_WORDS] = {(uint32_t)(n + (size_t)k * SMALL_QUADS), cnt, (uint32_t)r, cta0, ctas};
      segs.insert(segs.end(), d, d + SMALL_SEG_WORDS);
    }
    n += len;
  }
  if (n && (!bases_xy || !scalars)) return fail(TB200_E_ARG, "null pointer");
  const size_t nblk = segs.size() / SMALL_SEG_WORDS;
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  StreamScratch sc(g.stream);
  uint4 *d_b = nullptr, *d_o = nullptr, *scratch = nullptr;
  uint32_t *d_s = nullptr, *d_segs = nullptr;
  const size_t ticket_bytes = ((rows * 4 + 15) / 16) * 16;
  CU(sc.alloc(&d_b, n * 96));
  CU(sc.alloc(&d_s, n * 32));
  CU(sc.alloc(&d_o, rows * 96));
  CU(sc.alloc(&d_segs, segs.size() * 4));
  CU(sc.alloc(&scratch, nblk * 192 + ticket_bytes));
  CU(cudaMemsetAsync(scratch + nblk * 12, 0, ticket_bytes, g.stream));
  if (n) {
    CU(cudaMemcpyAsync(d_b, bases_xy, n * 96, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, g.stream));
  }
  CU(cudaMemcpyAsync(d_segs, segs.data(), segs.size() * 4, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_msm_small, (uint32_t)nblk, 4 * SMALL_QUADS, g.stream, d_b, d_s, (uint32_t)n, (flags & TB200_SCALARS_MONT) ? 1 : 0,
         scratch, d_o, 0u, (const uint32_t*)d_segs);
  CU(cudaMemcpyAsync(out_xy, d_o, rows * 96, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));   // `segs` (pageable host memory) is consumed by then as well
  sc.done = true;
  return 0;
}

// ---- a single MSM in flight next to other calls ------------------------------------------------------------------------
// The reference runs independent MSMs side by side (`try_par!` / `rayon::join`, src/macros.rs:1-17, src/mipp.rs:77-85) and
// computes cross-checks whose result nothing waits for (`MultilinearPC::commit(ck, &q)` for the debug_assert of
// src/sqrt_pst.rs:205-206). _begin uploads on the job's own stream and enqueues the MSM on a side pipeline (own stream,
// own workspace); _end waits and downloads the point. The host buffers must stay valid until _end returns.
}  // extern "C"
struct tb200_msm_job {
  cudaStream_t st = nullptr;
  cudaEvent_t ev = nullptr;
  uint4* d_b = nullptr;
  uint32_t* d_s = nullptr;
  uint4* d_out = nullptr;
};
namespace {
void msm_job_release(tb200_msm_job* j) {
  if (!j) return;
  if (j->st) {
    for (void* p : {(void*)j->d_b, (void*)j->d_s, (void*)j->d_out})
      if (p) cudaFreeAsync(p, j->st);
    cudaStreamDestroy(j->st);
  }
  if (j->ev) cudaEventDestroy(j->ev);
  delete j;
}
}  // namespace
extern "C" {
int tb200_msm_g1_begin(const uint64_t* bases_xy, const uint64_t* scalars, size_t n, unsigned flags, tb200_msm_job_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && (!bases_xy || !scalars))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (size_t(1) << 26)) return fail(TB200_E_LIMIT, "tb200_msm_g1_begin is meant for MSMs that run NEXT TO other work (n < 2^26)");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  constexpr int sl = Ctx::SIDE - 1;   // the side pipeline a concurrent PST opening uses least
  if (!g.side_stream[sl]) {
    CU(cudaStreamCreateWithFlags(&g.side_stream[sl], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&g.side_done[sl], cudaEventDisableTiming));
  }
  tb200_msm_job* j = new tb200_msm_job();
  auto bail = [&](int rc) {
    if (j->st) cudaStreamSynchronize(j->st);
    cudaStreamSynchronize(g.side_stream[sl]);
    msm_job_release(j);
    return rc;
  };
#define CUJ(expr)                                                                                         \
  do {                                                                                                    \
    cudaError_t e__ = (expr);                                                                             \
    if (e__ != cudaSuccess) return bail(fail((int)e__, "%s failed: %s", #expr, cudaGetErrorString(e__))); \
  } while (0)
  CUJ(cudaStreamCreateWithFlags(&j->st, cudaStreamNonBlocking));
  CUJ(cudaEventCreateWithFlags(&j->ev, cudaEventDisableTiming));
  CUJ(cudaMallocAsync((void**)&j->d_b, std::max<size_t>(n, 1) * 96, j->st));
  CUJ(cudaMallocAsync((void**)&j->d_s, std::max<size_t>(n, 1) * 32, j->st));
  CUJ(cudaMallocAsync((void**)&j->d_out, 96, j->st));
  if (n) {
    CUJ(cudaMemcpyAsync(j->d_b, bases_xy, n * 96, cudaMemcpyHostToDevice, j->st));
    CUJ(cudaMemcpyAsync(j->d_s, scalars, n * 32, cudaMemcpyHostToDevice, j->st));
  }
  CUJ(cudaEventRecord(j->ev, j->st));
  CUJ(cudaStreamWaitEvent(g.side_stream[sl], j->ev, 0));
  const bool prof = g.profiling;
  g.profiling = false;
  int rc = msm_dev(g, j->d_b, j->d_s, n, flags, j->d_out, g.side_stream[sl], nullptr, &g.side_arena[sl], false, false);
  g.profiling = prof;
  g.marks.clear();
  if (rc) return bail(rc);
  CUJ(cudaEventRecord(g.side_done[sl], g.side_stream[sl]));
  CUJ(cudaStreamWaitEvent(j->st, g.side_done[sl], 0));
#undef CUJ
  *out = j;
  return 0;
}
int tb200_msm_g1_end(tb200_msm_job_t job, uint64_t out_xy[12]) {
  if (!job) return fail(TB200_E_ARG, "null handle");
  if (need_ready()) return TB200_E_STATE;
  {   // no library lock while waiting: other calls run meanwhile
    std::lock_guard<std::mutex> lk(g_mu);
    CU(cudaSetDevice(primary().device));
  }
  cudaError_t e = cudaSuccess;
  if (out_xy) e = cudaMemcpyAsync(out_xy, job->d_out, 96, cudaMemcpyDeviceToHost, job->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(job->st);
  const int rc = e == cudaSuccess ? 0 : fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
  msm_job_release(job);
  return rc;
}

// ---- MIPP (G1 side) ---------------------------------------------------------------------------------------------------
int tb200_mipp_g1_begin(const uint64_t* a_xy, const uint64_t* y, size_t n, unsigned flags, tb200_mipp_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || !a_xy || !y || n == 0 || (n & (n - 1)) || n >= (1u << 28))
    return fail(TB200_E_ARG, "MIPP vectors must have a power-of-two length (got %zu)", n);
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  tb200_mipp* h = new tb200_mipp();
  h->n = (uint32_t)n;
  h->flags = flags;
  cudaError_t e = cudaMallocAsync((void**)&h->a, n * 96, g.stream);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&h->y, n * 32, g.stream);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&h->scal, 64 * 64, g.stream);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&h->digits, 64 * 32, g.stream);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&h->scal_host, 64 * 64);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->a, a_xy, n * 96, cudaMemcpyHostToDevice, g.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->y, y, n * 32, cudaMemcpyHostToDevice, g.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
  if (e == cudaSuccess && n >= 2 && g1_fold_mult_bytes(n) <= FOLD_MULT_BYTES_MAX) {
    // two-phase fold (kernels_pairing.cuh): the multiples of the first round's right half start right away
    e = cudaMallocAsync((void**)&h->mult, g1_fold_mult_bytes(n), g.stream);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->pre_st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_pre, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fold, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&h->sel, 64 * glv::SEL_MAX * 2, g.stream);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&h->sel_host, 64 * glv::SEL_MAX * 2);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);   // the stream-ordered allocations exist for every stream
    if (e == cudaSuccess && g1_fold_pre(h->pre_st, h->a, (uint32_t)(n / 2), (uint32_t)(n / 2), h->mult)) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaEventRecord(h->ev_pre, h->pre_st);
  }
  if (e != cudaSuccess) {
    cudaFreeAsync(h->mult, g.stream);
    cudaFreeAsync(h->sel, g.stream);
    cudaFreeHost(h->sel_host);
    if (h->pre_st) cudaStreamDestroy(h->pre_st);
    if (h->ev_pre) cudaEventDestroy(h->ev_pre);
    if (h->ev_fold) cudaEventDestroy(h->ev_fold);
    cudaFreeAsync(h->a, g.stream);
    cudaFreeAsync(h->y, g.stream);
    cudaFreeAsync(h->scal, g.stream);
    cudaFreeAsync(h->digits, g.stream);
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
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  const uint32_t split = h->n / 2;
  // comm_u_l = MSM(a[:split], y[split:]), comm_u_r = MSM(a[split:], y[:split])   (src/mipp.rs:82-84)
  // the two MSMs are independent and latency-bound (sequential Horner + inversion tail): run them concurrently on
  // two streams with separate workspaces (the reference runs them as two rayon tasks, src/mipp.rs:77-85)
  const bool prof = g.profiling;
  g.profiling = false;
  CU(cudaEventRecord(g.ev_join, g.stream));  // the folds are only enqueued: stream2 must follow them
  CU(cudaStreamWaitEvent(g.stream2, g.ev_join, 0));
  int rc = msm_dev(g, h->a, h->y + 8 * (size_t)split, split, h->flags, g.d_result, g.stream, nullptr, nullptr, false);
  if (rc == 0)
    rc = msm_dev(g, h->a + 6 * (size_t)split, h->y, split, h->flags, g.d_result + 6, g.stream2, nullptr, &g.arena2, false);
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
  Ctx& g = primary();
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
  // a_l + c a_r over the G1 endomorphism (engine_pairing.cu): 127 doublings instead of 253
  if (h->mult) {
    // the scalar is decomposed on the host (one value per round): the device gets the list of stored multiples to add
    uint16_t* hsel = h->sel_host + (size_t)glv::SEL_MAX * h->round;
    uint16_t* dsel = h->sel + (size_t)glv::SEL_MAX * h->round;
    glv::select_g1(c, mont != 0, hsel);
    CU(cudaMemcpyAsync(dsel, hsel, (size_t)(hsel[0] + 1) * 2, cudaMemcpyHostToDevice, g.stream));
    CU(cudaStreamWaitEvent(g.stream, h->ev_pre, 0));         // the multiples of this round's right half
    if (int rc = g1_fold_apply(g.stream, dsel, h->a, split, h->mult)) return rc;
    if (split >= 2) {                                        // phase A of the next round, off the critical path
      CU(cudaEventRecord(h->ev_fold, g.stream));
      CU(cudaStreamWaitEvent(h->pre_st, h->ev_fold, 0));
      if (int rc = g1_fold_pre(h->pre_st, h->a, split / 2, split / 2, h->mult)) return rc;
      CU(cudaEventRecord(h->ev_pre, h->pre_st));
    }
  } else if (int rc = g1_fold_glv(g.stream, ds, mont, h->digits + 8 * h->round, h->a, split)) {
    return rc;
  }
  if (int rc = fr_fold(g.stream, h->y, split, ds + 8, mont)) return rc;
  h->round++;
  h->n = split;
  return 0;
}

int tb200_mipp_g1_read(tb200_mipp_t h, uint64_t* a_xy, uint64_t* y) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h) return fail(TB200_E_ARG, "null handle");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  if (a_xy) CU(cudaMemcpyAsync(a_xy, h->a, (size_t)h->n * 96, cudaMemcpyDeviceToHost, g.stream));
  if (y) CU(cudaMemcpyAsync(y, h->y, (size_t)h->n * 32, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return 0;
}
int tb200_mipp_g1_end(tb200_mipp_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!h) return fail(TB200_E_ARG, "null handle");
  if (E.ready) {
    Ctx& g = primary();
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.stream);
    if (h->pre_st) {
      cudaStreamSynchronize(h->pre_st);
      cudaStreamDestroy(h->pre_st);
      cudaEventDestroy(h->ev_pre);
      cudaEventDestroy(h->ev_fold);
    }
    cudaFreeAsync(h->mult, g.stream);
    cudaFreeAsync(h->sel, g.stream);
    cudaFreeHost(h->sel_host);
    cudaFreeAsync(h->a, g.stream);
    cudaFreeAsync(h->y, g.stream);
    cudaFreeAsync(h->scal, g.stream);
    cudaFreeAsync(h->digits, g.stream);
    cudaFreeHost(h->scal_host);
  }
  delete h;
  return 0;
}

int tb200_compress_g1(uint64_t* vec_xy, size_t split, const uint64_t scaler[4], unsigned flags) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!vec_xy || !scaler) return fail(TB200_E_ARG, "null pointer");
  if (split == 0) return 0;
  if (split >= (1u << 27)) return fail(TB200_E_LIMIT, "split too large");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  uint4* d_v = nullptr;
  uint32_t* d_k = nullptr;
  CU(cudaMallocAsync((void**)&d_v, 2 * split * 96, g.stream));
  CU(cudaMallocAsync((void**)&d_k, 32, g.stream));
  CU(cudaMemcpyAsync(d_v, vec_xy, 2 * split * 96, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d_k, scaler, 32, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_compress_g1, cdiv(split, 128), 128, g.stream, d_v, (uint32_t)split, d_k, (flags & TB200_SCALARS_MONT) ? 1 : 0);
  CU(cudaMemcpyAsync(vec_xy, d_v, split * 96, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  cudaFreeAsync(d_v, g.stream);
  cudaFreeAsync(d_k, g.stream);
  return 0;
}

// ---- sqrt_pst scalar work on the device -------------------------------------------------------------------------------
static int fr_chis_locked(const uint64_t* b, size_t m, uint64_t* chis_out, int subset) {
  if (!chis_out || (m && !b) || m > 28) return fail(TB200_E_ARG, "bad arguments (m = %zu)", m);
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  const size_t n = size_t(1) << m;
  uint32_t *d_b = nullptr, *d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_b, std::max<size_t>(m, 1) * 32, g.stream));
  CU(cudaMallocAsync((void**)&d_o, n * 32, g.stream));
  if (m) CU(cudaMemcpyAsync(d_b, b, m * 32, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_fr_chis, cdiv(n, 128), 128, g.stream, d_b, (uint32_t)m, d_o, subset);
  CU(cudaMemcpyAsync(chis_out, d_o, n * 32, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  cudaFreeAsync(d_b, g.stream);
  cudaFreeAsync(d_o, g.stream);
  return 0;
}
int tb200_fr_chis(const uint64_t* b, size_t m, uint64_t* chis_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  return fr_chis_locked(b, m, chis_out, 0);
}
int tb200_fr_subset_products(const uint64_t* b, size_t m, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  return fr_chis_locked(b, m, out, 1);
}
static int fr_matvec_enqueue(Ctx& g, const void* d_Z, size_t rows, size_t cols, const void* d_v, void* d_out, cudaStream_t st) {
  if (rows >= (1ull << 31) || cols >= (1ull << 31)) return fail(TB200_E_LIMIT, "matrix too large");
  if (((uintptr_t)d_Z | (uintptr_t)d_v | (uintptr_t)d_out) & 15) return fail(TB200_E_ARG, "device pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  (void)g;
  LAUNCH(k_fr_matvec, cdiv(rows * 32, 256), 256, st, (const uint32_t*)d_Z, (uint32_t)rows, (uint32_t)cols,
         (const uint32_t*)d_v, (uint32_t*)d_out);
  return 0;
}
int tb200_fr_matvec_dev(const void* d_Z, size_t rows, size_t cols, const void* d_v, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || ((rows && cols) && (!d_Z || !d_v))) return fail(TB200_E_ARG, "null pointer");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  return fr_matvec_enqueue(g, d_Z, rows, cols, d_v, d_out, stream ? (cudaStream_t)stream : g.stream);
}
int tb200_fr_matvec(const uint64_t* Z, size_t rows, size_t cols, const uint64_t* v, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || !Z || !v || rows == 0 || cols == 0) return fail(TB200_E_ARG, "bad arguments");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  uint32_t *d_z = nullptr, *d_v = nullptr, *d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_z, rows * cols * 32, g.stream));
  CU(cudaMallocAsync((void**)&d_v, cols * 32, g.stream));
  CU(cudaMallocAsync((void**)&d_o, rows * 32, g.stream));
  CU(cudaMemcpyAsync(d_z, Z, rows * cols * 32, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d_v, v, cols * 32, cudaMemcpyHostToDevice, g.stream));
  int rc = fr_matvec_enqueue(g, d_z, rows, cols, d_v, d_o, g.stream);
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
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  return g1_sum_dev(g, d_pts_xy, n, d_out_xy, stream ? (cudaStream_t)stream : g.stream);
}
int tb200_g1_sum(const uint64_t* pts_xy, size_t n, uint64_t out_xy[12]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out_xy || (n && !pts_xy)) return fail(TB200_E_ARG, "null pointer");
  if (n > 128) return fail(TB200_E_LIMIT, "tb200_g1_sum (host form) takes at most 128 points");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  if (n) CU(cudaMemcpyAsync(g.d_result + 6, pts_xy, n * 96, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_g1_sum, 1, 32, g.stream, g.d_result + 6, (uint32_t)n, g.d_result);
  CU(cudaMemcpyAsync(g.h_result, g.d_result, 96, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  memcpy(out_xy, g.h_result, 96);
  return 0;
}
// `rows` independent MSMs of `per_row` (1..8) points each in ONE launch of the Straus kernel: out[i] = sum_j
// scalars[i][j] * bases[i][j]. The verifier's `g_mask[i] - z_i g` (ark-poly-commit `check`) and `C - v g`.
int tb200_msm_g1_each(const uint64_t* bases_xy, const uint64_t* scalars, size_t rows, size_t per_row, unsigned flags,
                      uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (rows == 0) return 0;
  if (!bases_xy || !scalars || !out_xy) return fail(TB200_E_ARG, "null pointer");
  if (per_row == 0 || per_row > (size_t)SMALL_QUADS) return fail(TB200_E_ARG, "per_row = %zu: between 1 and %d", per_row, SMALL_QUADS);
  if (rows > (1u << 16)) return fail(TB200_E_LIMIT, "tb200_msm_g1_each takes at most 65536 rows");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  const size_t n = rows * per_row;
  StreamScratch sc(g.stream);
  uint4 *d_b = nullptr, *d_o = nullptr;
  uint32_t* d_s = nullptr;
  CU(sc.alloc(&d_b, n * 96));
  CU(sc.alloc(&d_s, n * 32));
  CU(sc.alloc(&d_o, rows * 96));
  CU(cudaMemcpyAsync(d_b, bases_xy, n * 96, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_msm_small, (uint32_t)rows, 4 * SMALL_QUADS, g.stream, d_b, d_s, (uint32_t)n, (flags & TB200_SCALARS_MONT) ? 1 : 0,
         (uint4*)nullptr, d_o, (uint32_t)per_row, (const uint32_t*)nullptr);
  CU(cudaMemcpyAsync(out_xy, d_o, rows * 96, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  sc.done = true;
  return 0;
}
int tb200_g1_outer_sum_dev(const void* d_a_xy, size_t na, const void* d_b_xy, size_t nb, void* d_out_xy, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_a_xy || !d_b_xy || !d_out_xy || na == 0 || nb == 0) return fail(TB200_E_ARG, "bad arguments");
  if ((uint64_t)na * nb >= (1ull << 31)) return fail(TB200_E_LIMIT, "outer sum too large");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  LAUNCH(k_g1_outer_sum, cdiv((uint64_t)na * nb, 128), 128, st, (const uint4*)d_a_xy, (uint32_t)na, (const uint4*)d_b_xy,
         (uint32_t)nb, (uint4*)d_out_xy);
  return 0;
}

// ---- microbenchmarks / unit-test hooks ------------------------------------------------------------------------------
int tb200_int_pipe_peak(int kind, int iters, double* out_per_s) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out_per_s || iters <= 0 || kind < 0 || kind > 2) return fail(TB200_E_ARG, "bad arguments");
  Ctx& g = primary();
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

namespace tbe {
// unit-test scaffolding shared with the other units: blocking copies around one launch on the primary's stream
int with_buffers(const void* a, size_t abytes, const void* b, size_t bbytes, void* o1, size_t o1bytes, void* o2,
                 size_t o2bytes, const std::function<int(char*, char*, char*, char*)>& launch) {
  char *d_a = nullptr, *d_b = nullptr, *d_o1 = nullptr, *d_o2 = nullptr;
  CU(cudaMalloc((void**)&d_a, std::max<size_t>(abytes, 16)));
  CU(cudaMalloc((void**)&d_b, std::max<size_t>(bbytes, 16)));
  CU(cudaMalloc((void**)&d_o1, std::max<size_t>(o1bytes, 16)));
  CU(cudaMalloc((void**)&d_o2, std::max<size_t>(o2bytes, 16)));
  CU(cudaMemcpy(d_a, a, abytes, cudaMemcpyHostToDevice));
  if (b) CU(cudaMemcpy(d_b, b, bbytes, cudaMemcpyHostToDevice));
  int rc = launch(d_a, d_b, d_o1, d_o2);
  if (rc == 0) {
    CU(cudaStreamSynchronize(primary().stream));
    CU(cudaMemcpy(o1, d_o1, o1bytes, cudaMemcpyDeviceToHost));
    if (o2) CU(cudaMemcpy(o2, d_o2, o2bytes, cudaMemcpyDeviceToHost));
  }
  cudaFree(d_a);
  cudaFree(d_b);
  cudaFree(d_o1);
  cudaFree(d_o2);
  return rc;
}
}  // namespace tbe

extern "C" {

int tb200_test_fq_mul(const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !b || !out || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  cudaStream_t st = primary().stream;
  return with_buffers(a, n * 48, b, n * 48, out, n * 48, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    LAUNCH(k_test_fq_mul, cdiv(n, 128), 128, st, (const uint32_t*)da, (const uint32_t*)db, (uint32_t)n, (uint32_t*)d1);
    return 0;
  });
}
int tb200_test_fq_addsub(const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out_add, uint64_t* out_sub) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !b || !out_add || !out_sub || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  cudaStream_t st = primary().stream;
  return with_buffers(a, n * 48, b, n * 48, out_add, n * 48, out_sub, n * 48, [&](char* da, char* db, char* d1, char* d2) {
    LAUNCH(k_test_fq_addsub, cdiv(n, 128), 128, st, (const uint32_t*)da, (const uint32_t*)db, (uint32_t)n, (uint32_t*)d1,
           (uint32_t*)d2);
    return 0;
  });
}
int tb200_test_g1_add(const uint64_t* p_xy, const uint64_t* q_xy, size_t n, uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!p_xy || !q_xy || !out_xy || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  cudaStream_t st = primary().stream;
  return with_buffers(p_xy, n * 96, q_xy, n * 96, out_xy, n * 96, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    LAUNCH(k_test_g1_add, cdiv(n, 64), 64, st, (const uint4*)da, (const uint4*)db, (uint32_t)n, (uint4*)d1);
    return 0;
  });
}
int tb200_test_g1_mul(const uint64_t* p_xy, const uint64_t* k, size_t n, uint64_t* out_xy) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!p_xy || !k || !out_xy || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  cudaStream_t st = primary().stream;
  return with_buffers(p_xy, n * 96, k, n * 32, out_xy, n * 96, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    LAUNCH(k_test_g1_mul, cdiv(n, 64), 64, st, (const uint4*)da, (const uint32_t*)db, (uint32_t)n, (uint4*)d1);
    return 0;
  });
}

}  // extern "C"
