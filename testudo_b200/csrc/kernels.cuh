// MSM pipeline kernels (sm_100a). One pipeline serves every reference call site:
//
//   scalars --k_digits<COUNT>--> bucket histogram --k_scan*--> bucket_start[] --k_digits<SCATTER>--> entries[]
//   entries[] --k_accumulate--> buckets[] (+ heads[]) --k_fixup--> buckets[] --k_reduce_pass*--> group sums
//   group sums --k_finalize_single (Horner over windows) | k_finalize_batch--> canonical affine results
//
// A "group" is a bucket set that reduces to one point: a window of a single MSM (group sums are combined with
// 2^(c*w) by Horner), or a whole row of a shared-base batch (window tables 2^(c*w)*G_j are precomputed, so all
// windows of a row share one bucket set -- SURVEY.md App. D).
//
// entries[] holds (sign << 31 | point_ref) sorted by global bucket id = group * nb + |digit| - 1, nb = 2^(c-1).
// The sort is a counting sort: histogram with L2 atomics, exclusive scan, scatter with returning atomics. The
// order inside a bucket is arbitrary; the result does not depend on it (group addition is commutative and every
// output is normalised to the canonical affine point).
//
// k_accumulate is the hot kernel (>95% of the time at 2^24): it is load-balanced by construction -- every thread
// owns exactly K consecutive entries of the sorted array regardless of bucket boundaries (robust against the
// skewed scalar distributions of real witnesses) and does one XYZZ mixed addition
// (8M+2S, 2750 IMAD.WIDE) per entry.
#pragma once
#include "digits.cuh"
#include "g1_fast.cuh"

namespace tb {

// ------------------------------------------------------------------------------------------------------------
// vectorised global <-> register moves (16-byte accesses; all buffers are 16-byte aligned)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_fq2_nc(Affine& p, const uint4* __restrict__ src) {
  uint32_t* d = reinterpret_cast<uint32_t*>(&p);  // the whole packed object (x, y: 24 words)
#pragma unroll
  for (int i = 0; i < 6; i++) {
    uint4 v = __ldg(src + i);
    d[4 * i + 0] = v.x;
    d[4 * i + 1] = v.y;
    d[4 * i + 2] = v.z;
    d[4 * i + 3] = v.w;
  }
}
__device__ __forceinline__ void load_affine(Affine& p, const uint4* src) {
  uint32_t* d = reinterpret_cast<uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    uint4 v = src[i];
    d[4 * i + 0] = v.x;
    d[4 * i + 1] = v.y;
    d[4 * i + 2] = v.z;
    d[4 * i + 3] = v.w;
  }
}
__device__ __forceinline__ void store_affine(uint4* dst, const Affine& p) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 6; i++) dst[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
}
__device__ __forceinline__ void load_xyzz(Xyzz& p, const uint4* src) {
  uint32_t* d = reinterpret_cast<uint32_t*>(&p);  // the whole packed object (x, y, zz, zzz: 48 words)
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint4 v = src[i];
    d[4 * i + 0] = v.x;
    d[4 * i + 1] = v.y;
    d[4 * i + 2] = v.z;
    d[4 * i + 3] = v.w;
  }
}
__device__ __forceinline__ void store_xyzz(uint4* dst, const Xyzz& p) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 12; i++) dst[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
}
static_assert(sizeof(Affine) == 96 && sizeof(Xyzz) == 192, "packed layouts");

// out-of-line group operations for the non-hot kernels (keeps their code small; they are latency-, not
// throughput-bound)
static __device__ __noinline__ void xyzz_add_ni(Xyzz* p, const Xyzz* q) { xyzz_add(*p, *q); }
static __device__ __noinline__ void xyzz_dbl_ni(Xyzz* p) { xyzz_dbl(*p); }
static __device__ __noinline__ void xyzz_madd_ni(Xyzz* p, const Affine* q) { xyzz_madd(*p, *q); }
static __device__ __noinline__ void xyzz_to_affine_ni(Affine* r, const Xyzz* p) { xyzz_to_affine(*r, *p); }
// lazy-reduction versions (g1_fast.cuh) behind one call site each: the few dozen local-memory words per call are
// noise next to 9-14 multiplications, and the kernels that use them stay small and spill-free
static __device__ __noinline__ void xyzz_add_fast_ni(Xyzz* p, const Xyzz* q) {
  Xyzz a = *p;
  xyzz_add_fast(a, *q);
  *p = a;
}
static __device__ __noinline__ void xyzz_madd_fast_ni(Xyzz* p, const Affine* q) {
  Xyzz a = *p;
  xyzz_madd_fast(a, *q);
  *p = a;
}
static __device__ __noinline__ void xyzz_dbl_fast_ni(Xyzz* p) {
  Xyzz a = *p;
  xyzz_dbl_fast(a);
  *p = a;
}

// ------------------------------------------------------------------------------------------------------------
// geometry of one MSM call
// ------------------------------------------------------------------------------------------------------------
struct MsmGeom {
  uint32_t rows, cols;          // single MSM: rows = 1, cols = n
  long long row_stride, col_stride;  // in scalars (32 bytes)
  int c, W;                     // window bits, windows per scalar
  uint32_t nb;                  // buckets per group = 2^(c-1)
  uint32_t groups;              // single: W; batch: rows
  int batch;                    // 1: shared-base batch (group = row, ref = w*cols + j); 0: single (group = w, ref = j)
  int mont;                     // scalars are Montgomery-form Fr
  uint32_t ref_base;            // single MSM processed in point-range chunks: entry ref = ref_base + col
  int w_lo, w_hi;               // single MSM: the windows [w_lo, w_hi) this pass sorts (group = w - w_lo); batch: 0, W
};

// Translation units that only need the device functions of this header (engine_g2.cu, engine_pairing.cu) define
// TB_NO_G1_KERNELS: a __global__ function must live in exactly one unit of the library.
#ifndef TB_NO_G1_KERNELS
// ------------------------------------------------------------------------------------------------------------
// digits: histogram and scatter share one body
// ------------------------------------------------------------------------------------------------------------
// `only_window` >= 0 restricts the pass to one window: the single-MSM scatter runs window by window so that the
// window's slice of entries[] (n * 4 B = 64 MB at 2^24) and its cursors stay L2-resident while they are filled;
// a fused pass scatters 4-byte writes over the whole 0.9 GB array and was 2.5x slower at c = 20.
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(const uint32_t* __restrict__ scalars, MsmGeom g,
                                                uint32_t* __restrict__ counters, uint32_t* __restrict__ entries,
                                                int only_window) {
  const uint64_t total = (uint64_t)g.rows * g.cols;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t row, col;
    if (g.col_stride == 1 || g.rows == 1) {  // columns are the unit-stride dimension
      row = (uint32_t)(t / g.cols);
      col = (uint32_t)(t % g.cols);
    } else {  // rows are (un-transposed sqrt_pst matrix: Z[(j << m_col) | i])
      col = (uint32_t)(t / g.rows);
      row = (uint32_t)(t % g.rows);
    }
    const uint4* sp = reinterpret_cast<const uint4*>(scalars + 8 * ((long long)row * g.row_stride +
                                                                   (long long)col * g.col_stride));
    uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    uint32_t s[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    if (g.mont) {
      uint32_t cnv[8];
      mont_to_canonical<FrParams>(cnv, s);
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = cnv[i];
    }
    if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) continue;  // zero scalar: nothing to add
    DigitIter it(s, g.c);
    for (int w = 0; w < g.w_hi; w++) {       // the carry chain starts at window 0 whatever the range
      int32_t d = it.next(w == g.W - 1);
      if (d == 0 || w < g.w_lo || (only_window >= 0 && w != only_window)) continue;
      uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      uint32_t group = g.batch ? row : (uint32_t)(w - g.w_lo);
      uint32_t bucket = group * g.nb + (mag - 1);
      if (SCATTER) {
        uint32_t ref = g.batch ? (uint32_t)w * g.cols + col : g.ref_base + col;
        uint32_t pos = atomicAdd(&counters[bucket], 1u);
        entries[pos] = ref | (d < 0 ? 0x80000000u : 0u);
      } else {
        atomicAdd(&counters[bucket], 1u);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// shared-base batches: per-row counting sort in shared memory
// ------------------------------------------------------------------------------------------------------------
// In a batch every entry of row i falls into row i's own nb buckets, so one CTA per row keeps the row's histogram /
// scatter cursors in shared memory (nb * 4 bytes: 16 KB at c = 13) instead of hammering L2 with 1.3e9 atomics, and
// the scattered 4-byte writes of a row stay inside its 0.6 MB slice of entries[] while the CTA lives.
//   SCATTER = false: smem histogram -> counts[row*nb ..] (plain coalesced stores, no global atomics)
//   SCATTER = true : smem cursors loaded from the scanned bucket_start[], returning smem atomics give positions
// Launch shape: ONE 1024-thread CTA per SM (the host pads the dynamic shared memory to enforce it). With many
// small CTAs ~1200 rows were in flight and their 0.6 MB output slices (~760 MB) overflowed L2: the scattered 4-byte
// writes then became partial-sector read-modify-writes in DRAM (47 ms at 2^26); 148 rows in flight stay L2-resident.
template <bool SCATTER>
__global__ void __launch_bounds__(1024) k_batch_digits(const uint32_t* __restrict__ scalars, MsmGeom g,
                                                      uint32_t* __restrict__ counters /* counts or bucket_start */,
                                                      uint32_t* __restrict__ entries) {
  extern __shared__ uint32_t s_cnt[];
  const uint32_t row = blockIdx.x;
  uint32_t* row_counters = counters + (uint64_t)row * g.nb;
  for (uint32_t k = threadIdx.x; k < g.nb; k += blockDim.x) s_cnt[k] = SCATTER ? row_counters[k] : 0u;
  __syncthreads();
  for (uint32_t col = threadIdx.x; col < g.cols; col += blockDim.x) {
    const uint4* sp = reinterpret_cast<const uint4*>(scalars + 8 * ((long long)row * g.row_stride +
                                                                   (long long)col * g.col_stride));
    uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    uint32_t s[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    if (g.mont) {
      uint32_t cnv[8];
      mont_to_canonical<FrParams>(cnv, s);
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = cnv[i];
    }
    if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) continue;
    DigitIter it(s, g.c);
    for (int w = 0; w < g.W; w++) {
      int32_t d = it.next(w == g.W - 1);
      if (d == 0) continue;
      uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      if (SCATTER) {
        uint32_t pos = atomicAdd(&s_cnt[mag - 1], 1u);
        entries[pos] = ((uint32_t)w * g.cols + col) | (d < 0 ? 0x80000000u : 0u);
      } else {
        atomicAdd(&s_cnt[mag - 1], 1u);
      }
    }
  }
  if (!SCATTER) {
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < g.nb; k += blockDim.x) row_counters[k] = s_cnt[k];
  }
}

// ------------------------------------------------------------------------------------------------------------
// exclusive scan of the bucket histogram (3 kernels; B <= 2^31). SCAN_TILE items per block.
// ------------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
    uint32_t winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = winc - ws;  // exclusive warp offsets
    if (lane == 31) *total = winc;                              // lane 31 holds the block total (>= 16 warps padded)
  }
  __syncthreads();
  uint32_t r = warp_sums[wid] + inc - v;
  __syncthreads();
  return r;
}

// pass 1: per-tile totals
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ in, uint32_t n,
                                                                 uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t total;
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) s += (base + i < n) ? in[base + i] : 0u;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// pass 2: scan the tile totals in place (single block, loops over chunks)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_offsets(uint32_t* __restrict__ tile_sums, uint32_t ntiles,
                                                                    uint32_t* __restrict__ grand_total) {
  __shared__ uint32_t total;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < ntiles; base += SCAN_THREADS) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0u;
    uint32_t ex = block_exclusive_scan(v, &total);
    if (i < ntiles) tile_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}
// pass 3: out[i] = exclusive prefix; also out[n] = total (written by the last tile)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ in, uint32_t n,
                                                             const uint32_t* __restrict__ tile_offsets,
                                                             uint32_t* __restrict__ out, uint32_t* __restrict__ out2) {
  __shared__ uint32_t total;
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    s += v[i];
  }
  uint32_t ex = block_exclusive_scan(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    if (base + i < n) {
      out[base + i] = ex;
      out2[base + i] = ex;  // second copy: the scatter cursors
    }
    ex += v[i];
  }
  if (base <= n && n < base + SCAN_ITEMS) out[n] = ex - 0;  // ex == total prefix at position n
}

// ------------------------------------------------------------------------------------------------------------
// bucket accumulation -- the hot kernel
// ------------------------------------------------------------------------------------------------------------
// Thread t owns entries [t*K, min((t+1)*K, M)). Runs of equal bucket id inside the segment are summed with
// mixed additions. A run that begins at its bucket's first entry is written to buckets[b] (each non-empty
// bucket has exactly one such writer); a run that continues a bucket begun in an earlier segment is the
// thread's "head" and goes to heads[t] / head_bucket[t] for k_fixup.
// The register-operand form below is kept for reference only (TB_EXPERIMENTAL_REG_ACCUMULATE): the shared-memory operand
// slots of k_accumulate_s (kernels_smem.cuh) are 30 % faster and the only accumulate kernel in the default build.
constexpr int ACC_THREADS = 128;

#ifdef TB_EXPERIMENTAL_REG_ACCUMULATE
__global__ void __launch_bounds__(ACC_THREADS, 3)
    k_accumulate(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ bucket_start, uint32_t B,
                 uint32_t K, const uint4* __restrict__ points, uint4* __restrict__ buckets,
                 uint4* __restrict__ heads, int32_t* __restrict__ head_bucket) {
  const uint32_t M = __ldg(bucket_start + B);  // number of sorted entries: only known on the device
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t lo64 = t * K;
  if (lo64 >= M) return;
  const uint32_t lo = (uint32_t)lo64;
  const uint32_t hi = (uint32_t)min((uint64_t)M, lo64 + K);
  // b = last bucket with bucket_start[b] <= lo (it is non-empty because bucket_start[b+1] > lo)
  uint32_t l = 0, r = B;  // invariant: bucket_start[l] <= lo < bucket_start[r]  (bucket_start[B] = M > lo)
  while (r - l > 1) {
    uint32_t mid = l + ((r - l) >> 1);
    if (__ldg(bucket_start + mid) <= lo) l = mid;
    else r = mid;
  }
  uint32_t b = l;
  bool is_head = __ldg(bucket_start + b) < lo;
  head_bucket[t] = is_head ? (int32_t)b : -1;
  uint32_t end_b = __ldg(bucket_start + b + 1);
  Xyzz acc;
  xyzz_set_inf(acc);
  for (uint32_t pos = lo; pos < hi; pos++) {
    if (pos == end_b) {  // bucket b complete (or its part inside this segment)
      xyzz_canon(acc);   // the loop keeps lazily reduced coordinates; everything downstream is canonical
      store_xyzz(is_head ? heads + 12 * t : buckets + 12 * (uint64_t)b, acc);
      is_head = false;
      xyzz_set_inf(acc);
      do {
        b++;
        end_b = __ldg(bucket_start + b + 1);
      } while (end_b == pos);  // skip empty buckets
    }
    const uint32_t e = __ldg(entries + pos);
    Affine q;
    load_fq2_nc(q, points + 6 * (uint64_t)(e & 0x7fffffffu));
    if (e >> 31) fq_neg(q.y, q.y);
    xyzz_madd_fast(acc, q);
  }
  xyzz_canon(acc);
  store_xyzz(is_head ? heads + 12 * t : buckets + 12 * (uint64_t)b, acc);
}
#endif  // TB_EXPERIMENTAL_REG_ACCUMULATE

// Heads of one bucket occupy consecutive segments first..last (first = smallest t with t*K > bucket_start[b]).
// They are summed by pointer jumping: in round r, head `first + i` with i % 2^(r+1) == 0 absorbs head
// `first + i + 2^r` (if it belongs to the same bucket). After ceil(log2(#heads)) rounds heads[first] holds the sum;
// k_fixup_final adds it to buckets[b]. Heavy buckets (skewed witnesses: most scalars 0/1; degenerate top
// windows) therefore cost log-depth instead of a sequential walk over thousands of partial sums.
// `need` (one word per round, zeroed by the host; may be null): rounds are ENQUEUED for the largest bucket that could
// exist (log2 of n / K), but with well-spread scalars no bucket has more than a few heads. Round r tells round r+1 whether
// any bucket still has a partner at distance 2^(r+1); rounds that are not needed return after one load (17 launches:
// 1.17 ms -> ~0.2 ms per pass; it is paid once per point-range chunk of a host-facing MSM).
__global__ void __launch_bounds__(128) k_fixup_round(const uint32_t* __restrict__ bucket_start, uint32_t B,
                                                     uint32_t K, uint32_t round, uint4* __restrict__ heads,
                                                     const int32_t* __restrict__ head_bucket,
                                                     uint32_t* __restrict__ need) {
  if (need && round > 0 && __ldcg(need + round) == 0) return;
  const uint32_t M = bucket_start[B];
  const uint64_t S = ((uint64_t)M + K - 1) / K;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 || t >= S) return;
  const int32_t hb = head_bucket[t];
  if (hb < 0) return;
  const uint64_t first = bucket_start[hb] / K + 1;
  const uint64_t i = t - first;
  const uint64_t stride = 1ull << round;
  if (i & (2 * stride - 1)) return;
  const uint64_t u = t + stride;
  const uint32_t bucket_end = bucket_start[hb + 1];
  if (u >= S || u * K >= bucket_end) return;  // partner is not a head of this bucket
  const uint64_t u2 = t + 2 * stride;
  if (need && u2 < S && u2 * K < bucket_end) need[round + 1] = 1u;   // some bucket has a head 2^(r+1) further on
  Xyzz acc, h;
  load_xyzz(acc, heads + 12 * t);
  load_xyzz(h, heads + 12 * u);
  xyzz_add_fast_ni(&acc, &h);
  store_xyzz(heads + 12 * t, acc);
}
__global__ void __launch_bounds__(128) k_fixup_final(const uint32_t* __restrict__ bucket_start, uint32_t B,
                                                     uint32_t K, uint4* __restrict__ buckets,
                                                     const uint4* __restrict__ heads,
                                                     const int32_t* __restrict__ head_bucket) {
  const uint32_t M = bucket_start[B];
  const uint64_t S = ((uint64_t)M + K - 1) / K;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 || t >= S) return;
  const int32_t hb = head_bucket[t];
  if (hb < 0) return;
  if (t != bucket_start[hb] / K + 1) return;  // only the first head of a bucket
  Xyzz acc, h;
  load_xyzz(acc, buckets + 12 * (uint64_t)hb);
  load_xyzz(h, heads + 12 * t);
  xyzz_add_fast_ni(&acc, &h);
  store_xyzz(buckets + 12 * (uint64_t)hb, acc);
}

// ------------------------------------------------------------------------------------------------------------
// bucket reduction: sum_k k * B_k per group by hierarchical running sums
// ------------------------------------------------------------------------------------------------------------
// Elements of a level are (S, W) pairs describing a span of `ell` original buckets: S = plain sum, W = sum with
// local weights 1..ell. Combining L consecutive elements i = 0..L-1 into a span of L*ell buckets:
//     S' = sum S_i,   W' = sum W_i + ell * sum_i i * S_i      (sum_i i*S_i by a descending running sum)
// Level 0 reads raw buckets (S_i = W_i = B_i, ell = 1; empty buckets are identified by their zero count).
__global__ void __launch_bounds__(128, 4) k_reduce_pass(const uint4* __restrict__ inS, const uint4* __restrict__ inW,
                                                     const uint32_t* __restrict__ bucket_start /* level 0 only */,
                                                     uint4* __restrict__ outS, uint4* __restrict__ outW, uint32_t L,
                                                     int log2_ell, uint64_t total_out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_out) return;
  // lazily reduced coordinates throughout (g1_fast.cuh); the stored (S, W) pairs keep the lazy invariant.
  // Two sweeps keep at most three XYZZ values live (the one-sweep form needed four and spilled).
  Xyzz run, acc, x;
  xyzz_set_inf(run);
  xyzz_set_inf(acc);
  for (int i = (int)L - 1; i >= 0; i--) {
    const uint64_t idx = t * L + i;
    bool empty = false;
    if (bucket_start) empty = bucket_start[idx + 1] == bucket_start[idx];
    if (!empty) {
      load_xyzz(x, inS + 12 * idx);
      xyzz_add_fast_ni(&run, &x);
    }
    if (i > 0) xyzz_add_fast_ni(&acc, &run);
  }
  store_xyzz(outS + 12 * t, run);
  for (int k = 0; k < log2_ell; k++) xyzz_dbl_fast_ni(&acc);
  if (inW) {
    xyzz_set_inf(run);  // reused as sum of the W_i
    for (int i = 0; i < (int)L; i++) {
      load_xyzz(x, inW + 12 * (t * L + i));
      xyzz_add_fast_ni(&run, &x);
    }
  }
  xyzz_add_fast_ni(&acc, &run);
  store_xyzz(outW + 12 * t, acc);
}

// single MSM: result = sum_w 2^(c*w) * Wsum[w]  (Horner, high window first), then canonical affine.
// The ~(W-1)*c = 240 doublings are a strictly sequential chain; one thread needs 9 dependent multiplications per
// doubling (2.7 ms at c = 20). Four lanes of one warp share a doubling instead: its products form three levels of
// mutually independent multiplications (V, X^2 | W, S, M^2 | M(S-X3), W Y, V ZZ, W ZZZ), exchanged through shared
// memory, so the chain costs three multiplication latencies per doubling. Same formulas and lazy bounds as
// xyzz_dbl_fast (g1_fast.cuh).
#endif  // TB_NO_G1_KERNELS
__device__ __forceinline__ void fq_lazy_double(Fq& r, const Fq& a) {  // r = 2a, plain limb add (a < 2^383)
  Carry c;
  r.l[0] = add_cc(a.l[0], a.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = addc_cc(a.l[i], a.l[i], c);
}
__device__ __forceinline__ void fq_lazy_add(Fq& r, const Fq& a, const Fq& b) {
  Carry c;
  r.l[0] = add_cc(a.l[0], b.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = addc_cc(a.l[i], b.l[i], c);
}
#ifndef TB_NO_G1_KERNELS
__global__ void __launch_bounds__(32) k_finalize_single(const uint4* __restrict__ group_w, int W, int c,
                                                        uint4* __restrict__ out_affine) {
  // shared operand file of the cooperative doubling; all lanes run the SAME multiplier call on different operands
  enum { PX = 0, PY, PZZ, PZZZ, U, V, WW, S, M, T0, T1, D, NV };
  __shared__ Fq sv[NV];
  __shared__ int s_inf;
  const int lane = threadIdx.x;
  if (blockIdx.x != 0) return;
  if (lane == 0) {
    Xyzz t;
    xyzz_set_inf(t);
    sv[PX] = t.x; sv[PY] = t.y; sv[PZZ] = t.zz; sv[PZZZ] = t.zzz;
  }
  __syncwarp();
  for (int w = W - 1; w >= 0; w--) {
    if (lane == 0) {
      Xyzz total, x;
      total.x = sv[PX]; total.y = sv[PY]; total.zz = sv[PZZ]; total.zzz = sv[PZZZ];
      load_xyzz(x, group_w + 12 * w);
      xyzz_add_fast_ni(&total, &x);
      sv[PX] = total.x; sv[PY] = total.y; sv[PZZ] = total.zz; sv[PZZZ] = total.zzz;
      s_inf = xyzz_is_inf(total) ? 1 : 0;
    }
    __syncwarp();
    if (w == 0 || s_inf) continue;         // 2 * identity = identity (uniform branch: s_inf is shared)
    for (int k = 0; k < c; k++) {
      if (lane == 0) {                     // U = 2Y < 8q
        Fq u;
        fq_lazy_double(u, sv[PY]);
        sv[U] = u;
      }
      __syncwarp();
      // level 1: V = U^2 (lane 0) | X^2 (lane 1)
      {
        const int ia = lane == 0 ? U : PX;
        Fq r = fq_mul_call(sv[ia], sv[ia]);
        if (lane == 0) sv[V] = r;                            // V < 1.5q
        if (lane == 1) {
          Fq t;
          fq_lazy_double(t, r);
          fq_lazy_add(r, t, r);                              // M = 3 X^2 < 4.5q
          sv[M] = r;
        }
      }
      __syncwarp();
      // level 2: W = U V (lane 0) | S = X V (lane 1) | M^2 (lane 2)
      {
        const int ia = lane == 0 ? U : (lane == 1 ? PX : M);
        const int ib = lane == 2 ? M : V;
        Fq r = fq_mul_call(sv[ia], sv[ib]);
        if (lane == 0) sv[WW] = r;                           // W < 1.1q
        if (lane == 1) sv[S] = r;                            // S < 1.1q
        if (lane == 2) sv[T0] = r;                           // M^2 < 1.2q
      }
      __syncwarp();
      if (lane == 0) {                     // X3 = M^2 + 4q - 2S < 5.2q;  D = S + 8q - X3 < 9.1q
        Fq x3, d;
        fq_sub_lazy<0>(x3, sv[T0], sv[S]);
        fq_sub_lazy<0>(x3, x3, sv[S]);
        fq_sub_lazy<2>(d, sv[S], x3);
        sv[PX] = x3;
        sv[D] = d;
      }
      __syncwarp();
      // level 3: M D (lane 0) | W Y (lane 1) | V ZZ (lane 2) | W ZZZ (lane 3)
      {
        const int ia = lane == 0 ? M : (lane == 2 ? V : WW);
        const int ib = lane == 0 ? D : (lane == 1 ? PY : (lane == 2 ? PZZ : PZZZ));
        Fq r = fq_mul_call(sv[ia], sv[ib]);
        __syncwarp();                                        // every operand has been read
        if (lane == 0) sv[T0] = r;                           // < 1.4q
        if (lane == 1) sv[T1] = r;                           // < 1.1q
        if (lane == 2) sv[PZZ] = r;                          // < 2q
        if (lane == 3) sv[PZZZ] = r;                         // < 2q
      }
      __syncwarp();
      if (lane == 0) {                     // Y3 = M D + 2q - W Y < 3.4q
        Fq y3;
        fq_sub_lazy<0>(y3, sv[T0], sv[T1]);
        sv[PY] = y3;
      }
      __syncwarp();
    }
  }
  if (lane == 0) {
    Xyzz total;
    total.x = sv[PX]; total.y = sv[PY]; total.zz = sv[PZZ]; total.zzz = sv[PZZZ];
    xyzz_canon(total);
    Affine a;
    xyzz_to_affine_ni(&a, &total);
    store_affine(out_affine, a);
  }
}
// batch: every group sum is a finished row commitment; normalise each to affine
__global__ void __launch_bounds__(128) k_finalize_batch(const uint4* __restrict__ group_w, uint32_t groups,
                                                        uint4* __restrict__ out_affine) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= groups) return;
  Xyzz x;
  load_xyzz(x, group_w + 12 * (uint64_t)t);
  xyzz_canon(x);  // group sums arrive lazily reduced
  Affine a;
  xyzz_to_affine_ni(&a, &x);
  store_affine(out_affine + 6 * (uint64_t)t, a);
}
// n == 0 / all-zero scalars shortcut and generic "write identity"
__global__ void k_write_identity(uint4* out_affine, uint32_t count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count * 6) out_affine[t] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------------------
// SRS window tables: table[w * n + j] = 2^(c*w) * G_j  (affine), w < W
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_srs_tables(const uint4* __restrict__ bases, uint32_t n, int c, int W,
                                                    uint4* __restrict__ table) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  Affine g;
  load_affine(g, bases + 6 * (uint64_t)j);
  store_affine(table + 6 * (uint64_t)j, g);
  Xyzz p;
  xyzz_from_affine(p, g);
  for (int w = 1; w < W; w++) {
    for (int k = 0; k < c; k++) xyzz_dbl_ni(&p);
    Affine a;
    xyzz_to_affine_ni(&a, &p);
    store_affine(table + 6 * ((uint64_t)w * n + j), a);
    xyzz_from_affine(p, a);  // continue from the normalised point (ZZ = 1 keeps doublings cheap and exact)
  }
}

// ------------------------------------------------------------------------------------------------------------
// MIPP fold (src/mipp.rs:354-383): a_l[i] += c * a_r[i] (affine out), y_l[i] += c_inv * y_r[i]
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_compress_g1(uint4* __restrict__ a, uint32_t split,
                                                     const uint32_t* __restrict__ scaler /* 8 limbs */, int mont) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= split) return;
  uint32_t k[8];
#pragma unroll
  for (int j = 0; j < 8; j++) k[j] = scaler[j];
  if (mont) {
    uint32_t cnv[8];
    mont_to_canonical<FrParams>(cnv, k);
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = cnv[j];
  }
  Affine l, r;
  load_affine(l, a + 6 * (uint64_t)i);
  load_affine(r, a + 6 * ((uint64_t)split + i));
  Xyzz acc;
  xyzz_set_inf(acc);
  bool started = false;
  for (int limb = 7; limb >= 0; limb--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) xyzz_dbl_ni(&acc);
      if ((k[limb] >> bit) & 1) {
        xyzz_madd_ni(&acc, &r);
        started = true;
      }
    }
  }
  xyzz_madd_ni(&acc, &l);
  Affine o;
  xyzz_to_affine_ni(&o, &acc);
  store_affine(a + 6 * (uint64_t)i, o);
}
// y_l[i] += s * y_r[i] in Fr. Representation-agnostic for Montgomery inputs; for canonical inputs the scalar is
// first lifted to Montgomery form (s * R) so that mont_mul(sR, y) = s*y stays canonical.
__global__ void __launch_bounds__(128) k_compress_fr(uint32_t* __restrict__ y, uint32_t split,
                                                     const uint32_t* __restrict__ scaler, int mont) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= split) return;
  uint32_t s[8], l[8], r[8], t[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    s[j] = scaler[j];
    l[j] = y[8 * (uint64_t)i + j];
    r[j] = y[8 * ((uint64_t)split + i) + j];
  }
  if (!mont) {
    uint32_t r2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r2[j] = FrParams::r2(j);
    mont_mul<FrParams>(t, s, r2);  // s * R
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = t[j];
  }
  mont_mul<FrParams>(t, s, r);
  mod_add<FrParams>(l, l, t);
#pragma unroll
  for (int j = 0; j < 8; j++) y[8 * (uint64_t)i + j] = l[j];
}

// ------------------------------------------------------------------------------------------------------------
// sqrt_pst scalar work: chi products, q = Z * chi, dot product -- Fr, Montgomery form
// ------------------------------------------------------------------------------------------------------------
// chis[i] = prod_j (bit_{m-1-j}(i) ? b[j] : 1 - b[j])   (Polynomial::get_chi_i, src/sqrt_pst.rs:152-166)
// subset != 0: the clear bits contribute 1 instead of 1 - b[j]: out[i] = prod_{j : bit_{m-1-j}(i)} b[j], the structured
// polynomial of MIPP (`polynomial_evaluations_from_transcript`, src/mipp.rs:159-180, with b = cs_inv)
__global__ void __launch_bounds__(128) k_fr_chis(const uint32_t* __restrict__ b, uint32_t m, uint32_t* __restrict__ out,
                                                 int subset) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << m)) return;
  uint32_t prod[8], one[8];
#pragma unroll
  for (int k = 0; k < 8; k++) prod[k] = one[k] = FrParams::one(k);
  for (uint32_t j = 0; j < m; j++) {
    uint32_t f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) f[k] = b[8 * j + k];
    if (!((i >> (m - j - 1)) & 1)) {
      if (subset) continue;
      mod_sub<FrParams>(f, one, f);
    }
    mont_mul<FrParams>(prod, prod, f);
  }
#pragma unroll
  for (int k = 0; k < 8; k++) out[8 * (uint64_t)i + k] = prod[k];
}
__device__ __forceinline__ void fr_warp_sum(uint32_t* acc) {  // butterfly sum over the 32 lanes
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t other[8];
#pragma unroll
    for (int k = 0; k < 8; k++) other[k] = __shfl_xor_sync(0xffffffffu, acc[k], o);
    mod_add<FrParams>(acc, acc, other);
  }
}
// q[j] = sum_i Z[j * cols + i] * v[i]  (get_q, src/sqrt_pst.rs:81-101: Z[(j << m_col) | i] * chis[i]); one warp per j,
// lanes read 32 consecutive scalars (1 KiB) per step.
__global__ void __launch_bounds__(256) k_fr_matvec(const uint32_t* __restrict__ Z, uint32_t rows, uint32_t cols,
                                                   const uint32_t* __restrict__ v, uint32_t* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= rows) return;
  // The kernel moves 32 B per product but a canonical Fr product is 120 wide MACs: at 2^26 products that is 0.9 ms of the
  // integer pipe against 0.33 ms of HBM time -- the pipe, not the memory, is the bound. Two products share ONE
  // Montgomery reduction (mont_mul2_lazy: 184 instead of 240 wide MACs per pair; b + d + r < 2^256 holds for canonical
  // operands, the result is below 1.15 r: one conditional subtraction).
  uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t i = lane;
  for (; i + 32 < cols; i += 64) {
    const uint4* zp = reinterpret_cast<const uint4*>(Z + 8 * (j * cols + i));
    const uint4* vp = reinterpret_cast<const uint4*>(v + 8 * (uint64_t)i);
    const uint4 z0 = __ldg(zp), z1 = __ldg(zp + 1), v0 = __ldg(vp), v1 = __ldg(vp + 1);
    const uint4 y0 = __ldg(zp + 64), y1 = __ldg(zp + 65), w0 = __ldg(vp + 64), w1 = __ldg(vp + 65);   // element i + 32
    uint32_t a[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
    uint32_t c[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t b[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
    uint32_t d[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    uint32_t t[8];
    mont_mul2_lazy<FrParams>(t, a, c, b, d);
    mod_reduce_once<FrParams>(t);
    mod_add<FrParams>(acc, acc, t);
  }
  for (; i < cols; i += 32) {
    const uint4* zp = reinterpret_cast<const uint4*>(Z + 8 * (j * cols + i));
    const uint4* vp = reinterpret_cast<const uint4*>(v + 8 * (uint64_t)i);
    uint4 z0 = __ldg(zp), z1 = __ldg(zp + 1), v0 = __ldg(vp), v1 = __ldg(vp + 1);
    uint32_t a[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
    uint32_t c[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t t[8];
    mont_mul<FrParams>(t, a, c);
    mod_add<FrParams>(acc, acc, t);
  }
  fr_warp_sum(acc);
  if (lane == 0) {
    uint4* op = reinterpret_cast<uint4*>(out + 8 * j);
    op[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    op[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// One level of `MultilinearPC::open` / `open_g1` (ark-poly-commit 0.4 multilinear_pc; call sites
// src/sqrt_pst.rs:225, src/mipp.rs:144): for b < half
//     q[b] = r[2b+1] - r[2b],      r'[b] = r[2b] (1 - p) + r[2b+1] p = r[2b] + p q[b]
// and the MSM scalars of the level are q duplicated to length 2 half (`cur_q[x >> 1]`). Everything in Montgomery form;
// canonical callers convert with k_fr_to_mont first.
__global__ void __launch_bounds__(128) k_pst_level(const uint32_t* __restrict__ r_in, uint32_t half,
                                                   const uint32_t* __restrict__ p, uint32_t* __restrict__ r_out,
                                                   uint32_t* __restrict__ q_dup) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= half) return;
  const uint4* src = reinterpret_cast<const uint4*>(r_in + 16 * (uint64_t)b);
  const uint4 l0 = src[0], l1 = src[1], h0 = src[2], h1 = src[3];
  uint32_t lo[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
  uint32_t hi[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  uint32_t pt[8], q[8], t[8];
#pragma unroll
  for (int k = 0; k < 8; k++) pt[k] = p[k];
  mod_sub<FrParams>(q, hi, lo);
  mont_mul<FrParams>(t, pt, q);
  mod_add<FrParams>(t, t, lo);
  uint4* ro = reinterpret_cast<uint4*>(r_out + 8 * (uint64_t)b);
  ro[0] = make_uint4(t[0], t[1], t[2], t[3]);
  ro[1] = make_uint4(t[4], t[5], t[6], t[7]);
  uint4* qo = reinterpret_cast<uint4*>(q_dup + 16 * (uint64_t)b);
  qo[0] = qo[2] = make_uint4(q[0], q[1], q[2], q[3]);
  qo[1] = qo[3] = make_uint4(q[4], q[5], q[6], q[7]);
}
// canonical -> Montgomery in place (a * R^2 * R^-1)
__global__ void __launch_bounds__(128) k_fr_to_mont(uint32_t* __restrict__ v, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[8], r2[8], t[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    a[k] = v[8 * (uint64_t)i + k];
    r2[k] = FrParams::r2(k);
  }
  mont_mul<FrParams>(t, a, r2);
#pragma unroll
  for (int k = 0; k < 8; k++) v[8 * (uint64_t)i + k] = t[k];
}

// ------------------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------------------
// out = sum of n affine points (n small: per-GPU partial results)
__global__ void k_g1_sum(const uint4* __restrict__ pts, uint32_t n, uint4* __restrict__ out_affine) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Xyzz acc;
  xyzz_set_inf(acc);
  for (uint32_t i = 0; i < n; i++) {
    Affine p;
    load_affine(p, pts + 6 * (uint64_t)i);
    xyzz_madd_ni(&acc, &p);
  }
  Affine a;
  xyzz_to_affine_ni(&a, &acc);
  store_affine(out_affine, a);
}
// out[i * nb + j] = A_i + B_j, affine
__global__ void __launch_bounds__(128) k_g1_outer_sum(const uint4* __restrict__ A, uint32_t na,
                                                      const uint4* __restrict__ Bp, uint32_t nb,
                                                      uint4* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)na * nb) return;
  Affine a, b;
  load_affine(a, A + 6 * (t / nb));
  load_affine(b, Bp + 6 * (t % nb));
  Xyzz acc;
  xyzz_from_affine(acc, a);
  xyzz_madd_ni(&acc, &b);
  Affine o;
  xyzz_to_affine_ni(&o, &acc);
  store_affine(out + 6 * t, o);
}

// unit-test kernels ------------------------------------------------------------------------------------------
__global__ void k_test_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t n, uint32_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x, y, z;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    x.l[j] = a[12 * (uint64_t)i + j];
    y.l[j] = b[12 * (uint64_t)i + j];
  }
  fq_mul(z, x, y);
#pragma unroll
  for (int j = 0; j < 12; j++) out[12 * (uint64_t)i + j] = z.l[j];
}
__global__ void k_test_fq_addsub(const uint32_t* a, const uint32_t* b, uint32_t n, uint32_t* oadd, uint32_t* osub) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x, y, z, w;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    x.l[j] = a[12 * (uint64_t)i + j];
    y.l[j] = b[12 * (uint64_t)i + j];
  }
  fq_add(z, x, y);
  fq_sub(w, x, y);
#pragma unroll
  for (int j = 0; j < 12; j++) {
    oadd[12 * (uint64_t)i + j] = z.l[j];
    osub[12 * (uint64_t)i + j] = w.l[j];
  }
}
__global__ void k_test_g1_add(const uint4* p, const uint4* q, uint32_t n, uint4* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine a, b, o;
  load_affine(a, p + 6 * (uint64_t)i);
  load_affine(b, q + 6 * (uint64_t)i);
  Xyzz acc;
  xyzz_from_affine(acc, a);
  xyzz_madd(acc, b);  // the inlined hot-loop version
  xyzz_to_affine_ni(&o, &acc);
  store_affine(out + 6 * (uint64_t)i, o);
}
__global__ void k_test_g1_mul(const uint4* p, const uint32_t* k, uint32_t n, uint4* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine a, o;
  load_affine(a, p + 6 * (uint64_t)i);
  Xyzz acc;
  xyzz_set_inf(acc);
  bool started = false;
  for (int limb = 7; limb >= 0; limb--) {
    const uint32_t kw = k[8 * (uint64_t)i + limb];
    for (int bit = 31; bit >= 0; bit--) {
      if (started) xyzz_dbl_ni(&acc);
      if ((kw >> bit) & 1) {
        xyzz_madd_ni(&acc, &a);
        started = true;
      }
    }
  }
  xyzz_to_affine_ni(&o, &acc);
  store_affine(out + 6 * (uint64_t)i, o);
}

// integer-pipe microbenchmarks (the roofline denominator is measured) --------------------------
// Eight chains per thread; every chain's multiplier is the low word another chain produced one step earlier, so
// nothing is loop-invariant (ptxas strength-reduces an invariant product to adds) and nothing is warp-uniform
// (ptxas moves uniform chains to the uniform datapath) -- both were observed with a naive version.
// kind 0: IMAD.WIDE.U32.X (32x32+64 -> 64 with carry in / out);  kind 1: IMAD (32x32+32 -> 32)
__global__ void __launch_bounds__(256) k_int_pipe(int kind, int iters, uint32_t seed, uint64_t* sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t b = (seed * 40503u + tid * 2654435761u) | 1u;
  if (kind == 0) {
    uint64_t c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = ((uint64_t)(tid + k) << 32) | (seed + 77u * k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          // the carry-chained pair the Montgomery products issue (SASS: IMAD.WIDE.U32.X); the plain
          // `mad.wide.u32 c, a, b, c` with a 64-bit addend issues 25 % slower (benches/pipes.cu) and under-states the peak
          uint32_t a = (uint32_t)c[(k + 1) & 7];
          uint32_t lo = (uint32_t)c[k], hi = (uint32_t)(c[k] >> 32);
          if (k == 0) asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
          else asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
          asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(a), "r"(b));
          c[k] = ((uint64_t)hi << 32) | lo;
        }
      }
    }
    sink[tid] = c[0] ^ c[1] ^ c[2] ^ c[3] ^ c[4] ^ c[5] ^ c[6] ^ c[7];
  } else {
    uint32_t c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = seed + 77u * k + tid;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          uint32_t a = c[(k + 1) & 7];
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[k]) : "r"(a), "r"(b));
        }
      }
    }
    sink[tid] = c[0] ^ c[1] ^ c[2] ^ c[3] ^ c[4] ^ c[5] ^ c[6] ^ c[7];
  }
}
// kind 2: Fq Montgomery multiplications, 2 independent per-thread chains (lane-varying operands)
__global__ void __launch_bounds__(128) k_fq_mul_peak(int iters, uint32_t seed, uint32_t* sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  Fq x, y, z, w;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    x.l[j] = (seed + j * 77u + tid * 2654435761u) & 0x00ffffffu;
    y.l[j] = (seed * 3u + j * 1013u + tid * 40503u) & 0x00ffffffu;
    z.l[j] = (tid + j) & 0x00ffffffu;
    w.l[j] = (tid * 7u + j) & 0x00ffffffu;
  }
  for (int i = 0; i < iters; i++) {
    fq_mul(z, z, x);
    fq_mul(w, w, y);
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 12; j++) acc ^= z.l[j] ^ w.l[j];
  sink[tid] = acc;
}

#endif  // TB_NO_G1_KERNELS
}  // namespace tb
