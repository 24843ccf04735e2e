// Batched-affine bucket accumulation (alternative to k_accumulate for large inputs).
//
// The sorted entries of every bucket are summed by ROUNDS OF PAIRWISE AFFINE ADDITIONS: round r turns a bucket of
// n points into ceil(n/2) points (pairs (2k, 2k+1) are added, an odd leftover is copied), until every non-empty
// bucket holds one affine point. An affine addition needs one field inversion; Montgomery's trick shares one
// inversion among all additions a thread performs in a round:
//     forward : prefix products of the denominators d_i (1M each), prefix stored to scratch
//     invert  : one Fermat inversion per thread per round (~480M, amortised over T >= 256 additions)
//     backward: 1/d_i from the running inverse (2M), lambda = num_i / d_i (1M), x3 = lambda^2 - x1 - x2 (1S),
//               y3 = lambda (x1 - x3) - y1 (1M)
// => ~6M + 480/T per addition instead of the 8M+2S of an XYZZ mixed addition; the result of every round is affine,
// so no accumulator state is carried and no cross-thread fix-up exists: every output point is independent, threads
// own fixed-size, perfectly balanced ranges of OUTPUT positions whatever the bucket sizes are.
// Exceptional cases are folded into the same inversion batch:
//     x1 == x2, y1 == y2  (P + P)     : num = 3 x1^2, d = 2 y1        (same x3/y3 formulas)
//     x1 == x2, y1 == -y2 (P + (-P))  : result = identity, d := 1
//     either input identity / no partner: result = the other input, d := 1
// All coordinates here are canonical (equality tests are exact); identity = (0, 0).
#pragma once
#include "../kernels.cuh"

namespace tb {

// canonical out-of-line multiplier for this file (keeps the round kernel's loops compact)
__device__ __noinline__ Fq fq_mulc_call(Fq a, Fq b) {
  Fq r;
  mont_mul<FqParams>(r.l, a.l, b.l);
  return r;
}
// a^(q-2) with a fixed 4-bit window: 377 squarings + ~95 multiplications
__device__ __noinline__ Fq fq_inv_call(Fq a) {
  Fq tab[16];
  tab[0] = fq_one();
  tab[1] = a;
#pragma unroll 1
  for (int i = 2; i < 16; i++) tab[i] = fq_mulc_call(tab[i - 1], a);
  Fq acc = fq_one();
  bool started = false;
#pragma unroll 1
  for (int i = 11; i >= 0; i--) {
    uint32_t w = FqParams::p(i);
    if (i == 0) w = 0xffffffffu;  // q - 2: low limb 1 - 2 borrows from limb 1
    if (i == 1) w -= 1;
#pragma unroll 1
    for (int nib = 7; nib >= 0; nib--) {
      if (started) {
        acc = fq_mulc_call(acc, acc);
        acc = fq_mulc_call(acc, acc);
        acc = fq_mulc_call(acc, acc);
        acc = fq_mulc_call(acc, acc);
      }
      uint32_t d = (w >> (4 * nib)) & 15u;
      if (d) {
        acc = started ? fq_mulc_call(acc, tab[d]) : tab[d];
        started = true;
      }
    }
  }
  return acc;
}

// sizes of the next round: n' = ceil(n / 2); also tracks nothing else (the host halves the known maximum itself)
__global__ void __launch_bounds__(256) k_half_sizes(const uint32_t* __restrict__ off, uint32_t B,
                                                    uint32_t* __restrict__ sizes) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  sizes[b] = (off[b + 1] - off[b] + 1u) >> 1;
}
// largest bucket of the first round (read back once by the host to know the number of rounds)
__global__ void __launch_bounds__(256) k_max_size(const uint32_t* __restrict__ off, uint32_t B,
                                                  uint32_t* __restrict__ out_max) {
  uint32_t m = 0;
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x)
    m = max(m, off[b + 1] - off[b]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out_max, m);
}

// pidx[p] = index of the first input of output p in the current array | (has_partner << 31)
// Thread t covers outputs [8t, 8t+8): binary search for the bucket of the first one, then walk.
__global__ void __launch_bounds__(256) k_pair_index(const uint32_t* __restrict__ off_in,
                                                    const uint32_t* __restrict__ off_out, uint32_t B,
                                                    uint32_t* __restrict__ pidx) {
  const uint32_t n_out = off_out[B];
  const uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (p0 >= n_out) return;
  uint32_t l = 0, r = B;  // off_out[l] <= p0 < off_out[r]
  while (r - l > 1) {
    uint32_t mid = l + ((r - l) >> 1);
    if (off_out[mid] <= (uint32_t)p0) l = mid;
    else r = mid;
  }
  uint32_t b = l;
  uint32_t out_lo = off_out[b], out_hi = off_out[b + 1], in_lo = off_in[b], in_hi = off_in[b + 1];
  const uint32_t pend = (uint32_t)min((uint64_t)n_out, p0 + 8);
  for (uint32_t p = (uint32_t)p0; p < pend; p++) {
    while (p >= out_hi) {  // next non-empty output bucket
      b++;
      out_lo = out_hi;
      out_hi = off_out[b + 1];
      in_lo = off_in[b];
      in_hi = off_in[b + 1];
    }
    const uint32_t i0 = in_lo + 2 * (p - out_lo);
    pidx[p] = i0 | ((i0 + 1 < in_hi) ? 0x80000000u : 0u);
  }
}

// ---- round kernel, second generation --------------------------------------------------------------------------
// What ncu showed about the first version (profiles/r01_summary.md 3): both sweeps gathered both full points
// (105 GB of DRAM traffic at 2^24) behind three dependent loads (pidx -> entries -> point) with one multiplication
// of work per forward iteration, so `long_scoreboard` dominated. Now
//   * the forward sweep reads x-coordinates only (y is touched only when x0 == x1) and stores the classification,
//     so the backward sweep neither re-tests nor re-derives anything;
//   * the dependent load chain is software-pipelined three deep (index, entry, coordinates of the NEXT outputs are
//     in flight while the current product is multiplied);
//   * field work is lazily reduced with the out-of-line multiplier / squarer of the XYZZ loop; only the stored
//     output coordinates are made canonical (exact equality tests in the next round stay trivial).
__device__ __forceinline__ void load_fq_nc(Fq& v, const uint4* __restrict__ src) {
  uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
  v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
  v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
  v.l[8] = c.x; v.l[9] = c.y; v.l[10] = c.z; v.l[11] = c.w;
}
__device__ __forceinline__ void load_fq(Fq& v, const uint4* src) {
  uint4 a = src[0], b = src[1], c = src[2];
  v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
  v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
  v.l[8] = c.x; v.l[9] = c.y; v.l[10] = c.z; v.l[11] = c.w;
}
__device__ __forceinline__ void store_fq(uint4* dst, const Fq& v) {
  dst[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  dst[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
  dst[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}
enum : uint32_t { PAIR_COPY0 = 0, PAIR_COPY1 = 1, PAIR_INF = 2, PAIR_ADD = 3, PAIR_DBL = 4 };

// address of point `i` of the round's input (round 0: through entries[]); *neg = sign bit of the entry
template <bool FIRST>
__device__ __forceinline__ const uint4* input_ptr(uint32_t i, const uint32_t* __restrict__ entries,
                                                  const uint4* __restrict__ points, const uint4* __restrict__ cur,
                                                  uint32_t* neg) {
  if (FIRST) {
    const uint32_t e = __ldg(entries + i);
    *neg = e >> 31;
    return points + 6 * (uint64_t)(e & 0x7fffffffu);
  }
  *neg = 0;
  return cur + 6 * (uint64_t)i;
}
// rare path of the forward sweep: x0 == x1 (or an identity input): look at y and decide
template <bool FIRST>
__device__ __noinline__ uint32_t classify_slow(uint32_t i0, const uint32_t* __restrict__ entries,
                                               const uint4* __restrict__ points, const uint4* __restrict__ cur,
                                               Fq* denom) {
  uint32_t n0, n1;
  Affine a, b;
  load_affine(a, input_ptr<FIRST>(i0, entries, points, cur, &n0));
  load_affine(b, input_ptr<FIRST>(i0 + 1, entries, points, cur, &n1));
  if (n0) fq_neg(a.y, a.y);
  if (n1) fq_neg(b.y, b.y);
  if (affine_is_inf(b)) return PAIR_COPY0;
  if (affine_is_inf(a)) return PAIR_COPY1;
  if (!fq_eq(a.x, b.x)) {  // one of them had x == 0 without being the identity
    fq_sub(*denom, b.x, a.x);
    return PAIR_ADD;
  }
  if (fq_eq(a.y, b.y) && !fq_is_zero(a.y)) {
    fq_dbl(*denom, a.y);
    return PAIR_DBL;
  }
  return PAIR_INF;
}

template <bool FIRST>
__global__ void __launch_bounds__(128, 3)
    k_affine_round(const uint32_t* __restrict__ pidx, const uint32_t* __restrict__ off_out, uint32_t B, uint32_t T,
                   const uint32_t* __restrict__ entries, const uint4* __restrict__ points,
                   const uint4* __restrict__ cur, uint4* __restrict__ nxt, uint4* __restrict__ scratch,
                   uint8_t* __restrict__ kinds) {
  const uint32_t n_out = off_out[B];
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t base = warp * 32ull * T;
  if (base + lane >= n_out) return;
  const uint32_t cnt = (uint32_t)min((uint64_t)T, (n_out - (base + lane) + 31) / 32);
  // ---- forward sweep: prefix products of the denominators, three-stage software pipeline -------------------
  // stage A (2 ahead): pair index; stage B (1 ahead): addresses (round 0: entries) and the two x loads
  uint32_t pi_next2 = cnt > 1 ? __ldg(pidx + base + 32 + lane) : 0;
  uint32_t pi_next = __ldg(pidx + base + lane);
  Fq x0n, x1n;
  {
    uint32_t ng;
    const uint32_t i0 = pi_next & 0x7fffffffu;
    load_fq_nc(x0n, input_ptr<FIRST>(i0, entries, points, cur, &ng));
    if (pi_next >> 31) load_fq_nc(x1n, input_ptr<FIRST>(i0 + 1, entries, points, cur, &ng));
    else x1n = x0n;
  }
  Fq run = fq_one();
  for (uint32_t i = 0; i < cnt; i++) {
    const uint64_t p = base + (uint64_t)i * 32 + lane;
    const uint32_t pi = pi_next;
    Fq x0 = x0n, x1 = x1n;
    pi_next = pi_next2;
    if (i + 2 < cnt) pi_next2 = __ldg(pidx + p + 64);
    if (i + 1 < cnt) {  // issue the coordinate loads of the next output before working on this one
      uint32_t ng;
      const uint32_t j0 = pi_next & 0x7fffffffu;
      load_fq_nc(x0n, input_ptr<FIRST>(j0, entries, points, cur, &ng));
      if (pi_next >> 31) load_fq_nc(x1n, input_ptr<FIRST>(j0 + 1, entries, points, cur, &ng));
      else x1n = x0n;
    }
    uint32_t kind;
    Fq d;
    if (!(pi >> 31)) {
      kind = PAIR_COPY0;
    } else {
      fq_sub(d, x1, x0);  // inputs are canonical
      kind = PAIR_ADD;
      if (fq_is_zero(d) || fq_is_zero(x0) || fq_is_zero(x1)) kind = classify_slow<FIRST>(pi & 0x7fffffffu, entries, points, cur, &d);
    }
    store_fq(scratch + 3 * p, run);  // prefix BEFORE this output
    kinds[p] = (uint8_t)kind;
    if (kind >= PAIR_ADD) run = fq_mul_call(run, d);
  }
  fq_canon(run);
  Fq inv = fq_inv_call(run);
  // ---- backward sweep ----------------------------------------------------------------------------------------------
  for (uint32_t i = cnt; i-- > 0;) {
    const uint64_t p = base + (uint64_t)i * 32 + lane;
    const uint32_t pi = __ldg(pidx + p);
    const uint32_t i0 = pi & 0x7fffffffu;
    const uint32_t kind = kinds[p];
    Affine a, b, o;
    uint32_t n0 = 0, n1 = 0;
    if (kind != PAIR_COPY1 && kind != PAIR_INF) {
      load_fq2_nc(a, input_ptr<FIRST>(i0, entries, points, cur, &n0));
      if (n0) fq_neg(a.y, a.y);
    }
    if (kind == PAIR_COPY1 || kind == PAIR_ADD) {
      load_fq2_nc(b, input_ptr<FIRST>(i0 + 1, entries, points, cur, &n1));
      if (n1) fq_neg(b.y, b.y);
    }
    if (kind == PAIR_COPY0) {
      o = a;
    } else if (kind == PAIR_COPY1) {
      o = b;
    } else if (kind == PAIR_INF) {
      o.x = fq_zero();
      o.y = fq_zero();
    } else {
      Fq pre, d, num, lam, t;
      load_fq(pre, scratch + 3 * p);
      if (kind == PAIR_ADD) {
        fq_sub(d, b.x, a.x);
        fq_sub_lazy<0>(num, b.y, a.y);   // y1 + 2q - y0 < 3q
      } else {                           // doubling: d = 2 y0, num = 3 x0^2
        fq_dbl(d, a.y);
        t = fq_sqr_call(a.x);            // < 2q
        Carry c;
        num.l[0] = add_cc(t.l[0], t.l[0], c);
#pragma unroll
        for (int k = 1; k < 12; k++) num.l[k] = addc_cc(t.l[k], t.l[k], c);
        Carry c2;
        num.l[0] = add_cc(num.l[0], t.l[0], c2);
#pragma unroll
        for (int k = 1; k < 12; k++) num.l[k] = addc_cc(num.l[k], t.l[k], c2);   // < 6q
        b.x = a.x;
      }
      Fq dinv = fq_mul_call(inv, pre);   // 1 / d                 < 2q
      inv = fq_mul_call(inv, d);         // running inverse        < 2q
      lam = fq_mul_call(num, dinv);      //                        < 2q
      t = fq_sqr_call(lam);              //                        < 2q
      fq_sub_lazy<0>(t, t, a.x);         // + 2q - x0              < 4q
      fq_sub_lazy<0>(o.x, t, b.x);       // x3 = .. + 2q - x1      < 6q
      fq_sub_lazy<2>(t, a.x, o.x);       // x0 + 8q - x3           < 9q
      t = fq_mul_call(lam, t);           //                        < 2q
      fq_sub_lazy<0>(o.y, t, a.y);       // y3 = .. + 2q - y0      < 4q
      fq_canon(o.x);
      fq_canon(o.y);
    }
    store_affine(nxt + 6 * p, o);
  }
}

// level 0 of the bucket reduction when the buckets are single affine points (after the rounds): same running-sum
// scheme as k_reduce_pass, `run += B_i` is a mixed addition
__global__ void __launch_bounds__(128) k_reduce_pass0_affine(const uint4* __restrict__ pts,
                                                             const uint32_t* __restrict__ off,
                                                             uint4* __restrict__ outS, uint4* __restrict__ outW,
                                                             uint32_t L, uint64_t total_out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_out) return;
  Xyzz run, acc;
  xyzz_set_inf(run);
  xyzz_set_inf(acc);
  for (int i = (int)L - 1; i >= 0; i--) {
    const uint64_t idx = t * L + i;
    const uint32_t o0 = off[idx];
    if (off[idx + 1] != o0) {
      Affine a;
      load_affine(a, pts + 6 * (uint64_t)o0);
      xyzz_madd_fast_ni(&run, &a);
    }
    if (i > 0) xyzz_add_fast_ni(&acc, &run);
  }
  store_xyzz(outS + 12 * t, run);
  xyzz_add_fast_ni(&acc, &run);
  store_xyzz(outW + 12 * t, acc);
}

}  // namespace tb
