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
#include "kernels.cuh"

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

// load input i of the current round: round 0 gathers +-points[entries[i]], later rounds read cur[i]
template <bool FIRST>
__device__ __forceinline__ void load_input(Affine& a, uint32_t i, const uint32_t* __restrict__ entries,
                                           const uint4* __restrict__ points, const uint4* __restrict__ cur) {
  if (FIRST) {
    const uint32_t e = __ldg(entries + i);
    load_fq2_nc(a, points + 6 * (uint64_t)(e & 0x7fffffffu));
    if (e >> 31) fq_neg(a.y, a.y);
  } else {
    load_affine(a, cur + 6 * (uint64_t)i);
  }
}

// classification of one output: what the batch inversion has to invert and which formula applies
enum : uint32_t { PAIR_COPY0 = 0, PAIR_COPY1 = 1, PAIR_INF = 2, PAIR_ADD = 3, PAIR_DBL = 4 };
__device__ __forceinline__ uint32_t classify(const Affine& a, const Affine& b, bool has_partner, Fq& denom) {
  if (!has_partner || affine_is_inf(b)) return PAIR_COPY0;
  if (affine_is_inf(a)) return PAIR_COPY1;
  fq_sub(denom, b.x, a.x);
  if (!fq_is_zero(denom)) return PAIR_ADD;
  if (fq_eq(a.y, b.y) && !fq_is_zero(a.y)) {
    fq_dbl(denom, a.y);
    return PAIR_DBL;
  }
  return PAIR_INF;  // P + (-P) (or a 2-torsion point doubled, which cannot occur in the prime-order subgroup)
}

// One round. Outputs are dealt to the lanes of a warp round-robin (output = base + i*32 + lane) so that the lanes'
// accesses to cur[]/nxt[]/scratch stay adjacent; every thread handles T outputs.
template <bool FIRST>
__global__ void __launch_bounds__(128, 3)
    k_affine_round(const uint32_t* __restrict__ pidx, const uint32_t* __restrict__ off_out, uint32_t B, uint32_t T,
                   const uint32_t* __restrict__ entries, const uint4* __restrict__ points,
                   const uint4* __restrict__ cur, uint4* __restrict__ nxt, uint4* __restrict__ scratch) {
  const uint32_t n_out = off_out[B];
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t base = warp * 32ull * T;
  if (base >= n_out) return;
  // forward: prefix products of the denominators
  Fq run = fq_one();
  for (uint32_t i = 0; i < T; i++) {
    const uint64_t p = base + (uint64_t)i * 32 + lane;
    if (p >= n_out) break;
    const uint32_t pi = __ldg(pidx + p);
    const uint32_t i0 = pi & 0x7fffffffu;
    const bool partner = pi >> 31;
    Affine a, b;
    load_input<FIRST>(a, i0, entries, points, cur);
    if (partner) load_input<FIRST>(b, i0 + 1, entries, points, cur);
    Fq d;
    const uint32_t kind = classify(a, b, partner, d);
    // prefix BEFORE this output (what the backward sweep multiplies the running inverse with)
    uint4* sp = scratch + 3 * p;
    sp[0] = make_uint4(run.l[0], run.l[1], run.l[2], run.l[3]);
    sp[1] = make_uint4(run.l[4], run.l[5], run.l[6], run.l[7]);
    sp[2] = make_uint4(run.l[8], run.l[9], run.l[10], run.l[11]);
    if (kind >= PAIR_ADD) run = fq_mulc_call(run, d);
  }
  Fq inv = fq_inv_call(run);
  // backward
  uint32_t cnt = 0;
  {
    const uint64_t first = base + lane;
    if (first < n_out) cnt = (uint32_t)min((uint64_t)T, (n_out - first + 31) / 32);
  }
  for (uint32_t i = cnt; i-- > 0;) {
    const uint64_t p = base + (uint64_t)i * 32 + lane;
    const uint32_t pi = __ldg(pidx + p);
    const uint32_t i0 = pi & 0x7fffffffu;
    const bool partner = pi >> 31;
    Affine a, b, o;
    load_input<FIRST>(a, i0, entries, points, cur);
    if (partner) load_input<FIRST>(b, i0 + 1, entries, points, cur);
    Fq d;
    const uint32_t kind = classify(a, b, partner, d);
    if (kind == PAIR_COPY0) {
      o = a;
    } else if (kind == PAIR_COPY1) {
      o = b;
    } else if (kind == PAIR_INF) {
      o.x = fq_zero();
      o.y = fq_zero();
    } else {
      Fq pre;
      const uint4* sp = scratch + 3 * p;
      uint4 v0 = sp[0], v1 = sp[1], v2 = sp[2];
      pre.l[0] = v0.x; pre.l[1] = v0.y; pre.l[2] = v0.z; pre.l[3] = v0.w;
      pre.l[4] = v1.x; pre.l[5] = v1.y; pre.l[6] = v1.z; pre.l[7] = v1.w;
      pre.l[8] = v2.x; pre.l[9] = v2.y; pre.l[10] = v2.z; pre.l[11] = v2.w;
      Fq dinv = fq_mulc_call(inv, pre);  // 1 / d
      inv = fq_mulc_call(inv, d);        // running inverse without this d
      Fq num, lam, t;
      if (kind == PAIR_ADD) {
        fq_sub(num, b.y, a.y);
      } else {  // doubling: 3 x^2
        t = fq_mulc_call(a.x, a.x);
        fq_dbl(num, t);
        fq_add(num, num, t);
        b.x = a.x;
      }
      lam = fq_mulc_call(num, dinv);
      t = fq_mulc_call(lam, lam);
      fq_sub(t, t, a.x);
      fq_sub(o.x, t, b.x);     // x3 = lambda^2 - x1 - x2
      fq_sub(t, a.x, o.x);
      t = fq_mulc_call(lam, t);
      fq_sub(o.y, t, a.y);     // y3 = lambda (x1 - x3) - y1
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
