// k_accumulate_s: the bucket-accumulation loop with its field elements in SHARED-MEMORY SLOTS.
//
// Same algorithm, entries, outputs and lazy-reduction bounds as k_accumulate / xyzz_madd_fast (g1_fast.cuh). The
// difference is where operands live: the accumulator (X, Y, ZZ, ZZZ) and the five temporaries of a mixed addition
// occupy 9 slots of 48 B per thread in shared memory, laid out [slot][16-byte group][thread] so a warp's LDS.128 /
// STS.128 touch 512 consecutive bytes (conflict-free). The out-of-line multiplier loads its two operands with six
// LDS.128 and stores the product with three STS.128 -- on the otherwise idle LSU pipe -- instead of receiving them
// through ~36 call-boundary IMAD.MOVs on the FMA pipe (ncu/SASS: 373 of them per mixed addition, ~6% of the
// integer-pipe time), nothing spills, and the freed registers hold the prefetched next point.
#pragma once
#include "kernels.cuh"
#ifdef TB_EXPERIMENTAL_KARATSUBA  // measured 3 % slower on sm_100a; not part of the default build
#include "experimental/mont_kara.cuh"
#endif

namespace tb {

constexpr int ACCS_THREADS = 128;
constexpr int ACCS_SLOTS = 9;
constexpr int ACCS_SMEM = ACCS_SLOTS * 3 * ACCS_THREADS * 16;  // 55,296 B per CTA -> 4 CTAs per SM

enum : int { SX = 0, SY = 1, SZZ = 2, SZZZ = 3, SA = 4, SB = 5, SPP = 6, SRR = 7, SQQ = 8 };

__device__ __forceinline__ void slot_load(Fq& v, const uint4* sm, int slot) {
  const uint4* p = sm + slot * 3 * ACCS_THREADS;
  uint4 a = p[0], b = p[ACCS_THREADS], c = p[2 * ACCS_THREADS];
  v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
  v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
  v.l[8] = c.x; v.l[9] = c.y; v.l[10] = c.z; v.l[11] = c.w;
}
__device__ __forceinline__ void slot_store(uint4* sm, int slot, const Fq& v) {
  uint4* p = sm + slot * 3 * ACCS_THREADS;
  p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  p[ACCS_THREADS] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
  p[2 * ACCS_THREADS] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}
// d = a * b (lazy Montgomery product); `sm` already points at this thread's lane of slot 0
static __device__ __noinline__ void slot_mul(uint4* sm, int d, int a, int b) {
  Fq x, y, r;
  slot_load(x, sm, a);
  slot_load(y, sm, b);
  mont_mul_lazy<FqParams>(r.l, x.l, y.l);
  slot_store(sm, d, r);
}
#ifdef TB_EXPERIMENTAL_KARATSUBA
// Karatsuba variant (experimental/mont_kara.cuh): 240 instead of 276 wide MACs but ~170 more ALU-pipe instructions
static __device__ __noinline__ void slot_mul_k(uint4* sm, int d, int a, int b) {
  Fq x, y, r;
  slot_load(x, sm, a);
  slot_load(y, sm, b);
  mont_mul_kara(r.l, x.l, y.l);
  slot_store(sm, d, r);
}
#endif
// d = a * b + c * e, even/odd CIOS with one reduction (mont_mul2_lazy: 420 wide MACs)
static __device__ __noinline__ void slot_mul2_c(uint4* sm, int d, int a, int b, int c, int e) {
  Fq x, y, z, w, r;
  slot_load(x, sm, a);
  slot_load(y, sm, b);
  slot_load(z, sm, c);
  slot_load(w, sm, e);
  mont_mul2_lazy<FqParams>(r.l, x.l, y.l, z.l, w.l);
  slot_store(sm, d, r);
}
// V: bit 0 = Karatsuba for the single products (measured: no gain -- on sm_100a ALU-pipe instructions do not issue for
// free next to IMAD.WIDE, benches/pipes.cu); bit 1 = Y3 as ONE fused sum of two products (-132 wide MACs, default)
template <int V>
__device__ __forceinline__ void slot_mulv(uint4* sm, int d, int a, int b) {
#ifdef TB_EXPERIMENTAL_KARATSUBA
  if ((V & 1) != 0) {
    slot_mul_k(sm, d, a, b);
    return;
  }
#endif
  slot_mul(sm, d, a, b);
}
// d = a * a (dedicated squaring)
static __device__ __noinline__ void slot_sqr(uint4* sm, int d, int a) {
  Fq x, r;
  slot_load(x, sm, a);
  mont_sqr_lazy<FqParams>(r.l, x.l);
  slot_store(sm, d, r);
}
// d = a + k*q - b; returns the low limb of the result (for the P = 0 mod q filter)
template <int SEL>
__device__ __forceinline__ uint32_t slot_sub(uint4* sm, int d, int a, int b) {
  Fq x, y;
  slot_load(x, sm, a);
  slot_load(y, sm, b);
  fq_sub_lazy<SEL>(x, x, y);
  slot_store(sm, d, x);
  return x.l[0];
}

// acc (slots SX..SZZZ) += q; `inf` tracks whether the accumulator is the identity. Bounds as in xyzz_madd_fast.
// d = 4q - a  (a <= 4q)
__device__ __forceinline__ void slot_negy(uint4* sm, int d, int a) {
  Fq x, r;
  slot_load(x, sm, a);
  Carry c;
  r.l[0] = sub_cc(fq_kq(1, 0), x.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = subc_cc(fq_kq(1, i), x.l[i], c);
  slot_store(sm, d, r);
}

template <int V>
__device__ __forceinline__ void madd_slots(uint4* sm, const Affine& q, bool& inf) {
  if (affine_is_inf(q)) return;
  if (inf) {
    slot_store(sm, SX, q.x);
    slot_store(sm, SY, q.y);
    Fq one = fq_one();
    slot_store(sm, SZZ, one);
    slot_store(sm, SZZZ, one);
    inf = false;
    return;
  }
  slot_store(sm, SA, q.x);
  slot_store(sm, SB, q.y);
  slot_mulv<V>(sm, SPP, SA, SZZ);                 // U2
  slot_mulv<V>(sm, SRR, SB, SZZZ);                // S2
  const uint32_t plo = slot_sub<2>(sm, SPP, SPP, SX);   // P = U2 + 8q - X1
  slot_sub<1>(sm, SRR, SRR, SY);                  // R = S2 + 4q - Y1
  if (plo - 1u < 9u) {                            // P = k q possible: decide exactly, out of line
    Fq chk;
    slot_load(chk, sm, SPP);
    fq_canon(chk);
    if (fq_is_zero(chk)) {
      Xyzz p;
      slot_load(p.x, sm, SX);
      slot_load(p.y, sm, SY);
      slot_load(p.zz, sm, SZZ);
      slot_load(p.zzz, sm, SZZZ);
      p = xyzz_madd_exact(p, q);
      slot_store(sm, SX, p.x);
      slot_store(sm, SY, p.y);
      slot_store(sm, SZZ, p.zz);
      slot_store(sm, SZZZ, p.zzz);
      inf = xyzz_is_inf(p);
      return;
    }
  }
  slot_sqr(sm, SA, SPP);                          // PP   (slot A)
  slot_mulv<V>(sm, SB, SPP, SA);                  // PPP  (slot B)
  slot_mulv<V>(sm, SQQ, SX, SA);                  // Q
  slot_mulv<V>(sm, SZZ, SZZ, SA);                 // ZZ3
  slot_mulv<V>(sm, SZZZ, SZZZ, SB);               // ZZZ3
  slot_sqr(sm, SA, SRR);                          // RR   (slot A)
  slot_sub<0>(sm, SA, SA, SB);                    // RR + 2q - PPP
  slot_sub<0>(sm, SA, SA, SQQ);                   //    + 2q - Q
  slot_sub<0>(sm, SX, SA, SQQ);                   // X3 < 7.3q
  slot_sub<2>(sm, SQQ, SQQ, SX);                  // Q + 8q - X3 < 9.2q
  if ((V & 2) == 0) {
    slot_mulv<V>(sm, SQQ, SRR, SQQ);              // R (Q - X3)
    slot_mulv<V>(sm, SA, SY, SB);                 // Y1 PPP
    slot_sub<0>(sm, SY, SQQ, SA);                 // Y3
  } else {
    // Y3 = R (Q - X3) + (4q - Y1) PPP with ONE Montgomery reduction: R < 6q, Q - X3 < 9.2q, 4q - Y1 <= 4q,
    // PPP < 1.2q  =>  Y3 < q + (55.2 + 4.8) q^2 / 2^384 < 1.5q   (invariant Y < 4q)
    slot_negy(sm, SA, SY);                        // 4q - Y1
    slot_mul2_c(sm, SY, SA, SB, SRR, SQQ);        // b + d = PPP + (Q - X3) < 10.4q (mont_mul2_lazy's bound)
  }
}

__device__ __forceinline__ void flush_slots(uint4* sm, uint4* dst, bool inf) {
  Xyzz acc;
  if (inf) {
    xyzz_set_inf(acc);
  } else {
    slot_load(acc.x, sm, SX);
    slot_load(acc.y, sm, SY);
    slot_load(acc.zz, sm, SZZ);
    slot_load(acc.zzz, sm, SZZZ);
    // stored lazily reduced (X < 8q, Y < 4q, ZZ, ZZZ < 2q): every consumer (k_fixup_*, k_reduce_pass) takes the
    // lazy invariant, and a flush is a divergent branch of the hot loop, so it should be as short as possible
  }
  store_xyzz(dst, acc);
}

// accumulator slots <- stored XYZZ bucket; returns true if it is the identity
__device__ __forceinline__ bool load_bucket_slots(uint4* sm, const uint4* src) {
  uint4 v[12];
#pragma unroll
  for (int i = 0; i < 12; i++) v[i] = src[i];
#pragma unroll
  for (int i = 0; i < 12; i++) sm[(i / 3) * 3 * ACCS_THREADS + (i % 3) * ACCS_THREADS] = v[i];  // slots SX..SZZZ = 0..3
  return (v[6].x | v[6].y | v[6].z | v[6].w | v[7].x | v[7].y | v[7].z | v[7].w | v[8].x | v[8].y | v[8].z | v[8].w) == 0;
}

template <int V>
__global__ void __launch_bounds__(ACCS_THREADS, 4)
    k_accumulate_s(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ bucket_start, uint32_t B,
                   uint32_t K, const uint4* __restrict__ points, uint4* __restrict__ buckets,
                   uint4* __restrict__ heads, int32_t* __restrict__ head_bucket, int merge) {
  extern __shared__ uint4 s_slots[];
  uint4* sm = s_slots + threadIdx.x;
  const uint32_t M = __ldg(bucket_start + B);
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t lo64 = t * K;
  if (lo64 >= M) return;
  const uint32_t lo = (uint32_t)lo64;
  const uint32_t hi = (uint32_t)min((uint64_t)M, lo64 + K);
  uint32_t l = 0, r = B;
  while (r - l > 1) {
    uint32_t mid = l + ((r - l) >> 1);
    if (__ldg(bucket_start + mid) <= lo) l = mid;
    else r = mid;
  }
  uint32_t b = l;
  bool is_head = __ldg(bucket_start + b) < lo;
  head_bucket[t] = is_head ? (int32_t)b : -1;
  uint32_t end_b = __ldg(bucket_start + b + 1);
  bool inf = true;
  // `merge`: buckets[] already holds the sums of earlier point-range chunks of the same MSM (identity = all zero):
  // the thread that owns the first entry of a bucket continues from the stored value -- no extra group operation
  if (merge && !is_head) inf = load_bucket_slots(sm, buckets + 12 * (uint64_t)b);
  // software pipeline: the point of entry pos+1 is gathered into registers while entry pos is being added
  Affine nxt;
  uint32_t e = __ldg(entries + lo);
  load_fq2_nc(nxt, points + 6 * (uint64_t)(e & 0x7fffffffu));
  for (uint32_t pos = lo; pos < hi; pos++) {
    if (pos == end_b) {
      flush_slots(sm, is_head ? heads + 12 * t : buckets + 12 * (uint64_t)b, inf);
      is_head = false;
      inf = true;
      do {
        b++;
        end_b = __ldg(bucket_start + b + 1);
      } while (end_b == pos);
      if (merge) inf = load_bucket_slots(sm, buckets + 12 * (uint64_t)b);
    }
    Affine q = nxt;
    const uint32_t neg = e >> 31;
    if (pos + 1 < hi) {
      e = __ldg(entries + pos + 1);
      load_fq2_nc(nxt, points + 6 * (uint64_t)(e & 0x7fffffffu));
    }
    if (neg) fq_neg(q.y, q.y);
    madd_slots<V>(sm, q, inf);
  }
  flush_slots(sm, is_head ? heads + 12 * t : buckets + 12 * (uint64_t)b, inf);
}

}  // namespace tb
