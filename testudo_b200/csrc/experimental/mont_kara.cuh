// Karatsuba Montgomery multiplication for Fq (12 x 32-bit limbs): 240 wide MACs instead of 276.
//
// Why: k_accumulate_s is bound by the FMA-heavy pipe (IMAD.WIDE issues once per 4 cycles per SM sub-partition;
// profiles/r01_summary.md: 87 % busy), while the ALU pipe that executes IADD3 / LOP3 is mostly idle. One level of
// (subtractive) Karatsuba trades 36 wide MACs of the 12 x 12 product for ~110 ALU-pipe instructions:
//     a = a0 + a1 W, b = b0 + b1 W  (W = 2^192)
//     a b = z0 + (z0 + z2 + (a0 - a1)(b1 - b0)) W + z2 W^2,   z0 = a0 b0, z2 = a1 b1
// three 6 x 6 products (36 MACs each, even/odd accumulators as in mont.cuh so every product is one IMAD.WIDE on an
// aligned register pair) -> 108, followed by a separate word-serial Montgomery reduction of the 24-limb product
// (12 rows x 11 MACs; m = -T_i and m p_0 = m need no multiplier because q = 1, -q^-1 = 2^32 - 1 mod 2^32).
// mont_mul2_kara reduces a b + c d ONCE (348 MACs instead of 2 x 240): Y3 = R (Q - X3) - Y1 PPP of the mixed
// addition is such a sum.
//
// Same contract as mont_mul_lazy: the result is NOT fully reduced, r < q + T / 2^384 for the reduced total T, and
// must stay below 2^384 (callers track the bounds; g1_fast.cuh). Host build emulates the carry chains
// (tests/host_check).
#pragma once
#include "../mont.cuh"

namespace tb {

// r[0..11] = a[0..5] * b[0..5]. Products with i + j even accumulate in E (limb i + j), with i + j odd in O (limb
// i + j - 1): every product sits on an even index of its accumulator, i.e. on an aligned register pair.
// Row order: the rows that OPEN new top limbs (their last product lands on untouched limbs, addend RZ) run first,
// the rows that only add to existing limbs run last and ripple their carry to the top with plain ADDCs. That way no
// product ever needs a (carry limb, zero) addend pair, which ptxas would not fuse into one IMAD.WIDE.
TB_HD void mul6x6(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t e[12], o[11];
  // row 0: E <- a0 * (b0, b2, b4) at 0, 2, 4;  O <- a0 * (b1, b3, b5) at 0, 2, 4
#pragma unroll
  for (int j = 0; j < 6; j += 2) {
    mul_wide(e[j], e[j + 1], a[0], b[j]);
    mul_wide(o[j], o[j + 1], a[0], b[j + 1]);
  }
  // E, odd rows: a_i * (b1, b3, b5) at i+1, i+3, i+5; limbs i+5, i+6 are new
#pragma unroll
  for (int i = 1; i < 6; i += 2) {
    Carry c;
    e[i + 1] = mad_lo_cc(a[i], b[1], e[i + 1], c);
    e[i + 2] = madc_hi_cc(a[i], b[1], e[i + 2], c);
    e[i + 3] = madc_lo_cc(a[i], b[3], e[i + 3], c);
    e[i + 4] = madc_hi_cc(a[i], b[3], e[i + 4], c);
    e[i + 5] = madc_lo_cc(a[i], b[5], 0u, c);
    e[i + 6] = madc_hi(a[i], b[5], 0u, c);
  }
  // O, even rows: a_i * (b1, b3, b5) at i, i+2, i+4 (O index); limbs i+4, i+5 are new
#pragma unroll
  for (int i = 2; i < 6; i += 2) {
    Carry c;
    o[i] = mad_lo_cc(a[i], b[1], o[i], c);
    o[i + 1] = madc_hi_cc(a[i], b[1], o[i + 1], c);
    o[i + 2] = madc_lo_cc(a[i], b[3], o[i + 2], c);
    o[i + 3] = madc_hi_cc(a[i], b[3], o[i + 3], c);
    o[i + 4] = madc_lo_cc(a[i], b[5], 0u, c);
    o[i + 5] = madc_hi(a[i], b[5], 0u, c);
  }
  o[10] = 0;
  // E, even rows: a_i * (b0, b2, b4) at i, i+2, i+4, then ripple to limb 11
#pragma unroll
  for (int i = 2; i < 6; i += 2) {
    Carry c;
    e[i] = mad_lo_cc(a[i], b[0], e[i], c);
    e[i + 1] = madc_hi_cc(a[i], b[0], e[i + 1], c);
    e[i + 2] = madc_lo_cc(a[i], b[2], e[i + 2], c);
    e[i + 3] = madc_hi_cc(a[i], b[2], e[i + 3], c);
    e[i + 4] = madc_lo_cc(a[i], b[4], e[i + 4], c);
    e[i + 5] = madc_hi_cc(a[i], b[4], e[i + 5], c);
#pragma unroll
    for (int k = i + 6; k < 11; k++) e[k] = addc_cc(e[k], 0u, c);
    e[11] = addc(e[11], 0u, c);
  }
  // O, odd rows: a_i * (b0, b2, b4) at i-1, i+1, i+3 (O index), then ripple to limb 10
#pragma unroll
  for (int i = 1; i < 6; i += 2) {
    Carry c;
    o[i - 1] = mad_lo_cc(a[i], b[0], o[i - 1], c);
    o[i] = madc_hi_cc(a[i], b[0], o[i], c);
    o[i + 1] = madc_lo_cc(a[i], b[2], o[i + 1], c);
    o[i + 2] = madc_hi_cc(a[i], b[2], o[i + 2], c);
    o[i + 3] = madc_lo_cc(a[i], b[4], o[i + 3], c);
    o[i + 4] = madc_hi_cc(a[i], b[4], o[i + 4], c);
#pragma unroll
    for (int k = i + 5; k < 10; k++) o[k] = addc_cc(o[k], 0u, c);
    o[10] = addc(o[10], 0u, c);
  }
  // r = E + O * 2^32   (O occupies limbs 1..11: o[0..10])
  Carry c;
  r[0] = e[0];
  r[1] = add_cc(e[1], o[0], c);
#pragma unroll
  for (int k = 2; k < 11; k++) r[k] = addc_cc(e[k], o[k - 1], c);
  r[11] = addc(e[11], o[10], c);
}

// d = |x - y| over 6 limbs; returns the sign mask (all ones if x < y)
TB_HD uint32_t absdiff6(uint32_t* d, const uint32_t* x, const uint32_t* y) {
  uint32_t t[6];
  Carry c;
  t[0] = sub_cc(x[0], y[0], c);
#pragma unroll
  for (int i = 1; i < 6; i++) t[i] = subc_cc(x[i], y[i], c);
  const uint32_t m = subc_mask(c);
  // conditional negate: (t ^ m) - m
  Carry b;
  d[0] = sub_cc(t[0] ^ m, m, b);
#pragma unroll
  for (int i = 1; i < 5; i++) d[i] = subc_cc(t[i] ^ m, m, b);
  d[5] = subc_cc(t[5] ^ m, m, b);
  return m;
}

// T[0..23] = a * b for 12-limb a, b (any values)
TB_HD void kara_mul12(uint32_t* T, const uint32_t* a, const uint32_t* b) {
  uint32_t z0[12], z2[12], zm[12], da[6], db[6];
  mul6x6(z0, a, b);
  mul6x6(z2, a + 6, b + 6);
  const uint32_t sa = absdiff6(da, a, a + 6);      // a0 - a1
  const uint32_t sb = absdiff6(db, b + 6, b);      // b1 - b0
  mul6x6(zm, da, db);
  const uint32_t neg = sa ^ sb;                    // all ones: (a0 - a1)(b1 - b0) = -zm
  // mid = z0 + z2 +- zm  (13 limbs, non-negative)
  uint32_t mid[13];
  Carry c;
  mid[0] = add_cc(z0[0], z2[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) mid[i] = addc_cc(z0[i], z2[i], c);
  mid[12] = addc(0u, 0u, c);
  Carry d;
  (void)add_cc(neg, neg, d);                       // carry-in 1 iff neg (two's complement of zm)
#pragma unroll
  for (int i = 0; i < 12; i++) mid[i] = addc_cc(mid[i], zm[i] ^ neg, d);
  mid[12] = addc(mid[12], neg, d);
  // T = z0 + mid * 2^192 + z2 * 2^384
#pragma unroll
  for (int i = 0; i < 6; i++) T[i] = z0[i];
  Carry f;
  T[6] = add_cc(z0[6], mid[0], f);
#pragma unroll
  for (int i = 1; i < 6; i++) T[6 + i] = addc_cc(z0[6 + i], mid[i], f);
#pragma unroll
  for (int i = 0; i < 6; i++) T[12 + i] = addc_cc(z2[i], mid[6 + i], f);
  T[18] = addc_cc(z2[6], mid[12], f);
#pragma unroll
  for (int i = 7; i < 11; i++) T[12 + i] = addc_cc(z2[i], 0u, f);
  T[23] = addc(z2[11], 0u, f);
}

// r = T / 2^384 mod q for a 24-limb T (word-serial Montgomery reduction), lazily reduced: r < q + T / 2^384 and the
// caller guarantees that is below 2^384. The 12-limb window (x, E, O) is the one of mont_mul_lazy; row i cancels
// limb i with m = -limb, then the next high limb T[12 + i] is shifted in at the top.
template <class P>
TB_HD void mont_redc24(uint32_t* r, const uint32_t* T) {
  constexpr int N = P::N;
  static_assert(N == 12, "Fq only");
  uint32_t e[N + 1], o[N], x = 0;
#pragma unroll
  for (int j = 0; j < N; j++) {
    e[j] = T[j];
    o[j] = 0;
  }
  e[N] = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    Carry c;
    uint32_t s = add_cc(e[0], x, c);
    uint32_t k = addc(0, 0, c);
    uint32_t u = add_cc(s, 0xffffffffu, c);  // carries iff s != 0;  u = s - 1
    k = addc(k, 0, c);
    uint32_t m = ~u;                         // -s = ~(s - 1): LOP3, ALU pipe (a PTX neg becomes an FMA-pipe IMAD.MOV)
    e[1] = add_cc(e[1], k, c);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      e[j] = madc_lo_cc(m, P::p(j), e[j], c);
      e[j + 1] = madc_hi_cc(m, P::p(j), e[j + 1], c);
    }
    e[N] = addc(e[N], 0, c);
    o[0] = mad_lo_cc(m, P::p(1), o[0], c);
    o[1] = madc_hi_cc(m, P::p(1), o[1], c);
#pragma unroll
    for (int j = 2; j < N - 2; j += 2) {
      o[j] = madc_lo_cc(m, P::p(j + 1), o[j], c);
      o[j + 1] = madc_hi_cc(m, P::p(j + 1), o[j + 1], c);
    }
    o[N - 2] = madc_lo_cc(m, P::p(N - 1), o[N - 2], c);
    o[N - 1] = madc_hi(m, P::p(N - 1), o[N - 1], c);
    // window /= 2^32:  x' = e[1];  E' = O;  O' = E >> 64;  then T[12 + i] enters at the new limb 11 = o'[10]
    x = e[1];
    uint32_t t[N];
#pragma unroll
    for (int j = 0; j < N; j++) t[j] = o[j];
#pragma unroll
    for (int j = 0; j < N - 2; j++) o[j] = e[j + 2];
    o[N - 2] = add_cc(e[N], T[N + i], c);
    o[N - 1] = addc(0u, 0u, c);
#pragma unroll
    for (int j = 0; j < N; j++) e[j] = t[j];
    e[N] = 0;
  }
  Carry c;
  r[0] = add_cc(e[0], x, c);
#pragma unroll
  for (int j = 1; j < N; j++) r[j] = addc_cc(e[j], o[j - 1], c);
}

// r = a * b / 2^384 (mod q), lazily reduced (same bounds as mont_mul_lazy)
TB_HD void mont_mul_kara(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t T[24];
  kara_mul12(T, a, b);
  mont_redc24<FqParams>(r, T);
}

// r = (a * b + c * d) / 2^384 (mod q), lazily reduced: r < q + (a b + c d) / 2^384
TB_HD void mont_mul2_kara(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d) {
  uint32_t T[24], U[24];
  kara_mul12(T, a, b);
  kara_mul12(U, c, d);
  Carry k;
  T[0] = add_cc(T[0], U[0], k);
#pragma unroll
  for (int i = 1; i < 23; i++) T[i] = addc_cc(T[i], U[i], k);
  T[23] = addc(T[23], U[23], k);
  mont_redc24<FqParams>(r, T);
}

}  // namespace tb
