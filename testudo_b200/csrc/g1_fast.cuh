// Hot-loop group law: XYZZ mixed addition with LAZY field reduction and an out-of-line multiplier.
//
// Why (ncu, profiles/r01_accumulate_v1.md): with every Montgomery multiplication inlined the bucket loop was
// ~180 KB of SASS and the top stall was `no_instruction` (I-cache misses, 3.1 cycles per issue); ~60% of the
// executed instructions were not multiply-accumulates (conditional subtractions, selects, carry fix-ups), so the
// ALU pipe, not the FMA pipe, set the pace. Here
//   * fq_mul_call is ONE out-of-line copy of the multiplier; operands and result travel in registers
//     (by-value structs; ptxas keeps them out of local memory), so the loop body is ~15 KB;
//   * nothing in the loop is reduced to [0, q): Fq has 7 spare bits in 384 (128 q < R), so values may grow to
//     ~10 q. A product of a < A q and b < B q leaves the multiplier below (1 + A*B/128) q; a subtraction is
//     a + k q - b with a constant k q >= the bound of b. The bounds are tracked line by line below.
//   * P = 0 (mod q) -- the P+P / P+(-P) cases -- is detected by a one-instruction filter: q = 1 (mod 2^32), so
//     P = k q (1 <= k <= 9) forces the low limb of P to be k. Only then is the exact (canonical) path taken.
// Invariant of the accumulator between additions:  X < 8q, Y < 4q, ZZ < 2q, ZZZ < 2q, identity <=> ZZ == 0
// (ZZ is a product of factors that are non-zero mod q on this path, so it is never a non-zero multiple of q).
#pragma once
#include "g1.cuh"

namespace tb {

// k * q for k = 2, 4, 8 (subtraction offsets), little-endian limbs
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t FQ_KQ_DEV[3][12] = {
    {0x00000002u, 0x0a118000u, 0x60000001u, 0x2e16ba88u, 0x74129000u, 0x3de6c45fu, 0x01ea271eu, 0x3445b3e6u,
     0xd9429276u, 0x8c760b80u, 0x2f8a21d5u, 0x035c748cu},
    {0x00000004u, 0x14230000u, 0xc0000002u, 0x5c2d7510u, 0xe8252000u, 0x7bcd88beu, 0x03d44e3cu, 0x688b67ccu,
     0xb28524ecu, 0x18ec1701u, 0x5f1443abu, 0x06b8e918u},
    {0x00000008u, 0x28460000u, 0x80000004u, 0xb85aea21u, 0xd04a4000u, 0xf79b117du, 0x07a89c78u, 0xd116cf98u,
     0x650a49d8u, 0x31d82e03u, 0xbe288756u, 0x0d71d230u}};
#endif
TB_HD uint32_t fq_kq(int sel, int i) {  // sel: 0 -> 2q, 1 -> 4q, 2 -> 8q
#ifdef __CUDA_ARCH__
  return FQ_KQ_DEV[sel][i];
#else
  constexpr uint32_t V[3][12] = {
      {0x00000002u, 0x0a118000u, 0x60000001u, 0x2e16ba88u, 0x74129000u, 0x3de6c45fu, 0x01ea271eu, 0x3445b3e6u,
       0xd9429276u, 0x8c760b80u, 0x2f8a21d5u, 0x035c748cu},
      {0x00000004u, 0x14230000u, 0xc0000002u, 0x5c2d7510u, 0xe8252000u, 0x7bcd88beu, 0x03d44e3cu, 0x688b67ccu,
       0xb28524ecu, 0x18ec1701u, 0x5f1443abu, 0x06b8e918u},
      {0x00000008u, 0x28460000u, 0x80000004u, 0xb85aea21u, 0xd04a4000u, 0xf79b117du, 0x07a89c78u, 0xd116cf98u,
       0x650a49d8u, 0x31d82e03u, 0xbe288756u, 0x0d71d230u}};
  return V[sel][i];
#endif
}

// r = a + k*q - b  (k = 2 << SEL); exact in 384 bits when b <= k*q and a + k*q < 2^384
template <int SEL>
TB_HD void fq_sub_lazy(Fq& r, const Fq& a, const Fq& b) {
  Carry c;
  uint32_t t[12];
  t[0] = sub_cc(a.l[0], b.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) t[i] = subc_cc(a.l[i], b.l[i], c);
  Carry d;
  r.l[0] = add_cc(t[0], fq_kq(SEL, 0), d);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = addc_cc(t[i], fq_kq(SEL, i), d);
}

// value < 16 q  ->  canonical [0, q): conditional subtraction of 8q, 4q, 2q, q
TB_HD void fq_canon(Fq& a) {
#pragma unroll
  for (int sel = 2; sel >= -1; sel--) {
    uint32_t t[12];
    Carry c;
    t[0] = sub_cc(a.l[0], sel >= 0 ? fq_kq(sel, 0) : FqParams::p(0), c);
#pragma unroll
    for (int i = 1; i < 12; i++) t[i] = subc_cc(a.l[i], sel >= 0 ? fq_kq(sel, i) : FqParams::p(i), c);
    uint32_t borrow = subc_mask(c);
#pragma unroll
    for (int i = 0; i < 12; i++) a.l[i] = borrow ? a.l[i] : t[i];
  }
}

// the one out-of-line multiplier of the hot loop (register-passed on the device)
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ Fq fq_mul_call(Fq a, Fq b) {
  Fq r;
  mont_mul_lazy<FqParams>(r.l, a.l, b.l);
  return r;
}
#else
inline Fq fq_mul_call(Fq a, Fq b) {
  Fq r;
  mont_mul_lazy<FqParams>(r.l, a.l, b.l);
  return r;
}
#endif

// dedicated squaring (78 + 132 wide MACs instead of 144 + 132), same calling convention
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ Fq fq_sqr_call(Fq a) {
  Fq r;
  mont_sqr_lazy<FqParams>(r.l, a.l);
  return r;
}
#else
inline Fq fq_sqr_call(Fq a) {
  Fq r;
  mont_sqr_lazy<FqParams>(r.l, a.l);
  return r;
}
#endif

// exact path for the exceptional cases; operands by value so the caller's registers never get an address
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
    Xyzz
    xyzz_madd_exact(Xyzz p, Affine q) {
  fq_canon(p.x);
  fq_canon(p.y);
  fq_canon(p.zz);
  fq_canon(p.zzz);
  xyzz_madd(p, q);
  return p;
}

// acc += q (q affine with canonical coordinates, y possibly == q after negation of 0 -- harmless)
TB_HD void xyzz_madd_fast(Xyzz& p, const Affine& q) {
  if (affine_is_inf(q)) return;
  if (xyzz_is_inf(p)) {
    p.x = q.x;
    p.y = q.y;
    p.zz = fq_one();
    p.zzz = fq_one();
    return;
  }
  Fq pp, rr;
  pp = fq_mul_call(q.x, p.zz);     // U2  < 2q
  rr = fq_mul_call(q.y, p.zzz);    // S2  < 2q
  fq_sub_lazy<2>(pp, pp, p.x);     // P = U2 + 8q - X1   in (0, 10q)
  fq_sub_lazy<1>(rr, rr, p.y);     // R = S2 + 4q - Y1   in (0, 6q)
  if (pp.l[0] - 1u < 9u) {         // P = k q possible (q = 1 mod 2^32): decide exactly, out of line
    Fq chk = pp;
    fq_canon(chk);
    if (fq_is_zero(chk)) {
      p = xyzz_madd_exact(p, q);
      return;
    }
  }
  Fq t, ppp, qq;
  t = fq_sqr_call(pp);             // PP  < 1.8q
  ppp = fq_mul_call(pp, t);        // PPP < 1.2q
  qq = fq_mul_call(p.x, t);        // Q   < 1.2q
  p.zz = fq_mul_call(p.zz, t);     // ZZ3  < 2q
  p.zzz = fq_mul_call(p.zzz, ppp); // ZZZ3 < 2q
  t = fq_sqr_call(rr);             // RR  < 1.3q
  fq_sub_lazy<0>(t, t, ppp);       // RR + 2q - PPP
  fq_sub_lazy<0>(t, t, qq);        //    + 2q - Q
  fq_sub_lazy<0>(p.x, t, qq);      // X3 = ... + 2q - Q  < 7.3q  (invariant X < 8q)
  fq_sub_lazy<2>(qq, qq, p.x);     // Q + 8q - X3        < 9.2q
  qq = fq_mul_call(rr, qq);        // R (Q - X3)         < 1.5q
  t = fq_mul_call(p.y, ppp);       // Y1 PPP             < 1.1q
  fq_sub_lazy<0>(p.y, qq, t);      // Y3 = .. + 2q - ..  < 3.5q  (invariant Y < 4q)
}

// exact XYZZ + XYZZ for the exceptional cases of the lazy add (by value: no addresses of caller registers)
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
    Xyzz
    xyzz_add_exact(Xyzz p, Xyzz q) {
  fq_canon(p.x); fq_canon(p.y); fq_canon(p.zz); fq_canon(p.zzz);
  fq_canon(q.x); fq_canon(q.y); fq_canon(q.zz); fq_canon(q.zzz);
  xyzz_add(p, q);
  return p;
}

// p += q, both XYZZ with the lazy invariant (X < 8q, Y < 4q, ZZ, ZZZ < 2q); result satisfies it again.
// 12M + 2S through the out-of-line multiplier. Used by the bucket-reduction and fix-up kernels.
TB_HD void xyzz_add_fast(Xyzz& p, const Xyzz& q) {
  if (xyzz_is_inf(q)) return;
  if (xyzz_is_inf(p)) {
    p = q;
    return;
  }
  Fq u1, u2, s1, s2;
  u1 = fq_mul_call(p.x, q.zz);     // U1 < 1.2q
  u2 = fq_mul_call(q.x, p.zz);     // U2 < 1.2q
  s1 = fq_mul_call(p.y, q.zzz);    // S1 < 1.1q
  s2 = fq_mul_call(q.y, p.zzz);    // S2 < 1.1q
  fq_sub_lazy<0>(u2, u2, u1);      // P = U2 + 2q - U1 in (0, 3.2q)
  fq_sub_lazy<0>(s2, s2, s1);      // R = S2 + 2q - S1 in (0, 3.1q)
  if (u2.l[0] - 1u < 3u) {         // P = k q (k <= 3) possible: decide exactly
    Fq chk = u2;
    fq_canon(chk);
    if (fq_is_zero(chk)) {
      p = xyzz_add_exact(p, q);
      return;
    }
  }
  Fq pp, ppp, t;
  pp = fq_sqr_call(u2);            // PP  < 1.1q
  ppp = fq_mul_call(u2, pp);       // PPP < 1.1q
  u1 = fq_mul_call(u1, pp);        // Q   < 1.1q
  t = fq_mul_call(p.zz, q.zz);
  p.zz = fq_mul_call(t, pp);       // ZZ3  < 2q
  t = fq_mul_call(p.zzz, q.zzz);
  p.zzz = fq_mul_call(t, ppp);     // ZZZ3 < 2q
  t = fq_sqr_call(s2);             // RR < 1.1q
  fq_sub_lazy<0>(t, t, ppp);
  fq_sub_lazy<0>(t, t, u1);
  fq_sub_lazy<0>(p.x, t, u1);      // X3 < 7.1q
  fq_sub_lazy<2>(u1, u1, p.x);     // Q + 8q - X3 < 9.1q
  u1 = fq_mul_call(s2, u1);        // < 1.3q
  t = fq_mul_call(s1, ppp);        // < 1.1q
  fq_sub_lazy<0>(p.y, u1, t);      // Y3 < 3.3q
}

// p = 2p with the lazy invariant in and out (6M + 3S)
TB_HD void xyzz_dbl_fast(Xyzz& p) {
  if (xyzz_is_inf(p)) return;
  Fq u, v, w, s, m, t;
  Carry c;
  // U = 2Y < 8q (plain limb add: no reduction needed below 2^384)
  u.l[0] = add_cc(p.y.l[0], p.y.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) u.l[i] = addc_cc(p.y.l[i], p.y.l[i], c);
  v = fq_sqr_call(u);              // V < 1.5q
  w = fq_mul_call(u, v);           // W < 1.1q
  s = fq_mul_call(p.x, v);         // S < 1.1q
  m = fq_sqr_call(p.x);            // X^2 < 1.5q
  Carry d;
  t.l[0] = add_cc(m.l[0], m.l[0], d);
#pragma unroll
  for (int i = 1; i < 12; i++) t.l[i] = addc_cc(m.l[i], m.l[i], d);
  Carry e;
  m.l[0] = add_cc(t.l[0], m.l[0], e);
#pragma unroll
  for (int i = 1; i < 12; i++) m.l[i] = addc_cc(t.l[i], m.l[i], e);   // M = 3 X^2 < 4.5q
  t = fq_mul_call(w, p.y);         // W * Y1 < 1.1q
  Fq mm = fq_sqr_call(m);          // < 1.2q
  fq_sub_lazy<0>(mm, mm, s);
  fq_sub_lazy<0>(p.x, mm, s);      // X3 = M^2 + 4q - 2S < 5.2q
  fq_sub_lazy<2>(s, s, p.x);       // S + 8q - X3 < 9.1q
  s = fq_mul_call(m, s);           // < 1.4q
  fq_sub_lazy<0>(p.y, s, t);       // Y3 < 3.4q
  p.zz = fq_mul_call(v, p.zz);     // < 2q
  p.zzz = fq_mul_call(w, p.zzz);   // < 2q
}

TB_HD void xyzz_canon(Xyzz& p) {
  fq_canon(p.x);
  fq_canon(p.y);
  fq_canon(p.zz);
  fq_canon(p.zzz);
}

}  // namespace tb
