// BLS12-377 G2: the quadratic extension Fq2 = Fq[u] / (u^2 + 5) and the XYZZ group law on the twist
// E'(Fq2): y^2 = x^3 + B', B' = 1/u = (0, -1/5)  (D-type twist; ark-bls12-377 0.4 `g2::Config`, `Fq2Config::NONRESIDUE
// = -5`). Replaces ark-ec `short_weierstrass::{Affine, Projective}<ark_bls12_377::g2::Config>` arithmetic under the
// reference's G2 multi-scalar multiplications: `MultilinearPC::open` (src/sqrt_pst.rs:225), `commit_g2` and the G2
// `compress` of MIPP (src/mipp.rs:114,133) -- SURVEY.md 8f rank 1.
//
// Same formulas as g1.cuh (EFD shortw/xyzz, a = 0; the curve coefficient b never appears in them), written once over
// a field-operations class so the exceptional cases (inf + P, P + P, P + (-P)) are handled identically. Fq2 elements
// are canonical (both coordinates in [0, q)) between operations: the G2 MSMs of the reference are sqrt(n)-sized
// (<= 2^13 points), so the kernels favour simplicity over the lazy-reduction machinery of the G1 hot loop.
//
// Wire format (C ABI): G2 affine point = x.c0 || x.c1 || y.c0 || y.c1, 4 x 12 u32 (== ark's 4 x u64[6]), Montgomery
// form; all-zero == identity ((0,0) is not on the curve since B' != 0).
#pragma once
#include "g1.cuh"

namespace tb {

struct Fq2 {
  Fq c0, c1;
};

// The three (two) Fq products of an Fq2 product (square) are independent: they are inlined into ONE out-of-line
// function per Fq2 operation so that ptxas interleaves their carry chains. The G2 paths are latency-bound
// (sqrt(n)-sized MSMs, 253-step doubling chains), and a lone Montgomery product is a ~600-cycle dependent chain.
#if defined(__CUDA_ARCH__)
#define TB_G2_OL __device__ __noinline__
#else
#define TB_G2_OL inline
#endif

TB_HD void fq2_add(Fq2& r, const Fq2& a, const Fq2& b) {
  fq_add(r.c0, a.c0, b.c0);
  fq_add(r.c1, a.c1, b.c1);
}
TB_HD void fq2_sub(Fq2& r, const Fq2& a, const Fq2& b) {
  fq_sub(r.c0, a.c0, b.c0);
  fq_sub(r.c1, a.c1, b.c1);
}
TB_HD void fq2_dbl(Fq2& r, const Fq2& a) {
  fq_dbl(r.c0, a.c0);
  fq_dbl(r.c1, a.c1);
}
TB_HD void fq2_neg(Fq2& r, const Fq2& a) {
  fq_neg(r.c0, a.c0);
  fq_neg(r.c1, a.c1);
}
TB_HD bool fq2_is_zero(const Fq2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }
TB_HD Fq2 fq2_zero() {
  Fq2 r;
  r.c0 = fq_zero();
  r.c1 = fq_zero();
  return r;
}
TB_HD Fq2 fq2_one() {
  Fq2 r;
  r.c0 = fq_one();
  r.c1 = fq_zero();
  return r;
}
// r = 5 a (mod q)
TB_HD void fq_mul5(Fq& r, const Fq& a) {
  Fq t;
  fq_dbl(t, a);
  fq_dbl(t, t);
  fq_add(r, t, a);
}
// (a0 + a1 u)(b0 + b1 u) = (a0 b0 - 5 a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u     (Karatsuba: 3 Fq products)
TB_G2_OL void fq2_mul_ol(Fq2* rp, const Fq2* ap, const Fq2* bp) {
  const Fq2 a = *ap, b = *bp;
  Fq v0, v1, s, t, m;
  fq_add(s, a.c0, a.c1);
  fq_add(t, b.c0, b.c1);
  mont_mul_lazy<FqParams>(v0.l, a.c0.l, b.c0.l);   // three independent chains
  mont_mul_lazy<FqParams>(v1.l, a.c1.l, b.c1.l);
  mont_mul_lazy<FqParams>(m.l, s.l, t.l);
  mod_reduce_once<FqParams>(v0.l);
  mod_reduce_once<FqParams>(v1.l);
  mod_reduce_once<FqParams>(m.l);
  fq_sub(m, m, v0);
  Fq2 r;
  fq_sub(r.c1, m, v1);
  fq_mul5(t, v1);
  fq_sub(r.c0, v0, t);
  *rp = r;
}
TB_HD void fq2_mul(Fq2& r, const Fq2& a, const Fq2& b) { fq2_mul_ol(&r, &a, &b); }
// (a0 + a1 u)^2 = ((a0 + a1)(a0 - 5 a1) + 4 a0 a1) + 2 a0 a1 u                            (2 Fq products)
TB_G2_OL void fq2_sqr_ol(Fq2* rp, const Fq2* ap) {
  const Fq2 a = *ap;
  Fq v, s, t, m;
  fq_add(s, a.c0, a.c1);
  fq_mul5(t, a.c1);
  fq_sub(t, a.c0, t);
  mont_mul_lazy<FqParams>(v.l, a.c0.l, a.c1.l);    // two independent chains
  mont_mul_lazy<FqParams>(m.l, s.l, t.l);
  mod_reduce_once<FqParams>(v.l);
  mod_reduce_once<FqParams>(m.l);
  Fq2 r;
  fq_dbl(t, v);          // 2 a0 a1
  fq_dbl(s, t);          // 4 a0 a1
  fq_add(r.c0, m, s);
  r.c1 = t;
  *rp = r;
}
TB_HD void fq2_sqr(Fq2& r, const Fq2& a) { fq2_sqr_ol(&r, &a); }
// 1 / (a0 + a1 u) = (a0 - a1 u) / (a0^2 + 5 a1^2)
TB_HD void fq2_inv(Fq2& r, const Fq2& a) {
  Fq n, t, ni;
  fq_sqr(n, a.c0);
  fq_sqr(t, a.c1);
  fq_mul5(t, t);
  fq_add(n, n, t);
  fq_inv(ni, n);
  fq_mul(r.c0, a.c0, ni);
  fq_mul(t, a.c1, ni);
  fq_neg(r.c1, t);
}

struct Affine2 {
  Fq2 x, y;
};
struct Xyzz2 {  // x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity <=> ZZ == 0
  Fq2 x, y, zz, zzz;
};

TB_HD bool affine2_is_inf(const Affine2& p) { return fq2_is_zero(p.x) && fq2_is_zero(p.y); }
TB_HD bool xyzz2_is_inf(const Xyzz2& p) { return fq2_is_zero(p.zz); }
TB_HD void xyzz2_set_inf(Xyzz2& p) {
  p.x = fq2_zero();
  p.y = fq2_zero();
  p.zz = fq2_zero();
  p.zzz = fq2_zero();
}

// p = 2 p (EFD dbl-2008-s-1, a = 0)
TB_HD void xyzz2_dbl(Xyzz2& p) {
  if (xyzz2_is_inf(p)) return;
  Fq2 u, v, w, s, m, t;
  fq2_dbl(u, p.y);
  fq2_sqr(v, u);
  fq2_mul(w, u, v);
  fq2_mul(s, p.x, v);
  fq2_sqr(m, p.x);
  fq2_dbl(t, m);
  fq2_add(m, t, m);      // 3 x^2
  fq2_mul(t, w, p.y);    // W Y1 (old Y)
  fq2_sqr(p.x, m);
  fq2_sub(p.x, p.x, s);
  fq2_sub(p.x, p.x, s);
  fq2_sub(s, s, p.x);
  fq2_mul(s, m, s);
  fq2_sub(p.y, s, t);
  fq2_mul(p.zz, v, p.zz);
  fq2_mul(p.zzz, w, p.zzz);
  // y == 0 (a point of order 2 outside the prime-order subgroup) gives v = w = 0: the identity, correctly
}

// p += q, q affine (EFD madd-2008-s): 8M + 2S over Fq2
TB_HD void xyzz2_madd(Xyzz2& p, const Affine2& q) {
  if (affine2_is_inf(q)) return;
  if (xyzz2_is_inf(p)) {
    p.x = q.x;
    p.y = q.y;
    p.zz = fq2_one();
    p.zzz = fq2_one();
    return;
  }
  Fq2 pp, rr, t, ppp, qq;
  fq2_mul(pp, q.x, p.zz);   // U2
  fq2_mul(rr, q.y, p.zzz);  // S2
  fq2_sub(pp, pp, p.x);     // P
  fq2_sub(rr, rr, p.y);     // R
  if (fq2_is_zero(pp)) {
    if (fq2_is_zero(rr)) {  // P + P
      p.x = q.x;
      p.y = q.y;
      p.zz = fq2_one();
      p.zzz = fq2_one();
      xyzz2_dbl(p);
    } else {
      xyzz2_set_inf(p);
    }
    return;
  }
  fq2_sqr(t, pp);           // PP
  fq2_mul(ppp, pp, t);      // PPP
  fq2_mul(qq, p.x, t);      // Q
  fq2_mul(p.zz, p.zz, t);
  fq2_mul(p.zzz, p.zzz, ppp);
  fq2_sqr(t, rr);
  fq2_sub(t, t, ppp);
  fq2_sub(t, t, qq);
  fq2_sub(p.x, t, qq);      // X3 = R^2 - PPP - 2Q
  fq2_sub(qq, qq, p.x);
  fq2_mul(qq, rr, qq);
  fq2_mul(t, p.y, ppp);
  fq2_sub(p.y, qq, t);      // Y3 = R (Q - X3) - Y1 PPP
}

// p += q, both XYZZ (EFD add-2008-s): 12M + 2S over Fq2
TB_HD void xyzz2_add(Xyzz2& p, const Xyzz2& q) {
  if (xyzz2_is_inf(q)) return;
  if (xyzz2_is_inf(p)) {
    p = q;
    return;
  }
  Fq2 u1, u2, s1, s2, pp, ppp, t;
  fq2_mul(u1, p.x, q.zz);
  fq2_mul(u2, q.x, p.zz);
  fq2_mul(s1, p.y, q.zzz);
  fq2_mul(s2, q.y, p.zzz);
  fq2_sub(u2, u2, u1);  // P
  fq2_sub(s2, s2, s1);  // R
  if (fq2_is_zero(u2)) {
    if (fq2_is_zero(s2)) xyzz2_dbl(p);
    else xyzz2_set_inf(p);
    return;
  }
  fq2_sqr(pp, u2);
  fq2_mul(ppp, u2, pp);
  fq2_mul(u1, u1, pp);  // Q
  fq2_mul(t, p.zz, q.zz);
  fq2_mul(p.zz, t, pp);
  fq2_mul(t, p.zzz, q.zzz);
  fq2_mul(p.zzz, t, ppp);
  fq2_sqr(t, s2);
  fq2_sub(t, t, ppp);
  fq2_sub(t, t, u1);
  fq2_sub(p.x, t, u1);
  fq2_sub(u1, u1, p.x);
  fq2_mul(u1, s2, u1);
  fq2_mul(t, s1, ppp);
  fq2_sub(p.y, u1, t);
}

// canonical affine output (one Fq inversion); identity -> all-zero
TB_HD void xyzz2_to_affine(Affine2& r, const Xyzz2& p) {
  if (xyzz2_is_inf(p)) {
    r.x = fq2_zero();
    r.y = fq2_zero();
    return;
  }
  Fq2 t, ti, a, b;
  fq2_mul(t, p.zz, p.zzz);
  fq2_inv(ti, t);
  fq2_mul(a, ti, p.zzz);  // 1/ZZ
  fq2_mul(b, ti, p.zz);   // 1/ZZZ
  fq2_mul(r.x, p.x, a);
  fq2_mul(r.y, p.y, b);
}

// k * p by left-to-right double-and-add over a canonical 8-limb scalar (MIPP `compress` on G2, src/mipp.rs:133)
TB_HD void xyzz2_scalar_mul(Xyzz2& r, const Affine2& p, const uint32_t k[8]) {
  xyzz2_set_inf(r);
  bool started = false;
  for (int i = 7; i >= 0; i--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) xyzz2_dbl(r);
      if ((k[i] >> bit) & 1) {
        xyzz2_madd(r, p);
        started = true;
      }
    }
  }
}

}  // namespace tb
