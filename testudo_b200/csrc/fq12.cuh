// BLS12-377 pairing arithmetic: the tower Fq2 -> Fq6 -> Fq12, the optimal-ate Miller loop and the final
// exponentiation. Replaces ark-ec 0.4 `models::bls12::{Bls12::multi_miller_loop, final_exponentiation}` and
// ark-ff `Fp6_3over2 / Fp12_2over3over2` under the reference's pairing products on the commitment path:
//     t = E::multi_pairing(comm_list, h_vec)              src/sqrt_pst.rs:131-144
//     comm_t_l / comm_t_r = pairings_product(a, h)        src/mipp.rs:87-94,396-398
// (SURVEY.md 8f rank 3).
//
//     Fq2  = Fq[u]  / (u^2 + 5)         (g2.cuh)
//     Fq6  = Fq2[v] / (v^3 - u)
//     Fq12 = Fq6[w] / (w^2 - v)         ark-bls12-377 0.4 Fq6Config::NONRESIDUE = u, Fq12Config::NONRESIDUE = v
// In-memory order == ark's: c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 (each an Fq2 = c0 || c1 of 12 u32 Montgomery limbs),
// i.e. the coefficients of w^0, w^2, w^4, w^1, w^3, w^5.
//
// Only the VALUE in GT is part of the contract (one canonical Fq12 element per input); intermediate Miller values may
// differ from ark's by subfield factors that the final exponentiation removes. The final exponentiation follows
// ark's choice of exponent exactly: easy part (q^6 - 1)(q^2 + 1), hard part (x-1)^2 (x+q)(x^2+q^2-1) + 3
// = 3 (q^4 - q^2 + 1)/r (eprint 2020/875), so the result is the cube of the textbook reduced pairing, as in ark.
//
// Everything is __host__ __device__ so tests/host_check can run the same code on a CPU-only box against
// oracle/pairing.py; the product only ever runs it on the device (kernels_pairing.cuh).
#pragma once
#include "g2.cuh"

namespace tb {

// constant tables: a host copy (tests/host_check) and, under nvcc, a __constant__ copy; FQ12_C(x) picks per pass
#define TB_FQ12_CONST static const
#define TB_FQ12_NAME(x) FQ12_##x##_H
#include "fq12_consts.inc"
#undef TB_FQ12_CONST
#undef TB_FQ12_NAME
#if defined(__CUDACC__)
#define TB_FQ12_CONST static __device__ __constant__
#define TB_FQ12_NAME(x) FQ12_##x##_D
#include "fq12_consts.inc"
#undef TB_FQ12_CONST
#undef TB_FQ12_NAME
#endif
#ifdef __CUDA_ARCH__
#define FQ12_C(x) FQ12_##x##_D
#else
#define FQ12_C(x) FQ12_##x##_H
#endif

struct Fq6 {
  Fq2 c0, c1, c2;
};
struct Fq12 {
  Fq6 c0, c1;
};

TB_HD Fq fq_from_table(const uint32_t* t) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = t[i];
  return r;
}
// ONE out-of-line copy of the 700-instruction Montgomery product for the pairing code: with the products inlined the
// Miller loop's working set (doubling / addition steps, scalings by P's coordinates, twist coefficient) ran to well
// over 100 KB of SASS and a lone warp stalled on instruction fetch (ncu: `no_instruction` 1.24 per issue)
TB_G2_OL void fq_mul_ol(Fq* r, const Fq* a, const Fq* b) { fq_mul(*r, *a, *b); }
// a * (b in Fq)
TB_HD void fq2_scale(Fq2& r, const Fq2& a, const Fq& k) {
  fq_mul_ol(&r.c0, &a.c0, &k);
  fq_mul_ol(&r.c1, &a.c1, &k);
}
// u * (a0 + a1 u) = -5 a1 + a0 u
TB_HD void fq2_mul_xi(Fq2& r, const Fq2& a) {
  Fq t;
  fq_mul5(t, a.c1);
  Fq a0 = a.c0;
  fq_neg(r.c0, t);
  r.c1 = a0;
}
TB_HD void fq2_conj(Fq2& r, const Fq2& a) {
  r.c0 = a.c0;
  fq_neg(r.c1, a.c1);
}

// ---- Fq6 -------------------------------------------------------------------------------------------------------
TB_HD void fq6_add(Fq6& r, const Fq6& a, const Fq6& b) {
  fq2_add(r.c0, a.c0, b.c0);
  fq2_add(r.c1, a.c1, b.c1);
  fq2_add(r.c2, a.c2, b.c2);
}
TB_HD void fq6_sub(Fq6& r, const Fq6& a, const Fq6& b) {
  fq2_sub(r.c0, a.c0, b.c0);
  fq2_sub(r.c1, a.c1, b.c1);
  fq2_sub(r.c2, a.c2, b.c2);
}
TB_HD void fq6_neg(Fq6& r, const Fq6& a) {
  fq2_neg(r.c0, a.c0);
  fq2_neg(r.c1, a.c1);
  fq2_neg(r.c2, a.c2);
}
// v * (c0 + c1 v + c2 v^2) = u c2 + c0 v + c1 v^2
TB_HD void fq6_mul_v(Fq6& r, const Fq6& a) {
  Fq2 t;
  fq2_mul_xi(t, a.c2);
  r.c2 = a.c1;
  r.c1 = a.c0;
  r.c0 = t;
}
// Karatsuba over Fq2: 6 products
TB_G2_OL void fq6_mul_ol(Fq6* rp, const Fq6* ap, const Fq6* bp) {
  const Fq6& a = *ap;
  const Fq6& b = *bp;
  Fq2 v0, v1, v2, s, t, m;
  Fq6 r;
  fq2_mul(v0, a.c0, b.c0);
  fq2_mul(v1, a.c1, b.c1);
  fq2_mul(v2, a.c2, b.c2);
  fq2_add(s, a.c1, a.c2);
  fq2_add(t, b.c1, b.c2);
  fq2_mul(m, s, t);
  fq2_sub(m, m, v1);
  fq2_sub(m, m, v2);
  fq2_mul_xi(m, m);
  fq2_add(r.c0, v0, m);
  fq2_add(s, a.c0, a.c1);
  fq2_add(t, b.c0, b.c1);
  fq2_mul(m, s, t);
  fq2_sub(m, m, v0);
  fq2_sub(m, m, v1);
  fq2_mul_xi(s, v2);
  fq2_add(r.c1, m, s);
  fq2_add(s, a.c0, a.c2);
  fq2_add(t, b.c0, b.c2);
  fq2_mul(m, s, t);
  fq2_sub(m, m, v0);
  fq2_sub(m, m, v2);
  fq2_add(r.c2, m, v1);
  *rp = r;
}
TB_HD void fq6_mul(Fq6& r, const Fq6& a, const Fq6& b) { fq6_mul_ol(&r, &a, &b); }
// a * (d0 + d1 v): 5 products
TB_G2_OL void fq6_mul_by_01_ol(Fq6* rp, const Fq6* ap, const Fq2* d0p, const Fq2* d1p) {
  const Fq6& a = *ap;
  const Fq2 d0 = *d0p, d1 = *d1p;
  Fq2 aa, bb, s, t, m;
  Fq6 r;
  fq2_mul(aa, a.c0, d0);
  fq2_mul(bb, a.c1, d1);
  fq2_add(s, a.c1, a.c2);
  fq2_mul(m, s, d1);
  fq2_sub(m, m, bb);
  fq2_mul_xi(m, m);
  fq2_add(r.c0, m, aa);
  fq2_add(s, a.c0, a.c1);
  fq2_add(t, d0, d1);
  fq2_mul(m, s, t);
  fq2_sub(m, m, aa);
  fq2_sub(r.c1, m, bb);
  fq2_add(s, a.c0, a.c2);
  fq2_mul(m, s, d0);
  fq2_sub(m, m, aa);
  fq2_add(r.c2, m, bb);
  *rp = r;
}
TB_HD void fq6_inv(Fq6& r, const Fq6& a) {
  Fq2 t0, t1, t2, s, d;
  fq2_sqr(t0, a.c0);
  fq2_mul(s, a.c1, a.c2);
  fq2_mul_xi(s, s);
  fq2_sub(t0, t0, s);   // a0^2 - xi a1 a2
  fq2_sqr(t1, a.c2);
  fq2_mul_xi(t1, t1);
  fq2_mul(s, a.c0, a.c1);
  fq2_sub(t1, t1, s);   // xi a2^2 - a0 a1
  fq2_sqr(t2, a.c1);
  fq2_mul(s, a.c0, a.c2);
  fq2_sub(t2, t2, s);   // a1^2 - a0 a2
  fq2_mul(d, a.c2, t1);
  fq2_mul(s, a.c1, t2);
  fq2_add(d, d, s);
  fq2_mul_xi(d, d);
  fq2_mul(s, a.c0, t0);
  fq2_add(d, d, s);     // norm to Fq2
  fq2_inv(d, d);
  fq2_mul(r.c0, t0, d);
  fq2_mul(r.c1, t1, d);
  fq2_mul(r.c2, t2, d);
}

// ---- Fq12 ------------------------------------------------------------------------------------------------------
TB_HD Fq12 fq12_one() {
  Fq12 r;
  r.c0.c0 = fq2_one();
  r.c0.c1 = fq2_zero();
  r.c0.c2 = fq2_zero();
  r.c1.c0 = fq2_zero();
  r.c1.c1 = fq2_zero();
  r.c1.c2 = fq2_zero();
  return r;
}
TB_G2_OL void fq12_mul_ol(Fq12* rp, const Fq12* ap, const Fq12* bp) {
  const Fq12& a = *ap;
  const Fq12& b = *bp;
  Fq6 v0, v1, s, t;
  fq6_mul(v0, a.c0, b.c0);
  fq6_mul(v1, a.c1, b.c1);
  fq6_add(s, a.c0, a.c1);
  fq6_add(t, b.c0, b.c1);
  fq6_mul(s, s, t);
  fq6_sub(s, s, v0);
  fq6_sub(rp->c1, s, v1);
  fq6_mul_v(v1, v1);
  fq6_add(rp->c0, v0, v1);
}
TB_HD void fq12_mul(Fq12& r, const Fq12& a, const Fq12& b) { fq12_mul_ol(&r, &a, &b); }
// complex squaring: 2 Fq6 products
TB_G2_OL void fq12_sqr_ol(Fq12* rp, const Fq12* ap) {
  const Fq12& a = *ap;
  Fq6 ab, s, t;
  fq6_mul(ab, a.c0, a.c1);
  fq6_add(s, a.c0, a.c1);
  fq6_mul_v(t, a.c1);
  fq6_add(t, t, a.c0);
  fq6_mul(s, s, t);        // (a0 + a1)(a0 + v a1) = a0^2 + v a1^2 + ab + v ab
  fq6_sub(s, s, ab);
  fq6_mul_v(t, ab);
  fq6_sub(rp->c0, s, t);
  fq6_add(rp->c1, ab, ab);
}
TB_HD void fq12_sqr(Fq12& r, const Fq12& a) { fq12_sqr_ol(&r, &a); }
// a^(q^6): w -> -w. Inverse of a unitary element (anything after the easy part of the final exponentiation).
TB_HD void fq12_conj(Fq12& r, const Fq12& a) {
  r.c0 = a.c0;
  fq6_neg(r.c1, a.c1);
}
TB_G2_OL void fq12_inv_ol(Fq12* rp, const Fq12* ap) {
  Fq12& r = *rp;
  const Fq12& a = *ap;
  Fq6 t, s;
  fq6_mul(t, a.c0, a.c0);
  fq6_mul(s, a.c1, a.c1);
  fq6_mul_v(s, s);
  fq6_sub(t, t, s);        // a0^2 - v a1^2
  fq6_inv(t, t);
  fq6_mul(r.c0, a.c0, t);
  fq6_mul(s, a.c1, t);
  fq6_neg(r.c1, s);
}
TB_HD void fq12_inv(Fq12& r, const Fq12& a) { fq12_inv_ol(&r, &a); }
// coefficient of w^i (i = 0..5) inside the tower layout
TB_HD Fq2& fq12_coeff(Fq12& a, int i) {
  Fq6& h = (i & 1) ? a.c1 : a.c0;
  int j = i >> 1;
  return j == 0 ? h.c0 : (j == 1 ? h.c1 : h.c2);
}
// a^(q^k), k = 1 or 2: coefficient of w^i -> conj^k(.) * u^(i (q^k - 1)/6)
TB_G2_OL void fq12_frobenius_ol(Fq12* rp, const Fq12* ap, int k) {
  Fq12 r = *ap;
  for (int i = 0; i < 6; i++) {
    Fq2& c = fq12_coeff(r, i);
    if (k == 1) fq2_conj(c, c);
    if (i == 0) continue;
    if (k == 1) {
      Fq2 g;
      g.c0 = fq_from_table(FQ12_C(FROB1)[i - 1][0]);
      g.c1 = fq_from_table(FQ12_C(FROB1)[i - 1][1]);
      fq2_mul(c, c, g);
    } else {
      Fq g = fq_from_table(FQ12_C(FROB2)[i - 1]);
      fq2_scale(c, c, g);
    }
  }
  *rp = r;
}
TB_HD void fq12_frobenius(Fq12& r, const Fq12& a, int k) { fq12_frobenius_ol(&r, &a, k); }
// a * ((l0, 0, 0) + (l3, l4, 0) w): the line of a D-type twist (ark `mul_by_034`), 13 Fq2 products
TB_G2_OL void fq12_mul_by_034_ol(Fq12* ap, const Fq2* l0p, const Fq2* l3p, const Fq2* l4p) {
  Fq12& a = *ap;
  const Fq2 l0 = *l0p, l3 = *l3p, l4 = *l4p;
  Fq6 x, y, e;
  fq2_mul(x.c0, a.c0.c0, l0);
  fq2_mul(x.c1, a.c0.c1, l0);
  fq2_mul(x.c2, a.c0.c2, l0);             // x = a0 * l0
  fq6_mul_by_01_ol(&y, &a.c1, &l3, &l4);  // y = a1 * (l3 + l4 v)
  Fq2 s;
  fq2_add(s, l0, l3);
  fq6_add(e, a.c0, a.c1);
  fq6_mul_by_01_ol(&e, &e, &s, &l4);      // (a0 + a1)(l0 + l3 + l4 v)
  fq6_sub(e, e, x);
  fq6_sub(a.c1, e, y);
  fq6_mul_v(y, y);
  fq6_add(a.c0, x, y);
}
// Granger-Scott squaring of a unitary element (ark `cyclotomic_square_in_place`): 9 Fq2 products instead of 12
TB_G2_OL void fq12_cyclotomic_sqr_ol(Fq12* rp, const Fq12* ap) {
  const Fq2 z0 = ap->c0.c0, z4 = ap->c0.c1, z3 = ap->c0.c2, z2 = ap->c1.c0, z1 = ap->c1.c1, z5 = ap->c1.c2;
  Fq2 t0, t1, t2, t3, t4, t5, tmp, s, m;
  // (z0 + z1 y)^2, y^2 = xi
  fq2_mul(tmp, z0, z1);
  fq2_add(s, z0, z1);
  fq2_mul_xi(m, z1);
  fq2_add(m, m, z0);
  fq2_mul(t0, s, m);
  fq2_sub(t0, t0, tmp);
  fq2_mul_xi(m, tmp);
  fq2_sub(t0, t0, m);
  fq2_dbl(t1, tmp);
  // (z2 + z3 y)^2
  fq2_mul(tmp, z2, z3);
  fq2_add(s, z2, z3);
  fq2_mul_xi(m, z3);
  fq2_add(m, m, z2);
  fq2_mul(t2, s, m);
  fq2_sub(t2, t2, tmp);
  fq2_mul_xi(m, tmp);
  fq2_sub(t2, t2, m);
  fq2_dbl(t3, tmp);
  // (z4 + z5 y)^2
  fq2_mul(tmp, z4, z5);
  fq2_add(s, z4, z5);
  fq2_mul_xi(m, z5);
  fq2_add(m, m, z4);
  fq2_mul(t4, s, m);
  fq2_sub(t4, t4, tmp);
  fq2_mul_xi(m, tmp);
  fq2_sub(t4, t4, m);
  fq2_dbl(t5, tmp);
  Fq2 o;
  // z0' = 3 t0 - 2 z0
  fq2_sub(o, t0, z0);
  fq2_dbl(o, o);
  fq2_add(rp->c0.c0, o, t0);
  // z1' = 3 t1 + 2 z1
  fq2_add(o, t1, z1);
  fq2_dbl(o, o);
  fq2_add(rp->c1.c1, o, t1);
  // z2' = 3 xi t5 + 2 z2
  fq2_mul_xi(tmp, t5);
  fq2_add(o, tmp, z2);
  fq2_dbl(o, o);
  fq2_add(rp->c1.c0, o, tmp);
  // z3' = 3 t4 - 2 z3
  fq2_sub(o, t4, z3);
  fq2_dbl(o, o);
  fq2_add(rp->c0.c2, o, t4);
  // z4' = 3 t2 - 2 z4
  fq2_sub(o, t2, z4);
  fq2_dbl(o, o);
  fq2_add(rp->c0.c1, o, t2);
  // z5' = 3 t3 + 2 z5
  fq2_add(o, t3, z5);
  fq2_dbl(o, o);
  fq2_add(rp->c1.c2, o, t3);
}

constexpr uint64_t BLS_X = 0x8508c00000000001ull;  // ark-bls12-377 Config::X (positive)

// a^x for a unitary a (ark `exp_by_x` = cyclotomic_exp, X_IS_NEGATIVE = false)
TB_G2_OL void fq12_exp_by_x_ol(Fq12* rp, const Fq12* ap) {
  const Fq12 a = *ap;
  Fq12 acc = a;
  for (int bit = 62; bit >= 0; bit--) {
    fq12_cyclotomic_sqr_ol(&acc, &acc);
    if ((BLS_X >> bit) & 1) fq12_mul_ol(&acc, &acc, &a);
  }
  *rp = acc;
}
TB_HD void fq12_exp_by_x(Fq12& r, const Fq12& a) { fq12_exp_by_x_ol(&r, &a); }

// ark `Bls12::final_exponentiation`, step for step
TB_G2_OL void fq12_final_exp_ol(Fq12* outp, const Fq12* fp, int stop) {
  Fq12& out = *outp;
  const Fq12 f = *fp;
  Fq12 r, f2, y0, y1, y2;
  int step = 0;
#define TB_FE_STEP(x) if (++step == stop) { out = x; return; }
  fq12_conj(r, f);
  TB_FE_STEP(r)                // 1
  fq12_inv(f2, f);
  TB_FE_STEP(f2)               // 2
  fq12_mul(r, r, f2);          // f^(q^6 - 1)
  TB_FE_STEP(r)                // 3
  f2 = r;
  fq12_frobenius(r, r, 2);
  TB_FE_STEP(r)                // 4
  fq12_mul(r, r, f2);          // f^((q^6 - 1)(q^2 + 1))
  TB_FE_STEP(r)                // 5
  fq12_cyclotomic_sqr_ol(&y0, &r);
  TB_FE_STEP(y0)               // 6
  fq12_exp_by_x(y1, r);
  TB_FE_STEP(y1)               // 7
  fq12_conj(y2, r);
  fq12_mul(y1, y1, y2);
  TB_FE_STEP(y1)               // 8
  fq12_exp_by_x(y2, y1);
  TB_FE_STEP(y2)               // 9
  fq12_conj(y1, y1);
  TB_FE_STEP(y1)               // 10
  fq12_mul(y1, y1, y2);
  TB_FE_STEP(y1)               // 11
  fq12_exp_by_x(y2, y1);
  fq12_frobenius(y1, y1, 1);
  TB_FE_STEP(y1)               // 12
  fq12_mul(y1, y1, y2);
  fq12_mul(r, r, y0);
  fq12_exp_by_x(y0, y1);
  fq12_exp_by_x(y2, y0);
  fq12_frobenius(y0, y1, 2);
  fq12_conj(y1, y1);
  fq12_mul(y1, y1, y2);
  fq12_mul(y1, y1, y0);
  fq12_mul(out, r, y1);
#undef TB_FE_STEP
}
// `stop` < 1000 returns the value after that many steps of the chain (test hook: tests/test_gpu_pairing.py)
TB_HD void fq12_final_exp(Fq12& out, const Fq12& f, int stop = 1000) { fq12_final_exp_ol(&out, &f, stop); }

// ---- Miller loop ---------------------------------------------------------------------------------------------
struct G2Hom {  // homogeneous projective point on the twist (ark `G2HomProjective`)
  Fq2 x, y, z;
};

// (0 + b1 u) * t with B' = (0, b1) the twist coefficient: -5 b1 t1 + b1 t0 u
TB_HD void fq2_mul_twist_b(Fq2& e, const Fq2& t) {
  const Fq b1 = fq_from_table(FQ12_C(TWIST_B1));
  Fq m0, m1;
  fq_mul_ol(&m0, &t.c1, &b1);
  fq_mul_ol(&m1, &t.c0, &b1);
  fq_mul5(m0, m0);
  fq_neg(e.c0, m0);
  e.c1 = m1;
}

// ark `G2HomProjective::double_in_place` (TwistType::D): r = 2 r; the line through r, r evaluated at P as the three
// non-zero coefficients of ark's `ell` / `mul_by_034`: (l0, l3, l4) = (-h py, 3 j px, i)
TB_HD void g2_double_line(G2Hom& r, Fq2& l0, Fq2& l3, Fq2& l4, const Fq& px, const Fq& py) {
  Fq2 a, b, c, e, ff, g, h, i, j, e2, t;
  fq2_mul(a, r.x, r.y);
  fq_halve(a.c0, a.c0);
  fq_halve(a.c1, a.c1);
  fq2_sqr(b, r.y);
  fq2_sqr(c, r.z);
  fq2_dbl(t, c);
  fq2_add(t, t, c);            // 3 z^2
  fq2_mul_twist_b(e, t);       // e = B' 3 z^2
  fq2_dbl(ff, e);
  fq2_add(ff, ff, e);          // 3 e
  fq2_add(g, b, ff);
  fq_halve(g.c0, g.c0);
  fq_halve(g.c1, g.c1);
  fq2_add(t, r.y, r.z);
  fq2_sqr(h, t);
  fq2_add(t, b, c);
  fq2_sub(h, h, t);            // 2 y z
  fq2_sub(i, e, b);
  fq2_sqr(j, r.x);
  fq2_sqr(e2, e);
  fq2_sub(t, b, ff);
  fq2_mul(r.x, a, t);
  fq2_sqr(g, g);
  fq2_dbl(t, e2);
  fq2_add(t, t, e2);
  fq2_sub(r.y, g, t);
  fq2_mul(r.z, b, h);
  fq2_neg(l0, h);
  fq2_scale(l0, l0, py);
  fq2_dbl(t, j);
  fq2_add(t, t, j);
  fq2_scale(l3, t, px);
  l4 = i;
}
TB_HD void miller_double_step(Fq12& f, G2Hom& r, const Fq& px, const Fq& py) {
  Fq2 l0, l3, l4;
  g2_double_line(r, l0, l3, l4, px, py);
  fq12_mul_by_034_ol(&f, &l0, &l3, &l4);
}

// ark `G2HomProjective::add_in_place`: r = r + q; line coefficients (l0, l3, l4) = (lambda py, -theta px, j)
TB_HD void g2_add_line(G2Hom& r, Fq2& l0, Fq2& l3, Fq2& l4, const Affine2& q, const Fq& px, const Fq& py) {
  Fq2 theta, lambda, c, d, e, ff, g, h, j, t;
  fq2_mul(t, q.y, r.z);
  fq2_sub(theta, r.y, t);
  fq2_mul(t, q.x, r.z);
  fq2_sub(lambda, r.x, t);
  fq2_sqr(c, theta);
  fq2_sqr(d, lambda);
  fq2_mul(e, lambda, d);
  fq2_mul(ff, r.z, c);
  fq2_mul(g, r.x, d);
  fq2_add(h, e, ff);
  fq2_sub(h, h, g);
  fq2_sub(h, h, g);
  fq2_mul(r.x, lambda, h);
  fq2_sub(t, g, h);
  fq2_mul(t, theta, t);
  fq2_mul(g, e, r.y);
  fq2_sub(r.y, t, g);
  fq2_mul(r.z, r.z, e);
  fq2_mul(j, theta, q.x);
  fq2_mul(t, lambda, q.y);
  fq2_sub(j, j, t);
  fq2_scale(l0, lambda, py);
  fq2_neg(l3, theta);
  fq2_scale(l3, l3, px);
  l4 = j;
}
TB_HD void miller_add_step(Fq12& f, G2Hom& r, const Affine2& q, const Fq& px, const Fq& py) {
  Fq2 l0, l3, l4;
  g2_add_line(r, l0, l3, l4, q, px, py);
  fq12_mul_by_034_ol(&f, &l0, &l3, &l4);
}

// f_{x,Q}(P) (up to factors the final exponentiation removes); 1 if either point is the identity, as
// ark's multi_miller_loop skips such pairs
TB_G2_OL void miller_loop_ol(Fq12* fp, const Affine* pp, const Affine2* qp) {
  const Affine p = *pp;
  const Affine2 q = *qp;
  Fq12 f = fq12_one();
  if (affine_is_inf(p) || affine2_is_inf(q)) {
    *fp = f;
    return;
  }
  G2Hom r;
  r.x = q.x;
  r.y = q.y;
  r.z = fq2_one();
  for (int bit = 62; bit >= 0; bit--) {
    if (bit != 62) fq12_sqr(f, f);
    miller_double_step(f, r, p.x, p.y);
    if ((BLS_X >> bit) & 1) miller_add_step(f, r, q, p.x, p.y);
  }
  *fp = f;
}
TB_HD void miller_loop(Fq12& f, const Affine& p, const Affine2& q) { miller_loop_ol(&f, &p, &q); }

}  // namespace tb
