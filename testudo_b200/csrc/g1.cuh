// BLS12-377 G1 group law for the MSM kernels: affine inputs, XYZZ accumulators (EFD shortw/xyzz, a = 0).
//
// Replaces ark-ec 0.4 `short_weierstrass::{Affine, Projective}<ark_bls12_377::g1::Config>` arithmetic that
// every reference call site reaches through `VariableBaseMSM`. arkworks uses
// Jacobian coordinates; the representation is free because every consumer takes `.into_affine()`
// (src/sqrt_pst.rs:198, src/mipp.rs:117,363) or serialises through the affine form.
// XYZZ is chosen because the mixed addition costs 8M+2S (vs 7M+4S Jacobian) and needs no field doubling
// chains. All exceptional cases are exact: P+inf, inf+P, P+P (doubling branch), P+(-P) = inf.
//
// Wire format (C ABI): affine point = x[12] || y[12] u32 (== ark's u64[6] || u64[6]), Montgomery form;
// all-zero == identity ((0,0) is not on y^2 = x^3 + 1).
#pragma once
#include "mont.cuh"

// cold paths (exceptional cases of the group law, conversions): out-of-line on the device so the hot bucket
// loop keeps its register budget and instruction footprint
#if defined(__CUDA_ARCH__)
#define TB_COLD __device__ __noinline__
#elif defined(__CUDACC__)
#define TB_COLD __host__ __device__ inline
#else
#define TB_COLD inline
#endif

namespace tb {

struct Affine {
  Fq x, y;
};
struct Xyzz {  // x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity <=> ZZ == 0
  Fq x, y, zz, zzz;
};

TB_HD bool affine_is_inf(const Affine& p) { return fq_is_zero(p.x) && fq_is_zero(p.y); }
TB_HD bool xyzz_is_inf(const Xyzz& p) { return fq_is_zero(p.zz); }
TB_HD void xyzz_set_inf(Xyzz& p) {
  p.x = fq_zero();
  p.y = fq_zero();
  p.zz = fq_zero();
  p.zzz = fq_zero();
}
TB_HD void xyzz_from_affine(Xyzz& r, const Affine& p) {
  if (affine_is_inf(p)) {
    xyzz_set_inf(r);
    return;
  }
  r.x = p.x;
  r.y = p.y;
  r.zz = fq_one();
  r.zzz = fq_one();
}

// r = 2 * p for affine p != inf (EFD mdbl-2008-s-1, a = 0)
TB_COLD void xyzz_dbl_affine(Xyzz& r, const Affine& p) {
  Fq u, v, w, s, m, t;
  fq_dbl(u, p.y);
  fq_sqr(v, u);
  fq_mul(w, u, v);
  fq_mul(s, p.x, v);
  fq_sqr(m, p.x);
  fq_dbl(t, m);
  fq_add(m, t, m);  // 3 x^2
  fq_sqr(r.x, m);
  fq_sub(r.x, r.x, s);
  fq_sub(r.x, r.x, s);
  fq_sub(t, s, r.x);
  fq_mul(t, m, t);
  fq_mul(u, w, p.y);
  fq_sub(r.y, t, u);
  r.zz = v;
  r.zzz = w;
  // y == 0 cannot happen in the prime-order subgroup; if it did, v = w = 0 encodes the identity correctly
}

// p = 2 * p (EFD dbl-2008-s-1, a = 0)
TB_HD void xyzz_dbl(Xyzz& p) {
  if (xyzz_is_inf(p)) return;
  Fq u, v, w, s, m, t;
  fq_dbl(u, p.y);
  fq_sqr(v, u);
  fq_mul(w, u, v);
  fq_mul(s, p.x, v);
  fq_sqr(m, p.x);
  fq_dbl(t, m);
  fq_add(m, t, m);
  fq_mul(t, w, p.y);  // W * Y1 (uses old Y)
  fq_sqr(p.x, m);
  fq_sub(p.x, p.x, s);
  fq_sub(p.x, p.x, s);
  fq_sub(s, s, p.x);
  fq_mul(s, m, s);
  fq_sub(p.y, s, t);
  fq_mul(p.zz, v, p.zz);
  fq_mul(p.zzz, w, p.zzz);
}

// p += q, q affine (EFD madd-2008-s); the bucket-accumulation workhorse: 8M + 2S
TB_HD void xyzz_madd(Xyzz& p, const Affine& q) {
  if (affine_is_inf(q)) return;
  if (xyzz_is_inf(p)) {
    p.x = q.x;
    p.y = q.y;
    p.zz = fq_one();
    p.zzz = fq_one();
    return;
  }
  Fq pp, rr, t, ppp, qq;
  fq_mul(pp, q.x, p.zz);   // U2
  fq_mul(rr, q.y, p.zzz);  // S2
  fq_sub(pp, pp, p.x);     // P
  fq_sub(rr, rr, p.y);     // R
  if (fq_is_zero(pp)) {
    if (fq_is_zero(rr)) {
      // P + P: out-of-line doubling on copies, so that neither the accumulator nor the loaded point ever has
      // its address taken on the hot path (otherwise ptxas keeps them in local memory for the whole loop)
      Affine qc = q;
      Xyzz d;
      xyzz_dbl_affine(d, qc);
      p = d;
    } else {
      xyzz_set_inf(p);
    }
    return;
  }
  fq_sqr(t, pp);         // PP
  fq_mul(ppp, pp, t);    // PPP
  fq_mul(qq, p.x, t);    // Q
  fq_mul(p.zz, p.zz, t);
  fq_mul(p.zzz, p.zzz, ppp);
  fq_sqr(t, rr);
  fq_sub(t, t, ppp);
  fq_sub(t, t, qq);
  fq_sub(p.x, t, qq);    // X3 = R^2 - PPP - 2Q
  fq_sub(qq, qq, p.x);
  fq_mul(qq, rr, qq);
  fq_mul(t, p.y, ppp);
  fq_sub(p.y, qq, t);    // Y3 = R (Q - X3) - Y1 PPP
}

// p += q, both XYZZ (EFD add-2008-s): 12M + 2S
TB_HD void xyzz_add(Xyzz& p, const Xyzz& q) {
  if (xyzz_is_inf(q)) return;
  if (xyzz_is_inf(p)) {
    p = q;
    return;
  }
  Fq u1, u2, s1, s2, pp, ppp, t;
  fq_mul(u1, p.x, q.zz);
  fq_mul(u2, q.x, p.zz);
  fq_mul(s1, p.y, q.zzz);
  fq_mul(s2, q.y, p.zzz);
  fq_sub(u2, u2, u1);  // P
  fq_sub(s2, s2, s1);  // R
  if (fq_is_zero(u2)) {
    if (fq_is_zero(s2)) xyzz_dbl(p);
    else xyzz_set_inf(p);
    return;
  }
  fq_sqr(pp, u2);
  fq_mul(ppp, u2, pp);
  fq_mul(u1, u1, pp);  // Q
  fq_mul(t, p.zz, q.zz);
  fq_mul(p.zz, t, pp);
  fq_mul(t, p.zzz, q.zzz);
  fq_mul(p.zzz, t, ppp);
  fq_sqr(t, s2);
  fq_sub(t, t, ppp);
  fq_sub(t, t, u1);
  fq_sub(p.x, t, u1);
  fq_sub(u1, u1, p.x);
  fq_mul(u1, s2, u1);
  fq_mul(t, s1, ppp);
  fq_sub(p.y, u1, t);
}

TB_HD void xyzz_neg(Xyzz& p) { fq_neg(p.y, p.y); }

// canonical affine output (one field inversion); identity -> all-zero
TB_HD void xyzz_to_affine(Affine& r, const Xyzz& p) {
  if (xyzz_is_inf(p)) {
    r.x = fq_zero();
    r.y = fq_zero();
    return;
  }
  Fq t, ti, a, b;
  fq_mul(t, p.zz, p.zzz);
  fq_inv(ti, t);
  fq_mul(a, ti, p.zzz);  // 1/ZZ
  fq_mul(b, ti, p.zz);   // 1/ZZZ
  fq_mul(r.x, p.x, a);
  fq_mul(r.y, p.y, b);
}

// k * p by left-to-right double-and-add over a canonical 8-limb scalar (MIPP `compress`, src/mipp.rs:354-367)
TB_HD void xyzz_scalar_mul(Xyzz& r, const Affine& p, const uint32_t k[8]) {
  xyzz_set_inf(r);
  bool started = false;
  for (int i = 7; i >= 0; i--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) xyzz_dbl(r);
      if ((k[i] >> bit) & 1) {
        xyzz_madd(r, p);
        started = true;
      }
    }
  }
}

}  // namespace tb
