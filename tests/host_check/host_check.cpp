// Test-only shim: compiles the product's __host__ __device__ field/group headers with g++ (carry chains
// emulated) so the algorithm structure can be checked on a CPU-only box. Not part of the product library.
#include <cstring>
#include <initializer_list>
#include "../../testudo_b200/csrc/g1_fast.cuh"
#include "../../testudo_b200/csrc/digits.cuh"
#include "../../testudo_b200/csrc/experimental/mont_kara.cuh"
#include "../../testudo_b200/csrc/g2.cuh"
#include "../../testudo_b200/csrc/fq12.cuh"
#include "../../testudo_b200/csrc/fq12_coop.cuh"
using namespace tb;
extern "C" {
void hc_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { mont_mul<FqParams>(r, a, b); }
void hc_fq_sqr(const uint32_t* a, uint32_t* r) { mont_sqr<FqParams>(r, a); }
void hc_fq_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { mod_add<FqParams>(r, a, b); }
void hc_fq_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { mod_sub<FqParams>(r, a, b); }
void hc_fq_neg(const uint32_t* a, uint32_t* r) { mod_neg<FqParams>(r, a); }
void hc_fq_inv(const uint32_t* a, uint32_t* r) { Fq x, y; memcpy(x.l, a, 48); fq_inv(y, x); memcpy(r, y.l, 48); }
void hc_fq_inv_fermat(const uint32_t* a, uint32_t* r) { Fq x, y; memcpy(x.l, a, 48); fq_inv_fermat(y, x); memcpy(r, y.l, 48); }
void hc_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { mont_mul<FrParams>(r, a, b); }
void hc_fr_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { mod_add<FrParams>(r, a, b); }
// (a b + c d) R^-1 with ONE reduction, then one conditional subtraction: the pair product of k_fr_matvec
void hc_fr_mul2(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* r) {
  mont_mul2_lazy<FrParams>(r, a, b, c, d);
  mod_reduce_once<FrParams>(r);
}
void hc_fr_to_canonical(const uint32_t* a, uint32_t* r) { mont_to_canonical<FrParams>(r, a); }
void hc_consts(uint32_t* q, uint32_t* q_one, uint32_t* q_r2, uint32_t* r, uint32_t* r_one, uint32_t* r_r2) {
  for (int i = 0; i < 12; i++) { q[i] = FqParams::p(i); q_one[i] = FqParams::one(i); q_r2[i] = FqParams::r2(i); }
  for (int i = 0; i < 8; i++) { r[i] = FrParams::p(i); r_one[i] = FrParams::one(i); r_r2[i] = FrParams::r2(i); }
}
// acc (xyzz as affine in, inf allowed) + q -> affine out, exercising madd / add / dbl / to_affine
void hc_madd(const uint32_t* p_aff, const uint32_t* q_aff, uint32_t* out_aff) {
  Affine p, q, r; memcpy(&p, p_aff, 96); memcpy(&q, q_aff, 96);
  Xyzz acc; xyzz_from_affine(acc, p);
  xyzz_madd(acc, q);
  xyzz_to_affine(r, acc); memcpy(out_aff, &r, 96);
}
// (p scaled to a non-trivial ZZ by adding and subtracting t) + q, through the full XYZZ+XYZZ add
void hc_add(const uint32_t* p_aff, const uint32_t* q_aff, const uint32_t* t_aff, uint32_t* out_aff) {
  Affine p, q, t, r; memcpy(&p, p_aff, 96); memcpy(&q, q_aff, 96); memcpy(&t, t_aff, 96);
  Xyzz a, b; xyzz_from_affine(a, p); xyzz_from_affine(b, q);
  xyzz_madd(a, t); Affine nt = t; fq_neg(nt.y, nt.y); xyzz_madd(a, nt);   // a == p with ZZ != 1
  xyzz_madd(b, t); xyzz_madd(b, nt);                                    // b == q with ZZ != 1
  xyzz_add(a, b);
  xyzz_to_affine(r, a); memcpy(out_aff, &r, 96);
}
void hc_dbl(const uint32_t* p_aff, const uint32_t* t_aff, uint32_t* out_aff) {
  Affine p, t, r; memcpy(&p, p_aff, 96); memcpy(&t, t_aff, 96);
  Xyzz a; xyzz_from_affine(a, p);
  xyzz_madd(a, t); Affine nt = t; fq_neg(nt.y, nt.y); xyzz_madd(a, nt);
  xyzz_dbl(a);
  xyzz_to_affine(r, a); memcpy(out_aff, &r, 96);
}
void hc_scalar_mul(const uint32_t* p_aff, const uint32_t* k, uint32_t* out_aff) {
  Affine p, r; memcpy(&p, p_aff, 96);
  Xyzz a; xyzz_scalar_mul(a, p, k);
  xyzz_to_affine(r, a); memcpy(out_aff, &r, 96);
}
// hot-loop lazy madd: acc = sum of n affine points (with sign flags), then canonical affine out. Also reports the
// largest limb-12 headroom seen (bound check): returns 1 if any coordinate ever reached 2^383.
int hc_madd_fast_chain(const uint32_t* pts_aff, const uint32_t* neg, int n, uint32_t* out_aff) {
  Xyzz acc; xyzz_set_inf(acc);
  int overflow = 0;
  for (int i = 0; i < n; i++) {
    Affine q; memcpy(&q, pts_aff + 24 * i, 96);
    if (neg[i]) fq_neg(q.y, q.y);
    xyzz_madd_fast(acc, q);
    if ((acc.x.l[11] | acc.y.l[11] | acc.zz.l[11] | acc.zzz.l[11]) >> 31) overflow = 1;
  }
  xyzz_canon(acc);
  Affine r; xyzz_to_affine(r, acc); memcpy(out_aff, &r, 96);
  return overflow;
}
// lazy add / double chains: acc = sum_i (p_i), each p_i first pushed to a non-trivial XYZZ representation; then
// `dbls` doublings; canonical affine out
int hc_add_fast_chain(const uint32_t* pts_aff, int n, int dbls, const uint32_t* t_aff, uint32_t* out_aff) {
  Affine t; memcpy(&t, t_aff, 96);
  Affine nt = t; fq_neg(nt.y, nt.y);
  Xyzz acc; xyzz_set_inf(acc);
  int overflow = 0;
  for (int i = 0; i < n; i++) {
    Affine q; memcpy(&q, pts_aff + 24 * i, 96);
    Xyzz x; xyzz_from_affine(x, q);
    if (!xyzz_is_inf(x) && (i & 1)) { xyzz_madd_fast(x, t); xyzz_madd_fast(x, nt); }  // ZZ != 1, lazy coords
    xyzz_add_fast(acc, x);
    if ((acc.x.l[11] | acc.y.l[11] | acc.zz.l[11] | acc.zzz.l[11]) >> 31) overflow = 1;
  }
  for (int k = 0; k < dbls; k++) {
    xyzz_dbl_fast(acc);
    if ((acc.x.l[11] | acc.y.l[11] | acc.zz.l[11] | acc.zzz.l[11]) >> 31) overflow = 1;
  }
  xyzz_canon(acc);
  Affine r; xyzz_to_affine(r, acc); memcpy(out_aff, &r, 96);
  return overflow;
}
void hc_fq_canon(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.l, a, 48); fq_canon(x); memcpy(r, x.l, 48); }
void hc_fq_sqr_lazy(const uint32_t* a, uint32_t* r) { mont_sqr_lazy<FqParams>(r, a); }
void hc_fr_sqr_lazy(const uint32_t* a, uint32_t* r) { mont_sqr_lazy<FrParams>(r, a); }
void hc_kara_mul12(const uint32_t* a, const uint32_t* b, uint32_t* T) { kara_mul12(T, a, b); }
void hc_fq_mul_kara(const uint32_t* a, const uint32_t* b, uint32_t* r) { mont_mul_kara(r, a, b); }
void hc_fq_mul2_kara(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* r) {
  mont_mul2_kara(r, a, b, c, d);
}
void hc_fq_mul2_lazy(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* r) {
  mont_mul2_lazy<FqParams>(r, a, b, c, d);
}
// G2 (g2.cuh): Fq2 arithmetic and the XYZZ group law over the twist
void hc_fq2_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fq2 x, y, z; memcpy(&x, a, 96); memcpy(&y, b, 96); fq2_mul(z, x, y); memcpy(r, &z, 96); }
void hc_fq2_sqr(const uint32_t* a, uint32_t* r) { Fq2 x, z; memcpy(&x, a, 96); fq2_sqr(z, x); memcpy(r, &z, 96); }
void hc_fq2_inv(const uint32_t* a, uint32_t* r) { Fq2 x, z; memcpy(&x, a, 96); fq2_inv(z, x); memcpy(r, &z, 96); }
static void g2_scale(Xyzz2& a, const Affine2& t) {  // same point, ZZ != 1
  xyzz2_madd(a, t); Affine2 nt = t; fq2_neg(nt.y, nt.y); xyzz2_madd(a, nt);
}
void hc_g2_madd(const uint32_t* p_aff, const uint32_t* q_aff, uint32_t* out_aff) {
  Affine2 p, q, r; memcpy(&p, p_aff, 192); memcpy(&q, q_aff, 192);
  Xyzz2 acc; xyzz2_set_inf(acc); xyzz2_madd(acc, p);
  xyzz2_madd(acc, q);
  xyzz2_to_affine(r, acc); memcpy(out_aff, &r, 192);
}
void hc_g2_add(const uint32_t* p_aff, const uint32_t* q_aff, const uint32_t* t_aff, uint32_t* out_aff) {
  Affine2 p, q, t, r; memcpy(&p, p_aff, 192); memcpy(&q, q_aff, 192); memcpy(&t, t_aff, 192);
  Xyzz2 a, b; xyzz2_set_inf(a); xyzz2_madd(a, p); xyzz2_set_inf(b); xyzz2_madd(b, q);
  g2_scale(a, t); g2_scale(b, t);
  xyzz2_add(a, b);
  xyzz2_to_affine(r, a); memcpy(out_aff, &r, 192);
}
void hc_g2_dbl(const uint32_t* p_aff, const uint32_t* t_aff, uint32_t* out_aff) {
  Affine2 p, t, r; memcpy(&p, p_aff, 192); memcpy(&t, t_aff, 192);
  Xyzz2 a; xyzz2_set_inf(a); xyzz2_madd(a, p); g2_scale(a, t);
  xyzz2_dbl(a);
  xyzz2_to_affine(r, a); memcpy(out_aff, &r, 192);
}
void hc_g2_scalar_mul(const uint32_t* p_aff, const uint32_t* k, uint32_t* out_aff) {
  Affine2 p, r; memcpy(&p, p_aff, 192);
  Xyzz2 a; xyzz2_scalar_mul(a, p, k);
  xyzz2_to_affine(r, a); memcpy(out_aff, &r, 192);
}
void hc_fq_mul_lazy(const uint32_t* a, const uint32_t* b, uint32_t* r) { mont_mul_lazy<FqParams>(r, a, b); }
// signed digits of one canonical scalar with window c: out[w] in [-2^(c-1), 2^(c-1)]
int hc_digits(const uint32_t* s, int c, int32_t* out) {
  int W = num_windows(c);
  DigitIter it(s, c);
  for (int w = 0; w < W; w++) out[w] = it.next(w == W - 1);
  return W;
}
// ---- Fq12 tower / pairing (fq12.cuh): 144 u32 per Fq12 in ark's in-memory order -------------------------------------
void hc_fq12_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) {
  Fq12 x, y; memcpy(&x, a, 576); memcpy(&y, b, 576); fq12_mul(x, x, y); memcpy(r, &x, 576);
}
void hc_fq12_sqr(const uint32_t* a, uint32_t* r) { Fq12 x; memcpy(&x, a, 576); fq12_sqr(x, x); memcpy(r, &x, 576); }
void hc_fq12_inv(const uint32_t* a, uint32_t* r) { Fq12 x; memcpy(&x, a, 576); fq12_inv(x, x); memcpy(r, &x, 576); }
void hc_fq12_frobenius(const uint32_t* a, int k, uint32_t* r) {
  Fq12 x; memcpy(&x, a, 576); fq12_frobenius(x, x, k); memcpy(r, &x, 576);
}
void hc_fq12_cyclotomic_sqr(const uint32_t* a, uint32_t* r) {
  Fq12 x; memcpy(&x, a, 576); fq12_cyclotomic_sqr_ol(&x, &x); memcpy(r, &x, 576);
}
void hc_fq12_mul_by_034(const uint32_t* a, const uint32_t* l0, const uint32_t* l3, const uint32_t* l4, uint32_t* r) {
  Fq12 x; Fq2 c0, c3, c4; memcpy(&x, a, 576); memcpy(&c0, l0, 96); memcpy(&c3, l3, 96); memcpy(&c4, l4, 96);
  fq12_mul_by_034_ol(&x, &c0, &c3, &c4); memcpy(r, &x, 576);
}
void hc_fq12_exp_by_x(const uint32_t* a, uint32_t* r) { Fq12 x; memcpy(&x, a, 576); fq12_exp_by_x(x, x); memcpy(r, &x, 576); }
void hc_miller_loop(const uint32_t* p_aff, const uint32_t* q_aff, uint32_t* r) {
  Affine p; Affine2 q; Fq12 f; memcpy(&p, p_aff, 96); memcpy(&q, q_aff, 192);
  miller_loop(f, p, q); memcpy(r, &f, 576);
}
void hc_final_exp(const uint32_t* a, uint32_t* r) { Fq12 x, y; memcpy(&x, a, 576); fq12_final_exp(y, x); memcpy(r, &y, 576); }

// ---- the cooperative engine's lazily reduced per-item bodies (fq12_coop.cuh), run item by item in sequence -----------------
}  // extern "C"
namespace {
void hcw_fq6_products(WScratch* w, int count) {
  for (int t = 0; t < 18 * count; t++) wp_kar(w, t);
  for (int t = 0; t < 12 * count; t++) wp_fq2(w, t);
  for (int t = 0; t < 6 * count; t++) wp_fq6(w, t);
}
void hcw_mul(Fq12* dst, const Fq12* a, const Fq12* b) {
  WScratch w;
  for (int t = 0; t < 36; t++) wp_mul_xy(&w, a, b, t);
  hcw_fq6_products(&w, 3);
  for (int t = 0; t < 12; t++) wp_mul_out(dst, &w, t);
}
void hcw_sqr(Fq12* dst, const Fq12* a) {
  WScratch w;
  for (int t = 0; t < 24; t++) wp_sqr_xy(&w, a, t);
  hcw_fq6_products(&w, 2);
  for (int t = 0; t < 12; t++) wp_sqr_out(dst, &w, t);
}
void hcw_cyc(Fq12* dst, const Fq12* a) {
  WScratch w;
  for (int t = 0; t < 18; t++) wp_cyc_kar(&w, a, t);
  for (int t = 0; t < 12; t++) wp_cyc_fq2(&w, t);
  for (int t = 0; t < 12; t++) wp_cyc_out(dst, a, &w, t);
}
void hcw_conj(Fq12* dst, const Fq12* a) {
  Fq v[12];
  for (int t = 0; t < 12; t++) v[t] = wp_conj(a, t);
  for (int t = 0; t < 12; t++) w12_q(dst)[t] = v[t];
}
void hcw_frob(Fq12* dst, const Fq12* a, int k) {
  Fq v[12];
  for (int t = 0; t < 12; t++) v[t] = wp_frobenius(a, k, t);
  for (int t = 0; t < 12; t++) w12_q(dst)[t] = v[t];
}
void hcw_canon(Fq12* a) {
  for (int t = 0; t < 12; t++) lz_canon(w12_q(a)[t]);
}
void hcw_exp_by_x(Fq12* dst, const Fq12* a) {
  Fq12 acc = *a;
  for (int bit = 62; bit >= 0; bit--) {
    hcw_cyc(&acc, &acc);
    if ((BLS_X >> bit) & 1) hcw_mul(&acc, &acc, a);
  }
  *dst = acc;
}
// mirrors w12_final_exp (kernels_pairing.cuh)
void hcw_final_exp(Fq12* out, const Fq12* in) {
  Fq12 f = *in, r, f2, y0, y1, y2;
  hcw_conj(&r, &f);
  hcw_mul(&y0, &f, &r);
  {
    WScratch w;
    for (int t = 0; t < 18; t++) wp_inv6_r1(&w, &y0, t);
    for (int t = 0; t < 12; t++) wp_inv6_p1(&w, t);
    for (int t = 0; t < 6; t++) wp_inv6_p2(&w, t);
    for (int t = 0; t < 9; t++) wp_inv6_r2(&w, &y0, t);
    wp_inv6_d(&w);
    for (int t = 0; t < 9; t++) wp_inv6_r3(&w, t);
    for (int t = 0; t < 12; t++) wp_inv6_out(&f2, &w, t);
  }
  hcw_sqr(&r, &r);
  hcw_mul(&r, &r, &f2);
  f2 = r;
  hcw_frob(&r, &r, 2);
  hcw_mul(&r, &r, &f2);
  hcw_cyc(&y0, &r);
  hcw_exp_by_x(&y1, &r);
  hcw_conj(&y2, &r);
  hcw_mul(&y1, &y1, &y2);
  hcw_exp_by_x(&y2, &y1);
  hcw_conj(&y1, &y1);
  hcw_mul(&y1, &y1, &y2);
  hcw_exp_by_x(&y2, &y1);
  hcw_frob(&y1, &y1, 1);
  hcw_mul(&y1, &y1, &y2);
  hcw_mul(&r, &r, &y0);
  hcw_exp_by_x(&y0, &y1);
  hcw_exp_by_x(&y2, &y0);
  hcw_frob(&y0, &y1, 2);
  hcw_conj(&y1, &y1);
  hcw_mul(&y1, &y1, &y2);
  hcw_mul(&y1, &y1, &y0);
  hcw_mul(&r, &r, &y1);
  hcw_canon(&r);
  *out = r;
}
void hcw_double_step(WDouble* d, Fq12* line) {
  for (int t = 0; t < 11; t++) wp_dbl_r1(d, t);
  for (int t = 0; t < 12; t++) wp_dbl_p2(d, t);
  for (int t = 0; t < 10; t++) wp_dbl_p3(d, t);
  for (int t = 0; t < 14; t++) wp_dbl_r2(d, t);
  for (int t = 0; t < 12; t++) wp_dbl_p5(d, line, t);
}
void hcw_add_step(WDouble* d, const Affine2* q, Fq12* line) {
  for (int t = 0; t < 6; t++) wp_add_rA(d, q, t);
  for (int t = 0; t < 4; t++) wp_add_pA(d, t);
  for (int t = 0; t < 14; t++) wp_add_rB(d, q, t);
  for (int t = 0; t < 10; t++) wp_add_pB(d, line, t);
  for (int t = 0; t < 9; t++) wp_add_rC(d, t);
  for (int t = 0; t < 6; t++) wp_add_pC(d, t);
  for (int t = 0; t < 12; t++) wp_add_rD(d, t);
  for (int t = 0; t < 6; t++) wp_add_pD(d, t);
}
}  // namespace
extern "C" {
// lane-parallel XYZZ arithmetic over Fq2 (window combine of a G2 MSM) vs the canonical routines. p, e: XYZZ (4 Fq2 each),
// canonical; op 0: p = 2 p, 1: p += e. Returns 1 when the canonicalised results agree.
int hc_coop_g2_op(int op, const uint32_t* p_in, const uint32_t* e_in, uint32_t* out) {
  WG2 s;
  memcpy(&s.p, p_in, 384);
  memcpy(&s.e, e_in, 384);
  Xyzz2 ref = s.p;
  if (op == 0) {
    xyzz2_dbl(ref);
    for (int t = 0; t < 4; t++) wp_g2dbl_r1(&s, t);
    for (int t = 0; t < 6; t++) wp_g2dbl_p1(&s, t);
    for (int t = 0; t < 11; t++) wp_g2dbl_r2(&s, t);
    for (int t = 0; t < 8; t++) wp_g2dbl_p2(&s, t);
    for (int t = 0; t < 9; t++) wp_g2dbl_r3(&s, t);
    for (int t = 0; t < 4; t++) wp_g2dbl_p3(&s, t);
  } else {
    xyzz2_add(ref, s.e);
    wp_g2add_flags(&s);
    if (s.flag == 2) s.p = s.e;
    if (s.flag == 0) {
      for (int t = 0; t < 12; t++) wp_g2add_r1(&s, t);
      for (int t = 0; t < 8; t++) wp_g2add_p1(&s, t);
      wp_g2add_check(&s);
    }
    if (s.flag == 0) {
      for (int t = 0; t < 10; t++) wp_g2add_r2(&s, t);
      for (int t = 0; t < 8; t++) wp_g2add_p2(&s, t);
      for (int t = 0; t < 9; t++) wp_g2add_r3(&s, t);
      for (int t = 0; t < 8; t++) wp_g2add_p3(&s, t);
      for (int t = 0; t < 9; t++) wp_g2add_r4(&s, t);
      for (int t = 0; t < 4; t++) wp_g2add_p4(&s, t);
    }
  }
  for (int t = 0; t < 8; t++) lz_canon(reinterpret_cast<Fq*>(&s.p)[t]);
  memcpy(out, &s.p, 384);
  // compare as points: both to canonical affine (the XYZZ representatives agree too, but affine is what leaves the kernel)
  Affine2 a, b;
  xyzz2_to_affine(a, s.p);
  xyzz2_to_affine(b, ref);
  return memcmp(&a, &b, 192) == 0 && (op == 1 && s.flag == 3 ? 1 : memcmp(&s.p, &ref, 384) == 0 || xyzz2_is_inf(ref));
}
// product of two line values (tower slots 0, 3, 4 of a and b; the other slots are ignored) vs the general product
int hc_coop_line_mul(const uint32_t* a, const uint32_t* b) {
  Fq12 x, y, z, ref;
  memcpy(&x, a, 576);
  memcpy(&y, b, 576);
  for (int sl : {1, 2, 5}) { *w12_c(&x, sl) = fq2_zero(); *w12_c(&y, sl) = fq2_zero(); }
  WScratch w;
  for (int t = 0; t < 12; t++) wp_ll_xy(&w, &x, &y, t);
  for (int t = 0; t < 18; t++) wp_kar(&w, t);
  for (int t = 0; t < 12; t++) wp_fq2(&w, t);
  for (int t = 0; t < 12; t++) wp_ll_out(&z, &w, t);
  hcw_canon(&z);
  fq12_mul(ref, x, y);
  return memcmp(&z, &ref, 576) == 0;
}
// lazily reduced addition step vs the canonical g2_add_line: r (3 Fq2), q (2 Fq2), px, py canonical; 1 when they agree
int hc_coop_add_step(const uint32_t* r_in, const uint32_t* q_in, const uint32_t* px, const uint32_t* py) {
  WDouble d;
  Fq12 line;
  memset(&line, 0, sizeof line);
  Affine2 q;
  memcpy(&d.r, r_in, 288);
  memcpy(&q, q_in, 192);
  memcpy(&d.px, px, 48);
  memcpy(&d.py, py, 48);
  G2Hom r = d.r;
  hcw_add_step(&d, &q, &line);
  Fq2 l0, l3, l4;
  g2_add_line(r, l0, l3, l4, q, d.px, d.py);
  return memcmp(&r, &d.r, 288) == 0 && memcmp(&l0, w12_c(&line, 0), 96) == 0 && memcmp(&l3, w12_c(&line, 3), 96) == 0 &&
         memcmp(&l4, w12_c(&line, 4), 96) == 0;
}
// op: 0 mul, 1 sqr, 2 cyclotomic sqr, 3 frobenius 1, 4 frobenius 2, 5 conj, 6 the chain ((a b)^2 a)^2 conj, 7 exp_by_x,
// 8 final exponentiation. a, b: raw 12-limb coefficients (any representative < 1.02 q); r canonical.
void hc_coop_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
  Fq12 x, y, z;
  memcpy(&x, a, 576);
  memcpy(&y, b, 576);
  switch (op) {
    case 0: hcw_mul(&z, &x, &y); break;
    case 1: hcw_sqr(&z, &x); break;
    case 2: hcw_cyc(&z, &x); break;
    case 3: hcw_frob(&z, &x, 1); break;
    case 4: hcw_frob(&z, &x, 2); break;
    case 5: hcw_conj(&z, &x); break;
    case 6:
      hcw_mul(&z, &x, &y);
      hcw_sqr(&z, &z);
      hcw_mul(&z, &z, &x);
      hcw_sqr(&z, &z);
      hcw_conj(&z, &z);
      break;
    case 7: hcw_exp_by_x(&z, &x); break;
    default: hcw_final_exp(&z, &x); break;
  }
  hcw_canon(&z);
  memcpy(r, &z, 576);
}
// the chain of k_fq12_pow_coop: a^e by square-and-multiply from the top set bit with the GENERIC lazily reduced square /
// product (e: 8 canonical 32-bit limbs, not zero); r canonical
void hc_coop_pow(const uint32_t* a, const uint32_t* e, uint32_t* r) {
  Fq12 x, z;
  memcpy(&x, a, 576);
  int top = -1;
  for (int i = 7; i >= 0 && top < 0; i--)
    if (e[i]) top = 32 * i + 31 - __builtin_clz(e[i]);
  z = x;
  for (int bit = top - 1; bit >= 0; bit--) {
    hcw_sqr(&z, &z);
    if ((e[bit >> 5] >> (bit & 31)) & 1) hcw_mul(&z, &z, &x);
  }
  hcw_canon(&z);
  memcpy(r, &z, 576);
}
// the largest top limb over the 12 output coefficients of op BEFORE canonicalisation (the invariant: < 1.02 q)
uint32_t hc_coop_op_top(int op, const uint32_t* a, const uint32_t* b) {
  Fq12 x, y, z;
  memcpy(&x, a, 576);
  memcpy(&y, b, 576);
  if (op == 0) hcw_mul(&z, &x, &y);
  else if (op == 1) hcw_sqr(&z, &x);
  else if (op == 2) hcw_cyc(&z, &x);
  else if (op == 3) hcw_frob(&z, &x, 1);
  else if (op == 4) hcw_frob(&z, &x, 2);
  else hcw_conj(&z, &x);
  uint32_t top = 0;
  for (int t = 0; t < 12; t++) top = w12_q(&z)[t].l[11] > top ? w12_q(&z)[t].l[11] : top;
  return top;
}
// lazily reduced doubling step vs the canonical g2_double_line on the same inputs: r (3 Fq2, canonical), px, py.
// Writes both results (r: 3 Fq2, line: l0, l3, l4) and returns 1 when they agree.
int hc_coop_double_step(const uint32_t* r_in, const uint32_t* px, const uint32_t* py, uint32_t* r_out, uint32_t* line_out) {
  WDouble d;
  Fq12 line;
  memset(&line, 0, sizeof line);
  memcpy(&d.r, r_in, 288);
  memcpy(&d.px, px, 48);
  memcpy(&d.py, py, 48);
  hcw_double_step(&d, &line);
  G2Hom r;
  memcpy(&r, r_in, 288);
  Fq2 l0, l3, l4;
  Fq x, y;
  memcpy(&x, px, 48);
  memcpy(&y, py, 48);
  g2_double_line(r, l0, l3, l4, x, y);
  memcpy(r_out, &d.r, 288);
  memcpy(line_out, w12_c(&line, 0), 96);
  memcpy(line_out + 24, w12_c(&line, 3), 96);
  memcpy(line_out + 48, w12_c(&line, 4), 96);
  return memcmp(&r, &d.r, 288) == 0 && memcmp(&l0, w12_c(&line, 0), 96) == 0 && memcmp(&l3, w12_c(&line, 3), 96) == 0 &&
         memcmp(&l4, w12_c(&line, 4), 96) == 0;
}
// the cooperative Miller loop as w_miller_loop runs it (kernels_pairing.cuh), item by item
void hc_coop_miller(const uint32_t* p_aff, const uint32_t* q_aff, uint32_t* out) {
  Affine p;
  Affine2 q;
  memcpy(&p, p_aff, 96);
  memcpy(&q, q_aff, 192);
  Fq12 f = fq12_one(), line;
  memset(&line, 0, sizeof line);
  if (!(affine_is_inf(p) || affine2_is_inf(q))) {
    WDouble d;
    d.r.x = q.x;
    d.r.y = q.y;
    d.r.z = fq2_one();
    d.px = p.x;
    d.py = p.y;
    for (int bit = 62; bit >= 0; bit--) {
      if (bit != 62) hcw_sqr(&f, &f);
      hcw_double_step(&d, &line);
      hcw_mul(&f, &f, &line);
      if ((BLS_X >> bit) & 1) {
        hcw_add_step(&d, &q, &line);
        hcw_mul(&f, &f, &line);
      }
    }
  }
  hcw_canon(&f);
  memcpy(out, &f, 576);
}
// `iters` random products / squarings / doubling steps on representatives biased towards the extremes (0, q - 1, q,
// 1.02 q - 1): lazily reduced bodies vs the canonical tower. Returns the number of mismatches.
int hc_coop_stress(uint64_t seed, int iters) {
  auto next = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return seed; };
  Fq qm1, top, qq;
  for (int i = 0; i < 12; i++) qq.l[i] = FqParams::p(i);
  qm1 = qq; qm1.l[0] -= 1;
  // top = q + floor(q / 64) - 1 < 1.02 q
  { Fq sh; for (int i = 0; i < 11; i++) sh.l[i] = (qq.l[i] >> 6) | (qq.l[i + 1] << 26); sh.l[11] = qq.l[11] >> 6; lz_add(top, qq, sh); top.l[0] -= 2; }
  auto rnd = [&](Fq& v) {
    const uint64_t sel = next() % 8;
    if (sel == 0) v = fq_zero();
    else if (sel == 1) v = qm1;
    else if (sel == 2) v = qq;
    else if (sel == 3) v = top;
    else { for (int i = 0; i < 12; i++) v.l[i] = (uint32_t)next(); v.l[11] %= 0x01ae3a46u; }
  };
  int bad = 0;
  for (int it = 0; it < iters; it++) {
    Fq12 a, b, ca, cb, z, ref;
    for (int t = 0; t < 12; t++) { rnd(w12_q(&a)[t]); rnd(w12_q(&b)[t]); }
    ca = a; cb = b;
    hcw_canon(&ca); hcw_canon(&cb);
    hcw_mul(&z, &a, &b); hcw_canon(&z);
    fq12_mul(ref, ca, cb);
    bad += memcmp(&z, &ref, 576) != 0;
    hcw_sqr(&z, &a); hcw_canon(&z);
    fq12_sqr(ref, ca);
    bad += memcmp(&z, &ref, 576) != 0;
    // doubling step on canonical inputs
    WDouble d; Fq12 line; memset(&line, 0, sizeof line);
    memcpy(&d.r, &ca, 288); d.px = w12_q(&cb)[0]; d.py = w12_q(&cb)[1];
    G2Hom r; memcpy(&r, &ca, 288);
    Fq2 l0, l3, l4;
    hcw_double_step(&d, &line);
    g2_double_line(r, l0, l3, l4, w12_q(&cb)[0], w12_q(&cb)[1]);
    bad += !(memcmp(&r, &d.r, 288) == 0 && memcmp(&l0, w12_c(&line, 0), 96) == 0 && memcmp(&l3, w12_c(&line, 3), 96) == 0 &&
             memcmp(&l4, w12_c(&line, 4), 96) == 0);
    // addition step: r = the first three Fq2 of ca, q = the next two
    Affine2 q; memcpy(&q, w12_c(&ca, 3), 192);
    memcpy(&d.r, &ca, 288); memcpy(&r, &ca, 288);
    memset(&line, 0, sizeof line);
    hcw_add_step(&d, &q, &line);
    g2_add_line(r, l0, l3, l4, q, d.px, d.py);
    bad += !(memcmp(&r, &d.r, 288) == 0 && memcmp(&l0, w12_c(&line, 0), 96) == 0 && memcmp(&l3, w12_c(&line, 3), 96) == 0 &&
             memcmp(&l4, w12_c(&line, 4), 96) == 0);
  }
  return bad;
}
}
