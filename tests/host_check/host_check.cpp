// Test-only shim: compiles the product's __host__ __device__ field/group headers with g++ (carry chains
// emulated) so the algorithm structure can be checked on a CPU-only box. Not part of the product library.
#include <cstring>
#include "../../testudo_b200/csrc/g1_fast.cuh"
#include "../../testudo_b200/csrc/digits.cuh"
#include "../../testudo_b200/csrc/experimental/mont_kara.cuh"
#include "../../testudo_b200/csrc/g2.cuh"
#include "../../testudo_b200/csrc/fq12.cuh"
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
}
