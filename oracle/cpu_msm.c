/* TEST INFRASTRUCTURE ONLY -- C restatement of the CPU algorithm Testudo runs for its G1 MSMs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library. The product (testudo_b200/) never links or calls it.
 *
 * What is restated, and from where (all citations into /root/reference; the arithmetic itself lives
 * in un-vendored crates -- SURVEY.md G2 -- so their *published* algorithms are restated):
 *   - ark-ff 0.4 `Fp384<MontBackend<_, 6>>`: 6 x u64 little-endian limbs, Montgomery form, R = 2^384
 *     (Cargo.toml:21-22,77; constants SURVEY.md App. B).
 *   - ark-ec 0.4 short-Weierstrass `Projective` (Jacobian) with mixed `+= Affine`, a = 0.
 *   - ark-ec 0.4 `VariableBaseMSM::msm_bigint` -> `msm_bigint_wnaf`: window
 *     c = 3 if n < 32 else ln_without_floats(n) + 2, signed radix-2^c digits (`make_digits`), one task
 *     per window (rayon `parallel` feature, Cargo.toml:68) doing bucket accumulation with mixed adds and
 *     a running-sum reduction, then a Horner combine.
 *   - the callers: row fan-out of src/sqrt_pst.rs:121-125 (rows in parallel over a shared SRS),
 *     src/mipp.rs:354-367 `compress` (per-element scalar mul + add + into_affine),
 *     src/commitments.rs:79-86.
 *
 * PARITY UNPINNED: the reference has no golden vectors for this path and cannot be
 * compiled here. This file is cross-checked against oracle/bls12_377.py (big-int definition of the MSM)
 * by tests/test_oracle.py and against curve KATs.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fq;
typedef struct { fq x, y; } g1_affine;      /* identity <=> x == y == 0 (not on the curve) */
typedef struct { fq x, y, z; } g1_jac;      /* identity <=> z == 0 */

static const fq FQ_MOD = {{0x8508c00000000001ULL, 0x170b5d4430000000ULL, 0x1ef3622fba094800ULL,
                           0x1a22d9f300f5138fULL, 0xc63b05c06ca1493bULL, 0x01ae3a4617c510eaULL}};
static const fq FQ_ONE = {{0x02cdffffffffff68ULL, 0x51409f837fffffb1ULL, 0x9f7db3a98a7d3ff2ULL,
                           0x7b4e97b76e7c6305ULL, 0x4cf495bf803c84e8ULL, 0x008d6661e2fdf49aULL}}; /* R mod q */
__attribute__((unused)) static const fq FQ_R2 = {{0xb786686c9400cd22ULL, 0x0329fcaab00431b1ULL, 0x22a5f11162d6b46dULL,
                          0xbfdf7d03827dc3acULL, 0x837e92f041790bf9ULL, 0x006dfccb1e914b88ULL}};
#define FQ_INV 0x8508bfffffffffffULL /* -q^{-1} mod 2^64 */

static const uint64_t FR_MOD[4] = {0x0a11800000000001ULL, 0x59aa76fed0000001ULL, 0x60b44d1e5c37b001ULL,
                                   0x12ab655e9a2ca556ULL};
#define FR_INV 0x0a117fffffffffffULL

/* ------------------------------------------------------------------ Fq */
static inline int fq_is_zero(const fq *a) {
  return (a->l[0] | a->l[1] | a->l[2] | a->l[3] | a->l[4] | a->l[5]) == 0;
}
static inline int fq_eq(const fq *a, const fq *b) {
  uint64_t d = 0;
  for (int i = 0; i < 6; i++) d |= a->l[i] ^ b->l[i];
  return d == 0;
}
static inline int fq_geq_mod(const fq *a) {
  for (int i = 5; i >= 0; i--) {
    if (a->l[i] > FQ_MOD.l[i]) return 1;
    if (a->l[i] < FQ_MOD.l[i]) return 0;
  }
  return 1;
}
static inline void fq_sub_mod_inplace(fq *a) {
  u128 br = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a->l[i] - FQ_MOD.l[i] - (uint64_t)br;
    a->l[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
}
static inline void fq_add(fq *r, const fq *a, const fq *b) {
  u128 c = 0;
  for (int i = 0; i < 6; i++) {
    c += (u128)a->l[i] + b->l[i];
    r->l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (fq_geq_mod(r)) fq_sub_mod_inplace(r);
}
static inline void fq_sub(fq *r, const fq *a, const fq *b) {
  u128 br = 0;
  uint64_t t[6];
  for (int i = 0; i < 6; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)br;
    t[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
  if (br) {
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
      c += (u128)t[i] + FQ_MOD.l[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  memcpy(r->l, t, sizeof t);
}
static inline void fq_neg(fq *r, const fq *a) {
  if (fq_is_zero(a)) { *r = *a; return; }
  fq z = FQ_MOD;
  u128 br = 0;
  for (int i = 0; i < 6; i++) {
    u128 d = (u128)z.l[i] - a->l[i] - (uint64_t)br;
    r->l[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
}
/* CIOS Montgomery multiplication, 6 x 64-bit limbs (the ark-ff `MontBackend::mul_assign` algorithm). */
static void fq_mul(fq *r, const fq *a, const fq *b) {
  uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 6; i++) {
    u128 c = 0;
    for (int j = 0; j < 6; j++) {
      c += (u128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[6];
    t[6] = (uint64_t)c;
    t[7] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * FQ_INV;
    c = (u128)m * FQ_MOD.l[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 6; j++) {
      c += (u128)m * FQ_MOD.l[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[6];
    t[5] = (uint64_t)c;
    t[6] = t[7] + (uint64_t)(c >> 64);
  }
  memcpy(r->l, t, 48);
  if (fq_geq_mod(r)) fq_sub_mod_inplace(r);
}
static inline void fq_sqr(fq *r, const fq *a) { fq_mul(r, a, a); }
static inline void fq_dbl(fq *r, const fq *a) { fq_add(r, a, a); }
static void fq_inv(fq *r, const fq *a) { /* a^(q-2), Fermat */
  uint64_t e[6];
  memcpy(e, FQ_MOD.l, 48);
  e[0] -= 2; /* q is odd and ends ...0001, no borrow */
  fq acc = FQ_ONE, base = *a;
  for (int i = 0; i < 384; i++) {
    if ((e[i / 64] >> (i % 64)) & 1) fq_mul(&acc, &acc, &base);
    fq_sqr(&base, &base);
  }
  *r = acc;
}

/* ------------------------------------------------------------------ G1, Jacobian, a = 0 */
static inline int aff_is_inf(const g1_affine *p) { return fq_is_zero(&p->x) && fq_is_zero(&p->y); }
static inline void jac_set_inf(g1_jac *p) { memset(p, 0, sizeof *p); p->x = FQ_ONE; p->y = FQ_ONE; }
static inline int jac_is_inf(const g1_jac *p) { return fq_is_zero(&p->z); }

static void jac_double(g1_jac *r, const g1_jac *p) { /* dbl-2009-l, as ark-ec double_in_place for a = 0 */
  if (jac_is_inf(p)) { *r = *p; return; }
  fq a, b, c, d, e, f, t;
  fq_sqr(&a, &p->x);
  fq_sqr(&b, &p->y);
  fq_sqr(&c, &b);
  fq_add(&t, &p->x, &b);
  fq_sqr(&t, &t);
  fq_sub(&t, &t, &a);
  fq_sub(&t, &t, &c);
  fq_dbl(&d, &t);
  fq_dbl(&e, &a);
  fq_add(&e, &e, &a);
  fq_sqr(&f, &e);
  fq z3;
  fq_mul(&z3, &p->y, &p->z);
  fq_dbl(&z3, &z3);
  fq x3;
  fq_sub(&x3, &f, &d);
  fq_sub(&x3, &x3, &d);
  fq y3;
  fq_sub(&t, &d, &x3);
  fq_mul(&y3, &e, &t);
  fq_dbl(&c, &c);
  fq_dbl(&c, &c);
  fq_dbl(&c, &c);
  fq_sub(&y3, &y3, &c);
  r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add_mixed(g1_jac *r, const g1_jac *p, const g1_affine *q) { /* madd-2007-bl */
  if (aff_is_inf(q)) { *r = *p; return; }
  if (jac_is_inf(p)) { r->x = q->x; r->y = q->y; r->z = FQ_ONE; return; }
  fq z1z1, u2, s2, h, hh, i, j, rr, v, t;
  fq_sqr(&z1z1, &p->z);
  fq_mul(&u2, &q->x, &z1z1);
  fq_mul(&s2, &q->y, &p->z);
  fq_mul(&s2, &s2, &z1z1);
  if (fq_eq(&u2, &p->x)) {
    if (fq_eq(&s2, &p->y)) { jac_double(r, p); return; }
    jac_set_inf(r);
    return;
  }
  fq_sub(&h, &u2, &p->x);
  fq_sqr(&hh, &h);
  fq_dbl(&i, &hh);
  fq_dbl(&i, &i);
  fq_mul(&j, &h, &i);
  fq_sub(&rr, &s2, &p->y);
  fq_dbl(&rr, &rr);
  fq_mul(&v, &p->x, &i);
  fq x3, y3, z3;
  fq_sqr(&x3, &rr);
  fq_sub(&x3, &x3, &j);
  fq_sub(&x3, &x3, &v);
  fq_sub(&x3, &x3, &v);
  fq_sub(&t, &v, &x3);
  fq_mul(&y3, &rr, &t);
  fq_mul(&t, &p->y, &j);
  fq_dbl(&t, &t);
  fq_sub(&y3, &y3, &t);
  fq_add(&z3, &p->z, &h);
  fq_sqr(&z3, &z3);
  fq_sub(&z3, &z3, &z1z1);
  fq_sub(&z3, &z3, &hh);
  r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add(g1_jac *r, const g1_jac *p, const g1_jac *q) { /* add-2007-bl */
  if (jac_is_inf(p)) { *r = *q; return; }
  if (jac_is_inf(q)) { *r = *p; return; }
  fq z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
  fq_sqr(&z1z1, &p->z);
  fq_sqr(&z2z2, &q->z);
  fq_mul(&u1, &p->x, &z2z2);
  fq_mul(&u2, &q->x, &z1z1);
  fq_mul(&s1, &p->y, &q->z);
  fq_mul(&s1, &s1, &z2z2);
  fq_mul(&s2, &q->y, &p->z);
  fq_mul(&s2, &s2, &z1z1);
  if (fq_eq(&u1, &u2)) {
    if (fq_eq(&s1, &s2)) { jac_double(r, p); return; }
    jac_set_inf(r);
    return;
  }
  fq_sub(&h, &u2, &u1);
  fq_dbl(&i, &h);
  fq_sqr(&i, &i);
  fq_mul(&j, &h, &i);
  fq_sub(&rr, &s2, &s1);
  fq_dbl(&rr, &rr);
  fq_mul(&v, &u1, &i);
  fq x3, y3, z3;
  fq_sqr(&x3, &rr);
  fq_sub(&x3, &x3, &j);
  fq_sub(&x3, &x3, &v);
  fq_sub(&x3, &x3, &v);
  fq_sub(&t, &v, &x3);
  fq_mul(&y3, &rr, &t);
  fq_mul(&t, &s1, &j);
  fq_dbl(&t, &t);
  fq_sub(&y3, &y3, &t);
  fq_add(&z3, &p->z, &q->z);
  fq_sqr(&z3, &z3);
  fq_sub(&z3, &z3, &z1z1);
  fq_sub(&z3, &z3, &z2z2);
  fq_mul(&z3, &z3, &h);
  r->x = x3; r->y = y3; r->z = z3;
}
static void jac_to_affine(g1_affine *r, const g1_jac *p) {
  if (jac_is_inf(p)) { memset(r, 0, sizeof *r); return; }
  fq zi, zi2, zi3;
  fq_inv(&zi, &p->z);
  fq_sqr(&zi2, &zi);
  fq_mul(&zi3, &zi2, &zi);
  fq_mul(&r->x, &p->x, &zi2);
  fq_mul(&r->y, &p->y, &zi3);
}
/* batch normalisation (Montgomery's trick), ark `normalize_batch` */
static void jac_batch_to_affine(g1_affine *out, const g1_jac *in, size_t n) {
  fq *pref = (fq *)malloc((n + 1) * sizeof(fq));
  pref[0] = FQ_ONE;
  for (size_t i = 0; i < n; i++) {
    if (jac_is_inf(&in[i])) pref[i + 1] = pref[i];
    else fq_mul(&pref[i + 1], &pref[i], &in[i].z);
  }
  fq inv;
  fq_inv(&inv, &pref[n]);
  for (size_t k = n; k-- > 0;) {
    if (jac_is_inf(&in[k])) { memset(&out[k], 0, sizeof out[k]); continue; }
    fq zi, zi2, zi3;
    fq_mul(&zi, &inv, &pref[k]);
    fq_mul(&inv, &inv, &in[k].z);
    fq_sqr(&zi2, &zi);
    fq_mul(&zi3, &zi2, &zi);
    fq_mul(&out[k].x, &in[k].x, &zi2);
    fq_mul(&out[k].y, &in[k].y, &zi3);
  }
  free(pref);
}

/* ------------------------------------------------------------------ scalars */
static void fr_from_mont(uint64_t out[4], const uint64_t in[4]) { /* REDC(in) = in * R^-1 mod r */
  uint64_t t[5] = {in[0], in[1], in[2], in[3], 0};
  for (int i = 0; i < 4; i++) {
    uint64_t m = t[0] * FR_INV;
    u128 c = (u128)m * FR_MOD[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * FR_MOD[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = (uint64_t)(c >> 64);
  }
  int ge = 1;
  for (int i = 3; i >= 0; i--) {
    if (t[i] > FR_MOD[i]) { ge = 1; break; }
    if (t[i] < FR_MOD[i]) { ge = 0; break; }
  }
  if (ge) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)t[i] - FR_MOD[i] - (uint64_t)br;
      t[i] = (uint64_t)d;
      br = (d >> 64) & 1;
    }
  }
  memcpy(out, t, 32);
}

static int ark_log2(size_t x) { /* ark_std::log2 = ceil(log2(x)), 0 for x <= 1 */
  if (x <= 1) return 0;
  int n = 0;
  size_t v = x - 1;
  while (v) { n++; v >>= 1; }
  return n;
}
int oracle_ark_window_bits(size_t n) { return n < 32 ? 3 : ark_log2(n) * 69 / 100 + 2; }

/* ark-ec 0.4 `make_digits`: signed radix-2^w digits of a 253-bit scalar (4 x u64, canonical) */
static void make_digits(const uint64_t s[4], int w, int ndig, int32_t *out) {
  uint64_t radix = 1ULL << w, mask = radix - 1, carry = 0;
  for (int i = 0; i < ndig; i++) {
    int bit = i * w, limb = bit / 64, off = bit % 64;
    uint64_t v = limb < 4 ? s[limb] >> off : 0;
    if (off + w > 64 && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
    uint64_t coef = carry + (v & mask);
    carry = (coef + radix / 2) >> w;
    int64_t d = (int64_t)coef - (int64_t)(carry << w);
    if (i == ndig - 1) d += (int64_t)(carry << w);
    out[i] = (int32_t)d;
  }
}

/* ark-ec 0.4 msm_bigint_wnaf restated (App. A.1). scalars canonical 4 x u64. result Jacobian. */
static void msm_wnaf(g1_jac *result, const g1_affine *bases, const uint64_t *scalars, size_t n, int par_windows) {
  int c = oracle_ark_window_bits(n);
  int ndig = (253 + c - 1) / c;
  int32_t *digits = (int32_t *)malloc(n * (size_t)ndig * sizeof(int32_t));
  for (size_t i = 0; i < n; i++) make_digits(scalars + 4 * i, c, ndig, digits + i * ndig);
  g1_jac *wsum = (g1_jac *)malloc(ndig * sizeof(g1_jac));
  size_t nb = (size_t)1 << c; /* ark allocates 1 << c buckets */
#pragma omp parallel for schedule(dynamic, 1) if (par_windows)
  for (int w = 0; w < ndig; w++) {
    g1_jac *buckets = (g1_jac *)malloc(nb * sizeof(g1_jac));
    for (size_t b = 0; b < nb; b++) jac_set_inf(&buckets[b]);
    for (size_t i = 0; i < n; i++) {
      int32_t d = digits[i * ndig + w];
      if (d > 0) {
        jac_add_mixed(&buckets[d - 1], &buckets[d - 1], &bases[i]);
      } else if (d < 0) {
        g1_affine nq = bases[i];
        if (!aff_is_inf(&nq)) fq_neg(&nq.y, &nq.y);
        jac_add_mixed(&buckets[-d - 1], &buckets[-d - 1], &nq);
      }
    }
    g1_jac running, res;
    jac_set_inf(&running);
    jac_set_inf(&res);
    for (size_t b = nb; b-- > 0;) {
      jac_add(&running, &running, &buckets[b]);
      jac_add(&res, &res, &running);
    }
    wsum[w] = res;
    free(buckets);
  }
  g1_jac total;
  jac_set_inf(&total);
  for (int w = ndig - 1; w >= 1; w--) {
    jac_add(&total, &total, &wsum[w]);
    for (int k = 0; k < c; k++) jac_double(&total, &total);
  }
  jac_add(&total, &total, &wsum[0]);
  *result = total;
  free(wsum);
  free(digits);
}

/* ------------------------------------------------------------------ exported (ctypes) API
 * Layouts are the C-ABI ones of include/testudo_b200.h: points = x[6] || y[6] u64 Montgomery LE,
 * all-zero == identity; scalars = 4 x u64, canonical unless `mont` is set. */
static void load_scalars(uint64_t *dst, const uint64_t *src, size_t n, int mont) {
  if (!mont) { memcpy(dst, src, n * 32); return; }
  for (size_t i = 0; i < n; i++) fr_from_mont(dst + 4 * i, src + 4 * i);
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* VariableBaseMSM::msm_bigint + into_affine (one MSM; windows in parallel like ark's `parallel`) */
void oracle_msm_g1(const uint64_t *bases_xy, const uint64_t *scalars, size_t n, int scalars_mont,
                   uint64_t out_xy[12]) {
  uint64_t *sc = (uint64_t *)malloc((n ? n : 1) * 32);
  load_scalars(sc, scalars, n, scalars_mont);
  g1_jac r;
  if (n == 0) jac_set_inf(&r);
  else msm_wnaf(&r, (const g1_affine *)bases_xy, sc, n, 1);
  g1_affine a;
  jac_to_affine(&a, &r);
  memcpy(out_xy, &a, 96);
  free(sc);
}

/* Row fan-out of src/sqrt_pst.rs:121-125 / src/dense_mlpoly.rs:315-329: `rows` MSMs over shared bases.
 * scalar (i, j) of row i lives at scalars[4 * (i*row_stride + j*col_stride)]. Rows in parallel (rayon
 * par_iter over rows; nested window parallelism adds nothing once rows >= cores). */
void oracle_msm_g1_batch(const uint64_t *bases_xy, const uint64_t *scalars, size_t rows, size_t cols,
                         ptrdiff_t row_stride, ptrdiff_t col_stride, int scalars_mont, uint64_t *out_xy) {
  g1_jac *res = (g1_jac *)malloc((rows ? rows : 1) * sizeof(g1_jac));
#pragma omp parallel for schedule(dynamic, 1)
  for (ptrdiff_t i = 0; i < (ptrdiff_t)rows; i++) {
    uint64_t *sc = (uint64_t *)malloc((cols ? cols : 1) * 32);
    for (size_t j = 0; j < cols; j++) {
      const uint64_t *s = scalars + 4 * (i * row_stride + (ptrdiff_t)j * col_stride);
      if (scalars_mont) fr_from_mont(sc + 4 * j, s);
      else memcpy(sc + 4 * j, s, 32);
    }
    if (cols == 0) jac_set_inf(&res[i]);
    else msm_wnaf(&res[i], (const g1_affine *)bases_xy, sc, cols, 0);
    free(sc);
  }
  jac_batch_to_affine((g1_affine *)out_xy, res, rows);
  free(res);
}

/* k * P by left-to-right double-and-add (ark `mul_bigint`), into_affine */
static void scalar_mul(g1_jac *r, const g1_affine *p, const uint64_t k[4]) {
  g1_jac acc;
  jac_set_inf(&acc);
  for (int i = 255; i >= 0; i--) {
    jac_double(&acc, &acc);
    if ((k[i / 64] >> (i % 64)) & 1) jac_add_mixed(&acc, &acc, p);
  }
  *r = acc;
}
void oracle_g1_mul(const uint64_t p_xy[12], const uint64_t k[4], uint64_t out_xy[12]) {
  g1_jac r;
  scalar_mul(&r, (const g1_affine *)p_xy, k);
  g1_affine a;
  jac_to_affine(&a, &r);
  memcpy(out_xy, &a, 96);
}
void oracle_g1_add(const uint64_t p_xy[12], const uint64_t q_xy[12], uint64_t out_xy[12]) {
  g1_jac r;
  const g1_affine *p = (const g1_affine *)p_xy;
  if (aff_is_inf(p)) jac_set_inf(&r);
  else { r.x = p->x; r.y = p->y; r.z = FQ_ONE; }
  jac_add_mixed(&r, &r, (const g1_affine *)q_xy);
  g1_affine a;
  jac_to_affine(&a, &r);
  memcpy(out_xy, &a, 96);
}
/* src/mipp.rs:354-367 `compress`: left[i] = (right[i] * scaler + left[i]).into_affine(), i < split */
void oracle_compress_g1(uint64_t *vec_xy, size_t split, const uint64_t scaler[4], int scaler_mont) {
  uint64_t k[4];
  load_scalars(k, scaler, 1, scaler_mont);
  g1_affine *v = (g1_affine *)vec_xy;
  g1_jac *tmp = (g1_jac *)malloc((split ? split : 1) * sizeof(g1_jac));
#pragma omp parallel for schedule(static)
  for (ptrdiff_t i = 0; i < (ptrdiff_t)split; i++) {
    scalar_mul(&tmp[i], &v[split + i], k);
    jac_add_mixed(&tmp[i], &tmp[i], &v[i]);
  }
  jac_batch_to_affine(v, tmp, split);
  free(tmp);
}
/* Fq helpers for unit-testing the device field code (Montgomery in, Montgomery out) */
void oracle_fq_mul(const uint64_t a[6], const uint64_t b[6], uint64_t out[6]) {
  fq r;
  fq_mul(&r, (const fq *)a, (const fq *)b);
  memcpy(out, &r, 48);
}
void oracle_fq_add(const uint64_t a[6], const uint64_t b[6], uint64_t out[6]) {
  fq r;
  fq_add(&r, (const fq *)a, (const fq *)b);
  memcpy(out, &r, 48);
}
void oracle_fq_sub(const uint64_t a[6], const uint64_t b[6], uint64_t out[6]) {
  fq r;
  fq_sub(&r, (const fq *)a, (const fq *)b);
  memcpy(out, &r, 48);
}
void oracle_fq_inv(const uint64_t a[6], uint64_t out[6]) {
  fq r;
  fq_inv(&r, (const fq *)a);
  memcpy(out, &r, 48);
}
void oracle_fr_from_mont(const uint64_t a[4], uint64_t out[4]) { fr_from_mont(out, a); }

/* Synthetic bases with known discrete logs: P_k = start + k * step, affine, via
 * chunked Jacobian walk + batch normalisation. start/step are affine points supplied by the caller. */
void oracle_gen_points(const uint64_t start_xy[12], const uint64_t step_xy[12], size_t n, uint64_t *out_xy) {
  const size_t CH = 4096;
  g1_jac cur;
  const g1_affine *st = (const g1_affine *)start_xy;
  cur.x = st->x; cur.y = st->y; cur.z = FQ_ONE;
  g1_jac *buf = (g1_jac *)malloc(CH * sizeof(g1_jac));
  for (size_t base = 0; base < n; base += CH) {
    size_t m = n - base < CH ? n - base : CH;
    for (size_t k = 0; k < m; k++) {
      buf[k] = cur;
      jac_add_mixed(&cur, &cur, (const g1_affine *)step_xy);
    }
    jac_batch_to_affine((g1_affine *)(out_xy + 12 * base), buf, m);
  }
  free(buf);
}
