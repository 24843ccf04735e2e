// Lane-parallel Fq12 arithmetic with LAZY field reduction: the per-item bodies of the cooperative pairing engine.
//
// kernels_pairing.cuh runs every Fq12 operation of a final exponentiation / a Miller loop on a team of lanes: the 54
// (36, 18) independent Fq products of a Karatsuba Fq12 product, one per lane, and the linear recombinations one Fq
// coefficient per lane, phases separated by a team barrier. ncu (profiles/r01_summary.md 5b) showed ~1 600 instructions
// on the critical lane of one Fq12 product of which only ~400 belong to the Montgomery product: the canonical linear
// operations (add + conditional subtraction 38 instructions, 5 a = three of those, -5 a = four) were the chain.
//
// Here nothing between two products is reduced to [0, q). Fq is 377 bits in a 384-bit container: 2^384 = 152.3 q, so
//   * lz_add / lz_sub / lz_mul5 / lz_dbl are plain 384-bit operations (12, 12, 24, 12 instructions), wrapping mod 2^384;
//   * a chain with subtractions is made non-negative by ONE constant k q added per lane and phase (lz_add_kq<k>, k q an
//     immediate): the TRUE value of every stored quantity lies in [0, 2^384), so the wrapped arithmetic is exact;
//   * the multiplier is mont_mul_lazy without its conditional subtraction: a < A q, b < B q  ->  < (1 + A B / 152.3) q;
//   * lz_reduce subtracts floor(top limb / (top limb of q + 1)) q: any value < 2^384 -> < (1 + 2^-16) q, ~40 instructions.
//     It runs twice per Fq12 operation (after the Fq6 recombination and on the 12 output coefficients).
// Every function below documents the bound (in multiples of q) of what it reads and writes; tests/test_coop_lazy.py runs
// these same bodies on the host (tests/host_check) over extreme inputs with an overflow detector, against the oracle.
//
// INVARIANT of an Fq12 value in shared memory between operations: every coefficient < 1.02 q (RHO). Values that leave a
// kernel are canonicalised (lz_canon).
//
// The bodies are __host__ __device__ and take the item index t: the device drivers (kernels_pairing.cuh) call them as
//   for (t = lane; t < items; t += team) body(..., t);   barrier;
// the host check calls them for t = 0 .. items-1 in sequence -- equivalent because a phase only reads what earlier
// phases wrote (the in-place phases read into registers, pass a barrier, then write).
#pragma once
#include "fq12.cuh"

namespace tb {

// ---- lazily reduced Fq operations ------------------------------------------------------------------------------------------
struct KQ {
  uint32_t l[12];
};
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr KQ make_kq(unsigned k) {
  constexpr uint32_t Q[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u, 0xba094800u, 0x1ef3622fu,
                              0x00f5138fu, 0x1a22d9f3u, 0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
  KQ r{};
  uint64_t c = 0;
  for (int j = 0; j < 12; j++) {
    const uint64_t t = (uint64_t)Q[j] * k + c;
    r.l[j] = (uint32_t)t;
    c = t >> 32;
  }
  return r;
}
TB_HD void lz_add(Fq& r, const Fq& a, const Fq& b) {
  Carry c;
  r.l[0] = add_cc(a.l[0], b.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = addc_cc(a.l[i], b.l[i], c);
}
TB_HD void lz_sub(Fq& r, const Fq& a, const Fq& b) {   // wraps mod 2^384
  Carry c;
  r.l[0] = sub_cc(a.l[0], b.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = subc_cc(a.l[i], b.l[i], c);
}
TB_HD void lz_dbl(Fq& r, const Fq& a) { lz_add(r, a, a); }
TB_HD void lz_mul5(Fq& r, const Fq& a) {   // 4 a (funnel shifts: no carry chain) + a
  Fq t;
#pragma unroll
  for (int i = 11; i >= 1; i--) t.l[i] = (a.l[i] << 2) | (a.l[i - 1] >> 30);
  t.l[0] = a.l[0] << 2;
  lz_add(r, t, a);
}
TB_HD void lz_mul3(Fq& r, const Fq& a) {
  Fq t;
  lz_dbl(t, a);
  lz_add(r, t, a);
}
template <unsigned K>
TB_HD void lz_add_kq(Fq& r) {
  constexpr KQ kq = make_kq(K);
  Carry c;
  r.l[0] = add_cc(r.l[0], kq.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = addc_cc(r.l[i], kq.l[i], c);
}
// r = k q - a (a <= k q)
template <unsigned K>
TB_HD void lz_neg_kq(Fq& r, const Fq& a) {
  constexpr KQ kq = make_kq(K);
  Carry c;
  r.l[0] = sub_cc(kq.l[0], a.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = subc_cc(kq.l[i], a.l[i], c);
}
// any a < 2^384  ->  a - floor(a_11 / (q_11 + 1)) q  <  (1 + 2^-16) q:  with D = q_11 + 1 the quotient digit never exceeds
// a / q, and  a - t q < q (1 - 1/D) + 2^352 + a_11 2^352 / D.
TB_HD void lz_reduce(Fq& a) {
  const uint32_t t = a.l[11] / 0x01ae3a47u;   // <= 152; a constant divisor: multiply-high + shift
  uint32_t y[12];
#pragma unroll
  for (int i = 0; i < 12; i++) y[i] = mul_lo(t, FqParams::p(i));
  Carry c;
  y[1] = add_cc(y[1], mul_hi(t, FqParams::p(0)), c);
#pragma unroll
  for (int i = 2; i < 12; i++) y[i] = addc_cc(y[i], mul_hi(t, FqParams::p(i - 1)), c);
  Carry d;
  a.l[0] = sub_cc(a.l[0], y[0], d);
#pragma unroll
  for (int i = 1; i < 12; i++) a.l[i] = subc_cc(a.l[i], y[i], d);
}
TB_HD void lz_canon(Fq& a) {   // any a < 2^384 -> [0, q)
  lz_reduce(a);
  mod_reduce_once<FqParams>(a.l);
}
// (a + (a odd ? q : 0)) / 2 for a + q < 2^384: < (a + q) / 2
TB_HD void lz_halve(Fq& r, const Fq& a) { fq_halve(r, a); }

// the one out-of-line multiplier of the cooperative engine; operands and result in registers.
// a < A q, b < B q (A, B <= 150)  ->  r < (1 + A B / 152.3) q
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ Fq fq_mul_lz(Fq a, Fq b) {
  Fq r;
  mont_mul_lazy<FqParams>(r.l, a.l, b.l);
  return r;
}
#else
inline Fq fq_mul_lz(Fq a, Fq b) {
  Fq r;
  mont_mul_lazy<FqParams>(r.l, a.l, b.l);
  return r;
}
#endif

// ---- shared-memory scratch of one team -----------------------------------------------------------------------------------
struct WScratch {
  Fq xy[2][3][3][2];   // materialised Fq6 operands X_i, Y_i (i < 3) of up to three Fq6 products, [t][c] = Fq2 coefficient t, part c
  Fq kar[54];          // [product 0..17][a0 b0, a1 b1, (a0 + a1)(b0 + b1)]
  Fq prod[18][2];      // Fq2 products: [i][j], j = 0..5 -> x0y0, x1y1, x2y2, (x1+x2)(y1+y2), (x0+x1)(y0+y1), (x0+x2)(y0+y2)
  Fq r6[3][3][2];      // the three Fq6 results
};
TB_HD Fq* w12_q(Fq12* a) { return reinterpret_cast<Fq*>(a); }              // [2 * slot + c]
TB_HD const Fq* w12_q(const Fq12* a) { return reinterpret_cast<const Fq*>(a); }
TB_HD Fq2* w12_c(Fq12* a, int idx) { return reinterpret_cast<Fq2*>(a) + idx; }
TB_HD const Fq2* w12_c(const Fq12* a, int idx) { return reinterpret_cast<const Fq2*>(a) + idx; }

// lane `part` of the Karatsuba split of the Fq2 value v (or v0 + v1): c0, c1 or c0 + c1. Bound: the sum of the bounds.
TB_HD void wl_kar_operand(Fq& o, const Fq* v0, const Fq* v1, bool two, int part) {
  if (part < 2) {
    o = v0[part];
    if (two) lz_add(o, o, v1[part]);
  } else {
    lz_add(o, v0[0], v0[1]);
    if (two) {
      lz_add(o, o, v1[0]);
      lz_add(o, o, v1[1]);
    }
  }
}
// kar[0..2] (each < K q) -> coefficient c of the Fq2 product:  c0 = v0 - 5 v1 + O0 q,  c1 = m - v0 - v1 + O1 q.
// Needs O0 >= 5 K, O1 >= 2 K; result < (K + O0) q resp. (K + O1) q.
template <unsigned O0, unsigned O1>
TB_HD void wl_fq2_from_kar(Fq& o, const Fq* kar, int c) {
  if (c == 0) {
    Fq t;
    lz_mul5(t, kar[1]);
    lz_sub(o, kar[0], t);
    lz_add_kq<O0>(o);
  } else {
    lz_sub(o, kar[2], kar[0]);
    lz_sub(o, o, kar[1]);
    lz_add_kq<O1>(o);
  }
}

// ---- the three phases shared by every Fq6-product based operation ---------------------------------------------------------
// Operand bounds (checked per caller): X side <= 8.2 q, Y side <= 13.2 q per lane  =>  kar < KMAX q = 1.75 q.
// phase 1, item t < 18 * count: one Fq product
TB_HD void wp_kar(WScratch* w, int t) {
  const int pr = t / 3, part = t % 3, i = pr / 6, j = pr % 6;
  const int t0 = (j < 3) ? j : (j == 3 ? 1 : 0);
  const int t1 = (j == 4) ? 1 : 2;
  Fq a, b;
  wl_kar_operand(a, w->xy[0][i][t0], w->xy[0][i][t1], j >= 3, part);
  wl_kar_operand(b, w->xy[1][i][t0], w->xy[1][i][t1], j >= 3, part);
  w->kar[t] = fq_mul_lz(a, b);
}
// phase 2, item t < 12 * count: coefficient t & 1 of Fq2 product t / 2.  prod[.][0] < 10.75 q, prod[.][1] < 5.75 q
TB_HD void wp_fq2(WScratch* w, int t) {
  Fq o;
  wl_fq2_from_kar<9, 4>(o, &w->kar[3 * (t / 2)], t & 1);
  w->prod[t / 2][t & 1] = o;
}
// phase 3, item t < 6 * count: coefficient (tt, c) of Fq6 result i, REDUCED (< 1.00002 q).
// With P0 = 10.75, P1 = 5.75 (bounds of prod[.][0], prod[.][1]):
//   tt = 0: v0 + xi (m12 - v1 - v2):  c0 = p0.0 - 5 (p3.1 - p1.1 - p2.1) + 29 q < 97.3 q;  c1 = p0.1 + p3.0 - p1.0 - p2.0 + 22 q < 38.5 q
//   tt = 1: m01 - v0 - v1 + xi v2:    c0 = p4.0 - p0.0 - p1.0 - 5 p2.1 + 51 q < 61.8 q;     c1 = p4.1 - p0.1 - p1.1 + p2.0 + 12 q < 28.5 q
//   tt = 2: m02 - v0 - v2 + v1:       c0 = ... + 22 q < 43.5 q;                             c1 = ... + 12 q < 23.5 q
TB_HD void wp_fq6(WScratch* w, int t) {
  const int i = t / 6, tt = (t % 6) / 2, c = t & 1;
  const Fq(*p)[2] = &w->prod[6 * i];
  Fq r, m;
  if (tt == 0) {
    lz_sub(m, p[3][1 - c], p[1][1 - c]);
    lz_sub(m, m, p[2][1 - c]);
    if (c == 0) {
      lz_mul5(m, m);
      lz_sub(r, p[0][0], m);
      lz_add_kq<29>(r);
    } else {
      lz_add(r, p[0][1], m);
      lz_add_kq<22>(r);
    }
  } else if (tt == 1) {
    lz_sub(r, p[4][c], p[0][c]);
    lz_sub(r, r, p[1][c]);
    if (c == 0) {
      lz_mul5(m, p[2][1]);
      lz_sub(r, r, m);
      lz_add_kq<51>(r);
    } else {
      lz_add(r, r, p[2][0]);
      lz_add_kq<12>(r);
    }
  } else {
    lz_sub(r, p[5][c], p[0][c]);
    lz_sub(r, r, p[2][c]);
    lz_add(r, r, p[1][c]);
    if (c == 0) lz_add_kq<22>(r);
    else lz_add_kq<12>(r);
  }
  lz_reduce(r);
  w->r6[i][tt][c] = r;
}

// ---- dst = a * b -------------------------------------------------------------------------------------------------------------
// X = {a0, a1, a0 + a1}, Y = {b0, b1, b0 + b1}; C0 = R0 + v R1, C1 = R2 - R0 - R1.
// item t < 36: xy coefficient; inputs < RHO q = 1.02 q -> xy < 2.04 q -> product operands < 8.16 q (A B < 67)
TB_HD void wp_mul_xy(WScratch* w, const Fq12* a, const Fq12* b, int t) {
  const int which = t / 18, rem = t % 18, i = rem / 6, tt = (rem % 6) / 2, c = rem & 1;
  const Fq* src = w12_q(which ? b : a);
  Fq v;
  if (i < 2) v = src[2 * (3 * i + tt) + c];
  else lz_add(v, src[2 * tt + c], src[2 * (3 + tt) + c]);
  w->xy[which][i][tt][c] = v;
}
// item t < 12: output coefficient, reduced. r6 < 1.00002 q:  slot 0 c 0: R0 - 5 R1.2.1 + 6 q < 7.1 q; C1: R2 - R0 - R1 + 3 q < 4.1 q
TB_HD void wp_mul_out(Fq12* dst, const WScratch* w, int t) {
  const int slot = t / 2, c = t & 1;
  Fq r, m;
  if (slot == 0 && c == 0) {
    lz_mul5(m, w->r6[1][2][1]);
    lz_sub(r, w->r6[0][0][0], m);
    lz_add_kq<6>(r);
  } else if (slot < 3) {
    lz_add(r, w->r6[0][slot][c], slot == 0 ? w->r6[1][2][0] : w->r6[1][slot - 1][c]);
  } else {
    lz_sub(r, w->r6[2][slot - 3][c], w->r6[0][slot - 3][c]);
    lz_sub(r, r, w->r6[1][slot - 3][c]);
    lz_add_kq<3>(r);
  }
  lz_reduce(r);
  w12_q(dst)[t] = r;
}

// ---- dst = a^2 (complex squaring) -------------------------------------------------------------------------------------------
// R0 = a0 a1, R1 = (a0 + a1)(a0 + v a1); C0 = R1 - R0 - v R0, C1 = 2 R0.
// item t < 24. X0 = a0, Y0 = a1 (< 1.02 q), X1 = a0 + a1 (< 2.04 q), Y1 = a0 + v a1: coefficient (0, 0) = a0.0.0 - 5 a1.2.1
// + 6 q < 7.02 q, the others < 2.04 q  =>  X-side operands < 8.16 q, Y-side < 13.14 q, A B < 107.3, kar < 1.705 q
TB_HD void wp_sqr_xy(WScratch* w, const Fq12* a, int t) {
  const int which = t / 12, rem = t % 12, i = rem / 6, tt = (rem % 6) / 2, c = rem & 1;
  const Fq* src = w12_q(a);
  Fq v;
  if (i == 0) v = src[2 * (3 * which + tt) + c];
  else if (which == 0) lz_add(v, src[2 * tt + c], src[2 * (3 + tt) + c]);
  else if (tt == 0 && c == 0) {
    Fq m;
    lz_mul5(m, src[6 + 2 * 2 + 1]);
    lz_sub(v, src[0], m);
    lz_add_kq<6>(v);
  } else {
    lz_add(v, src[2 * tt + c], tt == 0 ? src[6 + 2 * 2 + 0] : src[6 + 2 * (tt - 1) + c]);
  }
  w->xy[which][i][tt][c] = v;
}
// item t < 12, reduced. slot s < 3: R1.s - R0.s - (v R0).s: s = 0, c = 0: R1 - R0 + 5 R0.2.1 + 2 q < 8.1 q; otherwise + 3 q < 4.1 q;
// slot >= 3: 2 R0 < 2.1 q
TB_HD void wp_sqr_out(Fq12* dst, const WScratch* w, int t) {
  const int slot = t / 2, c = t & 1;
  Fq r, m;
  if (slot == 0 && c == 0) {
    lz_mul5(m, w->r6[0][2][1]);
    lz_sub(r, w->r6[1][0][0], w->r6[0][0][0]);
    lz_add(r, r, m);
    lz_add_kq<2>(r);
  } else if (slot < 3) {
    lz_sub(r, w->r6[1][slot][c], w->r6[0][slot][c]);
    lz_sub(r, r, slot == 0 ? w->r6[0][2][0] : w->r6[0][slot - 1][c]);
    lz_add_kq<3>(r);
  } else {
    lz_dbl(r, w->r6[0][slot - 3][c]);
  }
  lz_reduce(r);
  w12_q(dst)[t] = r;
}

// ---- dst = a^2 for a unitary a (Granger-Scott) ----------------------------------------------------------------------------------
// tower slot of z_k: z0 = c[0], z1 = c[4], z2 = c[3], z3 = c[2], z4 = c[1], z5 = c[5]. Six Fq2 products = 18 Fq products:
// pair p = (za, zb) = (z_{2p}, z_{2p+1}); even product: za zb, odd product: (za + zb)(za + xi zb).
TB_HD int w12_cyc_slot(int k) { return k == 0 ? 0 : k == 1 ? 4 : k == 2 ? 3 : k == 3 ? 2 : k == 4 ? 1 : 5; }
// item t < 18. Inputs < 1.02 q: operands < 4.08 q and < 9.06 q (B.0 = za.0 - 5 zb.1 + 6 q < 7.02 q), A B < 37: kar < 1.25 q
TB_HD void wp_cyc_kar(WScratch* w, const Fq12* a, int t) {
  const int l = t / 3, part = t % 3, p = l >> 1;
  const Fq* za = w12_q(a) + 2 * w12_cyc_slot(2 * p);
  const Fq* zb = w12_q(a) + 2 * w12_cyc_slot(2 * p + 1);
  Fq x, y;
  if ((l & 1) == 0) {
    wl_kar_operand(x, za, za, false, part);
    wl_kar_operand(y, zb, zb, false, part);
  } else {
    wl_kar_operand(x, za, zb, true, part);
    Fq B[2], m;
    lz_mul5(m, zb[1]);
    lz_sub(B[0], za[0], m);
    lz_add_kq<6>(B[0]);
    lz_add(B[1], za[1], zb[0]);
    wl_kar_operand(y, B, B, false, part);
  }
  w->kar[t] = fq_mul_lz(x, y);
}
// item t < 12: prod[t / 2][t & 1]; prod[.][0] < 8.25 q, prod[.][1] < 4.25 q
TB_HD void wp_cyc_fq2(WScratch* w, int t) {
  Fq o;
  wl_fq2_from_kar<7, 3>(o, &w->kar[3 * (t / 2)], t & 1);
  w->prod[t / 2][t & 1] = o;
}
// item t < 12: z_k' = 3 t_k -/+ 2 z_k with t = t0, t1, xi t5, t4, t2, t3 for k = 0..5, where t_{2p} = prod[2p+1] - tmp - xi tmp and
// t_{2p+1} = 2 tmp (tmp = prod[2p]). With P0 = 8.25, P1 = 4.25:
//   even, c0: P.0 - tmp.0 + 5 tmp.1 + 9 q < 38.5 q;  c1: P.1 - tmp.1 - tmp.0 + 13 q < 17.3 q;  odd: < 16.5 q / 8.5 q
//   k = 2: c0 = -5 (2 tmp.1) + 43 q < 43 q;  c1 = 2 tmp.0 < 16.5 q
//   output 3 t + 2 z < 131.1 q (k = 2) or 3 t - 2 z + 3 q < 118.5 q, reduced
// Each item rewrites only the coefficient it read: dst may alias a.
TB_HD void wp_cyc_out(Fq12* dst, const Fq12* a, const WScratch* w, int t) {
  const int k = t / 2, c = t & 1;
  const int ti = k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 5 : k == 3 ? 4 : k == 4 ? 2 : 3;
  const int p = ti >> 1;
  Fq tv;
  if (k == 2) {            // xi t5, t5 = 2 prod[4]
    if (c == 0) {
      Fq u;
      lz_dbl(u, w->prod[4][1]);
      lz_mul5(u, u);
      lz_neg_kq<43>(tv, u);
    } else {
      lz_dbl(tv, w->prod[4][0]);
    }
  } else if ((ti & 1) == 0) {
    if (c == 0) {
      Fq m;
      lz_mul5(m, w->prod[2 * p][1]);
      lz_sub(tv, w->prod[2 * p + 1][0], w->prod[2 * p][0]);
      lz_add(tv, tv, m);
      lz_add_kq<9>(tv);
    } else {
      lz_sub(tv, w->prod[2 * p + 1][1], w->prod[2 * p][1]);
      lz_sub(tv, tv, w->prod[2 * p][0]);
      lz_add_kq<13>(tv);
    }
  } else {
    lz_dbl(tv, w->prod[2 * p][c]);
  }
  const int slot = w12_cyc_slot(k);
  const Fq z = w12_q(a)[2 * slot + c];
  Fq o;
  if (k == 0 || k == 3 || k == 4) {
    lz_sub(o, tv, z);
    lz_dbl(o, o);
    lz_add(o, o, tv);
    lz_add_kq<3>(o);
  } else {
    lz_add(o, tv, z);
    lz_dbl(o, o);
    lz_add(o, o, tv);
  }
  lz_reduce(o);
  w12_q(dst)[2 * slot + c] = o;
}

// ---- unary operations, item t < 12: value computed here, stored by the caller after a barrier ----------------------------------
// conjugation: coefficients 6..11 negated (2 q - v, one conditional subtraction: <= q)
TB_HD Fq wp_conj(const Fq12* a, int t) {
  Fq v = w12_q(a)[t];
  if (t >= 6) {
    lz_neg_kq<2>(v, v);
    mod_reduce_once<FqParams>(v.l);
  }
  return v;
}
// a^(q^k), k = 1, 2: tower slot idx holds the coefficient of w^e, e = 2 idx (idx < 3) or 2 (idx - 3) + 1; every Frobenius
// coefficient u^(e (q^k - 1)/6) lies in Fq (tests/test_oracle_pairing.py): 12 independent Fq products. Output < 1.02 q.
TB_HD Fq wp_frobenius(const Fq12* a, int k, int t) {
  const int sl = t / 2;
  const int e = sl < 3 ? 2 * sl : 2 * (sl - 3) + 1;
  Fq c = w12_q(a)[t];
  if (k == 1 && (t & 1)) lz_neg_kq<2>(c, c);      // < 2 q
  if (e > 0) {
    const Fq g = fq_from_table(k == 1 ? FQ12_C(FROB1)[e - 1][0] : FQ12_C(FROB2)[e - 1]);
    c = fq_mul_lz(c, g);                             // < (1 + 2 / 152) q
  } else if (k == 1 && (t & 1)) {
    mod_reduce_once<FqParams>(c.l);
  }
  return c;
}

// ---- doubling step of the Miller loop (ark `double_in_place`, see g2_double_line) on parallel lanes -------------------------------
// r = (x, y, z) canonical on entry and on exit; px, py canonical. Products in two rounds (11 and 14 lanes) plus the twist
// coefficient; squarings of v = v0 + v1 u use (v0 + v1)(v0 - 5 v1) = s0 and v0 v1 = s1:  v^2 = (s0 + 4 s1) + 2 s1 u.
struct WDouble {
  G2Hom r;
  Fq px, py;
  Fq kar[16];        // Fq products of the current round
  Fq v[11][2];       // Fq2 intermediates
};
enum { WD_A = 0, WD_B, WD_C, WD_YZ, WD_J, WD_E, WD_D, WD_G, WD_H, WD_NH, WD_J3 };
// operands of the squaring products: which = 0 -> (v0 + v1, v0 - 5 v1 + OFF q), 1 -> (v0, v1)
template <unsigned OFF>
TB_HD void wl_sqr_operands(Fq& a, Fq& b, const Fq* v, int which) {
  if (which == 0) {
    Fq t;
    lz_add(a, v[0], v[1]);
    lz_mul5(t, v[1]);
    lz_sub(b, v[0], t);
    lz_add_kq<OFF>(b);
  } else {
    a = v[0];
    b = v[1];
  }
}
TB_HD void wl_sqr_asm(Fq& o, const Fq& s0, const Fq& s1, int c) {   // (s0 + 4 s1, 2 s1)
  Fq t;
  lz_dbl(t, s1);
  if (c == 0) {
    lz_dbl(t, t);
    lz_add(o, s0, t);
  } else {
    o = t;
  }
}
// round 1, item t < 11: 0-2 x y (Karatsuba), 3-4 y^2, 5-6 z^2, 7-8 (y+z)^2, 9-10 x^2.
// operands < 4 q and < 12 q (v0 - 5 v1 + 10 q with v < 2 q): kar < 1.32 q
TB_HD void wp_dbl_r1(WDouble* s, int t) {
  const Fq* rx = reinterpret_cast<const Fq*>(&s->r.x);
  const Fq* ry = reinterpret_cast<const Fq*>(&s->r.y);
  const Fq* rz = reinterpret_cast<const Fq*>(&s->r.z);
  Fq a, b;
  if (t < 3) {
    wl_kar_operand(a, rx, rx, false, t);
    wl_kar_operand(b, ry, ry, false, t);
  } else {
    const int q = (t - 3) >> 1;
    Fq w[2];
    const Fq* src = q == 1 ? rz : (q == 3 ? rx : ry);
    w[0] = src[0];
    w[1] = src[1];
    if (q == 2) {
      lz_add(w[0], w[0], rz[0]);
      lz_add(w[1], w[1], rz[1]);
    }
    wl_sqr_operands<10>(a, b, w, (t - 3) & 1);
  }
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 12: a = x y / 2 (< 4.7 q, 2.7 q); b, c, (y+z)^2, j (< 6.6 q, 2.64 q); e = B' 3 c with B' = (0, b1), b1 = -1/5:
// e.0 = 3 c.1 (< 7.92 q), e.1 = b1 * 3 c.0 (operand < 19.8 q: < 1.14 q)
TB_HD void wp_dbl_p2(WDouble* s, int t) {
  const int c = t & 1;
  Fq o;
  if (t < 2) {
    wl_fq2_from_kar<7, 3>(o, &s->kar[0], c);
    lz_halve(o, o);
    s->v[WD_A][c] = o;
  } else if (t < 10) {
    const int q = (t - 2) >> 1;
    wl_sqr_asm(o, s->kar[3 + 2 * q], s->kar[4 + 2 * q], c);
    s->v[WD_B + q][c] = o;
  } else {
    Fq cc, tt;
    wl_sqr_asm(cc, s->kar[5], s->kar[6], 1 - c);
    lz_mul3(tt, cc);
    if (c == 0) o = tt;
    else o = fq_mul_lz(tt, fq_from_table(FQ12_C(TWIST_B1)));
    s->v[WD_E][c] = o;
  }
}
// item t < 10: d = b - 3 e + 24 q (< 30.6 q), g = (b + 3 e) / 2 (< 15.7 q), h = (y+z)^2 - b - c + 14 q (< 20.6 q),
// -h = b + c - (y+z)^2 + 7 q (< 20.2 q), 3 j (< 19.8 q)
TB_HD void wp_dbl_p3(WDouble* s, int t) {
  const int c = t & 1, q = t >> 1;
  Fq(*v)[2] = s->v;
  Fq o, tt;
  if (q < 2) {
    lz_mul3(tt, v[WD_E][c]);
    if (q == 0) {
      lz_sub(o, v[WD_B][c], tt);
      lz_add_kq<24>(o);
    } else {
      lz_add(o, v[WD_B][c], tt);
      lz_halve(o, o);
    }
    v[WD_D + q][c] = o;
  } else if (q == 2) {
    lz_sub(o, v[WD_YZ][c], v[WD_B][c]);
    lz_sub(o, o, v[WD_C][c]);
    lz_add_kq<14>(o);
    v[WD_H][c] = o;
  } else if (q == 3) {
    lz_add(o, v[WD_B][c], v[WD_C][c]);
    lz_sub(o, o, v[WD_YZ][c]);
    lz_add_kq<7>(o);
    v[WD_NH][c] = o;
  } else {
    lz_mul3(o, v[WD_J][c]);
    v[WD_J3][c] = o;
  }
}
// round 2, item t < 14: 0-2 a d, 3-4 g^2, 5-6 e^2, 7-9 b h, 10-11 (-h) py, 12-13 (3 j) px.
// a d: 7.4 q x 57.3 q -> < 3.8 q; g^2: 19.3 q x 33.7 q -> < 5.3 q (v0 - 5 v1 + 18 q); b h: 9.3 q x 37.3 q -> < 3.3 q
TB_HD void wp_dbl_r2(WDouble* s, int t) {
  Fq(*v)[2] = s->v;
  Fq a, b;
  if (t < 3) {
    wl_kar_operand(a, v[WD_A], v[WD_A], false, t);
    wl_kar_operand(b, v[WD_D], v[WD_D], false, t);
  } else if (t < 7) {
    wl_sqr_operands<18>(a, b, t < 5 ? v[WD_G] : v[WD_E], (t - 3) & 1);
  } else if (t < 10) {
    wl_kar_operand(a, v[WD_B], v[WD_B], false, t - 7);
    wl_kar_operand(b, v[WD_H], v[WD_H], false, t - 7);
  } else if (t < 12) {
    a = v[WD_NH][t - 10];
    b = s->py;
  } else {
    a = v[WD_J3][t - 12];
    b = s->px;
  }
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 12: x' = a d, y' = g^2 - 3 e^2 + 21 q (< 31.7 q), z' = b h (Fq2 from kar < 3.8 q: + 19 q / + 8 q), l0 = -h py,
// l3 = 3 j px, l4 = e - b + 7 q (< 15 q); all canonicalised. r was last read in round 1.
TB_HD void wp_dbl_p5(WDouble* s, Fq12* line, int t) {
  const int c = t & 1, q = t >> 1;
  Fq(*v)[2] = s->v;
  Fq o;
  if (q == 0) wl_fq2_from_kar<19, 8>(o, &s->kar[0], c);
  else if (q == 1) {
    Fq g2v, e2v, tt;
    wl_sqr_asm(g2v, s->kar[3], s->kar[4], c);
    wl_sqr_asm(e2v, s->kar[5], s->kar[6], c);
    lz_mul3(tt, e2v);
    lz_sub(o, g2v, tt);
    lz_add_kq<21>(o);
  } else if (q == 2) wl_fq2_from_kar<19, 8>(o, &s->kar[7], c);
  else if (q == 3) o = s->kar[10 + c];
  else if (q == 4) o = s->kar[12 + c];
  else {
    lz_sub(o, v[WD_E][c], v[WD_B][c]);
    lz_add_kq<7>(o);
  }
  lz_canon(o);
  Fq* dst = q == 0 ? reinterpret_cast<Fq*>(&s->r.x)
          : q == 1 ? reinterpret_cast<Fq*>(&s->r.y)
          : q == 2 ? reinterpret_cast<Fq*>(&s->r.z)
          : q == 3 ? w12_q(line) + 0       // line = (l0, 0, 0) + (l3, l4, 0) w: tower slots 0, 3, 4
          : q == 4 ? w12_q(line) + 6
                   : w12_q(line) + 8;
  dst[c] = o;
}

// ---- addition step of the Miller loop (ark `add_in_place`, see g2_add_line) on parallel lanes ----------------------------------
// r = r + q and the line (l0, l3, l4) = (lambda py, -theta px, theta q.x - lambda q.y). On one thread the step is ~30
// dependent Fq products (~60 us); here four product rounds (6, 14, 9, 12 lanes) with the recombinations in between.
// r, q, px, py canonical on entry; r canonical and the line coefficients canonical on exit. Uses WDouble's scratch.
enum { WA_TH = 0, WA_LA, WA_C, WA_D, WA_E, WA_H, WA_GMH };
// S = kar[2] - kar[0] - kar[1] (wrapping): the u coefficient of a Karatsuba Fq2 product
TB_HD void wl_kar_c1(Fq& o, const Fq* kar) {
  lz_sub(o, kar[2], kar[0]);
  lz_sub(o, o, kar[1]);
}
// round A, item t < 6: q.y r.z (0-2), q.x r.z (3-5); operands < 2 q: kar < 1.03 q
TB_HD void wp_add_rA(WDouble* s, const Affine2* q, int t) {
  const Fq* rz = reinterpret_cast<const Fq*>(&s->r.z);
  const Fq* qc = reinterpret_cast<const Fq*>(t < 3 ? &q->y : &q->x);
  Fq a, b;
  wl_kar_operand(a, qc, qc, false, t % 3);
  wl_kar_operand(b, rz, rz, false, t % 3);
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 4: theta = r.y - q.y r.z, lambda = r.x - q.x r.z:  c0 = src.0 - k0 + 5 k1 + 2 q < 8.2 q,  c1 = src.1 - S + 2 q < 5.1 q
TB_HD void wp_add_pA(WDouble* s, int t) {
  const int which = t >> 1, c = t & 1;
  const Fq* src = reinterpret_cast<const Fq*>(which == 0 ? &s->r.y : &s->r.x);
  const Fq* k = &s->kar[3 * which];
  Fq o, m;
  if (c == 0) {
    lz_mul5(m, k[1]);
    lz_sub(o, src[0], k[0]);
    lz_add(o, o, m);
  } else {
    wl_kar_c1(m, k);
    lz_sub(o, src[1], m);
  }
  lz_add_kq<2>(o);
  s->v[WA_TH + which][c] = o;
}
// round B, item t < 14: 0-1 theta^2, 2-3 lambda^2 (operands < 13.3 q, < 34.2 q: kar < 4 q / < 1.28 q), 4-6 theta q.x,
// 7-9 lambda q.y (< 1.18 q), 10-11 lambda py, 12-13 theta px (< 1.06 q)
TB_HD void wp_add_rB(WDouble* s, const Affine2* q, int t) {
  Fq(*v)[2] = s->v;
  Fq a, b;
  if (t < 4) {
    wl_sqr_operands<26>(a, b, v[WA_TH + (t >> 1)], t & 1);
  } else if (t < 10) {
    const int which = (t - 4) / 3, part = (t - 4) % 3;
    const Fq* qc = reinterpret_cast<const Fq*>(which == 0 ? &q->x : &q->y);
    wl_kar_operand(a, v[WA_TH + which], v[WA_TH + which], false, part);
    wl_kar_operand(b, qc, qc, false, part);
  } else if (t < 12) {
    a = v[WA_LA][t - 10];
    b = s->py;
  } else {
    a = v[WA_TH][t - 12];
    b = s->px;
  }
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 10: 0-3 c = theta^2, d = lambda^2 (< 9.1 q, 2.6 q); 4-5 l4 = theta q.x - lambda q.y; 6-7 l0 = lambda py;
// 8-9 l3 = -theta px; the line coefficients canonical
TB_HD void wp_add_pB(WDouble* s, Fq12* line, int t) {
  const int c = t & 1;
  Fq o, m;
  if (t < 4) {
    const int which = t >> 1;
    wl_sqr_asm(o, s->kar[2 * which], s->kar[2 * which + 1], c);
    s->v[WA_C + which][c] = o;
    return;
  }
  if (t < 6) {
    const Fq* ka = &s->kar[4];   // theta q.x
    const Fq* kb = &s->kar[7];   // lambda q.y
    if (c == 0) {                // ka0 - kb0 + 5 (kb1 - ka1) + 8 q < 15.1 q
      lz_sub(m, kb[1], ka[1]);
      lz_mul5(m, m);
      lz_sub(o, ka[0], kb[0]);
      lz_add(o, o, m);
      lz_add_kq<8>(o);
    } else {                     // S(a) - S(b) + 4 q < 7.6 q
      wl_kar_c1(o, ka);
      wl_kar_c1(m, kb);
      lz_sub(o, o, m);
      lz_add_kq<4>(o);
    }
    lz_canon(o);
    w12_q(line)[8 + c] = o;
  } else if (t < 8) {
    o = s->kar[10 + c];
    lz_canon(o);
    w12_q(line)[0 + c] = o;
  } else {
    lz_neg_kq<2>(o, s->kar[12 + c]);
    lz_canon(o);
    w12_q(line)[6 + c] = o;
  }
}
// round C, item t < 9: e = lambda d (0-2: < 2.02 q), f = r.z c (3-5: < 1.16 q), g = r.x d (6-8: < 1.16 q)
TB_HD void wp_add_rC(WDouble* s, int t) {
  Fq(*v)[2] = s->v;
  const int pr = t / 3, part = t % 3;
  const Fq* x = pr == 0 ? v[WA_LA] : reinterpret_cast<const Fq*>(pr == 1 ? &s->r.z : &s->r.x);
  const Fq* y = pr == 1 ? v[WA_C] : v[WA_D];
  Fq a, b;
  wl_kar_operand(a, x, x, false, part);
  wl_kar_operand(b, y, y, false, part);
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 6: e (< 13.1 q, 7.1 q), h = e + f - 2 g (reduced), g - h = 3 g - e - f (reduced), straight from the products:
//   with X_i = ke_i + kf_i - 2 kg_i:   h.0 = X_0 - 5 X_1 + 19 q < 33.8 q,   h.1 = S(e) + S(f) - 2 S(g) + 9 q < 16.9 q
//   with Y_i = 3 kg_i - ke_i - kf_i:   (g-h).0 = Y_0 - 5 Y_1 + 21 q < 40.4 q,   (g-h).1 = 3 S(g) - S(e) - S(f) + 11 q < 20.9 q
TB_HD void wp_add_pC(WDouble* s, int t) {
  const int what = t >> 1, c = t & 1;
  const Fq* ke = &s->kar[0];
  const Fq* kf = &s->kar[3];
  const Fq* kg = &s->kar[6];
  Fq o;
  if (what == 0) {
    wl_fq2_from_kar<11, 5>(o, ke, c);
    s->v[WA_E][c] = o;
    return;
  }
  Fq xe, xf, xg, m;
  if (c == 0) {
    Fq x0, x1;
    if (what == 1) {
      lz_dbl(m, kg[0]);
      lz_add(x0, ke[0], kf[0]);
      lz_sub(x0, x0, m);
      lz_dbl(m, kg[1]);
      lz_add(x1, ke[1], kf[1]);
      lz_sub(x1, x1, m);
    } else {
      lz_mul3(m, kg[0]);
      lz_sub(x0, m, ke[0]);
      lz_sub(x0, x0, kf[0]);
      lz_mul3(m, kg[1]);
      lz_sub(x1, m, ke[1]);
      lz_sub(x1, x1, kf[1]);
    }
    lz_mul5(x1, x1);
    lz_sub(o, x0, x1);
    if (what == 1) lz_add_kq<19>(o);
    else lz_add_kq<21>(o);
  } else {
    wl_kar_c1(xe, ke);
    wl_kar_c1(xf, kf);
    wl_kar_c1(xg, kg);
    if (what == 1) {
      lz_dbl(m, xg);
      lz_add(o, xe, xf);
      lz_sub(o, o, m);
      lz_add_kq<9>(o);
    } else {
      lz_mul3(m, xg);
      lz_sub(o, m, xe);
      lz_sub(o, o, xf);
      lz_add_kq<11>(o);
    }
  }
  lz_reduce(o);
  s->v[what == 1 ? WA_H : WA_GMH][c] = o;
}
// round D, item t < 12: lambda h (0-2), theta (g - h) (3-5): < 1.18 q; e r.y (6-8), r.z e (9-11): < 1.27 q
TB_HD void wp_add_rD(WDouble* s, int t) {
  Fq(*v)[2] = s->v;
  const int pr = t / 3, part = t % 3;
  const Fq* x = pr == 0 ? v[WA_LA] : pr == 1 ? v[WA_TH] : v[WA_E];
  const Fq* y = pr == 0 ? v[WA_H] : pr == 1 ? v[WA_GMH] : reinterpret_cast<const Fq*>(pr == 2 ? &s->r.y : &s->r.z);
  Fq a, b;
  wl_kar_operand(a, x, x, false, part);
  wl_kar_operand(b, y, y, false, part);
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 6: r.x = lambda h, r.y = theta (g - h) - e r.y, r.z = r.z e; canonical
TB_HD void wp_add_pD(WDouble* s, int t) {
  const int which = t >> 1, c = t & 1;
  Fq o, m;
  if (which == 0) wl_fq2_from_kar<7, 3>(o, &s->kar[0], c);
  else if (which == 2) wl_fq2_from_kar<7, 3>(o, &s->kar[9], c);
  else {
    const Fq* ka = &s->kar[3];
    const Fq* kb = &s->kar[6];
    if (c == 0) {                // ka0 - kb0 + 5 (kb1 - ka1) + 8 q < 15.6 q
      lz_sub(m, kb[1], ka[1]);
      lz_mul5(m, m);
      lz_sub(o, ka[0], kb[0]);
      lz_add(o, o, m);
      lz_add_kq<8>(o);
    } else {                     // S(a) - S(b) + 4 q < 7.7 q
      wl_kar_c1(o, ka);
      wl_kar_c1(m, kb);
      lz_sub(o, o, m);
      lz_add_kq<4>(o);
    }
  }
  lz_canon(o);
  Fq* dst = reinterpret_cast<Fq*>(which == 0 ? &s->r.x : which == 1 ? &s->r.y : &s->r.z);
  dst[c] = o;
}

// ---- f^(q^6 - 1) = conj(f) / f without a serial Fq12 inversion ----------------------------------------------------------------
// With g = conj(f):  g / f = g^2 / (f g)  and  f g = a0^2 - v a1^2 = N lies in Fq6 (f = a0 + a1 w). The drivers compute
// n12 = f g and g^2 with the cooperative product / squaring; the bodies below invert N in Fq6 on parallel lanes:
//   t0 = n0^2 - xi n1 n2,  t1 = xi n2^2 - n0 n1,  t2 = n1^2 - n0 n2,   d = xi (n2 t1 + n1 t2) + n0 t0  (in Fq2),
//   N^-1 = (t0, t1, t2) / d,   1 / d = conj(d) / (d0^2 + 5 d1^2)
// -- three product rounds and ONE Fq inversion (binary Euclid, one lane) instead of ~150 dependent products and the
// inversion on one thread (a third of the final exponentiation's time). N's coefficients are read from tower slots
// 0..2 of n12 (< 1.02 q); the result is written to slots 0..2 of `out`, slots 3..5 are zeroed.
// round 1, item t < 18: Fq2 products 0: n0 n0, 1: n1 n2, 2: n2 n2, 3: n0 n1, 4: n1 n1, 5: n0 n2 (operands < 2.04 q: kar < 1.03 q)
TB_HD void wp_inv6_r1(WScratch* w, const Fq12* n12, int t) {
  const int pr = t / 3, part = t % 3;
  const int ia = pr == 0 ? 0 : pr == 1 ? 1 : pr == 2 ? 2 : pr == 3 ? 0 : pr == 4 ? 1 : 0;
  const int ib = pr == 0 ? 0 : pr == 1 ? 2 : pr == 2 ? 2 : pr == 3 ? 1 : pr == 4 ? 1 : 2;
  const Fq* x = w12_q(n12) + 2 * ia;
  const Fq* y = w12_q(n12) + 2 * ib;
  Fq a, b;
  wl_kar_operand(a, x, x, false, part);
  wl_kar_operand(b, y, y, false, part);
  w->kar[t] = fq_mul_lz(a, b);
}
// item t < 12: prod[t / 2][t & 1] (< 7.03 q, < 4.03 q)
TB_HD void wp_inv6_p1(WScratch* w, int t) {
  Fq o;
  wl_fq2_from_kar<6, 3>(o, &w->kar[3 * (t / 2)], t & 1);
  w->prod[t / 2][t & 1] = o;
}
// item t < 6: t_i coefficient c -> r6[0][i][c], reduced.  (xi z).0 = -5 z.1, (xi z).1 = z.0
//   t0 = P0 - xi P1: (P0.0 + 5 P1.1 < 27.2 q,  P0.1 - P1.0 + 8 q < 12.1 q)
//   t1 = xi P2 - P3: (-5 P2.1 - P3.0 + 28 q < 28 q,  P2.0 - P3.1 + 5 q < 12.1 q)
//   t2 = P4 - P5:    (+ 8 q < 15.1 q,  + 5 q < 9.1 q)
TB_HD void wp_inv6_p2(WScratch* w, int t) {
  const int i = t >> 1, c = t & 1;
  const Fq(*P)[2] = w->prod;
  Fq o, m;
  if (i == 0) {
    if (c == 0) {
      lz_mul5(m, P[1][1]);
      lz_add(o, P[0][0], m);
    } else {
      lz_sub(o, P[0][1], P[1][0]);
      lz_add_kq<8>(o);
    }
  } else if (i == 1) {
    if (c == 0) {
      lz_mul5(m, P[2][1]);
      lz_add(m, m, P[3][0]);
      lz_neg_kq<28>(o, m);
    } else {
      lz_sub(o, P[2][0], P[3][1]);
      lz_add_kq<5>(o);
    }
  } else {
    lz_sub(o, P[4][c], P[5][c]);
    if (c == 0) lz_add_kq<8>(o);
    else lz_add_kq<5>(o);
  }
  lz_reduce(o);
  w->r6[0][i][c] = o;
}
// round 2, item t < 9: Fq2 products 0: n2 t1, 1: n1 t2, 2: n0 t0 (kar < 1.03 q)
TB_HD void wp_inv6_r2(WScratch* w, const Fq12* n12, int t) {
  const int pr = t / 3, part = t % 3;
  const Fq* x = w12_q(n12) + 2 * (2 - pr);
  const Fq* y = w->r6[0][pr == 0 ? 1 : pr == 1 ? 2 : 0];
  Fq a, b;
  wl_kar_operand(a, x, x, false, part);
  wl_kar_operand(b, y, y, false, part);
  w->kar[t] = fq_mul_lz(a, b);
}
// ONE item (the team's lane 0): d = xi (Q0 + Q1) + Q2, then 1 / d = (d0, -d1) / (d0^2 + 5 d1^2) -> r6[1][0][0..1], canonical.
//   d.0 = k6 - 5 k7 - 5 (S0 + S1) + 16 q < 37.8 q,   d.1 = (k0 - 5 k1) + (k3 - 5 k4) + S2 + 13 q < 16.2 q
TB_HD void wp_inv6_d(WScratch* w) {
  const Fq* k = w->kar;
  Fq d0, d1, m, s;
  wl_kar_c1(m, k + 0);
  wl_kar_c1(s, k + 3);
  lz_add(m, m, s);
  lz_add(m, m, k[7]);
  lz_mul5(m, m);
  lz_sub(d0, k[6], m);
  lz_add_kq<16>(d0);
  lz_canon(d0);
  lz_add(m, k[1], k[4]);
  lz_mul5(m, m);
  wl_kar_c1(s, k + 6);
  lz_add(d1, k[0], k[3]);
  lz_add(d1, d1, s);
  lz_sub(d1, d1, m);
  lz_add_kq<13>(d1);
  lz_canon(d1);
  Fq n, t5, ni;
  fq_mul(n, d0, d0);
  fq_mul(m, d1, d1);
  fq_mul5(t5, m);
  fq_add(n, n, t5);
  fq_inv(ni, n);
  fq_mul(w->r6[1][0][0], d0, ni);
  fq_mul(m, d1, ni);
  fq_neg(w->r6[1][0][1], m);
}
// round 3, item t < 9: Fq2 products t_i * (1 / d), i = t / 3
TB_HD void wp_inv6_r3(WScratch* w, int t) {
  const int pr = t / 3, part = t % 3;
  Fq a, b;
  wl_kar_operand(a, w->r6[0][pr], w->r6[0][pr], false, part);
  wl_kar_operand(b, w->r6[1][0], w->r6[1][0], false, part);
  w->kar[t] = fq_mul_lz(a, b);
}
// item t < 12: coefficient t of `out`: slots 0..2 = N^-1 (reduced), slots 3..5 = 0
TB_HD void wp_inv6_out(Fq12* out, const WScratch* w, int t) {
  Fq o;
  if (t < 6) {
    wl_fq2_from_kar<6, 3>(o, &w->kar[3 * (t / 2)], t & 1);
    lz_reduce(o);
  } else {
    o = fq_zero();
  }
  w12_q(out)[t] = o;
}

// ---- L = line_a * line_b for two line values (l0, 0, 0) + (l3, l4, 0) w ------------------------------------------------------
// With x = (a0, a3, a4), y = (b0, b3, b4) the product is
//   (x0 y0 + xi x2 y2,  x1 y1,  x1 y2 + x2 y1)  +  (x0 y1 + x1 y0,  x0 y2 + x2 y0,  0) w
// -- the six Fq2 products of one Karatsuba Fq6 product (wp_kar / wp_fq2 with count = 1), 18 Fq products instead of the 54
// of a general product. Used where two pairs share one Miller accumulator: f <- f^2 * (line_a line_b).
// item t < 12: operands from the (canonical) line coefficients
TB_HD void wp_ll_xy(WScratch* w, const Fq12* la, const Fq12* lb, int t) {
  const int which = t / 6, rem = t % 6, tt = rem / 2, c = rem & 1;
  const Fq* src = w12_q(which ? lb : la);
  const int slot = tt == 0 ? 0 : tt == 1 ? 3 : 4;
  w->xy[which][0][tt][c] = src[2 * slot + c];
}
// item t < 12: coefficient t of L, reduced. prod (P0 < 10.1 q, P1 < 5.1 q; operands < 4 q, kar < 1.11 q):
//   slot 0 = p0 + xi p2 (c0: p0.0 - 5 p2.1 + 26 q < 36.1 q; c1: p0.1 + p2.0 < 15.2 q), slot 1 = p1,
//   slot 2 = p3 - p1 - p2, slot 3 = p4 - p0 - p1, slot 4 = p5 - p0 - p2 (+ 21 q < 31.1 q / + 11 q < 16.1 q), slot 5 = 0
TB_HD void wp_ll_out(Fq12* L, const WScratch* w, int t) {
  const int slot = t / 2, c = t & 1;
  const Fq(*p)[2] = w->prod;
  Fq o, m;
  if (slot == 0) {
    if (c == 0) {
      lz_mul5(m, p[2][1]);
      lz_sub(o, p[0][0], m);
      lz_add_kq<26>(o);
    } else {
      lz_add(o, p[0][1], p[2][0]);
    }
  } else if (slot == 1) {
    o = p[1][c];
  } else if (slot == 5) {
    o = fq_zero();
  } else {
    const int hi = slot == 2 ? 3 : slot == 3 ? 4 : 5;
    const int s1 = slot == 2 ? 1 : 0;
    const int s2 = slot == 3 ? 1 : 2;
    lz_sub(o, p[hi][c], p[s1][c]);
    lz_sub(o, o, p[s2][c]);
    if (c == 0) lz_add_kq<21>(o);
    else lz_add_kq<11>(o);
  }
  lz_reduce(o);
  w12_q(L)[t] = o;
}

// ---- XYZZ arithmetic over Fq2 on parallel lanes (the window combine of a single G2 MSM) ---------------------------------------
// The Horner form sum_w 2^(c w) S_w is ~253 dependent doublings + W additions: on one thread 27 / 46 us each. Here the
// Fq2 products of a doubling (dbl-2008-s-1) run as three rounds of 4 / 11 / 9 Fq products, those of an addition
// (add-2008-s) as four rounds of 12 / 10 / 9 / 9, with the lazily reduced recombinations in between.
// Accumulator p and operand e: coefficients < 1.01 q (identity <=> ZZ = 0 mod q; any representative of 0).
struct alignas(16) WG2 {
  Xyzz2 p, e;
  Fq kar[12];
  Fq v[9][2];
  int flag;          // addition: 0 = generic, 1 = e is the identity, 2 = p is the identity, 3 = equal x (exact path)
};
enum { G2V_U = 0, G2V_V, G2V_M, G2V_W, G2V_T, G2V_S1, G2V_PP, G2V_ZZ12, G2V_ZZZ12 };
// (k0 - 5 k1) - (k0' - 5 k1') and S(k) - S(k') of two Karatsuba triples, + OFF q
template <unsigned O0, unsigned O1>
TB_HD void wl_fq2_diff_from_kar(Fq& o, const Fq* ka, const Fq* kb, int c) {
  Fq m;
  if (c == 0) {
    lz_sub(m, kb[1], ka[1]);
    lz_mul5(m, m);
    lz_sub(o, ka[0], kb[0]);
    lz_add(o, o, m);
    lz_add_kq<O0>(o);
  } else {
    wl_kar_c1(o, ka);
    wl_kar_c1(m, kb);
    lz_sub(o, o, m);
    lz_add_kq<O1>(o);
  }
}
// doubling, round 1, item t < 4: U^2 (U = 2 Y: operands < 4.1 q, < 13.2 q) and X^2 as (v0 + v1)(v0 - 5 v1), v0 v1
TB_HD void wp_g2dbl_r1(WG2* s, int t) {
  const Fq* y = reinterpret_cast<const Fq*>(&s->p.y);
  const Fq* x = reinterpret_cast<const Fq*>(&s->p.x);
  Fq a, b;
  if (t < 2) {
    Fq w[2];
    lz_dbl(w[0], y[0]);
    lz_dbl(w[1], y[1]);
    wl_sqr_operands<11>(a, b, w, t & 1);
  } else {
    wl_sqr_operands<6>(a, b, x, t & 1);
  }
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 6: V = U^2 (< 5.6 q, 2.1 q), M = 3 X^2 (< 15.5 q, 6.1 q), U = 2 Y
TB_HD void wp_g2dbl_p1(WG2* s, int t) {
  const int c = t & 1, what = t >> 1;
  Fq o;
  if (what == 0) wl_sqr_asm(o, s->kar[0], s->kar[1], c);
  else if (what == 1) {
    wl_sqr_asm(o, s->kar[2], s->kar[3], c);
    lz_mul3(o, o);
  } else {
    lz_dbl(o, reinterpret_cast<const Fq*>(&s->p.y)[c]);
  }
  s->v[what == 0 ? G2V_V : what == 1 ? G2V_M : G2V_U][c] = o;
}
// round 2, item t < 11: 0-2 W = U V (< 1.22 q), 3-5 S = X V (< 1.12 q), 6-7 M^2 (operands < 21.6 q, < 46.5 q: < 7.6 q; m0 m1
// < 1.63 q), 8-10 V ZZ (< 1.12 q)
TB_HD void wp_g2dbl_r2(WG2* s, int t) {
  Fq(*v)[2] = s->v;
  Fq a, b;
  if (t < 3) {
    wl_kar_operand(a, v[G2V_U], v[G2V_U], false, t);
    wl_kar_operand(b, v[G2V_V], v[G2V_V], false, t);
  } else if (t < 6) {
    const Fq* x = reinterpret_cast<const Fq*>(&s->p.x);
    wl_kar_operand(a, x, x, false, t - 3);
    wl_kar_operand(b, v[G2V_V], v[G2V_V], false, t - 3);
  } else if (t < 8) {
    wl_sqr_operands<31>(a, b, v[G2V_M], t & 1);
  } else {
    const Fq* zz = reinterpret_cast<const Fq*>(&s->p.zz);
    wl_kar_operand(a, v[G2V_V], v[G2V_V], false, t - 8);
    wl_kar_operand(b, zz, zz, false, t - 8);
  }
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 8: X' = M^2 - 2 S (reduced), T = S - X' = 3 S - M^2 (< 34.5 q, 13.5 q), W (< 8.3 q, 4.3 q), ZZ' = V ZZ (reduced).
// With M^2 = (Q0, Q1) = (s0 + 4 s1, 2 s1) < (14.2 q, 3.3 q) and S from kar[3..5]:
//   X'.0 = Q0 - 2 k3 + 10 k4 + 3 q < 28.3 q,   X'.1 = Q1 - 2 S(k3..5) + 3 q < 10.8 q
//   T.0 = 3 k3 - 15 k4 - Q0 + 31 q,           T.1 = 3 S(k3..5) - Q1 + 10 q
TB_HD void wp_g2dbl_p2(WG2* s, int t) {
  const int c = t & 1, what = t >> 1;
  const Fq* ks = &s->kar[3];
  Fq o, m, q;
  if (what < 2) {
    wl_sqr_asm(q, s->kar[6], s->kar[7], c);
    if (c == 0) {                       // sv = k3 - 5 k4 (wrapping)
      lz_mul5(m, ks[1]);
      lz_sub(m, ks[0], m);
    } else {
      wl_kar_c1(m, ks);
    }
    if (what == 0) {
      lz_dbl(m, m);
      lz_sub(o, q, m);
      lz_add_kq<3>(o);
      lz_reduce(o);
      reinterpret_cast<Fq*>(&s->p.x)[c] = o;
    } else {
      lz_mul3(m, m);
      lz_sub(o, m, q);
      if (c == 0) lz_add_kq<31>(o);
      else lz_add_kq<10>(o);
      s->v[G2V_T][c] = o;
    }
  } else if (what == 2) {
    wl_fq2_from_kar<7, 3>(o, &s->kar[0], c);
    s->v[G2V_W][c] = o;
  } else {
    wl_fq2_from_kar<6, 3>(o, &s->kar[8], c);
    lz_reduce(o);
    reinterpret_cast<Fq*>(&s->p.zz)[c] = o;
  }
}
// round 3, item t < 9: 0-2 M T (< 7.8 q), 3-5 W Y (< 1.18 q), 6-8 W ZZZ (< 1.18 q)
TB_HD void wp_g2dbl_r3(WG2* s, int t) {
  Fq(*v)[2] = s->v;
  const int pr = t / 3, part = t % 3;
  const Fq* x = pr == 0 ? v[G2V_M] : v[G2V_W];
  const Fq* y = pr == 0 ? v[G2V_T] : reinterpret_cast<const Fq*>(pr == 1 ? &s->p.y : &s->p.zzz);
  Fq a, b;
  wl_kar_operand(a, x, x, false, part);
  wl_kar_operand(b, y, y, false, part);
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 4: Y' = M T - W Y (+ 40 q < 53.7 q / + 17 q < 27.2 q), ZZZ' = W ZZZ; reduced
TB_HD void wp_g2dbl_p3(WG2* s, int t) {
  const int c = t & 1;
  Fq o;
  if (t < 2) {
    wl_fq2_diff_from_kar<40, 17>(o, &s->kar[0], &s->kar[3], c);
    lz_reduce(o);
    reinterpret_cast<Fq*>(&s->p.y)[c] = o;
  } else {
    wl_fq2_from_kar<6, 3>(o, &s->kar[6], c);
    lz_reduce(o);
    reinterpret_cast<Fq*>(&s->p.zzz)[c] = o;
  }
}

// addition p += e. ONE item first: the exceptional shapes (identity operands), decided on canonical values
TB_HD void wp_g2add_flags(WG2* s) {
  Fq2 pz = s->p.zz, ez = s->e.zz;
  lz_canon(pz.c0);
  lz_canon(pz.c1);
  lz_canon(ez.c0);
  lz_canon(ez.c1);
  s->flag = fq2_is_zero(ez) ? 1 : (fq2_is_zero(pz) ? 2 : 0);
}
// round 1, item t < 12: U1 = X1 ZZ2 (0-2), U2 = X2 ZZ1 (3-5), S1 = Y1 ZZZ2 (6-8), S2 = Y2 ZZZ1 (9-11); kar < 1.03 q
TB_HD void wp_g2add_r1(WG2* s, int t) {
  const int pr = t / 3, part = t % 3;
  const Fq2* xa = pr == 0 ? &s->p.x : pr == 1 ? &s->e.x : pr == 2 ? &s->p.y : &s->e.y;
  const Fq2* xb = pr == 0 ? &s->e.zz : pr == 1 ? &s->p.zz : pr == 2 ? &s->e.zzz : &s->p.zzz;
  const Fq* x = reinterpret_cast<const Fq*>(xa);
  const Fq* y = reinterpret_cast<const Fq*>(xb);
  Fq a, b;
  wl_kar_operand(a, x, x, false, part);
  wl_kar_operand(b, y, y, false, part);
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 8: P = U2 - U1 -> v[T] (canonical), R = S2 - S1 -> v[M] (canonical), U1 -> v[U], S1 -> v[S1] (< 7.1 q, 4.1 q)
TB_HD void wp_g2add_p1(WG2* s, int t) {
  const int c = t & 1, what = t >> 1;
  Fq o;
  if (what == 0) {
    wl_fq2_diff_from_kar<7, 4>(o, &s->kar[3], &s->kar[0], c);
    lz_canon(o);
    s->v[G2V_T][c] = o;
  } else if (what == 1) {
    wl_fq2_diff_from_kar<7, 4>(o, &s->kar[9], &s->kar[6], c);
    lz_canon(o);
    s->v[G2V_M][c] = o;
  } else if (what == 2) {
    wl_fq2_from_kar<6, 3>(o, &s->kar[0], c);
    s->v[G2V_U][c] = o;
  } else {
    wl_fq2_from_kar<6, 3>(o, &s->kar[6], c);
    s->v[G2V_S1][c] = o;
  }
}
// ONE item: equal x coordinates (P = 0): doubling or the identity, exactly, with the canonical routines
TB_HD void wp_g2add_check(WG2* s) {
  Fq2 pv;
  pv.c0 = s->v[G2V_T][0];
  pv.c1 = s->v[G2V_T][1];
  if (!fq2_is_zero(pv)) return;
  s->flag = 3;
  Xyzz2 a = s->p, e = s->e;
  for (int i = 0; i < 8; i++) {
    lz_canon(reinterpret_cast<Fq*>(&a)[i]);
    lz_canon(reinterpret_cast<Fq*>(&e)[i]);
  }
  xyzz2_add(a, e);
  s->p = a;
}
// round 2, item t < 10: 0-1 P^2, 2-3 R^2 (< 1.1 q), 4-6 ZZ1 ZZ2, 7-9 ZZZ1 ZZZ2 (< 1.03 q)
TB_HD void wp_g2add_r2(WG2* s, int t) {
  Fq(*v)[2] = s->v;
  Fq a, b;
  if (t < 4) {
    wl_sqr_operands<6>(a, b, t < 2 ? v[G2V_T] : v[G2V_M], t & 1);
  } else {
    const int part = (t - 4) % 3;
    const Fq* x = reinterpret_cast<const Fq*>(t < 7 ? &s->p.zz : &s->p.zzz);
    const Fq* y = reinterpret_cast<const Fq*>(t < 7 ? &s->e.zz : &s->e.zzz);
    wl_kar_operand(a, x, x, false, part);
    wl_kar_operand(b, y, y, false, part);
  }
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 8: PP -> v[PP] (< 5.2 q, 2.1 q), RR -> v[V], ZZ1 ZZ2 -> v[ZZ12], ZZZ1 ZZZ2 -> v[ZZZ12] (< 7.1 q, 4.1 q)
TB_HD void wp_g2add_p2(WG2* s, int t) {
  const int c = t & 1, what = t >> 1;
  Fq o;
  if (what == 0) wl_sqr_asm(o, s->kar[0], s->kar[1], c);
  else if (what == 1) wl_sqr_asm(o, s->kar[2], s->kar[3], c);
  else if (what == 2) wl_fq2_from_kar<6, 3>(o, &s->kar[4], c);
  else wl_fq2_from_kar<6, 3>(o, &s->kar[7], c);
  s->v[what == 0 ? G2V_PP : what == 1 ? G2V_V : what == 2 ? G2V_ZZ12 : G2V_ZZZ12][c] = o;
}
// round 3, item t < 9: 0-2 PPP = P PP (< 1.1 q), 3-5 Q = U1 PP (< 1.53 q), 6-8 ZZ' = ZZ12 PP (< 1.53 q)
TB_HD void wp_g2add_r3(WG2* s, int t) {
  Fq(*v)[2] = s->v;
  const int pr = t / 3, part = t % 3;
  const Fq* x = pr == 0 ? v[G2V_T] : pr == 1 ? v[G2V_U] : v[G2V_ZZ12];
  Fq a, b;
  wl_kar_operand(a, x, x, false, part);
  wl_kar_operand(b, v[G2V_PP], v[G2V_PP], false, part);
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 8: X' = RR - PPP - 2 Q (reduced -> p.x), T = Q - X' = 3 Q + PPP - RR -> v[T], PPP -> v[W] (< 7.1 q, 4.1 q),
// ZZ' (reduced -> p.zz). With PPP from kar[0..2] (< 1.1 q), Q from kar[3..5] (< 1.53 q), RR = v[V]:
//   X'.0 = RR.0 - k0 - 2 k3 + 5 k1 + 10 k4 + 5 q < 30.9 q,     X'.1 = RR.1 - S(k0..2) - 2 S(k3..5) + 5 q < 15.4 q
//   T.0 = 3 k3 - 15 k4 + k0 - 5 k1 - RR.0 + 34 q < 39.8 q,     T.1 = 3 S(k3..5) + S(k0..2) - RR.1 + 14 q < 19.8 q
TB_HD void wp_g2add_p3(WG2* s, int t) {
  const int c = t & 1, what = t >> 1;
  const Fq* kp = &s->kar[0];
  const Fq* kq = &s->kar[3];
  Fq o, a, b;
  if (what < 2) {
    if (c == 0) {                       // a = PPP.0, b = Q.0 (wrapping, no offsets)
      lz_mul5(a, kp[1]);
      lz_sub(a, kp[0], a);
      lz_mul5(b, kq[1]);
      lz_sub(b, kq[0], b);
    } else {
      wl_kar_c1(a, kp);
      wl_kar_c1(b, kq);
    }
    if (what == 0) {
      lz_sub(o, s->v[G2V_V][c], a);
      lz_dbl(b, b);
      lz_sub(o, o, b);
      lz_add_kq<5>(o);
      lz_reduce(o);
      reinterpret_cast<Fq*>(&s->p.x)[c] = o;
    } else {
      lz_mul3(b, b);
      lz_add(o, b, a);
      lz_sub(o, o, s->v[G2V_V][c]);
      if (c == 0) lz_add_kq<34>(o);
      else lz_add_kq<14>(o);
      s->v[G2V_T][c] = o;
    }
  } else if (what == 2) {
    wl_fq2_from_kar<6, 3>(o, kp, c);
    s->v[G2V_W][c] = o;
  } else {
    wl_fq2_from_kar<8, 4>(o, &s->kar[6], c);
    lz_reduce(o);
    reinterpret_cast<Fq*>(&s->p.zz)[c] = o;
  }
}
// round 4, item t < 9: 0-2 R T (< 1.79 q), 3-5 S1 PPP (< 1.82 q), 6-8 ZZZ' = ZZZ12 PPP (< 1.82 q)
TB_HD void wp_g2add_r4(WG2* s, int t) {
  Fq(*v)[2] = s->v;
  const int pr = t / 3, part = t % 3;
  const Fq* x = pr == 0 ? v[G2V_M] : pr == 1 ? v[G2V_S1] : v[G2V_ZZZ12];
  const Fq* y = pr == 0 ? v[G2V_T] : v[G2V_W];
  Fq a, b;
  wl_kar_operand(a, x, x, false, part);
  wl_kar_operand(b, y, y, false, part);
  s->kar[t] = fq_mul_lz(a, b);
}
// item t < 4: Y' = R T - S1 PPP (+ 11 q < 21.9 q / + 6 q < 11.5 q), ZZZ'; reduced
TB_HD void wp_g2add_p4(WG2* s, int t) {
  const int c = t & 1;
  Fq o;
  if (t < 2) {
    wl_fq2_diff_from_kar<11, 6>(o, &s->kar[0], &s->kar[3], c);
    lz_reduce(o);
    reinterpret_cast<Fq*>(&s->p.y)[c] = o;
  } else {
    wl_fq2_from_kar<10, 4>(o, &s->kar[6], c);
    lz_reduce(o);
    reinterpret_cast<Fq*>(&s->p.zzz)[c] = o;
  }
}

}  // namespace tb
