// Pairing products on the device: `E::multi_pairing(g1s, g2s)` = one Miller loop per pair, the
// product of the Miller values, ONE final exponentiation -- src/sqrt_pst.rs:131-144 (t, 2^m_col pairs) and
// src/mipp.rs:87-94 (comm_t_l / comm_t_r, two products per round over the halves of a and h).
//
// Shape of the work: 2^13 pairs at the C3 size, each ~7 000 dependent Fq products (63 doubling steps + 6 addition steps,
// every step one Fq12 squaring and a 13-product sparse line multiplication). Pairs are independent, so the Miller stage is
// one thread per pair (2^13 threads ~ two warps per SM sub-partition; the integer pipe is fed by the three interleaved
// carry chains of every Fq2 product); the product tree has fan-in 8 per launch; the final exponentiation is a single
// dependent chain and runs on one thread per product (the two products of a MIPP round side by side).
// Fq12 values are 576 B (36 uint4) in ark's in-memory order.
#pragma once
#include "fq12.cuh"
#include "kernels_g2.cuh"

namespace tb {

static_assert(sizeof(Fq12) == 576, "packed Fq12");

// NB: the word pointer is taken from the WHOLE object. Indexing past the 12-word array of the first member
// (`f.c0.c0.c0.l[4 * i]`, i dynamic) is undefined behaviour that NVVM exploits: it treated the words it could not
// prove written as undefined and dropped their copies (found as a single wrong limb in fq12_conj's output).
__device__ __forceinline__ void load_fq12(Fq12& f, const uint4* src) {
  uint32_t* d = reinterpret_cast<uint32_t*>(&f);
#pragma unroll
  for (int i = 0; i < 36; i++) {
    uint4 v = src[i];
    d[4 * i + 0] = v.x;
    d[4 * i + 1] = v.y;
    d[4 * i + 2] = v.z;
    d[4 * i + 3] = v.w;
  }
}
__device__ __forceinline__ void store_fq12(uint4* dst, const Fq12& f) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&f);
#pragma unroll
  for (int i = 0; i < 36; i++) dst[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
}

// f[j] = Miller(g1[j], g2[j ^ xor_mask]). xor_mask = 0: plain pairing product. xor_mask = split (a power of two,
// n = 2 split): the two cross products of a MIPP round in one launch -- f[0, split) pairs a_l with h_r and
// f[split, 2 split) pairs a_r with h_l (src/mipp.rs:89-93).
__global__ void __launch_bounds__(32) k_miller(const uint4* __restrict__ g1, const uint4* __restrict__ g2, uint32_t n,
                                               uint32_t xor_mask, uint4* __restrict__ f_out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  Affine p;
  Affine2 q;
  load_affine(p, g1 + 6 * (size_t)j);
  load_affine2(q, g2 + 12 * (size_t)(j ^ xor_mask));
  Fq12 f;
  miller_loop(f, p, q);
  store_fq12(f_out + 36 * (size_t)j, f);
}

// one level of the product tree over `segs` independent segments of length len: out[s][t] = prod_k in[s][t + k m],
// m = ceil(len / FAN); blockIdx.y = segment
constexpr int FQ12_FAN = 8;
__global__ void __launch_bounds__(32) k_fq12_prod_level(const uint4* __restrict__ in, uint32_t len, uint32_t m,
                                                        uint4* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const uint4* src = in + 36 * (size_t)blockIdx.y * len;
  Fq12 acc, x;
  load_fq12(acc, src + 36 * (size_t)t);
  for (int k = 1; k < FQ12_FAN; k++) {
    const uint64_t idx = (uint64_t)t + (uint64_t)k * m;
    if (idx >= len) break;
    load_fq12(x, src + 36 * idx);
    fq12_mul_ol(&acc, &acc, &x);
  }
  store_fq12(out + 36 * ((size_t)blockIdx.y * m + t), acc);
}

// ---- warp-cooperative Fq12 arithmetic --------------------------------------------------------------------------------
// A final exponentiation is ONE dependent chain of ~315 cyclotomic squarings and ~45 Fq12 products; run by a single
// thread it took 20 ms (31 idle lanes, every Fq2 product in sequence). Here one WARP owns the chain: the Fq12 values live
// in shared memory and the 18 (12, 6) independent Fq2 products inside every Fq12 product (squaring, cyclotomic squaring)
// run on 18 lanes at once; the cheap linear recombinations are spread over lanes the same way. Phases are separated by
// __syncwarp(). All pointers are shared-memory objects of the calling warp; dst may alias the inputs.
struct WScratch {
  Fq2 xy[2][3][3];   // materialised Fq6 operands X_i, Y_i (i < 3) of up to three Fq6 products
  Fq2 prod[18];      // Karatsuba products: [i][j], j = 0..5 -> a0b0, a1b1, a2b2, (a1+a2)(b1+b2), (a0+a1)(b0+b1), (a0+a2)(b0+b2)
  Fq2 r6[3][3];      // the three Fq6 results
};
__device__ __forceinline__ Fq2* w12_c(Fq12* a, int idx) { return reinterpret_cast<Fq2*>(a) + idx; }
__device__ __forceinline__ const Fq2* w12_c(const Fq12* a, int idx) { return reinterpret_cast<const Fq2*>(a) + idx; }

// phases 1-2 of every product: lane (i, j) multiplies the Karatsuba operands of Fq6 product i; lane (i, t) then
// assembles coefficient t of result i. `count` = number of Fq6 products (1..3).
__device__ __noinline__ void w_fq6_products(WScratch* w, int count) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane < 6 * count) {
    const int i = lane / 6, j = lane % 6;
    const int t0 = (j < 3) ? j : (j == 3 ? 1 : 0);
    const int t1 = (j == 4) ? 1 : 2;
    Fq2 a = w->xy[0][i][t0], b = w->xy[1][i][t0];
    if (j >= 3) {
      fq2_add(a, a, w->xy[0][i][t1]);
      fq2_add(b, b, w->xy[1][i][t1]);
    }
    fq2_mul_ol(&w->prod[lane], &a, &b);
  }
  __syncwarp();
  if (lane < 3 * count) {
    const int i = lane / 3, t = lane % 3;
    const Fq2* p = &w->prod[6 * i];
    Fq2 r, m;
    if (t == 0) {          // v0 + xi (m12 - v1 - v2)
      fq2_sub(m, p[3], p[1]);
      fq2_sub(m, m, p[2]);
      fq2_mul_xi(m, m);
      fq2_add(r, p[0], m);
    } else if (t == 1) {   // m01 - v0 - v1 + xi v2
      fq2_sub(r, p[4], p[0]);
      fq2_sub(r, r, p[1]);
      fq2_mul_xi(m, p[2]);
      fq2_add(r, r, m);
    } else {               // m02 - v0 - v2 + v1
      fq2_sub(r, p[5], p[0]);
      fq2_sub(r, r, p[2]);
      fq2_add(r, r, p[1]);
    }
    w->r6[i][t] = r;
  }
  __syncwarp();
}

// dst = a * b: X = {a0, a1, a0 + a1}, Y = {b0, b1, b0 + b1}; C0 = R0 + v R1, C1 = R2 - R0 - R1
__device__ __noinline__ void w12_mul(Fq12* dst, const Fq12* a, const Fq12* b, WScratch* w) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane < 18) {
    const int which = lane / 9, i = (lane % 9) / 3, t = lane % 3;
    const Fq12* src = which ? b : a;
    Fq2 v;
    if (i < 2) v = *w12_c(src, 3 * i + t);
    else fq2_add(v, *w12_c(src, t), *w12_c(src, 3 + t));
    w->xy[which][i][t] = v;
  }
  w_fq6_products(w, 3);
  if (lane < 6) {
    Fq2 r;
    if (lane < 3) {        // C0.c[t] = R0[t] + (v R1)[t], v (x0, x1, x2) = (xi x2, x0, x1)
      Fq2 m = w->r6[1][(lane + 2) % 3];
      if (lane == 0) fq2_mul_xi(m, m);
      fq2_add(r, w->r6[0][lane], m);
    } else {
      const int t = lane - 3;
      fq2_sub(r, w->r6[2][t], w->r6[0][t]);
      fq2_sub(r, r, w->r6[1][t]);
    }
    *w12_c(dst, lane) = r;
  }
  __syncwarp();
}

// dst = a^2 (complex squaring): R0 = a0 a1, R1 = (a0 + a1)(a0 + v a1); C0 = R1 - R0 - v R0, C1 = 2 R0
__device__ __noinline__ void w12_sqr(Fq12* dst, const Fq12* a, WScratch* w) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane < 12) {
    const int which = lane / 6, i = (lane % 6) / 3, t = lane % 3;
    Fq2 v;
    if (i == 0) v = *w12_c(a, 3 * which + t);                   // X0 = a0, Y0 = a1
    else if (which == 0) fq2_add(v, *w12_c(a, t), *w12_c(a, 3 + t));   // X1 = a0 + a1
    else {                                                      // Y1 = a0 + v a1
      Fq2 m = *w12_c(a, 3 + (t + 2) % 3);
      if (t == 0) fq2_mul_xi(m, m);
      fq2_add(v, *w12_c(a, t), m);
    }
    w->xy[which][i][t] = v;
  }
  w_fq6_products(w, 2);
  if (lane < 6) {
    Fq2 r;
    if (lane < 3) {
      Fq2 m = w->r6[0][(lane + 2) % 3];
      if (lane == 0) fq2_mul_xi(m, m);
      fq2_sub(r, w->r6[1][lane], w->r6[0][lane]);
      fq2_sub(r, r, m);
    } else {
      fq2_dbl(r, w->r6[0][lane - 3]);
    }
    *w12_c(dst, lane) = r;
  }
  __syncwarp();
}

// dst = a^2 for a unitary a (Granger-Scott, as fq12_cyclotomic_sqr_ol): six Fq2 products on six lanes.
// tower slot of z_k: z0 = c[0], z1 = c[4], z2 = c[3], z3 = c[2], z4 = c[1], z5 = c[5]
__device__ __noinline__ void w12_cyclotomic_sqr(Fq12* dst, const Fq12* a, WScratch* w) {
  const int lane = threadIdx.x & 31;
  constexpr int slot[6] = {0, 4, 3, 2, 1, 5};
  __syncwarp();
  if (lane < 6) {
    const int p = lane >> 1;
    const Fq2 za = *w12_c(a, slot[2 * p]), zb = *w12_c(a, slot[2 * p + 1]);
    if ((lane & 1) == 0) fq2_mul_ol(&w->prod[lane], &za, &zb);          // tmp = za zb
    else {
      Fq2 s, m;
      fq2_add(s, za, zb);
      fq2_mul_xi(m, zb);
      fq2_add(m, m, za);
      fq2_mul_ol(&w->prod[lane], &s, &m);                               // (za + zb)(za + xi zb)
    }
  }
  __syncwarp();
  if (lane < 6) {
    // z_k' = 3 t - 2 z_k (k = 0, 3, 4) or 3 t + 2 z_k (k = 1, 2, 5), with t = t0, t1, xi t5, t4, t2, t3 for k = 0..5 where
    // t_{2p} = prod[2p+1] - tmp - xi tmp and t_{2p+1} = 2 tmp (tmp = prod[2p])
    constexpr int tsel[6] = {0, 1, 5, 4, 2, 3};
    const int ti = tsel[lane], p = ti >> 1;
    Fq2 t, m;
    if ((ti & 1) == 0) {
      fq2_sub(t, w->prod[2 * p + 1], w->prod[2 * p]);
      fq2_mul_xi(m, w->prod[2 * p]);
      fq2_sub(t, t, m);
    } else {
      fq2_dbl(t, w->prod[2 * p]);
    }
    if (lane == 2) fq2_mul_xi(t, t);
    const Fq2 z = *w12_c(a, slot[lane]);
    Fq2 o;
    if (lane == 0 || lane == 3 || lane == 4) fq2_sub(o, t, z);
    else fq2_add(o, t, z);
    fq2_dbl(o, o);
    fq2_add(o, o, t);
    *w12_c(dst, slot[lane]) = o;   // each lane rewrites only the slot it read in this phase: dst may alias a
  }
  __syncwarp();
}

__device__ __forceinline__ void w12_copy(Fq12* dst, const Fq12* a) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  Fq2 v;
  if (lane < 6) v = *w12_c(a, lane);
  __syncwarp();
  if (lane < 6) *w12_c(dst, lane) = v;
  __syncwarp();
}
__device__ __forceinline__ void w12_conj(Fq12* dst, const Fq12* a) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  Fq2 v;
  if (lane < 6) {
    v = *w12_c(a, lane);
    if (lane >= 3) fq2_neg(v, v);
  }
  __syncwarp();
  if (lane < 6) *w12_c(dst, lane) = v;
  __syncwarp();
}
// a^(q^k), k = 1, 2: tower slot idx holds the coefficient of w^e, e = 2 idx (idx < 3) or 2 (idx - 3) + 1
__device__ __noinline__ void w12_frobenius(Fq12* dst, const Fq12* a, int k) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  Fq2 c;
  if (lane < 6) {
    const int e = lane < 3 ? 2 * lane : 2 * (lane - 3) + 1;
    c = *w12_c(a, lane);
    if (k == 1) fq2_conj(c, c);
    if (e > 0) {
      if (k == 1) {
        Fq2 g;
        g.c0 = fq_from_table(FQ12_C(FROB1)[e - 1][0]);
        g.c1 = fq_from_table(FQ12_C(FROB1)[e - 1][1]);
        fq2_mul_ol(&c, &c, &g);
      } else {
        const Fq g = fq_from_table(FQ12_C(FROB2)[e - 1]);
        fq2_scale(c, c, g);
      }
    }
  }
  __syncwarp();
  if (lane < 6) *w12_c(dst, lane) = c;
  __syncwarp();
}
__device__ __noinline__ void w12_exp_by_x(Fq12* dst, const Fq12* a, Fq12* acc, WScratch* w) {
  w12_copy(acc, a);
  for (int bit = 62; bit >= 0; bit--) {
    w12_cyclotomic_sqr(acc, acc, w);
    if ((BLS_X >> bit) & 1) w12_mul(acc, acc, a, w);
  }
  w12_copy(dst, acc);
}

struct WFinalExp {
  Fq12 f, r, f2, y0, y1, y2, acc;
  WScratch w;
};

// the chain of fq12_final_exp_ol (ark `final_exponentiation`), one warp; s->f holds the input, the result lands in s->r
__device__ __noinline__ void w12_final_exp(WFinalExp* s) {
  const int lane = threadIdx.x & 31;
  WScratch* w = &s->w;
  w12_conj(&s->r, &s->f);
  if (lane == 0) fq12_inv_ol(&s->f2, &s->f);      // one inversion per product: a lone Fq inversion dominates it
  __syncwarp();
  w12_mul(&s->r, &s->r, &s->f2, w);
  w12_copy(&s->f2, &s->r);
  w12_frobenius(&s->r, &s->r, 2);
  w12_mul(&s->r, &s->r, &s->f2, w);
  w12_cyclotomic_sqr(&s->y0, &s->r, w);
  w12_exp_by_x(&s->y1, &s->r, &s->acc, w);
  w12_conj(&s->y2, &s->r);
  w12_mul(&s->y1, &s->y1, &s->y2, w);
  w12_exp_by_x(&s->y2, &s->y1, &s->acc, w);
  w12_conj(&s->y1, &s->y1);
  w12_mul(&s->y1, &s->y1, &s->y2, w);
  w12_exp_by_x(&s->y2, &s->y1, &s->acc, w);
  w12_frobenius(&s->y1, &s->y1, 1);
  w12_mul(&s->y1, &s->y1, &s->y2, w);
  w12_mul(&s->r, &s->r, &s->y0, w);
  w12_exp_by_x(&s->y0, &s->y1, &s->acc, w);
  w12_exp_by_x(&s->y2, &s->y0, &s->acc, w);
  w12_frobenius(&s->y0, &s->y1, 2);
  w12_conj(&s->y1, &s->y1);
  w12_mul(&s->y1, &s->y1, &s->y2, w);
  w12_mul(&s->y1, &s->y1, &s->y0, w);
  w12_mul(&s->r, &s->r, &s->y1, w);
}

// ---- warp-cooperative Miller loop (one warp per pair) --------------------------------------------------------------
// Thread-per-pair keeps the integer pipe fed only when there are thousands of pairs; the MIPP rounds of the reference
// issue products of 2^12 ... 1 pairs and every one of them waits for a full 10 ms single-thread Miller loop. Below ~2^11
// pairs one warp owns a pair: f^2 and f * line are the cooperative Fq12 operations above, the doubling step runs its
// five + six independent Fq2 products on parallel lanes, and only the six addition steps of the loop are serial.
struct WMiller {
  Fq12 f, line;      // line = (l0, 0, 0) + (l3, l4, 0) w in tower slots 0, 3, 4; slots 1, 2, 5 stay zero
  G2Hom r;
  Affine2 q;
  Affine p;
  Fq2 t[6];
  WScratch w;
};

// ark `double_in_place` (see g2_double_line) on parallel lanes; updates s->r and the line slots of s->line
__device__ __noinline__ void w_double_step(WMiller* s) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane < 5) {
    Fq2 o;
    if (lane == 0) {                       // a = x y / 2
      fq2_mul_ol(&o, &s->r.x, &s->r.y);
      fq_halve(o.c0, o.c0);
      fq_halve(o.c1, o.c1);
      s->t[0] = o;
    } else if (lane == 1) {                // b = y^2
      fq2_sqr_ol(&o, &s->r.y);
      s->t[1] = o;
    } else if (lane == 2) {                // c = z^2, e = B' 3 c
      Fq2 c, t3;
      fq2_sqr_ol(&c, &s->r.z);
      fq2_dbl(t3, c);
      fq2_add(t3, t3, c);
      fq2_mul_twist_b(o, t3);
      s->t[2] = c;
      s->t[3] = o;
    } else if (lane == 3) {                // (y + z)^2
      Fq2 yz;
      fq2_add(yz, s->r.y, s->r.z);
      fq2_sqr_ol(&o, &yz);
      s->t[4] = o;
    } else {                               // j = x^2
      fq2_sqr_ol(&o, &s->r.x);
      s->t[5] = o;
    }
  }
  __syncwarp();
  Fq2 out;
  if (lane < 6) {
    const Fq2 b = s->t[1], e = s->t[3];
    if (lane == 0) {                       // x' = a (b - 3 e)
      Fq2 f3, d;
      fq2_dbl(f3, e);
      fq2_add(f3, f3, e);
      fq2_sub(d, b, f3);
      fq2_mul_ol(&out, &s->t[0], &d);
    } else if (lane == 1) {                // y' = ((b + 3 e) / 2)^2 - 3 e^2
      Fq2 f3, g, e2, t3;
      fq2_dbl(f3, e);
      fq2_add(f3, f3, e);
      fq2_add(g, b, f3);
      fq_halve(g.c0, g.c0);
      fq_halve(g.c1, g.c1);
      fq2_sqr_ol(&g, &g);
      fq2_sqr_ol(&e2, &e);
      fq2_dbl(t3, e2);
      fq2_add(t3, t3, e2);
      fq2_sub(out, g, t3);
    } else if (lane == 2 || lane == 3) {   // h = (y + z)^2 - (b + c);  z' = b h;  l0 = -h py
      Fq2 h, bc;
      fq2_add(bc, b, s->t[2]);
      fq2_sub(h, s->t[4], bc);
      if (lane == 2) fq2_mul_ol(&out, &b, &h);
      else {
        fq2_neg(h, h);
        fq2_scale(out, h, s->p.y);
      }
    } else if (lane == 4) {                // l3 = 3 j px
      Fq2 j3;
      fq2_dbl(j3, s->t[5]);
      fq2_add(j3, j3, s->t[5]);
      fq2_scale(out, j3, s->p.x);
    } else {                               // l4 = i = e - b
      fq2_sub(out, e, b);
    }
  }
  __syncwarp();
  if (lane == 0) s->r.x = out;
  else if (lane == 1) s->r.y = out;
  else if (lane == 2) s->r.z = out;
  else if (lane == 3) *w12_c(&s->line, 0) = out;
  else if (lane == 4) *w12_c(&s->line, 3) = out;
  else if (lane == 5) *w12_c(&s->line, 4) = out;
  __syncwarp();
}

// s->p, s->q loaded; result in s->f
__device__ __noinline__ void w_miller_loop(WMiller* s) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane < 6) {
    *w12_c(&s->f, lane) = lane == 0 ? fq2_one() : fq2_zero();
    *w12_c(&s->line, lane) = fq2_zero();
  }
  if (lane == 6) {
    s->r.x = s->q.x;
    s->r.y = s->q.y;
    s->r.z = fq2_one();
  }
  __syncwarp();
  if (affine_is_inf(s->p) || affine2_is_inf(s->q)) return;   // uniform across the warp
  for (int bit = 62; bit >= 0; bit--) {
    if (bit != 62) w12_sqr(&s->f, &s->f, &s->w);
    w_double_step(s);
    w12_mul(&s->f, &s->f, &s->line, &s->w);
    if ((BLS_X >> bit) & 1) {
      if (lane == 0) {
        Fq2 l0, l3, l4;
        G2Hom r = s->r;
        g2_add_line(r, l0, l3, l4, s->q, s->p.x, s->p.y);
        s->r = r;
        *w12_c(&s->line, 0) = l0;
        *w12_c(&s->line, 3) = l3;
        *w12_c(&s->line, 4) = l4;
      }
      __syncwarp();
      w12_mul(&s->f, &s->f, &s->line, &s->w);
    }
  }
}

// f[j] = Miller(g1[j], g2[j ^ xor_mask]), one warp (= one block) per pair
__global__ void __launch_bounds__(32) k_miller_coop(const uint4* __restrict__ g1, const uint4* __restrict__ g2,
                                                    uint32_t xor_mask, uint4* __restrict__ f_out) {
  __shared__ WMiller s;
  const int lane = threadIdx.x;
  const uint32_t j = blockIdx.x;
  uint4* p4 = reinterpret_cast<uint4*>(&s.p);
  uint4* q4 = reinterpret_cast<uint4*>(&s.q);
  if (lane < 6) p4[lane] = g1[6 * (size_t)j + lane];
  if (lane >= 8 && lane < 20) q4[lane - 8] = g2[12 * (size_t)(j ^ xor_mask) + (lane - 8)];
  __syncwarp();
  w_miller_loop(&s);
  __syncwarp();
  const uint4* f4 = reinterpret_cast<const uint4*>(&s.f);
  for (int i = lane; i < 36; i += 32) f_out[36 * (size_t)j + i] = f4[i];
}

// product tree level, one warp per output: out[s][t] = prod_k in[s][t + k m]
__global__ void __launch_bounds__(32) k_fq12_prod_level_coop(const uint4* __restrict__ in, uint32_t len, uint32_t m,
                                                             uint4* __restrict__ out) {
  __shared__ Fq12 acc, x;
  __shared__ WScratch w;
  const int lane = threadIdx.x;
  const uint32_t t = blockIdx.x;
  const uint4* src = in + 36 * (size_t)blockIdx.y * len;
  uint4* a4 = reinterpret_cast<uint4*>(&acc);
  uint4* x4 = reinterpret_cast<uint4*>(&x);
  for (int i = lane; i < 36; i += 32) a4[i] = src[36 * (size_t)t + i];
  for (int k = 1; k < FQ12_FAN; k++) {
    const uint64_t idx = (uint64_t)t + (uint64_t)k * m;
    if (idx >= len) break;
    __syncwarp();
    for (int i = lane; i < 36; i += 32) x4[i] = src[36 * idx + i];
    __syncwarp();
    w12_mul(&acc, &acc, &x, &w);
  }
  __syncwarp();
  for (int i = lane; i < 36; i += 32) out[36 * ((size_t)blockIdx.y * m + t) + i] = a4[i];
}

// out[b] = final_exponentiation(in[b]); one warp per product
__global__ void __launch_bounds__(32) k_final_exp(const uint4* __restrict__ in, uint4* __restrict__ out) {
  __shared__ WFinalExp s;
  const int lane = threadIdx.x;
  uint4* f4 = reinterpret_cast<uint4*>(&s.f);
  for (int i = lane; i < 36; i += 32) f4[i] = in[36 * (size_t)blockIdx.x + i];
  __syncwarp();
  w12_final_exp(&s);
  const uint4* r4 = reinterpret_cast<const uint4*>(&s.r);
  for (int i = lane; i < 36; i += 32) out[36 * (size_t)blockIdx.x + i] = r4[i];
}

// out[b] = in[b] with no pairs at all (n == 0): the empty product
__global__ void k_fq12_set_one(uint4* out, uint32_t count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  Fq12 one = fq12_one();
  store_fq12(out + 36 * (size_t)t, one);
}

// a^e for GT elements (the verifier's `tx.pow(c)`, src/mipp.rs:252-255): square-and-multiply over a canonical
// 8-limb exponent, one thread per (element, exponent) pair
__global__ void __launch_bounds__(32) k_fq12_pow(const uint4* __restrict__ in, const uint32_t* __restrict__ exps,
                                                 uint32_t n, int exps_mont, uint4* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  Fq12 a, acc = fq12_one();
  load_fq12(a, in + 36 * (size_t)t);
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = exps[8 * (size_t)t + i];
  if (exps_mont) mont_to_canonical<FrParams>(e, e);
  bool started = false;
  for (int i = 7; i >= 0; i--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) fq12_sqr_ol(&acc, &acc);
      if ((e[i] >> bit) & 1) {
        if (started) fq12_mul_ol(&acc, &acc, &a);
        else {
          acc = a;
          started = true;
        }
      }
    }
  }
  store_fq12(out + 36 * (size_t)t, acc);
}

// ---- MIPP `compress` with the curve endomorphisms ---------------------------------------------------------------------
// h_l + c_inv * h_r over 253-bit c_inv is a 253-step doubling chain per element and was the longest stage of a MIPP round
// (12 ms for ONE element). On the order-r subgroup the twisted Frobenius psi(x, y) = (conj(x) g_2, conj(y) g_3),
// g_i = u^(i (q-1)/6), acts as multiplication by x (the curve parameter; q = x mod r), and r < x^4: writing
// k = k0 + k1 x + k2 x^2 + k3 x^3 with 0 <= k_i < x < 2^64 turns k P into a 4-way simultaneous multiplication with 64
// doublings. PRECONDITION: the points lie in G2 (true for every CRS / folded-CRS vector the reference passes,
// src/mipp.rs:43,114) -- outside the subgroup psi(P) != [x] P.
// G1: phi(x, y) = (beta x, y) is multiplication by lambda = x^2 - 1 (127-bit halves, 127 doublings).

// digits[0..3] = base-x digits of the canonical scalar (u64 each, as 8 u32 words); one thread
__global__ void k_glv4_digits(const uint32_t* __restrict__ scaler, int mont, uint32_t* __restrict__ digits) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t k[8];
  for (int j = 0; j < 8; j++) k[j] = scaler[j];
  if (mont) mont_to_canonical<FrParams>(k, k);
  uint64_t q[4];
  for (int j = 0; j < 4; j++) q[j] = (uint64_t)k[2 * j] | ((uint64_t)k[2 * j + 1] << 32);
  for (int d = 0; d < 4; d++) {
    uint64_t rem = 0, nq[4] = {0, 0, 0, 0};
    for (int bit = 255; bit >= 0; bit--) {
      const uint64_t top = rem >> 63;
      rem = (rem << 1) | ((q[bit >> 6] >> (bit & 63)) & 1);
      if (top || rem >= BLS_X) {
        rem -= BLS_X;
        nq[bit >> 6] |= 1ull << (bit & 63);
      }
    }
    digits[2 * d] = (uint32_t)rem;
    digits[2 * d + 1] = (uint32_t)(rem >> 32);
    for (int j = 0; j < 4; j++) q[j] = nq[j];
  }
}

__device__ __forceinline__ void g2_psi(Affine2& r, const Affine2& p) {
  const Fq g2c = fq_from_table(FQ12_C(FROB1)[1][0]);   // u^(2 (q-1)/6), in Fq
  const Fq g3c = fq_from_table(FQ12_C(FROB1)[2][0]);   // u^(3 (q-1)/6), in Fq
  fq2_conj(r.x, p.x);
  fq2_scale(r.x, r.x, g2c);
  fq2_conj(r.y, p.y);
  fq2_scale(r.y, r.y, g3c);
}

// a[i] <- a[i] + k * a[split + i], k given by its four base-x digits
__global__ void __launch_bounds__(64) k_compress_g2_glv(uint4* __restrict__ a, uint32_t split,
                                                        const uint32_t* __restrict__ digits) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= split) return;
  uint64_t kd[4];
#pragma unroll
  for (int j = 0; j < 4; j++) kd[j] = (uint64_t)digits[2 * j] | ((uint64_t)digits[2 * j + 1] << 32);
  Affine2 l, b[4];
  load_affine2(l, a + 12 * (uint64_t)i);
  load_affine2(b[0], a + 12 * ((uint64_t)split + i));
  for (int j = 1; j < 4; j++) g2_psi(b[j], b[j - 1]);
  Xyzz2 T[16];                       // T[m] = sum over the set bits j of m of psi^j(r)
  xyzz2_set_inf(T[0]);
  for (int m = 1; m < 16; m++) {
    const int low = __ffs(m) - 1;
    T[m] = T[m & (m - 1)];
    xyzz2_madd_ni(&T[m], &b[low]);
  }
  Xyzz2 acc;
  xyzz2_set_inf(acc);
  bool started = false;
  for (int bit = 63; bit >= 0; bit--) {
    if (started) xyzz2_dbl_ni(&acc);
    const int m = (int)((kd[0] >> bit) & 1) | ((int)((kd[1] >> bit) & 1) << 1) | ((int)((kd[2] >> bit) & 1) << 2) |
                  ((int)((kd[3] >> bit) & 1) << 3);
    if (m) {
      xyzz2_add_ni(&acc, &T[m]);
      started = true;
    }
  }
  xyzz2_madd_ni(&acc, &l);
  Affine2 o;
  xyzz2_to_affine_ni(&o, &acc);
  store_affine2(a + 12 * (uint64_t)i, o);
}

// test hook: one Fq12 operation per thread (tests/test_gpu_pairing.py drives every op against the oracle)
//   0 mul(a,b)  1 sqr(a)  2 inv(a)  3 frobenius(a,1)  4 frobenius(a,2)  5 cyclotomic_sqr(a)  6 exp_by_x(a)
//   7 final_exp(a)  8 mul_by_034(a; b = l0 || l3 || l4)  9 miller(a = G1 affine || G2 affine)
//   20 w12_mul  21 w12_sqr  22 w12_cyclotomic_sqr  23 w12_final_exp  24 w12_frobenius(1)  25 w12_frobenius(2)
//   (warp-cooperative versions: element i is processed by the whole warp of block i / launched with n blocks)
__global__ void __launch_bounds__(32) k_test_w12_op(int op, const uint4* a, const uint4* b, uint4* out) {
  __shared__ WFinalExp s;
  const int lane = threadIdx.x;
  uint4* f4 = reinterpret_cast<uint4*>(&s.f);
  uint4* g4 = reinterpret_cast<uint4*>(&s.f2);
  for (int i = lane; i < 36; i += 32) {
    f4[i] = a[36 * (size_t)blockIdx.x + i];
    g4[i] = b[36 * (size_t)blockIdx.x + i];
  }
  __syncwarp();
  switch (op) {
    case 20: w12_mul(&s.r, &s.f, &s.f2, &s.w); break;
    case 21: w12_sqr(&s.r, &s.f, &s.w); break;
    case 22: w12_cyclotomic_sqr(&s.r, &s.f, &s.w); break;
    case 23: w12_final_exp(&s); break;
    case 24: w12_frobenius(&s.r, &s.f, 1); break;
    case 25: w12_frobenius(&s.r, &s.f, 2); break;
    case 26: w12_mul(&s.f, &s.f, &s.f, &s.w); w12_copy(&s.r, &s.f); break;     // aliasing
    case 28: {                                                                  // cooperative Miller loop: a = G1 || G2
      __shared__ WMiller ms;
      uint4* p4 = reinterpret_cast<uint4*>(&ms.p);
      uint4* q4 = reinterpret_cast<uint4*>(&ms.q);
      if (lane < 6) p4[lane] = a[36 * (size_t)blockIdx.x + lane];
      if (lane >= 8 && lane < 20) q4[lane - 8] = a[36 * (size_t)blockIdx.x + 6 + (lane - 8)];
      __syncwarp();
      w_miller_loop(&ms);
      w12_copy(&s.r, &ms.f);
      break;
    }
    case 27: w12_cyclotomic_sqr(&s.f, &s.f, &s.w); w12_conj(&s.r, &s.f); break;
    default: break;
  }
  __syncwarp();
  const uint4* r4 = reinterpret_cast<const uint4*>(&s.r);
  for (int i = lane; i < 36; i += 32) out[36 * (size_t)blockIdx.x + i] = r4[i];
}

__global__ void __launch_bounds__(32) k_test_fq12_op(int op, const uint4* a, const uint4* b, uint32_t n, uint4* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq12 x, y, r;
  load_fq12(x, a + 36 * (size_t)i);
  load_fq12(y, b + 36 * (size_t)i);
  switch (op) {
    case 0: fq12_mul(r, x, y); break;
    case 1: fq12_sqr(r, x); break;
    case 2: fq12_inv(r, x); break;
    case 3: fq12_frobenius(r, x, 1); break;
    case 4: fq12_frobenius(r, x, 2); break;
    case 5: fq12_cyclotomic_sqr_ol(&r, &x); break;
    case 6: fq12_exp_by_x(r, x); break;
    case 7: fq12_final_exp(r, x); break;
    case 8: r = x; fq12_mul_by_034_ol(&r, &y.c0.c0, &y.c0.c1, &y.c0.c2); break;
    case 9: {
      Affine p;
      Affine2 q;
      load_affine(p, a + 36 * (size_t)i);
      load_affine2(q, a + 36 * (size_t)i + 6);
      miller_loop(r, p, q);
      break;
    }
    default: fq12_final_exp(r, x, op - 100);
  }
  store_fq12(out + 36 * (size_t)i, r);
}

}  // namespace tb
