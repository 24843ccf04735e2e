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
#include "fq12_coop.cuh"
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

// ---- block-cooperative Fq12 arithmetic (a team of lanes, one Fq product per lane) ------------------------------------------
// A final exponentiation is ONE dependent chain of ~315 cyclotomic squarings and ~45 Fq12 products; run by a single
// thread it took 20 ms (31 idle lanes, every product in sequence). Here a team of lanes owns the chain: the Fq12 values
// live in shared memory and the 54 (36, 18) independent Fq products inside every Karatsuba Fq12 product (complex squaring,
// Granger-Scott squaring) run on 54 lanes at once -- one 276-MAC Montgomery product per lane; the linear recombinations
// are spread the same way, one Fq coefficient per lane. The per-item bodies, with LAZY field reduction, are in
// fq12_coop.cuh (round 2: the canonical linear operations were ~3/4 of the instructions on the critical lane); this file
// holds the drivers: `for (t = lane; t < items; t += team) body(t); barrier`.
// Teams: 64 lanes = a whole two-warp CTA (__syncthreads), 32 lanes = a one-warp CTA (__syncwarp: the 54 / 36-item phases
// take two passes; 15-20 % slower per pair but twice as many pairs are resident, so it wins when the Miller stage is
// throughput-bound), or -- in the pipelined Miller kernel -- warps 0-1 of a three-warp CTA (named barrier 1) next to a
// one-warp team on warp 2.
// Every function must be called by all lanes of the team. Pointers are shared-memory objects; dst may alias the inputs.
constexpr int W12_THREADS = 64;   // the largest team
struct Team {
  int tid, size, mode;            // mode 0: __syncthreads, 1: __syncwarp, 2: bar.sync 1, 64
  // Item index of the LINEAR phases. Their items are numbered (coefficient, c) with c in the lowest bit, and the c = 0 and
  // c = 1 recombinations are different code (lanes of one warp that diverge run one after the other): a two-warp team puts
  // the c = 0 items on warp 0 and the c = 1 items on warp 1, which halves the number of paths each warp walks through.
  int lin;
};
__device__ __forceinline__ int team_lin(int tid, int size) { return size == 64 ? (((tid & 31) << 1) | (tid >> 5)) : tid; }
__device__ __forceinline__ Team team_cta() {
  Team t;
  t.tid = threadIdx.x;
  t.size = blockDim.x;
  t.mode = blockDim.x > 32 ? 0 : 1;
  t.lin = team_lin(t.tid, t.size);
  return t;
}
__device__ __forceinline__ void team_sync(const Team& t) {
  if (t.mode == 0) __syncthreads();
  else if (t.mode == 1) __syncwarp();
  else asm volatile("bar.sync 1, 64;" ::: "memory");
}

// phases 1-3 of every product: 18 * count lanes multiply, 12 * count lanes assemble the Fq2 products, 6 * count lanes
// assemble (and reduce) coefficient (t, c) of Fq6 result i. `count` = number of Fq6 products (1..3).
static __device__ __noinline__ void w_fq6_products(WScratch* w, int count, Team tm) {
  team_sync(tm);
  for (int t = tm.tid; t < 18 * count; t += tm.size) wp_kar(w, t);
  team_sync(tm);
  for (int t = tm.lin; t < 12 * count; t += tm.size) wp_fq2(w, t);
  team_sync(tm);
  if (tm.lin < 6 * count) wp_fq6(w, tm.lin);
  team_sync(tm);
}
static __device__ __noinline__ void w12_mul(Fq12* dst, const Fq12* a, const Fq12* b, WScratch* w, Team tm) {
  team_sync(tm);
  for (int t = tm.lin; t < 36; t += tm.size) wp_mul_xy(w, a, b, t);
  w_fq6_products(w, 3, tm);
  if (tm.lin < 12) wp_mul_out(dst, w, tm.lin);
  team_sync(tm);
}
static __device__ __noinline__ void w12_sqr(Fq12* dst, const Fq12* a, WScratch* w, Team tm) {
  team_sync(tm);
  if (tm.lin < 24) wp_sqr_xy(w, a, tm.lin);
  w_fq6_products(w, 2, tm);
  if (tm.lin < 12) wp_sqr_out(dst, w, tm.lin);
  team_sync(tm);
}
static __device__ __noinline__ void w12_cyclotomic_sqr(Fq12* dst, const Fq12* a, WScratch* w, Team tm) {
  team_sync(tm);
  if (tm.tid < 18) wp_cyc_kar(w, a, tm.tid);
  team_sync(tm);
  if (tm.lin < 12) wp_cyc_fq2(w, tm.lin);
  team_sync(tm);
  if (tm.lin < 12) wp_cyc_out(dst, a, w, tm.lin);
  team_sync(tm);
}
__device__ __forceinline__ void w12_copy(Fq12* dst, const Fq12* a, Team tm) {
  team_sync(tm);
  Fq v;
  if (tm.tid < 12) v = w12_q(a)[tm.tid];
  team_sync(tm);
  if (tm.tid < 12) w12_q(dst)[tm.tid] = v;
  team_sync(tm);
}
__device__ __forceinline__ void w12_conj(Fq12* dst, const Fq12* a, Team tm) {
  team_sync(tm);
  Fq v;
  if (tm.lin < 12) v = wp_conj(a, tm.lin);
  team_sync(tm);
  if (tm.lin < 12) w12_q(dst)[tm.lin] = v;
  team_sync(tm);
}
static __device__ __noinline__ void w12_frobenius(Fq12* dst, const Fq12* a, int k, Team tm) {
  team_sync(tm);
  Fq c;
  if (tm.lin < 12) c = wp_frobenius(a, k, tm.lin);
  team_sync(tm);
  if (tm.lin < 12) w12_q(dst)[tm.lin] = c;
  team_sync(tm);
}
// every coefficient to [0, q): values that leave the kernel
__device__ __forceinline__ void w12_canon(Fq12* a, Team tm) {
  team_sync(tm);
  if (tm.tid < 12) {
    Fq v = w12_q(a)[tm.tid];
    lz_canon(v);
    w12_q(a)[tm.tid] = v;
  }
  team_sync(tm);
}
// out = (N^-1, 0) for the Fq6 element N in tower slots 0..2 of n12 (fq12_coop.cuh); out must not alias n12
static __device__ __noinline__ void w12_inv_fq6(Fq12* out, const Fq12* n12, WScratch* w, Team tm) {
  team_sync(tm);
  if (tm.tid < 18) wp_inv6_r1(w, n12, tm.tid);
  team_sync(tm);
  if (tm.lin < 12) wp_inv6_p1(w, tm.lin);
  team_sync(tm);
  if (tm.lin < 6) wp_inv6_p2(w, tm.lin);
  team_sync(tm);
  if (tm.tid < 9) wp_inv6_r2(w, n12, tm.tid);
  team_sync(tm);
  if (tm.tid == 0) wp_inv6_d(w);
  team_sync(tm);
  if (tm.tid < 9) wp_inv6_r3(w, tm.tid);
  team_sync(tm);
  if (tm.lin < 12) wp_inv6_out(out, w, tm.lin);
  team_sync(tm);
}
static __device__ __noinline__ void w12_exp_by_x(Fq12* dst, const Fq12* a, Fq12* acc, WScratch* w, Team tm) {
  w12_copy(acc, a, tm);
  for (int bit = 62; bit >= 0; bit--) {
    w12_cyclotomic_sqr(acc, acc, w, tm);
    if ((BLS_X >> bit) & 1) w12_mul(acc, acc, a, w, tm);
  }
  w12_copy(dst, acc, tm);
}

struct WFinalExp {
  Fq12 f, r, f2, y0, y1, y2, acc;
  WScratch w;
};

// the chain of fq12_final_exp_ol (ark `final_exponentiation`), one team; s->f holds the (canonical) input, the canonical
// result lands in s->r
static __device__ __noinline__ void w12_final_exp(WFinalExp* s, Team tm) {
  WScratch* w = &s->w;
  w12_conj(&s->r, &s->f, tm);                       // f^(q^6 - 1) = conj(f)^2 / (f conj(f)), the divisor in Fq6
  w12_mul(&s->y0, &s->f, &s->r, w, tm);
  w12_inv_fq6(&s->f2, &s->y0, w, tm);
  w12_sqr(&s->r, &s->r, w, tm);
  w12_mul(&s->r, &s->r, &s->f2, w, tm);
  w12_copy(&s->f2, &s->r, tm);
  w12_frobenius(&s->r, &s->r, 2, tm);
  w12_mul(&s->r, &s->r, &s->f2, w, tm);
  w12_cyclotomic_sqr(&s->y0, &s->r, w, tm);
  w12_exp_by_x(&s->y1, &s->r, &s->acc, w, tm);
  w12_conj(&s->y2, &s->r, tm);
  w12_mul(&s->y1, &s->y1, &s->y2, w, tm);
  w12_exp_by_x(&s->y2, &s->y1, &s->acc, w, tm);
  w12_conj(&s->y1, &s->y1, tm);
  w12_mul(&s->y1, &s->y1, &s->y2, w, tm);
  w12_exp_by_x(&s->y2, &s->y1, &s->acc, w, tm);
  w12_frobenius(&s->y1, &s->y1, 1, tm);
  w12_mul(&s->y1, &s->y1, &s->y2, w, tm);
  w12_mul(&s->r, &s->r, &s->y0, w, tm);
  w12_exp_by_x(&s->y0, &s->y1, &s->acc, w, tm);
  w12_exp_by_x(&s->y2, &s->y0, &s->acc, w, tm);
  w12_frobenius(&s->y0, &s->y1, 2, tm);
  w12_conj(&s->y1, &s->y1, tm);
  w12_mul(&s->y1, &s->y1, &s->y2, w, tm);
  w12_mul(&s->y1, &s->y1, &s->y0, w, tm);
  w12_mul(&s->r, &s->r, &s->y1, w, tm);
  w12_canon(&s->r, tm);
}

// ---- cooperative Miller loop (one team per pair) ----------------------------------------------------------------------------
// Thread-per-pair keeps the integer pipe fed only when there are thousands of pairs; the MIPP rounds of the reference
// issue products of 2^12 ... 1 pairs and every one of them waits for a full 10 ms single-thread Miller loop. Below ~2^11
// pairs a team owns a pair: f^2 and f * line are the cooperative Fq12 operations above, the doubling step runs its
// 11 + 14 independent Fq products on parallel lanes (fq12_coop.cuh), and so do the six addition steps of the loop.
struct WMiller {
  Fq12 f, line;      // line = (l0, 0, 0) + (l3, l4, 0) w in tower slots 0, 3, 4; slots 1, 2, 5 stay zero
  WDouble d;         // the running point r, P's coordinates, the doubling step's scratch
  Affine2 q;
  Affine p;
  WScratch w;
};

// SIMT rule that shaped the doubling step: lanes of a warp that take DIFFERENT branches run one after the other, so a first
// version with one Fq2 product per lane in lane-specific branches took 44 us per step. Every expensive operation (the
// Montgomery products) is ONE uniform call whose operands were selected per lane beforehand; only cheap additions sit in
// lane-specific branches.
static __device__ __noinline__ void w_double_step(WDouble* d, Fq12* line, Team tm) {
  team_sync(tm);
  if (tm.tid < 11) wp_dbl_r1(d, tm.tid);
  team_sync(tm);
  if (tm.lin < 12) wp_dbl_p2(d, tm.lin);
  team_sync(tm);
  if (tm.lin < 10) wp_dbl_p3(d, tm.lin);
  team_sync(tm);
  if (tm.tid < 14) wp_dbl_r2(d, tm.tid);
  team_sync(tm);
  if (tm.lin < 12) wp_dbl_p5(d, line, tm.lin);
  team_sync(tm);
}
// the addition step (six per loop), four product rounds on parallel lanes (fq12_coop.cuh); on one thread it was ~30
// dependent Fq products, ~60 us of a 1.15 ms loop each
static __device__ __noinline__ void w_add_step(WDouble* d, const Affine2* q, Fq12* line, Team tm) {
  team_sync(tm);
  if (tm.tid < 6) wp_add_rA(d, q, tm.tid);
  team_sync(tm);
  if (tm.lin < 4) wp_add_pA(d, tm.lin);
  team_sync(tm);
  if (tm.tid < 14) wp_add_rB(d, q, tm.tid);
  team_sync(tm);
  if (tm.lin < 10) wp_add_pB(d, line, tm.lin);
  team_sync(tm);
  if (tm.tid < 9) wp_add_rC(d, tm.tid);
  team_sync(tm);
  if (tm.lin < 6) wp_add_pC(d, tm.lin);
  team_sync(tm);
  if (tm.tid < 12) wp_add_rD(d, tm.tid);
  team_sync(tm);
  if (tm.lin < 6) wp_add_pD(d, tm.lin);
  team_sync(tm);
}

// s->p, s->q loaded; result in s->f (canonical)
static __device__ __noinline__ void w_miller_loop(WMiller* s, Team tm) {
  team_sync(tm);
  if (tm.tid < 6) {
    *w12_c(&s->f, tm.tid) = tm.tid == 0 ? fq2_one() : fq2_zero();
    *w12_c(&s->line, tm.tid) = fq2_zero();
  }
  if (tm.tid == 6) {
    s->d.r.x = s->q.x;
    s->d.r.y = s->q.y;
    s->d.r.z = fq2_one();
    s->d.px = s->p.x;
    s->d.py = s->p.y;
  }
  team_sync(tm);
  if (affine_is_inf(s->p) || affine2_is_inf(s->q)) return;   // uniform across the team
  for (int bit = 62; bit >= 0; bit--) {
    if (bit != 62) w12_sqr(&s->f, &s->f, &s->w, tm);
    w_double_step(&s->d, &s->line, tm);
    w12_mul(&s->f, &s->f, &s->line, &s->w, tm);
    if ((BLS_X >> bit) & 1) {
      w_add_step(&s->d, &s->q, &s->line, tm);
      w12_mul(&s->f, &s->f, &s->line, &s->w, tm);
    }
  }
  w12_canon(&s->f, tm);
}

// f[j] = Miller(g1[j], g2[j ^ xor_mask]), one CTA (64 or 32 threads) per pair
__global__ void __launch_bounds__(W12_THREADS, 8) k_miller_coop(const uint4* __restrict__ g1, const uint4* __restrict__ g2,
                                                    uint32_t xor_mask, uint4* __restrict__ f_out) {
  __shared__ WMiller s;
  const Team tm = team_cta();
  const int lane = threadIdx.x;
  const uint32_t j = blockIdx.x;
  uint4* p4 = reinterpret_cast<uint4*>(&s.p);
  uint4* q4 = reinterpret_cast<uint4*>(&s.q);
  if (lane < 6) p4[lane] = g1[6 * (size_t)j + lane];
  if (lane >= 8 && lane < 20) q4[lane - 8] = g2[12 * (size_t)(j ^ xor_mask) + (lane - 8)];
  team_sync(tm);
  w_miller_loop(&s, tm);
  team_sync(tm);
  const uint4* f4 = reinterpret_cast<const uint4*>(&s.f);
  for (int i = lane; i < 36; i += blockDim.x) f_out[36 * (size_t)j + i] = f4[i];
}

// ---- two pairs per warp with one shared accumulator -----------------------------------------------------------------------
// prod_j f_j = the Miller value of a product: pairs may share the accumulator, f <- f^2 * line_a * line_b (ark's
// multi_miller_loop does the same over chunks of four). In the throughput regime (> 512 pairs) what counts is the number
// of product passes and linear phases a warp issues per pair (ncu, profiles/r02_summary.md: the integer pipe is 47 %
// busy with 11 of 32 lanes active on average). One warp, two pairs:
//   f^2 once (2 passes)  |  BOTH doubling steps in the same three passes (lanes 0-15 pair a, 16-31 pair b)  |
//   line_a * line_b as one sparse product (1 pass, 18 Fq products)  |  f * L (2 passes)
// = 8 passes per iteration for two pairs where the one-pair kernel needs 7 per pair; the linear phases shrink the same way.
// An identity point or a missing second pair (odd count) leaves that pair's line at 1.
struct WMillerDuo {
  Fq12 f, L;
  Fq12 line[2];
  WDouble d[2];
  Affine2 q[2];
  Affine p[2];
  WScratch w;
};
// L = la * lb for line-shaped la, lb (fq12_coop.cuh); L must not alias them
static __device__ __noinline__ void w12_line_mul(Fq12* L, const Fq12* la, const Fq12* lb, WScratch* w, Team tm) {
  team_sync(tm);
  if (tm.lin < 12) wp_ll_xy(w, la, lb, tm.lin);
  team_sync(tm);
  if (tm.tid < 18) wp_kar(w, tm.tid);
  team_sync(tm);
  if (tm.lin < 12) wp_fq2(w, tm.lin);
  team_sync(tm);
  if (tm.lin < 12) wp_ll_out(L, w, tm.lin);
  team_sync(tm);
}
// the doubling / addition step of both pairs at once: lane = 16 * pair + item; `on` = this lane's pair is live
static __device__ __noinline__ void w_double_step2(WDouble* d, Fq12* line, bool on) {
  const int t = threadIdx.x & 15;
  __syncwarp();
  if (on && t < 11) wp_dbl_r1(d, t);
  __syncwarp();
  if (on && t < 12) wp_dbl_p2(d, t);
  __syncwarp();
  if (on && t < 10) wp_dbl_p3(d, t);
  __syncwarp();
  if (on && t < 14) wp_dbl_r2(d, t);
  __syncwarp();
  if (on && t < 12) wp_dbl_p5(d, line, t);
  __syncwarp();
}
static __device__ __noinline__ void w_add_step2(WDouble* d, const Affine2* q, Fq12* line, bool on) {
  const int t = threadIdx.x & 15;
  __syncwarp();
  if (on && t < 6) wp_add_rA(d, q, t);
  __syncwarp();
  if (on && t < 4) wp_add_pA(d, t);
  __syncwarp();
  if (on && t < 14) wp_add_rB(d, q, t);
  __syncwarp();
  if (on && t < 10) wp_add_pB(d, line, t);
  __syncwarp();
  if (on && t < 9) wp_add_rC(d, t);
  __syncwarp();
  if (on && t < 6) wp_add_pC(d, t);
  __syncwarp();
  if (on && t < 12) wp_add_rD(d, t);
  __syncwarp();
  if (on && t < 6) wp_add_pD(d, t);
  __syncwarp();
}
// f_out[b] = Miller(pair 2b) * Miller(pair 2b + 1); n pairs, ceil(n / 2) one-warp CTAs. Pair j takes g2[j ^ xor_mask].
__global__ void __launch_bounds__(32, 16) k_miller_duo(const uint4* __restrict__ g1, const uint4* __restrict__ g2, uint32_t n,
                                                       uint32_t xor_mask, uint4* __restrict__ f_out) {
  __shared__ WMillerDuo s;
  Team tm;
  tm.tid = threadIdx.x;
  tm.size = 32;
  tm.mode = 1;
  tm.lin = threadIdx.x;
  const int lane = threadIdx.x, pi = lane >> 4, t = lane & 15;
  const uint32_t j = 2 * blockIdx.x + pi;
  const bool present = j < n;
  uint4* p4 = reinterpret_cast<uint4*>(&s.p[pi]);
  uint4* q4 = reinterpret_cast<uint4*>(&s.q[pi]);
  if (present) {
    if (t < 6) p4[t] = g1[6 * (size_t)j + t];
    if (t < 12) q4[t] = g2[12 * (size_t)(j ^ xor_mask) + t];
  }
  if (t < 6) {
    *w12_c(&s.line[pi], t) = t == 0 ? fq2_one() : fq2_zero();
    if (pi == 0) *w12_c(&s.f, t) = t == 0 ? fq2_one() : fq2_zero();
  }
  __syncwarp();
  const bool on = present && !(affine_is_inf(s.p[pi]) || affine2_is_inf(s.q[pi]));
  WDouble* d = &s.d[pi];
  if (on && t == 0) {
    d->r.x = s.q[pi].x;
    d->r.y = s.q[pi].y;
    d->r.z = fq2_one();
    d->px = s.p[pi].x;
    d->py = s.p[pi].y;
  }
  __syncwarp();
  for (int bit = 62; bit >= 0; bit--) {
    if (bit != 62) w12_sqr(&s.f, &s.f, &s.w, tm);
    w_double_step2(d, &s.line[pi], on);
    w12_line_mul(&s.L, &s.line[0], &s.line[1], &s.w, tm);
    w12_mul(&s.f, &s.f, &s.L, &s.w, tm);
    if ((BLS_X >> bit) & 1) {
      w_add_step2(d, &s.q[pi], &s.line[pi], on);
      w12_line_mul(&s.L, &s.line[0], &s.line[1], &s.w, tm);
      w12_mul(&s.f, &s.f, &s.L, &s.w, tm);
    }
  }
  w12_canon(&s.f, tm);
  const uint4* f4 = reinterpret_cast<const uint4*>(&s.f);
  for (int i = lane; i < 36; i += 32) f_out[36 * (size_t)blockIdx.x + i] = f4[i];
}

// ---- pipelined Miller loop: the point chain and the f chain on different warps ----------------------------------------------
// The line functions depend on (P, Q) only: r -> 2 r (+ Q) never looks at f. In the sequential loop above a round is
// f^2 | doubling step | f * line, one after the other on the same lanes -- and the doubling step (three dependent product
// rounds) is the longest of the three. Here warp 2 of a three-warp CTA runs the point chain on its own (doubling and
// addition steps, one-warp team) and hands the lines to warps 0-1 through a ring of MP_RING slots; warps 0-1 run
// f <- f^2 * line. The two chains overlap completely; the loop takes max(point chain, f chain) instead of their sum.
// Hand-over with named barriers, the producer / consumer pattern of bar.arrive + bar.sync: barrier 2 + b = "slot b is
// full" (warp 2 arrives, warps 0-1 wait), 2 + MP_RING + b = "slot b is free" (warps 0-1 arrive, warp 2 waits); barrier 1
// is the f team's own.
constexpr int MP_RING = 4;
constexpr int MP_THREADS = 96;
constexpr int mp_events() {
  int e = 0;
  for (int bit = 62; bit >= 0; bit--) e += 1 + (int)((BLS_X >> bit) & 1);
  return e;
}
constexpr int MP_EVENTS = mp_events();
struct WMillerPipe {
  Fq12 f;
  Fq12 line[MP_RING];   // slots 1, 2, 5 of every entry stay zero
  WDouble d;
  Affine2 q;
  Affine p;
  WScratch w;
};
__device__ __forceinline__ void mp_bar_sync(int id) { asm volatile("bar.sync %0, 96;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void mp_bar_arrive(int id) {
  __threadfence_block();
  asm volatile("bar.arrive %0, 96;" ::"r"(id) : "memory");
}
__global__ void __launch_bounds__(MP_THREADS, 5) k_miller_pipe(const uint4* __restrict__ g1, const uint4* __restrict__ g2,
                                                               uint32_t xor_mask, uint4* __restrict__ f_out) {
  __shared__ WMillerPipe s;
  const int tid = threadIdx.x;
  const uint32_t j = blockIdx.x;
  uint4* p4 = reinterpret_cast<uint4*>(&s.p);
  uint4* q4 = reinterpret_cast<uint4*>(&s.q);
  if (tid < 6) p4[tid] = g1[6 * (size_t)j + tid];
  if (tid >= 8 && tid < 20) q4[tid - 8] = g2[12 * (size_t)(j ^ xor_mask) + (tid - 8)];
  uint4* l4 = reinterpret_cast<uint4*>(&s.line[0]);
  for (int i = tid; i < 36 * MP_RING; i += MP_THREADS) l4[i] = make_uint4(0, 0, 0, 0);
  if (tid >= 32 && tid < 38) *w12_c(&s.f, tid - 32) = tid == 32 ? fq2_one() : fq2_zero();
  __syncthreads();
  if (affine_is_inf(s.p) || affine2_is_inf(s.q)) {   // uniform across the CTA: the Miller value is 1
    const uint4* f4 = reinterpret_cast<const uint4*>(&s.f);
    for (int i = tid; i < 36; i += MP_THREADS) f_out[36 * (size_t)j + i] = f4[i];
    return;
  }
  if (tid >= 64) {
    // the point chain
    Team tm;
    tm.tid = tid - 64;
    tm.size = 32;
    tm.mode = 1;
    tm.lin = tm.tid;
    if (tm.tid == 0) {
      s.d.r.x = s.q.x;
      s.d.r.y = s.q.y;
      s.d.r.z = fq2_one();
      s.d.px = s.p.x;
      s.d.py = s.p.y;
    }
    __syncwarp();
    int e = 0;
    for (int bit = 62; bit >= 0; bit--) {
      if (e >= MP_RING) mp_bar_sync(2 + MP_RING + e % MP_RING);
      w_double_step(&s.d, &s.line[e % MP_RING], tm);
      mp_bar_arrive(2 + e % MP_RING);
      e++;
      if ((BLS_X >> bit) & 1) {
        if (e >= MP_RING) mp_bar_sync(2 + MP_RING + e % MP_RING);
        w_add_step(&s.d, &s.q, &s.line[e % MP_RING], tm);
        mp_bar_arrive(2 + e % MP_RING);
        e++;
      }
    }
  } else {
    // the f chain
    Team tm;
    tm.tid = tid;
    tm.size = 64;
    tm.mode = 2;
    tm.lin = team_lin(tid, 64);
    int e = 0;
    for (int bit = 62; bit >= 0; bit--) {
      if (bit != 62) w12_sqr(&s.f, &s.f, &s.w, tm);
      for (int step = 0; step <= (int)((BLS_X >> bit) & 1); step++) {
        mp_bar_sync(2 + e % MP_RING);
        w12_mul(&s.f, &s.f, &s.line[e % MP_RING], &s.w, tm);
        if (e + MP_RING < MP_EVENTS) mp_bar_arrive(2 + MP_RING + e % MP_RING);
        e++;
      }
    }
    w12_canon(&s.f, tm);
    const uint4* f4 = reinterpret_cast<const uint4*>(&s.f);
    if (tid < 36) f_out[36 * (size_t)j + tid] = f4[tid];
  }
}

// product tree level, one CTA per output: out[s][t] = prod_k in[s][t + k m]
__global__ void __launch_bounds__(W12_THREADS) k_fq12_prod_level_coop(const uint4* __restrict__ in, uint32_t len, uint32_t m,
                                                             uint4* __restrict__ out) {
  __shared__ Fq12 acc, x;
  __shared__ WScratch w;
  const Team tm = team_cta();
  const int lane = threadIdx.x;
  const uint32_t t = blockIdx.x;
  const uint4* src = in + 36 * (size_t)blockIdx.y * len;
  uint4* a4 = reinterpret_cast<uint4*>(&acc);
  uint4* x4 = reinterpret_cast<uint4*>(&x);
  for (int i = lane; i < 36; i += blockDim.x) a4[i] = src[36 * (size_t)t + i];
  for (int k = 1; k < FQ12_FAN; k++) {
    const uint64_t idx = (uint64_t)t + (uint64_t)k * m;
    if (idx >= len) break;
    team_sync(tm);
    for (int i = lane; i < 36; i += blockDim.x) x4[i] = src[36 * idx + i];
    team_sync(tm);
    w12_mul(&acc, &acc, &x, &w, tm);
  }
  w12_canon(&acc, tm);
  for (int i = lane; i < 36; i += blockDim.x) out[36 * ((size_t)blockIdx.y * m + t) + i] = a4[i];
}

// out[b] = final_exponentiation(in[b]); one CTA per product
__global__ void __launch_bounds__(W12_THREADS) k_final_exp(const uint4* __restrict__ in, uint4* __restrict__ out) {
  __shared__ WFinalExp s;
  const Team tm = team_cta();
  const int lane = threadIdx.x;
  uint4* f4 = reinterpret_cast<uint4*>(&s.f);
  for (int i = lane; i < 36; i += blockDim.x) f4[i] = in[36 * (size_t)blockIdx.x + i];
  team_sync(tm);
  w12_final_exp(&s, tm);
  const uint4* r4 = reinterpret_cast<const uint4*>(&s.r);
  for (int i = lane; i < 36; i += blockDim.x) out[36 * (size_t)blockIdx.x + i] = r4[i];
}

// out[b] = in[b] with no pairs at all (n == 0): the empty product
__global__ void k_fq12_set_one(uint4* out, uint32_t count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  Fq12 one = fq12_one();
  store_fq12(out + 36 * (size_t)t, one);
}

// a^e for GT elements (the verifier's `tx.pow(c)`, src/mipp.rs:258-261): square-and-multiply over a canonical
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

// a[i] <- a[i] + k * a[split + i], k given by its four base-x digits, with the four dimensions on four WARPS of a CTA
// (32 elements per CTA): warp j multiplies psi^j(r) by the 64-bit digit k_j (the digit is warp-uniform, so its
// double-and-add does not diverge), the three partial sums travel through shared memory to warp 0, which adds them, adds
// a[i] and normalises. Per thread 64 doublings + ~32 mixed additions (a first version did all four dimensions in one
// thread with a 15-entry table: 64 doublings + 60 full additions + 11 for the table, 5.3 ms instead of 3.6): the latency of
// a fold is what a MIPP round waits for.
__global__ void __launch_bounds__(128) k_compress_g2_glv4w(uint4* __restrict__ a, uint32_t split,
                                                           const uint32_t* __restrict__ digits) {
  __shared__ Xyzz2 part[3][32];
  const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * 32 + lane;
  const bool live = i < split;
  Xyzz2 acc;
  xyzz2_set_inf(acc);
  if (live) {
    const uint64_t kd = (uint64_t)digits[2 * j] | ((uint64_t)digits[2 * j + 1] << 32);
    Affine2 b;
    load_affine2(b, a + 12 * ((uint64_t)split + i));
    for (int t = 0; t < j; t++) g2_psi(b, b);
    bool started = false;
    for (int bit = 63; bit >= 0; bit--) {
      if (started) xyzz2_dbl_ni(&acc);
      if ((kd >> bit) & 1) {
        xyzz2_madd_ni(&acc, &b);
        started = true;
      }
    }
    if (j > 0) part[j - 1][lane] = acc;
  }
  __syncthreads();
  if (j != 0 || !live) return;
  for (int t = 0; t < 3; t++) xyzz2_add_ni(&acc, &part[t][lane]);
  Affine2 l, o;
  load_affine2(l, a + 12 * (uint64_t)i);
  xyzz2_madd_ni(&acc, &l);
  xyzz2_to_affine_ni(&o, &acc);
  store_affine2(a + 12 * (uint64_t)i, o);
}

// Window combine of a single G2 MSM over the endomorphism. The Horner form sum_w 2^(c w) S_w is a chain of ~250
// doublings on ONE thread (11 ms: the latency floor of every G2 MSM the reference issues -- commit_g2, the PST opening
// proofs). Here thread (w, j) multiplies psi^j(S_w) by the j-th base-x digit of 2^(c w) mod r (64 doublings), and a
// shared-memory tree adds the 4 W partial sums. psi on XYZZ coordinates: (conj X g_2, conj Y g_3, conj ZZ, conj ZZZ).
// One CTA of 4 W threads (W <= 96); the partial sums live in global scratch (4 W x 384 B).
__device__ __forceinline__ void xyzz2_psi(Xyzz2& r, const Xyzz2& p) {
  const Fq g2c = fq_from_table(FQ12_C(FROB1)[1][0]);
  const Fq g3c = fq_from_table(FQ12_C(FROB1)[2][0]);
  fq2_conj(r.x, p.x);
  fq2_scale(r.x, r.x, g2c);
  fq2_conj(r.y, p.y);
  fq2_scale(r.y, r.y, g3c);
  fq2_conj(r.zz, p.zz);
  fq2_conj(r.zzz, p.zzz);
}
__global__ void __launch_bounds__(384) k_finalize_single_g2_glv(const uint4* __restrict__ group_w, int W, int c,
                                                                uint4* __restrict__ scratch, uint4* __restrict__ out_affine) {
  const int t = threadIdx.x, w = t >> 2, j = t & 3;
  const int total = 4 * W;
  if (t < total) {
    // base-x digit j of 2^(c w) mod r  (c w <= 253 < 2 * 253: at most one subtraction of r)
    uint64_t q[4] = {0, 0, 0, 0};
    const int e = c * w;
    q[e >> 6] = 1ull << (e & 63);
    if (e >= 252) {   // r has 253 bits: 2^252 < r < 2^253
      uint64_t rr[4];
      for (int i = 0; i < 4; i++) rr[i] = (uint64_t)FrParams::p(2 * i) | ((uint64_t)FrParams::p(2 * i + 1) << 32);
      bool ge = true;
      for (int i = 3; i >= 0; i--) {
        if (q[i] != rr[i]) {
          ge = q[i] > rr[i];
          break;
        }
      }
      if (ge) {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
          const uint64_t d = q[i] - rr[i] - borrow;
          borrow = (q[i] < rr[i] + borrow) || (rr[i] + borrow < borrow) ? 1 : 0;
          q[i] = d;
        }
      }
    }
    uint64_t digit = 0;
    for (int d = 0; d <= j; d++) {
      uint64_t rem = 0, nq[4] = {0, 0, 0, 0};
      for (int bit = 255; bit >= 0; bit--) {
        const uint64_t top = rem >> 63;
        rem = (rem << 1) | ((q[bit >> 6] >> (bit & 63)) & 1);
        if (top || rem >= BLS_X) {
          rem -= BLS_X;
          nq[bit >> 6] |= 1ull << (bit & 63);
        }
      }
      digit = rem;
      for (int i = 0; i < 4; i++) q[i] = nq[i];
    }
    Xyzz2 b, acc;
    load_xyzz2(b, group_w + 24 * w);
    for (int i = 0; i < j; i++) xyzz2_psi(b, b);
    xyzz2_set_inf(acc);
    bool started = false;
    for (int bit = 63; bit >= 0; bit--) {
      if (started) xyzz2_dbl_ni(&acc);
      if ((digit >> bit) & 1) {
        xyzz2_add_ni(&acc, &b);
        started = true;
      }
    }
    store_xyzz2(scratch + 24 * t, acc);
  }
  // tree sum over the 4 W partial values
  for (int stride = 256; stride >= 1; stride >>= 1) {
    __syncthreads();
    if (t < stride && t + stride < total) {
      Xyzz2 x, y;
      load_xyzz2(x, scratch + 24 * t);
      load_xyzz2(y, scratch + 24 * (t + stride));
      xyzz2_add_ni(&x, &y);
      store_xyzz2(scratch + 24 * t, x);
    }
  }
  __syncthreads();
  if (t == 0) {
    Xyzz2 x;
    load_xyzz2(x, scratch);
    Affine2 a;
    xyzz2_to_affine_ni(&a, &x);
    store_affine2(out_affine, a);
  }
}

// Window combine of a single G2 MSM as a Horner chain on ONE warp with lane-parallel group operations (fq12_coop.cuh:
// wp_g2dbl_*, wp_g2add_*): ~253 doublings of 3 product rounds + W additions of 4. The psi-based kernel above needs only 64
// doublings per thread but runs them with single-thread Fq2 arithmetic (27 us per doubling, 46 us per addition: ~5 ms for
// a 2^13-point MSM); the chain here is four times as long and each link eight times as fast. It also accepts ANY curve
// point (no subgroup precondition).
static __device__ __noinline__ void w_g2_double(WG2* s) {
  const int t = threadIdx.x;
  __syncwarp();
  if (t < 4) wp_g2dbl_r1(s, t);
  __syncwarp();
  if (t < 6) wp_g2dbl_p1(s, t);
  __syncwarp();
  if (t < 11) wp_g2dbl_r2(s, t);
  __syncwarp();
  if (t < 8) wp_g2dbl_p2(s, t);
  __syncwarp();
  if (t < 9) wp_g2dbl_r3(s, t);
  __syncwarp();
  if (t < 4) wp_g2dbl_p3(s, t);
  __syncwarp();
}
static __device__ __noinline__ void w_g2_add(WG2* s) {
  const int t = threadIdx.x;
  __syncwarp();
  if (t == 0) wp_g2add_flags(s);
  __syncwarp();
  if (s->flag == 1) return;                      // uniform: s->flag is shared
  if (s->flag == 2) {
    if (t < 24) reinterpret_cast<uint4*>(&s->p)[t] = reinterpret_cast<const uint4*>(&s->e)[t];
    __syncwarp();
    return;
  }
  if (t < 12) wp_g2add_r1(s, t);
  __syncwarp();
  if (t < 8) wp_g2add_p1(s, t);
  __syncwarp();
  if (t == 0) wp_g2add_check(s);
  __syncwarp();
  if (s->flag == 3) return;
  if (t < 10) wp_g2add_r2(s, t);
  __syncwarp();
  if (t < 8) wp_g2add_p2(s, t);
  __syncwarp();
  if (t < 9) wp_g2add_r3(s, t);
  __syncwarp();
  if (t < 8) wp_g2add_p3(s, t);
  __syncwarp();
  if (t < 9) wp_g2add_r4(s, t);
  __syncwarp();
  if (t < 4) wp_g2add_p4(s, t);
  __syncwarp();
}
__global__ void __launch_bounds__(32) k_finalize_single_g2_coop(const uint4* __restrict__ group_w, int W, int c,
                                                                uint4* __restrict__ out_affine) {
  __shared__ WG2 s;
  const int t = threadIdx.x;
  if (t < 24) reinterpret_cast<uint4*>(&s.p)[t] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  for (int w = W - 1; w >= 0; w--) {
    if (t < 24) reinterpret_cast<uint4*>(&s.e)[t] = group_w[24 * w + t];
    __syncwarp();
    w_g2_add(&s);
    if (w > 0)
      for (int k = 0; k < c; k++) w_g2_double(&s);
  }
  __syncwarp();
  if (t < 8) lz_canon(reinterpret_cast<Fq*>(&s.p)[t]);
  __syncwarp();
  if (t == 0) {
    Affine2 a;
    xyzz2_to_affine_ni(&a, &s.p);
    store_affine2(out_affine, a);
  }
}

// One level of the hierarchical bucket reduction of a G2 MSM (as k_reduce_pass_g2: running sums S, weighted sums W), one
// WARP per output with the lane-parallel group law: these levels hold few elements (sqrt(n)-sized MSMs) and a thread of
// k_reduce_pass_g2 walks through 2 L + log2(ell) dependent group operations of ~50 us each (2.8 of the 6 ms of a
// 2^13-point MSM).
__global__ void __launch_bounds__(32) k_reduce_pass_g2_coop(const uint4* __restrict__ inS, const uint4* __restrict__ inW,
                                                            const uint32_t* __restrict__ bucket_start,
                                                            uint4* __restrict__ outS, uint4* __restrict__ outW, uint32_t L,
                                                            int log2_ell, uint64_t total_out) {
  __shared__ WG2 run, acc;      // run.p = running sum, acc.p = weighted sum
  const int lane = threadIdx.x;
  const uint64_t t = blockIdx.x;
  if (t >= total_out) return;
  uint4* rp = reinterpret_cast<uint4*>(&run.p);
  uint4* re = reinterpret_cast<uint4*>(&run.e);
  uint4* ap = reinterpret_cast<uint4*>(&acc.p);
  uint4* ae = reinterpret_cast<uint4*>(&acc.e);
  if (lane < 24) {
    rp[lane] = make_uint4(0, 0, 0, 0);
    ap[lane] = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  for (int i = (int)L - 1; i >= 0; i--) {
    const uint64_t idx = t * L + i;
    bool empty = false;
    if (bucket_start) empty = bucket_start[idx + 1] == bucket_start[idx];
    if (!empty) {
      if (lane < 24) re[lane] = inS[24 * idx + lane];
      __syncwarp();
      w_g2_add(&run);
    }
    if (i > 0) {
      if (lane < 24) ae[lane] = rp[lane];
      __syncwarp();
      w_g2_add(&acc);
    }
  }
  __syncwarp();
  if (lane < 8) lz_canon(reinterpret_cast<Fq*>(&run.p)[lane]);
  __syncwarp();
  if (lane < 24) outS[24 * t + lane] = rp[lane];
  for (int k = 0; k < log2_ell; k++) w_g2_double(&acc);
  if (inW) {
    if (lane < 24) rp[lane] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    for (int i = 0; i < (int)L; i++) {
      if (lane < 24) re[lane] = inW[24 * (t * L + i) + lane];
      __syncwarp();
      w_g2_add(&run);
    }
  }
  __syncwarp();
  if (lane < 24) ae[lane] = rp[lane];
  __syncwarp();
  w_g2_add(&acc);
  __syncwarp();
  if (lane < 8) lz_canon(reinterpret_cast<Fq*>(&acc.p)[lane]);
  __syncwarp();
  if (lane < 24) outW[24 * t + lane] = ap[lane];
}

// G1 fold over phi(x, y) = (beta x, y) = [lambda](x, y), lambda = x^2 - 1 < 2^127: k = k1 lambda + k0 with k0 < lambda and
// k1 = floor(k / lambda) < 2^127 (r < lambda^2 + lambda + 1), then a 2-way simultaneous multiplication, 127 doublings.
// digits[0..3] = k0, digits[4..7] = k1 (128 bits each). Same precondition: the points lie in G1's order-r subgroup.
__global__ void k_glv2_digits(const uint32_t* __restrict__ scaler, int mont, uint32_t* __restrict__ digits) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t k[8];
  for (int j = 0; j < 8; j++) k[j] = scaler[j];
  if (mont) mont_to_canonical<FrParams>(k, k);
  // lambda = BLS_X^2 - 1 as two 64-bit halves
  const uint64_t xl = BLS_X & 0xffffffffull, xh = BLS_X >> 32;
  const uint64_t ll = xl * xl, lh = xl * xh, hh = xh * xh;     // x^2 = hh 2^64 + 2 lh 2^32 + ll
  uint64_t lo = ll + (lh << 33);
  uint64_t hi = hh + (lh >> 31) + (lo < ll ? 1 : 0);
  hi -= (lo == 0);
  lo -= 1;
  uint64_t q[4], nq[4] = {0, 0, 0, 0}, r0 = 0, r1 = 0;         // remainder r1:r0 < lambda < 2^127
  for (int j = 0; j < 4; j++) q[j] = (uint64_t)k[2 * j] | ((uint64_t)k[2 * j + 1] << 32);
  for (int bit = 255; bit >= 0; bit--) {
    r1 = (r1 << 1) | (r0 >> 63);
    r0 = (r0 << 1) | ((q[bit >> 6] >> (bit & 63)) & 1);
    if (r1 > hi || (r1 == hi && r0 >= lo)) {
      const uint64_t b = r0 < lo;
      r0 -= lo;
      r1 -= hi + b;
      nq[bit >> 6] |= 1ull << (bit & 63);
    }
  }
  digits[0] = (uint32_t)r0;
  digits[1] = (uint32_t)(r0 >> 32);
  digits[2] = (uint32_t)r1;
  digits[3] = (uint32_t)(r1 >> 32);
  digits[4] = (uint32_t)nq[0];
  digits[5] = (uint32_t)(nq[0] >> 32);
  digits[6] = (uint32_t)nq[1];
  digits[7] = (uint32_t)(nq[1] >> 32);     // nq[2], nq[3] are zero for k < r
}

// a[i] <- a[i] + k * a[split + i], k = k0 + k1 lambda
__global__ void __launch_bounds__(128) k_compress_g1_glv(uint4* __restrict__ a, uint32_t split,
                                                         const uint32_t* __restrict__ digits) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= split) return;
  uint32_t k0[4], k1[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    k0[j] = digits[j];
    k1[j] = digits[4 + j];
  }
  Affine l, r, pr;
  load_affine(l, a + 6 * (uint64_t)i);
  load_affine(r, a + 6 * ((uint64_t)split + i));
  const Fq beta = fq_from_table(FQ12_C(G1_BETA));
  fq_mul_ol(&pr.x, &r.x, &beta);           // phi(r); the identity (0, 0) stays (0, 0)
  pr.y = r.y;
  // the lazily reduced group law of the MSM hot loop (g1_fast.cuh: one out-of-line multiplier): with the canonical
  // operations the three inlined group routines are ~350 KB of SASS and the kernel stalled on instruction fetch
  // (ncu: `no_instruction` 2.06 per issue)
  Xyzz t3;                                 // r + phi(r)
  xyzz_set_inf(t3);
  xyzz_madd_fast_ni(&t3, &r);
  xyzz_madd_fast_ni(&t3, &pr);
  Xyzz acc;
  xyzz_set_inf(acc);
  bool started = false;
  for (int limb = 3; limb >= 0; limb--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) xyzz_dbl_fast_ni(&acc);
      const int m = (int)((k0[limb] >> bit) & 1) | ((int)((k1[limb] >> bit) & 1) << 1);
      if (m == 1) xyzz_madd_fast_ni(&acc, &r);
      else if (m == 2) xyzz_madd_fast_ni(&acc, &pr);
      else if (m == 3) xyzz_add_fast_ni(&acc, &t3);
      started = started || m != 0;
    }
  }
  xyzz_madd_fast_ni(&acc, &l);
  xyzz_canon(acc);
  Affine o;
  xyzz_to_affine_ni(&o, &acc);
  store_affine(a + 6 * (uint64_t)i, o);
}

// ---- two-phase folds: the doubling chain leaves the critical path of a MIPP round ----------------------------------------
// A round is   cross values (MSMs, Miller loops, final exponentiations) -> challenge c -> fold -> next round,   and the
// fold used to be a 127-step (G1) / 64-step (G2) doubling chain per element that starts only once c is known: 3.6-4.3 ms
// of a 7 ms round. The multiples 2^j * a_r[i] do NOT depend on c: phase A computes them on a side stream while the
// round's Miller loops run (those are latency-bound and leave the SMs mostly idle); phase B -- after c is known -- adds
// the multiples c selects. c is ONE scalar for the whole vector, so the selection is the same for every element: the
// host decomposes c over the endomorphism (glv_host.h) into a list of (bit position j, endomorphism power d) pairs,
// T lanes share an element and split the list evenly (no lane is predicated off), a shuffle tree combines their partial
// sums, lane 0 adds a_l[i] and normalises. T = 32 for the short vectors of the late rounds (latency: ~4 + 5 additions),
// T = 8 for long ones (throughput: 4 elements per warp). The endomorphisms are applied on the fly:
// phi(X, Y, ZZ, ZZZ) = (beta X, Y, ZZ, ZZZ), psi as xyzz2_psi. Precondition as for the one-phase kernels: the points lie
// in the order-r subgroups.
constexpr int FOLD_G1_STEPS = 127;   // k0, k1 < 2^127
constexpr int FOLD_G2_STEPS = 64;    // base-x digits < 2^64

// phase A, G1: mult[i * 127 + j] = 2^j * a[first + i] as lazily reduced XYZZ (identity stays the identity)
__global__ void __launch_bounds__(128) k_fold_pre_g1(const uint4* __restrict__ a, uint32_t first, uint32_t count,
                                                     uint4* __restrict__ mult) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Affine r;
  load_affine(r, a + 6 * ((uint64_t)first + i));
  Xyzz t;
  xyzz_set_inf(t);
  xyzz_madd_fast_ni(&t, &r);
  uint4* dst = mult + (uint64_t)i * FOLD_G1_STEPS * 12;
  for (int j = 0; j < FOLD_G1_STEPS; j++) {
    store_xyzz(dst + 12 * j, t);
    if (j + 1 < FOLD_G1_STEPS) xyzz_dbl_fast_ni(&t);
  }
}

template <int T>
__device__ __forceinline__ void xyzz_shfl_down(Xyzz& o, const Xyzz& p, int off) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
  uint32_t* d = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int i = 0; i < 48; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], off, T);
}

// phase B, G1: a[i] <- a[i] + (k0 + k1 lambda) * a[split + i] from the stored multiples; T lanes per element.
// sel[0] = number of entries, sel[1..] = j | (d << 8). Launched with ceil(split T / 128) CTAs of 128 threads.
template <int T>
__global__ void __launch_bounds__(128) k_fold_apply_g1(uint4* __restrict__ a, uint32_t split,
                                                       const uint16_t* __restrict__ sel,
                                                       const uint4* __restrict__ mult) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t e = tid / T;
  const int lane = tid % T;
  const bool live = e < split;             // dead lane groups only take part in the shuffles
  const int count = sel[0];
  const Fq beta = fq_from_table(FQ12_C(G1_BETA));
  Xyzz acc;
  xyzz_set_inf(acc);
  if (live) {
    const uint4* src = mult + (uint64_t)e * FOLD_G1_STEPS * 12;
    for (int idx = lane; idx < count; idx += T) {
      const uint32_t ent = sel[1 + idx];
      Xyzz t;
      load_xyzz(t, src + 12 * (ent & 255u));
      if (ent >> 8) t.x = fq_mul_call(t.x, beta);   // phi: X < 2q, inside the lazy invariant
      xyzz_add_fast_ni(&acc, &t);
    }
  }
#pragma unroll
  for (int off = T / 2; off >= 1; off >>= 1) {
    Xyzz o;
    xyzz_shfl_down<T>(o, acc, off);
    if (lane < off) xyzz_add_fast_ni(&acc, &o);
  }
  if (lane != 0 || !live) return;
  Affine l, out;
  load_affine(l, a + 6 * (uint64_t)e);
  xyzz_madd_fast_ni(&acc, &l);
  xyzz_canon(acc);
  xyzz_to_affine_ni(&out, &acc);
  store_affine(a + 6 * (uint64_t)e, out);
}

// phase A, G2: mult[i * 64 + j] = 2^j * h[first + i] as (canonical) XYZZ over Fq2
__global__ void __launch_bounds__(64) k_fold_pre_g2(const uint4* __restrict__ h, uint32_t first, uint32_t count,
                                                    uint4* __restrict__ mult) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Affine2 r;
  load_affine2(r, h + 12 * ((uint64_t)first + i));
  Xyzz2 t;
  xyzz2_set_inf(t);
  xyzz2_madd_ni(&t, &r);
  uint4* dst = mult + (uint64_t)i * FOLD_G2_STEPS * 24;
  for (int j = 0; j < FOLD_G2_STEPS; j++) {
    store_xyzz2(dst + 24 * j, t);
    if (j + 1 < FOLD_G2_STEPS) xyzz2_dbl_ni(&t);
  }
}

template <int T>
__device__ __forceinline__ void xyzz2_shfl_down(Xyzz2& o, const Xyzz2& p, int off) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
  uint32_t* d = reinterpret_cast<uint32_t*>(&o);
#pragma unroll 8
  for (int i = 0; i < 96; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], off, T);
}
static __device__ __noinline__ void xyzz2_psi_ni(Xyzz2* p) {
  Xyzz2 r;
  xyzz2_psi(r, *p);
  *p = r;
}

// phase B, G2: h[i] <- h[i] + (k0 + k1 x + k2 x^2 + k3 x^3) * h[split + i]; T lanes per element, list as for G1 with
// d = the power of psi. Launched with ceil(split T / 64) CTAs of 64 threads.
template <int T>
__global__ void __launch_bounds__(64) k_fold_apply_g2(uint4* __restrict__ h, uint32_t split,
                                                      const uint16_t* __restrict__ sel,
                                                      const uint4* __restrict__ mult) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t e = tid / T;
  const int lane = tid % T;
  const bool live = e < split;
  const int count = sel[0];
  Xyzz2 acc;
  xyzz2_set_inf(acc);
  if (live) {
    const uint4* src = mult + (uint64_t)e * FOLD_G2_STEPS * 24;
    for (int idx = lane; idx < count; idx += T) {
      const uint32_t ent = sel[1 + idx];
      Xyzz2 t;
      load_xyzz2(t, src + 24 * (ent & 255u));
      for (uint32_t d = 0; d < (ent >> 8); d++) xyzz2_psi_ni(&t);
      xyzz2_add_ni(&acc, &t);
    }
  }
#pragma unroll 1
  for (int off = T / 2; off >= 1; off >>= 1) {
    Xyzz2 o;
    xyzz2_shfl_down<T>(o, acc, off);
    if (lane < off) xyzz2_add_ni(&acc, &o);
  }
  if (lane != 0 || !live) return;
  Affine2 l, out;
  load_affine2(l, h + 12 * (uint64_t)e);
  xyzz2_madd_ni(&acc, &l);
  xyzz2_to_affine_ni(&out, &acc);
  store_affine2(h + 12 * (uint64_t)e, out);
}

// test hook: one Fq12 operation per thread (tests/test_gpu_pairing.py drives every op against the oracle)
//   0 mul(a,b)  1 sqr(a)  2 inv(a)  3 frobenius(a,1)  4 frobenius(a,2)  5 cyclotomic_sqr(a)  6 exp_by_x(a)
//   7 final_exp(a)  8 mul_by_034(a; b = l0 || l3 || l4)  9 miller(a = G1 affine || G2 affine)
//   20 w12_mul  21 w12_sqr  22 w12_cyclotomic_sqr  23 w12_final_exp  24 w12_frobenius(1)  25 w12_frobenius(2)
//   (warp-cooperative versions: element i is processed by the whole warp of block i / launched with n blocks)
__global__ void __launch_bounds__(W12_THREADS) k_test_w12_op(int op, const uint4* a, const uint4* b, uint4* out) {
  __shared__ WFinalExp s;
  const Team tm = team_cta();
  const int lane = threadIdx.x;
  uint4* f4 = reinterpret_cast<uint4*>(&s.f);
  uint4* g4 = reinterpret_cast<uint4*>(&s.f2);
  for (int i = lane; i < 36; i += blockDim.x) {
    f4[i] = a[36 * (size_t)blockIdx.x + i];
    g4[i] = b[36 * (size_t)blockIdx.x + i];
  }
  team_sync(tm);
  switch (op) {
    case 20: w12_mul(&s.r, &s.f, &s.f2, &s.w, tm); break;
    case 21: w12_sqr(&s.r, &s.f, &s.w, tm); break;
    case 22: w12_cyclotomic_sqr(&s.r, &s.f, &s.w, tm); break;
    case 23: w12_final_exp(&s, tm); break;
    case 24: w12_frobenius(&s.r, &s.f, 1, tm); break;
    case 25: w12_frobenius(&s.r, &s.f, 2, tm); break;
    case 26: w12_mul(&s.f, &s.f, &s.f, &s.w, tm); w12_copy(&s.r, &s.f, tm); break;     // aliasing
    case 28: {                                                                  // cooperative Miller loop: a = G1 || G2
      __shared__ WMiller ms;
      uint4* p4 = reinterpret_cast<uint4*>(&ms.p);
      uint4* q4 = reinterpret_cast<uint4*>(&ms.q);
      if (lane < 6) p4[lane] = a[36 * (size_t)blockIdx.x + lane];
      if (lane >= 8 && lane < 20) q4[lane - 8] = a[36 * (size_t)blockIdx.x + 6 + (lane - 8)];
      team_sync(tm);
      w_miller_loop(&ms, tm);
      w12_copy(&s.r, &ms.f, tm);
      break;
    }
    case 27: w12_cyclotomic_sqr(&s.f, &s.f, &s.w, tm); w12_conj(&s.r, &s.f, tm); break;
    case 29: {                                                                  // a chain: ((a b)^2 a)^2 conj, lazily reduced throughout
      w12_mul(&s.r, &s.f, &s.f2, &s.w, tm);
      w12_sqr(&s.r, &s.r, &s.w, tm);
      w12_mul(&s.r, &s.r, &s.f, &s.w, tm);
      w12_sqr(&s.r, &s.r, &s.w, tm);
      w12_conj(&s.r, &s.r, tm);
      break;
    }
    default: break;
  }
  w12_canon(&s.r, tm);
  const uint4* r4 = reinterpret_cast<const uint4*>(&s.r);
  for (int i = lane; i < 36; i += blockDim.x) out[36 * (size_t)blockIdx.x + i] = r4[i];
}

// a^e for GT elements on ONE TEAM per element (the verifier's `tx.pow(c)`, src/mipp.rs:258-261): the thread-per-element
// kernel above walks ~380 Fq12 operations of 36-54 serial Fq products each (~17 ms whatever n is); the verifier raises 2 m
// <= 28 values, so the chain's latency is all that counts: square-and-multiply from the top set bit with the cooperative
// Fq12 square / product (54 / 36 lanes busy per phase). Generic squarings: the inputs are proof values, nothing says they
// lie in the cyclotomic subgroup. Canonical input, canonical output.
struct WPow {
  Fq12 base, acc;
  WScratch w;
  uint32_t e[8];
};
__global__ void __launch_bounds__(W12_THREADS) k_fq12_pow_coop(const uint4* __restrict__ in,
                                                               const uint32_t* __restrict__ exps, int exps_mont,
                                                               uint4* __restrict__ out) {
  __shared__ WPow s;
  const Team tm = team_cta();
  const int lane = threadIdx.x;
  uint4* b4 = reinterpret_cast<uint4*>(&s.base);
  for (int i = lane; i < 36; i += blockDim.x) b4[i] = in[36 * (size_t)blockIdx.x + i];
  if (lane == 0) {
    uint32_t e[8];
    bool zero = true;
    for (int i = 0; i < 8; i++) {
      e[i] = exps[8 * (size_t)blockIdx.x + i];
      zero = zero && e[i] == 0;
    }
    if (exps_mont && !zero) mont_to_canonical<FrParams>(e, e);
    for (int i = 0; i < 8; i++) s.e[i] = e[i];
  }
  team_sync(tm);
  int top = -1;                                     // team-uniform: every lane reads the same shared words
  for (int i = 7; i >= 0 && top < 0; i--)
    if (s.e[i]) top = 32 * i + 31 - __clz(s.e[i]);
  if (top < 0) {                                    // a^0 = 1
    if (lane < 12) w12_q(&s.acc)[lane] = lane == 0 ? fq_one() : fq_zero();
    team_sync(tm);
  } else {
    w12_copy(&s.acc, &s.base, tm);
    for (int bit = top - 1; bit >= 0; bit--) {
      w12_sqr(&s.acc, &s.acc, &s.w, tm);
      if ((s.e[bit >> 5] >> (bit & 31)) & 1) w12_mul(&s.acc, &s.acc, &s.base, &s.w, tm);
    }
  }
  w12_canon(&s.acc, tm);
  const uint4* r4 = reinterpret_cast<const uint4*>(&s.acc);
  for (int i = lane; i < 36; i += blockDim.x) out[36 * (size_t)blockIdx.x + i] = r4[i];
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
