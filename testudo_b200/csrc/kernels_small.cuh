// Small-n G1 MSM (n <= 1024) as a Straus multiplication, 8 points per one-warp CTA: no sort pipeline, no Horner chain.
//
// Why: the bucket pipeline needs ~40 launches and ends in a serial window combine, so a 2-point MSM costs as much as a
// 2^12-point one (2.5-2.9 ms; profiles/r01_summary.md section 4) -- and the reference issues many tiny MSMs:
// `commit_scalar` (src/commitments.rs:70-77), the bullet rounds (src/nizk/bullet.rs:93-118), the last MIPP rounds
// (src/mipp.rs:77-85). Below ~2^6 points the work is nothing and the LATENCY of the doubling chain is everything.
//
// Shape: Straus / fixed windows. Every point gets a QUAD of lanes. Per point: a table {1..8} B, then 64 signed radix-16
// digits (digits.cuh, the same recoding as the pipeline) -> 4 doublings + 1 table addition per digit. The doubling chain
// cannot be shortened (no endomorphism here: like ark's MSM this path accepts ANY curve point, not only the order-r
// subgroup), so the quad shortens every link instead: the 9 products of an XYZZ doubling run as 3 rounds of <= 4
// independent products on the 4 lanes, the 14 of a full addition as 4 rounds -- one uniform out-of-line multiplier call
// per round with the operands selected per lane (lanes of a warp that take different branches run one after the other,
// so only cheap limb additions sit in lane-specific code). Operands travel through shared-memory slots laid out
// [slot][16-byte group][quad] (conflict-free LDS.128 / STS.128). The per-point results are summed by a shared-memory
// tree with the exact (exceptional-case complete) XYZZ addition. One WARP per CTA (8 points): two warps on one SM already
// share a scheduler's multiplier pipe and stretch the chain (measured: 64 points in one CTA 2.08 ms, in eight 1.6 ms);
// the CTAs leave their partial sums in global scratch and the last one to finish (atomic ticket) adds them and normalises.
// Arithmetic: the lazily reduced group law of the hot loop (g1_fast.cuh), same bounds (X < 8q, Y < 4q, ZZ, ZZZ < 2q).
#pragma once
#include "kernels.cuh"

namespace tb {

constexpr int SMALL_MAX_POINTS = 1024;
constexpr int SMALL_QUADS = 8;                // points per CTA = quads of one warp
constexpr int SMALL_TABLE = 8;                // multiples 1..8 of the point
constexpr int SMALL_DIGITS = 64;              // num_windows(4)
// working slots of a quad
enum { QS_X = 0, QS_Y, QS_ZZ, QS_ZZZ, QS_EX, QS_EY, QS_EZZ, QS_EZZZ, QS_A, QS_B, QS_C, QS_D, QS_F, QS_G, QS_H, QS_SLOTS };
constexpr int SMALL_SLOTS_PER_QUAD = QS_SLOTS + 4 * SMALL_TABLE;   // 15 working + 32 table = 47 Fq values (2256 B)

struct QuadCtx {
  uint4* sm;       // slot storage
  int nq;          // quads in the CTA
  int quad, role;  // this lane
};
__device__ __forceinline__ Fq q_ld(const QuadCtx& c, int slot) {
  const uint4* p = c.sm + (size_t)slot * 3 * c.nq + c.quad;
  const uint4 a = p[0], b = p[c.nq], d = p[2 * c.nq];
  Fq v;
  v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
  v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
  v.l[8] = d.x; v.l[9] = d.y; v.l[10] = d.z; v.l[11] = d.w;
  return v;
}
__device__ __forceinline__ void q_st(const QuadCtx& c, int slot, const Fq& v) {
  uint4* p = c.sm + (size_t)slot * 3 * c.nq + c.quad;
  p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  p[c.nq] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
  p[2 * c.nq] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}
// one round of products: lane-selected operand slots (-1 = this lane has no product in this round); ONE multiplier call
__device__ __forceinline__ Fq q_round(const QuadCtx& c, int sa, int sb) {
  Fq r = fq_zero();
  if (sa >= 0) {
    const Fq a = q_ld(c, sa), b = q_ld(c, sb);
    r = fq_mul_call(a, b);
  }
  return r;
}
__device__ __forceinline__ bool fq_all_zero(const Fq& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) o |= a.l[i];
  return o == 0;
}
__device__ __forceinline__ void fq_dbl_plain(Fq& r, const Fq& a) {   // 2a as integers (no reduction, a < 2^383)
  Carry cy;
  r.l[0] = add_cc(a.l[0], a.l[0], cy);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = addc_cc(a.l[i], a.l[i], cy);
}
__device__ __forceinline__ void fq_add_plain(Fq& r, const Fq& a, const Fq& b) {
  Carry cy;
  r.l[0] = add_cc(a.l[0], b.l[0], cy);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = addc_cc(a.l[i], b.l[i], cy);
}

// acc = 2 acc on the quad (dbl-2008-s-1, as xyzz_dbl_fast): rounds {U^2, X^2}, {U V, X V, V ZZ, M^2}, {W Y, M T, W ZZZ}.
// The identity (ZZ == 0) stays the identity: ZZ' = V * 0.
static __device__ __noinline__ void quad_dbl(QuadCtx c) {
  const int r = c.role;
  if (r == 0) {                                  // U = 2 Y < 8q
    Fq u;
    fq_dbl_plain(u, q_ld(c, QS_Y));
    q_st(c, QS_A, u);
  }
  __syncwarp();
  {
    Fq p = q_round(c, r == 0 ? QS_A : (r == 1 ? QS_X : -1), r == 0 ? QS_A : QS_X);
    if (r == 0) q_st(c, QS_B, p);                // V = U^2
    if (r == 1) {                                // M = 3 X^2 < 4.5q
      Fq t, m;
      fq_dbl_plain(t, p);
      fq_add_plain(m, t, p);
      q_st(c, QS_C, m);
    }
  }
  __syncwarp();
  {
    const int sa = r == 0 ? QS_A : (r == 1 ? QS_X : (r == 2 ? QS_B : QS_C));
    const int sb = r == 0 ? QS_B : (r == 1 ? QS_B : (r == 2 ? QS_ZZ : QS_C));
    Fq p = q_round(c, sa, sb);
    if (r == 0) q_st(c, QS_D, p);                // W = U V
    if (r == 1) q_st(c, QS_F, p);                // S = X V
    if (r == 2) q_st(c, QS_ZZ, p);               // ZZ' = V ZZ  (ZZ has no other reader in this doubling)
    if (r == 3) q_st(c, QS_G, p);                // M^2
  }
  __syncwarp();
  if (r == 1) {                                  // X' = M^2 + 4q - 2S < 5.2q ; T = S + 8q - X' < 9.1q
    const Fq s = q_ld(c, QS_F);
    Fq x = q_ld(c, QS_G), t;
    fq_sub_lazy<0>(x, x, s);
    fq_sub_lazy<0>(x, x, s);
    q_st(c, QS_X, x);
    fq_sub_lazy<2>(t, s, x);
    q_st(c, QS_F, t);
  }
  __syncwarp();
  {
    const int sa = r == 0 ? QS_D : (r == 1 ? QS_C : (r == 2 ? QS_D : -1));
    const int sb = r == 0 ? QS_Y : (r == 1 ? QS_F : QS_ZZZ);
    Fq p = q_round(c, sa, sb);
    if (r == 0) q_st(c, QS_G, p);                // W Y
    if (r == 1) q_st(c, QS_H, p);                // M T
    if (r == 2) q_st(c, QS_ZZZ, p);              // ZZZ' = W ZZZ
  }
  __syncwarp();
  if (r == 1) {                                  // Y' = M T + 2q - W Y < 3.4q
    Fq y;
    fq_sub_lazy<0>(y, q_ld(c, QS_H), q_ld(c, QS_G));
    q_st(c, QS_Y, y);
  }
  __syncwarp();
}

// acc += E (both XYZZ in slots; add-2008-s as xyzz_add_fast). All cases exact: acc or E the identity, acc == E
// (doubling) and acc == -E (identity) are decided on the canonical values and, being rare, run on one lane.
static __device__ __noinline__ void quad_add(QuadCtx c, int* quad_flag) {
  const int r = c.role;
  if (r == 0) {
    const bool a_inf = fq_all_zero(q_ld(c, QS_ZZ)), e_inf = fq_all_zero(q_ld(c, QS_EZZ));
    quad_flag[c.quad] = e_inf ? 2 : (a_inf ? 1 : 0);
  }
  __syncwarp();
  int mode = quad_flag[c.quad];
  if (mode == 1) q_st(c, QS_X + r, q_ld(c, QS_EX + r));      // acc = E: every lane copies one coordinate
  __syncwarp();
  const bool act = mode == 0;
  {
    const int sa = r == 0 ? QS_X : (r == 1 ? QS_EX : (r == 2 ? QS_Y : QS_EY));
    const int sb = r == 0 ? QS_EZZ : (r == 1 ? QS_ZZ : (r == 2 ? QS_EZZZ : QS_ZZZ));
    Fq p = q_round(c, act ? sa : -1, sb);
    if (act) q_st(c, QS_A + r, p);               // U1, U2, S1, S2 -> A, B, C, D
  }
  __syncwarp();
  if (act && r == 1) {                           // P = U2 + 2q - U1 in (0, 3.2q); P = 0 (mod q) <=> same x
    Fq p;
    fq_sub_lazy<0>(p, q_ld(c, QS_B), q_ld(c, QS_A));
    q_st(c, QS_B, p);
    bool zero = false;
    if (p.l[0] - 1u < 3u || fq_all_zero(p)) {
      Fq chk = p;
      fq_canon(chk);
      zero = fq_is_zero(chk);
    }
    if (zero) quad_flag[c.quad] = 3;
  }
  if (act && r == 3) {                           // R = S2 + 2q - S1
    Fq t;
    fq_sub_lazy<0>(t, q_ld(c, QS_D), q_ld(c, QS_C));
    q_st(c, QS_D, t);
  }
  __syncwarp();
  mode = quad_flag[c.quad];
  if (mode == 3) {                               // same x: doubling or identity, exactly, on lane 0
    if (r == 0) {
      Xyzz a, e;
      a.x = q_ld(c, QS_X); a.y = q_ld(c, QS_Y); a.zz = q_ld(c, QS_ZZ); a.zzz = q_ld(c, QS_ZZZ);
      e.x = q_ld(c, QS_EX); e.y = q_ld(c, QS_EY); e.zz = q_ld(c, QS_EZZ); e.zzz = q_ld(c, QS_EZZZ);
      a = xyzz_add_exact(a, e);
      q_st(c, QS_X, a.x); q_st(c, QS_Y, a.y); q_st(c, QS_ZZ, a.zz); q_st(c, QS_ZZZ, a.zzz);
    }
  }
  const bool go = mode == 0;
  {
    const int sa = r == 0 ? QS_B : (r == 1 ? QS_D : (r == 2 ? QS_ZZ : QS_ZZZ));
    const int sb = r == 0 ? QS_B : (r == 1 ? QS_D : (r == 2 ? QS_EZZ : QS_EZZZ));
    Fq p = q_round(c, go ? sa : -1, sb);
    if (go) {
      if (r == 0) q_st(c, QS_F, p);              // PP
      if (r == 1) q_st(c, QS_G, p);              // RR
      if (r == 2) q_st(c, QS_ZZ, p);             // ZZ1 ZZ2
      if (r == 3) q_st(c, QS_ZZZ, p);            // ZZZ1 ZZZ2
    }
  }
  __syncwarp();
  {
    const int sa = r == 0 ? QS_B : (r == 1 ? QS_A : (r == 2 ? QS_ZZ : -1));
    Fq p = q_round(c, go ? sa : -1, QS_F);
    if (go) {
      if (r == 0) q_st(c, QS_H, p);              // PPP = P PP
      if (r == 1) q_st(c, QS_A, p);              // Q = U1 PP
      if (r == 2) q_st(c, QS_ZZ, p);             // ZZ' = ZZ1 ZZ2 PP < 2q
    }
  }
  __syncwarp();
  if (go && r == 1) {                            // X' = RR + 6q - PPP - 2Q < 7.1q ; T = Q + 8q - X'
    const Fq qq = q_ld(c, QS_A);
    Fq x = q_ld(c, QS_G), t;
    fq_sub_lazy<0>(x, x, q_ld(c, QS_H));
    fq_sub_lazy<0>(x, x, qq);
    fq_sub_lazy<0>(x, x, qq);
    q_st(c, QS_X, x);
    fq_sub_lazy<2>(t, qq, x);
    q_st(c, QS_A, t);
  }
  __syncwarp();
  {
    const int sa = r == 0 ? QS_C : (r == 1 ? QS_D : (r == 2 ? QS_ZZZ : -1));
    const int sb = r == 0 ? QS_H : (r == 1 ? QS_A : QS_H);
    Fq p = q_round(c, go ? sa : -1, sb);
    if (go) {
      if (r == 0) q_st(c, QS_C, p);              // S1 PPP
      if (r == 1) q_st(c, QS_G, p);              // R T
      if (r == 2) q_st(c, QS_ZZZ, p);            // ZZZ' < 2q
    }
  }
  __syncwarp();
  if (go && r == 1) {                            // Y' = R T + 2q - S1 PPP < 3.3q
    Fq y;
    fq_sub_lazy<0>(y, q_ld(c, QS_G), q_ld(c, QS_C));
    q_st(c, QS_Y, y);
  }
  __syncwarp();
}

// table entry k (1..8) <-> the accumulator / operand slots; every lane moves one coordinate
__device__ __forceinline__ void quad_copy(const QuadCtx& c, int dst, int src) { q_st(c, dst + c.role, q_ld(c, src + c.role)); }
__device__ __forceinline__ int quad_table_slot(int k) { return QS_SLOTS + 4 * (k - 1); }

// out = sum_i scalars[i] * bases[i], n <= SMALL_MAX_POINTS. Grid = ceil(n / 8) CTAs of 32 threads. scratch: gridDim.x XYZZ
// partial sums (12 uint4 each) followed by one uint4 whose .x is the ticket counter (zeroed by the caller).
// `per_row` > 0 (<= 8): gridDim.x INDEPENDENT MSMs of per_row points each -- CTA b multiplies points
// [b * per_row, (b + 1) * per_row) and writes its own affine result to out_affine + 6 b; no cross-CTA sum, no scratch
// (the verifier's g_mask[i] - z_i g, ark-poly-commit `check`: nv two-point MSMs for the latency of one).
// `segs` != nullptr: a RAGGED batch of independent MSMs (rows of 0 .. 1024 points) in one launch. CTA b reads its
// descriptor segs[5 b ..] = {first point, points (<= 8), row, first CTA of the row, CTAs of the row}; the CTAs of a row
// leave their partial sums at scratch + 12 b, the last one to finish (one ticket per row: the uint32 array behind the
// gridDim.x partials) adds them and writes out_affine + 6 row.
constexpr int SMALL_SEG_WORDS = 5;
__global__ void __launch_bounds__(4 * SMALL_QUADS) k_msm_small(const uint4* __restrict__ bases,
                                                               const uint32_t* __restrict__ scalars, uint32_t n, int mont,
                                                               uint4* __restrict__ scratch, uint4* __restrict__ out_affine,
                                                               uint32_t per_row, const uint32_t* __restrict__ segs) {
  __shared__ uint4 small_sm[SMALL_SLOTS_PER_QUAD * 3 * SMALL_QUADS + (SMALL_QUADS * SMALL_DIGITS + SMALL_QUADS * 4 + 16) / 16];
  __shared__ int is_last;
  constexpr int nq = SMALL_QUADS;
  QuadCtx c;
  c.sm = small_sm;
  c.nq = nq;
  c.quad = threadIdx.x >> 2;
  c.role = threadIdx.x & 3;
  uint32_t first, per, out_row, row_cta0, nb;        // this CTA's points, its row, and the row's CTAs
  if (segs) {
    const uint32_t* sg = segs + (size_t)SMALL_SEG_WORDS * blockIdx.x;
    first = sg[0]; per = sg[1]; out_row = sg[2]; row_cta0 = sg[3]; nb = sg[4];
  } else if (per_row) {
    first = blockIdx.x * per_row; per = per_row; out_row = blockIdx.x; row_cta0 = blockIdx.x; nb = 1;
  } else {
    first = blockIdx.x * nq; per = nq; out_row = 0; row_cta0 = 0; nb = gridDim.x;
  }
  const uint32_t pt = first + c.quad;                // this quad's point
  int8_t* digits = reinterpret_cast<int8_t*>(small_sm + (size_t)SMALL_SLOTS_PER_QUAD * 3 * nq);   // [nq][64]
  int* quad_flag = reinterpret_cast<int*>(digits + (size_t)nq * SMALL_DIGITS);                    // [nq]
  if (threadIdx.x == 0) is_last = 0;
  const bool live = (uint32_t)c.quad < per && pt < n;
  // ---- per point: digits, table entry 1 = the point itself (or the identity), accumulator = identity --------------------
  bool any = false;
  if (c.role == 0) {
    Affine b;
    if (live) load_affine(b, bases + 6 * (size_t)pt);
    uint32_t k[8];
    bool kz = true;
    for (int i = 0; i < 8; i++) {
      k[i] = live ? scalars[8 * (size_t)pt + i] : 0u;
      kz = kz && k[i] == 0;
    }
    if (live && mont && !kz) mont_to_canonical<FrParams>(k, k);
    const bool dead = !live || kz || affine_is_inf(b);
    DigitIter it(k, 4);
    for (int w = 0; w < SMALL_DIGITS; w++) digits[c.quad * SMALL_DIGITS + w] = dead ? 0 : (int8_t)it.next(w == SMALL_DIGITS - 1);
    const int t1 = quad_table_slot(1);
    q_st(c, t1 + 0, dead ? fq_zero() : b.x);
    q_st(c, t1 + 1, dead ? fq_zero() : b.y);
    q_st(c, t1 + 2, dead ? fq_zero() : fq_one());
    q_st(c, t1 + 3, dead ? fq_zero() : fq_one());
    quad_flag[c.quad] = 0;
  }
  q_st(c, QS_X + c.role, fq_zero());             // accumulator = identity
  __syncwarp();
  // ---- table 2..8: 2B = dbl(B), 3B = 2B + B, 4B = dbl(2B), 5B = 4B + B, 6B = dbl(3B), 7B = 6B + B, 8B = dbl(4B) --------------
  quad_copy(c, QS_EX, quad_table_slot(1));       // E = B for the three additions
  __syncwarp();
#pragma unroll 1
  for (int k = 2; k <= SMALL_TABLE; k++) {
    if (k & 1) {                                 // acc still holds (k-1) B from the previous step
      quad_add(c, quad_flag);
    } else {
      quad_copy(c, QS_X, quad_table_slot(k / 2));
      __syncwarp();
      quad_dbl(c);
    }
    quad_copy(c, quad_table_slot(k), QS_X);
    __syncwarp();
  }
  q_st(c, QS_X + c.role, fq_zero());
  __syncwarp();
  // ---- 64 windows, most significant first ----------------------------------------------------------------------------------
#pragma unroll 1
  for (int w = SMALL_DIGITS - 1; w >= 0; w--) {
    if (any) {                                   // warp-uniform: nothing to double before the first non-zero digit
      quad_dbl(c);
      quad_dbl(c);
      quad_dbl(c);
      quad_dbl(c);
    }
    const int d = digits[c.quad * SMALL_DIGITS + w];
    const int mag = d < 0 ? -d : d;
    if (mag) {                                   // operand = +-(|d| B): y negated as 4q - y (y < 4q)
      Fq v = q_ld(c, quad_table_slot(mag) + c.role);
      if (c.role == 1 && d < 0) {
        Fq z = fq_zero();
        fq_sub_lazy<1>(v, z, v);
      }
      q_st(c, QS_EX + c.role, v);
    } else {
      q_st(c, QS_EX + c.role, fq_zero());        // the identity: quad_add skips it
    }
    __syncwarp();
    if (__any_sync(0xffffffffu, mag != 0)) {
      quad_add(c, quad_flag);
      any = true;
    }
  }
  // ---- sum of the CTA's 8 results (exact addition), then the cross-CTA sum by the last CTA to arrive ----------------------
  __syncthreads();
  uint4* tree = small_sm + (size_t)QS_SLOTS * 3 * nq;      // the table area, free now: nq XYZZ values of 12 uint4
  Xyzz acc;
  xyzz_set_inf(acc);
  if (c.role == 0) {
    acc.x = q_ld(c, QS_X); acc.y = q_ld(c, QS_Y); acc.zz = q_ld(c, QS_ZZ); acc.zzz = q_ld(c, QS_ZZZ);
  }
  __syncthreads();
  if (c.role == 0) store_xyzz(tree + 12 * c.quad, acc);
  for (int stride = nq >> 1; stride >= 1; stride >>= 1) {
    __syncthreads();
    if (c.role == 0 && c.quad < stride) {
      Xyzz o;
      load_xyzz(o, tree + 12 * (c.quad + stride));
      xyzz_add_fast_ni(&acc, &o);
      store_xyzz(tree + 12 * c.quad, acc);
    }
  }
  if (nb > 1) {
    if (threadIdx.x == 0) {
      store_xyzz(scratch + 12 * (size_t)blockIdx.x, acc);
      __threadfence();
      const uint32_t ticket = atomicAdd(reinterpret_cast<uint32_t*>(scratch + 12 * (size_t)gridDim.x) + out_row, 1u);
      is_last = ticket == nb - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // lane l sums the partials l, l + 32, ...; shuffle-free tree through shared memory
    xyzz_set_inf(acc);
    for (uint32_t i = threadIdx.x; i < nb; i += 32) {
      Xyzz o;
      const uint4* src = scratch + 12 * ((size_t)row_cta0 + i);
      uint32_t* d = reinterpret_cast<uint32_t*>(&o);
      for (int w = 0; w < 12; w++) {
        const uint4 v = __ldcg(src + w);           // written by other CTAs: bypass L1
        d[4 * w] = v.x; d[4 * w + 1] = v.y; d[4 * w + 2] = v.z; d[4 * w + 3] = v.w;
      }
      xyzz_add_fast_ni(&acc, &o);
    }
    uint4* t32 = small_sm;                         // 32 XYZZ values = 384 uint4 (the slot area is free now)
    store_xyzz(t32 + 12 * threadIdx.x, acc);
    for (int stride = 16; stride >= 1; stride >>= 1) {
      __syncthreads();
      if ((int)threadIdx.x < stride) {
        Xyzz o;
        load_xyzz(o, t32 + 12 * (threadIdx.x + stride));
        xyzz_add_fast_ni(&acc, &o);
        store_xyzz(t32 + 12 * threadIdx.x, acc);
      }
    }
  }
  if (threadIdx.x == 0) {
    xyzz_canon(acc);
    Affine o;
    xyzz_to_affine_ni(&o, &acc);
    store_affine(out_affine + 6 * (size_t)out_row, o);
  }
}

}  // namespace tb
