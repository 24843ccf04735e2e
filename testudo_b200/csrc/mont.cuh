// Montgomery arithmetic on 32-bit limbs for the two BLS12-377 prime fields (Fq: 12 limbs, Fr: 8 limbs).
//
// Replaces ark-ff 0.4 `Fp384<MontBackend<FqConfig,6>>` / `Fp256<MontBackend<FrConfig,4>>` (un-vendored
// dependency of the reference, Cargo.toml:21-22,77) underneath every G1 MSM the reference issues
// (SURVEY.md 2.3). Same value representation as ark in memory: little-endian limbs, Montgomery form,
// R = 2^(32*N) -- a u64[6] ark limb array is bit-identical to our u32[12].
//
// Design (B200 integer pipe): a 32x32->64 multiply-accumulate is one IMAD.WIDE on the FMA pipe; carries
// ride on the ALU pipe (IADD3.X) and predicate registers. The product is accumulated in TWO interleaved
// accumulators so that every wide product lands on a 64-bit-aligned register pair and a whole row of
// products is one carry chain:
//     T = E + O * 2^32 + x          (E: limbs e[0..N], O: limbs o[0..N-1], x: one orphan limb)
// Row i adds a_i * B: products a_i*b_j with j even go to E, j odd go to O. Because both moduli satisfy
//     p = 1 (mod 2^32)   and   -p^{-1} = 0xffffffff (mod 2^32)
// the Montgomery quotient digit is m = -(T mod 2^32) (a negate, no multiply) and the product m*p_0 = m
// needs no multiplier either: it only produces the carry k into limb 1. After the row, T is divisible
// by 2^32; dividing swaps the roles of E and O (E' = O, O' = E >> 64) and leaves the old e[1] as the new
// orphan x'. Cost per N-limb multiplication: N*(2N-1) IMAD.WIDE (276 for Fq) -- SURVEY.md 8d counts
// 2*N^2 = 288 wide MACs (576 IMAD) per modmul for the roofline.
//
// Every function is __host__ __device__: on the device the carry chains are PTX (add.cc / madc.lo.cc /
// madc.hi.cc), on the host they are emulated with an explicit carry so tests/test_field_host.py can check
// the algorithm's structure on a CPU-only box. The product never *runs* the host path (see capi.cu).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
#endif

namespace tb {

// ---------------------------------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------------------------------
struct Carry {
#ifndef __CUDA_ARCH__
  uint32_t f = 0;
#endif
};

TB_HD uint32_t add_cc(uint32_t a, uint32_t b, Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  uint64_t t = (uint64_t)a + b;
  c.f = (uint32_t)(t >> 32);
  return (uint32_t)t;
#endif
}
TB_HD uint32_t addc_cc(uint32_t a, uint32_t b, Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  uint64_t t = (uint64_t)a + b + c.f;
  c.f = (uint32_t)(t >> 32);
  return (uint32_t)t;
#endif
}
TB_HD uint32_t addc(uint32_t a, uint32_t b, Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  return a + b + c.f;
#endif
}
TB_HD uint32_t sub_cc(uint32_t a, uint32_t b, Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  uint64_t t = (uint64_t)a - b;
  c.f = (uint32_t)(t >> 63);  // borrow
  return (uint32_t)t;
#endif
}
TB_HD uint32_t subc_cc(uint32_t a, uint32_t b, Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  uint64_t t = (uint64_t)a - b - c.f;
  c.f = (uint32_t)(t >> 63);
  return (uint32_t)t;
#endif
}
// returns 0 - borrow (0 or 0xffffffff): the mask form of the final borrow
TB_HD uint32_t subc_mask(Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(r));
  return r;
#else
  return 0u - c.f;
#endif
}
// m = -s through PTX: written as C (`0u - s`) the negate gets folded into the multiplies that consume m and
// ptxas then no longer fuses their lo/hi halves into one IMAD.WIDE.U32.X (120 IMAD.X + 120 IMAD.HI.U32.X instead)
TB_HD uint32_t neg32(uint32_t s) {
#ifdef __CUDA_ARCH__
  uint32_t m;
  asm volatile("sub.u32 %0, 0, %1;" : "=r"(m) : "r"(s));
  return m;
#else
  return 0u - s;
#endif
}
TB_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
TB_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// lo/hi of a 32x32 product through ONE PTX mul.wide: ptxas emits IMAD.WIDE.U32 and cannot peephole the high half
// into a later addition (which splits the product into IMAD + IMAD.HI.U32 plus two moves)
TB_HD void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm volatile("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#else
  uint64_t t = (uint64_t)a * b;
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
#endif
}
TB_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t d, Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(d));
  return r;
#else
  uint64_t t = (uint64_t)(uint32_t)(a * b) + d;
  c.f = (uint32_t)(t >> 32);
  return (uint32_t)t;
#endif
}
TB_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t d, Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(d));
  return r;
#else
  uint64_t t = (uint64_t)(uint32_t)(a * b) + d + c.f;
  c.f = (uint32_t)(t >> 32);
  return (uint32_t)t;
#endif
}
TB_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t d, Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(d));
  return r;
#else
  uint64_t t = (((uint64_t)a * b) >> 32) + d + c.f;
  c.f = (uint32_t)(t >> 32);
  return (uint32_t)t;
#endif
}
TB_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t d, Carry& c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(d));
  return r;
#else
  return (uint32_t)(((uint64_t)a * b) >> 32) + d + c.f;
#endif
}

// ---------------------------------------------------------------------------------------------------
// field parameter packs (checked against oracle/bls12_377.py by tests/test_constants.py)
// ---------------------------------------------------------------------------------------------------
// Modulus limbs on the device live in constant memory (uniform-register / c[bank] operands of IMAD.WIDE).
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t FQ_P_DEV[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u,
                                                        0xba094800u, 0x1ef3622fu, 0x00f5138fu, 0x1a22d9f3u,
                                                        0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
static __device__ __constant__ uint32_t FR_P_DEV[8] = {0x00000001u, 0x0a118000u, 0xd0000001u, 0x59aa76feu,
                                                       0x5c37b001u, 0x60b44d1eu, 0x9a2ca556u, 0x12ab655eu};
#endif

struct FqParams {
  static constexpr int N = 12;
  // q, little-endian 32-bit limbs
  TB_HD static uint32_t p(int i) {
#ifdef __CUDA_ARCH__
    return FQ_P_DEV[i];
#else
    constexpr uint32_t P[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u, 0xba094800u, 0x1ef3622fu,
                                0x00f5138fu, 0x1a22d9f3u, 0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
    return P[i];
#endif
  }
  // R mod q (Montgomery one)
  TB_HD static uint32_t one(int i) {
    constexpr uint32_t V[12] = {0xffffff68u, 0x02cdffffu, 0x7fffffb1u, 0x51409f83u, 0x8a7d3ff2u, 0x9f7db3a9u,
                                0x6e7c6305u, 0x7b4e97b7u, 0x803c84e8u, 0x4cf495bfu, 0xe2fdf49au, 0x008d6661u};
    return V[i];
  }
  // R^2 mod q
  TB_HD static uint32_t r2(int i) {
    constexpr uint32_t V[12] = {0x9400cd22u, 0xb786686cu, 0xb00431b1u, 0x0329fcaau, 0x62d6b46du, 0x22a5f111u,
                                0x827dc3acu, 0xbfdf7d03u, 0x41790bf9u, 0x837e92f0u, 0x1e914b88u, 0x006dfccbu};
    return V[i];
  }
};
struct FrParams {
  static constexpr int N = 8;
  TB_HD static uint32_t p(int i) {
#ifdef __CUDA_ARCH__
    return FR_P_DEV[i];
#else
    constexpr uint32_t P[8] = {0x00000001u, 0x0a118000u, 0xd0000001u, 0x59aa76feu,
                               0x5c37b001u, 0x60b44d1eu, 0x9a2ca556u, 0x12ab655eu};
    return P[i];
#endif
  }
  TB_HD static uint32_t one(int i) {
    constexpr uint32_t V[8] = {0xfffffff3u, 0x7d1c7fffu, 0x6ffffff2u, 0x7257f50fu,
                               0x512c0feeu, 0x16d81575u, 0x2bbb9a9du, 0x0d4bda32u};
    return V[i];
  }
  TB_HD static uint32_t r2(int i) {
    constexpr uint32_t V[8] = {0xb861857bu, 0x25d577bau, 0x8860591fu, 0xcc2c27b5u,
                               0xe5dc8593u, 0xa7cc008fu, 0xeff1c939u, 0x011fdae7u};
    return V[i];
  }
};

// ---------------------------------------------------------------------------------------------------
// generic N-limb modular add / sub / Montgomery mul; all values canonical (< p) in and out
// ---------------------------------------------------------------------------------------------------
template <class P>
TB_HD void mod_reduce_once(uint32_t* r) {  // r in [0, 2p) -> [0, p)
  constexpr int N = P::N;
  uint32_t t[N];
  Carry c;
  t[0] = sub_cc(r[0], P::p(0), c);
#pragma unroll
  for (int i = 1; i < N; i++) t[i] = subc_cc(r[i], P::p(i), c);
  uint32_t borrow = subc_mask(c);  // all-ones if r < p
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}

template <class P>
TB_HD void mod_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  Carry c;
  r[0] = add_cc(a[0], b[0], c);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = addc_cc(a[i], b[i], c);
  // both fields leave >= 3 spare top bits, so a + b never carries out of N limbs
  mod_reduce_once<P>(r);
}

template <class P>
TB_HD void mod_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  Carry c;
  uint32_t t[N];
  t[0] = sub_cc(a[0], b[0], c);
#pragma unroll
  for (int i = 1; i < N; i++) t[i] = subc_cc(a[i], b[i], c);
  uint32_t borrow = subc_mask(c);
  Carry d;
  r[0] = add_cc(t[0], P::p(0) & borrow, d);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = addc_cc(t[i], P::p(i) & borrow, d);
}

template <class P>
TB_HD void mod_neg(uint32_t* r, const uint32_t* a) {  // r = -a mod p (0 stays 0)
  constexpr int N = P::N;
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < N; i++) nz |= a[i];
  uint32_t mask = nz ? 0xffffffffu : 0u;
  Carry c;
  r[0] = sub_cc(P::p(0) & mask, a[0], c);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = subc_cc(P::p(i) & mask, a[i], c);
}

template <class P>
TB_HD bool mod_is_zero(const uint32_t* a) {
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) nz |= a[i];
  return nz == 0;
}
template <class P>
TB_HD bool mod_eq(const uint32_t* a, const uint32_t* b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) d |= a[i] ^ b[i];
  return d == 0;
}

// r = a * b * R^-1 (mod p), NOT fully reduced: r < p + a*b/R. Even/odd accumulator CIOS described in the
// file header. Valid for any a < 2^(32N) and b < 2^(32N) - p (the running total stays below b + p + epsilon).
template <class P>
TB_HD void mont_mul_lazy(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  static_assert(N % 2 == 0, "even limb count");
  uint32_t e[N + 1], o[N], x = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t ai = a[i];
    Carry c;
    if (i == 0) {
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        e[j] = mul_lo(ai, b[j]);
        e[j + 1] = mul_hi(ai, b[j]);
        o[j] = mul_lo(ai, b[j + 1]);
        o[j + 1] = mul_hi(ai, b[j + 1]);
      }
      e[N] = 0;
    } else {
      // E += a_i * (b_0, b_2, ...)
      e[0] = mad_lo_cc(ai, b[0], e[0], c);
      e[1] = madc_hi_cc(ai, b[0], e[1], c);
#pragma unroll
      for (int j = 2; j < N; j += 2) {
        e[j] = madc_lo_cc(ai, b[j], e[j], c);
        e[j + 1] = madc_hi_cc(ai, b[j], e[j + 1], c);
      }
      e[N] = addc(e[N], 0, c);
      // O += a_i * (b_1, b_3, ...)
      o[0] = mad_lo_cc(ai, b[1], o[0], c);
      o[1] = madc_hi_cc(ai, b[1], o[1], c);
#pragma unroll
      for (int j = 2; j < N - 2; j += 2) {
        o[j] = madc_lo_cc(ai, b[j + 1], o[j], c);
        o[j + 1] = madc_hi_cc(ai, b[j + 1], o[j + 1], c);
      }
      o[N - 2] = madc_lo_cc(ai, b[N - 1], o[N - 2], c);
      o[N - 1] = madc_hi(ai, b[N - 1], o[N - 1], c);  // O < 2^(32N-5): no carry out
    }
    // Montgomery digit: m = -(T mod 2^32); m * p_0 = m cancels limb 0 and carries k into limb 1
    // k = carry(e0 + x) + (s != 0), both via the carry flag (s + 0xffffffff carries iff s != 0): four ALU-pipe
    // adds. Written as `k += (s != 0)` ptxas emits select logic built from IMAD.MOVs, which occupy the FMA pipe.
    uint32_t s = add_cc(e[0], x, c);
    uint32_t k = addc(0, 0, c);
    (void)add_cc(s, 0xffffffffu, c);
    k = addc(k, 0, c);
    uint32_t m = neg32(s);
    // E += m * (p_2, p_4, ...) from limb 1 up (p_0 handled above)
    e[1] = add_cc(e[1], k, c);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      e[j] = madc_lo_cc(m, P::p(j), e[j], c);
      e[j + 1] = madc_hi_cc(m, P::p(j), e[j + 1], c);
    }
    e[N] = addc(e[N], 0, c);
    // O += m * (p_1, p_3, ...)
    o[0] = mad_lo_cc(m, P::p(1), o[0], c);
    o[1] = madc_hi_cc(m, P::p(1), o[1], c);
#pragma unroll
    for (int j = 2; j < N - 2; j += 2) {
      o[j] = madc_lo_cc(m, P::p(j + 1), o[j], c);
      o[j + 1] = madc_hi_cc(m, P::p(j + 1), o[j + 1], c);
    }
    o[N - 2] = madc_lo_cc(m, P::p(N - 1), o[N - 2], c);
    o[N - 1] = madc_hi(m, P::p(N - 1), o[N - 1], c);
    // T /= 2^32:  x' = e[1];  E' = O;  O' = E >> 64
    x = e[1];
    uint32_t t[N];
#pragma unroll
    for (int j = 0; j < N; j++) t[j] = o[j];
#pragma unroll
    for (int j = 0; j < N - 1; j++) o[j] = e[j + 2];
    o[N - 1] = 0;
#pragma unroll
    for (int j = 0; j < N; j++) e[j] = t[j];
    e[N] = 0;
  }
  // r = x + E + O * 2^32  (< 2p), then one conditional subtraction
  Carry c;
  r[0] = add_cc(e[0], x, c);
#pragma unroll
  for (int j = 1; j < N; j++) r[j] = addc_cc(e[j], o[j - 1], c);
}

// r = (a * b + c * d) * R^-1 (mod p) with ONE Montgomery reduction (2 N^2 + N (N - 1) wide MACs instead of
// 2 N (2N - 1)): every row adds a_i * B and c_i * D to the even/odd accumulators before the quotient digit is taken.
// NOT fully reduced: r < p + (a b + c d) / R; requires b + d + p < 2^(32N) - the running total stays below
// b + d + p + epsilon, so the top-limb arguments of mont_mul_lazy carry over when b + d < 2^(32N - 4).
template <class P>
TB_HD void mont_mul2_lazy(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* cc, const uint32_t* d) {
  constexpr int N = P::N;
  uint32_t e[N + 1], o[N], x = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    Carry c;
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const uint32_t ai = half ? cc[i] : a[i];
      const uint32_t* bb = half ? d : b;
      if (i == 0 && half == 0) {
#pragma unroll
        for (int j = 0; j < N; j += 2) {
          mul_wide(e[j], e[j + 1], ai, bb[j]);
          mul_wide(o[j], o[j + 1], ai, bb[j + 1]);
        }
        e[N] = 0;
      } else {
        e[0] = mad_lo_cc(ai, bb[0], e[0], c);
        e[1] = madc_hi_cc(ai, bb[0], e[1], c);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
          e[j] = madc_lo_cc(ai, bb[j], e[j], c);
          e[j + 1] = madc_hi_cc(ai, bb[j], e[j + 1], c);
        }
        e[N] = addc(e[N], 0, c);
        o[0] = mad_lo_cc(ai, bb[1], o[0], c);
        o[1] = madc_hi_cc(ai, bb[1], o[1], c);
#pragma unroll
        for (int j = 2; j < N - 2; j += 2) {
          o[j] = madc_lo_cc(ai, bb[j + 1], o[j], c);
          o[j + 1] = madc_hi_cc(ai, bb[j + 1], o[j + 1], c);
        }
        o[N - 2] = madc_lo_cc(ai, bb[N - 1], o[N - 2], c);
        o[N - 1] = madc_hi(ai, bb[N - 1], o[N - 1], c);
      }
    }
    uint32_t s = add_cc(e[0], x, c);
    uint32_t k = addc(0, 0, c);
    (void)add_cc(s, 0xffffffffu, c);
    k = addc(k, 0, c);
    uint32_t m = neg32(s);
    e[1] = add_cc(e[1], k, c);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      e[j] = madc_lo_cc(m, P::p(j), e[j], c);
      e[j + 1] = madc_hi_cc(m, P::p(j), e[j + 1], c);
    }
    e[N] = addc(e[N], 0, c);
    o[0] = mad_lo_cc(m, P::p(1), o[0], c);
    o[1] = madc_hi_cc(m, P::p(1), o[1], c);
#pragma unroll
    for (int j = 2; j < N - 2; j += 2) {
      o[j] = madc_lo_cc(m, P::p(j + 1), o[j], c);
      o[j + 1] = madc_hi_cc(m, P::p(j + 1), o[j + 1], c);
    }
    o[N - 2] = madc_lo_cc(m, P::p(N - 1), o[N - 2], c);
    o[N - 1] = madc_hi(m, P::p(N - 1), o[N - 1], c);
    x = e[1];
    uint32_t t[N];
#pragma unroll
    for (int j = 0; j < N; j++) t[j] = o[j];
#pragma unroll
    for (int j = 0; j < N - 1; j++) o[j] = e[j + 2];
    o[N - 1] = 0;
#pragma unroll
    for (int j = 0; j < N; j++) e[j] = t[j];
    e[N] = 0;
  }
  Carry c;
  r[0] = add_cc(e[0], x, c);
#pragma unroll
  for (int j = 1; j < N; j++) r[j] = addc_cc(e[j], o[j - 1], c);
}

// r = a * a * R^-1 (mod p), lazily reduced like mont_mul_lazy, with 78 instead of 144 product MACs (N = 12):
// row i takes a_i * a_i (weight 2^(64 i)) and a_i * 2 a_j for j > i only. The doubled operand is prepared once:
// d[j] = limb j of 2a, dc[j] = (a_j << 1) without the bit shifted in from a_{j-1} (the lowest doubled limb of a row
// must not contain a_i's top bit). Relative to the row's base the products land on the same limb positions as in
// mont_mul_lazy, so the even/odd accumulators, the Montgomery step and all bounds are unchanged (the row sums only
// regroup the same total). Requires a < 2^(32N-1) (true for every lazily reduced Fq value: < 16 q < 2^381).
template <class P>
TB_HD void mont_sqr_lazy(uint32_t* r, const uint32_t* a) {
  constexpr int N = P::N;
  uint32_t d[N], dc[N];
#pragma unroll
  for (int j = 0; j < N; j++) {
    dc[j] = a[j] << 1;
    d[j] = j ? (dc[j] | (a[j - 1] >> 31)) : dc[j];
  }
  uint32_t e[N + 1], o[N], x = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t ai = a[i];
    Carry c;
    // v(j): the row's operand at relative limb j (j >= i)
#define TB_SQR_V(j) ((j) == i ? ai : ((j) == i + 1 ? dc[(j)] : d[(j)]))
    if (i == 0) {
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        e[j] = mul_lo(ai, TB_SQR_V(j));
        e[j + 1] = mul_hi(ai, TB_SQR_V(j));
        o[j] = mul_lo(ai, TB_SQR_V(j + 1));
        o[j + 1] = mul_hi(ai, TB_SQR_V(j + 1));
      }
      e[N] = 0;
    } else {
      // E: even relative limbs j >= i
      constexpr int dummy = 0;
      (void)dummy;
      const int je = (i & 1) ? i + 1 : i;  // first even j >= i
      const int jo = (i & 1) ? i : i + 1;  // first odd  j >= i
      if (je < N) {
        e[je] = mad_lo_cc(ai, TB_SQR_V(je), e[je], c);
        e[je + 1] = madc_hi_cc(ai, TB_SQR_V(je), e[je + 1], c);
#pragma unroll
        for (int j = 0; j < N; j += 2) {
          if (j > je) {
            e[j] = madc_lo_cc(ai, TB_SQR_V(j), e[j], c);
            e[j + 1] = madc_hi_cc(ai, TB_SQR_V(j), e[j + 1], c);
          }
        }
        e[N] = addc(e[N], 0, c);
      }
      if (jo < N) {
        o[jo - 1] = mad_lo_cc(ai, TB_SQR_V(jo), o[jo - 1], c);
        if (jo == N - 1) {
          o[jo] = madc_hi(ai, TB_SQR_V(jo), o[jo], c);
        } else {
          o[jo] = madc_hi_cc(ai, TB_SQR_V(jo), o[jo], c);
#pragma unroll
          for (int j = 1; j < N; j += 2) {
            if (j > jo && j < N - 1) {
              o[j - 1] = madc_lo_cc(ai, TB_SQR_V(j), o[j - 1], c);
              o[j] = madc_hi_cc(ai, TB_SQR_V(j), o[j], c);
            }
          }
          o[N - 2] = madc_lo_cc(ai, TB_SQR_V(N - 1), o[N - 2], c);
          o[N - 1] = madc_hi(ai, TB_SQR_V(N - 1), o[N - 1], c);
        }
      }
    }
#undef TB_SQR_V
    uint32_t s = add_cc(e[0], x, c);
    uint32_t k = addc(0, 0, c);
    (void)add_cc(s, 0xffffffffu, c);
    k = addc(k, 0, c);
    uint32_t m = neg32(s);
    e[1] = add_cc(e[1], k, c);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      e[j] = madc_lo_cc(m, P::p(j), e[j], c);
      e[j + 1] = madc_hi_cc(m, P::p(j), e[j + 1], c);
    }
    e[N] = addc(e[N], 0, c);
    o[0] = mad_lo_cc(m, P::p(1), o[0], c);
    o[1] = madc_hi_cc(m, P::p(1), o[1], c);
#pragma unroll
    for (int j = 2; j < N - 2; j += 2) {
      o[j] = madc_lo_cc(m, P::p(j + 1), o[j], c);
      o[j + 1] = madc_hi_cc(m, P::p(j + 1), o[j + 1], c);
    }
    o[N - 2] = madc_lo_cc(m, P::p(N - 1), o[N - 2], c);
    o[N - 1] = madc_hi(m, P::p(N - 1), o[N - 1], c);
    x = e[1];
    uint32_t t[N];
#pragma unroll
    for (int j = 0; j < N; j++) t[j] = o[j];
#pragma unroll
    for (int j = 0; j < N - 1; j++) o[j] = e[j + 2];
    o[N - 1] = 0;
#pragma unroll
    for (int j = 0; j < N; j++) e[j] = t[j];
    e[N] = 0;
  }
  Carry c;
  r[0] = add_cc(e[0], x, c);
#pragma unroll
  for (int j = 1; j < N; j++) r[j] = addc_cc(e[j], o[j - 1], c);
}

// canonical product: inputs < p, output < p
template <class P>
TB_HD void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  mont_mul_lazy<P>(r, a, b);
  mod_reduce_once<P>(r);
}

template <class P>
TB_HD void mont_sqr(uint32_t* r, const uint32_t* a) {
  mont_mul<P>(r, a, a);
}

// Montgomery -> canonical (multiply by 1): used for Fr scalars handed over in ark's in-memory form
template <class P>
TB_HD void mont_to_canonical(uint32_t* r, const uint32_t* a) {
  uint32_t one[P::N];
  one[0] = 1;
#pragma unroll
  for (int i = 1; i < P::N; i++) one[i] = 0;
  mont_mul<P>(r, a, one);
}

// thin typed wrappers -----------------------------------------------------------------------------------
struct Fq {
  uint32_t l[12];
};
struct Fr {
  uint32_t l[8];
};
TB_HD void fq_mul(Fq& r, const Fq& a, const Fq& b) { mont_mul<FqParams>(r.l, a.l, b.l); }
TB_HD void fq_sqr(Fq& r, const Fq& a) { mont_sqr<FqParams>(r.l, a.l); }
TB_HD void fq_add(Fq& r, const Fq& a, const Fq& b) { mod_add<FqParams>(r.l, a.l, b.l); }
TB_HD void fq_sub(Fq& r, const Fq& a, const Fq& b) { mod_sub<FqParams>(r.l, a.l, b.l); }
TB_HD void fq_dbl(Fq& r, const Fq& a) { mod_add<FqParams>(r.l, a.l, a.l); }
TB_HD void fq_neg(Fq& r, const Fq& a) { mod_neg<FqParams>(r.l, a.l); }
TB_HD bool fq_is_zero(const Fq& a) { return mod_is_zero<FqParams>(a.l); }
TB_HD bool fq_eq(const Fq& a, const Fq& b) { return mod_eq<FqParams>(a.l, b.l); }
TB_HD Fq fq_zero() {
  Fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = 0;
  return r;
}
TB_HD Fq fq_one() {
  Fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = FqParams::one(i);
  return r;
}

// a^(q-2) by square-and-multiply over the fixed exponent: 377 squarings + ~190 products, a ~400 000-instruction
// dependent chain. Kept as the cross-check of fq_inv (tests/test_host_logic.py, tb200_test_fq_inv).
TB_HD void fq_inv_fermat(Fq& r, const Fq& a) {
  // q - 2: limb 0 of q is 1 -> 0xffffffff with a borrow into limb 1
  Fq acc = fq_one();
  bool started = false;
  for (int i = 11; i >= 0; i--) {
    uint32_t w = FqParams::p(i);
    if (i == 0) w = 0xffffffffu;
    if (i == 1) w -= 1;
    for (int bit = 31; bit >= 0; bit--) {
      if (started) fq_sqr(acc, acc);
      if ((w >> bit) & 1) {
        if (started) fq_mul(acc, acc, a);
        else {
          acc = a;
          started = true;
        }
      }
    }
  }
  r = acc;
}

// r = a / 2 (mod q): (a + (a odd ? q : 0)) >> 1
TB_HD void fq_halve(Fq& r, const Fq& a) {
  const uint32_t mask = 0u - (a.l[0] & 1u);
  uint32_t t[12];
  Carry c;
  t[0] = add_cc(a.l[0], FqParams::p(0) & mask, c);
#pragma unroll
  for (int i = 1; i < 12; i++) t[i] = addc_cc(a.l[i], FqParams::p(i) & mask, c);
#pragma unroll
  for (int i = 0; i < 11; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
  r.l[11] = t[11] >> 1;
}
TB_HD void fq_shr1(Fq& a) {
#pragma unroll
  for (int i = 0; i < 11; i++) a.l[i] = (a.l[i] >> 1) | (a.l[i + 1] << 31);
  a.l[11] >>= 1;
}
TB_HD bool fq_is_one_int(const Fq& a) {  // the INTEGER 1 (not the Montgomery one)
  uint32_t nz = a.l[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < 12; i++) nz |= a.l[i];
  return nz == 0;
}
// plain 12-limb a >= b
TB_HD bool fq_geq(const Fq& a, const Fq& b) {
  Carry c;
  (void)sub_cc(a.l[0], b.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) (void)subc_cc(a.l[i], b.l[i], c);
  return subc_mask(c) == 0;
}
TB_HD void fq_sub_plain(Fq& r, const Fq& a, const Fq& b) {  // a - b for a >= b, no modular wrap
  Carry c;
  r.l[0] = sub_cc(a.l[0], b.l[0], c);
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = subc_cc(a.l[i], b.l[i], c);
}

// Inversion by the binary extended Euclidean algorithm (~2 x 377 iterations of 12-limb shifts / subtractions, ~10x fewer
// instructions than the Fermat ladder above): every latency-bound tail of the engine ends in one of these (to-affine of
// an MSM result, of every folded MIPP element, the final exponentiation's Fq12 inverse). Input and output in Montgomery
// form: for X = a R the loop yields X^-1 = a^-1 R^-1 as an integer; one Montgomery product with R^3 restores a^-1 R.
// 0 -> 0 (as the ladder).
TB_HD void fq_inv(Fq& r, const Fq& a) {
  if (fq_is_zero(a)) {
    r = fq_zero();
    return;
  }
  Fq u = a, v, x1 = fq_zero(), x2 = fq_zero();
#pragma unroll
  for (int i = 0; i < 12; i++) v.l[i] = FqParams::p(i);
  x1.l[0] = 1;
  while (!fq_is_one_int(u) && !fq_is_one_int(v)) {
    while ((u.l[0] & 1u) == 0) {
      fq_shr1(u);
      fq_halve(x1, x1);
    }
    while ((v.l[0] & 1u) == 0) {
      fq_shr1(v);
      fq_halve(x2, x2);
    }
    if (fq_geq(u, v)) {
      fq_sub_plain(u, u, v);
      fq_sub(x1, x1, x2);
    } else {
      fq_sub_plain(v, v, u);
      fq_sub(x2, x2, x1);
    }
  }
  const Fq res = fq_is_one_int(u) ? x1 : x2;
  Fq r2, r3;
#pragma unroll
  for (int i = 0; i < 12; i++) r2.l[i] = FqParams::r2(i);
  fq_mul(r3, r2, r2);       // R^2 R^2 R^-1 = R^3
  fq_mul(r, res, r3);       // a^-1 R^-1 R^3 R^-1 = a^-1 R
}

}  // namespace tb
