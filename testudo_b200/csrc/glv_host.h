// Host-side scalar decompositions for the MIPP folds (kernels_pairing.cuh, "two-phase folds").
//
// The fold scalar of a round is ONE value for the whole vector (src/mipp.rs:106-114: compress(&mut m_a, split, &c)), so
// the set of stored multiples 2^j * P that it selects is the same for every element. The host decomposes the scalar over
// the curve endomorphisms and hands the device a SELECTION LIST of (bit position, endomorphism power) pairs; the lanes
// that share an element split the list evenly, every lane does the same number of additions, nothing is predicated off.
//   G1: k = k0 + k1 lambda,  lambda = x^2 - 1 (phi(x, y) = (beta x, y) = [lambda](x, y)),  k0, k1 < 2^127
//   G2: k = k0 + k1 x + k2 x^2 + k3 x^3,  x = the curve parameter (psi = [x] on G2),  k_i < x < 2^64
// An entry is  j | (d << 8)  with j the bit position and d the endomorphism power; list[0] holds the count.
#pragma once
#include <cstdint>
#include <cstring>

namespace tbe {
namespace glv {

typedef unsigned __int128 u128;
constexpr uint64_t X = 0x8508c00000000001ull;  // ark-bls12-377 Config::X
constexpr uint64_t R_LIMBS[4] = {0x0a11800000000001ull, 0x59aa76fed0000001ull, 0x60b44d1e5c37b001ull, 0x12ab655e9a2ca556ull};
constexpr uint64_t RINV_LIMBS[4] = {0x4122dd1a1beeec02ull, 0xbd1eae9574fee875ull, 0x838557e227b28e2full,
                                    0x07b301912290c02cull};  // (2^256)^-1 mod r
constexpr int SEL_MAX = 264;  // uint16 entries per list including the count (G1: <= 254, G2: <= 256)

inline bool geq(const uint64_t a[4], const uint64_t b[4]) {
  for (int i = 3; i >= 0; i--)
    if (a[i] != b[i]) return a[i] > b[i];
  return true;
}
inline void sub(uint64_t a[4], const uint64_t b[4]) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    const u128 d = (u128)a[i] - b[i] - borrow;
    a[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
}
// ark's in-memory Fr (Montgomery, R = 2^256) -> canonical integer: a * R^-1 mod r
inline void from_mont(const uint64_t a[4], uint64_t out[4]) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 4; j++) {
      const u128 p = (u128)a[i] * RINV_LIMBS[j] + t[i + j] + carry;
      t[i + j] = (uint64_t)p;
      carry = (uint64_t)(p >> 64);
    }
    t[i + 4] = carry;
  }
  uint64_t rem[4] = {0, 0, 0, 0};  // binary long division of the 512-bit product by r (r < 2^253: no overflow)
  for (int bit = 511; bit >= 0; bit--) {
    for (int i = 3; i > 0; i--) rem[i] = (rem[i] << 1) | (rem[i - 1] >> 63);
    rem[0] = (rem[0] << 1) | ((t[bit >> 6] >> (bit & 63)) & 1);
    if (geq(rem, R_LIMBS)) sub(rem, R_LIMBS);
  }
  memcpy(out, rem, 32);
}
inline void canonical(const uint64_t k[4], bool mont, uint64_t out[4]) {
  if (mont) {
    from_mont(k, out);
  } else {
    memcpy(out, k, 32);
    while (geq(out, R_LIMBS)) sub(out, R_LIMBS);  // values >= r are reduced, as `BigInt -> Fr` would
  }
}

// G1: list of (j, d) with bit j of k_d set, d in {0, 1}
inline void select_g1(const uint64_t k_in[4], bool mont, uint16_t list[SEL_MAX]) {
  uint64_t k[4];
  canonical(k_in, mont, k);
  const u128 lam = (u128)X * X - 1;
  u128 rem = 0, quo = 0;  // k = quo * lam + rem; quo < 2^127 for k < r
  for (int bit = 255; bit >= 0; bit--) {
    const bool top = (rem >> 127) != 0;
    rem = (rem << 1) | ((k[bit >> 6] >> (bit & 63)) & 1);
    quo <<= 1;
    if (top || rem >= lam) {
      rem -= lam;
      quo |= 1;
    }
  }
  int n = 0;
  for (int j = 0; j < 127; j++) {
    if ((uint64_t)(rem >> j) & 1) list[1 + n++] = (uint16_t)j;
    if ((uint64_t)(quo >> j) & 1) list[1 + n++] = (uint16_t)(j | (1 << 8));
  }
  list[0] = (uint16_t)n;
}

// G2: list of (j, d) with bit j of the d-th base-x digit set, d in 0..3
inline void select_g2(const uint64_t k_in[4], bool mont, uint16_t list[SEL_MAX]) {
  uint64_t q[4];
  canonical(k_in, mont, q);
  uint64_t digit[4];
  for (int d = 0; d < 4; d++) {
    u128 rem = 0;
    for (int i = 3; i >= 0; i--) {
      const u128 cur = (rem << 64) | q[i];
      q[i] = (uint64_t)(cur / X);
      rem = cur % X;
    }
    digit[d] = (uint64_t)rem;
  }
  int n = 0;
  for (int j = 0; j < 64; j++)
    for (int d = 0; d < 4; d++)
      if ((digit[d] >> j) & 1) list[1 + n++] = (uint16_t)(j | (d << 8));
  list[0] = (uint16_t)n;
}

}  // namespace glv
}  // namespace tbe
