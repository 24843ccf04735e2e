// Signed radix-2^c recoding of a 253-bit canonical Fr scalar.
//
// Restates ark-ec 0.4 `make_digits`: coef = carry + window; carry = (coef + 2^(c-1)) >> c;
// digit = coef - (carry << c), digits in [-2^(c-1), 2^(c-1)), so only 2^(c-1) buckets per window are needed
// (negation of an affine point is free). Difference from ark: we size the window count as ceil(254 / c) so the
// top window always has at most c-1 payload bits and can absorb the final carry without a special case
// (ark instead adds the carry back into the last digit). The digit string represents the same integer.
#pragma once
#include "mont.cuh"

namespace tb {

constexpr int SCALAR_BITS = 253;  // Fr::MODULUS_BIT_SIZE

TB_HD int num_windows(int c) { return (SCALAR_BITS + 1 + c - 1) / c; }

struct DigitIter {
  const uint32_t* s;  // 8 limbs, canonical (< r)
  int c;
  int bit;
  uint32_t carry;
  TB_HD DigitIter(const uint32_t* s_, int c_) : s(s_), c(c_), bit(0), carry(0) {}
  // next signed digit; `last` = this is the top window (no outgoing carry)
  TB_HD int32_t next(bool last) {
    int limb = bit >> 5, off = bit & 31;
    uint32_t lo = limb < 8 ? s[limb] : 0u;
    uint32_t hi = limb + 1 < 8 ? s[limb + 1] : 0u;
    uint64_t v = (((uint64_t)hi << 32) | lo) >> off;
    uint32_t coef = ((uint32_t)v & ((1u << c) - 1u)) + carry;
    bit += c;
    if (last) return (int32_t)coef;
    carry = (coef + (1u << (c - 1))) >> c;
    return (int32_t)coef - (int32_t)(carry << c);
  }
};

}  // namespace tb
