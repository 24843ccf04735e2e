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

// out[b] = final_exponentiation(in[b]); one thread per product
__global__ void __launch_bounds__(32) k_final_exp(const uint4* __restrict__ in, uint4* __restrict__ out) {
  if (threadIdx.x != 0) return;
  Fq12 f, e;
  load_fq12(f, in + 36 * (size_t)blockIdx.x);
  fq12_final_exp(e, f);
  store_fq12(out + 36 * (size_t)blockIdx.x, e);
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

// test hook: one Fq12 operation per thread (tests/test_gpu_pairing.py drives every op against the oracle)
//   0 mul(a,b)  1 sqr(a)  2 inv(a)  3 frobenius(a,1)  4 frobenius(a,2)  5 cyclotomic_sqr(a)  6 exp_by_x(a)
//   7 final_exp(a)  8 mul_by_034(a; b = l0 || l3 || l4)  9 miller(a = G1 affine || G2 affine)
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
