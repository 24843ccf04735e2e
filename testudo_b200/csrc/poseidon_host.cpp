// Host-side Poseidon sponge for the Fiat-Shamir transcript of `Polynomial::open` / `MippProof::prove`
// (src/poseidon_transcript.rs:12-125). The transcript is CPU code in the reference as well (a few
// dozen permutations per MIPP round between GPU calls); it lives in the library so that a host in any language gets the
// same challenges without re-implementing the sponge, and so that the Python mirror does not spend ~3 ms per round in
// big-integer arithmetic. This is NOT a CPU path for the MSM engine: nothing here touches group arithmetic.
//
// Restates ark-crypto-primitives 0.4.0 `sponge::poseidon::PoseidonSponge` (dependency of the reference, Cargo.toml:28,
// un-vendored) from its published source:
//   * state = [capacity | rate] field elements, all zero; mode Absorbing{next_absorb_index = 0}
//   * permute: R_F/2 full rounds, R_P partial rounds, R_F/2 full rounds; a round = add round constants, S-box x^alpha on
//     every element (full) or on element 0 (partial), multiply by the MDS matrix (new[i] = sum_j state[j] mds[i][j])
//   * absorb(elements): when squeezing, permute first and restart at rate index 0; when the absorb index reached the
//     rate, permute first; then add elements into state[capacity + index ..], permuting whenever the rate is full
//   * squeeze_native_field_elements(n): when absorbing, permute first; when the squeeze index reached the rate, permute
//     first; then read state[capacity + index ..], permuting between full rates (not after the last read)
//   * absorb(&Vec<u8>) = absorb(pack(le64(len) || bytes)) with (MODULUS_BIT_SIZE - 1) / 8 bytes per element, little endian
//   * squeeze_field_elements::<F2>(1) for a foreign field F2 = squeeze_bits(F2::MODULUS_BIT_SIZE - 1): the low
//     MODULUS_BIT_SIZE - 1 bits (little endian) of ceil(bits / usable) native elements, as an integer mod F2's modulus
// PARITY UNPINNED against the arkworks binary except for the parameters, which reproduce the reference's
// constants (tests/test_poseidon_transcript.py); ffi/kat/ prints known answers from a Rust run for tests/test_ark_kat.py.
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include <cstdlib>

#include "../../include/testudo_b200.h"

namespace {

typedef unsigned __int128 u128;

// ---- BMI2 / ADX path of the Fq multiplier (6 limbs): each row `t += a * b` is ONE pass of mulx with the low halves on the
// CF chain (adcx) and the high halves on the OF chain (adox) -- two independent carry chains the compiler cannot express
// from C. Selected at run time (cpu_has_adx); results are identical to the portable path (tests/test_poseidon_transcript.py
// runs both against the Python sponge).
#if defined(__x86_64__) && defined(__GNUC__)
#define TB200_HAVE_ADX_PATH 1
static inline void adx_row7(uint64_t& t0, uint64_t& t1, uint64_t& t2, uint64_t& t3, uint64_t& t4, uint64_t& t5, uint64_t& t6, const uint64_t* a, uint64_t b) {
  uint64_t lo, hi, z;
  asm(
    "xorl %k[z], %k[z]\n\t"
    "mulx 0(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t0]\n\t adox %[hi], %[t1]\n\t"
    "mulx 8(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t1]\n\t adox %[hi], %[t2]\n\t"
    "mulx 16(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t2]\n\t adox %[hi], %[t3]\n\t"
    "mulx 24(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t3]\n\t adox %[hi], %[t4]\n\t"
    "mulx 32(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t4]\n\t adox %[hi], %[t5]\n\t"
    "mulx 40(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t5]\n\t adox %[hi], %[t6]\n\t"
    "adcx %[z], %[t6]\n\t"
      : [t0] "+r"(t0), [t1] "+r"(t1), [t2] "+r"(t2), [t3] "+r"(t3), [t4] "+r"(t4), [t5] "+r"(t5), [t6] "+r"(t6), [lo] "=&r"(lo), [hi] "=&r"(hi), [z] "=&r"(z)
      : [a] "r"(a), "d"(b), "m"(*(const uint64_t(*)[6])a)
      : "cc");
}
static inline void adx_row8(uint64_t& t0, uint64_t& t1, uint64_t& t2, uint64_t& t3, uint64_t& t4, uint64_t& t5, uint64_t& t6, uint64_t& t7, const uint64_t* a, uint64_t b) {
  uint64_t lo, hi, z;
  asm(
    "xorl %k[z], %k[z]\n\t"
    "mulx 0(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t0]\n\t adox %[hi], %[t1]\n\t"
    "mulx 8(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t1]\n\t adox %[hi], %[t2]\n\t"
    "mulx 16(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t2]\n\t adox %[hi], %[t3]\n\t"
    "mulx 24(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t3]\n\t adox %[hi], %[t4]\n\t"
    "mulx 32(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t4]\n\t adox %[hi], %[t5]\n\t"
    "mulx 40(%[a]), %[lo], %[hi]\n\t adcx %[lo], %[t5]\n\t adox %[hi], %[t6]\n\t"
    "adcx %[z], %[t6]\n\t"
    "adox %[z], %[t7]\n\t adcx %[z], %[t7]\n\t"
      : [t0] "+r"(t0), [t1] "+r"(t1), [t2] "+r"(t2), [t3] "+r"(t3), [t4] "+r"(t4), [t5] "+r"(t5), [t6] "+r"(t6), [t7] "+r"(t7), [lo] "=&r"(lo), [hi] "=&r"(hi), [z] "=&r"(z)
      : [a] "r"(a), "d"(b), "m"(*(const uint64_t(*)[6])a)
      : "cc");
}
static inline bool cpu_has_adx() {
  static const bool yes = __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("adx") && !getenv("TB200_NO_ADX");
  return yes;
}
// r = x * y * R^-1 mod p for the 6-limb field with a clear top bit (no-carry CIOS), r < p
static inline void adx_mul6(uint64_t* r, const uint64_t* x, const uint64_t* y, const uint64_t* p, uint64_t inv) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0;
#define TB200_ADX_STEP(a0, a1, a2, a3, a4, a5, a6, yi)  \
  adx_row7(a0, a1, a2, a3, a4, a5, a6, x, yi);          \
  adx_row7(a0, a1, a2, a3, a4, a5, a6, p, a0 * inv);    \
  a0 = 0; /* the low limb cancelled: it becomes the (empty) top limb of the next row */
  TB200_ADX_STEP(t0, t1, t2, t3, t4, t5, t6, y[0])
  TB200_ADX_STEP(t1, t2, t3, t4, t5, t6, t0, y[1])
  TB200_ADX_STEP(t2, t3, t4, t5, t6, t0, t1, y[2])
  TB200_ADX_STEP(t3, t4, t5, t6, t0, t1, t2, y[3])
  TB200_ADX_STEP(t4, t5, t6, t0, t1, t2, t3, y[4])
  TB200_ADX_STEP(t5, t6, t0, t1, t2, t3, t4, y[5])
#undef TB200_ADX_STEP
  // the value sits in (t6, t0, t1, t2, t3, t4), below 2p: one branch-free subtraction
  const uint64_t v[6] = {t6, t0, t1, t2, t3, t4};
  uint64_t s[6], borrow = 0;
  for (int j = 0; j < 6; j++) {
    const u128 d = (u128)v[j] - p[j] - borrow;
    s[j] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  for (int j = 0; j < 6; j++) r[j] = borrow ? v[j] : s[j];
}
// r = (a0 b0 + a1 b1 + a2 b2) * R^-1 mod p with ONE reduction per row (the MDS row of the width-3 sponge), r < p
static inline void adx_dot3_6(uint64_t* r, const uint64_t* a0, const uint64_t* b0, const uint64_t* a1, const uint64_t* b1,
                              const uint64_t* a2, const uint64_t* b2, const uint64_t* p, uint64_t inv) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0;
#define TB200_ADX_STEP(c0, c1, c2, c3, c4, c5, c6, c7, i)  \
  adx_row8(c0, c1, c2, c3, c4, c5, c6, c7, a0, b0[i]);     \
  adx_row8(c0, c1, c2, c3, c4, c5, c6, c7, a1, b1[i]);     \
  adx_row8(c0, c1, c2, c3, c4, c5, c6, c7, a2, b2[i]);     \
  adx_row8(c0, c1, c2, c3, c4, c5, c6, c7, p, c0 * inv);   \
  c0 = 0;
  TB200_ADX_STEP(t0, t1, t2, t3, t4, t5, t6, t7, 0)
  TB200_ADX_STEP(t1, t2, t3, t4, t5, t6, t7, t0, 1)
  TB200_ADX_STEP(t2, t3, t4, t5, t6, t7, t0, t1, 2)
  TB200_ADX_STEP(t3, t4, t5, t6, t7, t0, t1, t2, 3)
  TB200_ADX_STEP(t4, t5, t6, t7, t0, t1, t2, t3, 4)
  TB200_ADX_STEP(t5, t6, t7, t0, t1, t2, t3, t4, 5)
#undef TB200_ADX_STEP
  // the value sits in (t6, t7, t0, t1, t2, t3 | t4), below 4p
  uint64_t v[7] = {t6, t7, t0, t1, t2, t3, t4};
  for (;;) {
    uint64_t s[6], borrow = 0;
    for (int j = 0; j < 6; j++) {
      const u128 d = (u128)v[j] - p[j] - borrow;
      s[j] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
    if (v[6] == 0 && borrow) break;          // v < p
    v[6] -= borrow;
    for (int j = 0; j < 6; j++) v[j] = s[j];
  }
  for (int j = 0; j < 6; j++) r[j] = v[j];
}
#endif

// Montgomery arithmetic over an odd modulus of N 64-bit limbs (N = 4: Fr, N = 6: Fq)
template <int N>
struct Field {
  uint64_t p[N], r2[N], one[N], inv;  // inv = -p^-1 mod 2^64
  bool nocarry = false;               // top bit of p clear: the interleaved CIOS of mul() applies
  bool adx = false;                   // BMI2 + ADX available: the two-carry-chain rows (6-limb field)
  static bool geq(const uint64_t* a, const uint64_t* b) {
    for (int i = N - 1; i >= 0; i--)
      if (a[i] != b[i]) return a[i] > b[i];
    return true;
  }
  static uint64_t sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t borrow = 0;
    for (int i = 0; i < N; i++) {
      const u128 d = (u128)a[i] - b[i] - borrow;
      r[i] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
  }
  void add(uint64_t* r, const uint64_t* a, const uint64_t* b) const {
    uint64_t carry = 0;
    for (int i = 0; i < N; i++) {
      const u128 s = (u128)a[i] + b[i] + carry;
      r[i] = (uint64_t)s;
      carry = (uint64_t)(s >> 64);
    }
    if (carry || geq(r, p)) sub_n(r, r, p);
  }
  // lo(a b + c + d), hi in `hi`: never overflows 128 bits
  static inline uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t& hi) {
    const u128 s = (u128)a * b + c + d;
    hi = (uint64_t)(s >> 64);
    return (uint64_t)s;
  }
  // t (N limbs, < 2p) -> r = t mod p without a branch on the data
  void reduce_once(uint64_t* r, const uint64_t* t) const {
    uint64_t s[N], borrow = 0;
#pragma GCC unroll 8
    for (int j = 0; j < N; j++) {
      const u128 d = (u128)t[j] - p[j] - borrow;
      s[j] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
#pragma GCC unroll 8
    for (int j = 0; j < N; j++) r[j] = borrow ? t[j] : s[j];
  }
  // CIOS with the multiplication and the reduction row interleaved on two carry words (A, C). The top bit of the
  // modulus' top limb is clear for both fields (Fq: 377 of 384 bits, Fr: 253 of 256), so the running value stays inside
  // N limbs (t[N-1] = A + C cannot overflow) -- checked in init(), which falls back to mul_generic otherwise. Fully
  // unrolled: ~2x the speed of the generic loop (the sponge spends all its time here: ~630 products per permutation).
  void mul(uint64_t* r, const uint64_t* x, const uint64_t* y) const {
    if (!nocarry) return mul_generic(r, x, y);
#ifdef TB200_HAVE_ADX_PATH
    if (N == 6 && adx) return adx_mul6(r, x, y, p, inv);
#endif
    uint64_t t[N];
#pragma GCC unroll 8
    for (int j = 0; j < N; j++) t[j] = 0;
#pragma GCC unroll 8
    for (int i = 0; i < N; i++) {
      uint64_t A, C;
      const uint64_t yi = y[i];
      const uint64_t t0 = mac(x[0], yi, t[0], 0, A);
      const uint64_t m = t0 * inv;
      (void)mac(m, p[0], t0, 0, C);
#pragma GCC unroll 8
      for (int j = 1; j < N; j++) {
        const uint64_t tj = mac(x[j], yi, t[j], A, A);
        t[j - 1] = mac(m, p[j], tj, C, C);
      }
      t[N - 1] = C + A;
    }
    reduce_once(r, t);
  }
  void mul_generic(uint64_t* r, const uint64_t* a, const uint64_t* b) const {  // CIOS, any odd modulus
    uint64_t t[N + 2] = {0};
    for (int i = 0; i < N; i++) {
      uint64_t carry = 0;
      for (int j = 0; j < N; j++) {
        const u128 s = (u128)a[j] * b[i] + t[j] + carry;
        t[j] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
      }
      u128 s = (u128)t[N] + carry;
      t[N] = (uint64_t)s;
      t[N + 1] = (uint64_t)(s >> 64);
      const uint64_t m = t[0] * inv;
      s = (u128)m * p[0] + t[0];
      carry = (uint64_t)(s >> 64);
      for (int j = 1; j < N; j++) {
        s = (u128)m * p[j] + t[j] + carry;
        t[j - 1] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
      }
      s = (u128)t[N] + carry;
      t[N - 1] = (uint64_t)s;
      t[N] = t[N + 1] + (uint64_t)(s >> 64);
    }
    if (t[N] || geq(t, p)) sub_n(t, t, p);
    memcpy(r, t, N * 8);
  }
  // r = sum_k a[k] * b[k] (K products) with ONE Montgomery reduction: the rows of the K products are added before each
  // quotient digit is taken (the MDS row of a Poseidon round: three products, one reduction instead of three).
  // a[k], b[k] point at N-limb values < p; K <= 16 (the running total stays below (K + 1) p 2^64, inside N + 2 limbs).
  void dot(uint64_t* r, const uint64_t* const* a, const uint64_t* const* b, unsigned K) const {
    uint64_t t[N + 2] = {0};
    for (int i = 0; i < N; i++) {
      for (unsigned k = 0; k < K; k++) {
        const uint64_t bi = b[k][i];
        const uint64_t* ak = a[k];
        uint64_t carry = 0;
        for (int j = 0; j < N; j++) {
          const u128 s = (u128)ak[j] * bi + t[j] + carry;
          t[j] = (uint64_t)s;
          carry = (uint64_t)(s >> 64);
        }
        const u128 s = (u128)t[N] + carry;
        t[N] = (uint64_t)s;
        t[N + 1] += (uint64_t)(s >> 64);
      }
      const uint64_t m = t[0] * inv;
      u128 s = (u128)m * p[0] + t[0];
      uint64_t carry = (uint64_t)(s >> 64);
      for (int j = 1; j < N; j++) {
        s = (u128)m * p[j] + t[j] + carry;
        t[j - 1] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
      }
      s = (u128)t[N] + carry;
      t[N - 1] = (uint64_t)s;
      t[N] = t[N + 1] + (uint64_t)(s >> 64);
      t[N + 1] = 0;
    }
    while (t[N] || geq(t, p)) t[N] -= sub_n(t, t, p);   // the total is below (K + 1) p
    memcpy(r, t, N * 8);
  }
  // the same for exactly three terms (the MDS row of the reference's width-3 sponge), fully unrolled
  void dot3(uint64_t* r, const uint64_t* a0, const uint64_t* b0, const uint64_t* a1, const uint64_t* b1, const uint64_t* a2,
            const uint64_t* b2) const {
#ifdef TB200_HAVE_ADX_PATH
    if (N == 6 && adx && nocarry) return adx_dot3_6(r, a0, b0, a1, b1, a2, b2, p, inv);
#endif
    uint64_t t[N + 2];
#pragma GCC unroll 8
    for (int j = 0; j < N + 2; j++) t[j] = 0;
#pragma GCC unroll 8
    for (int i = 0; i < N; i++) {
      uint64_t c0 = 0, c1 = 0, c2 = 0;
      const uint64_t y0 = b0[i], y1 = b1[i], y2 = b2[i];
#pragma GCC unroll 8
      for (int j = 0; j < N; j++) {
        uint64_t v = mac(a0[j], y0, t[j], c0, c0);
        v = mac(a1[j], y1, v, c1, c1);
        t[j] = mac(a2[j], y2, v, c2, c2);
      }
      u128 top = (u128)t[N] + c0 + c1 + c2;
      t[N] = (uint64_t)top;
      t[N + 1] += (uint64_t)(top >> 64);
      const uint64_t m = t[0] * inv;
      uint64_t C;
      (void)mac(m, p[0], t[0], 0, C);
#pragma GCC unroll 8
      for (int j = 1; j < N; j++) t[j - 1] = mac(m, p[j], t[j], C, C);
      top = (u128)t[N] + C;
      t[N - 1] = (uint64_t)top;
      t[N] = t[N + 1] + (uint64_t)(top >> 64);
      t[N + 1] = 0;
    }
    while (t[N] || geq(t, p)) t[N] -= sub_n(t, t, p);   // the total is below 4 p
    memcpy(r, t, N * 8);
  }
  void to_mont(uint64_t* r, const uint64_t* a) const { mul(r, a, r2); }
  void from_mont(uint64_t* r, const uint64_t* a) const {
    uint64_t o[N] = {1};
    mul(r, a, o);
  }
  void init(const uint64_t* modulus) {
    memcpy(p, modulus, N * 8);
    nocarry = (p[N - 1] >> 63) == 0;
#ifdef TB200_HAVE_ADX_PATH
    adx = cpu_has_adx();
#endif
    uint64_t x = 1;  // Newton: x = p^-1 mod 2^64
    for (int i = 0; i < 6; i++) x *= 2 - p[0] * x;
    inv = (uint64_t)(0 - x);
    // R mod p and R^2 mod p by repeated doubling of 1
    uint64_t v[N] = {1};
    for (int i = 0; i < 2 * 64 * N; i++) {
      uint64_t carry = 0;
      for (int j = 0; j < N; j++) {
        const uint64_t nv = (v[j] << 1) | carry;
        carry = v[j] >> 63;
        v[j] = nv;
      }
      if (carry || geq(v, p)) sub_n(v, v, p);
      if (i == 64 * N - 1) memcpy(one, v, N * 8);
    }
    memcpy(r2, v, N * 8);
  }
};

const uint64_t FR_MOD[4] = {0x0a11800000000001ull, 0x59aa76fed0000001ull, 0x60b44d1e5c37b001ull, 0x12ab655e9a2ca556ull};
const uint64_t FQ_MOD[6] = {0x8508c00000000001ull, 0x170b5d4430000000ull, 0x1ef3622fba094800ull,
                            0x1a22d9f300f5138full, 0xc63b05c06ca1493bull, 0x01ae3a4617c510eaull};

struct SpongeBase {
  virtual ~SpongeBase() {}
  virtual void reset() = 0;
  virtual void absorb_bytes(const uint8_t* data, size_t len) = 0;
  virtual int absorb_native(const uint64_t* canonical, size_t n) = 0;
  virtual void squeeze_native(uint64_t* canonical, size_t n) = 0;
  virtual void squeeze_fr(uint64_t out[4]) = 0;
  virtual int limbs() const = 0;
};

template <int N>
struct Sponge : SpongeBase {
  Field<N> F;
  unsigned full, partial, rate, cap, width, bits;  // bits = MODULUS_BIT_SIZE
  uint64_t alpha;
  std::vector<uint64_t> ark, mds, state;  // Montgomery form
  bool absorbing = true;
  unsigned index = 0;  // next_absorb_index / next_squeeze_index

  int limbs() const override { return N; }
  uint64_t* st(unsigned i) { return state.data() + (size_t)N * i; }
  void reset() override {
    std::fill(state.begin(), state.end(), 0);
    absorbing = true;
    index = 0;
  }
  void sbox(uint64_t* x) const {   // x^alpha, left to right from the top bit (alpha = 17: four squarings and one product)
    uint64_t acc[N], base[N];
    memcpy(base, x, N * 8);
    memcpy(acc, x, N * 8);
    int top = 63;
    while (!((alpha >> top) & 1)) top--;
    for (int b = top - 1; b >= 0; b--) {
      F.mul(acc, acc, acc);
      if ((alpha >> b) & 1) F.mul(acc, acc, base);
    }
    memcpy(x, acc, N * 8);
  }
  void permute() {
    std::vector<uint64_t> nxt((size_t)N * width);
    const unsigned half = full / 2;
    for (unsigned r = 0; r < full + partial; r++) {
      for (unsigned i = 0; i < width; i++) F.add(st(i), st(i), ark.data() + (size_t)N * (r * width + i));
      const bool is_full = r < half || r >= half + partial;
      if (is_full)
        for (unsigned i = 0; i < width; i++) sbox(st(i));
      else
        sbox(st(0));
      const uint64_t* sp[16];
      const uint64_t* mp[16];
      for (unsigned j = 0; j < width; j++) sp[j] = st(j);
      for (unsigned i = 0; i < width; i++) {               // row i of the MDS matrix: one reduction for the whole row
        for (unsigned j = 0; j < width; j++) mp[j] = mds.data() + (size_t)N * (i * width + j);
        if (width == 3) F.dot3(nxt.data() + (size_t)N * i, sp[0], mp[0], sp[1], mp[1], sp[2], mp[2]);
        else F.dot(nxt.data() + (size_t)N * i, sp, mp, width);
      }
      state.swap(nxt);
    }
  }
  // `absorb` of a slice of native elements given in Montgomery form
  void absorb_mont(const uint64_t* elems, size_t n) {
    if (n == 0) return;
    unsigned start;
    if (absorbing) {
      start = index;
      if (start == rate) {
        permute();
        start = 0;
      }
    } else {
      permute();
      start = 0;
    }
    size_t done = 0;
    for (;;) {
      if (start + (n - done) <= rate) {
        for (size_t i = 0; done + i < n; i++) F.add(st(cap + start + (unsigned)i), st(cap + start + (unsigned)i), elems + N * (done + i));
        absorbing = true;
        index = start + (unsigned)(n - done);
        return;
      }
      const unsigned take = rate - start;
      for (unsigned i = 0; i < take; i++) F.add(st(cap + start + i), st(cap + start + i), elems + N * (done + i));
      permute();
      done += take;
      start = 0;
    }
  }
  int absorb_native(const uint64_t* canonical, size_t n) override {
    std::vector<uint64_t> m((size_t)N * n);
    for (size_t i = 0; i < n; i++) {
      if (Field<N>::geq(canonical + N * i, F.p)) return TB200_E_ARG;
      F.to_mont(m.data() + N * i, canonical + N * i);
    }
    absorb_mont(m.data(), n);
    return 0;
  }
  void absorb_bytes(const uint8_t* data, size_t len) override {
    // Absorb for Vec<u8>: le64(len) || bytes, packed (MODULUS_BIT_SIZE - 1) / 8 bytes per element, little endian
    std::vector<uint8_t> buf(8 + len);
    const uint64_t l = (uint64_t)len;
    for (int i = 0; i < 8; i++) buf[i] = (uint8_t)(l >> (8 * i));
    if (len) memcpy(buf.data() + 8, data, len);
    const size_t chunk = (bits - 1) / 8;
    const size_t n = (buf.size() + chunk - 1) / chunk;
    std::vector<uint64_t> m((size_t)N * n, 0);
    for (size_t i = 0; i < n; i++) {
      uint8_t tmp[N * 8] = {0};
      const size_t take = std::min(chunk, buf.size() - i * chunk);
      memcpy(tmp, buf.data() + i * chunk, take);
      uint64_t v[N];
      for (int k = 0; k < N; k++) {
        v[k] = 0;
        for (int b = 0; b < 8; b++) v[k] |= (uint64_t)tmp[8 * k + b] << (8 * b);
      }
      F.to_mont(m.data() + N * i, v);  // chunk < modulus: at most MODULUS_BIT_SIZE - 1 bits
    }
    absorb_mont(m.data(), n);
  }
  void squeeze_native(uint64_t* canonical, size_t n) override {
    if (n == 0) return;
    unsigned start;
    if (absorbing) {
      permute();
      start = 0;
    } else {
      start = index;
      if (start == rate) {
        permute();
        start = 0;
      }
    }
    size_t done = 0;
    for (;;) {
      if (start + (n - done) <= rate) {
        for (size_t i = 0; done + i < n; i++) F.from_mont(canonical + N * (done + i), st(cap + start + (unsigned)i));
        absorbing = false;
        index = start + (unsigned)(n - done);
        return;
      }
      const unsigned take = rate - start;
      for (unsigned i = 0; i < take; i++) F.from_mont(canonical + N * (done + i), st(cap + start + i));
      // ark: "Unless we are done with squeezing in this call, permute" -- its test compares the length of the output
      // slice BEFORE the elements just read are cut off (so a read that starts mid-rate with exactly `rate` elements left
      // continues without a permutation); restated as it is
      if (n - done != rate) permute();
      done += take;
      start = 0;
    }
  }
  // challenge_scalar::<Fr>: native when the sponge is over Fr, else the foreign-field path through squeeze_bits
  void squeeze_fr(uint64_t out[4]) override {
    if (N == 4) {
      squeeze_native(out, 1);
      return;
    }
    const unsigned want = 253 - 1;                       // FieldElementSize::Full for Fr
    const unsigned usable = bits - 1;
    const unsigned cnt = (want + usable - 1) / usable;   // 1 for Fq
    std::vector<uint64_t> e((size_t)N * cnt);
    squeeze_native(e.data(), cnt);
    uint64_t v[4] = {0, 0, 0, 0};                        // the low 252 bits of the first element: < 2^252 < r
    for (unsigned b = 0; b < want; b++)
      if ((e[b >> 6] >> (b & 63)) & 1) v[b >> 6] |= 1ull << (b & 63);
    memcpy(out, v, 32);
  }
};

template <int N>
SpongeBase* make(const uint64_t* modulus, unsigned bits, unsigned full, unsigned partial, uint64_t alpha, unsigned rate,
                 unsigned cap, const uint64_t* ark, const uint64_t* mds) {
  Sponge<N>* s = new Sponge<N>();
  s->F.init(modulus);
  s->full = full;
  s->partial = partial;
  s->alpha = alpha;
  s->rate = rate;
  s->cap = cap;
  s->width = rate + cap;
  s->bits = bits;
  const size_t na = (size_t)(full + partial) * s->width, nm = (size_t)s->width * s->width;
  s->ark.resize(N * na);
  s->mds.resize(N * nm);
  for (size_t i = 0; i < na; i++) {
    if (Field<N>::geq(ark + N * i, s->F.p)) {
      delete s;
      return nullptr;
    }
    s->F.to_mont(s->ark.data() + N * i, ark + N * i);
  }
  for (size_t i = 0; i < nm; i++) {
    if (Field<N>::geq(mds + N * i, s->F.p)) {
      delete s;
      return nullptr;
    }
    s->F.to_mont(s->mds.data() + N * i, mds + N * i);
  }
  s->state.assign((size_t)N * s->width, 0);
  return s;
}

// ---- `transcript.append(label, &value)` (src/poseidon_transcript.rs:21-27): the value's UNCOMPRESSED CanonicalSerialize
// bytes (ark-serialize 0.4) from its in-memory words, restated next to the Python encoder (testudo_b200/serialize.py):
//   Fp: canonical value, little endian (Fq 48 bytes, Fr 32); Fq2 = c0 || c1; Fq12 = twelve Fq in tower order;
//   SW affine point: x || y with SWFlags in the top bits of the LAST byte -- bit 7: y > -y (Fq2: c1 decides, then c0),
//   bit 6: the point at infinity (all coordinates written as zero).
struct WordCodec {
  Field<6> fq;
  Field<4> fr;
  uint64_t half[6];   // (q - 1) / 2: y > -y  <=>  y > half
  WordCodec() {
    fq.init(FQ_MOD);
    fr.init(FR_MOD);
    for (int i = 0; i < 6; i++) half[i] = (FQ_MOD[i] >> 1) | (i + 1 < 6 ? FQ_MOD[i + 1] << 63 : 0);
  }
  static void put(std::vector<uint8_t>& out, const uint64_t* v, int limbs) {
    for (int i = 0; i < limbs; i++)
      for (int b = 0; b < 8; b++) out.push_back((uint8_t)(v[i] >> (8 * b)));
  }
  bool above_half(const uint64_t* v) const { return !Field<6>::geq(half, v); }   // v > half
  static bool is_zero(const uint64_t* v, int limbs) {
    uint64_t acc = 0;
    for (int i = 0; i < limbs; i++) acc |= v[i];
    return acc == 0;
  }
  // nwords: 4 = Fr, 12 = G1 affine, 24 = G2 affine, 72 = Fq12; false for anything else
  bool encode(std::vector<uint8_t>& out, const uint64_t* w, size_t nwords) const {
    uint64_t c[6];
    if (nwords == 4) {
      uint64_t v[4];
      fr.from_mont(v, w);
      put(out, v, 4);
      return true;
    }
    if (nwords == 72) {
      for (int i = 0; i < 12; i++) {
        fq.from_mont(c, w + 6 * i);
        put(out, c, 6);
      }
      return true;
    }
    if (nwords != 12 && nwords != 24) return false;
    const int nc = (int)nwords / 12;                       // Fq coefficients per coordinate
    if (is_zero(w, (int)nwords)) {                         // the C ABI's identity
      out.insert(out.end(), nwords * 8, 0);
      out.back() |= 0x40;
      return true;
    }
    bool neg = false;
    for (int i = 0; i < 2 * nc; i++) {
      fq.from_mont(c, w + 6 * i);
      put(out, c, 6);
      if (i >= nc) {                                       // y: the HIGHEST non-zero coefficient decides (c1 before c0)
        if (!is_zero(c, 6)) neg = above_half(c);
      }
    }
    if (neg) out.back() |= 0x80;
    return true;
  }
};
const WordCodec& codec() {
  static const WordCodec c;
  return c;
}

}  // namespace

struct tb200_poseidon {
  SpongeBase* s = nullptr;
  std::mutex mu;
};

extern "C" {

int tb200_poseidon_new(int field, unsigned full_rounds, unsigned partial_rounds, uint64_t alpha, unsigned rate,
                       unsigned capacity, const uint64_t* ark, const uint64_t* mds, tb200_poseidon_t* out) {
  if (!out || !ark || !mds || rate == 0 || capacity == 0 || (full_rounds & 1) || alpha == 0 || rate + capacity > 16 ||
      full_rounds + partial_rounds == 0 || full_rounds + partial_rounds > 4096)
    return TB200_E_ARG;
  SpongeBase* s = nullptr;
  if (field == 0) s = make<4>(FR_MOD, 253, full_rounds, partial_rounds, alpha, rate, capacity, ark, mds);
  else if (field == 1) s = make<6>(FQ_MOD, 377, full_rounds, partial_rounds, alpha, rate, capacity, ark, mds);
  if (!s) return TB200_E_ARG;
  tb200_poseidon* h = new tb200_poseidon();
  h->s = s;
  *out = h;
  return 0;
}
int tb200_poseidon_reset(tb200_poseidon_t h) {
  if (!h) return TB200_E_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  h->s->reset();
  return 0;
}
int tb200_poseidon_absorb_bytes(tb200_poseidon_t h, const uint8_t* data, size_t len) {
  if (!h || (len && !data)) return TB200_E_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  h->s->absorb_bytes(data, len);
  return 0;
}
int tb200_poseidon_append_words(tb200_poseidon_t h, const uint64_t* words, size_t nwords) {
  if (!h || !words) return TB200_E_ARG;
  std::vector<uint8_t> bytes;
  bytes.reserve(nwords * 8);
  if (!codec().encode(bytes, words, nwords)) return TB200_E_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  h->s->absorb_bytes(bytes.data(), bytes.size());
  return 0;
}
int tb200_poseidon_absorb_native(tb200_poseidon_t h, const uint64_t* elems, size_t n) {
  if (!h || (n && !elems)) return TB200_E_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  return h->s->absorb_native(elems, n);
}
int tb200_poseidon_squeeze_native(tb200_poseidon_t h, uint64_t* out, size_t n) {
  if (!h || (n && !out)) return TB200_E_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  h->s->squeeze_native(out, n);
  return 0;
}
int tb200_poseidon_squeeze_fr(tb200_poseidon_t h, uint64_t out[4]) {
  if (!h || !out) return TB200_E_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  h->s->squeeze_fr(out);
  return 0;
}
int tb200_poseidon_limbs(tb200_poseidon_t h) { return h ? h->s->limbs() : 0; }
int tb200_poseidon_free(tb200_poseidon_t h) {
  if (!h) return TB200_E_ARG;
  delete h->s;
  delete h;
  return 0;
}

}  // extern "C"
