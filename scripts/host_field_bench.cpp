// Host-side microbenchmark of the Poseidon sponge's field arithmetic (csrc/poseidon_host.cpp): ns per Montgomery product
// over Fq (interleaved CIOS vs the generic loop), per three-term dot product, and per permutation of the width-3 sponge.
//   g++ -O3 -std=c++17 -I include scripts/host_field_bench.cpp -o /tmp/hfb && /tmp/hfb
#include "../testudo_b200/csrc/poseidon_host.cpp"
#include <chrono>
#include <cstdio>
static double ns(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b, double n) {
  return std::chrono::duration<double, std::nano>(b - a).count() / n;
}
int main() {
  Field<6> F;
  const uint64_t q[6] = {0x8508c00000000001ull, 0x170b5d4430000000ull, 0x1ef3622fba094800ull,
                         0x1a22d9f300f5138full, 0xc63b05c06ca1493bull, 0x01ae3a4617c510eaull};
  F.init(q);
  uint64_t a[6] = {1, 2, 3, 4, 5, 6}, b[6] = {7, 8, 9, 10, 11, 12}, r[6];
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < 2000000; i++) F.mul(a, a, b);
  auto t1 = std::chrono::steady_clock::now();
  printf("mul (interleaved CIOS): %.1f ns   [%llx]\n", ns(t0, t1, 2e6), (unsigned long long)a[0]);
  t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < 2000000; i++) F.mul_generic(a, a, b);
  t1 = std::chrono::steady_clock::now();
  printf("mul (generic loop):     %.1f ns   [%llx]\n", ns(t0, t1, 2e6), (unsigned long long)a[0]);
  t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < 1000000; i++) {
    F.dot3(r, a, b, b, b, a, a);
    a[0] ^= r[0] & 1;
  }
  t1 = std::chrono::steady_clock::now();
  printf("dot3 (one reduction):   %.1f ns   [%llx]\n", ns(t0, t1, 1e6), (unsigned long long)r[0]);
  return 0;
}
