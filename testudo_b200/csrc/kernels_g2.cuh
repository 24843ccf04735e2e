// G2 stages of the Pippenger pipeline: the digit / histogram / scan / scatter kernels of
// kernels.cuh are point-agnostic and are reused as they are; only the stages that touch points are restated here over
// Fq2 (g2.cuh): bucket accumulation, head fix-up, hierarchical reduction, Horner finalisation, and the per-element
// `compress` of MIPP's G2 vector. Layouts: affine point 192 B (12 uint4), XYZZ bucket 384 B (24 uint4).
//
// The reference's G2 MSMs are sqrt(n)-sized (<= 2^13 points: `MultilinearPC::open` src/sqrt_pst.rs:225, `commit_g2`
// src/mipp.rs:114), so these kernels keep canonical coordinates and out-of-line group operations (the accumulators
// live in local memory, L1-resident) instead of the shared-memory-slot / lazy-reduction design of the G1 hot loop.
#pragma once
#include "g2.cuh"
#include "kernels.cuh"

namespace tb {

__device__ __forceinline__ void load_affine2(Affine2& p, const uint4* src) {
  uint32_t* d = reinterpret_cast<uint32_t*>(&p);  // the whole packed object: 48 words
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint4 v = src[i];
    d[4 * i + 0] = v.x;
    d[4 * i + 1] = v.y;
    d[4 * i + 2] = v.z;
    d[4 * i + 3] = v.w;
  }
}
__device__ __forceinline__ void store_affine2(uint4* dst, const Affine2& p) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 12; i++) dst[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
}
__device__ __forceinline__ void load_xyzz2(Xyzz2& p, const uint4* src) {
  uint32_t* d = reinterpret_cast<uint32_t*>(&p);  // the whole packed object: 96 words
#pragma unroll
  for (int i = 0; i < 24; i++) {
    uint4 v = src[i];
    d[4 * i + 0] = v.x;
    d[4 * i + 1] = v.y;
    d[4 * i + 2] = v.z;
    d[4 * i + 3] = v.w;
  }
}
__device__ __forceinline__ void store_xyzz2(uint4* dst, const Xyzz2& p) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 24; i++) dst[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
}
static_assert(sizeof(Affine2) == 192 && sizeof(Xyzz2) == 384, "packed layouts");

static __device__ __noinline__ void xyzz2_add_ni(Xyzz2* p, const Xyzz2* q) { xyzz2_add(*p, *q); }
static __device__ __noinline__ void xyzz2_dbl_ni(Xyzz2* p) { xyzz2_dbl(*p); }
static __device__ __noinline__ void xyzz2_madd_ni(Xyzz2* p, const Affine2* q) { xyzz2_madd(*p, *q); }
static __device__ __noinline__ void xyzz2_to_affine_ni(Affine2* r, const Xyzz2* p) { xyzz2_to_affine(*r, *p); }

#ifndef TB_NO_G2_KERNELS  // see kernels.cuh: a __global__ function lives in exactly one unit
// same segment scheme as k_accumulate (kernels.cuh): thread t owns sorted entries [t K, (t+1) K)
__global__ void __launch_bounds__(64) k_accumulate_g2(const uint32_t* __restrict__ entries,
                                                      const uint32_t* __restrict__ bucket_start, uint32_t B, uint32_t K,
                                                      const uint4* __restrict__ points, uint4* __restrict__ buckets,
                                                      uint4* __restrict__ heads, int32_t* __restrict__ head_bucket) {
  const uint32_t M = __ldg(bucket_start + B);
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t lo64 = t * K;
  if (lo64 >= M) return;
  const uint32_t lo = (uint32_t)lo64;
  const uint32_t hi = (uint32_t)min((uint64_t)M, lo64 + K);
  uint32_t l = 0, r = B;
  while (r - l > 1) {
    uint32_t mid = l + ((r - l) >> 1);
    if (__ldg(bucket_start + mid) <= lo) l = mid;
    else r = mid;
  }
  uint32_t b = l;
  bool is_head = __ldg(bucket_start + b) < lo;
  head_bucket[t] = is_head ? (int32_t)b : -1;
  uint32_t end_b = __ldg(bucket_start + b + 1);
  Xyzz2 acc;
  xyzz2_set_inf(acc);
  for (uint32_t pos = lo; pos < hi; pos++) {
    if (pos == end_b) {
      store_xyzz2(is_head ? heads + 24 * t : buckets + 24 * (uint64_t)b, acc);
      is_head = false;
      xyzz2_set_inf(acc);
      do {
        b++;
        end_b = __ldg(bucket_start + b + 1);
      } while (end_b == pos);
    }
    const uint32_t e = __ldg(entries + pos);
    Affine2 q;
    load_affine2(q, points + 12 * (uint64_t)(e & 0x7fffffffu));
    if (e >> 31) fq2_neg(q.y, q.y);
    xyzz2_madd_ni(&acc, &q);
  }
  store_xyzz2(is_head ? heads + 24 * t : buckets + 24 * (uint64_t)b, acc);
}

__global__ void __launch_bounds__(64) k_fixup_round_g2(const uint32_t* __restrict__ bucket_start, uint32_t B, uint32_t K,
                                                       uint32_t round, uint4* __restrict__ heads,
                                                       const int32_t* __restrict__ head_bucket) {
  const uint32_t M = bucket_start[B];
  const uint64_t S = ((uint64_t)M + K - 1) / K;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 || t >= S) return;
  const int32_t hb = head_bucket[t];
  if (hb < 0) return;
  const uint64_t first = bucket_start[hb] / K + 1;
  const uint64_t i = t - first;
  const uint64_t stride = 1ull << round;
  if (i & (2 * stride - 1)) return;
  const uint64_t u = t + stride;
  if (u >= S || u * K >= bucket_start[hb + 1]) return;
  Xyzz2 acc, h;
  load_xyzz2(acc, heads + 24 * t);
  load_xyzz2(h, heads + 24 * u);
  xyzz2_add_ni(&acc, &h);
  store_xyzz2(heads + 24 * t, acc);
}
__global__ void __launch_bounds__(64) k_fixup_final_g2(const uint32_t* __restrict__ bucket_start, uint32_t B, uint32_t K,
                                                       uint4* __restrict__ buckets, const uint4* __restrict__ heads,
                                                       const int32_t* __restrict__ head_bucket) {
  const uint32_t M = bucket_start[B];
  const uint64_t S = ((uint64_t)M + K - 1) / K;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 || t >= S) return;
  const int32_t hb = head_bucket[t];
  if (hb < 0) return;
  if (t != bucket_start[hb] / K + 1) return;
  Xyzz2 acc, h;
  load_xyzz2(acc, buckets + 24 * (uint64_t)hb);
  load_xyzz2(h, heads + 24 * t);
  xyzz2_add_ni(&acc, &h);
  store_xyzz2(buckets + 24 * (uint64_t)hb, acc);
}

// (S, W) hierarchical running sums exactly as k_reduce_pass
__global__ void __launch_bounds__(64) k_reduce_pass_g2(const uint4* __restrict__ inS, const uint4* __restrict__ inW,
                                                       const uint32_t* __restrict__ bucket_start, uint4* __restrict__ outS,
                                                       uint4* __restrict__ outW, uint32_t L, int log2_ell,
                                                       uint64_t total_out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_out) return;
  Xyzz2 run, acc, x;
  xyzz2_set_inf(run);
  xyzz2_set_inf(acc);
  for (int i = (int)L - 1; i >= 0; i--) {
    const uint64_t idx = t * L + i;
    bool empty = false;
    if (bucket_start) empty = bucket_start[idx + 1] == bucket_start[idx];
    if (!empty) {
      load_xyzz2(x, inS + 24 * idx);
      xyzz2_add_ni(&run, &x);
    }
    if (i > 0) xyzz2_add_ni(&acc, &run);
  }
  store_xyzz2(outS + 24 * t, run);
  for (int k = 0; k < log2_ell; k++) xyzz2_dbl_ni(&acc);
  if (inW) {
    xyzz2_set_inf(run);
    for (int i = 0; i < (int)L; i++) {
      load_xyzz2(x, inW + 24 * (t * L + i));
      xyzz2_add_ni(&run, &x);
    }
  }
  xyzz2_add_ni(&acc, &run);
  store_xyzz2(outW + 24 * t, acc);
}

__global__ void k_finalize_single_g2(const uint4* __restrict__ group_w, int W, int c, uint4* __restrict__ out_affine) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Xyzz2 total, x;
  xyzz2_set_inf(total);
  for (int w = W - 1; w >= 0; w--) {
    load_xyzz2(x, group_w + 24 * w);
    xyzz2_add_ni(&total, &x);
    if (w > 0)
      for (int k = 0; k < c; k++) xyzz2_dbl_ni(&total);
  }
  Affine2 a;
  xyzz2_to_affine_ni(&a, &total);
  store_affine2(out_affine, a);
}

// MIPP `compress` on a G2 vector (src/mipp.rs:133, 354-367): a[i] <- a[i] + scaler * a[split + i]
__global__ void __launch_bounds__(64) k_compress_g2(uint4* __restrict__ a, uint32_t split,
                                                    const uint32_t* __restrict__ scaler /* 8 limbs */, int mont) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= split) return;
  uint32_t k[8];
#pragma unroll
  for (int j = 0; j < 8; j++) k[j] = scaler[j];
  if (mont) {
    uint32_t cnv[8];
    mont_to_canonical<FrParams>(cnv, k);
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = cnv[j];
  }
  Affine2 l, r;
  load_affine2(l, a + 12 * (uint64_t)i);
  load_affine2(r, a + 12 * ((uint64_t)split + i));
  Xyzz2 acc;
  xyzz2_set_inf(acc);
  bool started = false;
  for (int limb = 7; limb >= 0; limb--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) xyzz2_dbl_ni(&acc);
      if ((k[limb] >> bit) & 1) {
        xyzz2_madd_ni(&acc, &r);
        started = true;
      }
    }
  }
  xyzz2_madd_ni(&acc, &l);
  Affine2 o;
  xyzz2_to_affine_ni(&o, &acc);
  store_affine2(a + 12 * (uint64_t)i, o);
}

// field / group unit-test kernels (tests/test_gpu_g2.py)
__global__ void k_test_g2_add(const uint4* p, const uint4* q, uint32_t n, uint4* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine2 a, b, o;
  load_affine2(a, p + 12 * (uint64_t)i);
  load_affine2(b, q + 12 * (uint64_t)i);
  Xyzz2 acc;
  xyzz2_set_inf(acc);
  xyzz2_madd_ni(&acc, &a);
  xyzz2_madd_ni(&acc, &b);
  xyzz2_to_affine_ni(&o, &acc);
  store_affine2(out + 12 * (uint64_t)i, o);
}
__global__ void k_test_g2_mul(const uint4* p, const uint32_t* k, uint32_t n, uint4* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine2 a, o;
  load_affine2(a, p + 12 * (uint64_t)i);
  Xyzz2 acc;
  xyzz2_set_inf(acc);
  bool started = false;
  for (int limb = 7; limb >= 0; limb--) {
    const uint32_t kw = k[8 * (uint64_t)i + limb];
    for (int bit = 31; bit >= 0; bit--) {
      if (started) xyzz2_dbl_ni(&acc);
      if ((kw >> bit) & 1) {
        xyzz2_madd_ni(&acc, &a);
        started = true;
      }
    }
  }
  xyzz2_to_affine_ni(&o, &acc);
  store_affine2(out + 12 * (uint64_t)i, o);
}

#endif  // TB_NO_G2_KERNELS
}  // namespace tb
