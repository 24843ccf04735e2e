// G2 unit of the engine: the point-touching stages of the MSM pipeline over Fq2 coordinates (kernels_g2.cuh) and the G2
// entry points. The sort stages (digits / scan / scatter) are shared with G1 and live in engine_g1.cu.
#define TB_NO_G1_KERNELS
#include <cstring>

#include "engine.h"
#include "kernels_g2.cuh"

using namespace tb;

namespace tbe {

int g2_accumulate(cudaStream_t st, uint32_t S_max, const uint32_t* entries, const uint32_t* starts, uint32_t B, uint32_t K,
                  const uint4* points, uint4* buckets, uint4* heads, int32_t* head_bucket) {
  LAUNCH(k_accumulate_g2, cdiv(S_max, 64), 64, st, entries, starts, B, K, points, buckets, heads, head_bucket);
  return 0;
}
int g2_fixup_round(cudaStream_t st, uint32_t S_max, const uint32_t* starts, uint32_t B, uint32_t K, uint32_t round,
                   uint4* heads, const int32_t* head_bucket) {
  LAUNCH(k_fixup_round_g2, cdiv(S_max, 64), 64, st, starts, B, K, round, heads, head_bucket);
  return 0;
}
int g2_fixup_final(cudaStream_t st, uint32_t S_max, const uint32_t* starts, uint32_t B, uint32_t K, uint4* buckets,
                   const uint4* heads, const int32_t* head_bucket) {
  LAUNCH(k_fixup_final_g2, cdiv(S_max, 64), 64, st, starts, B, K, buckets, heads, head_bucket);
  return 0;
}
int g2_reduce_pass(cudaStream_t st, const uint4* inS, const uint4* inW, const uint32_t* level0, uint4* outS, uint4* outW,
                   uint32_t L, int log2_ell, uint64_t n) {
  // few outputs (every single G2 MSM of the reference: sqrt(n)-sized): one warp per output with the lane-parallel group law
  if (n <= 8192 && !(getenv("TB200_G2_COMBINE") && !strcmp(getenv("TB200_G2_COMBINE"), "psi")))
    return g2_reduce_pass_coop(st, inS, inW, level0, outS, outW, L, log2_ell, n);
  LAUNCH(k_reduce_pass_g2, cdiv(n, 64), 64, st, inS, inW, level0, outS, outW, L, log2_ell, n);
  return 0;
}
int g2_finalize_single(cudaStream_t st, const uint4* group_w, int W, int c, uint4* d_out) {
  LAUNCH(k_finalize_single_g2, 1, 32, st, group_w, W, c, d_out);
  return 0;
}

}  // namespace tbe

using namespace tbe;

extern "C" {

// ---- G2 (MultilinearPC::open, commit_g2, G2 compress) -------------------------------------------
int tb200_msm_g2_dev(const void* d_bases, const void* d_scalars, size_t n, unsigned flags, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || (n && (!d_bases || !d_scalars))) return fail(TB200_E_ARG, "null pointer");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  return msm_dev(g, d_bases, d_scalars, n, flags, d_out, stream ? (cudaStream_t)stream : g.stream, nullptr, nullptr, true,
                 true);
}

int tb200_msm_g2(const uint64_t* bases, const uint64_t* scalars, size_t n, unsigned flags, uint64_t out[24]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && (!bases || !scalars))) return fail(TB200_E_ARG, "null pointer");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  uint4 *d_b = nullptr, *d_s = nullptr;
  if (n) {
    CU(cudaMallocAsync((void**)&d_b, n * 192, g.stream));
    CU(cudaMallocAsync((void**)&d_s, n * 32, g.stream));
    CU(cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_b, bases, n * 192, cudaMemcpyHostToDevice, g.stream));
  }
  int rc = msm_dev(g, d_b ? d_b : g.d_result, d_s ? d_s : g.d_result, n, flags, g.d_result, g.stream, nullptr, nullptr, true,
                   true);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(g.h_result, g.d_result, 192, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
    else memcpy(out, g.h_result, 192);
  } else {
    cudaStreamSynchronize(g.stream);
  }
  if (d_b) cudaFreeAsync(d_b, g.stream);
  if (d_s) cudaFreeAsync(d_s, g.stream);
  return rc;
}

int tb200_compress_g2(uint64_t* vec, size_t split, const uint64_t scaler[4], unsigned flags) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!vec || !scaler) return fail(TB200_E_ARG, "null pointer");
  if (split == 0) return 0;
  if (split >= (1u << 26)) return fail(TB200_E_LIMIT, "split too large");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  uint4* d_v = nullptr;
  uint32_t* d_k = nullptr;
  CU(cudaMallocAsync((void**)&d_v, 2 * split * 192, g.stream));
  CU(cudaMallocAsync((void**)&d_k, 32, g.stream));
  CU(cudaMemcpyAsync(d_v, vec, 2 * split * 192, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d_k, scaler, 32, cudaMemcpyHostToDevice, g.stream));
  LAUNCH(k_compress_g2, cdiv(split, 64), 64, g.stream, d_v, (uint32_t)split, d_k, (flags & TB200_SCALARS_MONT) ? 1 : 0);
  CU(cudaMemcpyAsync(vec, d_v, split * 192, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  cudaFreeAsync(d_v, g.stream);
  cudaFreeAsync(d_k, g.stream);
  return 0;
}

int tb200_test_g2_add(const uint64_t* p, const uint64_t* q, size_t n, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!p || !q || !out || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  cudaStream_t st = primary().stream;
  return with_buffers(p, n * 192, q, n * 192, out, n * 192, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    LAUNCH(k_test_g2_add, cdiv(n, 64), 64, st, (const uint4*)da, (const uint4*)db, (uint32_t)n, (uint4*)d1);
    return 0;
  });
}
int tb200_test_g2_mul(const uint64_t* p, const uint64_t* k, size_t n, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!p || !k || !out || n == 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  cudaStream_t st = primary().stream;
  return with_buffers(p, n * 192, k, n * 32, out, n * 192, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    LAUNCH(k_test_g2_mul, cdiv(n, 64), 64, st, (const uint4*)da, (const uint32_t*)db, (uint32_t)n, (uint4*)d1);
    return 0;
  });
}

}  // extern "C"
