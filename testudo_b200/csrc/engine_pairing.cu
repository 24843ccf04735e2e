// Pairing unit of the engine: the Fq12 tower, Miller loops, final exponentiation (kernels_pairing.cuh) and the folds /
// window combine that run over the curve endomorphisms, plus the entry points built on them (pairing products, the G2
// side of MIPP, a whole MIPP round).
#define TB_NO_G1_KERNELS
#define TB_NO_G2_KERNELS
#include <algorithm>
#include <cstring>

#include "engine.h"
#include "glv_host.h"
#include "kernels_pairing.cuh"

using namespace tb;

namespace tbe {

int g2_finalize_single_glv(cudaStream_t st, const uint4* group_w, int W, int c, uint4* fin_scratch, uint4* d_out) {
  // TB200_G2_COMBINE=psi selects the earlier kernel (four endomorphism dimensions per window, single-thread group law)
  static const bool psi = getenv("TB200_G2_COMBINE") && !strcmp(getenv("TB200_G2_COMBINE"), "psi");
  if (psi) LAUNCH(k_finalize_single_g2_glv, 1, 384, st, group_w, W, c, fin_scratch, d_out);
  else LAUNCH(k_finalize_single_g2_coop, 1, 32, st, group_w, W, c, d_out);
  return 0;
}

int g2_reduce_pass_coop(cudaStream_t st, const uint4* inS, const uint4* inW, const uint32_t* level0, uint4* outS,
                        uint4* outW, uint32_t L, int log2_ell, uint64_t n) {
  LAUNCH(k_reduce_pass_g2_coop, (uint32_t)n, 32, st, inS, inW, level0, outS, outW, L, log2_ell, n);
  return 0;
}

int g1_fold_pre(cudaStream_t st, const uint4* a, uint32_t first, uint32_t count, uint4* mult) {
  LAUNCH(k_fold_pre_g1, cdiv(count, 128), 128, st, a, first, count, mult);
  return 0;
}
// lanes per element of the two-phase folds: a warp per element while the vector is short (latency), 8 lanes when it is
// long (four elements per warp: throughput)
static inline int fold_lanes(uint32_t split) { return split >= 1024 ? 8 : 32; }
int g1_fold_apply(cudaStream_t st, const uint16_t* d_sel, uint4* a, uint32_t split, const uint4* mult) {
  if (fold_lanes(split) == 8) LAUNCH(k_fold_apply_g1<8>, cdiv((uint64_t)split * 8, 128), 128, st, a, split, d_sel, mult);
  else LAUNCH(k_fold_apply_g1<32>, cdiv((uint64_t)split * 32, 128), 128, st, a, split, d_sel, mult);
  return 0;
}
int g1_fold_glv(cudaStream_t st, const uint32_t* d_scaler, int mont, uint32_t* d_digits, uint4* a, uint32_t split) {
  LAUNCH(k_glv2_digits, 1, 32, st, d_scaler, mont, d_digits);
  LAUNCH(k_compress_g1_glv, cdiv(split, 128), 128, st, a, split, d_digits);
  return 0;
}

// Miller values of `n` pairs (g2 index = j ^ xor_mask) -> `segs` products (segment s = pairs [s n/segs, (s+1) n/segs))
// -> final exponentiation of each, written to d_out (segs x 576 B). Everything is enqueued on `st`; the scratch is
// stream-ordered. `after_miller`, if given, is recorded once the Miller kernel (the only reader of g1 / g2) is enqueued.
// `d_gt_in` != nullptr: skip the Miller stage, the n inputs are Fq12 values (per-GPU Miller products to be combined).
// `final_exp` = false: stop after the product tree (a partial Miller product for a sharded pairing product).
// Stage marks ("miller", "gt_product", "final_exp") are appended to the context's list; the caller finishes them.
int pairing_products(Ctx& g, const uint4* d_g1, const uint4* d_g2, uint32_t n, uint32_t xor_mask, uint32_t segs,
                     uint4* d_out, cudaStream_t st, cudaEvent_t after_miller, const uint4* d_gt_in, bool final_exp) {
  if (n == 0) {
    LAUNCH(k_fq12_set_one, 1, 32, st, d_out, segs);
    return 0;
  }
  uint32_t len = n / segs;
  uint4 *buf_a = nullptr, *buf_b = nullptr;
  CU(cudaMallocAsync((void**)&buf_a, (size_t)n * 576, st));
  CU(cudaMallocAsync((void**)&buf_b, (size_t)segs * cdiv(len, FQ12_FAN) * 576 + 576, st));
  if (mark(g, st, "pairing_begin")) return 1;
  // below ~2 waves of resident CTAs one CTA per pair (latency-bound); above, one thread per pair
  const bool coop = n <= (uint32_t)E.pairing_coop_max;
  // Miller kernels (measured, scripts/time_pairing.py): up to 512 pairs the loop's LATENCY counts and the pipelined
  // kernel is used (three warps per pair: the point chain next to the f chain, 0.50 ms); above, the stage is
  // throughput-bound and one warp takes TWO pairs with a shared accumulator (k_miller_duo; needs an even segment length).
  // The chain kernels (product tree, final exponentiation) keep two warps. tb200_set_pairing_team forces 96 (pipelined),
  // 64 / 32 (one pair per two-warp / one-warp CTA, sequential loop) or 33 (two pairs per warp).
  const int forced = E.pairing_team;
  int miller_team = forced ? forced : (n > 512 ? 33 : 96);
  if (miller_team == 33 && (len & 1)) miller_team = 32;
  const int chain_team = (forced == 32 || forced == 64) ? forced : 64;
  if (d_gt_in) CU(cudaMemcpyAsync(buf_a, d_gt_in, (size_t)n * 576, cudaMemcpyDeviceToDevice, st));
  else if (coop && miller_team == 96) LAUNCH(k_miller_pipe, n, MP_THREADS, st, d_g1, d_g2, xor_mask, buf_a);
  else if (coop && miller_team == 33) {
    LAUNCH(k_miller_duo, n / 2, 32, st, d_g1, d_g2, n, xor_mask, buf_a);
    len /= 2;   // one value per two pairs; pairs 2b, 2b + 1 lie in the same segment
  } else if (coop) LAUNCH(k_miller_coop, n, miller_team, st, d_g1, d_g2, xor_mask, buf_a);
  else LAUNCH(k_miller, cdiv(n, 32), 32, st, d_g1, d_g2, n, xor_mask, buf_a);
  if (after_miller) CU(cudaEventRecord(after_miller, st));
  if (mark(g, st, "miller")) return 1;
  uint4 *cur = buf_a, *nxt = buf_b;
  while (len > 1) {
    const uint32_t m = cdiv(len, FQ12_FAN);
    if ((uint64_t)m * segs <= 4096) LAUNCH(k_fq12_prod_level_coop, dim3(m, segs), chain_team, st, cur, len, m, nxt);
    else LAUNCH(k_fq12_prod_level, dim3(cdiv(m, 32), segs), 32, st, cur, len, m, nxt);
    std::swap(cur, nxt);
    len = m;
  }
  if (mark(g, st, "gt_product")) return 1;
  if (final_exp) LAUNCH(k_final_exp, segs, chain_team, st, cur, d_out);
  else CU(cudaMemcpyAsync(d_out, cur, (size_t)segs * 576, cudaMemcpyDeviceToDevice, st));
  if (mark(g, st, "final_exp")) return 1;
  CU(cudaFreeAsync(buf_a, st));
  CU(cudaFreeAsync(buf_b, st));
  return 0;
}

}  // namespace tbe

using namespace tbe;

namespace {
// one pairing-product call on the primary device over device pointers
int pairing_dev_call(const void* d_g1, const void* d_g2, size_t n, void* d_out, void* stream, const void* d_gt_in,
                     bool final_exp) {
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
  g.marks.clear();
  int rc = pairing_products(g, (const uint4*)d_g1, (const uint4*)d_g2, (uint32_t)n, 0, 1, (uint4*)d_out, st, nullptr,
                            (const uint4*)d_gt_in, final_exp);
  return rc ? rc : finish_marks(g, st);
}
// out[i] = in[i]^exps[i]: up to GT_POW_COOP_MAX elements one TEAM each (the chain's latency is what a verifier waits
// for: ~380 cooperative Fq12 operations instead of ~17 ms on one thread), larger batches one thread each
constexpr size_t GT_POW_COOP_MAX = 8192;
int gt_pow_launch(cudaStream_t st, const uint4* d_in, const uint32_t* d_exps, size_t n, int mont, uint4* d_out) {
  if (n <= GT_POW_COOP_MAX) LAUNCH(k_fq12_pow_coop, (uint32_t)n, W12_THREADS, st, d_in, d_exps, mont, d_out);
  else LAUNCH(k_fq12_pow, cdiv(n, 32), 32, st, d_in, d_exps, (uint32_t)n, mont, d_out);
  return 0;
}
}  // namespace

extern "C" {

int tb200_multi_pairing_dev(const void* d_g1_xy, const void* d_g2, size_t n, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || (n && (!d_g1_xy || !d_g2))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "too many pairs");
  return pairing_dev_call(d_g1_xy, d_g2, n, d_out, stream, nullptr, true);
}
int tb200_miller_product_dev(const void* d_g1_xy, const void* d_g2, size_t n, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || (n && (!d_g1_xy || !d_g2))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "too many pairs");
  return pairing_dev_call(d_g1_xy, d_g2, n, d_out, stream, nullptr, false);
}
// the combination of a sharded pairing product: the product of `n` partial Miller values, then ONE final exponentiation
int tb200_gt_product_final_exp_dev(const void* d_parts, size_t n, void* d_out, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!d_out || (n && !d_parts)) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 20)) return fail(TB200_E_LIMIT, "too many partial products");
  return pairing_dev_call(nullptr, nullptr, n, d_out, stream, d_parts, true);
}
int tb200_gt_product_final_exp(const uint64_t* parts, size_t n, uint64_t out[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && !parts)) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 20)) return fail(TB200_E_LIMIT, "too many partial products");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  uint4 *d_in = nullptr, *d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, 576, g.stream));
  if (n) {
    CU(cudaMallocAsync((void**)&d_in, n * 576, g.stream));
    CU(cudaMemcpyAsync(d_in, parts, n * 576, cudaMemcpyHostToDevice, g.stream));
  }
  int rc = pairing_dev_call(nullptr, nullptr, n, d_o, nullptr, d_in, true);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out, d_o, 576, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "result copy failed: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(g.stream);
  }
  if (d_in) cudaFreeAsync(d_in, g.stream);
  cudaFreeAsync(d_o, g.stream);
  return rc;
}

// ---- MIPP's G2 commitment key ------------------------------------------------------------------------------------------
int tb200_mipp_g2_begin(const uint64_t* h_vec, size_t n, unsigned flags, tb200_mipp_g2_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h_vec || !out || n == 0) return fail(TB200_E_ARG, "bad arguments");
  if (n & (n - 1)) return fail(TB200_E_ARG, "MIPP vectors must have a power-of-two length (n = %zu)", n);
  if (n >= (1u << 26)) return fail(TB200_E_LIMIT, "vector too long");
  CU(cudaSetDevice(primary().device));
  tb200_mipp_g2* m = new tb200_mipp_g2();
  m->n = (uint32_t)n;
  m->flags = flags;
  cudaError_t e = cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking);
  cudaStream_t m_st = m->st;
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&m->h, n * 192, m_st);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&m->scal, 64 * 32, m_st);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&m->digits, 64 * 32, m_st);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&m->scal_host, 64 * 32);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->h, h_vec, n * 192, cudaMemcpyHostToDevice, m_st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(m_st);  // h_vec is only borrowed for the duration of the call
  if (e == cudaSuccess && n >= 2 && g2_fold_mult_bytes(n) <= FOLD_MULT_BYTES_MAX) {
    // two-phase fold: the multiples of the first round's right half start right away, next to the first cross values
    e = cudaMallocAsync((void**)&m->mult, g2_fold_mult_bytes(n), m_st);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->pre_st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_pre, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_fold, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&m->sel, 64 * glv::SEL_MAX * 2, m_st);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&m->sel_host, 64 * glv::SEL_MAX * 2);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m_st);   // the stream-ordered allocations exist for every stream
    if (e == cudaSuccess) {
      k_fold_pre_g2<<<cdiv(n / 2, 64), 64, 0, m->pre_st>>>(m->h, (uint32_t)(n / 2), (uint32_t)(n / 2), m->mult);
      g_launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(m->ev_pre, m->pre_st);
  }
  if (e != cudaSuccess) {
    cudaFreeAsync(m->mult, m_st);
    cudaFreeAsync(m->sel, m_st);
    cudaFreeHost(m->sel_host);
    if (m->pre_st) cudaStreamDestroy(m->pre_st);
    if (m->ev_pre) cudaEventDestroy(m->ev_pre);
    if (m->ev_fold) cudaEventDestroy(m->ev_fold);
    cudaFreeAsync(m->h, m_st);
    cudaFreeAsync(m->scal, m_st);
    cudaFreeAsync(m->digits, m_st);
    cudaFreeHost(m->scal_host);
    if (m->st) cudaStreamDestroy(m->st);
    delete m;
    return fail((int)e, "mipp_g2_begin failed: %s", cudaGetErrorString(e));
  }
  *out = m;
  return 0;
}
size_t tb200_mipp_g2_len(tb200_mipp_g2_t h) { return h ? h->n : 0; }
/* h[i] <- h[i] + c_inv * h[split + i]; returns after ENQUEUEING on the handle's own stream */
int tb200_mipp_g2_fold(tb200_mipp_g2_t h, const uint64_t c_inv[4]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h || !c_inv) return fail(TB200_E_ARG, "null pointer");
  if (h->n < 2) return fail(TB200_E_STATE, "MIPP vector is already folded to length 1");
  if (h->round >= 64) return fail(TB200_E_LIMIT, "too many rounds");
  CU(cudaSetDevice(primary().device));
  const uint32_t split = h->n / 2;
  cudaStream_t m_st = h->st;
  if (h->mult) {
    // the scalar is decomposed on the host (one value per round): the device gets the list of stored multiples to add
    uint16_t* hs = h->sel_host + (size_t)glv::SEL_MAX * h->round;
    uint16_t* ds = h->sel + (size_t)glv::SEL_MAX * h->round;
    glv::select_g2(c_inv, (h->flags & TB200_SCALARS_MONT) != 0, hs);
    CU(cudaMemcpyAsync(ds, hs, (size_t)(hs[0] + 1) * 2, cudaMemcpyHostToDevice, m_st));
    CU(cudaStreamWaitEvent(m_st, h->ev_pre, 0));             // the multiples of this round's right half
    if (fold_lanes(split) == 8)
      LAUNCH(k_fold_apply_g2<8>, cdiv((uint64_t)split * 8, 64), 64, m_st, h->h, split, ds, h->mult);
    else
      LAUNCH(k_fold_apply_g2<32>, cdiv((uint64_t)split * 32, 64), 64, m_st, h->h, split, ds, h->mult);
    if (split >= 2) {                                        // phase A of the next round, off the critical path
      CU(cudaEventRecord(h->ev_fold, m_st));
      CU(cudaStreamWaitEvent(h->pre_st, h->ev_fold, 0));
      LAUNCH(k_fold_pre_g2, cdiv(split / 2, 64), 64, h->pre_st, h->h, split / 2, split / 2, h->mult);
      CU(cudaEventRecord(h->ev_pre, h->pre_st));
    }
  } else {
    memcpy(h->scal_host + 8 * h->round, c_inv, 32);
    CU(cudaMemcpyAsync(h->scal + 8 * h->round, h->scal_host + 8 * h->round, 32, cudaMemcpyHostToDevice, m_st));
    // 4-dimensional decomposition over the twisted Frobenius (kernels_pairing.cuh): 64 doublings instead of 253
    LAUNCH(k_glv4_digits, 1, 32, m_st, h->scal + 8 * h->round, (h->flags & TB200_SCALARS_MONT) ? 1 : 0,
           h->digits + 8 * h->round);
    LAUNCH(k_compress_g2_glv4w, cdiv(split, 32), 128, m_st, h->h, split, h->digits + 8 * h->round);
  }
  h->round++;
  h->n = split;
  return 0;
}
/* the current vector (len() points); waits for the enqueued folds */
int tb200_mipp_g2_read(tb200_mipp_g2_t h, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!h || !out) return fail(TB200_E_ARG, "null pointer");
  CU(cudaSetDevice(primary().device));
  cudaStream_t m_st = h->st;
  CU(cudaMemcpyAsync(out, h->h, (size_t)h->n * 192, cudaMemcpyDeviceToHost, m_st));
  CU(cudaStreamSynchronize(m_st));
  return 0;
}
int tb200_mipp_g2_end(tb200_mipp_g2_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!h) return fail(TB200_E_ARG, "null handle");
  if (E.ready) {
    cudaSetDevice(primary().device);
    cudaStreamSynchronize(h->st);
    if (h->pre_st) {
      cudaStreamSynchronize(h->pre_st);
      cudaStreamDestroy(h->pre_st);
      cudaEventDestroy(h->ev_pre);
      cudaEventDestroy(h->ev_fold);
    }
    cudaFreeAsync(h->mult, h->st);
    cudaFreeAsync(h->sel, h->st);
    cudaFreeHost(h->sel_host);
    cudaFreeAsync(h->h, h->st);
    cudaFreeAsync(h->scal, h->st);
    cudaFreeAsync(h->digits, h->st);
    cudaFreeHost(h->scal_host);
    cudaStreamDestroy(h->st);
  }
  delete h;
  return 0;
}

int tb200_mipp_pairing_cross(tb200_mipp_t a, tb200_mipp_g2_t h, uint64_t comm_t_l[72], uint64_t comm_t_r[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !h || !comm_t_l || !comm_t_r) return fail(TB200_E_ARG, "null pointer");
  if (a->n != h->n) return fail(TB200_E_ARG, "MIPP vectors differ in length (%u vs %u)", a->n, h->n);
  if (a->n < 2) return fail(TB200_E_STATE, "MIPP vectors are already folded to length 1");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  const uint32_t n = a->n, split = n / 2;
  // the G2 key is folded on its own stream: wait for the folds enqueued so far, and make later folds (which rewrite h
  // in place) wait for this round's Miller kernel
  CU(cudaEventRecord(g.ev_join, h->st));
  CU(cudaStreamWaitEvent(g.stream, g.ev_join, 0));
  uint4* d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, 2 * 576, g.stream));
  g.marks.clear();
  int rc = pairing_products(g, a->a, h->h, n, split, 2, d_o, g.stream, g.ev_join, nullptr, true);
  if (rc == 0) rc = finish_marks(g, g.stream);
  if (rc == 0) {
    cudaError_t e = cudaStreamWaitEvent(h->st, g.ev_join, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(comm_t_l, d_o, 576, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(comm_t_r, d_o + 36, 576, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) rc = fail((int)e, "pairing result copy failed: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(g.stream);
  }
  cudaFreeAsync(d_o, g.stream);
  return rc;
}

// One MIPP round's four values in one call: the two cross MSMs (library streams) and the two cross pairing products
// (their own stream) run side by side; a single host synchronisation at the end.
int tb200_mipp_cross_all(tb200_mipp_t a, tb200_mipp_g2_t h, uint64_t comm_u_l[12], uint64_t comm_u_r[12],
                         uint64_t comm_t_l[72], uint64_t comm_t_r[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !h || !comm_u_l || !comm_u_r || !comm_t_l || !comm_t_r) return fail(TB200_E_ARG, "null pointer");
  if (a->n != h->n) return fail(TB200_E_ARG, "MIPP vectors differ in length (%u vs %u)", a->n, h->n);
  if (a->n < 2) return fail(TB200_E_STATE, "MIPP vectors are already folded to length 1");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  if (!g.pair_stream) {
    CU(cudaStreamCreateWithFlags(&g.pair_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&g.ev_pair, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g.ev_pair2, cudaEventDisableTiming));
  }
  const uint32_t n = a->n, split = n / 2;
  const bool prof = g.profiling;
  g.profiling = false;
  // pairing stream: behind the G1 folds (library stream) and the G2 folds (the key's stream)
  CU(cudaEventRecord(g.ev_pair, g.stream));
  CU(cudaStreamWaitEvent(g.pair_stream, g.ev_pair, 0));
  CU(cudaStreamWaitEvent(g.stream2, g.ev_pair, 0));  // the second cross MSM reads the folded a, y too
  CU(cudaEventRecord(g.ev_pair2, h->st));
  CU(cudaStreamWaitEvent(g.pair_stream, g.ev_pair2, 0));
  uint4* d_o = nullptr;
  CU(cudaMallocAsync((void**)&d_o, 2 * 576, g.pair_stream));
  int rc = pairing_products(g, a->a, h->h, n, split, 2, d_o, g.pair_stream, g.ev_pair, nullptr, true);
  if (rc == 0) {
    cudaError_t e = cudaStreamWaitEvent(h->st, g.ev_pair, 0);  // later G2 folds rewrite h: behind this round's Miller kernel
    if (e != cudaSuccess) rc = fail((int)e, "event wait failed: %s", cudaGetErrorString(e));
  }
  // cross MSMs as in tb200_mipp_g1_cross. They are off the round's critical path (the Miller loops and the final
  // exponentiation next to them take longer), so what matters is how little they disturb those: above 64 points the sort
  // pipeline (a handful of resident warps) is used instead of the Straus path (one long-running warp per 8 points).
  const int small_keep = E.small_msm_max;
  E.small_msm_max = std::min(small_keep, E.mipp_cross_small_max);
  if (rc == 0) rc = msm_dev(g, a->a, a->y + 8 * (size_t)split, split, a->flags, g.d_result, g.stream, nullptr, nullptr, false);
  if (rc == 0)
    rc = msm_dev(g, a->a + 6 * (size_t)split, a->y, split, a->flags, g.d_result + 6, g.stream2, nullptr, &g.arena2, false);
  E.small_msm_max = small_keep;
  g.profiling = prof;
  if (rc == 0) {
    cudaError_t e = cudaEventRecord(g.ev_join, g.stream2);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g.stream, g.ev_join, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g.h_result, g.d_result, 192, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(comm_t_l, d_o, 576, cudaMemcpyDeviceToHost, g.pair_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(comm_t_r, d_o + 36, 576, cudaMemcpyDeviceToHost, g.pair_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.pair_stream);
    if (e != cudaSuccess) rc = fail((int)e, "MIPP round failed: %s", cudaGetErrorString(e));
    else {
      memcpy(comm_u_l, g.h_result, 96);
      memcpy(comm_u_r, (char*)g.h_result + 96, 96);
    }
  } else {
    cudaStreamSynchronize(g.stream);
    cudaStreamSynchronize(g.stream2);
    cudaStreamSynchronize(g.pair_stream);
  }
  cudaFreeAsync(d_o, g.pair_stream);
  return rc;
}

int tb200_gt_pow(const uint64_t* bases, const uint64_t* exps, size_t n, unsigned flags, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (n == 0) return 0;
  if (!bases || !exps || !out) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 22)) return fail(TB200_E_LIMIT, "too many elements");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  StreamScratch sc(g.stream);
  uint4 *d_b = nullptr, *d_o = nullptr;
  uint32_t* d_e = nullptr;
  CU(sc.alloc(&d_b, n * 576));
  CU(sc.alloc(&d_o, n * 576));
  CU(sc.alloc(&d_e, n * 32));
  CU(cudaMemcpyAsync(d_b, bases, n * 576, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d_e, exps, n * 32, cudaMemcpyHostToDevice, g.stream));
  if (int rc = gt_pow_launch(g.stream, d_b, d_e, n, (flags & TB200_SCALARS_MONT) ? 1 : 0, d_o)) return rc;
  CU(cudaMemcpyAsync(out, d_o, n * 576, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  sc.done = true;
  return 0;
}

// prod_i bases[i]^exps[i]: the TC half of the verifier's fold / reduce (src/mipp.rs:240-271) -- the powers on one team
// each, then the product tree of the pairing engine (no final exponentiation); n == 0 yields 1
int tb200_gt_multi_pow(const uint64_t* bases, const uint64_t* exps, size_t n, unsigned flags, uint64_t out[72]) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!out || (n && (!bases || !exps))) return fail(TB200_E_ARG, "null pointer");
  if (n >= (1u << 20)) return fail(TB200_E_LIMIT, "too many elements");
  Ctx& g = primary();
  CU(cudaSetDevice(g.device));
  StreamScratch sc(g.stream);
  uint4 *d_b = nullptr, *d_p = nullptr, *d_o = nullptr;
  uint32_t* d_e = nullptr;
  CU(sc.alloc(&d_o, 576));
  if (n) {
    CU(sc.alloc(&d_b, n * 576));
    CU(sc.alloc(&d_p, n * 576));
    CU(sc.alloc(&d_e, n * 32));
    CU(cudaMemcpyAsync(d_b, bases, n * 576, cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(d_e, exps, n * 32, cudaMemcpyHostToDevice, g.stream));
    if (int rc = gt_pow_launch(g.stream, d_b, d_e, n, (flags & TB200_SCALARS_MONT) ? 1 : 0, d_p)) return rc;
  }
  if (int rc = pairing_dev_call(nullptr, nullptr, n, d_o, nullptr, d_p, false)) return rc;
  CU(cudaMemcpyAsync(out, d_o, 576, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  sc.done = true;
  return 0;
}

int tb200_test_fq12_op(int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready()) return TB200_E_STATE;
  if (!a || !b || !out || n == 0 || op < 0) return fail(TB200_E_ARG, "bad arguments");
  CU(cudaSetDevice(primary().device));
  cudaStream_t st = primary().stream;
  return with_buffers(a, n * 576, b, n * 576, out, n * 576, nullptr, 0, [&](char* da, char* db, char* d1, char*) {
    if (op >= 20 && op < 100)
      LAUNCH(k_test_w12_op, (uint32_t)n, E.pairing_team == 32 ? 32 : 64, st, op, (const uint4*)da, (const uint4*)db, (uint4*)d1);
    else
      LAUNCH(k_test_fq12_op, cdiv(n, 32), 32, st, op, (const uint4*)da, (const uint4*)db, (uint32_t)n, (uint4*)d1);
    return 0;
  });
}

}  // extern "C"
