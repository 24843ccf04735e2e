// C++ host side above the C ABI (include/testudo_b200.h): mirrors the reference's Rust interfaces for the MSM path
// with the same names, argument meaning and error behaviour, so a maintainer can read it next to the Rust.
//   msm::{msm, msm_unchecked, msm_bigint}        ark-ec VariableBaseMSM
//   CommitterKey, Polynomial::{from_evaluations, commit, get_q, eval, open}   src/sqrt_pst.rs:14-230 (G1 work)
//   mipp::{multiexponentiation, compress, prove_g1}                          src/mipp.rs:31-153,354-394 (G1 work)
//   commitments::{MultiCommitGens, PedersenCommit, commit_inner}             src/commitments.rs, src/dense_mlpoly.rs:315-329
// The reference is Rust and no Rust toolchain exists in the build image, hence C++ (INTEGRATION.md has the Rust stub).
// Curve arithmetic never runs here: every group operation is a call into the CUDA library. The Fr scalar glue the
// reference computes on the CPU (chi products, q = Z*chi, challenge inversion) uses the same Montgomery code as the
// kernels through its host path (csrc/mont.cuh).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/testudo_b200.h"
#include "../csrc/mont.cuh"

namespace testudo_b200 {

struct Fr {  // ark-ff Fp256<MontBackend>: 4 x u64 Montgomery limbs
  uint64_t l[4];
  bool operator==(const Fr& o) const { return std::memcmp(l, o.l, 32) == 0; }
};
struct G1Affine {  // x[6] || y[6] Montgomery limbs; all-zero == identity
  uint64_t w[12];
  bool is_identity() const {
    for (auto v : w)
      if (v) return false;
    return true;
  }
  bool operator==(const G1Affine& o) const { return std::memcmp(w, o.w, 96) == 0; }
};
struct G2Affine {  // x.c0[6] || x.c1[6] || y.c0[6] || y.c1[6] Montgomery limbs (Fq2 = Fq[u]/(u^2+5)); all-zero == identity
  uint64_t w[24];
  bool operator==(const G2Affine& o) const { return std::memcmp(w, o.w, 192) == 0; }
};
struct Gt {  // ark Fq12 in memory: c0.c0.c0, c0.c0.c1, ..., c1.c2.c1, each 6 Montgomery limbs
  uint64_t w[72];
  bool operator==(const Gt& o) const { return std::memcmp(w, o.w, sizeof w) == 0; }
};
static_assert(sizeof(Fr) == 32 && sizeof(G1Affine) == 96 && sizeof(G2Affine) == 192 && sizeof(Gt) == 576, "ABI layouts");

struct EngineError : std::runtime_error {  // a CUDA failure has no error channel in commit/open: it unwinds
  int code;
  EngineError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc != 0) throw EngineError(rc, tb200_last_error());
}
inline void init(int device = -1) { check(tb200_init(device)); }

// ---- Fr glue (host path of the kernels' Montgomery code) ---------------------------------------------------------
namespace fr {
inline Fr mul(const Fr& a, const Fr& b) {
  Fr r;
  tb::mont_mul<tb::FrParams>((uint32_t*)r.l, (const uint32_t*)a.l, (const uint32_t*)b.l);
  return r;
}
inline Fr add(const Fr& a, const Fr& b) {
  Fr r;
  tb::mod_add<tb::FrParams>((uint32_t*)r.l, (const uint32_t*)a.l, (const uint32_t*)b.l);
  return r;
}
inline Fr sub(const Fr& a, const Fr& b) {
  Fr r;
  tb::mod_sub<tb::FrParams>((uint32_t*)r.l, (const uint32_t*)a.l, (const uint32_t*)b.l);
  return r;
}
inline Fr one() {
  Fr r;
  for (int i = 0; i < 8; i++) ((uint32_t*)r.l)[i] = tb::FrParams::one(i);
  return r;
}
inline Fr zero() { return Fr{{0, 0, 0, 0}}; }
inline Fr inverse(const Fr& a) {  // a^(r-2)
  static const uint64_t E[4] = {0x0a11800000000001ULL - 2, 0x59aa76fed0000001ULL, 0x60b44d1e5c37b001ULL,
                                0x12ab655e9a2ca556ULL};
  Fr acc = one(), base = a;
  for (int i = 0; i < 256; i++) {
    if ((E[i / 64] >> (i % 64)) & 1) acc = mul(acc, base);
    base = mul(base, base);
  }
  return acc;
}
}  // namespace fr

// ---- ark-ec VariableBaseMSM ------------------------------------------------------------------------------------------
namespace msm {
// msm_bigint(bases, bigints): canonical scalars; truncates to min(len)
inline G1Affine msm_bigint(const std::vector<G1Affine>& bases, const std::vector<Fr>& bigints) {
  size_t n = std::min(bases.size(), bigints.size());
  G1Affine out;
  check(tb200_msm_g1((const uint64_t*)bases.data(), (const uint64_t*)bigints.data(), n, 0, out.w));
  return out;
}
// msm_unchecked(bases, scalars): Montgomery-form Fr; silently truncates to min(len) like arkworks
inline G1Affine msm_unchecked(const G1Affine* bases, const Fr* scalars, size_t n) {
  G1Affine out;
  check(tb200_msm_g1((const uint64_t*)bases, (const uint64_t*)scalars, n, TB200_SCALARS_MONT, out.w));
  return out;
}
inline G1Affine msm_unchecked(const std::vector<G1Affine>& bases, const std::vector<Fr>& scalars) {
  return msm_unchecked(bases.data(), scalars.data(), std::min(bases.size(), scalars.size()));
}
// a ragged batch of independent small MSMs (rows of 0..1024 points, Montgomery-form scalars) in one launch
inline std::vector<G1Affine> msm_rows(const std::vector<G1Affine>& bases, const std::vector<Fr>& scalars,
                                      const std::vector<size_t>& row_len) {
  size_t total = 0;
  for (size_t l : row_len) total += l;
  if (bases.size() != total || scalars.size() != total) throw std::invalid_argument("bases / scalars must hold sum(row_len) entries");
  std::vector<G1Affine> out(row_len.size());
  check(tb200_msm_g1_rows((const uint64_t*)bases.data(), (const uint64_t*)scalars.data(), row_len.data(), row_len.size(),
                          TB200_SCALARS_MONT, (uint64_t*)out.data()));
  return out;
}
// msm(bases, scalars) -> Result<G1, usize>: .second == true on Ok; on Err .first is unspecified and err_len = min(len)
struct MsmResult {
  bool ok;
  G1Affine value;
  size_t err_len;
};
inline MsmResult msm(const std::vector<G1Affine>& bases, const std::vector<Fr>& scalars) {
  if (bases.size() != scalars.size()) return {false, {}, std::min(bases.size(), scalars.size())};
  return {true, msm_unchecked(bases, scalars), 0};
}
}  // namespace msm

// ---- sqrt_pst -----------------------------------------------------------------------------------------------------------
class CommitterKey {  // G1 side of ark-poly-commit CommitterKey: powers_of_g[0], resident on the GPU with window tables
 public:
  explicit CommitterKey(const std::vector<G1Affine>& powers_of_g0, int window_bits = 0) : n_(powers_of_g0.size()) {
    check(tb200_srs_load((const uint64_t*)powers_of_g0.data(), n_, window_bits, &h_));
  }
  ~CommitterKey() {
    if (h_) tb200_srs_free(h_);
  }
  CommitterKey(const CommitterKey&) = delete;
  CommitterKey& operator=(const CommitterKey&) = delete;
  tb200_srs_t handle() const { return h_; }
  size_t size() const { return n_; }

 private:
  tb200_srs_t h_ = nullptr;
  size_t n_;
};

// MultilinearPC::commit(ck, poly).g_product for one polynomial
inline G1Affine pc_commit(const CommitterKey& ck, const std::vector<Fr>& evals) {
  G1Affine out;
  check(tb200_msm_g1_batch(ck.handle(), (const uint64_t*)evals.data(), 1, evals.size(), (ptrdiff_t)evals.size(), 1,
                           TB200_SCALARS_MONT, out.w));
  return out;
}

namespace mipp {
struct InvalidIPVectorLength : std::runtime_error {  // Error::InvalidIPVectorLength, src/mipp.rs:400-421
  InvalidIPVectorLength() : std::runtime_error("InvalidIPVectorLength") {}
};
// src/mipp.rs:385-394
inline G1Affine multiexponentiation(const std::vector<G1Affine>& left, const std::vector<Fr>& right) {
  if (left.size() != right.size()) throw InvalidIPVectorLength();
  return msm::msm_unchecked(left, right);
}
// src/mipp.rs:354-367 (G1): vec[i] += vec[i + split]^scaler; vec is truncated to split
inline void compress(std::vector<G1Affine>& vec, size_t split, const Fr& scaler) {
  if (vec.size() < 2 * split) throw std::invalid_argument("compress: split too large");
  check(tb200_compress_g1((uint64_t*)vec.data(), split, scaler.l, TB200_SCALARS_MONT));
  vec.resize(split);
}
// src/mipp.rs:370-383
inline void compress_field(std::vector<Fr>& vec, size_t split, const Fr& scaler) {
  for (size_t i = 0; i < split; i++) vec[i] = fr::add(vec[i], fr::mul(vec[split + i], scaler));
  vec.resize(split);
}
struct MippProofG1 {  // G1 fields of MippProof<E> (src/mipp.rs:22-28)
  std::vector<std::pair<G1Affine, G1Affine>> comms_u;
  G1Affine final_a;
  Fr final_y;
  std::vector<Fr> xs, xs_inv;
};
// challenge(label, points appended by the reference) -> c_inv (the Poseidon transcript itself is out of scope)
using Challenge = std::function<Fr(const char* label, const std::vector<G1Affine>& appended)>;
// G1 part of MippProof::prove (src/mipp.rs:31-153); vectors stay on the GPU across rounds
inline MippProofG1 prove_g1(const Challenge& challenge, const std::vector<G1Affine>& a, const std::vector<Fr>& y,
                            const G1Affine& U) {
  if (a.size() != y.size()) throw InvalidIPVectorLength();
  MippProofG1 out;
  challenge("U", {U});  // src/mipp.rs:56
  tb200_mipp_t h = nullptr;
  check(tb200_mipp_g1_begin((const uint64_t*)a.data(), (const uint64_t*)y.data(), a.size(), TB200_SCALARS_MONT, &h));
  try {
    while (tb200_mipp_g1_len(h) > 1) {  // src/mipp.rs:58
      G1Affine ul, ur;
      check(tb200_mipp_g1_cross(h, ul.w, ur.w));                // :77-85
      Fr c_inv = challenge("challenge_i", {ul, ur});             // :97-101
      Fr c = fr::inverse(c_inv);                                 // :106
      check(tb200_mipp_g1_fold(h, c.l, c_inv.l));                // :110-112
      out.comms_u.push_back({ul, ur});                           // :117
      out.xs.push_back(c);
      out.xs_inv.push_back(c_inv);
    }
    check(tb200_mipp_g1_read(h, out.final_a.w, out.final_y.l));  // :122
  } catch (...) {
    tb200_mipp_g1_end(h);
    throw;
  }
  check(tb200_mipp_g1_end(h));
  return out;
}
}  // namespace mipp

class Polynomial {  // src/sqrt_pst.rs:14-20
 public:
  // src/sqrt_pst.rs:32-75. Z is kept un-transposed: row i is the strided view Z[(j << m_col) | i]
  static Polynomial from_evaluations(std::vector<Fr> Z) {
    size_t n = Z.size();
    if (n == 0 || (n & (n - 1))) throw std::invalid_argument("evaluation list must be a power of two");
    size_t nv = 0;
    while ((size_t(1) << nv) < n) nv++;
    Polynomial p;
    p.Z_ = std::move(Z);
    p.m_ = nv / 2;
    p.odd_ = nv % 2;
    return p;
  }
  size_t m() const { return m_; }
  size_t odd() const { return odd_; }
  // src/sqrt_pst.rs:117-149 -> comm_list (g_products). The pairing product t is out of scope.
  std::vector<G1Affine> commit(const CommitterKey& ck) const {
    size_t rows = size_t(1) << m_, cols = size_t(1) << (m_ + odd_);
    if (cols != ck.size()) throw std::invalid_argument("ck.powers_of_g[0] must have 2^m_row points");
    std::vector<G1Affine> out(rows);
    check(tb200_msm_g1_batch(ck.handle(), (const uint64_t*)Z_.data(), rows, cols, 1, (ptrdiff_t)rows, TB200_SCALARS_MONT,
                             (uint64_t*)out.data()));
    return out;
  }
  // src/sqrt_pst.rs:152-166 (bits of i MSB first)
  static Fr get_chi_i(const std::vector<Fr>& b, size_t i) {
    size_t m = b.size();
    Fr prod = fr::one();
    for (size_t j = 0; j < m; j++) {
      if ((i >> (m - j - 1)) & 1) prod = fr::mul(prod, b[j]);
      else prod = fr::mul(prod, fr::sub(fr::one(), b[j]));
    }
    return prod;
  }
  // src/sqrt_pst.rs:81-101 (the reference's CPU code, unchanged)
  void get_q(const std::vector<Fr>& point) {
    std::vector<Fr> b(point.begin() + m_ + odd_, point.end());
    size_t pow_m = size_t(1) << m_;
    chis_b_.resize(pow_m);
    for (size_t i = 0; i < pow_m; i++) chis_b_[i] = get_chi_i(b, i);
    q_.assign(pow_m << odd_, fr::zero());
    for (size_t j = 0; j < q_.size(); j++)
      for (size_t i = 0; i < pow_m; i++) q_[j] = fr::add(q_[j], fr::mul(Z_[(j << m_) | i], chis_b_[i]));
  }
  // src/sqrt_pst.rs:105-115
  Fr eval(const std::vector<Fr>& point) {
    std::vector<Fr> a(point.begin(), point.begin() + point.size() / 2 + odd_);
    if (q_.empty()) get_q(point);
    Fr acc = fr::zero();
    for (size_t j = 0; j < q_.size(); j++) acc = fr::add(acc, fr::mul(q_[j], get_chi_i(a, j)));
    return acc;
  }
  struct OpenG1 {
    G1Affine u, comm_q;
    mipp::MippProofG1 mipp;
  };
  // G1 work of src/sqrt_pst.rs:168-230
  OpenG1 open(const mipp::Challenge& challenge, const std::vector<G1Affine>& comm_list, const CommitterKey& ck,
              const std::vector<Fr>& point) {
    if (q_.empty()) get_q(point);
    if (chis_b_.size() != comm_list.size()) throw std::logic_error("chis.len() == comm_list.len()");  // :194
    OpenG1 o;
    o.u = msm::msm_unchecked(comm_list, chis_b_);  // :198
    o.comm_q = pc_commit(ck, q_);                   // :205
    if (!(o.u == o.comm_q)) throw std::logic_error("debug_assert!(c_u == comm.g_product) failed");  // :206
    o.mipp = mipp::prove_g1(challenge, comm_list, chis_b_, o.u);  // :212-213
    return o;
  }
  const std::vector<Fr>& q() const { return q_; }
  const std::vector<Fr>& chis_b() const { return chis_b_; }

 private:
  std::vector<Fr> Z_, q_, chis_b_;
  size_t m_ = 0, odd_ = 0;
};

// ---- commitments ---------------------------------------------------------------------------------------------------------
namespace commitments {
struct MultiCommitGens {  // src/commitments.rs:10-15 (generator derivation :17-39 is out of scope)
  std::vector<G1Affine> G;
  G1Affine h;
  size_t n() const { return G.size(); }
};
struct PedersenCommit {
  // src/commitments.rs:70-77
  static G1Affine commit_scalar(const Fr& scalar, const Fr& blind, const MultiCommitGens& gens_n) {
    if (gens_n.n() != 1) throw std::invalid_argument("assert_eq!(gens_n.n, 1)");
    std::vector<G1Affine> b{gens_n.G[0], gens_n.h};
    std::vector<Fr> s{scalar, blind};
    return msm::msm_unchecked(b, s);
  }
  // src/commitments.rs:79-86: msm_unchecked(G, scalars) + h * blind as one MSM over G || h
  static G1Affine commit_slice(const std::vector<Fr>& scalars, const Fr& blind, const MultiCommitGens& gens_n) {
    if (scalars.size() != gens_n.n()) throw std::invalid_argument("assert_eq!(scalars.len(), gens_n.n)");
    std::vector<G1Affine> b = gens_n.G;
    b.push_back(gens_n.h);
    std::vector<Fr> s = scalars;
    s.push_back(blind);
    return msm::msm_unchecked(b, s);
  }
};
// DensePolynomial::commit_inner (src/dense_mlpoly.rs:315-329) with all-zero blinds (every `commit(gens, false)` site)
inline std::vector<G1Affine> commit_inner(const std::vector<Fr>& Z, size_t L_size, const CommitterKey& gens_srs) {
  size_t R = Z.size() / L_size;
  if (L_size * R != Z.size() || R != gens_srs.size()) throw std::invalid_argument("L_size * R_size == Z.len()");
  std::vector<G1Affine> out(L_size);
  check(tb200_msm_g1_batch(gens_srs.handle(), (const uint64_t*)Z.data(), L_size, R, (ptrdiff_t)R, 1, TB200_SCALARS_MONT,
                           (uint64_t*)out.data()));
  return out;
}
}  // namespace commitments

// ---- ark-ec VariableBaseMSM for G2Projective (commit_g2, src/mipp.rs:133) ----------------------------------------------
namespace msm_g2 {
inline G2Affine msm_unchecked(const std::vector<G2Affine>& bases, const std::vector<Fr>& scalars) {
  G2Affine out;
  check(tb200_msm_g2((const uint64_t*)bases.data(), (const uint64_t*)scalars.data(),
                     std::min(bases.size(), scalars.size()), TB200_SCALARS_MONT, out.w));
  return out;
}
// MIPP `compress` on the G2 key (src/mipp.rs:114, 354-367): vec[i] += scaler * vec[split + i], truncated to split
inline void compress(std::vector<G2Affine>& vec, size_t split, const Fr& scaler) {
  if (vec.size() < 2 * split) throw std::invalid_argument("compress: vector shorter than 2 * split");
  check(tb200_compress_g2((uint64_t*)vec.data(), split, scaler.l, TB200_SCALARS_MONT));
  vec.resize(split);
}
}  // namespace msm_g2

// ---- ark-poly-commit multilinear_pc::MultilinearPC::{open, open_g1} (src/sqrt_pst.rs:225, src/mipp.rs:144) -----------
namespace multilinear_pc {
// level_bases[i] = ck.powers_of_h[off + i] (2^(nv - i) points); evals = poly.to_evaluations(); point has nv entries
inline std::vector<G2Affine> open(const std::vector<std::vector<G2Affine>>& level_bases, const std::vector<Fr>& evals,
                                  const std::vector<Fr>& point) {
  const size_t nv = point.size();
  if (evals.size() != (size_t(1) << nv) || level_bases.size() < nv) throw std::invalid_argument("open: sizes");
  std::vector<const uint64_t*> ptrs(nv);
  for (size_t i = 0; i < nv; i++) {
    if (level_bases[i].size() != (size_t(1) << (nv - i))) throw std::invalid_argument("open: CRS level size");
    ptrs[i] = (const uint64_t*)level_bases[i].data();
  }
  std::vector<G2Affine> proofs(nv);
  check(tb200_pst_open_g2((const uint64_t*)evals.data(), nv, (const uint64_t*)point.data(), ptrs.data(),
                          TB200_SCALARS_MONT, (uint64_t*)proofs.data()));
  return proofs;
}
inline std::vector<G1Affine> open_g1(const std::vector<std::vector<G1Affine>>& level_bases, const std::vector<Fr>& evals,
                                     const std::vector<Fr>& point) {
  const size_t nv = point.size();
  if (evals.size() != (size_t(1) << nv) || level_bases.size() < nv) throw std::invalid_argument("open_g1: sizes");
  std::vector<const uint64_t*> ptrs(nv);
  for (size_t i = 0; i < nv; i++) {
    if (level_bases[i].size() != (size_t(1) << (nv - i))) throw std::invalid_argument("open_g1: CRS level size");
    ptrs[i] = (const uint64_t*)level_bases[i].data();
  }
  std::vector<G1Affine> proofs(nv);
  check(tb200_pst_open_g1((const uint64_t*)evals.data(), nv, (const uint64_t*)point.data(), ptrs.data(),
                          TB200_SCALARS_MONT, (uint64_t*)proofs.data()));
  return proofs;
}
}  // namespace multilinear_pc

// ---- ark-ec Pairing for Bls12_377: E::multi_pairing / pairings_product (src/sqrt_pst.rs:131-144, src/mipp.rs:396-398) ---
namespace pairing {
// `E::multi_pairing(a, b).0`: zips (the shorter side bounds the product), pairs with an identity contribute 1
inline Gt multi_pairing(const std::vector<G1Affine>& a, const std::vector<G2Affine>& b) {
  Gt out;
  check(tb200_multi_pairing((const uint64_t*)a.data(), (const uint64_t*)b.data(), std::min(a.size(), b.size()), out.w));
  return out;
}
inline Gt pairings_product(const std::vector<G1Affine>& gs, const std::vector<G2Affine>& hs) {  // src/mipp.rs:396-398
  return multi_pairing(gs, hs);
}
inline Gt pairing(const G1Affine& p, const G2Affine& q) { return multi_pairing({p}, {q}); }
// `TargetField::pow(c.into_bigint())` on Montgomery-form exponents (the verifier's tx.pow(c), src/mipp.rs:258-261)
inline Gt pow(const Gt& base, const Fr& exp) {
  Gt out;
  check(tb200_gt_pow(base.w, exp.l, 1, TB200_SCALARS_MONT, out.w));
  return out;
}
// prod_i bases[i].pow(exps[i]): the TC half of the verifier's fold / reduce over MippTU (src/mipp.rs:240-271) in one call
inline Gt multi_pow(const std::vector<Gt>& bases, const std::vector<Fr>& exps) {
  if (bases.size() != exps.size()) throw std::invalid_argument("bases and exponents differ in length");
  Gt out;
  check(tb200_gt_multi_pow((const uint64_t*)bases.data(), (const uint64_t*)exps.data(), bases.size(), TB200_SCALARS_MONT, out.w));
  return out;
}
// several independent pairing products in ONE pass (the verifier's five: src/mipp.rs:307,311, src/sqrt_pst.rs:261);
// shorter products are padded with identity pairs, which contribute 1 as in ark
inline std::vector<Gt> multi_pairing_batch(const std::vector<std::pair<std::vector<G1Affine>, std::vector<G2Affine>>>& products) {
  size_t width = 1;
  for (const auto& p : products) width = std::max(width, std::min(p.first.size(), p.second.size()));
  std::vector<G1Affine> a(products.size() * width, G1Affine{});
  std::vector<G2Affine> b(products.size() * width, G2Affine{});
  for (size_t i = 0; i < products.size(); i++) {
    const size_t n = std::min(products[i].first.size(), products[i].second.size());
    std::copy_n(products[i].first.begin(), n, a.begin() + i * width);
    std::copy_n(products[i].second.begin(), n, b.begin() + i * width);
  }
  std::vector<Gt> out(products.size());
  check(tb200_multi_pairing_batch((const uint64_t*)a.data(), (const uint64_t*)b.data(), products.size(), width,
                                  (uint64_t*)out.data()));
  return out;
}
// the IPP commitment of Polynomial::commit: t = multi_pairing(comm_list, ck.powers_of_h[odd])  (src/sqrt_pst.rs:128-143)
inline Gt ipp_commitment(const std::vector<G1Affine>& comm_list, const std::vector<G2Affine>& h_vec) {
  if (comm_list.size() != h_vec.size()) throw std::invalid_argument("assert!(comm_list.len() == h_vec.len())");
  return multi_pairing(comm_list, h_vec);
}
}  // namespace pairing

}  // namespace testudo_b200
