// Test driver for the C++ host mirror (testudo_b200/host/testudo_b200.hpp): reads a little-endian blob written by
// tests/test_gpu_host_cpp.py, runs commit / open / MIPP / Pedersen through the mirror, writes the results back.
// blob in : u64 nv | Z[2^nv] Fr | srs[2^m_row] G1 | point[nv] Fr | h G1 | blind Fr
// blob out: comm_list[2^m_col] | u | comm_q | eval Fr | n_rounds u64 | (ul, ur)[rounds] | final_a | final_y | slice | err_len u64
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../../testudo_b200/host/testudo_b200.hpp"
using namespace testudo_b200;

template <class T>
static std::vector<T> rd(std::ifstream& f, size_t n) {
  std::vector<T> v(n);
  f.read((char*)v.data(), n * sizeof(T));
  return v;
}
template <class T>
static void wr(std::ofstream& f, const T* p, size_t n) { f.write((const char*)p, n * sizeof(T)); }

// deterministic stand-in for the Poseidon transcript: FNV-style mix of everything appended (same in the python test)
struct FakeTranscript {
  uint64_t st = 0xcbf29ce484222325ULL;
  void absorb(const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { st ^= b[i]; st *= 0x100000001b3ULL; }
  }
  Fr challenge(const char* label, const std::vector<G1Affine>& pts) {
    absorb(label, strlen(label));
    for (auto& p : pts) absorb(p.w, 96);
    // small canonical value lifted to Montgomery form by multiplying with R^2
    Fr v{{st | 1, st >> 7, 0, 0}};
    Fr r2;
    for (int i = 0; i < 8; i++) ((uint32_t*)r2.l)[i] = tb::FrParams::r2(i);
    return fr::mul(v, r2);
  }
};

// mode "g2": blob in : u64 nv | evals[2^nv] Fr | point[nv] Fr | G2 levels (2^nv, 2^(nv-1), .., 2 points) | G1 levels
//            blob out: G2 proofs[nv] | G1 proofs[nv] | msm_g2(level 0, evals) | compress(level 0, split, point[0])[split]
static int run_g2(const char* fin, const char* fout) {
  std::ifstream in(fin, std::ios::binary);
  uint64_t nv;
  in.read((char*)&nv, 8);
  auto evals = rd<Fr>(in, size_t(1) << nv);
  auto point = rd<Fr>(in, nv);
  std::vector<std::vector<G2Affine>> h_levels;
  std::vector<std::vector<G1Affine>> g_levels;
  for (size_t i = 0; i < nv; i++) h_levels.push_back(rd<G2Affine>(in, size_t(1) << (nv - i)));
  for (size_t i = 0; i < nv; i++) g_levels.push_back(rd<G1Affine>(in, size_t(1) << (nv - i)));
  std::ofstream out(fout, std::ios::binary);
  auto p2 = multilinear_pc::open(h_levels, evals, point);
  wr(out, p2.data(), p2.size());
  auto p1 = multilinear_pc::open_g1(g_levels, evals, point);
  wr(out, p1.data(), p1.size());
  G2Affine c = msm_g2::msm_unchecked(h_levels[0], evals);
  wr(out, &c, 1);
  auto v = h_levels[0];
  msm_g2::compress(v, v.size() / 2, point[0]);
  wr(out, v.data(), v.size());
  return 0;
}

// in: n, n G1 points, n G2 points, one Fr exponent -> out: multi_pairing, pairing of the first pair, its power, and
// multi_pairing with the G2 side one element shorter (zip semantics), the length-rule flag, (t * first)^e by multi_pow,
// and {multi_pairing, first pairing} again through multi_pairing_batch
static int run_pairing(const char* fin, const char* fout) {
  std::ifstream in(fin, std::ios::binary);
  uint64_t n;
  in.read((char*)&n, 8);
  auto a = rd<G1Affine>(in, n);
  auto b = rd<G2Affine>(in, n);
  auto e = rd<Fr>(in, 1);
  std::ofstream out(fout, std::ios::binary);
  Gt t = pairing::ipp_commitment(a, b);
  wr(out, &t, 1);
  Gt one = pairing::pairing(a[0], b[0]);
  wr(out, &one, 1);
  Gt pw = pairing::pow(one, e[0]);
  wr(out, &pw, 1);
  std::vector<G2Affine> shorter(b.begin(), b.end() - 1);
  Gt z = pairing::pairings_product(a, shorter);
  wr(out, &z, 1);
  uint64_t threw = 0;
  try { pairing::ipp_commitment(a, shorter); } catch (const std::invalid_argument&) { threw = 1; }
  wr(out, &threw, 1);
  // the verifier's primitives: t^e * one^e in one call, and two products (all pairs / the first pair) in one pass
  Gt mp = pairing::multi_pow({t, one}, {e[0], e[0]});
  wr(out, &mp, 1);
  auto batch = pairing::multi_pairing_batch({{a, b}, {{a[0]}, {b[0]}}});
  wr(out, batch.data(), batch.size());
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  try {
    init(-1);
    if (argc > 3 && std::string(argv[3]) == "g2") return run_g2(argv[1], argv[2]);
    if (argc > 3 && std::string(argv[3]) == "pairing") return run_pairing(argv[1], argv[2]);
    std::ifstream in(argv[1], std::ios::binary);
    uint64_t nv;
    in.read((char*)&nv, 8);
    size_t m_col = nv / 2, m_row = nv - m_col;
    auto Z = rd<Fr>(in, size_t(1) << nv);
    auto srs = rd<G1Affine>(in, size_t(1) << m_row);
    auto point = rd<Fr>(in, nv);
    auto hpt = rd<G1Affine>(in, 1);
    auto blind = rd<Fr>(in, 1);
    std::ofstream out(argv[2], std::ios::binary);

    CommitterKey ck(srs);
    Polynomial poly = Polynomial::from_evaluations(Z);
    auto comm_list = poly.commit(ck);
    wr(out, comm_list.data(), comm_list.size());
    FakeTranscript tr;
    auto opened = poly.open([&](const char* l, const std::vector<G1Affine>& p) { return tr.challenge(l, p); }, comm_list,
                            ck, point);
    wr(out, &opened.u, 1);
    wr(out, &opened.comm_q, 1);
    Fr ev = poly.eval(point);
    wr(out, &ev, 1);
    uint64_t rounds = opened.mipp.comms_u.size();
    wr(out, &rounds, 1);
    for (auto& pr : opened.mipp.comms_u) { wr(out, &pr.first, 1); wr(out, &pr.second, 1); }
    wr(out, &opened.mipp.final_a, 1);
    wr(out, &opened.mipp.final_y, 1);
    // Pedersen commit_slice over gens = (srs, h) with the first row of Z and a blind
    commitments::MultiCommitGens gens{srs, hpt[0]};
    std::vector<Fr> row0(Z.begin(), Z.begin() + srs.size());
    G1Affine sl = commitments::PedersenCommit::commit_slice(row0, blind[0], gens);
    wr(out, &sl, 1);
    // VariableBaseMSM::msm length rule
    std::vector<Fr> shorter(Z.begin(), Z.begin() + srs.size() - 1);
    auto r = msm::msm(srs, shorter);
    uint64_t err = r.ok ? ~0ull : r.err_len;
    wr(out, &err, 1);
    // multiexponentiation error behaviour
    uint64_t threw = 0;
    try { mipp::multiexponentiation(srs, shorter); } catch (const mipp::InvalidIPVectorLength&) { threw = 1; }
    wr(out, &threw, 1);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "host_driver: " << e.what() << std::endl;
    return 1;
  }
}
