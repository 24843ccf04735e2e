//! Prints tests/golden/ark_kat.json: values computed by ARKWORKS and by the reference crate itself on fixed inputs.
//! tests/test_ark_kat.py loads the file (skipped while it is absent) and requires the oracle AND the GPU engine to
//! reproduce every value bit for bit -- the step that turns "parity unpinned"  into "pinned".
//! Needs cargo + the reference checkout; NOT runnable in the build image. Usage: cargo run --release > ark_kat.json
//!
//! Encodings: field elements and points as lowercase hex of their ark `CanonicalSerialize` UNCOMPRESSED bytes.
use ark_bls12_377::{Bls12_377, Fr, G1Affine, G1Projective, G2Affine};
use ark_ec::{pairing::Pairing, scalar_mul::variable_base::VariableBaseMSM, AffineRepr, CurveGroup};
use ark_ff::{PrimeField, UniformRand};
use ark_poly_commit::multilinear_pc::MultilinearPC;
use ark_serialize::{CanonicalSerialize, Compress};
use libtestudo::{parameters::get_bls12377_fq_params, poseidon_transcript::PoseidonTranscript, sqrt_pst::Polynomial,
                 transcript::Transcript};
use serde_json::json;

fn hex<T: CanonicalSerialize>(v: &T) -> String {
    let mut b = Vec::new();
    v.serialize_with_mode(&mut b, Compress::No).unwrap();
    b.iter().map(|x| format!("{:02x}", x)).collect()
}

fn main() {
    // deterministic inputs that need no shared RNG: bases (i + 1) G, scalars from a counter hashed by squaring in Fr
    let g = G1Affine::generator();
    let g2 = G2Affine::generator();
    let mut kats = vec![];
    for &n in &[1usize, 2, 3, 31, 32, 33, 256, 1024] {
        let bases: Vec<G1Affine> = (0..n).map(|i| (g * Fr::from((i + 1) as u64)).into_affine()).collect();
        let mut s = Fr::from(0x9E3779B97F4A7C15u64);
        let scalars: Vec<Fr> = (0..n).map(|_| { s = s * s + Fr::from(7u64); s }).collect();
        let bigints: Vec<_> = scalars.iter().map(|x| x.into_bigint()).collect();
        let r = G1Projective::msm_bigint(&bases, &bigints).into_affine();
        kats.push(json!({"kind": "msm_g1", "n": n, "result": hex(&r)}));
    }
    // pairings: e(G1, G2), e(a G1, b G2), a product with an identity
    let (a, b) = (Fr::from(0x1234567u64), Fr::from(0x89ABCDEFu64));
    let e = Bls12_377::pairing(g, g2);
    let eab = Bls12_377::pairing((g * a).into_affine(), (g2 * b).into_affine());
    let prod = Bls12_377::multi_pairing([g, (g * a).into_affine(), G1Affine::identity()], [g2, (g2 * b).into_affine(), g2]);
    kats.push(json!({"kind": "pairing", "e_g1_g2": hex(&e.0), "e_aG1_bG2": hex(&eab.0), "a": "0x1234567", "b": "0x89abcdef",
                     "product": hex(&prod.0)}));
    // Poseidon transcript: the appends of one MIPP round, then three challenges
    let mut tr = PoseidonTranscript::new(&get_bls12377_fq_params());
    tr.append(b"U", &g);
    tr.append(b"comm_u_l", &(g * a).into_affine());
    tr.append(b"comm_t_l", &e.0);
    let c1: Fr = tr.challenge_scalar(b"challenge_i");
    let c2: Fr = tr.challenge_scalar(b"challenge_i");
    tr.append(b"comm_u_r", &G1Affine::identity());
    let c3: Fr = tr.challenge_scalar(b"random_point");
    kats.push(json!({"kind": "poseidon_fq", "appends": ["G1", "a*G1", "e(G1,G2)", "squeeze", "squeeze", "identity", "squeeze"],
                     "challenges": [hex(&c1), hex(&c2), hex(&c3)]}));
    // sqrt_pst commit / open on a CRS with a fixed test RNG (ark_std::test_rng is seeded: the CRS is reproducible from
    // its serialisation below), num_vars = 4 and 5
    for &nv in &[4usize, 5] {
        let mut rng = ark_std::test_rng();
        let z: Vec<Fr> = (0..1usize << nv).map(|_| Fr::rand(&mut rng)).collect();
        let r: Vec<Fr> = (0..nv).map(|_| Fr::rand(&mut rng)).collect();
        let m = (nv + 1) / 2;
        let gens = MultilinearPC::<Bls12_377>::setup(m, &mut rng);
        let (ck, _vk) = MultilinearPC::<Bls12_377>::trim(&gens, m);
        let mut pl = Polynomial::from_evaluations(&z);
        let v = pl.eval(&r);
        let (comm_list, t) = pl.commit(&ck);
        let mut tr = PoseidonTranscript::new(&get_bls12377_fq_params());
        let (u, pst_proof, mipp_proof) = pl.open(&mut tr, comm_list.clone(), &ck, &r, &t);
        kats.push(json!({"kind": "sqrt_pst", "num_vars": nv, "z": z.iter().map(hex).collect::<Vec<_>>(),
                         "r": r.iter().map(hex).collect::<Vec<_>>(), "eval": hex(&v),
                         "powers_of_g": ck.powers_of_g.iter().map(|l| l.iter().map(hex).collect::<Vec<_>>()).collect::<Vec<_>>(),
                         "powers_of_h": ck.powers_of_h.iter().map(|l| l.iter().map(hex).collect::<Vec<_>>()).collect::<Vec<_>>(),
                         "comm_list": comm_list.iter().map(|c| hex(&c.g_product)).collect::<Vec<_>>(), "t": hex(&t.0),
                         "u": hex(&u.g_product), "pst_proof": pst_proof.proofs.iter().map(hex).collect::<Vec<_>>(),
                         "mipp": {"comms_u": mipp_proof.comms_u.iter().map(|(l, r)| [hex(l), hex(r)]).collect::<Vec<_>>(),
                                  "comms_t": mipp_proof.comms_t.iter().map(|(l, r)| [hex(&l.0), hex(&r.0)]).collect::<Vec<_>>(),
                                  "final_a": hex(&mipp_proof.final_a), "final_h": hex(&mipp_proof.final_h),
                                  "pst_proof_h": mipp_proof.pst_proof_h.proofs.iter().map(hex).collect::<Vec<_>>()}}));
    }
    println!("{}", serde_json::to_string_pretty(&json!({"generator": "ffi/kat (arkworks 0.4 + rosariocannavo/testudo)", "kats": kats})).unwrap());
}
