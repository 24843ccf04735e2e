//! Safe wrapper over `testudo-b200-sys` for the call sites of rosariocannavo/testudo.
//! NOT COMPILED in the build image (no Rust toolchain). Every function names the reference line it replaces.
//!
//! Layout contract (include/testudo_b200.h): a G1 affine point crosses the boundary as x[6] || y[6] Montgomery limbs
//! (all-zero = identity), a G2 point as x.c0 || x.c1 || y.c0 || y.c1, an `Fr` as its 4 Montgomery limbs
//! (`TB200_SCALARS_MONT`), an `Fq12` as its twelve `Fq` in declaration order. `ark_ec::short_weierstrass::Affine` is a
//! Rust-layout struct with an `infinity: bool`, so points are REPACKED; `Fr` slices are passed as they lie in memory.
use ark_bls12_377::{Bls12_377, Fq, Fq12, Fq2, Fr, G1Affine, G1Projective, G2Affine};
use ark_ec::{pairing::PairingOutput, AffineRepr};
use ark_ff::{BigInt, Fp};
use std::os::raw::c_void;
use testudo_b200_sys as sys;

fn check(rc: i32) {
    // `commit` / `open` have no error channel in the reference: a CUDA failure panics
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::tb200_last_error()) }.to_string_lossy().into_owned();
        panic!("testudo_b200 error {rc}: {msg}");
    }
}

/// One-time initialisation: every listed GPU gets a context, `devices[0]` is the primary (tb200_init_devices).
pub fn init(devices: &[i32]) {
    check(unsafe { sys::tb200_init_devices(devices.as_ptr(), devices.len() as i32) });
}

// ---- limb repacking ----------------------------------------------------------------------------------------------------
fn fq_limbs(x: &Fq) -> [u64; 6] { x.0 .0 }                                    // Fp(BigInt([u64; 6]), _): Montgomery form
fn fq_from(w: &[u64]) -> Fq { Fp::new_unchecked(BigInt::new(w.try_into().unwrap())) }

pub fn pack_g1(p: &[G1Affine]) -> Vec<u64> {
    let mut v = vec![0u64; 12 * p.len()];
    for (i, a) in p.iter().enumerate() {
        if let Some((x, y)) = a.xy() {
            v[12 * i..12 * i + 6].copy_from_slice(&fq_limbs(x));
            v[12 * i + 6..12 * i + 12].copy_from_slice(&fq_limbs(y));
        }
    }
    v
}
pub fn unpack_g1(w: &[u64]) -> G1Affine {
    if w.iter().all(|&l| l == 0) { return G1Affine::identity(); }
    G1Affine::new_unchecked(fq_from(&w[0..6]), fq_from(&w[6..12]))
}
pub fn pack_g2(p: &[G2Affine]) -> Vec<u64> {
    let mut v = vec![0u64; 24 * p.len()];
    for (i, a) in p.iter().enumerate() {
        if let Some((x, y)) = a.xy() {
            for (k, c) in [&x.c0, &x.c1, &y.c0, &y.c1].iter().enumerate() {
                v[24 * i + 6 * k..24 * i + 6 * k + 6].copy_from_slice(&fq_limbs(c));
            }
        }
    }
    v
}
pub fn unpack_g2(w: &[u64]) -> G2Affine {
    if w.iter().all(|&l| l == 0) { return G2Affine::identity(); }
    G2Affine::new_unchecked(Fq2::new(fq_from(&w[0..6]), fq_from(&w[6..12])), Fq2::new(fq_from(&w[12..18]), fq_from(&w[18..24])))
}
pub fn unpack_gt(w: &[u64; 72]) -> Fq12 {
    let c: Vec<Fq> = w.chunks(6).map(fq_from).collect();
    let f2 = |i: usize| Fq2::new(c[2 * i], c[2 * i + 1]);
    Fq12::new(ark_bls12_377::Fq6::new(f2(0), f2(1), f2(2)), ark_bls12_377::Fq6::new(f2(3), f2(4), f2(5)))
}
fn fr_limbs(s: &[Fr]) -> *const u64 { s.as_ptr() as *const u64 }              // Fp256<MontBackend>: [u64; 4] each

// ---- ark-ec VariableBaseMSM (src/sqrt_pst.rs:198, src/mipp.rs:385-394, src/commitments.rs:70-86) --------------------------
/// `<G1Projective as VariableBaseMSM>::msm_unchecked`: truncates to min(len) like arkworks
pub fn msm_unchecked(bases: &[G1Affine], scalars: &[Fr]) -> G1Projective {
    let n = bases.len().min(scalars.len());
    let (b, mut out) = (pack_g1(&bases[..n]), [0u64; 12]);
    check(unsafe { sys::tb200_msm_g1(b.as_ptr(), fr_limbs(&scalars[..n]), n, sys::TB200_SCALARS_MONT, out.as_mut_ptr()) });
    unpack_g1(&out).into_group()
}
/// `VariableBaseMSM::msm`: Err(min_len) on a length mismatch (src/dense_mlpoly.rs:554, src/nizk/bullet.rs:237)
pub fn msm(bases: &[G1Affine], scalars: &[Fr]) -> Result<G1Projective, usize> {
    if bases.len() != scalars.len() { Err(bases.len().min(scalars.len())) } else { Ok(msm_unchecked(bases, scalars)) }
}
pub fn msm_g2_unchecked(bases: &[G2Affine], scalars: &[Fr]) -> G2Affine {
    let n = bases.len().min(scalars.len());
    let (b, mut out) = (pack_g2(&bases[..n]), [0u64; 24]);
    check(unsafe { sys::tb200_msm_g2(b.as_ptr(), fr_limbs(&scalars[..n]), n, sys::TB200_SCALARS_MONT, out.as_mut_ptr()) });
    unpack_g2(&out)
}

// ---- the SRS ck.powers_of_g[0] with its window tables, replicated on every GPU ------------------------------------------
pub struct Srs(*mut sys::tb200_srs);
unsafe impl Send for Srs {}
unsafe impl Sync for Srs {}
impl Srs {
    pub fn load(powers_of_g0: &[G1Affine]) -> Self {
        let (b, mut h) = (pack_g1(powers_of_g0), std::ptr::null_mut());
        check(unsafe { sys::tb200_srs_load(b.as_ptr(), powers_of_g0.len(), 0, &mut h) });
        Srs(h)
    }
}
impl Drop for Srs { fn drop(&mut self) { unsafe { sys::tb200_srs_free(self.0) }; } }

/// `Polynomial::commit` (src/sqrt_pst.rs:117-149) in one call: rows = `self.polys[i].Z`, h_vec = `ck.powers_of_h[odd]`.
/// Returns (comm_list g_products, t). With several GPUs the library shards by row range and all-gathers the partial
/// Miller products itself.
pub fn sqrt_pst_commit(srs: &Srs, rows: &[&[Fr]], h_vec: &[G2Affine]) -> (Vec<G1Affine>, PairingOutput<Bls12_377>) {
    let cols = rows[0].len();
    let ptrs: Vec<*const u64> = rows.iter().map(|r| fr_limbs(r)).collect();
    let h = pack_g2(h_vec);
    let mut out = vec![0u64; 12 * rows.len()];
    let mut t = [0u64; 72];
    check(unsafe { sys::tb200_sqrt_pst_commit(srs.0, ptrs.as_ptr(), rows.len(), cols, sys::TB200_SCALARS_MONT, h.as_ptr(),
                                              out.as_mut_ptr(), t.as_mut_ptr()) });
    (out.chunks(12).map(unpack_g1).collect(), PairingOutput(unpack_gt(&t)))
}
/// `MultilinearPC::commit(ck, &q).g_product` (src/sqrt_pst.rs:205): one row over the same SRS
pub fn pc_commit(srs: &Srs, evals: &[Fr]) -> G1Affine {
    let mut out = [0u64; 12];
    check(unsafe { sys::tb200_msm_g1_batch(srs.0, fr_limbs(evals), 1, evals.len(), evals.len() as isize, 1,
                                           sys::TB200_SCALARS_MONT, out.as_mut_ptr()) });
    unpack_g1(&out)
}

// ---- `E::multi_pairing` (src/sqrt_pst.rs:131-144, src/mipp.rs:396-398) ------------------------------------------------------
pub fn multi_pairing(a: &[G1Affine], b: &[G2Affine]) -> PairingOutput<Bls12_377> {
    let (pa, pb, mut t) = (pack_g1(a), pack_g2(b), [0u64; 72]);
    check(unsafe { sys::tb200_multi_pairing(pa.as_ptr(), pb.as_ptr(), a.len().min(b.len()), t.as_mut_ptr()) });
    PairingOutput(unpack_gt(&t))
}

// ---- the verifier side: `MippProof::verify` (src/mipp.rs:182-333), `MultilinearPC::check` (src/sqrt_pst.rs:261) ------------
fn pack_gt(v: &[Fq12]) -> Vec<u64> {
    let mut w = Vec::with_capacity(72 * v.len());
    for f in v {
        for c6 in [&f.c0, &f.c1] { for c2 in [&c6.c0, &c6.c1, &c6.c2] { w.extend_from_slice(&c2.c0.0 .0); w.extend_from_slice(&c2.c1.0 .0); } }
    }
    w
}
/// prod_i bases[i].pow(exps[i]): the TC half of the fold / reduce over `MippTU` (src/mipp.rs:240-271) in one call
pub fn gt_multi_pow(bases: &[Fq12], exps: &[Fr]) -> Fq12 {
    assert_eq!(bases.len(), exps.len());
    let (b, mut out) = (pack_gt(bases), [0u64; 72]);
    check(unsafe { sys::tb200_gt_multi_pow(b.as_ptr(), fr_limbs(exps), bases.len(), sys::TB200_SCALARS_MONT, out.as_mut_ptr()) });
    unpack_gt(&out)
}
/// `rows` independent MSMs of `per_row` (<= 8) points each in one launch: `commitment - g*value`, `g_mask[i] - g*point[i]`
pub fn msm_g1_each(bases: &[G1Affine], scalars: &[Fr], per_row: usize) -> Vec<G1Affine> {
    assert!(per_row > 0 && bases.len() == scalars.len() && bases.len() % per_row == 0);
    let rows = bases.len() / per_row;
    let (b, mut out) = (pack_g1(bases), vec![0u64; 12 * rows]);
    check(unsafe { sys::tb200_msm_g1_each(b.as_ptr(), fr_limbs(scalars), rows, per_row, sys::TB200_SCALARS_MONT, out.as_mut_ptr()) });
    out.chunks(12).map(unpack_g1).collect()
}
/// a ragged batch of independent small MSMs (rows of 0..1024 points) in one launch: `rows[i]` = (bases, scalars) of row i
pub fn msm_g1_rows(rows: &[(&[G1Affine], &[Fr])]) -> Vec<G1Affine> {
    let lens: Vec<usize> = rows.iter().map(|(b, s)| b.len().min(s.len())).collect();
    let (mut pb, mut ps) = (Vec::new(), Vec::new());
    for ((b, s), &n) in rows.iter().zip(&lens) {
        pb.extend(pack_g1(&b[..n]));
        for k in &s[..n] { ps.extend_from_slice(&k.0 .0); }
    }
    let mut out = vec![0u64; 12 * rows.len()];
    check(unsafe { sys::tb200_msm_g1_rows(pb.as_ptr(), ps.as_ptr(), lens.as_ptr(), rows.len(), sys::TB200_SCALARS_MONT, out.as_mut_ptr()) });
    out.chunks(12).map(unpack_g1).collect()
}
/// several `E::multi_pairing` in ONE pass of the pairing engine; shorter products are padded with identity pairs
pub fn multi_pairing_batch(products: &[(&[G1Affine], &[G2Affine])]) -> Vec<PairingOutput<Bls12_377>> {
    let width = products.iter().map(|(a, b)| a.len().min(b.len())).max().unwrap_or(0).max(1);
    let (mut pa, mut pb) = (vec![0u64; 12 * width * products.len()], vec![0u64; 24 * width * products.len()]);
    for (i, (a, b)) in products.iter().enumerate() {
        let n = a.len().min(b.len());
        pa[12 * width * i..12 * (width * i + n)].copy_from_slice(&pack_g1(&a[..n]));
        pb[24 * width * i..24 * (width * i + n)].copy_from_slice(&pack_g2(&b[..n]));
    }
    let mut out = vec![0u64; 72 * products.len()];
    check(unsafe { sys::tb200_multi_pairing_batch(pa.as_ptr(), pb.as_ptr(), products.len(), width, out.as_mut_ptr()) });
    out.chunks(72).map(|w| PairingOutput(unpack_gt(w.try_into().unwrap()))).collect()
}

// ---- `MippProof::prove` (src/mipp.rs:31-153): vectors resident on the GPU across rounds ------------------------------------
pub struct Mipp { a: *mut sys::tb200_mipp, h: *mut sys::tb200_mipp_g2 }
pub struct Round { pub comm_u_l: G1Affine, pub comm_u_r: G1Affine, pub comm_t_l: Fq12, pub comm_t_r: Fq12 }
impl Mipp {
    pub fn begin(a: &[G1Affine], y: &[Fr], h: &[G2Affine]) -> Self {
        let (pa, ph) = (pack_g1(a), pack_g2(h));
        let (mut ha, mut hh) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(unsafe { sys::tb200_mipp_g1_begin(pa.as_ptr(), fr_limbs(y), a.len(), sys::TB200_SCALARS_MONT, &mut ha) });
        check(unsafe { sys::tb200_mipp_g2_begin(ph.as_ptr(), h.len(), sys::TB200_SCALARS_MONT, &mut hh) });
        Mipp { a: ha, h: hh }
    }
    pub fn len(&self) -> usize { unsafe { sys::tb200_mipp_g1_len(self.a) } }
    /// src/mipp.rs:77-94: try_par!{ multiexponentiation x 2 } and par!{ pairings_product x 2 } in one call
    pub fn cross(&mut self) -> Round {
        let (mut ul, mut ur, mut tl, mut tr) = ([0u64; 12], [0u64; 12], [0u64; 72], [0u64; 72]);
        check(unsafe { sys::tb200_mipp_cross_all(self.a, self.h, ul.as_mut_ptr(), ur.as_mut_ptr(), tl.as_mut_ptr(), tr.as_mut_ptr()) });
        Round { comm_u_l: unpack_g1(&ul), comm_u_r: unpack_g1(&ur), comm_t_l: unpack_gt(&tl), comm_t_r: unpack_gt(&tr) }
    }
    /// src/mipp.rs:106-114: compress(m_a, c), compress_field(m_y, c_inv), compress(m_h, c_inv) -- enqueued, not awaited
    pub fn fold(&mut self, c: &Fr, c_inv: &Fr) {
        check(unsafe { sys::tb200_mipp_g1_fold(self.a, c as *const Fr as *const u64, c_inv as *const Fr as *const u64) });
        check(unsafe { sys::tb200_mipp_g2_fold(self.h, c_inv as *const Fr as *const u64) });
    }
    /// src/mipp.rs:121-122: (final_a, final_h)
    pub fn finish(self) -> (G1Affine, G2Affine) {
        let (mut a, mut y, mut h) = ([0u64; 12], [0u64; 4], [0u64; 24]);
        check(unsafe { sys::tb200_mipp_g1_read(self.a, a.as_mut_ptr(), y.as_mut_ptr()) });
        check(unsafe { sys::tb200_mipp_g2_read(self.h, h.as_mut_ptr()) });
        (unpack_g1(&a), unpack_g2(&h))
    }
}
impl Drop for Mipp { fn drop(&mut self) { unsafe { sys::tb200_mipp_g1_end(self.a); sys::tb200_mipp_g2_end(self.h); } } }

// ---- pinned host memory placed next to the GPUs (rows of Z, bases of a large MSM) ------------------------------------------
pub struct PinnedBuf { pub ptr: *mut c_void, pub bytes: usize }
impl PinnedBuf {
    /// `units` elements of `unit_bytes`, split into the per-GPU ranges the sharded entry points use
    pub fn sharded(units: usize, unit_bytes: usize) -> Self {
        let mut p = std::ptr::null_mut();
        check(unsafe { sys::tb200_host_alloc_sharded(units, unit_bytes, &mut p) });
        PinnedBuf { ptr: p, bytes: units * unit_bytes }
    }
}
impl Drop for PinnedBuf { fn drop(&mut self) { unsafe { sys::tb200_host_free(self.ptr) }; } }
