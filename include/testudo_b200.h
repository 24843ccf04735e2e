/* testudo_b200 -- C ABI of the B200-native BLS12-377 G1 MSM engine for Testudo's commitment hot path.
 *
 * This is the drop-in boundary: the entry points are exactly what a Rust `-sys` crate for the reference would
 * bind (INTEGRATION.md shows the stub). All citations are into the reference tree (/root/reference).
 *
 * Data layouts (identical to arkworks' in-memory values, so the Rust side copies limbs verbatim):
 *   G1 affine point : 12 x u64 = x[6] || y[6], little-endian limbs, Montgomery form (R = 2^384);
 *                     all-zero == point at infinity ((0,0) is not on y^2 = x^3 + 1).
 *   Fr scalar       : 4 x u64 little-endian. Canonical (`BigInt<4>`, what `msm_bigint` takes) by default;
 *                     pass TB200_SCALARS_MONT when handing over `Fr` values as stored by ark-ff
 *                     (Montgomery, R = 2^256) -- the conversion `into_bigint()` then happens on the GPU.
 * Results are canonical affine points, i.e. the value of `.into_affine()` on what arkworks returns.
 *
 * Return value: 0 on success; negative = argument error (TB200_E_*); positive = cudaError_t.
 * `tb200_last_error()` gives a thread-local message. All calls are blocking and thread-safe (they are issued
 * concurrently by rayon workers in the reference: src/sqrt_pst.rs:121-125, src/mipp.rs:77-85 via
 * src/macros.rs:1-17); the library serialises them.
 * Multi-GPU: ONE process drives all the GPUs it names in tb200_init_devices -- the reference is one
 * process that fans rows out internally (src/sqrt_pst.rs:121-125). The host-facing entry points then shard by
 * themselves: a large single MSM by point range, row commitments by row range over the replicated SRS, a pairing
 * product by pair range; each GPU's share is uploaded and computed by its own worker thread, and the per-GPU partial
 * results are combined by one ncclAllGather (single-process NCCL clique over NVLink) inside the call. The *_dev
 * entry points, MIPP and the PST openings run on the primary device (devices[0]). One process per GPU (torchrun,
 * testudo_b200/parallel.py) remains possible: every process then initialises its own single device.
 */
#ifndef TESTUDO_B200_H
#define TESTUDO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB200_OK 0
#define TB200_E_ARG (-1)      /* null pointer / bad size / bad flag */
#define TB200_E_STATE (-2)    /* library not initialised, bad handle */
#define TB200_E_LIMIT (-3)    /* problem exceeds an engine limit */

#define TB200_SCALARS_MONT 1u /* scalars are ark-ff Montgomery-form Fr, not canonical BigInt<4> */

/* ---- lifecycle -------------------------------------------------------------------------------------- */
/* Select `device` (-1 = current) and create the context; idempotent. Fails loudly without a CUDA device:
 * there is no CPU fallback anywhere in this library. */
int tb200_init(int device);
/* One context per listed CUDA ordinal, devices[0] = primary; builds the NCCL clique when ndevices > 1 (libnccl is
 * dlopen'ed then, not before). Idempotent; a later call may append devices behind the same primary (SRS handles loaded
 * before must be reloaded). Replaces nothing in the reference (it has no device notion) -- called once by the -sys
 * crate's initialiser, INTEGRATION.md. */
int tb200_init_devices(const int* devices, int ndevices);
int tb200_device_count(void); /* contexts created so far (0 before init) */
void tb200_shutdown(void);
const char* tb200_last_error(void);
/* number of kernels this library has launched since init / reset (bench.py reports it as gpu_launches) */
uint64_t tb200_launch_count(void);
void tb200_reset_launch_count(void);

/* ---- single variable-base MSM ------------------------------------------------------------------------
 * Replaces `<G1Projective as VariableBaseMSM>::msm_bigint(bases, bigints)` /
 * `msm_unchecked(bases, scalars)` followed by `.into_affine()` (ark-ec 0.4; call sites
 * src/sqrt_pst.rs:198, src/mipp.rs:385-394, src/commitments.rs:70-86, src/nizk/bullet.rs:93-118).
 * n == 0 yields the identity. `msm_unchecked`'s truncate-to-min(len) and `msm`'s Err(min_len) length rules
 * live in the host wrapper (testudo_b200/msm.py, host/msm.hpp), which passes n = min(len). */
int tb200_msm_g1(const uint64_t* bases_xy, const uint64_t* scalars, size_t n, unsigned flags,
                 uint64_t out_xy[12]);
/* Same, all pointers are DEVICE pointers (16-byte aligned); result (96 bytes) written to d_out_xy.
 * `stream` is a cudaStream_t (NULL = the library's stream); the call returns after enqueueing. */
int tb200_msm_g1_dev(const void* d_bases_xy, const void* d_scalars, size_t n, unsigned flags,
                     void* d_out_xy, void* stream);
/* One MSM whose inputs are RESIDENT on the GPUs: device slot i (order of tb200_init_devices) holds n[i] points and
 * scalars at d_bases_xy[i] / d_scalars[i] (n[i] may be 0). Per-GPU partial points, one all-gather of 96 bytes per GPU,
 * sum on the primary, result to the host ("single large MSM"). Blocking. */
int tb200_msm_g1_sharded_dev(const void* const* d_bases_xy, const void* const* d_scalars, const size_t* n,
                             unsigned flags, uint64_t out_xy[12]);

/* ---- G2 multi-scalar multiplication ------------------------------------------------
 * Replaces `<E::G2 as VariableBaseMSM>::msm_unchecked / msm_bigint` + `.into_affine()` for
 * `ark_bls12_377::G2Projective`: the G2 openings of `MultilinearPC::open` (src/sqrt_pst.rs:225) and MIPP's
 * `commit_g2` (src/mipp.rs:114). A G2 affine point is 24 u64 = x.c0[6] || x.c1[6] || y.c0[6] || y.c1[6]
 * (Fq2 = Fq[u]/(u^2+5), coordinates in Montgomery form, little-endian limbs); all-zero == identity.
 * Scalars and `flags` as for tb200_msm_g1. The result is the canonical affine point.
 * PRECONDITION: the bases lie in G2, the order-r subgroup of the twist (always true for a CRS): the window combine
 * uses the twisted Frobenius, which is multiplication by the curve parameter only there. */
int tb200_msm_g2(const uint64_t* bases, const uint64_t* scalars, size_t n, unsigned flags, uint64_t out[24]);
/* Same with DEVICE pointers (16-byte aligned); result (192 bytes) written to d_out; returns after enqueueing. */
int tb200_msm_g2_dev(const void* d_bases, const void* d_scalars, size_t n, unsigned flags, void* d_out,
                     void* stream);
/* `compress` on a G2 vector (src/mipp.rs:133, 354-367): vec[i] = vec[i] + scaler * vec[split + i], i < split */
int tb200_compress_g2(uint64_t* vec, size_t split, const uint64_t scaler[4], unsigned flags);

/* MIPP's G2 commitment key m_h (src/mipp.rs:43,114) kept on the device and folded on a stream of its own, so
 * the G2 folds overlap the G1 rounds: begin uploads `h` (n x 24 u64, n a power of two), fold(c_inv) ENQUEUES
 * h[i] <- h[i] + c_inv * h[n/2 + i] and halves the length, read waits and downloads the current vector (len() points;
 * final_h after the last round), end frees. `flags` selects the representation of c_inv. */
typedef struct tb200_mipp_g2* tb200_mipp_g2_t;
int tb200_mipp_g2_begin(const uint64_t* h_vec, size_t n, unsigned flags, tb200_mipp_g2_t* out);
size_t tb200_mipp_g2_len(tb200_mipp_g2_t h);
int tb200_mipp_g2_fold(tb200_mipp_g2_t h, const uint64_t c_inv[4]);
int tb200_mipp_g2_read(tb200_mipp_g2_t h, uint64_t* out);
int tb200_mipp_g2_end(tb200_mipp_g2_t h);

/* ---- PST openings: `MultilinearPC::open` (G2 proofs, src/sqrt_pst.rs:225) and the fork's `open_g1` (G1 proofs,
 * src/mipp.rs:144) of ark-poly-commit 0.4 multilinear_pc (call sites M6 and X1) ------------
 * evals: the 2^nv evaluations of the polynomial (`to_evaluations()` order); point: nv field elements; both in the
 * representation selected by `flags` (TB200_SCALARS_MONT = ark's in-memory Fr). level_bases[i] = the CRS level used
 * for variable i: 2^(nv - i) affine points (`ck.powers_of_h[off + i]` resp. `ck.powers_of_g[off + i]`, off =
 * ck.nv - nv for the variable-CRS fork). For i in 0..nv, k = nv - i:
 *     q_k[b] = r_k[2b+1] - r_k[2b],  r_{k-1}[b] = r_k[2b] (1 - point[i]) + r_k[2b+1] point[i],
 *     proofs[i] = MSM(level_bases[i], q_k duplicated to length 2^k).into_affine()
 * The quotient loop runs on the device (no host pass over the evaluations). proofs: nv x 12 (G1) / nv x 24 (G2). */
int tb200_pst_open_g1(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                      unsigned flags, uint64_t* proofs);
int tb200_pst_open_g2(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                      unsigned flags, uint64_t* proofs);
/* The same in two halves: _begin uploads and ENQUEUES the whole opening (nv >= 1) on streams of its own and returns;
 * _end waits, writes the proofs and releases the handle. `Polynomial::open` uses it to run the G2 opening of q
 * (src/sqrt_pst.rs:218-225), which does not depend on the MIPP transcript, NEXT TO the MIPP rounds (src/sqrt_pst.rs:212). */
typedef struct tb200_pst_open* tb200_pst_open_t;
int tb200_pst_open_g1_begin(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                            unsigned flags, tb200_pst_open_t* out);
int tb200_pst_open_g2_begin(const uint64_t* evals, size_t nv, const uint64_t* point, const uint64_t* const* level_bases,
                            unsigned flags, tb200_pst_open_t* out);
int tb200_pst_open_end(tb200_pst_open_t h, uint64_t* proofs);

/* ---- shared-base (SRS) batched MSM -------------------------------------------------------------------
 * Replaces the row fan-out `self.polys.par_iter().map(|p| MultilinearPC::commit(ck, p))`
 * (src/sqrt_pst.rs:121-125: 2^m_col MSMs over ck.powers_of_g[0]) and the Hyrax fan-out
 * `DensePolynomial::commit_inner` (src/dense_mlpoly.rs:315-329 over gens_n.G) by ONE call.
 * The SRS is uploaded once; window tables 2^(c*w) * G_j are precomputed on the GPU so every row needs a
 * single bucket set. window_bits = 0 picks c automatically. */
typedef struct tb200_srs* tb200_srs_t;
int tb200_srs_load(const uint64_t* bases_xy, size_t n, int window_bits, tb200_srs_t* out);
int tb200_srs_free(tb200_srs_t srs);
size_t tb200_srs_size(tb200_srs_t srs);
/* out_xy[i] = sum_j scalars[i*row_stride + j*col_stride] * G_j, i < rows, j < cols <= srs size.
 * Strides are in scalars (4 x u64). Un-transposed Z of sqrt_pst (Z[(j << m_col) | i], src/sqrt_pst.rs:58):
 * row_stride = 1, col_stride = 2^m_col. Contiguous rows (Hyrax, src/dense_mlpoly.rs:325): row_stride = cols,
 * col_stride = 1. */
int tb200_msm_g1_batch(tb200_srs_t srs, const uint64_t* scalars, size_t rows, size_t cols, ptrdiff_t row_stride,
                       ptrdiff_t col_stride, unsigned flags, uint64_t* out_xy);
/* rows given as separate heap buffers (what `Polynomial::commit` holds: self.polys[i].Z, src/sqrt_pst.rs:48-62) */
int tb200_msm_g1_batch_ptrs(tb200_srs_t srs, const uint64_t* const* row_ptrs, size_t rows, size_t cols,
                            unsigned flags, uint64_t* out_xy);
int tb200_msm_g1_batch_dev(tb200_srs_t srs, const void* d_scalars, size_t rows, size_t cols, ptrdiff_t row_stride,
                           ptrdiff_t col_stride, unsigned flags, void* d_out_xy, void* stream);
/* Hyrax rows WITH blinds: out[i] = MSM(gens_n.G, row_i) + blinds[i] * h -- `PedersenCommit::commit_slice`
 * (src/commitments.rs:80-86) under `DensePolynomial::commit_inner` (src/dense_mlpoly.rs:315-329) with
 * `commit(gens, random_tape)` (src/dense_mlpoly.rs:349-377). The blinding base h is loaded as one extra column of the
 * SRS (tb200_srs_load_blinded) and the blind rides along as the row's extra scalar: still ONE batched call.
 * scalars: rows x cols contiguous; cols == tb200_srs_size(srs) (the size WITHOUT h). */
int tb200_srs_load_blinded(const uint64_t* bases_xy, size_t n, const uint64_t h_xy[12], int window_bits, tb200_srs_t* out);
int tb200_msm_g1_batch_blinded(tb200_srs_t srs, const uint64_t* scalars, size_t rows, size_t cols, const uint64_t* blinds,
                               unsigned flags, uint64_t* out_xy);
/* `Polynomial::commit` in one call (src/sqrt_pst.rs:117-149): the row commitments (-> out_rows_xy, rows x 12) AND the IPP
 * commitment t = prod_i e(C_i, h_vec[i]) (-> out_t; h_vec = ck.powers_of_h[odd], rows x 24 u64, src/sqrt_pst.rs:128-144).
 * With several GPUs every GPU commits its row range, pairs ITS rows with its slice of h_vec (the rows never leave the GPU
 * between the two stages) and contributes one partial Miller product; one all-gather of 576 bytes per GPU and ONE final
 * exponentiation on the primary. Rows as separate heap buffers (self.polys) or as a strided view of Z. */
int tb200_sqrt_pst_commit(tb200_srs_t srs, const uint64_t* const* row_ptrs, size_t rows, size_t cols, unsigned flags,
                          const uint64_t* h_vec, uint64_t* out_rows_xy, uint64_t out_t[72]);
int tb200_sqrt_pst_commit_strided(tb200_srs_t srs, const uint64_t* scalars, size_t rows, size_t cols,
                                  ptrdiff_t row_stride, ptrdiff_t col_stride, unsigned flags, const uint64_t* h_vec,
                                  uint64_t* out_rows_xy, uint64_t out_t[72]);

/* ---- MIPP G1 steps, device-resident across rounds --------------------------------------------------------
 * Replaces the G1 work of `MippProof::prove` (src/mipp.rs:58-120): per round
 *   cross: comm_u_l = MSM(a_l, y_r), comm_u_r = MSM(a_r, y_l)           (src/mipp.rs:77-85, :385-394)
 *   fold : a_l[i] += c * a_r[i] (into_affine), y_l[i] += c_inv * y_r[i]  (src/mipp.rs:110-112, :354-383)
 * The vectors stay on the GPU; only two points per round come back for the transcript. */
typedef struct tb200_mipp* tb200_mipp_t;
int tb200_mipp_g1_begin(const uint64_t* a_xy, const uint64_t* y, size_t n, unsigned flags, tb200_mipp_t* out);
size_t tb200_mipp_g1_len(tb200_mipp_t h);
int tb200_mipp_g1_cross(tb200_mipp_t h, uint64_t comm_u_l[12], uint64_t comm_u_r[12]);
/* c and c_inv are Fr scalars in the representation selected by `flags` at begin(). The fold is only ENQUEUED (the
 * scalars are copied before the call returns); every later call on the handle is ordered behind it. It runs over the
 * G1 endomorphism (127 doublings), so the vector must lie in the order-r subgroup -- true for commitments. */
int tb200_mipp_g1_fold(tb200_mipp_t h, const uint64_t c[4], const uint64_t c_inv[4]);
/* current vectors (len() entries): a as affine points, y as scalars in the begin() representation */
int tb200_mipp_g1_read(tb200_mipp_t h, uint64_t* a_xy, uint64_t* y);
int tb200_mipp_g1_end(tb200_mipp_t h);
/* stand-alone `compress` for G1 (src/mipp.rs:354-367): vec[i] = vec[i] + scaler * vec[split + i], i < split */
int tb200_compress_g1(uint64_t* vec_xy, size_t split, const uint64_t scaler[4], unsigned flags);

/* ---- device buffers (for hosts that keep data resident across calls, e.g. Z between commit and open) -------------
 * Plain cudaMalloc / blocking copies on the library's device; pointers are valid for every *_dev entry point. */
int tb200_dev_alloc(size_t bytes, void** out);
int tb200_dev_free(void* d_ptr);
int tb200_dev_upload(void* d_dst, const void* h_src, size_t bytes);
int tb200_dev_download(void* h_dst, const void* d_src, size_t bytes);
int tb200_stream_sync(void); /* wait for work enqueued on the library's stream by *_dev calls with stream == NULL */
/* Page-locked host memory usable from every device of the library (cudaHostAlloc portable / cudaHostRegister): uploads
 * from such buffers are asynchronous and run at full PCIe rate; pageable buffers work too but are staged by the driver. */
int tb200_host_alloc(size_t bytes, void** out);
/* NUMA-placed pinned buffers. Page-locked pages live where they were first touched; eight GPUs reading 2 GiB each from
 * one socket's memory are bound by that socket, not by PCIe (round 1: 0.72 end-to-end efficiency at 8 GPUs). _near places
 * the buffer on the NUMA node of device slot `device_slot`; _sharded splits `units` elements of `unit_bytes` into the
 * SAME contiguous per-device ranges the sharded entry points use (point ranges of tb200_msm_g1, row ranges of a row-major
 * batch) and places every range next to the GPU that will read it. Without topology in sysfs both are plain pinned
 * allocations. Free with tb200_host_free. tb200_device_numa_node: the node of a device slot, -1 if unknown. */
int tb200_host_alloc_near(size_t bytes, int device_slot, void** out);
int tb200_host_alloc_sharded(size_t units, size_t unit_bytes, void** out);
int tb200_device_numa_node(int device_slot);
int tb200_host_free(void* h_ptr);
int tb200_host_register(void* h_ptr, size_t bytes);
int tb200_host_unregister(void* h_ptr);

/* ---- sqrt_pst scalar work on the device (Fr values in ark Montgomery form) ---------------
 * chis_out[i] = prod_j (bit(i, m-1-j) ? b[j] : 1 - b[j]), i < 2^m  -- `Polynomial::get_chi_i`, src/sqrt_pst.rs:152-166 */
int tb200_fr_chis(const uint64_t* b, size_t m, uint64_t* chis_out);
/* out[i] = prod_{j : bit(i, m-1-j) set} b[j], i < 2^m -- `MippProof::polynomial_evaluations_from_transcript(cs_inv)`
 * (src/mipp.rs:159-180) with b = cs_inv: the evaluations of the structured polynomial behind final_h, produced directly in
 * ark's Montgomery form for commit_g2 / open_g1 (src/mipp.rs:128-144). */
int tb200_fr_subset_products(const uint64_t* b, size_t m, uint64_t* out);
/* out[j] = sum_i Z[j*cols + i] * v[i], j < rows -- `get_q` (src/sqrt_pst.rs:81-101) with Z[(j << m_col) | i], v = chis,
 * and `eval` (src/sqrt_pst.rs:105-115) with rows = 1. HBM-bound: 32 B per multiply-add. */
int tb200_fr_matvec(const uint64_t* Z, size_t rows, size_t cols, const uint64_t* v, uint64_t* out);
int tb200_fr_matvec_dev(const void* d_Z, size_t rows, size_t cols, const void* d_v, void* d_out, void* stream);

/* ---- group utilities ------------------------------------------------------------------------------------ */
/* `rows` INDEPENDENT tiny MSMs in one launch: out[i] = sum_{j < per_row} scalars[i*per_row + j] * bases[i*per_row + j],
 * per_row <= 8 (one warp per row, a quad of lanes per point, the small-n Straus kernel). The verifier side of the path:
 * ark-poly-commit 0.4 `MultilinearPC::check` forms `commitment - g*value` and `g_mask_random[i] - g*point[i]` for every
 * variable (reached from `Polynomial::verify`, src/sqrt_pst.rs:261) -- nv + 1 two-point MSMs for the latency of one.
 * Accepts any curve point; scalars canonical, or Montgomery with TB200_SCALARS_MONT; (0, 0) in / out = identity. */
int tb200_msm_g1_each(const uint64_t* bases_xy, const uint64_t* scalars, size_t rows, size_t per_row, unsigned flags,
                      uint64_t* out_xy);
/* The same for rows of DIFFERENT lengths (0 .. 1024 points each): row i takes the next row_len[i] entries of bases /
 * scalars. `MippProof::verify` + `MultilinearPC::check` need a (2m + 2)-point fold (src/mipp.rs:240-276,310-316), an
 * m-point fold (check_2, :307) and nv + 1 two-point rows (src/sqrt_pst.rs:261): three kinds of independent MSMs, one
 * launch. An empty row yields the identity. */
int tb200_msm_g1_rows(const uint64_t* bases_xy, const uint64_t* scalars, const size_t* row_len, size_t rows, unsigned flags,
                      uint64_t* out_xy);
/* A single G1 MSM IN FLIGHT next to other calls: _begin uploads and enqueues on a side pipeline of the library (own stream,
 * own workspace) and returns, _end waits and writes the affine point. The reference runs independent MSMs side by side
 * (`try_par!` / `rayon::join`, src/macros.rs:1-17) and computes `MultilinearPC::commit(ck, &q)` only to feed a debug_assert
 * (src/sqrt_pst.rs:205-206) -- nothing on the prover's critical path waits for it. bases / scalars as for tb200_msm_g1;
 * the host buffers must stay valid until _end returns; _end(job, NULL) discards the result. n < 2^26. */
typedef struct tb200_msm_job* tb200_msm_job_t;
int tb200_msm_g1_begin(const uint64_t* bases_xy, const uint64_t* scalars, size_t n, unsigned flags, tb200_msm_job_t* out);
int tb200_msm_g1_end(tb200_msm_job_t job, uint64_t out_xy[12]);
/* out = sum of n affine points (combining per-GPU partial results after the NCCL all-gather) */
int tb200_g1_sum(const uint64_t* pts_xy, size_t n, uint64_t out_xy[12]);
int tb200_g1_sum_dev(const void* d_pts_xy, size_t n, void* d_out_xy, void* stream);
/* d_out[i * nb + j] = A_i + B_j (affine): synthetic bases with known discrete logs */
int tb200_g1_outer_sum_dev(const void* d_a_xy, size_t na, const void* d_b_xy, size_t nb, void* d_out_xy,
                           void* stream);

/* ---- profiling / tuning ----------------------------------------------------------------------------------- */
/* When enabled, every MSM call records CUDA events around its stages on the launching stream. */
void tb200_set_profiling(int enabled);
/* milliseconds of the named stage in the most recent profiled call: "digits", "scan", "scatter",
 * "accumulate", "fixup", "reduce", "finalize", "total". Returns < 0 if unknown / not profiled. */
double tb200_stage_ms(const char* stage);
/* geometry of the most recent MSM call: window bits c, windows W, sorted entries M, buckets B, segment K */
int tb200_last_geometry(int* c, int* windows, uint64_t* entries, uint64_t* buckets, int* segment);
/* override the automatic window choice for single MSMs (0 = automatic) */
void tb200_set_window_bits(int c);
/* bucket-accumulation method (XYZZ mixed additions over balanced segments, operands in shared-memory slots,
 * k_accumulate_s): 0 = automatic (= 4); 3 = plain CIOS products; 4 = Y3 as one fused sum of two products
 * (mont_mul2_lazy). Identical results. The measured dead ends (register operands, batched-affine rounds, Karatsuba)
 * live in csrc/experimental/ and are not part of the library. */
void tb200_set_accumulate_mode(int mode);
/* test / tuning hooks: the number of sorted entries one pipeline pass may index (default 2^32 - 1024; a single MSM
 * with more entries, n * windows, runs as point-range passes over one persistent bucket array; 0 restores the default),
 * and the work per GPU (points of a single MSM, scalars of a batch) below which a host-facing call is NOT sharded
 * (default 2^18; 0 restores it). */
void tb200_set_pass_entries_max(uint64_t entries);
void tb200_set_shard_min(size_t units);
/* tb200_sqrt_pst_commit: 1 = run the Miller loops of every row chunk on a side stream next to the row MSMs of the
 * following chunk (one partial product per chunk); 0 (default) = all Miller loops behind the row stage. See DESIGN.md
 * (multi-GPU) for the measurements behind the default. */
void tb200_set_commit_pipeline(int enabled);
/* single G1 MSMs of up to `n` points (default and maximum 1024; 0 disables) skip the sort pipeline: Straus with radix-16
 * signed digits, a quad of lanes per point, 8 points per one-warp CTA: `commit_scalar` (src/commitments.rs:70-77), the bullet rounds
 * (src/nizk/bullet.rs:93-118), the last MIPP rounds (src/mipp.rs:77-85). Identical results. */
void tb200_set_small_msm_max(int n);
/* large single G1 MSMs: 1 = the pipeline runs as three staggered window ranges on three streams (device-resident inputs)
 * resp. sorts the next point-range chunk next to the accumulation of the current one (host inputs), so that the
 * memory-bound sort stages and the latency-bound reduction tails could hide behind the integer-pipe-bound accumulation;
 * 0 (default) = everything on one stream. Identical results. Measured on B200: 84.2 -> 84.0 ms resident, 92.6 -> 91.4 ms
 * from host buffers at 2^24 points -- the accumulation fills the register file (4 CTAs x 128 threads x 128 registers per
 * SM), so the side streams' CTAs only displace accumulation CTAs; kept as an option, off by default.
 * Ignored while stage profiling is on. */
void tb200_set_msm_overlap(int enabled);
/* host-facing single MSMs of >= 2^23 points per GPU are uploaded in point-range chunks. pace != 0 (default): chunk k is
 * uploaded once chunk k-2 has been accumulated (double buffering) instead of as early as the copy engine allows, so that
 * GPUs behind a fast host link do not take bandwidth from the ones behind a slow link up front. `sixteenths`: the chunk
 * sizes in sixteenths of the points (count <= 16, sum 16); count = 0 restores the built-in schedule. */
int tb200_set_host_upload(int pace, const int* sixteenths, int count);
/* integer-pipe microbenchmark: runs `iters` dependent-chain iterations of wide MACs on every SM and returns the
 * achieved 32x32->64 multiply-accumulates per second in *out_macs_per_s (kind: 0 = IMAD.WIDE.U32 reg-reg,
 * 1 = IMAD (32-bit lo), 2 = full Fq Montgomery multiplications per second). */
int tb200_int_pipe_peak(int kind, int iters, double* out_per_s);

/* ---- kernel unit-test hooks (device field / group arithmetic on n independent lanes) ----------------------- */
int tb200_test_fq_mul(const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out);      /* 6 u64 each */
int tb200_test_fq_addsub(const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out_add, uint64_t* out_sub);
int tb200_test_g1_add(const uint64_t* p_xy, const uint64_t* q_xy, size_t n, uint64_t* out_xy); /* via XYZZ madd */
int tb200_test_g1_mul(const uint64_t* p_xy, const uint64_t* k, size_t n, uint64_t* out_xy);    /* k canonical */
int tb200_test_g2_add(const uint64_t* p, const uint64_t* q, size_t n, uint64_t* out);           /* 24 u64 each */
int tb200_test_g2_mul(const uint64_t* p, const uint64_t* k, size_t n, uint64_t* out);           /* k canonical */
/* one Fq12 operation per element (see kernels_pairing.cuh k_test_fq12_op for the op codes); a, b, out: n x 72 u64 */
int tb200_test_fq12_op(int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out);

/* ---- pairing products ---------------------------------------------------------------
 * GT element = Fq12 = 72 x u64: twelve Fq (6 LE limbs, Montgomery form) in ark's in-memory tower order
 * c0.c0.c0, c0.c0.c1, c0.c1.c0, ... c1.c2.c1 (Fq12 = Fq6[w]/(w^2 - v), Fq6 = Fq2[v]/(v^3 - u), Fq2 = Fq[u]/(u^2 + 5)).
 * tb200_multi_pairing replaces `E::multi_pairing(g1s, g2s).0` (ark-ec 0.4 `Bls12`: Miller loops, product, one final
 * exponentiation with ark's exponent 3 (q^12 - 1)/r-equivalent chain) at src/sqrt_pst.rs:131-144 (`t`) and
 * `pairings_product`, src/mipp.rs:396-398. Pairs with an identity on either side contribute 1, as in ark; n == 0
 * yields 1. */
int tb200_multi_pairing(const uint64_t* g1_xy, const uint64_t* g2, size_t n, uint64_t out[72]);
/* `products` independent pairing products of `pairs_each` pairs each in ONE pass of the pairing engine; product p takes the
 * pairs [p * pairs_each, (p + 1) * pairs_each) and lands at out + 72 p. Identity pairs contribute 1, so shorter products
 * are padded with (0, 0) points. The verifier side of the path evaluates five products -- `E::pairing(final_a, final_h)`
 * (src/mipp.rs:311), both sides of `check_2` (:307) and of `MultilinearPC::check` (src/sqrt_pst.rs:261) -- for the latency
 * of one. */
int tb200_multi_pairing_batch(const uint64_t* g1_xy, const uint64_t* g2, size_t products, size_t pairs_each, uint64_t* out);
/* Same with DEVICE pointers; d_out receives 576 bytes; returns after enqueueing on `stream` (NULL = library stream). */
int tb200_multi_pairing_dev(const void* d_g1_xy, const void* d_g2, size_t n, void* d_out, void* stream);
/* Sharded pairing product (one process per GPU): each rank reduces ITS slice of the pairs to one partial
 * value -- the product of the Miller-loop values, NO final exponentiation (576 bytes) --, the partials are all-gathered,
 * and tb200_gt_product_final_exp multiplies them and applies the single final exponentiation:
 *   multi_pairing(a, b) == gt_product_final_exp([miller_product(slice_r) for every rank r]).
 * The partial value is only meaningful as input of the combination (it depends on the Miller-loop formulas). */
int tb200_miller_product(const uint64_t* g1_xy, const uint64_t* g2, size_t n, uint64_t out[72]);
int tb200_miller_product_dev(const void* d_g1_xy, const void* d_g2, size_t n, void* d_out, void* stream);
int tb200_gt_product_final_exp(const uint64_t* parts, size_t n, uint64_t out[72]);
int tb200_gt_product_final_exp_dev(const void* d_parts, size_t n, void* d_out, void* stream);
/* The two cross pairing products of a MIPP round over the device-resident vectors (src/mipp.rs:87-94):
 * comm_t_l = prod_i e(a[i], h[split + i]), comm_t_r = prod_i e(a[split + i], h[i]), split = len / 2.
 * Both handles must have the same current length (>= 2). Waits for the G2 folds enqueued so far. */
int tb200_mipp_pairing_cross(tb200_mipp_t a, tb200_mipp_g2_t h, uint64_t comm_t_l[72], uint64_t comm_t_r[72]);
/* A whole MIPP round's prover values in one call (src/mipp.rs:77-94): the cross MSMs comm_u_l / comm_u_r and the cross
 * pairing products comm_t_l / comm_t_r are enqueued on separate streams and awaited together. */
int tb200_mipp_cross_all(tb200_mipp_t a, tb200_mipp_g2_t h, uint64_t comm_u_l[12], uint64_t comm_u_r[12],
                         uint64_t comm_t_l[72], uint64_t comm_t_r[72]);
/* out[i] = base[i] ^ exps[i] in GT (the verifier's `tx.pow(c)`, src/mipp.rs:258-261); exponents are Fr values
 * (canonical, or Montgomery with TB200_SCALARS_MONT). */
/* tuning/test hook: pairing products of up to `n` pairs run one WARP per Miller loop, larger ones one THREAD (default 8192) */
void tb200_set_pairing_coop_max(int n);
/* tuning/test hook, the Miller kernel of the cooperative pairing engine: 96 = pipelined (three warps per pair, the point
 * chain next to the f chain: lowest latency), 64 / 32 = one pair per two-warp / one-warp CTA, 33 = two pairs per warp with
 * a shared accumulator (highest throughput; even products only); the product tree and the final exponentiation run on
 * two warps unless 32 is forced. 0 (default) = by size: 96 up to 512 pairs, 33 above. Identical results. */
void tb200_set_pairing_team(int lanes);
int tb200_gt_pow(const uint64_t* bases, const uint64_t* exps, size_t n, unsigned flags, uint64_t* out);
/* out = prod_i bases[i] ^ exps[i] in GT: the TC half of the verifier's parallel fold / reduce over `MippTU`
 * (src/mipp.rs:240-271: `tx.pow(c)`, `res.tc.mul_assign(&tx)`, `merge`) in one call -- every power on its own team of
 * lanes, then the product tree of the pairing engine. Generic field arithmetic (proof values need not be unitary);
 * n == 0 yields 1. Exponents as for tb200_gt_pow. */
int tb200_gt_multi_pow(const uint64_t* bases, const uint64_t* exps, size_t n, unsigned flags, uint64_t out[72]);

/* ---- Poseidon sponge of the Fiat-Shamir transcript (HOST code, as in the reference) ----------------
 * Restates ark-crypto-primitives 0.4 `PoseidonSponge<F>` behind `PoseidonTranscript<F>` (src/poseidon_transcript.rs:12-125).
 * field: 0 = BLS12-377 Fr (benches/pst.rs:23,57), 1 = BLS12-377 Fq (`PoseidonTranscript<E::BaseField>`, src/mipp.rs:32,
 * src/sqrt_pst.rs:170,325). ark / mds: (full + partial) x (rate + capacity) round constants and the square MDS matrix,
 * canonical little-endian limbs of `field` (4 or 6 u64 each), row-major -- `PoseidonConfig::new(full, partial, alpha, mds,
 * ark, rate, capacity)`, src/parameters.rs:156-185. absorb_bytes = `sponge.absorb(&Vec<u8>)` (what `append` does with the
 * uncompressed serialisation, :21-27), absorb_native = `absorb(&F)`, squeeze_fr = `challenge_scalar::<Fr>` (:29-31; the
 * foreign-field squeeze when the sponge is over Fq), canonical limbs. No CUDA involved; usable before tb200_init. */
typedef struct tb200_poseidon* tb200_poseidon_t;
int tb200_poseidon_new(int field, unsigned full_rounds, unsigned partial_rounds, uint64_t alpha, unsigned rate,
                       unsigned capacity, const uint64_t* ark, const uint64_t* mds, tb200_poseidon_t* out);
int tb200_poseidon_reset(tb200_poseidon_t h);
int tb200_poseidon_absorb_bytes(tb200_poseidon_t h, const uint8_t* data, size_t len);
/* `transcript.append(label, &value)` (src/poseidon_transcript.rs:21-27) for a value in the C ABI's word layout: nwords = 4
 * (Fr), 12 (G1 affine), 24 (G2 affine) or 72 (Fq12 / GT), Montgomery limbs. Encodes the value's uncompressed
 * `CanonicalSerialize` bytes (ark-serialize 0.4: canonical little-endian coordinates, SWFlags in the last byte) and absorbs
 * them exactly as tb200_poseidon_absorb_bytes would -- without a round trip through the host language's big integers. */
int tb200_poseidon_append_words(tb200_poseidon_t h, const uint64_t* words, size_t nwords);
int tb200_poseidon_absorb_native(tb200_poseidon_t h, const uint64_t* elems, size_t n);
int tb200_poseidon_squeeze_native(tb200_poseidon_t h, uint64_t* out, size_t n);
int tb200_poseidon_squeeze_fr(tb200_poseidon_t h, uint64_t out[4]);
int tb200_poseidon_limbs(tb200_poseidon_t h);
int tb200_poseidon_free(tb200_poseidon_t h);

#ifdef __cplusplus
}
#endif
#endif /* TESTUDO_B200_H */
