/*
 * mira_oracle.h — CPU ORACLE for the Mira commitment hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or the reported CPU baseline.  The CUDA product (mira_b200/csrc) never links it.
 *
 * PARITY STATUS: "parity unpinned" at the commit() boundary.  The reference
 * (/root/reference, Rust) cannot be compiled here (no cargo/rustc) and its MSM arithmetic
 * lives in an un-vendored git dependency (halo2_proofs, branch joshbeal/dev-mira,
 * Cargo.toml:60-62; halo2curves re-exported through it, src/lib.rs:24-27).  The reference
 * holds NO golden vector for commit().  What this oracle IS pinned against (tests/test_oracle.py):
 *   - src/fft.rs:239-258            8-point FFT known-answer test over BN254 Fr   (field mul/add/sub)
 *   - src/polynomial/lagrange.rs:113-126   Lagrange values at X=2               (field inversion)
 *   - src/digest.rs:99-114          (r-1)*G == -G                                (curve add/double)
 *   - an independent Python big-integer evaluation of sum(s_i * P_i)             (tests/pyref.py)
 *   - the algebraic properties the reference's own tests rely on
 *     (src/nifs/vanilla/tests.rs:137-244 + src/plonk/mod.rs:547-557: commit is a homomorphism).
 *
 * Memory layouts restated from halo2curves (published crate, version un-pinned upstream):
 *   field element  : 4 x u64 little-endian limbs, Montgomery form, R = 2^256
 *   affine point   : { x, y } = 64 bytes; the identity is (0, 0)
 *   scalar         : 4 x u64 little-endian limbs, Montgomery form of the curve's scalar field
 */
#ifndef MIRA_ORACLE_H
#define MIRA_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_BN254_G1 = 0, ORACLE_GRUMPKIN_G1 = 1 };
/* field ids: the BN254 base field Fq (modulus p) and scalar field Fr (modulus r).
 * BN254 G1   : coordinates in Fq, scalars in Fr.
 * Grumpkin G1: coordinates in Fr, scalars in Fq (the 2-cycle). */
enum { ORACLE_FQ = 0, ORACLE_FR = 1 };

/* ---- field layer (all elements 32 bytes, Montgomery form unless noted) ---- */
void oracle_fe_from_u64(int field, uint64_t v, void *out);
void oracle_fe_from_canonical(int field, const void *canon_le32, void *out); /* canonical -> Montgomery */
void oracle_fe_to_canonical(int field, const void *a, void *canon_le32);     /* = PrimeField::to_repr  */
void oracle_fe_add(int field, const void *a, const void *b, void *out);
void oracle_fe_sub(int field, const void *a, const void *b, void *out);
void oracle_fe_mul(int field, const void *a, const void *b, void *out);
void oracle_fe_inv(int field, const void *a, void *out);                     /* 0 -> 0 */
/* n independent products out[i] = a[i]*b[i] (used to check the CUDA field kernels) */
void oracle_fe_mul_many(int field, const void *a, const void *b, size_t n, void *out);

/* ---- curve layer ---- */
void oracle_generator(int curve, void *out_affine64);
int  oracle_is_on_curve(int curve, const void *affine64);                    /* identity counts as on-curve */
void oracle_point_add_affine(int curve, const void *p64, const void *q64, void *out64);
void oracle_point_neg_affine(int curve, const void *p64, void *out64);
/* double-and-add scalar multiplication, scalar in Montgomery form of the scalar field */
void oracle_scalar_mul(int curve, const void *base64, const void *scalar32, void *out64);

/* ---- the hot path: CommitmentKey::commit (src/commitment.rs:78-87) ----
 * best_multiexp(v, &ck[..v.len()]).to_affine(), restated from halo2_proofs::arithmetic
 * (rayon chunk-per-thread Pippenger, unsigned windows c = ceil(ln n)).
 * threads <= 0 means "all online cores" (rayon's default pool size).
 * returns 0, or -1 for Error::TooLongInput (n > n_bases; nothing is written). */
int oracle_commit(int curve, const void *bases, size_t n_bases,
                  const void *scalars, size_t n, int threads, void *out_affine64);
/* the same sum by plain double-and-add per term (second opinion, O(n*256) group ops) */
void oracle_commit_naive(int curve, const void *bases, const void *scalars, size_t n, void *out_affine64);

/* ---- deterministic synthetic inputs (shared definition with mira_b200/csrc/testgen.cu) ----
 * word(seed, k) = splitmix64 output k of the stream seeded with `seed`.
 * scalar i  : limbs word(seed, 4i..4i+3), top limb masked to 62 bits, minus modulus if >= modulus,
 *             then converted to Montgomery form.
 * dist 0 = uniform; dist 1 = "witness-like" (60% zero, 25% {0,1}, 10% < 2^32, 5% uniform).
 * base i    : (k_i * G) with k_i = the uniform scalar i of stream `seed`, normalised to affine. */
void oracle_gen_scalars(int curve, uint64_t seed, size_t first, size_t n, int dist, void *out);
void oracle_gen_bases(int curve, uint64_t seed, size_t first, size_t n, int threads, void *out);

int oracle_num_cores(void);

#ifdef __cplusplus
}
#endif
#endif
