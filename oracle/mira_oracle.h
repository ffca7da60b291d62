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

/* ==== witness side (mira_oracle_witness.c): rows a5, a7-a9, a12 of SURVEY.md §8 ==================== */

/* RelaxedPlonkWitness::fold (src/plonk/mod.rs:1097-1134): out[i] = w1[i] + r*w2[i] */
void oracle_fold_w(int field, const void *w1, const void *w2, size_t n, const void *r, void *out);
/* out[i] = e[i] + sum_k r^(k+1) * terms[k][i] */
void oracle_fold_e(int field, const void *e, const void *const *terms, size_t n_terms, size_t n, const void *r,
                   void *out);
/* util::concatenate_with_padding (src/util.rs:189-193); returns elements written (out may be NULL) */
size_t oracle_concat_pad(const void *const *cols, const size_t *lens, size_t n_cols, size_t pad_size, void *out);

/* PlonkEvalDomain (src/plonk/eval.rs:93-106) as plain pointers; all HOST memory here. */
typedef struct {
  uint64_t row_size;                 /* GetDataForEval::row_size() */
  uint32_t num_selectors, num_fixed, num_advice, num_lookup, num_challenges, num_w1, num_w2, flags;
  const void *const *selectors;      /* [num_selectors] -> row_size bytes (Vec<bool>: 0/1 per row) */
  const void *const *fixed;          /* [num_fixed]     -> row_size x 32 B */
  const void *const *w1;             /* W1s[i] */
  const uint64_t *w1_len;            /* W1s[i].len() in elements */
  const void *const *w2;             /* W2s[i] */
  const uint64_t *w2_len;
  const void *challenges;            /* num_challenges x 32 B */
} oracle_eval_domain;

/* flags bit 0: the domain is a LookupEvalDomain (src/plonk/eval.rs:84-135): advice variable `index` is the separate
 * column w1[index] (w1_len[index] rows) instead of a slice of the concatenated W; w2 is unused. */
enum { ORACLE_EVAL_LOOKUP_DOMAIN = 1 };
/* plonk::eval::Error (src/plonk/eval.rs:3-25) */
enum {
  ORACLE_EVAL_OK = 0,
  ORACLE_EVAL_CHALLENGE_OUT_OF_BOUNDARY = -11,
  ORACLE_EVAL_COLUMN_OUT_OF_BOUNDARY = -12,
  ORACLE_EVAL_ROW_OUT_OF_BOUNDARY = -13,
  ORACLE_EVAL_INVALID_WITNESS_INDEX = -14,
  ORACLE_EVAL_BAD_PROGRAM = -15
};
/* GraphEvaluator::evaluate for rows [row_begin, row_end) of a serialised program (encoding: see
 * include/mira_b200.h, mira_eval_program_create).  out: (row_end-row_begin) x 32 B. */
int oracle_eval_rows(int field, const uint32_t *code, size_t code_words, const void *constants, size_t n_constants,
                     const int32_t *rotations, size_t n_rotations, uint32_t num_intermediates,
                     const oracle_eval_domain *dom, size_t row_begin, size_t row_end, void *out);

/* the same three loops split over `threads` pthreads (<= 0: all cores), as the reference runs them under rayon */
int oracle_eval_rows_mt(int field, const uint32_t *code, size_t code_words, const void *constants, size_t n_constants,
                        const int32_t *rotations, size_t n_rotations, uint32_t num_intermediates,
                        const oracle_eval_domain *dom, int threads, void *out);
void oracle_fold_w_mt(int field, const void *w1, const void *w2, size_t n, const void *r, int threads, void *out);
void oracle_fold_e_mt(int field, const void *e, const void *const *terms, size_t n_terms, size_t n, const void *r,
                      int threads, void *out);

/* lookup argument, SURVEY.md 8 row a6 (src/plonk/lookup.rs:278-319) */
void oracle_lookup_m(int field, const void *l, size_t n_l, const void *t, size_t n_t, void *out_m);
void oracle_lookup_h_g(int field, const void *l, const void *t, const void *m, size_t n, const void *r, void *out_h,
                       void *out_g);

/* fft::best_fft (src/fft.rs:51-115) in place over 2^log_n elements, and its helpers / wrappers */
void oracle_fft(int field, void *a, uint32_t log_n, const void *omega);
int  oracle_fft_omega(int field, uint32_t k, int is_inverse, void *out);   /* get_omega_or_inv; -1 if k > S */
void oracle_fft_divisor(int field, uint32_t k, void *out);                 /* get_ifft_divisor */
int  oracle_fft_forward(int field, void *a, uint32_t log_n);               /* fft  (src/fft.rs:160-162) */
int  oracle_fft_inverse(int field, void *a, uint32_t log_n);               /* ifft (src/fft.rs:165-175) */

int oracle_num_cores(void);

#ifdef __cplusplus
}
#endif
#endif
