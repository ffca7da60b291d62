/*
 * mira_b200.h — C ABI of the B200-native commitment engine (drop-in for Mira's commit hot path).
 *
 * The reference has no FFI today; the seam this library replaces is the inherent method
 *     CommitmentKey::<C>::commit(&self, v: &[C::Scalar]) -> Result<C, commitment::Error>
 * (/root/reference/src/commitment.rs:78-87), used with C = bn256::G1Affine and grumpkin::G1Affine
 * (examples/groth16/main.rs:82-88).  INTEGRATION.md shows the Rust `extern "C"` shim a maintainer
 * adds; every entry point below cites the reference interface it stands in for.
 *
 * Byte layouts are the reference's in-memory layouts (halo2curves), used without conversion:
 *   scalar : 32 B = 4 x u64 LE limbs, MONTGOMERY form (R = 2^256) of the curve's scalar field
 *   point  : 64 B = { x, y }, each 4 x u64 LE limbs Montgomery form of the base field; (0,0) = identity
 * All pointers are plain host or device addresses; there are no torch types in this ABI.
 * There is no CPU fallback: every compute entry point fails with MIRA_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef MIRA_B200_H
#define MIRA_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mira_msm_ctx mira_msm_ctx;

enum { MIRA_BN254_G1 = 0, MIRA_GRUMPKIN_G1 = 1 };
enum { MIRA_FQ = 0, MIRA_FR = 1 };   /* BN254 base field (p) / scalar field (r) */

enum {
  MIRA_OK = 0,
  MIRA_ERR_TOO_LONG_INPUT = -1, /* commitment::Error::TooLongInput, src/commitment.rs:20-24,82-85 */
  MIRA_ERR_CUDA = -2,           /* any CUDA failure; the reference has no such path => caller aborts */
  MIRA_ERR_INVALID = -3,        /* bad argument (null pointer, unknown curve, ...) */
  MIRA_ERR_NOT_ON_CURVE = -4,   /* load_or_setup_cache's check, src/commitment.rs:145-153 */
  /* plonk::eval::Error (src/plonk/eval.rs:3-25), raised when a program is bound to a domain */
  MIRA_ERR_EVAL_CHALLENGE = -11,     /* ChallengeIndexOutOfBoundary */
  MIRA_ERR_EVAL_COLUMN = -12,        /* ColumnVariableIndexOutOfBoundary */
  MIRA_ERR_EVAL_ROW = -13,           /* RowIndexOutOfBoundary */
  MIRA_ERR_EVAL_WITNESS_INDEX = -14, /* InvalidWitnessIndex */
  MIRA_ERR_EVAL_PROGRAM = -15        /* malformed program encoding (no reference analogue) */
};

/* Text of the last error raised on the calling thread ("" if none). */
const char *mira_last_error(void);

/* ---- CommitmentKey<C> (src/commitment.rs:26-49) -------------------------------------------------
 * Uploads `n_bases` generators (the `ck: Box<[C]>` memory image, or the `.cache/.../{k}.bin` file
 * content, src/commitment.rs:96-124) to GPU `device` and keeps them resident.  Fixed-base window
 * tables (2^(c*j) * P_i) are derived lazily on first use of a window size and cached in the context.
 * `bases` may be a host pointer (bases_on_device = 0) or a device pointer on `device` (= 1). */
int mira_msm_ctx_create(int curve, const void *bases, size_t n_bases, int bases_on_device, int device,
                        mira_msm_ctx **out);
/* The same key spread over several GPUs of ONE process -- what the reference's prover is (one process, rayon threads,
 * Cargo.toml:36; `commit` is a plain method call, src/commitment.rs:78-87).  `bases` is a HOST pointer; device
 * devices[g] keeps the contiguous point range g of n_devices balanced ranges (first n % n_devices ranges one point
 * longer) and its own window tables (a device may be listed more than once; it then holds several ranges -- only
 * useful to exercise the sharded path on a box with fewer GPUs).  mira_msm_commit on the returned context cuts the host scalar vector the same
 * way (prefix semantics: range g commits v[lo_g .. min(hi_g, n))), drives every device from its own host thread (own
 * PCIe link, own streams), moves the 128-byte XYZZ partial sums to devices[0] by peer copy and adds and normalises
 * them there: the result is bit-identical to a single-device commit.  On such a context: mira_msm_ctx_len,
 * _check_on_curve, _prepare, _prepare_for (host vectors), mira_msm_commit, mira_msm_get_stats (pairs and launches
 * summed over the devices) and the mira_msm_set_* knobs (applied to every device) work as usual; entry points that
 * take device-resident vectors (commit_device, commit_batch, partial*, scalars_device) return MIRA_ERR_INVALID /
 * NULL -- a device vector lives on one GPU; row-sharded provers use one context per device (partial_batch_dev). */
int mira_msm_ctx_create_sharded(int curve, const void *bases, size_t n_bases, const int *devices, size_t n_devices,
                                mira_msm_ctx **out);
/* number of devices a context spans (1 for mira_msm_ctx_create) */
size_t mira_msm_ctx_num_devices(const mira_msm_ctx *ctx);
void mira_msm_ctx_destroy(mira_msm_ctx *ctx);
/* CommitmentKey::len (src/commitment.rs:44-46) */
size_t mira_msm_ctx_len(const mira_msm_ctx *ctx);
/* `key.par_iter().all(|p| p.is_on_curve())` (src/commitment.rs:145-146) evaluated on the GPU.
 * Returns MIRA_OK or MIRA_ERR_NOT_ON_CURVE. */
int mira_msm_ctx_check_on_curve(mira_msm_ctx *ctx);
/* Build (or fetch) the fixed-base table that commits of length `n` will use, so that the first timed
 * commit does not pay for it.  Optional.  The table cache is bounded: tables are only built while 6 GiB of device
 * memory stay free beside them, least-recently-used ones are evicted to make room, and a commit whose table cannot
 * be built falls back to the cached table covering n with the closest window (the result never depends on it). */
int mira_msm_ctx_prepare(mira_msm_ctx *ctx, size_t n);
/* The same for vectors that LOOK LIKE `scalars` (n x 32 B, host or device): the adaptive window (see
 * mira_msm_set_adaptive_window) is chosen from a sample of them, exactly as a commit of that vector would, and that
 * table is built -- mira_msm_ctx_prepare(n) builds the one uniform scalars use, which sparse witness columns never
 * touch. */
int mira_msm_ctx_prepare_for(mira_msm_ctx *ctx, const void *scalars, size_t n, int scalars_on_device);

/* ---- CommitmentKey::commit (src/commitment.rs:78-87) --------------------------------------------
 * out_affine (HOST, 64 B) = to_affine( sum_{i<n} scalars[i] * bases[i] ); identity -> 64 zero bytes.
 * n > len  => MIRA_ERR_TOO_LONG_INPUT, nothing written (checked before any arithmetic, as upstream).
 * `scalars` is a HOST pointer; the host->device copy is part of the call. */
int mira_msm_commit(mira_msm_ctx *ctx, const void *scalars, size_t n, void *out_affine);
/* The device copy of the scalars of the LAST successful mira_msm_commit / host-scalar mira_msm_partial on this context
 * (*n_out elements; NULL if none).  It stays valid until the next host-buffer commit on the context, so a caller that
 * committed a witness from host memory can feed the same bytes to mira_eval_rows* / mira_fold_w without a second
 * H2D copy (run_sps_protocol then commit_cross_terms then fold, src/nifs/vanilla/mod.rs:220-251). */
const void *mira_msm_scalars_device(const mira_msm_ctx *ctx, size_t *n_out);
/* Same, with `scalars` already resident on the context's device (e.g. produced by the cross-term
 * kernels); `stream` is a cudaStream_t, NULL = the legacy default stream -- the SAME convention as the witness
 * kernels below, so mira_fold_w(..., NULL) followed by mira_msm_commit_device(..., NULL) is ordered.  (This holds
 * for every entry point that reads device scalars: commit_device, commit_batch, partial with scalars_on_device,
 * partial_batch_dev.  Host-buffer commits run on the context's private stream.)  A fixed-base table built lazily by
 * the call is built on `stream` too.  The result is still returned to the host because every caller hashes it
 * (src/poseidon/poseidon_hash.rs:129-143). */
int mira_msm_commit_device(mira_msm_ctx *ctx, const void *scalars_dev, size_t n, void *out_affine, void *stream);

/* `cross_terms.iter().map(|v| ck.commit(v))` (src/nifs/vanilla/mod.rs:124-127) as ONE call: `count` (<= 32)
 * device vectors of the same length n against the same key.  scalars_dev is a HOST array of device pointers;
 * out_affine receives count x 64 B.  Each result is bit-identical to mira_msm_commit_device on that vector; the
 * vectors share one sort, one accumulation and one reduction launch sequence, which is what makes the 2^19-row
 * commitments of a fold step efficient.  n > len => MIRA_ERR_TOO_LONG_INPUT, nothing written. */
int mira_msm_commit_batch(mira_msm_ctx *ctx, const void *const *scalars_dev, size_t count, size_t n, void *out_affine,
                          void *stream);

/* ---- point-range sharding --------------------------------------------------------
 * A rank that owns bases [lo, hi) creates its context over that slice and calls *_partial with the
 * matching scalar slice; the result is the un-normalised partial sum as 128 B XYZZ
 * {X, Y, ZZ, ZZZ} (ZZ = 0 <=> identity) written to HOST memory (scalars_on_device selects where the
 * scalars live).  The partials of all ranks are gathered (NCCL all_gather of 128 B per rank in the
 * Python host layer) and folded by mira_msm_combine on one rank. */
int mira_msm_partial(mira_msm_ctx *ctx, const void *scalars, size_t n, int scalars_on_device, void *out_xyzz,
                     void *stream);
/* out_affine (HOST, 64 B) = to_affine( sum of `count` XYZZ partials (HOST, 128 B each) ), on `device`. */
int mira_msm_combine(int curve, const void *partials_xyzz, size_t count, int device, void *out_affine);
/* The same for a row-sharded prover that keeps everything on the device: the XYZZ partial sums of `count` (<= 32)
 * device vectors of length n against this rank's key shard, written to DEVICE memory (count x 128 B) on `stream`
 * with no host synchronisation (one vector: sampled window as in mira_msm_commit_device; several: one batched
 * pass as in mira_msm_commit_batch).  A rank queues all the commitments of a fold step into one buffer, gathers the
 * ranks' buffers with a single all_gather and calls mira_msm_combine_dev. */
int mira_msm_partial_batch_dev(mira_msm_ctx *ctx, const void *const *scalars_dev, size_t count, size_t n,
                               void *out_xyzz_dev, void *stream);
/* out_affine (HOST, n_commits x 64 B): out[j] = to_affine( sum over g < n_ranks of the XYZZ partial at
 * partials_dev + g * rank_stride + j * 128 ) -- the layout an all_gather of the per-rank buffers produces.
 * Runs on `device`/`stream` and returns when the results are in host memory. */
int mira_msm_combine_dev(int curve, const void *partials_dev, size_t n_ranks, size_t n_commits, size_t rank_stride,
                         int device, void *out_affine, void *stream);

/* Plain device memory, for callers that have no CUDA runtime binding of their own (the Rust shim of INTEGRATION.md
 * keeps fixed columns, witnesses and cross terms in HBM with these; Python callers use torch tensors instead).
 * upload is asynchronous on `stream` (NULL = the legacy default stream, like every witness-side call); download
 * returns when the bytes are in host memory; sync waits for `stream`. */
int mira_dev_alloc(int device, size_t bytes, void **out_dev);
int mira_dev_free(int device, void *dev_ptr);
int mira_dev_upload(int device, void *dst_dev, const void *src_host, size_t bytes, void *stream);
int mira_dev_download(int device, void *dst_host, const void *src_dev, size_t bytes, void *stream);
int mira_dev_sync(int device, void *stream);

/* Page-lock a host buffer the caller will commit from repeatedly (a witness column arena, the `Vec<C::Scalar>` a
 * prover re-uses every step).  mira_msm_commit works with any host memory, but only page-locked memory lets the
 * H2D copies of its slices run asynchronously at PCIe speed behind the accumulation of the previous slice; with
 * pageable memory the driver stages every copy through its own buffers, synchronously.  Thin wrappers over
 * cudaHostRegister / cudaHostUnregister; registering an already registered range is not an error. */
int mira_host_register(void *host_ptr, size_t bytes);
int mira_host_unregister(void *host_ptr);

/* ---- introspection used by bench.py / DESIGN.md roofline accounting ------------------------------ */
typedef struct {
  int window_bits;        /* c: signed window width of the last commit */
  int windows;            /* W = ceil(255 / c) */
  uint64_t entries;       /* n * W (point, window) pairs sorted and accumulated */
  uint64_t buckets;       /* 2^(c-1) */
  uint64_t kernel_launches; /* kernels launched by the last commit */
  float ms_digits, ms_sort, ms_accumulate, ms_reduce, ms_total; /* CUDA-event times of the last commit
                                                                     (only when profiling is enabled) */
} mira_msm_stats;
int mira_msm_get_stats(const mira_msm_ctx *ctx, mira_msm_stats *out);
/* enable (1) / disable (0) per-phase CUDA-event timing inside commit (adds synchronisation). */
int mira_msm_set_profiling(mira_msm_ctx *ctx, int enabled);
/* override the window width chosen by the size heuristic (0 = automatic). */
int mira_msm_set_window(mira_msm_ctx *ctx, int window_bits);
/* By default a commit of >= 2^18 scalars looks at a sample of them (8192 scalars, 64 chunks spread over the vector)
 * and picks the window width from the pair count the sample predicts: witness vectors are mostly zeros and small
 * values and want a much narrower window than uniform scalars.  0 = always use the size-only heuristic.  The result
 * never depends on the window. */
int mira_msm_set_adaptive_window(mira_msm_ctx *ctx, int enabled);
/* Host-buffer commits (mira_msm_commit, mira_msm_partial with host scalars) are pipelined: the vector is cut
 * into up to 4 slices of geometrically growing size (1 : 4 : 16 : 64), the smallest of at least
 * `min_scalars_per_slice` scalars (default 2^19), and slice k+1 crosses PCIe while slice k is accumulated into
 * the same bucket set.  0 disables slicing.  The result does not depend on it. */
int mira_msm_set_slice_min(mira_msm_ctx *ctx, size_t min_scalars_per_slice);
/* Slice pipeline: part k+1 of the scalar vector is decomposed into window digits and radix-sorted on a second stream
 * while part k is accumulated into the (shared) buckets.  Page-locked host-buffer commits always do this with their
 * H2D slices (mira_msm_set_slice_min); `slices` = 1 turns that off.  Device-resident commits are cut into `slices`
 * (2..16) equal parts of at least `min_scalars_per_slice` (0 = 2^20) scalars only when asked: measured on B200 the
 * slicing costs more than the overlap returns, so the default (0) leaves them whole.  The result does
 * not depend on any of it. */
int mira_msm_set_pipeline(mira_msm_ctx *ctx, int slices, size_t min_scalars_per_slice);
/* Experimental, off by default (0): before the XYZZ accumulation, add the entries of every bucket two by two in
 * AFFINE coordinates `levels` times (0..6), each level sharing its inversions by Montgomery's trick
 * (mira_b200/csrc/affine_levels.cuh).  5M + 1S per addition instead of 8M + 2S, but two passes over the gathered
 * points: measured neutral on B200, kept as a tested alternative.  The result does not depend on it.
 * levels = MIRA_AFFINE_THREAD_LOCAL_PAIRS selects round 2's variant instead: every accumulation thread adds the
 * entries 2k and 2k+1 of its own chunk in affine coordinates (one inversion per thread, mira_b200/csrc/inv30.cuh) and
 * feeds the sums to the XYZZ accumulation (k_pair_up / k_pair_add / k_pair_acc).  Measured slower on B200 (36.8 ms
 * against 31.8 ms of accumulation at 2^24 points: bound by the second gather of the table points). */
enum { MIRA_AFFINE_THREAD_LOCAL_PAIRS = -2 };
int mira_msm_set_affine_levels(mira_msm_ctx *ctx, int levels);

/* ==== field vectors in HBM: the witness side of the hot path =====
 * `field` is the SCALAR field of the curve being committed to: MIRA_FR for BN254 G1, MIRA_FQ for
 * Grumpkin G1.  Every *_dev pointer is device memory on `device`; elements are 32 B Montgomery form.
 * `stream` is a cudaStream_t (NULL = the legacy default stream).  All calls are asynchronous on that
 * stream unless noted, so their outputs can feed mira_msm_commit_device without touching the host. */

/* RelaxedPlonkWitness::fold, W part (src/plonk/mod.rs:1100-1110): out[i] = w1[i] + r * w2[i].
 * r_host: 32 B on the host.  out_dev may alias w1_dev. */
int mira_fold_w(int field, const void *w1_dev, const void *w2_dev, size_t n, const void *r_host, void *out_dev,
                int device, void *stream);
/* E part (src/plonk/mod.rs:1118-1131): out[i] = e[i] + sum_{k<n_terms} r^(k+1) * terms[k][i].
 * terms_dev: HOST array of n_terms device pointers (the cross-term vectors T_1..T_d).  n_terms <= 16. */
int mira_fold_e(int field, const void *e_dev, const void *const *terms_dev, size_t n_terms, size_t n,
                const void *r_host, void *out_dev, int device, void *stream);
/* util::concatenate_with_padding (src/util.rs:189-193), the W_i = concat(columns, 2^k) step of
 * run_sps_protocol_* (src/plonk/mod.rs:682-684): column c (lens[c] elements) followed by zeros up to
 * pad_size.  *out_len = elements written; fails with MIRA_ERR_INVALID if it exceeds out_capacity. */
int mira_concat_pad(const void *const *cols_dev, const size_t *lens, size_t n_cols, size_t pad_size, void *out_dev,
                    size_t out_capacity, size_t *out_len, int device, void *stream);

/* ---- GraphEvaluator (src/polynomial/graph_evaluator.rs:163-388) as a device row program ------------
 * A program is the serialised `GraphEvaluator { constants, rotations, num_intermediates, calculations }`.
 * `code` is a sequence of u32 words, one record per `CalculationInfo { calculation, target }` in order:
 *     word 0 : opcode | (n_operands << 8)     opcode: 0 Add(a,b) 1 Sub(a,b) 2 Mul(a,b) 3 Square(a)
 *     word 1 : target                                 4 Double(a) 5 Negate(a) 6 Horner(start, factor, parts..)
 *     then per operand (ValueSource):                 7 Store(a)
 *     word 0 : kind | (rotation_index << 8)   kind:   0 Constant 1 Intermediate 2 Fixed 3 Poly 4 Challenge
 *     word 1 : index
 * Horner's operands are ordered (start_value, factor, parts[0], parts[1], ...).  `constants` are n x 32 B
 * Montgomery elements (host), `rotations` the i32 rotation table (host).  The result of a row is the value
 * of the LAST calculation's target (ZERO for an empty program), as GraphEvaluator::evaluate returns. */
typedef struct mira_eval_program mira_eval_program;
int mira_eval_program_create(int field, const uint32_t *code, size_t code_words, const void *constants,
                             size_t n_constants, const int32_t *rotations, size_t n_rotations,
                             uint32_t num_intermediates, mira_eval_program **out);
void mira_eval_program_destroy(mira_eval_program *prog);

/* PlonkEvalDomain (src/plonk/eval.rs:93-106).  The pointer ARRAYS and `challenges` live on the host; the
 * columns they point to live on the device.  Poly{index} operands are resolved exactly as
 * GetDataForEval::eval_column_var (src/plonk/eval.rs:57-70: selectors, then fixed, then advice) and
 * PlonkEvalDomain::eval_advice_var (src/plonk/eval.rs:153-228: W1s / W2s, lookup sub-columns). */
typedef struct {
  uint64_t row_size;            /* GetDataForEval::row_size() */
  uint32_t num_selectors, num_fixed, num_advice, num_lookup, num_challenges, num_w1, num_w2;
  uint32_t flags;               /* MIRA_EVAL_LOOKUP_DOMAIN: see below */
  const void *const *selectors; /* [num_selectors] -> row_size bytes, Rust Vec<bool> image (0 / 1) */
  const void *const *fixed;     /* [num_fixed]     -> row_size x 32 B */
  const void *const *w1;        /* W1s[i] */
  const uint64_t *w1_len;       /* W1s[i].len() in elements */
  const void *const *w2;        /* W2s[i] */
  const uint64_t *w2_len;
  const void *challenges;       /* num_challenges x 32 B: U1.challenges | U1.u | U2.challenges | 1
                                   (src/nifs/vanilla/mod.rs:91) */
} mira_eval_domain;

/* flags bit 0: the domain is a LookupEvalDomain (src/plonk/eval.rs:84-135, used by evaluate_ls / evaluate_ts,
 * src/plonk/lookup.rs:212-276): advice variable `index` is the separate column w1[index] (w1_len[index] rows)
 * instead of a slice of the concatenated W; w2 is unused. */
enum { MIRA_EVAL_LOOKUP_DOMAIN = 1 };

/* out_dev[row] = evaluator.evaluate(&domain, row) for row in [0, row_size)
 * (the `(0..row_size).into_par_iter().map(..)` of src/nifs/vanilla/mod.rs:109-116).
 * Index errors the reference raises per row are raised here once, when the program is bound to the domain.
 * The linked device program is cached with the program object: binding the same program(s) to the same column
 * pointers again (a prover does that every step) only refreshes the challenges, asynchronously on `stream`.
 * Calls that share a program object must be issued on one stream. */
int mira_eval_rows(const mira_eval_program *prog, const mira_eval_domain *dom, void *out_dev, int device,
                   void *stream);
/* The same for rows [row_begin, row_end) only; out_dev holds row_end - row_begin elements.  This is the row-range
 * shard of SURVEY.md 8e: each rank evaluates its range (rotations still read the whole, read-only columns) and
 * commits it against its slice of the key; row_end > row_size is RowIndexOutOfBoundary. */
int mira_eval_rows_range(const mira_eval_program *prog, const mira_eval_domain *dom, uint64_t row_begin,
                         uint64_t row_end, void *out_dev, int device, void *stream);
/* All cross terms of a fold in ONE launch: `n_progs` (<= 16) programs over the same domain, outs_dev[k] receiving
 * program k's vector (HOST array of device pointers).  The programs are merged by value numbering — the reference
 * builds one GraphEvaluator per cross term (src/nifs/vanilla/mod.rs:100-121) but the terms share most of their
 * sub-products (63 % of the multiplications of the primary IVC circuit's six terms) — so a shared value is computed
 * once per row.  Field arithmetic is exact: every output is bit-identical to evaluating its program alone.
 * Programs whose targets are re-assigned cannot be merged (MIRA_ERR_EVAL_PROGRAM); evaluate those one by one. */
int mira_eval_rows_multi(const mira_eval_program *const *progs, size_t n_progs, const mira_eval_domain *dom,
                         uint64_t row_begin, uint64_t row_end, void *const *outs_dev, int device, void *stream);
typedef struct {
  uint32_t instructions;   /* device instructions after Store-forwarding and Horner expansion */
  uint32_t slots;          /* live intermediates kept per row (local memory) */
  uint32_t accesses;       /* distinct (column, rotation) loads */
  uint32_t uniforms;       /* constants + challenges */
  uint32_t muls, adds;     /* field multiplications (incl. squarings) / additive ops per row */
  uint32_t loads;          /* column loads executed per row */
  uint32_t fused;          /* a*b +- c*d pairs executed as one dual-product instruction (counted in muls/adds too) */
} mira_eval_stats;
/* statistics of the last binding that involved this program (of the MERGED program after a multi call) */
int mira_eval_program_stats(const mira_eval_program *prog, mira_eval_stats *out);

/* ---- lookup argument of the SPS rounds 2 / 3 (src/plonk/mod.rs:748-907, src/plonk/lookup.rs:278-319) ----
 * evaluate_m: out_m[i] = F::from(#{ j : l[j] == t[i] }) at the first occurrence of each distinct t value, ZERO at
 * later duplicates (the reference's `processed_t`).  Exact: equality is checked on the 32 bytes of the elements. */
int mira_lookup_m(int field, const void *l_dev, size_t n_l, const void *t_dev, size_t n_t, void *out_m_dev,
                  int device, void *stream);
/* evaluate_h_g: out_h[i] = 1 / (l[i] + r), out_g[i] = m[i] / (t[i] + r), with 0 for a zero denominator
 * (`invert().unwrap_or(ZERO)`).  r_host: 32 B on the host.  Batched inversion (Montgomery's trick) on the device. */
int mira_lookup_h_g(int field, const void *l_dev, const void *t_dev, const void *m_dev, size_t n, const void *r_host,
                    void *out_h_dev, void *out_g_dev, int device, void *stream);

/* ---- fft::best_fft (src/fft.rs:51-115): in-place radix-2 transform of 2^log_n elements ------------
 * omega_host: 32 B element of multiplicative order 2^log_n (host).  Output order and values are those of
 * the reference: bit-reversal permutation, then log_n butterfly stages with twiddles omega^i. */
int mira_fft(int field, void *a_dev, uint32_t log_n, const void *omega_host, int device, void *stream);
/* fft (inverse = 0, src/fft.rs:160-162) / ifft (inverse = 1, src/fft.rs:165-175, including the division
 * by 2^log_n): omega = get_omega_or_inv(log_n) derived on the device from PrimeField::ROOT_OF_UNITY.
 * Only MIRA_FR has a usable 2-adic subgroup (S = 28); log_n > S fails with MIRA_ERR_INVALID as the
 * reference's assert does. */
int mira_fft_std(int field, void *a_dev, uint32_t log_n, int inverse, int device, void *stream);

/* ---- device-side synthetic inputs and unit-test hooks ---------------------------------------------
 * Deterministic generators shared bit-for-bit with oracle/mira_oracle.c (oracle_gen_scalars /
 * oracle_gen_bases) so that bench.py can build 2^24..2^26-point keys on the GPU in seconds.
 * `out` is a DEVICE pointer on `device`.  dist: 0 uniform, 1 witness-like. */
int mira_gen_scalars(int curve, uint64_t seed, size_t first, size_t n, int dist, int device, void *out_dev);
int mira_gen_bases(int curve, uint64_t seed, size_t first, size_t n, int device, void *out_dev);

/* Element-wise field kernels over HOST arrays of n 32-byte elements (tests/test_field_gpu.py).
 * op: 0 mul, 1 add, 2 sub, 3 sqr(a), 4 inv(a), 5 to_canonical(a), 6 from_canonical(a) */
int mira_test_field_op(int field, int op, const void *a, const void *b, size_t n, int device, void *out);
/* Element-wise group kernels over HOST arrays of n 64-byte affine points.
 * op: 0 p+q via XYZZ mixed add, 1 p+q via XYZZ full add, 2 2p, 3 k*p with k = (uint32) first word of q */
int mira_test_point_op(int curve, int op, const void *p, const void *q, size_t n, int device, void *out);

/* Host-only unit-test hook: binds the programs to `dom` (pointers are only recorded, never dereferenced) and returns
 * the linked device program, so the CPU test-suite can check the linker (Store forwarding, Horner expansion, value
 * numbering across programs, product-pair fusion, liveness-based slot allocation) without a GPU.
 * instr_words: 4 u32 per device instruction (op | akind << 4 | bkind << 8 | dst << 16, a, b, c | ckind << 14 | d << 16 |
 * dkind << 30; kinds 0 slot, 1 uniform, 2 access; ops 0 add 1 sub 2 mul 3 square 4 double 5 negate 6 copy 7 a*b+c*d
 * 8 a*b-c*d 9 out[dst] = a); access_words: 3 u64 per access (column pointer, rotation as i64, is_selector);
 * uniform_bytes: n_uniforms x 32 B (constants, challenges, then ZERO). */
int mira_test_eval_link_multi(const mira_eval_program *const *progs, size_t n_progs, const mira_eval_domain *dom,
                              uint32_t *instr_words, size_t instr_cap, size_t *n_instr, uint64_t *access_words,
                              size_t access_cap, size_t *n_access, void *uniform_bytes, size_t uniform_cap,
                              size_t *n_uniforms, uint32_t *n_slots);

#ifdef __cplusplus
}
#endif
#endif
