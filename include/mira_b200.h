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
  MIRA_ERR_NOT_ON_CURVE = -4    /* load_or_setup_cache's check, src/commitment.rs:145-153 */
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
void mira_msm_ctx_destroy(mira_msm_ctx *ctx);
/* CommitmentKey::len (src/commitment.rs:44-46) */
size_t mira_msm_ctx_len(const mira_msm_ctx *ctx);
/* `key.par_iter().all(|p| p.is_on_curve())` (src/commitment.rs:145-146) evaluated on the GPU.
 * Returns MIRA_OK or MIRA_ERR_NOT_ON_CURVE. */
int mira_msm_ctx_check_on_curve(mira_msm_ctx *ctx);
/* Build (or fetch) the fixed-base table that commits of length `n` will use, so that the first timed
 * commit does not pay for it.  Optional. */
int mira_msm_ctx_prepare(mira_msm_ctx *ctx, size_t n);

/* ---- CommitmentKey::commit (src/commitment.rs:78-87) --------------------------------------------
 * out_affine (HOST, 64 B) = to_affine( sum_{i<n} scalars[i] * bases[i] ); identity -> 64 zero bytes.
 * n > len  => MIRA_ERR_TOO_LONG_INPUT, nothing written (checked before any arithmetic, as upstream).
 * `scalars` is a HOST pointer; the host->device copy is part of the call. */
int mira_msm_commit(mira_msm_ctx *ctx, const void *scalars, size_t n, void *out_affine);
/* Same, with `scalars` already resident on the context's device (e.g. produced by the cross-term
 * kernels); `stream` is a cudaStream_t (NULL = the context's own stream).  The result is still
 * returned to the host because every caller hashes it (src/poseidon/poseidon_hash.rs:129-143). */
int mira_msm_commit_device(mira_msm_ctx *ctx, const void *scalars_dev, size_t n, void *out_affine, void *stream);

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

#ifdef __cplusplus
}
#endif
#endif
