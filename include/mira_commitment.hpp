// mira_commitment.hpp — C++ host-side mirror of the reference's `CommitmentKey<C>` over the C ABI.
//
// The reference is Rust (/root/reference/src/commitment.rs:26-167) and this image has no Rust
// toolchain, so this header stands where the Rust shim of INTEGRATION.md would: same method names,
// argument meaning and error behaviour, all arithmetic in libmira_b200.so (CUDA, sm_100a).
//
//   mira::CommitmentKey<mira::Bn256G1> ck(bases, n);      // CommitmentKey { ck: Box<[C]> }
//   auto c = ck.commit(scalars, m);                        // Result<C, Error>  ->  Affine or throws
//
// Header-only; link with -lmira_b200.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "mira_b200.h"

namespace mira {

// 4 x u64 little-endian limbs, Montgomery form: the in-memory layout of halo2curves' Fq / Fr.
struct alignas(8) Scalar {
  uint64_t limbs[4];
};
// {x, y}; (0, 0) is the identity (halo2curves G1Affine, `C::identity()`).
struct alignas(8) Affine {
  uint64_t x[4];
  uint64_t y[4];
  bool is_identity() const {
    uint64_t o = 0;
    for (int i = 0; i < 4; i++) o |= x[i] | y[i];
    return o == 0;
  }
  bool operator==(const Affine& r) const {
    for (int i = 0; i < 4; i++)
      if (x[i] != r.x[i] || y[i] != r.y[i]) return false;
    return true;
  }
};
static_assert(sizeof(Scalar) == 32 && sizeof(Affine) == 64, "layout must match size_of::<C>() == 64");

struct Bn256G1 { static constexpr int id = MIRA_BN254_G1; };       // halo2curves::bn256::G1Affine
struct GrumpkinG1 { static constexpr int id = MIRA_GRUMPKIN_G1; }; // halo2curves::grumpkin::G1Affine

// commitment::Error::TooLongInput { input_len, limit }  (src/commitment.rs:20-24)
struct TooLongInput : std::length_error {
  size_t input_len, limit;
  TooLongInput(size_t n, size_t lim)
      : std::length_error("Can't commit too long input: input len: " + std::to_string(n) + ", but limit is " + std::to_string(lim)),
        input_len(n), limit(lim) {}
};
// The reference has no device error path: a CUDA failure is fatal to the caller (no CPU fallback).
struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// io::ErrorKind::InvalidData "Wrong file in cache, some ptr out of curve" (src/commitment.rs:145-153)
struct NotOnCurve : std::runtime_error {
  using std::runtime_error::runtime_error;
};

template <class Curve>
class CommitmentKey {
 public:
  // `CommitmentKey { ck }`: uploads the generators once; they stay resident for the key's life.
  CommitmentKey(const Affine* bases, size_t n, int device = 0) : n_(n) {
    check(mira_msm_ctx_create(Curve::id, bases, n, /*bases_on_device=*/0, device, &ctx_));
  }
  // The same key spread over several GPUs of this process (mira_msm_ctx_create_sharded): contiguous point ranges on
  // `devices`; commit(host vector) runs on all of them and returns the same 64 bytes.  Device-vector methods throw
  // std::invalid_argument on such a key.
  CommitmentKey(const Affine* bases, size_t n, const std::vector<int>& devices) : n_(n) {
    check(mira_msm_ctx_create_sharded(Curve::id, bases, n, devices.data(), devices.size(), &ctx_));
  }
  size_t num_devices() const { return mira_msm_ctx_num_devices(ctx_); }
  CommitmentKey(const CommitmentKey&) = delete;
  CommitmentKey& operator=(const CommitmentKey&) = delete;
  CommitmentKey(CommitmentKey&& o) noexcept : ctx_(o.ctx_), n_(o.n_) { o.ctx_ = nullptr; }
  ~CommitmentKey() {
    if (ctx_) mira_msm_ctx_destroy(ctx_);
  }

  static Affine default_value() { return Affine{}; }           // src/commitment.rs:40-42
  size_t len() const { return mira_msm_ctx_len(ctx_); }        // :44-46
  bool is_empty() const { return len() == 0; }                 // :48-50

  // src/commitment.rs:78-87.  The length check happens before any arithmetic, as upstream.
  Affine commit(const Scalar* v, size_t n) const {
    if (n > n_) throw TooLongInput(n, n_);
    Affine out{};
    check(mira_msm_commit(ctx_, v, n, &out), n);
    return out;
  }
  Affine commit(const std::vector<Scalar>& v) const { return commit(v.data(), v.size()); }
  // scalars already in HBM on the key's device (produced by mira_eval_cross_terms / mira_fold_witness)
  Affine commit_device(const void* scalars_dev, size_t n, void* stream = nullptr) const {
    if (n > n_) throw TooLongInput(n, n_);
    Affine out{};
    check(mira_msm_commit_device(ctx_, scalars_dev, n, &out, stream), n);
    return out;
  }

  // `cross_terms.iter().map(|v| ck.commit(v))` (src/nifs/vanilla/mod.rs:124-127) in one call: device vectors of length n
  std::vector<Affine> commit_batch(const std::vector<const void*>& scalars_dev, size_t n, void* stream = nullptr) const {
    if (n > n_) throw TooLongInput(n, n_);
    std::vector<Affine> out(scalars_dev.size());
    check(mira_msm_commit_batch(ctx_, scalars_dev.data(), scalars_dev.size(), n, out.data(), stream), n);
    return out;
  }
  // Row-/point-range-sharded provers (one process per GPU): this rank's XYZZ partial sums (128 B each) of the vectors,
  // written to device memory on `stream` without a host round trip; all_gather the ranks' buffers, then
  // combine_partials_device.
  void partial_batch_device(const std::vector<const void*>& scalars_dev, size_t n, void* out_xyzz_dev, void* stream = nullptr) const {
    if (n > n_) throw TooLongInput(n, n_);
    check(mira_msm_partial_batch_dev(ctx_, scalars_dev.data(), scalars_dev.size(), n, out_xyzz_dev, stream), n);
  }

  // src/commitment.rs:109-124: the file is the memory image of [C], 64 << k bytes.
  static CommitmentKey load_from_file(const char* path, unsigned k, int device = 0) {
    size_t n = size_t(1) << k;
    std::vector<Affine> buf(n);
    FILE* f = std::fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    size_t got = std::fread(buf.data(), sizeof(Affine), n, f);
    std::fclose(f);
    if (got != n) throw std::runtime_error(std::string("short read (read_exact) from ") + path);
    return CommitmentKey(buf.data(), n, device);
  }
  // `key.par_iter().all(|p| p.is_on_curve())` (src/commitment.rs:145-146), evaluated on the GPU.
  void check_on_curve() const { check(mira_msm_ctx_check_on_curve(ctx_)); }
  // build the fixed-base table for commits of length n ahead of time (no reference analogue)
  void prepare(size_t n) const { check(mira_msm_ctx_prepare(ctx_, n)); }

  mira_msm_ctx* raw() const { return ctx_; }

 private:
  void check(int rc, size_t n = 0) const {
    if (rc == MIRA_OK) return;
    if (rc == MIRA_ERR_TOO_LONG_INPUT) throw TooLongInput(n, n_);
    if (rc == MIRA_ERR_NOT_ON_CURVE) throw NotOnCurve(mira_last_error());
    if (rc == MIRA_ERR_INVALID) throw std::invalid_argument(mira_last_error());
    throw CudaError(mira_last_error());
  }
  mira_msm_ctx* ctx_ = nullptr;
  size_t n_ = 0;
};

// out[j] = to_affine(sum over ranks of the partial at partials_dev + g * rank_stride + j * 128): the layout an all_gather
// of the per-rank buffers of partial_batch_device leaves on the device.
template <class Curve>
inline std::vector<Affine> combine_partials_device(const void* partials_dev, size_t n_ranks, size_t n_commits, size_t rank_stride,
                                                   int device = 0, void* stream = nullptr) {
  std::vector<Affine> out(n_commits);
  int rc = mira_msm_combine_dev(Curve::id, partials_dev, n_ranks, n_commits, rank_stride, device, out.data(), stream);
  if (rc == MIRA_ERR_INVALID) throw std::invalid_argument(mira_last_error());
  if (rc != MIRA_OK) throw CudaError(mira_last_error());
  return out;
}

}  // namespace mira
