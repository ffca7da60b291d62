// Exercises the C++ host-side mirror of the reference's CommitmentKey (include/mira_commitment.hpp) the way a
// compiled caller would: same method names and error behaviour as /root/reference/src/commitment.rs:26-87.
// Test infrastructure: links the CPU oracle as the checker.
//   usage: commitment_mirror_test <n> [gpu]      without "gpu": only the paths that must work (or fail loudly) on a CPU box
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mira_commitment.hpp"
#include "mira_oracle.h"

template <class Curve>
static int run(int curve_id, size_t n, bool gpu) {
  std::vector<mira::Affine> bases(n);
  std::vector<mira::Scalar> scalars(n);
  oracle_gen_bases(curve_id, 0x4D495241, 0, n, 0, bases.data());
  oracle_gen_scalars(curve_id, 99, 0, n, 1, scalars.data());
  if (!gpu) {
    // no device: constructing the key must throw CudaError (there is no CPU fallback to fall into)
    try {
      mira::CommitmentKey<Curve> ck(bases.data(), n);
    } catch (const mira::CudaError& e) {
      std::printf("curve %d: no GPU -> CudaError: %s\n", curve_id, e.what());
      return 0;
    }
    std::printf("curve %d: expected CudaError without a GPU\n", curve_id);
    return 1;
  }
  mira::CommitmentKey<Curve> ck(bases.data(), n);
  ck.check_on_curve();
  if (ck.len() != n || ck.is_empty() || !mira::CommitmentKey<Curve>::default_value().is_identity()) return 2;
  mira::Affine want{};
  if (oracle_commit(curve_id, bases.data(), n, scalars.data(), n, 0, &want) != 0) return 3;
  mira::Affine got = ck.commit(scalars);
  if (!(got == want)) { std::printf("curve %d: commit differs from the oracle\n", curve_id); return 4; }
  // prefix of the key
  size_t m = n / 3 + 1;
  oracle_commit(curve_id, bases.data(), n, scalars.data(), m, 0, &want);
  if (!(ck.commit(scalars.data(), m) == want)) return 5;
  // zero vector -> identity == default_value()
  std::vector<mira::Scalar> zeros(n, mira::Scalar{});
  if (!ck.commit(zeros).is_identity()) return 6;
  // TooLongInput { input_len, limit }, checked before any arithmetic
  try {
    std::vector<mira::Scalar> too_long(n + 1, mira::Scalar{});
    ck.commit(too_long);
    return 7;
  } catch (const mira::TooLongInput& e) {
    if (e.input_len != n + 1 || e.limit != n) return 8;
    std::printf("curve %d: %s\n", curve_id, e.what());
  }
  // the same key as three point ranges driven from this one process (mira_msm_ctx_create_sharded; device 0 three times
  // on a one-GPU box): identical bytes, prefix included; device vectors are refused
  {
    mira::CommitmentKey<Curve> sharded(bases.data(), n, std::vector<int>{0, 0, 0});
    if (sharded.num_devices() != 3 || sharded.len() != n) return 9;
    if (!(sharded.commit(scalars) == got)) return 10;
    if (!(sharded.commit(scalars.data(), m) == want)) return 11;
    try {
      sharded.commit_device(scalars.data(), 1);
      return 12;
    } catch (const std::invalid_argument&) {
    }
  }
  std::printf("curve %d: C++ mirror ok (n = %zu)\n", curve_id, n);
  return 0;
}

int main(int argc, char** argv) {
  size_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1000;
  bool gpu = argc > 2 && !std::strcmp(argv[2], "gpu");
  int rc = run<mira::Bn256G1>(MIRA_BN254_G1, n, gpu);
  if (!rc) rc = run<mira::GrumpkinG1>(MIRA_GRUMPKIN_G1, n, gpu);
  return rc;
}
