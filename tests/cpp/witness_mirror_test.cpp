// Exercises the C++ mirror of the witness-side interfaces (include/mira_witness.hpp) against the CPU oracle.
// Test infrastructure (links oracle/).  Built with nvcc for the CUDA runtime (device buffers only; no kernels here).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "mira_oracle.h"
#include "mira_witness.hpp"

static void* to_dev(const void* h, size_t bytes) {
  void* d = nullptr;
  cudaMalloc(&d, bytes ? bytes : 32);
  if (bytes) cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
  return d;
}
static std::vector<mira::Scalar> from_dev(const void* d, size_t n) {
  std::vector<mira::Scalar> h(n);
  cudaMemcpy(h.data(), d, n * 32, cudaMemcpyDeviceToHost);
  return h;
}
static bool same(const std::vector<mira::Scalar>& a, const std::vector<mira::Scalar>& b) {
  return a.size() == b.size() && !std::memcmp(a.data(), b.data(), a.size() * 32);
}

int main() {
  const int F = MIRA_FR;
  const size_t rows = 1000;
  // domain: 1 selector, 2 fixed, 2 advice columns per instance, 1 challenge
  std::vector<uint8_t> sel(rows);
  for (size_t i = 0; i < rows; i++) sel[i] = (i * 7 + 1) % 3 == 0;
  std::vector<mira::Scalar> f0(rows), f1(rows), w1(2 * rows), w2(2 * rows), ch(1), r(1);
  oracle_gen_scalars(ORACLE_BN254_G1, 1, 0, rows, 1, f0.data());
  oracle_gen_scalars(ORACLE_BN254_G1, 2, 0, rows, 0, f1.data());
  oracle_gen_scalars(ORACLE_BN254_G1, 3, 0, 2 * rows, 0, w1.data());
  oracle_gen_scalars(ORACLE_BN254_G1, 4, 0, 2 * rows, 1, w2.data());
  oracle_gen_scalars(ORACLE_BN254_G1, 5, 0, 1, 0, ch.data());
  oracle_gen_scalars(ORACLE_BN254_G1, 6, 0, 1, 0, r.data());
  // program (GraphEvaluator encoding):  t0 = Store(Poly 3 = advice 0 of instance 1)   t1 = Store(Poly 6 = advice 1 of instance 2, rot +1)
  //   t2 = Store(Challenge 0)  t3 = Mul(t0, t1)  t4 = Store(Poly 1 = fixed 0)  t5 = Mul(t4, t2)  t6 = Add(t3, t5)  t7 = Store(Poly 0 = selector)
  //   t8 = Mul(t6, t7)  t9 = Store(t8)
  auto OP = [](uint32_t op, uint32_t n) { return op | (n << 8); };
  auto VS = [](uint32_t kind, uint32_t rot) { return kind | (rot << 8); };
  std::vector<uint32_t> code = {
      OP(7, 1), 0, VS(3, 0), 3,  OP(7, 1), 1, VS(3, 1), 6,  OP(7, 1), 2, VS(4, 0), 0,
      OP(2, 2), 3, VS(1, 0), 0, VS(1, 0), 1,  OP(7, 1), 4, VS(3, 0), 1,  OP(2, 2), 5, VS(1, 0), 4, VS(1, 0), 2,
      OP(0, 2), 6, VS(1, 0), 3, VS(1, 0), 5,  OP(7, 1), 7, VS(3, 0), 0,  OP(2, 2), 8, VS(1, 0), 6, VS(1, 0), 7,
      OP(7, 1), 9, VS(1, 0), 8};
  std::vector<int32_t> rotations = {0, 1};
  std::vector<mira::Scalar> constants(3);
  oracle_fe_from_u64(ORACLE_FR, 0, &constants[0]);
  oracle_fe_from_u64(ORACLE_FR, 1, &constants[1]);
  oracle_fe_from_u64(ORACLE_FR, 2, &constants[2]);

  // oracle side
  const void* h_sel[1] = {sel.data()};
  const void* h_fix[2] = {f0.data(), f1.data()};
  const void* h_w1[1] = {w1.data()};
  const void* h_w2[1] = {w2.data()};
  uint64_t wl[1] = {2 * rows};
  oracle_eval_domain od{};
  od.row_size = rows; od.num_selectors = 1; od.num_fixed = 2; od.num_advice = 2; od.num_lookup = 0; od.num_challenges = 1;
  od.num_w1 = 1; od.num_w2 = 1; od.selectors = h_sel; od.fixed = h_fix; od.w1 = h_w1; od.w1_len = wl; od.w2 = h_w2; od.w2_len = wl;
  od.challenges = ch.data();
  std::vector<mira::Scalar> want(rows);
  if (oracle_eval_rows(ORACLE_FR, code.data(), code.size(), constants.data(), 3, rotations.data(), 2, 10, &od, 0, rows, want.data())) return 1;

  // GPU side through the C++ mirror
  mira::PlonkEvalDomain dom;
  dom.row_size = rows; dom.num_advice = 2; dom.num_lookup = 0; dom.challenges = ch;
  dom.selectors = {to_dev(sel.data(), rows)};
  dom.fixed = {to_dev(f0.data(), rows * 32), to_dev(f1.data(), rows * 32)};
  dom.W1s = {to_dev(w1.data(), 2 * rows * 32)};
  dom.W2s = {to_dev(w2.data(), 2 * rows * 32)};
  dom.W1_len = {2 * rows};
  dom.W2_len = {2 * rows};
  void* out = to_dev(nullptr, rows * 32);
  mira::GraphEvaluator ev(F, code, constants, rotations, 10);
  ev.evaluate_rows(dom, out);
  cudaDeviceSynchronize();
  if (!same(from_dev(out, rows), want)) { std::printf("evaluate_rows differs from the oracle\n"); return 2; }
  mira_eval_stats st = ev.stats();
  std::printf("evaluate_rows ok: %u device instructions, %u slots, %u fused\n", st.instructions, st.slots, st.fused);
  // error behaviour: challenge index out of boundary
  try {
    std::vector<uint32_t> bad = {OP(7, 1), 0, VS(4, 0), 5};
    mira::GraphEvaluator(F, bad, constants, rotations, 1).evaluate_rows(dom, out);
    return 3;
  } catch (const mira::EvalError& e) {
    if (e.kind != mira::EvalError::ChallengeIndexOutOfBoundary) return 4;
    std::printf("EvalError: %s\n", e.what());
  }
  // fold
  std::vector<mira::Scalar> fw(2 * rows);
  oracle_fold_w(ORACLE_FR, w1.data(), w2.data(), 2 * rows, r.data(), fw.data());
  void* dfw = to_dev(nullptr, 2 * rows * 32);
  mira::fold_W(F, dom.W1s[0], dom.W2s[0], 2 * rows, r[0], dfw);
  cudaDeviceSynchronize();
  if (!same(from_dev(dfw, 2 * rows), fw)) { std::printf("fold_W differs\n"); return 5;}
  const void* terms_h[2] = {f0.data(), f1.data()};
  std::vector<mira::Scalar> fe(rows);
  oracle_fold_e(ORACLE_FR, want.data(), terms_h, 2, rows, r.data(), fe.data());
  void* dfe = to_dev(nullptr, rows * 32);
  mira::fold_E(F, out, {dom.fixed[0], dom.fixed[1]}, rows, r[0], dfe);
  cudaDeviceSynchronize();
  if (!same(from_dev(dfe, rows), fe)) { std::printf("fold_E differs\n"); return 6; }
  // fft round trip + oracle
  std::vector<mira::Scalar> a(1 << 10);
  oracle_gen_scalars(ORACLE_BN254_G1, 9, 0, a.size(), 0, a.data());
  void* da = to_dev(a.data(), a.size() * 32);
  mira::fft(F, da, 10);
  cudaDeviceSynchronize();
  std::vector<mira::Scalar> fa = a;
  oracle_fft_forward(ORACLE_FR, fa.data(), 10);
  if (!same(from_dev(da, a.size()), fa)) { std::printf("fft differs\n"); return 7; }
  mira::ifft(F, da, 10);
  cudaDeviceSynchronize();
  if (!same(from_dev(da, a.size()), a)) { std::printf("ifft round trip differs\n"); return 8; }
  std::printf("C++ witness mirror ok\n");
  return 0;
}
