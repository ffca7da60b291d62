// mira_witness.hpp — C++ host-side mirror of the reference's witness-side hot path over the C ABI
// (cross-term evaluation, witness folding, column concatenation, lookup coefficients, FFT).
//
// Mirrors, with the reference's names, argument meaning and error behaviour:
//   GraphEvaluator::evaluate over PlonkEvalDomain / LookupEvalDomain   /root/reference/src/polynomial/graph_evaluator.rs:361-388,
//                                                                       src/plonk/eval.rs:84-228, src/nifs/vanilla/mod.rs:100-121
//   RelaxedPlonkWitness::fold                                           src/plonk/mod.rs:1097-1134
//   util::concatenate_with_padding                                      src/util.rs:189-193
//   lookup::Arguments::{evaluate_m, evaluate_h_g}                       src/plonk/lookup.rs:278-319
//   fft::{best_fft, fft, ifft}                                          src/fft.rs:51-115,160-175
//
// All vectors are DEVICE pointers (32-byte Montgomery elements) on `device`; nothing here computes on the CPU.
// Header-only; link with -lmira_b200.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "mira_b200.h"
#include "mira_commitment.hpp"

namespace mira {

// plonk::eval::Error (src/plonk/eval.rs:3-25)
struct EvalError : std::runtime_error {
  enum Kind { ChallengeIndexOutOfBoundary, ColumnVariableIndexOutOfBoundary, RowIndexOutOfBoundary, InvalidWitnessIndex, InvalidProgram };
  Kind kind;
  EvalError(Kind k, const std::string& what) : std::runtime_error(what), kind(k) {}
};

namespace detail {
inline void check_witness(int rc) {
  if (rc == MIRA_OK) return;
  const char* msg = mira_last_error();
  switch (rc) {
    case MIRA_ERR_EVAL_CHALLENGE: throw EvalError(EvalError::ChallengeIndexOutOfBoundary, msg);
    case MIRA_ERR_EVAL_COLUMN: throw EvalError(EvalError::ColumnVariableIndexOutOfBoundary, msg);
    case MIRA_ERR_EVAL_ROW: throw EvalError(EvalError::RowIndexOutOfBoundary, msg);
    case MIRA_ERR_EVAL_WITNESS_INDEX: throw EvalError(EvalError::InvalidWitnessIndex, msg);
    case MIRA_ERR_EVAL_PROGRAM: throw EvalError(EvalError::InvalidProgram, msg);
    case MIRA_ERR_INVALID: throw std::invalid_argument(msg);
    default: throw CudaError(msg);
  }
}
}  // namespace detail

// PlonkEvalDomain (src/plonk/eval.rs:93-106) / LookupEvalDomain (:84-91): device columns, host challenges.
struct PlonkEvalDomain {
  size_t row_size = 0;
  size_t num_advice = 0, num_lookup = 0;
  std::vector<Scalar> challenges;                 // U1.challenges | U1.u | U2.challenges | 1  (src/nifs/vanilla/mod.rs:91)
  std::vector<const void*> selectors;             // device, row_size bytes each (Vec<bool> image)
  std::vector<const void*> fixed;                 // device, row_size elements each
  std::vector<const void*> W1s, W2s;              // device
  std::vector<uint64_t> W1_len, W2_len;           // W1s[i].len()
  bool lookup_domain = false;                     // LookupEvalDomain: W1s are the separate advice columns

  mira_eval_domain raw() const {
    mira_eval_domain d{};
    d.row_size = row_size;
    d.num_selectors = (uint32_t)selectors.size();
    d.num_fixed = (uint32_t)fixed.size();
    d.num_advice = (uint32_t)num_advice;
    d.num_lookup = (uint32_t)num_lookup;
    d.num_challenges = (uint32_t)challenges.size();
    d.num_w1 = (uint32_t)W1s.size();
    d.num_w2 = (uint32_t)W2s.size();
    d.flags = lookup_domain ? MIRA_EVAL_LOOKUP_DOMAIN : 0;
    d.selectors = selectors.data();
    d.fixed = fixed.data();
    d.w1 = W1s.data();
    d.w1_len = W1_len.data();
    d.w2 = W2s.data();
    d.w2_len = W2_len.data();
    d.challenges = challenges.data();
    return d;
  }
};

// A serialised GraphEvaluator (src/polynomial/graph_evaluator.rs:163-178); encoding in mira_b200.h.
class GraphEvaluator {
 public:
  GraphEvaluator(int field, const std::vector<uint32_t>& code, const std::vector<Scalar>& constants, const std::vector<int32_t>& rotations,
                 uint32_t num_intermediates) {
    detail::check_witness(mira_eval_program_create(field, code.data(), code.size(), constants.data(), constants.size(), rotations.data(),
                                                   rotations.size(), num_intermediates, &p_));
  }
  GraphEvaluator(const GraphEvaluator&) = delete;
  GraphEvaluator& operator=(const GraphEvaluator&) = delete;
  GraphEvaluator(GraphEvaluator&& o) noexcept : p_(o.p_) { o.p_ = nullptr; }
  ~GraphEvaluator() {
    if (p_) mira_eval_program_destroy(p_);
  }
  // `(0..row_size).into_par_iter().map(|row| evaluator.evaluate(&data, row))` (src/nifs/vanilla/mod.rs:109-116)
  void evaluate_rows(const PlonkEvalDomain& data, void* out_dev, int device = 0, void* stream = nullptr) const {
    mira_eval_domain d = data.raw();
    detail::check_witness(mira_eval_rows(p_, &d, out_dev, device, stream));
  }
  void evaluate_rows(const PlonkEvalDomain& data, uint64_t row_begin, uint64_t row_end, void* out_dev, int device = 0,
                     void* stream = nullptr) const {
    mira_eval_domain d = data.raw();
    detail::check_witness(mira_eval_rows_range(p_, &d, row_begin, row_end, out_dev, device, stream));
  }
  mira_eval_stats stats() const {
    mira_eval_stats s{};
    detail::check_witness(mira_eval_program_stats(p_, &s));
    return s;
  }
  const mira_eval_program* raw() const { return p_; }

 private:
  mira_eval_program* p_ = nullptr;
};

// All cross terms of one fold in a single launch (the loop of src/nifs/vanilla/mod.rs:100-121).
inline void evaluate_cross_terms(const std::vector<const GraphEvaluator*>& evaluators, const PlonkEvalDomain& data,
                                 const std::vector<void*>& outs_dev, int device = 0, void* stream = nullptr) {
  std::vector<const mira_eval_program*> raw;
  for (auto* e : evaluators) raw.push_back(e->raw());
  mira_eval_domain d = data.raw();
  detail::check_witness(mira_eval_rows_multi(raw.data(), raw.size(), &d, 0, data.row_size, outs_dev.data(), device, stream));
}

// RelaxedPlonkWitness::fold (src/plonk/mod.rs:1097-1134)
inline void fold_W(int field, const void* W1_dev, const void* W2_dev, size_t n, const Scalar& r, void* out_dev, int device = 0,
                   void* stream = nullptr) {
  detail::check_witness(mira_fold_w(field, W1_dev, W2_dev, n, &r, out_dev, device, stream));
}
inline void fold_E(int field, const void* E_dev, const std::vector<const void*>& cross_terms_dev, size_t n, const Scalar& r, void* out_dev,
                   int device = 0, void* stream = nullptr) {
  detail::check_witness(mira_fold_e(field, E_dev, cross_terms_dev.data(), cross_terms_dev.size(), n, &r, out_dev, device, stream));
}

// util::concatenate_with_padding (src/util.rs:189-193); returns the number of elements written
inline size_t concatenate_with_padding(const std::vector<const void*>& cols_dev, const std::vector<size_t>& lens, size_t pad_size,
                                       void* out_dev, size_t out_capacity, int device = 0, void* stream = nullptr) {
  size_t n = 0;
  detail::check_witness(mira_concat_pad(cols_dev.data(), lens.data(), cols_dev.size(), pad_size, out_dev, out_capacity, &n, device, stream));
  return n;
}

// lookup::Arguments::evaluate_m / evaluate_h_g (src/plonk/lookup.rs:278-319)
inline void evaluate_m(int field, const void* l_dev, size_t n_l, const void* t_dev, size_t n_t, void* m_dev, int device = 0,
                       void* stream = nullptr) {
  detail::check_witness(mira_lookup_m(field, l_dev, n_l, t_dev, n_t, m_dev, device, stream));
}
inline void evaluate_h_g(int field, const void* l_dev, const void* t_dev, const void* m_dev, size_t n, const Scalar& r, void* h_dev,
                         void* g_dev, int device = 0, void* stream = nullptr) {
  detail::check_witness(mira_lookup_h_g(field, l_dev, t_dev, m_dev, n, &r, h_dev, g_dev, device, stream));
}

// fft / ifft / best_fft (src/fft.rs:51-115,160-175), in place
inline void best_fft(int field, void* a_dev, const Scalar& omega, uint32_t log_n, int device = 0, void* stream = nullptr) {
  detail::check_witness(mira_fft(field, a_dev, log_n, &omega, device, stream));
}
inline void fft(int field, void* a_dev, uint32_t log_n, int device = 0, void* stream = nullptr) {
  detail::check_witness(mira_fft_std(field, a_dev, log_n, 0, device, stream));
}
inline void ifft(int field, void* a_dev, uint32_t log_n, int device = 0, void* stream = nullptr) {
  detail::check_witness(mira_fft_std(field, a_dev, log_n, 1, device, stream));
}

}  // namespace mira
