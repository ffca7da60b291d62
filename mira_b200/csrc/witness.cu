// witness.cu — the witness side of the hot path in HBM: cross-term row evaluator, fold AXPY, column
// concatenation and radix-2 FFT, plus their C ABI (include/mira_b200.h, "field vectors in HBM").
//
// Replaces, on the device:
//   GraphEvaluator::evaluate under into_par_iter   /root/reference/src/nifs/vanilla/mod.rs:100-121,
//                                                   src/polynomial/graph_evaluator.rs:93-149,361-388,
//                                                   src/plonk/eval.rs:57-70,153-228
//   RelaxedPlonkWitness::fold                       src/plonk/mod.rs:1097-1134
//   concatenate_with_padding                        src/util.rs:189-193
//   best_fft / fft / ifft                           src/fft.rs:12-27,51-115,160-175
// Outputs stay in HBM so they feed mira_msm_commit_device without crossing PCIe.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <map>
#include <string>
#include <tuple>

#include "ctx.hpp"
#include "field.cuh"

namespace mira {

// ------------------------------------------------------------------ fold
// One thread per element, 128-bit loads/stores; HBM-bound: 96 B per element for W (two reads, one write),
// (n_terms + 2) * 32 B per element for E.
template <class F>
__global__ void __launch_bounds__(256) k_fold_w(const void* __restrict__ w1, const void* __restrict__ w2, size_t n, Fe<F> r,
                                                void* __restrict__ out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Fe<F> a = fe_load<F>(reinterpret_cast<const char*>(w1) + i * 32);
    Fe<F> b = fe_load<F>(reinterpret_cast<const char*>(w2) + i * 32);
    fe_store<F>(reinterpret_cast<char*>(out) + i * 32, fe_add(a, fe_mul(r, b)));
  }
}

constexpr int MAX_FOLD_TERMS = 16;
struct FoldTerms {
  const void* t[MAX_FOLD_TERMS];
};
// powers r^1..r^n_terms are computed once per block into shared memory (n_terms sequential products)
template <class F>
__global__ void __launch_bounds__(256) k_fold_e(const void* __restrict__ e, FoldTerms terms, int n_terms, size_t n, Fe<F> r,
                                                void* __restrict__ out) {
  __shared__ uint32_t pw[MAX_FOLD_TERMS][8];
  if (threadIdx.x == 0) {
    Fe<F> p = r;
    for (int k = 0; k < n_terms; k++) {
      for (int j = 0; j < 8; j++) pw[k][j] = p.v[j];
      p = fe_mul(p, r);
    }
  }
  __syncthreads();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Fe<F> acc = fe_load<F>(reinterpret_cast<const char*>(e) + i * 32);
    for (int k = 0; k < n_terms; k++) {
      Fe<F> p;
#pragma unroll
      for (int j = 0; j < 8; j++) p.v[j] = pw[k][j];
      Fe<F> t = fe_load<F>(reinterpret_cast<const char*>(terms.t[k]) + i * 32);
      acc = fe_add(acc, fe_mul(p, t));
    }
    fe_store<F>(reinterpret_cast<char*>(out) + i * 32, acc);
  }
}

// ------------------------------------------------------------------ row evaluator
// Device instruction (16 B): x = op | akind << 4 | bkind << 8 | dst << 16, y = a, z = b,
// w = c | ckind << 14 | d << 16 | dkind << 30 for the fused sum / difference of two products a*b +- c*d.
enum : uint32_t { DOP_ADD = 0, DOP_SUB, DOP_MUL, DOP_SQUARE, DOP_DOUBLE, DOP_NEGATE, DOP_COPY, DOP_MUL2ADD, DOP_MUL2SUB, DOP_OUT };
// DOP_OUT: out[dst][row] = a  (dst = index of the output vector: several programs may be merged into one)
enum : uint32_t { DK_SLOT = 0, DK_UNIFORM = 1, DK_ACCESS = 2 };
struct Access {          // one distinct (column, rotation) load
  const void* ptr;       // column base (32 B elements, or 1 B selectors)
  int32_t rot;
  uint32_t is_selector;
};

struct EvalOutsDev {
  void* p[16];
};

template <class F, int S>
__device__ __forceinline__ Fe<F> ev_fetch(uint32_t kind, uint32_t idx, const Fe<F>* slots, const void* __restrict__ uniforms,
                                          const Access* __restrict__ acc, uint64_t row, uint64_t row_size) {
  if (kind == DK_SLOT) return slots[idx];
  if (kind == DK_UNIFORM) return fe_load<F>(reinterpret_cast<const char*>(uniforms) + (size_t)idx * 32);
  Access a = acc[idx];
  uint64_t r = row;
  if (a.rot) {                       // get_rotation_idx: (row + rot).rem_euclid(row_size)
    int64_t v = ((int64_t)row + a.rot) % (int64_t)row_size;
    r = (uint64_t)(v < 0 ? v + (int64_t)row_size : v);
  }
  if (a.is_selector) return reinterpret_cast<const uint8_t*>(a.ptr)[r] ? fe_one<F>() : fe_zero<F>();
  return fe_load<F>(reinterpret_cast<const char*>(a.ptr) + r * 32);
}

// One thread per row (grid-stride).  The linked program sits in shared memory and is warp-uniform, so the
// opcode switch never diverges; live intermediates sit in a per-thread local array (L1-resident for the slot
// counts the linker produces).  Integer-pipe bound: ~1 Montgomery product per MUL/SQUARE instruction.
template <class F, int S>
__global__ void __launch_bounds__(128) k_eval_rows(const uint4* __restrict__ prog, uint32_t n_instr, const void* __restrict__ uniforms,
                                                   const Access* __restrict__ acc, uint64_t row_size, uint64_t row_begin,
                                                   uint64_t row_end, EvalOutsDev outs) {
  extern __shared__ uint4 s_prog[];
  for (uint32_t i = threadIdx.x; i < n_instr; i += blockDim.x) s_prog[i] = prog[i];
  __syncthreads();
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t row = row_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < row_end; row += stride) {
    Fe<F> slots[S];
    for (uint32_t pc = 0; pc < n_instr; pc++) {
      uint4 ins = s_prog[pc];
      uint32_t op = ins.x & 0xf, ak = (ins.x >> 4) & 0xf, bk = (ins.x >> 8) & 0xf, dst = ins.x >> 16;
      Fe<F> a = ev_fetch<F, S>(ak, ins.y, slots, uniforms, acc, row, row_size);
      if (op == DOP_OUT) {
        fe_store<F>(reinterpret_cast<char*>(outs.p[dst]) + (row - row_begin) * 32, a);
        continue;
      }
      Fe<F> r;
      if (op <= DOP_MUL) {
        Fe<F> b = ev_fetch<F, S>(bk, ins.z, slots, uniforms, acc, row, row_size);
        if (op == DOP_MUL) r = fe_mul(a, b);
        else if (op == DOP_ADD) r = fe_add(a, b);
        else r = fe_sub(a, b);
      } else if (op == DOP_MUL2ADD || op == DOP_MUL2SUB) {        // a*b +- c*d under ONE Montgomery reduction (field.cuh: mont_mul2)
        Fe<F> b = ev_fetch<F, S>(bk, ins.z, slots, uniforms, acc, row, row_size);
        Fe<F> c = ev_fetch<F, S>((ins.w >> 14) & 3u, ins.w & 0x3fffu, slots, uniforms, acc, row, row_size);
        Fe<F> d = ev_fetch<F, S>(ins.w >> 30, (ins.w >> 16) & 0x3fffu, slots, uniforms, acc, row, row_size);
        r = op == DOP_MUL2ADD ? fe_mul_add_mul(a, b, c, d) : fe_mul_sub_mul(a, b, c, d);
      } else if (op == DOP_SQUARE) r = fe_sqr(a);
      else if (op == DOP_DOUBLE) r = fe_dbl(a);
      else if (op == DOP_NEGATE) r = fe_neg(a);
      else r = a;
      slots[dst] = r;
    }
  }
}

// ------------------------------------------------------------------ FFT
// consts[0] = omega, consts[1] = scale (ifft divisor) — derived on the device for the fft/ifft wrappers.
template <class F>
__global__ void k_fft_consts(uint32_t k, int inverse, void* consts) {
  if (threadIdx.x || blockIdx.x) return;
  // PrimeField::ROOT_OF_UNITY of bn256::Fr, S = 28 (7^((r-1)/2^28)); validated by src/fft.rs:239-258
  Fe<F> c;
  const uint32_t root[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
  for (int i = 0; i < 8; i++) c.v[i] = root[i];
  Fe<F> w = fe_from_canonical(c);
  if (inverse) w = fe_inv(w);
  for (uint32_t i = k; i < 28; i++) w = fe_sqr(w);
  Fe<F> two = fe_add(fe_one<F>(), fe_one<F>());
  Fe<F> inv2 = fe_inv(two), d = fe_one<F>();
  for (uint32_t i = 0; i < k; i++) d = fe_mul(d, inv2);
  fe_store<F>(consts, w);
  fe_store<F>(reinterpret_cast<char*>(consts) + 32, d);
}

// tw[i] = omega^(i * stride), i < count (square-and-multiply per thread: exact, so equal to the reference's running
// product).  Used directly for small transforms and for the two small tables of the large ones.
template <class F>
__global__ void __launch_bounds__(256) k_fft_twiddles(const void* __restrict__ consts, size_t count, size_t stride, void* __restrict__ tw) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fe<F> base = fe_load<F>(consts), acc = fe_one<F>();
  for (size_t e = i * stride; e; e >>= 1) {
    if (e & 1) acc = fe_mul(acc, base);
    base = fe_sqr(base);
  }
  fe_store<F>(reinterpret_cast<char*>(tw) + i * 32, acc);
}

// Large tables: tw[i] = hi[i >> FFT_TW_LOW] * tw[i & (2^FFT_TW_LOW - 1)] for i >= 2^FFT_TW_LOW -- one multiplication
// per twiddle instead of a ~2 log2(i)-multiplication power (which cost as much as the transform itself at 2^24).
constexpr int FFT_TW_LOW = 10;
template <class F>
__global__ void __launch_bounds__(256) k_fft_twiddles_combine(const void* __restrict__ hi, size_t half, void* __restrict__ tw) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + ((size_t)1 << FFT_TW_LOW);
  if (i >= half) return;
  const char* t = reinterpret_cast<const char*>(tw);
  Fe<F> h = fe_load<F>(reinterpret_cast<const char*>(hi) + (i >> FFT_TW_LOW) * 32);
  Fe<F> l = fe_load<F>(t + (i & (((size_t)1 << FFT_TW_LOW) - 1)) * 32);
  fe_store<F>(reinterpret_cast<char*>(tw) + i * 32, fe_mul(h, l));
}

template <class F>
__global__ void __launch_bounds__(256) k_fft_bitrev(void* a, uint32_t log_n) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ((size_t)1 << log_n)) return;
  size_t rk = (size_t)(__brevll((unsigned long long)k) >> (64 - log_n));
  if (log_n == 0) rk = 0;
  if (k < rk) {
    char* p = reinterpret_cast<char*>(a);
    const uint4* pk = reinterpret_cast<const uint4*>(p + k * 32);
    const uint4* pr = reinterpret_cast<const uint4*>(p + rk * 32);
    uint4 k0 = pk[0], k1 = pk[1], r0 = pr[0], r1 = pr[1];
    reinterpret_cast<uint4*>(p + k * 32)[0] = r0; reinterpret_cast<uint4*>(p + k * 32)[1] = r1;
    reinterpret_cast<uint4*>(p + rk * 32)[0] = k0; reinterpret_cast<uint4*>(p + rk * 32)[1] = k1;
  }
}

// Stages [0, stages) fused in shared memory on tiles of 2^stages consecutive (already bit-reversed) elements.
constexpr int FFT_TILE_LOG = 10;
template <class F>
__global__ void __launch_bounds__(512) k_fft_tile(void* a, uint32_t log_n, uint32_t stages, const void* __restrict__ tw) {
  extern __shared__ uint4 sm_fft[];
  const size_t tile = (size_t)1 << stages, base = (size_t)blockIdx.x * tile;
  char* g = reinterpret_cast<char*>(a) + base * 32;
  for (size_t i = threadIdx.x; i < tile * 2; i += blockDim.x) sm_fft[i] = reinterpret_cast<const uint4*>(g)[i];
  __syncthreads();
  const size_t n = (size_t)1 << log_n;
  for (uint32_t s = 0; s < stages; s++) {
    size_t half = (size_t)1 << s, twiddle_chunk = n >> (s + 1);
    for (size_t t = threadIdx.x; t < tile / 2; t += blockDim.x) {
      size_t i = t & (half - 1), left = ((t >> s) << (s + 1)) + i, right = left + half;
      Fe<F> l, r;
      {
        uint4 lo = sm_fft[2 * left], hi = sm_fft[2 * left + 1];
        l.v[0] = lo.x; l.v[1] = lo.y; l.v[2] = lo.z; l.v[3] = lo.w; l.v[4] = hi.x; l.v[5] = hi.y; l.v[6] = hi.z; l.v[7] = hi.w;
        lo = sm_fft[2 * right]; hi = sm_fft[2 * right + 1];
        r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w; r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
      }
      if (i) r = fe_mul(r, fe_load<F>(reinterpret_cast<const char*>(tw) + (i * twiddle_chunk) * 32));
      Fe<F> x = fe_add(l, r), y = fe_sub(l, r);
      sm_fft[2 * left] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]); sm_fft[2 * left + 1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
      sm_fft[2 * right] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]); sm_fft[2 * right + 1] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
    }
    __syncthreads();
  }
  for (size_t i = threadIdx.x; i < tile * 2; i += blockDim.x) reinterpret_cast<uint4*>(g)[i] = sm_fft[i];
}

// One butterfly stage straight on HBM (stages >= FFT_TILE_LOG): 64 B read + 64 B written per butterfly.
template <class F>
__global__ void __launch_bounds__(256) k_fft_stage(void* a, uint32_t log_n, uint32_t s, const void* __restrict__ tw) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)1 << log_n;
  if (t >= n / 2) return;
  size_t half = (size_t)1 << s, i = t & (half - 1), left = ((t >> s) << (s + 1)) + i, right = left + half;
  char* p = reinterpret_cast<char*>(a);
  Fe<F> l = fe_load<F>(p + left * 32), r = fe_load<F>(p + right * 32);
  if (i) r = fe_mul(r, fe_load<F>(reinterpret_cast<const char*>(tw) + (i * (n >> (s + 1))) * 32));
  fe_store<F>(p + left * 32, fe_add(l, r));
  fe_store<F>(p + right * 32, fe_sub(l, r));
}

// Stages [s0, s0 + R) fused in shared memory for the stages beyond the contiguous tile: a block takes 2^R rows that
// are 2^s0 apart and FFT_COLS consecutive columns (8 x 32 B = 256 B contiguous per row, so the strided gather is
// still made of full sectors), runs the R butterfly stages over the row dimension and writes back.  One HBM round
// trip for up to 7 stages instead of one per stage.
constexpr int FFT_COLS = 8, FFT_STRIDED_LOG = 7;
template <class F>
__global__ void __launch_bounds__(256) k_fft_strided(void* a, uint32_t log_n, uint32_t s0, uint32_t R, const void* __restrict__ tw) {
  extern __shared__ uint4 sm_fft[];                       // [2^R rows][FFT_COLS] elements
  const size_t n = (size_t)1 << log_n;
  const uint32_t rows = 1u << R;
  const size_t col_blocks = ((size_t)1 << s0) / FFT_COLS;
  const size_t group = blockIdx.x / col_blocks, cb = blockIdx.x % col_blocks;
  const size_t base = (group << (s0 + R)) | (cb * FFT_COLS);
  char* g = reinterpret_cast<char*>(a);
  for (uint32_t e = threadIdx.x; e < rows * FFT_COLS * 2; e += blockDim.x) {       // uint4 granularity: 2 per element
    uint32_t el = e >> 1, row = el / FFT_COLS, c = el % FFT_COLS;
    sm_fft[e] = reinterpret_cast<const uint4*>(g + (base + ((size_t)row << s0) + c) * 32)[e & 1];
  }
  __syncthreads();
  for (uint32_t k = 0; k < R; k++) {
    const uint32_t s = s0 + k, half = 1u << k;
    const size_t twiddle_chunk = n >> (s + 1);
    for (uint32_t t = threadIdx.x; t < rows / 2 * FFT_COLS; t += blockDim.x) {
      uint32_t c = t % FFT_COLS, p = t / FFT_COLS;                                  // p-th butterfly over the rows
      uint32_t i = p & (half - 1), left = ((p >> k) << (k + 1)) + i, right = left + half;
      // position of the left element inside its 2^(s+1) chunk = (low k bits of the row) << s0 | column
      size_t pos = ((size_t)i << s0) | (cb * FFT_COLS + c);
      uint32_t li = (left * FFT_COLS + c) * 2, ri = (right * FFT_COLS + c) * 2;
      Fe<F> l, r;
      {
        uint4 lo = sm_fft[li], hi = sm_fft[li + 1];
        l.v[0] = lo.x; l.v[1] = lo.y; l.v[2] = lo.z; l.v[3] = lo.w; l.v[4] = hi.x; l.v[5] = hi.y; l.v[6] = hi.z; l.v[7] = hi.w;
        lo = sm_fft[ri]; hi = sm_fft[ri + 1];
        r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w; r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
      }
      if (pos) r = fe_mul(r, fe_load<F>(reinterpret_cast<const char*>(tw) + (pos * twiddle_chunk) * 32));
      Fe<F> x = fe_add(l, r), y = fe_sub(l, r);
      sm_fft[li] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]); sm_fft[li + 1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
      sm_fft[ri] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]); sm_fft[ri + 1] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
    }
    __syncthreads();
  }
  for (uint32_t e = threadIdx.x; e < rows * FFT_COLS * 2; e += blockDim.x) {
    uint32_t el = e >> 1, row = el / FFT_COLS, c = el % FFT_COLS;
    reinterpret_cast<uint4*>(g + (base + ((size_t)row << s0) + c) * 32)[e & 1] = sm_fft[e];
  }
}

template <class F>
__global__ void __launch_bounds__(256) k_scale(void* a, size_t n, const void* __restrict__ scale) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<F> d = fe_load<F>(scale);
  char* p = reinterpret_cast<char*>(a) + i * 32;
  fe_store<F>(p, fe_mul(fe_load<F>(p), d));
}

// ------------------------------------------------------------------ lookup argument (row a6)
// evaluate_m (src/plonk/lookup.rs:278-305) with two open-addressing hash tables in HBM keyed by the full 256-bit
// element (a slot remembers one representative INDEX; equality is checked on the 32 input bytes, so there are no
// false matches): table L counts the occurrences of every distinct l value, table T records the smallest index at
// which every distinct t value occurs (the reference's `processed_t`: later duplicates report ZERO).
__device__ __forceinline__ uint32_t lk_hash(const uint4& lo, const uint4& hi) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int i = 0; i < 8; i++) {
    h ^= w[i];
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
  }
  return (uint32_t)(h ^ (h >> 32));
}
__device__ __forceinline__ bool lk_eq(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return ((a0.x ^ b0.x) | (a0.y ^ b0.y) | (a0.z ^ b0.z) | (a0.w ^ b0.w) | (a1.x ^ b1.x) | (a1.y ^ b1.y) | (a1.z ^ b1.z) | (a1.w ^ b1.w)) == 0;
}
// Finds (or, with insert, claims) the slot of element `v` = vals[i].  rep[s] = representative index + 1, 0 = empty.
__device__ __forceinline__ int lk_probe(const void* __restrict__ vals, uint32_t i, uint32_t* rep, uint32_t mask, bool insert) {
  const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(vals) + (size_t)i * 32);
  uint4 v0 = p[0], v1 = p[1];
  uint32_t s = lk_hash(v0, v1) & mask;
  for (;;) {
    uint32_t cur = reinterpret_cast<volatile uint32_t*>(rep)[s];
    if (cur == 0) {
      if (!insert) return -1;
      cur = atomicCAS(&rep[s], 0u, i + 1);
      if (cur == 0) return (int)s;
    }
    const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(vals) + (size_t)(cur - 1) * 32);
    if (lk_eq(v0, v1, q[0], q[1])) return (int)s;
    s = (s + 1) & mask;
  }
}
// same, but the probed element comes from another vector than the table's representatives
__device__ __forceinline__ int lk_find(const void* __restrict__ probe_vals, uint32_t i, const void* __restrict__ table_vals,
                                       const uint32_t* __restrict__ rep, uint32_t mask) {
  const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(probe_vals) + (size_t)i * 32);
  uint4 v0 = p[0], v1 = p[1];
  uint32_t s = lk_hash(v0, v1) & mask;
  for (;;) {
    uint32_t cur = rep[s];
    if (cur == 0) return -1;
    const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(table_vals) + (size_t)(cur - 1) * 32);
    if (lk_eq(v0, v1, q[0], q[1])) return (int)s;
    s = (s + 1) & mask;
  }
}
__global__ void __launch_bounds__(256) k_lookup_count_l(const void* __restrict__ l, uint32_t n_l, uint32_t* rep, uint32_t* cnt, uint32_t mask) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_l) return;
  int s = lk_probe(l, i, rep, mask, true);
  atomicAdd(&cnt[s], 1u);
}
__global__ void __launch_bounds__(256) k_lookup_first_t(const void* __restrict__ t, uint32_t n_t, uint32_t* rep, uint32_t* first, uint32_t mask) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_t) return;
  int s = lk_probe(t, i, rep, mask, true);
  atomicMin(&first[s], i);
}
template <class F>
__global__ void __launch_bounds__(256) k_lookup_m(const void* __restrict__ l, const void* __restrict__ t, uint32_t n_t,
                                                  const uint32_t* __restrict__ rep_l, const uint32_t* __restrict__ cnt_l, uint32_t mask_l,
                                                  uint32_t* rep_t, const uint32_t* __restrict__ first_t, uint32_t mask_t,
                                                  void* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_t) return;
  Fe<F> m = fe_zero<F>();
  int st = lk_probe(t, i, rep_t, mask_t, false);
  if (st >= 0 && first_t[st] == i) {                    // first occurrence of this t value
    int sl = lk_find(t, i, l, rep_l, mask_l);
    if (sl >= 0) {
      m.v[0] = cnt_l[sl];
      m = fe_from_canonical(m);                         // F::from_u128(count)
    }
  }
  fe_store<F>(reinterpret_cast<char*>(out) + (size_t)i * 32, m);
}

// out[i] = mul[i] * (in[i] + r)^{-1}  (mul == nullptr: 1), 0 when in[i] + r == 0  (evaluate_h_g, lookup.rs:307-319).
// Thread t owns elements t, t + T, t + 2T, ... (coalesced), K at a time: Montgomery's trick — prefix products, ONE
// inversion, back-substitution — so an element costs ~3 products plus 1/K of an inversion.  All lanes invert at the
// same time, hence the branch-uniform inversion (field.cuh); it is still worth ~190 products, so long vectors use
// K = 64 (prefixes in local memory) and only short ones, which need the threads, K = 16.
template <class F, int INV_K>
__global__ void __launch_bounds__(128) k_shift_inv_mul(const void* __restrict__ in, const void* __restrict__ mul, size_t n, Fe<F> r,
                                                       void* __restrict__ out) {
  const size_t T = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t base = t; base < n; base += T * INV_K) {
    Fe<F> pre[INV_K];                                   // pre[k] = product of the non-zero denominators before k
    Fe<F> acc = fe_one<F>();
#pragma unroll 1
    for (int k = 0; k < INV_K; k++) {
      size_t i = base + (size_t)k * T;
      pre[k] = acc;
      if (i < n) {
        Fe<F> d = fe_add(fe_load<F>(reinterpret_cast<const char*>(in) + i * 32), r);
        if (!fe_is_zero(d)) acc = fe_mul(acc, d);
      }
    }
    Fe<F> inv = fe_inv_uniform(acc);
#pragma unroll 1
    for (int k = INV_K - 1; k >= 0; k--) {
      size_t i = base + (size_t)k * T;
      if (i < n) {
        Fe<F> d = fe_add(fe_load<F>(reinterpret_cast<const char*>(in) + i * 32), r);
        Fe<F> res = fe_zero<F>();
        if (!fe_is_zero(d)) {
          res = fe_mul(inv, pre[k]);                    // 1 / d
          inv = fe_mul(inv, d);
          if (mul) res = fe_mul(res, fe_load<F>(reinterpret_cast<const char*>(mul) + i * 32));
        }
        fe_store<F>(reinterpret_cast<char*>(out) + i * 32, res);
      }
    }
  }
}

}  // namespace mira

// =================================================================================== host side
struct mira_eval_program {
  int field = 0;
  std::vector<uint32_t> code;
  std::vector<uint8_t> constants;     // n x 32 B
  std::vector<int32_t> rotations;
  uint32_t num_intermediates = 0;
  mira_eval_stats stats{};
  // device copies of the last binding (re-used when the same programs meet the same domain pointers again: a prover
  // binds the same circuit every step, only the challenges change)
  mira_host::DevBuf d_prog, d_uniforms, d_access;
  int device = -1;
  uint64_t serial = 0;                 // unique per created program (addresses can be re-used, serials cannot)
  uint64_t bind_signature = 0;         // of (programs, domain layout, column pointers); 0 = nothing cached
  uint32_t bind_slots = 0, bind_instr = 0, bind_challenge_base = 0, bind_uniforms = 0;
  mira_eval_stats bind_stats{};
  void* h_stage = nullptr;             // pinned staging for the asynchronous uploads
  size_t h_stage_cap = 0;
  cudaEvent_t stage_free = nullptr;    // the previous upload from h_stage has been consumed
};

namespace mira_host {
using namespace mira;

static int set_device(int device) {
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(MIRA_ERR_CUDA, "CUDA device %d not available (%d visible)", device, count);
  CU(cudaSetDevice(device));
  return MIRA_OK;
}
static int valid_field(int f) { return f == MIRA_FQ || f == MIRA_FR; }
static unsigned grid_for(size_t n, int threads, int per_sm) {
  size_t want = (n + threads - 1) / threads, cap = (size_t)148 * per_sm;
  return (unsigned)std::max<size_t>(1, std::min(want, cap));
}

template <class F> static Fe<F> fe_from_host(const void* p) {
  Fe<F> r;
  memcpy(r.v, p, 32);
  return r;
}

template <class F>
static int fold_w_impl(const void* w1, const void* w2, size_t n, const void* r, void* out, cudaStream_t st) {
  if (!n) return MIRA_OK;
  k_fold_w<F><<<grid_for(n, 256, 8), 256, 0, st>>>(w1, w2, n, fe_from_host<F>(r), out);
  CU(cudaGetLastError());
  return MIRA_OK;
}
template <class F>
static int fold_e_impl(const void* e, const void* const* terms, size_t n_terms, size_t n, const void* r, void* out, cudaStream_t st) {
  if (!n) return MIRA_OK;
  FoldTerms ft{};
  for (size_t k = 0; k < n_terms; k++) ft.t[k] = terms[k];
  k_fold_e<F><<<grid_for(n, 256, 8), 256, 0, st>>>(e, ft, (int)n_terms, n, fe_from_host<F>(r), out);
  CU(cudaGetLastError());
  return MIRA_OK;
}

// ---- linker: GraphEvaluator programs + PlonkEvalDomain -> ONE device program ------------------------------
// Several programs bound to the same domain (the 5-6 cross terms of a fold, src/nifs/vanilla/mod.rs:100-121) are
// merged by global value numbering: the reference builds one GraphEvaluator per term, but the terms share most of
// their sub-products (63 % of the multiplications of the primary IVC circuit's six terms are duplicates across
// terms), and a value computed once is exact for every term that uses it.
struct Opnd {
  uint32_t kind, idx;   // DK_*
  bool operator==(const Opnd& o) const { return kind == o.kind && idx == o.idx; }
  bool operator<(const Opnd& o) const { return kind != o.kind ? kind < o.kind : idx < o.idx; }
};
struct LInstr {
  uint32_t op;
  int32_t dst_var;      // variable defined (DOP_OUT: index of the output vector)
  Opnd a, b;
  Opnd c{DK_UNIFORM, 0}, d{DK_UNIFORM, 0};   // only DOP_MUL2ADD / DOP_MUL2SUB
  bool fused() const { return op == DOP_MUL2ADD || op == DOP_MUL2SUB; }
};

struct Linker {          // program-independent part: the domain's columns
  const mira_eval_domain& D;
  std::vector<Access> access;
  std::map<std::pair<const void*, int32_t>, uint32_t> access_ix;
  explicit Linker(const mira_eval_domain& d) : D(d) {}

  uint32_t add_access(const void* ptr, int32_t rot, bool sel) {
    auto key = std::make_pair(ptr, rot);
    auto it = access_ix.find(key);
    if (it != access_ix.end()) return it->second;
    access.push_back(Access{ptr, rot, sel ? 1u : 0u});
    access_ix[key] = (uint32_t)access.size() - 1;
    return (uint32_t)access.size() - 1;
  }
  // PlonkEvalDomain::eval_advice_var (src/plonk/eval.rs:153-228); row-independent part, bounds for the last row
  int advice(size_t index, int32_t rot, Opnd* out) {
    if (D.flags & MIRA_EVAL_LOOKUP_DOMAIN) {     // LookupEvalDomain::eval_advice_var (src/plonk/eval.rs:125-135)
      if (index >= D.num_w1) return fail(MIRA_ERR_EVAL_COLUMN, "column variable index out of boundary: %zu", index);
      if (D.w1_len[index] < D.row_size) return fail(MIRA_ERR_EVAL_ROW, "column variable row index out of boundary: %llu", (unsigned long long)D.w1_len[index]);
      *out = Opnd{DK_ACCESS, add_access(D.w1[index], rot, false)};
      return MIRA_OK;
    }
    size_t row_size = D.row_size, num_advice = D.num_advice, num_lookup = D.num_lookup;
    size_t max_width = num_advice + num_lookup * 5;
    bool first = index < max_width;
    if (!first) index -= max_width;
    size_t num_witness = first ? D.num_w1 : D.num_w2, i, j;
    if (index < num_advice) {
      i = 0; j = index;
    } else {
      size_t li = (index - num_advice) / 5, ls = (index - num_advice) % 5;
      bool first_round = ls < 3;
      if (!first_round) ls -= 3;
      if (num_witness == 2) {
        if (first_round) { i = 0; j = num_advice + li * 3 + ls; } else { i = 1; j = li * 2 + ls; }
      } else if (num_witness == 3) {
        if (first_round) { i = 1; j = li * 3 + ls; } else { i = 2; j = li * 2 + ls; }
      } else {
        return fail(MIRA_ERR_EVAL_WITNESS_INDEX, "Invalid witness index. num_witness: %zu, num_advice: %zu, num_lookup: %zu, index: %zu",
                    num_witness, num_advice, num_lookup, index);
      }
    }
    const void* const* W = first ? D.w1 : D.w2;
    const uint64_t* L = first ? D.w1_len : D.w2_len;
    if (num_witness <= i || L[i] < (j + 1) * row_size)
      return fail(MIRA_ERR_EVAL_WITNESS_INDEX, "Invalid witness index. num_witness: %zu, num_advice: %zu, num_lookup: %zu, index: %zu",
                  num_witness, num_advice, num_lookup, index);
    *out = Opnd{DK_ACCESS, add_access(reinterpret_cast<const char*>(W[i]) + j * row_size * 32, rot, false)};
    return MIRA_OK;
  }
};

// One program decoded into symbolic instructions over its own variable ids.
struct ProgIR {
  std::vector<LInstr> ins;
  Opnd result{DK_UNIFORM, 0};
  bool ssa = true;
  uint32_t NV = 0;
};

// const_map: index of this program's constant i in the uniform table; challenge i sits at challenge_base + i.
static int decode_program(const mira_eval_program& P, Linker& L, const std::vector<uint32_t>& const_map, uint32_t challenge_base,
                          Opnd zero_uniform, ProgIR* out) {
  const mira_eval_domain& D = L.D;
  const std::vector<uint32_t>& code = P.code;
  const uint32_t NV = P.num_intermediates;
  std::vector<LInstr>& ins = out->ins;
  out->NV = NV;
  // pass 1: are targets unique (the form GraphEvaluator::add_calculation produces)?
  std::vector<int> defs(NV, 0);
  for (size_t pc = 0; pc < code.size();) {
    if (pc + 2 > code.size()) return fail(MIRA_ERR_EVAL_PROGRAM, "truncated program");
    uint32_t nops = code[pc] >> 8, target = code[pc + 1];
    if (pc + 2 + 2 * (size_t)nops > code.size() || target >= NV) return fail(MIRA_ERR_EVAL_PROGRAM, "malformed record at word %zu", pc);
    defs[target]++;
    pc += 2 + 2 * (size_t)nops;
  }
  bool ssa = true;
  for (uint32_t v = 0; v < NV; v++) ssa = ssa && defs[v] <= 1;
  out->ssa = ssa;
  // the get_value closure of Calculation::evaluate (graph_evaluator.rs:101-131); intermediates stay symbolic
  auto value = [&](const uint32_t* o, Opnd* r, int32_t* var) -> int {
    uint32_t kind = o[0] & 0xff, rix = o[0] >> 8, index = o[1];
    *var = -1;
    switch (kind) {
      case 0:
        if (index >= const_map.size()) return fail(MIRA_ERR_EVAL_PROGRAM, "constant index %u out of range", index);
        *r = Opnd{DK_UNIFORM, const_map[index]};
        return MIRA_OK;
      case 1:
        if (index >= NV) return fail(MIRA_ERR_EVAL_PROGRAM, "intermediate index %u out of range", index);
        *var = (int32_t)index;
        *r = Opnd{DK_SLOT, index};
        return MIRA_OK;
      case 2:
        if (index >= D.num_fixed) return fail(MIRA_ERR_EVAL_COLUMN, "column variable index out of boundary: %u", index);
        if (rix >= P.rotations.size()) return fail(MIRA_ERR_EVAL_PROGRAM, "rotation index %u out of range", rix);
        *r = Opnd{DK_ACCESS, L.add_access(D.fixed[index], P.rotations[rix], false)};
        return MIRA_OK;
      case 3: {
        if (rix >= P.rotations.size()) return fail(MIRA_ERR_EVAL_PROGRAM, "rotation index %u out of range", rix);
        int32_t rot = P.rotations[rix];
        if (index < D.num_selectors) {       // eval_column_var: selectors, then fixed, then advice
          *r = Opnd{DK_ACCESS, L.add_access(D.selectors[index], rot, true)};
          return MIRA_OK;
        }
        if (index - D.num_selectors < D.num_fixed) {
          *r = Opnd{DK_ACCESS, L.add_access(D.fixed[index - D.num_selectors], rot, false)};
          return MIRA_OK;
        }
        return L.advice((size_t)index - D.num_selectors - D.num_fixed, rot, r);
      }
      case 4:
        if (index >= D.num_challenges)
          return fail(MIRA_ERR_EVAL_CHALLENGE, "challenge index out of boundary: %u (len %u)", index, D.num_challenges);
        *r = Opnd{DK_UNIFORM, challenge_base + index};
        return MIRA_OK;
    }
    return fail(MIRA_ERR_EVAL_PROGRAM, "unknown value source kind %u", kind);
  };
  // pass 2: resolve operands; forward Store(x) (x not an intermediate, or any x when targets are unique)
  std::vector<Opnd> alias(NV, Opnd{DK_SLOT, 0});
  std::vector<char> aliased(NV, 0);
  int32_t last_target = -1;
  auto resolve = [&](const uint32_t* o, Opnd* r) -> int {
    int32_t var;
    int rc = value(o, r, &var);
    if (rc) return rc;
    if (var >= 0 && aliased[var]) *r = alias[var];
    return MIRA_OK;
  };
  for (size_t pc = 0; pc < code.size();) {
    uint32_t op = code[pc] & 0xff, nops = code[pc] >> 8, target = code[pc + 1];
    const uint32_t* o = &code[pc + 2];
    int rc;
    Opnd a{}, b{};
    last_target = (int32_t)target;
    switch (op) {
      case 0: case 1: case 2:
        if (nops != 2) return fail(MIRA_ERR_EVAL_PROGRAM, "binary op with %u operands", nops);
        if ((rc = resolve(o, &a)) || (rc = resolve(o + 2, &b))) return rc;
        ins.push_back(LInstr{op == 0 ? DOP_ADD : op == 1 ? DOP_SUB : DOP_MUL, (int32_t)target, a, b});
        aliased[target] = 0;
        break;
      case 3: case 4: case 5:
        if (nops != 1) return fail(MIRA_ERR_EVAL_PROGRAM, "unary op with %u operands", nops);
        if ((rc = resolve(o, &a))) return rc;
        ins.push_back(LInstr{op == 3 ? DOP_SQUARE : op == 4 ? DOP_DOUBLE : DOP_NEGATE, (int32_t)target, a, a});
        aliased[target] = 0;
        break;
      case 7:
        if (nops != 1) return fail(MIRA_ERR_EVAL_PROGRAM, "Store with %u operands", nops);
        if ((rc = resolve(o, &a))) return rc;
        if (ssa || a.kind != DK_SLOT) {
          alias[target] = a;
          aliased[target] = 1;
        } else {
          ins.push_back(LInstr{DOP_COPY, (int32_t)target, a, a});
          aliased[target] = 0;
        }
        break;
      case 6: {   // Horner(start, factor, parts...): value = value * factor + part
        if (nops < 2) return fail(MIRA_ERR_EVAL_PROGRAM, "Horner with %u operands", nops);
        Opnd fac{}, cur{};
        if ((rc = resolve(o + 2, &fac)) || (rc = resolve(o, &cur))) return rc;
        if (nops == 2) {
          ins.push_back(LInstr{DOP_COPY, (int32_t)target, cur, cur});
        } else {
          for (uint32_t k = 2; k < nops; k++) {
            Opnd part{};
            if ((rc = resolve(o + 2 * k, &part))) return rc;
            ins.push_back(LInstr{DOP_MUL, (int32_t)target, cur, fac});
            cur = Opnd{DK_SLOT, target};
            ins.push_back(LInstr{DOP_ADD, (int32_t)target, cur, part});
          }
        }
        aliased[target] = 0;
        break;
      }
      default:
        return fail(MIRA_ERR_EVAL_PROGRAM, "unknown opcode %u", op);
    }
    pc += 2 + 2 * (size_t)nops;
  }
  // the value of an empty program, and of an intermediate read before any calculation wrote it, is ZERO
  // (the reference zero-initialises `intermediates`, graph_evaluator.rs:354-359)
  out->result = zero_uniform;
  if (last_target >= 0) out->result = aliased[last_target] ? alias[last_target] : Opnd{DK_SLOT, (uint32_t)last_target};
  return MIRA_OK;
}

// MIRA_EVAL_FUSE=0 disables the product-pair fusion (A/B timing and the linker tests exercise both forms)
static bool ctx_fuse_enabled() {
  const char* e = getenv("MIRA_EVAL_FUSE");
  return !(e && e[0] == '0');
}

struct LinkedProgram {
  std::vector<uint4> prog;
  std::vector<Access> access;
  std::vector<uint8_t> uniforms;       // n x 32 B
  uint32_t slots = 0;
  mira_eval_stats stats{};
};

static int link_programs(mira_eval_program* const* progs, size_t n_progs, const mira_eval_domain* D, LinkedProgram* out) {
  Linker L(*D);
  // ---- uniform table: constants (de-duplicated by value across programs when there are several), challenges, ZERO
  std::vector<std::vector<uint32_t>> const_maps(n_progs);
  std::vector<uint8_t>& uni = out->uniforms;
  uni.clear();
  if (n_progs == 1) {
    const auto& c = progs[0]->constants;
    uni.assign(c.begin(), c.end());
    const_maps[0].resize(c.size() / 32);
    for (size_t i = 0; i < const_maps[0].size(); i++) const_maps[0][i] = (uint32_t)i;
  } else {
    std::map<std::string, uint32_t> seen;
    for (size_t p = 0; p < n_progs; p++) {
      const auto& c = progs[p]->constants;
      const_maps[p].resize(c.size() / 32);
      for (size_t i = 0; i < const_maps[p].size(); i++) {
        std::string key(reinterpret_cast<const char*>(c.data() + 32 * i), 32);
        auto it = seen.find(key);
        if (it == seen.end()) {
          it = seen.emplace(key, (uint32_t)(uni.size() / 32)).first;
          uni.insert(uni.end(), c.begin() + 32 * i, c.begin() + 32 * (i + 1));
        }
        const_maps[p][i] = it->second;
      }
    }
  }
  const uint32_t challenge_base = (uint32_t)(uni.size() / 32);
  if (D->num_challenges) uni.insert(uni.end(), (const uint8_t*)D->challenges, (const uint8_t*)D->challenges + (size_t)D->num_challenges * 32);
  const Opnd zero_uniform{DK_UNIFORM, (uint32_t)(uni.size() / 32)};
  uni.insert(uni.end(), 32, 0);

  // ---- decode every program, then merge them into one instruction list over global value ids
  std::vector<ProgIR> irs(n_progs);
  bool all_ssa = true;
  for (size_t p = 0; p < n_progs; p++) {
    int rc = decode_program(*progs[p], L, const_maps[p], challenge_base, zero_uniform, &irs[p]);
    if (rc) return rc;
    all_ssa = all_ssa && irs[p].ssa;
  }
  if (n_progs > 1 && !all_ssa)
    return fail(MIRA_ERR_EVAL_PROGRAM, "programs with re-assigned targets cannot be merged; evaluate them one by one");
  std::vector<LInstr> ins;
  uint32_t NV = 0;
  if (!all_ssa) {                       // single program with re-assigned targets: keep its variables as they are
    ins = irs[0].ins;
    NV = irs[0].NV;
    ins.push_back(LInstr{DOP_OUT, 0, irs[0].result, irs[0].result});
  } else {
    typedef std::tuple<uint32_t, Opnd, Opnd> Key;
    std::map<Key, uint32_t> numbering;
    for (size_t p = 0; p < n_progs; p++) {
      std::vector<Opnd> g(irs[p].NV, zero_uniform);           // local variable -> global operand
      auto map_op = [&](const Opnd& o) -> Opnd { return o.kind == DK_SLOT ? g[o.idx] : o; };
      for (const LInstr& I : irs[p].ins) {
        Opnd a = map_op(I.a), b = I.op <= DOP_MUL ? map_op(I.b) : a;
        if (I.op == DOP_COPY) {
          g[I.dst_var] = a;
          continue;
        }
        if ((I.op == DOP_ADD || I.op == DOP_MUL) && b < a) std::swap(a, b);
        uint32_t op = I.op;
        if (op == DOP_MUL && a == b) op = DOP_SQUARE;
        Key key(op, a, op <= DOP_MUL ? b : a);
        auto it = numbering.find(key);
        if (it == numbering.end()) {
          it = numbering.emplace(key, NV).first;
          ins.push_back(LInstr{op, (int32_t)NV, a, op <= DOP_MUL ? b : a});
          NV++;
        }
        g[I.dst_var] = Opnd{DK_SLOT, it->second};
      }
      Opnd res = map_op(irs[p].result);
      ins.push_back(LInstr{DOP_OUT, (int32_t)p, res, res});     // right after the program's own section
    }
  }

  mira_eval_stats& st = out->stats;
  st = mira_eval_stats{};
  auto operands_of = [](const LInstr& I, Opnd* o) -> int {
    o[0] = I.a;
    if (I.fused()) { o[1] = I.b; o[2] = I.c; o[3] = I.d; return 4; }
    if (I.op <= DOP_MUL) { o[1] = I.b; return 2; }
    return 1;
  };
  // ---- fuse  t1 = a*b; t2 = c*d; r = t1 +- t2  (t1, t2 read nowhere else) into ONE instruction that forms both
  // products under a single Montgomery reduction: 200 wide MACs instead of 272 and two slot round trips fewer.
  // Exact arithmetic makes this value-preserving.  Only when every variable is assigned once.
  if (all_ssa && ctx_fuse_enabled()) {
    std::vector<int> uses(NV, 0), def(NV, -1);
    for (size_t i = 0; i < ins.size(); i++) {
      if (ins[i].op != DOP_OUT) def[ins[i].dst_var] = (int)i;
      Opnd o[4];
      int no = operands_of(ins[i], o);
      for (int k = 0; k < no; k++)
        if (o[k].kind == DK_SLOT) uses[o[k].idx]++;
    }
    auto product_of = [&](const Opnd& o, Opnd* x, Opnd* y) -> bool {
      if (o.kind != DK_SLOT || uses[o.idx] != 1 || def[o.idx] < 0) return false;
      const LInstr& M = ins[def[o.idx]];
      if (M.op == DOP_MUL) { *x = M.a; *y = M.b; return true; }
      if (M.op == DOP_SQUARE) { *x = M.a; *y = M.a; return true; }
      return false;
    };
    for (size_t i = 0; i < ins.size(); i++) {
      LInstr& I = ins[i];
      if (I.op != DOP_ADD && I.op != DOP_SUB) continue;
      if (I.a.kind == DK_SLOT && I.b.kind == DK_SLOT && I.a.idx == I.b.idx) continue;
      Opnd a1, a2, b1, b2;
      if (!product_of(I.a, &a1, &a2) || !product_of(I.b, &b1, &b2)) continue;
      I.op = I.op == DOP_ADD ? DOP_MUL2ADD : DOP_MUL2SUB;
      I.a = a1; I.b = a2; I.c = b1; I.d = b2;       // the two MUL/SQUARE instructions become dead code below
      st.fused++;
    }
  }
  // ---- liveness (last read of every variable), dead-code removal, slot allocation
  const int NI = (int)ins.size();
  std::vector<int> last_use(NV, -1);
  std::vector<char> live(NI, 0);
  {
    std::vector<char> needed(NV, 0);
    for (int i = NI - 1; i >= 0; i--) {
      LInstr& I = ins[i];
      Opnd o[4];
      int no = operands_of(I, o);
      if (I.op != DOP_OUT) {
        if (!needed[I.dst_var]) continue;
        bool self = false;
        for (int k = 0; k < no; k++) self = self || (o[k].kind == DK_SLOT && (int32_t)o[k].idx == I.dst_var);
        if (!self) needed[I.dst_var] = 0;
      }
      live[i] = 1;
      for (int k = 0; k < no; k++)
        if (o[k].kind == DK_SLOT) needed[o[k].idx] = 1;
    }
  }
  for (int i = 0; i < NI; i++) {
    if (!live[i]) continue;
    Opnd o[4];
    int no = operands_of(ins[i], o);
    for (int k = 0; k < no; k++)
      if (o[k].kind == DK_SLOT) last_use[o[k].idx] = i;
  }
  std::vector<int32_t> slot_of(NV, -1);
  std::vector<uint32_t> free_slots;
  uint32_t n_slots = 0;
  out->prog.clear();
  for (int i = 0; i < NI; i++) {
    if (!live[i]) continue;
    LInstr I = ins[i];
    auto map_op = [&](Opnd o) -> Opnd {
      if (o.kind != DK_SLOT) return o;
      return slot_of[o.idx] >= 0 ? Opnd{DK_SLOT, (uint32_t)slot_of[o.idx]} : zero_uniform;
    };
    Opnd src[4], mapped[4];
    int no = operands_of(I, src);
    for (int k = 0; k < no; k++) mapped[k] = map_op(src[k]);
    for (int k = no; k < 4; k++) mapped[k] = mapped[0];
    // operands whose last read is this instruction release their slot before the destination is chosen
    for (int k = 0; k < no; k++) {
      const Opnd& o = src[k];
      if (o.kind == DK_SLOT && last_use[o.idx] == i && (I.op == DOP_OUT || (int32_t)o.idx != I.dst_var) && slot_of[o.idx] >= 0) {
        free_slots.push_back((uint32_t)slot_of[o.idx]);
        slot_of[o.idx] = -1;
      }
    }
    uint32_t dst;
    if (I.op == DOP_OUT) {
      dst = (uint32_t)I.dst_var;
    } else {
      if (slot_of[I.dst_var] < 0) {
        if (!free_slots.empty()) {
          slot_of[I.dst_var] = (int32_t)free_slots.back();
          free_slots.pop_back();
        } else {
          slot_of[I.dst_var] = (int32_t)n_slots++;
        }
      }
      dst = (uint32_t)slot_of[I.dst_var];
    }
    if (dst > 0xffff) return fail(MIRA_ERR_EVAL_PROGRAM, "program needs more than 65535 live intermediates");
    const Opnd &a = mapped[0], &b = mapped[1];
    uint32_t w = 0;
    if (I.fused()) {
      if (mapped[2].idx >= 0x4000u || mapped[3].idx >= 0x4000u) return fail(MIRA_ERR_EVAL_PROGRAM, "fused operand index out of range");
      w = mapped[2].idx | (mapped[2].kind << 14) | (mapped[3].idx << 16) | (mapped[3].kind << 30);
    }
    out->prog.push_back(make_uint4(I.op | (a.kind << 4) | (b.kind << 8) | (dst << 16), a.idx, b.idx, w));
    if (I.fused()) { st.muls += 2; st.adds++; }
    else if (I.op == DOP_MUL || I.op == DOP_SQUARE) st.muls++;
    else if (I.op != DOP_COPY && I.op != DOP_OUT) st.adds++;
    for (int k = 0; k < no; k++) st.loads += mapped[k].kind == DK_ACCESS;
    if (I.op != DOP_OUT) st.instructions++;
  }
  out->access = L.access;
  out->slots = n_slots;
  st.slots = n_slots;
  st.accesses = (uint32_t)L.access.size();
  st.uniforms = (uint32_t)(uni.size() / 32);
  return MIRA_OK;
}

constexpr int MAX_EVAL_OUTPUTS = 16;
typedef mira::EvalOutsDev EvalOuts;

template <class F, int S>
static int launch_eval(const uint4* prog, uint32_t n_instr, const void* uni, const Access* acc, uint64_t rows, uint64_t row_begin,
                       uint64_t row_end, const EvalOuts& outs, cudaStream_t st) {
  size_t smem = (size_t)n_instr * 16;
  if (smem > 200 * 1024) return fail(MIRA_ERR_EVAL_PROGRAM, "program of %u device instructions does not fit shared memory", n_instr);
  if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_eval_rows<F, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const int per_sm = [] { const char* e = getenv("MIRA_EVAL_BLOCKS_PER_SM"); return e ? atoi(e) : 16; }();
  k_eval_rows<F, S><<<grid_for(row_end - row_begin, 128, per_sm), 128, smem, st>>>(prog, n_instr, uni, acc, rows, row_begin, row_end, outs);
  CU(cudaGetLastError());
  return MIRA_OK;
}

static uint64_t fnv(uint64_t h, const void* p, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 0x100000001b3ull;
  return h;
}
// Everything the linked program depends on except the challenge VALUES (those only fill the uniform table).
static uint64_t bind_signature(mira_eval_program* const* progs, size_t n_progs, const mira_eval_domain* D) {
  uint64_t h = 0xcbf29ce484222325ull;
  h = fnv(h, &n_progs, sizeof n_progs);
  for (size_t p = 0; p < n_progs; p++) {
    h = fnv(h, &progs[p]->serial, sizeof progs[p]->serial);
  }
  h = fnv(h, &D->row_size, sizeof D->row_size);
  uint32_t f[8] = {D->num_selectors, D->num_fixed, D->num_advice, D->num_lookup, D->num_challenges, D->num_w1, D->num_w2, D->flags};
  h = fnv(h, f, sizeof f);
  if (D->num_selectors) h = fnv(h, D->selectors, D->num_selectors * sizeof(void*));
  if (D->num_fixed) h = fnv(h, D->fixed, D->num_fixed * sizeof(void*));
  if (D->num_w1) { h = fnv(h, D->w1, D->num_w1 * sizeof(void*)); h = fnv(h, D->w1_len, D->num_w1 * sizeof(uint64_t)); }
  if (D->num_w2) { h = fnv(h, D->w2, D->num_w2 * sizeof(void*)); h = fnv(h, D->w2_len, D->num_w2 * sizeof(uint64_t)); }
  const char* e = getenv("MIRA_EVAL_FUSE");
  uint8_t fuse = !(e && e[0] == '0');
  h = fnv(h, &fuse, 1);
  return h ? h : 1;
}

template <class F>
static int eval_impl(mira_eval_program* const* progs, size_t n_progs, const mira_eval_domain* D, uint64_t row_begin, uint64_t row_end,
                     void* const* outs, cudaStream_t st) {
  int rc;
  mira_eval_program* P = progs[0];          // device copies of the linked program live with the first program
  const uint64_t sig = bind_signature(progs, n_progs, D);
  auto stage = [&](size_t bytes) -> int {   // pinned staging buffer, free again once the previous upload has been consumed
    if (!P->stage_free) CU(cudaEventCreateWithFlags(&P->stage_free, cudaEventDisableTiming));
    CU(cudaEventSynchronize(P->stage_free));
    if (bytes > P->h_stage_cap) {
      if (P->h_stage) cudaFreeHost(P->h_stage);
      P->h_stage = nullptr;
      P->h_stage_cap = 0;
      CU(cudaMallocHost(&P->h_stage, bytes));
      P->h_stage_cap = bytes;
    }
    return MIRA_OK;
  };
  if (sig == P->bind_signature && P->d_prog.p) {
    // same circuit, same columns: only the challenges (a few uniforms) are refreshed, asynchronously on `st`
    for (size_t p = 0; p < n_progs; p++) progs[p]->stats = P->bind_stats;
    if (row_begin >= row_end) return MIRA_OK;
    if (D->num_challenges) {
      size_t bytes = (size_t)D->num_challenges * 32;
      if ((rc = stage(bytes))) return rc;
      memcpy(P->h_stage, D->challenges, bytes);
      CU(cudaMemcpyAsync((char*)P->d_uniforms.p + (size_t)P->bind_challenge_base * 32, P->h_stage, bytes, cudaMemcpyHostToDevice, st));
      CU(cudaEventRecord(P->stage_free, st));
    }
  } else {
    LinkedProgram lp;
    if ((rc = link_programs(progs, n_progs, D, &lp))) return rc;
    for (size_t p = 0; p < n_progs; p++) progs[p]->stats = lp.stats;
    if (row_begin >= row_end) return MIRA_OK;
    const size_t b_prog = std::max<size_t>(lp.prog.size(), 1) * 16, b_uni = lp.uniforms.size(),
                 b_acc = std::max<size_t>(lp.access.size(), 1) * sizeof(Access);
    P->bind_signature = 0;
    // a kernel still running on another stream may read the old device program: drain before replacing it
    if (P->d_prog.p) CU(cudaDeviceSynchronize());
    if ((rc = P->d_prog.ensure(b_prog)) || (rc = P->d_uniforms.ensure(b_uni)) || (rc = P->d_access.ensure(b_acc))) return rc;
    if ((rc = stage(b_prog + b_uni + b_acc))) return rc;
    char* h = (char*)P->h_stage;
    memcpy(h, lp.prog.data(), lp.prog.size() * 16);
    memcpy(h + b_prog, lp.uniforms.data(), b_uni);
    if (!lp.access.empty()) memcpy(h + b_prog + b_uni, lp.access.data(), lp.access.size() * sizeof(Access));
    CU(cudaMemcpyAsync(P->d_prog.p, h, b_prog, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(P->d_uniforms.p, h + b_prog, b_uni, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(P->d_access.p, h + b_prog + b_uni, b_acc, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(P->stage_free, st));
    P->bind_slots = lp.slots;
    P->bind_instr = (uint32_t)lp.prog.size();
    P->bind_uniforms = (uint32_t)(b_uni / 32);
    P->bind_challenge_base = P->bind_uniforms - 1 - D->num_challenges;     // constants | challenges | ZERO
    P->bind_stats = lp.stats;
    P->bind_signature = sig;
  }
  EvalOuts eo{};
  for (size_t p = 0; p < n_progs; p++) eo.p[p] = outs[p];
  const uint4* dp = (const uint4*)P->d_prog.p;
  const Access* da = (const Access*)P->d_access.p;
  const uint32_t ni = P->bind_instr, slots = P->bind_slots;
  if (slots <= 16) return launch_eval<F, 16>(dp, ni, P->d_uniforms.p, da, D->row_size, row_begin, row_end, eo, st);
  if (slots <= 64) return launch_eval<F, 64>(dp, ni, P->d_uniforms.p, da, D->row_size, row_begin, row_end, eo, st);
  if (slots <= 128) return launch_eval<F, 128>(dp, ni, P->d_uniforms.p, da, D->row_size, row_begin, row_end, eo, st);
  if (slots <= 256) return launch_eval<F, 256>(dp, ni, P->d_uniforms.p, da, D->row_size, row_begin, row_end, eo, st);
  P->bind_signature = 0;
  return fail(MIRA_ERR_EVAL_PROGRAM, "program keeps %u intermediates live; the device interpreter supports 256", slots);
}

// ---- FFT -------------------------------------------------------------------------------------------
template <class F>
static int fft_impl(void* a, uint32_t log_n, const void* d_consts, bool scale, cudaStream_t st) {
  const size_t n = (size_t)1 << log_n, half = n / 2;
  if (log_n == 0) {
    if (scale) k_scale<F><<<1, 256, 0, st>>>(a, 1, reinterpret_cast<const char*>(d_consts) + 32);
    CU(cudaGetLastError());
    return MIRA_OK;
  }
  void* tw = nullptr;
  CU(cudaMallocAsync(&tw, half * 32, st));
  const size_t low = (size_t)1 << FFT_TW_LOW;
  if (half <= low) {
    k_fft_twiddles<F><<<(unsigned)((half + 255) / 256), 256, 0, st>>>(d_consts, half, 1, tw);
  } else {
    void* hi = nullptr;
    if (cudaMallocAsync(&hi, (half >> FFT_TW_LOW) * 32, st) != cudaSuccess) {
      cudaFreeAsync(tw, st);
      return fail(MIRA_ERR_CUDA, "fft: cannot allocate the twiddle table");
    }
    k_fft_twiddles<F><<<(unsigned)(low / 256), 256, 0, st>>>(d_consts, low, 1, tw);
    k_fft_twiddles<F><<<(unsigned)(((half >> FFT_TW_LOW) + 255) / 256), 256, 0, st>>>(d_consts, half >> FFT_TW_LOW, low, hi);
    k_fft_twiddles_combine<F><<<(unsigned)((half - low + 255) / 256), 256, 0, st>>>(hi, half, tw);
    cudaFreeAsync(hi, st);
  }
  k_fft_bitrev<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, log_n);
  uint32_t fused = log_n < (uint32_t)FFT_TILE_LOG ? log_n : (uint32_t)FFT_TILE_LOG;
  size_t tile = (size_t)1 << fused;
  k_fft_tile<F><<<(unsigned)(n / tile), 512, tile * 32, st>>>(a, log_n, fused, tw);
  // remaining stages: up to 7 per HBM round trip (strided tiles); a lone last stage goes through the per-stage kernel
  for (uint32_t s = fused; s < log_n;) {
    uint32_t R = log_n - s < (uint32_t)FFT_STRIDED_LOG ? log_n - s : (uint32_t)FFT_STRIDED_LOG;
    if (R == 1) {
      k_fft_stage<F><<<(unsigned)((half + 255) / 256), 256, 0, st>>>(a, log_n, s, tw);
    } else {
      size_t blocks = (n >> (s + R)) * (((size_t)1 << s) / FFT_COLS);
      k_fft_strided<F><<<(unsigned)blocks, 256, ((size_t)1 << R) * FFT_COLS * 32, st>>>(a, log_n, s, R, tw);
    }
    s += R;
  }
  if (scale) k_scale<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, n, reinterpret_cast<const char*>(d_consts) + 32);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(tw, st);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "fft launch failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

// ---- lookup argument ---------------------------------------------------------------------------------
template <class F>
static int lookup_m_impl(const void* l, size_t n_l, const void* t, size_t n_t, void* out, cudaStream_t st) {
  auto cap_for = [](size_t n) { size_t c = 64; while (c < 2 * n) c <<= 1; return c; };
  const size_t cl = cap_for(n_l), ct = cap_for(n_t);
  uint32_t* buf = nullptr;                 // rep_l | cnt_l | rep_t | first_t
  CU(cudaMallocAsync((void**)&buf, (2 * cl + 2 * ct) * 4, st));
  uint32_t *rep_l = buf, *cnt_l = buf + cl, *rep_t = buf + 2 * cl, *first_t = buf + 2 * cl + ct;
  CU(cudaMemsetAsync(buf, 0, (2 * cl + ct) * 4, st));
  CU(cudaMemsetAsync(first_t, 0xff, ct * 4, st));
  if (n_l) mira::k_lookup_count_l<<<(unsigned)((n_l + 255) / 256), 256, 0, st>>>(l, (uint32_t)n_l, rep_l, cnt_l, (uint32_t)cl - 1);
  mira::k_lookup_first_t<<<(unsigned)((n_t + 255) / 256), 256, 0, st>>>(t, (uint32_t)n_t, rep_t, first_t, (uint32_t)ct - 1);
  mira::k_lookup_m<F><<<(unsigned)((n_t + 255) / 256), 256, 0, st>>>(l, t, (uint32_t)n_t, rep_l, cnt_l, (uint32_t)cl - 1, rep_t, first_t,
                                                                     (uint32_t)ct - 1, out);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(buf, st);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "lookup_m launch failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

template <class F>
static int shift_inv_impl(const void* in, const void* mul, size_t n, const void* r, void* out, cudaStream_t st) {
  if (n >= ((size_t)1 << 21)) {
    mira::k_shift_inv_mul<F, 64><<<grid_for((n + 63) / 64, 128, 8), 128, 0, st>>>(in, mul, n, fe_from_host<F>(r), out);
  } else {
    mira::k_shift_inv_mul<F, 16><<<grid_for((n + 15) / 16, 128, 8), 128, 0, st>>>(in, mul, n, fe_from_host<F>(r), out);
  }
  CU(cudaGetLastError());
  return MIRA_OK;
}

}  // namespace mira_host

using namespace mira_host;

extern "C" {

int mira_fold_w(int field, const void* w1, const void* w2, size_t n, const void* r, void* out, int device, void* stream) {
  if (!valid_field(field)) return fail(MIRA_ERR_INVALID, "unknown field %d", field);
  if (!r || (n && (!w1 || !w2 || !out))) return fail(MIRA_ERR_INVALID, "null argument");
  int rc = set_device(device);
  if (rc) return rc;
  return field == MIRA_FQ ? fold_w_impl<mira::FqTag>(w1, w2, n, r, out, (cudaStream_t)stream)
                          : fold_w_impl<mira::FrTag>(w1, w2, n, r, out, (cudaStream_t)stream);
}

int mira_fold_e(int field, const void* e, const void* const* terms, size_t n_terms, size_t n, const void* r, void* out, int device,
                void* stream) {
  if (!valid_field(field)) return fail(MIRA_ERR_INVALID, "unknown field %d", field);
  if (!r || (n && (!e || !out)) || (n_terms && !terms)) return fail(MIRA_ERR_INVALID, "null argument");
  if (n_terms > (size_t)mira::MAX_FOLD_TERMS) return fail(MIRA_ERR_INVALID, "at most %d cross terms", mira::MAX_FOLD_TERMS);
  for (size_t k = 0; k < n_terms; k++)
    if (n && !terms[k]) return fail(MIRA_ERR_INVALID, "cross term %zu is null", k);
  int rc = set_device(device);
  if (rc) return rc;
  return field == MIRA_FQ ? fold_e_impl<mira::FqTag>(e, terms, n_terms, n, r, out, (cudaStream_t)stream)
                          : fold_e_impl<mira::FrTag>(e, terms, n_terms, n, r, out, (cudaStream_t)stream);
}

int mira_concat_pad(const void* const* cols, const size_t* lens, size_t n_cols, size_t pad_size, void* out, size_t cap, size_t* out_len,
                    int device, void* stream) {
  if (n_cols && (!cols || !lens)) return fail(MIRA_ERR_INVALID, "null argument");
  size_t total = 0;
  for (size_t c = 0; c < n_cols; c++) total += lens[c] > pad_size ? lens[c] : pad_size;
  if (out_len) *out_len = total;
  if (total > cap) return fail(MIRA_ERR_INVALID, "concatenation needs %zu elements, output holds %zu", total, cap);
  if (total && !out) return fail(MIRA_ERR_INVALID, "null output");
  int rc = set_device(device);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = reinterpret_cast<char*>(out);
  for (size_t c = 0; c < n_cols; c++) {     // one copy-engine transfer + one memset per column
    size_t l = lens[c], tot = l > pad_size ? l : pad_size;
    if (l) CU(cudaMemcpyAsync(w, cols[c], l * 32, cudaMemcpyDeviceToDevice, st));
    if (tot > l) CU(cudaMemsetAsync(w + l * 32, 0, (tot - l) * 32, st));
    w += tot * 32;
  }
  return MIRA_OK;
}

int mira_eval_program_create(int field, const uint32_t* code, size_t code_words, const void* constants, size_t n_constants,
                             const int32_t* rotations, size_t n_rotations, uint32_t num_intermediates, mira_eval_program** out) {
  if (!out) return fail(MIRA_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!valid_field(field)) return fail(MIRA_ERR_INVALID, "unknown field %d", field);
  if ((code_words && !code) || (n_constants && !constants) || (n_rotations && !rotations)) return fail(MIRA_ERR_INVALID, "null argument");
  static std::atomic<uint64_t> next_serial{1};
  auto* p = new mira_eval_program();
  p->serial = next_serial.fetch_add(1);
  p->field = field;
  p->code.assign(code, code + code_words);
  p->constants.assign((const uint8_t*)constants, (const uint8_t*)constants + n_constants * 32);
  p->rotations.assign(rotations, rotations + n_rotations);
  p->num_intermediates = num_intermediates;
  *out = p;
  return MIRA_OK;
}

void mira_eval_program_destroy(mira_eval_program* p) {
  if (!p) return;
  if (p->device >= 0) cudaSetDevice(p->device);
  p->d_prog.release();
  p->d_uniforms.release();
  p->d_access.release();
  if (p->h_stage) cudaFreeHost(p->h_stage);
  if (p->stage_free) cudaEventDestroy(p->stage_free);
  delete p;
}

static int check_domain(const mira_eval_domain* dom) {
  if ((dom->num_selectors && !dom->selectors) || (dom->num_fixed && !dom->fixed) || (dom->num_w1 && (!dom->w1 || !dom->w1_len)) ||
      (dom->num_w2 && (!dom->w2 || !dom->w2_len)) || (dom->num_challenges && !dom->challenges))
    return fail(MIRA_ERR_INVALID, "domain has a null column table");
  if (dom->row_size >= ((uint64_t)1 << 40)) return fail(MIRA_ERR_INVALID, "row_size too large");
  return MIRA_OK;
}

int mira_eval_rows(const mira_eval_program* prog, const mira_eval_domain* dom, void* out, int device, void* stream) {
  if (!dom) return fail(MIRA_ERR_INVALID, "null argument");
  return mira_eval_rows_range(prog, dom, 0, dom->row_size, out, device, stream);
}

int mira_eval_rows_range(const mira_eval_program* prog, const mira_eval_domain* dom, uint64_t row_begin, uint64_t row_end, void* out,
                         int device, void* stream) {
  void* outs[1] = {out};
  return mira_eval_rows_multi(&prog, 1, dom, row_begin, row_end, outs, device, stream);
}

int mira_eval_rows_multi(const mira_eval_program* const* progs, size_t n_progs, const mira_eval_domain* dom, uint64_t row_begin,
                         uint64_t row_end, void* const* outs, int device, void* stream) {
  if (!progs || !dom || !n_progs || !outs) return fail(MIRA_ERR_INVALID, "null argument");
  if (n_progs > (size_t)MAX_EVAL_OUTPUTS) return fail(MIRA_ERR_INVALID, "at most %d programs per call", MAX_EVAL_OUTPUTS);
  if (row_begin > row_end || row_end > dom->row_size)
    return fail(MIRA_ERR_EVAL_ROW, "column variable row index out of boundary: %llu", (unsigned long long)row_end);
  int rc = check_domain(dom);
  if (rc) return rc;
  for (size_t p = 0; p < n_progs; p++) {
    if (!progs[p]) return fail(MIRA_ERR_INVALID, "program %zu is null", p);
    if (progs[p]->field != progs[0]->field) return fail(MIRA_ERR_INVALID, "programs of different fields in one call");
    if (row_end > row_begin && !outs[p]) return fail(MIRA_ERR_INVALID, "null output");
  }
  if ((rc = set_device(device))) return rc;
  auto** ps = const_cast<mira_eval_program**>(progs);
  mira_eval_program* p0 = ps[0];
  if (p0->device >= 0 && p0->device != device) {
    cudaSetDevice(p0->device);
    p0->d_prog.release(); p0->d_uniforms.release(); p0->d_access.release();
    p0->bind_signature = 0;
    cudaSetDevice(device);
  }
  p0->device = device;
  return p0->field == MIRA_FQ ? eval_impl<mira::FqTag>(ps, n_progs, dom, row_begin, row_end, outs, (cudaStream_t)stream)
                              : eval_impl<mira::FrTag>(ps, n_progs, dom, row_begin, row_end, outs, (cudaStream_t)stream);
}

int mira_test_eval_link_multi(const mira_eval_program* const* progs, size_t n_progs, const mira_eval_domain* dom, uint32_t* instr_words,
                              size_t instr_cap, size_t* n_instr, uint64_t* access_words, size_t access_cap, size_t* n_access,
                              void* uniform_bytes, size_t uniform_cap, size_t* n_uniforms, uint32_t* n_slots) {
  if (!progs || !n_progs || !dom || !n_instr || !n_access || !n_uniforms || !n_slots) return fail(MIRA_ERR_INVALID, "null argument");
  LinkedProgram lp;
  int rc = link_programs(const_cast<mira_eval_program**>(progs), n_progs, dom, &lp);
  if (rc) return rc;
  for (size_t p = 0; p < n_progs; p++) const_cast<mira_eval_program*>(progs[p])->stats = lp.stats;
  *n_instr = lp.prog.size();
  *n_access = lp.access.size();
  *n_uniforms = lp.uniforms.size() / 32;
  *n_slots = lp.slots;
  if (lp.prog.size() > instr_cap || lp.access.size() > access_cap || lp.uniforms.size() / 32 > uniform_cap)
    return fail(MIRA_ERR_INVALID, "output buffers too small");
  for (size_t i = 0; i < lp.prog.size(); i++) {
    instr_words[4 * i] = lp.prog[i].x; instr_words[4 * i + 1] = lp.prog[i].y; instr_words[4 * i + 2] = lp.prog[i].z; instr_words[4 * i + 3] = lp.prog[i].w;
  }
  for (size_t i = 0; i < lp.access.size(); i++) {
    access_words[3 * i] = (uint64_t)(uintptr_t)lp.access[i].ptr;
    access_words[3 * i + 1] = (uint64_t)(int64_t)lp.access[i].rot;
    access_words[3 * i + 2] = lp.access[i].is_selector;
  }
  if (uniform_bytes) memcpy(uniform_bytes, lp.uniforms.data(), lp.uniforms.size());
  return MIRA_OK;
}

int mira_eval_program_stats(const mira_eval_program* prog, mira_eval_stats* out) {
  if (!prog || !out) return fail(MIRA_ERR_INVALID, "null argument");
  *out = prog->stats;
  return MIRA_OK;
}

int mira_lookup_m(int field, const void* l, size_t n_l, const void* t, size_t n_t, void* out_m, int device, void* stream) {
  if (!valid_field(field)) return fail(MIRA_ERR_INVALID, "unknown field %d", field);
  if ((n_l && !l) || (n_t && (!t || !out_m))) return fail(MIRA_ERR_INVALID, "null argument");
  if (n_l >= ((size_t)1 << 30) || n_t >= ((size_t)1 << 30)) return fail(MIRA_ERR_INVALID, "lookup vectors too long");
  int rc = set_device(device);
  if (rc || !n_t) return rc;
  return field == MIRA_FQ ? lookup_m_impl<mira::FqTag>(l, n_l, t, n_t, out_m, (cudaStream_t)stream)
                          : lookup_m_impl<mira::FrTag>(l, n_l, t, n_t, out_m, (cudaStream_t)stream);
}

int mira_lookup_h_g(int field, const void* l, const void* t, const void* m, size_t n, const void* r, void* out_h, void* out_g, int device,
                    void* stream) {
  if (!valid_field(field)) return fail(MIRA_ERR_INVALID, "unknown field %d", field);
  if (!r || (n && (!l || !t || !m || !out_h || !out_g))) return fail(MIRA_ERR_INVALID, "null argument");
  int rc = set_device(device);
  if (rc || !n) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (field == MIRA_FQ) {
    if ((rc = shift_inv_impl<mira::FqTag>(l, nullptr, n, r, out_h, st))) return rc;
    return shift_inv_impl<mira::FqTag>(t, m, n, r, out_g, st);
  }
  if ((rc = shift_inv_impl<mira::FrTag>(l, nullptr, n, r, out_h, st))) return rc;
  return shift_inv_impl<mira::FrTag>(t, m, n, r, out_g, st);
}

int mira_fft(int field, void* a, uint32_t log_n, const void* omega, int device, void* stream) {
  if (!valid_field(field)) return fail(MIRA_ERR_INVALID, "unknown field %d", field);
  if (!a || !omega) return fail(MIRA_ERR_INVALID, "null argument");
  if (log_n > 31) return fail(MIRA_ERR_INVALID, "log_n = %u too large", log_n);
  int rc = set_device(device);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  void* consts = nullptr;
  CU(cudaMallocAsync(&consts, 64, st));
  CU(cudaMemcpyAsync(consts, omega, 32, cudaMemcpyHostToDevice, st));
  rc = field == MIRA_FQ ? fft_impl<mira::FqTag>(a, log_n, consts, false, st) : fft_impl<mira::FrTag>(a, log_n, consts, false, st);
  cudaFreeAsync(consts, st);
  return rc;
}

int mira_fft_std(int field, void* a, uint32_t log_n, int inverse, int device, void* stream) {
  if (!valid_field(field)) return fail(MIRA_ERR_INVALID, "unknown field %d", field);
  if (!a) return fail(MIRA_ERR_INVALID, "null argument");
  // get_omega_or_inv's assert (src/fft.rs:13): k <= F::S.  bn256::Fr: S = 28; bn256::Fq: S = 1, no ROOT_OF_UNITY here
  if (field != MIRA_FR || log_n > 28) return fail(MIRA_ERR_INVALID, "k=%u should no larger than F::S=%d", log_n, field == MIRA_FR ? 28 : 0);
  int rc = set_device(device);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  void* consts = nullptr;
  CU(cudaMallocAsync(&consts, 64, st));
  mira::k_fft_consts<mira::FrTag><<<1, 32, 0, st>>>(log_n, inverse, consts);
  rc = fft_impl<mira::FrTag>(a, log_n, consts, inverse != 0, st);
  cudaFreeAsync(consts, st);
  return rc;
}

}  // extern "C"
