// fp52.cuh — EXPERIMENT, not part of the shipped library: BN254 Fq / Fr Montgomery products on
// the FP64 pipe, 5 signed 52-bit limbs held in doubles, R = 2^260, to be compared with the IMAD.WIDE product of
// mira_b200/csrc/field.cuh (device replacement of halo2curves' `Fq` / `Fr`, src/lib.rs:24-27, src/commitment.rs:11).
//   * a limb product a_i*b_j (|a_i*b_j| <= 2^103) is split exactly into H*2^52 + L, 0 <= L < 2^52, by two DFMA.RZ and
//     one DADD:  hi = fma_rz(a, b, 1.5*2^104)  ->  hi = 1.5*2^104 + H*2^52 with H = floor(a*b / 2^52);
//                lo = fma_rz(a, b, (1.5*2^104 + 2^52) - hi) = 2^52 + L   (exact);
//   * the bit patterns of hi and lo are  K1 + H  and  K2 + L  as 64-bit integers (the exponent fields are fixed), so
//     column sums are plain integer additions on the ALU pipe and the constants are subtracted once per column;
//   * Montgomery reduction is word-serial over the five 52-bit columns; q_i = -col_i / p mod 2^52 is a 64-bit
//     integer product, taken as a SIGNED 52-bit value so that |q_i * p_j| <= 2^103 as well.
// Limbs are signed ("balanced", |limb| <= 2^51 after a product), so field subtraction is five DADDs with no borrow
// handling, and values are only congruent, not reduced: a product of |a|, |b| < 8p returns |r| < 1.3p (R / p = 2^6.4
// absorbs the slack).  fd_from_std / fd_to_std convert from and to the reference's byte layout.
// Every function is __host__ __device__; fp52_check.py drives the host build against Python big integers (24,000
// cases, bit-exact), modmul.cu checks fp52 == field.cuh on the device (2.4 M cases) and times both.
//
// RESULT on B200 (profiles/r02_modmul_microbench.jsonl, profiles/r02_modmul_ncu.txt): 160 FP64 + 159 ALU + 28 other
// instructions per product; 6.0-6.2e10 products/s against 6.5e10 for the IMAD.WIDE product (sqr 7.1e10 vs 6.5e10):
// FP64 and ALU pipes both sit at ~52 %, issue at 56 % (math-pipe-throttle, dispatch and fixed-latency waits with 4
// warps per scheduler) — no faster than the integer product, so the shipped multiplier stays IMAD.WIDE.  Warps running
// this product beside warps running the IMAD product on the same SM do not add up either (./modmul mix,
// profiles/r02_modmul_mixed_warps.jsonl: the combined rate never exceeds IMAD.WIDE alone).
#pragma once
#include <cstdint>
#include <cstring>
#include <cfenv>
#include <cmath>
#include "../../mira_b200/csrc/field.cuh"

namespace mira {
namespace fp52 {

#if defined(__CUDA_ARCH__)
#define MIRA_FD_HD __host__ __device__ __forceinline__
__device__ __forceinline__ double fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
__device__ __forceinline__ int64_t d2l(double x) { return __double_as_longlong(x); }
__device__ __forceinline__ double l2d(int64_t x) { return __longlong_as_double(x); }
#else
#define MIRA_FD_HD __host__ __device__ inline
// host (tests only): one correctly rounded fused multiply-add toward zero; every other FP operation here is exact or
// round-to-nearest on both sides
inline double fma_rz(double a, double b, double c) {
  volatile double va = a, vb = b, vc = c;
  std::fesetround(FE_TOWARDZERO);
  volatile double r = std::fma(va, vb, vc);
  std::fesetround(FE_TONEAREST);
  return r;
}
inline int64_t d2l(double x) { int64_t r; std::memcpy(&r, &x, 8); return r; }
inline double l2d(int64_t x) { double r; std::memcpy(&r, &x, 8); return r; }
#endif

constexpr double C1 = 0x1.8p104;                    // 1.5 * 2^104: a*b + C1 stays in [2^104, 2^105) for |a*b| < 2^103
constexpr double C1P = 0x1.8000000000001p104;       // C1 + 2^52
constexpr double C2 = 0x1.8p52;                     // 1.5 * 2^52: int -> double conversion offset
constexpr uint64_t K1 = (0x467ull << 52) | (1ull << 51);   // bits(C1 + H*2^52) = K1 + H
constexpr uint64_t K2 = 0x433ull << 52;                     // bits(2^52 + L)     = K2 + L
constexpr uint64_t KC = K2 + (1ull << 51);                  // bits(C2 + v)       = KC + v  for |v| < 2^51

template <class F> struct P52;
template <> struct P52<FqTag> {
  // p in balanced 52-bit limbs; NP = -p^-1 mod 2^52
  __host__ __device__ static constexpr double p(int i) {
    constexpr double v[5] = {154029749239111.0, -1945555279752254.0, 423691504025963.0, -1685982885422232.0, 53207371014450.0};
    return v[i];
  }
  static constexpr uint64_t NP = 0x20782e4866389ull;
  // 2^264 mod p (std Montgomery form -> R = 2^260 form) and 2^260 mod p (one)
  __host__ __device__ static constexpr double r264(int i) {
    constexpr double v[5] = {-1390697610713478.0, -323933227188290.0, -1721143775100325.0, -504184215139471.0, 14813684363143.0};
    return v[i];
  }
  __host__ __device__ static constexpr double one(int i) {
    constexpr double v[5] = {572299946026164.0, 1297056913851477.0, 438710680783112.0, 2010973926982104.0, 34180462156727.0};
    return v[i];
  }
};
template <> struct P52<FrTag> {
  __host__ __device__ static constexpr double p(int i) {
    constexpr double v[5] = {551490712240129.0, 1275002230477886.0, 423691496731624.0, -1685982885422232.0, 53207371014450.0};
    return v[i];
  }
  static constexpr uint64_t NP = 0x1f593efffffffull;
  __host__ __device__ static constexpr double r264(int i) {
    constexpr double v[5] = {879113770367670.0, -1474362784157842.0, -1721133898566287.0, -504184215139471.0, 14813684363143.0};
    return v[i];
  }
  __host__ __device__ static constexpr double one(int i) {
    constexpr double v[5] = {-1289223554465876.0, 986203696749470.0, 438711293507528.0, 2010973926982104.0, 34180462156727.0};
    return v[i];
  }
};

template <class F>
struct Fd {
  double v[5];       // value = sum v[i] * 2^(52 i); congruent to x * 2^260 mod the modulus
};

// number of (i, j) in [0,5)^2 with i + j == k; the same over i <= j only (squaring)
__host__ __device__ constexpr int cnt5(int k) { return k < 0 || k > 8 ? 0 : (k < 5 ? k + 1 : 9 - k); }
__host__ __device__ constexpr int cnt5_tri(int k) { return k < 0 || k > 8 ? 0 : (cnt5(k) + 1) / 2; }

// one limb product into two columns
MIRA_FD_HD void mac_split(uint64_t& c_lo, uint64_t& c_hi, double a, double b) {
  double hi = fma_rz(a, b, C1);
  double lo = fma_rz(a, b, C1P - hi);
  c_lo += (uint64_t)d2l(lo);
  c_hi += (uint64_t)d2l(hi);
}

// signed 52-bit integer -> double (exact)
MIRA_FD_HD double s52_to_double(int64_t v) { return l2d((int64_t)(KC + (uint64_t)v)) - C2; }

// Montgomery reduction of the 10 columns (col[k] weighs 2^(52k); constants of the NR*25 products already folded in by
// the caller's initial values) and carry normalisation into balanced limbs.
template <class F, int NPROD>
MIRA_FD_HD void fd_reduce_cols(double (&r)[5], uint64_t (&col)[11]) {
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint64_t t = col[i] * P52<F>::NP;
    int64_t q = (int64_t)(t << 12) >> 12;                 // signed low 52 bits
    double qd = s52_to_double(q);
#pragma unroll
    for (int j = 0; j < 5; j++) mac_split(col[i + j], col[i + j + 1], qd, P52<F>::p(j));
    col[i + 1] += (uint64_t)((int64_t)col[i] >> 52);      // col[i] is now a multiple of 2^52
  }
#pragma unroll
  for (int k = 5; k < 9; k++) {
    int64_t c = (int64_t)(col[k] + (1ull << 51)) >> 52;
    int64_t l = (int64_t)col[k] - (int64_t)((uint64_t)c << 52);
    col[k + 1] += (uint64_t)c;
    r[k - 5] = s52_to_double(l);
  }
  r[4] = s52_to_double((int64_t)col[9]);
}

// initial column values: minus the bit-pattern constants of every lo / hi term that will be added
// (nprod full 5x5 products plus the 5x5 products of the reduction)
template <int NPROD>
MIRA_FD_HD void fd_init_cols(uint64_t (&col)[11]) {
#pragma unroll
  for (int k = 0; k < 11; k++)
    col[k] = 0ull - ((uint64_t)((NPROD + 1) * cnt5(k)) * K2 + (uint64_t)((NPROD + 1) * cnt5(k - 1)) * K1);
}

// r = a * b / 2^260  (|a_i * b_j| < 2^103 for all limb pairs: e.g. one operand un-normalised by one addition)
template <class F>
MIRA_FD_HD Fd<F> fd_mul(const Fd<F>& a, const Fd<F>& b) {
  uint64_t col[11];
  fd_init_cols<1>(col);
#pragma unroll
  for (int i = 0; i < 5; i++)
#pragma unroll
    for (int j = 0; j < 5; j++) mac_split(col[i + j], col[i + j + 1], a.v[i], b.v[j]);
  Fd<F> r;
  fd_reduce_cols<F, 1>(r.v, col);
  return r;
}

// r = a^2 / 2^260: the 10 off-diagonal products are formed once with a doubled operand (|a_i| <= 2^51 required)
template <class F>
MIRA_FD_HD Fd<F> fd_sqr(const Fd<F>& a) {
  uint64_t col[11];
#pragma unroll
  for (int k = 0; k < 11; k++)      // i <= j terms of the square plus the 25 terms of the reduction
    col[k] = 0ull - ((uint64_t)(cnt5_tri(k) + cnt5(k)) * K2 + (uint64_t)(cnt5_tri(k - 1) + cnt5(k - 1)) * K1);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    mac_split(col[2 * i], col[2 * i + 1], a.v[i], a.v[i]);
    double a2 = a.v[i] + a.v[i];
#pragma unroll
    for (int j = i + 1; j < 5; j++) mac_split(col[i + j], col[i + j + 1], a2, a.v[j]);
  }
  Fd<F> r;
  fd_reduce_cols<F, 1>(r.v, col);
  return r;
}

// r = (a*b + c*d) / 2^260 under one reduction
template <class F>
MIRA_FD_HD Fd<F> fd_mul_add_mul(const Fd<F>& a, const Fd<F>& b, const Fd<F>& c, const Fd<F>& d) {
  uint64_t col[11];
  fd_init_cols<2>(col);
#pragma unroll
  for (int i = 0; i < 5; i++)
#pragma unroll
    for (int j = 0; j < 5; j++) {
      mac_split(col[i + j], col[i + j + 1], a.v[i], b.v[j]);
      mac_split(col[i + j], col[i + j + 1], c.v[i], d.v[j]);
    }
  Fd<F> r;
  fd_reduce_cols<F, 2>(r.v, col);
  return r;
}

template <class F> MIRA_FD_HD Fd<F> fd_add(const Fd<F>& a, const Fd<F>& b) {
  Fd<F> r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.v[i] = a.v[i] + b.v[i];
  return r;
}
template <class F> MIRA_FD_HD Fd<F> fd_sub(const Fd<F>& a, const Fd<F>& b) {
  Fd<F> r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.v[i] = a.v[i] - b.v[i];
  return r;
}
template <class F> MIRA_FD_HD Fd<F> fd_neg(const Fd<F>& a) {
  Fd<F> r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.v[i] = -a.v[i];
  return r;
}

// carry normalisation: same value, limbs 0..3 back in [-2^51, 2^51]  (input limbs: integers of magnitude < 2^100).
// (x + C1) - C1 rounds x to the nearest multiple of 2^52 (ulp of the binade of C1); the rest is exact.
template <class F> MIRA_FD_HD Fd<F> fd_norm(const Fd<F>& a) {
  Fd<F> r;
  double carry = 0.0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double x = a.v[i] + carry;
    double c = (x + C1) - C1;
    r.v[i] = x - c;
    carry = c * 0x1p-52;
  }
  r.v[4] = a.v[4] + carry;
  return r;
}

// ---- conversions from / to the reference layout (Fe<F>: 8 x u32 LE, Montgomery R = 2^256, canonical < p)
// raw 256-bit integer -> balanced limbs (value preserved)
template <class F> MIRA_FD_HD Fd<F> fd_from_u256(const uint32_t (&w)[8]) {
  uint64_t x[4];
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
  const uint64_t M = (1ull << 52) - 1;
  int64_t l[5];
  l[0] = (int64_t)(x[0] & M);
  l[1] = (int64_t)(((x[0] >> 52) | (x[1] << 12)) & M);
  l[2] = (int64_t)(((x[1] >> 40) | (x[2] << 24)) & M);
  l[3] = (int64_t)(((x[2] >> 28) | (x[3] << 36)) & M);
  l[4] = (int64_t)(x[3] >> 16);
  Fd<F> r;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int64_t c = (l[i] + (1ll << 51)) >> 52;
    l[i] -= c << 52;
    l[i + 1] += c;
    r.v[i] = s52_to_double(l[i]);
  }
  r.v[4] = s52_to_double(l[4]);
  return r;
}
// x*2^256 (canonical, < p) -> x*2^260 mod p with |r| < 1.3p: one product by 2^264 mod p
template <class F> MIRA_FD_HD Fd<F> fd_from_std(const Fe<F>& a) {
  Fd<F> k;
#pragma unroll
  for (int i = 0; i < 5; i++) k.v[i] = P52<F>::r264(i);
  return fd_mul(fd_from_u256<F>(a.v), k);
}

// balanced limbs (|value| < 2^259) -> canonical representative in [0, p) as 8 x u32
template <class F> MIRA_FD_HD void fd_to_u256_mod(uint32_t (&w)[8], const Fd<F>& a) {
  // to integers; accumulate into a signed 5 x 52 -> 4 x 64 two's complement value, then fix the range with +-p steps
  int64_t l[5];
#pragma unroll
  for (int i = 0; i < 5; i++) l[i] = (int64_t)a.v[i];       // exact integers
  // propagate into non-negative 52-bit limbs with a signed top
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int64_t c = l[i] >> 52;
    l[i] -= c << 52;
    l[i + 1] += c;
  }
  // l[4] signed, |l[4]| < 2^51.  Value v = sum l[i] 2^(52 i).  Bring into [0, p): add/subtract p a few times.
  uint64_t x[4];
  int64_t top;        // bits 256.. of the two's complement value (sign extension)
  x[0] = (uint64_t)l[0] | ((uint64_t)l[1] << 52);
  x[1] = ((uint64_t)l[1] >> 12) | ((uint64_t)l[2] << 40);
  x[2] = ((uint64_t)l[2] >> 24) | ((uint64_t)l[3] << 28);
  x[3] = ((uint64_t)l[3] >> 36) | ((uint64_t)l[4] << 16);
  top = l[4] >> 48;
  uint64_t pm[4];
#pragma unroll
  for (int i = 0; i < 4; i++) pm[i] = (uint64_t)FieldParams<F>::mod(2 * i) | ((uint64_t)FieldParams<F>::mod(2 * i + 1) << 32);
  for (int it = 0; it < 40; it++) {
    bool neg = top < 0;
    bool ge = false;
    if (!neg) {
      ge = top > 0;
      if (!ge) {
        ge = true;
        for (int i = 3; i >= 0; i--) {
          if (x[i] != pm[i]) { ge = x[i] > pm[i]; break; }
        }
      }
    }
    if (!neg && !ge) break;
    if (neg) {           // += p
      uint64_t c = 0;
      for (int i = 0; i < 4; i++) {
        uint64_t s1 = x[i] + pm[i], c1 = s1 < x[i];
        uint64_t s2 = s1 + c, c2 = s2 < s1;
        x[i] = s2; c = c1 | c2;
      }
      top += (int64_t)c;
    } else {             // -= p
      uint64_t br = 0;
      for (int i = 0; i < 4; i++) {
        uint64_t d = x[i] - pm[i], b1 = x[i] < pm[i];
        uint64_t d2 = d - br, b2 = d < br;
        x[i] = d2; br = b1 | b2;
      }
      top -= (int64_t)br;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) { w[2 * i] = (uint32_t)x[i]; w[2 * i + 1] = (uint32_t)(x[i] >> 32); }
}
// x*2^260 (any representative) -> x*2^256 canonical: one product by 2^256 (= limb 4 set to 2^48), then canonicalise
template <class F> MIRA_FD_HD Fe<F> fd_to_std(const Fd<F>& a) {
  Fd<F> k;
  k.v[0] = k.v[1] = k.v[2] = k.v[3] = 0.0;
  k.v[4] = 0x1p48;
  Fd<F> t = fd_mul(a, k);
  Fe<F> r;
  fd_to_u256_mod<F>(r.v, t);
  return r;
}

}  // namespace fp52
}  // namespace mira
