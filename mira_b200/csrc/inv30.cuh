// inv30.cuh — modular inversion by batched division steps ("safegcd", Bernstein & Yang 2019) on signed 30-bit limbs.
//
// Used where every lane of a warp inverts at once (the thread-local pair pre-addition of k_pair_up, the
// fixed-base table build): the control flow is the same for every input, and the cost is ~14,000 instructions per
// inversion instead of the ~115,000 of round 1's bit-by-bit uniform binary GCD (which worked on four 256-bit numbers
// per bit; this works on two 32-bit words per bit and touches the long numbers once per 30 bits).  It is what
// fe_inv_uniform (field.cuh) runs.
//
// The algorithm is the published one (a divstep maps (delta, f, g) with f odd to (1 - delta, g, (g - f) / 2) if
// delta > 0 and g is odd, and to (1 + delta, f, (g + (g odd) f) / 2) otherwise; 590 steps bring any 256-bit g to 0
// when delta starts at 1/2).  30 steps are run on the low words of f and g alone and collected into a 2 x 2
// integer matrix t (scaled by 2^30); the matrix is then applied to the full f, g (exact division by 2^30) and to
// the cofactors d, e (division by 2^30 modulo M), which satisfy d * x = f and e * x = g (mod M) throughout.  At
// the end g = 0, f = +-1 and d = +-1/x.  Replaces, for this use, Field::invert of halo2curves (reached from
// Curve::to_affine at /root/reference/src/commitment.rs:80); plain integer C++, host- and device-compilable so
// that tests/cpp can check it against big-integer arithmetic without a GPU.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace mira {
namespace inv30 {

constexpr uint32_t M30 = 0x3fffffffu;

struct Limbs {      // value = sum v[i] * 2^(30 i); limbs 0..7 in [0, 2^30), limb 8 signed
  int32_t v[9];
};

// 8 x 32-bit little-endian words (value < 2^256) -> 9 x 30-bit limbs
__host__ __device__ inline void from_words(Limbs& r, const uint32_t (&a)[8]) {
  r.v[0] = (int32_t)(a[0] & M30);
  r.v[1] = (int32_t)(((a[0] >> 30) | (a[1] << 2)) & M30);
  r.v[2] = (int32_t)(((a[1] >> 28) | (a[2] << 4)) & M30);
  r.v[3] = (int32_t)(((a[2] >> 26) | (a[3] << 6)) & M30);
  r.v[4] = (int32_t)(((a[3] >> 24) | (a[4] << 8)) & M30);
  r.v[5] = (int32_t)(((a[4] >> 22) | (a[5] << 10)) & M30);
  r.v[6] = (int32_t)(((a[5] >> 20) | (a[6] << 12)) & M30);
  r.v[7] = (int32_t)(((a[6] >> 18) | (a[7] << 14)) & M30);
  r.v[8] = (int32_t)(a[7] >> 16);
}
// non-negative normalised limbs (value < 2^256) -> words
__host__ __device__ inline void to_words(uint32_t (&a)[8], const Limbs& r) {
  const uint32_t* v = reinterpret_cast<const uint32_t*>(r.v);
  a[0] = v[0] | (v[1] << 30);
  a[1] = (v[1] >> 2) | (v[2] << 28);
  a[2] = (v[2] >> 4) | (v[3] << 26);
  a[3] = (v[3] >> 6) | (v[4] << 24);
  a[4] = (v[4] >> 8) | (v[5] << 22);
  a[5] = (v[5] >> 10) | (v[6] << 20);
  a[6] = (v[6] >> 12) | (v[7] << 18);
  a[7] = (v[7] >> 14) | (v[8] << 16);
}

struct Trans {      // 2^30 * (f', g') = (u f + v g, q f + r g)
  int32_t u, v, q, r;
};

// 30 division steps on the low words; zeta = -(delta + 1/2).  Branch-free.
__host__ __device__ inline int32_t divsteps30(int32_t zeta, uint32_t f0, uint32_t g0, Trans& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
  uint32_t f = f0, g = g0;
#pragma unroll 6
  for (int i = 0; i < 30; i++) {
    uint32_t c1 = (uint32_t)(zeta >> 31);          // all ones if delta > 0
    const uint32_t c2 = 0u - (g & 1u);             // all ones if g is odd
    const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;     // -f, -u, -v when delta > 0
    g += x & c2;
    q += y & c2;
    r += z & c2;
    c1 &= c2;                                      // the swap happens
    zeta = (int32_t)((uint32_t)zeta ^ c1) - 1;
    f += g & c1;
    u += q & c1;
    v += r & c1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
  return zeta;
}

// (f, g) <- (u f + v g, q f + r g) / 2^30, exactly
__host__ __device__ inline void update_fg(Limbs& f, Limbs& g, const Trans& t) {
  const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
  int64_t cf = u * f.v[0] + v * g.v[0];
  int64_t cg = q * f.v[0] + r * g.v[0];
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    cf += u * f.v[i] + v * g.v[i];
    cg += q * f.v[i] + r * g.v[i];
    f.v[i - 1] = (int32_t)((uint32_t)cf & M30);
    g.v[i - 1] = (int32_t)((uint32_t)cg & M30);
    cf >>= 30;
    cg >>= 30;
  }
  f.v[8] = (int32_t)cf;
  g.v[8] = (int32_t)cg;
}

// (d, e) <- (u d + v e, q d + r e) / 2^30 modulo M; d, e stay in (-2M, M).  MOD = M's limbs, MINV = 1/M mod 2^30.
template <class MOD>
__host__ __device__ inline void update_de(Limbs& d, Limbs& e, const Trans& t) {
  const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
  const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;        // sign masks
  int32_t md = (t.u & sd) + (t.v & se);                        // one M per negative input keeps the range
  int32_t me = (t.q & sd) + (t.r & se);
  int64_t cd = u * d.v[0] + v * e.v[0];
  int64_t ce = q * d.v[0] + r * e.v[0];
  md -= (int32_t)((MOD::minv30() * (uint32_t)cd + (uint32_t)md) & M30);      // makes the low 30 bits vanish
  me -= (int32_t)((MOD::minv30() * (uint32_t)ce + (uint32_t)me) & M30);
  cd += (int64_t)MOD::limb(0) * md;
  ce += (int64_t)MOD::limb(0) * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    cd += u * d.v[i] + v * e.v[i] + (int64_t)MOD::limb(i) * md;
    ce += q * d.v[i] + r * e.v[i] + (int64_t)MOD::limb(i) * me;
    d.v[i - 1] = (int32_t)((uint32_t)cd & M30);
    e.v[i - 1] = (int32_t)((uint32_t)ce & M30);
    cd >>= 30;
    ce >>= 30;
  }
  d.v[8] = (int32_t)cd;
  e.v[8] = (int32_t)ce;
}

// d in (-2M, M), sign = top limb of f (negative: the result is -d) -> d in [0, M), limbs normalised
template <class MOD>
__host__ __device__ inline void normalize(Limbs& d, int32_t sign) {
  int32_t cond_add = d.v[8] >> 31;
  const int32_t cond_neg = sign >> 31;
  int32_t c = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {                  // + M if negative, then negate if f = -1
    int32_t x = d.v[i] + (MOD::limb(i) & cond_add);
    x = (x ^ cond_neg) - cond_neg;
    x += c;
    c = x >> 30;
    d.v[i] = i < 8 ? (x & (int32_t)M30) : x;
  }
  cond_add = d.v[8] >> 31;                       // may be negative again (by less than M): + M once more
  c = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int32_t x = d.v[i] + (MOD::limb(i) & cond_add) + c;
    c = x >> 30;
    d.v[i] = i < 8 ? (x & (int32_t)M30) : x;
  }
}

// r = 1 / a mod M for a in [0, M) given as 8 words; 0 -> 0.  MOD::word(i): M's 32-bit words.
template <class MOD>
__host__ __device__ inline void modinv(uint32_t (&r)[8], const uint32_t (&a)[8]) {
  Limbs d, e, f, g;
  uint32_t mw[8];
#pragma unroll
  for (int i = 0; i < 8; i++) mw[i] = MOD::word(i);
  from_words(f, mw);
  from_words(g, a);
#pragma unroll
  for (int i = 0; i < 9; i++) { d.v[i] = 0; e.v[i] = 0; }
  e.v[0] = 1;
  int32_t zeta = -1;
#pragma unroll 1
  for (int it = 0; it < 20; it++) {              // 600 >= 590 division steps
    Trans t;
    zeta = divsteps30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
    update_de<MOD>(d, e, t);
    update_fg(f, g, t);
  }
  normalize<MOD>(d, f.v[8]);
  to_words(r, d);
}

}  // namespace inv30
}  // namespace mira
