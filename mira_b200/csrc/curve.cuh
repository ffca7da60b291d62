// curve.cuh — short-Weierstrass a = 0 group law in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2; identity <=> ZZ == 0), for BN254 G1 (over Fq) and Grumpkin G1 (over Fr).
//
// Replaces the Jacobian `add_assign` / mixed-add bucket arithmetic inside halo2's
// multiexp_serial that the reference calls at src/commitment.rs:80.  Formulas: EFD
// madd-2008-s (8M+2S), add-2008-s (12M+2S), dbl-2008-s-1 / mdbl-2008-s-1.  All exceptional
// cases (identity operand, P == Q, P == -Q) are handled so the result is the group sum for
// every input the reference accepts (duplicate or cancelling key points included).
#pragma once
#include "field.cuh"

// 1 (shipped): three products of the mixed addition are called out of line (xyzz_madd below); 0: all inlined
#ifndef MIRA_MADD_CALLS
#define MIRA_MADD_CALLS 1
#endif

namespace mira {

// Affine point as the reference stores it: {x, y}, 64 bytes, identity = (0, 0).
template <class F>
struct Affine {
  Fe<F> x, y;
};
template <class F>
struct Xyzz {
  Fe<F> x, y, zz, zzz;
};

template <class F> __device__ __forceinline__ bool aff_is_identity(const Affine<F>& p) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= p.x.v[i] | p.y.v[i];
  return o == 0;
}
template <class F> __device__ __forceinline__ bool xyzz_is_identity(const Xyzz<F>& p) { return fe_is_zero(p.zz); }
template <class F> __device__ __forceinline__ Xyzz<F> xyzz_identity() {
  Xyzz<F> r;
  r.x = fe_zero<F>(); r.y = fe_zero<F>(); r.zz = fe_zero<F>(); r.zzz = fe_zero<F>();
  return r;
}
template <class F> __device__ __forceinline__ Xyzz<F> xyzz_from_affine(const Affine<F>& p) {
  Xyzz<F> r;
  if (aff_is_identity(p)) return xyzz_identity<F>();
  r.x = p.x; r.y = p.y; r.zz = fe_one<F>(); r.zzz = fe_one<F>();
  return r;
}

// 64-byte affine load as four 128-bit loads (points are 64-byte aligned in every buffer we own).
template <class F> __device__ __forceinline__ Affine<F> aff_load(const void* p) {
  Affine<F> r;
  r.x = fe_load<F>(p);
  r.y = fe_load<F>(reinterpret_cast<const char*>(p) + 32);
  return r;
}
template <class F> __device__ __forceinline__ void aff_store(void* p, const Affine<F>& a) {
  fe_store<F>(p, a.x);
  fe_store<F>(reinterpret_cast<char*>(p) + 32, a.y);
}
template <class F> __device__ __forceinline__ Xyzz<F> xyzz_load(const void* p) {
  const char* c = reinterpret_cast<const char*>(p);
  Xyzz<F> r;
  r.x = fe_load<F>(c); r.y = fe_load<F>(c + 32); r.zz = fe_load<F>(c + 64); r.zzz = fe_load<F>(c + 96);
  return r;
}
// plain (non read-only-path) load, for shared memory or buffers written by the same kernel
template <class F> __device__ __forceinline__ Xyzz<F> xyzz_load_shared(const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  Xyzz<F> r;
  Fe<F>* f[4] = {&r.x, &r.y, &r.zz, &r.zzz};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint4 lo = q[2 * k], hi = q[2 * k + 1];
    f[k]->v[0] = lo.x; f[k]->v[1] = lo.y; f[k]->v[2] = lo.z; f[k]->v[3] = lo.w;
    f[k]->v[4] = hi.x; f[k]->v[5] = hi.y; f[k]->v[6] = hi.z; f[k]->v[7] = hi.w;
  }
  return r;
}
template <class F> __device__ __forceinline__ void xyzz_store(void* p, const Xyzz<F>& a) {
  char* c = reinterpret_cast<char*>(p);
  fe_store<F>(c, a.x); fe_store<F>(c + 32, a.y); fe_store<F>(c + 64, a.zz); fe_store<F>(c + 96, a.zzz);
}

// 2*P for affine P != identity (mdbl-2008-s-1, a = 0).  Rare path (duplicate points in a bucket).
template <class F> __device__ __noinline__ Xyzz<F> xyzz_dbl_affine(const Affine<F>& p) {
  Xyzz<F> r;
  if (fe_is_zero(p.y)) return xyzz_identity<F>();     // order-2 point cannot exist on these curves; be safe
  Fe<F> u = fe_dbl(p.y);
  Fe<F> v = fe_sqr(u);
  Fe<F> w = fe_mul(u, v);
  Fe<F> s = fe_mul(p.x, v);
  Fe<F> xx = fe_sqr(p.x);
  Fe<F> m = fe_add(fe_dbl(xx), xx);
  r.x = fe_sub(fe_sqr(m), fe_dbl(s));
  r.y = fe_sub(fe_mul(m, fe_sub(s, r.x)), fe_mul(w, p.y));
  r.zz = v;
  r.zzz = w;
  return r;
}
// 2*P for XYZZ P (dbl-2008-s-1, a = 0).
template <class F> __device__ __noinline__ Xyzz<F> xyzz_dbl(const Xyzz<F>& p) {
  if (xyzz_is_identity(p) || fe_is_zero(p.y)) return xyzz_identity<F>();
  Xyzz<F> r;
  Fe<F> u = fe_dbl(p.y);
  Fe<F> v = fe_sqrc(u);
  Fe<F> w = fe_mulc(u, v);
  Fe<F> s = fe_mulc(p.x, v);
  Fe<F> xx = fe_sqrc(p.x);
  Fe<F> m = fe_add(fe_dbl(xx), xx);
  r.x = fe_sub(fe_sqrc(m), fe_dbl(s));
  r.y = fe_sub(fe_mulc(m, fe_sub(s, r.x)), fe_mulc(w, p.y));
  r.zz = fe_mulc(v, p.zz);
  r.zzz = fe_mulc(w, p.zzz);
  return r;
}

// acc += q (q affine, possibly negated by the caller).  madd-2008-s.
// Three of its ten products (Q = X1*PP and the two ZZ updates) are called out of line: k_accumulate's loop body shrinks
// from 37 KB to 26 KB and fetch stalls less, while the products ptxas overlaps with their neighbours stay inlined.
// Measured at 2^24 points, accumulation phase (ms): all inlined 31.66 | ZZ, ZZZ 31.44 | + Q 31.35 | U2, S2, ZZ, ZZZ 31.48 |
// PPP, Q, ZZ, ZZZ 31.48 | all six plain products 32.31.  MIRA_MADD_CALLS=0 restores the fully inlined form.
#ifndef MIRA_MADD_MASK
#define MIRA_MADD_MASK (MIRA_MADD_CALLS ? 400 : 0)      // bit k: product k of the list below goes out of line
#endif
#define MIRA_MM(bit, inl, call) ((MIRA_MADD_MASK & (bit)) ? (call) : (inl))
template <class F> __device__ __forceinline__ void xyzz_madd(Xyzz<F>& acc, const Affine<F>& q) {
  if (aff_is_identity(q)) return;
  if (xyzz_is_identity(acc)) {
    acc.x = q.x; acc.y = q.y; acc.zz = fe_one<F>(); acc.zzz = fe_one<F>();
    return;
  }
  Fe<F> u2 = MIRA_MM(1, fe_mul(q.x, acc.zz), fe_mulc(q.x, acc.zz));
  Fe<F> s2 = MIRA_MM(2, fe_mul(q.y, acc.zzz), fe_mulc(q.y, acc.zzz));
  Fe<F> p = fe_sub(u2, acc.x);
  Fe<F> r = fe_sub(s2, acc.y);
  if (fe_is_zero(p)) {                       // same x: doubling or cancellation (rare)
    if (fe_is_zero(r)) acc = xyzz_dbl_affine(q);
    else acc = xyzz_identity<F>();
    return;
  }
  Fe<F> pp = MIRA_MM(4, fe_sqr(p), fe_sqrc(p));
  Fe<F> ppp = MIRA_MM(8, fe_mul(p, pp), fe_mulc(p, pp));
  Fe<F> qq = MIRA_MM(16, fe_mul(acc.x, pp), fe_mulc(acc.x, pp));
  Fe<F> rr = MIRA_MM(32, fe_sqr(r), fe_sqrc(r));
  Fe<F> x3 = fe_sub(fe_sub(rr, ppp), fe_dbl(qq));
  Fe<F> y3 = MIRA_MM(64, fe_mul_sub_mul(r, fe_sub(qq, x3), acc.y, ppp), fe_mul_sub_mulc(r, fe_sub(qq, x3), acc.y, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = MIRA_MM(128, fe_mul(acc.zz, pp), fe_mulc(acc.zz, pp));
  acc.zzz = MIRA_MM(256, fe_mul(acc.zzz, ppp), fe_mulc(acc.zzz, ppp));
}

// acc += q (both XYZZ).  add-2008-s.
template <class F> __device__ __forceinline__ void xyzz_add(Xyzz<F>& acc, const Xyzz<F>& q) {
  if (xyzz_is_identity(q)) return;
  if (xyzz_is_identity(acc)) { acc = q; return; }
  Fe<F> u1 = fe_mulc(acc.x, q.zz);
  Fe<F> u2 = fe_mulc(q.x, acc.zz);
  Fe<F> s1 = fe_mulc(acc.y, q.zzz);
  Fe<F> s2 = fe_mulc(q.y, acc.zzz);
  Fe<F> p = fe_sub(u2, u1);
  Fe<F> r = fe_sub(s2, s1);
  if (fe_is_zero(p)) {
    if (fe_is_zero(r)) acc = xyzz_dbl(acc);
    else acc = xyzz_identity<F>();
    return;
  }
  Fe<F> pp = fe_sqrc(p);
  Fe<F> ppp = fe_mulc(p, pp);
  Fe<F> qq = fe_mulc(u1, pp);
  Fe<F> x3 = fe_sub(fe_sub(fe_sqrc(r), ppp), fe_dbl(qq));
  Fe<F> y3 = fe_mul_sub_mulc(r, fe_sub(qq, x3), s1, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = fe_mulc(fe_mulc(acc.zz, q.zz), pp);
  acc.zzz = fe_mulc(fe_mulc(acc.zzz, q.zzz), ppp);
}

// Curve::to_affine: identity -> (0, 0); one inversion.
template <class F> __device__ __noinline__ Affine<F> xyzz_to_affine(const Xyzz<F>& p) {
  Affine<F> r;
  if (xyzz_is_identity(p)) { r.x = fe_zero<F>(); r.y = fe_zero<F>(); return r; }
  Fe<F> i = fe_inv(fe_mul(p.zz, p.zzz));       // 1 / (ZZ*ZZZ)
  Fe<F> izz = fe_mul(i, p.zzz);                // 1 / ZZ
  Fe<F> izzz = fe_mul(i, p.zz);                // 1 / ZZZ
  r.x = fe_mul(p.x, izz);
  r.y = fe_mul(p.y, izzz);
  return r;
}

// k * P for a small unsigned k (used by the bucket reduction to weight chunk sums)
template <class F> __device__ __noinline__ Xyzz<F> xyzz_mul_u32(const Xyzz<F>& p, uint32_t k) {
  Xyzz<F> acc = xyzz_identity<F>();
  for (int b = 31 - __clz(k | 1u); b >= 0; b--) {
    acc = xyzz_dbl(acc);
    if ((k >> b) & 1u) xyzz_add(acc, p);
  }
  return acc;
}

}  // namespace mira
