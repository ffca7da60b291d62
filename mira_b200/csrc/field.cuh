// field.cuh — BN254 Fq / Fr arithmetic on 8 x 32-bit limbs, Montgomery form with R = 2^256.
//
// Replaces (device side) halo2curves' 4x64-limb `Fq` / `Fr` that the reference reaches through
// halo2_proofs (src/lib.rs:24-27, src/commitment.rs:11).  Byte layout is identical: 32 bytes
// little-endian of value*2^256 mod m, so reference buffers are used without conversion.
//
// The multiplier bodies are generated (tools/gen_field_ptx.py -> field_gen.cuh) as inline PTX
// whose mad.lo.cc/madc.hi.cc pairs ptxas fuses into IMAD.WIDE.U32[.X]; modulus limbs are
// immediates.  B200 measured issue rate for IMAD.WIDE.U32 is 32 lanes/clk/SM (profiles/imad_peak.json,
// r02_modmul_ncu.txt), which is the roofline denominator of every kernel built on this.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "field_gen.cuh"
#include "inv30.cuh"

// Multiplier: the interleaved word-serial product of field_gen.cuh (128 wide MACs + 8 IMAD per product), with
// MIRA_DUAL_MUL a*b +- c*d as ONE interleaved pass under one reduction (192 wide MACs instead of 256).
// A second generator (separated 16-limb product, optional Karatsuba level, dedicated square, lazy wide subtraction:
// fewer wide MACs) was measured SLOWER on B200 in round 1 (k_accumulate at 2^24, ms: this one 34.0 | separated
// schoolbook + lazy 36.7 | Karatsuba + lazy 39.3 | Karatsuba 41.6: the saved MACs came back as IADD3/MOV traffic and
// registers, 118 -> 158) and has been removed from the tree (git history: field_gen_v2.cuh, tools/gen_field_ptx_v2.py);
// an FP64-pipe product was measured in round 2 and is no faster either (tools/microbench/fp52.cuh).
#ifndef MIRA_DUAL_MUL
#define MIRA_DUAL_MUL 1
#endif

namespace mira {

struct FqTag {};   // modulus p (BN254 base field;  Grumpkin scalar field)
struct FrTag {};   // modulus r (BN254 scalar field; Grumpkin base field)

template <class F> struct FieldParams;
template <> struct FieldParams<FqTag> {
  // p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
  __host__ __device__ static constexpr uint32_t mod(int i) {
    constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return v[i];
  }
  // R mod p  (Montgomery form of 1)
  __host__ __device__ static constexpr uint32_t one(int i) {
    constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return v[i];
  }
  // R^2 mod p
  __host__ __device__ static constexpr uint32_t r2(int i) {
    constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return v[i];
  }
};
template <> struct FieldParams<FrTag> {
  // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
  __host__ __device__ static constexpr uint32_t mod(int i) {
    constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return v[i];
  }
  __host__ __device__ static constexpr uint32_t one(int i) {
    constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return v[i];
  }
  __host__ __device__ static constexpr uint32_t r2(int i) {
    constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return v[i];
  }
};

// A field element in registers.  F selects the modulus at compile time.
template <class F>
struct Fe {
  uint32_t v[8];
};

template <class F> __device__ __forceinline__ Fe<F> fe_zero() {
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_one() {
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = FieldParams<F>::one(i);
  return r;
}
template <class F> __device__ __forceinline__ bool fe_is_zero(const Fe<F>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i];
  return o == 0;
}
template <class F> __device__ __forceinline__ bool fe_eq(const Fe<F>& a, const Fe<F>& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

// ---- dispatch onto the generated bodies ------------------------------------------------------
__device__ __forceinline__ void mont_mul_raw(FqTag, uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { gen::mont_mul_Fq(r, a, b); }
__device__ __forceinline__ void mont_mul_raw(FrTag, uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { gen::mont_mul_Fr(r, a, b); }
__device__ __forceinline__ void mont_mul2_raw(FqTag, uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&bcd)[24]) { gen::mont_mul2_Fq(r, a, bcd); }
__device__ __forceinline__ void mont_mul2_raw(FrTag, uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&bcd)[24]) { gen::mont_mul2_Fr(r, a, bcd); }
__device__ __forceinline__ void mont_sqr_raw(FqTag, uint32_t (&r)[8], const uint32_t (&a)[8]) { gen::mont_sqr_Fq(r, a); }
__device__ __forceinline__ void mont_sqr_raw(FrTag, uint32_t (&r)[8], const uint32_t (&a)[8]) { gen::mont_sqr_Fr(r, a); }
__device__ __forceinline__ void mont_redc_raw(FqTag, uint32_t (&r)[8], const uint32_t (&a)[8]) { gen::mont_redc_Fq(r, a); }
__device__ __forceinline__ void mont_redc_raw(FrTag, uint32_t (&r)[8], const uint32_t (&a)[8]) { gen::mont_redc_Fr(r, a); }
__device__ __forceinline__ void mod_add_raw(FqTag, uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { gen::mod_add_Fq(r, a, b); }
__device__ __forceinline__ void mod_add_raw(FrTag, uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { gen::mod_add_Fr(r, a, b); }
__device__ __forceinline__ void mod_sub_raw(FqTag, uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { gen::mod_sub_Fq(r, a, b); }
__device__ __forceinline__ void mod_sub_raw(FrTag, uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { gen::mod_sub_Fr(r, a, b); }

template <class F> __device__ __forceinline__ Fe<F> fe_mul(const Fe<F>& a, const Fe<F>& b) {
  Fe<F> r;
  mont_mul_raw(F{}, r.v, a.v, b.v);
  return r;
}
// The same product as an out-of-line function (operands and result in registers: no stack frame).  The group
// operations of the reduction and stitching kernels (xyzz_add, xyzz_dbl: 14 / 9 products each, inlined several times
// per kernel) use it: those kernels run a few hundred warps on long dependent chains, their bodies were 100+ KB of
// straight-line code, and instruction fetch was what they waited for (bucket reduction 1.76 -> 1.59 ms at 2^21 buckets,
// 0.32 -> 0.25 ms at 2^16).  k_accumulate calls three of the mixed addition's ten products this way (curve.cuh); with all
// of them out of line it is 2 % SLOWER (32.3 against 31.7 ms: ptxas no longer overlaps neighbouring products).
template <class F> __device__ __noinline__ Fe<F> fe_mul_call(Fe<F> a, Fe<F> b) {
  Fe<F> r;
  mont_mul_raw(F{}, r.v, a.v, b.v);
  return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_mulc(const Fe<F>& a, const Fe<F>& b) { return fe_mul_call<F>(a, b); }
template <class F> __device__ __noinline__ Fe<F> fe_sqr_call(Fe<F> a) {
  Fe<F> r;
  mont_sqr_raw(F{}, r.v, a.v);
  return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_sqrc(const Fe<F>& a) { return fe_sqr_call<F>(a); }
template <class F> __device__ __forceinline__ Fe<F> fe_sqr(const Fe<F>& a) {
  Fe<F> r;
  mont_sqr_raw(F{}, r.v, a.v);
  return r;
}
// a*b + c*d with ONE reduction
template <class F> __device__ __forceinline__ Fe<F> fe_mul_add_mul(const Fe<F>& a, const Fe<F>& b, const Fe<F>& c, const Fe<F>& d) {
#if MIRA_DUAL_MUL
  uint32_t bcd[24];
#pragma unroll
  for (int i = 0; i < 8; i++) { bcd[i] = b.v[i]; bcd[8 + i] = c.v[i]; bcd[16 + i] = d.v[i]; }
  Fe<F> r;
  mont_mul2_raw(F{}, r.v, a.v, bcd);
  return r;
#else
  return fe_add(fe_mul(a, b), fe_mul(c, d));
#endif
}
// a*b - c*d with ONE reduction
template <class F> __device__ __forceinline__ Fe<F> fe_mul_sub_mul(const Fe<F>& a, const Fe<F>& b, const Fe<F>& c, const Fe<F>& d) {
#if MIRA_DUAL_MUL
  // a*b + (MOD - c)*d: both products accumulate into one interleaved Montgomery pass
  Fe<F> nc = fe_sub(fe_zero<F>(), c);
  uint32_t bcd[24];
#pragma unroll
  for (int i = 0; i < 8; i++) { bcd[i] = b.v[i]; bcd[8 + i] = nc.v[i]; bcd[16 + i] = d.v[i]; }
  Fe<F> r;
  mont_mul2_raw(F{}, r.v, a.v, bcd);
  return r;
#else
  return fe_sub(fe_mul(a, b), fe_mul(c, d));
#endif
}
// ... and out of line, for xyzz_add / xyzz_dbl (see fe_mul_call)
template <class F> __device__ __noinline__ Fe<F> fe_mul_sub_mul_call(Fe<F> a, Fe<F> b, Fe<F> c, Fe<F> d) { return fe_mul_sub_mul(a, b, c, d); }
template <class F> __device__ __forceinline__ Fe<F> fe_mul_sub_mulc(const Fe<F>& a, const Fe<F>& b, const Fe<F>& c, const Fe<F>& d) {
  return fe_mul_sub_mul_call<F>(a, b, c, d);
}
template <class F> __device__ __forceinline__ Fe<F> fe_add(const Fe<F>& a, const Fe<F>& b) {
  Fe<F> r;
  mod_add_raw(F{}, r.v, a.v, b.v);
  return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_sub(const Fe<F>& a, const Fe<F>& b) {
  Fe<F> r;
  mod_sub_raw(F{}, r.v, a.v, b.v);
  return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_dbl(const Fe<F>& a) { return fe_add(a, a); }
template <class F> __device__ __forceinline__ Fe<F> fe_neg(const Fe<F>& a) { return fe_sub(fe_zero<F>(), a); }

// Montgomery -> canonical (PrimeField::to_repr): a / R, i.e. the reduction half of a product alone
// (gen::mont_redc_*: 64 wide MACs instead of the 128 of a product by the raw integer 1).
template <class F> __device__ __forceinline__ Fe<F> fe_to_canonical(const Fe<F>& a) {
  Fe<F> r;
  mont_redc_raw(F{}, r.v, a.v);
  return r;
}
// canonical -> Montgomery
template <class F> __device__ __forceinline__ Fe<F> fe_from_canonical(const Fe<F>& a) {
  Fe<F> r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.v[i] = FieldParams<F>::r2(i);
  return fe_mul(a, r2);
}

// ---- inversion -----------------------------------------------------------------------------------
// Plain 256-bit helpers for the binary extended Euclid below (values < 2^255, so no carry out of limb 7).
__device__ __forceinline__ void u256_shr1(uint32_t (&a)[8]) {
#pragma unroll
  for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
  a[7] >>= 1;
}
__device__ __forceinline__ void u256_add(uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a[i] + b[i];
    a[i] = (uint32_t)c;
    c >>= 32;
  }
}
__device__ __forceinline__ uint32_t u256_sub(uint32_t (&a)[8], const uint32_t (&b)[8]) {   // returns the borrow
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)a[i] - b[i] - br;
    a[i] = (uint32_t)d;
    br = (d >> 32) & 1;
  }
  return (uint32_t)br;
}
__device__ __forceinline__ bool u256_ge(const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return true;
}
__device__ __forceinline__ bool u256_is_one(const uint32_t (&a)[8]) {
  uint32_t o = a[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < 8; i++) o |= a[i];
  return o == 0;
}

// 1/a in Montgomery form; 0 -> 0 (as Field::invert().unwrap_or(0) is used by Curve::to_affine).
// Binary extended Euclid on the raw residue: shifts, adds and subtracts only, i.e. ~500 short iterations on the
// ALU pipe instead of the ~380 dependent Montgomery products of a Fermat ladder (0.6 ms on one thread).
// With A = a*R the loop yields A^{-1} = a^{-1} R^{-1}; two products by R^2 bring it to a^{-1} R.
template <class F> __device__ __noinline__ Fe<F> fe_inv(const Fe<F>& a) {
  if (fe_is_zero(a)) return a;
  uint32_t u[8], v[8], x1[8], x2[8], p[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u[i] = a.v[i];
    p[i] = v[i] = FieldParams<F>::mod(i);
    x1[i] = (i == 0);
    x2[i] = 0;
  }
  while (!u256_is_one(u) && !u256_is_one(v)) {
    while (!(u[0] & 1u)) {
      u256_shr1(u);
      if (x1[0] & 1u) u256_add(x1, p);
      u256_shr1(x1);
    }
    while (!(v[0] & 1u)) {
      u256_shr1(v);
      if (x2[0] & 1u) u256_add(x2, p);
      u256_shr1(x2);
    }
    if (u256_ge(u, v)) {
      u256_sub(u, v);
      if (u256_sub(x1, x2)) u256_add(x1, p);
    } else {
      u256_sub(v, u);
      if (u256_sub(x2, x1)) u256_add(x2, p);
    }
  }
  Fe<F> r, r2;
  const bool from_u = u256_is_one(u);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.v[i] = from_u ? x1[i] : x2[i];
    r2.v[i] = FieldParams<F>::r2(i);
  }
  return fe_mul(fe_mul(r, r2), r2);
}

// The same inverse for code where all 32 lanes of a warp invert at once (fe_inv's data-dependent inner loops would
// serialise there): batched division steps on 30-bit limbs (inv30.cuh), branch-free, ~14,000 instructions per
// inversion.  Round 1's version of this function was a branch-uniform bit-by-bit binary GCD on four 256-bit numbers
// (~115,000 executed instructions per inversion, ncu); it is what made a per-thread inversion unaffordable inside
// the accumulation kernel.
template <class F> struct Inv30Mod {
  __host__ __device__ static constexpr uint32_t word(int i) { return FieldParams<F>::mod(i); }
  __host__ __device__ static constexpr int32_t limb(int i) {
    const int bit = 30 * i, k = bit >> 5, sh = bit & 31;
    const uint64_t lo = FieldParams<F>::mod(k), hi = k + 1 < 8 ? FieldParams<F>::mod(k + 1) : 0u;
    return (int32_t)((((hi << 32) | lo) >> sh) & 0x3fffffffu);
  }
  __host__ __device__ static constexpr uint32_t minv30() {       // 1 / M mod 2^30 by Newton's iteration
    const uint32_t m0 = FieldParams<F>::mod(0);
    uint32_t x = m0;
    for (int i = 0; i < 5; i++) x *= 2u - m0 * x;
    return x & 0x3fffffffu;
  }
};
template <class F> __device__ __noinline__ Fe<F> fe_inv_uniform(const Fe<F>& a) {
  Fe<F> r, r2;
  inv30::modinv<Inv30Mod<F>>(r.v, a.v);          // (a R)^-1 = a^-1 R^-1 as a plain residue; 0 -> 0
#pragma unroll
  for (int i = 0; i < 8; i++) r2.v[i] = FieldParams<F>::r2(i);
  return fe_mul(fe_mul(r, r2), r2);
}

// ---- 128-bit vectorised global-memory access (elements are 32-byte aligned in all our buffers)
template <class F> __device__ __forceinline__ Fe<F> fe_load(const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = __ldg(q), hi = __ldg(q + 1);
  Fe<F> r;
  r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
  r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
  return r;
}
// plain (not read-only-path) load, for memory written earlier by the same kernel
template <class F> __device__ __forceinline__ Fe<F> fe_load_plain(const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = q[0], hi = q[1];
  Fe<F> r;
  r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
  r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
  return r;
}
template <class F> __device__ __forceinline__ void fe_store(void* p, const Fe<F>& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

}  // namespace mira
