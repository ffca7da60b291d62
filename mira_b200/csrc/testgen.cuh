// testgen.cuh — device-side deterministic synthetic inputs + element-wise unit-test kernels.
// Definitions are shared bit-for-bit with oracle/mira_oracle.c (oracle_gen_scalars / oracle_gen_bases)
// and tests/pyref.py, so GPU-generated 2^24..2^26-point workloads can be spot-checked on the CPU.
#pragma once
#include "curve.cuh"

namespace mira {

__host__ __device__ __forceinline__ uint64_t sm64_word(uint64_t seed, uint64_t k) {
  uint64_t z = seed + (k + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

// canonical pseudo-random element i of stream `seed`: 254 random bits, minus the modulus if >= modulus
template <class F> __device__ __forceinline__ Fe<F> gen_canonical(uint64_t seed, uint64_t i) {
  Fe<F> c;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint64_t w = sm64_word(seed, 4 * i + k);
    if (k == 3) w &= 0x3FFFFFFFFFFFFFFFULL;
    c.v[2 * k] = (uint32_t)w;
    c.v[2 * k + 1] = (uint32_t)(w >> 32);
  }
  bool ge = true;   // c >= MOD ?
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    uint32_t m = FieldParams<F>::mod(k);
    if (c.v[k] != m) { ge = c.v[k] > m; break; }
  }
  if (ge) {
    uint64_t br = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uint64_t d = (uint64_t)c.v[k] - FieldParams<F>::mod(k) - br;
      c.v[k] = (uint32_t)d;
      br = (d >> 32) & 1;
    }
  }
  return c;
}

template <class SF>
__global__ void k_gen_scalars(uint64_t seed, uint64_t first, uint64_t n, int dist, void* out) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint64_t i = first + k;
  Fe<SF> c = gen_canonical<SF>(seed, i);
  if (dist == 1) {
    uint64_t sel = sm64_word(seed ^ 0x5EEDULL, i) % 100;
    if (sel < 60) c = fe_zero<SF>();
    else if (sel < 85) { c.v[0] &= 1u; for (int j = 1; j < 8; j++) c.v[j] = 0; }
    else if (sel < 95) { for (int j = 1; j < 8; j++) c.v[j] = 0; }
  }
  fe_store<SF>(reinterpret_cast<char*>(out) + k * 32, fe_from_canonical(c));
}

template <class CF> __device__ __forceinline__ Affine<CF> curve_generator();
template <> __device__ __forceinline__ Affine<FqTag> curve_generator<FqTag>() {   // BN254 G1: (1, 2)
  Affine<FqTag> g;
  g.x = fe_one<FqTag>();
  g.y = fe_dbl(g.x);
  return g;
}
template <> __device__ __forceinline__ Affine<FrTag> curve_generator<FrTag>() {   // Grumpkin: (1, sqrt(-16))
  Affine<FrTag> g;
  g.x = fe_one<FrTag>();
  Fe<FrTag> y;   // canonical 0x2cf135e7506a45d632d270d45f1181294833fc48d823f272c
  y.v[0] = 0x823f272cu; y.v[1] = 0x833fc48du; y.v[2] = 0xf1181294u; y.v[3] = 0x2d270d45u;
  y.v[4] = 0x06a45d63u; y.v[5] = 0xcf135e75u; y.v[6] = 0x00000002u; y.v[7] = 0u;
  g.y = fe_from_canonical(y);
  return g;
}

// table[w*256 + d] = d * 2^(8w) * G
template <class CF>
__global__ void k_gen_table(void* table) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 32 * 256) return;
  uint32_t w = t >> 8, d = t & 255;
  Xyzz<CF> base = xyzz_from_affine(curve_generator<CF>());
  for (uint32_t k = 0; k < 8 * w; k++) base = xyzz_dbl(base);
  Xyzz<CF> r = xyzz_mul_u32(base, d);
  aff_store<CF>(reinterpret_cast<char*>(table) + (size_t)t * 64, xyzz_to_affine(r));
}

template <class CF, class SF>
__global__ void __launch_bounds__(128) k_gen_bases(uint64_t seed, uint64_t first, uint64_t n, const void* table, void* out) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  Fe<SF> c = gen_canonical<SF>(seed, first + k);
  Xyzz<CF> acc = xyzz_identity<CF>();
  for (int w = 0; w < 32; w++) {
    uint32_t d = (c.v[w >> 2] >> (8 * (w & 3))) & 0xffu;
    if (d) {
      Affine<CF> p = aff_load<CF>(reinterpret_cast<const char*>(table) + (size_t)(w * 256 + d) * 64);
      xyzz_madd(acc, p);
    }
  }
  aff_store<CF>(reinterpret_cast<char*>(out) + k * 64, xyzz_to_affine(acc));
}

// op: 0 mul, 1 add, 2 sub, 3 sqr(a), 4 inv(a), 5 to_canonical(a), 6 from_canonical(a)
template <class F>
__global__ void k_test_field(int op, const void* a_, const void* b_, uint64_t n, void* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<F> a = fe_load<F>(reinterpret_cast<const char*>(a_) + i * 32);
  Fe<F> b = fe_load<F>(reinterpret_cast<const char*>(b_) + i * 32);
  Fe<F> r;
  switch (op) {
    case 0: r = fe_mul(a, b); break;
    case 1: r = fe_add(a, b); break;
    case 2: r = fe_sub(a, b); break;
    case 3: r = fe_sqr(a); break;
    case 4: r = fe_inv(a); break;
    case 5: r = fe_to_canonical(a); break;
    default: r = fe_from_canonical(a); break;
  }
  fe_store<F>(reinterpret_cast<char*>(out) + i * 32, r);
}

// op: 0 p+q (XYZZ mixed add), 1 p+q (XYZZ full add), 2 2p, 3 k*p with k = low 32 bits of q's first word
template <class CF>
__global__ void k_test_point(int op, const void* p_, const void* q_, uint64_t n, void* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<CF> p = aff_load<CF>(reinterpret_cast<const char*>(p_) + i * 64);
  Affine<CF> q = aff_load<CF>(reinterpret_cast<const char*>(q_) + i * 64);
  Xyzz<CF> r = xyzz_from_affine(p);
  switch (op) {
    case 0: xyzz_madd(r, q); break;
    case 1: { Xyzz<CF> t = xyzz_from_affine(q); t = xyzz_dbl(t); xyzz_add(t, xyzz_from_affine(q)); /* t = 3q, non-trivial ZZ */
              Xyzz<CF> m2 = xyzz_dbl(xyzz_from_affine(q)); Affine<CF> nq = q; if (!fe_is_zero(nq.y)) nq.y = fe_neg(nq.y);
              Xyzz<CF> neg2 = xyzz_dbl(xyzz_from_affine(nq));   // -2q with non-trivial ZZ
              xyzz_add(r, t); xyzz_add(r, neg2); (void)m2; break; }   // p + 3q - 2q = p + q through full adds
    case 2: r = xyzz_dbl(r); break;
    default: r = xyzz_mul_u32(r, q.x.v[0]); break;
  }
  aff_store<CF>(reinterpret_cast<char*>(out) + i * 64, xyzz_to_affine(r));
}

}  // namespace mira
