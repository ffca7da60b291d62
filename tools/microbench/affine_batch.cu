// Bucket-accumulation arithmetic A/B on sm_100a: XYZZ mixed addition (what k_accumulate runs, 8M+2S per point)
// against batched affine addition (Montgomery's trick: 5M+1S per addition plus one inversion per batch).
// Both kernels gather random 64-byte affine points from a 2^20-point table, as the accumulation does.
//
// Build: nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -I../../mira_b200/csrc -o affine_batch affine_batch.cu
// Output: one JSON line per variant with ns per group addition (whole GPU).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "curve.cuh"
#include "testgen.cuh"

using namespace mira;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr uint32_t LOG_POINTS = 20, N_POINTS = 1u << LOG_POINTS;

__device__ __forceinline__ uint32_t pick(uint32_t t, uint32_t e) {
  uint32_t h = (t * 0x9E3779B1u) ^ (e * 0x85EBCA77u);
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
  return h & (N_POINTS - 1);
}

// A: one XYZZ accumulator per thread, L mixed additions
template <class CF>
__global__ void __launch_bounds__(128) k_xyzz(const void* __restrict__ table, int L, void* __restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  Xyzz<CF> acc = xyzz_identity<CF>();
  for (int e = 0; e < L; e++) {
    Affine<CF> p = aff_load<CF>(reinterpret_cast<const char*>(table) + (size_t)pick(t, e) * 64);
    xyzz_madd(acc, p);
  }
  xyzz_store<CF>(reinterpret_cast<char*>(out) + (size_t)t * 128, acc);
}

// B: K independent affine additions per thread sharing one inversion (per-thread Montgomery trick)
template <class CF, int K>
__global__ void __launch_bounds__(128) k_affine(const void* __restrict__ table, void* __restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  Fe<CF> pref[K];
  const char* tb = reinterpret_cast<const char*>(table);
  Fe<CF> run = fe_one<CF>();
#pragma unroll 1
  for (int j = 0; j < K; j++) {
    Fe<CF> ax = fe_load<CF>(tb + (size_t)pick(t, 2 * j) * 64), bx = fe_load<CF>(tb + (size_t)pick(t, 2 * j + 1) * 64);
    Fe<CF> dx = fe_sub(bx, ax);
    if (fe_is_zero(dx)) dx = fe_one<CF>();
    pref[j] = run;                       // product of the denominators before j
    run = fe_mul(run, dx);
  }
  Fe<CF> inv = fe_inv(run);
#pragma unroll 1
  for (int j = K - 1; j >= 0; j--) {
    Affine<CF> a = aff_load<CF>(tb + (size_t)pick(t, 2 * j) * 64), b = aff_load<CF>(tb + (size_t)pick(t, 2 * j + 1) * 64);
    Fe<CF> dx = fe_sub(b.x, a.x);
    if (fe_is_zero(dx)) dx = fe_one<CF>();
    Fe<CF> idx = fe_mul(inv, pref[j]);
    inv = fe_mul(inv, dx);
    Fe<CF> lam = fe_mul(fe_sub(b.y, a.y), idx);
    Affine<CF> r;
    r.x = fe_sub(fe_sub(fe_sqr(lam), a.x), b.x);
    r.y = fe_sub(fe_mul(lam, fe_sub(a.x, r.x)), a.y);
    aff_store<CF>(reinterpret_cast<char*>(out) + ((size_t)t * K + j) * 64, r);
  }
}

// C: as B, but the 32 lanes of a warp share ONE inversion: warp product by shuffles (5 steps up, 5 down)
template <class CF>
__device__ __forceinline__ Fe<CF> fe_shfl(const Fe<CF>& a, int src) {
  Fe<CF> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
  return r;
}
template <class CF, int K>
__global__ void __launch_bounds__(128) k_affine_warp(const void* __restrict__ table, void* __restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  Fe<CF> pref[K];
  const char* tb = reinterpret_cast<const char*>(table);
  Fe<CF> run = fe_one<CF>();
#pragma unroll 1
  for (int j = 0; j < K; j++) {
    Fe<CF> ax = fe_load<CF>(tb + (size_t)pick(t, 2 * j) * 64), bx = fe_load<CF>(tb + (size_t)pick(t, 2 * j + 1) * 64);
    Fe<CF> dx = fe_sub(bx, ax);
    if (fe_is_zero(dx)) dx = fe_one<CF>();
    pref[j] = run;
    run = fe_mul(run, dx);
  }
  // inclusive prefix product over lanes (Hillis-Steele), exclusive part kept for the way down
  Fe<CF> incl = run;
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    Fe<CF> o = fe_shfl(incl, lane - d < 0 ? lane : lane - d);
    Fe<CF> m = fe_mul(incl, o);
    if (lane >= d) incl = m;
  }
  Fe<CF> total_inv;
  {
    Fe<CF> total = fe_shfl(incl, 31);
    Fe<CF> ti = fe_one<CF>();
    if (lane == 0) ti = fe_inv(total);
    total_inv = fe_shfl(ti, 0);
  }
  // suffix product over lanes > lane
  Fe<CF> suf = run;
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    Fe<CF> o = fe_shfl(suf, lane + d > 31 ? lane : lane + d);
    Fe<CF> m = fe_mul(suf, o);
    if (lane + d <= 31) suf = m;
  }
  // 1/run_lane = total_inv * (product of lanes < lane) * (product of lanes > lane)
  Fe<CF> excl = fe_shfl(incl, lane ? lane - 1 : 0);
  Fe<CF> after = fe_shfl(suf, lane < 31 ? lane + 1 : 31);
  Fe<CF> inv = total_inv;
  if (lane) inv = fe_mul(inv, excl);
  if (lane < 31) inv = fe_mul(inv, after);
#pragma unroll 1
  for (int j = K - 1; j >= 0; j--) {
    Affine<CF> a = aff_load<CF>(tb + (size_t)pick(t, 2 * j) * 64), b = aff_load<CF>(tb + (size_t)pick(t, 2 * j + 1) * 64);
    Fe<CF> dx = fe_sub(b.x, a.x);
    if (fe_is_zero(dx)) dx = fe_one<CF>();
    Fe<CF> idx = fe_mul(inv, pref[j]);
    inv = fe_mul(inv, dx);
    Fe<CF> lam = fe_mul(fe_sub(b.y, a.y), idx);
    Affine<CF> r;
    r.x = fe_sub(fe_sub(fe_sqr(lam), a.x), b.x);
    r.y = fe_sub(fe_mul(lam, fe_sub(a.x, r.x)), a.y);
    aff_store<CF>(reinterpret_cast<char*>(out) + ((size_t)t * K + j) * 64, r);
  }
}

// correctness: r == a + b through the XYZZ formulas, for thread t's K pairs
template <class CF, int K>
__global__ void k_check(const void* __restrict__ table, const void* __restrict__ out, uint32_t threads, unsigned long long* bad) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= threads) return;
  const char* tb = reinterpret_cast<const char*>(table);
  for (int j = 0; j < K; j++) {
    Affine<CF> a = aff_load<CF>(tb + (size_t)pick(t, 2 * j) * 64), b = aff_load<CF>(tb + (size_t)pick(t, 2 * j + 1) * 64);
    Xyzz<CF> s = xyzz_from_affine(a);
    xyzz_madd(s, b);
    Affine<CF> want = xyzz_to_affine(s);
    Affine<CF> got = aff_load<CF>(reinterpret_cast<const char*>(out) + ((size_t)t * K + j) * 64);
    if (!fe_eq(want.x, got.x) || !fe_eq(want.y, got.y)) atomicAdd(bad, 1ull);
  }
}

template <class Fn>
static float time_ms(Fn fn, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  fn(); fn();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; i++) fn();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

template <int K>
static void run_affine(const void* pts, void* out, unsigned long long* d_bad, uint32_t threads, double xyzz_ns) {
  using CF = FqTag;
  uint32_t th = threads * 32 / (2 * K) ;          // keep total additions comparable
  th = th / 128 * 128;
  double adds = (double)th * K;
  float ms = time_ms([&] { k_affine<CF, K><<<th / 128, 128>>>(pts, out); });
  CK(cudaMemset(d_bad, 0, 8));
  k_check<CF, K><<<64, 128>>>(pts, out, 8192, d_bad);
  unsigned long long bad; CK(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
  printf("{\"variant\": \"affine, per-thread inversion\", \"K\": %d, \"ms\": %.3f, \"ns_per_add\": %.4f, \"vs_xyzz\": %.3f, \"mismatches\": %llu}\n", K, ms,
         ms * 1e6 / adds, xyzz_ns / (ms * 1e6 / adds), bad);
  ms = time_ms([&] { k_affine_warp<CF, K><<<th / 128, 128>>>(pts, out); });
  CK(cudaMemset(d_bad, 0, 8));
  k_check<CF, K><<<64, 128>>>(pts, out, 8192, d_bad);
  CK(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
  printf("{\"variant\": \"affine, per-warp inversion\", \"K\": %d, \"ms\": %.3f, \"ns_per_add\": %.4f, \"vs_xyzz\": %.3f, \"mismatches\": %llu}\n", K, ms,
         ms * 1e6 / adds, xyzz_ns / (ms * 1e6 / adds), bad);
  fflush(stdout);
}

int main() {
  using CF = FqTag;
  void *table, *pts, *out;
  unsigned long long* d_bad;
  CK(cudaMalloc(&table, 32 * 256 * 64));
  CK(cudaMalloc(&pts, (size_t)N_POINTS * 64));
  const uint32_t threads = 148 * 4 * 128 * 4;       // 4 waves of 4 blocks per SM
  CK(cudaMalloc(&out, (size_t)threads * 32 * 64 + (1 << 20)));
  CK(cudaMalloc(&d_bad, 8));
  k_gen_table<CF><<<64, 128>>>(table);
  k_gen_bases<CF, FrTag><<<N_POINTS / 128, 128>>>(0x4D495241ull, 0, N_POINTS, table, pts);
  CK(cudaDeviceSynchronize());
  const int L = 32;
  float ms = time_ms([&] { k_xyzz<CF><<<threads / 128, 128>>>(pts, L, out); });
  double xyzz_ns = ms * 1e6 / ((double)threads * L);
  printf("{\"variant\": \"xyzz madd (k_accumulate's arithmetic)\", \"L\": %d, \"ms\": %.3f, \"ns_per_add\": %.4f}\n", L, ms, xyzz_ns);
  fflush(stdout);
  run_affine<8>(pts, out, d_bad, threads, xyzz_ns);
  run_affine<16>(pts, out, d_bad, threads, xyzz_ns);
  run_affine<32>(pts, out, d_bad, threads, xyzz_ns);
  return 0;
}
