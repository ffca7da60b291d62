// Field-product throughput on B200: the shipped IMAD.WIDE Montgomery product (field.cuh, 8 x 32-bit limbs) against
// the FP64-pipe product (fp52.cuh, 5 x 52-bit signed limbs in doubles), plus the raw issue-rate plateaus of
// IMAD.WIDE.U32 and DFMA that the rooflines divide by (>= 100 ms per test, CUDA events AND
// clock64/globaltimer).
//
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o modmul modmul.cu
// Output: one JSON line per test.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "fp52.cuh"

using namespace mira;
using namespace mira::fp52;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

struct Clk { long long c0, c1; unsigned long long t0, t1; };
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ Fe<FqTag> rnd_fe(uint64_t& s) {
  Fe<FqTag> r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    r.v[i] = (uint32_t)(s >> 32);
  }
  r.v[7] &= 0x1fffffffu;        // < 2^253 < p
  return r;
}

// ---- modular products ------------------------------------------------------------------------------------------
// mode 0: x = x*y   1: x = x^2   2: x = x*y + x*z (one reduction)
template <int C, int MODE>
__global__ void __launch_bounds__(128) k_imad(uint32_t* out, int iters, Clk* clk) {
  uint64_t s = 0x1234567ull + blockIdx.x * 131ull + threadIdx.x;
  Fe<FqTag> x[C], y[C], z[C];
#pragma unroll
  for (int k = 0; k < C; k++) { x[k] = rnd_fe(s); y[k] = rnd_fe(s); z[k] = rnd_fe(s); }
  long long c0 = clock64(); unsigned long long t0 = gtime();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < C; k++) {
      if (MODE == 0) x[k] = fe_mul(x[k], y[k]);
      else if (MODE == 1) x[k] = fe_sqr(x[k]);
      else x[k] = fe_mul_add_mul(x[k], y[k], x[k], z[k]);
    }
  }
  long long c1 = clock64(); unsigned long long t1 = gtime();
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < C; k++)
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[k].v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = Clk{c0, c1, t0, t1};
}

template <int C, int MODE>
__global__ void __launch_bounds__(128) k_dfma(uint32_t* out, int iters, Clk* clk) {
  uint64_t s = 0x1234567ull + blockIdx.x * 131ull + threadIdx.x;
  Fd<FqTag> x[C], y[C], z[C];
#pragma unroll
  for (int k = 0; k < C; k++) { x[k] = fd_from_std(rnd_fe(s)); y[k] = fd_from_std(rnd_fe(s)); z[k] = fd_from_std(rnd_fe(s)); }
  long long c0 = clock64(); unsigned long long t0 = gtime();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < C; k++) {
      if (MODE == 0) x[k] = fd_mul(x[k], y[k]);
      else if (MODE == 1) x[k] = fd_sqr(x[k]);
      else x[k] = fd_mul_add_mul(x[k], y[k], x[k], z[k]);
    }
  }
  long long c1 = clock64(); unsigned long long t1 = gtime();
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < C; k++) {
    Fe<FqTag> r = fd_to_std(x[k]);
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r.v[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = Clk{c0, c1, t0, t1};
}

// ---- heterogeneous warps: NI of the block's 4 warps run the IMAD.WIDE product, the others the FP64-pipe product, at
// the same time on the same SM.  Do the two pipes add up?  (iters_i / iters_d are chosen by the host so that both
// kinds finish together; per-kind clocks come back through clk[2 * block + kind].)
template <int NI>
__global__ void __launch_bounds__(128) k_mix(uint32_t* out, int iters_i, int iters_d, Clk* clk) {
  uint64_t s = 0x1234567ull + blockIdx.x * 131ull + threadIdx.x;
  const int warp = threadIdx.x >> 5;
  uint32_t acc = 0;
  if (warp < NI) {
    Fe<FqTag> x[2], y[2];
#pragma unroll
    for (int k = 0; k < 2; k++) { x[k] = rnd_fe(s); y[k] = rnd_fe(s); }
    long long c0 = clock64(); unsigned long long t0 = gtime();
#pragma unroll 1
    for (int i = 0; i < iters_i; i++) {
#pragma unroll
      for (int k = 0; k < 2; k++) x[k] = fe_mul(x[k], y[k]);
    }
    long long c1 = clock64(); unsigned long long t1 = gtime();
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
      for (int i = 0; i < 8; i++) acc ^= x[k].v[i];
    if (threadIdx.x == 0) clk[2 * blockIdx.x] = Clk{c0, c1, t0, t1};
  } else {
    Fd<FqTag> x[2], y[2];
#pragma unroll
    for (int k = 0; k < 2; k++) { x[k] = fd_from_std(rnd_fe(s)); y[k] = fd_from_std(rnd_fe(s)); }
    long long c0 = clock64(); unsigned long long t0 = gtime();
#pragma unroll 1
    for (int i = 0; i < iters_d; i++) {
#pragma unroll
      for (int k = 0; k < 2; k++) x[k] = fd_mul(x[k], y[k]);
    }
    long long c1 = clock64(); unsigned long long t1 = gtime();
#pragma unroll
    for (int k = 0; k < 2; k++) {
      Fe<FqTag> r = fd_to_std(x[k]);
#pragma unroll
      for (int i = 0; i < 8; i++) acc ^= r.v[i];
    }
    if (threadIdx.x == 96) clk[2 * blockIdx.x + 1] = Clk{c0, c1, t0, t1};
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NI>
static void run_mix(int sms, int bps) {
  const int blocks = sms * bps;
  uint32_t* out; Clk* clk;
  CK(cudaMalloc(&out, (size_t)blocks * 128 * 4)); CK(cudaMalloc(&clk, 2 * blocks * sizeof(Clk)));
  std::vector<Clk> h(2 * blocks);
  auto kind_ns = [&](int kind) {
    double ns = 0;
    for (int b = 0; b < blocks; b++) ns += (double)(h[2 * b + kind].t1 - h[2 * b + kind].t0);
    return ns / blocks;
  };
  // calibrate each kind alone at this geometry (the other kind's warps idle), then run both for ~150 ms
  int it_i = 20000, it_d = 20000;
  k_mix<NI><<<blocks, 128>>>(out, it_i, 0, clk); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h.data(), clk, 2 * blocks * sizeof(Clk), cudaMemcpyDeviceToHost));
  const double ns_i = NI > 0 ? kind_ns(0) / it_i : 1.0;
  k_mix<NI><<<blocks, 128>>>(out, 0, it_d, clk); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h.data(), clk, 2 * blocks * sizeof(Clk), cudaMemcpyDeviceToHost));
  const double ns_d = NI < 4 ? kind_ns(1) / it_d : 1.0;
  // co-running slows both: start from the solo rates, then rebalance once from the measured co-running rates
  it_i = NI > 0 ? (int)(1.5e8 / ns_i) : 0;
  it_d = NI < 4 ? (int)(1.5e8 / ns_d) : 0;
  double rate_i = 0, rate_d = 0, ms_evt = 0;
  int it_i_last = 0, it_d_last = 0;
  for (int round = 0; round < 4; round++) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_mix<NI><<<blocks, 128>>>(out, it_i, it_d, clk);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms_evt = ms;
    it_i_last = it_i; it_d_last = it_d;
    CK(cudaMemcpy(h.data(), clk, 2 * blocks * sizeof(Clk), cudaMemcpyDeviceToHost));
    const double t_i = NI > 0 ? kind_ns(0) : 0, t_d = NI < 4 ? kind_ns(1) : 0;
    rate_i = NI > 0 ? 2.0 * it_i * blocks * 32.0 * NI / (t_i * 1e-9) : 0;           // modmul/s while the IMAD warps ran
    rate_d = NI < 4 ? 2.0 * it_d * blocks * 32.0 * (4 - NI) / (t_d * 1e-9) : 0;
    if (NI > 0 && NI < 4) {                    // make both kinds take the same time
      const double target = 0.5 * (t_i + t_d);
      it_i = (int)(it_i * target / t_i);
      it_d = (int)(it_d * target / t_d);
    }
  }
  // whole-kernel accounting (CUDA events): every product of either kind over the kernel's duration -- a lower bound of
  // the combined rate (the per-warp clocks above only balance the two iteration counts; warps do not finish together)
  const double tot_i = 2.0 * it_i_last * blocks * 32.0 * NI, tot_d = 2.0 * it_d_last * blocks * 32.0 * (4 - NI);
  printf("{\"test\": \"mixed warps: %d IMAD + %d FP64 of 4 per block\", \"warps_per_sm\": %d, \"imad_modmul_per_s\": %.4e, "
         "\"fp52_modmul_per_s\": %.4e, \"sum_per_s\": %.4e, \"ms_events\": %.1f, \"per_warp_clock_rates\": [%.3e, %.3e]}\n",
         NI, 4 - NI, bps * 4, tot_i / (ms_evt * 1e-3), tot_d / (ms_evt * 1e-3), (tot_i + tot_d) / (ms_evt * 1e-3), ms_evt, rate_i, rate_d);
  fflush(stdout);
  CK(cudaFree(out)); CK(cudaFree(clk));
}

// correctness on the device: fp52 path == IMAD path, bit for bit, over random operands and a dependent chain
__global__ void k_check(unsigned long long* mismatches, int rounds) {
  uint64_t s = 0xabcdefull + (blockIdx.x * blockDim.x + threadIdx.x) * 7919ull;
  unsigned long long bad = 0;
  for (int r = 0; r < rounds; r++) {
    Fe<FqTag> a = rnd_fe(s), b = rnd_fe(s), c = rnd_fe(s), d = rnd_fe(s);
    Fd<FqTag> fa = fd_from_std(a), fb = fd_from_std(b), fc = fd_from_std(c), fdd = fd_from_std(d);
    // ((a*b)^2 - c) * d + a*c
    Fe<FqTag> e1 = fe_mul(a, b);
    e1 = fe_sqr(e1);
    e1 = fe_sub(e1, c);
    e1 = fe_mul_add_mul(e1, d, a, c);
    Fd<FqTag> f1 = fd_mul(fa, fb);
    f1 = fd_sqr(f1);
    f1 = fd_norm(fd_sub(f1, fc));
    f1 = fd_mul_add_mul(f1, fdd, fa, fc);
    if (!fe_eq(e1, fd_to_std(f1))) bad++;
  }
  if (bad) atomicAdd(mismatches, bad);
}

// ---- raw issue-rate plateaus -----------------------------------------------------------------------------------
constexpr int UNROLL = 8;
__global__ void __launch_bounds__(128) k_wide_pair(uint32_t* out, int iters, Clk* clk) {
  uint32_t lo[UNROLL], hi[UNROLL];
  uint32_t b = 0x9e3779b9u ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) { lo[j] = j + threadIdx.x + 12345u; hi[j] = j; }
  long long c0 = clock64(); unsigned long long t0 = gtime();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++)
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                   : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 4) % UNROLL]), "r"(b));
  }
  long long c1 = clock64(); unsigned long long t1 = gtime();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s ^= lo[j] ^ hi[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = Clk{c0, c1, t0, t1};
}
__global__ void __launch_bounds__(128) k_dfma_raw(uint32_t* out, int iters, Clk* clk) {
  double acc[UNROLL];
  double a = 1.000001 + threadIdx.x * 1e-9, b = 0.999999;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) acc[j] = j + threadIdx.x;
  long long c0 = clock64(); unsigned long long t0 = gtime();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[j]) : "d"(a), "d"(b));
  }
  long long c1 = clock64(); unsigned long long t1 = gtime();
  double s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double_as_longlong(s);
  if (threadIdx.x == 0) clk[blockIdx.x] = Clk{c0, c1, t0, t1};
}
// DFMA with integer adds beside it (2 IADD3-pairs per DFMA: what fp52 issues)
__global__ void __launch_bounds__(128) k_dfma_iadd(uint32_t* out, int iters, Clk* clk) {
  double acc[UNROLL];
  uint64_t col[UNROLL];
  double a = 1.000001 + threadIdx.x * 1e-9, b = 0.999999;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) { acc[j] = j + threadIdx.x; col[j] = j; }
  long long c0 = clock64(); unsigned long long t0 = gtime();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++) {
      asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[j]) : "d"(a), "d"(b));
      col[j] += (uint64_t)__double_as_longlong(acc[(j + 3) % UNROLL]);
    }
  }
  long long c1 = clock64(); unsigned long long t1 = gtime();
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s += col[j] + (uint64_t)__double_as_longlong(acc[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s;
  if (threadIdx.x == 0) clk[blockIdx.x] = Clk{c0, c1, t0, t1};
}

template <typename L>
static void run(const char* name, const char* unit, double units_per_thread_iter, int sms, int bps, int iters0, L launch) {
  const int threads = 128, blocks = sms * bps;
  uint32_t* out; Clk* clk;
  CK(cudaMalloc(&out, (size_t)blocks * threads * 4)); CK(cudaMalloc(&clk, blocks * sizeof(Clk)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // calibrate to >= 120 ms
  int iters = iters0;
  float ms = 0;
  for (int attempt = 0; attempt < 6; attempt++) {
    launch(blocks, threads, out, iters, clk);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    launch(blocks, threads, out, iters, clk);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms >= 120.f) break;
    double f = 150.0 / (ms > 0.01f ? ms : 0.01f);
    iters = (int)(iters * (f > 20 ? 20 : f)) + 1;
  }
  std::vector<Clk> h(blocks);
  CK(cudaMemcpy(h.data(), clk, blocks * sizeof(Clk), cudaMemcpyDeviceToHost));
  double cyc = 0, ns = 0;
  for (auto& c : h) { cyc += (double)(c.c1 - c.c0); ns += (double)(c.t1 - c.t0); }
  cyc /= blocks; ns /= blocks;
  double total = units_per_thread_iter * (double)iters * blocks * threads;
  printf("{\"test\": \"%s\", \"unit\": \"%s\", \"warps_per_sm\": %d, \"iters\": %d, \"ms_events\": %.3f, \"per_s\": %.4e, "
         "\"per_clk_per_sm_clock64\": %.3f, \"sm_mhz_clock64_over_globaltimer\": %.0f, \"per_clk_per_sm_at_1965\": %.3f}\n",
         name, unit, bps * 4, iters, ms, total / (ms * 1e-3), total / sms / cyc, cyc / ns * 1e3, total / (ms * 1e-3) / sms / 1.965e9);
  fflush(stdout);
  CK(cudaFree(out)); CK(cudaFree(clk));
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  if (argc > 1 && !strcmp(argv[1], "mix")) {      // do the IMAD and FP64 products add up when they share an SM?
    for (int bps : {2, 4}) {
      run_mix<4>(sms, bps);
      run_mix<3>(sms, bps);
      run_mix<2>(sms, bps);
      run_mix<1>(sms, bps);
      run_mix<0>(sms, bps);
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "ncu")) {      // one short launch of each product kernel, for ncu --set full
    uint32_t* out; Clk* clk;
    CK(cudaMalloc(&out, (size_t)sms * 4 * 128 * 4)); CK(cudaMalloc(&clk, sms * 4 * sizeof(Clk)));
    k_imad<2, 0><<<sms * 4, 128>>>(out, 3000, clk);
    k_dfma<2, 0><<<sms * 4, 128>>>(out, 3000, clk);
    k_dfma<1, 0><<<sms * 4, 128>>>(out, 6000, clk);
    CK(cudaDeviceSynchronize());
    return 0;
  }
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", p.name, sms, p.major, p.minor, p.clockRate);
  {
    unsigned long long* bad; CK(cudaMalloc(&bad, 8)); CK(cudaMemset(bad, 0, 8));
    k_check<<<sms * 2, 128>>>(bad, 64);
    unsigned long long hb = 1; CK(cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost));
    printf("{\"test\": \"fp52 == imad on device\", \"cases\": %d, \"mismatches\": %llu}\n", sms * 2 * 128 * 64, hb);
    CK(cudaFree(bad));
  }
  for (int bps : {2, 4, 8}) {
    run("IMAD.WIDE pair (mad.lo.cc+madc.hi), 8 independent", "wide-MAC", UNROLL, sms, bps, 20000,
        [](int b, int t, uint32_t* o, int it, Clk* c) { k_wide_pair<<<b, t>>>(o, it, c); });
    run("DFMA.RZ, 8 independent", "dfma", UNROLL, sms, bps, 20000,
        [](int b, int t, uint32_t* o, int it, Clk* c) { k_dfma_raw<<<b, t>>>(o, it, c); });
    run("DFMA.RZ + 64-bit integer add each", "dfma", UNROLL, sms, bps, 20000,
        [](int b, int t, uint32_t* o, int it, Clk* c) { k_dfma_iadd<<<b, t>>>(o, it, c); });
  }
  for (int bps : {2, 3, 4, 6, 8}) {
    run("imad mul, 2 chains", "modmul", 2, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_imad<2, 0><<<b, t>>>(o, it, c); });
    run("fp52 mul, 2 chains", "modmul", 2, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_dfma<2, 0><<<b, t>>>(o, it, c); });
  }
  for (int bps : {2, 4}) {
    run("imad mul, 1 chain", "modmul", 1, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_imad<1, 0><<<b, t>>>(o, it, c); });
    run("fp52 mul, 1 chain", "modmul", 1, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_dfma<1, 0><<<b, t>>>(o, it, c); });
    run("imad mul, 4 chains", "modmul", 4, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_imad<4, 0><<<b, t>>>(o, it, c); });
    run("fp52 mul, 4 chains", "modmul", 4, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_dfma<4, 0><<<b, t>>>(o, it, c); });
    run("imad sqr, 2 chains", "modsqr", 2, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_imad<2, 1><<<b, t>>>(o, it, c); });
    run("fp52 sqr, 2 chains", "modsqr", 2, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_dfma<2, 1><<<b, t>>>(o, it, c); });
    run("imad a*b+c*d, 2 chains", "dual", 2, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_imad<2, 2><<<b, t>>>(o, it, c); });
    run("fp52 a*b+c*d, 2 chains", "dual", 2, sms, bps, 2000, [](int b, int t, uint32_t* o, int it, Clk* c) { k_dfma<2, 2><<<b, t>>>(o, it, c); });
  }
  return 0;
}
