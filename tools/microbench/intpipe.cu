// Integer / FP64 pipe issue-rate microbenchmark for sm_100a (B200).
// Measures warp-instruction throughput of the instructions a 8x32-bit-limb
// Montgomery multiplier is made of, so that the MSM roofline denominator
// ("measured mad.wide.u32 issue rate") is a measured number.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o intpipe intpipe.cu
// Output: one JSON line per test: lane-ops per clock per SM (from clock64 deltas)
// and lane-ops/s (from CUDA events).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int UNROLL = 8;   // independent chains per thread

// ---- 1. mad.wide.u32, independent 64-bit accumulators -----------------------
__global__ void k_mad_wide(uint64_t* out, uint32_t a0, uint32_t b0, long long* cyc) {
  uint64_t acc[UNROLL];
  uint32_t a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) acc[j] = j + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++)
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[(j + 4) % UNROLL]), "r"(b));
  }
  long long t1 = clock64();
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s ^= acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 2. mad.lo.cc / madc.hi.cc carry chain (what ptxas fuses to IMAD.WIDE.U32[.X])
__global__ void k_mad_chain(uint32_t* out, uint32_t a0, uint32_t b0, long long* cyc) {
  uint32_t acc[2 * UNROLL];
  uint32_t a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < 2 * UNROLL; j++) acc[j] = j + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
    // one chain of UNROLL wide MACs with carry propagated pair to pair
    asm volatile(
        "mad.lo.cc.u32 %0, %16, %17, %0;\n\t"  "madc.hi.cc.u32 %1, %16, %17, %1;\n\t"
        "madc.lo.cc.u32 %2, %16, %17, %2;\n\t" "madc.hi.cc.u32 %3, %16, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %16, %17, %4;\n\t" "madc.hi.cc.u32 %5, %16, %17, %5;\n\t"
        "madc.lo.cc.u32 %6, %16, %17, %6;\n\t" "madc.hi.cc.u32 %7, %16, %17, %7;\n\t"
        "madc.lo.cc.u32 %8, %16, %17, %8;\n\t" "madc.hi.cc.u32 %9, %16, %17, %9;\n\t"
        "madc.lo.cc.u32 %10, %16, %17, %10;\n\t" "madc.hi.cc.u32 %11, %16, %17, %11;\n\t"
        "madc.lo.cc.u32 %12, %16, %17, %12;\n\t" "madc.hi.cc.u32 %13, %16, %17, %13;\n\t"
        "madc.lo.cc.u32 %14, %16, %17, %14;\n\t" "madc.hi.u32 %15, %16, %17, %15;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]),
          "+r"(acc[12]), "+r"(acc[13]), "+r"(acc[14]), "+r"(acc[15])
        : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 2 * UNROLL; j++) s ^= acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 3/4. mad.lo.u32 and mad.hi.u32 (32-bit results) --------------------------
template <int HI>
__global__ void k_mad32(uint32_t* out, uint32_t a0, uint32_t b0, long long* cyc) {
  uint32_t acc[UNROLL];
  uint32_t a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) acc[j] = j + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++) {
      if (HI) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(acc[(j + 4) % UNROLL]), "r"(b));
      else    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(acc[(j + 4) % UNROLL]), "r"(b));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s ^= acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 5. IADD3 with carry (add.cc / addc.cc chain, 8 limbs) ---------------------
__global__ void k_addc(uint32_t* out, uint32_t a0, long long* cyc) {
  uint32_t acc[UNROLL], b[UNROLL];
#pragma unroll
  for (int j = 0; j < UNROLL; j++) { acc[j] = j + threadIdx.x; b[j] = a0 + j * threadIdx.x; }
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
    asm volatile(
        "add.cc.u32 %0, %0, %8;\n\t"  "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t" "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t" "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t" "addc.u32 %7, %7, %15;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s ^= acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 6. DFMA ----------------------------------------------------------------------
__global__ void k_dfma(double* out, double a0, double b0, long long* cyc) {
  double acc[UNROLL];
  double a = a0 + threadIdx.x * 1e-9, b = b0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) acc[j] = j + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++)
      asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(acc[j]) : "d"(acc[(j + 4) % UNROLL]), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 7. mixed: mad.wide + DFMA in the same warp (do the two pipes overlap?) ----------
__global__ void k_mix_wide_dfma(uint64_t* out, uint32_t a0, uint32_t b0, double da, double db, long long* cyc) {
  uint64_t acc[UNROLL];
  double dacc[UNROLL];
  uint32_t a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) { acc[j] = j + threadIdx.x; dacc[j] = j; }
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[(j + 4) % UNROLL]), "r"(b));
      asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(dacc[j]) : "d"(dacc[(j + 4) % UNROLL]), "d"(db));
    }
  }
  long long t1 = clock64();
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s ^= acc[j] ^ (uint64_t)__double_as_longlong(dacc[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 8. mixed: mad.wide + 2x IADD3 (plain-C style MAC: product, add lo, addc hi) -------
__global__ void k_mix_wide_iadd(uint64_t* out, uint32_t a0, uint32_t b0, long long* cyc) {
  uint64_t acc[UNROLL];
  uint32_t lo[UNROLL], hi[UNROLL];
  uint32_t a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) { acc[j] = j + threadIdx.x; lo[j] = j; hi[j] = j * 3; }
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[(j + 4) % UNROLL]), "r"(b));
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 4) % UNROLL]), "r"(b));
    }
  }
  long long t1 = clock64();
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s ^= acc[j] ^ lo[j] ^ ((uint64_t)hi[j] << 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 9. mul.wide.u32 only (no accumulate): IMAD.WIDE.U32 Rd, Ra, Rb, RZ ---------------
__global__ void k_mul_wide(uint64_t* out, uint32_t a0, uint32_t b0, long long* cyc) {
  uint64_t acc[UNROLL];
  uint32_t b = b0 ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) acc[j] = j + threadIdx.x + a0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++)
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[j]) : "r"((uint32_t)acc[(j + 4) % UNROLL] + 1u), "r"(b));
  }
  long long t1 = clock64();
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s ^= acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 10. mad.lo.cc + madc.hi pairs, independent (IMAD.WIDE.U32 with 64-bit accumulate, no .X) ---
__global__ void k_mad_pair(uint32_t* out, uint32_t a0, uint32_t b0, long long* cyc) {
  uint32_t lo[UNROLL], hi[UNROLL];
  uint32_t b = b0 ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) { lo[j] = j + threadIdx.x + a0; hi[j] = j; }
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int j = 0; j < UNROLL; j++)
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                   : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 4) % UNROLL]), "r"(b));
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < UNROLL; j++) s ^= lo[j] ^ hi[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, double lane_ops_per_thread, int blocks, int threads, long long* d_cyc, F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) launch();
  CK(cudaDeviceSynchronize());
  const int reps = 5;
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  long long* h = (long long*)malloc(sizeof(long long) * blocks);
  CK(cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
  double mean = 0; for (int i = 0; i < blocks; i++) mean += (double)h[i]; mean /= blocks;
  free(h);
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double total = lane_ops_per_thread * (double)blocks * threads;
  double per_sm = total / sms;
  // all blocks are co-resident (blocks = sms * k with k small), so every SM is busy for ~mean cycles
  printf("{\"test\": \"%s\", \"blocks\": %d, \"threads\": %d, \"ms\": %.4f, \"lane_ops_per_s\": %.4e, "
         "\"lane_ops_per_clk_per_sm\": %.2f, \"mean_cycles\": %.0f, \"implied_mhz\": %.0f}\n",
         name, blocks, threads, ms, total / (ms * 1e-3), per_sm / mean, mean, mean / (ms * 1e-3) / 1e6);
  fflush(stdout);
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", p.name, sms, p.major, p.minor, p.clockRate);
  for (int wps = 8; wps <= 32; wps *= 2) {   // warps per SM: 4, 8, 16, 32
    int threads = 128, bps = wps / 4, blocks = sms * bps;
    void* out; long long* cyc;
    CK(cudaMalloc(&out, (size_t)blocks * threads * 8)); CK(cudaMalloc(&cyc, blocks * sizeof(long long)));
    char nm[96];
    double n = (double)ITERS * UNROLL;
    snprintf(nm, 96, "mad.wide.u32 (ptxas: mul.wide+IADD3+IMAD.X) w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_mad_wide<<<blocks, threads>>>((uint64_t*)out, 12345u, 0x9e3779b9u, cyc); });
    snprintf(nm, 96, "mad.lo.cc/madc.hi.cc chain (wide MACs) w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_mad_chain<<<blocks, threads>>>((uint32_t*)out, 12345u, 0x9e3779b9u, cyc); });
    snprintf(nm, 96, "mul.wide.u32 (no acc) w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_mul_wide<<<blocks, threads>>>((uint64_t*)out, 12345u, 0x9e3779b9u, cyc); });
    snprintf(nm, 96, "mad.lo.cc+madc.hi pair (wide acc) w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_mad_pair<<<blocks, threads>>>((uint32_t*)out, 12345u, 0x9e3779b9u, cyc); });
    snprintf(nm, 96, "mad.lo.u32 w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_mad32<0><<<blocks, threads>>>((uint32_t*)out, 12345u, 0x9e3779b9u, cyc); });
    snprintf(nm, 96, "mad.hi.u32 w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_mad32<1><<<blocks, threads>>>((uint32_t*)out, 12345u, 0x9e3779b9u, cyc); });
    snprintf(nm, 96, "addc chain (IADD3.X) w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_addc<<<blocks, threads>>>((uint32_t*)out, 12345u, cyc); });
    snprintf(nm, 96, "fma.f64 w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_dfma<<<blocks, threads>>>((double*)out, 1.000001, 0.999999, cyc); });
    snprintf(nm, 96, "mix mad.wide+dfma (count=wide only) w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_mix_wide_dfma<<<blocks, threads>>>((uint64_t*)out, 12345u, 0x9e3779b9u, 1.000001, 0.999999, cyc); });
    snprintf(nm, 96, "mix mad.wide+2 iadd (count=wide only) w%d", wps);
    run(nm, n, blocks, threads, cyc, [&] { k_mix_wide_iadd<<<blocks, threads>>>((uint64_t*)out, 12345u, 0x9e3779b9u, cyc); });
    CK(cudaFree(out)); CK(cudaFree(cyc));
  }
  return 0;
}
