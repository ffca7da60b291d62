// Reference point for the hand-written radix sort (mira_b200/csrc/sort.cu): CUB's DeviceRadixSort::SortPairs (onesweep)
// on the same input -- 2^24 x 12 = 201,326,592 (22-bit key, 32-bit value) pairs -- and a plain device copy of the same
// bytes as the bandwidth floor.  Library code is used HERE ONLY, as a yardstick; the shipped sort is sort.cu.
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -o sort_ref sort_ref.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cub/cub.cuh>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_fill(uint32_t* k, uint32_t* v, size_t n, int bits) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t s = i * 0x9E3779B97F4A7C15ull + 0x1234567;
  s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
  k[i] = (uint32_t)s & ((1u << bits) - 1u);
  v[i] = (uint32_t)i;
}

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : (size_t)201326592;
  const int bits = argc > 2 ? atoi(argv[2]) : 22;
  uint32_t *ka, *va, *kb, *vb;
  CK(cudaMalloc(&ka, n * 4)); CK(cudaMalloc(&va, n * 4)); CK(cudaMalloc(&kb, n * 4)); CK(cudaMalloc(&vb, n * 4));
  k_fill<<<(unsigned)((n + 255) / 256), 256>>>(ka, va, n, bits);
  void* tmp = nullptr; size_t tmp_bytes = 0;
  cub::DoubleBuffer<uint32_t> dk(ka, kb), dv(va, vb);
  CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)n, 0, bits));
  CK(cudaMalloc(&tmp, tmp_bytes));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  for (int rep = 0; rep < 4; rep++) {
    k_fill<<<(unsigned)((n + 255) / 256), 256>>>(ka, va, n, bits);
    cub::DoubleBuffer<uint32_t> k2(ka, kb), v2(va, vb);
    CK(cudaEventRecord(e0));
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k2, v2, (int)n, 0, bits));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"what\": \"cub::DeviceRadixSort::SortPairs\", \"pairs\": %zu, \"key_bits\": %d, \"ms\": %.3f, \"temp_bytes\": %zu}\n", n, bits, ms, tmp_bytes);
  }
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    CK(cudaMemcpyAsync(kb, ka, n * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpyAsync(vb, va, n * 4, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"what\": \"device copy of the same pairs (one read + one write of 8 B per pair)\", \"pairs\": %zu, \"ms\": %.3f, \"gb_per_s\": %.0f}\n", n, ms, n * 16.0 / ms / 1e6);
  }
  return 0;
}
