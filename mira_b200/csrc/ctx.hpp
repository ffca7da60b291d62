// ctx.hpp — host-side context shared by the per-curve translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mira_b200.h"
#include "stager.hpp"

namespace mira_host {

int fail(int code, const char* fmt, ...);   // records the thread-local error text, returns `code`

#define CU(x)                                                                                          \
  do {                                                                                                 \
    cudaError_t e_ = (x);                                                                              \
    if (e_ != cudaSuccess)                                                                             \
      return ::mira_host::fail(MIRA_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return MIRA_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cap = bytes;
    return MIRA_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// Fixed-base table for one window width: table[j * n_cover + i] = 2^(c*j) * P_i (affine, 64 B)
struct Table {
  int c = 0, W = 0;
  uint32_t n_cover = 0;
  void* d = nullptr;
  size_t bytes = 0;
  uint64_t last_use = 0;      // ctx->table_clock at the last commit that used it (LRU eviction under memory pressure)
};

inline int windows_for(int c) { return (255 + c - 1) / c; }
int choose_window(size_t n);
int choose_window_sampled(size_t n, const uint32_t* bitlen_hist, size_t samples, double* costs /* [25], may be null */);

}  // namespace mira_host

struct mira_msm_ctx {
  int curve = 0;
  int device = 0;
  size_t n_bases = 0;
  void* d_bases = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;          // H2D of host-buffer commits, overlapped with compute slice by slice
  cudaEvent_t copy_done[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t compute_idle = nullptr;
  mira_host::Stager stager;                    // pageable host scalars go through page-locked slots (stager.hpp)
  std::vector<mira_host::Table> tables;
  uint64_t table_clock = 0;
  // workspace (grown on demand, reused across commits)
  mira_host::DevBuf scalars, cursor, buckets, part_keys, part_pts, red_a, red_b, red_c, red_d, result;
  // Pair-list workspace of one slice: unsorted and sorted (key, ref) arrays, counters, radix-sort scratch.  Two sets, so
  // that slice k+1 can be decomposed and sorted (prep_stream) while slice k is accumulated (pipeline.cuh).
  struct SortBufs {
    mira_host::DevBuf keys, refs, skeys, srefs, counts, tile_sums;
  } sb[2];
  cudaStream_t prep_stream = nullptr;          // digits + sort of the next slice, overlapped with the accumulation
  cudaEvent_t prep_done[2] = {nullptr, nullptr}, acc_done[2] = {nullptr, nullptr}, pipe_start = nullptr;
  int pipe_slices = -1;                        // device-resident commits: -1 = MIRA_PIPE_SLICES from the environment (default 1 = not sliced)
  size_t pipe_min_slice = (size_t)1 << 20;     // slices below ~1 M scalars cost more in launches than they hide
  mira_host::DevBuf pa_a, pa_b, pa_work;       // batched-affine levels (affine_levels.cuh): two ping-pong (key, point) lists, scratch
  int affine_levels = -1;                      // -1 = MIRA_AFFINE_LEVELS from the environment (default 0 = off)
  size_t scalars_valid = 0;                    // scalars.p holds the device copy of the last host-buffer commit (this many)
  void* h_result = nullptr;  // pinned, 4 KiB (up to 32 affine results of a batched commit)
  int forced_window = 0;
  bool adaptive_window = true;                 // pick the window from a sample of the scalars (sparse witnesses want a narrow one)
  mira_host::DevBuf sample;                    // sampled scalars + bit-length histogram
  uint32_t* h_hist = nullptr;                  // pinned, 2048 u32 (257 bit-length bins; word 258: pair count read back for sparse vectors;
                                               // 260..1283: hash counts of the sampled scalars)
  bool sampled_heavy = false;                  // the sample of the CURRENT commit holds a value that makes up > ~1.5 % of the vector
  double sampled_pairs_per_scalar = 0.0;       // density the sample of the CURRENT commit predicts (0 = not sampled)
  size_t slice_min = (size_t)1 << 19;          // host-buffer commits: smallest (first) slice of the geometric H2D pipeline
  bool profiling = false;
  mira_msm_stats stats{};
  std::mutex mu;
  // Single-process multi-GPU key (mira_msm_ctx_create_sharded): the parent holds no bases itself; shard g is an ordinary
  // context over key indices [shard_lo[g], shard_lo[g + 1]) on its own device, and `gather` (on the parent's device)
  // receives the shards' 128-byte XYZZ partial sums by peer copy.
  std::vector<mira_msm_ctx*> shards;
  std::vector<size_t> shard_lo;
  mira_host::DevBuf gather;
};

namespace mira_host {

// Per-curve entry points (curve_bn254.cu / curve_grumpkin.cu instantiate the templates in pipeline.cuh)
struct CurveOps {
  int (*commit)(mira_msm_ctx*, const void* scalars, size_t n, int on_device, void* out, bool want_affine, cudaStream_t st);
  int (*commit_batch)(mira_msm_ctx*, const void* const* scalar_sets_dev, size_t count, size_t n, void* out, cudaStream_t st);
  int (*prepare)(mira_msm_ctx*, size_t n, const void* like_scalars, int on_device);
  int (*check_on_curve)(mira_msm_ctx*);
  int (*combine)(const void* partials_host, size_t count, void* out_affine_host);
  int (*partial_batch_dev)(mira_msm_ctx*, const void* const* scalar_sets_dev, size_t count, size_t n, void* out_xyzz_dev, cudaStream_t st);
  int (*combine_dev)(const void* partials_dev, size_t n_ranks, size_t n_commits, size_t rank_stride, void* out_affine_host, cudaStream_t st);
  int (*partial_to_peer)(mira_msm_ctx*, const void* scalars_host, size_t n, void* dst_dev, int dst_device, cudaStream_t st);
  int (*gen_scalars)(uint64_t seed, size_t first, size_t n, int dist, void* out_dev);
  int (*gen_bases)(uint64_t seed, size_t first, size_t n, void* out_dev);
  int (*test_point_op)(int op, const void* p_dev, const void* q_dev, size_t n, void* out_dev);
};
extern const CurveOps OPS_BN254, OPS_GRUMPKIN;
inline const CurveOps& ops_for(int curve) { return curve == MIRA_BN254_G1 ? OPS_BN254 : OPS_GRUMPKIN; }

// device radix sort of (key, ref) pairs (sort.cu)
size_t radix_sort_temp_bytes(size_t max_pairs);
int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, const uint32_t* n_ptr,
                     size_t max_pairs, int key_bits, void* temp, cudaStream_t st, int* sorted_in_b, uint64_t* launches,
                     int first_pass = 0);

// field test hook (field_test.cu)
int test_field_op_dev(int field, int op, const void* a_dev, const void* b_dev, size_t n, void* out_dev);

}  // namespace mira_host
