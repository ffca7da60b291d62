// sort.cu — device LSD radix sort of (bucket, point-ref) pairs, staged through shared memory.
//
// There is no analogue in the reference (its CPU Pippenger walks buckets serially inside
// halo2's multiexp_serial, called at /root/reference/src/commitment.rs:80); on the GPU the n*W
// (bucket, point) pairs must be grouped by bucket before the bucket sums can be formed in parallel.
//
// 8 bits per pass, ceil(key_bits / 8) passes, each pass = histogram -> exclusive scan -> stable scatter:
//   * k_radix_hist    per-tile digit histogram (shared-memory atomics), written digit-major so ONE
//                     exclusive scan over [256][n_tiles] yields the global base of every (digit, tile);
//   * k_radix_scatter warp-private stable ranking with __match_any_sync, tile re-ordered in shared
//                     memory, then written out so each digit's run is one coalesced burst.
// The number of valid pairs lives in device memory (*n_ptr, produced by the digit kernel's compaction);
// grids are sized for the worst case and surplus tiles exit, so no host synchronisation is needed.
// HBM traffic per pass: 4 B (hist) + 8 B read + 8 B write per pair.
#include "ctx.hpp"

#include <cstdlib>

namespace mira_host {

constexpr int RS_BINS = 256;
#ifndef MIRA_RS_ITEMS
#define MIRA_RS_ITEMS 8
#endif
#ifndef MIRA_RS_THREADS
#define MIRA_RS_THREADS 512
#endif
constexpr int RS_ITEMS = MIRA_RS_ITEMS;            // pairs per thread

// Tile geometry: THREADS x RS_ITEMS pairs per tile.  Measured at 2^24 points (3 passes, ms): 256x16 5.61, 256x8 5.51,
// 384x8 5.17, 512x8 4.91 (shipped), 512x4 6.02, 768x4 6.05, 1024x4 5.67, 1024x8 6.09.  ncu (profiles/r01_msm_aux_v7.txt): the scatter is
// latency-bound (46 % of stall samples wait for the tile's own key loads), not bandwidth-bound (24 % of HBM peak).
// A persistent variant that prefetched the NEXT tile (keys in registers, values via a cp.async double buffer) was
// measured SLOWER (5.7 vs 4.9 ms): nothing is saturated (L1/smem 51 %, L2 19 %, DRAM 24 %, issue 41 %), the five
// block-wide barriers per tile are what serialises it; a one-sweep design with decoupled look-back is the next step.
template <int THREADS> struct RsCfg {
  static constexpr int WARPS = THREADS / 32;
  static constexpr int TILE = THREADS * RS_ITEMS;
  static constexpr int WARP_TILE = 32 * RS_ITEMS;  // consecutive pairs owned by one warp
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_radix_hist(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr,
                                                        int shift, uint32_t* __restrict__ hist, uint32_t n_tiles) {
  constexpr int TILE = RsCfg<THREADS>::TILE;
  __shared__ uint32_t sh[RS_BINS];
  const uint32_t n = *n_ptr;
  const uint32_t base = blockIdx.x * TILE;
  for (int i = threadIdx.x; i < RS_BINS; i += THREADS) sh[i] = 0;
  __syncthreads();
  if (base < n) {
    uint32_t kk[RS_ITEMS];                       // all loads in flight before the first shared-memory atomic
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
      uint32_t idx = base + k * THREADS + threadIdx.x;
      kk[k] = idx < n ? __ldg(keys + idx) : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
      uint32_t idx = base + k * THREADS + threadIdx.x;
      if (idx < n) atomicAdd(&sh[(kk[k] >> shift) & 0xffu], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RS_BINS; i += THREADS) hist[(size_t)i * n_tiles + blockIdx.x] = sh[i];
}

// ---- exclusive scan over the [256 * n_tiles] histogram (3 phases, same scheme as the bucket offsets)
constexpr int SC_THREADS = 512;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* sm, uint32_t& total) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sm[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    sm[lane] = w;
  }
  __syncthreads();
  uint32_t off = warp ? sm[warp - 1] : 0u;
  total = sm[(blockDim.x >> 5) - 1];
  __syncthreads();
  return off + x - v;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_sums(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ sums) {
  __shared__ uint32_t sm[32];
  uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) s += (base + k < n) ? in[base + k] : 0u;
  uint32_t total;
  block_excl_scan(s, sm, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SC_THREADS) k_scan_top(uint32_t* __restrict__ data, uint32_t n) {
  __shared__ uint32_t sm[32];
  uint32_t running = 0;
  for (uint32_t base = 0; base < n; base += SC_THREADS) {
    uint32_t idx = base + threadIdx.x;
    uint32_t v = (idx < n) ? data[idx] : 0u;
    uint32_t total;
    uint32_t ex = block_excl_scan(v, sm, total);
    if (idx < n) data[idx] = running + ex;
    running += total;
  }
}
__global__ void __launch_bounds__(SC_THREADS) k_scan_down(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ offs) {
  __shared__ uint32_t sm[32];
  uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
  uint32_t v[SC_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    v[k] = (base + k < n) ? data[base + k] : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t ex = block_excl_scan(s, sm, total) + offs[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    if (base + k < n) data[base + k] = ex;
    ex += v[k];
  }
}

// ---- stable scatter of one tile
template <int THREADS>
struct RsSmem {
  uint32_t cnt[RsCfg<THREADS>::WARPS][RS_BINS + 1];   // per-warp digit counts -> per-warp exclusive offsets (bin 256 = padding)
  uint32_t dstart[RS_BINS + 1];                       // first tile-local slot of each digit
  uint32_t gbase[RS_BINS];                            // global base of (digit, this tile)
  uint32_t scan_tmp[32];
  uint32_t keys[RsCfg<THREADS>::TILE];
  uint32_t vals[RsCfg<THREADS>::TILE];
  uint32_t vals_in[RsCfg<THREADS>::TILE];             // the tile's values, staged with cp.async while keys are ranked
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_radix_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                           const uint32_t* __restrict__ n_ptr, int shift,
                                                           const uint32_t* __restrict__ hist_scanned, uint32_t n_tiles,
                                                           uint32_t* __restrict__ out_keys, uint32_t* __restrict__ out_vals) {
  using Cfg = RsCfg<THREADS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RsSmem<THREADS>& S = *reinterpret_cast<RsSmem<THREADS>*>(smem_raw);
  const uint32_t n = *n_ptr;
  const uint32_t tile_base = blockIdx.x * Cfg::TILE;
  if (tile_base >= n) return;
  const uint32_t tile_count = (n - tile_base) < (uint32_t)Cfg::TILE ? (n - tile_base) : (uint32_t)Cfg::TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (int i = threadIdx.x; i < Cfg::WARPS * (RS_BINS + 1); i += THREADS) (&S.cnt[0][0])[i] = 0;
  for (int i = threadIdx.x; i < RS_BINS; i += THREADS) S.gbase[i] = hist_scanned[(size_t)i * n_tiles + blockIdx.x];
  __syncthreads();

  // phase 0: every load of the tile is issued up front (the kernel is latency-bound otherwise): keys into
  // registers, values through cp.async into shared memory where phase 3 picks them up.
  uint32_t k[RS_ITEMS];
  uint16_t rank[RS_ITEMS];
  uint32_t* my_cnt = S.cnt[warp];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    uint32_t local = warp * Cfg::WARP_TILE + r * 32 + lane;
    k[r] = local < tile_count ? __ldg(keys + tile_base + local) : 0u;
  }
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    uint32_t local = warp * Cfg::WARP_TILE + r * 32 + lane;
    if (local < tile_count) {
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(&S.vals_in[local]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(vals + tile_base + local) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // phase 1: warp-private stable ranks.  Warp w owns the WARP_TILE consecutive pairs starting at w*WARP_TILE.
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    uint32_t local = warp * Cfg::WARP_TILE + r * 32 + lane;
    bool valid = local < tile_count;
    uint32_t d = valid ? ((k[r] >> shift) & 0xffu) : (uint32_t)RS_BINS;
    // lanes holding the same digit: one ballot per digit bit (cost independent of how many distinct digits the
    // warp holds, unlike MATCH.ANY); invalid lanes (ragged last tile) form their own group on bin 256
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
    if (!valid) peers = ~peers;
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
      uint32_t b = __ballot_sync(0xffffffffu, (d >> bit) & 1u);
      peers &= ((d >> bit) & 1u) ? b : ~b;
    }
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader) {
      old = my_cnt[d];
      my_cnt[d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = (uint16_t)(old + __popc(peers & lt_mask));
    __syncwarp();
  }
  __syncthreads();

  // phase 2: per digit, exclusive prefix over warps; then exclusive scan over digits
  uint32_t total = 0;
  if (threadIdx.x < RS_BINS) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < Cfg::WARPS; w++) {
      uint32_t t = S.cnt[w][threadIdx.x];
      S.cnt[w][threadIdx.x] = run;
      run += t;
    }
    total = run;
  }
  uint32_t tile_total;
  uint32_t ds = block_excl_scan(total, S.scan_tmp, tile_total);
  if (threadIdx.x < RS_BINS) S.dstart[threadIdx.x] = ds;
  if (threadIdx.x == 0) S.dstart[RS_BINS] = tile_total;
  __syncthreads();

  // phase 3: place every pair at its tile-local sorted slot (each thread reads back the values it staged itself)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    uint32_t local = warp * Cfg::WARP_TILE + r * 32 + lane;
    if (local < tile_count) {
      uint32_t d = (k[r] >> shift) & 0xffu;
      uint32_t q = S.dstart[d] + my_cnt[d] + rank[r];
      S.keys[q] = k[r];
      S.vals[q] = S.vals_in[local];
    }
  }
  __syncthreads();

  // phase 4: stream the re-ordered tile out; a digit's run is contiguous in both smem and HBM
  for (uint32_t q = threadIdx.x; q < tile_count; q += THREADS) {
    uint32_t key = S.keys[q];
    uint32_t d = (key >> shift) & 0xffu;
    uint32_t dst = S.gbase[d] + (q - S.dstart[d]);
    out_keys[dst] = key;
    out_vals[dst] = S.vals[q];
  }
}

constexpr int RS_THREADS = MIRA_RS_THREADS;        // the shipped geometry
constexpr int RS_THREADS_SMALL = 256;              // alternative: smaller blocks that fit onto an SM beside accumulation blocks

// MIRA_RS_THREADS_RT=256 selects the small geometry at run time (development knob for the overlap experiments)
static int rs_threads_runtime() {
  static const int v = [] { const char* e = getenv("MIRA_RS_THREADS_RT"); return e && atoi(e) == RS_THREADS_SMALL ? RS_THREADS_SMALL : RS_THREADS; }();
  return v;
}

size_t radix_sort_temp_bytes(size_t max_pairs) {
  const size_t tile = RsCfg<RS_THREADS_SMALL>::TILE;       // the smaller tile needs the larger histogram
  size_t n_tiles = (max_pairs + tile - 1) / tile;
  size_t hist = (size_t)RS_BINS * n_tiles;
  size_t sums = (hist + SC_TILE - 1) / SC_TILE + 1;
  return (hist + sums + 64) * 4;
}

template <int THREADS>
static int radix_sort_pairs_t(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, const uint32_t* n_ptr,
                              size_t max_pairs, int key_bits, void* temp, cudaStream_t st, int* sorted_in_b, uint64_t* launches,
                              int first_pass) {
  constexpr int TILE = RsCfg<THREADS>::TILE;
  static std::once_flag attr_once[64];       // the attribute is per device; contexts of several keys may race here
  int dev = 0;
  CU(cudaGetDevice(&dev));
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once[dev & 63], [&] {
    attr_err = cudaFuncSetAttribute(k_radix_scatter<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem<THREADS>));
  });
  CU(attr_err);
  const uint32_t n_tiles = (uint32_t)((max_pairs + TILE - 1) / TILE);
  if (n_tiles == 0) {
    *sorted_in_b = 0;
    return MIRA_OK;
  }
  const uint32_t hist_n = RS_BINS * n_tiles;
  const uint32_t n_sums = (hist_n + SC_TILE - 1) / SC_TILE;
  uint32_t* hist = reinterpret_cast<uint32_t*>(temp);
  uint32_t* sums = hist + hist_n;
  int passes = (key_bits + 7) / 8;
  if (passes < 1) passes = 1;
  uint32_t *ka = keys_a, *va = vals_a, *kb = keys_b, *vb = vals_b;
  for (int p = first_pass; p < passes; p++) {
    int shift = 8 * p;
    k_radix_hist<THREADS><<<n_tiles, THREADS, 0, st>>>(ka, n_ptr, shift, hist, n_tiles);
    k_scan_sums<<<n_sums, SC_THREADS, 0, st>>>(hist, hist_n, sums);
    k_scan_top<<<1, SC_THREADS, 0, st>>>(sums, n_sums);
    k_scan_down<<<n_sums, SC_THREADS, 0, st>>>(hist, hist_n, sums);
    k_radix_scatter<THREADS><<<n_tiles, THREADS, sizeof(RsSmem<THREADS>), st>>>(ka, va, n_ptr, shift, hist, n_tiles, kb, vb);
    if (launches) *launches += 5;
    std::swap(ka, kb);
    std::swap(va, vb);
  }
  CU(cudaGetLastError());
  *sorted_in_b = passes > first_pass ? ((passes - first_pass) & 1) : 0;
  return MIRA_OK;
}

// Sorts the first *n_ptr pairs of (keys_a, vals_a) by the low `key_bits` bits of the key.  Ping-pongs
// between the a/b buffers; *sorted_in_b tells where the result ends up.  All launches go to `st`.
// first_pass = 1: the pairs are already grouped by the key's low byte (the digit kernel's fused first pass,
// msm_kernels.cuh: k_digits_scatter), so only the passes from bit 8 up are run.
int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, const uint32_t* n_ptr,
                     size_t max_pairs, int key_bits, void* temp, cudaStream_t st, int* sorted_in_b, uint64_t* launches, int first_pass) {
  if (rs_threads_runtime() == RS_THREADS_SMALL)
    return radix_sort_pairs_t<RS_THREADS_SMALL>(keys_a, vals_a, keys_b, vals_b, n_ptr, max_pairs, key_bits, temp, st, sorted_in_b, launches, first_pass);
  return radix_sort_pairs_t<RS_THREADS>(keys_a, vals_a, keys_b, vals_b, n_ptr, max_pairs, key_bits, temp, st, sorted_in_b, launches, first_pass);
}

}  // namespace mira_host
