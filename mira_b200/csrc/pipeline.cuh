// pipeline.cuh — host orchestration of the MSM pipeline, templated on <coordinate field, scalar field>.
// Instantiated once per curve (curve_bn254.cu, curve_grumpkin.cu) so the two curves compile in parallel.
#pragma once
#include <algorithm>
#include <cstdlib>
#include "ctx.hpp"
#include "msm_kernels.cuh"
#include "affine_levels.cuh"
#include "coop.cuh"
#include "testgen.cuh"

namespace mira_host {
using namespace mira;

// The table is built on the stream of the commit that needs it (`st`): every kernel that reads it is queued behind
// k_precompute on the same stream, so a lazily built table can never be read half-written (ctx->stream, where it used
// to be built, is non-blocking and unordered with respect to a caller's stream).
template <class CF>
int build_table(mira_msm_ctx* ctx, int c, uint32_t n_cover, Table* out, cudaStream_t st) {
  int W = windows_for(c);
  void* d = nullptr;
  size_t bytes = (size_t)W * n_cover * 64;
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(MIRA_ERR_CUDA, "cudaMalloc(%zu) for the fixed-base table failed: %s", bytes, cudaGetErrorString(e));
  }
  k_precompute<CF><<<(n_cover + 127) / 128, 128, 0, st>>>(ctx->d_bases, n_cover, c, W, d);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(MIRA_ERR_CUDA, "k_precompute launch failed: %s", cudaGetErrorString(e));
  }
  out->c = c;
  out->W = W;
  out->n_cover = n_cover;
  out->d = d;
  out->bytes = bytes;
  return MIRA_OK;
}

// Returns the table for window c covering at least n points, building it on first use.  The cache is bounded:
//   * a table is only built if it fits beside TABLE_RESERVE bytes of free device memory (workspace of the commit
//     itself); least-recently-used tables are evicted to make room (cudaFree synchronises the device, so no kernel
//     still reads an evicted table);
//   * if it cannot be built at all, the commit falls back to the cached table covering n whose window is closest to
//     the one asked for (`*out` then has a different c: callers read the window from the table, not from the request).
constexpr size_t TABLE_RESERVE = (size_t)6 << 30;
template <class CF>
int get_table(mira_msm_ctx* ctx, int c, size_t n, cudaStream_t st, Table** out) {
  ctx->table_clock++;
  for (auto& t : ctx->tables)
    if (t.c == c && t.n_cover >= n) {
      t.last_use = ctx->table_clock;
      *out = &t;
      return MIRA_OK;
    }
  // cover a power-of-two prefix (commit lengths recur; the prefix keeps small commits on small tables)
  size_t cover = 1;
  while (cover < n) cover <<= 1;
  if (cover > ctx->n_bases) cover = ctx->n_bases;
  // drop a smaller table for the same c
  for (auto it = ctx->tables.begin(); it != ctx->tables.end();) {
    if (it->c == c) {
      cudaFree(it->d);
      it = ctx->tables.erase(it);
    } else {
      ++it;
    }
  }
  const size_t need = (size_t)windows_for(c) * cover * 64;
  int rc = MIRA_ERR_CUDA;
  for (;;) {
    size_t free_b = 0, total_b = 0;
    bool fits = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && need + TABLE_RESERVE <= free_b;
    if (!fits && !ctx->tables.empty()) {          // evict the least recently used table and look again
      auto lru = ctx->tables.begin();
      for (auto it = ctx->tables.begin(); it != ctx->tables.end(); ++it)
        if (it->last_use < lru->last_use) lru = it;
      // ... unless it is the only fallback that covers n and the new table would not fit even without it
      if (lru->n_cover >= n && need + TABLE_RESERVE > free_b + lru->bytes) break;
      cudaFree(lru->d);
      ctx->tables.erase(lru);
      continue;
    }
    Table t;
    rc = build_table<CF>(ctx, c, (uint32_t)cover, &t, st);
    if (rc == MIRA_OK) {
      t.last_use = ctx->table_clock;
      ctx->tables.push_back(t);
      *out = &ctx->tables.back();
      return MIRA_OK;
    }
    break;
  }
  // fallback: the cached table covering n with the closest window
  Table* best = nullptr;
  for (auto& t : ctx->tables)
    if (t.n_cover >= n && (!best || std::abs(t.c - c) < std::abs(best->c - c))) best = &t;
  if (!best) return rc != MIRA_OK ? rc : fail(MIRA_ERR_CUDA, "no device memory for a fixed-base table of %zu bytes (window %d, %zu points)", need, c, cover);
  best->last_use = ctx->table_clock;
  *out = best;
  return MIRA_OK;
}

struct PhaseTimer {
  cudaEvent_t ev[6];
  bool on;
  cudaStream_t s;
  PhaseTimer(bool enabled, cudaStream_t st) : on(enabled), s(st) {
    if (on)
      for (auto& e : ev) cudaEventCreate(&e);
  }
  void mark(int i) {
    if (on) cudaEventRecord(ev[i], s);
  }
  float ms(int a, int b) {
    float v = 0;
    if (on) cudaEventElapsedTime(&v, ev[a], ev[b]);
    return v;
  }
  ~PhaseTimer() {
    if (on)
      for (auto& e : ev) cudaEventDestroy(e);
  }
};

// The device pipeline is split so that a host-buffer commit can run it slice by slice behind the H2D copies:
//   msm_begin   picks the window, fetches the fixed-base table, sizes the workspace, clears the buckets
//   msm_slice   digits -> sort -> accumulate for scalars [first, first + n_slice); add_mode folds into the buckets
//   msm_finish  bucket reduction; leaves the XYZZ sum (128 B) in ctx->result.p
constexpr int MAX_BATCH = 32;
// Window width for this vector: from a sample of the scalars when the vector is long enough for the choice to matter.
// `scalars` is a device pointer (on_device) or a host pointer; 64 chunks of 128 scalars spread over the vector are
// examined (host: 64 small H2D copies of 4 KiB).  Costs ~30 us (device) / ~0.2 ms (host) including the round trip.
constexpr size_t ADAPT_MIN_N = (size_t)1 << 18;
constexpr uint32_t ADAPT_CHUNKS = 64, ADAPT_CHUNK_LEN = 128;
template <class SF>
int pick_window(mira_msm_ctx* ctx, const void* scalars, size_t n, bool on_device, cudaStream_t st, int* window,
                double* pairs_per_scalar = nullptr) {
  *window = 0;
  if (pairs_per_scalar) *pairs_per_scalar = 0.0;        // 0 = not sampled
  ctx->sampled_pairs_per_scalar = 0.0;
  ctx->sampled_heavy = false;
  if (ctx->forced_window || !ctx->adaptive_window || n < ADAPT_MIN_N) return MIRA_OK;
  int rc;
  const size_t samples = (size_t)ADAPT_CHUNKS * ADAPT_CHUNK_LEN, stride = n / ADAPT_CHUNKS;
  if ((rc = ctx->sample.ensure(samples * 32 + (260 + SAMPLE_HASH_BINS) * 4))) return rc;
  uint32_t* d_hist = (uint32_t*)((char*)ctx->sample.p + samples * 32);
  CU(cudaMemsetAsync(d_hist, 0, (260 + SAMPLE_HASH_BINS) * 4, st));
  const void* src = scalars;
  size_t src_stride = stride;
  if (!on_device) {
    for (uint32_t k = 0; k < ADAPT_CHUNKS; k++)
      CU(cudaMemcpyAsync((char*)ctx->sample.p + (size_t)k * ADAPT_CHUNK_LEN * 32, (const char*)scalars + (size_t)k * stride * 32,
                         ADAPT_CHUNK_LEN * 32, cudaMemcpyDefault, st));
    src = ctx->sample.p;
    src_stride = ADAPT_CHUNK_LEN;
  }
  k_bitlen_hist<SF><<<(unsigned)((samples + 255) / 256), 256, 0, st>>>(src, ADAPT_CHUNKS, ADAPT_CHUNK_LEN, src_stride, d_hist);
  CU(cudaMemcpyAsync(ctx->h_hist, d_hist, (260 + SAMPLE_HASH_BINS) * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  {                             // heavy hitters: a hash bin far above the samples / bins background (8 of 8192 per bin)
    uint32_t mx = 0;
    for (int b = 0; b < SAMPLE_HASH_BINS; b++) mx = std::max(mx, ctx->h_hist[260 + b]);
    ctx->sampled_heavy = mx > samples / 64;              // one value in > 1/64 of the sampled scalars
  }
  double costs[25];
  int best = choose_window_sampled(n, ctx->h_hist, samples, costs);
  // hysteresis: a table costs seconds and gigabytes to build, so a window that already has one covering n is kept
  // unless the sample says it is clearly (> 5 %) worse
  int keep = 0;
  for (auto& t : ctx->tables)
    if (t.c >= 6 && t.c <= 24 && t.n_cover >= n && costs[t.c] <= 1.05 * costs[best] && (!keep || costs[t.c] < costs[keep])) keep = t.c;
  *window = keep ? keep : best;
  {                             // non-zero signed digits per scalar the sample predicts for the chosen window
    double pairs = 0;
    for (int L = 1; L <= 256; L++) pairs += (double)ctx->h_hist[L] * (double)((L + *window - 1) / *window);
    ctx->sampled_pairs_per_scalar = pairs / (double)samples;
    if (pairs_per_scalar) *pairs_per_scalar = ctx->sampled_pairs_per_scalar;
  }
  return MIRA_OK;
}

// Batched-affine levels before the XYZZ accumulation (affine_levels.cuh): off unless asked for
// (mira_msm_set_affine_levels, or MIRA_AFFINE_LEVELS in the environment for sweeps).  Measured on B200 at 2^24 points
// (201 M pairs, 96 per bucket): 0 levels 32.2 ms, 3 levels 32.2 ms, 5 levels 32.3 ms — the cheaper additions
// (0.10 vs 0.17 ns in isolation) are paid back by the second pass over the gathered points and the inversion kernels.
constexpr int PA_MAX_LEVELS = 6;
inline size_t pa_align(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
inline size_t pa_keys_bytes(size_t entries) { return pa_align(entries * 4); }
inline size_t pa_list_bytes(size_t entries) { return pa_keys_bytes(entries) + entries * 64 + 256; }
inline int affine_levels_for(const mira_msm_ctx* ctx) {
  static const int env = [] { const char* e = getenv("MIRA_AFFINE_LEVELS"); return e ? atoi(e) : 0; }();
  int want = ctx->affine_levels >= 0 ? ctx->affine_levels : env;
  return want < 0 ? 0 : (want > PA_MAX_LEVELS ? PA_MAX_LEVELS : want);
}

// Entries per accumulation thread: as long as possible (each chunk edge that falls inside a bucket's run costs one XYZZ
// full add in k_combine, and a thread's first pair pays its load latency alone) while keeping the machine full: two
// waves of 148 SMs x 512 resident threads below 32 M pairs, four above, never fewer than 32 pairs per thread
// (measured sweep, 2^16..2^22 points: 2^19 2.07 -> 1.64 ms, 2^16 0.40 -> 0.30 ms against the former 4 waves / 16).
inline int acc_chunk_len(size_t EA) {
  static const unsigned waves_env = [] { const char* e = getenv("MIRA_ACC_WAVES"); return e ? (unsigned)atoi(e) : 0u; }();
  static const int lmin = [] { const char* e = getenv("MIRA_ACC_LMIN"); return e ? atoi(e) : 32; }();
  const unsigned waves = waves_env ? waves_env : (EA < ((size_t)32 << 20) ? 2u : 4u);
  int L = (int)(EA / (148u * 512u * waves));
  L = L < lmin ? lmin : (L > 256 ? 256 : L);
  // Long lists run many waves of 148 SMs x 4 resident blocks x 128 threads; a last wave that is only partly full runs
  // at low occupancy (2^24 points: 10.38 waves at L = 256).  Shorten the chunks a little so that the thread count is
  // just under a whole number of waves (MIRA_ACC_WAVE_FIT=0 turns it off).
  static const int fit = [] { const char* e = getenv("MIRA_ACC_WAVE_FIT"); return e ? atoi(e) : 1; }();
  if (fit && L >= 64) {
    const double per_wave = 148.0 * 4.0 * 128.0;
    const double w = (double)EA / (L * per_wave);
    const double target = std::ceil(w);
    if (target >= 2.0 && w < target - 0.04) {
      int L2 = (int)std::ceil((double)EA / (target * per_wave - 64.0));
      if (L2 >= lmin && L2 <= L) L = L2;
    }
  }
  return L;
}

struct MsmPlan {
  struct Prepared {                 // a decomposed and sorted slice, waiting in buffer set `bs` for its accumulation
    const uint32_t* skeys = nullptr;
    const uint32_t* srefs = nullptr;
    uint32_t* d_npairs = nullptr;
    size_t E = 0;                   // host-side upper bound of *d_npairs (the exact count for sparse vectors)
    size_t E_chunking = 0;          // pair count the accumulation's chunk length is chosen for
  } prep[2];
  int affine_levels = 0;
  bool heavy = false;           // the sample found a heavy hitter: keep the LSD sort (the MSD partition's per-group pass would crawl)
  bool exact_count = false;     // sparse vector: read the pair count back after the digit kernel and size everything from it
  int c = 0, W = 0;
  Table* tab = nullptr;
  uint32_t B = 0;
  int n_sets = 1;           // bucket sets (scalar vectors committed together against the same key)
  size_t max_slice = 0;
  uint64_t launches = 0;
  uint64_t entries = 0;
};

template <class CF>
int msm_begin(mira_msm_ctx* ctx, size_t n, size_t max_slice, cudaStream_t st, MsmPlan* plan, int n_sets = 1, int window = 0,
              int n_bufsets = 1) {
  int rc;
  int c = ctx->forced_window ? ctx->forced_window : (window ? window : choose_window(n));
  Table* tab = nullptr;
  if ((rc = get_table<CF>(ctx, c, n, st, &tab))) return rc;
  c = tab->c;                 // the cache may have handed out a neighbouring window (memory budget)
  const int W = tab->W;
  const size_t E = max_slice * (size_t)W * (size_t)n_sets;
  const uint32_t B = 1u << (c - 1);
  if (n * (size_t)W * (size_t)n_sets >= (size_t)0x7fffffff || (size_t)W * tab->n_cover >= (size_t)0x7fffffff ||
      ((size_t)B + 1) * (size_t)n_sets >= (size_t)PK_KEY_MASK)
    return fail(MIRA_ERR_INVALID, "commit of %d x %zu scalars needs %zu (point, window) pairs: exceeds the 2^31 reference space; shard it",
                n_sets, n, n * (size_t)W * (size_t)n_sets);
  const size_t bucket_bytes = ((size_t)B + 1) * 128 * (size_t)n_sets;
  for (int b = 0; b < n_bufsets; b++) {
    auto& sb = ctx->sb[b];
    if ((rc = sb.keys.ensure(E * 4 + 16)) || (rc = sb.refs.ensure(E * 4 + 16)) || (rc = sb.skeys.ensure(E * 4 + 16)) ||
        (rc = sb.srefs.ensure(E * 4 + 16)) || (rc = sb.counts.ensure((size_t)1 << 20)) || (rc = sb.tile_sums.ensure(radix_sort_temp_bytes(E))))
      return rc;
  }
  if ((rc = ctx->buckets.ensure(bucket_bytes))) return rc;
  {
    // the accumulation's scratch is sized here, once, from the largest slice: growing it between slices would
    // cudaFree (a device-wide synchronisation) in the middle of the slice pipeline
    const int L = acc_chunk_len(E);
    const size_t n_chunks = (E + L - 1) / L, heavy_cap = n_chunks / HEAVY_CHUNKS + 2;
    if ((rc = ctx->part_keys.ensure(n_chunks * 8)) || (rc = ctx->part_pts.ensure(n_chunks * 256)) ||
        (rc = ctx->cursor.ensure((heavy_cap * 2 + 4) * 4)))
      return rc;
  }
  CU(cudaMemsetAsync(ctx->buckets.p, 0, bucket_bytes, st));
  plan->c = c; plan->W = W; plan->tab = tab; plan->B = B; plan->max_slice = max_slice; plan->n_sets = n_sets;
  // A sparse vector (the sample predicts well under W pairs per scalar: witness columns are ~60 % zeros) fills only a
  // fraction of the n * W upper bound the launches are sized from: the sort would launch mostly empty tiles (2^24
  // witness-like scalars: 1.10 ms of sort for 18 M pairs, 0.48 ms when sized right).  One 4-byte read-back after the
  // digit kernel (the host synchronises once more, ~30 us) sizes the sort and the accumulation grid from the real count.
  static const int exact_on = [] { const char* e = getenv("MIRA_EXACT_COUNT"); return e ? atoi(e) : 1; }();
  plan->exact_count = exact_on && n_sets == 1 && ctx->sampled_pairs_per_scalar > 0.0 && ctx->sampled_pairs_per_scalar < 0.6 * W;
  plan->heavy = ctx->sampled_heavy;
  ctx->sampled_heavy = false;
  ctx->sampled_pairs_per_scalar = 0.0;          // belongs to the commit that sampled it (pick_window runs right before msm_begin)
  plan->launches = 0; plan->entries = 0;
  return MIRA_OK;
}

// First half of a slice: scalars [first, first + n) -> sorted (bucket, point reference) pairs in buffer set `bs`.
template <class CF, class SF>
int msm_prep(mira_msm_ctx* ctx, MsmPlan* plan, const void* const* d_scalar_sets, size_t first, size_t n, bool add_mode, int bs,
             cudaStream_t st, PhaseTimer* pt) {
  int rc;
  const int c = plan->c, W = plan->W;
  Table* tab = plan->tab;
  auto& sb = ctx->sb[bs];
  const size_t E = n * (size_t)W * (size_t)plan->n_sets;
  uint32_t* d_npairs = (uint32_t*)sb.counts.p;     // number of (bucket, ref) pairs, produced on the device
  if (pt) pt->mark(0);
  // ---- digits.  Default (round 2): fused with the sort's first pass — histogram of the low key byte, scan, then the
  // digit kernel scatters its pairs into the 256 bins (k_digits_scatter).  MIRA_FUSED_DIGITS=0 keeps the compacted
  // list + full sort of round 1 (A/B measurements).
  // Sparse vectors (exact_count: the sample predicts well under W pairs per scalar) keep the compacted list: their
  // sort is cheap because the pairs are few, and the fused form decomposes every scalar three times instead of twice
  // (2^24 witness-like scalars: 5.36 ms against 6.13 ms fused).
  static const int fused_env = [] { const char* e = getenv("MIRA_FUSED_DIGITS"); return e ? atoi(e) : 1; }();
  const bool fused_on = fused_env == 2 || (fused_env == 1 && !plan->exact_count);
  if (!add_mode) CU(cudaMemsetAsync((char*)sb.counts.p + 48, 0, 8, st));      // affine additions of this commit
  uint32_t* d_hist = (uint32_t*)((char*)sb.counts.p + 1024);                  // [256] low-byte histogram
  uint32_t* d_cursor = (uint32_t*)((char*)sb.counts.p + 2048);                // [256] next free slot of each bin
  // MSD partition instead of the LSD passes (msm_kernels.cuh): for 17..24 key bits, from MSD_MIN_PAIRS pairs (below that
  // its fixed costs — 65,536 groups, two single-block scans — outweigh the cheaper ranking) and while an average group
  // fits k_msd_low's shared-memory path (larger commits keep the LSD passes), and not when the sample of the vector shows
  // a heavy hitter (a value repeated in > 1/64 of the scalars puts hundreds of thousands of pairs into W single
  // buckets, and k_msd_low sorts a group with ONE block).  MIRA_SORT_MSD=0 keeps the LSD sort.
  static const int msd_env = [] { const char* e = getenv("MIRA_SORT_MSD"); return e ? atoi(e) : 1; }();
  static const size_t msd_min_pairs = [] { const char* e = getenv("MIRA_SORT_MSD_MIN_LOG"); return (size_t)1 << (e ? atoi(e) : 25); }();
  int key_bits = c;          // keys are < n_sets * (B + 1)
  while (((uint64_t)1 << key_bits) < (uint64_t)plan->n_sets * (plan->B + 1)) key_bits++;
  int msd_bits = 1;          // the MSD partition works on key - 1 < n_sets * (B + 1) - 1 (a single set: c - 1 bits, evenly filled)
  while (((uint64_t)1 << msd_bits) < (uint64_t)plan->n_sets * (plan->B + 1) - 1) msd_bits++;
  const bool msd_on = fused_on && msd_env && !plan->heavy && msd_bits >= MSD_GROUP_BITS + 1 && msd_bits <= MSD_GROUP_BITS + 8 && E >= msd_min_pairs &&
                      (E >> MSD_GROUP_BITS) <= (size_t)MSD_LOW_MAX_AVG_CHUNKED;
  const int low_bits = msd_bits - MSD_GROUP_BITS;
  uint32_t* d_offs1 = (uint32_t*)((char*)sb.counts.p + 4096);                 // [257] segment starts
  uint32_t* d_tile_tab = d_offs1 + 320;                                       // [257]
  uint32_t* d_hist16 = (uint32_t*)((char*)sb.counts.p + 8192);                // [65536]
  uint32_t* d_offs16 = d_hist16 + MSD_GROUPS;                                 // [65537] (+ padding)
  uint32_t* d_cursor16 = d_offs16 + MSD_GROUPS + 64;                          // [65536]
  if (msd_on) {
    CU(cudaMemsetAsync(d_hist, 0, 1024, st));
    CU(cudaMemsetAsync(d_hist16, 0, (size_t)MSD_GROUPS * 4, st));
    const unsigned hist_blocks = (unsigned)std::min<size_t>((n + DS_THREADS - 1) / DS_THREADS, 148 * 8);
    for (int s = 0; s < plan->n_sets; s++) {
      k_digit_hist<SF><<<hist_blocks, DS_THREADS, 0, st>>>(d_scalar_sets[s], (uint32_t)n, c, W, (uint32_t)s * (plan->B + 1), msd_bits - 8, 1u, d_hist);
      plan->launches++;
    }
    k_msd_scan1<<<1, DS_THREADS, 0, st>>>(d_hist, d_offs1, d_cursor, d_tile_tab, d_npairs);
    plan->launches++;
  } else if (fused_on) {
    CU(cudaMemsetAsync(d_hist, 0, 1024, st));
    const unsigned hist_blocks = (unsigned)std::min<size_t>((n + DS_THREADS - 1) / DS_THREADS, 148 * 8);
    for (int s = 0; s < plan->n_sets; s++) {
      k_digit_hist<SF><<<hist_blocks, DS_THREADS, 0, st>>>(d_scalar_sets[s], (uint32_t)n, c, W, (uint32_t)s * (plan->B + 1), 0, 0u, d_hist);
      plan->launches++;
    }
    k_digit_scan<<<1, DS_THREADS, 0, st>>>(d_hist, d_cursor, d_npairs);
    plan->launches++;
  } else {
    CU(cudaMemsetAsync(d_npairs, 0, 4, st));
    for (int s = 0; s < plan->n_sets; s++) {
      k_digits<SF><<<(unsigned)((n + DG_THREADS - 1) / DG_THREADS), DG_THREADS, (size_t)W * DG_WARPS * 4, st>>>(
          d_scalar_sets[s], (uint32_t)n, (uint32_t)first, c, W, tab->n_cover, (uint32_t)s * (plan->B + 1), (uint32_t*)sb.keys.p,
          (uint32_t*)sb.refs.p, d_npairs);
      plan->launches++;
    }
  }
  size_t E_sort = E;
  if (plan->exact_count) {
    uint32_t* h_count = ctx->h_hist + 258;          // pinned
    CU(cudaMemcpyAsync(h_count, d_npairs, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    E_sort = *h_count ? *h_count : 1;               // an all-zero vector still launches (and finds nothing to do)
    if (E_sort > E) return fail(MIRA_ERR_CUDA, "pair count %zu exceeds its bound %zu", E_sort, E);
  }
  int in_b = 0;
  if (fused_on) {
    const uint32_t spb = ds_scalars_per_block(W);
    const size_t dyn = (size_t)spb * 9 * 4 + (size_t)spb * W * 8;
    static std::once_flag ds_once[64];
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    std::call_once(ds_once[dev_id & 63], [&] {
      // the smallest block (32 scalars) of the narrowest window (c = 2: 128 digits per scalar) stages 4096 pairs
      cudaFuncSetAttribute(k_digits_scatter<SF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           std::max(DS_SPB_MAX * 9 * 4 + DS_CAP * 8, 32 * 9 * 4 + 32 * 128 * 8));
    });
    for (int s = 0; s < plan->n_sets; s++) {
      k_digits_scatter<SF><<<(unsigned)((n + spb - 1) / spb), DS_THREADS, dyn, st>>>(
          d_scalar_sets[s], (uint32_t)n, (uint32_t)first, c, W, tab->n_cover, (uint32_t)s * (plan->B + 1), spb, msd_on ? msd_bits - 8 : 0,
          msd_on ? 1u : 0u, d_cursor, (uint32_t*)sb.skeys.p, (uint32_t*)sb.srefs.p);
      plan->launches++;
    }
    if (pt) pt->mark(1);
    if (msd_on) {
      // ---- MSD: group counts of the next 8 bits, the partition by them (s-buffers -> plain buffers), then every
      // 16-bit group by its low bits (back into the s-buffers)
      const unsigned tiles = (unsigned)(E_sort / MSD_TILE + 257);
      k_msd_hist16<<<tiles, MSD_M_THREADS, 0, st>>>((const uint32_t*)sb.skeys.p, d_offs1, d_tile_tab, low_bits, d_hist16);
      k_msd_scan16<<<1, MSD_S_THREADS, 0, st>>>(d_hist16, d_offs16, d_cursor16);
      k_msd_mid<<<tiles, MSD_M_THREADS, 0, st>>>((const uint32_t*)sb.skeys.p, (const uint32_t*)sb.srefs.p, d_offs1, d_tile_tab, d_cursor16,
                                                 low_bits, (uint32_t*)sb.keys.p, (uint32_t*)sb.refs.p);
      // groups per k_msd_low block: as many as keep the block's range around 3,000 pairs and its bins within 256
      uint32_t gpb = 1;
      for (int span = low_bits; span < 8 && (E_sort / MSD_GROUPS) * (gpb * 2) <= 3000; span++) gpb <<= 1;
      k_msd_low<<<MSD_GROUPS / gpb, MSD_L_THREADS, 0, st>>>((const uint32_t*)sb.keys.p, (const uint32_t*)sb.refs.p, d_offs16, gpb, low_bits,
                                                           (uint32_t*)sb.skeys.p, (uint32_t*)sb.srefs.p);
      plan->launches += 4;
      CU(cudaGetLastError());
      in_b = 1;
    } else {
      // ---- the remaining passes (bit 8 up) of the LSD radix sort, from the s-buffers into the plain ones and back
      int in_plain = 0;
      if ((rc = radix_sort_pairs((uint32_t*)sb.skeys.p, (uint32_t*)sb.srefs.p, (uint32_t*)sb.keys.p, (uint32_t*)sb.refs.p, d_npairs,
                                 E_sort, key_bits, sb.tile_sums.p, st, &in_plain, &plan->launches, 1)))
        return rc;
      in_b = !in_plain;
    }
  } else {
    if (pt) pt->mark(1);
    // ---- group pairs by bucket: LSD radix sort on the c-bit key
    if ((rc = radix_sort_pairs((uint32_t*)sb.keys.p, (uint32_t*)sb.refs.p, (uint32_t*)sb.skeys.p, (uint32_t*)sb.srefs.p, d_npairs, E_sort,
                               key_bits, sb.tile_sums.p, st, &in_b, &plan->launches)))
      return rc;
  }
  plan->prep[bs].skeys = (const uint32_t*)(in_b ? sb.skeys.p : sb.keys.p);
  plan->prep[bs].srefs = (const uint32_t*)(in_b ? sb.srefs.p : sb.refs.p);
  plan->prep[bs].d_npairs = d_npairs;
  plan->prep[bs].E = E_sort;
  // The chunk length stays the one of the n * W bound: sparse vectors pile most of their pairs into a few buckets, and
  // every chunk inside such a run leaves a partial sum that the heavy-run kernels add up run by run; shorter chunks
  // measured slower there (2^24 witness-like scalars: accumulation phase 3.69 ms with L = 256, 4.20 ms with L = 60).
  plan->prep[bs].E_chunking = E;
  if (pt) pt->mark(2);
  return MIRA_OK;
}

// Second half: the sorted pairs of buffer set `bs` are added into the buckets (add_mode: on top of earlier slices).
template <class CF>
int msm_acc(mira_msm_ctx* ctx, MsmPlan* plan, bool add_mode, int bs, cudaStream_t st, PhaseTimer* pt) {
  int rc;
  Table* tab = plan->tab;
  auto& sb = ctx->sb[bs];
  const uint32_t* skeys = plan->prep[bs].skeys;
  const uint32_t* srefs = plan->prep[bs].srefs;
  uint32_t* d_npairs = plan->prep[bs].d_npairs;
  const size_t E = plan->prep[bs].E;
  // ---- batched-affine levels: add the entries of every bucket two by two while the list is long (affine_levels.cuh)
  const uint32_t* acc_keys = skeys;
  const uint32_t* acc_refs = srefs;
  const void* acc_points = tab->d;
  const uint32_t* acc_n = d_npairs;
  size_t acc_bound = E;            // upper bound of *acc_n known to the host (grids are sized from it)
  bool acc_direct = false;
  {
    const size_t key_space = (size_t)plan->n_sets * ((size_t)plan->B + 1);
    int levels = affine_levels_for(ctx);
    size_t bound[PA_MAX_LEVELS + 1];
    bound[0] = E;
    for (int l = 0; l < levels; l++) {
      // every run leaves ceil(len / 2) entries; runs <= keys + tiles (a run is cut at tile edges)
      size_t tiles = (bound[l] + PA_TILE - 1) / PA_TILE;
      size_t runs = std::min(bound[l], key_space + tiles);
      bound[l + 1] = (bound[l] + runs + 1) / 2;
    }
    const size_t tiles0 = (bound[0] + PA_TILE - 1) / PA_TILE, threads0 = tiles0 * PA_THREADS;
    if (levels > 0) {
      const size_t need_a = pa_list_bytes(bound[1]), need_b = levels > 1 ? pa_list_bytes(bound[2]) : 0;
      const size_t need_w = pa_align((tiles0 + 1) * 8) + pa_align(threads0 * sizeof(PaMeta)) + threads0 * 64 + threads0 * PA_MAX_ADDS * 32 + 256;
      size_t grow = (need_a > ctx->pa_a.cap ? need_a : 0) + (need_b > ctx->pa_b.cap ? need_b : 0) + (need_w > ctx->pa_work.cap ? need_w : 0);
      size_t free_b = ~(size_t)0, total_b = 0;
      if (grow) cudaMemGetInfo(&free_b, &total_b);             // only when something has to be allocated
      if (grow && grow + ((size_t)2 << 30) > free_b) levels = 0;      // not worth evicting anything: the XYZZ path needs no extra memory
      else if ((rc = ctx->pa_a.ensure(need_a)) || (rc = ctx->pa_b.ensure(need_b)) || (rc = ctx->pa_work.ensure(need_w)))
        return rc;
    }
    uint32_t* d_counts = (uint32_t*)sb.counts.p;               // [0] pairs, [1 + l] list length after level l
    unsigned long long* d_adds = (unsigned long long*)((char*)sb.counts.p + 48);
    for (int l = 0; l < levels; l++) {
      const unsigned tiles = (unsigned)((bound[l] + PA_TILE - 1) / PA_TILE);
      const size_t threads = (size_t)tiles * PA_THREADS;
      // pa_work: tile counts | tile offsets | per-thread meta | per-thread products | their inverses | per-addition prefixes
      char* w = (char*)ctx->pa_work.p;
      uint32_t* tile_cnt = (uint32_t*)w;
      uint32_t* tile_off = tile_cnt + tiles0 + 1;
      PaMeta* meta = (PaMeta*)(w + pa_align((size_t)(tiles0 + 1) * 8));
      void* runs = (char*)meta + pa_align(threads0 * sizeof(PaMeta));
      void* invs = (char*)runs + threads0 * 32;
      void* pref = (char*)invs + threads0 * 32;
      DevBuf& out = (l & 1) ? ctx->pa_b : ctx->pa_a;
      uint32_t* keys_out = (uint32_t*)out.p;
      void* pts_out = (char*)out.p + pa_keys_bytes(bound[l + 1]);
      const unsigned inv_blocks = (unsigned)((threads / PA_INV_GROUP + 127) / 128 + 1);
      if (l == 0) {
        k_pa_up<CF, true><<<tiles, PA_THREADS, 0, st>>>(acc_keys, acc_refs, nullptr, tab->d, acc_n, tile_cnt, meta, pref, runs, d_adds);
        k_pa_scan<<<1, 1024, 0, st>>>(acc_n, tile_cnt, tile_off, d_counts + 1 + l);
        k_pa_inv<CF><<<inv_blocks, 128, 0, st>>>(runs, acc_n, invs);
        k_pa_down<CF, true><<<tiles, PA_THREADS, 0, st>>>(acc_keys, acc_refs, nullptr, tab->d, acc_n, tile_off, meta, pref, invs, keys_out, pts_out);
      } else {
        k_pa_up<CF, false><<<tiles, PA_THREADS, 0, st>>>(acc_keys, nullptr, acc_points, nullptr, acc_n, tile_cnt, meta, pref, runs, d_adds);
        k_pa_scan<<<1, 1024, 0, st>>>(acc_n, tile_cnt, tile_off, d_counts + 1 + l);
        k_pa_inv<CF><<<inv_blocks, 128, 0, st>>>(runs, acc_n, invs);
        k_pa_down<CF, false><<<tiles, PA_THREADS, 0, st>>>(acc_keys, nullptr, acc_points, nullptr, acc_n, tile_off, meta, pref, invs, keys_out, pts_out);
      }
      plan->launches += 4;
      acc_keys = keys_out;
      acc_refs = nullptr;
      acc_points = pts_out;
      acc_n = d_counts + 1 + l;
      acc_bound = bound[l + 1];
      acc_direct = true;
    }
    plan->affine_levels = levels;
  }
  // ---- accumulate
  {
    // Sized from the upper bound n*W; grids cover that bound and surplus threads exit on the device-side pair count.
    const size_t EA = acc_bound;
    const int L = acc_chunk_len(acc_direct ? EA : std::max(EA, plan->prep[bs].E_chunking));
    uint32_t n_chunks = (uint32_t)((EA + L - 1) / L);
    if ((rc = ctx->part_keys.ensure((size_t)n_chunks * 8)) || (rc = ctx->part_pts.ensure((size_t)n_chunks * 256))) return rc;
    // MIRA_ACC_PAD_KB (development knob): unused dynamic shared memory per accumulation block, to cap the blocks
    // per SM and leave room for the sort blocks of the NEXT slice (overlap experiments)
    static const int pad_kb = [] { const char* e = getenv("MIRA_ACC_PAD_KB"); return e ? atoi(e) : 0; }();
    if (pad_kb > 0) {
      static std::once_flag pad_once[64];
      int dev_id = 0;
      cudaGetDevice(&dev_id);
      std::call_once(pad_once[dev_id & 63], [&] {
        cudaFuncSetAttribute(k_accumulate<CF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad_kb * 1024);
      });
    }
    // Thread-local pair pre-addition (k_pair_up / k_pair_add / k_pair_acc, msm_kernels.cuh): OFF by default — measured
    // slower than the plain kernel on B200 (2^24 points: 6.1 + 13.8 + 16.3 = 36.2 ms against 31.8 ms; the two extra
    // gathers of every table point cost 134 B of DRAM traffic each and make the first two kernels memory-bound,
    // profiles/r02_pair_preadd.txt).  mira_msm_set_affine_levels(ctx, MIRA_AFFINE_THREAD_LOCAL_PAIRS) turns it on for
    // a context (any chunk length: what the parity tests use), MIRA_ACC_PAIR=1 for a process (chunks of >= 128 pairs,
    // MIRA_ACC_PAIR_MIN_L moves that).
    static const int pair_env = [] { const char* e = getenv("MIRA_ACC_PAIR"); return e ? atoi(e) : 0; }();
    static const int pair_env_min_l = [] { const char* e = getenv("MIRA_ACC_PAIR_MIN_L"); return e ? atoi(e) : 128; }();
    const bool pair_ctx = ctx->affine_levels == MIRA_AFFINE_THREAD_LOCAL_PAIRS;
    const bool pair_on = pair_ctx || (pair_env && ctx->affine_levels < 0);
    const int pair_min_l = pair_ctx ? 2 : pair_env_min_l;
    const unsigned acc_blocks = (n_chunks + 127) / 128;
    bool use_pair = pair_on && !acc_direct && L >= pair_min_l && pad_kb <= 0;
    if (use_pair) {
      const size_t pa_threads = (size_t)acc_blocks * 128, pa_slots = (size_t)((L + 1) / 2) * pa_threads;     // [pair][thread]
      const size_t need = pa_slots * (32 + 64) + pa_threads * 32 + 512;       // suffix products | sums | per-thread inverses
      if (need > ctx->pa_work.cap) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || need + ((size_t)2 << 30) > free_b + ctx->pa_work.cap) use_pair = false;
      }
      if (use_pair && (rc = ctx->pa_work.ensure(need))) return rc;
    }
    if (use_pair) {
      const size_t pa_threads = (size_t)acc_blocks * 128, pa_slots = (size_t)((L + 1) / 2) * pa_threads;
      char* sfx = (char*)ctx->pa_work.p;
      char* sums = sfx + ((pa_slots * 32 + 255) & ~(size_t)255);
      char* invs = sums + ((pa_slots * 64 + 255) & ~(size_t)255);
      k_pair_up<CF><<<acc_blocks, 128, 0, st>>>(acc_keys, acc_refs, acc_n, L, acc_points, sfx, invs);
      k_pair_add<CF><<<acc_blocks, 128, 0, st>>>(acc_keys, acc_refs, acc_n, L, acc_points, sfx, invs, sums);
      k_pair_acc<CF><<<acc_blocks, 128, 0, st>>>(acc_keys, acc_refs, acc_n, L, acc_points, ctx->buckets.p, (uint32_t*)ctx->part_keys.p,
                                                ctx->part_pts.p, add_mode ? 1 : 0, sums);
      plan->launches += 2;
    }
    else if (acc_direct)
      k_accumulate<CF, true><<<(n_chunks + 127) / 128, 128, 0, st>>>(acc_keys, nullptr, acc_n, L, acc_points, ctx->buckets.p,
                                                                    (uint32_t*)ctx->part_keys.p, ctx->part_pts.p, add_mode ? 1 : 0);
    else
      k_accumulate<CF, false><<<(n_chunks + 127) / 128, 128, (size_t)(pad_kb > 0 ? pad_kb : 0) * 1024, st>>>(
          acc_keys, acc_refs, acc_n, L, acc_points, ctx->buckets.p, (uint32_t*)ctx->part_keys.p, ctx->part_pts.p, add_mode ? 1 : 0);
    uint32_t heavy_cap = n_chunks / HEAVY_CHUNKS + 2;
    if ((rc = ctx->cursor.ensure(((size_t)heavy_cap * 2 + 4) * 4))) return rc;
    uint32_t* d_heavy = (uint32_t*)ctx->cursor.p;     // [0], [1] = counts, then the medium and the huge leader lists
    CU(cudaMemsetAsync(d_heavy, 0, 8, st));
    k_combine<CF><<<(n_chunks + 127) / 128, 128, 0, st>>>(acc_keys, (const uint32_t*)ctx->part_keys.p, ctx->part_pts.p, acc_n, L,
                                                             ctx->buckets.p, d_heavy, heavy_cap);
    k_combine_heavy<CF, 32><<<148 * 2, HV_THREADS, 0, st>>>(acc_keys, (const uint32_t*)ctx->part_keys.p, ctx->part_pts.p, acc_n, L,
                                                           ctx->buckets.p, d_heavy, heavy_cap);
    k_combine_heavy<CF, HV_THREADS><<<148 * 2, HV_THREADS, 0, st>>>(acc_keys, (const uint32_t*)ctx->part_keys.p, ctx->part_pts.p, acc_n,
                                                                   L, ctx->buckets.p, d_heavy, heavy_cap);
    plan->launches += 4;
  }
  if (pt) pt->mark(3);
  plan->entries += E;
  return MIRA_OK;
}

// One slice on one stream (batched commits, profiled commits, small commits).
template <class CF, class SF>
int msm_slice(mira_msm_ctx* ctx, MsmPlan* plan, const void* const* d_scalar_sets, size_t first, size_t n, bool add_mode,
              cudaStream_t st, PhaseTimer* pt) {
  int rc;
  if ((rc = msm_prep<CF, SF>(ctx, plan, d_scalar_sets, first, n, add_mode, 0, st, pt))) return rc;
  return msm_acc<CF>(ctx, plan, add_mode, 0, st, pt);
}

// Software pipeline over slices: the digits and the sort of slice k+1 run on ctx->prep_stream while slice k is
// accumulated on `st`.  The two are complementary on paper — the sort is latency/bandwidth-bound on the ALU and LSU
// pipes, the accumulation integer-multiply-bound with 11 % of the DRAM bandwidth — and two INDEPENDENT commits do
// overlap that way (tools/overlap_probe.py: 42.9 -> 39.8 ms per pair of 2^23-point commits).  Inside one commit the
// slices pay for it: every later slice spends one extra mixed addition per non-empty bucket, and the 512-thread sort
// blocks only get onto an SM when several accumulation blocks have retired.  Measured at 2^24 points, device-resident:
// 1 slice 39.9 ms, 2 slices 40.5-41.2, 4 slices 41.6-42.9 (low / high priority prep stream), 8 slices 44.4.  So a
// device-resident commit is NOT sliced by default (mira_msm_set_pipeline turns it on); a page-locked host-buffer
// commit, which is cut into H2D slices anyway, does overlap the preparation of slice k+1 with the accumulation of
// slice k (2^24: 41.96 -> 41.47 ms end to end).  `ready[k]`, if given, is an event the preparation of slice k has to
// wait for (the H2D copy).
inline int pipe_slices_for(const mira_msm_ctx* ctx, size_t n) {
  static const int env = [] { const char* e = getenv("MIRA_PIPE_SLICES"); return e ? atoi(e) : 1; }();
  int k = ctx->pipe_slices > 0 ? ctx->pipe_slices : env;
  if (k < 1) k = 1;
  if (k > 16) k = 16;
  while (k > 1 && n / (size_t)k < ctx->pipe_min_slice) k--;
  return k;
}
inline int pipe_setup(mira_msm_ctx* ctx) {
  if (!ctx->prep_stream) {
    // MIRA_PREP_PRIO=1: the preparation stream outranks the accumulation, so its (short) sort blocks are placed
    // ahead of the accumulation blocks still waiting for an SM
    static const int prio = [] { const char* e = getenv("MIRA_PREP_PRIO"); return e ? atoi(e) : 0; }();
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&ctx->prep_stream, cudaStreamNonBlocking, prio ? hi : lo));
  }
  for (cudaEvent_t* e : {&ctx->prep_done[0], &ctx->prep_done[1], &ctx->acc_done[0], &ctx->acc_done[1], &ctx->pipe_start})
    if (!*e) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  return MIRA_OK;
}
template <class CF, class SF>
int msm_pipeline(mira_msm_ctx* ctx, MsmPlan* plan, const void* d_scalars, const size_t* bounds, int n_slices, cudaEvent_t* ready,
                 cudaStream_t st) {
  int rc;
  if ((rc = pipe_setup(ctx))) return rc;
  // the preparation must not start before work already queued on `st` (it may produce the scalars) has finished
  CU(cudaEventRecord(ctx->pipe_start, st));
  CU(cudaStreamWaitEvent(ctx->prep_stream, ctx->pipe_start, 0));
  for (int k = 0; k < n_slices; k++) {
    const int bs = k & 1;
    const size_t first = bounds[k], cnt = bounds[k + 1] - bounds[k];
    if (ready) CU(cudaStreamWaitEvent(ctx->prep_stream, ready[k], 0));
    if (k >= 2) CU(cudaStreamWaitEvent(ctx->prep_stream, ctx->acc_done[bs], 0));     // buffer set free again
    const void* sets[1] = {(const char*)d_scalars + first * 32};
    if ((rc = msm_prep<CF, SF>(ctx, plan, sets, first, cnt, k > 0, bs, ctx->prep_stream, nullptr))) return rc;
    CU(cudaEventRecord(ctx->prep_done[bs], ctx->prep_stream));
    CU(cudaStreamWaitEvent(st, ctx->prep_done[bs], 0));
    if ((rc = msm_acc<CF>(ctx, plan, k > 0, bs, st, nullptr))) return rc;
    CU(cudaEventRecord(ctx->acc_done[bs], st));
  }
  return MIRA_OK;
}

// Chunk width (log2) of a reduction level over n elements; 0 = small enough for the closing k_reduce_chunks pass.
// MIRA_RED_LEVELS=0 turns the level passes off (the round-1 single pass).
inline int reduce_level_log_m(uint32_t n, int levels_done, unsigned n_sets) {
  static const int on = [] { const char* e = getenv("MIRA_RED_LEVELS"); return e ? atoi(e) : 1; }();
  // measured (profiles/r02_reduce_levels.txt): 2^21 buckets 2.10 -> 1.75 ms, 2^19 0.71 -> 0.71; below that a level
  // only adds latency (2^16 buckets 0.45 -> 0.52 ms, 2^14 0.30 -> 0.40), so small bucket sets keep the single pass
  // (ncu, profiles/r02_accumulate_v1.txt: after one level of a 2^21-bucket set the closing pass over the remaining
  // 131,071 elements still took 0.58 ms against 1.18 ms for the level itself, so a big reduction keeps levelling down
  // to 2^13 elements)
  if (!on) return 0;
  // MIRA_RED_MIN_LOG: log2 of the total bucket count (all sets of a batched commit) from which the first level pays:
  // the single pass spends ~370 products per thread on its weighting, which is throughput, not latency, once a batch
  // of 5-6 sets brings a few hundred thousand threads
  static const int min_log = [] { const char* e = getenv("MIRA_RED_MIN_LOG"); return e ? atoi(e) : 17; }();
  static const int min_n_log = [] { const char* e = getenv("MIRA_RED_MIN_N_LOG"); return e ? atoi(e) : 14; }();
  static const int batch_log_m = [] { const char* e = getenv("MIRA_RED_LOG_M"); return e ? atoi(e) : 0; }();
  if (levels_done == 0) {
    // single commits: MIRA_RED1_MIN_LOG / MIRA_RED1_LOG_M / MIRA_RED1_NEXT_MIN_LOG move the thresholds (sweeps)
    static const int one_min_log = [] { const char* e = getenv("MIRA_RED1_MIN_LOG"); return e ? atoi(e) : 20; }();
    static const int one_log_m = [] { const char* e = getenv("MIRA_RED1_LOG_M"); return e ? atoi(e) : 4; }();
    if (n_sets == 1) return n >= ((uint32_t)1 << one_min_log) ? one_log_m : 0;
    // batched commits (profiles/r02_reduce_levels.txt): 6 x 2^19 points (c = 17) 11.39 -> 9.78 ms, 6 x 2^16 (c = 15)
    // 2.29 -> 2.09 ms with chunks of 8 (2.18 with 16; the larger sets prefer 16: 9.80 against 9.90); below 2^17 buckets in
    // all the cooperative kernels (coop.cuh) are a little faster still (6 x 2^16 points: 2.04 ms)
    if ((uint64_t)n * n_sets < ((uint64_t)1 << min_log) || n < ((uint32_t)1 << min_n_log)) return 0;
    return batch_log_m ? batch_log_m : (n < ((uint32_t)1 << 15) ? 3 : 4);
  }
  static const int next_min_log = [] { const char* e = getenv("MIRA_RED1_NEXT_MIN_LOG"); return e ? atoi(e) : 13; }();
  return n >= ((uint32_t)1 << next_min_log) ? 4 : 0;
}

template <class CF>
int msm_finish(mira_msm_ctx* ctx, MsmPlan* plan, cudaStream_t st, PhaseTimer* pt) {
  int rc;
  const uint32_t B = plan->B;
  const unsigned S = (unsigned)plan->n_sets;
  // S = sum_b b * bucket[b].  For large bucket sets (>= 2^20) level passes (k_reduce_level) shrink the array 16x each
  // with two running sums per chunk and four doublings; the rest goes through k_reduce_chunks, whose per-thread
  // weighting by a small scalar multiplication is affordable there; every pass leaves partial sums in one array that
  // a tree sum (k_sum_points) folds.  Sizes first:
  uint32_t level_n[8], level_chunks[8];
  int level_log_m[8], n_levels = 0;
  uint32_t n = B;
  size_t a_total = 0;
  while (n_levels < 8) {
    const int lm = reduce_level_log_m(n, n_levels, S);
    if (!lm) break;
    const uint32_t chunks = (uint32_t)(((uint64_t)n + (1u << lm) - 1) >> lm);
    level_n[n_levels] = n; level_chunks[n_levels] = chunks; level_log_m[n_levels] = lm;
    n_levels++;
    a_total += chunks;
    n = chunks - 1;
  }
  // Small bucket sets (no level pass, at most 2^18 buckets in all): the reduction is a chain of dependent additions on
  // an idle GPU, so it runs on the cooperative kernels of coop.cuh (four warps per addition, ~3x lower latency per
  // group operation): one launch of at most 256 blocks — every lane owns m buckets, m chosen so that one wave covers
  // them — and one launch that sums the blocks' results.  MIRA_RED_COOP=0 keeps the round-1 kernels.
  static const int coop_on = [] { const char* e = getenv("MIRA_RED_COOP"); return e ? atoi(e) : 1; }();
  static const int coop_max_log = [] { const char* e = getenv("MIRA_RED_COOP_MAX_LOG"); return e ? atoi(e) : 18; }();
  if (coop_on && n_levels == 0 && B >= 2 && (uint64_t)B * S <= ((uint64_t)1 << coop_max_log)) {
    uint32_t mc = 2;
    while ((uint64_t)mc * 8192 < (uint64_t)B * S) mc <<= 1;
    const uint32_t n_lanes = (B + mc - 1) / mc, blocks = (n_lanes + 31) / 32;
    const size_t ca_stride = (size_t)(blocks + 1) * 128;
    if ((rc = ctx->red_a.ensure(ca_stride * S)) || (rc = ctx->red_b.ensure((size_t)256 * S)) || (rc = ctx->result.ensure((size_t)S * 192 + 256)))
      return rc;
    k_reduce_coop<CF><<<dim3(blocks, S), COOP_THREADS, 0, st>>>(ctx->buckets.p, B, mc, ctx->red_a.p, ((size_t)B + 1) * 128, ca_stride);
    plan->launches++;
    const void* fin = ctx->red_a.p;
    size_t fin_stride = ca_stride;
    if (blocks > 1) {
      k_sum_coop<CF><<<dim3(1, S), COOP_THREADS, 0, st>>>(ctx->red_a.p, blocks, (blocks + 31) / 32, ctx->red_b.p, ca_stride, 256);
      plan->launches++;
      fin = ctx->red_b.p;
      fin_stride = 256;
    }
    CU(cudaMemcpy2DAsync(ctx->result.p, 128, fin, fin_stride, 128, S, cudaMemcpyDeviceToDevice, st));
    if (pt) pt->mark(4);
    CU(cudaGetLastError());
    ctx->stats.window_bits = plan->c;
    ctx->stats.windows = plan->W;
    ctx->stats.entries = plan->entries;
    ctx->stats.buckets = B;
    ctx->stats.kernel_launches = plan->launches;
    return MIRA_OK;
  }
  // closing pass over the remaining n elements: buckets per thread m: the running sums are a serial chain of 2m full
  // adds per thread, so small sets get a small m (more, shorter chains)
  uint32_t m = n >> 15;
  m = m < 2 ? 2 : (m > 32 ? 32 : m);
  if (n < 2) m = 1;
  const uint32_t n_red = n ? (n + m - 1) / m : 0;
  a_total += n_red;
  const size_t a_stride = (a_total + 1) * 128, b_stride = (size_t)(a_total / 128 + 2) * 128;
  const size_t y_stride = ((size_t)(n_levels ? level_chunks[0] : 1) + 1) * 128;
  if ((rc = ctx->red_a.ensure(a_stride * S)) || (rc = ctx->red_b.ensure(b_stride * S)) || (rc = ctx->result.ensure((size_t)S * 192 + 256)))
    return rc;
  if (n_levels && ((rc = ctx->red_c.ensure(y_stride * S)) || (rc = ctx->red_d.ensure(y_stride * S)))) return rc;
  // bucket b of set s lives at buckets + s * in_stride + b * 128; element i of the level-0 array is bucket i + 1
  const char* src_lvl = reinterpret_cast<const char*>(ctx->buckets.p) + 128;
  size_t src_stride = ((size_t)B + 1) * 128;
  size_t a_off = 0;
  for (int l = 0; l < n_levels; l++) {
    void* next = (l & 1) ? ctx->red_d.p : ctx->red_c.p;
    k_reduce_level<CF><<<dim3((level_chunks[l] + 127) / 128, S), 128, 0, st>>>(src_lvl, level_n[l], level_log_m[l],
                                                                               reinterpret_cast<char*>(ctx->red_a.p) + a_off * 128, next,
                                                                               src_stride, a_stride, y_stride);
    plan->launches++;
    a_off += level_chunks[l];
    src_lvl = reinterpret_cast<const char*>(next);
    src_stride = y_stride;
  }
  if (n_red) {
    // k_reduce_chunks indexes buckets 1..n: element i of the current array is its "bucket" i + 1
    k_reduce_chunks<CF><<<dim3((n_red + 127) / 128, S), 128, 0, st>>>(src_lvl - 128, n, m, reinterpret_cast<char*>(ctx->red_a.p) + a_off * 128,
                                                                      src_stride, a_stride);
    plan->launches++;
  }
  void* src = ctx->red_a.p;
  void* dst = ctx->red_b.p;
  size_t dst_stride = b_stride;
  src_stride = a_stride;
  uint32_t cnt = (uint32_t)a_total;
  if (cnt == 0) {                 // B == 1 and nothing to weight beyond bucket 1 itself cannot happen (B >= 2); be safe
    CU(cudaMemsetAsync(ctx->red_a.p, 0, a_stride * S, st));
    cnt = 1;
  }
  while (cnt > 1) {
    // few points: one per thread (the serial part of a thread is as slow as a tree level, and the GPU is idle anyway)
    uint32_t per_thread = cnt > ((uint32_t)1 << 16) ? 8 : 1;
    uint32_t per_block = per_thread * 128;
    uint32_t blocks = (cnt + per_block - 1) / per_block;
    k_sum_points<CF><<<dim3(blocks, S), 128, 0, st>>>(src, cnt, per_thread, dst, src_stride, dst_stride);
    plan->launches++;
    cnt = blocks;
    std::swap(src, dst);
    std::swap(src_stride, dst_stride);
  }
  // result.p: S x 128 B XYZZ sums, followed (at S * 128) by S x 64 B for the affine results
  CU(cudaMemcpy2DAsync(ctx->result.p, 128, src, src_stride, 128, S, cudaMemcpyDeviceToDevice, st));
  if (pt) pt->mark(4);
  CU(cudaGetLastError());
  ctx->stats.window_bits = plan->c;
  ctx->stats.windows = plan->W;
  ctx->stats.entries = plan->entries;
  ctx->stats.buckets = B;
  ctx->stats.kernel_launches = plan->launches;
  return MIRA_OK;
}

// Scalars resident in HBM: one slice.
template <class CF, class SF>
int msm_device(mira_msm_ctx* ctx, const void* d_scalars, size_t n, cudaStream_t st) {
  int rc;
  if ((rc = ctx->result.ensure(256))) return rc;
  if (n == 0) {
    CU(cudaMemsetAsync(ctx->result.p, 0, 128, st));
    ctx->stats = mira_msm_stats{};
    return MIRA_OK;
  }
  MsmPlan plan;
  int window = 0;
  if ((rc = pick_window<SF>(ctx, d_scalars, n, true, st, &window))) return rc;
  PhaseTimer pt(ctx->profiling, st);
  // profiled commits and commits with affine levels run their phases one after the other (the phase times are what
  // the profile is for); everything else of sufficient length is software-pipelined over equal slices
  int K = (ctx->profiling || affine_levels_for(ctx) > 0) ? 1 : pipe_slices_for(ctx, n);
  size_t bounds[17];
  if (K > 1) {              // equal parts on 256-scalar boundaries; parts that round to nothing are dropped
    int parts = 0;
    bounds[0] = 0;
    for (int k = 1; k <= K; k++) {
      size_t b = k == K ? n : ((n / (size_t)K * (size_t)k) & ~(size_t)255);
      if (b > bounds[parts]) bounds[++parts] = b;
    }
    K = parts;
  }
  if (K > 1) {
    size_t max_slice = 0;
    for (int k = 0; k < K; k++) max_slice = std::max(max_slice, bounds[k + 1] - bounds[k]);
    if ((rc = msm_begin<CF>(ctx, n, max_slice, st, &plan, 1, window, 2))) return rc;
    if ((rc = msm_pipeline<CF, SF>(ctx, &plan, d_scalars, bounds, K, nullptr, st))) return rc;
    return msm_finish<CF>(ctx, &plan, st, nullptr);
  }
  if ((rc = msm_begin<CF>(ctx, n, n, st, &plan, 1, window))) return rc;
  const void* sets[1] = {d_scalars};
  if ((rc = msm_slice<CF, SF>(ctx, &plan, sets, 0, n, false, st, &pt))) return rc;
  if ((rc = msm_finish<CF>(ctx, &plan, st, &pt))) return rc;
  if (ctx->profiling) {
    CU(cudaStreamSynchronize(st));
    ctx->stats.ms_digits = pt.ms(0, 1);
    ctx->stats.ms_sort = pt.ms(1, 2);
    ctx->stats.ms_accumulate = pt.ms(2, 3);
    ctx->stats.ms_reduce = pt.ms(3, 4);
    ctx->stats.ms_total = pt.ms(0, 4);
  }
  return MIRA_OK;
}

// Scalars in HOST memory: the vector is cut into slices; slice k+1 crosses PCIe on the copy stream while slice k
// runs digits -> sort -> accumulate, so only the first slice's copy is exposed.  Page-locked sources are copied by
// the copy engine directly, pageable ones are staged through page-locked slots by worker threads (stager.hpp).
constexpr int SLICE_MAX_COUNT = 4;                 // ctx->slice_min: smallest (first) slice, in scalars

template <class CF, class SF>
int msm_host(mira_msm_ctx* ctx, const void* h_scalars, size_t n, cudaStream_t st) {
  int rc;
  if ((rc = ctx->result.ensure(256))) return rc;
  if (n == 0) {
    CU(cudaMemsetAsync(ctx->result.p, 0, 128, st));
    ctx->stats = mira_msm_stats{};
    return MIRA_OK;
  }
  ctx->scalars_valid = 0;
  if ((rc = ctx->scalars.ensure(n * 32))) return rc;
  // sampled BEFORE the slice copies are queued (its 64 small copies must not wait behind them); also tells how dense
  // the vector is, i.e. how much accumulation time a copied scalar buys
  int window = 0;
  double pairs_per_scalar = 0.0;
  if ((rc = pick_window<SF>(ctx, h_scalars, n, false, st, &window, &pairs_per_scalar))) return rc;
  // Slice sizes grow geometrically: slice k+1 crosses PCIe while slice k is accumulated, so it may be as much larger
  // as accumulating a scalar takes longer than copying it.  Uniform scalars (~12 pairs each at 2^24): ~4x, hence
  // 1 : 4 : 16 — the exposed first copy is then as small as ctx->slice_min allows (2^24 scalars: 0.8 M first =>
  // ~0.5 ms of the 9.7 ms H2D stays visible).  Witness columns are sparse (~1 pair per scalar): their accumulation is
  // no longer than their copy, a 4x slice would wait for its own copy, and equal slices are right (the commit is then
  // copy-bound and ends one slice's accumulation after the last copy).
  size_t growth = 4;
  if (pairs_per_scalar > 0.0) {
    const double ratio = pairs_per_scalar * 0.2 / 0.58;      // ~0.2 ns per pair (digits, sort, accumulate) vs 32 B at ~55 GB/s
    growth = ratio >= 3.0 ? 4 : (ratio >= 1.5 ? 2 : 1);
  }
  int K = 1;
  {
    size_t weight = 1, sum = 1;
    while (K < SLICE_MAX_COUNT && ctx->slice_min && n / (sum + weight * growth) >= ctx->slice_min) {
      weight *= growth;
      sum += weight;
      K++;
    }
  }
  size_t bounds[SLICE_MAX_COUNT + 2];
  int n_slices = 0;
  bounds[0] = 0;
  {
    size_t total_w = 0, w = 1;
    for (int k = 0; k < K; k++, w *= growth) total_w += w;
    size_t pos = 0;
    w = 1;
    for (int k = 0; k < K - 1; k++, w *= growth) {
      size_t len = ((n * w / total_w) + 255) & ~(size_t)255;
      if (pos + len >= n) break;
      pos += len;
      bounds[++n_slices] = pos;
    }
    bounds[++n_slices] = n;
  }
  size_t max_slice = 0;
  for (int k = 0; k < n_slices; k++) max_slice = std::max(max_slice, bounds[k + 1] - bounds[k]);
  if (!ctx->copy_stream) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < SLICE_MAX_COUNT; k++)
    if (!ctx->copy_done[k]) CU(cudaEventCreateWithFlags(&ctx->copy_done[k], cudaEventDisableTiming));
  if (!ctx->compute_idle) CU(cudaEventCreateWithFlags(&ctx->compute_idle, cudaEventDisableTiming));
  // the copies must not overtake work already queued on `st` that still reads ctx->scalars (a previous commit)
  CU(cudaEventRecord(ctx->compute_idle, st));
  CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->compute_idle, 0));
  // page-locked source: all copies are queued now and run by the copy engine on their own;
  // pageable source (a plain Rust Vec): each slice is staged through page-locked slots by worker threads right
  // before its compute is queued, so the staging of slice k+1 overlaps the accumulation of slice k
  cudaPointerAttributes attr{};
  bool pageable = true;
  if (cudaPointerGetAttributes(&attr, h_scalars) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
  else cudaGetLastError();
  if (pageable) CU(ctx->stager.init());
  if (!pageable)
    for (int k = 0; k < n_slices; k++) {
      size_t first = bounds[k], cnt = bounds[k + 1] - bounds[k];
      CU(cudaMemcpyAsync((char*)ctx->scalars.p + first * 32, (const char*)h_scalars + first * 32, cnt * 32, cudaMemcpyDefault,
                         ctx->copy_stream));
      CU(cudaEventRecord(ctx->copy_done[k], ctx->copy_stream));
    }
  MsmPlan plan;
  const bool overlap = n_slices > 1 && affine_levels_for(ctx) == 0 && ctx->pipe_slices != 1;      // set_pipeline(1) turns it off
  if ((rc = msm_begin<CF>(ctx, n, max_slice, st, &plan, 1, window, overlap ? 2 : 1))) return rc;
  if (overlap && !pageable) {
    // every copy is already queued: slice k+1 is decomposed and sorted behind its copy while slice k is accumulated
    if ((rc = msm_pipeline<CF, SF>(ctx, &plan, ctx->scalars.p, bounds, n_slices, ctx->copy_done, st))) return rc;
    ctx->scalars_valid = n;
    return msm_finish<CF>(ctx, &plan, st, nullptr);
  }
  for (int k = 0; k < n_slices; k++) {
    size_t first = bounds[k], cnt = bounds[k + 1] - bounds[k];
    if (pageable) {
      CU(ctx->stager.copy((char*)ctx->scalars.p + first * 32, (const char*)h_scalars + first * 32, cnt * 32, ctx->copy_stream));
      CU(cudaEventRecord(ctx->copy_done[k], ctx->copy_stream));
    }
    CU(cudaStreamWaitEvent(st, ctx->copy_done[k], 0));
    const void* sets[1] = {(const char*)ctx->scalars.p + first * 32};
    if ((rc = msm_slice<CF, SF>(ctx, &plan, sets, first, cnt, k > 0, st, nullptr))) return rc;
  }
  ctx->scalars_valid = n;
  return msm_finish<CF>(ctx, &plan, st, nullptr);
}

template <class CF, class SF>
int commit_impl(mira_msm_ctx* ctx, const void* scalars, size_t n, int on_device, void* out, bool want_affine, cudaStream_t st) {
  int rc;
  if (on_device) rc = msm_device<CF, SF>(ctx, scalars, n, st);
  else rc = msm_host<CF, SF>(ctx, scalars, n, st);
  if (rc) return rc;
  if (want_affine) {
    k_finalize<CF><<<1, 32, 0, st>>>(ctx->result.p, (char*)ctx->result.p + 128);
    ctx->stats.kernel_launches++;
    CU(cudaMemcpyAsync(ctx->h_result, (char*)ctx->result.p + 128, 64, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(out, ctx->h_result, 64);
  } else {
    CU(cudaMemcpyAsync(ctx->h_result, ctx->result.p, 128, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(out, ctx->h_result, 128);
  }
  return MIRA_OK;
}


// Several vectors of the same length against the same key (the 5-6 cross-term commitments of one fold,
// src/nifs/vanilla/mod.rs:124-127): one digit kernel per vector, then ONE sort, accumulation and reduction over
// `count` bucket sets, one normalisation launch, one D2H of count x 64 B.
template <class CF, class SF>
int commit_batch_impl(mira_msm_ctx* ctx, const void* const* d_scalar_sets, size_t count, size_t n, void* out, cudaStream_t st) {
  int rc;
  if ((rc = ctx->result.ensure(count * 192 + 256))) return rc;
  if (n == 0) {
    memset(out, 0, count * 64);
    ctx->stats = mira_msm_stats{};
    return MIRA_OK;
  }
  MsmPlan plan;
  if ((rc = msm_begin<CF>(ctx, n, n, st, &plan, (int)count))) return rc;
  if ((rc = msm_slice<CF, SF>(ctx, &plan, d_scalar_sets, 0, n, false, st, nullptr))) return rc;
  if ((rc = msm_finish<CF>(ctx, &plan, st, nullptr))) return rc;
  char* aff = (char*)ctx->result.p + count * 128;
  k_finalize<CF><<<(unsigned)count, 32, 0, st>>>(ctx->result.p, aff);
  ctx->stats.kernel_launches++;
  CU(cudaMemcpyAsync(ctx->h_result, aff, count * 64, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  memcpy(out, ctx->h_result, count * 64);
  return MIRA_OK;
}

// Row-sharded provers: `count` vectors against this rank's key shard, the un-normalised XYZZ sums
// (count x 128 B) written to DEVICE memory on `st` without a host round trip, so that a rank can queue all the
// commitments of a fold step, all_gather them once and combine.  One vector goes through the adaptive-window path
// (witness columns are sparse), several through the batched one.
template <class CF, class SF>
int partial_batch_dev_impl(mira_msm_ctx* ctx, const void* const* d_scalar_sets, size_t count, size_t n, void* d_out, cudaStream_t st) {
  int rc;
  if ((rc = ctx->result.ensure(count * 192 + 256))) return rc;
  if (n == 0) {
    CU(cudaMemsetAsync(d_out, 0, count * 128, st));
    ctx->stats = mira_msm_stats{};
    return MIRA_OK;
  }
  if (count == 1) {
    if ((rc = msm_device<CF, SF>(ctx, d_scalar_sets[0], n, st))) return rc;
  } else {
    MsmPlan plan;
    if ((rc = msm_begin<CF>(ctx, n, n, st, &plan, (int)count))) return rc;
    if ((rc = msm_slice<CF, SF>(ctx, &plan, d_scalar_sets, 0, n, false, st, nullptr))) return rc;
    if ((rc = msm_finish<CF>(ctx, &plan, st, nullptr))) return rc;
  }
  CU(cudaMemcpyAsync(d_out, ctx->result.p, count * 128, cudaMemcpyDeviceToDevice, st));
  return MIRA_OK;
}

// out[j] = to_affine( sum over ranks g of partials[g * rank_stride + j * 128] ), j < n_commits; partials on the device
// (as an all_gather leaves them), results to the host.  One warp per commitment: lane g sums ranks g, g + 32, ..., a
// shuffle tree folds the lanes (log2 instead of n_ranks dependent additions: the call sits on the critical path of
// every multi-rank step), lane 0 normalises.
template <class CF>
__device__ __forceinline__ Xyzz<CF> xyzz_shfl_down(const Xyzz<CF>& p, int delta) {
  Xyzz<CF> r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], delta);
    r.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], delta);
    r.zz.v[i] = __shfl_down_sync(0xffffffffu, p.zz.v[i], delta);
    r.zzz.v[i] = __shfl_down_sync(0xffffffffu, p.zzz.v[i], delta);
  }
  return r;
}
template <class CF>
__global__ void __launch_bounds__(32) k_combine_partials(const void* __restrict__ partials, uint32_t n_ranks, size_t rank_stride,
                                                         void* __restrict__ out_affine) {
  const uint32_t lane = threadIdx.x;
  Xyzz<CF> acc = xyzz_identity<CF>();
  for (uint32_t g = lane; g < n_ranks; g += 32) {
    Xyzz<CF> p = xyzz_load<CF>(reinterpret_cast<const char*>(partials) + (size_t)g * rank_stride + (size_t)blockIdx.x * 128);
    xyzz_add(acc, p);
  }
  uint32_t top = 1;
  while (top < n_ranks && top < 32) top <<= 1;
  for (uint32_t d = top >> 1; d > 0; d >>= 1) {
    Xyzz<CF> o = xyzz_shfl_down(acc, (int)d);
    if (lane < d) xyzz_add(acc, o);
  }
  if (lane == 0) aff_store<CF>(reinterpret_cast<char*>(out_affine) + (size_t)blockIdx.x * 64, xyzz_to_affine(acc));
}

// scratch of combine_dev_impl, one per device: a device buffer for the affine results and a page-locked mirror
struct CombineScratch {
  void* d = nullptr;
  void* h = nullptr;
  size_t cap = 0;
};
inline CombineScratch* combine_scratch(int device, size_t bytes, std::mutex** mu_out) {
  static CombineScratch scratch[64];
  static std::mutex mu[64];
  CombineScratch& s = scratch[device & 63];
  *mu_out = &mu[device & 63];
  std::lock_guard<std::mutex> lk(**mu_out);
  if (bytes > s.cap) {
    if (s.d) cudaFree(s.d);
    if (s.h) cudaFreeHost(s.h);
    s.d = s.h = nullptr;
    s.cap = 0;
    size_t cap = bytes < 4096 ? 4096 : bytes;
    if (cudaMalloc(&s.d, cap) != cudaSuccess || cudaMallocHost(&s.h, cap) != cudaSuccess) {
      cudaGetLastError();
      if (s.d) cudaFree(s.d);
      s.d = nullptr;
      return nullptr;
    }
    s.cap = cap;
  }
  return &s;
}

template <class CF>
int combine_dev_impl(const void* d_partials, size_t n_ranks, size_t n_commits, size_t rank_stride, void* out_affine_host, cudaStream_t st) {
  int device = 0;
  CU(cudaGetDevice(&device));
  std::mutex* mu = nullptr;
  CombineScratch* s = combine_scratch(device, n_commits * 64, &mu);
  if (!s) return fail(MIRA_ERR_CUDA, "combine: no scratch memory for %zu commitments", n_commits);
  std::lock_guard<std::mutex> lk(*mu);          // one combine at a time per device (the scratch is shared)
  k_combine_partials<CF><<<(unsigned)n_commits, 32, 0, st>>>(d_partials, (uint32_t)n_ranks, rank_stride, s->d);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(s->h, s->d, n_commits * 64, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  memcpy(out_affine_host, s->h, n_commits * 64);
  return MIRA_OK;
}

// One shard of a single-process multi-GPU commit (mira_msm_ctx_create_sharded): this device's slice of the HOST scalar
// vector through the ordinary host-buffer pipeline, then the 128-byte XYZZ partial sum straight from this device's
// memory into the gather buffer on the combining device (peer copy: NVLink when peer access is enabled, staged by the
// driver otherwise).  Returns when the partial has landed.
template <class CF, class SF>
int partial_to_peer_impl(mira_msm_ctx* ctx, const void* h_scalars, size_t n, void* d_dst, int dst_device, cudaStream_t st) {
  int rc;
  if ((rc = msm_host<CF, SF>(ctx, h_scalars, n, st))) return rc;
  if (dst_device == ctx->device) CU(cudaMemcpyAsync(d_dst, ctx->result.p, 128, cudaMemcpyDeviceToDevice, st));
  else CU(cudaMemcpyPeerAsync(d_dst, dst_device, ctx->result.p, ctx->device, 128, st));
  CU(cudaStreamSynchronize(st));
  return MIRA_OK;
}

template <class CF, class SF>
int prepare_impl(mira_msm_ctx* ctx, size_t n, const void* like_scalars, int on_device) {
  int rc, window = 0;
  if (like_scalars && (rc = pick_window<SF>(ctx, like_scalars, n, on_device != 0, ctx->stream, &window))) return rc;
  int c = ctx->forced_window ? ctx->forced_window : (window ? window : choose_window(n));
  Table* t = nullptr;
  if ((rc = get_table<CF>(ctx, c, n, ctx->stream, &t))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MIRA_OK;
}

template <class CF>
int check_on_curve_impl(mira_msm_ctx* ctx, uint32_t b_small, int b_negative) {
  int rc;
  if ((rc = ctx->result.ensure(256))) return rc;
  uint32_t* flag = (uint32_t*)((char*)ctx->result.p + 192);
  CU(cudaMemsetAsync(flag, 0, 4, ctx->stream));
  if (ctx->n_bases) {
    unsigned blocks = (unsigned)((ctx->n_bases + 255) / 256);
    k_check_on_curve<CF><<<blocks, 256, 0, ctx->stream>>>(ctx->d_bases, (uint32_t)ctx->n_bases, b_small, b_negative, flag);
  }
  uint32_t h = 0;
  CU(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (h) return fail(MIRA_ERR_NOT_ON_CURVE, "Wrong file in cache, some ptr out of curve");
  return MIRA_OK;
}

template <class CF>
int combine_impl(const void* partials, size_t count, void* out_affine) {
  void* d = nullptr;
  size_t bytes = (count + 2) * 128;
  CU(cudaMalloc(&d, bytes * 2));
  cudaError_t e = cudaMemset(d, 0, bytes * 2);
  if (e == cudaSuccess && count) e = cudaMemcpy(d, partials, count * 128, cudaMemcpyHostToDevice);
  void* src = d;
  void* dst = (char*)d + bytes;
  uint32_t cnt = count ? (uint32_t)count : 1;
  while (e == cudaSuccess && cnt > 1) {
    uint32_t blocks = (cnt + 127) / 128;
    k_sum_points<CF><<<blocks, 128>>>(src, cnt, 1, dst);
    cnt = blocks;
    std::swap(src, dst);
  }
  if (e == cudaSuccess) {
    k_finalize<CF><<<1, 32>>>(src, (char*)dst);
    e = cudaMemcpy(out_affine, dst, 64, cudaMemcpyDeviceToHost);
  }
  cudaFree(d);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "combine failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

template <class SF>
int gen_scalars_impl(uint64_t seed, size_t first, size_t n, int dist, void* out_dev) {
  k_gen_scalars<SF><<<(unsigned)((n + 255) / 256), 256>>>(seed, first, n, dist, out_dev);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return MIRA_OK;
}

template <class CF, class SF>
int gen_bases_impl(uint64_t seed, size_t first, size_t n, void* out_dev) {
  void* table = nullptr;
  CU(cudaMalloc(&table, 32 * 256 * 64));
  k_gen_table<CF><<<32 * 256 / 128, 128>>>(table);
  k_gen_bases<CF, SF><<<(unsigned)((n + 127) / 128), 128>>>(seed, first, n, table, out_dev);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(table);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "gen_bases failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

template <class CF>
int test_point_op_impl(int op, const void* p, const void* q, size_t n, void* out) {
  k_test_point<CF><<<(unsigned)((n + 127) / 128), 128>>>(op, p, q, n, out);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return MIRA_OK;
}

}  // namespace mira_host
