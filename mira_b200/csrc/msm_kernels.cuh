// msm_kernels.cuh — device kernels of the fixed-base signed-window Pippenger MSM.
//
// Replaces halo2_proofs::arithmetic::best_multiexp as called from CommitmentKey::commit
// (/root/reference/src/commitment.rs:78-87).  Pipeline:
//   digits        scalar: Montgomery -> canonical, signed c-bit digits, (bucket, point-ref) pairs.  k_digit_hist +
//                 k_digits_scatter place the pairs straight into 256 bins (the sort's first pass, fused); sparse
//                 vectors keep k_digits' compacted list
//   grouping      pairs of one bucket made contiguous: the MSD partition k_msd_* below (from 2^25 pairs), else the
//                 stable LSD passes of sort.cu
//   k_accumulate  load-balanced segmented XYZZ mixed-add over the grouped list (fixed-size chunks)
//   k_combine     stitch runs that straddle chunk boundaries
//   k_reduce_*    S = sum_b b * bucket[b] by chunked running sums + tree sum
//   k_finalize    to_affine
// The key is static, so windows j > 0 use precomputed points 2^(c*j) * P_i (k_precompute) and ALL
// windows share ONE bucket set: no per-window doublings, one reduction.
#pragma once
#include "curve.cuh"

namespace mira {

constexpr uint32_t REF_NEG = 0x80000000u;       // sign flag in a point reference
constexpr uint32_t PK_OPEN_LEFT = 0x80000000u;  // partial continues a run from the previous chunk
constexpr uint32_t PK_OPEN_RIGHT = 0x40000000u; // partial's run continues in the next chunk
constexpr uint32_t PK_KEY_MASK = 0x3fffffffu;

// ------------------------------------------------------------------ digits
// One thread per scalar: Montgomery -> canonical, then W signed c-bit digits d_j in [-2^(c-1)+1, 2^(c-1)].
// Every non-zero digit becomes one (key, ref) pair: key = |d_j| (the bucket), ref = j*n_cover + i with the
// sign in bit 31 (j*n_cover + i indexes the fixed-base table entry 2^(c*j) * P_i).  Zero digits are dropped
// here (witness-like scalars are mostly zero), so the pair list is compacted: a block counts its pairs,
// reserves a range with ONE atomicAdd, and each warp writes window by window at ballot-derived positions,
// i.e. 128-byte coalesced rows.  *n_out receives the total number of pairs.
constexpr int DG_THREADS = 256;
constexpr int DG_WARPS = DG_THREADS / 32;

// digit j of the canonical scalar whose 8 limbs (+ a zero 9th) sit in shared memory at `sc` (stride-9 rows are
// bank-conflict free): two LDS and a funnel shift instead of a register-select ladder
__device__ __forceinline__ uint32_t signed_digit(const uint32_t* sc, int j, int c, uint32_t& carry, uint32_t& neg) {
  const uint32_t half = 1u << (c - 1);
  const uint32_t mask = (1u << c) - 1u;     // c <= 26
  int bit = j * c;
  int w = bit >> 5, sh = bit & 31;
  uint32_t lo = w < 8 ? sc[w] : 0u, hi = w < 7 ? sc[w + 1] : 0u;
  uint32_t d = __funnelshift_r(lo, hi, sh) & mask;
  d += carry;
  neg = 0;
  carry = 0;
  if (d > half) {                            // d in (half, 2^c] -> d - 2^c in (-half, 0]
    d = (1u << c) - d;
    neg = d ? REF_NEG : 0u;
    carry = 1;
  }
  return d;
}

template <class SF>
__global__ void __launch_bounds__(DG_THREADS) k_digits(const void* __restrict__ scalars, uint32_t n, uint32_t first, int c, int W,
                                                       uint32_t n_cover, uint32_t key_offset, uint32_t* __restrict__ keys,
                                                       uint32_t* __restrict__ refs, uint32_t* __restrict__ n_out) {
  // key_offset: batched commits (several scalar vectors against the same key) give every vector its own bucket set;
  // the pairs of all vectors then go through ONE sort and ONE accumulation
  extern __shared__ uint32_t sm_cnt[];       // [W][DG_WARPS] pair counts -> exclusive offsets
  __shared__ uint32_t sm_sc[DG_THREADS * 9];
  __shared__ uint32_t sm_base;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool valid = i < n;
  Fe<SF> s = fe_zero<SF>();
  if (valid) s = fe_to_canonical(fe_load<SF>(reinterpret_cast<const char*>(scalars) + (size_t)i * 32));
  uint32_t* sc = sm_sc + threadIdx.x * 9;
#pragma unroll
  for (int k = 0; k < 8; k++) sc[k] = s.v[k];
  // pass 1: count the non-zero digits per (window, warp)
  uint32_t carry = 0, neg;
  for (int j = 0; j < W; j++) {
    uint32_t d = signed_digit(sc, j, c, carry, neg);
    uint32_t b = __ballot_sync(0xffffffffu, d != 0);
    if (lane == 0) sm_cnt[j * DG_WARPS + warp] = __popc(b);
  }
  __syncthreads();
  if (warp == 0) {                           // exclusive scan of the W * DG_WARPS counts by one warp
    const int total = W * DG_WARPS, per = (total + 31) / 32;
    uint32_t sum = 0;
    for (int k = 0; k < per; k++) {
      int idx = lane * per + k;
      if (idx < total) sum += sm_cnt[idx];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    uint32_t run = incl - sum;
    for (int k = 0; k < per; k++) {
      int idx = lane * per + k;
      if (idx < total) {
        uint32_t t = sm_cnt[idx];
        sm_cnt[idx] = run;
        run += t;
      }
    }
    uint32_t block_total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) sm_base = block_total ? atomicAdd(n_out, block_total) : 0u;
  }
  __syncthreads();
  const uint32_t base = sm_base;
  const uint32_t lt = (1u << lane) - 1u;
  carry = 0;
  for (int j = 0; j < W; j++) {
    uint32_t d = signed_digit(sc, j, c, carry, neg);
    uint32_t b = __ballot_sync(0xffffffffu, d != 0);
    if (d) {
      uint32_t pos = base + sm_cnt[j * DG_WARPS + warp] + __popc(b & lt);
      keys[pos] = key_offset + d;
      refs[pos] = ((uint32_t)j * n_cover + first + i) | neg;     // `scalars` is the slice starting at key index `first`
    }
  }
}

// ---- digits fused with the FIRST radix pass (round 2) ----------------------------------------------------------
// The LSD sort's first pass needs no stability (the order the digit kernel emits pairs in carries no meaning), so the
// digit kernel can place every pair straight into the 256 bins of the key's low byte instead of writing a compacted
// list that the first pass re-reads, histograms and scatters.  Three launches:
//   k_digit_hist     the 256-bin histogram of all pairs' low key bytes (scalars re-read in the scatter: 32 B per
//                    scalar, against the 8 B per PAIR, read and written, of the pass it replaces);
//   k_digit_scan     exclusive scan of the 256 counts -> per-bin cursors; the total is the pair count (*n_out);
//   k_digits_scatter a block decomposes `spb` scalars, ranks its pairs by bin with shared-memory atomics, reserves one
//                    range per bin with a global atomicAdd on that bin's cursor and streams the pairs out bin by bin
//                    (runs of ~spb * W / 256 pairs).
// At 2^24 points: digits 0.71 + first pass 1.64 ms -> see DESIGN.md section 6.  Block-to-block order inside a bin
// depends on the atomics' arrival order; the commitment does not (group addition is commutative and the result is
// normalised), and the later passes are stable.
// Block geometry, measured at 2^24 points (digit phase, ms): 512 scalars / 6144 pairs 1.375 | 384 / 4608 1.370 |
// 256 / 3072 1.329 (profiles/r02_fused_digits.txt)
#ifndef MIRA_DS_CAP
#define MIRA_DS_CAP 3072
#endif
#ifndef MIRA_DS_SPB_MAX
#define MIRA_DS_SPB_MAX 256
#endif
constexpr int DS_THREADS = 256;
constexpr int DS_CAP = MIRA_DS_CAP;        // pairs staged per block (24 KiB of keys + refs)
constexpr int DS_SPB_MAX = MIRA_DS_SPB_MAX;   // scalars per block

__host__ __device__ inline uint32_t ds_scalars_per_block(int W) {
  uint32_t spb = (uint32_t)DS_CAP / (uint32_t)W;
  spb &= ~31u;
  return spb > (uint32_t)DS_SPB_MAX ? (uint32_t)DS_SPB_MAX : (spb < 32u ? 32u : spb);
}

template <class SF>
__global__ void __launch_bounds__(DS_THREADS) k_digit_hist(const void* __restrict__ scalars, uint32_t n, int c, int W, uint32_t key_offset,
                                                           int bin_shift, uint32_t bin_bias, uint32_t* __restrict__ g_hist) {
  __shared__ uint32_t sh[256];
  __shared__ uint32_t sm_sc[DS_THREADS * 9];
  sh[threadIdx.x] = 0;
  __syncthreads();
  uint32_t* sc = sm_sc + threadIdx.x * 9;
  for (uint32_t i = blockIdx.x * DS_THREADS + threadIdx.x; i < n; i += gridDim.x * DS_THREADS) {
    Fe<SF> s = fe_to_canonical(fe_load<SF>(reinterpret_cast<const char*>(scalars) + (size_t)i * 32));
#pragma unroll
    for (int k = 0; k < 8; k++) sc[k] = s.v[k];
    uint32_t carry = 0, neg;
    for (int j = 0; j < W; j++) {
      uint32_t d = signed_digit(sc, j, c, carry, neg);
      if (d) atomicAdd(&sh[((key_offset + d - bin_bias) >> bin_shift) & 0xffu], 1u);
    }
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&g_hist[threadIdx.x], sh[threadIdx.x]);
}

// exclusive scan of 256 counts by one block of 256 threads; returns the thread's exclusive prefix, `total` = sum
__device__ __forceinline__ uint32_t ds_scan256(uint32_t v, uint32_t* tmp8, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) tmp8[warp] = x;
  __syncthreads();
  uint32_t off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < DS_THREADS / 32; w++) {
    uint32_t t = tmp8[w];
    if (w < warp) off += t;
    tot += t;
  }
  total = tot;
  __syncthreads();
  return off + x - v;
}

static __global__ void __launch_bounds__(DS_THREADS) k_digit_scan(const uint32_t* __restrict__ g_hist, uint32_t* __restrict__ g_cursor,
                                                           uint32_t* __restrict__ n_out) {
  __shared__ uint32_t tmp[8];
  uint32_t total;
  uint32_t ex = ds_scan256(g_hist[threadIdx.x], tmp, total);
  g_cursor[threadIdx.x] = ex;
  if (threadIdx.x == 0) *n_out = total;
}

template <class SF>
__global__ void __launch_bounds__(DS_THREADS) k_digits_scatter(const void* __restrict__ scalars, uint32_t n, uint32_t first, int c, int W,
                                                               uint32_t n_cover, uint32_t key_offset, uint32_t spb, int bin_shift, uint32_t bin_bias,
                                                               uint32_t* __restrict__ g_cursor, uint32_t* __restrict__ out_keys,
                                                               uint32_t* __restrict__ out_refs) {
  // bin = ((key - bin_bias) >> bin_shift) & 255: shift 0, bias 0 for the LSD sort's first pass; the TOP byte of key - 1
  // (keys start at 1, so key - 1 fills its bit range evenly) for the MSD partition below
  extern __shared__ uint32_t ds_dyn[];
  uint32_t* sc_all = ds_dyn;                       // [spb][9] canonical limbs
  uint32_t* st_keys = ds_dyn + spb * 9;            // [spb * W] pairs grouped by bin
  uint32_t* st_refs = st_keys + spb * (uint32_t)W;
  __shared__ uint32_t cnt[256], dstart[256], gbase[256], tmp[8];
  const uint32_t base_i = blockIdx.x * spb;
  const uint32_t cnt_i = n - base_i < spb ? n - base_i : spb;
  cnt[threadIdx.x] = 0;
  for (uint32_t s = threadIdx.x; s < cnt_i; s += DS_THREADS) {
    Fe<SF> v = fe_to_canonical(fe_load<SF>(reinterpret_cast<const char*>(scalars) + (size_t)(base_i + s) * 32));
#pragma unroll
    for (int k = 0; k < 8; k++) sc_all[s * 9 + k] = v.v[k];
  }
  __syncthreads();
  for (uint32_t s = threadIdx.x; s < cnt_i; s += DS_THREADS) {
    uint32_t carry = 0, neg;
    for (int j = 0; j < W; j++) {
      uint32_t d = signed_digit(sc_all + s * 9, j, c, carry, neg);
      if (d) atomicAdd(&cnt[((key_offset + d - bin_bias) >> bin_shift) & 0xffu], 1u);
    }
  }
  __syncthreads();
  const uint32_t mine = cnt[threadIdx.x];
  uint32_t total;
  const uint32_t ex = ds_scan256(mine, tmp, total);
  dstart[threadIdx.x] = ex;
  gbase[threadIdx.x] = mine ? atomicAdd(&g_cursor[threadIdx.x], mine) : 0u;
  cnt[threadIdx.x] = ex;                           // from here on: the bin's next free staging slot
  __syncthreads();
  for (uint32_t s = threadIdx.x; s < cnt_i; s += DS_THREADS) {
    uint32_t carry = 0, neg;
    for (int j = 0; j < W; j++) {
      uint32_t d = signed_digit(sc_all + s * 9, j, c, carry, neg);
      if (d) {
        const uint32_t key = key_offset + d;
        const uint32_t q = atomicAdd(&cnt[((key - bin_bias) >> bin_shift) & 0xffu], 1u);
        st_keys[q] = key;
        st_refs[q] = ((uint32_t)j * n_cover + first + base_i + s) | neg;     // `scalars` is the slice starting at key index `first`
      }
    }
  }
  __syncthreads();
  for (uint32_t q = threadIdx.x; q < total; q += DS_THREADS) {
    const uint32_t key = st_keys[q];
    const uint32_t b = ((key - bin_bias) >> bin_shift) & 0xffu;
    const uint32_t dst = gbase[b] + (q - dstart[b]);
    out_keys[dst] = key;
    out_refs[dst] = st_refs[q];
  }
}

// ---- MSD partition: grouping by key without a stable pass (round 2) ----------------------------------------------
// k_accumulate needs equal keys to be contiguous (and ascending), nothing more, and a most-significant-digit-first
// partition needs no stability anywhere: once the pairs are grouped by the key's top bits, every group is partitioned
// on its own.  Everything below works on key - 1 (keys start at 1: bucket b of set s is s * (B + 1) + b, b >= 1), whose K
// bits are evenly filled; with K in [17, 24]:
//   k_digit_hist     (above, bin_shift = K - 8) histogram of the pairs' TOP key byte;
//   k_msd_scan1      segment offsets, cursors, the tile table of the 256 top-byte segments, *n_out;
//   k_digits_scatter (above, bin_shift = K - 8) the digit kernel scatters its pairs into the 256 segments;
//   k_msd_hist16     tiles of a segment: histogram of the next 8 bits -> 65,536 group counts; k_msd_scan16 scans them;
//   k_msd_mid        tiles of a segment: rank by the next 8 bits with shared-memory atomics, one global atomicAdd per
//                    (tile, sub-bin) on that group's cursor, the tile re-ordered in shared memory and written in runs;
//   k_msd_low        a block per run of groups: 2^(K-16)-bin counting sort of each group in shared memory (groups
//                    beyond its capacity, i.e. skewed vectors, through two passes over global memory).
// Against the LSD sort's two stable passes (135 instructions per pair each: warp-ballot ranking; a histogram and three
// scan kernels in front of each) the unstable passes rank with one shared-memory atomic per pair.
constexpr int MSD_GROUP_BITS = 16;
constexpr int MSD_GROUPS = 1 << MSD_GROUP_BITS;
#ifndef MIRA_MSD_TILE
#define MIRA_MSD_TILE 4096
#endif
// k_msd_mid geometry, sort phase at 2^24 points (ms): 4096 pairs x 512 threads 2.001 | 2048 x 256 2.135 | 4096 x 256 1.956
#ifndef MIRA_MSD_M_THREADS
#define MIRA_MSD_M_THREADS 256
#endif
constexpr int MSD_TILE = MIRA_MSD_TILE;    // pairs per k_msd_hist16 / k_msd_mid block
constexpr int MSD_M_THREADS = MIRA_MSD_M_THREADS;
constexpr int MSD_S_THREADS = 1024;        // k_msd_scan16

// hist1[256] -> offs1[257] (segment starts, offs1[256] = pair count), cursor1 (a copy k_digits_scatter advances),
// tile_tab[b] = number of MSD_TILE tiles in segments < b (tile_tab[256] = all tiles), *n_out.  One block of 256 threads.
static __global__ void __launch_bounds__(DS_THREADS) k_msd_scan1(const uint32_t* __restrict__ hist1, uint32_t* __restrict__ offs1,
                                                                 uint32_t* __restrict__ cursor1, uint32_t* __restrict__ tile_tab,
                                                                 uint32_t* __restrict__ n_out) {
  __shared__ uint32_t tmp[8];
  const uint32_t mine = hist1[threadIdx.x];
  uint32_t total, tiles_total;
  const uint32_t ex = ds_scan256(mine, tmp, total);
  offs1[threadIdx.x] = ex;
  cursor1[threadIdx.x] = ex;
  const uint32_t tex = ds_scan256((mine + MSD_TILE - 1) / MSD_TILE, tmp, tiles_total);
  tile_tab[threadIdx.x] = tex;
  if (threadIdx.x == 0) {
    offs1[256] = total;
    tile_tab[256] = tiles_total;
    *n_out = total;
  }
}

// which segment and which slice [lo, hi) of the pair list does tile `blockIdx.x` cover?  (thread 0 fills the three words)
__device__ __forceinline__ void msd_locate_tile(const uint32_t* __restrict__ offs1, const uint32_t* __restrict__ tile_tab, uint32_t* seg_lo_hi) {
  uint32_t lo = 0, hi = 255;
  while (lo < hi) {
    uint32_t mid = (lo + hi + 1) >> 1;
    if (tile_tab[mid] <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const uint32_t seg_hi = offs1[lo + 1];
  const uint32_t t_lo = offs1[lo] + (blockIdx.x - tile_tab[lo]) * MSD_TILE;
  seg_lo_hi[0] = lo;
  seg_lo_hi[1] = t_lo;
  seg_lo_hi[2] = t_lo + MSD_TILE < seg_hi ? t_lo + MSD_TILE : seg_hi;
}

static __global__ void __launch_bounds__(MSD_M_THREADS) k_msd_hist16(const uint32_t* __restrict__ in_keys, const uint32_t* __restrict__ offs1,
                                                                     const uint32_t* __restrict__ tile_tab, int low_bits,
                                                                     uint32_t* __restrict__ hist16) {
  __shared__ uint32_t cnt[256], where[3];
  if (blockIdx.x >= tile_tab[256]) return;
  if (threadIdx.x == 0) msd_locate_tile(offs1, tile_tab, where);
  if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t seg = where[0], lo = where[1], hi = where[2];
  constexpr int ITEMS = MSD_TILE / MSD_M_THREADS;
  uint32_t key[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    const uint32_t idx = lo + k * MSD_M_THREADS + threadIdx.x;
    key[k] = idx < hi ? __ldg(in_keys + idx) : 0u;
  }
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    const uint32_t idx = lo + k * MSD_M_THREADS + threadIdx.x;
    if (idx < hi) atomicAdd(&cnt[((key[k] - 1u) >> low_bits) & 255u], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 256 && cnt[threadIdx.x]) atomicAdd(&hist16[(seg << 8) + threadIdx.x], cnt[threadIdx.x]);
}

// hist16[65536] -> offs16[65537], cursor16 (a copy k_msd_mid advances).  One block of 1024 threads: warp w owns the
// 2048 consecutive groups [2048 w, 2048 (w + 1)) and walks them 32 at a time (coalesced loads and stores, warp scans).
static __global__ void __launch_bounds__(MSD_S_THREADS) k_msd_scan16(const uint32_t* __restrict__ hist16, uint32_t* __restrict__ offs16,
                                                                     uint32_t* __restrict__ cursor16) {
  __shared__ uint32_t warp_sums[32];
  constexpr int PER_WARP = MSD_GROUPS / (MSD_S_THREADS / 32);      // 2048
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t* src = hist16 + warp * PER_WARP;
  uint32_t sum = 0;
#pragma unroll 8
  for (int i = 0; i < PER_WARP / 32; i++) sum += src[i * 32 + lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) warp_sums[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_sums[lane] = w;                             // inclusive over warps
  }
  __syncthreads();
  uint32_t run = warp ? warp_sums[warp - 1] : 0u;
#pragma unroll 4
  for (int i = 0; i < PER_WARP / 32; i++) {
    const uint32_t v = src[i * 32 + lane];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    const uint32_t ex = run + x - v;
    offs16[warp * PER_WARP + i * 32 + lane] = ex;
    cursor16[warp * PER_WARP + i * 32 + lane] = ex;
    run += __shfl_sync(0xffffffffu, x, 31);
  }
  if (threadIdx.x == MSD_S_THREADS - 1) offs16[MSD_GROUPS] = run;
}

static __global__ void __launch_bounds__(MSD_M_THREADS) k_msd_mid(const uint32_t* __restrict__ in_keys, const uint32_t* __restrict__ in_refs,
                                                                  const uint32_t* __restrict__ offs1, const uint32_t* __restrict__ tile_tab,
                                                                  uint32_t* __restrict__ cursor16, int low_bits,
                                                                  uint32_t* __restrict__ out_keys, uint32_t* __restrict__ out_refs) {
  __shared__ uint32_t cnt[256], dstart[256], gbase[256], tmp[MSD_M_THREADS / 32], where[3];
  __shared__ uint32_t st_keys[MSD_TILE], st_refs[MSD_TILE];       // the tile grouped by sub-bin, so that it leaves in runs
  if (blockIdx.x >= tile_tab[256]) return;
  if (threadIdx.x == 0) msd_locate_tile(offs1, tile_tab, where);
  if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t seg = where[0], lo = where[1], hi = where[2];
  constexpr int ITEMS = MSD_TILE / MSD_M_THREADS;
  uint32_t key[ITEMS], ref[ITEMS], rank[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    const uint32_t idx = lo + k * MSD_M_THREADS + threadIdx.x;
    key[k] = idx < hi ? __ldg(in_keys + idx) : 0u;
    ref[k] = idx < hi ? __ldg(in_refs + idx) : 0u;
  }
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    const uint32_t idx = lo + k * MSD_M_THREADS + threadIdx.x;
    if (idx < hi) rank[k] = atomicAdd(&cnt[((key[k] - 1u) >> low_bits) & 255u], 1u);
  }
  __syncthreads();
  {                                                  // exclusive scan of the 256 counts (warps 0..7), global reservations
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t mine = threadIdx.x < 256 ? cnt[threadIdx.x] : 0u;
    uint32_t x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) tmp[warp] = x;
    __syncthreads();
    if (threadIdx.x < 256) {
      uint32_t off = 0;
      for (int w = 0; w < warp; w++) off += tmp[w];
      dstart[threadIdx.x] = off + x - mine;
      gbase[threadIdx.x] = mine ? atomicAdd(&cursor16[(seg << 8) + threadIdx.x], mine) : 0u;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    const uint32_t idx = lo + k * MSD_M_THREADS + threadIdx.x;
    if (idx < hi) {
      const uint32_t q = dstart[((key[k] - 1u) >> low_bits) & 255u] + rank[k];
      st_keys[q] = key[k];
      st_refs[q] = ref[k];
    }
  }
  __syncthreads();
  const uint32_t count = hi - lo;
  for (uint32_t q = threadIdx.x; q < count; q += MSD_M_THREADS) {
    const uint32_t kk = st_keys[q];
    const uint32_t sub = ((kk - 1u) >> low_bits) & 255u;
    const uint32_t dst = gbase[sub] + (q - dstart[sub]);
    out_keys[dst] = kk;
    out_refs[dst] = st_refs[q];
  }
}

// Block b sorts groups [b * gpb, (b + 1) * gpb) by the key's low `low_bits` (1..8) bits.  A range
// of up to MSD_LOW_CAP pairs is ranked from registers, re-ordered in shared memory and written back in order (coalesced);
// a larger one is histogrammed first and then takes the same steps MSD_LOW_CAP pairs at a time.
#ifndef MIRA_MSD_L_ITEMS
#define MIRA_MSD_L_ITEMS 16
#endif
constexpr int MSD_L_THREADS = 256;
constexpr int MSD_LOW_ITEMS = MIRA_MSD_L_ITEMS;
constexpr int MSD_LOW_CAP = MSD_L_THREADS * MSD_LOW_ITEMS;      // 4096 pairs, 32 KiB
constexpr int MSD_LOW_MAX_AVG = 3400;                           // groups average at most this for the one-chunk path to be the common case
constexpr int MSD_LOW_MAX_AVG_CHUNKED = 12000;                  // ... and this with the chunked path (a 2^26-point commit: 11,264)

static __global__ void __launch_bounds__(MSD_L_THREADS) k_msd_low(const uint32_t* __restrict__ in_keys, const uint32_t* __restrict__ in_refs,
                                                                  const uint32_t* __restrict__ offs16, uint32_t gpb, int low_bits,
                                                                  uint32_t* __restrict__ out_keys, uint32_t* __restrict__ out_refs) {
  __shared__ uint32_t cnt[256], tmp[8];
  __shared__ uint32_t st_keys[MSD_LOW_CAP], st_refs[MSD_LOW_CAP];
  // gpb (a power of two) adjacent groups are one contiguous range of pairs that share the top bits, so sorting the
  // range by the low bits PLUS the log2(gpb) bits above them sorts every group in it: one pass per block whatever gpb is
  // (the host keeps low_bits + log2(gpb) <= 8)
  uint32_t span_bits = (uint32_t)low_bits;
  for (uint32_t x = gpb; x > 1; x >>= 1) span_bits++;
  const uint32_t mask = (1u << span_bits) - 1u;
  {
    const uint32_t g = blockIdx.x * gpb;
    const uint32_t lo = offs16[g], hi = offs16[g + gpb];
    if (lo == hi) return;                            // uniform per block: offs16 is read by every thread alike
    cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t total;
    if (hi - lo <= (uint32_t)MSD_LOW_CAP) {
      uint32_t key[MSD_LOW_ITEMS], ref[MSD_LOW_ITEMS], rank[MSD_LOW_ITEMS];
#pragma unroll
      for (int k = 0; k < MSD_LOW_ITEMS; k++) {
        const uint32_t i = lo + k * MSD_L_THREADS + threadIdx.x;
        key[k] = i < hi ? __ldg(in_keys + i) : 0u;
        ref[k] = i < hi ? __ldg(in_refs + i) : 0u;
      }
#pragma unroll
      for (int k = 0; k < MSD_LOW_ITEMS; k++) {
        const uint32_t i = lo + k * MSD_L_THREADS + threadIdx.x;
        if (i < hi) rank[k] = atomicAdd(&cnt[(key[k] - 1u) & mask], 1u);
      }
      __syncthreads();
      const uint32_t ex = ds_scan256(cnt[threadIdx.x], tmp, total);
      cnt[threadIdx.x] = ex;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < MSD_LOW_ITEMS; k++) {
        const uint32_t i = lo + k * MSD_L_THREADS + threadIdx.x;
        if (i < hi) {
          const uint32_t q = cnt[(key[k] - 1u) & mask] + rank[k];
          st_keys[q] = key[k];
          st_refs[q] = ref[k];
        }
      }
      __syncthreads();
      for (uint32_t q = threadIdx.x; q < hi - lo; q += MSD_L_THREADS) {
        out_keys[lo + q] = st_keys[q];
        out_refs[lo + q] = st_refs[q];
      }
      __syncthreads();
    } else {
      // oversized range (a 2^25..2^26-point commit, the top window's heavy buckets, skewed vectors): histogram of the
      // whole range, then MSD_LOW_CAP pairs at a time through the same rank / re-order / write-in-runs steps, every bin's
      // write position carried from chunk to chunk
      __shared__ uint32_t pos[256], cstart[256];
      for (uint32_t i = lo + threadIdx.x; i < hi; i += MSD_L_THREADS) atomicAdd(&cnt[(__ldg(in_keys + i) - 1u) & mask], 1u);
      __syncthreads();
      const uint32_t ex = ds_scan256(cnt[threadIdx.x], tmp, total);
      pos[threadIdx.x] = lo + ex;
      for (uint32_t base = lo; base < hi; base += MSD_LOW_CAP) {
        const uint32_t top = base + MSD_LOW_CAP < hi ? base + MSD_LOW_CAP : hi;
        cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t key[MSD_LOW_ITEMS], ref[MSD_LOW_ITEMS], rank[MSD_LOW_ITEMS];
#pragma unroll
        for (int k = 0; k < MSD_LOW_ITEMS; k++) {
          const uint32_t i = base + k * MSD_L_THREADS + threadIdx.x;
          key[k] = i < top ? __ldg(in_keys + i) : 0u;
          ref[k] = i < top ? __ldg(in_refs + i) : 0u;
        }
#pragma unroll
        for (int k = 0; k < MSD_LOW_ITEMS; k++) {
          const uint32_t i = base + k * MSD_L_THREADS + threadIdx.x;
          if (i < top) rank[k] = atomicAdd(&cnt[(key[k] - 1u) & mask], 1u);
        }
        __syncthreads();
        const uint32_t mine = cnt[threadIdx.x];
        cstart[threadIdx.x] = ds_scan256(mine, tmp, total);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < MSD_LOW_ITEMS; k++) {
          const uint32_t i = base + k * MSD_L_THREADS + threadIdx.x;
          if (i < top) {
            const uint32_t q = cstart[(key[k] - 1u) & mask] + rank[k];
            st_keys[q] = key[k];
            st_refs[q] = ref[k];
          }
        }
        __syncthreads();
        for (uint32_t q = threadIdx.x; q < top - base; q += MSD_L_THREADS) {
          const uint32_t kk = st_keys[q];
          const uint32_t b = (kk - 1u) & mask;
          const uint32_t dst = pos[b] + (q - cstart[b]);
          out_keys[dst] = kk;
          out_refs[dst] = st_refs[q];
        }
        __syncthreads();
        pos[threadIdx.x] += mine;                    // thread t owns bin t
      }
    }
  }
}

// Bit-length histogram of a SAMPLE of the scalars (n_chunks chunks of chunk_len consecutive scalars, `stride` apart):
// hist[L] += 1 for every sampled scalar whose canonical value has bit length L (0 for zero).  Feeds
// choose_window_sampled.  `src_stride` = distance between chunks in the source, in scalars.
constexpr int SAMPLE_HASH_BINS = 1024;
template <class SF>
__global__ void __launch_bounds__(256) k_bitlen_hist(const void* __restrict__ scalars, uint32_t n_chunks, uint32_t chunk_len,
                                                     size_t src_stride, uint32_t* __restrict__ hist) {
  // hist[0..256]: bit lengths; hist[260 .. 260 + SAMPLE_HASH_BINS): how many sampled scalars of bit length > 32 hash to
  // each bin — a value that makes up more than a per cent or two of the vector stands out of the ~samples / bins
  // background (heavy-hitter test for the MSD partition, pipeline.cuh)
  __shared__ uint32_t sh[257];
  for (int i = threadIdx.x; i < 257; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_chunks * chunk_len) {
    size_t idx = (size_t)(t / chunk_len) * src_stride + (t % chunk_len);
    Fe<SF> s = fe_to_canonical(fe_load<SF>(reinterpret_cast<const char*>(scalars) + idx * 32));
    int L = 0;
#pragma unroll
    for (int k = 7; k >= 0; k--)
      if (L == 0 && s.v[k]) L = 32 * k + 32 - __clz(s.v[k]);
    atomicAdd(&sh[L], 1u);
    if (L > 32) {
      uint32_t h = s.v[0] * 0x9E3779B1u;
#pragma unroll
      for (int k = 1; k < 8; k++) h = (h ^ s.v[k]) * 0x85EBCA77u;
      atomicAdd(&hist[260 + ((h >> 16) & (SAMPLE_HASH_BINS - 1))], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 257; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// ------------------------------------------------------------------ accumulate
// Sorted non-zero entries occupy [0, n_sorted).  Thread t owns entries [t*L, (t+1)*L): every thread
// does the same number of mixed adds whatever the bucket-size distribution is.  A run (maximal
// range of equal keys) that lies wholly inside one chunk is written straight to its bucket; a run
// that touches a chunk edge and continues beyond it becomes a "partial" (slot 2t for the chunk's
// first run, 2t+1 for its last) that k_combine stitches.
template <class CF>
__device__ __forceinline__ Affine<CF> load_ref(const void* __restrict__ table, uint32_t ref) {
  Affine<CF> p = aff_load<CF>(reinterpret_cast<const char*>(table) + (size_t)(ref & ~REF_NEG) * 64);
  if (ref & REF_NEG) {
    if (!fe_is_zero(p.y)) p.y = fe_neg(p.y);
  }
  return p;
}

// L2 prefetch distance of the table gather, in entries (0 = off).  Measured at 2^24 points, accumulation phase in ms
// (profiles/r02_acc_variants.txt): off 31.99 | 2 ahead 31.78 | 6 ahead 32.04; __launch_bounds__(128, 5) 32.27.
#ifndef MIRA_ACC_PREFETCH
#define MIRA_ACC_PREFETCH 2
#endif
#ifndef MIRA_ACC_MIN_BLOCKS
#define MIRA_ACC_MIN_BLOCKS 1
#endif
// DIRECT: entry e's point is the e-th element of a dense affine list (`table`; the output of the batched-affine
// levels, affine_levels.cuh) instead of the table entry named by srefs[e].
template <class CF, bool DIRECT = false>
__global__ void __launch_bounds__(128, MIRA_ACC_MIN_BLOCKS) k_accumulate(const uint32_t* __restrict__ skeys,
                                                    const uint32_t* __restrict__ srefs,
                                                    const uint32_t* __restrict__ n_ptr, int L,
                                                    const void* __restrict__ table, void* __restrict__ buckets,
                                                    uint32_t* __restrict__ part_keys, void* __restrict__ part_pts, int add_mode) {
  // add_mode: the buckets already hold the sums of earlier scalar slices (host-buffer commits are pipelined slice
  // by slice behind their H2D copies); the piece that owns a run's left end then starts from the bucket's value
  // instead of the identity, so folding a slice in costs one 128-byte load per run and no extra group operation.
  const uint32_t n_sorted = *n_ptr;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  size_t begin = (size_t)t * L;
  if (begin >= n_sorted) return;
  size_t end = begin + L < n_sorted ? begin + L : n_sorted;
  uint32_t prev_key = begin ? skeys[begin - 1] : 0u;
  uint32_t next_key = end < n_sorted ? skeys[end] : 0u;
  uint32_t cur = skeys[begin];
  bool first = true;
  uint32_t pk0 = 0, pk1 = 0;
  Xyzz<CF> acc = xyzz_identity<CF>();
  if (add_mode && prev_key != cur) acc = xyzz_load_shared<CF>(reinterpret_cast<const char*>(buckets) + (size_t)cur * 128);
  for (size_t e = begin; e < end; e++) {
    uint32_t k = skeys[e];
    if (k != cur) {
      // the finished run ends strictly inside the chunk: it can only be open to the left
      if (first && prev_key == cur) {
        pk0 = cur | PK_OPEN_LEFT;
        xyzz_store<CF>(reinterpret_cast<char*>(part_pts) + (size_t)(2 * t) * 128, acc);
      } else {
        xyzz_store<CF>(reinterpret_cast<char*>(buckets) + (size_t)cur * 128, acc);
      }
      first = false;
      cur = k;
      if (add_mode) acc = xyzz_load_shared<CF>(reinterpret_cast<const char*>(buckets) + (size_t)cur * 128);
      else acc = xyzz_identity<CF>();
    }
#if MIRA_ACC_PREFETCH
    if (!DIRECT && e + MIRA_ACC_PREFETCH < end) {    // pull the point two adds ahead into L2: the gather is a dependent DRAM access
      const char* nxt = reinterpret_cast<const char*>(table) + (size_t)(srefs[e + MIRA_ACC_PREFETCH] & ~REF_NEG) * 64;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
    }
#endif
    Affine<CF> p = DIRECT ? aff_load<CF>(reinterpret_cast<const char*>(table) + e * 64) : load_ref<CF>(table, srefs[e]);
    xyzz_madd(acc, p);
  }
  {
    bool open_left = first && prev_key == cur;
    bool open_right = next_key == cur;
    if (!open_left && !open_right) {
      xyzz_store<CF>(reinterpret_cast<char*>(buckets) + (size_t)cur * 128, acc);
    } else {
      uint32_t pk = cur | (open_left ? PK_OPEN_LEFT : 0u) | (open_right ? PK_OPEN_RIGHT : 0u);
      int slot = first ? 0 : 1;
      if (slot == 0) pk0 = pk; else pk1 = pk;
      xyzz_store<CF>(reinterpret_cast<char*>(part_pts) + (size_t)(2 * t + slot) * 128, acc);
    }
  }
  part_keys[2 * t] = pk0;
  part_keys[2 * t + 1] = pk1;
}

// ------------------------------------------------------------------ accumulate with thread-local pair pre-addition
// The same chunked walk as k_accumulate, but entries 2k and 2k+1 of a chunk, when they belong to the same bucket,
// are first added in AFFINE coordinates and only their sum goes through the XYZZ mixed addition: 6 field products
// (5M + 1S, Montgomery's trick included) + one 10-product mixed addition per TWO table points instead of two mixed
// additions, i.e. ~2,090 wide MACs against ~2,540.  The batch that shares an inversion is the thread's own chunk
// (~120 additions): no inter-thread exchange; every lane inverts its own product with the branch-free safegcd of
// inv30.cuh (~14,000 instructions).  Three launches over the same chunking (thread t owns entries [t*L, (t+1)*L)):
//   k_pair_up    descending: suffix[k] = prod_{j > k} d_j over the chunk's pairable pairs, d_k = x(2k+1) - x(2k),
//                to a [pair][thread] scratch array (32 B, coalesced); then inv[t] = 1 / prod_j d_j.  Gathers the
//                x-coordinates only; memory-latency-bound, many warps per SM.
//   k_pair_add   ascending: 1/d_k = inv * suffix[k], inv *= d_k; lambda = (y2 - y1)/d_k; the sum goes to sums[k]
//                ([pair][thread], 64 B); (0, 0) marks "not paired" (a sum is never the identity: opposite points
//                are not pairable).  Integer-multiply-bound, five products per pair.
//   k_pair_acc   k_accumulate's run bookkeeping over the sums (or, where not paired, the two table points).
// One fused kernel was measured first (profiles/r02_pair_preadd.txt): with its three loops and the inversion in one
// instruction stream (80 KB) and the warps of an SM in different loops, 22 % of all stall samples were instruction
// fetch, and the latency-bound up pass (25 % of the samples) ran at 3 warps per scheduler.
// A pair is skipped (both points take the mixed addition, which handles every exceptional case) when the keys
// differ, when d_k = 0 (equal or opposite points) or when an x-coordinate is 0 (the identity is stored as (0, 0)).
template <class CF>
__device__ __forceinline__ Fe<CF> load_ref_x(const void* __restrict__ table, uint32_t ref) {
  return fe_load<CF>(reinterpret_cast<const char*>(table) + (size_t)(ref & ~REF_NEG) * 64);
}
template <class CF>
__device__ __forceinline__ bool pa_pairable(const Fe<CF>& x1, const Fe<CF>& x2, const Fe<CF>& d) {
  return !fe_is_zero(d) && !fe_is_zero(x1) && !fe_is_zero(x2);
}
constexpr int PAIR_UP_AHEAD = 4;       // pairs whose x-coordinates are in flight while one product is computed

template <class CF>
__global__ void __launch_bounds__(128) k_pair_up(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ srefs,
                                                 const uint32_t* __restrict__ n_ptr, int L, const void* __restrict__ table,
                                                 void* __restrict__ suffix, void* __restrict__ inv_out) {
  const uint32_t n_sorted = *n_ptr;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  const size_t begin = (size_t)t * L;
  if (begin >= n_sorted) return;
  const size_t end = begin + L < n_sorted ? begin + L : n_sorted;
  const int m = (int)((end - begin + 1) >> 1);                 // pairs (the last one may hold a single entry)
  char* const my_suffix = reinterpret_cast<char*>(suffix) + (size_t)t * 32;
  Fe<CF> r = fe_one<CF>();
  // ring of PAIR_UP_AHEAD pairs: the gathers of pairs k-1 .. k-AHEAD are in flight during pair k's product
  Fe<CF> x1[PAIR_UP_AHEAD], x2[PAIR_UP_AHEAD];
  bool same[PAIR_UP_AHEAD];
#pragma unroll
  for (int a = 0; a < PAIR_UP_AHEAD; a++) {
    const int k = m - 1 - a;
    same[a] = false;
    x1[a] = fe_zero<CF>();
    x2[a] = fe_zero<CF>();
    if (k >= 0) {
      const size_t e = begin + 2 * (size_t)k;
      if (e + 1 < end && skeys[e] == skeys[e + 1]) {
        same[a] = true;
        x1[a] = load_ref_x<CF>(table, srefs[e]);
        x2[a] = load_ref_x<CF>(table, srefs[e + 1]);
      }
    }
  }
  for (int k0 = m - 1; k0 >= 0; k0 -= PAIR_UP_AHEAD) {
#pragma unroll
    for (int a = 0; a < PAIR_UP_AHEAD; a++) {
      const int k = k0 - a;
      if (k < 0) break;
      const Fe<CF> cx1 = x1[a], cx2 = x2[a];
      const bool csame = same[a];
      {                                                         // refill slot a with pair k - AHEAD
        const int kn = k - PAIR_UP_AHEAD;
        same[a] = false;
        if (kn >= 0) {
          const size_t e = begin + 2 * (size_t)kn;
          if (skeys[e] == skeys[e + 1]) {
            same[a] = true;
            x1[a] = load_ref_x<CF>(table, srefs[e]);
            x2[a] = load_ref_x<CF>(table, srefs[e + 1]);
          }
        }
      }
      if (csame) {
        Fe<CF> d = fe_sub(cx2, cx1);
        if (pa_pairable(cx1, cx2, d)) {
          fe_store<CF>(my_suffix + (size_t)k * nth * 32, r);
          r = fe_mul(r, d);
        }
      }
    }
  }
  fe_store<CF>(reinterpret_cast<char*>(inv_out) + (size_t)t * 32, fe_inv_uniform(r));
}

template <class CF>
__global__ void __launch_bounds__(128) k_pair_add(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ srefs,
                                                  const uint32_t* __restrict__ n_ptr, int L, const void* __restrict__ table,
                                                  const void* __restrict__ suffix, const void* __restrict__ inv_in,
                                                  void* __restrict__ sums) {
  const uint32_t n_sorted = *n_ptr;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  const size_t begin = (size_t)t * L;
  if (begin >= n_sorted) return;
  const size_t end = begin + L < n_sorted ? begin + L : n_sorted;
  const int m = (int)((end - begin + 1) >> 1);
  const char* const my_suffix = reinterpret_cast<const char*>(suffix) + (size_t)t * 32;
  char* const my_sums = reinterpret_cast<char*>(sums) + (size_t)t * 64;
  Fe<CF> inv = fe_load_plain<CF>(reinterpret_cast<const char*>(inv_in) + (size_t)t * 32);
  for (int k = 0; k < m; k++) {
    const size_t e = begin + 2 * (size_t)k;
    const bool have_q = e + 1 < end;
    Affine<CF> S;
    S.x = fe_zero<CF>();
    S.y = fe_zero<CF>();
    if (have_q && skeys[e] == skeys[e + 1]) {
      if (e + 5 < end) {                                        // pull the points two pairs ahead into L2
        const char* n0 = reinterpret_cast<const char*>(table) + (size_t)(srefs[e + 4] & ~REF_NEG) * 64;
        const char* n1 = reinterpret_cast<const char*>(table) + (size_t)(srefs[e + 5] & ~REF_NEG) * 64;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(n0));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(n1));
      }
      const Affine<CF> P = load_ref<CF>(table, srefs[e]);
      const Affine<CF> Q = load_ref<CF>(table, srefs[e + 1]);
      Fe<CF> d = fe_sub(Q.x, P.x);
      if (pa_pairable(P.x, Q.x, d)) {
        Fe<CF> sfx = fe_load<CF>(my_suffix + (size_t)k * nth * 32);
        Fe<CF> ik = fe_mul(inv, sfx);                           // 1 / d
        inv = fe_mul(inv, d);
        Fe<CF> lam = fe_mul(fe_sub(Q.y, P.y), ik);
        S.x = fe_sub(fe_sub(fe_sqr(lam), P.x), Q.x);
        S.y = fe_sub(fe_mul(lam, fe_sub(P.x, S.x)), P.y);
      }
    }
    aff_store<CF>(my_sums + (size_t)k * nth * 64, S);
  }
}

template <class CF>
__global__ void __launch_bounds__(128, MIRA_ACC_MIN_BLOCKS) k_pair_acc(const uint32_t* __restrict__ skeys,
                                                  const uint32_t* __restrict__ srefs,
                                                  const uint32_t* __restrict__ n_ptr, int L,
                                                  const void* __restrict__ table, void* __restrict__ buckets,
                                                  uint32_t* __restrict__ part_keys, void* __restrict__ part_pts, int add_mode,
                                                  const void* __restrict__ sums) {
  const uint32_t n_sorted = *n_ptr;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  const size_t begin = (size_t)t * L;
  if (begin >= n_sorted) return;
  const size_t end = begin + L < n_sorted ? begin + L : n_sorted;
  const char* const my_sums = reinterpret_cast<const char*>(sums) + (size_t)t * 64;
  const uint32_t prev_key = begin ? skeys[begin - 1] : 0u;
  const uint32_t next_key = end < n_sorted ? skeys[end] : 0u;
  uint32_t cur = skeys[begin];
  bool first = true;
  uint32_t pk0 = 0, pk1 = 0;
  Xyzz<CF> acc = xyzz_identity<CF>();
  if (add_mode && prev_key != cur) acc = xyzz_load_shared<CF>(reinterpret_cast<const char*>(buckets) + (size_t)cur * 128);
  for (size_t e = begin; e < end;) {                            // ONE copy of the mixed addition in the instruction stream
    const uint32_t key = skeys[e];
    const uint32_t off = (uint32_t)(e - begin);
    Affine<CF> P;
    P.x = fe_zero<CF>();
    P.y = fe_zero<CF>();
    if (!(off & 1u)) P = aff_load<CF>(my_sums + (size_t)(off >> 1) * nth * 64);
    if (aff_is_identity(P)) {                                   // rare: run edges, equal / opposite / identity points
      P = load_ref<CF>(table, srefs[e]);
      e += 1;
    } else {
      e += 2;
    }
    if (key != cur) {
      if (first && prev_key == cur) {
        pk0 = cur | PK_OPEN_LEFT;
        xyzz_store<CF>(reinterpret_cast<char*>(part_pts) + (size_t)(2 * t) * 128, acc);
      } else {
        xyzz_store<CF>(reinterpret_cast<char*>(buckets) + (size_t)cur * 128, acc);
      }
      first = false;
      cur = key;
      if (add_mode) acc = xyzz_load_shared<CF>(reinterpret_cast<const char*>(buckets) + (size_t)cur * 128);
      else acc = xyzz_identity<CF>();
    }
    xyzz_madd(acc, P);
  }
  {
    bool open_left = first && prev_key == cur;
    bool open_right = next_key == cur;
    if (!open_left && !open_right) {
      xyzz_store<CF>(reinterpret_cast<char*>(buckets) + (size_t)cur * 128, acc);
    } else {
      uint32_t pk = cur | (open_left ? PK_OPEN_LEFT : 0u) | (open_right ? PK_OPEN_RIGHT : 0u);
      int slot = first ? 0 : 1;
      if (slot == 0) pk0 = pk; else pk1 = pk;
      xyzz_store<CF>(reinterpret_cast<char*>(part_pts) + (size_t)(2 * t + slot) * 128, acc);
    }
  }
  part_keys[2 * t] = pk0;
  part_keys[2 * t + 1] = pk1;
}

// One thread per partial slot; the leftmost piece of each straddling run (not OPEN_LEFT) walks right
// over slot 0 of the following chunks while the run stays open, and writes the bucket.  A run that
// spans more than HEAVY_CHUNKS chunks (a "heavy" bucket: witness-like scalars put a large share of all
// pairs into a few buckets) is not walked serially: its leader chunk is appended to `heavy` and
// k_combine_heavy folds its pieces with a whole block.
constexpr uint32_t HEAVY_CHUNKS = 8;
constexpr uint32_t HUGE_CHUNKS = 256;     // runs spanning more chunks than this get a whole block, the rest a warp

template <class CF>
__global__ void __launch_bounds__(128) k_combine(const uint32_t* __restrict__ skeys,
                                                 const uint32_t* __restrict__ part_keys,
                                                 const void* __restrict__ part_pts,
                                                 const uint32_t* __restrict__ n_ptr, int L,
                                                 void* __restrict__ buckets, uint32_t* __restrict__ heavy,
                                                 uint32_t heavy_cap) {
  // heavy = [count_medium, count_huge, medium leaders (heavy_cap), huge leaders (heavy_cap)]
  // One thread per CHUNK: a chunk has at most one piece that leads a straddling run (its last run if that continues
  // into the next chunk — slot 1, or slot 0 when the whole chunk is one run that starts at its left edge); a thread
  // per slot left every other lane idle.
  const uint32_t n = *n_ptr;
  const uint32_t n_chunks = (n + L - 1) / L;
  uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= n_chunks) return;
  const uint2 pks = reinterpret_cast<const uint2*>(part_keys)[chunk];
  uint32_t q = pks.y ? 2 * chunk + 1 : 2 * chunk;
  uint32_t pk = pks.y ? pks.y : pks.x;
  if (pk == 0 || (pk & PK_OPEN_LEFT)) return;
  {
    uint64_t probe = (uint64_t)(chunk + HEAVY_CHUNKS) * (uint32_t)L;
    if (probe < n && skeys[probe] == (pk & PK_KEY_MASK)) {
      uint64_t probe2 = (uint64_t)(chunk + HUGE_CHUNKS) * (uint32_t)L;
      bool huge = probe2 < n && skeys[probe2] == (pk & PK_KEY_MASK);
      uint32_t slot = atomicAdd(&heavy[huge ? 1 : 0], 1u);
      if (slot < heavy_cap) {            // heavy_cap >= n_chunks / HEAVY_CHUNKS + 1: cannot overflow
        heavy[2 + (huge ? heavy_cap : 0u) + slot] = q;
        return;
      }
    }
  }
  Xyzz<CF> acc = xyzz_load<CF>(reinterpret_cast<const char*>(part_pts) + (size_t)q * 128);
  while (pk & PK_OPEN_RIGHT) {
    chunk++;
    uint32_t nq = 2 * chunk;
    pk = part_keys[nq];
    Xyzz<CF> nxt = xyzz_load<CF>(reinterpret_cast<const char*>(part_pts) + (size_t)nq * 128);
    xyzz_add(acc, nxt);
  }
  xyzz_store<CF>(reinterpret_cast<char*>(buckets) + (size_t)(pk & PK_KEY_MASK) * 128, acc);
}

// Heavy runs.  The run's pieces are the leader slot q0 and slot 0 of chunks t0+1 .. t1, where t1 is the last chunk
// whose first pair still carries the key (binary search over the sorted keys).  GROUP threads (a warp for runs of
// up to HUGE_CHUNKS chunks - e.g. the top window, whose few significant bits put thousands of pairs into each of
// a few thousand buckets - or a whole block beyond that) sum the pieces strided, then fold through shared memory.
constexpr int HV_THREADS = 256;
template <class CF, int GROUP>
__global__ void __launch_bounds__(HV_THREADS) k_combine_heavy(const uint32_t* __restrict__ skeys,
                                                              const uint32_t* __restrict__ part_keys,
                                                              const void* __restrict__ part_pts,
                                                              const uint32_t* __restrict__ n_ptr, int L,
                                                              void* __restrict__ buckets,
                                                              const uint32_t* __restrict__ heavy, uint32_t heavy_cap) {
  __shared__ uint4 sm[HV_THREADS * 8];
  constexpr int GROUPS = HV_THREADS / GROUP;
  const uint32_t n = *n_ptr;
  const uint32_t n_chunks = (n + L - 1) / L;
  const uint32_t which = GROUP == HV_THREADS ? 1u : 0u;
  const uint32_t* list = heavy + 2 + which * heavy_cap;
  uint32_t n_heavy = heavy[which] < heavy_cap ? heavy[which] : heavy_cap;
  const uint32_t g = threadIdx.x / GROUP, lane = threadIdx.x % GROUP;
  // every group of a block runs the same number of iterations (block-wide barriers below)
  const uint32_t iters = (n_heavy + gridDim.x * GROUPS - 1) / (gridDim.x * GROUPS);
  for (uint32_t it = 0; it < iters; it++) {
    const uint32_t h = (it * gridDim.x + blockIdx.x) * GROUPS + g;
    const bool active = h < n_heavy;
    Xyzz<CF> acc = xyzz_identity<CF>();
    uint32_t key = 0;
    if (active) {
      const uint32_t q0 = list[h];
      key = part_keys[q0] & PK_KEY_MASK;
      const uint32_t t0 = q0 >> 1;
      // t1 = largest chunk index with skeys[t1 * L] == key  (keys are sorted, chunk t0+1 qualifies)
      uint32_t lo = t0 + 1, hi = n_chunks - 1;
      while (lo < hi) {
        uint32_t mid = lo + (hi - lo + 1) / 2;
        if (skeys[(uint64_t)mid * (uint32_t)L] <= key) lo = mid; else hi = mid - 1;
      }
      const uint32_t t1 = lo;
      if (lane == 0) acc = xyzz_load_shared<CF>(reinterpret_cast<const char*>(part_pts) + (size_t)q0 * 128);
      for (uint32_t t = t0 + 1 + lane; t <= t1; t += GROUP) {
        Xyzz<CF> p = xyzz_load_shared<CF>(reinterpret_cast<const char*>(part_pts) + (size_t)(2 * t) * 128);
        xyzz_add(acc, p);
      }
    }
    char* my = reinterpret_cast<char*>(sm) + threadIdx.x * 128;
    xyzz_store<CF>(my, acc);
    __syncthreads();
    for (int s = GROUP >> 1; s > 0; s >>= 1) {
      if ((int)lane < s) {
        Xyzz<CF> o = xyzz_load_shared<CF>(reinterpret_cast<char*>(sm) + (threadIdx.x + s) * 128);
        xyzz_add(acc, o);
        xyzz_store<CF>(my, acc);
      }
      __syncthreads();
    }
    if (active && lane == 0) xyzz_store<CF>(reinterpret_cast<char*>(buckets) + (size_t)key * 128, acc);
    __syncthreads();
  }
}

// ------------------------------------------------------------------ bucket reduction
// buckets[1..B]; thread t owns b in [t*m+1, (t+1)*m]:  out[t] = sum (b - t*m) * bucket[b] + (t*m) * sum bucket[b]
template <class CF>
__global__ void __launch_bounds__(128) k_reduce_chunks(const void* __restrict__ buckets_all, uint32_t B, uint32_t m,
                                                       void* __restrict__ out_all, size_t in_stride, size_t out_stride) {
  // blockIdx.y selects the bucket set of a batched commit (strides in bytes)
  const void* buckets = reinterpret_cast<const char*>(buckets_all) + blockIdx.y * in_stride;
  void* out = reinterpret_cast<char*>(out_all) + blockIdx.y * out_stride;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo = t * m;
  if (lo >= B) return;
  uint32_t hi = lo + m < B ? lo + m : B;
  Xyzz<CF> run = xyzz_identity<CF>(), acc = xyzz_identity<CF>();
  for (uint32_t b = hi; b > lo; b--) {
    Xyzz<CF> bk = xyzz_load<CF>(reinterpret_cast<const char*>(buckets) + (size_t)b * 128);
    xyzz_add(run, bk);
    xyzz_add(acc, run);
  }
  if (lo) {
    Xyzz<CF> w = xyzz_mul_u32(run, lo);
    xyzz_add(acc, w);
  }
  xyzz_store<CF>(reinterpret_cast<char*>(out) + (size_t)t * 128, acc);
}

// One LEVEL of the recursive form of the same reduction, without scalar multiplications.  For an array X[0..n) with
// S = sum (i + 1) * X[i], cut into chunks of m = 2^log_m elements:
//     S = sum_t A_t  +  sum_{t >= 1} t * (m * T_t),      A_t = sum_j (j + 1) * X[t*m + j],   T_t = sum_j X[t*m + j]
// so a thread computes A_t and T_t with the two running sums (2m additions), multiplies T_t by m with log_m doublings
// and hands Y[t - 1] = m * T_t to the next level, whose sum  sum (t' + 1) * Y[t']  is the same problem m times
// smaller.  Against k_reduce_chunks this replaces the ~1.5 * bits(t*m) group operations of the per-thread weighting
// (21 doublings + ~10 additions at 2^21 buckets) by log_m doublings.  a_out[t] receives A_t (all levels' A's are tree-
// summed at the end).  blockIdx.y selects the bucket set of a batched commit (strides in bytes).
template <class CF>
__global__ void __launch_bounds__(128) k_reduce_level(const void* __restrict__ in_all, uint32_t n, int log_m, void* __restrict__ a_all,
                                                      void* __restrict__ next_all, size_t in_stride, size_t a_stride, size_t next_stride) {
  const char* in = reinterpret_cast<const char*>(in_all) + blockIdx.y * in_stride;
  char* a_out = reinterpret_cast<char*>(a_all) + blockIdx.y * a_stride;
  char* next = reinterpret_cast<char*>(next_all) + blockIdx.y * next_stride;
  const uint32_t m = 1u << log_m;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t lo = (uint64_t)t * m;
  if (lo >= n) return;
  const uint32_t hi = lo + m < n ? (uint32_t)(lo + m) : n;
  Xyzz<CF> run = xyzz_identity<CF>(), acc = xyzz_identity<CF>();
  for (uint32_t i = hi; i > (uint32_t)lo; i--) {
    Xyzz<CF> x = xyzz_load<CF>(in + (size_t)(i - 1) * 128);
    xyzz_add(run, x);
    xyzz_add(acc, run);
  }
  xyzz_store<CF>(a_out + (size_t)t * 128, acc);
  if (t) {
    for (int k = 0; k < log_m; k++) run = xyzz_dbl(run);
    xyzz_store<CF>(next + (size_t)(t - 1) * 128, run);
  }
}

// Sum `n` XYZZ points into ceil(n / per_block) points: each block sums a contiguous slice.
template <class CF>
__global__ void __launch_bounds__(128) k_sum_points(const void* __restrict__ in_all, uint32_t n, uint32_t per_thread,
                                                    void* __restrict__ out_all, size_t in_stride = 0, size_t out_stride = 0) {
  __shared__ uint4 sm[128 * 8];   // 128 XYZZ points
  const void* in = reinterpret_cast<const char*>(in_all) + blockIdx.y * in_stride;
  void* out = reinterpret_cast<char*>(out_all) + blockIdx.y * out_stride;
  uint32_t per_block = per_thread * blockDim.x;
  uint32_t base = blockIdx.x * per_block + threadIdx.x * per_thread;
  Xyzz<CF> acc = xyzz_identity<CF>();
  for (uint32_t k = 0; k < per_thread; k++) {
    uint32_t idx = base + k;
    if (idx < n) {
      Xyzz<CF> p = xyzz_load<CF>(reinterpret_cast<const char*>(in) + (size_t)idx * 128);
      xyzz_add(acc, p);
    }
  }
  char* my = reinterpret_cast<char*>(sm) + threadIdx.x * 128;
  xyzz_store<CF>(my, acc);
  __syncthreads();
  // tree over the threads that hold data only (the last launches of a reduction see a handful of points: every level
  // skipped is one ~7 us group addition off the critical path of a small commit)
  const uint32_t first = blockIdx.x * per_block;
  const uint32_t have = first < n ? (n - first < per_block ? n - first : per_block) : 0u;
  const uint32_t holders = (have + per_thread - 1) / per_thread;
  int top = 1;
  while ((uint32_t)top < holders) top <<= 1;
  for (int s = top >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      Xyzz<CF> o = xyzz_load_shared<CF>(reinterpret_cast<char*>(sm) + (threadIdx.x + s) * 128);
      xyzz_add(acc, o);
      xyzz_store<CF>(my, acc);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) xyzz_store<CF>(reinterpret_cast<char*>(out) + (size_t)blockIdx.x * 128, acc);
}

template <class CF>
__global__ void k_finalize(const void* __restrict__ in_xyzz, void* __restrict__ out_affine) {
  if (threadIdx.x) return;       // one block per point (a batched commit normalises all its results in one launch)
  Xyzz<CF> p = xyzz_load<CF>(reinterpret_cast<const char*>(in_xyzz) + (size_t)blockIdx.x * 128);
  aff_store<CF>(reinterpret_cast<char*>(out_affine) + (size_t)blockIdx.x * 64, xyzz_to_affine(p));
}

// ------------------------------------------------------------------ fixed-base table
// table[j*n + i] = 2^(c*j) * P_i in affine form, j = 0..W-1.  One thread per point walks the windows, doubling in
// XYZZ without normalising in between; the normalisations of PC_GROUP consecutive windows share one inversion
// (Montgomery's trick; every lane inverts at the same time, hence the branch-uniform inversion).  2^24 points x 12
// windows: 1.48 s with one inversion per table entry, 0.85 s this way (the c doublings per entry remain).
constexpr int PC_GROUP = 8;
template <class CF>
__global__ void __launch_bounds__(128) k_precompute(const void* __restrict__ bases, uint32_t n, int c, int W,
                                                    void* __restrict__ table) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<CF> p = aff_load<CF>(reinterpret_cast<const char*>(bases) + (size_t)i * 64);
  aff_store<CF>(reinterpret_cast<char*>(table) + (size_t)i * 64, p);
  Xyzz<CF> cur = xyzz_from_affine(p);
  for (int j0 = 1; j0 < W; j0 += PC_GROUP) {
    const int g = W - j0 < PC_GROUP ? W - j0 : PC_GROUP;
    Xyzz<CF> q[PC_GROUP];
    Fe<CF> pre[PC_GROUP];
    Fe<CF> acc = fe_one<CF>();
#pragma unroll 1
    for (int t = 0; t < g; t++) {
      for (int k = 0; k < c; k++) cur = xyzz_dbl(cur);
      q[t] = cur;
      pre[t] = acc;
      if (!xyzz_is_identity(cur)) acc = fe_mul(acc, fe_mul(cur.zz, cur.zzz));
    }
    Fe<CF> inv = fe_inv_uniform(acc);
#pragma unroll 1
    for (int t = g - 1; t >= 0; t--) {
      Affine<CF> a;
      a.x = fe_zero<CF>();
      a.y = fe_zero<CF>();
      if (!xyzz_is_identity(q[t])) {
        Fe<CF> id = fe_mul(inv, pre[t]);                       // 1 / (ZZ * ZZZ)
        inv = fe_mul(inv, fe_mul(q[t].zz, q[t].zzz));
        a.x = fe_mul(q[t].x, fe_mul(id, q[t].zzz));            // X / ZZ
        a.y = fe_mul(q[t].y, fe_mul(id, q[t].zz));             // Y / ZZZ
      }
      aff_store<CF>(reinterpret_cast<char*>(table) + ((size_t)(j0 + t) * n + i) * 64, a);
    }
  }
}

// is_on_curve over the whole key (src/commitment.rs:145-146): flag != 0 if any point is off-curve
template <class CF>
__global__ void __launch_bounds__(256) k_check_on_curve(const void* __restrict__ bases, uint32_t n, uint32_t b_small,
                                                        int b_negative, uint32_t* __restrict__ flag) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<CF> p = aff_load<CF>(reinterpret_cast<const char*>(bases) + (size_t)i * 64);
  if (aff_is_identity(p)) return;
  Fe<CF> bc = fe_zero<CF>();
  bc.v[0] = b_small;
  bc = fe_from_canonical(bc);
  if (b_negative) bc = fe_neg(bc);
  Fe<CF> lhs = fe_sqr(p.y);
  Fe<CF> rhs = fe_add(fe_mul(fe_sqr(p.x), p.x), bc);
  if (!fe_eq(lhs, rhs)) atomicOr(flag, 1u);
}

}  // namespace mira
