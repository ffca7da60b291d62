// affine_levels.cuh — batched-affine pre-reduction of the sorted (bucket, point) list.
//
// The bucket sums of CommitmentKey::commit (/root/reference/src/commitment.rs:78-87, halo2's multiexp buckets) are
// sums of AFFINE table points.  An XYZZ mixed addition costs 8M + 2S; an affine addition costs 5M + 1S if the
// inversion of its denominator is shared by a large batch (Montgomery's trick: 3 of the 5 products).  One "level"
// adds the entries of every run of equal keys two by two, ~1000 additions sharing one inversion,
// and halves the list; after a few levels the XYZZ accumulation finishes the much shorter list
// (tools/microbench/affine_batch.cu: 0.10 ns per affine addition against 0.17 ns per mixed addition on B200).
//
// A level is four launches:
//   k_pa_up     per tile of PA_TILE consecutive entries: which entries pair up and how many entries the tile leaves
//               behind; per thread (64 consecutive entries, up to 32 additions) the running product of its denominators
//   k_pa_scan   exclusive scan of the tiles' counts -> where each tile writes, and the new list length
//   k_pa_inv    inverses of all the threads' products (Montgomery's trick again, one real inversion per 32 threads'
//               worth, every lane of a warp inverting at once)
//   k_pa_down   the additions; level 0 gathers its points from the fixed-base table through the sorted references,
//               later levels read the dense point list written by the level before
// (A single kernel with one inversion per block was measured first: the lone inverting thread cost 23 % of the
// block's issue slots and the barrier around it stalled the other warps; splitting removes both.)
// Pairing is by position inside a run (entries 2j and 2j+1 of the run are added); a run is cut at tile edges, which
// leaves at most one extra entry per tile and level.  The output is again sorted by key with the same per-key sums,
// so every exceptional case of the group law is handled where it occurs: an identity operand passes the other one
// through, P + P doubles (denominator 2y), P + (-P) yields the identity (0, 0), and those lanes feed a 1 into the
// shared product.
#pragma once
#include "msm_kernels.cuh"

namespace mira {

constexpr int PA_THREADS = 128;
constexpr int PA_WARPS = PA_THREADS / 32;
constexpr int PA_EPT = 64;                       // entries per thread (the pairing masks are 64-bit words)
constexpr int PA_MAX_ADDS = PA_EPT / 2;          // a thread owns at most this many additions
constexpr int PA_TILE = PA_THREADS * PA_EPT;     // entries per block
constexpr uint32_t PA_NO_KEY = 0xffffffffu;      // beyond the list / beyond the tile (real keys are < PK_KEY_MASK)
constexpr int PA_SK_WORDS = PA_TILE + PA_TILE / 64 + 4;

// keys of the tile in shared memory, one pad word per 64 so that a thread's stride (65 words) is conflict-free
__device__ __forceinline__ uint32_t& pa_sk(uint32_t* sk, uint32_t p) { return sk[p + (p >> 6)]; }

struct PaMasks {
  uint64_t left;    // bit i: entry r0 + i is the even-numbered entry of its run -> it leaves one output
  uint64_t add;     // bit i: ... and entry r0 + i + 1 belongs to the same run -> the output is their sum
};

// Loads the tile's keys and works out which entries pair up.  Must be called by all threads of the block.
__device__ __forceinline__ PaMasks pa_pairing(const uint32_t* __restrict__ keys_in, size_t base, uint32_t n_tile, uint32_t* sk,
                                              int* warp_scratch) {
  for (uint32_t i = threadIdx.x; i <= (uint32_t)PA_TILE; i += PA_THREADS) pa_sk(sk, i) = i < n_tile ? keys_in[base + i] : PA_NO_KEY;
  __syncthreads();
  const uint32_t r0 = threadIdx.x * PA_EPT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t valid = 0, same = 0, head = 0;
  if (r0 < n_tile) {
    uint32_t prev = r0 ? pa_sk(sk, r0 - 1) : PA_NO_KEY;
    uint32_t cur = pa_sk(sk, r0);
#pragma unroll 8
    for (int i = 0; i < PA_EPT; i++) {
      uint32_t nxt = pa_sk(sk, r0 + i + 1);
      const uint64_t bit = 1ull << i;
      if (cur != PA_NO_KEY) {
        valid |= bit;
        if (cur != prev) head |= bit;
        if (nxt == cur) same |= bit;
      }
      prev = cur;
      cur = nxt;
    }
  }
  // start of the run that is open at r0: the last head before r0 (max-scan over the threads; entry 0 is a head)
  int last_head = head ? (int)r0 + 63 - __clzll((long long)head) : -1;
  int incl = last_head;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl = max(incl, o);
  }
  if (lane == 31) warp_scratch[warp] = incl;
  int before = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) before = -1;
  __syncthreads();
  for (int w = 0; w < warp; w++) before = max(before, warp_scratch[w]);
  __syncthreads();
  // parity of the position inside the run, entry by entry (registers only)
  PaMasks m{0, 0};
  uint32_t odd = before >= 0 ? ((r0 - (uint32_t)before) & 1u) : 0u;    // parity entry r0 would have if it is not a head
#pragma unroll 8
  for (int i = 0; i < PA_EPT; i++) {
    const uint64_t bit = 1ull << i;
    if (head & bit) odd = 0;
    if ((valid & bit) && !odd) {
      m.left |= bit;
      if (same & bit) m.add |= bit;
    }
    odd ^= 1u;
  }
  return m;
}

// tile_off = exclusive scan of tile_cnt over the tiles of the current list; *n_out = new list length
static __global__ void __launch_bounds__(1024) k_pa_scan(const uint32_t* __restrict__ n_ptr, const uint32_t* __restrict__ tile_cnt,
                                                  uint32_t* __restrict__ tile_off, uint32_t* __restrict__ n_out) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  const uint32_t n = *n_ptr;
  const uint32_t n_tiles = (uint32_t)(((size_t)n + PA_TILE - 1) / PA_TILE);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t start = 0; start < n_tiles; start += 1024) {
    uint32_t i = start + threadIdx.x;
    uint32_t v = i < n_tiles ? tile_cnt[i] : 0u, incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += o;
      }
      warp_sums[lane] = wi - w;        // exclusive
    }
    __syncthreads();
    uint32_t excl = carry + warp_sums[warp] + incl - v;
    if (i < n_tiles) tile_off[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = carry;
}

template <class F>
__device__ __forceinline__ Fe<F> fe_shfl(const Fe<F>& a, int src) {
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
  return r;
}

// where entry p of the tile lives
template <class CF, bool LEVEL0>
struct PaSource {
  const char* table;            // LEVEL0: fixed-base table
  const uint32_t* refs;         // LEVEL0: sorted references of this tile
  const char* pts;              // later levels: dense points of this tile
  __device__ __forceinline__ const char* addr(uint32_t p, bool& negate) const {
    if (LEVEL0) {
      uint32_t ref = refs[p];
      negate = (ref & REF_NEG) != 0;
      return table + (size_t)(ref & ~REF_NEG) * 64;
    }
    negate = false;
    return pts + (size_t)p * 64;
  }
  __device__ __forceinline__ Affine<CF> load(uint32_t p) const {
    bool negate;
    const char* a = addr(p, negate);
    Affine<CF> r = aff_load<CF>(a);
    if (LEVEL0 && negate && !fe_is_zero(r.y)) r.y = fe_neg(r.y);
    return r;
  }
  __device__ __forceinline__ Fe<CF> load_x(uint32_t p) const {
    bool negate;
    return fe_load<CF>(addr(p, negate));
  }
};

// a + b by the chord/tangent rule.  kind: 0 chord, 1 tangent (a == b), 2 result is b (a is the identity),
// 3 result is a (b is the identity), 4 result is the identity (a == -b).  den is the value whose inverse the
// rule needs (1 for the kinds that need none), so that every lane contributes a non-zero factor to the shared product.
enum { PA_CHORD = 0, PA_TANGENT = 1, PA_TAKE_B = 2, PA_TAKE_A = 3, PA_CANCEL = 4 };
template <class CF>
__device__ __forceinline__ int pa_classify(const Affine<CF>& a, const Affine<CF>& b, Fe<CF>& den) {
  den = fe_one<CF>();
  if (aff_is_identity(a)) return PA_TAKE_B;
  if (aff_is_identity(b)) return PA_TAKE_A;
  if (fe_eq(a.x, b.x)) {
    if (fe_eq(a.y, b.y)) {
      den = fe_dbl(a.y);
      return PA_TANGENT;
    }
    return PA_CANCEL;
  }
  den = fe_sub(b.x, a.x);
  return PA_CHORD;
}

// What a thread of the up pass leaves for the down pass.
struct PaMeta {
  uint64_t left, add;
  uint32_t out0;      // where the thread's first output goes in the new list
  uint32_t pad;
};

// ---- way up: pairing, output positions, and the running product of each thread's denominators.
// pref[(thread) * PA_MAX_ADDS + k] = product of the thread's denominators before its k-th addition; runs[thread] = all of them.
template <class CF, bool LEVEL0>
__global__ void __launch_bounds__(PA_THREADS) k_pa_up(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ refs_in,
                                                      const void* __restrict__ pts_in, const void* __restrict__ table,
                                                      const uint32_t* __restrict__ n_ptr, uint32_t* __restrict__ tile_cnt,
                                                      PaMeta* __restrict__ meta, void* __restrict__ pref, void* __restrict__ runs,
                                                      unsigned long long* __restrict__ add_counter) {
  __shared__ uint32_t sk[PA_SK_WORDS];
  __shared__ int scratch[PA_WARPS];
  const uint32_t n = *n_ptr;
  const size_t base = (size_t)blockIdx.x * PA_TILE;
  if (base >= n) return;
  const uint32_t n_tile = n - base < (size_t)PA_TILE ? (uint32_t)(n - base) : (uint32_t)PA_TILE;
  const PaMasks m = pa_pairing(keys_in, base, n_tile, sk, scratch);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t r0 = threadIdx.x * PA_EPT;
  const size_t gt = (size_t)blockIdx.x * PA_THREADS + threadIdx.x;

  // where this thread's outputs go: tile offset + exclusive scan of the per-thread output counts
  {
    int cnt = __popcll(m.left), incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    int before = incl - cnt;
    int outs = 0;
    for (int w = 0; w < PA_WARPS; w++) {
      if (w < warp) before += scratch[w];
      outs += scratch[w];
    }
    PaMeta mt;
    mt.left = m.left;
    mt.add = m.add;
    mt.out0 = (uint32_t)before;            // inside the tile; the down pass adds the tile's offset (k_pa_scan runs in between)
    mt.pad = 0;
    meta[gt] = mt;
    if (threadIdx.x == 0) {
      tile_cnt[blockIdx.x] = (uint32_t)outs;
      if (add_counter) atomicAdd(add_counter, (unsigned long long)(n_tile - (uint32_t)outs));   // one addition per entry removed
    }
  }

  PaSource<CF, LEVEL0> src;
  src.table = reinterpret_cast<const char*>(table);
  src.refs = refs_in + base;
  src.pts = reinterpret_cast<const char*>(pts_in) + base * 64;

  Fe<CF> run = fe_one<CF>();
  char* my_pref = reinterpret_cast<char*>(pref) + gt * (size_t)PA_MAX_ADDS * 32;
  uint64_t todo = m.add;
  int k = 0;
  while (todo) {
    // four additions at a time: their eight x loads are issued before the first product needs one
    uint32_t pos[4];
    Fe<CF> ax[4], bx[4];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (todo) {
        const int i = __ffsll((long long)todo) - 1;
        todo &= todo - 1;
        pos[j] = r0 + (uint32_t)i;
        ax[j] = src.load_x(pos[j]);
        bx[j] = src.load_x(pos[j] + 1);
        cnt = j + 1;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (j < cnt) {
        Fe<CF> den;
        if (fe_is_zero(ax[j]) || fe_is_zero(bx[j]) || fe_eq(ax[j], bx[j])) {     // rare: identity operand, tangent or cancellation
          Affine<CF> a = src.load(pos[j]), b = src.load(pos[j] + 1);
          pa_classify(a, b, den);
        } else {
          den = fe_sub(bx[j], ax[j]);
        }
        fe_store<CF>(my_pref + (size_t)k * 32, run);
        k++;
        run = fe_mul(run, den);
      }
    }
  }
  fe_store<CF>(reinterpret_cast<char*>(runs) + gt * 32, run);
}

// ---- inverses of all the threads' products: Montgomery's trick over PA_INV_GROUP consecutive products per thread,
// every lane running its own (branch-uniform) inversion, so the SIMD width is used and an inversion is shared by
// PA_INV_GROUP * PA_MAX_ADDS additions.
constexpr int PA_INV_GROUP = 32;
template <class CF>
__global__ void __launch_bounds__(128) k_pa_inv(const void* __restrict__ runs, const uint32_t* __restrict__ n_ptr, void* __restrict__ invs) {
  const uint32_t n = *n_ptr;
  const size_t count = (((size_t)n + PA_TILE - 1) / PA_TILE) * PA_THREADS;       // products written by the up pass
  const size_t first = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * PA_INV_GROUP;
  if (first >= count) return;
  const int g_n = count - first < (size_t)PA_INV_GROUP ? (int)(count - first) : PA_INV_GROUP;
  const char* r = reinterpret_cast<const char*>(runs) + first * 32;
  char* o = reinterpret_cast<char*>(invs) + first * 32;
  Fe<CF> acc = fe_one<CF>();
  for (int g = 0; g < g_n; g++) {
    fe_store<CF>(o + (size_t)g * 32, acc);                       // product of the group's earlier entries
    acc = fe_mul(acc, fe_load<CF>(r + (size_t)g * 32));
  }
  Fe<CF> inv = fe_inv_uniform(acc);
  for (int g = g_n - 1; g >= 0; g--) {
    Fe<CF> before = fe_load_plain<CF>(o + (size_t)g * 32);
    fe_store<CF>(o + (size_t)g * 32, fe_mul(inv, before));
    inv = fe_mul(inv, fe_load<CF>(r + (size_t)g * 32));
  }
}

// ---- way down: the additions themselves, outputs in descending order; `inv` is the inverse of the product of the
// denominators not yet used.
template <class CF, bool LEVEL0>
__global__ void __launch_bounds__(PA_THREADS) k_pa_down(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ refs_in,
                                                        const void* __restrict__ pts_in, const void* __restrict__ table,
                                                        const uint32_t* __restrict__ n_ptr, const uint32_t* __restrict__ tile_off,
                                                        const PaMeta* __restrict__ meta,
                                                        const void* __restrict__ pref, const void* __restrict__ invs,
                                                        uint32_t* __restrict__ keys_out, void* __restrict__ pts_out) {
  const uint32_t n = *n_ptr;
  const size_t base = (size_t)blockIdx.x * PA_TILE;
  if (base >= n) return;
  const size_t gt = (size_t)blockIdx.x * PA_THREADS + threadIdx.x;
  const PaMeta mt = meta[gt];
  if (!mt.left) return;
  const uint32_t r0 = threadIdx.x * PA_EPT;
  PaSource<CF, LEVEL0> src;
  src.table = reinterpret_cast<const char*>(table);
  src.refs = refs_in + base;
  src.pts = reinterpret_cast<const char*>(pts_in) + base * 64;
  const char* my_pref = reinterpret_cast<const char*>(pref) + gt * (size_t)PA_MAX_ADDS * 32;
  Fe<CF> inv = fe_load<CF>(reinterpret_cast<const char*>(invs) + gt * 32);
  uint64_t todo = mt.left;
  int k = __popcll(mt.add);
  uint32_t o = tile_off[blockIdx.x] + mt.out0 + (uint32_t)__popcll(mt.left);
  // operands of the entry being worked on; the next entry's are loaded before this one's arithmetic starts
  int i = 63 - __clzll((long long)todo);
  todo &= ~(1ull << i);
  Affine<CF> r = src.load(r0 + (uint32_t)i), b = r;
  bool is_add = (mt.add >> i) & 1ull;
  if (is_add) b = src.load(r0 + (uint32_t)i + 1);
  for (;;) {
    const uint32_t p = r0 + (uint32_t)i;
    const bool more = todo != 0;
    int i_next = 0;
    bool add_next = false;
    Affine<CF> r_next = r, b_next = b;
    if (more) {
      i_next = 63 - __clzll((long long)todo);
      todo &= ~(1ull << i_next);
      add_next = (mt.add >> i_next) & 1ull;
      r_next = src.load(r0 + (uint32_t)i_next);
      if (add_next) b_next = src.load(r0 + (uint32_t)i_next + 1);
    }
    o--;
    if (is_add) {
      k--;
      Fe<CF> den;
      const int kind = pa_classify(r, b, den);
      if (kind == PA_TAKE_B) {
        r = b;
      } else if (kind == PA_CANCEL) {
        r.x = fe_zero<CF>();
        r.y = fe_zero<CF>();
      } else if (kind != PA_TAKE_A) {
        Fe<CF> inv_den = fe_mul(inv, fe_load<CF>(my_pref + (size_t)k * 32));
        inv = fe_mul(inv, den);
        Fe<CF> num;
        if (kind == PA_TANGENT) {
          Fe<CF> xx = fe_sqr(r.x);
          num = fe_add(fe_dbl(xx), xx);
        } else {
          num = fe_sub(b.y, r.y);
        }
        Fe<CF> lam = fe_mul(num, inv_den);
        Fe<CF> x3 = fe_sub(fe_sub(fe_sqr(lam), r.x), b.x);
        r.y = fe_sub(fe_mul(lam, fe_sub(r.x, x3)), r.y);
        r.x = x3;
      }
    }
    keys_out[o] = keys_in[base + p];
    aff_store<CF>(reinterpret_cast<char*>(pts_out) + (size_t)o * 64, r);
    if (!more) break;
    i = i_next;
    is_add = add_next;
    r = r_next;
    b = b_next;
  }
}

}  // namespace mira
