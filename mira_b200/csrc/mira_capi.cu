// mira_capi.cu — host orchestration + C ABI (include/mira_b200.h) of the B200 commitment engine.
//
// Drop-in for CommitmentKey::commit (/root/reference/src/commitment.rs:78-87).  No CPU fallback:
// without a usable CUDA device every entry point returns MIRA_ERR_CUDA.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mira_b200.h"
#include "msm_kernels.cuh"
#include "testgen.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(x)                                                                                          \
  do {                                                                                                 \
    cudaError_t e_ = (x);                                                                              \
    if (e_ != cudaSuccess)                                                                             \
      return fail(MIRA_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
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

struct Table {
  int c = 0, W = 0;
  uint32_t n_cover = 0;
  void* d = nullptr;
};

int windows_for(int c) { return (255 + c - 1) / c; }

// Cost model (in field multiplications) used to pick the window width for a commit of length n:
// n*W mixed adds (10 muls) + 2 * 2^(c-1) full adds (14 muls) for the running sums, the latter
// weighted up because the reduction runs at lower occupancy.
int choose_window(size_t n) {
  int best = 8;
  double best_cost = 1e300;
  for (int c = 6; c <= 24; c++) {
    double W = windows_for(c);
    double cost = (double)n * W * 10.0 + 2.0 * std::ldexp(1.0, c - 1) * 14.0 * 1.5 + 40000.0 * 14.0;
    if (cost < best_cost) {
      best_cost = cost;
      best = c;
    }
  }
  return best;
}

}  // namespace

struct mira_msm_ctx {
  int curve = 0;
  int device = 0;
  size_t n_bases = 0;
  void* d_bases = nullptr;
  cudaStream_t stream = nullptr;
  std::vector<Table> tables;
  // workspace
  DevBuf scalars, keys, refs, skeys, srefs, counts, cursor, tile_sums, buckets, part_keys, part_pts, red_a, red_b, result;
  void* h_result = nullptr;  // pinned, 128 B
  int forced_window = 0;
  bool profiling = false;
  mira_msm_stats stats{};
  std::mutex mu;
};

namespace {

using namespace mira;

template <class CF>
int build_table(mira_msm_ctx* ctx, int c, uint32_t n_cover, Table* out) {
  int W = windows_for(c);
  void* d = nullptr;
  size_t bytes = (size_t)W * n_cover * 64;
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e != cudaSuccess)
    return fail(MIRA_ERR_CUDA, "cudaMalloc(%zu) for the fixed-base table failed: %s", bytes, cudaGetErrorString(e));
  k_precompute<CF><<<(n_cover + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_bases, n_cover, c, W, d);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(MIRA_ERR_CUDA, "k_precompute launch failed: %s", cudaGetErrorString(e));
  }
  out->c = c;
  out->W = W;
  out->n_cover = n_cover;
  out->d = d;
  return MIRA_OK;
}

// returns the table for window c covering at least n points (building it on first use)
template <class CF>
int get_table(mira_msm_ctx* ctx, int c, size_t n, Table** out) {
  for (auto& t : ctx->tables)
    if (t.c == c && t.n_cover >= n) {
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
  Table t;
  int rc = build_table<CF>(ctx, c, (uint32_t)cover, &t);
  if (rc) return rc;
  ctx->tables.push_back(t);
  *out = &ctx->tables.back();
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

// Runs the device pipeline; leaves the XYZZ sum (128 B) in ctx->result.p
template <class CF, class SF>
int msm_device(mira_msm_ctx* ctx, const void* d_scalars, size_t n, cudaStream_t st) {
  uint64_t launches = 0;
  int rc;
  if ((rc = ctx->result.ensure(256))) return rc;
  if (n == 0) {
    CU(cudaMemsetAsync(ctx->result.p, 0, 128, st));
    ctx->stats = mira_msm_stats{};
    return MIRA_OK;
  }
  int c = ctx->forced_window ? ctx->forced_window : choose_window(n);
  Table* tab = nullptr;
  if ((rc = get_table<CF>(ctx, c, n, &tab))) return rc;
  const int W = tab->W;
  const size_t E = n * (size_t)W;
  if (E >= (size_t)0x7fffffff || (size_t)W * tab->n_cover >= (size_t)0x7fffffff)
    return fail(MIRA_ERR_INVALID, "commit of %zu scalars needs %zu (point, window) pairs: exceeds the 2^31 reference space; shard it", n, E);
  const uint32_t B = 1u << (c - 1);

  if ((rc = ctx->keys.ensure(E * 4)) || (rc = ctx->refs.ensure(E * 4)) || (rc = ctx->skeys.ensure(E * 4 + 16)) ||
      (rc = ctx->srefs.ensure(E * 4)) || (rc = ctx->counts.ensure(((size_t)B + 2) * 4)) ||
      (rc = ctx->cursor.ensure(((size_t)B + 2) * 4)) || (rc = ctx->buckets.ensure(((size_t)B + 1) * 128)))
    return rc;
  const uint32_t n_counts = B + 1;
  const uint32_t n_tiles = (n_counts + SCAN_TILE - 1) / SCAN_TILE;
  if ((rc = ctx->tile_sums.ensure((size_t)n_tiles * 4 + 16))) return rc;

  PhaseTimer pt(ctx->profiling, st);
  pt.mark(0);
  // ---- digits + histogram
  CU(cudaMemsetAsync(ctx->counts.p, 0, ((size_t)B + 2) * 4, st));
  k_digits<SF><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_scalars, (uint32_t)n, c, W, tab->n_cover, (uint32_t*)ctx->keys.p,
                                                           (uint32_t*)ctx->refs.p, (uint32_t*)ctx->counts.p);
  launches++;
  pt.mark(1);
  // ---- exclusive scan of the histogram -> bucket start offsets, then scatter
  k_scan_tile_sums<<<n_tiles, SCAN_THREADS, 0, st>>>((const uint32_t*)ctx->counts.p, n_counts, (uint32_t*)ctx->tile_sums.p);
  k_scan_small<<<1, SCAN_THREADS, 0, st>>>((uint32_t*)ctx->tile_sums.p, n_tiles + 1);
  k_scan_apply<<<n_tiles, SCAN_THREADS, 0, st>>>((const uint32_t*)ctx->counts.p, n_counts, (const uint32_t*)ctx->tile_sums.p,
                                                 (uint32_t*)ctx->cursor.p);
  k_scatter<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((const uint32_t*)ctx->keys.p, (const uint32_t*)ctx->refs.p, E,
                                                        (uint32_t*)ctx->cursor.p, (uint32_t*)ctx->skeys.p, (uint32_t*)ctx->srefs.p);
  launches += 4;
  // number of non-zero entries = tile_sums[n_tiles] after the scan (total); read it back
  uint32_t n_sorted = 0;
  CU(cudaMemcpyAsync(&n_sorted, (uint32_t*)ctx->tile_sums.p + n_tiles, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  pt.mark(2);
  // ---- accumulate
  CU(cudaMemsetAsync(ctx->buckets.p, 0, ((size_t)B + 1) * 128, st));
  if (n_sorted) {
    const int L = 32;
    uint32_t n_chunks = (n_sorted + L - 1) / L;
    if ((rc = ctx->part_keys.ensure((size_t)n_chunks * 8)) || (rc = ctx->part_pts.ensure((size_t)n_chunks * 256))) return rc;
    k_accumulate<CF><<<(n_chunks + 127) / 128, 128, 0, st>>>((const uint32_t*)ctx->skeys.p, (const uint32_t*)ctx->srefs.p, n_sorted, L,
                                                            tab->d, ctx->buckets.p, (uint32_t*)ctx->part_keys.p, ctx->part_pts.p);
    k_combine<CF><<<(2 * n_chunks + 127) / 128, 128, 0, st>>>((const uint32_t*)ctx->part_keys.p, ctx->part_pts.p, n_chunks, ctx->buckets.p);
    launches += 2;
  }
  pt.mark(3);
  // ---- bucket reduction
  const uint32_t m = B >= (1u << 14) ? 32 : (B >= 1024 ? 8 : 1);
  uint32_t n_red = (B + m - 1) / m;
  if ((rc = ctx->red_a.ensure((size_t)n_red * 128)) || (rc = ctx->red_b.ensure((size_t)(n_red / 128 + 2) * 128))) return rc;
  k_reduce_chunks<CF><<<(n_red + 127) / 128, 128, 0, st>>>(ctx->buckets.p, B, m, ctx->red_a.p);
  launches++;
  void* src = ctx->red_a.p;
  void* dst = ctx->red_b.p;
  uint32_t cnt = n_red;
  while (cnt > 1) {
    uint32_t per_thread = cnt > 128 * 8 ? 8 : 1;
    uint32_t per_block = per_thread * 128;
    uint32_t blocks = (cnt + per_block - 1) / per_block;
    k_sum_points<CF><<<blocks, 128, 0, st>>>(src, cnt, per_thread, dst);
    launches++;
    cnt = blocks;
    std::swap(src, dst);
  }
  CU(cudaMemcpyAsync(ctx->result.p, src, 128, cudaMemcpyDeviceToDevice, st));
  pt.mark(4);
  CU(cudaGetLastError());
  ctx->stats.window_bits = c;
  ctx->stats.windows = W;
  ctx->stats.entries = E;
  ctx->stats.buckets = B;
  ctx->stats.kernel_launches = launches;
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

template <class CF, class SF>
int commit_impl(mira_msm_ctx* ctx, const void* scalars, size_t n, int on_device, void* out, bool want_affine, cudaStream_t st) {
  int rc;
  const void* d_scalars = scalars;
  if (!on_device && n) {
    if ((rc = ctx->scalars.ensure(n * 32))) return rc;
    CU(cudaMemcpyAsync(ctx->scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
    d_scalars = ctx->scalars.p;
  }
  if ((rc = msm_device<CF, SF>(ctx, d_scalars, n, st))) return rc;
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

int dispatch_commit(mira_msm_ctx* ctx, const void* scalars, size_t n, int on_device, void* out, bool want_affine, void* stream) {
  if (!ctx || !out || (n && !scalars)) return fail(MIRA_ERR_INVALID, "null argument");
  if (n > ctx->n_bases)   // src/commitment.rs:79-86: checked before any arithmetic
    return fail(MIRA_ERR_TOO_LONG_INPUT, "Can't commit too long input: input len: %zu, but limit is %zu", n, ctx->n_bases);
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  if (ctx->curve == MIRA_BN254_G1) return commit_impl<FqTag, FrTag>(ctx, scalars, n, on_device, out, want_affine, st);
  return commit_impl<FrTag, FqTag>(ctx, scalars, n, on_device, out, want_affine, st);
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

const char* mira_last_error(void) { return g_err.c_str(); }

int mira_msm_ctx_create(int curve, const void* bases, size_t n_bases, int bases_on_device, int device, mira_msm_ctx** out) {
  if (!out) return fail(MIRA_ERR_INVALID, "out is null");
  *out = nullptr;
  if (curve != MIRA_BN254_G1 && curve != MIRA_GRUMPKIN_G1) return fail(MIRA_ERR_INVALID, "unknown curve %d", curve);
  if (n_bases && !bases) return fail(MIRA_ERR_INVALID, "bases is null");
  if (n_bases >= ((size_t)1 << 31)) return fail(MIRA_ERR_INVALID, "key of %zu points is too large for one device context", n_bases);
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(MIRA_ERR_CUDA, "CUDA device %d not available (%d visible)", device, count);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(MIRA_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only (no fallback path)", device, prop.major, prop.minor);
  auto* ctx = new mira_msm_ctx();
  ctx->curve = curve;
  ctx->device = device;
  ctx->n_bases = n_bases;
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_bases, n_bases ? n_bases * 64 : 64);
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_result, 256);
  if (e == cudaSuccess && n_bases)
    e = cudaMemcpyAsync(ctx->d_bases, bases, n_bases * 64, bases_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    int rc = fail(MIRA_ERR_CUDA, "context creation failed: %s", cudaGetErrorString(e));
    mira_msm_ctx_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return MIRA_OK;
}

void mira_msm_ctx_destroy(mira_msm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& t : ctx->tables) cudaFree(t.d);
  for (DevBuf* b : {&ctx->scalars, &ctx->keys, &ctx->refs, &ctx->skeys, &ctx->srefs, &ctx->counts, &ctx->cursor, &ctx->tile_sums,
                    &ctx->buckets, &ctx->part_keys, &ctx->part_pts, &ctx->red_a, &ctx->red_b, &ctx->result})
    b->release();
  if (ctx->d_bases) cudaFree(ctx->d_bases);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

size_t mira_msm_ctx_len(const mira_msm_ctx* ctx) { return ctx ? ctx->n_bases : 0; }

int mira_msm_ctx_check_on_curve(mira_msm_ctx* ctx) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ctx->result.ensure(256))) return rc;
  uint32_t* flag = (uint32_t*)((char*)ctx->result.p + 192);
  CU(cudaMemsetAsync(flag, 0, 4, ctx->stream));
  if (ctx->n_bases) {
    unsigned blocks = (unsigned)((ctx->n_bases + 255) / 256);
    if (ctx->curve == MIRA_BN254_G1)
      mira::k_check_on_curve<mira::FqTag><<<blocks, 256, 0, ctx->stream>>>(ctx->d_bases, (uint32_t)ctx->n_bases, 3u, 0, flag);
    else
      mira::k_check_on_curve<mira::FrTag><<<blocks, 256, 0, ctx->stream>>>(ctx->d_bases, (uint32_t)ctx->n_bases, 17u, 1, flag);
  }
  uint32_t h = 0;
  CU(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (h) return fail(MIRA_ERR_NOT_ON_CURVE, "Wrong file in cache, some ptr out of curve");
  return MIRA_OK;
}

int mira_msm_ctx_prepare(mira_msm_ctx* ctx, size_t n) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  if (n > ctx->n_bases) return fail(MIRA_ERR_TOO_LONG_INPUT, "Can't commit too long input: input len: %zu, but limit is %zu", n, ctx->n_bases);
  if (n == 0) return MIRA_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  int c = ctx->forced_window ? ctx->forced_window : choose_window(n);
  Table* t = nullptr;
  int rc = ctx->curve == MIRA_BN254_G1 ? get_table<mira::FqTag>(ctx, c, n, &t) : get_table<mira::FrTag>(ctx, c, n, &t);
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MIRA_OK;
}

int mira_msm_commit(mira_msm_ctx* ctx, const void* scalars, size_t n, void* out_affine) {
  return dispatch_commit(ctx, scalars, n, 0, out_affine, true, nullptr);
}
int mira_msm_commit_device(mira_msm_ctx* ctx, const void* scalars_dev, size_t n, void* out_affine, void* stream) {
  return dispatch_commit(ctx, scalars_dev, n, 1, out_affine, true, stream);
}
int mira_msm_partial(mira_msm_ctx* ctx, const void* scalars, size_t n, int scalars_on_device, void* out_xyzz, void* stream) {
  return dispatch_commit(ctx, scalars, n, scalars_on_device, out_xyzz, false, stream);
}

int mira_msm_combine(int curve, const void* partials, size_t count, int device, void* out_affine) {
  if (!out_affine || (count && !partials)) return fail(MIRA_ERR_INVALID, "null argument");
  if (curve != MIRA_BN254_G1 && curve != MIRA_GRUMPKIN_G1) return fail(MIRA_ERR_INVALID, "unknown curve %d", curve);
  CU(cudaSetDevice(device));
  void* d = nullptr;
  size_t bytes = (count + 2) * 128;
  CU(cudaMalloc(&d, bytes * 2));
  CU(cudaMemset(d, 0, bytes * 2));
  if (count) CU(cudaMemcpy(d, partials, count * 128, cudaMemcpyHostToDevice));
  void* src = d;
  void* dst = (char*)d + bytes;
  uint32_t cnt = count ? (uint32_t)count : 1;
  while (cnt > 1) {
    uint32_t blocks = (cnt + 127) / 128;
    if (curve == MIRA_BN254_G1) mira::k_sum_points<mira::FqTag><<<blocks, 128>>>(src, cnt, 1, dst);
    else mira::k_sum_points<mira::FrTag><<<blocks, 128>>>(src, cnt, 1, dst);
    cnt = blocks;
    std::swap(src, dst);
  }
  if (curve == MIRA_BN254_G1) mira::k_finalize<mira::FqTag><<<1, 32>>>(src, (char*)dst);
  else mira::k_finalize<mira::FrTag><<<1, 32>>>(src, (char*)dst);
  cudaError_t e = cudaMemcpy(out_affine, dst, 64, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "combine failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

int mira_msm_get_stats(const mira_msm_ctx* ctx, mira_msm_stats* out) {
  if (!ctx || !out) return fail(MIRA_ERR_INVALID, "null argument");
  *out = ctx->stats;
  return MIRA_OK;
}
int mira_msm_set_profiling(mira_msm_ctx* ctx, int enabled) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  ctx->profiling = enabled != 0;
  return MIRA_OK;
}
int mira_msm_set_window(mira_msm_ctx* ctx, int window_bits) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  if (window_bits != 0 && (window_bits < 2 || window_bits > 26)) return fail(MIRA_ERR_INVALID, "window must be 0 or in [2, 26]");
  ctx->forced_window = window_bits;
  return MIRA_OK;
}

// ---------------------------------------------------------------------- generators / test hooks
int mira_gen_scalars(int curve, uint64_t seed, size_t first, size_t n, int dist, int device, void* out_dev) {
  if (n && !out_dev) return fail(MIRA_ERR_INVALID, "null argument");
  CU(cudaSetDevice(device));
  if (!n) return MIRA_OK;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (curve == MIRA_BN254_G1) mira::k_gen_scalars<mira::FrTag><<<blocks, 256>>>(seed, first, n, dist, out_dev);
  else mira::k_gen_scalars<mira::FqTag><<<blocks, 256>>>(seed, first, n, dist, out_dev);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return MIRA_OK;
}

int mira_gen_bases(int curve, uint64_t seed, size_t first, size_t n, int device, void* out_dev) {
  if (n && !out_dev) return fail(MIRA_ERR_INVALID, "null argument");
  CU(cudaSetDevice(device));
  if (!n) return MIRA_OK;
  void* table = nullptr;
  CU(cudaMalloc(&table, 32 * 256 * 64));
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (curve == MIRA_BN254_G1) {
    mira::k_gen_table<mira::FqTag><<<32 * 256 / 128, 128>>>(table);
    mira::k_gen_bases<mira::FqTag, mira::FrTag><<<blocks, 128>>>(seed, first, n, table, out_dev);
  } else {
    mira::k_gen_table<mira::FrTag><<<32 * 256 / 128, 128>>>(table);
    mira::k_gen_bases<mira::FrTag, mira::FqTag><<<blocks, 128>>>(seed, first, n, table, out_dev);
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(table);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "gen_bases failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

int mira_test_field_op(int field, int op, const void* a, const void* b, size_t n, int device, void* out) {
  if (!a || !out || n == 0) return fail(MIRA_ERR_INVALID, "null argument");
  CU(cudaSetDevice(device));
  void *da = nullptr, *db = nullptr, *dout = nullptr;
  CU(cudaMalloc(&da, n * 32));
  CU(cudaMalloc(&db, n * 32));
  CU(cudaMalloc(&dout, n * 32));
  CU(cudaMemcpy(da, a, n * 32, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(db, b ? b : a, n * 32, cudaMemcpyHostToDevice));
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (field == MIRA_FQ) mira::k_test_field<mira::FqTag><<<blocks, 128>>>(op, da, db, n, dout);
  else mira::k_test_field<mira::FrTag><<<blocks, 128>>>(op, da, db, n, dout);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out, dout, n * 32, cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(db); cudaFree(dout);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "test_field_op failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

int mira_test_point_op(int curve, int op, const void* p, const void* q, size_t n, int device, void* out) {
  if (!p || !out || n == 0) return fail(MIRA_ERR_INVALID, "null argument");
  CU(cudaSetDevice(device));
  void *dp = nullptr, *dq = nullptr, *dout = nullptr;
  CU(cudaMalloc(&dp, n * 64));
  CU(cudaMalloc(&dq, n * 64));
  CU(cudaMalloc(&dout, n * 64));
  CU(cudaMemcpy(dp, p, n * 64, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dq, q ? q : p, n * 64, cudaMemcpyHostToDevice));
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (curve == MIRA_BN254_G1) mira::k_test_point<mira::FqTag><<<blocks, 128>>>(op, dp, dq, n, dout);
  else mira::k_test_point<mira::FrTag><<<blocks, 128>>>(op, dp, dq, n, dout);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out, dout, n * 64, cudaMemcpyDeviceToHost);
  cudaFree(dp); cudaFree(dq); cudaFree(dout);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "test_point_op failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

}  // extern "C"
