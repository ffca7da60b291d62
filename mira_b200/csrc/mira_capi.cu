// mira_capi.cu — the C ABI (include/mira_b200.h) of the B200 commitment engine.
//
// Drop-in for CommitmentKey::commit (/root/reference/src/commitment.rs:78-87).  No CPU fallback:
// without a usable sm_100 CUDA device every compute entry point returns MIRA_ERR_CUDA.
// Per-curve device code lives in curve_bn254.cu / curve_grumpkin.cu (templates in pipeline.cuh).
#include "ctx.hpp"

#include <algorithm>
#include <thread>

namespace mira_host {

static thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

// Cost model (in field multiplications) used to pick the window width for a commit of length n:
// n*W mixed adds (10 muls each) + 2 * 2^(c-1) full adds (14 muls) for the running sums — weighted up
// because the reduction runs at lower occupancy — + a fixed launch/latency term.
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

// The same model fed with the pair counts a SAMPLE of the scalars predicts: hist[L] = sampled scalars whose canonical
// value has bit length L (hist[0] = zeros), scale = n / samples.  A scalar of bit length L contributes about
// ceil(L / c) non-zero signed digits.  Witness vectors (60 % zero, most of the rest below 2^32) produce an order of
// magnitude fewer pairs than uniform scalars, and the bucket reduction, whose cost does not shrink with them, then
// asks for a much narrower window (2^23 witness-like scalars: c = 22 -> 5.9 ms, c = 16 -> 4.1 ms).
int choose_window_sampled(size_t n, const uint32_t* hist, size_t samples, double* costs) {
  if (!samples) return choose_window(n);
  const double scale = (double)n / (double)samples;
  int best = 8;
  double best_cost = 1e300;
  for (int c = 6; c <= 24; c++) {
    double pairs = 0;
    for (int L = 1; L <= 256; L++)
      if (hist[L]) pairs += (double)hist[L] * (double)((L + c - 1) / c);
    pairs *= scale;
    double cost = pairs * 10.0 + 2.0 * std::ldexp(1.0, c - 1) * 14.0 * 1.5 + 40000.0 * 14.0;
    if (costs) costs[c] = cost;
    if (cost < best_cost) {
      best_cost = cost;
      best = c;
    }
  }
  return best;
}

static int valid_curve(int curve) { return curve == MIRA_BN254_G1 || curve == MIRA_GRUMPKIN_G1; }

// Stream convention (include/mira_b200.h): entry points that read DEVICE scalars run on the caller's stream, and NULL
// there is the legacy default stream, exactly as for the witness kernels (mira_fold_w, mira_eval_rows, ...), so that
// `mira_fold_w(..., NULL)` followed by `mira_msm_commit_device(..., NULL)` is ordered.  Host-buffer commits have no
// device-side producer and run on the context's own non-blocking stream.
static cudaStream_t stream_for(mira_msm_ctx* ctx, int scalars_on_device, void* stream) {
  return scalars_on_device ? (cudaStream_t)stream : (stream ? (cudaStream_t)stream : ctx->stream);
}

// ---- single-process multi-GPU key (mira_msm_ctx_create_sharded) ------------------------------------------------
// CommitmentKey::commit uses the key prefix ck[..v.len()] (src/commitment.rs:80): shard g, which owns key indices
// [lo, hi), commits v[lo .. min(hi, n)) -- possibly nothing.  One host thread per device drives that shard's ordinary
// host-buffer pipeline (its own PCIe link, its own streams); the 128-byte XYZZ partial sums land in `gather` on the
// parent's device by peer copy, and one kernel there adds them and normalises.  No collective library is needed inside
// one process; across processes the same partials travel by an NCCL all_gather (mira_b200/sharding.py, bench.py).
static int sharded_commit(mira_msm_ctx* ctx, const void* scalars, size_t n, void* out_affine) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  const size_t G = ctx->shards.size();
  CU(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ctx->gather.ensure(G * 128))) return rc;
  std::vector<int> rcs(G, MIRA_OK);
  std::vector<std::string> errs(G);
  std::vector<std::thread> workers;
  workers.reserve(G);
  auto run_shard = [&](size_t g, size_t lo, size_t cnt) {
    mira_msm_ctx* sub = ctx->shards[g];
    std::lock_guard<std::mutex> l2(sub->mu);
    cudaError_t e = cudaSetDevice(sub->device);
    if (e != cudaSuccess) {
      rcs[g] = MIRA_ERR_CUDA;
      errs[g] = cudaGetErrorString(e);
      return;
    }
    rcs[g] = ops_for(sub->curve).partial_to_peer(sub, (const char*)scalars + lo * 32, cnt, (char*)ctx->gather.p + g * 128, ctx->device,
                                                  sub->stream);
    if (rcs[g]) errs[g] = g_err;       // the error text is thread-local: carry it to the calling thread
  };
  for (size_t g = 0; g < G; g++) {
    const size_t lo = ctx->shard_lo[g], hi = ctx->shard_lo[g + 1];
    const size_t cnt = n > lo ? std::min(n, hi) - lo : 0;
    if (g + 1 == G) {
      run_shard(g, lo, cnt);             // the last range on the calling thread
      break;
    }
    try {
      workers.emplace_back(run_shard, g, lo, cnt);
    } catch (const std::exception&) {    // no thread to be had: the range still has to be committed
      run_shard(g, lo, cnt);
    }
  }
  for (auto& w : workers) w.join();
  CU(cudaSetDevice(ctx->device));      // the calling thread ran the last range on that range's device
  for (size_t g = 0; g < G; g++)
    if (rcs[g]) return fail(rcs[g], "shard %zu (device %d): %s", g, ctx->shards[g]->device, errs[g].c_str());
  if ((rc = ops_for(ctx->curve).combine_dev(ctx->gather.p, G, 1, 128, out_affine, ctx->stream))) return rc;
  // stats of the whole commit: pairs and launches summed over the shards, phase times of the slowest
  mira_msm_stats agg{};
  for (size_t g = 0; g < G; g++) {
    const mira_msm_stats& s = ctx->shards[g]->stats;
    if (g == 0) { agg.window_bits = s.window_bits; agg.windows = s.windows; agg.buckets = s.buckets; }
    agg.entries += s.entries;
    agg.kernel_launches += s.kernel_launches;
    agg.ms_total = std::max(agg.ms_total, s.ms_total);
  }
  agg.kernel_launches += 1;
  ctx->stats = agg;
  return MIRA_OK;
}

static int dispatch_commit(mira_msm_ctx* ctx, const void* scalars, size_t n, int on_device, void* out, bool want_affine, void* stream) {
  if (!ctx || !out || (n && !scalars)) return fail(MIRA_ERR_INVALID, "null argument");
  if (n > ctx->n_bases)   // src/commitment.rs:79-86: checked before any arithmetic
    return fail(MIRA_ERR_TOO_LONG_INPUT, "Can't commit too long input: input len: %zu, but limit is %zu", n, ctx->n_bases);
  if (!ctx->shards.empty()) {
    if (on_device || !want_affine)
      return fail(MIRA_ERR_INVALID, "a sharded (multi-device) context commits HOST scalar vectors to affine results only: device-resident "
                                    "vectors live on one device; use one context per device and mira_msm_partial_batch_dev / mira_msm_combine_dev");
    return sharded_commit(ctx, scalars, n, out);
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  return ops_for(ctx->curve).commit(ctx, scalars, n, on_device, out, want_affine, stream_for(ctx, on_device, stream));
}

// applies `fn` to every shard of a sharded context (first failure wins); false if the context is not sharded
template <class Fn>
static bool for_each_shard(mira_msm_ctx* ctx, int* rc, Fn fn) {
  if (ctx->shards.empty()) return false;
  *rc = MIRA_OK;
  for (mira_msm_ctx* sub : ctx->shards)
    if ((*rc = fn(sub)) != MIRA_OK) break;
  return true;
}

}  // namespace mira_host

using namespace mira_host;

extern "C" {

const char* mira_last_error(void) { return g_err.c_str(); }

int mira_msm_ctx_create(int curve, const void* bases, size_t n_bases, int bases_on_device, int device, mira_msm_ctx** out) {
  if (!out) return fail(MIRA_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!valid_curve(curve)) return fail(MIRA_ERR_INVALID, "unknown curve %d", curve);
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
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_result, 4096);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&ctx->h_hist, 2048 * 4);
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

int mira_msm_ctx_create_sharded(int curve, const void* bases, size_t n_bases, const int* devices, size_t n_devices, mira_msm_ctx** out) {
  if (!out) return fail(MIRA_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!valid_curve(curve)) return fail(MIRA_ERR_INVALID, "unknown curve %d", curve);
  if (!devices || n_devices == 0 || n_devices > 64) return fail(MIRA_ERR_INVALID, "a sharded context needs 1..64 devices");
  if (n_bases && !bases) return fail(MIRA_ERR_INVALID, "bases is null");
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  for (size_t g = 0; g < n_devices; g++)
    if (devices[g] < 0 || devices[g] >= count) return fail(MIRA_ERR_CUDA, "CUDA device %d not available (%d visible)", devices[g], count);
  auto* ctx = new mira_msm_ctx();
  ctx->curve = curve;
  ctx->device = devices[0];
  ctx->n_bases = n_bases;
  // balanced contiguous ranges, the first n % G shards one point longer (mira_b200/sharding.py: shard_range)
  const size_t q = n_bases / n_devices, r = n_bases % n_devices;
  ctx->shard_lo.resize(n_devices + 1);
  for (size_t g = 0; g <= n_devices; g++) ctx->shard_lo[g] = g * q + std::min(g, r);
  int rc = MIRA_OK;
  for (size_t g = 0; g < n_devices && rc == MIRA_OK; g++) {
    mira_msm_ctx* sub = nullptr;
    const size_t lo = ctx->shard_lo[g], cnt = ctx->shard_lo[g + 1] - lo;
    rc = mira_msm_ctx_create(curve, (const char*)bases + lo * 64, cnt, 0, devices[g], &sub);
    if (rc == MIRA_OK) {
      ctx->shards.push_back(sub);
      if (devices[g] != devices[0]) {       // direct peer copies of the partial sums; without peer access the driver stages them
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[g], devices[0]) == cudaSuccess && can) {
          cudaSetDevice(devices[g]);
          cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
          if (e != cudaSuccess) cudaGetLastError();       // already enabled, or not supported: not an error
        }
      }
    }
  }
  if (rc == MIRA_OK) {
    cudaError_t e = cudaSetDevice(devices[0]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) rc = fail(MIRA_ERR_CUDA, "sharded context creation failed: %s", cudaGetErrorString(e));
  }
  if (rc != MIRA_OK) {
    std::string keep = g_err;
    mira_msm_ctx_destroy(ctx);
    g_err = keep;
    return rc;
  }
  *out = ctx;
  return MIRA_OK;
}

size_t mira_msm_ctx_num_devices(const mira_msm_ctx* ctx) { return ctx ? (ctx->shards.empty() ? 1 : ctx->shards.size()) : 0; }

void mira_msm_ctx_destroy(mira_msm_ctx* ctx) {
  if (!ctx) return;
  for (mira_msm_ctx* sub : ctx->shards) mira_msm_ctx_destroy(sub);
  ctx->shards.clear();
  cudaSetDevice(ctx->device);
  ctx->gather.release();
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& t : ctx->tables) cudaFree(t.d);
  for (DevBuf* b : {&ctx->scalars, &ctx->cursor, &ctx->buckets, &ctx->part_keys, &ctx->part_pts, &ctx->red_a, &ctx->red_b, &ctx->red_c, &ctx->red_d, &ctx->result,
                    &ctx->pa_a, &ctx->pa_b, &ctx->pa_work})
    b->release();
  for (auto& sb : ctx->sb)
    for (DevBuf* b : {&sb.keys, &sb.refs, &sb.skeys, &sb.srefs, &sb.counts, &sb.tile_sums}) b->release();
  for (cudaEvent_t e : {ctx->prep_done[0], ctx->prep_done[1], ctx->acc_done[0], ctx->acc_done[1], ctx->pipe_start})
    if (e) cudaEventDestroy(e);
  if (ctx->prep_stream) cudaStreamDestroy(ctx->prep_stream);
  if (ctx->d_bases) cudaFree(ctx->d_bases);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  if (ctx->h_hist) cudaFreeHost(ctx->h_hist);
  ctx->sample.release();
  for (auto& e : ctx->copy_done)
    if (e) cudaEventDestroy(e);
  if (ctx->compute_idle) cudaEventDestroy(ctx->compute_idle);
  ctx->stager.release();
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

size_t mira_msm_ctx_len(const mira_msm_ctx* ctx) { return ctx ? ctx->n_bases : 0; }

int mira_msm_ctx_check_on_curve(mira_msm_ctx* ctx) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  int frc;
  if (for_each_shard(ctx, &frc, [](mira_msm_ctx* s) { return mira_msm_ctx_check_on_curve(s); })) return frc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  return ops_for(ctx->curve).check_on_curve(ctx);
}

int mira_msm_ctx_prepare(mira_msm_ctx* ctx, size_t n) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  if (n > ctx->n_bases) return fail(MIRA_ERR_TOO_LONG_INPUT, "Can't commit too long input: input len: %zu, but limit is %zu", n, ctx->n_bases);
  if (n == 0) return MIRA_OK;
  if (!ctx->shards.empty()) {           // every shard prepares for its share of a commit of length n
    for (size_t g = 0; g < ctx->shards.size(); g++) {
      const size_t lo = ctx->shard_lo[g], hi = ctx->shard_lo[g + 1];
      int rc = n > lo ? mira_msm_ctx_prepare(ctx->shards[g], std::min(n, hi) - lo) : MIRA_OK;
      if (rc) return rc;
    }
    return MIRA_OK;
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  return ops_for(ctx->curve).prepare(ctx, n, nullptr, 0);
}

int mira_msm_ctx_prepare_for(mira_msm_ctx* ctx, const void* scalars, size_t n, int scalars_on_device) {
  if (!ctx || (n && !scalars)) return fail(MIRA_ERR_INVALID, "null argument");
  if (n > ctx->n_bases) return fail(MIRA_ERR_TOO_LONG_INPUT, "Can't commit too long input: input len: %zu, but limit is %zu", n, ctx->n_bases);
  if (n == 0) return MIRA_OK;
  if (!ctx->shards.empty()) {
    if (scalars_on_device) return fail(MIRA_ERR_INVALID, "a sharded context takes host vectors");
    for (size_t g = 0; g < ctx->shards.size(); g++) {
      const size_t lo = ctx->shard_lo[g], hi = ctx->shard_lo[g + 1];
      int rc = n > lo ? mira_msm_ctx_prepare_for(ctx->shards[g], (const char*)scalars + lo * 32, std::min(n, hi) - lo, 0) : MIRA_OK;
      if (rc) return rc;
    }
    return MIRA_OK;
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  return ops_for(ctx->curve).prepare(ctx, n, scalars, scalars_on_device);
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

int mira_msm_commit_batch(mira_msm_ctx* ctx, const void* const* scalars_dev, size_t count, size_t n, void* out_affine, void* stream) {
  if (!ctx || (count && (!scalars_dev || !out_affine))) return fail(MIRA_ERR_INVALID, "null argument");
  if (count > 32) return fail(MIRA_ERR_INVALID, "at most 32 vectors per batched commit");
  if (n > ctx->n_bases)
    return fail(MIRA_ERR_TOO_LONG_INPUT, "Can't commit too long input: input len: %zu, but limit is %zu", n, ctx->n_bases);
  for (size_t k = 0; k < count; k++)
    if (n && !scalars_dev[k]) return fail(MIRA_ERR_INVALID, "vector %zu is null", k);
  if (!ctx->shards.empty()) return fail(MIRA_ERR_INVALID, "device-resident vectors live on one device: not available on a sharded context");
  if (!count) return MIRA_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  return ops_for(ctx->curve).commit_batch(ctx, scalars_dev, count, n, out_affine, stream_for(ctx, 1, stream));
}

int mira_msm_combine(int curve, const void* partials, size_t count, int device, void* out_affine) {
  if (!out_affine || (count && !partials)) return fail(MIRA_ERR_INVALID, "null argument");
  if (!valid_curve(curve)) return fail(MIRA_ERR_INVALID, "unknown curve %d", curve);
  CU(cudaSetDevice(device));
  return ops_for(curve).combine(partials, count, out_affine);
}

int mira_msm_partial_batch_dev(mira_msm_ctx* ctx, const void* const* scalars_dev, size_t count, size_t n, void* out_xyzz_dev, void* stream) {
  if (!ctx || (count && (!scalars_dev || !out_xyzz_dev))) return fail(MIRA_ERR_INVALID, "null argument");
  if (count > 32) return fail(MIRA_ERR_INVALID, "at most 32 vectors per batched commit");
  if (n > ctx->n_bases)
    return fail(MIRA_ERR_TOO_LONG_INPUT, "Can't commit too long input: input len: %zu, but limit is %zu", n, ctx->n_bases);
  for (size_t k = 0; k < count; k++)
    if (n && !scalars_dev[k]) return fail(MIRA_ERR_INVALID, "vector %zu is null", k);
  if (!ctx->shards.empty()) return fail(MIRA_ERR_INVALID, "device-resident vectors live on one device: not available on a sharded context");
  if (!count) return MIRA_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  return ops_for(ctx->curve).partial_batch_dev(ctx, scalars_dev, count, n, out_xyzz_dev, stream_for(ctx, 1, stream));
}

int mira_msm_combine_dev(int curve, const void* partials_dev, size_t n_ranks, size_t n_commits, size_t rank_stride, int device,
                         void* out_affine, void* stream) {
  if (n_commits && (!out_affine || (n_ranks && !partials_dev))) return fail(MIRA_ERR_INVALID, "null argument");
  if (!valid_curve(curve)) return fail(MIRA_ERR_INVALID, "unknown curve %d", curve);
  if (rank_stride < n_commits * 128 || rank_stride % 32) return fail(MIRA_ERR_INVALID, "rank stride %zu does not hold %zu partials", rank_stride, n_commits);
  if (!n_commits) return MIRA_OK;
  CU(cudaSetDevice(device));
  return ops_for(curve).combine_dev(partials_dev, n_ranks, n_commits, rank_stride, out_affine, (cudaStream_t)stream);
}

int mira_msm_get_stats(const mira_msm_ctx* ctx, mira_msm_stats* out) {
  if (!ctx || !out) return fail(MIRA_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(const_cast<mira_msm_ctx*>(ctx)->mu);
  *out = ctx->stats;
  return MIRA_OK;
}
int mira_msm_set_profiling(mira_msm_ctx* ctx, int enabled) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  { int frc; if (for_each_shard(ctx, &frc, [&](mira_msm_ctx* sub) { return mira_msm_set_profiling(sub, enabled); })) return frc; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->profiling = enabled != 0;
  return MIRA_OK;
}
int mira_msm_set_window(mira_msm_ctx* ctx, int window_bits) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  { int frc; if (for_each_shard(ctx, &frc, [&](mira_msm_ctx* sub) { return mira_msm_set_window(sub, window_bits); })) return frc; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (window_bits != 0 && (window_bits < 2 || window_bits > 26)) return fail(MIRA_ERR_INVALID, "window must be 0 or in [2, 26]");
  ctx->forced_window = window_bits;
  return MIRA_OK;
}

const void* mira_msm_scalars_device(const mira_msm_ctx* ctx, size_t* n_out) {
  if (!ctx) return nullptr;
  std::lock_guard<std::mutex> lk(const_cast<mira_msm_ctx*>(ctx)->mu);
  if (n_out) *n_out = ctx->scalars_valid;      // 0 on a sharded context: the device copies are spread over the shards' devices
  return ctx->scalars_valid ? ctx->scalars.p : nullptr;
}

// ---- plain device memory for callers without a CUDA runtime binding of their own (the Rust shim)
int mira_dev_alloc(int device, size_t bytes, void** out_dev) {
  if (!out_dev) return fail(MIRA_ERR_INVALID, "null argument");
  *out_dev = nullptr;
  CU(cudaSetDevice(device));
  cudaError_t e = cudaMalloc(out_dev, bytes ? bytes : 1);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  return MIRA_OK;
}
int mira_dev_free(int device, void* dev_ptr) {
  if (!dev_ptr) return MIRA_OK;
  CU(cudaSetDevice(device));
  CU(cudaFree(dev_ptr));
  return MIRA_OK;
}
int mira_dev_upload(int device, void* dst_dev, const void* src_host, size_t bytes, void* stream) {
  if (bytes && (!dst_dev || !src_host)) return fail(MIRA_ERR_INVALID, "null argument");
  CU(cudaSetDevice(device));
  CU(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return MIRA_OK;
}
int mira_dev_download(int device, void* dst_host, const void* src_dev, size_t bytes, void* stream) {
  if (bytes && (!dst_host || !src_dev)) return fail(MIRA_ERR_INVALID, "null argument");
  CU(cudaSetDevice(device));
  CU(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return MIRA_OK;
}
int mira_dev_sync(int device, void* stream) {
  CU(cudaSetDevice(device));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return MIRA_OK;
}

int mira_host_register(void* host_ptr, size_t bytes) {
  if (!host_ptr || !bytes) return fail(MIRA_ERR_INVALID, "null argument");
  cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return MIRA_OK;
  }
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "cudaHostRegister(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  return MIRA_OK;
}
int mira_host_unregister(void* host_ptr) {
  if (!host_ptr) return fail(MIRA_ERR_INVALID, "null argument");
  cudaError_t e = cudaHostUnregister(host_ptr);
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

int mira_msm_set_adaptive_window(mira_msm_ctx* ctx, int enabled) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  { int frc; if (for_each_shard(ctx, &frc, [&](mira_msm_ctx* sub) { return mira_msm_set_adaptive_window(sub, enabled); })) return frc; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->adaptive_window = enabled != 0;
  return MIRA_OK;
}

int mira_msm_set_slice_min(mira_msm_ctx* ctx, size_t min_scalars_per_slice) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  { int frc; if (for_each_shard(ctx, &frc, [&](mira_msm_ctx* sub) { return mira_msm_set_slice_min(sub, min_scalars_per_slice); })) return frc; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->slice_min = min_scalars_per_slice ? min_scalars_per_slice : ~(size_t)0;   // 0 = never slice
  return MIRA_OK;
}

int mira_msm_set_pipeline(mira_msm_ctx* ctx, int slices, size_t min_scalars_per_slice) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  { int frc; if (for_each_shard(ctx, &frc, [&](mira_msm_ctx* sub) { return mira_msm_set_pipeline(sub, slices, min_scalars_per_slice); })) return frc; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (slices < 0 || slices > 16) return fail(MIRA_ERR_INVALID, "pipeline slices %d out of range [0, 16]", slices);
  ctx->pipe_slices = slices ? slices : -1;
  ctx->pipe_min_slice = min_scalars_per_slice ? min_scalars_per_slice : (size_t)1 << 20;
  return MIRA_OK;
}

int mira_msm_set_affine_levels(mira_msm_ctx* ctx, int levels) {
  if (!ctx) return fail(MIRA_ERR_INVALID, "null context");
  { int frc; if (for_each_shard(ctx, &frc, [&](mira_msm_ctx* sub) { return mira_msm_set_affine_levels(sub, levels); })) return frc; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (levels != MIRA_AFFINE_THREAD_LOCAL_PAIRS && (levels < 0 || levels > 6))
    return fail(MIRA_ERR_INVALID, "affine levels %d out of range [0, 6] (or MIRA_AFFINE_THREAD_LOCAL_PAIRS)", levels);
  ctx->affine_levels = levels;
  return MIRA_OK;
}

// ---------------------------------------------------------------------- generators / test hooks
int mira_gen_scalars(int curve, uint64_t seed, size_t first, size_t n, int dist, int device, void* out_dev) {
  if (n && !out_dev) return fail(MIRA_ERR_INVALID, "null argument");
  if (!valid_curve(curve)) return fail(MIRA_ERR_INVALID, "unknown curve %d", curve);
  CU(cudaSetDevice(device));
  if (!n) return MIRA_OK;
  return ops_for(curve).gen_scalars(seed, first, n, dist, out_dev);
}

int mira_gen_bases(int curve, uint64_t seed, size_t first, size_t n, int device, void* out_dev) {
  if (n && !out_dev) return fail(MIRA_ERR_INVALID, "null argument");
  if (!valid_curve(curve)) return fail(MIRA_ERR_INVALID, "unknown curve %d", curve);
  CU(cudaSetDevice(device));
  if (!n) return MIRA_OK;
  return ops_for(curve).gen_bases(seed, first, n, out_dev);
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
  int rc = test_field_op_dev(field, op, da, db, n, dout);
  cudaError_t e = cudaSuccess;
  if (rc == MIRA_OK) e = cudaMemcpy(out, dout, n * 32, cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(db); cudaFree(dout);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "test_field_op failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

int mira_test_point_op(int curve, int op, const void* p, const void* q, size_t n, int device, void* out) {
  if (!p || !out || n == 0) return fail(MIRA_ERR_INVALID, "null argument");
  if (!valid_curve(curve)) return fail(MIRA_ERR_INVALID, "unknown curve %d", curve);
  CU(cudaSetDevice(device));
  void *dp = nullptr, *dq = nullptr, *dout = nullptr;
  CU(cudaMalloc(&dp, n * 64));
  CU(cudaMalloc(&dq, n * 64));
  CU(cudaMalloc(&dout, n * 64));
  CU(cudaMemcpy(dp, p, n * 64, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dq, q ? q : p, n * 64, cudaMemcpyHostToDevice));
  int rc = ops_for(curve).test_point_op(op, dp, dq, n, dout);
  cudaError_t e = cudaSuccess;
  if (rc == MIRA_OK) e = cudaMemcpy(out, dout, n * 64, cudaMemcpyDeviceToHost);
  cudaFree(dp); cudaFree(dq); cudaFree(dout);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(MIRA_ERR_CUDA, "test_point_op failed: %s", cudaGetErrorString(e));
  return MIRA_OK;
}

}  // extern "C"
