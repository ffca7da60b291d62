// coop.cuh — low-latency group operations: four warps share one XYZZ addition.
//
// The bucket reduction of a SMALL commit (2^14..2^18 buckets: the 2^16..2^20-point commits of BASELINE.json's sweep,
// and every commit of a row-sharded fold step) is a chain of a few dozen dependent group additions on a nearly idle
// GPU.  A lone warp needs ~7 us per XYZZ addition: its 14 field products are ~1,900 IMAD.WIDE warp instructions,
// each occupying the FMA-heavy pipe of its scheduler for 4 cycles, one after the other.  An SM has four schedulers.
// Here a block of 128 threads = 4 warps works on 32 independent additions (one per lane index): all four warps hold
// the operands, warp w computes the w-th product of each formula stage on ITS scheduler, the products are exchanged
// through shared memory, and the addition is 4 product latencies deep instead of 14 (doubling: 3 instead of 9).
// (Round 1 tried four LANES per addition: pipe occupancy is per warp instruction, so that could not help.)
//
// Replaces, for small bucket sets, the running sums / weighting / tree sums of halo2's multiexp_serial "summation by
// parts" (the reference's bucket reduction behind src/commitment.rs:80).  Every function here must be called by all
// 128 threads of the block with block-uniform control flow (it synchronises).
#pragma once
#include "curve.cuh"

namespace mira {

constexpr int COOP_THREADS = 128;

template <class CF>
struct CoopSmem {
  uint32_t prod[2][4][8][32];     // [buffer][warp][limb][lane]: the four products of a stage (double-buffered)
  uint32_t pt[32][32];            // [word][lane]: one XYZZ per lane, for exchanges between lanes
  int flag;
};

template <class CF>
struct Coop {
  CoopSmem<CF>& sm;
  const int lane, warp;
  int buf;
  __device__ Coop(CoopSmem<CF>& s) : sm(s), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), buf(0) {}

  // every warp contributes one product; afterwards all warps hold all four
  __device__ __forceinline__ void exchange(const Fe<CF>& mine, Fe<CF>& p0, Fe<CF>& p1, Fe<CF>& p2, Fe<CF>& p3) {
#pragma unroll
    for (int i = 0; i < 8; i++) sm.prod[buf][warp][i][lane] = mine.v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; i++) {
      p0.v[i] = sm.prod[buf][0][i][lane];
      p1.v[i] = sm.prod[buf][1][i][lane];
      p2.v[i] = sm.prod[buf][2][i][lane];
      p3.v[i] = sm.prod[buf][3][i][lane];
    }
    buf ^= 1;       // the next stage writes the other buffer: one barrier per stage is enough
  }

  // acc + q (both XYZZ), every exceptional case of xyzz_add included
  __device__ Xyzz<CF> add(const Xyzz<CF>& acc, const Xyzz<CF>& q) {
    const bool acc_id = xyzz_is_identity(acc), q_id = xyzz_is_identity(q);
    Fe<CF> mine = fe_zero<CF>(), u1, u2, s1, s2;
    // stage A: u1 = x1*zz2, u2 = x2*zz1, s1 = y1*zzz2, s2 = y2*zzz1
    if (warp == 0) mine = fe_mulc(acc.x, q.zz);
    else if (warp == 1) mine = fe_mulc(q.x, acc.zz);
    else if (warp == 2) mine = fe_mulc(acc.y, q.zzz);
    else mine = fe_mulc(q.y, acc.zzz);
    exchange(mine, u1, u2, s1, s2);
    const Fe<CF> p = fe_sub(u2, u1), r = fe_sub(s2, s1);
    const bool same_x = !acc_id && !q_id && fe_is_zero(p);      // doubling or cancellation: rare, handled below
    // stage B: pp = p^2, rr = r^2, zzm = zz1*zz2, zzzm = zzz1*zzz2
    Fe<CF> pp, rr, zzm, zzzm;
    if (warp == 0) mine = fe_sqrc(p);
    else if (warp == 1) mine = fe_sqrc(r);
    else if (warp == 2) mine = fe_mulc(acc.zz, q.zz);
    else mine = fe_mulc(acc.zzz, q.zzz);
    exchange(mine, pp, rr, zzm, zzzm);
    // stage C: ppp = p*pp, qq = u1*pp, zz3 = zzm*pp
    Fe<CF> ppp, qq, zz3, unused;
    if (warp == 0) mine = fe_mulc(p, pp);
    else if (warp == 1) mine = fe_mulc(u1, pp);
    else if (warp == 2) mine = fe_mulc(zzm, pp);
    exchange(mine, ppp, qq, zz3, unused);
    const Fe<CF> x3 = fe_sub(fe_sub(rr, ppp), fe_dbl(qq));
    // stage D: t1 = r*(qq - x3), t2 = s1*ppp, zzz3 = zzzm*ppp
    Fe<CF> t1, t2, zzz3;
    if (warp == 0) mine = fe_mulc(r, fe_sub(qq, x3));
    else if (warp == 1) mine = fe_mulc(s1, ppp);
    else if (warp == 2) mine = fe_mulc(zzzm, ppp);
    exchange(mine, t1, t2, zzz3, unused);
    Xyzz<CF> out;
    out.x = x3;
    out.y = fe_sub(t1, t2);
    out.zz = zz3;
    out.zzz = zzz3;
    if (q_id) out = acc;
    else if (acc_id) out = q;
    // same x: every warp redoes that lane's addition alone (tangent or cancellation), block-uniformly
    if (__syncthreads_or(same_x ? 1 : 0)) {
      if (same_x) {
        Xyzz<CF> t = acc;
        xyzz_add(t, q);
        out = t;
      }
    }
    return out;
  }

  // 2 * p
  __device__ Xyzz<CF> dbl(const Xyzz<CF>& p) {
    const bool id = xyzz_is_identity(p) || fe_is_zero(p.y);
    const Fe<CF> u = fe_dbl(p.y);
    Fe<CF> mine = fe_zero<CF>(), v, xx, unused, unused2;
    // stage A: v = u^2, xx = x^2
    if (warp == 0) mine = fe_sqrc(u);
    else if (warp == 1) mine = fe_sqrc(p.x);
    exchange(mine, v, xx, unused, unused2);
    const Fe<CF> m = fe_add(fe_dbl(xx), xx);
    // stage B: w = u*v, s = x*v, zz3 = v*zz, mm = m^2
    Fe<CF> w, s, zz3, mm;
    if (warp == 0) mine = fe_mulc(u, v);
    else if (warp == 1) mine = fe_mulc(p.x, v);
    else if (warp == 2) mine = fe_mulc(v, p.zz);
    else mine = fe_sqrc(m);
    exchange(mine, w, s, zz3, mm);
    const Fe<CF> x3 = fe_sub(mm, fe_dbl(s));
    // stage C: t1 = m*(s - x3), t2 = w*y, zzz3 = w*zzz
    Fe<CF> t1, t2, zzz3;
    if (warp == 0) mine = fe_mulc(m, fe_sub(s, x3));
    else if (warp == 1) mine = fe_mulc(w, p.y);
    else if (warp == 2) mine = fe_mulc(w, p.zzz);
    exchange(mine, t1, t2, zzz3, unused);
    Xyzz<CF> out;
    out.x = x3;
    out.y = fe_sub(t1, t2);
    out.zz = zz3;
    out.zzz = zzz3;
    if (id) out = xyzz_identity<CF>();
    return out;
  }

  // the value lane `lane + delta` holds (identity beyond lane 31); all warps hold the same values, warp 0 publishes
  __device__ Xyzz<CF> from_lane_above(const Xyzz<CF>& mine, int delta) {
    __syncthreads();                       // earlier readers of sm.pt are done
    if (warp == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        sm.pt[i][lane] = mine.x.v[i];
        sm.pt[8 + i][lane] = mine.y.v[i];
        sm.pt[16 + i][lane] = mine.zz.v[i];
        sm.pt[24 + i][lane] = mine.zzz.v[i];
      }
    }
    __syncthreads();
    Xyzz<CF> o = xyzz_identity<CF>();
    const int src = lane + delta;
    if (src < 32) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        o.x.v[i] = sm.pt[i][src];
        o.y.v[i] = sm.pt[8 + i][src];
        o.zz.v[i] = sm.pt[16 + i][src];
        o.zzz.v[i] = sm.pt[24 + i][src];
      }
    }
    return o;
  }

  // sum over the 32 lanes, valid in lane 0
  __device__ Xyzz<CF> lane_sum(Xyzz<CF> v) {
    for (int d = 16; d > 0; d >>= 1) {
      Xyzz<CF> o = from_lane_above(v, d);
      if (lane >= d) o = xyzz_identity<CF>();
      v = add(v, o);
    }
    return v;
  }
};

// ------------------------------------------------------------------ small bucket sets: reduction in two launches
// Launch 1: lane l of block b owns buckets (t*m, (t+1)*m], t = b*32 + l:  out[b] = sum over its lanes of
//   sum_j (j + 1 + t*m) * bucket[t*m + j + 1]   (running sums, weighting by double-and-add, lane sum).
// Launch 2 (k_sum_coop): sums the blocks' outputs.  blockIdx.y selects the bucket set of a batched commit.
template <class CF>
__global__ void __launch_bounds__(COOP_THREADS) k_reduce_coop(const void* __restrict__ buckets_all, uint32_t B, uint32_t m,
                                                              void* __restrict__ out_all, size_t in_stride, size_t out_stride) {
  __shared__ CoopSmem<CF> sm;
  Coop<CF> co(sm);
  const char* buckets = reinterpret_cast<const char*>(buckets_all) + blockIdx.y * in_stride;
  char* out = reinterpret_cast<char*>(out_all) + blockIdx.y * out_stride;
  const uint32_t t = blockIdx.x * 32 + co.lane;
  const uint64_t lo64 = (uint64_t)t * m;
  const bool valid = lo64 < B;
  const uint32_t lo = valid ? (uint32_t)lo64 : 0u;
  const uint32_t hi = valid ? (lo + m < B ? lo + m : B) : 0u;
  Xyzz<CF> run = xyzz_identity<CF>(), acc = xyzz_identity<CF>();
  for (uint32_t j = 0; j < m; j++) {                     // block-uniform trip count; lanes past their range add the identity
    Xyzz<CF> x = xyzz_identity<CF>();
    if (valid && hi - lo > j) x = xyzz_load<CF>(buckets + (size_t)(hi - j) * 128);
    run = co.add(run, x);
    acc = co.add(acc, run);
  }
  // acc += lo * run: double-and-add over the bits of the block's largest lo (uniform), most significant first
  const uint32_t lo_max = (blockIdx.x * 32 + 31) * m;
  Xyzz<CF> wsum = xyzz_identity<CF>();
  for (int b = 31 - __clz(lo_max | 1u); b >= 0; b--) {
    wsum = co.dbl(wsum);
    Xyzz<CF> c = co.add(wsum, run);
    if ((lo >> b) & 1u) wsum = c;
  }
  acc = co.add(acc, wsum);
  acc = co.lane_sum(acc);
  if (threadIdx.x == 0) xyzz_store<CF>(out + (size_t)blockIdx.x * 128, acc);
}

// out[block] = sum of in[block * 32 * per_lane ...]: per_lane sequential (cooperative) additions per lane, then the lane sum
template <class CF>
__global__ void __launch_bounds__(COOP_THREADS) k_sum_coop(const void* __restrict__ in_all, uint32_t n, uint32_t per_lane,
                                                           void* __restrict__ out_all, size_t in_stride, size_t out_stride) {
  __shared__ CoopSmem<CF> sm;
  Coop<CF> co(sm);
  const char* in = reinterpret_cast<const char*>(in_all) + blockIdx.y * in_stride;
  char* out = reinterpret_cast<char*>(out_all) + blockIdx.y * out_stride;
  const uint32_t base = blockIdx.x * 32 * per_lane;
  Xyzz<CF> acc = xyzz_identity<CF>();
  for (uint32_t k = 0; k < per_lane; k++) {
    const uint32_t idx = base + k * 32 + co.lane;            // consecutive lanes read consecutive points
    Xyzz<CF> x = xyzz_identity<CF>();
    if (idx < n) x = xyzz_load<CF>(in + (size_t)idx * 128);
    acc = co.add(acc, x);
  }
  acc = co.lane_sum(acc);
  if (threadIdx.x == 0) xyzz_store<CF>(out + (size_t)blockIdx.x * 128, acc);
}

}  // namespace mira
