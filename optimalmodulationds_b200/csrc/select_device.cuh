// Candidate band of ONE sample from its approximate (tensor-core) distances: every obstacle within `band` of the
// K-th smallest approximate distance gets a row in the shared list and is re-scored in fp32.  Executed by one warp
// (select_candidates_kernel, rollout_kernels.cu).
#pragma once
#include <cfloat>

#include "internal.cuh"

// `ld(t)` returns the approximate distance of obstacle t of sample w
template <typename Loader>
__device__ __forceinline__ void select_candidates_warp(Loader ld, int M, int K, float band,
                                                       int cap_rows, int w, int lane, int* __restrict__ cand_cnt,
                                                       int* __restrict__ row_base, int* __restrict__ row_sample,
                                                       int* __restrict__ row_obs, int* __restrict__ counters) {
#define LD(p) ld(p)
  // K-th smallest approximate value (with multiplicity), one streaming pass: every lane keeps the K smallest of
  // its own elements sorted in registers (ties: lower obstacle index first), then the warp merges the 32 lists
  float lv[MAXK];
  int lj[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k) { lv[k] = FLT_MAX; lj[k] = 0x7fffffff; }
#pragma unroll 4
  for (int t = lane; t < M; t += 32) {
    float v = LD(t);
    int j = t;
    if (v < lv[MAXK - 1]) {
#pragma unroll
      for (int k = 0; k < MAXK; ++k) {
        if (v < lv[k]) {                      // strict: an equal value seen later (higher index) stays behind
          const float tv = lv[k]; lv[k] = v; v = tv;
          const int tj = lj[k]; lj[k] = j; j = tj;
        }
      }
    }
  }
  float last_v = -FLT_MAX;
  for (int kk = 0; kk < K; ++kk) {
    float bv = lv[0];
    int bj = lj[0];
    int src = lane;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
      const int os = __shfl_xor_sync(0xffffffffu, src, off);
      if (ov < bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; src = os; }
    }
    last_v = bv;
    if (lane == src) {                         // pop the winner's head
#pragma unroll
      for (int k = 0; k + 1 < MAXK; ++k) { lv[k] = lv[k + 1]; lj[k] = lj[k + 1]; }
      lv[MAXK - 1] = FLT_MAX;
      lj[MAXK - 1] = 0x7fffffff;
    }
  }
  const float thr = last_v + band;
  // count, reserve a contiguous row range, then fill (ascending obstacle index within the sample).  EVERY obstacle
  // inside the band gets a row: a crowded band simply takes more of the shared list (budgeted at CAND_MAX rows per
  // sample, typically a third used).  If the list itself runs out, the high-water mark in counters[8] tells the host,
  // which grows the list and runs the rollout again (capi.cu: prefilter_verdict) -- nothing is ever dropped silently.
  int mine = 0;
#pragma unroll 4
  for (int t = lane; t < M; t += 32) mine += (LD(t) <= thr) ? 1 : 0;
  int total = mine;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
  int base = 0;
  if (lane == 0) {
    base = atomicAdd(&counters[0], total);
    atomicMax(&counters[8], base + total);
    if (total > CAND_MAX) atomicAdd(&counters[1], 1);
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  // rows of this sample that fit (all of them unless the list is exhausted; the rows written stay valid indices, so
  // the scoring launch that follows reads nothing out of bounds before the rollout is repeated)
  const int room = base < cap_rows ? cap_rows - base : 0;
  const int keep = total < room ? total : room;
  if (lane == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(&counters[2]), (unsigned long long)keep);
    cand_cnt[w] = keep;
    row_base[w] = base < cap_rows ? base : 0;
  }
  int written = 0;
  for (int t0 = 0; t0 < M && written < keep; t0 += 32) {
    const int t = t0 + lane;
    const bool in = t < M && LD(t) <= thr;
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    const int pos = written + __popc(bal & ((1u << lane) - 1u));
    if (in && pos < keep) {
      row_sample[base + pos] = w;
      row_obs[base + pos] = t;
    }
    written += __popc(bal);
  }
#undef LD
}

