// Per-sample device code of one rollout step, shared by the stand-alone step kernel (rollout_kernels.cu) and the
// single-launch whole-horizon kernel for small obstacle sets (exact_mlp.cu): gradient blend, nominal DS,
// modulation, RBF policy blend, Euler step.
#pragma once
#include <cfloat>

#include "internal.cuh"

// phase stamps of ONE thread (block 0, thread 0) inside step_sample, only in -DDSMPPI_TCX_PROF builds (tools/tcx_prof.py)
#ifdef DSMPPI_TCX_PROF
__device__ long long* g_step_prof = nullptr;
__device__ int g_step_prof_n = 0;
#define STEP_PROF(id)                                                                  \
  do {                                                                                 \
    if (g_step_prof && blockIdx.x == 0 && threadIdx.x == 0 && g_step_prof_n < 1000) {  \
      g_step_prof[2 * g_step_prof_n] = (id);                                           \
      g_step_prof[2 * g_step_prof_n + 1] = clock64();                                  \
      ++g_step_prof_n;                                                                 \
    }                                                                                  \
  } while (0)
#else
#define STEP_PROF(id) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// Gradient blend (MPPI.py:270-280): w = softmax(-10 * dist_k), grad = sum_k w_k grad_k, distance = dist_0
// ------------------------------------------------------------------------------------------------
// rows[k] = row of row_dist / row_grad holding the k-th closest obstacle of this sample
// (no __restrict__: the whole-horizon kernel writes these rows in the same launch that blends them)
template <int DD = MAXD>
__device__ __forceinline__ void blend(const float* row_dist, const float* row_grad, const int* rows, int K, int d,
                                      float& dist, float (&g)[MAXD]) {
  // three rolled passes over the K rows (maximum, normaliser, weighted sum) instead of eight unrolled copies: the
  // exponentials of the last pass are recomputed, which costs K expf and returns the same bits, and the code stays
  // short -- in the whole-horizon kernels this runs on one warp that is bound by instruction fetch
  float mx = -FLT_MAX;
#pragma unroll 1
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, -10.f * row_dist[rows[k]]);
  float den = 0.f;
#pragma unroll 1
  for (int k = 0; k < K; ++k) den += expf(-10.f * row_dist[rows[k]] - mx);
#pragma unroll
  for (int a = 0; a < DD; ++a) g[a] = 0.f;
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    const int r = rows[k];
    const float w = expf(-10.f * row_dist[r] - mx) / den;
    const float* sg = row_grad + (size_t)r * d;
#pragma unroll
    for (int a = 0; a < DD; ++a)
      if (a < d) g[a] += sg[a] * w;
  }
  dist = row_dist[rows[0]];
}

// ------------------------------------------------------------------------------------------------
// One rollout step for all samples: nominal DS, modulation, RBF policy blend, integration
// (MPPI.py:101-223, LinDS.py:11-21, policy.py:186-199).
// ------------------------------------------------------------------------------------------------
struct StepArgs {
  int N, H, d, t, nk, K;
  float dt, dst_thr, lin_thr, p;
  float goal[MAXD];
  dsmppi_modulation mod;
  const float* seds; int seds_G; float seds_thr;     // DSMPPI_DS_SEDS parameters (dsmppi_ctx::seds)
  const float* row_dist; const float* row_grad; const int* sel_rows;
  const float* mu; const float* sigma; const float* alpha;
  float* traj; float* closest; float* kval; float* dots; float* acts; float* qdot; float* grads;
};

__device__ __forceinline__ float gsigmoid(float x, float y_min, float y_max, float mid, float k) {
  // MPPI.py:352-353 with mid = (x0 + x1) / 2 folded on the host side of the expression
  return y_min + (y_max - y_min) / (1.f + expf(k * (-x + mid)));
}

__device__ __forceinline__ float nan_to_num(float v) {
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? FLT_MAX : -FLT_MAX;
  return v;
}

// ||x||_p for p != 2 (policy.py:186-199 with Policy.p set by a script), out of line: two powf per component would
// otherwise sit inside the RBF loop of every caller, which in the whole-horizon kernels is bound by instruction fetch
static_assert(MAXD == 8, "pnorm_general takes the components by value");
static __device__ __noinline__ float pnorm_general(float x0, float x1, float x2, float x3, float x4, float x5, float x6, float x7,
                                            int d, float p) {
  const float x[MAXD] = {x0, x1, x2, x3, x4, x5, x6, x7};
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < MAXD; ++a)
    if (a < d) acc += powf(x[a], p);
  return powf(acc, 1.f / p);
}

// Where one sample's step finds its inputs: the differentiated rows (distance, gradient) and the ranked row indices of
// this sample, and optionally the state itself.  The stand-alone step kernel reads all of it from global memory; the
// whole-horizon tensor-core kernel keeps the rows in shared memory and the state in registers, because every dependent
// L2 round trip of the ONE warp that steps a CTA's samples (~2 k cycles on this part) is on the rollout's critical path.
struct StepIO {
  const float* row_dist; const float* row_grad; const int* rows;
  const float* q_in;     // nullptr: all_traj[i, t-1]
  float* q_next;         // nullptr, or where q + dt * m goes besides all_traj[i, t]
};

// the stand-alone form: everything from global memory
__device__ __forceinline__ StepIO step_io_global(const StepArgs& s, int i) {
  return StepIO{s.row_dist, s.row_grad, s.sel_rows + (size_t)i * s.K, nullptr, nullptr};
}

// one sample i of step t (s.t is ignored: the whole-horizon kernels keep ONE argument block in the constant bank and
// pass the step they are at); every per-sample vector stays in registers
// D > 0: the joint count is a compile-time constant (no predicated padding lanes, no `a < d` tests); D == 0: generic
template <int D>
__device__ __forceinline__ void step_sample_t(const StepArgs& s, int i, int t, const StepIO& io) {
  constexpr int DD = D > 0 ? D : MAXD;               // components actually walked
  const int d = D > 0 ? D : s.d;
  const size_t st = (size_t)i * s.H + (t - 1);       // state-step index
  float q[MAXD], v[MAXD], vhat[MAXD], e0[MAXD], g[MAXD], u[MAXD], vt[MAXD], m[MAXD];   // only [0, DD) is ever touched
  // S0 nominal DS and its norm (MPPI.py:106-108)
  float ss = 0.f;
#pragma unroll
  for (int a = 0; a < DD; ++a) q[a] = a < d ? (io.q_in ? io.q_in[a] : s.traj[st * d + a]) : 0.f;
  if (s.mod.ds_kind == DSMPPI_DS_MATRIX) {
    // v = (q - q_goal) @ A, not normalised (MPPI_toy.py:89)
#pragma unroll
    for (int cc = 0; cc < DD; ++cc) {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < DD; ++a)
        if (a < d && cc < d) acc += (q[a] - s.goal[a]) * s.mod.ds_A[a * MAXD + cc];
      v[cc] = acc;
    }
  } else if (s.mod.ds_kind == DSMPPI_DS_SEDS) {
    // Gaussian-mixture regression on x = q - q_goal (SEDS.py:36-76)
    const int G = s.seds_G;
    const float* pri = s.seds;
    const float* den = pri + G;
    const float* mux = den + G;
    const float* muy = mux + G * d;
    const float* sinv = muy + G * d;
    const float* Am = sinv + G * d * d;
    float x[MAXD], y[MAXD];
    float dst2 = 0.f;
#pragma unroll
    for (int a = 0; a < DD; ++a) {
      x[a] = a < d ? q[a] - s.goal[a] : 0.f;
      y[a] = 0.f;
      dst2 += x[a] * x[a];
    }
    float psum = 0.f;
    for (int j = 0; j < G; ++j) {                      // responsibilities Priors_j N(x; Mu_j, Sigma_j)  (:28-34,48-49)
      float quad = 0.f;
#pragma unroll
      for (int r = 0; r < DD; ++r) {
        if (r < d) {
          float row = 0.f;
#pragma unroll
          for (int cc = 0; cc < DD; ++cc)
            if (cc < d) row += (x[cc] - mux[j * d + cc]) * sinv[(j * d + cc) * d + r];
          quad += row * (x[r] - mux[j * d + r]);
        }
      }
      psum += pri[j] * (expf(-0.5f * quad) / den[j]);
    }
    for (int j = 0; j < G; ++j) {
      float quad = 0.f;
#pragma unroll
      for (int r = 0; r < DD; ++r) {
        if (r < d) {
          float row = 0.f;
#pragma unroll
          for (int cc = 0; cc < DD; ++cc)
            if (cc < d) row += (x[cc] - mux[j * d + cc]) * sinv[(j * d + cc) * d + r];
          quad += row * (x[r] - mux[j * d + r]);
        }
      }
      float beta = nan_to_num(pri[j] * (expf(-0.5f * quad) / den[j]) / psum);     // :50-51
      beta = fmaxf(beta, 1e-8f);                                                  // :52
#pragma unroll
      for (int r = 0; r < DD; ++r) {
        if (r < d) {
          float yr = 0.f;
#pragma unroll
          for (int cc = 0; cc < DD; ++cc)
            if (cc < d) yr += Am[(j * d + r) * d + cc] * (x[cc] - mux[j * d + cc]);
          y[r] += beta * (muy[j * d + r] + yr);                                   // :53-59
        }
      }
    }
    float yn2 = 0.f;
#pragma unroll
    for (int a = 0; a < DD; ++a) yn2 += y[a] * y[a];
    const float ynorm = sqrtf(yn2), dst = sqrtf(dst2);
    const bool far = dst > s.lin_thr;                                             // :63
    const bool weak = ynorm < s.seds_thr;                                         // :72
#pragma unroll
    for (int a = 0; a < DD; ++a) {
      float va = y[a];
      if (far) va = weak ? -x[a] / dst : y[a] / ynorm;                            // :68-75 (|-x| == dst)
      v[a] = a < d ? va : 0.f;
    }
  } else {
    // unit-speed attractor, linear inside lin_thr (LinDS.py:11-21)
#pragma unroll
    for (int a = 0; a < DD; ++a) {
      v[a] = a < d ? -(q[a] - s.goal[a]) : 0.f;
      ss += v[a] * v[a];
    }
    const float dst = sqrtf(ss);
    if (dst > s.lin_thr) {
#pragma unroll
      for (int a = 0; a < DD; ++a) v[a] = v[a] / dst;
    }
  }
  ss = 0.f;
#pragma unroll
  for (int a = 0; a < DD; ++a) ss += v[a] * v[a];
  const float vn = sqrtf(ss);
#pragma unroll
  for (int a = 0; a < DD; ++a) vhat[a] = a < d ? v[a] / vn : 0.f;

  STEP_PROF(201);
  // S2e blended distance / gradient
  float dist;
  blend<DD>(io.row_dist, io.row_grad, io.rows, s.K, d, dist, g);
  dist -= s.dst_thr;                                  // MPPI.py:117
  s.closest[st] = dist;
  ss = 0.f;
#pragma unroll
  for (int a = 0; a < DD; ++a) {
    if (a < d) s.grads[st * d + a] = g[a];
    ss += g[a] * g[a];
  }
  const float gn = sqrtf(ss);
  float dot = 0.f;
#pragma unroll
  for (int a = 0; a < DD; ++a) {
    e0[a] = a < d ? g[a] / gn : 0.f;                  // MPPI.py:126
    dot += e0[a] * vhat[a];                           // MPPI.py:129
  }
  s.dots[st] = dot;
  STEP_PROF(202);
  // S3 modulation coefficients (MPPI.py:132,149-155)
  const float l_vel = gsigmoid(dot, 0.f, 1.f, s.mod.lvel_mid, s.mod.lvel_k);
  const float l_n = gsigmoid(dist, 0.f, 1.f, s.mod.dist_mid, s.mod.dist_k);
  const float l_tau = gsigmoid(dist, s.mod.ltau_max, 1.f, s.mod.dist_mid, s.mod.dist_k);
  const float l_nv = l_vel * 1.f + (1.f - l_vel) * l_n;

  // activations (MPPI.py:191-196)
  float ga = 0.f;
#pragma unroll
  for (int a = 0; a < DD; ++a)
    if (a < d) ga += sqrtf(fabsf(q[a] - s.goal[a]));
  ga = ga * ga;                                        // (sum |x|^0.5)^(1/0.5)
  ga = fminf(fmaxf(ga, 0.f), 1.f);
  if (ga < s.mod.goal_act_thr) ga = 0.f;
  const float act = (1.f - l_n) * (1.f - l_vel) * ga;
  s.acts[st] = act;
  const float kv_scale = s.mod.fold_activation ? act : 1.f;   // MPPI_toy.py:178-179
  STEP_PROF(203);
  // S4 RBF policy (policy.py:186-199, MPPI.py:165-186)
#pragma unroll
  for (int a = 0; a < DD; ++a) u[a] = 0.f;
  // the sampled policy is read-only for the whole rollout: ld.global.nc (the whole-horizon kernel prefetches the rows
  // into L1 before it ranks).  A rolled loop: in the whole-horizon kernels ONE warp per CTA runs this between two
  // network tiles and is bound by instruction fetch, not by arithmetic
#pragma unroll 1
  for (int k = 0; k < s.nk; ++k) {
    const float* mu = s.mu + ((size_t)i * NKMAX + k) * d;
    const float* al = s.alpha + ((size_t)i * NKMAX + k) * d;
    float muv[MAXD], alv[MAXD];
#pragma unroll
    for (int a = 0; a < DD; ++a) {
      muv[a] = a < d ? __ldg(mu + a) : 0.f;
      alv[a] = a < d ? __ldg(al + a) : 0.f;
    }
    const float sg = __ldg(s.sigma + (size_t)i * NKMAX + k);
    float acc = 0.f;
    if (s.p == 2.f) {
#pragma unroll
      for (int a = 0; a < DD; ++a)
        if (a < d) { const float df = q[a] - muv[a]; acc += df * df; }
      acc = sqrtf(acc);
    } else {
      float df[MAXD];
#pragma unroll
      for (int a = 0; a < MAXD; ++a) df[a] = (a < DD && a < d) ? fabsf(q[a] - muv[a]) : 0.f;
      acc = pnorm_general(df[0], df[1], df[2], df[3], df[4], df[5], df[6], df[7], d, s.p);
    }
    const float num = acc * acc;                       // norm ** 2
    const float phi = expf(-sg * num);
    s.kval[st * NKMAX + k] = s.mod.fold_activation ? phi * kv_scale : phi;   // MPPI.py:184
#pragma unroll
    for (int a = 0; a < DD; ++a)
      if (a < d) u[a] += alv[a] * phi;                 // MPPI.py:174-177
  }
  STEP_PROF(204);
  // total velocity and modulation M = l_tau I + (l_nv - l_tau) e0 e0^T  (MPPI.py:158-161,197-209)
  float proj = 0.f;
#pragma unroll
  for (int a = 0; a < DD; ++a) {
    vt[a] = v[a] + act * u[a] * vn;
    proj += e0[a] * vt[a];
  }
  ss = 0.f;
#pragma unroll
  for (int a = 0; a < DD; ++a) {
    m[a] = l_tau * vt[a] + (l_nv - l_tau) * e0[a] * proj;
    if (a >= d) m[a] = 0.f;
    ss += m[a] * m[a];
  }
  float mn = sqrtf(ss);
  if (mn <= 0.5f) mn = 1.f;                            // MPPI.py:211-212
  const bool coll = dist < 0.f;
#pragma unroll
  for (int a = 0; a < DD; ++a) {
    float mv = nan_to_num(m[a] / mn);                  // MPPI.py:213
    if (coll) mv = mv * 0.1f + e0[a] * vn * s.mod.repulsion;   // MPPI.py:215-217
    m[a] = mv;
  }
#pragma unroll
  for (int a = 0; a < DD; ++a) {
    if (a < d) {
      const float qn = q[a] + s.dt * m[a];                        // MPPI.py:220-221
      if (t < s.H) s.traj[(st + 1) * d + a] = qn;
      if (io.q_next) io.q_next[a] = qn;
    }
  }
  STEP_PROF(205);
  if (t == 1) {
#pragma unroll
    for (int a = 0; a < DD; ++a)
      if (a < d) s.qdot[(size_t)i * d + a] = m[a];                // MPPI.py:222-223
  }
}


// ------------------------------------------------------------------------------------------------
// The same step on a GROUP of GL lanes (rank_step_kernel): lane a owns joint a, lane k owns candidate row k and the
// policy kernels k, k + GL, ...  step_sample_t is one thread walking ~5 k dependent instructions and two dozen dependent
// loads; here the loads of a phase are issued together and the per-joint / per-row / per-kernel arithmetic runs side
// by side.  EVERY floating-point operation is the one step_sample_t performs, on the same operands, and every sum is
// accumulated in the same order (the library is built with -fmad=false; a sum over joints is re-played sequentially
// from shuffled terms), so the result is bit-identical -- the tests that compare the prefilter path with the
// all-pairs path (which steps with step_sample_t) hold the two together.
// Covers what the shipped robots use: linear-attractor nominal DS, p = 2, d = D in {2, 7} (step_group_supported); the
// callers keep the one-thread form for everything else.  All GL lanes of every group of the warp must call it (live
// or not).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool step_group_supported(const StepArgs& s) {
  return s.mod.ds_kind == DSMPPI_DS_LINEAR_ATTRACTOR && s.p == 2.f && (s.d == 7 || s.d == 2);
}

template <int DD, int GL>
__device__ __forceinline__ float group_ordered_sum(float term) {
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < DD; ++a) acc += __shfl_sync(0xffffffffu, term, a, GL);
  return acc;
}

// io: as for step_sample_t (rows / state from global or shared memory); a group that is not `live` (padding of the
// last CTA) computes on sample i's state and row 0 and stores nothing.
template <int D, int GL>
// q_lane (optional): this lane's joint of the state, carried in a register by a caller that steps the same sample again
// (the whole-horizon kernel): read instead of io.q_in / all_traj, and updated.
__device__ __forceinline__ void step_group_t(const StepArgs& s, int i, int t, const StepIO& io, int gl, bool live,
                                             float* q_lane = nullptr) {
  static_assert(D >= 1 && D <= GL && GL <= 32, "one lane per joint");
  static_assert(MAXK <= GL, "one lane per ranked row");
  constexpr int d = D;
  const size_t st = (size_t)i * s.H + (t - 1);
  const bool mine = gl < d;                                   // this lane owns a joint
  const int K = s.K;
  // ---- loads of the whole step's row / state inputs, issued together
  const float q = mine ? (q_lane ? *q_lane : io.q_in ? io.q_in[gl] : s.traj[st * d + gl]) : 0.f;
  const int row_k = live ? io.rows[gl < K ? gl : 0] : 0;
  const float dist_k = io.row_dist[row_k];
  float gk[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k) gk[k] = (k < K && mine) ? io.row_grad[(size_t)(live ? io.rows[k] : 0) * d + gl] : 0.f;
  const float goal = mine ? s.goal[gl] : 0.f;
  // S0 nominal DS: unit-speed attractor, linear inside lin_thr (LinDS.py:11-21), and its norm (MPPI.py:106-108)
  float v = mine ? -(q - goal) : 0.f;
  {
    const float dst = sqrtf(group_ordered_sum<D, GL>(v * v));
    if (dst > s.lin_thr) v = v / dst;
  }
  const float vn = sqrtf(group_ordered_sum<D, GL>(v * v));
  const float vhat = mine ? v / vn : 0.f;
  // S2e blended distance / gradient (MPPI.py:270-280): lane k holds row k's softmax weight
  float mx = gl < K ? -10.f * dist_k : -FLT_MAX;
#pragma unroll
  for (int off = GL / 2; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off, GL));
  const float ek = expf(-10.f * dist_k - mx);
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < MAXK; ++k) {
    const float e = __shfl_sync(0xffffffffu, ek, k, GL);
    if (k < K) den += e;
  }
  const float wk = ek / den;
  float g = 0.f;
#pragma unroll
  for (int k = 0; k < MAXK; ++k) {
    const float w = __shfl_sync(0xffffffffu, wk, k, GL);
    if (k < K) g += gk[k] * w;
  }
  float dist = __shfl_sync(0xffffffffu, dist_k, 0, GL);
  dist -= s.dst_thr;                                          // MPPI.py:117
  if (live && gl == 0) s.closest[st] = dist;
  if (live && mine) s.grads[st * d + gl] = g;
  const float gn = sqrtf(group_ordered_sum<D, GL>(g * g));
  const float e0 = mine ? g / gn : 0.f;                       // MPPI.py:126
  const float dot = group_ordered_sum<D, GL>(e0 * vhat);      // MPPI.py:129
  if (live && gl == 0) s.dots[st] = dot;
  // S3 modulation coefficients (MPPI.py:132,149-155) and activation (MPPI.py:191-196): scalars, every lane the same
  const float l_vel = gsigmoid(dot, 0.f, 1.f, s.mod.lvel_mid, s.mod.lvel_k);
  const float l_n = gsigmoid(dist, 0.f, 1.f, s.mod.dist_mid, s.mod.dist_k);
  const float l_tau = gsigmoid(dist, s.mod.ltau_max, 1.f, s.mod.dist_mid, s.mod.dist_k);
  const float l_nv = l_vel * 1.f + (1.f - l_vel) * l_n;
  float ga = group_ordered_sum<D, GL>(mine ? sqrtf(fabsf(q - goal)) : 0.f);
  ga = ga * ga;
  ga = fminf(fmaxf(ga, 0.f), 1.f);
  if (ga < s.mod.goal_act_thr) ga = 0.f;
  const float act = (1.f - l_n) * (1.f - l_vel) * ga;
  if (live && gl == 0) s.acts[st] = act;
  const float kv_scale = s.mod.fold_activation ? act : 1.f;   // MPPI_toy.py:178-179
  // S4 RBF policy (policy.py:186-199, MPPI.py:165-186): lane j evaluates kernels j, j + GL, ...; lane a accumulates
  // joint a's velocity over the kernels in ascending order
  float qa[D];
#pragma unroll
  for (int a = 0; a < D; ++a) qa[a] = __shfl_sync(0xffffffffu, q, a, GL);
  float u = 0.f;
  for (int k0 = 0; k0 < s.nk; k0 += GL) {
    const int k = k0 + gl;
    const bool has = k < s.nk;
    const float* mu = s.mu + ((size_t)i * NKMAX + (has ? k : 0)) * d;
    float muv[D], alv[GL];
#pragma unroll
    for (int a = 0; a < D; ++a) muv[a] = __ldg(mu + a);
    const float sg = __ldg(s.sigma + (size_t)i * NKMAX + (has ? k : 0));
#pragma unroll
    for (int j = 0; j < GL; ++j)
      alv[j] = (mine && k0 + j < s.nk) ? __ldg(s.alpha + ((size_t)i * NKMAX + k0 + j) * d + gl) : 0.f;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < D; ++a) { const float df = qa[a] - muv[a]; acc += df * df; }
    acc = sqrtf(acc);
    const float num = acc * acc;                              // norm ** 2
    const float phi = expf(-sg * num);
    if (live && has) s.kval[st * NKMAX + k] = s.mod.fold_activation ? phi * kv_scale : phi;   // MPPI.py:184
#pragma unroll
    for (int j = 0; j < GL; ++j) {
      const float ph = __shfl_sync(0xffffffffu, phi, j, GL);
      if (k0 + j < s.nk) u += alv[j] * ph;                    // MPPI.py:174-177
    }
  }
  // total velocity and modulation M = l_tau I + (l_nv - l_tau) e0 e0^T  (MPPI.py:158-161,197-209)
  const float vt = v + act * u * vn;
  const float proj = group_ordered_sum<D, GL>(e0 * vt);
  float m = l_tau * vt + (l_nv - l_tau) * e0 * proj;
  if (!mine) m = 0.f;
  float mn = sqrtf(group_ordered_sum<D, GL>(m * m));
  if (mn <= 0.5f) mn = 1.f;                                   // MPPI.py:211-212
  float mv = nan_to_num(m / mn);                              // MPPI.py:213
  if (dist < 0.f) mv = mv * 0.1f + e0 * vn * s.mod.repulsion; // MPPI.py:215-217
  if (live && mine) {
    const float qn = q + s.dt * mv;                           // MPPI.py:220-221
    if (t < s.H) s.traj[(st + 1) * d + gl] = qn;
    if (io.q_next) io.q_next[gl] = qn;
    if (q_lane) *q_lane = qn;
    if (t == 1) s.qdot[(size_t)i * d + gl] = mv;              // MPPI.py:222-223
  }
}

// dispatch on the joint counts the shipped robots have (planar 2-DoF, planar / Franka 7-DoF); anything else runs generic
__device__ __forceinline__ void step_sample(const StepArgs& s, int i, int t, const StepIO& io) {
  if (s.d == 7) step_sample_t<7>(s, i, t, io);
  else if (s.d == 2) step_sample_t<2>(s, i, t, io);
  else step_sample_t<0>(s, i, t, io);
}

// host side: the argument block of step `t` of rollout `a`
static inline StepArgs make_step_args(const dsmppi_ctx* c, const dsmppi_rollout_args* a, int t) {
  StepArgs s;
  s.N = a->N; s.H = a->H; s.d = c->d; s.t = t; s.nk = a->n_kernels; s.K = a->n_closest;
  s.dt = a->dt; s.dst_thr = a->dst_thr; s.lin_thr = a->lin_thr; s.p = a->rbf_p;
  for (int i = 0; i < MAXD; ++i) s.goal[i] = a->q_goal[i];
  s.mod = a->mod;
  s.seds = c->seds; s.seds_G = c->seds_G; s.seds_thr = c->seds_thr;
  s.row_dist = c->row_dist; s.row_grad = c->row_grad; s.sel_rows = c->sel_rows;
  s.mu = a->mu_tmp_dev; s.sigma = a->sigma_tmp_dev; s.alpha = a->alpha_tmp_dev;
  s.traj = a->all_traj_dev; s.closest = a->closest_dist_all_dev; s.kval = a->kernel_val_all_dev;
  s.dots = a->dot_products_dev; s.acts = a->kernel_activations_dev; s.qdot = a->qdot_dev;
  s.grads = a->nn_grad_all_dev;
  return s;
}
