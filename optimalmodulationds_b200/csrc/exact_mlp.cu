// fp32-exact evaluation of the learned distance network on (sample, obstacle) rows.
//
//   exact_mlp_kernel<false, RPT>: forward only  -> masked minimum link distance per row  (MPPI.py:235-242)
//   exact_mlp_kernel<true,  RPT>: forward + analytic VJP at argmin_l of the raw output   (robot_sdf.py:153-158)
//
// One CTA owns R = 4*RPT rows (RPT = 8, 4 or 2).  Activations live in shared memory feature-major
// (act[k][row]) and are updated in place layer by layer.  A warp owns an R-row x (8*FT)-feature block of the layer
// output, its lanes form a 4 x 8 grid and each thread keeps an RPT-row x FT-feature register tile (FT = 8, 4 or 2
// => 4, 8 or 16 warps per CTA), so one k-step of a warp is FT*RPT FFMAs fed by single-wavefront shared-memory
// loads (broadcast over the row / feature groups).  FT = 8 amortises the loads best and is used whenever there are
// enough row tiles to fill the SMs; the 8- and 16-warp shapes put 2 or 4 warps on every scheduler of an SM that
// holds a single tile, which is what the latency of a small batch (and of the whole-horizon kernel) is made of
// (pick_shape below has the measurements; only 8x8, 8x4 and 4x4 are instantiated).
// The weights of all seven GEMMs (4 forward, 3 backward) are one stream of 8-row stages that the CTA pulls from L2
// through a 4-deep cp.async ring, three stages (24 k-steps) ahead of the FFMAs and straight across layer
// boundaries: a lone CTA on an SM is then FFMA-bound instead of waiting ~500 cycles for L2 every four k-steps,
// which is what set the pace of the first versions whenever there were few rows.  ReLU masks stay in
// registers as bit masks (the forward and backward tilings coincide), so the backward pass needs no extra memory.
// The host picks RPT per launch from the (estimated) row count: big tiles amortise the weight stream, small tiles
// fill the 148 SMs when there are few rows and shorten the tail of the last wave (pick_rpt below).
// All arithmetic is IEEE fp32 (no fast-math): this is the path that has to agree with the reference's torch-CPU
// numbers to ~1e-6.
#include <cstdlib>

#include "internal.cuh"
#include "step_device.cuh"

namespace {

__host__ __device__ constexpr int nthreads(int FT) { return 32 * (HID / (8 * FT)); }   // 128 / 256 / 512 threads for FT = 8 / 4 / 2
constexpr int XS = 12;    // padded row stride of the raw-input scratch (nin <= 11)
constexpr int WS = 8;     // k-rows of weights per pipeline stage (8 KB)
constexpr int NSTAGE = 4; // ring depth

__device__ __forceinline__ bool row_lookup(const RowSrc& s, int r, int n_rows, int& i, int& j) {
  if (r >= n_rows) return false;
  if (s.mode == ROWS_DENSE) {
    i = r / s.M;
    j = r - i * s.M;
  } else if (s.mode == ROWS_SELECTED) {
    i = r / s.K;
    j = s.sel[r];
  } else {
    i = s.row_sample[r];
    j = s.row_obs[r];
  }
  return true;
}

// Thread tile: rows r0 .. r0+RPT-1 and FT features: FT = 8: {fa .. fa+3} (j = 0..3) and {fa+32 .. fa+35} (j = 4..7);
// FT = 4 / 2: {fa .. fa+FT-1}.
template <int FT>
struct Tile {
  int r0, fa;
  __device__ __forceinline__ void init(int tid, int RPT) {
    r0 = ((tid & 31) >> 3) * RPT;                                  // lane / 8: one of four RPT-row groups
    fa = (tid >> 5) * (8 * FT) + (tid & 7) * (FT == 8 ? 4 : FT);   // warp: feature block; lane % 8: feature group
  }
  __device__ __forceinline__ int feat(int j) const { return FT == 8 ? fa + (j & 3) + ((j >> 2) << 5) : fa + j; }
};

// FT consecutive-group values of a 256-wide row (weights of one k, or a bias vector) for this thread's tile
template <int FT>
__device__ __forceinline__ void load_feats(const float* p, float (&w)[FT]) {
  if constexpr (FT == 8) {
    const float4 w0 = *reinterpret_cast<const float4*>(p), w1 = *reinterpret_cast<const float4*>(p + 32);
    w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
  } else if constexpr (FT == 4) {
    const float4 w0 = *reinterpret_cast<const float4*>(p);
    w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
  } else {
    const float2 w0 = *reinterpret_cast<const float2*>(p);
    w[0] = w0.x; w[1] = w0.y;
  }
}

template <int RPT>
__device__ __forceinline__ void load_rows(const float* p, float (&a)[RPT]) {
  if constexpr (RPT == 8) {
    const float4 a0 = *reinterpret_cast<const float4*>(p), a1 = *reinterpret_cast<const float4*>(p + 4);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
  } else if constexpr (RPT == 4) {
    const float4 a0 = *reinterpret_cast<const float4*>(p);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
  } else {
    const float2 a0 = *reinterpret_cast<const float2*>(p);
    a[0] = a0.x; a[1] = a0.y;
  }
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Packed fp32 FMA (Blackwell fma.rn.f32x2): two IEEE fused multiply-adds per issue slot, bit-identical to two
// fmaf().  The inner loop is issue-bound with scalar FFMAs (ncu: issue 68 %, fma pipe 51 %), so halving the FFMA
// issue count is what lets the pipe fill.
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2f(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// The weight stream: GEMM g = 0..3 are the forward layers (Wf[g], K = nenc or 256), g = 4..6 the backward ones
// (Wb[3], Wb[2], Wb[1]); stage indices run through all of them.
struct WeightStream {
  const NetDev* net;
  float* ring;            // [NSTAGE][WS][HID]
  int n0;                 // stages of GEMM 0 = ceil(nenc / WS)
  int per_pass;           // stages of one pass over a row tile (forward, or forward + backward)
  int total;              // stages in the whole stream: per_pass x number of passes this CTA makes
  int issued;             // next stage to request

  __device__ __forceinline__ void locate(int stage, const float*& src, int& rows) const {
    int g, st;
    stage %= per_pass;
    if (stage < n0) { g = 0; st = stage; }
    else { g = 1 + (stage - n0) / (HID / WS); st = (stage - n0) % (HID / WS); }
    const int K = g == 0 ? net->nenc : HID;
    const float* W = g < 4 ? net->Wf[g] : net->Wb[7 - g];
    src = W + (size_t)st * WS * HID;
    rows = min(WS, K - st * WS);
  }
  // every thread requests its share of the next stage (or nothing past the end) and closes one group
  template <int NT>
  __device__ __forceinline__ void request_next() {
    if (issued < total) {
      const float* src;
      int rows;
      locate(issued, src, rows);
      float* dst = ring + (size_t)(issued % NSTAGE) * WS * HID;
      for (int c = threadIdx.x; c < rows * (HID / 4); c += NT) cp_async16(dst + c * 4, src + c * 4);
    }
    ++issued;
    cp_async_commit();
  }
};

// acc[i][j] = sum_k act[k][r0+i] * W[k*256 + feat(j)], W arriving through the ring; `stage` is the stream position
// of this GEMM's first stage and is advanced past its last one
template <int RPT, int FT>
__device__ __forceinline__ void gemm_tile(WeightStream& ws, int& stage, int K, const float (*act)[4 * RPT],
                                          const Tile<FT>& t, float (&acc)[RPT][FT]) {
  constexpr int NT = nthreads(FT);
  uint64_t acc2[RPT][FT / 2];             // acc2[i][p] = (acc[i][2p], acc[i][2p+1])
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int p = 0; p < FT / 2; ++p) acc2[i][p] = 0ull;
  for (int k0 = 0; k0 < K; k0 += WS, ++stage) {
    cp_async_wait<NSTAGE - 2>();          // this thread's share of `stage` has landed ...
    __syncthreads();                      // ... and everybody's; the buffer of stage-1 is free again
    ws.template request_next<NT>();       // refill it with stage + NSTAGE - 1
    const float* wb = ws.ring + (size_t)(stage % NSTAGE) * WS * HID + t.fa;
    const int rows = min(WS, K - k0);
    if (rows == WS) {
#pragma unroll
      for (int kk = 0; kk < WS; ++kk) {
        float a[RPT], w[FT];
        load_rows<RPT>(&act[k0 + kk][t.r0], a);
        load_feats<FT>(wb + kk * HID, w);
        uint64_t wp[FT / 2];
#pragma unroll
        for (int p = 0; p < FT / 2; ++p) wp[p] = pack2f(w[2 * p], w[2 * p + 1]);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const uint64_t ad = pack2f(a[i], a[i]);
#pragma unroll
          for (int p = 0; p < FT / 2; ++p) acc2[i][p] = fma2(ad, wp[p], acc2[i][p]);
        }
      }
    } else {
      for (int kk = 0; kk < rows; ++kk) {
        float a[RPT], w[FT];
        load_rows<RPT>(&act[k0 + kk][t.r0], a);
        load_feats<FT>(wb + kk * HID, w);
        uint64_t wp[FT / 2];
#pragma unroll
        for (int p = 0; p < FT / 2; ++p) wp[p] = pack2f(w[2 * p], w[2 * p + 1]);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const uint64_t ad = pack2f(a[i], a[i]);
#pragma unroll
          for (int p = 0; p < FT / 2; ++p) acc2[i][p] = fma2(ad, wp[p], acc2[i][p]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int p = 0; p < FT / 2; ++p) unpack2f(acc2[i][p], acc[i][2 * p], acc[i][2 * p + 1]);
}

template <int RPT, int FT>
__device__ __forceinline__ void store_tile(float (*act)[4 * RPT], const Tile<FT>& t, const float (&v)[RPT][FT]) {
#pragma unroll
  for (int j = 0; j < FT; ++j) {
    float* p = &act[t.feat(j)][t.r0];
    if constexpr (RPT == 8) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(v[4][j], v[5][j], v[6][j], v[7][j]);
    } else if constexpr (RPT == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
    } else {
      *reinterpret_cast<float2*>(p) = make_float2(v[0][j], v[1][j]);
    }
  }
}

template <int RPT>
struct TileSmem {
  static constexpr int R = 4 * RPT;
  float (*act)[R];   // [256][R]
  float* ring;       // [NSTAGE][WS][256] weight stages
  float* xs;         // [R][XS]  raw inputs x = [q, p]
  float* zs;         // [R][MAXO] raw outputs
  float* rad;        // [R]
  int* lst;          // [R] argmin link
  __device__ __forceinline__ explicit TileSmem(float* smem) {
    act = reinterpret_cast<float (*)[R]>(smem);
    ring = smem + HID * R;
    xs = ring + NSTAGE * WS * HID;
    zs = xs + R * XS;
    rad = zs + R * MAXO;
    lst = reinterpret_cast<int*>(rad + R);
  }
};

// starts the weight stream of a CTA that will make `passes` passes over row tiles: the first three stages fly
// while the inputs are encoded
template <bool BWD, int NT>
__device__ __forceinline__ void stream_begin(WeightStream& ws, const NetDev* net, float* ring, int passes) {
  ws.net = net;
  ws.ring = ring;
  ws.n0 = (net->nenc + WS - 1) / WS;
  ws.per_pass = ws.n0 + (BWD ? 6 : 3) * (HID / WS);
  ws.total = ws.per_pass * passes;
  ws.issued = 0;
#pragma unroll
  for (int i = 0; i < NSTAGE - 1; ++i) ws.template request_next<NT>();
}

// One pass of the network over the R rows [row0, row0 + R) of `src` (rows >= n_rows are padding): forward, and with
// BWD the analytic VJP.  Called by all NT threads of the CTA; `ws` / `stage` carry the weight stream across calls.
template <bool BWD, int RPT, int FT>
__device__ __forceinline__ void mlp_tile(const NetDev& net, const RowSrc& src, int row0, int n_rows,
                                         const float* q, int q_stride, const float* __restrict__ obs,
                                         uint32_t ignore_mask, float* out_m, float* out_dist, float* out_grad,
                                         const TileSmem<RPT>& sm, WeightStream& ws, int& stage) {
  constexpr int R = 4 * RPT;
  constexpr int NT = nthreads(FT);
  float (*act)[R] = sm.act;
  float* xs = sm.xs;
  float* zs = sm.zs;
  float* rad = sm.rad;
  int* lst = sm.lst;
  const int tid = threadIdx.x;
  Tile<FT> t;
  t.init(tid, RPT);
  const int d = net.d, nin = net.nin, nenc = net.nenc, O = net.O;

  // ---- rows -> encoded inputs [x, sin x, cos x]  (network_macros_mod.py:139-140)
  for (int idx = tid; idx < R * nin; idx += NT) {
    const int r = idx % R, c = idx / R;
    int i = 0, j = 0;
    const bool valid = row_lookup(src, row0 + r, n_rows, i, j);
    float x = 0.f;
    if (valid) x = (c < d) ? q[(size_t)i * q_stride + c] : obs[j * 4 + (c - d)];
    xs[r * XS + c] = x;
    act[c][r] = x;
    act[nin + c][r] = sinf(x);
    act[2 * nin + c][r] = cosf(x);
    if (c == 0) rad[r] = valid ? obs[j * 4 + 3] : 0.f;
  }
  __syncthreads();

  uint64_t mk[4];          // ReLU masks of this thread's tile, bit i*FT+j
  float acc[RPT][FT];
  // ---- hidden layers: h = relu(W h + b)
#pragma unroll 1
  for (int l = 0; l < 4; ++l) {
    gemm_tile<RPT, FT>(ws, stage, l == 0 ? nenc : HID, act, t, acc);
    __syncthreads();
    float bb[FT];
    load_feats<FT>(net.b[l] + t.fa, bb);
    uint64_t m = 0;
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int j = 0; j < FT; ++j) {
        const float v = acc[i][j] + bb[j];
        const bool on = v > 0.f;
        if (BWD) m |= (uint64_t)(on ? 1u : 0u) << (i * FT + j);
        acc[i][j] = on ? v : 0.f;
      }
    mk[l] = m;
    store_tile<RPT, FT>(act, t, acc);
    __syncthreads();
  }

  // ---- output layer (no activation)
  for (int idx = tid; idx < R * O; idx += NT) {
    const int r = idx % R, o = idx / R;
    const float* w = net.W4 + o * HID;
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < HID; ++k) s = fmaf(act[k][r], __ldg(w + k), s);
    zs[r * MAXO + o] = s + __ldg(net.b[4] + o);
  }
  __syncthreads();

  if constexpr (!BWD) {
    // MPPI.py:236-242: /100 for the 9-link Franka net, minus radius, ignored links := 1e6, min over links
    if (tid < R && row0 + tid < n_rows) {
      float m = 3.0e38f;
      for (int o = 0; o < O; ++o) {
        float y = zs[tid * MAXO + o];
        if (net.scale != 1.f) y = y / 100.f;
        y -= rad[tid];
        if ((ignore_mask >> o) & 1u) y = 1e6f;
        m = fminf(m, y);
      }
      out_m[src.out_row ? src.out_row[row0 + tid] : row0 + tid] = m;
    }
  } else {
  // ---- pass 2: l* = argmin of the RAW output (robot_sdf.py:155), distance of that link (MPPI.py:265-274);
  //      optionally also the pass-1 ranking key (masked minimum) so one launch serves both passes
  if (tid < R) {
    int best = 0;
    float bv = zs[tid * MAXO];
    float m = 3.0e38f;
    for (int o = 0; o < O; ++o) {
      const float v = zs[tid * MAXO + o];
      if (v < bv) { bv = v; best = o; }
      float y = v;
      if (net.scale != 1.f) y = y / 100.f;
      y -= rad[tid];
      if ((ignore_mask >> o) & 1u) y = 1e6f;
      m = fminf(m, y);
    }
    lst[tid] = best;
    if (row0 + tid < n_rows) {
      float y = bv;
      if (net.scale != 1.f) y = y / 100.f;
      const int orow = src.out_row ? src.out_row[row0 + tid] : row0 + tid;
      out_dist[orow] = y - rad[tid];
      if (out_m) out_m[orow] = m;
    }
  }
  __syncthreads();

  // g4 = W5[l*, :] * s4
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const float* w = net.W4 + lst[t.r0 + i] * HID;
    const uint32_t bits = (uint32_t)(mk[3] >> (i * FT)) & 0xffu;
#pragma unroll
    for (int j = 0; j < FT; ++j) acc[i][j] = ((bits >> j) & 1u) ? __ldg(w + t.feat(j)) : 0.f;
  }
  store_tile<RPT, FT>(act, t, acc);
  __syncthreads();

  // g_{l-1} = (W_l^T g_l) * s_{l-1},  l = 3, 2, 1   (Wb[l] is torch's [out][in]: out = k, in = n)
#pragma unroll 1
  for (int l = 3; l >= 1; --l) {
    gemm_tile<RPT, FT>(ws, stage, HID, act, t, acc);
    __syncthreads();
    const uint64_t m = mk[l - 1];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const uint32_t bits = (uint32_t)(m >> (i * FT)) & 0xffu;
#pragma unroll
      for (int j = 0; j < FT; ++j) acc[i][j] = ((bits >> j) & 1u) ? acc[i][j] : 0.f;
    }
    store_tile<RPT, FT>(act, t, acc);
    __syncthreads();
  }

  // a = W_1^T g_1 (only the three entries per joint that are consumed), then the encoding Jacobian:
  // dz/dx_c = a[c] + cos(x_c) a[nin+c] - sin(x_c) a[2nin+c]                       
  for (int idx = tid; idx < R * d; idx += NT) {
    const int r = idx % R, c = idx / R;
    const float* w = net.Wb[0];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
    for (int k = 0; k < HID; ++k) {
      const float g = act[k][r];
      a0 = fmaf(g, __ldg(w + k * nenc + c), a0);
      a1 = fmaf(g, __ldg(w + k * nenc + nin + c), a1);
      a2 = fmaf(g, __ldg(w + k * nenc + 2 * nin + c), a2);
    }
    const float x = xs[r * XS + c];
    if (row0 + r < n_rows) {
      const int orow = src.out_row ? src.out_row[row0 + r] : row0 + r;
      out_grad[(size_t)orow * d + c] = a0 + cosf(x) * a1 - sinf(x) * a2;
    }
  }
  }  // BWD
}

// resident CTAs per SM the register budget is sized for: the 4-warp shape packs 4 (32-row) or 6 (16-row) tiles on
// an SM, the 8-warp shape 2, the 16-warp shape 1
__host__ __device__ constexpr int min_ctas(int RPT, int FT) { return FT == 8 ? (RPT == 8 ? 4 : 6) : (FT == 4 ? (RPT == 8 ? 2 : 3) : 1); }

template <bool BWD, int RPT, int FT>
__global__ void __launch_bounds__(nthreads(FT), min_ctas(RPT, FT))
exact_mlp_kernel(NetDev net, RowSrc src, const float* __restrict__ q, int q_stride, const float* __restrict__ obs,
                 uint32_t ignore_mask, float* __restrict__ out_m, float* __restrict__ out_dist,
                 float* __restrict__ out_grad) {
  constexpr int R = 4 * RPT;
  extern __shared__ __align__(16) float smem[];
  const int n_rows = src.n_rows_dev ? min(*src.n_rows_dev, src.n_rows) : src.n_rows;
  const int tiles = (n_rows + R - 1) / R;
  if ((int)blockIdx.x >= tiles) return;
  TileSmem<RPT> sm(smem);
  WeightStream ws;
  // one tile per CTA in the regular launches; the re-scoring launch (launch_exact_fixup) strides a bounded grid
  stream_begin<BWD, nthreads(FT)>(ws, &net, sm.ring, (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
  int stage = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    mlp_tile<BWD, RPT, FT>(net, src, tile * R, n_rows, q, q_stride, obs, ignore_mask, out_m, out_dist, out_grad, sm, ws,
                           stage);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Whole-horizon rollout in ONE launch for small obstacle sets (M <= R): a CTA owns S = R / M samples and all of
// their (sample, obstacle) rows, and samples never interact, so it can run all H steps by itself -- network
// forward + VJP on its tile, then one thread per sample ranks its M rows, blends the K closest gradients and takes
// the modulation / policy / Euler step (step_device.cuh) -- with no grid-wide dependency and no launch per step.
// The weight stream simply keeps running across steps.  This is the path of the reference's own CPU-sized
// problems (planar scripts: M = 2..4; Franka integrator tick: N = 1, H = 2) where launch latency, not FLOPs,
// set the pace of the per-step launch sequence.
// ------------------------------------------------------------------------------------------------
template <int RPT, int FT>
__global__ void __launch_bounds__(nthreads(FT), FT == 8 ? 3 : 2)
rollout_fused_kernel(NetDev net, StepArgs sa, const float* __restrict__ obs, int M, uint32_t ignore_mask,
                     float* m_rows, float* row_dist, float* row_grad, int* sel_rows) {
  constexpr int R = 4 * RPT;
  extern __shared__ __align__(16) float smem[];
  const int S = R / M;                                  // samples per CTA
  const int i0 = blockIdx.x * S;
  if (i0 >= sa.N) return;
  const int ns = min(S, sa.N - i0);
  const int row0 = i0 * M;                              // dense rows: row = i * M + j
  const int n_rows = row0 + ns * M;
  TileSmem<RPT> sm(smem);
  WeightStream ws;
  stream_begin<true, nthreads(FT)>(ws, &net, sm.ring, sa.H);
  int stage = 0;
  RowSrc src{};
  src.mode = ROWS_DENSE;
  src.M = M;
  const int d = net.d, K = sa.K;
  const int tid = threadIdx.x;
  for (int t = 1; t <= sa.H; ++t) {
    const float* q = sa.traj + (size_t)(t - 1) * d;     // q_prev = all_traj[:, t-1, :]
    mlp_tile<true, RPT, FT>(net, src, row0, n_rows, q, sa.H * d, obs, ignore_mask, m_rows, row_dist, row_grad, sm,
                            ws, stage);
    __syncthreads();                                    // the tile's rows are visible to the whole CTA
    if (tid < ns) {
      const int i = i0 + tid;
      // the K closest obstacles, ascending by (masked distance, obstacle index)  (MPPI.py:243-247)
      const float* mr = m_rows + (size_t)i * M;
      float last_v = -3.4e38f;
      int last_j = -1;
      for (int kk = 0; kk < K; ++kk) {
        float bv = 3.4e38f;
        int bj = -1;
        for (int j = 0; j < M; ++j) {
          const float v = mr[j];
          const bool after = kk == 0 || v > last_v || (v == last_v && j > last_j);
          if (after && (bj < 0 || v < bv)) { bv = v; bj = j; }
        }
        if (bj < 0) bj = last_j < 0 ? 0 : last_j;
        sel_rows[(size_t)i * K + kk] = i * M + bj;
        last_v = bv; last_j = bj;
      }
      StepArgs s = sa;
      s.t = t;
      step_sample(s, i);
    }
    __syncthreads();                                    // next state written before the next tile reads it
  }
}

constexpr size_t smem_bytes(int R) {
  return (size_t)(HID * R + NSTAGE * WS * HID + R * XS + R * MAXO + R + R) * sizeof(float);
}

// Tile shape (rows per thread, features per thread).  Measured on B200 (tools/exact_shape_sweep.py, ms per
// propagate, planar-7 net, M = 4, H = 30; shapes as RPTxFT):
//     rows per step   8x8     4x8     8x4     4x4     8x2     4x2
//        1 000        5.52    4.20    4.79    3.77    5.61    4.13
//        4 000        5.54    5.06    4.86    5.80    5.68    8.19
//        8 000        7.63    7.91    7.84   10.64   11.17   16.19
//       16 000       13.84   13.73   15.29   18.57   22.13   28.28
// While the batch gives an SM at most one 32-row tile, the 8-warp shapes win: two warps per scheduler hide part of
// the shared-memory / FMA latency a lone 4-warp CTA exposes (ncu: 27 % -> 39 % issue-active), 16-row tiles while even
// those leave SMs idle.  16 warps (FT = 2) lose again: 2-feature register tiles need a load per two FFMA2.  From two
// tiles per SM on, the 4-warp 8 x 8 register tile is best (fewest shared-memory loads per FMA; 87 % of the FMA
// peak in steady state with 4 CTAs per SM).
struct Shape { int rpt, ft; };
Shape pick_shape(const dsmppi_ctx* c, long long rows, int min_rows_per_tile) {
  const long long sms = c->sm_count;
  const long long tiles32 = (rows + 31) / 32, tiles16 = (rows + 15) / 16;
  Shape s = {8, 8};
  if (tiles32 <= sms) s = {(min_rows_per_tile <= 16 && tiles16 <= sms) ? 4 : 8, 4};
  const char* fr = std::getenv("DSMPPI_EXACT_RPT");
  const char* ff = std::getenv("DSMPPI_EXACT_FT");
  if (ff) { const int f = std::atoi(ff); if (f == 8 || f == 4) s.ft = f; }
  if (fr) { const int f = std::atoi(fr); if (f == 8 || (f == 4 && min_rows_per_tile <= 16)) s.rpt = f; }
  if (s.ft == 8) s.rpt = 8;
  return s;
}

template <bool BWD, int RPT, int FT>
int launch_variant(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask, float* out_m,
                   float* out_dist, float* out_grad, cudaStream_t st) {
  constexpr int R = 4 * RPT;
  static bool init = false;
  if (!init) {
    CUDA_TRY(cudaFuncSetAttribute(exact_mlp_kernel<BWD, RPT, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_bytes(R)));
    init = true;
  }
  const long long grid = ((long long)src.n_rows + R - 1) / R;
  exact_mlp_kernel<BWD, RPT, FT><<<(unsigned)grid, nthreads(FT), smem_bytes(R), st>>>(
      c->net, src, q, q_stride, c->obs, ignore_mask, out_m, out_dist, out_grad);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return 0;
}

#define SHAPE_DISPATCH(sh, CALL)                                  \
  do {                                                            \
    if (sh.ft == 8) return CALL(8, 8);                            \
    if (sh.rpt == 8) return CALL(8, 4);                           \
    return CALL(4, 4);                                            \
  } while (0)

template <bool BWD>
int launch(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask, float* out_m,
           float* out_dist, float* out_grad, long long rows_estimate, cudaStream_t st) {
  if (src.n_rows <= 0) return 0;
  if (rows_estimate <= 0 || rows_estimate > src.n_rows) rows_estimate = src.n_rows;
  const Shape sh = pick_shape(c, rows_estimate, 1);
#define CALL(RPT, FT) launch_variant<BWD, RPT, FT>(c, q, q_stride, src, ignore_mask, out_m, out_dist, out_grad, st)
  SHAPE_DISPATCH(sh, CALL);
#undef CALL
}

template <int RPT, int FT>
int launch_fused_variant(dsmppi_ctx* c, const dsmppi_rollout_args* a, cudaStream_t st) {
  constexpr int R = 4 * RPT;
  static bool init = false;
  if (!init) {
    CUDA_TRY(cudaFuncSetAttribute(rollout_fused_kernel<RPT, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_bytes(R)));
    init = true;
  }
  const int S = R / c->M;
  const long long grid = ((long long)a->N + S - 1) / S;
  const StepArgs sa = make_step_args(c, a, 0);
  rollout_fused_kernel<RPT, FT><<<(unsigned)grid, nthreads(FT), smem_bytes(R), st>>>(
      c->net, sa, c->obs, c->M, a->ignored_link_mask, c->m_rows, c->row_dist, c->row_grad, c->sel_rows);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return 0;
}

}  // namespace

// whole-horizon single launch; the caller has checked M <= 32 and initialised all_traj[:, 0]
int launch_rollout_fused(dsmppi_ctx* c, const dsmppi_rollout_args* a, cudaStream_t st) {
  const int M = c->M;
  // a CTA holds whole samples: S = R / M of them, so the row count that matters is N * (R / S) ~ N * M
  const Shape sh = pick_shape(c, (long long)a->N * M, M);
#define CALL(RPT, FT) launch_fused_variant<RPT, FT>(c, a, st)
  SHAPE_DISPATCH(sh, CALL);
#undef CALL
}

int launch_exact_fixup(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask,
                       float* m_rows, float* row_dist, float* row_grad, bool bwd, cudaStream_t st) {
  constexpr int RPT = 8, FT = 8, R = 4 * RPT;
  static bool init = false;
  if (!init) {
    CUDA_TRY(cudaFuncSetAttribute(exact_mlp_kernel<true, RPT, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_bytes(R)));
    CUDA_TRY(cudaFuncSetAttribute(exact_mlp_kernel<false, RPT, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_bytes(R)));
    init = true;
  }
  long long grid = ((long long)src.n_rows + R - 1) / R;
  if (grid > 4LL * c->sm_count) grid = 4LL * c->sm_count;       // usually no row is flagged: keep the launch tiny
  if (bwd)
    exact_mlp_kernel<true, RPT, FT><<<(unsigned)grid, nthreads(FT), smem_bytes(R), st>>>(
        c->net, src, q, q_stride, c->obs, ignore_mask, m_rows, row_dist, row_grad);
  else
    exact_mlp_kernel<false, RPT, FT><<<(unsigned)grid, nthreads(FT), smem_bytes(R), st>>>(
        c->net, src, q, q_stride, c->obs, ignore_mask, m_rows, nullptr, nullptr);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return 0;
}

int launch_exact_forward(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask,
                         float* m_rows, cudaStream_t st) {
  if (use_tc_scoring(c)) return launch_tc_exact(c, q, q_stride, src, ignore_mask, m_rows, nullptr, nullptr, false, st);
  return launch<false>(c, q, q_stride, src, ignore_mask, m_rows, nullptr, nullptr, 0, st);
}

int launch_exact_fwdbwd(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask,
                        float* m_rows, float* row_dist, float* row_grad, long long rows_estimate, cudaStream_t st) {
  if (use_tc_scoring(c)) return launch_tc_exact(c, q, q_stride, src, ignore_mask, m_rows, row_dist, row_grad, true, st);
  return launch<true>(c, q, q_stride, src, ignore_mask, m_rows, row_dist, row_grad, rows_estimate, st);
}
