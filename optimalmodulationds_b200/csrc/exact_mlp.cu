// fp32-exact evaluation of the learned distance network on (sample, obstacle) rows.
//
//   exact_mlp_kernel<false, RPT>: forward only  -> masked minimum link distance per row  (MPPI.py:235-242)
//   exact_mlp_kernel<true,  RPT>: forward + analytic VJP at argmin_l of the raw output   (robot_sdf.py:153-158)
//
// One CTA of 4 warps owns R = 4*RPT rows (RPT = 8, 4 or 2).  Activations live in shared memory feature-major
// (act[k][row]) and are updated in place layer by layer.  A warp owns an R-row x 64-feature block of the layer
// output, its lanes form a 4 x 8 grid and each thread keeps an RPT-row x 8-feature register tile, so one k-step of a
// warp is 8*RPT FFMAs fed by single-wavefront shared-memory loads (broadcast over the row / feature groups).
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

namespace {

constexpr int NT = 128;   // threads per CTA: 4 warps, one per 64-feature quarter
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

// Thread tile: rows r0 .. r0+RPT-1, features {fa .. fa+3} (j = 0..3) and {fa+32 .. fa+35} (j = 4..7).
struct Tile {
  int r0, fa;
  __device__ __forceinline__ int feat(int j) const { return fa + (j & 3) + ((j >> 2) << 5); }
};

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
  int total;              // stages in the whole stream
  int issued;             // next stage to request

  __device__ __forceinline__ void locate(int stage, const float*& src, int& rows) const {
    int g, st;
    if (stage < n0) { g = 0; st = stage; }
    else { g = 1 + (stage - n0) / (HID / WS); st = (stage - n0) % (HID / WS); }
    const int K = g == 0 ? net->nenc : HID;
    const float* W = g < 4 ? net->Wf[g] : net->Wb[7 - g];
    src = W + (size_t)st * WS * HID;
    rows = min(WS, K - st * WS);
  }
  // every thread requests its share of the next stage (or nothing past the end) and closes one group
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
template <int RPT>
__device__ __forceinline__ void gemm_tile(WeightStream& ws, int& stage, int K, const float (*act)[4 * RPT], const Tile& t,
                                          float (&acc)[RPT][8]) {
  uint64_t acc2[RPT][4];                  // acc2[i][p] = (acc[i][2p], acc[i][2p+1])
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc2[i][p] = 0ull;
  for (int k0 = 0; k0 < K; k0 += WS, ++stage) {
    cp_async_wait<NSTAGE - 2>();          // this thread's share of `stage` has landed ...
    __syncthreads();                      // ... and everybody's; the buffer of stage-1 is free again
    ws.request_next();                    // refill it with stage + NSTAGE - 1
    const float* wb = ws.ring + (size_t)(stage % NSTAGE) * WS * HID + t.fa;
    const int rows = min(WS, K - k0);
    if (rows == WS) {
#pragma unroll
      for (int kk = 0; kk < WS; ++kk) {
        float a[RPT];
        load_rows<RPT>(&act[k0 + kk][t.r0], a);
        const float4 w0 = *reinterpret_cast<const float4*>(wb + kk * HID);
        const float4 w1 = *reinterpret_cast<const float4*>(wb + kk * HID + 32);
        const uint64_t wp[4] = {pack2f(w0.x, w0.y), pack2f(w0.z, w0.w), pack2f(w1.x, w1.y), pack2f(w1.z, w1.w)};
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const uint64_t ad = pack2f(a[i], a[i]);
#pragma unroll
          for (int p = 0; p < 4; ++p) acc2[i][p] = fma2(ad, wp[p], acc2[i][p]);
        }
      }
    } else {
      for (int kk = 0; kk < rows; ++kk) {
        float a[RPT];
        load_rows<RPT>(&act[k0 + kk][t.r0], a);
        const float4 w0 = *reinterpret_cast<const float4*>(wb + kk * HID);
        const float4 w1 = *reinterpret_cast<const float4*>(wb + kk * HID + 32);
        const uint64_t wp[4] = {pack2f(w0.x, w0.y), pack2f(w0.z, w0.w), pack2f(w1.x, w1.y), pack2f(w1.z, w1.w)};
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const uint64_t ad = pack2f(a[i], a[i]);
#pragma unroll
          for (int p = 0; p < 4; ++p) acc2[i][p] = fma2(ad, wp[p], acc2[i][p]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int p = 0; p < 4; ++p) unpack2f(acc2[i][p], acc[i][2 * p], acc[i][2 * p + 1]);
}

template <int RPT>
__device__ __forceinline__ void store_tile(float (*act)[4 * RPT], const Tile& t, const float (&v)[RPT][8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
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

template <bool BWD, int RPT>
__global__ void __launch_bounds__(NT, RPT == 8 ? 4 : (RPT == 4 ? 6 : 8))
exact_mlp_kernel(NetDev net, RowSrc src, const float* __restrict__ q, int q_stride, const float* __restrict__ obs,
                 uint32_t ignore_mask, float* __restrict__ out_m, float* __restrict__ out_dist,
                 float* __restrict__ out_grad) {
  constexpr int R = 4 * RPT;
  extern __shared__ __align__(16) float smem[];
  float (*act)[R] = reinterpret_cast<float (*)[R]>(smem);   // [256][R]
  float* ring = smem + HID * R;                             // [NSTAGE][WS][256] weight stages
  float* xs = ring + NSTAGE * WS * HID;                     // [R][XS]  raw inputs x = [q, p]
  float* zs = xs + R * XS;                                  // [R][MAXO] raw outputs
  float* rad = zs + R * MAXO;                               // [R]
  int* lst = reinterpret_cast<int*>(rad + R);               // [R] argmin link

  const int n_rows = src.n_rows_dev ? min(*src.n_rows_dev, src.n_rows) : src.n_rows;
  const int row0 = blockIdx.x * R;
  if (row0 >= n_rows) return;
  const int tid = threadIdx.x;
  Tile t;
  t.r0 = ((tid & 31) >> 3) * RPT;                     // lane / 8: one of four RPT-row groups
  t.fa = (tid >> 5) * 64 + (tid & 7) * 4;             // warp: 64-feature quarter; lane % 8: 4-feature group
  const int d = net.d, nin = net.nin, nenc = net.nenc, O = net.O;

  // start the weight stream before anything else: the first three stages fly while the inputs are encoded
  WeightStream ws;
  ws.net = &net;
  ws.ring = ring;
  ws.n0 = (nenc + WS - 1) / WS;
  ws.total = ws.n0 + (BWD ? 6 : 3) * (HID / WS);
  ws.issued = 0;
#pragma unroll
  for (int i = 0; i < NSTAGE - 1; ++i) ws.request_next();
  int stage = 0;

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

  uint64_t mk[4];          // ReLU masks of this thread's tile, bit i*8+j
  float acc[RPT][8];
  // ---- hidden layers: h = relu(W h + b)
#pragma unroll 1
  for (int l = 0; l < 4; ++l) {
    gemm_tile<RPT>(ws, stage, l == 0 ? nenc : HID, act, t, acc);
    __syncthreads();
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(net.b[l] + t.fa));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(net.b[l] + t.fa + 32));
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint64_t m = 0;
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = acc[i][j] + bb[j];
        const bool on = v > 0.f;
        if (BWD) m |= (uint64_t)(on ? 1u : 0u) << (i * 8 + j);
        acc[i][j] = on ? v : 0.f;
      }
    mk[l] = m;
    store_tile<RPT>(act, t, acc);
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

  if (!BWD) {
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
      out_m[row0 + tid] = m;
    }
    return;
  }

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
      out_dist[row0 + tid] = y - rad[tid];
      if (out_m) out_m[row0 + tid] = m;
    }
  }
  __syncthreads();

  // g4 = W5[l*, :] * s4
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const float* w = net.W4 + lst[t.r0 + i] * HID;
    const uint32_t bits = (uint32_t)(mk[3] >> (i * 8)) & 0xffu;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = ((bits >> j) & 1u) ? __ldg(w + t.feat(j)) : 0.f;
  }
  store_tile<RPT>(act, t, acc);
  __syncthreads();

  // g_{l-1} = (W_l^T g_l) * s_{l-1},  l = 3, 2, 1   (Wb[l] is torch's [out][in]: out = k, in = n)
#pragma unroll 1
  for (int l = 3; l >= 1; --l) {
    gemm_tile<RPT>(ws, stage, HID, act, t, acc);
    __syncthreads();
    const uint64_t m = mk[l - 1];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const uint32_t bits = (uint32_t)(m >> (i * 8)) & 0xffu;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = ((bits >> j) & 1u) ? acc[i][j] : 0.f;
    }
    store_tile<RPT>(act, t, acc);
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
    if (row0 + r < n_rows) out_grad[(size_t)(row0 + r) * d + c] = a0 + cosf(x) * a1 - sinf(x) * a2;
  }
}

constexpr size_t smem_bytes(int R) {
  return (size_t)(HID * R + NSTAGE * WS * HID + R * XS + R * MAXO + R + R) * sizeof(float);
}

// Rows-per-thread for a launch over about `rows` rows.  Measured on B200 (Franka shelf, ~24.6k candidate rows per
// step: 171.8 / 177.0 / 187.1 ms per iteration for RPT = 8 / 4 / 2; planar-7, 4000 rows: 6.32 / 6.02 / 6.81 ms):
// 32-row tiles win as soon as they give every SM two CTAs to overlap, 16-row tiles when rows are scarce; 8-row
// tiles re-read the weights from shared memory too often (LSU-bound) and are only kept for experiments.
int pick_rpt(const dsmppi_ctx* c, long long rows) {
  const char* force = std::getenv("DSMPPI_EXACT_RPT");
  if (force) { const int f = std::atoi(force); if (f == 8 || f == 4 || f == 2) return f; }
  return (rows + 31) / 32 >= 2LL * c->sm_count ? 8 : 4;
}

template <bool BWD, int RPT>
int launch_variant(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask, float* out_m,
                   float* out_dist, float* out_grad, cudaStream_t st) {
  constexpr int R = 4 * RPT;
  const long long grid = ((long long)src.n_rows + R - 1) / R;
  exact_mlp_kernel<BWD, RPT><<<(unsigned)grid, NT, smem_bytes(R), st>>>(c->net, src, q, q_stride, c->obs, ignore_mask,
                                                                       out_m, out_dist, out_grad);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return 0;
}

template <bool BWD>
int launch(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask, float* out_m,
           float* out_dist, float* out_grad, long long rows_estimate, cudaStream_t st) {
  static bool init[2] = {false, false};
  if (!init[BWD]) {
    CUDA_TRY(cudaFuncSetAttribute(exact_mlp_kernel<BWD, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(32)));
    CUDA_TRY(cudaFuncSetAttribute(exact_mlp_kernel<BWD, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(16)));
    init[BWD] = true;
  }
  if (src.n_rows <= 0) return 0;
  if (rows_estimate <= 0 || rows_estimate > src.n_rows) rows_estimate = src.n_rows;
  switch (pick_rpt(c, rows_estimate)) {
    case 8: return launch_variant<BWD, 8>(c, q, q_stride, src, ignore_mask, out_m, out_dist, out_grad, st);
    case 4: return launch_variant<BWD, 4>(c, q, q_stride, src, ignore_mask, out_m, out_dist, out_grad, st);
    default: return launch_variant<BWD, 2>(c, q, q_stride, src, ignore_mask, out_m, out_dist, out_grad, st);
  }
}

}  // namespace

int launch_exact_forward(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask,
                         float* m_rows, cudaStream_t st) {
  return launch<false>(c, q, q_stride, src, ignore_mask, m_rows, nullptr, nullptr, 0, st);
}

int launch_exact_fwdbwd(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask,
                        float* m_rows, float* row_dist, float* row_grad, long long rows_estimate, cudaStream_t st) {
  return launch<true>(c, q, q_stride, src, ignore_mask, m_rows, row_dist, row_grad, rows_estimate, st);
}
