// fp32-exact evaluation of the learned distance network on (sample, obstacle) rows.
//
//   exact_mlp_kernel<false>: forward only  -> masked minimum link distance per row  (MPPI.py:235-242)
//   exact_mlp_kernel<true> : forward + analytic VJP at argmin_l of the raw output   (robot_sdf.py:153-158)
//
// One CTA owns 64 rows.  Activations live in shared memory feature-major (act[k][row]) and are updated in
// place layer by layer; each of the 256 threads owns an 8-row x 8-feature register tile, so a layer is
// 256 rank-1 updates of 64 FFMAs.  ReLU masks stay in registers (the forward and backward tilings
// coincide), so the backward pass needs no extra memory.  All arithmetic is IEEE fp32 (no fast-math):
// this is the path that has to agree with the reference's torch-CPU numbers to ~1e-6.
#include "internal.cuh"

namespace {

constexpr int R = 64;     // rows per CTA
constexpr int NT = 256;   // threads per CTA
constexpr int XS = 12;    // padded row stride of the raw-input scratch (nin <= 11)

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

// acc[i][j] = sum_k act[k][rg*8+i] * W[k*256 + fg*8+j]
__device__ __forceinline__ void gemm_tile(const float* __restrict__ W, int K, const float (*act)[R], int rg, int fg,
                                          float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float4* wp = reinterpret_cast<const float4*>(W + fg * 8);
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(&act[k][rg * 8]);
    const float4 a1 = *reinterpret_cast<const float4*>(&act[k][rg * 8 + 4]);
    const float4 w0 = __ldg(wp + k * (HID / 4));
    const float4 w1 = __ldg(wp + k * (HID / 4) + 1);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
}

__device__ __forceinline__ void store_tile(float (*act)[R], int rg, int fg, const float (&v)[8][8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<float4*>(&act[fg * 8 + j][rg * 8]) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
    *reinterpret_cast<float4*>(&act[fg * 8 + j][rg * 8 + 4]) = make_float4(v[4][j], v[5][j], v[6][j], v[7][j]);
  }
}

template <bool BWD>
__global__ void __launch_bounds__(NT, 2)
exact_mlp_kernel(NetDev net, RowSrc src, const float* __restrict__ q, int q_stride, const float* __restrict__ obs,
                 uint32_t ignore_mask, float* __restrict__ out_m, float* __restrict__ out_dist,
                 float* __restrict__ out_grad) {
  extern __shared__ __align__(16) float smem[];
  float (*act)[R] = reinterpret_cast<float (*)[R]>(smem);   // [256][R]
  float* xs = smem + HID * R;                               // [R][XS]  raw inputs x = [q, p]
  float* zs = xs + R * XS;                                  // [R][MAXO] raw outputs
  float* rad = zs + R * MAXO;                               // [R]
  int* lst = reinterpret_cast<int*>(rad + R);               // [R] argmin link

  const int n_rows = src.n_rows_dev ? min(*src.n_rows_dev, src.n_rows) : src.n_rows;
  const int row0 = blockIdx.x * R;
  if (row0 >= n_rows) return;
  const int tid = threadIdx.x;
  const int rg = tid >> 5;      // warp = group of 8 rows
  const int fg = tid & 31;      // lane = group of 8 features
  const int d = net.d, nin = net.nin, nenc = net.nenc, O = net.O;

  // ---- rows -> encoded inputs [x, sin x, cos x]  (network_macros_mod.py:139-140)
  if (tid < R) {
    int i = 0, j = 0;
    const bool valid = row_lookup(src, row0 + tid, n_rows, i, j);
    for (int c = 0; c < nin; ++c) {
      float x = 0.f;
      if (valid) x = (c < d) ? q[(size_t)i * q_stride + c] : obs[j * 4 + (c - d)];
      xs[tid * XS + c] = x;
      act[c][tid] = x;
      act[nin + c][tid] = sinf(x);
      act[2 * nin + c][tid] = cosf(x);
    }
    rad[tid] = valid ? obs[j * 4 + 3] : 0.f;
  }
  __syncthreads();

  uint32_t mk[4][2];
  float acc[8][8];
  // ---- hidden layers: h = relu(W h + b)
#pragma unroll 1
  for (int l = 0; l < 4; ++l) {
    gemm_tile(net.Wf[l], l == 0 ? nenc : HID, act, rg, fg, acc);
    __syncthreads();
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(net.b[l] + fg * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(net.b[l] + fg * 8) + 1);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t m0 = 0, m1 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = acc[i][j] + bb[j];
        const bool on = v > 0.f;
        if (BWD) {
          if (i < 4) m0 |= (on ? 1u : 0u) << (i * 8 + j);
          else m1 |= (on ? 1u : 0u) << ((i - 4) * 8 + j);
        }
        acc[i][j] = on ? v : 0.f;
      }
    mk[l][0] = m0;
    mk[l][1] = m1;
    store_tile(act, rg, fg, acc);
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
  for (int i = 0; i < 8; ++i) {
    const float* w = net.W4 + lst[rg * 8 + i] * HID + fg * 8;
    const uint32_t bits = (i < 4 ? mk[3][0] >> (i * 8) : mk[3][1] >> ((i - 4) * 8)) & 0xffu;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = ((bits >> j) & 1u) ? __ldg(w + j) : 0.f;
  }
  store_tile(act, rg, fg, acc);
  __syncthreads();

  // g_{l-1} = (W_l^T g_l) * s_{l-1},  l = 3, 2, 1   (Wb[l] is torch's [out][in]: out = k, in = n)
#pragma unroll 1
  for (int l = 3; l >= 1; --l) {
    gemm_tile(net.Wb[l], HID, act, rg, fg, acc);
    __syncthreads();
    const uint32_t m0 = mk[l - 1][0], m1 = mk[l - 1][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t bits = (i < 4 ? m0 >> (i * 8) : m1 >> ((i - 4) * 8)) & 0xffu;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = ((bits >> j) & 1u) ? acc[i][j] : 0.f;
    }
    store_tile(act, rg, fg, acc);
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

constexpr size_t kSmemBytes = (size_t)(HID * R + R * XS + R * MAXO + R + R) * sizeof(float);

template <bool BWD>
int launch(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask, float* out_m,
           float* out_dist, float* out_grad, cudaStream_t st) {
  static bool attr_set[2] = {false, false};
  if (!attr_set[BWD]) {
    CUDA_TRY(cudaFuncSetAttribute(exact_mlp_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSmemBytes));
    attr_set[BWD] = true;
  }
  if (src.n_rows <= 0) return 0;
  const int grid = (src.n_rows + R - 1) / R;
  exact_mlp_kernel<BWD><<<grid, NT, kSmemBytes, st>>>(c->net, src, q, q_stride, c->obs, ignore_mask, out_m,
                                                       out_dist, out_grad);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return 0;
}

}  // namespace

int launch_exact_forward(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask,
                         float* m_rows, cudaStream_t st) {
  return launch<false>(c, q, q_stride, src, ignore_mask, m_rows, nullptr, nullptr, st);
}

int launch_exact_fwdbwd(dsmppi_ctx* c, const float* q, int q_stride, const RowSrc& src, uint32_t ignore_mask,
                        float* m_rows, float* row_dist, float* row_grad, cudaStream_t st) {
  return launch<true>(c, q, q_stride, src, ignore_mask, m_rows, row_dist, row_grad, st);
}
