// How should the live policy columns of a 10^6-sample batch cross PCIe?  (field_1m e2e, dsmppi_iteration_host.)
// The caller's tensors are (N, 50, d) / (N, 50) pinned host rows of which only the first nk kernels are live, so the
// copy is strided: 80 of every 400 bytes (mu, alpha; d = 2, nk = 10) and 40 of every 200 (sigma).
//   a) cudaMemcpy2DAsync (what round 1 shipped)       b) a gather kernel reading the pinned rows through UVA
//   c) one contiguous cudaMemcpyAsync of the same byte count (the PCIe ceiling)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/h2d_gather_microbench.cu -o /tmp/h2d_gather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// every thread moves 16 bytes; `live` and `pitch` in float4 units... rows are 16-byte aligned when live*4 % 16 == 0
__global__ void gather_rows(const float4* __restrict__ src, float4* __restrict__ dst, long long n_rows, int live4,
                            int pitch4) {
  const long long total = n_rows * live4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / live4;
    const int c = (int)(i - r * live4);
    dst[r * pitch4 + c] = src[r * pitch4 + c];
  }
}
__global__ void gather_rows2(const float2* __restrict__ src, float2* __restrict__ dst, long long n_rows, int live2,
                             int pitch2) {
  const long long total = n_rows * live2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / live2;
    const int c = (int)(i - r * live2);
    dst[r * pitch2 + c] = src[r * pitch2 + c];
  }
}

int main() {
  const long long N = 1000000;
  const int d = 2, NK = 50, nk = 10;
  const size_t mu_b = N * NK * d * 4, sg_b = N * NK * 4;
  float *h_mu, *h_al, *h_sg, *d_mu, *d_al, *d_sg, *h_flat, *d_flat;
  CK(cudaMallocHost(&h_mu, mu_b)); CK(cudaMallocHost(&h_al, mu_b)); CK(cudaMallocHost(&h_sg, sg_b));
  CK(cudaMalloc(&d_mu, mu_b)); CK(cudaMalloc(&d_al, mu_b)); CK(cudaMalloc(&d_sg, sg_b));
  const size_t live_b = N * nk * (2 * d + 1) * 4;
  CK(cudaMallocHost(&h_flat, live_b)); CK(cudaMalloc(&d_flat, live_b));
  for (size_t i = 0; i < mu_b / 4; ++i) { h_mu[i] = (float)i; h_al[i] = -(float)i; }
  for (size_t i = 0; i < sg_b / 4; ++i) h_sg[i] = 0.5f * i;
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto fn) {
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
      CK(cudaEventRecord(e0, st)); fn(); CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("%-52s %7.3f ms  %6.1f GB/s of live bytes\n", name, best, live_b / best * 1e-6);
  };
  timeit("contiguous cudaMemcpyAsync (same bytes)", [&] { CK(cudaMemcpyAsync(d_flat, h_flat, live_b, cudaMemcpyHostToDevice, st)); });
  timeit("cudaMemcpy2DAsync x3 (strided rows)", [&] {
    CK(cudaMemcpy2DAsync(d_mu, NK * d * 4, h_mu, NK * d * 4, nk * d * 4, N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(d_al, NK * d * 4, h_al, NK * d * 4, nk * d * 4, N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(d_sg, NK * 4, h_sg, NK * 4, nk * 4, N, cudaMemcpyHostToDevice, st));
  });
  for (int blocks : {148, 592, 2368}) for (int threads : {256, 1024}) {
    char name[96]; snprintf(name, sizeof name, "gather kernel over pinned rows, %d x %d", blocks, threads);
    timeit(name, [&] {
      gather_rows<<<blocks, threads, 0, st>>>((const float4*)h_mu, (float4*)d_mu, N, nk * d / 4, NK * d / 4);
      gather_rows<<<blocks, threads, 0, st>>>((const float4*)h_al, (float4*)d_al, N, nk * d / 4, NK * d / 4);
      gather_rows2<<<blocks, threads, 0, st>>>((const float2*)h_sg, (float2*)d_sg, N, nk / 2, NK / 2);
    });
  }
  // one fused launch with all three tensors would hide the tails; the three above already overlap poorly, so also try
  // the three kernels on three streams
  cudaStream_t s2, s3; CK(cudaStreamCreate(&s2)); CK(cudaStreamCreate(&s3));
  cudaEvent_t j2, j3, f0; CK(cudaEventCreate(&j2)); CK(cudaEventCreate(&j3)); CK(cudaEventCreate(&f0));
  timeit("gather kernels on three streams, 592 x 256", [&] {
    CK(cudaEventRecord(f0, st)); CK(cudaStreamWaitEvent(s2, f0, 0)); CK(cudaStreamWaitEvent(s3, f0, 0));
    gather_rows<<<592, 256, 0, st>>>((const float4*)h_mu, (float4*)d_mu, N, nk * d / 4, NK * d / 4);
    gather_rows<<<592, 256, 0, s2>>>((const float4*)h_al, (float4*)d_al, N, nk * d / 4, NK * d / 4);
    gather_rows2<<<592, 256, 0, s3>>>((const float2*)h_sg, (float2*)d_sg, N, nk / 2, NK / 2);
    CK(cudaEventRecord(j2, s2)); CK(cudaEventRecord(j3, s3)); CK(cudaStreamWaitEvent(st, j2, 0)); CK(cudaStreamWaitEvent(st, j3, 0));
  });
  // D2H side: kernel_val_all (N*H, 50) -> nk live columns, strided the same way
  timeit("D2H cudaMemcpy2DAsync kernel_val (40 of 200 B)", [&] {
    CK(cudaMemcpy2DAsync(h_sg, NK * 4, d_sg, NK * 4, nk * 4, N, cudaMemcpyDeviceToHost, st));
  });
  timeit("D2H scatter kernel kernel_val, 592 x 256", [&] {
    gather_rows2<<<592, 256, 0, st>>>((const float2*)d_sg, (float2*)h_sg, N, nk / 2, NK / 2);
  });
  timeit("D2H contiguous 40 MB", [&] { CK(cudaMemcpyAsync(h_flat, d_flat, (size_t)N * nk * 4, cudaMemcpyDeviceToHost, st)); });
  // check the gather result
  float probe[4];
  CK(cudaMemcpy(probe, d_mu + 999999ll * NK * d + 16, 16, cudaMemcpyDeviceToHost));
  printf("probe %g (expect %g)\n", probe[0], h_mu[999999ll * NK * d + 16]);
  return 0;
}
