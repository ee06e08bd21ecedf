// Device microbenchmarks used to pin roofline denominators on the box.
#include "common.cuh"

namespace qb {

// register-resident DMMA loop: 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters) {
  double a[4], b[2], c[8][4];
  const double s = 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = 1.0 + s * i;
  b[0] = 0.5 + s; b[1] = 0.25 - s;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int v = 0; v < 4; ++v) c[j][v] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma_16x8x8(c[j], a, b);
  }
  double acc = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc += c[j][v];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace qb

using namespace qb;

extern "C" int qb_measure_dmma_peak(double *tflops, void *stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = sm_count() * 2, threads = 256, iters = 4096;
  double *buf = nullptr;
  QB_CUDA_CHECK(cudaMalloc(&buf, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  QB_CUDA_CHECK(cudaEventCreate(&e0));
  QB_CUDA_CHECK(cudaEventCreate(&e1));
  dmma_peak_kernel<<<blocks, threads, 0, st>>>(buf, 64);  // warm-up
  QB_LAUNCH_CHECK();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    QB_CUDA_CHECK(cudaEventRecord(e0, st));
    dmma_peak_kernel<<<blocks, threads, 0, st>>>(buf, iters);
    QB_LAUNCH_CHECK();
    QB_CUDA_CHECK(cudaEventRecord(e1, st));
    QB_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    QB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  // flops: per warp per iter 8 mma x (16*8*8*2)
  double flops = (double)blocks * (threads / 32) * iters * 8.0 * 2048.0;
  *tflops = flops / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return 0;
}
