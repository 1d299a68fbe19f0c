// microbench.cu — FP64 issue-rate microbenchmark (measurement aid, not on the path): the non-FMA double-precision peak the
// episode kernel's arithmetic could reach at most, since the path is compiled with --fmad=false for bit-parity with the
// reference's separate multiply and add.
#include <cuda_runtime.h>
#include "../../include/eirgrid_b200.h"
#include "common.hpp"

namespace {

__global__ void __launch_bounds__(256) eg_fp64_rate_kernel(double* out, int iters, double a, double b) {
  // 8 independent chains per thread, each alternating a multiply and an add (no contraction: --fmad=false)
  double x0 = threadIdx.x * 1e-3 + 1.0, x1 = x0 + 0.1, x2 = x0 + 0.2, x3 = x0 + 0.3, x4 = x0 + 0.4, x5 = x0 + 0.5, x6 = x0 + 0.6, x7 = x0 + 0.7;
#pragma unroll 4
  for (int i = 0; i < iters; i++) {
    x0 = x0 * a; x1 = x1 * a; x2 = x2 * a; x3 = x3 * a; x4 = x4 * a; x5 = x5 * a; x6 = x6 * a; x7 = x7 * a;
    x0 = x0 + b; x1 = x1 + b; x2 = x2 + b; x3 = x3 + b; x4 = x4 + b; x5 = x5 + b; x6 = x6 + b; x7 = x7 + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace

extern "C" int eg_microbench_fp64(int device, double* tflops_out) {
  if (!tflops_out) return eg_fail(EG_ERR_INVALID, "eg_microbench_fp64: NULL argument");
  if (cudaSetDevice(device) != cudaSuccess) return eg_fail(EG_ERR_NO_DEVICE, "eg_microbench_fp64: no such device");
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, threads = 256, iters = 1 << 14;
  double* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(double) * blocks * threads) != cudaSuccess) return eg_fail(EG_ERR_CUDA, "eg_microbench_fp64: cudaMalloc failed");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  eg_fp64_rate_kernel<<<blocks, threads>>>(d, 256, 0.999999, 1e-7);  // warm-up
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    eg_fp64_rate_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (err != cudaSuccess) return eg_fail(EG_ERR_CUDA, cudaGetErrorString(err));
  *tflops_out = (double)blocks * threads * iters * 16.0 / (best * 1e-3) / 1e12;
  return EG_OK;
}
