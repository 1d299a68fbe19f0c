// microbench.cu — FP64 issue-rate microbenchmark (measurement aid, not on the path): the non-FMA double-precision peak the
// episode kernel's arithmetic could reach at most, since the path is compiled with --fmad=false for bit-parity with the
// reference's separate multiply and add.
#include <cuda_runtime.h>
#include "../../include/eirgrid_b200.h"
#include "common.hpp"
#include "update_rule.hpp"

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

__host__ __device__ inline double rule_math(uint32_t fn, double x, double y) {
  switch (fn) {
    case 0: return egm::exp(x);
    case 1: return egm::log(x);
    case 2: return egm::pow(x, y);
    case 3: {  // score_metrics of (net emissions x, opinion 0.5, total cost y, reliability 1)
      const double m[4] = {x, 0.5, y, 1.0};
      return egrule::score(m, false);
    }
    case 4: return egrule::contrast(2.0, x, (uint32_t)y, 0.2).penalty;
    case 5: return egrule::contrast(2.0, x, (uint32_t)y, 0.2).boost;
    case 6: return egrule::contrast(2.0, x, (uint32_t)y, 0.2).mild;
    case 7: return egrule::deficit_contrast((uint32_t)y, 0.2).penalty;
    case 8: return egrule::deficit_contrast((uint32_t)y, 0.2).boost;
    case 9: return egrule::update_uniform((uint32_t)x, 7u, (uint32_t)y, 3u);
  }
  return 0.0;
}

__global__ void eg_rule_math_kernel(uint32_t fn, const double* x, const double* y, uint32_t n, double* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rule_math(fn, x[i], y[i]);
}

}  // namespace

// Test aid: the update rule's shared arithmetic (csrc/eg_math.hpp, csrc/update_rule.hpp) evaluated on the host
// (device < 0) or on a GPU. fn: 0 exp(x), 1 ln(x), 2 pow(x, y), 3 score, 4-6 contrast penalty / boost / mild for
// (best 2.0, current x, stagnation counter y), 7-8 deficit penalty / boost for counter y, 9 a draw of the update stream.
extern "C" int eg_rule_math(int device, uint32_t fn, const double* x, const double* y, uint32_t n, double* out) {
  if (!x || !y || !out) return eg_fail(EG_ERR_INVALID, "eg_rule_math: NULL argument");
  if (device < 0) {
    for (uint32_t i = 0; i < n; i++) out[i] = rule_math(fn, x[i], y[i]);
    return EG_OK;
  }
  if (cudaSetDevice(device) != cudaSuccess) return eg_fail(EG_ERR_NO_DEVICE, "eg_rule_math: no such device");
  double *dx = nullptr, *dy = nullptr, *dout = nullptr;
  const size_t bytes = (size_t)(n ? n : 1) * sizeof(double);
  if (cudaMalloc((void**)&dx, bytes) != cudaSuccess || cudaMalloc((void**)&dy, bytes) != cudaSuccess || cudaMalloc((void**)&dout, bytes) != cudaSuccess) {
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return eg_fail(EG_ERR_CUDA, "eg_rule_math: cudaMalloc failed");
  }
  cudaMemcpy(dx, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(dy, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (n) eg_rule_math_kernel<<<(n + 255) / 256, 256>>>(fn, dx, dy, n, dout);
  cudaMemcpy(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
  const cudaError_t err = cudaGetLastError();
  cudaFree(dx); cudaFree(dy); cudaFree(dout);
  if (err != cudaSuccess) return eg_fail(EG_ERR_CUDA, cudaGetErrorString(err));
  return EG_OK;
}

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
