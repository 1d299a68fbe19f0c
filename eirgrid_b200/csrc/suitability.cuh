// suitability.cuh — location suitability analysis kernel (suitability.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct EgSuitabilityParams {
  int half;            // grid spans i, j in [-half, half]
  double step;         // metres between analysis points (reference: 2 * GRID_CELL_SIZE)
  uint32_t first, n;   // point range [first, first + n)
  int n_settlements;
  const double* sx;
  const double* sy;
  const uint32_t* pop;
  int n_generators;
  const double* gx;
  const double* gy;
  int n_coast;
  const double* cx;
  const double* cy;
  double* scores;      // [n][15]
};

cudaError_t eg_launch_suitability(const EgSuitabilityParams& p, cudaStream_t stream);
