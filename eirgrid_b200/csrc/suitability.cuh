// suitability.cuh — location suitability analysis kernel (suitability.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EG_SUIT_MAX_YEARS 26

struct EgSuitabilityParams {
  int mode;            // 0: analyze_map's grid, i, j in [-half, half]; 1: candidate-site grid, i, j in [0, side)
  int half;
  int side;
  double step;         // metres between points (analyze_map: 2 * GRID_CELL_SIZE)
  uint32_t first, n;   // point range [first, first + n)
  int year_first, n_years;  // simulated years 2025 + year_first .. (populations of those years)
  int n_settlements;
  const double* sx;    // device arrays padded to a multiple of 16 bytes (bulk copies)
  const double* sy;
  const uint32_t* pop;      // [26][n_settlements]
  const double* urban_r;    // [26][n_settlements] sqrt(pop) * 5, the radius of is_urban_area
  double urban_r_max;
  int n_generators;
  const double* gx;
  const double* gy;
  int n_coast;
  const double* cx;
  const double* cy;
  double* scores;      // [n][n_years][15]
};

// bytes of the per-site-row crossing lists the launch needs as scratch (`side` = points per axis of the grid)
size_t eg_suitability_rows_bytes(int side);
// 2 launches: the crossing lists of every site row, then the analysis itself
cudaError_t eg_launch_suitability(const EgSuitabilityParams& p, double* d_rows, cudaStream_t stream);
