// stats.cuh — batch-synchronous update statistics (stats.cu).
#pragma once
#include <cuda_runtime.h>
#include "tables.h"
#include "weights.hpp"

struct EgStatsParams {
  EgContrastConsts consts;
  const EgPolicyDevice* policy;   // snapshot the batch was sampled from (best lists)
  const eg_result* results;
  const eg_traj* trajs;
  uint32_t n;
  int64_t* stats;                 // [EG_STATS_WORDS], accumulated
  double* best_score;             // [1]
  unsigned long long* best_index; // [1]
  double ln100;
};

// 3 launches: reset of the best slot, accumulation + max score, lowest index holding the max
cudaError_t eg_launch_stats(const EgStatsParams& p, cudaStream_t stream);
