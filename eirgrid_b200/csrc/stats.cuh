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

// 1 launch: [score | global id | eg_result | eg_traj] of the shard's best episode as one flat record
cudaError_t eg_launch_pack_best(const eg_result* results, const eg_traj* trajs, uint32_t n, const double* best_score,
                                const unsigned long long* best_index, unsigned long long first_global, void* record, cudaStream_t stream);

// 3 launches: reset of the best slot, accumulation + max score, lowest index holding the max
cudaError_t eg_launch_stats(const EgStatsParams& p, cudaStream_t stream);
