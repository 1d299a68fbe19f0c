// update.cuh — the reference's per-episode weight update in episode order, on the device (update.cu).
#pragma once
#include <cuda_runtime.h>
#include "tables.h"

#define EG_UPD_CHUNK 16384          // episodes per pass of the pipeline (a longer batch is a sequence of passes)
#define EG_UPD_ROW 80               // bytes per (episode, year) row of multiply counts: 61 action keys, 15 deficit keys, pad
#define EG_UPD_ENTRIES 76            // table entries per year: 61 action keys followed by 15 deficit keys
#define EG_UPD_CAT_CAPACITY 3072    // best_actions ++ best_deficit_actions of all years of ONE strategy: a replay record holds
                                    // at most 2 x 984 + 984 + 4 x 26 entries
#define EG_UPD_IMP_INLINE 256       // improvement records copied to the host with the state (more: a second copy)

// Device-resident mirror of the eg_weights fields the rule reads and writes.
struct EgUpdState {
  double w[EG_NY][EG_N_ACTIONS];
  double dw[EG_NY][EG_N_DEFICIT_KEYS];
  double best_w[EG_NY][EG_N_ACTIONS];   // table at the moment of the last improvement (ActionWeights.best_weights)
  double best_metrics[4];
  double best_score;
  double learning_rate;
  double batch_best_score;
  long long batch_best_index;           // first episode of the call holding batch_best_score, -1 if none
  uint32_t has_best, iwi, iteration_count;
  uint32_t n_improvements, n_applied, n_flagged;
  // values at the start of the pass in flight (the scan advances the counters above before the other kernels read them)
  uint32_t pass_has_best, pass_iwi, pass_iteration_count;
  int32_t pass_last_improver;           // last improving episode of the pass (index inside the pass), -1 if none
  uint32_t pass_any_random;             // some episode of the pass takes the randomisation branch
  uint32_t cost_only;                   // ActionWeights.optimization_mode == "cost_only" (score_metrics' mode, quirk Q12)
};

// One best strategy as the rule compares against it: per year best_actions ++ best_deficit_actions.
struct EgUpdSlot {
  uint16_t len_run[EG_NY];              // len(best_actions[y])
  uint16_t len_def[EG_NY];              // len(best_deficit_actions[y])
  uint32_t off[EG_NY];                  // start of year y in cat[]
  unsigned long long mask_all[EG_NY];   // action codes that occur in the concatenation
  unsigned long long mask_def[EG_NY];   // action codes that occur in best_deficit_actions[y]
  uint8_t cat[EG_UPD_CAT_CAPACITY];
};

struct EgUpdImprovement {   // one entry of improvement_history (strategy.rs:72-84) as the host needs it
  long long episode;        // index inside the call
  double score;
  double metrics[4];
};

struct EgUpdPre {           // per episode, from the scan over the scores
  double best_before;       // score of the best strategy when the episode's update starts
  int32_t improver;         // episode of the pass that holds it, -1: the strategy the pass started with
  uint32_t iwi_before;
  uint32_t improved;
  uint32_t pad;
};

struct EgUpdCtl {           // per episode, what the table walk needs; 64 bytes
  double boost, penalty, mild, d_boost, d_penalty;
  uint32_t flags;           // EG_UPD_*
  uint32_t iteration;       // ActionWeights.iteration_count when the episode's update starts (Philox counter)
  uint32_t pad[4];
};
#define EG_UPD_APPLIED 1u
#define EG_UPD_RANDOMISE 2u
#define EG_UPD_D_APPLIED 4u
#define EG_UPD_D_RANDOMISE 8u
#define EG_UPD_IMPROVED 16u
// facts about the episode's factors that let the walk skip its repeat handling (valid for entries inside [MIN_WEIGHT, MAX_WEIGHT]):
// the first boost already saturates (MIN_WEIGHT * boost >= MAX_WEIGHT, boost >= 1), the first penalty already hits the floor
// (MAX_WEIGHT * p <= MIN_WEIGHT, or p is NaN: quirk Q9)
#define EG_UPD_BOOST_SAT 32u
#define EG_UPD_PEN_KILL 64u
#define EG_UPD_MILD_KILL 128u
#define EG_UPD_D_BOOST_SAT 256u
#define EG_UPD_D_PEN_KILL 512u
#define EG_UPD_PROPER 1024u     // boost >= 1 and no penalty factor above 1: the episode's steps keep an entry inside the clamp range
#define EG_UPD_D_PROPER 2048u

struct EgUpdBuffers {       // device scratch of one context, sized for EG_UPD_CHUNK episodes
  EgUpdState* state;
  EgUpdSlot* slots;         // [EG_UPD_CHUNK + 1]: slot 0 = strategy at the start of the pass, slot e + 1 = episode e's record
  double* score;            // [EG_UPD_CHUNK]
  EgUpdPre* pre;            // [EG_UPD_CHUNK]
  EgUpdCtl* ctl;            // [EG_UPD_CHUNK]
  uint8_t* counts;          // [EG_NY][EG_UPD_CHUNK][EG_UPD_ROW]
  double* factors;          // [EG_NY][EG_UPD_CHUNK][EG_UPD_ENTRIES] factors of the randomisation branch
  EgUpdImprovement* improvements;  // [capacity]
  uint32_t improvements_capacity;
};

// Queues one pass (n <= EG_UPD_CHUNK episodes, `base` = index of the first one inside the call) on `stream`;
// *launches receives the number of kernels launched.
cudaError_t eg_launch_update_pass(const EgUpdBuffers& b, const eg_result* d_results, const eg_traj* d_trajs, uint32_t n,
                                  uint32_t base, uint32_t replay, uint64_t rng_seed, cudaStream_t stream, int* launches);
