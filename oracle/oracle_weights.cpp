// oracle_weights.cpp — CPU ORACLE (test infrastructure): the policy table.
// Restates ai/learning/weights/{core,sampling,learning,strategy,deficit}.rs, ai/learning/constants.rs and
// ai/metrics/scoring.rs of the reference. See oracle.hpp for scope, deviations and parity status.
#include "oracle.hpp"
#include <cmath>
#include <algorithm>
#include <numeric>

namespace orc {

// ai/learning/constants.rs
static const double MIN_WEIGHT = 0.0001, MAX_WEIGHT = 0.999, DEFAULT_WEIGHT = 0.5;
static const double MAX_ACCEPTABLE_EMISSIONS = 1000000.0, MAX_ACCEPTABLE_COST = 50000000000.0;

// ---- Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants) -----------------------------
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

Rng::Rng(uint64_t seed, uint64_t episode, uint32_t stream_) {
  key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
  ctr_hi[0] = (uint32_t)episode; ctr_hi[1] = (uint32_t)(episode >> 32);
  stream = stream_;
}
uint64_t Rng::next_u64() {
  uint32_t c[4] = {ctr_hi[0], ctr_hi[1], draw, stream}, o[4];
  draw++;
  philox4x32_10(c, key, o);
  return (uint64_t)o[0] | ((uint64_t)o[1] << 32);
}
double Rng::next_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }  // rand's 53-bit method
uint64_t Rng::next_index(uint64_t n) { return (uint64_t)(((unsigned __int128)next_u64() * n) >> 64); }

// ---- scoring.rs ---------------------------------------------------------------------------------------
double score_metrics(const Metrics& m, bool cost_only) {  // :5-45
  if (cost_only) {
    double normalized_cost = std::max(m.cost / MAX_ACCEPTABLE_COST, 1.0);
    double log_cost = std::log(normalized_cost);
    double max_expected_log_cost = std::log(MAX_ACCEPTABLE_COST * 100.0 / MAX_ACCEPTABLE_COST);
    return 2.0 - std::min(log_cost / max_expected_log_cost, 1.0);
  }
  if (m.net > 0.0) return 1.0 - std::min(m.net / MAX_ACCEPTABLE_EMISSIONS, 1.0);
  double base_score = 1.0;
  double normalized_cost = std::max(m.cost / MAX_ACCEPTABLE_COST, 1.0);
  double log_cost = std::log(normalized_cost);
  double max_expected_log_cost = std::log(MAX_ACCEPTABLE_COST * 100.0 / MAX_ACCEPTABLE_COST);
  double cost_score = 1.0 - std::min(log_cost / max_expected_log_cost, 1.0);
  double opinion_score = m.opinion;
  double cost_weight = normalized_cost > 8.0 ? 0.8 : 0.5;
  double opinion_weight = 1.0 - cost_weight;
  return base_score + (cost_score * cost_weight + opinion_score * opinion_weight);
}

double evaluate_action_impact(const ActionResult& cur, const ActionResult& nw, bool cost_only) {  // :46-84
  if (cost_only) {
    double cost_change = nw.cost - cur.cost;
    return -cost_change / std::max(std::fabs(cur.cost), 1.0);
  }
  if (cur.net > 0.0) return (cur.net - nw.net) / std::max(std::fabs(cur.net), 1.0);
  double cost_change = nw.cost - cur.cost;
  double cost_improvement = -cost_change / std::max(std::fabs(cur.cost), 1.0);
  double opinion_improvement = (nw.opinion - cur.opinion) / std::max(std::fabs(cur.opinion), 1.0);
  double cost_weight = cur.cost > MAX_ACCEPTABLE_COST * 8.0 ? 0.8 : 0.5;
  double opinion_weight = 1.0 - cost_weight;
  return cost_improvement * cost_weight + opinion_improvement * opinion_weight;
}

// ---- ActionWeights::new (core.rs:25-250) ------------------------------------------------------------------
static const int kDeficitKeyType[14] = {GasPeaker, GasCombinedCycle, BatteryStorage, PumpedStorage, Biomass,
                                        OnshoreWind, OffshoreWind, UtilitySolar, HydroDam, Nuclear,
                                        DomesticSolar, CommercialSolar, TidalGenerator, WaveEnergy};  // core.rs:130-149
uint8_t deficit_key_action(int k) { return k < 14 ? (uint8_t)(3 * kDeficitKeyType[k]) : (uint8_t)EG_ACT_DO_NOTHING; }
int action_deficit_key(uint8_t code) {
  if (code == EG_ACT_DO_NOTHING) return 14;
  if (code < 45 && code % 3 == 0)
    for (int k = 0; k < 14; k++)
      if (kDeficitKeyType[k] == code / 3) return k;
  return -1;
}

Weights::Weights() {
  static const double gw[15] = {0.08, 0.08, 0.05, 0.05, 0.08, 0.03, 0.04, 0.06, 0.02, 0.04, 0.06, 0.06, 0.07, 0.05, 0.05};  // constants.rs:46-60
  static const double dwk[15] = {0.15, 0.15, 0.15, 0.10, 0.10, 0.07, 0.07, 0.06, 0.06, 0.05, 0.01, 0.01, 0.01, 0.01, 0.001};  // :66-77
  for (int y = 0; y < NY; y++) {
    for (int t = 0; t < 15; t++) {
      w[y][3 * t + 0] = gw[t];
      w[y][3 * t + 1] = gw[t] * 0.5;
      w[y][3 * t + 2] = gw[t] * 0.25;
    }
    for (int o = 0; o < 4; o++) {
      w[y][45 + 3 * o + 0] = 0.02;
      w[y][45 + 3 * o + 1] = 0.02 * 0.5;
      w[y][45 + 3 * o + 2] = 0.02 * 0.25;
    }
    w[y][EG_ACT_UPGRADE] = 0.04;
    w[y][EG_ACT_ADJUST] = 0.04;
    w[y][EG_ACT_CLOSE] = 0.02;
    w[y][EG_ACT_DO_NOTHING] = 0.1;
    for (int k = 0; k < 15; k++) dw[y][k] = dwk[k];
    // action-count weights, core.rs:157-186
    double total_weight = 0.0;
    for (int count = 0; count <= 20; count++) {
      double base_weight = std::exp(-0.8 * (double)count);
      double multiplier = count == 0 ? 4.0 : count == 1 ? 3.5 : count == 2 ? 3.0 : count == 3 ? 2.5 : count == 4 ? 2.0 : count == 5 ? 1.5 : 1.0;
      double weight = base_weight * multiplier;
      cw[y][count] = weight;
      total_weight += weight;
    }
    for (int count = 0; count <= 20; count++) cw[y][count] /= total_weight;
  }
}

void Weights::clear_current_run() {  // core.rs clear_current_run_actions + clear_replay_index
  for (int y = 0; y < NY; y++) {
    current_run_actions[y].clear();
    current_deficit_actions[y].clear();
    replay_index[y] = 0;
    deficit_replay_index[y] = 0;
  }
}

// ---- sampling.rs --------------------------------------------------------------------------------------------
static const uint8_t kGasPeaker100 = 3 * GasPeaker;

static uint8_t smart_fallback_action(int year, Rng& rng) {  // :445-490
  struct P { uint8_t a; uint32_t w; };
  uint32_t storage = year < 2035 ? 10 : 20;
  uint32_t offset = year < 2035 ? 5 : (year < 2045 ? 15 : 25);
  uint32_t gas = year < 2035 ? 15 : (year < 2045 ? 10 : 5);
  P pool[7] = {{(uint8_t)(3 * OnshoreWind), 15}, {(uint8_t)(3 * OffshoreWind), 10}, {(uint8_t)(3 * UtilitySolar), 15},
               {(uint8_t)(3 * BatteryStorage), storage}, {(uint8_t)(45 + 3 * Forest), offset},
               {(uint8_t)(45 + 3 * ActiveCapture), offset}, {(uint8_t)(3 * GasCombinedCycle), gas}};
  uint32_t total = 0;
  for (auto& p : pool) total += p.w;
  uint32_t choice = (uint32_t)rng.next_index(total);
  for (auto& p : pool) {
    if (choice < p.w) return p.a;
    choice -= p.w;
  }
  return (uint8_t)(3 * BatteryStorage);
}

static uint8_t smart_deficit_fallback_action(Rng& rng) {  // :492-528
  struct P { uint8_t a; uint32_t w; };
  // (0.07*0.5) as u32 == 0 ; (0.06*0.5*100.0) as u32 == 3
  P pool[6] = {{(uint8_t)(3 * GasPeaker), 30}, {(uint8_t)(3 * BatteryStorage), 30}, {(uint8_t)(3 * GasCombinedCycle), 20},
               {(uint8_t)(3 * OnshoreWind), 10}, {(uint8_t)(3 * OffshoreWind), (uint32_t)(0.07 * 0.5)},
               {(uint8_t)(3 * UtilitySolar), (uint32_t)(0.06 * 0.5 * 100.0)}};
  uint32_t total = 0;
  for (auto& p : pool) total += p.w;
  uint32_t choice = (uint32_t)rng.next_index(total);
  for (auto& p : pool) {
    if (choice < p.w) return p.a;
    choice -= p.w;
  }
  return (uint8_t)(3 * BatteryStorage);
}

uint8_t sample_action(Weights& W, int year, Rng& rng) {  // :76-238
  const int y = year - BASE_YEAR;
  if (W.force_best_actions) {
    uint8_t action;
    if (W.has_best && W.replay_index[y] < W.best_actions[y].size()) {
      action = W.best_actions[y][W.replay_index[y]];
      W.replay_index[y]++;
    } else {
      action = smart_fallback_action(year, rng);
    }
    W.current_run_actions[y].push_back(action);  // :97-99 (the caller records it a second time, quirk Q10)
    return action;
  }
  const double* yw = W.w[y];
  double current_exploration = W.iwi > 100 ? W.exploration_rate * (1.0 / (1.0 + 0.01 * (double)W.iwi)) : W.exploration_rate;
  bool should_explore = rng.next_f64() < current_exploration;
  if (should_explore) return (uint8_t)rng.next_index(EG_N_ACTIONS);
  double total_weight = 0.0;
  for (int k = 0; k < EG_N_ACTIONS; k++) total_weight += yw[k];
  if (total_weight <= 0.0) return kGasPeaker100;
  if (W.iwi > 500) {
    int idx[EG_N_ACTIONS];
    std::iota(idx, idx + EG_N_ACTIONS, 0);
    std::stable_sort(idx, idx + EG_N_ACTIONS, [&](int a, int b) { return yw[a] > yw[b]; });
    double stagnation_factor = std::min((double)W.iwi / 1000.0, 3.0);
    double power_scaling = 1.0 + (2.0 * stagnation_factor);
    double total_scaled = 0.0;
    for (int k = 0; k < EG_N_ACTIONS; k++) total_scaled += std::pow(yw[idx[k]], power_scaling);
    double random_val = rng.next_f64() * total_scaled;
    for (int k = 0; k < EG_N_ACTIONS; k++) {
      random_val -= std::pow(yw[idx[k]], power_scaling);
      if (random_val <= 0.0) return (uint8_t)idx[k];
    }
    return (uint8_t)idx[0];
  }
  double random_val = rng.next_f64() * total_weight;
  for (int k = 0; k < EG_N_ACTIONS; k++) {
    random_val -= yw[k];
    if (random_val <= 0.0) return (uint8_t)k;
  }
  return kGasPeaker100;
}

uint8_t sample_deficit_action(Weights& W, int year, Rng& rng) {  // :240-378
  const int y = year - BASE_YEAR;
  if (W.force_best_actions) {
    uint8_t action;
    if (W.has_best && W.deficit_replay_index[y] < W.best_deficit_actions[y].size()) {
      action = W.best_deficit_actions[y][W.deficit_replay_index[y]];
      W.deficit_replay_index[y]++;
    } else {
      action = smart_deficit_fallback_action(rng);
    }
    W.current_deficit_actions[y].push_back(action);  // :262-264
    return action;
  }
  const double* yw = W.dw[y];
  bool should_explore = rng.next_f64() < W.exploration_rate;
  if (should_explore) return deficit_key_action((int)rng.next_index(14));  // AddGenerator keys only (:334-336)
  double total_weight = 0.0;
  for (int k = 0; k < 14; k++) total_weight += yw[k];
  if (total_weight <= 0.0) return kGasPeaker100;
  double random_val = rng.next_f64() * total_weight;
  for (int k = 0; k < 14; k++) {
    random_val -= yw[k];
    if (random_val <= 0.0) return deficit_key_action(k);
  }
  return kGasPeaker100;
}

uint32_t sample_additional_actions(Weights& W, int year, Rng& rng) {  // :380-443
  const int y = year - BASE_YEAR;
  uint32_t deficit_count = (uint32_t)W.current_deficit_actions[y].size();
  uint32_t max_possible = deficit_count >= 20 ? 0 : 20 - deficit_count;
  if (max_possible == 0) return 0;
  double random_val = rng.next_f64();
  if (W.has_count_weights) {
    double total_weight = 0.0;
    for (int c = 0; c < EG_N_COUNT_KEYS; c++) total_weight += W.cw[y][c];
    if (total_weight <= 0.0) return 0;
    double random_choice = random_val * total_weight;
    for (int c = 0; c < EG_N_COUNT_KEYS; c++) {
      random_choice -= W.cw[y][c];
      if (random_choice <= 0.0) return std::min((uint32_t)c, max_possible);
    }
    return std::min(5u, max_possible);
  }
  double scaled_exploration = std::pow(W.exploration_rate, 0.5);
  uint32_t min_actions = (uint32_t)std::round(2.0 / scaled_exploration);
  uint32_t max_actions = (uint32_t)std::round(12.0 / scaled_exploration);
  uint32_t capped_max = std::min(max_actions, max_possible);
  uint32_t capped_min = std::min(min_actions, capped_max);
  if (capped_min == capped_max) return capped_min;
  return capped_min + (uint32_t)rng.next_index((uint64_t)(capped_max - capped_min) + 1);
}

// ---- learning.rs:21-88 ---------------------------------------------------------------------------------------
void update_weights(Weights& W, uint8_t action, int year, double improvement) {
  double* yw = W.w[year - BASE_YEAR];
  double current_weight = yw[action];
  double final_impact_score = W.has_best ? score_metrics(W.best_metrics, W.cost_only_mode) : 0.0;
  double relative_improvement;
  if (W.has_best) {
    double best_score = score_metrics(W.best_metrics, W.cost_only_mode);
    relative_improvement = best_score > 0.0 ? (final_impact_score - best_score) / best_score : final_impact_score;
  } else {
    relative_improvement = final_impact_score;
  }
  double immediate_weight = relative_improvement > 0.0 ? 0.7 : 0.3;
  double combined = immediate_weight * improvement + (1.0 - immediate_weight) * relative_improvement;
  double adjustment = combined > 0.0 ? 1.0 + (W.learning_rate * combined) : 1.0 / (1.0 + (W.learning_rate * std::fabs(combined)));
  yw[action] = std::min(std::max(current_weight * adjustment, MIN_WEIGHT), MAX_WEIGHT);
  if (combined < 0.0) {
    double boost = 1.0 + (W.learning_rate * 0.1);
    for (int k = 0; k < 45; k++)
      if (k != action) yw[k] = std::min(yw[k] * boost, MAX_WEIGHT);
    if (W.has_best && W.best_metrics.net <= 0.0 && W.best_metrics.cost > MAX_ACCEPTABLE_COST * 8.0)
      yw[EG_ACT_DO_NOTHING] = std::min(yw[EG_ACT_DO_NOTHING] * (1.0 + W.learning_rate * 0.2), MAX_WEIGHT);
  }
}

// ---- deficit.rs:82-135 -----------------------------------------------------------------------------------------
void update_deficit_weights(Weights& W, uint8_t action, int year, double improvement) {
  int key = action_deficit_key(action);
  if (key < 0) return;  // unreachable: every deficit action is an AddGenerator(_, 100) key
  double* yw = W.dw[year - BASE_YEAR];
  double current_weight = yw[key];
  double adjustment = improvement > 0.0 ? 1.0 + (W.learning_rate * improvement * 1.5)
                                        : 1.0 / (1.0 + (W.learning_rate * std::fabs(improvement) * 1.5));
  yw[key] = std::min(std::max(current_weight * adjustment, MIN_WEIGHT), MAX_WEIGHT);
  if (improvement < 0.0) {
    double boost = 1.0 + (W.learning_rate * 0.1);
    for (int k = 0; k < 14; k++)
      if (k != key) yw[k] = std::min(yw[k] * boost, MAX_WEIGHT);
  }
}

static bool contains(const std::vector<uint8_t>& v, uint8_t a) { return std::find(v.begin(), v.end(), a) != v.end(); }

// ---- learning.rs:131-283 ----------------------------------------------------------------------------------------
void apply_contrast_learning(Weights& W, const Metrics& cur, Rng* rng) {
  if (!W.has_best) return;
  double best_score = score_metrics(W.best_metrics, W.cost_only_mode);
  double current_score = score_metrics(cur, W.cost_only_mode);
  double deterioration = best_score > 0.0 ? (best_score - current_score) / best_score : 0.0;
  double iterations = (double)W.iwi;
  double dynamic_threshold = 0.1 * std::max(std::exp(-iterations / 500.0), 0.00001 / 0.1);
  bool force_contrast = W.iwi > 800;
  if (!(deterioration > dynamic_threshold || force_contrast)) return;
  double stagnation_iterations = (double)W.iwi / 10.0;
  double stagnation_factor = 1.0 + (0.2 * std::pow(stagnation_iterations, 1.8));
  double scaled_deterioration = std::pow(deterioration, 0.3);
  double combined_penalty = scaled_deterioration * stagnation_factor;
  double alr = W.learning_rate * (1.0 + 0.1 * (double)W.iwi);
  double penalty_factor = 1.0 / (1.0 + alr * 1.5 * combined_penalty);
  double best_boost_factor = 1.0 + (alr * 2.0 * stagnation_factor);
  for (int y = 0; y < NY; y++) {
    std::vector<uint8_t> current_year_actions = W.current_run_actions[y];
    current_year_actions.insert(current_year_actions.end(), W.current_deficit_actions[y].begin(), W.current_deficit_actions[y].end());
    std::vector<uint8_t> complete_best = W.best_actions[y];
    complete_best.insert(complete_best.end(), W.best_deficit_actions[y].begin(), W.best_deficit_actions[y].end());
    double* yw = W.w[y];
    for (uint8_t a : complete_best) yw[a] = std::min(yw[a] * best_boost_factor, MAX_WEIGHT);
    for (size_t i = 0; i < current_year_actions.size(); i++) {
      uint8_t a = current_year_actions[i];
      if (!contains(complete_best, a)) {
        // f64::max(NaN, MIN_WEIGHT) == MIN_WEIGHT (quirk Q9): std::fmax has the same NaN rule
        yw[a] = std::fmax(yw[a] * penalty_factor, MIN_WEIGHT);
      } else if (i < complete_best.size() && a != complete_best[i]) {
        double mild_penalty = 1.0 / (1.0 + alr * combined_penalty * 0.5);
        yw[a] = std::fmax(yw[a] * mild_penalty, MIN_WEIGHT);
      }
    }
  }
  if (W.iwi > 1200 && rng) {
    for (int y = 0; y < NY; y++)
      for (int k = 0; k < EG_N_ACTIONS; k++) {
        double random_factor = 1.0 + 0.25 * (rng->next_f64() * 2.0 - 1.0);
        W.w[y][k] = std::min(std::max(W.w[y][k] * random_factor, MIN_WEIGHT), MAX_WEIGHT);
      }
  }
}

// ---- strategy.rs:19-258 -------------------------------------------------------------------------------------------
void update_best_strategy(Weights& W, const Metrics& m) {
  double current_score = score_metrics(m, W.cost_only_mode);
  W.iteration_count += 1;
  bool should_update = !W.has_best || current_score > score_metrics(W.best_metrics, W.cost_only_mode);
  if (should_update) {
    W.improvement_history.push_back({W.iteration_count, current_score, m.net, m.cost, m.opinion, m.reliability});
    W.best_metrics = m;
    W.has_best = true;
    W.best_weights.assign(&W.w[0][0], &W.w[0][0] + NY * EG_N_ACTIONS);
    for (int y = 0; y < NY; y++) {
      W.best_actions[y] = W.current_run_actions[y];
      W.best_deficit_actions[y] = W.current_deficit_actions[y];
    }
    W.iwi = 0;
  } else {
    W.iwi += 1;
  }
}

// ---- learning.rs:285-373 --------------------------------------------------------------------------------------------
void apply_deficit_contrast_learning(Weights& W, Rng* rng) {
  if (!W.has_best) return;
  double deterioration = (double)W.iwi / 10.0;
  double iterations = (double)W.iwi;
  double dynamic_threshold = 0.05 * std::max(std::exp(-iterations / 400.0), 0.00001 / 0.05);
  bool force_contrast = W.iwi > 800;
  if (!(deterioration > dynamic_threshold || force_contrast)) return;
  double stagnation_iterations = (double)W.iwi / 10.0;
  double stagnation_factor = 1.0 + (0.2 * std::pow(stagnation_iterations, 1.8));
  double scaled_deterioration = std::pow(deterioration, 0.3);
  double combined_penalty = scaled_deterioration * stagnation_factor;
  double alr = W.learning_rate * (1.0 + 0.1 * (double)W.iwi);
  double penalty_factor = 1.0 / (1.0 + alr * 1.5 * combined_penalty);
  double best_boost_factor = 1.0 + (alr * 2.0 * stagnation_factor * 1.5);
  for (int y = 0; y < NY; y++) {
    const std::vector<uint8_t>& best = W.best_deficit_actions[y];
    double* yw = W.dw[y];
    for (uint8_t a : best) {
      int k = action_deficit_key(a);
      if (k >= 0) yw[k] = std::min(yw[k] * best_boost_factor, MAX_WEIGHT);
    }
    for (uint8_t a : W.current_deficit_actions[y]) {
      if (!contains(best, a)) {
        int k = action_deficit_key(a);
        if (k >= 0) yw[k] = std::fmax(yw[k] * penalty_factor, MIN_WEIGHT);
      }
    }
  }
  if (W.iwi > 1200 && rng) {
    for (int y = 0; y < NY; y++)
      for (int k = 0; k < EG_N_DEFICIT_KEYS; k++) {
        double random_factor = 1.0 + 0.25 * (rng->next_f64() * 2.0 - 1.0);
        W.dw[y][k] = std::min(std::max(W.dw[y][k] * random_factor, MIN_WEIGHT), MAX_WEIGHT);
      }
  }
}

// strategy.rs:313-342: shared.current_* <- local.current_*, reconstructed from the trajectory record.
// `replay` reproduces the double recording of replay iterations (quirk Q10): sampled actions are pushed by
// sample_*_action's replay branch AND by the caller; the forced battery (attempt >= 5) only by the caller.
void transfer_recorded_actions(Weights& W, const eg_traj& t, bool replay) {
  int row = 0;
  for (int y = 0; y < NY; y++) {
    W.current_run_actions[y].clear();
    W.current_deficit_actions[y].clear();
    int nd = std::min<int>(t.n_deficit[y], EG_TRAJ_CAPACITY - row);
    int na = std::min<int>(t.n_additional[y], EG_TRAJ_CAPACITY - row - nd);
    const uint8_t* a = t.actions + row;
    for (int i = 0; i < nd; i++) {
      W.current_run_actions[y].push_back(a[i]);
      W.current_deficit_actions[y].push_back(a[i]);
      if (replay && i < 4) W.current_deficit_actions[y].push_back(a[i]);
    }
    for (int i = nd; i < nd + na; i++) {
      W.current_run_actions[y].push_back(a[i]);
      if (replay) W.current_run_actions[y].push_back(a[i]);
    }
    row += nd + na;
  }
}

// strategy.rs:313-342 as the reference does it: the episode's own (unbounded) lists are copied over
void transfer_recorded_actions_from(Weights& W, const Weights& local) {
  for (int y = 0; y < NY; y++) {
    W.current_run_actions[y] = local.current_run_actions[y];
    W.current_deficit_actions[y] = local.current_deficit_actions[y];
  }
}

bool update_shared_from(Weights& shared, const eg_result& r, const Weights& local, Rng* rng) {  // multi_simulation.rs:494-508
  transfer_recorded_actions_from(shared, local);
  Metrics m{r.net_emissions, r.public_opinion, r.total_cost, r.power_reliability};
  uint32_t before = (uint32_t)shared.improvement_history.size();
  apply_contrast_learning(shared, m, rng);
  update_best_strategy(shared, m);
  apply_deficit_contrast_learning(shared, rng);
  return shared.improvement_history.size() != before;
}

bool update_shared(Weights& shared, const eg_result& r, const eg_traj& t, bool replay, Rng* rng) {  // the same from a C-ABI record
  transfer_recorded_actions(shared, t, replay);
  Metrics m{r.net_emissions, r.public_opinion, r.total_cost, r.power_reliability};
  uint32_t before = (uint32_t)shared.improvement_history.size();
  apply_contrast_learning(shared, m, rng);
  update_best_strategy(shared, m);
  apply_deficit_contrast_learning(shared, rng);
  return shared.improvement_history.size() != before;
}

}  // namespace orc
