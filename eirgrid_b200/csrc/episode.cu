// episode.cu — the batched 2025-2050 episode kernel (sm_100a), one episode per thread.
//
// Replaces, for n episodes at once, the body of the reference's rayon closure up to the write lock:
//   run_iteration / run_simulation / handle_power_deficit      core/iteration.rs:10-95, core/simulation.rs:22-522
//   apply_action / Map::add_generator / add_carbon_offset       core/actions.rs:40-204, utils/map_handler.rs:553-811
//   find_best_generator_location -> find_suitable_location      map_handler.rs:1084-1143, gpu/metal_location_search.rs:110-176
//   Map totals, opinion, capital cost, yearly metrics           map_handler.rs:813-985, analysis/metrics_calculation.rs:7-175
//   sample_action / sample_deficit_action / sample_additional   ai/learning/weights/sampling.rs:76-443
//   update_weights / update_deficit_weights (episode-local)     weights/learning.rs:21-88, weights/deficit.rs:82-135
//   score_metrics                                               ai/metrics/scoring.rs:5-45
//
// Design: every float sum/product of the reference runs over Vec<Generator> in insertion
// order, with the existing plants first. A thread therefore walks ITS episode's plants in the same order and
// reproduces the reference's rounding exactly; the existing-plant prefix of every accumulator and everything that
// needs pow/exp is tabulated per year on the host. Within a year a new plant extends each sequential sum by one
// term, so the sums are carried incrementally and re-walked only when the year (and so every term) changes.
// The 100x100 placement scan collapses to a walk down a per-(class, year) list of sites pre-sorted by their
// static score: only sites within the penalty radius of a plant built in this episode can differ from it.
// Compiled with --fmad=false: no contraction, so + - * / are the reference's IEEE operations.
#include "episode.cuh"

namespace {

constexpr int kBattery100 = 3 * 12;  // AddGenerator(BatteryStorage, 100 %), simulation.rs:376
constexpr int kGasPeaker100 = 3 * 8; // sampling fallbacks, sampling.rs:185,237,321,377
constexpr double kMinWeight = 0.0001, kMaxWeight = 0.999;      // ai/learning/constants.rs:14-15
constexpr double kMaxAcceptableCost = 50000000000.0;            // config/constants.rs:113
constexpr double kMaxAcceptableEmissions = 1000000.0;           // config/constants.rs:112

// ---- Philox4x32-10, counter = (episode lo, episode hi, draw, stream 0), key = seed ------------------------
__device__ __forceinline__ unsigned long long philox_u64(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (unsigned long long)c0 | ((unsigned long long)c1 << 32);
}

struct Rng {
  uint32_t k0, k1, e0, e1, draw;
  __device__ __forceinline__ unsigned long long u64() { return philox_u64(k0, k1, e0, e1, draw++, 0u); }
  __device__ __forceinline__ double f64() { return (double)(u64() >> 11) * (1.0 / 9007199254740992.0); }
  __device__ __forceinline__ uint32_t index(uint32_t n) { return (uint32_t)__umul64hi(u64(), (unsigned long long)n); }
};

struct State {  // ActionResult, ai/metrics/simulation_metrics.rs:14-19
  double net, opinion, balance, cost;
};

// evaluate_action_impact(.., None), scoring.rs:60-84
__device__ __forceinline__ double action_impact(const State& cur, const State& nw) {
  if (cur.net > 0.0) return (cur.net - nw.net) / fmax(fabs(cur.net), 1.0);
  double cost_change = nw.cost - cur.cost;
  double cost_improvement = -cost_change / fmax(fabs(cur.cost), 1.0);
  double opinion_improvement = (nw.opinion - cur.opinion) / fmax(fabs(cur.opinion), 1.0);
  double cost_weight = cur.cost > kMaxAcceptableCost * 8.0 ? 0.8 : 0.5;
  double opinion_weight = 1.0 - cost_weight;
  return cost_improvement * cost_weight + opinion_improvement * opinion_weight;
}

// packed plant: gi(8) gj(8) type(4) mult(2) build(5)
__device__ __forceinline__ uint32_t pack_gen(int gi, int gj, int t, int m, int b) {
  return (uint32_t)gi | ((uint32_t)gj << 8) | ((uint32_t)t << 16) | ((uint32_t)m << 20) | ((uint32_t)b << 22);
}

struct Episode {
  // sequential sums over the episode's fleet for the current year
  double gen[3];
  double co2;
  double op_sum;
  double gcost, gcost_prev;   // capital cost of new plants re-priced at year / year-1 (map_handler.rs:955-958)
  double ocost, ocost_prev;   // same for offsets (map_handler.rs:960-962)
  double off_amount;          // calc_total_carbon_offset(year)
  uint32_t n_gens, n_offs;
  uint32_t flags;
};

template <bool REPLAY>
struct Kernel {
  const EgEpisodeParams& p;
  const EgSmallTables* __restrict__ T;
  Episode e;
  uint32_t gens[EG_MAX_NEW_GENERATORS];
  uint16_t offs[EG_MAX_OFFSETS];       // otype(2) mult(2) build(5)
  double lw[EG_N_ACTIONS];             // episode-local copy of this year's weights once update_weights touched them
  double ldw[EG_N_DEFICIT_KEYS];
  bool lw_valid;
  uint8_t year_actions[EG_MAX_ACTIONS_PER_YEAR];
  uint16_t year_sites[EG_MAX_ACTIONS_PER_YEAR];
  uint8_t n_def_year[EG_NY], n_add_year[EG_NY];
  Rng rng;

  __device__ Kernel(const EgEpisodeParams& p_) : p(p_), T(p_.map.small) {}

  __device__ __forceinline__ double gen_cost(int t, int m, int b, int y) const {
    // get_current_cost: base_cost * inflation * technology_factor * location_modifier, then * multiplier
    // (const_funcs.rs:56, generator.rs:593)
    return __ldg(&T->base_cost[t][b]) * __ldg(&T->year[y].inflation) * __ldg(&T->tech[t][y]) * __ldg(&T->loc_mod[t]) * __ldg(&T->mult[m]);
  }
  __device__ __forceinline__ double off_cost(int o, int m, int y) const {
    return __ldg(&T->off_base_cost[o]) * __ldg(&T->year[y].inflation) * __ldg(&T->mult[m]);  // carbon_offset.rs:188-195
  }
  __device__ __forceinline__ double gen_opinion(int site, int t, int m, int b, int y) const {
    // calc_new_generator_opinion, map_handler.rs:946-948
    return 0.03 * __ldg(&p.map.site_opinion[site]) + __ldg(&T->op_type[y][t]) + __ldg(&p.map.op_cost[EG_OPC_INDEX(y, t, m, b)]);
  }

  // Re-walk the fleet for a new year: every per-plant term depends on the year.
  __device__ void year_start(int y) {
    const EgYearRow& yr = T->year[y];
    e.gen[0] = __ldg(&yr.ex_gen[0]); e.gen[1] = __ldg(&yr.ex_gen[1]); e.gen[2] = __ldg(&yr.ex_gen[2]);
    e.co2 = __ldg(&yr.ex_co2);
    e.op_sum = __ldg(&yr.ex_opinion_sum);
    e.gcost = 0.0; e.gcost_prev = 0.0;
    const int n = p.map.grid_n;
    for (uint32_t i = 0; i < e.n_gens; i++) {
      uint32_t g = gens[i];
      int gi = g & 0xFF, gj = (g >> 8) & 0xFF, t = (g >> 16) & 0xF, m = (g >> 20) & 0x3, b = (g >> 22) & 0x1F;
      e.gen[__ldg(&T->acc_class[t])] += __ldg(&T->net_mw[t]);
      e.co2 += __ldg(&T->co2[t]);
      e.op_sum += gen_opinion(gi * n + gj, t, m, b, y);
      e.gcost += gen_cost(t, m, b, y);
      if (y > 0) e.gcost_prev += gen_cost(t, m, b, y - 1);
    }
    e.ocost = 0.0; e.ocost_prev = 0.0; e.off_amount = 0.0;
    for (uint32_t i = 0; i < e.n_offs; i++) {
      uint32_t o = offs[i];
      int ot = o & 3, m = (o >> 2) & 3, b = (o >> 4) & 0x1F;
      double maturity = __ldg(&T->natural_offset[ot]) ? __ldg(&T->maturity[y - b]) : 1.0;
      e.off_amount += __ldg(&T->off_amount[ot]) * maturity;
      e.ocost += off_cost(ot, m, y);
      if (y > 0) e.ocost_prev += off_cost(ot, m, y - 1);
    }
  }

  __device__ __forceinline__ State state(int y) const {  // simulation.rs:122-135
    State s;
    s.net = e.co2 - e.off_amount;
    uint32_t cnt = __ldg(&T->year[y].ex_active) + e.n_gens;
    s.opinion = cnt > 0 ? e.op_sum / (double)cnt : 1.0;
    s.balance = (e.gen[0] + e.gen[1] + e.gen[2]) - __ldg(&T->year[y].usage_total);
    s.cost = e.gcost + e.ocost;
    return s;
  }

  // MetalLocationSearch::find_suitable_location (CPU branch) as a walk down the pre-sorted site list.
  __device__ int place(int t, int y) const {
    const int pc = __ldg(&T->pclass[t]);
    const int rc = __ldg(&T->rclass_of_pclass[pc]);
    const bool water = __ldg(&T->water_of_pclass[pc]) != 0;
    const int ns = p.map.n_sites, n = p.map.grid_n, km = p.map.kmax;
    const size_t base = ((size_t)pc * EG_NY + y) * ns;
    const uint16_t* __restrict__ order = p.map.order + base;
    const double* __restrict__ stat = p.map.static_score + base;
    const double* __restrict__ pref = p.map.prefix_score + base;
    const double* __restrict__ nf = p.map.near_factor + (size_t)rc * km * km;
    const double size_factor = __ldg(&T->size_factor);
    double best_score = 0.0;
    int best_site = -1;
    for (int k = 0; k < ns; k++) {
      const double s_static = __ldg(&stat[k]);
      // a site in range of new plants only loses score (factors < 1), so nothing below can beat the best so far
      if (s_static < best_score || !(s_static > 0.0)) break;
      const int site = __ldg(&order[k]);
      const int si = site / n, sj = site - si * n;
      double score = __ldg(&pref[k]);
      bool affected = false;
      for (uint32_t g = 0; g < e.n_gens; g++) {
        const uint32_t pk = gens[g];
        int di = si - (int)(pk & 0xFF), dj = sj - (int)((pk >> 8) & 0xFF);
        di = di < 0 ? -di : di; dj = dj < 0 ? -dj : dj;
        if (di < km && dj < km) {
          const double f = __ldg(&nf[di * km + dj]);
          if (f >= 0.0) { score *= f; affected = true; }   // score *= distance / penalty_radius
        }
      }
      if (affected) {
        if (water) score *= __ldg(&p.map.coast_factor[site]);
        score *= size_factor;
      } else {
        score = s_static;
      }
      // strict '>' in scan order (metal_location_search.rs:168): equal scores keep the lower site index
      if (score > best_score || (score == best_score && best_site >= 0 && site < best_site)) {
        best_score = score;
        best_site = site;
      }
    }
    return best_site;
  }

  __device__ __forceinline__ void add_generator(int site, int t, int m, int y) {
    if (e.n_gens >= EG_MAX_NEW_GENERATORS) { e.flags |= EG_FLAG_GEN_OVERFLOW; return; }
    const int n = p.map.grid_n;
    gens[e.n_gens++] = pack_gen(site / n, site % n, t, m, y);
    e.gen[__ldg(&T->acc_class[t])] += __ldg(&T->net_mw[t]);
    e.co2 += __ldg(&T->co2[t]);
    e.op_sum += gen_opinion(site, t, m, y, y);
    e.gcost += gen_cost(t, m, y, y);
    if (y > 0) e.gcost_prev += gen_cost(t, m, y, y - 1);
  }
  __device__ __forceinline__ void add_offset(int ot, int m, int y) {
    if (e.n_offs >= EG_MAX_OFFSETS) { e.flags |= EG_FLAG_OFFSET_OVERFLOW; return; }
    offs[e.n_offs++] = (uint16_t)(ot | (m << 2) | (y << 4));
    double maturity = __ldg(&T->natural_offset[ot]) ? __ldg(&T->maturity[0]) : 1.0;
    e.off_amount += __ldg(&T->off_amount[ot]) * maturity;
    e.ocost += off_cost(ot, m, y);
    if (y > 0) e.ocost_prev += off_cost(ot, m, y - 1);
  }

  // ---- episode-local learning (the deficit handler edits this year's rows of its private weights) ----------
  __device__ void touch_local(int y) {
    if (lw_valid) return;
    for (int k = 0; k < EG_N_ACTIONS; k++) lw[k] = __ldg(&p.policy->w[y][k]);
    for (int k = 0; k < EG_N_DEFICIT_KEYS; k++) ldw[k] = __ldg(&p.policy->dw[y][k]);
    lw_valid = true;
  }
  __device__ __forceinline__ double weight(int y, int k) const { return lw_valid ? lw[k] : __ldg(&p.policy->w[y][k]); }
  __device__ __forceinline__ double dweight(int y, int k) const { return lw_valid ? ldw[k] : __ldg(&p.policy->dw[y][k]); }

  __device__ static int deficit_key_of_type(int t) {  // weights/core.rs:130-149 insertion order
    switch (t) {
      case 8: return 0; case 7: return 1; case 12: return 2; case 11: return 3; case 9: return 4; case 0: return 5;
      case 1: return 6; case 4: return 7; case 10: return 8; case 5: return 9; case 2: return 10; case 3: return 11;
      case 13: return 12; case 14: return 13;
    }
    return -1;  // CoalPlant has no deficit key
  }
  __device__ static int deficit_key_action(int k) {
    const int type_of_key[14] = {8, 7, 12, 11, 9, 0, 1, 4, 10, 5, 2, 3, 13, 14};
    return 3 * type_of_key[k];
  }

  __device__ void update_deficit_weights(int y, int action, double improvement) {  // deficit.rs:82-135
    const int key = deficit_key_of_type(action / 3);
    if (key < 0 || action % 3 != 0) return;
    touch_local(y);
    const double lr = p.policy->learning_rate;
    double adj = improvement > 0.0 ? 1.0 + (lr * improvement * 1.5) : 1.0 / (1.0 + (lr * fabs(improvement) * 1.5));
    ldw[key] = fmin(fmax(ldw[key] * adj, kMinWeight), kMaxWeight);
    if (improvement < 0.0) {
      const double boost = 1.0 + (lr * 0.1);
      for (int k = 0; k < 14; k++)
        if (k != key) ldw[k] = fmin(ldw[k] * boost, kMaxWeight);
    }
  }
  __device__ void update_weights(int y, int action, double improvement) {  // learning.rs:21-88
    touch_local(y);
    const double lr = p.policy->learning_rate;
    const double rel = p.policy->relative_improvement;
    const double immediate = rel > 0.0 ? 0.7 : 0.3;
    const double combined = immediate * improvement + (1.0 - immediate) * rel;
    double adj = combined > 0.0 ? 1.0 + (lr * combined) : 1.0 / (1.0 + (lr * fabs(combined)));
    lw[action] = fmin(fmax(lw[action] * adj, kMinWeight), kMaxWeight);
    if (combined < 0.0) {
      const double boost = 1.0 + (lr * 0.1);
      for (int k = 0; k < 45; k++)
        if (k != action) lw[k] = fmin(lw[k] * boost, kMaxWeight);
      if (p.policy->noop_boost) lw[EG_ACT_DO_NOTHING] = fmin(lw[EG_ACT_DO_NOTHING] * (1.0 + lr * 0.2), kMaxWeight);
    }
  }

  // ---- sampling (canonical key order replaces HashMap iteration order) ---------------------------------------
  __device__ int sample_deficit_action(int y) {  // sampling.rs:315-378
    const bool explore = rng.f64() < p.policy->exploration_rate;
    if (explore) return deficit_key_action((int)rng.index(14));
    double total = 0.0;
    for (int k = 0; k < 14; k++) total += dweight(y, k);
    if (total <= 0.0) return kGasPeaker100;
    double rv = rng.f64() * total;
    for (int k = 0; k < 14; k++) {
      rv -= dweight(y, k);
      if (rv <= 0.0) return deficit_key_action(k);
    }
    return kGasPeaker100;
  }
  __device__ uint32_t sample_additional_actions(int y, uint32_t deficit_count) {  // sampling.rs:380-443
    const uint32_t max_possible = deficit_count >= 20 ? 0 : 20 - deficit_count;
    if (max_possible == 0) return 0;
    const double random_val = rng.f64();
    if (p.policy->has_count_weights) {
      double total = 0.0;
      for (int c = 0; c < EG_N_COUNT_KEYS; c++) total += __ldg(&p.policy->cw[y][c]);
      if (total <= 0.0) return 0;
      double rc = random_val * total;
      for (int c = 0; c < EG_N_COUNT_KEYS; c++) {
        rc -= __ldg(&p.policy->cw[y][c]);
        if (rc <= 0.0) return min((uint32_t)c, max_possible);
      }
      return min(5u, max_possible);
    }
    const double scaled = sqrt(p.policy->exploration_rate);  // powf(0.5) of a non-negative value
    const uint32_t min_actions = (uint32_t)round(2.0 / scaled), max_actions = (uint32_t)round(12.0 / scaled);
    const uint32_t cmax = min(max_actions, max_possible), cmin = min(min_actions, cmax);
    if (cmin == cmax) return cmin;
    return cmin + rng.index(cmax - cmin + 1);
  }
  __device__ int sample_action(int y) {  // sampling.rs:147-238
    const uint32_t iwi = p.policy->iwi;
    const double eps = p.policy->exploration_rate;
    const double cur_eps = iwi > 100 ? eps * (1.0 / (1.0 + 0.01 * (double)iwi)) : eps;
    const bool explore = rng.f64() < cur_eps;
    if (explore) return (int)rng.index(EG_N_ACTIONS);
    double total = 0.0;
    for (int k = 0; k < EG_N_ACTIONS; k++) total += weight(y, k);
    if (total <= 0.0) return kGasPeaker100;
    if (iwi > 500) {
      // stagnation branch: weights sorted descending (stable), raised to a power (sampling.rs:190-220)
      uint8_t idx[EG_N_ACTIONS];
      for (int k = 0; k < EG_N_ACTIONS; k++) {
        const double wk = weight(y, k);
        int j = k;
        while (j > 0 && weight(y, idx[j - 1]) < wk) { idx[j] = idx[j - 1]; j--; }
        idx[j] = (uint8_t)k;
      }
      const double sf = fmin((double)iwi / 1000.0, 3.0);
      const double power = 1.0 + (2.0 * sf);
      double total_scaled = 0.0;
      for (int k = 0; k < EG_N_ACTIONS; k++) total_scaled += pow(weight(y, idx[k]), power);
      double rv = rng.f64() * total_scaled;
      for (int k = 0; k < EG_N_ACTIONS; k++) {
        rv -= pow(weight(y, idx[k]), power);
        if (rv <= 0.0) return idx[k];
      }
      return idx[0];
    }
    double rv = rng.f64() * total;
    for (int k = 0; k < EG_N_ACTIONS; k++) {
      rv -= weight(y, k);
      if (rv <= 0.0) return k;
    }
    return kGasPeaker100;
  }

  __device__ __forceinline__ void record(int slot, int action, int site) {
    if (slot >= EG_MAX_ACTIONS_PER_YEAR) { e.flags |= EG_FLAG_YEAR_OVERFLOW; return; }
    year_actions[slot] = (uint8_t)action;
    year_sites[slot] = site >= 0 ? (uint16_t)site : (uint16_t)EG_SITE_NONE;
  }

  __device__ void run(uint32_t ep) {
    const unsigned long long id = p.same_stream ? 0ull : p.first_episode + ep;
    rng.k0 = (uint32_t)p.seed; rng.k1 = (uint32_t)(p.seed >> 32);
    rng.e0 = (uint32_t)id; rng.e1 = (uint32_t)(id >> 32); rng.draw = 0;
    e.n_gens = 0; e.n_offs = 0; e.flags = 0;
    const eg_traj* in = REPLAY ? p.replay_in + ep : nullptr;
    double total_cost = 0.0, total_credit = 0.0, total_sales = 0.0;
    uint32_t n_def_total = 0, n_add_total = 0;
    eg_result res;

    for (int y = 0; y < EG_NY; y++) {
      year_start(y);
      lw_valid = false;
      State cur = state(y);
      bool deficit_mode = cur.balance < 0.0;                 // simulation.rs:137
      const State initial = cur;                             // simulation.rs:341-356
      double remaining = deficit_mode ? -cur.balance : 0.0;  // Map::handle_power_deficit returns it unchanged (Q2)
      uint32_t attempts = 0, n_def = 0, n_add = 0, n_to_add = 0;
      bool counted = false;
      const uint32_t in_def = REPLAY ? in->n_deficit[y] : 0;

      for (;;) {
        int action;
        bool is_def;
        if (deficit_mode) {
          if (remaining > 0.0) {                             // simulation.rs:358
            attempts++;
            if (REPLAY) action = n_def < in_def ? in->actions[y][n_def] : kBattery100;
            else action = attempts < 5 ? sample_deficit_action(y) : kBattery100;
            is_def = true;
          } else {                                           // simulation.rs:490-519
            if (!REPLAY) {
              const State fin = state(y);
              const double overall_success = action_impact(initial, fin);
              if (fin.balance >= 0.0 && overall_success > 0.0 && n_def > 0) {
                const double success_factor = 0.1 * overall_success;
                for (uint32_t i = 0; i < n_def && i < EG_MAX_ACTIONS_PER_YEAR; i++) update_deficit_weights(y, year_actions[i], success_factor);
              }
            }
            deficit_mode = false;
            continue;
          }
        } else if (!counted) {                               // simulation.rs:144-187
          n_to_add = REPLAY ? in->n_additional[y] : sample_additional_actions(y, n_def);
          counted = true;
          continue;
        } else if (n_add < n_to_add) {                       // simulation.rs:189-198
          if (REPLAY) {
            const uint32_t pos = in_def + n_add;
            action = pos < EG_MAX_ACTIONS_PER_YEAR ? in->actions[y][pos] : EG_ACT_DO_NOTHING;
          } else {
            action = sample_action(y);
          }
          is_def = false;
        } else {
          break;
        }

        State before;
        if (is_def) before = state(y);                       // simulation.rs:380-395
        int site = -1;
        if (action < 45) {                                   // apply_action, actions.rs:42-76
          const int t = action / 3, m = action - 3 * t;
          site = place(t, y);
          if (site >= 0) add_generator(site, t, m, y);
          else e.flags |= EG_FLAG_NO_SITE;
        } else if (action < 57) {                            // actions.rs:129-179
          const int a = action - 45;
          add_offset(a / 3, a % 3, y);
        }                                                    // 57..60: no generator id matches / DoNothing (Q4)
        if (is_def && site < 0) { remaining = 0.0; continue; }  // no site with score > 0: flagged, loop left
        record(n_def + n_add, action, site);
        if (is_def) {
          n_def++;
          const State after = state(y);                      // simulation.rs:412-427
          if (!REPLAY) {
            const double overall = action_impact(before, after);
            const double emis = after.net < before.net ? (before.net - after.net) / fmax(fabs(before.net), 1.0) : 0.0;
            double cost_imp = 0.0;
            if (after.net < 1000.0) {
              const double cost_change = after.cost - before.cost;
              cost_imp = -cost_change / fmax(fabs(before.cost), 1.0);
            }
            const double op_imp = after.cost < kMaxAcceptableCost * 8.0 ? (after.opinion - before.opinion) / fmax(1.0 - before.opinion, 0.1) : 0.0;
            const double combined = overall * 0.7 + emis * 0.15 + cost_imp * 0.1 + op_imp * 0.05;
            update_deficit_weights(y, action, combined);     // simulation.rs:479
            update_weights(y, action, overall * 0.5);        // simulation.rs:482
          }
          remaining = -fmin(after.balance, 0.0);             // simulation.rs:486
        } else {
          n_add++;
        }
      }

      n_def_year[y] = (uint8_t)min(n_def, (uint32_t)EG_MAX_ACTIONS_PER_YEAR);
      n_add_year[y] = (uint8_t)min(n_add, (uint32_t)EG_MAX_ACTIONS_PER_YEAR - n_def_year[y]);
      n_def_total += n_def; n_add_total += n_add;
      const uint32_t used = min(n_def + n_add, (uint32_t)EG_MAX_ACTIONS_PER_YEAR);
      if (p.traj) {
        uint8_t* row = p.traj[ep].actions[y];
        for (uint32_t i = 0; i < used; i++) row[i] = year_actions[i];
        for (uint32_t i = used; i < EG_MAX_ACTIONS_PER_YEAR; i++) row[i] = 0;
      }
      if (p.sites) {
        uint16_t* row = p.sites[ep].site[y];
        for (uint32_t i = 0; i < used; i++) row[i] = year_sites[i];
        for (uint32_t i = used; i < EG_MAX_ACTIONS_PER_YEAR; i++) row[i] = (uint16_t)EG_SITE_NONE;
      }

      // calculate_yearly_metrics, analysis/metrics_calculation.rs:32-175
      const EgYearRow& yr = T->year[y];
      const double usage = __ldg(&yr.usage_total);
      const double generation = e.gen[0] + e.gen[1] + e.gen[2];
      const double balance = generation - usage;
      const double net = e.co2 - e.off_amount;
      const double credit = net >= 0.0 ? 0.0 : (-net) * __ldg(&yr.carbon_price);
      const uint32_t active = __ldg(&yr.ex_active) + e.n_gens;
      const double opinion = active > 0 ? e.op_sum / (double)active : 1.0;
      const double total_capital = e.gcost + e.ocost;
      const double yearly_capital = y == 0 ? total_capital : total_capital - (e.gcost_prev + e.ocost_prev);
      const double sales = (p.energy_sales && balance > 0.0) ? (balance * 8.76) * 50000.0 : 0.0;
      const double yearly_total = yearly_capital + 0.0 + 0.0 - credit - (p.energy_sales ? sales : 0.0);
      if (y == 0) { total_cost = yearly_total; total_credit = credit; total_sales = sales; }
      else { total_cost = total_cost + yearly_total; total_credit = total_credit + credit; total_sales = total_sales + sales; }
      if (p.yearly) {
        eg_year_metrics& m = p.yearly[ep].y[y];
        m.total_population = __ldg(&yr.pop_total);
        m.active_generators = active;
        m.total_power_usage = usage;
        m.total_power_generation = generation;
        m.power_balance = balance;
        m.average_public_opinion = opinion;
        m.yearly_capital_cost = yearly_capital;
        m.total_capital_cost = total_capital;
        m.inflation_factor = __ldg(&yr.inflation);
        m.total_co2_emissions = e.co2;
        m.total_carbon_offset = e.off_amount;
        m.net_co2_emissions = net;
        m.yearly_carbon_credit_revenue = credit;
        m.total_carbon_credit_revenue = total_credit;
        m.yearly_energy_sales_revenue = sales;
        m.total_energy_sales_revenue = total_sales;
        m.yearly_total_cost = yearly_total;
        m.total_cost = total_cost;
        m.reserved = 0.0;
      }
      if (y == EG_NY - 1) {  // iteration.rs:57-84
        res.net_emissions = net;
        res.public_opinion = opinion;
        res.total_cost = total_capital;
        res.power_reliability = balance >= 0.0 ? 1.0 : 0.0;
      }
    }

    // score_metrics, scoring.rs:5-45 (ln evaluated on the device: <= 1 ulp from the host libm)
    {
      const double normalized_cost = fmax(res.total_cost / kMaxAcceptableCost, 1.0);
      const double cost_term = fmin(log(normalized_cost) / p.ln100, 1.0);
      if (p.cost_only) res.score = 2.0 - cost_term;
      else if (res.net_emissions > 0.0) res.score = 1.0 - fmin(res.net_emissions / kMaxAcceptableEmissions, 1.0);
      else {
        const double cost_score = 1.0 - cost_term;
        const double cost_weight = normalized_cost > 8.0 ? 0.8 : 0.5;
        const double opinion_weight = 1.0 - cost_weight;
        res.score = 1.0 + (cost_score * cost_weight + res.public_opinion * opinion_weight);
      }
    }
    res.n_generators = e.n_gens;
    res.n_offsets = e.n_offs;
    res.n_deficit_actions = (uint16_t)n_def_total;
    res.n_additional_actions = (uint16_t)n_add_total;
    res.flags = e.flags;
    res.reserved = 0;
    p.out[ep] = res;
    if (p.traj) {
      for (int y = 0; y < EG_NY; y++) { p.traj[ep].n_deficit[y] = n_def_year[y]; p.traj[ep].n_additional[y] = n_add_year[y]; }
    }
  }
};

template <bool REPLAY>
__global__ void __launch_bounds__(EG_EPISODE_BLOCK) eg_episode_kernel(const __grid_constant__ EgEpisodeParams p) {
  const uint32_t ep = blockIdx.x * blockDim.x + threadIdx.x;
  if (ep >= p.n) return;
  Kernel<REPLAY> k(p);
  k.run(ep);
}

}  // namespace

cudaError_t eg_launch_rollout(const EgEpisodeParams& p, cudaStream_t stream) {
  if (p.n == 0) return cudaSuccess;
  const uint32_t blocks = (p.n + EG_EPISODE_BLOCK - 1) / EG_EPISODE_BLOCK;
  eg_episode_kernel<false><<<blocks, EG_EPISODE_BLOCK, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t eg_launch_replay(const EgEpisodeParams& p, cudaStream_t stream) {
  if (p.n == 0) return cudaSuccess;
  const uint32_t blocks = (p.n + EG_EPISODE_BLOCK - 1) / EG_EPISODE_BLOCK;
  eg_episode_kernel<true><<<blocks, EG_EPISODE_BLOCK, 0, stream>>>(p);
  return cudaGetLastError();
}
