// oracle_capi.cpp — CPU ORACLE (test infrastructure): extern "C" surface for ctypes (tests/, bench.py's
// cpu_baseline / --impl reference legs, __graft_entry__.smoke()). Not product code. See oracle.hpp.
#include "oracle.hpp"
#include <thread>
#include <functional>
#include <atomic>
#include <cstring>
#include <cmath>
#include <algorithm>

using namespace orc;

namespace orc { int radius_class_of_type(int t); }

extern "C" {

void* orc_world_new(int grid_n, double step) {
  World* w = new World();
  w->grid_n = grid_n;
  w->step = step;
  return w;
}
void orc_world_free(void* w) { delete (World*)w; }
void orc_world_add_settlement_raw(void* w, double lat, double lon, uint32_t pop) { world_add_settlement_raw(*(World*)w, lat, lon, pop); }
void orc_world_add_settlement_xy(void* w, double x, double y, uint32_t pop) { world_add_settlement_xy(*(World*)w, x, y, pop); }
void orc_world_add_existing_raw(void* w, double cap, double lat, double lon, int type) { world_add_existing_raw(*(World*)w, cap, lat, lon, type); }
void orc_world_add_existing_xy(void* w, double cap, double x, double y, int type) { world_add_existing_xy(*(World*)w, cap, x, y, type); }
void orc_world_add_coast(void* w, double x, double y) { world_add_coast(*(World*)w, x, y); }
void orc_world_build_fast(void* w) { ((World*)w)->build_fast_tables(); }
int orc_fuel_to_type(const char* fuel) { return fuel_to_type(fuel); }
void orc_world_counts(void* wp, uint32_t out[3]) {
  World* w = (World*)wp;
  out[0] = (uint32_t)w->settlements.size();
  out[1] = (uint32_t)w->existing.size();
  out[2] = (uint32_t)w->coastline.size();
}
// settlement/existing coordinates after the loader transforms (for feeding eg_map_set in tests)
void orc_world_settlement(void* wp, uint32_t i, double* x, double* y, uint32_t* pop) {
  World* w = (World*)wp;
  *x = w->settlements[i].c.x; *y = w->settlements[i].c.y; *pop = w->settlements[i].pop;
}
void orc_world_existing(void* wp, uint32_t i, double* x, double* y, int* type, double* cap, double* planning, double* construction) {
  World* w = (World*)wp;
  const Generator& g = w->existing[i];
  *x = g.c.x; *y = g.c.y; *type = g.type; *cap = g.power_out; *planning = g.planning_time; *construction = g.construction_time;
}

// fast-mode static tables, for checking the GPU-built site tables
int orc_world_prefix(void* wp, int yidx, int rclass, double* out) {
  World* w = (World*)wp;
  if (!w->fast_ready) return -1;
  const auto& v = w->fast.prefix[rclass][yidx];
  std::memcpy(out, v.data(), v.size() * sizeof(double));
  return 0;
}
int orc_world_site_static(void* wp, double* coast_factor, double* settle_opinion) {
  World* w = (World*)wp;
  if (!w->fast_ready) return -1;
  if (coast_factor) std::memcpy(coast_factor, w->fast.coast_factor.data(), w->fast.coast_factor.size() * sizeof(double));
  if (settle_opinion) std::memcpy(settle_opinion, w->fast.settle_opinion.data(), w->fast.settle_opinion.size() * sizeof(double));
  return 0;
}

// README.md:96-121 pins: population and power usage per year (action-independent)
void orc_world_demand(void* wp, uint32_t pop_out[26], double usage_out[26]) {
  World* w = (World*)wp;
  std::vector<Settlement> s = w->settlements;
  for (int year = BASE_YEAR; year <= END_YEAR; year++) {
    if (year > BASE_YEAR)
      for (Settlement& e : s) {
        uint32_t np = (uint32_t)std::round((double)e.pop * 1.01);
        e.pop = np;
        e.usage = (double)np * calc_power_usage_per_capita(year);
      }
    uint32_t tp = 0;
    double tu = 0.0;
    for (Settlement& e : s) { tp += e.pop; tu += e.usage; }
    pop_out[year - BASE_YEAR] = tp;
    usage_out[year - BASE_YEAR] = tu * (1.0 + ((double)year - 2024.0) * 0.02);
  }
}
// generation if every existing plant were Operational (README's 2025 row, earlier code revision)
double orc_world_existing_generation_if_operational(void* wp) {
  World* w = (World*)wp;
  double t = 0.0, in = 0.0, st = 0.0;
  for (Generator g : w->existing) {
    g.status = Operational;
    double o = g.current_power_output();
    if (is_intermittent(g.type)) in += o; else if (is_storage(g.type)) st += o; else t += o;
  }
  return t + in + st;
}
// year in which each existing plant becomes Operational (quirk Q1)
void orc_world_existing_online_year(void* wp, int* out) {
  World* w = (World*)wp;
  std::vector<Generator> g = w->existing;
  for (size_t i = 0; i < g.size(); i++) out[i] = 0;
  for (int year = BASE_YEAR; year <= END_YEAR; year++)
    for (size_t i = 0; i < g.size(); i++) {
      g[i].update_construction_status(year);
      if (!out[i] && g[i].is_active()) out[i] = year;
    }
}

void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
double orc_powi(double a, int b) { return powi(a, b); }
double orc_inflation(int year) { return calc_inflation_factor(year); }
double orc_gen_cost(int type, int build_year, int mult_idx, int year) {
  Generator g;
  g.type = type;
  g.base_cost = gen_base_cost(type, build_year);
  static const double m[3] = {1.0, 1.2, 1.5};
  g.mult = m[mult_idx];
  return g.current_cost(year);
}
double orc_score(double net, double opinion, double cost, double reliability, int cost_only) {
  return score_metrics(Metrics{net, opinion, cost, reliability}, cost_only != 0);
}

// ---- weights -----------------------------------------------------------------------------------------
void* orc_weights_new() { return new Weights(); }
void orc_weights_free(void* w) { delete (Weights*)w; }
void* orc_weights_clone(void* w) { return new Weights(*(Weights*)w); }
void orc_weights_get_table(void* wp, eg_weights_table* t) {
  Weights* w = (Weights*)wp;
  std::memset(t, 0, sizeof(*t));
  std::memcpy(t->weights, w->w, sizeof(w->w));
  std::memcpy(t->deficit_weights, w->dw, sizeof(w->dw));
  std::memcpy(t->count_weights, w->cw, sizeof(w->cw));
  t->learning_rate = w->learning_rate;
  t->exploration_rate = w->exploration_rate;
  t->best_metrics[0] = w->best_metrics.net; t->best_metrics[1] = w->best_metrics.opinion;
  t->best_metrics[2] = w->best_metrics.cost; t->best_metrics[3] = w->best_metrics.reliability;
  t->has_count_weights = w->has_count_weights; t->has_best = w->has_best;
  t->iteration_count = w->iteration_count; t->iterations_without_improvement = w->iwi;
}
void orc_weights_set_table(void* wp, const eg_weights_table* t) {
  Weights* w = (Weights*)wp;
  std::memcpy(w->w, t->weights, sizeof(w->w));
  std::memcpy(w->dw, t->deficit_weights, sizeof(w->dw));
  std::memcpy(w->cw, t->count_weights, sizeof(w->cw));
  w->learning_rate = t->learning_rate; w->exploration_rate = t->exploration_rate;
  w->has_count_weights = t->has_count_weights != 0;
  w->iteration_count = t->iteration_count; w->iwi = t->iterations_without_improvement;
}
// list lengths per year and the lists back to back (either buffer may be NULL / too small: written as far as it reaches)
int orc_weights_get_best(void* wp, uint32_t n_best[26], uint8_t* best, size_t best_cap, uint32_t n_best_deficit[26],
                         uint8_t* best_deficit, size_t best_deficit_cap) {
  Weights* w = (Weights*)wp;
  size_t ob = 0, obd = 0;
  for (int y = 0; y < 26; y++) {
    n_best[y] = (uint32_t)w->best_actions[y].size();
    n_best_deficit[y] = (uint32_t)w->best_deficit_actions[y].size();
    for (uint8_t a : w->best_actions[y]) { if (best && ob < best_cap) best[ob] = a; ob++; }
    for (uint8_t a : w->best_deficit_actions[y]) { if (best_deficit && obd < best_deficit_cap) best_deficit[obd] = a; obd++; }
  }
  return w->has_best ? 1 : 0;
}

}  // extern "C"

// ---- episodes ------------------------------------------------------------------------------------------
static void parallel_for_impl(uint32_t n, int threads, const std::function<void(uint32_t)>& fn) {
  if (threads <= 1 || n <= 1) {
    for (uint32_t i = 0; i < n; i++) fn(i);
    return;
  }
  std::atomic<uint32_t> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&]() {
      for (;;) {
        uint32_t i = next.fetch_add(1);
        if (i >= n) break;
        fn(i);
      }
    });
  for (auto& th : pool) th.join();
}

extern "C" {

// n episodes sampled against one weights snapshot (== the rayon closure body up to the write lock)
int orc_rollout(void* wp, void* weights, const eg_run_cfg* cfg, uint64_t seed, uint64_t first_episode, uint32_t n,
                int mode, int literal_scan, int threads, eg_result* out, eg_traj* traj, eg_sites* sites, eg_yearly* yearly) {
  World* world = (World*)wp;
  if (mode == FAST && !world->fast_ready) return -1;
  const Weights* snap = (const Weights*)weights;
  parallel_for_impl(n, threads, [&](uint32_t i) {
    Weights local = *snap;  // multi_simulation.rs:457-460
    EpisodeIO io;
    io.result = out ? out + i : nullptr;
    io.traj = traj ? traj + i : nullptr;
    io.sites = sites ? sites + i : nullptr;
    io.yearly = yearly ? yearly + i : nullptr;
    run_episode(*world, local, *cfg, seed, first_episode + i, io, (Mode)mode, literal_scan != 0);
  });
  return 0;
}

int orc_replay(void* wp, const eg_run_cfg* cfg, const eg_traj* in, uint32_t n, int mode, int threads,
               eg_result* out, eg_traj* traj_out, eg_sites* sites, eg_yearly* yearly) {
  World* world = (World*)wp;
  if (mode == FAST && !world->fast_ready) return -1;
  parallel_for_impl(n, threads, [&](uint32_t i) {
    Weights local;
    EpisodeIO io;
    io.replay_in = in + i;
    io.result = out ? out + i : nullptr;
    io.traj = traj_out ? traj_out + i : nullptr;
    io.sites = sites ? sites + i : nullptr;
    io.yearly = yearly ? yearly + i : nullptr;
    run_episode(*world, local, *cfg, 0, i, io, (Mode)mode, false);
  });
  return 0;
}

// the write-lock section for a batch, in episode-index order
int orc_update(void* weights, const eg_result* results, const eg_traj* trajs, uint32_t n, int replay, uint64_t rng_seed,
               eg_update_stats* stats) {
  Weights* W = (Weights*)weights;
  eg_update_stats s;
  std::memset(&s, 0, sizeof(s));
  s.batch_best_episode = -1;
  s.n_episodes = n;
  for (uint32_t i = 0; i < n; i++) {
    Rng rng(rng_seed, W->iteration_count, 0x55504454u);
    bool improved = update_shared(*W, results[i], trajs[i], replay != 0, &rng);
    if (improved) s.n_improvements++;
    Metrics m{results[i].net_emissions, results[i].public_opinion, results[i].total_cost, results[i].power_reliability};
    double sc = score_metrics(m, false);
    if (s.batch_best_episode < 0 || sc > s.batch_best_score) { s.batch_best_score = sc; s.batch_best_episode = i; }
  }
  s.iterations_without_improvement = W->iwi;
  s.best_score = W->has_best ? score_metrics(W->best_metrics, false) : 0.0;
  if (stats) *stats = s;
  return 0;
}

// One batch of the reference loop with the reference's own bookkeeping: n episodes run against ONE snapshot of the shared
// weights (threads), then the write-lock section for each of them in episode order, fed from the episode's own unbounded
// action lists (transfer_recorded_actions_from) — no eg_traj in between. Results and records are written out as well
// (a record longer than EG_TRAJ_CAPACITY is flagged there, the update still sees the full lists).
int orc_train_batch(void* wp, void* weights, const eg_run_cfg* cfg, uint64_t seed, uint64_t first_episode, uint32_t n,
                    int mode, int threads, uint64_t rng_seed, eg_result* out, eg_traj* traj, eg_update_stats* stats) {
  World* world = (World*)wp;
  if (mode == FAST && !world->fast_ready) return -1;
  Weights* W = (Weights*)weights;
  const Weights snap = *W;
  std::vector<Weights> locals(n, snap);  // multi_simulation.rs:457-460
  std::vector<eg_result> res(n);
  parallel_for_impl(n, threads, [&](uint32_t i) {
    EpisodeIO io;
    io.result = &res[i];
    io.traj = traj ? traj + i : nullptr;
    run_episode(*world, locals[i], *cfg, seed, first_episode + i, io, (Mode)mode, false);
  });
  eg_update_stats s;
  std::memset(&s, 0, sizeof(s));
  s.batch_best_episode = -1;
  s.n_episodes = n;
  for (uint32_t i = 0; i < n; i++) {
    Rng rng(rng_seed, W->iteration_count, 0x55504454u);
    if (update_shared_from(*W, res[i], locals[i], &rng)) s.n_improvements++;
    if (res[i].flags) s.n_flagged++;
    if (s.batch_best_episode < 0 || res[i].score > s.batch_best_score) { s.batch_best_score = res[i].score; s.batch_best_episode = i; }
    if (out) out[i] = res[i];
  }
  s.iterations_without_improvement = W->iwi;
  s.best_score = W->has_best ? score_metrics(W->best_metrics, false) : 0.0;
  if (stats) *stats = s;
  return 0;
}

// location analysis (map_handler.rs:61-142 loop order): point p = (i+half)*(2*half+1) + (j+half)
int orc_location_analysis(void* wp, int use_loaded_map, int half, double step, double* out, uint32_t first, uint32_t n) {
  World* world = (World*)wp;
  World empty;
  empty.coastline = world->coastline;
  const World& w = use_loaded_map ? *world : empty;
  const int side = 2 * half + 1;
  for (uint32_t k = 0; k < n; k++) {
    uint32_t p = first + k;
    int i = (int)(p / side) - half, j = (int)(p % side) - half;
    Coord c = Coord::make((double)i * step, (double)j * step);
    for (int t = 0; t < 15; t++) out[(size_t)k * 15 + t] = calculate_generator_suitability(w, w.existing, c, t);
  }
  return 0;
}

}  // extern "C"
