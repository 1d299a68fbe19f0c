// oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A literal C++17 restatement of the reference's episode hot path (ETM-Code/eirgrid, aiSimulator/src/...).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
// link, load or run anything in this directory. The product (eirgrid_b200/) never includes it.
//
// PARITY STATUS: the Rust reference cannot be compiled in this environment (no cargo/rustc, crates not
// vendored) and ships no tests. Pinned against reference artefacts: population + power-usage columns of
// README.md:96-121 (all 26 years), 2025 generation formula (README.md:96), cache/location_analysis.json
// (2601 sites x 15 types). Everything else (CO2, cost, opinion, score, placement, sampling, learning) is
// "PARITY UNPINNED": this file's reading of the code is the only definition.
//
// Deliberate, documented deviations from the reference (it is itself not reproducible there):
//   * HashMap iteration order (sampling.rs:228-233,367-374,414-419,167-178) -> canonical key order
//     (insertion order of ActionWeights::new, weights/core.rs:35-152; counts ascending).
//   * rand StdRng/thread_rng -> Philox4x32-10, counter = (episode id, draw index), key = seed.
//   * gen_range(0..n) -> floor(u64 * n / 2^64) (no rejection step; bias < n/2^64).
//   * carbon-offset coordinates (actions.rs:142-145) are never read -> not drawn.
#pragma once
#include <cstdint>
#include <vector>
#include <string>
#include "../include/eirgrid_b200.h"

namespace orc {

constexpr int BASE_YEAR = 2025;  // config/constants.rs:2
constexpr int END_YEAR = 2050;   // config/constants.rs:3
constexpr int NY = 26;

// models/generator.rs:11-36 (enum order)
enum GenType : int {
  OnshoreWind = 0, OffshoreWind, DomesticSolar, CommercialSolar, UtilitySolar, Nuclear, CoalPlant,
  GasCombinedCycle, GasPeaker, Biomass, HydroDam, PumpedStorage, BatteryStorage, TidalGenerator, WaveEnergy
};
// canonical order = insertion order of weights/core.rs:100-114 (enum order in carbon_offset.rs:10-15 differs)
enum OffType : int { Forest = 0, Wetland, ActiveCapture, CarbonCredit };

enum CStatus : int { Planned = 0, PlanningPermissionGranted, UnderConstruction, Operational, Decommissioned };

double powi(double a, int b);  // LLVM compiler-rt __powidf2, what Rust's f64::powi lowers to

// ---- config/const_funcs.rs -----------------------------------------------------------------------
double calc_inflation_factor(int year);
double calc_power_usage_per_capita(int year);
double cost_evolution_rate(int t);
double gen_base_cost(int t, int year);
double gen_base_power(int t);
bool requires_water(int t);
bool can_be_urban(int t);
bool is_intermittent(int t);
bool is_storage(int t);
double calc_generator_cost(int t, double base_cost, int year, bool is_urban, bool is_coastal, bool is_river);
double calc_type_opinion(int t, int year);
double calc_cost_opinion(double cost, int year);
double carbon_price(int year);
double calc_planning_permission_time(int t, int year, double opinion, double mult);
double calc_construction_time(int t, int year, double mult);
double calc_offset_planning_time(int o, int year, double opinion, double mult);
double calc_offset_construction_time(int o, int year, double mult);
double placement_penalty_radius(int t);

struct Coord {
  double x = 0, y = 0;
  static Coord make(double x, double y);  // data/poi.rs:11-15 (clamps to [0, 50000]^2)
  double distance_to(const Coord& o) const;
};
bool point_in_polygon(const Coord& p, const std::vector<Coord>& poly);  // const_funcs.rs:143-158

struct Settlement {  // models/settlement.rs
  Coord c;
  uint32_t pop = 0;
  double usage = 0;
};

struct Generator {  // models/generator.rs:374-449
  bool existing = false;  // id starts with "Existing_"
  int type = 0;
  Coord c;
  int site = -1;  // candidate-site index for simulation-built plants
  double base_cost = 0, power_out = 0, size = 1, co2_out = 0;
  double efficiency = 0.99, operation = 1.0;
  int commissioning_year = 0;
  bool active_flag = true;
  int status = Planned;
  double planning_time = 0, construction_time = 0;
  int planning_year = 0, construction_start = 0, construction_complete = 0;
  double mult = 1.0;
  int build_year = 2020;  // get_build_year(), generator.rs:689-701
  bool is_active() const { return active_flag && status == Operational; }
  void initialize_construction(int year, double opinion, bool delays);
  bool update_construction_status(int year);
  double current_power_output() const;  // hour == None
  double current_cost(int year) const;
  double co2_output() const;
};

struct CarbonOffset {  // models/carbon_offset.rs
  int type = 0;
  double base_cost = 0, size = 0, efficiency = 0.85;
  int status = Planned;
  double planning_time = 0, construction_time = 0;
  int planning_year = 0, construction_start = 0, construction_complete = 0, commissioning_year = 0;
  double mult = 1.0;
  void initialize_construction(int year, double opinion, bool delays);
  bool update_construction_status(int year);
  double current_cost(int year) const;
  double calc_carbon_offset(int year) const;
};

// Static tables used by the "fast" mode only (exact-preserving restructurings).
struct FastTables {
  int grid_n = 0;
  double step = 0;
  // settle_prod[y][site]: sequential product over settlements with year-y populations
  std::vector<std::vector<double>> settle_prod;
  // prefix[rc][y][site]: settle_prod then existing plants in list order for radius class rc
  std::vector<std::vector<std::vector<double>>> prefix;
  std::vector<double> coast_factor;   // 1/(1+min_coast_d/5000) per site
  std::vector<double> settle_opinion; // avg_settlement_opinion per site
  std::vector<double> existing_settle_opinion;  // per existing plant
  std::vector<std::vector<uint32_t>> pop;  // pop[y][s]
};

struct World {  // the base map handed to every episode (multi_simulation.rs:429-434)
  std::vector<Coord> coastline;
  std::vector<Settlement> settlements;
  std::vector<Generator> existing;
  int grid_n = 51;
  double step = 1000.0;
  FastTables fast;
  bool fast_ready = false;
  void build_fast_tables();
};

// initialize_map (main.rs:74-125): raw rows as parsed from the asset files
void world_add_settlement_raw(World& w, double lat, double lon, uint32_t population);  // settlements_loader.rs:23-42
void world_add_settlement_xy(World& w, double x, double y, uint32_t population);
int fuel_to_type(const std::string& fuel);                                              // generators_loader.rs:47-57
void world_add_existing_raw(World& w, double capacity, double lat, double lon, int type);  // generators_loader.rs:133-207 + map.add_generator at year 2024
void world_add_existing_xy(World& w, double capacity, double x, double y, int type);
void world_add_coast(World& w, double x, double y);

// ---- policy table (ai/learning/weights/*) ----------------------------------------------------------
struct Metrics { double net = 0, opinion = 0, cost = 0, reliability = 0; };  // SimulationMetrics
double score_metrics(const Metrics& m, bool cost_only);                       // scoring.rs:5-45
struct ActionResult { double net = 0, opinion = 0, balance = 0, cost = 0; };
double evaluate_action_impact(const ActionResult& a, const ActionResult& b, bool cost_only);  // scoring.rs:46-84

struct Improvement { uint32_t iteration; double score, net, cost, opinion, reliability; };

struct Weights {  // weights/mod.rs:49-107, dense
  double w[NY][EG_N_ACTIONS];
  double dw[NY][EG_N_DEFICIT_KEYS];
  double cw[NY][EG_N_COUNT_KEYS];
  bool has_count_weights = true;
  double learning_rate = 0.2, exploration_rate = 0.2;
  bool has_best = false;
  Metrics best_metrics;
  std::vector<double> best_weights;  // NY*61 when has_best
  std::vector<uint8_t> best_actions[NY];
  std::vector<uint8_t> best_deficit_actions[NY];
  uint32_t iteration_count = 0, iwi = 0;
  std::vector<uint8_t> current_run_actions[NY];
  std::vector<uint8_t> current_deficit_actions[NY];
  bool force_best_actions = false;
  size_t replay_index[NY] = {0};
  size_t deficit_replay_index[NY] = {0};
  bool cost_only_mode = false;  // ActionWeights.optimization_mode: never set by the driver (quirk Q12)
  std::vector<Improvement> improvement_history;
  Weights();  // ActionWeights::new, core.rs:25-250
  void clear_current_run();
};
uint8_t deficit_key_action(int k);
int action_deficit_key(uint8_t code);  // -1 if the action is not a deficit key

struct Rng {  // Philox4x32-10
  uint32_t key[2];
  uint32_t ctr_hi[2];
  uint32_t stream;
  uint32_t draw = 0;
  Rng(uint64_t seed, uint64_t episode, uint32_t stream_);
  uint64_t next_u64();
  double next_f64();
  uint64_t next_index(uint64_t n);
};
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

// sampling.rs
uint8_t sample_action(Weights& W, int year, Rng& rng);
uint8_t sample_deficit_action(Weights& W, int year, Rng& rng);
uint32_t sample_additional_actions(Weights& W, int year, Rng& rng);
// learning.rs / deficit.rs / strategy.rs
void update_weights(Weights& W, uint8_t action, int year, double improvement);
void update_deficit_weights(Weights& W, uint8_t action, int year, double improvement);
void apply_contrast_learning(Weights& W, const Metrics& cur, Rng* rng);
void update_best_strategy(Weights& W, const Metrics& m);
void apply_deficit_contrast_learning(Weights& W, Rng* rng);
void transfer_recorded_actions(Weights& W, const eg_traj& traj, bool replay);

enum Mode : int { FAITHFUL = 0, FAST = 1 };

struct EpisodeIO {
  const eg_traj* replay_in = nullptr;  // trajectory-replay mode: no sampling
  eg_result* result = nullptr;
  eg_traj* traj = nullptr;
  eg_sites* sites = nullptr;
  eg_yearly* yearly = nullptr;
};

// run_iteration + run_simulation (core/iteration.rs:10-95, core/simulation.rs:22-522).
// `local` is the episode's private clone of the shared weights and is mutated like the reference's local_weights.
void run_episode(const World& world, Weights& local, const eg_run_cfg& cfg, uint64_t seed, uint64_t episode_id,
                 const EpisodeIO& io, Mode mode, bool literal_scan = false);

// the write-lock section multi_simulation.rs:494-508 for one finished episode
bool update_shared(Weights& shared, const eg_result& r, const eg_traj& t, bool replay, Rng* rng);
// the same with the episode's own unbounded lists (transfer_recorded_actions_from, strategy.rs:313-342) instead of a record
bool update_shared_from(Weights& shared, const eg_result& r, const Weights& local, Rng* rng);

// location analysis (map_handler.rs:1319-1433)
double calculate_generator_suitability(const World& w, const std::vector<Generator>& gens, const Coord& c, int type);

}  // namespace orc
