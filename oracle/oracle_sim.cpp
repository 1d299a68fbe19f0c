// oracle_sim.cpp — CPU ORACLE (test infrastructure): one episode.
// Restates core/{iteration,simulation,actions}.rs, analysis/metrics_calculation.rs and the Map totals /
// placement search of utils/map_handler.rs + gpu/metal_location_search.rs. See oracle.hpp.
#include "oracle.hpp"
#include <cmath>
#include <algorithm>
#include <cstring>

namespace orc {

int radius_class_of_type(int t);  // oracle_world.cpp

namespace {

struct EpMap {  // utils/map_handler.rs Map, the per-episode clone (multi_simulation.rs:429-434, iteration.rs:24)
  const World* world;
  Mode mode;
  bool literal_scan;
  std::vector<Generator> generators;
  std::vector<Settlement> settlements;
  std::vector<CarbonOffset> offsets;
  int current_year = 2024;
  bool delays = true;
  size_t n_existing = 0;

  int yidx() const { return current_year - BASE_YEAR; }

  uint32_t calc_total_population() const {  // :813-817
    uint32_t t = 0;
    for (const Settlement& s : settlements) t += s.pop;
    return t;
  }
  double calc_total_power_usage(int year) const {  // :819-827
    double settlement_usage = 0.0;
    for (const Settlement& s : settlements) settlement_usage += s.usage;
    return settlement_usage * (1.0 + ((double)year - 2024.0) * 0.02);
  }
  double calc_total_power_generation() const {  // :829-868 (quirk Q2: the intermittent cap is computed, never applied)
    double total_generation = 0.0, intermittent_generation = 0.0, storage_generation = 0.0;
    for (const Generator& g : generators) {
      double output = g.current_power_output();
      if (is_intermittent(g.type)) intermittent_generation += output;
      else if (is_storage(g.type)) storage_generation += output;
      else total_generation += output;
    }
    return total_generation + intermittent_generation + storage_generation;
  }
  double calc_total_co2_emissions() const {  // :902-910
    double s = 0.0;
    for (const Generator& g : generators)
      if (g.is_active()) s += g.co2_output();
    return s;
  }
  double calc_total_carbon_offset(int year) const {  // :912-919
    double s = 0.0;
    for (const CarbonOffset& o : offsets) s += o.calc_carbon_offset(year);
    return s;
  }
  double calc_net_co2_emissions(int year) const { return calc_total_co2_emissions() - calc_total_carbon_offset(year); }  // :921-923

  double avg_settlement_opinion(const Generator& g, size_t index) const {  // :931-941
    if (mode == FAST) {
      if (index < n_existing) return world->fast.existing_settle_opinion[index];
      return world->fast.settle_opinion[g.site];
    }
    double sum = 0.0;
    for (const Settlement& s : settlements) sum += 1.0 / (1.0 + s.c.distance_to(g.c) / 10000.0);  // settlement.rs:103-106
    return settlements.empty() ? 1.0 : sum / (double)settlements.size();
  }
  double calc_new_generator_opinion(const Generator& g, size_t index, int year) const {  // :925-949
    double avg = avg_settlement_opinion(g, index);
    double type_opinion = calc_type_opinion(g.type, year);
    double cost_opinion = calc_cost_opinion(g.current_cost(year), year);
    return 0.03 * avg + 0.12 * type_opinion + 0.82 * cost_opinion;
  }
  double calculate_average_opinion(int year, uint32_t* count_out = nullptr) const {  // metrics_calculation.rs:7-30
    double total = 0.0;
    uint32_t count = 0;
    for (size_t i = 0; i < generators.size(); i++)
      if (generators[i].is_active()) {
        total += calc_new_generator_opinion(generators[i], i, year);
        count++;
      }
    if (count_out) *count_out = count;
    return count > 0 ? total / (double)count : 1.0;
  }
  double calc_total_capital_cost(int year) const {  // :951-965
    double generator_costs = 0.0;
    for (const Generator& g : generators)
      if (!g.existing) generator_costs += g.current_cost(year);
    double offset_costs = 0.0;
    for (const CarbonOffset& o : offsets) offset_costs += o.current_cost(year);
    return generator_costs + offset_costs;
  }
  double calc_yearly_capital_cost(int year) const {  // :968-985; get_start_year() is always 2025 (quirk Q6)
    double generator_costs = 0.0;
    for (const Generator& g : generators)
      if (g.build_year == year && !g.existing) generator_costs += g.current_cost(year);
    double offset_costs = 0.0;
    for (const CarbonOffset& o : offsets)
      if (2025 == year) offset_costs += o.current_cost(year);
    return generator_costs + offset_costs;
  }
  ActionResult state(int year) const {  // simulation.rs:122-135
    ActionResult r;
    r.net = calc_net_co2_emissions(year);
    r.opinion = calculate_average_opinion(year);
    r.balance = calc_total_power_generation() - calc_total_power_usage(year);
    r.cost = calc_total_capital_cost(year);
    return r;
  }
  void update_construction_status() {  // :1492-1504
    for (Generator& g : generators) g.update_construction_status(current_year);
    for (CarbonOffset& o : offsets) o.update_construction_status(current_year);
  }

  // MetalLocationSearch::find_suitable_location, CPU branch (gpu/metal_location_search.rs:110-176)
  int find_suitable_location(int type, float size_penalty) const {
    const int n = world->grid_n;
    const double step = world->step;
    const double penalty_radius = placement_penalty_radius(type);
    const bool water = (type == OffshoreWind || type == TidalGenerator || type == WaveEnergy);
    double best_score = 0.0;
    int best_site = -1;
    const int scan = literal_scan ? 100 : n;  // literal: num_x = (100000/1000) with Coordinate::new clamping to 50000
    for (int i = 0; i < scan; i++)
      for (int j = 0; j < scan; j++) {
        Coord location = Coord::make((double)i * step, (double)j * step);
        const int ci = literal_scan ? std::min(i, n - 1) : i, cj = literal_scan ? std::min(j, n - 1) : j;
        const int site = ci * n + cj;
        double score;
        size_t first_gen = 0;
        if (mode == FAST) {
          score = world->fast.prefix[radius_class_of_type(type)][yidx()][site];
          first_gen = n_existing;
        } else {
          score = 1.0;
          for (const Settlement& s : settlements) {
            double distance = location.distance_to(s.c);
            double population_factor = (double)s.pop / 1000000.0;
            score *= (1.0 + population_factor) / (1.0 + distance / 10000.0);
          }
        }
        for (size_t gi = first_gen; gi < generators.size(); gi++) {
          double distance = location.distance_to(generators[gi].c);
          if (distance < penalty_radius) score *= distance / penalty_radius;
        }
        if (water) {
          double f;
          if (mode == FAST) {
            f = world->fast.coast_factor[site];
          } else {
            double mind = 1.7976931348623157e308;
            bool first = true;
            for (const Coord& p : world->coastline) {
              double d = location.distance_to(p);
              if (first || d < mind) { mind = d; first = false; }
            }
            f = 1.0 / (1.0 + mind / 5000.0);
          }
          score *= f;
        }
        score *= 1.0 - ((double)size_penalty * 0.1);
        if (score > best_score) {
          best_score = score;
          best_site = site;
        }
      }
    return best_site;
  }
};

const double kOffsetSize[4] = {500.0, 300.0, 100.0, 1000.0};                        // actions.rs:134-139 (Forest, Wetland, ActiveCapture, CarbonCredit)
const double kOffsetBaseCost[4] = {1000000.0, 1000000.0, 1000000000.0, 50000000.0};  // constants.rs:245-248
const double kMult[3] = {100.0, 120.0, 150.0};
const double kCo2Rate[15] = {0, 0, 0, 0, 0, 0, 6300.0, 3500.0, 4800.0, 1500.0, 0, 0, 0, 0, 0};

// core/actions.rs:40-204. Returns the chosen site for AddGenerator, -1 otherwise; *no_site set when the search found none.
int apply_action(EpMap& map, uint8_t action, int year, bool* no_site) {
  if (action < 45) {
    const int type = action / 3;
    const double cost_multiplier = std::min(std::max(kMult[action % 3] / 100.0, 1.0), 5.0);
    const double gen_size = 100.0 / 100.0;  // DEFAULT_GENERATOR_SIZE as f64 / 100.0
    int site = map.find_suitable_location(type, (float)gen_size);
    if (site < 0) {
      // Reference: falls through to find_location_with_min_score / a fallback generator type
      // (map_handler.rs:1146-1174, actions.rs:77-89). Unreachable while any site scores > 0; flagged instead.
      *no_site = true;
      return -1;
    }
    Generator g;
    g.existing = false;
    g.type = type;
    g.site = site;
    g.c = Coord::make((double)(site / map.world->grid_n) * map.world->step, (double)(site % map.world->grid_n) * map.world->step);
    g.base_cost = gen_base_cost(type, year);
    g.power_out = gen_base_power(type);
    g.size = std::min(std::max(gen_size, 0.1), 1.0);
    g.co2_out = kCo2Rate[type] * gen_size;
    g.build_year = year;
    g.mult = std::min(std::max(cost_multiplier, 1.0), 5.0);  // set_construction_cost_multiplier (status Planned, commissioning_year 0 -> no recalculation)
    // Map::add_generator, map_handler.rs:553-709
    const double public_opinion = 0.65;
    if (map.delays) {
      double planning_time = calc_planning_permission_time(type, map.current_year, public_opinion, 1.0);
      double construction_time = calc_construction_time(type, map.current_year, 1.0);
      unsigned est = (unsigned)std::ceil((double)map.current_year + planning_time + construction_time);
      if (est > (unsigned)END_YEAR) return site;  // action cancelled; the search result is still reported
    }
    g.initialize_construction(map.current_year, public_opinion, map.delays);
    map.generators.push_back(g);
    return site;
  }
  if (action < 57) {
    const int otype = (action - 45) / 3;
    const double cost_multiplier = std::min(std::max(kMult[(action - 45) % 3] / 100.0, 1.0), 5.0);
    CarbonOffset o;
    o.type = otype;
    o.base_cost = kOffsetBaseCost[otype];
    o.size = kOffsetSize[otype];
    o.efficiency = std::min(std::max(0.85, 0.0), 1.0);
    o.mult = std::min(std::max(cost_multiplier, 1.0), 5.0);
    const double public_opinion = 0.65;
    if (map.delays) {  // map_handler.rs:791-806
      double planning_time = calc_offset_planning_time(otype, map.current_year, public_opinion, 1.0);
      double construction_time = calc_offset_construction_time(otype, map.current_year, 1.0);
      unsigned est = (unsigned)std::ceil((double)map.current_year + planning_time + construction_time);
      if (est > (unsigned)END_YEAR) return -1;
    }
    o.initialize_construction(map.current_year, public_opinion, map.delays);
    map.offsets.push_back(o);
    return -1;
  }
  // UpgradeEfficiency("") / AdjustOperation("",0) / CloseGenerator("") find no generator (quirk Q4); DoNothing
  return -1;
}

struct Recorder {
  eg_traj* traj;
  eg_sites* sites;
  uint32_t flags = 0;
  uint32_t n_def_total = 0, n_add_total = 0;
  void init() {
    if (traj) std::memset(traj, 0, sizeof(*traj));
    if (sites) std::memset(sites, 0xFF, sizeof(*sites));
  }
  // The C-ABI record eg_traj stores the year rows back to back in EG_TRAJ_CAPACITY slots (years are visited in order, and
  // within a year the deficit actions come first, so appending keeps the layout). The reference's own lists are the
  // unbounded Weights::current_run_actions / current_deficit_actions, which run_episode fills as well.
  int used = 0;
  void push(int y, bool deficit, uint8_t action, int site) {
    if (deficit) n_def_total++; else n_add_total++;
    if (used >= EG_TRAJ_CAPACITY) { flags |= EG_FLAG_RECORD_OVERFLOW; return; }
    const int s = used++;
    if (traj) {
      traj->actions[s] = action;
      if (deficit) traj->n_deficit[y]++; else traj->n_additional[y]++;
    }
    if (sites) sites->site[s] = site >= 0 ? (uint16_t)site : (uint16_t)EG_SITE_NONE;
  }
};

// first slot of year y's row in a record, and the row's (deficit, additional) lengths cut at the capacity
struct RowView { int row, nd, na; };
RowView row_of(const eg_traj& t, int y) {
  int row = 0;
  for (int k = 0; k < y; k++) {
    const int nd = std::min<int>(t.n_deficit[k], EG_TRAJ_CAPACITY - row);
    row += nd + std::min<int>(t.n_additional[k], EG_TRAJ_CAPACITY - row - nd);
  }
  const int nd = std::min<int>(t.n_deficit[y], EG_TRAJ_CAPACITY - row);
  return {row, nd, std::min<int>(t.n_additional[y], EG_TRAJ_CAPACITY - row - nd)};
}

// core/simulation.rs:319-522
void handle_power_deficit(EpMap& map, double deficit, int year, Weights& W, Rng& rng, const eg_traj* replay_in, Recorder& rec) {
  const int y = year - BASE_YEAR;
  double remaining_deficit = deficit;  // Map::handle_power_deficit returns it unchanged: storage is never charged (quirk Q2)
  uint32_t attempts = 0;
  ActionResult initial_state = map.state(year);
  int replay_pos = 0;
  while (remaining_deficit > 0.0) {
    attempts += 1;
    uint8_t action;
    if (replay_in) {
      // counts that run past the record's capacity (malformed input) are cut there
      const RowView rv = row_of(*replay_in, y);
      action = replay_pos < rv.nd ? replay_in->actions[rv.row + replay_pos] : (uint8_t)(3 * BatteryStorage);
      replay_pos++;
    } else if (attempts < 5) {
      action = sample_deficit_action(W, year, rng);
    } else {
      action = (uint8_t)(3 * BatteryStorage);
    }
    ActionResult current_state = map.state(year);
    if (action < 45) {
      bool no_site = false;
      int site = apply_action(map, action, year, &no_site);
      if (no_site) { rec.flags |= EG_FLAG_NO_SITE; break; }
      W.current_deficit_actions[y].push_back(action);  // record_deficit_action
      W.current_run_actions[y].push_back(action);      // record_action
      rec.push(y, true, action, site);
      ActionResult new_state = map.state(year);
      double overall_improvement = evaluate_action_impact(current_state, new_state, false);
      double emissions_improvement = new_state.net < current_state.net
                                         ? (current_state.net - new_state.net) / std::max(std::fabs(current_state.net), 1.0)
                                         : 0.0;
      double cost_improvement = 0.0;
      if (new_state.net < 1000.0) {
        double cost_change = new_state.cost - current_state.cost;
        cost_improvement = -cost_change / std::max(std::fabs(current_state.cost), 1.0);
      }
      double opinion_improvement = new_state.cost < 50000000000.0 * 8.0
                                       ? (new_state.opinion - current_state.opinion) / std::max(1.0 - current_state.opinion, 0.1)
                                       : 0.0;
      double combined_improvement = overall_improvement * 0.7 + emissions_improvement * 0.15 + cost_improvement * 0.1 + opinion_improvement * 0.05;
      update_deficit_weights(W, action, year, combined_improvement);
      update_weights(W, action, year, overall_improvement * 0.5);
      remaining_deficit = -std::min(new_state.balance, 0.0);
    }
  }
  ActionResult final_state = map.state(year);
  double overall_success = evaluate_action_impact(initial_state, final_state, false);
  if (final_state.balance >= 0.0 && overall_success > 0.0 && !W.current_deficit_actions[y].empty()) {
    double success_factor = 0.1 * overall_success;
    std::vector<uint8_t> acts = W.current_deficit_actions[y];
    for (uint8_t a : acts) update_deficit_weights(W, a, year, success_factor);
  }
}

}  // namespace

void run_episode(const World& world, Weights& W, const eg_run_cfg& cfg, uint64_t seed, uint64_t episode_id,
                 const EpisodeIO& io, Mode mode, bool literal_scan) {
  // run_iteration, core/iteration.rs:10-95
  EpMap map;
  map.world = &world;
  map.mode = mode;
  map.literal_scan = literal_scan;
  map.generators = world.existing;
  map.n_existing = world.existing.size();
  map.settlements = world.settlements;
  W.clear_current_run();
  W.force_best_actions = cfg.replay_best != 0;
  // run_simulation, core/simulation.rs:22-317
  map.delays = cfg.enable_construction_delays != 0;
  Rng rng(seed, cfg.same_stream_all_episodes ? 0 : episode_id, 0);
  Recorder rec{io.traj, io.sites};
  rec.init();
  eg_year_metrics prev;
  std::memset(&prev, 0, sizeof(prev));
  eg_year_metrics last;
  std::memset(&last, 0, sizeof(last));
  for (int year = BASE_YEAR; year <= END_YEAR; year++) {
    const int y = year - BASE_YEAR;
    map.current_year = year;
    map.update_construction_status();
    if (year > BASE_YEAR) {  // :107-120
      for (Settlement& s : map.settlements) {
        uint32_t new_pop = (uint32_t)std::round((double)s.pop * 1.01);
        s.pop = new_pop;
        double per_capita_usage = calc_power_usage_per_capita(year);
        s.usage = (double)new_pop * per_capita_usage;
      }
    }
    ActionResult current_state = map.state(year);
    if (current_state.balance < 0.0) handle_power_deficit(map, -current_state.balance, year, W, rng, io.replay_in, rec);
    uint32_t num_additional;
    if (io.replay_in) num_additional = (uint32_t)row_of(*io.replay_in, y).na;
    else if (W.force_best_actions) num_additional = W.has_best ? (uint32_t)W.best_actions[y].size() : 0;  // :146-162
    else num_additional = sample_additional_actions(W, year, rng);
    for (uint32_t i = 0; i < num_additional; i++) {
      uint8_t action;
      if (io.replay_in) {
        const RowView rv = row_of(*io.replay_in, y);
        action = io.replay_in->actions[rv.row + rv.nd + (int)i];
      } else {
        action = sample_action(W, year, rng);
      }
      bool no_site = false;
      int site = apply_action(map, action, year, &no_site);
      if (no_site) rec.flags |= EG_FLAG_NO_SITE;
      W.current_run_actions[y].push_back(action);  // record_action, :197
      rec.push(y, false, action, site);
    }
    // calculate_yearly_metrics, analysis/metrics_calculation.rs:32-175
    eg_year_metrics m;
    std::memset(&m, 0, sizeof(m));
    m.total_population = map.calc_total_population();
    m.total_power_usage = map.calc_total_power_usage(year);
    m.total_power_generation = map.calc_total_power_generation();
    m.power_balance = m.total_power_generation - m.total_power_usage;
    m.total_co2_emissions = map.calc_total_co2_emissions();
    m.total_carbon_offset = map.calc_total_carbon_offset(year);
    m.net_co2_emissions = map.calc_net_co2_emissions(year);
    double carbon_credit_revenue = m.net_co2_emissions >= 0.0 ? 0.0 : (-m.net_co2_emissions) * carbon_price(year);  // const_funcs.rs:206-218
    uint32_t active = 0;
    m.average_public_opinion = map.calculate_average_opinion(year, &active);
    m.active_generators = active;
    if (year == 2025) m.yearly_capital_cost = map.calc_yearly_capital_cost(year);
    else m.yearly_capital_cost = map.calc_total_capital_cost(year) - map.calc_total_capital_cost(year - 1);
    m.total_capital_cost = map.calc_total_capital_cost(year);
    m.inflation_factor = calc_inflation_factor(year);
    double sales = 0.0;
    if (cfg.enable_energy_sales && m.power_balance > 0.0) sales = (m.power_balance * 8.76) * 50000.0;  // const_funcs.rs:225-237
    m.yearly_energy_sales_revenue = sales;
    m.yearly_total_cost = m.yearly_capital_cost + 0.0 + 0.0 - carbon_credit_revenue - (cfg.enable_energy_sales ? sales : 0.0);
    m.yearly_carbon_credit_revenue = carbon_credit_revenue;
    if (year > BASE_YEAR) {
      m.total_cost = prev.total_cost + m.yearly_total_cost;
      m.total_carbon_credit_revenue = prev.total_carbon_credit_revenue + carbon_credit_revenue;
      m.total_energy_sales_revenue = prev.total_energy_sales_revenue + sales;
    } else {
      m.total_cost = m.yearly_total_cost;
      m.total_carbon_credit_revenue = carbon_credit_revenue;
      m.total_energy_sales_revenue = sales;
    }
    if (io.yearly) io.yearly->y[y] = m;
    prev = m;
    last = m;
  }
  if (io.result) {  // iteration.rs:57-84
    eg_result r;
    std::memset(&r, 0, sizeof(r));
    r.net_emissions = last.net_co2_emissions;
    r.public_opinion = last.average_public_opinion;
    r.total_cost = last.total_capital_cost;
    r.power_reliability = last.power_balance >= 0.0 ? 1.0 : 0.0;
    Metrics mm{r.net_emissions, r.public_opinion, r.total_cost, r.power_reliability};
    r.score = score_metrics(mm, cfg.cost_only != 0);
    r.n_generators = (uint32_t)(map.generators.size() - map.n_existing);
    r.n_offsets = (uint32_t)map.offsets.size();
    r.n_deficit_actions = (uint16_t)rec.n_def_total;
    r.n_additional_actions = (uint16_t)rec.n_add_total;
    r.flags = rec.flags;
    *io.result = r;
  }
}

}  // namespace orc
