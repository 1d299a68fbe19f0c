// oracle_world.cpp — CPU ORACLE (test infrastructure): constants, world model, placement, map totals.
// Restates config/{constants,const_funcs,tech_type}.rs, data/*.rs, models/*.rs, gpu/metal_location_search.rs
// and utils/map_handler.rs of the reference. See oracle.hpp for scope and parity status.
#include "oracle.hpp"
#include <cmath>
#include <algorithm>
#include <cctype>

namespace orc {

// ---- Rust f64::powi == llvm.powi.f64.i32 -> compiler-rt __powidf2 (lib/builtins/powidf2.c) ----------
double powi(double a, int b) {
  const bool recip = b < 0;
  double r = 1;
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? 1 / r : r;
}

// ---- config/const_funcs.rs ---------------------------------------------------------------------------
double calc_inflation_factor(int year) { return powi(1.0 + 0.0185, year - BASE_YEAR); }  // :13-15

double calc_power_usage_per_capita(int year) {  // :17-26
  const double BASE_USAGE = 0.001, ANNUAL_INCREASE = 0.02;
  double years_from_base = (double)(year - BASE_YEAR);
  return BASE_USAGE * std::pow(1.0 + ANNUAL_INCREASE, years_from_base);
}

double cost_evolution_rate(int t) {  // generator.rs:184-202, constants.rs:19-24
  switch (t) {
    case OnshoreWind: case OffshoreWind: return 0.99;
    case DomesticSolar: case CommercialSolar: case UtilitySolar: return 0.97;
    case Nuclear: return 0.99;
    case CoalPlant: return 1.10;
    case GasCombinedCycle: case GasPeaker: return 1.04;
    case Biomass: return 0.99;
    case HydroDam: case PumpedStorage: return 1.06;
    case BatteryStorage: return 0.97;
    case TidalGenerator: case WaveEnergy: return 0.95;
  }
  return 1.0;
}

double gen_base_cost(int t, int year) {  // generator.rs:244-298
  static const double base[15] = {1500000.0, 4000000.0, 10000000.0, 40000000.0, 240000000.0, 15000000000.0,
                                  1500000000.0, 560000000.0, 500000000.0, 150000000.0, 2500000000.0,
                                  1200000000.0, 150000000.0, 1000000000.0, 800000000.0};
  double years_from_base = (double)(year - BASE_YEAR);
  return base[t] * std::pow(cost_evolution_rate(t), years_from_base);
}

double gen_base_power(int t) {  // generator.rs:300-318, constants.rs:164-182
  static const double p[15] = {500.0, 800.0, 10.0, 50.0, 300.0, 1500.0, 1000.0, 800.0, 400.0, 50.0,
                               1200.0, 600.0, 500.0, 200.0, 100.0};
  return p[t];
}

bool requires_water(int t) { return t == OffshoreWind || t == TidalGenerator || t == WaveEnergy; }  // generator.rs:142-154
bool can_be_urban(int t) { return t == DomesticSolar || t == CommercialSolar || t == GasPeaker; }   // generator.rs:132-140
bool is_intermittent(int t) { return t <= UtilitySolar; }                                           // generator.rs:86-94
bool is_storage(int t) { return t == PumpedStorage || t == BatteryStorage; }                        // generator.rs:96-101

double calc_generator_cost(int t, double base_cost, int year, bool is_urban, bool is_coastal, bool is_river) {  // :28-57
  double inflation = calc_inflation_factor(year);
  double years_from_base = (double)(year - BASE_YEAR);
  double technology_factor = std::pow(cost_evolution_rate(t), years_from_base);
  double location_modifier = 1.0;
  if (is_urban) {
    double m = 1.0;
    if (t == DomesticSolar || t == CommercialSolar) m = 1.1;  // URBAN_SOLAR_BONUS
    else if (t == GasPeaker) m = 0.7;                         // URBAN_PEAKER_PENALTY
    location_modifier *= m;
  }
  if (requires_water(t)) {
    if (is_coastal) location_modifier *= 1.15;       // COASTAL_BONUS
    else if (is_river) location_modifier *= 1.10;    // RIVER_BONUS
  }
  return base_cost * inflation * technology_factor * location_modifier;
}

double calc_type_opinion(int t, int year) {  // :78-93
  double years = (double)(year - BASE_YEAR);
  double base = 0, chg = 0;
  switch (t) {
    case OnshoreWind: case OffshoreWind: base = 0.83; chg = 0.005; break;
    case DomesticSolar: case CommercialSolar: case UtilitySolar: base = 0.89; chg = 0.008; break;
    case Nuclear: base = 0.43; chg = 0.002; break;
    case CoalPlant: base = 0.41; chg = -0.015; break;
    case GasCombinedCycle: case GasPeaker: base = 0.42; chg = -0.008; break;
    case HydroDam: case PumpedStorage: base = 0.89; chg = 0.004; break;
    case TidalGenerator: case WaveEnergy: base = 0.75; chg = 0.005; break;
    case BatteryStorage: base = 0.85; chg = 0.003; break;
    case Biomass: base = 0.60; chg = 0.001; break;
  }
  double v = base + chg * years;
  return std::min(std::max(v, 0.0), 1.0);
}

double calc_cost_opinion(double cost, int year) {  // :95-106
  double inflation_adjusted_max = 1384000000.0 * calc_inflation_factor(year);
  double normalized_cost = cost / inflation_adjusted_max;
  if (normalized_cost <= 1.0) return 1.0 - normalized_cost;
  return 0.5 * std::exp(-0.5 * (normalized_cost - 1.0));
}

double carbon_price(int year) {  // :186-203
  if (year < 2030) return 75.0;
  if (year < 2040) {
    double phase_length = (double)(2040 - 2030);
    double t = (double)(year - 2030) / phase_length;
    return 75.0 + t * (130.0 - 75.0);
  }
  if (year <= 2050) {
    double phase_length = (double)(2050 - 2040);
    double t = (double)(year - 2040) / phase_length;
    return 130.0 + t * (300.0 - 130.0);
  }
  return 300.0;
}

static double calc_time_reduction_factor(double mult, double reduction_factor) {  // :339-351
  double b = std::min(std::max(mult, 1.0), 5.0);
  if (b <= 1.0) return 1.0;
  double log_reduction = std::min(std::log(b) * reduction_factor, 0.8);
  return 1.0 - log_reduction;
}

// config/tech_type.rs:53-200
enum Tech { T_OnshoreWind, T_OffshoreWind, T_SolarPV, T_Gas, T_Coal, T_Nuclear, T_Hydro, T_Biomass, T_Tidal, T_Wave, T_Storage };
static int map_to_tech_type(int t) {
  switch (t) {
    case OnshoreWind: return T_OnshoreWind;
    case OffshoreWind: return T_OffshoreWind;
    case DomesticSolar: case CommercialSolar: case UtilitySolar: return T_SolarPV;
    case GasCombinedCycle: case GasPeaker: return T_Gas;
    case CoalPlant: return T_Coal;
    case Nuclear: return T_Nuclear;
    case HydroDam: return T_Hydro;
    case PumpedStorage: case BatteryStorage: return T_Storage;
    case Biomass: return T_Biomass;
    case TidalGenerator: return T_Tidal;
    case WaveEnergy: return T_Wave;
  }
  return T_Gas;
}
static double interp_duration(int year, double base_2025, double end_2050) {
  int clamped = std::min(std::max(year, BASE_YEAR), 2050);
  double t = ((double)clamped - (double)BASE_YEAR) / (2050.0 - (double)BASE_YEAR);
  double years = base_2025 + t * (end_2050 - base_2025);
  return std::max(years, end_2050);
}
static double planning_duration(int year, int tech) {
  switch (tech) {
    case T_OnshoreWind: return interp_duration(year, 1.5, 0.5);
    case T_OffshoreWind: return interp_duration(year, 3.0, 1.0);
    case T_SolarPV: return interp_duration(year, 1.0, 0.3);
    case T_Gas: case T_Coal: return interp_duration(year, 2.0, 1.0);
    case T_Nuclear: return interp_duration(year, 5.0, 3.0);
    case T_Hydro: return interp_duration(year, 2.5, 1.5);
    case T_Storage: return interp_duration(year, 1.5, 0.8);
    case T_Biomass: return interp_duration(year, 2.0, 1.0);
    case T_Tidal: case T_Wave: return interp_duration(year, 3.0, 1.5);
  }
  return 0;
}
static double construction_duration(int year, int tech) {
  switch (tech) {
    case T_OnshoreWind: return interp_duration(year, 1.25, 0.75);
    case T_OffshoreWind: return interp_duration(year, 3.0, 2.0);
    case T_SolarPV: return interp_duration(year, 0.5, 0.25);
    case T_Gas: return interp_duration(year, 2.5, 2.0);
    case T_Coal: return interp_duration(year, 3.0, 3.0);
    case T_Nuclear: return interp_duration(year, 7.0, 4.0);
    case T_Hydro: return interp_duration(year, 4.0, 3.5);
    case T_Storage: return interp_duration(year, 1.0, 0.5);
    case T_Biomass: case T_Tidal: case T_Wave: return interp_duration(year, 2.0, 1.5);
  }
  return 0;
}

double calc_planning_permission_time(int t, int year, double opinion, double mult) {  // const_funcs.rs:269-283
  double base_time = planning_duration(year, map_to_tech_type(t));
  double opinion_factor = 1.0 - (opinion * 0.5);
  double cost_factor = calc_time_reduction_factor(mult, 0.25);
  return std::max(base_time * opinion_factor * cost_factor, 0.25);
}
double calc_construction_time(int t, int year, double mult) {  // :285-295
  double base_time = construction_duration(year, map_to_tech_type(t));
  double cost_factor = calc_time_reduction_factor(mult, 0.5);
  return std::max(base_time * cost_factor, 0.1);
}
double calc_offset_planning_time(int o, int year, double opinion, double mult) {  // :297-317
  static const double base[4] = {1.0 /*Forest*/, 1.5 /*Wetland*/, 2.0 /*ActiveCapture*/, 0.5 /*CarbonCredit*/};
  double years_from_base = (double)(year - BASE_YEAR);
  double year_factor = std::pow(1.0 - 0.02, years_from_base);
  double opinion_factor = 1.0 - (opinion * 0.5);
  double cost_factor = calc_time_reduction_factor(mult, 0.25);
  return std::max(base[o] * year_factor * opinion_factor * cost_factor, 0.25);
}
double calc_offset_construction_time(int o, int year, double mult) {  // :319-336
  static const double base[4] = {1.0, 2.0, 3.0, 0.2};
  double years_from_base = (double)(year - BASE_YEAR);
  double year_factor = std::pow(1.0 - 0.03, years_from_base);
  double cost_factor = calc_time_reduction_factor(mult, 0.5);
  return std::max(base[o] * year_factor * cost_factor, 0.1);
}

double placement_penalty_radius(int t) {  // gpu/metal_location_search.rs:139-146
  switch (t) {
    case Nuclear: return 12000.0;
    case CoalPlant: case GasCombinedCycle: return 8000.0;
    case OnshoreWind: case OffshoreWind: return 5000.0;
    case HydroDam: case PumpedStorage: return 7000.0;
    case TidalGenerator: case WaveEnergy: return 6000.0;
    default: return 3000.0;
  }
}

// ---- data/poi.rs ---------------------------------------------------------------------------------------
Coord Coord::make(double x, double y) {
  Coord c;
  c.x = std::min(std::max(x, 0.0), 50000.0);
  c.y = std::min(std::max(y, 0.0), 50000.0);
  return c;
}
double Coord::distance_to(const Coord& o) const {
  double dx = x - o.x, dy = y - o.y;
  return std::sqrt(dx * dx + dy * dy);
}

bool point_in_polygon(const Coord& p, const std::vector<Coord>& poly) {  // const_funcs.rs:143-158
  bool inside = false;
  if (poly.empty()) return false;
  size_t j = poly.size() - 1;
  for (size_t i = 0; i < poly.size(); i++) {
    if (((poly[i].y > p.y) != (poly[j].y > p.y)) &&
        (p.x < (poly[j].x - poly[i].x) * (p.y - poly[i].y) / (poly[j].y - poly[i].y) + poly[i].x)) {
      inside = !inside;
    }
    j = i;
  }
  return inside;
}

// ---- models/generator.rs -------------------------------------------------------------------------------
void Generator::initialize_construction(int year, double opinion, bool delays) {  // :451-480
  commissioning_year = year;
  if (!delays) {
    status = Operational;
    planning_year = construction_start = construction_complete = year;
    return;
  }
  planning_time = calc_planning_permission_time(type, year, opinion, mult);
  construction_time = calc_construction_time(type, year, mult);
  status = Planned;
}
bool Generator::update_construction_status(int year) {  // :482-517
  if (status == Operational || status == Decommissioned) return false;
  double since = (double)(year - commissioning_year);
  switch (status) {
    case Planned:
      if (since >= planning_time) { status = PlanningPermissionGranted; planning_year = year; return true; }
      break;
    case PlanningPermissionGranted:
      status = UnderConstruction; construction_start = year; return true;
    case UnderConstruction: {
      double s2 = (double)(year - construction_start);
      if (s2 >= construction_time) { status = Operational; construction_complete = year; active_flag = true; return true; }
      break;
    }
    default: break;
  }
  return false;
}
double Generator::current_power_output() const {  // :523-554 with hour == None
  if (!is_active()) return 0.0;
  double base_output = power_out * efficiency * operation;
  if (is_intermittent(type)) {
    if (type == OnshoreWind || type == OffshoreWind) return base_output * 0.35;  // WIND_CAPACITY_FACTOR
    return base_output * 0.20;                                                   // SOLAR_CAPACITY_FACTOR
  }
  return base_output;
}
double Generator::current_cost(int year) const {  // :582-594
  double c = calc_generator_cost(type, base_cost, year, can_be_urban(type), requires_water(type), requires_water(type));
  return c * mult;
}
double Generator::co2_output() const {  // :618-626
  if (!is_active()) return 0.0;
  return co2_out * operation * (1.0 - (efficiency - 0.99));
}

// ---- models/carbon_offset.rs ---------------------------------------------------------------------------
void CarbonOffset::initialize_construction(int year, double opinion, bool delays) {  // :115-144
  commissioning_year = year;
  if (!delays) {
    status = Operational;
    planning_year = construction_start = construction_complete = year;
    return;
  }
  planning_time = calc_offset_planning_time(type, year, opinion, mult);
  construction_time = calc_offset_construction_time(type, year, mult);
  status = Planned;
}
bool CarbonOffset::update_construction_status(int year) {  // :147-181
  if (status == Operational || status == Decommissioned) return false;
  double since = (double)(year - commissioning_year);
  switch (status) {
    case Planned:
      if (since >= planning_time) { status = PlanningPermissionGranted; planning_year = year; return true; }
      break;
    case PlanningPermissionGranted:
      status = UnderConstruction; construction_start = year; return true;
    case UnderConstruction: {
      double s2 = (double)(year - construction_start);
      if (s2 >= construction_time) { status = Operational; construction_complete = year; return true; }
      break;
    }
    default: break;
  }
  return false;
}
double CarbonOffset::current_cost(int year) const {  // :188-195
  double inflation_factor = powi(1.0 + 0.0185, year - BASE_YEAR);
  double bc = base_cost * inflation_factor;
  return bc * mult;
}
double CarbonOffset::calc_carbon_offset(int year) const {  // :210-260
  double base_offset = 0;
  switch (type) {
    case Forest: base_offset = size * 25.0; break;
    case ActiveCapture: base_offset = size * 500.0; break;
    case CarbonCredit: base_offset = size * 100.0; break;
    case Wetland: base_offset = size * 40.0; break;
  }
  if (status == Operational) {
    double maturity = 1.0;
    if (type == Forest || type == Wetland) {
      double years_from_start = (double)(year - construction_complete);
      double v = 1.0 - std::exp(-0.1 * years_from_start);
      maturity = std::min(std::max(v, 0.0), 1.0);
    }
    return base_offset * efficiency * maturity;
  }
  if (status == UnderConstruction) {
    double years_in = (double)(year - construction_start);
    double progress = std::min(std::max(years_in / construction_time, 0.0), 1.0);
    double eff = 0;
    if (type == Forest || type == Wetland) eff = std::pow(progress, 0.7) * 0.5;
    else if (type == CarbonCredit) eff = progress * 0.8;
    else eff = std::pow(progress, 2.0) * 0.3;
    return base_offset * efficiency * eff;
  }
  return 0.0;
}

// ---- loaders (data/settlements_loader.rs, data/generators_loader.rs, main.rs:74-125) --------------------
static bool transform_lat_lon_to_grid(double lat, double lon, Coord* out) {  // const_funcs.rs:124-136
  if (lat < 51.4 || lat > 55.4 || lon < -10.6 || lon > -5.9) return false;
  double x = (lon - (-10.6)) * 10638.297872340427;
  double y = (lat - 51.4) * 12500.0;
  *out = Coord::make(x, y);
  return true;
}

void world_add_settlement_xy(World& w, double x, double y, uint32_t population) {
  Settlement s;
  s.c = Coord::make(x, y);
  s.pop = population;
  s.usage = (double)population * calc_power_usage_per_capita(BASE_YEAR);
  w.settlements.push_back(s);
}
void world_add_settlement_raw(World& w, double lat, double lon, uint32_t population) {  // settlements_loader.rs:29-41
  Coord c;
  if (!transform_lat_lon_to_grid(lat, lon, &c)) return;  // reference prints a warning and skips
  world_add_settlement_xy(w, c.x, c.y, population);
}

int fuel_to_type(const std::string& fuel) {  // generators_loader.rs:47-57
  std::string f;
  for (char ch : fuel) f.push_back((char)std::tolower((unsigned char)ch));
  if (f == "gas") return GasCombinedCycle;
  if (f == "coal") return CoalPlant;
  if (f == "wind") return OnshoreWind;
  if (f == "hydro") return HydroDam;
  if (f == "oil") return GasPeaker;
  if (f == "biomass") return Biomass;
  return -1;
}

static double normalize_capacity(double capacity, int t) {  // generators_loader.rs:118-131
  double max_power;
  switch (t) {
    case OnshoreWind: max_power = 500.0; break;
    case OffshoreWind: max_power = 800.0; break;
    case CoalPlant: max_power = 1000.0; break;
    case GasCombinedCycle: max_power = 800.0; break;
    case GasPeaker: max_power = 400.0; break;
    case HydroDam: max_power = 1200.0; break;
    case Biomass: max_power = 50.0; break;
    default: max_power = 800.0; break;
  }
  return std::min(std::max(capacity / max_power, 0.1), 1.0);
}

static double initial_co2_rate(int t) {  // const_funcs.rs:113-122, constants.rs:125-128
  switch (t) {
    case CoalPlant: return 6300.0;
    case GasCombinedCycle: return 3500.0;
    case GasPeaker: return 4800.0;
    case Biomass: return 1500.0;
    default: return 0.0;
  }
}

void world_add_existing_xy(World& w, double capacity, double x, double y, int type) {
  // generators_loader.rs:167-203 (year = 2025) then Map::add_generator at current_year 2024 with
  // enable_construction_delays == true (map_handler.rs:394,553-577; quirk Q1)
  Generator g;
  g.existing = true;
  g.type = type;
  g.c = Coord::make(x, y);
  double size = normalize_capacity(capacity, type);
  // is_coastal only matters for requires_water types, none of which a fuel maps to
  g.base_cost = calc_generator_cost(type, gen_base_cost(type, BASE_YEAR), BASE_YEAR, false, false, false);
  g.power_out = capacity;
  g.size = std::min(std::max(size, 0.1), 1.0);
  g.co2_out = initial_co2_rate(type) * size;
  g.build_year = 2020;
  const int current_year = 2024;
  const double public_opinion = 0.65;  // calculate_public_opinion_at_location, map_handler.rs:1518-1522
  double planning_time = calc_planning_permission_time(type, current_year, public_opinion, 1.0);
  double construction_time = calc_construction_time(type, current_year, 1.0);
  unsigned estimated = (unsigned)std::ceil((double)current_year + planning_time + construction_time);
  if (estimated > (unsigned)END_YEAR) return;  // cancelled (map_handler.rs:566-573)
  g.initialize_construction(current_year, public_opinion, true);
  w.existing.push_back(g);
}

void world_add_existing_raw(World& w, double capacity, double lat, double lon, int type) {  // generators_loader.rs:59-116
  double la = lat, lo = lon;
  if (la < 51.4 || la > 55.4 || lo < -10.6 || lo > -5.9) {
    la = std::min(std::max(la, 51.4), 55.4);
    lo = std::min(std::max(lo, -10.6), -5.9);
  }
  Coord c;
  if (!transform_lat_lon_to_grid(la, lo, &c)) return;
  world_add_existing_xy(w, capacity, c.x, c.y, type);
}

void world_add_coast(World& w, double x, double y) { w.coastline.push_back(Coord::make(x, y)); }  // map_handler.rs:364-375

// ---- fast-mode tables: same arithmetic, evaluated once -------------------------------------------------
static int radius_class(double r) {
  if (r == 3000.0) return 0;
  if (r == 5000.0) return 1;
  if (r == 6000.0) return 2;
  if (r == 7000.0) return 3;
  if (r == 8000.0) return 4;
  return 5;
}
static const double kRadii[6] = {3000.0, 5000.0, 6000.0, 7000.0, 8000.0, 12000.0};

void World::build_fast_tables() {
  FastTables& f = fast;
  f.grid_n = grid_n;
  f.step = step;
  const int ns = grid_n * grid_n;
  const size_t S = settlements.size();
  f.pop.assign(NY, std::vector<uint32_t>(S));
  for (size_t s = 0; s < S; s++) f.pop[0][s] = settlements[s].pop;
  for (int y = 1; y < NY; y++)
    for (size_t s = 0; s < S; s++) f.pop[y][s] = (uint32_t)std::round((double)f.pop[y - 1][s] * 1.01);  // simulation.rs:112
  f.settle_prod.assign(NY, std::vector<double>(ns));
  f.coast_factor.assign(ns, 1.0);
  f.settle_opinion.assign(ns, 1.0);
  f.prefix.assign(6, std::vector<std::vector<double>>(NY, std::vector<double>(ns)));
  for (int i = 0; i < grid_n; i++)
    for (int j = 0; j < grid_n; j++) {
      const int site = i * grid_n + j;
      Coord loc = Coord::make((double)i * step, (double)j * step);
      for (int y = 0; y < NY; y++) {
        double score = 1.0;
        for (size_t s = 0; s < S; s++) {
          double distance = loc.distance_to(settlements[s].c);
          double population_factor = (double)f.pop[y][s] / 1000000.0;
          score *= (1.0 + population_factor) / (1.0 + distance / 10000.0);
        }
        f.settle_prod[y][site] = score;
        for (int rc = 0; rc < 6; rc++) {
          double sc = score;
          for (const Generator& g : existing) {
            double distance = loc.distance_to(g.c);
            if (distance < kRadii[rc]) sc *= distance / kRadii[rc];
          }
          f.prefix[rc][y][site] = sc;
        }
      }
      double mind = 1.7976931348623157e308;
      bool first = true;
      for (const Coord& p : coastline) {
        double d = loc.distance_to(p);
        if (first || d < mind) { mind = d; first = false; }
      }
      f.coast_factor[site] = 1.0 / (1.0 + mind / 5000.0);
      double sum = 0.0;
      for (size_t s = 0; s < S; s++) sum += 1.0 / (1.0 + settlements[s].c.distance_to(loc) / 10000.0);
      f.settle_opinion[site] = S ? sum / (double)S : 1.0;
    }
  f.existing_settle_opinion.resize(existing.size());
  for (size_t e = 0; e < existing.size(); e++) {
    double sum = 0.0;
    for (size_t s = 0; s < S; s++) sum += 1.0 / (1.0 + settlements[s].c.distance_to(existing[e].c) / 10000.0);
    f.existing_settle_opinion[e] = S ? sum / (double)S : 1.0;
  }
  fast_ready = true;
}

int radius_class_of_type(int t) { return radius_class(placement_penalty_radius(t)); }

// ---- location suitability (map_handler.rs:1178-1433) ----------------------------------------------------
static bool is_water_tile(const World& w, const Coord& c) { return !point_in_polygon(c, w.coastline); }
static bool is_urban_area(const World& w, const Coord& c) {  // :1199-1209
  for (const Settlement& s : w.settlements) {
    double distance = s.c.distance_to(c);
    double radius = std::sqrt((double)s.pop) * 5.0;
    if (distance < radius) return true;
  }
  return false;
}
static bool any_of_9_in_water(const World& w, const Coord& c, double d) {  // :1178-1193, 1211-1226
  for (int x = -1; x <= 1; x++)
    for (int y = -1; y <= 1; y++) {
      Coord p = Coord::make(c.x + ((double)x * d), c.y + ((double)y * d));
      if (is_water_tile(w, p)) return true;
    }
  return false;
}
static bool is_coastal_region(const World& w, const Coord& c) { return any_of_9_in_water(w, c, 8000.0); }
static bool is_near_water(const World& w, const Coord& c) { return any_of_9_in_water(w, c, 5000.0); }
static double get_distance_to_nearest_land(const World& w, const Coord& c) {  // :1398-1421
  double min_distance = 1.7976931348623157e308;
  for (int i = -10; i <= 10; i++)
    for (int j = -10; j <= 10; j++) {
      double x = c.x + ((double)i * 1000.0), y = c.y + ((double)j * 1000.0);
      if (x >= 0.0 && x <= 50000.0 && y >= 0.0 && y <= 50000.0) {
        Coord t = Coord::make(x, y);
        if (!is_water_tile(w, t)) min_distance = std::min(min_distance, c.distance_to(t));
      }
    }
  return min_distance;
}
static uint32_t get_nearby_population(const World& w, const Coord& c, double radius) {  // :1423-1433
  uint32_t total = 0;
  for (const Settlement& s : w.settlements)
    if (s.c.distance_to(c) <= radius) total += s.pop;
  return total;
}
static double get_terrain_suitability(const World& w, const Coord& c, int t) {  // :1228-1250 (elevation == 0)
  switch (t) {
    case OnshoreWind: return 1.0;
    case UtilitySolar: return (!is_near_water(w, c)) ? 1.2 : 1.0;
    case Nuclear: return (is_near_water(w, c) && !is_coastal_region(w, c)) ? 1.2 : 0.8;
    default: return 1.0;
  }
}

double calculate_generator_suitability(const World& w, const std::vector<Generator>& gens, const Coord& c, int t) {  // :1319-1396
  switch (t) {
    case OnshoreWind: {
      double base_score = is_urban_area(w, c) ? 0.0 : (is_coastal_region(w, c) ? 0.7 : 0.5);
      double nearby_penalty = 0.0;
      for (const Generator& g : gens) {
        double d = g.c.distance_to(c);
        if (d < 3000.0) nearby_penalty += 0.1 / (1.0 + d);
      }
      return base_score - nearby_penalty;
    }
    case OffshoreWind: case TidalGenerator: case WaveEnergy: {
      if (!is_water_tile(w, c)) return 0.0;
      double depth_factor = 0.8;
      double shore = get_distance_to_nearest_land(w, c);
      double distance_factor = shore < 2000.0 ? 0.3 : (shore > 10000.0 ? 0.5 : 0.7);
      return depth_factor * distance_factor;
    }
    case Nuclear: {
      if (is_urban_area(w, c) || is_water_tile(w, c)) return 0.0;
      double water_proximity = is_near_water(w, c) ? 0.3 : 0.0;
      double population_factor = get_nearby_population(w, c, 5000.0) < 10000 ? 0.7 : 0.0;
      return 0.4 * water_proximity + 0.6 * population_factor;
    }
    case UtilitySolar: case DomesticSolar: case CommercialSolar: {
      if (is_water_tile(w, c)) return 0.0;
      double terrain = get_terrain_suitability(w, c, t);
      return 0.6 * terrain + 0.4 * 0.8;
    }
    case HydroDam: case PumpedStorage: {
      if (!is_near_water(w, c) || is_urban_area(w, c)) return 0.0;
      double elevation = 0.0;
      double water_proximity = 0.8;
      return 0.5 * elevation + 0.5 * water_proximity;
    }
    default: {
      if (is_water_tile(w, c) || is_urban_area(w, c)) return 0.0;
      double terrain = get_terrain_suitability(w, c, t);
      return 0.7 * terrain + 0.3 * 0.5;
    }
  }
}

}  // namespace orc
