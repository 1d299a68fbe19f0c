"""Hand-derived known answers for the update rule, the 2025 year of a one-plant world, and the placement search.

Everything here is worked out from the REFERENCE'S RUST SOURCE (cited per step), in plain Python floats (IEEE double, one
operation per step, math.pow / math.exp = the platform libm like Rust's powf / exp), not from oracle/ and not from the
product: the oracle, the host library and the CUDA kernels are all held to these numbers (tests/test_known_answers.py),
which breaks the loop of comparing one restatement of the reference with another.
"""
import math

import numpy as np

from eirgrid_b200 import _abi

MIN_WEIGHT, MAX_WEIGHT = 0.0001, 0.999  # ai/learning/constants.rs:14-15

# action codes (include/eirgrid_b200.h): 3 * generator type + multiplier index; 45.. offsets; 60 DoNothing
ONSHORE, OFFSHORE, UTILITY_SOLAR, GAS_CC, GAS_PEAKER, BIOMASS, BATTERY, FOREST, DO_NOTHING = 0, 3, 12, 21, 24, 27, 36, 45, 60


def score_metrics(net, opinion, cost):
    """ai/metrics/scoring.rs:5-45, default mode"""
    normalized_cost = max(cost / 50000000000.0, 1.0)            # :9-10  MAX_ACCEPTABLE_COST
    log_cost = math.log(normalized_cost)                        # :12
    max_expected_log_cost = math.log(50000000000.0 * 100.0 / 50000000000.0)   # :13
    if net > 0.0:                                               # :18-22
        return 1.0 - min(net / 1000000.0, 1.0)
    cost_score = 1.0 - min(log_cost / max_expected_log_cost, 1.0)   # :27
    cost_weight = 0.8 if normalized_cost > 8.0 else 0.5         # :32-36
    opinion_weight = 1.0 - cost_weight
    return 1.0 + (cost_score * cost_weight + opinion * opinion_weight)   # :44


# ---- (a) one apply_contrast_learning + update_best_strategy + apply_deficit_contrast_learning step -------------------------
def record(rows):
    """rows: {year index: (deficit actions, additional actions)} -> one eg_traj record"""
    return _abi.pack_traj([(np.array(rows.get(y, ([], []))[0], np.uint8), np.array(rows.get(y, ([], []))[1], np.uint8)) for y in range(26)])


def result(net, opinion, cost, reliability=1.0):
    r = np.zeros(1, _abi.RESULT_DTYPE)
    r["net_emissions"], r["public_opinion"], r["total_cost"], r["power_reliability"] = net, opinion, cost, reliability
    r["score"] = score_metrics(net, opinion, cost)
    return r


def contrast_case():
    """Two episodes on fresh weights. Episode A becomes the best strategy (strategy.rs:58: no best yet); episode B is worse
    and goes through the whole rule. Returns (records, results, expected weights {(year, action): w},
    expected deficit weights {(year, deficit key): w})."""
    A = record({0: ([GAS_PEAKER, BATTERY], [ONSHORE]), 1: ([], [UTILITY_SOLAR, DO_NOTHING])})
    B = record({0: ([GAS_PEAKER, GAS_CC], [OFFSHORE, ONSHORE]), 1: ([], [DO_NOTHING, UTILITY_SOLAR, FOREST])})
    rA = result(-1000.0, 0.8, 2.0e10)      # net-zero, cost below MAX_ACCEPTABLE_COST: score 1 + (1 * 0.5 + 0.8 * 0.5) = 1.9
    rB = result(200000.0, 0.9, 1.0e10)     # still emitting: score 1 - 0.2 = 0.8
    best_score, current_score = score_metrics(-1000.0, 0.8, 2.0e10), score_metrics(200000.0, 0.9, 1.0e10)
    assert best_score == 1.9 and current_score == 0.8
    # initial table (weights/core.rs:35-152, learning/constants.rs:46-77)
    w = {(0, GAS_PEAKER): 0.02, (0, BATTERY): 0.07, (0, ONSHORE): 0.08, (0, GAS_CC): 0.06, (0, OFFSHORE): 0.08,
         (1, UTILITY_SOLAR): 0.08, (1, DO_NOTHING): 0.1, (1, FOREST): 0.02}
    dw = {(0, 0): 0.15, (0, 1): 0.15, (0, 2): 0.15}   # deficit keys 0 GasPeaker, 1 GasCombinedCycle, 2 BatteryStorage (core.rs:130-149)
    # episode A: apply_contrast_learning returns at once (no best), update_best_strategy stores A and sets the counter to 0,
    # apply_deficit_contrast_learning: deterioration 0 / 10 = 0 is not above the threshold 0.05 -> nothing. Table unchanged.
    # episode B, apply_contrast_learning (learning.rs:131-283) with iterations_without_improvement = 0:
    iwi = 0
    deterioration = (best_score - current_score) / best_score                            # :138-142
    threshold = 0.1 * max(math.exp(-float(iwi) / 500.0), 0.00001 / 0.1)                  # :153
    assert deterioration > threshold                                                     # :159
    stagnation = 1.0 + (0.2 * math.pow(float(iwi) / 10.0, 1.8))                          # :162-163  = 1
    combined = math.pow(deterioration, 0.3) * stagnation                                 # :167-170
    alr = 0.2 * (1.0 + 0.1 * float(iwi))                                                 # :173       = 0.2
    penalty = 1.0 / (1.0 + alr * 1.5 * combined)                                         # :176
    boost = 1.0 + (alr * 2.0 * stagnation)                                               # :179       = 1.4
    mild = 1.0 / (1.0 + alr * combined * 0.5)                                            # :245
    assert stagnation == 1.0 and boost == 1.4
    # year 2025: best = best_actions ++ best_deficit_actions = [GasPeaker, Battery, Onshore] ++ [GasPeaker, Battery]
    #            current = current_run_actions ++ current_deficit_actions = [GasPeaker, GasCC, Offshore, Onshore] ++ [GasPeaker, GasCC]
    # (simulation.rs:406-409: a deficit action is recorded in both lists)
    for a in (GAS_PEAKER, BATTERY, ONSHORE, GAS_PEAKER, BATTERY):                        # :222-226 boost every occurrence
        w[(0, a)] = min(w[(0, a)] * boost, MAX_WEIGHT)
    # :229-251, position i of current against best[i]
    #  i=0 GasPeaker: in best, best[0] is GasPeaker      -> nothing
    #  i=1 GasCC:     not in best                        -> penalty
    #  i=2 Offshore:  not in best                        -> penalty
    #  i=3 Onshore:   in best, best[3] is GasPeaker      -> mild penalty
    #  i=4 GasPeaker: in best, best[4] is Battery        -> mild penalty
    #  i=5 GasCC:     not in best (i >= len(best))       -> penalty
    w[(0, GAS_CC)] = max(w[(0, GAS_CC)] * penalty, MIN_WEIGHT)
    w[(0, OFFSHORE)] = max(w[(0, OFFSHORE)] * penalty, MIN_WEIGHT)
    w[(0, ONSHORE)] = max(w[(0, ONSHORE)] * mild, MIN_WEIGHT)
    w[(0, GAS_PEAKER)] = max(w[(0, GAS_PEAKER)] * mild, MIN_WEIGHT)
    w[(0, GAS_CC)] = max(w[(0, GAS_CC)] * penalty, MIN_WEIGHT)
    # year 2026: best = [UtilitySolar, DoNothing], current = [DoNothing, UtilitySolar, Forest]
    for a in (UTILITY_SOLAR, DO_NOTHING):
        w[(1, a)] = min(w[(1, a)] * boost, MAX_WEIGHT)
    w[(1, DO_NOTHING)] = max(w[(1, DO_NOTHING)] * mild, MIN_WEIGHT)       # i=0: in best, best[0] is UtilitySolar
    w[(1, UTILITY_SOLAR)] = max(w[(1, UTILITY_SOLAR)] * mild, MIN_WEIGHT)  # i=1: in best, best[1] is DoNothing
    w[(1, FOREST)] = max(w[(1, FOREST)] * penalty, MIN_WEIGHT)            # i=2: not in best
    # update_best_strategy (strategy.rs:19-258): 0.8 is not above 1.9 -> iterations_without_improvement = 1
    iwi = 1
    # apply_deficit_contrast_learning (learning.rs:285-373)
    d_det = float(iwi) / 10.0                                                            # :289
    d_thr = 0.05 * max(math.exp(-float(iwi) / 400.0), 0.00001 / 0.05)                    # :295-301
    assert d_det > d_thr
    d_stag = 1.0 + (0.2 * math.pow(float(iwi) / 10.0, 1.8))
    d_comb = math.pow(d_det, 0.3) * d_stag
    d_alr = 0.2 * (1.0 + 0.1 * float(iwi))
    d_pen = 1.0 / (1.0 + d_alr * 1.5 * d_comb)
    d_boost = 1.0 + (d_alr * 2.0 * d_stag * 1.5)
    # 2025: best_deficit_actions = [GasPeaker, Battery] boosted; current deficit [GasPeaker, GasCC]: GasCC is not in best -> penalty
    dw[(0, 0)] = min(dw[(0, 0)] * d_boost, MAX_WEIGHT)
    dw[(0, 2)] = min(dw[(0, 2)] * d_boost, MAX_WEIGHT)
    dw[(0, 1)] = max(dw[(0, 1)] * d_pen, MIN_WEIGHT)
    recs = np.concatenate([np.atleast_1d(A), np.atleast_1d(B)])
    ress = np.concatenate([rA, rB])
    return recs, ress, w, dw


# ---- (b) + (c): a toy world, one Biomass plant built by the 2025 deficit handler ---------------------------------------------
TOY = dict(sx=[10000.0, 30000.0], sy=[10000.0, 20000.0], pop=[20000, 10000],
           coast=[(0.0, 0.0), (50000.0, 0.0), (50000.0, 50000.0), (0.0, 50000.0)])


def toy_map_arrays():
    """arguments of Context.map_set / oracle World.from_arrays: two settlements, no existing plants, a square island"""
    return (np.array(TOY["sx"]), np.array(TOY["sy"]), np.array(TOY["pop"], np.uint32), np.zeros(0), np.zeros(0),
            np.zeros(0, np.uint8), np.zeros(0), np.array([c[0] for c in TOY["coast"]]), np.array([c[1] for c in TOY["coast"]]), 51, 1000.0)


def toy_record():
    """2025: the deficit handler's one action is AddGenerator(Biomass, 100 %); nothing else is recorded"""
    return np.atleast_1d(record({0: ([BIOMASS], [])}))


def placement_site():
    """MetalLocationSearch::find_suitable_location, CPU fallback (gpu/metal_location_search.rs:110-176) for the first plant:
    argmax over i, j in [0, 100) of prod_s (1 + pop_s / 1e6) / (1 + d_s / 1e4) * (1 - 1.0 * 0.1), first strict maximum.
    Returns (site index i * 51 + j, x, y)."""
    best, best_ij = 0.0, None
    for i in range(100):
        for j in range(100):
            x, y = min(max(i * 1000.0, 0.0), 50000.0), min(max(j * 1000.0, 0.0), 50000.0)   # Coordinate::new clamp, data/poi.rs:11-15
            score = 1.0
            for sx, sy, pop in zip(TOY["sx"], TOY["sy"], TOY["pop"]):
                distance = math.sqrt((x - sx) ** 2 + (y - sy) ** 2)
                score *= (1.0 + pop / 1000000.0) / (1.0 + distance / 10000.0)
            score *= 1.0 - (1.0 * 0.1)
            if score > best:
                best, best_ij = score, (i, j)
    i, j = best_ij
    return i * 51 + j, i * 1000.0, j * 1000.0


def toy_2025_metrics():
    """YearlyMetrics of 2025 (analysis/metrics_calculation.rs:32-175) after the one Biomass plant"""
    site, px, py = placement_site()
    m = {}
    m["total_population"] = 20000 + 10000                                             # map_handler.rs:813-817
    per_capita = 0.001 * math.pow(1.0 + 0.02, 0.0)                                    # const_funcs.rs:17-26 at 2025
    usage = 20000.0 * per_capita + 10000.0 * per_capita                               # settlements_loader.rs:30, summed in order
    m["total_power_usage"] = usage * (1.0 + (2025.0 - 2024.0) * 0.02)                 # map_handler.rs:819-827
    m["total_power_generation"] = 50.0 * 0.99 * 1.0                                   # generator.rs:528: power_out * efficiency * operation
    m["power_balance"] = m["total_power_generation"] - m["total_power_usage"]
    # calc_new_generator_opinion (map_handler.rs:925-949)
    settlement_opinions = 0.0
    for sx, sy in zip(TOY["sx"], TOY["sy"]):
        distance = math.sqrt((sx - px) ** 2 + (sy - py) ** 2)
        settlement_opinions += 1.0 / (1.0 + distance / 10000.0)                       # settlement.rs:103-106
    avg_settlement_opinion = settlement_opinions / 2.0
    type_opinion = min(max(0.60 + 0.001 * 0.0, 0.0), 1.0)                             # const_funcs.rs:78-93 (Biomass arm)
    inflation = math.pow(1.0 + 0.0185, 0)                                             # const_funcs.rs:13-15
    cost = 150000000.0 * inflation * math.pow(0.99, 0.0) * 1.0                        # const_funcs.rs:28-57: base * inflation * technology * location
    cost = cost * 1.0                                                                 # generator.rs:593 construction_cost_multiplier (100 %)
    normalized = cost / (1384000000.0 * inflation)                                    # const_funcs.rs:95-106
    cost_opinion = 1.0 - normalized
    m["average_public_opinion"] = 0.03 * avg_settlement_opinion + 0.12 * type_opinion + 0.82 * cost_opinion
    m["yearly_capital_cost"] = cost                                                   # map_handler.rs:968-985 (build year == 2025)
    m["total_capital_cost"] = cost
    m["inflation_factor"] = inflation
    m["total_co2_emissions"] = 1500.0 * 1.0 * 1.0 * (1.0 - (0.99 - 0.99))             # actions.rs:50-56, generator.rs:618-626
    m["total_carbon_offset"] = 0.0
    m["net_co2_emissions"] = m["total_co2_emissions"] - 0.0
    m["yearly_carbon_credit_revenue"] = 0.0                                           # const_funcs.rs:206-211: net >= 0
    m["yearly_energy_sales_revenue"] = m["power_balance"] * 8.76 * 50000.0            # const_funcs.rs:224-237
    m["yearly_total_cost"] = m["yearly_capital_cost"] + 0.0 + 0.0 - 0.0 - m["yearly_energy_sales_revenue"]   # metrics_calculation.rs:124-125
    m["total_cost"] = m["yearly_total_cost"]
    m["active_generators"] = 1
    return site, m


# ---- (d): the same world with three carbon offsets added in 2025; years 2025 and 2026 ---------------------------------------
FOREST_100, ACTIVE_CAPTURE_120, CARBON_CREDIT_150 = 45, 45 + 3 * 2 + 1, 45 + 3 * 3 + 2


def toy_offsets_record():
    """2025: Biomass by the deficit handler, then AddCarbonOffset(Forest, 100 %), (ActiveCapture, 120 %), (CarbonCredit, 150 %)"""
    return np.atleast_1d(record({0: ([BIOMASS], [FOREST_100, ACTIVE_CAPTURE_120, CARBON_CREDIT_150])}))


def carbon_price(year):
    """const_funcs.rs:186-203 with constants.rs:258-267"""
    if year < 2030:
        return 75.0
    if year < 2040:
        return 75.0 + ((year - 2030) / 10.0) * (130.0 - 75.0)
    if year <= 2050:
        return 130.0 + ((year - 2040) / 10.0) * (300.0 - 130.0)
    return 300.0


def toy_offsets_metrics():
    """YearlyMetrics of 2025 and 2026 (metrics_calculation.rs:32-175) for toy_offsets_record()"""
    site, px, py = placement_site()
    out = []
    prev = None
    pops = [20000, 10000]
    for year in (2025, 2026):
        m = {}
        if year > 2025:
            pops = [int(math.floor(p * 1.01 + 0.5)) for p in pops]                       # simulation.rs:110-113, f64::round
        per_capita = 0.001 * math.pow(1.0 + 0.02, float(year - 2025))                     # const_funcs.rs:17-26
        usage = 0.0
        for p in pops:
            usage += float(p) * per_capita                                                # simulation.rs:116-118 / settlements_loader.rs:30
        m["total_population"] = sum(pops)
        m["total_power_usage"] = usage * (1.0 + (float(year) - 2024.0) * 0.02)            # map_handler.rs:819-827
        m["total_power_generation"] = 50.0 * 0.99 * 1.0
        m["power_balance"] = m["total_power_generation"] - m["total_power_usage"]
        inflation = (1.0 + 0.0185) ** (year - 2025)                                       # powi, const_funcs.rs:13-15
        m["inflation_factor"] = inflation
        # emissions: the Biomass plant; offsets in insertion order (carbon_offset.rs:210-232, completion year 2025: delays are off)
        m["total_co2_emissions"] = 1500.0 * 1.0 * 1.0 * (1.0 - (0.99 - 0.99))
        maturity = min(max(1.0 - math.exp(-0.1 * float(year - 2025)), 0.0), 1.0)
        offset = 0.0
        offset += 500.0 * 25.0 * 0.85 * maturity          # Forest, 500 ha
        offset += 100.0 * 500.0 * 0.85 * 1.0              # ActiveCapture, 100 units
        offset += 1000.0 * 100.0 * 0.85 * 1.0             # CarbonCredit, 1000 units
        m["total_carbon_offset"] = offset
        m["net_co2_emissions"] = m["total_co2_emissions"] - offset
        credit = (-m["net_co2_emissions"]) * carbon_price(year) if m["net_co2_emissions"] < 0.0 else 0.0   # const_funcs.rs:206-219
        m["yearly_carbon_credit_revenue"] = credit

        def capital(y):  # calc_total_capital_cost(y): new plants, then offsets (map_handler.rs:951-965)
            infl = (1.0 + 0.0185) ** (y - 2025)
            gen = 150000000.0 * infl * math.pow(0.99, float(y - 2025)) * 1.0   # const_funcs.rs:56
            gen = gen * 1.0                                                     # generator.rs:593
            offs = 0.0
            for base, mult in ((1000000.0, 1.0), (1000000000.0, 1.2), (50000000.0, 1.5)):
                offs += (base * infl) * mult                                    # carbon_offset.rs:188-195
            return gen, gen + offs
        gen_cost, total_capital = capital(year)
        m["total_capital_cost"] = total_capital
        # 2025: plants built this year + offsets whose id "starts" in 2025 (always: quirk Q6); later: difference of the re-priced totals
        m["yearly_capital_cost"] = total_capital if year == 2025 else total_capital - capital(year - 1)[1]
        # opinion of the one plant
        settlement_opinions = 0.0
        for sx, sy in zip(TOY["sx"], TOY["sy"]):
            settlement_opinions += 1.0 / (1.0 + math.sqrt((sx - px) ** 2 + (sy - py) ** 2) / 10000.0)
        type_opinion = min(max(0.60 + 0.001 * float(year - 2025), 0.0), 1.0)
        normalized = gen_cost / (1384000000.0 * inflation)
        cost_opinion = 1.0 - normalized
        m["average_public_opinion"] = 0.03 * (settlement_opinions / 2.0) + 0.12 * type_opinion + 0.82 * cost_opinion
        sales = m["power_balance"] * 8.76 * 50000.0 if m["power_balance"] > 0.0 else 0.0
        m["yearly_energy_sales_revenue"] = sales
        m["yearly_total_cost"] = m["yearly_capital_cost"] + 0.0 + 0.0 - credit - sales
        m["total_cost"] = m["yearly_total_cost"] if prev is None else prev["total_cost"] + m["yearly_total_cost"]
        m["total_carbon_credit_revenue"] = credit if prev is None else prev["total_carbon_credit_revenue"] + credit
        m["total_energy_sales_revenue"] = sales if prev is None else prev["total_energy_sales_revenue"] + sales
        m["active_generators"] = 1
        out.append(m)
        prev = m
    return site, out
