// suitability.cu — location suitability for every analysis point x 15 generator types (sm_100a).
//
// Replaces LocationAnalysis::analyze_map (utils/map_handler.rs:61-142) with Map::calculate_generator_suitability
// (map_handler.rs:1319-1396) and its helpers is_water_tile / is_urban_area / is_near_water / is_coastal_region /
// get_distance_to_nearest_land / get_nearby_population / get_terrain_suitability (map_handler.rs:1178-1280,1398-1433)
// and is_point_inside_polygon (config/const_funcs.rs:143-158). It is the CUDA counterpart of the never-dispatched
// Metal kernel computeSuitability (aiSimulator/assets/metal_location_search.metal:239-258).
// One warp per point: the 441 nearest-land probes (the dominant cost, 441 x n_coast edge tests) are split over the
// lanes; lane 0 then evaluates the 15 type rules, sharing the point's predicates. Compiled with --fmad=false.
#include "suitability.cuh"

namespace {

struct Poly {
  const double* x;
  const double* y;
  int n;
};

__device__ __forceinline__ double clamp_map(double v) { return fmin(fmax(v, 0.0), 50000.0); }  // Coordinate::new

__device__ bool inside_polygon(double px, double py, const Poly& poly) {  // const_funcs.rs:143-158
  bool inside = false;
  if (poly.n == 0) return false;
  int j = poly.n - 1;
  for (int i = 0; i < poly.n; i++) {
    const double xi = __ldg(&poly.x[i]), yi = __ldg(&poly.y[i]), xj = __ldg(&poly.x[j]), yj = __ldg(&poly.y[j]);
    if (((yi > py) != (yj > py)) && (px < (xj - xi) * (py - yi) / (yj - yi) + xi)) inside = !inside;
    j = i;
  }
  return inside;
}
__device__ __forceinline__ bool water_tile(double px, double py, const Poly& poly) { return !inside_polygon(px, py, poly); }

// any of the 9 probes (x, y in {-d, 0, +d}, each clamped by Coordinate::new) lies on water; lanes 0..8 take one probe each
__device__ bool any9_water(double px, double py, double d, const Poly& poly, int lane) {
  bool w = false;
  if (lane < 9) {
    const int ix = lane / 3 - 1, iy = lane % 3 - 1;
    w = water_tile(clamp_map(px + ((double)ix * d)), clamp_map(py + ((double)iy * d)), poly);
  }
  return __any_sync(0xFFFFFFFFu, w);
}

__global__ void __launch_bounds__(128) eg_suitability_kernel(const EgSuitabilityParams p) {
  const int lane = threadIdx.x & 31;
  const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= p.n) return;  // whole warps exit together
  const uint32_t pt = p.first + k;
  const int side = 2 * p.half + 1;
  const int i = (int)(pt / side) - p.half, j = (int)(pt % side) - p.half;
  const double px = clamp_map((double)i * p.step), py = clamp_map((double)j * p.step);
  const Poly poly{p.cx, p.cy, p.n_coast};

  const bool water = water_tile(px, py, poly);
  const bool near_water = any9_water(px, py, 5000.0, poly, lane);   // is_near_water, :1211-1226
  const bool coastal = any9_water(px, py, 8000.0, poly, lane);      // is_coastal_region, :1178-1193

  // get_distance_to_nearest_land (:1398-1421): 21 x 21 probes at 1 km, lanes stride over them
  double min_distance = 1.7976931348623157e308;
  if (water) {
    for (int q = lane; q < 441; q += 32) {
      const int a = q / 21 - 10, b = q % 21 - 10;
      const double x = px + ((double)a * 1000.0), y = py + ((double)b * 1000.0);
      if (x >= 0.0 && x <= 50000.0 && y >= 0.0 && y <= 50000.0) {
        const double tx = clamp_map(x), ty = clamp_map(y);
        if (!water_tile(tx, ty, poly)) {
          const double dx = px - tx, dy = py - ty;
          min_distance = fmin(min_distance, sqrt(dx * dx + dy * dy));
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) min_distance = fmin(min_distance, __shfl_xor_sync(0xFFFFFFFFu, min_distance, o));
  }

  // is_urban_area (:1199-1209), get_nearby_population(5000) (:1423-1433) and the OnshoreWind neighbour penalty
  bool urban = false;
  unsigned int nearby_pop = 0;
  for (int s = lane; s < p.n_settlements; s += 32) {
    const double dx = __ldg(&p.sx[s]) - px, dy = __ldg(&p.sy[s]) - py;
    const double distance = sqrt(dx * dx + dy * dy);
    const unsigned int pop = __ldg(&p.pop[s]);
    if (distance < sqrt((double)pop) * 5.0) urban = true;
    if (distance <= 5000.0) nearby_pop += pop;
  }
  urban = __any_sync(0xFFFFFFFFu, urban);
  for (int o = 16; o > 0; o >>= 1) nearby_pop += __shfl_xor_sync(0xFFFFFFFFu, nearby_pop, o);

  if (lane != 0) return;
  // the neighbour penalty is a float sum in generator order: kept sequential on one lane
  double nearby_penalty = 0.0;
  for (int g = 0; g < p.n_generators; g++) {
    const double dx = __ldg(&p.gx[g]) - px, dy = __ldg(&p.gy[g]) - py;
    const double d = sqrt(dx * dx + dy * dy);
    if (d < 3000.0) nearby_penalty += 0.1 / (1.0 + d);
  }
  double* out = p.scores + (size_t)k * 15;
  // OnshoreWind (:1324-1339)
  out[0] = (urban ? 0.0 : (coastal ? 0.7 : 0.5)) - nearby_penalty;
  // OffshoreWind / TidalGenerator / WaveEnergy (:1340-1356)
  double marine = 0.0;
  if (water) {
    const double depth_factor = 0.8;
    const double distance_factor = min_distance < 2000.0 ? 0.3 : (min_distance > 10000.0 ? 0.5 : 0.7);
    marine = depth_factor * distance_factor;
  }
  out[1] = marine; out[13] = marine; out[14] = marine;
  // solar (:1367-1376); terrain only differs for UtilitySolar (:1237-1242, elevation == 0)
  const double sunlight = 0.8;
  out[2] = water ? 0.0 : 0.6 * 1.0 + 0.4 * sunlight;
  out[3] = out[2];
  out[4] = water ? 0.0 : 0.6 * (!near_water ? 1.2 : 1.0) + 0.4 * sunlight;
  // Nuclear (:1357-1366)
  if (urban || water) out[5] = 0.0;
  else {
    const double water_proximity = near_water ? 0.3 : 0.0;
    const double population_factor = nearby_pop < 10000u ? 0.7 : 0.0;
    out[5] = 0.4 * water_proximity + 0.6 * population_factor;
  }
  // HydroDam / PumpedStorage (:1377-1386)
  const double hydro = (!near_water || urban) ? 0.0 : 0.5 * 0.0 + 0.5 * 0.8;
  out[10] = hydro; out[11] = hydro;
  // everything else (:1387-1394): CoalPlant, GasCombinedCycle, GasPeaker, Biomass, BatteryStorage
  const double other = (water || urban) ? 0.0 : 0.7 * 1.0 + 0.3 * 0.5;
  out[6] = other; out[7] = other; out[8] = other; out[9] = other; out[12] = other;
}

}  // namespace

cudaError_t eg_launch_suitability(const EgSuitabilityParams& p, cudaStream_t stream) {
  if (p.n == 0) return cudaSuccess;
  const uint64_t threads = (uint64_t)p.n * 32;
  const uint32_t blocks = (uint32_t)((threads + 127) / 128);
  eg_suitability_kernel<<<blocks, 128, 0, stream>>>(p);
  return cudaGetLastError();
}
