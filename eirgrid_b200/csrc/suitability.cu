// suitability.cu — location suitability for every analysis point x 15 generator types x simulated years (sm_100a).
//
// Replaces LocationAnalysis::analyze_map (utils/map_handler.rs:61-142) with Map::calculate_generator_suitability
// (map_handler.rs:1319-1396) and its helpers is_water_tile / is_urban_area / is_near_water / is_coastal_region /
// get_distance_to_nearest_land / get_nearby_population / get_terrain_suitability (map_handler.rs:1178-1280,1398-1433)
// and is_point_inside_polygon (config/const_funcs.rs:143-158). It is the CUDA counterpart of the never-dispatched
// Metal kernel computeSuitability (aiSimulator/assets/metal_location_search.metal:239-258).
//
// One warp per point. What dominates in the reference is the point-in-polygon test: 1 + 9 + 9 probes per point and, for a
// point on water, the 21 x 21 probes of the nearest-land scan, each against every edge of the coastline polygon. Here the
// edges a horizontal probe line crosses are found once per (row of sites, probe row) — "crossing lists" below — and a block
// handles sites of ONE row: it stages that row's 25 lists, the settlement coordinates and (for the rare overlong list) the
// polygon into shared memory with bulk asynchronous copies (cp.async.bulk + mbarrier); the probes of a point are dealt out
// to the lanes and each compares its x with a handful of abscissae.
// The geometry of a point does not depend on the year; only the settlements' populations do (urban test, nearby-population
// rule). One pass therefore serves all requested years: lane y evaluates the 15 type rules of year y.
// Compiled with --fmad=false: every score is the reference's IEEE arithmetic.
#include "suitability.cuh"
#include <algorithm>

namespace {

__device__ __forceinline__ double clamp_map(double v) { return fmin(fmax(v, 0.0), 50000.0); }  // Coordinate::new

struct Poly {
  const double* x;
  const double* y;
  int n;
};

__device__ __forceinline__ bool inside_polygon(double px, double py, const Poly& poly) {  // const_funcs.rs:143-158
  bool inside = false;
  if (poly.n == 0) return false;
  double xj = poly.x[poly.n - 1], yj = poly.y[poly.n - 1];
  for (int i = 0; i < poly.n; i++) {
    const double xi = poly.x[i], yi = poly.y[i];
    if (((yi > py) != (yj > py)) && (px < (xj - xi) * (py - yi) / (yj - yi) + xi)) inside = !inside;
    xj = xi; yj = yi;
  }
  return inside;
}
__device__ __forceinline__ bool water_tile(double px, double py, const Poly& poly) { return !inside_polygon(px, py, poly); }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// bulk copy global -> shared, completion counted on the mbarrier (sizes are multiples of 16 bytes)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// ---- crossing lists ---------------------------------------------------------------------------------------------------
// is_point_inside_polygon toggles on every edge i with (yi > y) != (yj > y) whose intersection abscissa
// X_i(y) = (xj - xi) * (y - yi) / (yj - yi) + xi lies to the right of the probe. Both the set of such edges and the values
// X_i(y) depend on the probe's y alone, and the probes of a row of sites use only 25 different y: the site's own, +-5 km
// and +-8 km (clamped), and the 21 rows of the nearest-land scan. They are evaluated ONCE per (site row, probe row) by
// eg_suit_rows_kernel — the same expression, so the same bits, as inside the reference's loop — sorted, and a probe then
// finds by binary search how many of them lie to its right instead of testing every edge: the parity of that count is the
// reference's result.
constexpr int kSlots = 25;                 // probe rows per site row
constexpr int kListStride = 129;           // doubles per list: [count | up to 128 abscissae, ascending]; odd, so lanes on different lists hit different banks
constexpr int kMaxCrossings = 128;         // (the shipped coastline is 200 unordered points: a horizontal line crosses it ~50 times, at most 96)
constexpr int kRowDoubles = kSlots * kListStride + 1;   // 3226 doubles = 25,808 bytes, a multiple of 16 (bulk copy)
static_assert((kRowDoubles * 8) % 16 == 0, "a site row's lists are staged by one bulk copy");
constexpr int kSitesPerBlock = 64;

// y of probe row `slot` for a site row at py, and whether the nearest-land scan may use it (it skips rows outside the map)
__device__ __forceinline__ double slot_y(double py, int slot, bool* in_map) {
  if (slot < 21) {
    const double y = py + ((double)(slot - 10) * 1000.0);
    *in_map = y >= 0.0 && y <= 50000.0;
    return clamp_map(y);
  }
  *in_map = true;
  const double d = slot < 23 ? 5000.0 : 8000.0;
  return clamp_map(py + ((slot & 1) ? -1.0 : 1.0) * d);   // slots 21 / 23: -d, slots 22 / 24: +d
}

__device__ __forceinline__ double site_coord(const EgSuitabilityParams& p, int index) {
  return clamp_map((double)(index - (p.mode == 0 ? p.half : 0)) * p.step);
}

__global__ void __launch_bounds__(128) eg_suit_rows_kernel(const EgSuitabilityParams p, int side, double* rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= side * kSlots) return;
  const int j = t / kSlots, slot = t - j * kSlots;
  bool in_map;
  const double y = slot_y(site_coord(p, j), slot, &in_map);
  double* list = rows + (size_t)j * kRowDoubles + slot * kListStride;
  double sorted[kMaxCrossings];  // thread-local (L1-resident) while the list is built, written out once
  int c = 0;
  if (p.n_coast > 0) {
    double xj = p.cx[p.n_coast - 1], yj = p.cy[p.n_coast - 1];
    for (int i = 0; i < p.n_coast; i++) {
      const double xi = p.cx[i], yi = p.cy[i];
      if ((yi > y) != (yj > y)) {
        if (c < kMaxCrossings) {  // insertion into the ascending list
          const double x = (xj - xi) * (y - yi) / (yj - yi) + xi;
          int t = c;
          while (t > 0 && sorted[t - 1] > x) { sorted[t] = sorted[t - 1]; t--; }
          sorted[t] = x;
        }
        c++;
      }
      xj = xi; yj = yi;
    }
  }
  for (int t = 0; t < min(c, kMaxCrossings); t++) list[1 + t] = sorted[t];
  list[0] = c <= kMaxCrossings ? (double)c : -1.0;   // -1: more crossings than a list holds, probes of this row test every edge
  if (slot == 0) rows[(size_t)j * kRowDoubles + kSlots * kListStride] = 0.0;  // the pad element
}

// water test of one probe from its row's crossing list (inside = odd number of crossings to the right)
__device__ __forceinline__ bool water_from_list(const double* lists, int slot, double x, double y, const Poly& poly) {
  const double* list = lists + slot * kListStride;
  const int c = (int)list[0];
  if (c < 0) return water_tile(x, y, poly);
  // crossings to the right of the probe: the entries with x < X; the list is ascending
  int lo = 0, hi = c;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (x < list[1 + mid]) hi = mid;
    else lo = mid + 1;
  }
  return ((c - lo) & 1) == 0;   // inside = odd count; water = not inside
}

// One block = one row of sites (same y) x up to 64 consecutive site rows in x; one warp per site.
__global__ void __launch_bounds__(128) eg_suitability_kernel(const EgSuitabilityParams p, const double* rows, int side, int i_lo, int i_hi) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x % side, chunk = blockIdx.x / side;
  // staged per block: this site row's 25 crossing lists | coast x | coast y | settlement x | settlement y
  const uint32_t lb = kRowDoubles * 8, cb = (uint32_t)((p.n_coast * 8 + 15) & ~15), sb = (uint32_t)((p.n_settlements * 8 + 15) & ~15);
  double* s_lists = (double*)smem_raw;
  double* s_cx = (double*)(smem_raw + lb);
  double* s_cy = (double*)(smem_raw + lb + cb);
  double* s_sx = (double*)(smem_raw + lb + 2 * cb);
  double* s_sy = (double*)(smem_raw + lb + 2 * cb + sb);
  const uint32_t b = smem_addr(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(lb + 2 * cb + 2 * sb) : "memory");
    bulk_g2s(s_lists, rows + (size_t)j * kRowDoubles, lb, b);
    if (cb) { bulk_g2s(s_cx, p.cx, cb, b); bulk_g2s(s_cy, p.cy, cb, b); }
    if (sb) { bulk_g2s(s_sx, p.sx, sb, b); bulk_g2s(s_sy, p.sy, sb, b); }
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n selp.u32 %0, 1, 0, q;\n}" : "=r"(done) : "r"(b) : "memory");
  const Poly poly{s_cx, s_cy, p.n_coast};
  const double* sx = s_sx;
  const double* sy = s_sy;
  const double py = site_coord(p, j);
  for (int i = i_lo + chunk * kSitesPerBlock + warp; i <= i_hi && i < i_lo + (chunk + 1) * kSitesPerBlock; i += 4) {
    const long long pt = (long long)i * side + j;
    if (pt < (long long)p.first || pt >= (long long)p.first + p.n) continue;
    const uint32_t k = (uint32_t)(pt - p.first);
    const double px = site_coord(p, i);

    const bool water = water_from_list(s_lists, 10, px, py, poly);
    // is_near_water (:1211-1226) / is_coastal_region (:1178-1193): any of the 9 probes (x, y in {-d, 0, +d}, each clamped by
    // Coordinate::new) on water; lanes 0..8 take the 5 km probes, lanes 9..17 the 8 km probes
    bool w9 = false;
    if (lane < 18) {
      const int q = lane < 9 ? lane : lane - 9;
      const double d = lane < 9 ? 5000.0 : 8000.0;
      const int ix = q / 3 - 1, iy = q % 3 - 1;
      const double x = clamp_map(px + ((double)ix * d)), y = clamp_map(py + ((double)iy * d));
      const int slot = iy == 0 ? 10 : (lane < 9 ? (iy < 0 ? 21 : 22) : (iy < 0 ? 23 : 24));
      w9 = water_from_list(s_lists, slot, x, y, poly);
    }
    const unsigned w9_mask = __ballot_sync(0xFFFFFFFFu, w9);
    const bool near_water = (w9_mask & 0x1FFu) != 0, coastal = (w9_mask & 0x3FE00u) != 0;

    // get_distance_to_nearest_land (:1398-1421): 21 x 21 probes at 1 km, lanes stride over them
    double min_distance = 1.7976931348623157e308;
    if (water) {
      for (int q = lane; q < 441; q += 32) {
        const int a = q / 21 - 10, bb = q % 21 - 10;
        const double x = px + ((double)a * 1000.0), y = py + ((double)bb * 1000.0);
        if (x >= 0.0 && x <= 50000.0 && y >= 0.0 && y <= 50000.0) {
          const double tx = clamp_map(x), ty = clamp_map(y);
          if (!water_from_list(s_lists, bb + 10, tx, ty, poly)) {
            const double dx = px - tx, dy = py - ty;
            min_distance = fmin(min_distance, sqrt(dx * dx + dy * dy));
          }
        }
      }
      for (int o = 16; o > 0; o >>= 1) min_distance = fmin(min_distance, __shfl_xor_sync(0xFFFFFFFFu, min_distance, o));
    }

    // is_urban_area (:1199-1209) and get_nearby_population(5000) (:1423-1433) for every requested year: the distance to a
    // settlement is computed once; bit y of `urban` / entry y of `nearby` belong to year year_first + y
    uint32_t urban = 0u;
    uint32_t nearby[EG_SUIT_MAX_YEARS];
#pragma unroll
    for (int y = 0; y < EG_SUIT_MAX_YEARS; y++) nearby[y] = 0u;
    // (a settlement farther than every radius in play, with a metre of margin for the rounding of the square root, can
    //  satisfy neither test: its square root is not taken)
    const double reach = fmax(p.urban_r_max, 5000.0) + 1.0, reach2 = reach * reach;
    bool found = false;
    for (int s = lane; s < p.n_settlements; s += 32) {
      const double dx = sx[s] - px, dy = sy[s] - py;
      const double d2 = dx * dx + dy * dy;
      if (d2 < reach2) {
        const double distance = sqrt(d2);
        if (distance < p.urban_r_max) {
#pragma unroll
          for (int y = 0; y < EG_SUIT_MAX_YEARS; y++)
            if (y < p.n_years && distance < __ldg(&p.urban_r[(size_t)(p.year_first + y) * p.n_settlements + s])) urban |= 1u << y;
        }
        if (distance <= 5000.0) {
          found = true;
#pragma unroll
          for (int y = 0; y < EG_SUIT_MAX_YEARS; y++)
            if (y < p.n_years) nearby[y] += __ldg(&p.pop[(size_t)(p.year_first + y) * p.n_settlements + s]);
        }
      }
    }
    uint32_t my_nearby = 0u;
    if (__any_sync(0xFFFFFFFFu, found || urban != 0u)) {  // most sites have no settlement within reach: nothing to reduce
      urban = __reduce_or_sync(0xFFFFFFFFu, urban);
#pragma unroll
      for (int y = 0; y < EG_SUIT_MAX_YEARS; y++) {
        const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, nearby[y]);
        if (lane == y) my_nearby = total;
      }
    }

    // OnshoreWind's neighbour penalty (:1329-1337) is a float sum in generator order: the distances are computed by all
    // lanes, the terms of the plants in range are then added in index order (the same sum in every lane)
    double nearby_penalty = 0.0;
    for (int base = 0; base < p.n_generators; base += 32) {
      const int g = base + lane;
      double term = 0.0;
      bool hit = false;
      if (g < p.n_generators) {
        const double dx = __ldg(&p.gx[g]) - px, dy = __ldg(&p.gy[g]) - py;
        const double d2 = dx * dx + dy * dy;
        if (d2 < 3001.0 * 3001.0) {  // d < 3000 is impossible beyond 3001 m
          const double d = sqrt(d2);
          hit = d < 3000.0;
          term = 0.1 / (1.0 + d);
        }
      }
      unsigned hits = __ballot_sync(0xFFFFFFFFu, hit);
      while (hits) {
        const int src = __ffs(hits) - 1;
        hits &= hits - 1u;
        nearby_penalty += __shfl_sync(0xFFFFFFFFu, term, src);
      }
    }

    if (lane >= p.n_years) continue;
    const bool urban_y = (urban >> lane) & 1u;
    double* out = p.scores + ((size_t)k * p.n_years + lane) * 15;
    // OnshoreWind (:1324-1339)
    out[0] = (urban_y ? 0.0 : (coastal ? 0.7 : 0.5)) - nearby_penalty;
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
    if (urban_y || water) out[5] = 0.0;
    else {
      const double water_proximity = near_water ? 0.3 : 0.0;
      const double population_factor = my_nearby < 10000u ? 0.7 : 0.0;
      out[5] = 0.4 * water_proximity + 0.6 * population_factor;
    }
    // HydroDam / PumpedStorage (:1377-1386)
    const double hydro = (!near_water || urban_y) ? 0.0 : 0.5 * 0.0 + 0.5 * 0.8;
    out[10] = hydro; out[11] = hydro;
    // everything else (:1387-1394): CoalPlant, GasCombinedCycle, GasPeaker, Biomass, BatteryStorage
    const double other = (water || urban_y) ? 0.0 : 0.7 * 1.0 + 0.3 * 0.5;
    out[6] = other; out[7] = other; out[8] = other; out[9] = other; out[12] = other;
  }
}

}  // namespace

size_t eg_suitability_rows_bytes(int side) { return (size_t)side * kRowDoubles * sizeof(double); }

cudaError_t eg_launch_suitability(const EgSuitabilityParams& p, double* d_rows, cudaStream_t stream) {
  if (p.n == 0 || p.n_years == 0) return cudaSuccess;
  if (p.n_years > EG_SUIT_MAX_YEARS) return cudaErrorInvalidValue;
  const int side = p.mode == 0 ? 2 * p.half + 1 : p.side;
  const size_t staged_bytes = (size_t)kRowDoubles * 8 + 2 * (size_t)((p.n_coast * 8 + 15) & ~15) + 2 * (size_t)((p.n_settlements * 8 + 15) & ~15);
  if (staged_bytes > 200 * 1024) return cudaErrorInvalidValue;  // (a coastline of 10,000 points: not a map this path is for)
  eg_suit_rows_kernel<<<(side * kSlots + 127) / 128, 128, 0, stream>>>(p, side, d_rows);
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(eg_suitability_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    opted = true;
  }
  // site = i * side + j: the rows of sites i_lo..i_hi intersect the requested range
  const int i_lo = (int)(p.first / (uint32_t)side), i_hi = (int)(((uint64_t)p.first + p.n - 1) / (uint32_t)side);
  const int chunks = (i_hi - i_lo + kSitesPerBlock) / kSitesPerBlock;
  eg_suitability_kernel<<<(uint32_t)side * (uint32_t)chunks, 128, staged_bytes, stream>>>(p, d_rows, side, i_lo, i_hi);
  return cudaGetLastError();
}
