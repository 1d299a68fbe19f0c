// suitability.cu — location suitability for every analysis point x 15 generator types x simulated years (sm_100a).
//
// Replaces LocationAnalysis::analyze_map (utils/map_handler.rs:61-142) with Map::calculate_generator_suitability
// (map_handler.rs:1319-1396) and its helpers is_water_tile / is_urban_area / is_near_water / is_coastal_region /
// get_distance_to_nearest_land / get_nearby_population / get_terrain_suitability (map_handler.rs:1178-1280,1398-1433)
// and is_point_inside_polygon (config/const_funcs.rs:143-158). It is the CUDA counterpart of the never-dispatched
// Metal kernel computeSuitability (aiSimulator/assets/metal_location_search.metal:239-258).
//
// One warp per point, persistent blocks. What dominates is the point-in-polygon test: 1 + 9 + 9 probes per point and, for a
// point on water, the 21 x 21 probes of the nearest-land scan, each against every edge of the coastline polygon. The polygon
// (and the settlement coordinates) are therefore staged ONCE per block into shared memory with bulk asynchronous copies
// (cp.async.bulk + mbarrier) and every probe of every point the block handles reads them there; the probes of a point are
// dealt out to the lanes, so a polygon edge is one broadcast read per warp.
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

// any of the 9 probes (x, y in {-d, 0, +d}, each clamped by Coordinate::new) lies on water; lanes 0..8 take one probe each
__device__ __forceinline__ bool any9_water(double px, double py, double d, const Poly& poly, int lane) {
  bool w = false;
  if (lane < 9) {
    const int ix = lane / 3 - 1, iy = lane % 3 - 1;
    w = water_tile(clamp_map(px + ((double)ix * d)), clamp_map(py + ((double)iy * d)), poly);
  }
  return __any_sync(0xFFFFFFFFu, w);
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// bulk copy global -> shared, completion counted on the mbarrier (sizes are multiples of 16 bytes)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

template <bool STAGED>
__global__ void __launch_bounds__(128) eg_suitability_kernel(const EgSuitabilityParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x & 31;
  Poly poly{p.cx, p.cy, p.n_coast};
  const double* sx = p.sx;
  const double* sy = p.sy;
  if (STAGED) {
    // layout: coast x | coast y | settlement x | settlement y, each padded to a multiple of 16 bytes (the device arrays are too)
    const uint32_t cb = (uint32_t)((p.n_coast * 8 + 15) & ~15), sb = (uint32_t)((p.n_settlements * 8 + 15) & ~15);
    double* s_cx = (double*)smem_raw;
    double* s_cy = (double*)(smem_raw + cb);
    double* s_sx = (double*)(smem_raw + 2 * cb);
    double* s_sy = (double*)(smem_raw + 2 * cb + sb);
    const uint32_t b = smem_addr(&bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(2 * cb + 2 * sb) : "memory");
      if (cb) { bulk_g2s(s_cx, p.cx, cb, b); bulk_g2s(s_cy, p.cy, cb, b); }
      if (sb) { bulk_g2s(s_sx, p.sx, sb, b); bulk_g2s(s_sy, p.sy, sb, b); }
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n selp.u32 %0, 1, 0, q;\n}" : "=r"(done) : "r"(b) : "memory");
    poly.x = s_cx; poly.y = s_cy;
    sx = s_sx; sy = s_sy;
  }
  const int side = p.mode == 0 ? 2 * p.half + 1 : p.side;
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < p.n; k += warps) {
    const uint32_t pt = p.first + k;
    // analyze_map: i, j in [-half, half] (negative coordinates clamp to 0); candidate-site grid: i, j in [0, side)
    const int i = (int)(pt / side) - (p.mode == 0 ? p.half : 0), j = (int)(pt % side) - (p.mode == 0 ? p.half : 0);
    const double px = clamp_map((double)i * p.step), py = clamp_map((double)j * p.step);

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

    // is_urban_area (:1199-1209) and get_nearby_population(5000) (:1423-1433) for every requested year: the distance to a
    // settlement is computed once; bit y of `urban` / entry y of `nearby` belong to year year_first + y
    uint32_t urban = 0u;
    uint32_t nearby[EG_SUIT_MAX_YEARS];
#pragma unroll
    for (int y = 0; y < EG_SUIT_MAX_YEARS; y++) nearby[y] = 0u;
    for (int s = lane; s < p.n_settlements; s += 32) {
      const double dx = sx[s] - px, dy = sy[s] - py;
      const double distance = sqrt(dx * dx + dy * dy);
      if (distance < p.urban_r_max) {
#pragma unroll
        for (int y = 0; y < EG_SUIT_MAX_YEARS; y++)
          if (y < p.n_years && distance < __ldg(&p.urban_r[(size_t)(p.year_first + y) * p.n_settlements + s])) urban |= 1u << y;
      }
      if (distance <= 5000.0) {
#pragma unroll
        for (int y = 0; y < EG_SUIT_MAX_YEARS; y++)
          if (y < p.n_years) nearby[y] += __ldg(&p.pop[(size_t)(p.year_first + y) * p.n_settlements + s]);
      }
    }
    urban = __reduce_or_sync(0xFFFFFFFFu, urban);
    uint32_t my_nearby = 0u;
#pragma unroll
    for (int y = 0; y < EG_SUIT_MAX_YEARS; y++) {
      const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, nearby[y]);
      if (lane == y) my_nearby = total;
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
        const double d = sqrt(dx * dx + dy * dy);
        hit = d < 3000.0;
        term = 0.1 / (1.0 + d);
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

cudaError_t eg_launch_suitability(const EgSuitabilityParams& p, cudaStream_t stream) {
  if (p.n == 0 || p.n_years == 0) return cudaSuccess;
  if (p.n_years > EG_SUIT_MAX_YEARS) return cudaErrorInvalidValue;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint32_t warps_needed = p.n;
  // persistent blocks of 4 warps, 8 blocks per SM at most; fewer when there are fewer points than warps
  const uint32_t blocks = (uint32_t)std::min<uint64_t>((uint64_t)sms * 8, ((uint64_t)warps_needed + 3) / 4);
  const size_t staged_bytes = 2 * (size_t)((p.n_coast * 8 + 15) & ~15) + 2 * (size_t)((p.n_settlements * 8 + 15) & ~15);
  if (staged_bytes <= 96 * 1024) {
    static bool opted = false;
    if (!opted) {
      cudaError_t e = cudaFuncSetAttribute(eg_suitability_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) return e;
      opted = true;
    }
    eg_suitability_kernel<true><<<blocks, 128, staged_bytes, stream>>>(p);
  } else {
    eg_suitability_kernel<false><<<blocks, 128, 0, stream>>>(p);  // a coastline too long for shared memory is read through L1
  }
  return cudaGetLastError();
}
