// Harness (tests/test_host_tables.py, also under ASan/UBSan): the host-built tables the episode kernel's placement walk relies on,
// for the shipped map re-gridded to the sizes given as (grid_n step) pairs after the three asset files.
#include <cstdio>
#include <cstdlib>
#include <string>
#include "host_tables.hpp"
#include "common.hpp"
static std::string last;
int eg_fail(int code, const std::string& m) { last = m; return code; }
int main(int argc, char** argv) {
  EgHostMap m;
  int rc = eg_host_map_load(&m, argv[1], argv[2], argv[3]);
  if (rc) { std::printf("load rc=%d %s\n", rc, last.c_str()); return 1; }
  for (int a = 4; a + 1 < argc; a += 2) {
    m.grid_n = std::atoi(argv[a]);
    m.step = std::atof(argv[a + 1]);
    rc = eg_host_map_validate(m);
    if (rc) { std::printf("map %d %g rc=%d %s\n", m.grid_n, m.step, rc, last.c_str()); continue; }
    EgHostTables T;
    eg_host_build_tables(m, &T);
    std::printf("map %d %g geom=%d stride=%d entries=%d\n", m.grid_n, m.step, T.near_geom, T.r2_stride, T.r2_limit[2 * EG_N_RCLASS]);
    for (int rc2 = 0; rc2 < EG_N_RCLASS; rc2++) {
      const int lim = T.r2_limit[rc2], off = T.r2_limit[EG_N_RCLASS + rc2];
      // the factor of the first cell distance outside the radius (what a plant out of range multiplies by) and the last one inside
      std::printf("  rclass %d limit=%d offset=%d first=%.17g last_inside=%.17g at_limit=%.17g\n", rc2, lim, off,
                  T.near_factor[(size_t)rc2 * T.r2_stride], T.near_factor[(size_t)rc2 * T.r2_stride + lim - 1],
                  T.near_factor[(size_t)rc2 * T.r2_stride + lim]);
    }
    for (int t = 0; t < EG_NT; t++) {
      const EgSmallTables& S = T.small;
      std::printf("  type %d sums=%.17g,%.17g,%.17g,%.17g net_mw=%.17g co2=%.17g acc=%d info=%u,%u pclass=%d rclass=%d water=%d\n", t,
                  S.type_sums[t][0], S.type_sums[t][1], S.type_sums[t][2], S.type_sums[t][3], S.net_mw[t], S.co2[t], (int)S.acc_class[t],
                  S.place_info[t][0], S.place_info[t][1], (int)S.pclass[t], (int)S.rclass_of_pclass[S.pclass[t]],
                  (int)S.water_of_pclass[S.pclass[t]]);
    }
  }
  return 0;
}
