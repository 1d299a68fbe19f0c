// analyze_locations — the reference's location-analysis tool (aiSimulator/bin/analyze_locations.rs:1-46) on the B200 library.
//
//   analyze_locations [-m, --min-suitability <F>] [-o, --output-file <PATH>] [-c, --cache-dir <DIR>]
//
// Same flags and defaults (0.3, location_analysis.txt, cache). It analyses an EMPTY map like the reference's tool
// (Map::new: no settlements, no plants; the coastline is compiled into the reference and read from --assets here) and writes
// the text report and <cache-dir>/location_analysis.json, the file run_multi_simulation loads (core/multi_simulation.rs:149-154).
// Addition: --loaded-map analyses the map with its settlements and plants instead. Only the C ABI is used.
#include <cstdio>
#include <cstdlib>
#include <string>
#include "../include/eirgrid_b200.h"

static void die(const std::string& m) {
  std::fprintf(stderr, "error: %s\n", m.c_str());
  std::exit(2);
}

int main(int argc, char** argv) {
  double min_suitability = 0.3;
  std::string output_file = "location_analysis.txt", cache_dir = "cache", assets = "aiSimulator/assets";
  int loaded = 0, device = 0;
  for (int i = 1; i < argc; i++) {
    const std::string f = argv[i];
    auto need = [&]() -> std::string {
      if (i + 1 >= argc) die("missing value for " + f);
      return argv[++i];
    };
    if (f == "-m" || f == "--min-suitability") {
      const std::string v = need();
      char* end = nullptr;
      min_suitability = std::strtod(v.c_str(), &end);
      if (v.empty() || *end) die("invalid value '" + v + "' for '" + f + "'");
    } else if (f == "-o" || f == "--output-file") output_file = need();
    else if (f == "-c" || f == "--cache-dir") cache_dir = need();
    else if (f == "--assets") assets = need();
    else if (f == "--loaded-map") loaded = 1;
    else if (f == "--device") device = std::atoi(need().c_str());
    else if (f == "-h" || f == "--help") {
      std::puts("analyze_locations [-m, --min-suitability <F=0.3>] [-o, --output-file <PATH=location_analysis.txt>] [-c, --cache-dir <DIR=cache>]\n"
                "additions: --assets <DIR> --loaded-map --device <N>");
      return 0;
    } else die("unknown argument " + f);
  }
  eg_ctx* ctx = nullptr;
  if (eg_init(device, nullptr, &ctx) < 0) die(std::string("eg_init: ") + eg_last_error());
  if (eg_map_load(ctx, (assets + "/settlements.json").c_str(), (assets + "/ireland_generators.csv").c_str(),
                  (assets + "/coastline_points.json").c_str()) < 0)
    die(std::string("eg_map_load: ") + eg_last_error());
  std::puts("Starting location analysis...");
  std::printf("Minimum suitability threshold: %g\n", min_suitability);
  std::printf("\nSaving detailed results to %s...\n", output_file.c_str());
  std::printf("Saving location analysis cache to %s...\n", cache_dir.c_str());
  if (eg_location_analysis_write(ctx, loaded, min_suitability, cache_dir.c_str(), output_file.c_str()) < 0)
    die(std::string("eg_location_analysis_write: ") + eg_last_error());
  std::puts("Analysis complete!");
  eg_destroy(ctx);
  return 0;
}
