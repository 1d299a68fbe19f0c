// ASan/UBSan harness (tests/test_host_sanitizers.py): eg_update over records with arbitrary counts and action codes.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "weights.hpp"
int eg_fail(int code, const std::string&) { return code; }
int main() {
  eg_weights* w = nullptr;
  eg_weights_new(&w);
  std::vector<eg_result> res(512);
  std::vector<eg_traj> tr(512);
  srand(3);
  for (int round = 0; round < 6; round++) {
    for (size_t i = 0; i < res.size(); i++) {
      std::memset(&res[i], 0, sizeof(eg_result));
      res[i].score = 1.0 + (rand() % 1000) / 1000.0; res[i].net_emissions = (rand() % 3 == 0) ? 5e5 : -10.0; res[i].public_opinion = 0.7;
      res[i].total_cost = 1e10 * (1 + rand() % 50); res[i].power_reliability = 1.0;
      unsigned char* b = (unsigned char*)&tr[i];
      for (size_t k = 0; k < sizeof(eg_traj); k++) b[k] = (unsigned char)(rand() & 0xFF);   // arbitrary counts and action codes
    }
    w->iwi = round * 400;
    eg_update_stats st;
    int rc = eg_update(w, res.data(), tr.data(), (uint32_t)res.size(), round & 1, 77, &st);
    std::printf("round %d rc=%d improvements=%u iwi=%u\n", round, rc, st.n_improvements, st.iterations_without_improvement);
  }
  eg_weights_save_json(w, "fuzz_out.json");
  eg_weights* w2 = nullptr;
  std::printf("reload rc=%d\n", eg_weights_load_json("fuzz_out.json", &w2));
  return 0;
}
