// ASan/UBSan harness (tests/test_host_sanitizers.py): eg_update_combine_apply / fill_policy / merge / history / save+load over
// arbitrary statistics tables and winner records.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "weights.hpp"
int eg_fail(int code, const std::string&) { return code; }
int main() {
  eg_weights* w = nullptr;
  eg_weights_new(&w);
  srand(11);
  std::vector<int64_t> stats(EG_STATS_WORDS);
  std::vector<unsigned char> rec(EG_BEST_RECORD_BYTES * 8);
  for (int round = 0; round < 12; round++) {
    for (auto& s : stats) s = (rand() % 7 == 0) ? -(int64_t)(rand() % 100000) * 1000 : (rand() % 5 == 0 ? rand() % 50 : 0);
    stats[0] = 65536; stats[1] = rand() % 65536; stats[2] = rand() % 100;
    for (auto& b : rec) b = (unsigned char)(rand() & 0xFF);   // arbitrary winner records (scores, ids, results, trajectories)
    w->iwi = (round % 4) * 700;
    eg_update_stats st;
    int rc = eg_update_combine_apply(w, stats.data(), rec.data(), 8, 8ull * 65536, (uint64_t)round * 524288, &st);
    std::printf("round %d rc=%d iwi=%u best=%g improvements=%u\n", round, rc, st.iterations_without_improvement, st.best_score, st.n_improvements);
    EgPolicyDevice pol;
    eg_weights_fill_policy(*w, &pol);
  }
  eg_weights* c = nullptr;
  eg_weights_clone(w, &c);
  eg_weights_merge(c, w);
  eg_weights_history_append(w, 123, "fuzz_hist.json");
  eg_weights_history_append(c, 456, "fuzz_hist.json");
  eg_weights_save_json(c, "fuzz_out2.json");
  eg_weights* r = nullptr;
  std::printf("reload rc=%d\n", eg_weights_load_json("fuzz_out2.json", &r));
  return 0;
}
