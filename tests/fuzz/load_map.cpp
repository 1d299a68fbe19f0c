// ASan/UBSan harness (tests/test_host_sanitizers.py): eg_host_map_load + validate on the three asset files given as arguments.
#include <cstdio>
#include <string>
#include "host_tables.hpp"
#include "common.hpp"
static std::string last;
int eg_fail(int code, const std::string& m) { last = m; return code; }
int main(int argc, char** argv) {
  EgHostMap m;
  int rc = eg_host_map_load(&m, argv[1], argv[2], argv[3]);
  if (rc == 0) rc = eg_host_map_validate(m);
  std::printf("rc=%d %s | S=%zu E=%zu C=%zu\n", rc, rc ? last.c_str() : "", m.sx.size(), m.ex.size(), m.cx.size());
  return 0;
}
