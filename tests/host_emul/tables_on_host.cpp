// Test harness (not a CPU path): the host-side table computation of csrc/tables.cpp (fill_full_table, round 2) linked WITHOUT a device,
// printing chosen entries as plain integers for tests/test_tables_on_host.py to check against big-integer arithmetic.
//   tables_on_host <kind> <lg_m> <lg_rows> <index>...      kind: 0 inverse twiddle, 1 forward twiddle, 2 zk_shift
#include "internal.h"
#include "constants.inc"
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int kind = atoi(argv[1]);
    const uint32_t lg_m = (uint32_t)atoi(argv[2]), lg_rows = (uint32_t)atoi(argv[3]);
    std::vector<uint32_t> v;
    b200::fill_full_table(kind, lg_m, lg_rows, B200_ROU_FWD_MONT, B200_ROU_REV_MONT, v);
    printf("%zu", v.size());
    for (int i = 4; i < argc; i++) {
        const size_t idx = (size_t)strtoull(argv[i], nullptr, 10);
        printf(" %u", idx < v.size() ? b200::h_from_mont(v[idx]) : 0xffffffffu);
    }
    printf("\n");
    return 0;
}
