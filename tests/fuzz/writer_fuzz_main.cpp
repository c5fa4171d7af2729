// Fuzz harness of the page writer (tests/test_pages_fuzz_cpu.py): random small graphs, most of them corrupted in one place
// (neighbour ids, upper_base, levels, adj_up entries, entry point), through pgvrx_pages_write under ASan + UBSan with a page
// buffer of exactly the size the sizing call reported.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <random>
#include <string>
#include <vector>
#include "pgvrx_b200.h"
namespace pgvrx_host { void set_create_error(const std::string &) {} }
extern "C" size_t pgvrx_row_bytes(int32_t elem, int32_t dim) { return elem == 0 ? (size_t)dim * 4 : elem == 1 ? (size_t)dim * 2 : ((size_t)dim + 7) / 8; }
int main() {
    std::mt19937_64 rng(5);
    int ok = 0, bad = 0;
    for (int it = 0; it < 4000; it++) {
        int m = 2 + rng() % 20, dims = 1 + rng() % 40, n = rng() % 60;
        std::vector<uint8_t> level(n + 1);
        std::vector<int32_t> ub(n + 1), adj0((size_t)n * 2 * m + 1);
        int64_t up = 0;
        for (int i = 0; i < n; i++) {
            level[i] = (rng() % 5 == 0) ? rng() % 4 : 0;
            ub[i] = level[i] ? (int32_t)up : -1;
            up += level[i];
        }
        std::vector<int32_t> adj_up((size_t)up * m + 1);
        auto fill = [&](std::vector<int32_t> &a, int row_len) {
            for (size_t r = 0; r + row_len <= a.size(); r += row_len) {
                int c = rng() % (row_len + 1);
                for (int j = 0; j < row_len; j++) a[r + j] = j < c ? (int32_t)(rng() % (n ? n : 1)) : -1;
            }
        };
        fill(adj0, 2 * m);
        fill(adj_up, m);
        // corrupt something in most iterations
        int mode = it % 6;
        if (n > 0) {
            if (mode == 1) adj0[rng() % adj0.size()] = (int32_t)(rng() % 1000) - 200;
            if (mode == 2) ub[rng() % n] = (int32_t)(rng() % 1000) - 100;
            if (mode == 3) level[rng() % n] = (uint8_t)(rng() % 256);
            if (mode == 4 && adj_up.size() > 1) adj_up[rng() % adj_up.size()] = (int32_t)(rng() % 1000) - 200;
        }
        int64_t entry = n ? (int64_t)(rng() % n) : -1;
        if (mode == 5) entry = (int64_t)(rng() % 200) - 50;
        std::vector<uint8_t> rows((size_t)n * pgvrx_row_bytes(0, dims) + 1);
        int64_t np = 0;
        int rc = pgvrx_pages_write(0, dims, m, 16, n, rows.data(), level.data(), adj0.data(), ub.data(), up ? adj_up.data() : nullptr, up,
                                   entry, nullptr, nullptr, nullptr, 0, &np);
        if (rc != PGVRX_OK) { bad++; continue; }
        std::vector<uint8_t> pages((size_t)np * 8192);       // exact size: ASan sees any write past the relation
        rc = pgvrx_pages_write(0, dims, m, 16, n, rows.data(), level.data(), adj0.data(), ub.data(), up ? adj_up.data() : nullptr, up,
                               entry, nullptr, nullptr, pages.data(), np, &np);
        if (rc == PGVRX_OK) ok++; else bad++;
    }
    printf("writer fuzz: %d written, %d rejected\n", ok, bad);
    return 0;
}
