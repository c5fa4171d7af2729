// Fuzz harness of the page exporter (tests/test_pages_fuzz_cpu.py): every file named on the command line is read into a
// buffer of exactly its size and pushed through inspect / export / element_flags under ASan + UBSan, so any read outside
// the relation or write outside the output arrays aborts.  Stubs replace the two symbols the exporter takes from the library.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "pgvrx_b200.h"
namespace pgvrx_host { void set_create_error(const std::string &) {} }
extern "C" size_t pgvrx_row_bytes(int32_t elem, int32_t dim) { return elem == 0 ? (size_t)dim * 4 : elem == 1 ? (size_t)dim * 2 : ((size_t)dim + 7) / 8; }
int main(int argc, char **argv) {
    int ok = 0, bad = 0;
    for (int a = 1; a < argc; a++) {
        FILE *f = fopen(argv[a], "rb");
        fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<unsigned char> buf((size_t)sz);   // exact size: ASan sees any read past the relation
        if (fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) return 2;
        fclose(f);
        int64_t np = sz / 8192;
        pgvrx_pages_info info;
        if (pgvrx_pages_inspect(buf.data(), np, PGVRX_ELEM_F32, &info) != PGVRX_OK) { bad++; continue; }
        size_t n = (size_t)info.n_elements, m = (size_t)info.m;
        std::vector<unsigned char> rows(n * pgvrx_row_bytes(0, info.dims) + 1), level(n + 1), cnt(n + 1), flags(n + 1);
        std::vector<int32_t> adj0(n * 2 * m + 1), ub(n + 1), adj_up((size_t)info.n_upper_rows * m + 1);
        std::vector<uint64_t> tids(n * 10 + 1);
        int rc = pgvrx_pages_export(buf.data(), np, PGVRX_ELEM_F32, rows.data(), level.data(), adj0.data(), ub.data(), adj_up.data(), tids.data(), cnt.data());
        int rc2 = pgvrx_pages_element_flags(buf.data(), np, flags.data());
        if (rc == PGVRX_OK && rc2 == PGVRX_OK) ok++; else bad++;
    }
    printf("asan fuzz: %d accepted, %d rejected\n", ok, bad);
    return 0;
}
