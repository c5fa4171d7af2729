// pgvrx_relation_check.cpp -- point the page exporter at an HNSW index relation file and report what it finds.
//
// SURVEY 8(f) rank 1 asks for the exporter to be validated against index files written by the real extension on a
// machine that has PostgreSQL; none exists here, so this is the tool a maintainer runs there:
//
//   psql -Atc "SELECT pg_relation_filepath('items_embedding_idx')"        ->  base/16384/24601
//   pgvrx_relation_check $PGDATA/base/16384/24601 --elem f32 [--scan 1000]
//
// Host-only by default: pgvrx_pages_inspect / _export / _element_flags (the walk the reference's readers do,
// src/index/scan.rs:126-283, src/types/hnsw.rs:18-169), then structural checks of what came out -- every neighbour id
// is an element, lists are dense prefixes, no self-links, levels within hnsw_get_max_level(m), the entry point is a
// highest-level live element, heap TID counts -- and round-trips the arrays through pgvrx_pages_write to see whether
// the relation is byte-identical to what the reference's flush phase (build.rs:545-821) would write for that graph
// (true right after CREATE INDEX; inserts and vacuums rewrite tuples in place and break it, which is reported, not an
// error).  With --scan N (needs a GPU) the graph is loaded into HBM and N elements are looked up by their own rows:
// for the L2 / Hamming opclasses a live element must come back at distance 0.
// Relations above 1 GiB are split into segment files (file, file.1, file.2, ...): they are read in order.
// C++17 over include/pgvrx_b200.h only.  Exit status: 0 ok, 1 a check failed, 2 usage / IO, 3 library error.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pgvrx_b200.h"

namespace {

bool read_relation(const std::string &path, std::vector<uint8_t> &out) {
    for (int seg = 0;; seg++) {
        const std::string p = seg == 0 ? path : path + "." + std::to_string(seg);
        FILE *f = std::fopen(p.c_str(), "rb");
        if (!f) return seg > 0;
        uint8_t buf[1 << 16];
        size_t k;
        while ((k = std::fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + k);
        std::fclose(f);
    }
}

int lib_error(const char *what, const pgvrx_index *ix) {
    std::fprintf(stderr, "pgvrx_relation_check: %s: %s\n", what, pgvrx_last_error(ix));
    return 3;
}

}  // namespace

int main(int argc, char **argv) {
    std::string path;
    int elem = PGVRX_ELEM_F32, metric = -1;
    int64_t n_scan = 0;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--elem" && i + 1 < argc) {
            const std::string v = argv[++i];
            elem = v == "f32" ? PGVRX_ELEM_F32 : v == "f16" ? PGVRX_ELEM_F16 : v == "bit" ? PGVRX_ELEM_BIT : -1;
        } else if (a == "--metric" && i + 1 < argc) {
            const std::string v = argv[++i];
            metric = v == "l2" ? PGVRX_METRIC_L2SQ : v == "ip" ? PGVRX_METRIC_NEG_IP : v == "l1" ? PGVRX_METRIC_L1
                   : v == "hamming" ? PGVRX_METRIC_HAMMING : v == "jaccard" ? PGVRX_METRIC_JACCARD : -2;
        } else if (a == "--scan" && i + 1 < argc) n_scan = std::strtoll(argv[++i], nullptr, 10);
        else if (path.empty() && a[0] != '-') path = a;
        else elem = -1;
    }
    if (path.empty() || elem < 0 || metric == -2) {
        std::fprintf(stderr, "usage: pgvrx_relation_check <relation file> [--elem f32|f16|bit] [--metric l2|ip|l1|hamming|jaccard] [--scan N]\n");
        return 2;
    }
    if (metric < 0) metric = elem == PGVRX_ELEM_BIT ? PGVRX_METRIC_HAMMING : PGVRX_METRIC_L2SQ;
    std::vector<uint8_t> pages;
    if (!read_relation(path, pages) || pages.size() % 8192 != 0 || pages.empty()) {
        std::fprintf(stderr, "pgvrx_relation_check: cannot read %s as a whole number of 8 KiB pages\n", path.c_str());
        return 2;
    }
    const int64_t n_pages = (int64_t)(pages.size() / 8192);
    pgvrx_pages_info info;
    if (pgvrx_pages_inspect(pages.data(), n_pages, elem, &info)) return lib_error("pgvrx_pages_inspect", nullptr);
    const int64_t n = info.n_elements, nu = info.n_upper_rows;
    const int m = info.m;
    const size_t rb = pgvrx_row_bytes(elem, info.dims);
    std::vector<uint8_t> rows((size_t)std::max<int64_t>(n, 1) * rb), level((size_t)std::max<int64_t>(n, 1)), cnt((size_t)std::max<int64_t>(n, 1)), flags((size_t)std::max<int64_t>(n, 1));
    std::vector<int32_t> adj0((size_t)std::max<int64_t>(n, 1) * 2 * m), ub((size_t)std::max<int64_t>(n, 1)), adj_up((size_t)std::max<int64_t>(nu, 1) * m);
    std::vector<uint64_t> tids((size_t)std::max<int64_t>(n, 1) * PGVRX_HEAPTIDS);
    if (pgvrx_pages_export(pages.data(), n_pages, elem, rows.data(), level.data(), adj0.data(), ub.data(), adj_up.data(), tids.data(), cnt.data()))
        return lib_error("pgvrx_pages_export", nullptr);
    if (pgvrx_pages_element_flags(pages.data(), n_pages, flags.data())) return lib_error("pgvrx_pages_element_flags", nullptr);

    // ---- structural checks ----
    int64_t bad_prefix = 0, self_links = 0, to_deleted = 0, edges0 = 0, max_level = 0, no_tids = 0, empty0 = 0;
    const int level_cap = pgvrx_max_level(m);
    int64_t over_cap = 0;
    auto check_list = [&](const int32_t *l, int lm, int64_t self) {
        bool ended = false;
        int c = 0;
        for (int j = 0; j < lm; j++) {
            if (l[j] < 0) { ended = true; continue; }
            if (ended) bad_prefix++;
            if (l[j] == self) self_links++;
            if (flags[(size_t)l[j]] & PGVRX_ELEM_DELETED) to_deleted++;
            c++;
        }
        return c;
    };
    for (int64_t i = 0; i < n; i++) {
        max_level = std::max<int64_t>(max_level, level[(size_t)i]);
        over_cap += level[(size_t)i] > level_cap;
        no_tids += (flags[(size_t)i] & PGVRX_ELEM_NO_HEAPTIDS) != 0;
        if (flags[(size_t)i] & PGVRX_ELEM_DELETED) continue;
        const int c0 = check_list(&adj0[(size_t)i * 2 * m], 2 * m, i);
        edges0 += c0;
        empty0 += c0 == 0;
        for (int l = 1; l <= level[(size_t)i]; l++) check_list(&adj_up[((size_t)ub[(size_t)i] + l - 1) * m], m, i);
    }
    bool entry_ok = n == 0 ? info.entry == -1 : info.entry >= 0 && level[(size_t)info.entry] == info.entry_level;
    // byte round trip through the reference's flush layout
    int64_t np2 = 0;
    bool same_bytes = false;
    if (pgvrx_pages_write(elem, info.dims, m, info.ef_construction, n, rows.data(), level.data(), adj0.data(), ub.data(), adj_up.data(), nu,
                          info.entry, tids.data(), cnt.data(), nullptr, 0, &np2) == PGVRX_OK && np2 == n_pages) {
        std::vector<uint8_t> again((size_t)np2 * 8192);
        if (pgvrx_pages_write(elem, info.dims, m, info.ef_construction, n, rows.data(), level.data(), adj0.data(), ub.data(), adj_up.data(), nu,
                              info.entry, tids.data(), cnt.data(), again.data(), np2, &np2) == PGVRX_OK)
            same_bytes = again == pages;
    }
    const bool ok = bad_prefix == 0 && self_links == 0 && over_cap == 0 && entry_ok;

    // ---- optional: load into HBM and look elements up by their own rows ----
    int64_t scanned = 0, found_self = 0;
    if (n_scan > 0 && n > 0) {
        pgvrx_index_desc d;
        std::memset(&d, 0, sizeof d);
        d.elem_type = elem; d.metric = metric; d.dim = info.dims; d.m = m; d.ef_construction = info.ef_construction; d.capacity = n;
        pgvrx_index *ix = nullptr;
        if (pgvrx_index_create(&d, &ix)) return lib_error("pgvrx_index_create", nullptr);
        if (pgvrx_index_load_graph(ix, rows.data(), n, level.data(), adj0.data(), ub.data(), adj_up.data(), nu, info.entry)) return lib_error("pgvrx_index_load_graph", ix);
        if (info.n_deleted > 0 || no_tids > 0)
            if (pgvrx_index_set_element_flags(ix, flags.data(), n)) return lib_error("pgvrx_index_set_element_flags", ix);
        std::vector<int64_t> pick;
        for (int64_t i = 0; i < n && (int64_t)pick.size() < n_scan; i += std::max<int64_t>(1, n / n_scan))
            if (flags[(size_t)i] == 0) pick.push_back(i);
        std::vector<uint8_t> q(pick.size() * rb);
        for (size_t j = 0; j < pick.size(); j++) std::memcpy(&q[j * rb], &rows[(size_t)pick[j] * rb], rb);
        std::vector<int32_t> ids(pick.size()), c(pick.size());
        std::vector<float> ds(pick.size());
        if (pgvrx_search_batch(ix, q.data(), (int64_t)pick.size(), PGVRX_DEFAULT_EF_SEARCH, 1, ids.data(), ds.data(), c.data(), nullptr))
            return lib_error("pgvrx_search_batch", ix);
        scanned = (int64_t)pick.size();
        for (size_t j = 0; j < pick.size(); j++) found_self += c[j] == 1 && (ids[j] == pick[j] || ds[j] == 0.0f);
        pgvrx_index_destroy(ix);
    }

    std::printf("{\"relation\": \"%s\", \"pages\": %lld, \"elements\": %lld, \"deleted\": %lld, \"without_heap_tids\": %lld, \"heap_tids\": %lld, "
                "\"dims\": %d, \"m\": %d, \"ef_construction\": %d, \"entry\": %lld, \"entry_level\": %d, \"max_level\": %lld, \"upper_rows\": %lld, "
                "\"mean_degree_layer0\": %.2f, \"elements_without_neighbours\": %lld, \"links_to_deleted\": %lld, "
                "\"lists_not_dense_prefix\": %lld, \"self_links\": %lld, \"levels_over_cap\": %lld, \"entry_ok\": %s, "
                "\"identical_to_flush_layout\": %s, \"scanned\": %lld, \"found_self\": %lld, \"ok\": %s}\n",
                path.c_str(), (long long)n_pages, (long long)n, (long long)info.n_deleted, (long long)no_tids, (long long)info.n_heap_tids,
                info.dims, m, info.ef_construction, (long long)info.entry, info.entry_level, (long long)max_level, (long long)nu,
                n - info.n_deleted > 0 ? (double)edges0 / (double)(n - info.n_deleted) : 0.0, (long long)empty0, (long long)to_deleted,
                (long long)bad_prefix, (long long)self_links, (long long)over_cap, entry_ok ? "true" : "false",
                same_bytes ? "true" : "false", (long long)scanned, (long long)found_self, ok ? "true" : "false");
    return ok ? 0 : 1;
}
