// page_write.cpp -- flat graph arrays -> HNSW index pages: the flush phase of ambuild
// (src/index/build.rs:545-821) for a graph that was built on the GPU, byte for byte what the reference writes from
// its in-memory HnswBuildState, so that pgvrx_index_build + pgvrx_index_export_graph + this = CREATE INDEX.
//
//   create_meta_page      build.rs:545-569   block 0: PageInit + HnswMetaPageData at page + 24, pd_lower past it
//   create_graph_pages    build.rs:576-702   element tuple (72 B header + varlena) then its neighbour tuple, in element
//                                            order; a new page when the element does not fit, or when the pair would fit
//                                            on an empty page (combined <= HNSW_MAX_SIZE) but not here; the neighbour
//                                            tuple moves to the next page on its own when only it does not fit
//   write_neighbor_tuples build.rs:719-790   (level + 2) * m TIDs, layers from the top down, invalid TID after the last
//   update_meta_page      build.rs:797-821   entry point + insert page (= the last page written)
//
// Two passes: placement is computed from tuple sizes alone (PageAddItemExtended / PageGetFreeSpace arithmetic) in one
// sequential sweep, then the pages are filled -- by a few host threads for large graphs: every tuple has its own
// bytes and its own line pointer, and the page headers follow from the placement.  Pure host code.  Layouts: src/types/hnsw.rs:18-169; the exporter (page_export.cpp) is the
// inverse and the tests check both directions against the oracle's independent writer.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pgvrx_b200.h"

namespace pgvrx_host {
void set_create_error(const std::string &msg);   // capi.cu
}

namespace {

const size_t BLCKSZ = 8192, PAGE_HDR = 24, OPAQUE = 8, ITEMID = 4, ETUP_HDR = 72, NTUP_HDR = 4, TID = 6, VARHDR = 8;
const size_t SPECIAL = BLCKSZ - OPAQUE;                                  // PageInit: pd_special = pagesize - MAXALIGN(8)
const size_t MAX_SIZE = BLCKSZ - PAGE_HDR - OPAQUE - ITEMID;             // hnsw_max_size (types/hnsw.rs:325-331) = 8156

inline size_t maxalign(size_t x) { return (x + 7) & ~(size_t)7; }
inline void wr16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }
inline void wr32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }
inline void wr_tid(uint8_t *p, uint32_t blk, uint16_t off) { wr16(p, (uint16_t)(blk >> 16)); wr16(p + 2, (uint16_t)(blk & 0xffffu)); wr16(p + 4, off); }
inline void wr_invalid_tid(uint8_t *p) { wr_tid(p, 0xFFFFFFFFu, 0); }

int fail(const char *msg) {
    pgvrx_host::set_create_error(msg);
    return PGVRX_ERR_INVALID;
}

// where a tuple lands: block, line pointer number, byte offset of the tuple inside the page
struct Slot { uint32_t blk; uint16_t off; uint16_t at; };
struct Placement { Slot elem, nbr; };

// running state of the page being filled (what PageAddItemExtended / PageGetFreeSpace look at)
struct Cursor {
    uint32_t blk;
    uint16_t lower, upper, n_items;
    void fresh(uint32_t b) { blk = b; lower = (uint16_t)PAGE_HDR; upper = (uint16_t)SPECIAL; n_items = 0; }
    size_t free_space() const {                                          // PageGetFreeSpace
        const int space = (int)upper - (int)lower;
        return space < (int)ITEMID ? 0 : (size_t)space - ITEMID;
    }
    Slot add(size_t size) {                                              // PageAddItemExtended, offset = max + 1
        upper = (uint16_t)(upper - maxalign(size));
        lower = (uint16_t)(lower + ITEMID);
        n_items++;
        return Slot{blk, n_items, upper};
    }
};

void init_page(uint8_t *p) {                                             // PageInit + hnsw_init_page (build.rs:56-64)
    memset(p, 0, BLCKSZ);
    wr16(p + 12, (uint16_t)PAGE_HDR);                                    // pd_lower
    wr16(p + 14, (uint16_t)SPECIAL);                                     // pd_upper
    wr16(p + 16, (uint16_t)SPECIAL);                                     // pd_special
    wr16(p + 18, (uint16_t)(BLCKSZ | 4));                                // pd_pagesize_version (PG_PAGE_LAYOUT_VERSION 4)
    wr32(p + SPECIAL, 0xFFFFFFFFu);                                      // nextblkno = InvalidBlockNumber
    wr16(p + SPECIAL + 6, 0xFF90);                                       // page_id = HNSW_PAGE_ID
}

void put_item(uint8_t *page, const Slot &s, const uint8_t *tuple, size_t size) {
    memcpy(page + s.at, tuple, size);
    const uint32_t lp = (uint32_t)s.at | (1u << 15) /* LP_NORMAL */ | ((uint32_t)size << 17);
    wr32(page + PAGE_HDR + (size_t)(s.off - 1) * ITEMID, lp);
}

// fn(lo, hi) over [0, n) on a few threads; small graphs stay on the calling thread
template <class F> void parallel_ranges(int64_t n, int64_t grain, F fn) {
    int nt = (int)std::min<int64_t>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())), n / grain);
    if (nt <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> ts;
    const int64_t per = (n + nt - 1) / nt;
    int started = 0;
    try {                                                 // no exception may cross the C ABI: a thread that cannot be
        ts.reserve((size_t)nt);                           // created (EAGAIN) leaves its range to the calling thread
        for (; started < nt; started++) ts.emplace_back([=]() { fn(std::min(n, started * per), std::min(n, (started + 1) * per)); });
    } catch (...) {
    }
    if (started < nt) fn(std::min(n, started * per), n);
    for (auto &t : ts) t.join();
}

}  // namespace

static int pages_write_impl(int32_t elem_type, int32_t dims, int32_t m, int32_t ef_construction, int64_t n, const void *rows,
                                 const uint8_t *level, const int32_t *adj0, const int32_t *upper_base, const int32_t *adj_up,
                                 int64_t n_upper_rows, int64_t entry, const uint64_t *heap_tids, const uint8_t *heap_tid_counts, void *pages,
                                 int64_t cap_pages, int64_t *n_pages_out) {
    if (elem_type < PGVRX_ELEM_F32 || elem_type > PGVRX_ELEM_BIT) return fail("pgvector-rx: unknown element type");
    if (m < 2 || m > 100 || ef_construction < 4 || ef_construction > 1000 || dims < 1 || dims > 64000)
        return fail("pgvector-rx: index parameters out of range");
    if (n < 0 || (n > 0 && (!rows || !level || !adj0 || !upper_base))) return fail("pgvector-rx: null graph array");
    if (n > 0 && (entry < 0 || entry >= n)) return fail("pgvector-rx: entry point out of range");
    if ((heap_tids == nullptr) != (heap_tid_counts == nullptr)) return fail("pgvector-rx: heap_tids and heap_tid_counts go together");
    const size_t row_bytes = pgvrx_row_bytes(elem_type, dims);
    const size_t value_size = VARHDR + row_bytes;                        // varlena header + type header + data
    const size_t etup_size = maxalign(ETUP_HDR + value_size);            // hnsw_element_tuple_size (hnsw.rs:134-137)
    if (etup_size > MAX_SIZE) return fail("pgvector-rx: index tuple too large");   // build.rs:612-614

    // ---- pass 1: placement ----
    std::vector<Placement> loc((size_t)n);
    Cursor cur;
    cur.fresh(1);                                                        // block 0 is the meta page; the head page always exists
    int64_t up_rows = 0;                                                 // rows of adj_up consumed so far
    for (int64_t i = 0; i < n; i++) {
        // upper lists are packed in element order (pgvrx_index_export_graph, pgvrx_pages_export): anything else would
        // index adj_up out of bounds
        if (level[i] > 0 && (upper_base[i] != up_rows || !adj_up)) return fail("pgvector-rx: upper_base is not the running sum of levels");
        up_rows += level[i];
        const size_t ntup_size = maxalign(NTUP_HDR + (size_t)(level[i] + 2) * (size_t)m * TID);   // hnsw.rs:165-169
        if (ntup_size > MAX_SIZE) return fail("pgvector-rx: level too high for one neighbour tuple");
        if ((level[i] > 0) != (upper_base[i] >= 0)) return fail("pgvector-rx: upper_base inconsistent with level");
        const size_t combined = etup_size + ntup_size + ITEMID;
        const size_t fs = cur.free_space();
        if (fs < etup_size || (combined <= MAX_SIZE && fs < combined)) cur.fresh(cur.blk + 1);   // build.rs:636-642
        // the reference fixes the neighbour tuple's address BEFORE it is placed (build.rs:649-657); check it holds
        const uint32_t want_blk = combined <= MAX_SIZE ? cur.blk : cur.blk + 1;
        const uint16_t want_off = combined <= MAX_SIZE ? (uint16_t)(cur.n_items + 2) : (uint16_t)1;
        loc[(size_t)i].elem = cur.add(etup_size);
        if (cur.free_space() < ntup_size) cur.fresh(cur.blk + 1);        // build.rs:684-687
        loc[(size_t)i].nbr = cur.add(ntup_size);
        if (loc[(size_t)i].nbr.blk != want_blk || loc[(size_t)i].nbr.off != want_off)
            return fail("pgvector-rx: failed to add neighbor tuple to page");              // build.rs:699-701
    }
    if (up_rows != n_upper_rows) return fail("pgvector-rx: adj_up does not hold one row per upper layer of every element");
    const int64_t n_pages = (int64_t)cur.blk + 1;
    if (n_pages_out) *n_pages_out = n_pages;
    if (!pages) return PGVRX_OK;                                         // sizing call
    if (cap_pages < n_pages) return fail("pgvector-rx: page buffer too small");

    // ---- pass 2: fill ----
    uint8_t *base = (uint8_t *)pages;
    parallel_ranges(n_pages, 1024, [&](int64_t lo, int64_t hi) {
        for (int64_t b = lo; b < hi; b++) {
            init_page(base + (size_t)b * BLCKSZ);
            if (b >= 1 && b + 1 < n_pages) wr32(base + (size_t)b * BLCKSZ + SPECIAL, (uint32_t)(b + 1));   // hnsw_build_append_page
        }
    });
    {   // create_meta_page + update_meta_page
        uint8_t *mp = base + PAGE_HDR;
        wr32(mp, 0xA953A953u);                                           // HNSW_MAGIC_NUMBER
        wr32(mp + 4, 1);                                                 // HNSW_VERSION
        wr32(mp + 8, (uint32_t)dims);
        wr16(mp + 12, (uint16_t)m);
        wr16(mp + 14, (uint16_t)ef_construction);
        if (n > 0) {
            wr32(mp + 16, loc[(size_t)entry].elem.blk);
            wr16(mp + 20, loc[(size_t)entry].elem.off);
            wr16(mp + 22, (uint16_t)(int16_t)level[entry]);
        } else {
            wr32(mp + 16, 0xFFFFFFFFu);
            wr16(mp + 20, 0);
            wr16(mp + 22, (uint16_t)(int16_t)-1);
        }
        wr32(mp + 24, (uint32_t)cur.blk);                                // insert_page = last page written
        wr16(base + 12, (uint16_t)(PAGE_HDR + 28));                      // pd_lower past HnswMetaPageData
    }
    // page headers follow the last item placed on each page (placement order = element order, element before neighbour)
    for (int64_t i = 0; i < n; i++)
        for (const Slot *sl : {&loc[(size_t)i].elem, &loc[(size_t)i].nbr}) {
            wr16(base + (size_t)sl->blk * BLCKSZ + 12, (uint16_t)(PAGE_HDR + (size_t)sl->off * ITEMID));
            wr16(base + (size_t)sl->blk * BLCKSZ + 14, sl->at);
        }
    std::atomic<int> bad{0};                                             // 1 = heap TID count, 2 = neighbour id, 3 = heap TID form, 4 = no TIDs
    parallel_ranges(n, 16384, [&](int64_t lo, int64_t hi) {
        std::vector<uint8_t> etup(etup_size), ntup;
        for (int64_t i = lo; i < hi; i++) {
            const Placement &L = loc[(size_t)i];
            const int lv = level[i];
            // element tuple (hnsw.rs:110-128)
            memset(etup.data(), 0, etup_size);
            etup[0] = 1;                                                     // HNSW_ELEMENT_TUPLE_TYPE
            etup[1] = (uint8_t)lv;
            const int n_tids = heap_tids ? heap_tid_counts[i] : 1;
            if (n_tids < 0 || n_tids > PGVRX_HEAPTIDS) { bad.store(1); return; }
            // an element without heap TIDs is what a VACUUM leaves behind (deleted = 1, lists and value cleared, versions
            // bumped: vacuum.rs:655-790).  This writer lays out the CREATE INDEX flush, where every element has 1..10 TIDs;
            // writing such an element as a live zero-TID tuple would resurrect it in every later scan and repair
            if (n_tids == 0) { bad.store(4); return; }
            for (int t = 0; t < PGVRX_HEAPTIDS; t++) {
                uint8_t *slot = etup.data() + 4 + (size_t)t * TID;
                if (t < n_tids) {
                    // caller's TIDs are (block << 16 | offset); without them element i gets the synthetic TID of row i:
                    // block i / 65535, offset i % 65535 + 1 -- the offset is never 0, which would read as "no heap TID"
                    const uint64_t v = heap_tids ? heap_tids[(size_t)i * PGVRX_HEAPTIDS + t] : ((((uint64_t)i / 0xffffu) << 16) | ((uint64_t)i % 0xffffu + 1));
                    // offset 0 is InvalidOffsetNumber: readers stop at it (scan.rs:203-209) and the element would lose its rows
                    if ((v & 0xffffu) == 0 || (v >> 16) >= 0xFFFFFFFFull) { bad.store(3); return; }
                    wr_tid(slot, (uint32_t)(v >> 16), (uint16_t)(v & 0xffffu));
                } else {
                    wr_invalid_tid(slot);
                }
            }
            wr_tid(etup.data() + 4 + PGVRX_HEAPTIDS * TID, L.nbr.blk, L.nbr.off);        // neighbortid
            uint8_t *val = etup.data() + ETUP_HDR;
            wr32(val, (uint32_t)(value_size << 2));                          // 4-byte varlena header (SET_VARSIZE)
            if (elem_type == PGVRX_ELEM_BIT) wr32(val + 4, (uint32_t)dims);  // VarBitHeader.bit_len (bitvec.rs:28-32)
            else { wr16(val + 4, (uint16_t)dims); wr16(val + 6, 0); }        // {dim, unused} (vector.rs:43-48, halfvec.rs:41-46)
            memcpy(val + VARHDR, (const uint8_t *)rows + (size_t)i * row_bytes, row_bytes);
            put_item(base + (size_t)L.elem.blk * BLCKSZ, L.elem, etup.data(), etup_size);
            // neighbour tuple (hnsw.rs:150-160; build.rs:719-790)
            const size_t n_slots = (size_t)(lv + 2) * (size_t)m;
            const size_t ntup_size = maxalign(NTUP_HDR + n_slots * TID);
            ntup.assign(ntup_size, 0);
            ntup[0] = 2;                                                     // HNSW_NEIGHBOR_TUPLE_TYPE
            wr16(ntup.data() + 2, (uint16_t)n_slots);
            size_t t = 0;
            for (int lc = lv; lc >= 0; lc--) {
                const int lm = lc == 0 ? 2 * m : m;
                const int32_t *list = lc == 0 ? adj0 + (size_t)i * 2 * m : adj_up + ((size_t)upper_base[i] + (size_t)(lc - 1)) * (size_t)m;
                bool ended = false;
                for (int j = 0; j < lm; j++, t++) {
                    uint8_t *slot = ntup.data() + NTUP_HDR + t * TID;
                    const int32_t nb = ended ? -1 : list[j];
                    if (nb < 0) { ended = true; wr_invalid_tid(slot); continue; }
                    if (nb >= n) { bad.store(2); return; }
                    wr_tid(slot, loc[(size_t)nb].elem.blk, loc[(size_t)nb].elem.off);
                }
            }
            put_item(base + (size_t)L.nbr.blk * BLCKSZ, L.nbr, ntup.data(), ntup_size);
        }
    });
    if (bad.load() == 1) return fail("pgvector-rx: an element holds at most 10 heap TIDs");
    if (bad.load() == 2) return fail("pgvector-rx: neighbour id out of range");
    if (bad.load() == 3) return fail("pgvector-rx: heap TID is not (block << 16 | offset) with offset >= 1");
    if (bad.load() == 4) return fail("pgvector-rx: element without heap TIDs (a vacuumed index): pgvrx_pages_write covers the CREATE INDEX flush only");
    return PGVRX_OK;
}

// no exception crosses the C ABI: the placement table of a huge graph not fitting in host memory is a status code
extern "C" int pgvrx_pages_write(int32_t elem_type, int32_t dims, int32_t m, int32_t ef_construction, int64_t n, const void *rows,
                                 const uint8_t *level, const int32_t *adj0, const int32_t *upper_base, const int32_t *adj_up,
                                 int64_t n_upper_rows, int64_t entry, const uint64_t *heap_tids, const uint8_t *heap_tid_counts, void *pages,
                                 int64_t cap_pages, int64_t *n_pages_out) {
    try {
        return pages_write_impl(elem_type, dims, m, ef_construction, n, rows, level, adj0, upper_base, adj_up, n_upper_rows, entry,
                                heap_tids, heap_tid_counts, pages, cap_pages, n_pages_out);
    } catch (const std::bad_alloc &) {
        return fail("pgvector-rx: out of host memory"), PGVRX_ERR_NOMEM;
    } catch (...) {
        return fail("pgvector-rx: unexpected host error");
    }
}
