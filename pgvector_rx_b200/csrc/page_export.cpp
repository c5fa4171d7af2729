// page_export.cpp -- HNSW index pages -> flat graph arrays for pgvrx_index_load_graph.
//
// Walks the relation exactly as the reference's readers do: meta page at block 0
// (src/types/hnsw.rs:53-74, src/index/scan.rs:126-146), element pages chained from block 1 through
// HnswPageOpaqueData.nextblkno (hnsw.rs:18-27, hnsw_constants.rs:37), element tuples of 72 bytes +
// varlena (hnsw.rs:110-128, scan.rs:155-228) and neighbour tuples holding (level + 2) * m TIDs with
// layer lc at slots [(level - lc) * m, +lm) terminated by the first invalid TID
// (hnsw.rs:150-160, scan.rs:236-283).  Element ids are assigned in page-walk order, which is the
// order build.rs:600-702 wrote them, i.e. insertion order.  Pure host code: one sequential walk of the page chain
// assigns the element ids (a flat (block, offset) -> id table), then the elements are decoded by a few host threads --
// every element writes only its own output rows, and the layer rows of adj_up come from a prefix sum over the levels.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
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

const size_t BLCKSZ = 8192, PAGE_HDR = 24, ITEMID = 4, ETUP_HDR = 72, NTUP_HDR = 4, TID = 6;
const uint32_t INVALID_BLOCK = 0xFFFFFFFFu;

inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

struct Tid { uint32_t blk; uint16_t off; bool valid() const { return blk != INVALID_BLOCK && off != 0; } };
inline Tid rd_tid(const uint8_t *p) { return Tid{((uint32_t)rd16(p) << 16) | rd16(p + 2), rd16(p + 4)}; }

struct Item { const uint8_t *ptr; uint32_t len; bool normal; };

struct Pages {
    const uint8_t *base;
    int64_t n;
    const uint8_t *page(uint32_t b) const { return base + (size_t)b * BLCKSZ; }
    // pd_lower / pd_upper / pd_special of a sane page: 24 <= lower <= upper <= special <= 8192 (bufpage.h); anything else
    // would send the line-pointer walk outside the page
    bool header_ok(uint32_t b) const {
        const uint16_t lower = rd16(page(b) + 12), upper = rd16(page(b) + 14), special = rd16(page(b) + 16);
        return lower >= PAGE_HDR && lower <= upper && upper <= special && special <= BLCKSZ;
    }
    uint16_t max_offset(uint32_t b) const {
        if (!header_ok(b)) return 0;
        uint16_t lower = rd16(page(b) + 12);
        return lower <= PAGE_HDR ? 0 : (uint16_t)((lower - PAGE_HDR) / ITEMID);
    }
    Item item(uint32_t b, uint16_t off) const {
        const uint8_t *p = page(b);
        uint32_t lp = rd32(p + PAGE_HDR + (size_t)(off - 1) * ITEMID);
        uint32_t lp_off = lp & 0x7fff, flags = (lp >> 15) & 3, len = lp >> 17;
        return Item{p + lp_off, len, flags == 1 /* LP_NORMAL */ && len > 0 && lp_off + len <= BLCKSZ};
    }
    uint32_t next(uint32_t b) const { return rd32(page(b) + rd16(page(b) + 16)); }
    bool is_hnsw(uint32_t b) const { return rd16(page(b) + 16) + 8u <= BLCKSZ && rd16(page(b) + rd16(page(b) + 16) + 6) == 0xFF90; }
};

int err(const char *msg) {
    pgvrx_host::set_create_error(msg);
    return PGVRX_ERR_INVALID;
}

struct Walk {
    std::vector<uint32_t> blk;      // per element
    std::vector<uint16_t> off;
    std::vector<int64_t> slot0;     // per block: first entry of that block in id_flat, -1 for a block off the chain
    std::vector<uint16_t> slots;    // per block: max offset + 1 (entries in id_flat)
    std::vector<int32_t> id_flat;   // [slot0[block] + offset] -> element id or -1
    int32_t id_of(uint32_t b, uint16_t o) const {
        if ((size_t)b >= slot0.size() || slot0[b] < 0 || o >= slots[b]) return -1;
        return id_flat[(size_t)slot0[b] + o];
    }
};

int walk_elements(const Pages &pg, Walk &w, int64_t *n_deleted) {
    w.slot0.assign((size_t)pg.n, -1);
    w.slots.assign((size_t)pg.n, 0);
    *n_deleted = 0;
    for (uint32_t b = 1; b != INVALID_BLOCK;) {          // HNSW_HEAD_BLKNO
        if ((int64_t)b >= pg.n) return err("pgvector-rx: page chain points past the end of the relation");
        if (w.slot0[b] >= 0) return err("pgvector-rx: page chain loops");
        if (!pg.is_hnsw(b)) return err("pgvector-rx: hnsw index is not valid");
        if (!pg.header_ok(b)) return err("pgvector-rx: corrupt page header in the element chain");
        uint16_t mo = pg.max_offset(b);
        w.slot0[b] = (int64_t)w.id_flat.size();
        w.slots[b] = (uint16_t)(mo + 1);
        w.id_flat.resize(w.id_flat.size() + (size_t)mo + 1, -1);
        int32_t *ids = w.id_flat.data() + w.slot0[b];
        for (uint16_t o = 1; o <= mo; o++) {
            Item it = pg.item(b, o);
            if (!it.normal || it.len < 4) continue;
            if (it.ptr[0] != 1) continue;                 // HNSW_ELEMENT_TUPLE_TYPE
            if (it.len < ETUP_HDR + 8) return err("pgvector-rx: truncated element tuple");
            if (it.ptr[2] != 0) (*n_deleted)++;
            ids[o] = (int32_t)w.blk.size();
            w.blk.push_back(b);
            w.off.push_back(o);
        }
        b = pg.next(b);
    }
    return PGVRX_OK;
}

// fn(lo, hi) over [0, n) on a few threads; small relations stay on the calling thread
template <class F> void parallel_ranges(int64_t n, F fn) {
    int nt = (int)std::min<int64_t>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())), n / 16384);
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

// meta page + element walk + per-element validation, shared by inspect and export (which needs the walk as well)
int inspect_walk(const Pages &pg, int32_t elem_type, pgvrx_pages_info *info, Walk &w) {
    memset(info, 0, sizeof *info);
    const uint8_t *m = pg.page(0) + PAGE_HDR;
    if (rd32(m) != 0xA953A953u) return err("pgvector-rx: hnsw index is not valid");      // scan.rs:134-137
    if (rd32(m + 4) != 1) return err("pgvector-rx: unsupported hnsw index version");
    info->dims = (int32_t)rd32(m + 8);
    info->m = rd16(m + 12);
    info->ef_construction = rd16(m + 14);
    // reloption bounds (hnsw_constants.rs:53-65; dimensions up to the bit type's 64000): a meta page outside them is corrupt
    if (info->m < 2 || info->m > 100 || info->ef_construction < 4 || info->ef_construction > 1000 || info->dims < 1 || info->dims > 64000)
        return err("pgvector-rx: hnsw index is not valid (meta page parameters out of range)");
    Tid entry{rd32(m + 16), rd16(m + 20)};
    info->entry_level = (int16_t)rd16(m + 22);
    int rc = walk_elements(pg, w, &info->n_deleted);
    if (rc) return rc;
    info->n_elements = (int64_t)w.blk.size();
    info->entry = -1;
    if (entry.blk != INVALID_BLOCK) {
        if (w.id_of(entry.blk, entry.off) < 0) return err("pgvector-rx: entry point does not reference an element tuple");
        info->entry = w.id_of(entry.blk, entry.off);
    }
    const size_t want_row = pgvrx_row_bytes(elem_type, info->dims);
    std::atomic<int64_t> n_upper{0}, n_tids{0};
    std::atomic<int> bad{0};                 // 1 = varlena header, 2 = value size
    parallel_ranges((int64_t)w.blk.size(), [&](int64_t lo, int64_t hi) {
        int64_t up = 0, tids = 0;
        int b = 0;
        for (int64_t i = lo; i < hi && !b; i++) {
            Item it = pg.item(w.blk[(size_t)i], w.off[(size_t)i]);
            up += it.ptr[1];
            for (int t = 0; t < PGVRX_HEAPTIDS; t++) {
                if (!rd_tid(it.ptr + 4 + t * TID).valid()) break;
                tids++;
            }
            if (it.ptr[2] != 0) continue;       // deleted: vacuum zeroed the value, varlena header included (vacuum.rs:750-754)
            uint32_t vl = rd32(it.ptr + ETUP_HDR);
            if ((vl & 3u) != 0) { b = 1; break; }
            size_t vsize = vl >> 2;
            if (vsize != want_row + 8 || ETUP_HDR + vsize > it.len) { b = 2; break; }
        }
        n_upper += up; n_tids += tids;
        if (b) { int zero = 0; bad.compare_exchange_strong(zero, b); }
    });
    if (bad.load() == 1) return err("pgvector-rx: element value is not a plain 4-byte-header varlena");
    if (bad.load() == 2) return err("pgvector-rx: element value size does not match the index dimensions");
    info->n_upper_rows = n_upper.load();
    info->n_heap_tids = n_tids.load();
    return PGVRX_OK;
}

}  // namespace

extern "C" {

static int pgvrx_pages_inspect_impl(const void *pages, int64_t n_pages, int32_t elem_type, pgvrx_pages_info *info) {
    if (!pages || !info || n_pages < 2) return err("pgvector-rx: need the meta page and the head page");
    Pages pg{(const uint8_t *)pages, n_pages};
    Walk w;
    return inspect_walk(pg, elem_type, info, w);
}

static int pgvrx_pages_export_impl(const void *pages, int64_t n_pages, int32_t elem_type, void *rows, uint8_t *level, int32_t *adj0,
                       int32_t *upper_base, int32_t *adj_up, uint64_t *heap_tids, uint8_t *heap_tid_counts) {
    if (!pages || n_pages < 2) return err("pgvector-rx: need the meta page and the head page");
    Pages pg{(const uint8_t *)pages, n_pages};
    pgvrx_pages_info info;
    Walk w;
    int rc = inspect_walk(pg, elem_type, &info, w);
    if (rc) return rc;
    if (info.n_elements > 0 && (!rows || !level || !adj0 || !upper_base || (info.n_upper_rows > 0 && !adj_up)))
        return err("pgvector-rx: null output array");
    const int m = info.m;
    const size_t row_bytes = pgvrx_row_bytes(elem_type, info.dims);
    const int64_t n = info.n_elements;
    // first adj_up row of every element = sum of the levels before it (element order)
    std::vector<int64_t> up_of((size_t)n);
    {
        int64_t up = 0;
        for (int64_t i = 0; i < n; i++) { up_of[(size_t)i] = up; up += pg.item(w.blk[(size_t)i], w.off[(size_t)i]).ptr[1]; }
    }
    std::atomic<int> bad{0};
    parallel_ranges(n, [&](int64_t lo, int64_t hi) {
        for (int64_t ii = lo; ii < hi; ii++) {
            const size_t i = (size_t)ii;
            Item it = pg.item(w.blk[i], w.off[i]);
            const int lv = it.ptr[1];
            const uint8_t version = it.ptr[3];
            const int64_t up = up_of[i];
            level[i] = (uint8_t)lv;
            // a deleted element keeps its id (neighbour TIDs may still point at it) but has no value; its distance is never used
            if (it.ptr[2] != 0) memset((uint8_t *)rows + i * row_bytes, 0, row_bytes);
            else memcpy((uint8_t *)rows + i * row_bytes, it.ptr + ETUP_HDR + 8, row_bytes);
            int nh = 0;
            for (int t = 0; t < PGVRX_HEAPTIDS; t++) {                    // scan.rs:203-209: stop at the first invalid
                Tid h = rd_tid(it.ptr + 4 + t * TID);
                if (!h.valid()) break;
                if (heap_tids) heap_tids[i * PGVRX_HEAPTIDS + nh] = ((uint64_t)h.blk << 16) | h.off;
                nh++;
            }
            if (heap_tids) for (int t = nh; t < PGVRX_HEAPTIDS; t++) heap_tids[i * PGVRX_HEAPTIDS + t] = ~0ull;
            if (heap_tid_counts) heap_tid_counts[i] = (uint8_t)nh;
            upper_base[i] = lv > 0 ? (int32_t)up : -1;
            for (int j = 0; j < 2 * m; j++) adj0[i * 2 * m + j] = -1;
            for (int l = 1; l <= lv; l++) for (int j = 0; j < m; j++) adj_up[(size_t)(up + l - 1) * m + j] = -1;

            Tid nt = rd_tid(it.ptr + 4 + PGVRX_HEAPTIDS * TID);
            bool ok = nt.valid() && (int64_t)nt.blk < n_pages && nt.off <= pg.max_offset(nt.blk);
            Item ni{nullptr, 0, false};
            if (ok) ni = pg.item(nt.blk, nt.off);
            // scan.rs:254-265: wrong type / version / count => the element has no usable lists
            ok = ok && ni.normal && ni.len >= NTUP_HDR && ni.ptr[0] == 2 && ni.ptr[1] == version &&
                 rd16(ni.ptr + 2) == (lv + 2) * m && NTUP_HDR + (size_t)(lv + 2) * m * TID <= ni.len;
            if (!ok) continue;
            for (int lc = lv; lc >= 0; lc--) {
                const int lm = lc == 0 ? 2 * m : m;
                const size_t start = (size_t)(lv - lc) * m;            // scan.rs:268
                int32_t *dst = lc == 0 ? adj0 + i * 2 * m : adj_up + (size_t)(up + lc - 1) * m;
                for (int j = 0; j < lm; j++) {
                    Tid t = rd_tid(ni.ptr + NTUP_HDR + (start + j) * TID);
                    if (!t.valid()) break;                             // scan.rs:275-277
                    const int32_t id = w.id_of(t.blk, t.off);
                    if (id < 0) { bad.store(1); return; }
                    dst[j] = id;
                }
            }
        }
    });
    if (bad.load()) return err("pgvector-rx: neighbour TID does not reference an element tuple");
    return PGVRX_OK;
}

static int pgvrx_pages_element_flags_impl(const void *pages, int64_t n_pages, uint8_t *flags) {
    if (!pages || !flags || n_pages < 2) return err("pgvector-rx: need the meta page and the head page");
    Pages pg{(const uint8_t *)pages, n_pages};
    Walk w;
    int64_t nd;
    int rc = walk_elements(pg, w, &nd);
    if (rc) return rc;
    for (size_t i = 0; i < w.blk.size(); i++) {
        Item it = pg.item(w.blk[i], w.off[i]);
        uint8_t f = 0;
        if (it.ptr[2] != 0) f |= PGVRX_ELEM_DELETED;                              // scan.rs:179
        if (!rd_tid(it.ptr + 4).valid()) f |= PGVRX_ELEM_NO_HEAPTIDS;             // scan.rs:203-209,866-868
        flags[i] = f;
    }
    return PGVRX_OK;
}

int pgvrx_heap_tid_stream(const int32_t *elem_ids, int64_t n_ids, const uint64_t *heap_tids, const uint8_t *heap_tid_counts,
                          int64_t n_elements, uint64_t *out_tids, int64_t cap, int64_t *n_out) {
    if (n_ids < 0 || n_elements < 0 || cap < 0 || !n_out || (n_ids > 0 && (!elem_ids || !heap_tids || !heap_tid_counts)) || (cap > 0 && !out_tids))
        return err("pgvector-rx: null argument");
    int64_t k = 0;
    for (int64_t i = 0; i < n_ids; i++) {
        const int32_t e = elem_ids[i];
        if (e < 0) continue;                                             // padding of a result row
        if (e >= n_elements) return err("pgvector-rx: element id out of range");
        const int c = heap_tid_counts[e];
        if (c > PGVRX_HEAPTIDS) return err("pgvector-rx: an element holds at most 10 heap TIDs");
        for (int t = c - 1; t >= 0; t--, k++)                            // scan.rs:798: pop() takes the last heap TID first
            if (k < cap) out_tids[k] = heap_tids[(size_t)e * PGVRX_HEAPTIDS + t];
    }
    *n_out = k;                                                          // k > cap: call again with a larger buffer
    return PGVRX_OK;
}

// The entry points proper: host memory running out while the tables for a huge (or corrupt) relation are built comes
// back as a status code -- no exception crosses the C ABI.
#define PGVRX_NOTHROW(call)                                                           \
    try { return call; }                                                              \
    catch (const std::bad_alloc &) { pgvrx_host::set_create_error("pgvector-rx: out of host memory"); return PGVRX_ERR_NOMEM; } \
    catch (...) { pgvrx_host::set_create_error("pgvector-rx: unexpected host error"); return PGVRX_ERR_INVALID; }

int pgvrx_pages_inspect(const void *pages, int64_t n_pages, int32_t elem_type, pgvrx_pages_info *info) {
    PGVRX_NOTHROW(pgvrx_pages_inspect_impl(pages, n_pages, elem_type, info))
}
int pgvrx_pages_export(const void *pages, int64_t n_pages, int32_t elem_type, void *rows, uint8_t *level, int32_t *adj0,
                       int32_t *upper_base, int32_t *adj_up, uint64_t *heap_tids, uint8_t *heap_tid_counts) {
    PGVRX_NOTHROW(pgvrx_pages_export_impl(pages, n_pages, elem_type, rows, level, adj0, upper_base, adj_up, heap_tids, heap_tid_counts))
}
int pgvrx_pages_element_flags(const void *pages, int64_t n_pages, uint8_t *flags) {
    PGVRX_NOTHROW(pgvrx_pages_element_flags_impl(pages, n_pages, flags))
}

}  // extern "C"
