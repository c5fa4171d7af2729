/*
 * hnsw_oracle.cpp -- CPU restatement of pgvector-rx's HNSW hot path (see hnsw_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY.  Must be compiled with -ffp-contract=off and without
 * -ffast-math: the reference accumulates f32 strictly left to right with separate
 * multiply and add (src/types/vector.rs:518-567) and every result below has to be
 * bit-identical to that.
 *
 * Citations are file:line under /root/reference.
 */
#include "hnsw_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

/* ======================================================================================
 * a14: half codec  (src/types/halfvec.rs:54-143)
 * ==================================================================================== */
inline float bits_to_f32(uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; }
inline uint32_t f32_to_bits(float f) { uint32_t b; std::memcpy(&b, &f, 4); return b; }

float half_to_f32(uint16_t h) {
    uint32_t sign = (h >> 15) & 1u, exp = (h >> 10) & 0x1fu, mant = h & 0x3ffu;
    if (exp == 0) {
        if (mant == 0) return bits_to_f32(sign << 31);
        /* subnormal: renormalise (halfvec.rs:65-74) */
        uint32_t m = mant;
        int e = -14;
        while ((m & 0x400u) == 0) { m <<= 1; e -= 1; }
        m &= 0x3ffu;
        uint32_t fe = (uint32_t)(e + 127);
        return bits_to_f32((sign << 31) | (fe << 23) | (m << 13));
    }
    if (exp == 31) {
        if (mant == 0) return bits_to_f32((sign << 31) | (0xffu << 23));
        return bits_to_f32((sign << 31) | (0xffu << 23) | (mant << 13));
    }
    uint32_t fe = (uint32_t)((int)exp - 15 + 127);
    return bits_to_f32((sign << 31) | (fe << 23) | (mant << 13));
}

uint16_t f32_to_half(float f) {
    uint32_t bits = f32_to_bits(f);
    uint16_t sign = (uint16_t)((bits >> 31) & 1u);
    int exp = (int)((bits >> 23) & 0xffu);
    uint32_t mant = bits & 0x7fffffu;
    if (exp == 0xff) {
        if (mant == 0) return (uint16_t)((sign << 15) | (0x1f << 10));
        uint16_t m = (uint16_t)(mant >> 13);
        if (m < 1) m = 1; /* .max(1), halfvec.rs:106 */
        return (uint16_t)((sign << 15) | (0x1f << 10) | m);
    }
    if (exp > 142) return (uint16_t)((sign << 15) | (0x1f << 10));
    if (exp < 103) return (uint16_t)(sign << 15);
    if (exp < 113) {
        /* subnormal result (halfvec.rs:114-129) */
        int shift = 113 - exp;
        uint32_t full = mant | 0x800000u;
        uint32_t m = full >> (shift + 13);
        uint32_t round_bit = (full >> (shift + 12)) & 1u;
        uint32_t sticky = (full & ((1u << (shift + 12)) - 1u)) != 0 ? 1u : 0u;
        uint16_t result = (uint16_t)((sign << 15) | (uint16_t)m);
        if (round_bit != 0 && (sticky != 0 || (m & 1u) != 0)) return (uint16_t)(result + 1);
        return result;
    }
    uint16_t half_exp = (uint16_t)((exp - 127 + 15) & 0x1f);
    uint16_t half_mant = (uint16_t)(mant >> 13);
    uint32_t round_bit = (mant >> 12) & 1u;
    uint32_t sticky = mant & 0xfffu;
    uint16_t result = (uint16_t)((sign << 15) | (half_exp << 10) | half_mant);
    if (round_bit != 0 && (sticky != 0 || (half_mant & 1) != 0)) return (uint16_t)(result + 1);
    return result;
}

/* ======================================================================================
 * a13-a15: scalar distance kernels
 * ==================================================================================== */
/* PostgreSQL's pg_number_of_ones[256] (pg_bitutils.c) is the byte popcount table. */
struct PopTable {
    uint8_t t[256];
    PopTable() { for (int i = 0; i < 256; i++) { int c = 0; for (int b = 0; b < 8; b++) c += (i >> b) & 1; t[i] = (uint8_t)c; } }
};
const PopTable g_pop;

/* vector.rs:518-526 / :530-537 / :560-567 -- strict order, no contraction */
float f32_l2sq(int dim, const float *a, const float *b) {
    float d = 0.0f;
    for (int i = 0; i < dim; i++) { float diff = a[i] - b[i]; d += diff * diff; }
    return d;
}
float f32_ip(int dim, const float *a, const float *b) {
    float d = 0.0f;
    for (int i = 0; i < dim; i++) d += a[i] * b[i];
    return d;
}
float f32_l1(int dim, const float *a, const float *b) {
    float d = 0.0f;
    for (int i = 0; i < dim; i++) d += std::fabs(a[i] - b[i]);
    return d;
}
/* halfvec.rs:687-733 */
float f16_l2sq(int dim, const uint16_t *a, const uint16_t *b) {
    float d = 0.0f;
    for (int i = 0; i < dim; i++) { float diff = half_to_f32(a[i]) - half_to_f32(b[i]); d += diff * diff; }
    return d;
}
float f16_ip(int dim, const uint16_t *a, const uint16_t *b) {
    float d = 0.0f;
    for (int i = 0; i < dim; i++) d += half_to_f32(a[i]) * half_to_f32(b[i]);
    return d;
}
float f16_l1(int dim, const uint16_t *a, const uint16_t *b) {
    float d = 0.0f;
    for (int i = 0; i < dim; i++) d += std::fabs(half_to_f32(a[i]) - half_to_f32(b[i]));
    return d;
}
/* bitvec.rs:97-106 */
uint64_t bit_hamming(size_t nbytes, const uint8_t *a, const uint8_t *b) {
    uint64_t d = 0;
    for (size_t i = 0; i < nbytes; i++) d += g_pop.t[a[i] ^ b[i]];
    return d;
}
/* bitvec.rs:113-132 */
double bit_jaccard(size_t nbytes, const uint8_t *a, const uint8_t *b) {
    uint64_t ab = 0, aa = 0, bb = 0;
    for (size_t i = 0; i < nbytes; i++) {
        ab += g_pop.t[a[i] & b[i]];
        aa += g_pop.t[a[i]];
        bb += g_pop.t[b[i]];
    }
    if (ab == 0) return 1.0;
    return 1.0 - ((double)ab / (double)(aa + bb - ab));
}

size_t row_bytes_of(int elem_type, int dim) {
    switch (elem_type) {
        case ORC_ELEM_F32: return (size_t)dim * 4;
        case ORC_ELEM_F16: return (size_t)dim * 2;
        default: return ((size_t)dim + 7) / 8;
    }
}

/* the opclass FUNCTION 1 value (f64), wrappers at vector.rs:598-659, halfvec.rs:765-817,
 * bitvec.rs:144-167 */
double distance_f64(int elem_type, int metric, int dim, const void *a, const void *b) {
    if (elem_type == ORC_ELEM_F32) {
        const float *x = (const float *)a, *y = (const float *)b;
        switch (metric) {
            case ORC_METRIC_L2SQ: return (double)f32_l2sq(dim, x, y);
            case ORC_METRIC_NEG_IP: return -((double)f32_ip(dim, x, y));
            case ORC_METRIC_L1: return (double)f32_l1(dim, x, y);
        }
    } else if (elem_type == ORC_ELEM_F16) {
        const uint16_t *x = (const uint16_t *)a, *y = (const uint16_t *)b;
        switch (metric) {
            case ORC_METRIC_L2SQ: return (double)f16_l2sq(dim, x, y);
            case ORC_METRIC_NEG_IP: return -((double)f16_ip(dim, x, y));
            case ORC_METRIC_L1: return (double)f16_l1(dim, x, y);
        }
    } else {
        size_t nb = row_bytes_of(elem_type, dim);
        switch (metric) {
            case ORC_METRIC_HAMMING: return (double)bit_hamming(nb, (const uint8_t *)a, (const uint8_t *)b);
            case ORC_METRIC_JACCARD: return bit_jaccard(nb, (const uint8_t *)a, (const uint8_t *)b);
        }
    }
    return std::numeric_limits<double>::quiet_NaN();
}

/* ======================================================================================
 * a17: Rust std::collections::BinaryHeap restated (max-heap over a three-way compare).
 * push  = append + sift_up (move up while strictly greater than parent);
 * pop   = take last, swap with root, sift_down_to_bottom (always descend to the greater
 *         child, RIGHT child when left <= right), then sift_up from there;
 * into_vec = backing array order.
 * This code lives in the Rust standard library, not under /root/reference: UNPINNED.
 * ==================================================================================== */
template <class T, class Cmp>
struct RustHeap {
    std::vector<T> data;
    Cmp cmp; /* cmp(a,b) <0, 0, >0  in the heap's Ord */
    explicit RustHeap(Cmp c) : cmp(c) {}
    bool empty() const { return data.empty(); }
    size_t size() const { return data.size(); }
    const T &peek() const { return data[0]; }
    void sift_up(size_t start, size_t pos) {
        T elt = data[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (cmp(elt, data[parent]) <= 0) break;
            data[pos] = data[parent];
            pos = parent;
        }
        data[pos] = elt;
    }
    void sift_down_to_bottom(size_t pos) {
        size_t end = data.size(), start = pos;
        T elt = data[pos];
        size_t child = 2 * pos + 1;
        size_t lim = end >= 2 ? end - 2 : 0; /* end.saturating_sub(2) */
        while (child <= lim && end >= 2) {
            if (cmp(data[child], data[child + 1]) <= 0) child += 1;
            data[pos] = data[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (end >= 1 && child == end - 1) {
            data[pos] = data[child];
            pos = child;
        }
        data[pos] = elt;
        sift_up(start, pos);
    }
    void push(const T &v) {
        size_t old_len = data.size();
        data.push_back(v);
        sift_up(0, old_len);
    }
    T pop() {
        T item = data.back();
        data.pop_back();
        if (!data.empty()) {
            std::swap(item, data[0]);
            sift_down_to_bottom(0);
        }
        return item;
    }
};

int cmp_i(int32_t a, int32_t b) { return a < b ? -1 : (a > b ? 1 : 0); }

/* graph/mod.rs:17-22 */
struct Cand { float distance; int32_t idx; };
/* scan.rs:55-69 reduced to what the flat layout needs */
struct ScanCand { double distance; int32_t idx; };

template <class C> struct NearestCmp {   /* graph/mod.rs:103-112, scan.rs:85-93: reversed */
    bool total;
    int operator()(const C &a, const C &b) const {
        int c = a.distance < b.distance ? 1 : (a.distance > b.distance ? -1 : 0);
        if (c == 0 && total) c = -cmp_i(a.idx, b.idx);
        return c;
    }
};
template <class C> struct FurthestCmp {  /* graph/mod.rs:131-139, scan.rs:109-116 */
    bool total;
    int operator()(const C &a, const C &b) const {
        int c = a.distance < b.distance ? -1 : (a.distance > b.distance ? 1 : 0);
        if (c == 0 && total) c = cmp_i(a.idx, b.idx);
        return c;
    }
};

/* graph/mod.rs:58-84 */
struct Element {
    int level;
    std::vector<std::vector<Cand>> neighbors; /* [layer] -> stored order */
    std::vector<int64_t> heap_tids;           /* build.rs:245, <= HNSW_HEAPTIDS */
    /* state a vacuum leaves behind on disk (types/hnsw.rs:110-128,150-160); a fresh build has none of it */
    uint8_t deleted = 0;                      /* HnswElementTupleData.deleted (vacuum.rs:749)              */
    uint8_t version = 0;                      /* element tuple version (vacuum.rs:757)                     */
    uint8_t stale_neighbors = 0;              /* neighbour tuple version != element version (scan.rs:262)  */
};

const int HNSW_HEAPTIDS = 10; /* hnsw_constants.rs:85 */

inline int layer_m(int m, int layer) { return layer == 0 ? m * 2 : m; } /* hnsw_constants.rs:122-128 */

/* exact visited set: epoch-stamped array == HashSet<usize> semantics */
struct Visited {
    std::vector<uint32_t> stamp;
    uint32_t epoch = 0;
    void reset(size_t n) {
        if (stamp.size() < n) stamp.resize(n, 0);
        if (++epoch == 0) { std::fill(stamp.begin(), stamp.end(), 0); epoch = 1; }
    }
    bool contains(int32_t i) const { return stamp[(size_t)i] == epoch; }
    void insert(int32_t i) { stamp[(size_t)i] = epoch; }
};

} // namespace

struct orc_index {
    int elem_type, metric, dim, m, efc, tie_mode;
    size_t row_bytes;
    std::vector<uint8_t> values; /* row i at i*row_bytes */
    std::vector<Element> elements;
    int64_t entry = -1;
    orc_counters bc{};           /* build-time counters */
    Visited vis;                 /* single-threaded scratch */

    const void *row(int64_t i) const { return values.data() + (size_t)i * row_bytes; }
    double dist64(const void *a, const void *b) const { return distance_f64(elem_type, metric, dim, a, b); }
    /* build.rs:346-369: f64 -> f32 */
    float dist32(const void *a, const void *b) const { return (float)dist64(a, b); }
    bool total() const { return tie_mode == ORC_TIE_TOTAL_ORDER; }
};

namespace {

/* ======================================================================================
 * a1: search_layer  (src/graph/mod.rs:161-255)
 * ==================================================================================== */
std::vector<Cand> search_layer(orc_index &ix, Visited &vis, const std::vector<Cand> &entry_points,
                               size_t ef, int layer, const void *query, orc_counters &ctr) {
    vis.reset(ix.elements.size());
    RustHeap<Cand, NearestCmp<Cand>> candidates(NearestCmp<Cand>{ix.total()});
    RustHeap<Cand, FurthestCmp<Cand>> results(FurthestCmp<Cand>{ix.total()});
    size_t result_len = 0;

    for (const Cand &ep : entry_points) {                       /* :179-184 */
        vis.insert(ep.idx); ctr.n_visited++;
        candidates.push(ep);
        results.push(ep);
        result_len++;
    }
    while (!candidates.empty()) {                                /* :186 */
        Cand c = candidates.pop();
        float f_dist = results.empty() ? std::numeric_limits<float>::max() : results.peek().distance;
        if (c.distance > f_dist) break;                          /* :191 */
        const Element &ce = ix.elements[(size_t)c.idx];
        if (ce.level < layer) continue;                          /* :198 */
        if (layer == 0) ctr.n_expand0++; else ctr.n_expand_upper++;
        const std::vector<Cand> &nb = ce.neighbors[(size_t)layer];
        for (const Cand &n : nb) {                               /* :205 */
            if (vis.contains(n.idx)) continue;
            vis.insert(n.idx); ctr.n_visited++;                  /* :209 visited before level test */
            const Element &ee = ix.elements[(size_t)n.idx];
            if (ee.level < layer) continue;                      /* :214 */
            float e_distance = ix.dist32(query, ix.row(n.idx));  /* :218 */
            ctr.n_dist++;
            bool always_add = result_len < ef;
            float f2 = results.empty() ? std::numeric_limits<float>::max() : results.peek().distance;
            if (e_distance < f2 || always_add) {                 /* :229 */
                Cand e{e_distance, n.idx};
                candidates.push(e);
                results.push(e);
                result_len++;
                if (result_len > ef) { results.pop(); result_len--; } /* :239-242 */
            }
        }
    }
    std::vector<Cand> out = results.data;                        /* :248 into_iter = backing order */
    if (ix.total())
        std::sort(out.begin(), out.end(), [](const Cand &a, const Cand &b) {
            return a.distance < b.distance || (a.distance == b.distance && a.idx < b.idx); });
    else
        std::stable_sort(out.begin(), out.end(), [](const Cand &a, const Cand &b) {
            return a.distance < b.distance; });                  /* :249-253 stable sort_by */
    return out;
}

/* a3: check_element_closer (graph/mod.rs:315-339) */
bool check_element_closer(orc_index &ix, const Cand &e, const std::vector<Cand> &r, orc_counters &ctr) {
    const void *ev = ix.row(e.idx);
    for (const Cand &ri : r) {
        float d = ix.dist32(ev, ix.row(ri.idx));
        ctr.n_pair++;
        if (d <= e.distance) return false;                       /* :333 */
    }
    return true;
}

/* a3: select_neighbors (graph/mod.rs:269-308) */
std::vector<Cand> select_neighbors(orc_index &ix, const std::vector<Cand> &candidates, size_t max_neighbors,
                                   orc_counters &ctr) {
    if (candidates.size() <= max_neighbors) return candidates;   /* :276 */
    std::vector<Cand> result, discarded;
    result.reserve(max_neighbors);
    for (const Cand &e : candidates) {
        if (result.size() >= max_neighbors) break;               /* :285 */
        if (check_element_closer(ix, e, result, ctr)) result.push_back(e);
        else discarded.push_back(e);
    }
    for (const Cand &d : discarded) {                            /* :300-305 */
        if (result.size() >= max_neighbors) break;
        result.push_back(d);
    }
    return result;
}

/* a2: find_element_neighbors (graph/mod.rs:355-427) */
void find_element_neighbors(orc_index &ix, int32_t new_idx, int32_t entry_idx) {
    int new_level = ix.elements[(size_t)new_idx].level;
    int entry_level = ix.elements[(size_t)entry_idx].level;
    const void *q = ix.row(new_idx);
    float ep_distance = ix.dist32(q, ix.row(entry_idx));         /* :371-377 */
    ix.bc.n_dist++;
    std::vector<Cand> ep{Cand{ep_distance, entry_idx}};
    for (int lc = entry_level; lc >= new_level + 1; lc--) {      /* :385 */
        std::vector<Cand> w = search_layer(ix, ix.vis, ep, 1, lc, q, ix.bc);
        if (!w.empty()) ep = std::vector<Cand>{w[0]};            /* :396-398 */
    }
    int start_level = std::min(new_level, entry_level);
    for (int lc = start_level; lc >= 0; lc--) {                  /* :403 */
        size_t lm = (size_t)layer_m(ix.m, lc);
        std::vector<Cand> w = search_layer(ix, ix.vis, ep, (size_t)ix.efc, lc, q, ix.bc);
        ix.elements[(size_t)new_idx].neighbors[(size_t)lc] = select_neighbors(ix, w, lm, ix.bc);
        ep = std::move(w);                                       /* :425 */
    }
}

/* a4: update_neighbor_connections (graph/mod.rs:442-489) */
void update_neighbor_connections(orc_index &ix, int32_t new_idx) {
    int new_level = ix.elements[(size_t)new_idx].level;
    for (int lc = new_level; lc >= 0; lc--) {
        size_t lm = (size_t)layer_m(ix.m, lc);
        std::vector<Cand> snapshot = ix.elements[(size_t)new_idx].neighbors[(size_t)lc]; /* :455 */
        for (const Cand &hc : snapshot) {
            Cand new_candidate{hc.distance, new_idx};
            std::vector<Cand> &nb = ix.elements[(size_t)hc.idx].neighbors[(size_t)lc];
            if (nb.size() < lm) {
                nb.push_back(new_candidate);                     /* :469-471 */
            } else {
                std::vector<Cand> all = nb;                      /* :474-475 */
                all.push_back(new_candidate);
                std::stable_sort(all.begin(), all.end(), [](const Cand &a, const Cand &b) {
                    return a.distance < b.distance; });          /* :478-482 stable, distance only */
                std::vector<Cand> sel = select_neighbors(ix, all, lm, ix.bc);
                ix.elements[(size_t)hc.idx].neighbors[(size_t)lc] = std::move(sel); /* :485 */
            }
        }
    }
}

/* a5: one build_callback iteration (build.rs:400-535); row already normalised */
void build_insert(orc_index &ix, const void *rowp, int level, int64_t heap_tid) {
    size_t off = ix.values.size();
    ix.values.resize(off + ix.row_bytes);
    std::memcpy(ix.values.data() + off, rowp, ix.row_bytes);     /* :444-454 */
    Element el;
    el.level = level;
    el.neighbors.resize((size_t)level + 1);
    int32_t new_idx = (int32_t)ix.elements.size();
    ix.elements.push_back(std::move(el));                        /* :458-460 */

    if (ix.entry >= 0) {
        int32_t entry_idx = (int32_t)ix.entry;
        find_element_neighbors(ix, new_idx, entry_idx);          /* :464 */
        bool is_duplicate = false;
        for (const Cand &n : ix.elements[(size_t)new_idx].neighbors[0]) {   /* :484 */
            if (n.distance != 0.0f) break;                       /* :486 */
            if (std::memcmp(ix.row(new_idx), ix.row(n.idx), ix.row_bytes) == 0 &&
                (int)ix.elements[(size_t)n.idx].heap_tids.size() < HNSW_HEAPTIDS) {
                ix.elements[(size_t)n.idx].heap_tids.push_back(heap_tid);
                is_duplicate = true;
                break;
            }
        }
        if (is_duplicate) {                                      /* :505-512 */
            ix.elements.pop_back();
            ix.values.resize(off);
            return;
        }
        update_neighbor_connections(ix, new_idx);                /* :514 */
        if (ix.elements[(size_t)new_idx].level > ix.elements[(size_t)entry_idx].level)
            ix.entry = new_idx;                                  /* :523-525 strictly greater */
    } else {
        ix.entry = new_idx;                                      /* :526-529 */
    }
    ix.elements[(size_t)new_idx].heap_tids.push_back(heap_tid);  /* :532 */
}

/* ======================================================================================
 * a8: search_layer_disk  (src/index/scan.rs:302-448), flat-array form.
 * visited / discarded are caller-owned when non-null (iterative scan).
 * ==================================================================================== */
typedef RustHeap<ScanCand, NearestCmp<ScanCand>> DiscHeap;

std::vector<ScanCand> search_layer_scan(orc_index &ix, Visited &local_vis, std::vector<ScanCand> entry_points,
                                        size_t ef, int layer, const void *query, Visited *visited,
                                        DiscHeap *discarded, bool add_entry_to_visited, orc_counters &ctr,
                                        const std::vector<char> *skip_count = nullptr /* vacuum repair only, :314 */) {
    Visited *vis = visited;
    if (!vis) { local_vis.reset(ix.elements.size()); vis = &local_vis; }   /* :317-319 */
    RustHeap<ScanCand, NearestCmp<ScanCand>> candidates(NearestCmp<ScanCand>{ix.total()});
    RustHeap<ScanCand, FurthestCmp<ScanCand>> results(FurthestCmp<ScanCand>{ix.total()});
    size_t w_len = 0;
    for (const ScanCand &ep : entry_points) {                    /* :326-336 */
        if (add_entry_to_visited) { vis->insert(ep.idx); ctr.n_visited++; }
        candidates.push(ep);
        results.push(ep);
        if (!skip_count || !(*skip_count)[(size_t)ep.idx]) w_len++;   /* :330-335 */
    }
    while (!candidates.empty()) {                                /* :338 */
        ScanCand c = candidates.pop();
        double f_dist = results.empty() ? std::numeric_limits<double>::max() : results.peek().distance;
        if (c.distance > f_dist) {
            if (discarded) discarded->push(c);                   /* :340-345 */
            break;
        }
        const Element &ce = ix.elements[(size_t)c.idx];
        if (ce.level < layer) continue; /* load_neighbor_tids cannot address a layer above the level */
        if (layer == 0) ctr.n_expand0++; else ctr.n_expand_upper++;   /* the neighbour tuple is read either way */
        if (ce.stale_neighbors) continue;                        /* :262 version / count mismatch => :360 continue */
        const std::vector<Cand> &nb = ce.neighbors[(size_t)layer];  /* :349 -> :236-283 */
        for (const Cand &n : nb) {
            if (vis->contains(n.idx)) continue;                  /* :366 */
            vis->insert(n.idx); ctr.n_visited++;
            /* :179 a deleted element fails load_element before its distance is computed, with or without the
             * distance filter (:376-405): it stays in `visited` and goes nowhere else */
            if (ix.elements[(size_t)n.idx].deleted) continue;
            bool always_add = w_len < ef;                        /* :371 */
            double f2 = results.empty() ? std::numeric_limits<double>::max() : results.peek().distance;
            double d = ix.dist64(query, ix.row(n.idx));          /* :375 -> :185-192 */
            ctr.n_dist++;
            int e_level = ix.elements[(size_t)n.idx].level;
            if (!always_add && d >= f2) {                        /* :195-199 reject inside load_element */
                if (discarded && e_level >= layer) discarded->push(ScanCand{d, n.idx}); /* :385-405 */
                continue;
            }
            if (e_level < layer) continue;                       /* :410 */
            ScanCand e{d, n.idx};
            candidates.push(e);                                  /* :414-419 */
            results.push(e);
            if (!skip_count || !(*skip_count)[(size_t)e.idx]) w_len++;
            if (w_len > ef) {                                    /* :421-429: pops the furthest of W whatever it is */
                ScanCand ev = results.pop();
                w_len--;
                if (discarded) discarded->push(ev);
            }
        }
    }
    if (discarded) while (!candidates.empty()) discarded->push(candidates.pop()); /* :434-438 */
    std::vector<ScanCand> out = results.data;                    /* :441 */
    if (ix.total())
        std::sort(out.begin(), out.end(), [](const ScanCand &a, const ScanCand &b) {
            return a.distance > b.distance || (a.distance == b.distance && a.idx > b.idx); });
    else
        std::stable_sort(out.begin(), out.end(), [](const ScanCand &a, const ScanCand &b) {
            return a.distance > b.distance; });                  /* :442-446 descending, nearest last */
    return out;
}

/* a11: get_scan_items (scan.rs:458-530) */
std::vector<ScanCand> get_scan_items(orc_index &ix, Visited &scratch, const void *query, size_t ef_search,
                                     Visited *visited, DiscHeap *discarded, orc_counters &ctr) {
    if (ix.entry < 0) return {};                                 /* :470-472 */
    int32_t e = (int32_t)ix.entry;
    if (ix.elements[(size_t)e].deleted) return {};               /* :475-486 load_element => None (:179) */
    ScanCand ep{ix.dist64(query, ix.row(e)), e};                 /* :475-486 */
    ctr.n_dist++;
    int ep_level = ix.elements[(size_t)e].level;
    std::vector<ScanCand> ep_list{ep};
    for (int lc = ep_level; lc >= 1; lc--) {                     /* :492-512 fresh visited per layer */
        std::vector<ScanCand> w = search_layer_scan(ix, scratch, ep_list, 1, lc, query, nullptr, nullptr, true, ctr);
        if (w.empty()) return {};
        ep_list = std::vector<ScanCand>{w.back()};
    }
    if (visited) visited->reset(ix.elements.size());
    return search_layer_scan(ix, scratch, ep_list, ef_search, 0, query, visited, discarded, true, ctr); /* :515 */
}

/* ======================================================================================
 * aminsert on the flat arrays (src/index/insert.rs): the insert path of a LIVE index.  It differs from the
 * in-memory build (a2-a5) in three ways: the new element's lists are the nearest lm of W, no heuristic
 * (insert.rs:1096-1117); a full neighbour list is updated by REPLACING ONE SLOT, chosen by a heuristic in
 * which the new element is never pruned and never prunes (insert.rs:500-738); distances are recomputed
 * from the vectors in f64 (no cached edge distances on disk).  Page placement / slot reuse
 * (insert.rs:104-470) is storage-engine work and not restated: a new element always gets the next id.
 * ==================================================================================== */
/* find_element_neighbors_on_disk (insert.rs:1021-1123), skip_set = None */
std::vector<std::vector<ScanCand>> find_element_neighbors_on_disk(orc_index &ix, Visited &scratch, const void *query,
                                                                  int new_level, orc_counters &ctr,
                                                                  const std::vector<char> *skip = nullptr,
                                                                  int32_t entry_override = -1) {
    std::vector<std::vector<ScanCand>> nbl((size_t)new_level + 1);
    int32_t e = entry_override >= 0 ? entry_override : (int32_t)ix.entry;
    if (ix.elements[(size_t)e].deleted) return nbl;              /* :1036-1047 load_element => None */
    int entry_level = ix.elements[(size_t)e].level;
    std::vector<ScanCand> ep_list{ScanCand{ix.dist64(query, ix.row(e)), e}};
    ctr.n_dist++;
    for (int lc = entry_level; lc >= new_level + 1; lc--) {      /* :1052-1072 */
        std::vector<ScanCand> w = search_layer_scan(ix, scratch, ep_list, 1, lc, query, nullptr, nullptr, true, ctr, skip);
        if (w.empty()) return nbl;
        ep_list = std::vector<ScanCand>{w.back()};
    }
    for (int lc = std::min(new_level, entry_level); lc >= 0; lc--) {   /* :1075-1120 */
        size_t lm = (size_t)layer_m(ix.m, lc);
        size_t ef = (size_t)ix.efc + (skip ? 1 : 0);             /* :1080-1084 one more for the self-skip */
        std::vector<ScanCand> w = search_layer_scan(ix, scratch, ep_list, ef, lc, query, nullptr, nullptr, true, ctr, skip);
        std::vector<ScanCand> sel;                               /* :1101-1116 w is nearest-last: filter, rev().take(lm) */
        for (size_t i = w.size(); i-- > 0 && sel.size() < lm;)
            if (!skip || !(*skip)[(size_t)w[i].idx]) sel.push_back(w[i]);
        nbl[(size_t)lc] = std::move(sel);
        ep_list = std::move(w);
    }
    return nbl;
}

/* get_update_index (insert.rs:500-738): -1 = do not link, -2 = free slot, >= 0 = slot to overwrite */
int get_update_index(orc_index &ix, int32_t nb_idx, int layer, double new_distance, orc_counters &ctr) {
    const Element &nb = ix.elements[(size_t)nb_idx];
    if (nb.deleted) return -1;                                   /* :525-528 */
    if (nb.stale_neighbors || nb.level < layer) return -1;       /* :540-556 load_neighbor_tids => None */
    const std::vector<Cand> &tids = nb.neighbors[(size_t)layer];
    size_t lm = (size_t)layer_m(ix.m, layer);
    if (tids.size() < lm) return -2;                             /* :559-562 */
    struct HC { int32_t idx; double distance; bool is_new; };
    std::vector<HC> cand;
    int pruned_deleted = -1;
    for (size_t i = 0; i < tids.size(); i++) {                   /* :569-617 */
        const Element &x = ix.elements[(size_t)tids[i].idx];
        if (x.deleted || x.heap_tids.empty()) { if (pruned_deleted < 0) pruned_deleted = (int)i; continue; }
        cand.push_back(HC{tids[i].idx, ix.dist64(ix.row(nb_idx), ix.row(tids[i].idx)), false});
        ctr.n_pair++;
    }
    if (pruned_deleted >= 0) return pruned_deleted;              /* :620-623 */
    auto by_dist = [](const HC &a, const HC &b) { return a.distance < b.distance; };
    std::stable_sort(cand.begin(), cand.end(), by_dist);         /* :628-632 stable: ties keep slot order */
    cand.push_back(HC{-1, new_distance, true});                  /* :652-658 */
    std::stable_sort(cand.begin(), cand.end(), by_dist);         /* :659-663 the new element stays last among equals */
    std::vector<const HC *> selected, pruned;
    for (const HC &hc : cand) {                                  /* :671-705 */
        if (selected.size() >= lm) break;
        bool closer = true;
        for (const HC *sel : selected) {
            if (!hc.is_new && !sel->is_new) {                    /* :681-689; the new element is exempt both ways */
                double d = ix.dist64(ix.row(hc.idx), ix.row(sel->idx));
                ctr.n_pair++;
                if (d <= hc.distance) { closer = false; break; }
            }
        }
        if (closer) selected.push_back(&hc); else pruned.push_back(&hc);
    }
    for (const HC *pr : pruned) {                                /* :708-713 */
        if (selected.size() >= lm) break;
        selected.push_back(pr);
    }
    bool new_selected = false;
    for (const HC *sel : selected) new_selected = new_selected || sel->is_new;
    if (!new_selected) return -1;                                /* :716-720 */
    for (size_t i = 0; i < tids.size(); i++) {                   /* :731-738 first slot whose element was not kept */
        bool kept = false;
        for (const HC *sel : selected) kept = kept || (!sel->is_new && sel->idx == tids[i].idx);
        if (!kept) return (int)i;
    }
    return -1;
}

/* aminsert (insert.rs:1227-1480) for one row; row already normalised */
void am_insert(orc_index &ix, const void *rowp, int level, int64_t heap_tid) {
    size_t off = ix.values.size();
    ix.values.resize(off + ix.row_bytes);
    std::memcpy(ix.values.data() + off, rowp, ix.row_bytes);
    const void *query = ix.values.data() + off;
    int32_t new_idx = (int32_t)ix.elements.size();
    int entry_level = ix.entry >= 0 ? ix.elements[(size_t)ix.entry].level : -1;
    Visited scratch;
    std::vector<std::vector<ScanCand>> nbl((size_t)level + 1);
    if (ix.entry >= 0) nbl = find_element_neighbors_on_disk(ix, scratch, query, level, ix.bc);   /* :1317-1334 */
    for (const ScanCand &n : nbl[0]) {                           /* :1180-1213 find_duplicate_on_disk */
        if (n.distance != 0.0) break;
        if (std::memcmp(query, ix.row(n.idx), ix.row_bytes) != 0) continue;
        std::vector<int64_t> &t = ix.elements[(size_t)n.idx].heap_tids;
        if (t.empty() || (int)t.size() == HNSW_HEAPTIDS) continue;   /* :1159-1163 being deleted, or full */
        t.push_back(heap_tid);
        ix.values.resize(off);
        return;
    }
    Element el;
    el.level = level;
    el.neighbors.resize((size_t)level + 1);
    for (int lc = 0; lc <= level; lc++)                          /* :1390-1414 neighbour tuple, nearest first */
        for (const ScanCand &n : nbl[(size_t)lc]) el.neighbors[(size_t)lc].push_back(Cand{(float)n.distance, n.idx});
    el.heap_tids.push_back(heap_tid);
    ix.elements.push_back(std::move(el));
    for (int lc = level; lc >= 0; lc--) {                        /* :883-953 update_neighbors_on_disk */
        size_t lm = (size_t)layer_m(ix.m, lc);
        for (size_t t = 0; t < nbl[(size_t)lc].size() && t < lm; t++) {
            const ScanCand &sc = nbl[(size_t)lc][t];
            if (ix.elements[(size_t)sc.idx].deleted) continue;   /* :903-913 */
            int idx = get_update_index(ix, sc.idx, lc, sc.distance, ix.bc);
            if (idx == -1) continue;
            std::vector<Cand> &tids = ix.elements[(size_t)sc.idx].neighbors[(size_t)lc];   /* :793-870 */
            bool present = false;
            for (const Cand &c : tids) present = present || c.idx == new_idx;
            if (present) continue;
            Cand link{(float)sc.distance, new_idx};
            if (idx == -2) { if (tids.size() < lm) tids.push_back(link); }
            else tids[(size_t)idx] = link;
        }
    }
    if (ix.entry < 0 || level > entry_level) ix.entry = new_idx; /* :1454-1470 */
}

/* ======================================================================================
 * ambulkdelete on the flat arrays (src/index/vacuum.rs:816-842): pass 1 removes dead heap TIDs and collects the
 * elements left without any (:118-216), pass 2 re-finds the neighbours of every live element that points at one of
 * them or whose layer-0 list is not full (:228-281, :288-407, :413-640), pass 3 marks the emptied elements deleted
 * (:655-790).  Single backend, no concurrent inserts; page placement is not involved.
 * ==================================================================================== */
bool needs_updated(const orc_index &ix, const std::vector<char> &deleted, int32_t idx) {   /* :228-281 */
    const Element &e = ix.elements[(size_t)idx];
    for (const auto &l : e.neighbors)
        for (const Cand &c : l) if (deleted[(size_t)c.idx]) return true;
    return (int)e.neighbors[0].size() < layer_m(ix.m, 0);
}

void repair_graph_element(orc_index &ix, Visited &scratch, const std::vector<char> &deleted, int32_t idx,
                          int32_t entry, orc_counters &ctr) {                              /* :288-407 */
    if (idx == entry) return;                                    /* :300-302 */
    std::vector<char> skip = deleted;                            /* :338-339 deleted + self */
    skip[(size_t)idx] = 1;
    Element &e = ix.elements[(size_t)idx];
    std::vector<std::vector<ScanCand>> nbl =
        find_element_neighbors_on_disk(ix, scratch, ix.row(idx), e.level, ctr, &skip, entry);
    for (int lc = 0; lc <= e.level; lc++) {                      /* :366-377 the whole tuple is rewritten */
        std::vector<Cand> &l = e.neighbors[(size_t)lc];
        l.clear();
        for (const ScanCand &n : nbl[(size_t)lc]) l.push_back(Cand{(float)n.distance, n.idx});
    }
}

int64_t bulk_delete(orc_index &ix, const int64_t *dead, int64_t n_dead, orc_counters &ctr) {
    std::vector<int64_t> dead_sorted(dead, dead + n_dead);
    std::sort(dead_sorted.begin(), dead_sorted.end());
    auto is_dead = [&](int64_t t) { return std::binary_search(dead_sorted.begin(), dead_sorted.end(), t); };
    const size_t n = ix.elements.size();
    std::vector<char> deleted(n, 0);
    bool any_deleted = false;
    int64_t removed = 0;
    int32_t highest_point = -1;
    int highest_level = -1;
    for (size_t i = 0; i < n; i++) {                             /* pass 1, :118-216 */
        Element &e = ix.elements[i];
        if (!e.heap_tids.empty()) {
            std::vector<int64_t> keep;
            for (int64_t t : e.heap_tids) { if (is_dead(t)) removed++; else keep.push_back(t); }
            e.heap_tids.swap(keep);
        }
        if (e.heap_tids.empty()) { deleted[i] = 1; any_deleted = true; }       /* :188-190 (also elements deleted earlier) */
        else if (e.level > highest_level && (int64_t)i != ix.entry) { highest_point = (int32_t)i; highest_level = e.level; }
    }
    if (!any_deleted) return removed;
    Visited scratch;
    /* pass 2: repair_graph_entry_point (:413-523) */
    if (highest_point >= 0 && !ix.elements[(size_t)highest_point].deleted && needs_updated(ix, deleted, highest_point))
        repair_graph_element(ix, scratch, deleted, highest_point, (int32_t)ix.entry, ctr);
    if (ix.entry >= 0) {
        if (deleted[(size_t)ix.entry]) {
            ix.entry = highest_point;                            /* :464-484 (-1 clears the entry point) */
        } else if (!ix.elements[(size_t)ix.entry].deleted && needs_updated(ix, deleted, (int32_t)ix.entry)) {
            /* :497-514 searched from the highest point; with none, the element is its own entry and nothing happens */
            int32_t from = highest_point >= 0 ? highest_point : (int32_t)ix.entry;
            repair_graph_element(ix, scratch, deleted, (int32_t)ix.entry, from, ctr);
        }
    }
    for (size_t i = 0; i < n; i++) {                             /* :549-640 every other live element, in page order */
        Element &e = ix.elements[i];
        if (e.heap_tids.empty() || e.deleted) continue;          /* :566-570, :580-592 */
        if (!needs_updated(ix, deleted, (int32_t)i)) continue;
        int entry_level = ix.entry >= 0 ? ix.elements[(size_t)ix.entry].level : -1;
        if (ix.entry < 0) { ix.entry = (int32_t)i; continue; }   /* :606-625: no entry point: nothing to search, it becomes one */
        repair_graph_element(ix, scratch, deleted, (int32_t)i, (int32_t)ix.entry, ctr);
        if (e.level > entry_level) ix.entry = (int32_t)i;        /* :614-623 HNSW_UPDATE_ENTRY_GREATER */
    }
    for (size_t i = 0; i < n; i++) {                             /* pass 3, :655-790 */
        Element &e = ix.elements[i];
        if (e.deleted || !e.heap_tids.empty()) continue;
        for (auto &l : e.neighbors) l.clear();
        e.deleted = 1;
        std::memset(ix.values.data() + i * ix.row_bytes, 0, ix.row_bytes);
        e.version = e.version >= 15 ? 1 : (uint8_t)(e.version + 1);
    }
    return removed;
}

void add_counters(orc_counters &a, const orc_counters &b) {
    a.n_dist += b.n_dist; a.n_expand0 += b.n_expand0; a.n_expand_upper += b.n_expand_upper;
    a.n_visited += b.n_visited; a.n_pair += b.n_pair;
}

} // namespace

/* ======================================================================================
 * C API
 * ==================================================================================== */
extern "C" {

float orc_half_to_f32(uint16_t h) { return half_to_f32(h); }
uint16_t orc_f32_to_half(float f) { return f32_to_half(f); }

double orc_distance(int elem_type, int metric, int dim, const void *a, const void *b) {
    return distance_f64(elem_type, metric, dim, a, b);
}

double orc_normalize(int elem_type, int dim, const void *in, void *out) {
    if (elem_type == ORC_ELEM_F32) {                             /* vector.rs:106-126 */
        const float *ax = (const float *)in;
        float *rx = (float *)out;
        double norm = 0.0;
        for (int i = 0; i < dim; i++) { double v = (double)ax[i]; norm += v * v; }
        norm = std::sqrt(norm);
        for (int i = 0; i < dim; i++) rx[i] = norm > 0.0 ? (float)((double)ax[i] / norm) : 0.0f;
        return norm;
    }
    if (elem_type == ORC_ELEM_F16) {                             /* halfvec.rs:204-233 */
        const uint16_t *ax = (const uint16_t *)in;
        uint16_t *rx = (uint16_t *)out;
        double norm = 0.0;
        for (int i = 0; i < dim; i++) { double v = (double)half_to_f32(ax[i]); norm += v * v; }
        norm = std::sqrt(norm);
        for (int i = 0; i < dim; i++)
            rx[i] = norm > 0.0 ? f32_to_half((float)((double)half_to_f32(ax[i]) / norm)) : (uint16_t)0;
        for (int i = 0; i < dim; i++) if (std::isinf(half_to_f32(rx[i]))) return -1.0;
        return norm;
    }
    return std::numeric_limits<double>::quiet_NaN();
}

/* types/hnsw.rs:337-349 with BLCKSZ 8192, page header 24, opaque 8, ntup header 4, ItemId 4, TID 6 */
int orc_max_level(int m) {
    int available = 8192 - 24 - 8 - 4 - 4;
    int v = available / 6 / m - 2;
    return v < 255 ? v : 255;
}
/* build.rs:373-377 */
int orc_level_from_uniform(double r, int m) {
    double rr = r > std::numeric_limits<double>::min() ? r : std::numeric_limits<double>::min();
    double ml = 1.0 / std::log((double)m);                       /* hnsw_constants.rs:132-134 */
    double lv = std::floor(-std::log(rr) * ml);
    int maxl = orc_max_level(m);
    if (!(lv < (double)maxl)) return maxl;
    return (int)lv;
}

orc_index *orc_create(int elem_type, int metric, int dim, int m, int ef_construction, int tie_mode) {
    bool ok = (elem_type == ORC_ELEM_BIT) ? (metric == ORC_METRIC_HAMMING || metric == ORC_METRIC_JACCARD)
                                          : (metric >= ORC_METRIC_L2SQ && metric <= ORC_METRIC_L1);
    if (!ok || dim <= 0 || m < 2 || m > 100 || ef_construction < 4 || ef_construction > 1000 ||
        ef_construction < 2 * m)                                 /* options.rs:203-225, build.rs:858-867 */
        return nullptr;
    orc_index *ix = new orc_index();
    ix->elem_type = elem_type; ix->metric = metric; ix->dim = dim; ix->m = m; ix->efc = ef_construction;
    ix->tie_mode = tie_mode;
    ix->row_bytes = row_bytes_of(elem_type, dim);
    return ix;
}
void orc_destroy(orc_index *ix) { delete ix; }
size_t orc_row_bytes(const orc_index *ix) { return ix->row_bytes; }

int64_t orc_build(orc_index *ix, const void *rows, int64_t n, const int32_t *levels, int64_t base_tid) {
    const uint8_t *p = (const uint8_t *)rows;
    for (int64_t i = 0; i < n; i++) build_insert(*ix, p + (size_t)i * ix->row_bytes, levels[i], base_tid + i);
    return (int64_t)ix->elements.size();
}

int64_t orc_insert(orc_index *ix, const void *rows, int64_t n, const int32_t *levels, int64_t base_tid) {
    const uint8_t *p = (const uint8_t *)rows;
    for (int64_t i = 0; i < n; i++) am_insert(*ix, p + (size_t)i * ix->row_bytes, levels[i], base_tid + i);
    return (int64_t)ix->elements.size();
}

int64_t orc_bulk_delete(orc_index *ix, const int64_t *dead_tids, int64_t n_dead) {
    return bulk_delete(*ix, dead_tids, n_dead, ix->bc);
}

int64_t orc_n_elements(const orc_index *ix) { return (int64_t)ix->elements.size(); }
int64_t orc_entry_point(const orc_index *ix) { return ix->entry; }
int orc_element_level(const orc_index *ix, int64_t id) { return ix->elements[(size_t)id].level; }
int orc_neighbors(const orc_index *ix, int64_t id, int layer, int32_t *ids, float *dists, int cap) {
    const Element &e = ix->elements[(size_t)id];
    if (layer > e.level) return 0;
    const std::vector<Cand> &nb = e.neighbors[(size_t)layer];
    for (int i = 0; i < (int)nb.size() && i < cap; i++) {
        if (ids) ids[i] = nb[(size_t)i].idx;
        if (dists) dists[i] = nb[(size_t)i].distance;
    }
    return (int)nb.size();
}
int orc_heap_tids(const orc_index *ix, int64_t id, int64_t *tids, int cap) {
    const std::vector<int64_t> &t = ix->elements[(size_t)id].heap_tids;
    for (int i = 0; i < (int)t.size() && i < cap; i++) tids[i] = t[(size_t)i];
    return (int)t.size();
}
const void *orc_element_row(const orc_index *ix, int64_t id) { return ix->row(id); }
void orc_build_counters(const orc_index *ix, orc_counters *out) { *out = ix->bc; }

int64_t orc_export_graph(const orc_index *ix, uint8_t *level, int32_t *adj0, int32_t *upper_base,
                         int32_t *adj_up) {
    int64_t n = (int64_t)ix->elements.size(), rows = 0;
    int m = ix->m;
    for (int64_t i = 0; i < n; i++) {
        const Element &e = ix->elements[(size_t)i];
        if (level) level[i] = (uint8_t)e.level;
        if (adj0) {
            int32_t *r = adj0 + (size_t)i * 2 * m;
            for (int j = 0; j < 2 * m; j++) r[j] = j < (int)e.neighbors[0].size() ? e.neighbors[0][(size_t)j].idx : -1;
        }
        if (upper_base) upper_base[i] = e.level > 0 ? (int32_t)rows : -1;
        for (int l = 1; l <= e.level; l++) {
            if (adj_up) {
                int32_t *r = adj_up + (size_t)(rows + l - 1) * m;
                for (int j = 0; j < m; j++) r[j] = j < (int)e.neighbors[(size_t)l].size() ? e.neighbors[(size_t)l][(size_t)j].idx : -1;
            }
        }
        rows += e.level;
    }
    return rows;
}

int orc_import_graph(orc_index *ix, const void *rows, int64_t n, const uint8_t *level,
                     const int32_t *adj0, const int32_t *upper_base, const int32_t *adj_up, int64_t entry) {
    ix->values.assign((const uint8_t *)rows, (const uint8_t *)rows + (size_t)n * ix->row_bytes);
    ix->elements.clear();
    ix->elements.resize((size_t)n);
    int m = ix->m;
    for (int64_t i = 0; i < n; i++) {
        Element &e = ix->elements[(size_t)i];
        e.level = level[i];
        e.neighbors.resize((size_t)e.level + 1);
        e.heap_tids.push_back(i);
        for (int l = 0; l <= e.level; l++) {
            int lm = layer_m(m, l);
            const int32_t *r = l == 0 ? adj0 + (size_t)i * 2 * m : adj_up + (size_t)(upper_base[i] + l - 1) * m;
            for (int j = 0; j < lm && r[j] >= 0; j++)
                e.neighbors[(size_t)l].push_back(Cand{ix->dist32(ix->row(i), ix->row(r[j])), r[j]});
        }
    }
    ix->entry = entry;
    return 0;
}

/* Fixture states a vacuum leaves on disk.  The repair pass itself (vacuum.rs:288-640) is out of scope; these
 * only produce what a scan can meet while or after one runs. */
int orc_vacuum_mark(orc_index *ix, int64_t id, int what) {
    if (id < 0 || id >= (int64_t)ix->elements.size()) return -1;
    Element &e = ix->elements[(size_t)id];
    auto bump = [](uint8_t v) -> uint8_t { return v >= 15 ? 1 : (uint8_t)(v + 1); };   /* vacuum.rs:797-803 */
    if (what == ORC_VACUUM_REMOVE_HEAPTIDS) {          /* vacuum.rs:118-190: every heap TID of the element was dead */
        e.heap_tids.clear();
    } else if (what == ORC_VACUUM_MARK_DELETED) {      /* vacuum.rs:717-758 */
        e.heap_tids.clear();
        for (auto &l : e.neighbors) l.clear();         /* every neighbour TID set invalid */
        e.deleted = 1;
        std::memset(ix->values.data() + (size_t)id * ix->row_bytes, 0, ix->row_bytes);   /* value zeroed */
        e.version = bump(e.version);                   /* both tuples get the new version */
    } else if (what == ORC_VACUUM_STALE_NEIGHBORS) {   /* element rewritten, neighbour tuple not yet (scan.rs:262) */
        e.stale_neighbors = 1;
    } else return -1;
    return 0;
}

void orc_element_flags(const orc_index *ix, uint8_t *flags) {
    for (size_t i = 0; i < ix->elements.size(); i++) {
        const Element &e = ix->elements[i];
        flags[i] = (uint8_t)((e.deleted ? ORC_ELEM_DELETED : 0) | (e.heap_tids.empty() ? ORC_ELEM_NO_HEAPTIDS : 0));
    }
}

int orc_search_layer(orc_index *ix, const void *query, const int32_t *ep_ids, const float *ep_dists,
                     int n_ep, int ef, int layer, int32_t *out_ids, float *out_dists) {
    std::vector<Cand> ep;
    for (int i = 0; i < n_ep; i++) ep.push_back(Cand{ep_dists[i], ep_ids[i]});
    orc_counters c{};
    std::vector<Cand> w = search_layer(*ix, ix->vis, ep, (size_t)ef, layer, query, c);
    for (size_t i = 0; i < w.size(); i++) { out_ids[i] = w[i].idx; out_dists[i] = w[i].distance; }
    return (int)w.size();
}

int orc_select_neighbors(orc_index *ix, const int32_t *cand_ids, const float *cand_dists, int n_cand,
                         int max_neighbors, int32_t *out_ids, float *out_dists) {
    std::vector<Cand> c;
    for (int i = 0; i < n_cand; i++) c.push_back(Cand{cand_dists[i], cand_ids[i]});
    orc_counters ctr{};
    std::vector<Cand> r = select_neighbors(*ix, c, (size_t)max_neighbors, ctr);
    for (size_t i = 0; i < r.size(); i++) { out_ids[i] = r[i].idx; if (out_dists) out_dists[i] = r[i].distance; }
    return (int)r.size();
}

int orc_search_batch(orc_index *ix, const void *queries, int64_t nq, int ef_search, int k,
                     int32_t *out_ids, float *out_dists, int32_t *out_counts, orc_counters *counters,
                     int nthreads) {
    if (ef_search < 1 || ef_search > 1000 || k < 1) return -1;   /* options.rs:81,157-166 */
    if (nthreads < 1) nthreads = 1;
    std::vector<orc_counters> tc((size_t)nthreads);
    auto work = [&](int t) {
        Visited scratch;
        orc_counters c{};
        int64_t lo = nq * t / nthreads, hi = nq * (t + 1) / nthreads;
        for (int64_t q = lo; q < hi; q++) {
            const void *qp = (const uint8_t *)queries + (size_t)q * ix->row_bytes;
            std::vector<ScanCand> res = get_scan_items(*ix, scratch, qp, (size_t)ef_search, nullptr, nullptr, c);
            /* amgettuple pops from the back: nearest first (scan.rs:861) */
            int cnt = 0;
            for (size_t i = res.size(); i-- > 0 && cnt < k;) {
                if (ix->elements[(size_t)res[i].idx].heap_tids.empty()) continue;   /* scan.rs:866-868 */
                out_ids[(size_t)q * k + cnt] = res[i].idx;
                out_dists[(size_t)q * k + cnt] = (float)res[i].distance;
                cnt++;
            }
            if (out_counts) out_counts[q] = cnt;
            for (int j = cnt; j < k; j++) {
                out_ids[(size_t)q * k + j] = -1;
                out_dists[(size_t)q * k + j] = std::numeric_limits<float>::infinity();
            }
        }
        tc[(size_t)t] = c;
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    if (counters) { *counters = orc_counters{}; for (auto &c : tc) add_counters(*counters, c); }
    return 0;
}

/* a12: amgettuple stream (scan.rs:709-875) */
int64_t orc_scan_stream(orc_index *ix, const void *query, int ef_search, int iterative_mode,
                        int64_t max_scan_tuples, int64_t max_emit, int64_t *out_tids,
                        int32_t *out_elem_ids, double *out_dists, orc_counters *counters) {
    Visited scratch, visited;
    DiscHeap discarded(NearestCmp<ScanCand>{ix->total()});
    orc_counters c{};
    bool use_iter = iterative_mode != ORC_ITER_OFF;
    std::vector<ScanCand> results = get_scan_items(*ix, scratch, query, (size_t)ef_search,
                                                   use_iter ? &visited : nullptr,
                                                   use_iter ? &discarded : nullptr, c);   /* :764-784 */
    bool iterative_initialized = use_iter && ix->entry >= 0;
    int64_t tuples = 0, emitted = 0;
    double previous_distance = -std::numeric_limits<double>::infinity();
    bool have_current = false;
    ScanCand current{0, -1};
    std::vector<int64_t> cur_tids;
    while (emitted < max_emit) {                                 /* :794 */
        if (have_current) {
            if (!cur_tids.empty()) {
                int64_t tid = cur_tids.back();                   /* :798 pop from the back */
                cur_tids.pop_back();
                if (iterative_mode == ORC_ITER_STRICT) {         /* :801-806 */
                    if (current.distance < previous_distance) continue;
                    previous_distance = current.distance;
                }
                out_tids[emitted] = tid;
                if (out_elem_ids) out_elem_ids[emitted] = current.idx;
                if (out_dists) out_dists[emitted] = current.distance;
                emitted++;
                continue;
            }
            have_current = false;
        }
        if (results.empty()) {                                   /* :817 */
            if (iterative_mode == ORC_ITER_OFF) break;
            if (!iterative_initialized) break;
            if (tuples >= max_scan_tuples) {                     /* :831-841 */
                if (discarded.empty()) break;
                results.push_back(discarded.pop());
            } else {                                             /* :844-853 -> :538-577 */
                if (discarded.empty()) break;
                std::vector<ScanCand> ep;
                for (int i = 0; i < ef_search && !discarded.empty(); i++) ep.push_back(discarded.pop());
                results = search_layer_scan(*ix, scratch, ep, (size_t)ef_search, 0, query, &visited, &discarded,
                                            false, c);
            }
            if (results.empty()) break;
        }
        ScanCand sc = results.back();                            /* :861 */
        results.pop_back();
        const std::vector<int64_t> &t = ix->elements[(size_t)sc.idx].heap_tids;
        if (t.empty()) continue;                                 /* :867 */
        tuples++;                                                /* :871 */
        current = sc;
        cur_tids = t;
        have_current = true;
    }
    if (counters) *counters = c;
    return emitted;
}

int orc_flat_topk(int elem_type, int metric, int dim, const void *rows, int64_t n, const void *queries,
                  int64_t nq, int k, int32_t *out_ids, float *out_dists, int nthreads) {
    size_t rb = row_bytes_of(elem_type, dim);
    if (nthreads < 1) nthreads = 1;
    auto work = [&](int t) {
        std::vector<std::pair<double, int32_t>> d((size_t)n);
        int64_t lo = nq * t / nthreads, hi = nq * (t + 1) / nthreads;
        for (int64_t q = lo; q < hi; q++) {
            const void *qp = (const uint8_t *)queries + (size_t)q * rb;
            for (int64_t i = 0; i < n; i++)
                d[(size_t)i] = {distance_f64(elem_type, metric, dim, qp, (const uint8_t *)rows + (size_t)i * rb), (int32_t)i};
            size_t kk = (size_t)std::min<int64_t>(k, n);
            std::partial_sort(d.begin(), d.begin() + (long)kk, d.end());
            for (size_t j = 0; j < (size_t)k; j++) {
                out_ids[(size_t)q * k + j] = j < kk ? d[j].second : -1;
                out_dists[(size_t)q * k + j] = j < kk ? (float)d[j].first : std::numeric_limits<float>::infinity();
            }
        }
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    return 0;
}

} // extern "C"

/* ======================================================================================
 * a19: on-disk pages, written exactly as the flush phase of ambuild does
 * (src/index/build.rs:545-821; structs src/types/hnsw.rs:18-169; PostgreSQL bufpage.h layout).
 * Test fixture for the product's page exporter: no PostgreSQL is available in this image.
 * ==================================================================================== */
namespace {
const size_t BLCKSZ = 8192, PAGE_HDR = 24, OPAQUE = 8, ITEMID = 4, ETUP_HDR = 72, NTUP_HDR = 4, TID = 6;
inline size_t maxalign(size_t x) { return (x + 7) & ~(size_t)7; }                 /* hnsw.rs:316-320 */
inline void put16(uint8_t *p, uint16_t v) { std::memcpy(p, &v, 2); }
inline void put32(uint8_t *p, uint32_t v) { std::memcpy(p, &v, 4); }
inline uint16_t get16(const uint8_t *p) { uint16_t v; std::memcpy(&v, p, 2); return v; }
inline void put_tid(uint8_t *p, uint32_t blk, uint16_t off) { put16(p, (uint16_t)(blk >> 16)); put16(p + 2, (uint16_t)(blk & 0xffff)); put16(p + 4, off); }
inline void put_invalid_tid(uint8_t *p) { put_tid(p, 0xFFFFFFFFu, 0); }

struct PageWriter {
    std::vector<uint8_t> buf;
    size_t n_pages() const { return buf.size() / BLCKSZ; }
    uint8_t *page(size_t b) { return buf.data() + b * BLCKSZ; }
    size_t new_page() {                                   /* hnsw_new_buffer + hnsw_init_page, build.rs:37-64 */
        size_t b = n_pages();
        buf.resize(buf.size() + BLCKSZ, 0);
        uint8_t *p = page(b);
        put16(p + 12, (uint16_t)PAGE_HDR);                /* pd_lower   */
        put16(p + 14, (uint16_t)(BLCKSZ - OPAQUE));       /* pd_upper   */
        put16(p + 16, (uint16_t)(BLCKSZ - OPAQUE));       /* pd_special */
        put16(p + 18, (uint16_t)(BLCKSZ | 4));            /* pd_pagesize_version */
        put32(p + BLCKSZ - OPAQUE, 0xFFFFFFFFu);          /* nextblkno = InvalidBlockNumber */
        put16(p + BLCKSZ - OPAQUE + 4, 0);
        put16(p + BLCKSZ - OPAQUE + 6, 0xFF90);           /* HNSW_PAGE_ID */
        return b;
    }
    size_t free_space(size_t b) {                         /* PageGetFreeSpace */
        uint8_t *p = page(b);
        int space = (int)get16(p + 14) - (int)get16(p + 12);
        if (space < (int)ITEMID) return 0;
        return (size_t)space - ITEMID;
    }
    uint16_t max_offset(size_t b) { return (uint16_t)((get16(page(b) + 12) - PAGE_HDR) / ITEMID); }
    uint16_t add_item(size_t b, const uint8_t *item, size_t size) {   /* PageAddItemExtended */
        uint8_t *p = page(b);
        uint16_t lower = get16(p + 12), upper = get16(p + 14);
        size_t al = maxalign(size);
        uint16_t off = (uint16_t)((lower - PAGE_HDR) / ITEMID + 1);
        upper = (uint16_t)(upper - al);
        std::memcpy(p + upper, item, size);
        uint32_t lp = (uint32_t)upper | (1u << 15) /* LP_NORMAL */ | ((uint32_t)size << 17);
        put32(p + lower, lp);
        put16(p + 12, (uint16_t)(lower + ITEMID));
        put16(p + 14, upper);
        return off;
    }
    uint8_t *item_ptr(size_t b, uint16_t off) {
        uint8_t *p = page(b);
        uint32_t lp; std::memcpy(&lp, p + PAGE_HDR + (size_t)(off - 1) * ITEMID, 4);
        return p + (lp & 0x7fff);
    }
};
} // namespace

extern "C" int64_t orc_write_pages(const orc_index *ix, uint8_t *pages, int64_t cap_pages) {
    PageWriter w;
    const size_t max_size = BLCKSZ - maxalign(PAGE_HDR) - maxalign(OPAQUE) - ITEMID;   /* hnsw.rs:325-331 */
    /* varlena header bytes in front of the raw row: vector/halfvec {vl_len_, i16 dim, i16 unused},
     * bit {vl_len_, i32 bit_len} (vector.rs:43-48, halfvec.rs:41-46, bitvec.rs:28-32) */
    const size_t vhdr = 8, value_size = vhdr + ix->row_bytes;
    /* create_meta_page (build.rs:545-569) */
    size_t mb = w.new_page();
    {
        uint8_t *m = w.page(mb) + PAGE_HDR;
        put32(m, 0xA953A953u); put32(m + 4, 1); put32(m + 8, (uint32_t)ix->dim);
        put16(m + 12, (uint16_t)ix->m); put16(m + 14, (uint16_t)ix->efc);
        put32(m + 16, 0xFFFFFFFFu); put16(m + 20, 0); put16(m + 22, (uint16_t)(int16_t)-1); put32(m + 24, 0xFFFFFFFFu);
        put16(w.page(mb) + 12, (uint16_t)(PAGE_HDR + 28));
    }
    /* create_graph_pages (build.rs:576-702) */
    size_t cur = w.new_page();
    struct Loc { uint32_t blk; uint16_t off; uint32_t nblk; uint16_t noff; };
    std::vector<Loc> locs(ix->elements.size());
    std::vector<uint8_t> etup, ntup;
    for (size_t idx = 0; idx < ix->elements.size(); idx++) {
        const Element &e = ix->elements[idx];
        size_t etup_size = maxalign(ETUP_HDR + value_size);
        size_t ntup_size = maxalign(NTUP_HDR + (size_t)(e.level + 2) * ix->m * TID);
        size_t combined = etup_size + ntup_size + ITEMID;
        etup.assign(etup_size, 0);
        etup[0] = 1; etup[1] = (uint8_t)e.level; etup[2] = e.deleted; etup[3] = e.version;
        for (int i = 0; i < HNSW_HEAPTIDS; i++) {
            if (i < (int)e.heap_tids.size())   /* heap tid = (block tid >> 16, offset tid & 0xffff) for the fixture */
                put_tid(etup.data() + 4 + i * TID, (uint32_t)(e.heap_tids[(size_t)i] >> 16), (uint16_t)((e.heap_tids[(size_t)i] & 0xffff) + 1));
            else put_invalid_tid(etup.data() + 4 + i * TID);
        }
        uint8_t *val = etup.data() + ETUP_HDR;
        put32(val, (uint32_t)(value_size << 2));          /* 4-byte varlena header */
        if (ix->elem_type == ORC_ELEM_BIT) put32(val + 4, (uint32_t)ix->dim);
        else { put16(val + 4, (uint16_t)ix->dim); put16(val + 6, 0); }
        std::memcpy(val + vhdr, ix->row((int64_t)idx), ix->row_bytes);
        if (e.deleted) std::memset(val, 0, value_size);   /* vacuum.rs:750-754 zeroes the whole varlena, header included */
        size_t fs = w.free_space(cur);
        bool need_new = fs < etup_size || (combined <= max_size && fs < combined);      /* build.rs:636-638 */
        if (need_new) { size_t nb = w.new_page(); put32(w.page(cur) + BLCKSZ - OPAQUE, (uint32_t)nb); cur = nb; }
        uint32_t eblk = (uint32_t)cur;
        uint16_t eoff = (uint16_t)(w.max_offset(cur) + 1);
        Loc L;
        L.blk = eblk; L.off = eoff;
        if (combined <= max_size) { L.nblk = eblk; L.noff = (uint16_t)(eoff + 1); } else { L.nblk = eblk + 1; L.noff = 1; }
        locs[idx] = L;
        put_tid(etup.data() + 4 + HNSW_HEAPTIDS * TID, L.nblk, L.noff);                 /* neighbortid */
        w.add_item(cur, etup.data(), etup_size);
        if (w.free_space(cur) < ntup_size) { size_t nb = w.new_page(); put32(w.page(cur) + BLCKSZ - OPAQUE, (uint32_t)nb); cur = nb; }
        ntup.assign(ntup_size, 0);
        ntup[0] = 2;
        w.add_item(cur, ntup.data(), ntup_size);
    }
    size_t insert_page = cur;
    /* write_neighbor_tuples (build.rs:719-790): layers from the top down, lm slots each */
    for (size_t idx = 0; idx < ix->elements.size(); idx++) {
        const Element &e = ix->elements[idx];
        uint8_t *nt = w.item_ptr(locs[idx].nblk, locs[idx].noff);
        nt[0] = 2;
        nt[1] = e.stale_neighbors ? (uint8_t)(e.version + 1) : e.version;   /* hnsw.rs:150-160 version */
        size_t t = 0;
        for (int lc = e.level; lc >= 0; lc--) {
            int lm = layer_m(ix->m, lc);
            for (int i = 0; i < lm; i++, t++) {
                uint8_t *slot = nt + NTUP_HDR + t * TID;
                if (i < (int)e.neighbors[(size_t)lc].size()) {
                    const Loc &n = locs[(size_t)e.neighbors[(size_t)lc][(size_t)i].idx];
                    put_tid(slot, n.blk, n.off);
                } else put_invalid_tid(slot);
            }
        }
        put16(nt + 2, (uint16_t)t);
    }
    /* update_meta_page (build.rs:797-821) */
    {
        uint8_t *m = w.page(mb) + PAGE_HDR;
        if (ix->entry >= 0) {
            put32(m + 16, locs[(size_t)ix->entry].blk); put16(m + 20, locs[(size_t)ix->entry].off);
            put16(m + 22, (uint16_t)(int16_t)ix->elements[(size_t)ix->entry].level);
        }
        put32(m + 24, (uint32_t)insert_page);
    }
    int64_t np = (int64_t)w.n_pages();
    if (pages && cap_pages >= np) std::memcpy(pages, w.buf.data(), w.buf.size());
    return np;
}

