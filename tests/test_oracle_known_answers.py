"""Pins the oracle against every PG-free known answer the reference holds for the hot path.

Each test names the reference test / golden file it restates (file:line under /root/reference).
"""
import math

import numpy as np
import pytest

import oracle_ffi as O


def f32(*x):
    return np.array(x, dtype=np.float32)


def half(*x):
    L = O.lib()
    return np.array([L.orc_f32_to_half(float(v)) for v in x], dtype=np.uint16)


def bits(s):
    return np.packbits(np.array([int(c) for c in s], dtype=np.uint8))


# ---- src/types/vector.rs:956-1065 --------------------------------------------------------
def test_vector_distances():
    assert math.sqrt(O.distance(O.ELEM_F32, O.L2SQ, 2, f32(0, 0), f32(3, 4))) == 5.0          # :956-961
    assert O.distance(O.ELEM_F32, O.L2SQ, 3, f32(1, 2, 3), f32(1, 2, 3)) == 0.0               # :964-970
    assert O.distance(O.ELEM_F32, O.NEG_IP, 3, f32(1, 2, 3), f32(4, 5, 6)) == -32.0           # :973-980, :1045-1052
    assert O.distance(O.ELEM_F32, O.L1, 3, f32(1, 2, 3), f32(4, 6, 8)) == 12.0                # :1001-1008
    norm, _ = O.normalize(O.ELEM_F32, 2, f32(3, 4))                                           # :1019-1024
    assert norm == 5.0


def test_vector_cosine_via_normalised_ip():
    # cosine opclass = negative inner product over unit rows (vector.rs:852-856, build.rs:426-438)
    _, a = O.normalize(O.ELEM_F32, 3, f32(1, 2, 3))
    assert abs(1.0 + O.distance(O.ELEM_F32, O.NEG_IP, 3, a, a)) < 1e-6                        # :983-989
    _, x = O.normalize(O.ELEM_F32, 2, f32(1, 0))
    _, y = O.normalize(O.ELEM_F32, 2, f32(0, 1))
    assert 1.0 + O.distance(O.ELEM_F32, O.NEG_IP, 2, x, y) == 1.0                             # :992-998
    norm, z = O.normalize(O.ELEM_F32, 3, f32(0, 0, 0))
    assert norm == 0.0 and not z.any()


# ---- src/types/halfvec.rs:1083-1113, 1147-1184 ---------------------------------------------
def test_half_codec_known_answers():
    L = O.lib()
    rt = lambda v: L.orc_half_to_f32(L.orc_f32_to_half(v))
    assert rt(0.0) == 0.0 and rt(1.0) == 1.0 and rt(-1.0) == -1.0
    assert abs(rt(math.pi) - math.pi) < 0.002
    assert math.isinf(rt(float("inf"))) and math.isinf(rt(float("-inf")))
    assert math.isnan(rt(float("nan")))


def test_half_codec_matches_ieee_everywhere_it_should():
    # every half value decodes exactly like IEEE binary16 (numpy float16)
    L = O.lib()
    allh = np.arange(65536, dtype=np.uint16)
    ours = np.array([L.orc_half_to_f32(int(h)) for h in allh], dtype=np.float32)
    ref = allh.view(np.float16).astype(np.float32)
    nan = np.isnan(ref)       # signalling-NaN payloads get quieted on the ctypes float path
    assert np.array_equal(np.isnan(ours), nan)
    assert np.array_equal(ours.view(np.uint32)[~nan], ref.view(np.uint32)[~nan])
    # encode: round-to-nearest-even for normal results and in-range subnormals
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * 3,
                        (rng.standard_normal(5000) * 1e-5).astype(np.float32),
                        np.array([65504.0, 65519.9, 65520.0, 1e6, 6.1e-5, 5.96e-8], np.float32)])
    enc = np.array([L.orc_f32_to_half(float(v)) for v in x], dtype=np.uint16)
    ieee = x.astype(np.float16).view(np.uint16)
    # the reference flushes |x| < 2^-24 (f32 exponent < 103) to zero where IEEE may round up to
    # the smallest subnormal (halfvec.rs:111-113): exclude that sliver
    ok = np.abs(x) >= 2.0 ** -24
    assert np.array_equal(enc[ok], ieee[ok])
    tiny = np.float32(2.0 ** -24 * 0.75)
    assert L.orc_f32_to_half(float(tiny)) == 0 and np.float16(tiny).view(np.uint16) == 1


def test_halfvec_distances():
    assert math.sqrt(O.distance(O.ELEM_F16, O.L2SQ, 2, half(0, 0), half(3, 4))) == 5.0         # :1147-1160
    assert O.distance(O.ELEM_F16, O.NEG_IP, 3, half(1, 2, 3), half(4, 5, 6)) == -32.0          # :1163-1168
    assert O.distance(O.ELEM_F16, O.L1, 3, half(1, 2, 3), half(4, 6, 8)) == 12.0               # :1179-1184
    _, x = O.normalize(O.ELEM_F16, 2, half(1, 0))
    _, y = O.normalize(O.ELEM_F16, 2, half(0, 1))
    assert 1.0 + O.distance(O.ELEM_F16, O.NEG_IP, 2, x, y) == 1.0                              # :1171-1176


# ---- src/types/bitvec.rs:248-311 -------------------------------------------------------------
def test_bit_distances():
    assert O.distance(O.ELEM_BIT, O.HAMMING, 8, bits("10101010"), bits("10100000")) == 2.0    # :248-254
    assert O.distance(O.ELEM_BIT, O.HAMMING, 8, bits("11110000"), bits("11110000")) == 0.0    # :257-263
    assert O.distance(O.ELEM_BIT, O.JACCARD, 8, bits("11000000"), bits("10100000")) == 1.0 - 1.0 / 3.0  # :266-276
    assert O.distance(O.ELEM_BIT, O.JACCARD, 8, bits("11110000"), bits("11110000")) == 0.0    # :279-285
    assert O.distance(O.ELEM_BIT, O.JACCARD, 8, bits("11000000"), bits("00110000")) == 1.0    # :288-295


# ---- src/hnsw_constants.rs:136-164, src/types/hnsw.rs:337-349 ----------------------------------
def test_levels_and_constants():
    L = O.lib()
    assert L.orc_max_level(16) == 82 and L.orc_max_level(32) == 40
    assert L.orc_level_from_uniform(1.0, 16) == 0
    assert L.orc_level_from_uniform(1.0 / 16.0 - 1e-9, 16) == 1
    assert L.orc_level_from_uniform(0.0, 16) == 82                   # clamped by max(r, MIN_POSITIVE) + cap
    with pytest.raises(ValueError):
        O.Oracle(O.ELEM_F32, O.L2SQ, 3, m=1)                          # hnsw_vector.out:150-151
    with pytest.raises(ValueError):
        O.Oracle(O.ELEM_F32, O.L2SQ, 3, m=101)
    with pytest.raises(ValueError):
        O.Oracle(O.ELEM_F32, O.L2SQ, 3, ef_construction=3)
    with pytest.raises(ValueError):
        O.Oracle(O.ELEM_F32, O.L2SQ, 3, ef_construction=1001)
    with pytest.raises(ValueError):
        O.Oracle(O.ELEM_F32, O.L2SQ, 3, m=16, ef_construction=31)     # ef_construction >= 2*m


# ---- src/graph/mod.rs:538-778 (the only PG-free tests of the path) ----------------------------
def chain_graph(tie):
    # 5 elements at [0],[1],[2],[3],[4], chained 0-1-2-3-4 at layer 0 (graph/mod.rs:543-561)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 1, m=4, ef_construction=16, tie_mode=tie)
    rows = np.arange(5, dtype=np.float32)[:, None]
    adj0 = -np.ones((5, 8), np.int32)
    lists = {0: [1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3]}
    for i, l in lists.items():
        adj0[i, :len(l)] = l
    o.import_graph(dict(rows=rows, level=np.zeros(5, np.uint8), adj0=adj0,
                        upper_base=-np.ones(5, np.int32), adj_up=np.zeros((0, 4), np.int32), entry=0))
    return o


@pytest.mark.parametrize("tie", [O.TIE_RUST_HEAP, O.TIE_TOTAL_ORDER])
def test_search_layer_basic(tie):
    o = chain_graph(tie)
    q = f32(3.5)
    ep_d = np.float32(O.distance(O.ELEM_F32, O.L2SQ, 1, q, f32(0.0)))
    ids, ds = o.search_layer(q, [0], [ep_d], ef=2, layer=0)
    assert len(ids) == 2 and set(ids.tolist()) == {3, 4}              # :580-583
    assert ds[0] == 0.25 and ds[1] == 0.25


@pytest.mark.parametrize("tie", [O.TIE_RUST_HEAP, O.TIE_TOTAL_ORDER])
def test_select_neighbors_fits_and_prunes(tie):
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 2, m=4, ef_construction=16, tie_mode=tie)
    rows = np.array([[0, 0], [1, 0], [1.1, 0], [5, 0]], np.float32)
    o.import_graph(dict(rows=rows, level=np.zeros(4, np.uint8), adj0=-np.ones((4, 8), np.int32),
                        upper_base=-np.ones(4, np.int32), adj_up=np.zeros((0, 4), np.int32), entry=0))
    ids, _ = o.select_neighbors([1, 2], [1.0, 2.0], 4)                 # :587-611 fits -> all
    assert ids.tolist() == [1, 2]
    ids, _ = o.select_neighbors([1, 2, 3], [1.0, 1.21, 25.0], 2)       # :614-646
    assert len(ids) == 2 and ids[0] == 1
    # idx 2 (0.01 from idx 1) and idx 3 (16 from idx 1, <= 25) are both discarded; the free slot is
    # refilled from `discarded` in order (graph/mod.rs:300-305), so idx 2 comes back
    assert ids.tolist() == [1, 2]
    ids, _ = o.select_neighbors([1, 2, 3], [1.0, 1.21, 25.0], 3 - 1 + 1)  # lm=3 >= len -> returned as is
    assert ids.tolist() == [1, 2, 3]


@pytest.mark.parametrize("tie", [O.TIE_RUST_HEAP, O.TIE_TOTAL_ORDER])
def test_find_neighbors_and_backlinks(tie):
    # graph/mod.rs:649-726 via the build loop (levels all 0)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 1, m=4, ef_construction=16, tie_mode=tie)
    o.build(np.array([[0.0], [1.0]], np.float32), [0, 0])
    assert o.neighbors(1, 0).tolist() == [0]                           # :673-675
    assert o.neighbors(0, 0).tolist() == [1]                           # :706-708
    o.build(np.array([[2.0]], np.float32), [0], base_tid=2)
    n1 = o.neighbors(1, 0).tolist()
    assert 0 in n1 and 2 in n1                                         # :722-725


@pytest.mark.parametrize("tie", [O.TIE_RUST_HEAP, O.TIE_TOTAL_ORDER])
def test_multi_layer(tie):
    # graph/mod.rs:729-778: elem0 level 2 @[0], elem1 level 0 @[3], elem2 level 1 @[1]; m=2, efc=8
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 1, m=2, ef_construction=8, tie_mode=tie)
    o.build(np.array([[0.0], [3.0], [1.0]], np.float32), [2, 0, 1])
    assert len(o.neighbors(1, 0)) > 0                                  # :761
    assert len(o.neighbors(2, 0)) > 0 and len(o.neighbors(2, 1)) > 0   # :776-777
    assert o.entry == 0                                                # entry changes only on strictly greater level


# ---- tests/pg_regress/expected/*.out -----------------------------------------------------------
def _order(elem, metric, dim, rows, query, cosine=False, ef=40, iterative=O.ITER_OFF, tie=O.TIE_RUST_HEAP):
    o = O.Oracle(elem, metric, dim, tie_mode=tie)
    kept = []
    for r in rows:
        if cosine:
            norm, r2 = O.normalize(elem, dim, r)
            if norm == 0.0:
                continue        # build.rs:433-435
            r = r2
        kept.append(r)
    o.build(np.stack(kept), np.zeros(len(kept), np.int32))
    if cosine:
        _, query = O.normalize(elem, dim, query)
    tids, eids, _, _ = o.scan_stream(query, ef_search=ef, iterative=iterative)
    g = o.export_graph()
    return [tuple(float(v) if elem == O.ELEM_F32 else int(v) for v in g["rows"][e]) for e in eids], kept


@pytest.mark.parametrize("tie", [O.TIE_RUST_HEAP, O.TIE_TOTAL_ORDER])
def test_regress_vector_orders(tie):
    rows = [f32(0, 0, 0), f32(1, 2, 3), f32(1, 1, 1), f32(1, 2, 4)]   # NULL row skipped (build.rs:411)
    q = f32(3, 3, 3)
    got, _ = _order(O.ELEM_F32, O.L2SQ, 3, rows, q, tie=tie)           # hnsw_vector.out:7-14
    assert got == [(1, 2, 3), (1, 2, 4), (1, 1, 1), (0, 0, 0)]
    got, _ = _order(O.ELEM_F32, O.NEG_IP, 3, rows, q, tie=tie)         # :40-47
    assert got == [(1, 2, 4), (1, 2, 3), (1, 1, 1), (0, 0, 0)]
    got, kept = _order(O.ELEM_F32, O.NEG_IP, 3, rows, q, cosine=True, tie=tie)  # :61-67 (zero vector not indexed)
    assert len(got) == 3
    want = [f32(1, 1, 1), f32(1, 2, 3), f32(1, 2, 4)]
    for g, w in zip(got, want):
        assert np.allclose(g, O.normalize(O.ELEM_F32, 3, w)[1])
    got, _ = _order(O.ELEM_F32, O.L1, 3, rows, q, tie=tie)             # :87-94
    assert got == [(1, 2, 3), (1, 2, 4), (1, 1, 1), (0, 0, 0)]


@pytest.mark.parametrize("mode", [O.ITER_STRICT, O.ITER_RELAXED])
@pytest.mark.parametrize("tie", [O.TIE_RUST_HEAP, O.TIE_TOTAL_ORDER])
def test_regress_vector_iterative(mode, tie):
    rows = [f32(0, 0, 0), f32(1, 2, 3), f32(1, 1, 1)]
    got, _ = _order(O.ELEM_F32, O.L2SQ, 3, rows, f32(3, 3, 3), ef=1, iterative=mode, tie=tie)  # :107-124
    assert got == [(1, 2, 3), (1, 1, 1), (0, 0, 0)]
    # without iterative scan ef_search=1 yields a single row
    got, _ = _order(O.ELEM_F32, O.L2SQ, 3, rows, f32(3, 3, 3), ef=1, tie=tie)
    assert got == [(1, 2, 3)]


def test_regress_halfvec_orders():
    rows = [half(0, 0, 0), half(1, 2, 3), half(1, 1, 1), half(1, 2, 4)]
    q = half(3, 3, 3)
    h = lambda *x: tuple(int(v) for v in half(*x))
    got, _ = _order(O.ELEM_F16, O.L2SQ, 3, rows, q)                    # hnsw_halfvec.out:7-14
    assert got == [h(1, 2, 3), h(1, 2, 4), h(1, 1, 1), h(0, 0, 0)]
    got, _ = _order(O.ELEM_F16, O.NEG_IP, 3, rows, q)                  # :40-47
    assert got == [h(1, 2, 4), h(1, 2, 3), h(1, 1, 1), h(0, 0, 0)]
    got, _ = _order(O.ELEM_F16, O.NEG_IP, 3, rows, q, cosine=True)     # :61-67
    assert len(got) == 3
    got, _ = _order(O.ELEM_F16, O.L1, 3, rows, q)                      # :87-94
    assert got == [h(1, 2, 3), h(1, 2, 4), h(1, 1, 1), h(0, 0, 0)]


def test_regress_bit_orders():
    b = lambda s: tuple(int(v) for v in bits(s))
    rows = [bits("000"), bits("100"), bits("111"), bits("110")]
    got, _ = _order(O.ELEM_BIT, O.HAMMING, 3, rows, bits("111"))       # hnsw_bit.out:7-14
    assert got == [b("111"), b("110"), b("100"), b("000")]
    rows = [bits("0000"), bits("1100"), bits("1111"), bits("1110")]
    got, _ = _order(O.ELEM_BIT, O.JACCARD, 4, rows, bits("1111"))      # :28-35
    assert got == [b("1111"), b("1110"), b("1100"), b("0000")]


# ---- src/index/scan.rs:1078-1119 (10x10 grid) ----------------------------------------------------
@pytest.mark.parametrize("tie", [O.TIE_RUST_HEAP, O.TIE_TOTAL_ORDER])
def test_grid_scan(tie):
    rows = np.array([[i, j] for i in range(10) for j in range(10)], np.float32)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 2, tie_mode=tie)
    rng = np.random.default_rng(5)
    o.build(rows, O.levels_from_uniform(rng.random(100), 16))
    ids, ds, cnt = o.search_batch(f32(0, 0), ef_search=40, k=5)
    assert cnt[0] == 5 and ids[0, 0] == 0 and ds[0, 0] == 0.0
    assert set(ids[0, :4].tolist()) == {0, 1, 10, 11}
    assert ids[0, 4] in (2, 20) and ds[0, 4] == 4.0


# ---- tests/t/015_hnsw_vector_duplicates.pl:24-32 ---------------------------------------------------
def test_duplicates_merge_heap_tids():
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    rows = np.tile(f32(1, 1, 1), (25, 1))
    n = o.build(rows, np.zeros(25, np.int32))
    # every element holds at most 10 heap TIDs (hnsw_constants.rs:85); 25 copies -> 3 elements
    assert n == 3
    assert sorted(len(o.heap_tids(i)) for i in range(3)) == [5, 10, 10]
    tids, eids, _, _ = o.scan_stream(f32(1, 1, 1), ef_search=1)
    assert len(tids) == 10 and len(set(eids.tolist())) == 1            # exactly 10 TIDs at ef_search=1
    # TIDs are emitted from the back of the element's list (scan.rs:798)
    assert tids.tolist() == o.heap_tids(int(eids[0]))[::-1].tolist()


# ---- src/index/insert.rs:1492-1616 (pg_tests of aminsert, restated on the oracle) -------------------
def test_insert_after_build_is_found():
    # test_hnsw_insert_basic: index three rows, INSERT [0.5,0.5,0], nearest to it is itself
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    o.build(np.stack([f32(1, 0, 0), f32(0, 1, 0), f32(0, 0, 1)]), np.zeros(3, np.int32))
    assert o.insert(f32(0.5, 0.5, 0)[None], np.zeros(1, np.int32), base_tid=3) == 4
    ids, ds, _ = o.search_batch(f32(0.5, 0.5, 0)[None], k=1)
    assert ids[0, 0] == 3 and ds[0, 0] == 0.0


def test_insert_into_empty_index():
    # test_hnsw_insert_into_empty_index: every row goes through aminsert; nearest to [1,0,0] is [1,0,0]
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    for i, r in enumerate((f32(1, 0, 0), f32(0, 1, 0), f32(0, 0, 1))):
        o.insert(r[None], np.zeros(1, np.int32), base_tid=i)
    assert o.n == 3 and o.entry == 0
    ids, _, _ = o.search_batch(f32(1, 0, 0)[None], k=1)
    assert ids[0, 0] == 0
    # back-links went into free slots (insert.rs:559-562): everyone is linked to everyone
    for i in range(3):
        assert sorted(o.neighbors(i, 0).tolist()) == sorted(set(range(3)) - {i})


def test_insert_cosine_and_duplicates():
    # test_hnsw_insert_cosine
    o = O.Oracle(O.ELEM_F32, O.NEG_IP, 3)
    o.build(np.stack([f32(1, 0, 0), f32(0, 1, 0)]), np.zeros(2, np.int32))
    o.insert(f32(0, 0, 1)[None], np.zeros(1, np.int32), base_tid=2)
    ids, _, _ = o.search_batch(f32(0, 0, 1)[None], k=1)
    assert ids[0, 0] == 2
    # the duplicate walk stops at the first neighbour whose distance is not 0.0 (insert.rs:1189-1191): under -dot an
    # identical unit row is at distance -1, so cosine / inner-product indexes never merge duplicates
    o.insert(np.tile(f32(0, 0, 1), (3, 1)), np.zeros(3, np.int32), base_tid=3)
    assert o.n == 6
    # L2: a duplicate row adds a heap TID to an element holding 1..9 of them (insert.rs:1136-1170)
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 3)
    o.build(np.stack([f32(1, 0, 0), f32(0, 1, 0)]), np.zeros(2, np.int32))
    o.insert(np.tile(f32(0, 0, 1), (13, 1)), np.zeros(13, np.int32), base_tid=2)
    assert o.n == 4 and sorted(len(o.heap_tids(i)) for i in (2, 3)) == [3, 10]


def test_insert_replaces_one_slot_of_a_full_list():
    # get_update_index (insert.rs:500-738) on a full list: m = 2 => lm = 4 at layer 0.  Node 0 at the origin holds
    # 1, 2, 3, 4 at distances 1, 4, 9, 16 on a line; 5 arrives at distance 2.25 on the other side.  Sorted: 1, 5(new),
    # 2, 3, 4; 2 is pruned by 1 (d(2,1)=1 <= 4), 3 by 1, 4 by 1; kept = {1, new} + pruned 2, 3 => slot of 4 is overwritten.
    o = O.Oracle(O.ELEM_F32, O.L2SQ, 1, m=2, ef_construction=8)
    o.insert(np.array([[0.0], [1.0], [2.0], [3.0], [4.0]], np.float32), np.zeros(5, np.int32))
    assert o.neighbors(0, 0).tolist() == [1, 2, 3, 4]
    o.insert(np.array([[-1.5]], np.float32), np.zeros(1, np.int32), base_tid=5)
    assert o.neighbors(0, 0).tolist() == [1, 2, 3, 5]             # slot order kept, only the last slot changed
    assert o.neighbors(5, 0).tolist() == [0, 1, 2, 3]             # nearest lm of W, no heuristic (insert.rs:1110-1116)
