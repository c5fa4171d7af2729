"""CPU-side checks of the product library: it loads, exports every declared symbol, validates
options like the reference does, and its host-side opclass support is bit-identical to the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_ffi as O
import pgvector_rx_b200 as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pgvrx_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pgvrx_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = C.CDLL(P.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "libpgvrx_b200.so does not export %s" % name
    assert sorted(P.EXPORTS) == declared
    assert P.lib().pgvrx_abi_version() == 1


def test_option_validation_matches_reference_messages():
    # tests/pg_regress/expected/hnsw_vector.out:150-159
    # tests/pg_regress/expected/hnsw_vector.out:150-159, the ERROR lines verbatim
    for kw, text in ((dict(m=1), 'value 1 out of bounds for option "m"'), (dict(m=101), 'value 101 out of bounds for option "m"'),
                     (dict(ef_construction=3), 'value 3 out of bounds for option "ef_construction"'),
                     (dict(ef_construction=1001), 'value 1001 out of bounds for option "ef_construction"'),
                     (dict(m=16, ef_construction=31), "ef_construction must be greater than or equal to 2 * m")):
        with pytest.raises(P.PgvrxError) as e:
            P.HnswIndex("vector_l2_ops", 3, **kw)
        assert e.value.code == P.ERR_INVALID and e.value.message == text, str(e.value)
    with pytest.raises(P.PgvrxError) as e:
        P.HnswIndex("bit_hamming_ops", 64001)              # hnsw_bit.out:47-50
    assert e.value.message == "column cannot have more than 64000 dimensions for hnsw index"
    with pytest.raises(P.PgvrxError):
        P.HnswIndex("vector_l2_ops", 2001)
    with pytest.raises(P.PgvrxError):
        P.HnswIndex("varbit_hamming_ops", 3)


def test_no_cpu_fallback_without_device():
    if P.lib().pgvrx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(P.PgvrxError) as e:
        P.HnswIndex("vector_l2_ops", 8)
    assert e.value.code == P.ERR_CUDA and "no CPU path" in str(e.value)


def test_levels_match_oracle():
    L, OL = P.lib(), O.lib()
    rng = np.random.default_rng(0)
    for m in (2, 5, 16, 32, 100):
        assert L.pgvrx_max_level(m) == OL.orc_max_level(m)
        for r in np.concatenate([rng.random(2000), [0.0, 1.0, 1e-300, 1.0 / m, 1.0 / m - 1e-12]]):
            assert L.pgvrx_level_from_uniform(float(r), m) == OL.orc_level_from_uniform(float(r), m)


def test_normalize_bit_identical_to_oracle():
    rng = np.random.default_rng(1)
    for dim in (1, 3, 128, 769):
        x = (rng.standard_normal((50, dim)) * rng.choice([1e-6, 1.0, 300.0], size=(50, 1))).astype(np.float32)
        x[0] = 0
        for r in x:
            n1, u1 = P.normalize(P.ELEM_F32, dim, r)
            n2, u2 = O.normalize(O.ELEM_F32, dim, r)
            assert n1 == n2 and np.array_equal(u1.view(np.uint32), u2.view(np.uint32))
        h = x.astype(np.float16).view(np.uint16)
        for r in h:
            n1, u1 = P.normalize(P.ELEM_F16, dim, r)
            n2, u2 = O.normalize(O.ELEM_F16, dim, r)
            assert n1 == n2 and np.array_equal(u1, u2)
        # the batch entry point is the same arithmetic row by row
        for elem, rows in ((P.ELEM_F32, x), (P.ELEM_F16, h)):
            norms, unit = P.normalize_rows(elem, dim, rows)
            for i, r in enumerate(rows):
                n2, u2 = O.normalize(elem, dim, r)
                assert norms[i] == n2 and np.array_equal(unit[i].view(np.uint8), u2.view(np.uint8))
    # the reference flushes normalised components below 2^-24 to zero instead of rounding up
    r = np.array([1.0, 2.0 ** -24 * 0.75], np.float32).astype(np.float16).view(np.uint16)
    r = np.array([np.float16(1.0).view(np.uint16), np.float16(4.5e-8).view(np.uint16)], np.uint16)
    assert np.array_equal(P.normalize(P.ELEM_F16, 2, r)[1], O.normalize(O.ELEM_F16, 2, r)[1])
    with pytest.raises(P.PgvrxError):
        P.normalize(P.ELEM_BIT, 8, np.zeros(1, np.uint8))


def test_integration_doc_lists_every_export():
    # INTEGRATION.md shows the reference-side (pgrx) binding of the boundary: it has to name every entry point
    import os
    doc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    missing = [e for e in P.EXPORTS if e not in doc]
    assert not missing, missing


def test_normalize_rows_on_host_threads_is_the_row_by_row_result():
    # large batches are split over host threads (a CREATE INDEX worth of halfvec rows is seconds of software half
    # conversions): same bits as one pgvrx_normalize call per row, in place as well, first failing row's error
    rng = np.random.default_rng(21)
    x = rng.standard_normal((40_000, 96)).astype(np.float32)
    x[123] = 0.0                                                       # zero norm: comes out as a zero row
    norms, u = P.normalize_rows(P.ELEM_F32, 96, x)
    for i in (0, 123, 9_999, 39_999):
        nm, row = P.normalize(P.ELEM_F32, 96, x[i])
        assert nm == norms[i] and np.array_equal(row.view(np.uint32), u[i].view(np.uint32))
    assert norms[123] == 0.0 and not u[123].any()
    h = (rng.standard_normal((12_000, 256)) * 3).astype(np.float16).view(np.uint16)
    hn, hu = P.normalize_rows(P.ELEM_F16, 256, h)
    one = np.stack([P.normalize(P.ELEM_F16, 256, r)[1] for r in h[::997]])
    assert np.array_equal(hu[::997], one)
    inplace = h.copy()
    L = P.lib()
    assert L.pgvrx_normalize_rows(P.ELEM_F16, 256, P._p(inplace), P._p(inplace), len(inplace), None) == 0
    assert np.array_equal(inplace, hu)
    # an error from a worker thread reaches the caller's error state
    with pytest.raises(P.PgvrxError) as e:
        P.normalize_rows(P.ELEM_BIT, 256 * 8, np.zeros((8192, 256), np.uint8))
    assert "no norm support" in str(e.value)
