"""Robustness of the page exporter (a19): a relation file is untrusted input -- torn pages, stale headers, a wrong
relation.  Mutated copies of oracle-written relations (random byte flips, corrupt page headers and line pointers,
corrupt meta fields, truncation) go through pgvrx_pages_inspect / _export / _element_flags compiled with
AddressSanitizer + UBSan: every one must be accepted or rejected with a status code, never read or write out of bounds."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import make_rows, oracle_index, vacuumed_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mutate(pages, it, rng):
    bad = pages.copy()
    npg = len(bad) // 8192
    mode = it % 5
    if mode == 0:                                     # random byte flips anywhere
        for _ in range(int(rng.integers(1, 20))):
            bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
    elif mode == 1:                                   # pd_lower / pd_upper / pd_special
        pg = int(rng.integers(0, npg))
        for off in (12, 13, 14, 15, 16, 17):
            if rng.random() < 0.6:
                bad[pg * 8192 + off] = int(rng.integers(0, 256))
    elif mode == 2:                                   # line pointers
        pg = int(rng.integers(1, npg))
        for _ in range(6):
            bad[pg * 8192 + 24 + int(rng.integers(0, 200))] = int(rng.integers(0, 256))
    elif mode == 3:                                   # truncated relation
        bad = bad[: 8192 * int(rng.integers(1, npg))]
    else:                                             # meta page fields
        for _ in range(3):
            bad[24 + int(rng.integers(0, 28))] = int(rng.integers(0, 256))
    return bad


def test_exporter_survives_corrupt_relations(tmp_path):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "pages_fuzz")
    r = subprocess.run([cxx, "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-pthread",
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "fuzz", "pages_fuzz_main.cpp"),
                        os.path.join(ROOT, "pgvector_rx_b200", "csrc", "page_export.cpp"), "-o", exe],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-200:])
    o = oracle_index("vector_l2_ops", 8, make_rows("f32", 300, 8, 1), m=4, efc=16)
    o2, *_ = vacuumed_oracle("vector_l2_ops", "f32", 300, 8, 4, 16)
    srcs = [o.write_pages(), o2.write_pages()]
    rng = np.random.default_rng(7)
    files = []
    for it in range(600):
        f = str(tmp_path / ("in_%04d.bin" % it))
        mutate(srcs[it % 2], it, rng).tofile(f)
        files.append(f)
    good = str(tmp_path / "good.bin")
    srcs[0].tofile(good)
    r = subprocess.run([exe, good] + files, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert "asan fuzz:" in r.stdout
    accepted = int(r.stdout.split("asan fuzz:")[1].split("accepted")[0])
    assert 1 <= accepted < len(files)                 # the intact relation is accepted, most mutants are not


def test_writer_survives_corrupt_graph_arrays(tmp_path):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "writer_fuzz")
    r = subprocess.run([cxx, "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-pthread",
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "fuzz", "writer_fuzz_main.cpp"),
                        os.path.join(ROOT, "pgvector_rx_b200", "csrc", "page_write.cpp"), "-o", exe],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-200:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert "writer fuzz:" in r.stdout
    written = int(r.stdout.split("writer fuzz:")[1].split("written")[0])
    assert written > 100
