"""tools/pgvrx_relation_check.cpp: the exporter pointed at a relation FILE, the way a maintainer would validate it
against an index written by the real extension.  Here the files are written by the oracle's
independent page writer: a fresh CREATE INDEX (byte-identical to the flush layout), the same relation split into
segment files, and a relation a vacuum has worked on."""
import json
import os
import shutil
import subprocess

import pytest

from helpers import make_rows, oracle_index, vacuumed_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "pgvector_rx_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    out = str(tmp_path_factory.mktemp("relcheck") / "pgvrx_relation_check")
    subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tools", "pgvrx_relation_check.cpp"), "-L", LIBDIR, "-lpgvrx_b200",
                    "-Wl,-rpath," + LIBDIR, "-o", out], check=True, capture_output=True, text=True)
    return out


def run(exe, *args):
    r = subprocess.run([exe] + list(args), capture_output=True, text=True, timeout=120)
    return r.returncode, (json.loads(r.stdout) if r.stdout.strip().startswith("{") else None), r.stderr


def test_fresh_index_file_and_segments(exe, tmp_path):
    o = oracle_index("vector_l2_ops", 16, make_rows("f32", 3000, 16, 1), m=8, efc=32)
    pages = o.write_pages()
    f = str(tmp_path / "24601")
    pages.tofile(f)
    rc, out, err = run(exe, f)
    assert rc == 0 and out["ok"] and out["identical_to_flush_layout"], err
    assert out["elements"] == o.n and out["m"] == 8 and out["ef_construction"] == 32 and out["dims"] == 16
    assert out["entry"] == o.entry and out["deleted"] == 0 and out["heap_tids"] == 3000
    assert out["lists_not_dense_prefix"] == 0 and out["self_links"] == 0 and out["links_to_deleted"] == 0
    # a relation above 1 GiB comes as file, file.1, ...: same answer when the pages are split across segments
    half = (len(pages) // 8192 // 2) * 8192
    g = str(tmp_path / "24602")
    pages[:half].tofile(g)
    pages[half:].tofile(g + ".1")
    rc2, out2, _ = run(exe, g)
    out2["relation"] = out["relation"]
    assert rc2 == 0 and out2 == out
    # the wrong element type is refused by the exporter's size check, with its message
    rc3, _, err3 = run(exe, f, "--elem", "f16")
    assert rc3 == 3 and "does not match the index dimensions" in err3
    assert run(exe, str(tmp_path / "missing"))[0] == 2


def test_vacuumed_index_file(exe, tmp_path):
    o, deleted, no_tid, stale = vacuumed_oracle("vector_l2_ops", "f32", 3000, 16, 8, 32)
    f = str(tmp_path / "24603")
    o.write_pages().tofile(f)
    rc, out, err = run(exe, f)
    assert rc == 0 and out["ok"], err
    assert out["deleted"] == len(deleted) and out["without_heap_tids"] == len(deleted) + len(no_tid)
    assert out["links_to_deleted"] > 0                       # no repair was run on this fixture: live lists point at deleted elements
    assert out["elements_without_neighbours"] >= len(set(stale.tolist()) - set(deleted.tolist()))   # stale tuples read as empty
    assert not out["identical_to_flush_layout"]


def test_scan_option_needs_a_device(exe, tmp_path):
    import pgvector_rx_b200 as P
    if P.lib().pgvrx_device_count() > 0:
        pytest.skip("a device is present")
    o = oracle_index("vector_l2_ops", 4, make_rows("f32", 200, 4, 2), m=4, efc=16)
    f = str(tmp_path / "24604")
    o.write_pages().tofile(f)
    rc, out, err = run(exe, f, "--scan", "50")
    assert rc == 3 and out is None and "no CPU path" in err
