"""The boundary from plain C: include/pgvrx_b200.h must compile as C99 (-pedantic -Werror) and link against the
library the way a cgo / pgrx-sys / JNI binding would (examples/abi_smoke.c).  CPU: host-only entry points and the
loud failure of the compute entry points without a device; GPU: build, insert, scan and exact flat kNN through C."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "pgvector_rx_b200")


def compile_smoke(tmp_path):
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "abi_smoke")
    subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "abi_smoke.c"), "-L", LIBDIR, "-lpgvrx_b200",
                    "-Wl,-rpath," + LIBDIR, "-o", exe], check=True, capture_output=True, text=True)
    return exe


def test_header_is_c99_and_host_entry_points_work(tmp_path):
    exe = compile_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "host-only checks passed" in r.stdout


@pytest.mark.gpu
def test_c_caller_builds_inserts_and_scans(tmp_path):
    exe = compile_smoke(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "gpu checks passed" in r.stdout
