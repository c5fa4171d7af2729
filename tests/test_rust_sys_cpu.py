"""The Rust side ships as source (no Rust toolchain in this image): these checks keep it honest against the header.

* rust/pgvrx-b200-sys/src/lib.rs is exactly what tools/gen_rust_sys.py generates from include/pgvrx_b200.h;
* an independent, cruder reading of the header finds every prototype (same parameter count), every struct field and
  every constant mirrored there;
* every `ffi::` item the safe wrapper crate, its bench binary and the pgvector-rx shim use exists in the -sys crate, and
  every call passes as many arguments as the prototype takes.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pgvrx_b200.h")
SYS = os.path.join(ROOT, "rust", "pgvrx-b200-sys", "src", "lib.rs")
USERS = [os.path.join(ROOT, "rust", "pgvrx-b200", "src", "lib.rs"),
         os.path.join(ROOT, "rust", "pgvrx-b200", "src", "bin", "bench.rs"),
         os.path.join(ROOT, "rust", "pgvector-rx-shim", "src", "gpu_scan.rs"),
         os.path.join(ROOT, "rust", "pgvector-rx-shim", "src", "gpu_build.rs"),
         os.path.join(ROOT, "rust", "pgvector-rx-shim", "src", "guc.rs")]


def _no_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def _split_top(args):
    """Split on commas that are not inside (), [] or <>."""
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "([<":
            depth += 1
        elif ch in ")]>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def header_prototypes():
    t = _no_comments(open(HEADER).read())
    t = re.sub(r"typedef\s+(enum|struct)\s*\{.*?\}\s*\w+\s*;", " ", t, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(pgvrx_\w+)\s*\(([^;{]*?)\)\s*;", t, flags=re.S):
        args = " ".join(m.group(2).split())
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top(args))
    return protos


def sys_functions():
    t = open(SYS).read()
    return {m.group(1): (0 if not m.group(2).strip() else len(_split_top(m.group(2))))
            for m in re.finditer(r"pub fn (pgvrx_\w+)\((.*?)\)(?: -> [^;]+)?;", t)}


def test_sys_crate_is_what_the_generator_produces():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_prototype_is_mirrored_with_the_same_arity():
    protos, fns = header_prototypes(), sys_functions()
    assert len(protos) >= 50
    assert set(protos) == set(fns), (sorted(set(protos) - set(fns)), sorted(set(fns) - set(protos)))
    for name, n in protos.items():
        assert fns[name] == n, (name, n, fns[name])


def test_every_struct_field_and_constant_is_mirrored():
    t = _no_comments(open(HEADER).read())
    rs = open(SYS).read()
    n_fields = 0
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", t, flags=re.S):
        rname = "".join(p.capitalize() for p in m.group(2).split("_"))
        body = re.search(r"pub struct %s \{(.*?)\n\}" % rname, rs, flags=re.S)
        assert body, rname
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            first, *rest = [x.strip() for x in decl.split(",")]
            for nm in [first.rsplit(" ", 1)[1]] + rest:
                nm = nm.lstrip("*")
                arr = re.match(r"(\w+)\[(\d+)\]$", nm)
                if arr:
                    assert re.search(r"pub %s: \[\w+; %s\]," % (arr.group(1), arr.group(2)), body.group(1)), (rname, nm)
                else:
                    assert re.search(r"pub %s: " % re.escape(nm), body.group(1)), (rname, nm)
                n_fields += 1
    assert n_fields >= 60
    for m in re.finditer(r"^#define\s+(PGVRX_[A-Z0-9_]+)\s+(-?\d+)\s*$", t, flags=re.M):
        if m.group(1) == "PGVRX_B200_H":
            continue
        assert "pub const %s: i32 = %s;" % (m.group(1), m.group(2)) in rs, m.group(1)
    for m in re.finditer(r"(PGVRX_[A-Z0-9_]+)\s*=\s*(-?\d+)", t):
        assert "pub const %s: i32 = %s;" % (m.group(1), m.group(2)) in rs, m.group(1)


def _calls(text):
    """(name, n_args) for every `ffi::pgvrx_xxx(...)` call in a Rust source file."""
    out = []
    for m in re.finditer(r"ffi::(pgvrx_\w+)\s*\(", text):
        i, depth = m.end(), 1
        while depth:
            ch = text[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            i += 1
        inner = text[m.end():i - 1]
        inner = re.sub(r"//[^\n]*", "", inner)
        args = [a for a in _split_top(inner) if a.strip()]
        out.append((m.group(1), len(args)))
    return out


def test_wrapper_crates_only_use_what_the_sys_crate_declares():
    fns = sys_functions()
    rs = open(SYS).read()
    used = 0
    for path in USERS:
        assert os.path.exists(path), path
        text = open(path).read()
        for name, n in _calls(text):
            assert name in fns, (path, name)
            assert fns[name] == n, (os.path.basename(path), name, "call passes", n, "prototype takes", fns[name])
            used += 1
        for m in re.finditer(r"ffi::(PGVRX_[A-Z0-9_]+)", text):
            assert "pub const %s:" % m.group(1) in rs, (path, m.group(1))
        for m in re.finditer(r"ffi::(Pgvrx[A-Za-z]+)", text):
            assert "pub struct %s " % m.group(1) in rs, (path, m.group(1))
    assert used >= 30


def test_crate_manifests_exist_and_point_at_each_other():
    for d, needle in (("pgvrx-b200-sys", 'links = "pgvrx_b200"'), ("pgvrx-b200", "pgvrx-b200-sys"), ("pgvector-rx-shim", "pgvrx-b200")):
        p = os.path.join(ROOT, "rust", d, "Cargo.toml")
        assert os.path.exists(p), p
        assert needle in open(p).read(), (p, needle)
    assert "rustc-link-lib=dylib=pgvrx_b200" in open(os.path.join(ROOT, "rust", "pgvrx-b200-sys", "build.rs")).read()
