#!/usr/bin/env python
"""Generates rust/pgvrx-b200-sys/src/lib.rs -- the `extern "C"` mirror of include/pgvrx_b200.h a pgvector-rx maintainer
adds as a `-sys` crate -- from the header itself, so the two cannot drift (tests/test_rust_sys_cpu.py re-runs this and
compares with the committed file; there is no Rust toolchain in this image, so the crate ships as source).

    python tools/gen_rust_sys.py [--check]
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pgvrx_b200.h")
OUT = os.path.join(ROOT, "rust", "pgvrx-b200-sys", "src", "lib.rs")

SCALARS = {"int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "uint32_t": "u32", "uint8_t": "u8", "float": "f32",
           "double": "f64", "int": "c_int", "size_t": "usize", "char": "c_char", "void": "c_void"}
OPAQUE = {"pgvrx_index": "PgvrxIndex", "pgvrx_scan": "PgvrxScan", "pgvrx_group": "PgvrxGroup"}


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def camel(name):
    return "".join(p.capitalize() for p in name.split("_"))


def parse(text):
    """-> (defines, enums, structs, functions) in header order."""
    t = strip_comments(text)
    defines = [(m.group(1), m.group(2)) for m in re.finditer(r"^#define\s+(PGVRX_[A-Z0-9_]+)\s+(-?\d+)\s*$", t, flags=re.M)]
    enums = []
    for m in re.finditer(r"typedef\s+enum\s*\{(.*?)\}\s*(\w+)\s*;", t, flags=re.S):
        items = [(a.strip(), int(b)) for a, b in re.findall(r"(PGVRX_[A-Z0-9_]+)\s*=\s*(-?\d+)", m.group(1))]
        enums.append((m.group(2), items))
    structs = []
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", t, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            ty, names = decl.rsplit(" ", 1)[0], decl.rsplit(" ", 1)[1]
            # "uint64_t n_dist, n_expand0, n_expand_upper" style
            first, *rest = [x.strip() for x in decl.split(",")]
            ty, n0 = first.rsplit(" ", 1)
            for nm in [n0] + rest:
                fields.append((ty.strip(), nm.strip()))
        structs.append((m.group(2), fields))
    body = re.sub(r"typedef\s+(enum|struct)\s*\{.*?\}\s*\w+\s*;", " ", t, flags=re.S)
    body = re.sub(r"typedef\s+struct\s+\w+\s+\w+\s*;", " ", body)
    body = re.sub(r"^#.*$", " ", body, flags=re.M)
    body = body.replace('extern "C" {', " ").replace("}", " ")
    funcs = []
    for m in re.finditer(r"([\w\s\*]+?)\b(pgvrx_\w+)\s*\(([^;{]*?)\)\s*;", body, flags=re.S):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), " ".join(m.group(3).split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                params.append((mm.group(1).strip(), mm.group(2)))
        funcs.append((ret, name, params))
    return defines, enums, structs, funcs


def rust_type(c, struct_names):
    c = " ".join(c.replace("*", " * ").split())
    toks = c.split()
    const = toks[0] == "const"
    if const:
        toks = toks[1:]
    base = toks[0]
    if base in SCALARS:
        r = SCALARS[base]
    elif base in OPAQUE:
        r = OPAQUE[base]
    elif base in struct_names:
        r = camel(base)
    else:
        raise ValueError("unknown C type %r" % c)
    # every `*` points at what stands to its left; a `const` between two stars qualifies the pointer before it
    # (`const int32_t *const *` = pointer to const pointer to const int32_t)
    out, pointee_const = r, const
    for t in toks[1:]:
        if t == "*":
            out = ("*const " if pointee_const else "*mut ") + out
            pointee_const = False
        elif t == "const":
            pointee_const = True
        else:
            raise ValueError("unexpected token %r in C type %r" % (t, c))
    return out


RESERVED = {"in": "src", "type": "kind", "ref": "r", "box": "b", "out": "out", "as": "as_"}


def generate(text):
    defines, enums, structs, funcs = parse(text)
    struct_names = {n for n, _ in structs}
    L = []
    w = L.append
    w("//! pgvrx-b200-sys -- raw FFI bindings to libpgvrx_b200.so (the B200-native HNSW hot path for pgvector-rx).")
    w("//!")
    w("//! GENERATED from include/pgvrx_b200.h by tools/gen_rust_sys.py -- do not edit; re-run the generator instead.")
    w("//! The header is the contract (it cites the reference seam each entry point replaces); this file mirrors every")
    w("//! constant, enum value, struct field and prototype of it, in header order.  The safe wrapper crate is")
    w("//! rust/pgvrx-b200 (with the standalone bench binary), the modules pgvector-rx gains at its two seams")
    w("//! (src/index/scan.rs, src/index/build.rs) are rust/pgvector-rx-shim.")
    w("#![allow(non_camel_case_types, non_upper_case_globals)]")
    w("")
    w("use std::os::raw::{c_char, c_int, c_void};")
    w("")
    for name, val in defines:
        if name == "PGVRX_B200_H":
            continue
        w("pub const %s: i32 = %s;" % (name, val))
    w("")
    for ename, items in enums:
        w("// %s" % ename)
        for iname, val in items:
            w("pub const %s: i32 = %d;" % (iname, val))
        w("")
    for oname in OPAQUE.values():
        w("#[repr(C)]")
        w("pub struct %s {" % oname)
        w("    _private: [u8; 0],")
        w("}")
        w("")
    for sname, fields in structs:
        w("#[repr(C)]")
        w("#[derive(Debug, Clone, Copy)]")
        w("pub struct %s {" % camel(sname))
        for ty, nm in fields:
            arr = re.match(r"(\w+)\[(\d+)\]$", nm)            # fixed-size array member: `float phase_ms[4]`
            if arr:
                w("    pub %s: [%s; %s]," % (arr.group(1), rust_type(ty, struct_names), arr.group(2)))
            else:
                w("    pub %s: %s," % (nm, rust_type(ty, struct_names)))
        w("}")
        w("")
    w('#[link(name = "pgvrx_b200")]')
    w('extern "C" {')
    for ret, name, params in funcs:
        args = ", ".join("%s: %s" % (RESERVED.get(n, n), rust_type(t, struct_names)) for t, n in params)
        r = "" if ret == "void" else " -> " + rust_type(ret, struct_names)
        w("    pub fn %s(%s)%s;" % (name, args, r))
    w("}")
    w("")
    return "\n".join(L)


def main():
    text = open(HEADER).read()
    out = generate(text)
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != out:
            print("rust/pgvrx-b200-sys/src/lib.rs is stale: run python tools/gen_rust_sys.py", file=sys.stderr)
            sys.exit(1)
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(out)
    print(OUT)


if __name__ == "__main__":
    main()
