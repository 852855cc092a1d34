"""The Julia wrapper cannot be executed in this image (no julia binary), and the one class of bug a never-run `ccall` binding is
certain to grow is a struct that no longer matches the header.  This test parses `struct SabcConfig` out of SABCB200.jl, lays it
out by the C ABI rules Julia uses for isbits structs (natural alignment, declaration order) and compares every field's offset and
size -- and the struct size -- with offsetof()/sizeof() of include/sabc_b200.h as compiled by gcc.  The ctypes mirror
(_lib.Config, _lib.Timing) is checked against the same numbers."""
import ctypes as C
import os
import re
import subprocess

import sabc_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "simulatedannealingabc.jl_b200", "julia", "SABCB200.jl")
SIZES = {"Int64": 8, "UInt64": 8, "Float64": 8, "Int32": 4, "UInt32": 4, "Cstring": 8, "Cint": 4}


def julia_struct(name):
    text = open(JL).read()
    body = re.search(r"struct %s\n(.*?)\nend" % name, text, re.S).group(1)
    fields = []
    for line in body.splitlines():
        line = line.split("#")[0]
        for item in line.split(";"):
            item = item.strip()
            if not item:
                continue
            fname, ftype = [x.strip() for x in item.split("::")]
            if ftype.startswith("Ptr{"):
                size, align = 8, 8
            elif ftype.startswith("NTuple{"):
                n, t = re.match(r"NTuple\{(\d+),\s*(\w+)\}", ftype).groups()
                size, align = int(n) * SIZES[t], SIZES[t]
            else:
                size = align = SIZES[ftype]
            fields.append((fname, size, align))
    off, out, amax = 0, {}, 1
    for fname, size, align in fields:
        off = (off + align - 1) // align * align
        out[fname] = (off, size)
        off += size
        amax = max(amax, align)
    return [f[0] for f in fields], out, (off + amax - 1) // amax * amax


def c_layout(tmp_path, struct, fields):
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sabc_b200.h"', 'int main(void) {']
    for f in fields:
        lines.append(f'  printf("{f} %zu %zu\\n", offsetof({struct}, {f}), sizeof((({struct}*)0)->{f}));')
    lines += [f'  printf("__size__ %zu 0\\n", sizeof({struct}));', '  return 0; }']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    return {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out.splitlines()}


def header_fields(struct):
    text = open(os.path.join(ROOT, "include", "sabc_b200.h")).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.sub(r"\[.*\]", "", part.strip().split()[-1].lstrip("*")))
    return names


def test_julia_config_matches_the_header(tmp_path):
    order, jl, jl_size = julia_struct("SabcConfig")
    hdr = header_fields("sabc_config")
    assert order == hdr, f"field order differs:\n julia  {order}\n header {hdr}"
    c = c_layout(tmp_path, "sabc_config", hdr)
    for f in hdr:
        assert jl[f] == c[f], f"{f}: julia (offset, size) {jl[f]} vs C {c[f]}"
    assert jl_size == c["__size__"][0]


def test_ctypes_mirrors_match_the_header(tmp_path):
    for struct, mirror in (("sabc_config", sb._lib.Config), ("sabc_timing", sb._lib.Timing)):
        hdr = header_fields(struct)
        assert [f[0] for f in mirror._fields_] == hdr
        c = c_layout(tmp_path, struct, hdr)
        for f in hdr:
            d = getattr(mirror, f)
            assert (d.offset, d.size) == c[f], (struct, f)
        assert C.sizeof(mirror) == c["__size__"][0]


def _split_top(s):
    """split a comma list at nesting depth 0 ({...} and (...) may contain commas)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        if ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_ccall_signatures_match_the_header():
    """every `ccall((:sabc_x, libsabc), Ret, (T1, T2, ...), ...)` of SABCB200.jl against the prototype of sabc_x in the header: same
    number of arguments, and argument by argument the same class (pointer / 64-bit integer / 32-bit integer / double)."""
    hdr = open(os.path.join(ROOT, "include", "sabc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(sabc_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if args == ["void"] else args

    def c_class(a):
        if "*" in a or "[" in a:
            return "ptr"
        t = a.split()
        ty = " ".join(t[:-1]) if len(t) > 1 else t[0]
        return {"int64_t": "i64", "uint64_t": "i64", "int32_t": "i32", "uint32_t": "i32", "int": "i32", "double": "f64"}[ty.replace("const ", "")]

    def jl_class(t):
        if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
            return "ptr"
        return {"Int64": "i64", "UInt64": "i64", "Int32": "i32", "UInt32": "i32", "Cint": "i32", "Float64": "f64"}[t]

    text = open(JL).read()
    seen = set()
    for m in re.finditer(r"ccall\(\(:(sabc_\w+), libsabc\),\s*(\w+),\s*\(", text):
        name = m.group(1)
        depth, i = 1, m.end()
        while depth:                                    # the matching parenthesis of the argument-type tuple
            depth += {"(": 1, ")": -1}.get(text[i], 0); i += 1
        types = [t for t in _split_top(text[m.end():i - 1]) if t]
        assert name in protos, f"{name} is not declared in the header"
        want = protos[name]
        assert len(types) == len(want), f"{name}: julia passes {len(types)} arguments, the header declares {len(want)}"
        for k, (jt, ca) in enumerate(zip(types, want)):
            assert jl_class(jt) == c_class(ca), f"{name} argument {k + 1}: julia {jt} vs C `{ca}`"
        seen.add(name)
    assert {"sabc_create", "sabc_init", "sabc_update", "sabc_update_host", "sabc_set_tuning", "sabc_get_population", "sabc_get_state",
            "sabc_get_history", "sabc_history_len", "sabc_destroy", "sabc_last_error"} <= seen
