"""Static checks of the drop-in boundary (CPU): the C header, the ctypes binding and the Julia `ccall` shim must agree on
every entry point's name and argument count -- the shim cannot be executed here (the image has no Julia), so its signatures
are checked against include/gsa_cuda.h instead."""
import os
import re

import gsa_loader

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
g = gsa_loader.load()


def _split_args(s: str):
    """split a parenthesised argument list at top-level commas"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def header_declarations():
    src = open(os.path.join(ROOT, "include", "gsa_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"GSA_API\s+int\s+(\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = [] if args in ("", "void") else _split_args(args)
    return decls


def test_ctypes_signatures_match_header_arity():
    from importlib import import_module
    sigs = import_module("gsa_b200._lib").SIGNATURES
    decls = header_declarations()
    assert set(decls) == set(sigs)
    for name, args in decls.items():
        assert len(args) == len(sigs[name]), (name, args, sigs[name])


def test_julia_shim_ccalls_match_header():
    decls = header_declarations()
    src = open(os.path.join(ROOT, "julia", "GSACuda.jl")).read()
    calls = list(re.finditer(r"ccall\(\(:(\w+),\s*libgsa\),\s*Cint,\s*\(", src))
    assert len(calls) >= 10
    for m in calls:
        name = m.group(1)
        assert name in decls, f"the Julia shim calls {name}, which include/gsa_cuda.h does not declare"
        # the argument-type tuple starts at the '(' that ends the match
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types = [t for t in _split_args(src[m.end():i - 1]) if t]
        assert len(types) == len(decls[name]), (name, types, decls[name])
        for t, c in zip(types, decls[name]):                     # pointer-ness and 64-bit counts must line up
            is_ptr_c = "*" in c
            is_ptr_j = t.startswith(("Ptr", "Ref", "Cstring"))
            assert is_ptr_c == is_ptr_j, (name, t, c)
            if "int64_t" in c and not is_ptr_c:
                assert t == "Int64", (name, t, c)
            if "size_t" in c and not is_ptr_c:
                assert t == "Csize_t", (name, t, c)


def test_struct_layouts_match_header():
    """gsa_sobol_opts / gsa_sobol_result field order in the header, the ctypes structures and the Julia structs"""
    from importlib import import_module
    lib = import_module("gsa_b200._lib")
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gsa_cuda.h")).read(), flags=re.S)
    jl = open(os.path.join(ROOT, "julia", "GSACuda.jl")).read()
    for cname, cstruct, jname in (("gsa_sobol_opts", lib.SobolOpts, "GsaSobolOpts"), ("gsa_sobol_result", lib.SobolResultC, "GsaSobolResult")):
        body = re.search(r"typedef struct " + cname + r"\s*\{(.*?)\}", hdr, flags=re.S).group(1)
        fields_h = [re.split(r"[\s\*]+", f.strip())[-1] for f in body.split(";") if f.strip()]
        assert fields_h == [f[0] for f in cstruct._fields_], (cname, fields_h)
        jbody = re.search(r"struct " + jname + r"\b(.*?)\bend", jl, flags=re.S).group(1)
        fields_j = re.findall(r"(\w+)::", jbody)
        assert fields_j == fields_h, (jname, fields_j, fields_h)
