"""julia/CalipsoB200.jl is the ccall glue a CALIPSO.jl maintainer would add (src/solver/linear_solver.jl:1-60 seam, the hot
calls of solve.jl).  No Julia toolchain exists in the build image, so the file is checked mechanically against
include/calipso_b200.h: every `ccall((:name, LIB), ret, (argtypes...), args...)` names an exported function with the right
return type, argument count and argument types, and every CB200_* constant has the value the header's enums give it."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

JULIA_TO_C = {"Cint": "int", "Cdouble": "double", "Cstring": "const char *", "Cvoid": "void", "Ptr{Cvoid}": "ptr", "Ptr{Cint}": "int *",
              "Ptr{Cdouble}": "double *", "Ptr{Clonglong}": "long long *"}


def header_text():
    text = open(os.path.join(ROOT, "include", "calipso_b200.h")).read()
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def c_prototypes():
    """name -> (return type class, [argument type classes]) with handles / opaque pointers folded to 'ptr'"""
    protos = {}
    for m in re.finditer(r"^\s*([A-Za-z_][\w \*]*?)\s*\b(cb200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header_text(), flags=re.M | re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)

        def norm(t):
            t = re.sub(r"\bconst\b", "", t).strip()
            t = re.sub(r"\s+", " ", t)
            if "cb200_handle" in t or "cb200_options" in t or t in ("void *", "char *"):
                return "ptr" if t != "char *" else "char *"
            return t
        arglist = []
        if args.strip() and args.strip() != "void":
            for a in args.split(","):
                a = a.strip()
                a = re.sub(r"\b[A-Za-z_]\w*$", "", a).strip() if not a.endswith("*") else a      # drop the parameter name
                arglist.append(norm(a))
        protos[name] = (norm(ret), arglist)
    return protos


def header_enums():
    vals = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", header_text(), flags=re.S):
        cur = -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, v = [x.strip() for x in item.split("=")]
                cur = int(v, 0)
            else:
                name, cur = item, cur + 1
            vals[name] = cur
    return vals


def julia_source():
    return open(os.path.join(ROOT, "julia", "CalipsoB200.jl")).read()


def split_top_level(s):
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


def julia_ccalls():
    src = re.sub(r"#=.*?=#", "", julia_source(), flags=re.S)
    src = "\n".join(line.split("#")[0] for line in src.splitlines())
    calls = []
    for m in re.finditer(r"ccall\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        parts = split_top_level(src[m.end():i - 1])
        name = re.match(r"\(:(\w+),\s*LIBCB200\)", parts[0]).group(1)
        argtypes = split_top_level(parts[2].strip()[1:-1])
        argtypes = [a for a in argtypes if a]
        calls.append((name, parts[1], argtypes, parts[3:]))
    return calls


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 18
    for name, ret, argtypes, args in calls:
        assert name in protos, f"{name} is not declared in include/calipso_b200.h"
        cret, cargs = protos[name]
        jret = JULIA_TO_C[ret]
        assert jret == cret or (jret == "const char *" and cret == "char *"), (name, ret, cret)
        assert len(argtypes) == len(cargs) == len(args), (name, argtypes, cargs, args)
        for jt, ct in zip(argtypes, cargs):
            assert JULIA_TO_C[jt] == ct, (name, jt, ct)


def test_constants_match_the_header_enums():
    enums = header_enums()
    consts = re.findall(r"^const (CB200_\w+) = (?:Cint\()?(-?\d+)\)?", julia_source(), flags=re.M)
    assert len(consts) >= 30
    for name, value in consts:
        assert name in enums, name
        assert enums[name] == int(value), (name, value, enums[name])


def test_glue_covers_the_linear_solver_seam():
    """the four functions a LinearSolver must honour (linear_solver.jl:19,33,46,52) and the hot calls of solve!"""
    src = julia_source()
    for fn in ("b200_ldl_solver", "factorize!", "compute_inertia!", "linear_solve!", "residual!", "search_direction!", "cone!",
               "cone_search!", "apply_step!", "initialize!", "differentiate!"):
        assert re.search(r"^\s*(function )?" + re.escape(fn) + r"\(", src, flags=re.M), fn
    used = {c[0] for c in julia_ccalls()}
    for sym in ("cb200_ldl_create", "cb200_ldl_factorize", "cb200_ldl_inertia", "cb200_ldl_linear_solve", "cb200_create", "cb200_residual",
                "cb200_search_direction", "cb200_cone", "cb200_cone_search", "cb200_apply_step", "cb200_destroy", "cb200_amd_order"):
        assert sym in used, sym
