"""CPU: the Julia binding (julia/GPB200.jl) cannot be executed here (no Julia in the image), so it is checked
statically against include/gpb200.h: every `ccall` names an exported symbol, passes as many argument types as the C
prototype has parameters, as many values as types, and C-compatible type names; struct constructors pass as many
values to `new` as the struct declares fields; the dense update_cK! methods are typed like the reference's pair
(src/GPE.jl:169,177) so that dispatch is unambiguous."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "gaussianprocesses.jl_b200", "julia", "GPB200.jl")
HDR = os.path.join(ROOT, "include", "gpb200.h")


def _split_top(s):
    """split on commas that are not nested in (), [] or {}"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(s, start):
    """s[start] == '(' -> index one past the matching ')'"""
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise AssertionError("unbalanced parentheses")


def header_prototypes():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|void|int64_t|const char\*)\s+(gpb200_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        protos[m.group(2)] = [] if args in ("", "void") else _split_top(args)
    return protos


def julia_ccalls():
    src = open(JL).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), LIB\)", src):
        end = _balanced(src, m.start() + len("ccall"))
        parts = _split_top(src[m.start() + len("ccall("):end - 1])
        # parts: (sym, LIB) | RetType | (ArgTypes...) | values...
        types = parts[2].strip()
        assert types.startswith("(") and types.endswith(")"), parts
        tlist = _split_top(types[1:-1])
        calls.append((m.group(1), parts[1].strip(), tlist, parts[3:]))
    return calls


C2JL = {"int": {"Cint"}, "int32_t": {"Int32", "Cint"}, "int64_t": {"Int64"}, "double": {"Float64", "Cdouble"},
        "const double*": {"Ptr{Float64}", "Ref{Float64}"}, "double*": {"Ptr{Float64}", "Ref{Float64}"},
        "const int32_t*": {"Ptr{Int32}"}, "const int64_t*": {"Ptr{Int64}"}, "int64_t*": {"Ptr{Int64}"}, "gpb200_handle*": {"Ptr{Cvoid}"}, "gpb200_fitc*": {"Ptr{Cvoid}"},
        "gpb200_handle**": {"Ref{Ptr{Cvoid}}"}, "gpb200_fitc**": {"Ref{Ptr{Cvoid}}"}, "const char*": {"Cstring", "Ptr{UInt8}"},
        "char*": {"Ptr{UInt8}"}, "void*": {"Ptr{Cvoid}"}}


def _ctype(decl):
    d = re.sub(r"\b\w+$", "", decl.strip()).strip() if not decl.strip().endswith("*") else decl.strip()
    return re.sub(r"\s*\*", "*", d)


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 20
    for name, ret, tlist, values in calls:
        assert name in protos, "ccall of %s: not declared in include/gpb200.h" % name
        assert len(tlist) == len(protos[name]), "%s: %d argument types vs %d C parameters" % (name, len(tlist), len(protos[name]))
        assert len(values) == len(tlist), "%s: %d values for %d argument types" % (name, len(values), len(tlist))
        for jt, cdecl in zip(tlist, protos[name]):
            ct = _ctype(cdecl)
            assert ct in C2JL, (name, cdecl, ct)
            assert jt in C2JL[ct], "%s: Julia type %s for C parameter '%s'" % (name, jt, cdecl)


def test_struct_constructors_match_their_fields():
    src = open(JL).read()
    for m in re.finditer(r"mutable struct (\w+)[^\n]*\n(.*?)\nend\n", src, flags=re.S):
        body = m.group(2)
        head = body.split("function")[0]
        fields = [ln for ln in head.splitlines() if re.match(r"\s+\w+::", ln)]
        for nm in re.finditer(r"\bnew\(", body):
            end = _balanced(body, nm.end() - 1)
            nargs = len(_split_top(body[nm.end():end - 1]))
            assert nargs == len(fields), "%s: new(...) passes %d values, struct has %d fields" % (m.group(1), nargs, len(fields))
        for use in re.finditer(r"cK\.(\w+)\s*=", src):
            pass
    # every field assigned through `cK.<f> =` exists in one of the PDMat structs
    declared = set(re.findall(r"^\s+(\w+)::", src, flags=re.M))
    for f in set(re.findall(r"\bcK\.(\w+)\s*=[^=]", src)):
        assert f in declared, "assignment to undeclared field cK.%s" % f


def test_dense_update_ck_is_typed_like_the_reference_pair():
    src = open(JL).read()
    sigs = re.findall(r"update_cK!\(cK::B200PDMat,[^)]*\)", src)
    assert any("logNoise::Real" in s for s in sigs) and any("logNoise::AbstractVector" in s for s in sigs), sigs
    assert not any(re.search(r"logNoise\s*,", s) for s in sigs), "untyped logNoise is ambiguous with src/GPE.jl:169"
    # predictMVN is bound to the device (predict_full / rand reach it), not an error stub
    body = src[src.index("function predictMVN("):]
    body = body[:body.index("\nend\n")]
    assert "error(" not in body and "predict_raw" in body


def test_shim_binds_the_crossvalidation_and_elastic_entry_points():
    """crossvalidation.jl:142,180,311 and GPEelastic.jl:13 get device-backed methods: each is imported from the reference
    module (so the definitions EXTEND its generic functions) and reaches its C entry point."""
    src = open(JL).read()
    bound = {c[0] for c in julia_ccalls()}
    for sym in ("gpb200_cv_param", "gpb200_cv_block", "gpb200_append", "gpb200_rand", "gpb200_set_option",
                "gpb200_get_inverse_diag", "gpb200_fitc_predict_cov"):
        assert sym in bound, sym
    imports = src[src.index("import GaussianProcesses:"):src.index("import PDMats")]
    for fn in ("predict_LOO", "predict_CVfold", "dlogpdθ_LOO", "dlogpdθ_CVfold", "Folds", "update_target!", "get_value"):
        assert fn in re.split(r"[\s,:]+", imports), fn
    for sig in (r"function dlogpdθ_LOO\(gp::GPE\{X,Y,M,K,CS,D,P\}; noise::Bool, domean::Bool, kern::Bool\)",
                r"function predict_CVfold\(cK::B200PDMat, alpha::AbstractVector\{<:Real\}, y::AbstractVector\{<:Real\}, folds::Folds\)",
                r"function dlogpdθ_CVfold\(gp::GPE\{X,Y,M,K,CS,D,P\}, folds::Folds; noise::Bool, domean::Bool, kern::Bool\)",
                r"function Base\.append!\(gp::GPE\{X,Y,M,K,CS,D,P\}, x::AbstractMatrix, y::AbstractVector\)"):
        assert re.search(sig, src), sig
    # 0-based parameter / row indices cross the ABI (include/gpb200.h), 1-based ones stay in Julia
    assert "cv_param(cK, j - 1, alpha)" in src and "Vector{Int64}(V .- 1)" in src
