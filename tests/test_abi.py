"""CPU tests of the drop-in boundary: the library loads, exports every declared symbol, and never computes on the host."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, cuda_available


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hvb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hvb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(hvb):
    L = hvb._abi.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), name
    assert set(hvb._abi.EXPORTS) == set(names)
    assert b"sm_100a" in L.hvb_version()


def test_default_params_match_reference(hvb):
    p = hvb.RaycastParameter()
    # raycast-types.jl:226-230
    assert (p.variance_tol, p.break_tol, p.b_nodes_tol, p.plane_tolerance, p.ray_tol) == (1e-15, 1e-5, 1e-7, 1e-12, 1e-12)
    assert p.method == hvb.RCStandard == hvb.RCNonGeneralHP and p.world == 1 and p.fp32_filter == 1 and p.sort_output == 1
    assert p.persistent == 3                  # the persistent walk with warp-aggregated atomics (include/hvb200.h)
    assert ctypes.sizeof(hvb._abi.hvb_params) == 5 * 8 + 12 * 4 + 8 + 8 + 8 + 8
    assert ctypes.sizeof(hvb._abi.hvb_stats_t) == 31 * 8


def test_argument_errors(hvb):
    L = hvb._abi.lib()
    ctx = ctypes.c_void_p()
    xs = np.random.default_rng(0).random((3, 3))
    # sysvoronoi.jl:25-27: "There are not enough points to create a Voronoi tessellation"
    rc = L.hvb_create(ctypes.byref(ctx), 3, 3, xs.ctypes.data_as(ctypes.c_void_p), 0, None, None, None)
    assert rc == hvb._abi.HVB_EINVAL and b"not enough points" in L.hvb_last_error(None)
    rc = L.hvb_create(ctypes.byref(ctx), 7, 100, xs.ctypes.data_as(ctypes.c_void_p), 0, None, None, None)
    assert rc == hvb._abi.HVB_EINVAL
    assert L.hvb_search(None, None, 0, None, None, 0, 0) == hvb._abi.HVB_EINVAL


def test_cuboid_planes_match_reference_layout(hvb):
    # boundary.jl:510-534: plane 2i-1 = upper face (normal +e_i, base offset + dim_i e_i), plane 2i = lower face
    b = hvb.cuboid(3, periodic=[])
    assert len(b) == 6
    assert np.array_equal(b.normal[0], [1, 0, 0]) and np.array_equal(b.base[0], [1, 0, 0])
    assert np.array_equal(b.normal[1], [-1, 0, 0]) and np.array_equal(b.base[1], [0, 0, 0])
    assert hvb.cuboid(2).periodic == (1, 2)


@pytest.mark.skipif(cuda_available(), reason="a CUDA device is present")
def test_no_cpu_fallback(hvb):
    with pytest.raises(hvb.HVBError) as e:
        hvb.Raycast(np.random.default_rng(0).random((50, 3)))
    assert e.value.code == hvb._abi.HVB_ENOGPU


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "highvoronoi.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "hv_oracle" not in text and "qhull_oracle" not in text and "hostsim" not in text.replace("tests/hostsim", ""), f


def _c_struct_fields(src, name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return [m.group(2) for m in re.finditer(r"\b(double|int32_t|int64_t)\s+([a-z_0-9]+)\s*;", body)]


def test_struct_layouts_agree_across_header_python_and_julia(hvb):
    """hvb_params / hvb_stats_t are mirrored by hand in _abi.py and in the Julia shim: same fields, same order"""
    hdr = open(os.path.join(ROOT, "include", "hvb200.h")).read()
    hdr = hdr.replace("typedef struct hvb_params {", "typedef struct hvb_params {").replace("} hvb_params;", "} hvb_params;")
    params = _c_struct_fields(hdr, "hvb_params")
    assert params == [f[0] for f in hvb._abi.hvb_params._fields_]
    stats_body = re.search(r"typedef struct \{(.*?)\} hvb_stats_t;", hdr, flags=re.S)
    if stats_body is None:
        stats_body = re.search(r"typedef struct hvb_stats_t \{(.*?)\} hvb_stats_t;", hdr, flags=re.S)
    body = re.sub(r"/\*.*?\*/", "", stats_body.group(1), flags=re.S)
    stats = [m.group(2) for m in re.finditer(r"\b(double|int32_t|int64_t)\s+([a-z_0-9]+)\s*;", body)]
    assert stats == [f[0] for f in hvb._abi.hvb_stats_t._fields_]
    jl = open(os.path.join(ROOT, "julia", "HighVoronoiB200.jl")).read()
    jbody = re.search(r"mutable struct HvbParams(.*?)HvbParams\(\) = new\(\)", jl, flags=re.S).group(1)
    jfields = re.findall(r"([a-z_0-9]+)::(?:Cdouble|Int32|Int64)", jbody)
    assert jfields == params


# ---- the ccalls of the Julia shim against the C prototypes ---------------------------------------------------------------
def _split_top(s):
    """split at commas that are not nested in (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(s, start):
    """the text between the parenthesis at s[start] and its partner"""
    depth = 0
    for i in range(start, len(s)):
        depth += s[i] == "("
        depth -= s[i] == ")"
        if depth == 0:
            return s[start + 1:i]
    raise ValueError("unbalanced")


def c_prototypes():
    src = open(os.path.join(ROOT, "include", "hvb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|void|const char\s*\*)\s+(hvb_[a-z_0-9]+)\s*\(", src):
        args = _balanced(src, m.end() - 1)
        protos[m.group(2)] = (m.group(1).replace(" ", ""), [] if args.strip() == "void" else [" ".join(a.split()) for a in _split_top(args)])
    return protos


def c_kind(arg):
    """pointer class or scalar type of a C parameter declaration"""
    if "*" in arg:
        base = arg.replace("const", "").split("*")[0].strip()
        return "ptr:" + base + ("*" if arg.count("*") > 1 else "")
    return arg.replace("const", "").split()[0]


JL_SCALAR = {"Cint": "int", "Int32": "int", "Int64": "int64_t", "Cdouble": "double", "Float64": "double"}
JL_POINTEE = {"Cvoid": {"hvb_ctx", "void"}, "Float64": {"double"}, "Int64": {"int64_t"}, "Int32": {"int32_t", "int"}, "UInt8": {"uint8_t"},
              "HvbParams": {"hvb_params"}, "Ptr{Cvoid}": {"hvb_ctx*"}}


def test_every_ccall_of_the_julia_shim_matches_its_c_prototype():
    """the shim cannot be executed here (no Julia): what CAN be checked is that every ccall names an exported function and
    passes the number and kinds of arguments the header declares -- the mistakes an unexecuted binding typically carries"""
    jl = open(os.path.join(ROOT, "julia", "HighVoronoiB200.jl")).read()
    jl = "\n".join(l.split("#")[0] if not l.lstrip().startswith("#") else "" for l in jl.split("\n"))
    protos = c_prototypes()
    seen = set()
    for m in re.finditer(r"ccall\(", jl):
        parts = _split_top(_balanced(jl, m.end() - 1))
        name = re.match(r"\(:(hvb_[a-z_0-9]+),\s*LIB\)", parts[0]).group(1)
        assert name in protos, name
        ret, cargs = protos[name]
        seen.add(name)
        assert {"int": "Cint", "void": "Cvoid", "constchar*": "Cstring"}[ret] == parts[1], (name, parts[1])
        types = _split_top(parts[2].strip()[1:-1].rstrip(","))
        values = parts[3:]
        assert len(types) == len(cargs) == len(values), (name, types, cargs, values)
        for jt, ca in zip(types, cargs):
            ck = c_kind(ca)
            if jt in JL_SCALAR:
                assert ck == JL_SCALAR[jt], (name, jt, ca)
            else:
                inner = re.match(r"(?:Ptr|Ref)\{(.*)\}$", jt)
                assert inner and ck.startswith("ptr:"), (name, jt, ca)
                assert ck[4:] in JL_POINTEE[inner.group(1)], (name, jt, ca)
    # the entry points the seam needs are all bound
    assert {"hvb_create", "hvb_create_multi", "hvb_create_periodic", "hvb_set_points", "hvb_search", "hvb_counts", "hvb_fetch_vertices",
            "hvb_fetch_vertices_var", "hvb_fetch_rays", "hvb_last_error", "hvb_destroy", "hvb_default_params", "hvb_convex_hull"} <= seen


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/hvb200.h compiles as C99 (no C++ types at the boundary), examples/c_example.c links against libhvb200.so, and on a
    box without a GPU the run ends at hvb_create with HVB_ENOGPU -- never with a result"""
    import shutil
    import subprocess
    import hvb200
    hvb200._abi.lib()
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    libdir = os.path.join(ROOT, "highvoronoi.jl_b200", "lib")
    exe = str(tmp_path / "c_example")
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "examples", "c_example.c"), "-L" + libdir, "-lhvb200", "-Wl,-rpath," + libdir, "-lm", "-o", exe],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    from conftest import cuda_available
    if cuda_available():
        return                                              # the run itself belongs to the GPU suite (test_gpu_parity.py)
    run = subprocess.run([exe, "500", "3"], capture_output=True, text=True)
    assert run.returncode == 3 and "no host fallback" in run.stderr and "vertices" not in run.stdout
