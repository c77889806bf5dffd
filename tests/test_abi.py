"""The C-ABI library loads and exports every symbol include/osep_skel.h declares (no compute)."""
import os
import re

import pytest

import osep_path_planner_b200 as osep
from osep_path_planner_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "osep_skel.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(osep_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    L = osep.load()
    syms = header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, "declared in include/osep_skel.h but not exported: %s" % missing
    assert sorted(_lib.SYMBOLS) == syms, "python binding table out of date"
    assert L.osep_abi_version() == 1


def test_config_defaults_are_the_reference_ros_parameter_defaults(built):
    """rosa_node.cpp:50-60 and gskel_node.cpp:62-70."""
    L = osep.load()
    c = osep.RosaConfig(max_points=1); L.osep_rosa_default_config(c)
    assert (c.max_points, c.min_points, c.ne_knn, c.nb_knn, c.niter_drosa, c.niter_dcrosa, c.niter_smooth) == (1000, 100, 20, 10, 6, 5, 3)
    assert (c.pts_dist_lim, c.max_proj_range, c.radius_smooth) == (50.0, 10.0, 5.0) and abs(c.alpha_recenter - 0.3) < 1e-7
    g = osep.GSkelConfig(); L.osep_gskel_default_config(g)
    assert (g.gnd_th, g.fuse_dist_th, g.max_obs_wo_conf, g.niter_smooth_vertex, g.min_branch_length) == (60.0, 3.0, 5, 3, 5)
    assert abs(g.fuse_conf_th - 0.1) < 1e-7 and abs(g.lkf_pn - 1e-4) < 1e-10 and abs(g.lkf_mn - 0.1) < 1e-7
    # python-side defaults agree with the library's
    d = osep.RosaConfig()
    assert all(getattr(d, f) == getattr(c, f) for f, _ in osep.RosaConfig._fields_)


def test_no_cpu_fallback_without_a_gpu(built):
    """On a box with no usable B200 the product must fail loudly, never route elsewhere."""
    L = osep.load()
    if L.osep_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(osep.OsepError):
        osep.Rosa()
    with pytest.raises(osep.OsepError):
        from osep_path_planner_b200 import GSkel
        GSkel()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "osep_path_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), "%s mentions the oracle" % os.path.join(dirpath, f)


def test_header_is_plain_c_and_shims_compile(tmp_path):
    """include/osep_skel.h must be consumable by a C compiler (cgo / JNI / ctypes-style bindings) and the C++ shim
    classes must compile against it alone (no PCL / Eigen / ROS, no CUDA headers)."""
    import subprocess
    c = tmp_path / "abi.c"
    c.write_text('#include "osep_skel.h"\nint main(void) { osep_rosa_config c; osep_rosa_default_config(&c); return (int)sizeof(osep_graph_view) == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(c)])
    cpp = tmp_path / "shim.cpp"
    cpp.write_text('#include "rosa.hpp"\n#include "gskel.hpp"\nint main() { return 0; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "osep_path_planner_b200", "host"), str(cpp)])
