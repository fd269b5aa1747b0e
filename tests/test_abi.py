"""The C-ABI library builds for sm_100a, loads, and exports exactly what include/artiboost_b200.h declares.
No compute calls here (CPU suite): only argument-validation paths that return before touching the device."""
import ctypes as C
import os
import re
import subprocess

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "artiboost_b200.h")).read()
    return sorted(set(re.findall(r"AB_API\s+[\w\s\*]+?\b(ab_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for s in ["ab_version", "ab_last_error", "ab_launch_count", "ab_mano_forward", "ab_ccv_sample", "ab_view_from_id",
              "ab_pose_generate", "ab_pose_generate_workspace_bytes", "ab_render_batch", "ab_render_workspace_bytes"]:
        assert s in syms


def test_library_exports_every_declared_symbol(lib_built):
    from artiboost_b200 import lib
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert [s for s in exported if s.startswith("ab_")] == header_symbols()
    assert set(lib.EXPORTS) == set(header_symbols())
    for s in header_symbols():
        assert hasattr(lib_built, s)


def test_library_is_sm100a_with_lineinfo(lib_built):
    from artiboost_b200 import lib
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_argument_errors_are_reported_without_a_device(lib_built):
    L = lib_built
    assert L.ab_version() >= 100
    rc = L.ab_mano_forward(None, 4, None, None, None, -1, None, None, None, None)
    assert rc == -1 and b"null model" in L.ab_last_error()
    rc = L.ab_ccv_sample(None, 0, 1, 1, None, 1, None, None, None, None, None, None)
    assert rc == -1 and b"empty CCV space" in L.ab_last_error()
    rc = L.ab_view_from_id(None, -1, 12, 24, 0.45, 0.55, None, None, None, None, None)
    assert rc == -1
    assert L.ab_pose_generate_workspace_bytes(256) == 256 * 60 * 4
    rc = L.ab_render_batch(None, None, 1, 1, None, None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1
    # empty batches are no-ops
    from artiboost_b200.lib import ManoModelStruct
    m = ManoModelStruct(1, 1, 1, 1, 1, 1)
    assert L.ab_mano_forward(C.byref(m), 0, None, None, None, -1, None, None, None, None) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from artiboost_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        lib.load()
    except lib.AbError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the CUDA library is missing")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "artiboost_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f
