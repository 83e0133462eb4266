"""The C-ABI library builds, loads and exports every symbol include/b200ba.h declares.
No compute calls: this runs without a GPU."""
import ctypes as C
import os
import re

import pytest

from visual_slam_core_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200ba.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ba_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = _build.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b200ba.h but not exported"
    assert set(names) == set(_lib.EXPORTS)


def test_struct_layout_matches_header():
    lib = _lib.load()
    p = _lib.BAParams()
    lib.ba_default_params(C.byref(p))
    assert (p.huber_delta, p.chi2_thr, p.lambda0, p.max_iter) == (5.0, 5.991, 1e-4, 10)
    assert (p.lambda_up, p.lambda_down, p.lambda_min, p.lambda_max) == (10.0, 0.1, 1e-6, 1e8)
    assert p.use_graph == 1 and p.profile == 0 and p.pcg_tol == 1e-10
    assert C.sizeof(_lib.BAParams) == 8 * 8 + 4 * 4 + 8 + 2 * 4
    assert p.accept_rtol == 1e-8
    assert lib.ba_version() >= 100


def test_error_strings_and_no_cpu_fallback():
    import torch
    lib = _lib.load()
    assert lib.ba_error_string(0) == b"ok"
    assert b"no CPU fallback" in lib.ba_error_string(_lib.BA_ERR_NO_DEVICE)
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert lib.ba_create(None, 0, C.byref(h)) == _lib.BA_ERR_NO_DEVICE
        from visual_slam_core_b200 import LocalBASolver
        with pytest.raises(_lib.B200BAError):
            LocalBASolver()


def test_sass_is_sm100a():
    """The shipped cubin targets sm_100a (cuobjdump is in the image)."""
    import shutil
    import subprocess
    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cu):
        pytest.skip("cuobjdump not present")
    out = subprocess.run([cu, "-lelf", _build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_diagnostic_library_override(monkeypatch, tmp_path):
    """`BA_LIB_PATH` points `_lib` at a diagnostic build of the same library (`_build.build(out=..., extra=[...])`);
    without it the in-tree `libb200ba.so` is the only library the product loads."""
    from visual_slam_core_b200 import _build, _lib
    monkeypatch.delenv("BA_LIB_PATH", raising=False)
    assert _lib.lib_path() == _build.LIB_PATH
    other = str(tmp_path / "libb200ba_trace.so")
    monkeypatch.setenv("BA_LIB_PATH", other)
    assert _lib.lib_path() == other


def test_report_mirror_matches_the_header():
    """Every field of `ba_report` in include/b200ba.h has its ctypes twin, in order (a mismatch would shift the report)."""
    import re
    from visual_slam_core_b200 import _lib
    src = open(os.path.join(ROOT, "include", "b200ba.h")).read()
    body = src[src.index("typedef struct ba_report"):src.index("} ba_report;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b(?:int32_t|uint8_t|float|double)\s+(\w+)\s*(?:\[[^\]]*\])?\s*;", body)
    assert names == [f[0] for f in _lib.BAReport._fields_]
