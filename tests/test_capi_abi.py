"""The C-ABI library loads and exports every symbol include/vpl_capi.h declares (no GPU
compute here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "vpl_capi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vpl_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for n in ("vpl_create", "vpl_destroy", "vpl_lsd_detect_batch", "vpl_lbd_compute_batch", "vpl_match_batch",
              "vpl_frontend_batch", "vpl_last_error", "vpl_edlines_detect_batch", "vpl_linematch_batch",
              "vpl_linefront_batch"):
        assert n in names


def test_library_exports_every_declared_symbol(vpl):
    lib = vpl.capi.load()
    for n in declared_functions():
        assert hasattr(lib, n), f"{n} declared in vpl_capi.h but not exported"
    assert set(vpl.capi.EXPORTS) == set(declared_functions())


def test_struct_layouts(vpl):
    assert vpl.capi.KEYLINE_DTYPE.itemsize == 68   # cv::line_descriptor::KeyLine, 17 x 4 B
    assert vpl.capi.DMATCH_DTYPE.itemsize == 16    # cv::DMatch
    assert vpl.capi.SEGMENT_DTYPE.itemsize == 40
    assert ctypes.sizeof(vpl.capi.VplConfig) == 40
    assert vpl.capi.LINE_DTYPE.itemsize == 56      # struct Line's numeric fields (line.h:8-12)
    assert ctypes.sizeof(vpl.capi.EDLineParam) == 32 and ctypes.sizeof(vpl.capi.LineMatchParam) == 72
    assert b"sm_100a" in vpl.capi.load().vpl_version()


def test_no_cpu_fallback(vpl):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    if vpl.capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(vpl.VplError) as e:
        vpl.Context()
    assert e.value.code == vpl.capi.VPL_E_NODEVICE
    with pytest.raises(vpl.VplError):
        vpl.LSDDetector.createLSDDetector().detect(__import__("numpy").zeros((64, 64), "uint8"), 2, 1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "vplines-slam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+[\"<].*vpl_oracle", txt, flags=re.M), \
                    f"{f} references the oracle"
