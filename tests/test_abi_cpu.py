"""CPU checks: the C-ABI library loads and exports exactly what include/fdt_b200.h declares; the host-side
argument validation answers with error codes (no compute, no GPU needed); no CPU fallback exists."""
import ctypes as C
import os
import re

import pytest
import torch

import fdt_b200
from fdt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fdt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fdt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/fdt_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES table out of sync with the header"


def test_version_and_error_string():
    lib = _lib.lib()
    assert lib.fdt_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_argument_validation_returns_codes_without_touching_a_gpu():
    lib = _lib.lib()
    # nms_thresh <= 0 -> FDT_E_INVALID (mirrors the ValueError of detection.py:28-29); checked before any CUDA call
    rc = lib.fdt_detect_sort_nms(None, None, 1, 100, 2, 10, 5000, 0.0, 0.1, 0.2, None, None, None, C.c_void_p(256), 1 << 20, None)
    assert rc == _lib.FDT_E_INVALID and "nms_thresh" in _lib.last_error()
    rc = lib.fdt_detect_sort_nms(None, None, 1, 100, 2, 10, 9000, 0.3, 0.1, 0.2, None, None, None, C.c_void_p(256), 1 << 20, None)
    assert rc == _lib.FDT_E_UNSUPPORTED
    rc = lib.fdt_detect_threshold_compact(None, 1, 100, 2, 0.05, C.c_void_p(256), 16, None)
    assert rc == _lib.FDT_E_WORKSPACE
    rc = lib.fdt_priorbox(640.0, 640.0, 4.0, 16.0, 9, None, 0, None, 4, 4, None, None)
    assert rc == _lib.FDT_E_UNSUPPORTED
    with pytest.raises(ValueError):
        _lib.check(_lib.FDT_E_INVALID)
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.FDT_E_UNSUPPORTED)


def test_workspace_size_queries():
    lib = _lib.lib()
    assert lib.fdt_detect_workspace_bytes(64, 34125, 2) >= 64 * 34125 * 8
    assert lib.fdt_detect_workspace_bytes(64, 34125, 1) == 256
    assert lib.fdt_nms_workspace_bytes(5000) >= 5000 * 8
    assert lib.fdt_multibox_workspace_bytes(32, 34125, 2, 3200) > lib.fdt_match_workspace_bytes(32, 34125, 3200)
    assert lib.fdt_iou_track_workspace_bytes(10000, 1500000, 300) >= 1500000 * 10 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a GPU-less machine")
def test_no_cpu_fallback():
    from fdt_b200.layers import Detect, PriorBoxLayer
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Detect(2, 0, 750, 0.05, 0.3)(torch.zeros(1, 10, 4), torch.zeros(1, 10, 2), torch.zeros(10, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PriorBoxLayer(640, 640)(0, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fdt_b200.tracker.iou_track([torch.zeros(1, 5).numpy()])


def test_detect_constructor_contract():
    from fdt_b200.layers import Detect
    d = Detect(2, 0, 750, 0.05, 0.3)
    assert (d.num_classes, d.background_label, d.top_k, d.nms_thresh, d.conf_thresh, d.nms_top_k) == (2, 0, 750, 0.3, 0.05, 5000)
    assert d.variance == [0.1, 0.2]
    with pytest.raises(ValueError):
        Detect(2, 0, 750, 0.05, 0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "face-detection-and-tracking_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|libfdt_oracle|orc_[a-z_]+\(", txt, flags=re.M), \
                    f"{f} reaches into oracle/"


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/fdt_b200.h must compile as C (no C++ or torch types in the signatures)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "h.c"
    src.write_text('#include "fdt_b200.h"\nint main(void) { return fdt_version() > 0 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_pack_targets_layout():
    """MultiBoxLoss target packing (layers/modules/multibox_loss.py): list[B] of [G_i,5] -> gt[total,5] fp32 + int64 offsets; images
    without boxes get an empty range, other dtypes / shapes take the converting path, an all-empty batch a 1-row placeholder."""
    from fdt_b200.layers.modules.multibox_loss import pack_targets
    cpu = torch.device("cpu")
    a = torch.arange(10, dtype=torch.float32).reshape(2, 5)
    b = torch.zeros((0, 5), dtype=torch.float32)
    c = torch.arange(15, dtype=torch.float32).reshape(3, 5) + 100
    gt, off, total = pack_targets([a, b, c], cpu)
    assert total == 5 and off.dtype == torch.int64 and off.tolist() == [0, 2, 2, 5]
    assert gt.dtype == torch.float32 and torch.equal(gt, torch.cat([a, c], 0))
    gt3, off3, total3 = pack_targets([a.double(), None, c.double()], cpu)           # float64 and a missing entry
    assert total3 == 5 and off3.tolist() == [0, 2, 2, 5] and gt3.dtype == torch.float32 and torch.equal(gt3, torch.cat([a, c], 0))
    gt4, off4, total4 = pack_targets([b, None], cpu)
    assert total4 == 0 and off4.tolist() == [0, 0, 0] and tuple(gt4.shape) == (1, 5)


def test_multibox_workspace_grows_with_the_batch_and_stays_aligned():
    L = _lib.lib()
    w1 = L.fdt_multibox_workspace_bytes(1, 34125, 2, 10)
    w32 = L.fdt_multibox_workspace_bytes(32, 34125, 2, 3000)
    assert 0 < w1 < w32 and w1 % 256 == 0 and w32 % 256 == 0
    assert L.fdt_multibox_workspace_bytes(0, 0, 2, 0) == 256
