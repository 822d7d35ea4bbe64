"""CPU: the C-ABI library loads and exports every symbol include/pcd.h declares; without a GPU the
product path fails loudly (no CPU fallback); the product never imports the oracle."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pcd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcd_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pcd):
    L = ctypes.CDLL(pcd.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/pcd.h but not exported"
    assert sorted(n for n, _, _ in pcd.ABI) == syms, "python binding table out of sync with include/pcd.h"
    assert pcd.lib().pcd_abi_version() == 2


def test_no_cpu_fallback(pcd):
    if pcd.device_count() > 0:
        pytest.skip("a GPU is present")
    import numpy as np
    with pytest.raises(pcd.PcdError) as e:
        pcd.poisson_solver(np.zeros((4, 4)), np.zeros((4, 4)), 4, 4, 10, 1e-7)
    assert e.value.status == pcd.PCD_ERR_NO_DEVICE
    with pytest.raises(pcd.PcdError):
        pcd.Solver(8, 8)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "poisson_caustic_design_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("pcd_oracle.h", "libpcd_oracle", "libpcd_ref", "pcdo_", "import oracle", "from oracle"):
                    assert needle not in src, f"{f} references the oracle ({needle})"


def test_argument_errors_need_no_device(pcd):
    """Entry points reject bad handles / arguments with PCD_ERR_INVALID before touching CUDA (and say why)."""
    L = pcd.lib()
    null = ctypes.c_void_p()
    buf = (ctypes.c_ubyte * 256)()
    t = ctypes.c_int(0)
    assert L.pcd_set_solve_hook(null, None, None) == pcd.PCD_ERR_INVALID
    assert L.pcd_slab_peer_export(null, buf) == pcd.PCD_ERR_INVALID
    assert L.pcd_slab_peer_connect_ipc(null, 0, buf, 0, 16) == pcd.PCD_ERR_INVALID
    assert L.pcd_slab_peer_connect_local(null, 0, null) == pcd.PCD_ERR_INVALID
    assert L.pcd_slab_peer_run(null, 2, 0) == pcd.PCD_ERR_INVALID
    assert L.pcd_slab_peer_status(null, ctypes.byref(t)) == pcd.PCD_ERR_INVALID
    assert L.pcd_slab_load_device(null, null, null) == pcd.PCD_ERR_INVALID
    assert L.pcd_slab_store_device(null, null) == pcd.PCD_ERR_INVALID
    assert L.pcd_slab_pass(null, 2, 0) == pcd.PCD_ERR_INVALID
    assert L.pcd_slab_peer_handle_bytes() == 128
    assert L.pcd_slab_ghost_rows() == 2 * L.pcd_slab_sweeps_per_pass() + 1
    out = ctypes.c_void_p()
    assert L.pcd_slab_create(16, 16, 8, 16, 0, None, ctypes.byref(out)) == pcd.PCD_ERR_INVALID   # rows beyond the grid
    assert b"not a valid slab" in ctypes.c_char_p(L.pcd_last_error()).value


def test_auto_path_selection_table(pcd):
    """pcd_solver_plan: what SOLVER_AUTO chooses on a 148-SM device (host logic of csrc/sor_resident.cu: resident_plan)."""
    plan = pcd.solver_plan
    # the metric's grids
    assert plan(1024, 1024) == {"path": "resident", "rows_per_cta": 7, "ctas": 147, "transposed": False, "deep_only": False}
    assert plan(400, 400) == {"path": "resident", "rows_per_cta": 3, "ctas": 134, "transposed": False, "deep_only": False}
    assert plan(1024, 512)["rows_per_cta"] == 4 and plan(1024, 512)["ctas"] == 128
    assert plan(8192, 8192)["path"] == "tiled" and plan(2048, 2048)["path"] == "tiled"
    # short even-width grids: fewer CTAs of three rows (so that the deep-halo kernel applies); odd widths keep one/two rows
    assert plan(300, 157) == {"path": "resident", "rows_per_cta": 3, "ctas": 53, "transposed": False, "deep_only": False}
    assert plan(157, 300)["rows_per_cta"] == 3 and plan(301, 157)["rows_per_cta"] == 2 and plan(64, 2)["rows_per_cta"] == 1
    # taller than 7 rows per CTA: deep-halo kernel only, up to 9 rows (even widths)
    assert plan(1024, 1036)["deep_only"] is False and plan(1024, 1037) == {"path": "resident", "rows_per_cta": 8, "ctas": 130,
                                                                            "transposed": False, "deep_only": True}
    assert plan(1024, 1332)["rows_per_cta"] == 9 and plan(1024, 1333)["path"] == "tiled" and plan(1023, 1100)["path"] == "tiled"
    # wider than 1024 columns: transposed when the height fits the thread layout and is even
    assert plan(1280, 720) == {"path": "resident", "rows_per_cta": 9, "ctas": 143, "transposed": True, "deep_only": True}
    assert plan(1280, 320)["transposed"] and plan(1332, 1024)["transposed"] and plan(1030, 4)["transposed"]
    assert plan(1280, 721)["path"] == "tiled" and plan(1334, 1024)["path"] == "tiled" and plan(1100, 1100)["path"] == "tiled"
    # every slab has two or three.. rows: the even split never produces an empty or one-row slab for the deep-halo kernel
    for (w, h) in ((64, 3), (64, 4), (64, 5), (64, 7), (64, 443), (64, 445), (1024, 1037), (1200, 64)):
        p = plan(w, h)
        rows_total = w if p["transposed"] else h
        n_small = p["ctas"] * p["rows_per_cta"] - rows_total
        assert 0 <= n_small <= p["ctas"] and (p["rows_per_cta"] >= 3 and p["rows_per_cta"] - 1 >= 2), (w, h, p)
