"""The C-ABI shared library: builds for sm_100a, loads, and exports every symbol include/pfn_b200.h declares.
No compute calls (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from poweflownet_b200.build import build
    return build()


def _declared():
    text = open(os.path.join(ROOT, "include", "pfn_b200.h")).read()
    return sorted(set(re.findall(r"^PFN_API[^;(]*?\b(pfn_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("pfn_graph_prep", "pfn_ea_fwd", "pfn_ea_bwd", "pfn_spmm_hop", "pfn_linear_fwd", "pfn_linear_dgrad",
                 "pfn_linear_wgrad", "pfn_mpn_forward", "pfn_mpn_backward", "pfn_mse_fwd_bwd", "pfn_version", "pfn_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(libpath):
    handle = ctypes.CDLL(libpath)
    for name in _declared():
        assert hasattr(handle, name), name


def test_python_binding_covers_the_header(libpath):
    from poweflownet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert _lib.lib().pfn_version().startswith(b"pfn_b200")


def test_plain_c_signatures_no_torch_types():
    text = open(os.path.join(ROOT, "include", "pfn_b200.h")).read()
    assert "torch" not in text.lower().replace("pytorch", "") or "no torch types" in text
    assert "at::" not in text and "c10::" not in text and "#include <torch" not in text


def test_model_descriptor_and_workspace_queries(libpath):
    import ctypes as C
    from poweflownet_b200._lib import GraphLayout, MpnDesc, lib
    d = MpnDesc(4, 2, 4, 129, 4, 3, 0.2, 0)
    assert lib().pfn_mpn_num_params(C.byref(d)) == 35
    act, scr = C.c_size_t(), C.c_size_t()
    assert lib().pfn_mpn_workspace(C.byref(d), 15104, 23808, C.byref(act), C.byref(scr)) == 0
    assert act.value > 15104 * 132 * 4 * 24 and scr.value > 0
    lay = GraphLayout()
    assert lib().pfn_graph_layout_get(15104, 23808, C.byref(lay)) == 0 and lay.e_cap == 47616
    bad = MpnDesc(4, 5, 4, 129, 4, 3, 0.2, 0)
    assert lib().pfn_mpn_num_params(C.byref(bad)) == -1
    assert b"efeature_dim" in lib().pfn_last_error()


def test_sass_is_sm100a_only(libpath):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
