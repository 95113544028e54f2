"""CPU: the C-ABI library builds, loads and exports every symbol include/lbm_b200.h declares
(no compute calls without a GPU), and argument validation that needs no device works."""
import ctypes
import os
import re

from lbm_b200 import _capi as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    C.build()
    raw = ctypes.CDLL(C.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(raw, s), "missing export " + s
    # the ctypes table binds exactly the declared set
    assert sorted(C.SIGNATURES) == syms
    assert C.lib().lbm_abi_version() == 1


def test_struct_layout_matches_header():
    assert ctypes.sizeof(C.LbmCfg) == 4 * 8 + 2 * 8 + 4 * 4
    assert ctypes.sizeof(C.LbmLayout) == 6 * 8
    assert ctypes.sizeof(C.LbmPeerInfo) == 3 * 64 + 3 * 8 + 8 * 8      # lbm_peer_info


def test_null_arguments_are_rejected_without_a_device():
    L = C.lib()
    assert L.lbm_create(None, None) == -1
    h = C.c_vp()
    cfg = C.LbmCfg(nx=2, ny=2, x0=0, nxl=2, om_p=1.0, om_m=1.0)
    assert L.lbm_create(ctypes.byref(cfg), ctypes.byref(h)) == -1     # too small
    assert b"3x3" in L.lbm_last_error()
    cfg = C.LbmCfg(nx=16, ny=16, x0=8, nxl=16, om_p=1.0, om_m=1.0)
    assert L.lbm_create(ctypes.byref(cfg), ctypes.byref(h)) == -1     # slab outside
    assert L.lbm_destroy(None) == 0
    assert L.lbm_wall_row_len(None) == 0
