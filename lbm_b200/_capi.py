"""ctypes binding of liblbm_b200.so (include/lbm_b200.h).

There is no CPU fallback: if the CUDA library is missing or a call fails, an
exception is raised.  Nothing here imports or calls oracle/.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# (LBM_B200_LIB: development only -- A/B runs of two builds of the library on the same GPU box)
LIB_PATH = os.environ.get("LBM_B200_LIB") or os.path.join(_HERE, "liblbm_b200.so")
_lib = None

LBM_F64, LBM_F32 = 0, 1
LBM_ARITH_FUSED, LBM_ARITH_STRICT = 0, 1
LBM_RIGHT_VELOCITY, LBM_RIGHT_PRESSURE = 0, 1
LBM_STEP_MACRO_LAST = 1
LBM_POP_POST_COLLISION, LBM_POP_STREAMED = 0, 1

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_vp = ctypes.c_void_p


class LbmCfg(ctypes.Structure):
    _fields_ = [("nx", c_i64), ("ny", c_i64), ("x0", c_i64), ("nxl", c_i64),
                ("om_p", ctypes.c_double), ("om_m", ctypes.c_double),
                ("dtype", c_i32), ("arith", c_i32), ("right_wall", c_i32), ("device", c_i32)]


class LbmLayout(ctypes.Structure):
    _fields_ = [("elems", c_i64), ("origin", c_i64), ("plane", c_i64), ("pitch", c_i64),
                ("elem_size", c_i64), ("halo", c_i64)]


class LbmPeerInfo(ctypes.Structure):
    _fields_ = [("mem", (ctypes.c_ubyte * 64) * 2), ("flags", ctypes.c_ubyte * 64),
                ("addr", ctypes.c_uint64 * 2), ("flags_addr", ctypes.c_uint64),
                ("pid", c_i64), ("device", c_i64),
                ("x0", c_i64), ("nxl", c_i64), ("origin", c_i64), ("plane", c_i64), ("pitch", c_i64),
                ("elem_size", c_i64)]


class LbmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lbm_b200 error %d: %s" % (code, msg))
        self.code = code


# every symbol declared in include/lbm_b200.h: name -> (restype, argtypes)
SIGNATURES = {
    "lbm_abi_version": (ctypes.c_int, []),
    "lbm_last_error": (ctypes.c_char_p, []),
    "lbm_create": (ctypes.c_int, [ctypes.POINTER(LbmCfg), ctypes.POINTER(c_vp)]),
    "lbm_destroy": (ctypes.c_int, [c_vp]),
    "lbm_get_layout": (ctypes.c_int, [c_vp, ctypes.POINTER(LbmLayout)]),
    "lbm_bind_state": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t]),
    "lbm_state_ptrs": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)]),
    "lbm_set_stream": (ctypes.c_int, [c_vp, c_vp]),
    "lbm_sync": (ctypes.c_int, [c_vp]),
    "lbm_set_right_wall": (ctypes.c_int, [c_vp, c_i32]),
    "lbm_set_populations": (ctypes.c_int, [c_vp, c_vp]),
    "lbm_set_post_collision": (ctypes.c_int, [c_vp, c_vp]),
    "lbm_init_equilibrium": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "lbm_equilibrium": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "lbm_set_links": (ctypes.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp, c_i32]),
    "lbm_wall_row_len": (c_i64, [c_vp]),
    "lbm_set_walls": (ctypes.c_int, [c_vp, c_i64, c_vp]),
    "lbm_set_wall_profiles": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "lbm_set_ramp": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64]),
    "lbm_step": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i64, ctypes.c_uint32]),
    "lbm_step_columns": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, ctypes.c_uint32]),
    "lbm_flip": (ctypes.c_int, [c_vp]),
    "lbm_step2_columns": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i64, c_i64]),
    "lbm_set_temporal_blocking": (ctypes.c_int, [c_vp, c_i32]),
    "lbm_stepn_columns": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i32, c_vp]),
    "lbm_can_stepn": (ctypes.c_int, [c_vp]),
    "lbm_set_temporal_depth": (ctypes.c_int, [c_vp, c_i32]),
    "lbm_set_tuning": (ctypes.c_int, [c_vp, ctypes.c_char_p, c_i64]),
    "lbm_apply_bc": (ctypes.c_int, [c_vp, c_i64]),
    "lbm_get_forces": (ctypes.c_int, [c_vp, c_i64, c_i64, c_vp]),
    "lbm_get_forces_async": (ctypes.c_int, [c_vp, c_i64, c_i64, c_vp]),
    "lbm_forces_now": (ctypes.c_int, [c_vp, c_vp]),
    "lbm_get_populations": (ctypes.c_int, [c_vp, c_i32, c_vp]),
    "lbm_get_macro": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "lbm_get_speed": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "lbm_probe_line": (ctypes.c_int, [c_vp, c_i32, c_i64, c_i64, c_vp]),
    "lbm_peer_export": (ctypes.c_int, [c_vp, ctypes.POINTER(LbmPeerInfo)]),
    "lbm_peer_attach": (ctypes.c_int, [c_vp, c_i32, ctypes.POINTER(LbmPeerInfo)]),
    "lbm_peer_detach": (ctypes.c_int, [c_vp]),
    "lbm_peer_push": (ctypes.c_int, [c_vp, c_i32]),
    "lbm_peer_signal": (ctypes.c_int, [c_vp]),
    "lbm_state_checksum": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_uint64)]),
    "lbm_launch_count": (c_i64, [c_vp]),
    "lbm_resident_plan": (ctypes.c_int, [c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64]),
    "lbm_last_step_ms": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_float)]),
}


def build(force=False):
    """Compile liblbm_b200.so in-tree with the committed Makefile (nvcc, sm_100a)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir)]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "lbm_b200.h"))
    stale = (not os.path.exists(LIB_PATH)
             or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", src_dir, "NVFLAGS=-O3 -std=c++17 -lineinfo "
                               "-gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"])
    return LIB_PATH


def lib():
    """Load the library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LbmError(-2, "CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.lbm_abi_version() != 1:
        raise LbmError(-1, "ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise LbmError(rc, lib().lbm_last_error().decode("utf-8", "replace"))
