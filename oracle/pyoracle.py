"""Loader for the CPU oracle (TEST INFRASTRUCTURE). Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libfw_oracle.so"

_u32, _u64, _f32, _i32, _vp = C.c_uint32, C.c_uint64, C.c_float, C.c_int, C.c_void_p
EXTRA = {
    "smoother_run": (_i32, [_f32, _u32, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    "percent_volume_to_raw_gain": (_f32, [_f32]),
    "db_to_gain_clamped_neg_100_db": (_f32, [_f32]),
    "bf16_round": (_f32, [_f32]),
    "pan_to_gains": (None, [_f32, C.POINTER(_f32), C.POINTER(_f32)]),
    "silence_mask_new_all_silent": (_u64, [_u32]),
    "silence_mask_query": (_i32, [_u64, _u32, _i32]),
    "deinterleave": (_u64, [_vp, _u32, _u32, _vp, _u32, _i32]),
    "interleave": (None, [_vp, _u32, _u32, _vp, _u32, _i32, _u64, _i32]),
}

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)


def load():
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        sys.path.insert(0, str(HERE.parent))
        from firewheel_b200 import _capi
        _lib = _capi.bind(LIB, "fwo_", EXTRA)
    return _lib
