"""Loader for the test plugin (tests/plugins/fir1_plugin.cu): a third-party node behind include/fw_b200.h's fw_node_vtable."""
import ctypes
from pathlib import Path

import numpy as np

PLUGIN = Path(__file__).resolve().parent / "plugins" / "libfw_test_plugin.so"
RULE_NONE, RULE_PASSTHROUGH, RULE_ALL_IF_ALL_INPUTS = 0, 1, 2
_lib = None


def load():
    global _lib
    if _lib is None:
        if not PLUGIN.exists():
            raise RuntimeError(f"{PLUGIN} is missing: run __graft_entry__.build()")
        _lib = ctypes.CDLL(str(PLUGIN))
        _lib.fw_test_plugin_vtable.restype = ctypes.c_void_p
        _lib.fw_test_plugin_new.restype = ctypes.c_void_p
        _lib.fw_test_plugin_new.argtypes = [ctypes.c_float, ctypes.c_uint32, ctypes.c_int]
        _lib.fw_test_plugin_counters.argtypes = [ctypes.c_void_p]
    return _lib


def new_node(k0=0.5, rule=RULE_ALL_IF_ALL_INPUTS, fail_activate=False):
    """(vtable pointer, node pointer) for AudioGraph.add_custom_node"""
    p = load()
    return p.fw_test_plugin_vtable(), p.fw_test_plugin_new(k0, rule, 1 if fail_activate else 0)


def counters():
    """activate, deactivate, drop_node, drop_processor, update"""
    out = np.zeros(5, np.uint32)
    load().fw_test_plugin_counters(out.ctypes.data)
    return dict(zip(("activate", "deactivate", "drop_node", "drop_processor", "update"), (int(v) for v in out)))


def reset_counters():
    load().fw_test_plugin_reset_counters()


def fir1_reference(x, k0, state=None):
    """numpy restatement of the plugin's arithmetic: x [V][C][T] -> y, new state [V][C]"""
    V, C, T = x.shape
    st = np.zeros((V, C), np.float32) if state is None else state
    kv = (np.float32(k0) + np.float32(0.01) * np.arange(V, dtype=np.float32)).astype(np.float32)
    xp = np.concatenate([st[:, :, None], x[:, :, :-1]], axis=2)
    y = (x - (kv[:, None, None] * xp).astype(np.float32)).astype(np.float32)
    return y, x[:, :, -1].copy()
