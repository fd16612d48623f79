"""firewheel-b200: B200-native per-block audio-graph DSP path of BillyDM/firewheel.

The product is `firewheel_b200/lib/libfirewheel_b200.so` (CUDA sm_100a, C ABI declared in
include/fw_b200.h). This package is the thin host-side mirror of the reference's graph /
context / processor API on top of that ABI. There is NO CPU fallback: if the library is
missing, `load()` raises.
"""
import os
from pathlib import Path

from . import _capi
from ._capi import FW_ALL_VOICES, FW_ID_DANGLING  # noqa: F401
from .graph import (AddEdgeError, AudioGraphConfig, BiquadNode, CompileGraphError, ConvReverbNode, DelayNode,  # noqa: F401
                    DummyAudioNode, EdgeID, FirewheelGraphCtx, FirewheelProcessor, HardClipNode, MonoToStereoNode,
                    NodeID, PanNode, ResamplerNode, SamplerError, SamplerNode, StereoToMonoNode, SumNode, SvfNode, UpdateStatus, VolumeNode,
                    design_rbj, design_resampler, design_svf)

REPO_ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = Path(__file__).resolve().parent / "lib" / "libfirewheel_b200.so"

_lib = None


def load():
    """Load the CUDA product library. Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(firewheel-b200 has no CPU fallback)")
        _lib = _capi.bind(LIB_PATH, "fw_")
    return _lib
