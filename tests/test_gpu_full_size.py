"""Parity at BASELINE.json's FULL sizes, straight against the oracle (it finishes these in seconds): config 2 (1024 stereo
voices x 256 blocks of 256 frames, gain -> pan -> master bus) and config 3 (4096 stereo voices x 32 blocks of 512 frames,
4-stage biquad + 12000-frame delay). Bit-exact, two calls each so that state carried across calls is covered at size.
Config 4's full size is covered through the FFT property test in test_gpu_parity.py (the direct-form oracle needs minutes)."""
import numpy as np
import pytest

from conftest import synth
from firewheel_b200 import BiquadNode, DelayNode, PanNode, VolumeNode
from helpers import assert_bit_exact, chain, f32, run_planar

pytestmark = pytest.mark.gpu


def test_config2_full_size(gpu, oracle):
    V, F, K = 1024, 256, 256
    rng = np.random.default_rng(2)
    pct = (25 + 75 * rng.random(V)).astype(f32); pan = rng.uniform(-1, 1, V).astype(f32)

    def setup(cx, ids):
        cx.graph.set_percent_volume(ids[0], pct); cx.graph.set_pan(ids[1], pan)
    nodes = [(lambda: VolumeNode(100.0), 2, 2), (lambda: PanNode(0.0), 2, 2)]
    x = synth((V, 2, F * K), 20)
    outs = []
    for lib in (gpu, oracle):
        cx, proc, ids = chain(lib, 2, nodes, voices=V, master_bus=True, max_block=F, setup=setup)
        a = run_planar(proc, x, 2, True)
        cx.graph.set_percent_volume(ids[0], pct[::-1].copy())   # second call ramps every gain
        b = run_planar(proc, x, 2, True)
        outs.append((a, b))
        proc.free(); cx.update(); cx.free()
    for (yg, mg), (yo, mo) in zip(*outs):
        assert_bit_exact(yg, yo, "config 2 full size")
        assert mg == mo


def test_config3_full_size(gpu, oracle):
    from firewheel_b200 import design_rbj
    V, F, K, D = 4096, 512, 32, 12000
    rng = np.random.default_rng(3)
    co = np.zeros((V, 4, 5), f32)
    for v in range(V):
        for s in range(4):
            co[v, s] = design_rbj(gpu, 0 if s % 2 == 0 else 4, 200.0 * 40.0 ** rng.random(), rng.uniform(0.5, 2.0), rng.uniform(-3, 3), 48000)
    nodes = [(lambda: BiquadNode(4), 2, 2), (lambda: DelayNode(D), 2, 2)]
    x = synth((V, 2, F * K), 30)
    outs = []
    for lib in (gpu, oracle):
        cx, proc, ids = chain(lib, 2, nodes, voices=V, max_block=F, setup=lambda cx, ids: cx.graph.set_biquad_coeffs(ids[0], co))
        a = run_planar(proc, x, 2)[0]
        b = run_planar(proc, x, 2)[0]      # the ring (12000 frames) and the filter state carry over
        outs.append((a, b))
        proc.free(); cx.update(); cx.free()
    for yg, yo in zip(*outs):
        assert_bit_exact(yg, yo, "config 3 full size")
