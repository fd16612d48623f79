"""Block-stamped control and chunked calls on the product, against the CPU oracle: parameter stores, sampler transport and filter
retunes placed at chosen blocks of a long call (the reference's per-block polling, processor.rs:214 / volume.rs:92 / sampler.rs:331),
and calls longer than the reserved stretch (fw_graph_config::max_call_frames) processed as chunks."""
import numpy as np
import pytest

import timed_scenarios as ts
from conftest import synth
from helpers import assert_bit_exact, run_planar

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scenario,kw", [(ts.gain_pan_timed, dict(bus=False)), (ts.gain_pan_timed, dict(bus=True, V=70)), (ts.sampler_timed, {}), (ts.filters_timed, {})])
def test_stamped_stores_match_the_oracle(gpu, oracle, scenario, kw):
    a, b = scenario(gpu, timed=True, **kw), scenario(oracle, timed=True, **kw)
    for i, ((ya, ma), (yb, mb)) in enumerate(zip(a, b)):
        assert_bit_exact(ya, yb, f"{scenario.__name__} call {i}")
        assert ma == mb


@pytest.mark.parametrize("bus", [False, True])
def test_calls_longer_than_the_reserve_are_chunked(gpu, oracle, bus):
    """max_call_frames = 3 blocks, calls of 12 blocks: four chunks per call, timed stores on top"""
    a = ts.gain_pan_timed(gpu, V=9, bus=bus, timed=True, max_call_frames=3 * ts.F)
    b = ts.gain_pan_timed(oracle, V=9, bus=bus, timed=True)
    for i, ((ya, ma), (yb, mb)) in enumerate(zip(a, b)):
        assert_bit_exact(ya, yb, f"call {i}")
        assert ma == mb


def test_chunked_temporal_and_interleaved_entry(gpu, oracle):
    """biquad -> delay chain and the interleaved entry point with a 2-block reserve and a ragged 7.5-block call"""
    from firewheel_b200 import AudioGraphConfig, BiquadNode, DelayNode, FirewheelGraphCtx, design_rbj
    V, F, T = 5, 64, 7 * 64 + 32
    x = synth((V, T, 2), 3)  # interleaved [voice][frame][ch]
    outs = []
    for lib in (gpu, oracle):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V, max_call_frames=2 * F))
        g = cx.graph
        bq, dl = g.add_node(2, 2, BiquadNode(2)), g.add_node(2, 2, DelayNode(160))
        for c in range(2):
            g.connect(g.graph_in_node(), c, bq, c, False); g.connect(bq, c, dl, c, False); g.connect(dl, c, g.graph_out_node(), c, False)
        g.set_biquad_coeffs(bq, np.stack([[design_rbj(lib, 0, 900.0 + 70 * v, 0.8, 0.0, 48000), design_rbj(lib, 4, 2500.0, 1.1, 4.0, 48000)] for v in range(V)]).astype(np.float32))
        proc = cx.activate(48000, 2, 2, F)
        assert cx.update().graph_error is None, cx.last_error()
        res = []
        for _ in range(2):
            y = np.zeros((V, T, 2), np.float32)
            assert proc.process_interleaved(x, y, 2, 2, T) == 0
            res.append(y)
        proc.free(); cx.update(); cx.free()
        outs.append(res)
    for i, (yg, yo) in enumerate(zip(*outs)):
        assert_bit_exact(yg, yo, f"call {i}")
