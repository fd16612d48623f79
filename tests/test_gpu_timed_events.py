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


def test_steady_calls_replay_a_captured_cuda_graph(gpu, oracle):
    """a DAG on the generic lowering (dry + gained path into a 4-port sum, then pan): identical consecutive calls are captured into a
    CUDA graph at their second sight and replayed; a parameter store in between does not invalidate the capture, results stay bit-exact"""
    from firewheel_b200 import AudioGraphConfig, FirewheelGraphCtx, PanNode, SumNode, VolumeNode
    V, F, K = 33, 64, 3
    outs, replays = [], 0
    for lib in (gpu, oracle):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V))
        g = cx.graph
        vol, mix, pan = g.add_node(2, 2, VolumeNode(60.0)), g.add_node(4, 2, SumNode()), g.add_node(2, 2, PanNode(0.2))
        for c in range(2):
            g.connect(g.graph_in_node(), c, vol, c, False); g.connect(g.graph_in_node(), c, mix, c, False)
            g.connect(vol, c, mix, 2 + c, False); g.connect(mix, c, pan, c, False); g.connect(pan, c, g.graph_out_node(), c, False)
        proc = cx.activate(48000, 2, 2, F)
        assert cx.update().graph_error is None, cx.last_error()
        res = []
        for i in range(7):
            if i == 4:
                g.set_percent_volume(vol, 35.0, voice=5)
            res.append(run_planar(proc, synth((V, 2, K * F), 100 + i), 2))
        if lib is gpu:
            replays = proc.graph_replays()
        proc.free(); cx.update(); cx.free()
        outs.append(res)
    for i, ((yg, mg), (yo, mo)) in enumerate(zip(*outs)):
        assert_bit_exact(yg, yo, f"call {i}")
        assert mg == mo
    assert replays >= 4, replays


def test_port_count_change_on_a_live_temporal_node_reactivates_it(gpu):
    """set_num_inputs / set_num_outputs on an activated BiquadNode (graph.rs:315-393): its per-channel device state no longer fits, so the
    node is activated again with the new counts (fresh state) instead of running out of bounds"""
    from firewheel_b200 import AudioGraphConfig, BiquadNode, FirewheelGraphCtx, design_rbj
    V, F = 6, 64
    co = np.stack([[design_rbj(gpu, 0, 1000.0 + 100 * v, 0.8, 0.0, 48000)] for v in range(V)]).astype(np.float32)
    x = synth((V, 4, 4 * F), 8)

    def build(ports):
        cx = FirewheelGraphCtx(gpu, AudioGraphConfig(num_graph_inputs=4, num_graph_outputs=4, num_voices=V))
        g = cx.graph
        bq = g.add_node(ports, ports, BiquadNode(1))
        g.set_biquad_coeffs(bq, co)
        for c in range(4):
            if c < ports:
                g.connect(g.graph_in_node(), c, bq, c, False); g.connect(bq, c, g.graph_out_node(), c, False)
            else:
                g.connect(g.graph_in_node(), c, g.graph_out_node(), c, False)
        proc = cx.activate(48000, 4, 4, F)
        assert cx.update().graph_error is None, cx.last_error()
        return cx, proc, bq
    cx, proc, bq = build(2)
    run_planar(proc, x, 4)
    g = cx.graph
    for c in (2, 3):
        assert g.disconnect(g.graph_in_node(), c, g.graph_out_node(), c)
    g.set_num_inputs(bq, 4); g.set_num_outputs(bq, 4)
    for c in (2, 3):
        g.connect(g.graph_in_node(), c, bq, c, False); g.connect(bq, c, g.graph_out_node(), c, False)
    assert cx.update().graph_error is None, cx.last_error()
    run_planar(proc, x, 4)          # Q11: the first block after the swap reads zero inputs
    y_live, _ = run_planar(proc, x, 4)
    proc.free(); cx.update(); cx.free()
    cx2, proc2, _ = build(4)
    run_planar(proc2, np.concatenate([np.zeros((V, 4, F), np.float32), x[:, :, F:]], axis=2), 4)   # what the re-activated node saw in the swap call
    y_fresh, _ = run_planar(proc2, x, 4)
    proc2.free(); cx2.update(); cx2.free()
    assert_bit_exact(y_live, y_fresh, "re-activated 4-channel biquad")
