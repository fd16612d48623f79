"""GPU parity for the SVF cascade and the polyphase resampler (a13): product vs oracle, bit-exact —
both specs are ours and fix the order of every f32 operation (include/fw_b200.h)."""
import numpy as np
import pytest

import test_svf_resampler_oracle as pins
from conftest import synth
from firewheel_b200 import (AudioGraphConfig, FirewheelGraphCtx, PanNode, ResamplerNode, SumNode, SvfNode, VolumeNode, design_resampler,
                            design_svf)
from helpers import SR, assert_bit_exact, chain, f32, run_planar

pytestmark = pytest.mark.gpu


def svf_coeffs(lib, V, ns, seed):
    rng = np.random.default_rng(seed)
    return np.stack([[design_svf(lib, int(rng.integers(0, 6)), 150.0 * 2.0 ** rng.uniform(0, 6), rng.uniform(0.5, 4.0), SR) for _ in range(ns)]
                     for _ in range(V)]).astype(f32)


def compare(gpu, oracle, build, calls, n_out, bus=False):
    outs = []
    for lib in (gpu, oracle):
        cx, proc, hook = build(lib)
        res = []
        for x, arg in calls:
            if arg is not None:
                hook(arg)
            res.append(run_planar(proc, x, n_out, bus))
        outs.append(res)
        proc.free(); cx.update(); cx.free()
    for i, ((yg, mg), (yo, mo)) in enumerate(zip(*outs)):
        assert_bit_exact(yg, yo, f"call {i}")
        assert mg == mo, f"call {i}"
    return outs[0]


@pytest.mark.parametrize("ns,F,T,bus", [(1, 256, 1024, False), (3, 64, 777, False), (8, 128, 512, True), (2, 100, 1000, True),
                                        (4, 128, 480, False), (5, 256, 2048, True), (2, 64, 4096, False), (3, 32, 96, False)])  # T % 32 == 0: the lanes kernel
def test_svf_in_a_chain(gpu, oracle, ns, F, T, bus):
    V = 37
    co = svf_coeffs(gpu, V, ns, 10 + ns)
    nodes = [(lambda: VolumeNode(80.0), 2, 2), (lambda: SvfNode(ns), 2, 2), (lambda: PanNode(-0.4), 2, 2)]

    def build(lib):
        cx, proc, ids = chain(lib, 2, nodes, voices=V, master_bus=bus, max_block=F, setup=lambda cx, ids: cx.graph.set_svf_coeffs(ids[1], co))
        return cx, proc, lambda pct: cx.graph.set_percent_volume(ids[0], pct)
    x = synth((V, 2, T), 40 + ns)
    compare(gpu, oracle, build, [(x, None), (x[:, :, ::-1].copy(), 30.0), (x, None)], 2, bus)


def test_svf_in_a_dag(gpu, oracle):
    """dry + band-passed wet summed: the generic lowering runs the SVF per channel on pool buffers."""
    V, F, T = 20, 128, 640
    co = svf_coeffs(gpu, V, 2, 3)

    def build(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V))
        g = cx.graph
        sv = g.add_node(2, 2, SvfNode(2)); mix = g.add_node(4, 2, SumNode())
        for c in range(2):
            g.connect(g.graph_in_node(), c, sv, c, False); g.connect(g.graph_in_node(), c, mix, c, False)
            g.connect(sv, c, mix, 2 + c, False); g.connect(mix, c, g.graph_out_node(), c, False)
        g.set_svf_coeffs(sv, co)
        proc = cx.activate(SR, 2, 2, F)
        assert cx.update().graph_error is None, cx.last_error()
        return cx, proc, None
    x = synth((V, 2, T), 6)
    compare(gpu, oracle, build, [(x, None), (x, None)], 2)


def test_resampler_voices(gpu, oracle):
    """Per-voice ratio / resource / loop / seek, several calls (position carried on the device), then gain + master bus."""
    V, F = 24, 128
    tab = design_resampler(oracle, 256, 32, 0.9, 9.0)
    rng = np.random.default_rng(12)
    ratios = 2.0 ** rng.uniform(-1.5, 1.5, V)
    i16 = rng.integers(-32768, 32768, size=(900, 2), dtype=np.int64).astype(np.int16)

    def build(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=2, num_voices=V, master_bus=True))
        g = cx.graph
        rs = g.add_node(0, 2, ResamplerNode(tab)); vol = g.add_node(2, 2, VolumeNode(90.0))
        for c in range(2):
            g.connect(rs, c, vol, c, False); g.connect(vol, c, g.graph_out_node(), c, False)
        proc = cx.activate(SR, 0, 2, F)
        assert cx.update().graph_error is None, cx.last_error()
        res = [g.create_sample_resource(synth((2, 3000), 1)), g.create_sample_resource(synth((1, 1111), 2)), g.create_sample_resource(i16, interleaved=True)]
        for v in range(V):
            g.resampler_set(rs, res[v % 3], ratio=float(ratios[v]), playing=(v % 7 != 6), loop=(v % 2 == 0), voice=v)

        def hook(arg):
            if arg == "seek":
                for v in range(0, V, 3):
                    g.resampler_seek(rs, 17 * v, voice=v)
            elif arg == "retune":
                for v in range(V):
                    g.resampler_set(rs, res[(v + 1) % 3], ratio=float(ratios[::-1][v]), playing=True, loop=(v % 3 == 0), voice=v)
        return cx, proc, hook
    z = lambda T: np.zeros((V, 0, T), f32)
    compare(gpu, oracle, build, [(z(3 * F), None), (z(5 * F + 9), "seek"), (z(2 * F), "retune"), (z(4 * F), None)], 2, True)


@pytest.mark.parametrize("fn", [pins.test_svf_matches_the_bilinear_biquad, pins.test_svf_state_carries_across_blocks_and_calls,
                                pins.test_resampler_table_and_unit_ratio, pins.test_resampler_interpolates_a_sine,
                                pins.test_resampler_loop_and_channel_mapping], ids=lambda f: f.__name__)
def test_pins_hold_on_device(gpu, fn):
    fn(gpu)
