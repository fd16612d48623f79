"""Scenarios for block-stamped control (ctx_set_event_block): built once, run on any implementation of the C ABI."""
import numpy as np

from conftest import synth
from firewheel_b200 import (AudioGraphConfig, BiquadNode, FirewheelGraphCtx, PanNode, SamplerNode, SumNode, SvfNode, VolumeNode, design_rbj,
                            design_svf)
from helpers import run_planar

F = 64


def chain_gain_pan(lib, V, bus, max_call_frames=0):
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V, master_bus=bus, max_call_frames=max_call_frames))
    g = cx.graph
    vol, pan = g.add_node(2, 2, VolumeNode(100.0)), g.add_node(2, 2, PanNode(0.0))
    for c in range(2):
        g.connect(g.graph_in_node(), c, vol, c, False); g.connect(vol, c, pan, c, False); g.connect(pan, c, g.graph_out_node(), c, False)
    proc = cx.activate(48000, 2, 2, F)
    assert cx.update().graph_error is None, cx.last_error()
    return cx, proc, vol, pan


def gain_pan_timed(lib, V=6, bus=False, timed=True, max_call_frames=0):
    """12 blocks per call, three calls. Stores land at blocks 0, 3, 7 and 14 (= block 2 of the following call). timed=False performs
    the same stores by splitting the calls by hand — the reference's way."""
    cx, proc, vol, pan = chain_gain_pan(lib, V, bus, max_call_frames)
    g = cx.graph
    K = 12
    x = synth((V, 2, 3 * K * F), 41)
    stores = {0: lambda: g.set_percent_volume(vol, 40.0, voice=1), 3: lambda: (g.set_percent_volume(vol, 10.0), g.set_pan(pan, -0.5, voice=2)),
              7: lambda: g.set_pan(pan, np.linspace(-1, 1, V).astype(np.float32)), 14: lambda: g.set_percent_volume(vol, 150.0, voice=V - 1)}
    outs = []
    if timed:
        for b, fn in stores.items():
            g.set_event_block(b); fn()
        g.set_event_block(0)
        for k in range(3):
            outs.append(run_planar(proc, np.ascontiguousarray(x[:, :, k * K * F:(k + 1) * K * F]), 2, bus))
    else:
        cuts = sorted(set(list(stores) + [0, K, 2 * K, 3 * K]))
        pieces = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            if a in stores:
                stores[a]()
            pieces.append(run_planar(proc, np.ascontiguousarray(x[:, :, a * F:b * F]), 2, bus))
        ys = np.concatenate([p[0] for p in pieces], axis=-1)
        for k in range(3):
            last = [p for (a, b), p in zip(zip(cuts[:-1], cuts[1:]), pieces) if b <= (k + 1) * K][-1]
            outs.append((np.ascontiguousarray(ys[..., k * K * F:(k + 1) * K * F]), last[1]))
    proc.free(); cx.update(); cx.free()
    return outs


def sampler_timed(lib, V=5, timed=True):
    """sampler -> gain -> bus; transport messages (play / pause / seek / loop / volume) stamped with block offsets inside one long call"""
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=2, num_voices=V, master_bus=True))
    g = cx.graph
    smp, vol = g.add_node(0, 2, SamplerNode(100.0)), g.add_node(2, 2, VolumeNode(80.0))
    for c in range(2):
        g.connect(smp, c, vol, c, False); g.connect(vol, c, g.graph_out_node(), c, False)
    proc = cx.activate(48000, 0, 2, F)
    assert cx.update().graph_error is None, cx.last_error()
    res = [g.create_sample_resource(synth((2, 700 + 130 * r), 90 + r)) for r in range(3)]
    K = 16
    ev = {0: lambda: ([g.sampler_set_sample(smp, res[v % 3], True, voice=v) for v in range(V)], g.sampler_play(smp)),
          2: lambda: g.sampler_pause(smp, voice=1),
          5: lambda: (g.sampler_set_loop_range(smp, "full", voice=0), g.sampler_set_playhead(smp, 0.005, voice=2), g.sampler_play(smp, voice=1)),
          9: lambda: (g.sampler_set_percent_volume(smp, 30.0, voice=3), g.sampler_stop(smp, voice=4)),
          13: lambda: g.sampler_play(smp, voice=4)}
    x0 = np.zeros((V, 0, K * F), np.float32)
    if timed:
        for b, fn in ev.items():
            g.set_event_block(b); fn()
        g.set_event_block(0)
        out = run_planar(proc, x0, 2, True)
    else:
        cuts = sorted(set(list(ev) + [0, K]))
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            ev[a]()
            parts.append(run_planar(proc, np.zeros((V, 0, (b - a) * F), np.float32), 2, True))
        out = (np.concatenate([p[0] for p in parts], axis=-1), parts[-1][1])
    proc.free(); cx.update(); cx.free()
    return [out]


def filters_timed(lib, V=4, timed=True):
    """biquad -> svf with coefficient retunes stamped at blocks 4 and 9 (a per-block coefficient ramp over blocks 9..11 for one voice)"""
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V))
    g = cx.graph
    bq, sv = g.add_node(2, 2, BiquadNode(2)), g.add_node(2, 2, SvfNode(1))
    for c in range(2):
        g.connect(g.graph_in_node(), c, bq, c, False); g.connect(bq, c, sv, c, False); g.connect(sv, c, g.graph_out_node(), c, False)
    co = np.stack([[design_rbj(lib, 0, 500.0 + 100 * v, 0.9, 0.0, 48000), design_rbj(lib, 4, 3000.0, 1.2, 3.0, 48000)] for v in range(V)]).astype(np.float32)
    g.set_biquad_coeffs(bq, co)
    g.set_svf_coeffs(sv, np.stack([[design_svf(lib, 0, 1500.0, 0.8, 48000)] for _ in range(V)]).astype(np.float32))
    proc = cx.activate(48000, 2, 2, F)
    assert cx.update().graph_error is None, cx.last_error()
    K = 14
    x = synth((V, 2, K * F), 17)
    ramp = [design_rbj(lib, 0, 800.0 + 400.0 * i, 0.7, 0.0, 48000) for i in range(3)]
    ev = {4: lambda: g.set_svf_coeffs(sv, np.asarray([design_svf(lib, 2, 2500.0, 1.1, 48000)]), voice=1),
          9: lambda: g.set_biquad_coeffs(bq, ramp[0], voice=2, stage=0), 10: lambda: g.set_biquad_coeffs(bq, ramp[1], voice=2, stage=0),
          11: lambda: g.set_biquad_coeffs(bq, ramp[2], voice=2, stage=0)}
    if timed:
        for b, fn in ev.items():
            g.set_event_block(b); fn()
        g.set_event_block(0)
        out = run_planar(proc, x, 2)
    else:
        cuts = sorted(set(list(ev) + [0, K]))
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            if a in ev:
                ev[a]()
            parts.append(run_planar(proc, np.ascontiguousarray(x[:, :, a * F:b * F]), 2))
        out = (np.concatenate([p[0] for p in parts], axis=-1), parts[-1][1])
    proc.free(); cx.update(); cx.free()
    return [out]
