"""Small fixed scenarios shared by tests/golden/make_golden.py (which renders them with the CPU oracle and commits the
vectors), tests/test_golden.py (oracle == committed vectors, CPU) and its GPU half (product == committed vectors, without
the oracle in the loop). Inputs are stored in the fixtures too: nothing depends on a random generator staying stable."""
import numpy as np

from firewheel_b200 import (AudioGraphConfig, BiquadNode, ConvReverbNode, DelayNode, FirewheelGraphCtx, HardClipNode, MonoToStereoNode,
                            PanNode, ResamplerNode, SamplerNode, StereoToMonoNode, SumNode, SvfNode, VolumeNode, design_rbj,
                            design_resampler, design_svf)

f32 = np.float32
SR = 48000


def _activate(cx, n_in, n_out, F):
    proc = cx.activate(SR, n_in, n_out, F)
    st = cx.update()
    assert proc is not None and st.graph_error is None, (st, cx.last_error())
    return proc


def _calls(proc, inputs, n_out, bus, hooks=None):
    outs = []
    for i, x in enumerate(inputs):
        if hooks and hooks.get(i):
            hooks[i]()
        V, n_in, T = x.shape
        out = np.full((n_out, T) if bus else (V, n_out, T), np.nan, f32)
        rc, mask = proc.process_planar(np.ascontiguousarray(x), out, n_in, n_out, T)
        assert rc == 0
        outs.append((out, mask))
    return outs


def gain_pan_bus(lib, inputs):
    """gain -> pan -> master bus over 7 voices; the second call ramps every gain, the third mutes two voices."""
    V = 7
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V, master_bus=True))
    g = cx.graph
    vol, pan = g.add_node(2, 2, VolumeNode(100.0)), g.add_node(2, 2, PanNode(0.0))
    for c in range(2):
        g.connect(g.graph_in_node(), c, vol, c, False); g.connect(vol, c, pan, c, False); g.connect(pan, c, g.graph_out_node(), c, False)
    g.set_percent_volume(vol, np.linspace(30, 100, V).astype(f32)); g.set_pan(pan, np.linspace(-1, 1, V).astype(f32))
    proc = _activate(cx, 2, 2, 64)
    hooks = {1: lambda: g.set_percent_volume(vol, np.linspace(100, 20, V).astype(f32)),
             2: lambda: (g.set_percent_volume(vol, 0.0, voice=1), g.set_percent_volume(vol, 0.0, voice=4))}
    outs = _calls(proc, inputs, 2, True, hooks)
    proc.free(); cx.update(); cx.free()
    return outs


def dag_with_sums(lib, inputs):
    """fan-out, 2-port and 6-port SumNodes with unconnected ports, clip, mono<->stereo: the generic lowering's territory."""
    V = 4
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=3, num_graph_outputs=3, num_voices=V))
    g = cx.graph
    gi, go = g.graph_in_node(), g.graph_out_node()
    s2m, m2s = g.add_node(2, 1, StereoToMonoNode()), g.add_node(1, 2, MonoToStereoNode())
    clip, vol = g.add_node(2, 2, HardClipNode(-6.0)), g.add_node(1, 1, VolumeNode(60.0))
    sum2, sum6 = g.add_node(4, 2, SumNode()), g.add_node(6, 1, SumNode())
    g.connect(gi, 0, s2m, 0, False); g.connect(gi, 1, s2m, 1, False); g.connect(s2m, 0, m2s, 0, False)
    g.connect(m2s, 0, clip, 0, False); g.connect(m2s, 1, clip, 1, False)
    g.connect(clip, 0, sum2, 0, False); g.connect(clip, 1, sum2, 1, False); g.connect(gi, 2, sum2, 2, False)  # port 1 right unconnected
    g.connect(gi, 2, vol, 0, False)
    g.connect(vol, 0, sum6, 1, False); g.connect(gi, 0, sum6, 2, False); g.connect(gi, 1, sum6, 4, False); g.connect(s2m, 0, sum6, 5, False)
    g.connect(sum2, 0, go, 0, False); g.connect(sum2, 1, go, 1, False); g.connect(sum6, 0, go, 2, False)
    proc = _activate(cx, 3, 3, 100)
    hooks = {1: lambda: g.set_percent_volume(vol, 0.0), 2: lambda: g.set_percent_volume(vol, 90.0)}
    outs = _calls(proc, inputs, 3, False, hooks)
    proc.free(); cx.update(); cx.free()
    return outs


def filters(lib, inputs):
    """gain -> 3-stage biquad -> 2-stage SVF -> 160-frame delay; state carried across three calls."""
    V = 3
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V))
    g = cx.graph
    nodes = [g.add_node(2, 2, VolumeNode(80.0)), g.add_node(2, 2, BiquadNode(3)), g.add_node(2, 2, SvfNode(2)), g.add_node(2, 2, DelayNode(160))]
    prev = g.graph_in_node()
    for n in nodes + [g.graph_out_node()]:
        for c in range(2):
            g.connect(prev, c, n, c, False)
        prev = n
    bq = np.stack([[design_rbj(lib, t, fc, 0.9, 3.0, SR) for t, fc in ((0, 900.0 + 300 * v), (4, 2500.0), (1, 120.0))] for v in range(V)]).astype(f32)
    sv = np.stack([[design_svf(lib, t, fc, 1.5, SR) for t, fc in ((1, 700.0 + 100 * v), (5, 3000.0))] for v in range(V)]).astype(f32)
    g.set_biquad_coeffs(nodes[1], bq); g.set_svf_coeffs(nodes[2], sv)
    proc = _activate(cx, 2, 2, 128)
    outs = _calls(proc, inputs, 2, False)
    proc.free(); cx.update(); cx.free()
    return outs


def players(lib, inputs, extra):
    """A SamplerNode (i16 interleaved resource, looping / one-shot per voice, transport messages between calls) and a
    polyphase ResamplerNode (f32 planar resource, per-voice ratios) summed into the master bus."""
    V = 5
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=2, num_voices=V, master_bus=True))
    g = cx.graph
    smp, rs, mix = g.add_node(0, 2, SamplerNode(90.0)), g.add_node(0, 2, ResamplerNode(extra["table"])), g.add_node(4, 2, SumNode())
    for c in range(2):
        g.connect(smp, c, mix, c, False); g.connect(rs, c, mix, 2 + c, False); g.connect(mix, c, g.graph_out_node(), c, False)
    proc = _activate(cx, 0, 2, 64)
    r_i16, r_f32 = g.create_sample_resource(extra["pcm_i16"], interleaved=True), g.create_sample_resource(extra["pcm_f32"])
    g.sampler_set_sample(smp, r_i16, True)
    for v in range(V):
        if v % 2 == 0:
            g.sampler_set_loop_range(smp, "full", voice=v)
        g.resampler_set(rs, r_f32, step_q32=int(extra["steps"][v]), playing=True, loop=(v % 2 == 1), voice=v)
    g.sampler_play(smp)
    hooks = {1: lambda: (g.sampler_pause(smp, voice=0), g.sampler_set_playhead(smp, 0.004, voice=2), g.resampler_seek(rs, 40, voice=3)),
             2: lambda: (g.sampler_play(smp, voice=0), g.sampler_set_percent_volume(smp, 30.0))}
    outs = _calls(proc, inputs, 2, True, hooks)
    proc.free(); cx.update(); cx.free()
    return outs


def reverb(lib, inputs, extra):
    """200-tap stereo FIR reverb, history carried across calls (tolerance 1e-5: bf16 operands, fp32 tensor-core sums)."""
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=3))
    g = cx.graph
    rv = g.add_node(2, 2, ConvReverbNode(extra["ir"]))
    for c in range(2):
        g.connect(g.graph_in_node(), c, rv, c, False); g.connect(rv, c, g.graph_out_node(), c, False)
    proc = _activate(cx, 2, 2, 128)
    outs = _calls(proc, inputs, 2, False)
    proc.free(); cx.update(); cx.free()
    return outs


# name -> (function, tolerance or None for bit-exact)
SCENARIOS = {"gain_pan_bus": (gain_pan_bus, None), "dag_with_sums": (dag_with_sums, None), "filters": (filters, None),
             "players": (players, None), "reverb": (reverb, 1e-5)}


def make_inputs(name, lib):
    """Inputs + extra data of a scenario (only used by the generator; the tests read them back from the fixture)."""
    rng = np.random.default_rng(hash(name) % 2 ** 32 if False else sum(map(ord, name)))
    u = lambda shape: (rng.integers(0, 1 << 24, size=shape).astype(f32) * f32(2.0 ** -23) - f32(1.0)).astype(f32)
    if name == "gain_pan_bus":
        return [u((7, 2, 256)), u((7, 2, 256)), u((7, 2, 200))], {}
    if name == "dag_with_sums":
        x = [u((4, 3, 300)), u((4, 3, 300)), u((4, 3, 128))]
        x[0][:, :, 50:60] = -0.0
        return x, {}
    if name == "filters":
        return [u((3, 2, 384)), u((3, 2, 100)), u((3, 2, 512))], {}
    if name == "players":
        z = lambda T: np.zeros((5, 0, T), f32)
        extra = {"table": design_resampler(lib, 64, 16, 0.9, 8.0), "pcm_i16": rng.integers(-32768, 32768, size=(300, 2)).astype(np.int16),
                 "pcm_f32": u((2, 500)), "steps": np.array([int(r * 2 ** 32) for r in (0.5, 0.75, 1.0, 1.37, 2.2)], np.uint64)}
        return [z(256), z(192), z(320)], extra
    if name == "reverb":
        L = 200
        ir = (rng.standard_normal((2, L)) * np.exp(-6.9 * np.arange(L) / L)).astype(f32)
        return [u((3, 2, 256)), u((3, 2, 256))], {"ir": ir}
    raise KeyError(name)


def run(lib, name, inputs, extra):
    fn, _ = SCENARIOS[name]
    return fn(lib, inputs, extra) if extra else fn(lib, inputs)
