"""GPU parity: the CUDA product against the CPU oracle on the same seeded inputs, through the C ABI.
Bar: bit-exact (f32 bit patterns, including the sign of zero) and identical silence masks."""
import numpy as np
import pytest

from conftest import synth
from firewheel_b200 import HardClipNode, MonoToStereoNode, PanNode, StereoToMonoNode, SumNode, VolumeNode
from helpers import assert_bit_exact, chain, f32, run_planar

pytestmark = pytest.mark.gpu


def both(gpu, oracle, build, calls, n_out, master_bus=False):
    """build(lib) -> (cx, proc, ids); calls: list of (x, pre) where pre(cx, ids) may retune parameters."""
    outs = []
    for lib in (gpu, oracle):
        cx, proc, ids = build(lib)
        res = []
        for x, pre in calls:
            if pre:
                pre(cx, ids)
            res.append(run_planar(proc, x, n_out, master_bus))
        outs.append(res)
        proc.free()
        cx.update()
        cx.free()
    for i, ((yg, mg), (yo, mo)) in enumerate(zip(*outs)):
        assert_bit_exact(yg, yo, f"call {i}")
        assert mg == mo, f"call {i}: silence mask {mg:#x} != {mo:#x}"
    return outs[0]


def vol_pan(V, master_bus, max_block, pct, pan):
    def setup(cx, ids):
        cx.graph.set_percent_volume(ids[0], pct)
        cx.graph.set_pan(ids[1], pan)
    return lambda lib: chain(lib, 2, [(lambda: VolumeNode(100.0), 2, 2), (lambda: PanNode(0.0), 2, 2)], voices=V,
                             master_bus=master_bus, max_block=max_block, setup=setup)


@pytest.mark.parametrize("V,bus", [(1, False), (3, False), (70, False), (1, True), (2, True), (7, True), (64, True), (65, True), (130, True), (1024, True)])
def test_gain_pan_chain_constant(gpu, oracle, V, bus):
    rng = np.random.default_rng(V)
    pct = rng.uniform(25, 100, V).astype(f32)
    pan = rng.uniform(-1, 1, V).astype(f32)
    T = 256 * 3
    x = synth((V, 2, T), 100 + V)
    both(gpu, oracle, vol_pan(V, bus, 256, pct, pan), [(x, None), (x[:, :, :256].copy(), None)], 2, bus)


@pytest.mark.parametrize("F,T", [(256, 256 * 2 + 4), (256, 777), (100, 1000), (64, 640), (128, 131), (4, 64), (255, 1020)])
def test_block_shapes_and_ragged_tails(gpu, oracle, F, T):
    V = 5
    pct = np.linspace(30, 150, V).astype(f32)
    pan = np.linspace(-1, 1, V).astype(f32)
    x = synth((V, 2, T), F * 7 + T)
    both(gpu, oracle, vol_pan(V, True, F, pct, pan), [(x, None)], 2, True)
    both(gpu, oracle, vol_pan(V, False, F, pct, pan), [(x, None)], 2, False)


def test_ramps_settle_stall_and_mute(gpu, oracle):
    V, F = 9, 256
    pct0 = np.array([100, 100, 100, 50, 0, 100, 200, 100, 30], f32)
    pct1 = np.array([0, 50, 100, 100, 100, 25, 100, 0.05, 30], f32)   # to-mute, down, unchanged, up (stalls), from-mute, ...
    pan0 = np.zeros(V, f32)
    pan1 = np.linspace(-1, 1, V).astype(f32)
    x = -np.abs(synth((V, 2, F * 30), 7)) - f32(0.01)

    def retune(cx, ids):
        cx.graph.set_percent_volume(ids[0], pct1)
        cx.graph.set_pan(ids[1], pan1)

    def back(cx, ids):
        cx.graph.set_percent_volume(ids[0], pct0)

    calls = [(x[:, :, :F * 2].copy(), None), (x, retune), (x[:, :, :F * 4].copy(), None), (x, back), (x[:, :, :F].copy(), None)]
    both(gpu, oracle, vol_pan(V, False, F, pct0, pan0), calls, 2, False)
    both(gpu, oracle, vol_pan(V, True, F, pct0, pan0), calls, 2, True)


def test_all_muted_bus_is_silent(gpu, oracle):
    V = 4
    (y, m), = both(gpu, oracle, vol_pan(V, True, 256, np.zeros(V, f32), np.zeros(V, f32)), [(synth((V, 2, 512), 3), None)], 2, True)
    assert m == 0b11 and not np.any(np.signbit(y))


def test_clip_mono_stereo_chains(gpu, oracle):
    x1 = synth((6, 1, 900), 11)
    both(gpu, oracle, lambda lib: chain(lib, 1, [(MonoToStereoNode, 1, 2), (lambda: VolumeNode(70.0), 2, 2), (lambda: HardClipNode(-9.0), 2, 2)], voices=6), [(x1, None)], 2)
    x2 = synth((6, 2, 900), 12)
    both(gpu, oracle, lambda lib: chain(lib, 2, [(lambda: HardClipNode(-3.0), 2, 2), (StereoToMonoNode, 2, 1), (lambda: VolumeNode(120.0), 1, 1)], voices=6, master_bus=True), [(x2, None)], 1, True)
    both(gpu, oracle, lambda lib: chain(lib, 2, [(SumNode, 2, 2)], voices=3), [(x2[:3].copy(), None)], 2)  # 1-port sum == copy
    both(gpu, oracle, lambda lib: chain(lib, 2, [], voices=3), [(x2[:3].copy(), None)], 2)  # graph_in -> graph_out


def test_interleaved_entry_point(gpu, oracle):
    V, T = 3, 700
    x = synth((V, T, 2), 21)
    res = []
    for lib in (gpu, oracle):
        for bus in (False, True):
            cx, proc, ids = vol_pan(V, bus, 256, np.array([50, 0, 150], f32), np.array([-0.5, 0, 0.5], f32))(lib)
            out = np.full((T, 2) if bus else (V, T, 2), np.nan, f32)
            assert proc.process_interleaved(x, out, 2, 2, T) == 0
            res.append(out)
            proc.free(); cx.update(); cx.free()
    assert_bit_exact(res[0], res[2], "interleaved per-voice")
    assert_bit_exact(res[1], res[3], "interleaved bus")


def test_schedule_swap_first_block_reads_zero_inputs(gpu, oracle):  # Q11
    V, F = 3, 128
    x = synth((V, 2, 4 * F), 31)
    outs = []
    for lib in (gpu, oracle):
        cx, proc, (vol,) = chain(lib, 2, [(lambda: VolumeNode(100.0), 2, 2)], voices=V, max_block=F)
        g = cx.graph
        y0 = run_planar(proc, x, 2)
        clip = g.add_node(2, 2, HardClipNode(-12.0))
        for c in range(2):
            assert g.disconnect(vol, c, g.graph_out_node(), c)
            g.connect(vol, c, clip, c, False)
            g.connect(clip, c, g.graph_out_node(), c, False)
        assert cx.update().graph_error is None
        y1 = run_planar(proc, x, 2)
        y2 = run_planar(proc, x, 2)
        g.remove_node(clip)
        for c in range(2):
            g.connect(vol, c, g.graph_out_node(), c, False)
        assert cx.update().graph_error is None
        y3 = run_planar(proc, x, 2)
        outs.append((y0, y1, y2, y3))
        proc.free(); cx.update(); cx.free()
    for (a, ma), (b, mb) in zip(*outs):
        assert_bit_exact(a, b)
        assert ma == mb
    assert np.all(outs[0][1][0][:, :, :F] == 0) and np.any(outs[0][1][0][:, :, F:] != 0)


def test_config2_full_size_blocks(gpu, oracle):
    """BASELINE config[1]: 1024 stereo voices, gain -> pan -> sum, 256-frame blocks (8 blocks here)."""
    V, F, K = 1024, 256, 8
    rng = np.random.default_rng(2)
    pct = (25 + 75 * rng.random(V)).astype(f32)
    pan = rng.uniform(-1, 1, V).astype(f32)
    x = synth((V, 2, F * K), 99)
    (y, m), = both(gpu, oracle, vol_pan(V, True, F, pct, pan), [(x, None)], 2, True)
    assert m == 0 and np.all(np.isfinite(y))


def test_unsupported_topology_fails_loudly(gpu):
    """A DummyAudioNode with outputs inside the graph leaves stale buffer contents behind in the reference (dummy.rs:34-41);
    the device path refuses the graph instead of inventing a result — and never falls back to a CPU path."""
    from firewheel_b200 import AudioGraphConfig, DummyAudioNode, FirewheelGraphCtx
    cx = FirewheelGraphCtx(gpu, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2))
    g = cx.graph
    d = g.add_node(2, 2, DummyAudioNode())
    for c in range(2):
        g.connect(g.graph_in_node(), c, d, c, False)
        g.connect(d, c, g.graph_out_node(), c, False)
    proc = cx.activate(48000, 2, 2, 256)
    st = cx.update()
    assert st.graph_error is not None and st.graph_error.kind == "UnsupportedOnDevice"
    out = np.full((1, 2, 256), np.nan, f32)
    rc, _ = proc.process_planar(synth((1, 2, 256), 1), out, 2, 2, 256)
    assert rc == 0 and np.all(out == 0)  # no schedule => silence (processor.rs:86-89)
    proc.free()


# ---- temporal nodes: biquad cascade + delay line (spec ours; parity = bit-exact vs the f32 oracle) ------------
def biquad_coeffs(lib, V, ns, seed):
    from firewheel_b200 import design_rbj
    rng = np.random.default_rng(seed)
    k = np.zeros((V, ns, 5), f32)
    for v in range(V):
        for s in range(ns):
            k[v, s] = design_rbj(lib, [0, 4, 1, 5][s % 4], 200.0 * (2.0 ** rng.uniform(0, 5.3)), rng.uniform(0.5, 2.0), rng.uniform(-6, 6), 48000)
    return k


def temporal_chain(gpu, V, ns, D, F, pre=(), post=(), bus=False):
    from firewheel_b200 import BiquadNode, DelayNode
    nodes = list(pre)
    if ns is not None:
        nodes.append((lambda: BiquadNode(ns), 2, 2))
    if D is not None:
        nodes.append((lambda: DelayNode(D), 2, 2))
    nodes += list(post)
    bq_index = len(pre) if ns is not None else None
    coeffs = biquad_coeffs(gpu, V, ns, 7 * V + (ns or 0)) if ns else None

    def setup(cx, ids):
        if ns:
            cx.graph.set_biquad_coeffs(ids[bq_index], coeffs)
    return lambda lib: chain(lib, 2, nodes, voices=V, master_bus=bus, max_block=F, setup=setup)


@pytest.mark.parametrize("ns,D,F,T", [(4, 12000, 512, 2048), (4, 96, 256, 1024), (1, None, 128, 512), (8, 128, 64, 640), (2, 3200, 512, 1024),
                                      (None, 160, 256, 512), (None, 0, 256, 512),            # delay only; delay(0) == copy
                                      (4, 300, 256, 1000), (3, 50, 100, 777), (4, 12000, 512, 36)])  # generic path: ragged T / D
def test_biquad_delay(gpu, oracle, ns, D, F, T):
    V = 19
    x = synth((V, 2, T), 500 + T + (ns or 0))
    calls = [(x, None), (x[:, :, ::-1].copy(), None), (x, None)]  # three calls: filter state and ring carry across calls
    both(gpu, oracle, temporal_chain(gpu, V, ns, D, F), calls, 2)


def test_mixed_pointwise_and_temporal_stages(gpu, oracle):
    V, F, T = 70, 256, 1024
    pre = [(lambda: VolumeNode(80.0), 2, 2)]
    post = [(lambda: PanNode(0.25), 2, 2), (lambda: HardClipNode(-1.0), 2, 2)]
    x = synth((V, 2, T), 77)
    both(gpu, oracle, temporal_chain(gpu, V, 4, 480, F, pre, post, bus=True), [(x, None), (x, None)], 2, True)
    both(gpu, oracle, temporal_chain(gpu, V, 2, None, F, pre, post, bus=False), [(x, None)], 2, False)
    # temporal stage last with a master bus: an empty pointwise stage carries the bus
    both(gpu, oracle, temporal_chain(gpu, V, 4, 96, F, pre, (), bus=True), [(x, None), (x, None)], 2, True)


def test_config3_shape(gpu, oracle):
    """BASELINE config[2] shape at reduced voice count: 4-stage biquad cascade + 12000-frame delay, 512-frame blocks."""
    V, F, K = 256, 512, 26  # 13312 frames > D: the ring wraps inside one call
    x = synth((V, 2, F * K), 3)
    both(gpu, oracle, temporal_chain(gpu, V, 4, 12000, F), [(x, None), (x[:, :, :F * 2].copy(), None)], 2)


# ---- FIR convolutional reverb (spec ours): bf16 operands, fp32 tensor-core accumulation -----------------------
def reverb_ir(L, ch, seed):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal((ch, L)) * np.exp(-6.9 * np.arange(L) / L)
    h /= np.sqrt((h ** 2).sum(axis=1, keepdims=True))
    return h.astype(f32)


def norm_max_err(got, ref):
    return float(np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64))) / np.max(np.abs(ref)))


@pytest.mark.parametrize("L,V,T,F", [(200, 5, 512, 256), (1000, 130, 1024, 512), (4096, 3, 777, 128), (64, 2, 256, 256)])
def test_conv_reverb_vs_oracle(gpu, oracle, L, V, T, F):
    """Tolerance (north_star): normalised max error <= 1e-5 against the oracle (bf16-rounded x and h, f64 accumulate)."""
    from firewheel_b200 import ConvReverbNode
    ir = reverb_ir(L, 2, L)
    x = synth((V, 2, T), 900 + L)
    outs = []
    for lib in (gpu, oracle):
        cx, proc, _ = chain(lib, 2, [(lambda: ConvReverbNode(ir), 2, 2)], voices=V, max_block=F)
        ys = [run_planar(proc, x, 2)[0], run_planar(proc, x[:, :, ::-1].copy(), 2)[0]]  # second call: history carries over
        outs.append(ys)
        proc.free(); cx.update(); cx.free()
    for yg, yo in zip(*outs):
        assert norm_max_err(yg, yo) <= 1e-5, norm_max_err(yg, yo)


def test_conv_reverb_full_length_vs_fft(gpu, oracle):
    """BASELINE config[3] IR length (48000 taps, stereo IR) on a few voices; reference = f64 FFT convolution of the
    bf16-rounded operands (the direct-form oracle would take minutes at this length)."""
    import scipy.signal
    from firewheel_b200 import ConvReverbNode
    L, V, T = 48000, 6, 2048
    ir = reverb_ir(L, 2, 7)
    x = synth((V, 2, T * 2), 31)
    cx, proc, _ = chain(gpu, 2, [(lambda: ConvReverbNode(ir), 2, 2)], voices=V, max_block=512)
    y = np.concatenate([run_planar(proc, np.ascontiguousarray(x[:, :, :T]), 2)[0], run_planar(proc, np.ascontiguousarray(x[:, :, T:]), 2)[0]], axis=2)
    proc.free(); cx.update(); cx.free()
    rb = np.vectorize(oracle.bf16_round, otypes=[f32])
    xb, hb = rb(x).astype(np.float64), rb(ir).astype(np.float64)
    for v in range(V):
        for c in range(2):
            ref = scipy.signal.fftconvolve(xb[v, c], hb[c])[: 2 * T]
            assert norm_max_err(y[v, c], ref) <= 1e-5, (v, c, norm_max_err(y[v, c], ref))


@pytest.mark.parametrize("V,n_tiles,L", [(300, 19, 4100), (129, 41, 4100), (520, 13, 4100)])
def test_conv_reverb_many_tiles_vs_fft(gpu, oracle, V, n_tiles, L):
    """More output tiles than CTA pairs on the chip (76, 82, 78 tiles on 74 pairs): full waves of whole tiles plus a short tail wave whose tiles
    are split along K between many pairs with a fix-up add (reverb.cu). Reference = f64 FFT convolution of the bf16-rounded operands."""
    import scipy.signal
    from firewheel_b200 import AudioGraphConfig, ConvReverbNode, FirewheelGraphCtx
    T = n_tiles * 256
    ir = reverb_ir(L, 2, 5)
    x = synth((V, 2, T), 77)
    cx = FirewheelGraphCtx(gpu, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V, max_call_frames=T))
    g = cx.graph
    rv = g.add_node(2, 2, ConvReverbNode(ir))
    for c in range(2):
        g.connect(g.graph_in_node(), c, rv, c, False); g.connect(rv, c, g.graph_out_node(), c, False)
    proc = cx.activate(48000, 2, 2, 256)
    assert cx.update().graph_error is None, cx.last_error()
    y1, _ = run_planar(proc, x, 2)
    y2, _ = run_planar(proc, x, 2)   # second call: history carries over
    proc.free(); cx.update(); cx.free()
    rb = np.vectorize(oracle.bf16_round, otypes=[f32])
    hb = rb(ir).astype(np.float64)
    for v in list(range(0, V, max(1, V // 7))) + [V - 1]:
        xb = rb(np.concatenate([x[v], x[v]], axis=1)).astype(np.float64)
        for c in range(2):
            ref = scipy.signal.fftconvolve(xb[c], hb[c])[: 2 * T]
            assert norm_max_err(y1[v, c], ref[:T]) <= 1e-5, (v, c, norm_max_err(y1[v, c], ref[:T]))
            assert norm_max_err(y2[v, c], ref[T:]) <= 1e-5, (v, c, norm_max_err(y2[v, c], ref[T:]))


def test_reverb_in_a_mixed_chain_with_bus(gpu, oracle):
    from firewheel_b200 import ConvReverbNode
    ir = reverb_ir(300, 2, 3)
    V, T = 70, 512
    x = synth((V, 2, T), 5)
    nodes = [(lambda: VolumeNode(70.0), 2, 2), (lambda: ConvReverbNode(ir), 2, 2), (lambda: PanNode(-0.3), 2, 2)]
    outs = []
    for lib in (gpu, oracle):
        cx, proc, _ = chain(lib, 2, nodes, voices=V, master_bus=True, max_block=256)
        outs.append(run_planar(proc, x, 2, True)[0])
        proc.free(); cx.update(); cx.free()
    assert norm_max_err(outs[0], outs[1]) <= 1e-5


def test_config5_chain_shape(gpu, oracle):
    """BASELINE config[4] voice graph at reduced size: gain -> pan -> 4-stage biquad -> FIR reverb -> master bus.
    Everything up to the reverb is bit-exact, so the bus differs from the oracle only by the reverb's fp32 accumulation."""
    from firewheel_b200 import BiquadNode, ConvReverbNode
    V, F, T, L = 96, 256, 1024, 600
    ir = reverb_ir(L, 2, 17)
    co = biquad_coeffs(gpu, V, 4, 23)
    rng = np.random.default_rng(8)
    pct = (25 + 75 * rng.random(V)).astype(f32); pan = rng.uniform(-1, 1, V).astype(f32)
    nodes = [(lambda: VolumeNode(100.0), 2, 2), (lambda: PanNode(0.0), 2, 2), (lambda: BiquadNode(4), 2, 2), (lambda: ConvReverbNode(ir), 2, 2)]

    def setup(cx, ids):
        cx.graph.set_percent_volume(ids[0], pct); cx.graph.set_pan(ids[1], pan); cx.graph.set_biquad_coeffs(ids[2], co)
    x = synth((V, 2, T), 12)
    outs = []
    for lib in (gpu, oracle):
        cx, proc, _ = chain(lib, 2, nodes, voices=V, master_bus=True, max_block=F, setup=setup)
        outs.append([run_planar(proc, x, 2, True)[0], run_planar(proc, x, 2, True)[0]])
        proc.free(); cx.update(); cx.free()
    for yg, yo in zip(*outs):
        assert norm_max_err(yg, yo) <= 1e-5, norm_max_err(yg, yo)
