"""GPU parity for voice graphs that are NOT linear chains: the generic per-node lowering (DESIGN.md "generic
lowering") against the CPU oracle, which runs the reference's own executor over the reference's buffer assignment
(schedule.rs:289-342, compiler.rs:302-412). Bar: bit-exact outputs and identical silence masks; graphs holding a
ConvReverbNode are held to the reverb's 1e-5 normalised tolerance instead."""
import numpy as np
import pytest

from conftest import synth
from firewheel_b200 import (AudioGraphConfig, BiquadNode, ConvReverbNode, DelayNode, FirewheelGraphCtx, HardClipNode, MonoToStereoNode,
                            PanNode, StereoToMonoNode, SumNode, VolumeNode, design_rbj)
from helpers import SR, assert_bit_exact, f32, run_planar

pytestmark = pytest.mark.gpu


def activate(cx, n_in, n_out, F):
    proc = cx.activate(SR, n_in, n_out, F)
    assert proc is not None
    st = cx.update()
    assert st.kind == "Active" and st.graph_error is None, (st, cx.last_error())
    return proc


def compare(gpu, oracle, build, calls, n_out, bus=False, tol=None):
    """build(lib) -> (cx, proc, retune); calls: [(x, retune_arg or None)]."""
    outs = []
    for lib in (gpu, oracle):
        cx, proc, retune = build(lib)
        res = []
        for x, arg in calls:
            if arg is not None:
                retune(arg)
            res.append(run_planar(proc, x, n_out, bus))
        outs.append(res)
        proc.free(); cx.update(); cx.free()
    for i, ((yg, mg), (yo, mo)) in enumerate(zip(*outs)):
        if tol is None:
            assert_bit_exact(yg, yo, f"call {i}")
            assert mg == mo, f"call {i}: silence mask {mg:#x} != {mo:#x}"
        else:
            err = float(np.max(np.abs(yg.astype(np.float64) - yo)) / max(np.max(np.abs(yo)), 1e-30))
            assert err <= tol, (i, err)
    return outs[0]


# ---- hand-built topologies -----------------------------------------------------------------------------------
def test_four_to_two_sum_node(gpu, oracle):
    """The 2-port stereo SumNode straight between graph_in(4) and graph_out(2) (sum.rs:69-81)."""
    V, F, T = 37, 128, 640

    def build(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=4, num_graph_outputs=2, num_voices=V))
        g = cx.graph
        s = g.add_node(4, 2, SumNode())
        for i in range(4):
            g.connect(g.graph_in_node(), i, s, i, False)
        for c in range(2):
            g.connect(s, c, g.graph_out_node(), c, False)
        return cx, activate(cx, 4, 2, F), None
    compare(gpu, oracle, build, [(synth((V, 4, T), 1), None), (synth((V, 4, 100), 2), None)], 2)


@pytest.mark.parametrize("bus", [False, True])
def test_dry_wet_split(gpu, oracle, bus):
    """graph_in fans out to a dry gain and a filtered + delayed wet path; a SumNode mixes them (fan-out, reconvergence,
    a temporal node inside a DAG, buffer reuse across branches)."""
    V, F, T = 70, 256, 1024
    rng = np.random.default_rng(4)
    co = np.stack([[design_rbj(gpu, s % 2, 300.0 * (s + 1 + v % 5), 0.8, 0.0, SR) for s in range(2)] for v in range(V)]).astype(f32)
    dry_pct = (20 + 80 * rng.random(V)).astype(f32)

    def build(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V, master_bus=bus))
        g = cx.graph
        dry = g.add_node(2, 2, VolumeNode(100.0))
        bq = g.add_node(2, 2, BiquadNode(2))
        dl = g.add_node(2, 2, DelayNode(333))
        wet = g.add_node(2, 2, VolumeNode(50.0))
        mix = g.add_node(4, 2, SumNode())
        pan = g.add_node(2, 2, PanNode(0.2))
        for c in range(2):
            g.connect(g.graph_in_node(), c, dry, c, False)
            g.connect(g.graph_in_node(), c, bq, c, False)
            g.connect(bq, c, dl, c, False)
            g.connect(dl, c, wet, c, False)
            g.connect(dry, c, mix, c, False)
            g.connect(wet, c, mix, 2 + c, False)
            g.connect(mix, c, pan, c, False)
            g.connect(pan, c, g.graph_out_node(), c, False)
        g.set_biquad_coeffs(bq, co)
        g.set_percent_volume(dry, dry_pct)
        proc = activate(cx, 2, 2, F)
        return cx, proc, lambda pct: g.set_percent_volume(wet, pct)
    x = synth((V, 2, T), 9)
    compare(gpu, oracle, build, [(x, None), (x, 90.0), (x[:, :, :300].copy(), 0.0), (x, None), (x, 35.0)], 2, bus)


def test_wide_sum_skips_silent_ports(gpu, oracle):
    """6-port mono SumNode (sum.rs:111-133): two ports unconnected (cleared + flagged), one fed by a muted gain. The
    reference skips flagged ports; with the first port flagged it still starts from that port's zeros."""
    V, F, T = 33, 64, 512

    def build(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=3, num_graph_outputs=1, num_voices=V))
        g = cx.graph
        s = g.add_node(6, 1, SumNode())
        mute = g.add_node(1, 1, VolumeNode(100.0))
        g.connect(g.graph_in_node(), 0, mute, 0, False)
        g.connect(mute, 0, s, 1, False)           # port 0 unconnected, port 1 = gain
        g.connect(g.graph_in_node(), 1, s, 2, False)
        g.connect(g.graph_in_node(), 2, s, 3, False)
        g.connect(g.graph_in_node(), 0, s, 5, False)  # port 4 unconnected
        g.connect(s, 0, g.graph_out_node(), 0, False)
        proc = activate(cx, 3, 1, F)
        return cx, proc, lambda pct: g.set_percent_volume(mute, pct)
    x = synth((V, 3, T), 21)
    pcts = np.where(np.arange(V) % 3 == 0, 0.0, 80.0).astype(f32)
    compare(gpu, oracle, build, [(x, None), (x, pcts), (x, None), (x, 100.0)], 1)


def test_negative_gains_and_zero_signs(gpu, oracle):
    """Flagged-silent channels can hold -0.0 (a cleared channel times a negative pan gain); the reference's non-fused node
    bodies then write +0.0 where arithmetic would keep -0.0 (hard_clip.rs:78-82, stereo_to_mono.rs:41-47, sum.rs:52-56)."""
    V, F, T = 9, 32, 128

    def build(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=1, num_graph_outputs=4, num_voices=V))
        g = cx.graph
        mute = g.add_node(2, 2, VolumeNode(0.0))       # outputs cleared + flagged
        pan = g.add_node(2, 2, PanNode(0.0))
        clip = g.add_node(2, 2, HardClipNode(-3.0))
        s2m = g.add_node(2, 1, StereoToMonoNode())
        m2s = g.add_node(1, 2, MonoToStereoNode())
        sm = g.add_node(4, 2, SumNode())
        vol3 = g.add_node(3, 3, VolumeNode(70.0))
        g.connect(g.graph_in_node(), 0, mute, 0, False)
        g.connect(g.graph_in_node(), 0, pan, 0, False)  # pan input 1 unconnected: flagged, * negative gain -> -0.0
        for c in range(2):
            g.connect(pan, c, clip, c, False)
        g.connect(pan, 1, s2m, 0, False)
        g.connect(mute, 1, s2m, 1, False)
        g.connect(pan, 1, m2s, 0, False)
        g.connect(pan, 1, sm, 0, False); g.connect(pan, 1, sm, 1, False); g.connect(mute, 0, sm, 2, False); g.connect(pan, 1, sm, 3, False)
        g.connect(pan, 1, vol3, 0, False); g.connect(clip, 0, vol3, 1, False)
        g.connect(clip, 1, g.graph_out_node(), 0, False)
        g.connect(s2m, 0, g.graph_out_node(), 1, False)
        g.connect(sm, 1, g.graph_out_node(), 2, False)
        g.connect(vol3, 0, g.graph_out_node(), 3, False)
        g.set_pan_gains(pan, 0.7, -0.6)
        return cx, activate(cx, 1, 4, F), None
    x = synth((V, 1, T), 3)
    compare(gpu, oracle, build, [(x, None), (x, None)], 4)


def test_reverb_send_bus(gpu, oracle):
    """Dry path + reverb send summed into the master bus: the reverb runs per channel on pool buffers."""
    V, F, T, L = 40, 256, 768, 200
    rng = np.random.default_rng(11)
    h = (rng.standard_normal((2, L)) * np.exp(-6.9 * np.arange(L) / L)).astype(f32)

    def build(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V, master_bus=True))
        g = cx.graph
        send = g.add_node(2, 2, VolumeNode(40.0))
        rv = g.add_node(2, 2, ConvReverbNode(h))
        mix = g.add_node(4, 2, SumNode())
        for c in range(2):
            g.connect(g.graph_in_node(), c, send, c, False)
            g.connect(send, c, rv, c, False)
            g.connect(g.graph_in_node(), c, mix, c, False)
            g.connect(rv, c, mix, 2 + c, False)
            g.connect(mix, c, g.graph_out_node(), c, False)
        return cx, activate(cx, 2, 2, F), None
    x = synth((V, 2, T), 14)
    compare(gpu, oracle, build, [(x, None), (x, None), (x[:, :, :100].copy(), None), (x, None)], 2, True, tol=1e-5)


@pytest.mark.parametrize("bus,F,mcf", [(False, 128, 0), (True, 128, 256), (False, 100, 300), (True, 64, 0)])
def test_fused_runs_and_caller_row_reads(gpu, oracle, bus, F, mcf):
    """The second step of the generic lowering (runtime.cu fuse_generic): a run gain -> pan -> gain fed by graph_in's channels (2, 0) is one
    launch reading the caller's rows; an SVF and a delay fed by graph_in read them too (the pool copy of the inputs disappears); the run
    gain -> pan behind the SumNode rides in graph_out's launch (strided copy, or the bus stage). Odd call lengths, chunked calls
    (max_call_frames), ramps through the fused programs, and a schedule swap in between (Q11: the first block after it reads zero inputs,
    which every reader of the caller's rows has to do itself now)."""
    from firewheel_b200 import SvfNode, design_svf
    V = 37
    rng = np.random.default_rng(11)
    kf = np.stack([[design_svf(gpu, 0, 900.0 + 40 * v, 0.9, SR)] for v in range(V)]).astype(f32)
    pct = (30 + 90 * rng.random(V)).astype(f32)

    def build(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=3, num_graph_outputs=2, num_voices=V, master_bus=bus, max_call_frames=mcf))
        g = cx.graph
        gin, gout = g.graph_in_node(), g.graph_out_node()
        a1, a2, a3 = g.add_node(2, 2, VolumeNode(80.0)), g.add_node(2, 2, PanNode(-0.3)), g.add_node(2, 2, VolumeNode(120.0))
        svf, dl = g.add_node(2, 2, SvfNode(1)), g.add_node(1, 1, DelayNode(160))
        mix = g.add_node(5, 1, SumNode())
        up = g.add_node(1, 2, MonoToStereoNode())
        b1, b2 = g.add_node(2, 2, VolumeNode(70.0)), g.add_node(2, 2, PanNode(0.4))
        for c, port in enumerate((2, 0)):
            g.connect(gin, port, a1, c, False)
        for c, port in enumerate((1, 2)):
            g.connect(gin, port, svf, c, False)
        g.connect(gin, 1, dl, 0, False)
        for c in range(2):
            g.connect(a1, c, a2, c, False); g.connect(a2, c, a3, c, False)
            g.connect(a3, c, mix, c, False); g.connect(svf, c, mix, 2 + c, False)
            g.connect(b1, c, b2, c, False); g.connect(b2, c, gout, c, False)
        g.connect(dl, 0, mix, 4, False)
        g.connect(mix, 0, up, 0, False)
        for c in range(2):
            g.connect(up, c, b1, c, False)
        g.set_svf_coeffs(svf, kf)
        g.set_percent_volume(a1, pct)
        proc = activate(cx, 3, 2, F)
        state = {}

        def retune(arg):
            if arg == "swap":  # splice a HardClipNode in front of the mix's mono output and take it out again two calls later
                state["clip"] = g.add_node(1, 1, HardClipNode(-6.0))
                assert g.disconnect(mix, 0, up, 0)
                g.connect(mix, 0, state["clip"], 0, False); g.connect(state["clip"], 0, up, 0, False)
                assert cx.update().graph_error is None
            elif arg == "unswap":
                g.remove_node(state.pop("clip")); g.connect(mix, 0, up, 0, False)
                assert cx.update().graph_error is None
            else:
                g.set_percent_volume(a3, float(arg)); g.set_pan(b2, float(arg) / 100.0 - 0.5); g.set_percent_volume(b1, 100.0 - float(arg) / 2)
        return cx, proc, retune
    T = 5 * F + 17
    x = synth((V, 3, T), 21)
    x[:, :, T // 3: T // 3 + 5] = -0.0
    compare(gpu, oracle, build, [(x, None), (x, 40.0), (x, "swap"), (x[:, :, : 2 * F + 3].copy(), 90.0), (x, "unswap"), (x, None)], 2, bus)



# ---- random DAGs -----------------------------------------------------------------------------------------------
KINDS = ["vol1", "vol2", "vol3", "pan", "clip1", "clip2", "clip3", "m2s", "s2m", "sum2x1", "sum2x2", "sum3x1", "sum4x1", "sum6x1", "sum5x2", "sum1x2",
         "biquad1", "biquad2", "delay1", "delay2"]


def random_dag(lib, seed, V, F, bus, n_nodes, temporal):
    rng = np.random.default_rng(seed)
    n_in = int(rng.integers(1, 4))
    n_out = int(rng.integers(1, 3)) if bus else int(rng.integers(1, 5))
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=n_in, num_graph_outputs=n_out, num_voices=V, master_bus=bus))
    g = cx.graph
    sources = [(g.graph_in_node(), p) for p in range(n_in)]
    vols, pans = [], []
    for _ in range(n_nodes):
        kinds = KINDS if temporal else KINDS[:16]
        k = kinds[int(rng.integers(len(kinds)))]
        if k.startswith("vol"):
            ch = int(k[3]); nid = g.add_node(ch, ch, VolumeNode(float(rng.choice([0.0, 50.0, 100.0, 120.0])))); ni = no = ch; vols.append(nid)
        elif k == "pan":
            nid = g.add_node(2, 2, PanNode(float(rng.uniform(-1, 1)))); ni = no = 2; pans.append(nid)
        elif k.startswith("clip"):
            ch = int(k[4]); nid = g.add_node(ch, ch, HardClipNode(float(rng.uniform(-12, 0)))); ni = no = ch
        elif k == "m2s":
            nid = g.add_node(1, 2, MonoToStereoNode()); ni, no = 1, 2
        elif k == "s2m":
            nid = g.add_node(2, 1, StereoToMonoNode()); ni, no = 2, 1
        elif k.startswith("sum"):
            ports, ch = int(k[3]), int(k[5]); ni, no = ports * ch, ch; nid = g.add_node(ni, no, SumNode())
        elif k.startswith("biquad"):
            ch = int(k[6]); ns = int(rng.integers(1, 4)); nid = g.add_node(ch, ch, BiquadNode(ns)); ni = no = ch
            co = np.stack([[design_rbj(lib, int(rng.integers(0, 2)), float(rng.uniform(200, 4000)), 0.9, 0.0, SR) for _ in range(ns)] for _ in range(V)]).astype(f32)
            g.set_biquad_coeffs(nid, co)
        else:
            ch = int(k[5]); nid = g.add_node(ch, ch, DelayNode(int(rng.choice([0, 17, 160, 480])))); ni = no = ch
        for p in range(ni):
            if rng.random() < 0.85:
                src = sources[int(rng.integers(max(0, len(sources) - 8), len(sources)))]
                g.connect(src[0], src[1], nid, p, False)
        sources += [(nid, p) for p in range(no)]
    for p in range(n_out):
        if rng.random() < 0.9:
            src = sources[int(rng.integers(max(0, len(sources) - 5), len(sources)))]
            g.connect(src[0], src[1], g.graph_out_node(), p, False)
    for nid in pans:  # negative gains make -0.0 out of flagged channels
        for v in range(0, V, 3):
            g.set_pan_gains(nid, float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), voice=v)

    def retune(seed2):
        r2 = np.random.default_rng(seed2)
        for nid in vols:
            g.set_percent_volume(nid, r2.choice(np.array([0.0, 30.0, 100.0], f32), size=V).astype(f32))
        for nid in pans:
            g.set_pan(nid, r2.uniform(-1, 1, V).astype(f32))
    return cx, n_in, n_out, retune


@pytest.mark.parametrize("seed", range(24))
def test_random_dags(gpu, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    V = int(rng.choice([1, 5, 33, 70]))
    F = int(rng.choice([32, 100, 256]))
    bus = bool(seed % 3 == 0)
    temporal = seed % 2 == 1
    n_nodes = int(rng.integers(3, 14))
    shape = {}

    def build(lib):
        cx, n_in, n_out, retune = random_dag(lib, seed, V, F, bus, n_nodes, temporal)
        shape["io"] = (n_in, n_out)
        return cx, activate(cx, n_in, n_out, F), retune
    # learn the channel counts from one dry build on the oracle (cheap), then drive both
    cx, proc, _ = build(oracle); proc.free(); cx.update(); cx.free()
    n_in, n_out = shape["io"]
    T = int(rng.choice([F, 3 * F, 4 * F + 17]))
    x = synth((V, n_in, T), 50 + seed)
    x[:, :, T // 2: T // 2 + 7] = -0.0
    calls = [(x, None), (x, seed + 1), (x[:, :, : max(1, T // 3)].copy(), None), (x, seed + 2), (x, None)]
    compare(gpu, oracle, build, calls, n_out, bus)
