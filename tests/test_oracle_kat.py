"""Known-answer tests that pin the CPU oracle's numerics.

The reference holds no numeric tests, so the anchors are (i) values derived by hand from the
reference formulas (file:line cited), (ii) numpy float32 re-evaluations of the same single-op
sequences, (iii) scipy cross-checks for the nodes whose spec is ours (parity unpinned).
"""
import ctypes as C

import numpy as np
import pytest
import scipy.signal

from conftest import synth
from firewheel_b200 import (AudioGraphConfig, BiquadNode, ConvReverbNode, DelayNode, FirewheelGraphCtx, HardClipNode,
                            MonoToStereoNode, PanNode, StereoToMonoNode, SumNode, VolumeNode, design_rbj)

f32 = np.float32
SR, F = 48000, 256


# ---- helpers ----------------------------------------------------------------------------------
def chain(lib, n_ch, nodes, voices=1, master_bus=False, max_block=F, n_out=None, setup=None):
    """graph_in(n_ch) -> nodes[0] -> ... -> graph_out, port i to port i. nodes: [(node, n_in, n_out)]."""
    n_out = n_out if n_out is not None else (nodes[-1][2] if nodes else n_ch)
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=n_ch, num_graph_outputs=n_out, num_voices=voices, master_bus=master_bus))
    g = cx.graph
    prev, prev_w, ids = g.graph_in_node(), n_ch, []
    for node, ni, no in nodes:
        nid = g.add_node(ni, no, node)
        for p in range(min(prev_w, ni)):
            g.connect(prev, p, nid, p, False)
        prev, prev_w = nid, no
        ids.append(nid)
    for p in range(min(prev_w, n_out)):
        g.connect(prev, p, g.graph_out_node(), p, False)
    if setup:
        setup(cx, ids)  # parameters set before activation start un-smoothed (volume.rs:67-75)
    proc = cx.activate(SR, n_ch, n_out, max_block)
    st = cx.update()
    assert st.kind == "Active" and st.graph_error is None, st
    return cx, proc, ids


def run_planar(proc, x, n_out, master_bus=False):
    V, n_in, T = x.shape
    out = np.full((n_out, T) if master_bus else (V, n_out, T), np.nan, dtype=f32)
    rc, mask = proc.process_planar(np.ascontiguousarray(x), out, n_in, n_out, T)
    assert rc == 0
    return out, mask


def smoother_run(lib, initial, targets, frames, max_block=F, sr=SR):
    n = len(targets)
    t = np.asarray(targets, dtype=f32)
    fr = np.asarray(frames if np.ndim(frames) else [frames] * n, dtype=np.uint32)
    curves = np.zeros((n, max_block), dtype=f32)
    lens, stat, ab = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(2, f32)
    final = lib.smoother_run(f32(initial), sr, max_block, t.ctypes.data, fr.ctypes.data, n, curves.ctypes.data,
                             lens.ctypes.data, stat.ctypes.data, ab.ctypes.data)
    return curves, lens, stat, ab, final


# ---- scalar helpers -----------------------------------------------------------------------------
def test_percent_volume_to_raw_gain(oracle):  # range.rs:32-35
    assert oracle.percent_volume_to_raw_gain(50.0) == 0.25
    assert oracle.percent_volume_to_raw_gain(100.0) == 1.0
    assert oracle.percent_volume_to_raw_gain(200.0) == 4.0
    assert oracle.percent_volume_to_raw_gain(-5.0) == 0.0
    n = f32(37.5) * (f32(1.0) / f32(100.0))
    assert oracle.percent_volume_to_raw_gain(37.5) == float(n * n)


def test_db_to_gain_clamped(oracle):  # util.rs:21-27
    assert oracle.db_to_gain_clamped_neg_100_db(-100.0) == 0.0
    assert oracle.db_to_gain_clamped_neg_100_db(0.0) == 1.0
    assert abs(oracle.db_to_gain_clamped_neg_100_db(-6.0) - 10 ** (-0.3)) < 1e-7


def test_silence_mask(oracle):  # silence_mask.rs:23-72
    assert oracle.silence_mask_new_all_silent(0) == 0
    assert oracle.silence_mask_new_all_silent(2) == 0b11
    assert oracle.silence_mask_new_all_silent(64) == 2 ** 64 - 1
    assert oracle.silence_mask_query(0b10, 2, 0) == 1   # any
    assert oracle.silence_mask_query(0b10, 2, 1) == 0   # all
    assert oracle.silence_mask_query(0b11, 2, 1) == 1
    assert oracle.silence_mask_query(0b111, 2, 1) == 1  # bits above n ignored
    assert oracle.silence_mask_query(0b10, 1, 2) == 1   # is_channel_silent(1)


def test_bf16_round(oracle):
    assert oracle.bf16_round(1.0) == 1.0
    assert oracle.bf16_round(1.00390625) == 1.0          # 1 + 2^-8: tie -> even
    assert oracle.bf16_round(1.01171875) == 1.015625     # 1 + 3*2^-8: tie -> even (up)
    assert oracle.bf16_round(-0.3) == float(np.frombuffer(np.array([0xBE9A0000], np.uint32).tobytes(), f32)[0])


# ---- ParamSmoother (smoother.rs:93-205) ---------------------------------------------------------
def test_smoother_coefficients(oracle):
    _, _, _, ab, _ = smoother_run(oracle, 1.0, [1.0], F)
    import math
    arg = f32(-1.0) / ((f32(10.0) / f32(1000.0)) * f32(SR))  # smoother.rs:20,99
    b = f32(math.exp(float(arg)))  # correctly rounded f32 exp == glibc expf (numpy's SIMD exp is 1 ulp off here)
    assert float(arg) == float(f32(-1.0) / f32(480.0))
    assert ab[1] == b and ab[0] == f32(1.0) - b


def test_smoother_inactive_returns_full_buffer(oracle):  # Q1
    curves, lens, stat, _, _ = smoother_run(oracle, 0.5, [0.5, 0.5], [100, 7])
    assert list(lens) == [F, F] and list(stat) == [0, 0]
    assert np.all(curves == f32(0.5))


def test_smoother_recurrence_bit_exact_and_closed_form(oracle):
    curves, lens, stat, ab, _ = smoother_run(oracle, 0.25, [1.0, 1.0, 1.0], F)
    a, b = ab
    t = f32(1.0) * a
    y, exp = f32(0.25), []
    for _ in range(3 * F):
        y = f32(t + f32(y * b))
        exp.append(y)
    got = curves.reshape(-1)
    assert np.array_equal(got, np.array(exp, dtype=f32))
    assert list(lens) == [F] * 3 and list(stat) == [1, 1, 1]
    n = np.arange(1, 3 * F + 1, dtype=np.float64)
    closed = 1.0 + (0.25 - 1.0) * np.float64(b) ** n
    assert np.max(np.abs(got - closed)) < 2e-5


def test_smoother_settle_quirks(oracle):  # Q2, Q3
    nblk = 40
    curves, lens, stat, _, final = smoother_run(oracle, 0.0, [0.5] * nblk, F)
    k = int(np.argmax(stat == 2))           # first Deactivating block
    assert stat[k] == 2 and np.all(stat[:k] == 1)
    # Q3: the block that settles is *replaced* by the target over the whole curve
    assert np.all(curves[k] == f32(0.5))
    assert abs(0.5 - curves[k - 1][-1]) < 2e-5 and curves[k - 1][0] != 0.5
    # Q2: Deactivating -> Inactive never happens; later blocks return the full constant buffer
    assert np.all(stat[k:] == 2) and final == 2
    assert np.all(lens[k + 1:] == F) and np.all(curves[k + 1:] == f32(0.5))
    # expected settle point: 0.5 * b^n < 1e-5  =>  n = ln(5e4) * 480 = 5193
    assert 5000 < k * F < 5500


def test_smoother_can_stall_and_never_settle(oracle):  # Q10 (found while pinning): f32 fixed point outside epsilon
    # Near 1.0 the per-sample increment a*(1-y) drops below half an ulp of y once |1-y| ~ 1.4e-5 > settle_epsilon,
    # so y = t + y*b reproduces y exactly: the smoother stays Active forever on a constant curve.
    curves, lens, stat, ab, final = smoother_run(oracle, 0.25, [1.0] * 200, F)
    assert np.all(stat == 1) and final == 1
    y = curves[-1][-1]
    assert 1e-5 < abs(1.0 - y) < 2e-5
    assert np.all(curves[-1] == y)
    a, b = ab
    assert f32(f32(1.0) * a + f32(y * b)) == y  # the fixed point


def test_smoother_retarget_mid_ramp(oracle):
    curves, _, stat, ab, _ = smoother_run(oracle, 0.0, [1.0, 0.5, 0.5], F)
    a, b = ab
    y, exp = f32(0.0), []
    for tgt in (1.0, 0.5, 0.5):
        t = f32(tgt) * a
        for _ in range(F):
            y = f32(t + f32(y * b))
            exp.append(y)
    assert np.array_equal(curves.reshape(-1), np.array(exp, f32))


# ---- (de)interleave (util.rs:44-147) -------------------------------------------------------------
def test_deinterleave_stale_mask_quirk(oracle):  # Q4
    frames, n = 8, 2
    inter = synth((frames * n,), 1)
    planar = np.zeros((3, frames), f32)
    planar[1, 3] = 0.5  # stale non-zero in channel 1's destination
    mask = oracle.deinterleave(planar.ctypes.data, 3, frames, inter.ctypes.data, n, 1)
    assert np.array_equal(planar[0], inter[0::2]) and np.array_equal(planar[1], inter[1::2])
    assert np.all(planar[2] == 0)
    assert mask == 0b101  # ch0 judged silent from STALE zeros; ch1 not; ch2 (extra) cleared + silent


def test_interleave_masks(oracle):
    frames = 16
    planar = synth((2, frames), 2)
    out = np.full(frames * 2, np.nan, f32)
    oracle.interleave(planar.ctypes.data, 2, frames, out.ctypes.data, 2, 1, 0b10, 1)  # stereo fast path ignores per-channel bits
    assert np.array_equal(out[0::2], planar[0]) and np.array_equal(out[1::2], planar[1])
    oracle.interleave(planar.ctypes.data, 2, frames, out.ctypes.data, 2, 1, 0b11, 1)
    assert np.all(out == 0)
    oracle.interleave(planar.ctypes.data, 2, frames, out.ctypes.data, 2, 1, 0b10, 0)  # generic path skips silent ch
    assert np.array_equal(out[0::2], planar[0]) and np.all(out[1::2] == 0)


# ---- VolumeNode (volume.rs:85-144) ---------------------------------------------------------------
def test_volume_constant_gain_bit_exact(oracle):
    cx, proc, _ = chain(oracle, 2, [(VolumeNode(50.0), 2, 2)])
    x = synth((1, 2, 3 * F + 17), 3)
    y, mask = run_planar(proc, x, 2)
    assert np.array_equal(y, x * f32(0.25)) and mask == 0


def test_volume_interleaved_equals_planar(oracle):
    cx, proc, _ = chain(oracle, 2, [(VolumeNode(80.0), 2, 2)])
    T = 2 * F + 5
    x = synth((1, 2, T), 4)
    inter = np.ascontiguousarray(x[0].T).reshape(-1)
    out = np.zeros(T * 2, f32)
    assert proc.process_interleaved(inter, out, 2, 2, T) == 0
    g = f32(oracle.percent_volume_to_raw_gain(80.0))
    assert np.array_equal(out.reshape(T, 2).T, x[0] * g)


def test_volume_mute_clears_and_flags(oracle):
    cx, proc, _ = chain(oracle, 2, [(VolumeNode(0.0), 2, 2)])
    x = synth((1, 2, F), 5)
    y, mask = run_planar(proc, x, 2)
    assert np.all(y == 0) and not np.any(np.signbit(y)) and mask == 0b11


def test_volume_ramp_and_stuck_smoothing_sign_of_zero(oracle):  # Q2 consequence
    cx, proc, (vol,) = chain(oracle, 2, [(VolumeNode(100.0), 2, 2)])
    nblk = 30
    x = -np.abs(synth((1, 2, nblk * F), 6)) - f32(0.01)  # strictly negative input
    cx.graph.set_percent_volume(vol, 0.0)
    y, mask = run_planar(proc, x, 2)
    curves, _, stat, _, _ = smoother_run(oracle, 1.0, [0.0] * nblk, F)
    assert np.array_equal(y[0, 0], x[0, 0] * curves.reshape(-1))
    # after settling the gain is exactly 0 but is_smoothing() stays true: x*0 = -0.0, NOT the +0.0 of the mute path
    tail = y[0, :, -F:]
    assert np.all(tail == 0) and np.all(np.signbit(tail)) and mask == 0


def test_volume_mono_and_multichannel_paths(oracle):
    cx, proc, _ = chain(oracle, 3, [(VolumeNode(150.0), 3, 3)])
    x = synth((1, 3, F), 7)
    y, _ = run_planar(proc, x, 3)
    assert np.array_equal(y, x * f32(oracle.percent_volume_to_raw_gain(150.0)))


def test_volume_activation_failure_is_reported(oracle):  # volume.rs:63-65 -> graph.rs:603-609
    cx = FirewheelGraphCtx(oracle, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2))
    cx.graph.add_node(2, 1, VolumeNode(100.0))
    proc = cx.activate(SR, 2, 2, F)
    st = cx.update()
    assert st.kind == "Active" and st.graph_error.kind == "NodeActivationFailed"
    assert "must equal the number of outputs" in str(st.graph_error)


# ---- SumNode (sum.rs:42-135) ---------------------------------------------------------------------
@pytest.mark.parametrize("ports", [1, 2, 3, 4, 5, 8])
def test_sum_left_to_right(oracle, ports):
    n_out = 2
    cx = FirewheelGraphCtx(oracle, AudioGraphConfig(num_graph_inputs=ports * n_out, num_graph_outputs=n_out))
    g = cx.graph
    s = g.add_node(ports * n_out, n_out, SumNode())
    for i in range(ports * n_out):
        g.connect(g.graph_in_node(), i, s, i, False)
    for c in range(n_out):
        g.connect(s, c, g.graph_out_node(), c, False)
    proc = cx.activate(SR, ports * n_out, n_out, F)
    assert cx.update().graph_error is None
    x = synth((1, ports * n_out, F), 10 + ports) * f32(1e3)
    y, _ = run_planar(proc, x, n_out)
    for c in range(n_out):
        acc = x[0, c].copy()
        for p in range(1, ports):
            acc = (acc + x[0, p * n_out + c]).astype(f32)  # port-major layout, strict left fold
        assert np.array_equal(y[0, c], acc)


def test_sum_generic_path_skips_silent_ports(oracle):  # sum.rs:111-133: unconnected (cleared+flagged) ports are skipped
    cx = FirewheelGraphCtx(oracle, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=1))
    g = cx.graph
    s = g.add_node(6, 1, SumNode())
    g.connect(g.graph_in_node(), 0, s, 0, False)
    g.connect(g.graph_in_node(), 1, s, 4, False)
    g.connect(s, 0, g.graph_out_node(), 0, False)
    proc = cx.activate(SR, 2, 1, F)
    assert cx.update().graph_error is None
    x = synth((1, 2, F), 21)
    x[0, 0, :4] = f32(-0.0)
    x[0, 1, :4] = f32(-0.0)
    y, _ = run_planar(proc, x, 1)
    assert np.array_equal(y[0, 0], (x[0, 0] + x[0, 1]).astype(f32))
    assert np.all(np.signbit(y[0, 0, :4]))  # -0 + -0 = -0: the four silent ports were skipped, not added as +0


# ---- small nodes -----------------------------------------------------------------------------------
def test_mono_stereo_roundtrip(oracle):
    cx, proc, _ = chain(oracle, 1, [(MonoToStereoNode(), 1, 2), (StereoToMonoNode(), 2, 1)])
    x = synth((1, 1, F), 30)
    y, _ = run_planar(proc, x, 1)
    assert np.array_equal(y[0, 0], ((x[0, 0] + x[0, 0]) * f32(0.5)).astype(f32))


def test_hard_clip(oracle):
    cx, proc, _ = chain(oracle, 2, [(HardClipNode(-6.0), 2, 2)])
    t = f32(oracle.db_to_gain_clamped_neg_100_db(-6.0))
    x = synth((1, 2, F), 31)
    y, _ = run_planar(proc, x, 2)
    assert np.array_equal(y, np.maximum(np.minimum(x, t), -t))


# ---- our-spec nodes (parity unpinned; scipy cross-checks) -----------------------------------------
def test_pan_equal_power(oracle):
    gl, gr = C.c_float(), C.c_float()
    oracle.pan_to_gains(0.0, C.byref(gl), C.byref(gr))
    assert abs(gl.value - np.sqrt(0.5)) < 1e-7 and abs(gr.value - np.sqrt(0.5)) < 1e-7
    oracle.pan_to_gains(-1.0, C.byref(gl), C.byref(gr))
    assert gl.value == 1.0 and abs(gr.value) < 1e-7
    cx, proc, _ = chain(oracle, 2, [(PanNode(0.3), 2, 2)])
    oracle.pan_to_gains(0.3, C.byref(gl), C.byref(gr))
    x = synth((1, 2, F), 32)
    y, _ = run_planar(proc, x, 2)
    assert np.array_equal(y[0, 0], x[0, 0] * f32(gl.value)) and np.array_equal(y[0, 1], x[0, 1] * f32(gr.value))


def test_biquad_cascade_vs_scipy(oracle):
    ns = 4
    cx, proc, (bq,) = chain(oracle, 2, [(BiquadNode(ns), 2, 2)], max_block=128)
    rng = np.random.default_rng(5)
    sos, co = [], []
    for s in range(ns):
        k = design_rbj(oracle, s % 2 * 4, 200.0 * (2.0 ** rng.uniform(0, 5)), rng.uniform(0.5, 2.0), 3.0, SR)
        co.append(k)
        sos.append([k[0], k[1], k[2], 1.0, k[3], k[4]])
    cx.graph.set_biquad_coeffs(bq, np.array(co))
    x = synth((1, 2, 1000), 33)
    y, _ = run_planar(proc, x, 2)  # 1000 frames = 7 full blocks + a ragged one: state carries across
    ref = scipy.signal.sosfilt(np.array(sos, dtype=np.float64), x[0].astype(np.float64), axis=-1)
    # f32 TDF-II state noise vs an f64 evaluation of the same coefficients: a sanity check of the SPEC only
    # (measured 1.8e-5 here). The parity contract is GPU == this f32 oracle bit for bit (same op order).
    assert np.max(np.abs(y[0] - ref)) / np.max(np.abs(ref)) < 1e-4


def test_delay_line(oracle):
    D = 300
    cx, proc, _ = chain(oracle, 2, [(DelayNode(D), 2, 2)], max_block=128)
    x = synth((1, 2, 1000), 34)
    y, _ = run_planar(proc, x, 2)
    assert np.all(y[0, :, :D] == 0) and np.array_equal(y[0, :, D:], x[0, :, :-D])
    x2 = synth((1, 2, 100), 35)
    y2, _ = run_planar(proc, x2, 2)  # ring persists across calls
    assert np.array_equal(y2[0], x[0, :, -D:-D + 100])


def test_conv_reverb_vs_numpy(oracle):
    L = 200
    rng = np.random.default_rng(6)
    ir = (rng.standard_normal((2, L)) * np.exp(-6.9 * np.arange(L) / L)).astype(f32)
    cx, proc, _ = chain(oracle, 2, [(ConvReverbNode(ir), 2, 2)], max_block=64)
    x = synth((1, 2, 500), 36)
    y, _ = run_planar(proc, x, 2)
    rb = np.vectorize(oracle.bf16_round, otypes=[f32])
    for c in range(2):
        ref = np.convolve(rb(x[0, c]).astype(np.float64), rb(ir[c]).astype(np.float64))[:500]
        assert np.max(np.abs(y[0, c] - ref)) <= 1e-6 * np.max(np.abs(ref))


# ---- batching + master bus -----------------------------------------------------------------------
@pytest.mark.parametrize("V", [1, 2, 5, 8, 13])
def test_master_bus_is_balanced_tree(oracle, V):
    pct = np.linspace(30, 170, V).astype(f32)
    cx, proc, (vol,) = chain(oracle, 2, [(VolumeNode(100.0), 2, 2)], voices=V, master_bus=True,
                             setup=lambda cx, ids: cx.graph.set_percent_volume(ids[0], pct))
    x = synth((V, 2, F + 9), 40 + V) * f32(100.0)
    y, mask = run_planar(proc, x, 2, master_bus=True)
    level = [x[v] * f32(oracle.percent_volume_to_raw_gain(pct[v])) for v in range(V)]
    while len(level) > 1:
        nxt = [(level[i] + level[i + 1]).astype(f32) for i in range(0, len(level) - 1, 2)]
        if len(level) & 1:
            nxt.append(level[-1])
        level = nxt
    assert np.array_equal(y, level[0]) and mask == 0


def test_master_bus_all_muted_is_flagged_silent(oracle):
    cx, proc, _ = chain(oracle, 2, [(VolumeNode(0.0), 2, 2)], voices=4, master_bus=True)
    y, mask = run_planar(proc, synth((4, 2, F), 50), 2, master_bus=True)
    assert np.all(y == 0) and mask == 0b11


def test_per_voice_outputs_without_bus(oracle):
    V = 3
    def setup(cx, ids):
        cx.graph.set_percent_volume(ids[0], np.array([50, 100, 200], f32))
        cx.graph.set_pan(ids[1], np.array([-1.0, 0.0, 0.5], f32))
    cx, proc, (vol, pan) = chain(oracle, 2, [(VolumeNode(100.0), 2, 2), (PanNode(0.0), 2, 2)], voices=V, setup=setup)
    x = synth((V, 2, 2 * F), 51)
    y, _ = run_planar(proc, x, 2)
    gl, gr = C.c_float(), C.c_float()
    for v, (pc, pn) in enumerate([(50, -1.0), (100, 0.0), (200, 0.5)]):
        oracle.pan_to_gains(pn, C.byref(gl), C.byref(gr))
        g = f32(oracle.percent_volume_to_raw_gain(pc))
        assert np.array_equal(y[v, 0], (x[v, 0] * g).astype(f32) * f32(gl.value))
        assert np.array_equal(y[v, 1], (x[v, 1] * g).astype(f32) * f32(gr.value))


# ---- lifecycle ---------------------------------------------------------------------------------------
def test_lifecycle_no_schedule_outputs_silence_then_activates(oracle):  # processor.rs:76-89
    cx = FirewheelGraphCtx(oracle, AudioGraphConfig(num_graph_inputs=1, num_graph_outputs=1))
    g = cx.graph
    g.connect(g.graph_in_node(), 0, g.graph_out_node(), 0, False)
    assert cx.update().kind == "Inactive"
    proc = cx.activate(SR, 1, 1, F)
    assert cx.activate(SR, 1, 1, F) is None  # already active
    x = synth((F,), 60)
    out = np.full(F, np.nan, f32)
    assert proc.process_interleaved(x, out, 1, 1, F) == 0 and np.all(out == 0)  # no schedule yet
    assert cx.update().kind == "Active"
    assert proc.process_interleaved(x, out, 1, 1, F) == 0 and np.array_equal(out, x)  # graph_in -> graph_out passthrough
    proc.free()
    assert cx.update().kind == "Deactivated"
    assert not cx.is_activated()


def test_recompile_swaps_schedule_and_keeps_node_state(oracle):  # processor.rs:167-206
    cx, proc, (vol,) = chain(oracle, 2, [(VolumeNode(100.0), 2, 2)])
    g = cx.graph
    x = synth((1, 2, F), 61)
    y, _ = run_planar(proc, x, 2)
    assert np.array_equal(y, x)
    clip = g.add_node(2, 2, HardClipNode(-12.0))
    assert g.disconnect(vol, 0, g.graph_out_node(), 0) and g.disconnect(vol, 1, g.graph_out_node(), 1)
    for c in range(2):
        g.connect(vol, c, clip, c, False)
        g.connect(clip, c, g.graph_out_node(), c, False)
    assert g.needs_compile()
    assert cx.update().graph_error is None
    t = f32(oracle.db_to_gain_clamped_neg_100_db(-12.0))
    # Q11: graph inputs are written into the OLD schedule's pool (processor.rs:99-115) before process_block
    # polls and adopts the new schedule (processor.rs:214), whose fresh pool is all zeros: the first block
    # after a swap processes silence-valued (but not silence-flagged) inputs.
    x2 = np.concatenate([x, x], axis=2)
    y, _ = run_planar(proc, x2, 2)
    assert np.all(y[:, :, :F] == 0)
    assert np.array_equal(y[:, :, F:], np.maximum(np.minimum(x, t), -t))
    g.remove_node(clip)
    for c in range(2):
        g.connect(vol, c, g.graph_out_node(), c, False)
    assert cx.update().graph_error is None
    y, _ = run_planar(proc, x2, 2)
    assert np.all(y[:, :, :F] == 0) and np.array_equal(y[:, :, F:], x)
