"""CPU pins for the two nodes whose spec is ours (SURVEY §8 a11 SVF, a13 polyphase resampler; parity unpinned by the
reference): the oracle against scipy / closed forms, so that "bit-exact against the oracle" on the GPU means something."""
import numpy as np
import scipy.signal

from conftest import synth
from firewheel_b200 import (AudioGraphConfig, FirewheelGraphCtx, ResamplerNode, SvfNode, design_resampler, design_svf)

f32 = np.float32
SR = 48000


def run(proc, x, n_out):
    V, n_in, T = x.shape
    out = np.full((V, n_out, T), np.nan, f32)
    rc, mask = proc.process_planar(np.ascontiguousarray(x), out, n_in, n_out, T)
    assert rc == 0
    return out, mask


def svf_ctx(lib, coeffs, F=256):
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=1, num_graph_outputs=1))
    g = cx.graph
    n = g.add_node(1, 1, SvfNode(len(coeffs)))
    g.connect(g.graph_in_node(), 0, n, 0, False); g.connect(n, 0, g.graph_out_node(), 0, False)
    g.set_svf_coeffs(n, np.stack(coeffs))
    proc = cx.activate(SR, 1, 1, F)
    assert cx.update().graph_error is None
    return cx, proc


def test_svf_matches_the_bilinear_biquad(oracle):
    """The trapezoidal SVF is the bilinear transform of the analog 2-pole prototype: its lowpass / highpass / bandpass
    outputs equal scipy's bilinear-transformed H(s) = {1, s^2, s/Q} / (s^2 + s/Q + 1) (pre-warped), to f32 accuracy."""
    fc, q, T = 1500.0, 0.9, 2048
    x = np.zeros((1, 1, T), f32); x[0, 0, 0] = 1.0
    wc = 2 * SR * np.tan(np.pi * fc / SR)
    protos = {0: ([0, 0, wc * wc], [1, wc / q, wc * wc]), 2: ([1, 0, 0], [1, wc / q, wc * wc]), 1: ([0, wc, 0], [1, wc / q, wc * wc])}
    for ftype, (b, a) in protos.items():
        cx, proc = svf_ctx(oracle, [design_svf(oracle, ftype, fc, q, SR)])
        y, _ = run(proc, x, 1)
        bz, az = scipy.signal.bilinear(b, a, fs=SR)
        ref = scipy.signal.lfilter(bz, az, x[0, 0].astype(np.float64))
        assert np.max(np.abs(y[0, 0] - ref)) <= 2e-6 * max(1.0, np.max(np.abs(ref))), ftype
        proc.free(); cx.update(); cx.free()


def test_svf_state_carries_across_blocks_and_calls(oracle):
    k = [design_svf(oracle, 0, 800.0, 2.0, SR), design_svf(oracle, 4, 3000.0, 1.2, SR)]
    x = synth((1, 1, 1000), 5)
    cx, proc = svf_ctx(oracle, k, F=64)
    whole, _ = run(proc, x, 1)
    proc.free(); cx.update(); cx.free()
    cx, proc = svf_ctx(oracle, k, F=100)
    parts = np.concatenate([run(proc, np.ascontiguousarray(x[:, :, :333]), 1)[0], run(proc, np.ascontiguousarray(x[:, :, 333:]), 1)[0]], axis=2)
    assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32))
    proc.free(); cx.update(); cx.free()


def rs_ctx(lib, table, n_out, F=128):
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=n_out))
    g = cx.graph
    n = g.add_node(0, n_out, ResamplerNode(table))
    for c in range(n_out):
        g.connect(n, c, g.graph_out_node(), c, False)
    proc = cx.activate(SR, 0, n_out, F)
    assert cx.update().graph_error is None
    return cx, g, n, proc


def test_resampler_table_and_unit_ratio(oracle):
    """Phase 0 of the table is a unit impulse at tap T/2-1 (sinc at integers), every phase has DC gain ~1, and a ratio of
    exactly 1.0 therefore reproduces the resource sample for sample."""
    tab = design_resampler(oracle, 256, 32, 1.0, 9.0)
    assert tab[0, 15] == 1.0 and np.all(np.abs(np.delete(tab[0], 15)) < 1e-7)
    assert np.max(np.abs(tab.sum(axis=1) - 1.0)) < 2e-4
    x = synth((2, 700), 8)
    cx, g, n, proc = rs_ctx(oracle, tab, 2)
    g.resampler_set(n, g.create_sample_resource(x), ratio=1.0)
    y, m = run(proc, np.zeros((1, 0, 512), f32), 2)
    assert np.array_equal(y[0], x[:, :512]) and m == 0
    y2, _ = run(proc, np.zeros((1, 0, 300), f32), 2)   # position carried across calls; zeros past the end (one-shot)
    assert np.array_equal(y2[0, :, :188], x[:, 512:]) and np.all(y2[0, :, 188 + 16:] == 0)
    proc.free(); cx.update(); cx.free()


def test_resampler_interpolates_a_sine(oracle):
    """Upsampling by 8/3 and downsampling by 0.6 (with the cutoff lowered accordingly) follow the analytic sine."""
    n = np.arange(4000)
    for ratio, cutoff, f0 in [(3.0 / 8.0, 1.0, 1234.0), (1.0 / 0.6, 0.55, 900.0)]:
        tab = design_resampler(oracle, 256, 32, cutoff, 9.0)
        src = np.sin(2 * np.pi * f0 / SR * n).astype(f32)[None, :]
        cx, g, node, proc = rs_ctx(oracle, tab, 1)
        g.resampler_set(node, g.create_sample_resource(src), ratio=ratio)
        g.resampler_seek(node, 100)
        T = 1024
        y, _ = run(proc, np.zeros((1, 0, T), f32), 1)
        step = int(round(ratio * 2 ** 32))
        pos = 100 + np.arange(T) * step / 2 ** 32
        ref = np.sin(2 * np.pi * f0 / SR * pos)
        assert np.max(np.abs(y[0, 0] - ref)) < 4e-3, (ratio, np.max(np.abs(y[0, 0] - ref)))   # 256 phases, nearest-lower phase
        proc.free(); cx.update(); cx.free()


def test_resampler_loop_and_channel_mapping(oracle):
    tab = design_resampler(oracle, 64, 16, 1.0, 8.0)
    mono = synth((1, 50), 2)
    cx, g, node, proc = rs_ctx(oracle, tab, 3)
    g.resampler_set(node, g.create_sample_resource(mono), ratio=1.0, loop=True)
    y, m = run(proc, np.zeros((1, 0, 128), f32), 3)
    assert np.array_equal(y[0, 0], np.tile(mono[0], 3)[:128]) and np.all(y[0, 1:] == 0) and m == 0b110   # wraps modulo the length
    g.resampler_set(node, 0, ratio=1.0, playing=False)
    y, m = run(proc, np.zeros((1, 0, 128), f32), 3)
    assert np.all(y == 0) and m == 0b111
    proc.free(); cx.update(); cx.free()


def test_host_design_helpers_agree_between_product_and_oracle(oracle, product):
    """The design helpers are host-only f64 code on both sides of the parity boundary; tests hand the same table / coefficients
    to both anyway, but the helpers themselves must agree bit for bit (no GPU needed)."""
    from firewheel_b200 import design_rbj
    for args in [(256, 32, 1.0, 9.0), (64, 16, 0.45, 6.5), (1024, 64, 0.9, 12.0)]:
        assert np.array_equal(design_resampler(oracle, *args).view(np.uint32), design_resampler(product, *args).view(np.uint32)), args
    for ftype in range(6):
        for fc, q in [(80.0, 0.5), (1234.5, 0.707), (15000.0, 8.0)]:
            assert np.array_equal(design_svf(oracle, ftype, fc, q, SR).view(np.uint32), design_svf(product, ftype, fc, q, SR).view(np.uint32))
    for ftype in range(7):
        a, b = design_rbj(oracle, ftype, 997.0, 1.3, 4.5, SR), design_rbj(product, ftype, 997.0, 1.3, 4.5, SR)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
