"""The pull-style stream backend (the role of firewheel-cpal's DataCallback, lib.rs:378-449): a producer
thread renders periods ahead of the consumer; what the consumer pulls must be exactly the concatenation of
process_interleaved calls of one period each — checked against the CPU oracle driven that way."""
import time

import numpy as np
import pytest

from conftest import synth
from firewheel_b200 import AudioGraphConfig, FirewheelGraphCtx, PanNode, SamplerNode, VolumeNode
from helpers import SR, f32

pytestmark = pytest.mark.gpu
UNDERFLOW = 2  # FW_STREAM_OUTPUT_UNDERFLOW


def build(lib, V, F):
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=2, num_voices=V, master_bus=True))
    g = cx.graph
    smp = g.add_node(0, 2, SamplerNode(80.0)); vol = g.add_node(2, 2, VolumeNode(70.0)); pan = g.add_node(2, 2, PanNode(0.1))
    for c in range(2):
        g.connect(smp, c, vol, c, False); g.connect(vol, c, pan, c, False); g.connect(pan, c, g.graph_out_node(), c, False)
    proc = cx.activate(SR, 0, 2, F)
    assert cx.update().graph_error is None, cx.last_error()
    res = [g.create_sample_resource(synth((2, 5000 + 37 * i), 60 + i)) for i in range(3)]
    for v in range(V):
        g.sampler_set_sample(smp, res[v % 3], True, voice=v)
        g.sampler_set_loop_range(smp, "full", voice=v)
    g.sampler_play(smp)
    return cx, proc


def wait_ready(st, frames, timeout=10.0):
    t0 = time.time()
    while st.frames_ready() < frames:
        assert time.time() - t0 < timeout, "producer thread made no progress"
        time.sleep(0.001)


def test_pulled_audio_equals_period_by_period_rendering(gpu, oracle):
    V, F, period, n_periods = 9, 128, 384, 10
    cx, proc = build(gpu, V, F)
    st = proc.open_stream(2, SR, period, ring_periods=4)
    assert st is not None, gpu.last_device_error()
    got, pulled, sizes = [], 0, [100, 1, 383, 384, 700, 5, 64]
    total = period * n_periods
    i = 0
    while pulled < total:
        n = min(sizes[i % len(sizes)], total - pulled); i += 1
        wait_ready(st, n)
        y, k, status, t = st.pull(n)
        assert k == n and status == 0 and t == pulled / SR
        got.append(y); pulled += n
    st.close()
    proc.free(); cx.update(); cx.free()
    got = np.concatenate(got)

    cx, proc = build(oracle, V, F)
    want = []
    for _ in range(n_periods):
        out = np.full((period, 2), np.nan, f32)
        assert proc.process_interleaved(np.zeros(0, f32), out, 0, 2, period) == 0
        want.append(out)
    proc.free(); cx.update(); cx.free()
    want = np.concatenate(want)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_underflow_is_reported_and_zero_filled(gpu):
    cx, proc = build(gpu, 4, 64)
    st = proc.open_stream(2, SR, 256, ring_periods=2)
    wait_ready(st, 512)
    y, k, status, _ = st.pull(4096)        # far more than the ring can hold: the consumer outruns the producer
    assert 512 <= k < 4096 and status & UNDERFLOW and np.all(y[k:] == 0) and np.any(y[:k] != 0)
    wait_ready(st, 256)
    y2, k2, status2, t2 = st.pull(256)     # the stream recovers; its clock counts delivered frames only
    assert k2 == 256 and status2 == 0 and t2 == k / SR
    st.close()
    proc.free(); cx.update(); cx.free()


def test_streams_need_a_single_output(gpu):
    cx = FirewheelGraphCtx(gpu, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=2, num_voices=3, master_bus=False))
    g = cx.graph
    smp = g.add_node(0, 2, SamplerNode(80.0))
    for c in range(2):
        g.connect(smp, c, g.graph_out_node(), c, False)
    proc = cx.activate(SR, 0, 2, 64)
    assert cx.update().graph_error is None
    assert proc.open_stream(2, SR, 64) is None
    proc.free(); cx.update(); cx.free()
