"""Known-answer tests that pin the oracle's SamplerNode / SampleResource restatement (CPU only).

The reference has no sampler tests, so every expected value below is derived here, in numpy, from the reference source
(file:line cited) — never from the oracle's own output."""
import numpy as np
import pytest

from conftest import synth
from firewheel_b200 import AudioGraphConfig, FirewheelGraphCtx, SamplerError, SamplerNode

f32 = np.float32
SR = 48000


def sampler_ctx(lib, n_out, F, V=1, pct=100.0, activate=True):
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=n_out, num_voices=V))
    g = cx.graph
    smp = g.add_node(0, n_out, SamplerNode(pct))
    for c in range(n_out):
        g.connect(smp, c, g.graph_out_node(), c, False)
    proc = None
    if activate:
        proc = cx.activate(SR, 0, n_out, F)
        st = cx.update()
        assert st.kind == "Active" and st.graph_error is None
    return cx, g, smp, proc


def run(proc, V, n_out, T):
    out = np.full((V, n_out, T), np.nan, f32)
    rc, mask = proc.process_planar(np.zeros((V, 0, T), f32), out, 0, n_out, T)
    assert rc == 0
    return out, mask


def test_pcm_conversions(oracle):
    """pcm_i16_to_f32 = s as f32 * (1.0 / i16::MAX as f32); pcm_u16_to_f32 = s as f32 * (2.0 / u16::MAX as f32) - 1.0
    (sample_resource.rs:337-345), each a single f32 op sequence."""
    F = 64
    i16 = np.array([-32768, -32767, -1, 0, 1, 12345, 32767] + list(range(-28, 29)), np.int16)
    u16 = np.array([0, 1, 32767, 32768, 65534, 65535] + list(range(100, 158)), np.uint16)
    for data, want in [(i16, (i16.astype(f32) * (f32(1.0) / f32(32767.0))).astype(f32)),
                       (u16, ((u16.astype(f32) * (f32(2.0) / f32(65535.0))).astype(f32) - f32(1.0)).astype(f32))]:
        for interleaved in (True, False):
            cx, g, smp, proc = sampler_ctx(oracle, 1, F)
            h = g.create_sample_resource(data[:, None] if interleaved else data[None, :], interleaved=interleaved)
            g.sampler_set_sample(smp, h, True); g.sampler_play(smp)
            y, m = run(proc, 1, 1, F)
            assert np.array_equal(y[0, 0].view(np.uint32), want.view(np.uint32)) and m == 0
            proc.free(); cx.update(); cx.free()
    assert f32(-32768) * (f32(1.0) / f32(32767.0)) < -1.0  # the reference's scale maps i16::MIN just below -1.0


def test_one_shot_end_and_zero_tail(oracle):
    """sampler.rs:485-516: the last partial block is zero-filled, playing = false, playhead = 0; later blocks are cleared
    and flagged; gain = (percent/100)^2 (range.rs:32-35) multiplies every copied frame (sampler.rs:522-543)."""
    F, L = 32, 80
    x = synth((2, L), 3)
    cx, g, smp, proc = sampler_ctx(oracle, 2, F, pct=50.0)
    g.sampler_set_sample(smp, g.create_sample_resource(x), True); g.sampler_play(smp)
    y, m = run(proc, 1, 2, 4 * F)
    gain = f32(0.5) * f32(0.5)
    want = np.zeros((2, 4 * F), f32); want[:, :L] = x * gain
    assert np.array_equal(y[0].view(np.uint32), want.view(np.uint32))
    assert m == 0b11                      # last block: not playing any more -> clear_all_outputs
    y2, m2 = run(proc, 1, 2, F)
    assert np.all(y2 == 0) and m2 == 0b11
    g.sampler_play(smp)                   # node side still believes it is playing (sampler.rs:82-98): no message, stays silent
    y3, _ = run(proc, 1, 2, F)
    assert np.all(y3 == 0)
    g.sampler_pause(smp); g.sampler_play(smp)   # pause resets the node-side flag; play restarts from playhead 0 (sampler.rs:505)
    y4, _ = run(proc, 1, 2, F)
    assert np.array_equal(y4[0], x[:, :F] * gain)
    proc.free(); cx.update(); cx.free()


def test_loop_wrap_inside_and_across_blocks(oracle):
    """sampler.rs:445-484: at most one wrap per block; the second copy starts at the loop start and is not re-checked
    against the loop end."""
    F, L = 48, 200
    x = synth((1, L), 9)
    cx, g, smp, proc = sampler_ctx(oracle, 1, F)
    g.sampler_set_sample(smp, g.create_sample_resource(x), True)
    lo, hi = 30, 130                      # loop 30..130 (100 frames)
    g.sampler_set_loop_range(smp, (lo / SR, hi / SR))
    g.sampler_play(smp)
    y, _ = run(proc, 1, 1, 6 * F)
    want, ph = [], 0                      # SetLoopRange: playhead 0 is outside the range, stays 0 (sampler.rs:405-411)
    for _ in range(6):
        if ph >= hi:
            ph = lo
        first = min(F, hi - ph)
        blk = list(x[0, ph:ph + first])
        if first < F:
            blk += list(x[0, lo:lo + F - first]); ph = lo + (F - first)
        else:
            ph += F
        want += blk
    assert np.array_equal(y[0, 0], np.array(want, f32))
    proc.free(); cx.update(); cx.free()


def test_channel_mapping_and_masks(oracle):
    """sampler.rs:545-559: mono sample into a stereo node is duplicated; any other surplus channel is zeroed and flagged;
    a sample with more channels than the node drops the extra ones (sample_resource.rs:20-25)."""
    F = 16
    mono, tri = synth((1, F), 1), synth((3, F), 2)
    for data, n_out, want, mask in [(mono, 2, np.stack([mono[0], mono[0]]), 0),
                                    (mono, 3, np.stack([mono[0], np.zeros(F, f32), np.zeros(F, f32)]), 0b110),
                                    (tri, 2, tri[:2], 0),
                                    (tri, 5, np.concatenate([tri, np.zeros((2, F), f32)]), 0b11000)]:
        cx, g, smp, proc = sampler_ctx(oracle, n_out, F)
        g.sampler_set_sample(smp, g.create_sample_resource(data), True); g.sampler_play(smp)
        y, m = run(proc, 1, n_out, F)
        assert np.array_equal(y[0], want) and m == mask, (n_out, m)
        proc.free(); cx.update(); cx.free()


def test_message_semantics(oracle):
    F, L = 10, 1000
    x = synth((1, L), 4)
    cx, g, smp, proc = sampler_ctx(oracle, 1, F)
    h = g.create_sample_resource(x)
    g.sampler_set_sample(smp, h, True); g.sampler_play(smp)
    g.sampler_set_playhead(smp, 105.4 / SR)            # round() -> frame 105 (sampler.rs:394)
    y, _ = run(proc, 1, 1, F)
    assert np.array_equal(y[0, 0], x[0, 105:115])
    g.sampler_pause(smp)
    assert np.all(run(proc, 1, 1, F)[0] == 0)           # paused: cleared, playhead kept
    g.sampler_play(smp)
    assert np.array_equal(run(proc, 1, 1, F)[0][0, 0], x[0, 115:125])
    g.sampler_stop(smp); g.sampler_play(smp)            # stop: playhead -> loop start or 0 (sampler.rs:379-391)
    assert np.array_equal(run(proc, 1, 1, F)[0][0, 0], x[0, 0:10])
    g.sampler_set_loop_range(smp, (0.0, 50 / SR))       # playhead 10 is inside the new range -> jumps to its start (:405-411)
    assert np.array_equal(run(proc, 1, 1, F)[0][0, 0], x[0, 0:10])
    g.sampler_set_loop_range(smp, "full"); g.sampler_set_playhead(smp, (L - 5) / SR)
    y, _ = run(proc, 1, 1, F)
    assert np.array_equal(y[0, 0], np.concatenate([x[0, L - 5:], x[0, :5]]))   # full range = 0..len_frames (:241-248)
    with pytest.raises(SamplerError):
        g.sampler_set_loop_range(smp, (0.5, 0.25))      # start >= end: the reference's u64 subtraction underflows
    proc.free(); cx.update(); cx.free()


def test_ring_capacity_and_activation(oracle):
    """The node -> processor ring holds 128 messages (sampler.rs:14): the 129th push fails with Err(()); before activation
    the reference reaches todo!() — reported as NotActivated."""
    cx, g, smp, _ = sampler_ctx(oracle, 1, 16, activate=False)
    with pytest.raises(SamplerError) as e:
        g.sampler_play(smp)
    assert e.value.kind == "NotActivated"
    cx.free()
    cx, g, smp, proc = sampler_ctx(oracle, 1, 16)
    for i in range(128):
        g.sampler_set_playhead(smp, i / SR)
    with pytest.raises(SamplerError) as e:
        g.sampler_set_playhead(smp, 1.0)
    assert e.value.kind == "RingFull"
    run(proc, 1, 1, 16)                                 # drained by the first block of the next call (sampler.rs:331)
    g.sampler_set_playhead(smp, 0.0)
    info = g.node_info(smp)
    assert info.debug_name == b"beep_test" and info.updates  # Q8 (sampler.rs:186), sampler.rs:193
    proc.free(); cx.update(); cx.free()
