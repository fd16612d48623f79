"""The C++ oracle against tests/pyref.py, an independent pure-Python restatement of the same reference source: executor
silence flags, ParamSmoother, Volume / Sum / MonoToStereo / StereoToMono / HardClip on random small DAGs with parameter
changes between calls. Two restatements written separately agreeing bit for bit is the strongest pin available while the
reference itself cannot be compiled here. CPU only."""
import numpy as np
import pytest

import pyref
from conftest import synth
from firewheel_b200 import AudioGraphConfig, FirewheelGraphCtx, HardClipNode, MonoToStereoNode, StereoToMonoNode, SumNode, VolumeNode

f32 = np.float32
SR = 48000
KINDS = ["vol1", "vol2", "vol3", "clip1", "clip2", "clip3", "m2s", "s2m", "sum2x1", "sum2x2", "sum3x1", "sum4x2", "sum6x1", "sum5x2", "sum1x2"]


def build(lib, seed, F):
    """Random DAG on `lib`; returns (ctx, processor, n_in, n_out, python-side processors keyed by node id, volume node ids)."""
    rng = np.random.default_rng(seed)
    n_in, n_out = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=n_in, num_graph_outputs=n_out))
    g = cx.graph
    py = {int(g.graph_in_node()): pyref.Dummy(), int(g.graph_out_node()): pyref.Dummy()}
    sources = [(g.graph_in_node(), p) for p in range(n_in)]
    vols = []
    for _ in range(int(rng.integers(3, 10))):
        k = KINDS[int(rng.integers(len(KINDS)))]
        if k.startswith("vol"):
            ch, pct = int(k[3]), float(rng.choice([0.0, 40.0, 100.0, 130.0]))
            nid = g.add_node(ch, ch, VolumeNode(pct)); ni = no = ch
            py[int(nid)] = pyref.Volume(pct, SR, F); vols.append(nid)
        elif k.startswith("clip"):
            ch, db = int(k[4]), float(rng.uniform(-12, 0))
            nid = g.add_node(ch, ch, HardClipNode(db)); ni = no = ch
            py[int(nid)] = pyref.HardClip(pyref.db_to_gain_clamped(db))
        elif k == "m2s":
            nid = g.add_node(1, 2, MonoToStereoNode()); ni, no = 1, 2; py[int(nid)] = pyref.MonoToStereo()
        elif k == "s2m":
            nid = g.add_node(2, 1, StereoToMonoNode()); ni, no = 2, 1; py[int(nid)] = pyref.StereoToMono()
        else:
            ports, ch = int(k[3]), int(k[5]); ni, no = ports * ch, ch
            nid = g.add_node(ni, no, SumNode()); py[int(nid)] = pyref.Sum(ni, no)
        for p in range(ni):
            if rng.random() < 0.8:
                src = sources[int(rng.integers(max(0, len(sources) - 6), len(sources)))]
                g.connect(src[0], src[1], nid, p, False)
        sources += [(nid, p) for p in range(no)]
    for p in range(n_out):
        if rng.random() < 0.9:
            src = sources[int(rng.integers(max(0, len(sources) - 4), len(sources)))]
            g.connect(src[0], src[1], g.graph_out_node(), p, False)
    sched, num_buffers = g.compile_internal(F)
    ex = pyref.Executor([(int(s.id), s.input_buffers, s.output_buffers) for s in sched], num_buffers, F, py)
    proc = cx.activate(SR, n_in, n_out, F)
    st = cx.update()
    assert proc is not None and st.graph_error is None
    return cx, proc, n_in, n_out, ex, py, vols


@pytest.mark.parametrize("seed", range(16))
def test_oracle_equals_the_python_restatement(oracle, seed):
    F = int(np.random.default_rng(seed).choice([8, 16, 24]))
    cx, proc, n_in, n_out, ex, py, vols = build(oracle, seed, F)
    rng = np.random.default_rng(100 + seed)
    for call in range(5):
        if call in (1, 3):  # retune every gain: ramps, mutes, stalls (Q2, Q3, Q10)
            for nid in vols:
                pct = float(rng.choice([0.0, 25.0, 100.0]))
                cx.graph.set_percent_volume(nid, pct); py[int(nid)].set_percent(pct)
        T = int(rng.choice([F, 3 * F, 2 * F + 5]))
        x = synth((1, n_in, T), 7 * seed + call)
        if call == 2:
            x[:, :, : T // 2] = -0.0
        out = np.full((1, n_out, T), np.nan, f32)
        rc, mask = proc.process_planar(np.ascontiguousarray(x), out, n_in, n_out, T)
        assert rc == 0
        want, want_mask = ex.process(x[0], n_out)
        assert np.array_equal(out[0].view(np.uint32), want.view(np.uint32)), (seed, call, np.argwhere(out[0].view(np.uint32) != want.view(np.uint32))[:3])
        assert mask == want_mask, (seed, call, hex(mask), hex(want_mask))
    proc.free(); cx.update(); cx.free()


@pytest.mark.parametrize("seed", range(12))
def test_sampler_oracle_equals_the_python_restatement(oracle, seed):
    """SamplerNode under random transport: resources of every sample type, set_sample / play / pause / stop / set_playhead /
    loop ranges / volume changes between calls, samples ending and loops wrapping inside blocks."""
    from firewheel_b200 import SamplerNode
    rng = np.random.default_rng(500 + seed)
    F, n_out = int(rng.choice([8, 16, 20])), int(rng.integers(1, 4))
    cx = FirewheelGraphCtx(oracle, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=n_out))
    g = cx.graph
    smp = g.add_node(0, n_out, SamplerNode(100.0))
    for c in range(n_out):
        g.connect(smp, c, g.graph_out_node(), c, False)
    sched, nb = g.compile_internal(F)
    py_s = pyref.Sampler(100.0, SR, F)
    py = {int(g.graph_in_node()): pyref.Dummy(), int(g.graph_out_node()): pyref.Dummy(), int(smp): py_s}
    ex = pyref.Executor([(int(s.id), s.input_buffers, s.output_buffers) for s in sched], nb, F, py)
    proc = cx.activate(SR, 0, n_out, F)
    assert cx.update().graph_error is None
    datas = [synth((2, 50), seed), synth((1, 37), seed + 1), rng.integers(-32768, 32768, size=(2, 64)).astype(np.int16),
             rng.integers(0, 65536, size=(1, 45)).astype(np.uint16), synth((3, 29), seed + 2)]
    res = [(g.create_sample_resource(d), pyref.Resource(d)) for d in datas]
    node_playing = False  # SamplerNode::playing (sampler.rs:51): play / pause / stop are filtered on the node side
    for call in range(10):
        for _ in range(int(rng.integers(0, 4))):
            op = rng.choice(["set_sample", "play", "pause", "stop", "playhead", "loop_none", "loop_full", "loop_range", "volume"])
            if op == "set_sample":
                h, r = res[int(rng.integers(len(res)))]; stop = bool(rng.integers(2))
                g.sampler_set_sample(smp, h, stop); py_s.msgs.append(("set_sample", r, stop))
            elif op == "play":
                g.sampler_play(smp)
                if not node_playing:
                    py_s.msgs.append(("play",)); node_playing = True
            elif op in ("pause", "stop"):
                getattr(g, "sampler_" + op)(smp)
                if node_playing:
                    py_s.msgs.append((op,)); node_playing = False
            elif op == "playhead":
                secs = float(rng.integers(0, 70)) / SR
                g.sampler_set_playhead(smp, secs); py_s.msgs.append(("playhead", secs))
            elif op == "loop_none":
                g.sampler_set_loop_range(smp, None); py_s.msgs.append(("loop", None))
            elif op == "loop_full":
                g.sampler_set_loop_range(smp, "full"); py_s.msgs.append(("loop", "full"))
            elif op == "loop_range":
                a = int(rng.integers(0, 20)); b = a + int(rng.integers(3, 25))   # may reach past short samples: zeros there
                g.sampler_set_loop_range(smp, (a / SR, b / SR)); py_s.msgs.append(("loop", (a / SR, b / SR)))
            else:
                pct = float(rng.choice([0.0, 50.0, 100.0]))
                g.sampler_set_percent_volume(smp, pct); py_s.set_percent(pct)
        T = int(rng.choice([F, 2 * F, 3 * F]))  # whole blocks (Q1: the reference asserts on short blocks while the smoother idles)
        out = np.full((1, n_out, T), np.nan, f32)
        rc, mask = proc.process_planar(np.zeros((1, 0, T), f32), out, 0, n_out, T)
        assert rc == 0
        want, want_mask = ex.process(np.zeros((0, T), f32), n_out)
        assert np.array_equal(out[0].view(np.uint32), want.view(np.uint32)), (seed, call)
        assert mask == want_mask, (seed, call, hex(mask), hex(want_mask))
    proc.free(); cx.update(); cx.free()


@pytest.mark.parametrize("seed", range(16, 28))
def test_interleaved_entry_point_equals_the_python_restatement(oracle, seed):
    """process_interleaved (processor.rs:61-165): deinterleave, block loop, and the two interleave flavours with their different
    treatment of flagged channels (util.rs:90-147)."""
    F = int(np.random.default_rng(seed).choice([8, 16]))
    cx, proc, n_in, n_out, ex, py, vols = build(oracle, seed, F)
    rng = np.random.default_rng(900 + seed)
    for call in range(4):
        if call == 2:
            for nid in vols:
                pct = float(rng.choice([0.0, 60.0]))
                cx.graph.set_percent_volume(nid, pct); py[int(nid)].set_percent(pct)
        T = int(rng.choice([F, 2 * F + 3]))
        x = np.ascontiguousarray(synth((T, n_in), 11 * seed + call))
        out = np.full((T, n_out), np.nan, f32)
        assert proc.process_interleaved(x, out, n_in, n_out, T) == 0
        want = ex.process_interleaved(x, n_out)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (seed, call)
    proc.free(); cx.update(); cx.free()
