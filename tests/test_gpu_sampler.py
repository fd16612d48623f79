"""GPU parity for the SamplerNode (sampler.rs:283-560) and the sample resources (sample_resource.rs) on the device:
the CUDA product against the oracle's restatement, bit-exact including silence masks, through the C ABI."""
import numpy as np
import pytest

from conftest import synth
from firewheel_b200 import AudioGraphConfig, FirewheelGraphCtx, PanNode, SamplerNode, SumNode, VolumeNode
from helpers import SR, assert_bit_exact, f32, run_planar

pytestmark = pytest.mark.gpu


def make_resources(g, seed=0):
    """One resource per SampleResource impl (sample_resource.rs:28-335), lengths chosen to end / wrap mid-block."""
    rng = np.random.default_rng(seed)
    i16 = lambda shape: rng.integers(-32768, 32768, size=shape, dtype=np.int64).astype(np.int16)
    u16 = lambda shape: rng.integers(0, 65536, size=shape, dtype=np.int64).astype(np.uint16)
    return {
        "f32p_stereo": g.create_sample_resource(synth((2, 1000), seed + 1)),
        "f32p_mono": g.create_sample_resource(synth((1, 777), seed + 2)),
        "f32i_stereo": g.create_sample_resource(synth((1500, 2), seed + 3), interleaved=True),
        "i16i_stereo": g.create_sample_resource(i16((900, 2)), interleaved=True),
        "u16i_mono": g.create_sample_resource(u16((640, 1)), interleaved=True),
        "i16p_3ch": g.create_sample_resource(i16((3, 512))),
        "u16p_stereo": g.create_sample_resource(u16((2, 2048))),
        "f32i_3ch": g.create_sample_resource(synth((333, 3), seed + 4), interleaved=True),
        "tiny": g.create_sample_resource(synth((2, 40), seed + 5)),  # shorter than a block: loop wraps inside one block
    }


def run_scenario(lib, V, n_out, F, steps, bus=False, post=None, n_graph_out=None, stray=False):
    n_graph_out = n_graph_out or n_out
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=n_graph_out, num_voices=V, master_bus=bus))
    g = cx.graph
    smp = g.add_node(0, n_out, SamplerNode(100.0))
    last, ids = smp, {"smp": smp}
    if post:
        last = post(g, smp, ids)
    for c in range(n_graph_out):
        g.connect(last, c, g.graph_out_node(), c, False)
    if stray:  # an unconnected node makes the graph a non-chain: the generic per-node lowering runs instead of the fused one
        g.add_node(1, 1, VolumeNode(100.0))
    proc = cx.activate(SR, 0, n_graph_out, F)
    st = cx.update()
    assert st.kind == "Active" and st.graph_error is None, (st, cx.last_error())
    res = make_resources(g)
    outs = []
    for act, T in steps:
        if act:
            act(g, ids, res)
        outs.append(run_planar(proc, np.zeros((V, 0, T), f32), n_graph_out, bus))
    proc.free(); cx.update(); cx.free()
    return outs


def both(gpu, oracle, *args, **kw):
    og, oo = run_scenario(gpu, *args, **kw), run_scenario(oracle, *args, **kw)
    for i, ((yg, mg), (yo, mo)) in enumerate(zip(og, oo)):
        assert_bit_exact(yg, yo, f"call {i}")
        assert mg == mo, f"call {i}: silence mask {mg:#x} != {mo:#x}"
    return og


@pytest.mark.parametrize("name,n_out", [("f32p_stereo", 2), ("f32p_mono", 2), ("f32p_mono", 1), ("f32i_stereo", 2), ("i16i_stereo", 2), ("u16i_mono", 2),
                                        ("i16p_3ch", 2), ("i16p_3ch", 4), ("u16p_stereo", 1), ("f32i_3ch", 3), ("u16p_stereo", 3)])
def test_one_shot_every_resource_type(gpu, oracle, name, n_out):
    """Play once to the end: conversion (pcm_i16_to_f32 / pcm_u16_to_f32), channel mapping (mono duplicated into a stereo
    node, surplus channels zeroed + flagged), the zero tail of the last block and the cleared blocks after it."""
    V, F = 5, 128
    def start(g, ids, res):
        g.sampler_set_sample(ids["smp"], res[name], True)
        g.sampler_play(ids["smp"])
    outs = both(gpu, oracle, V, n_out, F, [(None, F), (start, 4 * F), (None, 8 * F), (None, 8 * F)])
    assert np.any(outs[1][0] != 0) and np.all(outs[0][0] == 0)


def test_transport_messages_and_loops(gpu, oracle):
    """Pause / stop / set_playhead / loop ranges, different per voice, applied at call boundaries in push order."""
    V, F = 12, 64
    sm = lambda ids: ids["smp"]

    def a0(g, ids, res):
        names = list(res)
        for v in range(V):
            g.sampler_set_sample(sm(ids), res[names[v % len(names)]], False, voice=v)
        g.sampler_play(sm(ids))

    def a1(g, ids, res):
        for v in range(0, V, 3):
            g.sampler_pause(sm(ids), voice=v)
        g.sampler_set_loop_range(sm(ids), "full", voice=1)
        g.sampler_set_loop_range(sm(ids), (0.002, 0.0071), voice=2)  # frames 96 .. 341
        g.sampler_set_playhead(sm(ids), 0.001, voice=4)

    def a2(g, ids, res):
        for v in range(0, V, 3):
            g.sampler_play(sm(ids), voice=v)
        g.sampler_stop(sm(ids), voice=5)
        g.sampler_set_loop_range(sm(ids), "full", voice=7)
        g.sampler_set_sample(sm(ids), res["tiny"], False, voice=7)   # loop shorter than a block
        g.sampler_set_sample(sm(ids), res["f32p_mono"], True, voice=8)
        g.sampler_play(sm(ids), voice=8)   # node side still thinks it plays: no message (sampler.rs:82-98)

    def a3(g, ids, res):
        g.sampler_set_loop_range(sm(ids), None, voice=1)
        g.sampler_play(sm(ids), voice=5)
        g.sampler_pause(sm(ids), voice=8); g.sampler_play(sm(ids), voice=8)
        g.sampler_set_percent_volume(sm(ids), 50.0, voice=2)
        g.sampler_set_percent_volume(sm(ids), 0.0, voice=3)   # ramps to zero; Q2: never "muted" again after a ramp

    steps = [(a0, 5 * F), (a1, 7 * F), (a2, 6 * F + 17), (a3, 9 * F), (None, 20 * F), (None, 3)]
    both(gpu, oracle, V, 2, F, steps)
    both(gpu, oracle, V, 2, F, steps, stray=True)


def test_muted_sampler_keeps_its_playhead(gpu, oracle):
    """gain < 1e-5 with an idle smoother clears the outputs and does NOT advance the playhead (sampler.rs:437-443)."""
    V, F = 3, 100
    def a0(g, ids, res):
        g.sampler_set_percent_volume(ids["smp"], 0.0, voice=1)
        g.sampler_set_sample(ids["smp"], res["f32p_stereo"], True)
        g.sampler_play(ids["smp"])
    def a1(g, ids, res):
        g.sampler_set_percent_volume(ids["smp"], 80.0, voice=1)
    both(gpu, oracle, V, 2, F, [(a0, 3 * F), (a1, 4 * F), (None, 6 * F)])


@pytest.mark.parametrize("bus,stray", [(False, False), (True, False), (False, True), (True, True)])
def test_sampler_feeding_a_voice_graph(gpu, oracle, bus, stray):
    """Config-5 shape in miniature: sampler -> gain -> pan (-> master bus); the sample ends mid-call, so the gain node sees
    its input go silent (smoother reset, cleared outputs) in the middle of a call."""
    V, F = 70, 64
    rng = np.random.default_rng(5)
    pct = (20 + 80 * rng.random(V)).astype(f32)

    def post(g, smp, ids):
        vol = g.add_node(2, 2, VolumeNode(100.0)); pan = g.add_node(2, 2, PanNode(0.3))
        for c in range(2):
            g.connect(smp, c, vol, c, False); g.connect(vol, c, pan, c, False)
        ids["vol"], ids["pan"] = vol, pan
        return pan

    def a0(g, ids, res):
        names = ["f32p_stereo", "f32i_stereo", "i16i_stereo", "u16p_stereo", "f32p_mono"]
        for v in range(V):
            g.sampler_set_sample(ids["smp"], res[names[v % 5]], False, voice=v)
            if v % 4 == 0:
                g.sampler_set_loop_range(ids["smp"], "full", voice=v)
        g.sampler_play(ids["smp"])
        g.set_percent_volume(ids["vol"], pct)

    def a1(g, ids, res):
        g.set_percent_volume(ids["vol"], pct[::-1].copy())
        for v in range(0, V, 7):
            g.sampler_stop(ids["smp"], voice=v)

    def a2(g, ids, res):
        g.sampler_play(ids["smp"])

    both(gpu, oracle, V, 2, F, [(a0, 12 * F), (a1, 30 * F), (a2, 8 * F), (None, 40 * F)], bus=bus, post=post, stray=stray)


def test_two_samplers_summed(gpu, oracle):
    V, F = 9, 32
    def post(g, smp, ids):
        smp2 = g.add_node(0, 1, SamplerNode(70.0)); mix = g.add_node(4, 2, SumNode()); m2s_src = smp2
        ids["smp2"] = smp2
        g.connect(smp, 0, mix, 0, False); g.connect(smp, 1, mix, 1, False)
        g.connect(m2s_src, 0, mix, 2, False)  # mix port 1 right channel stays unconnected
        return mix
    def a0(g, ids, res):
        g.sampler_set_sample(ids["smp"], res["u16p_stereo"], True); g.sampler_play(ids["smp"])
        g.sampler_set_sample(ids["smp2"], res["f32p_mono"], True); g.sampler_set_loop_range(ids["smp2"], (0.0, 0.004))
    def a1(g, ids, res):
        g.sampler_play(ids["smp2"])
    both(gpu, oracle, V, 2, F, [(a0, 10 * F), (a1, 80 * F), (None, 10 * F)], post=post)


# the oracle's known-answer tests (hand-derived from the reference source) hold for the device path as well
import test_sampler_oracle as kat  # noqa: E402


@pytest.mark.parametrize("fn", [kat.test_pcm_conversions, kat.test_one_shot_end_and_zero_tail, kat.test_loop_wrap_inside_and_across_blocks,
                                kat.test_channel_mapping_and_masks, kat.test_message_semantics, kat.test_ring_capacity_and_activation],
                         ids=lambda f: f.__name__)
def test_known_answers_on_device(gpu, fn):
    fn(gpu)
