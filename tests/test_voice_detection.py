"""Isomorphic-voice detection (include/fw_b200.h graph_detect_voices / ctx_new_batched) — CPU only.

The reference runs ONE graph: a mixer of V identical voices is V copies of a sub-graph under a tree of SumNodes. Three things are
checked here without a GPU:
  1. the batching extension itself: `num_voices = V, master_bus = 1` on the CPU oracle is bit-identical to the restated reference
     executor running the FLAT graph with the explicit SumNode tree (sum.rs:58-81) — the bus IS that tree;
  2. the product's host logic finds the voices of a flat graph (any node creation order), maps template nodes to their copies,
     and refuses / falls back with a reason when the voices are not disjoint, not isomorphic or the tree is not the canonical one;
  3. the batched context it builds (graph + per-voice parameter tables), cloned node by node onto the oracle, reproduces the flat
     graph's output bit for bit.
"""
import numpy as np
import pytest

from conftest import synth
from firewheel_b200 import (AudioGraphConfig, BiquadNode, DelayNode, FirewheelGraphCtx, HardClipNode, PanNode, SumNode, SvfNode, VolumeNode,
                            design_rbj, design_svf)
from firewheel_b200 import _capi as K
from helpers import SR, assert_bit_exact, f32, run_planar

F = 64


def voice_params(V, seed):
    rng = np.random.default_rng(seed)
    return dict(pct=(20 + 100 * rng.random(V)).astype(f32), pan=rng.uniform(-1, 1, V).astype(f32),
                fc=rng.uniform(300, 5000, V), wet=(10 + 60 * rng.random(V)).astype(f32))


def build_voice(lib, g, prm, v, src, kind, records, batched=False):
    """One voice on graph `g`, reading the two channels `src` = [(node, port)] * 2; returns its two output endpoints. `v` = the voice
    whose parameters to use, or None with batched=True: per-voice tables."""
    sel = (lambda a: a) if batched else (lambda a: a[v])
    bq_co = lambda vv: np.array([design_rbj(lib, 0, float(prm["fc"][vv]), 0.8, 0.0, SR), design_rbj(lib, 4, 2000.0, 1.1, 3.0, SR)], f32)
    sv_co = lambda vv: np.array([design_svf(lib, 0, float(prm["fc"][vv]) * 0.5, 0.9, SR)], f32)
    V = len(prm["pct"])

    def add(ctor, ni, no):
        nid = g.add_node(ni, no, ctor())
        records[nid] = (ctor, ni, no)
        return nid

    def wire(srcs, dst):
        for c, (n, p) in enumerate(srcs):
            g.connect(n, p, dst, c, False)
        return dst
    vol = wire(src, add(lambda: VolumeNode(100.0), 2, 2))
    g.set_percent_volume(vol, sel(prm["pct"]) if batched else float(prm["pct"][v]))
    pan = wire([(vol, 0), (vol, 1)], add(lambda: PanNode(0.0), 2, 2))
    g.set_pan(pan, sel(prm["pan"]) if batched else float(prm["pan"][v]))
    if kind == "chain":  # gain -> pan -> biquad -> delay
        bq = wire([(pan, 0), (pan, 1)], add(lambda: BiquadNode(2), 2, 2))
        g.set_biquad_coeffs(bq, np.stack([bq_co(vv) for vv in range(V)]) if batched else bq_co(v)[None])
        dl = wire([(bq, 0), (bq, 1)], add(lambda: DelayNode(17), 2, 2))
        return [(dl, 0), (dl, 1)]
    # "drywet": gain -> pan -> {dry clip | svf -> wet gain} -> the voice's OWN 2-port SumNode (a SumNode that is not part of the bus tree)
    clip = wire([(pan, 0), (pan, 1)], add(lambda: HardClipNode(-3.0), 2, 2))
    sv = wire([(pan, 1), (pan, 0)], add(lambda: SvfNode(1), 2, 2))  # channels swapped on purpose
    g.set_svf_coeffs(sv, np.stack([sv_co(vv) for vv in range(V)]) if batched else sv_co(v)[None])
    wet = wire([(sv, 0), (sv, 1)], add(lambda: VolumeNode(100.0), 2, 2))
    g.set_percent_volume(wet, sel(prm["wet"]) if batched else float(prm["wet"][v]))
    mix = wire([(clip, 0), (clip, 1), (wet, 0), (wet, 1)], add(lambda: SumNode(), 4, 2))
    return [(mix, 0), (mix, 1)]


def build_tree(g, leaves, records=None, carries=True):
    """the canonical bus tree over `leaves` (each 2 endpoints): pairs (2i, 2i+1) through a 4 -> 2 SumNode, an unpaired last element through a
    2 -> 2 SumNode (the 1-port copy path) — or straight up with carries=False (NOT canonical)"""
    level = leaves
    while len(level) > 1:
        nxt = []
        for i in range(0, len(level) - 1, 2):
            s = g.add_node(4, 2, SumNode())
            for c in range(2):
                g.connect(level[i][c][0], level[i][c][1], s, c, False)
                g.connect(level[i + 1][c][0], level[i + 1][c][1], s, 2 + c, False)
            nxt.append([(s, 0), (s, 1)])
        if len(level) % 2:
            if carries:
                s = g.add_node(2, 2, SumNode())
                for c in range(2):
                    g.connect(level[-1][c][0], level[-1][c][1], s, c, False)
                nxt.append([(s, 0), (s, 1)])
            else:
                nxt.append(level[-1])
        level = nxt
    return level[0]


def build_flat(lib, V, kind, seed, order=None, carries=True, mutate=None):
    """The reference-style flat graph: V voices (created in `order`) + the SumNode tree. Returns ctx, {voice: [node ids in creation order]}, records."""
    prm = voice_params(V, seed)
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2 * V, num_graph_outputs=2))
    g = cx.graph
    records, leaves, voice_ids = {}, {}, {}
    for v in (order if order is not None else range(V)):
        before = set(records)
        leaves[v] = build_voice(lib, g, prm, v, [(g.graph_in_node(), 2 * v), (g.graph_in_node(), 2 * v + 1)], kind, records)
        voice_ids[v] = [n for n in records if n not in before]
    if mutate:
        mutate(g, voice_ids, leaves)
    root = build_tree(g, [leaves[v] for v in range(V)], carries=carries) if V > 1 else leaves[0]
    for c in range(2):
        g.connect(root[c][0], root[c][1], g.graph_out_node(), c, False)
    return cx, voice_ids, records


def run(cx, x, n_in, bus, calls=3):
    proc = cx.activate(SR, n_in, 2, F)
    assert proc is not None
    st = cx.update()
    assert st.graph_error is None, (st, cx.last_error())
    outs = [run_planar(proc, x, 2, bus)[0] for _ in range(calls)]
    proc.free(); cx.update()
    return outs


# ---- 1. the bus is the tree ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("V,kind", [(2, "chain"), (3, "chain"), (5, "drywet"), (8, "chain"), (13, "drywet")])
def test_master_bus_equals_the_flat_reference_graph_with_a_sum_tree(oracle, V, kind):
    T = 5 * F + 9
    x = synth((V, 2, T), 300 + V)
    x[V // 2, :, :] = 0.0  # a silent voice: the tree's SumNodes see flagged inputs (sum.rs:58-65 copy path, silence masks)
    flat, _, _ = build_flat(oracle, V, kind, seed=V)
    y_flat = run(flat, x.reshape(1, 2 * V, T), 2 * V, False)
    prm = voice_params(V, V)
    bcx = FirewheelGraphCtx(oracle, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V, master_bus=True))
    g = bcx.graph
    out = build_voice(oracle, g, prm, None, [(g.graph_in_node(), 0), (g.graph_in_node(), 1)], kind, {}, batched=True)
    for c in range(2):
        g.connect(out[c][0], out[c][1], g.graph_out_node(), c, False)
    y_bus = run(bcx, x, 2, True)
    for i, (a, b) in enumerate(zip(y_flat, y_bus)):
        assert_bit_exact(b, a[0], f"V={V} call {i}")
    flat.free(); bcx.free()


# ---- 2. detection ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("V,kind", [(1, "chain"), (2, "drywet"), (3, "chain"), (5, "drywet"), (8, "chain"), (13, "drywet"), (32, "chain")])
def test_voices_are_found_in_any_creation_order(product, V, kind):
    order = list(np.random.default_rng(V).permutation(V))
    cx, voice_ids, records = build_flat(product, V, kind, seed=V, order=order)
    t = cx.graph.detect_voices()
    n_per_voice = len(voice_ids[0])
    assert (t.num_voices, t.num_template_nodes, t.voice_inputs, t.voice_outputs) == (V, n_per_voice, 2, 2)
    n_tree = 0
    n = V
    while n > 1:
        n = (n + 1) // 2; n_tree += n
    assert t.num_tree_nodes == n_tree
    seen = [set() for _ in range(V)]
    for i in range(t.num_template_nodes):
        ids = cx.graph.voice_nodes(i)
        assert len(ids) == V
        ctor0 = records[ids[0]]
        for v, nid in enumerate(ids):
            assert nid in voice_ids[v], f"template node {i}: {nid} is not a node of voice {v} (leaf order = voice order)"
            assert records[nid][1:] == ctor0[1:]
            seen[v].add(nid)
    assert all(seen[v] == set(voice_ids[v]) for v in range(V))
    cx.free()


def test_fallbacks_and_refusals(product):
    V = 4

    def detect(**kw):
        cx, voice_ids, _ = build_flat(product, V, "chain", seed=1, **kw)
        t = cx.graph.detect_voices()
        why = cx.last_error()
        cx.free()
        return t, why
    # a different delay length in one voice: not isomorphic -> the whole graph is ONE voice (it still runs, through the generic lowering)
    def other_delay(g, voice_ids, leaves):
        dl = leaves[2][0][0]
        src = [(e.src_node, e.src_port) for e in map(g.edge, g.edges()) if e.dst_node == dl]
        g.remove_node(dl)
        nd = g.add_node(2, 2, DelayNode(18))
        for c, (n, p) in enumerate(sorted(src, key=lambda s: s[1])):
            g.connect(n, p, nd, c, False)
        leaves[2] = [(nd, 0), (nd, 1)]
    t, why = detect(mutate=other_delay)
    assert t.num_voices == 1 and "not isomorphic" in why
    # one voice taps another voice's gain: the voices are not disjoint
    def cross_feed(g, voice_ids, leaves):
        g.disconnect(voice_ids[1][0], 0, voice_ids[1][1], 0)
        g.connect(voice_ids[0][0], 0, voice_ids[1][1], 0, False)
    t, why = detect(mutate=cross_feed)
    assert t.num_voices == 1 and ("share a node" in why or "outside itself" in why)
    # a voice reading another voice's input channels
    def wrong_inputs(g, voice_ids, leaves):
        g.disconnect(g.graph_in_node(), 6, voice_ids[3][0], 0)
        g.connect(g.graph_in_node(), 0, voice_ids[3][0], 0, False)
    t, why = detect(mutate=wrong_inputs)
    assert t.num_voices == 1 and "graph_in" in why
    # an unpaired element wired straight up instead of through a 1-port SumNode: not the canonical tree (it would differ in the sign of zero)
    cx, _, _ = build_flat(product, 3, "chain", seed=1, carries=False)
    t = cx.graph.detect_voices()
    assert t.num_voices == 1 and "balanced pairwise tree" in cx.last_error()
    cx.free()
    # a linear accumulation ((v0 + v1) + v2) + v3 is not the bus tree either
    cx = FirewheelGraphCtx(product, AudioGraphConfig(num_graph_inputs=8, num_graph_outputs=2))
    g = cx.graph
    prm, rec = voice_params(4, 1), {}
    leaves = [build_voice(product, g, prm, v, [(g.graph_in_node(), 2 * v), (g.graph_in_node(), 2 * v + 1)], "chain", rec) for v in range(4)]
    acc = leaves[0]
    for v in range(1, 4):
        s = g.add_node(4, 2, SumNode())
        for c in range(2):
            g.connect(acc[c][0], acc[c][1], s, c, False); g.connect(leaves[v][c][0], leaves[v][c][1], s, 2 + c, False)
        acc = [(s, 0), (s, 1)]
    for c in range(2):
        g.connect(acc[c][0], acc[c][1], g.graph_out_node(), c, False)
    assert cx.graph.detect_voices().num_voices == 1
    cx.free()
    # a batched context is not a flat graph; new_batched needs a detection first
    cx = FirewheelGraphCtx(product, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=4))
    with pytest.raises(ValueError, match="flat graph"):
        cx.graph.detect_voices()
    cx.free()
    # the oracle does not batch
    import pyoracle
    ocx, _, _ = build_flat(pyoracle.load(), 2, "chain", seed=1)
    with pytest.raises(ValueError, match="does not batch"):
        ocx.graph.detect_voices()
    ocx.free()


# ---- 3. the batched context the product builds, cloned onto the oracle, equals the flat graph ---------------------------------------------
def clone_onto(lib, bcx, template_ids, ctor_of):
    """rebuild the product's batched context `bcx` on `lib` (the oracle) from what the C ABI shows: nodes, ports, edges, parameter tables"""
    src = bcx.graph
    cfg = bcx.config
    ocx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=cfg.num_graph_inputs, num_graph_outputs=cfg.num_graph_outputs, num_voices=cfg.num_voices,
                                                  master_bus=cfg.master_bus))
    g = ocx.graph
    ids = {src.graph_in_node(): g.graph_in_node(), src.graph_out_node(): g.graph_out_node()}
    V = cfg.num_voices
    for tid in template_ids:
        info = src.node_info(tid)
        nid = g.add_node(info.num_inputs, info.num_outputs, ctor_of[tid]())
        ids[tid] = nid
        name = info.debug_name if isinstance(info.debug_name, str) else info.debug_name.decode()
        if name == "volume":
            g.set_percent_volume(nid, src.read_params(tid, K.FW_PARAM_PERCENT_VOLUME))
        elif name == "pan":
            g.set_pan(nid, src.read_params(tid, K.FW_PARAM_PAN))
        elif name == "biquad":
            g.set_biquad_coeffs(nid, src.read_params(tid, K.FW_PARAM_COEFFS).reshape(V, -1, 5))
        elif name == "svf":
            g.set_svf_coeffs(nid, src.read_params(tid, K.FW_PARAM_COEFFS).reshape(V, -1, 6))
    assert set(src.nodes()) == set(ids), "the batched graph holds exactly graph_in, graph_out and the template nodes"
    for e in map(src.edge, src.edges()):
        g.connect(ids[e.src_node], e.src_port, ids[e.dst_node], e.dst_port, False)
    return ocx


@pytest.mark.parametrize("V,kind", [(1, "drywet"), (2, "chain"), (3, "drywet"), (6, "chain"), (8, "drywet"), (11, "chain")])
def test_the_batched_context_reproduces_the_flat_graph(oracle, product, V, kind):
    order = list(np.random.default_rng(100 + V).permutation(V))
    T = 4 * F + 5
    x = synth((V, 2, T), 500 + V)
    oflat, _, _ = build_flat(oracle, V, kind, seed=V, order=order)
    y_flat = run(oflat, x.reshape(1, 2 * V, T), 2 * V, False)
    pflat, _, records = build_flat(product, V, kind, seed=V, order=order)  # same call sequence: same ids as on the oracle
    bcx, template_ids = FirewheelGraphCtx.new_batched(pflat)
    assert bcx.config.num_voices == V and bcx.config.master_bus == (V > 1)
    ctor_of = {tid: records[pflat.graph.voice_nodes(i)[0]][0] for i, tid in enumerate(template_ids)}
    ocx = clone_onto(oracle, bcx, template_ids, ctor_of)
    y_b = run(ocx, x, 2, V > 1)
    for i, (a, b) in enumerate(zip(y_flat, y_b)):
        assert_bit_exact(b if V > 1 else b[0], a[0], f"V={V} call {i}")
    # the per-voice tables are the copies' values in leaf order
    prm = voice_params(V, V)
    vol_tid = template_ids[[pflat.graph.node_info(pflat.graph.voice_nodes(i)[0]).debug_name for i in range(len(template_ids))].index(
        pflat.graph.node_info(pflat.graph.voice_nodes(0)[0]).debug_name)]
    tables = [bcx.graph.read_params(t, K.FW_PARAM_PERCENT_VOLUME) for t in template_ids]
    assert any(len(t) == V and np.array_equal(t, prm["pct"]) for t in tables), "the first gain's table is pct in voice order"
    del vol_tid
    for c in (oflat, pflat, bcx, ocx):
        c.free()


def test_source_voices_without_inputs(product):
    """config 5's shape as a flat graph: every voice starts at its own SamplerNode (graph_in has no channels), gain -> pan behind it;
    the sampler's volume and the gain / pan tables are gathered per voice"""
    from firewheel_b200 import SamplerNode
    V = 6
    prm = voice_params(V, 3)
    cx = FirewheelGraphCtx(product, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=2))
    g = cx.graph
    leaves, smp_ids = [], []
    for v in range(V):
        smp = g.add_node(0, 2, SamplerNode(float(50 + v)))
        vol, pan = g.add_node(2, 2, VolumeNode(float(prm["pct"][v]))), g.add_node(2, 2, PanNode(float(prm["pan"][v])))
        for c in range(2):
            g.connect(smp, c, vol, c, False); g.connect(vol, c, pan, c, False)
        leaves.append([(pan, 0), (pan, 1)]); smp_ids.append(smp)
    root = build_tree(g, leaves)
    for c in range(2):
        g.connect(root[c][0], root[c][1], g.graph_out_node(), c, False)
    bcx, tids = FirewheelGraphCtx.new_batched(cx)
    assert bcx.config.num_voices == V and bcx.config.num_graph_inputs == 0 and len(tids) == 3
    by_name = {}
    for tid in tids:
        info = bcx.graph.node_info(tid)
        by_name[info.debug_name if isinstance(info.debug_name, str) else info.debug_name.decode()] = tid
    assert cx.graph.voice_nodes(tids.index(by_name["beep_test"])) == smp_ids  # Q8: the sampler's debug name (sampler.rs:186)
    assert np.array_equal(bcx.graph.read_params(by_name["beep_test"], K.FW_PARAM_PERCENT_VOLUME), np.arange(50, 50 + V, dtype=f32))
    assert np.array_equal(bcx.graph.read_params(by_name["volume"], K.FW_PARAM_PERCENT_VOLUME), prm["pct"])
    assert np.array_equal(bcx.graph.read_params(by_name["pan"], K.FW_PARAM_PAN), prm["pan"])
    sched, _ = bcx.graph.compile_internal(F)
    assert len(sched) == 5  # graph_in, sampler, gain, pan, graph_out: the chain the fused device path takes (config 5's head)
    bcx.free(); cx.free()


@pytest.mark.parametrize("V", [3, 4, 7])
def test_mono_voices(oracle, product, V):
    """C = 1: the tree's SumNodes are 2 -> 1, the carry is a 1 -> 1 SumNode; voice = gain -> hard clip -> 3-stage biquad"""
    T = 3 * F + 2
    x = synth((V, 1, T), 40 + V)
    pct = (30 + 15 * np.arange(V)).astype(f32)
    co = np.stack([[design_rbj(product, s % 2, 400.0 * (v + 1) + 100 * s, 0.7, 0.0, SR) for s in range(3)] for v in range(V)]).astype(f32)

    def flat_on(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=V, num_graph_outputs=1))
        g = cx.graph
        level = []
        for v in reversed(range(V)):  # created back to front: leaf order, not creation order, numbers the voices
            vol, clip, bq = g.add_node(1, 1, VolumeNode(float(pct[v]))), g.add_node(1, 1, HardClipNode(-2.0)), g.add_node(1, 1, BiquadNode(3))
            g.set_biquad_coeffs(bq, co[v][None])
            g.connect(g.graph_in_node(), v, vol, 0, False); g.connect(vol, 0, clip, 0, False); g.connect(clip, 0, bq, 0, False)
            level.insert(0, bq)
        while len(level) > 1:
            nxt = []
            for i in range(0, len(level) - 1, 2):
                s_ = g.add_node(2, 1, SumNode())
                g.connect(level[i], 0, s_, 0, False); g.connect(level[i + 1], 0, s_, 1, False)
                nxt.append(s_)
            if len(level) % 2:
                s_ = g.add_node(1, 1, SumNode())
                g.connect(level[-1], 0, s_, 0, False)
                nxt.append(s_)
            level = nxt
        g.connect(level[0], 0, g.graph_out_node(), 0, False)
        return cx

    def run1(cx, xin, n_in, bus):
        proc = cx.activate(SR, n_in, 1, F)
        assert cx.update().graph_error is None, cx.last_error()
        y = run_planar(proc, xin, 1, bus)[0]
        proc.free(); cx.update()
        return y
    oflat = flat_on(oracle)
    y_flat = run1(oflat, x.reshape(1, V, T), V, False)
    pflat = flat_on(product)
    t = pflat.graph.detect_voices()
    assert (t.num_voices, t.num_template_nodes, t.voice_inputs, t.voice_outputs) == (V, 3, 1, 1)
    bcx, tids = FirewheelGraphCtx.new_batched(pflat)
    tables = {(bcx.graph.node_info(tid).debug_name if isinstance(bcx.graph.node_info(tid).debug_name, str) else bcx.graph.node_info(tid).debug_name.decode()): tid for tid in tids}
    assert np.array_equal(bcx.graph.read_params(tables["volume"], K.FW_PARAM_PERCENT_VOLUME), pct)
    assert np.array_equal(bcx.graph.read_params(tables["biquad"], K.FW_PARAM_COEFFS).reshape(V, 3, 5), co)
    # the same batched form on the oracle
    ocx = FirewheelGraphCtx(oracle, AudioGraphConfig(num_graph_inputs=1, num_graph_outputs=1, num_voices=V, master_bus=True))
    g = ocx.graph
    vol, clip, bq = g.add_node(1, 1, VolumeNode(100.0)), g.add_node(1, 1, HardClipNode(-2.0)), g.add_node(1, 1, BiquadNode(3))
    g.set_percent_volume(vol, pct); g.set_biquad_coeffs(bq, co)
    g.connect(g.graph_in_node(), 0, vol, 0, False); g.connect(vol, 0, clip, 0, False); g.connect(clip, 0, bq, 0, False); g.connect(bq, 0, g.graph_out_node(), 0, False)
    assert_bit_exact(run1(ocx, x, 1, True), y_flat[0], f"mono V={V}")
    for c in (oflat, pflat, bcx, ocx):
        c.free()
