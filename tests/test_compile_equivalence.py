"""Host logic: the product's graph/compiler must produce the SAME schedule (node order, buffer indices,
should_clear flags, buffer count), ids and errors as the oracle's literal restatement of compiler.rs, on
randomly grown and mutated graphs. CPU only — the product does not touch CUDA before activate()."""
import random

import pytest

from firewheel_b200 import AddEdgeError, AudioGraphConfig, DummyAudioNode, FirewheelGraphCtx


def mutate_both(go, gp, rng, steps):
    """Apply the same random edit script to both graphs; ids must agree at every step."""
    nodes = [go.graph_in_node(), go.graph_out_node()]
    assert nodes == [gp.graph_in_node(), gp.graph_out_node()]
    edges = []
    for _ in range(steps):
        r = rng.random()
        if r < 0.30 or len(nodes) < 4:
            ni, no = rng.randint(0, 4), rng.randint(0, 4)
            a, b = go.add_node(ni, no, DummyAudioNode()), gp.add_node(ni, no, DummyAudioNode())
            assert a == b
            nodes.append(a)
        elif r < 0.80:
            s, d = rng.choice(nodes), rng.choice(nodes)
            sp, dp, chk = rng.randint(0, 4), rng.randint(0, 4), rng.random() < 0.5
            res = []
            for g in (go, gp):
                try:
                    res.append(("ok", g.connect(s, sp, d, dp, chk)))
                except AddEdgeError as e:
                    res.append((e.kind, e.node if e.kind in ("InputPortAlreadyConnected",) else None, e.port if e.kind == "InputPortAlreadyConnected" else None))
            assert res[0] == res[1], (res, s, sp, d, dp, chk)
            if res[0][0] == "ok":
                edges.append(res[0][1])
        elif r < 0.88 and edges:
            e = edges.pop(rng.randrange(len(edges)))
            assert go.disconnect_by_edge_id(e) == gp.disconnect_by_edge_id(e)
        elif r < 0.94 and len(nodes) > 2:
            n = nodes[rng.randrange(2, len(nodes))] if len(nodes) > 2 else None
            ro = rp = None
            try:
                ro = go.remove_node(n)
            except KeyError:
                ro = "err"
            try:
                rp = gp.remove_node(n)
            except KeyError:
                rp = "err"
            assert ro == rp
            if ro != "err":
                nodes.remove(n)
                edges = [e for e in edges if e not in ro]
        else:
            n = rng.choice(nodes)
            k = rng.randint(0, 4)
            fn = rng.choice(["set_num_inputs", "set_num_outputs"])
            out = []
            for g in (go, gp):
                try:
                    out.append(getattr(g, fn)(n, k))
                except KeyError:
                    out.append("err")
            assert out[0] == out[1]
            if out[0] != "err":
                edges = [e for e in edges if e not in out[0]]
        assert go.cycle_detected() == gp.cycle_detected()
        assert go.nodes() == gp.nodes() and go.edges() == gp.edges()


@pytest.mark.parametrize("seed", range(12))
def test_random_graph_schedules_match(oracle, product, seed):
    rng = random.Random(seed)
    cfg = AudioGraphConfig(num_graph_inputs=rng.randint(0, 3), num_graph_outputs=rng.randint(1, 3))
    co, cp = FirewheelGraphCtx(oracle, cfg), FirewheelGraphCtx(product, cfg)
    mutate_both(co.graph, cp.graph, rng, 120)
    results = []
    for g in (co.graph, cp.graph):
        try:
            results.append(g.compile_internal(64))
        except Exception as e:  # CompileGraphError
            results.append(("err", e.kind))
    assert results[0] == results[1]
