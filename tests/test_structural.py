"""The reference's own five unit tests (crates/firewheel-graph/src/graph/compiler/schedule.rs:392-711),
re-implemented line for line against BOTH the CPU oracle and the product's host-side compiler.
These are the only tests the reference holds for this path, so they are what pins the oracle.
The product's graph/compile code is host-only until `activate`, so these run without a GPU.
"""
import pytest

from firewheel_b200 import (AddEdgeError, AudioGraphConfig, DummyAudioNode, FirewheelGraphCtx)


@pytest.fixture(params=["oracle", "product"])
def lib(request):
    return request.getfixturevalue(request.param)


def new_graph(lib, **kw):
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(**kw))
    return cx, cx.graph


def find(schedule, node_id):
    return next(s for s in schedule if s.id == node_id)


def verify_node(graph, schedule, node_id, in_ports_that_should_clear):  # schedule.rs:600-635
    info = graph.node_info(node_id)
    sn = find(schedule, node_id)
    assert len(sn.input_buffers) == info.num_inputs
    assert len(sn.output_buffers) == info.num_outputs
    assert len(in_ports_that_should_clear) == info.num_inputs
    for (buf, should_clear), expect in zip(sn.input_buffers, in_ports_that_should_clear):
        assert should_clear == expect
    seen = set()
    for buf, _ in sn.input_buffers:
        assert buf not in seen
        seen.add(buf)
    for buf in sn.output_buffers:
        assert buf not in seen
        seen.add(buf)


def verify_edge(graph, schedule, edge_id):  # schedule.rs:637-660
    e = graph.edge(edge_id)
    src = find(schedule, e.src_node).output_buffers[e.src_port]
    dst = find(schedule, e.dst_node).input_buffers[e.dst_port][0]
    assert src == dst


def test_simplest_graph_compile(lib):  # schedule.rs:407-436
    cx, graph = new_graph(lib, num_graph_inputs=1, num_graph_outputs=1)
    node0, node1 = graph.graph_in_node(), graph.graph_out_node()
    edge0 = graph.connect(node0, 0, node1, 0, False)
    schedule, num_buffers = graph.compile_internal(128)
    assert len(schedule) == 2
    assert num_buffers * 128 > 0
    assert schedule[0].id == node0
    assert schedule[1].id == node1
    verify_node(graph, schedule, node0, [])
    verify_node(graph, schedule, node1, [False])
    verify_edge(graph, schedule, edge0)


def test_graph_compile_1(lib):  # schedule.rs:451-524
    cx, graph = new_graph(lib, num_graph_inputs=2, num_graph_outputs=2)
    node0 = graph.graph_in_node()
    node1 = graph.add_node(1, 2, DummyAudioNode())
    node2 = graph.add_node(1, 1, DummyAudioNode())
    node3 = graph.add_node(2, 2, DummyAudioNode())
    node4 = graph.add_node(2, 2, DummyAudioNode())
    node5 = graph.add_node(5, 2, DummyAudioNode())
    node6 = graph.graph_out_node()
    edges = [graph.connect(node0, 0, node1, 0, False), graph.connect(node0, 1, node2, 0, False),
             graph.connect(node1, 0, node3, 0, False), graph.connect(node1, 1, node4, 1, False),
             graph.connect(node3, 0, node5, 0, False), graph.connect(node3, 1, node5, 1, False),
             graph.connect(node4, 0, node5, 2, False), graph.connect(node4, 1, node5, 3, False),
             graph.connect(node2, 0, node5, 4, False), graph.connect(node5, 0, node6, 0, False),
             graph.connect(node5, 1, node6, 1, False)]
    schedule, num_buffers = graph.compile_internal(128)
    assert len(schedule) == 7
    assert num_buffers * 128 > 6  # "Node 5 needs at-least 7 buffers" (len of the flat pool)
    assert num_buffers >= 7
    assert schedule[0].id == node0
    assert {schedule[1].id, schedule[2].id} == {node1, node2}
    assert {schedule[3].id, schedule[4].id} == {node3, node4}
    assert schedule[5].id == node5
    assert schedule[6].id == node6
    verify_node(graph, schedule, node0, [])
    verify_node(graph, schedule, node1, [False])
    verify_node(graph, schedule, node2, [False])
    verify_node(graph, schedule, node3, [False, True])
    verify_node(graph, schedule, node4, [True, False])
    verify_node(graph, schedule, node5, [False] * 5)
    verify_node(graph, schedule, node6, [False, False])
    for e in edges:
        verify_edge(graph, schedule, e)


def test_graph_compile_2(lib):  # schedule.rs:539-598
    cx, graph = new_graph(lib, num_graph_inputs=2, num_graph_outputs=2)
    node0 = graph.graph_in_node()
    node1 = graph.add_node(1, 1, DummyAudioNode())
    node2 = graph.add_node(2, 2, DummyAudioNode())
    node3 = graph.add_node(2, 2, DummyAudioNode())
    node4 = graph.add_node(5, 4, DummyAudioNode())
    node5 = graph.graph_out_node()
    node6 = graph.add_node(1, 1, DummyAudioNode())
    edges = [graph.connect(node0, 0, node2, 0, False), graph.connect(node0, 0, node3, 1, False),
             graph.connect(node2, 0, node4, 0, False), graph.connect(node3, 1, node4, 3, False),
             graph.connect(node1, 0, node4, 4, False), graph.connect(node4, 0, node5, 0, False),
             graph.connect(node4, 2, node6, 0, False)]
    schedule, num_buffers = graph.compile_internal(128)
    assert len(schedule) == 7
    assert num_buffers >= 8  # "Node 4 needs at-least 8 buffers"
    assert {schedule[0].id, schedule[1].id} == {node0, node1}
    assert {schedule[2].id, schedule[3].id} == {node2, node3}
    assert schedule[4].id == node4
    assert {schedule[5].id, schedule[6].id} == {node5, node6}
    for e in edges:
        verify_edge(graph, schedule, e)
    verify_node(graph, schedule, node0, [])
    verify_node(graph, schedule, node1, [True])
    verify_node(graph, schedule, node2, [False, True])
    verify_node(graph, schedule, node3, [True, False])
    verify_node(graph, schedule, node4, [False, True, True, False, False])
    verify_node(graph, schedule, node5, [False, True])
    verify_node(graph, schedule, node6, [False])


def test_many_to_one_detection(lib):  # schedule.rs:662-683
    cx, graph = new_graph(lib, num_graph_inputs=2, num_graph_outputs=1)
    node1, node2 = graph.graph_in_node(), graph.graph_out_node()
    graph.connect(node1, 0, node2, 0, False)
    with pytest.raises(AddEdgeError) as ei:
        graph.connect(node1, 1, node2, 0, False)
    assert ei.value.kind == "InputPortAlreadyConnected"
    assert ei.value.node == node2
    assert ei.value.port == 0


def test_cycle_detection(lib):  # schedule.rs:685-710
    cx, graph = new_graph(lib, num_graph_inputs=0, num_graph_outputs=2)
    node1 = graph.add_node(1, 1, DummyAudioNode())
    node2 = graph.add_node(2, 1, DummyAudioNode())
    node3 = graph.add_node(1, 1, DummyAudioNode())
    graph.connect(node1, 0, node2, 0, False)
    graph.connect(node2, 0, node3, 0, False)
    edge3 = graph.connect(node3, 0, node1, 0, False)
    assert graph.cycle_detected()
    graph.disconnect_by_edge_id(edge3)
    assert not graph.cycle_detected()
    graph.connect(node3, 0, node2, 1, False)
    assert graph.cycle_detected()


# ---- further pins on graph.rs semantics (same on both implementations) ----------------------
def test_connect_errors(lib):  # graph.rs:396-446
    cx, graph = new_graph(lib, num_graph_inputs=1, num_graph_outputs=1)
    gin, gout = graph.graph_in_node(), graph.graph_out_node()
    n = graph.add_node(1, 1, DummyAudioNode())
    with pytest.raises(AddEdgeError) as ei:
        graph.connect(gin, 1, n, 0, False)
    assert ei.value.kind == "OutPortOutOfRange"
    with pytest.raises(AddEdgeError) as ei:
        graph.connect(gin, 0, n, 3, False)
    assert ei.value.kind == "InPortOutOfRange"
    with pytest.raises(AddEdgeError) as ei:
        graph.connect(n, 0, n, 0, False)
    assert ei.value.kind == "CycleDetected"
    graph.connect(gin, 0, n, 0, False)
    with pytest.raises(AddEdgeError) as ei:
        graph.connect(gin, 0, n, 0, False)
    assert ei.value.kind == "EdgeAlreadyExists"
    removed = graph.remove_node(n)
    assert len(removed) == 1
    with pytest.raises(AddEdgeError) as ei:
        graph.connect(gin, 0, n, 0, False)
    assert ei.value.kind == "DstNodeNotFound"
    with pytest.raises(AddEdgeError) as ei:
        graph.connect(n, 0, gout, 0, False)
    assert ei.value.kind == "SrcNodeNotFound"
    with pytest.raises(KeyError):
        graph.remove_node(gin)  # Err(()) graph.rs:269-271


def test_arena_slot_reuse_and_generations(lib):  # thunderdome semantics visible through NodeID
    cx, graph = new_graph(lib)
    a = graph.add_node(1, 1, DummyAudioNode())
    b = graph.add_node(1, 1, DummyAudioNode())
    assert (a.slot, a.generation) == (2, 1) and (b.slot, b.generation) == (3, 1)
    graph.remove_node(a)
    graph.remove_node(b)
    c = graph.add_node(1, 1, DummyAudioNode())  # LIFO free list: slot 3 first
    d = graph.add_node(1, 1, DummyAudioNode())
    assert (c.slot, c.generation) == (3, 2) and (d.slot, d.generation) == (2, 2)
    assert graph.node_info(a) is None and graph.node_info(c) is not None
    assert [n.slot for n in graph.nodes()] == [0, 1, 2, 3]


def test_check_for_cycles_leaves_port_marked(lib):  # Q9: graph.rs:466-472 stale bookkeeping
    cx, graph = new_graph(lib, num_graph_inputs=0, num_graph_outputs=2)
    n1 = graph.add_node(1, 1, DummyAudioNode())
    n2 = graph.add_node(1, 1, DummyAudioNode())
    graph.connect(n1, 0, n2, 0, True)
    with pytest.raises(AddEdgeError) as ei:
        graph.connect(n2, 0, n1, 0, True)
    assert ei.value.kind == "CycleDetected"
    assert not graph.cycle_detected()  # the edge itself was removed again
    assert len(graph.edges()) == 1
    with pytest.raises(AddEdgeError) as ei:  # ... but the reference keeps the edge hash: "already exists"
        graph.connect(n2, 0, n1, 0, False)
    assert ei.value.kind == "EdgeAlreadyExists"


def test_set_num_ports_removes_edges(lib):  # graph.rs:315-375
    cx, graph = new_graph(lib, num_graph_inputs=2, num_graph_outputs=2)
    gin, gout = graph.graph_in_node(), graph.graph_out_node()
    n = graph.add_node(2, 2, DummyAudioNode())
    graph.connect(gin, 0, n, 0, False)
    e1 = graph.connect(gin, 1, n, 1, False)
    graph.connect(n, 0, gout, 0, False)
    e3 = graph.connect(n, 1, gout, 1, False)
    assert graph.set_num_inputs(n, 1) == [e1]
    assert graph.set_num_outputs(n, 1) == [e3]
    assert graph.node_info(n).num_inputs == 1 and graph.node_info(n).num_outputs == 1
    with pytest.raises(KeyError):
        graph.set_num_inputs(gin, 1)
    with pytest.raises(KeyError):
        graph.set_num_outputs(gout, 1)
    graph.connect(gin, 1, gout, 1, False)  # port freed again
    schedule, _ = graph.compile_internal(64)
    assert schedule[0].id == gin and schedule[-1].id == gout


def test_fanout_shares_one_buffer_and_reuse(lib):  # compiler.rs:387-399, :123-130
    cx, graph = new_graph(lib, num_graph_inputs=1, num_graph_outputs=2)
    gin, gout = graph.graph_in_node(), graph.graph_out_node()
    a = graph.add_node(1, 1, DummyAudioNode())
    b = graph.add_node(1, 1, DummyAudioNode())
    graph.connect(gin, 0, a, 0, False)
    graph.connect(gin, 0, b, 0, False)
    graph.connect(a, 0, gout, 0, False)
    graph.connect(b, 0, gout, 1, False)
    schedule, nb = graph.compile_internal(32)
    sa, sb = find(schedule, a), find(schedule, b)
    assert sa.input_buffers[0][0] == sb.input_buffers[0][0] == find(schedule, gin).output_buffers[0]
    assert nb == 3  # in, a.out, b.out — the shared input cannot be recycled before b has run
    assert sa.output_buffers[0] != sb.output_buffers[0]


def test_node_info_of_every_built_in_kind_matches(oracle, product):
    """AudioNodeInfo (node.rs:57-79) + debug_name of every built-in node kind: the product's host tables against the oracle's
    restated nodes (no GPU needed: nothing is activated)."""
    import numpy as np
    from firewheel_b200 import (AudioGraphConfig, BiquadNode, ConvReverbNode, DelayNode, DummyAudioNode, FirewheelGraphCtx, HardClipNode,
                                MonoToStereoNode, PanNode, ResamplerNode, SamplerNode, StereoToMonoNode, SumNode, SvfNode, VolumeNode)
    table = np.zeros((64, 16), np.float32)
    kinds = [(DummyAudioNode, 3, 2), (lambda: VolumeNode(50.0), 2, 2), (SumNode, 4, 2), (MonoToStereoNode, 1, 2), (StereoToMonoNode, 2, 1),
             (lambda: HardClipNode(-3.0), 2, 2), (lambda: PanNode(0.5), 2, 2), (lambda: BiquadNode(3), 2, 2), (lambda: DelayNode(100), 1, 1),
             (lambda: ConvReverbNode(np.ones((1, 8), np.float32)), 2, 2), (lambda: SamplerNode(90.0), 0, 2), (lambda: SvfNode(2), 2, 2),
             (lambda: ResamplerNode(table), 0, 1)]
    infos = []
    for lib in (oracle, product):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2))
        g = cx.graph
        rows = []
        for mk, ni, no in kinds:
            i = g.node_info(g.add_node(ni, no, mk()))
            rows.append((i.kind, i.num_inputs, i.num_outputs, i.num_min_supported_inputs, i.num_max_supported_inputs,
                         i.num_min_supported_outputs, i.num_max_supported_outputs, i.updates, i.debug_name))
        infos.append(rows)
        cx.free()
    assert infos[0] == infos[1]
