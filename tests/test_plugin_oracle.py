"""The plugin boundary (trait AudioNode / AudioNodeProcessor, node.rs:6-53, as include/fw_b200.h's fw_node_vtable) on the CPU
oracle: a user node compiled as its own shared library runs inside graphs of built-in nodes, its life cycle follows the
reference's (activate once, update() from ctx.update(), deactivate(Some(processor)) on roll-back and on removal while active,
drop otherwise), and its output equals the plugin's arithmetic restated in numpy."""
import numpy as np

import plugin_fixture as pf
from conftest import synth
from firewheel_b200 import AudioGraphConfig, CompileGraphError, FirewheelGraphCtx, VolumeNode
from helpers import assert_bit_exact, run_planar


def build(lib, V, F=64, k0=0.5, rule=pf.RULE_ALL_IF_ALL_INPUTS, fail=False, with_gain=True):
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V))
    g = cx.graph
    vt, node = pf.new_node(k0, rule, fail)
    cu = g.add_custom_node(2, 2, vt, node)
    prev = g.graph_in_node()
    if with_gain:
        vol = g.add_node(2, 2, VolumeNode(50.0))
        for c in range(2):
            g.connect(prev, c, vol, c, False)
        prev = vol
    for c in range(2):
        g.connect(prev, c, cu, c, False)
        g.connect(cu, c, g.graph_out_node(), c, False)
    proc = cx.activate(48000, 2, 2, F)
    return cx, proc, cu


def test_custom_node_runs_in_the_oracle_and_matches_its_arithmetic(oracle):
    pf.reset_counters()
    V, F = 5, 64
    cx, proc, cu = build(oracle, V, F)
    st = cx.update()
    assert st.graph_error is None, (st, cx.last_error())
    info = cx.graph.node_info(cu)
    assert info.debug_name == b"fir1_plugin" and info.kind == 13 and info.updates and (info.num_min_supported_inputs, info.num_max_supported_inputs) == (1, 8)
    x = synth((V, 2, 4 * F + 17), 3)
    y1, m1 = run_planar(proc, x, 2)
    y2, m2 = run_planar(proc, x, 2)   # state carries across calls
    g = np.float32(0.25)  # 50 % -> raw gain 0.25
    xs = (x * g).astype(np.float32)
    r1, s = pf.fir1_reference(xs, 0.5)
    r2, _ = pf.fir1_reference(xs, 0.5, s)
    assert_bit_exact(y1, r1, "call 1"); assert_bit_exact(y2, r2, "call 2")
    assert m1 == 0 and m2 == 0
    c = pf.counters()
    assert c["activate"] == 1 and c["update"] >= 1 and c["deactivate"] == 0 and c["drop_node"] == 0, c  # one activation for all voices
    proc.free(); cx.update(); cx.free()
    c = pf.counters()
    assert c["drop_node"] == 1 and c["drop_processor"] + c["deactivate"] == 1, c


def test_silent_inputs_follow_the_declared_rule(oracle):
    """unconnected inputs are cleared and flagged by the executor (schedule.rs:310-313): the node sees all-silent masks"""
    cx = FirewheelGraphCtx(oracle, AudioGraphConfig(num_graph_inputs=0, num_graph_outputs=2, num_voices=2))
    g = cx.graph
    vt, node = pf.new_node(0.25, pf.RULE_ALL_IF_ALL_INPUTS)
    cu = g.add_custom_node(2, 2, vt, node)
    for c in range(2):
        g.connect(cu, c, g.graph_out_node(), c, False)
    proc = cx.activate(48000, 0, 2, 32)
    assert cx.update().graph_error is None
    y, mask = run_planar(proc, np.zeros((2, 0, 96), np.float32), 2)
    assert not y.any() and mask == 3
    proc.free(); cx.update(); cx.free()


def test_activation_failure_is_reported_and_rolled_back(oracle):
    pf.reset_counters()
    cx, proc, cu = build(oracle, 3, fail=True)
    st = cx.update()
    assert st.graph_error is not None and st.graph_error.kind == "NodeActivationFailed" and st.graph_error.node == cu
    assert pf.counters()["activate"] == 0
    proc.free(); cx.update(); cx.free()
    assert pf.counters()["drop_node"] == 1
