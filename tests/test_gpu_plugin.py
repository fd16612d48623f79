"""The plugin boundary on the product: the same third-party node (tests/plugins/fir1_plugin.cu, built against include/fw_b200.h
only) runs through fw_node_vtable::process_device on the GPU and through ::process on the CPU oracle; outputs, silence masks
and life-cycle calls must agree."""
import numpy as np
import pytest

import plugin_fixture as pf
from conftest import synth
from firewheel_b200 import AudioGraphConfig, CompileGraphError, FirewheelGraphCtx, SumNode, VolumeNode
from helpers import assert_bit_exact, run_planar
from test_plugin_oracle import build

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("V,F,frames", [(1, 64, 256), (7, 64, 4 * 64 + 17), (130, 128, 1024)])
def test_custom_node_gpu_matches_oracle(gpu, oracle, V, F, frames):
    x = synth((V, 2, frames), 11)
    outs = []
    for lib in (gpu, oracle):
        pf.reset_counters()
        cx, proc, cu = build(lib, V, F)
        st = cx.update()
        assert st.graph_error is None, (st, cx.last_error())
        res = [run_planar(proc, x, 2) for _ in range(3)]
        assert pf.counters()["activate"] == 1
        proc.free(); cx.update(); cx.free()
        c = pf.counters()
        assert c["drop_node"] == 1 and c["drop_processor"] + c["deactivate"] == 1, c
        outs.append(res)
    for i, ((yg, mg), (yo, mo)) in enumerate(zip(*outs)):
        assert_bit_exact(yg, yo, f"call {i}")
        assert mg == mo


def test_custom_node_in_a_dag_with_silent_and_live_branches(gpu, oracle):
    """graph_in -> custom A -> sum <- custom B (unconnected inputs: all-silent masks every block) -> gain -> graph_out"""
    V, F = 9, 64
    x = synth((V, 2, 5 * F), 21)

    def run(lib):
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V))
        g = cx.graph
        a = g.add_custom_node(2, 2, *pf.new_node(0.3, pf.RULE_ALL_IF_ALL_INPUTS))
        b = g.add_custom_node(2, 2, *pf.new_node(0.7, pf.RULE_ALL_IF_ALL_INPUTS))
        mix, vol = g.add_node(4, 2, SumNode()), g.add_node(2, 2, VolumeNode(80.0))
        for c in range(2):
            g.connect(g.graph_in_node(), c, a, c, False)
            g.connect(a, c, mix, c, False); g.connect(b, c, mix, 2 + c, False)
            g.connect(mix, c, vol, c, False); g.connect(vol, c, g.graph_out_node(), c, False)
        proc = cx.activate(48000, 2, 2, F)
        assert cx.update().graph_error is None, cx.last_error()
        r = [run_planar(proc, x, 2) for _ in range(2)]
        proc.free(); cx.update(); cx.free()
        return r
    for i, ((yg, mg), (yo, mo)) in enumerate(zip(run(gpu), run(oracle))):
        assert_bit_exact(yg, yo, f"call {i}")
        assert mg == mo


def test_custom_node_removed_while_active_is_deactivated_with_its_processor(gpu):
    pf.reset_counters()
    V, F = 4, 64
    cx, proc, cu = build(gpu, V, F, with_gain=False)
    assert cx.update().graph_error is None
    x = synth((V, 2, 2 * F), 5)
    run_planar(proc, x, 2)
    g = cx.graph
    g.remove_node(cu)
    for c in range(2):
        g.connect(g.graph_in_node(), c, g.graph_out_node(), c, False)
    assert cx.update().graph_error is None          # new schedule without the node
    y, _ = run_planar(proc, x, 2)                   # the processor adopts it and returns the old plan
    cx.update()                                     # main thread: on_schedule_returned -> deactivate(Some(processor)) (graph.rs:644-648)
    c = pf.counters()
    assert c["deactivate"] == 1 and c["drop_node"] == 1 and c["drop_processor"] == 0, c
    proc.free(); cx.update(); cx.free()


def test_activation_failure_on_the_product(gpu):
    pf.reset_counters()
    cx, proc, cu = build(gpu, 3, fail=True)
    st = cx.update()
    assert st.graph_error is not None and st.graph_error.kind == "NodeActivationFailed" and st.graph_error.node == cu
    assert "asked to fail" in cx.last_error()
    proc.free(); cx.update(); cx.free()
    assert pf.counters()["drop_node"] == 1
