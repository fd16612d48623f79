"""Shared builders for parity tests: the same code drives the CUDA product and the CPU oracle."""
import numpy as np

from firewheel_b200 import AudioGraphConfig, FirewheelGraphCtx

f32 = np.float32
SR = 48000


def chain(lib, n_ch, nodes, voices=1, master_bus=False, max_block=256, n_out=None, setup=None):
    """graph_in(n_ch) -> nodes[0] -> ... -> graph_out, port i to port i. nodes: [(node_factory, n_in, n_out)]."""
    n_out = n_out if n_out is not None else (nodes[-1][2] if nodes else n_ch)
    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=n_ch, num_graph_outputs=n_out, num_voices=voices, master_bus=master_bus))
    g = cx.graph
    prev, prev_w, ids = g.graph_in_node(), n_ch, []
    for node, ni, no in nodes:
        nid = g.add_node(ni, no, node() if callable(node) else node)
        for p in range(min(prev_w, ni)):
            g.connect(prev, p, nid, p, False)
        prev, prev_w = nid, no
        ids.append(nid)
    for p in range(min(prev_w, n_out)):
        g.connect(prev, p, g.graph_out_node(), p, False)
    if setup:
        setup(cx, ids)  # parameters set before activation start un-smoothed (volume.rs:67-75)
    proc = cx.activate(SR, n_ch, n_out, max_block)
    assert proc is not None
    st = cx.update()
    assert st.kind == "Active" and st.graph_error is None, (st, cx.last_error())
    return cx, proc, ids


def run_planar(proc, x, n_out, master_bus=False):
    V, n_in, T = x.shape
    out = np.full((n_out, T) if master_bus else (V, n_out, T), np.nan, dtype=f32)
    rc, mask = proc.process_planar(np.ascontiguousarray(x), out, n_in, n_out, T)
    assert rc == 0, (rc, proc._lib.last_device_error())
    return out, mask


def bits(a):
    return np.ascontiguousarray(a, dtype=f32).view(np.uint32)


def assert_bit_exact(got, ref, what=""):
    gb, rb = bits(got), bits(ref)
    if not np.array_equal(gb, rb):
        bad = np.argwhere(gb != rb)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {gb.size} samples differ; first at {i}: got {got[i]!r} ({gb[i]:#010x}) want {ref[i]!r} ({rb[i]:#010x})")
