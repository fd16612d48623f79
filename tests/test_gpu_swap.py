"""Schedule swaps on the temporal and generic lowerings (processor.rs:167-206 + Q11: the first block after a swap reads zero
inputs; node state survives the swap like the reference's processors do). The pointwise-chain case lives in test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import synth
from firewheel_b200 import BiquadNode, DelayNode, HardClipNode, SamplerNode, SumNode, VolumeNode, design_rbj
from helpers import assert_bit_exact, chain, f32, run_planar

pytestmark = pytest.mark.gpu


def _swap_scenario(lib, V, F, kind):
    """kind: 'temporal' (biquad -> delay chain), 'generic' (dry + gained path into a SumNode) — a HardClipNode is spliced in
    before graph_out between calls and removed again."""
    from firewheel_b200 import AudioGraphConfig, FirewheelGraphCtx
    x = synth((V, 2, 4 * F), 77)
    if kind == "temporal":
        co = np.stack([[design_rbj(lib, 0, 700.0 + 50 * v, 0.8, 0.0, 48000), design_rbj(lib, 4, 2500.0, 1.1, 4.0, 48000)] for v in range(V)]).astype(f32)
        cx, proc, ids = chain(lib, 2, [(lambda: BiquadNode(2), 2, 2), (lambda: DelayNode(160), 2, 2)], voices=V, max_block=F,
                              setup=lambda cx, ids: cx.graph.set_biquad_coeffs(ids[0], co))
        last = ids[-1]
        g = cx.graph
    else:
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=V))
        g = cx.graph
        vol, mix = g.add_node(2, 2, VolumeNode(60.0)), g.add_node(4, 2, SumNode())
        for c in range(2):
            g.connect(g.graph_in_node(), c, vol, c, False); g.connect(g.graph_in_node(), c, mix, c, False)
            g.connect(vol, c, mix, 2 + c, False); g.connect(mix, c, g.graph_out_node(), c, False)
        proc = cx.activate(48000, 2, 2, F)
        assert cx.update().graph_error is None
        last = mix
    outs = [run_planar(proc, x, 2)]
    clip = g.add_node(2, 2, HardClipNode(-9.0))
    for c in range(2):
        assert g.disconnect(last, c, g.graph_out_node(), c)
        g.connect(last, c, clip, c, False); g.connect(clip, c, g.graph_out_node(), c, False)
    assert cx.update().graph_error is None
    outs += [run_planar(proc, x, 2), run_planar(proc, x, 2)]
    g.remove_node(clip)
    for c in range(2):
        g.connect(last, c, g.graph_out_node(), c, False)
    assert cx.update().graph_error is None
    outs.append(run_planar(proc, x, 2))
    proc.free(); cx.update(); cx.free()
    return outs


@pytest.mark.parametrize("kind,F", [("temporal", 128), ("temporal", 100), ("generic", 128)])
def test_schedule_swaps_on_temporal_and_generic_paths(gpu, oracle, kind, F):
    a, b = _swap_scenario(gpu, 5, F, kind), _swap_scenario(oracle, 5, F, kind)
    for i, ((yg, mg), (yo, mo)) in enumerate(zip(a, b)):
        assert_bit_exact(yg, yo, f"{kind} call {i}")
        assert mg == mo
