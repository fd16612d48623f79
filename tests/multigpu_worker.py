"""Run under torchrun (one rank per GPU; torch itself is not imported): voices shard by rank, the master bus crosses ranks
through the product's NCCL all-gather + fixed-order tree. Every rank checks its result bit-for-bit against the CPU oracle's
tree of trees; the communicator id travels through firewheel_b200.rendezvous, verdicts through the communicator itself."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))


def main():
    import firewheel_b200 as fw
    import pyoracle
    from conftest import synth
    from firewheel_b200 import PanNode, VolumeNode
    from sharding import tree_sum, voice_range
    from helpers import chain, run_planar
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    from firewheel_b200 import rendezvous
    gpu, oracle = fw.load(), pyoracle.load()
    V = int(os.environ.get("FW_TEST_VOICES", "200"))
    T, F = 1024, 256
    rng = np.random.default_rng(3)
    pct = (25 + 75 * rng.random(V)).astype(np.float32); pan = rng.uniform(-1, 1, V).astype(np.float32)
    x = synth((V, 2, T), 9)

    def build(lib, a, b, device=0):
        from firewheel_b200 import AudioGraphConfig, FirewheelGraphCtx
        cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2, num_voices=b - a, master_bus=True, device=device))
        g = cx.graph
        vol, pn = g.add_node(2, 2, VolumeNode(100.0)), g.add_node(2, 2, PanNode(0.0))
        for c in range(2):
            g.connect(g.graph_in_node(), c, vol, c, False); g.connect(vol, c, pn, c, False); g.connect(pn, c, g.graph_out_node(), c, False)
        g.set_percent_volume(vol, pct[a:b]); g.set_pan(pn, pan[a:b])
        proc = cx.activate(48000, 2, 2, F)
        st = cx.update()
        assert st.graph_error is None, (st, cx.last_error())
        return cx, proc

    lo, hi = voice_range(V, rank, world)
    cx, proc = build(gpu, lo, hi, device=local)
    rendezvous.init_comm(gpu, proc, rank, world)
    out = np.zeros((2, T), np.float32)
    for _ in range(2):  # two calls: staging reuse
        rc, mask = proc.process_planar(np.ascontiguousarray(x[lo:hi]), out, 2, 2, T)
        assert rc == 0, (rc, gpu.last_device_error())
    parts = []
    for r in range(world):
        a, b = voice_range(V, r, world)
        ocx, oproc = build(oracle, a, b)
        y = np.zeros((2, T), np.float32)
        orc, _ = oproc.process_planar(np.ascontiguousarray(x[a:b]), y, 2, 2, T)
        oproc.process_planar(np.ascontiguousarray(x[a:b]), y, 2, 2, T)
        parts.append(y.copy())
        oproc.free(); ocx.update(); ocx.free()
    ref = tree_sum(parts)
    ok = np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    all_ok = bool(proc.comm_allgather(np.array([1 if ok else 0], np.int64)).min() == 1)
    proc.comm_allgather(np.zeros(1, np.int64))  # barrier: nobody tears its mailbox down while a peer still runs
    proc.free(); cx.update(); cx.free()
    rendezvous.cleanup(rank)
    if rank == 0:
        print("multigpu parity", "OK" if all_ok else "MISMATCH", f"world={world} voices={V}")
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
