"""Run under torchrun (one rank per GPU): voices shard by rank, the master bus crosses ranks through the product's
NCCL all-gather + fixed-order tree. Every rank checks its result bit-for-bit against the CPU oracle's tree of trees."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    import firewheel_b200 as fw
    import pyoracle
    from conftest import synth
    from firewheel_b200 import PanNode, VolumeNode
    from firewheel_b200.sharding import tree_sum, voice_range
    from helpers import chain, run_planar
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
    ids = [bytes(128)]
    if rank == 0:
        import ctypes
        buf = (ctypes.c_uint8 * 128)()
        assert gpu.comm_unique_id(buf) == 0, gpu.last_device_error()
        ids = [bytes(buf)]
    dist.broadcast_object_list(ids, src=0)
    assert proc.comm_init(rank, world, ids[0]) == 0, gpu.last_device_error()
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
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    proc.free(); cx.update(); cx.free()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("multigpu parity", "OK" if int(flag.item()) == 1 else "MISMATCH", f"world={world} voices={V}")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
