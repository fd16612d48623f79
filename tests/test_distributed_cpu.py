"""N > 1 host logic on CPU (gloo, world_size 2): voice sharding + the fixed-order bus exchange. Each rank runs the CPU
oracle on its shard, the per-rank buses are all-gathered and tree-summed in rank order; the result must equal the
tree-of-trees bus, and the flat single-rank bus whenever the shards are equal powers of two."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, V, port, out_dir):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    import pyoracle
    from conftest import synth
    from firewheel_b200 import PanNode, VolumeNode
    from sharding import tree_sum, voice_range
    from helpers import chain, run_planar
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = pyoracle.load()
    T = 512
    rng = np.random.default_rng(11)
    pct = (25 + 75 * rng.random(V)).astype(np.float32); pan = rng.uniform(-1, 1, V).astype(np.float32)
    x = synth((V, 2, T), 5)
    lo, hi = voice_range(V, rank, world)

    def bus_of(a, b):
        def setup(cx, ids):
            cx.graph.set_percent_volume(ids[0], pct[a:b]); cx.graph.set_pan(ids[1], pan[a:b])
        cx, proc, _ = chain(lib, 2, [(lambda: VolumeNode(100.0), 2, 2), (lambda: PanNode(0.0), 2, 2)], voices=b - a, master_bus=True, setup=setup)
        y, _ = run_planar(proc, np.ascontiguousarray(x[a:b]), 2, True)
        proc.free(); cx.update(); cx.free()
        return y

    mine = torch.from_numpy(bus_of(lo, hi))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    bus = tree_sum([g.numpy() for g in gathered])
    ref_tot = tree_sum([bus_of(*voice_range(V, r, world)) for r in range(world)])
    ok = np.array_equal(bus.view(np.uint32), ref_tot.view(np.uint32))
    sizes = [voice_range(V, r, world)[1] - voice_range(V, r, world)[0] for r in range(world)]
    if len(set(sizes)) == 1 and sizes[0] & (sizes[0] - 1) == 0:
        ok = ok and np.array_equal(bus.view(np.uint32), bus_of(0, V).view(np.uint32))
    Path(out_dir, f"rank{rank}.ok").write_text("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("V", [8, 7])
def test_sharded_bus_over_gloo(tmp_path, V):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 500) + V
    mp.spawn(_worker, args=(2, V, port, str(tmp_path)), nprocs=2, join=True)
    assert [Path(tmp_path, f"rank{r}.ok").read_text() for r in range(2)] == ["1", "1"]


def test_voice_range_partitions():
    from sharding import voice_range
    for V in (1, 7, 8, 1024, 65536):
        for w in (1, 2, 3, 8):
            spans = [voice_range(V, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == V
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
