"""Regenerates tests/golden/*.npz from the CPU oracle:  python tests/golden/make_golden.py
Each fixture holds a scenario's inputs, its extra data (tables, PCM, IR) and the oracle's outputs + silence masks per call."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import golden_scenarios as gs  # noqa: E402
import pyoracle  # noqa: E402


def main():
    lib = pyoracle.load()
    for name in gs.SCENARIOS:
        inputs, extra = gs.make_inputs(name, lib)
        outs = gs.run(lib, name, inputs, extra)
        d = {f"in{i}": x for i, x in enumerate(inputs)}
        d.update({f"extra_{k}": v for k, v in extra.items()})
        for i, (y, m) in enumerate(outs):
            d[f"out{i}"] = y; d[f"mask{i}"] = np.uint64(m)
        path = Path(__file__).resolve().parent / f"{name}.npz"
        np.savez_compressed(path, **d)
        print(name, [y.shape for y, _ in outs], f"{path.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
