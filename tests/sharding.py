"""Voice sharding across ranks: rank r owns a contiguous voice range including all per-voice state;
the only exchange is the master bus. The multi-rank bus is the balanced tree (2i, 2i+1; unpaired carried) over the
ranks' buses, each of which is the same balanced tree over the rank's own voices ("tree of trees"); it equals the
flat tree over all voices whenever every rank owns the same power-of-two number of voices."""
import numpy as np


def voice_range(num_voices, rank, world_size):
    """Contiguous split; the first `num_voices % world_size` ranks own one voice more."""
    base, extra = divmod(num_voices, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def tree_sum(parts):
    """Balanced pairwise f32 sum with carry of an unpaired last element — the order the 2-port SumNode tree uses."""
    level = [np.asarray(p, dtype=np.float32) for p in parts]
    while len(level) > 1:
        nxt = [(level[i] + level[i + 1]).astype(np.float32) for i in range(0, len(level) - 1, 2)]
        if len(level) & 1:
            nxt.append(level[-1])
        level = nxt
    return level[0]
