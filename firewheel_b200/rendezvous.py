"""Torch-free bootstrap of the master-bus communicator for one-node launches (torchrun or any launcher that sets
RANK / WORLD_SIZE / MASTER_PORT). Only the 128-byte NCCL unique id has to travel out of band: rank 0 publishes it in a
file under /tmp, the other ranks poll for it. Everything after that (barriers, max-over-ranks, result cross-checks) goes
through `FirewheelProcessor.comm_allgather`, i.e. NCCL itself."""
import ctypes
import os
import time
from pathlib import Path

_seq = 0
_T0 = time.time()


def _path(tag):
    # all ranks of one launch share the launcher as parent and MASTER_PORT; `tag` counts communicators within the launch
    return Path(os.environ.get("FW_RDV_DIR", "/tmp")) / f"fw_b200_rdv_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}_{tag}"


def exchange_unique_id(lib, rank, timeout_s=120.0):
    """Returns the communicator id created on rank 0 (bytes, 128)."""
    global _seq
    _seq += 1
    p = _path(_seq)
    if rank == 0:
        buf = (ctypes.c_uint8 * 128)()
        if lib.comm_unique_id(buf) != 0:
            raise RuntimeError("ncclGetUniqueId failed: " + (lib.last_device_error() or b"").decode())
        tmp = p.with_suffix(f".tmp{os.getpid()}")
        tmp.write_bytes(bytes(buf))
        os.replace(tmp, p)  # atomic publish
        return bytes(buf)
    t_end = time.time() + timeout_s
    while time.time() < t_end:
        try:
            if p.stat().st_mtime >= _T0 - 300.0:  # not a leftover of an older launch that happened to share pid and port
                data = p.read_bytes()
                if len(data) == 128:
                    return data
        except FileNotFoundError:
            pass
        time.sleep(0.01)
    raise TimeoutError(f"rank {rank}: no communicator id at {p}")


def cleanup(rank):
    if rank != 0:
        return
    for k in range(1, _seq + 1):
        try:
            _path(k).unlink()
        except OSError:
            pass


def init_comm(lib, proc, rank, world):
    """Give `proc` a communicator over all ranks (no-op for world == 1)."""
    if world <= 1:
        proc.world_size = 1
        return
    uid = exchange_unique_id(lib, rank)
    if proc.comm_init(rank, world, uid) != 0:
        raise RuntimeError("ncclCommInitRank failed: " + (lib.last_device_error() or b"").decode())
