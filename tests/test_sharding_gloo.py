"""CPU, world_size 2 over gloo: the multi-GPU host logic (slab partition, halo arithmetic, the count all-gather
and the global offsets).  The extraction itself is replaced by the numpy oracle here -- the CUDA path of the same
decomposition is covered by tests/test_gpu_mt3d.py::test_slab_sharding_matches_single_run."""
import os
import socket
import sys

import numpy as np
import pytest

from contourist_b200 import sharding


def test_slab_bounds_cover_and_halo():
    for n0 in (5, 64, 513):
        for world in (1, 2, 3, 8):
            b = sharding.slab_bounds(n0, world)
            assert b[0][0] == 0 and b[-1][1] == n0
            assert all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
            for a, e in b:
                lo, hi, kw = sharding.slab_with_halo(a, e, n0)
                assert lo == max(a - 1, 0) and hi == min(e + 2, n0)
                assert kw["plane_offset"] == lo and kw["i_lo"] == a - lo and kw["i_hi"] == e - lo


def test_exclusive_offsets():
    off, tot = sharding.exclusive_offsets([[3, 5], [2, 1], [7, 0]])
    assert off.tolist() == [[0, 0], [3, 5], [5, 6]] and tot.tolist() == [12, 6]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import mt3d
    g = np.linspace(-1, 1, 24)
    X, Y, Z = np.meshgrid(g, g[:17], g[:19], indexing="ij")
    f = np.sin(4 * X) * np.cos(3 * Y) + Z * Z
    n0 = f.shape[0]
    a, b = sharding.slab_bounds(n0, world)[rank]
    # this rank's share of the oracle mesh: vertices owned by planes [a, b), triangles of voxel layers [a, b)
    r = mt3d.extract(f, 0.2)
    plane = (r["keys"] >> np.uint64(3)).astype(np.int64) // (f.shape[1] * f.shape[2])
    nv = int(((plane >= a) & (plane < b)).sum())
    layer = r["tri_cell"] // ((f.shape[1] - 1) * (f.shape[2] - 1))
    nt = int(((layer >= a) & (layer < b)).sum())
    off, tot, counts = sharding.allgather_counts(nv, nt)
    q.put((rank, nv, nt, off.tolist(), tot.tolist(), counts.tolist(), len(r["keys"]), len(r["tris"])))
    dist.destroy_process_group()


def test_allgather_offsets_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, nv0, nt0, off0, tot0, c0, V, T), (r1, nv1, nt1, off1, tot1, c1, _, _) = res
    assert off0 == [0, 0] and off1 == [nv0, nt0]
    assert tot0 == tot1 == [V, T] == [nv0 + nv1, nt0 + nt1]
    assert c0 == c1 == [[nv0, nt0], [nv1, nt1]]
