"""Worker of tests/test_gpu_multirank.py: launched with torchrun, one rank per GPU (NCCL)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from contourist_b200 import engine as E
    from contourist_b200 import sharding
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = E.Engine(local)
    sharding.init_native_comm(eng)
    # every rank builds the same volume (seeded) and cuts out its own slab
    n0, n1, n2 = 67, 45, 96
    x, y, z = np.meshgrid(np.arange(n0), np.arange(n1), np.arange(n2), indexing="ij")
    rng = np.random.default_rng(3)
    f = np.zeros((n0, n1, n2))
    for _ in range(9):
        c = rng.uniform(0, 1, 3) * (n0, n1, n2)
        r = rng.uniform(6, 15)
        f += np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (r * r))
    f = (f + 0.02 * rng.standard_normal(f.shape)).astype(np.float32)
    value = 0.45
    flags = E.WANT_NORMALS | E.GEOM_F64
    a, b = sharding.slab_bounds(n0, world)[rank]
    lo, hi, _ = sharding.slab_with_halo(a, b, n0)
    c, counts, off, tot, mesh = sharding.extract_and_gather(eng, np.ascontiguousarray(f[lo:hi]), a, b, n0, value, flags=flags,
                                                            origin=(-1.0, 0.5, 2.0), delta=(0.5, 0.25, 2.0))
    # the native all-gather agrees with torch.distributed's
    off2, tot2, counts2 = sharding.allgather_counts(c.n_verts, c.n_tris, device=dev)
    assert np.array_equal(counts, counts2) and np.array_equal(off, off2) and np.array_equal(tot, tot2)
    assert counts[rank, 0] == c.n_verts and counts[rank, 1] == c.n_tris
    if rank == 0:
        single = E.Engine(local)
        cs = single.mt3d_run(f, value, flags=flags, origin=(-1.0, 0.5, 2.0), delta=(0.5, 0.25, 2.0))
        ref = single.mt3d_fetch()
        assert (int(tot[0]), int(tot[1])) == (cs.n_verts, cs.n_tris)
        # vertex ids are ordered by owner word (plane-major) and triangles by voxel: the slabs concatenated in rank
        # order ARE the single-GPU arrays
        assert np.array_equal(mesh["verts"], ref["verts"])
        assert np.array_equal(mesh["normals"], ref["normals"])
        assert np.array_equal(mesh["tris"], ref["tris"])
        print("MULTIRANK OK: %d ranks, %d vertices, %d triangles gathered == single-GPU mesh" % (world, cs.n_verts, cs.n_tris))
    else:
        assert mesh is None
    # a second round on a different isovalue reuses the communicator and the buffers
    c, counts, off, tot, mesh = sharding.extract_and_gather(eng, np.ascontiguousarray(f[lo:hi]), a, b, n0, 0.3, flags=E.WANT_NORMALS)
    if rank == 0:
        cs = single.mt3d_run(f, 0.3, flags=E.WANT_NORMALS)
        ref = single.mt3d_fetch()
        assert np.array_equal(mesh["verts"], ref["verts"]) and np.array_equal(mesh["tris"], ref["tris"])
        print("MULTIRANK OK (second round)")
    eng.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
