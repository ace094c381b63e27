"""z-slab sharding of a volume across ranks: one process per GPU, no data-path collective.

Each rank extracts voxel layers / owner planes [a, b) of the global volume from a slab that carries
1 halo plane below (gradient normals) and 2 above (ids of the next shard's first owner plane), so every rank
can write GLOBAL vertex ids with nothing but an all-gather of (n_verts, n_tris) -> exclusive offsets.
torch.distributed is used for exactly that collective (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def slab_bounds(n0, world):
    """Owner-plane ranges [a_r, b_r) per rank: contiguous, covering [0, n0), as even as possible."""
    return [(int(round(r * n0 / float(world))), int(round((r + 1) * n0 / float(world)))) for r in range(world)]


def slab_with_halo(a, b, n0):
    """Planes [lo, hi) a rank must hold, and the engine arguments (i_lo, i_hi, plane_offset) for them."""
    lo = max(a - 1, 0)
    hi = min(b + 2, n0)
    return lo, hi, dict(i_lo=a - lo, i_hi=b - lo, plane_offset=lo)


def exclusive_offsets(counts):
    """counts [world, k] -> offsets [world, k] (exclusive prefix sum over ranks) and totals [k]."""
    counts = np.asarray(counts, dtype=np.int64)
    off = np.zeros_like(counts)
    off[1:] = np.cumsum(counts[:-1], axis=0)
    return off, counts.sum(axis=0)


def allgather_counts(n_verts, n_tris, device=None):
    """All-gather of this rank's (n_verts, n_tris); returns (offsets_of_this_rank [2], totals [2], all counts)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        c = np.array([[n_verts, n_tris]], dtype=np.int64)
        off, tot = exclusive_offsets(c)
        return off[0], tot, c
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([int(n_verts), int(n_tris)], dtype=torch.int64, device=device)
    out = torch.zeros(2 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine)
    counts = out.view(world, 2).cpu().numpy()
    off, tot = exclusive_offsets(counts)
    return off[rank], tot, counts


def extract_sharded(engine, slab, a, b, n0, value, origin=(0, 0, 0), delta=(1, 1, 1), flags=0, device=None,
                    shape=None, dtype=None):
    """Run one rank's slab through the engine and globalise its triangle ids.
    slab: numpy array or device pointer holding planes [lo, hi) from slab_with_halo(a, b, n0).
    Returns (counts, fetch dict with int64 global triangle ids, vertex offset of this rank, totals)."""
    lo, hi, kw = slab_with_halo(a, b, n0)
    c = engine.mt3d_run(slab, value, origin=origin, delta=delta, flags=flags, shape=shape, dtype=dtype, **kw)
    off, tot, _ = allgather_counts(c.n_verts, c.n_tris, device=device)
    out = engine.mt3d_fetch()
    if out["tris"] is not None:
        out["tris"] = out["tris"].astype(np.int64) + int(off[0])
    return c, out, off, tot


def init_native_comm(engine):
    """Give `engine` an NCCL communicator inside the C library (ctr_comm_init) spanning the ranks of the initialised
    torch.distributed group: rank 0 makes the NCCL id, torch.distributed carries its 128 bytes to the others (the only
    thing it is used for here).  Afterwards engine.allgather_offsets() / engine.gather_mesh() are the collectives of the
    path, issued by the library on the engine's stream."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.comm_init(box[0], rank, world)
    return rank, world


def extract_and_gather(engine, slab, a, b, n0, value, origin=(0, 0, 0), delta=(1, 1, 1), flags=0, root=0, shape=None, dtype=None):
    """One rank's part of a sharded extraction with the mesh gathered on `root` (north_star: "NCCL ... only for an
    allgather of per-rank triangle counts and vertex offsets, plus an optional gather of the mesh to rank 0").
    slab holds planes slab_with_halo(a, b, n0); engine must have a native communicator (init_native_comm).
    Returns (counts of this rank, all counts [world, 2], offsets of this rank, totals, gathered mesh or None)."""
    lo, hi, kw = slab_with_halo(a, b, n0)
    c = engine.mt3d_run(slab, value, origin=origin, delta=delta, flags=flags, shape=shape, dtype=dtype, **kw)
    counts, off, tot = engine.allgather_offsets()
    mesh = engine.gather_mesh(root)
    return c, counts, off, tot, mesh
