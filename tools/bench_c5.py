#!/usr/bin/env python
"""BASELINE configs[4]: 2048^3 fp32 'turbulence', isovalue 0, z-slab sharded over N B200
(strong scaling: the volume is fixed, every rank takes 1/N of the owner planes).  bench.py stays the headline line
(configs[2]); this prints one JSON line for C5.

    python tools/bench_c5.py [--size 2048] [--chunk 256] [--steps 3] [--warmup 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_c5.py --size 2048

A rank walks its slab in chunks of `--chunk` owner planes (i_lo / i_hi / plane_offset of the C ABI; 1 halo plane
below, 2 above, generated on the device with the chunk, not timed): 32 GiB of field never has to be resident at
once and the engine's work buffers stay at chunk size.  Per chunk: W warm-up runs, K timed runs (CUDA events on the
engine's stream), mean taken.  A rank's time = sum over its chunks; the job's time = max over ranks + the all-gather
of (n_verts, n_tris) that yields the global vertex / triangle offsets (the only collective on the path).
The totals (n_verts, n_tris) are independent of N and of --chunk (every vertex has exactly one owner plane): compare
the lines of two runs to check the sharding.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=2048, help="samples per axis (not --n: torchrun's parser claims that prefix)")
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--value", type=float, default=0.0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from contourist_b200 import engine as E
    from contourist_b200 import sharding, synthetic
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # whatever NCCL logs: not on the JSON line's stream
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    a_r, b_r = sharding.slab_bounds(n, world)[rank]
    eng = E.Engine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    flags = E.WANT_NORMALS
    tot_ms, nv, nt, ncells, nchunks = 0.0, 0, 0, 0, 0
    for a in range(a_r, b_r, args.chunk):
        b = min(a + args.chunk, b_r)
        lo, hi, kw = sharding.slab_with_halo(a, b, n)
        field = synthetic.turbulence(n, lo, hi, device=dev, n_total=n)
        shape = (hi - lo, n, n)

        def run():
            return eng.mt3d_run(field.data_ptr(), args.value, shape=shape, dtype=np.float32, flags=flags, **kw)
        for _ in range(max(args.warmup, 1)):
            c = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            c = run()
        e1.record(stream)
        torch.cuda.synchronize()
        tot_ms += e0.elapsed_time(e1) / args.steps
        nv += int(c.n_verts)
        nt += int(c.n_tris)
        ncells += int(c.n_active_cells)
        nchunks += 1
        del field
    # the collective: global offsets of this rank's vertices / triangles
    ag_ms = 0.0
    mine = torch.tensor([nv, nt, ncells], dtype=torch.int64, device=dev)
    if world > 1:
        out = torch.zeros(3 * world, dtype=torch.int64, device=dev)
        for _ in range(3):
            dist.all_gather_into_tensor(out, mine)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_gather_into_tensor(out, mine)
        e1.record()
        torch.cuda.synchronize()
        ag_ms = e0.elapsed_time(e1) / 10
        counts = out.view(world, 3).cpu().numpy()
        t = torch.tensor([tot_ms, ag_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        slowest, ag_ms = float(t[0]), float(t[1])
        per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(per_rank, torch.tensor([tot_ms], dtype=torch.float64, device=dev))
        per_rank = [float(x) for x in per_rank.cpu()]
    else:
        counts = mine.view(1, 3).cpu().numpy()
        slowest, per_rank = tot_ms, [tot_ms]
    if rank == 0:
        off, tot = sharding.exclusive_offsets(counts[:, :2])
        V, T = int(tot[0]), int(tot[1])
        ms = slowest + ag_ms
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        pk = float(json.load(open(pk_path))["hbm_gbs"]) if os.path.exists(pk_path) else 6650.0
        alg = float(n) ** 3 * 4 + V * 24.0 + T * 12.0            # field once + fp32 positions and normals + int32 triangles
        print(json.dumps({
            "path": "C5: %d^3 fp32 turbulence, isovalue %g, z-slabs over %d GPU(s) (BASELINE configs[4])" % (n, args.value, world),
            "metric": "Gvoxels/s", "value": float(n) ** 3 / ms / 1e6, "mtris_per_s": T / ms / 1e3, "ms_total": ms,
            "ms_slowest_rank": slowest, "ms_per_rank": per_rank, "allgather_ms": ag_ms, "n_gpus": world, "scaling": "strong",
            "chunk_planes": args.chunk, "chunks_per_rank": nchunks, "steps": args.steps, "warmup": max(args.warmup, 1),
            "n_verts": V, "n_tris": T, "n_active_cells": int(counts[:, 2].sum()),
            "vertex_offsets": [int(x) for x in off[:, 0]], "dtype": "f32", "data": "synthetic (sum of 64 sines, rng(1))",
            "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6 / world, "peak": pk, "unit": "GB/s per GPU",
                         "frac": alg / ms / 1e6 / world / pk, "algorithmic_bytes": alg}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
