#!/usr/bin/env python
"""bench.py -- headline benchmark: marching-tetrahedra extraction of a 512^3 fp32 volume on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 512]

One "step" = one pass of the hot path over one synthetic 512^3 volume per GPU (BASELINE.json configs[2]:
CT-like fp32 volume, single isovalue, indexed mesh with normals).  N > 1 (torchrun, one rank per GPU):
weak scaling, each rank owns a 512-plane z-slab (+halo) of a (512*N) x 512 x 512 volume, extracts it
independently and joins an NCCL all-gather of (n_verts, n_tris) -> global vertex offsets.

Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ISOVALUE = 0.5
# dram__bytes_read.sum + dram__bytes_write.sum of the seven kernels of one extraction, from profiles/ (ncu --set full);
TRAFFIC_BYTES = 1037.3e6     # profiles/r1q_ncu_full_summary.txt, dram read + write summed over the seven kernels (algorithmic: 725.9 MB)
METRIC = "Gvoxels/s, 512^3 fp32 marching-tetrahedra extraction (indexed mesh + normals)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.n_loaded = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        "end of the loaded window: rows that arrive later were sampled on an idle GPU and are not used"
        self.n_loaded = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        rows = self.rows[:self.n_loaded] if self.n_loaded else self.rows
        for r in rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def _oracle_slab(args):
    """Worker: numpy oracle (restatement of the reference's algorithm) on one z-slab."""
    sub, value = args
    from oracle import mt3d
    r = mt3d.extract(sub, value, np.float32)
    mt3d.normals(sub, value, r["keys"], np.float32)
    return len(r["keys"]), len(r["tris"])


def cpu_field(n_planes, n, seed=0):
    """Host sample of the CT-like field family: n_planes x n x n fp32 (same generator family, numpy)."""
    rng = np.random.default_rng(seed)
    cen = rng.uniform(0.15, 0.85, size=(48, 3))
    sig = rng.uniform(0.03, 0.12, size=(48, 3))
    amp = rng.uniform(0.5, 1.0, size=48)
    x = (np.arange(n_planes, dtype=np.float32) / max(n_planes - 1, 1) * 0.5 + 0.25).reshape(-1, 1, 1)
    y = (np.arange(n, dtype=np.float32) / (n - 1)).reshape(1, -1, 1)
    z = (np.arange(n, dtype=np.float32) / (n - 1)).reshape(1, 1, -1)
    f = np.zeros((n_planes, n, n), dtype=np.float32)
    for q in range(48):
        f += np.float32(amp[q]) * np.exp(-0.5 * (((x - cen[q, 0]) / sig[q, 0]) ** 2 + ((y - cen[q, 1]) / sig[q, 1]) ** 2
                                                 + ((z - cen[q, 2]) / sig[q, 2]) ** 2)).astype(np.float32)
    f += (0.02 * rng.standard_normal(f.shape)).astype(np.float32) / 3.0
    return f


def time_oracle(field, value, cores):
    """Oracle port over `cores` processes (independent z-slabs with a one-plane overlap)."""
    import multiprocessing as mp
    n0 = field.shape[0]
    bounds = [round(r * (n0 - 1) / cores) for r in range(cores + 1)]
    jobs = [(np.ascontiguousarray(field[bounds[r]:bounds[r + 1] + 1]), value) for r in range(cores)
            if bounds[r + 1] > bounds[r]]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_oracle_slab(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_oracle_slab, jobs)
    dt = time.perf_counter() - t0
    return dt, sum(r[1] for r in res)


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself cannot
    travel to the GPU box and would need hours at this size) on all host cores,
    on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, min(os.cpu_count() or 1, 32))
    n = 192
    planes = max(cores * 6 + 1, 49)
    field = cpu_field(planes, n)
    vox = field.size
    times = []
    tris = 0
    for s in range(args.warmup + args.steps):
        dt, tris = time_oracle(field, ISOVALUE, cores)
        if s >= args.warmup:
            times.append(dt)
    tot = sum(times)
    val = vox * len(times) / tot / 1e9
    sample = "%dx%dx%d fp32 sub-volume of the CT-like field per step, numpy oracle port, %d processes" % (planes, n, n, cores)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gvoxels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / len(times) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "512^3 fp32 CT-like volume, isovalue 0.5 (BASELINE configs[2]); bounded sample: " + sample},
            "mtris_per_s": tris * len(times) / tot / 1e6,
            "cpu_baseline": {"value": val, "unit": "Gvoxels/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from contourist_b200 import engine as E
    from contourist_b200 import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG=VERSION makes NCCL print its version on rank 0's stdout, in front of the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # whatever NCCL logs: not on the JSON line's stream
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    # this rank's slab of the (n*world) x n x n volume, with halo planes (1 below, 2 above)
    n_total = n * world
    a, b = rank * n, (rank + 1) * n
    lo, hi = max(a - 1, 0), min(b + 2, n_total)
    field = synthetic.ct_like(n, lo, hi, device=dev, n_total=n_total)
    shape = (hi - lo, n, n)
    eng = E.Engine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_timing(True)
    flags = E.WANT_NORMALS if not os.environ.get("CTR_BENCH_NO_NORMALS") else 0   # (diagnostic switch; the metric needs normals)
    # the collective of the path: all-gather of (n_verts, n_tris) -> exclusive scan = global vertex / triangle offsets
    # ctr_mt3d_finish waits for the extraction only (an event behind its last kernel), not for the all-gather queued
    # after it, so the host is already enqueueing step k+1 while the collective of step k-1 runs.  The pinned staging
    # buffers therefore rotate: slot s is reused 4 steps later, two finishes after its copy was consumed.
    counts_dev = torch.zeros(2, dtype=torch.int64, device=dev)
    counts_pin = [torch.zeros(2, dtype=torch.int64, pin_memory=True) for _ in range(4)]
    gathered = torch.zeros(2 * world, dtype=torch.int64, device=dev)
    n_gather = [0]
    # CTR_BENCH_AG_SIDE=1 (diagnostic): issue the collective on a side stream instead; measured no better at 2 GPUs
    # (0.480 vs 0.472 ms per step)
    side = torch.cuda.Stream(device=dev) if (world > 1 and os.environ.get("CTR_BENCH_AG_SIDE")) else None

    prev = [None]

    def gather(c):
        pin = counts_pin[n_gather[0] & 3]
        n_gather[0] += 1
        pin[0] = int(c.n_verts)
        pin[1] = int(c.n_tris)
        if side is None:
            counts_dev.copy_(pin, non_blocking=True)
            dist.all_gather_into_tensor(gathered, counts_dev)
        else:
            with torch.cuda.stream(side):
                counts_dev.copy_(pin, non_blocking=True)
                dist.all_gather_into_tensor(gathered, counts_dev)

    def step():
        if world == 1:
            return eng.mt3d_run(field.data_ptr(), ISOVALUE, shape=shape, dtype=np.float32, flags=flags,
                                i_lo=a - lo, i_hi=b - lo, plane_offset=lo)
        # N > 1: the extraction is queued, the all-gather of the PREVIOUS step's counts is issued while the GPU works
        # (its host-side cost is ~0.1 ms, a fifth of a step), then the host waits; flush() issues the last one
        eng.mt3d_enqueue(field.data_ptr(), ISOVALUE, shape=shape, dtype=np.float32, flags=flags,
                         i_lo=a - lo, i_hi=b - lo, plane_offset=lo)
        if prev[0] is not None:
            gather(prev[0])
        c = eng.mt3d_finish()
        prev[0] = c
        return c

    def flush():
        if world > 1 and prev[0] is not None:
            gather(prev[0])
            prev[0] = None
        if side is not None:
            stream.wait_stream(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        c = step()
    barrier()
    # per-stage CUDA-event times come from a separate instrumented pass (the event records and their read-back sit
    # between the kernels and after every run); the timed region below runs with the instrumentation off
    stage_acc = np.zeros(8)
    n_inst = max(3, min(args.steps, 10))
    for _ in range(n_inst):
        c = step()
        stage_acc += np.array(eng.stage_times(8)) / n_inst
    eng.set_timing(False)
    c = step()
    flush()
    barrier()
    # nvidia-smi samples every 100 ms and K steps may last a few ms: the same steps run untimed for ~0.4 s right before
    # the timed region (same count on every rank: each step holds a collective), so that the clock samples are taken
    # under this load and the timed steps start on a GPU already at its sustained state
    PRELOAD = 800
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(PRELOAD):
        c = step()
    flush()
    l0 = eng.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c = step()
    flush()                                           # one all-gather per step inside the timed region
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "%d untimed steps of the same loop directly before the timed region, and the timed region" % PRELOAD
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([c.n_tris, c.n_verts], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step = float(tmax.item()) / args.steps
    vox_all = float(n) ** 3 * world
    value = vox_all / (ms_step * 1e-3) / 1e9
    n_tris_all, n_verts_all = int(tot[0].item()), int(tot[1].item())
    if world > 1:                                     # the offsets the last timed step gathered are the real ones
        g = gathered.view(world, 2).sum(0)
        assert int(g[0]) == n_verts_all and int(g[1]) == n_tris_all, (g, n_verts_all, n_tris_all)

    # ---- end-to-end through the C ABI with HOST buffers (H2D of the field + D2H of the mesh in the timed region)
    host_field = torch.empty(field.shape, dtype=torch.float32, pin_memory=True)
    host_field.copy_(field)
    hf = host_field.numpy()
    e2e_times = []
    outs = None
    for s in range(0 if args.no_e2e else 2 + max(1, min(args.steps, 5))):
        barrier()
        t1 = time.perf_counter()
        if world == 1:
            # the host-array call of the engine: slabs uploaded / extracted / downloaded in a pipeline
            tot_e, outs = eng.mt3d_extract_host(hf, ISOVALUE, flags=flags, nslabs=args.e2e_slabs)
            ce = argparse.Namespace(n_verts=tot_e["n_verts"], n_tris=tot_e["n_tris"])
        else:
            ce = eng.mt3d_run(hf, ISOVALUE, flags=flags, i_lo=a - lo, i_hi=b - lo, plane_offset=lo)
            outs = eng.mt3d_fetch(pinned=True)
        if world > 1:
            counts_dev.copy_(torch.tensor([ce.n_verts, ce.n_tris], dtype=torch.int64))
            dist.all_gather_into_tensor(gathered, counts_dev)
        barrier()
        if s >= 2:
            e2e_times.append(time.perf_counter() - t1)
    e2e_t = torch.tensor([float(np.mean(e2e_times)) if e2e_times else float('nan')], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = vox_all / float(e2e_t.item()) / 1e9
    h2d = hf.nbytes
    d2h = (outs["verts"].nbytes + outs["normals"].nbytes + outs["tris"].nbytes) if outs else 0

    # ---- the mesh passes that follow an extraction (not part of the metric): reference orientation, seeded selection
    post = None
    if world == 1 and not args.no_e2e:
        post = {}
        idx = int(torch.argmax(field))
        pi, pj, pk = idx // (shape[1] * shape[2]), (idx // shape[2]) % shape[1], idx % shape[2]
        row = field[pi, pj].cpu().numpy()
        while pk + 1 < shape[2] and row[pk + 1] >= ISOVALUE:
            pk += 1                                   # last sample of the row still above the isovalue: a border voxel
        for name in ("orient_reference", "select_seeded"):
            ts = []
            for _ in range(3):
                run_c = eng.mt3d_run(field.data_ptr(), ISOVALUE, shape=shape, dtype=np.float32, flags=flags)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                res = eng.mt3d_orient_reference() if name == "orient_reference" else \
                    eng.mt3d_select_seeded(np.array([[min(pi, shape[0] - 2), min(pj, shape[1] - 2), min(pk, shape[2] - 2)]], np.int32))
                ts.append((time.perf_counter() - t1) * 1e3)
            post[name] = {"ms": min(ts), "result": list(res), "n_tris_in": int(run_c.n_tris)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    st = stage_acc                                    # ms: 0 H2D, 1 classify, 2 count+scan, 3 verts, 4 tris
    field_bytes = float(np.prod(shape)) * 4
    # SURVEY 8(d) algorithmic bytes: 4 B per voxel read (stage 1), 24 B per vertex (stage 3), 12 B per triangle (stage 4);
    # the scan (stage 2) moves no algorithmic bytes -- it is pure overhead and is charged to the whole-step figure.
    stages = {"k_bitplane_tma (stage 1: field -> low/near bitplanes, the only full-field pass)": (st[1], field_bytes),
              "k_emit_verts (stage 3: positions + normals)": (st[3], c.n_verts * 24.0),
              "k_emit_tris (stage 4: indexed triangles)": (st[4], c.n_tris * 12.0)}
    per_stage = {k.split(" ")[0]: {"ms": v[0], "algorithmic_bytes": v[1],
                                   "gbs": (v[1] / (v[0] * 1e-3) / 1e9) if v[0] > 0 else None,
                                   "frac": (v[1] / (v[0] * 1e-3) / 1e9 / peak) if v[0] > 0 else None}
                 for k, v in stages.items()}
    per_stage["k_count_a+k_count_b+k_tile_scan3+k_scan"] = {"ms": st[2], "algorithmic_bytes": 0.0, "gbs": 0.0, "frac": 0.0}
    kern_ms = float(st[1] + st[2] + st[3] + st[4])    # the four kernels of one extraction, from CUDA events on their stream
    dom = max(per_stage, key=lambda k: per_stage[k]["ms"])
    alg_total = float(n) ** 3 * 4 + c.n_verts * 24 + c.n_tris * 12        # field + V*(3p+3p) + T*12
    pipe_gbs = alg_total / (ms_step * 1e-3) / 1e9
    # CPU baseline: oracle port, 1 core, bounded sample of the same field family
    sub = cpu_field(33, 160)
    t_cpu, _ = time_oracle(sub, ISOVALUE, 1) if not args.no_e2e else (float('nan'), 0)
    cpu_val = sub.size / t_cpu / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3) + n_inst + 1 + PRELOAD, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: %d^3 fp32 CT-like volume per GPU (48 Gaussian blobs + smoothed noise), "
                               "isovalue 0.5, indexed mesh + gradient normals, fp32 geometry" % n,
                   "volume": [n_total, n, n],
                   "stacking": "one CT-like block (own 48 blobs) per GPU along the first axis, the neighbours' Gaussian tails "
                               "summed in: one continuous volume, the same amount of surface in every slab",
                   "sharding": "z-slabs, 1 plane halo below / 2 above, NCCL all-gather of counts",
                   "l2": "inputs (%.0f MB field) larger than the 126 MB L2; no explicit flush" % (field_bytes / 1e6)},
        "mtris_per_s": n_tris_all / (ms_step * 1e-3) / 1e6, "n_tris": n_tris_all, "n_verts": n_verts_all,
        "stage_ms": {"bitplane": st[1], "count_scan": st[2], "emit_verts": st[3], "emit_tris": st[4],
                     "wall_ms_per_step": wall / args.steps * 1e3},
        # one extraction = one launch each of four kernels; SURVEY 8(d) defines the algorithmic bytes per extraction, so the
        # roofline is taken over the four launches together (the scan kernel moves no algorithmic bytes of its own)
        "roofline": {"bound": "hbm", "kernel": "mt3d extraction = k_bitplane_tma + k_count_a + k_count_b + k_tile_scan3 + k_scan + k_emit_verts + k_emit_tris "
                                                "(largest share: %s, %.0f%% of the kernel time)" % (dom, 100.0 * per_stage[dom]["ms"] / kern_ms),
                     "achieved": alg_total / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg_total / (kern_ms * 1e-3) / 1e9 / peak, "traffic": TRAFFIC_BYTES, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_total, "kernel_ms": kern_ms, "per_stage": per_stage},
        "pipeline_roofline": {"algorithmic_bytes": alg_total, "achieved": pipe_gbs, "frac": pipe_gbs / peak,
                              "note": "whole step: (field + V*24 + T*12) / device time of the step"},
        "cpu_baseline": {"value": cpu_val, "unit": "Gvoxels/s", "cores": 1, "kind": "port",
                         "sample": "33x160x160 fp32 sub-volume of the CT-like field, numpy oracle port (extract + normals)"},
        "e2e": {"value": e2e_val, "unit": "Gvoxels/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "note": "Engine.mt3d_extract_host (ctr_mt3d_run + ctr_mt3d_fetch per z-slab on two contexts, page-locked host "
                        "buffers): H2D of the field and D2H of vertices, normals and triangles inside the timed region"},
        "gpu_launches": int(launches), "clocks": clocks,
        "post_passes": post,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--e2e-slabs", type=int, default=4, help="slabs of the pipelined host-array call (1 = run + fetch)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg and the CPU baseline (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
