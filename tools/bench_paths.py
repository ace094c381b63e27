#!/usr/bin/env python
"""Measured numbers for the other two hot-path rows: BASELINE configs[1] (2D, 16384^2 fp32,
16 levels) and configs[3] (4D, 128^3 x 64 fp32 morph field).  bench.py stays the single headline line (configs[2]);
this prints one JSON line per path with the same roofline conventions.

    python tools/bench_paths.py [--steps 10] [--warmup 3] [--n2d 16384] [--n4d 128 --nt 64]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def timed(fn, steps, warmup, stream, acc):
    import torch
    for _ in range(max(warmup, 3)):
        c = fn()
    torch.cuda.synchronize()
    acc[:] = 0                                       # stage times of the timed steps only
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        c = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return c, e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n2d", type=int, default=16384)
    ap.add_argument("--n4d", type=int, default=128)
    ap.add_argument("--nt", type=int, default=64)
    args = ap.parse_args()
    import torch
    from contourist_b200 import engine as E
    from contourist_b200 import synthetic
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    eng = E.Engine(0)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_timing(True)
    pk, src = peak()

    # ---- 2D: configs[1]
    n = args.n2d
    f = synthetic.field2d(n, device=dev)
    mn, mx = float(f.min()), float(f.max())
    levels = [(mx - mn) / 17 * i for i in range(1, 17)]          # Linear2DContour rule, multiple_2d_contour.py:102-107
    acc = np.zeros(8)

    def step2():
        c = eng.mt2d_run(f.data_ptr(), levels, shape=(n, n), dtype=np.float32)
        acc[:] += np.array(eng.stage_times(8))
        return c
    c, ms = timed(step2, args.steps, args.warmup, stream, acc)
    st = acc / args.steps
    S = int(c.n_segments)
    alg = n * n * 4.0 + S * (1 + 16 + 16)                        # field once for all levels + level tag, 2 keys, 2 fp32 points
    # CPU baseline (oracle port, 1 core, bounded sample of the same field) and end to end with host buffers
    import time
    from oracle import mt2d as o2, mp4d as o4           # the cpu_baseline leg: the oracle as the thing timed beside, never the product
    c0 = n // 2 - 512
    sub = f[c0:c0 + 1024, c0:c0 + 1024].cpu().numpy()
    t0 = time.perf_counter()
    for z in levels:
        o2.extract_level(sub, z, np.float32)
    cpu2 = sub.size / (time.perf_counter() - t0) / 1e9
    hf2 = eng.pinned_empty("bench_field2d", (n, n), np.float32)
    hf2[:] = f.cpu().numpy()
    e2 = []
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.mt2d_run(hf2, levels)
        o2d = eng.mt2d_fetch()
        e2.append(time.perf_counter() - t0)
    e2e2 = {"value": n * n / min(e2[1:]) / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": int(hf2.nbytes),
            "d2h_bytes_per_step": int(o2d["level"].nbytes + o2d["keys"].nbytes + o2d["pos"].nbytes)}
    del o2d
    print(json.dumps({"path": "2D marching triangles (BASELINE configs[1])", "metric": "Gsamples/s", "value": n * n / ms / 1e6,
                      "cpu_baseline": {"value": cpu2, "unit": "Gsamples/s", "cores": 1, "kind": "port",
                                       "sample": "1024^2 block from the middle of the same field, 16 levels, numpy oracle port"},
                      "e2e": e2e2,
                      "msegments_per_s": S / ms / 1e3, "ms_per_step": ms, "n_segments": S, "levels": 16, "dtype": "f32",
                      "stage_ms": {"count_scan": float(st[1]), "emit": float(st[2])},
                      "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": pk, "unit": "GB/s",
                                   "frac": alg / ms / 1e6 / pk, "algorithmic_bytes": alg, "peak_source": src}}))
    # ---- 2D polylines end to end (a23 / f4): host array in, {level: [(closed, points)]} out, through the drop-in class
    from contourist_b200 import grid_field, multiple_2d_contour, triangulated
    hf = hf2                                                       # page-locked host field, as the e2e contract asks
    grid = grid_field.FunctionGrid((0.0, 0.0), (n - 1.0, n - 1.0), (1.0, 1.0), hf)
    C = multiple_2d_contour.Multiple2DContourGrid(grid, levels)
    D = C.get_contours_dictionary()                                # warm-up (buffers)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        D = C.get_contours_dictionary()
        ts.append(time.perf_counter() - t0)
    n_poly = sum(len(v) for v in D.values())
    n_pts = sum(len(p) for v in D.values() for _, p in v)
    # the chaining alone on the device, and the host restatement (numpy list ranking) on one level as CPU baseline
    eng.mt2d_run(f.data_ptr(), levels, shape=(n, n), dtype=np.float32, flags=E.GEOM_F64)
    from contourist_b200.engine import PolyCounts
    import ctypes
    pc = PolyCounts()
    tk = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.lib.ctr_mt2d_polylines(eng.h, ctypes.byref(pc))
        tk.append(time.perf_counter() - t0)
    seg = eng.mt2d_fetch()
    sel = seg["level"] == 7
    t0 = time.perf_counter()
    host = triangulated.chain_segments(seg["keys"][sel], seg["pos"][sel])
    t_host = time.perf_counter() - t0
    print(json.dumps({"path": "2D polylines end to end (BASELINE configs[1]): Multiple2DContourGrid.get_contours_dictionary(), host array in, polylines out",
                      "metric": "Gsamples/s", "value": n * n / min(ts) / 1e9, "ms_total": min(ts) * 1e3,
                      "n_polylines": n_poly, "n_points": n_pts, "levels": 16, "dtype": "f32 field, f64 geometry",
                      "kernel_ms": {"extract": ms, "chain_on_device (ctr_mt2d_polylines, host-timed)": min(tk) * 1e3},
                      "e2e": {"value": n * n / min(ts) / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": int(hf.nbytes),
                              "d2h_bytes_per_step": int(n_pts * 16 + n_poly * 21)},
                      "cpu_baseline": {"value": int(sel.sum()) / t_host / 1e6, "unit": "Msegments/s chained", "cores": 1, "kind": "port",
                                       "sample": "level 7 of the same run (%d segments), triangulated.chain_segments (numpy list ranking, the "
                                                 "restatement of triangulated.py:221-305)" % int(sel.sum())},
                      "device_chain_msegments_per_s": S / min(tk) / 1e6}))
    del f, hf, hf2, seg, D
    torch.cuda.empty_cache()

    # ---- 4D: configs[3]
    f = synthetic.morph4d(args.n4d, args.nt, device=dev)
    shape = tuple(f.shape)
    acc[:] = 0

    def step4():
        c = eng.mp4d_run(f.data_ptr(), 1.2, shape=shape, dtype=np.float32, flags=E.MORPH)
        acc[:] += np.array(eng.stage_times(8))
        return c
    c, ms = timed(step4, args.steps, args.warmup, stream, acc)
    st = acc / args.steps
    V, T, M = int(c.n_verts), int(c.n_tets), int(c.n_morph_tris)
    nsamp = float(np.prod(shape))
    alg = nsamp * 4 + V * 16 + T * 16 + M * 24                   # field + 4D points + tetrahedra + morph triangles (3 segments x 2 ids)
    sub4 = f[56:72, 56:72, 56:72, 0:40].cpu().numpy()
    t0 = time.perf_counter()
    r4 = o4.extract(sub4, 1.2, np.float32)
    bp = o4.bin_times(r4["pos"].astype(np.float64), sub4.shape[3] - 1)
    cpu4 = sub4.size / (time.perf_counter() - t0) / 1e9
    hf4 = eng.pinned_empty("bench_field4d", shape, np.float32)
    hf4[:] = f.cpu().numpy()
    e4 = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.mp4d_run(hf4, 1.2, flags=E.MORPH)
        o4d = eng.mp4d_fetch()
        e4.append(time.perf_counter() - t0)
    e2e4 = {"value": nsamp / min(e4[1:]) / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": int(hf4.nbytes),
            "d2h_bytes_per_step": int(sum(v.nbytes for v in o4d.values() if isinstance(v, np.ndarray)))}
    del o4d
    print(json.dumps({"path": "4D marching pentatopes + morph triangles (BASELINE configs[3])", "metric": "Gsamples/s",
                      "cpu_baseline": {"value": cpu4, "unit": "Gsamples/s", "cores": 1, "kind": "port",
                                       "sample": "16x16x16x40 block of the same field, numpy oracle port (extract + bin_times)"},
                      "e2e": e2e4,
                      "value": nsamp / ms / 1e6, "mtets_per_s": T / ms / 1e3, "ms_per_step": ms, "shape": list(shape),
                      "n_verts": V, "n_tets": T, "n_morph_tris": M, "dtype": "f32",
                      "stage_ms": {"bitplane": float(st[1]), "count_scan": float(st[2]), "emit_verts": float(st[3]),
                                   "emit_tets": float(st[4]), "morph": float(st[5])},
                      "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": pk, "unit": "GB/s",
                                   "frac": alg / ms / 1e6 / pk, "algorithmic_bytes": alg, "peak_source": src}}))


if __name__ == "__main__":
    main()
