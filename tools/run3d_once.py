"""One 512^3 (or --n) CT-like extraction, repeated a few times: the command line profiled under ncu.

    python tools/run3d_once.py [--n 512] [--reps 3] [--f64]
Prints the counts and the per-stage CUDA-event times of the last repetition.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from contourist_b200 import engine as E          # noqa: E402
from contourist_b200 import synthetic            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--f64", action="store_true")
    ap.add_argument("--field", default="ct")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    if args.field == "ct":
        f, iso = synthetic.ct_like(args.n), 0.5
    else:
        f, iso = synthetic.turbulence(args.n), 0.0
    torch.cuda.synchronize()
    eng = E.Engine(0)
    eng.set_timing(True)
    flags = E.WANT_NORMALS | (E.GEOM_F64 if args.f64 else 0)
    for _ in range(args.reps):
        c = eng.mt3d_run(f.data_ptr(), iso, flags=flags, shape=tuple(f.shape), dtype="float32")
        st = eng.stage_times(8)
    print(json.dumps(dict(n_verts=c.n_verts, n_tris=c.n_tris, n_cells=c.n_active_cells, stage_ms=st)))


if __name__ == "__main__":
    main()
