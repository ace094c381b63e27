"""Synthetic fields of the BASELINE.json configs, generated on the device with torch (plumbing only).

SURVEY.md section 8(d): C3 "CT-like" blobs + smoothed voxel noise; C5 "turbulence" sum of sines;
C1 sphere; C2 2D oscillatory field (misc/2d test.ipynb cell 12); C4 torus->bar morph (triangle_json.py:17-35).
All generators can produce a slab [i0, i1) of planes so multi-GPU ranks never hold the full volume.
"""
import math

import numpy as np
import torch


def ct_like(n, i0=0, i1=None, device="cuda", seed=0, n_total=None, chunk=32):
    """[i1-i0, n, n] fp32 slab of the n_total x n x n CT-like volume (isovalue 0.5).

    n_total > n (a multiple of n) stacks n_total / n CT-like blocks along the first axis: block r has its own 48
    blobs (rng(seed + 7919 r), block-local coordinates), and the Gaussian tails of the neighbouring blocks are summed
    in, so the volume is one continuous field and every n-plane slab carries the same amount of surface (weak
    scaling: per-GPU work fixed).  n_total == n is the single block, C3 itself."""
    n_total = n if n_total is None else n_total
    i1 = n_total if i1 is None else i1
    assert n_total % n == 0, "the stacked volume is a whole number of n-plane blocks"
    nblocks = n_total // n
    nb = 48
    blobs = []
    for r in range(nblocks):
        rng = np.random.default_rng(seed + 7919 * r)
        blobs.append((rng.uniform(0.15, 0.85, size=(nb, 3)), rng.uniform(0.03, 0.12, size=(nb, 3)), rng.uniform(0.5, 1.0, size=nb)))
    out = torch.empty((i1 - i0, n, n), dtype=torch.float32, device=device)
    y = (torch.arange(n, device=device, dtype=torch.float32) / (n - 1)).view(1, n, 1)
    z = (torch.arange(n, device=device, dtype=torch.float32) / (n - 1)).view(1, 1, n)
    gen = torch.Generator(device=device)
    for a in range(i0, i1, chunk):
        b = min(a + chunk, i1)
        acc = torch.zeros((b - a, n, n), dtype=torch.float32, device=device)
        for r in range(max(a // n - 1, 0), min((b - 1) // n + 1, nblocks - 1) + 1):
            cen, sig, amp = blobs[r]
            x = ((torch.arange(a, b, device=device, dtype=torch.float32) - float(r * n)) / (n - 1)).view(-1, 1, 1)
            for q in range(nb):
                ex = ((x - cen[q, 0]) / sig[q, 0]) ** 2
                ey = ((y - cen[q, 1]) / sig[q, 1]) ** 2
                ez = ((z - cen[q, 2]) / sig[q, 2]) ** 2
                acc += float(amp[q]) * torch.exp(-0.5 * (ex + ey + ez))
        # noise needs one extra plane each side for the 3-tap box filter; seeded per global plane
        planes = []
        for gi in range(a - 1, b + 1):
            gen.manual_seed(1_000_003 * (seed + 1) + (gi % (1 << 30)))
            planes.append(torch.randn((n, n), generator=gen, device=device, dtype=torch.float32))
        noise = torch.stack(planes)
        sm = (noise[:-2] + noise[1:-1] + noise[2:]) / 3.0
        sm = (torch.roll(sm, 1, 1) + sm + torch.roll(sm, -1, 1)) / 3.0
        sm = (torch.roll(sm, 1, 2) + sm + torch.roll(sm, -1, 2)) / 3.0
        out[a - i0:b - i0] = acc + 0.02 * sm
    return out


def turbulence(n, i0=0, i1=None, device="cuda", seed=1, n_total=None, chunk=16):
    """[i1-i0, n, n] fp32 slab of the 'turbulence' field: sum of 64 sines, |k| log-uniform (isovalue 0)."""
    n_total = n if n_total is None else n_total
    i1 = n_total if i1 is None else i1
    rng = np.random.default_rng(seed)
    m = 64
    kn = np.exp(rng.uniform(np.log(2 * np.pi * 2), np.log(2 * np.pi * 64), size=m))
    dirs = rng.standard_normal((m, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    kv = dirs * kn[:, None]
    amp = kn ** (-5.0 / 6.0)
    phi = rng.uniform(0, 2 * np.pi, size=m)
    out = torch.empty((i1 - i0, n, n), dtype=torch.float32, device=device)
    y = (torch.arange(n, device=device, dtype=torch.float32) / (n - 1)).view(1, n, 1)
    z = (torch.arange(n, device=device, dtype=torch.float32) / (n - 1)).view(1, 1, n)
    for a in range(i0, i1, chunk):
        b = min(a + chunk, i1)
        x = (torch.arange(a, b, device=device, dtype=torch.float32) / (n_total - 1)).view(-1, 1, 1)
        acc = torch.zeros((b - a, n, n), dtype=torch.float32, device=device)
        for q in range(m):
            acc += float(amp[q]) * torch.sin(float(kv[q, 0]) * x + float(kv[q, 1]) * y + float(kv[q, 2]) * z + float(phi[q]))
        out[a - i0:b - i0] = acc
    return out


def sphere(n, device="cpu", dtype=torch.float64):
    """C1: f = x^2+y^2+z^2 on [-1,1]^3, n samples per axis (isovalue 0.5)."""
    g = torch.linspace(-1.0, 1.0, n, dtype=dtype, device=device)
    return g.view(-1, 1, 1) ** 2 + g.view(1, -1, 1) ** 2 + g.view(1, 1, -1) ** 2


def field2d(n, device="cuda", dtype=torch.float32, lo=-2.0, hi=2.0):
    """C2: f(x,y) = ||(sin(3x+y^2), cos(4y+x^2))|| on [-2,2]^2, n samples per axis."""
    g = torch.linspace(lo, hi, n, dtype=torch.float64, device=device)
    x = g.view(-1, 1)
    y = g.view(1, -1)
    f = torch.sqrt(torch.sin(3 * x + y * y) ** 2 + torch.cos(4 * y + x * x) ** 2)
    return f.to(dtype)


def morph4d(n, nt, device="cuda", dtype=torch.float32):
    """C4: fg(x,y,z,t) = t*3||(x,z)|| + (1-t)*3||(1-||(x,y)||, z)|| on [-2,2]^3 x [0,1] (isovalue 1.2)."""
    g = torch.linspace(-2.0, 2.0, n, dtype=torch.float64, device=device)
    t = torch.linspace(0.0, 1.0, nt, dtype=torch.float64, device=device).view(1, 1, 1, -1)
    x = g.view(-1, 1, 1, 1)
    y = g.view(1, -1, 1, 1)
    z = g.view(1, 1, -1, 1)
    bar = 3 * torch.sqrt(x * x + z * z)
    alpha = torch.sqrt(x * x + y * y)
    g2 = 3 * torch.sqrt((1 - alpha) ** 2 + z * z)
    return (t * bar + (1 - t) * g2).to(dtype)
