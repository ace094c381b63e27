"""CPU: the 2D numpy oracle (oracle/mt2d.py) against golden vectors produced by the unmodified reference and the
reference's own known answers (contourist/test/test_triangulated.py)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mt2d

FILES = sorted(glob.glob(os.path.join(GOLDEN, "mt2d_*.npz")))


def golden_keys(g, li):
    n0, n1 = g["field"].shape
    low, high = g["L%d_low" % li], g["L%d_high" % li]
    pm = np.minimum(low, high)
    d = np.maximum(low, high) - pm
    lowmin = (low == pm).all(axis=1)
    return ((((pm[:, 0] * n1 + pm[:, 1]).astype(np.uint64) * np.uint64(4) + (d[:, 0] * 2 + d[:, 1]).astype(np.uint64))
             << np.uint64(1)) | lowmin.astype(np.uint64))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_keys_positions_match_reference(path):
    g = np.load(path)
    assert len(g["levels"]) >= 3
    for li, z in enumerate(g["levels"]):
        r = mt2d.extract_level(g["field"], z)
        gk = golden_keys(g, li)
        o = np.argsort(gk)
        assert np.array_equal(gk[o], r["keys"])                      # P-keys incl. orientation
        assert np.array_equal(g["L%d_pos" % li][o], r["pos"])        # P-pos, 0 ulp


@pytest.mark.parametrize("name", ["mt2d_osc21.npz", "mt2d_noise.npz"])
def test_segments_match_reference_polylines(name):
    """Generic fields (no sample equals a level): every consecutive pair of points of the reference's
    polylines is one of the oracle's segments and together (with the closing segments) they use every segment
    exactly once.  The reference drops a point that is np.allclose to its predecessor
    (triangulated.py:269); levels where that happened (fewer polyline points than keys) are checked up to the
    dropped points."""
    g = np.load(os.path.join(GOLDEN, name))
    strict_levels = 0
    for li, z in enumerate(g["levels"]):
        r = mt2d.extract_level(g["field"], z)
        lut = {tuple(p): int(k) for k, p in zip(r["keys"], r["pos"])}
        assert len(lut) == len(r["keys"])
        segs = set((int(a), int(b)) for a, b in r["seg_keys"])
        dropped = len(r["keys"]) - int(g["L%d_len" % li].sum())
        assert dropped >= 0
        used, missing, off, ref = set(), 0, 0, []
        for closed, ln in zip(g["L%d_closed" % li], g["L%d_len" % li]):
            pts = g["L%d_pts" % li][off:off + ln]
            off += ln
            ks = [lut[tuple(p)] for p in pts]
            ref.append((bool(closed), tuple(sorted(ks))))
            pairs = list(zip(ks[:-1], ks[1:]))
            if closed and len(ks) > 2:
                pairs.append((ks[-1], ks[0]))
            for a, b in pairs:
                e = (min(a, b), max(a, b))
                if e in segs:
                    used.add(e)
                else:
                    missing += 1
        assert missing <= dropped
        assert len(segs - used) <= 2 * dropped
        if dropped == 0:
            strict_levels += 1
            assert used == segs
            mine = sorted((bool(c), tuple(sorted(int(k) for k in ks)))
                          for c, ks, _ in mt2d.polylines(r["keys"], r["pos"], r["seg_keys"]))
            assert mine == sorted(ref)
    assert strict_levels >= 2


def test_reference_known_answer_line():
    """test_triangulated.py:81-91: x + y = 1.5 on a 2x2 grid -> (1,.5), (.75,.75), (.5,1)."""
    f = np.array([[0.0, 1.0], [1.0, 2.0]])
    r = mt2d.extract_level(f, 1.5)
    [(closed, ks, pts)] = mt2d.polylines(r["keys"], r["pos"], r["seg_keys"])
    assert not closed
    expected = np.array([(1.0, 0.5), (0.75, 0.75), (0.5, 1.0)])
    assert np.allclose(pts, expected) or np.allclose(pts[::-1], expected)


def test_reference_known_answer_dot():
    """test_triangulated.py:93-106: single dot -> closed hexagon."""
    f = np.zeros((3, 3))
    f[1, 1] = 2.0
    r = mt2d.extract_level(f, 1.0)
    [(closed, ks, pts)] = mt2d.polylines(r["keys"], r["pos"], r["seg_keys"])
    assert closed
    expected = [[0.5, 0.5], [1.0, 0.5], [1.5, 1.0], [1.5, 1.5], [1.0, 1.5], [0.5, 1.0]]
    got = [tuple(p) for p in pts]
    assert sorted(got) == sorted(tuple(e) for e in expected)
    # cyclic order up to rotation / reversal
    n = len(expected)
    rots = [expected[s:] + expected[:s] for s in range(n)]
    rots += [list(reversed(x)) for x in rots]
    assert any(np.allclose(np.array(x), pts) for x in rots)


def test_linear_levels_restates_reference_quirk():
    f = np.array([[1.0, 2.0], [3.0, 5.0]])
    assert mt2d.linear_levels(f, 4) == [1.0, 2.0, 3.0]      # (5-1)/4 * i, NOT shifted by the minimum


def test_level_rules_known_answers_from_the_reference_classes():
    """multiple_2d_contour.py:91-108: values the unmodified reference's Linear2DContour(breakpoints=5) and
    Percentile2DContour(breakpoints=4) computed for f = sin(3x+y^2)+cos(4y+x^2) on [-2,2]^2, delta 0.25 (17 x 17)."""
    f = np.array([[np.sin(3 * x + y * y) + np.cos(4 * y + x * x) for y in -2 + 0.25 * np.arange(17)] for x in -2 + 0.25 * np.arange(17)])
    lin = mt2d.linear_levels(f, 5)
    per = mt2d.percentile_levels(f, 4)
    assert np.allclose(lin[:3], [0.7961310269006759, 1.5922620538013519, 2.388393080702028], rtol=1e-12, atol=0) and len(lin) == 4
    assert np.allclose(per[:3], [-0.6838886577943704, 0.05132507684886245, 0.681422313928007], rtol=1e-12, atol=0)
