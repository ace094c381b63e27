"""Load the UNMODIFIED reference (/root/reference/contourist) under Python 3 / numpy 2.

Used ONLY by tests/golden/make_golden.py in the build container to generate the
committed golden fixtures.  Nothing in tests/, bench.py or the product imports this
at run time (the GPU box has no /root/reference).

The shim does no algorithmic edits:
  * numpy aliases removed in numpy>=1.24: np.int, np.float, np.sometrue
  * Python-2 syntax in pentatopes.py / morph_geometry.py: print statements and one
    bare-tuple comprehension are rewritten textually at import time; one dict.keys() that Python 2
    returns as a list (and the code indexes) is wrapped in list()
  * the reference's implicit-relative imports are satisfied by putting the package
    directory itself on sys.path under a private module namespace.
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("CONTOURIST_REF", "/root/reference")
PKG = os.path.join(REF_ROOT, "contourist")

_loaded = {}


def _patch_numpy():
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "sometrue"):
        np.sometrue = np.any


_PRINT_RE = re.compile(r"^(\s*)print\s+(?!\()(.*)$")


def _py3_source(src):
    out = []
    for line in src.split("\n"):
        m = _PRINT_RE.match(line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = line.replace("for l in 0,1]", "for l in (0,1)]")
        # Python 2's dict.keys() is a list (morph_geometry.py:245 indexes it)
        line = line.replace("pair_order = active_pairs.keys()", "pair_order = list(active_pairs.keys())")
        out.append(line)
    return "\n".join(out)


def load(name):
    """Return reference module `name` (e.g. 'tetrahedral') as refcontourist.<name>."""
    _patch_numpy()
    full = "refcontourist." + name
    if full in sys.modules:
        return sys.modules[full]
    if "refcontourist" not in sys.modules:
        pkg = types.ModuleType("refcontourist")
        pkg.__path__ = [PKG]
        sys.modules["refcontourist"] = pkg
    path = os.path.join(PKG, name + ".py")
    with open(path) as f:
        src = _py3_source(f.read())
    mod = types.ModuleType(full)
    mod.__file__ = path
    mod.__package__ = "refcontourist"
    sys.modules[full] = mod
    # implicit relative imports ("import tetrahedral") resolve to the same modules
    for dep in ("grid_field", "triangulated", "surface_geometry", "lp_tools", "tetrahedral",
                "morph_geometry", "field2d", "pentatopes", "multiple_2d_contour"):
        if dep == name:
            continue
        if re.search(r"^\s*import\s+%s\b" % dep, src, re.M) or re.search(
                r"^\s*from\s+\.\s+import\s+%s\b" % dep, src, re.M):
            depmod = load(dep)
            sys.modules.setdefault(dep, depmod)
            setattr(sys.modules["refcontourist"], dep, depmod)
    code = compile(src, path, "exec")
    exec(code, mod.__dict__)
    setattr(sys.modules["refcontourist"], name, mod)
    return mod
