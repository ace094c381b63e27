#!/usr/bin/env python
"""Per-kernel launch counts, average device time and share of a step from an ncu launch list.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file launches.csv python bench.py --steps 2 --warmup 3 --no-e2e
    python tools/launch_shares.py launches.csv [--last 5] > profiles/rNN_launch_shares.txt

Times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else 5
    rows = [r for r in csv.reader(open(path)) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = OrderedDict()
    for r in rows[hdr_i + 1:]:
        if len(r) <= vi:
            continue
        m = re.search(r"(k\w+)\s*[<(]", r[ki])
        name = m.group(1) if m else r[ki][:40]
        t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        per.setdefault(name, []).append(t)
    avg = OrderedDict((k, sum(v[-last:]) / len(v[-last:])) for k, v in per.items())
    tot = sum(avg.values())
    print("per-launch device time under ncu (cold-cache, serialised: compare SHARES, not absolutes); average of each kernel's last %d launches" % last)
    print("%-28s %9s %10s %7s" % ("kernel", "launches", "avg us", "share"))
    for k, a in avg.items():
        print("%-28s %9d %10.1f %6.1f%%" % (k, len(per[k]), a, 100.0 * a / tot))
    print("sum %.1f us per step" % tot)


if __name__ == "__main__":
    main()
