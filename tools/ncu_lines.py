"""Attribute ncu per-instruction counters to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [<lib.so>] [--top N]

ncu's CSV source page has one row per SASS instruction (no line numbers); `nvdisasm -g` of the cubin embedded in
the shared library gives the source line of every instruction in the same order.  The two are zipped by ordinal.
The library must be the build that was profiled.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(lib, kernel_re):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    out = {}
    for cub in sorted(os.listdir(tmp)):
        if "-" in cub.split(".")[0]:
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        cur, line, rows = None, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur = m.group(1)
                rows = out.setdefault(cur, [])
                line = None
                continue
            m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur is not None:
                rows.append((line, m.group(2).strip()))
    return {k: v for k, v in out.items() if re.search(kernel_re, k)}


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "contourist_b200/libcontourist_b200.so"
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    # first kernel only: rows until the next "Kernel Name" header
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    kname = rows[hdr_i - 1][1] if hdr_i else "?"
    col = {h: i for i, h in enumerate(hdr)}
    sl = sass_lines(lib, kre)
    # choose the function whose instruction count matches
    cands = [v for v in sl.values() if len(v) == len(body)]
    if not cands:
        print("no cubin function with %d instructions matches %r (have %s)" % (len(body), kre, {k[-40:]: len(v) for k, v in sl.items()}))
        sys.exit(1)
    lines = cands[0]
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for (ln, sass), r in zip(lines, body):
        inst = int(float(r[col["Instructions Executed"]]))
        samp = int(float(r[col["Warp Stall Sampling (All Samples)"]]))
        tin = int(float(r[col["Thread Instructions Executed"]]))
        a = agg[ln]
        a[0] += inst; a[1] += samp; a[2] += tin
        tot[0] += inst; tot[1] += samp; tot[2] += tin
    print("kernel:", kname[:100])
    print("total warp-instructions %d, thread-instructions %d, stall samples %d" % (tot[0], tot[2], tot[1]))
    print("%-26s %12s %6s %10s %6s" % ("line", "warp-inst", "%", "samples", "%"))
    key = 0 if "--by-inst" in sys.argv else 1
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
        name = "%s:%d" % ln if ln else "?"
        print("%-26s %12d %6.2f %10d %6.2f" % (name, a[0], 100.0 * a[0] / max(tot[0], 1), a[1], 100.0 * a[1] / max(tot[1], 1)))


if __name__ == "__main__":
    main()
