"""Summarise an .ncu-rep: per-kernel time, DRAM bytes, throughput, occupancy, instruction counts.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json profiles/traffic.json] [> profiles/xxx.txt]

--json writes the measured DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per kernel launch, and their sum
over the kernels of one extraction) that bench.py reports as roofline.traffic.  The capture must hold exactly one launch
of each kernel of one extraction (tools/run3d_once.py under `ncu --set full -s ... -c ...`).
"""
import json
import os
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ,
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def to_bytes(val, unit):
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val.replace(",", "")) * mult.get(unit, 1.0)


def main():
    rep = sys.argv[1]
    jpath = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    kernels = {}
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('==', r[idx['Kernel Name']][:110])
        if jpath and 'dram__bytes_read.sum' in idx:
            import re
            m = re.search(r'(k\w+)\s*[<(]', r[idx['Kernel Name']])
            name = m.group(1) if m else r[idx['Kernel Name']][:40]
            rd = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']])
            wr = to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
            us = float(r[idx['gpu__time_duration.sum']].replace(",", "")) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units[idx['gpu__time_duration.sum']], 1.0)
            k = kernels.setdefault(name, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "us": 0.0})
            k["launches"] += 1
            k["dram_read_bytes"] += rd
            k["dram_write_bytes"] += wr
            k["us"] += us
        for w in WANT:
            if w in idx:
                print('   %-82s %s %s' % (w, r[idx[w]], units[idx[w]]))


    if jpath:
        total = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in kernels.values())
        with open(jpath, "w") as f:
            json.dump({"source": os.path.basename(rep) + " (ncu --set full, one launch of each kernel of a 512^3 extraction)",
                       "total_bytes": total, "kernels": kernels}, f, indent=1)


if __name__ == '__main__':
    main()
