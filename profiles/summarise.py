#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel, and optionally the key
metrics of one `ncu --set full` report.   python profiles/summarise.py launches.csv [prof.ncu-rep]"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_bytes.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sector_hit_rate.pct']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4].split('(')[0], []).append((float(r[-1]) / 1000, r[8], r[7]))
    tot = sum(sum(x[0] for x in v) / len(v) for v in agg.values())
    print("| kernel | launches | mean us | share | grid | block |\n|---|---|---|---|---|---|")
    for k, v in agg.items():
        m = sum(x[0] for x in v) / len(v)
        print("| %s | %d | %.1f | %.1f %% | %s | %s |" % (k, len(v), m, 100 * m / tot, v[0][1], v[0][2]))
    print("sum of means: %.1f us" % tot)


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h = r[0]
    for row in r[2:]:
        print("--", row[h.index("Kernel Name")][:60] if "Kernel Name" in h else "")
        for i, k in enumerate(h):
            if k in KEYS:
                print("   %-62s %s %s" % (k, row[i], r[1][i]))


if __name__ == "__main__":
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        report(sys.argv[2])
