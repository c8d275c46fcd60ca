#!/usr/bin/env python
"""profiles/kernel_profile.json from an ncu CSV of the accumulation kernels (tools/load_width_sweep.sh / tools/prof_msm.py):
the static, per-launch facts bench.py quotes beside its live CUDA-event times (DRAM traffic, multiplier-pipe utilisation).
Usage: python tools/ncu_summary.py profiles/r2_ncu_accumulate.csv 20 > profiles/kernel_profile.json"""
import collections
import csv
import json
import sys

path, log_n = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(path) if not l.startswith('==')]
data = collections.OrderedDict()
for row in csv.DictReader(lines):
    data.setdefault((int(row['ID']), row['Kernel Name']), {})[row['Metric Name']] = float(row['Metric Value'].replace(',', ''))


def avg(kind, metric):
    v = [m[metric] for (i, n), m in data.items() if ('Fq2' in n) == (kind == 'g2') and 'msm_accumulate' in n and metric in m]
    return sum(v) / len(v) if v else None


out = {"source": f"ncu per-launch metrics, {path} (B200, chain 2^{log_n}, cudaLimitMaxL2FetchGranularity = 64)", "log_n": log_n,
       "g1_kernels": "msm_accumulate_kernel<G1>", "g2_kernels": "msm_accumulate_kernel<G2>"}
for k in ('g1', 'g2'):
    out[k + "_time_us_ncu"] = avg(k, 'gpu__time_duration.sum') / 1e3
    out[k + "_dram_bytes_per_launch"] = avg(k, 'dram__bytes_read.sum') + avg(k, 'dram__bytes_write.sum')
    out[k + "_dram_read_bytes"] = avg(k, 'dram__bytes_read.sum')
    out[k + "_dram_write_bytes"] = avg(k, 'dram__bytes_write.sum')
    out[k + "_fmaheavy_pct"] = avg(k, 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active')
print(json.dumps(out, indent=1))
