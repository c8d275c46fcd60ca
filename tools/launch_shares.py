#!/usr/bin/env python
"""Per-proof kernel shares from an ncu launch list of bench.py (gpu__time_duration.sum per launch):
python tools/launch_shares.py profiles/r2_launches_ntt8.csv > profiles/r2_launch_shares_ntt8.md
One proof = the LAST glue_pre_kernel .. glue_post_kernel span of the list (the captured proof graph's kernel nodes, which ncu
profiles one by one, cold-cache and serialised: compare SHARES, not absolute times)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
rows = [r for r in rows if r['Metric Name'] == 'gpu__time_duration.sum']
names = [r['Kernel Name'] for r in rows]
post = max(i for i, n in enumerate(names) if n.startswith('glue_post_kernel'))
pre = max(i for i, n in enumerate(names[:post]) if n.startswith('glue_pre_kernel'))
span = rows[pre:post + 1]


def short(n):
    n = re.sub(r'\(.*', '', n).replace('void ', '')
    n = n.replace('Curve<Fp<FqParams>>, Fp<FqParams>', 'G1').replace('Curve<Fq2>, Fq2', 'G2')
    return n


agg = collections.OrderedDict()
for r in span:
    k = short(r['Kernel Name'])
    a = agg.setdefault(k, [0, 0.0, 0])
    a[0] += 1; a[1] += float(r['Metric Value'].replace(',', '')) / 1e3
    g = [int(x) for x in re.findall(r'\d+', r['Grid Size'])]
    a[2] = max(a[2], g[0] * g[1] * g[2])
total = sum(a[1] for a in agg.values())
print(f"# Launch list of one proof ({path})\n")
print(f"One proof = launches {pre}..{post} of the list ({len(span)} kernels: glue_pre .. glue_post); ncu replays every kernel node alone "
      "(cold cache, serialised), so compare SHARES.\n")
print("| kernel | launches | us | share of the serial sum | CTAs (largest launch) |")
print("|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.0f} | {100 * a[1] / total:.1f} % | {a[2]} |")
print(f"| serial sum | {len(span)} | {total:.0f} | 100 % | |")
acc = sum(a[1] for k, a in agg.items() if 'msm_accumulate' in k)
ntt = sum(a[1] for k, a in agg.items() if 'ntt_pass' in k)
print(f"\nAccumulation kernels: {acc:.0f} us = {100 * acc / total:.1f} % of the serial sum; NTT passes: {ntt:.0f} us = {100 * ntt / total:.1f} %.")
