#!/usr/bin/env python
"""One rank's share of a base-sharded 2^20 proof on ONE GPU (shard r of R), for different window sizes (B2G_MSM_C):
CUDA-event time of b2g_prove_partial (direct launches).  Usage: python tools/shard_window_sweep.py [R] [c ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from circom_compat_b200 import Context, Groth16, fr_to_mont, synth, release_all  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cs = [int(x) for x in sys.argv[2:]] or [0, 15, 16, 17]
setup_ctx = Context(0)
circ = synth.chain_circuit(1 << 20); w = synth.chain_witness(1 << 20)
pk, td = synth.setup(setup_ctx, circ)
cm = circ.matrices()
wm = fr_to_mont(w)
for c in cs:
    if c:
        os.environ['B2G_MSM_C'] = str(c)
    else:
        os.environ.pop('B2G_MSM_C', None)
    os.environ['B2G_GRAPH'] = '0'
    cx = Context(0, 3, R)
    for _ in range(3):
        Groth16.prove_partial(pk, cm, wm, cx, 5, 7)
    ts = []
    for _ in range(5):
        Groth16.prove_partial(pk, cm, wm, cx, 5, 7)
        ts.append(cx.last_timings())
    best = min(ts, key=lambda t: t['total'])
    print('R', R, 'c', c or 'default', {k: round(v, 3) for k, v in best.items() if not k.startswith('host')}, flush=True)
    cx.close()
    release_all()
