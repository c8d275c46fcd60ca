#!/usr/bin/env python
"""Witness map of a 2^LOG_N squaring chain through the C ABI, REPS times (for `ncu -k regex:ntt_pass --metrics gpu__time_duration.sum`):
python tools/prof_ntt.py 22 2.  The pass schedule follows the B2G_NTT_* environment at matrices-load time."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from circom_compat_b200 import CircomReduction, Context, fr_to_mont, synth, release  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
circ = synth.chain_circuit(1 << log_n)
wm = fr_to_mont(synth.chain_witness(1 << log_n))
cm = circ.matrices()
ctx = Context(0)
for i in range(reps):
    t = time.time()
    h = CircomReduction.witness_map_from_matrices(cm, circ.num_inputs, circ.num_constraints, wm, ctx)
    print('witness map 2^%d: %.1f ms wall clock incl. host copies' % (log_n, 1e3 * (time.time() - t)), flush=True)
release(cm)
ctx.close()
