#!/usr/bin/env python
"""Profiling driver (run under ncu on the GPU box): synthetic 2^k chain key, one whole proof, then the H (G1) and B2 (G2)
MSMs alone.  Usage: python tools/prof_msm.py [log_n] [iters]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from circom_compat_b200 import Context, Groth16, fr_to_mont, synth  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = Context(0)
circ = synth.chain_circuit(1 << log_n); w = synth.chain_witness(1 << log_n)
t0 = time.time()
pk, td = synth.setup(ctx, circ)
cm = circ.matrices()
print('setup %.1fs' % (time.time() - t0), flush=True)
wm = fr_to_mont(w)
p = Groth16.create_proof_with_reduction_and_matrices(pk, 0x1234, 0x5678, cm, circ.num_inputs, circ.num_constraints, wm, ctx)
print('proof', p.data[:8].hex(), flush=True)
print('G1 msm (whole, accumulate) ms', ctx.bench_msm(pk, cm, 0, iters), flush=True)
print('G2 msm (whole, accumulate) ms', ctx.bench_msm(pk, cm, 4, iters), flush=True)
