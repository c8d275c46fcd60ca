import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np
from circom_compat_b200 import Context, Groth16, fr_to_mont, synth, release_all
c0 = Context(0)
circ = synth.chain_circuit(1 << 20); w = synth.chain_witness(1 << 20)
pk, td = synth.setup(c0, circ); cm = circ.matrices(); wm = fr_to_mont(w)
ref = None
for env in sys.argv[1:]:
    kv = dict(x.split('=') for x in env.split(','))
    for k in ('B2G_MSM_CHUNK', 'B2G_MSM_CHUNK_G2'):
        os.environ.pop(k, None)
    os.environ.update(kv)
    cx = Context(0)
    p = Groth16.create_proof_with_reduction_and_matrices(pk, 5, 7, cm, circ.num_inputs, circ.num_constraints, wm, cx)
    ref = ref or p.data
    assert p.data == ref
    g1 = cx.bench_msm(pk, cm, 0, 5); g2 = cx.bench_msm(pk, cm, 4, 3); dev = cx.bench_device(pk, cm, 6)
    print(env, 'G1 acc %.3f whole %.3f | G2 acc %.3f whole %.3f | proof %.2f ms' % (g1[1], g1[0], g2[1], g2[0], dev), flush=True)
    cx.close()
release_all()
