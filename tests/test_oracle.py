"""CPU tests of the oracle itself (oracle/pyref.py big-int, oracle/cref.c restatement) against the reference's
golden vectors and KATs.  Mirrors the unit tests of /root/reference/src/zkey.rs:465-543 and the end-to-end
`verified` assertions of src/zkey.rs:846-919."""
import hashlib
import random

import numpy as np
import pytest

from oracle import cref as c
from oracle import pyref as o


def _limbs_rand(rng, n, mod):
    return c.ints_to_limbs([rng.randrange(mod) for _ in range(n)])


def test_kat_field_and_generators(golden):
    # can_deser_fq / can_deser_g1 / can_deser_g2 (src/zkey.rs:465-517): snarkjs Montgomery encodings of one / G1 / G2
    assert o._fq_from_mont(bytes(golden['kat_fq_one'])) == 1
    assert o._g1_from(bytes(golden['kat_g1_one'])) == o.G1_GEN
    assert o._g2_from(bytes(golden['kat_g2_one'])) == o.G2_GEN
    # the C oracle decodes the same bytes
    one = c.fq_from_mont(np.frombuffer(bytes(golden['kat_fq_one']), dtype='<u8'))
    assert c.limbs_to_ints(one) == [1]
    g1 = c.fq_from_mont(np.frombuffer(bytes(golden['kat_g1_one']), dtype='<u8'))
    assert c.limbs_to_ints(g1) == [1, 2]


def test_deser_key_kat_every_reader(golden, test_zkey_bytes):
    """`fn deser_key` (/root/reference/src/zkey.rs:545-763): all 20 points of test.zkey's IC / A / B1 / B2 / L / H queries as
    the reference pins them (bytes extracted from that test by tests/golden/make_golden.py).  Checked against all three
    readers of this repository: the oracle's (pyref.read_zkey), the product's Python reader (zkey.read_zkey) and the
    product's C++ reader (ark_circom::read_zkey, via groth16_bench --parse-only --dump-key)."""
    import os, re, subprocess
    from circom_compat_b200 import read_zkey
    kat = golden['deser_key']
    assert sum(len(v) for v in kat.values()) == 20
    z = o.read_zkey(test_zkey_bytes)
    pk, _ = read_zkey(test_zkey_bytes)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([os.path.join(root, 'circom_compat_b200', 'host', 'groth16_bench'), '--parse-only',
                                   os.path.join(root, 'tests', 'golden', 'test.zkey'), '--dump-key'], text=True)
    cpp = {}
    for name, idx, hx in re.findall(r'^(\w+)\[(\d+)\]=([0-9a-f]+)$', out, re.M):
        cpp.setdefault(name, []).append(bytes.fromhex(hx))
    oracle_fields = {'gamma_abc_g1': z.ic, 'a_query': z.a_query, 'b_g1_query': z.b_g1_query, 'b_g2_query': z.b_g2_query,
                     'l_query': z.l_query, 'h_query': z.h_query}
    for name, pts in kat.items():
        g2 = name == 'b_g2_query'
        exp = [bytes(b) for b in pts]
        # oracle: decoded affine points (None = infinity, on-curve checked like G1Affine::new) equal the decoded KAT bytes
        assert oracle_fields[name] == [(o._g2_from(b) if g2 else o._g1_from(b)) for b in exp], name
        # product readers keep the zkey's Montgomery bytes verbatim (= arkworks' in-memory Fp256 limbs)
        arr = np.ascontiguousarray(getattr(pk, name))
        assert [arr[i].tobytes() for i in range(arr.shape[0])] == exp, name
        assert cpp[name] == exp, name
    # the points at infinity the reference pins (a_query[3], b_g1_query[0..2], ...) decode to identity in the oracle
    assert z.a_query[3] is None and z.b_g1_query[:3] == [None, None, None]


def test_zkey_header_and_sections(test_zkey_bytes):
    # src/zkey.rs:519-543: n_vars = 4, n_public = 1, domain_size = 4; section sizes
    z = o.read_zkey(test_zkey_bytes)
    assert (z.n_vars, z.n_public, z.domain_size) == (4, 1, 4)
    assert len(z.a_query) == 4 and len(z.b_g2_query) == 4 and len(z.l_query) == 2 and len(z.h_query) == 4 and len(z.ic) == 2
    assert z.num_constraints == 1 and z.num_inputs == 2
    assert z.mat_a == [[(o.R_MOD - 1, 2)]] and z.mat_b == [[(1, 3)]]      # coefficients decode to -1 / +1


def test_verification_key_json(test_zkey_bytes):
    # deser_vk (src/zkey.rs:765-779) against test-vectors/verification_key.json
    import json, os
    vk = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'verification_key.json')))
    z = o.read_zkey(test_zkey_bytes)
    assert z.alpha_g1 == (int(vk['vk_alpha_1'][0]), int(vk['vk_alpha_1'][1]))
    def g2(v): return ((int(v[0][0]), int(v[0][1])), (int(v[1][0]), int(v[1][1])))
    assert z.beta_g2 == g2(vk['vk_beta_2']) and z.gamma_g2 == g2(vk['vk_gamma_2']) and z.delta_g2 == g2(vk['vk_delta_2'])
    assert z.ic == [(int(p[0]), int(p[1])) for p in vk['IC']]


def test_pyref_matches_golden_and_verifies(golden, test_zkey_bytes):
    z = o.read_zkey(test_zkey_bytes)
    g = golden['test_zkey']
    w = [int(x) for x in g['witness']]
    assert [str(x) for x in o.witness_map_from_matrices(z.mat_a, z.mat_b, z.num_inputs, z.num_constraints, w)] == g['h']
    case = g['proofs'][0]
    A, B, C = o.prove(z, int(case['r']), int(case['s']), w)
    assert o.proof_to_bytes(A, B, C).hex() == case['proof_hex']
    assert o.verify(z, w[1:z.num_inputs], (A, B, C))                       # verify_proof_with_zkey_* assert `verified`
    assert not o.verify(z, [34], (A, B, C))                                # tests/groth16.rs:42-74 wrong public input


def test_cref_matches_golden_small(golden, test_zkey_bytes):
    za = c.zkey_arrays(test_zkey_bytes)
    g = golden['test_zkey']
    wm = c.fr_to_mont(c.ints_to_limbs([int(x) for x in g['witness']]))
    for case in g['proofs']:
        pb, h = c.prove(za, int(case['r']), int(case['s']), wm, want_h=True)
        assert pb.hex() == case['proof_hex']
        assert [str(x) for x in c.limbs_to_ints(c.fr_from_mont(h))] == g['h']


def test_cref_matches_golden_complex(golden, complex_zkey_bytes):
    za = c.zkey_arrays(complex_zkey_bytes)
    assert (za['n_vars'], za['domain_size'], za['num_constraints']) == (10002, 16384, 10000)
    g = golden['complex_zkey']
    w = o.chain_witness(za['n_vars'], g['a'])
    wm = c.fr_to_mont(c.ints_to_limbs(w))
    pb, h = c.prove(za, int(g['r']), int(g['s']), wm, want_h=True)
    hc = c.fr_from_mont(h)
    assert [str(x) for x in c.limbs_to_ints(hc[:4])] == g['h_head']
    assert hashlib.sha256(np.ascontiguousarray(hc).tobytes()).hexdigest() == g['h_sha256_canon_le']
    assert pb.hex() == g['proof_hex']


def test_cref_field_vs_bigint():
    rng = random.Random(1)
    a = [rng.randrange(o.Q_MOD) for _ in range(200)] + [0, 1, o.Q_MOD - 1]
    b = [rng.randrange(o.Q_MOD) for _ in range(200)] + [o.Q_MOD - 1, o.Q_MOD - 1, o.Q_MOD - 1]
    am, bm = c.fq_to_mont(c.ints_to_limbs(a)), c.fq_to_mont(c.ints_to_limbs(b))
    assert c.limbs_to_ints(c.fq_from_mont(c.fq_mul(am, bm))) == [x * y % o.Q_MOD for x, y in zip(a, b)]
    a = [rng.randrange(o.R_MOD) for _ in range(200)]
    b = [rng.randrange(o.R_MOD) for _ in range(200)]
    am, bm = c.fr_to_mont(c.ints_to_limbs(a)), c.fr_to_mont(c.ints_to_limbs(b))
    assert c.limbs_to_ints(c.fr_from_mont(c.fr_mul(am, bm))) == [x * y % o.R_MOD for x, y in zip(a, b)]


@pytest.mark.parametrize('log_n', [1, 2, 5, 8])
def test_cref_ntt_vs_bigint(log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    v = [rng.randrange(o.R_MOD) for _ in range(n)]
    vm = c.fr_to_mont(c.ints_to_limbs(v))
    fwd = c.limbs_to_ints(c.fr_from_mont(c.ntt(vm)))
    assert fwd == o.fft(list(v))
    # definition check: X[k] = sum v[j] w^(jk)
    w = o.root_of_unity(n)
    assert fwd[1 % n] == sum(v[j] * pow(w, j * (1 % n), o.R_MOD) for j in range(n)) % o.R_MOD
    inv = c.limbs_to_ints(c.fr_from_mont(c.ntt(c.ntt(vm), inverse=True)))
    assert inv == v


def _rand_points_g1(rng, n):
    return c.fixed_base_g1(c.ints_to_limbs([rng.randrange(1, o.R_MOD) for _ in range(n)]))


def test_cref_msm_vs_bigint_with_edge_cases():
    rng = random.Random(7)
    n = 70
    ks = [rng.randrange(1, o.R_MOD) for _ in range(n)]
    bases = c.fixed_base_g1(c.ints_to_limbs(ks))
    bases[5] = 0                                   # a point at infinity (zkey convention: zeros)
    bases[9] = bases[8]                            # repeated base
    sc = [rng.randrange(o.R_MOD) for _ in range(n)]
    sc[0] = 0; sc[1] = 1; sc[2] = o.R_MOD - 1; sc[8] = 5; sc[9] = o.R_MOD - 5   # P*5 + P*(-5) cancels
    got = c.msm_g1(bases, c.ints_to_limbs(sc))
    pts = [None if not b.any() else tuple(c.limbs_to_ints(c.fq_from_mont(b))) for b in bases]
    exp = o.G1.msm(pts, sc)
    assert tuple(c.limbs_to_ints(c.fq_from_mont(got))) == exp
    # G2
    ks2 = [rng.randrange(1, o.R_MOD) for _ in range(20)]
    b2 = c.fixed_base_g2(c.ints_to_limbs(ks2))
    sc2 = [rng.randrange(o.R_MOD) for _ in range(20)]
    got2 = c.limbs_to_ints(c.fq_from_mont(c.msm_g2(b2, c.ints_to_limbs(sc2))))
    exp2 = o.G2.mul(o.G2_GEN, sum(a * b for a, b in zip(ks2, sc2)) % o.R_MOD)
    assert ((got2[0], got2[1]), (got2[2], got2[3])) == exp2
    # fixed-base spot check
    assert tuple(c.limbs_to_ints(c.fq_from_mont(bases[3]))) == o.G1.mul(o.G1_GEN, ks[3])


def test_msm_truncates_to_shorter_side():
    # msm_bigint pairs min(len(bases), len(scalars)) terms
    rng = random.Random(3)
    bases = _rand_points_g1(rng, 10)
    sc = c.ints_to_limbs([rng.randrange(o.R_MOD) for _ in range(6)])
    assert np.array_equal(c.msm_g1(bases, sc), c.msm_g1(bases[:6], sc))


def test_witness_map_domain_too_large():
    with pytest.raises(ValueError):
        o.domain_size_for((1 << 28) + 1)


def test_field_constants_held_by_the_reference(golden):
    """r as the hex string the reference's witness-calculator test asserts (src/witness/witness_calculator.rs:328-332), R^-1 mod r
    as the decimal literal of src/witness/memory.rs:45-48 and r as the little-endian bytes its .r1cs reader insists on
    (src/circom/r1cs_reader.rs:180-182), extracted by tests/golden/make_golden.py: the oracle's and the product's scalar-field
    modulus and Montgomery radix are these."""
    from circom_compat_b200 import fr_to_mont, fr_from_mont
    from circom_compat_b200.zkey import R_MOD
    k = golden['field_constants']
    r = int(k['r_hex'], 16)
    assert r == o.R_MOD == R_MOD
    assert int.from_bytes(bytes.fromhex(k['r_le_hex']), 'little') == r
    r_inv = int(k['r_inv_dec'])
    assert r_inv == pow(1 << 256, -1, r)
    # the Montgomery form used on the wire (zkey coefficients, witnesses, the C ABI) is x * 2^256 mod r: x_mont * R^-1 = x
    xs = [1, 2, 33, r - 1, 0x1234567890abcdef1234567890abcdef]
    mont = fr_to_mont(xs)
    for x, limbs in zip(xs, mont):
        m = sum(int(v) << (64 * i) for i, v in enumerate(limbs))
        assert m == (x << 256) % r and m * r_inv % r == x
    assert [int(v) for v in fr_from_mont(mont)] == xs
    assert np.array_equal(c.fr_to_mont(c.ints_to_limbs(xs)), mont)
