"""Host-side mirror of the proving entry points ark-circom users call, running on libb2groth.so.

    Groth16.create_proof_with_reduction_and_matrices(pk, r, s, matrices, num_inputs, num_constraints, full_assignment)
        <- Groth16::<Bn254, CircomReduction>::create_proof_with_reduction_and_matrices
           (/root/reference/src/zkey.rs:903-912, benches/groth16.rs:52-61)
    CircomReduction.witness_map_from_matrices(matrices, num_inputs, num_constraints, full_assignment)
        <- /root/reference/src/circom/qap.rs:23-88
    Groth16.prove(pk, matrices, full_assignment, rng)
        <- Groth16::<Bn254, CircomReduction>::prove (src/zkey.rs:866): draws r then s, then the call above.
Arguments keep the reference's meaning; field elements are (n, 4) uint64 Montgomery limb arrays (fr_to_mont).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .zkey import ConstraintMatrices, ProvingKey, R_MOD


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


def _c(a, dtype=np.uint64):
    return np.ascontiguousarray(a, dtype=dtype)


# Device-resident keys / matrices are cached per (host object, device, shard): every Context of that device and shard
# can use them, so several proofs can be in flight on one GPU (one Context per in-flight proof) without duplicating
# the 6 GiB of tables.  release(obj) / release_all() free them.
_PK_HANDLES, _MAT_HANDLES = {}, {}


def release(obj):
    for cache, free in ((_PK_HANDLES, 'b2g_pk_free'), (_MAT_HANDLES, 'b2g_matrices_free')):
        for key in [k for k in cache if k[0] == id(obj)]:
            getattr(N.lib(), free)(cache.pop(key)[0])


def release_all():
    for cache, free in ((_PK_HANDLES, 'b2g_pk_free'), (_MAT_HANDLES, 'b2g_matrices_free')):
        for key in list(cache):
            getattr(N.lib(), free)(cache.pop(key)[0])


class Context:
    """One b2g_ctx = one in-flight proof on one GPU (optionally one shard of a base-range-sharded prover)."""

    def __init__(self, device: int = 0, shard_rank: int = 0, shard_count: int = 1):
        self._h = C.c_void_p()
        N.check(N.lib().b2g_ctx_create(device, shard_rank, shard_count, C.byref(self._h)))
        self.device, self.shard_rank, self.shard_count = device, shard_rank, shard_count

    def close(self):
        if self._h:
            N.lib().b2g_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pk_handle(self, pk: ProvingKey):
        key = (id(pk), self.device, self.shard_rank, self.shard_count)
        if key not in _PK_HANDLES:
            d = N.PkDesc()
            d.n_vars, d.n_public, d.domain_size = pk.n_vars, pk.n_public, pk.domain_size
            keep = {}
            for name in ('alpha_g1', 'beta_g1', 'delta_g1', 'beta_g2', 'delta_g2', 'a_query', 'b_g1_query', 'b_g2_query', 'l_query', 'h_query'):
                keep[name] = _c(getattr(pk, name))
                setattr(d, name, keep[name].ctypes.data if keep[name].size else None)
            h = C.c_void_p()
            N.check(N.lib().b2g_pk_load(self._h, C.byref(d), C.byref(h)))
            _PK_HANDLES[key] = (h, pk)
        return _PK_HANDLES[key][0]

    def mat_handle(self, m: ConstraintMatrices, n_vars: int, reduction: int = N.REDUCTION_CIRCOM):
        key = (id(m), self.device, n_vars, reduction)
        if key not in _MAT_HANDLES:
            d = N.MatDesc()
            d.num_constraints, d.num_inputs, d.n_vars, d.reduction = m.num_constraints, m.num_instance_variables, n_vars, reduction
            keep = [_c(m.a[0], np.uint32), _c(m.a[1], np.uint32), _c(m.a[2]), _c(m.b[0], np.uint32), _c(m.b[1], np.uint32), _c(m.b[2])]
            names = ['a_rowptr', 'a_col', 'a_val', 'b_rowptr', 'b_col', 'b_val']
            if reduction == N.REDUCTION_LIBSNARK:
                if m.c is None:
                    raise ValueError("LibsnarkReduction needs the C matrix (R1CS route); zkey matrices have none")
                keep += [_c(m.c[0], np.uint32), _c(m.c[1], np.uint32), _c(m.c[2])]
                names += ['c_rowptr', 'c_col', 'c_val']
            for name, arr in zip(names, keep):
                setattr(d, name, arr.ctypes.data if arr.size else None)
            h = C.c_void_p()
            N.check(N.lib().b2g_matrices_load(self._h, C.byref(d), C.byref(h)))
            _MAT_HANDLES[key] = (h, m)
        return _MAT_HANDLES[key][0]

    def prepare(self, pk: ProvingKey, matrices: ConstraintMatrices, reduction_id: int = N.REDUCTION_CIRCOM) -> None:
        """load (pk, matrices) and allocate this context's scratch now instead of inside the first proof"""
        N.check(N.lib().b2g_ctx_prepare(self._h, self.pk_handle(pk), self.mat_handle(matrices, pk.n_vars, reduction_id)))

    def p2p_export(self) -> bytes:
        buf = np.zeros(N.IPC_HANDLE_BYTES, dtype=np.uint8)
        N.check(N.lib().b2g_p2p_export(self._h, _ptr(buf)))
        return buf.tobytes()

    def p2p_import(self, handles) -> None:
        blob = np.frombuffer(b''.join(handles), dtype=np.uint8).copy()
        N.check(N.lib().b2g_p2p_import(self._h, _ptr(blob), len(handles)))

    def last_timings(self) -> dict:
        buf = (C.c_float * 16)()
        N.check(N.lib().b2g_last_timings(self._h, buf))
        names = ('h2d', 'witness_map', 'msm_h', 'msm_l', 'msm_a', 'msm_b1', 'msm_b2', 'glue_d2h', 'total', 'host_upload_enqueued', 'host_all_enqueued', 'host_wait')
        return dict(zip(names, list(buf)[:12]))

    def bench_device(self, pk, matrices, iters: int) -> float:
        """average CUDA-event ms of witness map + 5 MSMs + glue with the witness already resident in HBM"""
        ms = C.c_float()
        N.check(N.lib().b2g_bench_device(self._h, self.pk_handle(pk), self.mat_handle(matrices, pk.n_vars), iters, C.byref(ms)))
        return ms.value

    def bench_msm(self, pk, matrices, query: int, iters: int):
        """(whole-MSM ms, accumulate-kernel ms) for one query run alone: 0 H, 1 L, 2 A, 3 B1, 4 B2"""
        out = (C.c_float * 2)()
        N.check(N.lib().b2g_bench_msm(self._h, self.pk_handle(pk), self.mat_handle(matrices, pk.n_vars), query, iters, out))
        return out[0], out[1]

    def launch_count(self) -> int:
        v = C.c_uint64()
        N.check(N.lib().b2g_launch_count(self._h, C.byref(v)))
        return v.value

    # kernel-level entry points -------------------------------------------------------------
    def msm_g1(self, bases, scalars, scalars_mont=False) -> np.ndarray:
        bases, scalars = _c(bases), _c(scalars)
        n = min(bases.size // 8, scalars.size // 4)
        out = np.zeros(8, dtype=np.uint64)
        N.check(N.lib().b2g_msm_g1(self._h, _ptr(bases), _ptr(scalars), n, int(scalars_mont), _ptr(out)))
        return out

    def msm_g2(self, bases, scalars, scalars_mont=False) -> np.ndarray:
        bases, scalars = _c(bases), _c(scalars)
        n = min(bases.size // 16, scalars.size // 4)
        out = np.zeros(16, dtype=np.uint64)
        N.check(N.lib().b2g_msm_g2(self._h, _ptr(bases), _ptr(scalars), n, int(scalars_mont), _ptr(out)))
        return out

    def ntt(self, data_mont, inverse=False) -> np.ndarray:
        d = _c(data_mont).copy()
        n = d.size // 4
        log_n = n.bit_length() - 1
        if n == 0 or (1 << log_n) != n:
            raise ValueError("ntt length must be a power of two")
        N.check(N.lib().b2g_ntt(self._h, _ptr(d), log_n, int(inverse)))
        return d

    def fixed_base_g1(self, scalars_canon) -> np.ndarray:
        sc = _c(scalars_canon); n = sc.size // 4
        out = np.zeros((n, 8), dtype=np.uint64)
        N.check(N.lib().b2g_fixed_base_g1(self._h, _ptr(sc), n, _ptr(out)))
        return out

    def fixed_base_g2(self, scalars_canon) -> np.ndarray:
        sc = _c(scalars_canon); n = sc.size // 4
        out = np.zeros((n, 16), dtype=np.uint64)
        N.check(N.lib().b2g_fixed_base_g2(self._h, _ptr(sc), n, _ptr(out)))
        return out

    def test_op(self, op: int, a, b=None) -> np.ndarray:
        a = _c(a); b = _c(b) if b is not None else None
        words = 4 if (op <= 7 or 14 <= op <= 16) else (16 if op in (9, 11, 13) else 8)
        n = a.size // words
        out = np.zeros_like(a)
        N.check(N.lib().b2g_test_op(self._h, op, _ptr(a), _ptr(b) if b is not None else None, n, _ptr(out)))
        return out


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


@dataclass
class Proof:
    """Proof<Bn254>{a: G1Affine, b: G2Affine, c: G1Affine}; `data` is the 256-byte uncompressed canonical view."""
    data: bytes

    def _int(self, i):
        return int.from_bytes(self.data[32 * i:32 * i + 32], 'little')

    @property
    def a(self):
        return (self._int(0), self._int(1))

    @property
    def b(self):
        return ((self._int(2), self._int(3)), (self._int(4), self._int(5)))

    @property
    def c(self):
        return (self._int(6), self._int(7))


class PendingProof:
    """A proof submitted with Groth16.submit; keeps the witness and output buffers alive until wait()."""

    def __init__(self, ctx, out, w):
        self._ctx, self._out, self._w = ctx, out, w

    def wait(self) -> Proof:
        N.check(N.lib().b2g_prove_wait(self._ctx._h))
        self._w = None
        return Proof(self._out.tobytes())


def _scalar_bytes(v) -> np.ndarray:
    if isinstance(v, (int, np.integer)):
        return np.frombuffer((int(v) % R_MOD).to_bytes(32, 'little'), dtype='<u8').copy()
    a = _c(v).reshape(-1)
    if a.size != 4:
        raise ValueError("scalar must be an int or 4 uint64 limbs (canonical)")
    return a


def fr_rand(rng) -> int:
    """Fr::rand of ark-ff 0.5, the rule Groth16::prove uses for r and s: four u64 limbs from the rng
    (limb 0 first), the top two bits of limb 3 cleared, rejected and redrawn if >= r, and the limbs taken AS the Montgomery
    representation (value = limbs * R^-1 mod r).  `rng` supplies 64-bit words through next_u64() if it has one (an adapter
    over a rand-compatible stream), else through getrandbits(64) (random.Random, secrets.SystemRandom).  Same rule as
    ark_circom::Fr::rand in the C++ mirror."""
    nxt = rng.next_u64 if hasattr(rng, 'next_u64') else (lambda: rng.getrandbits(64))
    while True:
        limbs = [nxt() & 0xFFFFFFFFFFFFFFFF for _ in range(4)]
        limbs[3] &= 0x3FFFFFFFFFFFFFFF
        v = sum(x << (64 * i) for i, x in enumerate(limbs))
        if v < R_MOD:
            return v * _R_INV_R % R_MOD


_R_INV_R = pow(1 << 256, -1, R_MOD)


class CircomReduction:
    """R1CSToQAP implementation selected by Groth16<Bn254, CircomReduction> (src/circom/qap.rs:12-14)."""
    ID = N.REDUCTION_CIRCOM

    @classmethod
    def witness_map_from_matrices(cls, matrices: ConstraintMatrices, num_inputs: int, num_constraints: int, full_assignment, ctx: Context = None) -> np.ndarray:
        ctx = ctx or default_context()
        if num_inputs != matrices.num_instance_variables or num_constraints != matrices.num_constraints:
            raise ValueError("num_inputs / num_constraints disagree with the matrices")
        w = _c(full_assignment)
        n_vars = w.size // 4
        mh = ctx.mat_handle(matrices, n_vars, cls.ID)
        n = 1
        while n < num_constraints + num_inputs:
            n <<= 1
        h = np.zeros((n, 4), dtype=np.uint64)
        dom = C.c_uint32()
        N.check(N.lib().b2g_witness_map(ctx._h, mh, _ptr(w), _ptr(h), C.byref(dom)))
        assert dom.value == n
        return h


class LibsnarkReduction(CircomReduction):
    """ark-groth16's default R1CSToQAP (`Groth16<Bn254>` in /root/reference/tests/groth16.rs:9,25-35): h = coefficients of
    (a*b - c)/Z, c from the real C matrix; keys from generate_random_parameters_with_reduction carry domain_size - 1 H bases."""
    ID = N.REDUCTION_LIBSNARK


class Groth16:
    """Groth16::<Bn254, QAP>; QAP = CircomReduction (snarkjs keys, default here) or LibsnarkReduction (arkworks keys)."""

    @staticmethod
    def create_proof_with_reduction_and_matrices(pk: ProvingKey, r, s, matrices: ConstraintMatrices, num_inputs: int,
                                                 num_constraints: int, full_assignment, ctx: Context = None, reduction=CircomReduction) -> Proof:
        ctx = ctx or default_context()
        if num_inputs != matrices.num_instance_variables or num_constraints != matrices.num_constraints:
            raise ValueError("num_inputs / num_constraints disagree with the matrices")
        w = _c(full_assignment)
        if w.size // 4 != pk.n_vars:
            raise ValueError("full_assignment length != n_vars")
        ph, mh = ctx.pk_handle(pk), ctx.mat_handle(matrices, pk.n_vars, reduction.ID)
        rr, ss = _scalar_bytes(r), _scalar_bytes(s)
        out = np.zeros(256, dtype=np.uint8)
        N.check(N.lib().b2g_prove(ctx._h, ph, mh, _ptr(rr), _ptr(ss), _ptr(w), _ptr(out)))
        return Proof(out.tobytes())

    @staticmethod
    def submit(pk: ProvingKey, r, s, matrices: ConstraintMatrices, full_assignment, ctx: Context, reduction=CircomReduction) -> 'PendingProof':
        """create_proof_with_reduction_and_matrices without the wait: enqueues the proof on `ctx` (one pending proof per
        context) and returns a handle whose .wait() yields the Proof.  One host thread + K contexts = K proofs in flight."""
        w = _c(full_assignment)
        if w.size // 4 != pk.n_vars:
            raise ValueError("full_assignment length != n_vars")
        ph, mh = ctx.pk_handle(pk), ctx.mat_handle(matrices, pk.n_vars, reduction.ID)
        rr, ss = _scalar_bytes(r), _scalar_bytes(s)
        out = np.zeros(256, dtype=np.uint8)
        N.check(N.lib().b2g_prove_submit(ctx._h, ph, mh, _ptr(rr), _ptr(ss), _ptr(w), _ptr(out)))
        return PendingProof(ctx, out, w)

    @staticmethod
    def prove(pk: ProvingKey, matrices: ConstraintMatrices, full_assignment, rng, ctx: Context = None, reduction=CircomReduction) -> Proof:
        """Draws r then s like create_random_proof_with_reduction (ark-groth16 0.5.0), each with the Fr::rand limb rule
        (fr_rand above): fed the same u64 stream as a seeded arkworks rng, it proves with the same (r, s)."""
        r = fr_rand(rng)
        s = fr_rand(rng)
        return Groth16.create_proof_with_reduction_and_matrices(pk, r, s, matrices, matrices.num_instance_variables,
                                                                matrices.num_constraints, full_assignment, ctx, reduction)

    @staticmethod
    def generate_random_parameters_with_reduction(circuit, rng, ctx: Context = None, reduction=CircomReduction) -> ProvingKey:
        """Setup on the GPU (tests/groth16.rs:25 flow); `circuit` is a synth.Circuit (R1CS as coordinate lists)."""
        from . import synth
        flavour = 'libsnark' if reduction.ID == N.REDUCTION_LIBSNARK else 'circom'
        return synth.generate_random_parameters_with_reduction(circuit, rng, ctx or default_context(), flavour)

    # ---- verification (host pairing; circom_compat_b200/verifier.py).  Call sites in the reference: src/zkey.rs:868-870,
    # 914-916 (process_vk + verify_with_processed_vk), tests/groth16.rs:33-35 (SNARK::verify).
    @staticmethod
    def process_vk(vk):
        """Groth16::process_vk(&params.vk): `vk` is a verifier.VerifyingKey or a ProvingKey (its vk part is used)."""
        from . import verifier
        return verifier.prepare_verifying_key(vk)

    @staticmethod
    def verify_with_processed_vk(pvk, public_inputs, proof) -> bool:
        """public_inputs = w[1..num_inputs] as integers (CircomCircuit::get_public_inputs, src/circom/circuit.rs:18-26)"""
        from . import verifier
        return verifier.verify_with_processed_vk(pvk, public_inputs, proof)

    @staticmethod
    def verify(vk, public_inputs, proof) -> bool:
        from . import verifier
        return verifier.verify(vk, public_inputs, proof)

    # base-range sharded variant: every rank calls prove_partial, the 768-byte partials are all-gathered by the caller
    # (torch.distributed / NCCL), then every rank calls prove_finish and obtains the same proof.
    @staticmethod
    def prove_partial(pk: ProvingKey, matrices: ConstraintMatrices, full_assignment, ctx: Context, r=None, s=None) -> np.ndarray:
        """r, s are optional here: when given, the (r, s)-only scalar multiplications start alongside the MSMs."""
        w = _c(full_assignment)
        ph, mh = ctx.pk_handle(pk), ctx.mat_handle(matrices, pk.n_vars)
        out = np.zeros(N.PARTIAL_BYTES, dtype=np.uint8)
        rr = _scalar_bytes(r) if r is not None else None
        ss = _scalar_bytes(s) if s is not None else None
        N.check(N.lib().b2g_prove_partial(ctx._h, ph, mh, _ptr(rr) if rr is not None else None, _ptr(ss) if ss is not None else None, _ptr(w), _ptr(out)))
        return out

    @staticmethod
    def prove_sharded_p2p(pk: ProvingKey, matrices: ConstraintMatrices, r, s, full_assignment, ctx: Context, reduction=CircomReduction) -> Proof:
        """Sharded proof whose exchange runs inside the kernels over NVLink peer memory (sharding.connect_p2p first)."""
        w = _c(full_assignment)
        ph, mh = ctx.pk_handle(pk), ctx.mat_handle(matrices, pk.n_vars, reduction.ID)
        rr, ss = _scalar_bytes(r), _scalar_bytes(s)
        out = np.zeros(256, dtype=np.uint8)
        N.check(N.lib().b2g_prove_sharded_p2p(ctx._h, ph, mh, _ptr(rr), _ptr(ss), _ptr(w), _ptr(out)))
        return Proof(out.tobytes())

    @staticmethod
    def prove_finish(pk: ProvingKey, partials: np.ndarray, r, s, ctx: Context) -> Proof:
        parts = np.ascontiguousarray(partials, dtype=np.uint8).reshape(-1, N.PARTIAL_BYTES)
        rr, ss = _scalar_bytes(r), _scalar_bytes(s)
        out = np.zeros(256, dtype=np.uint8)
        N.check(N.lib().b2g_prove_finish(ctx._h, ctx.pk_handle(pk), _ptr(parts), parts.shape[0], _ptr(rr), _ptr(ss), _ptr(out)))
        return Proof(out.tobytes())
