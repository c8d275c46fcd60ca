"""Host-side circom .r1cs / .wtns readers with the reference's shape (host I/O, never on the GPU path).

  R1CSFile.new(data) / R1CS.from(file)   <- /root/reference/src/circom/r1cs_reader.rs:54-146 (sections), :161-200 (header,
                                            BN254-only), :203-229 (constraints), :231-249 (wire map), :18-39 (R1CS)
  R1CS.to_circuit()                      <- the matrices CircomCircuit::generate_constraints produces
                                            (src/circom/circuit.rs:30-82: wire i -> Instance(i) / Witness(i - num_inputs),
                                            i.e. column index = wire index; wire_mapping unused, builder.rs:81-82)
  read_wtns(data)                        <- snarkjs witness file (the reference computes witnesses with WASM)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .zkey import R_MOD

_PRIME_LE = bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")   # r1cs_reader.rs:181


class SerializationError(ValueError):
    pass


@dataclass
class Header:
    field_size: int
    prime_size: bytes
    n_wires: int
    n_pub_out: int
    n_pub_in: int
    n_prv_in: int
    n_labels: int
    n_constraints: int


@dataclass
class R1CSFile:
    version: int
    header: Header
    constraints: list            # [(A, B, C)], each a list of (wire index, coefficient int)  - note (index, coeff) order
    wire_mapping: List[int]

    @staticmethod
    def new(data: bytes) -> 'R1CSFile':
        if data[:4] != b'r1cs':
            raise SerializationError("Invalid magic number")
        version, nsec = struct.unpack_from('<II', data, 4)
        if version != 1:
            raise SerializationError("Unsupported version")
        pos, off, size = 12, {}, {}
        for _ in range(nsec):
            t, sz = struct.unpack_from('<IQ', data, pos)
            pos += 12
            off[t], size[t] = pos, sz
            pos += sz
        for t, name in ((1, 'header'), (2, 'constraint'), (3, 'wire2label')):
            if t not in off:
                raise SerializationError(f"No section offset for {name} type found")
        p = off[1]
        field_size = struct.unpack_from('<I', data, p)[0]
        if field_size != 32:
            raise SerializationError("This parser only supports 32-byte fields")
        if size[1] != 32 + field_size:
            raise SerializationError("Invalid header section size")
        prime = data[p + 4:p + 36]
        if prime != _PRIME_LE:
            raise SerializationError("This parser only supports bn256")
        n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints = struct.unpack_from('<IIIIQI', data, p + 36)
        hdr = Header(field_size, prime, n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints)
        p = off[2]

        def lc():
            nonlocal p
            n = struct.unpack_from('<I', data, p)[0]
            p += 4
            out = []
            for _ in range(n):
                w = struct.unpack_from('<I', data, p)[0]
                v = int.from_bytes(data[p + 4:p + 36], 'little')
                if v >= R_MOD:
                    raise SerializationError("coefficient not reduced")
                out.append((w, v))
                p += 36
            return out
        cons = [(lc(), lc(), lc()) for _ in range(n_constraints)]
        if size[3] != n_wires * 8:
            raise SerializationError("Invalid map section size")
        wm = list(struct.unpack_from('<%dQ' % n_wires, data, off[3]))
        if wm and wm[0] != 0:
            raise SerializationError("Wire 0 should always be mapped to 0")
        return R1CSFile(version, hdr, cons, wm)


@dataclass
class R1CS:
    num_inputs: int
    num_aux: int
    num_variables: int
    constraints: list
    wire_mapping: Optional[List[int]]

    @staticmethod
    def from_file(f: R1CSFile) -> 'R1CS':
        ni = 1 + f.header.n_pub_in + f.header.n_pub_out                                        # r1cs_reader.rs:28
        return R1CS(ni, f.header.n_wires - ni, f.header.n_wires, f.constraints, list(f.wire_mapping))

    def to_circuit(self):
        from .synth import Circuit
        mats = []
        for k in range(3):
            rows, cols, vals = [], [], []
            for i, con in enumerate(self.constraints):
                for w, v in con[k]:
                    rows.append(i); cols.append(w); vals.append(v)
            mats.append((np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64), vals))
        return Circuit(self.num_variables, self.num_inputs, len(self.constraints), mats[0], mats[1], mats[2])


def read_wtns(data: bytes) -> List[int]:
    if data[:4] != b'wtns':
        raise SerializationError("not a wtns file")
    nsec = struct.unpack_from('<I', data, 8)[0]
    pos, out, nwit = 12, [], 0
    for _ in range(nsec):
        t, sz = struct.unpack_from('<IQ', data, pos)
        pos += 12
        if t == 1:
            n8 = struct.unpack_from('<I', data, pos)[0]
            if n8 != 32 or int.from_bytes(data[pos + 4:pos + 36], 'little') != R_MOD:
                raise SerializationError("only BN254 witnesses are supported")
            nwit = struct.unpack_from('<I', data, pos + 36)[0]
        elif t == 2:
            out = [int.from_bytes(data[pos + 32 * i:pos + 32 * i + 32], 'little') for i in range(nwit)]
        pos += sz
    return out
