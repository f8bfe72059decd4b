#!/usr/bin/env python3
"""Generates tests/golden/blake3_vectors.json with the `blake3` package (bindings of the official BLAKE3 crate -- the crate
the reference's `Blake3Hasher` = p3_blake3::Blake3 wraps, crates/crypto/src/hash/blake/mod.rs:16).  The package is in the
build image; the committed vectors make the tests independent of it.

  * `hash`: inputs byte i = i mod 251 (the official test-vector pattern) at lengths around every block / chunk / subtree
    boundary -> 32-byte digests;
  * `lmcs`: the reference's Blake3 LMCS semantics written out in Python on small matrices (ChainingHasher leaves with
    state lifting between heights, digest i = state[bitrev(i)], blake3(left || right) layers;
    crates/stateful-hasher/src/chaining.rs:31-52, crates/lifted-stark/src/lmcs/lifted_tree.rs:202-284,363-417) -> roots;
  * `challenger`: a scripted HashChallenger / SerializingChallenger64 session (observe / sample / sample_bits) -> values.
"""
import json
import os
import struct

import blake3

P = 0xFFFFFFFF00000001
HERE = os.path.dirname(os.path.abspath(__file__))


def h(b):
    return blake3.blake3(bytes(b)).digest()


def splitmix(x):
    x = (x + 0x9E3779B97F4A7C15) & (2**64 - 1)
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    return (z ^ (z >> 31)) % P


def bitrev(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


def lmcs_root(mats):
    """mats: list of (height, width, rows) in ascending height, rows in BIT-REVERSED domain order (the layout of a committed LDE)."""
    H = mats[-1][0]
    states = [bytes(32)] * mats[0][0]
    for height, width, rows in mats:
        if height > len(states):
            f = height // len(states)
            states = [s for s in states for _ in range(f)]          # nearest-neighbour duplication of the running states
        states = [h(states[r] + b"".join(struct.pack("<Q", v) for v in rows[r])) for r in range(height)]
    lg = H.bit_length() - 1
    layer = [states[bitrev(i, lg)] for i in range(H)]
    while len(layer) > 1:
        layer = [h(layer[2 * i] + layer[2 * i + 1]) for i in range(len(layer) // 2)]
    return layer[0]


class HashChallenger:
    def __init__(self, init=b""):
        self.inp, self.out = bytearray(init), bytearray()

    def observe_bytes(self, b):
        self.out = bytearray()
        self.inp += b

    def sample_byte(self):
        if not self.out:
            d = h(self.inp)
            self.out = bytearray(d)
            self.inp = bytearray(d)
        return self.out.pop()

    def sample_u64(self):
        return int.from_bytes(bytes(self.sample_byte() for _ in range(8)), "little")

    def observe_felt(self, v):
        self.observe_bytes(struct.pack("<Q", v))

    def sample_felt(self):
        while True:
            v = self.sample_u64()
            if v < P:
                return v

    def sample_bits(self, bits):
        return self.sample_u64() & ((1 << bits) - 1)


out = {"_about": __doc__.strip().splitlines()[0], "hash": [], "lmcs": [], "challenger": {}}
for n in [0, 1, 7, 8, 32, 63, 64, 65, 96, 127, 128, 440, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 3073, 4096, 4097, 8192, 8193, 31744, 102400]:
    out["hash"].append({"len": n, "digest": h(bytes(i % 251 for i in range(n))).hex()})
for shapes in ([(8, 3)], [(4, 2), (16, 5)], [(2, 1), (8, 9), (8, 4), (32, 2)], [(16, 0), (16, 3)], [(8, 130)]):
    mats, seed = [], 1
    for (height, width) in shapes:
        rows = [[splitmix(seed * 1000003 + r * 1009 + c) for c in range(width)] for r in range(height)]
        mats.append((height, width, rows))
        seed += 1
    out["lmcs"].append({"shapes": shapes, "rows": [m[2] for m in mats], "root": lmcs_root(mats).hex()})
ch = HashChallenger(b"".join(struct.pack("<Q", v) for v in (837197885082815666, 17812429367884914, 12945170128166309606, 6547471563106428306)))
script, res = [], []
for op, arg in [("observe", 27), ("observe", 16), ("sample", 0), ("sample", 0), ("bits", 12), ("observe", P - 1), ("observe", 0), ("sample", 0),
                ("bits", 23), ("sample", 0), ("sample", 0), ("sample", 0), ("sample", 0), ("sample", 0), ("observe", 5), ("bits", 1), ("sample", 0)]:
    script.append([op, arg])
    if op == "observe":
        ch.observe_felt(arg); res.append(0)
    elif op == "sample":
        res.append(ch.sample_felt())
    else:
        res.append(ch.sample_bits(arg))
out["challenger"] = {"initial_input_hex": b"".join(struct.pack("<Q", v) for v in (837197885082815666, 17812429367884914, 12945170128166309606, 6547471563106428306)).hex(),
                     "script": script, "results": res}
json.dump(out, open(os.path.join(HERE, "blake3_vectors.json"), "w"))
print("wrote blake3_vectors.json:", len(out["hash"]), "hashes,", len(out["lmcs"]), "trees")
