#!/usr/bin/env python3
"""Generates tests/golden/keccak_vectors.json from a pure-Python Keccak-f[1600] that this script first pins on hashlib.

No Keccak-256 (padding 0x01) implementation ships with the image, but SHA3-256 (hashlib / OpenSSL) is the same permutation and
the same sponge with padding byte 0x06: the script builds both hashes on its own permutation, requires SHA3-256 to agree with
hashlib at lengths around every block boundary, and checks the two published Keccak-256 answers ("" and "abc").  The reference's
`KeccakF` / `Keccak256Hash` are p3_keccak 0.6.2 (crates/crypto/src/hash/keccak/mod.rs:18), i.e. exactly this function.

  * `hash`: Keccak-256 of byte i = i mod 251 at lengths around the 136-byte rate -> digests;
  * `perm`: the permutation applied to lanes 0..24;
  * `lmcs`: the reference's Keccak LMCS semantics written out in Python on small matrices (overwrite-mode stateful sponge with
    rate 17 over canonical u64, state lifting between heights, digest i = lanes 0..4 of state[bitrev(i)], PaddingFreeSponge
    nodes; crates/stateful-hasher/src/field_sponge.rs:41-59, serializing_sponge.rs:72-86,
    crates/lifted-stark/src/lmcs/lifted_tree.rs:202-284,363-417, air/src/config.rs:309-353) -> roots;
  * `challenger`: a scripted HashChallenger<u8, Keccak256Hash, 32> / SerializingChallenger64 session -> values.
"""
import hashlib
import json
import os
import struct

P = 0xFFFFFFFF00000001
M = (1 << 64) - 1
HERE = os.path.dirname(os.path.abspath(__file__))


def rotl(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & M if n else x


def round_constants():
    rc, lfsr = [], 1
    for _ in range(24):
        c = 0
        for j in range(7):
            if lfsr & 1:
                c |= 1 << ((1 << j) - 1)
            lfsr = ((lfsr << 1) ^ (0x71 if lfsr & 0x80 else 0)) & 0xFF
        rc.append(c)
    return rc


RC = round_constants()
ROT = {}
x, y = 1, 0
for t in range(24):
    ROT[(x, y)] = ((t + 1) * (t + 2) // 2) % 64
    x, y = y, (2 * x + 3 * y) % 5
ROT[(0, 0)] = 0


def keccak_f(lanes):
    a = {(x, y): lanes[x + 5 * y] for x in range(5) for y in range(5)}
    for rnd in range(24):
        c = [a[(x, 0)] ^ a[(x, 1)] ^ a[(x, 2)] ^ a[(x, 3)] ^ a[(x, 4)] for x in range(5)]
        d = [c[(x - 1) % 5] ^ rotl(c[(x + 1) % 5], 1) for x in range(5)]
        a = {(x, y): a[(x, y)] ^ d[x] for x in range(5) for y in range(5)}
        b = {(y, (2 * x + 3 * y) % 5): rotl(a[(x, y)], ROT[(x, y)]) for x in range(5) for y in range(5)}
        a = {(x, y): b[(x, y)] ^ ((~b[((x + 1) % 5, y)] & M) & b[((x + 2) % 5, y)]) for x in range(5) for y in range(5)}
        a[(0, 0)] ^= RC[rnd]
    return [a[(i % 5, i // 5)] for i in range(25)]


def sponge256(data, pad):
    st = [0] * 25
    data = bytearray(data)
    data.append(pad)
    while len(data) % 136:
        data.append(0)
    data[-1] ^= 0x80
    for off in range(0, len(data), 136):
        for i in range(17):
            st[i] ^= int.from_bytes(data[off + 8 * i: off + 8 * i + 8], "little")
        st = keccak_f(st)
    return b"".join(struct.pack("<Q", v) for v in st[:4])


def keccak256(b):
    return sponge256(b, 0x01)


# ---- pin the permutation and the sponge before using them
for n in [0, 1, 8, 135, 136, 137, 271, 272, 273, 1000]:
    d = bytes(i % 251 for i in range(n))
    assert sponge256(d, 0x06) == hashlib.sha3_256(d).digest(), n
assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
assert keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"


def splitmix(x):
    x = (x + 0x9E3779B97F4A7C15) & M
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return (z ^ (z >> 31)) % P


def bitrev(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


def absorb(st, row):
    """StatefulSponge<KeccakF, 25, 17, 4>::absorb_into: overwrite mode, zero-filled partial chunk, nothing for an empty row"""
    st = list(st)
    for off in range(0, len(row), 17):
        chunk = list(row[off:off + 17])
        chunk += [0] * (17 - len(chunk))
        st[:17] = chunk
        st = keccak_f(st)
    return st


def compress(l, r):
    """CompressionFunctionFromHasher<PaddingFreeSponge<KeccakF, 25, 17, 4>, 2, 4>"""
    return keccak_f(list(l) + list(r) + [0] * 17)[:4]


def lmcs_root(mats):
    H = mats[-1][0]
    states = [[0] * 25] * mats[0][0]
    for height, width, rows in mats:
        if height > len(states):
            f = height // len(states)
            states = [s for s in states for _ in range(f)]
        states = [absorb(states[r], rows[r]) for r in range(height)]
    lg = H.bit_length() - 1
    layer = [states[bitrev(i, lg)][:4] for i in range(H)]
    while len(layer) > 1:
        layer = [compress(layer[2 * i], layer[2 * i + 1]) for i in range(len(layer) // 2)]
    return b"".join(struct.pack("<Q", v) for v in layer[0])


class HashChallenger:
    def __init__(self, init=b""):
        self.inp, self.out = bytearray(init), bytearray()

    def observe_bytes(self, b):
        self.out = bytearray()
        self.inp += b

    def sample_byte(self):
        if not self.out:
            d = keccak256(self.inp)
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


out = {"_about": __doc__.strip().splitlines()[0], "hash": [], "perm": [f"{v:016x}" for v in keccak_f(list(range(25)))], "lmcs": [], "challenger": {}}
for n in [0, 8, 16, 128, 136, 144, 264, 272, 280, 1024, 4352]:        # whole u64 words: what the transcript hashes
    out["hash"].append({"len": n, "digest": keccak256(bytes(i % 251 for i in range(n))).hex()})
for shapes in ([(8, 3)], [(4, 2), (16, 5)], [(2, 1), (8, 17), (8, 18), (32, 2)], [(16, 0), (16, 3)], [(8, 51)], [(4, 34), (8, 16)]):
    mats, seed = [], 1
    for (height, width) in shapes:
        rows = [[splitmix(seed * 1000003 + r * 1009 + c) for c in range(width)] for r in range(height)]
        mats.append((height, width, rows))
        seed += 1
    out["lmcs"].append({"shapes": shapes, "rows": [m[2] for m in mats], "root": lmcs_root(mats).hex()})
init = b"".join(struct.pack("<Q", v) for v in (837197885082815666, 17812429367884914, 12945170128166309606, 6547471563106428306))
ch = HashChallenger(init)
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
out["challenger"] = {"initial_input_hex": init.hex(), "script": script, "results": res}
json.dump(out, open(os.path.join(HERE, "keccak_vectors.json"), "w"))
print("wrote keccak_vectors.json:", len(out["hash"]), "hashes,", len(out["lmcs"]), "trees")
