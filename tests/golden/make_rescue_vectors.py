#!/usr/bin/env python3
"""Generates tests/golden/rescue_vectors.json from a pure-Python RPO / RPX (big-integer arithmetic, pow() for x^(1/7), polynomial
arithmetic modulo x^3 - x - 1 for the RPX (E) round) that the script first pins on the reference's own 19 RPO known answers
(tests/golden/rpo_reference_vectors.json = rescue/rpo/tests.rs:241-430).  RPX has no known answers in the reference tree; it
shares the MDS, the round constants and both S-boxes with RPO, and its (E) round is restated here from the definition.

  * `perm`: both permutations on a few states;
  * `lmcs`: roots of the algebraic LMCS (overwrite-mode sponge, rate 8; TruncatedPermutation nodes; crates/stateful-hasher/src/
    field_sponge.rs:41-59, crates/lifted-stark/src/lmcs/lifted_tree.rs:202-284,363-417) under each permutation;
  * `challenger`: a scripted DuplexChallenger<_, P, 12, 8> session (observe / sample / sample_bits) under each permutation.
"""
import json
import os
import re

P = 0xFFFFFFFF00000001
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
inc = open(os.path.join(ROOT, "oracle", "rescue_constants.inc")).read()


def table(name):
    body = re.search(r"%s\[\d+\] = \{(.*?)\};" % name, inc, re.S).group(1)
    return [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]+)ULL", body)]


ROW, ARK1, ARK2 = table("RESCUE_MDS_ROW"), table("RESCUE_ARK1"), table("RESCUE_ARK2")
INV7 = pow(7, -1, P - 1)
assert INV7 == 10540996611094048183


def mds(s):
    return [sum(ROW[(j - i) % 12] * s[j] for j in range(12)) % P for i in range(12)]


def fb(s, r):
    s = [pow((x + k) % P, 7, P) for x, k in zip(mds(s), ARK1[12 * r:12 * r + 12])]
    return [pow((x + k) % P, INV7, P) for x, k in zip(mds(s), ARK2[12 * r:12 * r + 12])]


def rpo(s):
    for r in range(7):
        s = fb(s, r)
    return s


def polymul(a, b):
    c = [0] * 5
    for i in range(3):
        for j in range(3):
            c[i + j] += a[i] * b[j]
    # x^3 = x + 1, x^4 = x^2 + x
    return [(c[0] + c[3]) % P, (c[1] + c[3] + c[4]) % P, (c[2] + c[4]) % P]


def ext(s, r):
    s = [(x + k) % P for x, k in zip(s, ARK1[12 * r:12 * r + 12])]
    out = []
    for k in range(0, 12, 3):
        a = s[k:k + 3]
        acc = [1, 0, 0]
        for _ in range(7):
            acc = polymul(acc, a)
        out += acc
    return out


def rpx(s):
    s = ext(fb(s, 0), 1)
    s = ext(fb(s, 2), 3)
    s = ext(fb(s, 4), 5)
    return [(x + k) % P for x, k in zip(mds(s), ARK1[72:84])]


def hash_elements(perm, e):
    st = [0] * 12
    st[8] = len(e) % 8
    i = 0
    for x in e:
        st[i] = x; i += 1
        if i == 8:
            st = perm(st); i = 0
    if i:
        st[i:8] = [0] * (8 - i)
        st = perm(st)
    return st[:4]


ref = json.load(open(os.path.join(HERE, "rpo_reference_vectors.json")))["hash_elements_prefixes"]
for n, want in enumerate(ref, 1):
    assert hash_elements(rpo, list(range(n))) == want, n


def splitmix(x):
    M = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & M
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return (z ^ (z >> 31)) % P


def bitrev(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


def absorb(perm, st, row):
    st = list(st)
    for off in range(0, len(row), 8):
        chunk = list(row[off:off + 8])
        st[:8] = chunk + [0] * (8 - len(chunk))
        st = perm(st)
    return st


def lmcs_root(perm, mats):
    H = mats[-1][0]
    states = [[0] * 12] * mats[0][0]
    for height, width, rows in mats:
        if height > len(states):
            f = height // len(states)
            states = [s for s in states for _ in range(f)]
        states = [absorb(perm, states[r], rows[r]) for r in range(height)]
    lg = H.bit_length() - 1
    layer = [states[bitrev(i, lg)][:4] for i in range(H)]
    while len(layer) > 1:
        layer = [perm(layer[2 * i] + layer[2 * i + 1] + [0] * 4)[:4] for i in range(len(layer) // 2)]
    return layer[0]


class Duplex:
    def __init__(self, perm, capacity):
        self.perm, self.st, self.inp, self.out = perm, [0] * 8 + list(capacity), [], []

    def duplex(self):
        if self.inp:
            n = len(self.inp)
            self.st[:8] = self.inp + [0] * (8 - n)
            self.st[8] = (self.st[8] + n) % P          # length tag (random_coin.masm:103-115)
            self.inp = []
        self.st = self.perm(self.st)
        self.out = self.st[:8]

    def observe(self, v):
        self.out = []
        self.inp.append(v)
        if len(self.inp) == 8:
            self.duplex()

    def sample(self):
        if self.inp or not self.out:
            self.duplex()
        return self.out.pop()

    def bits(self, b):
        return self.sample() & ((1 << b) - 1)


out = {"_about": __doc__.strip().splitlines()[0], "perm": {}, "lmcs": {}, "challenger": {}}
states = [[0] * 12, [P - 1] * 12, list(range(12))] + [[splitmix(100 * k + i) for i in range(12)] for k in range(5)]
for name, perm in (("rpo", rpo), ("rpx", rpx)):
    out["perm"][name] = [{"in": s, "out": perm(s)} for s in states]
    trees = []
    for shapes in ([(8, 3)], [(4, 2), (16, 5)], [(2, 1), (8, 9), (8, 4), (16, 2)], [(8, 0), (8, 17)]):
        mats, seed = [], 1
        for (height, width) in shapes:
            mats.append((height, width, [[splitmix(seed * 1000003 + r * 1009 + c) for c in range(width)] for r in range(height)]))
            seed += 1
        trees.append({"shapes": shapes, "rows": [m[2] for m in mats], "root": lmcs_root(perm, mats)})
    out["lmcs"][name] = trees
    ch = Duplex(perm, (837197885082815666, 17812429367884914, 12945170128166309606, 6547471563106428306))
    script, res = [], []
    for op, arg in [("observe", 27), ("observe", 16), ("sample", 0), ("sample", 0), ("bits", 12), ("observe", P - 1), ("observe", 0), ("sample", 0),
                    ("bits", 23)] + [("observe", i) for i in range(9)] + [("sample", 0)] * 9 + [("observe", 5), ("bits", 1), ("sample", 0)]:
        script.append([op, arg])
        if op == "observe":
            ch.observe(arg); res.append(0)
        elif op == "sample":
            res.append(ch.sample())
        else:
            res.append(ch.bits(arg))
    out["challenger"][name] = {"capacity": list(ch.st[8:12]) and [837197885082815666, 17812429367884914, 12945170128166309606, 6547471563106428306],
                               "script": script, "results": res}
json.dump(out, open(os.path.join(HERE, "rescue_vectors.json"), "w"))
print("wrote rescue_vectors.json")
