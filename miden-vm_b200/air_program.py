"""Builder for the AIR constraint op-list consumed by `mdn_air.program` (include/miden_b200.h).

The reference evaluates `LiftedAir::eval(builder)` -- generic Rust -- at every quotient-domain point
(crates/lifted-stark/src/prover/constraints/mod.rs:233).  A C-ABI backend cannot call that, so the
host lowers `eval` once into a DAG over the leaf/op vocabulary the reference itself enumerates in
crates/ace-codegen/src/dag/lower.rs:109-210 and ships it as a flat list.  This module is the
host-side constructor of that list (what the Rust shim's `SymbolicAirBuilder` capture would emit);
`dummy_miden_air` restates crates/lifted-stark/src/testing/airs/miden.rs:57-62.
"""
from __future__ import annotations

import numpy as np

P = 0xFFFFFFFF00000001
MAGIC = 0x5249414D

OP_MAIN, OP_AUX, OP_PUBLIC, OP_CHALLENGE, OP_AUX_VALUE = 0, 1, 2, 3, 4
OP_IS_FIRST, OP_IS_LAST, OP_IS_TRANSITION, OP_CONST, OP_EXT_CONST = 5, 6, 7, 8, 9
OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_PERIODIC, OP_PREPROCESSED = 10, 11, 12, 13, 14, 15


class Expr:
    __slots__ = ("b", "id")

    def __init__(self, b: "ProgramBuilder", id_: int):
        self.b, self.id = b, id_

    def _bin(self, op, other):
        if not isinstance(other, Expr):
            other = self.b.const(int(other))
        return self.b._node(op, self.id, other.id)

    def __add__(self, o): return self._bin(OP_ADD, o)
    def __sub__(self, o): return self._bin(OP_SUB, o)
    def __mul__(self, o): return self._bin(OP_MUL, o)
    def __neg__(self): return self.b._node(OP_NEG, self.id, 0)


class ProgramBuilder:
    """Mirrors the builder surface an AIR sees (main/aux windows, selectors, public values,
    randomness, permutation values, assert_zero / assert_zero_ext)."""

    def __init__(self):
        self.nodes: list[tuple[int, int, int]] = []
        self.constraints: list[int] = []
        self.consts: list[int] = []

    def _node(self, op, a=0, b=0) -> Expr:
        self.nodes.append((op, a, b))
        return Expr(self, len(self.nodes) - 1)

    def main(self, offset: int, col: int) -> Expr: return self._node(OP_MAIN, offset, col)
    def aux(self, offset: int, col: int) -> Expr: return self._node(OP_AUX, offset, col)
    def public(self, i: int) -> Expr: return self._node(OP_PUBLIC, i)
    def challenge(self, i: int) -> Expr: return self._node(OP_CHALLENGE, i)
    def aux_value(self, i: int) -> Expr: return self._node(OP_AUX_VALUE, i)
    def is_first_row(self) -> Expr: return self._node(OP_IS_FIRST)
    def is_last_row(self) -> Expr: return self._node(OP_IS_LAST)
    def is_transition(self) -> Expr: return self._node(OP_IS_TRANSITION)
    def periodic(self, col: int) -> Expr: return self._node(OP_PERIODIC, col)
    def preprocessed(self, offset: int, col: int) -> Expr: return self._node(OP_PREPROCESSED, offset, col)

    def const(self, v: int) -> Expr:
        self.consts.append(v % P)
        return self._node(OP_CONST, len(self.consts) - 1)

    def ext_const(self, c0: int, c1: int) -> Expr:
        self.consts += [c0 % P, c1 % P]
        return self._node(OP_EXT_CONST, len(self.consts) - 2)

    def assert_zero(self, e: Expr): self.constraints.append(e.id)
    assert_zero_ext = assert_zero

    def serialize(self) -> np.ndarray:
        w = [MAGIC, 1, len(self.nodes), len(self.constraints), len(self.consts)]
        for n in self.nodes:
            w += list(n)
        w += self.constraints
        for c in self.consts:
            w += [c & 0xFFFFFFFF, c >> 32]
        return np.array(w, dtype=np.uint32)


LOOKUP_MAGIC = 0x504B4C4D
NO_FLAG = 0xFFFFFFFF


class LookupProgramBuilder(ProgramBuilder):
    """Host-side lowering of `LookupAir::eval` (air/src/lookup/builder.rs) for the device aux build: the node
    surface a `LookupBuilder` exposes (main window, periodic values, challenges alpha = challenge(0) and
    beta = challenge(1), constants) and one record per interaction -- `insert(column, flag, multiplicity,
    denominator)` is `LookupGroup::insert` / `LookupBatch::insert` after `LookupMessage::encode`."""

    def __init__(self, num_columns: int):
        super().__init__()
        self.num_columns = num_columns
        self.interactions: list[tuple[int, int, int, int]] = []

    def bus_prefix(self, bus: int, max_message_width: int) -> Expr:
        """`Challenges::new`: bus_prefix[i] = alpha + (i + 1) * beta^W (air/src/lookup/challenges.rs)."""
        beta = self.challenge(1)
        g = beta
        for _ in range(max_message_width - 1):
            g = g * beta
        return self.challenge(0) + g * self.const(bus + 1)

    def encode(self, bus: int, max_message_width: int, elems) -> Expr:
        """`Challenges::encode`: bus_prefix[bus] + sum_i beta^i * elems[i]."""
        acc = self.bus_prefix(bus, max_message_width)
        bp = None
        for i, e in enumerate(elems):
            bp = self.const(1) if i == 0 else (self.challenge(1) if i == 1 else bp * self.challenge(1))
            acc = acc + (e if i == 0 else bp * e)
        return acc

    def insert(self, column: int, flag, multiplicity: Expr, denominator: Expr):
        self.interactions.append((column, NO_FLAG if flag is None else flag.id, multiplicity.id, denominator.id))

    def assert_zero(self, e):  # pragma: no cover - a lookup description emits no constraints
        raise TypeError("LookupBuilder has no assert_* surface")

    def serialize(self) -> np.ndarray:
        w = [LOOKUP_MAGIC, 1, len(self.nodes), len(self.interactions), len(self.consts)]
        for n in self.nodes:
            w += list(n)
        for it in self.interactions:
            w += list(it)
        for c in self.consts:
            w += [c & 0xFFFFFFFF, c >> 32]
        return np.array(w, dtype=np.uint32)


def dummy_miden_air() -> np.ndarray:
    """`local[0] * local[1] * ... * local[8] == 0` (testing/airs/miden.rs:57-62): degree 9, so
    log_quotient_degree = 3.  The fold starts from ONE exactly like the reference's `fold`."""
    b = ProgramBuilder()
    acc = b.const(1)
    for j in range(9):
        acc = acc * b.main(0, j)
    b.assert_zero(acc)
    return b.serialize()
