"""Miden production configuration and the synthetic workload named by BASELINE.json.

  * PCS parameters and Fiat-Shamir seeding: reference air/src/config.rs:55-67, 93-98, 188-198, 264-271.
  * Synthetic traces: the reference's bench fills DummyMidenAir traces from `rand` (third-party RNG,
    benches/miden-bench/src/main.rs, testing/airs/miden.rs:101-123: column 0 zero, the rest random);
    its stream cannot be reproduced without Rust, so the generator here is ours and stated:
    value(air, idx) = splitmix64(seed ^ (air << 56) ^ idx) mod p, seed = 2025
    (crates/lifted-stark/src/testing/params.rs:26), column 0 = 0.
  * Statement binding: default `MultiAir::observe` (crates/lifted-air/src/air.rs:307-324).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import air_program
from .binding import Air, Challenger, Lookup, Matrix, PcsParams, Statement, u32p, u64p

P = 0xFFFFFFFF00000001
MIDEN_WIDTHS = (51, 22, 16)        # core / chiplets / poseidon2 main widths
MIDEN_AUX_WIDTHS = (4, 3, 1)       # EF aux widths (air/src/lib.rs:273-274)
RELATION_DIGEST = (837197885082815666, 17812429367884914, 12945170128166309606, 6547471563106428306)
SEED = 2025


def miden_pcs_params() -> PcsParams:
    return PcsParams(3, 2, 7, 4, 12, 27, 16)


def fast_pcs_params() -> PcsParams:
    """Small-PoW variant for quick CPU tests (same structure, cheaper grinding)."""
    return PcsParams(3, 2, 2, 2, 3, 5, 4)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_trace(air_index: int, log_height: int, width: int, seed: int = SEED) -> np.ndarray:
    """Row-major (2^log_height, width) canonical Goldilocks values; column 0 is zero."""
    n = (1 << log_height) * width
    idx = np.arange(n, dtype=np.uint64)
    v = splitmix64(idx ^ np.uint64(seed ^ (air_index << 56)))
    v = np.where(v >= np.uint64(P), v - np.uint64(P), v).astype(np.uint64)
    v = v.reshape(1 << log_height, width)
    v[:, 0] = 0
    return np.ascontiguousarray(v)


class Workload:
    """Owns every buffer a `mdn_statement` / `mdn_matrix[]` points at."""

    def __init__(self, log_heights, widths=MIDEN_WIDTHS, aux_widths=MIDEN_AUX_WIDTHS, seed=SEED,
                 programs=None, num_randomness=2, public_values=(), log_quotient_degrees=None, traces=None,
                 num_aux_values=None, periodic=None, preprocessed=None, lookups=None):
        self.k = len(log_heights)
        self.log_heights = list(log_heights)
        self.widths = list(widths)[: self.k]
        self.aux_widths = list(aux_widths)[: self.k]
        self.traces = traces if traces is not None else [
            synthetic_trace(i, lh, w, seed) for i, (lh, w) in enumerate(zip(self.log_heights, self.widths))]
        self.programs = programs if programs is not None else [air_program.dummy_miden_air() for _ in range(self.k)]
        lqd = log_quotient_degrees if log_quotient_degrees is not None else [3] * self.k
        nav = num_aux_values if num_aux_values is not None else self.aux_widths
        self._airs = (Air * self.k)()
        for i in range(self.k):
            a = self._airs[i]
            a.width, a.aux_width = self.widths[i], self.aux_widths[i]
            a.num_aux_values, a.num_randomness = nav[i], num_randomness
            a.log_quotient_degree = lqd[i]
            a.program_words = len(self.programs[i])
            a.program = self.programs[i].ctypes.data_as(u32p)
            per = periodic[i] if periodic is not None else None
            if per is not None:
                per = np.ascontiguousarray(per, dtype=np.uint64)     # (max_period, n_cols) row-major
                self._periodic_keep = getattr(self, "_periodic_keep", []) + [per]
                a.periodic_values = per.ctypes.data_as(u64p)
                a.num_periodic_columns = per.shape[1]
                a.log_max_period = int(per.shape[0]).bit_length() - 1
        # lowered LookupAir per AIR: (num_columns, program words) or None -> aux trace built on the device
        self._lookups = []
        for i in range(self.k):
            lk = lookups[i] if lookups is not None else None
            if lk is not None:
                prog = np.ascontiguousarray(lk[1], dtype=np.uint32)
                st_ = Lookup(lk[0], len(prog), prog.ctypes.data_as(u32p))
                self._lookups.append((st_, prog))
                self._airs[i].lookup = C.pointer(st_)
        # BaseAir::preprocessed_trace per AIR (None = the AIR declares none); same height as the main trace
        self.preprocessed = None
        if preprocessed is not None and any(m is not None for m in preprocessed):
            self.preprocessed = [None if m is None else np.ascontiguousarray(m, dtype=np.uint64) for m in preprocessed]
            self.preprocessed_matrices = (Matrix * self.k)()
            for i, m in enumerate(self.preprocessed):
                if m is not None:
                    self._airs[i].preprocessed_width = m.shape[1]
                    self.preprocessed_matrices[i] = Matrix(m.ctypes.data_as(u64p), self.log_heights[i], m.shape[1])
        self.public_values = np.array(list(public_values), dtype=np.uint64)
        # default MultiAir::observe: len(air_inputs), air_inputs, max_aux_inputs (0), len(aux_inputs) (0)
        self.observe_felts = np.array([len(self.public_values), *self.public_values, 0, 0], dtype=np.uint64)
        self.statement = Statement(self._airs, self.k,
                                   self.public_values.ctypes.data_as(u64p), len(self.public_values),
                                   self.observe_felts.ctypes.data_as(u64p), len(self.observe_felts))
        self.matrices = (Matrix * self.k)()
        for i, t in enumerate(self.traces):
            self.matrices[i] = Matrix(t.ctypes.data_as(u64p), self.log_heights[i], self.widths[i])

    @property
    def cells(self) -> int:
        return sum((1 << lh) * w for lh, w in zip(self.log_heights, self.widths))


def initial_challenger(params: PcsParams, observe_fn) -> Challenger:
    """`config.challenger()` + `observe_protocol_params` (air/src/config.rs:188-198, 264-271):
    capacity = RELATION_DIGEST, then observe (num_queries, query_pow, deep_pow, folding_pow,
    log_blowup, log_final_degree, arity, 0).  `observe_fn(challenger, np.uint64[])` performs the
    sponge steps (mdn_challenger_observe from the product, or the oracle's script in tests)."""
    c = Challenger()
    for i in range(4):
        c.sponge_state[8 + i] = RELATION_DIGEST[i]
    felts = np.array([params.num_queries, params.query_pow_bits, params.deep_pow_bits, params.folding_pow_bits,
                      params.log_blowup, params.log_final_degree, 1 << params.log_folding_arity, 0], dtype=np.uint64)
    observe_fn(c, felts)
    return c


def initial_hash_challenger(params: PcsParams) -> bytes:
    """`blake3_256_config(..).challenger()` + `observe_protocol_params` (air/src/config.rs:299-307, 188-198): the
    HashChallenger's input buffer after observing the relation digest and the eight parameter felts, each as its
    canonical u64 in little-endian bytes; nothing has been sampled, so the output buffer is empty."""
    import struct
    felts = list(RELATION_DIGEST) + [params.num_queries, params.query_pow_bits, params.deep_pow_bits, params.folding_pow_bits,
                                     params.log_blowup, params.log_final_degree, 1 << params.log_folding_arity, 0]
    return b"".join(struct.pack("<Q", int(f)) for f in felts)
