// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Poseidon2 permutation over Goldilocks, width 12, and the three constructions the proving path
// builds from it.  Round structure restated from the reference's own round-by-round helpers:
//   crates/crypto/src/hash/algebraic_sponge/poseidon2/mod.rs:226-319 (M_E, M_I, add_rc, x^7)
//   precompiles-prover/src/transcript/poseidon2/trace.rs:380-500 (order of the 1+4+22+4 steps)
// Pinned by the known-answer test poseidon2/test.rs:7-39 (tests/test_oracle_kat.py).
#pragma once
#include "field.hpp"
#include "blake3.hpp"
#include "keccak.hpp"
#include "rescue.hpp"
#include <array>
#include <vector>

namespace orc {

#include "poseidon2_constants.inc"

using State = std::array<Fp, 12>;

inline Fp sbox7(Fp x) { Fp x2 = x * x; Fp x3 = x2 * x; Fp x4 = x2 * x2; return x4 * x3; }

// External linear layer: block-circulant of M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]]
// (mod.rs:234-276): apply M4 to each 4-chunk, then add the column sums to every chunk.
inline void p2_external(State& s) {
    for (int c = 0; c < 3; c++) {
        Fp x0 = s[4 * c], x1 = s[4 * c + 1], x2 = s[4 * c + 2], x3 = s[4 * c + 3];
        Fp sum = (x0 + x1) + (x2 + x3);
        s[4 * c + 0] = sum + x0 + (x1 + x1);      // 2 x0 + 3 x1 + x2 + x3
        s[4 * c + 1] = sum + x1 + (x2 + x2);
        s[4 * c + 2] = sum + x2 + (x3 + x3);
        s[4 * c + 3] = sum + x3 + (x0 + x0);
    }
    Fp col[4];
    for (int l = 0; l < 4; l++) col[l] = s[l] + s[4 + l] + s[8 + l];
    for (int i = 0; i < 12; i++) s[i] = s[i] + col[i % 4];
}

// Internal linear layer: s_i * diag_i + sum(s)  (mod.rs:283-292).
inline void p2_internal(State& s) {
    Fp sum;
    for (int i = 0; i < 12; i++) sum += s[i];
    for (int i = 0; i < 12; i++) s[i] = s[i] * Fp::raw(P2_INTERNAL_DIAG[i]) + sum;
}

inline void poseidon2_permute(State& s) {
    p2_external(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i] + Fp::raw(P2_RC_EXT_INITIAL[12 * r + i]));
        p2_external(s);
    }
    for (int r = 0; r < 22; r++) {
        s[0] = sbox7(s[0] + Fp::raw(P2_RC_INTERNAL[r]));
        p2_internal(s);
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i] + Fp::raw(P2_RC_EXT_TERMINAL[12 * r + i]));
        p2_external(s);
    }
}

using Digest = std::array<Fp, 4>;

// Which of the reference's STARK hash configurations the LMCS and the challenger follow (air/src/config.rs):
//   H_POSEIDON2: StatefulSponge<Poseidon2, 12, 8, 4> leaves (alignment 8), TruncatedPermutation nodes, DuplexChallenger
//                (:204-223, 255-273) -- the default;
//   H_BLAKE3   : ChainingHasher<Blake3> leaves (alignment 1: state <- blake3(state || little-endian u64 of every felt of
//                the row), crates/stateful-hasher/src/chaining.rs:31-52,161-169), blake3(left || right) nodes
//                (CompressionFunctionFromHasher<Blake3, 2, 32>), SerializingChallenger64<HashChallenger<u8, Blake3, 32>>
//                (:276-307).  A 32-byte digest travels as four raw little-endian u64 in a `Digest`.
//   H_KECCAK   : SerializingStatefulSponge<StatefulSponge<KeccakF, 25, 17, 4>> leaves (the overwrite-mode sponge over the
//                canonical u64 of every felt, alignment lcm(8, 17*8)/8 = 17, crates/stateful-hasher/src/serializing_sponge.rs:72-86,
//                163-200), PaddingFreeSponge<KeccakF, 25, 17, 4> nodes (zero state, the 8 words of the two digests overwrite
//                lanes 0..8, one permutation, lanes 0..4), SerializingChallenger64<HashChallenger<u8, Keccak256Hash, 32>>
//                (:309-353).  A digest is four u64 lanes, carried raw in a `Digest`.
// One process-wide switch: the oracle is test infrastructure with a single client at a time.
//   H_RPO, H_RPX: `rpo_config` / `rpx_config` (:225-248) -- everything of H_POSEIDON2 with the permutation replaced
//                (`alg_config<P>` is generic in it, :255-273); oracle/rescue.hpp.
enum HashKind { H_POSEIDON2 = 0, H_BLAKE3 = 1, H_KECCAK = 2, H_RPO = 3, H_RPX = 4 };
inline HashKind& hash_kind() { static HashKind k = H_POSEIDON2; return k; }
inline bool byte_hash() { return hash_kind() == H_BLAKE3 || hash_kind() == H_KECCAK; }      // hash challenger instead of the duplex one
// the permutation of the algebraic configurations
inline void algebraic_permute(State& s) {
    if (hash_kind() == H_RPO) rpo_permute(s);
    else if (hash_kind() == H_RPX) rpx_permute(s);
    else poseidon2_permute(s);
}
inline size_t lmcs_alignment() { return hash_kind() == H_BLAKE3 ? 1 : hash_kind() == H_KECCAK ? 17 : 8; }
// the 32-byte hash of the byte-oriented challengers
inline std::array<uint8_t, 32> hash32(const uint8_t* p, size_t n) { return hash_kind() == H_KECCAK ? keccak::hash256(p, n) : blake3::hash(p, n); }
inline std::array<uint8_t, 32> hash32(const std::vector<uint8_t>& v) { return hash32(v.data(), v.size()); }
using KeccakState = std::array<u64, 25>;

inline void digest_to_bytes(const Fp* d, uint8_t* out) { for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(d[i].v >> (8 * k)); }
inline void bytes_to_digest(const uint8_t* b, Fp* d) {
    for (int i = 0; i < 4; i++) { u64 v = 0; for (int k = 0; k < 8; k++) v |= (u64)b[8 * i + k] << (8 * k); d[i] = Fp::raw(v); }
}
// ChainingHasher::absorb_into: state <- H(state || into_byte_stream(row)); an empty row still re-hashes the state
inline void chaining_absorb_blake3(Fp* st4, const Fp* in, size_t n) {
    std::vector<uint8_t> buf(32 + 8 * n);
    digest_to_bytes(st4, buf.data());
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 8; k++) buf[32 + 8 * i + k] = (uint8_t)(in[i].v >> (8 * k));   // canonical u64, little-endian
    auto h = blake3::hash(buf);
    bytes_to_digest(h.data(), st4);
}

// Overwrite-mode sponge absorb (crates/stateful-hasher/src/field_sponge.rs:41-59), generic in the element type, the
// width/rate and the permutation like the reference's `StatefulSponge<P, WIDTH, RATE, OUT>`: each full chunk of RATE
// overwrites state[0..RATE] then permutes; a trailing partial chunk is zero-filled to the rate boundary and permuted;
// an empty input leaves the state untouched.  The production instance is (Fp, 12, 8, Poseidon2); the reference's own
// unit vectors use a mock permutation over u64 (tests/test_reference_vectors.py pins this function on them).
template <class T, size_t WIDTH, size_t RATE, class Perm>
inline void sponge_absorb_generic(std::array<T, WIDTH>& st, const T* in, size_t n, Perm&& permute) {
    static_assert(RATE < WIDTH, "rate must leave a capacity");
    size_t i = 0;
    while (i + RATE <= n) {
        for (size_t k = 0; k < RATE; k++) st[k] = in[i + k];
        permute(st);
        i += RATE;
    }
    if (i < n) {
        size_t rem = n - i;
        for (size_t k = 0; k < rem; k++) st[k] = in[i + k];
        for (size_t k = rem; k < RATE; k++) st[k] = T();
        permute(st);
    }
}
inline void sponge_absorb(State& st, const Fp* in, size_t n) {
    if (hash_kind() == H_BLAKE3) { chaining_absorb_blake3(st.data(), in, n); return; }
    sponge_absorb_generic<Fp, 12, 8>(st, in, n, [](State& s) { algebraic_permute(s); });
}
inline Digest sponge_squeeze(const State& st) { return Digest{st[0], st[1], st[2], st[3]}; }
// the Keccak leaf sponge: 25 u64 lanes, rate 17, the row's felts as canonical u64
inline void sponge_absorb(KeccakState& st, const Fp* in, size_t n) {
    std::vector<u64> w(n);
    for (size_t i = 0; i < n; i++) w[i] = in[i].v;
    sponge_absorb_generic<u64, 25, 17>(st, w.data(), n, [](KeccakState& s) { keccak::permute(s); });
}
inline Digest sponge_squeeze(const KeccakState& st) { return Digest{Fp::raw(st[0]), Fp::raw(st[1]), Fp::raw(st[2]), Fp::raw(st[3])}; }

// 2-to-1 compression = p3 TruncatedPermutation<_, 2, 4, 12>: perm([l | r | 0000])[0..4]
// (air/src/config.rs:217; equality with Poseidon2::merge shown by poseidon2/test.rs:208-230).
inline Digest compress2(const Digest& l, const Digest& r) {
    if (hash_kind() == H_BLAKE3) {
        uint8_t buf[64];
        digest_to_bytes(l.data(), buf); digest_to_bytes(r.data(), buf + 32);
        auto h = blake3::hash(buf, 64);
        Digest d; bytes_to_digest(h.data(), d.data());
        return d;
    }
    if (hash_kind() == H_KECCAK) {
        KeccakState k{};
        for (int i = 0; i < 4; i++) { k[i] = l[i].v; k[4 + i] = r[i].v; }
        keccak::permute(k);
        return sponge_squeeze(k);
    }
    State s;
    for (int i = 0; i < 4; i++) { s[i] = l[i]; s[4 + i] = r[i]; s[8 + i] = Fp(); }
    algebraic_permute(s);
    return Digest{s[0], s[1], s[2], s[3]};
}

}  // namespace orc
