// Host-side Fiat-Shamir pieces of the product: duplex challenger, prover transcript streams, and
// Merkle index bookkeeping.  These are the sequential, byte-sized steps the reference also keeps on
// the host; the only heavy part (proof-of-work search) is delegated to the GPU by the session.
//
//   DuplexChallenger semantics  : crates/lib/core/asm/stark/random_coin.masm:103-115 (length tag),
//                                 :181-210 (observe), :272-296 (lazy flush), :128-135 (sample order),
//                                 :151-166 (sample_bits)
//   ProverChannel               : crates/stark-transcript/src/prover.rs:116-145
//   TreeIndices / siblings      : crates/lifted-stark/src/lmcs/tree_indices.rs:34-45,113-126,197-…
#pragma once
#include "poseidon2.cuh"
#include "blake3.cuh"
#include "keccak.cuh"
#include "rescue.cuh"
#include <algorithm>
#include <utility>
#include <vector>

namespace hostfs {
using gl::u64;
using gl::u32;
using gl::E2;

// `hashed == true` turns the same object into p3's `SerializingChallenger64<Felt, HashChallenger<u8, H, 32>>`, the
// challenger of the Blake3_256 (H = Blake3) and Keccak (H = Keccak-256) configurations (air/src/config.rs:292-293,304-305,
// 335-336,350-351; p3-challenger 0.6.2 is not vendored, so
// this restates the published crate -- parity unpinned like the duplex case):
//   observe(felt) appends its canonical u64 as 8 little-endian bytes to the input buffer and clears the output buffer;
//   a digest is observed as its 32 bytes; a sampled byte pops from the BACK of the output buffer, which is refilled by
//   out = blake3(input buffer), input buffer <- out (chaining); sample() = u64::from_le_bytes(8 bytes), redrawn while >= p;
//   sample_bits(b) = low b bits of u64::from_le_bytes(8 bytes), no rejection.
struct Duplex {
    u64 st[12];
    u64 in[8];
    u32 in_len = 0, out_len = 0;
    int perm = 0;             // duplex mode: the permutation -- 0 Poseidon2, 3 RPO, 4 RPX (`alg_config<P>`, air/src/config.rs:255-273)
    bool hashed = false;
    bool keccak = false;      // hashed only: H = Keccak-256 (the Keccak configuration, air/src/config.rs:335-336) instead of Blake3
    std::vector<uint8_t> bin, bout;

    void observe_byte(uint8_t b) { bout.clear(); bin.push_back(b); }
    uint8_t sample_byte() {
        if (bout.empty() && keccak) {
            kk::Hash256 h; h.init();           // every observation is 8 or 32 bytes: the buffer is whole 64-bit words
            for (size_t i = 0; i + 8 <= bin.size(); i += 8) { u64 w = 0; for (int k = 0; k < 8; k++) w |= (u64)bin[i + k] << (8 * k); h.push64(w); }
            u64 o[4]; h.finish(o);
            bout.resize(32);
            for (int i = 0; i < 32; i++) bout[i] = (uint8_t)(o[i / 8] >> (8 * (i % 8)));
            bin = bout;
        } else if (bout.empty()) {
            b3::Hasher h; h.init();
            for (size_t i = 0; i + 4 <= bin.size(); i += 4) h.push((u32)bin[i] | ((u32)bin[i + 1] << 8) | ((u32)bin[i + 2] << 16) | ((u32)bin[i + 3] << 24));
            u32 o[8]; h.finish(o);             // every observation is 8 or 32 bytes: the buffer is whole words
            bout.resize(32);
            for (int i = 0; i < 32; i++) bout[i] = (uint8_t)(o[i / 4] >> (8 * (i % 4)));
            bin = bout;
        }
        uint8_t b = bout.back(); bout.pop_back();
        return b;
    }
    u64 sample_u64_bytes() { u64 v = 0; for (int k = 0; k < 8; k++) v |= (u64)sample_byte() << (8 * k); return v; }
    void observe_digest(const u64* d) {
        if (hashed) { for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) observe_byte((uint8_t)(d[i] >> (8 * k))); return; }
        for (int i = 0; i < 4; i++) observe(d[i]);
    }

    void duplex() {
        if (in_len) {
            for (u32 i = 0; i < 8; i++) st[i] = i < in_len ? in[i] : 0;
            st[8] = gl::add(st[8], in_len);
            in_len = 0;
        }
        if (perm == 3) rsc::rpo_permute(st);
        else if (perm == 4) rsc::rpx_permute(st);
        else p2::permute(st);
        out_len = 8;
    }
    void observe(u64 x) {
        if (hashed) { for (int k = 0; k < 8; k++) observe_byte((uint8_t)(x >> (8 * k))); return; }
        out_len = 0;
        in[in_len++] = x;
        if (in_len == 8) duplex();
    }
    u64 sample() {
        if (hashed) { for (;;) { u64 v = sample_u64_bytes(); if (v < gl::P) return v; } }
        if (in_len || !out_len) duplex();
        return st[--out_len];
    }
    E2 sample_ext() { u64 a = sample(); u64 b = sample(); return gl::e2(a, b); }
    u64 sample_bits(u32 bits) { return (hashed ? sample_u64_bytes() : sample()) & ((1ull << bits) - 1); }
};

struct Transcript {
    Duplex ch;
    std::vector<u64> fields;
    std::vector<u64> commitments;   // 4 per digest
    void send_field(u64 x) { fields.push_back(x); ch.observe(x); }
    void send_ext(E2 x) { send_field(x.a); send_field(x.b); }
    void send_commitment(const u64* d) { for (int i = 0; i < 4; i++) commitments.push_back(d[i]); ch.observe_digest(d); }
    void hint_field(u64 x) { fields.push_back(x); }
    void hint_commitment(const u64* d) { commitments.insert(commitments.end(), d, d + 4); }
};

struct Indices {
    std::vector<size_t> idx;   // sorted, unique
    u32 depth = 0;
    static Indices make(std::vector<size_t> v, u32 depth) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        Indices r; r.idx = std::move(v); r.depth = depth;
        return r;
    }
    Indices folded(u32 target) const {
        std::vector<size_t> v = idx;
        size_t mask = ((size_t)1 << target) - 1;
        for (auto& x : v) x &= mask;
        return make(std::move(v), target);
    }
};

// (depth, position) of the missing sibling digests, bottom-to-top, left-to-right.
inline std::vector<std::pair<u32, size_t>> missing_siblings(const Indices& ti) {
    std::vector<std::pair<u32, size_t>> out;
    std::vector<size_t> layer = ti.idx, parents;
    for (u32 d = ti.depth; d > 0; d--) {
        parents.clear();
        size_t k = 0;
        while (k < layer.size()) {
            size_t node = layer[k];
            bool has_sibling = (k + 1 < layer.size()) && layer[k + 1] == (node ^ 1);
            if (parents.empty() || parents.back() != (node >> 1)) parents.push_back(node >> 1);
            if (has_sibling) k += 2;
            else { out.emplace_back(d, node ^ 1); k += 1; }
        }
        layer.swap(parents);
    }
    return out;
}

}  // namespace hostfs
