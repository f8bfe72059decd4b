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
#include <algorithm>
#include <utility>
#include <vector>

namespace hostfs {
using gl::u64;
using gl::u32;
using gl::E2;

struct Duplex {
    u64 st[12];
    u64 in[8];
    u32 in_len = 0, out_len = 0;

    void duplex() {
        if (in_len) {
            for (u32 i = 0; i < 8; i++) st[i] = i < in_len ? in[i] : 0;
            st[8] = gl::add(st[8], in_len);
            in_len = 0;
        }
        p2::permute(st);
        out_len = 8;
    }
    void observe(u64 x) {
        out_len = 0;
        in[in_len++] = x;
        if (in_len == 8) duplex();
    }
    u64 sample() {
        if (in_len || !out_len) duplex();
        return st[--out_len];
    }
    E2 sample_ext() { u64 a = sample(); u64 b = sample(); return gl::e2(a, b); }
    u64 sample_bits(u32 bits) { return sample() & ((1ull << bits) - 1); }
};

struct Transcript {
    Duplex ch;
    std::vector<u64> fields;
    std::vector<u64> commitments;   // 4 per digest
    void send_field(u64 x) { fields.push_back(x); ch.observe(x); }
    void send_ext(E2 x) { send_field(x.a); send_field(x.b); }
    void send_commitment(const u64* d) { for (int i = 0; i < 4; i++) { commitments.push_back(d[i]); ch.observe(d[i]); } }
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
