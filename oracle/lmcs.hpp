// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Lifted Matrix Commitment Scheme (LMCS): Merkle tree over rows of several matrices of
// power-of-two heights, restated from crates/lifted-stark/src/lmcs/:
//   lifted_tree.rs:202-284  build_with_alignment (leaf states -> squeeze in domain order -> layers)
//   lifted_tree.rs:363-417  build_leaf_states_upsampled (state duplication between heights)
//   lifted_tree.rs:472-511  compress_uniform
//   lifted_tree.rs:155-180, 326-341  prove_batch / collect_rows (hint layout)
//   tree_indices.rs:34-45, 113-126, 197-…  TreeIndices / fold / MissingSiblingsIter
//   config.rs:152-…, merkle_witness.rs    open_batch (verifier side)
#pragma once
#include "ntt.hpp"
#include "transcript.hpp"
#include "poseidon2_x8.hpp"
#include <algorithm>
#include <map>
#include <type_traits>

namespace orc {

inline size_t aligned_len(size_t w, size_t alignment) { return (w + alignment - 1) / alignment * alignment; }

struct LmcsTree {
    // Stored matrices: rows in BIT-REVERSED domain order, ascending heights.
    std::vector<Matrix> leaves;
    // layers[d] has 2^d digests, layers[0][0] = root; last layer = leaf digests in DOMAIN order.
    std::vector<std::vector<Digest>> layers;
    size_t alignment = 1;

    size_t height() const { return leaves.back().height; }
    unsigned depth() const { return log2_strict(height()); }
    const Digest& root() const { return layers[0][0]; }

    // leaf states per height group (nearest-neighbour duplication between heights), squeezed in domain order; `St` is the
    // state of the configured stateful hasher (12 felts: Poseidon2 sponge / Blake3 chaining state in the first 4; 25 lanes: Keccak)
    template <class St>
    static std::vector<Digest> leaf_digests(const std::vector<Matrix>& leaves) {
        size_t H = leaves.back().height;
        std::vector<St> states(H, St{});
        size_t active = leaves.front().height;
        for (const Matrix& m : leaves) {
            size_t h = m.height;
            assert(h >= active && (h & (h - 1)) == 0);
            if (h > active) {   // nearest-neighbour duplication of the running states
                size_t f = h / active;
                for (size_t i = active; i-- > 0;)
                    for (size_t k = 0; k < f; k++) states[i * f + k] = states[i];
            }
            bool done = false;
#if ORC_HAVE_X8
            if constexpr (std::is_same_v<St, State>) {
                if (h >= 8 && x8_available() && hash_kind() == H_POSEIDON2) {   // 8 leaves per AVX-512 permutation
#pragma omp parallel for schedule(static) if (h > 256)
                    for (size_t r8 = 0; r8 < h / 8; r8++) sponge_absorb_x8(&states[8 * r8], m.row(8 * r8), m.width, m.width);
                    done = true;
                }
            }
#endif
            if (!done) {
#pragma omp parallel for schedule(static) if (h > 256)
                for (size_t r = 0; r < h; r++) sponge_absorb(states[r], m.row(r), m.width);
            }
            active = h;
        }
        unsigned lg = log2_strict(H);
        std::vector<Digest> layer(H);
        for (size_t i = 0; i < H; i++) layer[i] = sponge_squeeze(states[reverse_bits64(i, lg)]);
        return layer;
    }

    static LmcsTree build(std::vector<Matrix> mats, size_t alignment) {
        LmcsTree t;
        t.alignment = alignment;
        t.leaves = std::move(mats);
        size_t H = t.leaves.back().height;
        unsigned lg = log2_strict(H);
        std::vector<Digest> layer = hash_kind() == H_KECCAK ? leaf_digests<KeccakState>(t.leaves) : leaf_digests<State>(t.leaves);
        t.layers.assign(lg + 1, {});
        t.layers[lg] = std::move(layer);
        for (unsigned d = lg; d-- > 0;) {
            const std::vector<Digest>& prev = t.layers[d + 1];
            std::vector<Digest> next(prev.size() / 2);
#if ORC_HAVE_X8
            if (next.size() >= 8 && x8_available() && hash_kind() == H_POSEIDON2) {
#pragma omp parallel for schedule(static) if (next.size() > 256)
                for (size_t i8 = 0; i8 < next.size() / 8; i8++) compress2_x8(&prev[16 * i8], &next[8 * i8]);
            } else
#endif
            {
#pragma omp parallel for schedule(static) if (next.size() > 256)
                for (size_t i = 0; i < next.size(); i++) next[i] = compress2(prev[2 * i], prev[2 * i + 1]);
            }
            t.layers[d] = std::move(next);
        }
        return t;
    }

    // Rows opened for domain index `idx`: matrix j contributes physical row
    // bitrev(idx) >> log(H / h_j), zero-padded to the alignment (lifted_tree.rs:326-341).
    void write_rows(size_t idx, ProverTranscript& ch) const {
        unsigned lg = depth();
        size_t br = reverse_bits64(idx, lg);
        for (const Matrix& m : leaves) {
            unsigned sc = lg - log2_strict(m.height);
            const Fp* row = m.row(br >> sc);
            for (size_t c = 0; c < m.width; c++) ch.hint_field(row[c]);
            for (size_t c = m.width; c < aligned_len(m.width, alignment); c++) ch.hint_field(Fp());
        }
    }
};

// Sorted, de-duplicated index set at a given depth (tree_indices.rs:34-45).
struct TreeIndices {
    std::vector<size_t> idx;
    unsigned depth = 0;
    static TreeIndices make(std::vector<size_t> v, unsigned depth) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        TreeIndices t; t.idx = std::move(v); t.depth = depth;
        return t;
    }
    // tree_indices.rs:113-126: keep the low `target` bits.
    TreeIndices fold_to_depth(unsigned target) const {
        assert(target <= depth);
        std::vector<size_t> v = idx;
        size_t mask = (size_t(1) << target) - 1;
        for (auto& x : v) x &= mask;
        return make(std::move(v), target);
    }
    void shrink_depth(unsigned shift) { *this = fold_to_depth(depth > shift ? depth - shift : 0); }
};

// (depth, position) of the sibling digests a verifier is missing, bottom-to-top, left-to-right
// (tree_indices.rs MissingSiblingsIter).
inline std::vector<std::pair<unsigned, size_t>> missing_siblings(const TreeIndices& ti) {
    std::vector<std::pair<unsigned, size_t>> out;
    std::vector<size_t> cur = ti.idx;
    for (unsigned d = ti.depth; d > 0; d--) {
        std::vector<size_t> next;
        for (size_t k = 0; k < cur.size(); k++) {
            size_t node = cur[k], sib = node ^ 1;
            bool present = (k + 1 < cur.size() && cur[k + 1] == sib);
            if (next.empty() || next.back() != (node >> 1)) next.push_back(node >> 1);
            if (present) k++; else out.push_back({d, sib});
        }
        cur = std::move(next);
    }
    return out;
}

// lifted_tree.rs:155-180: per index the aligned rows, then the missing siblings.
inline void lmcs_prove_batch(const LmcsTree& t, const TreeIndices& ti, ProverTranscript& ch) {
    assert(ti.depth == t.depth());
    for (size_t i : ti.idx) t.write_rows(i, ch);
    for (auto& ds : missing_siblings(ti)) ch.hint_commitment(t.layers[ds.first][ds.second]);
}
inline void lmcs_prove_lifted_batch(const LmcsTree& t, const TreeIndices& q, ProverTranscript& ch) {
    lmcs_prove_batch(t, q.fold_to_depth(t.depth()), ch);
}

// Verifier side (config.rs open_batch + merkle_witness.rs): read rows for each index, hash them
// (one sponge absorb per matrix row, in order), rebuild the root from the streamed siblings.
// `widths` are the (aligned) row widths per matrix.  Returns rows keyed by tree index.
inline std::map<size_t, std::vector<Fp>> lmcs_open_batch(const Digest& root, const std::vector<size_t>& widths,
                                                         const TreeIndices& ti, VerifierTranscript& ch) {
    if (ti.idx.empty()) throw std::runtime_error("lmcs: empty index set");
    std::map<size_t, std::vector<Fp>> rows;
    std::vector<Digest> cur_hash;
    std::vector<size_t> cur = ti.idx;
    size_t total = 0;
    for (size_t w : widths) total += w;
    for (size_t i : ti.idx) {
        std::vector<Fp> r(total);
        for (auto& x : r) x = ch.hint_field();
        size_t off = 0;
        if (hash_kind() == H_KECCAK) {
            KeccakState st{};
            for (size_t w : widths) { sponge_absorb(st, r.data() + off, w); off += w; }
            cur_hash.push_back(sponge_squeeze(st));
        } else {
            State st{};
            for (size_t w : widths) { sponge_absorb(st, r.data() + off, w); off += w; }
            cur_hash.push_back(sponge_squeeze(st));
        }
        rows[i] = std::move(r);
    }
    for (unsigned d = ti.depth; d > 0; d--) {
        std::vector<size_t> next;
        std::vector<Digest> next_hash;
        for (size_t k = 0; k < cur.size(); k++) {
            size_t node = cur[k], sib = node ^ 1;
            bool present = (k + 1 < cur.size() && cur[k + 1] == sib);
            Digest sh = present ? cur_hash[k + 1] : ch.hint_commitment();
            Digest parent = (node & 1) ? compress2(sh, cur_hash[k]) : compress2(cur_hash[k], sh);
            next.push_back(node >> 1);
            next_hash.push_back(parent);
            if (present) k++;
        }
        cur = std::move(next);
        cur_hash = std::move(next_hash);
    }
    if (cur_hash.size() != 1 || cur_hash[0] != root) throw std::runtime_error("lmcs: root mismatch");
    return rows;
}
inline std::map<size_t, std::vector<Fp>> lmcs_open_lifted_batch(const Digest& root, const std::vector<size_t>& widths,
                                                                const TreeIndices& q, unsigned tree_log_height,
                                                                VerifierTranscript& ch) {
    TreeIndices leafs = q.fold_to_depth(tree_log_height);
    auto by_leaf = lmcs_open_batch(root, widths, leafs, ch);
    std::map<size_t, std::vector<Fp>> out;
    size_t mask = (size_t(1) << tree_log_height) - 1;
    for (size_t i : q.idx) out[i] = by_leaf.at(i & mask);
    return out;
}

}  // namespace orc
