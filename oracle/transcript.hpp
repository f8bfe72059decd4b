// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Duplex challenger + prover/verifier transcript channel.
//
// The challenger is p3-challenger 0.6.2 `DuplexChallenger<Felt, Poseidon2, 12, 8>` (un-vendored).
// Its behaviour is restated from the reference's in-tree MASM verifier, which must agree with
// the Rust prover for recursive verification to pass (miden-vm/tests/integration/prove_verify.rs:77):
//   crates/lib/core/asm/stark/random_coin.masm
//     :103-115  add_absorb_length_tag  -- every absorbing permutation adds the number of
//               absorbed elements to state[8]; squeeze-only permutations add nothing
//     :181-210  observe_felt           -- observe clears pending output, appends at
//               rate[input_len], permutes when 8 are buffered
//     :272-296  flush_buffer           -- lazily at the next sample: zero rate[input_len..8],
//               add tag, permute
//     :128-135  sample_felt            -- pops rate[7], rate[6], ... (output_len counter)
//     :151-166  sample_bits            -- low `bits` bits of the low 32-bit limb of one felt
//     :929-966  check_pow              -- observe(witness) then sample_bits(bits) == 0;
//               bits == 0: witness is zero and the sponge is untouched
// PARITY UNPINNED: p3's `grind` search order (parallel find_any under `concurrent`) is not
// pinned in-tree; this oracle returns the smallest valid witness (= sequential search order).
// Channel semantics: crates/stark-transcript/src/prover.rs:116-145 (send = record + observe,
// hint = record only, grind appends the witness to `fields`).
#pragma once
#include "poseidon2.hpp"
#include <stdexcept>

namespace orc {

// In H_BLAKE3 / H_KECCAK mode (H = Blake3 / Keccak-256, air/src/config.rs:335-336,350-351) the same struct is p3's `SerializingChallenger64<Felt, HashChallenger<u8, Blake3, 32>>`
// (air/src/config.rs:292-293,304-305; p3-challenger 0.6.2, un-vendored -- PARITY UNPINNED, restated from the published crate):
//   HashChallenger: observe(byte) clears the output buffer and appends to the input buffer; sample(): if the output buffer
//   is empty, flush = { out = H(input buffer); output buffer = out; input buffer = out (chaining) }, then POP FROM THE BACK;
//   SerializingChallenger64: observe(felt) = its canonical u64 as 8 little-endian bytes; observe(digest) = its 32 bytes;
//   sample() = u64::from_le_bytes(8 sampled bytes), rejected and redrawn while >= p; sample_bits(b) = low b bits of
//   u64::from_le_bytes(8 sampled bytes) (no rejection); grinding as in the duplex case (observe witness, sample_bits == 0).
struct Challenger {
    State st{};
    Fp in_buf[8];
    int in_len = 0;
    int out_len = 0;  // number of unread rate elements; next sample returns st[out_len-1]
    std::vector<uint8_t> bin, bout;   // H_BLAKE3: HashChallenger::input_buffer / output_buffer

    static Challenger from_bytes(const uint8_t* p, size_t n) { Challenger c; c.bin.assign(p, p + n); return c; }
    void observe_byte(uint8_t b) { bout.clear(); bin.push_back(b); }
    uint8_t sample_byte() {
        if (bout.empty()) {
            auto h = hash32(bin);
            bout.assign(h.begin(), h.end());
            bin.assign(h.begin(), h.end());
        }
        uint8_t b = bout.back(); bout.pop_back();
        return b;
    }
    u64 sample_u64_bytes() { u64 v = 0; for (int k = 0; k < 8; k++) v |= (u64)sample_byte() << (8 * k); return v; }

    static Challenger with_capacity(const Digest& cap) {
        Challenger c;
        for (int i = 0; i < 4; i++) c.st[8 + i] = cap[i];  // air/src/config.rs:264-271
        return c;
    }
    void duplex() {
        if (in_len > 0) {
            for (int i = 0; i < in_len; i++) st[i] = in_buf[i];
            for (int i = in_len; i < 8; i++) st[i] = Fp();
            st[8] = st[8] + Fp::raw((u64)in_len);
            in_len = 0;
        }
        algebraic_permute(st);
        out_len = 8;
    }
    void observe(Fp x) {
        if (byte_hash()) { for (int k = 0; k < 8; k++) observe_byte((uint8_t)(x.v >> (8 * k))); return; }
        out_len = 0;
        in_buf[in_len++] = x;
        if (in_len == 8) duplex();
    }
    void observe_digest(const Digest& d) {
        if (byte_hash()) { uint8_t b[32]; digest_to_bytes(d.data(), b); for (int i = 0; i < 32; i++) observe_byte(b[i]); return; }
        for (int i = 0; i < 4; i++) observe(d[i]);
    }
    Fp sample() {
        if (byte_hash()) { for (;;) { u64 v = sample_u64_bytes(); if (v < P) return Fp::raw(v); } }
        if (in_len > 0 || out_len == 0) duplex();
        return st[--out_len];
    }
    Ef sample_ext() { Fp a = sample(); Fp b = sample(); return Ef(a, b); }  // channel.rs:54-56
    u64 sample_bits(unsigned bits) {
        if (byte_hash()) return sample_u64_bytes() & ((u64(1) << bits) - 1);
        u64 v = sample().v;
        return v & ((u64(1) << bits) - 1);
    }
    bool check_witness(unsigned bits, Fp w) const {
        if (bits == 0) return true;
        Challenger c = *this;
        c.observe(w);
        return c.sample_bits(bits) == 0;
    }
    // Returns the witness and leaves the challenger in the post-check state.
    Fp grind(unsigned bits) {
        if (bits == 0) return Fp();
        for (u64 w = 0;; w++) {
            if (check_witness(bits, Fp::raw(w))) {
                observe(Fp::raw(w));
                u64 b = sample_bits(bits);
                (void)b;
                return Fp::raw(w);
            }
        }
    }
};

struct ProverTranscript {
    Challenger ch;
    std::vector<Fp> fields;
    std::vector<Digest> commitments;
    void send_field(Fp x) { fields.push_back(x); ch.observe(x); }
    void send_ext(Ef x) { send_field(x.a); send_field(x.b); }
    void send_commitment(const Digest& d) { commitments.push_back(d); ch.observe_digest(d); }
    void hint_field(Fp x) { fields.push_back(x); }
    void hint_commitment(const Digest& d) { commitments.push_back(d); }
    Fp grind(unsigned bits) { Fp w = ch.grind(bits); fields.push_back(w); return w; }
    Ef sample_ext() { return ch.sample_ext(); }
    u64 sample_bits(unsigned b) { return ch.sample_bits(b); }
};

struct VerifierTranscript {
    Challenger ch;
    const Fp* fields; size_t n_fields; size_t fpos = 0;
    const Digest* commitments; size_t n_commitments; size_t cpos = 0;
    Fp next_field() { if (fpos >= n_fields) throw std::runtime_error("transcript: out of fields"); return fields[fpos++]; }
    Digest next_commitment() { if (cpos >= n_commitments) throw std::runtime_error("transcript: out of commitments"); return commitments[cpos++]; }
    Fp receive_field() { Fp x = next_field(); ch.observe(x); return x; }
    Ef receive_ext() { Fp a = receive_field(); Fp b = receive_field(); return Ef(a, b); }
    Digest receive_commitment() { Digest d = next_commitment(); ch.observe_digest(d); return d; }
    Fp hint_field() { return next_field(); }
    Digest hint_commitment() { return next_commitment(); }
    void grind(unsigned bits) {
        Fp w = next_field();
        if (bits == 0) { if (!w.is_zero()) throw std::runtime_error("pow: nonzero witness for 0 bits"); return; }
        ch.observe(w);
        if (ch.sample_bits(bits) != 0) throw std::runtime_error("pow: invalid witness");
    }
    Ef sample_ext() { return ch.sample_ext(); }
    u64 sample_bits(unsigned b) { return ch.sample_bits(b); }
    bool is_empty() const { return fpos == n_fields && cpos == n_commitments; }
};

}  // namespace orc
