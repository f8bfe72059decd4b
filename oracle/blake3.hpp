// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// BLAKE3 hash (32-byte output) over byte strings, written from the published specification independently of the
// product's word-streaming csrc/blake3.cuh; tests/test_blake3.py pins both against the `blake3` package (bindings of the
// official crate, which is what the reference's `Blake3Hasher` = p3_blake3::Blake3 wraps, crates/crypto/src/hash/blake/mod.rs:16).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {
namespace blake3 {

static const uint32_t IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
static const int PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
inline void g(uint32_t* s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx; s[d] = rotr(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = rotr(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 7);
}
// full 16-word output of the compression function
inline std::array<uint32_t, 16> compress(const uint32_t cv[8], const uint8_t block[64], uint64_t counter, uint32_t block_len, uint32_t flags) {
    uint32_t m[16];
    for (int i = 0; i < 16; i++) m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) | ((uint32_t)block[4 * i + 3] << 24);
    uint32_t s[16] = {cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], cv[6], cv[7], IV[0], IV[1], IV[2], IV[3],
                      (uint32_t)counter, (uint32_t)(counter >> 32), block_len, flags};
    for (int r = 0; r < 7; r++) {
        g(s, 0, 4, 8, 12, m[0], m[1]); g(s, 1, 5, 9, 13, m[2], m[3]); g(s, 2, 6, 10, 14, m[4], m[5]); g(s, 3, 7, 11, 15, m[6], m[7]);
        g(s, 0, 5, 10, 15, m[8], m[9]); g(s, 1, 6, 11, 12, m[10], m[11]); g(s, 2, 7, 8, 13, m[12], m[13]); g(s, 3, 4, 9, 14, m[14], m[15]);
        uint32_t t[16];
        for (int i = 0; i < 16; i++) t[i] = m[PERM[i]];
        memcpy(m, t, sizeof m);
    }
    std::array<uint32_t, 16> out;
    for (int i = 0; i < 8; i++) { out[i] = s[i] ^ s[i + 8]; out[i + 8] = s[i + 8] ^ cv[i]; }
    return out;
}

// chaining value of chunk `index` (<= 1024 bytes), or its root output when it is the whole input
inline std::array<uint32_t, 8> chunk_cv(const uint8_t* p, size_t len, uint64_t index, bool root) {
    uint32_t cv[8];
    memcpy(cv, IV, sizeof cv);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        uint8_t block[64] = {0};
        size_t bl = std::min<size_t>(64, len - 64 * b);
        if (len) memcpy(block, p + 64 * b, bl); else bl = 0;
        uint32_t flags = (b == 0 ? CHUNK_START : 0) | (b + 1 == nblocks ? CHUNK_END | (root ? ROOT : 0) : 0);
        auto o = compress(cv, block, index, (uint32_t)bl, flags);
        for (int i = 0; i < 8; i++) cv[i] = o[i];
    }
    std::array<uint32_t, 8> r;
    for (int i = 0; i < 8; i++) r[i] = cv[i];
    return r;
}
inline std::array<uint32_t, 8> parent_cv(const std::array<uint32_t, 8>& l, const std::array<uint32_t, 8>& r, bool root) {
    uint8_t block[64];
    for (int i = 0; i < 8; i++) for (int k = 0; k < 4; k++) { block[4 * i + k] = (uint8_t)(l[i] >> (8 * k)); block[32 + 4 * i + k] = (uint8_t)(r[i] >> (8 * k)); }
    auto o = compress(IV, block, 0, 64, PARENT | (root ? ROOT : 0));
    std::array<uint32_t, 8> out;
    for (int i = 0; i < 8; i++) out[i] = o[i];
    return out;
}
// subtree over chunks [first, first + n) of the input: the left child takes the largest power of two < n chunks
inline std::array<uint32_t, 8> subtree(const uint8_t* p, size_t len, uint64_t first, bool root) {
    size_t n = (len + 1023) / 1024;
    if (n <= 1) return chunk_cv(p, len, first, root);
    size_t left = 1;
    while (left * 2 < n) left *= 2;
    auto l = subtree(p, left * 1024, first, false);
    auto r = subtree(p + left * 1024, len - left * 1024, first + left, false);
    return parent_cv(l, r, root);
}
inline std::array<uint8_t, 32> hash(const uint8_t* p, size_t len) {
    auto w = subtree(p, len, 0, true);
    std::array<uint8_t, 32> out;
    for (int i = 0; i < 8; i++) for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(w[i] >> (8 * k));
    return out;
}
inline std::array<uint8_t, 32> hash(const std::vector<uint8_t>& v) { return hash(v.data(), v.size()); }

}  // namespace blake3
}  // namespace orc
