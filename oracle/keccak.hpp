// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Keccak-f[1600] and Keccak-256, written from the published Keccak specification (FIPS 202 section 3: the step mappings
// theta, rho, pi, chi, iota over A[x][y]; rotation offsets from the (x, y) walk and round constants from the LFSR
// rc(t), both computed here rather than tabulated), independently of the product's flattened, table-driven
// csrc/keccak.cuh.  tests/test_keccak.py pins both on hashlib's SHA3-256 (same permutation and sponge, padding byte 0x06)
// and on the Keccak-256 of the empty string; the reference's `KeccakF` / `Keccak256Hash` re-export p3_keccak 0.6.2
// (crates/crypto/src/hash/keccak/mod.rs:18), un-vendored.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

namespace keccak {
using u64 = uint64_t;

inline u64 rotl(u64 x, unsigned n) { n &= 63; return n ? (x << n) | (x >> (64 - n)) : x; }

struct Tables {
    unsigned rot[5][5];
    u64 rc[24];
    Tables() {
        for (auto& r : rot) for (auto& v : r) v = 0;
        unsigned x = 1, y = 0;
        for (unsigned t = 0; t < 24; t++) {
            rot[x][y] = ((t + 1) * (t + 2) / 2) % 64;
            unsigned nx = y, ny = (2 * x + 3 * y) % 5;
            x = nx; y = ny;
        }
        uint8_t lfsr = 1;
        for (int round = 0; round < 24; round++) {
            u64 c = 0;
            for (int j = 0; j <= 6; j++) {
                if (lfsr & 1) c |= u64(1) << ((1u << j) - 1);
                lfsr = (uint8_t)((lfsr << 1) ^ ((lfsr & 0x80) ? 0x71 : 0));   // x^8 + x^6 + x^5 + x^4 + 1
            }
            rc[round] = c;
        }
    }
};
inline const Tables& tables() { static const Tables t; return t; }

// lanes: st[x + 5 y]
inline void permute(std::array<u64, 25>& st) {
    const Tables& T = tables();
    u64 A[5][5], B[5][5], C[5], D[5];
    for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) A[x][y] = st[x + 5 * y];
    for (int round = 0; round < 24; round++) {
        for (int x = 0; x < 5; x++) C[x] = A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl(C[(x + 1) % 5], 1);
        for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) A[x][y] ^= D[x];
        for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) B[y][(2 * x + 3 * y) % 5] = rotl(A[x][y], T.rot[x][y]);
        for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) A[x][y] = B[x][y] ^ (~B[(x + 1) % 5][y] & B[(x + 2) % 5][y]);
        A[0][0] ^= T.rc[round];
    }
    for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) st[x + 5 * y] = A[x][y];
}

// the sponge with rate 136 bytes and a 32-byte output; `pad` = 0x01 (Keccak-256) or 0x06 (SHA3-256)
inline std::array<uint8_t, 32> hash256(const uint8_t* p, size_t n, uint8_t pad = 0x01) {
    uint8_t block[200];
    std::array<u64, 25> st{};
    auto absorb_block = [&] {
        for (int i = 0; i < 17; i++) { u64 w = 0; for (int k = 0; k < 8; k++) w |= (u64)block[8 * i + k] << (8 * k); st[i] ^= w; }
        permute(st);
    };
    size_t off = 0;
    while (n - off >= 136) { memcpy(block, p + off, 136); absorb_block(); off += 136; }
    memset(block, 0, sizeof block);
    memcpy(block, p + off, n - off);
    block[n - off] ^= pad;
    block[135] ^= 0x80;
    absorb_block();
    std::array<uint8_t, 32> out;
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(st[i] >> (8 * k));
    return out;
}
inline std::array<uint8_t, 32> hash256(const std::vector<uint8_t>& v, uint8_t pad = 0x01) { return hash256(v.data(), v.size(), pad); }

}  // namespace keccak
