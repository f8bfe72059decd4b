// BLAKE3 (hash mode, 32-byte output) for host and device code of the product: the STARK hash of the reference's
// `HashFunction::Blake3_256` configuration (air/src/config.rs:276-307: `ChainingHasher<Blake3Hasher>` leaves,
// `CompressionFunctionFromHasher<Blake3Hasher, 2, 32>` nodes, `SerializingChallenger64<Felt, HashChallenger<u8,
// Blake3Hasher, 32>>` transcript; Blake3Hasher = p3_blake3::Blake3 = the `blake3` crate, crates/crypto/src/hash/blake/mod.rs:16).
// The algorithm is the published BLAKE3 specification (compression function, chunk chaining, binary tree of chunk
// chaining values); tests/test_blake3.py pins this file against the `blake3` package (the official crate's bindings)
// through tests/cpp and the C ABI.  Inputs here are always whole 32-bit words (digests and little-endian u64 field
// elements), so the streaming state takes words, not bytes.
#pragma once
#include "gl.cuh"

namespace b3 {
using gl::u32;
using gl::u64;

static constexpr u32 CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8;

GL_HD u32 rotr(u32 x, int n) { return (x >> n) | (x << (32 - n)); }
GL_HD u32 bswap(u32 x) { return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24); }
GL_HD u32 iv(int i) {
    constexpr u32 IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    return IV[i];
}

#define B3_G(a, b, c, d, mx, my) \
    do { a = a + b + (mx); d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12); a = a + b + (my); d = rotr(d ^ a, 8); c = c + d; b = rotr(b ^ c, 7); } while (0)

// One compression; out[0..8) = the new chaining value (or the first 32 output bytes when `flags` has ROOT).
// The message schedule is written out per round (permutation 2,6,3,10,7,0,4,13,1,11,12,5,9,14,15,8 applied r times),
// so every index is a compile-time constant and the block stays in registers.
GL_HD void compress(const u32* cv, const u32* m, u64 counter, u32 block_len, u32 flags, u32* out) {
    u32 s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    u32 s8 = 0x6A09E667u, s9 = 0xBB67AE85u, s10 = 0x3C6EF372u, s11 = 0xA54FF53Au;
    u32 s12 = (u32)counter, s13 = (u32)(counter >> 32), s14 = block_len, s15 = flags;
#define B3_ROUND(i0, i1, i2, i3, i4, i5, i6, i7, i8, i9, i10, i11, i12, i13, i14, i15) \
    B3_G(s0, s4, s8, s12, m[i0], m[i1]); B3_G(s1, s5, s9, s13, m[i2], m[i3]); B3_G(s2, s6, s10, s14, m[i4], m[i5]); B3_G(s3, s7, s11, s15, m[i6], m[i7]); \
    B3_G(s0, s5, s10, s15, m[i8], m[i9]); B3_G(s1, s6, s11, s12, m[i10], m[i11]); B3_G(s2, s7, s8, s13, m[i12], m[i13]); B3_G(s3, s4, s9, s14, m[i14], m[i15]);
    B3_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B3_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    B3_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    B3_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    B3_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    B3_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    B3_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
#undef B3_ROUND
    out[0] = s0 ^ s8; out[1] = s1 ^ s9; out[2] = s2 ^ s10; out[3] = s3 ^ s11;
    out[4] = s4 ^ s12; out[5] = s5 ^ s13; out[6] = s6 ^ s14; out[7] = s7 ^ s15;
}

// Streaming hasher over 32-bit words.  The current block stays buffered until more input arrives or `finish` runs,
// because only then is it known whether it closes its chunk (CHUNK_END) or the whole input (ROOT).
struct Hasher {
    u32 cv[8];            // chaining value of the current chunk
    u32 blk[16];          // buffered block
    u32 n;                // words buffered in blk
    u32 blocks;           // blocks of the current chunk already compressed
    u64 chunk;            // index of the current chunk
    u32 stack[8][8];      // chaining values of completed subtrees (inputs up to 256 KiB)
    u32 depth;

    GL_HD void init() {
        for (int i = 0; i < 8; i++) cv[i] = iv(i);
        n = 0; blocks = 0; chunk = 0; depth = 0;
    }
    GL_HD void parent(const u32* l, const u32* r, u32 flags, u32* out) {
        u32 m[16], key[8];
        for (int i = 0; i < 8; i++) { m[i] = l[i]; m[8 + i] = r[i]; key[i] = iv(i); }
        compress(key, m, 0, 64, PARENT | flags, out);
    }
    // a full buffered block followed by more input: compress it as a non-final block of the input
    GL_HD void spill() {
        u32 flags = blocks == 0 ? CHUNK_START : 0u;
        if (blocks == 15) {                         // 16th block: closes the chunk
            u32 ncv[8];
            compress(cv, blk, chunk, 64, flags | CHUNK_END, ncv);
            u64 total = chunk + 1;                  // merge completed subtrees (one per trailing zero bit)
            while ((total & 1) == 0) { depth--; parent(stack[depth], ncv, 0, ncv); total >>= 1; }
            for (int i = 0; i < 8; i++) { stack[depth][i] = ncv[i]; cv[i] = iv(i); }
            depth++; chunk++; blocks = 0;
        } else {
            compress(cv, blk, chunk, 64, flags, cv);
            blocks++;
        }
        n = 0;
    }
    GL_HD void push(u32 w) {
        if (n == 16) spill();
        blk[n++] = w;
    }
    GL_HD void push64(u64 v) { push((u32)v); push((u32)(v >> 32)); }   // little-endian u64
    // 32-byte digest as 8 little-endian words
    GL_HD void finish(u32* out) {
        u32 len = 4 * n;
        for (u32 i = n; i < 16; i++) blk[i] = 0;
        u32 flags = (blocks == 0 ? CHUNK_START : 0u) | CHUNK_END;
        if (depth == 0) { compress(cv, blk, chunk, len, flags | ROOT, out); return; }
        u32 acc[8];
        compress(cv, blk, chunk, len, flags, acc);
        while (depth > 1) { depth--; parent(stack[depth], acc, 0, acc); }
        parent(stack[0], acc, ROOT, out);
        depth = 0;
    }
};

// digest words <-> the 4 x u64 digest slots every tree of the product uses (little-endian)
GL_HD void words_to_u64(const u32* w, u64* d) { for (int i = 0; i < 4; i++) d[i] = (u64)w[2 * i] | ((u64)w[2 * i + 1] << 32); }

// hash(l || r): one compression (CompressionFunctionFromHasher<Blake3, 2, 32>)
GL_HD void compress2(const u64* l, const u64* r, u64* out) {
    u32 m[16], key[8], o[8];
    for (int i = 0; i < 4; i++) { m[2 * i] = (u32)l[i]; m[2 * i + 1] = (u32)(l[i] >> 32); m[8 + 2 * i] = (u32)r[i]; m[9 + 2 * i] = (u32)(r[i] >> 32); }
    for (int i = 0; i < 8; i++) key[i] = iv(i);
    compress(key, m, 0, 64, CHUNK_START | CHUNK_END | ROOT, o);
    words_to_u64(o, out);
}

}  // namespace b3
