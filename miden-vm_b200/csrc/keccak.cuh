// Keccak-f[1600] and the two constructions the reference's `HashFunction::Keccak` configuration builds from it
// (air/src/config.rs:310-353; KeccakF / Keccak256Hash = p3_keccak 0.6.2 = the published Keccak, crates/crypto/src/hash/keccak/mod.rs:18):
//   * leaves: `SerializingStatefulSponge<StatefulSponge<KeccakF, 25, 17, 4>>` -- the same OVERWRITE-mode sponge as the Poseidon2
//     one (crates/stateful-hasher/src/field_sponge.rs:41-59) over u64 lanes: each chunk of 17 canonical u64 overwrites
//     state[0..17), a partial chunk is zero-filled to the rate, one permutation per chunk; digest = state[0..4); alignment 17;
//   * nodes: `CompressionFunctionFromHasher<PaddingFreeSponge<KeccakF, 25, 17, 4>, 2, 4>` -- zero state, the 8 input words
//     overwrite state[0..8), one permutation, first 4 words;
//   * transcript: `HashChallenger<u8, Keccak256Hash, 32>` -- Keccak-256 (rate 136 bytes, XOR absorption, padding 0x01 .. 0x80).
// tests/test_keccak.py pins the permutation and the hash on hashlib's SHA3 (same permutation, padding byte 0x06) and on the
// Keccak-256 of the empty string.
#pragma once
#include "gl.cuh"

namespace kk {
using gl::u32;
using gl::u64;

GL_HD u64 rotl(u64 x, int n) { return (x << n) | (x >> (64 - n)); }

// iota round constants: constant memory on the device (the round loop is not unrolled, so the index is dynamic)
#define KK_RC_VALUES {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull, \
                      0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull, \
                      0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, \
                      0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull, \
                      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull}
static const u64 H_RC[24] = KK_RC_VALUES;
#ifdef __CUDACC__
__constant__ u64 D_RC[24] = KK_RC_VALUES;
#endif

GL_HD void permute(u64* st) {
#ifdef __CUDA_ARCH__
    const u64* RC = D_RC;
#else
    const u64* RC = H_RC;
#endif
    constexpr int ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    constexpr int PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        u64 bc[5];
#pragma unroll
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];          // theta
#pragma unroll
        for (int i = 0; i < 5; i++) {
            u64 t = bc[(i + 4) % 5] ^ rotl(bc[(i + 1) % 5], 1);
#pragma unroll
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        u64 t = st[1];                                                                                        // rho + pi
#pragma unroll
        for (int i = 0; i < 24; i++) { u64 b = st[PILN[i]]; st[PILN[i]] = rotl(t, ROTC[i]); t = b; }
#pragma unroll
        for (int j = 0; j < 25; j += 5) {                                                                     // chi
#pragma unroll
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
#pragma unroll
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= RC[round];                                                                                   // iota
    }
}

// PaddingFreeSponge<KeccakF, 25, 17, 4> on the 8 words of two digests
GL_HD void compress2(const u64* l, const u64* r, u64* out) {
    u64 st[25];
#pragma unroll
    for (int i = 0; i < 25; i++) st[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { st[i] = l[i]; st[4 + i] = r[i]; }
    permute(st);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = st[i];
}

// Keccak-256 over whole little-endian u64 words (every transcript observation is 8 or 32 bytes), streaming.
// `pad` = 0x01 for Keccak-256 (what the reference uses), 0x06 for SHA3-256 (what hashlib offers as an independent check).
struct Hash256 {
    u64 st[25];
    u32 k;          // words absorbed into the current block (< 17)
    GL_HD void init() { for (int i = 0; i < 25; i++) st[i] = 0; k = 0; }
    GL_HD void push64(u64 w) {
        // select the lane with compile-time indices so that the state stays in registers
#pragma unroll
        for (int i = 0; i < 17; i++) if ((u32)i == k) st[i] ^= w;
        if (++k == 17) { permute(st); k = 0; }
    }
    GL_HD void finish(u64* out4, u64 pad = 0x01) {
#pragma unroll
        for (int i = 0; i < 17; i++) if ((u32)i == k) st[i] ^= pad;
        st[16] ^= 0x8000000000000000ull;
        permute(st);
        for (int i = 0; i < 4; i++) out4[i] = st[i];
    }
};

}  // namespace kk
