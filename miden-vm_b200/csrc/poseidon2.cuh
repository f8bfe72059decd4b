// Poseidon2 permutation (Goldilocks, width 12, x^7, 4 + 22 + 4 rounds), device + host.
// Behaviour: p3 `default_goldilocks_poseidon2_12` as wrapped by the reference at
// crates/crypto/src/hash/algebraic_sponge/poseidon2/mod.rs:22-37; layer structure as in
// mod.rs:226-319.  The internal diagonal is [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 1/4, -1/4, 1/8]
// (constants.rs:18-31), applied with doublings/halvings instead of multiplications.
#pragma once
#include "gl.cuh"

namespace p2 {
using gl::u64;

#include "poseidon2_constants.inc"   // host copies (static const arrays)

#ifdef __CUDACC__
__constant__ u64 D_RC_EXT_INITIAL[48];
__constant__ u64 D_RC_INTERNAL[22];
__constant__ u64 D_RC_EXT_TERMINAL[48];
#endif

GL_HD u64 sbox(u64 x) {
    u64 x2 = gl::sqr(x), x3 = gl::mul(x2, x), x4 = gl::sqr(x2);
    return gl::mul(x3, x4);
}

// External layer: circ-like 4x4 MDS [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]] on each chunk, then
// every chunk receives the sum of all chunks (block-circulant [2M, M, M]).
GL_HD void external_layer(u64* s) {
#pragma unroll
    for (int c = 0; c < 12; c += 4) {
        u64 x0 = s[c], x1 = s[c + 1], x2 = s[c + 2], x3 = s[c + 3];
        u64 sum = gl::add(gl::add(x0, x1), gl::add(x2, x3));
        s[c] = gl::add(gl::add(sum, x0), gl::dbl(x1));
        s[c + 1] = gl::add(gl::add(sum, x1), gl::dbl(x2));
        s[c + 2] = gl::add(gl::add(sum, x2), gl::dbl(x3));
        s[c + 3] = gl::add(gl::add(sum, x3), gl::dbl(x0));
    }
#pragma unroll
    for (int l = 0; l < 4; l++) {
        u64 col = gl::add(gl::add(s[l], s[4 + l]), s[8 + l]);
        s[l] = gl::add(s[l], col);
        s[4 + l] = gl::add(s[4 + l], col);
        s[8 + l] = gl::add(s[8 + l], col);
    }
}

// Internal layer: s_i <- d_i * s_i + sum(s).
GL_HD void internal_layer(u64* s) {
    u64 sum = gl::add(gl::add(gl::add(s[0], s[1]), gl::add(s[2], s[3])),
                      gl::add(gl::add(gl::add(s[4], s[5]), gl::add(s[6], s[7])),
                              gl::add(gl::add(s[8], s[9]), gl::add(s[10], s[11]))));
    u64 h3 = gl::half(s[3]), h6 = gl::half(s[6]);
    u64 q9 = gl::half(gl::half(s[9])), q10 = gl::half(gl::half(s[10]));
    u64 e11 = gl::half(gl::half(gl::half(s[11])));
    u64 d4 = gl::dbl(s[4]), d5 = gl::dbl(gl::dbl(s[5])), d7 = gl::dbl(s[7]), d8 = gl::dbl(gl::dbl(s[8]));
    s[0] = gl::sub(sum, gl::dbl(s[0]));
    s[1] = gl::add(sum, s[1]);
    s[2] = gl::add(sum, gl::dbl(s[2]));
    s[3] = gl::add(sum, h3);
    s[4] = gl::add(sum, gl::add(d4, s[4]));
    s[5] = gl::add(sum, d5);
    s[6] = gl::sub(sum, h6);
    s[7] = gl::sub(sum, gl::add(d7, s[7]));
    s[8] = gl::sub(sum, d8);
    s[9] = gl::add(sum, q9);
    s[10] = gl::sub(sum, q10);
    s[11] = gl::add(sum, e11);
}

GL_HD void permute(u64* s) {
#ifdef __CUDA_ARCH__
    const u64* rci = D_RC_EXT_INITIAL; const u64* rcm = D_RC_INTERNAL; const u64* rct = D_RC_EXT_TERMINAL;
#else
    const u64* rci = P2_RC_EXT_INITIAL; const u64* rcm = P2_RC_INTERNAL; const u64* rct = P2_RC_EXT_TERMINAL;
#endif
    external_layer(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox(gl::add(s[i], rci[12 * r + i]));
        external_layer(s);
    }
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
        s[0] = sbox(gl::add(s[0], rcm[r]));
        internal_layer(s);
    }
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox(gl::add(s[i], rct[12 * r + i]));
        external_layer(s);
    }
}

}  // namespace p2
