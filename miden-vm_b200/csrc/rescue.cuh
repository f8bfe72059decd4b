// Rescue Prime Optimized (RPO) and RPX permutations over Goldilocks (width 12, 7 rounds) for host and device code: the
// permutations of the reference's `rpo_config` / `rpx_config` (air/src/config.rs:225-248), whose LMCS, node compression and duplex
// challenger are those of the Poseidon2 configuration with the permutation swapped (`alg_config<P>`, :255-273).
//   round structure : crates/crypto/src/hash/algebraic_sponge/rescue/rpo/mod.rs:185-207 (MDS, +ARK1, x^7, MDS, +ARK2, x^(1/7)),
//                     rescue/rpx/mod.rs:185-266 ((FB)(E)(FB)(E)(FB)(E)(M); (E): +ARK1, then x^7 in F_p[x]/(x^3 - x - 1) on 4 triples)
//   MDS             : circulant with the small first row of rescue/mds/mod.rs:47-62; here as two 12-term integer dot products over
//                     the 32-bit halves of the state (they fit 64 bits) recombined and reduced once per output
//   x^(1/7)         : exponent 10540996611094048183 = 0b1001001001001001001001001001000110110110110110110110110110110111; the
//                     repeated-pattern chain below takes 72 multiplications
// Canonical gl:: arithmetic throughout (these configurations are functional coverage: an RPO permutation is ~6400 field
// multiplications against Poseidon2's 472).  Pinned through tests/cpp/test_rescue.cpp on the reference's 19 RPO known answers
// (rescue/rpo/tests.rs:241-430) and on the oracle's independent restatement.
#pragma once
#include "gl.cuh"

namespace rsc {
using gl::u32;
using gl::u64;

#include "rescue_constants.inc"      // host copies: RESCUE_MDS_ROW, RESCUE_ARK1, RESCUE_ARK2
#ifdef __CUDACC__
__constant__ u64 D_ARK1[84];
__constant__ u64 D_ARK2[84];
#endif

GL_HD void mds(u64* s) {
    constexpr u32 ROW[12] = {7, 23, 8, 26, 13, 10, 9, 7, 6, 22, 21, 8};
    u64 o[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u64 lo = 0, hi = 0;                 // sums of at most 12 * 26 * 2^32 < 2^41
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const u32 c = ROW[(j - i + 12) % 12];
            lo += (u64)(u32)s[j] * c;
            hi += (u64)(u32)(s[j] >> 32) * c;
        }
        u64 l = lo + (hi << 32);
        u64 h = (hi >> 32) + (l < lo ? 1u : 0u);         // value = l + h * 2^64 with h < 2^10, and 2^64 = 2^32 - 1 (mod p)
        u64 r = l + h * 0xFFFFFFFFull;
        r += (r < l) ? 0xFFFFFFFFull : 0ull;            // wrapped: r < 2^42, the fold fits
        o[i] = r >= gl::P ? r - gl::P : r;
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = o[i];
}

GL_HD u64 pow7(u64 x) { u64 x2 = gl::sqr(x), x3 = gl::mul(x2, x), x4 = gl::sqr(x2); return gl::mul(x3, x4); }
GL_HD u64 sqn(u64 x, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) x = gl::sqr(x);
    return x;
}
// x^(1/7): bit pattern (100)^10 0 (011)^10 + tail, built from blocks of "100" doubled in length, then the run of "011"/"111"
GL_HD u64 inv_pow7(u64 x) {
    u64 p10 = gl::sqr(x);                                   // exponent 0b10
    u64 p100 = gl::sqr(p10);                                // 0b100
    u64 r2 = gl::mul(sqn(p100, 3), p100);                   // 0b100100
    u64 r4 = gl::mul(sqn(r2, 6), r2);                       // (100) x 4
    u64 r8 = gl::mul(sqn(r4, 12), r4);                      // (100) x 8
    u64 r10 = gl::mul(sqn(r8, 6), r2);                      // (100) x 10
    u64 mid = gl::mul(sqn(r10, 31), r10);                   // (100)x10 0 (100)x10
    u64 a = sqn(gl::mul(gl::sqr(mid), r10), 2);
    u64 b = gl::mul(gl::mul(p10, p100), x);                 // 0b111
    return gl::mul(a, b);
}

// ---- F_p[x] / (x^3 - x - 1) for the RPX (E) round: three-way Karatsuba product, x^3 = x + 1 and x^4 = x^2 + x
struct C3 { u64 c0, c1, c2; };
GL_HD C3 c3_reduce(u64 d0, u64 d1, u64 d2, u64 d3, u64 d4) {
    C3 r; r.c0 = gl::add(d0, d3); r.c1 = gl::add(gl::add(d1, d3), d4); r.c2 = gl::add(d2, d4);
    return r;
}
GL_HD C3 c3_mul(C3 a, C3 b) {
    u64 v0 = gl::mul(a.c0, b.c0), v1 = gl::mul(a.c1, b.c1), v2 = gl::mul(a.c2, b.c2);
    u64 m01 = gl::mul(gl::add(a.c0, a.c1), gl::add(b.c0, b.c1));
    u64 m02 = gl::mul(gl::add(a.c0, a.c2), gl::add(b.c0, b.c2));
    u64 m12 = gl::mul(gl::add(a.c1, a.c2), gl::add(b.c1, b.c2));
    u64 d1 = gl::sub(gl::sub(m01, v0), v1);
    u64 d2 = gl::add(gl::sub(gl::sub(m02, v0), v2), v1);
    u64 d3 = gl::sub(gl::sub(m12, v1), v2);
    return c3_reduce(v0, d1, d2, d3, v2);
}
GL_HD C3 c3_sqr(C3 a) {
    u64 a01 = gl::mul(a.c0, a.c1), a02 = gl::mul(a.c0, a.c2), a12 = gl::mul(a.c1, a.c2);
    return c3_reduce(gl::sqr(a.c0), gl::dbl(a01), gl::add(gl::dbl(a02), gl::sqr(a.c1)), gl::dbl(a12), gl::sqr(a.c2));
}
GL_HD C3 c3_pow7(C3 a) { C3 a2 = c3_sqr(a), a3 = c3_mul(a2, a), a4 = c3_sqr(a2); return c3_mul(a3, a4); }

GL_HD void fb_round(u64* s, const u64* ark1, const u64* ark2) {
    mds(s);
#pragma unroll 1
    for (int i = 0; i < 12; i++) s[i] = pow7(gl::add(s[i], ark1[i]));
    mds(s);
#pragma unroll 1
    for (int i = 0; i < 12; i++) s[i] = inv_pow7(gl::add(s[i], ark2[i]));
}
GL_HD void ext_round(u64* s, const u64* ark1) {
#pragma unroll
    for (int k = 0; k < 12; k += 3) {
        C3 a; a.c0 = gl::add(s[k], ark1[k]); a.c1 = gl::add(s[k + 1], ark1[k + 1]); a.c2 = gl::add(s[k + 2], ark1[k + 2]);
        C3 r = c3_pow7(a);
        s[k] = r.c0; s[k + 1] = r.c1; s[k + 2] = r.c2;
    }
}

#ifdef __CUDA_ARCH__
#define RSC_ARK1 rsc::D_ARK1
#define RSC_ARK2 rsc::D_ARK2
#else
#define RSC_ARK1 rsc::RESCUE_ARK1
#define RSC_ARK2 rsc::RESCUE_ARK2
#endif

// state: canonical felts in, canonical felts out.  The 12 state words are indexed by unrolled loops only, except in the S-box
// loops, which stay rolled (code size); the compiler keeps the state in local memory there -- acceptable for this coverage path.
GL_HD void rpo_permute(u64* s) {
#pragma unroll 1
    for (int r = 0; r < 7; r++) fb_round(s, RSC_ARK1 + 12 * r, RSC_ARK2 + 12 * r);
}
GL_HD void rpx_permute(u64* s) {
#pragma unroll 1
    for (int r = 0; r < 6; r += 2) {
        fb_round(s, RSC_ARK1 + 12 * r, RSC_ARK2 + 12 * r);
        ext_round(s, RSC_ARK1 + 12 * (r + 1));
    }
    mds(s);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::add(s[i], RSC_ARK1[72 + i]);
}

}  // namespace rsc
