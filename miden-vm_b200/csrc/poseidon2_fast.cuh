// Device-only Poseidon2 (Goldilocks, width 12) tuned for the sm_100a integer pipes.
//
// Same function as p2::permute (poseidon2.cuh; reference crates/crypto/src/hash/
// algebraic_sponge/poseidon2/mod.rs:226-319) but with lazy reduction:
//   * field elements travel as arbitrary u64 representatives (value mod p, not necessarily < p);
//   * multiplication reduces the 128-bit product with two carry-flag sequences and no final
//     compare/subtract;
//   * the linear layers accumulate in 96-bit (u64 + u32) registers with add-with-carry and reduce
//     once per output.
// The first version (canonical add/sub everywhere) spent ~30k instructions per permutation,
// 70 % of them on the ALU pipe in ISETP/SEL fix-ups (profiles/r1a_hash_kernels.md); this one is
// ~40 % of that.  Callers canonicalise (glf::canon) whatever they store.
#pragma once
#include "gl.cuh"

namespace glf {
using gl::u64;
using gl::u32;

struct W { u64 lo; u32 hi; };   // lo + hi * 2^64, hi stays tiny (< 2^8 here)

__device__ __forceinline__ u64 canon(u64 x) { return x >= gl::P ? x - gl::P : x; }

// 128-bit (hi:lo) -> u64 representative.  2^64 = 2^32 - 1, 2^96 = -1 (mod p).
__device__ __forceinline__ u64 red128(u64 lo, u64 hi) {
    u32 hl = (u32)hi, hh = (u32)(hi >> 32);
    u64 t, m, r;
    u32 b, c;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(t), "=r"(b) : "l"(lo), "l"((u64)hh));
    t -= (u64)b;                                    // b = 0xFFFFFFFF iff the subtraction borrowed
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(m) : "r"(hl), "r"(0xFFFFFFFFu));
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(t), "l"(m));
    r += (u64)(0u - c);                             // + (2^32 - 1) iff the addition carried
    return r;
}
__device__ __forceinline__ u64 mul(u64 a, u64 b) { return red128(a * b, __umul64hi(a, b)); }

// x + k for a canonical constant k (< p): one carry fix-up, cannot carry twice.
__device__ __forceinline__ u64 add_const(u64 x, u64 k) {
    u64 r; u32 c;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(x), "l"(k));
    return r + (u64)(0u - c);
}

__device__ __forceinline__ W wide(u64 x) { W w; w.lo = x; w.hi = 0; return w; }
__device__ __forceinline__ void wadd(W& w, u64 x) {
    asm("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+l"(w.lo), "+r"(w.hi) : "l"(x));
}
__device__ __forceinline__ void wadd(W& w, W x) {
    asm("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+l"(w.lo), "+r"(w.hi) : "l"(x.lo), "r"(x.hi));
}
// w + (8p - x) for a wide x < 8p: subtraction without going negative.  8p = 2^67 - 2^35 + 8.
__device__ __forceinline__ void wsub(W& w, W x) {
    u64 lo; u32 hi;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, %4, %5;" : "=l"(lo), "=r"(hi)
        : "l"(0xFFFFFFF800000008ull), "l"(x.lo), "r"(7u), "r"(x.hi));
    W t; t.lo = lo; t.hi = hi;
    wadd(w, t);
}
__device__ __forceinline__ W wshl(u64 x, int k) { W w; w.lo = x << k; w.hi = (u32)(x >> (64 - k)); return w; }
__device__ __forceinline__ u64 wred(W w) {
    u64 m, r; u32 c;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(m) : "r"(w.hi), "r"(0xFFFFFFFFu));
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(w.lo), "l"(m));
    return r + (u64)(0u - c);
}
// ---- canonical (< p) arithmetic built on the carry flag, for the NTT butterflies ----------------
__device__ __forceinline__ u64 canon_cc(u64 r) {      // r in [0, 2^64) -> r mod p
    u64 t; u32 m;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(t), "=r"(m) : "l"(r), "l"(0xFFFFFFFF00000001ull));
    return t - (u64)m;                                 // borrowed (r < p): add p back == subtract 2^32 - 1
}
__device__ __forceinline__ u64 csub(u64 a, u64 b) {   // a, b < p
    u64 d; u32 m;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(d), "=r"(m) : "l"(a), "l"(b));
    return d - (u64)m;
}
__device__ __forceinline__ u64 cadd(u64 a, u64 b) { return csub(a, 0xFFFFFFFF00000001ull - b); }   // a - (p - b)
__device__ __forceinline__ u64 cmul(u64 a, u64 b) { return canon_cc(mul(a, b)); }

// x / 2 for any representative: result < 2^64.
__device__ __forceinline__ u64 half(u64 x) { return (x >> 1) + ((x & 1) ? 0x7FFFFFFF80000001ull : 0ull); }

}  // namespace glf

namespace p2f {
using gl::u64;
using gl::u32;
using glf::W;

__device__ __forceinline__ u64 sbox(u64 x) {
    u64 x2 = glf::mul(x, x), x3 = glf::mul(x2, x), x4 = glf::mul(x2, x2);
    return glf::mul(x3, x4);
}

__device__ __forceinline__ void external_layer(u64* s) {
    W y[12];
#pragma unroll
    for (int c = 0; c < 12; c += 4) {
        W sum = glf::wide(s[c]);
        glf::wadd(sum, s[c + 1]); glf::wadd(sum, s[c + 2]); glf::wadd(sum, s[c + 3]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            W t = sum;
            u64 nx = s[c + ((i + 1) & 3)];
            glf::wadd(t, s[c + i]); glf::wadd(t, nx); glf::wadd(t, nx);   // sum + x_i + 2 x_{i+1}
            y[c + i] = t;
        }
    }
#pragma unroll
    for (int l = 0; l < 4; l++) {
        W col = y[l];
        glf::wadd(col, y[4 + l]); glf::wadd(col, y[8 + l]);
#pragma unroll
        for (int c = 0; c < 12; c += 4) {
            W o = y[c + l];
            glf::wadd(o, col);
            s[c + l] = glf::wred(o);
        }
    }
}

// s_i <- d_i s_i + sum,  d = [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 1/4, -1/4, 1/8]
__device__ __forceinline__ void internal_layer(u64* s) {
    W sum = glf::wide(s[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) glf::wadd(sum, s[i]);
    W o;
    u64 h3 = glf::half(s[3]), h6 = glf::half(s[6]);
    u64 q9 = glf::half(glf::half(s[9])), q10 = glf::half(glf::half(s[10]));
    u64 e11 = glf::half(glf::half(glf::half(s[11])));
    W t7 = glf::wshl(s[7], 1); glf::wadd(t7, s[7]);
    W t4 = glf::wshl(s[4], 1); glf::wadd(t4, s[4]);
    o = sum; glf::wsub(o, glf::wshl(s[0], 1)); s[0] = glf::wred(o);
    o = sum; glf::wadd(o, s[1]); s[1] = glf::wred(o);
    o = sum; glf::wadd(o, glf::wshl(s[2], 1)); s[2] = glf::wred(o);
    o = sum; glf::wadd(o, h3); s[3] = glf::wred(o);
    o = sum; glf::wadd(o, t4); s[4] = glf::wred(o);
    o = sum; glf::wadd(o, glf::wshl(s[5], 2)); s[5] = glf::wred(o);
    o = sum; glf::wsub(o, glf::wide(h6)); s[6] = glf::wred(o);
    o = sum; glf::wsub(o, t7); s[7] = glf::wred(o);
    o = sum; glf::wsub(o, glf::wshl(s[8], 2)); s[8] = glf::wred(o);
    o = sum; glf::wadd(o, q9); s[9] = glf::wred(o);
    o = sum; glf::wsub(o, glf::wide(q10)); s[10] = glf::wred(o);
    o = sum; glf::wadd(o, e11); s[11] = glf::wred(o);
}

// Output words are arbitrary representatives; canonicalise with glf::canon before storing.
__device__ __forceinline__ void permute(u64* s) {
    external_layer(s);
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        const u64* rc = phase ? p2::D_RC_EXT_TERMINAL : p2::D_RC_EXT_INITIAL;
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = sbox(glf::add_const(s[i], rc[12 * r + i]));
            external_layer(s);
        }
        if (phase == 0) {
#pragma unroll 1
            for (int r = 0; r < 22; r++) {
                s[0] = sbox(glf::add_const(s[0], p2::D_RC_INTERNAL[r]));
                internal_layer(s);
            }
        }
    }
}

}  // namespace p2f
