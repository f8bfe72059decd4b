// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Eight Poseidon2 permutations at once in AVX-512 lanes: the same round structure as poseidon2.hpp (which stays
// the definition and the fallback), used by LmcsTree::build for the leaf sponges and the compression layers so
// that the CPU baseline is not a purely scalar strawman.  The reference itself reaches its hashing throughput
// through p3's packed Goldilocks types (AVX2/AVX-512/NEON); this is the analogous packing, restated.
// Checked lane by lane against the scalar permutation in tests/test_oracle.py::test_poseidon2_x8_matches_scalar.
#pragma once
#include "poseidon2.hpp"
#if defined(__x86_64__)
#include <immintrin.h>
#define ORC_HAVE_X8 1
#define ORC_X8 __attribute__((target("avx512f,avx512dq"), always_inline)) inline
#define ORC_X8_FN __attribute__((target("avx512f,avx512dq")))

namespace orc {

inline bool x8_available() {
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq");
    return ok;
}

namespace x8 {
typedef __m512i V;

ORC_X8 V bc(u64 x) { return _mm512_set1_epi64((long long)x); }
// canonical (< p) operands and results
ORC_X8 V add(V a, V b) {
    V s = _mm512_add_epi64(a, b);
    __mmask8 fix = _mm512_cmplt_epu64_mask(s, a) | _mm512_cmpge_epu64_mask(s, bc(P));
    return _mm512_mask_sub_epi64(s, fix, s, bc(P));
}
ORC_X8 V sub(V a, V b) {
    V d = _mm512_sub_epi64(a, b);
    return _mm512_mask_add_epi64(d, _mm512_cmplt_epu64_mask(a, b), d, bc(P));
}
ORC_X8 V mul(V a, V b) {
    const V m32 = bc(0xFFFFFFFFULL);
    V ah = _mm512_srli_epi64(a, 32), bh = _mm512_srli_epi64(b, 32);
    V ll = _mm512_mul_epu32(a, b), lh = _mm512_mul_epu32(a, bh), hl = _mm512_mul_epu32(ah, b), hh = _mm512_mul_epu32(ah, bh);
    V mid = _mm512_add_epi64(lh, _mm512_srli_epi64(ll, 32));                 // < 2^64
    V mid2 = _mm512_add_epi64(hl, _mm512_and_si512(mid, m32));              // < 2^64
    V lo = _mm512_or_si512(_mm512_and_si512(ll, m32), _mm512_slli_epi64(mid2, 32));
    V hi = _mm512_add_epi64(hh, _mm512_add_epi64(_mm512_srli_epi64(mid, 32), _mm512_srli_epi64(mid2, 32)));
    // 2^64 = 2^32 - 1, 2^96 = -1 (mod p):  lo - (hi >> 32) + (hi & m32) * (2^32 - 1)
    V hh32 = _mm512_srli_epi64(hi, 32), hl32 = _mm512_and_si512(hi, m32);
    V t = _mm512_sub_epi64(lo, hh32);
    t = _mm512_mask_sub_epi64(t, _mm512_cmplt_epu64_mask(lo, hh32), t, m32);         // borrowed: + p  ==  - (2^32 - 1)
    V m = _mm512_sub_epi64(_mm512_slli_epi64(hl32, 32), hl32);
    V r = _mm512_add_epi64(t, m);
    r = _mm512_mask_add_epi64(r, _mm512_cmplt_epu64_mask(r, t), r, m32);             // carried: 2^64 = 2^32 - 1
    return _mm512_mask_sub_epi64(r, _mm512_cmpge_epu64_mask(r, bc(P)), r, bc(P));
}
ORC_X8 V sbox(V x) { V x2 = mul(x, x), x3 = mul(x2, x), x4 = mul(x2, x2); return mul(x3, x4); }

ORC_X8 void external(V* s) {
    for (int c = 0; c < 3; c++) {
        V x0 = s[4 * c], x1 = s[4 * c + 1], x2 = s[4 * c + 2], x3 = s[4 * c + 3];
        V sum = add(add(x0, x1), add(x2, x3));
        s[4 * c + 0] = add(add(sum, x0), add(x1, x1));
        s[4 * c + 1] = add(add(sum, x1), add(x2, x2));
        s[4 * c + 2] = add(add(sum, x2), add(x3, x3));
        s[4 * c + 3] = add(add(sum, x3), add(x0, x0));
    }
    V col[4];
    for (int l = 0; l < 4; l++) col[l] = add(add(s[l], s[4 + l]), s[8 + l]);
    for (int i = 0; i < 12; i++) s[i] = add(s[i], col[i % 4]);
}
ORC_X8 void internal(V* s) {
    V sum = s[0];
    for (int i = 1; i < 12; i++) sum = add(sum, s[i]);
    for (int i = 0; i < 12; i++) s[i] = add(mul(s[i], bc(P2_INTERNAL_DIAG[i])), sum);
}
ORC_X8_FN inline void permute(V* s) {
    external(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) s[i] = sbox(add(s[i], bc(P2_RC_EXT_INITIAL[12 * r + i])));
        external(s);
    }
    for (int r = 0; r < 22; r++) {
        s[0] = sbox(add(s[0], bc(P2_RC_INTERNAL[r])));
        internal(s);
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) s[i] = sbox(add(s[i], bc(P2_RC_EXT_TERMINAL[12 * r + i])));
        external(s);
    }
}
}  // namespace x8

// Eight array-of-struct states (consecutive State objects) through one permutation.
ORC_X8_FN inline void poseidon2_permute_x8(State* st8) {
    const __m512i idx = _mm512_set_epi64(84, 72, 60, 48, 36, 24, 12, 0);   // lane l -> st8[l]
    x8::V s[12];
    const long long* base = (const long long*)st8;
    for (int i = 0; i < 12; i++) s[i] = _mm512_i64gather_epi64(idx, base + i, 8);
    x8::permute(s);
    for (int i = 0; i < 12; i++) _mm512_i64scatter_epi64((long long*)st8 + i, idx, s[i], 8);
}

// sponge_absorb (poseidon2.hpp) for the 8 consecutive states st8[0..8) and the 8 rows rows + l * stride.
ORC_X8_FN inline void sponge_absorb_x8(State* st8, const Fp* rows, size_t stride, size_t n) {
    if (n == 0) return;
    const __m512i sidx = _mm512_set_epi64(84, 72, 60, 48, 36, 24, 12, 0);
    const __m512i ridx = _mm512_mullo_epi64(_mm512_set_epi64(7, 6, 5, 4, 3, 2, 1, 0), _mm512_set1_epi64((long long)stride));
    x8::V s[12];
    const long long* sb = (const long long*)st8;
    const long long* rb = (const long long*)rows;
    for (int i = 0; i < 12; i++) s[i] = _mm512_i64gather_epi64(sidx, sb + i, 8);
    for (size_t i = 0; i < n; i += 8) {
        size_t rem = n - i < 8 ? n - i : 8;
        for (size_t k = 0; k < rem; k++) s[k] = _mm512_i64gather_epi64(ridx, rb + i + k, 8);
        for (size_t k = rem; k < 8; k++) s[k] = _mm512_setzero_si512();
        x8::permute(s);
    }
    for (int i = 0; i < 12; i++) _mm512_i64scatter_epi64((long long*)st8 + i, sidx, s[i], 8);
}

// parents[0..8) = compress2(children[2l], children[2l + 1])
ORC_X8_FN inline void compress2_x8(const Digest* children, Digest* parents) {
    const __m512i cidx = _mm512_set_epi64(56, 48, 40, 32, 24, 16, 8, 0);    // 8 u64 of children per parent
    const __m512i pidx = _mm512_set_epi64(28, 24, 20, 16, 12, 8, 4, 0);
    x8::V s[12];
    const long long* cb = (const long long*)children;
    for (int k = 0; k < 8; k++) s[k] = _mm512_i64gather_epi64(cidx, cb + k, 8);
    for (int k = 8; k < 12; k++) s[k] = _mm512_setzero_si512();
    x8::permute(s);
    for (int k = 0; k < 4; k++) _mm512_i64scatter_epi64((long long*)parents + k, pidx, s[k], 8);
}

}  // namespace orc
#else
#define ORC_HAVE_X8 0
namespace orc { inline bool x8_available() { return false; } }
#endif
