// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Rescue Prime Optimized (RPO) and RPX permutations over Goldilocks, width 12, 7 rounds -- the permutations of the reference's
// `rpo_config` / `rpx_config` (air/src/config.rs:225-248), restated from
//   crates/crypto/src/hash/algebraic_sponge/rescue/mod.rs:24-116   S-box x^7, inverse S-box x^(1/7), constant addition
//   rescue/rpo/mod.rs:185-207                                      RPO round: MDS, +ARK1, x^7, MDS, +ARK2, x^(1/7)
//   rescue/rpx/mod.rs:185-266                                      RPX: (FB)(E)(FB)(E)(FB)(E)(M); E = +ARK1 then x^7 in F_p[x]/(x^3 - x - 1)
//   rescue/mds/mod.rs:14-45,47-213                                 the circulant MDS matrix
// written the plain way: the matrix-vector product entry by entry, the inverse S-box as a generic square-and-multiply with the
// exponent (7^-1 mod p-1), the cubic-extension product as a schoolbook product reduced with x^3 = x + 1 -- none of the
// reference's (or the product's) frequency-domain / addition-chain / Karatsuba shortcuts.
// PINNED on the reference's own 19 `hash_elements` known answers (rescue/rpo/tests.rs:241-430; tests/test_rescue.py) for RPO.
// RPX has no known-answer vector in the reference tree: its (E) round is checked against big-integer polynomial arithmetic in
// Python, everything else it shares with RPO.
#pragma once
#include "field.hpp"
#include <array>

namespace orc {

#include "rescue_constants.inc"

using RState = std::array<Fp, 12>;

constexpr u64 RESCUE_INV_ALPHA = 10540996611094048183ULL;     // 7 * INV_ALPHA = 1 (mod p - 1), rescue/rpo/tests.rs:12-14

inline void rescue_mds(RState& s) {
    RState o;
    for (int i = 0; i < 12; i++) {
        Fp acc;
        for (int j = 0; j < 12; j++) acc += Fp::raw(RESCUE_MDS_ROW[(j - i + 12) % 12]) * s[j];
        o[i] = acc;
    }
    s = o;
}
inline Fp rescue_pow7(Fp x) { Fp x2 = x * x, x4 = x2 * x2; return x4 * x2 * x; }
inline void rescue_fb_round(RState& s, int round) {
    rescue_mds(s);
    for (int i = 0; i < 12; i++) s[i] = rescue_pow7(s[i] + Fp::raw(RESCUE_ARK1[12 * round + i]));
    rescue_mds(s);
    for (int i = 0; i < 12; i++) s[i] = fp_pow(s[i] + Fp::raw(RESCUE_ARK2[12 * round + i]), RESCUE_INV_ALPHA);
}
inline void rpo_permute(RState& s) { for (int r = 0; r < 7; r++) rescue_fb_round(s, r); }

// F_p[x] / (x^3 - x - 1): schoolbook product, then x^3 = x + 1, x^4 = x^2 + x
using Cubic = std::array<Fp, 3>;
inline Cubic cubic_mul(const Cubic& a, const Cubic& b) {
    Fp c[5];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c[i + j] += a[i] * b[j];
    return Cubic{c[0] + c[3], c[1] + c[3] + c[4], c[2] + c[4]};
}
inline Cubic cubic_pow7(const Cubic& a) {
    Cubic a2 = cubic_mul(a, a), a4 = cubic_mul(a2, a2);
    return cubic_mul(cubic_mul(a4, a2), a);
}
inline void rpx_ext_round(RState& s, int round) {
    for (int i = 0; i < 12; i++) s[i] += Fp::raw(RESCUE_ARK1[12 * round + i]);
    for (int k = 0; k < 4; k++) {
        Cubic r = cubic_pow7(Cubic{s[3 * k], s[3 * k + 1], s[3 * k + 2]});
        for (int i = 0; i < 3; i++) s[3 * k + i] = r[i];
    }
}
inline void rpx_permute(RState& s) {
    rescue_fb_round(s, 0); rpx_ext_round(s, 1);
    rescue_fb_round(s, 2); rpx_ext_round(s, 3);
    rescue_fb_round(s, 4); rpx_ext_round(s, 5);
    rescue_mds(s);
    for (int i = 0; i < 12; i++) s[i] += Fp::raw(RESCUE_ARK1[12 * 6 + i]);
}

// AlgebraicSponge::hash_elements (crates/crypto/src/hash/algebraic_sponge/mod.rs:62-69,215-265) -- only what the reference's
// known-answer test needs: zero state, capacity[0] = len mod 8, overwrite the rate 8 elements at a time, zero-pad the last block.
template <class Perm>
inline std::array<Fp, 4> rescue_hash_elements(const Fp* e, size_t n, Perm&& permute) {
    RState st{};
    st[8] = Fp::raw(n % 8);
    size_t i = 0;
    for (size_t k = 0; k < n; k++) {
        st[i++] = e[k];
        if (i == 8) { permute(st); i = 0; }
    }
    if (i > 0) { while (i < 8) st[i++] = Fp(); permute(st); }
    return {st[0], st[1], st[2], st[3]};
}

}  // namespace orc
