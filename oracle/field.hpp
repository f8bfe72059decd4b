// TEST INFRASTRUCTURE ONLY -- CPU oracle for the Miden STARK proving path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or call anything under oracle/.  The product (miden-vm_b200/) never includes it.
//
// Parity status: the reference's arithmetic lives in un-vendored Plonky3 crates (p3-goldilocks,
// p3-field 0.6.2; reference Cargo.lock:2950-3365) and no Rust toolchain exists in this image, so
// this is a restatement.  Pinned here: modulus / generator / two-adic root (reference
// crates/lib/core/asm/stark/constants.masm:5 ROOT_UNITY = 1753635133440165772, g = 7,
// TWO_ADICITY = 32: random_coin.masm:426-440), extension u^2 = 7
// (air/src/constraints/ext_field.rs:11-12).
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <cassert>

namespace orc {

using u64 = uint64_t;
using u128 = unsigned __int128;

// Goldilocks prime p = 2^64 - 2^32 + 1 (reference crates/field/src/native/mod.rs:58: Felt is a
// transparent wrapper over p3 Goldilocks).  Values are kept canonical (< p) at all times.
constexpr u64 P = 0xFFFFFFFF00000001ULL;

struct Fp {
    u64 v;
    constexpr Fp() : v(0) {}
    constexpr explicit Fp(u64 x) : v(x % P) {}
    static constexpr Fp raw(u64 x) { Fp r; r.v = x; return r; }
    bool operator==(const Fp& o) const { return v == o.v; }
    bool operator!=(const Fp& o) const { return v != o.v; }
    bool is_zero() const { return v == 0; }
};

// branch-free forms (the compiler emits cmov/sbb): the proving path is dominated by these
inline Fp operator+(Fp a, Fp b) {
    u64 s;
    bool c = __builtin_add_overflow(a.v, b.v, &s);
    u64 t = s - P;                       // valid when the add overflowed or s >= P
    return Fp::raw((c | (s >= P)) ? t : s);
}
inline Fp operator-(Fp a, Fp b) {
    u64 d;
    bool bw = __builtin_sub_overflow(a.v, b.v, &d);
    return Fp::raw(bw ? d + P : d);
}
inline Fp operator-(Fp a) { return Fp::raw(a.v ? P - a.v : 0); }
// 128-bit product reduced with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p).
inline u64 reduce128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    u64 t;
    bool bw = __builtin_sub_overflow(lo, hh, &t);
    t -= bw ? 0xFFFFFFFFULL : 0;              // borrow: add p back (== subtract 2^32 - 1 mod 2^64)
    u64 m = hl * 0xFFFFFFFFULL;               // hl * (2^32 - 1) < 2^64
    u64 r;
    bool cy = __builtin_add_overflow(t, m, &r);
    r += cy ? 0xFFFFFFFFULL : 0;              // carry: 2^64 = 2^32 - 1
    return r >= P ? r - P : r;
}
inline Fp operator*(Fp a, Fp b) { return Fp::raw(reduce128((u128)a.v * b.v)); }
inline Fp& operator+=(Fp& a, Fp b) { a = a + b; return a; }
inline Fp& operator-=(Fp& a, Fp b) { a = a - b; return a; }
inline Fp& operator*=(Fp& a, Fp b) { a = a * b; return a; }

inline Fp fp_pow(Fp b, u64 e) {
    Fp r = Fp::raw(1);
    while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
    return r;
}
inline Fp fp_inv(Fp a) { assert(a.v != 0); return fp_pow(a, P - 2); }
inline Fp fp_exp_pow2(Fp a, unsigned k) { while (k--) a = a * a; return a; }

constexpr u64 GENERATOR = 7;        // multiplicative generator of F_p^*
constexpr unsigned TWO_ADICITY = 32;
constexpr u64 ROOT_UNITY_2_32 = 1753635133440165772ULL;  // = 7^((p-1)/2^32), constants.masm:5

// Primitive 2^bits-th root of unity (p3 `two_adic_generator`): successive squarings of the
// order-2^32 root.
inline Fp two_adic_generator(unsigned bits) {
    assert(bits <= TWO_ADICITY);
    return fp_exp_pow2(Fp::raw(ROOT_UNITY_2_32), TWO_ADICITY - bits);
}

// Quadratic extension F_p[u]/(u^2 - 7); basis coefficients (c0, c1) in transcript order.
struct Ef {
    Fp a, b;
    Ef() {}
    Ef(Fp a_, Fp b_) : a(a_), b(b_) {}
    explicit Ef(Fp a_) : a(a_), b() {}
    bool operator==(const Ef& o) const { return a == o.a && b == o.b; }
    bool operator!=(const Ef& o) const { return !(*this == o); }
    bool is_zero() const { return a.is_zero() && b.is_zero(); }
};
inline Ef operator+(Ef x, Ef y) { return Ef(x.a + y.a, x.b + y.b); }
inline Ef operator-(Ef x, Ef y) { return Ef(x.a - y.a, x.b - y.b); }
inline Ef operator-(Ef x) { return Ef(-x.a, -x.b); }
inline Ef operator*(Ef x, Ef y) {
    return Ef(x.a * y.a + Fp::raw(7) * (x.b * y.b), x.a * y.b + x.b * y.a);
}
inline Ef operator*(Ef x, Fp s) { return Ef(x.a * s, x.b * s); }
inline Ef operator-(Ef x, Fp s) { return Ef(x.a - s, x.b); }
inline Ef operator+(Ef x, Fp s) { return Ef(x.a + s, x.b); }
inline Ef& operator+=(Ef& x, Ef y) { x = x + y; return x; }
inline Ef& operator*=(Ef& x, Ef y) { x = x * y; return x; }
inline Ef ef_one() { return Ef(Fp::raw(1), Fp()); }
inline Ef ef_inv(Ef x) {
    // 1/(a + b u) = (a - b u) / (a^2 - 7 b^2)
    Fp n = x.a * x.a - Fp::raw(7) * (x.b * x.b);
    Fp ni = fp_inv(n);
    return Ef(x.a * ni, -(x.b * ni));
}
inline Ef ef_pow(Ef b, u64 e) {
    Ef r = ef_one();
    while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
    return r;
}
inline Ef ef_exp_pow2(Ef a, unsigned k) { while (k--) a = a * a; return a; }

inline unsigned reverse_bits(u64 x, unsigned bits) {
    u64 r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | ((x >> i) & 1); }
    return (unsigned)r;
}
inline u64 reverse_bits64(u64 x, unsigned bits) {
    u64 r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | ((x >> i) & 1); }
    return r;
}
inline unsigned log2_strict(size_t n) {
    unsigned l = 0;
    while ((size_t(1) << l) < n) l++;
    assert((size_t(1) << l) == n);
    return l;
}

// Montgomery batch inversion (p3 `batch_multiplicative_inverse`; exact, so any method agrees).
inline void batch_inverse(std::vector<Fp>& v) {
    size_t n = v.size();
    if (!n) return;
    std::vector<Fp> pre(n);
    Fp acc = Fp::raw(1);
    for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = acc * v[i]; }
    Fp inv = fp_inv(acc);
    for (size_t i = n; i-- > 0;) { Fp t = v[i]; v[i] = inv * pre[i]; inv = inv * t; }
}
inline void batch_inverse(std::vector<Ef>& v) {
    size_t n = v.size();
    if (!n) return;
    std::vector<Ef> pre(n);
    Ef acc = ef_one();
    for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = acc * v[i]; }
    Ef inv = ef_inv(acc);
    for (size_t i = n; i-- > 0;) { Ef t = v[i]; v[i] = inv * pre[i]; inv = inv * t; }
}

}  // namespace orc
