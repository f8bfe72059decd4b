// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Radix-2 NTTs over Goldilocks on row-major matrices, and the coset low-degree extension used by
// the reference through p3-dft 0.6.2 `Radix2DitParallel` (un-vendored).  Every function here is
// defined by exact field arithmetic, so any correct algorithm yields identical canonical values;
// correctness is pinned against a naive O(n^2) DFT in tests/test_oracle_ntt.py (mirroring the
// reference's NaiveDft differentials, e.g. crates/lifted-stark/src/prover/quotient.rs:254-265).
#pragma once
#include "field.hpp"
#include "poseidon2_x8.hpp"
#include <cstring>
#include <memory>
#include <new>
#include <cstdio>
#include <cstdlib>
#include <omp.h>

namespace orc {

// Row-major matrix of base-field values.
struct Matrix {
    size_t height = 0, width = 0;
    std::vector<Fp> v;
    Matrix() {}
    Matrix(size_t h, size_t w) : height(h), width(w), v(h * w) {}
    Fp* row(size_t r) { return v.data() + r * width; }
    const Fp* row(size_t r) const { return v.data() + r * width; }
};

inline void bit_reverse_rows(Matrix& m) {
    unsigned lg = log2_strict(m.height);
    // each unordered pair (i, bitrev(i)) is swapped by the iteration with the smaller index only
#pragma omp parallel if (m.height * m.width > (size_t(1) << 16))
    {
        std::vector<Fp> tmp(m.width);
#pragma omp for schedule(static)
        for (size_t i = 0; i < m.height; i++) {
            size_t j = reverse_bits64(i, lg);
            if (i < j) {
                memcpy(tmp.data(), m.row(i), m.width * sizeof(Fp));
                memcpy(m.row(i), m.row(j), m.width * sizeof(Fp));
                memcpy(m.row(j), tmp.data(), m.width * sizeof(Fp));
            }
        }
    }
}

// In-place decimation-in-frequency transform on rows: natural-order input, BIT-REVERSED output.
// out[bitrev(k)] = sum_j in[j] * root^(j k), root a primitive height-th root of unity.
// Reference form (one pass over the whole row-major matrix per stage); kept for small inputs and as the
// definition the column-wise routine below is tested against.
inline void dif_rows_rowmajor(Matrix& m, Fp root) {
    size_t n = m.height, w = m.width;
    if (n <= 1) return;
    unsigned lg = log2_strict(n);
    std::vector<Fp> tw(n / 2);
    tw[0] = Fp::raw(1);
    for (size_t i = 1; i < n / 2; i++) tw[i] = tw[i - 1] * root;
    for (unsigned s = 0; s < lg; s++) {
        size_t half = n >> (s + 1);           // butterfly span
        size_t nblocks = size_t(1) << s;
        for (size_t bj = 0; bj < nblocks * half; bj++) {
            size_t b = bj / half, j = bj % half;
            Fp t = tw[j << s];
            Fp* x = m.row(b * 2 * half + j);
            Fp* y = m.row(b * 2 * half + j + half);
            for (size_t c = 0; c < w; c++) {
                Fp a = x[c], d = y[c];
                x[c] = a + d;
                y[c] = (a - d) * t;
            }
        }
    }
}

// One column (contiguous, length n): the same DIF.  stage_tw[s] holds the `half = n >> (s+1)` twiddles of stage
// s contiguously, so the inner loop is unit-stride in data and twiddles (and runs 8 butterflies per AVX-512 step).
#if ORC_HAVE_X8
ORC_X8_FN inline void dif_stage_x8(Fp* x, Fp* y, const Fp* t, size_t half) {
    for (size_t j = 0; j < half; j += 8) {
        __m512i a = _mm512_loadu_si512((const void*)(x + j)), d = _mm512_loadu_si512((const void*)(y + j));
        __m512i tw = _mm512_loadu_si512((const void*)(t + j));
        _mm512_storeu_si512((void*)(x + j), x8::add(a, d));
        _mm512_storeu_si512((void*)(y + j), x8::mul(x8::sub(a, d), tw));
    }
}
#endif
inline void dif_column_stage(Fp* x, Fp* y, const Fp* t, size_t half, bool vec) {
#if ORC_HAVE_X8
    if (vec && half >= 8) { dif_stage_x8(x, y, t, half); return; }
#endif
    (void)vec;
    for (size_t j = 0; j < half; j++) {
        Fp a = x[j], d = y[j];
        x[j] = a + d;
        y[j] = (a - d) * t[j];
    }
}
// After s stages a DIF splits into 2^s independent transforms of size n >> s, so the stages that span more than
// BLOCK elements stream the whole column and everything below runs block by block while the block sits in cache
// (7 passes over a 2^21-element column instead of 21).
inline void dif_column(Fp* col, size_t n, const std::vector<std::vector<Fp>>& stage_tw, bool vec) {
    const size_t BLOCK = size_t(1) << 14;             // 128 KiB of field elements
    unsigned lg = log2_strict(n), s_split = 0;
    while ((n >> s_split) > BLOCK) s_split++;
    for (unsigned s = 0; s < s_split; s++) {
        size_t half = n >> (s + 1), nblocks = size_t(1) << s;
        for (size_t b = 0; b < nblocks; b++) dif_column_stage(col + b * 2 * half, col + b * 2 * half + half, stage_tw[s].data(), half, vec);
    }
    size_t bsz = n >> s_split;
    for (size_t blk = 0; blk < n / bsz; blk++) {
        Fp* base = col + blk * bsz;
        for (unsigned s = s_split; s < lg; s++) {
            size_t half = n >> (s + 1), nsub = bsz / (2 * half);
            for (size_t b = 0; b < nsub; b++) dif_column_stage(base + b * 2 * half, base + b * 2 * half + half, stage_tw[s].data(), half, vec);
        }
    }
}

// Matrix form: transpose to columns, one cache-resident transform per column (threads over columns), transpose
// back.  Same result as dif_rows_rowmajor (exact field arithmetic); tests/test_oracle.py checks both against the
// naive DFT.
inline void dif_rows(Matrix& m, Fp root) {
    size_t n = m.height, w = m.width;
    if (n <= 1 || w == 0) return;
    if (n * w < (size_t(1) << 12)) { dif_rows_rowmajor(m, root); return; }
    unsigned lg = log2_strict(n);
    std::vector<std::vector<Fp>> stage_tw(lg);
    {
        std::vector<Fp> tw(n / 2);
        const size_t CH = 4096;
#pragma omp parallel for schedule(static)
        for (size_t i0 = 0; i0 < n / 2; i0 += CH) {
            Fp x = fp_pow(root, i0);
            for (size_t i = i0; i < std::min(n / 2, i0 + CH); i++) { tw[i] = x; x = x * root; }
        }
        for (unsigned s = 0; s < lg; s++) {
            size_t half = n >> (s + 1);
            stage_tw[s].resize(half);
#pragma omp parallel for schedule(static) if (half > 4096)
            for (size_t j = 0; j < half; j++) stage_tw[s][j] = tw[j << s];
        }
    }
    const bool vec = x8_available();
    double t0 = omp_get_wtime();
    // uninitialised scratch: every page is first touched by the thread that fills it
    std::unique_ptr<Fp, void (*)(void*)> cols_mem((Fp*)malloc(n * w * sizeof(Fp)), free);
    if (!cols_mem) throw std::bad_alloc();
    Fp* cols = cols_mem.get();
#pragma omp parallel for schedule(static)
    for (size_t r0 = 0; r0 < n; r0 += 64)
        for (size_t c = 0; c < w; c++)
            for (size_t r = r0; r < std::min(n, r0 + 64); r++) cols[c * n + r] = m.v[r * w + c];
    double t1 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t c = 0; c < w; c++) dif_column(cols + c * n, n, stage_tw, vec);
    double t2 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (size_t r0 = 0; r0 < n; r0 += 64)
        for (size_t c = 0; c < w; c++)
            for (size_t r = r0; r < std::min(n, r0 + 64); r++) m.v[r * w + c] = cols[c * n + r];
    if (getenv("ORC_PROFILE_NTT")) fprintf(stderr, "[dif_rows %zux%zu] tw+alloc+transpose %.0f ms, columns %.0f ms, transpose back %.0f ms\n", n, w, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (omp_get_wtime() - t2) * 1e3);
}

// Forward DFT, natural in -> natural out:  out[k] = sum_j in[j] * omega_n^(j k).
inline void dft_rows(Matrix& m) {
    dif_rows(m, two_adic_generator(log2_strict(m.height)));
    bit_reverse_rows(m);
}
// Inverse DFT, natural in -> natural out.
inline void idft_rows(Matrix& m) {
    unsigned lg = log2_strict(m.height);
    dif_rows(m, fp_inv(two_adic_generator(lg)));
    bit_reverse_rows(m);
    Fp ninv = fp_inv(Fp::raw((u64)m.height));
#pragma omp parallel for schedule(static) if (m.v.size() > (1u << 16))
    for (size_t i = 0; i < m.v.size(); i++) m.v[i] = m.v[i] * ninv;
}

// p3 `coset_lde_batch(mat, added_bits, shift)` as called at
// crates/lifted-stark/src/prover/commit.rs:173: interpret each column as evaluations over the
// subgroup H (natural order), interpolate, and evaluate on the coset shift*K, |K| = |H| << added_bits.
// Output: row-major, rows in BIT-REVERSED order of the coset index (commit.rs:118-119).
inline Matrix coset_lde_bitrev(const Matrix& evals, unsigned added_bits, Fp shift) {
    Matrix c = evals;
    idft_rows(c);
    size_t n = c.height, w = c.width, big = n << added_bits;
    Matrix out(big, w);
#pragma omp parallel for schedule(static)
    for (size_t k0 = 0; k0 < n; k0 += 1024) {
        Fp sp = fp_pow(shift, k0);
        for (size_t k = k0; k < std::min(n, k0 + 1024); k++) {
            for (size_t j = 0; j < w; j++) out.row(k)[j] = c.row(k)[j] * sp;
            sp = sp * shift;
        }
    }
    dif_rows(out, two_adic_generator(log2_strict(big)));   // leaves rows bit-reversed
    return out;
}

// Naive O(n^2) evaluation of a coefficient column at point x (test helper).
inline Fp horner_eval(const std::vector<Fp>& coeffs, Fp x) {
    Fp acc;
    for (size_t i = coeffs.size(); i-- > 0;) acc = acc * x + coeffs[i];
    return acc;
}

}  // namespace orc
