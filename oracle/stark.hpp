// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Lifted-STARK prover and verifier, CPU restatement of crates/lifted-stark/src:
//   prover/mod.rs:230-578      prove()                       -> stark_prove()
//   prover/commit.rs:142-180   commit_traces                 -> commit_traces()
//   prover/constraints/mod.rs:83-278, domain.rs:698-749      -> eval_quotient_numerators()
//   prover/quotient.rs:45-217  upsample / accumulate / commit_quotient
//   pcs/prover.rs:34-102       open_with_channel             -> pcs_open()
//   pcs/deep/prover.rs:54-315, pcs/deep/interpolate.rs       -> DEEP section of pcs_open()
//   pcs/fri/prover.rs:93-269, pcs/fri/fold/arity4.rs:46-121  -> FRI section of pcs_open()
//   verifier/mod.rs:153-…, pcs/verifier.rs:72-125, pcs/deep/verifier.rs, pcs/fri/verifier.rs
//                                                            -> stark_verify()
// PARITY UNPINNED: the reference holds no golden proofs/roots for this path and
// cannot be built here; this restatement is pinned only at the Poseidon2 KAT and field
// constants, by NaiveDft-style differentials, and by its own prove->verify round trip (the
// reference's parity mechanism, crates/lifted-stark/src/testing/configs/goldilocks_poseidon2.rs:143-166).
#pragma once
#include "lmcs.hpp"
#include "air.hpp"
#include <functional>
#include <numeric>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace orc {

// optional phase timing (ORC_PROFILE=1)
struct OrcTimer {
    const char* name; std::chrono::steady_clock::time_point t0; bool on;
    OrcTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()), on(getenv("ORC_PROFILE") != nullptr) {}
    ~OrcTimer() { if (on) fprintf(stderr, "[oracle] %-28s %8.1f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); }
};
#define ORC_TIMER(n) OrcTimer orc_timer_##__LINE__(n)

struct PcsParams {   // crates/lifted-stark/src/pcs/params.rs:35-99; Miden values air/src/config.rs:55-67
    unsigned log_blowup = 3, log_folding_arity = 2, log_final_degree = 7;
    unsigned folding_pow_bits = 4, deep_pow_bits = 12, num_queries = 27, query_pow_bits = 16;
};

struct AirDesc {
    size_t width = 0;            // main trace width
    size_t aux_width = 0;        // aux width in EF columns
    size_t num_aux_values = 0;
    size_t num_randomness = 0;
    unsigned log_quotient_degree = 0;   // domain.rs:585-598 (symbolic degree analysis is host-side)
    AirProgram program;
    // periodic_columns_matrix(): max_period x n_periodic, row-major (prover/periodic.rs:49-98)
    std::vector<Fp> periodic; size_t n_periodic = 0; unsigned log_max_period = 0;
    size_t preprocessed_width = 0;
    LookupProgram lookup;            // optional: aux trace built from the lowered LookupAir (build_logup_aux_trace)       // BaseAir::preprocessed_width
    // coefficients (ascending) of periodic column c over the size-max_period subgroup
    std::vector<std::vector<Fp>> periodic_coeffs() const {
        size_t mp = size_t(1) << log_max_period;
        std::vector<std::vector<Fp>> out;
        for (size_t c = 0; c < n_periodic; c++) {
            Matrix m(mp, 1);
            for (size_t r = 0; r < mp; r++) m.row(r)[0] = periodic[r * n_periodic + c];
            idft_rows(m);
            out.push_back(m.v);
        }
        return out;
    }
};

// (aux trace as EF row-major flattened to base: N x 2*aux_width, aux values)
using AuxBuilder = std::function<void(size_t instance, const Matrix& main, const std::vector<Ef>& randomness,
                                      Matrix& aux_flat, std::vector<Ef>& aux_values)>;

struct Statement {
    std::vector<AirDesc> airs;           // instance order
    std::vector<Fp> public_values;       // shared air_inputs
    std::vector<Fp> observe_felts;       // what `Statement::observe` absorbs (host/AIR specific)
    // BaseAir::preprocessed_trace per AIR (instance order; empty matrix = none): setup-fixed columns
    // committed once (preprocessed.rs:41-110) and observed before everything else (prover/mod.rs:282-286)
    std::vector<Matrix> preprocessed;
    bool has_preprocessed() const { for (auto& m : preprocessed) if (m.width) return true; return false; }
};



struct Proof {
    std::vector<uint8_t> log_trace_heights;   // instance order (proof.rs:57-63)
    std::vector<Fp> fields;
    std::vector<Digest> commitments;
};

// ---------------------------------------------------------------------------------------------
// Domain helpers (domain.rs)
// ---------------------------------------------------------------------------------------------
inline Fp lde_shift(unsigned log_lde) { return fp_exp_pow2(Fp::raw(GENERATOR), TWO_ADICITY - log_lde); }  // :358-361

struct TraceOrder {   // order.rs: stable sort on (log_height, instance index)
    std::vector<unsigned> log_heights;       // instance order
    std::vector<size_t> proof_to_instance;   // proof position -> instance index
    static TraceOrder make(const std::vector<unsigned>& lh) {
        TraceOrder t; t.log_heights = lh;
        t.proof_to_instance.resize(lh.size());
        std::iota(t.proof_to_instance.begin(), t.proof_to_instance.end(), 0);
        std::stable_sort(t.proof_to_instance.begin(), t.proof_to_instance.end(),
                         [&](size_t a, size_t b) { return lh[a] < lh[b]; });
        return t;
    }
    unsigned max_log_height() const { return log_heights[proof_to_instance.back()]; }
};

inline unsigned fri_num_rounds(const PcsParams& p, unsigned log_lde) {   // pcs/fri/mod.rs:80-94
    unsigned target = p.log_final_degree + p.log_blowup;
    unsigned steps = log_lde > target ? log_lde - target : 0;
    return (steps + p.log_folding_arity - 1) / p.log_folding_arity;
}
inline size_t fri_final_poly_degree(const PcsParams& p, unsigned log_lde) {   // :105-115
    unsigned r = fri_num_rounds(p, log_lde);
    unsigned lf = log_lde > r * p.log_folding_arity ? log_lde - r * p.log_folding_arity : 0;
    unsigned ld = lf > p.log_blowup ? lf - p.log_blowup : 0;
    return size_t(1) << ld;
}

// ---------------------------------------------------------------------------------------------
// Commit (commit.rs:142-180)
// ---------------------------------------------------------------------------------------------
inline LmcsTree commit_traces(const std::vector<Matrix>& traces_proof_order, unsigned log_blowup) {
    std::vector<Matrix> ldes;
    for (const Matrix& t : traces_proof_order) {
        ORC_TIMER("  coset lde");
        unsigned lh = log2_strict(t.height);
        ldes.push_back(coset_lde_bitrev(t, log_blowup, lde_shift(lh + log_blowup)));
    }
    ORC_TIMER("  lmcs build");
    return LmcsTree::build(std::move(ldes), lmcs_alignment());   // build_aligned_tree: alignment = RATE (8) for the sponge, 1 for the chaining hasher
}

// Preprocessed::build (preprocessed.rs:41-110): AIRs with preprocessed columns sorted by (height, air index),
// LDE by the blowup on the canonical coset, aligned LMCS tree.  `air_of[i]` = AIR backing committed trace i.
struct PreprocessedBundle { LmcsTree tree; std::vector<size_t> air_of; std::vector<Matrix> traces; };
inline PreprocessedBundle build_preprocessed(const Statement& st, unsigned log_blowup) {
    PreprocessedBundle b;
    for (size_t i = 0; i < st.preprocessed.size(); i++) if (st.preprocessed[i].width) b.air_of.push_back(i);
    std::stable_sort(b.air_of.begin(), b.air_of.end(), [&](size_t x, size_t y) { return st.preprocessed[x].height < st.preprocessed[y].height; });
    for (size_t a : b.air_of) b.traces.push_back(st.preprocessed[a]);
    b.tree = commit_traces(b.traces, log_blowup);
    return b;
}

// ---------------------------------------------------------------------------------------------
// Constraint evaluation on gJ_j, natural order (constraints/mod.rs:83-278)
// ---------------------------------------------------------------------------------------------
inline std::vector<Ef> eval_quotient(const AirDesc& air, const Matrix& main_lde, const Matrix& aux_lde, const Matrix* prep_lde,
                                     unsigned log_n, unsigned log_blowup, Ef alpha,
                                     const std::vector<Ef>& randomness, const std::vector<Fp>& publics,
                                     const std::vector<Ef>& aux_values) {
    unsigned log_d = air.log_quotient_degree;
    size_t D = size_t(1) << log_d, n = size_t(1) << log_n, gj = n * D;
    unsigned log_gj = log_n + log_d;
    Fp s = lde_shift(log_n + log_blowup);
    // inv_vanishing_evals (domain.rs:742-749)
    Fp s_pow_n = fp_exp_pow2(s, log_n);
    Fp omega_d = two_adic_generator(log_d);
    std::vector<Fp> inv_zh(D), zh(D);
    { Fp x = Fp::raw(1); for (size_t i = 0; i < D; i++) { zh[i] = s_pow_n * x - Fp::raw(1); inv_zh[i] = zh[i]; x = x * omega_d; } }
    batch_inverse(inv_zh);
    Fp omega_j = two_adic_generator(log_gj);
    Fp omega_h_inv = fp_inv(two_adic_generator(log_n));
    // selectors (domain.rs:698-734)
    std::vector<Fp> xs(gj), d_first(gj), d_last(gj);
    { Fp x = s; for (size_t i = 0; i < gj; i++) { xs[i] = x; d_first[i] = x - Fp::raw(1); d_last[i] = x - omega_h_inv; x = x * omega_j; } }
    batch_inverse(d_first); batch_inverse(d_last);
    // The committed LDE (bit-reversed rows, height n*B) holds gJ as its first gj rows
    // (commit.rs:95-106): natural row i of gJ = physical row bitrev_{log_gj}(i).
    std::vector<Ef> out(gj);
    auto pcoef = air.periodic_coeffs();
#pragma omp parallel if (gj > 1024)
    {
        std::vector<Ef> scratch;
        std::vector<Ef> per(air.n_periodic);
#pragma omp for schedule(static)
        for (size_t i = 0; i < gj; i++) {
            size_t inext = (i + D) & (gj - 1);
            size_t pl = reverse_bits64(i, log_gj), pn = reverse_bits64(inext, log_gj);
            AirPoint pt{};
            pt.main_local = main_lde.row(pl); pt.main_next = main_lde.row(pn);
            pt.aux_local = aux_lde.width ? aux_lde.row(pl) : nullptr;
            pt.aux_next = aux_lde.width ? aux_lde.row(pn) : nullptr;
            if (prep_lde) { pt.prep_local = prep_lde->row(pl); pt.prep_next = prep_lde->row(pn); }
            pt.publics = publics.data(); pt.challenges = randomness.data(); pt.aux_values = aux_values.data();
            Fp z = zh[i & (D - 1)];
            pt.is_first = Ef(z * d_first[i]); pt.is_last = Ef(z * d_last[i]); pt.is_transition = Ef(xs[i] - omega_h_inv);
            if (air.n_periodic) {
                Fp y = fp_exp_pow2(xs[i], log_n - air.log_max_period);
                for (size_t c = 0; c < air.n_periodic; c++) per[c] = Ef(horner_eval(pcoef[c], y));
                pt.periodic = per.data();
            }
            Ef folded = air_eval_folded(air.program, pt, alpha, scratch);
            out[i] = folded * inv_zh[i & (D - 1)];
        }
    }
    return out;
}

// quotient.rs:45-56: same polynomial, same coset shift, larger subgroup (natural order).
inline std::vector<Ef> upsample_evals(const std::vector<Ef>& evals, unsigned added_bits) {
    if (!added_bits) return evals;
    size_t n = evals.size(), big = n << added_bits;
    Matrix m(n, 2);
    for (size_t i = 0; i < n; i++) { m.row(i)[0] = evals[i].a; m.row(i)[1] = evals[i].b; }
    idft_rows(m);
    Matrix o(big, 2);
    for (size_t i = 0; i < n; i++) { o.row(i)[0] = m.row(i)[0]; o.row(i)[1] = m.row(i)[1]; }
    dft_rows(o);
    std::vector<Ef> r(big);
    for (size_t i = 0; i < big; i++) r[i] = Ef(o.row(i)[0], o.row(i)[1]);
    return r;
}

// quotient.rs:78-111
inline void cyclic_extend_and_accumulate(std::vector<Ef>& acc, const std::vector<Ef>& contrib, Ef beta) {
    if (acc.empty()) { acc = contrib; return; }
    size_t old = acc.size();
    std::vector<Ef> next(contrib.size());
    for (size_t i = 0; i < contrib.size(); i++) next[i] = acc[i % old] * beta + contrib[i];
    acc = std::move(next);
}

// quotient.rs:143-217.  q_evals: N*D EF values, natural order on gJ.  Returns the LDE matrix
// (L x 2D base columns, bit-reversed rows) whose column pair t holds chunk q_t on gK.
inline Matrix quotient_chunk_lde(const std::vector<Ef>& q_evals, unsigned log_n, unsigned log_d, unsigned log_blowup) {
    size_t n = size_t(1) << log_n, D = size_t(1) << log_d, L = n << log_blowup;
    Matrix c(n, 2 * D);
    for (size_t r = 0; r < n; r++)
        for (size_t t = 0; t < D; t++) { c.row(r)[2 * t] = q_evals[r * D + t].a; c.row(r)[2 * t + 1] = q_evals[r * D + t].b; }
    idft_rows(c);
    Fp omega_j_inv = fp_inv(two_adic_generator(log_n + log_d));
    Matrix out(L, 2 * D);
    Fp row_base = Fp::raw(1);   // omega_J^{-k}
    for (size_t k = 0; k < n; k++) {
        Fp scale = Fp::raw(1);
        for (size_t t = 0; t < D; t++) {
            out.row(k)[2 * t] = c.row(k)[2 * t] * scale;
            out.row(k)[2 * t + 1] = c.row(k)[2 * t + 1] * scale;
            scale = scale * row_base;
        }
        row_base = row_base * omega_j_inv;
    }
    dif_rows(out, two_adic_generator(log2_strict(L)));   // plain DFT, bit-reversed rows
    return out;
}

// ---------------------------------------------------------------------------------------------
// FRI arity-4 fold of one physical row [y0, y2, y1, y3] (fold/arity4.rs:46-121)
// ---------------------------------------------------------------------------------------------
inline Ef fold4(const Ef* ev, Fp s_inv, Ef beta) {
    Fp w = two_adic_generator(2);
    Ef y0 = ev[0], y2 = ev[1], y1 = ev[2], y3 = ev[3];
    Ef s02 = y0 + y2, d02 = y0 - y2, s13 = y1 + y3, d31 = (y3 - y1) * w;
    Ef c0 = s02 + s13, c1 = d02 + d31, c2 = s02 - s13, c3 = d02 - d31;
    Ef x = beta * s_inv;
    Ef x2 = x * x, x3 = x2 * x;
    Ef sum = c0 + c1 * x + c2 * x2 + c3 * x3;
    Fp four_inv = fp_inv(Fp::raw(4));
    return sum * four_inv;
}
inline Ef fold2(const Ef* ev, Fp s_inv, Ef beta) {   // fold/arity2.rs: (y0+y1)/2 + beta/s (y0-y1)/2
    Fp half = fp_inv(Fp::raw(2));
    Ef s = (ev[0] + ev[1]) * half, d = (ev[0] - ev[1]) * half;
    return s + d * (beta * s_inv);
}
// Arity 8 (fold/arity8.rs:35-75): inverse DFT of the eight coset evaluations, evaluate at beta/s, divide by 8.
// The physical row is in bit-reversed order [y0, y4, y2, y6, y1, y5, y3, y7].
inline Ef fold8(const Ef* ev, Fp s_inv, Ef beta) {
    Ef y[8];
    for (unsigned j = 0; j < 8; j++) y[reverse_bits(j, 3)] = ev[j];
    Fp w_inv = fp_inv(two_adic_generator(3));
    Ef x = beta * s_inv, xp = ef_one(), acc;
    for (unsigned m = 0; m < 8; m++) {
        Ef c;
        for (unsigned k = 0; k < 8; k++) c = c + y[k] * fp_pow(w_inv, (u64)m * k);
        acc = acc + c * xp;
        xp = xp * x;
    }
    return acc * fp_inv(Fp::raw(8));
}
inline Ef fold_row(unsigned log_arity, const Ef* ev, Fp s_inv, Ef beta) {
    if (log_arity == 2) return fold4(ev, s_inv, beta);
    if (log_arity == 1) return fold2(ev, s_inv, beta);
    if (log_arity == 3) return fold8(ev, s_inv, beta);
    throw std::runtime_error("fri: unsupported arity");
}

// ---------------------------------------------------------------------------------------------
// PCS open (pcs/prover.rs:34-102)
// ---------------------------------------------------------------------------------------------
struct OpenDebug {   // intermediate values exported for stage-by-stage parity tests
    std::vector<Ef> ood_evals[2];      // aligned, flat, per point
    std::vector<Ef> deep_evals;        // bit-reversed
    std::vector<Digest> fri_roots;
    std::vector<Ef> fri_betas;
    std::vector<Ef> final_poly;
    std::vector<size_t> query_indices;
    Ef deep_alpha, deep_beta;
};

// coeff_mats[g][m]: coefficient matrix (natural order, height = trace height of that matrix) of
// each committed matrix, so that column c evaluates to sum_k coeff[k][c] * (x/shift_m)^k... we
// store plain coefficients of the polynomial f with f(shift*omega^i) = LDE value; evaluation at
// the lifted point is f(z^r).
inline void pcs_open(const PcsParams& params, unsigned log_max_n, Ef z, Ef z_next,
                     const std::vector<const LmcsTree*>& trees,
                     const std::vector<std::vector<Matrix>>& coeffs,   // [tree][matrix] plain coefficients
                     ProverTranscript& ch, OpenDebug* dbg) {
    unsigned log_lde = log_max_n + params.log_blowup;
    size_t L = size_t(1) << log_lde;
    Ef pts[2] = {z, z_next};

    // --- OOD evaluations: column c of matrix m (height n_m) evaluated at z^(N_max/n_m)
    //     (deep/interpolate.rs:127-203; value-equivalent to the barycentric route).
    std::vector<Ef> flat[2];
    std::vector<size_t> aligned_widths, widths;
    for (size_t g = 0; g < trees.size(); g++)
        for (size_t m = 0; m < trees[g]->leaves.size(); m++) {
            const Matrix& cm = coeffs[g][m];
            unsigned lr = log_max_n - log2_strict(cm.height);
            size_t w = cm.width, aw = aligned_len(w, lmcs_alignment());
            widths.push_back(w); aligned_widths.push_back(aw);
            for (int p = 0; p < 2; p++) {
                Ef x = ef_exp_pow2(pts[p], lr);
                // Horner in row chunks (threads), recombined with x^(chunk start): sum_k row_k x^k
                const size_t CH = 4096, nch = (cm.height + CH - 1) / CH;
                std::vector<std::vector<Ef>> part(nch, std::vector<Ef>(w));
#pragma omp parallel for schedule(static) if (nch > 1)
                for (size_t t = 0; t < nch; t++) {
                    std::vector<Ef>& a = part[t];
                    size_t k1 = std::min(cm.height, (t + 1) * CH);
                    for (size_t k = k1; k-- > t * CH;) {
                        const Fp* row = cm.row(k);
                        for (size_t c = 0; c < w; c++) a[c] = a[c] * x + row[c];
                    }
                }
                std::vector<Ef> acc(w);
                Ef xch = ef_one();
                for (size_t i = 0; i < CH && i < cm.height; i++) xch = xch * x;   // x^CH
                for (size_t t = nch; t-- > 0;)
                    for (size_t c = 0; c < w; c++) acc[c] = acc[c] * xch + part[t][c];
                for (size_t c = 0; c < w; c++) flat[p].push_back(acc[c]);
                for (size_t c = w; c < aw; c++) flat[p].push_back(Ef());
            }
        }
    for (int p = 0; p < 2; p++) for (const Ef& e : flat[p]) ch.send_ext(e);   // deep/prover.rs:150-154
    ch.grind(params.deep_pow_bits);                                          // :157-158
    Ef alpha = ch.sample_ext(), beta = ch.sample_ext();                      // :161-162
    size_t W = flat[0].size();
    Ef f_red_at[2];
    for (int p = 0; p < 2; p++) { Ef a; for (size_t i = 0; i < W; i++) a = a * alpha + flat[p][i]; f_red_at[p] = a; }
    // alpha powers: column i of the aligned index space gets alpha^(W-1-i)  (:195-212)
    std::vector<Ef> apow(W);
    { Ef a = ef_one(); for (size_t i = W; i-- > 0;) { apow[i] = a; a = a * alpha; } }

    // --- DEEP quotient over the LDE domain, bit-reversed order (:214-312)
    std::vector<Fp> xs(L);
    { Fp s = lde_shift(log_lde), w = two_adic_generator(log_lde), x = s; std::vector<Fp> nat(L);
      for (size_t i = 0; i < L; i++) { nat[i] = x; x = x * w; }
      for (size_t i = 0; i < L; i++) xs[i] = nat[reverse_bits64(i, log_lde)]; }
    std::vector<Ef> inv0(L), inv1(L);
    for (size_t i = 0; i < L; i++) { inv0[i] = z - xs[i]; inv1[i] = z_next - xs[i]; }
    batch_inverse(inv0); batch_inverse(inv1);
    std::vector<Ef> deep(L);
#pragma omp parallel for schedule(static) if (L > 1024)
    for (size_t i = 0; i < L; i++) {
        Ef fr;   // f_reduced(x_i)
        size_t off = 0, mi = 0;
        for (size_t g = 0; g < trees.size(); g++)
            for (const Matrix& m : trees[g]->leaves) {
                unsigned sc = log_lde - log2_strict(m.height);
                const Fp* row = m.row(i >> sc);
                for (size_t c = 0; c < m.width; c++) fr = fr + apow[off + c] * row[c];
                off += aligned_widths[mi++];
            }
        deep[i] = inv0[i] * (f_red_at[0] - fr) + beta * (inv1[i] * (f_red_at[1] - fr));
    }
    if (dbg) { dbg->ood_evals[0] = flat[0]; dbg->ood_evals[1] = flat[1]; dbg->deep_evals = deep; dbg->deep_alpha = alpha; dbg->deep_beta = beta; }

    // --- FRI commit phase (fri/prover.rs:93-242)
    unsigned la = params.log_folding_arity;
    size_t arity = size_t(1) << la;
    size_t final_deg = fri_final_poly_degree(params, log_lde);
    size_t final_domain = final_deg << params.log_blowup;
    unsigned log_dom = log_lde;
    std::vector<LmcsTree> fri_trees;
    std::vector<Ef> cur = std::move(deep);
    while ((size_t(1) << log_dom) > final_domain) {
        size_t rows = cur.size() / arity;
        Matrix m(rows, 2 * arity);   // physical row k = cur[k*arity .. ], flattened EF -> F
        for (size_t k = 0; k < rows; k++)
            for (size_t j = 0; j < arity; j++) { m.row(k)[2 * j] = cur[k * arity + j].a; m.row(k)[2 * j + 1] = cur[k * arity + j].b; }
        std::vector<Matrix> one; one.push_back(std::move(m));
        fri_trees.push_back(LmcsTree::build(std::move(one), 1));   // build_tree: unaligned (:157-165)
        ch.send_commitment(fri_trees.back().root());
        ch.grind(params.folding_pow_bits);
        Ef b = ch.sample_ext();
        if (dbg) { dbg->fri_roots.push_back(fri_trees.back().root()); dbg->fri_betas.push_back(b); }
        // s_inv[k] = omega_dom^{-bitrev(k, log_dom - la)}  (:137-142)
        unsigned log_rows = log_dom - la;
        Fp ginv = fp_inv(two_adic_generator(log_dom));
        std::vector<Fp> pw(rows);
        { Fp x = Fp::raw(1); for (size_t k = 0; k < rows; k++) { pw[k] = x; x = x * ginv; } }
        std::vector<Ef> next(rows);
#pragma omp parallel for schedule(static) if (rows > 1024)
        for (size_t k = 0; k < rows; k++) next[k] = fold_row(la, &cur[k * arity], pw[reverse_bits64(k, log_rows)], b);
        cur = std::move(next);
        log_dom -= la;
    }
    // final polynomial (:228-239): first final_deg values, bit-reverse, iDFT, descending order
    {
        unsigned lf = log2_strict(final_deg);
        Matrix m(final_deg, 2);
        for (size_t i = 0; i < final_deg; i++) { Ef e = cur[reverse_bits64(i, lf)]; m.row(i)[0] = e.a; m.row(i)[1] = e.b; }
        idft_rows(m);
        std::vector<Ef> fp(final_deg);
        for (size_t i = 0; i < final_deg; i++) fp[i] = Ef(m.row(final_deg - 1 - i)[0], m.row(final_deg - 1 - i)[1]);
        for (const Ef& e : fp) ch.send_ext(e);
        if (dbg) dbg->final_poly = fp;
    }
    // --- queries (pcs/prover.rs:73-101)
    ch.grind(params.query_pow_bits);
    std::vector<size_t> qs;
    for (unsigned q = 0; q < params.num_queries; q++) qs.push_back((size_t)ch.sample_bits(log_lde));
    if (dbg) dbg->query_indices = qs;
    TreeIndices ti = TreeIndices::make(qs, log_lde);
    for (const LmcsTree* t : trees) lmcs_prove_lifted_batch(*t, ti, ch);
    for (const LmcsTree& t : fri_trees) { ti.shrink_depth(la); lmcs_prove_batch(t, ti, ch); }   // fri/prover.rs:254-269
}

// ---------------------------------------------------------------------------------------------
// prove()  (prover/mod.rs:230-578)
// ---------------------------------------------------------------------------------------------
struct ProveDebug {
    Digest main_root, aux_root, quotient_root;
    std::vector<Ef> randomness;
    Ef alpha, beta, z;
    std::vector<Ef> quotient_acc;   // natural order on gJ_max
    OpenDebug open;
};

// `build_logup_aux_trace` (air/src/lookup/aux_builder.rs:49-97) on the lowered LookupAir, with the per-fraction
// semantics of `accumulate_slow` (:215-268): f_c(r) = sum of m/d over the active interactions of column c at
// row r; aux[r][c>0] = f_c(r); aux[r][0] = sum_{r'<r} sum_c f_c(r'); committed final = aux[N][0].
inline void build_logup_aux(const AirDesc& air, const Matrix& main, const std::vector<Ef>& challenges,
                            const std::vector<Fp>& publics, Matrix& aux, std::vector<Ef>& aux_values) {
    const LookupProgram& lp = air.lookup;
    size_t N = main.height, C = lp.num_columns;
    if (air.aux_width != C || air.num_aux_values != 1) throw std::runtime_error("lookup: aux shape mismatch");
    aux = Matrix(N, 2 * C);
    std::vector<Ef> totals(N);
    size_t maxp = size_t(1) << air.log_max_period;
    std::string err;
    #pragma omp parallel for schedule(static)
    for (size_t r = 0; r < N; r++) {
        std::vector<Ef> val(lp.nodes.size());
        const Fp* loc = main.row(r); const Fp* nxt = main.row((r + 1) % N);
        for (size_t i = 0; i < lp.nodes.size(); i++) {
            const AirNode& nd = lp.nodes[i];
            Ef v;
            switch (nd.op) {
                case OP_MAIN: v = Ef((nd.a ? nxt : loc)[nd.b]); break;
                case OP_PUBLIC: v = Ef(publics[nd.a]); break;
                case OP_CHALLENGE: v = challenges[nd.a]; break;
                case OP_CONST: v = Ef(Fp(lp.consts[nd.a])); break;
                case OP_EXT_CONST: v = Ef(Fp(lp.consts[nd.a]), Fp(lp.consts[nd.a + 1])); break;
                case OP_ADD: v = val[nd.a] + val[nd.b]; break;
                case OP_SUB: v = val[nd.a] - val[nd.b]; break;
                case OP_MUL: v = val[nd.a] * val[nd.b]; break;
                case OP_NEG: v = -val[nd.a]; break;
                case OP_PERIODIC: v = Ef(air.periodic[(r % maxp) * air.n_periodic + nd.a]); break;
                default: break;
            }
            val[i] = v;
        }
        std::vector<Ef> f(C);
        for (const LookupInteraction& it : lp.interactions) {
            if (it.flag != LookupProgram::NO_FLAG && val[it.flag].is_zero()) continue;
            Ef d = val[it.denominator];
            if (d.is_zero()) {
                #pragma omp critical
                err = "LogUp denominator must be non-zero";
                continue;
            }
            f[it.column] += ef_inv(d) * val[it.multiplicity].a;
        }
        Ef t;
        for (size_t c = 0; c < C; c++) {
            t += f[c];
            if (c > 0) { aux.row(r)[2 * c] = f[c].a; aux.row(r)[2 * c + 1] = f[c].b; }
        }
        totals[r] = t;
    }
    if (!err.empty()) throw std::runtime_error(err);
    Ef acc;
    for (size_t r = 0; r < N; r++) { aux.row(r)[0] = acc.a; aux.row(r)[1] = acc.b; acc += totals[r]; }
    aux_values.assign(1, acc);
}

inline Proof stark_prove(const PcsParams& params, const Statement& st, const std::vector<Matrix>& traces,
                         Challenger challenger, const AuxBuilder& build_aux, ProveDebug* dbg = nullptr) {
    size_t k = st.airs.size();
    if (traces.size() != k || k == 0) throw std::runtime_error("prove: trace count mismatch");
    std::vector<unsigned> lh(k);
    for (size_t i = 0; i < k; i++) {
        if (traces[i].width != st.airs[i].width) throw std::runtime_error("prove: trace width mismatch");
        lh[i] = log2_strict(traces[i].height);
    }
    TraceOrder ord = TraceOrder::make(lh);
    unsigned log_max_n = ord.max_log_height();
    unsigned lb = params.log_blowup;
    // preprocessed commitment first (mod.rs:282-286), then statement.observe + observe_shape (:290-291)
    PreprocessedBundle prep;
    bool has_prep = st.has_preprocessed();
    if (has_prep) {
        for (size_t i = 0; i < k; i++) {
            bool want = st.airs[i].preprocessed_width > 0, have = i < st.preprocessed.size() && st.preprocessed[i].width > 0;
            if (want != have) throw std::runtime_error("prove: preprocessed presence mismatch");
            if (have && (st.preprocessed[i].width != st.airs[i].preprocessed_width || st.preprocessed[i].height != traces[i].height))
                throw std::runtime_error("prove: preprocessed shape mismatch");
        }
        prep = build_preprocessed(st, lb);
        challenger.observe_digest(prep.tree.root());
    }
    for (Fp f : st.observe_felts) challenger.observe(f);
    challenger.observe(Fp::raw((u64)k));
    for (size_t i = 0; i < k; i++) challenger.observe(Fp::raw((u64)lh[i]));
    ProverTranscript ch; ch.ch = challenger;

    unsigned log_qd = 0;
    for (size_t j = 0; j < k; j++) log_qd = std::max(log_qd, st.airs[ord.proof_to_instance[j]].log_quotient_degree);
    if (log_qd > lb) throw std::runtime_error("prove: constraint degree too high");

    // 1. main commit
    std::vector<Matrix> main_p;
    for (size_t j = 0; j < k; j++) main_p.push_back(traces[ord.proof_to_instance[j]]);
    LmcsTree main_tree = [&] { ORC_TIMER("commit main"); return commit_traces(main_p, lb); }();
    ch.send_commitment(main_tree.root());
    // 2. randomness, aux traces (instance order), aux commit
    size_t max_rand = 0;
    for (auto& a : st.airs) max_rand = std::max(max_rand, a.num_randomness);
    std::vector<Ef> randomness;
    for (size_t i = 0; i < max_rand; i++) randomness.push_back(ch.sample_ext());
    std::vector<Matrix> aux_i(k); std::vector<std::vector<Ef>> auxv_i(k);
    for (size_t i = 0; i < k; i++) {
        std::vector<Ef> r(randomness.begin(), randomness.begin() + st.airs[i].num_randomness);
        aux_i[i] = Matrix(traces[i].height, 2 * st.airs[i].aux_width);
        auxv_i[i].assign(st.airs[i].num_aux_values, Ef());
        if (st.airs[i].lookup.present()) build_logup_aux(st.airs[i], traces[i], r, st.public_values, aux_i[i], auxv_i[i]);
        else if (build_aux) build_aux(i, traces[i], r, aux_i[i], auxv_i[i]);
    }
    std::vector<Matrix> aux_p; std::vector<std::vector<Ef>> auxv_p;
    for (size_t j = 0; j < k; j++) { aux_p.push_back(aux_i[ord.proof_to_instance[j]]); auxv_p.push_back(auxv_i[ord.proof_to_instance[j]]); }
    LmcsTree aux_tree = [&] { ORC_TIMER("commit aux"); return commit_traces(aux_p, lb); }();
    ch.send_commitment(aux_tree.root());
    for (auto& vs : auxv_p) for (const Ef& v : vs) ch.send_ext(v);
    // 3. alpha, beta
    Ef alpha = ch.sample_ext(), beta = ch.sample_ext();
    // 4. constraints -> accumulator
    std::vector<Ef> acc;
    for (size_t j = 0; j < k; j++) {
        ORC_TIMER("constraints");
        const AirDesc& air = st.airs[ord.proof_to_instance[j]];
        unsigned ln = lh[ord.proof_to_instance[j]];
        std::vector<Ef> r(randomness.begin(), randomness.begin() + air.num_randomness);
        const Matrix* pl = nullptr;
        if (has_prep) for (size_t q2 = 0; q2 < prep.air_of.size(); q2++) if (prep.air_of[q2] == ord.proof_to_instance[j]) pl = &prep.tree.leaves[q2];
        std::vector<Ef> q = eval_quotient(air, main_tree.leaves[j], aux_tree.leaves[j], pl, ln, lb, alpha, r, st.public_values, auxv_p[j]);
        q = upsample_evals(q, log_qd - air.log_quotient_degree);
        cyclic_extend_and_accumulate(acc, q, beta);
    }
    // 5. quotient commit
    Matrix qlde = [&] { ORC_TIMER("quotient lde"); return quotient_chunk_lde(acc, log_max_n, log_qd, lb); }();
    std::vector<Matrix> ql; ql.push_back(qlde);
    LmcsTree q_tree = [&] { ORC_TIMER("quotient tree"); return LmcsTree::build(std::move(ql), lmcs_alignment()); }();
    ch.send_commitment(q_tree.root());
    // 6. OOD point (domain.rs:539-552)
    unsigned log_lde = log_max_n + lb;
    Fp sinv = fp_inv(lde_shift(log_lde));
    Ef z;
    for (;;) {
        z = ch.sample_ext();
        if (z.is_zero()) continue;
        if (ef_exp_pow2(z, log_max_n) == ef_one()) continue;
        if (ef_exp_pow2(z * sinv, log_lde) == ef_one()) continue;
        break;
    }
    Ef z_next = z * two_adic_generator(log_max_n);
    // coefficient matrices for OOD evaluation
    OrcTimer* tco = new OrcTimer("ood coefficients");
    std::vector<std::vector<Matrix>> coeffs(3);
    for (size_t j = 0; j < k; j++) { Matrix c = main_p[j]; idft_rows(c); coeffs[0].push_back(std::move(c)); }
    for (size_t j = 0; j < k; j++) { Matrix c = aux_p[j]; idft_rows(c); coeffs[1].push_back(std::move(c)); }
    {   // quotient chunk polynomials: recover plain coefficients from the first N rows?  Simpler:
        // interpolate the committed LDE (bit-reversed) back: un-bitreverse, iDFT over K, divide by g^k.
        Matrix c = q_tree.leaves[0];
        bit_reverse_rows(c);
        idft_rows(c);
        size_t n = size_t(1) << log_max_n;
        Matrix cc(n, c.width);
        Fp gi = sinv, sp = Fp::raw(1);
        for (size_t r = 0; r < n; r++) { for (size_t col = 0; col < c.width; col++) cc.row(r)[col] = c.row(r)[col] * sp; sp = sp * gi; }
        coeffs[2].push_back(std::move(cc));
    }
    delete tco;
    std::vector<const LmcsTree*> trees{&main_tree, &aux_tree, &q_tree};
    if (has_prep) {   // group order [preprocessed, main, aux, quotient] (mod.rs:551-559)
        trees.insert(trees.begin(), &prep.tree);
        std::vector<Matrix> pc;
        for (const Matrix& t : prep.traces) { Matrix c = t; idft_rows(c); pc.push_back(std::move(c)); }
        coeffs.insert(coeffs.begin(), std::move(pc));
    }
    if (dbg) {
        dbg->main_root = main_tree.root(); dbg->aux_root = aux_tree.root(); dbg->quotient_root = q_tree.root();
        dbg->randomness = randomness; dbg->alpha = alpha; dbg->beta = beta; dbg->z = z; dbg->quotient_acc = acc;
    }
    { ORC_TIMER("pcs open"); pcs_open(params, log_max_n, z, z_next, trees, coeffs, ch, dbg ? &dbg->open : nullptr); }
    Proof pf;
    for (unsigned h : lh) pf.log_trace_heights.push_back((uint8_t)h);
    pf.fields = std::move(ch.fields);
    pf.commitments = std::move(ch.commitments);
    return pf;
}

// ---------------------------------------------------------------------------------------------
// verify()  (verifier/mod.rs:153-…)
// ---------------------------------------------------------------------------------------------
inline void stark_verify(const PcsParams& params, const Statement& st, const Proof& pf, Challenger challenger,
                         const Digest* preprocessed_commitment = nullptr) {
    size_t k = st.airs.size();
    if (pf.log_trace_heights.size() != k) throw std::runtime_error("verify: height count");
    std::vector<unsigned> lh(pf.log_trace_heights.begin(), pf.log_trace_heights.end());
    TraceOrder ord = TraceOrder::make(lh);
    unsigned log_max_n = ord.max_log_height(), lb = params.log_blowup, log_lde = log_max_n + lb;
    bool has_prep = false;
    for (auto& a : st.airs) has_prep |= a.preprocessed_width > 0;
    if (has_prep != (preprocessed_commitment != nullptr)) throw std::runtime_error("verify: preprocessed presence mismatch");
    if (has_prep) challenger.observe_digest(*preprocessed_commitment);
    for (Fp f : st.observe_felts) challenger.observe(f);
    challenger.observe(Fp::raw((u64)k));
    for (size_t i = 0; i < k; i++) challenger.observe(Fp::raw((u64)lh[i]));
    VerifierTranscript ch{challenger, pf.fields.data(), pf.fields.size(), 0, pf.commitments.data(), pf.commitments.size(), 0};

    unsigned log_qd = 0;
    for (auto& a : st.airs) log_qd = std::max(log_qd, a.log_quotient_degree);
    size_t D = size_t(1) << log_qd;
    Digest main_root = ch.receive_commitment();
    size_t max_rand = 0;
    for (auto& a : st.airs) max_rand = std::max(max_rand, a.num_randomness);
    std::vector<Ef> randomness;
    for (size_t i = 0; i < max_rand; i++) randomness.push_back(ch.sample_ext());
    Digest aux_root = ch.receive_commitment();
    std::vector<std::vector<Ef>> auxv(k);
    for (size_t j = 0; j < k; j++) {
        const AirDesc& air = st.airs[ord.proof_to_instance[j]];
        for (size_t i = 0; i < air.num_aux_values; i++) auxv[j].push_back(ch.receive_ext());
    }
    Ef alpha = ch.sample_ext(), beta = ch.sample_ext();
    Digest q_root = ch.receive_commitment();
    Fp shift = lde_shift(log_lde), sinv = fp_inv(shift);
    Ef z;
    for (;;) {
        z = ch.sample_ext();
        if (z.is_zero() || ef_exp_pow2(z, log_max_n) == ef_one() || ef_exp_pow2(z * sinv, log_lde) == ef_one()) continue;
        break;
    }
    Fp omega_h = two_adic_generator(log_max_n);
    Ef pts[2] = {z, z * omega_h};

    struct Group { Digest root; std::vector<size_t> widths, aligned; unsigned log_height; };
    std::vector<Group> groups(3);
    groups[0].root = main_root; groups[1].root = aux_root; groups[2].root = q_root;
    for (auto& g : groups) g.log_height = log_lde;
    for (size_t j = 0; j < k; j++) {
        const AirDesc& air = st.airs[ord.proof_to_instance[j]];
        groups[0].widths.push_back(air.width); groups[1].widths.push_back(2 * air.aux_width);
    }
    groups[2].widths.push_back(2 * D);
    std::vector<size_t> prep_pos(k, (size_t)-1);   // proof position -> index in the preprocessed group
    if (has_prep) {   // committed preprocessed traces = preprocessed AIRs in proof order (verifier/mod.rs)
        Group pg; pg.root = *preprocessed_commitment; pg.log_height = 0;
        for (size_t j = 0; j < k; j++) {
            const AirDesc& air = st.airs[ord.proof_to_instance[j]];
            if (!air.preprocessed_width) continue;
            prep_pos[j] = pg.widths.size();
            pg.widths.push_back(air.preprocessed_width);
            pg.log_height = std::max(pg.log_height, lh[ord.proof_to_instance[j]] + lb);
        }
        groups.insert(groups.begin(), pg);
    }
    const size_t gm = has_prep ? 1 : 0;   // index of the main group
    for (auto& g : groups) for (size_t w : g.widths) g.aligned.push_back(aligned_len(w, lmcs_alignment()));

    // --- DEEP oracle (pcs/deep/verifier.rs): read evals, grind, alpha/beta, reduced openings
    size_t W = 0;
    for (auto& g : groups) for (size_t w : g.aligned) W += w;
    std::vector<Ef> evals[2];
    for (int p = 0; p < 2; p++) for (size_t i = 0; i < W; i++) evals[p].push_back(ch.receive_ext());
    ch.grind(params.deep_pow_bits);
    Ef dalpha = ch.sample_ext(), dbeta = ch.sample_ext();
    Ef reduced[2];
    for (int p = 0; p < 2; p++) { Ef a; for (const Ef& e : evals[p]) a = a * dalpha + e; reduced[p] = a; }
    // --- FRI oracle (pcs/fri/verifier.rs:60-…)
    unsigned la = params.log_folding_arity;
    size_t arity = size_t(1) << la;
    unsigned rounds = fri_num_rounds(params, log_lde);
    std::vector<Digest> fri_roots; std::vector<Ef> fri_betas;
    for (unsigned r = 0; r < rounds; r++) {
        fri_roots.push_back(ch.receive_commitment());
        ch.grind(params.folding_pow_bits);
        fri_betas.push_back(ch.sample_ext());
    }
    size_t final_deg = fri_final_poly_degree(params, log_lde);
    std::vector<Ef> final_poly;
    for (size_t i = 0; i < final_deg; i++) final_poly.push_back(ch.receive_ext());
    ch.grind(params.query_pow_bits);
    std::vector<size_t> qs;
    for (unsigned q = 0; q < params.num_queries; q++) qs.push_back((size_t)ch.sample_bits(log_lde));
    TreeIndices ti = TreeIndices::make(qs, log_lde);
    // --- DEEP open_batch
    std::map<size_t, Ef> reduced_rows;
    for (size_t i : ti.idx) reduced_rows[i] = Ef();
    for (auto& g : groups) {
        auto rows = lmcs_open_lifted_batch(g.root, g.aligned, ti, g.log_height, ch);
        for (auto& kv : reduced_rows) { Ef a = kv.second; for (Fp f : rows.at(kv.first)) a = a * dalpha + Ef(f); kv.second = a; }
    }
    std::map<size_t, Ef> fevals;
    Fp omega_l = two_adic_generator(log_lde);
    for (auto& kv : reduced_rows) {
        Fp x = shift * fp_pow(omega_l, kv.first);
        Ef d; Ef bp = ef_one();
        for (int p = 0; p < 2; p++) { d = d + bp * (reduced[p] - kv.second) * ef_inv(pts[p] - x); bp = bp * dbeta; }
        fevals[kv.first] = d;
    }
    // --- FRI low-degree test (fri/verifier.rs:105-212)
    unsigned log_dom = log_lde;
    for (unsigned r = 0; r < rounds; r++) {
        unsigned log_folded = log_dom - la;
        size_t folded = size_t(1) << log_folded;
        ti.shrink_depth(la);
        auto rows = lmcs_open_batch(fri_roots[r], {2 * arity}, ti, ch);
        Fp ginv = fp_inv(two_adic_generator(log_dom));
        std::map<size_t, Ef> next;
        for (auto& kv : fevals) {
            size_t row_idx = kv.first & (folded - 1);
            size_t pos = reverse_bits64(kv.first >> log_folded, la);
            const std::vector<Fp>& fr = rows.at(row_idx);
            std::vector<Ef> row(arity);
            for (size_t j = 0; j < arity; j++) row[j] = Ef(fr[2 * j], fr[2 * j + 1]);
            if (row[pos] != kv.second) throw std::runtime_error("fri: evaluation mismatch");
            next[row_idx] = fold_row(la, row.data(), fp_pow(ginv, row_idx), fri_betas[r]);
        }
        fevals = std::move(next);
        log_dom = log_folded;
    }
    Fp gen = two_adic_generator(log_dom);
    for (auto& kv : fevals) {
        Ef x(fp_pow(gen, kv.first));
        Ef acc;
        for (const Ef& c : final_poly) acc = acc * x + c;   // descending order
        if (acc != kv.second) throw std::runtime_error("fri: final polynomial mismatch");
    }
    // --- constraint check at z (verifier/mod.rs step 9-12)
    size_t off = 0;
    std::vector<size_t> goff(groups.size());
    for (size_t g = 0; g < groups.size(); g++) { goff[g] = off; for (size_t w : groups[g].aligned) off += w; }
    Ef accumulated;
    size_t moff = goff[gm], aoff = goff[gm + 1];
    std::vector<Ef> scratch;
    for (size_t j = 0; j < k; j++) {
        const AirDesc& air = st.airs[ord.proof_to_instance[j]];
        unsigned ln = lh[ord.proof_to_instance[j]];
        std::vector<Ef> ml(evals[0].begin() + moff, evals[0].begin() + moff + air.width);
        std::vector<Ef> mn(evals[1].begin() + moff, evals[1].begin() + moff + air.width);
        std::vector<Ef> al, an;
        for (size_t c = 0; c < air.aux_width; c++) {
            // EF cell from two opened base-column evaluations: v0 + u*v1 (row_to_packed_ext)
            Ef u(Fp(), Fp::raw(1));
            al.push_back(evals[0][aoff + 2 * c] + u * evals[0][aoff + 2 * c + 1]);
            an.push_back(evals[1][aoff + 2 * c] + u * evals[1][aoff + 2 * c + 1]);
        }
        moff += aligned_len(air.width, lmcs_alignment()); aoff += aligned_len(2 * air.aux_width, lmcs_alignment());
        // selectors_at (domain.rs:518-530) at the lifted point
        Ef zl = ef_exp_pow2(z, log_max_n - ln);
        Ef van = ef_exp_pow2(zl, ln) - Fp::raw(1);
        Fp ohi = fp_inv(two_adic_generator(ln));
        AirPoint pt{};
        pt.main_local_ef = ml.data(); pt.main_next_ef = mn.data(); pt.aux_local_ef = al.data(); pt.aux_next_ef = an.data();
        std::vector<Ef> ppl, ppn;
        if (has_prep && prep_pos[j] != (size_t)-1) {
            size_t po = goff[0];
            for (size_t q2 = 0; q2 < prep_pos[j]; q2++) po += groups[0].aligned[q2];
            ppl.assign(evals[0].begin() + po, evals[0].begin() + po + air.preprocessed_width);
            ppn.assign(evals[1].begin() + po, evals[1].begin() + po + air.preprocessed_width);
            pt.prep_local_ef = ppl.data(); pt.prep_next_ef = ppn.data();
        }
        std::vector<Ef> r(randomness.begin(), randomness.begin() + air.num_randomness);
        pt.publics = st.public_values.data(); pt.challenges = r.data(); pt.aux_values = auxv[j].data();
        pt.is_first = van * ef_inv(zl - Fp::raw(1)); pt.is_last = van * ef_inv(zl - ohi); pt.is_transition = zl - ohi;
        std::vector<Ef> per;
        if (air.n_periodic) {   // verifier/periodic.rs:55-71: z^(N_max / period)
            Ef y = ef_exp_pow2(z, log_max_n - air.log_max_period);
            for (auto& cf : air.periodic_coeffs()) { Ef a; for (size_t q = cf.size(); q-- > 0;) a = a * y + cf[q]; per.push_back(a); }
            pt.periodic = per.data();
        }
        Ef folded = air_eval_folded(air.program, pt, alpha, scratch);
        accumulated = accumulated * beta + folded;
    }
    // reconstruct_quotient (domain.rs:773-795)
    {
        Ef u(Fp(), Fp::raw(1));
        std::vector<Ef> chunks;
        for (size_t t = 0; t < D; t++) chunks.push_back(evals[0][goff[gm + 2] + 2 * t] + u * evals[0][goff[gm + 2] + 2 * t + 1]);
        Fp omega_s = two_adic_generator(log_qd);
        Ef uu = ef_exp_pow2(z * sinv, log_max_n);
        Ef num, den; Fp ost = Fp::raw(1);
        for (size_t t = 0; t < D; t++) {
            Ef wt = ef_inv(uu - ost) * ost;
            num = num + wt * chunks[t]; den = den + wt; ost = ost * omega_s;
        }
        Ef qz = num * ef_inv(den);
        Ef van = ef_exp_pow2(z, log_max_n) - Fp::raw(1);
        if (accumulated != qz * van) throw std::runtime_error("verify: constraint mismatch");
    }
    if (!ch.is_empty()) throw std::runtime_error("verify: trailing transcript data");
}

}  // namespace orc
