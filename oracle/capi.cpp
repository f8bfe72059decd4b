// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// C entry points of the CPU oracle, loaded with ctypes by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.  The struct layouts deliberately equal the
// ones in include/miden_b200.h so a test can hand the same buffers to both sides; the two
// libraries share no code.
#include "stark.hpp"
#include <cstdio>
#include <cstring>
#include <string>
#include <malloc.h>

using namespace orc;

// Keep large buffers on the heap instead of returning them to the kernel after every phase
// (fresh mmap + page faults dominated the first version's run time).
__attribute__((constructor)) static void orc_tune_malloc() {
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, -1);
    mallopt(M_TOP_PAD, 256 << 20);
}

// StatefulSponge over the reference's MockBinaryPermutation (see orc_mock_sponge below)
template <size_t WIDTH, size_t RATE>
static void mock_sponge(const uint64_t* in, size_t n, uint64_t* state_out) {
    std::array<uint64_t, WIDTH> st{};
    sponge_absorb_generic<uint64_t, WIDTH, RATE>(st, in, n, [](std::array<uint64_t, WIDTH>& s) {
        uint64_t w = 0;
        for (size_t i = 0; i < WIDTH; i++) w += s[i] * (uint64_t)(i + 1);
        s.fill(w);
    });
    for (size_t i = 0; i < WIDTH; i++) state_out[i] = st[i];
}

extern "C" {

struct orc_pcs_params { uint32_t log_blowup, log_folding_arity, log_final_degree, folding_pow_bits, deep_pow_bits, num_queries, query_pow_bits; };
struct orc_challenger { uint64_t sponge_state[12]; uint64_t input_buffer[8]; uint32_t input_len, output_len; };
struct orc_lookup { uint32_t num_columns, program_words; const uint32_t* program; };
struct orc_air { uint32_t width, aux_width, num_aux_values, num_randomness, log_quotient_degree, program_words; const uint32_t* program;
                 const uint64_t* periodic_values; uint32_t num_periodic_columns, log_max_period, preprocessed_width;
                 const struct orc_lookup* lookup; };
struct orc_matrix { const uint64_t* values; uint32_t log_height, width; };
struct orc_statement { const orc_air* airs; uint32_t n_airs; const uint64_t* public_values; uint32_t n_public_values; const uint64_t* observe_felts; uint32_t n_observe_felts; };
typedef int (*orc_aux_builder)(void* ctx, uint32_t instance, const orc_matrix* main, const uint64_t* randomness, uint64_t* aux_out, uint64_t* aux_values);
struct orc_proof { const uint8_t* log_trace_heights; size_t n_heights; const uint64_t* fields; size_t n_fields; const uint64_t* commitments; size_t n_commitments; };

static thread_local std::string g_err;
const char* orc_last_error() { return g_err.c_str(); }

// OpenMP team size of every later call.  `torch.distributed.run` exports OMP_NUM_THREADS=1 to its workers, which
// would run the CPU baseline single-threaded; bench.py sets the team size explicitly (n <= 0: all online CPUs).
int orc_set_threads(int n) {
    if (n <= 0) n = omp_get_num_procs();
    omp_set_dynamic(0);
    omp_set_num_threads(n);
    return n;
}
int orc_get_threads() { return omp_get_max_threads(); }

void orc_poseidon2_permute(uint64_t* st, size_t n) {
    for (size_t i = 0; i < n; i++) {
        State s; for (int k = 0; k < 12; k++) s[k] = Fp(st[12 * i + k]);
        poseidon2_permute(s);
        for (int k = 0; k < 12; k++) st[12 * i + k] = s[k].v;
    }
}

// n states through the AVX-512 path (8 at a time; the tail and machines without AVX-512 use the scalar code).
// Returns 1 when the vector path ran.
int orc_poseidon2_permute_x8(uint64_t* st, size_t n) {
    std::vector<State> s(n);
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 12; k++) s[i][k] = Fp(st[12 * i + k]);
    size_t i = 0; int used = 0;
#if ORC_HAVE_X8
    if (x8_available()) { for (; i + 8 <= n; i += 8) poseidon2_permute_x8(&s[i]); used = n >= 8; }
#endif
    for (; i < n; i++) poseidon2_permute(s[i]);
    for (size_t j = 0; j < n; j++) for (int k = 0; k < 12; k++) st[12 * j + k] = s[j][k].v;
    return used;
}

uint64_t orc_fp_mul(uint64_t a, uint64_t b) { return (Fp(a) * Fp(b)).v; }
uint64_t orc_fp_inv(uint64_t a) { return fp_inv(Fp(a)).v; }
uint64_t orc_two_adic_generator(uint32_t bits) { return two_adic_generator(bits).v; }
uint64_t orc_lde_shift(uint32_t log_lde) { return lde_shift(log_lde).v; }

// One FRI fold of a physical (bit-reversed) row of 2^log_arity extension values, as the commit phase and
// the verifier apply it (pcs/fri/fold/arity{2,4,8}.rs).  row = 2 * 2^log_arity felts, beta/out = 2 felts.
// Exposed for tests/test_anchors.py, which compares it with the VM's own restatement of the arity-4 fold
// (processor/src/execution/operations/fri_ops/mod.rs:48-240).
int orc_fri_fold_row(uint32_t log_arity, const uint64_t* row, uint64_t s_inv, const uint64_t* beta, uint64_t* out) {
    try {
        Ef ev[8];
        for (unsigned j = 0; j < (1u << log_arity); j++) ev[j] = Ef(Fp(row[2 * j]), Fp(row[2 * j + 1]));
        Ef r = fold_row(log_arity, ev, Fp(s_inv), Ef(Fp(beta[0]), Fp(beta[1])));
        out[0] = r.a.v; out[1] = r.b.v;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ---- hooks for the reference's own unit vectors (tests/test_reference_vectors.py, tests/golden/reference_unit_vectors.json)
// StatefulSponge over the reference's MockBinaryPermutation (crates/stateful-hasher/src/testing.rs:38-49: every element
// becomes sum(state[i] * (i + 1)), wrapping u64), for the (WIDTH, RATE) pairs its tests use (field_sponge.rs:85-160).
int orc_mock_sponge(uint32_t width, uint32_t rate, const uint64_t* in, size_t n, uint64_t* state_out) {
    if (width == 4 && rate == 2) mock_sponge<4, 2>(in, n, state_out);
    else if (width == 6 && rate == 3) mock_sponge<6, 3>(in, n, state_out);
    else if (width == 8 && rate == 4) mock_sponge<8, 4>(in, n, state_out);
    else if (width == 12 && rate == 8) mock_sponge<12, 8>(in, n, state_out);
    else return -1;
    return 0;
}
// TreeIndices (lmcs/tree_indices.rs): op 0 = new (sorted, de-duplicated; -1 if an index is out of range),
// 1 = fold_to_depth(arg) (-1 if arg > depth), 2 = shrink_depth(arg), 3 = missing_siblings as (depth, position) pairs.
// Returns the number of u64 written to out (op 3: 2 per node) and the resulting depth in *depth_out.
long long orc_tree_indices(int op, const uint64_t* idx, size_t n, uint32_t depth, uint32_t arg, uint64_t* out, size_t cap, uint32_t* depth_out) {
    std::vector<size_t> v(idx, idx + n);
    for (size_t x : v) if (depth < 64 && x >= (size_t(1) << depth)) return -1;     // TreeIndices::new validation (:34-45)
    TreeIndices t = TreeIndices::make(v, depth);
    std::vector<uint64_t> res;
    if (op == 1) { if (arg > depth) return -1; t = t.fold_to_depth(arg); }
    else if (op == 2) t.shrink_depth(arg);
    if (op == 3) { for (auto& ds : missing_siblings(t)) { res.push_back(ds.first); res.push_back(ds.second); } }
    else for (size_t x : t.idx) res.push_back(x);
    if (depth_out) *depth_out = t.depth;
    for (size_t i = 0; i < res.size() && i < cap; i++) out[i] = res[i];
    return (long long)res.size();
}
// PcsParams::new validation (pcs/params.rs:53-99): 0 ok, 1 InvalidFoldingArity, 2 ZeroBlowup, 3 ZeroQueries,
// 4 FinalDegreeUnreachable; FriParams::num_rounds / final_poly_degree (pcs/fri/mod.rs:80-115) for an LDE of 2^log_lde.
int orc_pcs_params_check(const orc_pcs_params* p) {
    if (p->log_folding_arity < 1 || p->log_folding_arity > 3) return 1;
    if (p->log_blowup == 0) return 2;
    if (p->num_queries == 0) return 3;
    if (p->log_final_degree + p->log_blowup < p->log_folding_arity - 1) return 4;
    return 0;
}
void orc_fri_shape(const orc_pcs_params* p, uint32_t log_lde, uint32_t* rounds, uint64_t* final_degree) {
    PcsParams q; q.log_blowup = p->log_blowup; q.log_folding_arity = p->log_folding_arity; q.log_final_degree = p->log_final_degree;
    *rounds = fri_num_rounds(q, log_lde);
    *final_degree = fri_final_poly_degree(q, log_lde);
}

static Matrix to_matrix(const orc_matrix& m) {
    Matrix r(size_t(1) << m.log_height, m.width);
    for (size_t i = 0; i < r.v.size(); i++) r.v[i] = Fp(m.values[i]);
    return r;
}

// Naive O(n^2) DFT of one column (natural -> natural), the anchor for the fast transforms.
void orc_naive_dft(const uint64_t* in, uint32_t log_n, uint64_t* out) {
    size_t n = size_t(1) << log_n;
    Fp w = two_adic_generator(log_n);
    for (size_t k = 0; k < n; k++) {
        Fp wk = fp_pow(w, k), x = Fp::raw(1), acc;
        for (size_t j = 0; j < n; j++) { acc = acc + Fp(in[j]) * x; x = x * wk; }
        out[k] = acc.v;
    }
}
void orc_dft(const uint64_t* in, uint32_t log_n, uint32_t width, int inverse, uint64_t* out) {
    orc_matrix om{in, log_n, width};
    Matrix m = to_matrix(om);
    if (inverse) idft_rows(m); else dft_rows(m);
    for (size_t i = 0; i < m.v.size(); i++) out[i] = m.v[i].v;
}

// coset_lde_batch: bit-reversed rows, row-major.
void orc_coset_lde_batch(const orc_matrix* mat, uint32_t added_bits, uint64_t shift, uint64_t* out) {
    Matrix r = coset_lde_bitrev(to_matrix(*mat), added_bits, Fp(shift));
    for (size_t i = 0; i < r.v.size(); i++) out[i] = r.v[i].v;
}

// build_aligned_tree over matrices given in DOMAIN order (ascending heights); returns the root and,
// if layers_out != NULL, all digest layers top-down (2^(depth+1) - 1 digests).
void orc_lmcs_commit(const orc_matrix* mats, uint32_t n, uint64_t root[4], uint64_t* layers_out) {
    std::vector<Matrix> ms;
    for (uint32_t i = 0; i < n; i++) { Matrix m = to_matrix(mats[i]); bit_reverse_rows(m); ms.push_back(std::move(m)); }
    LmcsTree t = LmcsTree::build(std::move(ms), lmcs_alignment());
    for (int i = 0; i < 4; i++) root[i] = t.root()[i].v;
    if (layers_out) {
        size_t o = 0;
        for (auto& layer : t.layers) for (auto& d : layer) for (int i = 0; i < 4; i++) layers_out[o++] = d[i].v;
    }
}

// STARK hash configuration of every later call (0 = Poseidon2, the default; 1 = Blake3_256; 2 = Keccak; 3 = RPO; 4 = RPX) and, for the two byte-oriented ones, the
// pre-bound challenger = the bytes its HashChallenger input buffer holds after `config.challenger()` +
// `observe_protocol_params` (air/src/config.rs:299-307,188-198); the `orc_challenger*` argument is then ignored.
static std::vector<uint8_t> g_hash_challenger_input;
int orc_set_hash(int kind, const uint8_t* challenger_input, size_t n) {
    if (kind < 0 || kind > 4) return -1;
    hash_kind() = (HashKind)kind;
    g_hash_challenger_input.assign(challenger_input, challenger_input + (challenger_input ? n : 0));
    return 0;
}
// RPO / RPX: the permutation on 12 canonical felts (kind 3 / 4) and AlgebraicSponge::hash_elements (the reference's known answers)
void orc_rescue_permute(int kind, uint64_t st[12]) {
    RState s; for (int i = 0; i < 12; i++) s[i] = Fp(st[i]);
    if (kind == 4) rpx_permute(s); else rpo_permute(s);
    for (int i = 0; i < 12; i++) st[i] = s[i].v;
}
void orc_rescue_hash_elements(int kind, const uint64_t* e, size_t n, uint64_t out[4]) {
    std::vector<Fp> v(n); for (size_t i = 0; i < n; i++) v[i] = Fp(e[i]);
    auto d = kind == 4 ? rescue_hash_elements(v.data(), n, [](RState& s) { rpx_permute(s); }) : rescue_hash_elements(v.data(), n, [](RState& s) { rpo_permute(s); });
    for (int i = 0; i < 4; i++) out[i] = d[i].v;
}
// Keccak-256 (pad = 1) / SHA3-256 (pad = 6) of a byte string, and the permutation on 25 lanes
void orc_keccak256(const uint8_t* p, size_t n, uint8_t pad, uint8_t out[32]) { auto h = keccak::hash256(p, n, pad); memcpy(out, h.data(), 32); }
void orc_keccak_f(uint64_t st[25]) { std::array<u64, 25> a; memcpy(a.data(), st, 200); keccak::permute(a); memcpy(st, a.data(), 200); }
void orc_blake3(const uint8_t* p, size_t n, uint8_t out[32]) { auto h = blake3::hash(p, n); memcpy(out, h.data(), 32); }

static Challenger to_challenger(const orc_challenger* c) {
    if (byte_hash()) return Challenger::from_bytes(g_hash_challenger_input.data(), g_hash_challenger_input.size());
    Challenger ch;
    for (int i = 0; i < 12; i++) ch.st[i] = Fp(c->sponge_state[i]);
    for (uint32_t i = 0; i < c->input_len; i++) ch.in_buf[i] = Fp(c->input_buffer[i]);
    ch.in_len = c->input_len; ch.out_len = c->output_len;
    return ch;
}

// Challenger scripting for transcript parity tests: ops[i] = 0 observe(arg) / 1 sample / 2 sample_bits(arg)
// / 3 grind(arg).  out[i] receives the sampled value / witness (0 for observe).
void orc_challenger_script(orc_challenger* c, const uint32_t* ops, const uint64_t* args, size_t n, uint64_t* out) {
    Challenger ch;
    for (int i = 0; i < 12; i++) ch.st[i] = Fp(c->sponge_state[i]);
    for (uint32_t i = 0; i < c->input_len; i++) ch.in_buf[i] = Fp(c->input_buffer[i]);
    ch.in_len = c->input_len; ch.out_len = c->output_len;
    if (byte_hash()) ch = to_challenger(c);      // the hash challenger starts from the bytes given to orc_set_hash
    for (size_t i = 0; i < n; i++) {
        switch (ops[i]) {
            case 0: ch.observe(Fp(args[i])); out[i] = 0; break;
            case 1: out[i] = ch.sample().v; break;
            case 2: out[i] = ch.sample_bits((unsigned)args[i]); break;
            case 3: out[i] = ch.grind((unsigned)args[i]).v; break;
        }
    }
    for (int i = 0; i < 12; i++) c->sponge_state[i] = ch.st[i].v;
    for (int i = 0; i < 8; i++) c->input_buffer[i] = i < ch.in_len ? ch.in_buf[i].v : 0;
    c->input_len = ch.in_len; c->output_len = ch.out_len;
}

struct orc_prove_result {
    Proof proof;
    ProveDebug dbg;
    std::vector<uint64_t> fields, commitments;
};

static Statement to_statement(const orc_statement* st) {
    Statement s;
    for (uint32_t i = 0; i < st->n_airs; i++) {
        const orc_air& a = st->airs[i];
        AirDesc d;
        d.width = a.width; d.aux_width = a.aux_width; d.num_aux_values = a.num_aux_values;
        d.num_randomness = a.num_randomness; d.log_quotient_degree = a.log_quotient_degree;
        d.program = AirProgram::parse(a.program, a.program_words);
        d.n_periodic = a.num_periodic_columns; d.log_max_period = a.log_max_period; d.preprocessed_width = a.preprocessed_width;
        for (size_t q = 0; q < ((size_t)a.num_periodic_columns << a.log_max_period); q++) d.periodic.push_back(Fp(a.periodic_values[q]));
        for (auto& nd : d.program.nodes) if (nd.op == OP_PERIODIC && nd.a >= d.n_periodic) throw std::runtime_error("air program: periodic column out of range");
        if (a.lookup) {
            d.lookup = LookupProgram::parse(a.lookup->num_columns, a.lookup->program, a.lookup->program_words);
            for (auto& nd : d.lookup.nodes) {
                if (nd.op == OP_PERIODIC && nd.a >= d.n_periodic) throw std::runtime_error("lookup program: periodic column out of range");
                if (nd.op == OP_MAIN && (nd.a > 1 || nd.b >= d.width)) throw std::runtime_error("lookup program: main column out of range");
                if (nd.op == OP_CHALLENGE && nd.a >= d.num_randomness) throw std::runtime_error("lookup program: challenge out of range");
                if (nd.op == OP_PUBLIC && nd.a >= st->n_public_values) throw std::runtime_error("lookup program: public value out of range");
            }
        }
        s.airs.push_back(std::move(d));
    }
    for (uint32_t i = 0; i < st->n_public_values; i++) s.public_values.push_back(Fp(st->public_values[i]));
    for (uint32_t i = 0; i < st->n_observe_felts; i++) s.observe_felts.push_back(Fp(st->observe_felts[i]));
    return s;
}
static PcsParams to_params(const orc_pcs_params* p) {
    PcsParams q;
    q.log_blowup = p->log_blowup; q.log_folding_arity = p->log_folding_arity; q.log_final_degree = p->log_final_degree;
    q.folding_pow_bits = p->folding_pow_bits; q.deep_pow_bits = p->deep_pow_bits; q.num_queries = p->num_queries;
    q.query_pow_bits = p->query_pow_bits;
    return q;
}

// Full prove.  Returns an opaque handle (NULL on error); `out` points into it.  `preprocessed` (may be NULL):
// one matrix per AIR in instance order, width 0 = none.
void* orc_prove_pp(const orc_pcs_params* params, const orc_statement* st, const orc_matrix* traces, const orc_matrix* preprocessed,
                   const orc_challenger* challenger, orc_aux_builder cb, void* ctx, orc_proof* out, uint64_t* prep_commitment_out) {
    try {
        Statement s = to_statement(st);
        if (preprocessed) for (uint32_t i = 0; i < st->n_airs; i++) s.preprocessed.push_back(preprocessed[i].width ? to_matrix(preprocessed[i]) : Matrix());
        if (preprocessed && prep_commitment_out && s.has_preprocessed()) { PreprocessedBundle b = build_preprocessed(s, params->log_blowup); for (int i = 0; i < 4; i++) prep_commitment_out[i] = b.tree.root()[i].v; }
        std::vector<Matrix> tr;
        for (uint32_t i = 0; i < st->n_airs; i++) tr.push_back(to_matrix(traces[i]));
        AuxBuilder ab;
        if (cb) ab = [&](size_t inst, const Matrix& main, const std::vector<Ef>& r, Matrix& aux, std::vector<Ef>& av) {
            std::vector<uint64_t> rr; for (auto& e : r) { rr.push_back(e.a.v); rr.push_back(e.b.v); }
            std::vector<uint64_t> ao(aux.v.size()), avo(2 * av.size());
            std::vector<uint64_t> mv(main.v.size()); for (size_t i = 0; i < mv.size(); i++) mv[i] = main.v[i].v;
            orc_matrix mm{mv.data(), log2_strict(main.height), (uint32_t)main.width};
            if (cb(ctx, (uint32_t)inst, &mm, rr.data(), ao.data(), avo.data()) != 0) throw std::runtime_error("aux builder failed");
            for (size_t i = 0; i < ao.size(); i++) aux.v[i] = Fp(ao[i]);
            for (size_t i = 0; i < av.size(); i++) av[i] = Ef(Fp(avo[2 * i]), Fp(avo[2 * i + 1]));
        };
        auto* res = new orc_prove_result();
        res->proof = stark_prove(to_params(params), s, tr, to_challenger(challenger), ab, &res->dbg);
        for (auto& f : res->proof.fields) res->fields.push_back(f.v);
        for (auto& d : res->proof.commitments) for (int i = 0; i < 4; i++) res->commitments.push_back(d[i].v);
        out->log_trace_heights = res->proof.log_trace_heights.data(); out->n_heights = res->proof.log_trace_heights.size();
        out->fields = res->fields.data(); out->n_fields = res->fields.size();
        out->commitments = res->commitments.data(); out->n_commitments = res->proof.commitments.size();
        return res;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void* orc_prove(const orc_pcs_params* params, const orc_statement* st, const orc_matrix* traces,
                const orc_challenger* challenger, orc_aux_builder cb, void* ctx, orc_proof* out) {
    return orc_prove_pp(params, st, traces, nullptr, challenger, cb, ctx, out, nullptr);
}
void orc_prove_free(void* h) { delete (orc_prove_result*)h; }

// what: same numbering as mdn_info in include/miden_b200.h.
long long orc_prove_info(void* h, int what, uint64_t* out, size_t cap) {
    auto* r = (orc_prove_result*)h;
    std::vector<uint64_t> v;
    auto push_d = [&](const Digest& d) { for (int i = 0; i < 4; i++) v.push_back(d[i].v); };
    auto push_e = [&](const Ef& e) { v.push_back(e.a.v); v.push_back(e.b.v); };
    switch (what) {
        case 0: push_d(r->dbg.main_root); break;
        case 1: push_d(r->dbg.aux_root); break;
        case 2: push_d(r->dbg.quotient_root); break;
        case 3: push_e(r->dbg.z); break;
        case 4: for (auto& e : r->dbg.quotient_acc) push_e(e); break;
        case 5: for (auto& e : r->dbg.open.deep_evals) push_e(e); break;
        case 6: for (auto& d : r->dbg.open.fri_roots) push_d(d); break;
        case 7: for (auto q : r->dbg.open.query_indices) v.push_back(q); break;
        default: return -1;
    }
    for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
    return (long long)v.size();
}

// Verify a proof given as raw streams.  0 = accepted; -1 = rejected (orc_last_error has the reason).
int orc_verify_pp(const orc_pcs_params* params, const orc_statement* st, const orc_proof* pf, const orc_challenger* challenger,
                  const uint64_t* prep_commitment) {
    try {
        Statement s = to_statement(st);
        Proof p;
        p.log_trace_heights.assign(pf->log_trace_heights, pf->log_trace_heights + pf->n_heights);
        for (size_t i = 0; i < pf->n_fields; i++) {
            if (pf->fields[i] >= P) throw std::runtime_error("non-canonical field element");
            p.fields.push_back(Fp::raw(pf->fields[i]));
        }
        for (size_t i = 0; i < pf->n_commitments; i++) {
            Digest d; for (int k = 0; k < 4; k++) d[k] = Fp(pf->commitments[4 * i + k]);
            p.commitments.push_back(d);
        }
        Digest pc; if (prep_commitment) for (int k = 0; k < 4; k++) pc[k] = Fp(prep_commitment[k]);
        stark_verify(to_params(params), s, p, to_challenger(challenger), prep_commitment ? &pc : nullptr);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int orc_verify(const orc_pcs_params* params, const orc_statement* st, const orc_proof* pf, const orc_challenger* challenger) {
    return orc_verify_pp(params, st, pf, challenger, nullptr);
}

}  // extern "C"
