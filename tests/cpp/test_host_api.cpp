// Parity test of the C++ host layer (miden-vm_b200/host/miden_prover.hpp), written the way the reference tests
// its prover: prove -> verify -> tamper (crates/lifted-stark/src/testing/configs/goldilocks_poseidon2.rs:143-166),
// plus the constructor-time validation of ProverStatement / ProverInstance (prover/mod.rs:139-153).
// The verifier is the CPU oracle (TEST INFRASTRUCTURE; oracle/capi.cpp mirrors the C-ABI structs one to one).
// Exit code 0 = all checks passed; prints NO_DEVICE and exits 3 when no CUDA device is usable (expected on CPU).
#include "../../miden-vm_b200/host/miden_prover.hpp"
#include <cstdio>
#include <cstring>

extern "C" {
int orc_verify(const void* params, const void* st, const void* proof, const void* challenger);
int orc_verify_pp(const void* params, const void* st, const void* proof, const void* challenger, const uint64_t* prep_commitment);
const char* orc_last_error();
int orc_set_hash(int kind, const uint8_t* challenger_input, size_t n);
}

using namespace miden;
static const Felt P = 0xFFFFFFFF00000001ULL;

static Felt splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
// DummyMidenAir trace (testing/airs/miden.rs:101-123): random cells, column 0 zero
static RowMajorMatrix synthetic_trace(uint64_t air, uint32_t log_h, uint32_t w) {
    std::vector<Felt> v((size_t(1) << log_h) * w);
    for (size_t i = 0; i < v.size(); i++) { Felt x = splitmix64(i ^ (2025ULL ^ (air << 56))); v[i] = x >= P ? x - P : x; }
    for (size_t r = 0; r < (size_t(1) << log_h); r++) v[r * w] = 0;
    return RowMajorMatrix(std::move(v), w);
}
static Felt mulmod(Felt a, Felt b) { return (Felt)((unsigned __int128)a * b % P); }

// local[0] * ... * local[8] == 0 (testing/airs/miden.rs:57-62), folded from ONE like the reference
static Air dummy_miden_air(uint32_t width, uint32_t aux_width) {
    AirBuilder b;
    auto acc = b.constant(1);
    for (uint32_t j = 0; j < 9; j++) acc = b.mul(acc, b.main(0, j));
    b.assert_zero(acc);
    Air a; a.width = width; a.aux_width = aux_width; a.num_aux_values = aux_width; a.num_randomness = 2; a.log_quotient_degree = 3;
    a.program = b.finish();
    return a;
}
// main[0] = prep[0] * main[1], main[2] = prep_next[1] + main[1]
static Air preprocessed_air(uint32_t log_h, RowMajorMatrix& trace) {
    size_t n = size_t(1) << log_h;
    RowMajorMatrix pm = synthetic_trace(40, log_h, 2);
    for (size_t r = 0; r < n; r++) pm.values[2 * r] = splitmix64(r * 77 + 5) % P;
    trace = synthetic_trace(20, log_h, 3);
    for (size_t r = 0; r < n; r++) {
        Felt c1 = trace.values[3 * r + 1];
        trace.values[3 * r] = mulmod(pm.values[2 * r], c1);
        Felt s = pm.values[2 * ((r + 1) % n) + 1] + c1; if (s >= P || s < c1) s -= P;
        trace.values[3 * r + 2] = s;
    }
    AirBuilder b;
    b.assert_zero(b.sub(b.main(0, 0), b.mul(b.preprocessed(0, 0), b.main(0, 1))));
    b.assert_zero(b.sub(b.sub(b.main(0, 2), b.preprocessed(1, 1)), b.main(0, 1)));
    Air a; a.width = 3; a.log_quotient_degree = 1; a.preprocessed_width = 2; a.program = b.finish(); a.preprocessed_trace = pm;
    return a;
}

// config.challenger() + observe_protocol_params (air/src/config.rs:188-198, 264-271)
static Challenger initial_challenger(const PcsParams& p) {
    static const Felt RELATION_DIGEST[4] = {837197885082815666ULL, 17812429367884914ULL, 12945170128166309606ULL, 6547471563106428306ULL};
    Challenger c;
    for (int i = 0; i < 4; i++) c.raw.sponge_state[8 + i] = RELATION_DIGEST[i];
    c.observe_slice({p.num_queries, p.query_pow_bits, p.deep_pow_bits, p.folding_pow_bits, p.log_blowup, p.log_final_degree, Felt(1) << p.log_folding_arity, 0});
    return c;
}

struct OracleView {   // the C structs the oracle verifier reads (same layout as the product's)
    detail::Lowered low; mdn_pcs_params params; mdn_proof proof; std::vector<Felt> comm;
    OracleView(const Statement& s, const PcsParams& p, const StarkProofData& pf) : low(s), params(p.raw()) {
        for (auto& c : pf.transcript.commitments) comm.insert(comm.end(), c.begin(), c.end());
        proof = mdn_proof{pf.log_trace_heights.data(), pf.log_trace_heights.size(), pf.transcript.fields.data(), pf.transcript.fields.size(), comm.data(), pf.transcript.commitments.size()};
    }
};
static int verify(const Statement& s, const PcsParams& p, const StarkProofData& pf, const Challenger& ch, const Commitment* prep = nullptr) {
    OracleView v(s, p, pf);
    return prep ? orc_verify_pp(&v.params, &v.low.st, &v.proof, &ch.raw, prep->data()) : orc_verify(&v.params, &v.low.st, &v.proof, &ch.raw);
}
#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "CHECK failed at line %d: %s (%s)\n", __LINE__, #cond, orc_last_error()); return 1; } } while (0)

int main() {
    PcsParams params; params.log_final_degree = 2; params.folding_pow_bits = 2; params.deep_pow_bits = 3; params.num_queries = 5; params.query_pow_bits = 4;
    Challenger proto = initial_challenger(params);
    std::unique_ptr<StarkConfig> config;
    try { config.reset(new StarkConfig(params, proto, 0)); }
    catch (const ProverError& e) {
        if (e.kind == ProverError::NoDevice) { printf("NO_DEVICE %s\n", e.what()); return 3; }
        throw;
    }
    // 1. prove -> verify -> tamper, mixed heights
    {
        Statement st = Statement::with_default_observe({dummy_miden_air(11, 2), dummy_miden_air(9, 1), dummy_miden_air(10, 1)}, {});
        ProverStatement ps(st, {synthetic_trace(0, 6, 11), synthetic_trace(1, 7, 9), synthetic_trace(2, 5, 10)});
        StarkOutput out = ProverInstance(*config, ps, nullptr).prove(config->challenger());
        CHECK(out.proof.log_trace_heights == (std::vector<uint8_t>{6, 7, 5}));
        CHECK(verify(st, params, out.proof, proto) == 0);
        StarkProofData bad = out.proof; bad.transcript.fields[bad.transcript.fields.size() / 2] ^= 1;
        CHECK(verify(st, params, bad, proto) != 0);
        bad = out.proof; bad.transcript.commitments.back()[0] ^= 1;
        CHECK(verify(st, params, bad, proto) != 0);
        StarkOutput again = ProverInstance(*config, ps, nullptr).prove(config->challenger());      // deterministic
        CHECK(again.proof.transcript.fields == out.proof.transcript.fields);
        // a violated constraint yields a proof the verifier rejects
        ProverStatement ps_bad(st, {synthetic_trace(0, 6, 11), synthetic_trace(1, 7, 9), synthetic_trace(2, 5, 10)});
        ps_bad.traces[1].values[3 * 9] = 1;
        StarkOutput wrong = ProverInstance(*config, ps_bad, nullptr).prove(config->challenger());
        CHECK(verify(st, params, wrong.proof, proto) != 0);
    }
    // 2. InstanceError at construction: trace count / width
    {
        Statement st = Statement::with_default_observe({dummy_miden_air(9, 1)}, {});
        bool threw = false;
        try { ProverStatement ps(st, {synthetic_trace(0, 5, 10)}); } catch (const ProverError& e) { threw = e.kind == ProverError::Instance; }
        CHECK(threw);
        threw = false;
        try { ProverStatement ps(st, {}); } catch (const ProverError& e) { threw = e.kind == ProverError::Instance; }
        CHECK(threw);
        // DomainError from the backend: quotient degree above the blowup
        Statement st2 = st; st2.airs[0].log_quotient_degree = 4;
        ProverStatement ps2(st2, {synthetic_trace(0, 5, 9)});
        threw = false;
        try { ProverInstance(*config, ps2, nullptr).prove(config->challenger()); } catch (const ProverError& e) { threw = e.kind == ProverError::Domain; }
        CHECK(threw);
    }
    // 3. Preprocessed::build / presence parity / proof against the commitment
    {
        RowMajorMatrix t;
        Air a = preprocessed_air(6, t);
        Statement st = Statement::with_default_observe({a}, {});
        ProverStatement ps(st, {t});
        bool threw = false;
        try { ProverInstance inst(*config, ps, nullptr); } catch (const ProverError& e) { threw = e.kind == ProverError::Instance; }
        CHECK(threw);                                            // PresenceMismatch
        std::unique_ptr<Preprocessed> pp = Preprocessed::build(st, *config);
        CHECK(pp != nullptr);
        StarkOutput out = ProverInstance(*config, ps, pp.get()).prove(config->challenger());
        CHECK(verify(st, params, out.proof, proto, &pp->commitment()) == 0);
        Commitment wrong = pp->commitment(); wrong[2] ^= 1;
        CHECK(verify(st, params, out.proof, proto, &wrong) != 0);
        Statement plain = Statement::with_default_observe({dummy_miden_air(9, 1)}, {});
        CHECK(Preprocessed::build(plain, *config) == nullptr);
    }
    // 4. the other HashFunctions of prove_stark: the same statement under Blake3_256, Keccak, Rpo256, Rpx256; every proof is
    //    accepted by the oracle verifier in that mode only, and the config returns to Poseidon2
    {
        Statement st = Statement::with_default_observe({dummy_miden_air(18, 2), dummy_miden_air(9, 1)}, {});
        ProverStatement ps(st, {synthetic_trace(0, 6, 18), synthetic_trace(1, 5, 9)});
        StarkOutput p2 = ProverInstance(*config, ps, nullptr).prove(config->challenger());
        std::vector<uint8_t> init;                       // relation digest + parameter felts as little-endian bytes
        {
            static const Felt RELATION_DIGEST[4] = {837197885082815666ULL, 17812429367884914ULL, 12945170128166309606ULL, 6547471563106428306ULL};
            std::vector<Felt> f(RELATION_DIGEST, RELATION_DIGEST + 4);
            for (Felt v : {Felt(params.num_queries), Felt(params.query_pow_bits), Felt(params.deep_pow_bits), Felt(params.folding_pow_bits),
                           Felt(params.log_blowup), Felt(params.log_final_degree), Felt(1) << params.log_folding_arity, Felt(0)}) f.push_back(v);
            for (Felt v : f) for (int k = 0; k < 8; k++) init.push_back((uint8_t)(v >> (8 * k)));
        }
        const HashFunction kinds[4] = {HashFunction::Blake3_256, HashFunction::Keccak, HashFunction::Rpo256, HashFunction::Rpx256};
        for (HashFunction h : kinds) {
            config->with_hash(h, init);
            StarkOutput out = ProverInstance(*config, ps, nullptr).prove(config->challenger());
            CHECK(out.proof.transcript.commitments[0] != p2.proof.transcript.commitments[0]);
            CHECK(orc_set_hash((int)h, init.data(), init.size()) == 0);
            CHECK(verify(st, params, out.proof, proto) == 0);
            StarkProofData bad = out.proof; bad.transcript.fields[bad.transcript.fields.size() / 3] ^= 1;
            CHECK(verify(st, params, bad, proto) != 0);
            CHECK(orc_set_hash(0, nullptr, 0) == 0);
            CHECK(verify(st, params, out.proof, proto) != 0);          // the Poseidon2 verifier rejects it
        }
        config->with_hash(HashFunction::Poseidon2);
        StarkOutput back = ProverInstance(*config, ps, nullptr).prove(config->challenger());
        CHECK(back.proof.transcript.fields == p2.proof.transcript.fields && back.proof.transcript.commitments == p2.proof.transcript.commitments);
    }
    printf("HOST_API_OK\n");
    return 0;
}
