// Host-side mirror of the reference's proving interface, in C++ over the C ABI (include/miden_b200.h).
//
// The reference host is Rust; this image has no Rust toolchain, so the layer a Rust maintainer would write above
// the FFI is restated here in C++ with the reference's names and argument meaning:
//
//   reference (crates/lifted-stark, prover/src/lib.rs)            here
//   ------------------------------------------------------------  -----------------------------------------------
//   PcsParams::new(..)                      pcs/params.rs:35-99   miden::PcsParams
//   StarkConfig (pcs + lmcs + dft + challenger)  config.rs:26-45  miden::StarkConfig   (device instead of dft/lmcs)
//   LiftedAir: width / aux_width / num_randomness / eval ..       miden::Air           (eval lowered to an op-list)
//   Statement::new(airs, air_inputs, aux_inputs)                  miden::Statement
//   ProverStatement::new(statement, traces)                       miden::ProverStatement   (shape errors -> InstanceError)
//   Preprocessed::build(&statement, &config) / .commitment()      miden::Preprocessed::build / commitment()
//   ProverInstance::new(&config, &prover_statement, preprocessed) miden::ProverInstance
//   ProverInstance::prove(challenger) -> StarkOutput              ProverInstance::prove(challenger) -> StarkOutput
//   StarkProofData { log_trace_heights, transcript }              miden::StarkProofData
//   ProverError::{Instance, Domain, ..}  prover/mod.rs:582-596    miden::ProverError { kind, what() }
//   HashFunction::{Blake3_256, Keccak, Rpo256, Poseidon2, Rpx256} -> the config constructor (prover/src/lib.rs:245-301)
//                                                                 miden::HashFunction + StarkConfig::with_hash
//
// Header-only; link with -lmiden_b200.  There is no CPU fallback: constructing a StarkConfig without a usable
// CUDA device throws ProverError{NoDevice}.
#pragma once
#include "../../include/miden_b200.h"
#include <array>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace miden {

using Felt = uint64_t;                       // canonical Goldilocks element (crates/field/src/native/mod.rs:58)
using QuadFelt = std::array<Felt, 2>;        // (c0, c1) of F[u]/(u^2 - 7)
using Commitment = std::array<Felt, 4>;      // Hash<Felt, Felt, 4>

struct ProverError : std::runtime_error {
    enum Kind { Instance, Domain, Cuda, Unsupported, AuxBuilder, NoDevice, ExternalAssertion } kind;
    ProverError(Kind k, const std::string& m) : std::runtime_error(m), kind(k) {}
    static Kind from_status(int rc) {
        switch (rc) {
            case MDN_ERR_DOMAIN: return Domain;
            case MDN_ERR_CUDA: return Cuda;
            case MDN_ERR_UNSUPPORTED: return Unsupported;
            case MDN_ERR_AUX_BUILDER: return AuxBuilder;
            case MDN_ERR_NO_DEVICE: return NoDevice;
            case MDN_ERR_EXTERNAL_ASSERTION: return ExternalAssertion;
            default: return Instance;
        }
    }
};

struct PcsParams {
    uint32_t log_blowup = 3, log_folding_arity = 2, log_final_degree = 7;
    uint32_t folding_pow_bits = 4, deep_pow_bits = 12, num_queries = 27, query_pow_bits = 16;   // air/src/config.rs:55-67
    mdn_pcs_params raw() const { return {log_blowup, log_folding_arity, log_final_degree, folding_pow_bits, deep_pow_bits, num_queries, query_pow_bits}; }
};

// `HashFunction` of miden_prover::prove_stark's match (prover/src/lib.rs:245-301) = which miden_air::config constructor applies
enum class HashFunction { Poseidon2 = MDN_HASH_POSEIDON2, Blake3_256 = MDN_HASH_BLAKE3, Keccak = MDN_HASH_KECCAK, Rpo256 = MDN_HASH_RPO, Rpx256 = MDN_HASH_RPX };

// p3 DuplexChallenger<Felt, Poseidon2, 12, 8> state (air/src/config.rs:223,264-271)
struct Challenger {
    mdn_challenger raw{};
    void observe(Felt x) { mdn_challenger_observe(&raw, &x, 1); }
    void observe_slice(const std::vector<Felt>& xs) { if (!xs.empty()) mdn_challenger_observe(&raw, xs.data(), xs.size()); }
    Felt sample() { return mdn_challenger_sample(&raw); }
};

// RowMajorMatrix<Felt> view: height = 2^log_height
struct RowMajorMatrix {
    std::vector<Felt> values;
    uint32_t log_height = 0, width = 0;
    RowMajorMatrix() {}
    RowMajorMatrix(std::vector<Felt> v, uint32_t w) : values(std::move(v)), width(w) {
        size_t h = w ? values.size() / w : 0;
        if (!w || h * w != values.size() || (h & (h - 1)) || !h) throw ProverError(ProverError::Instance, "matrix height must be a power of two");
        while ((size_t(1) << log_height) < h) log_height++;
    }
    size_t height() const { return size_t(1) << log_height; }
    mdn_matrix raw() const { return {values.data(), log_height, width}; }
};

// Recording builder for `LiftedAir::eval` (the symbolic capture a Rust shim performs once per AIR)
class AirBuilder {
public:
    struct Expr { uint32_t id; };
    Expr main(uint32_t offset, uint32_t col) { return node(0, offset, col); }
    Expr aux(uint32_t offset, uint32_t col) { return node(1, offset, col); }
    Expr public_value(uint32_t i) { return node(2, i, 0); }
    Expr challenge(uint32_t i) { return node(3, i, 0); }
    Expr aux_value(uint32_t i) { return node(4, i, 0); }
    Expr is_first_row() { return node(5, 0, 0); }
    Expr is_last_row() { return node(6, 0, 0); }
    Expr is_transition() { return node(7, 0, 0); }
    Expr constant(Felt v) { consts_.push_back(v); return node(8, (uint32_t)consts_.size() - 1, 0); }
    Expr periodic(uint32_t col) { return node(14, col, 0); }
    Expr preprocessed(uint32_t offset, uint32_t col) { return node(15, offset, col); }
    Expr add(Expr a, Expr b) { return node(10, a.id, b.id); }
    Expr sub(Expr a, Expr b) { return node(11, a.id, b.id); }
    Expr mul(Expr a, Expr b) { return node(12, a.id, b.id); }
    Expr neg(Expr a) { return node(13, a.id, 0); }
    void assert_zero(Expr e) { constraints_.push_back(e.id); }
    void assert_zero_ext(Expr e) { constraints_.push_back(e.id); }
    std::vector<uint32_t> finish() const {
        std::vector<uint32_t> w = {0x5249414Du, 1u, (uint32_t)(nodes_.size() / 3), (uint32_t)constraints_.size(), (uint32_t)consts_.size()};
        w.insert(w.end(), nodes_.begin(), nodes_.end());
        w.insert(w.end(), constraints_.begin(), constraints_.end());
        for (Felt c : consts_) { w.push_back((uint32_t)c); w.push_back((uint32_t)(c >> 32)); }
        return w;
    }
private:
    Expr node(uint32_t op, uint32_t a, uint32_t b) { nodes_.insert(nodes_.end(), {op, a, b}); return Expr{(uint32_t)(nodes_.size() / 3 - 1)}; }
    std::vector<uint32_t> nodes_, constraints_;
    std::vector<Felt> consts_;
};

// One AIR of the MultiAir: the shape queries of LiftedAir plus the lowered eval
struct Air {
    uint32_t width = 0, aux_width = 0, num_aux_values = 0, num_randomness = 0, log_quotient_degree = 0, preprocessed_width = 0;
    std::vector<uint32_t> program;                 // AirBuilder::finish()
    std::vector<Felt> periodic_values;             // row-major (1 << log_max_period) x num_periodic_columns
    uint32_t num_periodic_columns = 0, log_max_period = 0;
    std::vector<uint32_t> lookup_program;          // lowered LookupAir (empty: aux trace from the host builder)
    uint32_t lookup_columns = 0;
    RowMajorMatrix preprocessed_trace;             // BaseAir::preprocessed_trace(); width 0 = None
};

struct Statement {
    std::vector<Air> airs;
    std::vector<Felt> air_inputs;                  // public values
    std::vector<Felt> observe_felts;               // what Statement::observe absorbs (AIR specific)
    // default MultiAir::observe: len(air_inputs), air_inputs, max_aux_inputs (0), len(aux_inputs) (0)
    static Statement with_default_observe(std::vector<Air> airs, std::vector<Felt> air_inputs) {
        Statement s; s.airs = std::move(airs); s.air_inputs = std::move(air_inputs);
        s.observe_felts.push_back(s.air_inputs.size());
        s.observe_felts.insert(s.observe_felts.end(), s.air_inputs.begin(), s.air_inputs.end());
        s.observe_felts.push_back(0); s.observe_felts.push_back(0);
        return s;
    }
};

struct ProverStatement {
    Statement statement;
    std::vector<RowMajorMatrix> traces;            // instance order
    ProverStatement(Statement s, std::vector<RowMajorMatrix> t) : statement(std::move(s)), traces(std::move(t)) {
        if (traces.size() != statement.airs.size()) throw ProverError(ProverError::Instance, "trace count does not match the AIR count");
        for (size_t i = 0; i < traces.size(); i++)
            if (traces[i].width != statement.airs[i].width) throw ProverError(ProverError::Instance, "trace width does not match its AIR");
    }
};

// GenericStarkConfig: PCS parameters + the pre-bound challenger prototype; owns the device session
class StarkConfig {
public:
    StarkConfig(PcsParams params, Challenger challenger_prototype, int cuda_device = 0) : params_(params), proto_(challenger_prototype) {
        mdn_pcs_params p = params.raw();
        mdn_session* s = nullptr;
        int rc = mdn_session_create(&p, cuda_device, &s);
        if (rc != MDN_OK) throw ProverError(ProverError::from_status(rc), mdn_last_error(nullptr));
        session_.reset(s, mdn_session_destroy);
    }
    // blake3_256_config / keccak_config / rpo_config / rpx_config instead of poseidon2_config.  For the two byte-oriented
    // configurations `hash_challenger_input` is the HashChallenger's input buffer after config.challenger() +
    // observe_protocol_params (relation digest + 8 parameter felts, little-endian u64) and the Challenger prototype is unused;
    // for RPO / RPX the prototype must be the duplex state built over that permutation.
    StarkConfig& with_hash(HashFunction h, const std::vector<uint8_t>& hash_challenger_input = {}) {
        int rc = mdn_session_set_hash(session_.get(), (mdn_hash_kind)h);
        if (rc != MDN_OK) throw ProverError(ProverError::from_status(rc), mdn_last_error(session_.get()));
        if (h == HashFunction::Blake3_256 || h == HashFunction::Keccak) {
            mdn_hash_challenger hc{hash_challenger_input.data(), hash_challenger_input.size(), nullptr, 0};
            rc = mdn_session_set_hash_challenger(session_.get(), &hc);
            if (rc != MDN_OK) throw ProverError(ProverError::from_status(rc), mdn_last_error(session_.get()));
        }
        hash_ = h;
        return *this;
    }
    HashFunction hash() const { return hash_; }
    bool hash_challenger() const { return hash_ == HashFunction::Blake3_256 || hash_ == HashFunction::Keccak; }
    const PcsParams& pcs() const { return params_; }
    Challenger challenger() const { return proto_; }
    mdn_session* session() const { return session_.get(); }
    const std::shared_ptr<mdn_session>& shared_session() const { return session_; }
private:
    PcsParams params_; Challenger proto_;
    HashFunction hash_ = HashFunction::Poseidon2;
    std::shared_ptr<mdn_session> session_;
};

struct TranscriptData { std::vector<Felt> fields; std::vector<Commitment> commitments; };
struct StarkProofData { std::vector<uint8_t> log_trace_heights; TranscriptData transcript; };
struct StarkOutput { StarkProofData proof; };

// LiftedAir::build_aux_trace for the AIRs that do not ship a lowered LookupAir
using AuxBuilder = std::function<void(uint32_t instance, const RowMajorMatrix& main, const std::vector<QuadFelt>& challenges,
                                      std::vector<Felt>& aux_flat /* height x 2*aux_width */, std::vector<QuadFelt>& aux_values)>;

namespace detail {
struct Lowered {                                   // C structs pointing into a Statement
    std::vector<mdn_lookup> lookups; std::vector<mdn_air> airs; mdn_statement st{};
    explicit Lowered(const Statement& s) {
        lookups.resize(s.airs.size()); airs.resize(s.airs.size());
        for (size_t i = 0; i < s.airs.size(); i++) {
            const Air& a = s.airs[i];
            mdn_air& r = airs[i];
            r = mdn_air{};
            r.width = a.width; r.aux_width = a.aux_width; r.num_aux_values = a.num_aux_values; r.num_randomness = a.num_randomness;
            r.log_quotient_degree = a.log_quotient_degree; r.program_words = (uint32_t)a.program.size(); r.program = a.program.data();
            r.periodic_values = a.periodic_values.empty() ? nullptr : a.periodic_values.data();
            r.num_periodic_columns = a.num_periodic_columns; r.log_max_period = a.log_max_period; r.preprocessed_width = a.preprocessed_width;
            if (!a.lookup_program.empty()) {
                lookups[i] = mdn_lookup{a.lookup_columns, (uint32_t)a.lookup_program.size(), a.lookup_program.data()};
                r.lookup = &lookups[i];
            }
        }
        st.airs = airs.data(); st.n_airs = (uint32_t)airs.size();
        st.public_values = s.air_inputs.data(); st.n_public_values = (uint32_t)s.air_inputs.size();
        st.observe_felts = s.observe_felts.data(); st.n_observe_felts = (uint32_t)s.observe_felts.size();
    }
};
inline void check(const StarkConfig& c, int rc) {
    if (rc != MDN_OK) throw ProverError(ProverError::from_status(rc), mdn_last_error(c.session()));
}
}  // namespace detail

// Preprocessed::build(&statement, &config): the LDE tree lives on the device inside the config's session
class Preprocessed {
public:
    // None when no AIR declares preprocessed columns (preprocessed.rs:83-89)
    static std::unique_ptr<Preprocessed> build(const Statement& s, const StarkConfig& config) {
        bool any = false;
        for (const Air& a : s.airs) any |= a.preprocessed_width > 0;
        if (!any) return nullptr;
        detail::Lowered low(s);
        std::vector<mdn_matrix> mats;
        for (const Air& a : s.airs) mats.push_back(a.preprocessed_width ? a.preprocessed_trace.raw() : mdn_matrix{nullptr, 0, 0});
        auto p = std::unique_ptr<Preprocessed>(new Preprocessed());
        detail::check(config, mdn_session_set_preprocessed(config.session(), &low.st, mats.data(), p->commitment_.data()));
        p->session_ = config.shared_session();
        return p;
    }
    // the bundle lives in the session while this object does (the reference lends `&Preprocessed` to each ProverInstance)
    ~Preprocessed() { if (session_) mdn_session_set_preprocessed(session_.get(), nullptr, nullptr, nullptr); }
    Preprocessed(const Preprocessed&) = delete;
    Preprocessed& operator=(const Preprocessed&) = delete;
    const Commitment& commitment() const { return commitment_; }
private:
    Preprocessed() {}
    Commitment commitment_{};
    std::shared_ptr<mdn_session> session_;
};

class ProverInstance {
public:
    // `preprocessed` must be non-null exactly when some AIR declares preprocessed columns (PresenceMismatch otherwise)
    ProverInstance(const StarkConfig& config, const ProverStatement& ps, const Preprocessed* preprocessed, AuxBuilder aux = nullptr)
        : config_(config), ps_(ps), aux_(std::move(aux)) {
        bool expected = false;
        for (const Air& a : ps.statement.airs) expected |= a.preprocessed_width > 0;
        if (expected != (preprocessed != nullptr)) throw ProverError(ProverError::Instance, "preprocessed presence mismatch");
    }
    StarkOutput prove(const Challenger& challenger) const {
        detail::Lowered low(ps_.statement);
        std::vector<mdn_matrix> mats;
        for (const RowMajorMatrix& t : ps_.traces) mats.push_back(t.raw());
        mdn_proof proof{};
        int rc = mdn_prove(config_.session(), &low.st, mats.data(), config_.hash_challenger() ? nullptr : &challenger.raw, aux_ ? &ProverInstance::trampoline : nullptr,
                           (void*)this, 0, &proof);
        detail::check(config_, rc);
        StarkOutput out;
        out.proof.log_trace_heights.assign(proof.log_trace_heights, proof.log_trace_heights + proof.n_heights);
        out.proof.transcript.fields.assign(proof.fields, proof.fields + proof.n_fields);
        out.proof.transcript.commitments.resize(proof.n_commitments);
        for (size_t i = 0; i < proof.n_commitments; i++) for (int k = 0; k < 4; k++) out.proof.transcript.commitments[i][k] = proof.commitments[4 * i + k];
        return out;
    }
private:
    static int trampoline(void* ctx, uint32_t instance, const mdn_matrix* main, const uint64_t* randomness, uint64_t* aux_out, uint64_t* aux_values) {
        const ProverInstance* self = (const ProverInstance*)ctx;
        try {
            const Air& a = self->ps_.statement.airs[instance];
            std::vector<QuadFelt> ch(a.num_randomness);
            for (uint32_t i = 0; i < a.num_randomness; i++) ch[i] = {randomness[2 * i], randomness[2 * i + 1]};
            std::vector<Felt> flat((size_t(1) << main->log_height) * 2 * a.aux_width, 0);
            std::vector<QuadFelt> vals(a.num_aux_values, QuadFelt{0, 0});
            self->aux_(instance, self->ps_.traces[instance], ch, flat, vals);
            for (size_t i = 0; i < flat.size(); i++) aux_out[i] = flat[i];
            for (size_t i = 0; i < vals.size(); i++) { aux_values[2 * i] = vals[i][0]; aux_values[2 * i + 1] = vals[i][1]; }
            return 0;
        } catch (...) { return -1; }
    }
    const StarkConfig& config_; const ProverStatement& ps_; AuxBuilder aux_;
};

}  // namespace miden
