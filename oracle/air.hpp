// TEST INFRASTRUCTURE ONLY (see field.hpp header).
// AIR constraint programs: a flat op-list evaluated at one point of the quotient domain.
//
// The reference evaluates `air.eval(builder)` (generic Rust) at every point
// (crates/lifted-stark/src/prover/constraints/mod.rs:83-278).  A C-ABI backend cannot call Rust
// generics, so the host shim lowers `eval` once to a DAG with the leaf/op vocabulary the reference
// already enumerates in crates/ace-codegen/src/dag/lower.rs:109-210, and ships it as the op-list
// below (format shared with include/miden_b200.h).  Constraints are folded in emission order
// with the verifier's orientation acc <- acc*alpha + C_k
// (crates/lifted-stark/src/verifier/constraints.rs:83,108).
#pragma once
#include "field.hpp"
#include <stdexcept>
#include <string>

namespace orc {

enum AirOp : uint32_t {
    OP_MAIN = 0,        // a = row offset (0 local, 1 next), b = column          -> F
    OP_AUX = 1,         // a = row offset, b = EF column (base cols 2b, 2b+1)     -> EF
    OP_PUBLIC = 2,      // a = public value index                                  -> F
    OP_CHALLENGE = 3,   // a = randomness index                                    -> EF
    OP_AUX_VALUE = 4,   // a = aux (permutation) value index                       -> EF
    OP_IS_FIRST = 5,    //                                                          -> F
    OP_IS_LAST = 6,     //                                                          -> F
    OP_IS_TRANSITION = 7,  //                                                       -> F
    OP_CONST = 8,       // a = constant-pool index                                 -> F
    OP_EXT_CONST = 9,   // a = constant-pool index of (c0, c1)                     -> EF
    OP_ADD = 10, OP_SUB = 11, OP_MUL = 12,  // a, b = node ids; EF if either is EF
    OP_NEG = 13,        // a = node id
    OP_PERIODIC = 14,   // a = periodic column index                               -> F
    OP_PREPROCESSED = 15,  // a = row offset, b = preprocessed column                 -> F
};

struct AirNode { uint32_t op, a, b; };

struct AirProgram {
    std::vector<AirNode> nodes;
    std::vector<uint32_t> constraints;   // node ids, emission order
    std::vector<u64> consts;
    std::vector<uint8_t> is_ext;         // derived per node

    static constexpr uint32_t MAGIC = 0x5249414Du;  // "MAIR"

    // Serialized form: u32 words [MAGIC, version=1, n_nodes, n_constraints, n_consts,
    //   nodes (3 words each), constraints, consts (lo, hi words each)].
    static AirProgram parse(const uint32_t* w, size_t n_words) {
        if (n_words < 5 || w[0] != MAGIC || w[1] != 1) throw std::runtime_error("air program: bad header");
        size_t nn = w[2], nc = w[3], nk = w[4];
        if (n_words != 5 + 3 * nn + nc + 2 * nk) throw std::runtime_error("air program: bad length");
        AirProgram p;
        const uint32_t* q = w + 5;
        for (size_t i = 0; i < nn; i++, q += 3) p.nodes.push_back({q[0], q[1], q[2]});
        for (size_t i = 0; i < nc; i++) p.constraints.push_back(*q++);
        for (size_t i = 0; i < nk; i++, q += 2) p.consts.push_back((u64)q[0] | ((u64)q[1] << 32));
        p.is_ext.resize(nn);
        for (size_t i = 0; i < nn; i++) {
            const AirNode& nd = p.nodes[i];
            switch (nd.op) {
                case OP_AUX: case OP_CHALLENGE: case OP_AUX_VALUE: case OP_EXT_CONST: p.is_ext[i] = 1; break;
                case OP_ADD: case OP_SUB: case OP_MUL:
                    if (nd.a >= i || nd.b >= i) throw std::runtime_error("air program: forward reference");
                    p.is_ext[i] = p.is_ext[nd.a] | p.is_ext[nd.b]; break;
                case OP_NEG:
                    if (nd.a >= i) throw std::runtime_error("air program: forward reference");
                    p.is_ext[i] = p.is_ext[nd.a]; break;
                default:
                    if (nd.op > OP_PREPROCESSED) throw std::runtime_error("air program: unknown op");
                    p.is_ext[i] = 0;
            }
        }
        for (uint32_t c : p.constraints) if (c >= nn) throw std::runtime_error("air program: bad constraint id");
        return p;
    }
};

// Lowering of `LookupAir::eval` (reference air/src/lookup/builder.rs): the node vocabulary of the constraint
// op-list restricted to what a LookupBuilder exposes (main window, periodic values, public values, the two
// challenges, constants, arithmetic), plus one record per interaction: the aux column it belongs to, its 0/1
// flag (NONE = unconditional), its signed base-field multiplicity and its encoded extension-field
// denominator -- `ProverGroup::insert` / `ProverBatch::insert` (air/src/lookup/prover.rs:338-362,421-444)
// push exactly (multiplicity, denominator) when the flag is non-zero.
struct LookupInteraction { uint32_t column, flag, multiplicity, denominator; };
struct LookupProgram {
    uint32_t num_columns = 0;
    std::vector<AirNode> nodes;
    std::vector<LookupInteraction> interactions;
    std::vector<u64> consts;
    static constexpr uint32_t MAGIC = 0x504B4C4Du;   // "MLKP"
    static constexpr uint32_t NO_FLAG = 0xFFFFFFFFu;
    bool present() const { return num_columns > 0; }

    // words [MAGIC, 1, n_nodes, n_interactions, n_consts, nodes (3 each), interactions (4 each), consts (lo, hi)]
    static LookupProgram parse(uint32_t num_columns, const uint32_t* w, size_t n_words) {
        if (n_words < 5 || w[0] != MAGIC || w[1] != 1) throw std::runtime_error("lookup program: bad header");
        size_t nn = w[2], ni = w[3], nk = w[4];
        if (n_words != 5 + 3 * nn + 4 * ni + 2 * nk) throw std::runtime_error("lookup program: bad length");
        LookupProgram p; p.num_columns = num_columns;
        const uint32_t* q = w + 5;
        std::vector<uint8_t> is_ext(nn);
        for (size_t i = 0; i < nn; i++, q += 3) {
            AirNode nd{q[0], q[1], q[2]};
            switch (nd.op) {
                case OP_MAIN: case OP_PUBLIC: case OP_CONST: case OP_PERIODIC: is_ext[i] = 0; break;
                case OP_CHALLENGE: case OP_EXT_CONST: is_ext[i] = 1; break;
                case OP_ADD: case OP_SUB: case OP_MUL:
                    if (nd.a >= i || nd.b >= i) throw std::runtime_error("lookup program: forward reference");
                    is_ext[i] = is_ext[nd.a] | is_ext[nd.b]; break;
                case OP_NEG:
                    if (nd.a >= i) throw std::runtime_error("lookup program: forward reference");
                    is_ext[i] = is_ext[nd.a]; break;
                default: throw std::runtime_error("lookup program: op not available to a LookupBuilder");
            }
            p.nodes.push_back(nd);
        }
        for (size_t i = 0; i < ni; i++, q += 4) {
            LookupInteraction it{q[0], q[1], q[2], q[3]};
            if (it.column >= num_columns) throw std::runtime_error("lookup program: column out of range");
            if ((it.flag != NO_FLAG && (it.flag >= nn || is_ext[it.flag])) || it.multiplicity >= nn || is_ext[it.multiplicity] || it.denominator >= nn)
                throw std::runtime_error("lookup program: bad interaction");
            p.interactions.push_back(it);
        }
        for (size_t i = 0; i < nk; i++, q += 2) p.consts.push_back((u64)q[0] | ((u64)q[1] << 32));
        return p;
    }
};

// Everything a program can read at one evaluation point.
struct AirPoint {
    const Fp* main_local; const Fp* main_next;
    const Fp* aux_local; const Fp* aux_next;   // base-field layout: EF column c = (aux[2c], aux[2c+1])
    const Fp* publics; const Ef* challenges; const Ef* aux_values;
    const Ef* periodic = nullptr;              // periodic column values at this point
    const Fp* prep_local = nullptr; const Fp* prep_next = nullptr;        // preprocessed window (base)
    const Ef* prep_local_ef = nullptr; const Ef* prep_next_ef = nullptr;  // ... at the OOD point
    Ef is_first, is_last, is_transition;       // EF so the same evaluator serves the OOD check
    // For the OOD check main/aux cells are EF; then these are used instead of the Fp pointers.
    const Ef* main_local_ef = nullptr; const Ef* main_next_ef = nullptr;
    const Ef* aux_local_ef = nullptr; const Ef* aux_next_ef = nullptr;
};

// Evaluate all nodes; returns the alpha-folded constraint sum.
inline Ef air_eval_folded(const AirProgram& p, const AirPoint& pt, Ef alpha, std::vector<Ef>& scratch) {
    scratch.resize(p.nodes.size());
    for (size_t i = 0; i < p.nodes.size(); i++) {
        const AirNode& nd = p.nodes[i];
        Ef v;
        switch (nd.op) {
            case OP_MAIN:
                if (pt.main_local_ef) v = (nd.a ? pt.main_next_ef : pt.main_local_ef)[nd.b];
                else v = Ef((nd.a ? pt.main_next : pt.main_local)[nd.b]);
                break;
            case OP_AUX:
                if (pt.aux_local_ef) v = (nd.a ? pt.aux_next_ef : pt.aux_local_ef)[nd.b];
                else { const Fp* r = nd.a ? pt.aux_next : pt.aux_local; v = Ef(r[2 * nd.b], r[2 * nd.b + 1]); }
                break;
            case OP_PUBLIC: v = Ef(pt.publics[nd.a]); break;
            case OP_CHALLENGE: v = pt.challenges[nd.a]; break;
            case OP_AUX_VALUE: v = pt.aux_values[nd.a]; break;
            case OP_IS_FIRST: v = pt.is_first; break;
            case OP_IS_LAST: v = pt.is_last; break;
            case OP_IS_TRANSITION: v = pt.is_transition; break;
            case OP_CONST: v = Ef(Fp(p.consts[nd.a])); break;
            case OP_EXT_CONST: v = Ef(Fp(p.consts[nd.a]), Fp(p.consts[nd.a + 1])); break;
            case OP_ADD: v = scratch[nd.a] + scratch[nd.b]; break;
            case OP_SUB: v = scratch[nd.a] - scratch[nd.b]; break;
            case OP_MUL: v = scratch[nd.a] * scratch[nd.b]; break;
            case OP_NEG: v = -scratch[nd.a]; break;
            case OP_PERIODIC: v = pt.periodic[nd.a]; break;
            case OP_PREPROCESSED:
                if (pt.prep_local_ef) v = (nd.a ? pt.prep_next_ef : pt.prep_local_ef)[nd.b];
                else v = Ef((nd.a ? pt.prep_next : pt.prep_local)[nd.b]);
                break;
        }
        scratch[i] = v;
    }
    Ef acc;
    for (uint32_t c : p.constraints) acc = acc * alpha + scratch[c];
    return acc;
}

}  // namespace orc
