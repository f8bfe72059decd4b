// Run-time specialisation of the constraint evaluator (a4) for large AIRs.
//
// The op-list interpreter (k_constraints) pays an instruction fetch, a switch and local-memory slot traffic
// per node -- fine for the 19-node DummyMidenAir, ~4x too slow for the ~5 k-node Miden AIRs
// (air/src/lib.rs:342-355,442-454,523-528).  For programs above a node threshold the session lowers the SAME
// validated op-list into straight-line CUDA C++ (typed: base values are u64, extension values E2; constants
// are literals; the alpha fold uses precomputed alpha powers, which is the reference's own base/ext split,
// prover/constraints/folder.rs:88-105), compiles it once per AIR with NVRTC for sm_100a and launches the
// cubin through the driver API.  Field arithmetic is exact, so the result is bit-identical to the interpreter
// (asserted in tests/test_gpu_parity.py).  libnvrtc / libcuda are dlopen'ed on first use; when NVRTC is
// missing the session keeps using the interpreter and says so in mdn_get_info(MDN_INFO_JIT).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace jit {

// Kernel arguments: layout must equal `struct JitArgs` in PRELUDE below (pointers and u64 first, then u32).
struct JitArgs {
    const uint64_t* main_lde; const uint64_t* aux_lde; const uint64_t* prep_lde;
    const uint64_t* publics; const uint64_t* challenges; const uint64_t* aux_values;
    const uint64_t* periodic; const uint64_t* apow;
    const uint64_t* acc_in; uint64_t* acc_out;
    const uint64_t* w_hi; const uint64_t* w_lo;
    uint64_t shift, w_l, w_h_inv;
    uint64_t zh[16], inv_zh[16];
    uint64_t beta_a, beta_b;
    uint32_t log_n, log_b, acc_in_log_n, lo_bits, log_max_period, pad;
};

// Arguments of the generated LogUp row kernel (same member names as JitArgs where the leaf code is shared).
struct LookupJitArgs {
    const uint64_t* main_lde;      // raw main trace, column-major [col][row]
    const uint64_t* publics; const uint64_t* challenges; const uint64_t* periodic;   // periodic: raw row-major matrix
    uint64_t* aux_cm; uint64_t* totals;
    uint32_t* bad_flag;
    uint32_t log_n, n_periodic, log_max_period, pad;
};

static const char PRELUDE[] = R"CUDA(
typedef unsigned long long u64;
typedef unsigned int u32;
struct E2 { u64 a, b; };
struct JitArgs {
    const u64* main_lde; const u64* aux_lde; const u64* prep_lde;
    const u64* publics; const u64* challenges; const u64* aux_values;
    const u64* periodic; const u64* apow;
    const u64* acc_in; u64* acc_out;
    const u64* w_hi; const u64* w_lo;
    u64 shift, w_l, w_h_inv;
    u64 zh[16], inv_zh[16];
    u64 beta_a, beta_b;
    u32 log_n, log_b, acc_in_log_n, lo_bits, log_max_period, pad;
};
#define GP 0xFFFFFFFF00000001ull
// canonical (< p) Goldilocks arithmetic on the carry flag
__device__ __forceinline__ u64 fsub(u64 a, u64 b) {
    u64 d; u32 m;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(d), "=r"(m) : "l"(a), "l"(b));
    return d - (u64)m;
}
__device__ __forceinline__ u64 fadd(u64 a, u64 b) { return fsub(a, GP - b); }
__device__ __forceinline__ u64 fneg(u64 a) { return a ? GP - a : 0ull; }
)CUDA";
// first-generation multiplication (the one every JIT measurement so far ran with)
static const char PRELUDE_FMUL_G1[] = R"CUDA(__device__ __forceinline__ u64 fmul(u64 x, u64 y) {
    u64 lo = x * y, hi = __umul64hi(x, y);
    u32 hl = (u32)hi, hh = (u32)(hi >> 32);
    u64 t, m, r; u32 b, c;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(t), "=r"(b) : "l"(lo), "l"((u64)hh));
    t -= (u64)b;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(m) : "r"(hl), "r"(0xFFFFFFFFu));
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(t), "l"(m));
    r += (u64)(0u - c);
    u64 q; u32 k;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(q), "=r"(k) : "l"(r), "l"(GP));
    return q - (u64)k;
}
)CUDA";
// second-generation multiplication (poseidon2_fast2.cuh: one 128-bit product, the carry folded by an IMAD.WIDE);
// the default since it was measured on the constraint kernels (tools/realistic_air_probe.py, profiles/r2_tuning.md)
static const char PRELUDE_FMUL_G2[] = R"CUDA(__device__ __forceinline__ u64 fmul(u64 x, u64 y) {
    unsigned __int128 q = (unsigned __int128)x * y;
    u64 lo = (u64)q, hi = (u64)(q >> 64);
    u32 hl = (u32)hi, hh = (u32)(hi >> 32);
    u64 t, m, r; u32 b, c;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(t), "=r"(b) : "l"(lo), "l"((u64)hh));
    t -= (u64)b;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(m) : "r"(hl), "r"(0xFFFFFFFFu));
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(t), "l"(m));
    r = (u64)c * 0xFFFFFFFFull + r;
    u64 q2; u32 k;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(q2), "=r"(k) : "l"(r), "l"(GP));
    return q2 - (u64)k;
}
)CUDA";
static const char PRELUDE_B[] = R"CUDA(__device__ __forceinline__ u64 fmul7(u64 x) { return fmul(x, 7ull); }
__device__ u64 fpow(u64 b, u64 e) { u64 r = 1; while (e) { if (e & 1) r = fmul(r, b); b = fmul(b, b); e >>= 1; } return r; }
__device__ u64 finv(u64 a) { return fpow(a, GP - 2); }
__device__ __forceinline__ E2 mk(u64 a, u64 b) { E2 r; r.a = a; r.b = b; return r; }
__device__ __forceinline__ E2 eadd(E2 x, E2 y) { return mk(fadd(x.a, y.a), fadd(x.b, y.b)); }
__device__ __forceinline__ E2 esub(E2 x, E2 y) { return mk(fsub(x.a, y.a), fsub(x.b, y.b)); }
__device__ __forceinline__ E2 eneg(E2 x) { return mk(fneg(x.a), fneg(x.b)); }
__device__ __forceinline__ E2 emul(E2 x, E2 y) {
    return mk(fadd(fmul(x.a, y.a), fmul7(fmul(x.b, y.b))), fadd(fmul(x.a, y.b), fmul(x.b, y.a)));
}
__device__ __forceinline__ E2 emulf(E2 x, u64 s) { return mk(fmul(x.a, s), fmul(x.b, s)); }
__device__ __forceinline__ E2 eaddf(E2 x, u64 s) { return mk(fadd(x.a, s), x.b); }
__device__ __forceinline__ E2 esubf(E2 x, u64 s) { return mk(fsub(x.a, s), x.b); }      // x - s
__device__ __forceinline__ E2 efsub(u64 s, E2 x) { return mk(fsub(s, x.a), fneg(x.b)); } // s - x
__device__ __forceinline__ E2 ldapow(const u64* __restrict__ p, u32 k) { return mk(p[2 * k], p[2 * k + 1]); }
__device__ E2 einv(E2 x) { u64 n = fsub(fmul(x.a, x.a), fmul7(fmul(x.b, x.b))); u64 ni = finv(n); return mk(fmul(x.a, ni), fneg(fmul(x.b, ni))); }
struct LookupJitArgs {
    const u64* main_lde; const u64* publics; const u64* challenges; const u64* periodic;
    u64* aux_cm; u64* totals;
    u32* bad_flag;
    u32 log_n, n_periodic, log_max_period, pad;
};
)CUDA";

// ---------------------------------------------------------------------------------------------
// code generation
// ---------------------------------------------------------------------------------------------
struct GenInfo { uint32_t n_constraints = 0; bool uses_sel = false; uint32_t n_chunks = 0, spill_base = 0, spill_ext = 0; };

// `w` is an op-list already validated by compile_oplist (session.cu).
//
// NVRTC's optimiser is superlinear in the size of one function (2 k nodes: 5 s, 10 k nodes: 8 min in one basic
// block), so the arithmetic nodes are cut into chunks of CHUNK nodes, one `__noinline__` device function each.
// Leaves (trace cells, constants, challenges ...) are re-materialised in every chunk that reads them; arithmetic
// values that live across a chunk boundary travel through two small per-thread arrays (Sb: base, Se: extension)
// whose slots are assigned by chunk-granular liveness.
inline bool jit_arith2() { const char* e = getenv("MDN_JIT_ARITH"); return !(e && atoi(e) == 1); }   // default since r2: 44.2 -> 39.9 ms on the 7.3 k-node probe (profiles/r2_tuning.md); MDN_JIT_ARITH=1 restores the first generation
inline uint32_t chunk_nodes() { const char* e = getenv("MDN_JIT_CHUNK"); uint32_t v = e ? (uint32_t)atoi(e) : 0; return v ? v : 512; }

// lookup == true: `w` is a lowered LookupAir ("MLKP": 4-word interactions instead of constraint ids) and the
// kernel is the row kernel of build_logup_aux_trace (one (V, U) rational per aux column, see k_logup_rows).
inline std::string generate(const uint32_t* w, GenInfo* info, bool lookup = false, uint32_t n_cols = 0) {
    const uint32_t nn = w[2], nc = w[3];
    const uint32_t* cons = w + 5 + 3 * (size_t)nn;
    const uint32_t* kw = cons + (lookup ? 4 : 1) * (size_t)nc;
    const uint32_t NOFLAG = 0xFFFFFFFFu;
    // operand nodes of item k (constraint: the node itself; interaction: flag?, multiplicity, denominator)
    auto item_ops = [&](uint32_t k, uint32_t out[3]) -> int {
        if (!lookup) { out[0] = cons[k]; return 1; }
        const uint32_t* it = cons + 4 * (size_t)k;
        int n = 0;
        if (it[1] != NOFLAG) out[n++] = it[1];
        out[n++] = it[2]; out[n++] = it[3];
        return n;
    };
    std::vector<uint8_t> ext(nn, 0), used(nn, 0), leaf(nn, 0);
    bool uses_sel = false;
    auto OP = [&](uint32_t j) { return w[5 + 3 * j]; };
    auto X = [&](uint32_t j) { return w[6 + 3 * j]; };
    auto Y = [&](uint32_t j) { return w[7 + 3 * j]; };
    for (uint32_t j = 0; j < nn; j++) {
        uint32_t op = OP(j);
        if (op >= 10 && op <= 12) ext[j] = ext[X(j)] | ext[Y(j)];
        else if (op == 13) ext[j] = ext[X(j)];
        else { ext[j] = (op == 1 || op == 3 || op == 4 || op == 9); leaf[j] = 1; }
        if (op >= 5 && op <= 7) uses_sel = true;
    }
    for (uint32_t k = 0; k < nc; k++) { uint32_t o[3]; int n = item_ops(k, o); for (int q = 0; q < n; q++) used[o[q]] = 1; }
    for (uint32_t j = nn; j-- > 0;) {
        if (!used[j] || leaf[j]) continue;
        used[X(j)] = 1;
        if (OP(j) != 13) used[Y(j)] = 1;
    }
    // chunk assignment of the arithmetic nodes, in definition order
    const uint32_t NONE = 0xFFFFFFFFu, CHUNK = chunk_nodes();
    std::vector<uint32_t> chunk_of(nn, NONE), last_chunk(nn, 0);
    uint32_t n_arith = 0;
    for (uint32_t j = 0; j < nn; j++) if (used[j] && !leaf[j]) chunk_of[j] = n_arith++ / CHUNK;
    const uint32_t n_chunks = std::max(1u, (n_arith + CHUNK - 1) / CHUNK);
    for (uint32_t j = 0; j < nn; j++) {
        if (chunk_of[j] == NONE) continue;
        last_chunk[j] = std::max(last_chunk[j], chunk_of[j]);
        uint32_t ops[2] = {X(j), OP(j) == 13 ? X(j) : Y(j)};
        for (uint32_t o : ops) if (chunk_of[o] != NONE) last_chunk[o] = std::max(last_chunk[o], chunk_of[j]);
    }
    // every item (constraint fold / interaction) runs in the last chunk that defines one of its operands
    std::vector<uint32_t> item_chunk(nc, 0);
    for (uint32_t k = 0; k < nc; k++) {
        uint32_t o[3]; int n = item_ops(k, o);
        for (int q = 0; q < n; q++) if (chunk_of[o[q]] != NONE) item_chunk[k] = std::max(item_chunk[k], chunk_of[o[q]]);
        for (int q = 0; q < n; q++) if (chunk_of[o[q]] != NONE) last_chunk[o[q]] = std::max(last_chunk[o[q]], item_chunk[k]);
    }
    // slots for values that cross a chunk boundary
    std::vector<uint32_t> slot(nn, NONE);
    std::vector<std::vector<uint32_t>> defined(n_chunks), dying(n_chunks);
    for (uint32_t j = 0; j < nn; j++) if (chunk_of[j] != NONE && last_chunk[j] > chunk_of[j]) { defined[chunk_of[j]].push_back(j); dying[last_chunk[j]].push_back(j); }
    std::vector<uint32_t> free_b, free_e;
    uint32_t nb = 0, ne = 0;
    for (uint32_t c = 0; c < n_chunks; c++) {
        for (uint32_t j : dying[c]) (ext[j] ? free_e : free_b).push_back(slot[j]);   // read at the top of chunk c, free afterwards
        for (uint32_t j : defined[c]) {
            std::vector<uint32_t>& fl = ext[j] ? free_e : free_b;
            if (!fl.empty()) { slot[j] = fl.back(); fl.pop_back(); } else slot[j] = ext[j] ? ne++ : nb++;
        }
    }
    // NB: a slot freed by a value dying in chunk c may be re-used by a value defined in chunk c; chunk c loads
    // all its live-ins before it stores any live-out, so that is safe.
    char buf[256];
    auto nm = [&](uint32_t j) { char b[24]; snprintf(b, sizeof b, "%c%u", ext[j] ? 'e' : 'b', j); return std::string(b); };
    auto leaf_def = [&](uint32_t j) {
        uint32_t op = OP(j), x = X(j), y = Y(j);
        std::string d = std::string("  const ") + (ext[j] ? "E2 " : "u64 ") + nm(j) + " = ";
        switch (op) {
            case 0: snprintf(buf, sizeof buf, "a.main_lde[(size_t)%u * L + %s];\n", y, x ? "pn" : "pos"); break;
            case 1: snprintf(buf, sizeof buf, "mk(a.aux_lde[(size_t)%u * L + %s], a.aux_lde[(size_t)%u * L + %s]);\n", 2 * y, x ? "pn" : "pos", 2 * y + 1, x ? "pn" : "pos"); break;
            case 2: snprintf(buf, sizeof buf, "a.publics[%u];\n", x); break;
            case 3: snprintf(buf, sizeof buf, "mk(a.challenges[%u], a.challenges[%u]);\n", 2 * x, 2 * x + 1); break;
            case 4: snprintf(buf, sizeof buf, "mk(a.aux_values[%u], a.aux_values[%u]);\n", 2 * x, 2 * x + 1); break;
            case 5: snprintf(buf, sizeof buf, "c.is_first;\n"); break;
            case 6: snprintf(buf, sizeof buf, "c.is_last;\n"); break;
            case 7: snprintf(buf, sizeof buf, "c.is_trans;\n"); break;
            case 8: { uint64_t v = (uint64_t)kw[2 * x] | ((uint64_t)kw[2 * x + 1] << 32); snprintf(buf, sizeof buf, "0x%llxull;\n", (unsigned long long)v); break; }
            case 9: {
                uint64_t v0 = (uint64_t)kw[2 * x] | ((uint64_t)kw[2 * x + 1] << 32), v1 = (uint64_t)kw[2 * x + 2] | ((uint64_t)kw[2 * x + 3] << 32);
                snprintf(buf, sizeof buf, "mk(0x%llxull, 0x%llxull);\n", (unsigned long long)v0, (unsigned long long)v1); break;
            }
            case 14: snprintf(buf, sizeof buf, "a.periodic[(size_t)%u * per_stride + per_idx];\n", x); break;
            case 15: snprintf(buf, sizeof buf, "a.prep_lde[(size_t)%u * L + %s];\n", y, x ? "pn" : "pos"); break;
            default: throw std::runtime_error("jit: unknown leaf op");
        }
        return d + buf;
    };
    auto arith_def = [&](uint32_t j) {
        uint32_t op = OP(j), x = X(j), y = Y(j);
        std::string d = std::string("  const ") + (ext[j] ? "E2 " : "u64 ") + nm(j) + " = ";
        if (op == 13) return d + (ext[x] ? "eneg(" : "fneg(") + nm(x) + ");\n";
        const char* fb = op == 10 ? "fadd" : op == 11 ? "fsub" : "fmul";
        const char* fe = op == 10 ? "eadd" : op == 11 ? "esub" : "emul";
        if (!ext[x] && !ext[y]) return d + fb + "(" + nm(x) + ", " + nm(y) + ");\n";
        if (ext[x] && ext[y]) return d + fe + "(" + nm(x) + ", " + nm(y) + ");\n";
        if (ext[x]) return d + (op == 10 ? "eaddf" : op == 11 ? "esubf" : "emulf") + "(" + nm(x) + ", " + nm(y) + ");\n";
        return d + (op == 10 ? "eaddf" : op == 11 ? "efsub" : "emulf") + "(" + (op == 11 ? nm(x) + ", " + nm(y) : nm(y) + ", " + nm(x)) + ");\n";
    };
    // constraint folds: in the chunk of their node (leaf constraints: chunk 0); weights alpha^(K-1-k) make the
    // order irrelevant
    std::vector<std::vector<uint32_t>> folds(n_chunks);
    for (uint32_t k = 0; k < nc; k++) folds[item_chunk[k]].push_back(k);

    std::string s;
    s.reserve(80 * (size_t)nn + 16384);
    s += PRELUDE;
    s += jit_arith2() ? PRELUDE_FMUL_G2 : PRELUDE_FMUL_G1;
    s += PRELUDE_B;
    s += lookup ? "typedef LookupJitArgs KArgs;\n" : "typedef JitArgs KArgs;\n";
    s += "struct Ctx { size_t L, pos, pn, per_idx, per_stride; u64 is_first, is_last, is_trans; };\n";
    std::vector<uint8_t> seen(nn, 0);
    for (uint32_t c = 0; c < n_chunks; c++) {
        if (lookup) snprintf(buf, sizeof buf, "__device__ __noinline__ void chunk%u(const KArgs& a, const Ctx& c, u64* __restrict__ Sb, E2* __restrict__ Se, E2* __restrict__ V, E2* __restrict__ U) {\n", c);
        else snprintf(buf, sizeof buf, "__device__ __noinline__ void chunk%u(const KArgs& a, const Ctx& c, u64* __restrict__ Sb, E2* __restrict__ Se, E2& acc) {\n", c);
        s += buf;
        s += "  const size_t L = c.L, pos = c.pos, pn = c.pn, per_idx = c.per_idx, per_stride = c.per_stride;\n"
             "  (void)L; (void)pos; (void)pn; (void)per_idx; (void)per_stride; (void)Sb; (void)Se;\n";
        // operands needed by this chunk: leaves (re-materialised) and live-in arithmetic values (loaded)
        std::vector<uint32_t> need;
        auto want = [&](uint32_t o) { if (!seen[o]) { seen[o] = 1; need.push_back(o); } };
        for (uint32_t j = 0; j < nn; j++) {
            if (chunk_of[j] != c) continue;
            uint32_t ops[2] = {X(j), OP(j) == 13 ? X(j) : Y(j)};
            for (uint32_t o : ops) if (leaf[o] || chunk_of[o] != c) want(o);
        }
        for (uint32_t k : folds[c]) { uint32_t o[3]; int n = item_ops(k, o); for (int q = 0; q < n; q++) if (leaf[o[q]] || chunk_of[o[q]] != c) want(o[q]); }
        for (uint32_t o : need) {
            if (leaf[o]) s += leaf_def(o);
            else { snprintf(buf, sizeof buf, "  const %s %s = %s[%u];\n", ext[o] ? "E2" : "u64", nm(o).c_str(), ext[o] ? "Se" : "Sb", slot[o]); s += buf; }
            seen[o] = 0;
        }
        for (uint32_t j = 0; j < nn; j++) if (chunk_of[j] == c) s += arith_def(j);
        for (uint32_t k : folds[c]) {
            if (!lookup) {
                uint32_t cn = cons[k];
                snprintf(buf, sizeof buf, "  acc = eadd(acc, %s(ldapow(a.apow, %u), %s));\n", ext[cn] ? "emul" : "emulf", k, nm(cn).c_str());
                s += buf;
                continue;
            }
            // ProverGroup::insert (air/src/lookup/prover.rs:338-362): push (multiplicity, denominator) unless the flag is zero
            const uint32_t* it = cons + 4 * (size_t)k;
            std::string den = ext[it[3]] ? nm(it[3]) : "mk(" + nm(it[3]) + ", 0ull)";
            std::string cond = it[1] == NOFLAG ? "true" : nm(it[1]) + " != 0ull";
            s += "  if (" + cond + ") { const E2 d = " + den + "; if ((d.a | d.b) == 0ull) *a.bad_flag = 2u; else { ";
            snprintf(buf, sizeof buf, "V[%u] = eadd(emul(V[%u], d), emulf(U[%u], %s)); U[%u] = emul(U[%u], d); } }\n", it[0], it[0], it[0], nm(it[2]).c_str(), it[0], it[0]);
            s += buf;
        }
        for (uint32_t j : defined[c]) { snprintf(buf, sizeof buf, "  %s[%u] = %s;\n", ext[j] ? "Se" : "Sb", slot[j], nm(j).c_str()); s += buf; }
        s += "}\n";
    }
    if (lookup) {
        s += "extern \"C\" __global__ void __launch_bounds__(128) k_jit(const KArgs a) {\n"
             "  Ctx c;\n"
             "  c.L = (size_t)1 << a.log_n;\n"
             "  c.pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;\n"
             "  if (c.pos >= c.L) return;\n"
             "  c.pn = (c.pos + 1) & (c.L - 1);\n"
             "  c.per_idx = (c.pos & (((size_t)1 << a.log_max_period) - 1)) * a.n_periodic;\n"
             "  c.per_stride = 1;\n"
             "  c.is_first = c.is_last = c.is_trans = 0ull;\n";
        snprintf(buf, sizeof buf, "  u64 Sb[%u]; E2 Se[%u];\n  E2 V[%u], U[%u];\n  for (int i = 0; i < %u; i++) { V[i] = mk(0ull, 0ull); U[i] = mk(1ull, 0ull); }\n",
                 std::max(1u, nb), std::max(1u, ne), std::max(1u, n_cols), std::max(1u, n_cols), n_cols);
        s += buf;
        for (uint32_t c = 0; c < n_chunks; c++) { snprintf(buf, sizeof buf, "  chunk%u(a, c, Sb, Se, V, U);\n", c); s += buf; }
        snprintf(buf, sizeof buf, "  E2 t = mk(0ull, 0ull);\n  for (u32 i = 0; i < %uu; i++) {\n", n_cols);
        s += buf;
        s += "    E2 f = V[i];\n"
             "    if (!(U[i].a == 1ull && U[i].b == 0ull)) f = emul(V[i], einv(U[i]));\n"
             "    t = eadd(t, f);\n"
             "    if (i > 0) { a.aux_cm[(size_t)(2 * i) * c.L + c.pos] = f.a; a.aux_cm[(size_t)(2 * i + 1) * c.L + c.pos] = f.b; }\n"
             "  }\n"
             "  a.totals[2 * c.pos] = t.a; a.totals[2 * c.pos + 1] = t.b;\n"
             "}\n";
        if (info) { info->n_constraints = nc; info->uses_sel = false; info->n_chunks = n_chunks; info->spill_base = nb; info->spill_ext = ne; }
        return s;
    }
    s += "extern \"C\" __global__ void __launch_bounds__(128) k_jit(const JitArgs a) {\n"
         "  Ctx c;\n"
         "  c.L = (size_t)1 << (a.log_n + a.log_b);\n"
         "  c.pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;\n"
         "  { const u32 ct0 = a.pad & 0xffu, cnt = (a.pad >> 8) & 0xffu;   /* cosets [ct0, ct0 + cnt); cnt == 0: all */\n"
         "    if (c.pos >= (cnt ? (size_t)cnt << a.log_n : c.L)) return;\n"
         "    c.pos += (size_t)ct0 << a.log_n; }\n"
         "  const u32 N = 1u << a.log_n;\n"
         "  const u32 t = (u32)(c.pos >> a.log_n), r = (u32)(c.pos & (N - 1));\n"
         "  c.pn = ((size_t)t << a.log_n) + ((r + 1) & (N - 1));\n"
         "  c.per_idx = ((size_t)(r & ((1u << a.log_max_period) - 1)) << a.log_b) | t;\n"
         "  c.per_stride = (size_t)1 << (a.log_max_period + a.log_b);\n"
         "  c.is_first = c.is_last = c.is_trans = 0ull;\n";
    if (uses_sel)
        s += "  {\n"
             "    u64 x = fmul(fmul(a.shift, fpow(a.w_l, t)), fmul(a.w_hi[r >> a.lo_bits], a.w_lo[r & ((1u << a.lo_bits) - 1)]));\n"
             "    u64 d_first = fsub(x, 1ull), d_last = fsub(x, a.w_h_inv);\n"
             "    u64 inv = finv(fmul(d_first, d_last));\n"
             "    c.is_first = fmul(a.zh[t], fmul(inv, d_last));\n"
             "    c.is_last = fmul(a.zh[t], fmul(inv, d_first));\n"
             "    c.is_trans = d_last;\n"
             "  }\n";
    snprintf(buf, sizeof buf, "  u64 Sb[%u]; E2 Se[%u];\n  E2 acc = mk(0ull, 0ull);\n", std::max(1u, nb), std::max(1u, ne));
    s += buf;
    for (uint32_t c = 0; c < n_chunks; c++) { snprintf(buf, sizeof buf, "  chunk%u(a, c, Sb, Se, acc);\n", c); s += buf; }
    s += "  E2 q = emulf(acc, a.inv_zh[t]);\n"
         "  if (a.acc_in) {\n"
         "    const size_t Lin = (size_t)1 << (a.acc_in_log_n + a.log_b);\n"
         "    const size_t pa = ((size_t)t << a.acc_in_log_n) + (r & ((1u << a.acc_in_log_n) - 1));\n"
         "    q = eadd(emul(mk(a.acc_in[pa], a.acc_in[Lin + pa]), mk(a.beta_a, a.beta_b)), q);\n"
         "  }\n"
         "  a.acc_out[c.pos] = q.a;\n"
         "  a.acc_out[c.L + c.pos] = q.b;\n"
         "}\n";
    if (info) { info->n_constraints = nc; info->uses_sel = uses_sel; info->n_chunks = n_chunks; info->spill_base = nb; info->spill_ext = ne; }
    return s;
}

// ---------------------------------------------------------------------------------------------
// NVRTC + driver API, loaded lazily
// ---------------------------------------------------------------------------------------------
struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why; int version = 0;
    bool ok() const { return h != nullptr; }
};
inline Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        // Toolkit paths first: a process that imported torch already holds torch's bundled libnvrtc (12.8 here)
        // under the bare soname, and that build miscompiles large chunk functions for sm_100a
        // (profiles/r1_summary.md "NVRTC 12.8"); a full path loads the toolkit's own copy next to it.
        std::vector<std::string> names;
        for (const char* env : {"MDN_NVRTC_PATH"}) if (const char* v = getenv(env)) names.push_back(v);
        for (const char* env : {"CUDA_HOME", "CUDA_PATH"}) if (const char* v = getenv(env)) names.push_back(std::string(v) + "/lib64/libnvrtc.so.12");
        names.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
        names.push_back("libnvrtc.so.12");
        for (const std::string& nm : names) {
            void* h = dlopen(nm.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (!h) continue;
            auto ver = (int (*)(int*, int*))dlsym(h, "nvrtcVersion");
            int maj = 0, mn = 0;
            if (ver) ver(&maj, &mn);
            if (maj > 12 || (maj == 12 && mn >= 9) || getenv("MDN_NVRTC_ALLOW_OLD")) { n.h = h; n.version = maj * 100 + mn; n.why.clear(); break; }
            n.why = "libnvrtc " + std::to_string(maj) + "." + std::to_string(mn) + " is older than 12.9";
            dlclose(h);
        }
        if (!n.h && n.why.empty()) n.why = "libnvrtc not found";
        if (!n.h) return;
        auto sym = [&](const char* s) { void* p = dlsym(n.h, s); if (!p) n.why = std::string("missing symbol ") + s; return p; };
        n.CreateProgram = (decltype(n.CreateProgram))sym("nvrtcCreateProgram");
        n.CompileProgram = (decltype(n.CompileProgram))sym("nvrtcCompileProgram");
        n.GetCUBINSize = (decltype(n.GetCUBINSize))sym("nvrtcGetCUBINSize");
        n.GetCUBIN = (decltype(n.GetCUBIN))sym("nvrtcGetCUBIN");
        n.GetProgramLogSize = (decltype(n.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
        n.GetProgramLog = (decltype(n.GetProgramLog))sym("nvrtcGetProgramLog");
        n.DestroyProgram = (decltype(n.DestroyProgram))sym("nvrtcDestroyProgram");
        n.GetErrorString = (decltype(n.GetErrorString))sym("nvrtcGetErrorString");
        if (!n.why.empty()) { dlclose(n.h); n.h = nullptr; }
    });
    return n;
}

// source -> sm_100a cubin (throws std::runtime_error with the compiler log on failure)
inline std::vector<char> compile(const std::string& src) {
    Nvrtc& n = nvrtc();
    if (!n.ok()) throw std::runtime_error("NVRTC unavailable: " + n.why);
    void* prog = nullptr;
    int rc = n.CreateProgram(&prog, src.c_str(), "mdn_constraints.cu", 0, nullptr, nullptr);
    if (rc) throw std::runtime_error(std::string("nvrtcCreateProgram: ") + n.GetErrorString(rc));
    // chunked functions keep NVRTC + ptxas -O3 linear (10 k nodes: 7 s); MDN_JIT_PTXAS overrides the level
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
    std::string po = std::string("--ptxas-options=") + (getenv("MDN_JIT_PTXAS") ? getenv("MDN_JIT_PTXAS") : "-O3");
    opts.push_back(po.c_str());
    if (jit_arith2()) opts.push_back("--device-int128");
    rc = n.CompileProgram(prog, (int)opts.size(), opts.data());
    if (rc) {
        size_t ls = 0; n.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0'); if (ls) n.GetProgramLog(prog, log.data());
        n.DestroyProgram(&prog);
        throw std::runtime_error(std::string("nvrtcCompileProgram: ") + n.GetErrorString(rc) + "\n" + log.substr(0, 2000));
    }
    size_t sz = 0; n.GetCUBINSize(prog, &sz);
    std::vector<char> cubin(sz);
    n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    return cubin;
}

struct Driver {
    void* h = nullptr;
    int (*ModuleLoadData)(void**, const void*) = nullptr;
    int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
    int (*ModuleUnload)(void*) = nullptr;
    int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
    int (*GetErrorString)(int, const char**) = nullptr;
    std::string why;
    bool ok() const { return h != nullptr; }
};
inline Driver& driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        d.h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!d.h) { d.why = "libcuda.so.1 not found"; return; }
        auto sym = [&](const char* s) { void* p = dlsym(d.h, s); if (!p) d.why = std::string("missing symbol ") + s; return p; };
        d.ModuleLoadData = (decltype(d.ModuleLoadData))sym("cuModuleLoadData");
        d.ModuleGetFunction = (decltype(d.ModuleGetFunction))sym("cuModuleGetFunction");
        d.ModuleUnload = (decltype(d.ModuleUnload))sym("cuModuleUnload");
        d.LaunchKernel = (decltype(d.LaunchKernel))sym("cuLaunchKernel");
        d.GetErrorString = (decltype(d.GetErrorString))sym("cuGetErrorString");
        if (!d.why.empty()) { dlclose(d.h); d.h = nullptr; }
    });
    return d;
}
inline std::string cu_err(int rc) { const char* s = nullptr; if (driver().GetErrorString) driver().GetErrorString(rc, &s); return s ? s : "unknown driver error"; }

struct Kernel {
    void* module = nullptr; void* func = nullptr;
    int checked = 0;   // 0: not yet compared with the interpreter, 1: agreed, -1: disagreed (never used again)
    Kernel() {}
    Kernel(const Kernel&) = delete; Kernel& operator=(const Kernel&) = delete;
    ~Kernel() { if (module && driver().ok()) driver().ModuleUnload(module); }
    // The caller's device must be current and its primary context initialised (any runtime call does that).
    void load(const std::vector<char>& cubin) {
        Driver& d = driver();
        if (!d.ok()) throw std::runtime_error("CUDA driver unavailable: " + d.why);
        int rc = d.ModuleLoadData(&module, cubin.data());
        if (rc) throw std::runtime_error("cuModuleLoadData: " + cu_err(rc));
        rc = d.ModuleGetFunction(&func, module, "k_jit");
        if (rc) throw std::runtime_error("cuModuleGetFunction: " + cu_err(rc));
    }
    template <class Args>
    void launch(const Args& a, unsigned blocks, unsigned threads, cudaStream_t st) const {
        void* params[] = {(void*)&a};
        int rc = driver().LaunchKernel(func, blocks, 1, 1, threads, 1, 1, 0, (void*)st, params, nullptr);
        if (rc) throw std::runtime_error("cuLaunchKernel: " + cu_err(rc));
    }
};

// process-wide cubin cache keyed by a hash of the program words (AIRs are fixed per deployment)
inline uint64_t fnv1a(const uint32_t* w, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= w[i]; h *= 1099511628211ull; }
    return h;
}
inline const std::vector<char>& cubin_for(const uint32_t* w, size_t n_words, GenInfo* info, bool lookup = false, uint32_t n_cols = 0) {
    static std::mutex mu;
    static std::map<uint64_t, std::pair<std::vector<char>, GenInfo>> cache;
    uint64_t key = fnv1a(w, n_words) ^ (uint64_t)n_words << 40 ^ (uint64_t)chunk_nodes() << 20 ^ (uint64_t)n_cols << 8 ^ (uint64_t)lookup;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
        GenInfo gi;
        std::string src = generate(w, &gi, lookup, n_cols);
        if (const char* dump = getenv("MDN_JIT_DUMP")) { if (FILE* f = fopen(dump, "w")) { fwrite(src.data(), 1, src.size(), f); fclose(f); } }
        it = cache.emplace(key, std::make_pair(compile(src), gi)).first;
        if (const char* dump = getenv("MDN_JIT_DUMP_CUBIN")) { if (FILE* f = fopen(dump, "wb")) { fwrite(it->second.first.data(), 1, it->second.first.size(), f); fclose(f); } }
    }
    if (info) *info = it->second.second;
    return it->second.first;
}

}  // namespace jit
