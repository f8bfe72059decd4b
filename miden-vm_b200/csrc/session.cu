// Host orchestration of the proving path + the C ABI (include/miden_b200.h).
//
// This file mirrors `miden_lifted_stark::prover::prove` (reference
// crates/lifted-stark/src/prover/mod.rs:230-578) phase by phase; each phase names the reference
// lines it replaces.  Everything data-parallel runs in the kernels of kernels.cu; the host keeps
// the Fiat-Shamir transcript (sequential) exactly like the reference's host does.
#include "../../include/miden_b200.h"
#include "host_transcript.hpp"
#include "kernels.cuh"
#include "ntt_tables.hpp"
#include "jit.hpp"
#include <cstddef>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

using gl::E2;
using gl::u32;
using gl::u64;
using hostfs::Duplex;
using hostfs::Indices;
using hostfs::Transcript;

namespace {

struct MdnError : std::runtime_error {
    int code;
    MdnError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] void fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    throw MdnError(code, buf);
}
#define CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) fail(MDN_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

std::string g_create_error;

// Proof-lifetime device memory.  Every buffer a proof allocates is gone when the proof ends, and a prover proves
// the same shapes over and over, so the session owns slabs obtained once with cudaMalloc and hands out blocks by
// bumping (plus size-keyed reuse of blocks released mid-proof); at the end of a proof the bump pointers go back
// to zero.  After the first proof of a shape no allocation reaches the driver.  The stream-ordered pool this
// replaces had to re-map physical memory whenever its free space was fragmented when a 1 GB request arrived --
// a driver call that stalled 100-200 ms under a loaded neighbour (profiles/r1_summary.md "step-time outliers").
// Blocks are used on the session's stream only (the copy stream is ordered behind it by events), so handing a
// released block to the next request is safe in stream order.
struct Arena {
    struct Slab { char* base; size_t size, used; };
    std::vector<Slab> slabs;
    std::multimap<size_t, char*> free_blocks;
    size_t live = 0, grow_events = 0;
    // One proof on several GPUs: every rank runs the same allocation sequence, so a block has the same slab index and
    // offset on every rank; the hooks map a new slab into the peers (CUDA IPC) and unmap before slabs are freed.
    std::function<void(char*, size_t)> on_new_slab;
    std::function<void()> before_drop_slabs;
    static size_t round_up(size_t b) { return (b + 511) & ~(size_t)511; }
    void* alloc(size_t bytes) {
        bytes = round_up(bytes);
        auto it = free_blocks.lower_bound(bytes);
        if (it != free_blocks.end() && it->first <= bytes + bytes / 4) { void* p = it->second; free_blocks.erase(it); live++; return p; }
        for (Slab& sl : slabs) if (sl.size - sl.used >= bytes) { void* p = sl.base + sl.used; sl.used += bytes; live++; return p; }
        size_t sz = std::max(bytes, (size_t)1 << 30);
        char* base = nullptr;
        CUDA_OK(cudaMalloc((void**)&base, sz));
        slabs.push_back(Slab{base, sz, bytes});
        grow_events++; live++;
        if (on_new_slab) on_new_slab(base, sz);
        return base;
    }
    void free(void* p, size_t bytes) { free_blocks.emplace(round_up(bytes), (char*)p); live--; }
    // end of a proof: everything must have been released; several slabs are merged into one of the total size
    void reset() {
        if (live) return;
        free_blocks.clear();
        for (Slab& sl : slabs) sl.used = 0;
        if (slabs.size() > 1) {
            size_t total = 0;
            for (Slab& sl : slabs) total += sl.size;
            if (before_drop_slabs) before_drop_slabs();
            for (Slab& sl : slabs) cudaFree(sl.base);
            slabs.clear();
            char* base = nullptr;
            if (cudaMalloc((void**)&base, total) == cudaSuccess) { slabs.push_back(Slab{base, total, 0}); if (on_new_slab) on_new_slab(base, total); }
            else cudaGetLastError();   // the next proof grows again
        }
    }
    void destroy() { if (before_drop_slabs && !slabs.empty()) before_drop_slabs(); for (Slab& sl : slabs) cudaFree(sl.base); slabs.clear(); free_blocks.clear(); live = 0; }
};
thread_local Arena* tl_arena = nullptr;   // set while an API call works on a proof
struct ArenaScope { Arena* prev; explicit ArenaScope(Arena* a) : prev(tl_arena) { tl_arena = a; } ~ArenaScope() { tl_arena = prev; } };

// device buffer: from the proof arena when one is active, else stream-ordered (persistent tables, tools)
struct DevBuf {
    u64* p = nullptr; size_t n = 0; cudaStream_t st = nullptr; Arena* owner = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept { p = o.p; n = o.n; st = o.st; owner = o.owner; o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { release(); p = o.p; n = o.n; st = o.st; owner = o.owner; o.p = nullptr; o.n = 0; return *this; }
    void alloc(size_t count, cudaStream_t s) {
        release(); st = s; n = count; owner = tl_arena;
        if (!count) return;
        if (owner) p = (u64*)owner->alloc(count * sizeof(u64));
        else CUDA_OK(cudaMallocAsync((void**)&p, count * sizeof(u64), s));
    }
    void release() {
        if (!p) return;
        if (owner) owner->free(p, n * sizeof(u64)); else cudaFreeAsync(p, st);
        p = nullptr; n = 0;
    }
    ~DevBuf() { release(); }
};

struct NttPlan {
    mk::NttTables T;
    DevBuf store;
};
struct PremulPlan {
    mk::PremulTables P;
    DevBuf store;
    u32 n_bases;
};

enum ProfCat { PC_TRANSPOSE = 0, PC_NTT, PC_LEAF, PC_COMPRESS, PC_CONSTRAINTS, PC_OOD, PC_DEEP, PC_FRI, PC_GRIND, PC_GATHER, PC_COUNT };
struct Prof {
    std::vector<cudaEvent_t> pool;
    struct Region { int cat; size_t e0, e1; };
    std::vector<Region> regions;
    size_t used = 0;
    cudaStream_t st = nullptr;
    size_t ev() { if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); } return used++; }
    size_t begin(int cat) { size_t a = ev(); cudaEventRecord(pool[a], st); regions.push_back({cat, a, a}); return regions.size() - 1; }
    void end(size_t r) { size_t b = ev(); cudaEventRecord(pool[b], st); regions[r].e1 = b; }
    void reset() { used = 0; regions.clear(); }
    void resolve(float* ms, unsigned* counts) {
        for (int i = 0; i < PC_COUNT; i++) { ms[i] = 0; counts[i] = 0; }
        for (auto& r : regions) { float t = 0; cudaEventElapsedTime(&t, pool[r.e0], pool[r.e1]); ms[r.cat] += t; counts[r.cat]++; }
    }
    ~Prof() { for (auto e : pool) cudaEventDestroy(e); }
};
struct ProfScope { Prof& p; size_t r; ProfScope(Prof& p_, int cat) : p(p_), r(p_.begin(cat)) {} ~ProfScope() { p.end(r); } };

struct Tree {
    DevBuf nodes;   // heap layout: layer d at ((1<<d)-1)*4, 4 u64 per digest
    u32 depth = 0;
    u64* layer(u32 d) { return nodes.p + (((size_t)1 << d) - 1) * 4; }
};

struct CommittedMat { u64* lde; u64* coef; u32 log_n, width; };
struct Committed {
    DevBuf lde_buf, coef_buf;
    std::vector<CommittedMat> mats;   // proof order (ascending height)
    Tree tree;
    u64 root[4];
};

struct AirHost {
    mdn_air desc;
    DevBuf program;   // nodes | constraints | consts(lo,hi pairs as u64)
    mk::AirDev dev;
    // lowered LookupAir (mdn_air.lookup): compiled program, raw periodic matrix, and the raw column-major main
    // trace kept from before the in-place inverse NTT (the LogUp fractions are evaluated on the trace domain)
    bool has_lookup = false;
    DevBuf lookup_program, raw_main;
    mk::AirDev lookup_dev;
    // NVRTC-specialised constraint kernel (jit.hpp) for large programs; NULL = interpreter
    std::shared_ptr<jit::Kernel> jit, lookup_jit;
    u32 n_constraints = 0;
};

}  // namespace

struct mdn_session {
    mdn_pcs_params params;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    // ONE proof on G ranks (mdn_session_set_shard; kernels.cuh "One proof on G GPUs"): rank g computes LDE cosets
    // [g*B/G, (g+1)*B/G) of every column -- forward NTTs, leaf sponge, constraints, DEEP, FRI folds -- and the Merkle
    // sub-tree over leaves [g*L/G, (g+1)*L/G).  Results cross ranks as peer-memory stores ordered by a device-side
    // barrier; the host callback is only the bootstrap transport for the CUDA IPC handles.
    u32 shard_rank = 0, shard_world = 1, shard_log_g = 0;
    mdn_allgather_fn allgather = nullptr; void* allgather_ctx = nullptr;
    bool shard_active = false;                     // inside a proof that is split over the ranks
    u32 shard_min_log = 10;                        // FRI layers below 2^shard_min_log leaves are replicated (MDN_SHARD_MIN_LOG)
    struct SlabView { char* base[mk::MAX_RANKS]; size_t size; };
    std::vector<SlabView> slab_views;              // every rank's mapping of arena slab i
    u64* sync_local = nullptr; mk::PeerPtrs sync_flags{}; u64 sync_epoch = 0;
    bool sharded() const { return shard_active && shard_world > 1; }
    bool tree_sharded(u32 depth) const { return sharded() && depth >= shard_log_g + 6; }
    bool fri_layer_sharded(u32 log_dom) const {    // domain 2^log_dom split by coset: folds and leaves stay rank-local
        const u32 la = params.log_folding_arity, lb = params.log_blowup;
        return sharded() && log_dom >= la + lb && log_dom >= la + shard_min_log;
    }
    u32 nt() const { return sharded() ? (1u << params.log_blowup) >> shard_log_g : (1u << params.log_blowup); }   // cosets of this rank
    u32 t0() const { return sharded() ? shard_rank * nt() : 0; }
    int coset_owner(u32 t) const { return sharded() ? (int)(t / nt()) : -1; }
    // STARK hash configuration (mdn_session_set_hash): Poseidon2 sponge + duplex challenger (default) or Blake3 chaining
    // hasher + hash challenger (air/src/config.rs:276-307).  `align` = LMCS alignment: 8 (sponge rate) or 1 (chaining).
    int hash_kind = MDN_HASH_POSEIDON2;
    u32 align() const { return hash_kind == MDN_HASH_BLAKE3 ? 1u : hash_kind == MDN_HASH_KECCAK ? 17u : 8u; }
    bool byte_hash() const { return hash_kind == MDN_HASH_BLAKE3 || hash_kind == MDN_HASH_KECCAK; }     // hash challenger, not the duplex one
    int perm() const { return byte_hash() ? 0 : hash_kind; }                                             // rescue.cuh permutations share the Poseidon2 kernels
    // Compression layers d_from-1 ... lg of the sub-tree of `rank` (the whole tree: lg = 0, rank = 0): one launch per layer while a
    // layer has more than 512 nodes, then ONE single-block launch for the rest (mk::launch_compress_top).  Returns the permutations.
    size_t compress_subtree(Tree& t, u32 d_from, u32 lg, u32 rank) {
        static constexpr u32 TOP = 10;
        size_t n_perm = 0;
        u32 d = d_from;
        for (; d > lg && d - 1 - lg >= TOP; d--) {
            size_t cnt = (size_t)1 << (d - 1 - lg), start = (size_t)rank << (d - 1 - lg);
            n_perm += cnt;
            compress_layer(t.layer(d) + 2 * start * 4, t.layer(d - 1) + start * 4, cnt);
        }
        if (d > lg) { n_perm += ((size_t)1 << (d - lg)) - 1; mk::launch_compress_top(t.nodes.p, d, lg, rank, hash_kind, stream); }
        return n_perm;
    }
    u32 state_words() const { return hash_kind == MDN_HASH_KECCAK ? 25u : 12u; }      // leaf state handed on between height groups
    std::vector<uint8_t> hash_ch_in, hash_ch_out;          // pre-bound HashChallenger state (mdn_session_set_hash_challenger)
    void hash_leaves(const mk::LeafArgs& a, u32 ln, u32 lb, const u64* prev, u32 prev_log, u64* states_out, const mk::PushDst* dig, u32 tb, u32 tn) {
        if (hash_kind == MDN_HASH_BLAKE3) mk::launch_leaf_hash_b3(a, ln, lb, prev, prev_log, states_out, dig, tb, tn, stream);
        else if (hash_kind == MDN_HASH_KECCAK) mk::launch_leaf_hash_kk(a, ln, lb, prev, prev_log, states_out, dig, tb, tn, stream);
        else mk::launch_leaf_hash(a, ln, lb, prev, prev_log, states_out, dig, tb, tn, stream, perm());
    }
    void compress_layer(const u64* children, u64* parents, size_t n) {
        if (hash_kind == MDN_HASH_BLAKE3) mk::launch_compress_layer_b3(children, parents, n, stream);
        else if (hash_kind == MDN_HASH_KECCAK) mk::launch_compress_layer_kk(children, parents, n, stream);
        else mk::launch_compress_layer(children, parents, n, stream, perm());
    }
    mdn_external_check external_check = nullptr; void* external_ctx = nullptr;   // Statement::eval_external (mdn_session_set_external_check)
    void shard_map_slab(char* base, size_t size);
    void shard_unmap_slabs();
    void shard_teardown();
    void shard_barrier();
    void shard_check(const char* where);
    // the same check folded into a stream synchronisation the caller performs anyway: enqueue the copy of the barrier
    // flag before that synchronisation, evaluate it after (one host round trip less per commitment of a split proof)
    u32 shard_flag_host = 0;
    void shard_check_enqueue();
    void shard_check_finish(const char* where);
    mk::PeerPtrs peers_of(const u64* p) const;
    mk::PushDst push_dst(u64* p, u32 mode, u32 owner_shift = 0) const;
    std::map<u32, std::unique_ptr<NttPlan>> ntt_plans;
    std::map<std::pair<u32, u32>, std::unique_ptr<PremulPlan>> premul_plans;   // (n, kind)

    // ---- per-proof state ----
    Arena arena;
    bool use_arena = getenv("MDN_NO_ARENA") == nullptr;   // off: every buffer is its own allocation (compute-sanitizer memcheck)
    void release_proof_memory();
    bool in_proof = false;
    std::vector<AirHost> airs;              // instance order
    std::vector<u32> log_heights;           // instance order
    std::vector<u32> order;                 // proof position -> instance
    std::vector<u64> publics;
    u32 log_max_n = 0;
    u32 log_qd = 0;                         // max log_quotient_degree over the AIRs (number of quotient chunks)
    Transcript tr;
    std::vector<E2> randomness;
    Committed main_c, aux_c, quot_c;
    // Preprocessed bundle (mdn_session_set_preprocessed): persists across proofs like the reference's borrowed
    // `Preprocessed` (preprocessed.rs:49-61).  prep_air[q] = instance of committed preprocessed trace q.
    // constraint JIT: programs with at least jit_min_nodes nodes are compiled with NVRTC (0 = never)
    u32 jit_min_nodes = 256;
    std::map<u64, std::shared_ptr<jit::Kernel>> jit_kernels;   // loaded modules by program hash
    std::vector<u64> jit_used;       // per AIR of the last proof: 1 = JIT kernel, 0 = interpreter
    std::string jit_note;            // why the JIT was not used, if it was wanted
    Committed prep_c; bool has_prep = false;
    std::vector<u32> prep_air, prep_log_h;
    void set_preprocessed(const mdn_statement* st, const mdn_matrix* mats);
    std::vector<std::vector<u64>> aux_values_p;   // proof order, EF pairs
    DevBuf d_publics, d_randomness, d_aux_values, d_flag;
    void check_input_flag(const char* what);
    std::vector<size_t> aux_values_off;           // proof order offsets (in u64) into d_aux_values
    // outputs
    std::vector<uint8_t> out_heights;
    std::vector<u64> out_fields, out_commitments;
    // introspection
    E2 ood_z{0, 0};
    u64 dbg_roots[3][4] = {};
    std::vector<u64> dbg_quot_acc, dbg_deep, dbg_fri_roots, dbg_queries;
    bool keep_debug = false;
    mdn_timings timings{};
    cudaEvent_t ev[16];
    Prof prof;
    double leaf_bytes = 0, ntt_bytes = 0; unsigned long long perms = 0;

    NttPlan& ntt(u32 n);
    PremulPlan& premul_trace(u32 n);
    PremulPlan& premul_quotient(u32 n, u32 log_d);
    void build_tree(Committed& c);
    void lde_matrix(CommittedMat& m);
    void keep_raw_main(u32 j);
    void build_logup_aux(u32 j, u64* aux_cm, u64 final_out[2]);
    void lde_and_commit(Committed& c, float* t_lde, float* t_hash, bool lde_done = false);
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_ev[8];
    // pinned bounce buffers for pageable host inputs (a Rust Vec<Felt> is pageable)
    u64* bounce[2] = {nullptr, nullptr}; cudaEvent_t bounce_ev[2]; bool bounce_busy[2] = {false, false};
    static constexpr size_t BOUNCE_WORDS = (size_t)4 << 20;   // 32 MiB each
    void host_to_device(u64* dst, const u64* src, size_t n);
    void upload_matrix(const mdn_matrix& m, bool on_device, u64* dst_cm);
    void prove_begin(const mdn_statement* st, const mdn_matrix* traces, const mdn_challenger* ch, u32 flags);
    void commit_aux(const mdn_matrix* aux, const u64* const* aux_values, bool zero_aux);
    void finish();
    u64 grind(u32 bits);
    void reset_proof();
};

namespace {

// ---------------------------------------------------------------------------------------------
// table construction (host, tiny): ntt_tables.hpp ------------------------------------------------
// ---------------------------------------------------------------------------------------------
using ntt_tables::split_n;

}  // namespace

NttPlan& mdn_session::ntt(u32 n) {
    auto it = ntt_plans.find(n);
    if (it != ntt_plans.end()) return *it->second;
    if (n > 22) fail(MDN_ERR_UNSUPPORTED, "trace height 2^%u exceeds the supported 2^22", n);
    auto plan = std::make_unique<NttPlan>();
    ntt_tables::NttHost host = ntt_tables::build_ntt(n);
    { ArenaScope persistent(nullptr); plan->store.alloc(host.data.size(), stream); }
    CUDA_OK(cudaMemcpyAsync(plan->store.p, host.data.data(), host.data.size() * sizeof(u64), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    plan->T = host.view(plan->store.p);
    auto& ref = *plan;
    ntt_plans[n] = std::move(plan);
    return ref;
}

namespace {
std::unique_ptr<PremulPlan> make_premul(const std::vector<u64>& bases, u32 n, cudaStream_t stream) {
    ntt_tables::PremulHost host = ntt_tables::build_premul(bases, n);
    auto plan = std::make_unique<PremulPlan>();
    { ArenaScope persistent(nullptr); plan->store.alloc(host.data.size(), stream); }
    CUDA_OK(cudaMemcpyAsync(plan->store.p, host.data.data(), host.data.size() * sizeof(u64), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    plan->P = host.view(plan->store.p);
    plan->n_bases = (u32)bases.size();
    return plan;
}
}  // namespace

// bases g_t = shift * w_L^t: coset t of the LDE of a height-2^n trace (commit.rs:137-141, domain.rs:358-361)
PremulPlan& mdn_session::premul_trace(u32 n) {
    auto key = std::make_pair(n, 0u);
    auto it = premul_plans.find(key);
    if (it != premul_plans.end()) return *it->second;
    u32 lb = params.log_blowup, B = 1u << lb;
    u64 s = gl::lde_shift(n + lb), wl = gl::two_adic_generator(n + lb);
    std::vector<u64> bases(B);
    u64 x = s;
    for (u32 t = 0; t < B; t++) { bases[t] = x; x = gl::mul(x, wl); }
    auto plan = make_premul(bases, n, stream);
    auto& ref = *plan;
    premul_plans[key] = std::move(plan);
    return ref;
}
// bases w_J^(-t) * w_L^(t'), id = t*B + t' for chunk t < D and LDE coset t' < B, with J the quotient
// domain of order N*D (w_J = w_L^(B/D))  (quotient.rs:186-209: chunk t scaled by w_J^(-kt), then a
// plain DFT over K evaluates on g*K because the iDFT over H left g^k baked into the coefficients)
PremulPlan& mdn_session::premul_quotient(u32 n, u32 log_d) {
    auto key = std::make_pair(n, 1u + log_d);
    auto it = premul_plans.find(key);
    if (it != premul_plans.end()) return *it->second;
    u32 lb = params.log_blowup, B = 1u << lb, D = 1u << log_d;
    u64 wl = gl::two_adic_generator(n + lb), wji = gl::inv(gl::two_adic_generator(n + log_d));
    std::vector<u64> bases(D * B);
    for (u32 t = 0; t < D; t++)
        for (u32 t2 = 0; t2 < B; t2++) bases[t * B + t2] = gl::mul(gl::pow(wji, t), gl::pow(wl, t2));
    auto plan = make_premul(bases, n, stream);
    auto& ref = *plan;
    premul_plans[key] = std::move(plan);
    return ref;
}

void mdn_session::reset_proof() {
    in_proof = false; shard_active = false;
    log_heights.clear(); order.clear(); publics.clear(); randomness.clear();
    aux_values_p.clear(); aux_values_off.clear();
    tr = Transcript();
    release_proof_memory();
}

// Called when no proof-lifetime buffer of a caller's frame is alive any more (API wrappers, error paths).
void mdn_session::release_proof_memory() {
    main_c = Committed(); aux_c = Committed(); quot_c = Committed();
    airs.clear();
    d_publics.release(); d_randomness.release(); d_aux_values.release();
    arena.reset();
}

// ---------------------------------------------------------------------------------------------
// One proof on G ranks: peer mappings of the arena, device barrier
// ---------------------------------------------------------------------------------------------
// Collective (every rank reaches it at the same point of the same allocation sequence): export the new slab, gather
// the G handles through the host callback, map the peers' slabs.
void mdn_session::shard_map_slab(char* base, size_t size) {
    if (shard_world <= 1) return;
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, base));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    const size_t W = 10;
    std::vector<u64> mine(W), all(W * shard_world);
    memcpy(mine.data(), &h, 64); mine[8] = size; mine[9] = slab_views.size();
    if (allgather(allgather_ctx, mine.data(), all.data(), W) != 0) fail(MDN_ERR_INVALID_ARG, "all-gather callback failed");
    SlabView v{}; v.size = size;
    for (u32 g = 0; g < shard_world; g++) {
        if (all[g * W + 8] != size || all[g * W + 9] != mine[9])
            fail(MDN_ERR_INVALID_ARG, "rank %u allocated a different proof arena (slab %llu of %llu bytes here, slab %llu of %llu bytes there): every rank must prove the same statement with the same flags",
                 g, (unsigned long long)mine[9], (unsigned long long)size, (unsigned long long)all[g * W + 9], (unsigned long long)all[g * W + 8]);
        if (g == shard_rank) { v.base[g] = base; continue; }
        cudaIpcMemHandle_t hg; memcpy(&hg, &all[g * W], 64);
        void* m = nullptr;
        CUDA_OK(cudaIpcOpenMemHandle(&m, hg, cudaIpcMemLazyEnablePeerAccess));
        v.base[g] = (char*)m;
    }
    slab_views.push_back(v);
    { std::vector<u64> one(1, 0), got(shard_world); if (allgather(allgather_ctx, one.data(), got.data(), 1) != 0) fail(MDN_ERR_INVALID_ARG, "all-gather callback failed"); }   // every rank has mapped it
}
// Collective: nobody frees a slab a peer still has mapped.
void mdn_session::shard_unmap_slabs() {
    if (slab_views.empty()) return;
    cudaStreamSynchronize(stream);
    for (auto& v : slab_views) for (u32 g = 0; g < shard_world; g++) if (g != shard_rank && v.base[g]) cudaIpcCloseMemHandle(v.base[g]);
    slab_views.clear();
    if (shard_world > 1 && allgather) { std::vector<u64> one(1, 0), all(shard_world); allgather(allgather_ctx, one.data(), all.data(), 1); }
}
void mdn_session::shard_teardown() {
    if (shard_world > 1) {
        release_proof_memory();
        arena.destroy();                    // unmaps the peers first (before_drop_slabs)
        if (sync_local) {
            for (u32 g = 0; g < shard_world; g++) if (g != shard_rank && sync_flags.p[g]) cudaIpcCloseMemHandle(sync_flags.p[g]);
            cudaFree(sync_local); sync_local = nullptr;
        }
    }
    arena.on_new_slab = nullptr; arena.before_drop_slabs = nullptr;
    sync_flags = mk::PeerPtrs{}; sync_epoch = 0;
    shard_rank = 0; shard_world = 1; shard_log_g = 0; allgather = nullptr; allgather_ctx = nullptr; shard_active = false;
}
mk::PeerPtrs mdn_session::peers_of(const u64* p) const {
    mk::PeerPtrs r{};
    if (shard_world <= 1) { r.p[0] = const_cast<u64*>(p); return r; }
    for (size_t i = 0; i < slab_views.size() && i < arena.slabs.size(); i++) {
        const char* b = arena.slabs[i].base;
        if ((const char*)p >= b && (const char*)p < b + arena.slabs[i].size) {
            size_t off = (const char*)p - b;
            for (u32 g = 0; g < shard_world; g++) r.p[g] = (u64*)(slab_views[i].base[g] + off);
            return r;
        }
    }
    fail(MDN_ERR_CUDA, "internal: buffer outside the shared proof arena");
}
mk::PushDst mdn_session::push_dst(u64* p, u32 mode, u32 owner_shift) const {
    if (!sharded() || mode == mk::PUSH_LOCAL) return mk::local_dst(p);
    mk::PushDst d{};
    d.pp = peers_of(p); d.rank = shard_rank; d.world = shard_world; d.mode = mode; d.owner_shift = owner_shift;
    return d;
}
void mdn_session::shard_barrier() {
    if (!sharded()) return;
    mk::launch_barrier(sync_flags, shard_rank, shard_world, ++sync_epoch, (u32*)d_flag.p, stream);
}
void mdn_session::shard_check(const char* where) {
    if (!sharded()) return;
    u32 flag = 0;
    CUDA_OK(cudaMemcpyAsync(&flag, d_flag.p, sizeof flag, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    if (flag & 8) { CUDA_OK(cudaMemsetAsync(d_flag.p, 0, 8, stream)); fail(MDN_ERR_CUDA, "cross-GPU barrier timed out (%s): a peer rank stopped or proves a different statement", where); }
}

void mdn_session::shard_check_enqueue() {
    shard_flag_host = 0;
    if (sharded()) CUDA_OK(cudaMemcpyAsync(&shard_flag_host, d_flag.p, sizeof shard_flag_host, cudaMemcpyDeviceToHost, stream));
}
void mdn_session::shard_check_finish(const char* where) {
    if (!sharded()) return;
    if (shard_flag_host & 8) { CUDA_OK(cudaMemsetAsync(d_flag.p, 0, 8, stream)); fail(MDN_ERR_CUDA, "cross-GPU barrier timed out (%s): a peer rank stopped or proves a different statement", where); }
}

void mdn_session::check_input_flag(const char* what) {
    u32 flag = 0;
    CUDA_OK(cudaMemcpyAsync(&flag, d_flag.p, sizeof flag, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    if (sharded()) {   // slices checked by the other ranks report into every rank's word (launch_transpose_slice_push)
        u32 peer_flag = 0;
        CUDA_OK(cudaMemcpyAsync(&peer_flag, sync_local + 16, sizeof peer_flag, cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        if (peer_flag) { CUDA_OK(cudaMemsetAsync(sync_local + 16, 0, 8, stream)); flag |= peer_flag & 1u; }
    }
    if (flag) {
        CUDA_OK(cudaMemsetAsync(d_flag.p, 0, 8, stream));
        if (flag & 8) fail(MDN_ERR_CUDA, "cross-GPU barrier timed out: a peer rank stopped or proves a different statement");
        fail(MDN_ERR_INVALID_ARG, "%s contains a non-canonical field element (>= p)", what);
    }
}

// Host -> device copy on the copy stream.  Pinned (or registered) memory goes straight to the DMA engine;
// pageable memory is staged through two pinned 32 MiB bounce buffers filled by a few host threads, which
// sustains ~3x the throughput of a plain cudaMemcpyAsync from pageable memory (profiles/r1_summary.md).
void mdn_session::host_to_device(u64* dst, const u64* src, size_t n) {
    cudaPointerAttributes at{};
    bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) { CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(u64), cudaMemcpyHostToDevice, copy_stream)); return; }
    for (int b = 0; b < 2; b++) if (!bounce[b]) { CUDA_OK(cudaHostAlloc((void**)&bounce[b], BOUNCE_WORDS * sizeof(u64), cudaHostAllocDefault)); CUDA_OK(cudaEventCreateWithFlags(&bounce_ev[b], cudaEventDisableTiming)); }
    unsigned nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency() / 2));
    int b = 0;
    for (size_t off = 0; off < n; off += BOUNCE_WORDS, b ^= 1) {
        size_t cnt = std::min(BOUNCE_WORDS, n - off);
        if (bounce_busy[b]) CUDA_OK(cudaEventSynchronize(bounce_ev[b]));
        std::vector<std::thread> th;
        size_t per = (cnt + nt - 1) / nt;
        for (unsigned q = 0; q < nt; q++) {
            size_t a = (size_t)q * per, e = std::min(cnt, a + per);
            if (a >= e) break;
            th.emplace_back([=] { memcpy(bounce[b] + a, src + off + a, (e - a) * sizeof(u64)); });
        }
        for (auto& t : th) t.join();
        CUDA_OK(cudaMemcpyAsync(dst + off, bounce[b], cnt * sizeof(u64), cudaMemcpyHostToDevice, copy_stream));
        CUDA_OK(cudaEventRecord(bounce_ev[b], copy_stream));
        bounce_busy[b] = true;
    }
}

// row-major (host or device) -> column-major device
void mdn_session::upload_matrix(const mdn_matrix& m, bool on_device, u64* dst_cm) {
    size_t N = (size_t)1 << m.log_height;
    if (m.width == 0) return;
    if (on_device) {
        ProfScope ps(prof, PC_TRANSPOSE);
        mk::launch_transpose_rm_to_cm(m.values, dst_cm, (u32)N, m.width, (u32*)d_flag.p, stream);
        return;
    }
    DevBuf staging; staging.alloc(N * m.width, stream);
    CUDA_OK(cudaEventRecord(copy_ev[7], stream));
    CUDA_OK(cudaStreamWaitEvent(copy_stream, copy_ev[7], 0));
    host_to_device(staging.p, m.values, N * m.width);
    CUDA_OK(cudaEventRecord(copy_ev[6], copy_stream));
    CUDA_OK(cudaStreamWaitEvent(stream, copy_ev[6], 0));
    ProfScope ps(prof, PC_TRANSPOSE);
    mk::launch_transpose_rm_to_cm(staging.p, dst_cm, (u32)N, m.width, (u32*)d_flag.p, stream);
}

// a1 + a2 + a3: coset LDE of every matrix of `c` (coefficients already in c.coef as natural
// evaluations over H), leaf hashing and tree compression.
//   reference: commit_traces (prover/commit.rs:142-180) -> coset_lde_batch (:173) +
//   build_aligned_tree (:178; lmcs/lifted_tree.rs:202-284)
void mdn_session::lde_matrix(CommittedMat& m) {
    if (!m.width) return;
    u32 lb = params.log_blowup;
    size_t N = (size_t)1 << m.log_n, L = N << lb;
    NttPlan& plan = ntt(m.log_n);
    PremulPlan& pm = premul_trace(m.log_n);
    ProfScope ps(prof, PC_NTT);
    ntt_bytes += (double)(N + (size_t)nt() * N) * m.width * 8.0;   // read the trace column once, write this rank's cosets of the LDE once
    if (sharded() && m.log_n >= shard_log_g + shard_min_log) {
        // interpolation split by column: rank g transforms columns [g*w/G, (g+1)*w/G) and stores the coefficients into
        // every rank (the all-gather of coefficient columns of SURVEY 8(e), as peer stores)
        u32 c0 = (u32)((u64)m.width * shard_rank / shard_world), c1 = (u32)((u64)m.width * (shard_rank + 1) / shard_world);
        if (c1 > c0) mk::launch_intt(m.coef + (size_t)c0 * N, N, c1 - c0, plan.T, stream);
        shard_barrier();     // every rank is done with the raw columns (transposes, the copy kept for the LogUp build)
        if (c1 > c0) mk::launch_push(m.coef + (size_t)c0 * N, peers_of(m.coef + (size_t)c0 * N), shard_rank, shard_world, (size_t)(c1 - c0) * N, stream);
        shard_barrier();
    } else mk::launch_intt(m.coef, N, m.width, plan.T, stream);
    // this rank's cosets only (all B of them on one GPU); column groups sized so a group's LDE (the fwd passes'
    // working set) stays L2-resident
    const u32 tb = t0(), tn = nt();
    size_t col_bytes = (size_t)tn * N * sizeof(u64);
    u32 group = (u32)std::max<size_t>(1, (48u << 20) / col_bytes);
    std::vector<mk::FwdItem> items;
    for (u32 cc = 0; cc < m.width; cc++)
        for (u32 t = tb; t < tb + tn; t++)
            items.push_back(mk::FwdItem{m.coef + (size_t)cc * N, m.lde + (size_t)cc * L + (size_t)t * N, t, 0});
    DevBuf d_items;
    d_items.alloc(items.size() * sizeof(mk::FwdItem) / sizeof(u64), stream);
    CUDA_OK(cudaMemcpyAsync(d_items.p, items.data(), items.size() * sizeof(mk::FwdItem), cudaMemcpyHostToDevice, stream));
    for (u32 c0 = 0; c0 < m.width; c0 += group) {
        u32 cn = std::min(group, m.width - c0);
        mk::launch_fwd_ntt((const mk::FwdItem*)d_items.p + (size_t)c0 * tn, cn * tn, plan.T, pm.P, stream);
    }
}

void mdn_session::lde_and_commit(Committed& c, float* t_lde, float* t_hash, bool lde_done) {
    cudaEvent_t e0 = ev[12], e1 = ev[13], e2 = ev[14];
    CUDA_OK(cudaEventRecord(e0, stream));
    if (!lde_done) for (auto& m : c.mats) lde_matrix(m);
    CUDA_OK(cudaEventRecord(e1, stream));
    build_tree(c);
    CUDA_OK(cudaEventRecord(e2, stream));
    CUDA_OK(cudaEventSynchronize(e2));
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
    if (t_lde) *t_lde += a;
    if (t_hash) *t_hash += b;
}

// leaf sponge states per height group (ascending), then the compression layers.
// Split over ranks: every rank hashes the leaves of its cosets and stores each digest into the rank that owns the
// leaf's sub-tree (leaf index = domain index r*B + t, owner = index >> (depth - log G)); after the barrier every rank
// compresses its own sub-tree, stores its sub-root into all ranks, and the top log G layers are recomputed everywhere.
// Trees too small to split are replicated: the digests go to every rank.
void mdn_session::build_tree(Committed& c) {
    u32 lb = params.log_blowup;
    u32 log_n_max = 0;
    for (auto& m : c.mats) log_n_max = std::max(log_n_max, m.log_n);
    u32 depth = log_n_max + lb;
    size_t L = (size_t)1 << depth;
    c.tree.depth = depth;
    c.tree.nodes.alloc((2 * L - 1) * 4, stream);
    const bool sh = sharded(), split = tree_sharded(depth);
    const u32 lg = shard_log_g, tb = t0(), tn = nt();
    mk::PushDst dig = push_dst(c.tree.layer(depth), sh ? (split ? mk::PUSH_OWNER : mk::PUSH_ALL) : mk::PUSH_LOCAL, depth - (split ? lg : 0));
    shard_barrier();   // the tree buffer is a fresh allocation: no rank may still be using the memory under its old identity
    DevBuf states_a, states_b;
    // One launch absorbs up to 8 matrices of one height (launch-argument space); a taller pile of equal-height matrices
    // is absorbed in several launches that hand the sponge states on through the SoA state buffers, exactly like the
    // hand-over between height groups.
    const u64* prev = nullptr; u32 prev_log = 0;
    size_t i = 0;
    while (i < c.mats.size()) {
        size_t j = i;
        mk::LeafArgs args; args.n_mats = 0;
        const u32 ln = c.mats[i].log_n;
        while (j < c.mats.size() && c.mats[j].log_n == ln && args.n_mats < 8) {
            // a zero-width matrix is a no-op for the sponge (an empty absorb leaves the state untouched) but not for the
            // chaining hasher, which re-hashes its state (crates/stateful-hasher/src/chaining.rs:43-46)
            if (c.mats[j].width || hash_kind == MDN_HASH_BLAKE3) args.m[args.n_mats++] = mk::LeafMat{c.mats[j].lde, c.mats[j].width, 0};
            j++;
        }
        bool last = (j == c.mats.size());
        DevBuf& out = (prev == states_a.p && prev) ? states_b : states_a;
        if (!last) out.alloc((size_t)state_words() << (ln + lb), stream);
        {
            ProfScope ps(prof, PC_LEAF);
            size_t Lg = (size_t)tn << ln;
            const u32 rate = hash_kind == MDN_HASH_KECCAK ? 17 : 8;
            for (int q = 0; q < args.n_mats; q++) { leaf_bytes += (double)Lg * args.m[q].width * 8.0; perms += Lg * ((args.m[q].width + rate - 1) / rate); }
            leaf_bytes += last ? (double)Lg * 32.0 : (double)Lg * 8.0 * state_words();
            hash_leaves(args, ln, lb, prev, prev_log, last ? nullptr : out.p, last ? &dig : nullptr, tb, tn);
        }
        prev = out.p; prev_log = ln;
        i = j;
    }
    shard_barrier();   // every digest of this rank's leaf range (or of the whole replicated tree) has arrived
    if (!split) {
        ProfScope ps(prof, PC_COMPRESS);
        perms += compress_subtree(c.tree, depth, 0, 0);
    } else {
        {
            ProfScope ps(prof, PC_COMPRESS);
            perms += compress_subtree(c.tree, depth, lg, shard_rank);
        }
        u64* mine = c.tree.layer(lg) + (size_t)shard_rank * 4;
        mk::launch_push(mine, peers_of(mine), shard_rank, shard_world, 4, stream);
        shard_barrier();
        ProfScope ps(prof, PC_COMPRESS);
        compress_subtree(c.tree, lg, 0, 0);
    }
    CUDA_OK(cudaMemcpyAsync(c.root, c.tree.layer(0), 4 * sizeof(u64), cudaMemcpyDeviceToHost, stream));
    shard_check_enqueue();
    CUDA_OK(cudaStreamSynchronize(stream));
    shard_check_finish("commitment");
}

// GrindingChallenger::grind on the device: smallest witness (sequential p3 order), then the
// challenger advances to the post-check state (random_coin.masm:929-966).
u64 mdn_session::grind(u32 bits) {
    if (bits == 0) { tr.fields.push_back(0); return 0; }
    Duplex& ch = tr.ch;
    if (ch.hashed) {
        // hash challenger: a candidate w is checked by hashing (input buffer || w as 8 little-endian bytes) and reading
        // the low bits of the first sampled u64; the buffer is whole 32-bit words (every observation is 8 or 32 bytes)
        if (ch.bin.size() % (ch.keccak ? 8 : 4)) fail(MDN_ERR_INVALID_ARG, "hash challenger input buffer is not a whole number of words");
        u32 nw = (u32)(ch.bin.size() / 4);
        DevBuf d; d.alloc((nw + 1) / 2 + 2, stream);
        u64 none = ~0ull;
        if (nw) CUDA_OK(cudaMemcpyAsync(d.p + 1, ch.bin.data(), ch.bin.size(), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(d.p, &none, sizeof none, cudaMemcpyHostToDevice, stream));
        u64 start = 0, found = ~0ull;
        u64 batch = std::max<u64>(1ull << 14, std::min<u64>(1ull << 20, 4ull << bits));
        ProfScope ps(prof, PC_GRIND);
        while (found == ~0ull) {
            if (start >= gl::P) fail(MDN_ERR_INVALID_ARG, "proof-of-work search exhausted");
            if (ch.keccak) mk::launch_grind_kk(d.p + 1, nw / 2, bits, start, batch, d.p, stream);
            else mk::launch_grind_b3((const u32*)(d.p + 1), nw, bits, start, batch, d.p, stream);
            CUDA_OK(cudaMemcpyAsync(&found, d.p, sizeof(u64), cudaMemcpyDeviceToHost, stream));
            CUDA_OK(cudaStreamSynchronize(stream));
            start += batch;
        }
        ch.observe(found);
        if (ch.sample_bits(bits) != 0) fail(MDN_ERR_CUDA, "device proof-of-work witness failed the host check");
        tr.fields.push_back(found);
        return found;
    }
    if (ch.in_len >= 8) fail(MDN_ERR_INVALID_ARG, "challenger buffer overflow");
    u64 base[12];
    for (int i = 0; i < 12; i++) base[i] = ch.st[i];
    for (u32 i = 0; i < ch.in_len; i++) base[i] = ch.in[i];
    DevBuf d; d.alloc(13, stream);
    u64 init[13];
    memcpy(init, base, sizeof base); init[12] = ~0ull;
    CUDA_OK(cudaMemcpyAsync(d.p, init, sizeof init, cudaMemcpyHostToDevice, stream));
    u64 start = 0, found = ~0ull;
    u64 batch = std::max<u64>(1ull << 14, std::min<u64>(1ull << 22, 4ull << bits));
    ProfScope ps(prof, PC_GRIND);
    while (found == ~0ull) {
        if (start >= gl::P) fail(MDN_ERR_INVALID_ARG, "proof-of-work search exhausted");
        mk::launch_grind(d.p, ch.in_len, bits, start, batch, d.p + 12, stream, perm());
        CUDA_OK(cudaMemcpyAsync(&found, d.p + 12, sizeof(u64), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        start += batch;
    }
    ch.observe(found);
    u64 chk = ch.sample_bits(bits);
    if (chk != 0) fail(MDN_ERR_CUDA, "device proof-of-work witness failed the host check");
    tr.fields.push_back(found);
    return found;
}

// ---------------------------------------------------------------------------------------------
// op-list compiler: validation, liveness, slot assignment (shared by constraint and lookup programs)
// ---------------------------------------------------------------------------------------------
struct Compiled { std::vector<u32> code; std::vector<u64> consts; u32 n_slots = 0; bool uses_sel = false; };
// lookup == false: words = [MAIR, 1, n_nodes, n_constraints, n_consts | nodes | constraint node ids | consts];
//                  a FOLD instruction (15) is placed right after the node it folds, in emission order.
// lookup == true : words = [MLKP, 1, n_nodes, n_interactions, n_consts | nodes | {column, flag, mult, denom} | consts];
//                  an EMIT instruction (17) is placed once its three operands are defined.
static Compiled compile_oplist(u32 i, const mdn_air& a, u32 n_publics, const u32* w, u32 n_words, bool lookup, u32 lookup_cols) {
    const char* what = lookup ? "lookup program" : "constraint program";
    const u32 magic = lookup ? 0x504B4C4Du : 0x5249414Du, item = lookup ? 4u : 1u;
    if (!w || n_words < 5 || w[0] != magic || w[1] != 1) fail(MDN_ERR_INVALID_ARG, "AIR %u: bad %s header", i, what);
    u32 nn = w[2], nc = w[3], nk = w[4];
    if ((size_t)n_words != 5 + 3 * (size_t)nn + (size_t)item * nc + 2 * (size_t)nk) fail(MDN_ERR_INVALID_ARG, "AIR %u: bad %s length", i, what);
    Compiled out;
    const u32* items = w + 5 + 3 * (size_t)nn;
    for (u32 j = 0; j < nn; j++) {
        u32 op = w[5 + 3 * j], x = w[6 + 3 * j], y = w[7 + 3 * j];
        if (op > 15) fail(MDN_ERR_INVALID_ARG, "AIR %u: unknown op %u", i, op);
        // a LookupBuilder exposes the main window, periodic values, public values, challenges and constants only
        if (lookup && (op == 1 || (op >= 4 && op <= 7) || op == 15)) fail(MDN_ERR_INVALID_ARG, "AIR %u: op %u is not available to a lookup program", i, op);
        if (op == 15 && (x > 1 || y >= a.preprocessed_width)) fail(MDN_ERR_INVALID_ARG, "AIR %u: preprocessed column out of range", i);
        if (op >= 5 && op <= 7) out.uses_sel = true;
        if (op >= 10 && op <= 12 && (x >= j || y >= j)) fail(MDN_ERR_INVALID_ARG, "AIR %u: forward reference", i);
        if (op == 13 && x >= j) fail(MDN_ERR_INVALID_ARG, "AIR %u: forward reference", i);
        if (op == 0 && (x > 1 || y >= a.width)) fail(MDN_ERR_INVALID_ARG, "AIR %u: main column out of range", i);
        if (op == 1 && (x > 1 || y >= a.aux_width)) fail(MDN_ERR_INVALID_ARG, "AIR %u: aux column out of range", i);
        if (op == 2 && x >= n_publics) fail(MDN_ERR_INVALID_ARG, "AIR %u: public value out of range", i);
        if (op == 3 && x >= a.num_randomness) fail(MDN_ERR_INVALID_ARG, "AIR %u: challenge out of range", i);
        if (op == 4 && x >= a.num_aux_values) fail(MDN_ERR_INVALID_ARG, "AIR %u: aux value out of range", i);
        if (op == 8 && x >= nk) fail(MDN_ERR_INVALID_ARG, "AIR %u: constant out of range", i);
        if (op == 9 && x + 1 >= nk) fail(MDN_ERR_INVALID_ARG, "AIR %u: constant out of range", i);
        if (op == 14 && x >= a.num_periodic_columns) fail(MDN_ERR_INVALID_ARG, "AIR %u: periodic column out of range", i);
    }
    // event stream: node definitions, each followed by the folds / emits that became ready
    struct Ev { u32 kind; u32 node; };   // kind 0: define node; 1: fold node / emit interaction `node`
    std::vector<Ev> evs; evs.reserve(nn + nc);
    auto ready_at = [&](u32 kk) -> u32 {
        if (!lookup) return items[kk];
        const u32* it = items + 4 * (size_t)kk;
        u32 m = std::max(it[2], it[3]);
        return it[1] == 0xFFFFFFFFu ? m : std::max(m, it[1]);
    };
    for (u32 kk = 0; kk < nc; kk++) {
        if (!lookup) { if (items[kk] >= nn) fail(MDN_ERR_INVALID_ARG, "AIR %u: bad constraint id", i); continue; }
        const u32* it = items + 4 * (size_t)kk;
        if (it[0] >= lookup_cols) fail(MDN_ERR_INVALID_ARG, "AIR %u: lookup column out of range", i);
        if ((it[1] != 0xFFFFFFFFu && it[1] >= nn) || it[2] >= nn || it[3] >= nn) fail(MDN_ERR_INVALID_ARG, "AIR %u: bad lookup interaction", i);
    }
    { u32 kk = 0;
      for (u32 j = 0; j < nn; j++) {
          evs.push_back({0, j});
          while (kk < nc && ready_at(kk) <= j) { evs.push_back({1, lookup ? kk : items[kk]}); kk++; }
      }
      if (kk != nc) fail(MDN_ERR_INVALID_ARG, "AIR %u: %s with no nodes", i, what); }
    std::vector<u32> last_use(nn, 0);
    std::vector<uint8_t> is_ext(nn, 0);
    for (u32 e = 0; e < evs.size(); e++) {
        if (evs[e].kind) {
            if (!lookup) { last_use[evs[e].node] = e; continue; }
            const u32* it = items + 4 * (size_t)evs[e].node;
            if (it[1] != 0xFFFFFFFFu) last_use[it[1]] = e;
            last_use[it[2]] = e; last_use[it[3]] = e;
            continue;
        }
        u32 j = evs[e].node, op = w[5 + 3 * j], x = w[6 + 3 * j], y = w[7 + 3 * j];
        last_use[j] = std::max(last_use[j], e);
        if (op >= 10 && op <= 12) { last_use[x] = e; last_use[y] = e; is_ext[j] = is_ext[x] | is_ext[y]; }
        else if (op == 13) { last_use[x] = e; is_ext[j] = is_ext[x]; }
        else is_ext[j] = (op == 1 || op == 3 || op == 4 || op == 9);
    }
    if (lookup) for (u32 kk = 0; kk < nc; kk++) {
        const u32* it = items + 4 * (size_t)kk;
        if ((it[1] != 0xFFFFFFFFu && is_ext[it[1]]) || is_ext[it[2]]) fail(MDN_ERR_INVALID_ARG, "AIR %u: lookup flag and multiplicity must be base-field expressions", i);
    }
    std::vector<u32> slot_of(nn, 0), free_slots;
    u32 n_slots = 0;
    std::vector<u32>& code = out.code; code.reserve(4 * evs.size());
    auto release = [&](u32 node, u32 e) { if (last_use[node] == e) free_slots.push_back(slot_of[node]); };
    for (u32 e = 0; e < evs.size(); e++) {
        if (evs[e].kind && !lookup) {
            code.insert(code.end(), {15u | ((u32)is_ext[evs[e].node] << 8), 0u, slot_of[evs[e].node], 0u});
            release(evs[e].node, e);
            continue;
        }
        if (evs[e].kind) {
            const u32* it = items + 4 * (size_t)evs[e].node;
            u32 fs = it[1] == 0xFFFFFFFFu ? 0xffffu : slot_of[it[1]];
            code.insert(code.end(), {17u, it[0], fs | (slot_of[it[2]] << 16), slot_of[it[3]]});
            // a node may serve several operands of one interaction: release each distinct node once
            u32 ops3[3] = {it[1], it[2], it[3]};
            for (int q = 0; q < 3; q++) {
                if (ops3[q] == 0xFFFFFFFFu) continue;
                bool dup = false;
                for (int q2 = 0; q2 < q; q2++) dup |= ops3[q2] == ops3[q];
                if (!dup) release(ops3[q], e);
            }
            continue;
        }
        u32 j = evs[e].node, op = w[5 + 3 * j], x = w[6 + 3 * j], y = w[7 + 3 * j];
        u32 ox = x, oy = y;
        if (op >= 10 && op <= 13) {
            ox = slot_of[x]; oy = (op == 13) ? 0 : slot_of[y];
            release(x, e);
            if (op != 13 && y != x) release(y, e);
        }
        u32 dst;
        if (!free_slots.empty()) { dst = free_slots.back(); free_slots.pop_back(); }
        else dst = n_slots++;
        slot_of[j] = dst;
        code.insert(code.end(), {(op == 15 ? 16u : op) | ((u32)is_ext[j] << 8), dst, ox, oy});   // compiled 15 = FOLD, 16 = PREPROCESSED
        if (last_use[j] == e) free_slots.push_back(dst);   // dead value
    }
    if (n_slots > 1024) fail(MDN_ERR_UNSUPPORTED, "AIR %u: %s needs %u live values (interpreter limit 1024)", i, what, n_slots);
    out.n_slots = n_slots;
    const u32* kw = items + (size_t)item * nc;
    for (u32 j = 0; j < nk; j++) {
        u64 v = (u64)kw[2 * j] | ((u64)kw[2 * j + 1] << 32);
        if (v >= gl::P) fail(MDN_ERR_INVALID_ARG, "AIR %u: non-canonical constant", i);
        out.consts.push_back(v);
    }
    return out;
}


// ---------------------------------------------------------------------------------------------
// prove_begin: validation, statement/shape binding, main commit, randomness  (mod.rs:240-349)
// ---------------------------------------------------------------------------------------------
void mdn_session::prove_begin(const mdn_statement* st, const mdn_matrix* traces, const mdn_challenger* chal, u32 flags) {
    reset_proof();
    log_qd = 0;
    memset(&timings, 0, sizeof timings);
    mk::reset_launch_count();
    prof.st = stream; prof.reset(); leaf_bytes = ntt_bytes = 0; perms = 0;
    if (!d_flag.p) { ArenaScope persistent(nullptr); d_flag.alloc(1, stream); CUDA_OK(cudaMemsetAsync(d_flag.p, 0, 8, stream)); }
    if (!st || !traces || (!chal && !byte_hash())) fail(MDN_ERR_INVALID_ARG, "null argument");
    if (st->n_airs == 0 || st->n_airs > 256) fail(MDN_ERR_INVALID_ARG, "AIR count must be in 1..=256");
    if (params.log_folding_arity < 1 || params.log_folding_arity > 3) fail(MDN_ERR_INVALID_ARG, "invalid folding arity: log_arity %u (must be 1, 2, or 3)", params.log_folding_arity);
    u32 lb = params.log_blowup;
    if (lb == 0 || lb > 4) fail(MDN_ERR_UNSUPPORTED, "log_blowup must be in 1..=4");
    if (st->n_public_values && !st->public_values) fail(MDN_ERR_INVALID_ARG, "public_values is NULL");
    if (st->n_observe_felts && !st->observe_felts) fail(MDN_ERR_INVALID_ARG, "observe_felts is NULL");
    for (u32 i = 0; i < st->n_public_values; i++) if (st->public_values[i] >= gl::P) fail(MDN_ERR_INVALID_ARG, "public value %u is not a canonical field element (>= p)", i);
    for (u32 i = 0; i < st->n_observe_felts; i++) if (st->observe_felts[i] >= gl::P) fail(MDN_ERR_INVALID_ARG, "observed statement felt %u is not a canonical field element (>= p)", i);
    bool on_device = (flags & MDN_FLAG_DEVICE_TRACES) != 0;
    if (shard_world > 1) {
        if (!use_arena) fail(MDN_ERR_UNSUPPORTED, "a proof split over several ranks needs the proof arena (unset MDN_NO_ARENA)");
        if (lb < shard_log_g) fail(MDN_ERR_UNSUPPORTED, "a proof can be split over at most 2^log_blowup = %u ranks (one LDE coset each)", 1u << lb);
        shard_active = true;
        shard_barrier();   // no rank stores into a peer's arena for this proof before every rank has left the previous one
    }
    CUDA_OK(cudaEventRecord(ev[0], stream));

    u32 k = st->n_airs;
    airs.resize(k); log_heights.resize(k);
    for (u32 i = 0; i < k; i++) {
        const mdn_air& a = st->airs[i];
        if (traces[i].width != a.width) fail(MDN_ERR_INVALID_ARG, "trace %u width %u does not match AIR width %u", i, traces[i].width, a.width);
        if (a.log_quotient_degree > lb)
            fail(MDN_ERR_DOMAIN, "log_quotient_degree %u > log_blowup %u", a.log_quotient_degree, lb);
        log_qd = std::max(log_qd, a.log_quotient_degree);
        if (traces[i].log_height + lb > 32) fail(MDN_ERR_DOMAIN, "LDE log order %u exceeds two-adicity 32", traces[i].log_height + lb);
        if (a.width == 0) fail(MDN_ERR_INVALID_ARG, "AIR %u has zero width", i);
        log_heights[i] = traces[i].log_height;
        // parse + upload the constraint program
        AirHost& h = airs[i];
        h.desc = a;
        Compiled cp = compile_oplist(i, a, st->n_public_values, a.program, a.program_words, false, 0);
        bool uses_sel = cp.uses_sel;
        u32 n_slots = cp.n_slots;
        std::vector<u32>& code = cp.code;
        u32 nk = (u32)cp.consts.size();
        std::vector<u64> hostp((code.size() + 1) / 2 + nk + 2, 0);
        memcpy(hostp.data(), code.data(), code.size() * sizeof(u32));
        size_t const_off = (code.size() + 1) / 2;
        for (u32 j = 0; j < nk; j++) hostp[const_off + j] = cp.consts[j];
        // periodic columns: table[col][m] = P_col(s^(n/maxp) * w_{maxp*B}^m), m < maxp*B
        // (prover/periodic.rs:49-98; the column is interpolated over the size-maxp subgroup)
        size_t per_off = hostp.size();
        if (a.num_periodic_columns) {
            if (!a.periodic_values) fail(MDN_ERR_INVALID_ARG, "AIR %u: periodic_values is NULL", i);
            if (a.log_max_period > traces[i].log_height) fail(MDN_ERR_INVALID_ARG, "AIR %u: periodic column period exceeds the trace height", i);
            size_t mp = (size_t)1 << a.log_max_period, np = a.num_periodic_columns, tl = mp << lb;
            u64 wp_inv = gl::inv(gl::two_adic_generator(a.log_max_period)), mp_inv = gl::inv((u64)mp);
            u64 sh = gl::exp_pow2(gl::lde_shift(traces[i].log_height + lb), traces[i].log_height - a.log_max_period);
            u64 wq = gl::two_adic_generator(a.log_max_period + lb);
            hostp.resize(per_off + np * tl);
            std::vector<u64> coef(mp);
            for (size_t c = 0; c < np; c++) {
                for (size_t kq = 0; kq < mp; kq++) {        // naive inverse DFT (periods are tiny)
                    u64 acc = 0, wk = gl::pow(wp_inv, kq), xx = 1;
                    for (size_t rr = 0; rr < mp; rr++) {
                        u64 v = a.periodic_values[rr * np + c];
                        if (v >= gl::P) fail(MDN_ERR_INVALID_ARG, "AIR %u: non-canonical periodic value", i);
                        acc = gl::add(acc, gl::mul(v, xx)); xx = gl::mul(xx, wk);
                    }
                    coef[kq] = gl::mul(acc, mp_inv);
                }
                u64 pt = sh;
                for (size_t m2 = 0; m2 < tl; m2++) {
                    u64 acc = 0;
                    for (size_t kq = mp; kq-- > 0;) acc = gl::add(gl::mul(acc, pt), coef[kq]);
                    hostp[per_off + c * tl + m2] = acc;
                    pt = gl::mul(pt, wq);
                }
            }
        }
        h.program.alloc(hostp.size(), stream);
        CUDA_OK(cudaMemcpyAsync(h.program.p, hostp.data(), hostp.size() * sizeof(u64), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        h.dev.code = (const u32*)h.program.p;
        h.dev.consts = h.program.p + const_off;
        h.dev.periodic = a.num_periodic_columns ? h.program.p + per_off : nullptr;
        h.dev.n_instr = (u32)(code.size() / 4); h.dev.n_slots = std::max(1u, n_slots); h.dev.uses_selectors = uses_sel;
        h.dev.log_max_period = a.log_max_period; h.dev.n_periodic = a.num_periodic_columns;
        // large programs: straight-line kernel compiled once per AIR (jit.hpp); the interpreter stays the fallback
        h.jit.reset(); h.n_constraints = a.program[3];
        if (jit_min_nodes && a.program[2] >= jit_min_nodes) {
            try {
                u64 key = jit::fnv1a(a.program, a.program_words) ^ ((u64)a.program_words << 40);
                auto it = jit_kernels.find(key);
                if (it == jit_kernels.end()) {
                    const std::vector<char>& cubin = jit::cubin_for(a.program, a.program_words, nullptr);
                    auto kn = std::make_shared<jit::Kernel>();
                    kn->load(cubin);
                    it = jit_kernels.emplace(key, kn).first;
                }
                h.jit = it->second;
            } catch (const std::exception& e) { jit_note = e.what(); }
        }
        // lowered LookupAir -> the aux trace of this AIR is built on the device (commit_aux)
        h.has_lookup = a.lookup != nullptr;
        if (h.has_lookup) {
            const mdn_lookup& lk = *a.lookup;
            if (lk.num_columns != a.aux_width || a.num_aux_values != 1) fail(MDN_ERR_INVALID_ARG, "AIR %u: a lookup program needs aux_width == num_columns and num_aux_values == 1", i);
            if (lk.num_columns == 0 || lk.num_columns > mk::LOGUP_MAX_COLS) fail(MDN_ERR_UNSUPPORTED, "AIR %u: at most %u lookup columns", i, mk::LOGUP_MAX_COLS);
            Compiled lc = compile_oplist(i, a, st->n_public_values, lk.program, lk.program_words, true, lk.num_columns);
            size_t c_off = (lc.code.size() + 1) / 2, p_off = c_off + lc.consts.size() + 2;
            size_t np = a.num_periodic_columns, mp = (size_t)1 << a.log_max_period;
            std::vector<u64> hp(p_off + np * mp, 0);
            memcpy(hp.data(), lc.code.data(), lc.code.size() * sizeof(u32));
            for (size_t j = 0; j < lc.consts.size(); j++) hp[c_off + j] = lc.consts[j];
            for (size_t q = 0; q < np * mp; q++) hp[p_off + q] = a.periodic_values[q];   // canonical: checked above
            h.lookup_program.alloc(hp.size(), stream);
            CUDA_OK(cudaMemcpyAsync(h.lookup_program.p, hp.data(), hp.size() * sizeof(u64), cudaMemcpyHostToDevice, stream));
            CUDA_OK(cudaStreamSynchronize(stream));
            h.lookup_dev = mk::AirDev{};
            h.lookup_dev.code = (const u32*)h.lookup_program.p; h.lookup_dev.consts = h.lookup_program.p + c_off;
            h.lookup_dev.periodic = np ? h.lookup_program.p + p_off : nullptr;
            h.lookup_dev.n_instr = (u32)(lc.code.size() / 4); h.lookup_dev.n_slots = std::max(1u, lc.n_slots);
            h.lookup_dev.log_max_period = a.log_max_period; h.lookup_dev.n_periodic = a.num_periodic_columns;
            h.lookup_jit.reset();
            if (jit_min_nodes && lk.program[2] >= jit_min_nodes) {
                try {
                    u64 key = jit::fnv1a(lk.program, lk.program_words) ^ ((u64)lk.program_words << 40) ^ 0x4C4B5550ull;
                    auto it = jit_kernels.find(key);
                    if (it == jit_kernels.end()) {
                        const std::vector<char>& cubin = jit::cubin_for(lk.program, lk.program_words, nullptr, true, lk.num_columns);
                        auto kn = std::make_shared<jit::Kernel>();
                        kn->load(cubin);
                        it = jit_kernels.emplace(key, kn).first;
                    }
                    h.lookup_jit = it->second;
                } catch (const std::exception& e) { jit_note = e.what(); }
            }
        }
    }
    // preprocessed presence / shape parity (ProverInstance::new, prover/mod.rs:139-153; validate_preprocessed,
    // preprocessed.rs:147-260): a bundle must be installed exactly when some AIR declares preprocessed columns,
    // and each committed trace must have its AIR's declared width and its main trace's height
    {
        bool expected = false;
        for (u32 i = 0; i < k; i++) expected |= st->airs[i].preprocessed_width > 0;
        if (expected != has_prep) fail(MDN_ERR_INVALID_ARG, "preprocessed presence mismatch: AIRs %s preprocessed columns but %s bundle is installed", expected ? "declare" : "declare no", has_prep ? "a" : "no");
        if (has_prep) {
            std::vector<int> idx(k, -1);
            for (size_t q = 0; q < prep_air.size(); q++) { if (prep_air[q] >= k) fail(MDN_ERR_INVALID_ARG, "preprocessed bundle was built for a different AIR list"); idx[prep_air[q]] = (int)q; }
            for (u32 i = 0; i < k; i++) {
                bool want = st->airs[i].preprocessed_width > 0;
                if (want != (idx[i] >= 0)) fail(MDN_ERR_INVALID_ARG, "AIR %u: preprocessed trace presence mismatch", i);
                if (!want) continue;
                if (prep_c.mats[idx[i]].width != st->airs[i].preprocessed_width) fail(MDN_ERR_INVALID_ARG, "AIR %u: preprocessed width %u does not match the declared %u", i, prep_c.mats[idx[i]].width, st->airs[i].preprocessed_width);
                if (prep_log_h[idx[i]] != traces[i].log_height) fail(MDN_ERR_INVALID_ARG, "AIR %u: preprocessed height 2^%u differs from the main trace height 2^%u", i, prep_log_h[idx[i]], traces[i].log_height);
            }
        }
    }
    publics.assign(st->public_values, st->public_values + st->n_public_values);
    // TraceOrder: stable sort on (log_height, instance)  (order.rs)
    order.resize(k);
    for (u32 i = 0; i < k; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return log_heights[a] < log_heights[b]; });
    log_max_n = log_heights[order.back()];
    for (u32 i = 0; i < k; i++) ntt(log_heights[i]);   // validates the supported range early

    // challenger: caller's pre-bound state, then Statement::observe + observe_shape (mod.rs:290-291)
    Duplex& ch = tr.ch;
    ch.perm = perm();
    if (byte_hash()) {
        ch.hashed = true; ch.keccak = (hash_kind == MDN_HASH_KECCAK); ch.bin = hash_ch_in; ch.bout = hash_ch_out;      // mdn_session_set_hash_challenger
        if (ch.keccak && ch.bin.size() % 8) fail(MDN_ERR_INVALID_ARG, "Keccak hash challenger: the input buffer must be whole 64-bit words");
    } else {
        if (!chal) fail(MDN_ERR_INVALID_ARG, "null challenger");
        for (int i = 0; i < 12; i++) ch.st[i] = chal->sponge_state[i];
        if (chal->input_len > 7 || chal->output_len > 8) fail(MDN_ERR_INVALID_ARG, "malformed challenger state");
        for (u32 i = 0; i < chal->input_len; i++) ch.in[i] = chal->input_buffer[i];
        ch.in_len = chal->input_len; ch.out_len = chal->output_len;
    }
    if (has_prep) ch.observe_digest(prep_c.root);   // preprocessed commitment first (mod.rs:282-286)
    for (u32 i = 0; i < st->n_observe_felts; i++) ch.observe(st->observe_felts[i]);
    ch.observe(k);
    for (u32 i = 0; i < k; i++) ch.observe(log_heights[i]);

    // 1. upload + transpose main traces (proof order), LDE, LMCS  (mod.rs:326-341)
    size_t coef_total = 0, lde_total = 0;
    for (u32 j = 0; j < k; j++) {
        u32 inst = order[j];
        size_t N = (size_t)1 << log_heights[inst];
        coef_total += N * airs[inst].desc.width; lde_total += (N << lb) * airs[inst].desc.width;
    }
    main_c.coef_buf.alloc(coef_total, stream);
    main_c.lde_buf.alloc(lde_total, stream);
    size_t co = 0, lo = 0;
    for (u32 j = 0; j < k; j++) {
        u32 inst = order[j];
        size_t N = (size_t)1 << log_heights[inst];
        u32 w = airs[inst].desc.width;
        main_c.mats.push_back(CommittedMat{main_c.lde_buf.p + lo, main_c.coef_buf.p + co, log_heights[inst], w});
        co += N * w; lo += (N << lb) * w;
    }
    // H2D copies run on a second stream, so the LDE of matrix j overlaps the copy of matrix j+1
    // (copies of pinned buffers are asynchronous; pageable buffers degrade to a staged copy).
    std::vector<DevBuf> staging(k);
    if (!on_device && sharded()) {
        // A proof split over G ranks: every rank copies 1/G of the rows of every trace to its GPU, transposes that slice
        // and stores it into the column-major trace of EVERY rank over NVLink, so each PCIe link carries 1/G of the
        // trace instead of all of it.  Traces too short to split are copied whole by every rank.
        const u32 lg = shard_log_g;
        auto slice_rows = [&](u32 ln) -> size_t { return ln >= lg + 5 ? ((size_t)1 << (ln - lg)) : ((size_t)1 << ln); };
        for (u32 j = 0; j < k; j++) staging[j].alloc(slice_rows(main_c.mats[j].log_n) * main_c.mats[j].width, stream);
        shard_barrier();                       // the coefficient buffer is a fresh allocation on every rank
        CUDA_OK(cudaEventRecord(copy_ev[7], stream));
        CUDA_OK(cudaStreamWaitEvent(copy_stream, copy_ev[7], 0));
        mk::PeerPtrs bad{};
        for (u32 g = 0; g < shard_world; g++) bad.p[g] = sync_flags.p[g] + 16;
        // smallest matrix first, and the LDE of a matrix is queued as soon as every rank's slice of it has arrived, so
        // the copies of the later matrices (copy stream) hide behind the LDE of the earlier ones, as on one GPU
        std::vector<u32> by_size(k);
        for (u32 j = 0; j < k; j++) by_size[j] = j;
        std::stable_sort(by_size.begin(), by_size.end(), [&](u32 a, u32 b) { return staging[a].n < staging[b].n; });
        for (u32 q = 0; q < k; q++) {
            const u32 j = by_size[q];
            const mdn_matrix& m = traces[order[j]];
            CommittedMat& cm = main_c.mats[j];
            const bool whole = cm.log_n < lg + 5;
            size_t rows = slice_rows(cm.log_n), row0 = whole ? 0 : rows * shard_rank;
            host_to_device(staging[j].p, m.values + row0 * cm.width, rows * cm.width);
            CUDA_OK(cudaEventRecord(copy_ev[q % 7], copy_stream));
            CUDA_OK(cudaStreamWaitEvent(stream, copy_ev[q % 7], 0));
            {
                ProfScope ps(prof, PC_TRANSPOSE);
                if (whole) mk::launch_transpose_rm_to_cm(staging[j].p, cm.coef, 1u << cm.log_n, cm.width, (u32*)d_flag.p, stream);
                else mk::launch_transpose_slice_push(staging[j].p, peers_of(cm.coef), bad, shard_world, (u32)row0, (u32)rows, 1u << cm.log_n, cm.width, stream);
            }
            shard_barrier();                   // every rank's slice of this matrix has arrived everywhere
            if (q + 1 == k) CUDA_OK(cudaEventRecord(ev[1], stream));
            keep_raw_main(j);
            lde_matrix(cm);
        }
    } else if (!on_device) {
        for (u32 j = 0; j < k; j++) staging[j].alloc(((size_t)1 << main_c.mats[j].log_n) * main_c.mats[j].width, stream);
        CUDA_OK(cudaEventRecord(copy_ev[7], stream));
        CUDA_OK(cudaStreamWaitEvent(copy_stream, copy_ev[7], 0));
        // smallest matrix first: only its copy is exposed, every later copy hides behind the LDE of its predecessors
        // (the committed order of the matrices does not depend on the order in which they are prepared)
        std::vector<u32> by_size(k);
        for (u32 j = 0; j < k; j++) by_size[j] = j;
        std::stable_sort(by_size.begin(), by_size.end(), [&](u32 a, u32 b) { return staging[a].n < staging[b].n; });
        for (u32 q = 0; q < k; q++) {
            u32 j = by_size[q];
            const mdn_matrix& m = traces[order[j]];
            host_to_device(staging[j].p, m.values, staging[j].n);
            CUDA_OK(cudaEventRecord(copy_ev[q % 7], copy_stream));
            CUDA_OK(cudaStreamWaitEvent(stream, copy_ev[q % 7], 0));
            {
                ProfScope ps(prof, PC_TRANSPOSE);
                mk::launch_transpose_rm_to_cm(staging[j].p, main_c.mats[j].coef, 1u << main_c.mats[j].log_n, main_c.mats[j].width, (u32*)d_flag.p, stream);
            }
            if (q + 1 == k) CUDA_OK(cudaEventRecord(ev[1], stream));
            keep_raw_main(j);
            lde_matrix(main_c.mats[j]);   // queued behind the copy of this matrix; overlaps the copy of the next one
        }
    } else {
        for (u32 j = 0; j < k; j++) upload_matrix(traces[order[j]], true, main_c.mats[j].coef);
        CUDA_OK(cudaEventRecord(ev[1], stream));
        for (u32 j = 0; j < k; j++) { keep_raw_main(j); lde_matrix(main_c.mats[j]); }
    }
    staging.clear();
    check_input_flag("a main trace");
    lde_and_commit(main_c, &timings.lde_main, &timings.hash_main, true);
    CUDA_OK(cudaEventRecord(ev[2], stream));
    tr.send_commitment(main_c.root);
    memcpy(dbg_roots[0], main_c.root, 32);
    // 2. randomness (mod.rs:344-349)
    u32 max_rand = 0;
    for (auto& a : airs) max_rand = std::max(max_rand, a.desc.num_randomness);
    for (u32 i = 0; i < max_rand; i++) randomness.push_back(tr.ch.sample_ext());
    in_proof = true;
}

// Preprocessed::build (preprocessed.rs:63-131): the declared matrices sorted by (height, AIR index), coset LDE on
// the canonical shift of their own height, one aligned LMCS tree.  Kept on the device until replaced.
void mdn_session::set_preprocessed(const mdn_statement* st, const mdn_matrix* mats) {
    if (in_proof) fail(MDN_ERR_INVALID_ARG, "set_preprocessed called inside a proof");
    prep_c = Committed(); has_prep = false; prep_air.clear(); prep_log_h.clear();
    if (!mats) return;
    if (!st) fail(MDN_ERR_INVALID_ARG, "null argument");
    u32 lb = params.log_blowup;
    if (lb == 0 || lb > 4) fail(MDN_ERR_UNSUPPORTED, "log_blowup must be in 1..=4");
    if (!d_flag.p) { ArenaScope persistent(nullptr); d_flag.alloc(1, stream); CUDA_OK(cudaMemsetAsync(d_flag.p, 0, 8, stream)); }
    std::vector<u32> ids;
    for (u32 i = 0; i < st->n_airs; i++) {
        if (mats[i].width != st->airs[i].preprocessed_width) fail(MDN_ERR_INVALID_ARG, "AIR %u: preprocessed matrix width %u does not match the declared %u", i, mats[i].width, st->airs[i].preprocessed_width);
        if (!mats[i].width) continue;
        if (!mats[i].values) fail(MDN_ERR_INVALID_ARG, "AIR %u: preprocessed matrix is NULL", i);
        if (mats[i].log_height + lb > 32) fail(MDN_ERR_DOMAIN, "LDE log order %u exceeds two-adicity 32", mats[i].log_height + lb);
        ntt(mats[i].log_height);
        ids.push_back(i);
    }
    if (ids.empty()) return;
    std::stable_sort(ids.begin(), ids.end(), [&](u32 a, u32 b) { return mats[a].log_height < mats[b].log_height; });
    size_t coef_total = 0, lde_total = 0;
    for (u32 i : ids) { size_t N = (size_t)1 << mats[i].log_height; coef_total += N * mats[i].width; lde_total += (N << lb) * mats[i].width; }
    prep_c.coef_buf.alloc(coef_total, stream); prep_c.lde_buf.alloc(lde_total, stream);
    size_t co = 0, lo = 0;
    for (u32 i : ids) {
        size_t N = (size_t)1 << mats[i].log_height;
        prep_c.mats.push_back(CommittedMat{prep_c.lde_buf.p + lo, prep_c.coef_buf.p + co, mats[i].log_height, mats[i].width});
        co += N * mats[i].width; lo += (N << lb) * mats[i].width;
        upload_matrix(mats[i], false, prep_c.mats.back().coef);
    }
    check_input_flag("a preprocessed trace");
    lde_and_commit(prep_c, nullptr, nullptr);
    prep_air = ids;
    for (u32 i : ids) prep_log_h.push_back(mats[i].log_height);
    has_prep = true;
}

void mdn_session::keep_raw_main(u32 j) {
    AirHost& h = airs[order[j]];
    if (!h.has_lookup) return;
    CommittedMat& m = main_c.mats[j];
    size_t n = ((size_t)1 << m.log_n) * m.width;
    h.raw_main.alloc(n, stream);
    CUDA_OK(cudaMemcpyAsync(h.raw_main.p, m.coef, n * sizeof(u64), cudaMemcpyDeviceToDevice, stream));
}

// build_logup_aux_trace (air/src/lookup/aux_builder.rs:49-97) for proof position j: fraction collection and
// per-row sums on the trace domain, exclusive EF prefix sum into column 0, committed final = the grand total.
void mdn_session::build_logup_aux(u32 j, u64* aux_cm, u64 final_out[2]) {
    AirHost& h = airs[order[j]];
    u32 ln = log_heights[order[j]];
    size_t N = (size_t)1 << ln;
    DevBuf totals, scratch, fin;
    totals.alloc(2 * N, stream); scratch.alloc(2 * ((N + 2047) / 2048) + 2, stream); fin.alloc(2, stream);
    mk::LogupArgs la;
    la.main_cm = h.raw_main.p; la.log_n = ln; la.n_cols = h.desc.aux_width; la.prog = h.lookup_dev;
    la.publics = d_publics.p; la.challenges = d_randomness.p; la.aux_cm = aux_cm; la.totals = totals.p; la.bad_flag = (u32*)d_flag.p;
    {
        ProfScope ps(prof, PC_CONSTRAINTS);
        if (h.lookup_jit && h.lookup_jit->checked < 0) h.lookup_jit.reset();
        if (h.lookup_jit) {
            jit::LookupJitArgs ja{};
            ja.main_lde = la.main_cm; ja.publics = la.publics; ja.challenges = la.challenges;
            ja.periodic = h.lookup_dev.periodic; ja.aux_cm = aux_cm; ja.totals = totals.p; ja.bad_flag = la.bad_flag;
            ja.log_n = ln; ja.n_periodic = h.lookup_dev.n_periodic; ja.log_max_period = h.lookup_dev.log_max_period;
            try { h.lookup_jit->launch(ja, (unsigned)((N + 127) / 128), 128, stream); }
            catch (const std::exception& e) { fail(MDN_ERR_CUDA, "%s", e.what()); }
            mk::count_launch();
            if (h.lookup_jit->checked == 0) {
                // first use: the interpreter builds the same rows into scratch buffers; fraction columns and row totals
                // must agree word for word, otherwise the interpreter's result is kept and the kernel is retired
                DevBuf aux2, tot2;
                u32 C = h.desc.aux_width;
                aux2.alloc(2 * (size_t)C * N, stream); tot2.alloc(2 * N, stream);
                mk::LogupArgs lb2 = la; lb2.aux_cm = aux2.p; lb2.totals = tot2.p;
                if (mk::launch_logup_rows(lb2, stream) != 0) fail(MDN_ERR_UNSUPPORTED, "lookup program too large for the interpreter");
                if (C > 1) mk::launch_compare(aux2.p + 2 * N, aux_cm + 2 * N, 2 * (size_t)(C - 1) * N, (u32*)d_flag.p, stream);
                mk::launch_compare(tot2.p, totals.p, 2 * N, (u32*)d_flag.p, stream);
                u32 flag = 0;
                CUDA_OK(cudaMemcpyAsync(&flag, d_flag.p, sizeof flag, cudaMemcpyDeviceToHost, stream));
                CUDA_OK(cudaStreamSynchronize(stream));
                if (flag & 4) {
                    // clear the comparison bit only: a zero-denominator (2) or input (1) report raised by the same kernels stays
                    u32 keep = flag & ~4u;
                    CUDA_OK(cudaMemsetAsync(d_flag.p, 0, 8, stream));
                    if (keep) { CUDA_OK(cudaMemcpyAsync(d_flag.p, &keep, sizeof keep, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream)); }
                    if (C > 1) CUDA_OK(cudaMemcpyAsync(aux_cm + 2 * N, aux2.p + 2 * N, 2 * (size_t)(C - 1) * N * sizeof(u64), cudaMemcpyDeviceToDevice, stream));
                    CUDA_OK(cudaMemcpyAsync(totals.p, tot2.p, 2 * N * sizeof(u64), cudaMemcpyDeviceToDevice, stream));
                    CUDA_OK(cudaStreamSynchronize(stream));
                    h.lookup_jit->checked = -1;
                    jit_note = "NVRTC lookup kernel disagreed with the interpreter on its first use; interpreter kept";
                } else h.lookup_jit->checked = 1;
            }
        } else if (mk::launch_logup_rows(la, stream) != 0) fail(MDN_ERR_UNSUPPORTED, "lookup program too large for the interpreter");
        mk::launch_ef_exclusive_scan(totals.p, N, aux_cm, aux_cm + N, fin.p, scratch.p, stream);
    }
    CUDA_OK(cudaMemcpyAsync(final_out, fin.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, stream));
    u32 flag = 0;
    CUDA_OK(cudaMemcpyAsync(&flag, d_flag.p, sizeof flag, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    h.raw_main.release();
    if (flag) {
        CUDA_OK(cudaMemsetAsync(d_flag.p, 0, 8, stream));
        if (flag & 2) fail(MDN_ERR_INVALID_ARG, "AIR %u: LogUp denominator must be non-zero", order[j]);   // aux_builder.rs:240-243
        fail(MDN_ERR_INVALID_ARG, "an aux trace contains a non-canonical field element (>= p)");
    }
}

// aux traces (instance order, EF flattened to base), aux values; commit + observe (mod.rs:397-422)
void mdn_session::commit_aux(const mdn_matrix* aux, const u64* const* aux_values, bool zero_aux) {
    if (!in_proof) fail(MDN_ERR_INVALID_ARG, "commit_aux called outside a proof");
    u32 k = (u32)airs.size(), lb = params.log_blowup;
    CUDA_OK(cudaEventRecord(ev[3], stream));
    size_t coef_total = 0, lde_total = 0;
    for (u32 j = 0; j < k; j++) {
        u32 inst = order[j];
        size_t N = (size_t)1 << log_heights[inst];
        u32 w = 2 * airs[inst].desc.aux_width;
        coef_total += N * w; lde_total += (N << lb) * w;
    }
    aux_c.coef_buf.alloc(coef_total, stream);
    aux_c.lde_buf.alloc(lde_total, stream);
    // device copies of the small per-proof vectors used by the lookup and constraint kernels
    d_publics.alloc(std::max<size_t>(1, publics.size()), stream);
    if (!publics.empty()) CUDA_OK(cudaMemcpyAsync(d_publics.p, publics.data(), publics.size() * sizeof(u64), cudaMemcpyHostToDevice, stream));
    d_randomness.alloc(std::max<size_t>(1, 2 * randomness.size()), stream);
    if (!randomness.empty()) CUDA_OK(cudaMemcpyAsync(d_randomness.p, randomness.data(), randomness.size() * sizeof(E2), cudaMemcpyHostToDevice, stream));
    size_t co = 0, lo = 0;
    std::vector<u64> flat_values;
    aux_values_p.assign(k, {}); aux_values_off.assign(k, 0);
    for (u32 j = 0; j < k; j++) {
        u32 inst = order[j];
        size_t N = (size_t)1 << log_heights[inst];
        u32 w = 2 * airs[inst].desc.aux_width;
        aux_c.mats.push_back(CommittedMat{aux_c.lde_buf.p + lo, aux_c.coef_buf.p + co, log_heights[inst], w});
        u64 logup_final[2] = {0, 0};
        const bool dev_aux = airs[inst].has_lookup;
        if (dev_aux) build_logup_aux(j, aux_c.coef_buf.p + co, logup_final);
        else if (w) {
            if (zero_aux) CUDA_OK(cudaMemsetAsync(aux_c.coef_buf.p + co, 0, N * w * sizeof(u64), stream));
            else {
                if (!aux[inst].values) fail(MDN_ERR_INVALID_ARG, "aux trace %u is NULL", inst);
                if (aux[inst].width != w || aux[inst].log_height != log_heights[inst]) fail(MDN_ERR_INVALID_ARG, "aux trace %u has the wrong shape", inst);
                upload_matrix(aux[inst], false, aux_c.coef_buf.p + co);
            }
        }
        u32 nav = airs[inst].desc.num_aux_values;
        aux_values_off[j] = flat_values.size();
        if (nav && !dev_aux && !zero_aux && (!aux_values || !aux_values[inst])) fail(MDN_ERR_INVALID_ARG, "aux values of instance %u are NULL", inst);
        for (u32 v = 0; v < 2 * nav; v++) {
            u64 x = dev_aux ? logup_final[v] : (zero_aux ? 0 : aux_values[inst][v]);
            if (x >= gl::P) fail(MDN_ERR_INVALID_ARG, "non-canonical aux value");
            aux_values_p[j].push_back(x); flat_values.push_back(x);
        }
        co += N * w; lo += (N << lb) * w;
    }
    if (!zero_aux) check_input_flag("an aux trace");
    // Statement::eval_external on the aux values in instance order -- including the finals of aux traces built on the
    // device -- before anything is committed (prover/mod.rs:383-395, ProverError::ExternalAssertionFailed)
    if (external_check) {
        std::vector<std::vector<u64>> by_inst(k);
        for (u32 j = 0; j < k; j++) by_inst[order[j]] = aux_values_p[j];
        std::vector<const u64*> vp(k); std::vector<u32> vn(k); std::vector<uint8_t> lh(k);
        for (u32 i = 0; i < k; i++) { vp[i] = by_inst[i].data(); vn[i] = (u32)by_inst[i].size() / 2; lh[i] = (uint8_t)log_heights[i]; }
        u32 failed = 0;
        int rc = external_check(external_ctx, (const u64*)randomness.data(), (u32)randomness.size(), vp.data(), vn.data(), lh.data(), k, &failed);
        if (rc > 0) fail(MDN_ERR_EXTERNAL_ASSERTION, "external assertion %u failed", failed);
        if (rc < 0) fail(MDN_ERR_EXTERNAL_ASSERTION, "eval_external reported a reduction error");
    }
    lde_and_commit(aux_c, nullptr, nullptr);
    tr.send_commitment(aux_c.root);
    memcpy(dbg_roots[1], aux_c.root, 32);
    for (auto& vs : aux_values_p) for (u64 v : vs) tr.send_field(v);
    d_aux_values.alloc(std::max<size_t>(1, flat_values.size()), stream);
    if (!flat_values.empty()) CUDA_OK(cudaMemcpyAsync(d_aux_values.p, flat_values.data(), flat_values.size() * sizeof(u64), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaEventRecord(ev[4], stream));
}

// ---------------------------------------------------------------------------------------------
// finish: constraints, quotient commit, OOD point, PCS opening  (mod.rs:424-577)
// ---------------------------------------------------------------------------------------------
void mdn_session::finish() {
    if (!in_proof) fail(MDN_ERR_INVALID_ARG, "finish called outside a proof");
    u32 k = (u32)airs.size(), lb = params.log_blowup, B = 1u << lb;
    u32 log_lde = log_max_n + lb;
    size_t Nmax = (size_t)1 << log_max_n, L = Nmax << lb;
    // 3. alpha, beta (mod.rs:425-426)
    E2 alpha = tr.ch.sample_ext(), beta = tr.ch.sample_ext();
    // 4. constraint evaluation + beta accumulation, ascending height (mod.rs:445-537)
    DevBuf acc_pp[2];
    std::vector<DevBuf> jit_apow;   // alive until the constraint kernels have run
    jit_used.clear();
    int acc_cur = -1;
    u32 acc_prev_log = 0;
    for (u32 j = 0; j < k; j++) {
        u32 inst = order[j];
        AirHost& air = airs[inst];
        u32 ln = log_heights[inst];
        int nxt = acc_cur < 0 ? 0 : 1 - acc_cur;
        acc_pp[nxt].alloc((size_t)2 << (ln + lb), stream);
        mk::ConstraintArgs ca;
        ca.main_lde = main_c.mats[j].lde; ca.main_width = main_c.mats[j].width;
        ca.aux_lde = aux_c.mats[j].lde; ca.aux_width_base = aux_c.mats[j].width;
        ca.prep_lde = nullptr;   // same height and coset as the main trace (mod.rs:463-476)
        if (has_prep) for (size_t q = 0; q < prep_air.size(); q++) if (prep_air[q] == inst) ca.prep_lde = prep_c.mats[q].lde;
        ca.log_n = ln; ca.log_blowup = lb; ca.air = air.dev;
        ca.publics = d_publics.p; ca.challenges = d_randomness.p; ca.aux_values = d_aux_values.p + aux_values_off[j];
        ca.alpha = alpha; ca.beta = beta;
        ca.acc_in = acc_cur < 0 ? nullptr : acc_pp[acc_cur].p; ca.acc_in_log_n = acc_prev_log; ca.acc_out = acc_pp[nxt].p;
        ca.T = &ntt(ln).T;
        ca.t0 = t0(); ca.nt = nt();            // this rank's cosets (all of them on one GPU)
        ProfScope ps(prof, PC_CONSTRAINTS);
        if (air.jit && air.jit->checked < 0) air.jit.reset();   // failed its self-check earlier in this session
        jit_used.push_back(air.jit ? 1 : 0);
        if (air.jit) {
            // alpha^(K-1-k) for the K constraints in emission order
            u32 K = air.n_constraints;
            std::vector<E2> apow(std::max(1u, K));
            { E2 x = gl::e2(1, 0); for (u32 q = K; q-- > 0;) { apow[q] = x; x = gl::e2_mul(x, alpha); } }
            jit_apow.emplace_back(); jit_apow.back().alloc(2 * (size_t)std::max(1u, K), stream);
            CUDA_OK(cudaMemcpyAsync(jit_apow.back().p, apow.data(), apow.size() * sizeof(E2), cudaMemcpyHostToDevice, stream));
            CUDA_OK(cudaStreamSynchronize(stream));   // apow is a stack vector
            jit::JitArgs ja{};
            ja.main_lde = ca.main_lde; ja.aux_lde = ca.aux_lde; ja.prep_lde = ca.prep_lde;
            ja.publics = ca.publics; ja.challenges = ca.challenges; ja.aux_values = ca.aux_values;
            ja.periodic = air.dev.periodic; ja.apow = jit_apow.back().p;
            ja.acc_in = ca.acc_in; ja.acc_out = ca.acc_out; ja.w_hi = ca.T->w_hi; ja.w_lo = ca.T->w_lo;
            ja.shift = gl::lde_shift(ln + lb); ja.w_l = gl::two_adic_generator(ln + lb); ja.w_h_inv = gl::inv(gl::two_adic_generator(ln));
            { u64 s_pow_n = gl::exp_pow2(ja.shift, ln), w_b = gl::two_adic_generator(lb), x = 1;   // Z_H on coset t (domain.rs:742-749)
              for (u32 t = 0; t < B; t++) { ja.zh[t] = gl::sub(gl::mul(s_pow_n, x), 1); ja.inv_zh[t] = gl::inv(ja.zh[t]); x = gl::mul(x, w_b); } }
            ja.beta_a = beta.a; ja.beta_b = beta.b;
            ja.log_n = ln; ja.log_b = lb; ja.acc_in_log_n = ca.acc_in_log_n; ja.lo_bits = ca.T->lo_bits; ja.log_max_period = air.dev.log_max_period;
            size_t Lj = (size_t)1 << (ln + lb), own = (size_t)ca.nt << ln;
            ja.pad = ca.t0 | (ca.nt << 8);
            try { air.jit->launch(ja, (unsigned)((own + 127) / 128), 128, stream); }
            catch (const std::exception& e) { fail(MDN_ERR_CUDA, "%s", e.what()); }
            mk::count_launch();
            if (air.jit->checked == 0) {
                // first use of this compiled kernel in the session: the interpreter evaluates the same points and
                // the two accumulators must agree word for word; on disagreement (a compiler defect) the interpreter's
                // result is kept, the kernel is retired and the reason is recorded
                DevBuf chk; chk.alloc(2 * Lj, stream);
                mk::ConstraintArgs cb = ca; cb.acc_out = chk.p;
                if (mk::launch_constraints(cb, stream) != 0) fail(MDN_ERR_UNSUPPORTED, "constraint program too large for the interpreter");
                for (u32 coord = 0; coord < 2; coord++) {     // the points this rank evaluated
                    size_t o = (size_t)coord * Lj + ((size_t)ca.t0 << ln);
                    mk::launch_compare(chk.p + o, ca.acc_out + o, own, (u32*)d_flag.p, stream);
                }
                u32 flag = 0;
                CUDA_OK(cudaMemcpyAsync(&flag, d_flag.p, sizeof flag, cudaMemcpyDeviceToHost, stream));
                CUDA_OK(cudaStreamSynchronize(stream));
                if (sharded()) {
                    // every rank must take the same branch (the ranks' allocation sequences have to stay identical)
                    std::vector<u64> mine(1, flag & 4), all(shard_world);
                    if (allgather(allgather_ctx, mine.data(), all.data(), 1) != 0) fail(MDN_ERR_INVALID_ARG, "all-gather callback failed");
                    for (u64 f : all) flag |= (u32)f;
                }
                if (flag & 4) {
                    CUDA_OK(cudaMemsetAsync(d_flag.p, 0, 8, stream));
                    if (flag & ~4u) { u32 keep = flag & ~4u; CUDA_OK(cudaMemcpyAsync(d_flag.p, &keep, sizeof keep, cudaMemcpyHostToDevice, stream)); CUDA_OK(cudaStreamSynchronize(stream)); }
                    CUDA_OK(cudaMemcpyAsync(ca.acc_out, chk.p, 2 * Lj * sizeof(u64), cudaMemcpyDeviceToDevice, stream));
                    CUDA_OK(cudaStreamSynchronize(stream));
                    air.jit->checked = -1; jit_used.back() = 0;
                    jit_note = "NVRTC kernel disagreed with the interpreter on its first use; interpreter kept";
                } else air.jit->checked = 1;
            }
        } else if (mk::launch_constraints(ca, stream) != 0) fail(MDN_ERR_UNSUPPORTED, "constraint program too large for the interpreter");
        acc_cur = nxt; acc_prev_log = ln;
    }
    DevBuf acc = std::move(acc_pp[acc_cur]);
    acc_pp[1 - acc_cur].release();
    CUDA_OK(cudaEventRecord(ev[5], stream));
    const u32 tb = t0(), tn = nt();
    if (keep_debug) {
        if (sharded()) {   // debug export only: collect every rank's cosets of the accumulator
            shard_barrier();
            for (u32 coord = 0; coord < 2; coord++) { u64* own = acc.p + (size_t)coord * L + (size_t)tb * Nmax; mk::launch_push(own, peers_of(own), shard_rank, shard_world, (size_t)tn * Nmax, stream); }
            shard_barrier();
        }
        // natural order on gJ: index r*B + t  <- planes [coord][t*N + r]
        std::vector<u64> planes(2 * L);
        CUDA_OK(cudaMemcpyAsync(planes.data(), acc.p, 2 * L * sizeof(u64), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        // natural order on gJ_max (N*D points): index r*D + t  <- coset t*(B/D) of the planes
        u32 Dq = 1u << log_qd, cs = B >> log_qd;
        dbg_quot_acc.assign(2 * Nmax * Dq, 0);
        for (size_t t = 0; t < Dq; t++)
            for (size_t r = 0; r < Nmax; r++) {
                dbg_quot_acc[2 * (r * Dq + t)] = planes[t * cs * Nmax + r];
                dbg_quot_acc[2 * (r * Dq + t) + 1] = planes[L + t * cs * Nmax + r];
            }
    }
    // 5. quotient commit (quotient.rs:143-217).  The accumulator was evaluated on all B cosets of gK
    //    (for a satisfied AIR that equals the reference's evaluate-on-gJ-then-upsample, quotient.rs:45-56,
    //    because C/Z_H is then a polynomial of degree < N*D); chunk t < D is the LDE coset t*(B/D).
    //    The committed matrix has column 2t + coord.
    //    Split over ranks: chunk t lives on the rank that owns coset t*(B/D); that rank interpolates it and stores
    //    the coefficients into every rank (the all-gather of chunk coefficients of SURVEY 8(e), as peer stores),
    //    then every rank evaluates all chunks on its own cosets.
    const u32 D = 1u << log_qd, cstep = B >> log_qd;
    {
        NttPlan& plan = ntt(log_max_n);
        PremulPlan& pm = premul_quotient(log_max_n, log_qd);
        size_t rq = prof.begin(PC_NTT);
        ntt_bytes += (double)(Nmax + L) * 2 * D * 8.0 / (sharded() ? shard_world : 1);
        // own chunks: c with c*cstep in [tb, tb + tn)
        u32 c_lo = (tb + cstep - 1) / cstep, c_hi = std::min(D, (tb + tn + cstep - 1) / cstep);
        if (c_hi > c_lo)
            for (u32 coord = 0; coord < 2; coord++)
                mk::launch_intt(acc.p + (size_t)coord * L + (size_t)c_lo * cstep * Nmax, (size_t)cstep * Nmax, c_hi - c_lo, plan.T, stream);
        if (sharded()) {
            shard_barrier();   // every rank is done writing / reading its accumulator planes
            for (u32 c = c_lo; c < c_hi; c++)
                for (u32 coord = 0; coord < 2; coord++) {
                    u64* chunk = acc.p + (size_t)coord * L + (size_t)c * cstep * Nmax;
                    mk::launch_push(chunk, peers_of(chunk), shard_rank, shard_world, Nmax, stream);
                }
            shard_barrier();
        }
        quot_c.lde_buf.alloc(L * 2 * D, stream);
        quot_c.coef_buf = std::move(acc);
        quot_c.mats.push_back(CommittedMat{quot_c.lde_buf.p, quot_c.coef_buf.p, log_max_n, 2 * D});
        std::vector<mk::FwdItem> items;
        for (u32 t = 0; t < D; t++)
            for (u32 coord = 0; coord < 2; coord++)
                for (u32 t2 = tb; t2 < tb + tn; t2++)
                    items.push_back(mk::FwdItem{quot_c.coef_buf.p + (size_t)coord * L + (size_t)t * cstep * Nmax,
                                                quot_c.lde_buf.p + (size_t)(2 * t + coord) * L + (size_t)t2 * Nmax, t * B + t2, 0});
        DevBuf d_items; d_items.alloc(items.size() * sizeof(mk::FwdItem) / sizeof(u64), stream);
        CUDA_OK(cudaMemcpyAsync(d_items.p, items.data(), items.size() * sizeof(mk::FwdItem), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        // groups of whole chunks keep the working set near L2 size
        u32 per = 2 * tn;   // items per chunk t
        u32 chunks_per_launch = std::max(1u, B / tn);
        for (u32 t = 0; t < D; t += chunks_per_launch) {
            u32 cn = std::min(chunks_per_launch, D - t);
            mk::launch_fwd_ntt((const mk::FwdItem*)d_items.p + (size_t)t * per, cn * per, plan.T, pm.P, stream);
        }
        prof.end(rq);
        build_tree(quot_c);
        tr.send_commitment(quot_c.root);
        memcpy(dbg_roots[2], quot_c.root, 32);
    }
    CUDA_OK(cudaEventRecord(ev[6], stream));
    // 6. OOD point: resample while z = 0, z in H, or z in gK  (domain.rs:539-552)
    u64 shift = gl::lde_shift(log_lde), shift_inv = gl::inv(shift);
    E2 z;
    for (;;) {
        z = tr.ch.sample_ext();
        if (z.a == 0 && z.b == 0) continue;
        if (gl::e2_eq(gl::e2_exp_pow2(z, log_max_n), gl::e2(1, 0))) continue;
        if (gl::e2_eq(gl::e2_exp_pow2(gl::e2_mulf(z, shift_inv), log_lde), gl::e2(1, 0))) continue;
        break;
    }
    ood_z = z;
    u64 omega_h = gl::two_adic_generator(log_max_n);
    E2 z_next = gl::e2_mulf(z, omega_h);

    // 7. PCS opening (pcs/prover.rs:34-102)
    // 7a. OOD evaluations of every committed column at z^(r_m), (z*w_H)^(r_m)
    //     (deep/interpolate.rs:127-203; computed here from the coefficient columns)
    // group order [preprocessed?, main, aux, quotient] (mod.rs:551-559)
    std::vector<Committed*> groups;
    if (has_prep) groups.push_back(&prep_c);
    groups.push_back(&main_c); groups.push_back(&aux_c); groups.push_back(&quot_c);
    const int ng = (int)groups.size(), gq = ng - 1;
    struct MatEval { std::vector<u64> v; };   // width x 4
    std::vector<std::vector<MatEval>> evals(ng);
    {
        // Split over ranks: a column's dot product with the weight vector is a sum over coefficient slots, so rank g
        // takes slots [g*N/G, (g+1)*N/G) of every (tall enough) column -- it builds only that slice of each weight
        // vector -- and the G partial sums per column meet in every rank's buffer (peer stores) and are added on the host.
        size_t ood_region = prof.begin(PC_OOD);
        // all dot products are queued first and fetched with ONE device->host copy
        size_t total_cols = 0;
        for (int g = 0; g < ng; g++) for (auto& cm : groups[g]->mats) total_cols += cm.width;
        const size_t T = std::max<size_t>(1, total_cols * 4);
        const u32 G = sharded() ? shard_world : 1, me = sharded() ? shard_rank : 0, lg = sharded() ? shard_log_g : 0;
        DevBuf d_all; d_all.alloc(T * G, stream);
        u64* d_out = d_all.p + (size_t)me * T;
        auto sliced = [&](u32 ln) { return G > 1 && ln >= lg + shard_min_log; };
        std::vector<DevBuf> keep;                              // weight vectors / partial sums stay alive until the copy
        std::map<u32, std::pair<u64*, u64*>> weights;          // per log height: (w0, w1), this rank's slice
        auto make_w = [&](u32 ln, E2 y0, E2 y1) {
            size_t S = sliced(ln) ? ((size_t)1 << (ln - lg)) : ((size_t)1 << ln), p0 = sliced(ln) ? S * me : 0;
            keep.emplace_back(); keep.back().alloc(2 * S, stream); u64* a = keep.back().p;
            keep.emplace_back(); keep.back().alloc(2 * S, stream); u64* b = keep.back().p;
            size_t tw = 2 * (((size_t)1 << (ln - ln / 2)) + ((size_t)1 << (ln / 2)));
            keep.emplace_back(); keep.back().alloc(2 * tw, stream); u64* sc = keep.back().p;
            mk::launch_pow_bitrev(y0, ln, a, sc, p0, S, stream);
            mk::launch_pow_bitrev(y1, ln, b, sc + tw, p0, S, stream);
            return std::make_pair(a, b);
        };
        size_t col_off = 0;
        struct Scale { size_t off; u64 n_inv; bool sliced; };
        std::vector<Scale> scale;                              // (first u64 index, 1/N, summed over ranks?) per matrix
        for (int g = 0; g < ng; g++) {
            evals[g].resize(groups[g]->mats.size());
            for (size_t m = 0; m < groups[g]->mats.size(); m++) {
                CommittedMat& cm = groups[g]->mats[m];
                evals[g][m].v.assign((size_t)cm.width * 4, 0);
                if (!cm.width) continue;
                u32 ln = cm.log_n, lr = log_max_n - ln;
                size_t Nm = (size_t)1 << ln;
                const bool sl = sliced(ln);
                const u32 ls = sl ? ln - lg : ln;                      // log of the slice this rank sums over
                const size_t S = (size_t)1 << ls, s0 = sl ? S * me : 0;
                u32 n_chunks = (u32)std::max<size_t>(1, std::min<size_t>(S / 4096, 256));   // k_ood_reduce walks the chunks serially
                if (g != gq) {
                    auto it = weights.find(ln);
                    if (it == weights.end()) it = weights.emplace(ln, make_w(ln, gl::e2_exp_pow2(z, lr), gl::e2_exp_pow2(z_next, lr))).first;
                    keep.emplace_back(); keep.back().alloc((size_t)cm.width * n_chunks * 4, stream);
                    mk::launch_ood_dot(cm.coef + s0, Nm, cm.width, ls, it->second.first, it->second.second, keep.back().p, n_chunks, stream);
                    mk::launch_ood_reduce(keep.back().p, cm.width, n_chunks, d_out + col_off * 4, stream);
                } else {
                    // quotient chunk t: stored coefficients are a_k * (g*w_J^t)^k (planes coord, column t*(B/D)),
                    // so q_t(y) is their evaluation at y / (g * w_J^t); outputs land at columns 2t, 2t+1.
                    u64 wj_inv = gl::inv(gl::two_adic_generator(log_max_n + log_qd));
                    for (u32 t = 0; t < D; t++) {
                        u64 f = gl::mul(shift_inv, gl::pow(wj_inv, t));
                        auto wv = make_w(ln, gl::e2_mulf(z, f), gl::e2_mulf(z_next, f));
                        keep.emplace_back(); keep.back().alloc((size_t)2 * n_chunks * 4, stream);
                        mk::launch_ood_dot(cm.coef + (size_t)t * cstep * Nm + s0, (size_t)B * Nm, 2, ls, wv.first, wv.second, keep.back().p, n_chunks, stream);
                        mk::launch_ood_reduce(keep.back().p, 2, n_chunks, d_out + (col_off + 2 * t) * 4, stream);
                    }
                }
                scale.push_back(Scale{col_off * 4, gl::inv((u64)Nm), sl});   // launch_intt leaves coefficients scaled by N
                col_off += cm.width;
            }
        }
        prof.end(ood_region);
        if (G > 1) {
            shard_barrier();                                   // d_all is a fresh allocation on every rank
            mk::launch_push(d_out, peers_of(d_out), shard_rank, shard_world, T, stream);
            shard_barrier();
        }
        std::vector<u64> host(T * G);
        CUDA_OK(cudaMemcpyAsync(host.data(), d_all.p, host.size() * sizeof(u64), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        size_t si = 0;
        for (int g = 0; g < ng; g++)
            for (size_t m = 0; m < groups[g]->mats.size(); m++) {
                CommittedMat& cm = groups[g]->mats[m];
                if (!cm.width) continue;
                const Scale& sc = scale[si++];
                for (size_t q = 0; q < (size_t)cm.width * 4; q++) {
                    u64 v = host[(size_t)me * T + sc.off + q];
                    if (sc.sliced) { v = 0; for (u32 r = 0; r < G; r++) v = gl::add(v, host[(size_t)r * T + sc.off + q]); }
                    evals[g][m].v[q] = gl::mul(v, sc.n_inv);
                }
            }
    }
    // aligned flat evaluation lists per point (deep/prover.rs:150-154)
    std::vector<E2> flat[2];
    std::vector<u32> aligned_off;   // per (group, matrix): offset in the aligned index space
    u32 W = 0;
    for (int g = 0; g < ng; g++)
        for (size_t m = 0; m < groups[g]->mats.size(); m++) {
            u32 w = groups[g]->mats[m].width, aw = (w + align() - 1) / align() * align();   // Lmcs alignment: 8 for the sponge, 1 for the chaining hasher
            aligned_off.push_back(W);
            for (int p = 0; p < 2; p++) {
                for (u32 c = 0; c < w; c++) {
                    flat[p].push_back(gl::e2(evals[g][m].v[4 * c + 2 * p], evals[g][m].v[4 * c + 2 * p + 1]));
                }
                for (u32 c = w; c < aw; c++) flat[p].push_back(gl::e2(0, 0));
            }
            W += aw;
        }
    for (int p = 0; p < 2; p++) for (const E2& e : flat[p]) tr.send_ext(e);
    // 7b. DEEP grind + challenges (deep/prover.rs:157-162)
    grind(params.deep_pow_bits);
    E2 dalpha = tr.ch.sample_ext(), dbeta = tr.ch.sample_ext();
    E2 fz[2];
    for (int p = 0; p < 2; p++) { E2 a = gl::e2(0, 0); for (const E2& e : flat[p]) a = gl::e2_add(gl::e2_mul(a, dalpha), e); fz[p] = a; }
    std::vector<E2> apow(W);
    { E2 a = gl::e2(1, 0); for (u32 i = W; i-- > 0;) { apow[i] = a; a = gl::e2_mul(a, dalpha); } }
    // 7c. DEEP quotient over the LDE domain (deep/prover.rs:214-312)
    DevBuf d_apow; d_apow.alloc(2 * (size_t)W, stream);
    CUDA_OK(cudaMemcpyAsync(d_apow.p, apow.data(), W * sizeof(E2), cudaMemcpyHostToDevice, stream));
    // FRI shape (fri/mod.rs:80-94) and which layers stay split by coset: layer r (domain 2^(log_lde - la*r)) is
    // produced rank-locally while fri_layer_sharded() holds; the first small layer is stored into every rank and
    // everything after it is replicated.  The debug export of the DEEP evaluations needs layer 0 everywhere.
    const u32 la = params.log_folding_arity;
    u32 rounds; size_t final_deg;
    {
        u32 target = params.log_final_degree + lb;
        u32 steps = log_lde > target ? log_lde - target : 0;
        rounds = (steps + la - 1) / la;
        u32 lf = log_lde > la * rounds ? log_lde - la * rounds : 0;
        final_deg = (size_t)1 << (lf > lb ? lf - lb : 0);
    }
    std::vector<char> layer_sh(rounds + 1, 0);
    for (u32 r = 0; r <= rounds && log_lde >= la * r; r++) {
        bool shd = fri_layer_sharded(log_lde - la * r) && r < rounds && !(r == 0 && keep_debug);
        layer_sh[r] = shd && (r == 0 || layer_sh[r - 1]);
        if (!layer_sh[r]) break;
    }
    std::vector<DevBuf> fri_layers;   // EF interleaved, natural domain order
    fri_layers.emplace_back(); fri_layers[0].alloc(2 * L, stream);
    {
        mk::DeepArgs da; da.n_mats = 0;
        std::vector<mk::DeepMat> all_mats;
        size_t mi = 0;
        for (int g = 0; g < ng; g++)
            for (size_t m = 0; m < groups[g]->mats.size(); m++, mi++) {
                CommittedMat& cm = groups[g]->mats[m];
                if (!cm.width) continue;
                all_mats.push_back(mk::DeepMat{cm.lde, cm.width, cm.log_n, aligned_off[mi], 0});
            }
        // descriptors travel through device memory, so the number of committed matrices is not limited
        DevBuf d_mats; d_mats.alloc(std::max<size_t>(1, all_mats.size() * sizeof(mk::DeepMat) / sizeof(u64)), stream);
        CUDA_OK(cudaMemcpyAsync(d_mats.p, all_mats.data(), all_mats.size() * sizeof(mk::DeepMat), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaStreamSynchronize(stream));   // all_mats is a stack vector
        da.m = (const mk::DeepMat*)d_mats.p; da.n_mats = (int)all_mats.size();
        // NB: the quotient matrix's device columns are already in committed order (2t + coord).
        da.log_n_max = log_max_n; da.log_blowup = lb; da.apow = d_apow.p; da.total_w = W;
        da.z0 = z; da.z1 = z_next; da.fz0 = fz[0]; da.fz1 = fz[1]; da.beta = dbeta;
        da.out = push_dst(fri_layers[0].p, sharded() && !layer_sh[0] ? mk::PUSH_ALL : mk::PUSH_LOCAL);
        da.T = &ntt(log_max_n).T; da.t0 = tb; da.nt = tn;
        if (sharded() && !layer_sh[0]) shard_barrier();   // fresh target buffer on every rank
        { ProfScope ps(prof, PC_DEEP); mk::launch_deep(da, stream); }
        if (sharded() && !layer_sh[0]) shard_barrier();
    }
    CUDA_OK(cudaStreamSynchronize(stream));
    if (keep_debug) {
        // export in the reference's bit-reversed order
        std::vector<u64> nat(2 * L);
        CUDA_OK(cudaMemcpy(nat.data(), fri_layers[0].p, 2 * L * sizeof(u64), cudaMemcpyDeviceToHost));
        dbg_deep.assign(2 * L, 0);
        for (size_t i = 0; i < L; i++) {
            size_t br = gl::bitrev32((u32)i, log_lde);
            dbg_deep[2 * br] = nat[2 * i]; dbg_deep[2 * br + 1] = nat[2 * i + 1];
        }
    }
    // 7d. FRI commit phase (fri/prover.rs:93-242)
    std::vector<Tree> fri_trees(rounds);
    std::vector<char> fri_split(rounds, 0);
    dbg_fri_roots.clear();
    u32 log_dom = log_lde;
    for (u32 r = 0; r < rounds; r++) {
        if (log_dom < la) fail(MDN_ERR_INVALID_ARG, "FRI domain too small for the folding arity");
        size_t q = (size_t)1 << (log_dom - la);
        Tree& t = fri_trees[r];
        t.depth = log_dom - la;
        t.nodes.alloc((2 * q - 1) * 4, stream);
        const bool in_sh = layer_sh[r] != 0;                       // this rank holds its cosets' entries of layer r only
        const bool split = in_sh && tree_sharded(t.depth);         // sub-tree per rank, else a replicated tree
        fri_split[r] = split;
        const u32 ft0 = in_sh ? tb : 0, fnt = in_sh ? tn : B;
        mk::PushDst dig = push_dst(t.layer(t.depth), in_sh ? (split ? mk::PUSH_OWNER : mk::PUSH_ALL) : mk::PUSH_LOCAL, t.depth - (split ? shard_log_g : 0));
        if (in_sh) shard_barrier();
        {
            ProfScope ps(prof, PC_FRI);
            perms += (q * (la == 3 ? 2 : 1) + q - 1) / (in_sh ? shard_world : 1);
            if (hash_kind == MDN_HASH_BLAKE3) mk::launch_fri_leaf_hash_b3(fri_layers[r].p, q, la, dig, lb, ft0, fnt, stream);
            else if (hash_kind == MDN_HASH_KECCAK) mk::launch_fri_leaf_hash_kk(fri_layers[r].p, q, la, dig, lb, ft0, fnt, stream);
            else mk::launch_fri_leaf_hash(fri_layers[r].p, q, la, dig, lb, ft0, fnt, stream, perm());
        }
        if (in_sh) shard_barrier();
        if (!split) {
            ProfScope ps(prof, PC_FRI);
            compress_subtree(t, t.depth, 0, 0);
        } else {
            const u32 lg = shard_log_g;
            {
                ProfScope ps(prof, PC_FRI);
                compress_subtree(t, t.depth, lg, shard_rank);
            }
            u64* mine = t.layer(lg) + (size_t)shard_rank * 4;
            mk::launch_push(mine, peers_of(mine), shard_rank, shard_world, 4, stream);
            shard_barrier();
            ProfScope ps(prof, PC_FRI);
            compress_subtree(t, lg, 0, 0);
        }
        u64 root[4];
        CUDA_OK(cudaMemcpyAsync(root, t.layer(0), sizeof root, cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        tr.send_commitment(root);
        dbg_fri_roots.insert(dbg_fri_roots.end(), root, root + 4);
        grind(params.folding_pow_bits);
        E2 fb = tr.ch.sample_ext();
        fri_layers.emplace_back(); fri_layers[r + 1].alloc(2 * q, stream);
        {
            const bool bcast = in_sh && !layer_sh[r + 1];          // the first replicated layer: stored into every rank
            mk::PushDst nxt = push_dst(fri_layers[r + 1].p, bcast ? mk::PUSH_ALL : mk::PUSH_LOCAL);
            if (bcast) shard_barrier();
            { ProfScope ps(prof, PC_FRI); mk::launch_fri_fold(fri_layers[r].p, log_dom, la, fb, nxt, lb, ft0, fnt, stream); }
            if (bcast) shard_barrier();
        }
        log_dom -= la;
    }
    shard_check_enqueue();        // evaluated after the synchronisation of the final-layer copy below
    // final polynomial (fri/prover.rs:228-239): values on the size-final_deg subgroup are the
    // final-layer entries at natural indices i*B; iDFT on the host, sent in descending order.
    {
        size_t dom = (size_t)1 << log_dom;
        std::vector<u64> lay(2 * dom);
        CUDA_OK(cudaMemcpyAsync(lay.data(), fri_layers[rounds].p, 2 * dom * sizeof(u64), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        shard_check_finish("FRI commit phase");
        size_t stride = dom / final_deg;
        u32 lf = 0; while (((size_t)1 << lf) < final_deg) lf++;
        u64 wi = gl::inv(gl::two_adic_generator(lf)), ninv = gl::inv((u64)final_deg);
        std::vector<E2> coeff(final_deg);
        for (size_t kk = 0; kk < final_deg; kk++) {
            u64 wk = gl::pow(wi, kk), x = 1;
            E2 a = gl::e2(0, 0);
            for (size_t i = 0; i < final_deg; i++) {
                E2 e = gl::e2(lay[2 * i * stride], lay[2 * i * stride + 1]);
                a = gl::e2_add(a, gl::e2_mulf(e, x));
                x = gl::mul(x, wk);
            }
            coeff[kk] = gl::e2_mulf(a, ninv);
        }
        for (size_t i = final_deg; i-- > 0;) tr.send_ext(coeff[i]);
    }
    // 7e. query grind + indices (pcs/prover.rs:73-84)
    grind(params.query_pow_bits);
    std::vector<size_t> qs;
    for (u32 i = 0; i < params.num_queries; i++) qs.push_back((size_t)tr.ch.sample_bits(log_lde));
    dbg_queries.assign(qs.begin(), qs.end());
    Indices ti = Indices::make(qs, log_lde);
    // 7f. openings: one pointer list, one gather (pcs/prover.rs:89-101; lifted_tree.rs:155-180).  Split over ranks,
    //     every opened word has an owner -- the rank holding that LDE coset / FRI coset / Merkle sub-tree -- which
    //     stores it into every rank's value buffer (-1: replicated data, read locally).
    std::vector<const u64*> ptrs;
    std::vector<int> owner;
    struct Emit { int kind; size_t count; size_t pad; };   // kind 0: `count` fields then `pad` zero fields; 1: commitment (4)
    std::vector<Emit> plan;
    for (int g = 0; g < ng; g++) {
        Committed& c = *groups[g];
        const bool replicated = (&c == &prep_c);   // the preprocessed bundle is committed once, whole, on every rank
        Indices leafs = ti.folded(c.tree.depth);
        for (size_t idx : leafs.idx)
            for (auto& cm : c.mats) {
                u32 ldm = cm.log_n + lb;
                size_t im = idx & (((size_t)1 << ldm) - 1);
                size_t t = im & (B - 1), rr = im >> lb;
                size_t pos = (t << cm.log_n) + rr, Lm = (size_t)1 << ldm;
                for (u32 col = 0; col < cm.width; col++) { ptrs.push_back(cm.lde + (size_t)col * Lm + pos); owner.push_back(replicated ? -1 : coset_owner((u32)t)); }
                plan.push_back(Emit{0, cm.width, (size_t)((cm.width + align() - 1) / align() * align() - cm.width)});
            }
        for (auto& ds : hostfs::missing_siblings(leafs)) {
            // split tree: a node below the sub-root level exists only on the rank owning its leaf range
            int own = -1;
            if (!replicated && tree_sharded(c.tree.depth) && ds.first > shard_log_g) own = (int)(ds.second >> (ds.first - shard_log_g));
            for (int q = 0; q < 4; q++) { ptrs.push_back(c.tree.layer(ds.first) + ds.second * 4 + q); owner.push_back(own); }
            plan.push_back(Emit{1, 4, 0});
        }
    }
    {
        Indices fi = ti;
        u32 ld = log_lde;
        for (u32 r = 0; r < rounds; r++) {
            fi = fi.folded(fi.depth > la ? fi.depth - la : 0);
            size_t q = (size_t)1 << (ld - la);
            const u64* lay = fri_layers[r].p;
            u32 a = 1u << la;
            for (size_t idx : fi.idx) {
                int own = layer_sh[r] ? coset_owner((u32)(idx & (B - 1))) : -1;    // the row's 2^la entries share idx's coset
                for (u32 e = 0; e < a; e++) {
                    size_t src = idx + (size_t)gl::bitrev32(e, la) * q;
                    ptrs.push_back(lay + 2 * src); ptrs.push_back(lay + 2 * src + 1);
                    owner.push_back(own); owner.push_back(own);
                }
                plan.push_back(Emit{0, 2 * (size_t)a, 0});
            }
            for (auto& ds : hostfs::missing_siblings(fi)) {
                int own = -1;
                if (fri_split[r] && ds.first > shard_log_g) own = (int)(ds.second >> (ds.first - shard_log_g));
                for (int qq = 0; qq < 4; qq++) { ptrs.push_back(fri_trees[r].layer(ds.first) + ds.second * 4 + qq); owner.push_back(own); }
                plan.push_back(Emit{1, 4, 0});
            }
            ld -= la;
        }
    }
    {
        DevBuf d_ptrs, d_vals, d_owner; d_ptrs.alloc(ptrs.size(), stream); d_vals.alloc(ptrs.size(), stream);
        CUDA_OK(cudaMemcpyAsync(d_ptrs.p, ptrs.data(), ptrs.size() * sizeof(u64), cudaMemcpyHostToDevice, stream));
        if (!sharded()) {
            ProfScope ps(prof, PC_GATHER); mk::launch_gather((const u64* const*)d_ptrs.p, d_vals.p, ptrs.size(), stream);
        } else {
            d_owner.alloc((owner.size() + 1) / 2, stream);
            CUDA_OK(cudaMemcpyAsync(d_owner.p, owner.data(), owner.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
            shard_barrier();
            { ProfScope ps(prof, PC_GATHER); mk::launch_gather_push((const u64* const*)d_ptrs.p, (const int*)d_owner.p, peers_of(d_vals.p), shard_rank, shard_world, ptrs.size(), stream); }
            shard_barrier();
        }
        std::vector<u64> vals(ptrs.size());
        CUDA_OK(cudaMemcpyAsync(vals.data(), d_vals.p, vals.size() * sizeof(u64), cudaMemcpyDeviceToHost, stream));
        shard_check_enqueue();
        CUDA_OK(cudaStreamSynchronize(stream));
        shard_check_finish("query openings");
        size_t o = 0;
        for (auto& e : plan) {
            if (e.kind == 0) { for (size_t i = 0; i < e.count; i++) tr.hint_field(vals[o++]); for (size_t i = 0; i < e.pad; i++) tr.hint_field(0); }
            else { tr.hint_commitment(&vals[o]); o += 4; }
        }
    }
    CUDA_OK(cudaEventRecord(ev[7], stream));
    CUDA_OK(cudaEventSynchronize(ev[7]));
    // 8. StarkProofData (mod.rs:572-577)
    out_heights.clear();
    for (u32 h : log_heights) out_heights.push_back((uint8_t)h);
    out_fields = std::move(tr.fields);
    out_commitments = std::move(tr.commitments);
    cudaEventElapsedTime(&timings.h2d_transpose, ev[0], ev[1]);
    cudaEventElapsedTime(&timings.commit_main, ev[1], ev[2]);
    cudaEventElapsedTime(&timings.commit_aux, ev[3], ev[4]);
    cudaEventElapsedTime(&timings.evaluate_constraints, ev[4], ev[5]);
    cudaEventElapsedTime(&timings.commit_quotient, ev[5], ev[6]);
    cudaEventElapsedTime(&timings.open, ev[6], ev[7]);
    cudaEventElapsedTime(&timings.total, ev[0], ev[7]);
    timings.kernel_launches = mk::launch_count();
    prof.resolve(timings.kernel_ms, timings.kernel_regions);
    timings.leaf_hash_bytes = leaf_bytes; timings.ntt_bytes = ntt_bytes; timings.permutations = perms;
    in_proof = false; shard_active = false;   // the API wrapper returns the proof's device memory to the arena once this frame is gone
}

// =============================================================================================
// C ABI
// =============================================================================================
#define API_TRY(s) try {
#define API_CATCH(s) } catch (const MdnError& e) { (s)->error = e.what(); (s)->reset_proof(); return e.code; } \
    catch (const std::exception& e) { (s)->error = e.what(); (s)->reset_proof(); return MDN_ERR_INVALID_ARG; } return MDN_OK;

extern "C" {

int mdn_session_create(const mdn_pcs_params* params, int cuda_device, mdn_session** out) {
    if (!params || !out) { g_create_error = "null argument"; return MDN_ERR_INVALID_ARG; }
    // PcsParams::new (pcs/params.rs:53-99), before any device is touched
    if (params->log_folding_arity < 1 || params->log_folding_arity > 3) { g_create_error = "invalid folding arity: log_arity " + std::to_string(params->log_folding_arity) + " (must be 1, 2, or 3)"; return MDN_ERR_INVALID_ARG; }
    if (params->log_blowup == 0) { g_create_error = "log_blowup must be at least 1"; return MDN_ERR_INVALID_ARG; }
    if (params->num_queries == 0) { g_create_error = "num_queries must be at least 1"; return MDN_ERR_INVALID_ARG; }
    if (params->log_final_degree + params->log_blowup < params->log_folding_arity - 1) {
        g_create_error = "log_final_degree " + std::to_string(params->log_final_degree) + " + log_blowup " + std::to_string(params->log_blowup) + " is below the minimum target " +
                         std::to_string(params->log_folding_arity - 1) + " reachable by fixed-arity folding";
        return MDN_ERR_INVALID_ARG;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no usable CUDA device: ") + cudaGetErrorString(e) + " (this backend has no CPU fallback)";
        return MDN_ERR_NO_DEVICE;
    }
    if (cuda_device < 0 || cuda_device >= count) { g_create_error = "CUDA device index out of range"; return MDN_ERR_NO_DEVICE; }
    if (params->num_queries == 0 || params->log_blowup == 0) { g_create_error = "invalid PCS parameters"; return MDN_ERR_INVALID_ARG; }
    auto* s = new mdn_session();
    s->params = *params; s->device = cuda_device;
    try {
        CUDA_OK(cudaSetDevice(cuda_device));
        CUDA_OK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        for (auto& evn : s->copy_ev) CUDA_OK(cudaEventCreateWithFlags(&evn, cudaEventDisableTiming));
        for (auto& evn : s->ev) CUDA_OK(cudaEventCreate(&evn));
        cudaMemPool_t pool;
        CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, cuda_device));
        uint64_t thr = ~0ull;
        CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        mk::upload_constants();
        CUDA_OK(cudaDeviceSynchronize());
    } catch (const std::exception& ex) { g_create_error = ex.what(); delete s; return MDN_ERR_CUDA; }
    *out = s;
    return MDN_OK;
}

void mdn_session_destroy(mdn_session* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    // every stream-ordered allocation must be returned before the stream goes away
    s->reset_proof();
    s->allgather = nullptr;     // the peers may be gone already: unmap without a rendezvous
    s->shard_teardown();
    s->prep_c = Committed();
    s->d_publics.release(); s->d_randomness.release(); s->d_aux_values.release(); s->d_flag.release();
    s->ntt_plans.clear(); s->premul_plans.clear();
    for (auto& a : s->airs) a.jit.reset();
    s->jit_kernels.clear();
    cudaStreamSynchronize(s->stream);
    s->arena.destroy();
    for (auto& evn : s->ev) cudaEventDestroy(evn);
    for (auto& evn : s->copy_ev) cudaEventDestroy(evn);
    for (int b = 0; b < 2; b++) if (s->bounce[b]) { cudaFreeHost(s->bounce[b]); cudaEventDestroy(s->bounce_ev[b]); }
    cudaStreamDestroy(s->copy_stream);
    cudaStreamDestroy(s->stream);
    delete s;
}

const char* mdn_last_error(const mdn_session* s) { return s ? s->error.c_str() : g_create_error.c_str(); }

static void fill_proof(mdn_session* s, mdn_proof* out) {
    out->log_trace_heights = s->out_heights.data(); out->n_heights = s->out_heights.size();
    out->fields = s->out_fields.data(); out->n_fields = s->out_fields.size();
    out->commitments = s->out_commitments.data(); out->n_commitments = s->out_commitments.size() / 4;
}

int mdn_prove_begin(mdn_session* s, const mdn_statement* st, const mdn_matrix* traces, const mdn_challenger* challenger,
                    uint32_t flags, uint64_t main_root[4], uint64_t* randomness_out) {
    if (!s) return MDN_ERR_INVALID_ARG;
    API_TRY(s)
    ArenaScope proof_memory(s->use_arena ? &s->arena : nullptr);
    CUDA_OK(cudaSetDevice(s->device));
    s->prove_begin(st, traces, challenger, flags);
    if (main_root) memcpy(main_root, s->main_c.root, 32);
    if (randomness_out) for (size_t i = 0; i < s->randomness.size(); i++) { randomness_out[2 * i] = s->randomness[i].a; randomness_out[2 * i + 1] = s->randomness[i].b; }
    API_CATCH(s)
}

int mdn_prove_commit_aux(mdn_session* s, const mdn_matrix* aux, const uint64_t* const* aux_values, uint64_t aux_root[4]) {
    if (!s) return MDN_ERR_INVALID_ARG;
    API_TRY(s)
    ArenaScope proof_memory(s->use_arena ? &s->arena : nullptr);
    CUDA_OK(cudaSetDevice(s->device));
    s->commit_aux(aux, aux_values, aux == nullptr);
    if (aux_root) memcpy(aux_root, s->aux_c.root, 32);
    API_CATCH(s)
}

int mdn_prove_finish(mdn_session* s, mdn_proof* out) {
    if (!s || !out) return MDN_ERR_INVALID_ARG;
    API_TRY(s)
    ArenaScope proof_memory(s->use_arena ? &s->arena : nullptr);
    CUDA_OK(cudaSetDevice(s->device));
    s->finish();
    fill_proof(s, out);
    s->release_proof_memory();
    API_CATCH(s)
}

int mdn_prove(mdn_session* s, const mdn_statement* st, const mdn_matrix* traces, const mdn_challenger* challenger,
              mdn_aux_builder build_aux, void* aux_ctx, uint32_t flags, mdn_proof* out) {
    if (!s || !out) return MDN_ERR_INVALID_ARG;
    API_TRY(s)
    ArenaScope proof_memory(s->use_arena ? &s->arena : nullptr);
    CUDA_OK(cudaSetDevice(s->device));
    s->prove_begin(st, traces, challenger, flags);
    if (!build_aux) {
        s->commit_aux(nullptr, nullptr, true);
    } else {
        if (flags & MDN_FLAG_DEVICE_TRACES) fail(MDN_ERR_UNSUPPORTED, "an aux builder needs host-resident main traces");
        u32 k = st->n_airs;
        std::vector<std::vector<u64>> aux_bufs(k), val_bufs(k);
        std::vector<mdn_matrix> aux_mats(k);
        std::vector<const u64*> val_ptrs(k);
        for (u32 i = 0; i < k; i++) {
            const mdn_air& a = st->airs[i];
            if (a.lookup) { aux_mats[i] = mdn_matrix{nullptr, traces[i].log_height, 2 * a.aux_width}; val_ptrs[i] = nullptr; continue; }   // built on the device
            size_t N = (size_t)1 << traces[i].log_height;
            aux_bufs[i].assign(N * 2 * a.aux_width, 0);
            val_bufs[i].assign(2 * (size_t)a.num_aux_values + 1, 0);
            std::vector<u64> r;
            for (u32 q = 0; q < a.num_randomness; q++) { r.push_back(s->randomness[q].a); r.push_back(s->randomness[q].b); }
            r.push_back(0);
            if (build_aux(aux_ctx, i, &traces[i], r.data(), aux_bufs[i].data(), val_bufs[i].data()) != 0)
                fail(MDN_ERR_AUX_BUILDER, "aux builder failed for instance %u", i);
            aux_mats[i] = mdn_matrix{aux_bufs[i].data(), traces[i].log_height, 2 * a.aux_width};
            val_ptrs[i] = val_bufs[i].data();
        }
        s->commit_aux(aux_mats.data(), val_ptrs.data(), false);
    }
    s->finish();
    fill_proof(s, out);
    s->release_proof_memory();
    API_CATCH(s)
}

size_t mdn_proof_serialize(const mdn_proof* p, uint8_t* out, size_t cap) {
    size_t need = 8 + p->n_heights + 8 + 8 * p->n_fields + 8 + 32 * p->n_commitments;
    if (!out || cap < need) return need;
    auto put64 = [&](uint64_t v) { for (int i = 0; i < 8; i++) *out++ = (uint8_t)(v >> (8 * i)); };
    put64(p->n_heights); memcpy(out, p->log_trace_heights, p->n_heights); out += p->n_heights;
    put64(p->n_fields); for (size_t i = 0; i < p->n_fields; i++) put64(p->fields[i]);
    put64(p->n_commitments); for (size_t i = 0; i < 4 * p->n_commitments; i++) put64(p->commitments[i]);
    return need;
}

int mdn_coset_lde_batch(mdn_session* s, const mdn_matrix* mat, uint32_t added_bits, uint64_t shift, uint64_t* out) {
    if (!s || !mat || !out) return MDN_ERR_INVALID_ARG;
    API_TRY(s)
    CUDA_OK(cudaSetDevice(s->device));
    if (added_bits != s->params.log_blowup) fail(MDN_ERR_UNSUPPORTED, "added_bits must equal the session's log_blowup");
    if (mat->log_height > 22) fail(MDN_ERR_UNSUPPORTED, "trace height 2^%u exceeds the supported 2^22", mat->log_height);
    if (mat->log_height + added_bits > 32) fail(MDN_ERR_DOMAIN, "LDE log order %u exceeds two-adicity 32", mat->log_height + added_bits);
    if (s->in_proof) fail(MDN_ERR_INVALID_ARG, "mdn_coset_lde_batch called inside a proof");
    if (shift != gl::lde_shift(mat->log_height + added_bits)) fail(MDN_ERR_UNSUPPORTED, "only the canonical LDE shift 7^(2^(32-log_lde)) is supported");
    if (!s->d_flag.p) { s->d_flag.alloc(1, s->stream); CUDA_OK(cudaMemsetAsync(s->d_flag.p, 0, 8, s->stream)); }
    Committed c;
    size_t N = (size_t)1 << mat->log_height, L = N << added_bits;
    c.coef_buf.alloc(N * mat->width, s->stream); c.lde_buf.alloc(L * mat->width, s->stream);
    c.mats.push_back(CommittedMat{c.lde_buf.p, c.coef_buf.p, mat->log_height, mat->width});
    s->upload_matrix(*mat, false, c.coef_buf.p);
    s->check_input_flag("the matrix");
    s->lde_matrix(c.mats[0]);          // the LDE only: no tree
    DevBuf rm; rm.alloc(L * mat->width, s->stream);
    mk::launch_export_lde_bitrev_rm(c.lde_buf.p, mat->log_height, added_bits, mat->width, rm.p, s->stream);
    CUDA_OK(cudaMemcpyAsync(out, rm.p, L * mat->width * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    API_CATCH(s)
}

// The matrices are trace-domain evaluations (height N); the committed tree is the one
// `commit_traces` builds: LDE by the session blowup then build_aligned_tree.
int mdn_lmcs_commit(mdn_session* s, const mdn_matrix* mats, uint32_t n_mats, uint64_t root[4]) {
    if (!s || !mats || !root || !n_mats) return MDN_ERR_INVALID_ARG;
    API_TRY(s)
    CUDA_OK(cudaSetDevice(s->device));
    u32 lb = s->params.log_blowup;
    if (!s->d_flag.p) { s->d_flag.alloc(1, s->stream); CUDA_OK(cudaMemsetAsync(s->d_flag.p, 0, 8, s->stream)); }
    Committed c;
    size_t ct = 0, lt = 0;
    for (u32 i = 0; i < n_mats; i++) {
        if (i && mats[i].log_height < mats[i - 1].log_height) fail(MDN_ERR_INVALID_ARG, "matrices must be sorted by ascending height");
        size_t N = (size_t)1 << mats[i].log_height; ct += N * mats[i].width; lt += (N << lb) * mats[i].width;
    }
    c.coef_buf.alloc(ct, s->stream); c.lde_buf.alloc(lt, s->stream);
    size_t co = 0, lo = 0;
    for (u32 i = 0; i < n_mats; i++) {
        size_t N = (size_t)1 << mats[i].log_height;
        c.mats.push_back(CommittedMat{c.lde_buf.p + lo, c.coef_buf.p + co, mats[i].log_height, mats[i].width});
        s->upload_matrix(mats[i], false, c.coef_buf.p + co);
        co += N * mats[i].width; lo += (N << lb) * mats[i].width;
    }
    s->check_input_flag("a matrix");
    s->lde_and_commit(c, nullptr, nullptr);
    memcpy(root, c.root, 32);
    API_CATCH(s)
}

int mdn_poseidon2_permute(mdn_session* s, uint64_t* states, size_t n) {
    if (!s || !states) return MDN_ERR_INVALID_ARG;
    API_TRY(s)
    CUDA_OK(cudaSetDevice(s->device));
    DevBuf d; d.alloc(12 * n, s->stream);
    CUDA_OK(cudaMemcpyAsync(d.p, states, 12 * n * sizeof(u64), cudaMemcpyHostToDevice, s->stream));
    mk::launch_poseidon2_batch(d.p, n, s->stream);
    CUDA_OK(cudaMemcpyAsync(states, d.p, 12 * n * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    API_CATCH(s)
}

void mdn_challenger_observe(mdn_challenger* c, const uint64_t* felts, size_t n) {
    Duplex d;
    memcpy(d.st, c->sponge_state, sizeof d.st);
    memcpy(d.in, c->input_buffer, sizeof d.in);
    d.in_len = c->input_len; d.out_len = c->output_len;
    for (size_t i = 0; i < n; i++) d.observe(felts[i]);
    memcpy(c->sponge_state, d.st, sizeof d.st);
    for (u32 i = 0; i < 8; i++) c->input_buffer[i] = i < d.in_len ? d.in[i] : 0;
    c->input_len = d.in_len; c->output_len = d.out_len;
}
uint64_t mdn_challenger_sample(mdn_challenger* c) {
    Duplex d;
    memcpy(d.st, c->sponge_state, sizeof d.st);
    memcpy(d.in, c->input_buffer, sizeof d.in);
    d.in_len = c->input_len; d.out_len = c->output_len;
    u64 v = d.sample();
    memcpy(c->sponge_state, d.st, sizeof d.st);
    for (u32 i = 0; i < 8; i++) c->input_buffer[i] = i < d.in_len ? d.in[i] : 0;
    c->input_len = d.in_len; c->output_len = d.out_len;
    return v;
}

long long mdn_get_info(mdn_session* s, mdn_info what, uint64_t* out, size_t cap) {
    if (!s) return -1;
    std::vector<u64> v;
    switch (what) {
        case MDN_INFO_MAIN_ROOT: v.assign(s->dbg_roots[0], s->dbg_roots[0] + 4); break;
        case MDN_INFO_AUX_ROOT: v.assign(s->dbg_roots[1], s->dbg_roots[1] + 4); break;
        case MDN_INFO_QUOTIENT_ROOT: v.assign(s->dbg_roots[2], s->dbg_roots[2] + 4); break;
        case MDN_INFO_OOD_POINT: v = {s->ood_z.a, s->ood_z.b}; break;
        case MDN_INFO_QUOTIENT_ACC: v = s->dbg_quot_acc; break;
        case MDN_INFO_DEEP_EVALS: v = s->dbg_deep; break;
        case MDN_INFO_FRI_ROOTS: v = s->dbg_fri_roots; break;
        case MDN_INFO_QUERY_INDICES: v = s->dbg_queries; break;
        case MDN_INFO_JIT: v = s->jit_used; break;
        case MDN_INFO_POOL: {
            cudaMemPool_t pool; uint64_t a[4] = {0, 0, 0, 0};
            if (cudaDeviceGetDefaultMemPool(&pool, s->device) == cudaSuccess) {
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &a[0]);
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemHigh, &a[1]);
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &a[2]);
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &a[3]);
            }
            v.assign(a, a + 4);
            { size_t cap = 0; for (auto& sl : s->arena.slabs) cap += sl.size; v.push_back(cap); v.push_back(s->arena.slabs.size()); v.push_back(s->arena.grow_events); v.push_back(s->arena.live); }
            break;
        }
        case MDN_INFO_BUILD: {
#ifdef MDN_GEN1
            v.push_back(1); v.push_back(1);
#else
            v.push_back(2); v.push_back(2);
#endif
            break;
        }
        default: return -1;
    }
    if (out) for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
    return (long long)v.size();
}

int mdn_session_set_shard(mdn_session* s, uint32_t rank, uint32_t world, mdn_allgather_fn fn, void* ctx) {
    if (!s) return MDN_ERR_INVALID_ARG;
    if (world == 0 || (world & (world - 1)) || world > mk::MAX_RANKS || rank >= world || (world > 1 && !fn)) {
        s->error = "invalid shard configuration (world must be a power of two <= 8, callback required)"; return MDN_ERR_INVALID_ARG;
    }
    if (s->in_proof) { s->error = "mdn_session_set_shard called inside a proof"; return MDN_ERR_INVALID_ARG; }
    try {
        CUDA_OK(cudaSetDevice(s->device));
        s->shard_teardown();
        if (world == 1) return MDN_OK;
        s->shard_rank = rank; s->shard_world = world; s->allgather = fn; s->allgather_ctx = ctx;
        s->shard_log_g = 0; while ((1u << s->shard_log_g) < world) s->shard_log_g++;
        if (const char* e = getenv("MDN_SHARD_MIN_LOG")) s->shard_min_log = (u32)atoi(e);
        // the proof arena starts over as a shared arena: every slab is mapped into every rank when it is created
        s->arena.destroy();
        s->arena.on_new_slab = [s](char* base, size_t size) { s->shard_map_slab(base, size); };
        s->arena.before_drop_slabs = [s]() { s->shard_unmap_slabs(); };
        // barrier flags: one slot per source rank, zeroed before any peer can see the buffer
        CUDA_OK(cudaMalloc((void**)&s->sync_local, 4096));
        CUDA_OK(cudaMemsetAsync(s->sync_local, 0, 4096, s->stream));
        CUDA_OK(cudaStreamSynchronize(s->stream));
        cudaIpcMemHandle_t h;
        CUDA_OK(cudaIpcGetMemHandle(&h, s->sync_local));
        std::vector<u64> mine(8), all(8 * (size_t)world);
        memcpy(mine.data(), &h, 64);
        if (fn(ctx, mine.data(), all.data(), 8) != 0) fail(MDN_ERR_INVALID_ARG, "all-gather callback failed");
        for (u32 g = 0; g < world; g++) {
            if (g == rank) { s->sync_flags.p[g] = s->sync_local; continue; }
            cudaIpcMemHandle_t hg; memcpy(&hg, &all[8 * (size_t)g], 64);
            void* m = nullptr;
            CUDA_OK(cudaIpcOpenMemHandle(&m, hg, cudaIpcMemLazyEnablePeerAccess));
            s->sync_flags.p[g] = (u64*)m;
        }
        s->sync_epoch = 0;
        // rendezvous: no rank may go on (and possibly fail and free its buffer) before every rank has mapped it
        { std::vector<u64> one(1, 0), got(world); if (fn(ctx, one.data(), got.data(), 1) != 0) fail(MDN_ERR_INVALID_ARG, "all-gather callback failed"); }
    } catch (const MdnError& e) { s->error = e.what(); return e.code; }
    catch (const std::exception& e) { s->error = e.what(); return MDN_ERR_INVALID_ARG; }
    return MDN_OK;
}

int mdn_session_set_hash(mdn_session* s, mdn_hash_kind kind) {
    if (!s) return MDN_ERR_INVALID_ARG;
    if (kind < MDN_HASH_POSEIDON2 || kind > MDN_HASH_RPX) { s->error = "unknown hash configuration"; return MDN_ERR_UNSUPPORTED; }
    if (s->in_proof) { s->error = "mdn_session_set_hash called inside a proof"; return MDN_ERR_INVALID_ARG; }
    if (s->has_prep && kind != s->hash_kind) { s->error = "the preprocessed bundle was committed under the other hash: remove it first"; return MDN_ERR_INVALID_ARG; }
    s->hash_kind = kind;
    return MDN_OK;
}
int mdn_session_set_hash_challenger(mdn_session* s, const mdn_hash_challenger* c) {
    if (!s || !c || (c->input_len && !c->input_buffer) || (c->output_len && !c->output_buffer)) return MDN_ERR_INVALID_ARG;
    if (c->input_len % 4 || c->output_len > 32) { s->error = "hash challenger: the input buffer must be whole 32-bit words and the output buffer at most 32 bytes"; return MDN_ERR_INVALID_ARG; }
    s->hash_ch_in.assign(c->input_buffer, c->input_buffer + c->input_len);
    s->hash_ch_out.assign(c->output_buffer, c->output_buffer + c->output_len);
    return MDN_OK;
}

int mdn_session_set_external_check(mdn_session* s, mdn_external_check fn, void* ctx) {
    if (!s) return MDN_ERR_INVALID_ARG;
    s->external_check = fn; s->external_ctx = ctx;
    return MDN_OK;
}

int mdn_session_set_preprocessed(mdn_session* s, const mdn_statement* st, const mdn_matrix* preprocessed, uint64_t commitment_out[4]) {
    if (!s) return MDN_ERR_INVALID_ARG;
    try {
        CUDA_OK(cudaSetDevice(s->device));
        s->set_preprocessed(st, preprocessed);
        if (commitment_out) { if (s->has_prep) memcpy(commitment_out, s->prep_c.root, 32); else memset(commitment_out, 0, 32); }
    } catch (const MdnError& e) { s->error = e.what(); s->prep_c = Committed(); s->has_prep = false; return e.code; }
    catch (const std::exception& e) { s->error = e.what(); s->prep_c = Committed(); s->has_prep = false; return MDN_ERR_INVALID_ARG; }
    return MDN_OK;
}

const char* mdn_jit_status(mdn_session* s) {
    static thread_local std::string msg;
    jit::Nvrtc& n = jit::nvrtc();
    msg = n.ok() ? "nvrtc " + std::to_string(n.version / 100) + "." + std::to_string(n.version % 100) : "nvrtc unavailable (" + n.why + ")";
    if (s && !s->jit_note.empty()) msg += "; " + s->jit_note;
    return msg.c_str();
}

int mdn_session_set_jit(mdn_session* s, uint32_t min_nodes) {
    if (!s) return MDN_ERR_INVALID_ARG;
    s->jit_min_nodes = min_nodes;
    return MDN_OK;
}

// Codegen + NVRTC only (no device needed): returns the cubin size, or a negative status with the compiler log
// in *err (static buffer).  Lets CPU-only CI check that an AIR lowers and compiles.
long long mdn_jit_compile_check(const uint32_t* program, uint32_t program_words, const char** err) {
    static thread_local std::string msg;
    try {
        const bool lookup = program && program_words >= 5 && program[0] == 0x504B4C4Du;
        if (!program || program_words < 5 || (program[0] != 0x5249414Du && !lookup) || program[1] != 1 ||
            (size_t)program_words != 5 + 3 * (size_t)program[2] + (lookup ? 4 : 1) * (size_t)program[3] + 2 * (size_t)program[4]) { msg = "bad program"; if (err) *err = msg.c_str(); return MDN_ERR_INVALID_ARG; }
        for (u32 j = 0; j < program[2]; j++) {
            u32 op = program[5 + 3 * j], x = program[6 + 3 * j], y = program[7 + 3 * j];
            if (op > 15 || (op >= 10 && op <= 12 && (x >= j || y >= j)) || (op == 13 && x >= j) || ((op == 8) && x >= program[4]) || (op == 9 && x + 1 >= program[4])) { msg = "malformed node"; if (err) *err = msg.c_str(); return MDN_ERR_INVALID_ARG; }
        }
        if (lookup) {
            u32 n_cols = 0;
            for (u32 q = 0; q < program[3]; q++) {
                const u32* it = program + 5 + 3 * (size_t)program[2] + 4 * (size_t)q;
                if ((it[1] != 0xFFFFFFFFu && it[1] >= program[2]) || it[2] >= program[2] || it[3] >= program[2] || it[0] >= 16) { msg = "bad interaction"; if (err) *err = msg.c_str(); return MDN_ERR_INVALID_ARG; }
                n_cols = std::max(n_cols, it[0] + 1);
            }
            return (long long)jit::cubin_for(program, program_words, nullptr, true, n_cols).size();
        }
        for (u32 q = 0; q < program[3]; q++) if (program[5 + 3 * (size_t)program[2] + q] >= program[2]) { msg = "bad constraint id"; if (err) *err = msg.c_str(); return MDN_ERR_INVALID_ARG; }
        return (long long)jit::cubin_for(program, program_words, nullptr).size();
    } catch (const std::exception& e) { msg = e.what(); if (err) *err = msg.c_str(); return MDN_ERR_UNSUPPORTED; }
}

// sizeof / offsetof of every struct of the boundary, so a binding (ctypes, Rust #[repr(C)]) can assert its layout
size_t mdn_abi_layout(uint32_t* out, size_t cap) {
    const uint32_t v[] = {
        (uint32_t)sizeof(mdn_pcs_params), (uint32_t)sizeof(mdn_challenger), (uint32_t)sizeof(mdn_lookup), (uint32_t)sizeof(mdn_air),
        (uint32_t)offsetof(mdn_air, program), (uint32_t)offsetof(mdn_air, periodic_values), (uint32_t)offsetof(mdn_air, preprocessed_width),
        (uint32_t)offsetof(mdn_air, lookup), (uint32_t)sizeof(mdn_matrix), (uint32_t)sizeof(mdn_statement), (uint32_t)sizeof(mdn_proof),
        (uint32_t)sizeof(mdn_timings), (uint32_t)offsetof(mdn_timings, kernel_ms), (uint32_t)offsetof(mdn_timings, permutations)};
    size_t n = sizeof v / sizeof v[0];
    if (out) for (size_t i = 0; i < n && i < cap; i++) out[i] = v[i];
    return n;
}

int mdn_set_debug(mdn_session* s, int enable) {
    if (!s) return MDN_ERR_INVALID_ARG;
    s->keep_debug = enable != 0;
    return MDN_OK;
}

int mdn_get_timings(mdn_session* s, mdn_timings* out) {
    if (!s || !out) return MDN_ERR_INVALID_ARG;
    *out = s->timings;
    return MDN_OK;
}

}  // extern "C"
