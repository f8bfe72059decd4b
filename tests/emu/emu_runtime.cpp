// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory).
// CPU execution of CUDA kernels: every CUDA thread of a block is a cooperative fiber (ucontext); __syncthreads()
// and warp shuffles hand control back to the block scheduler, which resumes the fibers warp by warp.  Blocks run
// one after another, "device memory" is host memory, streams are synchronous, events carry wall-clock time.
#include "cuda_runtime.h"
#include <ucontext.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace emu {
uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;
size_t g_max_dyn_smem_opt_in = 0;

enum State { RUNNABLE, AT_WARP_BAR, AT_BLOCK_BAR, DONE };
struct Fiber { ucontext_t ctx; State st; uint3 tid; char* stack; };
static constexpr size_t STACK_BYTES = 256 << 10;
static std::vector<Fiber> g_fibers;
static std::vector<char*> g_stacks;
static ucontext_t g_sched;
static Fiber* g_cur = nullptr;
static const std::function<void()>* g_body = nullptr;
static unsigned long long g_warp_slots[32];
static std::vector<unsigned char> g_smem_buf;

static void yield(State s) { Fiber* f = g_cur; f->st = s; swapcontext(&f->ctx, &g_sched); }
void syncthreads() { yield(AT_BLOCK_BAR); }
unsigned long long shfl_down(unsigned long long v, unsigned off) {
    unsigned lane = (g_cur->tid.x + g_cur->tid.y * g_blockDim.x) & 31u;
    g_warp_slots[lane] = v;
    yield(AT_WARP_BAR);                       // every lane of the warp has published its value
    unsigned src = lane + off;
    unsigned long long r = src < 32 ? g_warp_slots[src] : v;
    yield(AT_WARP_BAR);                       // every lane has read before the slots are reused
    return r;
}
static void fiber_main() {
    (*g_body)();
    g_cur->st = DONE;
    swapcontext(&g_cur->ctx, &g_sched);
}
static void resume(Fiber& f) {
    g_cur = &f;
    g_threadIdx = f.tid;
    f.st = RUNNABLE;
    swapcontext(&g_sched, &f.ctx);
}

void launch(const LaunchCfg& c, const std::function<void()>& body) {
    unsigned nthreads = c.block.x * c.block.y * c.block.z;
    if (nthreads == 0 || nthreads > 1024) { fprintf(stderr, "emu: bad block size %u\n", nthreads); abort(); }
    if (c.smem > (227u << 10)) { fprintf(stderr, "emu: %zu bytes of dynamic shared memory exceed the 227 KB an sm_100a block can have\n", c.smem); abort(); }
    if (c.smem > (48u << 10) && c.smem > g_max_dyn_smem_opt_in) { fprintf(stderr, "emu: %zu bytes of dynamic shared memory without cudaFuncAttributeMaxDynamicSharedMemorySize\n", c.smem); abort(); }
    while (g_stacks.size() < nthreads) g_stacks.push_back((char*)malloc(STACK_BYTES));
    g_fibers.resize(nthreads);
    g_smem_buf.assign(c.smem + 16, 0xCD);     // poison: reads of unwritten shared memory show up as garbage
    g_dyn_smem = g_smem_buf.data();
    g_blockDim = c.block; g_gridDim = c.grid;
    g_body = &body;
    for (unsigned bz = 0; bz < c.grid.z; bz++)
    for (unsigned by = 0; by < c.grid.y; by++)
    for (unsigned bx = 0; bx < c.grid.x; bx++) {
        g_blockIdx = uint3{bx, by, bz};
        for (unsigned t = 0; t < nthreads; t++) {
            Fiber& f = g_fibers[t];
            f.tid = uint3{t % c.block.x, (t / c.block.x) % c.block.y, t / (c.block.x * c.block.y)};
            f.st = RUNNABLE; f.stack = g_stacks[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, fiber_main, 0);
        }
        unsigned nwarps = (nthreads + 31) / 32;
        for (;;) {
            bool all_done = true;
            for (unsigned w = 0; w < nwarps; w++) {
                unsigned lo = w * 32, hi = lo + 32 < nthreads ? lo + 32 : nthreads;
                for (;;) {
                    for (unsigned t = lo; t < hi; t++) if (g_fibers[t].st == RUNNABLE) resume(g_fibers[t]);
                    bool any_warp_bar = false, all_warp_bar = true;
                    for (unsigned t = lo; t < hi; t++) {
                        if (g_fibers[t].st == AT_WARP_BAR) any_warp_bar = true;
                        else if (g_fibers[t].st != DONE) all_warp_bar = false;
                    }
                    if (!any_warp_bar) break;
                    if (!all_warp_bar) { fprintf(stderr, "emu: divergent warp shuffle in block (%u,%u)\n", bx, by); abort(); }
                    for (unsigned t = lo; t < hi; t++) if (g_fibers[t].st == AT_WARP_BAR) g_fibers[t].st = RUNNABLE;
                }
            }
            for (unsigned t = 0; t < nthreads; t++) {
                if (g_fibers[t].st == AT_BLOCK_BAR) { g_fibers[t].st = RUNNABLE; all_done = false; }
                else if (g_fibers[t].st != DONE) { fprintf(stderr, "emu: scheduler inconsistency\n"); abort(); }
            }
            if (all_done) break;
        }
    }
    g_body = nullptr;
}
}  // namespace emu

// ---- runtime API ----------------------------------------------------------------------------------------
struct emu_stream { int id; };
struct emu_event { std::chrono::steady_clock::time_point t; };
struct emu_pool { int id; };
static emu_pool g_pool;
cudaError_t cudaGetDeviceCount(int* n) { *n = getenv("MDN_EMU_NO_DEVICE") ? 0 : 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
// ---- device memory: plain heap, or (MDN_EMU_SHM=1, the multi-rank tests) named shared memory so that another
//      rank's process can map it through the cudaIpc* stand-ins ---------------------------------------------
namespace {
struct ShmBlock { std::string name; size_t bytes; bool owner; };
std::map<void*, ShmBlock> g_shm;
unsigned g_shm_seq = 0;
bool shm_mode() { static int m = getenv("MDN_EMU_SHM") ? 1 : 0; return m != 0; }
struct ShmCleanup { ~ShmCleanup() { for (auto& kv : g_shm) if (kv.second.owner) shm_unlink(kv.second.name.c_str()); } } g_shm_cleanup;
}
long long clock64() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
namespace emu { void cpu_relax() { sched_yield(); } }
static void* plain_alloc(size_t bytes) { return aligned_alloc(256, (bytes + 255) / 256 * 256 + 256); }
cudaError_t cudaMalloc(void** p, size_t bytes) {
    if (!shm_mode()) { *p = plain_alloc(bytes); return *p ? cudaSuccess : cudaErrorInvalidValue; }
    char name[64];
    snprintf(name, sizeof name, "/mdn_emu_%d_%u", (int)getpid(), g_shm_seq++);
    size_t sz = (bytes + 4095) / 4096 * 4096 + 4096;
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return cudaErrorInvalidValue;
    if (ftruncate(fd, (off_t)sz) != 0) { close(fd); shm_unlink(name); return cudaErrorInvalidValue; }
    void* m = mmap(nullptr, sz, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { shm_unlink(name); return cudaErrorInvalidValue; }
    g_shm[m] = ShmBlock{name, sz, true};
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    auto it = g_shm.find(p);
    if (it == g_shm.end()) { free(p); return cudaSuccess; }
    munmap(p, it->second.bytes);
    if (it->second.owner) shm_unlink(it->second.name.c_str());
    g_shm.erase(it);
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    auto it = g_shm.find(p);
    if (it == g_shm.end() || !it->second.owner) return cudaErrorInvalidValue;
    memset(h, 0, sizeof *h);
    snprintf(h->reserved, 48, "%s", it->second.name.c_str());
    memcpy(h->reserved + 48, &it->second.bytes, sizeof(size_t));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    size_t sz; memcpy(&sz, h.reserved + 48, sizeof sz);
    int fd = shm_open(h.reserved, O_RDWR, 0600);
    if (fd < 0) return cudaErrorInvalidValue;
    void* m = mmap(nullptr, sz, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return cudaErrorInvalidValue;
    g_shm[m] = ShmBlock{h.reserved, sz, false};
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) { return cudaFree(p); }
cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t) { *p = plain_alloc(bytes); return *p ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { *p = plain_alloc(bytes); return *p ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t) { memset(dst, v, bytes); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emu_stream{0}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event{std::chrono::steady_clock::now()}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) { *pool = &g_pool; return cudaSuccess; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
cudaError_t cudaMemPoolGetAttribute(cudaMemPool_t, cudaMemPoolAttr, void* v) { *(uint64_t*)v = 0; return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) { a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = (void*)p; a->hostPointer = (void*)p; return cudaSuccess; }

// Marker binding.py looks for: a library that exports it is this emulator and is refused unless the caller
// (tests/test_emulated.py) opted in with MDN_ALLOW_EMULATOR=1.  The product library has no such symbol.
extern "C" int mdn_emulated_build() { return 1; }
