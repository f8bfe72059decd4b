// TEST INFRASTRUCTURE ONLY -- not part of the product, never loaded by binding.py's default path.
//
// A stand-in for <cuda_runtime.h> that lets tests/emu/Makefile compile the product's kernels.cu and session.cu
// with g++ and run every kernel on the CPU, one cooperative fiber per CUDA thread (tests/emu/emu_runtime.cpp), so
// that the host orchestration and the kernels' index arithmetic can be exercised by `-m "not gpu"` tests in a
// container without a GPU.  It plays the role a CUDA simulator would: same sources, same launch geometry, same
// shared-memory tiles and barriers.  What it cannot show: PTX carry-flag primitives (the host build takes the
// unsigned __int128 branch of poseidon2_fast2.cuh), register/shared-memory limits, races, performance.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __host__
#define __device__
#define __global__
#define __constant__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define MDN_EMULATED 1

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { unsigned x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r; r.x = x; r.y = y; return r; }

namespace emu {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void syncthreads();
unsigned long long shfl_down(unsigned long long v, unsigned off);
extern unsigned char* g_dyn_smem;     // dynamic shared memory of the running block
}
#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
static inline void __syncthreads() { emu::syncthreads(); }
static inline unsigned long long __shfl_down_sync(unsigned, unsigned long long v, unsigned off) { return emu::shfl_down(v, off); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
// one OS thread runs all fibers, so plain read-modify-write is atomic here
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
// cross-process "peer memory" (one proof on several ranks): the arena slabs are POSIX shared memory here, the
// system-scope fence is a host fence, and clock64() counts nanoseconds
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
long long clock64();
namespace emu { void cpu_relax(); }

// ---- runtime API subset used by session.cu / kernels.cu ------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNoDevice = 100 };
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
typedef struct emu_pool* cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4, cudaMemPoolAttrReservedMemCurrent, cudaMemPoolAttrReservedMemHigh, cudaMemPoolAttrUsedMemCurrent, cudaMemPoolAttrUsedMemHigh };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };

extern "C++" {
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaDeviceSynchronize();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t s);
cudaError_t cudaFreeAsync(void* p, cudaStream_t s);
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned flags);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int device);
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t pool, cudaMemPoolAttr a, void* v);
cudaError_t cudaMemPoolGetAttribute(cudaMemPool_t pool, cudaMemPoolAttr a, void* v);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p);
// CUDA IPC stand-in: with MDN_EMU_SHM=1 every cudaMalloc is a named POSIX shared-memory object, the handle carries
// its name, and cudaIpcOpenMemHandle maps it into the opening process
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
}
namespace emu { extern size_t g_max_dyn_smem_opt_in; }
// not per kernel: remembers the largest opt-in so that a launch above the 48 KB default without ANY opt-in is caught
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int v) { if ((size_t)v > emu::g_max_dyn_smem_opt_in) emu::g_max_dyn_smem_opt_in = (size_t)v; return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t bytes) { memcpy((void*)&sym, src, bytes); return cudaSuccess; }

// ---- kernel launch: `k<<<grid, block, smem, stream>>>(args)` is rewritten by gen.py into
//      emu::launch(emu::cfg(grid, block, smem, stream), [&] { k(args); }); ----------------------------------------
#include <functional>
namespace emu {
struct LaunchCfg { dim3 grid, block; size_t smem; };
static inline LaunchCfg cfg(dim3 g, dim3 b, size_t smem = 0, cudaStream_t = nullptr) { LaunchCfg c{g, b, smem}; return c; }
void launch(const LaunchCfg& c, const std::function<void()>& body);
}
