// Minimal CPU emulation of the CUDA execution model - TEST INFRASTRUCTURE ONLY.
//
// Lets the CPU test-suite (pytest -m "not gpu", no GPU in the build container) execute the
// *same kernel sources* as the product and compare them with the oracle, so that indexing,
// carry-chain and synchronisation logic is checked before GPU minutes are spent.  It is
// built only by tests/emu/build_emu.py into tests/emu/_build/, is never loaded by the
// ark_plonk_b200 package, and is not a fallback: the product library requires a GPU.
//
// Model: one OS thread per CUDA thread of a block; blocks run one after another;
// __syncthreads() is a std::barrier; `__shared__` variables are function-local statics.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

namespace apb_emu {
inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline std::barrier<>* g_barrier = nullptr;
inline thread_local std::barrier<>* t_warp_barrier = nullptr;      // of this thread's warp (apb_emu::warp_sync)
inline uint8_t* g_dyn_smem = nullptr;
inline uint32_t g_shfl_buf[2048][16];

template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 block, size_t smem, A... args) {
    const unsigned nt = block.x * block.y * block.z;
    std::barrier<> bar(nt);
    g_barrier = &bar;
    std::vector<uint8_t> dyn(smem + 64);
    g_dyn_smem = dyn.data();
    std::vector<std::unique_ptr<std::barrier<>>> warp_bars;        // warps = 32 consecutive linear thread ids
    for (unsigned w = 0; w * 32 < nt; w++) warp_bars.emplace_back(new std::barrier<>(nt - w * 32 < 32 ? nt - w * 32 : 32));
    auto worker = [&](unsigned tid) {
        t_warp_barrier = warp_bars[tid / 32].get();
        t_blockDim = block;
        t_gridDim = grid;
        t_threadIdx = dim3(tid % block.x, (tid / block.x) % block.y, tid / (block.x * block.y));
        for (unsigned bz = 0; bz < grid.z; bz++)
            for (unsigned by = 0; by < grid.y; by++)
                for (unsigned bx = 0; bx < grid.x; bx++) {
                    t_blockIdx = dim3(bx, by, bz);
                    kernel(args...);
                    bar.arrive_and_wait();
                }
    };
    if (nt == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; t++) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    g_barrier = nullptr;
    g_dyn_smem = nullptr;
}
// a REAL barrier over the 32 lanes of the calling thread's warp: for kernels whose lanes exchange data through
// shared memory (every lane of the warp must call it the same number of times)
inline void warp_sync() { t_warp_barrier->arrive_and_wait(); }
}  // namespace apb_emu

#define threadIdx (apb_emu::t_threadIdx)
#define blockIdx (apb_emu::t_blockIdx)
#define blockDim (apb_emu::t_blockDim)
#define gridDim (apb_emu::t_gridDim)
#define APB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    apb_emu::launch(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#define APB_DYN_SMEM(name) unsigned char* name = apb_emu::g_dyn_smem

static inline void __syncthreads() { apb_emu::g_barrier->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline uint32_t atomicMin(uint32_t* p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline uint32_t atomicCAS(uint32_t* p, uint32_t cmp, uint32_t val) {
    uint32_t expected = cmp;
    __atomic_compare_exchange_n(p, &expected, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline uint32_t __brev(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
static inline uint4 __ldg(const uint4* p) { return *p; }
static inline uint32_t __ldg(const uint32_t* p) { return *p; }

// ---- runtime API shim -------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
#ifdef APB_EMU_ASAN      // exact-size allocations so AddressSanitizer sees every out-of-bounds access
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
#else
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
#endif
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
enum { cudaStreamNonBlocking = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return 0; }
