// Library-wide state: error reporting, the per-process stream, launch accounting.
#pragma once
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <string>

#include "../../include/apb.h"
#include "apb_cuda.h"
#include "ec.cuh"

namespace apb {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;
extern cudaStream_t g_own_stream;                    // created by apb_init; used when the caller set none
extern thread_local cudaStream_t t_user_stream;     // apb_set_stream: the calling host thread's stream
extern bool g_inited;
extern int g_device;                                // the device apb_init bound the library to
extern double g_last_ms;
// Entry points that take a handle (commitment key, domain) serialise on THAT handle's mutex and use
// only its workspaces, so two handles driven from two host threads / streams run concurrently.
// Handle-less entry points (polynomial utilities: one shared workspace) serialise on g_api_mutex.
extern std::recursive_mutex g_api_mutex;

// the stream every launch / copy of the current entry point is enqueued on
inline cudaStream_t cur_stream() { return t_user_stream ? t_user_stream : g_own_stream; }
// CUDA's current device is per host thread: entry points may be called from any thread
inline void bind_device() {
#ifndef APB_EMU
    if (g_inited) cudaSetDevice(g_device);
#endif
}
struct ApiGuard {
    std::unique_lock<std::recursive_mutex> lk;
    explicit ApiGuard(std::recursive_mutex& m) : lk(m) { bind_device(); }
};
// temporary device buffers / events of one entry point: released on every return path
struct DevBuf {
    void* p = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    EventPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
    EventPair(const EventPair&) = delete;
    EventPair& operator=(const EventPair&) = delete;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    float ms() const { float v = 0; cudaEventElapsedTime(&v, a, b); return v; }
};

int set_err(int code, const char* fmt, ...);

#define APB_CUDA_TRY(expr)                                                                  \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return apb::set_err(APB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                                cudaGetErrorString(e_), __FILE__, __LINE__);                \
    } while (0)

#define APB_KLAUNCH(kernel, grid, block, smem, ...)                                         \
    do {                                                                                    \
        apb::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
        APB_LAUNCH(kernel, grid, block, smem, apb::cur_stream(), __VA_ARGS__);                  \
    } while (0)

#define APB_API_LOCK() apb::ApiGuard apb_api_lock_(apb::g_api_mutex)
#define APB_HANDLE_LOCK(h) apb::ApiGuard apb_handle_lock_((h)->mu)

#define APB_CHECK_LAUNCH() APB_CUDA_TRY(cudaGetLastError())

#define APB_REQUIRE_INIT()                                                                  \
    do {                                                                                    \
        if (!apb::g_inited) {                                                               \
            int rc_ = apb_init(-1);                                                         \
            if (rc_ != APB_OK) return rc_;                                                  \
        }                                                                                   \
    } while (0)

// 32-byte (Fr) and 48-byte (Fq) elements move as 128-bit vectors
template <class P>
APB_D Fp<P> load_fp(const void* base, size_t idx) {
    Fp<P> r;
    const uint4* p = reinterpret_cast<const uint4*>(base) + idx * (P::N / 4);
#pragma unroll
    for (int i = 0; i < P::N / 4; i++) {
        uint4 t = p[i];
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
// Same load as an `asm volatile` statement: it keeps its program position relative to the (volatile)
// multiply-add chains, so a prefetch written at the top of a loop body is ISSUED there.  (A plain
// load whose value is only consumed at the bottom gets sunk to the bottom by the compiler under
// register pressure, and the software pipeline silently disappears.)  NC: read-only data path.
template <class P, int NC>
APB_D Fp<P> load_fp_early(const void* base, size_t idx) {
#ifdef __CUDA_ARCH__
    Fp<P> r;
    const uint4* p = reinterpret_cast<const uint4*>(base) + idx * (P::N / 4);
#pragma unroll
    for (int i = 0; i < P::N / 4; i++) {
        if (NC)
            asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r.v[4 * i]), "=r"(r.v[4 * i + 1]), "=r"(r.v[4 * i + 2]), "=r"(r.v[4 * i + 3]) : "l"(p + i));
        else
            asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r.v[4 * i]), "=r"(r.v[4 * i + 1]), "=r"(r.v[4 * i + 2]), "=r"(r.v[4 * i + 3]) : "l"(p + i) : "memory");
    }
    return r;
#else
    return load_fp<P>(base, idx);
#endif
}
template <class P>
APB_D void store_fp(void* base, size_t idx, const Fp<P>& a) {
    uint4* p = reinterpret_cast<uint4*>(base) + idx * (P::N / 4);
#pragma unroll
    for (int i = 0; i < P::N / 4; i++) p[i] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
}

struct Curve381 {
    typedef Fr381 FR;
    typedef Fq381 FQ;
};
struct Curve377 {
    typedef Fr377 FR;
    typedef Fq377 FQ;
};

}  // namespace apb
