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
extern cudaStream_t g_stream;
extern bool g_inited;
extern double g_last_ms;
extern std::recursive_mutex g_api_mutex;   // the library has one stream and shared workspaces: entry points serialise

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
        APB_LAUNCH(kernel, grid, block, smem, apb::g_stream, __VA_ARGS__);                  \
    } while (0)

#define APB_API_LOCK() std::lock_guard<std::recursive_mutex> apb_api_lock_(apb::g_api_mutex)

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
    typedef Fq381_28 FQ28;
};
struct Curve377 {
    typedef Fr377 FR;
    typedef Fq377 FQ;
    typedef Fq377_28 FQ28;
};

}  // namespace apb
