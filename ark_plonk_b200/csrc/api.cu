// Library state, initialisation, device-memory helpers and diagnostics of the apb C ABI.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace apb {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
cudaStream_t g_own_stream = nullptr;
thread_local cudaStream_t t_user_stream = nullptr;
int g_device = 0;
bool g_inited = false;
double g_last_ms = 0.0;
std::recursive_mutex g_api_mutex;
extern int g_num_sms;

int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

template <class P>
__global__ void k_field_op(int op, const void* a, const void* b, void* out, uint64_t count) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fp<P> x = load_fp<P>(a, i), r;
    if (op <= 2) {
        Fp<P> y = load_fp<P>(b, i);
        r = op == 0 ? x * y : (op == 1 ? x + y : x - y);
    } else if (op == 5) {
        r = x.sqr();
    } else if (op == 6) {
        r = x.is_zero() ? x : x.inverse_binary();
    } else {
        r = op == 3 ? x.to_mont() : x.from_mont();
    }
    store_fp<P>(out, i, r);
}

// Throughput microbenchmark: ILP independent carry-free multiply-add chains per thread.
// WIDE = 1: mad.wide.u32 (64-bit accumulate, the IMAD.WIDE the Montgomery chains compile to);
// WIDE = 0: mad.lo.u32.
template <int WIDE>
__global__ void __launch_bounds__(256) k_imad_bench(uint32_t* out, uint32_t iters, uint32_t seed) {
#ifdef __CUDA_ARCH__
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    if (WIDE) {
        unsigned long long acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = a + k;
        for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
            }
        }
        unsigned long long s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) s ^= acc[k];
        if (s == 0x1234567) out[0] = (uint32_t)s;
    } else {
        uint32_t acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = a + k;
        for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(a + k), "r"(b));
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) s ^= acc[k];
        if (s == 0x1234567) out[0] = s;
    }
#else
    (void)out; (void)iters; (void)seed;
#endif
}

// dependent Montgomery products: ILP independent chains of x <- x * y per thread
template <class P, int ILP>
__global__ void k_mul_bench(void* out, uint32_t iters) {
    Fp<P> x[ILP], y = Fp<P>::r2();
#pragma unroll
    for (int k = 0; k < ILP; k++) {
        x[k] = Fp<P>::one();
        x[k].v[0] += threadIdx.x + k;
    }
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = x[k] * y;
    }
    Fp<P> s = x[0];
#pragma unroll
    for (int k = 1; k < ILP; k++) s = s + x[k];
    if (s.v[0] == 0x12345 && s.v[1] == 77) store_fp<P>(out, 0, s);
}

}  // namespace apb

using namespace apb;

template <class P>
static int run_mul_bench(int threads, int blocks_per_sm, int ilp, uint32_t iters, double* muls_per_s) {
    DevBuf out_buf;
    APB_CUDA_TRY(out_buf.alloc(256));
    void* d_out = out_buf.p;
    EventPair ev;
    unsigned blocks = (unsigned)(g_num_sms * blocks_per_sm);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(ev.a, cur_stream());
        auto k1 = k_mul_bench<P, 1>;
        auto k2 = k_mul_bench<P, 2>;
        auto k4 = k_mul_bench<P, 4>;
        if (ilp == 1) APB_KLAUNCH(k1, blocks, threads, 0, d_out, iters);
        else if (ilp == 2) APB_KLAUNCH(k2, blocks, threads, 0, d_out, iters);
        else APB_KLAUNCH(k4, blocks, threads, 0, d_out, iters);
        cudaEventRecord(ev.b, cur_stream());
        APB_CUDA_TRY(cudaEventSynchronize(ev.b));
        const float ms = ev.ms();
        if (rep > 0 && ms < best) best = ms;
    }
    int eff_ilp = ilp == 1 ? 1 : (ilp == 2 ? 2 : 4);
    *muls_per_s = (double)blocks * threads * iters * eff_ilp / (best * 1e-3);
    return APB_OK;
}

extern "C" int apb_mul_bench(int field, int threads, int blocks_per_sm, int ilp, uint32_t iters, double* muls_per_s) {
    APB_API_LOCK();
    APB_REQUIRE_INIT();
    if (!muls_per_s) return set_err(APB_ERR_INVALID_ARG, "apb_mul_bench: null out");
    switch (field) {
        case 0: return run_mul_bench<Fr381>(threads, blocks_per_sm, ilp, iters, muls_per_s);
        case 1: return run_mul_bench<Fq381>(threads, blocks_per_sm, ilp, iters, muls_per_s);
        case 2: return run_mul_bench<Fr377>(threads, blocks_per_sm, ilp, iters, muls_per_s);
        default: return run_mul_bench<Fq377>(threads, blocks_per_sm, ilp, iters, muls_per_s);
    }
}

extern "C" int apb_init(int device) {
    APB_API_LOCK();
    if (g_inited) return APB_OK;
#ifndef APB_EMU
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_err(APB_ERR_CUDA, "apb_init: no CUDA device (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device >= 0) APB_CUDA_TRY(cudaSetDevice(device));
    int dev = 0;
    APB_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    APB_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    g_num_sms = prop.multiProcessorCount;
    g_device = dev;              // every later entry point re-binds its calling thread to this device
#else
    (void)device;
    g_num_sms = 1;
#endif
    APB_CUDA_TRY(cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking));
    g_inited = true;
    return APB_OK;
}

// Work of the calling host thread is enqueued on `stream` (a cudaStream_t of the library's device)
// from now on; NULL returns to the library's own stream.  Device buffers produced on that stream
// can then be handed to the _dev entry points without any host synchronisation, and sync = 0
// calls return with their work merely enqueued there.
extern "C" int apb_set_stream(void* stream) {
    t_user_stream = (cudaStream_t)stream;
    return APB_OK;
}

extern "C" const char* apb_last_error(void) { return g_err; }
extern "C" const char* apb_version(void) {
#ifdef APB_EMU
    return "apb 0.1 (CPU emulation build - tests only)";
#else
    return "apb 0.1 (sm_100a)";
#endif
}
extern "C" uint64_t apb_kernel_launches(void) { return g_launches.load(); }
extern "C" double apb_last_device_ms(void) { return g_last_ms; }
extern "C" void* apb_stream(void) { return (void*)cur_stream(); }   // the stream this thread's calls use

extern "C" int apb_dev_alloc(size_t bytes, void** d_ptr) {
    APB_API_LOCK();
    if (!d_ptr) return set_err(APB_ERR_INVALID_ARG, "apb_dev_alloc: null out");
    APB_REQUIRE_INIT();
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return set_err(APB_ERR_OOM, "apb_dev_alloc(%zu): %s", bytes, cudaGetErrorString(e));
    return APB_OK;
}
extern "C" int apb_dev_free(void* d_ptr) {
    APB_API_LOCK();
    if (d_ptr) APB_CUDA_TRY(cudaFree(d_ptr));
    return APB_OK;
}
extern "C" int apb_dev_upload(void* d_dst, const void* h_src, size_t bytes) {
    APB_API_LOCK();
    APB_REQUIRE_INIT();
    APB_CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, cur_stream()));
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return APB_OK;
}
extern "C" int apb_dev_download(void* h_dst, const void* d_src, size_t bytes) {
    APB_API_LOCK();
    APB_REQUIRE_INIT();
    APB_CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, cur_stream()));
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return APB_OK;
}
extern "C" int apb_dev_sync(void) {
    APB_API_LOCK();
    APB_REQUIRE_INIT();
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return APB_OK;
}

extern "C" int apb_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t count) {
    APB_API_LOCK();
    if (field < 0 || field > 3 || op < 0 || op > 6) return set_err(APB_ERR_INVALID_ARG, "apb_field_op: bad field/op");
    if (!a || !out || (op <= 2 && !b)) return set_err(APB_ERR_INVALID_ARG, "apb_field_op: null argument");
    if (count == 0) return APB_OK;
    APB_REQUIRE_INIT();
    const size_t esz = (field == 0 || field == 2) ? 32 : 48;
    DevBuf ba, bb, bout;
    APB_CUDA_TRY(ba.alloc(count * esz));
    APB_CUDA_TRY(bb.alloc(count * esz));
    APB_CUDA_TRY(bout.alloc(count * esz));
    void *da = ba.p, *db = bb.p, *dout = bout.p;
    APB_CUDA_TRY(cudaMemcpyAsync(da, a, count * esz, cudaMemcpyHostToDevice, cur_stream()));
    if (b) APB_CUDA_TRY(cudaMemcpyAsync(db, b, count * esz, cudaMemcpyHostToDevice, cur_stream()));
    unsigned blocks = (unsigned)((count + 127) / 128);
    switch (field) {
        case 0: APB_KLAUNCH(k_field_op<Fr381>, blocks, 128, 0, op, (const void*)da, (const void*)db, dout, (uint64_t)count); break;
        case 1: APB_KLAUNCH(k_field_op<Fq381>, blocks, 128, 0, op, (const void*)da, (const void*)db, dout, (uint64_t)count); break;
        case 2: APB_KLAUNCH(k_field_op<Fr377>, blocks, 128, 0, op, (const void*)da, (const void*)db, dout, (uint64_t)count); break;
        default: APB_KLAUNCH(k_field_op<Fq377>, blocks, 128, 0, op, (const void*)da, (const void*)db, dout, (uint64_t)count); break;
    }
    APB_CHECK_LAUNCH();
    APB_CUDA_TRY(cudaMemcpyAsync(out, dout, count * esz, cudaMemcpyDeviceToHost, cur_stream()));
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return APB_OK;
}

extern "C" int apb_imad_peak(double* wide_per_s, double* imad32_per_s) {
    APB_API_LOCK();
    APB_REQUIRE_INIT();
    DevBuf out_buf;
    APB_CUDA_TRY(out_buf.alloc(64));
    uint32_t* d_out = out_buf.as<uint32_t>();
    const uint32_t iters = 4096;
    const unsigned blocks = (unsigned)g_num_sms * 8, threads = 256;
    double res[2] = {0, 0};
    for (int wide = 0; wide < 2; wide++) {
        EventPair ev;
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(ev.a, cur_stream());
            if (wide) APB_KLAUNCH(k_imad_bench<1>, blocks, threads, 0, d_out, iters, 12345u + rep);
            else APB_KLAUNCH(k_imad_bench<0>, blocks, threads, 0, d_out, iters, 12345u + rep);
            cudaEventRecord(ev.b, cur_stream());
            APB_CUDA_TRY(cudaEventSynchronize(ev.b));
            const float ms = ev.ms();
            if (rep > 0 && ms < best) best = ms;
        }
        double ops = (double)blocks * threads * iters * 32.0;
        res[wide] = best > 0 ? ops / (best * 1e-3) : 0;
    }
    if (wide_per_s) *wide_per_s = res[1];
    if (imad32_per_s) *imad32_per_s = res[0];
    return APB_OK;
}
