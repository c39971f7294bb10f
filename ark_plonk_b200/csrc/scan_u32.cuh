// Multi-block exclusive scan of u32 counters (bucket histograms, multiset sizes).
#pragma once
#include "common.cuh"

namespace apb {

// multi-block exclusive scan of u32 (block = 1024 elements), three small kernels
static __global__ void __launch_bounds__(256) k_u32_scan_block(const uint32_t* in, uint32_t* out, uint32_t* block_tot, uint64_t n) {
    __shared__ uint32_t sm[256];
    const uint32_t tid = threadIdx.x;
    const uint64_t base = ((uint64_t)blockIdx.x * 256 + tid) * 4;
    uint32_t e[4], run = 0;
    for (int j = 0; j < 4; j++) { e[j] = base + j < n ? in[base + j] : 0; run += e[j]; }
    sm[tid] = run;
    __syncthreads();
    for (uint32_t off = 1; off < 256; off <<= 1) {
        uint32_t v = tid >= off ? sm[tid - off] : 0;
        __syncthreads();
        sm[tid] += v;
        __syncthreads();
    }
    uint32_t acc = tid ? sm[tid - 1] : 0;
    for (int j = 0; j < 4; j++) { if (base + j < n) out[base + j] = acc; acc += e[j]; }
    if (tid == 255) block_tot[blockIdx.x] = sm[255];
}
static __global__ void k_u32_scan_totals(uint32_t* block_tot, uint32_t nb, uint32_t* total) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < nb; i++) { uint32_t e = block_tot[i]; block_tot[i] = acc; acc += e; }
    *total = acc;
}
static __global__ void k_u32_scan_apply(uint32_t* out, const uint32_t* block_tot, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_tot[i >> 10];
}


}  // namespace apb

// out[i] = sum of in[0..i); *total = sum of all.  block_tot: scratch of ceil(n/1024)+1 words.
static inline int u32_scan(const uint32_t* in, uint32_t* out, size_t n, uint32_t* block_tot, uint32_t* total) {
    const unsigned nb = (unsigned)((n + 1023) / 1024);
    APB_KLAUNCH(apb::k_u32_scan_block, nb, 256, 0, in, out, block_tot, (uint64_t)n);
    APB_KLAUNCH(apb::k_u32_scan_totals, 1, 32, 0, block_tot, nb, total);
    APB_KLAUNCH(apb::k_u32_scan_apply, (unsigned)((n + 255) / 256), 256, 0, out, (const uint32_t*)block_tot, (uint64_t)n);
    APB_CHECK_LAUNCH();
    return APB_OK;
}

