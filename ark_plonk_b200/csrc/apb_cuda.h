// Platform switch for the kernel sources.
//
// Product build (nvcc, sm_100a): plain CUDA runtime.
// APB_EMU build (g++, tests/emu only): the same kernel sources are compiled against a tiny
// thread-pool emulation of the CUDA execution model so that index/carry logic can be checked
// by the CPU test-suite on a box without a GPU.  The emulation library is test
// infrastructure: the Python package never loads it and there is no CPU fallback.
#pragma once

#ifdef APB_EMU
#include "cuda_emu.h"      // tests/emu/cuda_emu.h
#else
#include <cuda_runtime.h>
#define APB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define APB_DYN_SMEM(name)                                         \
    extern __shared__ __align__(16) unsigned char name##_raw_[];   \
    unsigned char* name = name##_raw_
#endif

#include <stdint.h>
#include <stddef.h>

#define APB_HD __host__ __device__ __forceinline__
#define APB_D __device__ __forceinline__
