// cuda_emu.h -- TEST INFRASTRUCTURE.  A serial host emulation of the CUDA execution model for kernels that use neither shared
// memory nor barriers nor warp intrinsics (platipy_b200/csrc/distmap_kernels.cuh): every (block, thread) pair runs to
// completion, one after another.  The build container has no GPU; this lets the CPU test-suite check such a kernel's
// indexing and arithmetic against the oracle here, before the -m gpu tests run the real launch on a B200.  Nothing in
// platipy_b200 includes, links or loads this: the product runs the same source compiled by nvcc for sm_100a.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>

#include <algorithm>
using std::max;
using std::min;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define B200_HOST_EMU 1

struct emu_dim3 {
    unsigned x = 1, y = 1, z = 1;
};
static thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

// atomics of a serial emulation are plain read-modify-writes
template <typename T>
static T atomicAdd(T* address, T value)
{
    const T old = *address;
    *address = old + value;
    return old;
}

template <typename K, typename... A>
static void emu_launch(K kernel, unsigned grid, unsigned block, A... args)
{
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t) {
            blockIdx.x = b;
            threadIdx.x = t;
            kernel(args...);
        }
}
