// morph.cuh -- binary post-processing of fused probability maps (reference fusion.py:295-328
// process_probability_image; multiatlas/run.py:373-404, 423-424):
//   sitk.BinaryFillhole           (itk::BinaryFillholeImageFilter, FullyConnected = false)
//   sitk.ConnectedComponent + LabelShapeStatistics + argmax(voxel count) + (label == k) + Cast(UInt8)
//   == sitk.RelabelComponent(sitk.ConnectedComponent(x)) == 1   (largest object, lowest label on ties)
//
// Both are face-connected component labellings of a binary volume: of the foreground for the largest
// object, of the background for the hole filling (background components that do not reach the image border are
// holes).  Labelling is a union-find over x-run segments in global memory:
//   init     one warp per 32-voxel chunk of a row: ballot of the class mask, every voxel points at the first voxel
//            of its run inside the chunk
//   merge    run heads are joined to the run ending in the previous chunk; a voxel is joined to its y-1 / z-1
//            neighbour only where that join is not implied by the previous voxel's (same two runs)
//   flatten  every voxel points at its root (the smallest linear index of its component == the first voxel in
//            raster order, which is also the order in which itk::ConnectedComponentImageFilter numbers objects)
// All integer work; results are bit-exact by construction.
#pragma once
#include "common.cuh"

namespace b200 {

// (volatile reads: other threads lower parents concurrently with atomicMin; any value read is a valid ancestor)
__device__ __forceinline__ int uf_find(const int* L, int a)
{
    const volatile int* V = L;
    int p = V[a];
    while (p != a) {
        a = p;
        p = V[a];
    }
    return a;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b)
{
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&L[a], b);  // a > b: hang the larger root under the smaller
        if (old == a) return;
        a = old;
    }
}

// CLS: voxels with (in[i] != 0) == CLS are labelled, the others get -1.  blockDim = (32, 8): one warp per row chunk.
template <bool CLS>
__global__ void __launch_bounds__(256) ccl_init_kernel(const uint8_t* __restrict__ in, int* __restrict__ L, int nx, int ny, int nz)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int row = blockIdx.y * 8 + threadIdx.y;  // y + ny * z
    if (row >= ny * nz) return;
    const int i = row * nx + x;
    const bool m = x < nx && ((in[i] != 0) == CLS);
    const unsigned mask = __ballot_sync(0xffffffffu, m);
    if (x >= nx) return;
    if (!m) {
        L[i] = -1;
        return;
    }
    // first lane of the run of set bits that contains this lane
    const unsigned below = ~mask & ((1u << threadIdx.x) - 1u);  // cleared bits below this lane
    const int start = below ? 32 - __clz(below) : 0;
    L[i] = i - (threadIdx.x - start);
}

// Class membership of the neighbours is read from L (>= 0), not from the image, so that the union-find only ever
// follows labels the init kernel wrote.
__global__ void __launch_bounds__(256) ccl_merge_kernel(int* L, int nx, int ny, int nz)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int row = blockIdx.y * 8 + threadIdx.y;
    if (row >= ny * nz || x >= nx) return;
    const int i = row * nx + x;
    const volatile int* V = L;
    if (V[i] < 0) return;
    const int y = row % ny, z = row / ny;
    const bool prev = x > 0 && V[i - 1] >= 0;
    // run continues across the chunk boundary
    if (threadIdx.x == 0 && prev) uf_union(L, i, i - 1);
    if (y > 0 && V[i - nx] >= 0) {
        const bool implied = prev && V[i - 1 - nx] >= 0;
        if (!implied) uf_union(L, i, i - nx);
    }
    if (z > 0) {
        const int pz = nx * ny;
        if (V[i - pz] >= 0) {
            const bool implied = prev && V[i - 1 - pz] >= 0;
            if (!implied) uf_union(L, i, i - pz);
        }
    }
}

__global__ void __launch_bounds__(256) ccl_flatten_kernel(int* L, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int l = L[q];
        if (l >= 0) L[q] = uf_find(L, (int)q);
    }
}

// voxel count per root (cnt zero-initialised)
__global__ void __launch_bounds__(256) ccl_count_kernel(const int* __restrict__ L, int* __restrict__ cnt, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int l = L[q];
        if (l >= 0) atomicAdd(&cnt[l], 1);
    }
}
// best = max over roots of (count << 32 | ~root): largest object, first in raster order on ties; ncomp = number of roots
__global__ void __launch_bounds__(256) ccl_best_kernel(const int* __restrict__ L, const int* __restrict__ cnt, size_t n, unsigned long long* best,
                                                       unsigned long long* ncomp)
{
    unsigned long long b = 0, c = 0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        if (L[q] == (int)q) {
            const unsigned long long key = ((unsigned long long)(unsigned)cnt[q] << 32) | (unsigned)(~(unsigned)q);
            b = key > b ? key : b;
            ++c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ob = __shfl_down_sync(0xffffffffu, b, o);
        b = ob > b ? ob : b;
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (b) atomicMax(best, b);
        if (c) atomicAdd(ncomp, c);
    }
}
__global__ void __launch_bounds__(256) ccl_select_kernel(const int* __restrict__ L, const unsigned long long* __restrict__ best, uint8_t* __restrict__ out,
                                                         size_t n)
{
    const unsigned long long b = *best;
    const int root = b ? (int)(~(unsigned)(b & 0xffffffffull)) : -2;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = L[q] == root ? 1 : 0;
}

// background components that reach the image border are "outside": flag their roots
__global__ void __launch_bounds__(256) fillhole_border_kernel(const int* __restrict__ L, int* __restrict__ flag, int nx, int ny, int nz)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int l = L[q];
        if (l < 0) continue;
        const int x = (int)(q % nx), y = (int)((q / nx) % ny), z = (int)(q / ((size_t)nx * ny));
        if (x == 0 || x == nx - 1 || y == 0 || y == ny - 1 || z == 0 || z == nz - 1) flag[l] = 1;
    }
}
__global__ void __launch_bounds__(256) fillhole_write_kernel(const int* __restrict__ L, const int* __restrict__ flag, uint8_t* __restrict__ out, size_t n,
                                                             uint8_t foreground)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int l = L[q];
        out[q] = (l < 0 || !flag[l]) ? foreground : 0;  // foreground voxel, or background not connected to the border
    }
}

template <bool CLS>
inline int ccl_label(b200reg_ctx* ctx, const uint8_t* d_in, int* L, int nx, int ny, int nz)
{
    const size_t n = (size_t)nx * ny * nz;
    const dim3 blk(32, 8, 1);
    const dim3 grd((nx + 31) / 32, (unsigned)(((size_t)ny * nz + 7) / 8), 1);
    ccl_init_kernel<CLS><<<grd, blk, 0, ctx->stream>>>(d_in, L, nx, ny, nz);
    ccl_merge_kernel<<<grd, blk, 0, ctx->stream>>>(L, nx, ny, nz);
    ccl_flatten_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(L, n);
    ctx->launches += 3;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

inline int check_ccl_size(const int size[3])
{
    if (!size || size[0] <= 0 || size[1] <= 0 || size[2] <= 0) return set_error(B200REG_ERR_ARG, "invalid size");
    if ((size_t)size[0] * size[1] * size[2] >= (1ull << 31)) return set_error(B200REG_ERR_UNSUPPORTED, "volumes of 2^31 voxels or more are not supported");
    if ((size_t)size[1] * size[2] > 8ull * 65535ull) return set_error(B200REG_ERR_UNSUPPORTED, "more than 524280 rows are not supported");
    return B200REG_OK;
}

// itk::BinaryFillholeImageFilter (FullyConnected off): d_out may alias d_in.
inline int binary_fillhole(b200reg_ctx* ctx, const uint8_t* d_in, const int size[3], uint8_t foreground, uint8_t* d_out)
{
    B200_TRY(check_ccl_size(size));
    const size_t n = (size_t)size[0] * size[1] * size[2];
    TempBuf L, flag;
    B200_TRY(L.alloc(ctx, n * sizeof(int)));
    B200_TRY(flag.alloc(ctx, n * sizeof(int)));
    B200_CUDA(cudaMemsetAsync(flag.p, 0, n * sizeof(int), ctx->stream));
    B200_TRY(ccl_label<false>(ctx, d_in, L.as<int>(), size[0], size[1], size[2]));
    fillhole_border_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(L.as<int>(), flag.as<int>(), size[0], size[1], size[2]);
    fillhole_write_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(L.as<int>(), flag.as<int>(), d_out, n, foreground);
    ctx->launches += 2;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// ConnectedComponent -> LabelShapeStatistics -> largest (first on ties) -> UInt8 mask.  d_out may alias d_in.
// d_info (device, 2 x u64, optional read-back by the caller): best key, number of components.
inline int largest_component(b200reg_ctx* ctx, const uint8_t* d_in, const int size[3], uint8_t* d_out, unsigned long long* d_info)
{
    B200_TRY(check_ccl_size(size));
    const size_t n = (size_t)size[0] * size[1] * size[2];
    TempBuf L, cnt;
    B200_TRY(L.alloc(ctx, n * sizeof(int)));
    B200_TRY(cnt.alloc(ctx, n * sizeof(int)));
    B200_CUDA(cudaMemsetAsync(cnt.p, 0, n * sizeof(int), ctx->stream));
    B200_CUDA(cudaMemsetAsync(d_info, 0, 2 * sizeof(unsigned long long), ctx->stream));
    B200_TRY(ccl_label<true>(ctx, d_in, L.as<int>(), size[0], size[1], size[2]));
    ccl_count_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(L.as<int>(), cnt.as<int>(), n);
    ccl_best_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(L.as<int>(), cnt.as<int>(), n, d_info, d_info + 1);
    ccl_select_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(L.as<int>(), d_info, d_out, n);
    ctx->launches += 3;
    B200_CHECK_LAUNCH();
    return B200REG_OK;
}

// probability_image / max -> BinaryThreshold(lower = threshold, upper = 255) (fusion.py:304-308); the maximum is read
// from device memory (d_minmax[1]) so that the whole chain runs without a host round trip.
template <typename T>
__global__ void __launch_bounds__(256) normalise_threshold_kernel(const T* __restrict__ in, const double* __restrict__ d_minmax, double lower, double upper,
                                                                  uint8_t* __restrict__ out, size_t n)
{
    const double mx = d_minmax[1];
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        // itk::Functor::Div<T, double, T>: A / B in double, cast to T; B == 0 -> NumericTraits<T>::max()
        T v;
        if (mx != 0.0) v = (T)((double)in[q] / mx);
        else v = sizeof(T) == 4 ? (T)FLT_MAX : (T)DBL_MAX;
        const double d = (double)v;
        out[q] = (d >= lower && d <= upper) ? 1 : 0;
    }
}

}  // namespace b200
