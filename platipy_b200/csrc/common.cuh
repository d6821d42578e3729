// common.cuh -- context, error plumbing, device geometry and pixel traits shared by all kernels.
// Compiled with -fmad=false: no FMA contraction anywhere, so f64 arithmetic is reproducible against the
// CPU oracle wherever the operation order is the same.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges cost nothing unless a profiler injects its library

#include "../../include/b200reg.h"

// NVTX range for the enclosing scope (Nsight Systems / ncu --nvtx show the pyramid, each level's Demons loop, the glue and the
// fusion stages by name)
struct B200NvtxRange {
    bool open = true;
    explicit B200NvtxRange(const char* name) { nvtxRangePushA(name); }
    void end()
    {
        if (open) nvtxRangePop();
        open = false;
    }
    ~B200NvtxRange() { end(); }
    B200NvtxRange(const B200NvtxRange&) = delete;
    B200NvtxRange& operator=(const B200NvtxRange&) = delete;
};
#define B200_NVTX_CAT2(a, b) a##b
#define B200_NVTX_CAT(a, b) B200_NVTX_CAT2(a, b)
#define B200_NVTX(name) B200NvtxRange B200_NVTX_CAT(_nvtx_range_, __LINE__)(name)

namespace b200 {

// ---- error plumbing --------------------------------------------------------------------------------
inline std::string& last_error_ref()
{
    static thread_local std::string s;
    return s;
}
inline int set_error(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}
#define B200_CUDA(expr)                                                                                          \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess)                                                                                   \
            return ::b200::set_error(B200REG_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,      \
                                     cudaGetErrorString(_e));                                                    \
    } while (0)
#define B200_CHECK_LAUNCH() B200_CUDA(cudaGetLastError())
#define B200_TRY(expr)                       \
    do {                                     \
        int _s = (expr);                     \
        if (_s != B200REG_OK) return _s;     \
    } while (0)

}  // namespace b200

// Named semantic switches (process-wide): the ITK behaviours that could only be recalled (confidence M) and are
// cheap to state both ways.  Same names and values as the switches of the CPU checker used by tests/ (see include/b200reg.h); the defaults are
// the recalled behaviours.  DESIGN.md section 5 lists each switch with the kernels / host loops it reaches.
namespace b200 {
struct Semantics {
    int discrete_gaussian_axis_order = 0;     // 0: z, y, x   1: x, y, z                        (gauss.cuh: discrete_gaussian_f32)
    int recursive_gaussian_axis_order = 0;    // 0: z, x, y   1: x, y, z                        (deriche.cuh: recursive_gaussian_vec3)
    int resample_linear_scanline = 1;         // 1: scan-line continuous index for linear chains  0: per voxel   (resample.cuh: make_chain)
    int dvf_transform_interpolation = 0;      // 0: weighted sum of 8 neighbours  1: nested lerps  (resample.cuh: apply_chain)
    int vector_resample_interpolation = 0;    // 0: nested lerps  1: weighted sum                  (resample.cuh: resample_vec3_kernel)
    int binary_threshold_in_pixel_type = 0;   // 0: bounds compared as real numbers  1: bounds cast to the pixel type first (fusion.cuh)
};
inline Semantics& semantics()
{
    static Semantics s;
    return s;
}
inline int* semantic_slot(const char* name)
{
    Semantics& s = semantics();
    struct { const char* n; int* p; } tab[] = {
        { "discrete_gaussian_axis_order", &s.discrete_gaussian_axis_order }, { "recursive_gaussian_axis_order", &s.recursive_gaussian_axis_order },
        { "resample_linear_scanline", &s.resample_linear_scanline }, { "dvf_transform_interpolation", &s.dvf_transform_interpolation },
        { "vector_resample_interpolation", &s.vector_resample_interpolation }, { "binary_threshold_in_pixel_type", &s.binary_threshold_in_pixel_type },
    };
    for (auto& e : tab)
        if (name && std::strcmp(e.n, name) == 0) return e.p;
    return nullptr;
}
}  // namespace b200

struct b200reg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int64_t launches = 0;
    int sm_count = 148;
    // pinned scratch for small read-backs
    double* h_scratch = nullptr;  // 64 doubles
    double* h_trace = nullptr;    // pinned landing area of the Demons iteration trace (grown on demand)
    size_t h_trace_doubles = 0;
    std::set<const void*> smem_optin;  // kernels whose dynamic shared-memory limit was raised on this device
    // per-iteration (metric, RMS change) pairs of the most recent Demons call, one vector per level (b200reg_demons_trace)
    std::vector<std::vector<double>> traces;
    bool force_separable = false;  // B200REG_FORCE_SEPARABLE=1: unfused smoothing passes (A/B testing)
    bool unfused_force = false;    // B200REG_UNFUSED_FORCE=1: separate warp and force kernels (W through HBM)
    bool staple_voxelwise = false; // B200REG_STAPLE_VOXELWISE=1: per-voxel EM kernels instead of the pattern-histogram EM
    int zm_tma = 1;                // B200REG_ZM_TMA=0: cp.async (LDGSTS) staging of the fused smoothing kernel's plane tiles instead of one tensor-map TMA copy per tile
    int zm_tma_l2 = 3;             // B200REG_ZM_TMA_L2 = 0 | 1 | 2 | 3: L2 promotion of the tensor-map loads (none, 64, 128, 256 bytes)
    bool zm_regadd = false;        // B200REG_ZM_REGADD=0: add + smooth stages both operands in shared memory (first version)
    int pf_warp = 0;               // B200REG_PF_WARP=n: warp kernel prefetches the field n planes ahead into L2
    int pf_force = 0;              // B200REG_PF_FORCE=n: force kernel prefetches W / F n steps ahead into L2
    int warp_march = 0;            // B200REG_WARP_MARCH=n: z-marching warp kernel with n planes per thread (0: one-shot kernel)
    int zm_chunks = 0;             // B200REG_ZM_CHUNKS=n: z-chunks per tile column of the fused smoothing kernel (0: automatic)
    bool zm_tx32 = true;           // B200REG_ZM_TX32=0: 64-wide tiles (320 threads, 2 CTAs per SM) in the fused smoothing kernel; 32-wide: 4 CTAs per SM, -1 %
    bool zm_addout = true;         // B200REG_ZM_ADDOUT=0: D + U formed inside the displacement smoothing (staged twice) instead of at the end of the update smoothing
    double pyramid_restrict_cost = 0.6;  // B200REG_PYRAMID_RESTRICT_COST: largest restricted-work estimate (in full passes) that takes pyramid.cuh
    bool pyramid_restrict = true;  // B200REG_PYRAMID_RESTRICT=0: shrinking pyramid levels blur the whole image before resampling it (pyramid.cuh) (A/B)
    bool warp_resample = true;     // B200REG_WARP_RESAMPLE=0: a Float32 image resampled through a field on the output grid takes the generic batch kernel (A/B)
    bool conv_static_radius = true;  // B200REG_CONV_STATIC_RADIUS=0: the Float32 Gaussian passes always take the run-time-radius kernels (A/B)
    bool identity_copy = true;     // B200REG_IDENTITY_COPY=0: identity re-grids onto an identical grid always run the resampling kernel (A/B)
    bool pdl = true;               // B200REG_PDL=0: the Demons loop kernels are launched without programmatic dependent launch (A/B)
    bool pack_labels = true;       // B200REG_PACK_LABELS=0: UInt8 nearest-neighbour items of a resample batch are gathered one byte at a time (A/B)
    bool force_zm1 = false;        // B200REG_FORCE_ZM1=1: first-generation fused smoothing kernel
};

namespace b200 {

// Stream-ordered temporary buffer (cudaMallocAsync pool; release threshold is raised at ctx creation so
// freed blocks stay cached in the pool).
struct TempBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    TempBuf() = default;
    TempBuf(const TempBuf&) = delete;
    TempBuf& operator=(const TempBuf&) = delete;
    int alloc(b200reg_ctx* ctx, size_t bytes)
    {
        release();
        s = ctx->stream;
        B200_CUDA(cudaMallocAsync(&p, bytes ? bytes : 16, s));
        return B200REG_OK;
    }
    void release()
    {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    ~TempBuf() { release(); }
    template <typename T>
    T* as() const
    {
        return reinterpret_cast<T*>(p);
    }
};

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------------------
// The Demons loop is a chain of short dependent kernels (five per iteration; at the coarse pyramid levels each runs for 10-20 us).
// Launched with cudaLaunchAttributeProgrammaticStreamSerialization, a kernel's CTAs may become resident while the tail of its
// predecessor is still running: every loop kernel signals `pdl_launch_dependents()` at once, runs the part of its prologue that
// reads no data of the predecessor (index arithmetic, mbarrier / tensor-map set-up), and calls `pdl_wait()` before its first
// dependent access -- the wait returns when the predecessor grid has completed and its writes are visible.  A kernel launched
// without the attribute sees both as no-ops.  B200REG_PDL=0 turns the attribute off (A/B).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(b200reg_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = ctx->pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- small read-backs ------------------------------------------------------------------------------------------------------------
// A few bytes of statistics go back to the host after most calls (min / max, the Demons control block, metric sums).  As
// cudaMemcpyAsync they would queue on the device-to-host copy engine BEHIND any bulk download in flight on another stream -- the
// pipelined host API keeps a 1.6 GB field download running while the next registration starts, and a 16-byte read-back then
// blocks the host for 30 ms.  So they are stores by one warp into pinned host memory (directly addressable under UVA) instead:
// ordered on the compute stream, independent of the copy engines.
__global__ void small_d2h_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int nwords)
{
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
inline cudaError_t small_d2h(b200reg_ctx* ctx, void* h_pinned_dst, const void* d_src, size_t bytes)
{
    small_d2h_kernel<<<1, 128, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_src), reinterpret_cast<uint32_t*>(h_pinned_dst), (int)((bytes + 3) / 4));
    return cudaGetLastError();
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: opt in once per context and kernel.
template <typename K>
inline int ensure_dynamic_smem(b200reg_ctx* ctx, K kernel, size_t bytes)
{
    const void* key = reinterpret_cast<const void*>(kernel);
    if (ctx->smem_optin.count(key)) return B200REG_OK;
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    ctx->smem_optin.insert(key);
    return B200REG_OK;
}

// ---- geometry --------------------------------------------------------------------------------------
struct GeomD {
    int nx, ny, nz;
    double origin[3];
    double i2p[9];  // Direction * diag(Spacing)
    double p2i[9];  // inverse
    double spacing[3];
    double direction[9];
    double hi[3];  // size - 0.5: upper bound of the continuous-index range inside the buffer
    int small;     // fewer than 2^31 voxels: kernels may use 32-bit element offsets
};

inline void inv3(const double* m, double* o)
{
    // diagonal (identity direction): exact reciprocals
    if (m[1] == 0 && m[2] == 0 && m[3] == 0 && m[5] == 0 && m[6] == 0 && m[7] == 0) {
        for (int i = 0; i < 9; ++i) o[i] = 0.0;
        o[0] = 1.0 / m[0];
        o[4] = 1.0 / m[4];
        o[8] = 1.0 / m[8];
        return;
    }
    double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    o[0] = c00 / det;
    o[1] = (m[2] * m[7] - m[1] * m[8]) / det;
    o[2] = (m[1] * m[5] - m[2] * m[4]) / det;
    o[3] = c01 / det;
    o[4] = (m[0] * m[8] - m[2] * m[6]) / det;
    o[5] = (m[2] * m[3] - m[0] * m[5]) / det;
    o[6] = c02 / det;
    o[7] = (m[1] * m[6] - m[0] * m[7]) / det;
    o[8] = (m[0] * m[4] - m[1] * m[3]) / det;
}

inline GeomD make_geomd(const b200reg_geom& s)
{
    GeomD g;
    g.nx = s.size[0];
    g.ny = s.size[1];
    g.nz = s.size[2];
    for (int r = 0; r < 3; ++r) {
        g.origin[r] = s.origin[r];
        g.spacing[r] = s.spacing[r];
        for (int c = 0; c < 3; ++c) {
            g.i2p[r * 3 + c] = s.direction[r * 3 + c] * s.spacing[c];
            g.direction[r * 3 + c] = s.direction[r * 3 + c];
        }
    }
    inv3(g.i2p, g.p2i);
    g.hi[0] = g.nx - 0.5;
    g.hi[1] = g.ny - 0.5;
    g.hi[2] = g.nz - 0.5;
    g.small = ((size_t)g.nx * g.ny * g.nz < (1ull << 31)) ? 1 : 0;
    return g;
}
inline size_t nvox(const b200reg_geom& g) { return (size_t)g.size[0] * g.size[1] * g.size[2]; }
// identity direction cosines: index -> point and point -> index separate per axis (the zero terms add exact zeros)
inline bool geom_is_diag(const GeomD& g)
{
    const double* d = g.direction;
    return d[0] == 1.0 && d[4] == 1.0 && d[8] == 1.0 && d[1] == 0.0 && d[2] == 0.0 && d[3] == 0.0 && d[5] == 0.0 && d[6] == 0.0 && d[7] == 0.0;
}
inline bool valid_geom(const b200reg_geom* g)
{
    if (!g) return false;
    for (int i = 0; i < 3; ++i)
        if (g->size[i] <= 0 || !(g->spacing[i] > 0.0)) return false;
    return true;
}

// ImageBase::TransformIndexToPhysicalPoint: sum over columns, then + origin
__device__ __forceinline__ void idx2pt(const GeomD& g, double i0, double i1, double i2, double* p)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double sum = 0.0;
        sum += g.i2p[r * 3 + 0] * i0;
        sum += g.i2p[r * 3 + 1] * i1;
        sum += g.i2p[r * 3 + 2] * i2;
        p[r] = sum + g.origin[r];
    }
}
// ImageBase::TransformPhysicalPointToContinuousIndex
__device__ __forceinline__ void pt2cidx(const GeomD& g, const double* p, double* c)
{
    double v0 = p[0] - g.origin[0], v1 = p[1] - g.origin[1], v2 = p[2] - g.origin[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double sum = 0.0;
        sum += g.p2i[r * 3 + 0] * v0;
        sum += g.p2i[r * 3 + 1] * v1;
        sum += g.p2i[r * 3 + 2] * v2;
        c[r] = sum;
    }
}
// ImageFunction::IsInsideBuffer(ContinuousIndex): [-0.5, size - 0.5), NaN -> outside
__device__ __forceinline__ bool inside_buffer(const GeomD& g, const double* c)
{
    return (c[0] >= -0.5 && c[0] < g.hi[0] && c[1] >= -0.5 && c[1] < g.hi[1] && c[2] >= -0.5 && c[2] < g.hi[2]);
}

// ---- pixel traits: load as double; store with CastPixelWithBoundsChecking ------------------------------
template <typename T>
struct Px;
#define B200_INT_PX(T, LO, HI)                                                      \
    template <>                                                                     \
    struct Px<T> {                                                                  \
        __device__ static __forceinline__ double ld(const T* p, size_t i) { return (double)p[i]; } \
        __device__ static __forceinline__ T cast(double v)                          \
        {                                                                           \
            double w = v;                                                           \
            if (w < (double)(LO)) w = (double)(LO);                                 \
            if (w > (double)(HI)) w = (double)(HI);                                 \
            return (T)w;                                                            \
        }                                                                           \
        static double cast_host(double v) { return v < (double)(LO) ? (double)(LO) : (v > (double)(HI) ? (double)(HI) : v); } \
    };
B200_INT_PX(int8_t, INT8_MIN, INT8_MAX)
B200_INT_PX(uint8_t, 0, UINT8_MAX)
B200_INT_PX(int16_t, INT16_MIN, INT16_MAX)
B200_INT_PX(uint16_t, 0, UINT16_MAX)
B200_INT_PX(int32_t, INT32_MIN, INT32_MAX)
B200_INT_PX(uint32_t, 0, UINT32_MAX)
#undef B200_INT_PX
template <>
struct Px<int64_t> {
    __device__ static __forceinline__ double ld(const int64_t* p, size_t i) { return (double)p[i]; }
    static double cast_host(double v) { return v < -9.2e18 ? -9.2e18 : (v > 9.2e18 ? 9.2e18 : v); }
    __device__ static __forceinline__ int64_t cast(double v)
    {
        if (v <= -9223372036854775808.0) return INT64_MIN;
        if (v >= 9223372036854775808.0) return INT64_MAX;
        return (int64_t)v;
    }
};
template <>
struct Px<uint64_t> {
    __device__ static __forceinline__ double ld(const uint64_t* p, size_t i) { return (double)p[i]; }
    static double cast_host(double v) { return v < 0 ? 0 : (v > 1.8e19 ? 1.8e19 : v); }
    __device__ static __forceinline__ uint64_t cast(double v)
    {
        if (v <= 0.0) return 0;
        if (v >= 18446744073709551616.0) return UINT64_MAX;
        return (uint64_t)v;
    }
};
template <>
struct Px<float> {
    __device__ static __forceinline__ double ld(const float* p, size_t i) { return (double)p[i]; }
    static double cast_host(double v) { return v; }
    __device__ static __forceinline__ float cast(double v)
    {
        double w = v;
        if (w < -(double)FLT_MAX) w = -(double)FLT_MAX;
        if (w > (double)FLT_MAX) w = (double)FLT_MAX;
        return (float)w;
    }
};
template <>
struct Px<double> {
    __device__ static __forceinline__ double ld(const double* p, size_t i) { return p[i]; }
    static double cast_host(double v) { return v; }
    __device__ static __forceinline__ double cast(double v) { return v; }
};

inline size_t dtype_size(int dt)
{
    switch (dt) {
    case B200REG_I8: case B200REG_U8: return 1;
    case B200REG_I16: case B200REG_U16: return 2;
    case B200REG_I32: case B200REG_U32: case B200REG_F32: return 4;
    case B200REG_I64: case B200REG_U64: case B200REG_F64: return 8;
    default: return 0;
    }
}

// dispatch a generic lambda on the pixel type: f(T{}) with T the C type
#define B200_DISPATCH_DTYPE(dt, NAME, ...)                                       \
    switch (dt) {                                                                \
    case B200REG_I8: { using NAME = int8_t; __VA_ARGS__; } break;                \
    case B200REG_U8: { using NAME = uint8_t; __VA_ARGS__; } break;               \
    case B200REG_I16: { using NAME = int16_t; __VA_ARGS__; } break;              \
    case B200REG_U16: { using NAME = uint16_t; __VA_ARGS__; } break;             \
    case B200REG_I32: { using NAME = int32_t; __VA_ARGS__; } break;              \
    case B200REG_U32: { using NAME = uint32_t; __VA_ARGS__; } break;             \
    case B200REG_I64: { using NAME = int64_t; __VA_ARGS__; } break;              \
    case B200REG_U64: { using NAME = uint64_t; __VA_ARGS__; } break;             \
    case B200REG_F32: { using NAME = float; __VA_ARGS__; } break;                \
    case B200REG_F64: { using NAME = double; __VA_ARGS__; } break;               \
    default: return ::b200::set_error(B200REG_ERR_ARG, "unsupported pixel type %d", (int)(dt)); \
    }

// 3-D launch shape used by the per-voxel kernels: x fastest, 64x4 threads per block
constexpr int BX = 64, BY = 4;
inline dim3 grid3(int nx, int ny, int nz) { return dim3((nx + BX - 1) / BX, (ny + BY - 1) / BY, nz); }
inline dim3 block3() { return dim3(BX, BY, 1); }

// warp / block reductions (fixed order -> deterministic)
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace b200
