// cuda_runtime.h (host stand-in) -- TEST INFRASTRUCTURE, part of oracle/ref_shim.
//
// Lets g++ compile the REFERENCE's own CUDA translation units (kfusion/src/cuda/{tsdf_volume,imgproc,proj_icp}.cu, read from
// /root/reference where they lie) for the host, so that their per-thread kernels can be executed on the CPU and used to pin the
// oracle (oracle/orc_*.c).  nvcc 12.9 cannot build those files for a GPU any more (texture references were removed in CUDA 12).
//
// What this emulates: vector types, the runtime calls the three files make (malloc/memcpy over host memory), legacy texture
// references with point filtering (element and half channel formats), the math intrinsics they use, and kernel launches
// (oracle/ref_shim/Makefile rewrites `k<<<g, b>>>(args);` into cudahost::launch(...) on the fly; nothing is copied to disk).
// A launch runs the CUDA threads of every block ONE AFTER ANOTHER: exact for kernels whose threads do not communicate
// (integrate, raycast, project, extract_normals, all of imgproc).  Kernels that synchronise or rely on warp-synchronous
// shared memory (icp_helper_kernel's block reduction, extract_kernel's warp compaction) compile but abort if launched;
// kfref.cpp calls their per-thread device functions directly instead.
//
// Intrinsic semantics on the host: __fmaf_rn = fmaf, __fsqrt_rn = sqrtf, __float2int_rn = nearbyint (ties to even),
// __float2half_rn / __half2float = IEEE binary16 (F16C).  The APPROXIMATE GPU intrinsics have no bit-defined host equivalent
// and are mapped to the correctly rounded operation: __fdividef(a, b) = a / b, __expf = expf, __powf = powf, __sinf/__cosf,
// rsqrtf(x) = 1 / sqrtf(x).  That is the same numerics contract the oracle and the sm_100a kernels follow.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <immintrin.h>

#define __host__
#define __device__
#define __global__
#define __shared__ static
#define __constant__ static
#ifdef CUDAHOST_LOCKSTEP
#define __forceinline__ inline __attribute__((always_inline))      /* lock-step mode: machine code order must be source order (lockstep.h) */
#else
#define __forceinline__ inline
#endif
#define __launch_bounds__(...)
#define __restrict__

// ---- vector types ------------------------------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct alignas(4) ushort2 { unsigned short x, y; };
struct alignas(4) short2 { short x, y; };
struct uchar3 { unsigned char x, y, z; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { float2 v = {x, y}; return v; }
static inline float3 make_float3(float x, float y, float z) { float3 v = {x, y, z}; return v; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
static inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
static inline int3 make_int3(int x, int y, int z) { int3 v = {x, y, z}; return v; }
static inline ushort2 make_ushort2(unsigned short x, unsigned short y) { ushort2 v; v.x = x; v.y = y; return v; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { uchar4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
static inline uchar3 make_uchar3(unsigned char x, unsigned char y, unsigned char z) { uchar3 v = {x, y, z}; return v; }

// ---- thread coordinates (one emulated CUDA thread runs at a time) -----------------------------------------------------
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
static const int warpSize = 32;

#define CUDART_NAN_F (__builtin_nanf(""))
#define CUDART_INF_F (__builtin_inff())

// ---- runtime -------------------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef cudaError_t cudaError;
enum { cudaSuccess = 0 };
typedef void *cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncCache { cudaFuncCachePreferNone, cudaFuncCachePreferShared, cudaFuncCachePreferL1 };

static inline const char *cudaGetErrorString(cudaError_t) { return "cudahost error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocPitch(void **p, size_t *pitch, size_t width_bytes, size_t rows)
{
    *pitch = (width_bytes + 511) & ~(size_t)511;
    return cudaMalloc(p, *pitch * rows);
}
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind)
{
    for (size_t r = 0; r < h; ++r) std::memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMemcpyFromSymbol(void *d, const T &symbol, size_t n) { std::memcpy(d, &symbol, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetCacheConfig(F, cudaFuncCache) { return cudaSuccess; }

// ---- legacy texture references (point filtering only) ---------------------------------------------------------------
enum cudaTextureReadMode { cudaReadModeElementType, cudaReadModeNormalizedFloat };
enum cudaTextureFilterMode { cudaFilterModePoint, cudaFilterModeLinear };
enum cudaTextureAddressMode { cudaAddressModeWrap, cudaAddressModeClamp, cudaAddressModeMirror, cudaAddressModeBorder };
enum cudaChannelFormatKind { cudaChannelFormatKindSigned, cudaChannelFormatKindUnsigned, cudaChannelFormatKindFloat };
struct cudaChannelFormatDesc { int x, y, z, w; cudaChannelFormatKind f; };
template <class T> static inline cudaChannelFormatDesc cudaCreateChannelDesc()
{
    cudaChannelFormatDesc d = {(int)sizeof(T) * 8, 0, 0, 0, cudaChannelFormatKindUnsigned};
    return d;
}
static inline cudaChannelFormatDesc cudaCreateChannelDescHalf()
{
    cudaChannelFormatDesc d = {16, 0, 0, 0, cudaChannelFormatKindFloat};
    return d;
}
struct textureReference {
    int normalized;
    cudaTextureFilterMode filterMode;
    cudaTextureAddressMode addressMode[3];
    cudaChannelFormatDesc channelDesc;
    mutable const void *ptr;
    mutable size_t pitch;
    mutable int width, height;
    mutable bool half_elems;
};
template <class T, int dim = 1, cudaTextureReadMode mode = cudaReadModeElementType> struct texture : public textureReference {
    texture(int norm = 0, cudaTextureFilterMode f = cudaFilterModePoint, cudaTextureAddressMode a = cudaAddressModeClamp)
    {
        normalized = norm; filterMode = f; addressMode[0] = addressMode[1] = addressMode[2] = a; channelDesc = cudaCreateChannelDesc<T>();
        ptr = 0; pitch = 0; width = height = 0; half_elems = false;
    }
    texture(int norm, cudaTextureFilterMode f, cudaTextureAddressMode a, cudaChannelFormatDesc desc)
    {
        normalized = norm; filterMode = f; addressMode[0] = addressMode[1] = addressMode[2] = a; channelDesc = desc;
        ptr = 0; pitch = 0; width = height = 0; half_elems = false;
    }
};
template <class T, cudaTextureReadMode mode>
static inline cudaError_t cudaBindTexture2D(size_t *offset, const texture<T, 2, mode> &tex, const void *ptr, const cudaChannelFormatDesc &desc, size_t width,
                                            size_t height, size_t pitch)
{
    if (offset) *offset = 0;
    tex.ptr = ptr; tex.pitch = pitch; tex.width = (int)width; tex.height = (int)height;
    tex.half_elems = desc.f == cudaChannelFormatKindFloat && desc.x == 16;
    return cudaSuccess;
}
template <class T, cudaTextureReadMode mode>
static inline cudaError_t cudaBindTexture(size_t *offset, const texture<T, 1, mode> &tex, const void *ptr, const cudaChannelFormatDesc &, size_t bytes)
{
    if (offset) *offset = 0;
    tex.ptr = ptr; tex.pitch = bytes; tex.width = (int)(bytes / sizeof(T)); tex.height = 1; tex.half_elems = false;
    return cudaSuccess;
}
static inline cudaError_t cudaUnbindTexture(const textureReference *t) { t->ptr = 0; return cudaSuccess; }

static inline float __half2float(unsigned short h) { return _cvtsh_ss(h); }
static inline unsigned short __float2half_rn(float f) { return _cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC); }

namespace cudahost {
template <class T> struct TexFetch {
    static T at(const textureReference &t, int x, int y) { return *(const T *)((const char *)t.ptr + (size_t)y * t.pitch + (size_t)x * sizeof(T)); }
    static T zero() { T v; std::memset(&v, 0, sizeof(T)); return v; }
};
template <> struct TexFetch<float> {
    static float at(const textureReference &t, int x, int y)
    {
        if (t.half_elems) return __half2float(*(const unsigned short *)((const char *)t.ptr + (size_t)y * t.pitch + (size_t)x * 2));
        return *(const float *)((const char *)t.ptr + (size_t)y * t.pitch + (size_t)x * 4);
    }
    static float zero() { return 0.f; }
};
// unnormalised coordinates, point filtering: texel = floor(coordinate) (CUDA programming guide, texture fetching)
static inline bool resolve(const textureReference &t, float x, float y, int &ix, int &iy)
{
    if (t.filterMode != cudaFilterModePoint || t.normalized || !t.ptr) { std::fprintf(stderr, "cudahost: unsupported texture fetch\n"); std::abort(); }
    ix = (int)std::floor(x); iy = (int)std::floor(y);
    const bool inside = ix >= 0 && iy >= 0 && ix < t.width && iy < t.height;
    if (inside) return true;
    if (t.addressMode[0] == cudaAddressModeBorder) return false;
    ix = ix < 0 ? 0 : (ix >= t.width ? t.width - 1 : ix);          // clamp (wrap/mirror need normalised coordinates)
    iy = iy < 0 ? 0 : (iy >= t.height ? t.height - 1 : iy);
    return true;
}
}  // namespace cudahost
template <class T, cudaTextureReadMode mode> static inline T tex2D(const texture<T, 2, mode> &t, float x, float y)
{
    int ix, iy;
    return cudahost::resolve(t, x, y, ix, iy) ? cudahost::TexFetch<T>::at(t, ix, iy) : cudahost::TexFetch<T>::zero();
}

// ---- device intrinsics -------------------------------------------------------------------------------------------------
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __fdividef(float a, float b) { return a / b; }
namespace cudahost {
static inline float expf_(float a) { return std::exp(a); }
static inline float powf_(float a, float b) { return std::pow(a, b); }
static inline float sinf_(float a) { return std::sin(a); }
static inline float cosf_(float a) { return std::cos(a); }
}
#define __expf(a) cudahost::expf_(a)          /* glibc's <math.h> already declares __expf & co. as internal externs */
#define __powf(a, b) cudahost::powf_(a, b)
#define __sinf(a) cudahost::sinf_(a)
#define __cosf(a) cudahost::cosf_(a)
static inline float rsqrtf(float a) { return 1.f / std::sqrt(a); }
static inline float rsqrt(float a) { return 1.f / std::sqrt(a); }
static inline float __saturatef(float a) { return a != a ? 0.f : (a < 0.f ? 0.f : (a > 1.f ? 1.f : a)); }
static inline int __float2int_rn(float a) { return (int)std::nearbyint(a); }     // default rounding mode: to nearest, ties to even
static inline int __float2int_rd(float a) { return (int)std::floor(a); }
static inline int __float2int_rz(float a) { return (int)a; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline float min(float a, float b) { return std::fmin(a, b); }
static inline float max(float a, float b) { return std::fmax(a, b); }
using std::abs;
using std::fabs;          // CUDA resolves fabs(float) to the float overload; plain C ::fabs would silently promote to double
using std::isnan;
using std::isinf;

namespace cudahost {
[[noreturn]] static inline void not_emulated(const char *what)
{
    std::fprintf(stderr, "cudahost: %s needs inter-thread communication, which the sequential host emulation does not provide\n", what);
    std::abort();
}
}
static inline void __syncthreads() { cudahost::not_emulated("__syncthreads"); }
#ifdef CUDAHOST_LOCKSTEP
#include "lockstep.h"
static inline unsigned __ballot(int p) { return cudahost::warp_vote(p); }
static inline int __all(int p) { unsigned active; const unsigned m = cudahost::warp_vote(p, &active); return m == active; }
static inline int __any(int p) { return cudahost::warp_vote(p) != 0u; }
#else
static inline unsigned __ballot(int) { cudahost::not_emulated("__ballot"); }
static inline int __all(int) { cudahost::not_emulated("__all"); }
static inline int __any(int) { cudahost::not_emulated("__any"); }
#endif
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned atomicInc(unsigned *p, unsigned lim) { unsigned o = *p; *p = o >= lim ? 0 : o + 1; return o; }

// ---- kernel launch: every thread of every block, sequentially ---------------------------------------------------------
namespace cudahost {
struct LaunchCfg {
    dim3 grid, block;
    LaunchCfg(dim3 g, dim3 b, size_t = 0, cudaStream_t = 0) : grid(g), block(b) {}
};
#ifdef CUDAHOST_LOCKSTEP
// warp-lock-step launch (lockstep.h): blocks in launch order, the warps of a block one after another, 32 lanes per warp as fibers
template <class F> static inline void launch(const LaunchCfg &c, F body)
{
    gridDim = c.grid; blockDim = c.block;
    struct Thunk { static void call(void *p) { (*(F *)p)(); } };
    const unsigned per_block = c.block.x * c.block.y * c.block.z;
    for (unsigned bz = 0; bz < c.grid.z; ++bz)
        for (unsigned by = 0; by < c.grid.y; ++by)
            for (unsigned bx = 0; bx < c.grid.x; ++bx) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                // warps in DESCENDING order: extract_kernel's thread 0 publishes the point count in an epilogue that no block barrier
                // protects (tsdf_volume.cu:690-703); any interleaving is a legal GPU schedule, and this is the one under which that
                // epilogue sees the points of every warp of its block -- the outcome the reference relies on
                for (long first = (long)((per_block - 1) / 32) * 32; first >= 0; first -= 32) {
                    uint3 tids[32];
                    int n = 0;
                    for (unsigned f = (unsigned)first; f < per_block && f < (unsigned)first + 32; ++f, ++n) {
                        tids[n].x = f % c.block.x; tids[n].y = (f / c.block.x) % c.block.y; tids[n].z = f / (c.block.x * c.block.y);
                    }
                    run_warp(tids, n, &Thunk::call, (void *)&body);
                }
            }
}
#else
template <class F> static inline void launch(const LaunchCfg &c, F body)
{
    gridDim = c.grid; blockDim = c.block;
    for (unsigned bz = 0; bz < c.grid.z; ++bz)
        for (unsigned by = 0; by < c.grid.y; ++by)
            for (unsigned bx = 0; bx < c.grid.x; ++bx) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                for (unsigned tz = 0; tz < c.block.z; ++tz)
                    for (unsigned ty = 0; ty < c.block.y; ++ty)
                        for (unsigned tx = 0; tx < c.block.x; ++tx) {
                            threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
                            body();
                        }
            }
}
#endif
}  // namespace cudahost

// inline PTX (Warp::laneId, gmem::LdCs under __CUDA_ARCH__) has no host meaning; only kernels that are never launched here use it
#ifdef CUDAHOST_LOCKSTEP
// the only PTX the lock-step kernels execute reads a special register into a local named `ret` (Warp::laneId / laneMaskLt,
// temp_utils.hpp:462-484); other asm statements (gmem::LdCs / StCs) have no such local, bind the dummy below and abort if executed
static unsigned ret;
#define asm(...) (ret = cudahost::ptx_special(#__VA_ARGS__))
#else
#define asm(...) cudahost::not_emulated("inline PTX")
#endif
