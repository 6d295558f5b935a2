// tests/simt/simt.h -- TEST INFRASTRUCTURE ONLY.  Never included, linked or loaded by the product (mumemto_b200/).
//
// A small SIMT executor that lets the engine's kernel sources (mumemto_b200/csrc/*.cu) be compiled with g++ and run on the CPU
// inside `pytest -m "not gpu"`, because the development container has no GPU.  It exists to unit-test index arithmetic, tile/halo
// logic and host orchestration of the hand-written sm_100a kernels before they go to a B200; it is not a backend: it is orders of
// magnitude slower than one CPU core running the oracle, it is built only under tests/simt/_build/, and the shipped library is
// compiled by nvcc alone and refuses to run without a CUDA device.
//
// Model: blocks run one after another; the threads of a block are ucontext fibers scheduled round-robin on one OS thread, so
// __syncthreads(), warp shuffles and ballots have their CUDA meaning (all named lanes must arrive).  `__shared__` maps to a
// function-local static (one block is live at a time).  Atomics are plain read-modify-write (one OS thread).
#pragma once
#ifndef MMB_SIMT
#error "simt.h is only for the -DMMB_SIMT test build"
#endif

#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <vector>

// ---------------------------------------------------------------------------------------------- CUDA keywords
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct ulonglong2 { unsigned long long x, y; };
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

namespace simt {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void block_barrier();
void warp_barrier(unsigned mask);
unsigned long long* warp_slot(unsigned lane);   // per-warp exchange buffer of the calling fiber's warp
extern unsigned long long g_launches;
extern int g_block_order;      // 0 forward, 1 reverse, 2 scrambled
}  // namespace simt

#define threadIdx simt::g_threadIdx
#define blockIdx simt::g_blockIdx
#define blockDim simt::g_blockDim
#define gridDim simt::g_gridDim
#define warpSize 32

// ---------------------------------------------------------------------------------------------- synchronisation and warp collectives
inline void __syncthreads() { simt::block_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_barrier(mask); }
inline void __threadfence() {}
inline void __threadfence_block() {}

namespace simt {
inline unsigned lane_id() { return (g_threadIdx.x + g_threadIdx.y * g_blockDim.x) & 31; }
template <class T>
inline T exchange(unsigned mask, T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long raw = 0; memcpy(&raw, &v, sizeof(T));
    *warp_slot(lane_id()) = raw;
    warp_barrier(mask);
    unsigned long long got = *warp_slot(src_lane & 31);
    warp_barrier(mask);
    T out; memcpy(&out, &got, sizeof(T)); return out;
}
}  // namespace simt

template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    unsigned lane = simt::lane_id(); unsigned base = lane & ~(unsigned)(width - 1);
    return simt::exchange(mask, v, base + ((unsigned)src & (unsigned)(width - 1)));
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned lane = simt::lane_id(); unsigned base = lane & ~(unsigned)(width - 1);
    unsigned src = (lane - base >= d) ? lane - d : lane;
    return simt::exchange(mask, v, src);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned lane = simt::lane_id(); unsigned base = lane & ~(unsigned)(width - 1);
    unsigned src = (lane - base + d < (unsigned)width) ? lane + d : lane;
    return simt::exchange(mask, v, src);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    (void)width; return simt::exchange(mask, v, simt::lane_id() ^ (unsigned)x);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    *simt::warp_slot(simt::lane_id()) = pred ? 1ull : 0ull;
    simt::warp_barrier(mask);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) if ((mask >> l & 1) && *simt::warp_slot(l)) r |= 1u << l;
    simt::warp_barrier(mask);
    return r;
}
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    static_assert(sizeof(T) <= 8, "match payload");
    unsigned long long raw = 0; memcpy(&raw, &v, sizeof(T));
    *simt::warp_slot(simt::lane_id()) = raw;
    simt::warp_barrier(mask);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) if ((mask >> l & 1) && *simt::warp_slot(l) == raw) r |= 1u << l;
    simt::warp_barrier(mask);
    return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    *simt::warp_slot(simt::lane_id()) = v;
    simt::warp_barrier(mask);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) if (mask >> l & 1) r |= (unsigned)*simt::warp_slot(l);
    simt::warp_barrier(mask);
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }

// ---------------------------------------------------------------------------------------------- integer intrinsics
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32); }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    unsigned long long v = ((unsigned long long)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
inline unsigned __vcmpeq4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xffu) == ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i); return r; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
using std::max;
using std::min;

// ---------------------------------------------------------------------------------------------- atomics (single OS thread)
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---------------------------------------------------------------------------------------------- runtime API subset
typedef int cudaError_t;
typedef struct simt_stream* cudaStream_t;
typedef struct simt_event { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost, cudaMemcpyDefault };
inline const char* cudaGetErrorString(cudaError_t e) { return e ? "simt error" : "no error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
enum { cudaHostAllocPortable = 1 };
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { static int dummy; *s = (cudaStream_t)&dummy; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new simt_event{0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

#define MMB_LAUNCH(kernel, grid, block, smem, stream, ...)                                                          \
    do {                                                                                                            \
        ++::mmb::g_kernel_launches;                                                                                 \
        ::mmb::profile_before(#kernel, (stream)); ::mmb::profile_after(#kernel, (stream));                          \
        auto mmb_args_ = std::make_tuple(__VA_ARGS__); /* evaluate the arguments once, like a real launch does */   \
        auto mmb_kernel_ = kernel;                                                                                  \
        simt::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { std::apply(mmb_kernel_, mmb_args_); });       \
    } while (0)
#define MMB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(simt::g_dyn_smem)
