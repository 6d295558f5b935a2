// mumemto_b200/csrc/common.cuh -- shared definitions for the sm_100a engine.
//
// Everything under csrc/ is compiled by nvcc for sm_100a into libmumemto_b200.so.  The only other way these sources are ever
// compiled is the unit-test build under tests/simt/ (-DMMB_SIMT), which runs the kernel bodies on CPU fibers so that index logic
// can be tested in a container without a GPU; that build is never shipped or loaded by the product.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef MMB_SIMT
#include "simt.h"
#else
#include <cuda_runtime.h>
#endif

namespace mmb {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

extern std::atomic<unsigned long long> g_kernel_launches;   // every MMB_LAUNCH, for bench.py's "gpu_launches"
extern double g_host_alloc_ms, g_host_free_ms, g_host_sync_ms;      // host time inside cudaMallocAsync / cudaFreeAsync / stream syncs
extern unsigned long long g_host_allocs, g_host_syncs;
inline double host_now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

inline void check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess)
        throw Error(10, std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + file + ":" + std::to_string(line) + " (" + what + ")");
}
#define MMB_CUDA(x) ::mmb::check((x), #x, __FILE__, __LINE__)

// Optional per-kernel timing for bench.py's roofline: launches whose name contains the selected substring are bracketed by CUDA
// events on their own stream (no synchronisation); mmb_profile_result() adds the pairs up afterwards.
void profile_before(const char* kernel_name, cudaStream_t s);
void profile_after(const char* kernel_name, cudaStream_t s);

#ifndef MMB_SIMT
#define MMB_LAUNCH(kernel, grid, block, smem, stream, ...)                       \
    do {                                                                         \
        ++::mmb::g_kernel_launches;                                              \
        ::mmb::profile_before(#kernel, (stream));                                \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
        ::mmb::profile_after(#kernel, (stream));                                 \
        ::mmb::check(cudaPeekAtLastError(), #kernel, __FILE__, __LINE__);        \
    } while (0)
#define MMB_DYN_SMEM(type, name)                                  \
    extern __shared__ __align__(16) unsigned char name##_raw_[];  \
    type* name = reinterpret_cast<type*>(name##_raw_)
#endif

// a 16-byte aligned static shared-memory array (the unit-test build spells the attribute order differently)
#ifdef MMB_SIMT
#define MMB_SHARED_ALIGNED16(type, name, count) alignas(16) static type name[count]
#else
#define MMB_SHARED_ALIGNED16(type, name, count) __shared__ __align__(16) type name[count]
#endif

inline u32 ceil_div(u64 a, u64 b) { return (u32)((a + b - 1) / b); }

// MMB_DEBUG=1: wall-clock of a host-side phase (synchronises the stream at both ends; never active in measured runs)
struct PhaseTimer {
    const char* name; cudaStream_t s; bool on; double t0 = 0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    PhaseTimer(const char* n, cudaStream_t st) : name(n), s(st), on(getenv("MMB_DEBUG") != nullptr) { if (on) { cudaStreamSynchronize(s); t0 = now(); } }
    ~PhaseTimer() { if (on) { cudaStreamSynchronize(s); fprintf(stderr, "[phase] %-28s %9.3f ms\n", name, now() - t0); } }
};

// Device memory for the pipeline's many short-lived arrays.  All work of an engine is ordered on one stream, so a block released by
// one stage can be handed to the next stage at once; blocks are cached per process and only go back to the driver when an allocation
// fails or the engine is destroyed.  (cudaMallocAsync's pool showed 15..540 ms of allocator time per call of mmb_find on the tiled
// route -- see profiles/README.md -- which is why the engine keeps its own free lists.)
void* pool_alloc(size_t bytes, cudaStream_t s);
void pool_free(void* p, cudaStream_t s);
void pool_trim();          // give every cached block back to the driver

// Stream-ordered device buffer.
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream; n = count;
        double t0 = host_now_ms();
        p = (T*)pool_alloc((count ? count : 1) * sizeof(T), stream);
        g_host_alloc_ms += host_now_ms() - t0; ++g_host_allocs;
    }
    void zero() { if (n) MMB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void release() { if (p) { double t0 = host_now_ms(); pool_free(p, s); g_host_free_ms += host_now_ms() - t0; p = nullptr; n = 0; } }
    size_t bytes() const { return n * sizeof(T); }
    void swap(DevBuf& o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(s, o.s); }
};

template <class T>
inline std::vector<T> to_host(const T* d, size_t n, cudaStream_t s) {
    std::vector<T> h(n);
    double t0 = host_now_ms();
    if (n) MMB_CUDA(cudaMemcpyAsync(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaStreamSynchronize(s));
    g_host_sync_ms += host_now_ms() - t0; ++g_host_syncs;
    return h;
}

}  // namespace mmb
