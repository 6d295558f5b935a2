// mumemto_b200/csrc/prims.cu -- device-wide radix sort and scans.
// Round 1: these wrap CUB (the CUDA toolkit's own primitives, compiled here for sm_100a); the domain kernels that call them are
// hand-written.  The unit-test build replaces them with the obvious sequential loops.
#include "prims.h"

#ifdef MMB_SIMT
#include <algorithm>
#include <numeric>
namespace mmb { namespace prim {
template <class V>
static void sort_impl(const u64* ki, u64* ko, const V* vi, V* vo, size_t n, int b, int e) {
    std::vector<size_t> idx(n);
    std::iota(idx.begin(), idx.end(), (size_t)0);
    u64 mask = (e - b >= 64) ? ~0ull : (((1ull << (e - b)) - 1) << b);
    std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t c) { return (ki[a] & mask) < (ki[c] & mask); });
    for (size_t i = 0; i < n; ++i) { ko[i] = ki[idx[i]]; vo[i] = vi[idx[i]]; }
}
void sort_pairs(const u64* ki, u64* ko, const u32* vi, u32* vo, size_t n, int b, int e, cudaStream_t) { sort_impl(ki, ko, vi, vo, n, b, e); }
void sort_pairs64(const u64* ki, u64* ko, const u64* vi, u64* vo, size_t n, int b, int e, cudaStream_t) { sort_impl(ki, ko, vi, vo, n, b, e); }
void exclusive_sum(const u32* in, u32* out, size_t n, cudaStream_t) { u32 a = 0; for (size_t i = 0; i < n; ++i) { u32 v = in[i]; out[i] = a; a += v; } }
void exclusive_sum(const u32* in, u64* out, size_t n, cudaStream_t) { u64 a = 0; for (size_t i = 0; i < n; ++i) { u32 v = in[i]; out[i] = a; a += v; } }
void exclusive_sum(const u64* in, u64* out, size_t n, cudaStream_t) { u64 a = 0; for (size_t i = 0; i < n; ++i) { u64 v = in[i]; out[i] = a; a += v; } }
void inclusive_max(const u32* in, u32* out, size_t n, cudaStream_t) { u32 a = 0; for (size_t i = 0; i < n; ++i) { a = std::max(a, in[i]); out[i] = a; } }
void inclusive_max(const u64* in, u64* out, size_t n, cudaStream_t) { u64 a = 0; for (size_t i = 0; i < n; ++i) { a = std::max(a, in[i]); out[i] = a; } }
}}  // namespace mmb::prim
#else
#include <cub/cub.cuh>
namespace mmb { namespace prim {

struct Temp {
    void* p = nullptr; cudaStream_t s;
    Temp(size_t bytes, cudaStream_t st) : s(st) {
        if (getenv("MMB_DEBUG")) fprintf(stderr, "[prims] temp storage %zu bytes\n", bytes);
        p = pool_alloc(bytes ? bytes : 1, st);
    }
    ~Temp() { pool_free(p, s); }
};
struct MaxOp { template <class T> __device__ __forceinline__ T operator()(const T& a, const T& b) const { return a > b ? a : b; } };

template <class V>
static void sort_impl(const u64* ki, u64* ko, const V* vi, V* vo, size_t n, int b, int e, cudaStream_t s) {
    if (n == 0) return;
    size_t bytes = 0;
    MMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ki, ko, vi, vo, (int64_t)n, b, e, s));
    Temp t(bytes, s);
    ++g_kernel_launches;
    profile_before("cub_radix_sort", s);
    MMB_CUDA(cub::DeviceRadixSort::SortPairs(t.p, bytes, ki, ko, vi, vo, (int64_t)n, b, e, s));
    profile_after("cub_radix_sort", s);
}
void sort_pairs(const u64* ki, u64* ko, const u32* vi, u32* vo, size_t n, int b, int e, cudaStream_t s) { sort_impl(ki, ko, vi, vo, n, b, e, s); }
void sort_pairs64(const u64* ki, u64* ko, const u64* vi, u64* vo, size_t n, int b, int e, cudaStream_t s) { sort_impl(ki, ko, vi, vo, n, b, e, s); }

template <class I, class O>
static void exsum_impl(const I* in, O* out, size_t n, cudaStream_t s) {
    if (n == 0) return;
    size_t bytes = 0;
    MMB_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, bytes, in, out, cub::Sum(), (O)0, (int64_t)n, s));
    Temp t(bytes, s);
    ++g_kernel_launches;
    profile_before("cub_scan", s);
    MMB_CUDA(cub::DeviceScan::ExclusiveScan(t.p, bytes, in, out, cub::Sum(), (O)0, (int64_t)n, s));
    profile_after("cub_scan", s);
}
void exclusive_sum(const u32* in, u32* out, size_t n, cudaStream_t s) { exsum_impl(in, out, n, s); }
void exclusive_sum(const u32* in, u64* out, size_t n, cudaStream_t s) { exsum_impl(in, out, n, s); }
void exclusive_sum(const u64* in, u64* out, size_t n, cudaStream_t s) { exsum_impl(in, out, n, s); }

template <class T>
static void incmax_impl(const T* in, T* out, size_t n, cudaStream_t s) {
    if (n == 0) return;
    size_t bytes = 0;
    MMB_CUDA(cub::DeviceScan::InclusiveScan(nullptr, bytes, in, out, MaxOp(), (int64_t)n, s));
    Temp t(bytes, s);
    ++g_kernel_launches;
    profile_before("cub_scan", s);
    MMB_CUDA(cub::DeviceScan::InclusiveScan(t.p, bytes, in, out, MaxOp(), (int64_t)n, s));
    profile_after("cub_scan", s);
}
void inclusive_max(const u32* in, u32* out, size_t n, cudaStream_t s) { incmax_impl(in, out, n, s); }
void inclusive_max(const u64* in, u64* out, size_t n, cudaStream_t s) { incmax_impl(in, out, n, s); }

}}  // namespace mmb::prim
#endif
