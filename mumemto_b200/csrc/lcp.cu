// mumemto_b200/csrc/lcp.cu -- LCP array from (text, SA, ISA); BWT from (text, SA) when the arrays are dumped (-A).
//
// The reference obtains LCP from gsacak (include/direct_gsacak.hpp:62) or, on the PFP path, from lcpD / s_lcp_T range minima
// (include/pfp_lcp_mum.hpp:145,179,284-321); both equal the plain LCP of adjacent suffixes of T with '$' as an ordinary symbol.
// Here: Kasai's invariant PLCP[i+1] >= PLCP[i]-1 applied per chunk of consecutive text positions.  Each thread owns CHUNK
// positions, pays one from-scratch comparison for the first and amortised O(1) for the rest; comparisons run 8 bytes at a time.
// Traffic per suffix: isa 4 B + sa 4 B + lcp 4 B written + the compared text bytes (L2-resident neighbours) + 1 B bwt.
#include "sa.h"

namespace mmb {

namespace {

constexpr u32 TPB = 128;
constexpr u32 CHUNK = 32;

__device__ __forceinline__ u64 load8(const u8* __restrict__ t, u64 i) {      // unaligned 8-byte little-endian load
    const u32* w = reinterpret_cast<const u32*>(t + (i & ~3ull));
    u32 sh = (u32)(i & 3) * 8;
    u32 a = w[0], b = w[1], c = w[2];
    u32 lo = sh ? __funnelshift_r(a, b, sh) : a;
    u32 hi = sh ? __funnelshift_r(b, c, sh) : b;
    return ((u64)hi << 32) | lo;
}

// number of equal leading symbols of text[i..) and text[j..), given that the first `have` are known equal; never reads past n
__device__ __forceinline__ u32 extend(const u8* __restrict__ t, u64 n, u64 i, u64 j, u32 have) {
    u64 lim = n - max(i, j);                // symbols available in the shorter suffix
    u32 l = have;
    while ((u64)l + 8 <= lim) {
        u64 x = load8(t, i + l) ^ load8(t, j + l);
        if (x) return l + (u32)((__ffsll((long long)x) - 1) >> 3);
        l += 8;
    }
    while ((u64)l < lim && t[i + l] == t[j + l]) ++l;
    return l;
}

__global__ void __launch_bounds__(TPB) plcp_chunks(const u8* __restrict__ text, u64 n, const u32* __restrict__ sa, const u32* __restrict__ isa,
                                                   u32 m, u32* __restrict__ lcp) {
    u64 first = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * CHUNK;
    u32 l = 0;
    for (u32 k = 0; k < CHUNK; ++k) {
        u64 i = first + k;
        if (i >= m) return;
        u32 r = isa[i];
        if (r == 0) { lcp[0] = 0; l = 0; continue; }       // the sentinel suffix
        u64 j = sa[r - 1];
        l = (i < n && j < n) ? extend(text, n, i, j, l) : 0u;
        lcp[r] = l;
        l = l ? l - 1 : 0;
    }
}

__global__ void __launch_bounds__(256) gather_bwt(const u8* __restrict__ text, const u32* __restrict__ sa, u32 m, u8* __restrict__ bwt) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    u32 p = sa[r];
    bwt[r] = p ? text[p - 1] : (u8)0;
}

}  // namespace

void build_lcp(const u8* text, u64 n, const SuffixArray& sa, DevBuf<u32>& lcp, cudaStream_t s) {
    u32 m = sa.m;
    lcp.alloc((u64)m + 64, s);                       // 64 zero entries of slack: tile loads of the detector round up to 16 bytes
    MMB_CUDA(cudaMemsetAsync(lcp.p + m, 0, 64 * sizeof(u32), s));
    u64 threads = ((u64)m + CHUNK - 1) / CHUNK;
    MMB_LAUNCH(plcp_chunks, ceil_div(threads, TPB), TPB, 0, s, text, n, sa.sa.p, sa.isa.p, m, lcp.p);
}

void build_bwt(const u8* text, const SuffixArray& sa, DevBuf<u8>& bwt, cudaStream_t s) {
    bwt.alloc(sa.m, s);
    MMB_LAUNCH(gather_bwt, ceil_div(sa.m, 256), 256, 0, s, text, sa.sa.p, sa.m, bwt.p);
}

}  // namespace mmb
