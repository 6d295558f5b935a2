// mumemto_b200/csrc/detect.cu -- multi-MUM / multi-MEM interval detection over the (SA, LCP, BWT) stream.
//
// Replaces mem_finder::update / update_mems / check_doc_range / check_bwt_range (reference include/mem_finder.hpp:161-170,
// 304-355, 265-289, 189-192) -- rows E, F and I of SURVEY.md section 8.  The reference keeps a stack of open LCP intervals and pops
// the ones that close at rank j; the set popped at j is exactly the chain of previous-smaller-values walked leftwards from j-1, so
// every rank j with lcp[j-1] > lcp[j] is handled by one thread that walks left at most max_freq+1 ranks:
//   v = lcp[j-1]; extend left while lcp[k] >= v; the interval [k, j-1] with length v closes here; continue with v = lcp[k] while
//   v > lcp[j] and v >= min_len.  Document multiplicities accumulate in bit-sliced counters as the walk widens, the BWT
//   left-maximality flag is the OR of bwt[k] != bwt[k-1] over the pairs absorbed so far, and the next-best threshold of -M is
//   max(lcp[k], lcp[j]).  Intervals that never close (they touch the last rank) are never reported, like the reference (no flush).
// Reports are appended with an atomic counter and sorted afterwards by (closing rank, order in chain) = the reference's order.
// Traffic per suffix: lcp 4 B + bwt 1 B + sa 4/8 B for walked ranks (L1/L2 hits between neighbouring threads).
#include "detect.h"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 TPB = 128;

__device__ __forceinline__ u32 find_doc(const u64* doc_start, u32 ndocs, u64 p) {
    u32 lo = 0, hi = ndocs;                 // doc_start[lo] <= p < doc_start[hi]
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (doc_start[mid] <= p) lo = mid; else hi = mid; }
    return lo;
}

// W = 64-bit words per plane (ndocs <= 64 W), P = planes (counts 0 .. 2^P - 1)
template <int W, int P>
struct DocCounter {
    u64 plane[P][W];
    u32 distinct;
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int w = 0; w < W; ++w) plane[p][w] = 0;
        distinct = 0;
    }
    // returns false when document d would exceed max_doc_freq (0 = unlimited)
    __device__ __forceinline__ bool add(u32 d, u32 max_doc_freq) {
        u32 w = (W == 1) ? 0u : (d >> 6);
        u64 bit = 1ull << (d & 63);
        u32 c = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) c |= (u32)((plane[p][w] & bit) != 0) << p;
        if (c == 0) ++distinct;
        if (max_doc_freq == 0) { plane[0][w] |= bit; return true; }
        if (c >= max_doc_freq) return false;
#pragma unroll
        for (int p = 0; p < P; ++p) {          // ripple-carry increment
            bool was = (plane[p][w] & bit) != 0;
            plane[p][w] ^= bit;
            if (!was) break;
        }
        return true;
    }
};

template <class SAT, int W, int P>
__global__ void __launch_bounds__(TPB) detect_kernel(const u32* __restrict__ lcp, const SAT* __restrict__ sa, const u8* __restrict__ bwt, u64 m,
                                                     DetectParams prm, const u64* __restrict__ doc_start_g, Report* __restrict__ reports, u32 cap,
                                                     u32* __restrict__ count, u16* __restrict__ thresh) {
    MMB_DYN_SMEM(u64, doc_start);
    for (u32 i = threadIdx.x; i <= prm.ndocs; i += blockDim.x) doc_start[i] = doc_start_g[i];
    __syncthreads();
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (j >= m) return;
    u32 cur = lcp[j], v = lcp[j - 1];
    if (v <= cur || v < prm.min_len) return;

    DocCounter<W, P> docs; docs.clear();
    u64 k = j - 1;                  // lowest rank inside the current interval
    u64 anchor_sa = ~0ull;          // SA of the lowest-ranked document-0 suffix seen
    u32 size = 1, order = 0;
    bool left_differs = false;
    {
        u64 p = (u64)sa[k]; u32 d = find_doc(doc_start, prm.ndocs, p);
        if (d == 0) anchor_sa = p;
        if (!docs.add(d, prm.max_doc_freq)) return;
    }
    for (;;) {
        while (lcp[k] >= v) {       // rank k-1 shares v symbols with rank k: absorb it (k >= 2 here because lcp[0] = lcp[1] = 0 < v)
            left_differs |= (bwt[k] != bwt[k - 1]);
            --k; ++size;
            if (prm.max_freq && size > prm.max_freq) return;
            u64 p = (u64)sa[k]; u32 d = find_doc(doc_start, prm.ndocs, p);
            if (d == 0) anchor_sa = p;
            if (!docs.add(d, prm.max_doc_freq)) return;
        }
        u32 before = lcp[k];
        if (size >= prm.num_distinct && docs.distinct >= prm.num_distinct) {
            if (prm.merge && anchor_sa != ~0ull) thresh[anchor_sa] = (u16)min(max(before, cur), 65535u);
            if (left_differs) {
                u32 slot = atomicAdd(count, 1u);
                if (slot < cap) { Report r; r.key = (j << 20) | order; r.s = k; r.len = v; r.pad = 0; reports[slot] = r; }
                ++order;
            }
        }
        v = before;
        if (v <= cur || v < prm.min_len) return;
    }
}

__global__ void report_keys(const Report* __restrict__ rep, u32 n, u64* __restrict__ keys, u32* __restrict__ idx) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = rep[i].key; idx[i] = i; }
}
__global__ void report_gather(const Report* __restrict__ rep, const u32* __restrict__ idx, u32 n, Report* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rep[idx[i]];
}

template <class SAT, int W, int P>
void launch_detect(const u32* lcp, const SAT* sa, const u8* bwt, u64 m, const DetectParams& prm, const u64* d_doc_start, Report* rep, u32 cap,
                   u32* count, u16* thresh, cudaStream_t s) {
    auto k = detect_kernel<SAT, W, P>;
    size_t smem = (size_t)(prm.ndocs + 1) * sizeof(u64);
    MMB_LAUNCH(k, ceil_div(m - 1, TPB), TPB, smem, s, lcp, sa, bwt, m, prm, d_doc_start, rep, cap, count, thresh);
}

template <class SAT, int W>
void dispatch_planes(const u32* lcp, const SAT* sa, const u8* bwt, u64 m, const DetectParams& prm, const u64* d_doc_start, Report* rep, u32 cap,
                     u32* count, u16* thresh, cudaStream_t s) {
    u32 f = prm.max_doc_freq;
    if (f <= 1) launch_detect<SAT, W, 1>(lcp, sa, bwt, m, prm, d_doc_start, rep, cap, count, thresh, s);
    else if (f <= 3) launch_detect<SAT, W, 2>(lcp, sa, bwt, m, prm, d_doc_start, rep, cap, count, thresh, s);
    else if (f <= 255) launch_detect<SAT, W, 8>(lcp, sa, bwt, m, prm, d_doc_start, rep, cap, count, thresh, s);
    else throw Error(30, "per-sequence frequency limits above 255 are not supported");
}

}  // namespace

template <class SAT>
void detect_intervals(const u32* lcp, const SAT* sa, const u8* bwt, u64 m, const DetectParams& prm, const u64* d_doc_start, u16* thresh,
                      Reports& out, cudaStream_t s) {
    // the leftward walk is bounded by max_freq, or else by max_doc_freq * ndocs (a document overflows); with neither it is unbounded
    if (prm.max_freq == 0 && prm.max_doc_freq == 0)
        throw Error(31, "unbounded total frequency (-f 0 with -F 0) is not supported by the windowed interval scan");
    if (prm.max_freq >= (1u << 20)) throw Error(32, "maximum total frequency must be below 2^20");
    if (prm.ndocs > 1024) throw Error(33, "more than 1024 input sequences are not supported");
    out.count = 0;
    if (m < 2) { out.rep.alloc(0, s); return; }
    DevBuf<u32> d_count; d_count.alloc(1, s);
    u64 cap = std::max<u64>(1u << 16, m / 64);
    DevBuf<Report> raw;
    for (;;) {
        if (cap > 0xfffffff0ull) throw Error(34, "too many matches for one pass");
        raw.alloc(cap, s); d_count.zero();
        if (prm.ndocs <= 64) dispatch_planes<SAT, 1>(lcp, sa, bwt, m, prm, d_doc_start, raw.p, (u32)cap, d_count.p, thresh, s);
        else if (prm.ndocs <= 256) dispatch_planes<SAT, 4>(lcp, sa, bwt, m, prm, d_doc_start, raw.p, (u32)cap, d_count.p, thresh, s);
        else dispatch_planes<SAT, 16>(lcp, sa, bwt, m, prm, d_doc_start, raw.p, (u32)cap, d_count.p, thresh, s);
        u32 got = to_host(d_count.p, 1, s)[0];
        if (got <= cap) { out.count = got; break; }
        cap = (u64)got + 1024;          // the counter kept counting past the capacity: one retry with the exact size
    }
    u32 n = (u32)out.count;
    out.rep.alloc(n, s);
    if (n == 0) return;
    DevBuf<u64> ka, kb; DevBuf<u32> ia, ib;
    ka.alloc(n, s); kb.alloc(n, s); ia.alloc(n, s); ib.alloc(n, s);
    MMB_LAUNCH(report_keys, ceil_div(n, 256), 256, 0, s, raw.p, n, ka.p, ia.p);
    prim::sort_pairs(ka.p, kb.p, ia.p, ib.p, n, 0, 64, s);
    MMB_LAUNCH(report_gather, ceil_div(n, 256), 256, 0, s, raw.p, ib.p, n, out.rep.p);
}

template void detect_intervals<u32>(const u32*, const u32*, const u8*, u64, const DetectParams&, const u64*, u16*, Reports&, cudaStream_t);
template void detect_intervals<u64>(const u32*, const u64*, const u8*, u64, const DetectParams&, const u64*, u16*, Reports&, cudaStream_t);

}  // namespace mmb
