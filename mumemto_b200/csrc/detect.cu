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
// Reports are appended with an atomic counter and sorted afterwards by (closing rank, interval size) = the reference's order.
//
// Two kernels implement this:
//   detect_window_kernel  (used when at most 64 interval sizes can qualify, i.e. F - k < 64: strict / partial MUMs, bounded MEMs).
//       An interval of size m closing at j is [j-m, j-1]; it exists iff v = min(lcp[j-m+1 .. j-1]) exceeds both lcp[j-m] and lcp[j],
//       and its length is v.  Each block stages a tile of lcp (1024 closing ranks + halo F) in shared memory with one TMA bulk copy,
//       builds power-of-two window minima there, and every thread tests its closing ranks in O(1) per size.  The few survivors are
//       verified by the whole warp: document bitmask by OR-reduction, BWT left-maximality by neighbour comparison.
//       Traffic per suffix: lcp 4 B; sa (+ text for the BWT symbol) only for surviving candidates.
//   detect_walk_kernel    (any other parameters) the leftward walk described above, one thread per closing rank.
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

// symbol preceding the suffix at rank r
template <class SAT>
__device__ __forceinline__ u8 bwt_at(const u8* __restrict__ bwt, const u8* __restrict__ text, const SAT* __restrict__ sa, u64 r) {
    if (bwt) return bwt[r];
    u64 p = (u64)sa[r];
    return p ? text[p - 1] : (u8)0;
}

template <class SAT, int W, int P>
__global__ void __launch_bounds__(TPB) detect_walk_kernel(const u32* __restrict__ lcp, const SAT* __restrict__ sa, const u8* __restrict__ bwt,
                                                          const u8* __restrict__ text, u64 m, u64 j_first, DetectParams prm, const u64* __restrict__ doc_start_g, Report* __restrict__ reports, u32 cap,
                                                     u32* __restrict__ count, u16* __restrict__ thresh) {
    MMB_DYN_SMEM(u64, doc_start);
    for (u32 i = threadIdx.x; i <= prm.ndocs; i += blockDim.x) doc_start[i] = doc_start_g[i];
    __syncthreads();
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (j >= m || j < j_first) return;
    u32 cur = lcp[j], v = lcp[j - 1];
    if (v <= cur || v < prm.min_len) return;

    DocCounter<W, P> docs; docs.clear();
    u64 k = j - 1;                  // lowest rank inside the current interval
    u64 anchor_sa = ~0ull;          // SA of the lowest-ranked document-0 suffix seen
    u32 size = 1;
    bool left_differs = false;
    {
        u64 p = (u64)sa[k]; u32 d = find_doc(doc_start, prm.ndocs, p);
        if (d == 0) anchor_sa = p;
        if (!docs.add(d, prm.max_doc_freq)) return;
    }
    for (;;) {
        while (lcp[k] >= v) {       // rank k-1 shares v symbols with rank k: absorb it (k >= 2 here because lcp[0] = lcp[1] = 0 < v)
            left_differs |= (bwt_at(bwt, text, sa, k) != bwt_at(bwt, text, sa, k - 1));
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
                if (slot < cap) { Report r; r.key = (j << 20) | min(size, 0xfffffu); r.s = k; r.len = v; r.pad = 0; reports[slot] = r; }
            }
        }
        v = before;
        if (v <= cur || v < prm.min_len) return;
    }
}

// ------------------------------------------------------------------------------------------------ windowed-minimum detector
constexpr u32 DW_TPB = 256, DW_JPT = 4, DW_TILE = DW_TPB * DW_JPT;
constexpr u32 DW_MAX_SIZES = 64;

struct WindowPlan {
    u32 size_lo, size_hi;       // interval sizes that can qualify
    u32 lev_lo, lev_hi;         // floor(log2(size-1)) over that range: window-minimum levels kept in shared memory
    u32 nkept;                  // number of levels >= 1 kept (lev_hi - max(lev_lo, 1) + 1, or 0)
    u32 halo;                   // ranks staged to the left of the tile (multiple of 4, >= size_hi)
    u32 span;                   // halo + DW_TILE
};

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

// One TMA bulk copy global -> shared, completion on an mbarrier (cp.async.bulk; UBLKCP in SASS).  bytes and both addresses are
// multiples of 16.  Thread 0 issues, everybody waits.
__device__ __forceinline__ void tma_load_tile(void* dst, const void* src, u32 bytes, u64* bar) {
#ifdef MMB_SIMT
    if (threadIdx.x == 0) memcpy(dst, src, bytes);
    (void)bar;
    __syncthreads();
#else
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                     "r"(smem_u32(bar))
                     : "memory");
    }
    u32 done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar))
            : "memory");
    }
#endif
}


// Two bulk copies on two mbarriers (second optional: bytes2 = 0); everybody waits for both.  Leaves the block synchronised.
__device__ __forceinline__ void tma_load_two(void* dst1, const void* src1, u32 bytes1, void* dst2, const void* src2, u32 bytes2, u64* bar) {
#ifdef MMB_SIMT
    if (threadIdx.x == 0) { if (bytes1) memcpy(dst1, src1, bytes1); if (bytes2) memcpy(dst2, src2, bytes2); }
    (void)bar;
    __syncthreads();
#else
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    u32 total = bytes1 + bytes2;
    if (total == 0) return;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(total) : "memory");
        if (bytes1)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst1)), "l"(src1),
                         "r"(bytes1), "r"(smem_u32(bar))
                         : "memory");
        if (bytes2)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst2)), "l"(src2),
                         "r"(bytes2), "r"(smem_u32(bar))
                         : "memory");
    }
    u32 done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar))
            : "memory");
    }
#endif
}

template <class SAT, int W, int P>
__global__ void __launch_bounds__(DW_TPB) detect_window_kernel(const u32* __restrict__ lcp, const SAT* __restrict__ sa, const u8* __restrict__ bwt,
                                                              const u8* __restrict__ text, u64 m, u64 j_first, DetectParams prm, WindowPlan plan,
                                                              const u64* __restrict__ doc_start_g, Report* __restrict__ reports, u32 cap,
                                                              u32* __restrict__ count, u16* __restrict__ thresh) {
    MMB_DYN_SMEM(u64, smem64);
    u64* bar = smem64;                                  // 16 bytes reserved
    u64* doc_start = smem64 + 2;                        // ndocs + 1 (+1 pad to keep 16-byte alignment)
    u32* lc = reinterpret_cast<u32*>(doc_start + ((prm.ndocs + 2) & ~1u));      // level 0: the lcp tile itself
    u32* lev = lc + plan.span;                          // levels lev_lo' .. lev_hi, each plan.span entries (level 0 lives in lc)
    const u32 span = plan.span, halo = plan.halo;
    const u64 tile0 = (u64)blockIdx.x * DW_TILE;        // first closing rank handled by this block
    const i64 origin = (i64)tile0 - (i64)halo;          // rank held by lc[0]

    for (u32 i = threadIdx.x; i <= prm.ndocs; i += blockDim.x) doc_start[i] = doc_start_g[i];
    // stage lcp[origin .. origin + span): ranks below 0 or at/after m read as 0 (an lcp of 0 closes everything)
    {
        u64 lo = origin < 0 ? 0 : (u64)origin, hi = min((u64)(origin + span), m);
        u32 skip = (u32)((i64)lo - origin);
        for (u32 i = threadIdx.x; i < skip; i += blockDim.x) lc[i] = 0;
        u32 have = hi > lo ? (u32)(hi - lo) : 0u, padded = (have + 3) & ~3u;      // lcp is allocated with slack past m (engine)
        if (padded) tma_load_tile(lc + skip, lcp + lo, padded * 4, bar); else __syncthreads();
        for (u32 i = skip + have + threadIdx.x; i < span; i += blockDim.x) lc[i] = 0;
    }
    __syncthreads();
    // window minima: level t holds min(lc[i .. i + 2^t - 1]); built by doubling, levels lev_lo..lev_hi are kept
    {
        constexpr u32 PER = 6;                          // span <= DW_TPB * PER is checked on the host
        u32 v[PER];
#pragma unroll
        for (u32 q = 0; q < PER; ++q) { u32 i = threadIdx.x + q * DW_TPB; v[q] = i < span ? lc[i] : 0u; }
        u32* work = lev + (size_t)plan.nkept * span;      // scratch level after the kept ones
        for (u32 t = 1; t <= plan.lev_hi; ++t) {
            const u32* prev = (t == 1) ? lc : work;
            u32 stride = 1u << (t - 1);
            u32 nv[PER];
#pragma unroll
            for (u32 q = 0; q < PER; ++q) {
                u32 i = threadIdx.x + q * DW_TPB;
                nv[q] = v[q];
                if (i < span && i + stride < span) nv[q] = min(v[q], prev[i + stride]);
            }
            __syncthreads();
#pragma unroll
            for (u32 q = 0; q < PER; ++q) {
                u32 i = threadIdx.x + q * DW_TPB;
                v[q] = nv[q];
                if (i < span) { work[i] = nv[q]; if (t >= plan.lev_lo) lev[(size_t)(t - max(plan.lev_lo, 1u)) * span + i] = nv[q]; }
            }
            __syncthreads();
        }
    }
    const u32 lane = threadIdx.x & 31;
    const u32 lev_base = max(plan.lev_lo, 1u);
    auto window_min = [&](u32 idx_j, u32 size) -> u32 {      // min(lc[idx_j - size + 1 .. idx_j - 1]), size >= 2
        u32 w = size - 1;
        u32 t = 31 - __clz((int)w), pw = 1u << t;
        const u32* L = t ? lev + (size_t)(t - lev_base) * span : lc;
        return min(L[idx_j - size + 1], L[idx_j - pw]);
    };

    for (u32 q = 0; q < DW_JPT; ++q) {
        u32 idx = halo + q * DW_TPB + threadIdx.x;          // position of rank j in lc
        u64 j = tile0 + q * DW_TPB + threadIdx.x;
        bool active = j >= j_first && j >= 1 && j < m;
        u32 cur = lc[idx];
        u32 size = plan.size_lo;
        if (active) {                                       // the smallest window has the largest minimum: cheap rejection
            u32 v0 = (j >= size) ? window_min(idx, size) : 0u;
            if (v0 <= cur || v0 < prm.min_len) active = false;
        }
        for (;;) {                                          // warp-uniform: one candidate per lane per turn
            bool have = false; u32 v = 0;
            while (active && size <= plan.size_hi) {
                if (j < size) { active = false; break; }
                v = window_min(idx, size);
                if (v <= cur || v < prm.min_len) { active = false; break; }
                if (lc[idx - size] < v) { have = true; break; }
                ++size;
            }
            if (!have) active = false;
            unsigned cand = __ballot_sync(0xffffffffu, have);
            if (!cand) break;
            if (P == 1 && prm.max_doc_freq == 1) {
                // MUM modes: the warp verifies each candidate together
                while (cand) {
                    int src = __ffs((int)cand) - 1; cand &= cand - 1;
                    u64 cj = __shfl_sync(0xffffffffu, j, src);
                    u32 cs = __shfl_sync(0xffffffffu, size, src), cv = __shfl_sync(0xffffffffu, v, src), ccur = __shfl_sync(0xffffffffu, cur, src);
                    u32 cbefore = __shfl_sync(0xffffffffu, lc[idx - (have ? size : 0)], src);
                    u64 s0 = cj - cs;
                    if (bwt && !prm.merge) {
                        // cheap rejection first: an interval whose suffixes are all preceded by the same symbol is not left-maximal
                        // (the common case inside long shared regions); only the BWT bytes of the tile are read for this
                        bool d0 = false;
                        for (u32 t = 1 + lane; t < cs; t += 32) d0 |= bwt[s0 + t] != bwt[s0 + t - 1];
                        if (!__any_sync(0xffffffffu, d0)) continue;
                    }
                    u64 bits[W];
#pragma unroll
                    for (int w = 0; w < W; ++w) bits[w] = 0;
                    u64 anchor = ~0ull; bool differs = false; u32 carry = 0;
                    for (u32 t0 = 0; t0 < cs; t0 += 32) {
                        u32 t = t0 + lane;
                        u32 sym = 0;
                        if (t < cs) {
                            u64 r = s0 + t; u64 p = (u64)sa[r];
                            u32 d = find_doc(doc_start, prm.ndocs, p);
                            bits[(W == 1) ? 0 : (d >> 6)] |= 1ull << (d & 63);
                            if (d == 0) anchor = p;
                            sym = bwt ? bwt[r] : (p ? text[p - 1] : (u8)0);
                        }
                        u32 left = __shfl_up_sync(0xffffffffu, sym, 1);
                        if (lane == 0) left = carry;
                        if (t < cs && t > 0 && left != sym) differs = true;
                        carry = __shfl_sync(0xffffffffu, sym, 31);
                    }
                    u32 distinct = 0;
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        u64 b = bits[w];
#pragma unroll
                        for (u32 d = 16; d > 0; d >>= 1) b |= __shfl_xor_sync(0xffffffffu, b, d);
                        distinct += (u32)__popcll(b);
                    }
                    differs = __any_sync(0xffffffffu, differs);
                    if (distinct == cs && distinct >= prm.num_distinct) {          // every document at most once
                        if (prm.merge && anchor != ~0ull) thresh[anchor] = (u16)min(max(cbefore, ccur), 65535u);
                        if (differs && lane == 0) {
                            u32 slot = atomicAdd(count, 1u);
                            if (slot < cap) { Report r; r.key = (cj << 20) | cs; r.s = s0; r.len = cv; r.pad = 0; reports[slot] = r; }
                        }
                    }
                }
            } else if (have) {
                // MEM modes: the owning thread counts documents itself
                DocCounter<W, P> docs; docs.clear();
                u64 s0 = j - size; bool ok = true, differs = false;
                u8 prev_sym = 0;
                for (u32 t = 0; t < size && ok; ++t) {
                    u64 r = s0 + t; u64 p = (u64)sa[r];
                    ok = docs.add(find_doc(doc_start, prm.ndocs, p), prm.max_doc_freq);
                    u8 sym = bwt ? bwt[r] : (p ? text[p - 1] : (u8)0);
                    if (t > 0 && sym != prev_sym) differs = true;
                    prev_sym = sym;
                }
                if (ok && docs.distinct >= prm.num_distinct && differs) {
                    u32 slot = atomicAdd(count, 1u);
                    if (slot < cap) { Report r; r.key = (j << 20) | size; r.s = s0; r.len = v; r.pad = 0; reports[slot] = r; }
                }
            }
            if (have) ++size;
        }
    }
}

// Warp-cooperative check of one MUM-mode candidate (all lanes hold the same arguments): the interval [s0, s0 + cs) qualifies iff its
// suffixes come from cs distinct documents and cs >= num_distinct; it is reported iff it is also left-maximal (leftmax: 1 / 0 when the
// caller already knows, -1 = derive it from the symbols that precede the suffixes).
template <class SAT, int W>
__device__ __forceinline__ void verify_mum_candidate(u64 cj, u32 cs, u32 cv, u32 ccur, u32 cbefore, int leftmax, const SAT* __restrict__ sa,
                                                     const u8* __restrict__ bwt, const u8* __restrict__ text, const u64* doc_start,
                                                     const DetectParams& prm, Report* __restrict__ reports, u32 cap, u32* __restrict__ count,
                                                     u16* __restrict__ thresh, u32 lane) {
    const u64 s0 = cj - cs;
    u64 bits[W];
#pragma unroll
    for (int w = 0; w < W; ++w) bits[w] = 0;
    u64 anchor = ~0ull; bool differs = false; u32 carry = 0;
    for (u32 t0 = 0; t0 < cs; t0 += 32) {
        u32 t = t0 + lane;
        u32 sym = 0;
        if (t < cs) {
            u64 rk = s0 + t; u64 p = (u64)sa[rk];
            u32 dd;
            if (prm.doc_coarse) { dd = prm.doc_coarse[p >> prm.coarse_shift]; while (doc_start[dd + 1] <= p) ++dd; }
            else dd = find_doc(doc_start, prm.ndocs, p);
            bits[(W == 1) ? 0 : (dd >> 6)] |= 1ull << (dd & 63);
            if (dd == 0) anchor = p;
            if (leftmax < 0) sym = bwt ? bwt[rk] : (p ? text[p - 1] : (u8)0);
        }
        if (leftmax < 0) {
            u32 left = __shfl_up_sync(0xffffffffu, sym, 1);
            if (lane == 0) left = carry;
            if (t < cs && t > 0 && left != sym) differs = true;
            carry = __shfl_sync(0xffffffffu, sym, 31);
        }
    }
    u32 distinct = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
#ifdef MMB_SIMT
        u64 bb = bits[w];
        for (u32 dd = 16; dd > 0; dd >>= 1) bb |= __shfl_xor_sync(0xffffffffu, bb, dd);
        distinct += (u32)__popcll(bb);
#else
        distinct += (u32)__popc(__reduce_or_sync(0xffffffffu, (u32)bits[w])) + (u32)__popc(__reduce_or_sync(0xffffffffu, (u32)(bits[w] >> 32)));
#endif
    }
    differs = leftmax < 0 ? __any_sync(0xffffffffu, differs) : leftmax != 0;
    if (distinct == cs && distinct >= prm.num_distinct) {          // every document at most once
        if (prm.merge && anchor != ~0ull) thresh[anchor] = (u16)min(max(cbefore, ccur), 65535u);
        if (differs && lane == 0) {
            u32 slot = atomicAdd(count, 1u);
            if (slot < cap) { Report rp; rp.key = (cj << 20) | cs; rp.s = s0; rp.len = cv; rp.pad = 0; reports[slot] = rp; }
        }
    }
}

// The same check from a document-id byte per rank (StreamTile::doc, experimental): count bytes instead of count gathered SA values and
// lookups; SA is read once, by the lane that holds document 0, for the -M threshold.  leftmax must be known (0 / 1).
template <class SAT, int W>
__device__ __forceinline__ void verify_mum_candidate_docs(u64 cj, u32 cs, u32 cv, u32 ccur, u32 cbefore, int leftmax, const SAT* __restrict__ sa,
                                                          const u8* __restrict__ docs, const DetectParams& prm, Report* __restrict__ reports,
                                                          u32 cap, u32* __restrict__ count, u16* __restrict__ thresh, u32 lane) {
    const u64 s0 = cj - cs;
    u64 bits[W];
#pragma unroll
    for (int w = 0; w < W; ++w) bits[w] = 0;
    u64 anchor_rank = ~0ull;
    bool bad = false;                                   // a rank without a document (the sentinel) cannot be part of a match
    for (u32 t = lane; t < cs; t += 32) {
        u32 dd = docs[s0 + t];
        if (dd >= prm.ndocs) { bad = true; continue; }
        bits[(W == 1) ? 0 : (dd >> 6)] |= 1ull << (dd & 63);
        if (dd == 0) anchor_rank = s0 + t;
    }
    u32 distinct = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
#ifdef MMB_SIMT
        u64 bb = bits[w];
        for (u32 dd = 16; dd > 0; dd >>= 1) bb |= __shfl_xor_sync(0xffffffffu, bb, dd);
        distinct += (u32)__popcll(bb);
#else
        distinct += (u32)__popc(__reduce_or_sync(0xffffffffu, (u32)bits[w])) + (u32)__popc(__reduce_or_sync(0xffffffffu, (u32)(bits[w] >> 32)));
#endif
    }
    bad = __any_sync(0xffffffffu, bad);
    if (!bad && distinct == cs && distinct >= prm.num_distinct) {          // every document at most once
        if (prm.merge && anchor_rank != ~0ull) thresh[(u64)sa[anchor_rank]] = (u16)min(max(cbefore, ccur), 65535u);
        if (leftmax && lane == 0) {
            u32 slot = atomicAdd(count, 1u);
            if (slot < cap) { Report rp; rp.key = (cj << 20) | cs; rp.s = s0; rp.len = cv; rp.pad = 0; reports[slot] = rp; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ chunked window-minimum detector
// Same test as detect_window_kernel, but the minimum of the smallest window (size_lo - 1 entries ending at j-1) costs O(1) per rank
// instead of one shared-memory pass per power of two: the staged tile is cut into chunks of C = (size_lo - 1) | 1 ranks (C >= the
// window, C odd), one thread walks one chunk.  Pass 1 writes the suffix minima of every chunk; in pass 2 the window of rank j is the
// tail of the previous chunk (one suffix minimum) joined with the head of its own chunk (a running minimum in a register).  Lanes
// are C words apart and C is odd, so every shared-memory access is conflict-free.  Larger sizes extend the window one entry at a
// time.
// FAST (MUM modes over a stream tile that carries BWT symbols): the BWT bytes of the tile are staged too, pass 1 notes the
// last rank of every chunk whose symbol differs from its predecessor's, and pass 2 carries "last change at or before j-1" in a
// register -- the reference's own left-maximality test (last_bwt_change > start, include/mem_finder.hpp:165-166,189-192) -- so that
// intervals that are not left-maximal (the common case inside shared regions) are dropped by the lane that found them and only the
// survivors reach the warp-cooperative document check.  With -M every interval of the right size reaches that check (its threshold is
// recorded whether or not it is left-maximal), still from the once-per-eight-ranks convergence point of the single-size loop.
constexpr u32 DC_TPB = 128;
struct ChunkPlan {
    u32 size_lo, size_hi;
    u32 w0;                     // size_lo - 1: entries in the smallest window
    u32 chunk;                  // C
    u32 halo_chunks;            // H: chunks staged in front of the first closing rank (H * C >= size_hi)
    u32 rounds;                 // R: chunks per thread; a block stages R * DC_TPB chunks and closes (R * DC_TPB - H) * C ranks
    u32 span;                   // staged lcp entries: R * DC_TPB * C + 4, rounded up to a multiple of 4
    u32 span_b;                 // staged bwt bytes (FAST): R * DC_TPB * C + 32, rounded up to a multiple of 16
};

template <class SAT, int W, int P, bool FAST>
__global__ void __launch_bounds__(DC_TPB) detect_chunk_kernel(const u32* __restrict__ lcp, const SAT* __restrict__ sa, const u8* __restrict__ bwt,
                                                             const u8* __restrict__ text, u64 m, u64 j_first, DetectParams prm, ChunkPlan plan,
                                                             const u64* __restrict__ doc_start_g, Report* __restrict__ reports, u32 cap,
                                                             u32* __restrict__ count, u16* __restrict__ thresh, const u8* __restrict__ docs) {
    MMB_DYN_SMEM(u64, smem64);
    u64* bar = smem64;                                  // 2 mbarriers (16 bytes)
    u64* doc_start = smem64 + 2;                        // ndocs + 1 (+1 pad to keep 16-byte alignment)
    u32* lc = reinterpret_cast<u32*>(doc_start + ((prm.ndocs + 2) & ~1u));      // the lcp tile
    u32* suf = lc + plan.span;                          // suffix minima inside every chunk
    u8* bt = reinterpret_cast<u8*>(suf + plan.span);    // FAST: the bwt tile
    int* lastin = reinterpret_cast<int*>(bt + plan.span_b);                     // FAST: per chunk, tile index of its last symbol change (-1: none)
    const u32 C = plan.chunk, H = plan.halo_chunks, G = plan.rounds * DC_TPB, W0 = plan.w0, span = plan.span;
    const i64 origin = (i64)blockIdx.x * (i64)(G - H) * C - (i64)H * C;         // rank of tile index 0 (first entry of chunk 0)
    const i64 origin_al = origin - (((origin % 4) + 4) % 4);                    // staging starts on a 16-byte boundary of lcp
    const u32 d = (u32)(origin - origin_al);                                    // tile index x lives at lc[d + x]
    const i64 origin_b = (origin - 1) - ((((origin - 1) % 16) + 16) % 16);      // same for the bytes, one rank earlier
    const u32 db = (u32)(origin - origin_b);                                    // tile index x lives at bt[db + x], db >= 1

    for (u32 i = threadIdx.x; i <= prm.ndocs; i += blockDim.x) doc_start[i] = doc_start_g[i];
    {   // stage lcp[origin_al .. origin_al + span): ranks below 0 or at/after m read as 0 (an lcp of 0 closes everything)
        u64 lo = origin_al < 0 ? 0 : (u64)origin_al, hi = min((u64)(origin_al + (i64)span), m);
        u32 skip = (u32)((i64)lo - origin_al);
        u32 have = hi > lo ? (u32)(hi - lo) : 0u, padded = min((have + 3) & ~3u, span - skip);      // lcp is allocated with slack past m
        u32 skip_b = 0, have_b = 0, padded_b = 0; u64 lo_b = 0;
        if (FAST) {
            lo_b = origin_b < 0 ? 0 : (u64)origin_b;
            u64 hi_b = min((u64)(origin_b + (i64)plan.span_b), m);
            skip_b = (u32)((i64)lo_b - origin_b);
            have_b = hi_b > lo_b ? (u32)(hi_b - lo_b) : 0u; padded_b = min((have_b + 15) & ~15u, plan.span_b - skip_b);   // bwt has slack too
        }
        tma_load_two(lc + skip, lcp + lo, padded * 4, bt + skip_b, FAST ? bwt + lo_b : nullptr, padded_b, bar);
        for (u32 i = threadIdx.x; i < skip; i += blockDim.x) lc[i] = 0;
        for (u32 i = skip + have + threadIdx.x; i < span; i += blockDim.x) lc[i] = 0;
        if (FAST) {
            for (u32 i = threadIdx.x; i < skip_b; i += blockDim.x) bt[i] = 0;
            for (u32 i = skip_b + have_b + threadIdx.x; i < plan.span_b; i += blockDim.x) bt[i] = 0;
        }
    }
    __syncthreads();
    for (u32 g = threadIdx.x; g < G; g += DC_TPB) {     // pass 1
        u32 base = d + g * C, mn = 0xffffffffu;
        if (FAST) {
            const u8* b = bt + db + g * C;
            int last = -1; u32 right = b[C - 1];
#pragma unroll 4
            for (u32 k = C; k-- > 0;) {
                mn = min(mn, lc[base + k]); suf[base + k] = mn;
                u32 left = b[(int)k - 1];
                if (left != right && last < 0) last = (int)(g * C + k);
                right = left;
            }
            lastin[g] = last;
        } else {
#pragma unroll 4
            for (u32 k = C; k-- > 0;) { mn = min(mn, lc[base + k]); suf[base + k] = mn; }
        }
    }
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    for (u32 r = 0; r < plan.rounds; ++r) {             // pass 2 (warp-uniform loops: the verification below is warp-cooperative)
        const u32 g = r * DC_TPB + threadIdx.x;
        const u32 base = d + g * C;
        const i64 j0 = origin + (i64)g * C;             // rank of the chunk's first entry
        // closing ranks of this chunk that are examined: k in [k_lo, k_hi)
        u32 k_lo = C, k_hi = 0;
        if (g >= H) {
            i64 first = max((i64)max(j_first, (u64)1), j0), end = min((i64)m, j0 + (i64)C);
            if (first < end) { k_lo = (u32)(first - j0); k_hi = (u32)(end - j0); }
        }
        u32 run = 0xffffffffu;                          // min(lc[base .. base + k - 1])
        int last = -0x40000000;                         // FAST: tile index of the last symbol change at or before the previous rank
        u32 prev_sym = 0;
        const u8* b = bt + db + g * C;
        if (FAST && g >= H) {
            for (u32 back = 1; back <= H; ++back) { int l = lastin[g - back]; if (l >= 0) { last = l; break; } }
            prev_sym = b[-1];
        }
        if (FAST && plan.size_lo == plan.size_hi) {
            // One interval size (strict multi-MUMs): the lane decides alone whether rank j closes a left-maximal interval of that size
            // and only notes it in a bit mask; the warp converges once per eight ranks for the cooperative document check of the survivors
            // (rare, except with -M where every interval of that size needs its threshold recorded).
            const u32 S = plan.size_lo;
            const u32* plc = lc + base;
            const u32* psuf = suf + base - (g >= H ? W0 : 0u);      // chunks in the halo are walked for nothing: keep their reads in bounds
            const u32* pbef = lc + base - (g >= H ? S : 0u);
            const int x0 = (int)(g * C);
            for (u32 k0 = 0; k0 < C; k0 += 8) {
                u32 cmask = 0, lmask = 0;               // candidates among the eight ranks, and which of them are left-maximal
                u32 v8[8];                              // window minima of the eight ranks (statically indexed: registers)
#pragma unroll
                for (u32 u = 0; u < 8; ++u) {
                    const u32 k = k0 + u;
                    v8[u] = 0;
                    if (k < C) {
                        const u32 cur = plc[k];
                        const u32 v = (k >= W0) ? run : min(psuf[k], run);
                        v8[u] = v;
                        run = min(run, cur);
                        const int last_before = last;
                        const u32 sym = b[k];
                        if (sym != prev_sym) last = x0 + (int)k;
                        prev_sym = sym;
                        bool c = k >= k_lo && k < k_hi && v > cur && v >= prm.min_len;
                        // -M needs the threshold of every interval that passes the size / document test, left-maximal or not
                        const bool lm = last_before > x0 + (int)k - (int)S;
                        if (c) c = pbef[k] < v && (prm.merge || lm);
                        cmask |= (u32)c << u; lmask |= (u32)(c && lm) << u;
                    }
                }
                if (!__any_sync(0xffffffffu, cmask != 0)) continue;
                for (;;) {
                    const bool have = cmask != 0;
                    unsigned cand = __ballot_sync(0xffffffffu, have);
                    if (!cand) break;
                    u32 kc = 0, vv = 0; int lmc = 0;
                    if (have) {
                        const u32 bit = (u32)__ffs((int)cmask) - 1;
                        kc = k0 + bit; cmask &= cmask - 1; lmc = (int)((lmask >> bit) & 1);
#pragma unroll
                        for (u32 u = 0; u < 8; ++u) if (bit == u) vv = v8[u];
                    }
                    const u64 jc = (u64)(j0 + (i64)kc);
                    const u32 curc = lc[base + kc], befc = have ? pbef[kc] : 0u;
                    while (cand) {
                        int src = __ffs((int)cand) - 1; cand &= cand - 1;
                        u64 cj = __shfl_sync(0xffffffffu, jc, src);
                        u32 cv = __shfl_sync(0xffffffffu, vv, src), ccur = __shfl_sync(0xffffffffu, curc, src), cbef = __shfl_sync(0xffffffffu, befc, src);
                        int clm = __shfl_sync(0xffffffffu, lmc, src);
                        if (docs) verify_mum_candidate_docs<SAT, W>(cj, S, cv, ccur, cbef, clm, sa, docs, prm, reports, cap, count, thresh, lane);
                        else verify_mum_candidate<SAT, W>(cj, S, cv, ccur, cbef, clm, sa, bwt, text, doc_start, prm, reports, cap, count, thresh, lane);
                    }
                }
            }
            continue;
        }
        for (u32 k = 0; k < C; ++k) {
            const u32 idx = base + k;
            const u32 cur = lc[idx];
            u32 v = 0;
            if (g >= H) v = (k >= W0) ? run : min(suf[idx - W0], run);      // k = W0 only when C = W0 + 1: the window is the chunk's head
            run = min(run, cur);
            const int x = (int)(g * C + k);
            const int last_before = last;
            if (FAST && g >= H) { u32 sym = b[k]; if (sym != prev_sym) last = x; prev_sym = sym; }
            bool active = k >= k_lo && k < k_hi && v > cur && v >= prm.min_len;
            if (!__any_sync(0xffffffffu, active)) continue;
            const u64 j = (u64)(j0 + k);
            u32 size = plan.size_lo;
            for (;;) {                                  // one candidate per lane per turn
                bool have = false;
                while (active && size <= plan.size_hi) {
                    if (j < size || v <= cur || v < prm.min_len) { active = false; break; }
                    u32 before = lc[idx - size];
                    if (before < v && (!FAST || prm.merge || last_before > x - (int)size)) { have = true; break; }
                    ++size; v = min(v, before);
                }
                if (!have) active = false;
                unsigned cand = __ballot_sync(0xffffffffu, have);
                if (!cand) break;
                if (P == 1 && prm.max_doc_freq == 1) {
                    while (cand) {
                        int src = __ffs((int)cand) - 1; cand &= cand - 1;
                        u64 cj = __shfl_sync(0xffffffffu, j, src);
                        u32 cs = __shfl_sync(0xffffffffu, size, src), cv = __shfl_sync(0xffffffffu, v, src), ccur = __shfl_sync(0xffffffffu, cur, src);
                        u32 cbefore = __shfl_sync(0xffffffffu, lc[idx - (have ? size : 0)], src);
                        u64 s0 = cj - cs;
                        if (!FAST && bwt && !prm.merge) {
                            bool d0 = false;
                            for (u32 t = 1 + lane; t < cs; t += 32) d0 |= bwt[s0 + t] != bwt[s0 + t - 1];
                            if (!__any_sync(0xffffffffu, d0)) continue;
                        }
                        u64 bits[W];
#pragma unroll
                        for (int w = 0; w < W; ++w) bits[w] = 0;
                        u64 anchor = ~0ull; bool differs = false; u32 carry = 0;
                        for (u32 t0 = 0; t0 < cs; t0 += 32) {
                            u32 t = t0 + lane;
                            u32 sym = 0;
                            if (t < cs) {
                                u64 rk = s0 + t; u64 p = (u64)sa[rk];
                                u32 dd = find_doc(doc_start, prm.ndocs, p);
                                bits[(W == 1) ? 0 : (dd >> 6)] |= 1ull << (dd & 63);
                                if (dd == 0) anchor = p;
                                sym = bwt ? bwt[rk] : (p ? text[p - 1] : (u8)0);
                            }
                            u32 left = __shfl_up_sync(0xffffffffu, sym, 1);
                            if (lane == 0) left = carry;
                            if (t < cs && t > 0 && left != sym) differs = true;
                            carry = __shfl_sync(0xffffffffu, sym, 31);
                        }
                        u32 distinct = 0;
#pragma unroll
                        for (int w = 0; w < W; ++w) {
                            u64 bb = bits[w];
#pragma unroll
                            for (u32 dd = 16; dd > 0; dd >>= 1) bb |= __shfl_xor_sync(0xffffffffu, bb, dd);
                            distinct += (u32)__popcll(bb);
                        }
                        differs = __any_sync(0xffffffffu, differs);
                        if (distinct == cs && distinct >= prm.num_distinct) {          // every document at most once
                            if (prm.merge && anchor != ~0ull) thresh[anchor] = (u16)min(max(cbefore, ccur), 65535u);
                            if (differs && lane == 0) {
                                u32 slot = atomicAdd(count, 1u);
                                if (slot < cap) { Report rp; rp.key = (cj << 20) | cs; rp.s = s0; rp.len = cv; rp.pad = 0; reports[slot] = rp; }
                            }
                        }
                    }
                } else if (have) {
                    DocCounter<W, P> docs; docs.clear();
                    u64 s0 = j - size; bool ok = true, differs = false;
                    u8 psym = 0;
                    if (bwt) {      // left-maximality first, from the staged symbols alone (no SA gather, no document lookup): most intervals fail it
                        bool any = false;
                        for (u32 t = 1; t < size && !any; ++t) any = bwt[s0 + t] != bwt[s0 + t - 1];
                        ok = any;
                    }
                    for (u32 t = 0; t < size && ok; ++t) {
                        u64 rk = s0 + t; u64 p = (u64)sa[rk];
                        ok = docs.add(find_doc(doc_start, prm.ndocs, p), prm.max_doc_freq);
                        u8 sym = bwt ? bwt[rk] : (p ? text[p - 1] : (u8)0);
                        if (t > 0 && sym != psym) differs = true;
                        psym = sym;
                    }
                    if (ok && docs.distinct >= prm.num_distinct && differs) {
                        u32 slot = atomicAdd(count, 1u);
                        if (slot < cap) { Report rp; rp.key = (j << 20) | size; rp.s = s0; rp.len = v; rp.pad = 0; reports[slot] = rp; }
                    }
                }
                if (have) { v = min(v, lc[idx - size]); ++size; }
            }
        }
    }
}

__global__ void report_keys(const Report* __restrict__ rep, u32 n, u64* __restrict__ keys, u32* __restrict__ idx) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = rep[i].key; idx[i] = i; }
}
__global__ void report_size_keys(const Report* __restrict__ rep, u32 n, u64* __restrict__ keys, u32* __restrict__ idx) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = (rep[i].key >> 20) - rep[i].s; idx[i] = i; }
}
__global__ void report_rank_keys(const Report* __restrict__ rep, const u32* __restrict__ idx, u32 n, u64* __restrict__ keys) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = rep[idx[i]].key >> 20;
}
__global__ void report_gather(const Report* __restrict__ rep, const u32* __restrict__ idx, u32 n, Report* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rep[idx[i]];
}

struct DetectArgs {
    const u32* lcp; const u8* bwt; const u8* text; u64 m; u64 j_first; DetectParams prm; const u64* d_doc_start; Report* rep; u32 cap; u32* count; u16* thresh;
    bool use_window; WindowPlan plan;
    bool use_chunk = false; ChunkPlan cplan;
    const u8* docs = nullptr;
};

template <class SAT, int W, int P>
void launch_detect(const SAT* sa, const DetectArgs& a, cudaStream_t s) {
    if (a.use_chunk) {
        bool fast = a.bwt && P == 1 && a.prm.max_doc_freq == 1 && (!a.prm.merge || a.cplan.size_lo == a.cplan.size_hi);
        auto detect_chunk_k = fast ? detect_chunk_kernel<SAT, W, P, true> : detect_chunk_kernel<SAT, W, P, false>;
        size_t smem = 16 + (size_t)((a.prm.ndocs + 2) & ~1u) * 8 + (size_t)a.cplan.span * 4 * 2 + (fast ? a.cplan.span_b + (size_t)a.cplan.rounds * DC_TPB * 4 : 0);
#ifndef MMB_SIMT
        if (smem > 48 * 1024) MMB_CUDA(cudaFuncSetAttribute(detect_chunk_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
        u64 per_block = (u64)(a.cplan.rounds * DC_TPB - a.cplan.halo_chunks) * a.cplan.chunk;
        MMB_LAUNCH(detect_chunk_k, ceil_div(a.m, per_block), DC_TPB, smem, s, a.lcp, sa, a.bwt, a.text, a.m, a.j_first, a.prm, a.cplan, a.d_doc_start, a.rep, a.cap, a.count, a.thresh, fast ? a.docs : (const u8*)nullptr);
    } else if (a.use_window) {
        auto detect_window_k = detect_window_kernel<SAT, W, P>;
        size_t smem = 16 + (size_t)((a.prm.ndocs + 2) & ~1u) * 8 + (size_t)a.plan.span * 4 * (2 + a.plan.nkept);
#ifndef MMB_SIMT
        if (smem > 48 * 1024) MMB_CUDA(cudaFuncSetAttribute(detect_window_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
        MMB_LAUNCH(detect_window_k, ceil_div(a.m, DW_TILE), DW_TPB, smem, s, a.lcp, sa, a.bwt, a.text, a.m, a.j_first, a.prm, a.plan, a.d_doc_start, a.rep, a.cap, a.count, a.thresh);
    } else {
        auto detect_walk_k = detect_walk_kernel<SAT, W, P>;
        size_t smem = (size_t)(a.prm.ndocs + 1) * sizeof(u64);
        MMB_LAUNCH(detect_walk_k, ceil_div(a.m - 1, TPB), TPB, smem, s, a.lcp, sa, a.bwt, a.text, a.m, a.j_first, a.prm, a.d_doc_start, a.rep, a.cap, a.count, a.thresh);
    }
}

template <class SAT, int W>
void dispatch_planes(const SAT* sa, const DetectArgs& a, cudaStream_t s) {
    u32 f = a.prm.max_doc_freq;
    if (f <= 1) launch_detect<SAT, W, 1>(sa, a, s);
    else if (f <= 3) launch_detect<SAT, W, 2>(sa, a, s);
    else if (f <= 255) launch_detect<SAT, W, 8>(sa, a, s);
    else throw Error(30, "per-sequence frequency limits above 255 are not supported");
}

u32 floor_log2(u32 x) { u32 t = 0; while ((2u << t) <= x) ++t; return t; }

// Which interval sizes can qualify, and whether the windowed kernel covers them.  Returns false if nothing can qualify at all.
// The chunked kernel: any size range whose smallest window fits a chunk of <= 127 ranks (two staged arrays of DC_TPB chunks).
bool make_chunk_plan(u64 lo, u64 hi, ChunkPlan& cp) {
    const char* force = getenv("MMB_DETECT");
    if (force && strcmp(force, "chunk") != 0) return false;
    if (hi - lo + 1 > DW_MAX_SIZES || lo - 1 > 127) return false;
    cp.size_lo = (u32)lo; cp.size_hi = (u32)hi; cp.w0 = (u32)lo - 1;
    cp.chunk = cp.w0 | 1u;
    cp.halo_chunks = ((u32)hi + cp.chunk - 1) / cp.chunk;
    u32 want = 8192 / (DC_TPB * cp.chunk);
    cp.rounds = std::min(32u, std::max(1u, want));
    while (cp.rounds * DC_TPB < 2 * cp.halo_chunks + 1) ++cp.rounds;           // at least half of the staged chunks close ranks
    cp.span = (cp.rounds * DC_TPB * cp.chunk + 4 + 3) & ~3u;
    cp.span_b = (cp.rounds * DC_TPB * cp.chunk + 32 + 15) & ~15u;
    return true;
}

bool make_plan(const DetectParams& prm, bool& use_window, WindowPlan& pl) {
    u64 hi = prm.max_freq;
    if (prm.max_doc_freq) { u64 bound = (u64)prm.max_doc_freq * prm.ndocs; hi = hi ? std::min(hi, bound) : bound; }
    u64 lo = std::max<u64>(2, prm.num_distinct);
    use_window = false;
    if (hi == 0) return true;                       // no limit at all: the leftward walk
    if (hi < lo) return false;
    if (hi - lo + 1 > DW_MAX_SIZES) return true;
    pl.size_lo = (u32)lo; pl.size_hi = (u32)hi;
    pl.halo = ((u32)hi + 3) & ~3u; pl.span = pl.halo + DW_TILE;
    if (pl.span > DW_TPB * 6) return true;
    pl.lev_lo = floor_log2(pl.size_lo - 1); pl.lev_hi = floor_log2(pl.size_hi - 1);
    pl.nkept = pl.lev_hi >= 1 ? pl.lev_hi - std::max(pl.lev_lo, 1u) + 1 : 0;
    const char* force = getenv("MMB_DETECT");
    use_window = !getenv("MMB_FORCE_WALK") && !(force && strcmp(force, "walk") == 0);
    return true;
}

}  // namespace

template <class SAT>
void detect_intervals(const u32* lcp, const SAT* sa, const u8* bwt, const u8* text, u64 m, u64 j_first, const DetectParams& prm,
                      const u64* d_doc_start, u16* thresh, Reports& out, cudaStream_t s, const u8* docs) {
    // the leftward walk is bounded by max_freq, or else by max_doc_freq * ndocs (a document overflows); with neither it is unbounded
    const bool unbounded = prm.max_freq == 0 && prm.max_doc_freq == 0;
    if (unbounded && !(prm.whole_stream && j_first == 0))
        throw Error(31, "unbounded total frequency (-f 0 with -F 0) needs the whole suffix stream in one view (direct route, or one tile)");
    if (prm.max_freq >= (1u << 20)) throw Error(32, "maximum total frequency must be below 2^20");
    if (prm.ndocs > 1024) throw Error(33, "more than 1024 input sequences are not supported");
    out.count = 0;
    DetectArgs a;
    a.lcp = lcp; a.bwt = bwt; a.text = text; a.m = m; a.j_first = j_first; a.prm = prm; a.d_doc_start = d_doc_start; a.thresh = thresh;
    a.docs = prm.ndocs < 255 ? docs : nullptr;
    if (m < 2 || !make_plan(prm, a.use_window, a.plan)) { out.rep.alloc(0, s); return; }
    {
        u64 hi = prm.max_freq;
        if (prm.max_doc_freq) { u64 bound = (u64)prm.max_doc_freq * prm.ndocs; hi = hi ? std::min(hi, bound) : bound; }
        a.use_chunk = hi != 0 && !getenv("MMB_FORCE_WALK") && make_chunk_plan(std::max<u64>(2, prm.num_distinct), hi, a.cplan);
    }
    DevBuf<u32> d_count; d_count.alloc(1, s);
    u64 cap = std::max<u64>(1u << 16, (m - std::min(m, j_first)) / 64);
    DevBuf<Report> raw;
    for (;;) {
        if (cap > 0xfffffff0ull) throw Error(34, "too many matches for one pass");
        raw.alloc(cap, s); d_count.zero();
        a.rep = raw.p; a.cap = (u32)cap; a.count = d_count.p;
        if (prm.ndocs <= 64) dispatch_planes<SAT, 1>(sa, a, s);
        else if (prm.ndocs <= 256) dispatch_planes<SAT, 4>(sa, a, s);
        else dispatch_planes<SAT, 16>(sa, a, s);
        u32 got = to_host(d_count.p, 1, s)[0];
        if (got <= cap) { out.count = got; break; }
        cap = (u64)got + 1024;          // the counter kept counting past the capacity: one retry with the exact size
    }
    u32 n = (u32)out.count;
    out.rep.alloc(n, s);
    if (n == 0) return;
    DevBuf<u64> ka, kb; DevBuf<u32> ia, ib;
    ka.alloc(n, s); kb.alloc(n, s); ia.alloc(n, s); ib.alloc(n, s);
    if (!unbounded) {
        MMB_LAUNCH(report_keys, ceil_div(n, 256), 256, 0, s, raw.p, n, ka.p, ia.p);
        prim::sort_pairs(ka.p, kb.p, ia.p, ib.p, n, 0, 64, s);
    } else {
        // interval sizes may exceed the 20 bits the key has for them: stable sort by size first, then by closing rank
        MMB_LAUNCH(report_size_keys, ceil_div(n, 256), 256, 0, s, raw.p, n, ka.p, ia.p);
        prim::sort_pairs(ka.p, kb.p, ia.p, ib.p, n, 0, 64, s);
        MMB_LAUNCH(report_rank_keys, ceil_div(n, 256), 256, 0, s, raw.p, (const u32*)ib.p, n, ka.p);
        prim::sort_pairs(ka.p, kb.p, ib.p, ia.p, n, 0, 44, s);
        ia.swap(ib);
    }
    MMB_LAUNCH(report_gather, ceil_div(n, 256), 256, 0, s, raw.p, ib.p, n, out.rep.p);
}

template void detect_intervals<u32>(const u32*, const u32*, const u8*, const u8*, u64, u64, const DetectParams&, const u64*, u16*, Reports&, cudaStream_t, const u8*);
template void detect_intervals<u64>(const u32*, const u64*, const u8*, const u8*, u64, u64, const DetectParams&, const u64*, u16*, Reports&, cudaStream_t, const u8*);

}  // namespace mmb
