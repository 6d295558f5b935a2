// mumemto_b200/csrc/merge.cu -- anchor-based merge of K partition results on the device.
//
// Replaces parse_candidate / merge_partitions / fix_neg_strand and the left fold of anchor_merge's main (reference
// src/merge_candidates.cpp:62-157,208-219) -- row J of SURVEY.md section 8.  A partition is (MUM table, uint16 next-best thresholds
// nb[0..L_0]).  The reference folds partitions left to right; one fold step R (+) P is
//   nb'[i] = max(nbR[i], nbP[i]) if both > 0 else 0
//   for every i that starts a MUM in R or in P, with both nb > 0: take the LATEST-STARTED MUM of each side (it is not replaced by an
//   earlier one when it does not reach i: the reference `continue`s), delta = i - start must not exceed its length, rem = len - delta,
//   new = min(remR, remP); emit iff new > nb'[i] and new >= 20 (hard-coded, :141); '+' columns shift by delta, '-' columns by
//   rem - new; the anchor column of P is dropped.
// The fold is kept EXACTLY (the "K-way min/max" restatement of SURVEY.md section 9.7 is not equivalent: MUMs of one partition may
// overlap, and when an intermediate candidate is suppressed the next fold sees an earlier intermediate MUM, not the latest-started
// sources -- DESIGN.md section 6 has the counter-example), but it runs in INDEX SPACE:
//   * what a fold step needs of its left operand is only (start, length) per intermediate MUM and the folded thresholds AT candidate
//     coordinates; the latter are recomputed from the K source arrays on the fly (nbR[i] = all nb_q[i] > 0 ? max : 0 over the
//     partitions folded so far), so no intermediate threshold array is ever written;
//   * an intermediate MUM is (start, length, one source row per partition folded so far); the shifts of the reference telescope:
//     a '+' column ends up at off + (i - start_q), a '-' column at off + (len_q - (i - start_q) - new) with i / new of the FINAL row;
//   * K - 1 tiny fold steps (candidate test, co-ranking of the two sorted start lists by binary search, scan, compaction) with the row
//     count kept on the device -- no host round trip inside the fold -- then ONE pass that writes the merged rows and one pass that
//     writes the merged thresholds.
// The .mums text of the merged table is only formatted when somebody asks for it (warp per row).
// Traffic: 2 K bytes of thresholds per anchor coordinate + 9 bytes per merged cell.
#include "merge.h"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 TPB = 256;
constexpr u32 SCAN_ITEMS = 8;                       // items per thread in the scan kernels
constexpr u32 SCAN_TILE = TPB * SCAN_ITEMS;

// ---------------------------------------------------------------------------------------------- small device-wide exclusive scan
// (three launches, sized for the few million candidates of a fold step; no library call, no host round trip)
__device__ __forceinline__ u32 block_exclusive(u32 mine, u32* total) {
    __shared__ u32 warp_sums[TPB / 32];
    u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, inc = mine;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    u32 before = 0, all = 0;
    for (u32 w = 0; w < TPB / 32; ++w) { if (w < warp) before += warp_sums[w]; all += warp_sums[w]; }
    *total = all;
    __syncthreads();
    return before + inc - mine;
}
__global__ void __launch_bounds__(TPB) scan_tile_sums(const u32* __restrict__ in, u32 n, u32* __restrict__ tile_sum) {
    u32 base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS, s = 0;
    for (u32 k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) s += in[base + k];
    u32 total; block_exclusive(s, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(TPB) scan_tile_offsets(u32* __restrict__ tile_sum, u32 ntiles, u32* __restrict__ grand_total) {
    __shared__ u32 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 t0 = 0; t0 < ntiles; t0 += TPB) {
        u32 t = t0 + threadIdx.x, v = t < ntiles ? tile_sum[t] : 0u, total;
        u32 ex = block_exclusive(v, &total);
        if (t < ntiles) tile_sum[t] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}
__global__ void __launch_bounds__(TPB) scan_apply(const u32* __restrict__ in, u32 n, const u32* __restrict__ tile_off, u32* __restrict__ out) {
    u32 base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS, s = 0, v[SCAN_ITEMS];
    for (u32 k = 0; k < SCAN_ITEMS; ++k) { v[k] = base + k < n ? in[base + k] : 0u; s += v[k]; }
    u32 total, ex = block_exclusive(s, &total) + tile_off[blockIdx.x];
    for (u32 k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}
void exclusive_scan_u32(const u32* in, u32* out, u32 n, u32* total, DevBuf<u32>& tiles, cudaStream_t s) {
    u32 nt = ceil_div(n, SCAN_TILE);
    if (tiles.n < nt) tiles.alloc(nt, s);
    MMB_LAUNCH(scan_tile_sums, nt, TPB, 0, s, in, n, tiles.p);
    MMB_LAUNCH(scan_tile_offsets, 1, TPB, 0, s, tiles.p, nt, total);
    MMB_LAUNCH(scan_apply, nt, TPB, 0, s, in, n, (const u32*)tiles.p, out);
}

// ---------------------------------------------------------------------------------------------- inputs
struct Source {             // one partition as handed in (device pointers, rows in any order)
    const u32* len; const i64* off; const u8* strand; const u16* nb; u32 n, ndocs;
};
constexpr u32 MAX_PARTS = 64;
struct Sources { Source p[MAX_PARTS]; u32 k; };

__global__ void anchor_keys(const i64* __restrict__ off, u32 n, u32 ndocs, u64* __restrict__ keys, u32* __restrict__ idx) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = (u64)off[(u64)i * ndocs]; idx[i] = i; }
}
__global__ void sorted_lengths(const u32* __restrict__ order, const u32* __restrict__ len, u32 n, u32* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = len[order[i]];
}

// index of the last start <= i, or -1
__device__ __forceinline__ i64 latest_start(const u64* __restrict__ starts, u32 n, u64 i) {
    i64 lo = -1, hi = n;                          // starts[lo] <= i < starts[hi]
    while (hi - lo > 1) { i64 mid = (lo + hi) >> 1; if (starts[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

// thresholds of the fold of partitions 0 .. j-1 at coordinate i
__device__ __forceinline__ u16 folded_nb(const Sources& S, u32 j, u64 i) {
    u16 mx = 0;
    for (u32 q = 0; q < j; ++q) { u16 v = S.p[q].nb[i]; if (v == 0) return 0; mx = v > mx ? v : mx; }
    return mx;
}

struct Cand { u32 ia, ib, new_len; };

// One fold step, R (+) partition j.  Thread t < capR handles the start of R's MUM t (if t < *nR); thread t >= capR handles the start
// of P's MUM t - capR unless R starts a MUM at the same coordinate.  The candidate lands at its position in the merged order of the
// two sorted start lists (own index + rank in the other list), so that compaction keeps anchor order without a sort.
__global__ void __launch_bounds__(TPB) fold_step(Sources S, u32 j, const u64* __restrict__ r_start, const u32* __restrict__ r_len, const u32* __restrict__ r_count,
                                                 u32 capR, const u64* __restrict__ p_start, const u32* __restrict__ p_len, u32 np, u64 anchor_len,
                                                 u32* __restrict__ flag, Cand* __restrict__ cand, u64* __restrict__ cand_i) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 nr = *r_count;
    bool from_r = t < capR;
    if (from_r ? t >= nr : t - capR >= np) return;
    u64 i = from_r ? r_start[t] : p_start[t - capR];
    if (i >= anchor_len) return;
    i64 ia = from_r ? (i64)t : latest_start(r_start, nr, i);
    i64 ib = from_r ? latest_start(p_start, np, i) : (i64)(t - capR);
    if (ia < 0 || ib < 0) return;
    if (!from_r && r_start[ia] == i) return;                    // the R-side thread owns this coordinate
    // several MUMs with one start offset would make the reference step through them one coordinate late; MUM starts are unique.
    u16 x = folded_nb(S, j, i), y = S.p[j].nb[i];
    if (!(x > 0 && y > 0)) return;
    u64 da = i - r_start[ia], db = i - p_start[ib];
    if (da > r_len[ia] || db > p_len[ib]) return;
    u32 ra = r_len[ia] - (u32)da, rb = p_len[ib] - (u32)db, nl = ra < rb ? ra : rb;
    u16 nbmax = x > y ? x : y;
    if (!(nl > nbmax && nl >= 20)) return;
    // position in the merged start order: R's elements before it + P's elements before it
    u32 pos = from_r ? t + (u32)(ib + 1) - (p_start[ib] == i ? 1u : 0u)       // P starts strictly below i
                     : (t - capR) + (u32)(ia + 1);                            // R starts <= i (none equal)
    // when both sides start at i the R-side element is placed first and the P-side slot stays empty
    Cand c; c.ia = (u32)ia; c.ib = (u32)ib; c.new_len = nl;
    cand[pos] = c; cand_i[pos] = i; flag[pos] = 1;
}

// kept candidates -> the next intermediate table (start, length, one source row per partition folded so far)
__global__ void __launch_bounds__(TPB) fold_compact(const u32* __restrict__ flag, const u32* __restrict__ slot, u32 n, const Cand* __restrict__ cand,
                                                    const u64* __restrict__ cand_i, u32 j, const u32* __restrict__ r_src, u32 cap,
                                                    const u32* __restrict__ p_order, u64* __restrict__ o_start, u32* __restrict__ o_len,
                                                    u32* __restrict__ o_src) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n || !flag[t]) return;
    u32 s = slot[t];
    Cand c = cand[t];
    o_start[s] = cand_i[t]; o_len[s] = c.new_len;
    for (u32 q = 0; q < j; ++q) o_src[(u64)q * cap + s] = r_src[(u64)q * cap + c.ia];
    o_src[(u64)j * cap + s] = p_order[c.ib];
}
__global__ void init_first(const u64* __restrict__ starts, const u32* __restrict__ lens, const u32* __restrict__ order, u32 n, u64* __restrict__ o_start,
                           u32* __restrict__ o_len, u32* __restrict__ o_src, u32* __restrict__ count) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *count = n;
    if (t < n) { o_start[t] = starts[t]; o_len[t] = lens[t]; o_src[t] = order[t]; }
}

// merged thresholds: one pass over the K source arrays
__global__ void __launch_bounds__(TPB) merged_thresholds(Sources S, u64 n, u16* __restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = folded_nb(S, S.k, i);
}

// merged rows: thread per cell
__global__ void __launch_bounds__(TPB) merged_rows(Sources S, const u32* __restrict__ count, const u64* __restrict__ f_start, const u32* __restrict__ f_len,
                                                   const u32* __restrict__ f_src, u32 cap, u32 cols, const u16* __restrict__ col_part,
                                                   const u16* __restrict__ col_local, u32* __restrict__ len_o, i64* __restrict__ off_o,
                                                   u8* __restrict__ plus_o, u64* __restrict__ start_o) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 n = *count;
    if (t >= (u64)n * cols) return;
    u32 r = (u32)(t / cols), c = (u32)(t % cols);
    u32 q = col_part[c], lc = col_local[c];
    const Source& P = S.p[q];
    u32 row = f_src[(u64)q * cap + r];
    u64 i = f_start[r]; u32 nl = f_len[r];
    const i64* po = P.off + (u64)row * P.ndocs;
    u64 delta = i - (u64)po[0];
    u8 sv = P.strand[(u64)row * P.ndocs + lc];
    u8 p = (sv == 1 || sv == '+') ? 1 : 0;
    off_o[t] = po[lc] + (p ? (i64)delta : (i64)P.len[row] - (i64)delta - (i64)nl);
    plus_o[t] = p;
    if (c == 0) { len_o[r] = nl; start_o[r] = i; }
}

// ---- text: serialize_mum, reference include/mumsio.hpp:283-296:  len \t o,o,... \t +,-,... \n      (one warp per row)
__device__ __forceinline__ u32 ndigits(u64 x) { u32 d = 1; while (x >= 10) { x /= 10; ++d; } return d; }
__device__ __forceinline__ u32 nchars(i64 v) { return v < 0 ? 1 + ndigits((u64)(-v)) : ndigits((u64)v); }
__device__ __forceinline__ void put_signed(u8* dst, i64 v, u32 w) {
    u64 x = v < 0 ? (u64)(-v) : (u64)v;
    u32 first = v < 0 ? 1u : 0u;
    if (first) dst[0] = '-';
    for (u32 i = w; i-- > first;) { dst[i] = (u8)('0' + x % 10); x /= 10; }
}

template <bool WRITE>
__global__ void __launch_bounds__(TPB) merged_lines(const u32* __restrict__ len, const i64* __restrict__ off, const u8* __restrict__ plus, u32 n, u32 cols,
                                                    u32* __restrict__ line_len, const u64* __restrict__ line_off, u8* __restrict__ text) {
    u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= n) return;                                     // whole warp
    const i64* ro = off + (u64)r * cols; const u8* rp = plus + (u64)r * cols;
    u32 dl = ndigits(len[r]);
    if (!WRITE) {
        u32 w = 0;
        for (u32 c = lane; c < cols; c += 32) w += nchars(ro[c]);
#pragma unroll
        for (u32 d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
        if (lane == 0) line_len[r] = dl + 1 + w + (cols - 1) + 1 + cols + (cols - 1) + 1;
        return;
    }
    u8* d = text + line_off[r];
    if (lane == 0) { put_signed(d, (i64)len[r], dl); d[dl] = '\t'; }
    u32 at = dl + 1;                                        // running position inside the line (uniform over the warp)
    for (u32 c0 = 0; c0 < cols; c0 += 32) {
        u32 c = c0 + lane;
        i64 v = c < cols ? ro[c] : 0;
        u32 w = c < cols ? nchars(v) + (c + 1 < cols ? 1u : 0u) : 0u, inc = w;      // digits + the comma that follows
#pragma unroll
        for (u32 k = 1; k < 32; k <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, inc, k); if (lane >= k) inc += o; }
        if (c < cols) {
            u8* q = d + at + inc - w;
            u32 nd = w - (c + 1 < cols ? 1u : 0u);
            put_signed(q, v, nd);
            if (c + 1 < cols) q[nd] = ',';
        }
        at += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) d[at] = '\t';
    ++at;
    for (u32 c = lane; c < cols; c += 32) { d[at + 2 * c] = rp[c] ? '+' : '-'; d[at + 2 * c + 1] = c + 1 < cols ? ',' : '\n'; }
}

u32 bits_for(u64 x) { u32 b = 1; while ((1ull << b) <= x && b < 63) ++b; return b; }

}  // namespace

void merge_partitions(const MergeSource* src, size_t k, u64 anchor_len, MergePartition& R, cudaStream_t s) {
    if (k < 2 || k > MAX_PARTS) throw Error(61, "anchor merge takes 2 .. 64 partitions");
    Sources S; S.k = (u32)k;
    u64 cap64 = 0; u32 cols = 0;
    for (size_t p = 0; p < k; ++p) {
        if (src[p].n > 0xfffffff0ull) throw Error(60, "partition too large");
        if (src[p].ndocs < 2 || src[p].ndocs > 0xffffu) throw Error(61, "a partition needs at least two sequences");
        S.p[p] = Source{src[p].len, src[p].off, src[p].strand, src[p].nb, (u32)src[p].n, src[p].ndocs};
        cap64 += src[p].n; cols += src[p].ndocs - (p ? 1 : 0);
    }
    if (cap64 > 0xfffffff0ull) throw Error(60, "partitions too large");
    u32 cap = (u32)std::max<u64>(cap64, 1);
    R.ndocs = cols; R.n = 0;
    R.nb.alloc(anchor_len, s);
    if (anchor_len) MMB_LAUNCH(merged_thresholds, ceil_div(anchor_len, TPB), TPB, 0, s, S, anchor_len, R.nb.p);

    // ---- every partition: rows ordered by anchor offset (start list, lengths and original row in that order)
    std::vector<DevBuf<u64>> starts(k); std::vector<DevBuf<u32>> lens(k), order(k);
    int key_bits = (int)bits_for(anchor_len ? anchor_len : 1);
    for (size_t p = 0; p < k; ++p) {
        u32 n = S.p[p].n;
        starts[p].alloc(n, s); lens[p].alloc(n, s); order[p].alloc(n, s);
        if (!n) continue;
        DevBuf<u64> ka; DevBuf<u32> ia;
        ka.alloc(n, s); ia.alloc(n, s);
        MMB_LAUNCH(anchor_keys, ceil_div(n, TPB), TPB, 0, s, S.p[p].off, n, S.p[p].ndocs, ka.p, ia.p);
        prim::sort_pairs(ka.p, starts[p].p, ia.p, order[p].p, n, 0, std::min(64, key_bits + 1), s);
        MMB_LAUNCH(sorted_lengths, ceil_div(n, TPB), TPB, 0, s, (const u32*)order[p].p, S.p[p].len, n, lens[p].p);
    }

    // ---- the fold, in index space; the row count of the intermediate table stays on the device
    DevBuf<u64> st_a, st_b, cand_i; DevBuf<u32> ln_a, ln_b, src_a, src_b, count_a, count_b, flag, slot, tiles; DevBuf<Cand> cand;
    st_a.alloc(cap, s); st_b.alloc(cap, s); ln_a.alloc(cap, s); ln_b.alloc(cap, s); src_a.alloc((u64)cap * k, s); src_b.alloc((u64)cap * k, s);
    count_a.alloc(1, s); count_b.alloc(1, s); flag.alloc(cap, s); slot.alloc(cap, s); cand.alloc(cap, s); cand_i.alloc(cap, s);
    MMB_LAUNCH(init_first, ceil_div(std::max(S.p[0].n, 1u), TPB), TPB, 0, s, (const u64*)starts[0].p, (const u32*)lens[0].p, (const u32*)order[0].p, S.p[0].n,
               st_a.p, ln_a.p, src_a.p, count_a.p);
    u32 capR = S.p[0].n;                 // upper bound of the intermediate row count
    for (u32 j = 1; j < k; ++j) {
        u32 np = S.p[j].n, width = capR + np;
        if (width == 0) { MMB_CUDA(cudaMemsetAsync(count_b.p, 0, 4, s)); }
        else {
            MMB_CUDA(cudaMemsetAsync(flag.p, 0, (size_t)width * 4, s));
            MMB_LAUNCH(fold_step, ceil_div(width, TPB), TPB, 0, s, S, j, (const u64*)st_a.p, (const u32*)ln_a.p, (const u32*)count_a.p, capR,
                       (const u64*)starts[j].p, (const u32*)lens[j].p, np, anchor_len, flag.p, cand.p, cand_i.p);
            exclusive_scan_u32(flag.p, slot.p, width, count_b.p, tiles, s);
            MMB_LAUNCH(fold_compact, ceil_div(width, TPB), TPB, 0, s, (const u32*)flag.p, (const u32*)slot.p, width, (const Cand*)cand.p, (const u64*)cand_i.p, j,
                       (const u32*)src_a.p, cap, (const u32*)order[j].p, st_b.p, ln_b.p, src_b.p);
        }
        st_a.swap(st_b); ln_a.swap(ln_b); src_a.swap(src_b); count_a.swap(count_b);
        capR = std::min<u64>(width, cap);      // every kept candidate sits at a distinct start coordinate of R or P
    }

    // ---- the merged table: one pass
    u32 n = to_host(count_a.p, 1, s)[0];
    R.n = n;
    R.len.alloc(n, s); R.off.alloc((u64)n * cols, s); R.plus.alloc((u64)n * cols, s); R.start.alloc(n, s);
    if (!n) return;
    std::vector<u16> h_part(cols), h_local(cols);
    u32 c = 0;
    for (u32 p = 0; p < k; ++p)
        for (u32 lc = (p ? 1 : 0); lc < S.p[p].ndocs; ++lc, ++c) { h_part[c] = (u16)p; h_local[c] = (u16)lc; }
    DevBuf<u16> col_part, col_local; col_part.alloc(cols, s); col_local.alloc(cols, s);
    MMB_CUDA(cudaMemcpyAsync(col_part.p, h_part.data(), cols * 2, cudaMemcpyHostToDevice, s));
    MMB_CUDA(cudaMemcpyAsync(col_local.p, h_local.data(), cols * 2, cudaMemcpyHostToDevice, s));
    MMB_LAUNCH(merged_rows, ceil_div((u64)n * cols, TPB), TPB, 0, s, S, (const u32*)count_a.p, (const u64*)st_a.p, (const u32*)ln_a.p, (const u32*)src_a.p, cap, cols,
               (const u16*)col_part.p, (const u16*)col_local.p, R.len.p, R.off.p, R.plus.p, R.start.p);
    MMB_CUDA(cudaStreamSynchronize(s));       // h_part / h_local live on this stack frame
}

void merge_text(const MergePartition& P, DevBuf<u8>& text, u64& bytes, cudaStream_t s) {
    bytes = 0;
    if (!P.n) { text.alloc(0, s); return; }
    DevBuf<u32> ll; DevBuf<u64> lo; ll.alloc(P.n + 1, s); lo.alloc(P.n + 1, s);
    MMB_CUDA(cudaMemsetAsync(ll.p + P.n, 0, 4, s));
    auto merged_lines_len_k = merged_lines<false>; auto merged_lines_write_k = merged_lines<true>;
    MMB_LAUNCH(merged_lines_len_k, ceil_div((u64)P.n * 32, TPB), TPB, 0, s, P.len.p, P.off.p, P.plus.p, P.n, P.ndocs, ll.p, (const u64*)nullptr, (u8*)nullptr);
    prim::exclusive_sum(ll.p, lo.p, P.n + 1, s);
    bytes = to_host(lo.p + P.n, 1, s)[0];
    text.alloc(bytes, s);
    MMB_LAUNCH(merged_lines_write_k, ceil_div((u64)P.n * 32, TPB), TPB, 0, s, P.len.p, P.off.p, P.plus.p, P.n, P.ndocs, ll.p, lo.p, text.p);
}

}  // namespace mmb
