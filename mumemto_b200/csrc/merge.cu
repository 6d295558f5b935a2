// mumemto_b200/csrc/merge.cu -- anchor-based merge of partition results on the device.
//
// Replaces parse_candidate / merge_partitions / fix_neg_strand and the left fold of anchor_merge's main (reference
// src/merge_candidates.cpp:62-157,208-219) -- row J of SURVEY.md section 8.  A partition is (MUM table sorted by anchor offset,
// uint16 next-best thresholds nb[0..L_0]).  One fold step A (+) B:
//   nb'[i] = max(nbA[i], nbB[i]) if both > 0 else 0                                            (one thread per anchor coordinate)
//   for every i that starts a MUM in A or in B, with both nb > 0: take the latest-started MUM of each side (binary search in the
//   sorted start offsets), delta = i - start must not exceed its length, rem = len - delta, new = min(remA, remB); emit iff
//   new > nb'[i] and new >= 20 (hard-coded in the reference, :141); '+' columns shift by delta, '-' columns by rem - new; the
//   anchor column of B is dropped.  Rows come out in increasing i: candidates are keyed by i and radix-sorted.
// The reference folds partitions left to right; so do we (the fold is what defines column order and the `continue` quirks).
// Traffic: 2 K bytes of thresholds per anchor coordinate + the MUM tables.
#include "merge.h"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 TPB = 256;

__global__ void anchor_keys(const i64* __restrict__ off, u32 n, u32 ndocs, u64* __restrict__ keys, u32* __restrict__ idx) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = (u64)off[(u64)i * ndocs]; idx[i] = i; }
}

__global__ void anchor_keys_iota(u32 n, u32* __restrict__ idx) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = i;
}

__global__ void gather_rows(const u32* __restrict__ idx, u32 n, u32 ndocs, const u32* __restrict__ len, const i64* __restrict__ off,
                            const u8* __restrict__ strand, u32* __restrict__ len_o, i64* __restrict__ off_o, u8* __restrict__ plus_o,
                            u64* __restrict__ start_o) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (u64)n * ndocs) return;
    u32 r = (u32)(t / ndocs), d = (u32)(t % ndocs);
    u64 src = (u64)idx[r] * ndocs + d;
    off_o[t] = off[src];
    plus_o[t] = (strand[src] == 1 || strand[src] == '+') ? 1 : 0;
    if (d == 0) { len_o[r] = len[idx[r]]; start_o[r] = (u64)off[src]; }
}

__global__ void merge_thresholds(const u16* __restrict__ a, const u16* __restrict__ b, u64 n, u16* __restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u16 x = a[i], y = b[i];
    out[i] = (x > 0 && y > 0) ? (x > y ? x : y) : (u16)0;
}

// index of the last start <= i, or -1
__device__ __forceinline__ i64 latest_start(const u64* __restrict__ starts, u32 n, u64 i) {
    i64 lo = -1, hi = n;                          // starts[lo] <= i < starts[hi]
    while (hi - lo > 1) { i64 mid = (lo + hi) >> 1; if (starts[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

struct Cand { u32 ia, ib, new_len, da, db; };

// thread t < na handles the start of A's MUM t; thread t >= na handles B's MUM t - na unless A starts a MUM at the same coordinate
__global__ void fold_candidates(const u64* __restrict__ sa, const u32* __restrict__ la, u32 na, const u64* __restrict__ sb, const u32* __restrict__ lb,
                                u32 nb_, const u16* __restrict__ nba, const u16* __restrict__ nbb, u64 anchor_len, u64* __restrict__ keys,
                                u32* __restrict__ valid, Cand* __restrict__ cand) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= na + nb_) return;
    bool from_a = t < na;
    u64 i = from_a ? sa[t] : sb[t - na];
    keys[t] = i; valid[t] = 0;
    if (i >= anchor_len) return;
    i64 ia = from_a ? (i64)t : latest_start(sa, na, i);
    i64 ib = from_a ? latest_start(sb, nb_, i) : (i64)(t - na);
    if (ia < 0 || ib < 0) return;
    if (!from_a && sa[ia] == i) return;                         // the A-side thread owns this coordinate
    // several MUMs with one start offset would make the reference step through them one coordinate late; MUM starts are unique.
    u16 x = nba[i], y = nbb[i];
    if (!(x > 0 && y > 0)) return;
    u64 da = i - sa[ia], db = i - sb[ib];
    if (da > la[ia] || db > lb[ib]) return;
    u32 ra = la[ia] - (u32)da, rb = lb[ib] - (u32)db, nl = ra < rb ? ra : rb;
    u16 nbmax = x > y ? x : y;
    if (!(nl > nbmax && nl >= 20)) return;
    Cand c; c.ia = (u32)ia; c.ib = (u32)ib; c.new_len = nl; c.da = (u32)da; c.db = (u32)db;
    cand[t] = c; valid[t] = 1;
}

__global__ void compact_candidates(const u32* __restrict__ order, const u32* __restrict__ valid, const u32* __restrict__ slot, u32 n,
                                   const Cand* __restrict__ cand, Cand* __restrict__ out) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;     // t walks candidates in sorted (anchor coordinate) order
    if (t < n && valid[t]) out[slot[t]] = cand[order[t]];
}
__global__ void gather_valid(const u32* __restrict__ order, const u32* __restrict__ valid_in, u32 n, u32* __restrict__ valid_sorted) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) valid_sorted[t] = valid_in[order[t]];
}

__global__ void fold_rows(const Cand* __restrict__ cand, u32 n, const u32* __restrict__ la, const i64* __restrict__ oa, const u8* __restrict__ pa, u32 da_cols,
                          const u32* __restrict__ lb, const i64* __restrict__ ob, const u8* __restrict__ pb, u32 db_cols, u32* __restrict__ len_o,
                          i64* __restrict__ off_o, u8* __restrict__ plus_o, u64* __restrict__ start_o) {
    u32 cols = da_cols + db_cols - 1;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (u64)n * cols) return;
    u32 r = (u32)(t / cols), c = (u32)(t % cols);
    Cand k = cand[r];
    i64 o; u8 p;
    if (c < da_cols) {
        u64 src = (u64)k.ia * da_cols + c;
        p = pa[src]; o = oa[src] + (p ? (i64)k.da : (i64)(la[k.ia] - k.da) - (i64)k.new_len);
    } else {
        u64 src = (u64)k.ib * db_cols + (c - da_cols + 1);
        p = pb[src]; o = ob[src] + (p ? (i64)k.db : (i64)(lb[k.ib] - k.db) - (i64)k.new_len);
    }
    off_o[t] = o; plus_o[t] = p;
    if (c == 0) { len_o[r] = k.new_len; start_o[r] = (u64)o; }
}

// ---- text: serialize_mum, reference include/mumsio.hpp:283-296:  len \t o,o,... \t +,-,... \n
__device__ __forceinline__ u32 ndigits(u64 x) { u32 d = 1; while (x >= 10) { x /= 10; ++d; } return d; }
__device__ __forceinline__ u32 nchars(i64 v) { return v < 0 ? 1 + ndigits((u64)(-v)) : ndigits((u64)v); }
__device__ __forceinline__ void put_signed(u8* dst, i64 v, u32 w) {
    u64 x = v < 0 ? (u64)(-v) : (u64)v;
    u32 first = v < 0 ? 1u : 0u;
    if (first) dst[0] = '-';
    for (u32 i = w; i-- > first;) { dst[i] = (u8)('0' + x % 10); x /= 10; }
}

template <bool WRITE>
__global__ void merged_lines(const u32* __restrict__ len, const i64* __restrict__ off, const u8* __restrict__ plus, u32 n, u32 cols,
                             u32* __restrict__ line_len, const u64* __restrict__ line_off, u8* __restrict__ text) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;     // one thread per row: merged tables are small next to the threshold pass
    if (r >= n) return;
    const i64* ro = off + (u64)r * cols; const u8* rp = plus + (u64)r * cols;
    u32 dl = ndigits(len[r]);
    if (!WRITE) {
        u32 w = dl + 1;
        for (u32 c = 0; c < cols; ++c) w += nchars(ro[c]);
        line_len[r] = w + (cols - 1) + 1 + cols + (cols - 1) + 1;
        return;
    }
    u8* d = text + line_off[r];
    put_signed(d, (i64)len[r], dl); d += dl; *d++ = '\t';
    for (u32 c = 0; c < cols; ++c) { u32 w = nchars(ro[c]); put_signed(d, ro[c], w); d += w; if (c + 1 < cols) *d++ = ','; }
    *d++ = '\t';
    for (u32 c = 0; c < cols; ++c) { *d++ = rp[c] ? '+' : '-'; if (c + 1 < cols) *d++ = ','; }
    *d = '\n';
}

u32 tail_count(const u32* idx, const u32* flag, u32 n, cudaStream_t s) {
    if (!n) return 0;
    u32 h[2];
    MMB_CUDA(cudaMemcpyAsync(&h[0], idx + n - 1, 4, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaMemcpyAsync(&h[1], flag + n - 1, 4, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaStreamSynchronize(s));
    return h[0] + h[1];
}

}  // namespace

void merge_prepare(const u32* len, const i64* off, const u8* strand, u64 n_mums, u32 ndocs, const u16* nb, u64 anchor_len, MergePartition& out,
                   cudaStream_t s) {
    if (n_mums > 0xfffffff0ull) throw Error(60, "partition too large");
    u32 n = (u32)n_mums;
    out.n = n; out.ndocs = ndocs;
    out.len.alloc(n, s); out.off.alloc((u64)n * ndocs, s); out.plus.alloc((u64)n * ndocs, s); out.start.alloc(n, s);
    out.nb.alloc(anchor_len, s);
    MMB_CUDA(cudaMemcpyAsync(out.nb.p, nb, anchor_len * 2, cudaMemcpyDeviceToDevice, s));
    if (!n) return;
    DevBuf<u64> ka, kb; DevBuf<u32> ia, ib;
    ka.alloc(n, s); kb.alloc(n, s); ia.alloc(n, s); ib.alloc(n, s);
    MMB_LAUNCH(anchor_keys, ceil_div(n, TPB), TPB, 0, s, off, n, ndocs, ka.p, ia.p);
    prim::sort_pairs(ka.p, kb.p, ia.p, ib.p, n, 0, 64, s);
    MMB_LAUNCH(gather_rows, ceil_div((u64)n * ndocs, TPB), TPB, 0, s, ib.p, n, ndocs, len, off, strand, out.len.p, out.off.p, out.plus.p, out.start.p);
}

void merge_fold(const MergePartition& A, const MergePartition& B, u64 anchor_len, MergePartition& R, cudaStream_t s) {
    R.ndocs = A.ndocs + B.ndocs - 1;
    R.nb.alloc(anchor_len, s);
    if (anchor_len) MMB_LAUNCH(merge_thresholds, ceil_div(anchor_len, TPB), TPB, 0, s, A.nb.p, B.nb.p, anchor_len, R.nb.p);
    u32 total = A.n + B.n;
    R.n = 0;
    if (A.n == 0 || B.n == 0) { R.len.alloc(0, s); R.off.alloc(0, s); R.plus.alloc(0, s); R.start.alloc(0, s); return; }
    DevBuf<u64> ka, kb; DevBuf<u32> valid, ia, ib, vs, slot; DevBuf<Cand> cand, kept;
    ka.alloc(total, s); kb.alloc(total, s); valid.alloc(total, s); ia.alloc(total, s); ib.alloc(total, s); vs.alloc(total, s); slot.alloc(total, s);
    cand.alloc(total, s);
    MMB_LAUNCH(fold_candidates, ceil_div(total, TPB), TPB, 0, s, A.start.p, A.len.p, A.n, B.start.p, B.len.p, B.n, A.nb.p, B.nb.p, anchor_len, ka.p,
               valid.p, cand.p);
    // order candidates by anchor coordinate (A's and B's start lists are each sorted; a radix sort merges them)
    MMB_LAUNCH(anchor_keys_iota, ceil_div(total, TPB), TPB, 0, s, total, ia.p);
    prim::sort_pairs(ka.p, kb.p, ia.p, ib.p, total, 0, 64, s);
    MMB_LAUNCH(gather_valid, ceil_div(total, TPB), TPB, 0, s, ib.p, valid.p, total, vs.p);
    prim::exclusive_sum(vs.p, slot.p, total, s);
    u32 n = tail_count(slot.p, vs.p, total, s);
    R.n = n;
    R.len.alloc(n, s); R.off.alloc((u64)n * R.ndocs, s); R.plus.alloc((u64)n * R.ndocs, s); R.start.alloc(n, s);
    if (!n) return;
    kept.alloc(n, s);
    MMB_LAUNCH(compact_candidates, ceil_div(total, TPB), TPB, 0, s, ib.p, vs.p, slot.p, total, cand.p, kept.p);
    MMB_LAUNCH(fold_rows, ceil_div((u64)n * R.ndocs, TPB), TPB, 0, s, kept.p, n, A.len.p, A.off.p, A.plus.p, A.ndocs, B.len.p, B.off.p, B.plus.p, B.ndocs,
               R.len.p, R.off.p, R.plus.p, R.start.p);
}

void merge_text(const MergePartition& P, DevBuf<u8>& text, u64& bytes, cudaStream_t s) {
    bytes = 0;
    if (!P.n) { text.alloc(0, s); return; }
    DevBuf<u32> ll; DevBuf<u64> lo; ll.alloc(P.n + 1, s); lo.alloc(P.n + 1, s);
    MMB_CUDA(cudaMemsetAsync(ll.p + P.n, 0, 4, s));
    auto merged_lines_len_k = merged_lines<false>; auto merged_lines_write_k = merged_lines<true>;
    MMB_LAUNCH(merged_lines_len_k, ceil_div(P.n, TPB), TPB, 0, s, P.len.p, P.off.p, P.plus.p, P.n, P.ndocs, ll.p, (const u64*)nullptr, (u8*)nullptr);
    prim::exclusive_sum(ll.p, lo.p, P.n + 1, s);
    bytes = to_host(lo.p + P.n, 1, s)[0];
    text.alloc(bytes, s);
    MMB_LAUNCH(merged_lines_write_k, ceil_div(P.n, TPB), TPB, 0, s, P.len.p, P.off.p, P.plus.p, P.n, P.ndocs, ll.p, lo.p, text.p);
}

}  // namespace mmb
