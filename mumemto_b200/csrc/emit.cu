// mumemto_b200/csrc/emit.cu -- turn reported intervals into the reference's output records and text lines.
//
// Replaces mem_finder::write_mum (reference include/mem_finder.hpp:357-428), write_mem (:210-263) and the strand packing of
// write_bums (:460-476) -- rows G and H of SURVEY.md section 8, with the quirks listed there kept on purpose:
//   MUM mode: '-' occurrences use coordinate 2(L+1)-q-len-1; the whole MUM is dropped if a '-' occurrence reaches its doc's final
//   '$' (q+len >= 2(L+1)); dropped if the lowest-numbered document present is on '-'; '+' occurrences are not checked against '$'.
//   MEM mode: occurrences in rank order, no filter, unsigned 64-bit coordinates (they wrap), and the LAST occurrence of a line
//   uses 2(L+1)-q-len (no -1).
// One warp per interval.  Text is produced in two passes (line lengths -> exclusive scan -> bytes) so that the .mums/.mems file is
// a single device buffer copied out once.
#include "detect.h"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 TPB = 128;                 // 4 warps = 4 intervals per block
constexpr u32 FULL = 0xffffffffu;

__device__ __forceinline__ u32 find_doc(const u64* __restrict__ doc_start, u32 ndocs, u64 p) {
    u32 lo = 0, hi = ndocs;
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (doc_start[mid] <= p) lo = mid; else hi = mid; }
    return lo;
}
__device__ __forceinline__ u32 ndigits(u64 x) { u32 d = 1; while (x >= 10) { x /= 10; ++d; } return d; }
__device__ __forceinline__ void put_number(u8* __restrict__ dst, u64 x, u32 digits) {
    for (u32 i = digits; i-- > 0;) { dst[i] = (u8)('0' + x % 10); x /= 10; }
}
__device__ __forceinline__ u32 warp_inclusive_sum(u32 v, u32 lane) {
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(FULL, v, d); if (lane >= d) v += o; }
    return v;
}
__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
    for (u32 d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

struct Occ { u32 doc; u64 pos; bool neg, overrun; };
// symbol offset inside the document -> printed coordinate; minus_one = false only for the last occurrence of a MEM line
__device__ __forceinline__ Occ locate(u64 p, u32 len, const u64* __restrict__ doc_start, const u64* __restrict__ doc_len, u32 ndocs, int revcomp,
                                      bool minus_one) {
    Occ o; o.doc = find_doc(doc_start, ndocs, p);
    u64 q = p - doc_start[o.doc], half = doc_len[o.doc] + 1;
    o.neg = revcomp && q >= half; o.overrun = false; o.pos = q;
    if (o.neg) { o.overrun = q + len >= 2 * half; o.pos = 2 * half - q - len - (minus_one ? 1 : 0); }
    return o;
}

// ------------------------------------------------------------------------------------------------ MUM mode
template <class SAT>
__global__ void __launch_bounds__(TPB) mum_validate(const Report* __restrict__ rep, u32 n, const SAT* __restrict__ sa, const u64* __restrict__ doc_start,
                                                    const u64* __restrict__ doc_len, u32 ndocs, int revcomp, u32* __restrict__ valid) {
    u32 lane = threadIdx.x & 31, r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;                                  // whole warp leaves together
    Report R = rep[r];
    u64 e = (R.key >> 20) - 1;
    bool bad = false; u32 best = 0xffffffffu;            // best = lowest doc seen << 1 | neg
    for (u64 i = R.s + lane; i <= e; i += 32) {
        Occ o = locate((u64)sa[i], R.len, doc_start, doc_len, ndocs, revcomp, true);
        bad |= o.overrun;
        best = min(best, (o.doc << 1) | (u32)o.neg);
    }
#pragma unroll
    for (u32 d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(FULL, best, d));
    bad = __any_sync(FULL, bad);
    if (lane == 0) valid[r] = (!bad && !(best & 1)) ? 1u : 0u;
}

__global__ void compact_reports(const Report* __restrict__ rep, const u32* __restrict__ valid, const u32* __restrict__ idx, u32 n, Report* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && valid[i]) out[idx[i]] = rep[i];
}

template <class SAT>
__global__ void __launch_bounds__(TPB) mum_rows_kernel(const Report* __restrict__ rep, u32 n, const SAT* __restrict__ sa, const u64* __restrict__ doc_start,
                                                const u64* __restrict__ doc_len, u32 ndocs, int revcomp, u32* __restrict__ len_out,
                                                i64* __restrict__ off, u8* __restrict__ strand) {
    u32 lane = threadIdx.x & 31, r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    Report R = rep[r];
    u64 e = (R.key >> 20) - 1;
    if (lane == 0) len_out[r] = R.len;
    for (u64 i = R.s + lane; i <= e; i += 32) {          // each document occurs at most once in MUM mode: no write conflicts
        Occ o = locate((u64)sa[i], R.len, doc_start, doc_len, ndocs, revcomp, true);
        off[(u64)r * ndocs + o.doc] = (i64)o.pos;
        strand[(u64)r * ndocs + o.doc] = o.neg ? '-' : '+';
    }
}

// line = len \t o_0,o_1,...,o_{N-1} \t s_0,...,s_{N-1} \n   (absent documents leave an empty field)
template <bool WRITE>
__global__ void __launch_bounds__(TPB) mum_lines(const u32* __restrict__ len, const i64* __restrict__ off, const u8* __restrict__ strand, u32 n, u32 ndocs,
                                                 u32* __restrict__ line_len, const u64* __restrict__ line_off, u8* __restrict__ text) {
    u32 lane = threadIdx.x & 31, r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    u32 L = len[r], dl = ndigits(L);
    const i64* ro = off + (u64)r * ndocs;
    const u8* rs = strand + (u64)r * ndocs;
    u32 width_off = 0, present = 0;
    for (u32 d = lane; d < ndocs; d += 32) { i64 o = ro[d]; if (o >= 0) { width_off += ndigits((u64)o); ++present; } }
    width_off = warp_sum(width_off); present = warp_sum(present);
    u32 total = dl + 1 + width_off + (ndocs - 1) + 1 + present + (ndocs - 1) + 1;
    if (!WRITE) { if (lane == 0) line_len[r] = total; return; }
    u8* dst = text + line_off[r];
    u32 at_off = dl + 1, at_str = dl + 1 + width_off + (ndocs - 1) + 1;
    if (lane == 0) { put_number(dst, L, dl); dst[dl] = '\t'; dst[at_str - 1] = '\t'; dst[total - 1] = '\n'; }
    for (u32 base = 0; base < ndocs; base += 32) {
        u32 d = base + lane;
        i64 o = (d < ndocs) ? ro[d] : -1;
        u32 w = (o >= 0) ? ndigits((u64)o) : 0, pr = (o >= 0) ? 1u : 0u;
        u32 inc_w = warp_inclusive_sum(w, lane), inc_p = warp_inclusive_sum(pr, lane);
        if (d < ndocs) {
            u32 po = at_off + (inc_w - w) + d, ps = at_str + (inc_p - pr) + d;     // d commas precede field d
            if (o >= 0) { put_number(dst + po, (u64)o, w); dst[ps] = rs[d]; }
            if (d + 1 < ndocs) { dst[po + w] = ','; dst[ps + pr] = ','; }
        }
        at_off += __shfl_sync(FULL, inc_w, 31); at_str += __shfl_sync(FULL, inc_p, 31);
    }
}

// ------------------------------------------------------------------------------------------------ MEM mode
__global__ void mem_sizes(const Report* __restrict__ rep, u32 n, u64* __restrict__ occ) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) occ[i] = (rep[i].key >> 20) - rep[i].s;                     // e - s + 1
    if (i == n) occ[i] = 0;
}

template <class SAT>
__global__ void __launch_bounds__(TPB) mem_rows_kernel(const Report* __restrict__ rep, u32 n, const SAT* __restrict__ sa, const u64* __restrict__ doc_start,
                                                const u64* __restrict__ doc_len, u32 ndocs, int revcomp, const u64* __restrict__ first,
                                                u32* __restrict__ len_out, u64* __restrict__ off, u32* __restrict__ doc, u8* __restrict__ strand) {
    u32 lane = threadIdx.x & 31, r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    Report R = rep[r];
    u64 e = (R.key >> 20) - 1, base = first[r];
    if (lane == 0) len_out[r] = R.len;
    for (u64 i = R.s + lane; i <= e; i += 32) {
        Occ o = locate((u64)sa[i], R.len, doc_start, doc_len, ndocs, revcomp, i != e);
        u64 slot = base + (i - R.s);
        off[slot] = o.pos; doc[slot] = o.doc; strand[slot] = o.neg ? '-' : '+';
    }
}

// line = len \t p_0,...,p_{k-1} \t d_0,...,d_{k-1} \t s_0,...,s_{k-1} \n
template <bool WRITE>
__global__ void __launch_bounds__(TPB) mem_lines(const u32* __restrict__ len, const u64* __restrict__ first, const u64* __restrict__ off,
                                                 const u32* __restrict__ doc, const u8* __restrict__ strand, u32 n, u32* __restrict__ line_len,
                                                 const u64* __restrict__ line_off, u8* __restrict__ text) {
    u32 lane = threadIdx.x & 31, r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    u32 L = len[r], dl = ndigits(L);
    u64 b = first[r];
    u32 k = (u32)(first[r + 1] - b);
    u32 wp = 0, wd = 0;
    for (u32 i = lane; i < k; i += 32) { wp += ndigits(off[b + i]); wd += ndigits(doc[b + i]); }
    wp = warp_sum(wp); wd = warp_sum(wd);
    u32 total = dl + 1 + wp + (k - 1) + 1 + wd + (k - 1) + 1 + k + (k - 1) + 1;
    if (!WRITE) { if (lane == 0) line_len[r] = total; return; }
    u8* dst = text + line_off[r];
    u32 at_p = dl + 1, at_d = at_p + wp + (k - 1) + 1, at_s = at_d + wd + (k - 1) + 1;
    if (lane == 0) { put_number(dst, L, dl); dst[dl] = '\t'; dst[at_d - 1] = '\t'; dst[at_s - 1] = '\t'; dst[total - 1] = '\n'; }
    for (u32 base = 0; base < k; base += 32) {
        u32 i = base + lane;
        u64 p = (i < k) ? off[b + i] : 0; u32 d = (i < k) ? doc[b + i] : 0;
        u32 w1 = (i < k) ? ndigits(p) : 0, w2 = (i < k) ? ndigits(d) : 0;
        u32 i1 = warp_inclusive_sum(w1, lane), i2 = warp_inclusive_sum(w2, lane);
        if (i < k) {
            u32 pp = at_p + (i1 - w1) + i, pd = at_d + (i2 - w2) + i, ps = at_s + 2 * i;
            put_number(dst + pp, p, w1); put_number(dst + pd, d, w2); dst[ps] = strand[b + i];
            if (i + 1 < k) { dst[pp + w1] = ','; dst[pd + w2] = ','; dst[ps + 1] = ','; }
        }
        at_p += __shfl_sync(FULL, i1, 31); at_d += __shfl_sync(FULL, i2, 31);
    }
}

__global__ void pack_strand_bits(const u8* __restrict__ strand, u64 nbits, u8* __restrict__ packed) {
    u64 byte = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (byte * 8 >= nbits) return;
    u32 v = 0;
    for (u32 b = 0; b < 8; ++b) { u64 i = byte * 8 + b; if (i < nbits && strand[i] == '+') v |= 1u << (7 - b); }
    packed[byte] = (u8)v;
}

__global__ void rebase_first(const u64* __restrict__ in, u64 n, u64 add, u64* __restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + add;
}

u32 warps_grid(u32 n) { return ceil_div((u64)n * 32, TPB); }

// --trim (optional, off by default): mumemto/trim_mums.py:51-86 of the reference.  A '+' multi-MUM may run through the '$' that ends its
// document (quirk Q1); the reference's post-processing shortens such a row by its largest overhang max_d(start_d + len - L_d) -- the
// start coordinates stay as they are -- and drops every row that is then shorter than min_length.  One warp per row.
__global__ void __launch_bounds__(TPB) trim_lengths(const u32* __restrict__ len, const i64* __restrict__ off, u32 n, u32 N, const u64* __restrict__ doc_len,
                                                    u32 min_length, u32* __restrict__ new_len, u32* __restrict__ keep) {
    u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= n) return;
    i64 l = len[r], worst = INT64_MIN;
    for (u32 d = lane; d < N; d += 32) { i64 o = off[(u64)r * N + d]; if (o != -1) worst = max(worst, o + l - (i64)doc_len[d]); }
#pragma unroll
    for (u32 k = 16; k > 0; k >>= 1) worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, k));
    if (lane == 0) {
        i64 nl = worst > 0 ? l - worst : l;
        bool ok = nl >= (i64)min_length;
        new_len[r] = ok ? (u32)nl : 0u; keep[r] = ok ? 1u : 0u;
    }
}
__global__ void __launch_bounds__(TPB) trim_gather(const u32* __restrict__ new_len, const i64* __restrict__ off, const u8* __restrict__ strand,
                                                   const u32* __restrict__ keep, const u32* __restrict__ slot, u32 n, u32 N, u32* __restrict__ len_o,
                                                   i64* __restrict__ off_o, u8* __restrict__ strand_o) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (u64)n * N) return;
    u32 r = (u32)(t / N), d = (u32)(t % N);
    if (!keep[r]) return;
    u64 dst = (u64)slot[r] * N + d;
    off_o[dst] = off[t]; strand_o[dst] = strand[t];
    if (d == 0) len_o[slot[r]] = new_len[r];
}

}  // namespace

void trim_mum_table(MumTable& t, u32 N, const u64* d_doc_len, u32 min_length, cudaStream_t s) {
    u32 n = (u32)t.count;
    if (!n) return;
    DevBuf<u32> nl, keep, slot; nl.alloc(n, s); keep.alloc((u64)n + 1, s); slot.alloc((u64)n + 1, s);
    MMB_CUDA(cudaMemsetAsync(keep.p + n, 0, 4, s));
    MMB_LAUNCH(trim_lengths, warps_grid(n), TPB, 0, s, (const u32*)t.len.p, (const i64*)t.off.p, n, N, d_doc_len, min_length, nl.p, keep.p);
    prim::exclusive_sum(keep.p, slot.p, (u64)n + 1, s);
    u32 kept = to_host(slot.p + n, 1, s)[0];
    DevBuf<u32> len_o; DevBuf<i64> off_o; DevBuf<u8> strand_o;
    len_o.alloc(kept, s); off_o.alloc((u64)kept * N, s); strand_o.alloc((u64)kept * N, s);
    if (kept) MMB_LAUNCH(trim_gather, ceil_div((u64)n * N, TPB), TPB, 0, s, (const u32*)nl.p, (const i64*)t.off.p, (const u8*)t.strand.p, (const u32*)keep.p,
                         (const u32*)slot.p, n, N, len_o.p, off_o.p, strand_o.p);
    t.len = std::move(len_o); t.off = std::move(off_o); t.strand = std::move(strand_o); t.count = kept;
}

template <class SAT>
void mum_rows(const Reports& rep, const SAT* sa, const DetectParams& prm, const u64* d_doc_start, const u64* d_doc_len, MumTable& out, cudaStream_t s) {
    u32 n = (u32)rep.count, N = prm.ndocs;
    out.count = 0; out.text_bytes = 0; out.text.alloc(0, s);
    if (n == 0) { out.len.alloc(0, s); out.off.alloc(0, s); out.strand.alloc(0, s); return; }
    DevBuf<u32> valid, idx; valid.alloc(n, s); idx.alloc(n, s);
    auto mum_validate_k = mum_validate<SAT>;
    MMB_LAUNCH(mum_validate_k, warps_grid(n), TPB, 0, s, rep.rep.p, n, sa, d_doc_start, d_doc_len, N, prm.revcomp, valid.p);
    prim::exclusive_sum(valid.p, idx.p, n, s);
    u32 tail[2];
    MMB_CUDA(cudaMemcpyAsync(&tail[0], idx.p + n - 1, 4, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaMemcpyAsync(&tail[1], valid.p + n - 1, 4, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaStreamSynchronize(s));
    u32 nv = tail[0] + tail[1];
    out.count = nv;
    out.len.alloc(nv, s); out.off.alloc((u64)nv * N, s); out.strand.alloc((u64)nv * N, s);
    if (nv == 0) return;
    DevBuf<Report> kept; kept.alloc(nv, s);
    MMB_LAUNCH(compact_reports, ceil_div(n, 256), 256, 0, s, rep.rep.p, valid.p, idx.p, n, kept.p);
    MMB_CUDA(cudaMemsetAsync(out.off.p, 0xff, out.off.bytes(), s));
    MMB_CUDA(cudaMemsetAsync(out.strand.p, 0, out.strand.bytes(), s));
    auto mum_rows_k = mum_rows_kernel<SAT>;
    MMB_LAUNCH(mum_rows_k, warps_grid(nv), TPB, 0, s, kept.p, nv, sa, d_doc_start, d_doc_len, N, prm.revcomp, out.len.p, out.off.p, out.strand.p);
}

void mum_text(MumTable& out, u32 N, cudaStream_t s) {
    u32 nv = (u32)out.count;
    out.text_bytes = 0;
    if (nv == 0) { out.text.alloc(0, s); return; }
    DevBuf<u32> ll; DevBuf<u64> lo; ll.alloc(nv + 1, s); lo.alloc(nv + 1, s);
    MMB_CUDA(cudaMemsetAsync(ll.p + nv, 0, 4, s));
    auto mum_lines_len_k = mum_lines<false>; auto mum_lines_write_k = mum_lines<true>;
    MMB_LAUNCH(mum_lines_len_k, warps_grid(nv), TPB, 0, s, out.len.p, out.off.p, out.strand.p, nv, N, ll.p, (const u64*)nullptr, (u8*)nullptr);
    prim::exclusive_sum(ll.p, lo.p, nv + 1, s);
    out.text_bytes = to_host(lo.p + nv, 1, s)[0];
    out.text.alloc(out.text_bytes, s);
    MMB_LAUNCH(mum_lines_write_k, warps_grid(nv), TPB, 0, s, out.len.p, out.off.p, out.strand.p, nv, N, ll.p, lo.p, out.text.p);
}

template <class SAT>
void emit_mums(const Reports& rep, const SAT* sa, const DetectParams& prm, const u64* d_doc_start, const u64* d_doc_len, bool want_text,
               MumTable& out, cudaStream_t s) {
    mum_rows<SAT>(rep, sa, prm, d_doc_start, d_doc_len, out, s);
    if (want_text) mum_text(out, prm.ndocs, s);
}

template <class SAT>
void mem_rows(const Reports& rep, const SAT* sa, const DetectParams& prm, const u64* d_doc_start, const u64* d_doc_len, MemTable& out, cudaStream_t s) {
    u32 n = (u32)rep.count, N = prm.ndocs;
    out.count = n; out.total_occ = 0; out.text_bytes = 0; out.text.alloc(0, s);
    out.len.alloc(n, s); out.first.alloc((u64)n + 1, s);
    if (n == 0) { MMB_CUDA(cudaMemsetAsync(out.first.p, 0, 8, s)); out.off.alloc(0, s); out.doc.alloc(0, s); out.strand.alloc(0, s); return; }
    DevBuf<u64> occ; occ.alloc((u64)n + 1, s);
    MMB_LAUNCH(mem_sizes, ceil_div(n + 1, 256), 256, 0, s, rep.rep.p, n, occ.p);
    prim::exclusive_sum(occ.p, out.first.p, (u64)n + 1, s);
    out.total_occ = to_host(out.first.p + n, 1, s)[0];
    out.off.alloc(out.total_occ, s); out.doc.alloc(out.total_occ, s); out.strand.alloc(out.total_occ, s);
    auto mem_rows_k = mem_rows_kernel<SAT>;
    MMB_LAUNCH(mem_rows_k, warps_grid(n), TPB, 0, s, rep.rep.p, n, sa, d_doc_start, d_doc_len, N, prm.revcomp, out.first.p, out.len.p, out.off.p, out.doc.p,
               out.strand.p);
}

void mem_text(MemTable& out, cudaStream_t s) {
    u32 n = (u32)out.count;
    out.text_bytes = 0;
    if (n == 0) { out.text.alloc(0, s); return; }
    DevBuf<u32> ll; DevBuf<u64> lo; ll.alloc(n + 1, s); lo.alloc(n + 1, s);
    MMB_CUDA(cudaMemsetAsync(ll.p + n, 0, 4, s));
    auto mem_lines_len_k = mem_lines<false>; auto mem_lines_write_k = mem_lines<true>;
    MMB_LAUNCH(mem_lines_len_k, warps_grid(n), TPB, 0, s, out.len.p, out.first.p, out.off.p, out.doc.p, out.strand.p, n, ll.p, (const u64*)nullptr, (u8*)nullptr);
    prim::exclusive_sum(ll.p, lo.p, n + 1, s);
    out.text_bytes = to_host(lo.p + n, 1, s)[0];
    out.text.alloc(out.text_bytes, s);
    MMB_LAUNCH(mem_lines_write_k, warps_grid(n), TPB, 0, s, out.len.p, out.first.p, out.off.p, out.doc.p, out.strand.p, n, ll.p, lo.p, out.text.p);
}

template <class SAT>
void emit_mems(const Reports& rep, const SAT* sa, const DetectParams& prm, const u64* d_doc_start, const u64* d_doc_len, bool want_text,
               MemTable& out, cudaStream_t s) {
    mem_rows<SAT>(rep, sa, prm, d_doc_start, d_doc_len, out, s);
    if (want_text) mem_text(out, s);
}

void concat_mums(std::vector<MumTable>& parts, u32 N, MumTable& out, cudaStream_t s) {
    u64 total = 0;
    for (auto& p : parts) total += p.count;
    if (total > 0xfffffff0ull) throw Error(35, "too many matches");
    out.count = total; out.text_bytes = 0; out.text.alloc(0, s);
    out.len.alloc(total, s); out.off.alloc(total * N, s); out.strand.alloc(total * N, s);
    u64 at = 0;
    for (auto& p : parts) {
        if (!p.count) continue;
        MMB_CUDA(cudaMemcpyAsync(out.len.p + at, p.len.p, p.count * 4, cudaMemcpyDeviceToDevice, s));
        MMB_CUDA(cudaMemcpyAsync(out.off.p + at * N, p.off.p, p.count * N * 8, cudaMemcpyDeviceToDevice, s));
        MMB_CUDA(cudaMemcpyAsync(out.strand.p + at * N, p.strand.p, p.count * N, cudaMemcpyDeviceToDevice, s));
        at += p.count;
    }
    MMB_CUDA(cudaStreamSynchronize(s));
    parts.clear();
}

void concat_mems(std::vector<MemTable>& parts, MemTable& out, cudaStream_t s) {
    u64 total = 0, occ = 0;
    for (auto& p : parts) { total += p.count; occ += p.total_occ; }
    if (total > 0xfffffff0ull) throw Error(35, "too many matches");
    out.count = total; out.total_occ = occ; out.text_bytes = 0; out.text.alloc(0, s);
    out.len.alloc(total, s); out.first.alloc(total + 1, s); out.off.alloc(occ, s); out.doc.alloc(occ, s); out.strand.alloc(occ, s);
    u64 at = 0, oat = 0;
    for (auto& p : parts) {
        if (p.count) {
            MMB_CUDA(cudaMemcpyAsync(out.len.p + at, p.len.p, p.count * 4, cudaMemcpyDeviceToDevice, s));
            MMB_LAUNCH(rebase_first, ceil_div(p.count, 256), 256, 0, s, (const u64*)p.first.p, p.count, oat, out.first.p + at);
            if (p.total_occ) {
                MMB_CUDA(cudaMemcpyAsync(out.off.p + oat, p.off.p, p.total_occ * 8, cudaMemcpyDeviceToDevice, s));
                MMB_CUDA(cudaMemcpyAsync(out.doc.p + oat, p.doc.p, p.total_occ * 4, cudaMemcpyDeviceToDevice, s));
                MMB_CUDA(cudaMemcpyAsync(out.strand.p + oat, p.strand.p, p.total_occ, cudaMemcpyDeviceToDevice, s));
            }
        }
        at += p.count; oat += p.total_occ;
    }
    MMB_CUDA(cudaMemcpyAsync(out.first.p + total, &occ, 8, cudaMemcpyHostToDevice, s));
    MMB_CUDA(cudaStreamSynchronize(s));
    parts.clear();
}

void pack_strands(const u8* strand, u64 nbits, DevBuf<u8>& packed, cudaStream_t s) {
    u64 bytes = (nbits + 7) / 8;
    packed.alloc(bytes, s);
    if (bytes) MMB_LAUNCH(pack_strand_bits, ceil_div(bytes, 256), 256, 0, s, strand, nbits, packed.p);
}

template void mum_rows<u32>(const Reports&, const u32*, const DetectParams&, const u64*, const u64*, MumTable&, cudaStream_t);
template void mum_rows<u64>(const Reports&, const u64*, const DetectParams&, const u64*, const u64*, MumTable&, cudaStream_t);
template void mem_rows<u32>(const Reports&, const u32*, const DetectParams&, const u64*, const u64*, MemTable&, cudaStream_t);
template void mem_rows<u64>(const Reports&, const u64*, const DetectParams&, const u64*, const u64*, MemTable&, cudaStream_t);
template void emit_mums<u32>(const Reports&, const u32*, const DetectParams&, const u64*, const u64*, bool, MumTable&, cudaStream_t);
template void emit_mums<u64>(const Reports&, const u64*, const DetectParams&, const u64*, const u64*, bool, MumTable&, cudaStream_t);
template void emit_mems<u32>(const Reports&, const u32*, const DetectParams&, const u64*, const u64*, bool, MemTable&, cudaStream_t);
template void emit_mems<u64>(const Reports&, const u64*, const DetectParams&, const u64*, const u64*, bool, MemTable&, cudaStream_t);

}  // namespace mmb
