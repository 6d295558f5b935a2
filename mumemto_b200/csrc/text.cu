// mumemto_b200/csrc/text.cu -- FASTA bytes (device resident) -> concatenated text T = fwd_0 $ rc_0 $ fwd_1 $ rc_1 $ ...
//
// Replaces RefBuilder::build_input_file (reference src/ref_builder.cpp:211-314): kseq record rules (include/kseq.h:173-219),
// per-record toupper (:234-236), records of one file joined without separator, '$' after the forward text, then the reverse
// complement (seqtk table, :29-50) and a second '$' (:272-292).  Row A of SURVEY.md section 8.
//
// The parse is a 4-state line automaton evaluated in parallel: every 4 KiB chunk summarises itself as a map state->state, the maps
// are composed per file, and with the start state of each chunk known every byte decides locally whether it is sequence.
//   pass 1  fasta_chunk_maps     read raw                      -> 1 B map per chunk
//   (tiny)  fasta_chunk_states   one thread per file composes  -> start state per chunk
//   pass 2  fasta_chunk_counts   read raw                      -> kept bases / header starts per chunk
//   (scan)  exclusive sums over chunks; fasta_doc_layout       -> doc_start[], doc_len[]
//   pass 3  fasta_scatter        read raw, write T (fwd + rc)  -> 1 B read + 2 B written per base
// Algorithmic traffic: 3 reads of the FASTA bytes + 2 bytes written per base.
#include "text.h"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 CHUNK = 4096;     // bytes per block
constexpr u32 TPB = 256;        // threads per block, 16 bytes each
constexpr u32 FULL = 0xffffffffu;

enum : u32 { ST_PRE = 0, ST_BOL = 1, ST_HDR = 2, ST_SEQ = 3 };
constexpr u32 MAP_ID = 0xE4;    // 2 bits per input state: 3->3, 2->2, 1->1, 0->0

__device__ __forceinline__ bool is_marker(u8 c) { return c == '>' || c == '@'; }

__device__ __forceinline__ u32 map_compose(u32 a, u32 b) {   // apply a, then b
    u32 r = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) r |= ((b >> (2 * ((a >> (2 * s)) & 3))) & 3) << (2 * s);
    return r;
}

__device__ __forceinline__ u8 upper(u8 c) { return (c >= 'a' && c <= 'z') ? (u8)(c - 32) : c; }

// seqtk complement table restricted to what can follow toupper(); src/ref_builder.cpp:29-38 (note '`' -> '@' in that table).
__device__ __forceinline__ u8 complement_upper(u8 c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'U': return 'A';
        case 'M': return 'K'; case 'K': return 'M'; case 'R': return 'Y'; case 'Y': return 'R';
        case 'V': return 'B'; case 'B': return 'V'; case 'H': return 'D'; case 'D': return 'H';
        case '`': return '@';
        default: return c;
    }
}

struct ChunkView {
    u64 begin;      // absolute offset in raw of this thread's first byte
    u64 fend;       // absolute end of the file
    u32 nbytes;     // 0..16 valid bytes for this thread
    u8 b[16];
    u32 nl, mk, cr, pl;     // 16-bit masks over b[]: '\n', header marker ('>' '@'), '\r', '+'
    bool eof_or_nl_after;   // the byte after b[nbytes-1] is '\n' or the file ends there
};

// chunk index -> (file, chunk start).  chunk_first[f] = index of the first chunk of file f (N+1 entries).
__device__ __forceinline__ u32 find_file(const u32* __restrict__ chunk_first, u32 nfiles, u32 chunk) {
    u32 lo = 0, hi = nfiles;              // invariant: chunk_first[lo] <= chunk < chunk_first[hi]
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (chunk_first[mid] <= chunk) lo = mid; else hi = mid; }
    return lo;
}

__device__ __forceinline__ void load_view(const u8* __restrict__ raw, const u64* __restrict__ file_off, const u64* __restrict__ file_end,
                                          const u32* __restrict__ chunk_first, u32 nfiles, ChunkView& v, u32& file) {
    __shared__ u32 s_file;
    u32 chunk = blockIdx.x;
    if (threadIdx.x == 0) s_file = find_file(chunk_first, nfiles, chunk);
    __syncthreads();
    file = s_file;
    u64 fbeg = file_off[file];
    v.fend = file_end[file];
    v.begin = fbeg + (u64)(chunk - chunk_first[file]) * CHUNK + (u64)threadIdx.x * 16;
    v.nbytes = v.begin >= v.fend ? 0u : (u32)min((u64)16, v.fend - v.begin);
    if (v.nbytes == 16) {                                   // file slots are 16-byte aligned (engine.cu), so this is a 128-bit load
        uint4 w = *reinterpret_cast<const uint4*>(raw + v.begin);
        u32 ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) v.b[k] = (u8)(ww[k >> 2] >> (8 * (k & 3)));
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) v.b[k] = (k < (int)v.nbytes) ? raw[v.begin + k] : (u8)0;
    }
    v.nl = v.mk = v.cr = v.pl = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        u8 c = v.b[k];
        v.nl |= (u32)(c == '\n') << k; v.mk |= (u32)is_marker(c) << k; v.cr |= (u32)(c == '\r') << k; v.pl |= (u32)(c == '+') << k;
    }
    u32 valid = (1u << v.nbytes) - 1;
    v.nl &= valid; v.mk &= valid; v.cr &= valid; v.pl &= valid;
    v.eof_or_nl_after = true;
    if (v.nbytes && ((v.cr >> (v.nbytes - 1)) & 1) && v.begin + v.nbytes < v.fend) v.eof_or_nl_after = raw[v.begin + v.nbytes] == '\n';
}

// Exclusive composition scan of per-thread maps over the block; *total = map of the whole block.
__device__ __forceinline__ u32 block_map_scan(u32 mine, u32* total) {
    __shared__ u32 warp_maps[TPB / 32];
    u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = mine;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc = map_compose(o, inc);
    }
    if (lane == 31) warp_maps[warp] = inc;
    __syncthreads();
    u32 before = MAP_ID, all = MAP_ID;
    for (u32 w = 0; w < TPB / 32; ++w) { if (w < warp) before = map_compose(before, warp_maps[w]); all = map_compose(all, warp_maps[w]); }
    u32 excl = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) excl = MAP_ID;
    *total = all;
    __syncthreads();            // warp_maps is reused by the next call
    return map_compose(before, excl);
}

// state after the thread's bytes for each of the four start states, from the masks alone
__device__ __forceinline__ u32 thread_map(const ChunkView& v) {
    if (v.nbytes == 0) return MAP_ID;
    bool has_nl = v.nl != 0;
    u32 tail = ST_BOL;                                      // state after the last newline of these bytes
    if (has_nl) {
        u32 last = 31 - (u32)__clz((int)v.nl);
        if (last + 1 < v.nbytes) tail = ((v.mk >> (last + 1)) & 1) ? ST_HDR : ST_SEQ;
    }
    u32 from_bol = has_nl ? tail : ((v.mk & 1) ? ST_HDR : ST_SEQ);
    u32 from_hdr = has_nl ? tail : ST_HDR;
    u32 from_seq = has_nl ? tail : ST_SEQ;
    u32 from_pre = ST_PRE;
    if (v.mk) {
        u32 fm = (u32)__ffs((int)v.mk) - 1;
        from_pre = (v.nl >> fm) >> 1 ? tail : ST_HDR;        // a newline after the first marker ends that header
    }
    return (from_pre << (2 * ST_PRE)) | (from_bol << (2 * ST_BOL)) | (from_hdr << (2 * ST_HDR)) | (from_seq << (2 * ST_SEQ));
}

// Which of the thread's bytes are sequence symbols (keep) and which start a header, given the state before its first byte.
struct ByteMasks { u32 keep, header, fastq; };
__device__ __forceinline__ ByteMasks classify_bytes(const ChunkView& v, u32 st) {
    ByteMasks r; r.keep = r.header = r.fastq = 0;
    if (v.nbytes == 0) return r;
    u32 valid = (1u << v.nbytes) - 1;
    u32 cut = 0;                                            // bytes below this index are still before the first header of the file
    u32 pre_marker = 0;
    if (st == ST_PRE) {
        if (!v.mk) return r;
        cut = (u32)__ffs((int)v.mk) - 1; pre_marker = 1u << cut;
    }
    u32 from_cut = valid & ~((1u << cut) - 1);
    u32 line_start = ((v.nl << 1) | (st == ST_BOL ? 1u : 0u)) & valid;
    r.header = ((v.mk & line_start) | pre_marker) & from_cut;
    u32 run = ~v.nl & valid;                                // header text runs up to (not including) the next newline
    u32 start = r.header | ((st == ST_HDR ? 1u : 0u) & run);
    u32 in_header = ((start + run) ^ run) & run;            // carry trick: fill from each start bit to the end of its run of non-newlines
    u32 eol_next = (v.nl >> 1) | (v.eof_or_nl_after ? 1u << (v.nbytes - 1) : 0u);
    r.keep = from_cut & ~v.nl & ~in_header & ~(v.cr & eol_next);
    r.fastq = v.pl & line_start & from_cut & ~in_header;
    return r;
}

// Exclusive sum of one u32 per thread over the block; *total = block sum.
__device__ __forceinline__ u32 block_sum_scan(u32 mine, u32* total) {
    __shared__ u32 warp_sums[TPB / 32];
    u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = mine;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    u32 before = 0, all = 0;
    for (u32 w = 0; w < TPB / 32; ++w) { if (w < warp) before += warp_sums[w]; all += warp_sums[w]; }
    *total = all;
    __syncthreads();
    return before + inc - mine;
}

__global__ void __launch_bounds__(TPB) fasta_chunk_maps(const u8* __restrict__ raw, const u64* __restrict__ file_off, const u64* __restrict__ file_end,
                                                        const u32* __restrict__ chunk_first, u32 nfiles, u8* __restrict__ chunk_map) {
    ChunkView v; u32 file;
    load_view(raw, file_off, file_end, chunk_first, nfiles, v, file);
    u32 total;
    block_map_scan(thread_map(v), &total);
    if (threadIdx.x == 0) chunk_map[blockIdx.x] = (u8)total;
}

__global__ void fasta_chunk_states(const u8* __restrict__ chunk_map, const u32* __restrict__ chunk_first, u32 nfiles, u8* __restrict__ chunk_state) {
    u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfiles) return;
    u32 st = ST_PRE;
    for (u32 c = chunk_first[f]; c < chunk_first[f + 1]; ++c) { chunk_state[c] = (u8)st; st = (chunk_map[c] >> (2 * st)) & 3; }
}

// Write `total` staged symbols to text[dst0 ..): bytes up to the first 4-byte boundary, aligned words, trailing bytes.
// REVERSE: the run is the reverse complement of the staged symbols.
template <bool REVERSE>
__device__ __forceinline__ void write_run(u8* __restrict__ text, u64 dst0, const u8* stage, u32 total) {
    auto get = [&](u32 j) -> u32 { return REVERSE ? (u32)complement_upper(stage[total - 1 - j]) : (u32)stage[j]; };
    u32 head = min(total, (u32)((4 - (dst0 & 3)) & 3));
    for (u32 j = threadIdx.x; j < head; j += blockDim.x) text[dst0 + j] = (u8)get(j);
    u32 nwords = (total - head) >> 2;
    for (u32 w = threadIdx.x; w < nwords; w += blockDim.x) {
        u32 j = head + 4 * w;
        *reinterpret_cast<u32*>(text + dst0 + j) = get(j) | (get(j + 1) << 8) | (get(j + 2) << 16) | (get(j + 3) << 24);
    }
    for (u32 j = head + 4 * nwords + threadIdx.x; j < total; j += blockDim.x) text[dst0 + j] = (u8)get(j);
}

template <bool SCATTER>
__global__ void __launch_bounds__(TPB) fasta_pass(const u8* __restrict__ raw, const u64* __restrict__ file_off, const u64* __restrict__ file_end,
                                                  const u32* __restrict__ chunk_first, u32 nfiles, const u8* __restrict__ chunk_state, int raw_mode,
                                                  u32* __restrict__ chunk_kept, u32* __restrict__ chunk_hdr, u32* __restrict__ fastq_flag,
                                                  // scatter-only arguments
                                                  const u64* __restrict__ chunk_kept_base, const u32* __restrict__ chunk_hdr_base,
                                                  const u64* __restrict__ doc_start, const u64* __restrict__ doc_len, int revcomp,
                                                  u8* __restrict__ text, u64* __restrict__ rec_pos, u64* __restrict__ rec_before) {
    ChunkView v; u32 file;
    load_view(raw, file_off, file_end, chunk_first, nfiles, v, file);
    u32 st;
    if (raw_mode) st = ST_SEQ;
    else {
        u32 total;
        u32 prefix = block_map_scan(thread_map(v), &total);
        st = (prefix >> (2 * (u32)chunk_state[blockIdx.x])) & 3;
    }
    ByteMasks bm;
    if (raw_mode) { bm.keep = (1u << v.nbytes) - 1; bm.header = 0; bm.fastq = 0; }
    else bm = classify_bytes(v, st);
    u32 kept = (u32)__popc(bm.keep), hdr = (u32)__popc(bm.header);
    if (bm.fastq) *fastq_flag = 1;
    u32 kept_total, hdr_total;
    u32 kept_before = block_sum_scan(kept, &kept_total);
    u32 hdr_before = block_sum_scan(hdr, &hdr_total);
    if (!SCATTER) {
        if (threadIdx.x == 0) { chunk_kept[blockIdx.x] = kept_total; chunk_hdr[blockIdx.x] = hdr_total; }
        return;
    }
    // kept symbols go to shared memory first (a chunk's kept symbols form ONE contiguous run in the forward text and one in the
    // reverse complement), then the block writes both runs with aligned 4-byte stores
    __shared__ u8 stage[CHUNK];
    u64 kbase = chunk_kept_base[blockIdx.x] + kept_before;      // global index among kept bases of this thread's first kept byte
    u32 hidx = chunk_hdr_base[blockIdx.x] + hdr_before;
    u32 local = kept_before;
#pragma unroll
    for (u32 k = 0; k < 16; ++k) {
        if ((bm.header >> k) & 1) { rec_pos[hidx] = v.begin + k; rec_before[hidx] = kbase + (u32)__popc(bm.keep & ((1u << k) - 1)); ++hidx; }
        if ((bm.keep >> k) & 1) stage[local++] = upper(v.b[k]);
    }
    __syncthreads();
    u64 base = doc_start[file], L = doc_len[file];
    u64 q0 = chunk_kept_base[blockIdx.x] - chunk_kept_base[chunk_first[file]];      // document-relative index of the chunk's first kept base
    u32 total = kept_total;
    if (total == 0) return;
    write_run<false>(text, base + q0, stage, total);
    if (revcomp) write_run<true>(text, base + 2 * L - (q0 + total - 1), stage, total);
}

__global__ void fasta_doc_layout(const u64* __restrict__ chunk_kept_base, const u32* __restrict__ chunk_first, u32 nfiles, u32 nchunks,
                                 u64 total_kept, int revcomp, u64* __restrict__ doc_start, u64* __restrict__ doc_len) {
    // one thread: N is small (files), and each doc start depends on all earlier lengths
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u64 at = 0;
    for (u32 f = 0; f < nfiles; ++f) {
        u64 b0 = chunk_first[f] < nchunks ? chunk_kept_base[chunk_first[f]] : total_kept;
        u64 b1 = chunk_first[f + 1] < nchunks ? chunk_kept_base[chunk_first[f + 1]] : total_kept;
        u64 L = b1 - b0;
        doc_start[f] = at; doc_len[f] = L;
        at += revcomp ? 2 * (L + 1) : (L + 1);
    }
    doc_start[nfiles] = at;
}

__global__ void fasta_terminators(const u64* __restrict__ doc_start, const u64* __restrict__ doc_len, u32 nfiles, int revcomp, u8* __restrict__ text) {
    u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfiles) return;
    u64 L = doc_len[f];
    text[doc_start[f] + L] = '$';
    if (revcomp) text[doc_start[f] + 2 * L + 1] = '$';
}

}  // namespace

void build_text(const u8* d_raw, const std::vector<u64>& file_off, const std::vector<u64>& file_size, bool revcomp, bool raw_mode,
                TextLayout& out, cudaStream_t s) {
    u32 nfiles = (u32)file_off.size();
    std::vector<u32> chunk_first(nfiles + 1, 0);
    std::vector<u64> file_end(nfiles);
    for (u32 f = 0; f < nfiles; ++f) {
        if (file_off[f] % 16) throw Error(11, "file image offsets must be multiples of 16");
        u64 nch = (file_size[f] + CHUNK - 1) / CHUNK;
        if ((u64)chunk_first[f] + nch > 0xfffffff0ull) throw Error(12, "input too large for one device");
        chunk_first[f + 1] = chunk_first[f] + (u32)nch;
        file_end[f] = file_off[f] + file_size[f];
    }
    u32 nchunks = chunk_first[nfiles];
    DevBuf<u64> d_off, d_end, d_kept_base; DevBuf<u32> d_first, d_kept, d_hdr, d_hdr_base, d_flag; DevBuf<u8> d_map, d_state;
    d_off.alloc(nfiles, s); d_end.alloc(nfiles, s); d_first.alloc(nfiles + 1, s);
    MMB_CUDA(cudaMemcpyAsync(d_off.p, file_off.data(), nfiles * 8, cudaMemcpyHostToDevice, s));
    MMB_CUDA(cudaMemcpyAsync(d_end.p, file_end.data(), nfiles * 8, cudaMemcpyHostToDevice, s));
    MMB_CUDA(cudaMemcpyAsync(d_first.p, chunk_first.data(), (nfiles + 1) * 4, cudaMemcpyHostToDevice, s));
    d_map.alloc(nchunks, s); d_state.alloc(nchunks, s); d_kept.alloc(nchunks + 1, s); d_hdr.alloc(nchunks + 1, s);
    d_kept_base.alloc(nchunks + 1, s); d_hdr_base.alloc(nchunks + 1, s); d_flag.alloc(1, s);
    d_flag.zero(); d_kept.zero(); d_hdr.zero();
    out.doc_start.alloc(nfiles + 1, s); out.doc_len.alloc(nfiles, s);

    if (nchunks) {
        if (!raw_mode) {
            MMB_LAUNCH(fasta_chunk_maps, nchunks, TPB, 0, s, d_raw, d_off.p, d_end.p, d_first.p, nfiles, d_map.p);
            MMB_LAUNCH(fasta_chunk_states, ceil_div(nfiles, 128), 128, 0, s, d_map.p, d_first.p, nfiles, d_state.p);
        }
        auto fasta_count_pass = fasta_pass<false>;
        MMB_LAUNCH(fasta_count_pass, nchunks, TPB, 0, s, d_raw, d_off.p, d_end.p, d_first.p, nfiles, d_state.p, (int)raw_mode, d_kept.p, d_hdr.p, d_flag.p,
                   (const u64*)nullptr, (const u32*)nullptr, (const u64*)nullptr, (const u64*)nullptr, (int)revcomp, (u8*)nullptr, (u64*)nullptr, (u64*)nullptr);
    }
    prim::exclusive_sum(d_kept.p, d_kept_base.p, nchunks + 1, s);     // entry [nchunks] = total
    prim::exclusive_sum(d_hdr.p, d_hdr_base.p, nchunks + 1, s);
    u64 total_kept = 0; u32 total_hdr = 0, fastq = 0;
    MMB_CUDA(cudaMemcpyAsync(&total_kept, d_kept_base.p + nchunks, 8, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaMemcpyAsync(&total_hdr, d_hdr_base.p + nchunks, 4, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaMemcpyAsync(&fastq, d_flag.p, 4, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaStreamSynchronize(s));
    if (fastq) throw Error(13, "FASTQ records ('+' separator line) are not supported");

    MMB_LAUNCH(fasta_doc_layout, 1, 1, 0, s, d_kept_base.p, d_first.p, nfiles, nchunks, total_kept, (int)revcomp, out.doc_start.p, out.doc_len.p);
    out.n = (revcomp ? 2 : 1) * (total_kept + nfiles);
    out.text.alloc(out.n + 64, s);     // slack so that vector loads near the end stay in bounds
    MMB_CUDA(cudaMemsetAsync(out.text.p + out.n, 0, 64, s));
    DevBuf<u64> d_rec_pos, d_rec_before;
    d_rec_pos.alloc(total_hdr, s); d_rec_before.alloc(total_hdr, s);
    MMB_LAUNCH(fasta_terminators, ceil_div(nfiles, 128), 128, 0, s, out.doc_start.p, out.doc_len.p, nfiles, (int)revcomp, out.text.p);
    if (nchunks) {
        auto fasta_scatter_pass = fasta_pass<true>;
        MMB_LAUNCH(fasta_scatter_pass, nchunks, TPB, 0, s, d_raw, d_off.p, d_end.p, d_first.p, nfiles, d_state.p, (int)raw_mode, d_kept.p, d_hdr.p, d_flag.p,
                   d_kept_base.p, d_hdr_base.p, out.doc_start.p, out.doc_len.p, (int)revcomp, out.text.p, d_rec_pos.p, d_rec_before.p);
    }
    out.h_doc_start = to_host(out.doc_start.p, nfiles + 1, s);
    out.h_doc_len = to_host(out.doc_len.p, nfiles, s);
    out.h_rec_pos = to_host(d_rec_pos.p, total_hdr, s);
    out.h_rec_before = to_host(d_rec_before.p, total_hdr, s);
    out.total_bases = total_kept;
    out.revcomp = revcomp;
}

}  // namespace mmb
