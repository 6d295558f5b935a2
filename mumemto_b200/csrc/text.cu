// mumemto_b200/csrc/text.cu -- FASTA bytes (device resident) -> concatenated text T = fwd_0 $ rc_0 $ fwd_1 $ rc_1 $ ...
//
// Replaces RefBuilder::build_input_file (reference src/ref_builder.cpp:211-314): kseq record rules (include/kseq.h:173-219),
// per-record toupper (:234-236), records of one file joined without separator, '$' after the forward text, then the reverse
// complement (seqtk table, :29-50) and a second '$' (:272-292).  Row A of SURVEY.md section 8.
//
// The parse is a 4-state line automaton evaluated in parallel: every 4 KiB chunk summarises itself as a map state->state, the maps
// are composed per file, and with the start state of each chunk known every byte decides locally whether it is sequence.
//   pass 1  fasta_chunk_maps     read raw                      -> per chunk: 1 B map + kept bases / header starts for EACH of the four
//                                                                 states the chunk can start in (they only differ before the chunk's first
//                                                                 newline, so most threads classify their bytes once)
//   (tiny)  fasta_chunk_states   one thread per file composes  -> start state, kept bases and header starts per chunk
//   (scan)  exclusive sums over chunks; fasta_doc_layout       -> doc_start[], doc_len[]
//   pass 2  fasta_scatter        read raw, write T (fwd + rc)  -> 1 B read + 2 B written per base
// Algorithmic traffic: 2 reads of the FASTA bytes + 2 bytes written per base.
#include "text.h"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 CHUNK = 4096;     // bytes per block
constexpr u32 TPB = 256;        // threads per block, 16 bytes each
constexpr u32 FULL = 0xffffffffu;

enum : u32 { ST_PRE = 0, ST_BOL = 1, ST_HDR = 2, ST_SEQ = 3 };
constexpr u32 MAP_ID = 0xE4;    // 2 bits per input state: 3->3, 2->2, 1->1, 0->0

__device__ __forceinline__ u32 squeeze4(u32 m) { return ((m & 0x01010101u) * 0x01020408u) >> 24; }     // one bit per 0xff byte, byte 0 -> bit 0

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
    // the four 16-bit masks, four bytes at a time: a SIMD byte compare per class and word, then the four 0xff/0x00 bytes are squeezed
    // into four bits by one multiply (bits 0, 8, 16, 24 land on bits 24..27 of the product; no two partial products collide)
    v.nl = v.mk = v.cr = v.pl = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        u32 x = ((u32)v.b[4 * i]) | ((u32)v.b[4 * i + 1] << 8) | ((u32)v.b[4 * i + 2] << 16) | ((u32)v.b[4 * i + 3] << 24);
        v.nl |= squeeze4(__vcmpeq4(x, 0x0a0a0a0au)) << (4 * i);
        v.mk |= squeeze4(__vcmpeq4(x, 0x3e3e3e3eu) | __vcmpeq4(x, 0x40404040u)) << (4 * i);
        v.cr |= squeeze4(__vcmpeq4(x, 0x0d0d0d0du)) << (4 * i);
        v.pl |= squeeze4(__vcmpeq4(x, 0x2b2b2b2bu)) << (4 * i);
    }
    u32 valid = (1u << v.nbytes) - 1;
    v.nl &= valid; v.mk &= valid; v.cr &= valid; v.pl &= valid;
    v.eof_or_nl_after = true;
    if (v.nbytes && ((v.cr >> (v.nbytes - 1)) & 1) && v.begin + v.nbytes < v.fend) v.eof_or_nl_after = raw[v.begin + v.nbytes] == '\n';
}

// Inside the scans a map lives in BYTE form (byte s = image of state s): composition is then ONE byte permute whose selector is the
// first map squeezed to nibbles (the 2-bit form needed ~30 shift/mask instructions per composition, half of pass 1's instructions).
__device__ __forceinline__ u32 map_expand(u32 m) { return (m & 3u) | ((m & 0xcu) << 6) | ((m & 0x30u) << 12) | ((m & 0xc0u) << 18); }
__device__ __forceinline__ u32 map_contract(u32 y) { return (y & 3u) | ((y >> 6) & 0xcu) | ((y >> 12) & 0x30u) | ((y >> 18) & 0xc0u); }
__device__ __forceinline__ u32 map_compose_b(u32 a, u32 b) {      // apply a, then b; all in byte form
    u32 x = (a | (a >> 4)) & 0x00ff00ffu;
    u32 sel = (x | (x >> 8)) & 0xffffu;
    return __byte_perm(b, 0u, sel);
}
constexpr u32 MAP_ID_B = 0x03020100u;

// Exclusive composition scan of per-thread maps over the block; *total = map of the whole block.  (2-bit form in and out.)
__device__ __forceinline__ u32 block_map_scan(u32 mine, u32* total) {
    __shared__ u32 warp_maps[TPB / 32];
    u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = map_expand(mine);
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc = map_compose_b(o, inc);
    }
    if (lane == 31) warp_maps[warp] = inc;
    __syncthreads();
    u32 before = MAP_ID_B, all = MAP_ID_B;
    for (u32 w = 0; w < TPB / 32; ++w) { u32 m = warp_maps[w]; if (w < warp) before = map_compose_b(before, m); all = map_compose_b(all, m); }
    u32 excl = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) excl = MAP_ID_B;
    *total = map_contract(all);
    __syncthreads();            // warp_maps is reused by the next call
    return map_contract(map_compose_b(before, excl));
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

// State before the thread's first byte when the chunk is known to start in BOL, HDR or SEQ (the scatter pass; every chunk but a file's
// first).  Outside PRE the automaton forgets everything at a newline, so no map composition is needed: the state is the `tail` of the
// nearest earlier thread that holds a newline -- with "beginning of line" resolved by the first byte that follows -- or the chunk's
// start state if there is none.  One ballot and two shuffles per thread, one short walk over the warps before its own (the generic
// composition scan of 4-state maps was a quarter of the scatter pass's instructions).
__device__ __forceinline__ u32 block_state_before(const ChunkView& v, u32 cs) {
    __shared__ u8 w_has_nl[TPB / 32], w_end[TPB / 32], w_first_mk[TPB / 32], w_nonempty[TPB / 32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool has_nl = v.nl != 0;
    u32 tail = ST_BOL;
    if (has_nl) {
        u32 last = 31 - (u32)__clz((int)v.nl);
        if (last + 1 < v.nbytes) tail = ((v.mk >> (last + 1)) & 1) ? ST_HDR : ST_SEQ;
    }
    const u32 first_mk = v.mk & 1, nonempty = v.nbytes ? 1u : 0u;
    const u32 nlmask = __ballot_sync(FULL, has_nl);
    // state after the warp's last byte, when the warp holds a newline
    const u32 jl = nlmask ? 31 - (u32)__clz((int)nlmask) : 0;
    u32 t_last = __shfl_sync(FULL, tail, jl);
    const u32 nxt = min(jl + 1, 31u);
    const u32 mk_next = __shfl_sync(FULL, first_mk, nxt), ne_next = __shfl_sync(FULL, nonempty, nxt);
    if (t_last == ST_BOL && jl < 31 && ne_next) t_last = mk_next ? ST_HDR : ST_SEQ;
    if (lane == 0) { w_has_nl[warp] = nlmask ? 1 : 0; w_end[warp] = (u8)t_last; w_first_mk[warp] = (u8)first_mk; w_nonempty[warp] = (u8)nonempty; }
    __syncthreads();
    u32 sw = cs;                                                // state before the warp's first byte
    for (u32 w = 0; w < warp; ++w) {
        if (w_has_nl[w]) sw = w_end[w];
        else if (sw == ST_BOL && w_nonempty[w]) sw = w_first_mk[w] ? ST_HDR : ST_SEQ;
    }
    const u32 mk0 = __shfl_sync(FULL, first_mk, 0);
    const u32 prev = nlmask & ((1u << lane) - 1);
    const u32 j = prev ? 31 - (u32)__clz((int)prev) : 0;
    const u32 t = __shfl_sync(FULL, tail, j);
    const u32 mk_j1 = __shfl_sync(FULL, first_mk, min(j + 1, 31u));
    u32 st;
    if (prev) st = (t != ST_BOL) ? t : (j + 1 == lane ? (u32)ST_BOL : (mk_j1 ? ST_HDR : ST_SEQ));
    else st = (sw != ST_BOL) ? sw : (lane == 0 ? (u32)ST_BOL : (mk0 ? ST_HDR : ST_SEQ));
    __syncthreads();            // the shared arrays are reused by the next call
    return st;
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

// Sum of one u64 per thread over the block (every thread gets the total).
__device__ __forceinline__ u64 block_sum64(u64 mine) {
    __shared__ u64 warp_sums[TPB / 32];
    u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (u32 d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(FULL, mine, d);
    if (lane == 0) warp_sums[warp] = mine;
    __syncthreads();
    u64 all = 0;
    for (u32 w = 0; w < TPB / 32; ++w) all += warp_sums[w];
    __syncthreads();
    return all;
}

// Pass 1: the chunk's state map, and its kept-base / header-start counts for each of the four states it may start in, packed as
// four 16-bit fields (a chunk has at most 4096 bytes).
__global__ void __launch_bounds__(TPB) fasta_chunk_maps(const u8* __restrict__ raw, const u64* __restrict__ file_off, const u64* __restrict__ file_end,
                                                        const u32* __restrict__ chunk_first, u32 nfiles, int raw_mode, u8* __restrict__ chunk_map,
                                                        u64* __restrict__ chunk_kept4, u64* __restrict__ chunk_hdr4) {
    ChunkView v; u32 file;
    load_view(raw, file_off, file_end, chunk_first, nfiles, v, file);
    u64 kept4 = 0, hdr4 = 0; u32 total = MAP_ID;
    if (raw_mode) kept4 = (u64)v.nbytes * 0x0001000100010001ull;
    else {
        u32 prefix = block_map_scan(thread_map(v), &total);     // state before this thread's bytes, per chunk start state
        u32 done = 0xffffffffu; u32 k_prev = 0, h_prev = 0;
#pragma unroll
        for (u32 cs = 0; cs < 4; ++cs) {
            u32 st = (prefix >> (2 * cs)) & 3;
            if (st != done) { ByteMasks bm = classify_bytes(v, st); k_prev = (u32)__popc(bm.keep); h_prev = (u32)__popc(bm.header); done = st; }
            kept4 |= (u64)k_prev << (16 * cs); hdr4 |= (u64)h_prev << (16 * cs);
        }
    }
    kept4 = block_sum64(kept4); hdr4 = block_sum64(hdr4);
    if (threadIdx.x == 0) { chunk_map[blockIdx.x] = (u8)total; chunk_kept4[blockIdx.x] = kept4; chunk_hdr4[blockIdx.x] = hdr4; }
}

// One WARP per file: every lane composes the maps of its slice of the file's chunks, a warp scan of the 32 slice maps gives each lane the
// state its slice starts in, and a second walk over the slice writes the per-chunk start state and counts.
__global__ void __launch_bounds__(128) fasta_chunk_states(const u8* __restrict__ chunk_map, const u64* __restrict__ chunk_kept4, const u64* __restrict__ chunk_hdr4,
                                                         const u32* __restrict__ chunk_first, u32 nfiles, int raw_mode, u8* __restrict__ chunk_state,
                                                         u32* __restrict__ chunk_kept, u32* __restrict__ chunk_hdr) {
    u32 f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= nfiles) return;                                    // whole warp
    u32 c0 = chunk_first[f], c1 = chunk_first[f + 1];
    u32 per = (c1 - c0 + 31) / 32;
    u32 lo = min(c0 + lane * per, c1), hi = min(lo + per, c1);
    u32 mine = MAP_ID;
    if (!raw_mode) for (u32 c = lo; c < hi; ++c) mine = map_compose(mine, (u32)chunk_map[c]);
    u32 inc = mine;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc = map_compose(o, inc);
    }
    u32 before = __shfl_up_sync(FULL, inc, 1);                  // composition of the slices before mine
    if (lane == 0) before = MAP_ID;
    u32 st = raw_mode ? ST_SEQ : ((before >> (2 * ST_PRE)) & 3);        // every file starts before its first header
    for (u32 c = lo; c < hi; ++c) {
        chunk_state[c] = (u8)st;
        chunk_kept[c] = (u32)(chunk_kept4[c] >> (16 * st)) & 0xffffu;
        chunk_hdr[c] = (u32)(chunk_hdr4[c] >> (16 * st)) & 0xffffu;
        if (!raw_mode) st = (chunk_map[c] >> (2 * st)) & 3;
    }
}

// Copy `total` staged bytes to text[dst ..): src is shared memory staged at the same offset modulo 16 as dst, so after the bytes up to
// the first 16-byte boundary the body moves as aligned 128-bit words.
__device__ __forceinline__ void copy_run(u8* __restrict__ text, u64 dst, const u8* src, u32 total) {
    u32 head = min(total, (u32)((16 - (dst & 15)) & 15));
    for (u32 j = threadIdx.x; j < head; j += blockDim.x) text[dst + j] = src[j];
    u32 nvec = (total - head) >> 4;
    for (u32 w = threadIdx.x; w < nvec; w += blockDim.x)
        *reinterpret_cast<uint4*>(text + dst + head + 16 * (u64)w) = *reinterpret_cast<const uint4*>(src + head + 16 * w);
    for (u32 j = head + 16 * nvec + threadIdx.x; j < total; j += blockDim.x) text[dst + j] = src[j];
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
        const u32 cs = chunk_state[blockIdx.x];                // block-uniform
        if (cs == ST_PRE) {                                     // a file's first chunk: the general composition of state maps
            u32 total;
            u32 prefix = block_map_scan(thread_map(v), &total);
            st = (prefix >> (2 * ST_PRE)) & 3;
        } else st = block_state_before(v, cs);
    }
    ByteMasks bm;
    if (raw_mode) { bm.keep = (1u << v.nbytes) - 1; bm.header = 0; bm.fastq = 0; }
    else bm = classify_bytes(v, st);
    u32 kept = (u32)__popc(bm.keep), hdr = (u32)__popc(bm.header);
    if (bm.fastq) *fastq_flag = 1;
    u32 packed_total;                                           // one scan for both counts: each stays below 2^13 per block
    u32 packed_before = block_sum_scan(kept | (hdr << 16), &packed_total);
    u32 kept_total = packed_total & 0xffffu, hdr_total = packed_total >> 16;
    u32 kept_before = packed_before & 0xffffu, hdr_before = packed_before >> 16;
    if (!SCATTER) {
        if (threadIdx.x == 0) { chunk_kept[blockIdx.x] = kept_total; chunk_hdr[blockIdx.x] = hdr_total; }
        return;
    }
    // kept symbols go to shared memory first (a chunk's kept symbols form ONE contiguous run in the forward text and one in the
    // reverse complement): both runs are staged in final order -- upper-cased, the second one reversed and complemented through a
    // 256-byte table -- at the offset (mod 16) of their destination, then the block copies them with aligned 128-bit stores
    MMB_SHARED_ALIGNED16(u8, stage_f, CHUNK + 16);
    MMB_SHARED_ALIGNED16(u8, stage_r, CHUNK + 16);
    __shared__ u8 lut_u[256], lut_c[256];
    lut_u[threadIdx.x] = upper((u8)threadIdx.x); lut_c[threadIdx.x] = complement_upper(upper((u8)threadIdx.x));      // TPB = 256
    u64 kbase = chunk_kept_base[blockIdx.x] + kept_before;      // global index among kept bases of this thread's first kept byte
    u32 hidx = chunk_hdr_base[blockIdx.x] + hdr_before;
    u64 base = doc_start[file], L = doc_len[file];
    u64 q0 = chunk_kept_base[blockIdx.x] - chunk_kept_base[chunk_first[file]];      // document-relative index of the chunk's first kept base
    u32 total = kept_total;
    const u64 dst_f = base + q0, dst_r = base + 2 * L - (q0 + total - 1);
    const u32 of = (u32)(dst_f & 15), orr = (u32)(dst_r & 15);
    __syncthreads();                                            // the tables
    u32 local = kept_before;
#pragma unroll
    for (u32 k = 0; k < 16; ++k) {
        if ((bm.header >> k) & 1) { rec_pos[hidx] = v.begin + k; rec_before[hidx] = kbase + (u32)__popc(bm.keep & ((1u << k) - 1)); ++hidx; }
        if ((bm.keep >> k) & 1) {
            stage_f[of + local] = lut_u[v.b[k]];
            if (revcomp) stage_r[orr + (total - 1 - local)] = lut_c[v.b[k]];
            ++local;
        }
    }
    __syncthreads();
    if (total == 0) return;
    copy_run(text, dst_f, stage_f + of, total);
    if (revcomp) copy_run(text, dst_r, stage_r + orr, total);
}

__global__ void fasta_doc_layout(const u64* __restrict__ chunk_kept_base, const u32* __restrict__ chunk_first, u32 nfiles, u32 nchunks,
                                 u64 total_kept, int revcomp, u64 at0, u64* __restrict__ doc_start, u64* __restrict__ doc_len) {
    // one thread: N is small (files), and each doc start depends on all earlier lengths
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u64 at = at0;
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

void text_begin(const std::vector<u64>& file_size, bool revcomp, TextLayout& out, cudaStream_t s) {
    u64 worst = 0;
    for (u64 b : file_size) worst += b + 1;
    worst *= revcomp ? 2 : 1;
    out.text.alloc(worst + 128, s);     // zero slack behind the text so that vector loads near the end stay in bounds (pfp.cu stages 16-byte words up to n + 80)
    out.n = 0; out.total_bases = 0; out.revcomp = revcomp;
    out.h_doc_start.assign(1, 0); out.h_doc_len.clear(); out.h_rec_pos.clear(); out.h_rec_before.clear();
}

void text_append(const u8* d_raw, const std::vector<u64>& all_off, const std::vector<u64>& all_size, size_t f0, size_t f1, bool raw_mode,
                 TextLayout& out, cudaStream_t s) {
    std::vector<u64> file_off(all_off.begin() + f0, all_off.begin() + f1), file_size(all_size.begin() + f0, all_size.begin() + f1);
    const bool revcomp = out.revcomp;
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
    DevBuf<u64> d_off, d_end, d_kept_base, d_doc_start, d_doc_len; DevBuf<u32> d_first, d_kept, d_hdr, d_hdr_base, d_flag; DevBuf<u8> d_map, d_state;
    d_off.alloc(nfiles, s); d_end.alloc(nfiles, s); d_first.alloc(nfiles + 1, s);
    MMB_CUDA(cudaMemcpyAsync(d_off.p, file_off.data(), nfiles * 8, cudaMemcpyHostToDevice, s));
    MMB_CUDA(cudaMemcpyAsync(d_end.p, file_end.data(), nfiles * 8, cudaMemcpyHostToDevice, s));
    MMB_CUDA(cudaMemcpyAsync(d_first.p, chunk_first.data(), (nfiles + 1) * 4, cudaMemcpyHostToDevice, s));
    d_map.alloc(nchunks, s); d_state.alloc(nchunks, s); d_kept.alloc(nchunks + 1, s); d_hdr.alloc(nchunks + 1, s);
    d_kept_base.alloc(nchunks + 1, s); d_hdr_base.alloc(nchunks + 1, s); d_flag.alloc(1, s);
    d_flag.zero(); d_kept.zero(); d_hdr.zero();
    d_doc_start.alloc(nfiles + 1, s); d_doc_len.alloc(nfiles, s);

    if (nchunks) {
        DevBuf<u64> d_kept4, d_hdr4; d_kept4.alloc(nchunks, s); d_hdr4.alloc(nchunks, s);
        MMB_LAUNCH(fasta_chunk_maps, nchunks, TPB, 0, s, d_raw, d_off.p, d_end.p, d_first.p, nfiles, (int)raw_mode, d_map.p, d_kept4.p, d_hdr4.p);
        MMB_LAUNCH(fasta_chunk_states, ceil_div((u64)nfiles * 32, 128), 128, 0, s, (const u8*)d_map.p, (const u64*)d_kept4.p, (const u64*)d_hdr4.p, d_first.p, nfiles,
                   (int)raw_mode, d_state.p, d_kept.p, d_hdr.p);
    }
    prim::exclusive_sum(d_kept.p, d_kept_base.p, nchunks + 1, s);     // entry [nchunks] = total
    prim::exclusive_sum(d_hdr.p, d_hdr_base.p, nchunks + 1, s);
    u64 total_kept = 0; u32 total_hdr = 0;
    MMB_CUDA(cudaMemcpyAsync(&total_kept, d_kept_base.p + nchunks, 8, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaMemcpyAsync(&total_hdr, d_hdr_base.p + nchunks, 4, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaStreamSynchronize(s));

    // documents of this group start where the text ends so far
    MMB_LAUNCH(fasta_doc_layout, 1, 1, 0, s, d_kept_base.p, d_first.p, nfiles, nchunks, total_kept, (int)revcomp, out.n, d_doc_start.p, d_doc_len.p);
    DevBuf<u64> d_rec_pos, d_rec_before;
    d_rec_pos.alloc(total_hdr, s); d_rec_before.alloc(total_hdr, s);
    MMB_LAUNCH(fasta_terminators, ceil_div(nfiles, 128), 128, 0, s, d_doc_start.p, d_doc_len.p, nfiles, (int)revcomp, out.text.p);
    if (nchunks) {
        auto fasta_scatter_pass = fasta_pass<true>;
        MMB_LAUNCH(fasta_scatter_pass, nchunks, TPB, 0, s, d_raw, d_off.p, d_end.p, d_first.p, nfiles, d_state.p, (int)raw_mode, d_kept.p, d_hdr.p, d_flag.p,
                   d_kept_base.p, d_hdr_base.p, d_doc_start.p, d_doc_len.p, (int)revcomp, out.text.p, d_rec_pos.p, d_rec_before.p);
    }
    if (to_host(d_flag.p, 1, s)[0]) throw Error(13, "FASTQ records ('+' separator line) are not supported");      // seen by the scatter pass
    std::vector<u64> ds = to_host(d_doc_start.p, nfiles + 1, s), dl = to_host(d_doc_len.p, nfiles, s);
    std::vector<u64> rp = to_host(d_rec_pos.p, total_hdr, s), rb = to_host(d_rec_before.p, total_hdr, s);
    for (u32 f = 0; f < nfiles; ++f) { out.h_doc_len.push_back(dl[f]); out.h_doc_start.push_back(ds[f + 1]); }
    for (u32 r = 0; r < total_hdr; ++r) { out.h_rec_pos.push_back(rp[r]); out.h_rec_before.push_back(rb[r] + out.total_bases); }     // kept bases before it, over all groups
    out.n = ds[nfiles];
    out.total_bases += total_kept;
}

void text_finish(TextLayout& out, cudaStream_t s) {
    size_t nfiles = out.h_doc_len.size();
    out.doc_start.alloc(nfiles + 1, s); out.doc_len.alloc(nfiles, s);
    MMB_CUDA(cudaMemcpyAsync(out.doc_start.p, out.h_doc_start.data(), (nfiles + 1) * 8, cudaMemcpyHostToDevice, s));
    MMB_CUDA(cudaMemcpyAsync(out.doc_len.p, out.h_doc_len.data(), nfiles * 8, cudaMemcpyHostToDevice, s));
    MMB_CUDA(cudaMemsetAsync(out.text.p + out.n, 0, 128, s));
    MMB_CUDA(cudaStreamSynchronize(s));
}

void build_text(const u8* d_raw, const std::vector<u64>& file_off, const std::vector<u64>& file_size, bool revcomp, bool raw_mode,
                TextLayout& out, cudaStream_t s) {
    text_begin(file_size, revcomp, out, s);
    text_append(d_raw, file_off, file_size, 0, file_off.size(), raw_mode, out, s);
    text_finish(out, s);
}

}  // namespace mmb
