// mumemto_b200/csrc/detect.h -- LCP-interval detection (detect.cu) and output formatting (emit.cu).
#pragma once
#include "common.cuh"

namespace mmb {

struct DetectParams {
    u32 min_len;        // -l
    u32 num_distinct;   // -k after normalisation
    u32 max_freq;       // -F after normalisation; 0 = no limit (not supported by the windowed scan)
    u32 max_doc_freq;   // -f; 0 = no per-document limit
    u32 ndocs;
    int revcomp;
    int merge;          // -M: record next-best thresholds
    // optional: document of text position p = doc_coarse[p >> coarse_shift], then forward while doc_start[d + 1] <= p (at most 4096
    // entries, built by the engine when the text is loaded); nullptr = binary search over doc_start
    const u16* doc_coarse;
    u32 coarse_shift;
    // the arrays handed to detect_intervals are the WHOLE suffix stream (not a tile): required for max_freq == 0 && max_doc_freq == 0,
    // where an interval may be of any size
    int whole_stream;
};

// One reported interval [s, e] of ranks, all suffixes sharing a prefix of length len.  key = (e + 1) << 20 | order within the
// closing rank, so that sorting by key reproduces the reference's output order (closing rank ascending, length descending).
struct Report {
    u64 key;
    u64 s;
    u32 len;
    u32 pad;
};

struct Reports {
    DevBuf<Report> rep;     // sorted by key
    u64 count = 0;
};

// thresh: 2*(L_0+1) zero-initialised uint16 (only when prm.merge), else nullptr.  bwt may be nullptr: the preceding symbol is then
// read from text[sa - 1] for the few ranks that need it.  lcp must be readable (zero) for 64 entries past m.
// The arrays are a view of the suffix stream (the whole stream, or one tile with a halo in front); only closing ranks j >= j_first
// of the view are examined, and report keys / starts are view indices.
// docs (optional): one document-id byte per rank of the view (StreamTile::doc); the strict-MUM path then checks candidates from them.
template <class SAT>
void detect_intervals(const u32* lcp, const SAT* sa, const u8* bwt, const u8* text, u64 m, u64 j_first, const DetectParams& prm,
                      const u64* d_doc_start, u16* thresh, Reports& out, cudaStream_t s, const u8* docs = nullptr);

struct MumTable {           // strict / partial multi-MUMs that survive the strand rules, in output order
    u64 count = 0;
    DevBuf<u32> len;        // [count]
    DevBuf<i64> off;        // [count][ndocs], -1 = absent
    DevBuf<u8> strand;      // [count][ndocs], '+', '-', 0 = absent
    DevBuf<u8> text;        // .mums bytes
    u64 text_bytes = 0;
};

struct MemTable {           // multi-MEMs: CSR over occurrences
    u64 count = 0;
    u64 total_occ = 0;
    DevBuf<u32> len;        // [count]
    DevBuf<u64> first;      // [count+1] occurrence offsets
    DevBuf<u64> off;        // [total_occ] (unsigned, wraps like the reference)
    DevBuf<u32> doc;        // [total_occ]
    DevBuf<u8> strand;      // [total_occ]
    DevBuf<u8> text;        // .mems bytes
    u64 text_bytes = 0;
};

template <class SAT>
void emit_mums(const Reports& rep, const SAT* sa, const DetectParams& prm, const u64* d_doc_start, const u64* d_doc_len, bool want_text,
               MumTable& out, cudaStream_t s);
template <class SAT>
void emit_mems(const Reports& rep, const SAT* sa, const DetectParams& prm, const u64* d_doc_start, const u64* d_doc_len, bool want_text,
               MemTable& out, cudaStream_t s);

// Rows only (mum_rows / mem_rows), text from rows (mum_text / mem_text), and concatenation of per-tile tables in stream order.
template <class SAT>
void mum_rows(const Reports& rep, const SAT* sa, const DetectParams& prm, const u64* d_doc_start, const u64* d_doc_len, MumTable& out, cudaStream_t s);
void mum_text(MumTable& t, u32 ndocs, cudaStream_t s);
// optional post-processing of the reference (mumemto/trim_mums.py:51-86): shorten rows that overhang their document, drop rows below min_length
void trim_mum_table(MumTable& t, u32 ndocs, const u64* d_doc_len, u32 min_length, cudaStream_t s);
template <class SAT>
void mem_rows(const Reports& rep, const SAT* sa, const DetectParams& prm, const u64* d_doc_start, const u64* d_doc_len, MemTable& out, cudaStream_t s);
void mem_text(MemTable& t, cudaStream_t s);
void concat_mums(std::vector<MumTable>& parts, u32 ndocs, MumTable& out, cudaStream_t s);
void concat_mems(std::vector<MemTable>& parts, MemTable& out, cudaStream_t s);

// .bumbl strand bits: MSB-first, '+' = 1, bit index = row * ndocs + doc.
void pack_strands(const u8* strand, u64 nbits, DevBuf<u8>& packed, cudaStream_t s);

}  // namespace mmb
