// mumemto_b200/csrc/text.h -- device text layout produced by text.cu.
#pragma once
#include "common.cuh"

namespace mmb {

struct TextLayout {
    DevBuf<u8> text;            // n symbols (+64 zero bytes of slack)
    u64 n = 0;                  // |T|
    u64 total_bases = 0;        // sum of forward bases (the metric's numerator)
    bool revcomp = true;
    DevBuf<u64> doc_start;      // N+1: first symbol of each doc; [N] = n
    DevBuf<u64> doc_len;        // N: forward bases L_d
    std::vector<u64> h_doc_start, h_doc_len;
    std::vector<u64> h_rec_pos;     // per FASTA record: absolute offset of its '>' in the raw buffer
    std::vector<u64> h_rec_before;  // per FASTA record: number of kept bases before it (global)
};

// d_raw holds the files back to back at 16-byte aligned offsets file_off[f] (sizes file_size[f]).
// raw_mode: bytes are sequence already (library entry points), no FASTA syntax.
void build_text(const u8* d_raw, const std::vector<u64>& file_off, const std::vector<u64>& file_size, bool revcomp, bool raw_mode,
                TextLayout& out, cudaStream_t s);

// The same in steps, so that the upload of later files can overlap the text build of earlier ones (engine.cu, mmb_load_fasta):
// text_begin sizes the text for the worst case (every byte a base), text_append builds the documents of files [f0, f1) behind the ones
// already there (it only needs the raw bytes of those files to be on the device), text_finish publishes the layout.
void text_begin(const std::vector<u64>& file_size, bool revcomp, TextLayout& out, cudaStream_t s);
void text_append(const u8* d_raw, const std::vector<u64>& file_off, const std::vector<u64>& file_size, size_t f0, size_t f1, bool raw_mode,
                 TextLayout& out, cudaStream_t s);
void text_finish(TextLayout& out, cudaStream_t s);

}  // namespace mmb
