// mumemto_b200/csrc/merge.h -- anchor-based partition merge (merge.cu).
#pragma once
#include "common.cuh"

namespace mmb {

struct MergePartition {
    u32 n = 0, ndocs = 0;
    DevBuf<u32> len;        // [n]
    DevBuf<i64> off;        // [n][ndocs], rows sorted by column 0
    DevBuf<u8> plus;        // [n][ndocs], 1 = '+'
    DevBuf<u64> start;      // [n] = column 0 (anchor offsets, ascending)
    DevBuf<u16> nb;         // [anchor_len] next-best thresholds
};

// All pointers are device pointers.  strand cells may be 1 / '+' for the forward strand, anything else is '-'.
void merge_prepare(const u32* len, const i64* off, const u8* strand, u64 n_mums, u32 ndocs, const u16* nb, u64 anchor_len, MergePartition& out,
                   cudaStream_t s);
void merge_fold(const MergePartition& A, const MergePartition& B, u64 anchor_len, MergePartition& R, cudaStream_t s);
void merge_text(const MergePartition& P, DevBuf<u8>& text, u64& bytes, cudaStream_t s);

}  // namespace mmb
