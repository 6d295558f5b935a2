// mumemto_b200/csrc/merge.h -- anchor-based partition merge (merge.cu).
#pragma once
#include "common.cuh"

namespace mmb {

struct MergePartition {     // the merged table
    u32 n = 0, ndocs = 0;
    DevBuf<u32> len;        // [n]
    DevBuf<i64> off;        // [n][ndocs], rows in anchor order
    DevBuf<u8> plus;        // [n][ndocs], 1 = '+'
    DevBuf<u64> start;      // [n] = column 0 (anchor offsets, ascending)
    DevBuf<u16> nb;         // [anchor_len] next-best thresholds
};

// One input partition; all pointers are device pointers, rows in any order.  strand cells: 1 / '+' = forward strand, anything else '-'.
struct MergeSource {
    const u32* len; const i64* off; const u8* strand; const u16* nb;      // [n], [n][ndocs], [n][ndocs], [anchor_len]
    u64 n; u32 ndocs;
};

// Left fold of k partitions (2 <= k <= 64), exactly as `anchor_merge p1 p2 ...` of the reference folds them.
void merge_partitions(const MergeSource* src, size_t k, u64 anchor_len, MergePartition& out, cudaStream_t s);
void merge_text(const MergePartition& P, DevBuf<u8>& text, u64& bytes, cudaStream_t s);

}  // namespace mmb
