// mumemto_b200/csrc/sa.h -- suffix array / LCP / BWT producers (sa.cu, lcp.cu).
#pragma once
#include "common.cuh"

namespace mmb {

struct SuffixArray {
    DevBuf<u32> sa;     // m entries; sa[0] = n (sentinel suffix)
    DevBuf<u32> isa;    // m entries; inverse permutation once construction has finished
    u32 m = 0;
    u32 rounds = 0;             // doubling rounds run
    u64 sorted_elements = 0;    // total (key,value) pairs pushed through the radix sort
};

void build_suffix_array(const u8* text, u64 n, SuffixArray& out, cudaStream_t s);

// lcp[r] = longest common prefix of suffixes sa[r-1] and sa[r] (lcp[0] = 0); bwt[r] = text[sa[r]-1] (0 when sa[r] = 0).
void build_lcp_bwt(const u8* text, u64 n, const SuffixArray& sa, DevBuf<u32>& lcp, DevBuf<u8>& bwt, cudaStream_t s);

}  // namespace mmb
