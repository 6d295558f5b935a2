// mumemto_b200/csrc/sa.h -- suffix array / LCP / BWT producers (sa.cu, lcp.cu).
#pragma once
#include "common.cuh"

namespace mmb {

struct SaStats {
    u32 rounds = 0;             // doubling rounds run
    u32 initial_groups = 0;     // multi-member groups after the initial key sort
    u64 groups_sorted = 0;      // group sorts over all rounds
    u64 large_elements = 0;     // elements that went through the device-wide radix sort inside rounds
};

struct SuffixArray {
    DevBuf<u32> sa;     // m entries; sa[0] = n (sentinel suffix)
    DevBuf<u32> isa;    // m entries; inverse permutation once construction has finished
    u32 m = 0;
    SaStats stats;
};

void build_suffix_array(const u8* text, u64 n, SuffixArray& out, cudaStream_t s);

// Multi-member groups of one doubling round, one list per size class (sa.cu), each entry head << 32 | size.
struct GroupStore {
    static constexpr int kClasses = 9;
    DevBuf<u64> groups;
    DevBuf<u32> counts;             // kClasses counters
    u64 first[kClasses + 1] = {};   // offset of each class list inside groups
    void alloc(u32 m, cudaStream_t s);
};

// Shared refinement engine (also used by pfp.cu).  sa is ordered by the first h0 symbols, isa holds head ranks, groups holds the
// multi-member groups.  limit (optional): per suffix, the number of symbols after which nothing is left to compare (dictionary
// suffixes end at their terminator and may tie).
// isa_stride 2 (with limit): isa holds 2 m words, pairs (head rank, offset of the position inside its phrase); the caller fills the
// offsets beforehand and the rounds then never read limit[].
void order_by_initial_keys(u64* keys_in, u64* keys_tmp, u32* vals_in, u32 m, int key_bits, u32 h0, const u32* limit, u32* sa, u32* isa,
                           GroupStore& groups, cudaStream_t s, u32 isa_stride = 1);
void refine_by_doubling(u32* sa, u32* isa, u32 m, u32 h0, const u32* limit, GroupStore& groups, SaStats& st, cudaStream_t s, u32 isa_stride = 1);

// Pruned prefix-free-parse route: exact order of the dictionary suffixes inside the key groups whose ranks are marked in refine[]
// (sorted_keys = the initial cpk-symbol keys in rank order, sa = positions in that order), by comparing further cpk-symbol chunks of the
// dictionary bytes; ghead[rank] = first rank of the run of EQUAL strings the rank belongs to (its own rank when it is alone).  Returns
// false when max_rounds chunks were not enough (the caller falls back to the doubling engine above).
bool refine_dictionary_chunks(const u64* sorted_keys, u32* sa, u32 m, u32 cpk, u32 bits, const u8* dict, const u32* limit, const u8* lut,
                              const u8* refine, u32* ghead, SaStats& st, cudaStream_t s, u32 max_rounds);

// lcp[r] = longest common prefix of suffixes sa[r-1] and sa[r] (lcp[0] = 0), followed by 64 zero entries.
void build_lcp(const u8* text, u64 n, const SuffixArray& sa, DevBuf<u32>& lcp, cudaStream_t s);
// bwt[r] = text[sa[r]-1] (0 when sa[r] = 0)
void build_bwt(const u8* text, const SuffixArray& sa, DevBuf<u8>& bwt, cudaStream_t s);

}  // namespace mmb
