// mumemto_b200/csrc/multi_gpu.h -- partition-merge over several GPUs inside ONE process (the drop-in executable's --gpus N).
//
// Replaces the reference's multi-core workflow `mumemto merge` (mumemto/merge_mums.py:171-183): P x `mumemto_exec -M -n` on partitions
// that share their first sequence, then `anchor_merge` (src/merge_candidates.cpp:170-255).  Here: one host thread per GPU runs its
// partition through the C ABI (include/mumemto_b200.h), ONE ncclAllGather moves the packed partition tables (lengths, offsets, strands,
// thresholds) between the GPUs, and the merge kernels fold them on the first device; PREFIX.mums / PREFIX.athresh are byte-identical
// to what `anchor_merge p0.mums p1.mums ...` writes for the same partitions.  Host code only: plain C++ on top of the C ABI, the CUDA
// runtime (buffers, peer-less copies) and NCCL.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "fasta_io.h"

namespace mgpu {

struct Request {
    const std::vector<fio::Source>* sources = nullptr;      // the input files (fio::probe), anchor first
    int gpus = 1;                               // devices first_device .. first_device + gpus - 1
    int partitions = 0;                         // 0 = one per GPU; > gpus only with gpus == 1 (run one after the other, no NCCL)
    int first_device = 0;
    bool use_rcomp = true;
    uint32_t min_match_len = 20;
    bool binary = false;                        // PREFIX.bumbl instead of PREFIX.mums
    std::string output_prefix;
};

// Partition p holds the anchor (file 0) and a contiguous slice of the other files, so the merged columns come out in input order.
std::vector<std::vector<size_t>> make_partitions(size_t nfiles, size_t parts);

// Returns false with `error` set on failure.
bool run(const Request& rq, std::string& error);

// ONE collection over several GPUs: every GPU holds the whole collection and reports the matches of its slice of the
// suffix order (mmb_options.shard_index / shard_count; slices are cut at segment boundaries, which no reportable interval crosses); the
// slices are written one after the other, so PREFIX.mums / .mems / .bumbl are the bytes of the single-GPU run.  No collective: the only
// exchange is the final gather of output bytes through the host.  With gpus == 1 the slices run one after the other (tests).
struct ShardRequest {
    const std::vector<fio::Source>* sources = nullptr;
    int gpus = 1, shards = 2, first_device = 0;
    bool use_rcomp = true, binary = false;
    uint32_t min_match_len = 20;
    int num_distinct_docs = 0, rare_freq = 1, max_mem_freq = 0;
    uint32_t trim_min_len = 0;
    std::string output_prefix;
};
bool run_sharded(const ShardRequest& rq, std::string& error);

}  // namespace mgpu
