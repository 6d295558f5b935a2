/* include/mumemto_b200.h -- C ABI of the B200 multi-MUM / multi-MEM engine (libmumemto_b200.so).
 *
 * This is the drop-in boundary for the reference's FASTA -> .mums/.mems hot path.  Plain pointers and sizes only; no CUDA, torch or
 * C++ types.  Each entry point names the reference interface it stands in for (paths relative to the reference repository):
 *
 *   mmb_load_fasta / _device   RefBuilder::build_input_file            src/ref_builder.cpp:211-314
 *   mmb_load_records           RefBuilder::build_input_file_lib        src/ref_builder.cpp:330-384
 *   mmb_find                   the body of build_main after parsing    src/pfp_mum.cpp:76-94,116-145
 *   mmb_find_from_arrays       the -a branch: RefBuilder(prefix) +     src/ref_builder.cpp:140-168, src/pfp_mum.cpp:96-110,
 *                              file_lcp::process                       include/read_arrays.hpp:66-84
 *                              = gsacak_lcp / pf_parsing + pfp_lcp::process feeding mem_finder::update
 *                              (include/direct_gsacak.hpp:50-116, include/pfp_lcp_mum.hpp:115-231, include/mem_finder.hpp:161-170)
 *   mmb_options                BuildOptions + validate/set_parameters  include/pfp_mum.hpp:56-198
 *   mmb_text / mmb_mum_* / mmb_mem_*   mem_finder::write_mum / write_mem / write_bums   include/mem_finder.hpp:210-263,357-428,460-503
 *   mmb_thresholds             candidate_thresh (.athresh)             include/mem_finder.hpp:110-115,326-336
 *   mmb_arrays                 -A array dump (.sa/.lcp/.bwt)           include/pfp_lcp_mum.hpp:323-369
 *   mmb_anchor_merge           merge_partitions left fold              src/merge_candidates.cpp:106-157,208-219
 *
 * Every function returns 0 on success, a non-zero code otherwise; mmb_last_error() then holds the message (thread local).
 * The library has no CPU implementation: creating an engine without a usable CUDA device fails with MMB_ERR_NO_DEVICE.
 */
#ifndef MUMEMTO_B200_H
#define MUMEMTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMB_API __attribute__((visibility("default")))

enum {
    MMB_OK = 0,
    MMB_ERR_ARG = 1,          /* null / inconsistent argument */
    MMB_ERR_NO_DEVICE = 2,    /* no CUDA device: there is no CPU fallback */
    MMB_ERR_INPUT = 3,        /* bad input data (empty file, FASTQ, NUL byte, < 2 documents ...) */
    MMB_ERR_PARAM = 4,        /* parameter combination the reference rejects (FATAL_ERROR) or the engine does not support */
    MMB_ERR_CUDA = 5,
    MMB_ERR_STATE = 6,        /* call order: results requested before mmb_find, ... */
    MMB_ERR_INTERNAL = 7
};

typedef struct mmb_engine mmb_engine;

/* Flags of the reference CLI that reach the hot path (defaults: include/pfp_mum.hpp:56-78). */
typedef struct mmb_options {
    uint32_t min_match_len;      /* -l, default 20 */
    int32_t num_distinct_docs;   /* -k as typed (0, negative = relative to N); library callers pass the absolute count */
    int32_t rare_freq;           /* -f, default 1 (1 = MUM mode, other = MEM mode) */
    int32_t max_mem_freq;        /* -F, default 0 */
    uint8_t merge;               /* -M */
    uint8_t anchor_merge;        /* -n */
    uint8_t cli_normalise;       /* 1: apply BuildOptions::validate/set_parameters; 0: library rules (mumemto_api.cpp:332-411) */
    uint8_t want_text;           /* 1: also format the .mums/.mems bytes on the device */
    uint8_t route;               /* 0 auto; 1 direct suffix sorting of the text (|T| < 2^31); 2 prefix-free parse (dictionary + parse), any |T| < 2^40 with -F below 2^20 */
    uint8_t keep_arrays;         /* 1: keep SA/LCP of the whole stream for mmb_copy_arrays even on the tiled (route 2) path */
    uint8_t no_prune;            /* 1: enumerate every suffix on route 2 (no segment pruning; results are identical) */
    uint8_t shard_count;         /* > 1 = this call reports only the matches of shard `shard_index` of `shard_count` contiguous slices of
                                    the suffix order (cut at segment boundaries, which no reportable interval crosses); the outputs of the shards,
                                    concatenated in shard order, are the bytes of the unsharded call.  Needs the pruned prefix-free-parse route
                                    (forced); not with merge / keep_arrays / no_prune.  0 or 1 = everything */
    uint8_t shard_index;
    uint8_t reserved[3];
    uint64_t tile_ranks;         /* route 2: ranks per stream tile, 0 = default (2^30 in mmb_find = ~18 GB of tile buffers; 2^26 in mmb_find_from_arrays) */
    uint32_t trim_min_len;       /* 0 = off (the reference's output, bit-exact).  > 0: multi-MUM rows are post-processed like the reference's
                                    `mumemto trim` (mumemto/trim_mums.py:51-86): shortened by their overhang past the end of a sequence,
                                    dropped when shorter than this afterwards (its default is 20) */
    uint32_t reserved2;
} mmb_options;

typedef struct mmb_stats {
    uint64_t text_symbols;       /* |T| */
    uint64_t input_bases;        /* sum of forward bases */
    uint64_t num_matches;
    uint64_t kernel_launches;    /* launches issued by this library since the engine was created */
    uint32_t sa_rounds;
    uint64_t sa_sorted_elements;
    float ms_text, ms_sa, ms_lcp, ms_detect, ms_emit, ms_total;   /* CUDA-event times of the last mmb_find (ms_text: last load) */
    uint32_t route;              /* route taken by the last mmb_find (1 direct, 2 prefix-free parse, 3 arrays given by the caller) */
    uint32_t pfp_phrases, pfp_distinct_phrases, pfp_dict_bytes, pfp_tiles;
    uint32_t pruned;             /* route 2: 1 when segments that cannot hold a match were left out of the suffix stream */
    uint64_t stream_suffixes;    /* ranks that were enumerated and scanned (|T| + 1 without pruning) */
    uint64_t threshold_segments; /* -M, pruned: segments whose threshold was taken without enumerating them */
} mmb_stats;

MMB_API const char* mmb_last_error(void);
MMB_API void mmb_default_options(mmb_options* o);

/* cuda_stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL for the engine's own stream. */
MMB_API int mmb_create(int device, void* cuda_stream, mmb_engine** out);
MMB_API void mmb_destroy(mmb_engine* e);

/* Page-locked host memory for inputs and outputs (copies to and from it run at full PCIe speed).  An engine must exist on the calling
 * thread's device before the first call.  NULL on failure. */
MMB_API void* mmb_alloc_pinned(size_t bytes);
MMB_API void mmb_free_pinned(void* p);

/* Input.  files[i] is the (decompressed) byte image of FASTA file i in host memory.  The copy to the device happens here. */
MMB_API int mmb_load_fasta(mmb_engine* e, const uint8_t* const* files, const uint64_t* sizes, size_t nfiles, int use_revcomp);
/* Same, but the images already sit in device memory at d_raw + offsets[i] (offsets multiples of 16; the buffer must be readable for 64
 * bytes past the last image).  Record names for PREFIX.lengths are fetched from the device copy (cut at 255 bytes). */
MMB_API int mmb_load_fasta_device(mmb_engine* e, const void* d_raw, const uint64_t* offsets, const uint64_t* sizes, size_t nfiles, int use_revcomp);
/* Library-style input: document d has n_records[d] NUL-terminated records, flattened in records[] (no FASTA syntax). */
MMB_API int mmb_load_records(mmb_engine* e, const char* const* records, const size_t* n_records, size_t ndocs, int use_revcomp);

/* The hot path.  Everything stays on the device until one of the copy functions below is called. */
MMB_API int mmb_find(mmb_engine* e, const mmb_options* opts);
/* Detection over PRE-COMPUTED arrays in the reference's on-disk format (-A output: 5-byte little-endian SA and LCP entries, one BWT
 * byte per rank), host memory.  doc_lengths = forward bases per document (PREFIX.lengths); `entries` ranks are processed, starting at
 * rank 0 -- the reference's reader stops after |T| of the |T|+1 entries of the files, and so does its CLI twin here.  Replaces any
 * loaded text; results are read with the same accessors as after mmb_find (mmb_copy_arrays is not available afterwards). */
MMB_API int mmb_find_from_arrays(mmb_engine* e, const uint64_t* doc_lengths, size_t ndocs, int use_revcomp, const uint8_t* sa5,
                                 const uint8_t* lcp5, const uint8_t* bwt, uint64_t entries, const mmb_options* opts);
MMB_API int mmb_get_stats(const mmb_engine* e, mmb_stats* out);

/* Documents and FASTA records (for PREFIX.lengths, src/ref_builder.cpp:193-209). */
MMB_API size_t mmb_num_docs(const mmb_engine* e);
MMB_API int mmb_doc_lengths(const mmb_engine* e, uint64_t* lengths /* [num_docs] forward bases */);
MMB_API size_t mmb_num_records(const mmb_engine* e, size_t doc);
MMB_API int mmb_record(const mmb_engine* e, size_t doc, size_t rec, const char** name, uint64_t* length);

/* Results of the last mmb_find. */
MMB_API int mmb_is_mum_mode(const mmb_engine* e);
MMB_API uint64_t mmb_num_matches(const mmb_engine* e);
MMB_API uint64_t mmb_text_size(const mmb_engine* e);                 /* bytes of the .mums / .mems file */
MMB_API int mmb_copy_text(mmb_engine* e, void* host_dst);            /* device -> host */
MMB_API const void* mmb_text_device(const mmb_engine* e);            /* device pointer, valid until the next mmb_find */
/* MUM table: lengths[n], offsets[n][num_docs] (-1 absent), strands[n][num_docs] (1 '+', 0 '-' or absent). NULL skips a field. */
MMB_API int mmb_copy_mums(mmb_engine* e, uint32_t* lengths, int64_t* offsets, uint8_t* strands);
/* .bumbl strand bits (MSB first, '+' = 1), (n * num_docs + 7) / 8 bytes. */
MMB_API int mmb_copy_mum_strand_bits(mmb_engine* e, uint8_t* packed);
/* MEM table in CSR form: first[n+1]; offsets/seq_ids/strands[total occurrences]. */
MMB_API uint64_t mmb_num_mem_occurrences(const mmb_engine* e);
MMB_API int mmb_copy_mems(mmb_engine* e, uint32_t* lengths, uint64_t* first, int64_t* offsets, uint64_t* seq_ids, uint8_t* strands);
/* -M: 2 * (L_0 + 1) uint16 next-best thresholds; PREFIX.athresh is the first L_0 + 1 of them. */
MMB_API uint64_t mmb_threshold_count(const mmb_engine* e);
MMB_API int mmb_copy_thresholds(mmb_engine* e, uint16_t* dst);
/* -A: the suffix stream, |T| + 1 entries each. NULL skips an array. */
MMB_API int mmb_copy_arrays(mmb_engine* e, uint64_t* sa, uint64_t* lcp, uint8_t* bwt);
/* Device pointers of the last MUM table (lengths u32[n], offsets i64[n][num_docs], strands u8[n][num_docs] as '+', '-', 0) and of
 * the threshold array, e.g. to hand them to NCCL or to mmb_anchor_merge without leaving the GPU.  Valid until the next mmb_find. */
MMB_API int mmb_mum_table_device(mmb_engine* e, const void** lengths, const void** offsets, const void** strands, const void** thresholds);

/* Anchor-based merge of nparts partitions, folded left to right like `anchor_merge p1 p2 ...` (src/merge_candidates.cpp:208-219).
 * Partition p: n_mums[p] strict MUMs over ndocs[p] sequences whose first sequence is the shared anchor; lengths[p] u32[n],
 * offsets[p] i64[n][ndocs], strands[p] u8[n][ndocs] (1 or '+' = forward), thresholds[p] u16[anchor_len] (= PREFIX.athresh).
 * inputs_on_device: 1 if all those pointers are device pointers (e.g. an all-gather result), 0 for host memory. */
MMB_API int mmb_anchor_merge(mmb_engine* e, size_t nparts, const void* const* lengths, const uint64_t* n_mums, const void* const* offsets,
                             const void* const* strands, const void* const* thresholds, const uint64_t* ndocs, uint64_t anchor_len,
                             int inputs_on_device);
MMB_API uint64_t mmb_merge_num_mums(const mmb_engine* e);
MMB_API uint64_t mmb_merge_num_docs(const mmb_engine* e);
MMB_API int mmb_merge_copy(mmb_engine* e, uint32_t* lengths, int64_t* offsets, uint8_t* strands /* 1 = '+' */, uint16_t* thresholds);
MMB_API uint64_t mmb_merge_text_size(mmb_engine* e);                /* merged .mums bytes (include/mumsio.hpp:283-296,318-326); formatted on first use */
MMB_API int mmb_merge_copy_text(mmb_engine* e, void* host_dst);
MMB_API const void* mmb_merge_text_device(mmb_engine* e);           /* device pointer of the merged .mums bytes, valid until the next merge */

/* Measurement aid for bench.py: bracket every launch of the kernels whose name contains `name_substring` with CUDA events on the
 * launching stream (NULL or "" switches it off and clears the record); mmb_profile_result adds up the recorded launches. */
MMB_API int mmb_profile_kernel(const char* name_substring);
MMB_API int mmb_profile_result(double* total_ms, uint64_t* launches);
/* With mmb_profile_kernel("*") every launch is bracketed; this returns one line "kernel\tms\tlaunches\n" per kernel name (library
 * sorts and scans appear as cub_radix_sort / cub_scan).  *needed = bytes required including the terminating NUL. */
MMB_API int mmb_profile_table(char* dst, size_t cap, size_t* needed);

/* The text T itself (|T| bytes), for tests. */
MMB_API int mmb_copy_symbols(mmb_engine* e, uint8_t* dst);

#ifdef __cplusplus
}
#endif
#endif /* MUMEMTO_B200_H */
