/* include/mumemto.h -- the reference's library C ABI, re-declared for the GPU engine (drop-in for libmumemto).
 *
 * Symbol names, argument order and struct layouts follow mumemto_library/mumemto.h:33-94 of the reference so that a program
 * built against that header links against libmumemto_b200.so unchanged.  Implementation: mumemto_b200/csrc/library_api.cpp.
 * Return codes of mumemto_mum / mumemto_mem: 0 ok, 1 out_result is NULL, 2 docs is NULL, 3 error (see mumemto_last_error), 4 unknown
 * (mumemto_library/mumemto_api.cpp:495-537).  Results are owned by the caller until mum_free / mem_free. */
#ifndef MUMEMTO_H
#define MUMEMTO_H

#include <stddef.h>
#include <stdint.h>

#define MUMEMTO_C_API __attribute__((visibility("default")))

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mumemto_doc_view {
    const char* const* records;   /* NUL-terminated sequence records of one document */
    size_t num_records;
} mumemto_doc_view;

typedef struct mumemto_mum_result mumemto_mum_result;
typedef struct mumemto_mem_result mumemto_mem_result;

typedef struct mumemto_mum_match_view {
    uint32_t length;
    const int64_t* offsets;   /* num_docs entries, -1 = absent */
    const uint8_t* strands;   /* 1 => '+', 0 => '-' */
} mumemto_mum_match_view;

typedef struct mumemto_mem_match_view {
    uint32_t length;
    size_t occurrences;
    const int64_t* offsets;
    const size_t* seq_ids;
    const uint8_t* strands;
} mumemto_mem_match_view;

MUMEMTO_C_API const char* mumemto_last_error(void);

MUMEMTO_C_API int mumemto_mum(const mumemto_doc_view* docs, size_t num_docs, uint32_t min_match_len, uint8_t use_revcomp, size_t num_distinct,
                              uint8_t use_gsacak, mumemto_mum_result** out_result);
MUMEMTO_C_API int mumemto_mem(const mumemto_doc_view* docs, size_t num_docs, uint32_t min_match_len, uint8_t use_revcomp, size_t num_distinct,
                              size_t max_total_freq, size_t max_doc_freq, uint8_t use_gsacak, mumemto_mem_result** out_result);

MUMEMTO_C_API size_t num_docs(const mumemto_mum_result* r);
MUMEMTO_C_API const size_t* doc_record_offsets(const mumemto_mum_result* r);   /* num_docs(r) + 1 entries */
MUMEMTO_C_API const size_t* record_lengths(const mumemto_mum_result* r);
MUMEMTO_C_API size_t num_mums(const mumemto_mum_result* r);
MUMEMTO_C_API mumemto_mum_match_view mum_at(const mumemto_mum_result* r, size_t idx);
MUMEMTO_C_API void mum_free(mumemto_mum_result* r);

MUMEMTO_C_API size_t num_docs_mem(const mumemto_mem_result* r);
MUMEMTO_C_API const size_t* doc_record_offsets_mem(const mumemto_mem_result* r);
MUMEMTO_C_API const size_t* record_lengths_mem(const mumemto_mem_result* r);
MUMEMTO_C_API size_t num_mems(const mumemto_mem_result* r);
MUMEMTO_C_API mumemto_mem_match_view mem_at(const mumemto_mem_result* r, size_t idx);
MUMEMTO_C_API void mem_free(mumemto_mem_result* r);

#ifdef __cplusplus
}
#endif
#endif /* MUMEMTO_H */
