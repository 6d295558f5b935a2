/* oracle/ref_shims/gsacak.h -- TEST INFRASTRUCTURE.
 * Declarations matching the call sites of the reference's un-vendored dependency github.com/oma219/gsa-is
 * (unpinned; /root/reference/thirdparty/CMakeLists.txt:59-74) under -DM64:
 *   include/dictionary.hpp:133   gsacak(&d[0], &saD[0], &lcpD[0], nullptr, d.size())
 *   include/parse.hpp:85         sacak_int(&p[0], &saP[0], p.size(), alphabet_size)
 *   include/direct_gsacak.hpp:62 gsacak(&text[0], &sa[0], &lcp[0], nullptr, text.size())
 * Published semantics restated: symbols 0 and 1 are separators; the final symbol is 0 and sorts first; the
 * 1-separators sort next, in text order; the LCP never extends across a separator.  Implementation: gsacak_shim.cpp. */
#ifndef GSACAK_SHIM_H
#define GSACAK_SHIM_H
#include <stdint.h>
typedef uint64_t uint_t;
typedef int64_t int_t;
typedef uint32_t int_text;
typedef int32_t int_da;
int gsacak(unsigned char* s, uint_t* SA, int_t* LCP, int_da* DA, uint_t n);
int sacak_int(int_text* s, uint_t* SA, uint_t n, uint_t k);
#endif
