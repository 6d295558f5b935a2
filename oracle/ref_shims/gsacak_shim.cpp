// oracle/ref_shims/gsacak_shim.cpp -- TEST INFRASTRUCTURE.  See gsacak.h for what this replaces and why.
#include <vector>
#include "../sais.hpp"
extern "C" {
#include "gsacak.h"
}

extern "C" int gsacak(unsigned char* s, uint_t* SA, int_t* LCP, int_da* DA, uint_t n) {
    (void)DA;
    // Make every separator a distinct symbol: terminator -> 0, k-th '1' -> k, byte c -> c + m.
    int64_t m = 0;
    for (uint_t i = 0; i + 1 < n; ++i) if (s[i] <= 1) ++m;
    std::vector<int64_t> t(n);
    int64_t k = 0;
    for (uint_t i = 0; i + 1 < n; ++i) t[i] = (s[i] <= 1) ? ++k : (int64_t)s[i] + m;
    t[n - 1] = 0;
    std::vector<int64_t> sa(n);
    oracle::suffix_array(t.data(), sa.data(), (int64_t)n, 256 + m + 1);
    for (uint_t i = 0; i < n; ++i) SA[i] = (uint_t)sa[i];
    if (LCP) oracle::lcp_kasai(t.data(), sa.data(), LCP, (int64_t)n);
    return 0;
}

extern "C" int sacak_int(int_text* s, uint_t* SA, uint_t n, uint_t k) {
    std::vector<int64_t> t(s, s + n), sa(n);
    int64_t K = 0;
    for (uint_t i = 0; i < n; ++i) K = std::max<int64_t>(K, t[i] + 1);
    (void)k;
    oracle::suffix_array(t.data(), sa.data(), (int64_t)n, K);
    for (uint_t i = 0; i < n; ++i) SA[i] = (uint_t)sa[i];
    return 0;
}
