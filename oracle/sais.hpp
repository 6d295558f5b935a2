// oracle/sais.hpp -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
//
// Linear-time suffix array construction by induced sorting (SA-IS, Nong/Zhang/Chan 2009)
// over an integer alphabet, plus a Kasai-style LCP array.  Written from the published
// algorithm; it stands in for the reference's un-vendored third-party dependency
// github.com/oma219/gsa-is (FetchContent, unpinned; /root/reference/thirdparty/CMakeLists.txt:59-74),
// whose call sites are include/dictionary.hpp:133, include/parse.hpp:85 and
// include/direct_gsacak.hpp:62 in the reference.  A suffix array and an LCP array have
// exactly one correct value, so any correct builder pins the same result.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <algorithm>

namespace oracle {

// s[0..n): symbols in [0,K); s[n-1] must be the unique smallest symbol. SA: n entries.
template <class Sym, class Idx>
void sais_rec(const Sym* s, Idx* SA, Idx n, Idx K) {
    if (n == 1) { SA[0] = 0; return; }
    if (n == 2) { SA[0] = 1; SA[1] = 0; return; }
    std::vector<bool> stype(n);                   // true: S-type suffix
    stype[n - 1] = true;
    for (Idx i = n - 2; i >= 0; --i)
        stype[i] = (s[i] < s[i + 1]) || (s[i] == s[i + 1] && stype[i + 1]);
    auto is_lms = [&](Idx i) { return i > 0 && stype[i] && !stype[i - 1]; };

    std::vector<Idx> bkt((size_t)K);
    auto buckets = [&](bool ends) {
        std::fill(bkt.begin(), bkt.end(), (Idx)0);
        for (Idx i = 0; i < n; ++i) bkt[(size_t)s[i]]++;
        Idx sum = 0;
        for (Idx c = 0; c < K; ++c) { sum += bkt[c]; bkt[c] = ends ? sum : sum - bkt[c]; }
    };
    auto induce = [&]() {
        buckets(false);
        for (Idx i = 0; i < n; ++i) {
            Idx j = SA[i] - 1;
            if (SA[i] > 0 && !stype[j]) SA[bkt[(size_t)s[j]]++] = j;
        }
        buckets(true);
        for (Idx i = n - 1; i >= 0; --i) {
            Idx j = SA[i] - 1;
            if (SA[i] > 0 && stype[j]) SA[--bkt[(size_t)s[j]]] = j;
        }
    };

    // Stage 1: order the LMS substrings.
    std::fill(SA, SA + n, (Idx)-1);
    buckets(true);
    for (Idx i = 1; i < n; ++i) if (is_lms(i)) SA[--bkt[(size_t)s[i]]] = i;
    induce();

    Idx n1 = 0;
    for (Idx i = 0; i < n; ++i) if (is_lms(SA[i])) SA[n1++] = SA[i];
    std::fill(SA + n1, SA + n, (Idx)-1);
    Idx name = 0, prev = -1;
    for (Idx i = 0; i < n1; ++i) {
        Idx pos = SA[i];
        bool diff = (prev < 0);
        for (Idx d = 0; !diff; ++d) {
            if (s[pos + d] != s[prev + d] || stype[pos + d] != stype[prev + d]) { diff = true; break; }
            if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
        }
        if (diff) { ++name; prev = pos; }
        SA[n1 + pos / 2] = name - 1;
    }
    for (Idx i = n - 1, j = n - 1; i >= n1; --i) if (SA[i] >= 0) SA[j--] = SA[i];

    // Stage 2: order the LMS suffixes (recursively if names collide).
    Idx* s1 = SA + n - n1;
    Idx* SA1 = SA;
    if (name < n1) sais_rec<Idx, Idx>(s1, SA1, n1, name);
    else for (Idx i = 0; i < n1; ++i) SA1[s1[i]] = i;

    // Stage 3: induce the full order from the ordered LMS suffixes.
    for (Idx i = 1, j = 0; i < n; ++i) if (is_lms(i)) s1[j++] = i;
    for (Idx i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
    std::fill(SA + n1, SA + n, (Idx)-1);
    buckets(true);
    for (Idx i = n1 - 1; i >= 0; --i) {
        Idx j = SA[i]; SA[i] = -1;
        SA[--bkt[(size_t)s[j]]] = j;
    }
    induce();
}

// Plain suffix array of an integer string whose last symbol is the unique minimum.
inline void suffix_array(const int64_t* s, int64_t* SA, int64_t n, int64_t K) {
    sais_rec<int64_t, int64_t>(s, SA, n, K);
}
inline void suffix_array(const int32_t* s, int32_t* SA, int32_t n, int32_t K) {
    sais_rec<int32_t, int32_t>(s, SA, n, K);
}

// LCP[i] = lcp(suffix SA[i-1], suffix SA[i]), LCP[0] = 0 (Kasai et al. 2001).
template <class Sym, class Idx, class Out>
void lcp_kasai(const Sym* s, const Idx* SA, Out* LCP, Idx n) {
    std::vector<Idx> rank((size_t)n);
    for (Idx i = 0; i < n; ++i) rank[(size_t)SA[i]] = i;
    Idx h = 0;
    for (Idx i = 0; i < n; ++i) {
        Idx r = rank[(size_t)i];
        if (r == 0) { LCP[0] = 0; h = 0; continue; }
        Idx j = SA[r - 1];
        while (i + h < n && j + h < n && s[i + h] == s[j + h]) ++h;
        LCP[r] = (Out)h;
        if (h > 0) --h;
    }
}

}  // namespace oracle
