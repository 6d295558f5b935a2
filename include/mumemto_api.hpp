// include/mumemto_api.hpp -- native C++ API of the GPU engine, header-only over the C ABI of include/mumemto.h.
//
// Same namespace, function names, argument order/defaults and result shapes as the reference's mumemto_library/mumemto_api.hpp:29-57
// (`mumemto::mumemto_mum / mumemto_mem` over `std::vector<std::vector<std::string>>`, results `{matches, lengths}` with
// `mumsio::Mum {length, offsets, strands}` / `mumsio::Mem {length, offsets, seq_ids, strands}`, include/mumsio.hpp:17-28), so a
// program written against that header compiles against this one and links libmumemto_b200.so.  Errors throw std::runtime_error with
// the text of mumemto_last_error() (the reference throws from the same call).  There is no CPU path behind it.
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "mumemto.h"

namespace mumsio {
struct Mum { uint32_t length; std::vector<int64_t> offsets; std::vector<uint8_t> strands; };                                // -1 = absent, 1 = '+'
struct Mem { uint32_t length; std::vector<int64_t> offsets; std::vector<size_t> seq_ids; std::vector<uint8_t> strands; };
}  // namespace mumsio

namespace mumemto {

struct MumResult { std::vector<mumsio::Mum> matches; std::vector<std::vector<size_t>> lengths; };   // lengths[doc][record]
struct MemResult { std::vector<mumsio::Mem> matches; std::vector<std::vector<size_t>> lengths; };

namespace detail {
struct DocViews {
    std::vector<std::vector<const char*>> ptrs; std::vector<mumemto_doc_view> views;
    explicit DocViews(const std::vector<std::vector<std::string>>& s) : ptrs(s.size()), views(s.size()) {
        for (size_t d = 0; d < s.size(); ++d) {
            for (const std::string& r : s[d]) ptrs[d].push_back(r.c_str());
            views[d] = mumemto_doc_view{ptrs[d].data(), ptrs[d].size()};
        }
    }
};
inline std::vector<std::vector<size_t>> nested(size_t nd, const size_t* first, const size_t* len) {
    std::vector<std::vector<size_t>> out(nd);
    for (size_t d = 0; d < nd; ++d) out[d].assign(len + first[d], len + first[d + 1]);
    return out;
}
[[noreturn]] inline void fail(int rc) { const char* m = mumemto_last_error(); throw std::runtime_error(m && *m ? m : "mumemto: error " + std::to_string(rc)); }
}  // namespace detail

inline MumResult mumemto_mum(const std::vector<std::vector<std::string>>& sequences, std::uint32_t min_match_len = 20, bool use_revcomp = true,
                             size_t num_distinct = 0, bool use_gsacak = false) {
    detail::DocViews v(sequences);
    mumemto_mum_result* r = nullptr;
    int rc = ::mumemto_mum(v.views.data(), v.views.size(), min_match_len, use_revcomp, num_distinct, use_gsacak, &r);
    if (rc != 0) detail::fail(rc);
    MumResult out;
    size_t nd = ::num_docs(r), n = ::num_mums(r);
    out.lengths = detail::nested(nd, ::doc_record_offsets(r), ::record_lengths(r));
    out.matches.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        mumemto_mum_match_view m = ::mum_at(r, i);
        out.matches.push_back({m.length, std::vector<int64_t>(m.offsets, m.offsets + nd), std::vector<uint8_t>(m.strands, m.strands + nd)});
    }
    ::mum_free(r);
    return out;
}

inline MemResult mumemto_mem(const std::vector<std::vector<std::string>>& sequences, std::uint32_t min_match_len = 20, bool use_revcomp = true,
                             size_t num_distinct = 0, size_t max_total_freq = 0, size_t max_doc_freq = 2, bool use_gsacak = false) {
    detail::DocViews v(sequences);
    mumemto_mem_result* r = nullptr;
    int rc = ::mumemto_mem(v.views.data(), v.views.size(), min_match_len, use_revcomp, num_distinct, max_total_freq, max_doc_freq, use_gsacak, &r);
    if (rc != 0) detail::fail(rc);
    MemResult out;
    size_t nd = ::num_docs_mem(r), n = ::num_mems(r);
    out.lengths = detail::nested(nd, ::doc_record_offsets_mem(r), ::record_lengths_mem(r));
    out.matches.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        mumemto_mem_match_view m = ::mem_at(r, i);
        out.matches.push_back({m.length, std::vector<int64_t>(m.offsets, m.offsets + m.occurrences), std::vector<size_t>(m.seq_ids, m.seq_ids + m.occurrences),
                               std::vector<uint8_t>(m.strands, m.strands + m.occurrences)});
    }
    ::mem_free(r);
    return out;
}

}  // namespace mumemto
