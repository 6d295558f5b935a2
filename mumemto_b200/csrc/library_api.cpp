// mumemto_b200/csrc/library_api.cpp -- the reference's C library ABI (mumemto_library/mumemto.h:33-94) on the GPU engine.
//
// Same symbols, struct layouts, ownership and error codes as mumemto_library/mumemto_api.cpp:432-643: results are opaque heap
// objects owned by the caller until mum_free / mem_free, views point into them, return codes 0 ok / 1 null out / 2 null docs /
// 3 exception / 4 unknown, thread-local error text.  use_gsacak is accepted and ignored (one device route, same result).
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mumemto.h"
#include "mumemto_b200.h"

struct mumemto_mum_result {
    size_t num_docs = 0;
    std::vector<size_t> doc_record_offsets, record_lengths;
    std::vector<uint32_t> len; std::vector<int64_t> off; std::vector<uint8_t> strand;
};
struct mumemto_mem_result {
    size_t num_docs = 0;
    std::vector<size_t> doc_record_offsets, record_lengths;
    std::vector<uint32_t> len; std::vector<uint64_t> first; std::vector<int64_t> off; std::vector<size_t> ids; std::vector<uint8_t> strand;
};

namespace {
thread_local std::string g_last_error;

struct EngineGuard {
    mmb_engine* e = nullptr;
    EngineGuard() { if (mmb_create(0, nullptr, &e)) throw std::runtime_error(mmb_last_error()); }
    ~EngineGuard() { mmb_destroy(e); }
};

template <class R>
void load(EngineGuard& g, const mumemto_doc_view* docs, size_t num_docs, uint8_t revcomp, R* r) {
    std::vector<const char*> flat; std::vector<size_t> counts;
    r->num_docs = num_docs; r->doc_record_offsets.push_back(0);
    for (size_t d = 0; d < num_docs; ++d) {
        counts.push_back(docs[d].num_records);
        for (size_t i = 0; i < docs[d].num_records; ++i) {
            const char* s = docs[d].records ? docs[d].records[i] : nullptr;
            flat.push_back(s ? s : "");
            r->record_lengths.push_back(std::strlen(flat.back()));
        }
        r->doc_record_offsets.push_back(r->record_lengths.size());
    }
    if (mmb_load_records(g.e, flat.data(), counts.data(), num_docs, revcomp)) throw std::runtime_error(mmb_last_error());
}
}  // namespace

extern "C" {

const char* mumemto_last_error(void) { return g_last_error.c_str(); }

int mumemto_mum(const mumemto_doc_view* docs, size_t num_docs, uint32_t min_match_len, uint8_t use_revcomp, size_t num_distinct, uint8_t,
                mumemto_mum_result** out_result) {
    if (!out_result) { g_last_error = "out_result must be non-null"; return 1; }
    *out_result = nullptr;
    try {
        if (!docs && num_docs != 0) { g_last_error = "docs must be non-null when num_docs != 0"; return 2; }
        auto* r = new mumemto_mum_result();
        if (num_docs == 0) { r->doc_record_offsets.push_back(0); *out_result = r; return 0; }
        try {
            EngineGuard g;
            load(g, docs, num_docs, use_revcomp, r);
            mmb_options o; mmb_default_options(&o);
            o.min_match_len = min_match_len; o.num_distinct_docs = (int32_t)num_distinct; o.rare_freq = 1; o.max_mem_freq = 0; o.cli_normalise = 0; o.want_text = 0;
            if (mmb_find(g.e, &o)) throw std::runtime_error(mmb_last_error());
            uint64_t n = mmb_num_matches(g.e);
            r->len.resize(n); r->off.resize(n * num_docs); r->strand.resize(n * num_docs);
            if (mmb_copy_mums(g.e, r->len.data(), r->off.data(), r->strand.data())) throw std::runtime_error(mmb_last_error());
        } catch (...) { delete r; throw; }
        *out_result = r;
        return 0;
    } catch (const std::exception& e) { g_last_error = e.what(); return 3; }
    catch (...) { g_last_error = "unknown exception"; return 4; }
}

int mumemto_mem(const mumemto_doc_view* docs, size_t num_docs, uint32_t min_match_len, uint8_t use_revcomp, size_t num_distinct,
                size_t max_total_freq, size_t max_doc_freq, uint8_t, mumemto_mem_result** out_result) {
    if (!out_result) { g_last_error = "out_result must be non-null"; return 1; }
    *out_result = nullptr;
    try {
        if (!docs && num_docs != 0) { g_last_error = "docs must be non-null when num_docs != 0"; return 2; }
        auto* r = new mumemto_mem_result();
        if (num_docs == 0) { r->doc_record_offsets.push_back(0); *out_result = r; return 0; }
        try {
            if (max_doc_freq <= 1) throw std::invalid_argument("per-sequence MEM frequency f must be > 1 (use mumemto_mum instead)");
            EngineGuard g;
            load(g, docs, num_docs, use_revcomp, r);
            mmb_options o; mmb_default_options(&o);
            o.min_match_len = min_match_len; o.num_distinct_docs = (int32_t)num_distinct; o.rare_freq = (int32_t)max_doc_freq;
            o.max_mem_freq = (int32_t)max_total_freq; o.cli_normalise = 0; o.want_text = 0;
            if (mmb_find(g.e, &o)) throw std::runtime_error(mmb_last_error());
            uint64_t n = mmb_num_matches(g.e), occ = mmb_num_mem_occurrences(g.e);
            r->len.resize(n); r->first.resize(n + 1); r->off.resize(occ); r->strand.resize(occ);
            std::vector<uint64_t> ids(occ);
            if (mmb_copy_mems(g.e, r->len.data(), r->first.data(), r->off.data(), ids.data(), r->strand.data())) throw std::runtime_error(mmb_last_error());
            r->ids.assign(ids.begin(), ids.end());
        } catch (...) { delete r; throw; }
        *out_result = r;
        return 0;
    } catch (const std::exception& e) { g_last_error = e.what(); return 3; }
    catch (...) { g_last_error = "unknown exception"; return 4; }
}

size_t num_docs(const mumemto_mum_result* r) { return r ? r->num_docs : 0; }
const size_t* doc_record_offsets(const mumemto_mum_result* r) { return r ? r->doc_record_offsets.data() : nullptr; }
const size_t* record_lengths(const mumemto_mum_result* r) { return r ? r->record_lengths.data() : nullptr; }
size_t num_mums(const mumemto_mum_result* r) { return r ? r->len.size() : 0; }
mumemto_mum_match_view mum_at(const mumemto_mum_result* r, size_t idx) {
    mumemto_mum_match_view v{};
    if (!r || idx >= r->len.size()) return v;
    v.length = r->len[idx]; v.offsets = r->off.data() + idx * r->num_docs; v.strands = r->strand.data() + idx * r->num_docs;
    return v;
}
void mum_free(mumemto_mum_result* r) { delete r; }

size_t num_docs_mem(const mumemto_mem_result* r) { return r ? r->num_docs : 0; }
const size_t* doc_record_offsets_mem(const mumemto_mem_result* r) { return r ? r->doc_record_offsets.data() : nullptr; }
const size_t* record_lengths_mem(const mumemto_mem_result* r) { return r ? r->record_lengths.data() : nullptr; }
size_t num_mems(const mumemto_mem_result* r) { return r ? r->len.size() : 0; }
mumemto_mem_match_view mem_at(const mumemto_mem_result* r, size_t idx) {
    mumemto_mem_match_view v{};
    if (!r || idx >= r->len.size()) return v;
    size_t a = r->first[idx], b = r->first[idx + 1];
    v.length = r->len[idx]; v.occurrences = b - a; v.offsets = r->off.data() + a; v.seq_ids = r->ids.data() + a; v.strands = r->strand.data() + a;
    return v;
}
void mem_free(mumemto_mem_result* r) { delete r; }

}  // extern "C"
