// mumemto_b200/csrc/pybind_module.cpp -- `mumemto_b200._mumemto_core`: the reference's pybind11 module on the GPU engine.
//
// Same module surface as python_bindings/src/mumemto_pybind.cpp:93-172 (`mum`, `mem`, `MumResult`, `MemResult` with `num_docs`,
// `num_matches`, `__len__`, `__getitem__`, `match_at`; same argument names and defaults), so `from mumemto import mum, mem` users switch
// by changing the import.  `match_at` returns read-only NumPy arrays that VIEW the result object's memory (base = the Python result
// object; nothing is copied per match); the result object owns the C-ABI handle of include/mumemto.h and frees it when collected.
// `record_lengths()` (ours) gives lengths[doc][record].  Errors of the engine raise RuntimeError with mumemto_last_error().
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "mumemto.h"

namespace py = pybind11;

namespace {

template <typename T>
py::array view_1d(py::handle base, const T* ptr, size_t n, const char* dtype = nullptr) {
    py::dtype dt = dtype ? py::dtype(dtype) : py::dtype::of<T>();
    if (!ptr || n == 0) return py::array(dt, py::array::ShapeContainer{(py::ssize_t)0});
    py::array a(dt, {(py::ssize_t)n}, {(py::ssize_t)sizeof(T)}, ptr, base);
    a.attr("setflags")(false);
    return a;
}

struct DocViews {
    std::vector<std::vector<const char*>> ptrs; std::vector<mumemto_doc_view> views;
    explicit DocViews(const std::vector<std::vector<std::string>>& s) : ptrs(s.size()), views(s.size()) {
        for (size_t d = 0; d < s.size(); ++d) {
            for (const std::string& r : s[d]) ptrs[d].push_back(r.c_str());
            views[d] = mumemto_doc_view{ptrs[d].data(), ptrs[d].size()};
        }
    }
};

[[noreturn]] void fail(int rc) {
    const char* m = mumemto_last_error();
    throw std::runtime_error(m && *m ? std::string(m) : "mumemto: error " + std::to_string(rc));
}

py::list nested_lengths(size_t nd, const size_t* first, const size_t* len) {
    py::list out;
    for (size_t d = 0; d < nd; ++d) { py::list l; for (size_t i = first[d]; i < first[d + 1]; ++i) l.append(len[i]); out.append(l); }
    return out;
}

class MumResult {
public:
    explicit MumResult(mumemto_mum_result* r) : r_(r) {}
    MumResult(const MumResult&) = delete;
    ~MumResult() { if (r_) mum_free(r_); }
    size_t num_docs() const { return ::num_docs(r_); }
    size_t num_matches() const { return ::num_mums(r_); }
    py::list record_lengths() const { return nested_lengths(::num_docs(r_), ::doc_record_offsets(r_), ::record_lengths(r_)); }
    py::tuple match_at(size_t idx) const {
        if (idx >= num_matches()) throw py::index_error("MumResult index out of range");
        mumemto_mum_match_view m = ::mum_at(r_, idx);
        py::object self = py::cast(this);
        return py::make_tuple(m.length, view_1d<int64_t>(self, m.offsets, num_docs()), view_1d<uint8_t>(self, m.strands, num_docs(), "bool"));
    }
private:
    mumemto_mum_result* r_;
};

class MemResult {
public:
    explicit MemResult(mumemto_mem_result* r) : r_(r) {}
    MemResult(const MemResult&) = delete;
    ~MemResult() { if (r_) mem_free(r_); }
    size_t num_docs() const { return ::num_docs_mem(r_); }
    size_t num_matches() const { return ::num_mems(r_); }
    py::list record_lengths() const { return nested_lengths(::num_docs_mem(r_), ::doc_record_offsets_mem(r_), ::record_lengths_mem(r_)); }
    py::tuple match_at(size_t idx) const {
        if (idx >= num_matches()) throw py::index_error("MemResult index out of range");
        mumemto_mem_match_view m = ::mem_at(r_, idx);
        py::object self = py::cast(this);
        return py::make_tuple(m.length, view_1d<int64_t>(self, m.offsets, m.occurrences), view_1d<size_t>(self, m.seq_ids, m.occurrences),
                              view_1d<uint8_t>(self, m.strands, m.occurrences, "bool"));
    }
private:
    mumemto_mem_result* r_;
};

MumResult* mum_run(const std::vector<std::vector<std::string>>& sequences, uint32_t min_match_len, bool use_revcomp, size_t num_distinct, bool use_gsacak) {
    DocViews v(sequences);
    mumemto_mum_result* r = nullptr;
    int rc;
    { py::gil_scoped_release nogil; rc = mumemto_mum(v.views.data(), v.views.size(), min_match_len, use_revcomp, num_distinct, use_gsacak, &r); }
    if (rc != 0) fail(rc);
    return new MumResult(r);
}

MemResult* mem_run(const std::vector<std::vector<std::string>>& sequences, uint32_t min_match_len, bool use_revcomp, size_t num_distinct,
                   size_t max_total_freq, size_t max_doc_freq, bool use_gsacak) {
    DocViews v(sequences);
    mumemto_mem_result* r = nullptr;
    int rc;
    { py::gil_scoped_release nogil; rc = mumemto_mem(v.views.data(), v.views.size(), min_match_len, use_revcomp, num_distinct, max_total_freq, max_doc_freq, use_gsacak, &r); }
    if (rc != 0) fail(rc);
    return new MemResult(r);
}

template <class R>
py::tuple get_item(const R& self, py::ssize_t i, const char* what) {
    const py::ssize_t n = (py::ssize_t)self.num_matches();
    if (i < 0) i += n;
    if (i < 0 || i >= n) throw py::index_error(std::string(what) + " index out of range");
    return self.match_at((size_t)i);
}

}  // namespace

PYBIND11_MODULE(_mumemto_core, m) {
    m.doc() = "mumemto Python bindings on the B200 engine (backed by libmumemto_b200.so)";
    py::class_<MumResult>(m, "MumResult")
        .def("num_docs", &MumResult::num_docs).def("num_matches", &MumResult::num_matches).def("__len__", &MumResult::num_matches)
        .def("record_lengths", &MumResult::record_lengths)
        .def("__getitem__", [](const MumResult& s, py::ssize_t i) { return get_item(s, i, "MumResult"); }, py::arg("i"))
        .def("match_at", &MumResult::match_at, py::arg("idx"), "Return (length, offsets_ndarray, strands_ndarray) as zero-copy views.");
    py::class_<MemResult>(m, "MemResult")
        .def("num_docs", &MemResult::num_docs).def("num_matches", &MemResult::num_matches).def("__len__", &MemResult::num_matches)
        .def("record_lengths", &MemResult::record_lengths)
        .def("__getitem__", [](const MemResult& s, py::ssize_t i) { return get_item(s, i, "MemResult"); }, py::arg("i"))
        .def("match_at", &MemResult::match_at, py::arg("idx"), "Return (length, offsets_ndarray, seq_ids_ndarray, strands_ndarray) as zero-copy views.");
    m.def("mum", &mum_run, py::return_value_policy::take_ownership, py::arg("sequences"), py::arg("min_match_len") = 20, py::arg("use_revcomp") = true,
          py::arg("num_distinct") = (size_t)0, py::arg("use_gsacak") = false);
    m.def("mem", &mem_run, py::return_value_policy::take_ownership, py::arg("sequences"), py::arg("min_match_len") = 20, py::arg("use_revcomp") = true,
          py::arg("num_distinct") = (size_t)0, py::arg("max_total_freq") = (size_t)0, py::arg("max_doc_freq") = (size_t)2, py::arg("use_gsacak") = false);
}
