// oracle/ref_shims/sdsl/rmq_support.hpp -- TEST INFRASTRUCTURE.
// Stand-in for sdsl::rmq_succinct_sct<> (reference: include/dictionary.hpp:48,98,154; include/pfp.hpp:54,243;
// include/pfp_lcp_mum.hpp:290-292).  operator()(l, r) returns the index of a minimum of v[l..r] (inclusive, leftmost);
// the reference only ever reads the VALUE at that index, so any correct range-minimum answers identically.
// Layout: a copy of the values, block minima over 16-element blocks and a sparse table over the blocks.
#pragma once
#include "int_vector.hpp"

namespace sdsl {
template <bool MIN = true>
class rmq_succinct_sct {
    static constexpr uint64_t BS = 16;
    std::vector<int64_t> val_;
    std::vector<std::vector<uint32_t>> st_;   // st_[k][b] = index (in val_) of the min of blocks [b, b + 2^k)
    uint64_t scan(uint64_t l, uint64_t r) const {
        uint64_t m = l;
        for (uint64_t i = l + 1; i <= r; ++i) if (val_[i] < val_[m]) m = i;
        return m;
    }
    uint64_t better(uint64_t a, uint64_t b) const {   // leftmost minimum
        if (val_[b] < val_[a]) return b;
        if (val_[a] < val_[b]) return a;
        return a < b ? a : b;
    }
public:
    rmq_succinct_sct() = default;
    template <class Vec>
    explicit rmq_succinct_sct(const Vec* v) {
        if (!v) return;
        val_.assign(v->begin(), v->end());
        uint64_t n = val_.size();
        if (n == 0) return;
        uint64_t nb = (n + BS - 1) / BS;
        st_.emplace_back(nb);
        for (uint64_t b = 0; b < nb; ++b) st_[0][b] = (uint32_t)scan(b * BS, std::min(n, (b + 1) * BS) - 1);
        for (uint64_t k = 1; (1ULL << k) <= nb; ++k) {
            st_.emplace_back(nb - (1ULL << k) + 1);
            for (uint64_t b = 0; b + (1ULL << k) <= nb; ++b)
                st_[k][b] = (uint32_t)better(st_[k - 1][b], st_[k - 1][b + (1ULL << (k - 1))]);
        }
    }
    uint64_t operator()(uint64_t l, uint64_t r) const {
        uint64_t bl = l / BS, br = r / BS;
        if (bl == br) return scan(l, r);
        uint64_t m = better(scan(l, (bl + 1) * BS - 1), scan(br * BS, r));
        if (bl + 1 < br) {
            uint64_t len = br - bl - 1;
            uint64_t k = 63 - __builtin_clzll(len);
            uint64_t c = better(st_[k][bl + 1], st_[k][br - (1ULL << k)]);
            // keep the leftmost minimum overall
            if (val_[c] < val_[m] || (val_[c] == val_[m] && c < m)) m = c;
        }
        return m;
    }
    uint64_t serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
    void load(std::istream&) {}
};
}  // namespace sdsl
