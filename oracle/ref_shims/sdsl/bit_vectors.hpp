// oracle/ref_shims/sdsl/bit_vectors.hpp -- TEST INFRASTRUCTURE.
// Stand-in for sdsl::bit_vector, rank_support_v<1>, select_support (1-based select), as used by the reference at
// include/ref_builder.hpp:23-24, include/dictionary.hpp:49-51,115-124, include/parse.hpp:49-50,109-121.
// rank(i) = number of ones in [0,i); select(k) = position of the k-th one (k >= 1).  These are definitional.
#pragma once
#include "io.hpp"

namespace sdsl {

class bit_vector;

template <uint8_t B = 1>
class rank_support_v {
    const bit_vector* bv_ = nullptr;
    std::vector<uint64_t> cum_;   // ones before each 64-bit word
public:
    rank_support_v() = default;
    explicit rank_support_v(const bit_vector* bv);
    uint64_t operator()(uint64_t i) const;
    uint64_t rank(uint64_t i) const { return (*this)(i); }
    uint64_t serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
    void load(std::istream&, const bit_vector* = nullptr) {}
};

class select_support_shim {
    std::vector<uint64_t> pos_;   // position of every one
public:
    select_support_shim() = default;
    explicit select_support_shim(const bit_vector* bv);
    uint64_t operator()(uint64_t k) const { return pos_[k - 1]; }
    uint64_t select(uint64_t k) const { return pos_[k - 1]; }
    uint64_t serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
    void load(std::istream&, const bit_vector* = nullptr) {}
};

class bit_vector {
    std::vector<uint64_t> w_;
    uint64_t n_ = 0;
public:
    typedef rank_support_v<1> rank_1_type;
    typedef select_support_shim select_1_type;
    struct ref {
        uint64_t* word; uint64_t mask;
        operator bool() const { return (*word & mask) != 0; }
        ref& operator=(bool b) { if (b) *word |= mask; else *word &= ~mask; return *this; }
        ref& operator=(const ref& o) { return *this = (bool)o; }
    };
    bit_vector() = default;
    bit_vector(uint64_t n, int v) : w_((n + 63) / 64 + 1, v ? ~0ULL : 0ULL), n_(n) {}
    uint64_t size() const { return n_; }
    void resize(uint64_t n) { w_.resize((n + 63) / 64 + 1, 0); n_ = n; }
    ref operator[](uint64_t i) { return ref{&w_[i >> 6], 1ULL << (i & 63)}; }
    bool operator[](uint64_t i) const { return (w_[i >> 6] >> (i & 63)) & 1; }
    const std::vector<uint64_t>& words() const { return w_; }
    uint64_t serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
    void load(std::istream&) {}
};

template <uint8_t B>
rank_support_v<B>::rank_support_v(const bit_vector* bv) : bv_(bv) {
    const auto& w = bv->words();
    cum_.resize(w.size() + 1);
    uint64_t c = 0;
    for (size_t i = 0; i < w.size(); ++i) { cum_[i] = c; c += __builtin_popcountll(w[i]); }
    cum_[w.size()] = c;
}
template <uint8_t B>
uint64_t rank_support_v<B>::operator()(uint64_t i) const {
    const auto& w = bv_->words();
    uint64_t wi = i >> 6;
    if (wi >= w.size()) return cum_[w.size()];
    uint64_t r = i & 63;
    return cum_[wi] + (r ? __builtin_popcountll(w[wi] & ((1ULL << r) - 1)) : 0);
}
inline select_support_shim::select_support_shim(const bit_vector* bv) {
    for (uint64_t i = 0; i < bv->size(); ++i) if ((*bv)[i]) pos_.push_back(i);
}
}  // namespace sdsl
