// oracle/ref_shims/sdsl/int_vector.hpp -- TEST INFRASTRUCTURE.
// Stand-in for sdsl::int_vector<40> (reference: include/parse.hpp:45-48): a resizable array of unsigned integers.
#pragma once
#include "bit_vectors.hpp"

namespace sdsl {
template <uint8_t W = 0>
class int_vector {
    std::vector<uint64_t> v_;
public:
    int_vector() = default;
    explicit int_vector(uint64_t n, uint64_t x = 0) : v_(n, x) {}
    void resize(uint64_t n) { v_.resize(n); }
    uint64_t size() const { return v_.size(); }
    uint64_t& operator[](uint64_t i) { return v_[i]; }
    const uint64_t& operator[](uint64_t i) const { return v_[i]; }
    void swap(int_vector& o) { v_.swap(o.v_); }
    void clear() { v_.clear(); }
    uint64_t serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
    void load(std::istream&) {}
};
}  // namespace sdsl
