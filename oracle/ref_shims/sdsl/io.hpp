// oracle/ref_shims/sdsl/io.hpp -- TEST INFRASTRUCTURE.
// Minimal stand-in for the parts of sdsl-lite (github.com/vikshiv/sdsl-lite, un-vendored and unpinned in
// /root/reference/thirdparty/CMakeLists.txt:117-130) that the reference sources name.  Serialisation is a
// no-op: the oracle never writes or reads sdsl index files.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <iomanip>
#include <iostream>
#include <string>
#include <unordered_map>
#include <vector>

namespace sdsl {
struct structure_tree_node {};
struct structure_tree {
    static structure_tree_node* add_child(structure_tree_node*, const std::string&, const std::string&) { return nullptr; }
    static void add_size(structure_tree_node*, uint64_t) {}
};
namespace util {
template <class T> std::string class_name(const T&) { return "shim"; }
}
namespace conf { const uint64_t SDSL_BLOCK_SIZE = (uint64_t)1 << 22; }

template <class T> uint64_t serialize(const T&, std::ostream&, structure_tree_node* = nullptr, std::string = "") { return 0; }
template <class T> void load(T&, std::istream&) {}
template <class T> uint64_t write_member(const T&, std::ostream&, structure_tree_node* = nullptr, std::string = "") { return 0; }
template <class T> void read_member(T&, std::istream&) {}
template <class T> uint64_t size_in_bytes(const T&) { return 0; }
}  // namespace sdsl
