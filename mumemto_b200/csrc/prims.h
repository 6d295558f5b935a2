// mumemto_b200/csrc/prims.h -- device-wide sort / scan building blocks used by the suffix-array builder and the compaction steps.
#pragma once
#include "common.cuh"

namespace mmb {
namespace prim {

// Stable LSD radix sort of (key, value) pairs on key bits [begin_bit, end_bit).  in/out must not alias.
void sort_pairs(const u64* keys_in, u64* keys_out, const u32* vals_in, u32* vals_out, size_t n, int begin_bit, int end_bit, cudaStream_t s);
void sort_pairs64(const u64* keys_in, u64* keys_out, const u64* vals_in, u64* vals_out, size_t n, int begin_bit, int end_bit, cudaStream_t s);

// out[i] = sum of in[0..i)
void exclusive_sum(const u32* in, u32* out, size_t n, cudaStream_t s);
void exclusive_sum(const u32* in, u64* out, size_t n, cudaStream_t s);
void exclusive_sum(const u64* in, u64* out, size_t n, cudaStream_t s);
// out[i] = max of in[0..i]
void inclusive_max(const u32* in, u32* out, size_t n, cudaStream_t s);
void inclusive_max(const u64* in, u64* out, size_t n, cudaStream_t s);

}  // namespace prim
}  // namespace mmb
