// mumemto_b200/csrc/pfp_device.cuh -- device helpers shared by pfp.cu and stream.cu.
#pragma once
#include "common.cuh"

namespace mmb {

// min of v[l..r] (inclusive, l <= r) with the two-level index of PfpIndex: rmq[k][b] = min of blocks [b, b + 2^k), 32 entries per block
__device__ __forceinline__ u32 range_min(const u32* __restrict__ v, const u32* __restrict__ rmq, u32 nblocks, u32 l, u32 r) {
    u32 bl = l >> 5, br = r >> 5, mn = 0xffffffffu;
    if (r - l < 48) { for (u32 i = l; i <= r; ++i) mn = min(mn, v[i]); return mn; }      // short ranges (the common case): plain scan
    for (u32 i = l; i < (bl + 1) * 32; ++i) mn = min(mn, v[i]);
    for (u32 i = br * 32; i <= r; ++i) mn = min(mn, v[i]);
    if (bl + 1 < br) {
        u32 len = br - bl - 1, k = 31 - (u32)__clz((int)len);
        mn = min(mn, min(rmq[(u64)k * nblocks + bl + 1], rmq[(u64)k * nblocks + br - (1u << k)]));
    }
    return mn;
}

}  // namespace mmb
