// mumemto_b200/csrc/sa.cu -- suffix sorting by prefix doubling with group-local sorts.
//
// Replaces the reference's gsacak() / sacak_int() calls (include/direct_gsacak.hpp:56-67, include/dictionary.hpp:133,
// include/parse.hpp:85; third-party gsa-is) -- rows C/D of SURVEY.md section 8.  Three users share one refinement engine:
//   build_suffix_array      the text itself (direct route): ranks 0..n over T followed by a smallest sentinel, rank 0 = the sentinel
//   (pfp.cu)                the dictionary (suffixes end at their phrase's terminator and may TIE) and the parse (integer alphabet)
//
// State in HBM (m suffixes, 32-bit indices): sa[m] current order, isa[m] = rank of the first member of the suffix's group
// ("head rank"), nh[m] scratch for the new head ranks of one round, and a list of the groups that still have >1 member,
// each packed as head << 32 | size.
// Round h (all suffixes already ordered by their first h symbols): every listed group is sorted by key = isa[suffix + h] + 1
// (0 = nothing left to compare) by a kernel chosen by group size:
//   size <= 256   sub-warp / warp rank sort in registers (4/8/16/32 lanes per group, up to 8 elements per lane): each element
//                 counts the smaller and the equal keys of its group through shuffles, which gives its slot AND its new head
//                 rank in one go; the first element of each run of equal keys appends the run as a group of the next round
//   size <= 4096  one block, bitonic sort in shared memory
//   larger        flat list + device-wide radix sort (prims.h), as in round 1's first version
// New head ranks are written to nh[] and applied to isa[] by a second kernel, so that no group reads a rank of this round.
// Traffic per listed suffix per round: sa 4 B read + 4 B written, isa one random 4 B read and one random 4 B write, nh 4+4 B.
#include "sa.h"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 TPB = 256;
constexpr u32 FULL = 0xffffffffu;
constexpr u32 BLOCK_SORT_MAX = 2048;      // elements one block sorts in shared memory
constexpr int NCLASS = 9;                 // size classes: <=4, <=8, <=16, <=32, <=64, <=128, <=256, <=BLOCK_SORT_MAX, larger

typedef u64 Group;                        // head << 32 | size
__device__ __forceinline__ u32 g_head(Group g) { return (u32)(g >> 32); }
__device__ __forceinline__ u32 g_size(Group g) { return (u32)g; }
__device__ __forceinline__ int size_class(u32 size) {
    if (size <= 32) return size <= 4 ? 0 : (size <= 8 ? 1 : (size <= 16 ? 2 : 3));
    if (size <= 256) return size <= 64 ? 4 : (size <= 128 ? 5 : 6);
    return size <= BLOCK_SORT_MAX ? 7 : 8;
}

// Head ranks by suffix.  stride 1: plain array.  stride 2 (dictionary mode): pairs (head rank, offset of the position inside its
// phrase) -- the doubling key of suffix s at offset h then needs ONE random read, at s + h: the suffix is exhausted (nothing left to
// compare, key 0) exactly when that position lies less than h symbols into its phrase, i.e. the phrase started after s.
struct Isa {
    u32* p; u32 stride;
    __device__ __forceinline__ void set(u32 s, u32 rank) const { p[(u64)s * stride] = rank; }
};

// Next round's groups, one list per size class (so that no binning pass is needed).
struct GroupLists {
    Group* list[NCLASS];
    u32* count;                 // NCLASS counters
};

// Append groups from converged code: EVERY thread of the block must call this (want = false when it has nothing).  Lanes are matched
// by class inside the warp, one lane per (warp, class) reserves in a shared-memory counter and one thread per (block, class) in the
// global one -- thousands of warps adding to the same global counter serialise in L2 otherwise (the lists hold ~1e8 groups per round).
__device__ __forceinline__ void append_group_warp(bool want, u32 head, u32 size, const GroupLists& L) {
    __shared__ u32 s_cnt[NCLASS], s_base[NCLASS];
    if (threadIdx.x < NCLASS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned wanting = __ballot_sync(FULL, want);
    u32 lane = threadIdx.x & 31;
    int cls = 0; unsigned peers = 0; u32 wbase = 0;
    if (want) {
        cls = size_class(size);
        peers = __match_any_sync(wanting, cls);
        int leader = __ffs((int)peers) - 1;
        if ((int)lane == leader) wbase = atomicAdd(&s_cnt[cls], (u32)__popc(peers));
        wbase = __shfl_sync(peers, wbase, leader);
    }
    __syncthreads();
    if (threadIdx.x < NCLASS && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&L.count[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (want) L.list[cls][s_base[cls] + wbase + __popc(peers & ((1u << lane) - 1))] = ((u64)head << 32) | size;
}
// Same from divergent code (few callers per block).
__device__ __forceinline__ void append_group_single(u32 head, u32 size, const GroupLists& L) {
    int cls = size_class(size);
    L.list[cls][atomicAdd(&L.count[cls], 1u)] = ((u64)head << 32) | size;
}

// ------------------------------------------------------------------------------------------------ initial order
__global__ void __launch_bounds__(TPB) symbol_histogram(const u8* __restrict__ text, u64 n, u32* __restrict__ hist) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&h[text[i]], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(TPB) pack_text_keys(const u8* __restrict__ text, u64 n, const u8* __restrict__ lut, u32 bits, u32 cpk,
                                                      u64* __restrict__ keys, u32* __restrict__ vals) {
    __shared__ u8 code[256];
    code[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    u64 key = 0;
    for (u32 t = 0; t < cpk; ++t) {
        u64 p = i + t;
        u64 c = (p < n) ? code[text[p]] : 0;
        key = (key << bits) | c;
    }
    keys[i] = key;
    vals[i] = (u32)i;
}

// group starts / head-rank candidates after the initial sort
__global__ void __launch_bounds__(TPB) first_flags(const u64* __restrict__ keys, u32 m, u32* __restrict__ head_cand) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    head_cand[p] = (p == 0 || keys[p - 1] != keys[p]) ? p : 0u;
}

// isa from head ranks; the last member of every multi-member group appends the group (limit: see refine_by_doubling)
__global__ void __launch_bounds__(TPB) first_ranks(const u64* __restrict__ keys, const u32* __restrict__ sa, const u32* __restrict__ head, u32 m, u32 h0,
                                                   const u32* __restrict__ limit, Isa isa, GroupLists L) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false; u32 hd = 0;
    if (p < m) {
        hd = head[p];
        isa.set(sa[p], hd);
        bool last = (p + 1 == m) || keys[p + 1] != keys[p];
        want = last && hd != p;
        if (want && limit && limit[sa[hd]] <= h0) want = false;   // the whole (terminated) suffix is inside the key: equal strings, final tie
    }
    append_group_warp(want, hd, p - hd + 1, L);
}

// key of suffix s for the round with offset h: rank of the suffix h symbols further on (+1), 0 when there is nothing further
__device__ __forceinline__ u32 round_key(u32 s, u32 h, u32 m, const Isa& isa, const u32* __restrict__ limit) {
    u64 t = (u64)s + h;
    if (isa.stride == 2) {
        if (t >= m) return 0u;                              // past the last terminator
        uint2 v = *reinterpret_cast<const uint2*>(isa.p + 2 * t);
        return v.y < h ? 0u : v.x + 1u;
    }
    if (limit && limit[s] <= h) return 0u;
    return t < m ? isa.p[t] + 1u : 0u;
}

// ------------------------------------------------------------------------------------------------ (sub-)warp rank sort
// WIDTH lanes per group, K elements per lane (K > 1 only with WIDTH = 32).
template <int WIDTH, int K>
__global__ void __launch_bounds__(TPB) sort_groups_warp(const Group* __restrict__ groups, u32 ngroups, u32 h, u32 m, u32* __restrict__ sa,
                                                        Isa isa, const u32* __restrict__ limit, u32* __restrict__ nh,
                                                        u8* __restrict__ changed, GroupLists next) {
    u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
    u32 gi = tid / WIDTH, lane = tid % WIDTH;
    bool have = gi < ngroups;
    Group g = have ? groups[gi] : 0;
    u32 head = g_head(g), size = have ? g_size(g) : 0u;
    u32 s[K], key[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        u32 e = lane + (u32)WIDTH * k;
        s[k] = 0; key[k] = 0;
        if (e < size) { s[k] = sa[head + e]; key[k] = round_key(s[k], h, m, isa, limit); }
    }
    u32 lt[K], eq[K], before[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { lt[k] = 0; eq[k] = 0; before[k] = 0; }
#pragma unroll
    for (int k2 = 0; k2 < K; ++k2) {
        for (int j = 0; j < WIDTH; ++j) {
            u32 kj = __shfl_sync(FULL, key[k2], j, WIDTH);
            u32 ej = (u32)j + (u32)WIDTH * k2;
            if (ej < size) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    u32 e = lane + (u32)WIDTH * k;
                    lt[k] += kj < key[k];
                    eq[k] += kj == key[k];
                    before[k] += (kj == key[k]) && (ej < e);
                }
            }
        }
    }
    // A group whose keys are all equal did not split: nothing moves, no rank changes, and it is carried over as it is.
    bool split = (size > 0) && (eq[0] != size);           // element 0 exists in every listed group; uniform over the group's lanes
    if (lane == 0 && have) changed[gi] = split ? 1 : 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        u32 e = lane + (u32)WIDTH * k;
        bool valid = e < size;
        if (valid && split) {
            u32 slot = head + lt[k] + before[k];
            sa[slot] = s[k];
            nh[slot] = head + lt[k];
        }
        append_group_warp(valid && before[k] == 0 && eq[k] > 1 && key[k] != 0, head + lt[k], eq[k], next);
    }
}

// ------------------------------------------------------------------------------------------------ block bitonic sort
__global__ void __launch_bounds__(TPB) sort_groups_block(const Group* __restrict__ groups, u32 h, u32 m, u32* __restrict__ sa, Isa isa,
                                                         const u32* __restrict__ limit, u32* __restrict__ nh, GroupLists next) {
    __shared__ u64 item[BLOCK_SORT_MAX];          // key << 32 | suffix
    __shared__ u16 start_of[BLOCK_SORT_MAX];      // 1 + index of the first element of the run each element belongs to (<= 4096)
    Group g = groups[blockIdx.x];
    u32 head = g_head(g), size = g_size(g);
    u32 n2 = 1; while (n2 < size) n2 <<= 1;
    for (u32 e = threadIdx.x; e < n2; e += blockDim.x) {
        u64 v = ~0ull;
        if (e < size) { u32 s = sa[head + e]; v = ((u64)round_key(s, h, m, isa, limit) << 32) | s; }
        item[e] = v;
    }
    __syncthreads();
    for (u32 k = 2; k <= n2; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {     // one thread per compare-exchange pair
                u32 e = ((t & ~(j - 1)) << 1) | (t & (j - 1)), partner = e | j;
                u64 a = item[e], b = item[partner];
                bool up = (e & k) == 0;
                if ((a > b) == up) { item[e] = b; item[partner] = a; }
            }
            __syncthreads();
        }
    }
    // run starts -> for every element the index of the first element of its run (block-wide inclusive max-scan, stored +1)
    __shared__ u32 slice_max[TPB];
    u32 per = (size + blockDim.x - 1) / blockDim.x;
    u32 lo = min(threadIdx.x * per, size), hi = min(lo + per, size);
    u32 run = 0;
    for (u32 e = lo; e < hi; ++e) {
        if (e == 0 || (u32)(item[e] >> 32) != (u32)(item[e - 1] >> 32)) run = e + 1;
        start_of[e] = (u16)run;
    }
    slice_max[threadIdx.x] = run;
    __syncthreads();
    for (u32 d = 1; d < blockDim.x; d <<= 1) {
        u32 v = (threadIdx.x >= d) ? slice_max[threadIdx.x - d] : 0u;
        __syncthreads();
        slice_max[threadIdx.x] = max(slice_max[threadIdx.x], v);
        __syncthreads();
    }
    u32 carry = threadIdx.x ? slice_max[threadIdx.x - 1] : 0u;
    for (u32 e = lo; e < hi; ++e) {
        u32 st = max((u32)start_of[e], carry) - 1;
        u32 key = (u32)(item[e] >> 32);
        sa[head + e] = (u32)item[e];
        nh[head + e] = head + st;
        bool last = (e + 1 == size) || (u32)(item[e + 1] >> 32) != key;
        if (last && e != st && key != 0) append_group_single(head + st, e - st + 1, next);
    }
}

// ------------------------------------------------------------------------------------------------ applying new head ranks
template <int WIDTH, int K>
__global__ void __launch_bounds__(TPB) apply_groups_warp(const Group* __restrict__ groups, u32 ngroups, const u32* __restrict__ sa,
                                                         const u32* __restrict__ nh, const u8* __restrict__ changed, Isa isa) {
    u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
    u32 gi = tid / WIDTH, lane = tid % WIDTH;
    if (gi >= ngroups || !changed[gi]) return;
    Group g = groups[gi];
    u32 head = g_head(g), size = g_size(g);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        u32 e = lane + (u32)WIDTH * k;
        if (e < size) { u32 r = nh[head + e]; if (r != head) isa.set(sa[head + e], r); }      // the first run keeps the old head rank
    }
}
__global__ void __launch_bounds__(TPB) apply_groups_block(const Group* __restrict__ groups, const u32* __restrict__ sa, const u32* __restrict__ nh,
                                                          Isa isa) {
    Group g = groups[blockIdx.x];
    u32 head = g_head(g), size = g_size(g);
    for (u32 e = threadIdx.x; e < size; e += blockDim.x) isa.set(sa[head + e], nh[head + e]);
}

// ------------------------------------------------------------------------------------------------ large groups: flat list + radix sort
__global__ void large_sizes(const Group* __restrict__ groups, u32 n, u32* __restrict__ sizes) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sizes[i] = g_size(groups[i]);
    if (i == n) sizes[i] = 0;
}
__global__ void __launch_bounds__(TPB) large_gather(const Group* __restrict__ groups, const u32* __restrict__ base, u32 h, u32 m,
                                                    const u32* __restrict__ sa, Isa isa, const u32* __restrict__ limit,
                                                    u64* __restrict__ keys, u32* __restrict__ vals, u32* __restrict__ pos) {
    Group g = groups[blockIdx.x];
    u32 head = g_head(g), size = g_size(g), b = base[blockIdx.x];
    for (u32 e = blockIdx.y * blockDim.x + threadIdx.x; e < size; e += gridDim.y * blockDim.x) {
        u32 s = sa[head + e];
        keys[b + e] = ((u64)head << 32) | round_key(s, h, m, isa, limit);
        vals[b + e] = s;
        pos[b + e] = head + e;
    }
}
__global__ void __launch_bounds__(TPB) large_scatter(const u32* __restrict__ pos, const u64* __restrict__ keys, const u32* __restrict__ vals, u32 A,
                                                     u32* __restrict__ sa, u32* __restrict__ head_cand) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    sa[pos[a]] = vals[a];
    head_cand[a] = (a == 0 || keys[a - 1] != keys[a]) ? pos[a] : 0u;
}
__global__ void __launch_bounds__(TPB) large_finish(const u32* __restrict__ pos, const u64* __restrict__ keys, const u32* __restrict__ head, u32 A,
                                                    u32* __restrict__ nh, GroupLists next) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false; u32 hd = 0, p = 0;
    if (a < A) {
        hd = head[a]; p = pos[a];
        nh[p] = hd;
        bool last = (a + 1 == A) || keys[a + 1] != keys[a];
        want = last && hd != p && (u32)keys[a] != 0;
    }
    append_group_warp(want, hd, p - hd + 1, next);
}
__global__ void __launch_bounds__(TPB) large_apply(const u32* __restrict__ pos, u32 A, const u32* __restrict__ sa, const u32* __restrict__ nh,
                                                   Isa isa) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < A) isa.set(sa[pos[a]], nh[pos[a]]);
}

u32 bits_for(u64 x) { u32 b = 1; while ((1ull << b) <= x && b < 63) ++b; return b; }   // bits to represent values 0..x

template <int WIDTH, int K>
void launch_warp_class(const Group* groups, u32 n, u32 h, u32 m, u32* sa, Isa isa, const u32* limit, u32* nh, u8* changed, const GroupLists& next,
                       cudaStream_t s) {
    if (!n) return;
    auto sort_groups_warp_k = sort_groups_warp<WIDTH, K>;
    MMB_LAUNCH(sort_groups_warp_k, ceil_div((u64)n * WIDTH, TPB), TPB, 0, s, groups, n, h, m, sa, isa, limit, nh, changed, next);
}
template <int WIDTH, int K>
void launch_apply_class(const Group* groups, u32 n, const u32* sa, const u32* nh, const u8* changed, Isa isa, cudaStream_t s) {
    if (!n) return;
    auto apply_groups_warp_k = apply_groups_warp<WIDTH, K>;
    MMB_LAUNCH(apply_groups_warp_k, ceil_div((u64)n * WIDTH, TPB), TPB, 0, s, groups, n, sa, nh, changed, isa);
}

// ================================================================================================ dictionary refinement by symbol chunks
// The pruned prefix-free-parse route (pfp.cu) only needs the exact order of the dictionary suffixes inside the few key groups whose
// text suffixes can still take part in a match (or that hold a whole phrase).  Those groups are refined WITHOUT a rank array: round r
// compares the next `cpk` symbols of every listed suffix, packed from the dictionary bytes themselves (one or two random sectors per
// listed suffix and round, nothing written back but the permuted sa entries).  A suffix whose string has ended sorts with key 0 and ties
// with every other ended suffix of its group (equal strings: a member group of the stream).
// Output: sa reordered inside the listed groups; ghead[rank] = first rank of the run of equal strings the rank belongs to.
struct DictView {
    const u8* dict; const u32* limit; const u8* lut; u32 bits, cpk;
};

__device__ __forceinline__ u64 chunk_key(const DictView& D, const u8* code, u32 s, u32 h) {
    u32 lim = D.limit[s];                            // symbols from s up to and including the terminator
    if (lim <= h) return 0ull;
    u32 take = min(D.cpk, lim - h);
    const u8* q = D.dict + (u64)s + h;
    u64 key = 0;
    for (u32 t = 0; t < D.cpk; ++t) key = (key << D.bits) | (u64)(t < take ? code[q[t]] : 0);
    return key;
}

template <int WIDTH, int K>
__global__ void __launch_bounds__(TPB) sort_dgroups_warp(const Group* __restrict__ groups, u32 ngroups, u32 h, DictView D, u32* __restrict__ sa,
                                                         u32* __restrict__ ghead, GroupLists next) {
    __shared__ u8 code[256];
    code[threadIdx.x] = D.lut[threadIdx.x];
    __syncthreads();
    u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
    u32 gi = tid / WIDTH, lane = tid % WIDTH;
    bool have = gi < ngroups;
    Group g = have ? groups[gi] : 0;
    u32 head = g_head(g), size = have ? g_size(g) : 0u;
    u32 s[K]; u64 key[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        u32 e = lane + (u32)WIDTH * k;
        s[k] = 0; key[k] = 0;
        if (e < size) { s[k] = sa[head + e]; key[k] = chunk_key(D, code, s[k], h); }
    }
    u32 lt[K], eq[K], before[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { lt[k] = 0; eq[k] = 0; before[k] = 0; }
#pragma unroll
    for (int k2 = 0; k2 < K; ++k2) {
        for (int j = 0; j < WIDTH; ++j) {
            u64 kj = __shfl_sync(FULL, key[k2], j, WIDTH);
            u32 ej = (u32)j + (u32)WIDTH * k2;
            if (ej < size) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    u32 e = lane + (u32)WIDTH * k;
                    lt[k] += kj < key[k];
                    eq[k] += kj == key[k];
                    before[k] += (kj == key[k]) && (ej < e);
                }
            }
        }
    }
    bool split = (size > 0) && (eq[0] != size);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        u32 e = lane + (u32)WIDTH * k;
        bool valid = e < size;
        u32 run = head + lt[k], slot = run + before[k];
        if (valid) {
            if (split) sa[slot] = s[k];
            if (eq[k] == 1 || key[k] == 0) ghead[slot] = run;          // final: alone, or tied with equal strings
        }
        append_group_warp(valid && before[k] == 0 && eq[k] > 1 && key[k] != 0, run, eq[k], next);
    }
}

__global__ void __launch_bounds__(TPB) sort_dgroups_block(const Group* __restrict__ groups, u32 h, DictView D, u32* __restrict__ sa,
                                                          u32* __restrict__ ghead, GroupLists next) {
    __shared__ u64 key[BLOCK_SORT_MAX];
    __shared__ u32 suf[BLOCK_SORT_MAX];
    __shared__ u16 idx[BLOCK_SORT_MAX];           // element index carried through the sort
    __shared__ u16 start_of[BLOCK_SORT_MAX];
    __shared__ u8 code[256];
    code[threadIdx.x] = D.lut[threadIdx.x];
    __syncthreads();
    Group g = groups[blockIdx.x];
    u32 head = g_head(g), size = g_size(g);
    u32 n2 = 1; while (n2 < size) n2 <<= 1;
    for (u32 e = threadIdx.x; e < n2; e += blockDim.x) {
        u64 v = ~0ull;
        if (e < size) { u32 s = sa[head + e]; suf[e] = s; v = chunk_key(D, code, s, h); }
        key[e] = v; idx[e] = (u16)e;
    }
    __syncthreads();
    for (u32 k = 2; k <= n2; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {     // one thread per compare-exchange pair
                u32 e = ((t & ~(j - 1)) << 1) | (t & (j - 1)), partner = e | j;
                u64 a = key[e], b = key[partner]; u16 ia = idx[e], ib = idx[partner];
                bool up = (e & k) == 0;
                bool gt = a > b || (a == b && ia > ib);
                if (gt == up) { key[e] = b; key[partner] = a; idx[e] = ib; idx[partner] = ia; }
            }
            __syncthreads();
        }
    }
    __shared__ u32 slice_max[TPB];
    u32 per = (size + blockDim.x - 1) / blockDim.x;
    u32 lo = min(threadIdx.x * per, size), hi = min(lo + per, size);
    u32 run = 0;
    for (u32 e = lo; e < hi; ++e) {
        if (e == 0 || key[e] != key[e - 1]) run = e + 1;
        start_of[e] = (u16)run;
    }
    slice_max[threadIdx.x] = run;
    __syncthreads();
    for (u32 d = 1; d < blockDim.x; d <<= 1) {
        u32 v = (threadIdx.x >= d) ? slice_max[threadIdx.x - d] : 0u;
        __syncthreads();
        slice_max[threadIdx.x] = max(slice_max[threadIdx.x], v);
        __syncthreads();
    }
    u32 carry = threadIdx.x ? slice_max[threadIdx.x - 1] : 0u;
    for (u32 e = lo; e < hi; ++e) {
        u32 st = max((u32)start_of[e], carry) - 1;
        u64 kk = key[e];
        sa[head + e] = suf[idx[e]];
        bool last = (e + 1 == size) || key[e + 1] != kk;
        bool first = e == st;
        if (kk == 0 || (first && last)) ghead[head + e] = head + st;
        if (last && e != st && kk != 0) append_group_single(head + st, e - st + 1, next);
    }
}

// groups larger than one block sorts: flat list, sorted by chunk key and then (stably) by group
__global__ void __launch_bounds__(TPB) dlarge_gather(const Group* __restrict__ groups, const u32* __restrict__ base, u32 h, DictView D,
                                                     const u32* __restrict__ sa, u64* __restrict__ keys, u32* __restrict__ gid, u32* __restrict__ sfx,
                                                     u32* __restrict__ iota) {
    __shared__ u8 code[256];
    code[threadIdx.x] = D.lut[threadIdx.x];
    __syncthreads();
    Group g = groups[blockIdx.x];
    u32 head = g_head(g), size = g_size(g), b = base[blockIdx.x];
    for (u32 e = blockIdx.y * blockDim.x + threadIdx.x; e < size; e += gridDim.y * blockDim.x) {
        u32 s = sa[head + e];
        keys[b + e] = chunk_key(D, code, s, h); gid[b + e] = blockIdx.x; sfx[b + e] = s; iota[b + e] = b + e;
    }
}
__global__ void __launch_bounds__(TPB) dlarge_gid_keys(const u32* __restrict__ perm, const u32* __restrict__ gid, u32 A, u64* __restrict__ out) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < A) out[a] = gid[perm[a]];
}
__global__ void __launch_bounds__(TPB) dlarge_scatter(const u32* __restrict__ perm, const u64* __restrict__ keys, const u32* __restrict__ gid,
                                                      const u32* __restrict__ sfx, const Group* __restrict__ groups, const u32* __restrict__ base, u32 A,
                                                      u32* __restrict__ sa, u64* __restrict__ skey, u32* __restrict__ pos, u32* __restrict__ head_cand) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    u32 src = perm[a], g = gid[src];
    u32 p = g_head(groups[g]) + (a - base[g]);
    u64 k = keys[src];
    sa[p] = sfx[src]; skey[a] = k; pos[a] = p;
    bool first = (a == base[g]) || keys[perm[a - 1]] != k;
    head_cand[a] = first ? p : 0u;
}
__global__ void __launch_bounds__(TPB) dlarge_finish(const u32* __restrict__ pos, const u64* __restrict__ skey, const u32* __restrict__ head,
                                                     const u32* __restrict__ perm, const u32* __restrict__ gid, u32 A, u32* __restrict__ ghead,
                                                     GroupLists next) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false; u32 hd = 0, p = 0;
    if (a < A) {
        hd = head[a]; p = pos[a];
        u64 k = skey[a];
        bool last = (a + 1 == A) || skey[a + 1] != k || gid[perm[a + 1]] != gid[perm[a]];
        if (k == 0 || (last && hd == p)) ghead[p] = hd;
        want = last && hd != p && k != 0;
    }
    append_group_warp(want, hd, p - hd + 1, next);
}

template <int WIDTH, int K>
void launch_dwarp_class(const Group* groups, u32 n, u32 h, const DictView& D, u32* sa, u32* ghead, const GroupLists& next, cudaStream_t s) {
    if (!n) return;
    auto sort_dgroups_warp_k = sort_dgroups_warp<WIDTH, K>;
    MMB_LAUNCH(sort_dgroups_warp_k, ceil_div((u64)n * WIDTH, TPB), TPB, 0, s, groups, n, h, D, sa, ghead, next);
}

// Key groups after the initial sort: ghead of every rank, and the groups that need refinement (refine[p] != 0 on their ranks, more than
// one member, and strings that continue past the key).  Equal keys whose last symbol is 0 (terminator or the padding after it) are
// strings that ended inside the key: equal strings, a final tie.
__global__ void __launch_bounds__(TPB) list_dict_groups(const u64* __restrict__ keys, const u32* __restrict__ head, u32 m, u32 bits,
                                                        const u8* __restrict__ refine, u32* __restrict__ ghead, GroupLists L) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false; u32 hd = 0;
    if (p < m) {
        hd = head[p];
        u64 key = keys[p];
        bool last = (p + 1 == m) || keys[p + 1] != key;
        bool multi = hd != p || !last;
        bool final_tie = multi && (key & ((1ull << bits) - 1)) == 0;
        ghead[p] = final_tie ? hd : p;              // refined groups overwrite this when their runs become final
        want = last && hd != p && refine[p] && !final_tie;
    }
    append_group_warp(want, hd, p - hd + 1, L);
}

const u32 kClassMinSize[NCLASS] = {2, 5, 9, 17, 33, 65, 129, 257, BLOCK_SORT_MAX + 1};

}  // namespace

void GroupStore::alloc(u32 m, cudaStream_t s) {
    u64 total = 0;
    for (int c = 0; c < NCLASS; ++c) { first[c] = total; total += (u64)m / kClassMinSize[c] + 1; }
    first[NCLASS] = total;
    groups.alloc(total, s); counts.alloc(NCLASS, s); counts.zero();
}

static GroupLists lists_of(GroupStore& g) {
    GroupLists L;
    for (int c = 0; c < NCLASS; ++c) L.list[c] = g.groups.p + g.first[c];
    L.count = g.counts.p;
    return L;
}

// Refine an order that is already sorted by the first h0 symbols (sa, isa = head ranks, cur = multi-member groups by size class,
// all on the device) until every group is a single suffix or a set of suffixes that are equal up to their `limit` (dictionary mode).
void refine_by_doubling(u32* sa, u32* isa_p, u32 m, u32 h0, const u32* limit, GroupStore& cur, SaStats& st, cudaStream_t s, u32 isa_stride) {
    Isa isa{isa_p, isa_stride};
    GroupStore nxt; nxt.alloc(m, s);
    DevBuf<u32> nh; nh.alloc(m, s);
    DevBuf<u8> changed; changed.alloc(cur.first[NCLASS], s);      // one flag per listed group, laid out like the class lists
    int sort_bits = 32 + (int)bits_for(m);
    for (u64 h = h0;; h *= 2) {
        std::vector<u32> cnt = to_host(cur.counts.p, NCLASS, s);
        u64 total = 0;
        for (int c = 0; c < NCLASS; ++c) {
            if ((u64)cnt[c] > cur.first[c + 1] - cur.first[c]) throw Error(23, "internal: group list overflow");
            total += cnt[c];
        }
        if (getenv("MMB_DEBUG")) {
            fprintf(stderr, "[sa] m=%u h=%llu classes:", m, (unsigned long long)h);
            for (int c = 0; c < NCLASS; ++c) fprintf(stderr, " %u", cnt[c]);
            fprintf(stderr, "\n");
        }
        if (total == 0) break;
        if (h >= 2 * (u64)m) throw Error(22, "internal: prefix doubling did not converge");
        u32 hh = (u32)std::min<u64>(h, m);
        nxt.counts.zero();
        GroupLists next = lists_of(nxt);
        const Group* G[NCLASS];
        for (int c = 0; c < NCLASS; ++c) G[c] = cur.groups.p + cur.first[c];
        launch_warp_class<4, 1>(G[0], cnt[0], hh, m, sa, isa, limit, nh.p, changed.p + cur.first[0], next, s);
        launch_warp_class<8, 1>(G[1], cnt[1], hh, m, sa, isa, limit, nh.p, changed.p + cur.first[1], next, s);
        launch_warp_class<16, 1>(G[2], cnt[2], hh, m, sa, isa, limit, nh.p, changed.p + cur.first[2], next, s);
        launch_warp_class<32, 1>(G[3], cnt[3], hh, m, sa, isa, limit, nh.p, changed.p + cur.first[3], next, s);
        launch_warp_class<32, 2>(G[4], cnt[4], hh, m, sa, isa, limit, nh.p, changed.p + cur.first[4], next, s);
        launch_warp_class<32, 4>(G[5], cnt[5], hh, m, sa, isa, limit, nh.p, changed.p + cur.first[5], next, s);
        launch_warp_class<32, 8>(G[6], cnt[6], hh, m, sa, isa, limit, nh.p, changed.p + cur.first[6], next, s);
        if (cnt[7]) MMB_LAUNCH(sort_groups_block, cnt[7], TPB, 0, s, G[7], hh, m, sa, isa, limit, nh.p, next);
        DevBuf<u32> lpos; u32 A = 0;
        if (cnt[8]) {
            u32 nl = cnt[8];
            DevBuf<u32> sizes, base, lval_a, lval_b, cand, head; DevBuf<u64> lkey_a, lkey_b, by_head;
            sizes.alloc(nl + 1, s); base.alloc(nl + 1, s);
            // Groups were appended in arbitrary (atomic) order; the flat list below must follow suffix-array order, because after the
            // radix sort (head in the high key bits) the a-th pair belongs at the a-th listed position.
            by_head.alloc(nl, s); lval_a.alloc(nl, s); lval_b.alloc(nl, s);
            MMB_CUDA(cudaMemsetAsync(lval_a.p, 0, (size_t)nl * 4, s));
            prim::sort_pairs(G[8], by_head.p, lval_a.p, lval_b.p, nl, 32, 64, s);
            G[8] = by_head.p;
            MMB_LAUNCH(large_sizes, ceil_div(nl + 1, TPB), TPB, 0, s, G[8], nl, sizes.p);
            prim::exclusive_sum(sizes.p, base.p, nl + 1, s);
            A = to_host(base.p + nl, 1, s)[0];
            lpos.alloc(A, s); lkey_a.alloc(A, s); lkey_b.alloc(A, s); lval_a.alloc(A, s); lval_b.alloc(A, s); cand.alloc(A, s); head.alloc(A, s);
            dim3 grid(nl, 32);
            MMB_LAUNCH(large_gather, grid, TPB, 0, s, G[8], (const u32*)base.p, hh, m, (const u32*)sa, isa, limit, lkey_a.p, lval_a.p, lpos.p);
            prim::sort_pairs(lkey_a.p, lkey_b.p, lval_a.p, lval_b.p, A, 0, sort_bits, s);
            MMB_LAUNCH(large_scatter, ceil_div(A, TPB), TPB, 0, s, (const u32*)lpos.p, (const u64*)lkey_b.p, (const u32*)lval_b.p, A, sa, cand.p);
            prim::inclusive_max(cand.p, head.p, A, s);
            MMB_LAUNCH(large_finish, ceil_div(A, TPB), TPB, 0, s, (const u32*)lpos.p, (const u64*)lkey_b.p, (const u32*)head.p, A, nh.p, next);
        }
        launch_apply_class<4, 1>(G[0], cnt[0], sa, nh.p, changed.p + cur.first[0], isa, s);
        launch_apply_class<8, 1>(G[1], cnt[1], sa, nh.p, changed.p + cur.first[1], isa, s);
        launch_apply_class<16, 1>(G[2], cnt[2], sa, nh.p, changed.p + cur.first[2], isa, s);
        launch_apply_class<32, 1>(G[3], cnt[3], sa, nh.p, changed.p + cur.first[3], isa, s);
        launch_apply_class<32, 2>(G[4], cnt[4], sa, nh.p, changed.p + cur.first[4], isa, s);
        launch_apply_class<32, 4>(G[5], cnt[5], sa, nh.p, changed.p + cur.first[5], isa, s);
        launch_apply_class<32, 8>(G[6], cnt[6], sa, nh.p, changed.p + cur.first[6], isa, s);
        if (cnt[7]) MMB_LAUNCH(apply_groups_block, cnt[7], TPB, 0, s, G[7], (const u32*)sa, (const u32*)nh.p, isa);
        if (A) MMB_LAUNCH(large_apply, ceil_div(A, TPB), TPB, 0, s, (const u32*)lpos.p, A, (const u32*)sa, (const u32*)nh.p, isa);
        cur.groups.swap(nxt.groups); cur.counts.swap(nxt.counts);
        ++st.rounds; st.groups_sorted += total; st.large_elements += A;
    }
}

// Sort (key, suffix) pairs, derive head ranks and the first group lists.  keys are consumed.
void order_by_initial_keys(u64* keys_in, u64* keys_tmp, u32* vals_in, u32 m, int key_bits, u32 h0, const u32* limit, u32* sa, u32* isa_p,
                           GroupStore& groups, cudaStream_t s, u32 isa_stride) {
    Isa isa{isa_p, isa_stride};
    DevBuf<u32> cand, head;
    cand.alloc(m, s); head.alloc(m, s);
    groups.alloc(m, s);
    prim::sort_pairs(keys_in, keys_tmp, vals_in, sa, m, 0, key_bits, s);
    MMB_LAUNCH(first_flags, ceil_div(m, TPB), TPB, 0, s, (const u64*)keys_tmp, m, cand.p);
    prim::inclusive_max(cand.p, head.p, m, s);
    MMB_LAUNCH(first_ranks, ceil_div(m, TPB), TPB, 0, s, (const u64*)keys_tmp, (const u32*)sa, (const u32*)head.p, m, h0, limit, isa, lists_of(groups));
}

bool refine_dictionary_chunks(const u64* sorted_keys, u32* sa, u32 m, u32 cpk, u32 bits, const u8* dict, const u32* limit, const u8* lut,
                              const u8* refine, u32* ghead, SaStats& st, cudaStream_t s, u32 max_rounds) {
    DictView D{dict, limit, lut, bits, cpk};
    GroupStore cur, nxt;
    cur.alloc(m, s); nxt.alloc(m, s);
    {
        DevBuf<u32> cand, head; cand.alloc(m, s); head.alloc(m, s);
        MMB_LAUNCH(first_flags, ceil_div(m, TPB), TPB, 0, s, sorted_keys, m, cand.p);
        prim::inclusive_max(cand.p, head.p, m, s);
        MMB_LAUNCH(list_dict_groups, ceil_div(m, TPB), TPB, 0, s, sorted_keys, (const u32*)head.p, m, bits, refine, ghead, lists_of(cur));
    }
    for (u32 r = 1;; ++r) {
        std::vector<u32> cnt = to_host(cur.counts.p, NCLASS, s);
        u64 total = 0;
        for (int c = 0; c < NCLASS; ++c) {
            if ((u64)cnt[c] > cur.first[c + 1] - cur.first[c]) throw Error(23, "internal: group list overflow");
            total += cnt[c];
        }
        if (getenv("MMB_DEBUG")) {
            fprintf(stderr, "[dsort] m=%u round=%u classes:", m, r);
            for (int c = 0; c < NCLASS; ++c) fprintf(stderr, " %u", cnt[c]);
            fprintf(stderr, "\n");
        }
        if (total == 0) return true;
        if (r > max_rounds) return false;            // very long shared prefixes (long low-complexity phrases): the caller falls back to doubling
        u64 h64 = (u64)cpk * r;
        if (h64 >= 0xffffffffull) return false;
        u32 h = (u32)h64;
        nxt.counts.zero();
        GroupLists next = lists_of(nxt);
        const Group* G[NCLASS];
        for (int c = 0; c < NCLASS; ++c) G[c] = cur.groups.p + cur.first[c];
        launch_dwarp_class<4, 1>(G[0], cnt[0], h, D, sa, ghead, next, s);
        launch_dwarp_class<8, 1>(G[1], cnt[1], h, D, sa, ghead, next, s);
        launch_dwarp_class<16, 1>(G[2], cnt[2], h, D, sa, ghead, next, s);
        launch_dwarp_class<32, 1>(G[3], cnt[3], h, D, sa, ghead, next, s);
        launch_dwarp_class<32, 2>(G[4], cnt[4], h, D, sa, ghead, next, s);
        launch_dwarp_class<32, 4>(G[5], cnt[5], h, D, sa, ghead, next, s);
        launch_dwarp_class<32, 8>(G[6], cnt[6], h, D, sa, ghead, next, s);
        if (cnt[7]) MMB_LAUNCH(sort_dgroups_block, cnt[7], TPB, 0, s, G[7], h, D, sa, ghead, next);
        u32 A = 0;
        if (cnt[8]) {
            u32 nl = cnt[8];
            DevBuf<u32> sizes, base, gid, sfx, perm_a, perm_b, perm_c, pos, cand, head; DevBuf<u64> ck, ck_b, gk, gk_b, skey, by_head;
            sizes.alloc(nl + 1, s); base.alloc(nl + 1, s);
            // groups were appended in arbitrary order; the run-start scan below needs them in suffix-array order
            by_head.alloc(nl, s); perm_a.alloc(nl, s); perm_b.alloc(nl, s);
            MMB_CUDA(cudaMemsetAsync(perm_a.p, 0, (size_t)nl * 4, s));
            prim::sort_pairs(G[8], by_head.p, perm_a.p, perm_b.p, nl, 32, 64, s);
            G[8] = by_head.p;
            MMB_LAUNCH(large_sizes, ceil_div(nl + 1, TPB), TPB, 0, s, G[8], nl, sizes.p);
            prim::exclusive_sum(sizes.p, base.p, nl + 1, s);
            A = to_host(base.p + nl, 1, s)[0];
            ck.alloc(A, s); ck_b.alloc(A, s); gk.alloc(A, s); gk_b.alloc(A, s); skey.alloc(A, s); gid.alloc(A, s); sfx.alloc(A, s);
            perm_a.alloc(A, s); perm_b.alloc(A, s); perm_c.alloc(A, s); pos.alloc(A, s); cand.alloc(A, s); head.alloc(A, s);
            dim3 grid(nl, 32);
            MMB_LAUNCH(dlarge_gather, grid, TPB, 0, s, G[8], (const u32*)base.p, h, D, (const u32*)sa, ck.p, gid.p, sfx.p, perm_a.p);
            prim::sort_pairs(ck.p, ck_b.p, perm_a.p, perm_b.p, A, 0, (int)(bits * cpk), s);
            MMB_LAUNCH(dlarge_gid_keys, ceil_div(A, TPB), TPB, 0, s, (const u32*)perm_b.p, (const u32*)gid.p, A, gk.p);
            prim::sort_pairs(gk.p, gk_b.p, perm_b.p, perm_c.p, A, 0, (int)bits_for(nl), s);
            MMB_LAUNCH(dlarge_scatter, ceil_div(A, TPB), TPB, 0, s, (const u32*)perm_c.p, (const u64*)ck.p, (const u32*)gid.p, (const u32*)sfx.p, G[8],
                       (const u32*)base.p, A, sa, skey.p, pos.p, cand.p);
            prim::inclusive_max(cand.p, head.p, A, s);
            MMB_LAUNCH(dlarge_finish, ceil_div(A, TPB), TPB, 0, s, (const u32*)pos.p, (const u64*)skey.p, (const u32*)head.p, (const u32*)perm_c.p,
                       (const u32*)gid.p, A, ghead, next);
        }
        cur.groups.swap(nxt.groups); cur.counts.swap(nxt.counts);
        ++st.rounds; st.groups_sorted += total; st.large_elements += A;
    }
}

void build_suffix_array(const u8* text, u64 n, SuffixArray& out, cudaStream_t s) {
    if (n + 1 >= 0xfffffff0ull) throw Error(20, "text too long for the direct (32-bit) suffix array builder");
    u32 m = (u32)(n + 1);
    out.m = m;
    out.sa.alloc(m, s); out.isa.alloc(m, s);
    out.stats = SaStats{};

    // symbol codes: 1 + rank among the symbols present, 0 is the padding / sentinel code
    DevBuf<u32> d_hist; d_hist.alloc(256, s); d_hist.zero();
    MMB_LAUNCH(symbol_histogram, 148 * 8, TPB, 0, s, text, n, d_hist.p);
    std::vector<u32> hist = to_host(d_hist.p, 256, s);
    std::vector<u8> lut(256, 0);
    u32 sigma = 0;
    for (int c = 0; c < 256; ++c) if (hist[c]) lut[c] = (u8)(++sigma);
    if (hist[0]) throw Error(21, "NUL byte in sequence data");
    u32 bits = bits_for(sigma), cpk = 64 / bits;
    DevBuf<u8> d_lut; d_lut.alloc(256, s);
    MMB_CUDA(cudaMemcpyAsync(d_lut.p, lut.data(), 256, cudaMemcpyHostToDevice, s));

    GroupStore groups;
    {
        DevBuf<u64> key_a, key_b; DevBuf<u32> val_a;
        key_a.alloc(m, s); key_b.alloc(m, s); val_a.alloc(m, s);
        MMB_LAUNCH(pack_text_keys, ceil_div(m, TPB), TPB, 0, s, text, n, (const u8*)d_lut.p, bits, cpk, key_a.p, val_a.p);
        order_by_initial_keys(key_a.p, key_b.p, val_a.p, m, (int)(bits * cpk), cpk, nullptr, out.sa.p, out.isa.p, groups, s);
    }
    refine_by_doubling(out.sa.p, out.isa.p, m, cpk, nullptr, groups, out.stats, s);
}

}  // namespace mmb
