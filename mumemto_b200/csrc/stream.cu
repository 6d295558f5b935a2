// mumemto_b200/csrc/stream.cu -- enumerate the suffix stream of T (rank order: SA and LCP) from the prefix-free-parse index, tile by tile.
//
// Replaces the hot loop of the reference, pfp_lcp::process (include/pfp_lcp_mum.hpp:115-231): walk the dictionary suffix array, group
// equal dictionary suffixes, merge the occurrence lists of their phrases by the rank of the parse suffix that follows (a priority
// queue there), and for every text suffix derive SA (:337-341) and LCP (:145,179,284-321) -- row D of SURVEY.md section 8.
//
// Here a "member" is one owned dictionary suffix (longer than w) in dictionary-suffix order; it contributes one text suffix per
// occurrence of its phrase.  Members with equal strings form a group.
//   expand_members   one warp per member streams its phrase's occurrence list: SA = start of the following phrase - (|alpha| - w).
//                    A group with a single member is already in final order and its LCPs are |alpha| - w + (precomputed LCP between
//                    consecutive occurrences); the first suffix of a group gets lcp(alpha_prev, alpha) from a byte comparison.
//                    Groups of 2..8 members are merged in the same kernel by binary searches in the other members' (sorted) lists.
//   merge_small      groups with more members (repeats), <= 2048 suffixes: one block, bitonic sort by following-suffix rank in shared
//                    memory, LCP by range-minimum queries over s_lcp.      merge_large: the same through a device-wide radix sort.
// Tiles end on group boundaries and carry `halo` ranks of the previous tile so that the interval detector can look back.
// Traffic per text suffix: occurrence list 16 B read, SA 8 B + LCP 4 B written (+ 4 B key for multi-member groups).
#include "pfp.h"
#include "pfp_device.cuh"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 TPB = 256;
constexpr u32 MERGE_MAX = 2048;
constexpr u32 NOT_START = 0xffffffffu;
constexpr u32 KLCP_UNKNOWN = 255;        // mem_klcp: the key-level lcp with the previous member was not derived (more symbols than any key holds)

__device__ __forceinline__ u64 load8(const u8* __restrict__ t, u64 i) {
    const u32* w = reinterpret_cast<const u32*>(t + (i & ~3ull));
    u32 sh = (u32)(i & 3) * 8;
    u32 a = w[0], b = w[1], c = w[2];
    u32 lo = sh ? __funnelshift_r(a, b, sh) : a;
    u32 hi = sh ? __funnelshift_r(b, c, sh) : b;
    return ((u64)hi << 32) | lo;
}

__global__ void __launch_bounds__(TPB) owned_flags(const u32* __restrict__ sa_d, const u32* __restrict__ limit, u32 dsize, u32* __restrict__ flag) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < dsize) flag[p] = limit[sa_d[p]] >= PFP_W + 2 ? 1u : 0u;      // |alpha| = limit - 1 > w
    if (p == dsize) flag[p] = 0;
}

// Everything expand_members needs to know about a member, gathered once (16 bytes, read coalesced there).
constexpr u8 MEMBER_STAND_IN = 2;        // MemberMeta::inside of a stand-in rank (pruned route)
struct MemberMeta {
    u32 q0;         // first entry of the phrase's occurrence list
    u32 cnt;        // occurrences of the phrase
    u32 shift;      // |alpha| - w
    u8 before;      // symbol preceding the suffix inside the phrase (valid if inside)
    u8 inside;      // 0: the suffix starts its phrase, the preceding symbol comes from the text
    u8 back;        // distance to the first member of the group (FUSE_MAX_MEMBERS + 1 = group too large to merge in place)
    u8 members;     // members in the group (same cap)
};

// Everything members_fill needs to know about a dictionary position, gathered once in dictionary order (coalesced) so that the
// member pass does one 16-byte random read per member instead of four.
struct __align__(16) DictRecord { u32 limit, phrase, head, before; };      // before: symbol preceding the position, 0x100 = it starts its phrase
__global__ void __launch_bounds__(TPB) dict_records(const u32* __restrict__ limit, const u32* __restrict__ dphrase, const u32* __restrict__ isa_d,
                                                    const u32* __restrict__ doff, const u8* __restrict__ dict, u32 dsize, DictRecord* __restrict__ rec) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dsize) return;
    DictRecord r; r.limit = limit[i]; r.phrase = dphrase[i]; r.head = isa_d ? isa_d[i] : 0u;
    r.before = i > doff[r.phrase] ? (u32)dict[i - 1] : 0x100u;
    rec[i] = r;
}

__global__ void __launch_bounds__(TPB) members_fill(const u32* __restrict__ sa_d, const DictRecord* __restrict__ rec, const u32* __restrict__ prank,
                                                    const u32* __restrict__ occ_first, const u32* __restrict__ flag, const u32* __restrict__ idx, u32 dsize,
                                                    const u8* __restrict__ klcp, const u32* __restrict__ ghead, const u8* __restrict__ mkind,
                                                    u32* __restrict__ mem_i, u32* __restrict__ mem_head,
                                                    u32* __restrict__ mem_cnt, MemberMeta* __restrict__ meta, u8* __restrict__ mem_klcp) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dsize || !flag[p]) return;
    u32 k = idx[p], i = sa_d[p];
    if (klcp && ghead) mem_klcp[k] = klcp[p];      // pruned route: already the symbols shared with the previous suffix longer than w
    else if (klcp) {    // key symbols shared with the previous MEMBER: minimum over the ranks skipped in between (suffixes of <= w symbols)
        // Runs of skipped ranks are usually a few entries, but equal short strings (phrase ends) form runs of millions: give up after
        // 64 steps and let group_lcp compare the bytes from the start (KLCP_UNKNOWN).
        u32 mn = klcp[p], q = p, steps = 0;
        while (q > 0 && !flag[q - 1] && steps < 64) { --q; ++steps; mn = min(mn, (u32)klcp[q]); }
        if (q > 0 && !flag[q - 1]) mn = KLCP_UNKNOWN;
        mem_klcp[k] = (u8)mn;
    }
    DictRecord d = rec[i];
    u32 r = prank[d.phrase];
    u32 q0 = occ_first[r], cnt = occ_first[r + 1] - q0;
    const bool stand_in = mkind && (mkind[p] & 3) == 2;      // one rank that stands for the pruned suffixes after a surviving segment
    if (stand_in) cnt = 1;
    mem_i[k] = i; mem_head[k] = ghead ? ghead[p] : d.head; mem_cnt[k] = cnt;
    MemberMeta m;
    m.q0 = q0; m.cnt = cnt; m.shift = d.limit - 1 - PFP_W;
    m.inside = stand_in ? MEMBER_STAND_IN : (d.before < 0x100u ? 1 : 0); m.before = (u8)d.before; m.back = 0; m.members = 1;
    meta[k] = m;
}

// lcp between the strings of consecutive members that start a new group: read off the initial sort keys when the two differ inside
// them (mem_klcp < key_symbols), else a byte comparison inside the dictionary that starts after the key
__global__ void __launch_bounds__(TPB) group_lcp(const u8* __restrict__ dict, const MemberMeta* __restrict__ meta, const u32* __restrict__ mem_i,
                                                 const u32* __restrict__ mem_head, const u8* __restrict__ mem_klcp, u32 key_symbols, u32 K,
                                                 u32* __restrict__ glcp) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    if (k == 0) {       // the first member follows the sentinel suffix -- or, on the pruned route, suffixes that were left out
        u32 shared = mem_klcp ? mem_klcp[0] : 0u;
        glcp[0] = shared < key_symbols ? min(shared, meta[0].shift + PFP_W) : 0u;
        return;
    }
    if (mem_head[k] == mem_head[k - 1]) { glcp[k] = NOT_START; return; }
    u32 cap = min(meta[k - 1].shift, meta[k].shift) + PFP_W, l = 0;      // |alpha| of the shorter; different strings of a prefix-free set differ before either ends
    if (mem_klcp) {
        u32 shared = mem_klcp[k];
        // the two differ inside the key (on the pruned route member k - 1 need not be the suffix this count was taken against: the
        // suffixes in between were pruned -- only the member's own length bounds it)
        if (shared < key_symbols) { glcp[k] = min(shared, meta[k].shift + PFP_W); return; }
        if (shared != KLCP_UNKNOWN) l = min(key_symbols, cap);
    }
    u32 a = mem_i[k - 1], b = mem_i[k];
    while (l + 8 <= cap) {
        u64 x = load8(dict, (u64)a + l) ^ load8(dict, (u64)b + l);
        if (x) { l += (u32)((__ffsll((long long)x) - 1) >> 3); glcp[k] = l; return; }
        l += 8;
    }
    while (l < cap && dict[a + l] == dict[b + l]) ++l;
    glcp[k] = l;
}

constexpr u32 FUSE_MAX_MEMBERS = 8;      // groups with at most this many members are merged inside expand_members

// distance to the group start and group size for every member, capped at FUSE_MAX_MEMBERS + 1
__global__ void __launch_bounds__(TPB) group_shape(const u32* __restrict__ glcp, u32 K, MemberMeta* __restrict__ meta) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    u32 ks = k, ke = k + 1;
    for (u32 b = 0; glcp[ks] == NOT_START && b <= FUSE_MAX_MEMBERS; ++b) --ks;
    bool big = glcp[ks] == NOT_START;
    while (!big && ke < K && glcp[ke] == NOT_START) { ++ke; if (ke - ks > FUSE_MAX_MEMBERS) big = true; }
    meta[k].back = (u8)(big ? FUSE_MAX_MEMBERS + 1 : k - ks);
    meta[k].members = (u8)(big ? FUSE_MAX_MEMBERS + 1 : ke - ks);
}

__global__ void tile_bounds(const u64* __restrict__ out_base, const u32* __restrict__ glcp, u32 K, u64 tile_ranks, u32 nt, u32* __restrict__ bounds) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nt) return;
    if (t == nt) { bounds[t] = K; return; }
    u64 target = (u64)t * tile_ranks;
    u32 lo = 0, hi = K;                                     // first k with out_base[k] >= target
    while (lo < hi) { u32 mid = lo + ((hi - lo) >> 1); if (out_base[mid] >= target) hi = mid; else lo = mid + 1; }
    while (lo > 0 && lo < K && glcp[lo] == NOT_START) --lo;   // back to the start of that member's group
    bounds[t] = lo;
}

// first index in the ascending list v[0..c) whose value is >= x
__device__ __forceinline__ u32 lower_bound_u32(const u32* __restrict__ v, u32 c, u32 x) {
    u32 lo = 0, hi = c;
    while (lo < hi) { u32 mid = (lo + hi) >> 1; if (v[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

// Load-balanced expansion: each warp owns 32 consecutive members, scans their occurrence counts, and then walks the warp's output
// range 32 suffixes at a time, so every lane produces one text suffix per step no matter how the occurrences are spread over the
// members (phrases occur 1 .. 2N times).  The occurrence list of every member is already ordered by the rank of the following parse
// suffix, so in a group of M <= 8 members the final slot of an element is its own index plus, for each other member, the number of
// that member's elements with a smaller rank (one binary search each); the same searches give the predecessor in merged order, hence
// the LCP.  Groups with more members only get their (rank, SA) pairs written here and are sorted afterwards.
constexpr u32 EXP_TPB = 256;
// The kernel waits on memory (ncu: long-scoreboard stalls dominate), so what matters is how many loads are in flight: MINB = resident
// blocks per SM the register allocation is tuned for, PREFETCH = how many iterations ahead the list entries are touched.  Measured on
// C2 (profiles/README.md): <6,0> 170 ms, <8,0> 144 ms, <6,1> 159 ms, <8,1> 138 ms per step.
template <int MINB, int PREFETCH>
__global__ void __launch_bounds__(EXP_TPB, MINB) expand_members(u32 k0, u32 k1, const MemberMeta* __restrict__ meta, const u64* __restrict__ out_base,
                                                                const u32* __restrict__ glcp, const u32* __restrict__ occ_next_rank,
                                                                const u64* __restrict__ occ_next_start, const u32* __restrict__ occ_lcp,
                                                                const u32* __restrict__ s_lcp, const u32* __restrict__ rmq, u32 rmq_blocks,
                                                                const u8* __restrict__ text, u64 text_len, u64 lead, u64* __restrict__ sa, u32* __restrict__ lcp,
                                                                u32* __restrict__ key, u8* __restrict__ bwt) {
    __shared__ MemberMeta smeta[EXP_TPB / 32][32];
    __shared__ u32 soff[EXP_TPB / 32][33];
    u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u64 kw64 = (u64)k0 + ((u64)blockIdx.x * (EXP_TPB / 32) + w) * 32;          // first member of this warp
    if (kw64 >= k1) return;                                                    // whole warp
    u32 kw = (u32)kw64, k = kw + lane;
    MemberMeta me; me.q0 = 0; me.cnt = 0; me.shift = 0; me.before = 0; me.inside = 0; me.back = 0; me.members = 1;
    if (k < k1) me = meta[k];
    u32 incl = me.cnt;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    u32 T = __shfl_sync(0xffffffffu, incl, 31);
    smeta[w][lane] = me; soff[w][lane] = incl - me.cnt;
    if (lane == 31) soff[w][32] = T;
    __syncwarp();
    const u64 tile_base = out_base[k0];
    const u64 wbase = out_base[kw] - tile_base + lead;
    auto member_of = [&](u32 e) -> u32 {                   // largest index with soff[mi] <= e (every member has at least one occurrence)
        u32 mi = 0;
#pragma unroll
        for (u32 step = 16; step > 0; step >>= 1) if (mi + step < 32 && soff[w][mi + step] <= e) mi += step;
        return mi;
    };
    u32 mi_next = lane < T ? member_of(lane) : 0u;
    for (u32 e = lane; e < T; e += 32) {
        const u32 mi = mi_next;
        const MemberMeta m = smeta[w][mi];
        u32 t = e - soff[w][mi], kk = kw + mi, q = m.q0 + t;
        if (e + 32 < T) mi_next = member_of(e + 32);
        if (PREFETCH > 0 && e + 32 * PREFETCH < T) {
#ifndef MMB_SIMT
            const u32 ep = e + 32 * PREFETCH;
            const u32 mp = PREFETCH == 1 ? mi_next : member_of(ep);
            const u32 qn = smeta[w][mp].q0 + (ep - soff[w][mp]);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(occ_next_start + qn));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(occ_lcp + qn));
#endif
        }
        if (m.inside == MEMBER_STAND_IN) {                 // pruned route: closes the segment before it with the true LCP, matches nothing
            sa[wbase + e] = text_len; bwt[wbase + e] = 0; lcp[wbase + e] = glcp[kk];
            continue;
        }
        u64 pos = occ_next_start[q] - m.shift;
        u8 sym = m.inside ? m.before : (pos ? text[pos - 1] : (u8)0);
        if (m.members == 1) {
            sa[wbase + e] = pos; bwt[wbase + e] = sym;
            lcp[wbase + e] = t ? m.shift + occ_lcp[q] : glcp[kk];
        } else if (m.members > FUSE_MAX_MEMBERS) {
            sa[wbase + e] = pos; key[wbase + e] = occ_next_rank[q]; bwt[wbase + e] = sym;      // merge_small permutes all three; the large path re-reads the text
        } else {
            u32 ks = kk - m.back;
            u64 gbase = (ks >= kw) ? wbase + soff[w][ks - kw] : out_base[ks] - tile_base + lead;
            u32 r = occ_next_rank[q], slot = t;
            bool has_pred = t > 0;
            u32 pred = has_pred ? occ_next_rank[q - 1] : 0u;
            for (u32 j = 0; j < m.members; ++j) {
                u32 kj = ks + j;
                if (kj == kk) continue;
                u32 oq, oc;
                if (kj >= kw && kj < kw + 32) { oq = smeta[w][kj - kw].q0; oc = smeta[w][kj - kw].cnt; }
                else { MemberMeta o = meta[kj]; oq = o.q0; oc = o.cnt; }
                u32 lb = lower_bound_u32(occ_next_rank + oq, oc, r);
                slot += lb;
                if (lb) { u32 v = occ_next_rank[oq + lb - 1]; if (!has_pred || v > pred) pred = v; has_pred = true; }
            }
            sa[gbase + slot] = pos; bwt[gbase + slot] = sym;
            lcp[gbase + slot] = has_pred ? m.shift + range_min(s_lcp, rmq, rmq_blocks, pred + 1, r) : glcp[ks];
        }
    }
}

// group starts that have more than one member -> (first member, end member) by size
constexpr u32 MERGE_WARP_MAX = 256;      // groups of at most this many suffixes are ordered by one warp (merge_warp)
__global__ void __launch_bounds__(TPB) collect_multi(u32 k0, u32 k1, u32 K, const u32* __restrict__ glcp, const u64* __restrict__ out_base,
                                                     u64* __restrict__ tiny, u64* __restrict__ small, u64* __restrict__ large, u32* __restrict__ counts) {
    u32 k = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k1 || glcp[k] == NOT_START || k + 1 >= K || glcp[k + 1] != NOT_START) return;
    u32 e = k + 1;
    while (e < K && glcp[e] == NOT_START) ++e;
    if (e - k <= FUSE_MAX_MEMBERS) return;                     // merged inside expand_members
    u64 g = out_base[e] - out_base[k];
    if (g <= MERGE_WARP_MAX) tiny[atomicAdd(&counts[2], 1u)] = ((u64)k << 32) | e;
    else if (g <= MERGE_MAX) small[atomicAdd(&counts[0], 1u)] = ((u64)k << 32) | e;
    else large[atomicAdd(&counts[1], 1u)] = ((u64)k << 32) | e;
}

// One warp per group of at most 32 K suffixes: every element counts the smaller keys of its group through shuffles (the keys -- ranks of
// the following parse suffixes -- are distinct), which is its slot, and remembers the largest of them, which is its predecessor.
template <int K, int LO>
__global__ void __launch_bounds__(TPB) merge_warp(const u64* __restrict__ groups, u32 ngroups, u32 k0, const u64* __restrict__ out_base,
                                                  const u32* __restrict__ glcp, const u32* __restrict__ mem_i, const u32* __restrict__ limit,
                                                  const u32* __restrict__ s_lcp, const u32* __restrict__ rmq, u32 rmq_blocks, u64 lead,
                                                  u64* __restrict__ sa, u32* __restrict__ lcp, const u32* __restrict__ key, u8* __restrict__ bwt) {
    u32 gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gi >= ngroups) return;                                  // whole warp
    u64 g = groups[gi];
    u32 ks = (u32)(g >> 32), ke = (u32)g;
    u64 base = out_base[ks] - out_base[k0] + lead;
    u32 G = (u32)(out_base[ke] - out_base[ks]);
    if (G > 32u * K || G <= (u32)LO) return;                    // handled by another instance
    u32 shift = limit[mem_i[ks]] - 1 - PFP_W;
    u32 kv[K]; u64 pv[K]; u8 sv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        u32 e = lane + 32u * k;
        kv[k] = 0xffffffffu; pv[k] = 0; sv[k] = 0;
        if (e < G) { kv[k] = key[base + e]; pv[k] = sa[base + e]; sv[k] = bwt[base + e]; }
    }
    u32 lt[K], pred[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { lt[k] = 0; pred[k] = 0; }
#pragma unroll
    for (int k2 = 0; k2 < K; ++k2) {
        if (32u * k2 >= G) break;
        for (int j = 0; j < 32; ++j) {
            u32 kj = __shfl_sync(0xffffffffu, kv[k2], j);
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (kj < kv[k]) { ++lt[k]; pred[k] = max(pred[k], kj); }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        u32 e = lane + 32u * k;
        if (e < G) {
            u64 dst = base + lt[k];
            sa[dst] = pv[k]; bwt[dst] = sv[k];
            lcp[dst] = lt[k] ? shift + range_min(s_lcp, rmq, rmq_blocks, pred[k] + 1, kv[k]) : glcp[ks];
        }
    }
}

__global__ void __launch_bounds__(TPB) merge_small(const u64* __restrict__ groups, u32 k0, const u64* __restrict__ out_base, const u32* __restrict__ glcp,
                                                   const u32* __restrict__ mem_i, const u32* __restrict__ limit, const u32* __restrict__ s_lcp,
                                                   const u32* __restrict__ rmq, u32 rmq_blocks, u64 lead, u64* __restrict__ sa, u32* __restrict__ lcp,
                                                   const u32* __restrict__ key, u8* __restrict__ bwt) {
    __shared__ u64 item[MERGE_MAX];       // key << 32 | index
    __shared__ u64 pos[MERGE_MAX];
    __shared__ u8 sym[MERGE_MAX];         // symbol preceding each suffix (written by expand_members): permuted with the suffix
    u64 g = groups[blockIdx.x];
    u32 ks = (u32)(g >> 32), ke = (u32)g;
    u64 base = out_base[ks] - out_base[k0] + lead;
    u32 G = (u32)(out_base[ke] - out_base[ks]);
    u32 shift = limit[mem_i[ks]] - 1 - PFP_W;
    u32 n2 = 1; while (n2 < G) n2 <<= 1;
    for (u32 e = threadIdx.x; e < n2; e += blockDim.x) {
        if (e < G) { item[e] = ((u64)key[base + e] << 32) | e; pos[e] = sa[base + e]; sym[e] = bwt[base + e]; } else item[e] = ~0ull;
    }
    __syncthreads();
    for (u32 k = 2; k <= n2; k <<= 1)
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {     // one thread per compare-exchange PAIR (no idle partner half)
                u32 e = ((t & ~(j - 1)) << 1) | (t & (j - 1)), partner = e | j;
                u64 a = item[e], b = item[partner];
                if ((a > b) == ((e & k) == 0)) { item[e] = b; item[partner] = a; }
            }
            __syncthreads();
        }
    for (u32 e = threadIdx.x; e < G; e += blockDim.x) {
        u64 it = item[e];
        sa[base + e] = pos[(u32)it]; bwt[base + e] = sym[(u32)it];
        lcp[base + e] = e ? shift + range_min(s_lcp, rmq, rmq_blocks, (u32)(item[e - 1] >> 32) + 1, (u32)(it >> 32)) : glcp[ks];
    }
}

__global__ void large_sizes(const u64* __restrict__ groups, u32 n, const u64* __restrict__ out_base, u64* __restrict__ sizes) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { u64 g = groups[i]; sizes[i] = out_base[(u32)g] - out_base[(u32)(g >> 32)]; }
    if (i == n) sizes[i] = 0;
}
__global__ void __launch_bounds__(TPB) large_gather(const u64* __restrict__ groups, const u64* __restrict__ fbase, u32 k0, const u64* __restrict__ out_base,
                                                    u64 lead, const u64* __restrict__ sa, const u32* __restrict__ key, u64* __restrict__ fkeys,
                                                    u64* __restrict__ fvals) {
    u64 g = groups[blockIdx.x];
    u32 ks = (u32)(g >> 32), ke = (u32)g;
    u64 base = out_base[ks] - out_base[k0] + lead, G = out_base[ke] - out_base[ks], fb = fbase[blockIdx.x];
    for (u64 e = (u64)blockIdx.y * blockDim.x + threadIdx.x; e < G; e += (u64)gridDim.y * blockDim.x) {
        fkeys[fb + e] = ((u64)blockIdx.x << 32) | key[base + e];
        fvals[fb + e] = sa[base + e];
    }
}
__global__ void __launch_bounds__(TPB) large_scatter(const u64* __restrict__ groups, const u64* __restrict__ fbase, u64 total, u32 k0,
                                                     const u64* __restrict__ out_base, const u32* __restrict__ glcp, const u32* __restrict__ mem_i,
                                                     const u32* __restrict__ limit, const u32* __restrict__ s_lcp, const u32* __restrict__ rmq,
                                                     u32 rmq_blocks, u64 lead, const u64* __restrict__ fkeys, const u64* __restrict__ fvals,
                                                     u64* __restrict__ sa, u32* __restrict__ lcp) {
    u64 a = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= total) return;
    u64 fk = fkeys[a];
    u32 seq = (u32)(fk >> 32);
    u64 g = groups[seq];
    u32 ks = (u32)(g >> 32);
    u64 t = a - fbase[seq], dst = out_base[ks] - out_base[k0] + lead + t;
    u32 shift = limit[mem_i[ks]] - 1 - PFP_W;
    sa[dst] = fvals[a];
    lcp[dst] = t ? shift + range_min(s_lcp, rmq, rmq_blocks, (u32)fkeys[a - 1] + 1, (u32)fk) : glcp[ks];
}

__global__ void tile_edges(u64* __restrict__ sa, u32* __restrict__ lcp, u8* __restrict__ bwt, const u8* __restrict__ text, u64 n, int first_tile,
                           u64 count) {
    u32 t = threadIdx.x;
    if (t < 64) lcp[count + t] = 0;                              // zero slack for the detector's rounded tile loads
    if (first_tile && t == 0) { sa[0] = n; lcp[0] = 0; bwt[0] = n ? text[n - 1] : (u8)0; }      // rank 0: the sentinel suffix
}
// suffixes of groups that went through the sort kernels: the symbol before them from the text
__global__ void bwt_of_groups(const u64* __restrict__ groups, u32 k0, const u64* __restrict__ out_base, u64 lead, const u8* __restrict__ text,
                              const u64* __restrict__ sa, u8* __restrict__ bwt) {
    u64 g = groups[blockIdx.x];
    u64 base = out_base[(u32)(g >> 32)] - out_base[k0] + lead, G = out_base[(u32)g] - out_base[(u32)(g >> 32)];
    for (u64 e = (u64)blockIdx.y * blockDim.x + threadIdx.x; e < G; e += (u64)gridDim.y * blockDim.x) { u64 p = sa[base + e]; bwt[base + e] = p ? text[p - 1] : (u8)0; }
}

// document of every rank of the finished view (DocLookup): coarse table, then forward while the next document starts at or before p
__global__ void __launch_bounds__(TPB) tile_docs(const u64* __restrict__ sa, u64 count, DocLookup L, u64 n, u8* __restrict__ doc) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    u64 p = sa[r];
    if (p >= n) { doc[r] = 255; return; }
    u32 d = L.coarse[p >> L.shift];
    while (L.doc_start[d + 1] <= p) ++d;
    doc[r] = (u8)d;
}


// ---------------------------------------------------------------------------------------------- -M on the pruned route
// A threshold-only segment (pfp.h, PfpPrune) is ONE LCP interval: its N text suffixes share at least -l symbols, its neighbours share
// fewer.  The reference records a threshold for it when the N suffixes come from N distinct documents (include/mem_finder.hpp:326-336):
// thresh[position of the document-0 suffix] = min(max(LCP[first rank], LCP[closing rank]), 65535) -- and both LCPs are shared key symbols
// at the two segment boundaries.  No order inside the segment is needed: one warp gathers the occurrences of the segment's dictionary
// suffixes straight from the occurrence lists.  rk word layout: pfp.cu.
constexpr u32 RKW_HB = 1u << 31, RKW_OWN = 1u << 30;

template <int W>
__global__ void __launch_bounds__(TPB) segment_thresholds_kernel(const u32* __restrict__ closers, u32 n_closers, const u32* __restrict__ rk,
                                                                 const u32* __restrict__ prev_owned, const u8* __restrict__ klcp,
                                                                 const u32* __restrict__ sa_d, const u32* __restrict__ limit, const u32* __restrict__ dphrase,
                                                                 const u32* __restrict__ prank, const u32* __restrict__ occ_first,
                                                                 const u64* __restrict__ occ_next_start, DocLookup L, u16* __restrict__ thresh) {
    u32 wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wi >= n_closers) return;                                    // whole warp
    const u32 c = closers[wi];
    const u32 last = prev_owned[c - 1] - 1;                         // last rank of the segment that is longer than w (exists: c closes it)
    const u32 lcp_close = klcp[c];
    u64 bits[W];
#pragma unroll
    for (int w = 0; w < W; ++w) bits[w] = 0;
    u64 anchor = ~0ull; u32 lcp_open = 0, seen = 0;
    // walk the segment from its last rank to its first, 32 ranks at a time
    for (u32 hi = last;;) {
        u32 x = hi >= lane ? hi - lane : 0u;
        bool in_range = hi >= lane;
        u32 w = in_range ? rk[x] : RKW_HB;
        unsigned starts = __ballot_sync(0xffffffffu, (w & RKW_HB) != 0);
        int first_lane = __ffs((int)starts) - 1;                    // nearest segment start at or below hi (-1: further down)
        bool mine = in_range && (w & RKW_OWN) && (first_lane < 0 || (int)lane <= first_lane);
        // metadata of my dictionary suffix
        u32 q0 = 0, cnt = 0, shift = 0;
        if (mine) {
            u32 i = sa_d[x];
            shift = limit[i] - 1 - PFP_W;
            u32 r = prank[dphrase[i]];
            q0 = occ_first[r]; cnt = occ_first[r + 1] - q0;
        }
        unsigned owners = __ballot_sync(0xffffffffu, mine);
        if (owners) {                                               // the lowest rank seen so far that is longer than w opens the segment
            int low = 31 - __clz((int)owners);
            lcp_open = __shfl_sync(0xffffffffu, (u32)klcp[mine ? x : 0], low);
        }
        while (owners) {
            int src = __ffs((int)owners) - 1; owners &= owners - 1;
            u32 mq = __shfl_sync(0xffffffffu, q0, src), mc = __shfl_sync(0xffffffffu, cnt, src), ms = __shfl_sync(0xffffffffu, shift, src);
            for (u32 t = lane; t < mc; t += 32) {
                u64 p = occ_next_start[mq + t] - ms;
                u32 d = L.coarse[p >> L.shift];
                while (L.doc_start[d + 1] <= p) ++d;
                bits[(W == 1) ? 0 : (d >> 6)] |= 1ull << (d & 63);
                if (d == 0) anchor = p;
            }
            seen += mc;
        }
        if (first_lane >= 0 || hi < 32) break;
        hi -= 32;
    }
    u32 distinct = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
#ifdef MMB_SIMT
        u64 bb = bits[w];
        for (u32 dd = 16; dd > 0; dd >>= 1) bb |= __shfl_xor_sync(0xffffffffu, bb, dd);
        distinct += (u32)__popcll(bb);
#else
        distinct += (u32)__popc(__reduce_or_sync(0xffffffffu, (u32)bits[w])) + (u32)__popc(__reduce_or_sync(0xffffffffu, (u32)(bits[w] >> 32)));
#endif
    }
    if (distinct == seen && distinct == L.ndocs && anchor != ~0ull) thresh[anchor] = (u16)min(max(lcp_open, lcp_close), 65535u);
}


// The same for collections of at most 64 documents, one THREAD per closing rank (threads of the other ranks leave at once, so that the
// rank-indexed arrays are read coalesced): the documents of a segment are the union of its phrases' document masks (PhraseDocs, L2
// resident: one record per distinct phrase), and the document-0 suffix sits at the phrase's document-0 occurrence + the offset of the
// dictionary suffix inside the phrase.  Segments that touch a phrase with a repeated document or an occurrence across a document
// boundary go to `slow` (the warp kernel above decides them exactly from the occurrence lists).
__global__ void __launch_bounds__(TPB) segment_thresholds_masks(const u8* __restrict__ mkind, u32 dsize, const u32* __restrict__ rk,
                                                                const u32* __restrict__ prev_owned, const u8* __restrict__ klcp,
                                                                const u32* __restrict__ sa_d, const u32* __restrict__ rank_phrase,
                                                                const PhraseDocs* __restrict__ pdocs, u32 ndocs, u16* __restrict__ thresh,
                                                                u32* __restrict__ slow, u32* __restrict__ n_slow, int list_all) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    bool defer = false;
    if (c < dsize && (mkind[c] & 4)) {
        if (list_all) defer = true;                         // more than 64 documents: every segment goes to the warp kernel
        else {
            u32 x = c - 1;                                  // last rank of the segment that is longer than w: usually the rank before
            if (!(rk[x] & RKW_OWN)) x = prev_owned[c - 1] - 1;
            const u32 lcp_close = klcp[c];
            u64 mask = 0, anchor = ~0ull; u32 lcp_open = 0;
            for (;; --x) {
                u32 w = rk[x];
                if (w & RKW_OWN) {
                    u32 i = sa_d[x];
                    PhraseDocs r = pdocs[rank_phrase[x]];        // rank-indexed: coalesced with rk / sa_d / klcp
                    defer |= (r.flags & (PD_DUPLICATE | PD_CROSSES)) != 0;
                    mask |= r.mask;
                    if (r.flags & PD_HAS_ANCHOR) anchor = r.anchor + (i - r.first);
                    lcp_open = klcp[x];
                }
                if ((w & RKW_HB) || x == 0) break;
            }
            if (!defer && (u32)__popcll(mask) == ndocs && anchor != ~0ull) thresh[anchor] = (u16)min(max(lcp_open, lcp_close), 65535u);
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, defer);
    if (m) {
        u32 lane = threadIdx.x & 31, base = 0;
        int leader = __ffs((int)m) - 1;
        if ((int)lane == leader) base = atomicAdd(n_slow, (u32)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (defer) slow[base + __popc(m & ((1u << lane) - 1))] = c;
    }
}

}  // namespace

void segment_thresholds(const PfpIndex& ix, const DocLookup& docs, u16* thresh, cudaStream_t s) {
    if (!ix.pruned || !ix.n_closers || !thresh) return;
    const bool masks = ix.phrase_docs.p && ix.rank_phrase.p && docs.ndocs <= 64 && !getenv("MMB_THRESH_SLOW");
    DevBuf<u32> slow, n_slow;
    slow.alloc(masks ? std::max<u32>(ix.n_closers / 8, 1u << 16) : ix.n_closers, s); n_slow.alloc(1, s); n_slow.zero();
    // (with masks the deferred segments are the few that touch a document boundary or a phrase repeated inside a document; if they
    // ever outnumber the list, the kernel is simply run again with every segment deferred)
    MMB_LAUNCH(segment_thresholds_masks, ceil_div(ix.dsize, TPB), TPB, 0, s, (const u8*)ix.mkind.p, ix.dsize, (const u32*)ix.rk.p, (const u32*)ix.prev_owned.p,
               (const u8*)ix.klcp.p, (const u32*)ix.sa_d.p, (const u32*)ix.rank_phrase.p, (const PhraseDocs*)ix.phrase_docs.p, docs.ndocs, thresh, slow.p, n_slow.p,
               masks ? 0 : 1);
    u32 count = to_host(n_slow.p, 1, s)[0];
    if (count > slow.n) {
        slow.alloc(ix.n_closers, s); n_slow.zero();
        MMB_LAUNCH(segment_thresholds_masks, ceil_div(ix.dsize, TPB), TPB, 0, s, (const u8*)ix.mkind.p, ix.dsize, (const u32*)ix.rk.p, (const u32*)ix.prev_owned.p,
                   (const u8*)ix.klcp.p, (const u32*)ix.sa_d.p, (const u32*)ix.rank_phrase.p, (const PhraseDocs*)ix.phrase_docs.p, docs.ndocs, thresh, slow.p, n_slow.p, 1);
        count = to_host(n_slow.p, 1, s)[0];
    }
    const u32* list = slow.p;
    if (!count) return;
    u32 grid = ceil_div((u64)count * 32, TPB);
    auto launch = [&](auto segment_thresholds_k) {
        MMB_LAUNCH(segment_thresholds_k, grid, TPB, 0, s, list, count, (const u32*)ix.rk.p, (const u32*)ix.prev_owned.p, (const u8*)ix.klcp.p,
                   (const u32*)ix.sa_d.p, (const u32*)ix.limit.p, (const u32*)ix.dphrase.p, (const u32*)ix.prank.p, (const u32*)ix.occ_first.p,
                   (const u64*)ix.occ_next_start.p, docs, thresh);
    };
    if (docs.ndocs <= 64) launch(segment_thresholds_kernel<1>);
    else if (docs.ndocs <= 256) launch(segment_thresholds_kernel<4>);
    else launch(segment_thresholds_kernel<16>);
}

void stream_suffixes(const u8* text, const PfpIndex& ix, u64 tile_ranks, u32 halo, TileConsumer& consumer, cudaStream_t s, const DocLookup* docs) {
    u32 dsize = ix.dsize;
    PhaseTimer* pt = new PhaseTimer("stream members+bounds", s);
    // ---- members (owned dictionary suffixes in suffix order)
    DevBuf<u32> flag_own, idx; idx.alloc((u64)dsize + 1, s);
    const u32* flag = ix.owned.p;
    if (!flag) {        // alphabets too large for the key-based flag (pfp.cu)
        flag_own.alloc((u64)dsize + 1, s);
        MMB_LAUNCH(owned_flags, ceil_div((u64)dsize + 1, TPB), TPB, 0, s, (const u32*)ix.sa_d.p, (const u32*)ix.limit.p, dsize, flag_own.p);
        flag = flag_own.p;
    }
    prim::exclusive_sum(flag, idx.p, (u64)dsize + 1, s);
    u32 K = to_host(idx.p + dsize, 1, s)[0];
    if (K == 0) { const_cast<PfpIndex&>(ix).stream_ranks = 0; return; }      // pruned route: no segment can hold a match
    DevBuf<u32> mem_i, mem_head, mem_cnt, glcp; DevBuf<u64> out_base; DevBuf<MemberMeta> meta; DevBuf<u8> mem_klcp;
    if (ix.klcp.p) mem_klcp.alloc(K, s);
    meta.alloc(K, s); mem_i.alloc(K, s); mem_head.alloc(K, s); mem_cnt.alloc((u64)K + 1, s); glcp.alloc((u64)K + 1, s); out_base.alloc((u64)K + 1, s);
    MMB_CUDA(cudaMemsetAsync(mem_cnt.p + K, 0, 4, s));
    {
        DevBuf<DictRecord> rec; rec.alloc(dsize, s);
        MMB_LAUNCH(dict_records, ceil_div(dsize, TPB), TPB, 0, s, (const u32*)ix.limit.p, (const u32*)ix.dphrase.p, (const u32*)ix.isa_d.p, (const u32*)ix.doff.p,
                   (const u8*)ix.dict.p, dsize, rec.p);
        MMB_LAUNCH(members_fill, ceil_div(dsize, TPB), TPB, 0, s, (const u32*)ix.sa_d.p, (const DictRecord*)rec.p, (const u32*)ix.prank.p,
                   (const u32*)ix.occ_first.p, flag, (const u32*)idx.p, dsize, (const u8*)ix.klcp.p, (const u32*)ix.ghead.p, (const u8*)ix.mkind.p, mem_i.p, mem_head.p,
                   mem_cnt.p, meta.p, mem_klcp.p);
    }
    flag_own.release(); idx.release();
    prim::exclusive_sum(mem_cnt.p, out_base.p, (u64)K + 1, s);
    u64 total = to_host(out_base.p + K, 1, s)[0];
    const_cast<PfpIndex&>(ix).stream_ranks = total;
    if (getenv("MMB_DEBUG")) fprintf(stderr, "[stream] members %u, ranks %llu of %llu\n", K, (unsigned long long)total, (unsigned long long)ix.n);
    if (tile_ranks >= (1ull << 62) && total >= (1ull << 32))
        throw Error(31, "unbounded total frequency (-f 0 with -F 0): the pruned suffix stream still has " + std::to_string(total) + " ranks, too many for one view");
    if ((!ix.pruned || ix.full_stream) && total != ix.n) throw Error(76, "internal: the parse does not cover the text (" + std::to_string(total) + " of " + std::to_string(ix.n) + ")");
    MMB_LAUNCH(group_lcp, ceil_div(K, TPB), TPB, 0, s, (const u8*)ix.dict.p, (const MemberMeta*)meta.p, (const u32*)mem_i.p, (const u32*)mem_head.p, (const u8*)mem_klcp.p, ix.key_symbols, K, glcp.p);
    mem_klcp.release();
    mem_head.release();
    MMB_LAUNCH(group_shape, ceil_div(K, TPB), TPB, 0, s, (const u32*)glcp.p, K, meta.p);

    // ---- tile boundaries on group starts
    u32 nt = (u32)((total + tile_ranks - 1) / tile_ranks);
    DevBuf<u32> d_bounds; d_bounds.alloc(nt + 1, s);
    MMB_LAUNCH(tile_bounds, ceil_div(nt + 1, 128), 128, 0, s, (const u64*)out_base.p, (const u32*)glcp.p, K, tile_ranks, nt, d_bounds.p);
    std::vector<u32> raw_bounds = to_host(d_bounds.p, nt + 1, s);
    std::vector<u32> bounds;
    for (u32 b : raw_bounds) if (bounds.empty() || b > bounds.back()) bounds.push_back(b);
    if (bounds.empty() || bounds[0] != 0) bounds.insert(bounds.begin(), 0);
    std::vector<u64> h_base(bounds.size());
    for (size_t t = 0; t < bounds.size(); ++t) MMB_CUDA(cudaMemcpyAsync(&h_base[t], out_base.p + bounds[t], 8, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaStreamSynchronize(s));
    u64 max_new = 0;
    for (size_t t = 0; t + 1 < bounds.size(); ++t) max_new = std::max(max_new, h_base[t + 1] - h_base[t]);

    delete pt;
    // ---- tile buffers: [halo | new ranks | 64 slack]; the first tile starts with the sentinel rank instead of a halo
    u64 cap = (u64)halo + max_new + 1;
    DevBuf<u64> sa, halo_sa, small, large, tiny; DevBuf<u32> lcp, key, halo_lcp, counts; DevBuf<u8> bwt, halo_bwt;
    sa.alloc(cap, s); lcp.alloc(cap + 64, s); key.alloc(cap, s); bwt.alloc(cap + 64, s); halo_sa.alloc(halo + 1, s); halo_lcp.alloc(halo + 1, s);
    halo_bwt.alloc(halo + 1, s); counts.alloc(3, s);
    DevBuf<u8> doc;
    if (docs) doc.alloc(cap, s);
    const int expand_variant = getenv("MMB_EXPAND_VARIANT") ? atoi(getenv("MMB_EXPAND_VARIANT")) : 0;
    u64 rank_next = 0;                  // global rank of the first new entry of the coming tile
    u64 halo_have = 0;
    for (size_t t = 0; t + 1 < bounds.size(); ++t) {
        u32 k0 = bounds[t], k1 = bounds[t + 1];
        bool first_tile = (t == 0), last_tile = (t + 2 == bounds.size());
        PhaseTimer* tt = new PhaseTimer("stream tile: enumerate", s);
        u64 fresh = h_base[t + 1] - h_base[t];
        u64 lead = first_tile ? 1 : halo_have;          // view index of the first member-generated entry
        if (!first_tile && halo_have) {
            MMB_CUDA(cudaMemcpyAsync(sa.p, halo_sa.p, halo_have * 8, cudaMemcpyDeviceToDevice, s));
            MMB_CUDA(cudaMemcpyAsync(lcp.p, halo_lcp.p, halo_have * 4, cudaMemcpyDeviceToDevice, s));
            MMB_CUDA(cudaMemcpyAsync(bwt.p, halo_bwt.p, halo_have, cudaMemcpyDeviceToDevice, s));
        }
        u32 nm = k1 - k0;
        small.alloc((u64)nm / 2 + 1, s); large.alloc((u64)nm / 2 + 1, s); tiny.alloc((u64)nm / 2 + 1, s); counts.zero();
        {
            auto launch_expand = [&](auto expand_members_k) {
                MMB_LAUNCH(expand_members_k, ceil_div(nm, EXP_TPB), EXP_TPB, 0, s, k0, k1, (const MemberMeta*)meta.p, (const u64*)out_base.p, (const u32*)glcp.p,
                           (const u32*)ix.occ_next_rank.p, (const u64*)ix.occ_next_start.p, (const u32*)ix.occ_lcp.p, (const u32*)ix.s_lcp.p,
                           (const u32*)ix.rmq.p, ix.rmq_blocks, text, ix.n, lead, sa.p, lcp.p, key.p, bwt.p);
            };
            switch (expand_variant) {
                case 1: launch_expand(expand_members<8, 0>); break;
                case 2: launch_expand(expand_members<6, 1>); break;
                case 4: launch_expand(expand_members<8, 2>); break;
                case 5: launch_expand(expand_members<6, 0>); break;
                default: launch_expand(expand_members<8, 1>); break;
            }
        }
        MMB_LAUNCH(collect_multi, ceil_div(nm, TPB), TPB, 0, s, k0, k1, K, (const u32*)glcp.p, (const u64*)out_base.p, tiny.p, small.p, large.p, counts.p);
        std::vector<u32> c = to_host(counts.p, 3, s);
        if (getenv("MMB_DEBUG")) fprintf(stderr, "[stream] tile members %u fresh ranks %llu: multi-member groups by size: <=256: %u, <=2048: %u, larger: %u\n", nm, (unsigned long long)fresh, c[2], c[0], c[1]);
        if (c[2]) {         // two instances: up to 64 suffixes (2 per lane) and up to 256 (8 per lane); each skips the groups of the other
            auto merge_warp_2 = merge_warp<2, 0>; auto merge_warp_8 = merge_warp<8, 64>;
            MMB_LAUNCH(merge_warp_2, ceil_div((u64)c[2] * 32, TPB), TPB, 0, s, (const u64*)tiny.p, c[2], k0, (const u64*)out_base.p, (const u32*)glcp.p, (const u32*)mem_i.p,
                       (const u32*)ix.limit.p, (const u32*)ix.s_lcp.p, (const u32*)ix.rmq.p, ix.rmq_blocks, lead, sa.p, lcp.p, (const u32*)key.p, bwt.p);
            MMB_LAUNCH(merge_warp_8, ceil_div((u64)c[2] * 32, TPB), TPB, 0, s, (const u64*)tiny.p, c[2], k0, (const u64*)out_base.p, (const u32*)glcp.p, (const u32*)mem_i.p,
                       (const u32*)ix.limit.p, (const u32*)ix.s_lcp.p, (const u32*)ix.rmq.p, ix.rmq_blocks, lead, sa.p, lcp.p, (const u32*)key.p, bwt.p);
        }
        if (c[0]) {
            MMB_LAUNCH(merge_small, c[0], TPB, 0, s, (const u64*)small.p, k0, (const u64*)out_base.p, (const u32*)glcp.p, (const u32*)mem_i.p,
                       (const u32*)ix.limit.p, (const u32*)ix.s_lcp.p, (const u32*)ix.rmq.p, ix.rmq_blocks, lead, sa.p, lcp.p, (const u32*)key.p, bwt.p);
        }
        if (c[1]) {
            u32 nl = c[1];
            DevBuf<u64> sizes, fbase, fk_a, fk_b, fv_a, fv_b;
            sizes.alloc(nl + 1, s); fbase.alloc(nl + 1, s);
            MMB_LAUNCH(large_sizes, ceil_div(nl + 1, TPB), TPB, 0, s, (const u64*)large.p, nl, (const u64*)out_base.p, sizes.p);
            prim::exclusive_sum(sizes.p, fbase.p, nl + 1, s);
            u64 ftotal = to_host(fbase.p + nl, 1, s)[0];
            fk_a.alloc(ftotal, s); fk_b.alloc(ftotal, s); fv_a.alloc(ftotal, s); fv_b.alloc(ftotal, s);
            dim3 grid(nl, 32);
            MMB_LAUNCH(large_gather, grid, TPB, 0, s, (const u64*)large.p, (const u64*)fbase.p, k0, (const u64*)out_base.p, lead, (const u64*)sa.p,
                       (const u32*)key.p, fk_a.p, fv_a.p);
            int seq_bits = 1; while ((1ull << seq_bits) < (u64)nl) ++seq_bits;      // keys are (group number << 32 | rank): the bits above are zero
            prim::sort_pairs64(fk_a.p, fk_b.p, fv_a.p, fv_b.p, ftotal, 0, 32 + seq_bits, s);
            MMB_LAUNCH(large_scatter, ceil_div(ftotal, TPB), TPB, 0, s, (const u64*)large.p, (const u64*)fbase.p, ftotal, k0, (const u64*)out_base.p,
                       (const u32*)glcp.p, (const u32*)mem_i.p, (const u32*)ix.limit.p, (const u32*)ix.s_lcp.p, (const u32*)ix.rmq.p, ix.rmq_blocks, lead,
                       (const u64*)fk_b.p, (const u64*)fv_b.p, sa.p, lcp.p);
            dim3 gb(nl, 32);
            MMB_LAUNCH(bwt_of_groups, gb, TPB, 0, s, (const u64*)large.p, k0, (const u64*)out_base.p, lead, text, (const u64*)sa.p, bwt.p);
        }
        u64 count = lead + fresh;
        MMB_LAUNCH(tile_edges, 1, 64, 0, s, sa.p, lcp.p, bwt.p, text, ix.n, (int)first_tile, count);
        StreamTile tile;
        tile.rank0 = first_tile ? 0 : rank_next - halo_have;
        tile.count = count; tile.first = first_tile ? 0 : halo_have; tile.sa = sa.p; tile.lcp = lcp.p; tile.bwt = bwt.p; tile.last = last_tile;
        if (docs) {
            MMB_LAUNCH(tile_docs, ceil_div(count, TPB), TPB, 0, s, (const u64*)sa.p, count, *docs, ix.n, doc.p);
            tile.doc = doc.p;
        }
        delete tt;
        { PhaseTimer ct("stream tile: consume", s); consumer.consume(tile); }
        // stash the tail as the next tile's halo
        halo_have = std::min<u64>(halo, count);
        MMB_CUDA(cudaMemcpyAsync(halo_sa.p, sa.p + (count - halo_have), halo_have * 8, cudaMemcpyDeviceToDevice, s));
        MMB_CUDA(cudaMemcpyAsync(halo_lcp.p, lcp.p + (count - halo_have), halo_have * 4, cudaMemcpyDeviceToDevice, s));
        MMB_CUDA(cudaMemcpyAsync(halo_bwt.p, bwt.p + (count - halo_have), halo_have, cudaMemcpyDeviceToDevice, s));
        rank_next = (first_tile ? 1 : rank_next) + fresh;
    }
}

}  // namespace mmb
