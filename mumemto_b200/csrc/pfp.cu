// mumemto_b200/csrc/pfp.cu -- builds the prefix-free-parse index of the text on the device (see pfp.h for the construction).
//
// Stages (all kernels hand-written; the device-wide sorts and scans come from prims.h):
//   1 trigger scan            2 passes over T (count, then write): every position hashes its own 10-byte window
//   2 phrase identity         one 64-bit content hash per phrase, radix sort by hash, byte-wise verification against the group head
//                             (a hash collision between different phrases is detected there and the stage is retried with a new seed)
//   3 dictionary D'           distinct phrases + terminators; per-position `limit` (distance to the terminator) and phrase id
//   4 suffix order of D'      shared doubling engine (sa.cu) in "limit" mode: equal dictionary suffixes stay tied on purpose
//   5 phrase ranks -> parse   rank of the whole-phrase suffixes; parse = rank of every phrase occurrence
//   6 suffix order of parse   same engine on the integer alphabet
//   7 s_lcp                   text-level LCP of rank-adjacent parse suffixes: chunked Kasai over phrases + one byte comparison of the
//                             two differing phrases
//   8 range-minimum index over s_lcp (block minima + sparse table), occurrence lists of every phrase ordered by the rank of the parse
//     suffix that follows, with the text LCP between consecutive occurrences
#include "pfp.h"
#include "pfp_device.cuh"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 TPB = 256;
constexpr u32 FULL = 0xffffffffu;
constexpr u32 SCAN_BLOCK = 4096;        // text positions per block in the trigger passes (16 per thread)

__device__ __forceinline__ u64 load8(const u8* __restrict__ t, u64 i) {      // unaligned 8-byte little-endian load (reads <= 11 bytes past i)
    const u32* w = reinterpret_cast<const u32*>(t + (i & ~3ull));
    u32 sh = (u32)(i & 3) * 8;
    u32 a = w[0], b = w[1], c = w[2];
    u32 lo = sh ? __funnelshift_r(a, b, sh) : a;
    u32 hi = sh ? __funnelshift_r(b, c, sh) : b;
    return ((u64)hi << 32) | lo;
}
__host__ __device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
    return x;
}
// Trigger rule for the window T[e-9 .. e].  Any deterministic function of the ten bytes is a valid rule (the outputs of the path do not
// depend on it).  This one is a polynomial hash modulo 2^32 that ROLLS from one window to the next inside a thread
// (two multiply-adds per position instead of hashing ten bytes afresh: the first version cost ~28 instructions per position and the
// kernel was issue-bound at 1.08 TB/s), finished by one xor-shift and one multiply; the window triggers when the result falls into the
// lowest 1/PFP_P of the 32-bit range.
constexpr u32 TRIG_B = 0x9E3779B1u;
constexpr u32 trig_pow(u32 e) { u32 r = 1; for (u32 i = 0; i < e; ++i) r *= TRIG_B; return r; }
constexpr u32 TRIG_B10 = trig_pow(PFP_W);
constexpr u32 TRIG_LIMIT = 0xffffffffu / PFP_P;
__device__ __forceinline__ bool hash_triggers(u32 h) { h ^= h >> 15; h *= 0x2C1B3C6Du; return h < TRIG_LIMIT; }

// exclusive block scan of one u32 per thread; *total = block sum
__device__ __forceinline__ u32 block_scan(u32 mine, u32* total) {
    __shared__ u32 warp_sums[TPB / 32];
    u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, inc = mine;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    u32 before = 0, all = 0;
    for (u32 w = 0; w < TPB / 32; ++w) { if (w < warp) before += warp_sums[w]; all += warp_sums[w]; }
    *total = all;
    __syncthreads();
    return before + inc - mine;
}

// Pass 1: every thread reads the bytes around its 16 positions once (the window's history rounded up to whole words + one 128-bit load),
// tests the 16 windows from registers and keeps the result as a 16-bit mask; the block's trigger count goes to block_count.
constexpr u32 TRIG_HIST = (PFP_W - 1 + 3) / 4 * 4;      // bytes of history before the thread's first position, a multiple of 4
constexpr u32 TRIG_WORDS = TRIG_HIST / 4 + 4;
__global__ void __launch_bounds__(TPB) trigger_masks(const u8* __restrict__ text, u64 n, u32 block0, u16* __restrict__ masks, u32* __restrict__ block_count) {
    const u32 blk = blockIdx.x + block0;                // blocks below block0 were scanned ahead of the parse (PfpTriggers)
    u64 e0 = (u64)blk * SCAN_BLOCK + (u64)threadIdx.x * 16;
    u32 mask = 0;
    if (e0 < n) {
        u32 w[TRIG_WORDS];                              // bytes e0 - TRIG_HIST .. e0 + 15 (the text allocation is 16-byte aligned with 64 bytes of slack)
        if (e0 >= TRIG_HIST) {
            const u32* q = reinterpret_cast<const u32*>(text + e0 - TRIG_HIST);
#pragma unroll
            for (u32 i = 0; i < TRIG_HIST / 4; ++i) w[i] = q[i];
        } else {                                        // the first thread(s) of the text: positions before 0 read as zero
#pragma unroll
            for (u32 i = 0; i < TRIG_HIST / 4; ++i) {
                u32 v = 0;
#pragma unroll
                for (u32 b = 0; b < 4; ++b) { u64 at = e0 + 4 * i + b; if (at >= TRIG_HIST) v |= (u32)text[at - TRIG_HIST] << (8 * b); }
                w[i] = v;
            }
        }
        uint4 c = *reinterpret_cast<const uint4*>(text + e0);
        w[TRIG_HIST / 4] = c.x; w[TRIG_HIST / 4 + 1] = c.y; w[TRIG_HIST / 4 + 2] = c.z; w[TRIG_HIST / 4 + 3] = c.w;
#define TRIG_BYTE(b) ((w[(b) >> 2] >> (((b) & 3) * 8)) & 0xffu)
        // T[e0+k] is byte TRIG_HIST + k of w[]; its window starts PFP_W - 1 bytes earlier
        constexpr u32 first = TRIG_HIST - (PFP_W - 1);
        u32 h = 0;
#pragma unroll
        for (u32 i = 0; i < PFP_W; ++i) h = h * TRIG_B + TRIG_BYTE(first + i);
#pragma unroll
        for (u32 k = 0; k < 16; ++k) {
            u64 e = e0 + k;
            if (e >= PFP_W && e < n && hash_triggers(h)) mask |= 1u << k;
            if (k < 15) h = h * TRIG_B + TRIG_BYTE(first + PFP_W + k) - TRIG_BYTE(first + k) * TRIG_B10;
        }
#undef TRIG_BYTE
    }
    masks[(u64)blk * TPB + threadIdx.x] = (u16)mask;
    u32 total; block_scan((u32)__popc(mask), &total);
    if (threadIdx.x == 0) block_count[blk] = total;
}

// Pass 2: positions of the set bits, in text order.
__global__ void __launch_bounds__(TPB) trigger_positions(const u16* __restrict__ masks, const u32* __restrict__ block_base, u64* __restrict__ ends) {
    u64 e0 = (u64)blockIdx.x * SCAN_BLOCK + (u64)threadIdx.x * 16;
    u32 mask = masks[(u64)blockIdx.x * TPB + threadIdx.x];
    u32 total, before = block_scan((u32)__popc(mask), &total);
    u32 at = block_base[blockIdx.x] + before;
    while (mask) { u32 k = (u32)__ffs((int)mask) - 1; mask &= mask - 1; ends[at++] = e0 + k; }
}

__global__ void __launch_bounds__(TPB) phrase_starts(const u64* __restrict__ ends, u32 np, u64 n, u64* __restrict__ pstart) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > np) return;
    pstart[j] = (j == 0) ? 0 : ((j == np) ? n : ends[j - 1] - PFP_W + 1);
}

// phrase j = T'[pstart[j] .. pstart[j+1] + w - 1], length pstart[j+1] - pstart[j] + w; bytes at or past n read as 0 (text has zero slack)
// Content hash: ONE thread per phrase, 8 bytes at a time straight from global memory, a full 64-bit finaliser per word (the hash is a
// sum of per-word mixes).  Measured alternatives at C2 (this version: 12.7 ms): staging the block's text span in shared memory 15.6 ms;
// 8 lanes per phrase on 8-byte words 14.9 ms; 8 lanes on 16-byte chunks with a cheaper two-multiply mix ~13.5 ms AND structurally weak
// (differences in a chunk's last bytes only travelled upward: it collided on every seed at 4 million distinct phrases --
// scripts/phrase_hash_quality.py reproduces that on the CPU).  The per-phrase overhead of the cooperative variants eats what their
// coalescing wins; the verification below does profit from 8 lanes per phrase (13.8 -> 12.5 ms).
constexpr u32 PH_LANES = 8;
__global__ void __launch_bounds__(TPB) phrase_hash(const u8* __restrict__ text, const u64* __restrict__ pstart, u32 np, u64 seed, u64* __restrict__ keys,
                                                   u32* __restrict__ vals) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= np) return;
    u64 s = pstart[j], len = pstart[j + 1] - s + PFP_W;
    u64 h = seed ^ mix64(len);
    u64 k = 0;
    for (; (k + 1) * 8 <= len; ++k) h += mix64(load8(text, s + k * 8) + (k + 1) * 0x9E3779B97F4A7C15ull + seed);
    u32 rem = (u32)(len - k * 8);
    if (rem) { u64 wd = load8(text, s + k * 8) & ((1ull << (8 * rem)) - 1); h += mix64(wd + (k + 1) * 0x9E3779B97F4A7C15ull + seed); }
    keys[j] = h; vals[j] = j;
}

__global__ void __launch_bounds__(TPB) hash_boundaries(const u64* __restrict__ keys, u32 np, u32* __restrict__ flag) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < np) flag[q] = (q == 0 || keys[q] != keys[q - 1]) ? 1u : 0u;
}

// after the inclusive numbering of groups: representative (first member) per group, provisional id per phrase, byte-wise verification
__global__ void __launch_bounds__(TPB) assign_groups(const u32* __restrict__ flag, const u32* __restrict__ gnum_excl, const u32* __restrict__ order, u32 np,
                                                     u32* __restrict__ pid, u32* __restrict__ rep) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= np) return;
    u32 g = gnum_excl[q] + flag[q] - 1;
    pid[order[q]] = g;
    if (flag[q]) rep[g] = order[q];
}
__global__ void __launch_bounds__(TPB) verify_groups(const u8* __restrict__ text, const u64* __restrict__ pstart, const u32* __restrict__ pid,
                                                     const u32* __restrict__ rep, u32 np, u32* __restrict__ mismatch) {
    u32 j = (u32)(((u64)blockIdx.x * blockDim.x + threadIdx.x) / PH_LANES), lane = threadIdx.x % PH_LANES;      // PH_LANES lanes compare one phrase with its head
    if (j >= np) return;
    u32 r = rep[pid[j]];
    if (r == j) return;
    u64 a = pstart[j], b = pstart[r], la = pstart[j + 1] - a + PFP_W, lb = pstart[r + 1] - b + PFP_W;
    bool bad = la != lb;
    u64 full = la >> 3;
    for (u64 k = lane; !bad && k < full; k += PH_LANES) bad = load8(text, a + k * 8) != load8(text, b + k * 8);
    u32 rem = (u32)(la & 7);
    if (!bad && rem && (full % PH_LANES) == lane) { u64 msk = (1ull << (8 * rem)) - 1; bad = (load8(text, a + full * 8) & msk) != (load8(text, b + full * 8) & msk); }
    if (bad) *mismatch = 1;
}

__global__ void __launch_bounds__(TPB) dict_lengths(const u64* __restrict__ pstart, const u32* __restrict__ rep, u32 nd, u64* __restrict__ dlen) {
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nd) { u32 j = rep[g]; dlen[g] = pstart[j + 1] - pstart[j] + PFP_W + 1; }      // + terminator
    if (g == nd) dlen[g] = 0;
}
__global__ void narrow_offsets(const u64* __restrict__ in, u32 n, u32* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u32)in[i];
}
// one warp per distinct phrase: copy its bytes, append the terminator, fill limit[] and dphrase[]
__global__ void __launch_bounds__(TPB) dict_fill(const u8* __restrict__ text, u64 n, const u64* __restrict__ pstart, const u32* __restrict__ rep,
                                                 const u32* __restrict__ doff, u32 nd, u8* __restrict__ dict, u32* __restrict__ limit,
                                                 u32* __restrict__ dphrase) {
    u32 g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= nd) return;
    u64 s = pstart[rep[g]];
    u32 o = doff[g], total = doff[g + 1] - o;       // phrase length + 1
    for (u32 t = lane; t < total; t += 32) {
        u64 p = s + t;
        dict[o + t] = (t + 1 < total && p < n) ? text[p] : (u8)0;
        limit[o + t] = total - t;
        dphrase[o + t] = g;
    }
}

__global__ void __launch_bounds__(TPB) dict_histogram(const u8* __restrict__ d, u32 n, u32* __restrict__ hist) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    u32 stride = gridDim.x * blockDim.x;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&h[d[i]], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}
__global__ void __launch_bounds__(TPB) dict_keys(const u8* __restrict__ d, const u32* __restrict__ limit, u32 n, const u8* __restrict__ lut, u32 bits,
                                                 u32 cpk, u64* __restrict__ keys, u32* __restrict__ vals) {
    __shared__ u8 code[256];
    code[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 lim = limit[i];
    u64 key = 0;
    for (u32 t = 0; t < cpk; ++t) key = (key << bits) | (u64)(t < lim ? code[d[i + t]] : 0);
    keys[i] = key; vals[i] = i;
}

// Pair layout of the dictionary's rank array for the doubling rounds (sa.cu, Isa): word 2i+1 = offset of position i inside its phrase.
__global__ void __launch_bounds__(TPB) phrase_offsets(const u32* __restrict__ limit, const u32* __restrict__ dphrase, const u32* __restrict__ doff, u32 dsize,
                                                      u32* __restrict__ isa2) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dsize) return;
    u32 g = dphrase[i];
    isa2[2 * (u64)i + 1] = (doff[g + 1] - doff[g]) - limit[i];
}
__global__ void __launch_bounds__(TPB) unpair_ranks(const u32* __restrict__ isa2, u32 dsize, u32* __restrict__ isa) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dsize) isa[i] = isa2[2 * (u64)i];
}

// A dictionary suffix is "owned" (longer than w) iff symbol w of its initial sort key is a real symbol (codes are >= 1, the
// terminator and everything after it is 0).  All suffixes that share a key agree on this, and later rounds only reorder suffixes
// inside groups of equal keys, so the flag of a RANK is final right after the initial sort -- no random access needed.
__global__ void __launch_bounds__(TPB) owned_from_keys(const u64* __restrict__ sorted_keys, u32 dsize, u32 bits, u32 cpk, u32* __restrict__ flag,
                                                       u8* __restrict__ klcp) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < dsize) {
        u64 key = sorted_keys[p];
        flag[p] = ((u32)(key >> (bits * (cpk - 1 - PFP_W))) & ((1u << bits) - 1)) ? 1u : 0u;
        // symbols this key shares with the previous rank's (the same for every later order of the ranks inside a group of equal keys)
        u32 shared = 0;
        if (p) { u64 x = key ^ sorted_keys[p - 1]; shared = x ? cpk - 1 - (u32)(63 - __clzll((long long)x)) / bits : cpk; }
        klcp[p] = (u8)shared;
    }
    if (p == dsize) flag[p] = 0;
}

// The one exception: the last phrase ends with w padding zeros that ARE symbols of its string, so its last w owned suffixes show a
// zero at symbol w.  Their rank is looked up through isa_d: that is the head rank of their group, and because a padding zero and a
// terminator are the same key symbol such a suffix can share its group with shorter strings that end where its padding starts
// (e.g. "$" + padding with the "$" that ends some other phrase) -- so walk the group to the rank that really holds the suffix.
__global__ void owned_fix_last_phrase(const u32* __restrict__ pid, u32 np, const u32* __restrict__ doff, const u32* __restrict__ limit,
                                      const u32* __restrict__ isa_d, const u32* __restrict__ sa_d, u32 dsize, u32* __restrict__ flag) {
    u32 g = pid[np - 1];
    for (u32 i = doff[g] + threadIdx.x; i < doff[g + 1]; i += blockDim.x) {
        u32 lim = limit[i];
        if (lim >= PFP_W + 2 && lim <= 2 * PFP_W + 1) {
            u32 q = isa_d[i];
            while (q < dsize && sa_d[q] != i) ++q;
            if (q < dsize) flag[q] = 1;
        }
    }
}


// ================================================================================================ pruned route: rank words and segments
// rank word (one u32 per rank of the initial key order):
//   bit 31      the rank starts a segment (shares fewer than t key symbols with the rank before it)
//   bit 30      the suffix is longer than w (it owns text suffixes)
//   bit 29      the suffix is shorter than t symbols (its text suffixes may share fewer than t symbols with each other)
//   bits 28..9  occurrences of its phrase (saturating)
//   bits 8..0   symbol that precedes the suffix inside its phrase, 0x100 = the suffix is a whole phrase
constexpr u32 RK_HB = 1u << 31, RK_OWN = 1u << 30, RK_SHORT = 1u << 29, RK_CNT_MAX = (1u << 20) - 1, RK_START = 0x100u;
constexpr u32 SEG_LIMIT = 128;           // segments of more ranks than this are kept as they are (low-complexity repeats)
constexpr u8 CLS_STREAM = 1, CLS_REFINE = 2, CLS_THRESH = 4, CLS_HAS_START = 8;

__device__ __forceinline__ u32 key_shared(u64 a, u64 b, u32 bits, u32 cpk) {
    u64 x = a ^ b;
    return x ? cpk - 1 - (u32)(63 - __clzll((long long)x)) / bits : cpk;
}

__global__ void __launch_bounds__(TPB) phrase_counts(const u32* __restrict__ pid, u32 np, u32* __restrict__ count) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < np) atomicAdd(&count[pid[j]], 1u);
}
__global__ void __launch_bounds__(TPB) dict_info(const u32* __restrict__ dphrase, const u32* __restrict__ doff, const u8* __restrict__ dict,
                                                 const u32* __restrict__ count, u32 dsize, u64* __restrict__ pinfo) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dsize) return;
    u32 g = dphrase[i];
    u32 before = i > doff[g] ? (u32)dict[i - 1] : RK_START;
    pinfo[i] = ((u64)g << 32) | (min(count[g], RK_CNT_MAX) << 9) | before;      // phrase id rides along: the gather in rank_words fetches it for free
}
__global__ void __launch_bounds__(TPB) rank_words(const u64* __restrict__ keys, const u32* __restrict__ sa_d, const u64* __restrict__ pinfo, u32 dsize,
                                                  u32 bits, u32 cpk, u32 t, u32* __restrict__ rk, u32* __restrict__ rank_phrase) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dsize) return;
    u64 key = keys[p];
    u64 info = pinfo[sa_d[p]];
    u32 w = (u32)info;
    if (rank_phrase) rank_phrase[p] = (u32)(info >> 32);        // valid as long as the rank's key group is not refined (threshold-only segments never are)
    if ((u32)(key >> (bits * (cpk - 1 - PFP_W))) & ((1u << bits) - 1)) w |= RK_OWN;
    if (((u32)(key >> (bits * (cpk - t))) & ((1u << bits) - 1)) == 0) w |= RK_SHORT;           // symbol t - 1 is the terminator or past it
    if (p == 0 || key_shared(keys[p - 1], key, bits, cpk) < t) w |= RK_HB;
    rk[p] = w;
}
// the last phrase's w padding zeros are symbols of its string: its last w suffixes that are longer than w look short in their keys
// (see owned_fix_last_phrase); one block per such suffix finds its rank among the equal keys and sets the bit
__global__ void __launch_bounds__(TPB) rank_words_fix_last_phrase(const u32* __restrict__ pid, u32 np, const u32* __restrict__ doff, const u32* __restrict__ limit,
                                                                  const u8* __restrict__ dict, const u8* __restrict__ lut, u32 bits, u32 cpk,
                                                                  const u64* __restrict__ keys, const u32* __restrict__ sa_d, u32 dsize,
                                                                  u32* __restrict__ rk) {
    u32 g = pid[np - 1];
    u32 i = doff[g + 1] - 1 - (PFP_W + 1) - blockIdx.x;                // positions with limit = w + 2 + blockIdx.x
    if (i < doff[g] || i >= doff[g + 1]) return;
    u32 lim = limit[i];
    if (lim < PFP_W + 2 || lim > 2 * PFP_W + 1) return;
    u64 key = 0;
    for (u32 t = 0; t < cpk; ++t) key = (key << bits) | (u64)(t < lim ? lut[dict[i + t]] : 0);
    u32 lo = 0, hi = dsize;                                             // first rank whose key is >= key
    while (lo < hi) { u32 mid = lo + ((hi - lo) >> 1); if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    for (u32 q = lo + threadIdx.x; q < dsize && keys[q] == key; q += blockDim.x)
        if (sa_d[q] == i) rk[q] |= RK_OWN;
}

// Class of every rank that is longer than w: what its segment needs (CLS_*).  The block stages the rank words of its 256 ranks plus
// SEG_LIMIT ranks of context on both sides, as 17 windows of 32 entries.  No walks: every warp computes, for its windows, the ballot of
// segment starts and the warp prefix sums of two packed aggregates --
//   A: occurrences | whole phrases | short suffixes | saturated counts          B: suffixes n | sum of preceding symbols | sum of squares
// (all preceding symbols equal  <=>  n * sum(x^2) == (sum x)^2) -- and leaves them in shared memory with the window totals; a rank then
// finds the bounds of its segment by bit scans over the start bits and reads its aggregates as differences of prefix sums (plus whole
// window totals when the segment spans windows).  History: every rank walking its segment (18 warp instructions per dictionary position:
// a warp walked as far as its longest segment, twice) 14.7 ms at C2; a first walk-free variant with BLOCK-wide 64-bit scans and their
// barriers 20.2 ms; warp prefix sums with a fallback walk for segments crossing a 32-rank window 13.2 ms (about half of the segments
// cross one: with 64 haplotypes a segment is 10..30 ranks long).
__device__ __forceinline__ u8 segment_class(u64 total, bool uniform, bool has_start, bool has_short, bool saturated, const PfpPrune& prune, u32 seg_is_interval) {
    const u8 general = CLS_STREAM | CLS_REFINE;
    u8 c = 0;
    if (saturated) c = general;
    else if (total >= prune.size_lo) {
        if (!uniform) c = general;
        // -M: a segment of exactly N suffixes, all at least t = -l symbols long, is one interval of length >= -l
        else if (prune.merge) c = (seg_is_interval && !has_short && total == prune.ndocs) ? CLS_THRESH : general;
    }
    if (has_start) c |= CLS_REFINE | CLS_HAS_START;             // phrase ranks need the exact place of every whole phrase
    return c;
}
__global__ void __launch_bounds__(TPB) classify_segments(const u32* __restrict__ rk, u32 dsize, PfpPrune prune, u32 seg_is_interval, u8* __restrict__ cls) {
    constexpr u32 SPAN = TPB + 2 * SEG_LIMIT + 1;               // 513 staged ranks: tile index y <-> rank base + y
    constexpr u32 NWIN = (SPAN + 31) / 32;                      // 17 windows of 32 tile entries
    __shared__ u32 tile[SPAN];
    __shared__ u64 pa[NWIN * 32], pb[NWIN * 32];                // inclusive prefix sums of A / B inside each window
    __shared__ u64 wa[NWIN], wb[NWIN];                          // window totals
    __shared__ u32 hbm[NWIN + 1];                               // segment-start bits of each window; [NWIN] = all ones
    const i64 base = (i64)blockIdx.x * TPB - SEG_LIMIT;
    for (u32 y = threadIdx.x; y < SPAN; y += TPB) {
        i64 q = base + y;
        tile[y] = (q >= 0 && q < (i64)dsize) ? rk[q] : RK_HB;       // outside the array: a segment start that owns nothing
    }
    __syncthreads();
    const u8 all = CLS_STREAM | CLS_REFINE | CLS_HAS_START;         // without a verdict: keep, and assume the segment may hold a whole phrase
    const u32 p = blockIdx.x * TPB + threadIdx.x, x = threadIdx.x + SEG_LIMIT, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!prune.enabled) { if (p < dsize) cls[p] = (tile[x] & RK_OWN) ? all : (u8)0; return; }       // uniform over the grid
    // ---- phase 1: per window of 32 entries the start bits and the warp prefix sums of
    //   A: occurrences | whole phrases << 32 | short suffixes << 42 | saturated counts << 52      B: n | sum x << 10 | sum x^2 << 28
    for (u32 win = warp; win < NWIN; win += TPB / 32) {
        const u32 y = win * 32 + lane;
        const u32 v = y < SPAN ? tile[y] : RK_HB;
        const unsigned hb = __ballot_sync(FULL, (v & RK_HB) != 0);
        u64 sa = 0, sb = 0;
        if (v & RK_OWN) {
            u32 c = (v >> 9) & RK_CNT_MAX, sym = v & 0x1ffu;
            bool start = sym == RK_START;
            sa = (u64)c | ((u64)(start ? 1 : 0) << 32) | ((u64)((v & RK_SHORT) ? 1 : 0) << 42) | ((u64)(c == RK_CNT_MAX ? 1 : 0) << 52);
            if (!start) sb = 1ull | ((u64)sym << 10) | ((u64)(sym * sym) << 28);
        }
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) {
            u64 oa = __shfl_up_sync(FULL, sa, d), ob = __shfl_up_sync(FULL, sb, d);
            if (lane >= d) { sa += oa; sb += ob; }
        }
        pa[y] = sa; pb[y] = sb;
        if (lane == 31) { wa[win] = sa; wb[win] = sb; }
        if (lane == 0) hbm[win] = hb;
    }
    if (threadIdx.x == 0) hbm[NWIN] = 0xffffffffu;
    __syncthreads();
    if (p >= dsize) return;
    if (!(tile[x] & RK_OWN)) { cls[p] = 0; return; }
    // ---- phase 2: bounds of the rank's segment by bit scans over the start bits, its aggregates from the prefix sums
    const u32 kx = x >> 5, lx = x & 31;
    int k = (int)kx;
    u32 m = hbm[k] & (0xffffffffu >> (31 - lx));
    while (!m && k > 0 && kx - (u32)k <= SEG_LIMIT / 32) { --k; m = hbm[k]; }
    if (!m) { cls[p] = all; return; }
    const u32 a = (u32)k * 32 + (31 - (u32)__clz((int)m));                          // first rank of the segment
    k = (int)kx;
    m = lx == 31 ? 0u : hbm[k] & (0xffffffffu << (lx + 1));
    while (!m) { ++k; m = hbm[k]; }                                                 // hbm[NWIN] ends the search
    const u32 e = (u32)k * 32 + (u32)__ffs((int)m) - 1;                             // first rank of the next segment
    if (x - a > SEG_LIMIT || e - a > SEG_LIMIT + 1) { cls[p] = all; return; }       // more than SEG_LIMIT + 1 ranks: low-complexity repeats, kept as they are
    const u32 ka = a >> 5, ke = (e - 1) >> 5;
    u64 A = pa[e - 1], B = pb[e - 1];
    const u64 A0 = (a & 31) ? pa[a - 1] : 0ull, B0 = (a & 31) ? pb[a - 1] : 0ull;
    if (ka == ke) { A -= A0; B -= B0; }
    else {
        A += wa[ka] - A0; B += wb[ka] - B0;
        for (u32 w = ka + 1; w < ke; ++w) { A += wa[w]; B += wb[w]; }
    }
    const u64 n = B & 0x3ff, s1 = (B >> 10) & 0x3ffff, s2 = B >> 28;
    const bool has_start = ((A >> 32) & 0x3ff) != 0;
    cls[p] = segment_class(A & 0xffffffffull, !has_start && n * s2 == s1 * s1, has_start, ((A >> 42) & 0x3ff) != 0, ((A >> 52) & 0x3ff) != 0, prune, seg_is_interval);
}

// Sharding by suffix-order slice: slice bounds are moved forward to the next segment start (no reportable interval
// crosses a segment boundary, pfp.h fact 1), then every segment outside [lo, hi) loses its place in the stream; it keeps its refinement
// only when it holds a whole phrase (phrase ranks are global).  The first out-of-slice rank after the slice's last segment becomes the
// stand-in that closes it with the true boundary LCP, exactly like a pruned neighbour.
__global__ void __launch_bounds__(TPB) shard_snap(const u32* __restrict__ rk, u32 dsize, u32 raw_lo, u32 raw_hi, u32* __restrict__ bounds) {
    __shared__ u32 found;
    const u32 raw = blockIdx.x == 0 ? raw_lo : raw_hi;
    u32 result = dsize;
    if (raw == 0) result = 0;
    else if (raw < dsize) {
        for (u32 base = raw; base < dsize; base += TPB) {
            if (threadIdx.x == 0) found = 0xffffffffu;
            __syncthreads();
            u32 p = base + threadIdx.x;
            if (p < dsize && (rk[p] & RK_HB)) atomicMin(&found, p);
            __syncthreads();
            u32 f = found;
            __syncthreads();
            if (f != 0xffffffffu) { result = f; break; }
        }
    }
    if (threadIdx.x == 0) bounds[blockIdx.x] = result;
}
__global__ void __launch_bounds__(TPB) shard_mask(const u32* __restrict__ bounds, u32 dsize, u8* __restrict__ cls) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dsize || (p >= bounds[0] && p < bounds[1])) return;
    u8 c = cls[p];
    cls[p] = (c & CLS_HAS_START) ? (u8)(CLS_REFINE | CLS_HAS_START) : (u8)0;
}

__global__ void __launch_bounds__(TPB) owned_positions(const u32* __restrict__ rk, u32 dsize, u32* __restrict__ out) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < dsize) out[p] = (rk[p] & RK_OWN) ? p + 1 : 0u;
}

// Per rank that is longer than w: key symbols shared with the previous such rank; member flags of the stream (surviving ranks and the
// stand-in that follows a surviving segment); -M: mkind 3 = the rank closes a threshold-only segment (stream.cu, segment_thresholds).
__global__ void __launch_bounds__(TPB) member_flags(const u64* __restrict__ keys, const u32* __restrict__ rk, const u8* __restrict__ cls,
                                                    const u32* __restrict__ prev_owned, u32 dsize, u32 bits, u32 cpk, u32 t, u8* __restrict__ klcp,
                                                    u32* __restrict__ member, u8* __restrict__ mkind, u8* __restrict__ refine,
                                                    u32* __restrict__ n_closers) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    bool closes = false;
    if (p < dsize) {
        u32 w = rk[p];
        u8 c = cls[p], kind = 0; u32 shared = 0;
        if (w & RK_OWN) {
            u32 q = p ? prev_owned[p - 1] : 0u;          // 1 + previous rank that is longer than w
            if (q) shared = key_shared(keys[q - 1], keys[p], bits, cpk);
            u8 cq = q ? cls[q - 1] : (u8)0;
            if (c & CLS_STREAM) kind = 1;
            else if (cq & CLS_STREAM) kind = 2;          // first rank after a surviving segment: carries the true LCP that closes it
            closes = q && (cq & CLS_THRESH) && shared < t;
        }
        klcp[p] = (u8)shared; member[p] = kind ? 1u : 0u; mkind[p] = closes ? (u8)(kind | 4) : kind;
        refine[p] = (u8)(((c & CLS_REFINE) ? 1 : 0) | ((c & CLS_HAS_START) ? 2 : 0));
    }
    if (p == dsize) member[p] = 0;
    // how many segments get their threshold that way (statistics): one atomic per block
    __shared__ u32 s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    unsigned m = __ballot_sync(FULL, closes);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(&s_cnt, (u32)__popc(m));
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(n_closers, s_cnt);
}

// after the refinement (which permutes sa_d inside key groups, so the rank words' per-suffix fields no longer line up): ranks that hold
// a whole phrase, looked up only in the segments that have one
// -M, at most 64 documents: document mask / anchor occurrence of every distinct phrase (PhraseDocs)
__global__ void __launch_bounds__(TPB) phrase_docs_init(const u32* __restrict__ doff, u32 nd, PhraseDocs* __restrict__ rec) {
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nd) { PhraseDocs r; r.mask = 0; r.anchor = 0; r.flags = 0; r.first = doff[g]; rec[g] = r; }
}
__global__ void __launch_bounds__(TPB) phrase_docs_fill(const u64* __restrict__ pstart, const u32* __restrict__ pid, u32 np, PfpPrune L,
                                                        PhraseDocs* __restrict__ rec) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= np) return;
    u64 a = pstart[j], b = pstart[j + 1] - 1;              // the text suffixes this occurrence owns start in [a, b]
    u32 d = L.doc_coarse[a >> L.coarse_shift];
    while (L.doc_start[d + 1] <= a) ++d;
    PhraseDocs* r = rec + pid[j];
    u32 fl = L.doc_start[d + 1] <= b ? PD_CROSSES : 0u;
    u64 bit = 1ull << (d & 63);
    u64 old = atomicOr((unsigned long long*)&r->mask, (unsigned long long)bit);
    if (old & bit) fl |= PD_DUPLICATE;
    if (d == 0) { fl |= PD_HAS_ANCHOR; r->anchor = a; }     // a second occurrence in document 0 sets PD_DUPLICATE: the value is then unused
    if (fl) atomicOr(&r->flags, fl);
}

__global__ void __launch_bounds__(TPB) phrase_start_flags(const u8* __restrict__ refine, const u32* __restrict__ sa_d, const u32* __restrict__ dphrase,
                                                          const u32* __restrict__ doff, u32 dsize, u32* __restrict__ flag) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < dsize) {
        u32 f = 0;
        if (refine[p] & 2) { u32 i = sa_d[p]; f = doff[dphrase[i]] == i ? 1u : 0u; }
        flag[p] = f;
    }
    if (p == dsize) flag[p] = 0;
}
__global__ void __launch_bounds__(TPB) phrase_ranks_from_order(const u32* __restrict__ flag, const u32* __restrict__ before, const u32* __restrict__ sa_d,
                                                               const u32* __restrict__ dphrase, u32 dsize, u32* __restrict__ prank) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < dsize && flag[p]) prank[dphrase[sa_d[p]]] = before[p] + 1;
}

__global__ void __launch_bounds__(TPB) phrase_rank_keys(const u32* __restrict__ isa_d, const u32* __restrict__ doff, u32 nd, u64* __restrict__ keys,
                                                        u32* __restrict__ vals) {
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nd) { keys[g] = isa_d[doff[g]]; vals[g] = g; }
}
__global__ void __launch_bounds__(TPB) phrase_rank_scatter(const u32* __restrict__ order, u32 nd, u32* __restrict__ prank) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nd) prank[order[r]] = r + 1;
}
__global__ void __launch_bounds__(TPB) make_parse(const u32* __restrict__ pid, const u32* __restrict__ prank, u32 np, u32* __restrict__ parse) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < np) parse[j] = prank[pid[j]];
    if (j == np) parse[j] = 0;
}
__global__ void __launch_bounds__(TPB) parse_keys(const u32* __restrict__ parse, u32 m, u32 b, u64* __restrict__ keys, u32* __restrict__ vals) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    u64 nxt = (j + 1 < m) ? parse[j + 1] : 0;
    keys[j] = ((u64)parse[j] << b) | nxt; vals[j] = j;
}

// longest common prefix of T'[a..) and T'[b..) (symbols at or past n are 0), at most cap
__device__ __forceinline__ u64 text_lcp(const u8* __restrict__ t, u64 a, u64 b, u64 cap) {
    u64 l = 0;
    while (l + 8 <= cap) {
        u64 x = load8(t, a + l) ^ load8(t, b + l);
        if (x) return l + (u64)((__ffsll((long long)x) - 1) >> 3);
        l += 8;
    }
    while (l < cap && t[a + l] == t[b + l]) ++l;
    return l;
}

// consecutive phrases per thread (Kasai inheritance inside a chunk).  Measured at C2: chunks of 32 -> 33.0 ms, chunks of 256 -> 39.4 ms.
constexpr u32 KASAI_CHUNK = 32;
// LCP of the phrases of adjacent rank (dictionary strings; by_rank[r - 1] = distinct phrase with rank r): plcp[r] for r = 2 .. nd,
// plcp[0] = plcp[1] = 0.  With a range-minimum index over it the LCP of ANY two phrases is one L2-resident query, so parse_lcp never
// touches the text (the reference does the same through lcpD and an RMQ, include/dictionary.hpp:87-101).
__global__ void __launch_bounds__(TPB) phrase_by_rank(const u32* __restrict__ prank, u32 nd, u32* __restrict__ by_rank) {
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nd) by_rank[prank[g] - 1] = g;
}
__global__ void __launch_bounds__(TPB) phrase_lcp(const u8* __restrict__ dict, const u32* __restrict__ doff, const u32* __restrict__ by_rank, u32 nd,
                                                  u32* __restrict__ plcp) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;          // phrase rank, 1-based; slot 0 unused
    if (r > nd) return;
    if (r < 2) { plcp[r] = 0; return; }
    u32 ga = by_rank[r - 2], gb = by_rank[r - 1];
    u32 a = doff[ga], b = doff[gb];
    u32 cap = min(doff[ga + 1] - a, doff[gb + 1] - b) - 1;   // symbols of the shorter phrase (terminator excluded)
    u32 l = 0;
    while (l + 8 <= cap) {
        u64 x = load8(dict, (u64)a + l) ^ load8(dict, (u64)b + l);
        if (x) { l += (u32)((__ffsll((long long)x) - 1) >> 3); plcp[r] = l; return; }
        l += 8;
    }
    while (l < cap && dict[a + l] == dict[b + l]) ++l;
    plcp[r] = l;
}

__global__ void __launch_bounds__(128) parse_lcp(const u64* __restrict__ pstart, const u32* __restrict__ parse,
                                                 const u32* __restrict__ sa_p, const u32* __restrict__ isa_p, u32 m, const u32* __restrict__ plcp,
                                                 const u32* __restrict__ prmq, u32 prmq_blocks, u32* __restrict__ s_lcp, u32* __restrict__ overflow) {
    u64 first = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * KASAI_CHUNK;
    u32 l = 0;
    for (u32 k = 0; k < KASAI_CHUNK; ++k) {
        u64 j = first + k;
        if (j >= m) return;
        u32 r = isa_p[j];
        if (r == 0) { s_lcp[0] = 0; l = 0; continue; }
        u64 j2 = sa_p[r - 1];
        u32 ra, rb;
        while ((ra = parse[j + l]) == (rb = parse[j2 + l])) ++l;       // the terminator 0 at m-1 is unique, so this stops inside the array
        u64 common = pstart[j + l] - pstart[j];
        if (ra && rb) common += range_min(plcp, prmq, prmq_blocks, min(ra, rb) + 1, max(ra, rb));      // LCP of the two differing phrases
        if (common > 0xfffffffeull) { *overflow = 1; common = 0xfffffffeull; }
        s_lcp[r] = (u32)common;
        if (l) --l;
    }
}

// ---- range minimum over s_lcp
__global__ void __launch_bounds__(TPB) rmq_block_minima(const u32* __restrict__ v, u32 m, u32 nblocks, u32* __restrict__ out) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    u32 lo = b * 32, hi = min(lo + 32, m), mn = 0xffffffffu;
    for (u32 i = lo; i < hi; ++i) mn = min(mn, v[i]);
    out[b] = mn;
}
__global__ void __launch_bounds__(TPB) rmq_level(const u32* __restrict__ prev, u32 nblocks, u32 half, u32* __restrict__ out) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    u32 o = b + half;
    out[b] = o < nblocks ? min(prev[b], prev[o]) : prev[b];
}

__global__ void __launch_bounds__(TPB) occ_keys(const u32* __restrict__ parse, const u32* __restrict__ isa_p, u32 np, u64* __restrict__ keys,
                                                u32* __restrict__ vals) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < np) { keys[j] = ((u64)parse[j] << 32) | isa_p[j + 1]; vals[j] = j; }
}
__global__ void __launch_bounds__(TPB) occ_fill(const u64* __restrict__ keys, const u32* __restrict__ order, const u64* __restrict__ pstart, u32 np, u32 nd,
                                                const u32* __restrict__ s_lcp, const u32* __restrict__ rmq, u32 rmq_blocks, u32* __restrict__ occ_first,
                                                u32* __restrict__ next_rank, u64* __restrict__ next_start, u32* __restrict__ occ_lcp) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) { occ_first[0] = 0; occ_first[nd + 1] = np; }
    if (q >= np) return;
    u64 key = keys[q];
    u32 rank = (u32)(key >> 32), nr = (u32)key;
    bool first = (q == 0) || (u32)(keys[q - 1] >> 32) != rank;
    if (first) occ_first[rank] = q;
    next_rank[q] = nr;
    next_start[q] = pstart[order[q] + 1];
    occ_lcp[q] = first ? 0u : range_min(s_lcp, rmq, rmq_blocks, (u32)keys[q - 1] + 1, nr);
}

}  // namespace

void build_pfp_index(const u8* text, u64 n, PfpIndex& ix, cudaStream_t s) {
    PfpParse pp;
    pfp_parse(text, n, ix, pp, s);
    pfp_finish(text, n, ix, pp, s);
}

// stages 1-3: triggers, phrase identity, dictionary (cheap; tells how repetitive the collection is)
void pfp_triggers_begin(u64 worst_symbols, PfpTriggers& tr, cudaStream_t s) {
    tr.reset();
    u64 blocks = (worst_symbols + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (blocks == 0 || blocks > 0x7fffffffull) return;
    tr.masks.alloc(blocks * TPB, s); tr.bcount.alloc(blocks + 1, s);
    tr.capacity_blocks = blocks;
}
void pfp_triggers_extend(const u8* text, u64 prefix_symbols, PfpTriggers& tr, cudaStream_t s) {
    if (!tr.capacity_blocks) return;
    u32 full = (u32)std::min<u64>(prefix_symbols / SCAN_BLOCK, tr.capacity_blocks);      // blocks whose positions all lie inside the finished prefix
    if (full <= tr.blocks_done) return;
    MMB_LAUNCH(trigger_masks, full - tr.blocks_done, TPB, 0, s, text, prefix_symbols, tr.blocks_done, tr.masks.p, tr.bcount.p);
    tr.blocks_done = full; tr.text = text;
}

void pfp_parse(const u8* text, u64 n, PfpIndex& ix, PfpParse& pp, cudaStream_t s, PfpTriggers* ahead) {
    ix.n = n;
    if (n < PFP_W + 2) throw Error(70, "text too short for the prefix-free parse route");
    // ---- 1. triggers
    PhaseTimer* pt = new PhaseTimer("pfp 1 triggers", s);
    u64 nblk64 = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nblk64 > 0x7fffffffull) throw Error(71, "text too long");
    u32 nblk = (u32)nblk64;
    DevBuf<u32> bcount, bbase; DevBuf<u16> masks;
    u32 done = 0;
    if (ahead && ahead->text == text && ahead->blocks_done && ahead->blocks_done <= nblk && ahead->capacity_blocks >= nblk) {
        masks = std::move(ahead->masks); bcount = std::move(ahead->bcount); done = ahead->blocks_done;     // scanned while the upload was running
    } else { bcount.alloc(nblk + 1, s); masks.alloc((u64)nblk * TPB, s); }
    if (ahead) ahead->reset();
    if (getenv("MMB_DEBUG")) fprintf(stderr, "[pfp] trigger blocks scanned ahead of the parse: %u of %u\n", done, nblk);
    bbase.alloc(nblk + 1, s);
    MMB_CUDA(cudaMemsetAsync(bcount.p + nblk, 0, 4, s));
    if (done < nblk) MMB_LAUNCH(trigger_masks, nblk - done, TPB, 0, s, text, n, done, masks.p, bcount.p);
    prim::exclusive_sum(bcount.p, bbase.p, nblk + 1, s);
    u64 ntrig = to_host(bbase.p + nblk, 1, s)[0];
    if (ntrig + 2 >= 0xfffffff0ull) throw Error(72, "too many phrases");
    u32 np = (u32)ntrig + 1;
    ix.np = np;
    DevBuf<u64> ends; ends.alloc(np, s);
    MMB_LAUNCH(trigger_positions, nblk, TPB, 0, s, (const u16*)masks.p, (const u32*)bbase.p, ends.p);
    masks.release();
    u64 last_end = n + PFP_W - 1;
    MMB_CUDA(cudaMemcpyAsync(ends.p + np - 1, &last_end, 8, cudaMemcpyHostToDevice, s));
    ix.pstart.alloc(np + 1, s);
    MMB_LAUNCH(phrase_starts, ceil_div(np + 1, TPB), TPB, 0, s, (const u64*)ends.p, np, n, ix.pstart.p);
    MMB_CUDA(cudaStreamSynchronize(s));           // last_end lives on this stack frame
    ends.release(); bcount.release(); bbase.release();

    // ---- 2. phrase identity
    delete pt; pt = new PhaseTimer("pfp 2 phrase identity", s);
    DevBuf<u32>& pid = pp.pid; DevBuf<u32> rep; pid.alloc(np, s); rep.alloc(np, s);
    u32 nd = 0;
    for (u64 attempt = 0;; ++attempt) {
        if (attempt == 8) throw Error(73, "phrase hashing kept colliding");
        DevBuf<u64> ka, kb; DevBuf<u32> va, vb, flag, gnum, mismatch;
        ka.alloc(np, s); kb.alloc(np, s); va.alloc(np, s); vb.alloc(np, s); flag.alloc(np, s); gnum.alloc(np, s); mismatch.alloc(1, s); mismatch.zero();
        MMB_LAUNCH(phrase_hash, ceil_div(np, TPB), TPB, 0, s, text, (const u64*)ix.pstart.p, np, mix64(attempt + 0x1234567ull), ka.p, va.p);
        prim::sort_pairs(ka.p, kb.p, va.p, vb.p, np, 0, 64, s);
        MMB_LAUNCH(hash_boundaries, ceil_div(np, TPB), TPB, 0, s, (const u64*)kb.p, np, flag.p);
        prim::exclusive_sum(flag.p, gnum.p, np, s);
        MMB_LAUNCH(assign_groups, ceil_div(np, TPB), TPB, 0, s, (const u32*)flag.p, (const u32*)gnum.p, (const u32*)vb.p, np, pid.p, rep.p);
        MMB_LAUNCH(verify_groups, ceil_div((u64)np * PH_LANES, TPB), TPB, 0, s, text, (const u64*)ix.pstart.p, (const u32*)pid.p, (const u32*)rep.p, np, mismatch.p);
        u32 tail[3];
        MMB_CUDA(cudaMemcpyAsync(&tail[0], gnum.p + np - 1, 4, cudaMemcpyDeviceToHost, s));
        MMB_CUDA(cudaMemcpyAsync(&tail[1], flag.p + np - 1, 4, cudaMemcpyDeviceToHost, s));
        MMB_CUDA(cudaMemcpyAsync(&tail[2], mismatch.p, 4, cudaMemcpyDeviceToHost, s));
        MMB_CUDA(cudaStreamSynchronize(s));
        nd = tail[0] + tail[1];
        if (!tail[2]) break;
    }
    ix.nd = nd;

    // ---- 3. dictionary
    delete pt; pt = new PhaseTimer("pfp 3 dictionary", s);
    {
        DevBuf<u64> dlen, doff64; dlen.alloc(nd + 1, s); doff64.alloc(nd + 1, s);
        MMB_LAUNCH(dict_lengths, ceil_div(nd + 1, TPB), TPB, 0, s, (const u64*)ix.pstart.p, (const u32*)rep.p, nd, dlen.p);
        prim::exclusive_sum(dlen.p, doff64.p, nd + 1, s);
        u64 dsize = to_host(doff64.p + nd, 1, s)[0];
        if (dsize >= 0xffffff00ull) throw Error(74, "dictionary too large for 32-bit indices");
        ix.dsize = (u32)dsize;
        ix.doff.alloc(nd + 1, s);
        MMB_LAUNCH(narrow_offsets, ceil_div(nd + 1, TPB), TPB, 0, s, (const u64*)doff64.p, nd + 1, ix.doff.p);
    }
    u32 dsize = ix.dsize;
    ix.dict.alloc((u64)dsize + 64, s); ix.limit.alloc(dsize, s); ix.dphrase.alloc(dsize, s);
    MMB_CUDA(cudaMemsetAsync(ix.dict.p + dsize, 0, 64, s));
    MMB_LAUNCH(dict_fill, ceil_div((u64)nd * 32, TPB), TPB, 0, s, text, n, (const u64*)ix.pstart.p, (const u32*)rep.p, (const u32*)ix.doff.p, nd, ix.dict.p,
               ix.limit.p, ix.dphrase.p);
    rep.release();
    delete pt;
}

// stages 4-8
void pfp_finish(const u8* text, u64 n, PfpIndex& ix, PfpParse& pp, cudaStream_t s, const PfpPrune& prune) {
    u32 np = ix.np, nd = ix.nd, dsize = ix.dsize;
    DevBuf<u32>& pid = pp.pid;
    // ---- 4. suffix order of the dictionary (ties allowed)
    PhaseTimer* pt = new PhaseTimer("pfp 4 dictionary SA", s);
    bool pruned_done = false;
    ix.prank.alloc(nd, s);
    {
        DevBuf<u32> hist; hist.alloc(256, s); hist.zero();
        MMB_LAUNCH(dict_histogram, 148 * 8, TPB, 0, s, (const u8*)ix.dict.p, dsize, hist.p);
        std::vector<u32> h = to_host(hist.p, 256, s);
        std::vector<u8> lut(256, 0); u32 sigma = 0;
        for (int c = 1; c < 256; ++c) if (h[c]) lut[c] = (u8)(++sigma);
        u32 bits = 1; while ((1u << bits) <= sigma) ++bits;
        u32 cpk = 64 / bits;
        // 18 key symbols (54 bits = 7 radix passes of the initial sort instead of 8) when segments need not be exact intervals: the
        // segments get coarser (t = min(-l, key symbols)), at C2 the stream grows by 7.6 % and the sort saves 5.3 ms: 330.8 -> 328.5 ms;
        // 16 symbols lose (the stream doubles: 352.8 ms).  -M keeps all 21: its threshold-only segments need t = -l.  MMB_KEY_SYMBOLS
        // overrides (never below w + 2: the pruned route needs t > w).
        { const u32 narrow = std::max(18u, PFP_W + 1); if (!prune.merge && cpk > narrow) cpk = narrow; }      // the key must hold symbol w (longer-than-w flag)
        if (const char* ks = getenv("MMB_KEY_SYMBOLS")) { u32 v = (u32)atoi(ks); if (v >= PFP_W + 1 && v <= 64 / bits) cpk = v; }
        DevBuf<u8> d_lut; d_lut.alloc(256, s);
        MMB_CUDA(cudaMemcpyAsync(d_lut.p, lut.data(), 256, cudaMemcpyHostToDevice, s));
        MMB_CUDA(cudaStreamSynchronize(s));       // lut lives on this stack frame
        ix.sa_d.alloc(dsize, s);
        DevBuf<u64> ka, kb; DevBuf<u32> va; ka.alloc(dsize, s); kb.alloc(dsize, s); va.alloc(dsize, s);
        MMB_LAUNCH(dict_keys, ceil_div(dsize, TPB), TPB, 0, s, (const u8*)ix.dict.p, (const u32*)ix.limit.p, dsize, (const u8*)d_lut.p, bits, cpk, ka.p, va.p);
        // Pruned route: needs the longer-than-w flag in the key (cpk > w) and a segment prefix of at least w + 1 symbols.
        const u32 t = std::min(prune.min_len, cpk);
        // t may be at or below w: every owned suffix has more than w >= t symbols then (none is "short"), and a suffix of at most w symbols
        // that shares its first t symbols with owned ones sits inside their segment without owning anything (validated by the fuzzer)
        const bool try_pruned = cpk > PFP_W && t > 10 && !getenv("MMB_NO_PRUNE");
        if (try_pruned) {
            prim::sort_pairs(ka.p, kb.p, va.p, ix.sa_d.p, dsize, 0, (int)(bits * cpk), s);
            ka.release(); va.release();
            DevBuf<u32> count; DevBuf<u64> pinfo; count.alloc(nd, s); count.zero(); pinfo.alloc(dsize, s);
            MMB_LAUNCH(phrase_counts, ceil_div(np, TPB), TPB, 0, s, (const u32*)pid.p, np, count.p);
            if (prune.enabled && prune.merge && prune.ndocs <= 64 && prune.doc_coarse) {
                ix.phrase_docs.alloc(nd, s);
                MMB_LAUNCH(phrase_docs_init, ceil_div(nd, TPB), TPB, 0, s, (const u32*)ix.doff.p, nd, ix.phrase_docs.p);
                MMB_LAUNCH(phrase_docs_fill, ceil_div(np, TPB), TPB, 0, s, (const u64*)ix.pstart.p, (const u32*)pid.p, np, prune, ix.phrase_docs.p);
            }
            MMB_LAUNCH(dict_info, ceil_div(dsize, TPB), TPB, 0, s, (const u32*)ix.dphrase.p, (const u32*)ix.doff.p, (const u8*)ix.dict.p, (const u32*)count.p, dsize, pinfo.p);
            ix.rk.alloc(dsize, s);
            if (ix.phrase_docs.p) ix.rank_phrase.alloc(dsize, s);
            MMB_LAUNCH(rank_words, ceil_div(dsize, TPB), TPB, 0, s, (const u64*)kb.p, (const u32*)ix.sa_d.p, (const u64*)pinfo.p, dsize, bits, cpk, t, ix.rk.p,
                       ix.rank_phrase.p);
            MMB_LAUNCH(rank_words_fix_last_phrase, PFP_W, TPB, 0, s, (const u32*)pid.p, np, (const u32*)ix.doff.p, (const u32*)ix.limit.p, (const u8*)ix.dict.p,
                       (const u8*)d_lut.p, bits, cpk, (const u64*)kb.p, (const u32*)ix.sa_d.p, dsize, ix.rk.p);
            pinfo.release(); count.release();
            DevBuf<u8> cls, refine; cls.alloc(dsize, s); refine.alloc(dsize, s);
            MMB_LAUNCH(classify_segments, ceil_div(dsize, TPB), TPB, 0, s, (const u32*)ix.rk.p, dsize, prune, (u32)(t == prune.min_len ? 1 : 0), cls.p);
            if (prune.shard_count > 1) {
                DevBuf<u32> bounds; bounds.alloc(2, s);
                u32 raw_lo = (u32)((u64)dsize * prune.shard_index / prune.shard_count), raw_hi = (u32)((u64)dsize * (prune.shard_index + 1) / prune.shard_count);
                if (prune.shard_index + 1 == prune.shard_count) raw_hi = dsize;
                MMB_LAUNCH(shard_snap, 2, TPB, 0, s, (const u32*)ix.rk.p, dsize, raw_lo, raw_hi, bounds.p);
                MMB_LAUNCH(shard_mask, ceil_div(dsize, TPB), TPB, 0, s, (const u32*)bounds.p, dsize, cls.p);
            }
            ix.prev_owned.alloc(dsize, s);
            {
                DevBuf<u32> own_pos; own_pos.alloc(dsize, s);
                MMB_LAUNCH(owned_positions, ceil_div(dsize, TPB), TPB, 0, s, (const u32*)ix.rk.p, dsize, own_pos.p);
                prim::inclusive_max(own_pos.p, ix.prev_owned.p, dsize, s);
            }
            ix.owned.alloc((u64)dsize + 1, s); ix.klcp.alloc(dsize, s); ix.mkind.alloc(dsize, s); ix.key_symbols = cpk; ix.seg_symbols = t;
            DevBuf<u32> n_closers; n_closers.alloc(1, s); n_closers.zero();
            MMB_LAUNCH(member_flags, ceil_div((u64)dsize + 1, TPB), TPB, 0, s, (const u64*)kb.p, (const u32*)ix.rk.p, (const u8*)cls.p, (const u32*)ix.prev_owned.p, dsize,
                       bits, cpk, t, ix.klcp.p, ix.owned.p, ix.mkind.p, refine.p, n_closers.p);
            cls.release();
            ix.ghead.alloc(dsize, s);
            // rounds of 21 symbols: phrase lengths are geometric (mean ~100), the longest of 1e8 phrases has ~2000 symbols = ~100 rounds, and
            // the late rounds only touch the few groups of such phrases; beyond ~5 000 shared symbols (a long single-symbol run inside
            // one phrase: one suffix leaves its group per symbol) give up and sort by doubling, which needs log rounds there
            const u32 max_rounds = getenv("MMB_CHUNK_ROUNDS") ? (u32)atoi(getenv("MMB_CHUNK_ROUNDS")) : 256u;
            pruned_done = refine_dictionary_chunks(kb.p, ix.sa_d.p, dsize, cpk, bits, ix.dict.p, ix.limit.p, d_lut.p, refine.p, ix.ghead.p, ix.stats_d, s, max_rounds);
            if (pruned_done) {
                ix.pruned = true; ix.full_stream = !prune.enabled;
                ix.n_closers = to_host(n_closers.p, 1, s)[0];
                // phrase ranks: whole phrases in suffix order
                DevBuf<u32> flag, before; flag.alloc((u64)dsize + 1, s); before.alloc((u64)dsize + 1, s);
                MMB_LAUNCH(phrase_start_flags, ceil_div((u64)dsize + 1, TPB), TPB, 0, s, (const u8*)refine.p, (const u32*)ix.sa_d.p, (const u32*)ix.dphrase.p,
                           (const u32*)ix.doff.p, dsize, flag.p);
                prim::exclusive_sum(flag.p, before.p, (u64)dsize + 1, s);
                MMB_LAUNCH(phrase_ranks_from_order, ceil_div(dsize, TPB), TPB, 0, s, (const u32*)flag.p, (const u32*)before.p, (const u32*)ix.sa_d.p,
                           (const u32*)ix.dphrase.p, dsize, ix.prank.p);
            } else {
                // a phrase with a very long low-complexity stretch: start over with the doubling engine, nothing pruned
                ix.rk.release(); ix.prev_owned.release(); ix.owned.release(); ix.klcp.release(); ix.mkind.release(); ix.ghead.release(); ix.phrase_docs.release(); ix.rank_phrase.release();
                ix.stats_d = SaStats();
                ka.alloc(dsize, s); va.alloc(dsize, s);
                MMB_LAUNCH(dict_keys, ceil_div(dsize, TPB), TPB, 0, s, (const u8*)ix.dict.p, (const u32*)ix.limit.p, dsize, (const u8*)d_lut.p, bits, cpk, ka.p, va.p);
            }
        }
        if (!pruned_done) {
            ix.isa_d.alloc(dsize, s);
            DevBuf<u32> isa2; isa2.alloc(2 * (u64)dsize, s);
            MMB_LAUNCH(phrase_offsets, ceil_div(dsize, TPB), TPB, 0, s, (const u32*)ix.limit.p, (const u32*)ix.dphrase.p, (const u32*)ix.doff.p, dsize, isa2.p);
            GroupStore groups;
            order_by_initial_keys(ka.p, kb.p, va.p, dsize, (int)(bits * cpk), cpk, ix.limit.p, ix.sa_d.p, isa2.p, groups, s, 2);
            if (cpk > PFP_W) {
                ix.owned.alloc((u64)dsize + 1, s); ix.klcp.alloc(dsize, s); ix.key_symbols = cpk;
                MMB_LAUNCH(owned_from_keys, ceil_div((u64)dsize + 1, TPB), TPB, 0, s, (const u64*)kb.p, dsize, bits, cpk, ix.owned.p, ix.klcp.p);
            }
            ka.release(); kb.release(); va.release();
            refine_by_doubling(ix.sa_d.p, isa2.p, dsize, cpk, ix.limit.p, groups, ix.stats_d, s, 2);
            MMB_LAUNCH(unpair_ranks, ceil_div(dsize, TPB), TPB, 0, s, (const u32*)isa2.p, dsize, ix.isa_d.p);
            isa2.release();
            if (ix.owned.p) MMB_LAUNCH(owned_fix_last_phrase, 1, 256, 0, s, (const u32*)pid.p, np, (const u32*)ix.doff.p, (const u32*)ix.limit.p, (const u32*)ix.isa_d.p, (const u32*)ix.sa_d.p, dsize, ix.owned.p);
        }
        MMB_CUDA(cudaStreamSynchronize(s));
    }

    // ---- 5. phrase ranks and the parse
    delete pt; pt = new PhaseTimer("pfp 5 phrase ranks", s);
    if (!pruned_done) {
        DevBuf<u64> ka, kb; DevBuf<u32> va, vb; ka.alloc(nd, s); kb.alloc(nd, s); va.alloc(nd, s); vb.alloc(nd, s);
        MMB_LAUNCH(phrase_rank_keys, ceil_div(nd, TPB), TPB, 0, s, (const u32*)ix.isa_d.p, (const u32*)ix.doff.p, nd, ka.p, va.p);
        prim::sort_pairs(ka.p, kb.p, va.p, vb.p, nd, 0, 32, s);
        MMB_LAUNCH(phrase_rank_scatter, ceil_div(nd, TPB), TPB, 0, s, (const u32*)vb.p, nd, ix.prank.p);
    }
    u32 m = np + 1;
    ix.parse.alloc((u64)m + 1, s);
    MMB_LAUNCH(make_parse, ceil_div(m, TPB), TPB, 0, s, (const u32*)pid.p, (const u32*)ix.prank.p, np, ix.parse.p);
    MMB_CUDA(cudaMemsetAsync(ix.parse.p + m, 0, 4, s));
    pid.release();

    // ---- 6. suffix order of the parse
    delete pt; pt = new PhaseTimer("pfp 6 parse SA", s);
    DevBuf<u32> sa_p; sa_p.alloc(m, s); ix.isa_p.alloc(m, s);
    {
        u32 b = 1; while ((1ull << b) <= nd) ++b;
        GroupStore groups;
        {
            DevBuf<u64> ka, kb; DevBuf<u32> va; ka.alloc(m, s); kb.alloc(m, s); va.alloc(m, s);
            MMB_LAUNCH(parse_keys, ceil_div(m, TPB), TPB, 0, s, (const u32*)ix.parse.p, m, b, ka.p, va.p);
            order_by_initial_keys(ka.p, kb.p, va.p, m, (int)(2 * b), 2, nullptr, sa_p.p, ix.isa_p.p, groups, s);
        }
        refine_by_doubling(sa_p.p, ix.isa_p.p, m, 2, nullptr, groups, ix.stats_p, s);
    }

    // ---- 7. text-level LCP between rank-adjacent parse suffixes
    delete pt; pt = new PhaseTimer("pfp 7 s_lcp", s);
    ix.s_lcp.alloc(m, s);
    {
        // LCP of phrases of adjacent rank + range-minimum index (blocks of 32 + sparse table over the blocks: a few MB, L2 resident)
        DevBuf<u32> by_rank, plcp, prmq; by_rank.alloc(nd, s); plcp.alloc((u64)nd + 1, s);
        MMB_LAUNCH(phrase_by_rank, ceil_div(nd, TPB), TPB, 0, s, (const u32*)ix.prank.p, nd, by_rank.p);
        MMB_LAUNCH(phrase_lcp, ceil_div((u64)nd + 1, TPB), TPB, 0, s, (const u8*)ix.dict.p, (const u32*)ix.doff.p, (const u32*)by_rank.p, nd, plcp.p);
        u32 pblocks = (nd + 1 + 31) / 32, plevels = 1;
        while ((1u << plevels) <= pblocks) ++plevels;
        prmq.alloc((u64)plevels * pblocks, s);
        MMB_LAUNCH(rmq_block_minima, ceil_div(pblocks, TPB), TPB, 0, s, (const u32*)plcp.p, nd + 1, pblocks, prmq.p);
        for (u32 k = 1; k < plevels; ++k)
            MMB_LAUNCH(rmq_level, ceil_div(pblocks, TPB), TPB, 0, s, (const u32*)(prmq.p + (u64)(k - 1) * pblocks), pblocks, 1u << (k - 1),
                       prmq.p + (u64)k * pblocks);
        DevBuf<u32> overflow; overflow.alloc(1, s); overflow.zero();
        u64 threads = ((u64)m + KASAI_CHUNK - 1) / KASAI_CHUNK;
        MMB_LAUNCH(parse_lcp, ceil_div(threads, 128), 128, 0, s, (const u64*)ix.pstart.p, (const u32*)ix.parse.p, (const u32*)sa_p.p,
                   (const u32*)ix.isa_p.p, m, (const u32*)plcp.p, (const u32*)prmq.p, pblocks, ix.s_lcp.p, overflow.p);
        if (to_host(overflow.p, 1, s)[0]) throw Error(75, "a common prefix longer than 2^32 symbols: not supported");
    }
    sa_p.release();

    // ---- 8. range-minimum index and occurrence lists
    delete pt; pt = new PhaseTimer("pfp 8 rmq + occurrences", s);
    ix.rmq_blocks = (m + 31) / 32;
    ix.rmq_levels = 1; while ((1u << ix.rmq_levels) <= ix.rmq_blocks) ++ix.rmq_levels;
    ix.rmq.alloc((u64)ix.rmq_levels * ix.rmq_blocks, s);
    MMB_LAUNCH(rmq_block_minima, ceil_div(ix.rmq_blocks, TPB), TPB, 0, s, (const u32*)ix.s_lcp.p, m, ix.rmq_blocks, ix.rmq.p);
    for (u32 k = 1; k < ix.rmq_levels; ++k)
        MMB_LAUNCH(rmq_level, ceil_div(ix.rmq_blocks, TPB), TPB, 0, s, (const u32*)(ix.rmq.p + (u64)(k - 1) * ix.rmq_blocks), ix.rmq_blocks, 1u << (k - 1),
                   ix.rmq.p + (u64)k * ix.rmq_blocks);
    ix.occ_first.alloc(nd + 2, s); ix.occ_next_rank.alloc(np, s); ix.occ_next_start.alloc(np, s); ix.occ_lcp.alloc(np, s);
    {
        DevBuf<u64> ka, kb; DevBuf<u32> va, vb; ka.alloc(np, s); kb.alloc(np, s); va.alloc(np, s); vb.alloc(np, s);
        u32 b = 1; while ((1ull << b) <= nd) ++b;
        MMB_LAUNCH(occ_keys, ceil_div(np, TPB), TPB, 0, s, (const u32*)ix.parse.p, (const u32*)ix.isa_p.p, np, ka.p, va.p);
        prim::sort_pairs(ka.p, kb.p, va.p, vb.p, np, 0, (int)(32 + b), s);
        MMB_LAUNCH(occ_fill, ceil_div(np, TPB), TPB, 0, s, (const u64*)kb.p, (const u32*)vb.p, (const u64*)ix.pstart.p, np, nd, (const u32*)ix.s_lcp.p,
                   (const u32*)ix.rmq.p, ix.rmq_blocks, ix.occ_first.p, ix.occ_next_rank.p, ix.occ_next_start.p, ix.occ_lcp.p);
    }
    delete pt;
}

}  // namespace mmb
