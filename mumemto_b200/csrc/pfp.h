// mumemto_b200/csrc/pfp.h -- prefix-free parsing route (pfp.cu, stream.cu): suffix stream of T from the suffix arrays of a
// dictionary and a parse that are ~|T|/25 and |T|/100 in size on pangenome-like input.
//
// Reference: pfparser (include/newscan.hpp:310-423), dictionary / parse / pf_parsing (include/dictionary.hpp:103-157,
// include/parse.hpp:77-146, include/pfp.hpp:171-244) and the enumeration loop pfp_lcp::process (include/pfp_lcp_mum.hpp:115-231)
// -- rows B, C, D of SURVEY.md section 8.  The outputs of that path do not depend on the hash, window or modulus, so
// the device parse uses its own window hash; what is kept is the construction:
//   T' = T . 0^w.  Trigger positions e (w <= e < n) are where hash(T[e-w+1..e]) % p == 0, plus the forced last one n+w-1.
//   Phrase j = T'[s_j .. e_j], s_0 = 0, s_{j+1} = e_j - w + 1 (consecutive phrases overlap by w).  Text position i belongs to the
//   phrase with s_j <= i <= e_j - w; its "dictionary suffix" is alpha(i) = T'[i .. e_j], longer than w.  The set of all alpha is
//   prefix-free, so two text suffixes compare by (alpha, then the parse suffix that follows), and
//   lcp = lcp(alpha, alpha') if they differ, else |alpha| - w + lcp of the following phrase-aligned suffixes.
#pragma once
#include "common.cuh"
#include "sa.h"

namespace mmb {

#ifndef MMB_PFP_W
#define MMB_PFP_W 15
#endif
constexpr u32 PFP_W = MMB_PFP_W;    // window.  The reference defaults to -w 10; the outputs do not depend on it, the cost does.
                                    // A dictionary suffix is valid when it is longer than w.  The text positions whose valid suffix has L < -l
                                    // symbols cannot be pruned from the key; grouped by content they are ALL text positions sharing an L-mer that
                                    // ends in a trigger window: about 2 N |genome| / 4^L suffixes per group (64 x 100 Mbp: 3050 for L = 11, 763
                                    // for 12, 190 for 13, 48 for 14, 12 for 15, 3 for 16).  With w = 10 the groups of L = 11..13 survived every
                                    // size rule, made up more than half of the pruned stream (684 M ranks at C2) and caused all of its large
                                    // merges.  From w = 15 on the smallest valid suffix has 16 symbols and such groups fall under the size
                                    // rule.  Measured at C2 / C4 / one C5 partition (ms): w = 10: 330.7 / 392.4 / 223.5; w = 19: 305.4 / 341.8 /
                                    // 203.5; w = 17: 301.4 / 339.0 / 204.7; w = 15: 297.2 / 336.4 / 205.9 (a smaller w = a smaller dictionary:
                                    // phrases overlap by w; and w <= 17 lets the initial sort use 18-symbol keys = 7 radix passes).
constexpr u32 PFP_P = 100;          // modulus (reference default -m 100)

// -M on the pruned route, collections of at most 64 documents: which documents the occurrences of a phrase lie in.
struct PhraseDocs {
    u64 mask;               // bit d: an occurrence of the phrase starts in document d
    u64 anchor;             // text position where the phrase's occurrence in document 0 starts (valid with PD_HAS_ANCHOR)
    u32 flags;              // PD_*
    u32 first;              // dictionary position where the phrase starts (doff)
};
enum : u32 { PD_DUPLICATE = 1, PD_CROSSES = 2, PD_HAS_ANCHOR = 4 };     // a document twice; an occurrence spans a document boundary

struct PfpIndex {
    u64 n = 0;                      // |T|
    // parse
    u32 np = 0;                     // phrases; the parse has np + 1 symbols (terminator 0)
    DevBuf<u64> pstart;             // [np + 1] text position of each phrase start; pstart[np] = n
    DevBuf<u32> parse;              // [np + 1] lexicographic rank of each phrase (1-based), parse[np] = 0
    DevBuf<u32> isa_p;              // [np + 1] rank of each parse suffix
    DevBuf<u32> s_lcp;              // [np + 1] lcp in TEXT symbols between parse suffixes of adjacent rank (s_lcp[0] = 0)
    // range-minimum over s_lcp: minima of 32-entry blocks and a sparse table over the blocks
    u32 rmq_blocks = 0, rmq_levels = 0;
    DevBuf<u32> rmq;                // [rmq_levels][rmq_blocks]
    // dictionary D' = distinct phrases, each followed by a terminator 0
    u32 nd = 0;                     // distinct phrases
    u32 dsize = 0;                  // |D'|
    DevBuf<u8> dict;                // [dsize + 64]
    DevBuf<u32> limit;              // [dsize] symbols from here up to and including the terminator
    DevBuf<u32> sa_d, isa_d;        // suffix order of D' (equal strings share one head rank in isa_d; isa_d is empty on the pruned route)
    DevBuf<u32> owned;              // [dsize + 1] by rank in sa_d: 1 = the suffix is a member of the stream: longer than w and (pruned route)
                                    //   not pruned, or the stand-in that follows a surviving segment (may be empty: stream.cu then derives it)
    DevBuf<u8> klcp;                // [dsize] by rank in sa_d: symbols shared inside the initial sort key with the previous rank (full route) or
    u32 key_symbols = 0;            //   with the previous rank that is longer than w (pruned route), capped at key_symbols
    // ---- pruned route only (see PfpPrune)
    bool pruned = false;            // the dictionary order came from the chunk engine (rank words, ghead ... are filled)
    bool full_stream = true;        // nothing was left out: the stream has all |T| suffixes
    DevBuf<u32> ghead;              // [dsize] by rank: first rank of the run of equal strings the rank belongs to
    DevBuf<u8> mkind;               // [dsize] by rank: 1 = ordinary member, 2 = stand-in rank after a surviving segment; bit 2 (4): the rank
                                    //   closes a threshold-only segment (-M)
    DevBuf<u32> rk;                 // [dsize] by rank: segment start / longer-than-w / occurrences / preceding symbol (pfp.cu: rank words)
    DevBuf<u32> prev_owned;         // [dsize] by rank: 1 + the last rank <= p that is longer than w (0 = none)
    u32 n_closers = 0;              // -M: segments whose only possible interval is the whole segment (threshold only)
    DevBuf<PhraseDocs> phrase_docs; // -M, at most 64 documents: per distinct phrase (dictionary order)
    DevBuf<u32> rank_phrase;        // -M, with phrase_docs: distinct phrase of the suffix at every rank of the INITIAL key order (stays valid for
                                    //   ranks whose key group is never refined: the threshold-only segments)
    u32 seg_symbols = 0;            // symbols that define a segment (min(-l, key_symbols))
    u64 stream_ranks = 0;           // text suffixes that are enumerated (members' occurrences + stand-ins)
    DevBuf<u32> dphrase;            // [dsize] distinct-phrase id (in D' order) of every dictionary position
    DevBuf<u32> doff;               // [nd + 1] start of each distinct phrase in D'
    DevBuf<u32> prank;              // [nd] lexicographic rank (1-based) of each distinct phrase
    // occurrences of every phrase rank, ordered by the rank of the parse suffix that follows them
    DevBuf<u32> occ_first;          // [nd + 2] indexed by phrase rank
    DevBuf<u32> occ_next_rank;      // [np] isa_p of the following parse suffix
    DevBuf<u64> occ_next_start;     // [np] text position where the following phrase starts
    DevBuf<u32> occ_lcp;            // [np] text lcp between the following suffixes of this and the previous occurrence of the phrase
    SaStats stats_d, stats_p;
};

struct PfpParse { DevBuf<u32> pid; };      // provisional phrase id of every phrase occurrence, handed from pfp_parse to pfp_finish

// Segment pruning.  Sorted by their first t = min(-l, key symbols) symbols the dictionary suffixes fall into
// SEGMENTS; text suffixes of different segments share fewer than -l symbols, so no reportable interval crosses a segment boundary.  A
// segment can be left out of the suffix stream when none of its intervals can be reported:
//   * fewer text suffixes than the smallest interval size that qualifies (max(2, -k)), or
//   * every suffix is preceded by the same symbol inside its phrase: no interval is left-maximal (the reference's check_bwt_range,
//     include/mem_finder.hpp:189-192).  With -M the thresholds of such intervals are still needed (recorded before the BWT test,
//     :326-336): a segment of exactly N suffixes is one interval, whose threshold needs no suffix order at all (stream.cu,
//     segment_thresholds); any other same-symbol segment stays in the stream.
// Only surviving segments (and the key groups that hold a whole phrase: phrase ranks) are sorted beyond the initial key.
struct PfpPrune {
    u32 enabled = 0;
    u32 min_len = 20;       // -l
    u32 size_lo = 2;        // smallest interval size that can qualify
    u32 merge = 0;          // -M
    u32 ndocs = 0;
    u32 shard_index = 0, shard_count = 1;   // only the segments of this slice of the dictionary suffix order are streamed
    // -M: position -> document lookup (device pointers; coarse table + document starts, as DetectParams::doc_coarse)
    const u16* doc_coarse = nullptr; u32 coarse_shift = 0; const u64* doc_start = nullptr;
};


// Trigger masks computed AHEAD of the parse, while the collection is still being uploaded (engine.cu: load_common): a window only looks
// backwards, so every 4096-position block that lies inside the finished prefix of the text can be scanned as soon as that prefix exists.
// pfp_parse takes over what is there and scans the rest.
struct PfpTriggers {
    DevBuf<u16> masks; DevBuf<u32> bcount;
    u64 capacity_blocks = 0; u32 blocks_done = 0; const u8* text = nullptr;
    void reset() { masks.release(); bcount.release(); capacity_blocks = 0; blocks_done = 0; text = nullptr; }
};
void pfp_triggers_begin(u64 worst_symbols, PfpTriggers& tr, cudaStream_t s);
void pfp_triggers_extend(const u8* text, u64 prefix_symbols, PfpTriggers& tr, cudaStream_t s);

void build_pfp_index(const u8* text, u64 n, PfpIndex& out, cudaStream_t s);
// The same in two steps: parse + dictionary (stages 1-3, cheap) and everything else; `auto` routing looks at ix.dsize in between.
void pfp_parse(const u8* text, u64 n, PfpIndex& ix, PfpParse& pp, cudaStream_t s, PfpTriggers* ahead = nullptr);
void pfp_finish(const u8* text, u64 n, PfpIndex& ix, PfpParse& pp, cudaStream_t s, const PfpPrune& prune = PfpPrune());

// One window of the suffix stream in rank order, with `halo` ranks of the previous window in front of it.
struct StreamTile {
    u64 rank0 = 0;                  // global rank of view index 0
    u64 count = 0;                  // valid entries in the view (halo included)
    u64 first = 0;                  // view index of the first rank that is new in this tile
    const u64* sa = nullptr;        // [count]
    const u32* lcp = nullptr;       // [count + 64], zero padded
    const u8* bwt = nullptr;        // [count] symbol preceding each suffix (0 for text position 0)
    const u8* doc = nullptr;        // [count] document of each suffix (255 for the sentinel), only when stream_suffixes was given a DocLookup
    bool last = false;
};

// Optional: have every tile carry one document-id byte per rank (collections of at most 254 documents).  Experimental (MMB_DOC_STREAM=1):
// with -M every interval of the right size needs a document check, which then reads count bytes instead of gathering count SA values.
struct DocLookup { const u16* coarse; u32 shift; const u64* doc_start; u32 ndocs; };

// Calls consume(tile) for consecutive tiles of about tile_ranks ranks, each carrying `halo` ranks of overlap.
struct TileConsumer { virtual void consume(const StreamTile& t) = 0; virtual ~TileConsumer() {} };
void stream_suffixes(const u8* text, const PfpIndex& ix, u64 tile_ranks, u32 halo, TileConsumer& consumer, cudaStream_t s,
                     const DocLookup* docs = nullptr);
// -M on the pruned route: thresholds of the segments whose closing rank is flagged in ix.mkind (every one is a single interval of ndocs suffixes).
void segment_thresholds(const PfpIndex& ix, const DocLookup& docs, u16* thresh, cudaStream_t s);

}  // namespace mmb
