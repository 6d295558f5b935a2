// mumemto_b200/csrc/engine.cu -- host orchestration of the device pipeline and the C ABI declared in include/mumemto_b200.h.
//
// Pipeline of mmb_find (one CUDA stream, no host work between stages except scalar read-backs that size the next allocation):
//   text (text.cu)  ->  suffix array (sa.cu)  ->  LCP + BWT (lcp.cu)  ->  interval detection (detect.cu)  ->  records + text (emit.cu)
// which stands in for build_main's stages in the reference (src/pfp_mum.cpp:57-145).  Parameter normalisation follows
// BuildOptions::validate / set_parameters (include/pfp_mum.hpp:80-198) for the CLI and mumemto::mumemto_mum / mumemto_mem
// (mumemto_library/mumemto_api.cpp:332-411) for library callers.
#include "../../include/mumemto_b200.h"
#include "detect.h"
#include "merge.h"
#include "pfp.h"
#include "sa.h"
#include "text.h"

#include <atomic>
#include <cctype>
#include <map>
#include <mutex>
#include <unordered_map>

namespace mmb {
std::atomic<unsigned long long> g_kernel_launches{0};
double g_host_alloc_ms = 0, g_host_free_ms = 0, g_host_sync_ms = 0;
unsigned long long g_host_allocs = 0, g_host_syncs = 0;
}

// ---------------------------------------------------------------------------------------------- cached device allocator
// Free lists are keyed by (device, stream): a block released by one stage may be handed out again at once only to work that is
// ordered on the SAME stream.  Two engines in one process (concurrent mumemto_mum / mumemto_mem calls, engines on different GPUs)
// therefore never see each other's blocks.
namespace mmb {
namespace poolimpl {
typedef std::pair<int, cudaStream_t> Key;
struct BlockCache {
    std::mutex mu;
    std::map<Key, std::multimap<size_t, void*>> free_blocks;      // (device, stream) -> size -> block
    std::unordered_map<void*, std::pair<size_t, Key>> live;      // blocks handed out
} g_cache;
size_t round_block(size_t b) { return b <= (1u << 20) ? ((b + 511) & ~(size_t)511) : ((b + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1)); }
int current_device() { int d = 0; cudaGetDevice(&d); return d; }
// give the cached blocks of one key (or of every key on `device` when stream_only is false) back to the driver
void trim(int device, cudaStream_t s, bool stream_only) {
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        for (auto& kv : g_cache.free_blocks) {
            if (kv.first.first != device || (stream_only && kv.first.second != s)) continue;
            for (auto& b : kv.second) drop.push_back(b.second);
            kv.second.clear();
        }
    }
    if (drop.empty()) return;
    if (stream_only) cudaStreamSynchronize(s); else cudaDeviceSynchronize();
    for (void* p : drop) cudaFree(p);
}
}  // namespace poolimpl
using poolimpl::g_cache;
using poolimpl::round_block;

void pool_trim() { poolimpl::trim(poolimpl::current_device(), nullptr, false); }

void* pool_alloc(size_t bytes, cudaStream_t s) {
    size_t want = round_block(bytes);
    poolimpl::Key key(poolimpl::current_device(), s);
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        auto& fl = g_cache.free_blocks[key];
        auto it = fl.lower_bound(want);
        if (it != fl.end() && it->first <= want + want / 4 + 4096) {      // best fit, at most 25 % slack
            void* p = it->second; size_t sz = it->first;
            fl.erase(it); g_cache.live[p] = {sz, key};
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cudaGetLastError(); poolimpl::trim(key.first, nullptr, false); e = cudaMalloc(&p, want); }
    check(e, "cudaMalloc", __FILE__, __LINE__);
    std::lock_guard<std::mutex> lk(g_cache.mu);
    g_cache.live[p] = {want, key};
    return p;
}

void pool_free(void* p, cudaStream_t) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    auto it = g_cache.live.find(p);
    if (it == g_cache.live.end()) return;
    g_cache.free_blocks[it->second.second].emplace(it->second.first, p);      // back to the list of the stream it was allocated for
    g_cache.live.erase(it);
}
}  // namespace mmb

// ---------------------------------------------------------------------------------------------- per-kernel event timing
namespace mmb {
namespace profimpl {
std::mutex mu;
std::atomic<bool> on{false};
std::string filter;                                    // empty = off
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pairs;
std::vector<std::string> names;                        // kernel of each pair
bool open_pair = false;
}  // namespace profimpl
void profile_before(const char* name, cudaStream_t s) {
    using namespace profimpl;
    if (!on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(mu);
    if (filter.empty() || (filter != "*" && !strstr(name, filter.c_str()))) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, s);
    pairs.emplace_back(a, b); names.emplace_back(name); open_pair = true;
}
void profile_after(const char* name, cudaStream_t s) {
    using namespace profimpl;
    if (!on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(mu);
    if (!open_pair) return;
    (void)name;
    cudaEventRecord(pairs.back().second, s); open_pair = false;
}
}  // namespace mmb

using namespace mmb;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) { g_error = msg; return code; }

int map_error(const Error& e) {
    int c = e.code;
    if (c == 10) return MMB_ERR_CUDA;
    if (c == 1 || c == 11 || c == 61) return MMB_ERR_ARG;                          // null destination, misaligned offsets, partition shape
    if (c == 13 || c == 21 || c == 40 || c == 41 || c == 42 || c == 60 || c == 70) return MMB_ERR_INPUT;
    if (c == 12 || c == 20 || (c >= 30 && c <= 39) || (c >= 50 && c <= 59) || (c >= 71 && c <= 75)) return MMB_ERR_PARAM;   // 71-75: sizes the engine does not support
    return MMB_ERR_INTERNAL;
}

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr; cudaStream_t s;
    explicit Timer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, s); }
    float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

struct RecordInfo { std::string name; u64 length; };

}  // namespace

struct mmb_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;                 // uploads of mmb_load_fasta, so that they overlap the text build
    std::vector<cudaEvent_t> upload_events;
    unsigned long long launches_at_create = 0;

    bool loaded = false, found = false;
    TextLayout text;
    std::vector<std::vector<RecordInfo>> records;
    DevBuf<u8> raw;                     // device copy of the input images (host-load path)
    DevBuf<u16> doc_coarse;             // document of every 2^coarse_shift-th text position (DetectParams::doc_coarse)
    PfpTriggers trig;                   // trigger masks of the text prefix, scanned while the upload was still running (load_common)
    u32 coarse_shift = 0;

    SuffixArray sa;
    DevBuf<u32> lcp;
    DevBuf<u8> bwt;
    Reports reports;
    MumTable mums;
    MemTable mems;
    DevBuf<u16> thresh;
    u64 thresh_count = 0;
    bool mum_mode = true;
    DetectParams prm{};
    mmb_stats stats{};

    DevBuf<u64> full_sa;                // route 2 with keep_arrays: the whole stream
    DevBuf<u32> full_lcp;
    u32 route = 1;

    MergePartition merged;              // result of the last mmb_anchor_merge
    DevBuf<u8> merged_text;
    u64 merged_text_bytes = 0, merged_anchor_len = 0;
    bool has_merged = false, has_merged_text = false;
};

namespace {

// ---------------------------------------------------------------------------------------------- parameters
void normalise(const mmb_options& o, u32 N, DetectParams& p, bool& mum_mode, bool& merge) {
    long n = (long)N, k = o.num_distinct_docs, f = o.rare_freq, F = o.max_mem_freq;
    merge = o.merge || o.anchor_merge;
    if (o.cli_normalise) {
        if (f < 0) throw Error(50, "Per-sequence MEM frequency must be > 0 (or 0 for no limit).");
        if (k < -n) k = 2; else if (k <= 0) k = n + k; else if (k == 1) k = 2; else if (k >= n) k = n;
        if (merge && k != n) throw Error(51, "Merging not available for partial multi-MUM/MEMs");
        if (merge && f != 1) throw Error(52, "Merging not available for multi-MEMs");
        if (F < -n || F == 1) F = 0; else if (F < 0) F = n + F;
        if (f > 0 && (F == 0 || F > f * n)) F = f * n;
    } else {
        if (k == 0) k = n;
        if (f < 0 || F < 0 || k < 0) throw Error(53, "negative frequency / document count");
    }
    p.min_len = o.min_match_len; p.num_distinct = (u32)k; p.max_doc_freq = (u32)f; p.max_freq = (u32)F; p.ndocs = N; p.merge = merge ? 1 : 0;
    mum_mode = (f == 1);
}

// name of the FASTA record whose header marker sits at img[at]: up to the first whitespace (kseq, include/kseq.h:186-187)
std::string record_name(const u8* img, u64 size, u64 at) {
    std::string s;
    for (u64 i = at + 1; i < size && !isspace(img[i]); ++i) s += (char)img[i];
    return s;
}

// record names when the images only exist on the device: up to NAME_MAX_LEN bytes after each header marker, cut at the first whitespace
constexpr u32 NAME_MAX_LEN = 255;
__global__ void gather_record_names(const u8* __restrict__ raw, const u64* __restrict__ pos, const u64* __restrict__ end, u32 nrec, u8* __restrict__ out) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrec) return;
    u8* o = out + (u64)r * (NAME_MAX_LEN + 1);
    u32 k = 0;
    for (u64 i = pos[r] + 1; i < end[r] && k < NAME_MAX_LEN; ++i) {
        u8 c = raw[i];
        if (c == ' ' || (c >= 9 && c <= 13)) break;          // isspace
        o[k++] = c;
    }
    o[k] = 0;
}

void finish_load(mmb_engine* e, const u8* d_raw, const u8* const* host_imgs, const std::vector<u64>& off, const std::vector<u64>& size, bool raw_mode) {
    u32 N = (u32)off.size();
    for (u32 d = 0; d < N; ++d)
        if (e->text.h_doc_len[d] == 0) throw Error(40, "Empty input file found (document " + std::to_string(d) + ")");
    e->records.assign(N, {});
    if (raw_mode) return;
    // FASTA records: boundaries from the device scan, names from the host image when we have one
    const auto& pos = e->text.h_rec_pos; const auto& before = e->text.h_rec_before;
    std::vector<u8> dev_names;
    if (!host_imgs && d_raw && !pos.empty()) {
        // images were uploaded by the caller: fetch the names from the device copy (header lines longer than 255 bytes are cut)
        std::vector<u64> ends(pos.size());
        size_t rr = 0;
        for (u32 d = 0; d < N; ++d) while (rr < pos.size() && pos[rr] < off[d] + size[d]) { ends[rr] = off[d] + size[d]; ++rr; }
        DevBuf<u64> d_pos, d_end; DevBuf<u8> d_names;
        d_pos.alloc(pos.size(), e->stream); d_end.alloc(pos.size(), e->stream); d_names.alloc(pos.size() * (NAME_MAX_LEN + 1), e->stream);
        MMB_CUDA(cudaMemcpyAsync(d_pos.p, pos.data(), pos.size() * 8, cudaMemcpyHostToDevice, e->stream));
        MMB_CUDA(cudaMemcpyAsync(d_end.p, ends.data(), ends.size() * 8, cudaMemcpyHostToDevice, e->stream));
        MMB_LAUNCH(gather_record_names, ceil_div(pos.size(), 128), 128, 0, e->stream, d_raw, (const u64*)d_pos.p, (const u64*)d_end.p, (u32)pos.size(), d_names.p);
        dev_names = to_host(d_names.p, pos.size() * (NAME_MAX_LEN + 1), e->stream);
    }
    u64 cum = 0; size_t r = 0;
    for (u32 d = 0; d < N; ++d) {
        u64 doc_end_bases = cum + e->text.h_doc_len[d];
        while (r < pos.size() && pos[r] < off[d] + size[d]) {
            u64 next = (r + 1 < pos.size() && pos[r + 1] < off[d] + size[d]) ? before[r + 1] : doc_end_bases;
            RecordInfo ri; ri.length = next - before[r];
            if (host_imgs) ri.name = record_name(host_imgs[d], size[d], pos[r] - off[d]);
            else if (!dev_names.empty()) ri.name = (const char*)&dev_names[r * (NAME_MAX_LEN + 1)];
            e->records[d].push_back(std::move(ri));
            ++r;
        }
        cum = doc_end_bases;
    }
}

// coarse position -> document table for the detector's candidate checks (DetectParams::doc_coarse)
void build_doc_coarse(mmb_engine* e) {
    const std::vector<u64>& ds = e->text.h_doc_start;
    u32 shift = 0;
    while (((e->text.n >> shift) + 1) > 4096) ++shift;
    std::vector<u16> tab((e->text.n >> shift) + 1);
    u32 d = 0;
    for (size_t i = 0; i < tab.size(); ++i) {
        u64 pos = (u64)i << shift;
        while (d + 1 < ds.size() - 1 && ds[d + 1] <= pos) ++d;
        tab[i] = (u16)d;
    }
    e->doc_coarse.alloc(tab.size(), e->stream); e->coarse_shift = shift;
    MMB_CUDA(cudaMemcpyAsync(e->doc_coarse.p, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice, e->stream));
    MMB_CUDA(cudaStreamSynchronize(e->stream));       // tab lives on this stack frame
}

// groups: file ranges [f0, f1) whose bytes are on the device once the matching event has fired (nullptr / empty: everything is there)
struct UploadGroup { size_t f0, f1; cudaEvent_t ready; };

void load_common(mmb_engine* e, const u8* d_raw, const u8* const* host_imgs, const std::vector<u64>& off, const std::vector<u64>& size,
                 bool revcomp, bool raw_mode, const std::vector<UploadGroup>* groups = nullptr) {
    if (off.size() < 2) throw Error(41, "Multiple FASTA inputs required. Perhaps split a multi-FASTA into multiple files?");
    e->found = false; e->loaded = false;
    e->trig.reset();
    Timer t(e->stream); t.start();
    if (groups && groups->size() > 1) {
        // the text build of the files that have arrived runs while the copy stream is still uploading the next ones
        // ... and so does the first pass of the prefix-free parse: the stream would otherwise idle between groups (the copy is the slower side)
        text_begin(size, revcomp, e->text, e->stream);
        u64 worst = 0;
        for (u64 b : size) worst += (b + 1) * (revcomp ? 2 : 1);
        const char* amin = getenv("MMB_TRIGGER_AHEAD_MIN");                     // tests lower it; small texts may take the direct route anyway
        if (worst >= (amin ? (u64)atoll(amin) : (1ull << 24))) pfp_triggers_begin(worst, e->trig, e->stream);
        for (const UploadGroup& g : *groups) {
            MMB_CUDA(cudaStreamWaitEvent(e->stream, g.ready, 0));
            text_append(d_raw, off, size, g.f0, g.f1, raw_mode, e->text, e->stream);
            pfp_triggers_extend(e->text.text.p, e->text.n, e->trig, e->stream);
        }
        text_finish(e->text, e->stream);
    } else {
        if (groups && groups->size() == 1) MMB_CUDA(cudaStreamWaitEvent(e->stream, (*groups)[0].ready, 0));
        build_text(d_raw, off, size, revcomp, raw_mode, e->text, e->stream);
    }
    e->stats.ms_text = t.stop();
    build_doc_coarse(e);
    finish_load(e, d_raw, host_imgs, off, size, raw_mode);
    e->stats.text_symbols = e->text.n; e->stats.input_bases = e->text.total_bases;
    e->loaded = true;
}

// Host images -> device, on the engine's copy stream, in groups of at least 256 MB: one event per group tells the text build which files
// are in place (with pinned host memory the copies run asynchronously at PCIe speed; from pageable memory they are staged by the driver
// and simply finish earlier than the events are waited for).
void upload(mmb_engine* e, const u8* const* imgs, const u64* sizes, size_t n, std::vector<u64>& off, std::vector<u64>& size,
            std::vector<UploadGroup>& groups) {
    u64 at = 0;
    for (size_t i = 0; i < n; ++i) { off.push_back(at); size.push_back(sizes[i]); at += (sizes[i] + 15) & ~15ull; }
    e->raw.alloc(at + 64, e->stream);
    if (!e->copy_stream) MMB_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    MMB_CUDA(cudaStreamSynchronize(e->stream));        // nothing of an earlier call may still read the buffer the copy stream is about to fill
    MMB_CUDA(cudaMemsetAsync(e->raw.p + at, 0, 64, e->copy_stream));
    const char* gb = getenv("MMB_UPLOAD_GROUP_BYTES");                            // tests make the groups small
    const u64 GROUP_BYTES = gb ? (u64)atoll(gb) : (256ull << 20);
    size_t f0 = 0; u64 acc = 0;
    for (size_t i = 0; i < n; ++i) {
        if (sizes[i]) MMB_CUDA(cudaMemcpyAsync(e->raw.p + off[i], imgs[i], sizes[i], cudaMemcpyHostToDevice, e->copy_stream));
        acc += sizes[i];
        if (acc >= GROUP_BYTES || i + 1 == n) {
            while (e->upload_events.size() <= groups.size()) { cudaEvent_t ev; MMB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); e->upload_events.push_back(ev); }
            cudaEvent_t ev = e->upload_events[groups.size()];
            MMB_CUDA(cudaEventRecord(ev, e->copy_stream));
            groups.push_back(UploadGroup{f0, i + 1, ev});
            f0 = i + 1; acc = 0;
        }
    }
}

template <class F>
int guarded(F&& f) {
    try { f(); return MMB_OK; }
    catch (const Error& e) { return fail(map_error(e), e.what()); }
    catch (const std::exception& e) { return fail(MMB_ERR_INTERNAL, e.what()); }
    catch (...) { return fail(MMB_ERR_INTERNAL, "unknown exception"); }
}

template <class T>
void copy_out(const T* d, T* h, u64 n, cudaStream_t s) {
    if (h && n) MMB_CUDA(cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, s));
}

__global__ void widen_u32(const u32* __restrict__ in, u64 n, u64* __restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
// 5-byte little-endian entries (the reference's .sa / .lcp files) -> u64 / u32; *overflow is set when an lcp does not fit 32 bits
__global__ void unpack5(const u8* __restrict__ raw, u64 n, u64* __restrict__ out64, u32* __restrict__ out32, u32* __restrict__ overflow) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8* q = raw + 5 * i;
    u64 v = (u64)q[0] | ((u64)q[1] << 8) | ((u64)q[2] << 16) | ((u64)q[3] << 24) | ((u64)q[4] << 32);
    if (out64) out64[i] = v;
    if (out32) { if (v > 0xfffffffeull) { *overflow = 1; v = 0xfffffffeull; } out32[i] = (u32)v; }
}
__global__ void bwt_from_sa64(const u8* __restrict__ text, const u64* __restrict__ sa, u64 m, u8* __restrict__ bwt) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) { u64 p = sa[r]; bwt[r] = p ? text[p - 1] : (u8)0; }
}
__global__ void strands_to_flags(const u8* __restrict__ in, u64 n, u8* __restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] == '+' ? 1 : 0;
}
__global__ void mem_widen(const u64* __restrict__ off, const u32* __restrict__ doc, u64 n, i64* __restrict__ off_out, u64* __restrict__ doc_out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { off_out[i] = (i64)off[i]; doc_out[i] = doc[i]; }
}

}  // namespace

extern "C" {

const char* mmb_last_error(void) { return g_error.c_str(); }

void mmb_default_options(mmb_options* o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->min_match_len = 20; o->num_distinct_docs = 0; o->rare_freq = 1; o->max_mem_freq = 0; o->cli_normalise = 1; o->want_text = 1;
}

int mmb_create(int device, void* cuda_stream, mmb_engine** out) {
    if (!out) return fail(MMB_ERR_ARG, "out must be non-null");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return fail(MMB_ERR_NO_DEVICE, "no usable CUDA device " + std::to_string(device) + ": this library has no CPU path");
    }
    return guarded([&] {
        MMB_CUDA(cudaSetDevice(device));
        auto* e = new mmb_engine();
        e->device = device;
        if (cuda_stream) e->stream = (cudaStream_t)cuda_stream;
        else { MMB_CUDA(cudaStreamCreate(&e->stream)); e->own_stream = true; }
#ifndef MMB_SIMT
        cudaMemPool_t pool;                       // keep freed blocks cached: stages allocate and free in a steady pattern
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
#endif
        e->launches_at_create = g_kernel_launches.load();
        *out = e;
    });
}

void mmb_destroy(mmb_engine* e) {
    if (!e) return;
    cudaStreamSynchronize(e->stream);
    cudaStream_t s = e->stream; bool own = e->own_stream; int dev = e->device;
    cudaSetDevice(dev);
    if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
    for (cudaEvent_t ev : e->upload_events) cudaEventDestroy(ev);
    delete e;
    poolimpl::trim(dev, s, true);
    if (own) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
}

void* mmb_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void mmb_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int mmb_load_fasta(mmb_engine* e, const uint8_t* const* files, const uint64_t* sizes, size_t nfiles, int use_revcomp) {
    if (!e || !files || !sizes) return fail(MMB_ERR_ARG, "null argument");
    return guarded([&] {
        MMB_CUDA(cudaSetDevice(e->device));
        std::vector<u64> off, size; std::vector<UploadGroup> groups;
        upload(e, files, sizes, nfiles, off, size, groups);
        load_common(e, e->raw.p, files, off, size, use_revcomp != 0, false, &groups);
    });
}

int mmb_load_fasta_device(mmb_engine* e, const void* d_raw, const uint64_t* offsets, const uint64_t* sizes, size_t nfiles, int use_revcomp) {
    if (!e || !d_raw || !offsets || !sizes) return fail(MMB_ERR_ARG, "null argument");
    return guarded([&] {
        MMB_CUDA(cudaSetDevice(e->device));
        std::vector<u64> off(offsets, offsets + nfiles), size(sizes, sizes + nfiles);
        for (size_t i = 0; i < nfiles; ++i)
            if (off[i] & 15) throw Error(11, "mmb_load_fasta_device: offsets[" + std::to_string(i) + "] is not a multiple of 16");
        load_common(e, (const u8*)d_raw, nullptr, off, size, use_revcomp != 0, false);
    });
}

int mmb_load_records(mmb_engine* e, const char* const* records, const size_t* n_records, size_t ndocs, int use_revcomp) {
    if (!e || (!records && ndocs) || !n_records) return fail(MMB_ERR_ARG, "null argument");
    return guarded([&] {
        MMB_CUDA(cudaSetDevice(e->device));
        // concatenate the records of each document on the host (they are scattered C strings), then one copy per document
        std::vector<std::string> docs(ndocs);
        std::vector<std::vector<RecordInfo>> recs(ndocs);
        size_t r = 0;
        for (size_t d = 0; d < ndocs; ++d)
            for (size_t i = 0; i < n_records[d]; ++i, ++r) {
                const char* sp = records[r] ? records[r] : "";
                docs[d] += sp;
                recs[d].push_back(RecordInfo{"", (u64)strlen(sp)});
            }
        std::vector<const u8*> imgs; std::vector<u64> sizes;
        for (auto& d : docs) { imgs.push_back((const u8*)d.data()); sizes.push_back(d.size()); }
        std::vector<u64> off, size; std::vector<UploadGroup> groups;
        upload(e, imgs.data(), sizes.data(), ndocs, off, size, groups);
        MMB_CUDA(cudaStreamSynchronize(e->copy_stream));  // the host strings die at the end of this scope
        load_common(e, e->raw.p, nullptr, off, size, use_revcomp != 0, true);
        e->records = std::move(recs);
    });
}

int mmb_find(mmb_engine* e, const mmb_options* opts) {
    if (!e || !opts) return fail(MMB_ERR_ARG, "null argument");
    if (!e->loaded || !e->text.text.p) return fail(MMB_ERR_STATE, "mmb_find before a successful mmb_load_*");
    return guarded([&] {
        MMB_CUDA(cudaSetDevice(e->device));
        cudaStream_t s = e->stream;
        e->found = false;
        u32 N = (u32)e->text.h_doc_len.size();
        bool merge = false;
        normalise(*opts, N, e->prm, e->mum_mode, merge);
        e->prm.revcomp = e->text.revcomp ? 1 : 0;
        e->prm.doc_coarse = e->doc_coarse.p; e->prm.coarse_shift = e->coarse_shift;

        Timer total(s), t(s);
        double host0 = host_now_ms(), a0 = g_host_alloc_ms, f0 = g_host_free_ms, y0 = g_host_sync_ms;
        unsigned long long na0 = g_host_allocs, ns0 = g_host_syncs;
        total.start();
        e->thresh_count = 0; e->thresh.release();
        if (merge) {
            e->thresh_count = 2 * (e->text.h_doc_len[0] + 1);       // candidate_thresh.resize(doc_lens[0] * 2), include/mem_finder.hpp:79-81
            e->thresh.alloc(e->thresh_count, s); e->thresh.zero();
        }
        u16* thresh = merge ? e->thresh.p : nullptr;
        bool want_text = opts->want_text != 0;
        u64 n = e->text.n;
        // route: forced, or by size, or -- for texts either route can hold -- by how repetitive the collection is: the parse and the
        // dictionary (stages 1-3 of the PFP route) are cheap, and the route pays off when the dictionary is a small fraction of the text
        PfpIndex ix; PfpParse pp;
        bool parsed = false;
        // -f 0 without -F: intervals of any size.  The interval walk then needs the whole suffix stream in ONE view: the direct route below
        // 2^31 symbols, or the prefix-free-parse route when the PRUNED stream fits one tile (segment pruning holds for any interval size:
        // a segment is left out because it is too small or has a uniform BWT)
        const bool unbounded = e->prm.max_freq == 0 && e->prm.max_doc_freq == 0;
        const bool prune_ok = !(opts->keep_arrays || opts->no_prune);
        if (unbounded && !prune_ok && (opts->route == 2 || n + 1 >= (1ull << 31)))
            throw Error(31, "unbounded total frequency (-f 0 with -F 0) needs the whole suffix stream in one view: not together with keep_arrays / no_prune on the prefix-free-parse route");
        e->prm.whole_stream = 0;
        const u32 shards = opts->shard_count > 1 ? opts->shard_count : 1;
        if (shards > 1) {
            if (opts->shard_index >= shards) throw Error(32, "shard_index must be below shard_count");
            if (merge || !prune_ok || opts->route == 1) throw Error(32, "a sharded call needs the pruned prefix-free-parse route: not with merge, keep_arrays, no_prune or route = 1");
        }
        bool use_pfp = (opts->route == 2 || shards > 1 || (opts->route == 0 && n + 1 >= (1ull << 31)));
        Timer tparse(s); float ms_parse = 0;
        if (opts->route == 0 && !use_pfp && !unbounded && n >= (1u << 16)) {      // unbounded + auto: the direct route whenever it fits
            tparse.start();
            pfp_parse(e->text.text.p, n, ix, pp, s, &e->trig); parsed = true;
            ms_parse = tparse.stop();
            use_pfp = (u64)ix.dsize * 3 <= n;
            if (!use_pfp) { ix = PfpIndex(); pp = PfpParse(); }
        }
        e->route = use_pfp ? 2 : 1;
        e->full_sa.release(); e->full_lcp.release(); e->bwt.release();
        e->stats.pfp_phrases = e->stats.pfp_distinct_phrases = e->stats.pfp_dict_bytes = e->stats.pfp_tiles = 0;
        e->stats.stream_suffixes = n + 1; e->stats.pruned = 0; e->stats.threshold_segments = 0;
        if (!use_pfp) {
            e->trig.reset();
            t.start();
            build_suffix_array(e->text.text.p, n, e->sa, s);
            e->stats.ms_sa = t.stop() + ms_parse;
            t.start();
            build_lcp(e->text.text.p, n, e->sa, e->lcp, s);
            e->stats.ms_lcp = t.stop();
            e->sa.isa.release();
            t.start();
            e->prm.whole_stream = 1;
            detect_intervals<u32>(e->lcp.p, e->sa.sa.p, (const u8*)nullptr, e->text.text.p, e->sa.m, 0, e->prm, e->text.doc_start.p, thresh, e->reports, s);
            MMB_CUDA(cudaStreamSynchronize(s));
            e->stats.ms_detect = t.stop();
            t.start();
            if (e->mum_mode && opts->trim_min_len) {
                mum_rows<u32>(e->reports, e->sa.sa.p, e->prm, e->text.doc_start.p, e->text.doc_len.p, e->mums, s);
                trim_mum_table(e->mums, N, e->text.doc_len.p, opts->trim_min_len, s);
                if (want_text) mum_text(e->mums, N, s);
            } else if (e->mum_mode) emit_mums<u32>(e->reports, e->sa.sa.p, e->prm, e->text.doc_start.p, e->text.doc_len.p, want_text, e->mums, s);
            else emit_mems<u32>(e->reports, e->sa.sa.p, e->prm, e->text.doc_start.p, e->text.doc_len.p, want_text, e->mems, s);
            e->stats.ms_emit = t.stop();
            e->stats.sa_rounds = e->sa.stats.rounds; e->stats.sa_sorted_elements = e->sa.stats.groups_sorted;
        } else {
            e->sa.sa.release(); e->sa.isa.release(); e->lcp.release();
            t.start();
            if (!parsed) pfp_parse(e->text.text.p, n, ix, pp, s, &e->trig);
            // segment pruning (pfp.h): off when the caller wants the whole suffix stream back
            PfpPrune prune;
            prune.enabled = (opts->keep_arrays || opts->no_prune) ? 0 : 1;
            prune.min_len = e->prm.min_len; prune.size_lo = std::max<u32>(2, e->prm.num_distinct); prune.merge = merge ? 1 : 0; prune.ndocs = N;
            prune.doc_coarse = e->doc_coarse.p; prune.coarse_shift = e->coarse_shift; prune.doc_start = e->text.doc_start.p;
            prune.shard_index = shards > 1 ? opts->shard_index : 0; prune.shard_count = shards;
            pfp_finish(e->text.text.p, n, ix, pp, s, prune);
            if (shards > 1 && (!ix.pruned || ix.full_stream))
                throw Error(32, "a sharded call needs segment pruning, and this collection could not be pruned (-l below 11, or a very long low-complexity phrase)");
            pp = PfpParse();
            e->stats.ms_sa = t.stop() + ms_parse;
            e->stats.ms_lcp = 0;
            e->stats.pfp_phrases = ix.np; e->stats.pfp_distinct_phrases = ix.nd; e->stats.pfp_dict_bytes = ix.dsize;
            e->stats.sa_rounds = ix.stats_d.rounds + ix.stats_p.rounds; e->stats.sa_sorted_elements = ix.stats_d.groups_sorted + ix.stats_p.groups_sorted;
            // how far back an interval can reach: the detector's halo
            u64 hi = e->prm.max_freq;
            if (e->prm.max_doc_freq) { u64 bound = (u64)e->prm.max_doc_freq * N; hi = hi ? std::min(hi, bound) : bound; }
            u32 halo = (u32)std::min<u64>((hi + 8) & ~3ull, 1u << 22);
            // ranks per stream tile: 17 B per rank of tile buffers; C2 measured 745 / 731 / 722 ms per step at 2^28 / 2^29 / 2^30
            u64 tile_ranks = opts->tile_ranks ? opts->tile_ranks : (1ull << 30);
            if (hi == 0) {          // unbounded: one tile or nothing
                if (!ix.pruned || ix.full_stream) throw Error(31, "unbounded total frequency (-f 0 with -F 0): the suffix stream could not be pruned (-l below 11, or a very long low-complexity phrase); use the direct route (text below 2^31 symbols)");
                tile_ranks = 1ull << 62; halo = 0; e->prm.whole_stream = 1;
            }
            struct Consumer : TileConsumer {
                mmb_engine* e; cudaStream_t s; u16* thresh; bool keep; std::vector<MumTable> mparts; std::vector<MemTable> eparts; u32 tiles = 0;
                void consume(const StreamTile& tile) override {
                    ++tiles;
                    if (getenv("MMB_DEBUG_TILE")) {
                        std::vector<u64> hs = to_host(tile.sa, tile.count, s); std::vector<u32> hl = to_host(tile.lcp, tile.count, s); std::vector<u8> hb = to_host(tile.bwt, tile.count, s);
                        for (u64 r = 0; r < tile.count; ++r) fprintf(stderr, "[tile] %llu sa %llu lcp %u bwt %u\n", (unsigned long long)(tile.rank0 + r), (unsigned long long)hs[r], hl[r], hb[r]);
                    }
                    Reports rep;
                    detect_intervals<u64>(tile.lcp, tile.sa, tile.bwt, e->text.text.p, tile.count, tile.first, e->prm, e->text.doc_start.p, thresh, rep, s, tile.doc);
                    if (e->mum_mode) { MumTable part; mum_rows<u64>(rep, tile.sa, e->prm, e->text.doc_start.p, e->text.doc_len.p, part, s); mparts.push_back(std::move(part)); }
                    else { MemTable part; mem_rows<u64>(rep, tile.sa, e->prm, e->text.doc_start.p, e->text.doc_len.p, part, s); eparts.push_back(std::move(part)); }
                    if (keep) {
                        u64 fresh = tile.count - tile.first, at = tile.rank0 + tile.first;
                        MMB_CUDA(cudaMemcpyAsync(e->full_sa.p + at, tile.sa + tile.first, fresh * 8, cudaMemcpyDeviceToDevice, s));
                        MMB_CUDA(cudaMemcpyAsync(e->full_lcp.p + at, tile.lcp + tile.first, fresh * 4, cudaMemcpyDeviceToDevice, s));
                    }
                    MMB_CUDA(cudaStreamSynchronize(s));
                }
            } consumer;
            consumer.e = e; consumer.s = s; consumer.thresh = thresh; consumer.keep = opts->keep_arrays != 0;
            if (consumer.keep) { e->full_sa.alloc(n + 1, s); e->full_lcp.alloc(n + 1, s); }
            t.start();
            // experimental (not yet measured on a GPU, off by default): a document-id byte per rank for the -M candidate checks
            DocLookup dl{e->doc_coarse.p, e->coarse_shift, e->text.doc_start.p, N};
            const bool doc_stream = merge && N < 255 && getenv("MMB_DOC_STREAM") && atoi(getenv("MMB_DOC_STREAM"));
            stream_suffixes(e->text.text.p, ix, tile_ranks, halo, consumer, s, doc_stream ? &dl : nullptr);
            if (merge) segment_thresholds(ix, dl, thresh, s);
            e->stats.stream_suffixes = ix.stream_ranks + 1; e->stats.pruned = ix.full_stream ? 0 : 1;
            e->stats.threshold_segments = ix.n_closers;
            e->stats.ms_detect = t.stop();
            e->stats.pfp_tiles = consumer.tiles;
            t.start();
            if (e->mum_mode) {
                concat_mums(consumer.mparts, N, e->mums, s);
                if (opts->trim_min_len) trim_mum_table(e->mums, N, e->text.doc_len.p, opts->trim_min_len, s);
                if (want_text) mum_text(e->mums, N, s);
            }
            else { concat_mems(consumer.eparts, e->mems, s); if (want_text) mem_text(e->mems, s); }
            e->stats.ms_emit = t.stop();
        }
        e->stats.ms_total = total.stop();
        if (getenv("MMB_PROFILE_HOST"))
            fprintf(stderr, "[host] find %.2f ms wall; cudaMallocAsync %.2f ms in %llu calls; cudaFreeAsync %.2f ms; readback syncs %.2f ms in %llu; device %.2f ms\n",
                    host_now_ms() - host0, g_host_alloc_ms - a0, g_host_allocs - na0, g_host_free_ms - f0, g_host_sync_ms - y0, g_host_syncs - ns0, e->stats.ms_total);
        e->stats.num_matches = e->mum_mode ? e->mums.count : e->mems.count;
        e->stats.route = e->route;
        e->found = true;
    });
}

int mmb_find_from_arrays(mmb_engine* e, const uint64_t* doc_lengths, size_t ndocs, int use_revcomp, const uint8_t* sa5, const uint8_t* lcp5,
                         const uint8_t* bwt, uint64_t entries, const mmb_options* opts) {
    if (!e || !doc_lengths || !sa5 || !lcp5 || !bwt || !opts) return fail(MMB_ERR_ARG, "null argument");
    return guarded([&] {
        MMB_CUDA(cudaSetDevice(e->device));
        cudaStream_t s = e->stream;
        e->found = false; e->loaded = false;
        if (ndocs < 2) throw Error(41, "Multiple FASTA inputs required. Perhaps split a multi-FASTA into multiple files?");
        u32 N = (u32)ndocs;
        // document layout from the lengths alone (RefBuilder(prefix), src/ref_builder.cpp:140-168 + build_bv :171-191); no text
        TextLayout& tl = e->text;
        tl.text.release(); tl.revcomp = use_revcomp != 0; tl.h_rec_pos.clear(); tl.h_rec_before.clear();
        tl.h_doc_len.assign(doc_lengths, doc_lengths + ndocs); tl.h_doc_start.assign(ndocs + 1, 0); tl.total_bases = 0;
        for (u32 d = 0; d < N; ++d) {
            tl.h_doc_start[d + 1] = tl.h_doc_start[d] + (tl.revcomp ? 2 : 1) * (tl.h_doc_len[d] + 1);
            tl.total_bases += tl.h_doc_len[d];
        }
        tl.n = tl.h_doc_start[N];
        if (entries == 0 || entries > tl.n + 1) throw Error(42, "array length does not match the sequence lengths");
        tl.doc_start.alloc(N + 1, s); tl.doc_len.alloc(N, s);
        MMB_CUDA(cudaMemcpyAsync(tl.doc_start.p, tl.h_doc_start.data(), (N + 1) * 8, cudaMemcpyHostToDevice, s));
        MMB_CUDA(cudaMemcpyAsync(tl.doc_len.p, tl.h_doc_len.data(), N * 8, cudaMemcpyHostToDevice, s));
        MMB_CUDA(cudaStreamSynchronize(s));
        e->records.assign(N, {});
        build_doc_coarse(e);
        e->stats = mmb_stats{}; e->stats.text_symbols = tl.n; e->stats.input_bases = tl.total_bases;

        bool merge = false;
        normalise(*opts, N, e->prm, e->mum_mode, merge);
        e->prm.revcomp = tl.revcomp ? 1 : 0;
        e->prm.doc_coarse = e->doc_coarse.p; e->prm.coarse_shift = e->coarse_shift;
        Timer total(s), t(s);
        total.start();
        e->thresh_count = 0; e->thresh.release();
        if (merge) { e->thresh_count = 2 * (tl.h_doc_len[0] + 1); e->thresh.alloc(e->thresh_count, s); e->thresh.zero(); }
        u16* thresh = merge ? e->thresh.p : nullptr;
        e->sa.sa.release(); e->sa.isa.release(); e->lcp.release(); e->bwt.release(); e->full_sa.release(); e->full_lcp.release();
        e->route = 3;

        u64 hi = e->prm.max_freq;
        if (e->prm.max_doc_freq) { u64 bound = (u64)e->prm.max_doc_freq * N; hi = hi ? std::min(hi, bound) : bound; }
        u64 halo = std::min<u64>((hi + 8) & ~3ull, 1u << 22);
        u64 tile = opts->tile_ranks ? opts->tile_ranks : (1ull << 26);
        if (hi == 0) {                                  // no frequency limit: only with the whole stream in one view
            if (entries >= (1ull << 31)) throw Error(31, "unbounded total frequency (-f 0 with -F 0) needs the whole suffix stream in one view (below 2^31 ranks)");
            tile = entries; halo = 0;
        }
        e->prm.whole_stream = tile >= entries ? 1 : 0;
        u64 cap = halo + std::min(tile, entries);
        DevBuf<u8> raw_sa, raw_lcp, d_bwt; DevBuf<u64> d_sa; DevBuf<u32> d_lcp, overflow;
        raw_sa.alloc(cap * 5, s); raw_lcp.alloc(cap * 5, s); d_bwt.alloc(cap + 64, s); d_sa.alloc(cap, s); d_lcp.alloc(cap + 64, s);
        overflow.alloc(1, s); overflow.zero();
        std::vector<MumTable> mparts; std::vector<MemTable> eparts;
        t.start();
        for (u64 start = 0; start < entries; start += tile) {
            u64 first = std::min(halo, start), view0 = start - first, count = first + std::min(tile, entries - start);
            MMB_CUDA(cudaMemcpyAsync(raw_sa.p, sa5 + view0 * 5, count * 5, cudaMemcpyHostToDevice, s));
            MMB_CUDA(cudaMemcpyAsync(raw_lcp.p, lcp5 + view0 * 5, count * 5, cudaMemcpyHostToDevice, s));
            MMB_CUDA(cudaMemcpyAsync(d_bwt.p, bwt + view0, count, cudaMemcpyHostToDevice, s));
            MMB_CUDA(cudaMemsetAsync(d_bwt.p + count, 0, 64, s));
            MMB_CUDA(cudaMemsetAsync(d_lcp.p + count, 0, 64 * 4, s));
            MMB_LAUNCH(unpack5, ceil_div(count, 256), 256, 0, s, (const u8*)raw_sa.p, count, d_sa.p, (u32*)nullptr, overflow.p);
            MMB_LAUNCH(unpack5, ceil_div(count, 256), 256, 0, s, (const u8*)raw_lcp.p, count, (u64*)nullptr, d_lcp.p, overflow.p);
            Reports rep;
            detect_intervals<u64>(d_lcp.p, d_sa.p, d_bwt.p, (const u8*)nullptr, count, first, e->prm, tl.doc_start.p, thresh, rep, s);
            if (e->mum_mode) { MumTable part; mum_rows<u64>(rep, d_sa.p, e->prm, tl.doc_start.p, tl.doc_len.p, part, s); mparts.push_back(std::move(part)); }
            else { MemTable part; mem_rows<u64>(rep, d_sa.p, e->prm, tl.doc_start.p, tl.doc_len.p, part, s); eparts.push_back(std::move(part)); }
            MMB_CUDA(cudaStreamSynchronize(s));
            ++e->stats.pfp_tiles;
        }
        if (to_host(overflow.p, 1, s)[0]) throw Error(75, "a common prefix longer than 2^32 symbols: not supported");
        e->stats.ms_detect = t.stop();
        t.start();
        bool want_text = opts->want_text != 0;
        if (e->mum_mode) { concat_mums(mparts, N, e->mums, s); if (want_text) mum_text(e->mums, N, s); }
        else { concat_mems(eparts, e->mems, s); if (want_text) mem_text(e->mems, s); }
        e->stats.ms_emit = t.stop();
        e->stats.ms_total = total.stop();
        e->stats.num_matches = e->mum_mode ? e->mums.count : e->mems.count;
        e->stats.route = e->route;
        e->loaded = true; e->found = true;
    });
}

int mmb_get_stats(const mmb_engine* e, mmb_stats* out) {
    if (!e || !out) return fail(MMB_ERR_ARG, "null argument");
    *out = e->stats;
    out->kernel_launches = g_kernel_launches.load() - e->launches_at_create;
    return MMB_OK;
}

size_t mmb_num_docs(const mmb_engine* e) { return (e && e->loaded) ? e->text.h_doc_len.size() : 0; }
int mmb_doc_lengths(const mmb_engine* e, uint64_t* lengths) {
    if (!e || !lengths || !e->loaded) return fail(MMB_ERR_STATE, "no input loaded");
    for (size_t d = 0; d < e->text.h_doc_len.size(); ++d) lengths[d] = e->text.h_doc_len[d];
    return MMB_OK;
}
size_t mmb_num_records(const mmb_engine* e, size_t doc) { return (e && e->loaded && doc < e->records.size()) ? e->records[doc].size() : 0; }
int mmb_record(const mmb_engine* e, size_t doc, size_t rec, const char** name, uint64_t* length) {
    if (!e || !e->loaded || doc >= e->records.size() || rec >= e->records[doc].size()) return fail(MMB_ERR_ARG, "record index out of range");
    if (name) *name = e->records[doc][rec].name.c_str();
    if (length) *length = e->records[doc][rec].length;
    return MMB_OK;
}

int mmb_is_mum_mode(const mmb_engine* e) { return e && e->mum_mode; }
uint64_t mmb_num_matches(const mmb_engine* e) { return (e && e->found) ? (e->mum_mode ? e->mums.count : e->mems.count) : 0; }
uint64_t mmb_text_size(const mmb_engine* e) { return (e && e->found) ? (e->mum_mode ? e->mums.text_bytes : e->mems.text_bytes) : 0; }
const void* mmb_text_device(const mmb_engine* e) { return (e && e->found) ? (e->mum_mode ? (const void*)e->mums.text.p : (const void*)e->mems.text.p) : nullptr; }

int mmb_copy_text(mmb_engine* e, void* host_dst) {
    if (!e || !e->found) return fail(MMB_ERR_STATE, "no results: call mmb_find first");
    return guarded([&] {
        u64 n = mmb_text_size(e);
        if (n && !host_dst) throw Error(1, "null destination");
        copy_out((const u8*)mmb_text_device(e), (u8*)host_dst, n, e->stream);
        MMB_CUDA(cudaStreamSynchronize(e->stream));
    });
}

int mmb_copy_mums(mmb_engine* e, uint32_t* lengths, int64_t* offsets, uint8_t* strands) {
    if (!e || !e->found || !e->mum_mode) return fail(MMB_ERR_STATE, "no MUM results");
    return guarded([&] {
        cudaStream_t s = e->stream;
        u64 n = e->mums.count, cells = n * e->prm.ndocs;
        copy_out(e->mums.len.p, lengths, n, s);
        copy_out(e->mums.off.p, offsets, cells, s);
        DevBuf<u8> flags;
        if (strands && cells) {
            flags.alloc(cells, s);
            MMB_LAUNCH(strands_to_flags, ceil_div(cells, 256), 256, 0, s, e->mums.strand.p, cells, flags.p);
            copy_out(flags.p, strands, cells, s);
        }
        MMB_CUDA(cudaStreamSynchronize(s));
    });
}

int mmb_copy_mum_strand_bits(mmb_engine* e, uint8_t* packed) {
    if (!e || !e->found || !e->mum_mode) return fail(MMB_ERR_STATE, "no MUM results");
    return guarded([&] {
        u64 nbits = e->mums.count * e->prm.ndocs;
        DevBuf<u8> bits;
        pack_strands(e->mums.strand.p, nbits, bits, e->stream);
        copy_out(bits.p, packed, (nbits + 7) / 8, e->stream);
        MMB_CUDA(cudaStreamSynchronize(e->stream));
    });
}

uint64_t mmb_num_mem_occurrences(const mmb_engine* e) { return (e && e->found && !e->mum_mode) ? e->mems.total_occ : 0; }

int mmb_copy_mems(mmb_engine* e, uint32_t* lengths, uint64_t* first, int64_t* offsets, uint64_t* seq_ids, uint8_t* strands) {
    if (!e || !e->found || e->mum_mode) return fail(MMB_ERR_STATE, "no MEM results");
    return guarded([&] {
        cudaStream_t s = e->stream;
        u64 n = e->mems.count, occ = e->mems.total_occ;
        copy_out(e->mems.len.p, lengths, n, s);
        copy_out(e->mems.first.p, first, n + 1, s);
        DevBuf<i64> o64; DevBuf<u64> d64; DevBuf<u8> flags;
        if (occ && (offsets || seq_ids)) {
            o64.alloc(occ, s); d64.alloc(occ, s);
            MMB_LAUNCH(mem_widen, ceil_div(occ, 256), 256, 0, s, e->mems.off.p, e->mems.doc.p, occ, o64.p, d64.p);
            copy_out(o64.p, offsets, occ, s); copy_out(d64.p, seq_ids, occ, s);
        }
        if (occ && strands) {
            flags.alloc(occ, s);
            MMB_LAUNCH(strands_to_flags, ceil_div(occ, 256), 256, 0, s, e->mems.strand.p, occ, flags.p);
            copy_out(flags.p, strands, occ, s);
        }
        MMB_CUDA(cudaStreamSynchronize(s));
    });
}

uint64_t mmb_threshold_count(const mmb_engine* e) { return (e && e->found) ? e->thresh_count : 0; }
int mmb_copy_thresholds(mmb_engine* e, uint16_t* dst) {
    if (!e || !e->found || !e->thresh_count) return fail(MMB_ERR_STATE, "no thresholds: run mmb_find with merge set");
    return guarded([&] { copy_out(e->thresh.p, dst, e->thresh_count, e->stream); MMB_CUDA(cudaStreamSynchronize(e->stream)); });
}

int mmb_copy_arrays(mmb_engine* e, uint64_t* sa, uint64_t* lcp, uint8_t* bwt) {
    if (!e || !e->found) return fail(MMB_ERR_STATE, "no results: call mmb_find first");
    return guarded([&] {
        cudaStream_t s = e->stream;
        if (e->route == 3) throw Error(1, "the arrays were supplied by the caller (mmb_find_from_arrays)");
        if (e->route == 2) {
            if (!e->full_sa.p) throw Error(1, "the suffix stream was not kept: run mmb_find with keep_arrays = 1");
            u64 m = e->text.n + 1;
            copy_out(e->full_sa.p, sa, m, s);
            if (lcp) { DevBuf<u64> w; w.alloc(m, s); MMB_LAUNCH(widen_u32, ceil_div(m, 256), 256, 0, s, (const u32*)e->full_lcp.p, m, w.p); copy_out(w.p, lcp, m, s); MMB_CUDA(cudaStreamSynchronize(s)); }
            if (bwt) { DevBuf<u8> b; b.alloc(m, s); MMB_LAUNCH(bwt_from_sa64, ceil_div(m, 256), 256, 0, s, (const u8*)e->text.text.p, (const u64*)e->full_sa.p, m, b.p); copy_out(b.p, bwt, m, s); MMB_CUDA(cudaStreamSynchronize(s)); }
            MMB_CUDA(cudaStreamSynchronize(s));
            return;
        }
        u64 m = e->sa.m;
        DevBuf<u64> wide; wide.alloc(m, s);
        if (sa) { MMB_LAUNCH(widen_u32, ceil_div(m, 256), 256, 0, s, e->sa.sa.p, m, wide.p); copy_out(wide.p, sa, m, s); MMB_CUDA(cudaStreamSynchronize(s)); }
        if (lcp) { MMB_LAUNCH(widen_u32, ceil_div(m, 256), 256, 0, s, e->lcp.p, m, wide.p); copy_out(wide.p, lcp, m, s); MMB_CUDA(cudaStreamSynchronize(s)); }
        if (bwt) { if (!e->bwt.p) build_bwt(e->text.text.p, e->sa, e->bwt, s); copy_out(e->bwt.p, bwt, m, s); }
        MMB_CUDA(cudaStreamSynchronize(s));
    });
}

int mmb_mum_table_device(mmb_engine* e, const void** lengths, const void** offsets, const void** strands, const void** thresholds) {
    if (!e || !e->found || !e->mum_mode) return fail(MMB_ERR_STATE, "no MUM results");
    if (lengths) *lengths = e->mums.len.p;
    if (offsets) *offsets = e->mums.off.p;
    if (strands) *strands = e->mums.strand.p;
    if (thresholds) *thresholds = e->thresh_count ? e->thresh.p : nullptr;
    return MMB_OK;
}

int mmb_anchor_merge(mmb_engine* e, size_t nparts, const void* const* lengths, const uint64_t* n_mums, const void* const* offsets,
                     const void* const* strands, const void* const* thresholds, const uint64_t* ndocs, uint64_t anchor_len, int inputs_on_device) {
    if (!e || !lengths || !n_mums || !offsets || !strands || !thresholds || !ndocs) return fail(MMB_ERR_ARG, "null argument");
    if (nparts < 2) return fail(MMB_ERR_ARG, "requires at least two partitions");
    return guarded([&] {
        MMB_CUDA(cudaSetDevice(e->device));
        cudaStream_t s = e->stream;
        e->has_merged = false;
        std::vector<MergeSource> src(nparts);
        std::vector<DevBuf<u32>> hl(nparts); std::vector<DevBuf<i64>> ho(nparts); std::vector<DevBuf<u8>> hs(nparts); std::vector<DevBuf<u16>> ht(nparts);
        for (size_t p = 0; p < nparts; ++p) {
            u64 n = n_mums[p], nd = ndocs[p];
            if (nd < 2 || nd > 0xffffffffull) throw Error(61, "a partition needs at least two sequences");
            if (inputs_on_device) { src[p] = MergeSource{(const u32*)lengths[p], (const i64*)offsets[p], (const u8*)strands[p], (const u16*)thresholds[p], n, (u32)nd}; continue; }
            hl[p].alloc(n, s); ho[p].alloc(n * nd, s); hs[p].alloc(n * nd, s); ht[p].alloc(anchor_len, s);
            if (n) {
                MMB_CUDA(cudaMemcpyAsync(hl[p].p, lengths[p], n * 4, cudaMemcpyHostToDevice, s));
                MMB_CUDA(cudaMemcpyAsync(ho[p].p, offsets[p], n * nd * 8, cudaMemcpyHostToDevice, s));
                MMB_CUDA(cudaMemcpyAsync(hs[p].p, strands[p], n * nd, cudaMemcpyHostToDevice, s));
            }
            if (anchor_len) MMB_CUDA(cudaMemcpyAsync(ht[p].p, thresholds[p], anchor_len * 2, cudaMemcpyHostToDevice, s));
            src[p] = MergeSource{hl[p].p, ho[p].p, hs[p].p, ht[p].p, n, (u32)nd};
        }
        merge_partitions(src.data(), nparts, anchor_len, e->merged, s);
        e->merged_anchor_len = anchor_len;
        e->merged_text.release(); e->merged_text_bytes = 0; e->has_merged_text = false;      // formatted on demand (mmb_merge_text_size / _copy_text)
        MMB_CUDA(cudaStreamSynchronize(s));
        e->has_merged = true;
    });
}

static int ensure_merged_text(mmb_engine* e) {
    if (!e || !e->has_merged) return fail(MMB_ERR_STATE, "no merge result");
    if (e->has_merged_text) return MMB_OK;
    return guarded([&] {
        MMB_CUDA(cudaSetDevice(e->device));
        merge_text(e->merged, e->merged_text, e->merged_text_bytes, e->stream);
        e->has_merged_text = true;
    });
}

uint64_t mmb_merge_num_mums(const mmb_engine* e) { return (e && e->has_merged) ? e->merged.n : 0; }
uint64_t mmb_merge_num_docs(const mmb_engine* e) { return (e && e->has_merged) ? e->merged.ndocs : 0; }
uint64_t mmb_merge_text_size(mmb_engine* e) { return ensure_merged_text(e) == MMB_OK ? e->merged_text_bytes : 0; }

int mmb_merge_copy(mmb_engine* e, uint32_t* lengths, int64_t* offsets, uint8_t* strands, uint16_t* thresholds) {
    if (!e || !e->has_merged) return fail(MMB_ERR_STATE, "no merge result");
    return guarded([&] {
        cudaStream_t s = e->stream; u64 n = e->merged.n, cells = n * e->merged.ndocs;
        copy_out(e->merged.len.p, lengths, n, s);
        copy_out(e->merged.off.p, offsets, cells, s);
        copy_out(e->merged.plus.p, strands, cells, s);
        copy_out(e->merged.nb.p, thresholds, e->merged_anchor_len, s);
        MMB_CUDA(cudaStreamSynchronize(s));
    });
}

const void* mmb_merge_text_device(mmb_engine* e) { return ensure_merged_text(e) == MMB_OK ? (const void*)e->merged_text.p : nullptr; }

int mmb_merge_copy_text(mmb_engine* e, void* host_dst) {
    if (int rc = ensure_merged_text(e)) return rc;
    return guarded([&] { copy_out(e->merged_text.p, (u8*)host_dst, e->merged_text_bytes, e->stream); MMB_CUDA(cudaStreamSynchronize(e->stream)); });
}

int mmb_profile_kernel(const char* name_substring) {
    std::lock_guard<std::mutex> lk(profimpl::mu);
    for (auto& p : profimpl::pairs) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    profimpl::pairs.clear(); profimpl::names.clear(); profimpl::open_pair = false;
    profimpl::filter = name_substring ? name_substring : "";
    profimpl::on.store(!profimpl::filter.empty());
    return MMB_OK;
}

int mmb_profile_result(double* total_ms, uint64_t* launches) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(profimpl::mu);
    double t = 0;
    for (auto& p : profimpl::pairs) { float ms = 0; if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) t += ms; }
    if (total_ms) *total_ms = t;
    if (launches) *launches = profimpl::pairs.size();
    return MMB_OK;
}

int mmb_profile_table(char* dst, size_t cap, size_t* needed) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(profimpl::mu);
    std::map<std::string, std::pair<double, uint64_t>> agg;
    for (size_t i = 0; i < profimpl::pairs.size(); ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, profimpl::pairs[i].first, profimpl::pairs[i].second) != cudaSuccess) continue;
        auto& a = agg[profimpl::names[i]]; a.first += ms; ++a.second;
    }
    std::string out;
    for (auto& kv : agg) out += kv.first + "\t" + std::to_string(kv.second.first) + "\t" + std::to_string(kv.second.second) + "\n";
    if (needed) *needed = out.size() + 1;
    if (dst && cap) { size_t k = std::min(cap - 1, out.size()); memcpy(dst, out.data(), k); dst[k] = 0; }
    return MMB_OK;
}

int mmb_copy_symbols(mmb_engine* e, uint8_t* dst) {
    if (!e || !e->loaded || !e->text.text.p) return fail(MMB_ERR_STATE, "no input loaded");
    return guarded([&] { copy_out(e->text.text.p, dst, e->text.n, e->stream); MMB_CUDA(cudaStreamSynchronize(e->stream)); });
}

}  // extern "C"
