// mumemto_b200/csrc/sa.cu -- suffix array of the text (plus the smallest sentinel suffix) by prefix doubling with group filtering.
//
// Replaces the reference's gsacak() call on the whole text (include/direct_gsacak.hpp:56-67; third-party gsa-is) and defines the
// same stream the PFP enumeration produces (include/pfp_lcp_mum.hpp:115-231): ranks 0..n over T followed by a sentinel smaller
// than every symbol, rank 0 being the sentinel suffix (SA = n).  Rows C/D of SURVEY.md section 8.
//
// Layout in HBM (m = n + 1 suffixes, 32-bit indices, direct mode requires m < 2^32):
//   sa[m]   current order            isa[m]  rank of the first member of the suffix's group ("head rank")
//   round-local: pos[A] (ranks still in multi-member groups), key[A] = head<<32 | rank(suffix+h)+1, val[A] = suffix
// Round h: gather keys for the A active ranks (2 random 4 B reads each), radix sort (key,val), write sa back through pos, flag new
// group starts, max-scan them into head ranks, update isa, drop singletons.  Stops when no multi-member group is left.
// Initial order: the first `cpk` symbols packed `bits` per symbol into one 64-bit key (21 symbols for DNA + '$' + N).
#include "sa.h"
#include "prims.h"

namespace mmb {

namespace {

constexpr u32 TPB = 256;

__global__ void __launch_bounds__(TPB) symbol_histogram(const u8* __restrict__ text, u64 n, u32* __restrict__ hist) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&h[text[i]], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(TPB) pack_keys(const u8* __restrict__ text, u64 n, const u8* __restrict__ lut, u32 bits, u32 cpk,
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

// After a sort of all m suffixes: group starts, and which ranks sit in multi-member groups.
__global__ void __launch_bounds__(TPB) first_flags(const u64* __restrict__ keys, u32 m, u32* __restrict__ head_cand, u32* __restrict__ active) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    u64 k = keys[p];
    bool start = (p == 0) || keys[p - 1] != k;
    bool next_start = (p + 1 == m) || keys[p + 1] != k;
    head_cand[p] = start ? p : 0u;
    active[p] = (start && next_start) ? 0u : 1u;
}

__global__ void __launch_bounds__(TPB) first_ranks(const u32* __restrict__ sa, const u32* __restrict__ head, u32 m, u32* __restrict__ isa) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) isa[sa[p]] = head[p];
}

__global__ void __launch_bounds__(TPB) compact_first(const u32* __restrict__ active, const u32* __restrict__ idx, u32 m, u32* __restrict__ pos) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m && active[p]) pos[idx[p]] = p;
}

__global__ void __launch_bounds__(TPB) gather_keys(const u32* __restrict__ pos, const u32* __restrict__ sa, const u32* __restrict__ isa, u32 h, u32 m,
                                                   u32 A, u64* __restrict__ keys, u32* __restrict__ vals) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    u32 s = sa[pos[a]];
    u64 t = (u64)s + h;
    u32 r2 = (t < m) ? isa[t] + 1u : 0u;
    keys[a] = ((u64)isa[s] << 32) | r2;
    vals[a] = s;
}

__global__ void __launch_bounds__(TPB) scatter_sorted(const u32* __restrict__ pos, const u64* __restrict__ keys, const u32* __restrict__ vals, u32 A,
                                                      u32* __restrict__ sa, u32* __restrict__ head_cand, u32* __restrict__ still) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    u64 k = keys[a];
    bool start = (a == 0) || keys[a - 1] != k;
    bool next_start = (a + 1 == A) || keys[a + 1] != k;
    u32 p = pos[a];
    sa[p] = vals[a];
    head_cand[a] = start ? p : 0u;
    still[a] = (start && next_start) ? 0u : 1u;
}

__global__ void __launch_bounds__(TPB) update_ranks(const u32* __restrict__ vals, const u32* __restrict__ head, u32 A, u32* __restrict__ isa) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < A) isa[vals[a]] = head[a];
}

__global__ void __launch_bounds__(TPB) compact_active(const u32* __restrict__ pos, const u32* __restrict__ still, const u32* __restrict__ idx, u32 A,
                                                      u32* __restrict__ pos_out) {
    u32 a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < A && still[a]) pos_out[idx[a]] = pos[a];
}

u32 bits_for(u64 x) { u32 b = 1; while ((1ull << b) <= x && b < 63) ++b; return b; }   // bits to represent values 0..x

u32 read_count(const u32* idx, const u32* flag, u32 count, cudaStream_t s) {
    u32 h[2] = {0, 0};
    if (count == 0) return 0;
    MMB_CUDA(cudaMemcpyAsync(&h[0], idx + count - 1, 4, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaMemcpyAsync(&h[1], flag + count - 1, 4, cudaMemcpyDeviceToHost, s));
    MMB_CUDA(cudaStreamSynchronize(s));
    return h[0] + h[1];
}

}  // namespace

void build_suffix_array(const u8* text, u64 n, SuffixArray& out, cudaStream_t s) {
    if (n + 1 >= 0xfffffff0ull) throw Error(20, "text too long for the direct (32-bit) suffix array builder");
    u32 m = (u32)(n + 1);
    out.m = m;
    out.sa.alloc(m, s); out.isa.alloc(m, s);

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

    DevBuf<u64> key_a, key_b; DevBuf<u32> val_a, val_b, pos_a, pos_b, cand, head, flag, idx;
    key_a.alloc(m, s); key_b.alloc(m, s); val_a.alloc(m, s); val_b.alloc(m, s);
    cand.alloc(m, s); head.alloc(m, s); flag.alloc(m, s); idx.alloc(m, s);

    MMB_LAUNCH(pack_keys, ceil_div(m, TPB), TPB, 0, s, text, n, d_lut.p, bits, cpk, key_a.p, val_a.p);
    prim::sort_pairs(key_a.p, key_b.p, val_a.p, out.sa.p, m, 0, (int)(bits * cpk), s);
    MMB_LAUNCH(first_flags, ceil_div(m, TPB), TPB, 0, s, key_b.p, m, cand.p, flag.p);
    prim::inclusive_max(cand.p, head.p, m, s);
    MMB_LAUNCH(first_ranks, ceil_div(m, TPB), TPB, 0, s, out.sa.p, head.p, m, out.isa.p);
    prim::exclusive_sum(flag.p, idx.p, m, s);
    u32 A = read_count(idx.p, flag.p, m, s);
    pos_a.alloc(A, s); pos_b.alloc(A, s);
    if (A) MMB_LAUNCH(compact_first, ceil_div(m, TPB), TPB, 0, s, flag.p, idx.p, m, pos_a.p);

    int sort_bits = 32 + (int)bits_for(m);
    out.rounds = 0; out.sorted_elements = m;
    for (u64 h = cpk; A > 0; h *= 2) {
        if (h >= 2 * (u64)m) throw Error(22, "internal: prefix doubling did not converge");
        u32 g = ceil_div(A, TPB);
        MMB_LAUNCH(gather_keys, g, TPB, 0, s, pos_a.p, out.sa.p, out.isa.p, (u32)std::min<u64>(h, m), m, A, key_a.p, val_a.p);
        prim::sort_pairs(key_a.p, key_b.p, val_a.p, val_b.p, A, 0, sort_bits, s);
        MMB_LAUNCH(scatter_sorted, g, TPB, 0, s, pos_a.p, key_b.p, val_b.p, A, out.sa.p, cand.p, flag.p);
        prim::inclusive_max(cand.p, head.p, A, s);
        MMB_LAUNCH(update_ranks, g, TPB, 0, s, val_b.p, head.p, A, out.isa.p);
        prim::exclusive_sum(flag.p, idx.p, A, s);
        u32 A2 = read_count(idx.p, flag.p, A, s);
        if (A2) MMB_LAUNCH(compact_active, g, TPB, 0, s, pos_a.p, flag.p, idx.p, A, pos_b.p);
        pos_a.swap(pos_b);
        out.sorted_elements += A;
        A = A2; ++out.rounds;
    }
}

}  // namespace mmb
