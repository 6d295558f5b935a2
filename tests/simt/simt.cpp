// tests/simt/simt.cpp -- TEST INFRASTRUCTURE ONLY (see simt.h).  Round-robin fiber scheduler for one block at a time.
#define MMB_SIMT 1
#include "simt.h"

namespace simt {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;
unsigned long long g_launches = 0;
int g_block_order = 0;

namespace {
constexpr size_t kStack = 256 * 1024;
struct Fiber { ucontext_t ctx; bool done; uint3 tid; unsigned warp; };
struct MaskBarrier { unsigned mask; unsigned count; unsigned long gen; };
struct Warp { unsigned long long slot[32]; unsigned lanes; std::vector<MaskBarrier> bar; };
std::vector<Fiber> fibers;
std::vector<char*> stacks;
std::vector<Warp> warps;
ucontext_t sched_ctx;
int cur = -1, live = 0, bar_count = 0;
unsigned long bar_gen = 0, progress = 0;
const std::function<void()>* body = nullptr;

void yield_to_scheduler() { swapcontext(&fibers[cur].ctx, &sched_ctx); }

void fiber_main() {
    (*body)();
    fibers[cur].done = true;
    --live; ++progress;
    if (live > 0 && bar_count == live) { bar_count = 0; ++bar_gen; }   // exited threads no longer hold a barrier back
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}
}  // namespace

void block_barrier() {
    unsigned long g = bar_gen;
    if (++bar_count == live) { bar_count = 0; ++bar_gen; ++progress; return; }
    while (bar_gen == g) yield_to_scheduler();
}

// Disjoint lane sets of one warp may run collectives at the same time (divergent code), so barriers are keyed by their mask.
void warp_barrier(unsigned mask) {
    Warp& w = warps[fibers[cur].warp];
    mask &= w.lanes;
    size_t i = 0;
    while (i < w.bar.size() && w.bar[i].mask != mask) ++i;
    if (i == w.bar.size()) w.bar.push_back(MaskBarrier{mask, 0, 0});
    unsigned expected = (unsigned)__builtin_popcount(mask);
    unsigned long g = w.bar[i].gen;
    if (++w.bar[i].count >= expected) { w.bar[i].count = 0; ++w.bar[i].gen; ++progress; return; }
    while (w.bar[i].gen == g) yield_to_scheduler();      // entries are only appended within a block, so index i stays valid
}

unsigned long long* warp_slot(unsigned lane) { return &warps[fibers[cur].warp].slot[lane & 31]; }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn) {
    ++g_launches;
    size_t nthreads = (size_t)block.x * block.y * block.z;
    if (nthreads == 0 || (size_t)grid.x * grid.y * grid.z == 0) return;
    if (nthreads > 1024) { fprintf(stderr, "simt: block of %zu threads\n", nthreads); abort(); }
    while (stacks.size() < nthreads) stacks.push_back((char*)malloc(kStack));
    std::vector<unsigned char> dyn(smem + 16);
    g_dyn_smem = dyn.data();
    g_blockDim = block; g_gridDim = grid;
    body = &fn;
    fibers.resize(nthreads);
    warps.resize((nthreads + 31) / 32);
    // Blocks of a real grid run in no particular order; tests switch between forward, reverse and a scrambled order so that code
    // relying on the order of e.g. atomic appends does not pass by accident.
    size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    for (size_t bi = 0; bi < nblocks; ++bi) {
        size_t b = bi;
        if (g_block_order == 1) b = nblocks - 1 - bi;
        else if (g_block_order == 2) {                   // bijection on [0, nblocks): multiply by an odd constant modulo the next power of two, cycle-walk
            size_t p2 = 1; while (p2 < nblocks) p2 <<= 1;
            b = bi;
            do { b = (b * 0x9E3779B1ull + 12345) & (p2 - 1); } while (b >= nblocks);
        }
        unsigned bx = (unsigned)(b % grid.x), by = (unsigned)((b / grid.x) % grid.y), bz = (unsigned)(b / ((size_t)grid.x * grid.y));
        for (auto& w : warps) { w.bar.clear(); w.lanes = 0; }
        size_t t = 0;
        for (unsigned tz = 0; tz < block.z; ++tz) for (unsigned ty = 0; ty < block.y; ++ty) for (unsigned tx = 0; tx < block.x; ++tx, ++t) {
            Fiber& f = fibers[t];
            f.done = false; f.tid = uint3{tx, ty, tz}; f.warp = (unsigned)(t / 32);
            warps[f.warp].lanes |= 1u << (t & 31);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = stacks[t]; f.ctx.uc_stack.ss_size = kStack; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, fiber_main, 0);
        }
        live = (int)nthreads; bar_count = 0; bar_gen = 0;
        while (live > 0) {
            unsigned long before = progress;
            for (size_t i = 0; i < nthreads; ++i) {
                if (fibers[i].done) continue;
                cur = (int)i;
                g_threadIdx = fibers[i].tid; g_blockIdx = uint3{bx, by, bz};
                swapcontext(&sched_ctx, &fibers[i].ctx);
            }
            if (live > 0 && progress == before) { fprintf(stderr, "simt: deadlock (barrier never completed) in block %u\n", bx); abort(); }
        }
    }
    g_dyn_smem = nullptr;
}

}  // namespace simt

extern "C" __attribute__((visibility("default"))) void simt_set_block_order(int mode) { simt::g_block_order = mode; }
