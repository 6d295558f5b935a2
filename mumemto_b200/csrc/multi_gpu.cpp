// mumemto_b200/csrc/multi_gpu.cpp -- see multi_gpu.h.  Compiled into bin/mumemto_exec only (the library itself does not depend on NCCL).
#include "multi_gpu.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <thread>

#include "mumemto_b200.h"

namespace mgpu {

namespace {

struct Fail : std::runtime_error { using std::runtime_error::runtime_error; };
void cu(cudaError_t e, const char* what) { if (e != cudaSuccess) throw Fail(std::string("CUDA: ") + cudaGetErrorString(e) + " (" + what + ")"); }
void nc(ncclResult_t r, const char* what) { if (r != ncclSuccess) throw Fail(std::string("NCCL: ") + ncclGetErrorString(r) + " (" + what + ")"); }
void mm(int rc) { if (rc != 0) throw Fail(mmb_last_error()); }

// byte layout of one partition's payload: lengths u32[n] | offsets i64[n][nd] | strands u8[n][nd] | thresholds u16[L], padded to 16
struct Layout { uint64_t off, str, thr, bytes; };
Layout layout(uint64_t n, uint64_t nd, uint64_t L) {
    auto a8 = [](uint64_t x) { return (x + 7) & ~7ull; };
    Layout l; l.off = a8(4 * n); l.str = l.off + 8 * n * nd; l.thr = a8(l.str + n * nd); l.bytes = (l.thr + 2 * L + 15) & ~15ull;
    return l;
}

struct Part {
    std::vector<size_t> files;
    mmb_engine* e = nullptr;
    int device = 0;
    uint64_t n = 0, nd = 0, L = 0;
    const void *len = nullptr, *off = nullptr, *str = nullptr, *thr = nullptr;
    std::string error;
    std::string lengths_lines;      // this partition's share of PREFIX.lengths
};

void put(const std::string& path, const void* p, size_t n) {
    std::ofstream o(path, std::ios::binary);
    if (!o) throw Fail("cannot write " + path);
    o.write((const char*)p, (std::streamsize)n);
}

}  // namespace

std::vector<std::vector<size_t>> make_partitions(size_t nfiles, size_t parts) {
    std::vector<std::vector<size_t>> out;
    size_t rest = nfiles - 1;
    parts = std::max<size_t>(1, std::min(parts, rest));
    size_t at = 1;
    for (size_t p = 0; p < parts; ++p) {
        size_t take = rest / parts + (p < rest % parts ? 1 : 0);
        std::vector<size_t> g{0};
        for (size_t k = 0; k < take; ++k) g.push_back(at++);
        out.push_back(std::move(g));
    }
    return out;
}

bool run(const Request& rq, std::string& error) {
    try {
        int ndev = 0;
        cu(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount");
        const std::vector<fio::Source>& src = *rq.sources;
        if (src.size() < 3) throw Fail("partition-merge needs an anchor and at least two other sequences");
        if (rq.gpus < 1 || rq.first_device < 0 || rq.first_device + rq.gpus > ndev)
            throw Fail("--gpus " + std::to_string(rq.gpus) + " from device " + std::to_string(rq.first_device) + ": only " + std::to_string(ndev) + " CUDA devices");
        if (rq.gpus > 1 && rq.partitions > 0 && rq.partitions != rq.gpus) throw Fail("--partitions must equal --gpus when several GPUs are used (one partition per GPU)");
        size_t want = rq.gpus > 1 ? (size_t)rq.gpus : (size_t)std::max(2, rq.partitions);
        auto groups = make_partitions(src.size(), want);
        if (groups.size() < 2) throw Fail("not enough sequences for two partitions");
        const size_t P = groups.size();
        const int G = (int)std::min<size_t>((size_t)rq.gpus, P);
        std::vector<Part> parts(P);
        for (size_t p = 0; p < P; ++p) { parts[p].files = groups[p]; parts[p].device = rq.first_device + (G > 1 ? (int)p : 0); }

        mmb_options mo; mmb_default_options(&mo);
        mo.min_match_len = rq.min_match_len; mo.merge = 1; mo.anchor_merge = 1; mo.cli_normalise = 1; mo.want_text = 0;

        // ---- every partition through the hot path: one host thread per GPU (or one after the other on a single GPU)
        auto work = [&](Part& pt) {
            try {
                cu(cudaSetDevice(pt.device), "cudaSetDevice");
                mm(mmb_create(pt.device, nullptr, &pt.e));
                std::vector<const fio::Source*> mine;
                for (size_t f : pt.files) mine.push_back(&src[f]);
                fio::DeviceImages img = fio::upload(mine, pt.device);
                int rc = mmb_load_fasta_device(pt.e, img.d_raw, img.offsets.data(), img.sizes.data(), img.sizes.size(), rq.use_rcomp ? 1 : 0);
                img.release();
                mm(rc);
                mm(mmb_find(pt.e, &mo));
                pt.n = mmb_num_matches(pt.e); pt.nd = mmb_num_docs(pt.e); pt.L = mmb_threshold_count(pt.e) / 2;
                mm(mmb_mum_table_device(pt.e, &pt.len, &pt.off, &pt.str, &pt.thr));
                // PREFIX.lengths lines (src/ref_builder.cpp:193-209); the anchor is listed by the first partition only
                for (size_t d = (pt.files[0] == 0 && &pt != &parts[0]) ? 1 : 0; d < pt.files.size(); ++d) {
                    std::string cp = std::filesystem::canonical(src[pt.files[d]].path).string();
                    uint64_t total = 0; size_t nr = mmb_num_records(pt.e, d);
                    for (size_t r = 0; r < nr; ++r) { uint64_t len; mmb_record(pt.e, d, r, nullptr, &len); total += len; }
                    pt.lengths_lines += cp + " * " + std::to_string(total) + "\n";
                    for (size_t r = 0; r < nr; ++r) { const char* name; uint64_t len; mmb_record(pt.e, d, r, &name, &len); pt.lengths_lines += cp + " " + name + " " + std::to_string(len) + "\n"; }
                }
            } catch (const std::exception& ex) { pt.error = ex.what(); }
        };
        // payloads: packed per partition on its own device
        std::vector<uint8_t*> send(P, nullptr);
        auto pack = [&](Part& pt, uint8_t* dst) {
            Layout l = layout(pt.n, pt.nd, pt.L);
            cu(cudaMemsetAsync(dst, 0, l.bytes, 0), "memset payload");
            if (pt.n) {
                cu(cudaMemcpyAsync(dst, pt.len, 4 * pt.n, cudaMemcpyDeviceToDevice, 0), "pack lengths");
                cu(cudaMemcpyAsync(dst + l.off, pt.off, 8 * pt.n * pt.nd, cudaMemcpyDeviceToDevice, 0), "pack offsets");
                cu(cudaMemcpyAsync(dst + l.str, pt.str, pt.n * pt.nd, cudaMemcpyDeviceToDevice, 0), "pack strands");
            }
            if (pt.L) cu(cudaMemcpyAsync(dst + l.thr, pt.thr, 2 * pt.L, cudaMemcpyDeviceToDevice, 0), "pack thresholds");
        };

        uint8_t* gathered = nullptr;            // on the first device: P payloads of `stride` bytes
        uint64_t stride = 0;
        if (G > 1) {
            std::vector<std::thread> th;
            for (auto& pt : parts) th.emplace_back(work, std::ref(pt));
            for (auto& t : th) t.join();
            for (auto& pt : parts) if (!pt.error.empty()) throw Fail(pt.error);
            for (auto& pt : parts) {
                if (pt.L != parts[0].L) throw Fail("partitions do not share their first sequence");
                stride = std::max(stride, layout(pt.n, pt.nd, pt.L).bytes);
            }
            // ---- ONE all-gather of the padded payloads
            std::vector<int> devs(G); for (int g = 0; g < G; ++g) devs[g] = rq.first_device + g;
            std::vector<ncclComm_t> comm(G);
            nc(ncclCommInitAll(comm.data(), G, devs.data()), "ncclCommInitAll");
            std::vector<uint8_t*> recv(G, nullptr); std::vector<cudaStream_t> st(G);
            for (int g = 0; g < G; ++g) {
                cu(cudaSetDevice(devs[g]), "cudaSetDevice");
                cu(cudaStreamCreate(&st[g]), "cudaStreamCreate");
                cu(cudaMalloc((void**)&send[g], stride), "cudaMalloc send");
                cu(cudaMalloc((void**)&recv[g], stride * G), "cudaMalloc recv");
                cu(cudaMemsetAsync(send[g], 0, stride, 0), "memset");
                pack(parts[g], send[g]);
                cu(cudaDeviceSynchronize(), "pack");
            }
            auto t0 = std::chrono::steady_clock::now();
            nc(ncclGroupStart(), "ncclGroupStart");
            for (int g = 0; g < G; ++g) nc(ncclAllGather(send[g], recv[g], stride, ncclUint8, comm[g], st[g]), "ncclAllGather");
            nc(ncclGroupEnd(), "ncclGroupEnd");
            for (int g = 0; g < G; ++g) { cu(cudaSetDevice(devs[g]), "cudaSetDevice"); cu(cudaStreamSynchronize(st[g]), "all-gather"); }
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            std::fprintf(stderr, "\n\tall-gather of %d x %.1f MB over NCCL: %.2f ms", G, stride / 1e6, ms);
            gathered = recv[0];
            for (int g = 0; g < G; ++g) {
                cu(cudaSetDevice(devs[g]), "cudaSetDevice");
                cu(cudaFree(send[g]), "cudaFree"); if (g) cu(cudaFree(recv[g]), "cudaFree");
                cu(cudaStreamDestroy(st[g]), "cudaStreamDestroy");
                ncclCommDestroy(comm[g]);
            }
            cu(cudaSetDevice(devs[0]), "cudaSetDevice");
        } else {
            // one GPU: partitions one after the other; each payload is packed into the gathered buffer before the next run reuses the engine's memory
            std::vector<std::vector<uint8_t>> unused;
            std::vector<uint8_t*> keep(P, nullptr);
            for (size_t p = 0; p < P; ++p) {
                work(parts[p]);
                if (!parts[p].error.empty()) throw Fail(parts[p].error);
                if (parts[p].L != parts[0].L) throw Fail("partitions do not share their first sequence");
                Layout l = layout(parts[p].n, parts[p].nd, parts[p].L);
                cu(cudaMalloc((void**)&keep[p], l.bytes), "cudaMalloc payload");
                pack(parts[p], keep[p]);
                cu(cudaDeviceSynchronize(), "pack");
                stride = std::max(stride, l.bytes);
                if (p + 1 < P) { mmb_destroy(parts[p].e); parts[p].e = nullptr; }       // the last engine stays for the merge
            }
            cu(cudaMalloc((void**)&gathered, stride * P), "cudaMalloc gathered");
            for (size_t p = 0; p < P; ++p) {
                cu(cudaMemcpy(gathered + p * stride, keep[p], layout(parts[p].n, parts[p].nd, parts[p].L).bytes, cudaMemcpyDeviceToDevice), "gather");
                cu(cudaFree(keep[p]), "cudaFree");
            }
        }

        // ---- merge on the first device (left fold in partition order, like `anchor_merge p0 p1 ...`)
        mmb_engine* e0 = G > 1 ? parts[0].e : parts[P - 1].e;
        std::vector<const void*> L(P), O(P), S(P), T(P); std::vector<uint64_t> n(P), nd(P);
        for (size_t p = 0; p < P; ++p) {
            Layout l = layout(parts[p].n, parts[p].nd, parts[p].L);
            const uint8_t* base = gathered + p * stride;
            L[p] = base; O[p] = base + l.off; S[p] = base + l.str; T[p] = base + l.thr; n[p] = parts[p].n; nd[p] = parts[p].nd;
        }
        uint64_t anchor_len = parts[0].L;
        mm(mmb_anchor_merge(e0, P, L.data(), n.data(), O.data(), S.data(), T.data(), nd.data(), anchor_len, 1));
        uint64_t nm = mmb_merge_num_mums(e0), ns = mmb_merge_num_docs(e0);
        std::vector<uint16_t> nb(anchor_len);
        if (rq.binary) {
            if (nm == 0) throw Fail("no merged MUMs: nothing to write in binary format");
            std::vector<uint32_t> len(nm); std::vector<int64_t> off(nm * ns); std::vector<uint8_t> plus(nm * ns);
            mm(mmb_merge_copy(e0, len.data(), off.data(), plus.data(), nb.data()));
            std::vector<uint8_t> bits((nm * ns + 7) / 8, 0);
            for (uint64_t i = 0; i < nm * ns; ++i) if (plus[i]) bits[i >> 3] |= (uint8_t)(0x80u >> (i & 7));
            uint16_t flags = 0x8000;
            std::ofstream b(rq.output_prefix + ".bumbl", std::ios::binary);
            b.write((char*)&flags, 2); b.write((char*)&ns, 8); b.write((char*)&nm, 8);
            b.write((char*)len.data(), nm * 4); b.write((char*)off.data(), nm * ns * 8); b.write((char*)bits.data(), bits.size());
        } else {
            size_t tn = mmb_merge_text_size(e0);
            fio::write_device_buffer(rq.output_prefix + ".mums", mmb_merge_text_device(e0), tn);
            mm(mmb_merge_copy(e0, nullptr, nullptr, nullptr, nb.data()));
        }
        put(rq.output_prefix + ".athresh", nb.data(), nb.size() * 2);
        {   // PREFIX.lengths of the whole collection: the anchor once, then every partition's own sequences
            std::ofstream Lf(rq.output_prefix + ".lengths");
            for (auto& pt : parts) Lf << pt.lengths_lines;
        }
        std::fprintf(stderr, "\n\t%zu partitions, %llu merged multi-MUMs over %llu sequences", P, (unsigned long long)nm, (unsigned long long)ns);
        cu(cudaFree(gathered), "cudaFree");
        for (auto& pt : parts) if (pt.e) mmb_destroy(pt.e);
        return true;
    } catch (const std::exception& ex) {
        error = ex.what();
        return false;
    }
}

bool run_sharded(const ShardRequest& rq, std::string& error) {
    try {
        int ndev = 0;
        cu(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount");
        const std::vector<fio::Source>& src = *rq.sources;
        if (rq.shards < 2 || rq.shards > 255) throw Fail("--shards must be between 2 and 255");
        if (rq.gpus < 1 || rq.first_device < 0 || rq.first_device + rq.gpus > ndev)
            throw Fail("--gpus " + std::to_string(rq.gpus) + " from device " + std::to_string(rq.first_device) + ": only " + std::to_string(ndev) + " CUDA devices");
        if (rq.gpus > 1 && rq.gpus != rq.shards) throw Fail("--shards must equal --gpus when several GPUs are used (one slice per GPU)");
        const bool mum_mode = rq.rare_freq == 1;
        const int K = rq.shards;
        struct Slice {
            int device = 0; std::string error; std::vector<char> text; uint64_t n = 0, nd = 0;
            std::vector<uint32_t> len; std::vector<int64_t> off; std::vector<uint8_t> plus; std::string lengths_lines; float ms = 0; uint64_t streamed = 0;
        };
        std::vector<Slice> sl(K);
        for (int k = 0; k < K; ++k) sl[k].device = rq.first_device + (rq.gpus > 1 ? k : 0);
        auto work = [&](int k) {
            Slice& me = sl[k];
            mmb_engine* e = nullptr;
            try {
                cu(cudaSetDevice(me.device), "cudaSetDevice");
                mm(mmb_create(me.device, nullptr, &e));
                std::vector<const fio::Source*> all;
                for (auto& f : src) all.push_back(&f);
                fio::DeviceImages img = fio::upload(all, me.device);
                int rc = mmb_load_fasta_device(e, img.d_raw, img.offsets.data(), img.sizes.data(), img.sizes.size(), rq.use_rcomp ? 1 : 0);
                img.release();
                mm(rc);
                mmb_options mo; mmb_default_options(&mo);
                mo.min_match_len = rq.min_match_len; mo.num_distinct_docs = rq.num_distinct_docs; mo.rare_freq = rq.rare_freq; mo.max_mem_freq = rq.max_mem_freq;
                mo.cli_normalise = 1; mo.want_text = (mum_mode && rq.binary) ? 0 : 1; mo.trim_min_len = mum_mode ? rq.trim_min_len : 0;
                mo.shard_count = (uint8_t)K; mo.shard_index = (uint8_t)k;
                mm(mmb_find(e, &mo));
                mmb_stats st; mmb_get_stats(e, &st);
                me.ms = st.ms_total; me.streamed = st.stream_suffixes;
                me.n = mmb_num_matches(e); me.nd = mmb_num_docs(e);
                if (mo.want_text) { me.text.resize(mmb_text_size(e)); if (!me.text.empty()) mm(mmb_copy_text(e, me.text.data())); }
                else if (me.n) {
                    me.len.resize(me.n); me.off.resize(me.n * me.nd); me.plus.resize(me.n * me.nd);
                    mm(mmb_copy_mums(e, me.len.data(), me.off.data(), me.plus.data()));
                }
                if (k == 0)
                    for (size_t d = 0; d < src.size(); ++d) {
                        std::string cp = std::filesystem::canonical(src[d].path).string();
                        uint64_t total = 0; size_t nr = mmb_num_records(e, d);
                        for (size_t r = 0; r < nr; ++r) { uint64_t len; mmb_record(e, d, r, nullptr, &len); total += len; }
                        me.lengths_lines += cp + " * " + std::to_string(total) + "\n";
                        for (size_t r = 0; r < nr; ++r) { const char* name; uint64_t len; mmb_record(e, d, r, &name, &len); me.lengths_lines += cp + " " + name + " " + std::to_string(len) + "\n"; }
                    }
            } catch (const std::exception& ex) { me.error = ex.what(); }
            if (e) mmb_destroy(e);
        };
        if (rq.gpus > 1) {
            std::vector<std::thread> th;
            for (int k = 0; k < K; ++k) th.emplace_back(work, k);
            for (auto& t : th) t.join();
        } else for (int k = 0; k < K; ++k) work(k);
        for (auto& s : sl) if (!s.error.empty()) throw Fail(s.error);

        uint64_t total = 0;
        for (auto& s : sl) total += s.n;
        if (mum_mode && rq.binary) {
            if (total == 0) throw Fail("no multi-MUMs found: nothing to write in binary format");
            uint64_t ns = sl[0].nd, nm = total;
            long n = (long)ns, kk = rq.num_distinct_docs;
            kk = kk < -n ? 2 : (kk <= 0 ? n + kk : (kk == 1 ? 2 : std::min(kk, n)));
            uint16_t flags = (uint16_t)(0x8000 | (kk < n ? 0x2000 : 0));
            std::vector<uint8_t> bits((nm * ns + 7) / 8, 0);
            uint64_t at = 0;
            for (auto& s : sl) for (uint64_t i = 0; i < s.n * ns; ++i, ++at) if (s.plus[i]) bits[at >> 3] |= (uint8_t)(0x80u >> (at & 7));
            std::ofstream b(rq.output_prefix + ".bumbl", std::ios::binary);
            b.write((char*)&flags, 2); b.write((char*)&ns, 8); b.write((char*)&nm, 8);
            for (auto& s : sl) b.write((char*)s.len.data(), (std::streamsize)(s.n * 4));
            for (auto& s : sl) b.write((char*)s.off.data(), (std::streamsize)(s.n * ns * 8));
            b.write((char*)bits.data(), (std::streamsize)bits.size());
        } else {
            std::ofstream o(rq.output_prefix + (mum_mode ? ".mums" : ".mems"), std::ios::binary);
            if (!o) throw Fail("cannot write " + rq.output_prefix);
            for (auto& s : sl) o.write(s.text.data(), (std::streamsize)s.text.size());
        }
        { std::ofstream Lf(rq.output_prefix + ".lengths"); Lf << sl[0].lengths_lines; }
        std::fprintf(stderr, "\n\t%d slices of the suffix order:", K);
        for (int k = 0; k < K; ++k) std::fprintf(stderr, " [%d] %llu matches, %llu ranks streamed, %.1f ms;", k, (unsigned long long)sl[k].n, (unsigned long long)sl[k].streamed, sl[k].ms);
        std::fprintf(stderr, "\n\tFound %llu matches!", (unsigned long long)total);
        return true;
    } catch (const std::exception& ex) {
        error = ex.what();
        return false;
    }
}

}  // namespace mgpu
