// mumemto_b200/csrc/fasta_io.h -- host side of the drop-in executables: FASTA files on disk -> device memory.
//
// Stands in for the gzopen / kseq_read loop of RefBuilder::build_input_file (reference src/ref_builder.cpp:211-268) as far as the
// bytes are concerned (gzopen reads plain and gzip files alike); record parsing itself happens on the device (text.cu).
// A pool of threads reads 32 MiB slices (pread from the page cache, or a slice of an inflated gzip image) into per-thread pinned
// buffers and copies them to the device on per-thread streams, so disk reads, page-cache copies and PCIe transfers overlap and no
// multi-GB pinned allocation is needed.  Host C++ only (CUDA runtime API).
#pragma once
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace fio {

struct Source {
    std::string path;
    bool gz = false;
    uint64_t size = 0;              // decompressed bytes
    std::string inflated;           // gzip files only
};

inline std::string inflate_file(const std::string& path) {
    gzFile f = gzopen(path.c_str(), "r");
    if (!f) throw std::runtime_error("cannot open " + path);
    gzbuffer(f, 1 << 20);
    std::string out; std::vector<char> buf(1 << 22); int k;
    while ((k = gzread(f, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), k);
    gzclose(f);
    return out;
}

template <class Job>
inline void parallel_for(size_t n, unsigned workers, Job&& job) {
    std::atomic<size_t> next{0};
    std::vector<std::string> errors(workers);
    std::vector<std::thread> th;
    for (unsigned w = 0; w < workers; ++w)
        th.emplace_back([&, w] {
            try { for (size_t i; (i = next++) < n;) job(i, w); } catch (const std::exception& ex) { errors[w] = ex.what(); next = n; }
        });
    for (auto& t : th) t.join();
    for (auto& e : errors) if (!e.empty()) throw std::runtime_error(e);
}

inline unsigned pool_size(size_t items) {
    unsigned hw = std::thread::hardware_concurrency();
    return (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<unsigned>(hw ? hw : 4, 16), items));
}

// sizes of plain files, inflated images of gzip files
inline std::vector<Source> probe(const std::vector<std::string>& files) {
    std::vector<Source> src(files.size());
    parallel_for(files.size(), pool_size(files.size()), [&](size_t i, unsigned) {
        Source& s = src[i]; s.path = files[i];
        unsigned char h[2] = {0, 0};
        int fd = open(s.path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open " + s.path);
        ssize_t k = read(fd, h, 2);
        struct stat st; int rc = fstat(fd, &st); close(fd);
        if (rc != 0) throw std::runtime_error("cannot stat " + s.path);
        s.gz = k == 2 && h[0] == 0x1f && h[1] == 0x8b;
        if (s.gz) { s.inflated = inflate_file(s.path); s.size = s.inflated.size(); } else s.size = (uint64_t)st.st_size;
    });
    return src;
}

struct DeviceImages {
    uint8_t* d_raw = nullptr;                   // cudaMalloc'ed on `device`; images at 16-byte aligned offsets
    std::vector<uint64_t> offsets, sizes;
    int device = 0;
    void release() { if (d_raw) { cudaSetDevice(device); cudaFree(d_raw); d_raw = nullptr; } }
};

inline void cu(cudaError_t e, const char* what) { if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e) + " (" + what + ")"); }

inline DeviceImages upload(const std::vector<const Source*>& files, int device) {
    constexpr uint64_t SLICE = 32ull << 20;
    DeviceImages out; out.device = device;
    uint64_t total = 0;
    struct Item { size_t file; uint64_t at, bytes; };
    std::vector<Item> items;
    for (size_t i = 0; i < files.size(); ++i) {
        out.offsets.push_back(total); out.sizes.push_back(files[i]->size);
        for (uint64_t at = 0; at < files[i]->size; at += SLICE) items.push_back(Item{i, at, std::min(SLICE, files[i]->size - at)});
        total += (files[i]->size + 15) & ~15ull;
    }
    cu(cudaSetDevice(device), "cudaSetDevice");
    cu(cudaMalloc((void**)&out.d_raw, total + 64), "cudaMalloc input images");
    cu(cudaMemset(out.d_raw + (total & ~15ull), 0, 64 + (total & 15)), "cudaMemset");
    unsigned workers = pool_size(items.size());
    std::vector<int> fds(files.size(), -1);
    for (size_t i = 0; i < files.size(); ++i)
        if (!files[i]->gz) { fds[i] = open(files[i]->path.c_str(), O_RDONLY); if (fds[i] < 0) throw std::runtime_error("cannot open " + files[i]->path); }
    std::vector<uint8_t*> stage(workers, nullptr); std::vector<cudaStream_t> st(workers, nullptr);
    try {
        parallel_for(items.size(), workers, [&](size_t k, unsigned w) {
            if (!stage[w]) {
                cu(cudaSetDevice(device), "cudaSetDevice");
                cu(cudaHostAlloc((void**)&stage[w], SLICE, cudaHostAllocDefault), "cudaHostAlloc staging");
                cu(cudaStreamCreateWithFlags(&st[w], cudaStreamNonBlocking), "cudaStreamCreate");
            }
            const Item& it = items[k]; const Source& s = *files[it.file];
            if (s.gz) memcpy(stage[w], s.inflated.data() + it.at, it.bytes);
            else {
                uint64_t got = 0;
                while (got < it.bytes) {
                    ssize_t r = pread(fds[it.file], stage[w] + got, it.bytes - got, (off_t)(it.at + got));
                    if (r <= 0) throw std::runtime_error("short read on " + s.path);
                    got += (uint64_t)r;
                }
            }
            cu(cudaMemcpyAsync(out.d_raw + out.offsets[it.file] + it.at, stage[w], it.bytes, cudaMemcpyHostToDevice, st[w]), "H2D");
            cu(cudaStreamSynchronize(st[w]), "H2D");
        });
    } catch (...) {
        for (unsigned w = 0; w < workers; ++w) { if (stage[w]) cudaFreeHost(stage[w]); if (st[w]) cudaStreamDestroy(st[w]); }
        for (int fd : fds) if (fd >= 0) close(fd);
        out.release();
        throw;
    }
    for (unsigned w = 0; w < workers; ++w) { if (stage[w]) cudaFreeHost(stage[w]); if (st[w]) cudaStreamDestroy(st[w]); }
    for (int fd : fds) if (fd >= 0) close(fd);
    return out;
}

// device -> pinned slices -> write(2)
inline void write_device_buffer(const std::string& path, const void* d_src, uint64_t bytes) {
    constexpr uint64_t SLICE = 64ull << 20;
    int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) throw std::runtime_error("cannot write " + path);
    uint8_t* stage[2] = {nullptr, nullptr}; cudaStream_t st;
    uint64_t slice = std::min<uint64_t>(SLICE, std::max<uint64_t>(bytes, 1));
    cu(cudaHostAlloc((void**)&stage[0], slice, cudaHostAllocDefault), "cudaHostAlloc"); cu(cudaHostAlloc((void**)&stage[1], slice, cudaHostAllocDefault), "cudaHostAlloc");
    cu(cudaStreamCreate(&st), "cudaStreamCreate");
    uint64_t at = 0; int cur = 0;
    uint64_t n0 = std::min(slice, bytes);
    if (n0) cu(cudaMemcpyAsync(stage[0], d_src, n0, cudaMemcpyDeviceToHost, st), "D2H");
    while (at < bytes) {
        uint64_t n = std::min(slice, bytes - at);
        cu(cudaStreamSynchronize(st), "D2H");
        uint64_t nxt = at + n;
        if (nxt < bytes) cu(cudaMemcpyAsync(stage[cur ^ 1], (const uint8_t*)d_src + nxt, std::min(slice, bytes - nxt), cudaMemcpyDeviceToHost, st), "D2H");
        const uint8_t* q = stage[cur]; uint64_t left = n;
        while (left) { ssize_t k = write(fd, q, left); if (k <= 0) { close(fd); throw std::runtime_error("short write on " + path); } q += k; left -= (uint64_t)k; }
        at = nxt; cur ^= 1;
    }
    close(fd);
    cudaFreeHost(stage[0]); cudaFreeHost(stage[1]); cudaStreamDestroy(st);
}

}  // namespace fio
