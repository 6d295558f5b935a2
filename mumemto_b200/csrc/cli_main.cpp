// mumemto_b200/csrc/cli_main.cpp -- host C++ command line with the reference's flag table, on top of the C ABI.
//
// Mirrors src/pfp_mum.cpp of the reference: option table and parse_build_options (:255-313), BuildOptions::validate
// (include/pfp_mum.hpp:80-147), the RefBuilder input checks (src/ref_builder.cpp:52-138: file-list, extensions, de-duplication,
// at least two inputs), the output files of mem_finder (include/mem_finder.hpp:91-158,451-503) and PREFIX.lengths
// (src/ref_builder.cpp:193-209).  Everything between reading the files and writing the results runs on the GPU through
// include/mumemto_b200.h.  -a PREFIX (detection over PREFIX.sa/.lcp/.bwt + PREFIX.lengths, src/pfp_mum.cpp:96-110) is supported; the
// .dict/.parse stage files of the CPU algorithm (-p/-P/-K) have no counterpart here.
#include <fcntl.h>
#include <getopt.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "mumemto_b200.h"
#include "fasta_io.h"
#include "multi_gpu.h"

namespace fs = std::filesystem;

#define MUMEMTO_B200_VERSION "1.4.0-b200"

[[noreturn]] static void fatal(const std::string& msg) {
    std::fprintf(stderr, "\n\033[31mError:\033[m %s\n\n", msg.c_str());
    std::exit(1);
}
static void status(const char* what) { std::fprintf(stderr, "\033[32m[build_main] \033[m%s", what); }

struct Options {
    std::string input_list, output_prefix = "output", arrays_in;
    bool use_gsacak = false;
    std::vector<std::string> files;
    bool use_rcomp = true, arrays_out = false, binary = false, merge = false, anchor_merge = false;
    int num_distinct_docs = 0, max_mem_freq = 0, rare_freq = 1;
    size_t min_match_len = 20;
    int device = 0;
    int gpus = 1;           // --gpus N: partition-merge over N GPUs in this process (anchor = first sequence), NCCL all-gather + device merge
    int partitions = 0;     // --partitions P (with one GPU): the same workflow with the partitions run one after the other
    int shards = 0;         // --shards K: ONE collection, the suffix order in K slices (one per GPU with --gpus K; one after the other on one GPU)
    int trim = 0;           // --trim[=MINLEN]: post-process multi-MUMs like `mumemto trim` (mumemto/trim_mums.py); 0 = off
};

static bool ends_with(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && !s.compare(s.size() - suf.size(), suf.size(), suf); }
static bool is_fasta_name(const std::string& f) {
    for (const char* e : {".fa", ".fasta", ".fna", ".fa.gz", ".fasta.gz", ".fna.gz"}) if (ends_with(f, e)) return true;
    return false;
}

static int usage() {
    std::fprintf(stderr,
        "\nmumemto (B200 engine) - find maximal [unique | exact] matches on the GPU.\n"
        "Usage: mumemto_exec [options] [input_fasta [...]]\n\n"
        "I/O options:\n"
        "\t-h, --help                      prints this usage message\n"
        "\t-i, --input           [FILE]    path to a file-list of genomes to use (overrides positional args)\n"
        "\t-o, --output          [PREFIX]  output prefix path\n"
        "\t-r, --no-revcomp                do not include the reverse complement of the sequences\n"
        "\t-b, --binary                    output binary format (multi-MUMs only)\n"
        "\t-A, --arrays-out                write LCP, BWT, and SA to file\n"
        "\t-a, --arrays-in       [PREFIX]  compute matches from precomputed LCP, BWT, SA (with shared PREFIX)\n"
        "\t-M, --merge                     output extra metadata to enable merging multi-MUMs\n"
        "\t-n, --anchor                    use anchor-based merging (uses first sequence as anchor)\n"
        "Exact match parameters:\n"
        "\t-l, --min-match-len   [INT]     minimum MUM or MEM length (default: 20)\n"
        "\t-k, --minimum-genomes [INT]     find matches in at least k sequences (k < 0: relative to N)\n"
        "\t-f, --per-seq-freq    [INT]     maximum number of occurences per sequence (default: 1)\n"
        "\t-F, --max-total-freq  [INT]     maximum number of total occurences of match\n"
        "Accepted for compatibility, no effect on results: -w, -m, -g, -s\n"
        "Not available (stage files of the CPU algorithm): -p, -P, -K\n"
        "Engine:\n"
        "\t-d, --device          [INT]     CUDA device (default 0)\n"
        "\t    --gpus            [INT]     partition-merge on N GPUs: the first sequence is the anchor of every partition, the others are\n"
        "\t                                split into N groups (one per GPU, -M -n), one NCCL all-gather, merge on the device;\n"
        "\t                                writes PREFIX.mums / PREFIX.athresh like `anchor_merge` (rows in anchor order)\n"
        "\t    --partitions      [INT]     the same workflow with P partitions run one after the other on one GPU\n"
        "\t    --shards          [INT]     ONE collection on K GPUs (--gpus K) or in K passes on one GPU: every GPU holds the whole collection and\n"
        "\t                                reports its slice of the suffix order; the slices are written in order = the single-GPU bytes\n"
        "\t    --trim[=INT]                shorten multi-MUMs that run past the end of a sequence and drop those shorter than INT (default 20)\n"
        "\t                                afterwards, like `mumemto trim`; off by default (the reference's output, bit for bit)\n\n");
    return 0;
}

static std::string slurp_plain(const std::string& path) {    // binary array files: no transparent gzip detection
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) fatal("cannot open " + path);
    std::string out((size_t)in.tellg(), '\0');
    in.seekg(0); in.read(out.data(), (std::streamsize)out.size());
    return out;
}

static void put_file(const std::string& path, const void* p, size_t n) {
    std::ofstream o(path, std::ios::binary);
    if (!o) fatal("cannot write " + path);
    o.write((const char*)p, (std::streamsize)n);
}

#define CHECK(call) do { if ((call) != 0) fatal(mmb_last_error()); } while (0)

int main(int argc, char** argv) {
    std::fprintf(stderr, "\033[1m\033[31m\nmumemto version: %s\033[m\033[0m\n", MUMEMTO_B200_VERSION);
    if (argc == 1) { std::fprintf(stderr, "\nUsage: mumemto_exec [options] [input_fasta [...]]\n\t-h, --help   prints detailed usage message\n"); return 0; }
    static struct option long_options[] = {
        {"help", no_argument, nullptr, 'h'}, {"input", required_argument, nullptr, 'i'}, {"output", required_argument, nullptr, 'o'},
        {"revcomp", no_argument, nullptr, 'r'}, {"minimum-genomes", required_argument, nullptr, 'k'}, {"no-overlap", no_argument, nullptr, 's'},
        {"modulus", required_argument, nullptr, 'm'}, {"from-parse", no_argument, nullptr, 'p'}, {"only-parse", no_argument, nullptr, 'P'},
        {"min-match-len", required_argument, nullptr, 'l'}, {"max-freq", required_argument, nullptr, 'F'}, {"arrays-out", no_argument, nullptr, 'A'},
        {"arrays-in", required_argument, nullptr, 'a'}, {"keep-temp-files", no_argument, nullptr, 'K'}, {"window", required_argument, nullptr, 'w'},
        {"rare", required_argument, nullptr, 'f'}, {"binary", no_argument, nullptr, 'b'}, {"merge", no_argument, nullptr, 'M'},
        {"anchor", no_argument, nullptr, 'n'}, {"use-gsacak", no_argument, nullptr, 'g'}, {"device", required_argument, nullptr, 'd'},
        {"gpus", required_argument, nullptr, 1000}, {"partitions", required_argument, nullptr, 1001}, {"trim", optional_argument, nullptr, 1002}, {"shards", required_argument, nullptr, 1003},
        {nullptr, 0, nullptr, 0}};
    Options o;
    int c;
    while ((c = getopt_long(argc, argv, "hi:F:o:w:sl:ra:AKk:p:m:f:bgMnPd:", long_options, nullptr)) >= 0) {
        switch (c) {
            case 'h': usage(); return 0;
            case 'i': o.input_list = optarg; break;
            case 'o': o.output_prefix = optarg; break;
            case 'w': case 'm': case 's': break;                    // PFP window / modulus / overlap: results do not depend on them
            case 'g': o.use_gsacak = true; break;                   // no effect on results either; only checked against -a like the reference
            case 'a': o.arrays_in = optarg; break;
            case 'r': o.use_rcomp = false; break;
            case 'k': o.num_distinct_docs = std::atoi(optarg); break;
            case 'l': o.min_match_len = (size_t)std::atoi(optarg); break;
            case 'F': o.max_mem_freq = std::atoi(optarg); break;
            case 'A': o.arrays_out = true; break;
            case 'f': o.rare_freq = std::atoi(optarg); break;
            case 'b': o.binary = true; break;
            case 'M': o.merge = true; break;
            case 'n': o.anchor_merge = true; break;
            case 'd': o.device = std::atoi(optarg); break;
            case 1000: o.gpus = std::atoi(optarg); break;
            case 1001: o.partitions = std::atoi(optarg); break;
            case 1003: o.shards = std::atoi(optarg); break;
            case 1002: o.trim = optarg ? std::atoi(optarg) : 20; break;
            case 'p': case 'P': case 'K':
                fatal("-p/-P/-K use stage files of the CPU algorithm (.dict/.parse); the GPU engine has no such stages");
            default: usage(); return 1;
        }
    }
    for (int i = optind; i < argc; ++i) o.files.push_back(argv[i]);

    // BuildOptions::validate
    if (!o.input_list.empty() && !fs::is_regular_file(o.input_list)) fatal("The provided file-list is not valid.");
    if (!o.input_list.empty() && !o.files.empty()) { status("Using filelist, ignoring positional args\n"); o.files.clear(); }
    if (o.input_list.empty() && o.files.empty() && o.arrays_in.empty()) fatal("Need to provide a file-list or files as positional args for processing.");
    fs::path pp(o.output_prefix);
    if (pp.parent_path().string().empty() && !pp.string().empty()) o.output_prefix = "./" + o.output_prefix;
    else if (!fs::exists(pp.parent_path())) fs::create_directories(pp.parent_path());
    if (o.use_gsacak && !o.arrays_in.empty()) fatal("--use-gsacak flag is incompatible with --arrays-in flag");
    if (o.anchor_merge) o.merge = true;
    if (!o.arrays_in.empty()) {
        if (!fs::is_regular_file(o.arrays_in + ".sa")) fatal("Missing *.sa file. Expected file: " + o.arrays_in + ".sa");
        else if (!fs::is_regular_file(o.arrays_in + ".lcp")) fatal("Missing *.lcp file. Expected file: " + o.arrays_in + ".lcp");
        else if (!fs::is_regular_file(o.arrays_in + ".bwt")) fatal("Missing *.bwt file. Expected file: " + o.arrays_in + ".bwt");
    }
    if (o.rare_freq < 0) fatal("Per-sequence MEM frequency must be > 0 (or 0 for no limit).");
    if (o.binary && o.rare_freq != 1) { status("binary output is not supported for multi-MEMs, ignoring flag\n"); o.binary = false; }

    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&]() { auto t = std::chrono::steady_clock::now(); double s = std::chrono::duration<double>(t - t0).count(); t0 = t; return s; };
    mmb_engine* e = nullptr;
    std::vector<uint64_t> arr_lengths; std::string arr_sa, arr_lcp, arr_bwt;         // -a: inputs of mmb_find_from_arrays
    if (!o.arrays_in.empty()) {
        // RefBuilder(prefix) (src/ref_builder.cpp:140-168): document lengths from PREFIX.lengths, no sequence data at all
        std::string lf = o.arrays_in + ".lengths";
        if (!fs::is_regular_file(lf)) fatal("Lengths file required for using intermediate files. File should match output prefix: " + lf);
        std::ifstream in(lf); std::string line;
        while (std::getline(in, line)) {
            std::vector<std::string> w; std::string cur; std::istringstream ss(line);
            while (std::getline(ss, cur, ' ')) w.push_back(cur);
            if (w.size() == 2) arr_lengths.push_back(std::stoull(w[1]));
            else if (w.size() == 3 && w[1] == "*") arr_lengths.push_back(std::stoull(w[2]));
        }
        status("reading the suffix array, LCP and BWT files");
        arr_sa = slurp_plain(o.arrays_in + ".sa"); arr_lcp = slurp_plain(o.arrays_in + ".lcp"); arr_bwt = slurp_plain(o.arrays_in + ".bwt");
        std::fprintf(stderr, " done.  (%.3f sec)\n", lap());
        CHECK(mmb_create(o.device, nullptr, &e));
    } else {
    // RefBuilder constructors: collect, check, de-duplicate
    std::vector<std::string> inputs;
    if (!o.input_list.empty()) {
        std::ifstream in(o.input_list); std::string line;
        while (std::getline(in, line)) {
            std::istringstream ss(line); std::string w;
            if (!(ss >> w)) continue;
            if (!fs::is_regular_file(w)) fatal("The following path in the input list is not valid: " + w);
            if (!is_fasta_name(w)) fatal("The following input-file is not a FASTA file: " + w);
            inputs.push_back(w);
        }
        std::vector<std::string> u; std::unordered_set<std::string> seen;
        for (auto& f : inputs) if (seen.insert(f).second) u.push_back(f);
        inputs.swap(u);
    } else {
        std::unordered_set<std::string> seen;
        for (auto& f : o.files) {
            if (!fs::is_regular_file(f)) fatal("The following file path is not valid: " + f);
            if (!is_fasta_name(f)) fatal("The following input-file is not a FASTA file: " + f);
            std::string norm = fs::absolute(f).lexically_normal().string();
            if (seen.insert(norm).second) inputs.push_back(norm);
        }
    }
    if (inputs.size() <= 1) fatal("Multiple FASTA inputs required. Perhaps split a multi-FASTA into multiple files?");


    status("parsing input files");
    const bool timing = std::getenv("MMB_CLI_TIMING") != nullptr;
    auto mark = [&, tm = std::chrono::steady_clock::now()](const char* what) mutable {
        if (!timing) return;
        auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "\n\t[timing] %-28s %8.3f s", what, std::chrono::duration<double>(t - tm).count());
        tm = t;
    };
    std::vector<fio::Source> sources;
    try { sources = fio::probe(inputs); } catch (const std::exception& ex) { fatal(ex.what()); }
    mark("probe input files");
    if (o.shards > 1) {
        // one collection, K slices of the suffix order (multi_gpu.cpp: run_sharded)
        if (o.merge || o.arrays_out || !o.arrays_in.empty() || o.partitions > 1) fatal("--shards cannot be combined with -M, -A, -a or --partitions");
        std::fprintf(stderr, " done.  (%.3f sec)\n", lap());
        mgpu::ShardRequest rq;
        rq.sources = &sources; rq.gpus = std::max(1, o.gpus); rq.shards = o.shards; rq.first_device = o.device; rq.use_rcomp = o.use_rcomp;
        rq.binary = o.binary && o.rare_freq == 1; rq.min_match_len = (uint32_t)o.min_match_len; rq.num_distinct_docs = o.num_distinct_docs;
        rq.rare_freq = o.rare_freq; rq.max_mem_freq = o.max_mem_freq; rq.trim_min_len = (o.rare_freq == 1 && o.trim > 0) ? (uint32_t)o.trim : 0;
        rq.output_prefix = o.output_prefix;
        std::string err;
        status(o.rare_freq == 1 ? "finding multi-MUMs: one slice of the suffix order per GPU" : "finding multi-MEMs: one slice of the suffix order per GPU");
        if (!mgpu::run_sharded(rq, err)) fatal(err);
        std::fprintf(stderr, " done.  (%.3f sec)\n", lap());
        return 0;
    }
    if (o.gpus > 1 || o.partitions > 1) {
        // partition-merge inside this process (multi_gpu.cpp): replaces `mumemto_exec -M -n` x P + `anchor_merge` of the reference
        // workflow (mumemto/merge_mums.py:171-183 -> src/merge_candidates.cpp:170-255)
        if (o.rare_freq != 1 || o.num_distinct_docs != 0) fatal("Merging not available for partial multi-MUM/MEMs");
        std::fprintf(stderr, " done.  (%.3f sec)\n", lap());
        mgpu::Request rq;
        rq.sources = &sources; rq.gpus = std::max(1, o.gpus); rq.partitions = o.partitions; rq.first_device = o.device;
        rq.use_rcomp = o.use_rcomp; rq.min_match_len = (uint32_t)o.min_match_len; rq.binary = o.binary; rq.output_prefix = o.output_prefix;
        std::string err;
        status("finding multi-MUMs: one partition per GPU, all-gather, merge");
        if (!mgpu::run(rq, err)) fatal(err);
        std::fprintf(stderr, " done.  (%.3f sec)\n", lap());
        return 0;
    }
    CHECK(mmb_create(o.device, nullptr, &e));
    mark("CUDA context + engine");
    {   // pooled pread -> pinned slices -> device (fasta_io.h), then the text build on the device
        std::vector<const fio::Source*> all;
        for (auto& sc : sources) all.push_back(&sc);
        fio::DeviceImages img;
        try { img = fio::upload(all, o.device); } catch (const std::exception& ex) { fatal(ex.what()); }
        mark("read + upload");
        if (mmb_load_fasta_device(e, img.d_raw, img.offsets.data(), img.sizes.data(), img.sizes.size(), o.use_rcomp) != 0)
            fatal(std::string(mmb_last_error()) + "\nPlease check the input files and ensure that it contains valid FASTA files.");
        img.release();
        std::vector<fio::Source>().swap(sources);
        mark("text build on the device");
    }
    std::fprintf(stderr, " done.  (%.3f sec)\n", lap());
    {   // PREFIX.lengths
        std::ofstream L(o.output_prefix + ".lengths");
        for (size_t d = 0; d < inputs.size(); ++d) {
            std::string cp = fs::canonical(inputs[d]).string();
            uint64_t total = 0; size_t nr = mmb_num_records(e, d);
            for (size_t r = 0; r < nr; ++r) { uint64_t len; mmb_record(e, d, r, nullptr, &len); total += len; }
            L << cp << " * " << total << std::endl;
            for (size_t r = 0; r < nr; ++r) { const char* name; uint64_t len; mmb_record(e, d, r, &name, &len); L << cp << " " << name << " " << len << std::endl; }
        }
    }

    }

    mmb_options mo; mmb_default_options(&mo);
    mo.min_match_len = (uint32_t)o.min_match_len; mo.num_distinct_docs = o.num_distinct_docs; mo.rare_freq = o.rare_freq; mo.max_mem_freq = o.max_mem_freq;
    mo.merge = o.merge; mo.anchor_merge = o.anchor_merge; mo.cli_normalise = 1; mo.want_text = o.binary ? 0 : 1;
    mo.keep_arrays = o.arrays_out ? 1 : 0;      // the tiled route only keeps the whole stream when asked
    mo.trim_min_len = (o.rare_freq == 1 && o.trim > 0) ? (uint32_t)o.trim : 0;
    bool mum_mode = (o.rare_freq == 1);
    status(mum_mode ? "finding multi-MUMs on the GPU" : "finding multi-MEMs on the GPU");
    if (!o.arrays_in.empty()) {
        uint64_t total = 0;                 // file_lcp::process walks total_length = |T| ranks (include/read_arrays.hpp:66-84), one less than the files hold
        for (uint64_t l : arr_lengths) total += (o.use_rcomp ? 2 : 1) * (l + 1);
        if (arr_sa.size() < 5 * total || arr_lcp.size() < 5 * total || arr_bwt.size() < total) fatal("array files are shorter than the sequence lengths imply");
        CHECK(mmb_find_from_arrays(e, arr_lengths.data(), arr_lengths.size(), o.use_rcomp, (const uint8_t*)arr_sa.data(), (const uint8_t*)arr_lcp.data(),
                                   (const uint8_t*)arr_bwt.data(), total, &mo));
        o.arrays_out = false;               // -A only applies to the array producers (pfp_lcp / gsacak_lcp)
    } else
    CHECK(mmb_find(e, &mo));
    std::fprintf(stderr, " done.  (%.3f sec)\n", lap());

    size_t N = mmb_num_docs(e);
    uint64_t count = mmb_num_matches(e);
    if (!mum_mode || !o.binary) {
        // device -> pinned slices -> write(2), copy and write overlapped (fasta_io.h)
        try { fio::write_device_buffer(o.output_prefix + (mum_mode ? ".mums" : ".mems"), mmb_text_device(e), mmb_text_size(e)); }
        catch (const std::exception& ex) { fatal(ex.what()); }
    } else {
        if (count == 0) fatal("no multi-MUMs found: nothing to write in binary format");
        long n = (long)N, k = o.num_distinct_docs;
        k = k < -n ? 2 : (k <= 0 ? n + k : (k == 1 ? 2 : std::min(k, n)));
        uint16_t flags = (uint16_t)(0x8000 | (k < n ? 0x2000 : 0));
        uint64_t ns = N, nm = count;
        std::vector<uint32_t> len(nm); std::vector<int64_t> off(nm * ns); std::vector<uint8_t> bits((nm * ns + 7) / 8);
        CHECK(mmb_copy_mums(e, len.data(), off.data(), nullptr));
        CHECK(mmb_copy_mum_strand_bits(e, bits.data()));
        std::ofstream b(o.output_prefix + ".bumbl", std::ios::binary);
        b.write((char*)&flags, 2); b.write((char*)&ns, 8); b.write((char*)&nm, 8);
        b.write((char*)len.data(), nm * 4); b.write((char*)off.data(), nm * ns * 8); b.write((char*)bits.data(), bits.size());
    }
    if (o.merge) {
        std::vector<uint16_t> t(mmb_threshold_count(e));
        CHECK(mmb_copy_thresholds(e, t.data()));
        if (o.anchor_merge) put_file(o.output_prefix + ".athresh", t.data(), t.size());          // first L_0 + 1 entries = half of the bytes
        else {
            // string-based merge metadata, include/mem_finder.hpp:116-157: thresholds re-laid per MUM in anchor order
            std::vector<uint32_t> len(count); std::vector<int64_t> off(count * N);
            CHECK(mmb_copy_mums(e, len.data(), off.data(), nullptr));
            std::vector<std::pair<uint64_t, uint64_t>> pos;
            for (uint64_t i = 0; i < count; ++i) pos.push_back({(uint64_t)off[i * N], len[i]});
            std::sort(pos.begin(), pos.end(), [](auto& a, auto& b) { return a.first < b.first; });
            std::vector<uint16_t> f, r; uint64_t half = t.size() / 2;
            for (auto& [at, l] : pos) {
                uint64_t rev = 2 * half - at - l - 1;
                for (uint64_t x = 0; x < l; ++x) { f.push_back(t[at + x] < l - x ? t[at + x] : 0); r.push_back(t.at(rev + x) < l - x ? t[rev + x] : 0); }
                f.push_back(0); r.push_back(0);
            }
            put_file(o.output_prefix + ".thresh", f.data(), f.size() * 2);
            put_file(o.output_prefix + ".thresh_rev", r.data(), r.size() * 2);
        }
    }
    if (o.arrays_out) {
        mmb_stats st; mmb_get_stats(e, &st);
        uint64_t m = st.text_symbols + 1;
        std::vector<uint64_t> sa(m), lcp(m); std::vector<uint8_t> bwt(m);
        CHECK(mmb_copy_arrays(e, sa.data(), lcp.data(), bwt.data()));
        FILE* fa = std::fopen((o.output_prefix + ".sa").c_str(), "wb"); FILE* fl = std::fopen((o.output_prefix + ".lcp").c_str(), "wb");
        if (!fa || !fl) fatal("cannot write array files");
        for (uint64_t i = 0; i < m; ++i) { std::fwrite(&sa[i], 5, 1, fa); std::fwrite(&lcp[i], 5, 1, fl); }     // SSABYTES = THRBYTES = 5
        std::fclose(fa); std::fclose(fl);
        put_file(o.output_prefix + ".bwt", bwt.data(), bwt.size());
    }
    mmb_stats st; mmb_get_stats(e, &st);
    status(""); std::fprintf(stderr, "Found %llu matches!\n", (unsigned long long)count);
    std::fprintf(stderr, "\tdevice time: text %.2f ms, suffix array %.2f ms (%u rounds), LCP/BWT %.2f ms, detection %.2f ms, output %.2f ms; %llu kernel launches\n\n",
                 st.ms_text, st.ms_sa, st.sa_rounds, st.ms_lcp, st.ms_detect, st.ms_emit, (unsigned long long)st.kernel_launches);
    mmb_destroy(e);
    return 0;
}
