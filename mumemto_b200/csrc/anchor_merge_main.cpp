// mumemto_b200/csrc/anchor_merge_main.cpp -- `anchor_merge p1.mums p2.mums ... -o out` on the GPU engine.
//
// Mirrors main / parse_candidate / validate_candidate_path of the reference's src/merge_candidates.cpp (:40-95,170-255) and the
// .mums / .bumbl readers and writers of include/mumsio.hpp (:40-194,328-385); the fold itself is mmb_anchor_merge (csrc/merge.cu).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "mumemto_b200.h"

namespace fs = std::filesystem;

[[noreturn]] static void die(const std::string& m) { std::cerr << "anchor_merge: " << m << std::endl; std::exit(1); }
static bool ends_with(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && !s.compare(s.size() - suf.size(), suf.size(), suf); }

struct Part { std::vector<uint32_t> len; std::vector<int64_t> off; std::vector<uint8_t> plus; std::vector<uint16_t> nb; uint64_t ndocs = 0; };

static Part load(const std::string& path) {
    bool bumbl = ends_with(path, ".bumbl");
    if (!bumbl && !ends_with(path, ".mums")) die("Invalid input: " + path + ". Inputs must explicitly end with .mums or .bumbl.");
    std::string base = path.substr(0, path.size() - (bumbl ? 6 : 5));
    if (!fs::exists(path)) die("Could not find MUMs file: " + path);
    if (!fs::exists(base + ".athresh")) die("Could not find threshold file: " + base + ".athresh");
    Part P;
    {
        std::ifstream t(base + ".athresh", std::ios::binary | std::ios::ate);
        std::streamsize sz = t.tellg();
        if (sz % 2) die("File size is not a multiple of uint16_t");
        t.seekg(0); P.nb.resize(sz / 2); t.read((char*)P.nb.data(), sz);
    }
    if (bumbl) {
        std::ifstream in(path, std::ios::binary);
        uint16_t flags; uint64_t ns, nm;
        in.read((char*)&flags, 2); in.read((char*)&ns, 8); in.read((char*)&nm, 8);
        if (flags & (1u << 13)) die("Cannot parse partial bumbl: header indicates partial MUMs");
        P.ndocs = ns; P.len.resize(nm); P.off.resize(nm * ns); P.plus.resize(nm * ns);
        if (flags & (1u << 15)) in.read((char*)P.len.data(), nm * 4);
        else { std::vector<uint16_t> l16(nm); in.read((char*)l16.data(), nm * 2); for (uint64_t i = 0; i < nm; ++i) P.len[i] = l16[i]; }
        in.read((char*)P.off.data(), nm * ns * 8);
        std::vector<uint8_t> packed((nm * ns + 7) / 8); in.read((char*)packed.data(), packed.size());
        if (!in) die("Unexpected EOF while reading bumbl file");
        for (uint64_t i = 0; i < nm * ns; ++i) P.plus[i] = (packed[i / 8] >> (7 - i % 8)) & 1;
    } else {
        std::ifstream in(path); std::string line;
        while (std::getline(in, line)) {
            if (line.empty()) continue;
            std::istringstream ss(line); std::string a, b, c; ss >> a >> b >> c;
            if (a.empty() || b.empty() || c.empty()) die("Malformed MUMs line: " + line);
            P.len.push_back((uint32_t)std::stoul(a));
            uint64_t cols = 0;
            { std::stringstream x(b); std::string tok; while (std::getline(x, tok, ',')) { if (tok.empty()) die("Cannot parse partial MUMs: empty offset encountered. Filter to strict MUMs."); P.off.push_back(std::stoll(tok)); ++cols; } }
            uint64_t scols = 0;
            { std::stringstream x(c); std::string tok; while (std::getline(x, tok, ',')) { P.plus.push_back(tok == "+"); ++scols; } }
            if (cols != scols) die("Offsets and strands column size mismatch in MUMs file");
            if (P.ndocs && cols != P.ndocs) die("inconsistent number of columns in " + path);
            P.ndocs = cols;
        }
    }
    return P;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::cerr << "Usage: " << argv[0] << " <input_paths>... -o <output_prefix> [-v]" << std::endl; return 1; }
    std::vector<std::string> paths; std::string out = "merged"; bool verbose = false; int device = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-o" && i + 1 < argc) out = argv[++i]; else if (a == "-v") verbose = true;
        else if (a == "-d" && i + 1 < argc) device = std::atoi(argv[++i]); else paths.push_back(a);
    }
    if (paths.size() < 2) die("requires at least two input files");
    std::vector<Part> parts;
    for (auto& p : paths) { if (verbose) std::cerr << "reading " << p << std::endl; parts.push_back(load(p)); }
    uint64_t anchor_len = parts[0].nb.size();
    for (auto& p : parts) if (p.nb.size() != anchor_len) die("threshold files differ in length: the partitions do not share their first sequence");
    std::vector<const void*> L, O, S, T; std::vector<uint64_t> n, nd;
    for (auto& p : parts) {
        if (p.len.empty()) { p.len.push_back(0); p.off.assign(std::max<uint64_t>(p.ndocs, 2), 0); p.plus.assign(std::max<uint64_t>(p.ndocs, 2), 0); n.push_back(0); if (!p.ndocs) p.ndocs = 2; }
        else n.push_back(p.len.size());
        L.push_back(p.len.data()); O.push_back(p.off.data()); S.push_back(p.plus.data()); T.push_back(p.nb.data()); nd.push_back(p.ndocs);
    }
    mmb_engine* e = nullptr;
    if (mmb_create(device, nullptr, &e)) die(mmb_last_error());
    std::cerr << "merging " << fs::path(paths[0]).filename().string() << " and " << fs::path(paths[1]).filename().string() << std::endl;
    if (mmb_anchor_merge(e, parts.size(), L.data(), n.data(), O.data(), S.data(), T.data(), nd.data(), anchor_len, 0)) die(mmb_last_error());

    bool as_bumbl = ends_with(out, ".bumbl"), as_mums = ends_with(out, ".mums");
    std::string prefix = as_bumbl ? out.substr(0, out.size() - 6) : (as_mums ? out.substr(0, out.size() - 5) : out);
    uint64_t nm = mmb_merge_num_mums(e), ns = mmb_merge_num_docs(e);
    std::vector<uint16_t> nb(anchor_len);
    if (as_bumbl) {
        if (nm == 0) die("no merged MUMs: nothing to write in binary format");
        std::vector<uint32_t> len(nm); std::vector<int64_t> off(nm * ns); std::vector<uint8_t> plus(nm * ns);
        if (mmb_merge_copy(e, len.data(), off.data(), plus.data(), nb.data())) die(mmb_last_error());
        std::vector<uint8_t> packed((nm * ns + 7) / 8, 0);
        for (uint64_t i = 0; i < nm * ns; ++i) if (plus[i]) packed[i / 8] |= 1u << (7 - i % 8);
        uint16_t flags = 0x8000;
        std::ofstream b(prefix + ".bumbl", std::ios::binary);
        b.write((char*)&flags, 2); b.write((char*)&ns, 8); b.write((char*)&nm, 8);
        b.write((char*)len.data(), nm * 4); b.write((char*)off.data(), nm * ns * 8); b.write((char*)packed.data(), packed.size());
    } else {
        std::vector<char> text(mmb_merge_text_size(e));
        if (mmb_merge_copy_text(e, text.data())) die(mmb_last_error());
        if (mmb_merge_copy(e, nullptr, nullptr, nullptr, nb.data())) die(mmb_last_error());
        std::ofstream o(prefix + ".mums", std::ios::binary); o.write(text.data(), text.size());
    }
    std::ofstream t(prefix + ".athresh", std::ios::binary); t.write((char*)nb.data(), nb.size() * 2);
    mmb_destroy(e);
    std::cerr << "done." << std::endl;
    return 0;
}
