// mumemto_b200/csrc/extract_mums_main.cpp -- `extract_mums -m PREFIX.mums [-l PREFIX.lengths] [-o OUT.fa] [-t]`.
//
// Host-side file tool of the string-based merge: writes every multi-MUM of a partition as one FASTA record
// `>mum_<i>` + the MUM's bases in the partition's FIRST sequence + '#', in FILE order.  Same flags, defaults, messages and exit codes as
// the reference's src/extract_mums.cpp (:22-31 usage, :43-78 first sequence = first token of the first .lengths line, all records
// concatenated; :80-92 only (length, first offset) of every row is read, .mums or .bumbl; :94-118 output; :120-198 option handling).
// No device work: the MUM search on the extracted records is `mumemto_exec` again (mumemto_b200/merge_mums.py drives both).
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <getopt.h>

static bool ends_with(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && !s.compare(s.size() - suf.size(), suf.size(), suf); }
static bool is_file(const std::string& p) { return std::filesystem::is_regular_file(p); }

static void usage() {
    std::fprintf(stderr, "\nextract_mums - extract MUMs from a MUM and length file\n");
    std::fprintf(stderr, "Usage: extract_mums [options] -m mum_file -o output_file\n\n");
    std::fprintf(stderr, "Options:\n");
    std::fprintf(stderr, "\t%-32sprints this usage message\n", "-h, --help");
    std::fprintf(stderr, "\t%-22s%-10spath to a mum file\n", "-m, --mums", "[FILE]");
    std::fprintf(stderr, "\t%-22s%-10spath to a length file (optional, uses associated length file if not provided)\n", "-l, --lengths", "[FILE]");
    std::fprintf(stderr, "\t%-22s%-10soutput path\n", "-o, --output", "[FILE]");
    std::fprintf(stderr, "\t%-32sdo not add terminator (#) to end of each MUM sequence\n", "-t, --no-terminator");
}

// all records of the (optionally gzip-compressed) FASTA file, concatenated: header lines dropped, line ends stripped
static bool read_sequence(const std::string& path, std::string& seq) {
    gzFile f = gzopen(path.c_str(), "r");
    if (!f) { std::cerr << "Error: Unable to open FASTA file" << std::endl; return false; }
    gzbuffer(f, 1 << 20);
    std::vector<char> buf(1 << 22);
    bool line_start = true, in_header = false;
    int got;
    while ((got = gzread(f, buf.data(), (unsigned)buf.size())) > 0) {
        for (int i = 0; i < got; ++i) {
            char c = buf[i];
            if (c == '\n') { line_start = true; in_header = false; continue; }
            if (line_start && c == '>') in_header = true;
            line_start = false;
            if (!in_header && c != '\r') seq.push_back(c);
        }
    }
    gzclose(f);
    return true;
}

struct Row { uint64_t off0, len; };

static bool read_rows(const std::string& path, std::vector<Row>& rows) {
    if (ends_with(path, ".bumbl")) {
        std::ifstream in(path, std::ios::binary);
        uint16_t flags = 0; uint64_t ns = 0, nm = 0;
        in.read((char*)&flags, 2); in.read((char*)&ns, 8); in.read((char*)&nm, 8);
        if (!in) { std::cerr << "Error: truncated bumbl header" << std::endl; return false; }
        std::vector<uint32_t> len(nm);
        if (flags & (1u << 15)) in.read((char*)len.data(), nm * 4);
        else { std::vector<uint16_t> l16(nm); in.read((char*)l16.data(), nm * 2); for (uint64_t i = 0; i < nm; ++i) len[i] = l16[i]; }
        std::vector<int64_t> row(ns);
        for (uint64_t i = 0; i < nm; ++i) {
            in.read((char*)row.data(), ns * 8);
            if (!in) { std::cerr << "Error: truncated bumbl file" << std::endl; return false; }
            rows.push_back({(uint64_t)row[0], len[i]});
        }
        return true;
    }
    std::ifstream in(path);
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        size_t tab = line.find('\t');
        if (tab == std::string::npos) { std::cerr << "Error: malformed MUMs line: " << line << std::endl; return false; }
        size_t comma = line.find_first_of(",\t", tab + 1);
        std::string first = line.substr(tab + 1, comma == std::string::npos ? std::string::npos : comma - tab - 1);
        if (first.empty()) { std::cerr << "Error: Cannot parse partial MUMs: empty offset encountered" << std::endl; return false; }
        rows.push_back({std::stoull(first), std::stoull(line.substr(0, tab))});
    }
    return true;
}

int main(int argc, char** argv) {
    if (argc == 1) { usage(); return 1; }
    std::string mums, out, lengths;
    bool terminator = true;
    static struct option long_options[] = {{"help", no_argument, nullptr, 'h'}, {"mums", required_argument, nullptr, 'm'},
        {"lengths", required_argument, nullptr, 'l'}, {"output", required_argument, nullptr, 'o'}, {"no-terminator", no_argument, nullptr, 't'}, {0, 0, 0, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "hm:l:o:t", long_options, nullptr)) >= 0) {
        switch (c) {
            case 'h': usage(); return 0;
            case 'm': mums = optarg; break;
            case 'l': lengths = optarg; break;
            case 'o': out = optarg; break;
            case 't': terminator = false; break;
            default: usage(); return 1;
        }
    }
    if (optind < argc && mums.empty()) mums = argv[optind];
    if (mums.empty()) { std::cerr << "Error: No mum file provided\n"; return 1; }
    if (!ends_with(mums, ".mums") && !ends_with(mums, ".bumbl")) mums += ".mums";
    if (!is_file(mums)) { std::cerr << "Error: Invalid mum file provided: " << mums << std::endl; return 1; }
    std::string stem = mums.substr(0, mums.find_last_of('.'));
    if (lengths.empty()) lengths = stem + ".lengths";
    if (!is_file(lengths)) {
        std::cerr << "Error: Invalid lengths file: " << lengths << std::endl;
        std::cerr << "Provide a lengths file with the -l option or ensure the mum file has an associated lengths file" << std::endl;
        return 1;
    }
    if (out.empty()) out = stem + "_mums.fa"; else if (!ends_with(out, ".fa")) out += ".fa";

    std::string first_line;
    { std::ifstream lf(lengths); if (!std::getline(lf, first_line)) { std::cerr << "Error: Length file is empty" << std::endl; return 1; } }
    std::string fasta = first_line.substr(0, first_line.find_first_of(" \t\n\r"));
    if (!is_file(fasta)) { std::cerr << "Error: Invalid FASTA path in length file" << std::endl; return 1; }
    std::string seq;
    if (!read_sequence(fasta, seq)) return 1;
    if (seq.empty()) { std::cerr << "Error: Empty input file found" << std::endl; return 1; }
    std::vector<Row> rows;
    if (!read_rows(mums, rows)) return 1;

    std::string text;
    text.reserve(rows.size() * 64);
    uint64_t count = 0;
    for (const Row& r : rows) {
        if (r.off0 > seq.size()) { std::cerr << "Error: MUM offset beyond the first sequence" << std::endl; return 1; }
        text += ">mum_" + std::to_string(count++) + "\n";
        text.append(seq, r.off0, r.len);                       // clipped at the end of the sequence, like std::string::substr
        if (terminator) text.push_back('#');
        text.push_back('\n');
    }
    std::ofstream o(out, std::ios::binary);
    o.write(text.data(), (std::streamsize)text.size());
    return o.good() ? 0 : 1;
}
