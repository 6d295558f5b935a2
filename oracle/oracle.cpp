// oracle/oracle.cpp -- TEST INFRASTRUCTURE: single-threaded CPU restatement of the reference's FASTA -> .mums/.mems path and of
// its anchor-based partition merge.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may run this; the
// product (mumemto_b200/) never links, imports or executes it.
//
// Parity status: PINNED.  Its outputs are checked (tests/test_oracle.py) against (1) the md5 table of SURVEY.md Appendix A and
// (2) oracle/_ref/mumemto_exec + anchor_merge, i.e. the reference's own sources compiled unmodified (oracle/Makefile).
//
// What follows which reference lines (all paths relative to /root/reference):
//   read_fasta()          include/kseq.h:173-219 (record/line rules), src/ref_builder.cpp:232-241 (toupper per record)
//   build_text()          src/ref_builder.cpp:211-314 (fwd '$' rc '$' per file; records joined without separator), :29-50 (complement)
//   normalise()           include/pfp_mum.hpp:80-198 (validate + set_parameters)
//   suffix stream         include/pfp_lcp_mum.hpp:115-231,337-341 (rank j -> SA, LCP, BWT, doc over T followed by a smallest sentinel)
//   Detector::push()      include/mem_finder.hpp:161-170,304-355,430-449 (interval stack, no flush at end of stream)
//   Detector::docs_ok()   include/mem_finder.hpp:265-289
//   Detector::emit_mum()  include/mem_finder.hpp:357-428;  emit_mem(): :210-263;  finish(): :104-158,451-503
//   merge_main()          src/merge_candidates.cpp:62-157,170-255 and include/mumsio.hpp:40-103,328-385
//
// Usage mirrors the reference CLI:  mumemto_oracle [-o prefix] [-l n] [-k n] [-f n] [-F n] [-r] [-b] [-M] [-n] [-A] a.fa b.fa ...
//                                   mumemto_oracle merge p1.mums p2.mums ... -o out
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

#include "sais.hpp"

namespace fs = std::filesystem;
typedef uint64_t u64;

[[noreturn]] static void die(const std::string& m) { fprintf(stderr, "oracle: %s\n", m.c_str()); exit(1); }

// ------------------------------------------------------------------------------------------------ FASTA
struct Record { std::string name, seq; };

static std::string slurp(const std::string& path) {
    gzFile f = gzopen(path.c_str(), "r");
    if (!f) die("cannot open " + path);
    std::string out; char buf[1 << 16]; int k;
    while ((k = gzread(f, buf, sizeof buf)) > 0) out.append(buf, k);
    gzclose(f);
    return out;
}

// kseq rules: skip to the first '>' or '@'; name = up to first whitespace; rest of that line is a comment; then every following
// line is sequence until a line that starts with '>', '@' or '+'.  '\n' is dropped, a '\r' that ends a line is dropped.
static std::vector<Record> read_fasta(const std::string& path) {
    std::string b = slurp(path);
    std::vector<Record> recs;
    size_t i = 0, n = b.size();
    while (i < n && b[i] != '>' && b[i] != '@') ++i;
    while (i < n) {
        ++i;                                             // header marker
        Record r;
        while (i < n && !isspace((unsigned char)b[i])) r.name += b[i++];
        while (i < n && b[i] != '\n') ++i;               // comment
        if (i < n) ++i;
        bool fastq = false;
        while (i < n) {                                  // sequence lines
            char c = b[i];
            if (c == '>' || c == '@') break;
            if (c == '+') { fastq = true; break; }
            if (c == '\n') { ++i; continue; }
            size_t e = b.find('\n', i);
            if (e == std::string::npos) e = n;
            r.seq.append(b, i, e - i);
            if (r.seq.size() > 1 && r.seq.back() == '\r') r.seq.pop_back();
            i = (e < n) ? e + 1 : n;
        }
        if (fastq) die("FASTQ input is not supported by the oracle: " + path);
        recs.push_back(std::move(r));
    }
    return recs;
}

static unsigned char complement(unsigned char c) {      // seqtk table, src/ref_builder.cpp:29-38
    static const char* from = "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn";
    static const char* to   = "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn";
    const char* p = c ? strchr(from, c) : nullptr;
    return p ? (unsigned char)to[p - from] : c;
}

// ------------------------------------------------------------------------------------------------ parameters
struct Params {
    std::string prefix = "output";
    std::vector<std::string> files;
    bool revcomp = true, binary = false, merge = false, anchor = false, arrays_out = false;
    long min_len = 20, k = 0, f = 1, F = 0;
};

static void normalise(Params& p, long N) {
    if (p.anchor) p.merge = true;
    if (p.f < 0) die("Per-sequence MEM frequency must be > 0 (or 0 for no limit).");
    if (p.binary && p.f != 1) p.binary = false;
    if (p.k < -N) p.k = 2; else if (p.k <= 0) p.k = N + p.k; else if (p.k == 1) p.k = 2; else if (p.k >= N) p.k = N;
    if (p.merge && p.k != N) die("Merging not available for partial multi-MUM/MEMs");
    if (p.merge && p.f != 1) die("Merging not available for multi-MEMs");
    if (p.F < -N || p.F == 1) p.F = 0; else if (p.F < 0) p.F = N + p.F;
    if (p.f > 0 && (p.F == 0 || p.F > p.f * N)) p.F = p.f * N;
}

// ------------------------------------------------------------------------------------------------ text
struct Text {
    std::vector<uint8_t> t;
    std::vector<u64> doc_start;       // N+1 entries
    std::vector<u64> half;            // per doc: L+1 (revcomp) -- "doc_lens" after halving, include/mem_finder.hpp:74-78
};

static Text build_text(const Params& p, std::vector<std::vector<Record>>& all) {
    Text T; T.doc_start.push_back(0);
    for (auto& recs : all) {
        std::string fwd;
        for (auto& r : recs) { for (auto& c : r.seq) c = (char)toupper((unsigned char)c); fwd += r.seq; }
        if (fwd.empty()) die("Empty input file found");
        T.t.insert(T.t.end(), fwd.begin(), fwd.end()); T.t.push_back('$');
        if (p.revcomp) {
            for (size_t i = fwd.size(); i-- > 0;) T.t.push_back(complement((unsigned char)fwd[i]));
            T.t.push_back('$');
        }
        T.doc_start.push_back(T.t.size());
        u64 len = T.doc_start.back() - T.doc_start[T.doc_start.size() - 2];
        T.half.push_back(p.revcomp ? len / 2 : len);
    }
    return T;
}

// ------------------------------------------------------------------------------------------------ detector
struct Detector {
    const Params& p; const Text& T; size_t N;
    std::vector<u64> sa; std::vector<uint8_t> bwt; std::vector<u64> doc;   // whole stream kept: the oracle favours clarity
    struct Open { u64 first, len, before; };
    std::vector<Open> open;
    u64 last_change = 0, prev_lcp = 0, count = 0;
    std::string text_out;
    std::vector<uint32_t> b_len; std::vector<int64_t> b_start; std::vector<char> b_strand;
    std::vector<uint16_t> thresh; std::vector<std::pair<u64, u64>> mum_pos;

    Detector(const Params& p_, const Text& T_) : p(p_), T(T_), N(T_.half.size()) {
        open.push_back({0, 0, 0});
        if (p.merge) thresh.assign(T.half[0] * 2, 0);
    }
    bool docs_ok(u64 s, u64 e) const {
        std::vector<u64> seen(N + 1, 0); u64 distinct = 0;
        for (u64 i = s; i <= e; ++i) {
            u64 d = doc[i];
            if (seen[d]++ == 0) ++distinct;
            else if (p.f && seen[d] > (u64)p.f) return false;
        }
        return distinct >= (u64)p.k;
    }
    void emit_mum(u64 len, u64 s, u64 e) {
        std::vector<int64_t> off(N, -1); std::vector<char> st(N, 0);
        for (u64 i = s; i <= e; ++i) {
            u64 d = doc[i], q = sa[i] - T.doc_start[d]; char c = '+';
            if (p.revcomp && q >= T.half[d]) {
                c = '-';
                if (q + len >= 2 * T.half[d]) return;
                q = 2 * T.half[d] - q - len - 1;
            }
            off[d] = (int64_t)q; st[d] = c;
        }
        size_t i = 0;
        while (i + 1 < N && st[i] == 0) ++i;
        if (st[i] == '-') return;
        if (p.merge) mum_pos.push_back({(u64)off[0], len});
        ++count;
        if (p.binary) {
            b_len.push_back((uint32_t)len); b_start.insert(b_start.end(), off.begin(), off.end());
            b_strand.insert(b_strand.end(), st.begin(), st.end());
            return;
        }
        std::string a = std::to_string(len) + "\t", b;
        for (size_t d = 0; d < N; ++d) {
            if (off[d] >= 0) { a += std::to_string(off[d]); b += st[d]; }
            if (d + 1 < N) { a += ','; b += ','; }
        }
        text_out += a + "\t" + b + "\n";
    }
    void emit_mem(u64 len, u64 s, u64 e) {
        std::string a, b, c;
        for (u64 i = s; i <= e; ++i) {
            u64 d = doc[i], q = sa[i] - T.doc_start[d]; char st = '+';
            if (p.revcomp && q >= T.half[d]) { st = '-'; q = 2 * T.half[d] - q - len - (i == e ? 0 : 1); }   // unsigned wrap kept
            a += std::to_string(q); b += std::to_string(d); c += st;
            if (i < e) { a += ','; b += ','; c += ','; }
        }
        text_out += std::to_string(len) + "\t" + a + "\t" + b + "\t" + c + "\n"; ++count;
    }
    void push(u64 j, uint8_t ch, u64 d, u64 pos, u64 lcp) {
        u64 start = j - 1;
        while (lcp < open.back().len) {
            Open o = open.back(); open.pop_back();
            u64 m = j - o.first;
            if (o.len >= (u64)p.min_len && m >= (u64)p.k && (p.F == 0 || m <= (u64)p.F) && docs_ok(o.first, j - 1)) {
                if (p.merge) {
                    u64 nb = std::min<u64>(std::max(o.before, lcp), 65535);
                    for (u64 i = o.first; i < j; ++i) if (doc[i] == 0) { thresh[sa[i] - T.doc_start[0]] = (uint16_t)nb; break; }
                }
                if (last_change > o.first) { if (p.f == 1) emit_mum(o.len, o.first, j - 1); else emit_mem(o.len, o.first, j - 1); }
            }
            start = o.first; prev_lcp = o.before;
        }
        if (lcp > open.back().len && lcp >= (u64)p.min_len) open.push_back({start, lcp, prev_lcp});
        if (j == 0 || bwt.back() != ch) last_change = j;
        sa.push_back(pos); bwt.push_back(ch); doc.push_back(d);
        prev_lcp = lcp;
    }
    void finish() {
        auto put = [&](const std::string& ext, const void* ptr, size_t bytes) {
            std::ofstream o(p.prefix + ext, std::ios::binary); o.write((const char*)ptr, bytes);
        };
        if (p.f != 1) put(".mems", text_out.data(), text_out.size());
        else if (!p.binary) put(".mums", text_out.data(), text_out.size());
        else {
            if (b_len.empty()) die("zero MUMs with -b: the reference dereferences an empty vector here (include/mem_finder.hpp:461-462)");
            u64 nm = b_len.size(), ns = N;
            std::vector<uint8_t> packed((nm * ns + 7) / 8, 0);
            for (u64 i = 0; i < nm * ns; ++i) if (b_strand[i] == '+') packed[i / 8] |= 1u << (7 - i % 8);
            uint16_t flags = 0x8000 | (((u64)p.k < N) ? 0x2000 : 0);
            std::ofstream o(p.prefix + ".bumbl", std::ios::binary);
            o.write((char*)&flags, 2); o.write((char*)&ns, 8); o.write((char*)&nm, 8);
            o.write((char*)b_len.data(), nm * 4); o.write((char*)b_start.data(), nm * ns * 8); o.write((char*)packed.data(), packed.size());
        }
        if (p.anchor) put(".athresh", thresh.data(), T.half[0] * 2);
        else if (p.merge) {
            std::sort(mum_pos.begin(), mum_pos.end(), [](auto& a, auto& b) { return a.first < b.first; });
            std::vector<uint16_t> f, r;
            for (auto& [at, len] : mum_pos) {
                u64 rev = 2 * T.half[0] - at - len - 1;
                for (u64 x = 0; x < len; ++x) {
                    f.push_back(thresh[at + x] < len - x ? thresh[at + x] : 0);
                    r.push_back(thresh.at(rev + x) < len - x ? thresh[rev + x] : 0);
                }
                f.push_back(0); r.push_back(0);
            }
            put(".thresh", f.data(), f.size() * 2); put(".thresh_rev", r.data(), r.size() * 2);
        }
    }
};

// ------------------------------------------------------------------------------------------------ main path
static void put5(FILE* f, u64 v) { fwrite(&v, 5, 1, f); }

static int find_main(int argc, char** argv) {
    Params p;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) die("missing value for " + a); return argv[++i]; };
        if (a == "-o") p.prefix = val(); else if (a == "-l") p.min_len = atol(val().c_str());
        else if (a == "-k") p.k = atol(val().c_str()); else if (a == "-f") p.f = atol(val().c_str());
        else if (a == "-F") p.F = atol(val().c_str()); else if (a == "-r") p.revcomp = false;
        else if (a == "-b") p.binary = true; else if (a == "-M") p.merge = true; else if (a == "-n") p.anchor = true;
        else if (a == "-A") p.arrays_out = true; else if (a == "-g" || a == "-K") {}
        else if (a == "-w" || a == "-m") val();
        else if (a == "-i") { std::ifstream in(val()); std::string line; while (std::getline(in, line)) { std::istringstream s(line); std::string w; if (s >> w) p.files.push_back(w); } }
        else p.files.push_back(a);
    }
    std::vector<std::string> uniq; std::unordered_set<std::string> seen;
    for (auto& f : p.files) {
        if (!fs::is_regular_file(f)) die("The following file path is not valid: " + f);
        bool ok = false;
        for (const char* e : {".fa", ".fasta", ".fna", ".fa.gz", ".fasta.gz", ".fna.gz"})
            if (f.size() >= strlen(e) && f.compare(f.size() - strlen(e), strlen(e), e) == 0) ok = true;
        if (!ok) die("The following input-file is not a FASTA file: " + f);
        std::string norm = fs::absolute(f).lexically_normal().string();
        if (seen.insert(norm).second) uniq.push_back(norm);
    }
    if (uniq.size() <= 1) die("Multiple FASTA inputs required. Perhaps split a multi-FASTA into multiple files?");
    p.files = uniq;
    fs::path pp(p.prefix);
    if (pp.parent_path().string().empty() && !p.prefix.empty()) p.prefix = "./" + p.prefix;
    else if (!fs::exists(pp.parent_path())) fs::create_directories(pp.parent_path());
    normalise(p, (long)p.files.size());

    std::vector<std::vector<Record>> all;
    for (auto& f : p.files) all.push_back(read_fasta(f));
    {   // PREFIX.lengths, src/ref_builder.cpp:193-209
        std::ofstream L(p.prefix + ".lengths");
        for (size_t i = 0; i < all.size(); ++i) {
            std::string cp = fs::canonical(p.files[i]).string(); u64 tot = 0;
            for (auto& r : all[i]) tot += r.seq.size();
            L << cp << " * " << tot << "\n";
            for (auto& r : all[i]) L << cp << " " << r.name << " " << r.seq.size() << "\n";
        }
    }
    Text T = build_text(p, all);
    int64_t n = (int64_t)T.t.size() + 1;
    std::vector<int64_t> s(n), SA(n), LCP(n);
    for (int64_t i = 0; i + 1 < n; ++i) s[i] = (int64_t)T.t[i] + 1;
    s[n - 1] = 0;
    oracle::suffix_array(s.data(), SA.data(), n, 258);
    oracle::lcp_kasai(s.data(), SA.data(), LCP.data(), n);

    FILE *fsa = nullptr, *flcp = nullptr, *fbwt = nullptr;
    if (p.arrays_out) { fsa = fopen((p.prefix + ".sa").c_str(), "wb"); flcp = fopen((p.prefix + ".lcp").c_str(), "wb"); fbwt = fopen((p.prefix + ".bwt").c_str(), "wb"); }
    Detector D(p, T);
    for (int64_t j = 0; j < n; ++j) {
        u64 pos = (u64)SA[j];
        uint8_t ch = pos ? T.t[pos - 1] : 0;
        u64 d = std::upper_bound(T.doc_start.begin(), T.doc_start.end(), pos) - T.doc_start.begin() - 1;   // = rank of doc-end marks before pos
        if (fsa) { put5(fsa, pos); put5(flcp, (u64)LCP[j]); fputc(ch, fbwt); }
        D.push((u64)j, ch, d, pos, (u64)LCP[j]);
    }
    if (fsa) { fclose(fsa); fclose(flcp); fclose(fbwt); }
    D.finish();
    fprintf(stderr, "oracle: found %llu matches\n", (unsigned long long)D.count);
    return 0;
}

// ------------------------------------------------------------------------------------------------ anchor merge
struct Mum { uint32_t len; std::vector<int64_t> off; std::vector<uint8_t> plus; };
struct Partition { std::vector<Mum> mums; std::vector<uint8_t> starts; std::vector<uint16_t> nb; };

static Partition load_partition(const std::string& path) {
    if (path.size() < 5 || path.compare(path.size() - 5, 5, ".mums")) die("merge inputs must end with .mums: " + path);
    Partition P;
    { std::ifstream t(path.substr(0, path.size() - 5) + ".athresh", std::ios::binary | std::ios::ate);
      if (!t) die("missing .athresh for " + path);
      size_t bytes = t.tellg(); t.seekg(0); P.nb.resize(bytes / 2); t.read((char*)P.nb.data(), bytes); }
    P.starts.assign(P.nb.size(), 0);
    std::ifstream in(path); std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        std::istringstream ss(line); std::string a, b, c; ss >> a >> b >> c;
        Mum m; m.len = (uint32_t)std::stoul(a);
        { std::stringstream x(b); std::string tok; while (std::getline(x, tok, ',')) { if (tok.empty()) die("partial MUMs cannot be merged"); m.off.push_back(std::stoll(tok)); } }
        { std::stringstream x(c); std::string tok; while (std::getline(x, tok, ',')) m.plus.push_back(tok == "+"); }
        P.starts[m.off[0]] = 1; P.mums.push_back(std::move(m));
    }
    std::sort(P.mums.begin(), P.mums.end(), [](const Mum& a, const Mum& b) { return a.off[0] < b.off[0]; });
    return P;
}

static Partition fold(const Partition& A, const Partition& B) {
    Partition R; R.nb.assign(A.nb.size(), 0); R.starts.assign(A.nb.size(), 0);
    const Mum *ca = nullptr, *cb = nullptr; size_t ia = 0, ib = 0, la = 0, lb = 0;
    for (size_t i = 0; i < A.starts.size(); ++i) {
        bool both = A.nb[i] > 0 && B.nb[i] > 0;
        if (both) R.nb[i] = std::max(A.nb[i], B.nb[i]);
        if (A.starts[i]) { ca = &A.mums[ia++]; la = i; }
        if (B.starts[i]) { cb = &B.mums[ib++]; lb = i; }
        if (!(ca && cb && (A.starts[i] || B.starts[i]) && both)) continue;
        size_t da = i - la, db = i - lb;
        if (da > ca->len || db > cb->len) continue;
        uint32_t ra = ca->len - (uint32_t)da, rb = cb->len - (uint32_t)db, nl = std::min(ra, rb);
        if (!(nl > R.nb[i] && nl >= 20)) continue;
        Mum m; m.len = nl;
        auto shift = [&](const Mum& c, size_t delta, uint32_t rem, size_t from) {
            for (size_t d = from; d < c.off.size(); ++d) {
                m.off.push_back(c.off[d] + (c.plus[d] ? (int64_t)delta : (int64_t)rem - (int64_t)nl));
                m.plus.push_back(c.plus[d]);
            }
        };
        shift(*ca, da, ra, 0); shift(*cb, db, rb, 1);
        R.starts[m.off[0]] = 1; R.mums.push_back(std::move(m));
    }
    return R;
}

static int merge_main(int argc, char** argv) {
    std::vector<std::string> in; std::string out = "merged";
    for (int i = 2; i < argc; ++i) { std::string a = argv[i]; if (a == "-o" && i + 1 < argc) out = argv[++i]; else if (a != "-v") in.push_back(a); }
    if (in.size() < 2) die("merge needs at least two inputs");
    Partition acc = load_partition(in[0]);
    for (size_t i = 1; i < in.size(); ++i) acc = fold(acc, load_partition(in[i]));
    if (out.size() >= 5 && out.compare(out.size() - 5, 5, ".mums") == 0) out.resize(out.size() - 5);
    std::ofstream o(out + ".mums");
    for (auto& m : acc.mums) {
        o << m.len << "\t";
        for (size_t d = 0; d < m.off.size(); ++d) o << m.off[d] << (d + 1 < m.off.size() ? "," : "");
        o << "\t";
        for (size_t d = 0; d < m.plus.size(); ++d) o << (m.plus[d] ? "+" : "-") << (d + 1 < m.plus.size() ? "," : "");
        o << "\n";
    }
    std::ofstream t(out + ".athresh", std::ios::binary); t.write((char*)acc.nb.data(), acc.nb.size() * 2);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "merge")) return merge_main(argc, argv);
    return find_main(argc, argv);
}
