// rcx_kmercount — exact canonical k-mer counter that writes a Jellyfish-style
// text dump (">COUNT\nKMER\n", counts >= --min only).
//
// Test/bench infrastructure only: it stands in for stages 0-2 of the
// reference pipeline (run_rcorrector.pl:262-281: jellyfish bc / count / dump
// -L 2), which need Jellyfish 2.3.0 — not vendored in the reference and not
// installed in this image.  Stage 3 (the hot path) only needs the text file
// that main.cpp:294-308 parses.
//
// usage: rcx_kmercount -k K [-min 2] [-bin prefix] -o dump.txt reads1.fq [reads2.fq ...]
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static inline int base_code(char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
    }
}

static inline uint64_t revcomp(uint64_t x, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; ++i) {
        r = (r << 2) | (3 - (x & 3));
        x >>= 2;
    }
    return r;
}

int main(int argc, char** argv) {
    int k = 23, minc = 2;
    const char* out = nullptr;
    const char* binp = nullptr;
    std::vector<const char*> files;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-k")) k = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-min")) minc = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-o")) out = argv[++i];
        else if (!strcmp(argv[i], "-bin")) binp = argv[++i];
        else files.push_back(argv[i]);
    }
    if (!out || files.empty() || k < 1 || k > 32) {
        fprintf(stderr, "usage: rcx_kmercount -k K [-min 2] [-bin prefix] -o dump.txt reads.fq ...\n");
        return 1;
    }
    const uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    std::vector<uint64_t> all;
    std::vector<char> line(1 << 16);
    for (const char* f : files) {
        FILE* fp = fopen(f, "r");
        if (!fp) { fprintf(stderr, "cannot open %s\n", f); return 1; }
        long ln = 0;
        bool fastq = true;
        while (fgets(line.data(), (int)line.size(), fp)) {
            if (ln == 0) fastq = line[0] == '@';
            const int per = fastq ? 4 : 2;
            if (ln % per == 1) {
                uint64_t code = 0;
                int valid = 0;
                for (char* p = line.data(); *p && *p != '\n' && *p != '\r'; ++p) {
                    int b = base_code(*p);
                    if (b < 0) { valid = 0; code = 0; continue; }
                    code = ((code << 2) | (uint64_t)b) & mask;
                    if (++valid >= k) {
                        uint64_t rc = revcomp(code, k);
                        all.push_back(rc < code ? rc : code);
                    }
                }
            }
            ++ln;
        }
        fclose(fp);
    }
    std::sort(all.begin(), all.end());
    FILE* fo = fopen(out, "w");
    if (!fo) { fprintf(stderr, "cannot write %s\n", out); return 1; }
    std::vector<uint64_t> codes;
    std::vector<int32_t> counts;
    std::string kmer(k, 'A');
    size_t kept = 0;
    for (size_t i = 0; i < all.size();) {
        size_t j = i;
        while (j < all.size() && all[j] == all[i]) ++j;
        long c = (long)(j - i);
        if (c >= minc) {
            uint64_t x = all[i];
            for (int p = k - 1; p >= 0; --p) { kmer[p] = "ACGT"[x & 3]; x >>= 2; }
            fprintf(fo, ">%ld\n%s\n", c, kmer.c_str());
            if (binp) { codes.push_back(all[i]); counts.push_back((int32_t)c); }
            ++kept;
        }
        i = j;
    }
    fclose(fo);
    if (binp) {
        std::string a = std::string(binp) + ".codes", b = std::string(binp) + ".counts";
        FILE* fa = fopen(a.c_str(), "wb"); FILE* fb = fopen(b.c_str(), "wb");
        if (!fa || !fb) { fprintf(stderr, "cannot write %s.*\n", binp); return 1; }
        fwrite(codes.data(), 8, codes.size(), fa); fwrite(counts.data(), 4, counts.size(), fb);
        fclose(fa); fclose(fb);
    }
    fprintf(stderr, "rcx_kmercount: %zu k-mer instances, %zu distinct kept (count >= %d)\n", all.size(), kept, minc);
    return 0;
}
