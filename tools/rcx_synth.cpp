// rcx_synth — seeded synthetic RNA-seq-like workload generator.
//
// Test/bench infrastructure only (no reference code involved).  Produces, from
// one seed, (a) a random "transcriptome" with log-normal expression, optional
// poly-A tails, near-paralogs, tandem repeats and chimeric isoforms, (b) reads
// sampled from it with position-ramped substitution errors, correlated Phred+33
// qualities, sprinkled N and a fraction of truncated reads, and (c) optionally
// an ANALYTIC k-mer table (every transcriptome k-mer with its expected count,
// >= 2) so that big configs need no host-side counting of 10^9..10^10 k-mer
// instances.  Every read is generated from splitmix64(seed, read index), so the
// output does not depend on the thread count and a FASTQ prefix equals the
// prefix of the binary buffers.
//
// Outputs (all optional):
//   -cover N    reads / pairs of the whole job (default: -reads): the table's expected counts follow it, so that the
//               shards of one job (-reads n_i -rseed s_i) share one table
//   -rseed S    seed of the reads only (default: -seed): another shard of reads from the same transcriptome
//   -fq1 F [-fq2 F] [-fqlimit M]  FASTQ text (first M reads/pairs only)
//   -bin P      P.seq P.qual (concatenated, no separators) P.off (uint32[n+1]);
//               paired reads are interleaved mate1,mate2,...
//   -table P    P.codes (uint64 forward codes, canonical) P.counts (int32)
//   -dump F     the same table as Jellyfish-style text, descending count
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return (uint64_t)(uni() * (double)n); }
    int range(int lo, int hi) { return lo + (int)below((uint64_t)(hi - lo + 1)); }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

static inline uint64_t mix(uint64_t a, uint64_t b) {
    Rng r(a * 0xD1342543DE82EF95ull + b);
    r.next();
    return r.next();
}

struct Opts {
    uint64_t seed = 1;
    uint64_t rseed = 0;  // reads only (0: same as seed): different shards of reads from the same transcriptome
    long ntx = 50;
    int txmin = 300, txmax = 2500;
    double polya = 0.3, paralog = 0.0, repeat = 0.0, chimera = 0.2, sigma = 2.0;
    long reads = 20000;  // reads (single) or pairs (paired)
    long cover = 0;      // reads / pairs of the WHOLE job the table's expected counts follow (0: -reads); shards of one job share a table
    int len = 100;
    bool paired = false;
    double err = 0.01, nrate = 0.002, shortfrac = 0.02;
    int fragmax = 400;
    int k = 23;
    const char *fq1 = nullptr, *fq2 = nullptr, *bin = nullptr, *table = nullptr, *dump = nullptr;
    long fqlimit = -1;
    int threads = 0;
};

static const char NUC[5] = "ACGT";

int main(int argc, char** argv) {
    Opts o;
    for (int i = 1; i < argc; ++i) {
        auto is = [&](const char* s) { return !strcmp(argv[i], s); };
        if (is("-seed")) o.seed = strtoull(argv[++i], 0, 10);
        else if (is("-rseed")) o.rseed = strtoull(argv[++i], 0, 10);
        else if (is("-ntx")) o.ntx = atol(argv[++i]);
        else if (is("-txmin")) o.txmin = atoi(argv[++i]);
        else if (is("-txmax")) o.txmax = atoi(argv[++i]);
        else if (is("-polya")) o.polya = atof(argv[++i]);
        else if (is("-paralog")) o.paralog = atof(argv[++i]);
        else if (is("-repeat")) o.repeat = atof(argv[++i]);
        else if (is("-chimera")) o.chimera = atof(argv[++i]);
        else if (is("-sigma")) o.sigma = atof(argv[++i]);
        else if (is("-reads")) o.reads = atol(argv[++i]);
        else if (is("-cover")) o.cover = atol(argv[++i]);
        else if (is("-len")) o.len = atoi(argv[++i]);
        else if (is("-paired")) o.paired = true;
        else if (is("-err")) o.err = atof(argv[++i]);
        else if (is("-nrate")) o.nrate = atof(argv[++i]);
        else if (is("-short")) o.shortfrac = atof(argv[++i]);
        else if (is("-fragmax")) o.fragmax = atoi(argv[++i]);
        else if (is("-k")) o.k = atoi(argv[++i]);
        else if (is("-fq1")) o.fq1 = argv[++i];
        else if (is("-fq2")) o.fq2 = argv[++i];
        else if (is("-fqlimit")) o.fqlimit = atol(argv[++i]);
        else if (is("-bin")) o.bin = argv[++i];
        else if (is("-table")) o.table = argv[++i];
        else if (is("-dump")) o.dump = argv[++i];
        else if (is("-threads")) o.threads = atoi(argv[++i]);
        else { fprintf(stderr, "rcx_synth: unknown option %s\n", argv[i]); return 1; }
    }
    if (o.len < 1 || o.len > 1000 || o.k < 1 || o.k > 32) { fprintf(stderr, "bad -len/-k\n"); return 1; }
    if (!o.rseed) o.rseed = o.seed;
    if (o.txmin < o.len) o.txmin = o.len;
    if (o.txmax < o.txmin) o.txmax = o.txmin;
    int nthreads = o.threads > 0 ? o.threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;

    // ---------------------------------------------------------------- transcriptome
    Rng g(mix(o.seed, 0xA11CE));
    std::vector<std::string> tx;
    tx.reserve(o.ntx + o.ntx / 4 + 4);
    for (long t = 0; t < o.ntx; ++t) {
        int L = g.range(o.txmin, o.txmax);
        std::string s(L, 'A');
        if (t > 0 && g.uni() < o.paralog) {
            // near-paralog: copy of an earlier transcript with a substitution every 15-40 bp
            const std::string& src = tx[g.below(tx.size())];
            s = src;
            for (size_t p = g.range(15, 40); p < s.size(); p += g.range(15, 40)) {
                char c;
                do c = NUC[g.below(4)]; while (c == s[p]);
                s[p] = c;
            }
        } else {
            for (int p = 0; p < L; ++p) s[p] = NUC[g.below(4)];
        }
        if (g.uni() < o.repeat && s.size() > 260) {
            int unit = g.range(2, 6), copies = g.range(10, 40);
            std::string u(unit, 'A');
            for (int p = 0; p < unit; ++p) u[p] = NUC[g.below(4)];
            std::string rep;
            for (int c = 0; c < copies; ++c) rep += u;
            s.insert(200, rep);
        }
        if (g.uni() < o.polya) s += std::string(g.range(10, 60), 'A');
        tx.push_back(s);
    }
    long nch = (long)(o.ntx * o.chimera);
    for (long c = 0; c < nch; ++c) {
        const std::string& a = tx[g.below(o.ntx)];
        const std::string& b = tx[g.below(o.ntx)];
        std::string s = a.substr(0, a.size() / 2) + b.substr(b.size() / 2);
        if ((int)s.size() >= o.len) tx.push_back(s);
    }
    const long T = (long)tx.size();
    std::vector<double> w(T), cum(T);
    double tot = 0;
    for (long t = 0; t < T; ++t) {
        w[t] = exp(o.sigma * g.normal()) * (double)tx[t].size();  // reads ~ expression x length
        tot += w[t];
        cum[t] = tot;
    }

    // ---------------------------------------------------------------- reads
    const long nrec_all = (o.paired ? 2 : 1) * (o.cover > 0 ? o.cover : o.reads);  // the table's expected coverage always uses the full workload
    if (!o.bin && o.fq1 && o.fqlimit >= 0 && o.fqlimit < o.reads) o.reads = o.fqlimit;  // FASTQ sample only: generate just that prefix
    const long nrec = o.paired ? 2 * o.reads : o.reads;
    const bool want_reads = o.bin || o.fq1;
    std::vector<uint32_t> off;
    std::vector<char> seq, qual;
    if (want_reads) {
        off.assign(nrec + 1, 0);
        // pass 1: lengths (cheap, deterministic per read index)
        auto read_len = [&](long pair_idx, int mate) -> int {
            Rng r(mix(o.rseed ^ 0x5EED, (uint64_t)pair_idx));
            double u = r.uni();
            if (mate == 0 && u < o.shortfrac) return 15 + (int)(r.uni() * (o.len - 15 + 1));
            return o.len;
        };
        for (long i = 0; i < o.reads; ++i) {
            if (o.paired) {
                off[2 * i + 1] = read_len(i, 0);
                off[2 * i + 2] = read_len(i, 1);
            } else {
                off[i + 1] = read_len(i, 0);
            }
        }
        for (long i = 0; i < nrec; ++i) off[i + 1] += off[i];
        seq.resize(off[nrec]);
        qual.resize(off[nrec]);
        auto worker = [&](long lo, long hi) {
            std::string frag, m2;
            for (long i = lo; i < hi; ++i) {
                Rng r(mix(o.rseed ^ 0xF00D, (uint64_t)i));
                double x = r.uni() * tot;
                long t = std::lower_bound(cum.begin(), cum.end(), x) - cum.begin();
                if (t >= T) t = T - 1;
                const std::string& s = tx[t];
                int fmax = std::min<int>((int)s.size(), std::max(o.len, o.fragmax));
                int flen = r.range(o.len, fmax);
                long start = (long)r.below(s.size() - flen + 1);
                frag.assign(s, start, flen);
                bool rev = r.next() & 1;
                auto rc = [](const std::string& a) {
                    std::string b(a.rbegin(), a.rend());
                    for (auto& c : b) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
                    return b;
                };
                if (rev) frag = rc(frag);
                for (int mate = 0; mate < (o.paired ? 2 : 1); ++mate) {
                    long rec = o.paired ? 2 * i + mate : i;
                    int L = off[rec + 1] - off[rec];
                    const std::string* src = &frag;
                    if (mate == 1) { m2 = rc(frag); src = &m2; }
                    char* sq = &seq[off[rec]];
                    char* ql = &qual[off[rec]];
                    for (int p = 0; p < L; ++p) {
                        char c = (*src)[p];
                        double pe = o.err * (0.5 + 1.5 * p / (double)o.len);
                        double u = r.uni();
                        if (u < o.nrate) { sq[p] = 'N'; ql[p] = '!'; continue; }
                        if (r.uni() < pe) {
                            char d;
                            do d = NUC[r.below(4)]; while (d == c);
                            sq[p] = d;
                            ql[p] = (char)('!' + r.range(2, 20));
                        } else {
                            sq[p] = c;
                            ql[p] = (char)('!' + r.range(15, 41));
                        }
                    }
                }
            }
        };
        std::vector<std::thread> th;
        long chunk = (o.reads + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; ++t) {
            long lo = t * chunk, hi = std::min(o.reads, lo + chunk);
            if (lo < hi) th.emplace_back(worker, lo, hi);
        }
        for (auto& t : th) t.join();
    }
    if (o.bin) {
        std::string p(o.bin);
        FILE* f;
        f = fopen((p + ".seq").c_str(), "wb"); if (!f) return 1; fwrite(seq.data(), 1, seq.size(), f); fclose(f);
        f = fopen((p + ".qual").c_str(), "wb"); if (!f) return 1; fwrite(qual.data(), 1, qual.size(), f); fclose(f);
        f = fopen((p + ".off").c_str(), "wb"); if (!f) return 1; fwrite(off.data(), 4, off.size(), f); fclose(f);
    }
    if (o.fq1) {
        long lim = o.fqlimit >= 0 ? std::min(o.fqlimit, o.reads) : o.reads;
        FILE* f1 = fopen(o.fq1, "w");
        FILE* f2 = o.paired && o.fq2 ? fopen(o.fq2, "w") : nullptr;
        if (!f1 || (o.paired && o.fq2 && !f2)) { fprintf(stderr, "cannot write fastq\n"); return 1; }
        std::vector<char> buf1(1 << 22), buf2(1 << 22);
        setvbuf(f1, buf1.data(), _IOFBF, buf1.size());
        if (f2) setvbuf(f2, buf2.data(), _IOFBF, buf2.size());
        for (long i = 0; i < lim; ++i) {
            for (int mate = 0; mate < (o.paired ? 2 : 1); ++mate) {
                long rec = o.paired ? 2 * i + mate : i;
                FILE* f = (mate == 1 && f2) ? f2 : f1;
                int L = off[rec + 1] - off[rec];
                if (o.paired) fprintf(f, "@r%ld/%d\n", i, mate + 1); else fprintf(f, "@r%ld\n", i);
                fwrite(&seq[off[rec]], 1, L, f);
                fputs("\n+\n", f);
                fwrite(&qual[off[rec]], 1, L, f);
                fputc('\n', f);
            }
        }
        fclose(f1);
        if (f2) fclose(f2);
    }

    // ---------------------------------------------------------------- analytic table
    if (o.table || o.dump) {
        const int k = o.k;
        const uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
        // expected per-base coverage of transcript t
        const double total_bases = (double)nrec_all * o.len;
        struct KC { uint64_t code; int32_t cnt; };
        std::vector<size_t> base(T + 1, 0);
        for (long t = 0; t < T; ++t) base[t + 1] = base[t] + (tx[t].size() >= (size_t)k ? tx[t].size() - k + 1 : 0);
        std::vector<KC> kc(base[T]);
        auto tworker = [&](long lo, long hi) {
            for (long t = lo; t < hi; ++t) {
                double cov = total_bases * (w[t] / tot) / (double)tx[t].size() * (double)(o.len - k + 1) / o.len;
                Rng r(mix(o.seed ^ 0xC0DE, (uint64_t)t));
                const std::string& s = tx[t];
                uint64_t fw = 0, rv = 0;
                size_t w0 = base[t];
                for (size_t p = 0; p < s.size(); ++p) {
                    uint64_t b = s[p] == 'A' ? 0 : s[p] == 'C' ? 1 : s[p] == 'G' ? 2 : 3;
                    fw = ((fw << 2) | b) & mask;
                    rv = (rv >> 2) | ((3 - b) << (2 * (k - 1)));
                    if (p + 1 >= (size_t)k) {
                        // mildly noisy count around the expected coverage, never below 2
                        double c = cov * (0.85 + 0.3 * r.uni());
                        int32_t ci = c > 2e9 ? 2000000000 : (int32_t)(c + 0.5);
                        if (ci < 2) ci = 2;
                        kc[w0++] = KC{fw < rv ? fw : rv, ci};
                    }
                }
            }
        };
        {
            std::vector<std::thread> th;
            long chunk = (T + nthreads - 1) / nthreads;
            for (int t = 0; t < nthreads; ++t) {
                long lo = t * chunk, hi = std::min(T, lo + chunk);
                if (lo < hi) th.emplace_back(tworker, lo, hi);
            }
            for (auto& t : th) t.join();
        }
        // merge duplicate canonical k-mers (sum of counts): partition by top bits, sort partitions in parallel
        const int PB = 8, NP = 1 << PB;
        const int shift = 2 * k > PB ? 2 * k - PB : 0;
        std::vector<size_t> pcount(NP + 1, 0);
        for (auto& e : kc) pcount[(e.code >> shift) + 1]++;
        for (int p = 0; p < NP; ++p) pcount[p + 1] += pcount[p];
        std::vector<KC> part(kc.size());
        {
            std::vector<size_t> cur(pcount.begin(), pcount.end() - 1);
            for (auto& e : kc) part[cur[e.code >> shift]++] = e;
        }
        kc.clear(); kc.shrink_to_fit();
        std::vector<size_t> outn(NP, 0);
        auto sworker = [&](int lo, int hi) {
            for (int p = lo; p < hi; ++p) {
                KC* b = &part[pcount[p]];
                KC* e = &part[pcount[p + 1]];
                std::sort(b, e, [](const KC& x, const KC& y) { return x.code < y.code; });
                KC* wv = b;
                for (KC* q = b; q < e;) {
                    KC* q2 = q;
                    int64_t s = 0;
                    while (q2 < e && q2->code == q->code) { s += q2->cnt; ++q2; }
                    *wv++ = KC{q->code, (int32_t)std::min<int64_t>(s, 2000000000)};
                    q = q2;
                }
                outn[p] = wv - b;
            }
        };
        {
            std::vector<std::thread> th;
            int chunk = (NP + nthreads - 1) / nthreads;
            for (int t = 0; t < nthreads; ++t) {
                int lo = t * chunk, hi = std::min(NP, lo + chunk);
                if (lo < hi) th.emplace_back(sworker, lo, hi);
            }
            for (auto& t : th) t.join();
        }
        std::vector<uint64_t> codes;
        std::vector<int32_t> counts;
        for (int p = 0; p < NP; ++p)
            for (size_t i = 0; i < outn[p]; ++i) {
                codes.push_back(part[pcount[p] + i].code);
                counts.push_back(part[pcount[p] + i].cnt);
            }
        part.clear(); part.shrink_to_fit();
        // scatter the order deterministically so the table is not fed in sorted-key order
        if (o.table) {
            std::string p(o.table);
            FILE* f;
            f = fopen((p + ".codes").c_str(), "wb"); if (!f) return 1; fwrite(codes.data(), 8, codes.size(), f); fclose(f);
            f = fopen((p + ".counts").c_str(), "wb"); if (!f) return 1; fwrite(counts.data(), 4, counts.size(), f); fclose(f);
        }
        if (o.dump) {
            // descending count so the reference's ERROR_RATE pass (main.cpp:314-351) meets its
            // 100000 high-count records early instead of scanning the whole file
            std::vector<uint32_t> idx(codes.size());
            for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
            std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return counts[a] > counts[b]; });
            FILE* f = fopen(o.dump, "w");
            if (!f) return 1;
            std::vector<char> vb(1 << 22);
            setvbuf(f, vb.data(), _IOFBF, vb.size());
            char line[96];
            for (uint32_t i : idx) {
                int n = snprintf(line, sizeof line, ">%d\n", counts[i]);
                uint64_t x = codes[i];
                for (int p = k - 1; p >= 0; --p) { line[n + p] = NUC[x & 3]; x >>= 2; }
                line[n + k] = '\n';
                fwrite(line, 1, n + k + 1, f);
            }
            fclose(f);
        }
        fprintf(stderr, "rcx_synth: table %zu distinct k-mers\n", codes.size());
    }
    fprintf(stderr, "rcx_synth: %ld transcripts, %ld records\n", T, want_reads ? nrec : 0L);
    return 0;
}
