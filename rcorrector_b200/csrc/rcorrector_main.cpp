// rcorrector_main.cpp — the drop-in `rcorrector` stage-3 binary (process surface of SURVEY §8b).
//
// Keeps the reference's command line (main.cpp:165-268), stderr lines
// (main.cpp:308,358,364,81-86), exit codes and output naming / bytes
// (Reads.h:138-157,360-421), so `run_rcorrector.pl ... -stage 3` can call it
// unmodified.  Everything between "load the dump" and "write the records" is
// replaced: the dump is parsed in bulk into (code, count) arrays, the table is
// built in HBM through the C ABI (rcx_table_build), and where the reference
// creates/joins its pthreads per 512*t reads (main.cpp:476-483) this binary
// hands multi-million-read batches to rcx_correct_batch — one host thread per
// GPU, reads sharded by record range (mates stay together), output in input
// order.  There is no CPU correction path: without a GPU it exits 1.
//
// Extra options (not in the reference): -gpus N (default 1; env RCX_GPUS),
// -batch N (records per pipeline batch, default 1M), -countkmers (instead of
// -c: count the k-mers of the input files on the GPU, SURVEY 8f N3 -- what
// run_rcorrector.pl's stages 0-2 ask Jellyfish for).
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <future>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rcx.h"
#include "fastq_io.h"

using rcxio::Chunk;
using rcxio::LineReader;
using rcxio::Writer;

static void print_help() {  // main.cpp:50-71 (same text: scripts and users read it)
    fprintf(stderr,
            "Usage: ./rcorrector [OPTIONS]\n"
            "OPTIONS:\n"
            "Required parameters:\n"
            "\t-r seq_file: seq_file is the path to the sequence file. Can use multiple -r to specifiy multiple sequence files\n"
            "\t-p seq_file_left seq_file_right: the paths to the paired-end data set. Can use multiple -p to specifiy multiple sequence files\n"
            "\t-i seq_file: seq_file is the path to the interleaved mate-pair sequence file. Can use multiple -i\n"
            "\t-c jf_dump: the kmer counts dumped by JellyFish\n"
            "\t-k kmer_length\n"
            "Other parameters:\n"
            "\t-od output_file_directory (default: ./)\n"
            "\t-t number of threads to use (default: 1)\n"
            "\t-maxcor INT: the maximum number of correction every 100bp (default: 8)\n"
            "\t-maxcorK INT: the maximum number of correction within k-bp window (default: 4)\n"
            "\t-wk FLOAT: the proportion of kmers that are used to estimate weak kmer count threshold (default: 0.95)\n"
            "\t-stdout: output the corrected sequences to stdout (default: not used)\n"
            "\t-verbose: output some correction information to stdout (default: not used)\n");
}

static double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static const double g_t0 = now_s();
static const bool g_timing = getenv("RCX_TIMING") != nullptr;
// RCX_TIMING: seconds spent per pipeline stage, summed over the batches (the stages of different batches overlap)
static std::atomic<long long> g_us_read{0}, g_us_correct{0}, g_us_format{0}, g_us_write{0};
struct StageTimer {
    std::atomic<long long>& acc;
    double t0;
    explicit StageTimer(std::atomic<long long>& a) : acc(a), t0(g_timing ? now_s() : 0) {}
    ~StageTimer() { if (g_timing) acc += (long long)((now_s() - t0) * 1e6); }
};
static void tmark(const char* what) {
    if (g_timing) fprintf(stderr, "RCX_TIMING %8.3f s  %s\n", now_s() - g_t0, what);
}

// outputs opened so far: removed again when the run has to stop half way (a truncated *.cor.fq looks like a result)
static std::vector<std::string> g_out_paths;
static std::string g_od_created;
static void remove_partial_outputs() {
    for (auto& p : g_out_paths) unlink(p.c_str());
    if (!g_od_created.empty()) rmdir(g_od_created.c_str());  // (only if this run created it and it is empty)
}

static void die_rcx(const char* what) {
    fprintf(stderr, "%s: %s\n", what, rcx_last_error());
    exit(1);
}

struct ReadFile {
    std::string path;
    std::string mate_path;  // second mate for -p
    bool paired = false, interleaved = false;
    int type = 1;  // 0 FASTA, 1 FASTQ (Reads.h:118-132)
    std::string out_path, mate_out_path;
};

static int sniff_type(const std::string& path) {  // Reads.h:108-132
    LineReader r;
    if (!r.open(path)) {
        fprintf(stderr, "ERROR: Could not access file %s\n", path.c_str());  // File.h:71,93
        exit(1);
    }
    const char* p;
    size_t n;
    bool nl;
    const bool got = r.getline(p, n, nl);
    if (!got || n == 0 || (p[0] != '>' && p[0] != '@')) {
        // Reads.h:118-131 prints the buffer its Gets filled: the first line WITH its newline (an empty file leaves the
        // reference's buffer uninitialised; nothing is shown here)
        const std::string shown = got ? std::string(p, n) + (nl ? "\n" : "") : std::string();
        fprintf(stderr, "\"%s\"'s format is wrong: %s\n", path.c_str(), shown.c_str());
        exit(1);
    }
    return p[0] == '@' ? 1 : 0;
}

static std::string out_name(const std::string& od, const std::string& in, int type) { return rcxio::output_path(od, in, type); }

// ---------------------------------------------------------------------------------------------- dump loading
struct Dump {
    std::vector<uint64_t> codes;   // every record, in file order (KmerCode::Append over the k-mer token)
    std::vector<int32_t> counts;   // atoi(&token[1]); records with an invalid k-mer get count 0 (Store::Put ignores them)
    int64_t stored = 0;            // the "Stored %d kmers" number (main.cpp:306)
    uint64_t rate_records = 0;     // how many leading records the ERROR_RATE pass may use (validity quirk, main.cpp:323)
};

static inline int nuc(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

static void load_dump_serial(const char* path, int k, Dump& d) {  // main.cpp:294-308 with a block reader instead of fscanf("%s")
    FILE* fp = fopen(path, "r");
    if (!fp) {
        fprintf(stderr, "Could not open file %s\n", path);
        exit(1);
    }
    const uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    std::vector<char> buf(64u << 20);
    size_t have = 0;
    bool eof = false, want_count = true, rate_open = true;
    int last_inv_state = -1;  // KmerCode::invalidPos left by the last STORED record of this loop
    int cur_count = 0;
    while (!eof || have) {
        if (!eof) {
            size_t got = fread(buf.data() + have, 1, buf.size() - have, fp);
            if (got == 0) eof = true;
            have += got;
        }
        size_t pos = 0, consumed = 0;
        for (;;) {
            while (pos < have && (unsigned char)buf[pos] <= ' ') ++pos;  // %s skips whitespace
            size_t s = pos;
            while (pos < have && (unsigned char)buf[pos] > ' ') ++pos;
            if (pos == have && !eof) break;  // token may continue in the next block
            if (s == pos) { consumed = have; break; }
            if (want_count) {
                // atoi(&buffer[1])
                long v = 0;
                size_t q = s + 1;
                bool neg = false;
                if (q < pos && (buf[q] == '-' || buf[q] == '+')) { neg = buf[q] == '-'; ++q; }
                while (q < pos && buf[q] >= '0' && buf[q] <= '9') { v = v * 10 + (buf[q] - '0'); if (v > 2147483647L) v = 2147483647L; ++q; }
                cur_count = (int)(neg ? -v : v);
                want_count = false;
            } else {
                uint64_t code = 0;
                int inv = -1;
                for (size_t q = s; q < pos; ++q) {  // KmerCode::Append, KmerCode.cpp:7-21
                    const int v = nuc(buf[q]);
                    if (inv != -1) ++inv;
                    code = ((code << 2) & mask) | (uint64_t)(v & 3);
                    if (v < 0) inv = 0;
                    if (inv >= k) inv = -1;
                }
                d.codes.push_back(code);
                if (cur_count > 1) {
                    ++d.stored;
                    last_inv_state = inv;
                    d.counts.push_back(inv == -1 ? cur_count : 0);
                } else {
                    d.counts.push_back(0);
                }
                // the ERROR_RATE pass stops being able to use records once one of them leaves the shared KmerCode
                // invalid after the 4 last-base variants (an invalid character anywhere but the last position)
                if (rate_open) {
                    if (inv > 0) rate_open = false;
                    else d.rate_records = d.codes.size();
                }
                want_count = true;
            }
            consumed = pos;
        }
        memmove(buf.data(), buf.data() + consumed, have - consumed);
        have -= consumed;
        if (eof && have && consumed == 0) break;
    }
    fclose(fp);
    if (last_inv_state != -1) d.rate_records = 0;  // main.cpp:323 tests the state the load loop left behind
}


// The same loader for a regular dump (">count\nKMER\n" records, which is what `jellyfish dump` writes): the file is mapped,
// cut at record starts and parsed by `threads` host threads.  Anything irregular (a count token that does not start with
// '>', a k-mer token that does, blanks inside a line, a missing final newline) falls back to the serial token parser.
static bool load_dump_parallel(const char* path, int k, Dump& d, int threads) {
    rcxio::MappedFile mf;
    if (!mf.open(path) || mf.size < 4 || mf.data[0] != '>' || mf.data[mf.size - 1] != '\n') return false;
    const uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    const int T = std::max(1, threads);
    struct Part {
        std::vector<uint64_t> codes;
        std::vector<int32_t> counts;
        int64_t stored = 0;
        int64_t first_bad = -1;  // first local record whose k-mer leaves KmerCode invalid with invalidPos > 0
        int last_inv = -2;       // invalidPos after the last STORED record (-2: none stored here)
        bool ok = true;
    };
    std::vector<Part> part(T);
    std::vector<size_t> cut(T + 1, 0);
    cut[T] = mf.size;
    for (int t = 1; t < T; ++t) {  // a record starts at a '>' that follows a newline
        size_t p = mf.size * t / T;
        while (p < mf.size && !(mf.data[p] == '>' && mf.data[p - 1] == '\n')) ++p;
        cut[t] = p;
    }
    rcxio::parallel_for(T, [&](int t) {
        Part& P = part[t];
        const char* p = mf.data + cut[t];
        const char* e = mf.data + cut[t + 1];
        P.codes.reserve((size_t)(e - p) / (k + 4) + 16);
        P.counts.reserve((size_t)(e - p) / (k + 4) + 16);
        while (p < e) {
            if (*p != '>') { P.ok = false; return; }
            const char* q = (const char*)memchr(p, '\n', e - p);
            if (!q) { P.ok = false; return; }
            long v = 0;  // atoi(&buffer[1])
            const char* c = p + 1;
            bool neg = false;
            if (c < q && (*c == '-' || *c == '+')) { neg = *c == '-'; ++c; }
            const char* c0 = c;
            while (c < q && *c >= '0' && *c <= '9') { v = v * 10 + (*c - '0'); if (v > 2147483647L) v = 2147483647L; ++c; }
            if (c != q || c == c0) { P.ok = false; return; }  // blanks / other characters in the count line: let the token parser decide
            const int cur_count = (int)(neg ? -v : v);
            const char* s0 = q + 1;
            const char* q2 = (const char*)memchr(s0, '\n', e - s0);
            if (!q2 || q2 == s0) { P.ok = false; return; }
            uint64_t code = 0;
            int inv = -1;
            for (const char* x = s0; x < q2; ++x) {  // KmerCode::Append, KmerCode.cpp:7-21
                const unsigned char ch = (unsigned char)*x;
                if (ch <= ' ' || ch == '>') { P.ok = false; return; }
                const int b = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1;
                if (inv != -1) ++inv;
                code = ((code << 2) & mask) | (uint64_t)(b & 3);
                if (b < 0) inv = 0;
                if (inv >= k) inv = -1;
            }
            P.codes.push_back(code);
            if (cur_count > 1) {
                ++P.stored;
                P.last_inv = inv;
                P.counts.push_back(inv == -1 ? cur_count : 0);
            } else {
                P.counts.push_back(0);
            }
            if (P.first_bad < 0 && inv > 0) P.first_bad = (int64_t)P.codes.size() - 1;
            p = q2 + 1;
        }
    });
    size_t total = 0;
    for (auto& P : part) { if (!P.ok) return false; total += P.codes.size(); }
    d.codes.resize(total);
    d.counts.resize(total);
    std::vector<size_t> start(T + 1, 0);
    for (int t = 0; t < T; ++t) start[t + 1] = start[t] + part[t].codes.size();
    rcxio::parallel_for(T, [&](int t) {
        if (part[t].codes.empty()) return;
        memcpy(d.codes.data() + start[t], part[t].codes.data(), part[t].codes.size() * sizeof(uint64_t));
        memcpy(d.counts.data() + start[t], part[t].counts.data(), part[t].counts.size() * sizeof(int32_t));
    });
    d.stored = 0;
    d.rate_records = total;
    int last_inv_state = -1;
    bool rate_open = true;
    for (int t = 0; t < T; ++t) {
        d.stored += part[t].stored;
        if (part[t].last_inv != -2) last_inv_state = part[t].last_inv;
        if (rate_open && part[t].first_bad >= 0) { d.rate_records = start[t] + (size_t)part[t].first_bad; rate_open = false; }
    }
    if (last_inv_state != -1) d.rate_records = 0;  // main.cpp:323 tests the state the load loop left behind
    return true;
}

static void load_dump(const char* path, int k, Dump& d, int threads) {
    if (!getenv("RCX_SERIAL_IO") && load_dump_parallel(path, k, d, threads)) return;
    d = Dump();
    load_dump_serial(path, k, d);
}

// ---------------------------------------------------------------------------------------------- per-file processing
struct Gpu {
    rcx_ctx* ctx = nullptr;
};

struct HostBatch {
    char* seq = nullptr;
    char* qual = nullptr;
    uint32_t* off = nullptr;
    int32_t *corr = nullptr, *l = nullptr, *m = nullptr, *h = nullptr;
    size_t cap_bases = 0, cap_reads = 0;
    size_t n = 0, bases = 0;
    void alloc(size_t reads, size_t bases_) {
        cap_reads = reads;
        cap_bases = bases_;
        seq = (char*)rcx_host_alloc(bases_ + 64);
        qual = (char*)rcx_host_alloc(bases_ + 64);
        off = (uint32_t*)rcx_host_alloc((reads + 1) * 4);
        corr = (int32_t*)rcx_host_alloc(reads * 4);
        l = (int32_t*)rcx_host_alloc(reads * 4);
        m = (int32_t*)rcx_host_alloc(reads * 4);
        h = (int32_t*)rcx_host_alloc(reads * 4);
        if (!seq || !qual || !off || !corr || !l || !m || !h) die_rcx("pinned allocation failed");
    }
};

// read one record into the batch; returns false at end of file
static bool read_record(LineReader& r, int type, HostBatch& b, Chunk& c, const std::string& path) {
    const char* p;
    size_t n;
    bool nl;
    if (!r.getline(p, n, nl)) return false;
    c.ids.insert(c.ids.end(), p, p + n);
    c.id_off.push_back((uint32_t)c.ids.size());
    size_t L = 0;
    bool seq_nl = false;
    if (r.getline(p, n, nl)) {
        L = n;
        seq_nl = nl;
        if (L > RCX_MAX_READ_LENGTH) {
            fprintf(stderr, "ERROR: a read in %s is longer than %d bases.\n", path.c_str(), RCX_MAX_READ_LENGTH);
            exit(1);
        }
        memcpy(b.seq + b.bases, p, L);
    }
    if (type == 1) {
        r.getline(p, n, nl);  // '+' line: dropped (Reads.h:246)
        size_t qn = 0;
        bool qnl = false;
        const char* qp = nullptr;
        if (r.getline(qp, qn, qnl)) {
            // Reads.h:255-262: the newline is stripped at the SEQUENCE line's length
            const size_t strip_at = L + (seq_nl ? 1 : 0) - 1;
            if (qn == L || (qn == strip_at && qnl)) {
                memcpy(b.qual + b.bases, qp, std::min(qn, L));
                if (qn < L) memset(b.qual + b.bases + qn, 0, L - qn);
            } else {
                c.ragged_qual = true;
                memcpy(b.qual + b.bases, qp, std::min(qn, L));
                if (qn < L) memset(b.qual + b.bases + qn, 0, L - qn);
            }
            if (c.ragged_qual || qn != L) {
                // keep verbatim text (with its newline if it was not stripped) for the writer
                while (c.qual_text_off.size() < c.id_off.size() - 1) c.qual_text_off.push_back((uint32_t)c.qual_text.size());
                c.qual_text.insert(c.qual_text.end(), qp, qp + qn);
                if (qnl && qn != strip_at) c.qual_text.push_back('\n');
                c.qual_text_off.push_back((uint32_t)c.qual_text.size());
                c.ragged_qual = true;
            }
        } else {
            memset(b.qual + b.bases, 0, L);
        }
    } else {
        memset(b.qual + b.bases, 0, L);  // FASTA: qual[0] == '\0' (Reads.h:241)
    }
    b.bases += L;
    b.off[++b.n] = (uint32_t)b.bases;
    c.n = b.n;
    return true;
}

// Reads.h:360-421 OutputBatch for records [lo, hi) of the batch (stride/start select one mate of an interleaved batch),
// written through a sink: a byte counter (the size of a piece is known before its place in the file is reserved), a raw
// pointer into the mapped output file, or a string (gzip members, streams).
struct CountSink {
    size_t n = 0;
    void put(const char*, size_t len) { n += len; }
    void ch(char) { ++n; }
};
struct PtrSink {
    char* p;
    void put(const char* s, size_t len) { memcpy(p, s, len); p += len; }
    void ch(char c) { *p++ = c; }
};
struct StringSink {
    std::string& out;
    void put(const char* s, size_t len) { out.append(s, len); }
    void ch(char c) { out.push_back(c); }
};
template <class Sink>
static void emit_records(Sink& out, const HostBatch& b, const Chunk& c, int type, size_t lo, size_t hi, size_t stride) {
    char num[16];
    for (size_t r = lo; r < hi; r += stride) {
        if (c.ext_ids) out.put(c.id_ptr[r], c.id_len[r]);
        else out.put(c.ids.data() + c.id_off[r], c.id_off[r + 1] - c.id_off[r]);
        out.put(" l:", 3);
        out.put(num, rcxio::format_int(num, b.l[r]));
        out.put(" m:", 3);
        out.put(num, rcxio::format_int(num, b.m[r]));
        out.put(" h:", 3);
        out.put(num, rcxio::format_int(num, b.h[r]));
        if (b.corr[r] == -1) out.put(" unfixable_error", 16);
        else if (b.corr[r] > 0) out.put(" cor", 4);
        out.ch('\n');
        out.put(b.seq + b.off[r], b.off[r + 1] - b.off[r]);
        out.ch('\n');
        if (type == 1) {
            out.put("+\n", 2);
            if (c.ragged_qual && r + 1 < c.qual_text_off.size() && c.qual_text_off[r + 1] > c.qual_text_off[r])
                out.put(c.qual_text.data() + c.qual_text_off[r], c.qual_text_off[r + 1] - c.qual_text_off[r]);
            else
                out.put(b.qual + b.off[r], b.off[r + 1] - b.off[r]);
            out.ch('\n');
        }
    }
}
static void format_records(const HostBatch& b, const Chunk& c, int type, size_t lo, size_t hi, size_t stride, std::string& out) {
    out.clear();
    if (lo < hi) {  // one allocation: ids + twice the bases + at most 64 bytes of tags and separators per record
        const size_t nrec = (hi - lo + stride - 1) / stride;
        const size_t bases = (size_t)(b.off[hi] - b.off[lo]);
        const size_t idb = c.ext_ids ? nrec * (size_t)(c.id_len[lo] + 16) : (size_t)(c.id_off[hi] - c.id_off[lo]);
        const size_t want = idb + 2 * bases + 64 * nrec;
        if (out.capacity() < want) out.reserve(want + want / 8);
    }
    StringSink sink{out};
    emit_records(sink, b, c, type, lo, hi, stride);
}


// ---------------------------------------------------------------------------------------------- parallel FASTQ reader
// One input file.  Plain FASTQ files are mapped and read by `threads` host threads (newline index, then every thread
// copies its share of records into the pinned batch); gzip / FASTA input, and any file in which a record turns out to be
// irregular (quality line of another length than the sequence line, missing final newline), is read record by record
// with the reference's line semantics (read_record).
struct Source {
    std::string path;
    int type = 1;
    bool fast = false;
    rcxio::MappedFile mf;
    size_t pos = 0;             // fast mode: next unread byte
    LineReader lr;              // serial mode
    std::vector<uint32_t> nl;   // newline index of the current region, relative to reg_lo
    size_t reg_lo = 0;
    double rec_bytes = 300;     // running estimate of bytes per record
    bool open(const std::string& p, int t, bool allow_fast) {
        path = p; type = t; pos = 0;
        fast = allow_fast && t == 1 && !rcxio::ends_with_gz(p) && mf.open(p) && mf.size > 0;
        if (fast) return true;
        return lr.open(p);
    }
    void go_serial() {  // hand the rest of a mapped file to the serial reader
        lr.open_mem(mf.data, mf.size, pos);
        fast = false;
    }
    // index enough of the file for up to `want` records; returns how many whole records are available
    size_t peek(size_t want, int threads) {
        size_t hi = pos + (size_t)(rec_bytes * 1.05 * (double)want) + 4096;
        for (;;) {
            if (hi > mf.size) hi = mf.size;
            if (hi - pos > 0xfff00000ull) hi = pos + 0xfff00000ull;  // 32-bit relative offsets
            reg_lo = pos;
            rcxio::index_newlines(mf.data, pos, hi, threads, nl);
            const size_t have = nl.size() / 4;
            if (have >= want || hi >= mf.size || hi - pos >= 0xfff00000ull) return have < want ? have : want;
            hi = pos + (hi - pos) * 2;
        }
    }
    bool at_eof() const { return pos >= mf.size; }
    // true if what is left cannot be taken by the parallel reader (bytes after the last newline)
    const char* line(size_t i, uint32_t& len) const {  // line i of the indexed region
        const size_t b = i == 0 ? 0 : (size_t)nl[i - 1] + 1;
        len = (uint32_t)(nl[i] - b);
        return mf.data + reg_lo + b;
    }
    void consume(size_t nrec) {
        if (nrec == 0) return;
        const size_t end = reg_lo + (size_t)nl[4 * nrec - 1] + 1;
        rec_bytes = 0.5 * rec_bytes + 0.5 * (double)(end - pos) / (double)nrec;
        pos = end;
    }
};

// Fill `b` / `c` from the mapped sources: records r = stride * i + which come from source `which`.  Returns false (nothing
// consumed) when a record is irregular; `n_out` = records appended.
static bool fill_batch_fast(Source& A, Source* B2, bool interleaved, HostBatch& b, Chunk& c, int threads, size_t& n_out) {
    n_out = 0;
    const size_t per = B2 ? b.cap_reads / 2 : (interleaved ? (b.cap_reads & ~(size_t)1) : b.cap_reads);
    size_t na = A.peek(per, threads);
    size_t nb = B2 ? B2->peek(per, threads) : 0;
    if (B2) {
        if (na != nb) {
            // one file ends before the other (main.cpp:463-466), unless the shorter one only has a ragged tail left
            const size_t m = na < nb ? na : nb;
            if (m == 0) return false;  // let the serial reader sort out what is left (and report unpaired files)
            na = nb = m;
        }
    }
    if (!B2 && interleaved && (na & 1)) {
        // mates must stay in one batch: an odd count is only right when these records end the file exactly (a dangling
        // record, corrected single-end); with bytes left over (a last record without its final newline) the pair is
        // split between the readers otherwise -- leave the odd record to the serial reader as well
        const size_t end = A.reg_lo + (size_t)A.nl[4 * na - 1] + 1;
        if (end < A.mf.size) na -= 1;
    }
    if (na == 0) return !(A.pos < A.mf.size) ? true : false;  // leftover bytes without 4 complete lines: serial reader
    const size_t stride = B2 ? 2 : 1;
    const size_t n = na * stride;
    // lengths + regularity
    std::vector<uint32_t> len(n);
    std::vector<char> bad(threads > 0 ? threads : 1, 0), toolong(threads > 0 ? threads : 1, 0);
    rcxio::parallel_for(threads, [&](int t) {
        const size_t lo = na * t / threads, hi = na * (t + 1) / threads;
        for (size_t i = lo; i < hi; ++i)
            for (size_t w = 0; w < stride; ++w) {
                const Source& S = w == 0 ? A : *B2;
                uint32_t L, qn, pl;
                S.line(4 * i + 1, L);
                S.line(4 * i + 3, qn);
                S.line(4 * i + 2, pl);
                if (qn != L) bad[t] = 1;
                if (L > RCX_MAX_READ_LENGTH) toolong[t] = 1;
                len[i * stride + w] = L;
            }
    });
    for (int t = 0; t < threads; ++t)
        if (bad[t] || toolong[t]) return false;  // serial reader: reference line semantics / its error message
    // offsets (and the base budget of the pinned buffers)
    size_t keep = n, acc = 0;
    b.off[0] = 0;
    for (size_t r = 0; r < n; ++r) {
        acc += len[r];
        if (acc > b.cap_bases - 2 * (RCX_MAX_READ_LENGTH + 1)) { keep = r & ~(size_t)1; break; }
        b.off[r + 1] = (uint32_t)acc;
    }
    if (keep == 0) return false;
    const size_t nrec = keep / stride;
    c.clear();
    c.ext_ids = true;
    c.id_ptr.resize(keep);
    c.id_len.resize(keep);
    rcxio::parallel_for(threads, [&](int t) {
        const size_t lo = nrec * t / threads, hi = nrec * (t + 1) / threads;
        for (size_t i = lo; i < hi; ++i)
            for (size_t w = 0; w < stride; ++w) {
                const Source& S = w == 0 ? A : *B2;
                const size_t r = i * stride + w;
                uint32_t l0, L;
                const char* idp = S.line(4 * i, l0);
                const char* sp = S.line(4 * i + 1, L);
                uint32_t qn;
                const char* qp = S.line(4 * i + 3, qn);
                c.id_ptr[r] = idp;
                c.id_len[r] = l0;
                memcpy(b.seq + b.off[r], sp, L);
                memcpy(b.qual + b.off[r], qp, L);
            }
    });
    b.n = keep;
    b.bases = b.off[keep];
    c.n = keep;
    A.consume(nrec);
    if (B2) B2->consume(nrec);
    n_out = keep;
    return true;
}

// The CUDA path counts the reads it cannot reproduce exactly (include/rcx.h: rcx_stats.unsupported — a third pass of the
// correction loop, a k-mer count of 2^29 or more, more substitutions on one search path than its edit row holds).  None
// has been seen on real or synthetic data; if one ever is, the run stops here instead of writing output that may differ from
// the reference's.
static void check_supported(rcx_ctx* ctx) {
    rcx_stats st;
    if (rcx_get_stats(ctx, &st)) die_rcx("rcx_get_stats");
    if (st.unsupported) {
        fprintf(stderr, "ERROR: %llu reads of this batch need a case the GPU path does not reproduce (see rcx_stats.unsupported); "
                        "no output is kept\n", (unsigned long long)st.unsupported);
        fflush(stderr);
        remove_partial_outputs();
        _exit(1);  // (formatting / writing threads of earlier batches may still be running)
    }
}

static void correct_on_gpus(std::vector<Gpu>& gpus, int mode, HostBatch& b) {
    const size_t n = b.n;
    if (n == 0) return;
    const size_t G = gpus.size();
    if (G == 1) {
        if (rcx_correct_batch(gpus[0].ctx, mode, (int64_t)n, b.seq, b.qual, b.off, b.seq, b.corr, b.l, b.m, b.h)) die_rcx("rcx_correct_batch");
        check_supported(gpus[0].ctx);
        return;
    }
    // contiguous record-index shards, both mates of a pair on the same GPU
    std::vector<std::thread> th;
    std::vector<std::string> errs(G);
    const size_t unit = mode == RCX_INTERLEAVED ? 2 : 1;
    const size_t units = n / unit;
    for (size_t g = 0; g < G; ++g) {
        const size_t lo = units * g / G * unit, hi = units * (g + 1) / G * unit;
        if (lo == hi) continue;
        th.emplace_back([&, g, lo, hi]() {
            std::vector<uint32_t> off(hi - lo + 1);
            for (size_t i = lo; i <= hi; ++i) off[i - lo] = b.off[i] - b.off[lo];
            if (rcx_correct_batch(gpus[g].ctx, mode, (int64_t)(hi - lo), b.seq + b.off[lo], b.qual + b.off[lo], off.data(), b.seq + b.off[lo],
                                  b.corr + lo, b.l + lo, b.m + lo, b.h + lo))
                errs[g] = rcx_last_error();
        });
    }
    for (auto& t : th) t.join();
    for (auto& e : errs)
        if (!e.empty()) {
            fprintf(stderr, "rcx_correct_batch: %s\n", e.c_str());
            exit(1);
        }
    for (auto& g : gpus) check_supported(g.ctx);
}

int main(int argc, char* argv[]) {
    int k = 23, threads = 1, max_fix_per_k = 4, ngpu = 1;
    double wk = 0.95;
    bool to_stdout = false, count_kmers = false;
    const char* dump_path = nullptr;
    std::string od = "./";
    size_t batch_reads = 1u << 20;  // records per pipeline batch (3 pinned buffers cycle through read -> correct -> write)
    if (const char* e = getenv("RCX_GPUS")) ngpu = atoi(e);
    if (argc == 1) {
        print_help();
        exit(0);
    }
    for (int i = 1; i < argc; ++i) {  // first pass: options (main.cpp:165-247)
        if (!strcmp("-r", argv[i])) { ++i; continue; }
        else if (!strcmp("-p", argv[i])) { i += 2; continue; }
        else if (!strcmp("-i", argv[i])) { ++i; continue; }
        else if (!strcmp("-od", argv[i])) { if (mkdir(argv[i + 1], 0700) == 0) g_od_created = argv[i + 1]; od = argv[i + 1]; ++i; }
        else if (!strcmp("-c", argv[i])) { dump_path = argv[i + 1]; ++i; }
        else if (!strcmp("-k", argv[i])) { k = atoi(argv[i + 1]); ++i; }
        else if (!strcmp("-t", argv[i])) { threads = atoi(argv[i + 1]); ++i; }
        else if (!strcmp("-maxcor", argv[i])) { ++i; }  // parsed and unused by the reference as well (dead parameter)
        else if (!strcmp("-maxcorK", argv[i])) { max_fix_per_k = atoi(argv[i + 1]); ++i; }
        else if (!strcmp("-wk", argv[i])) { wk = atof(argv[i + 1]); ++i; }
        else if (!strcmp("-stdout", argv[i])) { to_stdout = true; }
        else if (!strcmp("-verbose", argv[i])) { /* per-read trace to stdout: not produced */ }
        else if (!strcmp("-gpus", argv[i])) { ngpu = atoi(argv[i + 1]); ++i; }
        else if (!strcmp("-countkmers", argv[i])) { count_kmers = true; }
        else if (!strcmp("-batch", argv[i])) { batch_reads = (size_t)atol(argv[i + 1]); ++i; }
        else if (!strcmp("-h", argv[i])) { print_help(); exit(0); }
        else { fprintf(stderr, "Unknown argument: %s\n", argv[i]); exit(0); }
    }
    if (!dump_path && !count_kmers) {
        fprintf(stderr, "Could not open file (no -c jf_dump given)\n");
        exit(1);
    }
    if (dump_path) {
        FILE* t = fopen(dump_path, "r");
        if (!t) {
            fprintf(stderr, "Could not open file %s\n", dump_path);
            exit(1);
        }
        fclose(t);
    }
    if (k < 2 || k > RCX_MAX_K) {
        fprintf(stderr, "The kmer size must be between 2 and %d.\n", RCX_MAX_K);
        exit(1);
    }
    if (threads < 1) threads = 1;
    if (ngpu < 1) ngpu = 1;
    if (batch_reads < 2) batch_reads = 2;
    batch_reads &= ~(size_t)1;

    std::vector<ReadFile> files;
    std::vector<std::string> mates;
    for (int i = 1; i < argc; ++i) {  // second pass: read files in argv order (main.cpp:250-268)
        if (!strcmp("-r", argv[i])) { ReadFile f; f.path = argv[i + 1]; files.push_back(f); ++i; }
        else if (!strcmp("-p", argv[i])) { ReadFile f; f.path = argv[i + 1]; f.paired = true; files.push_back(f); mates.push_back(argv[i + 2]); i += 2; }
        else if (!strcmp("-i", argv[i])) { ReadFile f; f.path = argv[i + 1]; f.interleaved = true; files.push_back(f); ++i; }
        else if (!strcmp("-od", argv[i]) || !strcmp("-c", argv[i]) || !strcmp("-k", argv[i]) || !strcmp("-t", argv[i]) || !strcmp("-maxcor", argv[i]) ||
                 !strcmp("-maxcorK", argv[i]) || !strcmp("-wk", argv[i]) || !strcmp("-gpus", argv[i]) || !strcmp("-batch", argv[i]))
            ++i;
    }
    if (files.size() > 100 || mates.size() > 100) {  // MAX_READ_FILE, Reads.h:11,110-114
        fprintf(stderr, "The number of read files exceeds the limit %d.\n", 100);
        exit(1);
    }
    {
        size_t mi = 0;
        for (auto& f : files) {
            f.type = sniff_type(f.path);
            f.out_path = out_name(od, f.path, f.type);
            if (f.paired) {
                f.mate_path = mates[mi++];
                const int t2 = sniff_type(f.mate_path);
                f.mate_out_path = out_name(od, f.mate_path, t2);
            }
        }
    }
    // the reference opens (truncates) every output when the file is added (Reads.h:159-160)
    std::vector<Writer> out1(files.size()), out2(files.size());
    for (size_t i = 0; i < files.size(); ++i) {
        if (!out1[i].open(files[i].out_path, to_stdout)) {
            fprintf(stderr, "ERROR: Could not access file %s\n", files[i].out_path.c_str());  // File.h:71,93
            exit(1);
        }
        if (!to_stdout) g_out_paths.push_back(files[i].out_path);
        if (files[i].paired && !out2[i].open(files[i].mate_out_path, to_stdout)) {
            fprintf(stderr, "ERROR: Could not access file %s\n", files[i].mate_out_path.c_str());
            exit(1);
        }
        if (!to_stdout && files[i].paired) g_out_paths.push_back(files[i].mate_out_path);
    }

    // ---- badQualityThreshold (main.cpp:88-128): first 1e6 records of the first Reads object.  Independent of the table, so the
    // pass over the reads (a serial inflate for *.gz input) runs while the dump is parsed and the table is built; its line is
    // printed where the reference prints it.
    std::future<int> bad_qual_future = std::async(std::launch::async, [&files]() -> int {
    int bad_qual = 0;
    if (!files.empty() && files[0].type == 1) {
        int first[300], last[300];
        memset(first, 0, sizeof first);
        memset(last, 0, sizeof last);
        long total = 0;
        std::string lastq;  // quality line of the last FASTQ record of this pass
        for (size_t fi = 0; fi < files.size() && total < 1000000; ++fi) {
            LineReader r;
            if (!r.open(files[fi].path)) continue;
            const char* p;
            size_t n;
            bool nl;
            const int per = files[fi].type == 1 ? 4 : 2;
            while (total < 1000000 && r.getline(p, n, nl)) {
                size_t L = 0;
                if (r.getline(p, n, nl)) L = n;
                if (per == 4) {
                    r.getline(p, n, nl);
                    if (r.getline(p, n, nl) && L > 0 && n > 0) {
                        ++last[(unsigned char)(L - 1 < n ? p[L - 1] : 0)];
                        ++first[(unsigned char)p[0]];
                        lastq.assign(p, n);
                    }
                } else if (L > 0) {
                    // a FASTA file after FASTQ ones: Reads::Next (Reads.h:186-222) leaves its quality buffer alone for a FASTA
                    // record, so GetBadQuality counts the LAST FASTQ record's quality line again, indexed by this record's length
                    ++last[(unsigned char)(L - 1 < lastq.size() ? lastq[L - 1] : 0)];
                    ++first[(unsigned char)(lastq.empty() ? 0 : lastq[0])];
                }
                ++total;
            }
        }
        int i, cnt = 0, t1, t2;
        for (i = 0; i < 300; ++i) { cnt += first[i]; if (cnt > total * 0.05) break; }
        t1 = i - 1;
        cnt = 0;
        for (i = 0; i < 300; ++i) { cnt += last[i]; if (cnt > total * 0.05) break; }
        t2 = i;
        bad_qual = (int)(char)(t2 < t1 ? t2 : t1);
    }
    return bad_qual;
    });

    // ---- GPUs and table
    // On a multi-GPU host the CUDA runtime initialises every visible device at the first call (several seconds for eight):
    // unless the user chose the devices, only the first -gpus N are made visible to this process.
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        std::string vis;
        for (int g = 0; g < ngpu; ++g) vis += (g ? "," : "") + std::to_string(g);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
    }
    // CUDA context creation and the pinned batch buffers overlap the parsing of the dump
    std::vector<Gpu> gpus(ngpu);
    std::future<void> gpus_ready = std::async(std::launch::async, [&]() {
        for (int g = 0; g < ngpu; ++g)
            if (rcx_create(g, &gpus[g].ctx)) die_rcx("rcx_create");
    });
    // the three pinned batch buffers are allocated while the dump is parsed (pinning memory is slow)
    const int NSETS = 3;
    struct BatchSet {
        HostBatch b;
        Chunk c;
        bool more = true;  // the file may have further records after this batch
        std::vector<std::string> pieces[2];   // formatted output per mate file and formatting thread (capacity kept across batches)
        std::vector<std::string> zpieces[2];  // the same as gzip members, for *.gz outputs
    };
    std::vector<BatchSet> sets(NSETS);
    std::future<void> sets_ready = std::async(std::launch::async, [&]() {
        gpus_ready.wait();
        // bases per batch: 160 per record (records of 100-150 bases fill the record budget first), at most 3 GB, and never
        // less than a few maximal reads (tiny -batch values are used by tests)
        size_t cap_bases = batch_reads * 160 < (size_t)3000000000u ? batch_reads * 160 : (size_t)3000000000u;
        if (cap_bases < (size_t)16 * (RCX_MAX_READ_LENGTH + 1)) cap_bases = (size_t)16 * (RCX_MAX_READ_LENGTH + 1);
        for (auto& st : sets) st.b.alloc(batch_reads, cap_bases);
    });
    Dump dump;
    if (dump_path) {
        load_dump(dump_path, k, dump, threads);
        tmark("dump parsed");
    }
    gpus_ready.wait();
    tmark("CUDA contexts ready");
    if (!dump_path) {
        // -countkmers: what `jellyfish count -C` + `dump -L 2` would have handed over -- the canonical k-mers
        // of every input file (mates included) seen at least twice -- counted on the first GPU.  Sequence lines only.
        rcx_ctx* c0 = gpus[0].ctx;
        if (rcx_count_kmers_begin(c0, k, 0)) die_rcx("rcx_count_kmers_begin");
        std::vector<char> cs;
        std::vector<uint32_t> co(1, 0);
        auto flush = [&]() {
            if (co.size() > 1 && rcx_count_kmers_add(c0, (int64_t)co.size() - 1, cs.data(), co.data())) die_rcx("rcx_count_kmers_add");
            cs.clear();
            co.assign(1, 0);
        };
        auto count_file = [&](const std::string& path, int type) {
            LineReader r;
            if (!r.open(path)) { fprintf(stderr, "ERROR: Could not access file %s\n", path.c_str()); exit(1); }
            const char* p;
            size_t n;
            bool nl;
            while (r.getline(p, n, nl)) {      // id line
                if (!r.getline(p, n, nl)) break;  // sequence line
                cs.insert(cs.end(), p, p + n);
                co.push_back((uint32_t)cs.size());
                if (type == 1) { r.getline(p, n, nl); r.getline(p, n, nl); }  // '+' and quality lines
                if (cs.size() > (256u << 20)) flush();
            }
        };
        for (auto& f : files) {
            count_file(f.path, f.type);
            if (f.paired) count_file(f.mate_path, sniff_type(f.mate_path));
        }
        flush();
        uint64_t kept = 0, distinct = 0, instances = 0;
        if (rcx_count_kmers_finish(c0, 2, &kept, &distinct, &instances)) die_rcx("rcx_count_kmers_finish");
        dump.codes.resize(kept);
        dump.counts.resize(kept);
        if (rcx_count_kmers_export(c0, dump.codes.data(), dump.counts.data(), kept)) die_rcx("rcx_count_kmers_export");
        dump.stored = (int64_t)kept;
        dump.rate_records = kept;
        tmark("k-mers counted");
    }
    if (dump_path || ngpu > 1)
    {   // one upload over PCIe, device-to-device copies to the other GPUs (NVLink), local inserts in parallel
        std::vector<rcx_ctx*> ctxs;
        for (auto& g : gpus) ctxs.push_back(g.ctx);
        if (rcx_table_build_multi(ctxs.data(), ngpu, k, dump.codes.data(), dump.counts.data(), dump.codes.size())) die_rcx("rcx_table_build");
    }
    fprintf(stderr, "Stored %d kmers\n", (int)dump.stored);
    tmark("table built");

    // ---- ERROR_RATE (main.cpp:310-358)
    double error_rate = 0.01;
    if (rcx_estimate_error_rate(gpus[0].ctx, dump.codes.data(), dump.rate_records, wk, &error_rate)) die_rcx("rcx_estimate_error_rate");
    fprintf(stderr, "Weak kmer threshold rate: %lf (estimated from %.3lf/1 of the chosen kmers)\n", error_rate, wk);
    tmark("error rate estimated");
    if (getenv("RCX_DEBUG_SCALARS")) fprintf(stderr, "RCX_DEBUG error_rate=%.17g\n", error_rate);
    { std::vector<uint64_t>().swap(dump.codes); std::vector<int32_t>().swap(dump.counts); }

    const int bad_qual = bad_qual_future.get();
    fprintf(stderr, "Bad quality threshold is '%c'\n", bad_qual);
    tmark("bad quality threshold found");

    rcx_params params;
    params.k = k;
    params.max_fix_per_k = max_fix_per_k;
    params.error_rate = error_rate;
    params.bad_qual = bad_qual;
    for (auto& g : gpus)
        if (rcx_set_params(g.ctx, &params)) die_rcx("rcx_set_params");

    // ---- correction loop (replaces main.cpp:368-523): a three-stage pipeline over three pinned batch buffers --
    // stage 1 reads + parses batch i+1 (parallel reader), stage 2 corrects batch i on the GPUs (this thread), stage 3
    // formats (in parallel) and writes batch i-1; output order = input order.
    uint64_t total_reads = 0, total_corrections = 0;
    sets_ready.get();
    tmark("pinned batch buffers ready");
    const int nfmt = std::min(threads, 32);
    const bool fast_io = !getenv("RCX_SERIAL_IO");
    std::shared_future<void> last_write;     // the latest batch's formatting + writing
    std::shared_future<void> last_reserved;  // chain of reservations: keeps the output in input order
    std::vector<std::shared_future<void>> set_write(NSETS);
    for (size_t fi = 0; fi < files.size(); ++fi) {
        const ReadFile& f = files[fi];
        Source s1, s2;
        if (!s1.open(f.path, f.type, fast_io) || (f.paired && !s2.open(f.mate_path, sniff_type(f.mate_path), fast_io))) {
            fprintf(stderr, "ERROR: Could not access file %s\n", f.path.c_str());
            exit(1);
        }
        if (f.paired && s1.fast != s2.fast) {  // both mates through the same reader
            if (s1.fast) s1.go_serial();
            if (s2.fast) s2.go_serial();
        }
        const int mode = (f.paired || f.interleaved) ? RCX_INTERLEAVED : RCX_SINGLE;
        // stage 1: fill a batch set; returns false when the file is exhausted and the set is empty
        std::atomic<bool> unpaired{false};  // set by stage 1 when one file of a -p pair ends before the other
        auto produce = [&](BatchSet& st) -> bool {
            StageTimer timer(g_us_read);
            HostBatch& b = st.b;
            Chunk& c = st.c;
            b.n = 0;
            b.bases = 0;
            b.off[0] = 0;
            c.clear();
            st.more = true;
            if (s1.fast) {
                size_t got = 0;
                if (fill_batch_fast(s1, f.paired ? &s2 : nullptr, f.interleaved, b, c, std::max(1, threads), got)) {
                    if (got == 0) { st.more = false; return false; }
                    return true;
                }
                s1.go_serial();  // irregular record ahead: the serial reader takes over from here
                if (f.paired) s2.go_serial();
                b.n = 0; b.bases = 0; b.off[0] = 0;
                c.clear();
            }
            LineReader& r1 = s1.lr;
            LineReader& r2 = s2.lr;
            bool more = true;
            while (b.n + 2 <= b.cap_reads && b.bases + 2 * (RCX_MAX_READ_LENGTH + 1) <= b.cap_bases) {
                if (!read_record(r1, f.type, b, c, f.path)) { more = false; break; }
                if (f.paired) {
                    if (!read_record(r2, f.type, b, c, f.mate_path)) {
                        // main.cpp:463-466.  The error is raised by the main loop once the batches before it are written
                        // (the reference, too, has written its earlier batches by then); the dangling mate is dropped.
                        unpaired = true;
                        b.n -= 1;
                        b.bases = b.off[b.n];
                        c.id_off.pop_back();
                        c.ids.resize(c.id_off.back());
                        if (c.qual_text_off.size() > b.n + 1) { c.qual_text_off.resize(b.n + 1); c.qual_text.resize(c.qual_text_off.back()); }
                        c.n = b.n;
                        more = false;
                        break;
                    }
                } else if (f.interleaved) {
                    if (!read_record(r1, f.type, b, c, f.path)) {
                        // odd record count in an interleaved file: the reference pairs it with a stale buffer; the dangling
                        // record is corrected single-end below
                        more = false;
                        break;
                    }
                }
            }
            if (f.paired && !more) {
                const char* p; size_t n; bool nl;
                if (!unpaired && r2.getline(p, n, nl)) unpaired = true;  // the second file is the longer one
            }
            st.more = more;
            return b.n > 0;
        };
        // stage 3: format + write.  Formatting runs on `nfmt` threads into per-thread buffers that keep their capacity from
        // batch to batch; plain files are then written by the same threads with positional writes at offsets reserved in
        // batch order (the batches' writes themselves may overlap); gzip files and stdout are sequential streams, written in
        // batch order.
        // `prev_reserved` becomes ready when the batch before this one has its place in the output (plain files: right after
        // the reservation, so the batches' formatting and page faults overlap; streams: when it is written); `reserved` is
        // this batch's turn of the same chain.
        auto consume = [&, fi](BatchSet& st, std::shared_future<void> prev_reserved, std::shared_ptr<std::promise<void>> reserved) {
            HostBatch& b = st.b;
            Chunk& c = st.c;
            struct Part { Writer* w; size_t start, stride, recs; int parts; std::vector<std::string>* pieces; std::vector<std::string>* out; };
            Part part[2];
            int nparts = 0;
            auto plan = [&](Writer& w, size_t start, size_t stride, int which) {
                Part& p = part[nparts++];
                p.w = &w; p.start = start; p.stride = stride;
                p.recs = b.n > start ? (b.n - start + stride - 1) / stride : 0;
                p.parts = (int)std::min<size_t>((size_t)nfmt, std::max<size_t>(1, p.recs / 4096));
                p.pieces = &st.pieces[which];
                p.out = w.gz() ? &st.zpieces[which] : p.pieces;  // what goes to the file
                if ((int)p.pieces->size() < p.parts) p.pieces->resize(p.parts);
                if ((int)p.out->size() < p.parts) p.out->resize(p.parts);
            };
            // ---- output in input order (main.cpp:485-516); -stdout interleaves the mates of a -p pair
            if (f.paired && !to_stdout) {
                plan(out1[fi], 0, 2, 0);
                plan(out2[fi], 1, 2, 1);
            } else {
                plan(out1[fi], 0, 1, 0);  // single, interleaved, or paired to stdout (mate1, mate2, mate1, ...)
            }
            auto range = [&](const Part& p, int t, size_t& lo, size_t& hi) {
                lo = p.start + p.stride * (p.recs * t / p.parts);
                hi = std::min(p.start + p.stride * (p.recs * (t + 1) / p.parts), b.n);
            };
            bool direct = !getenv("RCX_NO_DIRECT_OUT");
            for (int q = 0; q < nparts; ++q) direct = direct && part[q].w->can_scatter() && !part[q].w->gz();
            if (direct) {
                // plain files: piece sizes first (no bytes move), then the place in the file, then every thread formats its
                // piece straight into the mapping of that region -- no intermediate buffer, no second copy
                std::vector<size_t> sizes[2];
                {
                    StageTimer timer(g_us_format);
                    for (int q = 0; q < nparts; ++q) {
                        Part& p = part[q];
                        sizes[q].assign(p.parts, 0);
                        rcxio::parallel_for(p.parts, [&](int t) {
                            size_t lo, hi;
                            range(p, t, lo, hi);
                            CountSink cs;
                            emit_records(cs, b, c, f.type, lo, hi, p.stride);
                            sizes[q][t] = cs.n;
                        });
                    }
                }
                if (prev_reserved.valid()) prev_reserved.wait();
                uint64_t at0[2] = {0, 0}, total[2] = {0, 0};
                bool mapped_ok = true;
                void* mbase[2] = {nullptr, nullptr};
                size_t mlen[2] = {0, 0};
                char* dst[2] = {nullptr, nullptr};
                for (int q = 0; q < nparts; ++q) {
                    for (int t = 0; t < part[q].parts; ++t) total[q] += sizes[q][t];
                    at0[q] = part[q].w->reserve(total[q]);
                    dst[q] = total[q] ? part[q].w->map(at0[q], total[q], mbase[q], mlen[q]) : nullptr;  // (grows the file: in order)
                    if (total[q] && !dst[q]) mapped_ok = false;
                }
                reserved->set_value();
                StageTimer timer(g_us_write);
                for (int q = 0; q < nparts; ++q) {
                    Part& p = part[q];
                    if (!total[q]) continue;
                    std::vector<size_t> offs(p.parts, 0);
                    for (int t = 1; t < p.parts; ++t) offs[t] = offs[t - 1] + sizes[q][t - 1];
                    std::atomic<int> failed{0};
                    rcxio::parallel_for(p.parts, [&](int t) {
                        size_t lo, hi;
                        range(p, t, lo, hi);
                        if (mapped_ok) {
                            PtrSink ps{dst[q] + offs[t]};
                            emit_records(ps, b, c, f.type, lo, hi, p.stride);
                        } else {  // the file could not be mapped: format into a buffer, positional write
                            format_records(b, c, f.type, lo, hi, p.stride, (*p.pieces)[t]);
                            if (!p.w->write_at(at0[q] + offs[t], (*p.pieces)[t].data(), (*p.pieces)[t].size())) failed = 1;
                        }
                    });
                    Writer::unmap(mbase[q], mlen[q]);
                    if (failed) {
                        fprintf(stderr, "ERROR: Could not write the output of %s\n", f.path.c_str());
                        exit(1);
                    }
                }
                return;
            }
            {
                StageTimer timer(g_us_format);
                for (int q = 0; q < nparts; ++q) {
                    Part& p = part[q];
                    rcxio::parallel_for(p.parts, [&](int t) {
                        size_t lo, hi;
                        range(p, t, lo, hi);
                        format_records(b, c, f.type, lo, hi, p.stride, (*p.pieces)[t]);
                        if (p.w->gz()) {  // one gzip member per piece, level 1 like the reference's gzopen(.., "w1")
                            if ((*p.pieces)[t].empty()) (*p.out)[t].clear();
                            else if (!rcxio::gzip_member((*p.pieces)[t].data(), (*p.pieces)[t].size(), 1, (*p.out)[t])) {
                                fprintf(stderr, "ERROR: Could not compress the output of %s\n", f.path.c_str());
                                exit(1);
                            }
                        }
                    });
                }
            }
            if (prev_reserved.valid()) prev_reserved.wait();  // the previous batch has reserved its place (plain files) / is written (streams)
            StageTimer timer(g_us_write);
            for (int q = 0; q < nparts; ++q) {
                Part& p = part[q];
                if (p.w->can_scatter()) {
                    std::vector<uint64_t> at(p.parts);
                    uint64_t total = 0;
                    for (int t = 0; t < p.parts; ++t) { at[t] = p.w->reserve((*p.out)[t].size()); total += (*p.out)[t].size(); }
                    std::atomic<int> failed{0};
                    void* mbase = nullptr;
                    size_t mlen = 0;
                    char* dst = (p.parts > 1 && !getenv("RCX_NO_MMAP_OUT")) ? p.w->map(at[0], total, mbase, mlen) : nullptr;
                    rcxio::parallel_for(p.parts, [&](int t) {
                        if (dst) memcpy(dst + (at[t] - at[0]), (*p.out)[t].data(), (*p.out)[t].size());
                        else if (!p.w->write_at(at[t], (*p.out)[t].data(), (*p.out)[t].size())) failed = 1;
                    });
                    Writer::unmap(mbase, mlen);
                    if (failed) {
                        fprintf(stderr, "ERROR: Could not write the output of %s\n", f.path.c_str());
                        exit(1);
                    }
                } else {
                    for (int t = 0; t < p.parts; ++t) p.w->write((*p.pieces)[t].data(), (*p.pieces)[t].size());
                }
            }
            reserved->set_value();
        };
        int cur = 0;
        bool have = produce(sets[cur]);
        while (have) {
            BatchSet& st = sets[cur];
            HostBatch& b = st.b;
            const int nxt = (cur + 1) % NSETS;
            std::future<bool> fnext;
            const bool may_have_more = st.more;
            if (may_have_more) {
                if (set_write[nxt].valid()) set_write[nxt].wait();  // its previous contents are written
                fnext = std::async(std::launch::async, [&, nxt]() { return produce(sets[nxt]); });
            }
            size_t odd_tail = 0;
            if (mode == RCX_INTERLEAVED && (b.n & 1)) odd_tail = 1;
            const double t_correct0 = g_timing ? now_s() : 0;
            if (odd_tail) {
                // correct the even prefix as pairs, the dangling record single-end
                HostBatch t = b;
                t.n = b.n - 1;
                correct_on_gpus(gpus, RCX_INTERLEAVED, t);
                HostBatch sgl = b;
                sgl.seq = b.seq + b.off[b.n - 1]; sgl.qual = b.qual + b.off[b.n - 1];
                uint32_t so[2] = {0, b.off[b.n] - b.off[b.n - 1]};
                sgl.off = so; sgl.corr = b.corr + b.n - 1; sgl.l = b.l + b.n - 1; sgl.m = b.m + b.n - 1; sgl.h = b.h + b.n - 1; sgl.n = 1;
                correct_on_gpus(gpus, RCX_SINGLE, sgl);
            } else {
                correct_on_gpus(gpus, mode, b);
            }
            if (g_timing) g_us_correct += (long long)((now_s() - t_correct0) * 1e6);
            for (size_t r = 0; r < b.n; ++r) {  // UpdateSummary, main.cpp:73-79
                ++total_reads;
                if (b.corr[r] > 0) total_corrections += (uint64_t)b.corr[r];
            }
            std::shared_future<void> prev = last_reserved;
            auto rp = std::make_shared<std::promise<void>>();
            last_reserved = rp->get_future().share();
            last_write = std::async(std::launch::async, [&, cur, prev, rp]() { consume(sets[cur], prev, rp); }).share();
            set_write[cur] = last_write;
            have = may_have_more ? fnext.get() : false;
            cur = nxt;
        }
        for (auto& w : set_write) if (w.valid()) w.wait();  // every batch of this file is written: its sources go out of scope
        if (unpaired) {
            fprintf(stderr, "ERROR: The files are not paired!\n");  // main.cpp:463-466
            exit(1);
        }
    }
    for (auto& w : out1) w.close();
    for (auto& w : out2) w.close();
    for (auto& g : gpus) if (rcx_sync(g.ctx)) die_rcx("rcx_sync");
    fprintf(stderr, "Processed %" PRIu64 " reads\n\tCorrected %" PRIu64 " bases.\n", total_reads, total_corrections);
    tmark("done");
    if (g_timing)
        fprintf(stderr, "RCX_TIMING stage sums: read %.3f s, correct %.3f s, format %.3f s, write %.3f s\n", g_us_read.load() * 1e-6,
                g_us_correct.load() * 1e-6, g_us_format.load() * 1e-6, g_us_write.load() * 1e-6);
    fflush(stdout);
    fflush(stderr);
    _exit(0);  // everything is written and closed: skip the teardown of pinned buffers, mappings and CUDA contexts
}
