// fastq_io.h — host-side read I/O of the drop-in `rcorrector` binary.
//
// Behaviour follows the reference's reader/writer (Reads.h:108-162 AddReadFile,
// :224-266 NextWithBuffer, :360-421 OutputBatch, :39-75 GetFileName;
// File.h:51-69 gz detection by the last two characters of the path, gzopen
// "w1") so that the bytes written equal the reference's; the implementation is
// a chunked block reader + batch formatter instead of per-line fgets/Printf.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace rcxio {

// run fn(t) for t in [0, nthreads) on that many host threads
template <class F>
inline void parallel_for(int nthreads, F fn) {
    if (nthreads <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&fn, t]() { fn(t); });
    for (auto& x : th) x.join();
}

// A plain (not gzip) input file mapped read-only: the parallel readers index and slice it in place.
struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    bool open(const std::string& path) {
        close();
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd); return false; }
        size = (size_t)st.st_size;
        if (size > 0) {
            void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) { ::close(fd); size = 0; return false; }
            madvise(p, size, MADV_SEQUENTIAL);
            data = (const char*)p;
        }
        ::close(fd);
        return true;
    }
    void close() {
        if (data) munmap((void*)data, size);
        data = nullptr;
        size = 0;
    }
    ~MappedFile() { close(); }
};

// positions of every '\n' in [lo, hi) of a mapped file, found by `nthreads` threads (ascending, relative to `lo`)
inline void index_newlines(const char* base, size_t lo, size_t hi, int nthreads, std::vector<uint32_t>& out) {
    const size_t n = hi - lo;
    std::vector<std::vector<uint32_t>> part(nthreads);
    parallel_for(nthreads, [&](int t) {
        const size_t a = lo + n * t / nthreads, b = lo + n * (t + 1) / nthreads;
        std::vector<uint32_t>& v = part[t];
        v.reserve((b - a) / 48 + 16);
        const char* p = base + a;
        const char* e = base + b;
        while (p < e) {
            const char* q = (const char*)memchr(p, '\n', e - p);
            if (!q) break;
            v.push_back((uint32_t)(q - (base + lo)));
            p = q + 1;
        }
    });
    size_t tot = 0;
    for (auto& v : part) tot += v.size();
    out.resize(tot);
    std::vector<size_t> start(nthreads + 1, 0);
    for (int t = 0; t < nthreads; ++t) start[t + 1] = start[t] + part[t].size();
    parallel_for(nthreads, [&](int t) { if (!part[t].empty()) memcpy(out.data() + start[t], part[t].data(), part[t].size() * sizeof(uint32_t)); });
}

inline bool ends_with_gz(const std::string& p) {  // File.h:51-55
    return p.size() >= 2 && p[p.size() - 2] == 'g' && p[p.size() - 1] == 'z';
}

// Reads.h:39-75: basename without directory and without the last extension; for *.gz also a preceding
// .fastq/.fasta/.fq/.fa is dropped.
inline std::string output_stem(const std::string& in) {
    const int len = (int)in.size();
    int i = len - 1;  // last '.' or '/' scanning from the end
    while (i >= 0 && in[i] != '.' && in[i] != '/') --i;
    if (i >= 0 && in.compare(i, std::string::npos, ".gz") == 0) {
        const int tmp = i;
        int j = i - 1;
        while (j >= 0 && in[j] != '.' && in[j] != '/') --j;
        const std::string ext = j >= 0 ? in.substr(j, tmp - j) : std::string();
        if (j >= 0 && (ext == ".fastq" || ext == ".fasta" || ext == ".fq" || ext == ".fa")) i = j;
        else i = tmp;
    }
    int s = len - 1;
    while (s >= 0 && in[s] != '/') --s;
    if (i >= 0 && in[i] == '.') return in.substr(s + 1, i - (s + 1));
    return in.substr(s + 1);
}

// Reads.h:138-157: <od>/<stem>.cor.fq (FASTQ, type 1) or .cor.fa (FASTA), plus ".gz" when the input path ends in "gz"
inline std::string output_path(const std::string& od, const std::string& in, int type) {
    std::string s = od + "/" + output_stem(in) + (type == 1 ? ".cor.fq" : ".cor.fa");
    if (ends_with_gz(in)) s += ".gz";
    return s;
}

// Block reader over a plain or gzip file that hands out lines (without the '\n').  A gzip file is inflated by a thread of
// its own, a few blocks ahead of the consumer: inflating is the slowest step of reading *.gz input (and cannot be split
// inside one stream), so it should at least overlap the parsing — and the two files of a -p pair inflate side by side.
class LineReader {
   public:
    LineReader() = default;
    LineReader(const LineReader&) = delete;
    LineReader& operator=(const LineReader&) = delete;
    // serial reading of an already mapped file from byte `pos` (the parallel reader hands irregular input over here)
    void open_mem(const char* data, size_t size, size_t pos) {
        close();
        mem_ = data;
        beg_ = pos;
        end_ = size;
        eof_ = true;
    }
    size_t tell() const { return beg_; }  // memory mode: offset of the next unread byte
    bool open(const std::string& path) {
        close();
        gz_ = ends_with_gz(path);
        if (gz_) {
            g_ = gzopen(path.c_str(), "r");
            if (!g_) return false;
            gzbuffer(g_, 1 << 20);
            inflater_ = std::thread([this]() { inflate_ahead(); });
            return true;
        }
        f_ = fopen(path.c_str(), "r");
        return f_ != nullptr;
    }
    void close() {
        if (inflater_.joinable()) {
            { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
            cv_.notify_all();
            inflater_.join();
        }
        blocks_.clear();
        stop_ = false;
        inflate_done_ = false;
        if (f_) fclose(f_);
        if (g_) gzclose(g_);
        f_ = nullptr;
        g_ = nullptr;
        mem_ = nullptr;
        beg_ = end_ = 0;
        eof_ = false;
        gz_ = false;
    }
    ~LineReader() { close(); }
    // returns false at end of file; `had_nl` tells whether the line ended with '\n'
    bool getline(const char*& p, size_t& n, bool& had_nl) {
        for (;;) {
            if (beg_ < end_) {
                const char* s = (mem_ ? mem_ : buf_.data()) + beg_;
                const void* nl = memchr(s, '\n', end_ - beg_);
                if (nl) {
                    p = s;
                    n = (const char*)nl - s;
                    had_nl = true;
                    beg_ += n + 1;
                    return true;
                }
                if (eof_) {
                    p = s;
                    n = end_ - beg_;
                    had_nl = false;
                    beg_ = end_;
                    return true;
                }
            } else if (eof_) {
                return false;
            }
            fill();
        }
    }

   private:
    static constexpr size_t kBlock = 4u << 20;  // inflated bytes per hand-over
    static constexpr size_t kAhead = 8;         // blocks the inflater may be ahead
    void inflate_ahead() {  // the gzip file's own thread
        for (;;) {
            std::vector<char> blk(kBlock);
            const int got = gzread(g_, blk.data(), (unsigned)blk.size());
            std::unique_lock<std::mutex> lk(mu_);
            if (got <= 0) { inflate_done_ = true; cv_.notify_all(); return; }
            blk.resize((size_t)got);
            cv_.wait(lk, [this]() { return stop_ || blocks_.size() < kAhead; });
            if (stop_) return;
            blocks_.push_back(std::move(blk));
            cv_.notify_all();
        }
    }
    void fill() {
        if (beg_ > 0) {
            memmove(buf_.data(), buf_.data() + beg_, end_ - beg_);
            end_ -= beg_;
            beg_ = 0;
        }
        if (gz_) {
            std::vector<char> blk;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this]() { return !blocks_.empty() || inflate_done_; });
                if (blocks_.empty()) { eof_ = true; return; }
                blk = std::move(blocks_.front());
                blocks_.pop_front();
            }
            cv_.notify_all();
            if (buf_.size() - end_ < blk.size()) buf_.resize(end_ + blk.size() + (8u << 20));
            memcpy(buf_.data() + end_, blk.data(), blk.size());
            end_ += blk.size();
            return;
        }
        if (buf_.size() - end_ < (1u << 20)) buf_.resize(buf_.size() + (8u << 20));
        const long got = (long)fread(buf_.data() + end_, 1, buf_.size() - end_, f_);
        if (got <= 0) eof_ = true;
        else end_ += (size_t)got;
    }
    FILE* f_ = nullptr;
    gzFile g_ = nullptr;
    const char* mem_ = nullptr;
    bool gz_ = false, eof_ = false;
    std::vector<char> buf_ = std::vector<char>(8u << 20);
    size_t beg_ = 0, end_ = 0;
    // gzip: blocks inflated ahead by inflater_
    std::thread inflater_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::vector<char>> blocks_;
    bool stop_ = false, inflate_done_ = false;
};

// One complete gzip member (RFC 1952) holding p[0..n) at `level`.  A gzip file is a concatenation of members, so the
// pieces of a batch can be compressed by several threads and written back to back: gzread / gzgets (what the reference reads
// its input with), gzip -d and every other reader see the concatenation of the pieces.
inline bool gzip_member(const char* p, size_t n, int level, std::string& out) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, (uLong)n) + 64);
    zs.next_out = (Bytef*)&out[0];
    size_t out_left = out.size(), done = 0;
    int rc = Z_OK;
    do {
        const size_t in_now = std::min<size_t>(n - done, (size_t)1 << 30);  // avail_in / avail_out are 32-bit
        const size_t out_now = std::min<size_t>(out_left, (size_t)1 << 30);
        zs.next_in = (Bytef*)(p + done);
        zs.avail_in = (uInt)in_now;
        zs.avail_out = (uInt)out_now;
        rc = deflate(&zs, done + in_now == n ? Z_FINISH : Z_NO_FLUSH);
        done += in_now - zs.avail_in;
        out_left -= out_now - zs.avail_out;
        if (rc == Z_STREAM_ERROR || (rc == Z_BUF_ERROR && out_left == 0)) { deflateEnd(&zs); return false; }
    } while (rc != Z_STREAM_END);
    out.resize(out.size() - out_left);
    deflateEnd(&zs);
    return true;
}

// Output file.  Files are written through a file descriptor so that the pieces of a batch can be written by several
// threads at once: reserve() hands out the file offsets in output order, write_at() is a positional write.  For *.gz
// outputs (File.h:51-66: a path ending in "gz" is written through zlib at level 1) the pieces are gzip members, compressed
// by the threads that formatted them.  stdout stays a sequential stream.
class Writer {
   public:
    bool open(const std::string& path, bool to_stdout) {  // File.h:38-75
        if (to_stdout) { f_ = stdout; return true; }
        gz_ = ends_with_gz(path);
        fd_ = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0666);  // the file fopen(path, "w") / gzopen(path, "w1") create (mapped for writing)
        pos_ = 0;
        return fd_ >= 0;
    }
    bool can_scatter() const { return fd_ >= 0; }
    bool gz() const { return gz_; }  // pieces handed to reserve() / write_at() must be gzip members (gzip_member(), level 1)
    // the next n bytes of the file belong to the caller; call in output order
    uint64_t reserve(uint64_t n) { const uint64_t at = pos_; pos_ += n; return at; }
    // Grow the file to cover everything reserved so far and map [at, at + n) of it for the writing threads (memcpy into
    // the mapping takes its page faults on the thread that copies; positional writes to ONE file serialise on the inode
    // lock).  Returns nullptr when the file cannot be grown or mapped (the caller falls back to write_at()).
    char* map(uint64_t at, uint64_t n, void*& base, size_t& len) const {
        if (n == 0 || fd_ < 0 || ftruncate(fd_, (off_t)pos_) != 0) return nullptr;
        const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE), lo = at & ~(page - 1);
        len = (size_t)(at + n - lo);
        base = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, (off_t)lo);
        if (base == MAP_FAILED) { base = nullptr; return nullptr; }
        return (char*)base + (at - lo);
    }
    static void unmap(void* base, size_t len) { if (base) munmap(base, len); }
    // positional write (any thread); false on an I/O error
    bool write_at(uint64_t at, const char* p, size_t n) const {
        while (n) {
            const ssize_t w = ::pwrite(fd_, p, n, (off_t)at);
            if (w <= 0) return false;
            p += w; at += (uint64_t)w; n -= (size_t)w;
        }
        return true;
    }
    // sequential write of uncompressed bytes
    void write(const char* p, size_t n) {
        if (fd_ < 0) { fwrite(p, 1, n, f_); return; }
        if (gz_) {
            std::string z;
            if (n && gzip_member(p, n, 1, z)) write_at(reserve(z.size()), z.data(), z.size());
            return;
        }
        write_at(reserve(n), p, n);
    }
    void close() {
        if (f_) fflush(f_);
        if (fd_ >= 0) {
            if (gz_ && pos_ == 0) {  // gzopen + gzclose without data leave an empty member, not an empty file
                std::string z;
                if (gzip_member("", 0, 1, z)) write_at(reserve(z.size()), z.data(), z.size());
            }
            ::close(fd_);
        }
        f_ = nullptr;
        fd_ = -1;
    }
    ~Writer() { close(); }

   private:
    FILE* f_ = nullptr;  // stdout
    int fd_ = -1;
    bool gz_ = false;
    uint64_t pos_ = 0;
};

// One chunk of records of ONE input file, in the layout the C ABI takes.
struct Chunk {
    std::vector<char> ids;          // id lines concatenated (newline stripped)
    std::vector<uint32_t> id_off;   // n+1
    std::vector<char> qual_text;    // only used when a quality line's length differs from its sequence (kept verbatim)
    std::vector<uint32_t> qual_text_off;
    bool ragged_qual = false;
    size_t n = 0;
    // parallel reader: id lines stay in the mapped input file(s)
    std::vector<const char*> id_ptr;
    std::vector<uint32_t> id_len;
    bool ext_ids = false;
    void clear() {
        ids.clear(); id_off.assign(1, 0); qual_text.clear(); qual_text_off.assign(1, 0); ragged_qual = false; n = 0;
        id_ptr.clear(); id_len.clear(); ext_ids = false;
    }
};

inline size_t format_int(char* p, int v) {  // "%d"
    char tmp[16];
    int n = 0;
    unsigned int u = v < 0 ? 0u - (unsigned int)v : (unsigned int)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    size_t w = 0;
    if (v < 0) p[w++] = '-';
    while (n) p[w++] = tmp[--n];
    return w;
}

}  // namespace rcxio
