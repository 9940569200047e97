// CPU harness around rcorrector_b200/csrc/fastq_io.h (no CUDA): lets tests/test_host_io.py drive the host-side pieces of
// the drop-in binary — output naming, the block line reader (plain / gzip / mapped), the parallel newline index, "%d".
#include "../../rcorrector_b200/csrc/fastq_io.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    if (cmd == "name" && argc >= 5) {  // name <od> <type> <path>...
        for (int i = 4; i < argc; ++i) printf("%s\n", rcxio::output_path(argv[2], argv[i], atoi(argv[3])).c_str());
        return 0;
    }
    if (cmd == "lines" && argc == 3) {  // lines <file>: "<length> <had_nl> <first byte or -1>" per line
        rcxio::LineReader r;
        if (!r.open(argv[2])) return 1;
        const char* p; size_t n; bool nl;
        while (r.getline(p, n, nl)) printf("%zu %d %d\n", n, nl ? 1 : 0, n ? (int)(unsigned char)p[0] : -1);
        return 0;
    }
    if (cmd == "memlines" && argc == 4) {  // memlines <file> <start offset>: the same over a mapped file from an offset
        rcxio::MappedFile mf;
        if (!mf.open(argv[2])) return 1;
        rcxio::LineReader r;
        r.open_mem(mf.data, mf.size, (size_t)atoll(argv[3]));
        const char* p; size_t n; bool nl;
        while (r.getline(p, n, nl)) printf("%zu %d %d\n", n, nl ? 1 : 0, n ? (int)(unsigned char)p[0] : -1);
        printf("tell %zu\n", r.tell());
        return 0;
    }
    if (cmd == "newlines" && argc == 6) {  // newlines <file> <lo> <hi (0 = size)> <threads>: positions relative to lo
        rcxio::MappedFile mf;
        if (!mf.open(argv[2])) return 1;
        size_t lo = (size_t)atoll(argv[3]), hi = (size_t)atoll(argv[4]);
        if (hi == 0) hi = mf.size;
        std::vector<uint32_t> out;
        rcxio::index_newlines(mf.data, lo, hi, atoi(argv[5]), out);
        for (uint32_t v : out) printf("%u\n", v);
        return 0;
    }
    if (cmd == "int") {  // int <v>...
        char b[32];
        for (int i = 2; i < argc; ++i) { size_t n = rcxio::format_int(b, atoi(argv[i])); printf("%.*s\n", (int)n, b); }
        return 0;
    }
    if (cmd == "gzwrite" && argc == 4) {  // gzwrite <in> <out path>: copy through rcxio::Writer (gz level 1 when the name ends in gz)
        rcxio::MappedFile mf;
        if (!mf.open(argv[2])) return 1;
        rcxio::Writer w;
        if (!w.open(argv[3], false)) return 1;
        w.write(mf.data, mf.size);
        w.close();
        return 0;
    }
    return 2;
}
