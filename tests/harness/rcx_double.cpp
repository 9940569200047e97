// rcx_double.cpp — TEST DOUBLE of librcx.so for CPU tests of the drop-in binary's HOST side.
//
// TEST INFRASTRUCTURE ONLY.  It is never built by the product Makefiles, never installed and never loaded by the product:
// tests/test_binary_host_cpu.py compiles rcorrector_b200/csrc/rcorrector_main.cpp against THIS file (instead of
// librcx.so) into a temporary directory, so that everything of the binary that is not a CUDA kernel — argv grammar,
// dump ingest, badQualityThreshold, output naming, the parallel / serial FASTQ readers, the three-stage pipeline, the
// per-GPU sharding of a batch, the formatter and the gzip writer — can be checked on a machine without a GPU against
// the reference's bytes.  The "device" behind the C ABI here is the CPU oracle (oracle/librcx_oracle.so), used as the
// checker it is.  The product has no CPU path: the real rcx_create() fails without a CUDA device.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rcx.h"
#include "../../oracle/rcx_oracle.h"

struct rcx_ctx {
    int device = 0;
    orc_table* table = nullptr;
    int k = 0;
    orc_params params{};
    bool have_params = false;
};

static thread_local std::string g_err;

extern "C" int rcx_abi_version(void) { return RCX_ABI_VERSION; }
extern "C" const char* rcx_last_error(void) { return g_err.c_str(); }

extern "C" int rcx_create(int device, rcx_ctx** out) {
    if (!out || device < 0) { g_err = "rcx_create (double): bad argument"; return RCX_ERR_ARG; }
    *out = new rcx_ctx();
    (*out)->device = device;  // any index: the double stands in for as many "GPUs" as -gpus asks for
    return RCX_OK;
}
extern "C" void rcx_destroy(rcx_ctx* c) {
    if (!c) return;
    if (c->table) orc_table_free(c->table);
    delete c;
}

extern "C" int rcx_table_build(rcx_ctx* c, int k, const uint64_t* fwd_codes, const int32_t* counts, uint64_t n, int src_on_device) {
    if (!c || src_on_device || k < 2 || k > RCX_MAX_K) { g_err = "rcx_table_build (double): bad argument"; return RCX_ERR_ARG; }
    if (c->table) orc_table_free(c->table);
    c->table = orc_table_new(k);
    c->k = k;
    orc_table_put(c->table, fwd_codes, counts, n);
    return RCX_OK;
}

extern "C" int rcx_table_build_multi(rcx_ctx** ctxs, int nctx, int k, const uint64_t* fwd_codes, const int32_t* counts, uint64_t n) {
    for (int g = 0; g < nctx; ++g) {
        const int rc = rcx_table_build(ctxs[g], k, fwd_codes, counts, n, 0);
        if (rc) return rc;
    }
    return RCX_OK;
}

// -countkmers on the CPU (exact canonical counts; the product counts on the GPU, rcx_count.cu)
static std::unordered_map<uint64_t, uint32_t> g_cnt;
static std::vector<uint64_t> g_kept_codes;
static std::vector<int32_t> g_kept_counts;
static int g_cnt_k = 0;
extern "C" int rcx_count_kmers_begin(rcx_ctx* c, int k, uint64_t) {
    if (!c || k < 2 || k > RCX_MAX_K) { g_err = "rcx_count_kmers_begin (double): bad argument"; return RCX_ERR_ARG; }
    g_cnt.clear(); g_cnt_k = k;
    return RCX_OK;
}
extern "C" int rcx_count_kmers_add(rcx_ctx*, int64_t n, const char* seq, const uint32_t* off) {
    const int k = g_cnt_k;
    const uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    for (int64_t r = 0; r < n; ++r) {
        uint64_t fw = 0, rv = 0;
        int run = 0;
        for (uint32_t i = off[r]; i < off[r + 1]; ++i) {
            const char ch = seq[i];
            const int v = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1;
            fw = ((fw << 2) | (uint64_t)(v & 3)) & mask;
            rv = (rv >> 2) | ((uint64_t)(3 - (v & 3)) << (2 * (k - 1)));
            run = v >= 0 ? run + 1 : 0;
            if (run >= k) ++g_cnt[fw < rv ? fw : rv];
        }
    }
    return RCX_OK;
}
extern "C" int rcx_count_kmers_finish(rcx_ctx* c, int32_t min_count, uint64_t* kept, uint64_t* distinct, uint64_t* instances) {
    g_kept_codes.clear(); g_kept_counts.clear();
    uint64_t inst = 0;
    for (auto& kv : g_cnt) {
        inst += kv.second;
        if ((int64_t)kv.second >= min_count) { g_kept_codes.push_back(kv.first); g_kept_counts.push_back((int32_t)kv.second); }
    }
    if (kept) *kept = g_kept_codes.size();
    if (distinct) *distinct = g_cnt.size();
    if (instances) *instances = inst;
    return rcx_table_build(c, g_cnt_k, g_kept_codes.data(), g_kept_counts.data(), g_kept_codes.size(), 0);
}
extern "C" int rcx_count_kmers_export(rcx_ctx*, uint64_t* codes_out, int32_t* counts_out, uint64_t n) {
    if (n > g_kept_codes.size()) return RCX_ERR_ARG;
    if (codes_out) memcpy(codes_out, g_kept_codes.data(), n * sizeof(uint64_t));
    if (counts_out) memcpy(counts_out, g_kept_counts.data(), n * sizeof(int32_t));
    return RCX_OK;
}

extern "C" int rcx_set_params(rcx_ctx* c, const rcx_params* p) {
    if (!c || !p || !c->table || p->k != c->k) { g_err = "rcx_set_params (double): bad argument or no table"; return RCX_ERR_STATE; }
    c->params.k = p->k;
    c->params.max_fix_per_k = p->max_fix_per_k;
    c->params.error_rate = p->error_rate;
    c->params.bad_qual = p->bad_qual;
    c->have_params = true;
    return RCX_OK;
}

extern "C" int rcx_correct_batch(rcx_ctx* c, int mode, int64_t n, const char* seq_in, const char* qual, const uint32_t* off, char* seq_out,
                                 int32_t* correction, int32_t* l, int32_t* m, int32_t* h) {
    if (!c || !c->table || !c->have_params) { g_err = "rcx_correct_batch (double): no table or params"; return RCX_ERR_STATE; }
    if (n < 0 || (mode != RCX_SINGLE && mode != RCX_INTERLEAVED) || (mode == RCX_INTERLEAVED && (n & 1)) || (n > 0 && off[0] != 0)) {
        g_err = "rcx_correct_batch (double): bad n, mode or off[0]";
        return RCX_ERR_ARG;
    }
    if (n == 0) return RCX_OK;
    if (getenv("RCX_DOUBLE_PASSTHROUGH")) {  // host-throughput measurements (tools/host_rate.sh): no correction at all
        if (seq_out != seq_in) memcpy(seq_out, seq_in, off[n]);
        memset(correction, 0, (size_t)n * 4); memset(l, 0, (size_t)n * 4); memset(m, 0, (size_t)n * 4); memset(h, 0, (size_t)n * 4);
        return RCX_OK;
    }
    // the oracle writes seq_out from seq_in; the binary passes the same buffer for both
    std::vector<char> in(seq_in, seq_in + off[n]);
    orc_correct_batch(c->table, &c->params, mode == RCX_INTERLEAVED ? ORC_INTERLEAVED : ORC_SINGLE, n, in.data(), qual, off, seq_out,
                      correction, l, m, h, 2);
    return RCX_OK;
}

extern "C" int rcx_sync(rcx_ctx*) { return RCX_OK; }
extern "C" int rcx_get_stats(rcx_ctx* c, rcx_stats* out) {
    if (!c || !out) return RCX_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (const char* e = getenv("RCX_DOUBLE_UNSUPPORTED")) out->unsupported = (uint64_t)atoll(e);  // tests: the binary must stop
    return RCX_OK;
}
extern "C" void* rcx_host_alloc(uint64_t bytes) { return malloc(bytes ? bytes : 1); }
extern "C" void rcx_host_free(void* p) { free(p); }

// main.cpp:310-358 over the first n records: the 4 last-base variants of each record (ShiftRight(1) + Append(b)), records
// whose best variant has a count >= 1000, the first 100000 of them, ratio second / best, the wk quantile.
extern "C" int rcx_estimate_error_rate(rcx_ctx* c, const uint64_t* fwd_codes, uint64_t n, double wk, double* out) {
    if (!c || !c->table || !out) { g_err = "rcx_estimate_error_rate (double): bad argument"; return RCX_ERR_STATE; }
    std::vector<double> rates;
    for (uint64_t i = 0; i < n && rates.size() < 100000; ++i) {
        int best = 0, second = 0;
        for (int b = 0; b < 4; ++b) {
            const int v = orc_table_count(c->table, (fwd_codes[i] & ~3ull) | (uint64_t)b);
            if (v > best) { second = best; best = v; }
            else if (v > second) second = v;
        }
        if (best < 1000) continue;
        rates.push_back((double)second / (double)best);
    }
    double er = 0;
    const size_t kq = rates.size();
    if (kq > 0) {
        std::sort(rates.begin(), rates.end());
        rates.push_back(rates.back());
        er = rates[(size_t)(int)(kq * wk)];
    }
    if (er == 0 || kq < 100) er = 0.01;
    *out = er;
    return RCX_OK;
}
