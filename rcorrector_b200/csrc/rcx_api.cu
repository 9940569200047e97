// rcx_api.cu — the extern "C" boundary of librcx.so (include/rcx.h).
//
// One rcx_ctx = one GPU: it owns the k-mer count table in HBM, a stream, the
// intermediate buffers of the hot path (counts[], thresholds, flags) and grow-
// only staging buffers for host batches.  There is no CPU fallback anywhere in
// this file: every entry point either runs the sm_100a kernels or fails.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rcx.h"
#include "rcx_internal.h"

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(_e == cudaErrorMemoryAllocation ? RCX_ERR_NOMEM : RCX_ERR_CUDA,             \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
    } while (0)

// device allocations of one call that must not outlive an early return
struct DevGuard {
    std::vector<void**> ptrs;
    void watch(void** p) { ptrs.push_back(p); }
    void release_all() {
        for (void** p : ptrs)
            if (*p) { cudaFree(*p); *p = nullptr; }
    }
    ~DevGuard() { release_all(); }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// One pipeline slot = everything one batch (or one chunk of a host batch) needs on the device: a stream, the staging
// buffers, the intermediates of the hot path and the events that time it.  Host batches are cut into chunks that cycle
// through the slots so that the PCIe copies of one chunk overlap the kernels of another.  Buffers grow on first use, so
// a slot that is never used costs nothing.
#define RCX_NSLOTS 6
struct RcxSlot {
    cudaStream_t stream = nullptr;
    DevBuf<char> d_seq, d_qual;
    DevBuf<uint32_t> d_off;
    DevBuf<int32_t> d_counts, d_thr, d_corr, d_l, d_m, d_h;
    DevBuf<uint8_t> d_flags;
    DevBuf<uint32_t> d_worklist, d_worklist2;
    DevBuf<int32_t> d_state;
    DevBuf<signed char> d_fixbuf, d_dfs_path;
    DevBuf<unsigned char> d_posbuf;
    DevBuf<unsigned int> d_dfs_stack;
    // compact host interface: quality bits in, gathered edits out
    DevBuf<unsigned char> d_qbits, d_temp;
    DevBuf<uint32_t> d_scan;
    DevBuf<uint16_t> d_edits;
    bool edits_pending = false;  // the chunk's edits still wait on the device for their place in the caller's stream
    unsigned long long* d_counters = nullptr;
    unsigned long long* h_counters = nullptr;  // pinned landing area of the counters
    cudaEvent_t ev[6]{};
    cudaEvent_t kev[7]{};  // per-kernel marks inside the correction pipeline
    bool kev_valid = false;
    bool pending = false;  // work enqueued whose counters / timings have not been harvested yet
    int launches = 0;
    void release() {
        d_seq.release(); d_qual.release(); d_off.release(); d_counts.release(); d_thr.release(); d_corr.release(); d_l.release();
        d_m.release(); d_h.release(); d_flags.release(); d_worklist.release(); d_worklist2.release(); d_state.release();
        d_fixbuf.release(); d_dfs_path.release(); d_posbuf.release(); d_dfs_stack.release();
        d_qbits.release(); d_temp.release(); d_scan.release(); d_edits.release();
    }
};

struct rcx_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;  // = slot[0].stream: table build, device-resident batches
    // table
    RcxEntry* entries_dev = nullptr;     // the table: nlines * 4 sectors of 32 bytes
    unsigned long long nlines = 0;
    unsigned long long entries = 0;      // distinct canonical k-mers stored
    unsigned long long contexts = 0;     // distinct (k-1)-mer contexts = occupied sectors
    unsigned long long* big_keys = nullptr;
    unsigned int* big_counts = nullptr;
    int nbig = 0;
    bool wide_counts = false;  // some count >= 2^29: the thread-per-read kernels keep 29 bits next to their flags, so the warp-per-read ones run
    int k = 0;
    // params
    rcx_params params{};
    bool have_params = false;
    RcxSlot slot[RCX_NSLOTS];
    int nslots = 3;  // slots a host batch cycles through (RCX_SLOTS)
    int chain = 0;   // RCX_CHAIN: 1 = a chunk's kernels start when the previous chunk has reached its finalize kernel, 2 = when it is done
    rcx_stats stats{};
    bool vh_mode = false;  // the next batch's search starts with batches over its very heavy items (picked from the last harvested batch)
    // k-mer counting (rcx_count_kmers_*): the counting table, and after finish() the kept records
    unsigned long long* cnt_keys = nullptr;
    unsigned int* cnt_counts = nullptr;
    unsigned long long cnt_cap = 0, cnt_occupied = 0, cnt_instances = 0;
    unsigned long long* cnt_stat = nullptr;  // device: [0] occupied slots, [1] k-mer instances
    int cnt_k = 0;
    unsigned long long* kept_codes = nullptr;
    int32_t* kept_counts = nullptr;
    unsigned long long kept_n = 0;
    // the compact call in progress: where the chunks' edits go (harvest appends them in chunk order)
    uint16_t* cur_edits = nullptr;
    uint64_t cur_edits_cap = 0, cur_edits_n = 0;
    bool cur_edits_overflow = false;
};

extern "C" int rcx_abi_version(void) { return RCX_ABI_VERSION; }
extern "C" const char* rcx_last_error(void) { return g_err.c_str(); }

extern "C" int rcx_create(int device, rcx_ctx** out) {
    if (!out) return fail(RCX_ERR_ARG, "rcx_create: out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(RCX_ERR_CUDA, std::string("rcx_create: no CUDA device (") + cudaGetErrorString(e) +
                                      "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(RCX_ERR_ARG, "rcx_create: device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    rcx_ctx* c = new rcx_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    for (auto& sl : c->slot) {
        CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaMalloc(&sl.d_counters, sizeof(unsigned long long) * RCX_C_COUNT));
        CUDA_TRY(cudaMemset(sl.d_counters, 0, sizeof(unsigned long long) * RCX_C_COUNT));
        CUDA_TRY(cudaHostAlloc(&sl.h_counters, sizeof(unsigned long long) * RCX_C_COUNT, cudaHostAllocDefault));
        memset(sl.h_counters, 0, sizeof(unsigned long long) * RCX_C_COUNT);
        for (auto& ev : sl.ev) CUDA_TRY(cudaEventCreate(&ev));
        for (auto& ev : sl.kev) CUDA_TRY(cudaEventCreate(&ev));
    }
    c->stream = c->slot[0].stream;
    if (const char* e = getenv("RCX_SLOTS")) c->nslots = std::max(2, std::min(RCX_NSLOTS, atoi(e)));
    if (const char* e = getenv("RCX_CHAIN")) c->chain = atoi(e);
    *out = c;
    return RCX_OK;
}

static void free_table(rcx_ctx* c);
static void count_release(rcx_ctx* c, bool kept_too);

extern "C" void rcx_destroy(rcx_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& sl : c->slot) if (sl.stream) cudaStreamSynchronize(sl.stream);
    free_table(c);
    count_release(c, true);
    for (auto& sl : c->slot) {
        sl.release();
        if (sl.d_counters) cudaFree(sl.d_counters);
        if (sl.h_counters) cudaFreeHost(sl.h_counters);
        for (auto& ev : sl.ev) if (ev) cudaEventDestroy(ev);
        for (auto& ev : sl.kev) if (ev) cudaEventDestroy(ev);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    delete c;
}

extern "C" void* rcx_stream(rcx_ctx* c) { return c ? (void*)c->stream : nullptr; }

extern "C" void* rcx_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        g_err = "rcx_host_alloc: cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}
extern "C" void rcx_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

static int harvest(rcx_ctx* c, RcxSlot& sl);

extern "C" int rcx_sync(rcx_ctx* c) {
    if (!c) return fail(RCX_ERR_ARG, "rcx_sync: null context");
    CUDA_TRY(cudaSetDevice(c->device));
    for (auto& sl : c->slot) CUDA_TRY(cudaStreamSynchronize(sl.stream));
    return RCX_OK;
}

static void free_table(rcx_ctx* c) {
    if (c->entries_dev) { cudaFree(c->entries_dev); c->entries_dev = nullptr; }
    if (c->big_keys) { cudaFree(c->big_keys); c->big_keys = nullptr; }
    if (c->big_counts) { cudaFree(c->big_counts); c->big_counts = nullptr; }
    c->nbig = 0;
    c->wide_counts = false;
    c->nlines = 0;
    c->entries = 0;
}

extern "C" int rcx_table_build(rcx_ctx* c, int k, const uint64_t* fwd_codes, const int32_t* counts, uint64_t n,
                               int src_on_device) {
    if (!c) return fail(RCX_ERR_ARG, "rcx_table_build: null context");
    if (k < 2 || k > RCX_MAX_K) return fail(RCX_ERR_ARG, "rcx_table_build: k must be in [2, 32]");
    if (n && (!fwd_codes || !counts)) return fail(RCX_ERR_ARG, "rcx_table_build: null input arrays");
    if (n >= 0xffffffffull) return fail(RCX_ERR_ARG, "rcx_table_build: more than 2^32-2 records in one table");
    CUDA_TRY(cudaSetDevice(c->device));
    free_table(c);
    c->k = k;
    // One 32-byte entry per distinct (k-1)-mer context.  A transcriptome / genome has about as many contexts as k-mers
    // (the de Bruijn graph is mostly linear paths); unrelated random k-mers have two each.  Start from contexts ~ n at
    // the target load and grow if the claim pass says otherwise.
    // Entries per sector.  A lookup that does not end at its home sector costs the whole warp a dependent second probe, and
    // with 32 lanes per warp even a 10 % chance per lane makes that the common case (measured, 10 M reads per step:
    // load 0.15 / 0.2 / 0.3 / 0.4 / 0.5 -> 61.0 / 62.9 / 67.1 / 71.7 / 79.0 ms).  HBM is what a B200 has plenty of: take
    // the sparsest of 0.2 / 0.3 / 0.4 whose table stays below a third of the free memory (163 / 109 / 82 bytes per k-mer).
    double load = 0.40;
    {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const double cand[3] = {0.20, 0.30, 0.40};
        for (double cl : cand) {
            load = cl;
            if ((double)n * 1.02 / cl * 32.0 <= (double)free_b / 3.0) break;
        }
    }
    if (const char* e = getenv("RCX_TABLE_LOAD")) load = atof(e);
    if (!(load >= 0.05 && load <= 0.70)) load = 0.40;
    unsigned long long nlines = (unsigned long long)((double)n * 1.02 / load / 4.0) + 64;

    // records: device-resident if they are (or fit), else streamed through a bounded staging area for every pass
    const unsigned long long* d_codes_all = nullptr;
    const int32_t* d_counts_all = nullptr;
    unsigned long long* own_codes = nullptr;
    int32_t* own_counts = nullptr;
    DevGuard guard;  // frees whatever is still set when the function returns (every CUDA_TRY below may return)
    guard.watch((void**)&own_codes);
    guard.watch((void**)&own_counts);
    const uint64_t chunk = 1ull << 24;
    if (src_on_device) {
        d_codes_all = (const unsigned long long*)fwd_codes;
        d_counts_all = counts;
    } else {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t m = (n * 12 <= free_b / 4) ? std::max<uint64_t>(n, 1) : std::min<uint64_t>(chunk, std::max<uint64_t>(n, 1));
        CUDA_TRY(cudaMalloc(&own_codes, m * sizeof(unsigned long long)));
        CUDA_TRY(cudaMalloc(&own_counts, m * sizeof(int32_t)));
        if (m >= n) {
            CUDA_TRY(cudaMemcpyAsync(own_codes, fwd_codes, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(own_counts, counts, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            d_codes_all = own_codes;
            d_counts_all = own_counts;
        }
    }
    // run `fn(codes, counts, len, first_index)` over all records
    auto for_records = [&](auto fn) -> int {
        if (d_codes_all) return fn(d_codes_all, d_counts_all, (unsigned long long)n, 0ull);
        for (uint64_t s = 0; s < n; s += chunk) {
            const uint64_t len = std::min<uint64_t>(chunk, n - s);
            CUDA_TRY(cudaMemcpyAsync(own_codes, fwd_codes + s, len * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(own_counts, counts + s, len * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            int rc = fn(own_codes, own_counts, (unsigned long long)len, (unsigned long long)s);
            if (rc) return rc;
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
        return RCX_OK;
    };
    unsigned long long* d_stat = nullptr;  // [0] contexts claimed, [1] abort flag, [2] distinct k-mers, [3] saturated counts
    guard.watch((void**)&d_stat);
    CUDA_TRY(cudaMalloc(&d_stat, 4 * sizeof(unsigned long long)));
    unsigned long long h_stat[4] = {0, 0, 0, 0};
    RcxEntry* e = nullptr;
    guard.watch((void**)&e);
    for (int attempt = 0;; ++attempt) {
        CUDA_TRY(cudaMalloc(&e, nlines * 4 * sizeof(RcxEntry)));
        CUDA_TRY(rcx_launch_table_init(e, nlines, c->stream));
        CUDA_TRY(cudaMemsetAsync(d_stat, 0, 4 * sizeof(unsigned long long), c->stream));
        const unsigned long long limit = (unsigned long long)(0.75 * 4.0 * (double)nlines);
        for (int home_only = 1; home_only >= 0; --home_only) {
            int rc = for_records([&](const unsigned long long* dc, const int32_t* dn, unsigned long long len, unsigned long long) -> int {
                CUDA_TRY(rcx_launch_table_claim(e, nlines, k, dc, dn, len, home_only, d_stat, limit, c->stream));
                return RCX_OK;
            });
            if (rc) return rc;
        }
        CUDA_TRY(cudaMemcpyAsync(h_stat, d_stat, sizeof h_stat, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const double fill = (double)h_stat[0] / (4.0 * (double)nlines);
        if (!h_stat[1] && (fill <= load * 1.5 || attempt >= 1)) break;  // (after one resize the size is what memory allows)
        // more contexts than assumed: size for what was seen (an aborted pass saw at least 75 % of the sectors)
        cudaFree(e);
        e = nullptr;
        const unsigned long long seen = h_stat[1] ? std::max<unsigned long long>(h_stat[0], 2 * (unsigned long long)n) : h_stat[0];
        nlines = (unsigned long long)((double)seen / load / 4.0) + 64;
        {   // unrelated k-mers have two contexts each: if the sparse size does not fit, accept a denser table (up to 0.7)
            size_t free_b = 0, total_b = 0;
            CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            double budget = 0.8 * (double)free_b;
            if (const char* eb = getenv("RCX_TABLE_MAX_BYTES")) budget = atof(eb);  // tests
            if ((double)nlines * 128.0 > budget) {
                const unsigned long long fit = (unsigned long long)(budget / 128.0);
                const unsigned long long need_l = (unsigned long long)((double)seen / 0.70 / 4.0) + 64;
                if (fit < need_l) return fail(RCX_ERR_NOMEM, "rcx_table_build: the k-mer table does not fit the free device memory");
                nlines = fit;
            }
        }
        if (attempt >= 4) return fail(RCX_ERR_NOMEM, "rcx_table_build: could not size the table");
    }
    // "last record wins" per count slot, one partition of the table at a time (the side array is as large as the
    // sectors it covers, so it is cut to what is free)
    const unsigned long long nsec = nlines * 4;
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    unsigned long long part = nsec;
    const unsigned long long afford = (unsigned long long)(free_b * 0.6) / 32;
    if (part > afford) part = std::max<unsigned long long>(afford, 1ull << 16);
    if (const char* e = getenv("RCX_TABLE_WIN_SECTORS")) {  // tests: force the partitioned path
        const unsigned long long f = strtoull(e, nullptr, 10);
        if (f >= 1024 && f < part) part = f;
    }
    unsigned int* d_win = nullptr;
    guard.watch((void**)&d_win);
    CUDA_TRY(cudaMalloc(&d_win, part * 32));
    const unsigned long long big_cap = 1ull << 20;
    unsigned long long* d_big_keys = nullptr;
    unsigned int* d_big_counts = nullptr;
    guard.watch((void**)&d_big_keys);
    guard.watch((void**)&d_big_counts);
    CUDA_TRY(cudaMalloc(&d_big_keys, big_cap * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&d_big_counts, big_cap * sizeof(unsigned int)));
    for (unsigned long long lo = 0; lo < nsec; lo += part) {
        const unsigned long long hi = std::min(nsec, lo + part);
        CUDA_TRY(cudaMemsetAsync(d_win, 0, (hi - lo) * 32, c->stream));
        for (int write = 0; write < 2; ++write) {
            int rc = for_records([&](const unsigned long long* dc, const int32_t* dn, unsigned long long len, unsigned long long first) -> int {
                CUDA_TRY(rcx_launch_table_fill(e, nlines, k, dc, dn, len, first, d_win, lo, hi, write, d_stat, d_big_keys, d_big_counts, big_cap,
                                               c->stream));
                return RCX_OK;
            });
            if (rc) return rc;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(h_stat, d_stat, sizeof h_stat, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h_stat[3] > big_cap) return fail(RCX_ERR_ARG, "rcx_table_build: more than 2^20 k-mers with a count >= 2^24-1");
    // exact counts of the saturated 24-bit fields: sorted side list, binary-searched on the device
    c->nbig = (int)h_stat[3];
    if (c->nbig > 0) {
        std::vector<unsigned long long> bk(c->nbig);
        std::vector<unsigned int> bc(c->nbig);
        CUDA_TRY(cudaMemcpy(bk.data(), d_big_keys, c->nbig * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(bc.data(), d_big_counts, c->nbig * sizeof(unsigned int), cudaMemcpyDeviceToHost));
        std::vector<int> order(c->nbig);
        for (int i = 0; i < c->nbig; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return bk[a] < bk[b]; });
        std::vector<unsigned long long> sk(c->nbig);
        std::vector<unsigned int> sc(c->nbig);
        for (int i = 0; i < c->nbig; ++i) { sk[i] = bk[order[i]]; sc[i] = bc[order[i]]; }
        for (int i = 0; i < c->nbig; ++i) if (sc[i] >= (1u << 29)) c->wide_counts = true;
        CUDA_TRY(cudaMemcpy(d_big_keys, sk.data(), c->nbig * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d_big_counts, sc.data(), c->nbig * sizeof(unsigned int), cudaMemcpyHostToDevice));
    }
    c->big_keys = d_big_keys;
    c->big_counts = d_big_counts;
    c->entries_dev = e;
    d_big_keys = nullptr; d_big_counts = nullptr; e = nullptr;  // owned by the context now (the guard frees the temporaries)
    c->nlines = nlines;
    c->entries = h_stat[2];
    c->contexts = h_stat[0];
    return RCX_OK;
}

// The same table on several GPUs of one host (the drop-in binary's -gpus N): the (code, count) arrays cross PCIe ONCE, to
// the first context's GPU, and travel on device to device (binomial tree of peer copies: NVLink / NVSwitch where the GPUs
// have it); every GPU then inserts locally, in parallel.
extern "C" int rcx_table_build_multi(rcx_ctx** ctxs, int nctx, int k, const uint64_t* fwd_codes, const int32_t* counts, uint64_t n) {
    if (!ctxs || nctx < 1) return fail(RCX_ERR_ARG, "rcx_table_build_multi: no contexts");
    for (int g = 0; g < nctx; ++g)
        if (!ctxs[g]) return fail(RCX_ERR_ARG, "rcx_table_build_multi: null context");
    if (nctx == 1 || n == 0) {
        for (int g = 0; g < nctx; ++g) {
            const int rc = rcx_table_build(ctxs[g], k, fwd_codes, counts, n, 0);
            if (rc) return rc;
        }
        return RCX_OK;
    }
    if (!fwd_codes || !counts) return fail(RCX_ERR_ARG, "rcx_table_build_multi: null input arrays");
    std::vector<unsigned long long*> d_codes(nctx, nullptr);
    std::vector<int32_t*> d_counts(nctx, nullptr);
    auto release = [&]() {
        for (int g = 0; g < nctx; ++g) {
            cudaSetDevice(ctxs[g]->device);
            if (d_codes[g]) cudaFree(d_codes[g]);
            if (d_counts[g]) cudaFree(d_counts[g]);
        }
    };
    bool staged = true;
    for (int g = 0; g < nctx && staged; ++g) {
        cudaSetDevice(ctxs[g]->device);
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || n * 12 > free_b / 4 ||  // keep room for the table and its build
            cudaMalloc(&d_codes[g], n * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&d_counts[g], n * sizeof(int32_t)) != cudaSuccess)
            staged = false;
        for (int p = 0; p < nctx && staged; ++p)
            if (p != g) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, ctxs[g]->device, ctxs[p]->device);
                if (can && cudaDeviceEnablePeerAccess(ctxs[p]->device, 0) != cudaSuccess) cudaGetLastError();  // (already enabled)
            }
    }
    cudaGetLastError();
    if (!staged) {  // records too large to stage next to the table: every GPU streams them from the host
        release();
        std::vector<std::thread> th;
        std::vector<int> rcs(nctx, 0);
        std::vector<std::string> errs(nctx);
        for (int g = 0; g < nctx; ++g)
            th.emplace_back([&, g]() { rcs[g] = rcx_table_build(ctxs[g], k, fwd_codes, counts, n, 0); if (rcs[g]) errs[g] = g_err; });
        for (auto& t : th) t.join();
        for (int g = 0; g < nctx; ++g) if (rcs[g]) return fail(rcs[g], errs[g]);
        return RCX_OK;
    }
    cudaError_t e = cudaSetDevice(ctxs[0]->device);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_codes[0], fwd_codes, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctxs[0]->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_counts[0], counts, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctxs[0]->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctxs[0]->stream);
    for (int have = 1; have < nctx && e == cudaSuccess; have *= 2) {  // GPUs [0, have) hold the records: each feeds one more
        for (int g = 0; g < have && g + have < nctx && e == cudaSuccess; ++g) {
            const int src = g, dst = g + have;
            e = cudaSetDevice(ctxs[src]->device);
            if (e == cudaSuccess) e = cudaMemcpyPeerAsync(d_codes[dst], ctxs[dst]->device, d_codes[src], ctxs[src]->device, n * sizeof(uint64_t), ctxs[src]->stream);
            if (e == cudaSuccess) e = cudaMemcpyPeerAsync(d_counts[dst], ctxs[dst]->device, d_counts[src], ctxs[src]->device, n * sizeof(int32_t), ctxs[src]->stream);
        }
        for (int g = 0; g < have && e == cudaSuccess; ++g) {
            e = cudaSetDevice(ctxs[g]->device);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctxs[g]->stream);
        }
    }
    if (e != cudaSuccess) {
        release();
        return fail(RCX_ERR_CUDA, std::string("rcx_table_build_multi: ") + cudaGetErrorString(e));
    }
    std::vector<std::thread> th;
    std::vector<int> rcs(nctx, 0);
    std::vector<std::string> errs(nctx);
    for (int g = 0; g < nctx; ++g)
        th.emplace_back([&, g]() {
            rcs[g] = rcx_table_build(ctxs[g], k, (const uint64_t*)d_codes[g], d_counts[g], n, 1);
            if (rcs[g]) errs[g] = g_err;
        });
    for (auto& t : th) t.join();
    release();
    for (int g = 0; g < nctx; ++g) if (rcs[g]) return fail(rcs[g], errs[g]);
    return RCX_OK;
}

// ------------------------------------------------------------------------------------------------ k-mer counting (N3)
static void count_release(rcx_ctx* c, bool kept_too) {
    if (c->cnt_keys) { cudaFree(c->cnt_keys); c->cnt_keys = nullptr; }
    if (c->cnt_counts) { cudaFree(c->cnt_counts); c->cnt_counts = nullptr; }
    if (c->cnt_stat) { cudaFree(c->cnt_stat); c->cnt_stat = nullptr; }
    c->cnt_cap = c->cnt_occupied = c->cnt_instances = 0;
    if (kept_too) {
        if (c->kept_codes) { cudaFree(c->kept_codes); c->kept_codes = nullptr; }
        if (c->kept_counts) { cudaFree(c->kept_counts); c->kept_counts = nullptr; }
        c->kept_n = 0;
    }
}
static int count_alloc(rcx_ctx* c, unsigned long long cap, unsigned long long** keys, unsigned int** counts) {
    *keys = nullptr; *counts = nullptr;
    CUDA_TRY(cudaMalloc(keys, cap * sizeof(unsigned long long)));
    cudaError_t e = cudaMalloc(counts, cap * sizeof(unsigned int));
    if (e != cudaSuccess) { cudaFree(*keys); *keys = nullptr; return fail(RCX_ERR_NOMEM, "rcx_count_kmers: the counting table does not fit the device memory"); }
    CUDA_TRY(rcx_launch_count_init(*keys, *counts, cap, c->stream));
    return RCX_OK;
}
extern "C" int rcx_count_kmers_begin(rcx_ctx* c, int k, uint64_t capacity_hint) {
    if (!c) return fail(RCX_ERR_ARG, "rcx_count_kmers_begin: null context");
    if (k < 2 || k > RCX_MAX_K) return fail(RCX_ERR_ARG, "rcx_count_kmers_begin: k must be in [2, 32]");
    CUDA_TRY(cudaSetDevice(c->device));
    count_release(c, true);
    c->cnt_k = k;
    c->cnt_cap = std::max<unsigned long long>(1ull << 20, 2 * (unsigned long long)capacity_hint + 1024);
    int rc = count_alloc(c, c->cnt_cap, &c->cnt_keys, &c->cnt_counts);
    if (rc) return rc;
    CUDA_TRY(cudaMalloc(&c->cnt_stat, 2 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(c->cnt_stat, 0, 2 * sizeof(unsigned long long), c->stream));
    return RCX_OK;
}
extern "C" int rcx_count_kmers_add(rcx_ctx* c, int64_t n, const char* seq, const uint32_t* off) {
    if (!c || !c->cnt_keys) return fail(RCX_ERR_STATE, "rcx_count_kmers_add: rcx_count_kmers_begin first");
    if (n < 0 || (n && (!seq || !off))) return fail(RCX_ERR_ARG, "rcx_count_kmers_add: bad arguments");
    if (n == 0) return RCX_OK;
    if (off[0] != 0) return fail(RCX_ERR_ARG, "rcx_count_kmers_add: off[0] must be 0");
    CUDA_TRY(cudaSetDevice(c->device));
    RcxSlot& sl = c->slot[0];
    int rc = harvest(c, sl);
    if (rc) return rc;
    // chunks of at most 2^28 bases: the table is kept at most half full BEFORE a chunk that could add one k-mer per base
    const int k = c->cnt_k;
    int64_t lo = 0;
    while (lo < n) {
        int64_t hi = lo;
        while (hi < n && (uint64_t)(off[hi + 1] - off[lo]) <= (1ull << 28)) ++hi;
        if (hi == lo) hi = lo + 1;
        const uint64_t b0 = off[lo], span = off[hi] - b0;
        const unsigned long long worst = c->cnt_occupied + span;  // every window a new k-mer
        if (2 * worst > c->cnt_cap) {  // grow by re-insertion
            unsigned long long ncap = c->cnt_cap;
            while (2 * worst > ncap) ncap *= 2;
            unsigned long long* nk = nullptr;
            unsigned int* nc = nullptr;
            rc = count_alloc(c, ncap, &nk, &nc);
            if (rc) return rc;
            CUDA_TRY(rcx_launch_count_rehash(c->cnt_keys, c->cnt_counts, c->cnt_cap, nk, nc, ncap, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            cudaFree(c->cnt_keys); cudaFree(c->cnt_counts);
            c->cnt_keys = nk; c->cnt_counts = nc; c->cnt_cap = ncap;
        }
        CUDA_TRY(sl.d_seq.ensure(span + 96));
        CUDA_TRY(sl.d_off.ensure((size_t)(hi - lo) + 1));
        CUDA_TRY(cudaMemcpyAsync(sl.d_seq.p, seq + b0, span, cudaMemcpyHostToDevice, sl.stream));
        CUDA_TRY(cudaMemcpyAsync(sl.d_off.p, off + lo, ((size_t)(hi - lo) + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, sl.stream));
        CUDA_TRY(rcx_launch_count_reads(sl.d_seq.p - b0, sl.d_off.p, hi - lo, k, c->cnt_keys, c->cnt_counts, c->cnt_cap, c->cnt_stat, c->sm_count, sl.stream));
        unsigned long long h_stat[2];
        CUDA_TRY(cudaMemcpyAsync(h_stat, c->cnt_stat, sizeof h_stat, cudaMemcpyDeviceToHost, sl.stream));
        CUDA_TRY(cudaStreamSynchronize(sl.stream));
        c->cnt_occupied = h_stat[0];
        c->cnt_instances = h_stat[1];
        lo = hi;
    }
    return RCX_OK;
}
extern "C" int rcx_count_kmers_finish(rcx_ctx* c, int32_t min_count, uint64_t* kept, uint64_t* distinct, uint64_t* instances) {
    if (!c || !c->cnt_keys) return fail(RCX_ERR_STATE, "rcx_count_kmers_finish: rcx_count_kmers_begin first");
    if (min_count < 1) min_count = 1;
    CUDA_TRY(cudaSetDevice(c->device));
    // stable selection in slot order, a window of slots at a time: count, then allocate, then scatter
    const unsigned long long win = 1ull << 27;
    unsigned int *flags = nullptr, *pos = nullptr;
    void* temp = nullptr;
    const size_t tb = rcx_count_select_temp_bytes(win);
    CUDA_TRY(cudaMalloc(&flags, win * sizeof(unsigned int)));
    CUDA_TRY(cudaMalloc(&pos, win * sizeof(unsigned int)));
    CUDA_TRY(cudaMalloc(&temp, tb));
    auto cleanup = [&]() { cudaFree(flags); cudaFree(pos); cudaFree(temp); };
    unsigned long long total = 0;
    std::vector<unsigned long long> base;
    for (int pass = 0; pass < 2; ++pass) {
        unsigned long long run = 0;
        size_t wi = 0;
        for (unsigned long long lo = 0; lo < c->cnt_cap; lo += win, ++wi) {
            const unsigned long long n = std::min(win, c->cnt_cap - lo);
            cudaError_t e = rcx_launch_count_select(c->cnt_keys, c->cnt_counts, lo, n, (unsigned int)min_count, flags, pos, temp, tb,
                                                    pass ? base[wi] : 0, pass ? c->kept_codes : nullptr, pass ? c->kept_counts : nullptr, c->stream);
            if (e != cudaSuccess) { cleanup(); return fail(RCX_ERR_CUDA, std::string("rcx_count_kmers_finish: ") + cudaGetErrorString(e)); }
            if (!pass) {
                unsigned int last_pos = 0, last_flag = 0;
                cudaMemcpyAsync(&last_pos, pos + (n - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream);
                cudaMemcpyAsync(&last_flag, flags + (n - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream);
                if (cudaStreamSynchronize(c->stream) != cudaSuccess) { cleanup(); return fail(RCX_ERR_CUDA, "rcx_count_kmers_finish: selection failed"); }
                base.push_back(run);
                run += (unsigned long long)last_pos + last_flag;
            }
        }
        if (!pass) {
            total = run;
            if (c->kept_codes) { cudaFree(c->kept_codes); c->kept_codes = nullptr; }
            if (c->kept_counts) { cudaFree(c->kept_counts); c->kept_counts = nullptr; }
            if (cudaMalloc(&c->kept_codes, std::max<unsigned long long>(total, 1) * sizeof(unsigned long long)) != cudaSuccess ||
                cudaMalloc(&c->kept_counts, std::max<unsigned long long>(total, 1) * sizeof(int32_t)) != cudaSuccess) {
                cleanup();
                return fail(RCX_ERR_NOMEM, "rcx_count_kmers_finish: the kept k-mers do not fit the device memory");
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cleanup();
    c->kept_n = total;
    if (kept) *kept = total;
    if (distinct) *distinct = c->cnt_occupied;
    if (instances) *instances = c->cnt_instances;
    const int k = c->cnt_k;
    count_release(c, false);  // the counting table makes room for the count table
    return rcx_table_build(c, k, (const uint64_t*)c->kept_codes, c->kept_counts, total, 1);
}
extern "C" int rcx_count_kmers_export(rcx_ctx* c, uint64_t* codes_out, int32_t* counts_out, uint64_t n) {
    if (!c) return fail(RCX_ERR_ARG, "rcx_count_kmers_export: null context");
    if (n > c->kept_n) return fail(RCX_ERR_ARG, "rcx_count_kmers_export: more records asked for than kept");
    if (n == 0) return RCX_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    if (codes_out) CUDA_TRY(cudaMemcpy(codes_out, c->kept_codes, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (counts_out) CUDA_TRY(cudaMemcpy(counts_out, c->kept_counts, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return RCX_OK;
}

static RcxTable table_of(const rcx_ctx* c) {
    RcxTable t;
    t.e = c->entries_dev;
    t.nlines = c->nlines;
    t.k = c->k;
    t.big_keys = c->big_keys;
    t.big_counts = c->big_counts;
    t.nbig = c->nbig;
    return t;
}

extern "C" int rcx_table_lookup(rcx_ctx* c, const uint64_t* fwd_codes, int32_t* counts_out, uint64_t n) {
    if (!c) return fail(RCX_ERR_ARG, "rcx_table_lookup: null context");
    if (!c->entries_dev) return fail(RCX_ERR_STATE, "rcx_table_lookup: no table built");
    if (n && (!fwd_codes || !counts_out)) return fail(RCX_ERR_ARG, "rcx_table_lookup: null arrays");
    CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t chunk = 1ull << 24;
    unsigned long long* d_codes = nullptr;
    int32_t* d_out = nullptr;
    const uint64_t m = std::min<uint64_t>(chunk, std::max<uint64_t>(n, 1));
    CUDA_TRY(cudaMalloc(&d_codes, m * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&d_out, m * sizeof(int32_t)));
    for (uint64_t s = 0; s < n; s += chunk) {
        const uint64_t len = std::min<uint64_t>(chunk, n - s);
        CUDA_TRY(cudaMemcpyAsync(d_codes, fwd_codes + s, len * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(rcx_launch_table_lookup(table_of(c), d_codes, d_out, len, c->stream));
        CUDA_TRY(cudaMemcpyAsync(counts_out + s, d_out, len * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    cudaFree(d_codes);
    cudaFree(d_out);
    return RCX_OK;
}

extern "C" int rcx_table_info(rcx_ctx* c, uint64_t* entries, uint64_t* slots, uint64_t* bytes) {
    if (!c) return fail(RCX_ERR_ARG, "rcx_table_info: null context");
    if (entries) *entries = c->entries;
    if (slots) *slots = c->nlines * 4;  // one context entry per 32-byte sector
    if (bytes) *bytes = c->nlines * 4 * sizeof(RcxEntry);
    return RCX_OK;
}

extern "C" int rcx_set_params(rcx_ctx* c, const rcx_params* p) {
    if (!c || !p) return fail(RCX_ERR_ARG, "rcx_set_params: null argument");
    if (p->k < 2 || p->k > RCX_MAX_K) return fail(RCX_ERR_ARG, "rcx_set_params: k must be in [2, 32]");
    if (p->max_fix_per_k < 0) return fail(RCX_ERR_ARG, "rcx_set_params: max_fix_per_k < 0");
    if (!(p->error_rate > 0)) return fail(RCX_ERR_ARG, "rcx_set_params: error_rate must be > 0");
    c->params = *p;
    c->have_params = true;
    return RCX_OK;
}

// ------------------------------------------------------------------------------------------------ hot path
static int check_ready(const rcx_ctx* c, const char* who) {
    if (!c) return fail(RCX_ERR_ARG, std::string(who) + ": null context");
    if (!c->entries_dev) return fail(RCX_ERR_STATE, std::string(who) + ": no table built (rcx_table_build)");
    if (!c->have_params) return fail(RCX_ERR_STATE, std::string(who) + ": parameters not set (rcx_set_params)");
    if (c->params.k != c->k) return fail(RCX_ERR_STATE, std::string(who) + ": params.k differs from the table's k");
    return RCX_OK;
}

static int scan_lengths(const uint32_t* off, int64_t n, int32_t* max_len) {
    uint32_t mx = 0;
    for (int64_t r = 0; r < n; ++r) {
        if (off[r + 1] < off[r]) return fail(RCX_ERR_ARG, "offsets are not non-decreasing");
        mx = std::max(mx, off[r + 1] - off[r]);
    }
    if (mx > RCX_MAX_READ_LENGTH) return fail(RCX_ERR_ARG, "a read is longer than RCX_MAX_READ_LENGTH (1023)");
    *max_len = (int32_t)mx;
    return RCX_OK;
}


// Enqueue the kernels of the hot path for one batch on `sl`.  The pointers may be "virtual" bases (a slot buffer minus the
// chunk's first offset): the kernels index them with the ABSOLUTE offsets of `d_off`.
static int launch_pipeline(rcx_ctx* c, RcxSlot& sl, int mode, int64_t n, char* d_seq, const char* d_qual, const uint32_t* d_off,
                           uint64_t first_base, uint64_t end_base, int32_t max_len, int32_t* d_corr, int32_t* d_l, int32_t* d_m,
                           int32_t* d_h, bool correct, const unsigned char* d_qbits = nullptr, int has_qual = 1, bool emit_edits = false) {
    const uint64_t span = end_base - first_base;
    const uint64_t lead = first_base & 15;  // keeps every virtual base 16-byte aligned
    CUDA_TRY(sl.d_counts.ensure(span + 32));
    CUDA_TRY(sl.d_thr.ensure((size_t)n + 2));
    CUDA_TRY(sl.d_flags.ensure((size_t)n + 2));
    if (correct) {
        CUDA_TRY(sl.d_worklist.ensure((size_t)n + 2));
        CUDA_TRY(sl.d_worklist2.ensure(2 * (size_t)n + 4));  // second-pass list, then (from n) vh mode's list of the very heavy items
        CUDA_TRY(sl.d_state.ensure((size_t)n * rcx_state_words(max_len, c->k) + 64));
        CUDA_TRY(sl.d_fixbuf.ensure(2 * (span + 96)));  // dense fix[] (1 byte per base) or 16-bit edit lists (a count + up to len - 1 edits per read)
        CUDA_TRY(sl.d_posbuf.ensure(span + 96));
    }
    CUDA_TRY(cudaMemsetAsync(sl.d_counters, 0, sizeof(unsigned long long) * RCX_C_COUNT, sl.stream));
    RcxScanArgs sa;
    sa.t = table_of(c);
    sa.seq = d_seq; sa.off = d_off; sa.n = n; sa.max_len = max_len;
    sa.counts = sl.d_counts.p + lead - first_base; sa.thr = sl.d_thr.p; sa.flags = sl.d_flags.p; sa.counters = sl.d_counters;
    sa.triage = correct ? 1 : 0; sa.mode = correct ? mode : RCX_SINGLE; sa.error_rate = c->params.error_rate;
    sa.correction = d_corr; sa.l = d_l; sa.m = d_m; sa.h = d_h; sa.worklist = correct ? sl.d_worklist.p : nullptr;
    sa.total_bases = end_base; sa.max_fix_per_k = c->params.max_fix_per_k;
    sa.state = correct ? sl.d_state.p : nullptr;
    // ordinary read lengths: one thread per read (scan + thresholds + triage + first-pass preparation in one kernel);
    // long reads, RCX_SCAN_WARP=1, or a table with counts of 2^29 and more: the warp-per-read kernels
    const bool thread_scan = rcx_scan_thread_supported(max_len, c->k, d_seq) && !getenv("RCX_SCAN_WARP") && !c->wide_counts;
    // heavy = at least 2k untrusted k-windows (one isolated error explains k of them); RCX_HEAVY_WINDOWS overrides, 0 = off
    sa.heavy_windows = 2 * c->k;
    if (const char* e = getenv("RCX_HEAVY_WINDOWS")) sa.heavy_windows = atoi(e);
    // very heavy = at least three quarters of the longest read's windows untrusted: always counted; in vh mode (the previous batch
    // counted a quarter of its work items as such; RCX_VH_MODE=0/1 forces) also listed for the search kernel's opening batches
    sa.vheavy_windows = std::max(sa.heavy_windows + 1, 3 * (max_len - c->k + 1) / 4);
    sa.vh_mode = c->vh_mode ? 1 : 0;
    if (const char* e = getenv("RCX_VH_MODE")) sa.vh_mode = atoi(e) != 0;
    sa.vhlist = correct ? sl.d_worklist2.p + n : nullptr;
    if (!(thread_scan && rcx_finalize_thread_supported(max_len, c->k)))
        sa.heavy_windows = sa.vheavy_windows = sa.vh_mode = 0;  // (the warp-per-read kernels keep one list)
    sa.posbuf = correct ? sl.d_posbuf.p + lead - first_base : nullptr;
    sa.fixbuf = correct ? sl.d_fixbuf.p + lead - first_base : nullptr;
    sa.fixlist = correct ? reinterpret_cast<unsigned short*>(sl.d_fixbuf.p) + lead - first_base : nullptr;
    // ordinary read lengths: one thread per read (scan + thresholds + triage + first-pass preparation in one kernel);
    // long reads (or RCX_SCAN_WARP=1): the warp-per-read kernels
    sa.compact_fix = (thread_scan && rcx_finalize_thread_supported(max_len, c->k)) ? 1 : 0;
    // RCX_ROW24=1 (experimental, off by default): 24-bit count rows in the scan's 64-read tiles (reads of more than 128 bases)
    bool row24 = false;
    if (const char* e = getenv("RCX_ROW24")) row24 = atoi(e) != 0;
    CUDA_TRY(cudaEventRecord(sl.ev[1], sl.stream));
    if (thread_scan) CUDA_TRY(rcx_launch_scan_thread(sa, c->sm_count, sl.stream, row24));
    else CUDA_TRY(rcx_launch_scan(sa, c->sm_count, sl.stream));
    CUDA_TRY(cudaEventRecord(sl.ev[2], sl.stream));
    sl.launches = 1;
    sl.kev_valid = false;
    if (correct) {
        RcxCorrectArgs ca;
        ca.t = sa.t;
        ca.seq = d_seq; ca.qual = d_qual; ca.off = d_off; ca.n = n; ca.max_len = max_len; ca.mode = mode;
        ca.max_fix_per_k = c->params.max_fix_per_k; ca.error_rate = c->params.error_rate; ca.bad_qual = c->params.bad_qual;
        ca.qbits = d_qbits; ca.has_qual = has_qual; ca.emit_edits = emit_edits ? 1 : 0;
        ca.counts = sa.counts; ca.thr = sl.d_thr.p; ca.flags = sl.d_flags.p;
        ca.correction = d_corr; ca.l = d_l; ca.m = d_m; ca.h = d_h; ca.counters = sl.d_counters;
        ca.worklist = sl.d_worklist.p; ca.worklist2 = sl.d_worklist2.p; ca.vhlist = sl.d_worklist2.p + n; ca.vh_mode = sa.vh_mode; ca.state = sl.d_state.p; ca.fixbuf = sa.fixbuf; ca.pass = 0;
        int dfs_grid = 0;
        rcx_dfs_geometry(c->sm_count, max_len, c->k, c->params.max_fix_per_k, &dfs_grid, &ca.dfs_depth, &ca.dfs_lpad);
        CUDA_TRY(sl.d_dfs_stack.ensure(rcx_dfs_stack_words(dfs_grid, ca.dfs_depth)));
        CUDA_TRY(sl.d_dfs_path.ensure(rcx_dfs_path_bytes(dfs_grid, ca.dfs_lpad)));
        ca.posbuf = sa.posbuf; ca.dfs_stack = sl.d_dfs_stack.p; ca.dfs_path = sl.d_dfs_path.p; ca.fixlist = sa.fixlist;
        ca.compact_fix = (thread_scan && rcx_finalize_thread_supported(max_len, c->k)) ? 1 : 0;
        ca.dfs_batch = (thread_scan && max_len <= 128) ? 1 : 0;  // the scan sorted the light items within 128-read tiles
        if (const char* e = getenv("RCX_DFS_BATCH")) ca.dfs_batch = atoi(e);
        CUDA_TRY(rcx_launch_correct(ca, c->sm_count, dfs_grid, sl.stream, sl.kev, thread_scan));
        sl.kev_valid = true;
        sl.launches = thread_scan ? 5 : 6;  // scan (+ thresholds + triage [+ first preparation]), search, preparation + search of the second pass, finalize
        if (emit_edits) {  // applied edits of the changed reads, in record order (prefix sum + gather: 2 library + 1 own launch)
            const size_t tb = rcx_edits_temp_bytes(n);
            CUDA_TRY(sl.d_temp.ensure(tb));
            CUDA_TRY(sl.d_scan.ensure((size_t)n + 1));
            CUDA_TRY(sl.d_edits.ensure(span + 64));  // a read cannot have more edits than bases
            CUDA_TRY(rcx_launch_edits_gather(d_corr, d_off, n, ca.fixbuf, ca.compact_fix ? ca.fixlist : nullptr, sl.d_scan.p, sl.d_temp.p, tb, sl.d_edits.p, sl.d_counters, sl.stream));
            sl.launches += 1;
        }
    }
    CUDA_TRY(cudaEventRecord(sl.ev[3], sl.stream));
    return RCX_OK;
}

// add the counters and CUDA-event times of a finished slot to the statistics of the current call
static int harvest(rcx_ctx* c, RcxSlot& sl) {
    if (!sl.pending) return RCX_OK;
    CUDA_TRY(cudaStreamSynchronize(sl.stream));
    const unsigned long long* h = sl.h_counters;
    if (sl.edits_pending) {  // this chunk's edits follow those of the chunks before it (harvests run in chunk order)
        const uint64_t ne = h[RCX_C_NEDITS];
        if (c->cur_edits && c->cur_edits_n + ne <= c->cur_edits_cap) {
            if (ne) CUDA_TRY(cudaMemcpyAsync(c->cur_edits + c->cur_edits_n, sl.d_edits.p, ne * sizeof(uint16_t), cudaMemcpyDeviceToHost, sl.stream));
        } else {
            c->cur_edits_overflow = true;
        }
        c->cur_edits_n += ne;
        sl.edits_pending = false;
    }
    c->stats.scan_probes += h[RCX_C_SCAN_PROBES];
    c->stats.search_probes += h[RCX_C_SEARCH_PROBES];
    c->stats.rescan_probes += h[RCX_C_RESCAN_PROBES];
    c->stats.rescan_windows += h[RCX_C_RESCAN_WINDOWS];
    c->stats.search_nodes += h[RCX_C_SEARCH_NODES];
    c->stats.work_items += h[RCX_C_NWORK] + h[RCX_C_NWORKH];
    c->stats.heavy_items += h[RCX_C_NWORKH];
    c->stats.vheavy_items += h[RCX_C_NVH];
    if (h[RCX_C_NWORK] + h[RCX_C_NWORKH] >= 1024)  // (a batch large enough to say something about the next one)
        c->vh_mode = 4 * h[RCX_C_NVH] >= h[RCX_C_NWORK] + h[RCX_C_NWORKH];
    c->stats.second_pass += h[RCX_C_NWORK2];
    c->stats.unsupported += h[RCX_C_UNSUPPORTED];
    {   // test hook (tests/test_gpu_golden.py): RCX_TEST_FORCE_UNSUPPORTED=n flags the n-th harvested batch of the process
        static int forced = getenv("RCX_TEST_FORCE_UNSUPPORTED") ? atoi(getenv("RCX_TEST_FORCE_UNSUPPORTED")) : 0;
        if (forced > 0 && --forced == 0) c->stats.unsupported += 1;
    }
    c->stats.kernel_launches += (uint64_t)sl.launches;
    auto ms = [&](cudaEvent_t a, cudaEvent_t b) {
        float v = 0;
        if (cudaEventElapsedTime(&v, a, b) != cudaSuccess) { cudaGetLastError(); v = 0; }
        return v;
    };
    c->stats.ms_h2d += ms(sl.ev[0], sl.ev[1]);
    c->stats.ms_scan += ms(sl.ev[1], sl.ev[2]);
    c->stats.ms_correct += ms(sl.ev[2], sl.ev[3]);
    c->stats.ms_d2h += ms(sl.ev[4], sl.ev[5]);
    c->stats.ms_total += ms(sl.ev[0], sl.ev[5]);
    if (sl.kev_valid) {
        c->stats.ms_triage += ms(sl.kev[0], sl.kev[1]);
        c->stats.ms_prep += ms(sl.kev[1], sl.kev[2]) + ms(sl.kev[3], sl.kev[4]);
        c->stats.ms_dfs += ms(sl.kev[2], sl.kev[3]) + ms(sl.kev[4], sl.kev[5]);
        c->stats.ms_finalize += ms(sl.kev[5], sl.kev[6]);
    }
    sl.pending = false;
    return RCX_OK;
}
static int begin_call(rcx_ctx* c, uint64_t reads, uint64_t bases) {
    for (auto& sl : c->slot) {
        int rc = harvest(c, sl);
        if (rc) return rc;
    }
    c->stats = rcx_stats{};
    c->stats.reads = reads;
    c->stats.bases = bases;
    return RCX_OK;
}
// the tail every enqueue shares: counters to the pinned landing area, final event
static int finish_enqueue(RcxSlot& sl) {
    CUDA_TRY(cudaMemcpyAsync(sl.h_counters, sl.d_counters, sizeof(unsigned long long) * RCX_C_COUNT, cudaMemcpyDeviceToHost, sl.stream));
    CUDA_TRY(cudaEventRecord(sl.ev[5], sl.stream));
    sl.pending = true;
    return RCX_OK;
}

extern "C" int rcx_correct_batch_device(rcx_ctx* c, int mode, int64_t n, char* seq, const char* qual, const uint32_t* off,
                                        uint64_t total, int32_t max_len, int32_t* correction, int32_t* l, int32_t* m,
                                        int32_t* h) {
    int rc = check_ready(c, "rcx_correct_batch_device");
    if (rc) return rc;
    if (n < 0 || (mode != RCX_SINGLE && mode != RCX_INTERLEAVED)) return fail(RCX_ERR_ARG, "rcx_correct_batch_device: bad n or mode");
    if (mode == RCX_INTERLEAVED && (n & 1)) return fail(RCX_ERR_ARG, "rcx_correct_batch_device: interleaved batch needs an even n");
    if (n == 0) return RCX_OK;
    if (!seq || !qual || !off || !correction || !l || !m || !h) return fail(RCX_ERR_ARG, "rcx_correct_batch_device: null pointer");
    if (max_len < 1 || max_len > RCX_MAX_READ_LENGTH) return fail(RCX_ERR_ARG, "rcx_correct_batch_device: bad max_len");
    CUDA_TRY(cudaSetDevice(c->device));
    RcxSlot& sl = c->slot[0];
    // Asynchronous on purpose: a device batch enqueued behind another one does not wait for it on the host (the kernels are
    // stream-ordered; counters and event times then describe the LAST batch).  Waiting here put a host round trip between
    // the steps of a loop of device batches -- 8 ms per step with eight processes on one host (profiles/r02_n8_bench.json).
    if (sl.pending && cudaStreamQuery(sl.stream) == cudaSuccess) {  // the previous batch is already done: its counters cost nothing
        rc = harvest(c, sl);                                        // (and they pick the search kernel's mode for this batch)
        if (rc) return rc;
    }
    cudaGetLastError();
    sl.pending = false;
    rc = begin_call(c, (uint64_t)n, total);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(sl.ev[0], sl.stream));
    rc = launch_pipeline(c, sl, mode, n, seq, qual, off, 0, total, max_len, correction, l, m, h, true);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(sl.ev[4], sl.stream));
    return finish_enqueue(sl);
}

// Host batch: cut into chunks of whole records (pairs stay together) that cycle through the slots, so that the
// host -> device copy of chunk i+1 and the device -> host copy of chunk i-1 overlap the kernels of chunk i
// (needs pinned host buffers to be truly asynchronous: rcx_host_alloc).  Results do not depend on the chunking.
// `compact`: quality arrives as one bit per base (qbits; may be null when has_qual == 0), no sequence goes back; the applied
// edits are appended to c->cur_edits instead.
static int correct_host(rcx_ctx* c, const char* who, int mode, int64_t n, const char* seq_in, const char* qual, const unsigned char* qbits,
                        int has_qual, bool compact, const uint32_t* off, char* seq_out, int32_t* correction, int32_t* l, int32_t* m,
                        int32_t* h) {
    int rc = check_ready(c, who);
    if (rc) return rc;
    if (n < 0 || (mode != RCX_SINGLE && mode != RCX_INTERLEAVED)) return fail(RCX_ERR_ARG, std::string(who) + ": bad n or mode");
    if (mode == RCX_INTERLEAVED && (n & 1)) return fail(RCX_ERR_ARG, std::string(who) + ": interleaved batch needs an even n");
    if (n == 0) return RCX_OK;
    if (!seq_in || !off || !correction || !l || !m || !h || (!compact && (!qual || !seq_out)) || (compact && has_qual && !qbits))
        return fail(RCX_ERR_ARG, std::string(who) + ": null pointer");
    if (off[0] != 0) return fail(RCX_ERR_ARG, std::string(who) + ": off[0] must be 0");
    CUDA_TRY(cudaSetDevice(c->device));
    rc = begin_call(c, (uint64_t)n, off[n]);
    if (rc) return rc;
    // Chunks are sized in BASES: what a chunk costs (copies, kernels) follows its bases, not its records.  Measured per 10 M
    // 100-bp reads: 1 M / 1.5 M / 2 M / 3 M records -> 81.9 / 75.5 / 75.0 / 85.2 ms; per 6 M 150-bp reads: 1 M / 1.5 M / 2 M /
    // 3 M records -> 81.9 / 82.2 / 96.8 / 102.0 ms (profiles/r01_u_e2e_chunks.md).
    double chunk_bases = 208e6;  // = the 2 Mi-record chunks the 100-bp workload was tuned with
    if (const char* e = getenv("RCX_CHUNK_BASES")) chunk_bases = atof(e);
    int64_t chunk_reads = (int64_t)(chunk_bases / std::max(1.0, (double)off[n] / (double)n));
    if (const char* e = getenv("RCX_CHUNK_READS")) chunk_reads = atoll(e);
    if (chunk_reads < 2) chunk_reads = 2;
    chunk_reads &= ~(int64_t)1;
    // chunk schedule: the first and the last chunk are half-size -- their copies are the ones that cannot hide behind kernels
    // (more, smaller chunks cost more than they hide: every kernel launch ends with a tail of few busy blocks)
    std::vector<int64_t> cuts;
    {
        bool ramp = n >= 3 * chunk_reads;
        if (const char* e = getenv("RCX_CHUNK_RAMP")) ramp = ramp && atoi(e) != 0;
        const int64_t hlf = ramp ? (chunk_reads / 2) & ~(int64_t)1 : 0;
        int64_t lo = 0;
        cuts.push_back(0);
        if (ramp) cuts.push_back(lo += hlf);
        while (n - hlf - lo > 0) { lo = std::min<int64_t>(n - hlf, lo + chunk_reads); cuts.push_back(lo); }
        if (ramp) cuts.push_back(n);
    }
    int si = 0;
    RcxSlot* prev = nullptr;
    for (size_t ci = 0; ci + 1 < cuts.size(); ++ci) {
        const int64_t lo = cuts[ci], hi = cuts[ci + 1], cn = hi - lo;
        if (cn <= 0) continue;  // (an empty chunk must not advance the slot: the harvests below rely on the slot order)
        int32_t max_len = 0;
        rc = scan_lengths(off + lo, cn, &max_len);
        if (rc) return rc;
        if (max_len < 1) max_len = 1;
        RcxSlot& sl = c->slot[si];
        rc = harvest(c, sl);  // the slot's previous chunk (also makes its buffers reusable)
        if (rc) return rc;
        const uint64_t b0 = off[lo], b1 = off[hi], span = b1 - b0, lead = b0 & 15;
        const uint64_t q0 = b0 >> 3, q1 = (b1 + 7) >> 3;  // bytes of the quality bit plane that cover the chunk
        CUDA_TRY(sl.d_seq.ensure(span + 96));
        if (!compact) CUDA_TRY(sl.d_qual.ensure(span + 96));
        else if (qbits) CUDA_TRY(sl.d_qbits.ensure(q1 - q0 + 16));
        CUDA_TRY(sl.d_off.ensure((size_t)cn + 1));
        CUDA_TRY(sl.d_corr.ensure((size_t)cn));
        CUDA_TRY(sl.d_l.ensure((size_t)cn));
        CUDA_TRY(sl.d_m.ensure((size_t)cn));
        CUDA_TRY(sl.d_h.ensure((size_t)cn));
        CUDA_TRY(cudaEventRecord(sl.ev[0], sl.stream));
        CUDA_TRY(cudaMemcpyAsync(sl.d_seq.p + lead, seq_in + b0, span, cudaMemcpyHostToDevice, sl.stream));
        if (!compact) CUDA_TRY(cudaMemcpyAsync(sl.d_qual.p + lead, qual + b0, span, cudaMemcpyHostToDevice, sl.stream));
        else if (qbits) CUDA_TRY(cudaMemcpyAsync(sl.d_qbits.p, qbits + q0, q1 - q0, cudaMemcpyHostToDevice, sl.stream));
        CUDA_TRY(cudaMemcpyAsync(sl.d_off.p, off + lo, ((size_t)cn + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, sl.stream));
        if (prev && c->chain == 1 && prev->kev_valid) CUDA_TRY(cudaStreamWaitEvent(sl.stream, prev->kev[5], 0));
        if (prev && c->chain == 2) CUDA_TRY(cudaStreamWaitEvent(sl.stream, prev->ev[3], 0));
        prev = &sl;
        rc = launch_pipeline(c, sl, mode, cn, sl.d_seq.p + lead - b0, compact ? nullptr : sl.d_qual.p + lead - b0, sl.d_off.p, b0, b1, max_len,
                             sl.d_corr.p, sl.d_l.p, sl.d_m.p, sl.d_h.p, true, compact ? (qbits ? sl.d_qbits.p - q0 : nullptr) : nullptr,
                             compact ? (qbits ? has_qual : 0) : 1, compact);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(sl.ev[4], sl.stream));
        if (!compact) CUDA_TRY(cudaMemcpyAsync(seq_out + b0, sl.d_seq.p + lead, span, cudaMemcpyDeviceToHost, sl.stream));
        sl.edits_pending = compact;
        CUDA_TRY(cudaMemcpyAsync(correction + lo, sl.d_corr.p, (size_t)cn * 4, cudaMemcpyDeviceToHost, sl.stream));
        CUDA_TRY(cudaMemcpyAsync(l + lo, sl.d_l.p, (size_t)cn * 4, cudaMemcpyDeviceToHost, sl.stream));
        CUDA_TRY(cudaMemcpyAsync(m + lo, sl.d_m.p, (size_t)cn * 4, cudaMemcpyDeviceToHost, sl.stream));
        CUDA_TRY(cudaMemcpyAsync(h + lo, sl.d_h.p, (size_t)cn * 4, cudaMemcpyDeviceToHost, sl.stream));
        rc = finish_enqueue(sl);
        if (rc) return rc;
        si = (si + 1) % c->nslots;
    }
    for (int j = 0; j < c->nslots; ++j) {  // oldest chunk first: `si` is the slot the next chunk would have taken
        rc = harvest(c, c->slot[(si + j) % c->nslots]);
        if (rc) return rc;
    }
    for (int j = 0; j < c->nslots; ++j) CUDA_TRY(cudaStreamSynchronize(c->slot[j].stream));  // (the edits of the last chunks)
    return RCX_OK;
}

extern "C" int rcx_correct_batch(rcx_ctx* c, int mode, int64_t n, const char* seq_in, const char* qual, const uint32_t* off,
                                 char* seq_out, int32_t* correction, int32_t* l, int32_t* m, int32_t* h) {
    if (c) { c->cur_edits = nullptr; c->cur_edits_cap = c->cur_edits_n = 0; c->cur_edits_overflow = false; }
    return correct_host(c, "rcx_correct_batch", mode, n, seq_in, qual, nullptr, 1, false, off, seq_out, correction, l, m, h);
}

extern "C" int rcx_correct_batch_compact(rcx_ctx* c, int mode, int64_t n, const char* seq, const uint8_t* qual_bits, int has_qual,
                                         const uint32_t* off, int32_t* correction, int32_t* l, int32_t* m, int32_t* h, uint16_t* edits,
                                         uint64_t edits_cap, uint64_t* n_edits) {
    if (!c) return fail(RCX_ERR_ARG, "rcx_correct_batch_compact: null context");
    if (!edits && edits_cap) return fail(RCX_ERR_ARG, "rcx_correct_batch_compact: null edits buffer");
    c->cur_edits = edits; c->cur_edits_cap = edits_cap; c->cur_edits_n = 0; c->cur_edits_overflow = false;
    if (n_edits) *n_edits = 0;
    const int rc = correct_host(c, "rcx_correct_batch_compact", mode, n, seq, nullptr, qual_bits, has_qual ? 1 : 0, true, off, nullptr,
                                correction, l, m, h);
    c->cur_edits = nullptr;
    if (rc) return rc;
    if (n_edits) *n_edits = c->cur_edits_n;
    if (c->cur_edits_overflow)
        return fail(RCX_ERR_ARG, "rcx_correct_batch_compact: the edits buffer is too small (" + std::to_string(c->cur_edits_n) + " entries needed)");
    return RCX_OK;
}

extern "C" int rcx_scan_batch(rcx_ctx* c, int64_t n, const char* seq, const uint32_t* off, int32_t* counts_out,
                              int32_t* strong_threshold_out) {
    if (!c) return fail(RCX_ERR_ARG, "rcx_scan_batch: null context");
    if (!c->entries_dev) return fail(RCX_ERR_STATE, "rcx_scan_batch: no table built");
    if (n <= 0) return n == 0 ? RCX_OK : fail(RCX_ERR_ARG, "rcx_scan_batch: n < 0");
    if (!seq || !off || !counts_out || !strong_threshold_out) return fail(RCX_ERR_ARG, "rcx_scan_batch: null pointer");
    int32_t max_len = 0;
    int rc = scan_lengths(off, n, &max_len);
    if (rc) return rc;
    if (max_len < 1) max_len = 1;
    const uint64_t total = off[n];
    CUDA_TRY(cudaSetDevice(c->device));
    rc = begin_call(c, (uint64_t)n, total);
    if (rc) return rc;
    RcxSlot& sl = c->slot[0];
    CUDA_TRY(sl.d_seq.ensure(total + 96));
    CUDA_TRY(sl.d_off.ensure((size_t)n + 1));
    CUDA_TRY(cudaEventRecord(sl.ev[0], sl.stream));
    CUDA_TRY(cudaMemcpyAsync(sl.d_seq.p, seq, total, cudaMemcpyHostToDevice, sl.stream));
    CUDA_TRY(cudaMemcpyAsync(sl.d_off.p, off, ((size_t)n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, sl.stream));
    CUDA_TRY(sl.d_counts.ensure(total + 32));
    CUDA_TRY(cudaMemsetAsync(sl.d_counts.p, 0, (total + 1) * sizeof(int32_t), sl.stream));
    rc = launch_pipeline(c, sl, RCX_SINGLE, n, sl.d_seq.p, nullptr, sl.d_off.p, 0, total, max_len, nullptr, nullptr, nullptr, nullptr, false);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(sl.ev[4], sl.stream));
    CUDA_TRY(cudaMemcpyAsync(counts_out, sl.d_counts.p, total * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream));
    CUDA_TRY(cudaMemcpyAsync(strong_threshold_out, sl.d_thr.p, (size_t)n * 4, cudaMemcpyDeviceToHost, sl.stream));
    rc = finish_enqueue(sl);
    if (rc) return rc;
    return harvest(c, sl);
}

extern "C" int rcx_get_stats(rcx_ctx* c, rcx_stats* out) {
    if (!c || !out) return fail(RCX_ERR_ARG, "rcx_get_stats: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    for (auto& sl : c->slot) {
        int rc = harvest(c, sl);
        if (rc) return rc;
    }
    *out = c->stats;
    return RCX_OK;
}

// Measured random-access read ceiling of this GPU over the built table: GB/s of `sector_bytes`-sized (32 or 64)
// independent random reads (SURVEY §8d roofline denominator).
extern "C" int rcx_measure_random_read(rcx_ctx* c, int sector_bytes, double* gbps) {
    if (!c || !gbps) return fail(RCX_ERR_ARG, "rcx_measure_random_read: null argument");
    if (!c->entries_dev) return fail(RCX_ERR_STATE, "rcx_measure_random_read: no table built");
    if (sector_bytes != 32 && sector_bytes != 64) return fail(RCX_ERR_ARG, "rcx_measure_random_read: sector_bytes must be 32 or 64");
    CUDA_TRY(cudaSetDevice(c->device));
    RcxSlot& sl = c->slot[0];
    int rc = harvest(c, sl);
    if (rc) return rc;
    const int grid = c->sm_count * 8;
    const unsigned long long iters = 256;
    unsigned long long* sink = sl.d_counters + RCX_C_COUNT - 1;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
        CUDA_TRY(cudaEventRecord(sl.ev[0], sl.stream));
        CUDA_TRY(rcx_launch_random_read(table_of(c), sector_bytes, grid, iters, sink, sl.stream));
        CUDA_TRY(cudaEventRecord(sl.ev[1], sl.stream));
        CUDA_TRY(cudaStreamSynchronize(sl.stream));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, sl.ev[0], sl.ev[1]));
        if (rep > 0 && ms < best) best = ms;
    }
    *gbps = (double)grid * 256.0 * (double)iters * 8.0 * sector_bytes / (best * 1e-3) / 1e9;
    return RCX_OK;
}

// raw device counters of the last call on the first slot (debug / profiling builds: search breakdown and per-phase cycles)
extern "C" int rcx_debug_counters(rcx_ctx* c, uint64_t* out, int n) {
    if (!c || !out) return fail(RCX_ERR_ARG, "rcx_debug_counters: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    RcxSlot& sl = c->slot[0];
    CUDA_TRY(cudaStreamSynchronize(sl.stream));
    unsigned long long h[RCX_C_COUNT];
    CUDA_TRY(cudaMemcpy(h, sl.d_counters, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) out[i] = i < RCX_C_COUNT ? h[i] : 0;
    return RCX_OK;
}


// ------------------------------------------------------------------------------------------------ ERROR_RATE (main.cpp:310-358)
extern "C" int rcx_estimate_error_rate(rcx_ctx* c, const uint64_t* fwd_codes, uint64_t n, double wk, double* out) {
    if (!c || !out) return fail(RCX_ERR_ARG, "rcx_estimate_error_rate: null argument");
    if (!c->entries_dev) return fail(RCX_ERR_STATE, "rcx_estimate_error_rate: no table built");
    const uint64_t cap = 100000;
    std::vector<double> rates;
    rates.reserve(cap + 1);
    CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t chunk = 1ull << 20;
    std::vector<int32_t> cnt(4 * std::min<uint64_t>(chunk, std::max<uint64_t>(n, 1)));
    unsigned long long* d_codes = nullptr;
    int32_t* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_codes, cnt.size() / 4 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&d_out, cnt.size() * sizeof(int32_t)));
    for (uint64_t s = 0; s < n && rates.size() < cap; s += chunk) {
        const uint64_t len = std::min<uint64_t>(chunk, n - s);
        // the 4 last-base variants of a record (ShiftRight(1) + Append(b), main.cpp:327-331) share one context: one lookup
        CUDA_TRY(cudaMemcpyAsync(d_codes, fwd_codes + s, len * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(rcx_launch_table_lookup4(table_of(c), d_codes, d_out, len, c->stream));
        CUDA_TRY(cudaMemcpyAsync(cnt.data(), d_out, len * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (uint64_t i = 0; i < len && rates.size() < cap; ++i) {
            int best = 0, second = 0;
            for (int b = 0; b < 4; ++b) {
                const int v = cnt[4 * i + b];
                if (v > best) { second = best; best = v; }
                else if (v > second) second = v;
            }
            if (best < 1000) continue;
            rates.push_back((double)second / (double)best);
        }
    }
    cudaFree(d_codes);
    cudaFree(d_out);
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
