/* rcx.h — C ABI of the B200-native Rcorrector stage-3 hot path (librcx.so).
 *
 * The reference has no plugin/FFI surface: its hot path is the pthread fan-out
 * inside one binary (main.cpp:476-483 creating/joining ErrorCorrection_Thread,
 * ErrorCorrection.cpp:73-136).  These entry points are what a C-ABI replacement
 * of that region binds; each cites the reference code it replaces.  Plain
 * pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Conventions: every function returns 0 on success and a nonzero rcx_status on
 * failure (the caller prints rcx_last_error() and exit(1)s, like every fatal
 * path of the reference: main.cpp:193-197,465-466).  There is NO CPU fallback:
 * without a CUDA device rcx_create() fails with RCX_ERR_CUDA.
 */
#ifndef RCX_H
#define RCX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCX_ABI_VERSION 1
#define RCX_MAX_READ_LENGTH 1023 /* utils.h:7 MAX_READ_LENGTH 1024 incl. NUL */
#define RCX_MAX_K 32             /* run_rcorrector.pl:225-228 */

typedef enum {
    RCX_OK = 0,
    RCX_ERR_CUDA = 1,     /* no device / CUDA runtime failure */
    RCX_ERR_ARG = 2,      /* bad argument (k, lengths, null pointers, odd n in interleaved mode) */
    RCX_ERR_STATE = 3,    /* table or params not set */
    RCX_ERR_NOMEM = 4
} rcx_status;

/* how mates are laid out in one batch (ErrorCorrection.cpp:96-106) */
typedef enum {
    RCX_SINGLE = 0,      /* -r: no pair threshold (pairStrongTrustThreshold = -1) */
    RCX_INTERLEAVED = 1  /* -i, and -p after the host interleaves the two files: mates are records 2i, 2i+1 */
} rcx_mode;

/* The four live scalars the hot path reads (globals of main.cpp:17-30,
 * consumed via extern at ErrorCorrection.cpp:9-17). */
typedef struct {
    int32_t k;             /* kmerLength, 1..32 (-k, default 23) */
    int32_t max_fix_per_k; /* MAX_FIX_PER_K (-maxcorK, default 4) */
    double error_rate;     /* ERROR_RATE (main.cpp:310-358) */
    int32_t bad_qual;      /* badQualityThreshold as a char value (main.cpp:88-128); 0 for FASTA */
} rcx_params;

/* counters of the last rcx_correct_batch* call (device-side, for bench.py / profiles) */
typedef struct {
    uint64_t reads;
    uint64_t bases;
    uint64_t scan_probes;    /* 32-byte table sector requests of the scan kernel (one answers two adjacent k-windows) */
    uint64_t search_probes;  /* requests of the correction search (one answers the four extensions of a node / jump step) */
    uint64_t rescan_probes;  /* requests of the l/m/h re-scan (one answers two adjacent windows that cover a changed base) */
    uint64_t search_nodes;   /* SearchPaths_* node visits */
    uint64_t kernel_launches;/* CUDA kernels launched by the call (5 per chunk of a host batch) */
    float ms_scan, ms_correct, ms_total, ms_h2d, ms_d2h; /* CUDA-event times on the context's stream; ms_correct = triage..finalize */
    float ms_triage, ms_prep, ms_dfs, ms_finalize;       /* per-kernel split of ms_correct (both passes summed) */
    uint64_t work_items;     /* reads that needed the search (not all-trusted, not rejected) */
    uint64_t second_pass;    /* of those, reads that took the second pass of the A.4 loop */
    uint64_t heavy_items;    /* of the work items, those filed as likely long searches (searched first) */
    uint64_t vheavy_items;   /* of those, the ones with at least three quarters of their k-windows untrusted (a quarter of the work items being so
                                switches the NEXT batch's search to warp batches over them) */
    uint64_t rescan_windows; /* k-windows of the l/m/h re-scan (GetKmerInformation looks every one of them up) */
    uint64_t unsupported;    /* must be 0: a read that would need a third A.4 pass, a count >= 2^29, an edit-row overflow */
} rcx_stats;

typedef struct rcx_ctx rcx_ctx;

int rcx_abi_version(void);
/* human-readable message of the last failure on this thread */
const char* rcx_last_error(void);

/* one context = one GPU (one process per GPU, or one host thread per context) */
int rcx_create(int device, rcx_ctx** out);
void rcx_destroy(rcx_ctx* ctx);

/* ---- k-mer count table: replaces Store (Store.h:17-88) and the Put loop of
 * main.cpp:295-307.  `fwd_codes[i]` is the 2-bit code of record i's k-mer as
 * KmerCode::Append builds it (KmerCode.cpp:7-21: A0 C1 G2 T3, first base in the
 * highest bits); the library canonicalises (KmerCode.h:58-71), skips
 * count <= 1 (main.cpp:299) and keeps the LAST record of duplicate canonical
 * keys (Store.h:55).  src_on_device != 0: the arrays are device pointers on
 * this context's GPU (e.g. the landing buffers of an NCCL broadcast). */
int rcx_table_build(rcx_ctx* ctx, int k, const uint64_t* fwd_codes, const int32_t* counts, uint64_t n,
                    int src_on_device);
/* The same table on every context of `ctxs` (several GPUs of one host; the Put loop of main.cpp:295-307 once per GPU):
 * the host arrays cross PCIe once, to the first context's GPU, and go on device to device (peer copies over NVLink /
 * NVSwitch, binomial tree) before every GPU inserts locally.  Host pointers. */
int rcx_table_build_multi(rcx_ctx** ctxs, int nctx, int k, const uint64_t* fwd_codes, const int32_t* counts, uint64_t n);
/* Store::GetCount (Store.h:59-66) for n valid forward codes; host pointers. */
int rcx_table_lookup(rcx_ctx* ctx, const uint64_t* fwd_codes, int32_t* counts_out, uint64_t n);
/* number of distinct canonical k-mers stored, 32-byte sectors allocated (one (k-1)-mer context each), bytes of HBM */
int rcx_table_info(rcx_ctx* ctx, uint64_t* entries, uint64_t* slots, uint64_t* bytes);

/* ---- the k-mer counting stages of run_rcorrector.pl (:262-281: `jellyfish bc`, `jellyfish count -C --bc`,
 * `jellyfish dump -L 2`) on the GPU.  begin() opens an empty counting table; add() counts the canonical k-mers of a batch of
 * reads (same layout as rcx_correct_batch: host pointers, every record of every input file, mates included; windows that
 * hold a non-ACGT letter are skipped); finish() keeps the k-mers seen at least min_count times (2 = `dump -L 2`) and
 * builds the count table from them on the device -- no text dump in between.  The kept (canonical code, count) records can
 * be read back with export() (the order is the counting table's slot order: deterministic, but not Jellyfish's), e.g. for
 * rcx_estimate_error_rate.  capacity_hint = expected distinct k-mers (0: start small; the table grows as needed). */
int rcx_count_kmers_begin(rcx_ctx* ctx, int k, uint64_t capacity_hint);
int rcx_count_kmers_add(rcx_ctx* ctx, int64_t n, const char* seq, const uint32_t* off);
int rcx_count_kmers_finish(rcx_ctx* ctx, int32_t min_count, uint64_t* kept, uint64_t* distinct, uint64_t* instances);
int rcx_count_kmers_export(rcx_ctx* ctx, uint64_t* codes_out, int32_t* counts_out, uint64_t n);

int rcx_set_params(rcx_ctx* ctx, const rcx_params* p);

/* ---- the hot path: ErrorCorrection_Thread run to completion over one batch
 * (ErrorCorrection.cpp:73-136; the region main.cpp:476-483 brackets).
 * Reads are concatenated without separators: record r = seq[off[r] .. off[r+1]),
 * same offsets for qual; off has n+1 entries; every length <= RCX_MAX_READ_LENGTH.
 * Outputs per record: corrected bases (seq_out may alias seq_in), `correction`
 * (-1 unfixable / 0 / n bases changed), and the l,m,h k-mer counts of
 * GetKmerInformation (ErrorCorrection.cpp:1567-1602).  Host pointers; the call
 * copies host->device, runs the kernels, copies back and returns when done. */
int rcx_correct_batch(rcx_ctx* ctx, int mode, int64_t n, const char* seq_in, const char* qual,
                      const uint32_t* off, char* seq_out, int32_t* correction, int32_t* l, int32_t* m,
                      int32_t* h);
/* The same batch through a COMPACT host interface, for callers that keep their own copy of the reads (the drop-in binary's
 * reader, bench.py's e2e arm): PCIe is what a multi-GPU host runs out of first, so only what the path consumes crosses it.
 *  - quality: ErrorCorrection reads it only as "qual[i] > badQualityThreshold" (the three post-filters,
 *    ErrorCorrection.cpp:1314-1316,1407-1455) and "qual[0] != 0" (has quality at all, :1314).  qual_bits holds one bit per
 *    base of the batch: bit (i & 7) of byte i >> 3 = (signed char)qual[i] > bad_qual for batch base i (off[] indexes bases
 *    as in rcx_correct_batch); has_qual = 0 for FASTA input (qual_bits may then be null).  The threshold folded into the
 *    bits must be the bad_qual given to rcx_set_params.
 *  - result: correction / l / m / h per record as above, and instead of the corrected bases the applied substitutions:
 *    record r with correction[r] > 0 owns correction[r] consecutive entries of edits[], after those of the records before
 *    it; entry = position in the read | base << 10 (A0 C1 G2 T3), in no particular order inside a record.  *n_edits = total
 *    entries = sum of the positive correction[r]; RCX_ERR_ARG (with *n_edits set) if edits_cap is too small.
 * Per 100-bp read this moves ~117 bytes up and ~18 down instead of 204 up and 116 down. */
int rcx_correct_batch_compact(rcx_ctx* ctx, int mode, int64_t n, const char* seq, const uint8_t* qual_bits, int has_qual,
                              const uint32_t* off, int32_t* correction, int32_t* l, int32_t* m, int32_t* h, uint16_t* edits,
                              uint64_t edits_cap, uint64_t* n_edits);
/* Same as rcx_correct_batch with DEVICE pointers on the context's GPU; `seq` is corrected in place
 * (as the reference does with _Read::seq).  total_bases = off[n] and max_len =
 * the longest read, both known to the caller that built the batch.
 * Asynchronous on the context's stream: call rcx_sync() before reading results. */
int rcx_correct_batch_device(rcx_ctx* ctx, int mode, int64_t n, char* seq, const char* qual,
                             const uint32_t* off, uint64_t total_bases, int32_t max_len,
                             int32_t* correction, int32_t* l, int32_t* m, int32_t* h);
int rcx_sync(rcx_ctx* ctx);

/* stage output for parity tests: counts[] of the scan (ErrorCorrection.cpp:716-723),
 * written at counts_out[off[r] + j], and GetStrongTrustedThreshold per record
 * (ErrorCorrection.cpp:1482-1565).  Host pointers; counts_out has off[n] entries. */
int rcx_scan_batch(rcx_ctx* ctx, int64_t n, const char* seq, const uint32_t* off, int32_t* counts_out,
                   int32_t* strong_threshold_out);

int rcx_get_stats(rcx_ctx* ctx, rcx_stats* out);
/* roofline denominator: measured GB/s of independent random `sector_bytes` (32 or 64) reads over the built table */
int rcx_measure_random_read(rcx_ctx* ctx, int sector_bytes, double* gbps);
/* raw device counters of the last call (profiling builds add a search breakdown and per-phase cycle sums) */
int rcx_debug_counters(rcx_ctx* ctx, uint64_t* out, int n);
/* raw CUDA stream handle (cudaStream_t) of the context, for callers that enqueue their own copies */
void* rcx_stream(rcx_ctx* ctx);
/* pinned host memory for batch buffers (cudaHostAlloc / cudaFreeHost) */
void* rcx_host_alloc(uint64_t bytes);
void rcx_host_free(void* p);

/* ---- pre-pass scalars (main.cpp:310-358): ERROR_RATE from the
 * first `n` dump records in file order, using table probes on the GPU for the
 * 4 last-base variants of each record.  Host pointers. */
int rcx_estimate_error_rate(rcx_ctx* ctx, const uint64_t* fwd_codes, uint64_t n, double wk, double* out);

#ifdef __cplusplus
}
#endif
#endif
